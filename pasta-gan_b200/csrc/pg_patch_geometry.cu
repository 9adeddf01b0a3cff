// Patch routing, host side: the crop geometry of the reference's data loader for a whole batch in one native call.
//
// Replaces, per (sample, body part), ``get_crop`` of the reference (training/dataset.py:751-836): a few float32 operations that turn 1-4 pose
// joints into a source quadrilateral, and two ``cv2.getPerspectiveTransform`` calls (quadrilateral -> patch and back) -- plus the two 3 x 3 inversions
// ``cv2.warpPerspective`` performs on its matrix argument.  This is HOST code (no kernel): 2 x 8 x 8 eliminations per part are microseconds of scalar
// work, but 1.6 of the 2.1 ms of a batch-16 ``PatchRouter.normalize`` when evaluated with numpy.  Every value is produced by the same IEEE operations in
// the same order as OpenCV / numpy produce them (float32 for the quadrilateral, double for the systems; no fused multiply-adds: the file is compiled
// with -ffp-contract=off and uses no wider intermediates), so the matrices are bit-equal to the reference's -- tests/test_patch_routing.py holds them to
// tests/golden/warp.npz.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "pasta_b200.h"
#include "pg_common.cuh"

namespace pg {
namespace geom {

struct P2 { float x, y; };

constexpr int kJoints = 18, kParts = 10;
// joint indices in the reference's OpenPose-18 order (training/dataset.py:859-861)
enum { cnose = 0, cneck, rshoulder, relbow, rwrist, lshoulder, lelbow, lwrist, rhip, rknee, rankle, lhip, lknee, lankle };

struct Part { int n; int j[4]; int nf; int f[3]; };                 // joints of the part; joints of its fall-back (nf = 0: none)
static const Part kPartTable[kParts] = {                            // training/dataset.py:847-857, fall-backs of :756-776
    {4, {lshoulder, lhip, rhip, rshoulder}, 0, {0, 0, 0}},
    {3, {lshoulder, rshoulder, cnose, 0}, 3, {lshoulder, rshoulder, rshoulder}},
    {2, {lshoulder, lelbow, 0, 0}, 0, {0, 0, 0}},
    {2, {lelbow, lwrist, 0, 0}, 0, {0, 0, 0}},
    {2, {rshoulder, relbow, 0, 0}, 0, {0, 0, 0}},
    {2, {relbow, rwrist, 0, 0}, 0, {0, 0, 0}},
    {2, {lhip, lknee, 0, 0}, 1, {lhip, 0, 0}},
    {2, {lknee, lankle, 0, 0}, 0, {0, 0, 0}},
    {2, {rhip, rknee, 0, 0}, 1, {rhip, 0, 0}},
    {2, {rknee, rankle, 0, 0}, 0, {0, 0, 0}},
};

static inline bool confident(const double* kp, const int* j, int n) {          // valid_joints (:748-749)
    for (int i = 0; i < n; i++)
        if (!(kp[j[i] * 3 + 2] >= 0.1)) return false;
    return true;
}

static inline P2 joint(const double* kp, int j) {                              // np.float32(joints[...][:, :2]); x += 32 (the 192 -> 256 padding, :780)
    P2 p = {(float)kp[j * 3 + 0], (float)kp[j * 3 + 1]};
    p.x = p.x + 32.0f;
    return p;
}

// quadrilateral around the segment p0 -> p1, half_width x |segment| to either side (:822-829): [p0 + off, p0 - off, p1 - off, p1 + off]
static inline void segment_box(P2 p0, P2 p1, float half_width, P2 q[4]) {
    const float sx = p1.x - p0.x, sy = p1.y - p0.y;
    const float ox = half_width * (-sy), oy = half_width * sx;
    q[0] = {p0.x + ox, p0.y + oy};
    q[1] = {p0.x - ox, p0.y - oy};
    q[2] = {p1.x - ox, p1.y - oy};
    q[3] = {p1.x + ox, p1.y + oy};
}

// source quadrilateral of one part, false when its joints are not all confident even after the fall-back
static bool part_quadrilateral(const double* kp, int part, int o_h, double ar, P2 q[4]) {
    const Part& pt = kPartTable[part];
    const float hw = (float)(ar / 2.0);
    if (confident(kp, pt.j, pt.n)) {
        if (pt.n == 4) {
            for (int i = 0; i < 4; i++) q[i] = joint(kp, pt.j[i]);
        } else if (pt.n == 2) {
            segment_box(joint(kp, pt.j[0]), joint(kp, pt.j[1]), hw, q);
        } else {                                                               // head box: from twice the neck-to-nose vector down to the neck (:803-817)
            const P2 a = joint(kp, pt.j[0]), b = joint(kp, pt.j[1]), nose = joint(kp, pt.j[2]);
            const P2 neck = {0.5f * (a.x + b.x), 0.5f * (a.y + b.y)};
            const P2 top = {neck.x + 2.0f * (nose.x - neck.x), neck.y + 2.0f * (nose.y - neck.y)};
            P2 t[4];
            segment_box(top, neck, 0.5f, t);
            q[0] = t[1]; q[1] = t[2]; q[2] = t[3]; q[3] = t[0];
        }
        return true;
    }
    if (pt.nf == 0 || !confident(kp, pt.f, pt.nf)) return false;
    if (pt.nf == 1) {                                                          // hip without knee: a vertical from the hip to the bottom edge (:782-786)
        const P2 hip = joint(kp, pt.f[0]);
        const P2 foot = {hip.x, (float)(o_h - 1)};
        segment_box(hip, foot, hw, q);
    } else {                                                                   // shoulders without nose: a square above the shoulder line (:792-801)
        const P2 a = joint(kp, pt.f[0]), b = joint(kp, pt.f[1]);
        float nx = -(b.y - a.y), ny = b.x - a.x;
        if (ny > 0.0f) { nx = -nx; ny = -ny; }
        q[0] = {a.x + nx, a.y + ny};
        q[1] = a;
        q[2] = b;
        q[3] = {b.x + nx, b.y + ny};
    }
    return true;
}

// cv2.getPerspectiveTransform(src, dst): the 8 x 8 system of imgwarp.cpp solved as cv::solve(DECOMP_LU) does (LUImpl: row pivoting on the first largest
// magnitude, row updates a += alpha * pivot_row with alpha = a[j][i] * (-1 / a[i][i]), back substitution).  Singular: zeros, m[8] = 1.
static void perspective_transform(const P2 src[4], const P2 dst[4], double m[9]) {
    double A[8][8], b[8];
    std::memset(A, 0, sizeof(A));
    for (int i = 0; i < 4; i++) {
        A[i][0] = A[i + 4][3] = src[i].x;
        A[i][1] = A[i + 4][4] = src[i].y;
        A[i][2] = A[i + 4][5] = 1.0;
        A[i][6] = (double)(float)(-src[i].x * dst[i].x);                       // float product, then widened (as OpenCV and numpy do)
        A[i][7] = (double)(float)(-src[i].y * dst[i].x);
        A[i + 4][6] = (double)(float)(-src[i].x * dst[i].y);
        A[i + 4][7] = (double)(float)(-src[i].y * dst[i].y);
        b[i] = dst[i].x;
        b[i + 4] = dst[i].y;
    }
    const double eps = DBL_EPSILON * 100;
    bool ok = true;
    for (int i = 0; i < 8 && ok; i++) {
        int k = i;
        for (int j = i + 1; j < 8; j++)
            if (std::fabs(A[j][i]) > std::fabs(A[k][i])) k = j;
        if (std::fabs(A[k][i]) < eps) { ok = false; break; }
        if (k != i) {
            for (int c = 0; c < 8; c++) { const double t = A[i][c]; A[i][c] = A[k][c]; A[k][c] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1.0 / A[i][i];
        for (int j = i + 1; j < 8; j++) {
            const double alpha = A[j][i] * d;
            for (int c = i + 1; c < 8; c++) A[j][c] = A[j][c] + alpha * A[i][c];
            b[j] = b[j] + alpha * b[i];
        }
    }
    if (ok) {
        for (int i = 7; i >= 0; i--) {
            double s = b[i];
            for (int c = i + 1; c < 8; c++) s = s - A[i][c] * b[c];
            b[i] = s / A[i][i];
        }
        for (int i = 0; i < 8; i++) m[i] = b[i];
    } else {
        for (int i = 0; i < 8; i++) m[i] = 0.0;
    }
    m[8] = 1.0;
}

// cv::invert on a 3 x 3 CV_64F matrix: adjugate times 1 / det, zeros when det == 0 (what warpPerspective applies to its matrix argument)
static void invert3x3(const double S[9], double t[9]) {
    const double c00 = S[4] * S[8] - S[5] * S[7];
    const double c01 = S[3] * S[8] - S[5] * S[6];
    const double c02 = S[3] * S[7] - S[4] * S[6];
    const double det = S[0] * c00 - S[1] * c01 + S[2] * c02;
    const double d = det != 0.0 ? 1.0 / det : 0.0;
    t[0] = c00 * d;
    t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    t[6] = c02 * d;
    t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
}

}  // namespace geom
}  // namespace pg

extern "C" int pg_patch_crop_transforms(const double* keypoints, int32_t B, int32_t h, int32_t w, int32_t o_h, double ar,
                                        double* M, double* M_inv, double* to_patch, double* to_image, uint8_t* valid) {
    using namespace pg;
    using namespace pg::geom;
    if (B == 0) return PG_OK;
    PG_REQUIRE(keypoints && M && M_inv && valid, "pg_patch_crop_transforms: null pointer");
    PG_REQUIRE(B > 0 && h > 0 && w > 0 && o_h > 0, "pg_patch_crop_transforms: bad shape (B = %d, patch %d x %d, image height %d)", B, h, w, o_h);
    const P2 dst[4] = {{0.0f, 0.0f}, {0.0f, (float)h}, {(float)w, (float)h}, {(float)w, 0.0f}};        // wh * unit square (:831-832)
    for (int b = 0; b < B; b++) {
        const double* kp = keypoints + (size_t)b * kJoints * 3;
        for (int p = 0; p < kParts; p++) {
            const size_t o = ((size_t)b * kParts + p) * 9;
            P2 q[4];
            const bool ok = part_quadrilateral(kp, p, o_h, ar, q);
            valid[(size_t)b * kParts + p] = ok ? 1 : 0;
            if (ok) {
                perspective_transform(q, dst, M + o);
                perspective_transform(dst, q, M_inv + o);
            } else {
                for (int i = 0; i < 9; i++) M[o + i] = M_inv[o + i] = 0.0;
            }
            if (to_patch) invert3x3(M + o, to_patch + o);
            if (to_image) invert3x3(M_inv + o, to_image + o);
        }
    }
    return PG_OK;
}
