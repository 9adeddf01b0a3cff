/* oracle.c — plain-C restatement of the PASTA-GAN operator hot path.  TEST INFRASTRUCTURE ONLY (see oracle/ops_oracle.py).
 *
 * Scalar loops with double accumulation, written from the definitions the reference implements:
 *   orc_upfirdn2d   torch_utils/ops/upfirdn2d.py:169-208 (_upfirdn2d_ref) / upfirdn2d.cu:29-92 (the generic gather kernel)
 *   orc_bias_act    torch_utils/ops/bias_act.py:94-123 (_bias_act_ref) and the grad = 1 branch of bias_act.cu:54-142
 *   orc_get_perspective_transform / orc_warp_perspective_u8   the OpenCV calls of the data loader's patch routing (see the section at the end)
 *   orc_conv2d      what conv2d_gradfix.conv2d computes (conv2d_gradfix.py:35-38): cross-correlation, or true convolution with flip_weight = 0
 *                   (conv2d_resample.py:35-36)
 * It is a second, independent statement of the algorithm next to the torch-CPU oracle: tests/test_oracle_c.py pins it to the same golden
 * vectors (the .npz files under tests/golden, generated from the unmodified reference).  Parity status: PINNED.  Nothing in pasta-gan_b200/ links or loads it.
 * Build: make -C oracle  ->  oracle/_build/liboracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

/* y[n,c,oy,ox] = gain * sum_{i,j} k[i,j] * z[oy*downy + i, ox*downx + j],  z = pad(zero_insert(x)),  k = f flipped unless `flip`. */
int orc_upfirdn2d(const float* x, const float* f, float* y, int N, int C, int H, int W, int fh, int fw,
                  int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1, int flip, float gain) {
    const int outW = (W * upx + padx0 + padx1 - fw + downx) / downx;
    const int outH = (H * upy + pady0 + pady1 - fh + downy) / downy;
    if (outW < 1 || outH < 1) return 1;
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; n++)
        for (int c = 0; c < C; c++) {
            const float* xp = x + ((size_t)n * C + c) * H * W;
            float* yp = y + ((size_t)n * C + c) * outH * outW;
            for (int oy = 0; oy < outH; oy++)
                for (int ox = 0; ox < outW; ox++) {
                    double acc = 0.0;
                    for (int i = 0; i < fh; i++) {
                        const int uy = oy * downy + i - pady0;            /* row in the zero-inserted image */
                        if (uy < 0 || uy % upy != 0) continue;
                        const int iy = uy / upy;
                        if (iy >= H) continue;
                        for (int j = 0; j < fw; j++) {
                            const int ux = ox * downx + j - padx0;
                            if (ux < 0 || ux % upx != 0) continue;
                            const int ix = ux / upx;
                            if (ix >= W) continue;
                            const float kv = flip ? f[i * fw + j] : f[(fh - 1 - i) * fw + (fw - 1 - j)];
                            acc += (double)kv * (double)xp[(size_t)iy * W + ix];
                        }
                    }
                    yp[(size_t)oy * outW + ox] = (float)(acc * (double)gain);
                }
        }
    (void)floordiv;
    return 0;
}

static double act_fwd(int act, double t, double alpha) {
    switch (act) {
        case 1: return t;
        case 2: return t > 0 ? t : 0;
        case 3: return t > 0 ? t : t * alpha;
        case 4: return tanh(t);
        case 5: return 1.0 / (1.0 + exp(-t));
        case 6: return t >= 0 ? t : expm1(t);
        case 7: return 1.0507009873554804934193349852946 * (t >= 0 ? t : 1.6732632423543772848170429916717 * expm1(t));
        case 8: return t > 20 ? t : log1p(exp(t));
        case 9: return t / (1.0 + exp(-t));
    }
    return NAN;
}

/* grad == 0: y = clamp(act(x + b) * gain).   grad == 1: y = x * gain * act'(.) from yref (linear / relu / lrelu only), 0 where |yref| >= clamp. */
int orc_bias_act(const float* x, const float* b, const float* yref, float* y, int64_t n, int size_b, int64_t step_b,
                 int grad, int act, float alpha, float gain, float clamp) {
    if (grad == 1 && act > 3) return 2;
#pragma omp parallel for
    for (int64_t i = 0; i < n; i++) {
        const double bias = b ? (double)b[(i / step_b) % size_b] : 0.0;
        double v;
        if (grad == 0) {
            v = act_fwd(act, (double)x[i] + bias, alpha) * gain;
            if (clamp >= 0) v = v > clamp ? clamp : (v < -clamp ? -clamp : v);
        } else {
            const double yy = gain != 0 ? (double)yref[i] / gain : 0.0;
            const double d1 = act == 1 ? 1.0 : (act == 2 ? (yy > 0 ? 1.0 : 0.0) : (yy > 0 ? 1.0 : alpha));
            v = (double)x[i] * d1 * gain;
            if (clamp >= 0 && !((double)yref[i] > -clamp && (double)yref[i] < clamp)) v = 0.0;
        }
        y[i] = (float)v;
    }
    return 0;
}

/* y[n,o,oy,ox] = sum_{c,i,j} w'[o,c,i,j] * xpad[n,c,oy*stride + i, ox*stride + j];  w' = w (flip_weight != 0, cross-correlation) or w mirrored. */
int orc_conv2d(const float* x, const float* w, float* y, int N, int Cin, int H, int W, int Cout, int kh, int kw,
               int stride, int pad_y, int pad_x, int flip_weight) {
    const int outH = (H + 2 * pad_y - kh) / stride + 1, outW = (W + 2 * pad_x - kw) / stride + 1;
    if (outH < 1 || outW < 1) return 1;
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; n++)
        for (int o = 0; o < Cout; o++)
            for (int oy = 0; oy < outH; oy++)
                for (int ox = 0; ox < outW; ox++) {
                    double acc = 0.0;
                    for (int c = 0; c < Cin; c++)
                        for (int i = 0; i < kh; i++) {
                            const int iy = oy * stride + i - pad_y;
                            if (iy < 0 || iy >= H) continue;
                            for (int j = 0; j < kw; j++) {
                                const int ix = ox * stride + j - pad_x;
                                if (ix < 0 || ix >= W) continue;
                                const int wi = flip_weight ? i : kh - 1 - i, wj = flip_weight ? j : kw - 1 - j;
                                acc += (double)w[(((size_t)o * Cin + c) * kh + wi) * kw + wj] * (double)x[(((size_t)n * Cin + c) * H + iy) * W + ix];
                            }
                        }
                    y[(((size_t)n * Cout + o) * outH + oy) * outW + ox] = (float)acc;
                }
    return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Patch routing: cv2.getPerspectiveTransform and cv2.warpPerspective(uint8, INTER_LINEAR) in plain C, from OpenCV's published
 * algorithm (modules/imgproc/src/imgwarp.cpp: getPerspectiveTransform, WarpPerspectiveInvoker, remapBilinear with the fixed-point BilinearTab_i of
 * initInterTab2D; modules/core/src/lapack.cpp: invert for 3 x 3; matrix_decomp.cpp: LUImpl) -- the calls the reference's data loader makes
 * (training/dataset.py:834-835, :879-897).  A second restatement next to oracle/warp_oracle.py; tests/test_oracle_c.py pins it to
 * tests/golden/warp.npz (bytes written by the real cv2 4.13.0 and by the unmodified reference normalize).  Parity status: PINNED.
 * Compile without FMA contraction (the Makefile passes -ffp-contract=off): the double rounding sequence is part of the result. */

/* src, dst: 4 points each as float32 x0,y0,x1,y1,...; m: 9 doubles, row-major, m[8] = 1.  Singular system: zeros (m[8] = 1). */
int orc_get_perspective_transform(const float* src, const float* dst, double* m) {
    double A[8][8], b[8];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) A[i][j] = 0.0;
    for (int i = 0; i < 4; i++) {
        const float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        A[i][0] = A[i + 4][3] = sx;
        A[i][1] = A[i + 4][4] = sy;
        A[i][2] = A[i + 4][5] = 1.0;
        A[i][6] = (double)(float)(-sx * dx);                   /* float products, widened afterwards */
        A[i][7] = (double)(float)(-sy * dx);
        A[i + 4][6] = (double)(float)(-sx * dy);
        A[i + 4][7] = (double)(float)(-sy * dy);
        b[i] = dx;
        b[i + 4] = dy;
    }
    const double eps = 2.220446049250313e-16 * 100;
    int ok = 1;
    for (int i = 0; i < 8 && ok; i++) {
        int k = i;
        for (int j = i + 1; j < 8; j++) if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (fabs(A[k][i]) < eps) { ok = 0; break; }
        if (k != i) {
            for (int c = 0; c < 8; c++) { const double t = A[i][c]; A[i][c] = A[k][c]; A[k][c] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1.0 / A[i][i];
        for (int j = i + 1; j < 8; j++) {
            const double alpha = A[j][i] * d;
            for (int c = i + 1; c < 8; c++) A[j][c] += alpha * A[i][c];
            b[j] += alpha * b[i];
        }
    }
    if (ok)
        for (int i = 7; i >= 0; i--) {
            double s = b[i];
            for (int c = i + 1; c < 8; c++) s -= A[i][c] * b[c];
            b[i] = s / A[i][i];
        }
    for (int i = 0; i < 8; i++) m[i] = ok ? b[i] : 0.0;
    m[8] = 1.0;
    return 0;
}

static void orc_invert3x3(const double* S, double* t) {
    const double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[3] * S[8] - S[5] * S[6], c02 = S[3] * S[7] - S[4] * S[6];
    const double det = S[0] * c00 - S[1] * c01 + S[2] * c02;
    const double d = det != 0.0 ? 1.0 / det : 0.0;
    t[0] = c00 * d;                          t[1] = (S[2] * S[7] - S[1] * S[8]) * d;  t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d;  t[4] = (S[0] * S[8] - S[2] * S[6]) * d;  t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    t[6] = c02 * d;                          t[7] = (S[1] * S[6] - S[0] * S[7]) * d;  t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
}

/* BilinearTab_i[fy * 32 + fx] = {tl, tr, bl, br}: round(w * 2^15) as short, then the sum forced to 2^15 on the largest / smallest entry. */
static void orc_bilinear_tab(short tab[1024][4]) {
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            const float fy[2] = {1.0f - (float)i * (1.0f / 32.0f), (float)i * (1.0f / 32.0f)};
            const float fx[2] = {1.0f - (float)j * (1.0f / 32.0f), (float)j * (1.0f / 32.0f)};
            int v[4], sum = 0;
            for (int k1 = 0; k1 < 2; k1++)
                for (int k2 = 0; k2 < 2; k2++) {
                    long r = lrintf(fy[k1] * fx[k2] * 32768.0f);
                    if (r > 32767) r = 32767;
                    if (r < -32768) r = -32768;
                    v[k1 * 2 + k2] = (int)r;
                    sum += (int)r;
                }
            if (sum != 32768) {
                /* OpenCV scans the centre ksize/2 .. ksize/2+1 block -- for the 2 x 2 kernel that is entry (1,1) and whatever lies behind it, which
                 * only matters at (0,0): {32767,0,0,0} gets its missing unit on entry (1,1) */
                const int diff = sum - 32768;
                v[3] -= diff;
            }
            for (int k = 0; k < 4; k++) tab[i * 32 + j][k] = (short)v[k];
        }
}

static int orc_sat16(long long v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : (int)v); }

/* dst[h][w][C] = cv2.warpPerspective(src[H][W][C], M, (w, h), INTER_LINEAR, border (0 constant / 1 replicate), 0); M: source -> destination, 9 doubles */
int orc_warp_perspective_u8(const unsigned char* src, int H, int W, int C, const double* M, unsigned char* dst, int h, int w, int border) {
    static short tab[1024][4];
    static int tab_ready = 0;
    if (!tab_ready) { orc_bilinear_tab(tab); tab_ready = 1; }
    if (C < 1 || C > 4 || H < 1 || W < 1 || h < 1 || w < 1) return 1;
    double m[9];
    orc_invert3x3(M, m);
    const int bh0 = h < 16 ? h : 16;
    int bw0 = 1024 / bh0;
    if (bw0 > w) bw0 = w;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int xb = (x / bw0) * bw0, x1 = x - xb;
            const double X0 = (m[0] * xb + m[1] * y) + m[2], Y0 = (m[3] * xb + m[4] * y) + m[5], W0 = (m[6] * xb + m[7] * y) + m[8];
            double Wd = W0 + m[6] * x1;
            Wd = Wd != 0.0 ? 32.0 / Wd : 0.0;
            double fX = (X0 + m[0] * x1) * Wd, fY = (Y0 + m[3] * x1) * Wd;
            if (!(fX > -2147483648.0)) fX = -2147483648.0;
            if (!(fX < 2147483647.0)) fX = 2147483647.0;
            if (!(fY > -2147483648.0)) fY = -2147483648.0;
            if (!(fY < 2147483647.0)) fY = 2147483647.0;
            const long long X = llrint(fX), Y = llrint(fY);    /* round half to even (default rounding mode), as saturate_cast<int>(double) */
            const int sx = orc_sat16(X >> 5), sy = orc_sat16(Y >> 5);
            const short* wt = tab[(int)(Y & 31) * 32 + (int)(X & 31)];
            unsigned char* d = dst + ((size_t)y * w + x) * C;
            if (border == 0 && (sx >= W || sx + 1 < 0 || sy >= H || sy + 1 < 0)) {
                for (int c = 0; c < C; c++) d[c] = 0;
                continue;
            }
            for (int c = 0; c < C; c++) {
                int acc = 0;
                for (int t = 0; t < 4; t++) {
                    int yy = sy + (t >> 1), xx = sx + (t & 1), v;
                    if (border == 1) {
                        yy = yy < 0 ? 0 : (yy > H - 1 ? H - 1 : yy);
                        xx = xx < 0 ? 0 : (xx > W - 1 ? W - 1 : xx);
                        v = src[((size_t)yy * W + xx) * C + c];
                    } else {
                        v = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? src[((size_t)yy * W + xx) * C + c] : 0;
                    }
                    acc += v * wt[t];
                }
                const int o = (acc + (1 << 14)) >> 15;
                d[c] = (unsigned char)(o < 0 ? 0 : (o > 255 ? 255 : o));
            }
        }
    return 0;
}
