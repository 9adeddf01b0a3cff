// Patch routing on the device: the perspective warps of the reference's data loader, training/dataset.py:838-927
// (UvitonDatasetFull*.normalize), which calls cv2.warpPerspective 28 + 28 times per sample on the CPU.
//
//   pg_warp_perspective_u8   a table of independent warps (one job = one cv2.warpPerspective call on a uint8 image of <= 4 channels), one launch;
//                            strided source / destination so that the ten rectified patches are written straight into the channel-concatenated
//                            [h, w, 30] tensor the reference builds with np.concatenate (dataset.py:920-923)
//   pg_patch_denorm_u8       the "denormalised" composite: for each output pixel walk the parts in order, warp the rectified patch and its garment
//                            mask back (BORDER_CONSTANT), and keep the patch where the mask came back as exactly 255 (dataset.py:882-886, :892-897);
//                            also emits the per-part 0 / 1 masks the loader keeps for the arms (dataset.py:903-907)
//
// Arithmetic is OpenCV's fixed-point INTER_LINEAR path, restated so that results are bit-identical to it: destination-to-source coordinates in double
// with the rounding sequence of WarpPerspectiveInvoker (64-column blocks: X0 at the block's first column, then X0 + M0*x1; no fused multiply-add),
// rounded half-to-even to 1/32 pixel, int16-saturated integer parts, 2^15-scaled int16 bilinear weights (the (0,0) entry is {32767, 0, 0, 1}),
// (sum + 2^14) >> 15.  Byte work, HBM / latency bound: 0.8 MB in and 1.1 MB out per sample; one thread per destination pixel, coalesced stores.
#include "pg_common.cuh"

namespace pg {

struct WarpCoord { int sx, sy, fx, fy; };

// Source coordinate of destination pixel (x, y): WarpPerspectiveInvoker's scalar loop (imgwarp.cpp), block width bw0.
__device__ __forceinline__ WarpCoord warp_coord(const double* __restrict__ m, int x, int y, int bw0) {
    const int xb = (x / bw0) * bw0, x1 = x - xb;
    const double dxb = (double)xb, dy = (double)y, dx1 = (double)x1;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dxb), __dmul_rn(m[1], dy)), m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dxb), __dmul_rn(m[4], dy)), m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m[6], dxb), __dmul_rn(m[7], dy)), m[8]);
    double W = __dadd_rn(W0, __dmul_rn(m[6], dx1));
    W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(m[0], dx1)), W)));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(m[3], dx1)), W)));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    WarpCoord c;
    c.sx = max(-32768, min(32767, X >> 5));
    c.sy = max(-32768, min(32767, Y >> 5));
    c.fx = X & 31; c.fy = Y & 31;
    return c;
}

// int16 bilinear weights {tl, tr, bl, br} of BilinearTab_i[fy*32 + fx]
__device__ __forceinline__ void bilinear_weights(int fx, int fy, int (&w)[4]) {
    w[0] = (32 - fy) * (32 - fx) * 32; w[1] = (32 - fy) * fx * 32; w[2] = fy * (32 - fx) * 32; w[3] = fy * fx * 32;
    if ((fx | fy) == 0) { w[0] = 32767; w[3] = 1; }               // saturate_cast<short>(32768) and OpenCV's sum correction
}

// One channel of remapBilinear.  BORDER: 0 = constant 0, 1 = replicate.  ``base`` points at channel c of pixel (0, 0).
template <int BORDER>
__device__ __forceinline__ unsigned char remap_tap4(const unsigned char* __restrict__ base, int H, int W, int row_stride, int pix_stride,
                                                    const WarpCoord& c, const int (&w)[4]) {
    int v[4];
    if ((unsigned)c.sx < (unsigned)(W - 1) && (unsigned)c.sy < (unsigned)(H - 1)) {
        const unsigned char* p = base + (size_t)c.sy * row_stride + (size_t)c.sx * pix_stride;
        v[0] = p[0]; v[1] = p[pix_stride]; v[2] = p[row_stride]; v[3] = p[row_stride + pix_stride];
    } else if (BORDER == 1) {
        const int x0 = min(max(c.sx, 0), W - 1), x1 = min(max(c.sx + 1, 0), W - 1);
        const int y0 = min(max(c.sy, 0), H - 1), y1 = min(max(c.sy + 1, 0), H - 1);
        v[0] = base[(size_t)y0 * row_stride + (size_t)x0 * pix_stride]; v[1] = base[(size_t)y0 * row_stride + (size_t)x1 * pix_stride];
        v[2] = base[(size_t)y1 * row_stride + (size_t)x0 * pix_stride]; v[3] = base[(size_t)y1 * row_stride + (size_t)x1 * pix_stride];
    } else {
        if (c.sx >= W || c.sx + 1 < 0 || c.sy >= H || c.sy + 1 < 0) return 0;
        const bool xa = c.sx >= 0, xb = c.sx + 1 < W, ya = c.sy >= 0, yb = c.sy + 1 < H;
        v[0] = (xa && ya) ? base[(size_t)c.sy * row_stride + (size_t)c.sx * pix_stride] : 0;
        v[1] = (xb && ya) ? base[(size_t)c.sy * row_stride + (size_t)(c.sx + 1) * pix_stride] : 0;
        v[2] = (xa && yb) ? base[(size_t)(c.sy + 1) * row_stride + (size_t)c.sx * pix_stride] : 0;
        v[3] = (xb && yb) ? base[(size_t)(c.sy + 1) * row_stride + (size_t)(c.sx + 1) * pix_stride] : 0;
    }
    const int acc = v[0] * w[0] + v[1] * w[1] + v[2] * w[2] + v[3] * w[3];
    return (unsigned char)min(255, max(0, (acc + (1 << 14)) >> 15));
}

__device__ __forceinline__ int warp_block_width(int dst_h, int dst_w) {        // bw0 of WarpPerspectiveInvoker (BLOCK_SZ = 32)
    const int bh0 = min(16, dst_h);
    return min(1024 / bh0, dst_w);
}

// grid = (ceil(max_dst_pixels / 256), njobs)
__global__ void __launch_bounds__(256) warp_perspective_u8_kernel(const pg_warp_job* __restrict__ jobs) {
    __shared__ pg_warp_job j;
    if (threadIdx.x < sizeof(pg_warp_job) / 8)
        reinterpret_cast<unsigned long long*>(&j)[threadIdx.x] = reinterpret_cast<const unsigned long long*>(jobs + blockIdx.y)[threadIdx.x];
    __syncthreads();
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= j.dst_h * j.dst_w) return;
    const int y = pix / j.dst_w, x = pix - y * j.dst_w;
    const WarpCoord c = warp_coord(j.m, x, y, warp_block_width(j.dst_h, j.dst_w));
    int w[4];
    bilinear_weights(c.fx, c.fy, w);
    unsigned char* d = j.dst + (size_t)y * j.dst_row_stride + (size_t)x * j.dst_pix_stride;
    for (int ch = 0; ch < j.channels; ch++)
        d[ch] = j.border == 1 ? remap_tap4<1>(j.src + ch, j.src_h, j.src_w, j.src_row_stride, j.src_pix_stride, c, w)
                              : remap_tap4<0>(j.src + ch, j.src_h, j.src_w, j.src_row_stride, j.src_pix_stride, c, w);
}

// grid = (ceil(H*W / 256), B); patches / masks [B][h][w][P*3], m [B][P][9] (destination -> patch), valid [B][P]
__global__ void __launch_bounds__(256) patch_denorm_u8_kernel(const unsigned char* __restrict__ patches, const unsigned char* __restrict__ masks,
                                                              const double* __restrict__ m, const unsigned char* __restrict__ valid,
                                                              unsigned char* __restrict__ denorm, unsigned char* __restrict__ part_masks,
                                                              int P, int h, int w, int H, int W) {
    constexpr int kMaxParts = 16;
    __shared__ double sm[kMaxParts * 9];
    __shared__ unsigned char sv[kMaxParts];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < P * 9; i += 256) sm[i] = m[(size_t)b * P * 9 + i];
    if (threadIdx.x < P) sv[threadIdx.x] = valid[(size_t)b * P + threadIdx.x];
    __syncthreads();
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    const int bw0 = warp_block_width(H, W);
    const int row_stride = w * P * 3, pix_stride = P * 3;
    const unsigned char* pb = patches + (size_t)b * h * row_stride;
    const unsigned char* mb = masks + (size_t)b * h * row_stride;
    unsigned char out[3] = {0, 0, 0};
    for (int p = 0; p < P; p++) {
        unsigned char keep = 0;
        if (sv[p]) {
            const WarpCoord c = warp_coord(sm + p * 9, x, y, bw0);
            int wt[4];
            bilinear_weights(c.fx, c.fy, wt);
            keep = remap_tap4<0>(mb + p * 3, h, w, row_stride, pix_stride, c, wt) == 255;
            if (keep) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) out[ch] = remap_tap4<0>(pb + p * 3 + ch, h, w, row_stride, pix_stride, c, wt);
            }
        }
        if (part_masks) part_masks[((size_t)b * P + p) * H * W + pix] = keep;
    }
    unsigned char* d = denorm + ((size_t)b * H * W + pix) * 3;
    d[0] = out[0]; d[1] = out[1]; d[2] = out[2];
}

}  // namespace pg

extern "C" int pg_warp_perspective_u8(const pg_warp_job* jobs_device, int32_t njobs, int32_t max_dst_pixels, void* stream) {
    using namespace pg;
    static_assert(sizeof(pg_warp_job) == 128, "pg_warp_job must stay 128 bytes (the kernel copies it as 16 x 8-byte words; the Python binding mirrors it)");
    if (njobs == 0 || max_dst_pixels == 0) return PG_OK;
    PG_REQUIRE(jobs_device != nullptr && njobs > 0 && njobs <= 65535 && max_dst_pixels > 0, "pg_warp_perspective_u8: bad job table (njobs = %d, max_dst_pixels = %d)", njobs, max_dst_pixels);
    warp_perspective_u8_kernel<<<dim3((max_dst_pixels + 255) / 256, njobs), 256, 0, (cudaStream_t)stream>>>(jobs_device);
    return launch_status("warp_perspective_u8");
}

extern "C" int pg_patch_denorm_u8(const void* patches, const void* masks, const double* m, const void* valid, void* denorm, void* part_masks,
                                  int32_t B, int32_t P, int32_t h, int32_t w, int32_t H, int32_t W, void* stream) {
    using namespace pg;
    if (B == 0) return PG_OK;
    PG_REQUIRE(patches && masks && m && valid && denorm, "pg_patch_denorm_u8: null pointer");
    PG_REQUIRE(B > 0 && B <= 65535 && P > 0 && P <= 16 && h > 1 && w > 1 && H > 0 && W > 0, "pg_patch_denorm_u8: bad shape (B = %d, P = %d, patch %d x %d, image %d x %d)", B, P, h, w, H, W);
    patch_denorm_u8_kernel<<<dim3((H * W + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)patches, (const unsigned char*)masks, m, (const unsigned char*)valid, (unsigned char*)denorm, (unsigned char*)part_masks, P, h, w, H, W);
    return launch_status("patch_denorm_u8");
}
