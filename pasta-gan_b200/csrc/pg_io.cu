// Input / output pipeline on the device, either side of the generator (reference test.py:105-115 and :131-135):
//   pg_u8_normalize      uint8 loader tensors -> float32 generator inputs, x / 127.5 - 1 (images) or x (masks), several tensors per launch, with a
//                        destination batch stride so that `pose || retain` (torch.cat, test.py:115) is written in place
//   pg_image_to_u8_bgr   float32 NCHW generator output -> uint8 HWC, BGR, cropped to the photo: uint8(clip((v + 1) * 127.5, 0, 255))
// Both are pure streaming (1 + 4 and 4 + 1 bytes per element); the point is the PCIe side: 17 MB instead of 82 MB in, 2.4 MB instead of 25 MB out
// per batch of 16 at 256 x 192.  Arithmetic is written so that results are bit-identical to the reference's torch / numpy expressions.
#include "pg_common.cuh"

namespace pg {

constexpr int kIoMaxJobs = 8;
struct U8Jobs {
    const unsigned char* src[kIoMaxJobs]; float* dst[kIoMaxJobs];
    long long rows[kIoMaxJobs], row_len[kIoMaxJobs], src_stride[kIoMaxJobs], dst_stride[kIoMaxJobs];
    int normalize[kIoMaxJobs], vec[kIoMaxJobs];
    int njobs;
};

__device__ __forceinline__ float u8_to_float(unsigned int b, int normalize) {
    // torch on CUDA evaluates `x.to(float32) / 127.5 - 1` (test.py:105) as x * (1.0f / 127.5f) - 1: division by a host scalar becomes a multiply by
    // its float reciprocal (ATen BinaryDivTrueKernel).  Same two roundings here, no FMA contraction.
    const float v = (float)b;
    return normalize ? __fsub_rn(__fmul_rn(v, 1.0f / 127.5f), 1.0f) : v;
}

__global__ void __launch_bounds__(256) u8_normalize_kernel(const __grid_constant__ U8Jobs j) {
    const int job = blockIdx.y;
    const unsigned char* src = j.src[job]; float* dst = j.dst[job];
    const long long len = j.row_len[job], rows = j.rows[job];
    const int nz = j.normalize[job];
    if (j.vec[job]) {
        const long long per_row = len >> 4, total = rows * per_row;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / per_row, c = (i - r * per_row) << 4;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + r * j.src_stride[job] + c));
            float4* o = reinterpret_cast<float4*>(dst + r * j.dst_stride[job] + c);
            const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; k++)
                o[k] = make_float4(u8_to_float(w[k] & 0xffu, nz), u8_to_float((w[k] >> 8) & 0xffu, nz), u8_to_float((w[k] >> 16) & 0xffu, nz), u8_to_float(w[k] >> 24, nz));
        }
    } else {
        const long long total = rows * len;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const long long r = i / len, c = i - r * len;
            dst[r * j.dst_stride[job] + c] = u8_to_float(__ldg(src + r * j.src_stride[job] + c), nz);
        }
    }
}

__global__ void __launch_bounds__(256) image_to_u8_bgr_kernel(const float* __restrict__ img, unsigned char* __restrict__ out, int N, int H, int W, int x0, int x1) {
    const int wc = x1 - x0;
    const long long total = (long long)N * H * wc;
    const long long plane = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % wc); const long long r = i / wc; const int y = (int)(r % H); const int n = (int)(r / H);
        const float* p = img + (long long)n * 3 * plane + (long long)y * W + x0 + x;
        unsigned char* o = out + i * 3;
#pragma unroll
        for (int k = 0; k < 3; k++) {                                  // output channel k = input channel 2 - k (RGB -> BGR, test.py:133)
            float v = __fmul_rn(__fadd_rn(__ldg(p + (2 - k) * plane), 1.0f), 127.5f);
            v = fminf(fmaxf(v, 0.f), 255.f);                           // np.clip; NaN -> 0 like fmaxf
            o[k] = (unsigned char)(int)v;                              // astype(np.uint8) truncates
        }
    }
}

}  // namespace pg

extern "C" int pg_u8_normalize(const void* const* src, void* const* dst, const int64_t* rows, const int64_t* row_len, const int64_t* src_stride,
                               const int64_t* dst_stride, const int32_t* normalize, int32_t njobs, void* stream) {
    using namespace pg;
    PG_REQUIRE(njobs >= 0 && njobs <= kIoMaxJobs, "u8_normalize: at most %d tensors per call (got %d)", kIoMaxJobs, njobs);
    if (njobs == 0) return PG_OK;
    PG_REQUIRE(src && dst && rows && row_len && src_stride && dst_stride && normalize, "u8_normalize: job arrays must be host pointers");
    U8Jobs j; j.njobs = njobs;
    long long most = 0;
    for (int i = 0; i < kIoMaxJobs; i++) {
        const int k = i < njobs ? i : 0;
        PG_REQUIRE(src[k] && dst[k] && rows[k] >= 0 && row_len[k] >= 1 && src_stride[k] >= row_len[k] && dst_stride[k] >= row_len[k], "u8_normalize: bad job %d", k);
        j.src[i] = (const unsigned char*)src[k]; j.dst[i] = (float*)dst[k];
        j.rows[i] = rows[k]; j.row_len[i] = row_len[k]; j.src_stride[i] = src_stride[k]; j.dst_stride[i] = dst_stride[k]; j.normalize[i] = normalize[k];
        j.vec[i] = row_len[k] % 16 == 0 && src_stride[k] % 16 == 0 && dst_stride[k] % 4 == 0 && aligned16(src[k]) && aligned16(dst[k]);
        const long long work = rows[k] * (j.vec[i] ? row_len[k] / 16 : row_len[k]);
        if (work > most) most = work;
    }
    long long bx = (most + 255) / 256;
    if (bx > kNumSMs * 8) bx = kNumSMs * 8;
    if (bx < 1) bx = 1;
    u8_normalize_kernel<<<dim3((unsigned)bx, (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>(j);
    return launch_status("u8_normalize", 1);
}

extern "C" int pg_image_to_u8_bgr(const float* img, void* out, int32_t N, int32_t H, int32_t W, int32_t x0, int32_t x1, void* stream) {
    using namespace pg;
    PG_REQUIRE(N >= 0 && H >= 1 && W >= 1 && x0 >= 0 && x1 > x0 && x1 <= W, "image_to_u8_bgr: bad sizes / crop [%d, %d) of %d", x0, x1, W);
    if (N == 0) return PG_OK;
    PG_REQUIRE(img && out, "image_to_u8_bgr: img and out must be device pointers");
    const long long total = (long long)N * H * (x1 - x0);
    long long bx = (total + 255) / 256;
    if (bx > kNumSMs * 16) bx = kNumSMs * 16;
    image_to_u8_bgr_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(img, (unsigned char*)out, N, H, W, x0, x1);
    return launch_status("image_to_u8_bgr", 1);
}
