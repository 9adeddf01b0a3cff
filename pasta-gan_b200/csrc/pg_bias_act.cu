// bias_act for sm_100a: y = clamp(act(x + b) * gain), plus the first / second-order gradient forms.
//
// Replaces the reference's bias_act_kernel (torch_utils/ops/bias_act.cu:23-147, launcher
// bias_act.cpp:32-90).  The op is pure HBM streaming (8 B/element forward, 12 B/element grad=1), so
// the design is about bytes in flight, not math:
//   * 16-byte vector loads/stores (4 x fp32 / 8 x fp16 / 2 x fp64 per access) on the read-only path,
//   * kUnroll independent vectors per thread issued before any use (>= 64 B in flight per thread per
//     operand -> ~100 KB per SM at full occupancy, well above the ~35 KB latency-bandwidth product),
//   * one bias lookup per vector instead of the reference's div+mod per element (legal whenever
//     step_b is a multiple of the vector width, i.e. every NCHW plane with H*W % 4 == 0),
//   * a scalar kernel covers unaligned bases, odd step_b and the (size_x % vector) tail.
#include "pg_common.cuh"

namespace pg {

struct BiasActParams {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    uint32_t size_x; uint32_t size_b; uint32_t step_b;
    int grad; float alpha, gain, clamp;
    uint32_t elem_begin;   // scalar kernel: first element it owns
};

template <class S> __device__ __forceinline__ S s_exp(S v);
template <> __device__ __forceinline__ float  s_exp<float>(float v)   { return expf(v); }
template <> __device__ __forceinline__ double s_exp<double>(double v) { return exp(v); }
template <class S> __device__ __forceinline__ S s_log1p(S v);
template <> __device__ __forceinline__ float  s_log1p<float>(float v)   { return log1pf(v); }
template <> __device__ __forceinline__ double s_log1p<double>(double v) { return log1p(v); }
template <class S> __device__ __forceinline__ S s_tanh(S v);
template <> __device__ __forceinline__ float  s_tanh<float>(float v)   { return tanhf(v); }
template <> __device__ __forceinline__ double s_tanh<double>(double v) { return tanh(v); }

// One element.  `v` is x (grad 0) or the incoming gradient (grad >= 1); `xb` = xref + b; `yref` as stored.
template <class S, int A>
__device__ __forceinline__ S bias_act_elem(int G, S v, S b, S xref, S yref, S dy, S alpha, S gain, S clamp) {
    const S one = (S)1, two = (S)2;
    const S selu_scale = (S)1.0507009873554804934193349852946;
    const S selu_sa    = (S)(1.0507009873554804934193349852946 * 1.6732632423543772848170429916717);
    const S yy = (gain != (S)0) ? yref / gain : (S)0;
    S y = (S)0;
    if (G == 0) {
        const S t = v + b;
        if (A == PG_ACT_LINEAR)   y = t;
        if (A == PG_ACT_RELU)     y = t > 0 ? t : (S)0;
        if (A == PG_ACT_LRELU)    y = t > 0 ? t : t * alpha;
        if (A == PG_ACT_TANH)     y = s_tanh<S>(t);
        if (A == PG_ACT_SIGMOID)  y = one / (one + s_exp<S>(-t));
        if (A == PG_ACT_ELU)      y = t >= 0 ? t : s_exp<S>(t) - one;
        if (A == PG_ACT_SELU)     y = t >= 0 ? selu_scale * t : selu_sa * (s_exp<S>(t) - one);
        if (A == PG_ACT_SOFTPLUS) y = t > (S)20 ? t : s_log1p<S>(s_exp<S>(t));
        if (A == PG_ACT_SWISH)    y = t / (one + s_exp<S>(-t));
    } else {
        S d1 = one, d2 = (S)0;   // act'(.) and act''(.)
        if (A == PG_ACT_RELU)     { d1 = yy > 0 ? one : (S)0; }
        if (A == PG_ACT_LRELU)    { d1 = yy > 0 ? one : alpha; }
        if (A == PG_ACT_TANH)     { d1 = one - yy * yy;          d2 = d1 * (-two * yy); }
        if (A == PG_ACT_SIGMOID)  { d1 = yy * (one - yy);        d2 = d1 * (one - two * yy); }
        if (A == PG_ACT_ELU)      { d1 = yy >= 0 ? one : yy + one;              d2 = yy >= 0 ? (S)0 : yy + one; }
        if (A == PG_ACT_SELU)     { d1 = yy >= 0 ? selu_scale : yy + selu_sa;   d2 = yy >= 0 ? (S)0 : yy + selu_sa; }
        if (A == PG_ACT_SOFTPLUS) { const S c = s_exp<S>(-yy); d1 = one - c;    d2 = c * (one - c); }
        if (A == PG_ACT_SWISH) {
            const S t = xref + b;
            const S s = one / (one + s_exp<S>(-t));
            d1 = s * (one + t * (one - s));
            d2 = s * (one - s) * (two + t * (one - two * s));
            yref = t * s * gain;            // clamp mask is recomputed from x (bias_act.cu:128)
        }
        y = v * (G == 1 ? d1 : d2);
    }
    y *= gain * dy;
    if (clamp >= (S)0) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0 ? clamp : -clamp);
        else        y = (yref > -clamp && yref < clamp) ? y : (S)0;
    }
    return y;
}

constexpr int kThreads = 256;
constexpr int kUnroll  = 4;

template <class T, int A>
__global__ void __launch_bounds__(kThreads) bias_act_vec_kernel(BiasActParams p) {
    typedef typename Acc<T>::type S;
    constexpr int V = Vec16<T>::N;
    const uint32_t nvec = p.size_x / V;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const int G = p.grad;
    const T* __restrict__ px = (const T*)p.x;
    const T* __restrict__ pb = (const T*)p.b;
    const T* __restrict__ pxr = (const T*)p.xref;
    const T* __restrict__ pyr = (const T*)p.yref;
    const T* __restrict__ pdy = (const T*)p.dy;
    T* __restrict__ py = (T*)p.y;

    const uint32_t base = blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
    Vec16<T> vx[kUnroll], vxr[kUnroll], vyr[kUnroll], vdy[kUnroll];
    S bias[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
        const uint32_t i = base + u * kThreads;
        if (i < nvec) {
            vx[u] = ld16(px + (size_t)i * V);
            if (pxr) vxr[u] = ld16(pxr + (size_t)i * V);
            if (pyr) vyr[u] = ld16(pyr + (size_t)i * V);
            if (pdy) vdy[u] = ld16(pdy + (size_t)i * V);
            bias[u] = pb ? to_acc<T>(__ldg(pb + ((i * V) / p.step_b) % p.size_b)) : (S)0;
        }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; u++) {
        const uint32_t i = base + u * kThreads;
        if (i < nvec) {
            Vec16<T> out;
#pragma unroll
            for (int k = 0; k < V; k++) {
                const S xr = pxr ? to_acc<T>(vxr[u].v[k]) : (S)0;
                const S yr = pyr ? to_acc<T>(vyr[u].v[k]) : (S)0;
                const S d  = pdy ? to_acc<T>(vdy[u].v[k]) : (S)1;
                out.v[k] = from_acc<T, S>(bias_act_elem<S, A>(G, to_acc<T>(vx[u].v[k]), bias[u], xr, yr, d, alpha, gain, clamp));
            }
            st16(py + (size_t)i * V, out);
        }
    }
}

// Scalar form: any alignment, any step_b.  Owns elements [elem_begin, size_x).
template <class T, int A>
__global__ void __launch_bounds__(kThreads) bias_act_scalar_kernel(BiasActParams p) {
    typedef typename Acc<T>::type S;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const T* px = (const T*)p.x; const T* pb = (const T*)p.b; const T* pxr = (const T*)p.xref;
    const T* pyr = (const T*)p.yref; const T* pdy = (const T*)p.dy; T* py = (T*)p.y;
    for (uint64_t i = (uint64_t)p.elem_begin + (uint64_t)blockIdx.x * kThreads + threadIdx.x; i < p.size_x;
         i += (uint64_t)gridDim.x * kThreads) {
        const S b  = pb ? to_acc<T>(pb[(i / p.step_b) % p.size_b]) : (S)0;
        const S xr = pxr ? to_acc<T>(pxr[i]) : (S)0;
        const S yr = pyr ? to_acc<T>(pyr[i]) : (S)0;
        const S d  = pdy ? to_acc<T>(pdy[i]) : (S)1;
        py[i] = from_acc<T, S>(bias_act_elem<S, A>(p.grad, to_acc<T>(px[i]), b, xr, yr, d, alpha, gain, clamp));
    }
}

template <class T, int A>
static int launch_bias_act(BiasActParams p, cudaStream_t stream) {
    constexpr int V = Vec16<T>::N;
    const bool vec_ok = aligned16(p.x) && aligned16(p.y) && (!p.xref || aligned16(p.xref)) && (!p.yref || aligned16(p.yref)) &&
                        (!p.dy || aligned16(p.dy)) && (!p.b || p.step_b % V == 0) && p.size_x >= (uint32_t)V;
    uint32_t done = 0;
    int launches = 0;
    if (vec_ok) {
        launches++;
        const uint32_t nvec = p.size_x / V;
        const uint32_t blocks = (nvec + kThreads * kUnroll - 1) / (kThreads * kUnroll);
        bias_act_vec_kernel<T, A><<<blocks, kThreads, 0, stream>>>(p);
        done = nvec * V;
    }
    if (done < p.size_x) {
        launches++;
        p.elem_begin = done;
        const uint64_t rest = p.size_x - done;
        uint64_t blocks = (rest + kThreads - 1) / kThreads;
        if (blocks > (uint64_t)kNumSMs * 32) blocks = (uint64_t)kNumSMs * 32;
        bias_act_scalar_kernel<T, A><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
    }
    return launch_status("bias_act", launches);
}

template <class T>
static int dispatch_act(const BiasActParams& p, int act, cudaStream_t s) {
    switch (act) {
        case PG_ACT_LINEAR:   return launch_bias_act<T, PG_ACT_LINEAR>(p, s);
        case PG_ACT_RELU:     return launch_bias_act<T, PG_ACT_RELU>(p, s);
        case PG_ACT_LRELU:    return launch_bias_act<T, PG_ACT_LRELU>(p, s);
        case PG_ACT_TANH:     return launch_bias_act<T, PG_ACT_TANH>(p, s);
        case PG_ACT_SIGMOID:  return launch_bias_act<T, PG_ACT_SIGMOID>(p, s);
        case PG_ACT_ELU:      return launch_bias_act<T, PG_ACT_ELU>(p, s);
        case PG_ACT_SELU:     return launch_bias_act<T, PG_ACT_SELU>(p, s);
        case PG_ACT_SOFTPLUS: return launch_bias_act<T, PG_ACT_SOFTPLUS>(p, s);
        case PG_ACT_SWISH:    return launch_bias_act<T, PG_ACT_SWISH>(p, s);
    }
    return fail(PG_ERR_INVALID_ARGUMENT, "no CUDA kernel found for the specified activation func (act=%d)", act);
}

}  // namespace pg

extern "C" int pg_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                           int64_t size_x, int32_t size_b, int64_t step_b,
                           int32_t grad, int32_t act, float alpha, float gain, float clamp,
                           int32_t dtype, void* stream) {
    using namespace pg;
    PG_REQUIRE(size_x >= 0 && size_x <= INT32_MAX, "x is too large");
    PG_REQUIRE(grad >= 0 && grad <= 2, "grad must be 0, 1 or 2");
    PG_REQUIRE(b == nullptr || (size_b >= 1 && step_b >= 1), "b has wrong number of elements");
    if (size_x == 0) return PG_OK;
    PG_REQUIRE(x != nullptr && y != nullptr, "x and y must be device pointers");
    BiasActParams p;
    p.x = x; p.b = b; p.xref = xref; p.yref = yref; p.dy = dy; p.y = y;
    p.size_x = (uint32_t)size_x; p.size_b = b ? (uint32_t)size_b : 1u; p.step_b = b ? (uint32_t)(step_b > INT32_MAX ? INT32_MAX : step_b) : 1u;
    p.grad = grad; p.alpha = alpha; p.gain = gain; p.clamp = clamp; p.elem_begin = 0;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case PG_F32: return dispatch_act<float>(p, act, s);
        case PG_F16: return dispatch_act<__half>(p, act, s);
        case PG_F64: return dispatch_act<double>(p, act, s);
    }
    return fail(PG_ERR_INVALID_ARGUMENT, "unsupported dtype %d", dtype);
}
