// Shared helpers for the sm_100a kernels behind include/pasta_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/pasta_b200.h"

namespace pg {

// ---- thread-local error slot (pg_last_error) -------------------------------------------------
char* error_slot();
int   fail(int code, const char* fmt, ...);

#define PG_REQUIRE(cond, ...) do { if (!(cond)) return ::pg::fail(PG_ERR_INVALID_ARGUMENT, __VA_ARGS__); } while (0)
#define PG_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
    return ::pg::fail(PG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)

void count_launch(int n = 1);

inline int launch_status(const char* what, int launches = 1) {
    count_launch(launches);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PG_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    return PG_OK;
}

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- element-type traits: storage type T, arithmetic type acc_t (fp32, or fp64 for double) ----
template <class T> struct Acc            { typedef float  type; };
template <>        struct Acc<double>    { typedef double type; };

template <class T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v)      { return (typename Acc<T>::type)v; }
template <>        __device__ __forceinline__ float to_acc<__half>(__half v)           { return __half2float(v); }
template <class T, class A> __device__ __forceinline__ T from_acc(A v)                { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half, float>(float v)         { return __float2half_rn(v); }

// 16-byte vector of T
template <class T> struct Vec16 { static constexpr int N = 16 / sizeof(T); T v[16 / sizeof(T)]; };

template <class T> __device__ __forceinline__ Vec16<T> ld16(const T* p) {
    Vec16<T> r; *reinterpret_cast<int4*>(&r) = __ldg(reinterpret_cast<const int4*>(p)); return r;
}
template <class T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& r) {
    *reinterpret_cast<int4*>(p) = *reinterpret_cast<const int4*>(&r);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pg
