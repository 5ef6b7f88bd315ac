// common.cuh -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <utility>

#include "it_b200.h"

namespace itb {

// ---- error reporting (the C spelling of infini::Exception) -------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void take_prefetch_hint(const void *&ptr, long long &bytes);  // it_b200_l2_prefetch_hint (support.cu)

#define ITB_FAIL(...)                                                                          \
    do {                                                                                       \
        itb::set_error(__VA_ARGS__);                                                           \
        return 1;                                                                              \
    } while (0)

#define ITB_CHECK(cond, ...)                                                                   \
    do {                                                                                       \
        if (!(cond)) ITB_FAIL(__VA_ARGS__);                                                    \
    } while (0)

#define ITB_LAUNCH_CHECK(name)                                                                 \
    do {                                                                                       \
        cudaError_t e__ = cudaPeekAtLastError();                                               \
        if (e__ != cudaSuccess) ITB_FAIL("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
        itb::count_launch();                                                                   \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// Every kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization, and every kernel begins with
// pdl_trigger() (lets the NEXT kernel in the stream start its launch + prologue now) and executes pdl_wait() before
// its first read of an activation or its first global write (blocks until the PREVIOUS kernel has fully completed
// and flushed).  In a decode step of ~450 short kernels this hides the launch gap, and the GEMM kernels prefetch
// their weight tiles -- which no kernel in the step writes -- ahead of pdl_wait().
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();  // false when ITB_NO_PDL=1
// zero-initialised, self-cleaning int tickets private to (device, stream); nullptr if n exceeds the pool (65536)
int *stream_tickets(cudaStream_t st, int n);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline int dtype_size(int dt) {
    switch (dt) {
    case ITB_F32: case ITB_I32: case ITB_U32: return 4;
    case ITB_F16: case ITB_BF16: return 2;
    case ITB_I64: return 8;
    case ITB_U8: case ITB_I8: case ITB_BOOL: return 1;
    default: return 0;
    }
}

// ---- dtype traits --------------------------------------------------------------------------
template <typename T> struct TypeOps;
template <> struct TypeOps<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct TypeOps<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct TypeOps<__nv_bfloat16> {
    static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <typename T> __device__ __forceinline__ float to_f(T v) { return TypeOps<T>::to_f(v); }
template <typename T> __device__ __forceinline__ T from_f(float v) { return TypeOps<T>::from_f(v); }
// round a float to T and back (the reference's "(T)(expr)" points)
template <typename T> __device__ __forceinline__ float round_t(float v) { return to_f<T>(from_f<T>(v)); }

// inference BatchNorm arithmetic, spelled with intrinsics so the stand-alone kernel and the conv epilogue round alike:
//   y = scale * (x - mean) * rs + bias,  rs = 1 / sqrt(var + eps)      (reference batch_norm.cc:9-69 via cuDNN)
//   rs as the hardware reciprocal square root (what cudnnBatchNormalizationForwardInference evaluates): the reference's own
//   golden vector (test_cuda_batch_norm.cc:49-52: (8 - 9) / sqrt(9) expected "-0.333333" within its relative 1e-6) holds for
//   rsqrt(9) = 0.33333331 and NOT for the correctly rounded 1 / sqrt(9) = 0.33333334
__device__ __forceinline__ float bn_rs(float var, float eps) { return rsqrtf(var + eps); }
__device__ __forceinline__ float bn_apply(float x, float mean, float rs, float scale, float bias) {
    return __fmaf_rn(__fmul_rn(scale, __fsub_rn(x, mean)), rs, bias);
}

// 16-byte vector of T
template <typename T> struct Vec16 {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};
template <typename T> __device__ __forceinline__ Vec16<T> ld16(const T *p) {
    Vec16<T> r;
    *reinterpret_cast<uint4 *>(r.v) = *reinterpret_cast<const uint4 *>(p);
    return r;
}
template <typename T> __device__ __forceinline__ Vec16<T> ld16_stream(const T *p) {
    Vec16<T> r;
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    *reinterpret_cast<uint4 *>(r.v) = u;
    return r;
}
template <typename T> __device__ __forceinline__ void st16(T *p, const Vec16<T> &r) {
    *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(r.v);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- reductions -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; all threads get the result
__device__ __forceinline__ float block_sum(float v, float *red) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = lane < nw ? red[lane] : 0.f;
    return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float *red) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = lane < nw ? red[lane] : -INFINITY;
    return warp_max(r);
}

// grid sizing for HBM-bound grid-stride kernels: a few waves of 148 SMs
inline int grid_for(int64_t work_items, int threads, int max_ctas_per_sm = 8) {
    int64_t need = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)kNumSMs * max_ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

struct Dims8 {
    int64_t v[ITB_MAX_RANK];
};

#define ITB_DISPATCH_FLOAT(dt, NAME, ...)                                                      \
    switch (dt) {                                                                              \
    case ITB_F32: { using T = float; __VA_ARGS__; } break;                                     \
    case ITB_F16: { using T = __half; __VA_ARGS__; } break;                                    \
    case ITB_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break;                            \
    default: ITB_FAIL("%s: unsupported dtype %d", NAME, dt);                                   \
    }

}  // namespace itb
