// oracle/ref_cuda_shim.cu -- TEST INFRASTRUCTURE ONLY (never linked into or loaded by the product).
//
// extern "C" entry points over the REFERENCE's own CUDA kernel launchers, compiled by `make -C oracle ref_cuda` from
// the .cu files where they lie under /root/reference/src/kernels/cuda (nothing is copied) into
// oracle/_ref/libit_ref_cuda.so.  The GPU parity tests run them on the B200 next to this repo's kernels: it pins
// RMSNorm (for which the reference has no test at all), RoPE, AttentionKVCache, Softmax and LayerNorm against the
// reference's actual device code instead of a restatement of it.  Device pointers in, legacy default stream.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "cuda/cuda_attention_kvcache.h"
#include "cuda/cuda_common.h"

namespace infini {
// defined in src/cuda/cuda_runtime.cc:511, which drags the cudnn/cublas runtime in; the launchers only need the symbol
thread_local cudaStream_t CUDAStream::_stream = nullptr;

// launchers of the reference (declared in its headers or file-local prototypes)
void rmsnorm_kernel(int dType, void *input, void *weight, void *output, int num_tokens, int hidden_size);   // rms_norm.cu:105
void rope_kernel(int dType, int *pos, void *input, void *output, int size, int dim_model, int dim_head,
                 int hidden_stride, int pos_stride);                                                          // rope.cu:82
void softmax_kernel(int num_blocks, float *input, float *output, int size, int dimsize, int stride);         // softmax.cu:242
void softmax_kernel(int num_blocks, half *input, half *output, int size, int dimsize, int stride);           // softmax.cu:324
void LaynormKernel(const float *input, const float *scale, const float eps, int size, int scaleSize, const int dimsize,
                   const int stride, float *output, const float *bias, int biasSize);                        // layer_norm.cu:339
void LaynormKernel(const float *input, const float *scale, const float eps, int size, int scaleSize, const int dimsize,
                   const int stride, float *output);                                                         // layer_norm.cu:396
}  // namespace infini

#define GUARD(stmt)                                                 \
    try {                                                           \
        stmt;                                                       \
    } catch (...) {                                                 \
        return 1;                                                   \
    }                                                               \
    return cudaDeviceSynchronize() == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : 2

extern "C" int ref_cuda_rmsnorm(int dtype, void *x, void *w, void *y, int tokens, int hidden) {
    GUARD(infini::rmsnorm_kernel(dtype, x, w, y, tokens, hidden));
}

extern "C" int ref_cuda_rope(int dtype, int *pos, void *x, void *y, int size, int dim_model, int dim_head,
                             int hidden_stride, int pos_stride) {
    GUARD(infini::rope_kernel(dtype, pos, x, y, size, dim_model, dim_head, hidden_stride, pos_stride));
}

extern "C" int ref_cuda_attention_kvcache(float *kc, float *vc, float *q, float *k, float *v, int *pos, float *out, int B,
                                          int H, int S, int D, float *tmp_o, float *tmp_sum) {
    AttentionKVCacheMetadata m;  // attention_kvcache.cc:28-39: dims and contiguous strides of the cache
    m.dimSize[0] = B; m.dimSize[1] = H; m.dimSize[2] = S; m.dimSize[3] = D;
    m.stride[3] = 1; m.stride[2] = D; m.stride[1] = S * D; m.stride[0] = H * S * D;
    GUARD(infini::attention_kvcache_kernel(kc, vc, q, k, v, pos, out, m, tmp_o, tmp_sum));
}

extern "C" int ref_cuda_softmax_f32(float *x, float *y, int size, int dimsize, int stride) {
    GUARD(infini::softmax_kernel(size / dimsize, x, y, size, dimsize, stride));
}

extern "C" int ref_cuda_softmax_f16(void *x, void *y, int size, int dimsize, int stride) {
    GUARD(infini::softmax_kernel(size / dimsize, (half *)x, (half *)y, size, dimsize, stride));
}

extern "C" int ref_cuda_layernorm_f32(const float *x, const float *scale, float eps, int size, int scale_size,
                                      int dimsize, int stride, float *y, const float *bias, int bias_size) {
    if (bias) { GUARD(infini::LaynormKernel(x, scale, eps, size, scale_size, dimsize, stride, y, bias, bias_size)); }
    GUARD(infini::LaynormKernel(x, scale, eps, size, scale_size, dimsize, stride, y));
}
