// conv_shapes.h -- which convolution shapes each kernel path takes.  Header-only and free of CUDA so the scheduler's layout pass
// (schedule.cc, also compiled on its own for the host unit tests) and the launchers (kernels/gemm.cu, kernels/conv_nhwc.cu) decide
// from the SAME predicates: a step is only marked NHWC when the kernel is certain to take it.
#pragma once
#include <cstdint>

#include "it_b200.h"

namespace itb {

inline void conv_out_hw(int H, int W, int R, int S, int ph, int pw, int sh, int sw, int dh, int dw, int &OH, int &OW) {
    OH = (H + 2 * ph - dh * (R - 1) - 1) / sh + 1;  // reference src/operators/conv.cc:85-114
    OW = (W + 2 * pw - dw * (S - 1) - 1) / sw + 1;
}

inline bool conv_is_1x1_direct(int R, int S, int ph, int pw, int sh, int sw) {
    return R == 1 && S == 1 && ph == 0 && pw == 0 && sh == 1 && sw == 1;
}

// the folded im2col GEMM ([F, Kc] x [Kc, N*P], scattered back by the epilogue) runs on the tcgen05 kernel
inline bool conv_fold_ok(int dtype, int N, int64_t P, int64_t Kc, int F, int groups) {
    // (Kc itself need not be a multiple of 8: the im2col rows and a copy of the filters are zero-padded to Kp = ceil8(Kc))
    return (dtype == ITB_F16 || dtype == ITB_BF16) && groups == 1 && (N * P) % 8 == 0 && ((Kc + 7) & ~7ll) >= 64 &&
           N * P >= 64 && N * P < (1ll << 31) && F >= 1;
}

// NCHW input, NHWC output through the folded im2col GEMM (it_b200_conv2d_fused_nhwc_out)
inline bool conv_nchw_to_nhwc_ok(int dtype, int N, int C, int H, int W, int F, int R, int S, int ph, int pw, int sh, int sw, int dh,
                                 int dw, int groups) {
    int OH, OW;
    conv_out_hw(H, W, R, S, ph, pw, sh, sw, dh, dw, OH, OW);
    const int64_t P = (int64_t)OH * OW, Kc = (int64_t)C * R * S;
    const bool direct_tc = conv_is_1x1_direct(R, S, ph, pw, sh, sw) && P % 8 == 0;
    return !direct_tc && conv_fold_ok(dtype, N, P, Kc, F, groups) && F % 8 == 0;
}

// NHWC input through the implicit-GEMM kernel (it_b200_conv2d_nhwc): TMA im2col limits
inline bool conv_nhwc_ok(int dtype, int C, int F, int R, int S, int ph, int pw, int sh, int sw, int dh, int dw, int groups) {
    return (dtype == ITB_F16 || dtype == ITB_BF16) && groups == 1 && C % 8 == 0 && F % 8 == 0 && C >= 8 && F >= 8 && R >= 1 &&
           S >= 1 && sh >= 1 && sw >= 1 && sh <= 8 && sw <= 8 && dh >= 1 && dw >= 1 && ph >= 0 && pw >= 0 && ph <= 127 &&
           pw <= 127 && (R - 1) * dh - ph <= 128 && (S - 1) * dw - pw <= 128 && (R - 1) * dh < 65536 && (S - 1) * dw < 65536;
}

// NCHW input with <= 4 channels, NHWC output through the stem kernel (it_b200_conv2d_stem): no residual
inline bool conv_stem_ok(int dtype, int C, int F, int R, int S, int ph, int pw, int sh, int sw, int dh, int dw, int groups) {
    return (dtype == ITB_F16 || dtype == ITB_BF16) && groups == 1 && C >= 1 && C <= 4 && F >= 8 && F <= 64 && F % 8 == 0 && R >= 1 &&
           R <= 7 && S >= 1 && S <= 8 && sh >= 1 && sh <= 2 && sw >= 1 && sw <= 2 && dh == 1 && dw == 1 && ph >= 0 && pw >= 0;
}

}  // namespace itb
