/*
 * it_oracle.c -- CPU restatement of the InfiniTensor operator-kernel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library, and only as the checker or as
 * the reported CPU baseline -- never as the thing shipped or measured as ours.
 *
 * Parity pin status: every function below is checked in tests/test_oracle_golden.py
 * against the reference's own golden vectors (test/kernels/cuda/ *.cc) and, where
 * the reference's native-CPU backend implements the op, against the reference
 * itself compiled into oracle/_ref (see oracle/Makefile).  RMSNorm has no test
 * anywhere in the reference => "parity unpinned" for that one op; it is
 * restated from src/kernels/cuda/rms_norm.cu:36-54.
 *
 * Convention: all floating-point tensors travel as float32 arrays whose values
 * are already exactly representable in the storage dtype `dt`
 * (ONNX enum: 1 = f32, 10 = f16, 16 = bf16).  Every point at which the
 * reference stores or rounds to T is reproduced with orc_round(x, dt).
 * All citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DT_F32 1
#define DT_F16 10
#define DT_BF16 16

/* ---- storage-dtype rounding (round-to-nearest-even), src/utils/data_convert.cc:5-41 ---- */
static inline float round_bf16(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { /* NaN */
        u |= 0x00400000u;
        u &= 0xffff0000u;
    } else {
        uint32_t lsb = (u >> 16) & 1u;
        u += 0x7fffu + lsb;
        u &= 0xffff0000u;
    }
    memcpy(&x, &u, 4);
    return x;
}
static inline float round_f16(float x) { return (float)(_Float16)x; }

float orc_round(float x, int dt) {
    if (dt == DT_F16) return round_f16(x);
    if (dt == DT_BF16) return round_bf16(x);
    return x;
}

void orc_round_array(float *x, int64_t n, int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) x[i] = orc_round(x[i], dt);
}

/* ======================================================================
 * MatMul.  src/operators/matmul.cc:26-49 (shape rule), src/kernels/cuda/matmul.cc:66-174
 * C[b,m,n] = op(A)[b,m,k] . op(B)[b,k,n] (+ bias pre-expanded into C, beta = 1),
 * batch broadcast by stride 0 (matmul.cc:124-137), `act` ignored (quirk q5).
 * fp32 accumulate in a fixed k-ascending order, result rounded to T once.
 * biasFull is bias already broadcast to [b,m,n] (NULL = none).
 * ====================================================================== */
void orc_matmul(const float *A, const float *B, const float *biasFull, float *C,
                int64_t nb, int64_t strideA, int64_t strideB, int m, int n, int k,
                int transA, int transB, int dt) {
    const int JB = 256; /* column block: gives the host threads independent work at m=16 */
    int njb = (n + JB - 1) / JB;
    int64_t tasks = nb * (int64_t)m * njb;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < tasks; ++t) {
        int jb = (int)(t % njb);
        int64_t bi = t / njb;
        int i = (int)(bi % m);
        int64_t b = bi / m;
        const float *a = A + b * strideA;
        const float *bm = B + b * strideB;
        int j0 = jb * JB, j1 = j0 + JB < n ? j0 + JB : n;
        float acc[256];
        for (int j = j0; j < j1; ++j) acc[j - j0] = 0.f;
        if (!transB) {
            for (int kk = 0; kk < k; ++kk) {
                float av = transA ? a[(int64_t)kk * m + i] : a[(int64_t)i * k + kk];
                const float *brow = bm + (int64_t)kk * n;
                for (int j = j0; j < j1; ++j) acc[j - j0] += av * brow[j];
            }
        } else {
            for (int j = j0; j < j1; ++j) {
                const float *bcol = bm + (int64_t)j * k;
                float s = 0.f;
                if (!transA) {
                    const float *arow = a + (int64_t)i * k;
                    for (int kk = 0; kk < k; ++kk) s += arow[kk] * bcol[kk];
                } else {
                    for (int kk = 0; kk < k; ++kk) s += a[(int64_t)kk * m + i] * bcol[kk];
                }
                acc[j - j0] = s;
            }
        }
        float *c = C + (b * m + i) * (int64_t)n;
        const float *bias = biasFull ? biasFull + (b * m + i) * (int64_t)n : NULL;
        for (int j = j0; j < j1; ++j)
            c[j] = orc_round(acc[j - j0] + (bias ? bias[j] : 0.f), dt);
    }
}

/* ======================================================================
 * Conv (2-D cross-correlation, NCHW x FCRS -> NFHW, groups, symmetric padding).
 * src/operators/conv.cc:85-114 (output dims), src/kernels/cuda/conv.cc:143-168,
 * src/kernels/cpu/conv.cc:10-52 (native CPU loop nest).
 * ====================================================================== */
void orc_conv2d(const float *x, const float *w, float *y, int N, int C, int H, int W,
                int F, int R, int S, int ph, int pw, int sh, int sw, int dh, int dw,
                int groups, int dt) {
    int Cg = C / groups, Fg = F / groups;
    int OH = (H + 2 * ph - dh * (R - 1) - 1) / sh + 1;
    int OW = (W + 2 * pw - dw * (S - 1) - 1) / sw + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int f = 0; f < F; ++f) {
            int g = f / Fg;
            for (int oh = 0; oh < OH; ++oh)
                for (int ow = 0; ow < OW; ++ow) {
                    float acc = 0.f;
                    for (int c = 0; c < Cg; ++c)
                        for (int r = 0; r < R; ++r) {
                            int ih = oh * sh - ph + r * dh;
                            if (ih < 0 || ih >= H) continue;
                            for (int s = 0; s < S; ++s) {
                                int iw = ow * sw - pw + s * dw;
                                if (iw < 0 || iw >= W) continue;
                                acc += x[(((int64_t)n * C + g * Cg + c) * H + ih) * W + iw] *
                                       w[(((int64_t)f * Cg + c) * R + r) * S + s];
                            }
                        }
                    y[(((int64_t)n * F + f) * OH + oh) * OW + ow] = orc_round(acc, dt);
                }
        }
}

/* ======================================================================
 * AttentionKVCache (decode, q-len 1).  src/kernels/cuda/attention_kvcache.cu:8-145.
 *  - seq_length = position_id[0] + 1 for every batch row (.cu:17)
 *  - k, v are appended IN PLACE into the cache inputs at position_id[0] (.cu:49-53, 89-93)
 *  - P = exp(q.k / sqrt(128)) WITHOUT max subtraction (.cu:72-80), per 16-token chunk
 *    O_chunk = (sum P v) / sum_chunk, then merge: O = sum(O_chunk * sum_chunk) / sum(sum_chunk) (.cu:110-140)
 *  - cache layout [B, H, S_max, D] contiguous, q/k/v/out [B, H, 1, D]
 * The reference is fp32 + D = 128 only; for f16/bf16 storage we keep the same fp32
 * arithmetic and round the output once (documented extension).
 * ====================================================================== */
void orc_attention_kvcache(float *kcache, float *vcache, const float *q, const float *k,
                           const float *v, int64_t pos, float *out, int B, int H, int Smax,
                           int D, int dt) {
    const int SEQ_UNIT = 16;
    int seq = (int)pos + 1;
    float inv = (float)sqrt(128.0); /* scale fixed at sqrt(128) regardless of D (.cu:72) */
#pragma omp parallel for schedule(static)
    for (int bh = 0; bh < B * H; ++bh) {
        float *kc = kcache + (int64_t)bh * Smax * D;
        float *vc = vcache + (int64_t)bh * Smax * D;
        const float *qq = q + (int64_t)bh * D;
        memcpy(kc + (int64_t)pos * D, k + (int64_t)bh * D, sizeof(float) * D);
        memcpy(vc + (int64_t)pos * D, v + (int64_t)bh * D, sizeof(float) * D);
        float *osum = (float *)calloc(D, sizeof(float));
        float *oc = (float *)malloc(sizeof(float) * D);
        float tot = 0.f;
        for (int s0 = 0; s0 < seq; s0 += SEQ_UNIT) {
            float csum = 0.f;
            for (int d = 0; d < D; ++d) oc[d] = 0.f;
            for (int s = s0; s < s0 + SEQ_UNIT && s < seq; ++s) {
                float dot = 0.f;
                for (int d = 0; d < D; ++d) dot += qq[d] * kc[(int64_t)s * D + d];
                float p = expf(dot / inv);
                csum += p;
                for (int d = 0; d < D; ++d) oc[d] = fmaf(p, vc[(int64_t)s * D + d], oc[d]);
            }
            for (int d = 0; d < D; ++d) osum[d] += (oc[d] / csum) * csum;
            tot += csum;
        }
        for (int d = 0; d < D; ++d) out[(int64_t)bh * D + d] = orc_round(osum[d] / tot, dt);
        free(osum);
        free(oc);
    }
}

/* ======================================================================
 * Softmax along one axis with stride.  src/kernels/cuda/softmax.cu:3-97 (online max/sum),
 * dispatch :242-404.  x viewed as [outer, dim, inner].
 * ====================================================================== */
void orc_softmax(const float *x, float *y, int64_t outer, int dim, int64_t inner, int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t oi = 0; oi < outer * inner; ++oi) {
        int64_t o = oi / inner, i = oi % inner;
        const float *px = x + o * dim * inner + i;
        float *py = y + o * dim * inner + i;
        float mx = -INFINITY;
        for (int d = 0; d < dim; ++d) mx = fmaxf(mx, px[d * inner]);
        double sumd = 0.0; /* the reference reduces as a tree (cub::BlockReduce); a double sum is order-free */
        for (int d = 0; d < dim; ++d) sumd += (double)expf(px[d * inner] - mx);
        float sum = (float)sumd;
        for (int d = 0; d < dim; ++d) py[d * inner] = orc_round(expf(px[d * inner] - mx) / sum, dt);
    }
}

/* ======================================================================
 * LayerNormalization over ONE axis with stride (the reference normalises dims[axis]
 * only, src/kernels/cuda/layer_norm.cc:21-27, layer_norm.cu:4-90); scale / bias are
 * either length-dim vectors or scalars (.cu:41-89).  fp32 accumulation (quirk q4:
 * the reference accumulates in T; we do not copy that).
 * ====================================================================== */
void orc_layernorm(const float *x, const float *scale, const float *bias, float *y,
                   int64_t outer, int dim, int64_t inner, int scaleSize, int biasSize,
                   float eps, int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t oi = 0; oi < outer * inner; ++oi) {
        int64_t o = oi / inner, i = oi % inner;
        const float *px = x + o * dim * inner + i;
        float *py = y + o * dim * inner + i;
        double mud = 0.0; /* order-free sums (the reference reduces as a tree) */
        for (int d = 0; d < dim; ++d) mud += (double)px[d * inner];
        float mu = (float)(mud / dim);
        double vard = 0.0;
        for (int d = 0; d < dim; ++d) {
            float t = px[d * inner] - mu;
            vard += (double)(t * t);
        }
        float var = (float)(vard / dim);
        float rs = 1.0f / sqrtf(var + eps);
        for (int d = 0; d < dim; ++d) {
            float s = scale[scaleSize == dim ? d : 0];
            float b = biasSize > 0 ? bias[biasSize == dim ? d : 0] : 0.f;
            py[d * inner] = orc_round(s * (px[d * inner] - mu) * rs + b, dt);
        }
    }
}

/* ======================================================================
 * RMSNorm.  src/kernels/cuda/rms_norm.cu:36-54: eps 1e-5 hard-coded (:46), and
 * (T)(x * rsqrt) is rounded to T BEFORE the multiply by weight (:52) (quirk q3).
 * ====================================================================== */
void orc_rmsnorm(const float *x, const float *w, float *y, int64_t tokens, int hidden, int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < tokens; ++t) {
        const float *px = x + t * hidden;
        float var = 0.f;
        for (int d = 0; d < hidden; ++d) var += px[d] * px[d];
        float r = 1.0f / sqrtf(var / (float)hidden + 0.00001f);
        for (int d = 0; d < hidden; ++d)
            y[t * hidden + d] = orc_round(orc_round(px[d] * r, dt) * w[d], dt);
    }
}

/* ======================================================================
 * RoPE (rotate-half).  src/kernels/cuda/rope.cu:7-31; dim_head hard-coded 128
 * (rope.cc:25).  input [B, S, dim_model], pos [B, S].  The reference launch only
 * covers batch 0 / token 0 (rope.cu:84-85, quirk q2 = defect); the evident intent
 * -- every (b, s) row -- is what is restated here.  Arithmetic in T exactly as
 * written: in*T(cos) -/+ in2*T(sin), each product and the sum rounded to T.
 * ====================================================================== */
void orc_rope(const int64_t *pos, const float *x, float *y, int B, int S, int dim_model,
              int dim_head, int dt) {
    int half = dim_head / 2;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < S; ++s) {
            int64_t off = ((int64_t)b * S + s) * dim_model;
            int p = (int)pos[(int64_t)b * S + s];
            for (int i = 0; i < dim_model; ++i) {
                int col = i % dim_head;
                int c = col < half ? col : col - half;
                float freq = (float)p * powf(10000.f, -(float)(c * 2) / (float)dim_head);
                float cs = orc_round((float)cos((double)freq), dt);
                float sn = orc_round((float)sin((double)freq), dt);
                float a = orc_round(x[off + i] * cs, dt);
                float r;
                if (col < half)
                    r = a - orc_round(x[off + i + half] * sn, dt);
                else
                    r = a + orc_round(x[off + i - half] * sn, dt);
                y[off + i] = orc_round(r, dt);
            }
        }
}

/* ======================================================================
 * Unary family.  src/kernels/cuda/unary.cu:31-143 (formulas), cudnn Relu/Sigmoid/Tanh
 * unary.cc:70-122; native CPU src/kernels/cpu/unary.cc:24-60.
 * op codes are private to the oracle (see oracle/__init__.py).
 * ====================================================================== */
enum { U_RELU = 0, U_SIGMOID, U_TANH, U_GELU, U_SILU, U_ERF, U_NEG, U_ABS, U_SQRT,
       U_HARDSIGMOID, U_HARDSWISH, U_EXP };

void orc_unary(int op, const float *x, float *y, int64_t n, int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i], r;
        switch (op) {
        case U_RELU: r = v > 0.f ? v : 0.f; break;
        case U_SIGMOID: r = 1.f / (1.f + expf(-v)); break;
        case U_TANH: r = tanhf(v); break;
        case U_GELU: r = (float)(0.5 * v * (1 + erf(v / sqrtf(2.0f)))); break; /* unary.cu:113 */
        case U_SILU: r = (float)(v / (1.0 + expf(-v))); break;                /* unary.cu:123 */
        case U_ERF: r = erff(v); break;
        case U_NEG: r = -v; break;
        case U_ABS: r = v < 0 ? -v : v; break;
        case U_SQRT: r = sqrtf(v); break;
        case U_HARDSIGMOID: r = fmaxf(0.0f, fminf(1.0f, 0.2f * v + 0.5f)); break;
        case U_HARDSWISH: r = v * fmaxf(0.f, fminf(1.f, (1.f / 6.f) * v + 0.5f)); break;
        case U_EXP: r = expf(v); break;
        default: r = NAN;
        }
        y[i] = orc_round(r, dt);
    }
}

/* ======================================================================
 * Binary with numpy broadcasting (rank <= 8).  cudnnOpTensor path
 * src/kernels/cuda/element_wise.cc:13-82 (Add/Sub/Mul/Min/Max), SIMT path
 * element_wise.cu:9-131 (Div/Pow/Less).  Strides are in elements, 0 on broadcast dims.
 * Comparison ops write 0/1.
 * ====================================================================== */
enum { B_ADD = 0, B_SUB, B_MUL, B_DIV, B_POW, B_MIN, B_MAX, B_LESS, B_EQUAL, B_GREATER };

void orc_binary(int op, const float *a, const float *b, float *c, int rank,
                const int64_t *dims, const int64_t *sa, const int64_t *sb, int dt) {
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < n; ++idx) {
        int64_t rem = idx, oa = 0, ob = 0;
        for (int i = rank - 1; i >= 0; --i) {
            int64_t c_i = rem % dims[i];
            rem /= dims[i];
            oa += c_i * sa[i];
            ob += c_i * sb[i];
        }
        float x = a[oa], y = b[ob], r;
        switch (op) {
        case B_ADD: r = x + y; break;
        case B_SUB: r = x - y; break;
        case B_MUL: r = x * y; break;
        case B_DIV: r = x / y; break;
        case B_POW: r = powf(x, y); break;
        case B_MIN: r = fminf(x, y); break;
        case B_MAX: r = fmaxf(x, y); break;
        case B_LESS: r = x < y ? 1.f : 0.f; break;
        case B_EQUAL: r = x == y ? 1.f : 0.f; break;
        case B_GREATER: r = x > y ? 1.f : 0.f; break;
        default: r = NAN;
        }
        c[idx] = (op >= B_LESS) ? r : orc_round(r, dt);
    }
}

/* ======================================================================
 * MaxPool / AveragePool 2-D, NCHW.  src/kernels/cuda/pooling.cc:8-95; average counts
 * padding (CUDNN_POOLING_AVERAGE_COUNT_INCLUDE_PADDING, pooling.cc:88, quirk q9);
 * output dims src/operators/pooling.cc (floor unless ceilMode).
 * ====================================================================== */
void orc_pool2d(int is_max, const float *x, float *y, int N, int C, int H, int W, int kh,
                int kw, int dh, int dw, int ph, int pw, int sh, int sw, int OH, int OW,
                int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t nc = 0; nc < (int64_t)N * C; ++nc)
        for (int oh = 0; oh < OH; ++oh)
            for (int ow = 0; ow < OW; ++ow) {
                float acc = is_max ? -INFINITY : 0.f;
                for (int r = 0; r < kh; ++r)
                    for (int s = 0; s < kw; ++s) {
                        int ih = oh * sh - ph + r * dh, iw = ow * sw - pw + s * dw;
                        int inb = ih >= 0 && ih < H && iw >= 0 && iw < W;
                        if (is_max) {
                            if (inb) acc = fmaxf(acc, x[(nc * H + ih) * W + iw]);
                        } else if (inb)
                            acc += x[(nc * H + ih) * W + iw];
                    }
                if (!is_max) acc /= (float)(kh * kw);
                y[(nc * OH + oh) * OW + ow] = orc_round(acc, dt);
            }
}

/* ======================================================================
 * BatchNormalization, inference, spatial.  src/kernels/cuda/batch_norm.cc:9-69
 * (cudnnBatchNormalizationForwardInference): y = scale*(x-mean)/sqrt(var+eps)+bias.
 * ====================================================================== */
void orc_batchnorm(const float *x, const float *mean, const float *var, const float *scale,
                   const float *bias, float *y, int N, int C, int64_t HW, float eps, int dt) {
#pragma omp parallel for schedule(static)
    for (int64_t nc = 0; nc < (int64_t)N * C; ++nc) {
        int c = (int)(nc % C);
        float rs = 1.0f / sqrtf(var[c] + eps);
        for (int64_t i = 0; i < HW; ++i)
            y[nc * HW + i] = orc_round(scale[c] * (x[nc * HW + i] - mean[c]) * rs + bias[c], dt);
    }
}
