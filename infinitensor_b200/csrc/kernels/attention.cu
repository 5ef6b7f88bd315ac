// attention.cu -- AttentionKVCache decode kernel (q-len 1) for sm_100a.
//
// Replaces reference _attention_kvcache_kernel_128_1/_2 (src/kernels/cuda/attention_kvcache.cu:8-169,
// wrapper attention_kvcache.cc:8-56).  Contract kept from the reference:
//   * seq_length = position_id[0] + 1 for every batch row (.cu:17)
//   * k, v are appended IN PLACE into the cache INPUT tensors at position_id[0] (.cu:49-53, 89-93)
//   * cache layout [B, H, S_max, 128] contiguous; q/k/v/out [B, H, 1, 128]; scale 1/sqrt(128) (.cu:72)
// Differences by design: numerically stable online softmax instead of the reference's
// exp without max-subtraction (same mathematics; quirk q1), f16/bf16 caches in addition to fp32,
// no 2 GiB scratch round trip -- a single kernel when B*H alone fills the 148 SMs.
//
// HBM-bound: every K and V row up to position p is read exactly once with 128-bit loads; a warp
// holds RPW = 32/LPR rows per load instruction (LPR = lanes per 128-wide row: 16 for 2-byte types,
// 32 for fp32), U independent K and V loads are issued before any arithmetic so each lane keeps
// 2*U 16-byte requests in flight.
#include "common.cuh"

namespace itb {

constexpr int kD = 128;

template <typename T> struct RowCfg {
    static constexpr int EPL = 16 / sizeof(T);  // elements per lane per row
    static constexpr int LPR = kD / EPL;        // lanes per row
    static constexpr int RPW = 32 / LPR;        // rows per warp-load
};

template <typename P> __device__ __forceinline__ int read_pos(const void *p) { return (int)((const P *)p)[0]; }

// partial layout in workspace: [BH, nsplit] x { m, l, acc[128] }
// rotate-half RoPE of this lane's EPL dims of one 128-wide head row, arithmetic rounded to T exactly like rope_kernel
// (norm.cu / reference rope.cu:21-29); the partner dims (+-64) live in lane ^ (LPR/2) of the same row group
template <typename T, int EPL, int LPR>
__device__ __forceinline__ Vec16<T> rope_row(const Vec16<T> &x, int col, float p) {
    Vec16<T> partner, r;
    {
        uint4 mine = *reinterpret_cast<const uint4 *>(x.v), other;
        other.x = __shfl_xor_sync(0xffffffffu, mine.x, LPR / 2);
        other.y = __shfl_xor_sync(0xffffffffu, mine.y, LPR / 2);
        other.z = __shfl_xor_sync(0xffffffffu, mine.z, LPR / 2);
        other.w = __shfl_xor_sync(0xffffffffu, mine.w, LPR / 2);
        *reinterpret_cast<uint4 *>(partner.v) = other;
    }
    const bool lo = col < kD / 2;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
        const int c = (col + j) & (kD / 2 - 1);
        const float freq = p * powf(10000.f, -(float)(c * 2) / (float)kD);
        const float cs = round_t<T>(cosf(freq)), sn = round_t<T>(sinf(freq));
        const float a = round_t<T>(to_f(x.v[j]) * cs), b = round_t<T>(to_f(partner.v[j]) * sn);
        r.v[j] = from_f<T>(lo ? a - b : a + b);
    }
    return r;
}

// ROPE = true: q and kin are the PRE-RoPE projections; RoPE (position rope_pos[b]) is applied on load, so the two
// RoPE kernels of the layer disappear (the appended cache row is the rotated k, as in the unfused graph)
template <typename T, int WARPS, int U, bool ROPE>
__global__ void __launch_bounds__(WARPS * 32) attn_decode_kernel(T *__restrict__ kcache, T *__restrict__ vcache,
                                                                 const T *__restrict__ q,
                                                                 const T *__restrict__ kin,
                                                                 const T *__restrict__ vin,
                                                                 const void *__restrict__ position_id,
                                                                 int pos_dtype, T *__restrict__ out, int Smax,
                                                                 int nsplit, float *__restrict__ partial,
                                                                 const void *__restrict__ rope_pos,
                                                                 int rope_pos_dtype, int H) {
    pdl_trigger();
    pdl_wait();
    using C = RowCfg<T>;
    constexpr int EPL = C::EPL, LPR = C::LPR, RPW = C::RPW;
    __shared__ float s_m[WARPS], s_l[WARPS];
    __shared__ float s_acc[WARPS][kD];

    const int bh = blockIdx.x, split = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR;         // which of the RPW rows this lane works on
    const int col = (lane % LPR) * EPL; // first of this lane's EPL dims

    int pos = pos_dtype == ITB_I64 ? read_pos<int64_t>(position_id) : read_pos<int32_t>(position_id);
    if (pos < 0) pos = 0;
    if (pos >= Smax) pos = Smax - 1;
    const int seq = pos + 1;

    // this split's [s_begin, s_end)
    int chunk = (seq + nsplit - 1) / nsplit;
    chunk = ((chunk + RPW - 1) / RPW) * RPW;
    const int s_begin = split * chunk;
    const int s_end = min(seq, s_begin + chunk);

    T *kc = kcache + (int64_t)bh * Smax * kD;
    T *vc = vcache + (int64_t)bh * Smax * kD;
    const T *kn = kin + (int64_t)bh * kD;
    const T *vn = vin + (int64_t)bh * kD;

    Vec16<T> knew = ld16(kn + col);  // this lane's dims of the new k row
    Vec16<T> qv = ld16(q + (int64_t)bh * kD + col);
    if (ROPE) {
        const int b = bh / H;
        const float p = rope_pos_dtype == ITB_I64 ? (float)(int)((const int64_t *)rope_pos)[b]
                                                  : (float)((const int32_t *)rope_pos)[b];
        knew = rope_row<T, EPL, LPR>(knew, col, p);
        qv = rope_row<T, EPL, LPR>(qv, col, p);
    }

    // in-place append (one warp of the split that owns `pos`)
    if (pos >= s_begin && pos < s_end && warp == 0 && lane < LPR) {
        st16(kc + (int64_t)pos * kD + col, knew);
        st16(vc + (int64_t)pos * kD + col, ld16(vn + col));
    }

    float qf[EPL];
    {
#pragma unroll
        for (int j = 0; j < EPL; ++j) qf[j] = to_f(qv.v[j]) * 0.08838834764831845f;  // 1/sqrt(128)
    }

    float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) acc[j] = 0.f;

    // warp w takes row groups w, w+WARPS, ... ; each group = RPW*U rows
    for (int s0 = s_begin + warp * RPW * U; s0 < s_end; s0 += WARPS * RPW * U) {
        Vec16<T> kv[U], vv[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int s = s0 + u * RPW + sub;
            ok[u] = s < s_end;
            const T *kp = (s == pos) ? kn : kc + (int64_t)s * kD;  // the appended row comes from the k input
            const T *vp = (s == pos) ? vn : vc + (int64_t)s * kD;
            if (ok[u]) {
                kv[u] = ld16_stream(kp + col);
                vv[u] = ld16_stream(vp + col);
                if (ROPE && s == pos) kv[u] = knew;
            }
        }
        float sc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float d = 0.f;
            if (ok[u]) {
#pragma unroll
                for (int j = 0; j < EPL; ++j) d += qf[j] * to_f(kv[u].v[j]);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            sc[u] = ok[u] ? d : -INFINITY;
        }
        float mx = m;
#pragma unroll
        for (int u = 0; u < U; ++u) mx = fmaxf(mx, sc[u]);
        if (mx > -INFINITY) {
            float corr = expf(m - mx);  // m = -inf -> 0
            l *= corr;
#pragma unroll
            for (int j = 0; j < EPL; ++j) acc[j] *= corr;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ok[u]) {
                    float p = expf(sc[u] - mx);
                    l += p;
#pragma unroll
                    for (int j = 0; j < EPL; ++j) acc[j] = fmaf(p, to_f(vv[u].v[j]), acc[j]);
                }
            }
            m = mx;
        }
    }

    // merge the RPW sub-rows inside the warp (lanes with equal col, different sub)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        float m2 = __shfl_xor_sync(0xffffffffu, m, o);
        float l2 = __shfl_xor_sync(0xffffffffu, l, o);
        float mx = fmaxf(m, m2);
        float c1 = mx > -INFINITY ? expf(m - mx) : 0.f, c2 = mx > -INFINITY ? expf(m2 - mx) : 0.f;
        l = l * c1 + l2 * c2;
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            float a2 = __shfl_xor_sync(0xffffffffu, acc[j], o);
            acc[j] = acc[j] * c1 + a2 * c2;
        }
        m = mx;
    }
    // merge warps through shared memory
    if (sub == 0) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) s_acc[warp][col + j] = acc[j];
        if (lane == 0) {
            s_m[warp] = m;
            s_l[warp] = l;
        }
    }
    __syncthreads();
    if (threadIdx.x < kD) {
        int d = threadIdx.x;
        float mx = -INFINITY;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) mx = fmaxf(mx, s_m[w]);
        float L = 0.f, A = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            float c = s_m[w] > -INFINITY ? expf(s_m[w] - mx) : 0.f;
            L += s_l[w] * c;
            A += s_acc[w][d] * c;
        }
        if (nsplit == 1) {
            out[(int64_t)bh * kD + d] = from_f<T>(A / L);
        } else {
            float *pp = partial + ((int64_t)bh * nsplit + split) * (kD + 2);
            pp[2 + d] = A;
            if (d == 0) {
                pp[0] = mx;
                pp[1] = L;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kD) attn_merge_kernel(const float *__restrict__ partial, T *__restrict__ out,
                                                        int nsplit) {
    pdl_trigger();
    pdl_wait();
    int bh = blockIdx.x, d = threadIdx.x;
    const float *pp = partial + (int64_t)bh * nsplit * (kD + 2);
    float mx = -INFINITY;
    for (int s = 0; s < nsplit; ++s) mx = fmaxf(mx, pp[s * (kD + 2)]);
    float L = 0.f, A = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        float ms = pp[s * (kD + 2)];
        float c = ms > -INFINITY ? expf(ms - mx) : 0.f;
        L += pp[s * (kD + 2) + 1] * c;
        A += pp[s * (kD + 2) + 2 + d] * c;
    }
    out[(int64_t)bh * kD + d] = from_f<T>(A / L);
}

static int choose_nsplit(int BH, int Smax) {
    // enough CTAs for >= 2 per SM; never split below 64 positions per CTA
    int want = (2 * kNumSMs + BH - 1) / BH;
    int cap = Smax / 64 > 0 ? Smax / 64 : 1;
    int n = want < cap ? want : cap;
    return n < 1 ? 1 : n;
}

}  // namespace itb

using namespace itb;

extern "C" int64_t it_b200_attention_kvcache_workspace(int B, int H, int S_max, int D) {
    (void)D;
    int ns = choose_nsplit(B * H, S_max);
    return ns == 1 ? 0 : (int64_t)B * H * ns * (kD + 2) * sizeof(float);
}

static int attention_impl(int dtype, void *k_cache, void *v_cache, const void *q, const void *k, const void *v,
                          const void *position_id, int pos_dtype, const void *rope_pos, int rope_pos_dtype, void *out,
                          int B, int H, int S_max, int D, void *workspace, int64_t workspace_bytes, void *stream) {
    ITB_CHECK(D == kD, "AttentionKVCache: head dim %d != 128 (reference attention_kvcache.cu:154)", D);
    ITB_CHECK(pos_dtype == ITB_I32 || pos_dtype == ITB_U32 || pos_dtype == ITB_I64,
              "AttentionKVCache: position dtype %d must be int32/uint32/int64", pos_dtype);
    ITB_CHECK(!rope_pos || rope_pos_dtype == ITB_I32 || rope_pos_dtype == ITB_U32 || rope_pos_dtype == ITB_I64,
              "AttentionKVCache: rope position dtype %d must be int32/uint32/int64", rope_pos_dtype);
    ITB_CHECK(aligned16(k_cache) && aligned16(v_cache) && aligned16(q) && aligned16(k) && aligned16(v),
              "AttentionKVCache: tensors must be 16-byte aligned");
    int BH = B * H;
    if (BH == 0) return 0;
    int ns = choose_nsplit(BH, S_max);
    int64_t need = it_b200_attention_kvcache_workspace(B, H, S_max, D);
    ITB_CHECK(ns == 1 || (workspace && workspace_bytes >= need), "AttentionKVCache: workspace %lld < %lld bytes",
              (long long)workspace_bytes, (long long)need);
    auto st = (cudaStream_t)stream;
    dim3 grid(BH, ns);
    ITB_DISPATCH_FLOAT(dtype, "AttentionKVCache", {
        constexpr int WARPS = 8, U = 4;  // measured: U=2 at 64 regs (4 CTAs/SM) is 11 % slower end-to-end -- per-thread MLP wins
        if (rope_pos)
            launch_k(attn_decode_kernel<T, WARPS, U, true>, dim3(grid), dim3(WARPS * 32), 0, st, (T *)k_cache, (T *)v_cache,
                     (const T *)q, (const T *)k, (const T *)v, position_id, pos_dtype, (T *)out, S_max, ns,
                     (float *)workspace, rope_pos, rope_pos_dtype, H);
        else
            launch_k(attn_decode_kernel<T, WARPS, U, false>, dim3(grid), dim3(WARPS * 32), 0, st, (T *)k_cache, (T *)v_cache,
                     (const T *)q, (const T *)k, (const T *)v, position_id, pos_dtype, (T *)out, S_max, ns,
                     (float *)workspace, rope_pos, rope_pos_dtype, H);
        ITB_LAUNCH_CHECK("AttentionKVCache");
        if (ns > 1) {
            launch_k(attn_merge_kernel<T>, dim3(BH), dim3(kD), 0, st, (const float *)workspace, (T *)out, ns);
            ITB_LAUNCH_CHECK("AttentionKVCache.merge");
        }
    });
    return 0;
}

extern "C" int it_b200_attention_kvcache(int dtype, void *k_cache, void *v_cache, const void *q, const void *k,
                                         const void *v, const void *position_id, int pos_dtype, void *out, int B,
                                         int H, int S_max, int D, void *workspace, int64_t workspace_bytes,
                                         void *stream) {
    return attention_impl(dtype, k_cache, v_cache, q, k, v, position_id, pos_dtype, nullptr, 0, out, B, H, S_max, D,
                          workspace, workspace_bytes, stream);
}

extern "C" int it_b200_attention_kvcache_rope(int dtype, void *k_cache, void *v_cache, const void *q_pre,
                                              const void *k_pre, const void *v, const void *position_id, int pos_dtype,
                                              const void *rope_pos, int rope_pos_dtype, void *out, int B, int H,
                                              int S_max, int D, void *workspace, int64_t workspace_bytes, void *stream) {
    ITB_CHECK(rope_pos != nullptr, "AttentionKVCache+RoPE: rope positions missing");
    return attention_impl(dtype, k_cache, v_cache, q_pre, k_pre, v, position_id, pos_dtype, rope_pos, rope_pos_dtype, out,
                          B, H, S_max, D, workspace, workspace_bytes, stream);
}
