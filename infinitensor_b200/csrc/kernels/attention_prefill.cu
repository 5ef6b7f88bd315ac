// attention_prefill.cu -- fused multi-token (prefill) attention on the 5th-generation tensor core, sm_100a.
//
//   out[b,h] = softmax( (q[b,h] . k[b,h]^T) (/ or *) scale  + mask ) . v[b,h]        q, k, v, out: [B, H, S, D] f16 / bf16
//
// SURVEY 8(f-3).  The reference has no such kernel: its frontend lowers attention with q-len > 1 to
// Transpose -> MatMul -> Div -> Add(mask) -> Softmax -> MatMul (examples/python + pyinfinitensor/onnx.py; config C2, GPT-2),
// five launches of matmul.cc:66-211 / element_wise.cc / softmax.cu:242-404 per head group plus the [B,H,S,S] score tensor's
// round trips through HBM.  This kernel is the whole chain for one (batch, head, 128-query tile) per CTA:
//   * Q, K, V tiles arrive by TMA (3-D tensor maps over [B*H, S, D], 128B swizzle) into shared memory;
//   * S = Q K^T runs as tcgen05.mma (kind::f16, M = 128 queries, N = 128 keys, K = 16 per instruction, both operands K-major)
//     into a [128 lanes x 128 columns] fp32 accumulator in TENSOR MEMORY;
//   * 512 threads: warp w reads TMEM lanes [32 (w & 3), +32) = query rows, and the four warps sharing a lane quadrant split each
//     128-key tile into 32-column parts (w >> 2) -- one warp per scheduler, as in the first version (128 threads, a whole row
//     each), left every dependent exp / round chain exposed: 40 us per (head, tile) at GPT-2's size; the parts' running
//     (max, sum) meet once in shared memory after pass A.  Each thread reads its part of S back with tcgen05.ld, applies the graph's own
//     rounding points (MatMul output, Div / Mul by the scalar, Add mask -- each rounded to the storage type like the separate
//     kernels), and the softmax: pass A accumulates the row maximum / sum over all key tiles (online), pass B recomputes S,
//     writes P = exp(s - max) / sum (rounded like the Softmax kernel's output) as the K-major A operand into shared memory
//     and the tensor core accumulates O += P V (V consumed as the MN-major B operand, straight from its [S, D] layout);
//   * O leaves TMEM once, rounded to the storage type.
// Two passes over K instead of rescaling O in TMEM: the extra QK^T costs 2 S^2 D flops on an idle tensor pipe, and the
// result needs no correction step -- the probabilities are final when they meet V.
// No causal shortcut: the mask is whatever tensor the graph adds (broadcast strides), so every graph the frontend emits is
// reproduced, not only triangular ones.
#include <cstring>

#include "gemm.cuh"

namespace itb {

constexpr int AP_BQ = 128, AP_BK = 128;
constexpr int AP_PARTS = 4, AP_PCOLS = AP_BK / AP_PARTS, AP_THREADS = 128 * AP_PARTS;
constexpr int AP_MPAD = AP_BQ + 2;  // 65 words per key column: the transposing store hits 32 distinct banks

struct ApArgs {
    int BH, H, Sq, Skv, D;
    const void *scale;      // device scalar (nullptr: none)
    int scale_is_div;       // 1: s / scale, 0: s * scale
    const void *mask;       // additive mask (nullptr: none), element (b, h, i, j) at mask[b*mb + h*mh + i*mi + j*mj]
    int64_t mb, mh, mi, mj;
    void *out;              // element (b, h, i, d) at out[b*ob + h*oh + i*os + d]  ([B, H, S, D] dense: ob = H*S*D, oh = S*D, os = D)
    int64_t ob, oh, os;
    // q / k / v tensor maps: 0 = 3-D over a dense [B*H, S, D]; 1 = 4-D with dimensions (d, h, s, b); 2 = 4-D (d, s, h, b) -- strided
    // views of e.g. a fused q/k/v projection output [B, S, 3 H D] (it_b200_attention_prefill_strided)
    int map_mode;
};

__device__ __forceinline__ void tma_load_4d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
          "l"(policy)
        : "memory");
}

template <typename T, int D>
__global__ void __launch_bounds__(AP_THREADS, 1) attention_prefill_kernel(const __grid_constant__ CUtensorMap mapQ,
                                                                   const __grid_constant__ CUtensorMap mapK,
                                                                   const __grid_constant__ CUtensorMap mapV, const ApArgs a) {
    constexpr int DG = D / 64;                    // 64-column groups of the head dim
    constexpr int QK_BYTES = AP_BQ * D * 2;       // one Q or K tile (DG boxes of [128 rows x 64 cols])
    constexpr int V_BYTES = AP_BK * D * 2;        // one V tile (DG x 2 boxes of [64 kv rows x 64 cols])
    constexpr int P_BYTES = AP_BQ * AP_BK * 2;    // probabilities, two k-blocks of [128 rows x 64 keys]
    extern __shared__ uint8_t ap_smem_raw[];
    uint8_t *smem = ap_smem_raw + ((1024u - (smem_u32(ap_smem_raw) & 1023u)) & 1023u);
    uint8_t *q_sm = smem, *k_sm = q_sm + QK_BYTES, *v_sm = k_sm + QK_BYTES, *p_sm = v_sm + V_BYTES;
    // the mask tile of the running (query tile, key tile), TRANSPOSED and padded: m_sm[j * AP_MPAD + i] -- every thread reads its own
    // query row i for consecutive keys j (lanes = consecutive i: conflict-free), the cooperative fill writes with lanes = keys
    T *m_sm = reinterpret_cast<T *>(p_sm + P_BYTES);
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float red_m[AP_PARTS][AP_BQ], red_l[AP_PARTS][AP_BQ];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y, q0 = blockIdx.x * AP_BQ;
    const int quad = warp & 3, part = warp >> 2;
    const int row = quad * 32 + lane;  // query row inside the tile == TMEM lane
    const int pc0 = part * AP_PCOLS;   // this thread's columns of every key tile: [pc0, pc0 + 32)
    pdl_trigger();
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapV);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = tmem_slot, tmem_o = tmem_slot + 128;
    pdl_wait();

    const uint32_t idesc_s = umma_idesc_f16(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 0, 0, 128, AP_BK);
    const uint32_t idesc_o = umma_idesc_f16(std::is_same<T, __nv_bfloat16>::value ? 1 : 0, 0, /*B = V, MN-major*/ 1, 128, D);
    const uint64_t pol = l2_policy_evict_last();
    uint32_t ph_load = 0, ph_mma = 0;
    const int b = bh / a.H, h = bh % a.H;
    // one [rows x 64 columns] box of q / k / v: (column, sequence row) of head (b, h), whatever the tensor map's dimension order
    auto ld = [&](void *dst, const CUtensorMap *m, int c_d, int c_s) {
        if (a.map_mode == 0) tma_load_3d(dst, m, &bar_load, c_d, c_s, bh, pol);
        else if (a.map_mode == 1) tma_load_4d(dst, m, &bar_load, c_d, h, c_s, b, pol);
        else tma_load_4d(dst, m, &bar_load, c_d, c_s, h, b, pol);
    };

    // Q tile once -- together with the first K tile (and V when there is only one key tile): one round trip instead of three
    const int ntiles = (a.Skv + AP_BK - 1) / AP_BK;
    const bool single = ntiles == 1;  // one key tile (GPT-2's S = 128): K, V and the mask tile stay in shared memory for both passes
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_load, QK_BYTES + (ntiles > 0 ? QK_BYTES : 0) + (single ? V_BYTES : 0));
#pragma unroll
        for (int g = 0; g < DG; ++g) ld(q_sm + g * (AP_BQ * 128), &mapQ, g * 64, q0);
        if (ntiles > 0)
#pragma unroll
            for (int g = 0; g < DG; ++g) ld(k_sm + g * (AP_BK * 128), &mapK, g * 64, 0);
        if (single)
#pragma unroll
            for (int g = 0; g < DG; ++g)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    ld(v_sm + g * (AP_BK * 128) + hh * (64 * 128), &mapV, g * 64, hh * 64);
    }

    float sc = 1.f;
    if (a.scale) sc = to_f(*(const T *)a.scale);
    // s / 2^k == s * 2^-k exactly (GPT-2: sqrt(64)): spares a division per score and pass
    const bool div_pow2 = a.scale && a.scale_is_div && (__float_as_uint(sc) & 0x007fffffu) == 0u && sc != 0.f && fabsf(sc) < 1e30f && fabsf(sc) > 1e-30f;
    const float sc_mul = a.scale ? (a.scale_is_div ? (div_pow2 ? 1.f / sc : 0.f) : sc) : 1.f;
    const bool use_mul = a.scale && (!a.scale_is_div || div_pow2);
    const T *mbase = a.mask ? (const T *)a.mask + b * a.mb + h * a.mh : nullptr;
    const bool mvec = mbase && a.mj == 1 && (a.mi & 7) == 0 && (((uintptr_t)mbase & 15) == 0);
    // cooperative, coalesced fill of the mask tile (rows q0.., keys j0..): 16 threads cover one row's 128 keys with 16-byte loads
    auto load_mask = [&](int j0) {
        if (!mbase) return;
        const int jc = (threadIdx.x & 15) * 8, r0 = threadIdx.x >> 4;  // 32 rows per sweep
        // all four sweeps' loads in flight before the first transposing store (the stores may alias the mask as far as the
        // compiler knows: issued row by row, the dependent L2 round trips alone cost ~16 us per tile -- measured)
        {
            constexpr int NSW = AP_BQ / (AP_THREADS / 16);
            uint4 buf[NSW];
#pragma unroll
            for (int i = 0; i < NSW; ++i) {
                const int r = r0 + (AP_THREADS / 16) * i;
                const bool rok = q0 + r < a.Sq;
                const T *src = mbase + (int64_t)(q0 + r) * a.mi + (int64_t)(j0 + jc) * a.mj;
                if (rok && mvec && j0 + jc + 8 <= a.Skv && ((((int64_t)(q0 + r) * a.mi + j0 + jc) & 7) == 0)) {
                    buf[i] = __ldg(reinterpret_cast<const uint4 *>(src));
                } else {
                    T vals[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) vals[e] = (rok && j0 + jc + e < a.Skv) ? src[(int64_t)e * a.mj] : from_f<T>(0.f);
                    buf[i] = *reinterpret_cast<const uint4 *>(vals);
                }
            }
#pragma unroll
            for (int i = 0; i < NSW; ++i) {
                const int r = r0 + (AP_THREADS / 16) * i;
                const T *vals = reinterpret_cast<const T *>(&buf[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) m_sm[(jc + e) * AP_MPAD + r] = vals[e];
            }
        }
    };
    const bool row_ok = q0 + row < a.Sq;

    // score of (this row, key j0 + c) from the raw accumulator value, with the graph's rounding points
    auto score = [&](float acc, int j) -> float {
        float s = round_t<T>(acc);                                      // MatMul output
        if (a.scale) s = round_t<T>(use_mul ? s * sc_mul : s / sc);  // Div / Mul by the scalar constant
        if (mbase && row_ok) s = round_t<T>(s + to_f(m_sm[(j % AP_BK) * AP_MPAD + row]));  // Add(mask)
        return s;
    };
    auto load_k = [&](int t) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar_load, QK_BYTES);
#pragma unroll
            for (int g = 0; g < DG; ++g) ld(k_sm + g * (AP_BK * 128), &mapK, g * 64, t * AP_BK);
        }
    };
    auto mma_s = [&]() {  // S = Q K^T into tmem_s
        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t qb = smem_u32(q_sm), kb = smem_u32(k_sm);
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk) {
                const uint32_t off = (kk >> 2) * (128 * 128) + (kk & 3) * 32;
                tc_mma_f16(tmem_s, umma_desc_sw128(qb + off, 0, 1024), umma_desc_sw128(kb + off, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
            }
            tc_commit(&bar_mma);
        }
    };

    // ---------------- pass A: row maximum and sum over all keys ----------------
    float m = -INFINITY, l = 0.f;
    if (ntiles == 0) {  // no keys: only Q was requested
        mbar_wait(&bar_load, ph_load);
        ph_load ^= 1;
    }
    for (int t = 0; t < ntiles; ++t) {
        if (t > 0) load_k(t);  // (tile 0 came with Q)
        const int j0 = t * AP_BK;
        load_mask(j0);  // (the previous tile's readers are behind the __syncthreads that closed it)
        mbar_wait(&bar_load, ph_load);
        ph_load ^= 1;
        mma_s();
        __syncthreads();  // mask tile complete
        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
#pragma unroll 1
        for (int c0 = pc0; c0 < pc0 + AP_PCOLS; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_s + ((uint32_t)(quad * 32) << 16) + c0, v);
            float sv[16], mx = m;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                sv[j] = j0 + c0 + j < a.Skv ? score(__uint_as_float(v[j]), j0 + c0 + j) : -INFINITY;
                mx = fmaxf(mx, sv[j]);
            }
            if (mx > -INFINITY) {
                float add = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) add += expf(sv[j] - mx);  // exp(-inf) = 0
                l = l * expf(m - mx) + add;
                m = mx;
            }
        }
        tc_fence_before();
        __syncthreads();  // every row has read S before the next tile's MMA overwrites it (and K is reloaded)
    }
    // the four column parts of a row meet: global maximum, sums rescaled to it (same order for every part -> identical values)
    red_m[part][row] = m;
    red_l[part][row] = l;
    __syncthreads();
    {
        float mg = -INFINITY;
#pragma unroll
        for (int pp = 0; pp < AP_PARTS; ++pp) mg = fmaxf(mg, red_m[pp][row]);
        float lg = 0.f;
#pragma unroll
        for (int pp = 0; pp < AP_PARTS; ++pp) {
            const float mp = red_m[pp][row];
            if (mp > -INFINITY) lg += red_l[pp][row] * expf(mp - mg);
        }
        m = mg;
        l = lg;
    }
    const float inv_l = l > 0.f ? 1.f / l : 0.f;

    // ---------------- pass B: P = softmax row (final), O += P V ----------------
    for (int t = 0; t < ntiles; ++t) {
        if (threadIdx.x == 0 && !single) {
            mbar_expect_tx(&bar_load, QK_BYTES + V_BYTES);
#pragma unroll
            for (int g = 0; g < DG; ++g) ld(k_sm + g * (AP_BK * 128), &mapK, g * 64, t * AP_BK);
#pragma unroll
            for (int g = 0; g < DG; ++g)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
                    ld(v_sm + g * (AP_BK * 128) + hh * (64 * 128), &mapV, g * 64, t * AP_BK + hh * 64);
        }
        const int j0 = t * AP_BK;
        if (!single) {
            load_mask(j0);  // (every reader of the previous tile's mask passed the __syncthreads before that tile's P.V)
            mbar_wait(&bar_load, ph_load);
            ph_load ^= 1;
        }
        if (!single) {  // (one key tile: S of pass A is still in tensor memory)
            mma_s();
            __syncthreads();  // mask tile complete
            mbar_wait(&bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
        }
#pragma unroll 1
        for (int c0 = pc0; c0 < pc0 + AP_PCOLS; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_s + ((uint32_t)(quad * 32) << 16) + c0, v);
            T pv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float pr = 0.f;
                if (j0 + c0 + j < a.Skv && m > -INFINITY) pr = expf(score(__uint_as_float(v[j]), j0 + c0 + j) - m) * inv_l;
                pv[j] = from_f<T>(pr);  // the Softmax kernel's output rounding
            }
            // P is the K-major A operand of the second MMA: k-block = 64 keys, rows of 128 B, 128B swizzle
            uint8_t *blk = p_sm + (c0 >> 6) * (AP_BQ * 128) + row * 128;
            const int chunk = (c0 & 63) >> 3;  // first of the two 16-byte chunks these 16 keys fill
            *reinterpret_cast<uint4 *>(blk + (((chunk) ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4 *>(&pv[0]);
            *reinterpret_cast<uint4 *>(blk + (((chunk + 1) ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4 *>(&pv[8]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
            const uint32_t pb = smem_u32(p_sm), vb = smem_u32(v_sm);
#pragma unroll
            for (int kk = 0; kk < AP_BK / 16; ++kk) {
                // A = P: k-block (kk >> 2), 32 B per k16 step inside the 128 B row;  B = V (MN-major): column groups V_BYTES / DG
                // apart (LBO), 8-row key groups 1 KB apart (SBO), one k16 step = 16 key rows = 2 KB
                const uint64_t ad = umma_desc_sw128(pb + (kk >> 2) * (AP_BQ * 128) + (kk & 3) * 32, 0, 1024);
                const uint64_t bd = umma_desc_sw128(vb + kk * 2048, AP_BK * 128, 1024);
                tc_mma_f16(tmem_o, ad, bd, idesc_o, (t > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, ph_mma);  // P, K, V are free again (and on the last tile: O is complete)
        ph_mma ^= 1;
        tc_fence_after();
    }

    // ---------------- epilogue: O row -> out ----------------
    T *orow = (T *)a.out + (int64_t)b * a.ob + (int64_t)h * a.oh + (int64_t)(q0 + row) * a.os;
#pragma unroll 1
    for (int c0 = part * (D / AP_PARTS); c0 < (part + 1) * (D / AP_PARTS); c0 += 16) {
        uint32_t v[16];
        if (ntiles > 0) tmem_ld_32x32b_x16(tmem_o + ((uint32_t)(quad * 32) << 16) + c0, v);
        T ov[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) ov[j] = from_f<T>(ntiles > 0 ? __uint_as_float(v[j]) : 0.f);
        if (row_ok) {
            *reinterpret_cast<uint4 *>(orow + c0) = *reinterpret_cast<const uint4 *>(&ov[0]);
            *reinterpret_cast<uint4 *>(orow + c0 + 8) = *reinterpret_cast<const uint4 *>(&ov[8]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_slot, 256);
    }
    (void)lane;
}

// strides of a q / k / v view in elements: {batch, head, sequence row}; the head dimension is contiguous
struct ApStrides {
    int64_t b, h, s;
};

template <typename T, int D>
static int launch_ap(const void *q, const void *k, const void *v, ApArgs a, cudaStream_t st, const ApStrides *sq = nullptr,
                     const ApStrides *sk = nullptr, const ApStrides *sv = nullptr) {
    CUtensorMap mq, mk, mv;
    if (!sq) {
        a.map_mode = 0;
        if (!make_tma_3d_b16(&mq, q, (uint64_t)a.BH, (uint64_t)a.Sq, (uint64_t)D, (uint64_t)D, (uint64_t)a.Sq * D, AP_BQ, 64) ||
            !make_tma_3d_b16(&mk, k, (uint64_t)a.BH, (uint64_t)a.Skv, (uint64_t)D, (uint64_t)D, (uint64_t)a.Skv * D, AP_BK, 64) ||
            !make_tma_3d_b16(&mv, v, (uint64_t)a.BH, (uint64_t)a.Skv, (uint64_t)D, (uint64_t)D, (uint64_t)a.Skv * D, 64, 64))
            ITB_FAIL("attention_prefill: cuTensorMapEncodeTiled failed");
    } else {
        // 4-D maps, dimensions ordered by stride (TMA wants them ascending): (d, h, s, b) when heads are the inner stride (a fused
        // projection's [B, S, H D] rows), (d, s, h, b) for [B, H, S, D]
        const int B = a.BH / a.H;
        const bool h_inner = sq->h < sq->s;
        ITB_CHECK((sk->h < sk->s) == h_inner && (sv->h < sv->s) == h_inner, "attention_prefill: q / k / v views must share a dimension order");
        a.map_mode = h_inner ? 1 : 2;
        auto mk4 = [&](CUtensorMap *m, const void *base, const ApStrides &t, int S, uint32_t box_s) {
            return h_inner ? make_tma_4d_b16(m, base, (uint64_t)D, (uint64_t)a.H, (uint64_t)S, (uint64_t)B, (uint64_t)t.h, (uint64_t)t.s,
                                             (uint64_t)t.b, 64, 1, box_s, 1)
                           : make_tma_4d_b16(m, base, (uint64_t)D, (uint64_t)S, (uint64_t)a.H, (uint64_t)B, (uint64_t)t.s, (uint64_t)t.h,
                                             (uint64_t)t.b, 64, box_s, 1, 1);
        };
        if (!mk4(&mq, q, *sq, a.Sq, AP_BQ) || !mk4(&mk, k, *sk, a.Skv, AP_BK) || !mk4(&mv, v, *sv, a.Skv, 64))
            ITB_FAIL("attention_prefill: cuTensorMapEncodeTiled (strided views) failed");
    }
    const int smem = 2 * AP_BQ * D * 2 + AP_BK * D * 2 + AP_BQ * AP_BK * 2 + AP_BK * AP_MPAD * 2 + 1024;
    auto kern = attention_prefill_kernel<T, D>;
    static int attr_smem[64] = {0};  // per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    cudaError_t e = cudaSuccess;
    if (smem > attr_smem[dev]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        ITB_CHECK(e == cudaSuccess, "attention_prefill: smem attribute: %s", cudaGetErrorString(e));
        attr_smem[dev] = smem;
    }
    e = launch_k(kern, dim3((a.Sq + AP_BQ - 1) / AP_BQ, a.BH), dim3(AP_THREADS), (size_t)smem, st, mq, mk, mv, a);
    ITB_CHECK(e == cudaSuccess, "attention_prefill: launch failed: %s", cudaGetErrorString(e));
    ITB_LAUNCH_CHECK("attention_prefill");
    return 0;
}

}  // namespace itb

using namespace itb;

extern "C" int it_b200_attention_prefill(int dtype, const void *q, const void *k, const void *v, void *out, int B, int H, int S_q,
                                         int S_kv, int D, const void *scale, int scale_is_div, const void *mask,
                                         int64_t mask_stride_b, int64_t mask_stride_h, int64_t mask_stride_i,
                                         int64_t mask_stride_j, void *stream) {
    ITB_CHECK(dtype == ITB_F16 || dtype == ITB_BF16, "attention_prefill: dtype %d must be f16 / bf16", dtype);
    ITB_CHECK(D == 64 || D == 128, "attention_prefill: head dim %d must be 64 or 128", D);
    ITB_CHECK(B >= 0 && H >= 0 && S_q >= 0 && S_kv >= 0, "attention_prefill: negative dimension");
    ITB_CHECK((int64_t)B * H <= 65535, "attention_prefill: B * H beyond the grid limit");
    if ((int64_t)B * H * S_q == 0) return 0;
    ITB_CHECK(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "attention_prefill: tensors must be 16-byte aligned");
    ApArgs a{};
    a.BH = B * H;
    a.H = H;
    a.Sq = S_q;
    a.Skv = S_kv;
    a.D = D;
    a.scale = scale;
    a.scale_is_div = scale_is_div;
    a.mask = mask;
    a.mb = mask_stride_b;
    a.mh = mask_stride_h;
    a.mi = mask_stride_i;
    a.mj = mask_stride_j;
    a.out = out;
    a.ob = (int64_t)H * S_q * D;
    a.oh = (int64_t)S_q * D;
    a.os = D;
    auto st = (cudaStream_t)stream;
    if (dtype == ITB_BF16) return D == 64 ? launch_ap<__nv_bfloat16, 64>(q, k, v, a, st) : launch_ap<__nv_bfloat16, 128>(q, k, v, a, st);
    return D == 64 ? launch_ap<__half, 64>(q, k, v, a, st) : launch_ap<__half, 128>(q, k, v, a, st);
}

// The same attention over STRIDED views: element (b, h, i, d) of q at q[b*qs[0] + h*qs[1] + i*qs[2] + d] (likewise k, v, out) -- the
// frontend's Split -> Reshape -> Transpose([0,2,1,3]) of a fused q/k/v projection and the Transpose -> Reshape after the attention
// become addressing (4-D tensor maps, strided output rows) instead of five copy kernels per layer.
extern "C" int it_b200_attention_prefill_strided(int dtype, const void *q, const void *k, const void *v, void *out, int B, int H,
                                                 int S_q, int S_kv, int D, const int64_t *q_strides, const int64_t *k_strides,
                                                 const int64_t *v_strides, const int64_t *out_strides, const void *scale,
                                                 int scale_is_div, const void *mask, int64_t mask_stride_b, int64_t mask_stride_h,
                                                 int64_t mask_stride_i, int64_t mask_stride_j, void *stream) {
    ITB_CHECK(dtype == ITB_F16 || dtype == ITB_BF16, "attention_prefill: dtype %d must be f16 / bf16", dtype);
    ITB_CHECK(D == 64 || D == 128, "attention_prefill: head dim %d must be 64 or 128", D);
    ITB_CHECK(B >= 0 && H >= 0 && S_q >= 0 && S_kv >= 0, "attention_prefill: negative dimension");
    ITB_CHECK((int64_t)B * H <= 65535, "attention_prefill: B * H beyond the grid limit");
    if ((int64_t)B * H * S_q == 0) return 0;
    ITB_CHECK(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out), "attention_prefill: tensors must be 16-byte aligned");
    for (int i = 0; i < 3; ++i)
        ITB_CHECK(q_strides[i] % 8 == 0 && k_strides[i] % 8 == 0 && v_strides[i] % 8 == 0 && out_strides[i] % 8 == 0 && q_strides[i] > 0 &&
                      k_strides[i] > 0 && v_strides[i] > 0,
                  "attention_prefill: view strides must be positive multiples of 8 elements");
    ApArgs a{};
    a.BH = B * H;
    a.H = H;
    a.Sq = S_q;
    a.Skv = S_kv;
    a.D = D;
    a.scale = scale;
    a.scale_is_div = scale_is_div;
    a.mask = mask;
    a.mb = mask_stride_b;
    a.mh = mask_stride_h;
    a.mi = mask_stride_i;
    a.mj = mask_stride_j;
    a.out = out;
    a.ob = out_strides[0];
    a.oh = out_strides[1];
    a.os = out_strides[2];
    const ApStrides sq{q_strides[0], q_strides[1], q_strides[2]}, sk{k_strides[0], k_strides[1], k_strides[2]},
        sv{v_strides[0], v_strides[1], v_strides[2]};
    auto st = (cudaStream_t)stream;
    if (dtype == ITB_BF16)
        return D == 64 ? launch_ap<__nv_bfloat16, 64>(q, k, v, a, st, &sq, &sk, &sv) : launch_ap<__nv_bfloat16, 128>(q, k, v, a, st, &sq, &sk, &sv);
    return D == 64 ? launch_ap<__half, 64>(q, k, v, a, st, &sq, &sk, &sv) : launch_ap<__half, 128>(q, k, v, a, st, &sq, &sk, &sv);
}
