// gemm_streamk.cu -- persistent stream-K variant of the decode GEMM (M <= 64), sm_100a.
//
//   {C_i[M,N_i]} = X[M,K] . {W_i[K,N_i]}   (1..4 weight matrices sharing X; bf16 / fp16, fp32 accumulate)
//
// Why: the tile-per-CTA kernel (gemm_skinny.cu) quantises badly -- a decode GEMM is only 64..500 column tiles,
// so 148 SMs x 2 CTAs are rarely filled evenly and the measured bandwidth swings between 48 % and 86 % of peak
// with the shape (profiles/r01*_gemm_bench.md).  Here the work is cut at k-tile granularity instead:
//   * unit = (column tile of 64, k-tile of 64); U = sum_i ceil(N_i/64) * ceil(K/64) units
//   * a fixed grid of G = 148 x occupancy CTAs; CTA c owns the contiguous unit range [c*U/G, (c+1)*U/G) --
//     every CTA streams the same number of bytes (+-1 tile of 8 KB), whatever N and K are
//   * one TMA producer lane keeps a STAGES-deep ring full ACROSS tile boundaries (no pipeline drain per tile);
//     4 consumer warps run mma.sync m16n8k16 from the 128B-swizzled tiles
//   * a segment that covers a whole column tile is finished from registers; partial segments (at most two per
//     CTA) are written to a per-CTA fp32 slot, and the LAST contributor of each tile -- found with one atomic
//     ticket, no spinning -- sums the slots in CTA order (deterministic), applies bias / residual / activation
//     and stores.  The ticket counter is reset by that CTA, so the scratch is self-cleaning.
//   * PDL: weight tiles of the first ring are requested ahead of griddepcontrol.wait when B is constant.
#include <map>
#include <mutex>

#include "gemm.cuh"

namespace itb {

constexpr int KS_BN = 64, KS_BK = 64;
constexpr int KS_W_BYTES = KS_BN * KS_BK * 2;
constexpr int KS_THREADS = 160;  // 4 consumer warps + 1 producer warp
constexpr int KS_MAX_GROUPS = 4;
constexpr int KS_MAX_CTAS = 148 * 2;  // stream-K grid; the (tile, split) mode may launch up to 4x this
constexpr int KS_MAX_TILES = 4096;

struct StreamKGroups {
    CUtensorMap mapW[KS_MAX_GROUPS];
    void *C[KS_MAX_GROUPS];
    int n[KS_MAX_GROUPS];
    int tile_start[KS_MAX_GROUPS + 1];
    int ngroups;
};

template <int MT> struct StreamKCfg {
    static constexpr int X_BYTES = MT * 16 * KS_BK * 2;
    static constexpr int STAGES = MT == 1 ? 10 : (MT == 2 ? 8 : 6);
    static constexpr int RED_FLOATS = MT * 16 * KS_BN;
    static constexpr int SMEM = STAGES * (KS_W_BYTES + X_BYTES) + RED_FLOATS * 4 + 2 * STAGES * 8 + 64 + 1024;
};

__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <typename T, int MT>
__global__ void __launch_bounds__(KS_THREADS) gemm_streamk_kernel(const __grid_constant__ StreamKGroups grp,
                                                                  const __grid_constant__ CUtensorMap mapX,
                                                                  GemmArgs g, int ktiles, long long U, int tiles_mode,
                                                                  float *__restrict__ slots,
                                                                  int *__restrict__ tickets) {
    using Cfg = StreamKCfg<MT>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *w_sm = smem;
    uint8_t *x_sm = smem + S * KS_W_BYTES;
    float *red = reinterpret_cast<float *>(x_sm + S * Cfg::X_BYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(red + Cfg::RED_FLOATS);
    uint64_t *empty = full + S;
    int *s_flag = reinterpret_cast<int *>(empty + S);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, c = blockIdx.x;
    // tiles_mode == 0: stream-K (contiguous unit ranges).  tiles_mode = T > 0: classic (column tile, k split) grid --
    // CTA c owns tile c % T and split c / T, so CTAs that run together sweep ADJACENT column tiles at the same k rows
    // (DRAM-page friendly for row-major [K,N] weights); partial tiles still meet through the ticketed slot reduction.
    int u0, u1, nsplit_t = 1, per_t = ktiles;
    if (tiles_mode > 0) {
        nsplit_t = G / tiles_mode;
        per_t = (ktiles + nsplit_t - 1) / nsplit_t;
        const int t = c % tiles_mode, sp = c / tiles_mode;
        u0 = t * ktiles + sp * per_t;
        u1 = min(u0 + per_t, (t + 1) * ktiles);
        if (u1 < u0) u1 = u0;
    } else {
        u0 = (int)((long long)c * U / G);
        u1 = (int)((long long)(c + 1) * U / G);
    }
    const int nunits = u1 - u0;
    const int tile0 = u0 / ktiles, kt0 = u0 - tile0 * ktiles;  // the only divisions: once per CTA

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        fence_mbar_init();
    }
    __syncthreads();

    auto locate = [&](int tile, int &gi, int &n0) {
        gi = 0;
#pragma unroll
        for (int i = 1; i < KS_MAX_GROUPS; ++i)
            if (i < grp.ngroups && tile >= grp.tile_start[i]) gi = i;
        n0 = (tile - grp.tile_start[gi]) * KS_BN;
    };

    if (warp == 4) {
        // ===== TMA producer: one lane streams this CTA's unit range (tile / k-tile / stage tracked incrementally) =====
        if (lane == 0) {
            const uint64_t pol_w = l2_policy_evict_first(), pol_x = l2_policy_evict_last();
            const bool w_const = (g.act & ITB_MATMUL_B_CONST) != 0;
            const int pre = nunits < S ? nunits : S;
            int tile = tile0, kt = kt0, gi, n0;
            locate(tile, gi, n0);
            if (!w_const) pdl_wait();
            for (int it = 0; it < pre; ++it) {  // first ring: weights may go ahead of the PDL wait
                mbar_expect_tx(&full[it], KS_W_BYTES + Cfg::X_BYTES);
                tma_load_2d(w_sm + it * KS_W_BYTES, &grp.mapW[gi], &full[it], n0, kt * KS_BK, pol_w);
                if (++kt == ktiles) {
                    kt = 0;
                    locate(++tile, gi, n0);
                }
            }
            if (w_const) pdl_wait();
            {
                int kx = kt0;
                for (int it = 0; it < pre; ++it) {
                    tma_load_2d(x_sm + it * Cfg::X_BYTES, &mapX, &full[it], kx * KS_BK, 0, pol_x);
                    if (++kx == ktiles) kx = 0;
                }
            }
            int s = 0;
            uint32_t phase = 0;  // parity of the `empty` completion that frees stage s for its next use
            for (int it = pre; it < nunits; ++it) {
                mbar_wait(&empty[s], phase);
                mbar_expect_tx(&full[s], KS_W_BYTES + Cfg::X_BYTES);
                tma_load_2d(w_sm + s * KS_W_BYTES, &grp.mapW[gi], &full[s], n0, kt * KS_BK, pol_w);
                tma_load_2d(x_sm + s * Cfg::X_BYTES, &mapX, &full[s], kt * KS_BK, 0, pol_x);
                if (++kt == ktiles) {
                    kt = 0;
                    locate(++tile, gi, n0);
                }
                if (++s == S) {
                    s = 0;
                    phase ^= 1;
                }
            }
        }
        __syncwarp();
        return;  // the producer warp takes no part in the consumer-only barriers below
    }

    // ===== consumers: warp w owns columns [16w, 16w+16) of the current tile =====
    pdl_wait();
    const int mi = lane >> 3, r8 = lane & 7;
    const uint32_t w_base = smem_u32(w_sm), x_base = smem_u32(x_sm);
    const int tid = threadIdx.x;  // 0..127
    const T *bias = (const T *)g.bias;
    const bool round_first = (g.act & ITB_ACT_ROUND_BEFORE_BIAS) != 0;
    const int first_tile = tile0;

    float acc[MT][2][4];
    auto zero_acc = [&]() {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[mt][nb][i] = 0.f;
    };
    zero_acc();
    int tile = tile0, kt = kt0, seg_kt = kt0, s = 0;
    uint32_t phase = 0;

    for (int it = 0; it < nunits; ++it) {
        mbar_wait(&full[s], phase);
        const uint32_t wb = w_base + s * KS_W_BYTES, xb = x_base + s * Cfg::X_BYTES;
#pragma unroll
        for (int kk = 0; kk < KS_BK / 16; ++kk) {
            uint32_t a[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int row = mt * 16 + r8 + 8 * (mi & 1);
                const int chunk = kk * 2 + (mi >> 1);
                ldmatrix_x4(a[mt][0], a[mt][1], a[mt][2], a[mt][3], xb + row * 128 + ((chunk ^ (row & 7)) << 4));
            }
            uint32_t b0, b1, b2, b3;
            {
                const int krow = kk * 16 + r8 + 8 * (mi & 1);
                const int nchunk = warp * 2 + (mi >> 1);
                ldmatrix_x4_trans(b0, b1, b2, b3, wb + krow * 128 + ((nchunk ^ (krow & 7)) << 4));
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                mma_m16n8k16<T>(acc[mt][0], a[mt], b0, b1);
                mma_m16n8k16<T>(acc[mt][1], a[mt], b2, b3);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == S) {
            s = 0;
            phase ^= 1;
        }
        const int this_tile = tile, this_kt = kt;
        if (++kt == ktiles) {
            kt = 0;
            ++tile;
        }
        const bool seg_end = (it + 1 == nunits) || (kt == 0);
        if (!seg_end) continue;
        const bool whole = (seg_kt == 0) && (this_kt == ktiles - 1) && (tiles_mode == 0 || nsplit_t == 1);
        seg_kt = kt;
        (void)this_kt;

        // ---------------- flush the finished segment of column tile `this_tile` ----------------
        int gi, n0;
        locate(this_tile, gi, n0);
        const int gN = grp.n[gi];
        T *C = (T *)grp.C[gi];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
                const int row = mt * 16 + (lane >> 2);
                const int col = warp * 16 + nb * 8 + (lane & 3) * 2;
                *reinterpret_cast<float2 *>(&red[row * KS_BN + col]) = make_float2(acc[mt][nb][0], acc[mt][nb][1]);
                *reinterpret_cast<float2 *>(&red[(row + 8) * KS_BN + col]) = make_float2(acc[mt][nb][2], acc[mt][nb][3]);
            }
        bar_consumers();

        bool finish = whole;  // do I run the epilogue of this tile?
        int c_first = c, c_last = c;
        if (!whole) {
            // publish my partial, take a ticket; the last contributor finishes the tile
            const int j = (tiles_mode > 0 || this_tile == first_tile) ? 0 : 1;
            float *my = slots + ((size_t)c * 2 + j) * Cfg::RED_FLOATS;
            for (int i = tid * 4; i < Cfg::RED_FLOATS; i += 128 * 4)
                *reinterpret_cast<float4 *>(my + i) = *reinterpret_cast<const float4 *>(red + i);
            __threadfence();
            bar_consumers();
            const long long t0 = (long long)this_tile * ktiles;
            int ncontrib;
            if (tiles_mode > 0) {
                ncontrib = (ktiles + per_t - 1) / per_t;  // non-empty splits of this tile
                c_first = 0;
                c_last = ncontrib - 1;                    // split ordinals; CTA = tile + ordinal * T
            } else {
                c_first = (int)(((t0 + 1) * G - 1) / U);      // largest c' with c'*U/G <= t0
                c_last = (int)(((t0 + ktiles) * G - 1) / U);  // largest c' with c'*U/G <= t0 + ktiles - 1
                ncontrib = c_last - c_first + 1;
            }
            if (tid == 0) {
                const int old = atomicAdd(&tickets[this_tile], 1);
                const int last = (old == ncontrib - 1);
                if (last) tickets[this_tile] = 0;  // self-cleaning: every contributor has already arrived
                *s_flag = last;
            }
            bar_consumers();
            finish = (*s_flag != 0);
            if (finish) __threadfence();
        }
        if (finish) {
            const int act = g.act & 0xff;
            for (int idx = tid; idx < MT * 16 * (KS_BN / 2); idx += 128) {
                const int row = idx / (KS_BN / 2), col = (idx % (KS_BN / 2)) * 2;
                const int gn = n0 + col;
                if (row >= g.m || gn >= gN) continue;
                float2 v;
                if (whole) {
                    v = *reinterpret_cast<const float2 *>(&red[row * KS_BN + col]);
                } else {
                    v = make_float2(0.f, 0.f);
                    for (int cc = c_first; cc <= c_last; ++cc) {  // fixed order -> deterministic sum
                        size_t slot;
                        if (tiles_mode > 0) {
                            slot = (size_t)(this_tile + cc * tiles_mode) * 2;
                        } else {
                            const int jj = ((int)(((long long)cc * U / G) / ktiles) == this_tile) ? 0 : 1;
                            slot = (size_t)cc * 2 + jj;
                        }
                        const float2 p = __ldcg(reinterpret_cast<const float2 *>(
                            slots + slot * Cfg::RED_FLOATS + row * KS_BN + col));
                        v.x += p.x;
                        v.y += p.y;
                    }
                }
                if (bias) {
                    if (round_first) {
                        v.x = round_t<T>(v.x);
                        v.y = round_t<T>(v.y);
                    }
                    v.x += to_f(bias[row * g.bias_sm + gn * g.bias_sn]);
                    if (gn + 1 < gN) v.y += to_f(bias[row * g.bias_sm + (gn + 1) * g.bias_sn]);
                }
                v.x = gemm_act(act, v.x);
                v.y = gemm_act(act, v.y);
                T *dst = C + (int64_t)row * gN + gn;
                if (gn + 1 < gN && (gN & 1) == 0) {
                    if constexpr (std::is_same<T, __nv_bfloat16>::value)
                        *reinterpret_cast<__nv_bfloat162 *>(dst) = __floats2bfloat162_rn(v.x, v.y);
                    else
                        *reinterpret_cast<__half2 *>(dst) = __floats2half2_rn(v.x, v.y);
                } else {
                    dst[0] = from_f<T>(v.x);
                    if (gn + 1 < gN) dst[1] = from_f<T>(v.y);
                }
            }
        }
        bar_consumers();  // `red` is reused by the next segment
        zero_acc();
    }
}

// ---- per-(device, stream) scratch: 2 fp32 slots per CTA + one ticket per column tile, zero-initialised once ----
struct StreamKScratch {
    float *slots = nullptr;
    int *tickets = nullptr;
};
static StreamKScratch *get_scratch(cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, StreamKScratch> table;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto &sc = table[{dev, st}];
    if (!sc.slots) {
        const size_t slot_bytes = (size_t)KS_MAX_CTAS * 4 * 2 * StreamKCfg<4>::RED_FLOATS * 4;
        if (cudaMalloc(&sc.slots, slot_bytes) != cudaSuccess) return nullptr;
        if (cudaMalloc(&sc.tickets, KS_MAX_TILES * sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMemset(sc.tickets, 0, KS_MAX_TILES * sizeof(int)) != cudaSuccess) return nullptr;
    }
    return &sc;
}

static bool streamk_pure() {
    static int v = -1;
    if (v < 0) {
        const char *e = std::getenv("ITB_STREAMK_PURE");
        v = e && e[0] == '1';
    }
    return v == 1;
}
static int streamk_ctas_per_sm() {
    static int v = 0;
    if (!v) {
        const char *e = std::getenv("ITB_STREAMK_CTAS_PER_SM");
        v = e && e[0] == '1' ? 1 : 2;
    }
    return v;
}

template <typename T, int MT>
static int launch_streamk_t(const GemmArgs &g, int ngroups, const void *const *Ws, void *const *Cs, const int *Ns,
                            cudaStream_t st) {
    using Cfg = StreamKCfg<MT>;
    StreamKGroups grp{};
    CUtensorMap mapX;
    grp.ngroups = ngroups;
    int tiles_n = 0;
    for (int i = 0; i < ngroups; ++i) {
        if (!make_tma_2d_b16(&grp.mapW[i], Ws[i], (uint64_t)g.k, (uint64_t)Ns[i], (uint64_t)Ns[i], KS_BK, KS_BN, 128))
            ITB_FAIL("matmul(streamk): cuTensorMapEncodeTiled(W) failed");
        grp.C[i] = Cs[i];
        grp.n[i] = Ns[i];
        grp.tile_start[i] = tiles_n;
        tiles_n += (Ns[i] + KS_BN - 1) / KS_BN;
    }
    for (int i = ngroups; i <= KS_MAX_GROUPS; ++i) grp.tile_start[i] = tiles_n;
    if (tiles_n > KS_MAX_TILES) return -1;
    if (!make_tma_2d_b16(&mapX, g.A, (uint64_t)g.m, (uint64_t)g.k, (uint64_t)g.k, MT * 16, KS_BK, 128))
        ITB_FAIL("matmul(streamk): cuTensorMapEncodeTiled(X) failed");
    const int ktiles = (g.k + KS_BK - 1) / KS_BK;
    const long long U = (long long)tiles_n * ktiles;
    long long G = (long long)kNumSMs * streamk_ctas_per_sm();
    if (G > U) G = U;
    int tiles_mode = 0;
    if (!streamk_pure()) {
        // (tile, split) grid: split-K so that the grid comes close to (but not over) the resident-CTA budget
        int splitk = (int)(G / tiles_n);
        splitk = std::max(1, std::min(splitk, 16));
        splitk = std::min(splitk, std::max(1, ktiles / 4));
        int per = (ktiles + splitk - 1) / splitk;
        splitk = (ktiles + per - 1) / per;
        tiles_mode = tiles_n;
        G = (long long)tiles_n * splitk;
        if (G > KS_MAX_CTAS * 4) return -1;
    }
    StreamKScratch *sc = get_scratch(st);
    ITB_CHECK(sc != nullptr, "matmul(streamk): scratch allocation failed");

    static bool attr_done = false;
    auto kern = gemm_streamk_kernel<T, MT>;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        ITB_CHECK(e == cudaSuccess, "matmul(streamk): smem attribute: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    cudaError_t e = launch_k(kern, dim3((unsigned)G), dim3(KS_THREADS), Cfg::SMEM, st, grp, mapX, g, ktiles, U, tiles_mode, sc->slots,
                             sc->tickets);
    ITB_CHECK(e == cudaSuccess, "matmul(streamk): launch failed: %s", cudaGetErrorString(e));
    itb::count_launch();
    return 0;
}

static bool streamk_ok(int dtype, const GemmArgs &g) {
    if (dtype != ITB_BF16 && dtype != ITB_F16) return false;
    if (g.batch != 1 || g.trans_a || g.trans_b || g.m > 64 || g.m < 1) return false;
    if (g.n % 8 != 0 || g.k % 8 != 0 || g.n < 64 || g.k < 64) return false;
    if (!aligned16(g.A) || !aligned16(g.B) || ((uintptr_t)g.C & 3)) return false;
    return true;
}

int launch_gemm_streamk_grouped(int dtype, const GemmArgs &g0, int ngroups, const void *const *Ws, void *const *Cs,
                                const int *Ns, cudaStream_t st) {
    if (ngroups < 1 || ngroups > KS_MAX_GROUPS) return -1;
    for (int i = 0; i < ngroups; ++i) {
        GemmArgs g = g0;
        g.B = Ws[i];
        g.C = Cs[i];
        g.n = Ns[i];
        if (!streamk_ok(dtype, g)) return -1;
    }
    const int mt = (g0.m + 15) / 16;
#define KS_GO(TT)                                                                              \
    do {                                                                                       \
        if (mt == 1) return launch_streamk_t<TT, 1>(g0, ngroups, Ws, Cs, Ns, st);              \
        if (mt == 2) return launch_streamk_t<TT, 2>(g0, ngroups, Ws, Cs, Ns, st);              \
        return launch_streamk_t<TT, 4>(g0, ngroups, Ws, Cs, Ns, st);                           \
    } while (0)
    if (dtype == ITB_BF16) KS_GO(__nv_bfloat16);
    KS_GO(__half);
#undef KS_GO
}

int launch_gemm_streamk(int dtype, const GemmArgs &g, cudaStream_t st) {
    const void *Ws[1] = {g.B};
    void *Cs[1] = {g.C};
    int Ns[1] = {g.n};
    return launch_gemm_streamk_grouped(dtype, g, 1, Ws, Cs, Ns, st);
}

}  // namespace itb
