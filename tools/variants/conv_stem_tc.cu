// conv_stem_tc.cu -- the stem convolution (NCHW input with <= 4 channels -> NHWC output) on the tcgen05 tensor cores.
//
// Same idea as conv_stem.cu (mma.sync, 114 us on ResNet-50's 7x7 s2 stem at 40 % tensor-pipe activity: HMMA-issue bound), but the
// patch is turned into a REAL K-major UMMA operand in shared memory so one thread can issue the whole tile's arithmetic:
//   * a persistent CTA (one per SM) walks 8 x 16 blocks of output pixels (M = 128 = the TMEM lanes);
//   * the block's input patch sits in shared memory channel-interleaved ({c0..c3} = 8 B per pixel); k = (r, s -> S4, c -> 4), so a
//     16-byte chunk of an A row = two horizontally adjacent patch pixels: 256 threads copy 28 chunks per pixel row (two 8-byte
//     loads, one swizzled 16-byte store each) into the 128B-swizzled K-major tile [K/64 blocks][128 rows][128 B];
//   * filters [64 f][K] K-major, built once per CTA; D[128 px, 64 f] += A . B^T: K/16 tcgen05.mma (M 128, N 64) by one thread into
//     one of two TMEM accumulators; the next patch is fetched into registers and the previous tile's epilogue runs meanwhile;
//   * epilogue: tcgen05.ld, BatchNorm (one FMA) + ReLU, per-warp staging slab [2 rows x 16 px x 32 f], TMA store through a 4-D map
//     (f, ow, oh, n) that clips blocks cut by the image border.
// Replaces cudnnConvolutionForward (+ BatchNorm + Relu) for the stem (reference src/kernels/cuda/conv.cc:143-168).
#include <algorithm>
#include <cstdlib>

#include <cudaTypedefs.h>

#include "conv_shapes.h"
#include "gemm.cuh"

namespace itb {

constexpr int SC_TH = 8, SC_TW = 16;  // output block: 8 rows x 16 columns = 128 pixels
constexpr int SC_BUILD_WARPS = 8, SC_BUILD_THREADS = 256, SC_THREADS = 288;  // + 1 MMA warp
constexpr int SC_KB_MAX = 4;          // K <= 256
constexpr int SC_A_BLOCK = 128 * 128; // one 64-wide k-block of the A tile
constexpr int SC_B_BLOCK = 64 * 128;
constexpr int SC_SLAB = 32 * 64;      // [32 pixels x 32 filters] staging slab

struct StemTcParams {
    const void *x, *w;
    const float *bn_mean, *bn_var, *bn_scale, *bn_bias;
    float bn_eps;
    int relu;
    int N, C, H, W, F, R, S, OH, OW, ph, pw, sh, sw;
    int tiles_h, tiles_w, tiles;
    int PH, PW, S4, K, KB;  // patch size, taps per filter row (multiple of 4), K = R * S4 * 4, k-blocks of 64
    int chunks;             // 16-byte chunks per A row that carry data: K / 8
    int patch_elems;        // PH * PW
    int off_b, off_patch, off_stg, off_bn, off_bar;
    uint32_t idesc;
};

__device__ __forceinline__ void sc_tma_store_4d(const CUtensorMap *m, const void *smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void sc_tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <typename T> __device__ __forceinline__ uint32_t sc_pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t sc_pack2<__half>(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t sc_pack2<__nv_bfloat16>(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

constexpr int SC_PF = 4;  // patch pixels a thread prefetches: PH * PW <= 21 * 38 for every supported shape (R <= 7, S <= 8, strides <= 2)

template <typename T>
__global__ void __launch_bounds__(SC_THREADS, 1) conv_stem_tc_kernel(const __grid_constant__ CUtensorMap mapY, StemTcParams p) {
    extern __shared__ uint8_t sc_raw[];
    uint8_t *smem = sc_raw + ((1024u - (smem_u32(sc_raw) & 1023u)) & 1023u);
    uint8_t *a_sm = smem;                     // 2 x KB x [128 rows x 128 B]
    uint8_t *b_sm = smem + p.off_b;           // KB x [64 rows x 128 B]
    uint2 *patch = reinterpret_cast<uint2 *>(smem + p.off_patch);
    uint8_t *stg = smem + p.off_stg;          // 8 warps x 2 slabs
    float2 *bn_sm = reinterpret_cast<float2 *>(smem + p.off_bn);
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + p.off_bar);  // [2] builders -> MMA
    uint64_t *a_free = a_full + 2;                                       // [2] MMA (commit) -> builders
    uint64_t *acc_full = a_free + 2;                                     // [2] MMA (commit) -> epilogue
    uint64_t *acc_free = acc_full + 2;                                   // [2] epilogue -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_free[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], SC_BUILD_WARPS);
        }
        fence_mbar_init();
        tma_prefetch_desc(&mapY);
    }
    if (warp == SC_BUILD_WARPS) tmem_alloc(tmem_slot, 128);
    const int a_bytes = p.KB * SC_A_BLOCK;
    const T *X = (const T *)p.x;

    if (warp < SC_BUILD_WARPS) {
        // ---- constants: filters -> K-major swizzled B tile [kb][64 f][128 B]; zero the never-written tail of both A buffers
        const T *Wg = (const T *)p.w;
        const int RS = p.R * p.S, total_chunks = p.KB * 8;
        for (int i = threadIdx.x; i < 64 * total_chunks; i += SC_BUILD_THREADS) {
            const int f = i / total_chunks, j = i - f * total_chunks;  // 16-byte chunk j of row f: taps 2j, 2j + 1
            T v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int tap = 2 * j + (e >> 2), c = e & 3;
                const int r = tap / p.S4, s = tap - r * p.S4;
                v[e] = (f < p.F && r < p.R && s < p.S && c < p.C) ? Wg[(f * p.C + c) * RS + r * p.S + s] : from_f<T>(0.f);
            }
            *reinterpret_cast<uint4 *>(b_sm + (j >> 3) * SC_B_BLOCK + f * 128 + (((j & 7) ^ (f & 7)) << 4)) = *reinterpret_cast<const uint4 *>(v);
        }
        for (int i = threadIdx.x; i < 2 * 128 * (total_chunks - p.chunks); i += SC_BUILD_THREADS) {
            const int per = total_chunks - p.chunks;
            const int buf = i / (128 * per), rem = i - buf * 128 * per, m = rem / per, j = p.chunks + rem % per;
            *reinterpret_cast<uint4 *>(a_sm + buf * a_bytes + (j >> 3) * SC_A_BLOCK + m * 128 + (((j & 7) ^ (m & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int f = threadIdx.x; f < 64; f += SC_BUILD_THREADS) {
            float2 ab = make_float2(f < p.F ? 1.f : 0.f, 0.f);
            if (f < p.F && p.bn_scale) {
                ab.x = p.bn_scale[f] * bn_rs(p.bn_var[f], p.bn_eps);
                ab.y = __fmaf_rn(-p.bn_mean[f], ab.x, p.bn_bias[f]);
            }
            bn_sm[f] = ab;
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == SC_BUILD_WARPS) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint64_t a_desc0 = umma_desc_sw128(smem_u32(a_sm), 0, 1024), b_desc0 = umma_desc_sw128(smem_u32(b_sm), 0, 1024);
            const int ksteps = p.K / 16;
            int it = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1, use = it >> 1;
                mbar_wait(&a_full[buf], use & 1);
                if (it >= 2) mbar_wait(&acc_free[buf], (use - 1) & 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 64);
                for (int j = 0; j < ksteps; ++j) {
                    const uint32_t a_off = (uint32_t)((buf * a_bytes + (j >> 2) * SC_A_BLOCK + (j & 3) * 32) >> 4);
                    const uint32_t b_off = (uint32_t)(((j >> 2) * SC_B_BLOCK + (j & 3) * 32) >> 4);
                    tc_mma_f16(d_tmem, a_desc0 + a_off, b_desc0 + b_off, p.idesc, j > 0 ? 1u : 0u);
                }
                tc_commit(&a_free[buf]);
                tc_commit(&acc_full[buf]);
            }
        }
        __syncwarp();
    } else {
        // ===== builders / epilogue (256 threads) =====
        const int quad = warp & 3, half = warp >> 2;
        uint8_t *my_stg = stg + warp * 2 * SC_SLAB;
        // this thread's share of the A build: pixel row m, chunks j = jb, jb + 2, ... (two threads per row)
        const int m = threadIdx.x & 127, jb = threadIdx.x >> 7;
        const int ohl = m >> 4, owl = m & 15;
        const uint32_t *patch_w = reinterpret_cast<const uint32_t *>(patch);

        auto tile_coords = [&](int tile, int &n, int &oh0, int &ow0) {
            const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h;
            n = tile / (p.tiles_w * p.tiles_h);
            oh0 = th * SC_TH;
            ow0 = tw * SC_TW;
        };
        // patch pixels of a tile -> registers ({c0..c3} packed in 8 bytes; zero outside the image / beyond C)
        uint2 pf[SC_PF];
        int pf_py[SC_PF], pf_px[SC_PF];  // this thread's patch pixels: the same for every tile
#pragma unroll
        for (int i = 0; i < SC_PF; ++i) {
            const int e = threadIdx.x + i * SC_BUILD_THREADS;
            pf_py[i] = e < p.patch_elems ? e / p.PW : -(1 << 20);
            pf_px[i] = e < p.patch_elems ? e - (e / p.PW) * p.PW : 0;
        }
        const int64_t plane = (int64_t)p.H * p.W;
        auto fetch = [&](int tile) {
            int n, oh0, ow0;
            tile_coords(tile, n, oh0, ow0);
            const int ih0 = oh0 * p.sh - p.ph, iw0 = ow0 * p.sw - p.pw;
            const T *img = X + (int64_t)n * p.C * plane;
#pragma unroll
            for (int i = 0; i < SC_PF; ++i) {
                const int ih = ih0 + pf_py[i], iw = iw0 + pf_px[i];
                const bool in = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
                const T *src = img + (in ? ih * p.W + iw : 0);
                const T z = from_f<T>(0.f);
                const T v0 = in ? src[0] : z, v1 = (in && p.C > 1) ? src[plane] : z, v2 = (in && p.C > 2) ? src[2 * plane] : z,
                        v3 = (in && p.C > 3) ? src[3 * plane] : z;
                pf[i].x = (uint32_t)(*reinterpret_cast<const unsigned short *>(&v0)) | ((uint32_t)(*reinterpret_cast<const unsigned short *>(&v1)) << 16);
                pf[i].y = (uint32_t)(*reinterpret_cast<const unsigned short *>(&v2)) | ((uint32_t)(*reinterpret_cast<const unsigned short *>(&v3)) << 16);
            }
        };
        auto commit_patch = [&]() {
#pragma unroll
            for (int i = 0; i < SC_PF; ++i) {
                const int e = threadIdx.x + i * SC_BUILD_THREADS;
                if (e < p.patch_elems) patch[e] = pf[i];
            }
        };

        // tile (block) -> BatchNorm + ReLU -> staging slab -> TMA store; the accumulator is released as soon as it is in registers
        auto epilogue = [&](int t_prev, int it_prev) {
            const int pb = it_prev & 1, pu = it_prev >> 1;
            int n, oh0, ow0;
            tile_coords(t_prev, n, oh0, ow0);
            mbar_wait(&acc_full[pb], pu & 1);
            tc_fence_after();
            uint32_t v[32];
            sc_tmem_ld_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(pb * 64 + half * 32), v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[pb]);
            uint8_t *slab = my_stg + (it_prev & 1) * SC_SLAB;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store that used this slab two tiles ago
            __syncwarp();
#pragma unroll
            for (int j8 = 0; j8 < 32; j8 += 8) {
                uint32_t ow[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 ab0 = bn_sm[half * 32 + j8 + 2 * e], ab1 = bn_sm[half * 32 + j8 + 2 * e + 1];
                    float x0 = __fmaf_rn(__uint_as_float(v[j8 + 2 * e]), ab0.x, ab0.y);
                    float x1 = __fmaf_rn(__uint_as_float(v[j8 + 2 * e + 1]), ab1.x, ab1.y);
                    if (p.relu) {
                        x0 = fmaxf(x0, 0.f);
                        x1 = fmaxf(x1, 0.f);
                    }
                    ow[e] = sc_pack2<T>(x0, x1);
                }
                // slab = [2 image rows][16 columns][32 filters]: pixel row (lane) x 64 B, 64B swizzle
                *reinterpret_cast<uint4 *>(slab + lane * 64 + (((j8 >> 3) ^ ((lane >> 1) & 3)) << 4)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                sc_tma_store_4d(&mapY, slab, half * 32, ow0, oh0 + 2 * quad, n);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
        };
        int it = 0;
        if (blockIdx.x < p.tiles) {
            fetch(blockIdx.x);
            commit_patch();
        }
        named_bar_sync(1, SC_BUILD_THREADS);
        int prev_tile = -1;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1, use = it >> 1;
            const int next = tile + gridDim.x;
            if (next < p.tiles) fetch(next);  // global loads in flight under the build and the epilogue
            // ---- A tile of this block from the patch
            if (it >= 2) mbar_wait(&a_free[buf], (use - 1) & 1);  // the MMAs that read this buffer are done
            {
                uint8_t *arow = a_sm + buf * a_bytes + m * 128;
                const int base_px = (ohl * p.sh) * p.PW + owl * p.sw;  // patch pixel of tap (0, 0)
                // chunk j = taps 2j, 2j + 1 of filter row r: this thread takes j = jb, jb + 2, ... -> the tap advances by 4
                int s = 2 * jb, roff = base_px;  // (2 jb < 4 <= S4)
                for (int j = jb; j < p.chunks; j += 2) {
                    const uint32_t *src = patch_w + (roff + s) * 2;
                    const uint2 lo = *reinterpret_cast<const uint2 *>(src), hi = *reinterpret_cast<const uint2 *>(src + 2);
                    *reinterpret_cast<uint4 *>(arow + (j >> 3) * SC_A_BLOCK + (((j & 7) ^ (m & 7)) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
                    s += 4;
                    if (s >= p.S4) {
                        s -= p.S4;
                        roff += p.PW;
                    }
                }
            }
            fence_proxy_async();
            named_bar_sync(1, SC_BUILD_THREADS);  // A complete; every reader of the patch is done
            if (threadIdx.x == 0) mbar_arrive(&a_full[buf]);
            // ---- epilogue of the PREVIOUS tile (its MMAs ran while this tile was built), then the next patch (whose global loads
            // have had the build and the epilogue to land) goes to shared memory
            if (prev_tile >= 0) epilogue(prev_tile, it - 1);
            if (next < p.tiles) commit_patch();
            prev_tile = tile;
            named_bar_sync(1, SC_BUILD_THREADS);  // the next patch is in shared memory
        }
        if (prev_tile >= 0) epilogue(prev_tile, it - 1);  // the last tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == SC_BUILD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

// 4-D output map (f, ow, oh, n) over NHWC y, box [32 f][16 ow][2 oh][1], 64B swizzle
static bool make_stem_out_map(CUtensorMap *map, const void *y, int N, int OH, int OW, int F) {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return false;
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    cuuint64_t gdim[4] = {(cuuint64_t)F, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)F * 2, (cuuint64_t)OW * F * 2, (cuuint64_t)OH * OW * F * 2};
    cuuint32_t box[4] = {32, 16, 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(y), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// -1 = the shape is not this kernel's (conv_stem.cu's mma.sync kernel takes it)
int launch_conv_stem_tc(int dtype, const void *x, const void *w, void *y, int N, int C, int H, int W, int F, int R, int S, int OH,
                        int OW, int ph, int pw, int sh, int sw, const float *bn_mean, const float *bn_var, const float *bn_scale,
                        const float *bn_bias, float bn_eps, int relu, cudaStream_t st) {
    {
        const char *e = std::getenv("ITB_STEM_IMPL");  // "mma" pins conv_stem.cu's kernel (tests cover both)
        if (e && e[0] == 'm') return -1;
    }
    StemTcParams p{};
    p.S4 = ((S + 3) / 4) * 4;
    p.K = R * p.S4 * 4;
    p.KB = (p.K + 63) / 64;
    p.PH = (SC_TH - 1) * sh + R;
    p.PW = (SC_TW - 1) * sw + p.S4;
    p.patch_elems = p.PH * p.PW;
    if (p.KB > SC_KB_MAX || p.K % 16 != 0 || F > 64 || F % 8 != 0 || p.patch_elems > SC_PF * SC_BUILD_THREADS || !aligned16(y)) return -1;
    p.x = x;
    p.w = w;
    p.bn_mean = bn_mean;
    p.bn_var = bn_var;
    p.bn_scale = bn_scale;
    p.bn_bias = bn_bias;
    p.bn_eps = bn_eps;
    p.relu = relu;
    p.N = N;
    p.C = C;
    p.H = H;
    p.W = W;
    p.F = F;
    p.R = R;
    p.S = S;
    p.OH = OH;
    p.OW = OW;
    p.ph = ph;
    p.pw = pw;
    p.sh = sh;
    p.sw = sw;
    p.tiles_h = (OH + SC_TH - 1) / SC_TH;
    p.tiles_w = (OW + SC_TW - 1) / SC_TW;
    const int64_t tiles = (int64_t)N * p.tiles_h * p.tiles_w;
    if (tiles >= (1ll << 31)) return -1;
    p.tiles = (int)tiles;
    p.chunks = p.K / 8;
    p.off_b = 2 * p.KB * SC_A_BLOCK;
    p.off_patch = p.off_b + p.KB * SC_B_BLOCK;
    p.off_stg = (p.off_patch + p.patch_elems * 8 + 1023) & ~1023;
    p.off_bn = p.off_stg + SC_BUILD_WARPS * 2 * SC_SLAB;
    p.off_bar = p.off_bn + 64 * 8;
    const int smem = p.off_bar + 8 * 8 + 16 + 1024;
    if (smem > 227 * 1024) return -1;
    p.idesc = umma_idesc_f16(dtype == ITB_BF16 ? 1 : 0, 0, 0, 128, 64);
    alignas(64) CUtensorMap mapY;
    if (!make_stem_out_map(&mapY, y, N, OH, OW, F)) ITB_FAIL("conv(stem, tcgen05): cuTensorMapEncodeTiled(y) failed");
    auto go = [&](auto kern, int *attr_smem) -> int {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (smem > attr_smem[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            ITB_CHECK(e == cudaSuccess, "conv(stem, tcgen05): smem attribute: %s", cudaGetErrorString(e));
            attr_smem[dev] = smem;
        }
        cudaError_t e = launch_k(kern, dim3((unsigned)std::min<int64_t>(tiles, kNumSMs)), dim3(SC_THREADS), (size_t)smem, st, mapY, p);
        ITB_CHECK(e == cudaSuccess, "conv(stem, tcgen05): launch failed: %s", cudaGetErrorString(e));
        count_launch();
        return 0;
    };
    static int attr_h[64] = {0}, attr_b[64] = {0};
    if (dtype == ITB_F16) return go(conv_stem_tc_kernel<__half>, attr_h);
    return go(conv_stem_tc_kernel<__nv_bfloat16>, attr_b);
}

}  // namespace itb
