// decode_stack.cuh -- the device-side "program" of the persistent decode kernel (decode_stack.cu).
//
// A program is a list of PHASES executed by ONE persistent grid (one CTA per SM); consecutive phases are separated by a
// grid-wide barrier in L2 instead of a kernel boundary.  Two phase kinds:
//   DS_GEMM  out[g][16, N_g] = xform(X)[16, K] . W_g[K, N_g]   (up to 3 weight matrices sharing X: q/k/v, gate/up)
//            tcgen05 swap-AB: 128 weight columns per UMMA (A, MN-major, two [64k x 64n] TMA boxes), X is the 16-row
//            K-major B operand, fp32 accumulator [128 lanes x 16 columns] in TMEM.
//   DS_ATTN  AttentionKVCache (+ RoPE of q and k) over the (head, 64-row chunk) units of the batch.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace itb {

constexpr int DS_MAX_GROUPS = 3;
constexpr int DS_ROWS = 16;  // activation rows (tokens) per launch: the UMMA N dimension

enum { DS_GEMM = 0, DS_ATTN = 1 };
enum { DS_XF_NONE = 0, DS_XF_RMSNORM = 1, DS_XF_SILU_MUL = 2 };
enum { DS_EPI_STORE = 0, DS_EPI_RESIDUAL = 1 };

struct alignas(64) DsPhase {
    CUtensorMap mapW[DS_MAX_GROUPS];  // W_g [K, N_g] row-major, box [64 k x 64 n], 128B swizzle
    CUtensorMap mapX;                 // activation [rows, K], box [16 x 64 k], 128B swizzle (xform 2: the gate half)
    CUtensorMap mapX2;                // xform 2: the up half
    int kind;
    // ---- GEMM
    int ngroups, n_per_group, tiles_per_group, ntiles, kchunks, K;
    int xform, epi;
    void *out[DS_MAX_GROUPS];   // [rows, n_per_group]
    const void *residual;       // epi 1: [rows, n_per_group], added after the product has been rounded to the storage type
    float *ss_out;              // epi 1: per-tile sums of squares of the NEW residual rows [ntiles][16] (feeds the next RMSNorm)
    const float *ss_in;         // xform 1: the producer phase's ss_out (nullptr: every CTA reduces x_raw itself)
    int ss_in_tiles;
    const void *x_raw;          // xform 1: the un-normalised rows [rows, K]
    const void *norm_w;         // xform 1: RMSNorm weight [K]
    float *partial;             // split-K partial tiles [ntiles][slots_per_tile][16][128] fp32
    int slots_per_tile;         // == nsplit
    int nsplit;                 // split-K factor S: units = ntiles x S, unit -> (split = unit / ntiles, tile = unit % ntiles)
    int *tickets;               // [ntiles], self-cleaning
    // ---- attention
    void *kcache, *vcache;      // [B, H, S_max, 128]
    const void *q, *k, *v;      // [B, H*128] (q, k before RoPE)
    void *attn_out;             // [B, H*128]
    float *attn_partial;        // [B*H][slots_per_head][130]
    int *attn_tickets;          // [B*H]
    int H, S_max, slots_per_head;
};

struct DsProgram {
    const DsPhase *phases;  // device memory
    int nphases;
    int rows;               // valid activation rows (<= 16)
    const void *position_id;
    int pos_flags;          // dtype | ITB_POS_*
    const void *rope_pos;   // per-row RoPE positions (nullptr: no RoPE)
    int rope_pos_dtype;
    unsigned int *grid_bar; // {count, generation}, zero-initialised once per device
    unsigned long long *trace;  // optional (tools/ds_trace.py): [cta][phase][2] globaltimer at phase start / this CTA's arrival
};

// host side (decode_stack.cu)
struct DsGemmDesc {
    int ngroups;
    const void *W[DS_MAX_GROUPS];
    void *out[DS_MAX_GROUPS];
    int n_per_group, K;
    int xform, epi;
    const void *X, *X2;       // activation operand(s) in global memory ([rows, K])
    const void *residual;
    const void *norm_w;
    int ss_from_prev_residual;  // xform 1: 1 = the previous epi-1 phase's ss_out feeds rinv; 0 = reduce X directly
};

}  // namespace itb
