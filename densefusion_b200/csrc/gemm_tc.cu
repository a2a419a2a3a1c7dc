// Tensor-core GEMM / implicit-GEMM convolution of the pose path: tcgen05.mma with fp32 accumulators in TMEM, operands
// by TMA, persistent warp-specialised CTAs (or CTA pairs, cta_group::2), same contract and fused epilogues as gemm_simt.cu.
//
//   C[m,n] = act( sum_k A[m,k] W[n,k] + bias )         M tile 128 rows per CTA (one TMEM lane per row)
//
// fp32 parity on 11-bit tensor-core operands: x = hi + lo, D += A_hi W_hi + A_lo W_hi + A_hi W_lo (the dropped A_lo W_lo term
// is O(2^-22) relative).  "3xtf32": hi = x & 0xffffe000, all three products kind::tf32.  "hybrid": the two correction terms
// with bf16 operands (kind::f16).  "hybrid16": the main term on fp16 operands as well (see gemm_tc_q_kernel).  "tf32": the
// first product only (stated looser bound).
//
// The A operand is split in registers by stager warps and written to TMEM (tcgen05.st; the MMA reads it in .ts form), the
// weight operand is split once at pack time (df_split_tf32 / df_pack_*).  (Round 2 also tried splitting the fp32 weight tile on chip,
// "hybrid16w": bit-identical but slower -- the conversion instructions set the pace -- and removed again)
// The first two kernel generations of round 1 (one tile per CTA; persistent with A read straight from global memory) are
// gone: they were superseded by gemm_tc_q_kernel on every shape (DESIGN.md section 4 keeps their measurements).
#include "df_common.cuh"
#include "../../include/densefusion_b200.h"
#include <cuda.h>
#include <stdlib.h>
#include <stdio.h>
#include <vector>

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                       // floats per k-block = 128 B = one SW128 atom row
constexpr int UMMA_K = 8;                    // tf32

struct TcParams {
    const float* A; int lda;
    const float* bias; int bias_crop_stride;
    float* C; int ldc;
    int M, N, K, relu, precise;
    int rows_per_crop;
    long long a_gs, bias_gs, c_gs;
    float* pool_partial; int tiles_per_crop;
    // ---- implicit-convolution form (generation-2 kernels only; conv_taps == 0: plain GEMM) ----
    // A is an NHWC image (cB, cH, cW, K channels); an M tile is a (TB x TH x TW) patch of output pixels and the k loop
    // walks taps x 32-channel blocks, each A tile being the patch shifted by the tap -- a 4-D TMA box whose
    // out-of-image part is zero-filled, i.e. the convolution's zero padding costs nothing.
    int conv_taps, conv_dil;                    // 1 or 9 (3x3, stride 1, padding == dilation)
    int cW, cH, cB, TW, TH, TB, tiles_x, tiles_y, tiles_b;
    const float* residual; int ldr;             // added before the activation (BasicBlock skip), same pixel order as C
    const float* prelu;                         // relu == 2: y = x > 0 ? x : prelu[0] * x
    // ---- chunked accumulation (generation-2 kernels): the k loop of one tile is cut into k_chunks runs of <= kbc k-blocks;
    // every run starts a fresh TMEM accumulator and the epilogue adds the runs in fp32 (round-to-nearest) through C.
    // Why: the tensor core truncates when it aligns the accumulator with new products, a bias that grows with the length
    // of the chain (measured 3xTF32 error 3e-6 at K = 384 but 6.5e-5 at K = 9216); short chains keep long-K convolutions
    // at fp32-parity.  k_chunks == 1: plain single-run accumulation.
    int k_chunks, kbc;
    // Rounding-bias compensation: the tensor core TRUNCATES TOWARD ZERO when it adds an instruction's products into the fp32
    // accumulator -- measured -0.30 .. -0.45 ulp per MMA instruction for same-sign sums, almost free of scatter (rms == |mean|,
    // scripts/trunc_probe.py, profiles/r2_c7_trunc.jsonl) -- a bias that grows with the chain and, unlike rounding noise, adds
    // up coherently from layer to layer.  The epilogue multiplies every accumulation run by 1 + bias_comp * (MMA instructions
    // of the run); bias_comp = the measured mean truncation per instruction for the layer's operand statistics (0: off).
    float bias_comp;
    // "hybrid16s" (precise == 4): both operands as TWO fp16 planes, x = fp16(x s) + fp16(x s - fp16(x s)) with a power-of-two scale s
    // that lifts the remainder plane out of fp16's subnormals: the weight planes are packed once with the tensor's own scale
    // (df_pack_f16s; w_inv_scale = device pointer to 1 / s_w), the activation is scaled in the stagers' registers (a_scale, chosen by
    // the caller).  The epilogue multiplies every accumulation run by *w_inv_scale / a_scale.
    // a_scale == 0 (default): every CTA derives the scale from the SAME 4096-element sample of the operand (512 rows x 8 columns,
    // samp_rows x samp_cols addressable; see "activation scale" in the kernel), so it follows the data without a host round trip.
    const float* w_inv_scale; float a_scale;
    long long samp_rows; int samp_cols;
    // conv1 form (hybrid16s, A planes in TMEM): A is NOT a matrix in memory -- the stagers gather the 7x7 / stride 2 / pad 3 patches of the
    // NCHW image p.A (c1_B x 3 x c1_H x c1_W) straight into their registers, row m = output pixel (b, yo, xo), column k = c*49 + ky*7 +
    // kx (K = 160: 147 + zero padding), so the 0.58 GB patch matrix of df_enc_im2col_conv1 and its trip through HBM disappear.
    // c1_H == 0: off.
    int c1_H, c1_W, c1_Ho, c1_Wo;
    // DF_TC_DBG bit 256: cluster 0 / leader CTA records clock64() at the hand-over points of its first TRACE_KB k-blocks
    // (producer: slot free, TMA issued; stager groups: bytes landed, TMEM slot free, A handed over; issuer: operands ready, MMAs
    // committed) into `trace` [TRACE_EV][TRACE_KB]; df_tc_trace_read copies it out.  Timeline of the pipeline, no effect on results.
    unsigned long long* trace;
    int dbg;                                    // DF_TC_DBG knock-out bits (timing experiments only; results are wrong): 1 no global stores,
                                                // 2 no transpose, 4 no TMEM load, 8 no MMAs, 16 no store instruction (reads / math of the store path kept), 32 stagers skip the fp16 split, 64 stagers skip the TMEM store
    int run_steps;                              // host side only: MMA instructions per accumulation run, 0 = default
    // ---- weight-gradient form (df_conv_wgrad_tc): output column n = tap * wk_rows + ci multiplies row ci of the W operand
    // read wk_shift[tap] elements further along k (a 3x3 tap is an offset in the zero-padded, flattened pixel axis) ----
    int wk_rows;                                // 0: off
    int wk_shift[9], wk_row0[9];                // per tap: k offset (a multiple of 4 elements: TMA boxes start 16-byte aligned) and first row
    int tstore;                                 // plain single-run GEMM form: the epilogue hands its 32 x 32 chunks to TMA stores (tm_c) instead
                                                // of transposing them through shared memory and storing them itself
    int m_fastest;                              // tile order of the persistent loop (q_decode).  Default 0 = n fastest: the clusters
                                                // running at one time share activation rows; measured 12% faster at M = 64000 /
                                                // K = 1536 than sharing the weight tile (env DF_TC_TILE_ORDER=1)
};

}  // namespace
#include "tc_ptx.cuh"
namespace {
using namespace df_tc;

// ------------------------------------------------------------------------------------------------
// Generation 2 of the persistent kernel ("q"): the A operand also arrives by TMA.
//
// Why: in the kernel above every stager lane reads its own accumulator row straight from global memory, i.e. one
// LDG.128 of a warp touches 32 different 128-byte lines.  The L1/tex pipe retires about one such line ("wavefront")
// per cycle, so a 128x32 fp32 k-block costs >= 1024 cycles of LSU time -- more than the 768 cycles its twelve
// 128x128x8 TF32 MMAs need.  That, not the tensor pipe, set the pace (profiles/r1_call6_tc_probe.txt: one TF32 pass
// was no faster than three).  Here a stage is {A 128x32 fp32 | W_hi | W_lo}, all 128B-swizzled TMA tiles; the stagers
// read their row back from shared memory (conflict-free LDS.128 thanks to the swizzle), split it and tcgen05.st the
// hi / lo halves into TMEM as before.
//
// CTAS == 2 additionally pairs two SMs on one 256 x (2*bn_cta) tile with tcgen05.mma.cta_group::2: every CTA stages
// its own 128 rows of A (TMEM) and HALF of the weight tile (shared memory), the tensor cores of both SMs read both
// halves, so the L2 -> SM and shared-memory bytes per flop are halved -- the tile moves from the L2 / smem roof
// (48 KB per 768 MMA cycles) under the tensor roof (48 KB per 1536).
//   warp 0       TMA producer (own A rows, own half of W_hi / W_lo) -> local full[s]
//   warp 1       MMA issuer + TMEM owner (leader CTA only issues; commits are multicast to both CTAs)
//   warps 2-9    A stagers, two groups alternating k-blocks; arrive on the LEADER's a_full[s] (remote arrive from the peer)
//   warps 10-17  epilogue: two warps per TMEM lane quarter, alternate 32-column chunks
// TMEM (512 columns): CTAS == 1: two 128-column accumulators in [0,256), four A stages of 64 columns (hi | lo) in
// [256,512).  CTAS == 2: accumulators at 0 and 192 (two when the tile is <= 192 wide, so the stores of tile i overlap
// the MMAs of tile i+1 -- 123 MB of tower-1 output cannot be left to bursts between tiles; one 256-wide accumulator
// only for the store-free pooled epilogue), two A stages in [384,512).
// ------------------------------------------------------------------------------------------------
// Per-instruction truncation compensation per arithmetic mode (see TcParams::bias_comp), calibrated on B200 with
// scripts/trunc_probe.py: mean signed error of mixed-sign / post-ReLU dot products per MMA instruction of the chain (hybrid16
// -1.61e-6 over 96 instructions, hybrid -2.31e-6 over 128, 3xtf32 -3.72e-6 over 192).  Effect on the bench configuration
// (profiles/r2_c8_bias_comp.txt): worst pose error of 32 crops against the oracle 1.03e-4 -> 3.5e-5, rms GEMM error halved.
constexpr float BIAS_COMP_H16 = 1.6e-8f, BIAS_COMP_HYBRID = 1.8e-8f, BIAS_COMP_3XTF32 = 1.9e-8f, BIAS_COMP_H16S = 1.6e-8f;
constexpr int Q_THREADS = 19 * 32;                           // warp 18: the second TMA producer
constexpr int Q_MAX_STAGES = 8;
constexpr int Q_TILE = 128 * BK * 4;                         // 16 KB: one 128-row fp32 tile of 32 k
constexpr int Q_SMEM_STAGES = 12 * Q_TILE;                   // 192 KB of operand stages: 4 x {A 16 KB | W_hi 16 | W_lo 16} for 128
                                                             // weight rows per CTA ... 8 x {16 | 4 | 4} for 32 (deeper pipelines
                                                             // for the narrow tiles, which are TMA-latency bound)
constexpr int Q_SMEM_EPI = 8 * 32 * 32 * 4;                  // 32 KB: one swizzled 32x32 transpose tile per epilogue warp
constexpr int Q_SMEM_TOTAL = 1024 + Q_SMEM_STAGES + Q_SMEM_EPI + 512;
static_assert(Q_SMEM_TOTAL <= 227 * 1024, "shared memory overflow");

constexpr int TRACE_KB = 96;
constexpr int TRACE_EV = 17;                                  // 0-8 per k-block (see TcParams::trace); 9-13 per accumulation run: epilogue warp 10 waits for / has / has drained the
                                                             // accumulator, issuer waits for / has a free accumulator; 14-16 per chunk of warp 10: accumulator in registers / transposed / stored
#define DF_TRACE(ev, it_) do { if (DBG && p.trace && cid == 0 && rank == 0 && (it_) < (uint32_t)TRACE_KB && lane == 0) p.trace[(ev) * TRACE_KB + (it_)] = clock64(); } while (0)

// Balanced schedule of the long-K convolutions (FORM 4).  The default tile walk is round robin over whole tiles: 180 tiles of a 15 x 15
// layer4 convolution on 74 CTA pairs are 2.43 tiles per pair, i.e. three rounds with the last one a third full (every pair waits for
// the 32 that got a third tile; 19% of the launch, 10% at 200 tiles).  Here every cluster gets a CONTIGUOUS range of (tile,
// accumulation run) units of equal k-block weight, cut at run boundaries -- the runs of a tile already meet in C (first run stored,
// the others added) -- so a tile may be shared by two neighbouring clusters: cluster c + 1 STARTS with the last runs of the tile
// (stores / adds them like intermediate runs, then raises a flag per epilogue warp), cluster c ENDS with its first runs (waits for the
// flag before it touches C, adds its runs, and its last one takes the full path: read back, bias, skip connection, activation).
// The order of the additions into an element is fixed by the schedule, so the result is deterministic (it differs from the
// round-robin schedule's in the last bit: (r2 + r3) + r0 + r1 instead of r0 + r1 + r2 + r3).  The waiting cluster never blocks
// progress: the runs it waits for are the FIRST thing its neighbour does, before any wait of its own.
constexpr int Q_SCHED_MAX = 80;
struct QSched {
    short t0[Q_SCHED_MAX], r0[Q_SCHED_MAX];      // first tile of cluster c and the run it starts with (> 0: the tile's first runs belong to c - 1)
    short t1[Q_SCHED_MAX], r1[Q_SCHED_MAX];      // last tile and the run it stops before (0x7fff: the whole tile)
    unsigned int* flags;                         // [clusters][2 CTAs][8 epilogue warps], zero between launches
};

struct QTile {
    int g, n0, row0, rows_valid, crop, pool_tile;
    int x0, y0, b0;                              // convolution form: origin of this CTA's pixel patch
    uint32_t taps;                               // convolution form: taps the CTA pair has to visit (bit ky*3 + kx)
};

// Taps of a 3x3 convolution that can reach the image from a TH x TW patch at (x0, y0): with dilation d the tap (ky, kx)
// reads rows y + (ky-1)*d, so a patch within d of the top edge never sees ky = 0 etc.  Skipped taps are whole k-block
// runs the pair neither loads nor multiplies (layer4 at dilation 4 on a 10x10 map: half of all tap visits).
__host__ __device__ __forceinline__ void q_patch(const TcParams& p, int pt, int& tx, int& ty, int& tb)
{
    // batch fastest: the two patches of a CTA pair normally sit at the same (x0, y0) of neighbouring crop groups and need the
    // same taps (measured against x-fastest with the union of two neighbouring patches: 3-8% on the dilated layers)
    tb = pt % p.tiles_b;
    const int rest = pt / p.tiles_b;
    tx = rest % p.tiles_x; ty = rest / p.tiles_x;
}

__host__ __device__ __forceinline__ uint32_t q_tap_mask(const TcParams& p, int x0, int y0)
{
    if (p.conv_taps != 9) return 1u;
    const int yh = (y0 + p.TH < p.cH ? y0 + p.TH : p.cH) - 1, xh = (x0 + p.TW < p.cW ? x0 + p.TW : p.cW) - 1;
    uint32_t my = 0, mx = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int d = (t - 1) * p.conv_dil;
        if (yh + d >= 0 && y0 + d < p.cH) my |= 1u << t;
        if (xh + d >= 0 && x0 + d < p.cW) mx |= 1u << t;
    }
    return ((my & 1u) ? mx : 0u) | ((my & 2u) ? mx << 3 : 0u) | ((my & 4u) ? mx << 6 : 0u);
}

// FORM: which launch form the instantiation serves -- -1 any (every feature tested at run time), 0 plain single-run GEMM, 1 implicit
// convolution, 2 pooled GEMM, 3 the conv1 gather form, 4 implicit convolution on the balanced schedule (QSched).  The specialised forms
// read the features they do not have as constants, so their code disappears from the role loops (see gemm_tc_q_kernel).
template <int FORM> struct QForm {
    static constexpr bool ANY = FORM < 0;
    static constexpr bool CONV = FORM == 1 || FORM == 4;
    __device__ static __forceinline__ int conv_taps(const TcParams& p) { return (ANY || CONV) ? p.conv_taps : 0; }
    __device__ static __forceinline__ float* pool_partial(const TcParams& p) { return (ANY || FORM == 2) ? p.pool_partial : nullptr; }
    __device__ static __forceinline__ int c1_H(const TcParams& p) { return (ANY || FORM == 3) ? p.c1_H : 0; }
    __device__ static __forceinline__ int wk_rows(const TcParams& p) { return ANY ? p.wk_rows : 0; }
    __device__ static __forceinline__ const float* residual(const TcParams& p) { return (ANY || CONV) ? p.residual : nullptr; }
    __device__ static __forceinline__ int bias_crop_stride(const TcParams& p) { return (ANY || FORM == 0 || FORM == 2) ? p.bias_crop_stride : 0; }
    __device__ static __forceinline__ int k_chunks(const TcParams& p) { return (ANY || CONV) ? p.k_chunks : 1; }
    __device__ static __forceinline__ int m_fastest(const TcParams& p) { return ANY ? p.m_fastest : 0; }
};

template <int CTAS, int FORM>
__device__ __forceinline__ QTile q_decode(const TcParams& p, int t, int m_tiles, int n_tiles, int bnt, int rank)
{
    using F = QForm<FORM>;
    QTile c;
    const int per_group = m_tiles * n_tiles;
    c.g = t / per_group;
    const int rem = t - c.g * per_group;
    int mt, nt;
    if (F::m_fastest(p)) { nt = rem / m_tiles; mt = rem - nt * m_tiles; }      // concurrent clusters share the weight tile
    else { mt = rem / n_tiles; nt = rem - mt * n_tiles; }                  // ... or the activation rows
    c.n0 = nt * bnt;
    c.crop = 0; c.pool_tile = 0;
    c.x0 = c.y0 = c.b0 = 0; c.taps = 1u;
    if (F::conv_taps(p)) {
        const int pt = mt * CTAS + rank;
        int tx, ty, tb;
        q_patch(p, pt, tx, ty, tb);
        c.x0 = tx * p.TW; c.y0 = ty * p.TH; c.b0 = tb * p.TB;      // a patch past the last one is fully masked
        c.row0 = 0;
        c.rows_valid = (ty < p.tiles_y && tb < p.tiles_b) ? p.TW * p.TH * p.TB : 0;
        c.taps = 0;
#pragma unroll
        for (int r = 0; r < CTAS; ++r) {                           // union over the pair: one MMA stream serves both patches
            q_patch(p, mt * CTAS + r, tx, ty, tb);
            if (ty < p.tiles_y && tb < p.tiles_b) c.taps |= q_tap_mask(p, tx * p.TW, ty * p.TH);
        }
    } else if (F::pool_partial(p)) {
        const int pairs_per_crop = (p.rows_per_crop + 128 * CTAS - 1) / (128 * CTAS);
        c.crop = mt / pairs_per_crop;
        c.pool_tile = (mt - c.crop * pairs_per_crop) * CTAS + rank;        // index of this CTA's 128-row tile in the crop
        c.row0 = c.crop * p.rows_per_crop + c.pool_tile * 128;
        c.rows_valid = max(0, min(128, p.rows_per_crop - c.pool_tile * 128));
    } else {
        c.row0 = (mt * CTAS + rank) * 128;
        c.rows_valid = max(0, min(128, p.M - c.row0));
    }
    return c;
}

// Epilogue store of one transposed 32 x 32 chunk, single-run fast path (no earlier partial sums in C, no skip connection, one bias
// vector for the warp's 32 rows): per 4-row step one LDS.128, the bias, the activation, one address multiply-add and one STG.128.
template <int ACT>
__device__ __forceinline__ void epi_store_simple(const float* srow, int sw0, int sw1, char* cbase, uint32_t ldcb, const int (&roff)[8],
                                                 float4 b, float slope, bool no_store = false)
{
    // the rows are read ahead of the stores, four at a time: with the row mask tested first every row was its own branch region and the
    // LDS -> FADD -> STG chains ran one after the other through the same four registers (ncu source view: ~70 short-scoreboard
    // samples on each of the eight FADDs, the largest item of the epilogue warps' stalls)
#pragma unroll
    for (int pb = 0; pb < 8; pb += 4) {
    float4 rows[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rows[i] = *reinterpret_cast<const float4*>(srow + (pb + i) * 128 + ((i & 1) ? sw1 : sw0));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ps = pb + i;
        if (roff[ps] < 0) continue;
        float4 o = rows[i];
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        if (ACT == 1) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        } else if (ACT == 2) {
            o.x = o.x > 0.f ? o.x : slope * o.x; o.y = o.y > 0.f ? o.y : slope * o.y;
            o.z = o.z > 0.f ? o.z : slope * o.z; o.w = o.w > 0.f ? o.w : slope * o.w;
        }
        if (no_store && o.x != 12345.678f) continue;          // (DF_TC_DBG bit 16: everything but the store instruction itself)
        *reinterpret_cast<float4*>(cbase + (unsigned long long)(uint32_t)roff[ps] * ldcb) = o;
    }
    }
}

// DBG: the knock-out bits and the clock64 timeline (TcParams::dbg / trace) exist in a second instantiation only, launched when
// DF_TC_DBG is set -- as run-time tests inside the role loops they cost registers and instructions on every path (measured: a few more
// conditionals and 24 bytes of additional spills made the short-K GEMMs 10-18% slower, profiles/r2_s3_variants.jsonl).
template <int CTAS, int A_STAGES, int A_COLS, bool DBG, int FORM>
__global__ void __launch_bounds__(Q_THREADS, 1)
gemm_tc_q_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_whi,
                 const __grid_constant__ CUtensorMap tm_wlo, const TcParams p, const int bn_cta, const int m_tiles,
                 const int n_tiles, const int total_tiles, const __grid_constant__ QSched sched, const __grid_constant__ CUtensorMap tm_c)
{
    // A_STAGES TMEM A stages of A_COLS columns each at the top of TMEM (64: [hi | lo] / [fp16 | - | bf16 | bf16 lo]; 32: the two
    // fp16 planes of hybrid16s, the only arithmetic that instantiation carries); the accumulators share what is left below.
    // A_COLS == 0 (hybrid16s, "A_SMEM"): the two fp16 planes of A go to a shared-memory ring of A_STAGES 16 KB tiles (4, with 4 operand
    // stages of a 256-wide tile; 3 when the pooled epilogue needs its 8 KB: measured -2..3% / +9% for the other choice) in the MMA's
    // K-major 128B-swizzled layout instead (rows of [hi x32 | lo x32], like the weight tile) and the MMA reads both operands from
    // shared memory -- all 512 TMEM columns are then accumulators: TWO buffers of up to 256 columns, so that the drain of a 256-wide tile
    // (one accumulator with A in TMEM: exposed, ~20% of the pooled conv6 layer, ~12% of the long-K convolutions) overlaps the next run.
    constexpr bool A_SMEM = A_COLS == 0;
    constexpr bool S16 = A_COLS == 32 || A_SMEM;
    const int dbg = DBG ? p.dbg : 0;
    // features of the launch form: constants in the specialised instantiations (QForm)
    using F = QForm<FORM>;
    const int conv_taps = F::conv_taps(p), c1_H = F::c1_H(p), wk_rows = F::wk_rows(p), bias_crop_stride = F::bias_crop_stride(p);
    float* const pool_partial = F::pool_partial(p);
    const float* const residual = F::residual(p);
    const int k_chunks = F::k_chunks(p);
    constexpr int TMEM_A0 = 512 - A_STAGES * A_COLS;
    constexpr int ACC_STRIDE = TMEM_A0 / 2;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET from the __shared__ symbol (not a round trip through uintptr_t): the compiler then knows every
    // pointer derived from `smem` is shared memory and emits LDS / STS instead of generic LD / ST (measured in the round-2 ncu capture:
    // all 2.9 M tile reads of the stagers and the epilogue were generic loads on the long scoreboard)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* s_epi = reinterpret_cast<float*>(smem + Q_SMEM_STAGES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_SMEM_STAGES + Q_SMEM_EPI);
    uint64_t* full = bars;                          // [8]  TMA bytes of this CTA's stage
    uint64_t* empty = bars + 8;                     // [8]  MMAs that read the stage have retired (commit)
    uint64_t* a_full = bars + 16;                   // [8]  A of the stage is in TMEM (all stager warps of the pair)
    uint64_t* acc_full = bars + 24;                 // [2]
    uint64_t* acc_empty = bars + 26;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / CTAS, ncl = gridDim.x / CTAS;
    const int nkb = p.K / BK;
    const int bnt = bn_cta * CTAS;                  // tile width = accumulator columns
    const uint32_t w_bytes = (uint32_t)bn_cta * BK * 4;
    // A | W_hi | W_lo (or the bf16 pair tile), all 1024-B aligned; hybrid16s: A | [fp16 hi x32 | fp16 lo x32] rows
    const uint32_t stage_bytes = S16 ? Q_TILE + w_bytes : Q_TILE + 2 * w_bytes;
    // (the pooled epilogue keeps its per-warp column sums in the last 8 KB of the stage area: one stage less when the ring fills it)
    const uint32_t pool_bytes = pool_partial ? 8192u : (c1_H ? 2048u : 0u);      // (conv1 form: the patch offset table lives there)
    const uint32_t aring_bytes = A_SMEM ? (uint32_t)A_STAGES * Q_TILE : 0u;
    // An EVEN number of stages: the two stager groups (and the two producer warps) take alternate k-blocks, so with an even ring a stage
    // always belongs to the same group and that group sees every fill of it; with an odd ring a group would see every OTHER fill -- all of
    // the same barrier parity -- and a group running ahead could take fill i for fill i + 2.
    int q_stages = (int)min((uint32_t)Q_MAX_STAGES, ((uint32_t)Q_SMEM_STAGES - pool_bytes - aring_bytes) / stage_bytes);
    if (q_stages > 2) q_stages &= ~1;
    const int Q_STAGES = q_stages;
    uint8_t* const aring = smem + (size_t)Q_STAGES * stage_bytes;           // A_SMEM: the operand-plane ring of A (1024-byte aligned)
    // tile walk of this cluster: round robin over all tiles, or (FORM 4) the contiguous range of the balanced schedule, which starts
    // at run sp_r0 of tile t_first and stops before run sp_r1 of tile t_end - 1
    constexpr bool SPLIT = FORM == 4;
    const int t_first = SPLIT ? (int)sched.t0[cid] : cid, t_step = SPLIT ? 1 : ncl;
    const int t_end = SPLIT ? (int)sched.t1[cid] + 1 : total_tiles;
    const int sp_r0 = SPLIT ? (int)sched.r0[cid] : 0, sp_r1 = SPLIT ? (int)sched.r1[cid] : 0x7fff;
    const uint32_t ACC_BUFS = bnt <= ACC_STRIDE ? 2 : 1;
    const uint32_t acc_shift = ACC_BUFS - 1;        // ti % ACC_BUFS == ti & acc_shift, ti / ACC_BUFS == ti >> acc_shift

    if (threadIdx.x == 0) {
        for (int i = 0; i < Q_MAX_STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); mbar_init(a_full + i, 4 * CTAS); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8 * CTAS); }
        tmem_slot[1] = 0u;                                         // hybrid16s: bit pattern of the sampled activation maximum
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CTAS == 2) tmem_alloc_pair(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();              // the peer's barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, cluster rendezvous) may overlap the
    // tail of the previous kernel in the stream; global memory is only touched after the wait.  Dependents are released
    // at once -- they block in their own griddepcontrol.wait until this grid has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ---- hybrid16s, activation scale: the A operand is split into fp16(a s) + fp16(a s - fp16(a s)).  The second plane is ~2^-12 of the
    // first: it stays a normal fp16 number while |a s| >= 0.25 and carries an ABSOLUTE error of 2^-25 below, and fp16 saturates at
    // 65504 -- so s = 2^k is chosen to put the operand's largest entries near 2^5: entries up to ~2000x larger still fit, and the
    // absolute error is 2^-30 of the largest entry (fp32 itself rounds that entry to 2^-24).  "Largest entry" is estimated from a fixed
    // sample of 512 (row, 4-column group) positions along a diagonal, twice, read by the 16 stager / epilogue warps of EVERY CTA (same addresses, hence the same scale everywhere
    // and from run to run; L2 hits after the first CTA) while the TMA producer is already filling the pipeline.
    float a_sc = 1.0f;
    if (S16 && warp >= 2 && warp < 18) {
        a_sc = p.a_scale;
        if (a_sc == 0.0f) {
            const int t = (int)threadIdx.x - 64;                                  // 0..511
            const float* row = p.A + (size_t)(((long long)t * p.samp_rows) >> 9) * p.lda;
            uint32_t m = 0;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                // (row and column both move with t: every group of 4 channels is visited when the operand has <= 2048 columns, so a
                // channel that is systematically larger than the rest cannot hide from the sample)
                const int col = (((t << 2) + j * (p.samp_cols >> 1)) % p.samp_cols) & ~3;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + col));
                const uint32_t b0 = v.x & 0x7fffffffu, b1 = v.y & 0x7fffffffu, b2 = v.z & 0x7fffffffu, b3 = v.w & 0x7fffffffu;
                if (b0 < 0x7f800000u) m = max(m, b0);
                if (b1 < 0x7f800000u) m = max(m, b1);
                if (b2 < 0x7f800000u) m = max(m, b2);
                if (b3 < 0x7f800000u) m = max(m, b3);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
            if (lane == 0 && m) atomicMax(tmem_slot + 1, m);
            asm volatile("bar.sync 2, 512;" ::: "memory");
            const int e = (int)(tmem_slot[1] >> 23) - 127;                        // floor(log2(sample maximum)); all-zero sample: -127
            a_sc = e < -100 ? 1.0f : exp2f((float)(5 - e));
        }
    }

    // Ring positions are kept as (slot, phase) counters: the stage count is a runtime value, and `it % Q_STAGES` in the issuing warp's
    // loop was an integer division (MUFU.RCP + ~20 dependent instructions) per k-block on the kernel's critical path.
    if (warp == 0 || warp == 18) {
        // ------------------------------- TMA producers (two warps, alternating k-blocks) -------------------------------
        // The clock64 timeline (DF_TC_DBG bit 256) showed ONE producer warp needing ~700 clocks per k-block -- ~450 of them between
        // "slot free" and "both TMA instructions issued" -- which paced every launch shape once the issuer's loop had been shortened;
        // two warps take the even / odd k-blocks of the same sequence (each waits for and fills its own slots).
        const uint32_t pw = warp == 0 ? 0u : 1u;
        const uint32_t a_bytes = conv_taps ? (uint32_t)(p.TW * p.TH * p.TB) * BK * 4 : (uint32_t)Q_TILE;
        // 3xTF32: W_hi + W_lo; hybrid: W_hi + bf16 pair tile; hybrid16: pair tile + a half-width (64 B rows) correction tile
        const uint32_t bytes = (c1_H ? 0u : a_bytes) + (S16 ? w_bytes : p.precise == 3 ? w_bytes + w_bytes / 2 : (p.precise ? 2u : 1u) * w_bytes);
        const int cblocks = conv_taps ? p.K / (BK * conv_taps) : 1;        // 32-channel blocks per tap
        int s = 0;
        uint32_t ph = 1;                                                       // parity of "slot is free": passes at once in round 0
        uint32_t pit = 0, e_ready = 0;
        for (int t = t_first; t < t_end; t += t_step) {
            const QTile c = q_decode<CTAS, FORM>(p, t, m_tiles, n_tiles, bnt, rank);
            int wrow = c.g * p.N + c.n0 + rank * bn_cta;
            int wk0 = 0;                                                       // k offset of the W operand (weight-gradient form)
            if (wk_rows) {                                                   // groups = slices of the reduction axis (split-K)
                wrow = c.n0 + rank * bn_cta;
                const int tap = wrow / wk_rows;
                wk0 = p.wk_shift[tap] + c.g * p.K;
                wrow += p.wk_row0[tap] - tap * wk_rows;
            }
            const int acol = (int)(c.g * p.a_gs);
            const int nkb_t = conv_taps ? __popc(c.taps) * cblocks : nkb;
            uint64_t tap_list = 0;                                              // the taps the pair visits, 4 bits each, in order
            for (int tap = 8; tap >= 0; --tap)
                if ((c.taps >> tap) & 1u) tap_list = (tap_list << 4) | (uint64_t)tap;
            int cb = 0;
            int j0 = 0, j1 = nkb_t;
            if (SPLIT) {                                                        // this cluster's runs of a shared tile
                if (t == t_first) j0 = sp_r0 * p.kbc;
                if (t == t_end - 1) j1 = min(nkb_t, sp_r1 * p.kbc);
                if (j0) { const int ord = j0 / cblocks; cb = j0 - ord * cblocks; tap_list >>= 4 * ord; }
            }
            for (int j = j0; j < j1; ++j, ++pit) {
                const bool mine = (pit & 1u) == pw;
                if (mine) {
                if (!e_ready) mbar_wait(empty + s, ph);
                DF_TRACE(0, pit);
                {                                                               // (state of MY next slot: round trip overlaps the TMA issue)
                    int sn = s + 2; uint32_t phn = ph;
                    if (sn >= Q_STAGES) { sn -= Q_STAGES; phn ^= 1u; }
                    e_ready = mbar_try(empty + sn, phn);
                }
                }
                if (mine && elect_one()) {
                    // (DF_TC_DBG bit 128: the activation box is fetched for the first tap of a channel block only -- what a 3x3 convolution
                    // would pull through the port if the nine taps shared one haloed box; timing experiment, wrong results)
                    const bool skip_a = (dbg & 128) && conv_taps == 9 && j >= cblocks;
                    mbar_expect_tx(full + s, skip_a ? bytes - a_bytes : bytes);
                    uint8_t* dst = smem + (size_t)s * stage_bytes;
                    int kb = j;                                                 // k-block of the weight matrix
                    if (conv_taps) {
                        const int tap = (int)(tap_list & 15u);
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const int dy = conv_taps == 9 ? (ky - 1) * p.conv_dil : 0;
                        const int dx = conv_taps == 9 ? (kx - 1) * p.conv_dil : 0;
                        kb = tap * cblocks + cb;
                        if (!skip_a) tma_load_4d(&tm_a, dst, full + s, cb * BK, c.x0 + dx, c.y0 + dy, c.b0);
                    } else if (!c1_H) {
                        tma_load_2d(&tm_a, dst, full + s, acol + kb * BK, c.row0);
                    }
                    // first weight tile: 32 fp32 per row (TF32 hi part), or, hybrid16 / hybrid16s, 64 halves per row ([fp16(W) x32 | bf16(W) x32]
                    // / [fp16(W s) x32 | fp16 remainder x32])
                    tma_load_2d(&tm_whi, dst + Q_TILE, full + s, kb * (p.precise >= 3 ? 2 * BK : BK) + wk0, wrow);
                    // second weight tile: W_lo (fp32), or [bf16(W) x32 | bf16(W_lo) x32] (hybrid), or bf16(W_lo) x32 in 64-byte rows (hybrid16)
                    if (!S16 && p.precise) tma_load_2d(&tm_wlo, dst + Q_TILE + w_bytes, full + s, kb * (p.precise == 2 ? 2 * BK : BK) + wk0, wrow);
                }
                if (++cb == cblocks) { cb = 0; tap_list >>= 4; }
                if (++s == Q_STAGES) { s = 0; ph ^= 1u; }
                __syncwarp();
                if (mine) DF_TRACE(1, pit);
            }
        }
    } else if (warp == 1) {
        // ------------------------------- MMA issuer (leader CTA) --------------------
        if (rank == 0) {
            const uint32_t idesc = tf32_instr_desc(bnt, 128 * CTAS);
            const uint32_t idesc_bf = bf16_instr_desc(bnt, 128 * CTAS);
            const uint32_t idesc_h = f16_instr_desc(bnt, 128 * CTAS);
            uint32_t it = 0, ti = 0;
            int s = 0;
            uint32_t ph = 0, a_ready = 0;
            const int cblocks = conv_taps ? p.K / (BK * conv_taps) : 1;
            for (int t = t_first; t < t_end; t += t_step) {
              const int nkb_t = conv_taps == 9 ? __popc(q_decode<CTAS, FORM>(p, t, m_tiles, n_tiles, bnt, 0).taps) * cblocks : nkb;
              const int runs = (nkb_t + p.kbc - 1) / p.kbc;
              const int ra = (SPLIT && t == t_first) ? sp_r0 : 0, rb = (SPLIT && t == t_end - 1) ? min(runs, sp_r1) : runs;
              for (int kc = ra; kc < rb; ++kc, ++ti) {
                const int kb0 = kc * p.kbc, kb1 = min(nkb_t, kb0 + p.kbc);
                const uint32_t ab = ti & acc_shift;
                const uint32_t aph = ((ti >> acc_shift) & 1) ^ 1;
                DF_TRACE(12, ti);
                if (CTAS == 2) mbar_wait_cluster(acc_empty + ab, aph); else mbar_wait(acc_empty + ab, aph);
                DF_TRACE(13, ti);
                const uint32_t acc = tmem_base + ab * ACC_STRIDE;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    // a_full[s] alone: the stagers of BOTH CTAs arrive on it only after their own full[s] (A and weight tile of the
                    // stage, one barrier) has completed, so it implies that every byte the MMAs of this k-block read has landed --
                    // the peer's half of the weight tile never had another guard.  (Measured with the clock64 timeline, DF_TC_DBG
                    // bit 256: this warp's serial loop, ~900 clocks per k-block with two ~200-clock barrier round trips in it, paced
                    // tower-1 -- not the tensor pipe, not the operand stream.)  The state of the NEXT slot is probed before the MMAs
                    // of this one are issued, so its round trip overlaps them.
                    DF_TRACE(2, it);
                    if (!a_ready) mbar_wait(a_full + s, ph);
                    DF_TRACE(3, it);
                    {
                        int sn = s + 1; uint32_t phn = ph;
                        if (sn == Q_STAGES) { sn = 0; phn ^= 1u; }
                        a_ready = mbar_try(a_full + sn, phn);
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t w_hi = smem_u32(smem + (size_t)s * stage_bytes + Q_TILE);
                        const uint64_t bhi0 = sw128_desc(w_hi), blo0 = sw128_desc(w_hi + w_bytes);
                        const uint32_t a0 = tmem_base + TMEM_A0 + (it % A_STAGES) * A_COLS;
                        if (dbg & 8) {
                        } else if (A_SMEM) {
                            // hybrid16s with A in shared memory: rows [fp16(a) x32 | fp16(a - fp16(a)) x32], same layout as the weight tile
                            const uint64_t ad0 = sw128_desc(smem_u32(aring + (size_t)(it % A_STAGES) * Q_TILE));
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t b_h = bhi0 + (uint64_t)(j * 2), b_l = bhi0 + (uint64_t)(4 + j * 2);
                                const uint64_t a_h = ad0 + (uint64_t)(j * 2), a_l = ad0 + (uint64_t)(4 + j * 2);
                                const uint32_t acc_on = (kb != kb0) || (j != 0);
                                umma_ss_f16_pair(acc, a_h, b_h, idesc_h, acc_on);
                                umma_ss_f16_pair(acc, a_l, b_h, idesc_h, 1u);
                                umma_ss_f16_pair(acc, a_h, b_l, idesc_h, 1u);
                            }
                        } else if (S16) {
                            // hybrid16s: every term on fp16 operands.  TMEM A stage: [0,16) fp16(a) pairs | [16,32) fp16(a - fp16(a));
                            // weight tile rows: 64 B of fp16(w) then 64 B of fp16(w - fp16(w)) (descriptor +4 = +64 B).
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t b_h = bhi0 + (uint64_t)(j * 2), b_l = bhi0 + (uint64_t)(4 + j * 2);
                                const uint32_t acc_on = (kb != kb0) || (j != 0);
                                umma_ts_bf16_pair(acc, a0 + j * 8, b_h, idesc_h, acc_on);
                                umma_ts_bf16_pair(acc, a0 + 16 + j * 8, b_h, idesc_h, 1u);
                                umma_ts_bf16_pair(acc, a0 + j * 8, b_l, idesc_h, 1u);
                            }
                        } else if (p.precise == 3) {
                            // hybrid16: the main term on fp16 operands (11 significant bits like TF32, K = 16 per instruction:
                            // half the tensor time), the two correction terms on bf16 as in the hybrid mode -- 2 + 4
                            // instructions per k-block.  TMEM A stage: [0,16) fp16(a) pairs | [32,48) bf16(a) | [48,64)
                            // bf16(a - fp16(a)); weight tile 1 rows: fp16(w) then bf16(w); tile 2 (64-byte rows, 64B swizzle): bf16(w - fp16(w)).
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t b_h = bhi0 + (uint64_t)(j * 2), b_w = bhi0 + (uint64_t)(4 + j * 2);
                                const uint64_t b_wlo = sw64_desc(w_hi + w_bytes) + (uint64_t)(j * 2);
                                const uint32_t acc_on = (kb != kb0) || (j != 0);
                                if (CTAS == 2) {
                                    umma_ts_bf16_pair(acc, a0 + j * 8, b_h, idesc_h, acc_on);
                                    umma_ts_bf16_pair(acc, a0 + 48 + j * 8, b_w, idesc_bf, 1u);
                                    umma_ts_bf16_pair(acc, a0 + 32 + j * 8, b_wlo, idesc_bf, 1u);
                                } else {
                                    umma_ts_bf16(acc, a0 + j * 8, b_h, idesc_h, acc_on);
                                    umma_ts_bf16(acc, a0 + 48 + j * 8, b_w, idesc_bf, 1u);
                                    umma_ts_bf16(acc, a0 + 32 + j * 8, b_wlo, idesc_bf, 1u);
                                }
                            }
                        } else {
#pragma unroll
                        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                            const uint64_t bhi = bhi0 + (uint64_t)(ks * 2), blo = blo0 + (uint64_t)(ks * 2);
                            const uint32_t a_hi = a0 + ks * UMMA_K;
                            const uint32_t acc_on = (kb != kb0) || (ks != 0);
                            if (CTAS == 2) {
                                umma_ts_pair(acc, a_hi, bhi, idesc, acc_on);
                                if (p.precise == 1) { umma_ts_pair(acc, a_hi + BK, bhi, idesc, 1u); umma_ts_pair(acc, a_hi, blo, idesc, 1u); }
                            } else {
                                umma_ts(acc, a_hi, bhi, idesc, acc_on);
                                if (p.precise == 1) { umma_ts(acc, a_hi + BK, bhi, idesc, 1u); umma_ts(acc, a_hi, blo, idesc, 1u); }
                            }
                        }
                        }
                        if (!S16 && p.precise == 2 && !(dbg & 8)) {
                            // hybrid: the two correction terms are ~2^-11 of the main one, so bf16 operands (K = 16 per
                            // instruction, twice the TF32 rate) keep them to 2^-20 of the result: 4 + 4 instructions per
                            // k-block instead of 12.  TMEM A stage: [0,32) tf32 hi | [32,48) bf16(a) pairs | [48,64) bf16(a_lo);
                            // weight tile rows: 64 B of bf16(w) then 64 B of bf16(w_lo) (descriptor +4 = +64 B).
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const uint64_t b_w = blo0 + (uint64_t)(j * 2), b_wlo = blo0 + (uint64_t)(4 + j * 2);
                                if (CTAS == 2) {
                                    umma_ts_bf16_pair(acc, a0 + 48 + j * 8, b_w, idesc_bf, 1u);
                                    umma_ts_bf16_pair(acc, a0 + 32 + j * 8, b_wlo, idesc_bf, 1u);
                                } else {
                                    umma_ts_bf16(acc, a0 + 48 + j * 8, b_w, idesc_bf, 1u);
                                    umma_ts_bf16(acc, a0 + 32 + j * 8, b_wlo, idesc_bf, 1u);
                                }
                            }
                        }
                        if (CTAS == 2) {
                            umma_commit_pair(empty + s);
                            if (kb == kb1 - 1) umma_commit_pair(acc_full + ab);
                        } else {
                            umma_commit(empty + s);
                            if (kb == kb1 - 1) umma_commit(acc_full + ab);
                        }
                    }
                    if (++s == Q_STAGES) { s = 0; ph ^= 1u; }
                    __syncwarp();
                    DF_TRACE(4, it);
                }
              }
            }
        }
    } else if (warp < 10) {
        // ------------------------------- A stagers (two groups) ----------------------
        const int my_tiles = (t_end - t_first + t_step - 1) / t_step;
        uint32_t total_it = (uint32_t)my_tiles * nkb;
        if (conv_taps == 9 || SPLIT) {                                 // tiles near the border visit fewer taps
            const int cblocks = conv_taps == 9 ? p.K / (BK * 9) : nkb;
            total_it = 0;
            for (int t = t_first + lane * t_step; t < t_end; t += 32 * t_step) {
                const int nkb_t = conv_taps == 9 ? __popc(q_decode<CTAS, FORM>(p, t, m_tiles, n_tiles, bnt, 0).taps) * cblocks : nkb;
                int j0 = 0, j1 = nkb_t;
                if (SPLIT) {
                    if (t == t_first) j0 = sp_r0 * p.kbc;
                    if (t == t_end - 1) j1 = min(nkb_t, sp_r1 * p.kbc);
                }
                total_it += j1 - j0;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) total_it += __shfl_xor_sync(0xffffffffu, total_it, off);
        }
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int sw = r & 7;
        int s = grp;                                                     // slot / phase of iteration `it` ...
        uint32_t ph = 0;
        int sp = grp - A_STAGES;                                         // ... and of iteration it - A_STAGES (negative: none yet)
        uint32_t php = 0;
        // conv1 form: per k the offset of patch element (c, ky, kx) relative to the patch origin and its (ky - 3, kx - 3), built once
        // in shared memory (a __constant__ table indexed by the warp-uniform k was measured slower: 0.175 vs 0.153 ms at 64 x 160^2)
        int* c1_off = reinterpret_cast<int*>(smem + Q_SMEM_STAGES - 2048);
        int* c1_dyx = c1_off + 160;
        const bool conv1 = S16 && !A_SMEM && c1_H != 0;
        if (conv1) {
            const int k = (int)threadIdx.x - 64;                         // the 256 stager threads
            if (k < 160) {
                const int c = k / 49, rem = k - c * 49, ky = rem / 7, kx = rem - ky * 7;
                c1_off[k] = k < 147 ? (c * c1_H + (ky - 3)) * p.c1_W + (kx - 3) : 0;
                c1_dyx[k] = k < 147 ? (((ky - 3) & 0xff) | (((kx - 3) & 0xff) << 8)) : 0x8080;      // 0x80: never inside the image
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        int c1_tile = 0, c1_kb = grp;                                    // conv1 form: tile of this cluster / k-block inside it (nkb >= 2)
        for (uint32_t it = grp; it < total_it; it += 2) {
            // hi: TF32-exact part (raw value in single-pass mode); second[]: what goes into columns [32,64) of the TMEM stage --
            // lo (3xTF32), or 16 words of bf16(x) pairs followed by 16 words of bf16(lo) pairs (hybrid)
            uint32_t hi[32], second[32];
            if (conv1) {
                const long long m = ((long long)(t_first + c1_tile * t_step) * CTAS + rank) * 128 + r;       // output pixel of this lane
                const bool row_in = m < (long long)p.M;
                const int hw = p.c1_Ho * p.c1_Wo;
                const int b = row_in ? (int)(m / hw) : 0, rem = row_in ? (int)(m - (long long)b * hw) : 0;
                const int yo = rem / p.c1_Wo, xo = rem - yo * p.c1_Wo;
                const int y2 = yo * 2, x2 = xo * 2;
                const float* org = p.A + ((size_t)b * 3 * c1_H + y2) * p.c1_W + x2;      // patch origin + (3, 3)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int k = c1_kb * 32 + i;
                    const int d = c1_dyx[k];
                    const int y = y2 + (int)(signed char)(d & 0xff), x = x2 + (int)(signed char)((d >> 8) & 0xff);
                    const bool ok = row_in && (unsigned)y < (unsigned)c1_H && (unsigned)x < (unsigned)p.c1_W && d != 0x8080;
                    hi[i] = ok ? __float_as_uint(__ldg(org + c1_off[k])) : 0u;
                }
                c1_kb += 2;
                if (c1_kb >= nkb) { c1_kb -= nkb; ++c1_tile; }
            } else {
            mbar_wait(full + s, ph);
            if (q == 0) DF_TRACE(5, it);
            const uint8_t* arow = smem + (size_t)s * stage_bytes + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(arow + ((c ^ sw) << 4));
                hi[c * 4 + 0] = v.x; hi[c * 4 + 1] = v.y; hi[c * 4 + 2] = v.z; hi[c * 4 + 3] = v.w;
            }
            }
            if (S16 && (dbg & 32)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) second[j] = hi[j];
            } else if (S16) {
                const float sa = a_sc;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x0 = __uint_as_float(hi[2 * j]) * sa, x1 = __uint_as_float(hi[2 * j + 1]) * sa;
                    const uint32_t h = pack_f16x2_sat(x0, x1);
                    float f0, f1;
                    unpack_f16x2(h, f0, f1);
                    second[j] = h;
                    second[16 + j] = pack_f16x2_sat(x0 - f0, x1 - f1);
                }
            } else if (p.precise == 3) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x0 = __uint_as_float(hi[2 * j]), x1 = __uint_as_float(hi[2 * j + 1]);
                    const uint32_t h = pack_f16x2_sat(x0, x1);
                    float f0, f1;
                    unpack_f16x2(h, f0, f1);
                    second[j] = pack_bf16x2(x0, x1);
                    second[16 + j] = pack_bf16x2(x0 - f0, x1 - f1);
                    hi[j] = h;                                     // (j <= 2j: the words read above are never overwritten early)
                }
#pragma unroll
                for (int j = 16; j < 32; ++j) hi[j] = 0u;
            } else if (p.precise == 2) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float x0 = __uint_as_float(hi[2 * j]), x1 = __uint_as_float(hi[2 * j + 1]);
                    hi[2 * j] &= 0xffffe000u; hi[2 * j + 1] &= 0xffffe000u;
                    second[j] = pack_bf16x2(x0, x1);
                    second[16 + j] = pack_bf16x2(x0 - __uint_as_float(hi[2 * j]), x1 - __uint_as_float(hi[2 * j + 1]));
                }
            } else if (p.precise == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float x = __uint_as_float(hi[i]);
                    hi[i] &= 0xffffe000u;
                    second[i] = __float_as_uint(x - __uint_as_float(hi[i]));
                }
            }
            // TMEM A slot it % A_STAGES was last read by the MMAs of iteration it - A_STAGES.  With A_STAGES >= Q_STAGES that is
            // implied by full[s] (the commit that frees the slot is what let the TMA refill the stage); with fewer TMEM slots wait
            // for that iteration's commit explicitly (same barrier the TMA producer watches; the loads and the split above overlap it).
            if (q == 0) DF_TRACE(6, it);
            if (sp >= 0 && A_STAGES < Q_STAGES) mbar_wait(empty + sp, php);
            if (q == 0) DF_TRACE(7, it);
            if (A_SMEM) {
                uint8_t* prow = aring + (size_t)(it % A_STAGES) * Q_TILE + r * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<uint4*>(prow + ((c ^ sw) << 4)) = make_uint4(second[c * 4], second[c * 4 + 1], second[c * 4 + 2], second[c * 4 + 3]);
                // generic-proxy writes -> visible to the MMAs (async proxy).  The shared::cta form on purpose: the unqualified
                // fence.proxy.async costs a MEMBAR.ALL.GPU per k-block (measured in round 2: tower-1 0.56 ms instead of 0.30)
                fence_proxy_async();
            } else {
                tc_fence_after();
                const uint32_t ta = tmem_base + lane_base + TMEM_A0 + (it % A_STAGES) * A_COLS;
                if (dbg & 64) { if (second[0] == 0x12345678u) tmem_st32(ta, second); }
                else if (S16) tmem_st32(ta, second);
                else {
                    tmem_st32(ta, hi);
                    if (p.precise) tmem_st32(ta + BK, second);
                }
                tmem_st_wait();
                tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) {
                if (CTAS == 2) mbar_arrive_remote(a_full + s, 0); else mbar_arrive(a_full + s);
            }
            if (q == 0) DF_TRACE(8, it);
            s += 2; if (s >= Q_STAGES) { s -= Q_STAGES; ph ^= 1u; }
            sp += 2; if (sp >= Q_STAGES) { sp -= Q_STAGES; php ^= 1u; }
        }
    } else {
        // ------------------------------- epilogue (8 warps) --------------------------
        const int ew = warp - 10;
        const int q = warp & 3;
        const int half = ew >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float* stage = s_epi + ew * 1024;
        const int nchunks = bnt / 32;
        // store path: after the transpose lane (rr, cc) owns 4 columns of rows ps*4 + rr, ps = 0..7, of this warp's 32 rows; the
        // swizzle of row ps*4 + rr is rr (+ 4 for odd ps), so the tile is read at two fixed offsets plus an immediate per row
        const int rr = lane >> 3, cc = lane & 7;
        const float* srow = stage + rr * 32;
        const int sw0 = (cc ^ rr) << 2, sw1 = (cc ^ (rr + 4)) << 2;
        const uint32_t ldcb = (uint32_t)p.ldc * 4u, ldrb = (uint32_t)p.ldr * 4u;
        const float slope = p.relu == 2 ? __ldg(p.prelu) : 0.0f;
        const float w_inv = S16 ? __ldg(p.w_inv_scale) / a_sc : 1.0f;          // (powers of two: exact)
        const int per_kb = p.precise == 1 ? 12 : (p.precise == 2 ? 8 : (p.precise >= 3 ? 6 : 4));
        uint32_t ti = 0;
        for (int t = t_first; t < t_end; t += t_step) {
          const QTile c = q_decode<CTAS, FORM>(p, t, m_tiles, n_tiles, bnt, rank);
          const int nkb_t = conv_taps == 9 ? __popc(c.taps) * (p.K / (BK * 9)) : nkb;
          const int runs = conv_taps == 9 ? (nkb_t + p.kbc - 1) / p.kbc : k_chunks;
          // balanced schedule: a tile shared with the neighbouring cluster -- `tail_part`: this cluster has its LAST runs (and does them
          // first: all of them intermediate, the first one stores), `head_part`: its FIRST runs (after the neighbour's: all of them add, the
          // last one finishes the tile)
          int ra = 0, rb = runs;
          bool tail_part = false, head_part = false;
          if (SPLIT) {
              if (t == t_first && sp_r0 > 0) { ra = sp_r0; tail_part = true; }
              if (t == t_end - 1 && sp_r1 < runs) { rb = sp_r1; head_part = true; }
          }
          const int r = q * 32 + lane;
          const bool row_ok = r < c.rows_valid;
          const float* bias = p.bias ? p.bias + c.g * p.bias_gs : nullptr;
          // Per tile: where this lane's rows live in C (row / pixel index, -1 = masked) and which bias row they use.
          int roff[8];
          int myrow = -1;                                      // the row this lane holds after the TMEM load (lane == row)
          int crop_first = 0, crop_boundary = 0x7fffffff;
          bool straddle = false;
          if (!pool_partial) {
              if (conv_taps) {
                  if (runs > 1) {
                      const int rx = r % p.TW, rest = r / p.TW;
                      const int ry = rest % p.TH, rb = rest / p.TH;
                      const int x = c.x0 + rx, y = c.y0 + ry, b = c.b0 + rb;
                      if (row_ok && x < p.cW && y < p.cH && b < p.cB) myrow = (b * p.cH + y) * p.cW + x;
                  }
#pragma unroll
                  for (int ps = 0; ps < 8; ++ps) {
                      const int R = q * 32 + ps * 4 + rr;
                      const int rx = R % p.TW, rest = R / p.TW;
                      const int ry = rest % p.TH, rb = rest / p.TH;
                      const int x = c.x0 + rx, y = c.y0 + ry, b = c.b0 + rb;
                      const bool ok = R < c.rows_valid && x < p.cW && y < p.cH && b < p.cB;
                      roff[ps] = ok ? (b * p.cH + y) * p.cW + x : -1;
                  }
              } else {
#pragma unroll
                  for (int ps = 0; ps < 8; ++ps) {
                      const int R = q * 32 + ps * 4 + rr;
                      roff[ps] = R < c.rows_valid ? c.row0 + R : -1;
                      if ((dbg & 512) && roff[ps] >= 0) roff[ps] &= 1023;      // (knock-out: the output folded onto 1024 rows -- it stays in L2)
                  }
                  if (row_ok) myrow = c.row0 + r;
                  if (bias && bias_crop_stride) {
                      // clamped to the last crop: the warp's 32 rows may lie entirely in the masked tail of the last M tile
                      // (M = 3000: rows 3040..3071), and its bias vector is loaded before the row masks are looked at -- one
                      // row past the end of the (crops x N) bias buffer.  That read was the "ConvS2Fn" illegal address of
                      // round 1: harmless while the allocator happens to map the following bytes, a fault when it does not.
                      crop_first = min((c.row0 + q * 32) / p.rows_per_crop, (p.M - 1) / p.rows_per_crop);
                      crop_boundary = (crop_first + 1) * p.rows_per_crop;
                      straddle = c.row0 + q * 32 + 31 >= crop_boundary && crop_boundary < p.M;
                      bias += (size_t)crop_first * bias_crop_stride;
                  }
              }
          } else if (bias && bias_crop_stride) {
              bias += (size_t)c.crop * bias_crop_stride;          // pooled tiles are crop-aligned: one bias row per tile
          }
          float* const Cg = p.C + c.g * p.c_gs;
          // TMA-store epilogue (plain single-run GEMM form): lane == row keeps its 32 columns; its bias row is the one of its crop
          const bool tstore = FORM == 0 && p.tstore;
          const float* brow = nullptr;
          if (tstore && p.bias) {
              brow = p.bias + c.g * p.bias_gs;
              if (bias_crop_stride) brow += (size_t)(min(c.row0 + r, p.M - 1) / p.rows_per_crop) * bias_crop_stride;
          }
          for (int kc = ra; kc < rb; ++kc, ++ti) {
            const bool first_run = SPLIT ? (tail_part ? kc == ra : (!head_part && kc == 0)) : kc == 0;
            const bool last_run = SPLIT ? (!tail_part && kc == rb - 1) : kc == runs - 1;
            const uint32_t ab = ti & acc_shift;
            float* pool = reinterpret_cast<float*>(smem + Q_SMEM_STAGES - 8192) + (ti & 1) * 4 * 256;      // [2][4 lane quarters][256 columns]
            float run_scale = w_inv;
            if (p.bias_comp != 0.0f) {
                const int n_kb = min(nkb_t, (kc + 1) * p.kbc) - kc * p.kbc;
                run_scale *= 1.0f + p.bias_comp * (float)(n_kb * per_kb);
            }
            // the simple store: one run, no skip connection, one bias vector for all 32 rows of the warp
            const bool simple = first_run && last_run && !residual && !straddle;
            if (ew == 0) DF_TRACE(9, ti);
            mbar_wait(acc_full + ab, (ti >> acc_shift) & 1);
            if (ew == 0) DF_TRACE(10, ti);
            tc_fence_after();
            if (SPLIT && head_part && kc == 0) {
                // the neighbour's runs of this tile (same warp index there: same rows and columns) must be in C before this warp's first add
                unsigned int* flag = sched.flags + ((size_t)cid * 2 + rank) * 8 + ew;
                if (lane == 0) {
                    unsigned int v;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v == 0u);
                    *flag = 0u;                                // (only this warp reads it; the next launch finds it clear)
                }
                __syncwarp();
                __threadfence();
            }
            if (last_run && !first_run) {                      // the partial sums were written / added by OTHER lanes of this warp
                __threadfence();
                __syncwarp();
            }
#pragma unroll 1
            for (int ch = half; ch < nchunks; ch += 2) {
                uint32_t v[32];
                const int col = c.n0 + ch * 32;
                if (FORM == 0 && tstore) {
                    // The chunk goes to memory as ONE TMA store of the warp's 32 x 32 tile: every lane finishes its own row (scale, bias,
                    // activation -- the same operations in the same order as the store path below, so the results are bit-identical),
                    // writes it to the swizzled staging tile (the layout a 128B-swizzle box expects) and lane 0 hands the tile to the
                    // TMA unit, which clips rows past M and columns past N.  No transpose read-back, no per-row address arithmetic, no
                    // store instructions in the warp: the clock64 timeline put one chunk at ~3 400 clocks, 1 800 of them between
                    // "transposed" and "stores issued" -- three chunks per tile were as long as tower-1's whole k loop.
                    if (col >= p.N) continue;
                    tmem_ld32(tmem_base + lane_base + ab * ACC_STRIDE + ch * 32, v);
                    if (lane == 0) bulk_wait_group_read0();            // the previous chunk's store has read the staging tile
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (brow && col + j * 4 < p.N) b = __ldg(reinterpret_cast<const float4*>(brow + col + j * 4));
                        float4 o = make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]), __uint_as_float(v[j * 4 + 2]),
                                               __uint_as_float(v[j * 4 + 3]));
                        if (run_scale != 1.0f) {
                            o.x = __fmul_rn(o.x, run_scale); o.y = __fmul_rn(o.y, run_scale);
                            o.z = __fmul_rn(o.z, run_scale); o.w = __fmul_rn(o.w, run_scale);
                        }
                        o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
                        if (p.relu == 1) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = o;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && c.row0 + q * 32 < p.M) {
                        tma_store_2d(&tm_c, stage, (int)(c.g * p.c_gs) + col, c.row0 + q * 32);
                        bulk_commit_group();
                    }
                    continue;
                }
                const int cq = col + cc * 4;
                // this lane's 4 bias values (requested before the accumulator is read: the L2 round trip overlaps the TMEM load and the
                // transpose): one vector, or two when the warp's 32 rows straddle a crop boundary
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (bias && last_run && cq < p.N) {
                    b0 = __ldg(reinterpret_cast<const float4*>(bias + cq));
                    if (straddle) b1 = __ldg(reinterpret_cast<const float4*>(bias + bias_crop_stride + cq));
                }
                if (dbg & 4) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                } else
                tmem_ld32(tmem_base + lane_base + ab * ACC_STRIDE + ch * 32, v);
                if (ew == 0) DF_TRACE(14, ti * 4 + (ch >> 1));
                if (run_scale != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * run_scale);
                }
                if (!last_run && !pool_partial) {
                    // Partial sum of an intermediate run, straight from the registers (lane == row, 32 consecutive columns = one
                    // 128-byte line per lane): the first run is stored, the following ones are ADDED in L2 (red.global: no read round
                    // trip, no transpose) -- one thread per element in run order, so the sum is deterministic and, fp32 addition being
                    // commutative, the same as read-add-write.  Only the last run takes the full path below (the wide tiles have ONE
                    // accumulator: measured before this, the MMA issuer waited 43% of layer4.0 for the drains of its four runs).
                    if (myrow >= 0) {
                        float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(Cg + col) + (unsigned long long)(uint32_t)myrow * ldcb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (col + j * 4 >= p.N) break;
                            const float4 o = make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]), __uint_as_float(v[j * 4 + 2]),
                                                         __uint_as_float(v[j * 4 + 3]));
                            if (first_run) *reinterpret_cast<float4*>(dst + j * 4) = o;
                            else red_add_v4(dst + j * 4, o);
                        }
                    }
                    continue;
                }
                if (pool_partial) {
                    // column sums of act(x + bias) over this warp's 32 rows: through the same swizzled transpose as the store path
                    // (lane (rr, cc) then owns 4 columns of rows ps*4 + rr: 8 adds per column and two shuffle stages across rr, instead
                    // of a 31-shuffle butterfly over registers + 32 scalar bias loads per lane -- the pooled drain was 11.8 k clocks
                    // per tile against 16.6 k of main loop on a tile that has only one accumulator)
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                            make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                    __syncwarp();
                    const int lim = c.rows_valid - (q * 32 + rr);        // row ps*4 + rr of this warp is real iff ps*4 < lim
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int ps = 0; ps < 8; ++ps) {
                        if (ps * 4 >= lim) continue;
                        float4 o = *reinterpret_cast<const float4*>(srow + ps * 128 + ((ps & 1) ? sw1 : sw0));
                        o.x += b0.x; o.y += b0.y; o.z += b0.z; o.w += b0.w;
                        if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
                    }
#pragma unroll
                    for (int off = 8; off <= 16; off <<= 1) {
                        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
                    }
                    if (rr == 0) *reinterpret_cast<float4*>(pool + q * 256 + ch * 32 + cc * 4) = acc;
                } else {
                    // transpose through a swizzled 32x32 tile: lane == row on the way in, 8 lanes == one 128 B row out;
                    // bias / skip connection / activation are applied on the way out (coalesced float4 accesses)
                    __syncwarp();
                    if (!(dbg & 2)) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) =
                            make_uint4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                    }
                    __syncwarp();
                    if (ew == 0) DF_TRACE(15, ti * 4 + (ch >> 1));
                    if (cq < p.N && !(dbg & 1)) {
                        char* cbase = reinterpret_cast<char*>(Cg + cq);
                        if (simple) {
                            const bool ns = (dbg & 16) != 0;
                            if (p.relu == 1) epi_store_simple<1>(srow, sw0, sw1, cbase, ldcb, roff, b0, slope, ns);
                            else if (p.relu == 2) epi_store_simple<2>(srow, sw0, sw1, cbase, ldcb, roff, b0, slope, ns);
                            else epi_store_simple<0>(srow, sw0, sw1, cbase, ldcb, roff, b0, slope, ns);
                        } else {
                            // general store: the sum of the earlier runs of this tile comes back from C (written by this very thread)
                            // in the LAST run only, the skip connection from the residual tensor.  All loads of four rows are issued before the first store:
                            // left in program order every C read waited for the store before it (the compiler cannot prove
                            // the rows distinct), one L2 round trip per row -- 8 x 4 per 256-wide run, which the single accumulator of
                            // the wide tiles exposes in full (measured: the MMA issuer waited 43% of layer4.0 on acc_empty).
                            const char* rbase = (residual && last_run) ? reinterpret_cast<const char*>(residual + cq) : nullptr;
#pragma unroll
                            for (int pb = 0; pb < 8; pb += 4) {
                                float4 prev[4], resv[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int ps = pb + i;
                                    prev[i] = make_float4(0.f, 0.f, 0.f, 0.f); resv[i] = prev[i];
                                    if (roff[ps] < 0) continue;
                                    if (!first_run) prev[i] = __ldcg(reinterpret_cast<const float4*>(cbase + (unsigned long long)(uint32_t)roff[ps] * ldcb));   // (L2: where the adds happened)
                                    if (rbase) resv[i] = __ldg(reinterpret_cast<const float4*>(rbase + (unsigned long long)(uint32_t)roff[ps] * ldrb));
                                }
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int ps = pb + i;
                                    if (roff[ps] < 0) continue;
                                    float4 o = *reinterpret_cast<const float4*>(srow + ps * 128 + ((ps & 1) ? sw1 : sw0));
                                    float* dst = reinterpret_cast<float*>(cbase + (unsigned long long)(uint32_t)roff[ps] * ldcb);
                                    o.x += prev[i].x; o.y += prev[i].y; o.z += prev[i].z; o.w += prev[i].w;
                                    if (bias) {
                                        const float4 bv = roff[ps] >= crop_boundary ? b1 : b0;
                                        o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                                    }
                                    o.x += resv[i].x; o.y += resv[i].y; o.z += resv[i].z; o.w += resv[i].w;
                                    if (p.relu == 1) {
                                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                                    } else if (p.relu == 2) {
                                        o.x = o.x > 0.f ? o.x : slope * o.x; o.y = o.y > 0.f ? o.y : slope * o.y;
                                        o.z = o.z > 0.f ? o.z : slope * o.z; o.w = o.w > 0.f ? o.w : slope * o.w;
                                    }
                                    *reinterpret_cast<float4*>(dst) = o;
                                }
                            }
                        }
                    }
                    if (ew == 0) DF_TRACE(16, ti * 4 + (ch >> 1));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (ew == 0) DF_TRACE(11, ti);
            if (lane == 0) {
                if (CTAS == 2) mbar_arrive_remote(acc_empty + ab, 0); else mbar_arrive(acc_empty + ab);
            }
            if (SPLIT && tail_part && kc == rb - 1) {          // this cluster's runs of the shared tile are in C: release the neighbour
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    unsigned int* flag = sched.flags + ((size_t)(cid - 1) * 2 + rank) * 8 + ew;
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flag), "r"(1u) : "memory");
                }
            }
            if (pool_partial) {
                asm volatile("bar.sync 1, 256;" ::: "memory");          // the 8 epilogue warps
                const int tt = threadIdx.x - 320;
                if (tt < bnt && c.n0 + tt < p.N && c.rows_valid > 0) {
                    const float sum = ((pool[tt] + pool[256 + tt]) + pool[512 + tt]) + pool[768 + tt];
                    pool_partial[((size_t)c.crop * p.tiles_per_crop + c.pool_tile) * p.N + c.n0 + tt] = sum;
                }
            }
          }
        }
        if (FORM == 0 && lane == 0) bulk_wait_group0();                 // TMA stores of this warp are complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();              // nobody leaves while the peer may still read / signal this CTA
    if (warp == 1) {
        tc_fence_after();
        if (CTAS == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 2-D fp32 tensor (rows, K) with row pitch `ld` floats; box = 32 floats x box_rows, 128B swizzle
bool make_map(CUtensorMap* map, const float* base, long long rows, int K, int ld, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool use_pdl();

// hybrid mode: per weight row and 32-wide k-block 64 bf16 = [bf16(w) x32 | bf16(w - tf32_trunc(w)) x32]; row pitch 2K bf16
bool make_map_bf16_pairs(CUtensorMap* map, const void* base, long long rows, int K, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)2 * K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// hybrid16 mode, second operand: K bf16 per weight row (bf16(w - fp16(w))); box = 32 bf16 (64 B) x box_rows, 64B swizzle
bool make_map_bf16_rows64(CUtensorMap* map, const void* base, long long rows, int K, int box_rows)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// NHWC image (B, H, W, C) with pixel pitch `ld` floats as a 4-D tensor; box = 32 channels x TW x TH x TB pixels
bool make_map_nhwc(CUtensorMap* map, const float* base, int B, int H, int W, int C, int ld, int TW, int TH, int TB)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tile width of the CTA-pair kernel.  Grouped layers must not straddle groups; otherwise a tail tile is masked.  Widths up to
// `acc_stride` keep two accumulators in TMEM (stores overlap the next tile); a wider one (256 with A in TMEM) only when nothing is
// stored (pooled epilogue) or the k loop is long.  Among the admissible widths take the one with the least work after wave
// quantisation (+32: fixed cost per tile).  Returns 0 when no width fits.
long long split_busiest_kblocks(const TcParams& p, int groups, int total, int max_clusters);

int pair_tile_width(const TcParams& p, int groups, int m_tiles, int acc_stride, int max_clusters)
{
    const int widths[4] = {256, 192, 128, 64};
    long long best = -1;
    int width = 0;
    for (int i = 0; i < 4; ++i) {
        const int w = widths[i];
        // 256-wide tiles leave room for ONE accumulator next to TMEM A stages: always fine for the store-free pooled epilogue; with
        // stores the epilogue of a tile is exposed, which only pays once the k loop is long (>= 32 k-blocks: layer4 convolutions
        // -15..17%, pose step +2.7%, profiles/r2_c4_ab_wide.jsonl; DF_TC_WIDE_KB overrides the threshold, 0 = never)
        static const int wide_kb = getenv("DF_TC_WIDE_KB") ? atoi(getenv("DF_TC_WIDE_KB")) : 32;
        if (w == 256 && acc_stride < 256 && !p.pool_partial && !(wide_kb > 0 && p.K / BK >= wide_kb)) continue;
        if (w == 192 && acc_stride < 192) continue;
        if (w == 64 && p.N > 64) continue;                     // narrow layers only (64-channel decoder stages)
        // 256-wide tiles over an N that is not a multiple of 256 (tower-1: 1920 = 7.5 tiles): the last tile of a row is half masked --
        // its missing weight rows are out-of-bounds rows of the TMA box (zero fill, no bytes through the port), its missing columns are
        // skipped by the epilogue.  Only with two 256-column accumulators (A planes in shared memory), a plain GEMM and a tail of at
        // least half a tile; DF_TC_TAIL256 = 0 never, 1 by the cost below, 2 always.
        static const int tail256 = getenv("DF_TC_TAIL256") ? atoi(getenv("DF_TC_TAIL256")) : 0;
        const bool tail_ok = w == 256 && tail256 && groups == 1 && acc_stride >= 256 && !p.conv_taps && !p.pool_partial && !p.wk_rows &&
                             !p.c1_H && p.N >= 1024 && p.N % 256 >= 128;
        if (p.N % w != 0 && (groups > 1 || (w == 256 && !tail_ok))) continue;
        const long long tiles = (long long)m_tiles * ((p.N + w - 1) / w) * groups;
        long long cost = ((tiles + max_clusters - 1) / max_clusters) * (w + 32);
        if (tail_ok && p.N % w != 0) {
            // operand-port time of a row of tiles (128 activation rows + w / 2 weight rows per CTA and k-block; the masked half of the
            // tail tile is free), in the units of the cost above: 192-wide tiles come out at 224
            const long long nt = (p.N + w - 1) / w;
            cost = tail256 >= 2 ? 0 : ((tiles + max_clusters - 1) / max_clusters) * ((nt * 256 - 64) / nt);
        }
        if (w == 256 && acc_stride >= 256) {
            // long-K convolutions on the balanced schedule (QSched) are not quantised to whole rounds: the busiest cluster's k-blocks
            const long long kb = split_busiest_kblocks(p, groups, (int)tiles, max_clusters);
            if (kb > 0) cost = (kb * (w + 32) + p.K / BK - 1) / (p.K / BK);
        }
        if (best < 0 || cost < best) { best = cost; width = w; }
    }
    return width;
}

int pair_m_tiles(const TcParams& p)
{
    return p.conv_taps ? (p.tiles_x * p.tiles_y * p.tiles_b + 1) / 2
           : p.pool_partial ? (p.M / p.rows_per_crop) * ((p.rows_per_crop + 255) / 256) : (p.M + 255) / 256;
}

unsigned long long* g_trace = nullptr;

typedef void (*QKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams, const int, const int, const int, const int,
                        const QSched, const CUtensorMap);

// Instantiation that serves a launch: the specialised forms (QForm) exist for the hybrid16s variants, i.e. the inference path; everything
// else -- the older arithmetic modes, the weight-gradient form, multi-run GEMMs, PReLU on a GEMM, DF_TC_DBG -- runs the generic one.
// Accumulation runs of a launch (TcParams::k_chunks / kbc): run length in MMA instructions, p.run_steps or DF_TC_RUN_STEPS (216)
void plan_runs(TcParams& p)
{
    static const int env_steps = getenv("DF_TC_RUN_STEPS") ? atoi(getenv("DF_TC_RUN_STEPS")) : 216;
    const int run_steps = p.run_steps > 0 ? p.run_steps : env_steps;
    const int per_kb = p.precise == 1 ? 12 : (p.precise == 2 ? 8 : 6);
    const int run_kb = run_steps / per_kb > 0 ? run_steps / per_kb : 1;
    const int nkb = p.K / BK;
    p.k_chunks = 1; p.kbc = nkb;
    if (p.precise && !p.pool_partial && nkb > run_kb + run_kb / 3) {
        p.k_chunks = (nkb + run_kb - 1) / run_kb;
        p.kbc = (nkb + p.k_chunks - 1) / p.k_chunks;
        p.k_chunks = (nkb + p.kbc - 1) / p.kbc;
    }
}

// Balanced schedule (QSched) of a long-K convolution launch: contiguous ranges of equal k-block weight, cut at accumulation-run
// boundaries.  Returns false -- the launch keeps the round-robin walk -- when the form does not apply, a tile would be shared by more
// than two clusters, or the busiest cluster would not get at least 10% less to do than under round robin (measured, profiles/r2_s4_split_ab.jsonl: a predicted 8% at 200 tiles of 144 k-blocks came out 0.5-2% SLOWER, a predicted 12.5% at 126 tiles 8% faster).
__device__ unsigned int g_split_flags[256 * Q_SCHED_MAX * 16];     // 256 regions, one per launch in flight (taken round robin)

bool plan_split_ranges(const TcParams& p, int groups, int n_tiles, int total, int clusters, QSched* out, long long* rr_out, long long* sp_out)
{
    if (!p.conv_taps || p.k_chunks < 2 || p.wk_rows || p.m_fastest || p.dbg || groups != 1) return false;
    if (clusters < 2 || clusters > Q_SCHED_MAX || total < clusters || total >= 0x7fff) return false;
    const int cblocks = p.K / (BK * p.conv_taps);
    std::vector<int> nkb((size_t)total);
    long long sum = 0;
    for (int t = 0; t < total; ++t) {
        const int mt = t / n_tiles;
        uint32_t mask = 0;
        for (int r = 0; r < 2; ++r) {
            int tx, ty, tb;
            q_patch(p, mt * 2 + r, tx, ty, tb);
            if (ty < p.tiles_y && tb < p.tiles_b) mask |= q_tap_mask(p, tx * p.TW, ty * p.TH);
        }
        nkb[t] = __builtin_popcount(mask) * cblocks;
        if (nkb[t] <= 0) return false;
        sum += nkb[t];
    }
    // round robin: the busiest cluster
    long long rr_max = 0;
    for (int c = 0; c < clusters; ++c) {
        long long load = 0;
        for (int t = c; t < total; t += clusters) load += nkb[t];
        rr_max = load > rr_max ? load : rr_max;
    }
    // boundaries b_c = sum * c / clusters, moved to the nearest run boundary of the tile they fall into
    std::vector<int> bt((size_t)clusters + 1), br((size_t)clusters + 1);
    bt[0] = 0; br[0] = 0; bt[clusters] = total; br[clusters] = 0;
    {
        int t = 0;
        long long cum = 0;                                             // k-blocks before tile t
        for (int c = 1; c < clusters; ++c) {
            const long long target = sum * c / clusters;
            while (t < total && cum + nkb[t] <= target) cum += nkb[t++];
            if (t >= total) return false;
            const int runs = (nkb[t] + p.kbc - 1) / p.kbc;
            int r = (int)((target - cum + p.kbc / 2) / p.kbc);
            if (r >= runs) { bt[c] = t + 1; br[c] = 0; } else { bt[c] = t; br[c] = r; }
        }
    }
    long long sp_max = 0;
    for (int c = 0; c < clusters; ++c) {
        const int t0 = bt[c], r0 = br[c];
        int t1 = bt[c + 1], r1 = br[c + 1];
        if (r1 == 0) { t1 -= 1; r1 = 0x7fff; }
        if (t1 < t0 || (t1 == t0 && r0 > 0 && r1 != 0x7fff)) return false;      // empty range / a tile cut twice
        out->t0[c] = (short)t0; out->r0[c] = (short)r0; out->t1[c] = (short)t1; out->r1[c] = (short)r1;
        long long load = 0;
        for (int t = t0; t <= t1; ++t) {
            const int j0 = t == t0 ? r0 * p.kbc : 0;
            const int j1 = (t == t1 && r1 != 0x7fff) ? (r1 * p.kbc < nkb[t] ? r1 * p.kbc : nkb[t]) : nkb[t];
            if (j1 <= j0) return false;
            load += j1 - j0;
        }
        sp_max = load > sp_max ? load : sp_max;
    }
    if (rr_out) *rr_out = rr_max;
    if (sp_out) *sp_out = sp_max;
    return sp_max * 100 <= rr_max * 90;
}

static int split_enabled()
{
    static const int enabled = getenv("DF_TC_SPLIT") ? atoi(getenv("DF_TC_SPLIT")) : 1;
    return enabled;
}

// k-blocks of the busiest cluster when a launch on 256-wide tiles takes the balanced schedule, 0 when it would not (pair_tile_width)
long long split_busiest_kblocks(const TcParams& p_in, int groups, int total, int max_clusters)
{
    if (!split_enabled() || !p_in.conv_taps || p_in.precise != 4 || p_in.N % 256) return 0;
    TcParams p = p_in;
    if (!p.k_chunks) plan_runs(p);
    QSched tmp;
    long long rr = 0, sp = 0;
    const int clusters = total < max_clusters ? total : max_clusters;
    return plan_split_ranges(p, groups, p.N / 256, total, clusters, &tmp, &rr, &sp) ? sp : 0;
}

bool plan_split(const TcParams& p, int groups, int n_tiles, int total, int clusters, QSched* out)
{
    if (!split_enabled() || !plan_split_ranges(p, groups, n_tiles, total, clusters, out, nullptr, nullptr)) return false;
    static unsigned int* flags = nullptr;
    if (!flags && cudaGetSymbolAddress(reinterpret_cast<void**>(&flags), g_split_flags) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    static unsigned int region = 0;                                   // (atomic: launches may come from several host threads)
    out->flags = flags + (size_t)(__sync_fetch_and_add(&region, 1u) & 255u) * (Q_SCHED_MAX * 16);
    return true;
}

template <int CTAS, int A_STAGES, int A_COLS>
QKernel pick_q_kernel(const TcParams& p, int* form_out, bool split)
{
    if constexpr (A_COLS == 0 && A_STAGES == 4 && CTAS == 2) {
        if (split) { *form_out = 4; return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 4>; }
    }
    int form = -1;
    if (A_COLS != 64 && !p.dbg && !p.wk_rows && !p.m_fastest) {
        if (p.c1_H) form = A_COLS == 32 ? 3 : -1;
        else if (p.pool_partial) form = (A_COLS == 0 && A_STAGES == 3) ? 2 : -1;
        else if (p.conv_taps) form = (A_STAGES == 4) ? 1 : -1;
        else if (p.k_chunks == 1 && !p.residual && p.relu != 2 && A_STAGES == 4) form = 0;
    }
    *form_out = form;
    if (p.dbg) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, true, -1>;
    if constexpr (A_COLS == 32) {
        if (form == 0) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 0>;
        if (form == 1) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 1>;
        if (form == 3) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 3>;
    }
    if constexpr (A_COLS == 0 && A_STAGES == 4) {
        if (form == 0) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 0>;
        if (form == 1) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 1>;
    }
    if constexpr (A_COLS == 0 && A_STAGES == 3) {
        if (form == 2) return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, 2>;
    }
    *form_out = -1;
    return gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, -1>;
}

template <int CTAS, int A_STAGES, int A_COLS = 64>
int launch_q(const TcParams& p_in, const float* W_hi, const float* W_lo, int ldw, int groups, cudaStream_t s)
{
    static_assert(A_COLS == 64 || ((A_COLS == 32 || A_COLS == 0) && CTAS == 2), "the 32-column / shared-memory A stages are CTA-pair hybrid16s forms");
    constexpr int THREADS = Q_THREADS;
    constexpr int ACC_STRIDE = (512 - A_STAGES * A_COLS) / 2;
    if ((A_COLS != 64) != (p_in.precise == 4)) return DF_ERR_ARG;
    TcParams p = p_in;
    {
        static int order = -1;
        if (order < 0) { const char* e = getenv("DF_TC_TILE_ORDER"); order = e ? atoi(e) : 0; }
        p.m_fastest = order;
    }
    {   // accumulation runs once the chain is long enough to matter (see TcParams::k_chunks).  What grows the truncation bias
        // is the number of MMA instructions chained on one accumulator, so the run length is set in instructions: <= 216
        // (18 k-blocks of 3xTF32 at 12 per k-block, 27 of hybrid at 8, 36 of hybrid16 at 6).
        // Inference default 216; the training path asks for 108 (precision bits 8..15): the truncation is a BIAS, which the
        // sums over pixels of a weight gradient amplify (layer4.1.conv1 gradient error 9e-4 at 108, 5.9e-3 at 216, measured).
        {   // DF_TC_BIAS_COMP="h16,hybrid,3xtf32,h16s" (per-instruction factors; calibration knob)
            static float comp[4] = {0.f, 0.f, 0.f, 0.f};
            static bool parsed = false;
            if (!parsed) {
                const char* e = getenv("DF_TC_BIAS_COMP");
                comp[0] = BIAS_COMP_H16; comp[1] = BIAS_COMP_HYBRID; comp[2] = BIAS_COMP_3XTF32; comp[3] = BIAS_COMP_H16S;
                if (e) sscanf(e, "%f,%f,%f,%f", &comp[0], &comp[1], &comp[2], &comp[3]);
                parsed = true;
            }
            p.bias_comp = p.precise == 4 ? comp[3] : p.precise == 3 ? comp[0] : (p.precise == 2 ? comp[1] : (p.precise == 1 ? comp[2] : 0.0f));
        }
        static const int dbg = getenv("DF_TC_DBG") ? atoi(getenv("DF_TC_DBG")) : 0;
        p.dbg = dbg;
        p.trace = nullptr;
        if (dbg & 256) {
            if (!g_trace && cudaMalloc(&g_trace, TRACE_EV * TRACE_KB * sizeof(unsigned long long)) != cudaSuccess) return DF_ERR_UNSUPPORTED;
            cudaMemsetAsync(g_trace, 0, TRACE_EV * TRACE_KB * sizeof(unsigned long long), s);
            p.trace = g_trace;
        }
        plan_runs(p);
    }
    static int max_clusters = 0;
    if (!max_clusters) {
        int dev = 0, num_sms = 0;
        cudaGetDevice(&dev);
        cudaError_t e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_TOTAL);
        if (e != cudaSuccess) return (int)e;
        int n = num_sms / CTAS;
        if (CTAS == 2) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3(num_sms / 2 * 2); q.blockDim = dim3(THREADS); q.dynamicSmemBytes = Q_SMEM_TOTAL;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.attrs = at; q.numAttrs = 1;
            int occ = 0;
            if (cudaOccupancyMaxActiveClusters(&occ, gemm_tc_q_kernel<CTAS, A_STAGES, A_COLS, false, -1>, &q) == cudaSuccess && occ > 0 && occ < n) n = occ;
            (void)cudaGetLastError();
        }
        max_clusters = n;
    }
    int bn_cta = 0;
    const int rows_per_mtile = 128 * CTAS;
    const int m_tiles = p.conv_taps ? (p.tiles_x * p.tiles_y * p.tiles_b + CTAS - 1) / CTAS
                        : p.pool_partial ? (p.M / p.rows_per_crop) * ((p.rows_per_crop + rows_per_mtile - 1) / rows_per_mtile)
                                         : (p.M + rows_per_mtile - 1) / rows_per_mtile;
    if (CTAS == 1) {
        bn_cta = 128;
        if (groups > 1 && p.N % 128) return DF_ERR_UNSUPPORTED;
    } else if (p.wk_rows) {
        bn_cta = 64;                                               // every CTA's half tile stays inside one tap (wk_rows % 64 == 0)
    } else {
        bn_cta = pair_tile_width(p, groups, m_tiles, ACC_STRIDE, max_clusters) / 2;
        if (!bn_cta) return DF_ERR_UNSUPPORTED;
    }
    const int bnt = bn_cta * CTAS;
    // the A operand as a 2-D tensor: group g's columns start at g*a_gs inside the row
    const long long a_cols = (long long)(groups - 1) * p.a_gs + p.K;
    if (!p.conv_taps && !p.c1_H && (a_cols > p.lda || (groups > 1 && p.a_gs < 0))) return DF_ERR_UNSUPPORTED;
    if (p.c1_H && (A_COLS != 32 || bnt < p.N || groups != 1)) return DF_ERR_UNSUPPORTED;      // conv1 form: one n-tile, A planes in TMEM
    CUtensorMap ma, mhi, mlo;
    if (p.c1_H) {
        // (no A tensor map: the stagers gather the patches themselves; `ma` is set to the weight map below)
    } else if (p.conv_taps) {
        if (!make_map_nhwc(&ma, p.A, p.cB, p.cH, p.cW, p.K / p.conv_taps, p.lda, p.TW, p.TH, p.TB)) return DF_ERR_UNSUPPORTED;
    } else if (!make_map(&ma, p.A, p.M, (int)a_cols, p.lda, 128)) return DF_ERR_UNSUPPORTED;
    const long long wrows = p.wk_rows ? (long long)p.wk_rows * (p.N / p.wk_rows == 9 ? 3 : 1) : (long long)groups * p.N;
    if (p.wk_rows) {                                               // weight-gradient form: W spans every k slice (row length ldw)
        if (!make_map(&mhi, W_hi, wrows, ldw, ldw, bn_cta) || !make_map(&mlo, W_lo, wrows, ldw, ldw, bn_cta)) return DF_ERR_UNSUPPORTED;
    } else if (p.precise == 4) {                                   // hybrid16s: one packed tensor, [fp16 hi x32 | fp16 lo x32] per row and k-block
        if (ldw != p.K || !p.w_inv_scale) return DF_ERR_UNSUPPORTED;
        if (!make_map_bf16_pairs(&mhi, W_hi, wrows, p.K, bn_cta)) return DF_ERR_UNSUPPORTED;
        mlo = mhi;
    } else if (p.precise == 3) {                                   // hybrid16: both weight operands are packed 16-bit pair tensors
        if (ldw != p.K) return DF_ERR_UNSUPPORTED;
        if (!make_map_bf16_pairs(&mhi, W_hi, wrows, p.K, bn_cta) || !make_map_bf16_rows64(&mlo, W_lo, wrows, p.K, bn_cta))
            return DF_ERR_UNSUPPORTED;
    } else {
        if (!make_map(&mhi, W_hi, wrows, p.K, ldw, bn_cta)) return DF_ERR_UNSUPPORTED;
        if (p.precise == 2) {
            if (ldw != p.K) return DF_ERR_UNSUPPORTED;             // the packed pair tensor has the dense row pitch
            if (!make_map_bf16_pairs(&mlo, W_lo, wrows, p.K, bn_cta)) return DF_ERR_UNSUPPORTED;
        } else if (!make_map(&mlo, p.precise ? W_lo : W_hi, wrows, p.K, ldw, bn_cta)) return DF_ERR_UNSUPPORTED;
    }
    p.samp_rows = p.conv_taps ? (long long)p.cB * p.cH * p.cW : p.M;
    p.samp_cols = p.conv_taps ? p.K / p.conv_taps : (int)a_cols;
    if (p.c1_H) {                                                  // the image as a (B*3*H, W) matrix
        p.samp_rows = (long long)(p.M / (p.c1_Ho * p.c1_Wo)) * 3 * p.c1_H;
        p.samp_cols = p.c1_W;
        ma = mhi;
    }
    const int n_tiles = (p.N + bnt - 1) / bnt;
    const int total = m_tiles * n_tiles * groups;
    const int clusters = total < max_clusters ? total : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CTAS); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = Q_SMEM_TOTAL; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = use_pdl() ? 2 : 1;
    int form = -1;
    CUtensorMap mc = mhi;
    {   // TMA-store epilogue of the plain GEMM form: C as a 2-D tensor whose boxes are the epilogue warps' 32 x 32 chunks; grouped
        // launches must write column slices of one matrix (c_gs = column offset)
        static const int ts_env = getenv("DF_TC_TSTORE") ? atoi(getenv("DF_TC_TSTORE")) : 1;
        const long long width = groups == 1 ? p.N : (long long)(groups - 1) * p.c_gs + p.N;
        p.tstore = ts_env && A_COLS != 64 && !p.conv_taps && !p.pool_partial && !p.c1_H && !p.wk_rows && p.k_chunks == 1 && !p.residual &&
                   p.relu != 2 && (groups == 1 || (p.c_gs >= p.N && p.N % bnt == 0)) && width <= p.ldc && !(p.ldc & 3) && !((uintptr_t)p.C & 15) &&
                   make_map(&mc, p.C, p.M, (int)width, p.ldc, 32);
    }
    QSched sched = {};
    const bool split = (CTAS == 2 && A_COLS == 0 && A_STAGES == 4) ? plan_split(p, groups, n_tiles, total, clusters, &sched) : false;
    const QKernel kern = pick_q_kernel<CTAS, A_STAGES, A_COLS>(p, &form, split);
    {   // dynamic shared memory opt-in, once per instantiation (index: form + 1, debug build last)
        static bool smem_set[7] = {false, false, false, false, false, false, false};
        const int slot = p.dbg ? 6 : form + 1;
        if (!smem_set[slot]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_TOTAL);
            if (e != cudaSuccess) return (int)e;
            smem_set[slot] = true;
        }
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ma, mhi, mlo, p, bn_cta, m_tiles, n_tiles, total, sched, mc);
    return e == cudaSuccess ? 0 : (int)e;
}

__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[i] = h;
    lo[i] = v - h;
}

// hybrid16s: which launches keep the A planes in shared memory (two 256-column accumulators) instead of TMEM (four A stages, accumulators
// of up to 192 columns double-buffered, 256 single).  DF_TC_A_SMEM = 0 never, 1 always, unset: the shapes that run on 256-wide tiles.
bool s16_a_in_smem(const TcParams& p, int groups)
{
    static const int mode = getenv("DF_TC_A_SMEM") ? atoi(getenv("DF_TC_A_SMEM")) : 2;
    if (mode != 2) return mode == 1;
    // measured per shape (profiles/r2j_gemm_ab_asmem*.jsonl): the shared-memory A ring wins where it buys the second 256-column
    // accumulator (conv6 pooled -13%, conv5 -14%, tower-2 -10%, layer4 -6%) and loses 6-8% on narrower tiles (more shared-memory traffic)
    static int clusters = 0;
    if (!clusters) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        clusters = sms / 2;
    }
    return !p.wk_rows && pair_tile_width(p, groups, pair_m_tiles(p), 256, clusters) == 256;
}

// DF_TC_PDL=0 turns programmatic dependent launch off (A/B timing runs)
bool use_pdl()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DF_TC_PDL");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v == 1;
}

// DF_TC_VARIANT=4..7 overrides the automatic choice (A/B timing runs)
int default_variant()
{
    static int v = 0;
    if (!v) {
        const char* e = getenv("DF_TC_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 5 || v > 7) v = 6;
    }
    return v;
}

__global__ void pack_bf16_pairs_kernel(const float* __restrict__ w, uint32_t* __restrict__ out, long long rows, int K)
{
    // one thread per (row, k-block, pair j): writes bf16x2 of w[2j], w[2j+1] and of their low parts
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = rows * (K / 2);
    if (i >= total) return;
    const long long row = i / (K / 2);
    const int kp = (int)(i - row * (K / 2));
    const int kb = kp / 16, j = kp - kb * 16;
    const float x0 = w[row * K + kb * 32 + 2 * j], x1 = w[row * K + kb * 32 + 2 * j + 1];
    const float l0 = x0 - __uint_as_float(__float_as_uint(x0) & 0xffffe000u);
    const float l1 = x1 - __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
    uint32_t* o = out + row * K + kb * 32;                       // 64 bf16 = 32 words per k-block
    o[j] = pack_bf16x2(x0, x1);
    o[16 + j] = pack_bf16x2(l0, l1);
}

// hybrid16 mode: t1 = per row and k-block [fp16(w) x32 | bf16(w) x32], t2 = bf16(w - fp16(w)) row-major (rows, K); fp16 saturates
// at +-65504 and the remainder (as for values below fp16's normal range) moves into the correction term
__global__ void pack_f16_pairs_kernel(const float* __restrict__ w, uint32_t* __restrict__ t1, uint32_t* __restrict__ t2,
                                      long long rows, int K)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = rows * (K / 2);
    if (i >= total) return;
    const long long row = i / (K / 2);
    const int kp = (int)(i - row * (K / 2));
    const int kb = kp / 16, j = kp - kb * 16;
    const float x0 = w[row * K + kb * 32 + 2 * j], x1 = w[row * K + kb * 32 + 2 * j + 1];
    const uint32_t h = pack_f16x2_sat(x0, x1);
    float f0, f1;
    unpack_f16x2(h, f0, f1);
    t1[row * K + kb * 32 + j] = h;
    t1[row * K + kb * 32 + 16 + j] = pack_bf16x2(x0, x1);
    t2[i] = pack_bf16x2(x0 - f0, x1 - f1);                       // plain row-major bf16 (rows, K)
}

// hybrid16s mode: absolute maximum of the weight tensor (bit pattern of a non-negative float orders like an unsigned integer) ...
__global__ void absmax_kernel(const float* __restrict__ w, long long n, uint32_t* __restrict__ out)
{
    uint32_t m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t b = __float_as_uint(w[i]) & 0x7fffffffu;
        if (b <= 0x7f800000u && b > m) m = b;                      // (NaN is skipped: it poisons its own products only)
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// ... and the planes: with s = the power of two that brings that maximum into [2^14, 2^15), per row and k-block
// [fp16(w s) x32 | fp16(w s - fp16(w s)) x32].  The remainder of an entry is ~2^-12 of it, so every entry within 2^-16 of the largest
// keeps a NORMAL fp16 remainder (22 significant bits in all); smaller ones lose bits that are below 2^-39 of the largest entry.
// scale[0] = 1 / s (what the GEMM epilogue multiplies by), scale[1] = s; scale[2] holds the maximum's bit pattern.
__global__ void pack_f16s_kernel(const float* __restrict__ w, uint32_t* __restrict__ planes, float* __restrict__ scale,
                                 long long rows, int K)
{
    const uint32_t mb = reinterpret_cast<const uint32_t*>(scale)[2];
    int e = (int)(mb >> 23) - 127;                                  // floor(log2(max)) (0 / subnormal maximum: -127)
    if (mb == 0x7f800000u || e < -100) e = 14;                      // infinite or (sub)zero tensor: s = 1
    const float s = exp2f((float)(14 - e));
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { scale[0] = exp2f((float)(e - 14)); scale[1] = s; }
    const long long total = rows * (K / 2);
    if (i >= total) return;
    const long long row = i / (K / 2);
    const int kp = (int)(i - row * (K / 2));
    const int kb = kp / 16, j = kp - kb * 16;
    const float x0 = w[row * K + kb * 32 + 2 * j] * s, x1 = w[row * K + kb * 32 + 2 * j + 1] * s;
    const uint32_t h = pack_f16x2_sat(x0, x1);
    float f0, f1;
    unpack_f16x2(h, f0, f1);
    planes[row * K + kb * 32 + j] = h;
    planes[row * K + kb * 32 + 16 + j] = pack_f16x2_sat(x0 - f0, x1 - f1);
}

// Convolution weight (Cout, Cin, taps) -> GEMM operand (rows, taps*cols) tap-major, split for the tensor-core modes, in ONE
// pass (training repacks every step).  rotate == 0: rows = Cout, cols = Cin (forward).  rotate == 1: rows = Cin, cols = Cout,
// taps reversed -- the 180-degree rotated, in/out-transposed kernel of the data gradient.  One thread per two k.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo,
                                        uint32_t* __restrict__ pairs, int Cout, int Cin, int taps, int rotate)
{
    const int rows = rotate ? Cin : Cout, cols = rotate ? Cout : Cin;
    const int K = taps * cols;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * (K / 2)) return;
    const int row = (int)(i / (K / 2)), kp = (int)(i - (long long)row * (K / 2));
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int k = kp * 2 + e, tap = k / cols, c = k - tap * cols;
        x[e] = rotate ? w[((size_t)c * Cin + row) * taps + (taps - 1 - tap)] : w[((size_t)row * Cin + c) * taps + tap];
    }
    const float h0 = __uint_as_float(__float_as_uint(x[0]) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x[1]) & 0xffffe000u);
    float* hrow = hi + (size_t)row * K + kp * 2;
    const bool raw = !lo && !pairs;                              // plain fp32 repack (input of df_pack_f16s)
    hrow[0] = raw ? x[0] : h0; hrow[1] = raw ? x[1] : h1;
    if (lo) { float* lrow = lo + (size_t)row * K + kp * 2; lrow[0] = x[0] - h0; lrow[1] = x[1] - h1; }
    if (pairs) {
        const int kb = (kp * 2) / 32, j = kp - kb * 16;
        uint32_t* o = pairs + (size_t)row * K + kb * 32;
        o[j] = pack_bf16x2(x[0], x[1]);
        o[16 + j] = pack_bf16x2(x[0] - h0, x[1] - h1);
    }
}

// the same repacking for the hybrid16 mode: t1 = [fp16(w) x32 | bf16(w) x32] per k-block, t2 = bf16(w - fp16(w)) row-major
__global__ void pack_conv_weight16_kernel(const float* __restrict__ w, uint32_t* __restrict__ t1, uint32_t* __restrict__ t2,
                                          int Cout, int Cin, int taps, int rotate)
{
    const int rows = rotate ? Cin : Cout, cols = rotate ? Cout : Cin;
    const int K = taps * cols;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * (K / 2)) return;
    const int row = (int)(i / (K / 2)), kp = (int)(i - (long long)row * (K / 2));
    float x[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int k = kp * 2 + e, tap = k / cols, c = k - tap * cols;
        x[e] = rotate ? w[((size_t)c * Cin + row) * taps + (taps - 1 - tap)] : w[((size_t)row * Cin + c) * taps + tap];
    }
    const uint32_t h = pack_f16x2_sat(x[0], x[1]);
    float f0, f1;
    unpack_f16x2(h, f0, f1);
    const int kb = (kp * 2) / 32, j = kp - kb * 16;
    t1[(size_t)row * K + kb * 32 + j] = h;
    t1[(size_t)row * K + kb * 32 + 16 + j] = pack_bf16x2(x[0], x[1]);
    t2[i] = pack_bf16x2(x[0] - f0, x[1] - f1);
}

// NHWC activation (B,H,W,C; pixel pitch ld) -> channel-major rows over the zero-padded, flattened pixel axis:
// out[c][k] = padded input at flat position k, where position (b*(H+2d) + y+d)*Wp + x+d holds in[b,y,x,c] (Wp >= W+2d) and
// everything else up to the row length Ppad is zero.  `copies` == 3 writes three planes (plane stride `plane`), plane kx read
// (kx-1)*d positions further -- the horizontal tap offsets of a 3x3 weight gradient.  With `lo` every value is split for
// 3xTF32 (out = TF32-exact part, lo = remainder).  32-pixel x 32-channel tiles (+ d pixels of halo on both sides) through
// shared memory: reads run along the channels, writes along the pixels; the input is read once for all planes.
__global__ void __launch_bounds__(256)
cmajor_pad_kernel(const float* __restrict__ in, int ld, float* __restrict__ out, float* __restrict__ lo, int B, int H, int W,
                  int C, int d, int Wp, int Ppad, int copies, long long plane)
{
    __shared__ float tile[40][33];                                 // 32 + 2 * 4 rows: dilation <= 4
    __shared__ int src[40];                                        // source pixel of every tile row, -1 = padding (decoded once per block)
    const int Hp = H + 2 * d;
    const int halo = copies == 3 ? d : 0;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (threadIdx.x < 32 + 2 * halo) {
        const int pp = p0 - halo + (int)threadIdx.x;
        int q = -1;
        if (pp >= 0 && pp < B * Hp * Wp) {
            const int x = pp % Wp - d, r = pp / Wp, y = r % Hp - d, b = r / Hp;
            if (x >= 0 && x < W && y >= 0 && y < H) q = (b * H + y) * W + x;
        }
        src[threadIdx.x] = q;
    }
    __syncthreads();
    for (int j = ty; j < 32 + 2 * halo; j += 8) {
        const int q = src[j], c = c0 + tx;
        tile[j][tx] = (q >= 0 && c < C) ? __ldg(in + (size_t)q * ld + c) : 0.0f;
    }
    __syncthreads();
    if (p0 + tx >= Ppad) return;
    for (int kx = 0; kx < copies; ++kx) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty + i * 8;
            if (c >= C) continue;
            const float v = tile[tx + kx * halo][ty + i * 8];
            const size_t o = (size_t)kx * plane + (size_t)c * Ppad + p0 + tx;
            if (lo) {
                const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
                out[o] = h;
                lo[o] = v - h;
            } else {
                out[o] = v;
            }
        }
    }
}

}  // namespace

// Weight gradient of a stride-1 3x3 (padding == dilation) or 1x1 convolution on the tensor cores:
//   dW[co, tap*Cin + ci] = sum_p dY[p, co] * X[p + offset(tap), ci]
// as ONE 3xTF32 GEMM whose reduction runs over the pixels: both operands are first rewritten channel-major over the
// zero-padded, flattened pixel axis (cmajor_pad_kernel), where a tap is a plain offset along k -- the W-operand TMA box of
// output-column block (tap, ci..) simply starts wk_shift[tap] elements further (out-of-range parts are zero-filled), and the
// zero borders of dY make every product that would wrap into a neighbouring row or image vanish.
// TMA boxes must start 16-byte aligned, so only offsets that are multiples of 4 elements can be taken at load time: the padded
// row length Wp is rounded up to a multiple of 4 (vertical tap offsets (ky-1)*d*Wp), and the horizontal tap offset (kx-1)*d is
// baked into three pre-shifted copies of X (rows kx*Cin + ci of the W operand).
// Split-K: the reduction runs over thousands of pixels while the output (Cout x taps*Cin) is a few dozen tiles, so the pixel
// axis is cut into S slices (GEMM groups: A column offset and W k offset g*Kg, partial outputs summed by df_reduce_partials in
// fixed order) until about two waves of CTA pairs are busy.
struct WgradGeom { int d, Wp, Ppad, S, Kg, copies; };

static inline WgradGeom wgrad_geometry(int B, int H, int W, int Cin, int Cout, int taps, int dilation)
{
    WgradGeom g;
    g.d = taps == 9 ? dilation : 0;
    g.copies = taps == 9 ? 3 : 1;
    g.Wp = taps == 9 ? (W + 2 * g.d + 3) / 4 * 4 : W;
    const long long P = (long long)B * (H + 2 * g.d) * g.Wp;
    const int nkb = (int)((P + 31) / 32);
    const long long tiles = (long long)((Cout + 255) / 256) * ((taps * Cin + 127) / 128);
    int S = (int)((148 + tiles - 1) / tiles);
    if (S > nkb / 8) S = nkb / 8;
    if (S < 1) S = 1;
    const int kbs = (nkb + S - 1) / S;                 // k-blocks per slice
    g.S = (nkb + kbs - 1) / kbs;
    g.Kg = kbs * 32;
    g.Ppad = g.S * g.Kg;
    return g;
}

extern "C" long long df_conv_wgrad_scratch_floats(int B, int H, int W, int Cin, int Cout, int taps, int dilation)
{
    const WgradGeom g = wgrad_geometry(B, H, W, Cin, Cout, taps, dilation);
    return (long long)g.Ppad * ((long long)Cout + 2LL * g.copies * Cin) + (g.S > 1 ? (long long)g.S * Cout * taps * Cin : 0);
}

extern "C" int df_conv_wgrad_tc(const float* X, int ldx, const float* dY, int ldy, int B, int H, int W, int Cin, int Cout, int taps,
                                int dilation, float* scratch, float* dW, void* stream)
{
    if (!X || !dY || !scratch || !dW || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return DF_ERR_ARG;
    if ((taps != 1 && taps != 9) || dilation < 1 || Cin % 64 || Cout % 4 || ldx < Cin || ldy < Cout) return DF_ERR_ARG;
    if (((uintptr_t)scratch & 15) || ((uintptr_t)dW & 15)) return DF_ERR_ARG;
    if ((long long)B * (H + 2 * dilation) * (W + 2 * dilation + 3) >= (1LL << 30)) return DF_ERR_ARG;
    const WgradGeom g = wgrad_geometry(B, H, W, Cin, Cout, taps, dilation);
    const int d = g.d, Wp = g.Wp, Ppad = g.Ppad, copies = g.copies;
    float* dyT = scratch;
    float* xhi = dyT + (size_t)Cout * Ppad;
    float* xlo = xhi + (size_t)copies * Cin * Ppad;
    float* part = xlo + (size_t)copies * Cin * Ppad;                       // S partial outputs (S > 1)
    cudaStream_t s = (cudaStream_t)stream;
    if (d > 4) return DF_ERR_UNSUPPORTED;                                  // (halo rows of the transpose tile)
    cmajor_pad_kernel<<<dim3(Ppad / 32, (Cout + 31) / 32), 256, 0, s>>>(dY, ldy, dyT, nullptr, B, H, W, Cout, d, Wp, Ppad, 1, 0);
    cmajor_pad_kernel<<<dim3(Ppad / 32, (Cin + 31) / 32), 256, 0, s>>>(X, ldx, xhi, xlo, B, H, W, Cin, d, Wp, Ppad, copies,
                                                                     (long long)Cin * Ppad);

    const long long out_floats = (long long)Cout * taps * Cin;
    TcParams p = {};
    p.A = dyT; p.lda = Ppad; p.bias = nullptr; p.bias_crop_stride = 0; p.C = g.S > 1 ? part : dW; p.ldc = taps * Cin;
    p.M = Cout; p.N = taps * Cin; p.K = g.Kg; p.relu = 0; p.precise = 1;
    p.rows_per_crop = p.M; p.a_gs = g.Kg; p.bias_gs = 0; p.c_gs = out_floats; p.pool_partial = nullptr; p.tiles_per_crop = 0;
    p.wk_rows = Cin;
    for (int t = 0; t < taps; ++t) {
        p.wk_shift[t] = taps == 9 ? (t / 3 - 1) * d * Wp : 0;
        p.wk_row0[t] = taps == 9 ? (t % 3) * Cin : 0;
    }
    int rc = launch_q<2, 2>(p, xhi, xlo, Ppad, g.S, s);
    if (rc) return rc;
    if (g.S > 1) {
        rc = df_reduce_partials(part, g.S, out_floats, dW, 0, stream);
        if (rc) return rc;
    }
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pack_conv_weight16(const float* w, void* t1, void* t2, int Cout, int Cin, int taps, int rotate, void* stream)
{
    if (!w || !t1 || !t2 || Cout <= 0 || Cin <= 0 || taps <= 0) return DF_ERR_ARG;
    const int cols = rotate ? Cout : Cin, rows = rotate ? Cin : Cout;
    if ((taps * cols) % 32) return DF_ERR_ARG;
    const long long total = (long long)rows * (taps * cols / 2);
    pack_conv_weight16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (uint32_t*)t1, (uint32_t*)t2, Cout,
                                                                                            Cin, taps, rotate);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pack_conv_weight(const float* w, float* hi, float* lo, void* pairs, int Cout, int Cin, int taps, int rotate,
                                   void* stream)
{
    if (!w || !hi || Cout <= 0 || Cin <= 0 || taps <= 0) return DF_ERR_ARG;      // lo == pairs == NULL: hi = the repacked fp32 weights
    const int cols = rotate ? Cout : Cin, rows = rotate ? Cin : Cout;
    if ((taps * cols) % 32) return DF_ERR_ARG;
    const long long total = (long long)rows * (taps * cols / 2);
    pack_conv_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, hi, lo, (uint32_t*)pairs, Cout, Cin,
                                                                                          taps, rotate);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pack_bf16_pairs(const float* w, void* out, long long rows, int K, void* stream)
{
    if (!w || !out || rows <= 0 || K <= 0 || K % 32) return DF_ERR_ARG;
    const long long total = rows * (K / 2);
    pack_bf16_pairs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (uint32_t*)out, rows, K);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_pack_f16_pairs(const float* w, void* t1, void* t2, long long rows, int K, void* stream)
{
    if (!w || !t1 || !t2 || rows <= 0 || K <= 0 || K % 32) return DF_ERR_ARG;
    const long long total = rows * (K / 2);
    pack_f16_pairs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, (uint32_t*)t1, (uint32_t*)t2, rows, K);
    DF_RETURN_LAST_ERROR();
}

// DF_TC_DBG bit 256: the clock64() timeline of the last traced launch, [9 events][96 k-blocks] (see TcParams::trace)
extern "C" int df_tc_trace_read(unsigned long long* host_out, int count)
{
    if (!host_out || count <= 0 || count > TRACE_EV * TRACE_KB) return DF_ERR_ARG;
    if (!g_trace) return DF_ERR_UNSUPPORTED;
    cudaError_t e = cudaMemcpy(host_out, g_trace, (size_t)count * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? 0 : (int)e;
}

extern "C" int df_pack_f16s(const float* w, void* planes, float* scale, long long rows, int K, void* stream)
{
    if (!w || !planes || !scale || rows <= 0 || K <= 0 || K % 32) return DF_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scale, 0, 4 * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
    const long long n = rows * K;
    const int blocks = (int)((n + 1023) / 1024 < 1184 ? (n + 1023) / 1024 : 1184);
    absmax_kernel<<<blocks, 256, 0, s>>>(w, n, reinterpret_cast<uint32_t*>(scale) + 2);
    const long long total = rows * (K / 2);
    pack_f16s_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, (uint32_t*)planes, scale, rows, K);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream)
{
    if (!x || !hi || !lo || n <= 0) return DF_ERR_ARG;
    split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n);
    DF_RETURN_LAST_ERROR();
}

extern "C" int df_gemm_tc(const float* A, int lda, const float* W_hi, const float* W_lo, int ldw, const float* bias,
                          int bias_crop_stride, float* C, int ldc, int M, int N, int K, int relu, int rows_per_crop,
                          int groups, long long a_group_stride, long long bias_group_stride,
                          long long c_group_stride, float* pool_partial, int precision, int variant, void* stream)
{
    if (!A || !W_hi || (!C && !pool_partial)) return DF_ERR_ARG;
    const int run_units = (precision >> 8) & 0xff;                  // accumulation-run length in units of 12 MMA instructions
    const int a_byte = (precision >> 16) & 0xff;                    // hybrid16s: 0 = scale sampled by the kernel, else a fixed 2^k
    precision &= 0xff;
    if (precision < 1 || precision > 6 || precision == 5) return DF_ERR_ARG;       // (5: the on-chip weight split of round 2, removed)
    if (precision != 2 && !W_lo) return DF_ERR_ARG;
    if (M <= 0 || N <= 0 || K <= 0 || groups <= 0) return DF_ERR_ARG;
    if (K % BK || lda % 4 || ldw % 4 || N % 4 || a_group_stride % 4 || bias_group_stride % 4) return DF_ERR_ARG;
    if (((uintptr_t)A & 15) || ((uintptr_t)W_hi & 15) || ((uintptr_t)W_lo & 15)) return DF_ERR_ARG;
    if (C && (ldc % 4 || c_group_stride % 4 || ((uintptr_t)C & 15))) return DF_ERR_ARG;
    if ((bias_crop_stride || pool_partial) && rows_per_crop <= 0) return DF_ERR_ARG;
    if (pool_partial && M % rows_per_crop) return DF_ERR_ARG;
    if (groups > 1 && N % 128) return DF_ERR_UNSUPPORTED;      // group g's weight rows start at g*N: keep tiles inside a group

    TcParams p = {};
    p.A = A; p.lda = lda; p.bias = bias; p.bias_crop_stride = bias_crop_stride;
    p.C = pool_partial ? nullptr : C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.relu = relu;
    // kernel-side: 0 single TF32, 1 3xTF32, 2 hybrid, 3 hybrid16, 4 hybrid16s
    p.precise = precision == 1 ? 1 : (precision == 3 ? 2 : (precision == 6 ? 4 : (precision == 4 ? 3 : 0)));
    p.w_inv_scale = precision == 6 ? W_lo : nullptr; p.a_scale = a_byte == 0 ? 0.0f : (a_byte == 0x80 ? 1.0f : exp2f((float)(signed char)a_byte));
    p.run_steps = run_units * 12;
    p.rows_per_crop = rows_per_crop > 0 ? rows_per_crop : M;
    p.a_gs = a_group_stride; p.bias_gs = bias_group_stride; p.c_gs = c_group_stride;
    p.pool_partial = pool_partial;
    p.tiles_per_crop = pool_partial ? (p.rows_per_crop + BM - 1) / BM : 0;

    // variant: 0 = auto (6); 5 = one CTA per tile; 6 / 7 = CTA pairs (cta_group::2, 256-row tiles) with 2 / 4 TMEM A stages
    // (accumulators up to 192 / 128 columns double-buffered).  5 and 7 are kept for A/B measurements (DF_TC_VARIANT).
    int v = variant;
    if (v == 0) v = default_variant();
    int rc;
    if (precision == 6) {                 // hybrid16s: two fp16 planes per operand, 32-column TMEM A stages: CTA-pair kernel only
        if (variant != 0 && variant != 6) return DF_ERR_UNSUPPORTED;
        rc = s16_a_in_smem(p, groups) ? (p.pool_partial ? launch_q<2, 3, 0>(p, W_hi, W_hi, ldw, groups, (cudaStream_t)stream)     // (pool buffer: 3 + 4 stages)
                                                         : launch_q<2, 4, 0>(p, W_hi, W_hi, ldw, groups, (cudaStream_t)stream))
                                      : launch_q<2, 4, 32>(p, W_hi, W_hi, ldw, groups, (cudaStream_t)stream);
    } else if (v == 5) rc = launch_q<1, 4>(p, W_hi, W_lo, ldw, groups, (cudaStream_t)stream);
    else if (v == 6) rc = launch_q<2, 2>(p, W_hi, W_lo, ldw, groups, (cudaStream_t)stream);
    else if (v == 7) rc = launch_q<2, 4>(p, W_hi, W_lo, ldw, groups, (cudaStream_t)stream);
    else return DF_ERR_ARG;
    if (rc) return rc;                           // DF_ERR_UNSUPPORTED: a shape the TMA boxes cannot address (the caller runs df_gemm_fp32)
    DF_RETURN_LAST_ERROR();
}

// conv1 of the encoder (7x7, stride 2, padding 3, 3 -> Cout channels; lib/extractors.py:82) as a GEMM whose A operand is gathered from
// the NCHW image by the kernel's stagers (TcParams::c1_*): y (B*Ho*Wo, Cout) = act(patches . W^T).  `planes` / `scale`: df_pack_f16s of the
// (Cout, 160) weight matrix [c*49 + ky*7 + kx, zero-padded from 147].  hybrid16s arithmetic only.
extern "C" int df_enc_conv1_tc(const float* img, int B, int H, int W, const void* planes, const float* scale, float* Y, int ldy, int Cout,
                               int relu, void* stream)
{
    if (!img || !planes || !scale || !Y || B <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout > 64 || Cout % 4 || ldy % 4 || ldy < Cout) return DF_ERR_ARG;
    if (W % 4 || ((uintptr_t)img & 15) || ((uintptr_t)Y & 15) || ((uintptr_t)planes & 15)) return DF_ERR_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    if ((long long)B * Ho * Wo >= (1LL << 31) || (long long)B * 3 * H * W >= (1LL << 31)) return DF_ERR_ARG;
    TcParams p = {};
    p.A = img; p.lda = W; p.bias = nullptr; p.bias_crop_stride = 0; p.C = Y; p.ldc = ldy;
    p.M = B * Ho * Wo; p.N = Cout; p.K = 160; p.relu = relu ? 1 : 0; p.precise = 4;
    p.w_inv_scale = scale; p.a_scale = 0.0f;
    p.rows_per_crop = p.M; p.a_gs = 0; p.bias_gs = 0; p.c_gs = 0; p.pool_partial = nullptr; p.tiles_per_crop = 0;
    p.c1_H = H; p.c1_W = W; p.c1_Ho = Ho; p.c1_Wo = Wo;
    const int rc = launch_q<2, 4, 32>(p, reinterpret_cast<const float*>(planes), reinterpret_cast<const float*>(planes), 160, 1, (cudaStream_t)stream);
    if (rc) return rc;
    DF_RETURN_LAST_ERROR();
}

// Pixel patch (TB x TH x TW <= 128 rows) of the implicit-GEMM convolution with the fewest tap visits = patches x the taps each
// can reach (q_tap_mask: small patches of a dilated layer skip the taps that only see padding); ties: wider rows.
static void conv_patch_plan(TcParams& p, int B, int H, int W, int taps, int dilation)
{
    auto reach = [&](int extent, int t) {                  // sum over the patches along one axis of the taps (of 3) they reach
        if (taps != 9) return (extent + t - 1) / t;
        int sum = 0;
        for (int o = 0; o < extent; o += t) {
            const int hi = (o + t < extent ? o + t : extent) - 1;
            for (int k = -1; k <= 1; ++k) sum += (hi + k * dilation >= 0 && o + k * dilation < extent) ? 1 : 0;
        }
        return sum;
    };
    long long best_cost = -1;
    for (int tw = 1; tw <= (W < 128 ? W : 128); ++tw) {
        const int rx = reach(W, tw);
        for (int th = 1; th <= H && tw * th <= 128; ++th) {
            int tb = 128 / (tw * th);
            if (tb > B) tb = B;
            const long long cost = (long long)rx * reach(H, th) * ((B + tb - 1) / tb);
            if (best_cost < 0 || cost < best_cost || (cost == best_cost && tw > p.TW)) {
                best_cost = cost; p.TW = tw; p.TH = th; p.TB = tb;
            }
        }
    }
    p.tiles_x = (W + p.TW - 1) / p.TW; p.tiles_y = (H + p.TH - 1) / p.TH; p.tiles_b = (B + p.TB - 1) / p.TB;
}

// Multiply-adds df_conv_tc executes on real output pixels: sum over the pixel patches of (valid rows) x (taps the patch's CTA
// pair visits) x Cin x Cout -- the dense count minus the skipped all-padding taps.  For the time-weighted roofline of bench.py.
extern "C" long long df_conv_tc_macs(int B, int H, int W, int Cin, int Cout, int taps, int dilation)
{
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (taps != 1 && taps != 9) || dilation < 1) return DF_ERR_ARG;
    TcParams p = {};
    p.conv_taps = taps; p.conv_dil = dilation; p.cW = W; p.cH = H; p.cB = B;
    conv_patch_plan(p, B, H, W, taps, dilation);
    const int patches = p.tiles_x * p.tiles_y * p.tiles_b;
    long long macs = 0;
    for (int mt = 0; mt * 2 < patches; ++mt) {
        uint32_t mask = 0;
        long long rows = 0;
        for (int r = 0; r < 2; ++r) {
            int tx, ty, tb;
            q_patch(p, mt * 2 + r, tx, ty, tb);
            if (ty >= p.tiles_y || tb >= p.tiles_b) continue;
            mask |= q_tap_mask(p, tx * p.TW, ty * p.TH);
            const int w = (tx * p.TW + p.TW < W ? p.TW : W - tx * p.TW), h = (ty * p.TH + p.TH < H ? p.TH : H - ty * p.TH);
            const int b = (tb * p.TB + p.TB < B ? p.TB : B - tb * p.TB);
            rows += (long long)w * h * b;
        }
        macs += rows * __builtin_popcount(mask) * Cin * Cout;
    }
    return macs;
}

// The balanced schedule df_conv_tc would use for a hybrid16s 3x3 convolution of this geometry on `clusters` CTA pairs with 256-wide
// tiles (host-side arithmetic only; tests and profiling scripts).  out[0] = k-blocks of the busiest cluster under round robin, out[1] =
// under the balanced schedule, out[2] = tiles, out[3] = k-blocks per run, out[4] = k-blocks of a full tile, then per cluster
// (first tile, first run, last tile, end run or 0x7fff).  Returns 1 when the balanced schedule is taken, 0 when the launch keeps the
// round-robin walk (out[0..4] still set when a schedule exists), < 0 on bad arguments.
// The tile plan df_gemm_tc takes for a hybrid16s GEMM of this shape on `clusters` CTA pairs -- the launcher's own functions, no launch:
// out[0] tile width (columns of the CTA pair's accumulator), [1] 1 = A planes in a shared-memory ring / 0 = in TMEM, [2] accumulators
// in TMEM (2: the drain of a tile overlaps the next one), [3] tiles, [4] rounds of the persistent kernel, [5] accumulation runs per tile.
extern "C" int df_gemm_tc_plan(int M, int N, int K, int groups, int pooled, int rows_per_crop, int clusters, int* out)
{
    if (M <= 0 || N <= 0 || K <= 0 || K % BK || groups <= 0 || clusters < 1 || !out) return DF_ERR_ARG;
    if (groups > 1 && N % 128) return DF_ERR_UNSUPPORTED;
    if (pooled && (rows_per_crop <= 0 || M % rows_per_crop)) return DF_ERR_ARG;
    static float dummy = 0.0f;
    TcParams p = {};
    p.M = M; p.N = N; p.K = K; p.precise = 4;
    p.rows_per_crop = rows_per_crop > 0 ? rows_per_crop : M;
    p.pool_partial = pooled ? &dummy : nullptr;                          // (only tested for null by the planning functions)
    plan_runs(p);
    const int m_tiles = pair_m_tiles(p);
    const bool a_smem = !p.wk_rows && pair_tile_width(p, groups, m_tiles, 256, clusters) == 256;      // s16_a_in_smem's rule
    const int acc_stride = a_smem ? 256 : (512 - 4 * 32) / 2;
    const int w = pair_tile_width(p, groups, m_tiles, acc_stride, clusters);
    if (!w) return DF_ERR_UNSUPPORTED;
    const long long tiles = (long long)m_tiles * ((N + w - 1) / w) * groups;
    out[0] = w; out[1] = a_smem ? 1 : 0; out[2] = w <= acc_stride ? 2 : 1; out[3] = (int)tiles;
    out[4] = (int)((tiles + clusters - 1) / clusters); out[5] = p.k_chunks;
    return 0;
}

extern "C" int df_conv_tc_schedule(int B, int H, int W, int Cin, int Cout, int dilation, int clusters, int* out)
{
    if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cin % BK || Cout <= 0 || Cout % 256 || dilation < 1 || clusters < 1 || clusters > Q_SCHED_MAX || !out)
        return DF_ERR_ARG;
    TcParams p = {};
    p.conv_taps = 9; p.conv_dil = dilation; p.cW = W; p.cH = H; p.cB = B; p.precise = 4;
    p.M = B * H * W; p.N = Cout; p.K = 9 * Cin;
    conv_patch_plan(p, B, H, W, 9, dilation);
    plan_runs(p);
    const int m_tiles = (p.tiles_x * p.tiles_y * p.tiles_b + 1) / 2, n_tiles = Cout / 256, total = m_tiles * n_tiles;
    const int cl = total < clusters ? total : clusters;
    QSched sc = {};
    long long rr = 0, sp = 0;
    const bool ok = plan_split_ranges(p, 1, n_tiles, total, cl, &sc, &rr, &sp);
    out[0] = (int)rr; out[1] = (int)sp; out[2] = total; out[3] = p.kbc; out[4] = p.K / BK;
    for (int c = 0; c < cl; ++c) { out[5 + 4 * c] = sc.t0[c]; out[6 + 4 * c] = sc.r0[c]; out[7 + 4 * c] = sc.t1[c]; out[8 + 4 * c] = sc.r1[c]; }
    return ok ? 1 : 0;
}

// 3x3 (stride 1, padding == dilation) or 1x1 convolution on an NHWC image as an implicit GEMM on the paired tcgen05
// kernel: out[pixel, n] = act( sum_{tap, c} X[pixel + tap*dil, c] W[n, tap*Cin + c] + bias[n] + residual[pixel, n] ).
extern "C" int df_conv_tc(const float* X, int B, int H, int W, int Cin, int ldx, const float* W_hi, const float* W_lo,
                          int taps, int dilation, const float* bias, const float* residual, int ldr, const float* prelu,
                          int act, float* Y, int ldy, int Cout, int precision, void* stream)
{
    if (!X || !W_hi || !Y || B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0) return DF_ERR_ARG;
    const int run_units = (precision >> 8) & 0xff;
    const int a_byte = (precision >> 16) & 0xff;
    precision &= 0xff;
    if (precision < 1 || precision > 6 || precision == 5) return DF_ERR_ARG;       // (5: the on-chip weight split of round 2, removed)
    if (precision != 2 && !W_lo) return DF_ERR_ARG;
    if ((taps != 1 && taps != 9) || dilation < 1 || act < 0 || act > 2 || (act == 2 && !prelu)) return DF_ERR_ARG;
    if (Cin % BK || Cout % 4 || ldx % 4 || ldy % 4 || ldx < Cin || ldy < Cout || (residual && (ldr % 4 || ldr < Cout))) return DF_ERR_ARG;
    if (((uintptr_t)X & 15) || ((uintptr_t)Y & 15) || ((uintptr_t)W_hi & 15) || ((uintptr_t)W_lo & 15) ||
        ((uintptr_t)residual & 15) || ((uintptr_t)bias & 15))
        return DF_ERR_ARG;
    if ((long long)B * H * W >= (1LL << 31)) return DF_ERR_ARG;

    TcParams p = {};
    p.A = X; p.lda = ldx; p.bias = bias; p.bias_crop_stride = 0; p.C = Y; p.ldc = ldy;
    p.M = B * H * W; p.N = Cout; p.K = taps * Cin; p.relu = act;
    p.precise = precision == 1 ? 1 : (precision == 3 ? 2 : (precision == 6 ? 4 : (precision == 4 ? 3 : 0)));
    p.w_inv_scale = precision == 6 ? W_lo : nullptr; p.a_scale = a_byte == 0 ? 0.0f : (a_byte == 0x80 ? 1.0f : exp2f((float)(signed char)a_byte));
    p.run_steps = run_units * 12;
    p.rows_per_crop = p.M; p.a_gs = 0; p.bias_gs = 0; p.c_gs = 0; p.pool_partial = nullptr; p.tiles_per_crop = 0;
    p.conv_taps = taps; p.conv_dil = dilation; p.cW = W; p.cH = H; p.cB = B;
    p.residual = residual; p.ldr = ldr; p.prelu = prelu;
    conv_patch_plan(p, B, H, W, taps, dilation);
    const int rc = precision == 6 ? (s16_a_in_smem(p, 1) ? launch_q<2, 4, 0>(p, W_hi, W_hi, taps * Cin, 1, (cudaStream_t)stream)
                                                         : launch_q<2, 4, 32>(p, W_hi, W_hi, taps * Cin, 1, (cudaStream_t)stream))
                                  : launch_q<2, 2>(p, W_hi, W_lo, taps * Cin, 1, (cudaStream_t)stream);
    if (rc) return rc;
    DF_RETURN_LAST_ERROR();
}
