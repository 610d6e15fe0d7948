// conv_stream_kernel.cu -- K2a, second formulation: boards -> ReLU(conv2(ReLU(conv1(one-hot planes))))
// as fp16 in fc1's A-operand tile order (neural_network.py:22-23,30-31), streamed FILE BY FILE.
//
// Why: a tcgen05.mma (M=128, K=16) costs max(N/2, 40 + 0.22 N) cycles -- it fetches its 4 KB A
// operand and 32 N bytes of B from shared memory for every instruction (tests/cuda/umma_bw_probe.cu).
// The first kernel (conv_kernel.cu) used one N=64 MMA per 3x3 tap, each on its own shifted A:
// 54 cycles for 32 cycles of math.  Here one A operand feeds the three taps of a kernel row:
//
//   tile      = 16 positions; GEMM rows m = (rank y, position p), p fastest -> M = 128
//   A operand = one FILE x of those 16 boards: shared-memory slab [K/8 chunks][10 padded ranks]
//               [16 positions][16 B], K-major no-swizzle with SBO = 128 B and LBO = 2 560 B; every
//               core matrix is 128 contiguous, 128-byte-aligned bytes and the vertical tap dy is
//               a +-256 B start address
//   scatter   = source file x contributes to output files x+1, x, x-1 through taps dx = -1, 0, +1:
//               three MMAs (N = 64 each for conv2, 32 for conv1) on the SAME A with
//               collector::a::fill/use/lastuse, each accumulating into the TMEM slot of its
//               output file.  3 x 32 cycles of math per 4 KB A + 6 KB B fetched (96-cycle floor met).
//
// The kernel is a stream over file steps g = 8 * tile + x with ring buffers everywhere:
//   planes slabs (8 x 5 KB), act1 slabs (8 x 10 KB), conv1 TMEM slots (4 x 32 columns), conv2 TMEM
//   slots (6 x 64 columns); an output file is complete once source files x-1, x, x+1 have been
//   issued.  Warp roles (all hand-overs by mbarriers):
//     warp 0, 1     MMA issuers: conv1(g) and conv2(g), each paced only by its own barriers
//     warps 2-5     plane encoder: thread = (rank, position), one square per step
//     warps 6-13    epilogue 1 (two groups of four, alternating output files): conv1 slot -> ReLU ->
//                   fp16 -> act1 slab (+ optional training copies)
//     warps 14-21   epilogue 2 (two groups of four, one per half of the output channels): conv2 slot -> +bias -> ReLU -> fp16 -> global, fc1 tile order
//
// Algorithmic work: 2 801 664 FLOP/position (dense convention);
// traffic 72 B in + 8 192 B out per position.
#pragma once
#include <cstdlib>
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {
namespace cs {

typedef unsigned long long u64;

// ---------------------------------------------------------------- L2-resident conv -> fc hand-over (pipe_kernel.cu)
// The hand-over unit is a group of PIPE_GF files (PIPE_GF x 8 squares = PIPE_GF x 128 KB) of a tile: the
// conv side produces a tile file by file and the fc side consumes its squares in the same file-major
// order, so a unit can be read as soon as its files are written and overwritten as soon as they
// have been loaded -- half a tile's lifetime on either side compared with whole-tile hand-over.
#ifndef CDQ_PIPE_GF
#define CDQ_PIPE_GF 8     // files per hand-over unit: 8 = whole tiles (measured best), 2 or 4 = finer units
#endif
constexpr int PIPE_GF = CDQ_PIPE_GF, PIPE_UNITS = 8 / PIPE_GF;
static_assert(PIPE_UNITS >= 1 && PIPE_UNITS <= 4, "B_UNIT_DONE / B_UNIT_FREE hold 2 x 4 barriers each");
struct PipeCtl {
    uint8_t* ring;            // ring_tiles x 1 MiB act2 tiles (128 positions each)
    unsigned ring_tiles;
    unsigned* ready;          // [ring_tiles][PIPE_UNITS] += 1 per (epilogue-2 warp, 16-position sub-tile): 64 per use
    unsigned* consumed;       // [ring_tiles][PIPE_UNITS] += 1 per use the fc side has finished loading
};
#ifndef CDQ_PIPE_SPIN_LIMIT
#define CDQ_PIPE_SPIN_LIMIT (1u << 27)   // a hand-over that never happens traps instead of hanging the GPU
#endif
__device__ __forceinline__ unsigned pipe_load(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void pipe_wait_ge(const unsigned* p, unsigned want) {
    unsigned v, spins = 0;
    while (true) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= want) break;
        if (++spins > CDQ_PIPE_SPIN_LIMIT) __trap();
        __nanosleep(64);
    }
}

constexpr int TILE_POS = 16;                      // positions per tile
constexpr int CELL = 16;                          // bytes per (square, 8-channel chunk)
constexpr int RANK_B = TILE_POS * CELL;           // 256 B: one rank of the 16 boards (two core matrices)
constexpr int CHUNK_B = 10 * RANK_B;              // 2 560 B: 10 padded ranks (= LBO)
constexpr int PL_SLAB = 2 * CHUNK_B;              // 16 input channels of one file
constexpr int A1_SLAB = 4 * CHUNK_B;              // 32 channels of one file
constexpr int RP = 8, RA = 8;                     // slab rings
constexpr int S1 = 4, S2 = 6;                     // TMEM slot rings
constexpr int W1_TAP_BYTES = 32 * 16 * 2;
constexpr int W2_TAP_BYTES = 64 * 32 * 2;
constexpr int W1_BYTES = 9 * W1_TAP_BYTES;
constexpr int W2_BYTES = 9 * W2_TAP_BYTES;

constexpr int OFF_PLANES = 0;
constexpr int OFF_ACT1 = OFF_PLANES + RP * PL_SLAB;
constexpr int OFF_W1 = OFF_ACT1 + RA * A1_SLAB;
constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
constexpr int OFF_B2 = OFF_W2 + W2_BYTES;         // conv2 bias, 64 floats
constexpr int OFF_BAR = OFF_B2 + 256;
enum { B_W = 0, B_PL_FULL = 1, B_PL_EMPTY = B_PL_FULL + RP, B_C1_FULL = B_PL_EMPTY + RP, B_C1_EMPTY = B_C1_FULL + S1,
       B_A1_FULL = B_C1_EMPTY + S1, B_A1_EMPTY = B_A1_FULL + RA, B_C2_FULL = B_A1_EMPTY + RA, B_C2_EMPTY = B_C2_FULL + S2,
       B_UNIT_DONE = B_C2_EMPTY + S2,      // [2 tile parities][4]: pipe mode, epilogue 2 -> signal warp
       B_UNIT_FREE = B_UNIT_DONE + 2 * 4,  // [2 tile parities][4]: signal warp -> epilogue 2 (barrier reusable)
       B_COUNT = B_UNIT_FREE + 2 * 4 };
constexpr int SMEM = OFF_BAR + B_COUNT * 8 + 16;
static_assert(SMEM <= 227 * 1024, "conv stream kernel shared memory");
static_assert(OFF_BAR % 8 == 0 && OFF_W1 % 128 == 0 && OFF_ACT1 % 128 == 0, "alignment");

constexpr int THREADS = 22 * 32;
constexpr int THREADS_PIPE = 23 * 32;   // pipe mode adds warp 22: publishes finished hand-over units
constexpr int TM_C2 = 0;            // conv2 slot s: columns TM_C2 + 64 s
constexpr int TM_C1 = 64 * S2;      // conv1 slot s: columns TM_C1 + 32 s
static_assert(TM_C1 + 32 * S1 <= 512, "TMEM columns");
static_assert(RP == 8, "the plane encoder addresses slab = file");

// layout of the training copy of ReLU(conv1) (consumed by conv_bwd2_kernel.cu): zero-bordered boards
// per group of 8 positions, [4 chunks][10 ranks][8 positions][10 files][16 B]
// training boards of ReLU(conv1), RANK-major in groups of FOUR positions (the unit conv_bwd2_kernel.cu fetches with one bulk
// copy): [group][10 ranks][4 channel chunks][4 positions][10 files][8 halfs]; (kernel row dy, channel chunk) is ONE linear
// index for conv2's weight gradient
constexpr int TB_ROW = 160, TB_CHUNK = 4 * TB_ROW, TB_RANK = 4 * TB_CHUNK, TB_GROUP = 10 * TB_RANK;

// ---------------------------------------------------------------- MMA issue of one source file step
// A source file feeds output files x-1, x, x+1 (dx blocks 0, 1, 2 of the re-ordered weights) whose
// TMEM slots d0, d1, d2 are consecutive except where the slot ring wraps.  The three taps of a kernel
// row go out as ONE MMA of N = 3 BLK, or, at a wrap, as two MMAs on the same A (collector fill /
// lastuse: 96 cycles either way, tests/cuda/umma_bw_probe.cu).  The first row of a step is also
// split where "accumulate" changes: a slot touched for the first time starts from zero.
//   FC   0 = first file of the tile (blocks 1, 2), 1 = middle, 2 = last file (blocks 0, 1)
//   WRAP 0 = none, 1 = ring wraps between d0 and d1, 2 = between d1 and d2
template <int BLK, int FC, int WRAP, bool FIRST>
__device__ __forceinline__ void issue_row(uint32_t d0, uint32_t d1, uint32_t d2, uint64_t a, uint64_t b) {
    constexpr uint64_t T = (BLK * 16) >> 4;      // one dx block of B in descriptor units
    constexpr uint32_t I1 = make_idesc_f16(128, BLK), I2 = make_idesc_f16(128, 2 * BLK), I3 = make_idesc_f16(128, 3 * BLK);
    if (FC == 0) {
        if (WRAP != 2) umma_f16(d1, a, b + T, I2, !FIRST);
        else { umma_f16_ca<0>(d1, a, b + T, I1, !FIRST); umma_f16_ca<2>(d2, a, b + 2 * T, I1, !FIRST); }
    } else if (FC == 2) {
        if (WRAP != 1) umma_f16(d0, a, b, I2, true);
        else { umma_f16_ca<0>(d0, a, b, I1, true); umma_f16_ca<2>(d1, a, b + T, I1, true); }
    } else if (FIRST) {
        if (WRAP != 1) { umma_f16_ca<0>(d0, a, b, I2, true); umma_f16_ca<2>(d2, a, b + 2 * T, I1, false); }
        else { umma_f16_ca<0>(d0, a, b, I1, true); umma_f16_ca<1>(d1, a, b + T, I1, true); umma_f16_ca<2>(d2, a, b + 2 * T, I1, false); }
    } else {
        if (WRAP == 0) umma_f16(d0, a, b, I3, true);
        else if (WRAP == 1) { umma_f16_ca<0>(d0, a, b, I1, true); umma_f16_ca<2>(d1, a, b + T, I2, true); }
        else { umma_f16_ca<0>(d0, a, b, I2, true); umma_f16_ca<2>(d2, a, b + 2 * T, I1, true); }
    }
}
// all kernel rows of a step: 3 dy x NK K-slices of 16 channels
template <int BLK, int NK, int FC, int WRAP, int R0, int R1>
__device__ __forceinline__ void issue_rows(uint32_t d0, uint32_t d1, uint32_t d2, uint64_t a_base, uint64_t w_desc) {
#pragma unroll
    for (int dyi = 0; dyi < 3; dyi++) {
#pragma unroll
        for (int k = 0; k < NK; k++) {
            if (dyi * NK + k < R0 || dyi * NK + k >= R1) continue;   // rows [R0, R1) of the step
            const uint64_t a = a_base + (uint64_t)((dyi * RANK_B + k * 2 * CHUNK_B) >> 4);
            const uint64_t b = w_desc + (uint64_t)((dyi * NK * 2 * 3 * BLK * 16 + k * 2 * 3 * BLK * 16) >> 4);
            if (dyi == 0 && k == 0) issue_row<BLK, FC, WRAP, true>(d0, d1, d2, a, b);
            else issue_row<BLK, FC, WRAP, false>(d0, d1, d2, a, b);
        }
    }
}
// ---------------------------------------------------------------- one file step of an issuer, all indices
// compile-time.  The issuers' control flow repeats with period lcm(8 files, ring sizes): 8 steps
// for conv1 (slabs 8, slots 4), 24 steps for conv2 (slabs 8, slots 6).  J is the step inside that
// period, so barrier numbers, TMEM slots, operand offsets and the MMA segmentation are immediates
// and a step costs a few dozen instructions.  (A first version computed them at run time --
// g % 6, a 9-way switch, descriptors through R2UR -- and the issuing warps, not the tensor pipe,
// set the pace: 1 650 cycles per step for 760 cycles of MMAs, ncu tensor pipe 48 %.)
struct IssueCtx {
    uint64_t* bars;
    uint32_t tmem;
    uint64_t a1_desc, a2_desc, w1_desc, w2_desc;   // descriptors of planes slab 0 / act1 slab 0 / re-ordered weights
    int lane;
};
__device__ __forceinline__ void wait_bar(const IssueCtx& c, int bar, uint32_t parity) { mbar_wait_warp(c.lane, c.bars + bar, parity); }

// conv1: J = file of the tile; `odd` = parity of the tile index (the planes ring turns once per tile)
template <int J>
__device__ __forceinline__ void c1_ready(const IssueCtx& c, uint32_t odd) {
    wait_bar(c, B_PL_FULL + J, odd);
    if (J == 0) wait_bar(c, B_C1_EMPTY + 0, 1);                      // output file 0 of a tile: slot 0, an even use
    if (J < 7) wait_bar(c, B_C1_EMPTY + (J + 1) % S1, (((J + 1) / S1) & 1) ^ 1);
    tc_fence_after();
}
template <int J>
__device__ __forceinline__ void c1_step(const IssueCtx& c, uint32_t odd, bool more) {
    static_assert(RP == 8 && S1 == 4, "conv1 issuer period");
    constexpr int FC = J == 0 ? 0 : J == 7 ? 2 : 1, SL = J % S1, WR = SL == 0 ? 1 : SL == S1 - 1 ? 2 : 0;
    const uint64_t a = c.a1_desc + (uint64_t)((J * PL_SLAB) >> 4);
    const uint32_t d0 = c.tmem + TM_C1 + 32 * ((J + S1 - 1) % S1), d1 = c.tmem + TM_C1 + 32 * SL, d2 = c.tmem + TM_C1 + 32 * ((J + 1) % S1);
    if (elect_one_sync()) issue_rows<32, 1, FC, WR, 0, 2>(d0, d1, d2, a, c.w1_desc);
    __syncwarp();
    if (J < 7) c1_ready<(J + 1) & 7>(c, odd);
    else if (more) c1_ready<0>(c, odd ^ 1);
    if (elect_one_sync()) {
        issue_rows<32, 1, FC, WR, 2, 3>(d0, d1, d2, a, c.w1_desc);
        umma_commit(c.bars + B_PL_EMPTY + J);
        if (J > 0) umma_commit(c.bars + B_C1_FULL + (J - 1) % S1);
        if (J == 7) umma_commit(c.bars + B_C1_FULL + SL);
    }
    __syncwarp();
}
// conv2: J = step inside three tiles; `odd` = parity of the three-tile group index
template <int J>
__device__ __forceinline__ void c2_ready(const IssueCtx& c, uint32_t odd) {
    wait_bar(c, B_A1_FULL + (J & 7), odd ^ ((J >> 3) & 1));          // act1 ring: use = 3 group + J/8
    if ((J & 7) == 0) wait_bar(c, B_C2_EMPTY + J % S2, ((J / S2) & 1) ^ 1);   // slot ring: use = 4 group + J/6
    if ((J & 7) < 7) wait_bar(c, B_C2_EMPTY + (J + 1) % S2, (((J + 1) / S2) & 1) ^ 1);
    tc_fence_after();
}
template <int J>
__device__ __forceinline__ void c2_step(const IssueCtx& c, uint32_t odd, bool more) {
    static_assert(RA == 8 && S2 == 6, "conv2 issuer period");
    constexpr int F = J & 7, FC = F == 0 ? 0 : F == 7 ? 2 : 1, SL = J % S2, WR = SL == 0 ? 1 : SL == S2 - 1 ? 2 : 0;
    const uint64_t a = c.a2_desc + (uint64_t)((F * A1_SLAB) >> 4);
    const uint32_t d0 = c.tmem + TM_C2 + 64 * ((J + S2 - 1) % S2), d1 = c.tmem + TM_C2 + 64 * SL, d2 = c.tmem + TM_C2 + 64 * ((J + 1) % S2);
    if (elect_one_sync()) issue_rows<64, 2, FC, WR, 0, 4>(d0, d1, d2, a, c.w2_desc);
    __syncwarp();
    if (F < 7) c2_ready<(J + 1) % 24>(c, odd);
    else if (more) c2_ready<(J + 1) % 24>(c, J == 23 ? odd ^ 1 : odd);
    if (elect_one_sync()) {
        issue_rows<64, 2, FC, WR, 4, 6>(d0, d1, d2, a, c.w2_desc);
        umma_commit(c.bars + B_A1_EMPTY + F);
        if (F > 0) umma_commit(c.bars + B_C2_FULL + (J + S2 - 1) % S2);
        if (F == 7) umma_commit(c.bars + B_C2_FULL + SL);
    }
    __syncwarp();
}
template <int T>   // the eight steps of tile T (0..2) of a three-tile group
__device__ __forceinline__ void c2_tile(const IssueCtx& c, uint32_t odd, bool more) {
    c2_step<8 * T + 0>(c, odd, true); c2_step<8 * T + 1>(c, odd, true); c2_step<8 * T + 2>(c, odd, true);
    c2_step<8 * T + 3>(c, odd, true); c2_step<8 * T + 4>(c, odd, true); c2_step<8 * T + 5>(c, odd, true);
    c2_step<8 * T + 6>(c, odd, true); c2_step<8 * T + 7>(c, odd, more);
}

#ifdef CDQ_TRACE
// timeline study build (python chess-deep-q_b200/build.py -DCDQ_TRACE --out=libcdq_trace.so):
// clock64 stamps of CTA 0, steps [TR_FROM, TR_FROM + 32), layout [role 4][step 32][tag 8]
static __device__ long long* g_trace;
constexpr int TR_FROM = 96;
#define TR(role, step, tag)                                                                             \
    do {                                                                                                \
        if (tr_buf && lane == 0 && (step) >= TR_FROM && (step) < TR_FROM + 32)                            \
            tr_buf[((role) * 32 + (step) - TR_FROM) * 8 + (tag)] = clock64();                            \
    } while (0)
#else
#define TR(role, step, tag) do {} while (0)
#endif

template <bool PIPE>
__device__ __forceinline__ void conv_stream_body(uint8_t* smem, const CdqBoard* __restrict__ boards, long long n,
                                                 const PackedNet& net, __half* __restrict__ act2,
                                                 __half* __restrict__ dbg_act1, __half* __restrict__ act1_boards,
                                                 const int cta, const int n_cta, const PipeCtl& pipe) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef CDQ_TRACE
    long long* const tr_buf = (cta == 0 && !PIPE) ? g_trace : nullptr;
#endif

    // ---- one-time setup: zero both slab rings (rank borders stay zero forever), barriers, weights
    for (int i = tid; i < OFF_W1 / 16; i += (int)blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid < 64) reinterpret_cast<float*>(smem + OFF_B2)[tid] = net.conv2_b[tid];
    if (tid == 0) {
        mbar_init(bars + B_W, 1);
        for (int b = 0; b < RP; b++) { mbar_init(bars + B_PL_FULL + b, 4); mbar_init(bars + B_PL_EMPTY + b, 1); }
        for (int b = 0; b < RA; b++) { mbar_init(bars + B_A1_FULL + b, 4); mbar_init(bars + B_A1_EMPTY + b, 1); }
        for (int b = 0; b < S1; b++) { mbar_init(bars + B_C1_FULL + b, 1); mbar_init(bars + B_C1_EMPTY + b, 4); }
        for (int b = 0; b < S2; b++) { mbar_init(bars + B_C2_FULL + b, 1); mbar_init(bars + B_C2_EMPTY + b, 8); }
        for (int b = 0; b < 2 * 4; b++) { mbar_init(bars + B_UNIT_DONE + b, 8); mbar_init(bars + B_UNIT_FREE + b, 1); }
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        // weights, already in this kernel's order (PackedNet::wss, written by pack_weights_kernel):
        // [dy][K chunk][dx = +1, 0, -1][n groups], so that the taps of one kernel row form ONE B operand of
        // N = 3 x 64 (3 x 32) rows whose row order matches ascending output files
        mbar_arrive_expect_tx(bars + B_W, W1_BYTES + W2_BYTES);
        bulk_g2s(smem + OFF_W1, net.wss, W1_BYTES + W2_BYTES, bars + B_W);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // whole 128-position fc1 tiles are always written (missing boards count as empty boards)
    const long long n_tiles = ((n + 127) >> 7) * (128 / TILE_POS);
    const int n_local = (int)((n_tiles - cta + n_cta - 1) / n_cta);   // tiles of this CTA
    const int G = 8 * n_local;                                                        // file steps of this CTA

    if (warp <= 1) {
        // =============================== MMA issuers: warp 0 conv1, warp 1 conv2 (warp-uniform control, one
        // elected lane issues; two warps so that one's barrier waits overlap the other's queued MMAs)
        mbar_wait_warp(lane, bars + B_W, 0);
        const uint32_t smem_a = smem_u32(smem);
        constexpr uint64_t A_HI = ((uint64_t)(128 >> 4) << 32) | (1ull << 46) | ((uint64_t)(CHUNK_B >> 4) << 16);
        constexpr uint64_t B_HI = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        IssueCtx c;
        c.bars = bars; c.tmem = tmem; c.lane = lane;
        c.a1_desc = A_HI | (uint64_t)(((smem_a + OFF_PLANES) >> 4) & 0x3fff);
        c.a2_desc = A_HI | (uint64_t)(((smem_a + OFF_ACT1) >> 4) & 0x3fff);
        // B operands: LBO = one K chunk of all three dx blocks
        c.w1_desc = B_HI | ((uint64_t)(1536 >> 4) << 16) | (uint64_t)(((smem_a + OFF_W1) >> 4) & 0x3fff);
        c.w2_desc = B_HI | ((uint64_t)(3072 >> 4) << 16) | (uint64_t)(((smem_a + OFF_W2) >> 4) & 0x3fff);
        // Both issuers take the barrier waits of step g+1 in the MIDDLE of step g's MMAs, so the
        // tensor pipe always has queued work while the issuing thread waits.
        if (warp == 0) {
            if (n_local > 0) c1_ready<0>(c, 0);
            for (int i = 0; i < n_local; i++) {
                const uint32_t odd = i & 1;
                TR(0, 8 * i, 0);
                c1_step<0>(c, odd, true); c1_step<1>(c, odd, true); c1_step<2>(c, odd, true); c1_step<3>(c, odd, true);
                c1_step<4>(c, odd, true); c1_step<5>(c, odd, true); c1_step<6>(c, odd, true);
                c1_step<7>(c, odd, i + 1 < n_local);
            }
        } else {
            if (n_local > 0) c2_ready<0>(c, 0);
            for (int i = 0, grp = 0; i < n_local; i += 3, grp++) {
                const uint32_t odd = grp & 1;
                TR(0, 8 * i, 3);
                c2_tile<0>(c, odd, i + 1 < n_local);
                TR(0, 8 * i + 8, 3);
                if (i + 1 < n_local) c2_tile<1>(c, odd, i + 2 < n_local);
                TR(0, 8 * i + 16, 3);
                if (i + 2 < n_local) c2_tile<2>(c, odd, i + 3 < n_local);
            }
        }
    } else if (warp <= 5) {
        // =============================== plane encoder: thread = (rank r, position p)
        const int e = tid - 64, r = e >> 4, p = e & 15;
        const uint8_t* gb = reinterpret_cast<const uint8_t*>(boards);
        // rank byte r of the 8 bitboards of one board.  The bytes of tile i+1 are requested at the start of
        // tile i and stay RAW in registers until tile i+1 begins: nothing depends on them before that, so
        // the loads (DRAM latency) really are in flight during the tile's eight steps.  (With the
        // channel-pair words computed at once the warp stalled on them and every tile began with a
        // ~2 800-cycle bubble for all roles.)
        auto fetch = [&](int i, unsigned (&b)[8]) {
            const long long pos = (cta + (long long)i * n_cta) * TILE_POS + p;
            const bool ok = i < n_local && pos < n;
#pragma unroll
            for (int t = 0; t < 8; t++) b[t] = ok ? (unsigned)__ldg(gb + pos * 72 + t * 8 + r) : 0u;
        };
        unsigned raw[8], cur[6];
        fetch(0, raw);
        uint8_t* dst0 = smem + OFF_PLANES + (r + 1) * RANK_B + p * CELL;
        for (int i = 0; i < n_local; i++) {
            // channel-pair words: bit x of byte 0 / byte 2 = channel 2j / 2j+1 present on file x
            // (channel = type + 6*black, board_utils.py:17-30)
            const unsigned w = raw[6], k = raw[7] & ~raw[6];
            cur[0] = (raw[0] & w) | ((raw[1] & w) << 16);
            cur[1] = (raw[2] & w) | ((raw[3] & w) << 16);
            cur[2] = (raw[4] & w) | ((raw[5] & w) << 16);
            cur[3] = (raw[0] & k) | ((raw[1] & k) << 16);
            cur[4] = (raw[2] & k) | ((raw[3] & k) << 16);
            cur[5] = (raw[4] & k) | ((raw[5] & k) << 16);
            fetch(i + 1, raw);
            // two files per iteration: one proxy fence and one round of barrier latencies per pair
#pragma unroll 1
            for (int f = 0; f < 8; f += 2) {
                const int g = i * 8 + f;                // RP = 8: slab = file, ring use = tile index
                const uint32_t free_parity = (i & 1) ^ 1;
                TR(1, g, 0);
                mbar_wait_warp(lane, bars + B_PL_EMPTY + f, free_parity);
                mbar_wait_warp(lane, bars + B_PL_EMPTY + f + 1, free_parity);
                TR(1, g, 1);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    unsigned v[6];
#pragma unroll
                    for (int j = 0; j < 6; j++) v[j] = ((cur[j] >> (f + h)) & 0x00010001u) * 0x3C00u;   // fp16 1.0 per present channel
                    uint8_t* dst = dst0 + (f + h) * PL_SLAB;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<uint4*>(dst + CHUNK_B) = make_uint4(v[4], v[5], 0x3C00u, 0u);   // channel 12: constant one
                }
                TR(1, g, 2);
                fence_proxy_async();
                __syncwarp();
                TR(1, g, 3);
                if (lane == 0) { mbar_arrive(bars + B_PL_FULL + f); mbar_arrive(bars + B_PL_FULL + f + 1); }
            }
        }
    } else if (warp <= 13) {
        // =============================== epilogue 1: conv1 slot -> act1 slab; two groups of four warps,
        // warps 6-9 take the even output files, warps 10-13 the odd ones (a step's fixed latencies --
        // barrier polls, tcgen05.ld, proxy fence -- add up to more than the 760-cycle MMA budget)
        const int q = warp & 3;
        const int y = 2 * q + (lane >> 4), p = lane & 15;      // TMEM lane 32q + lane = row (rank y, position p)
        uint8_t* dst0 = smem + OFF_ACT1 + (y + 1) * RANK_B + p * CELL;
#pragma unroll 1
        for (int o = (warp - 6) >> 2; o < G; o += 2) {
            const int slot = o % S1, slab = o % RA, f = o & 7;
            TR(2, o, 0);
            mbar_wait_warp(lane, bars + B_C1_FULL + slot, (o / S1) & 1);
            tc_fence_after();
            TR(2, o, 1);
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_C1 + 32 * slot, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + B_C1_EMPTY + slot);
            TR(2, o, 2);
            mbar_wait_warp(lane, bars + B_A1_EMPTY + slab, ((o / RA) & 1) ^ 1);   // conv2 has read this slab's last use
            TR(2, o, 3);
            uint8_t* dst = dst0 + slab * A1_SLAB;
            long long pos = 0;
            if (act1_boards || dbg_act1) pos = (cta + (long long)(o >> 3) * n_cta) * TILE_POS + p;
#pragma unroll
            for (int kc = 0; kc < 4; kc++) {
                uint4 u;
                u.x = relu_pack_f16(__uint_as_float(v[kc * 8 + 0]), __uint_as_float(v[kc * 8 + 1]));
                u.y = relu_pack_f16(__uint_as_float(v[kc * 8 + 2]), __uint_as_float(v[kc * 8 + 3]));
                u.z = relu_pack_f16(__uint_as_float(v[kc * 8 + 4]), __uint_as_float(v[kc * 8 + 5]));
                u.w = relu_pack_f16(__uint_as_float(v[kc * 8 + 6]), __uint_as_float(v[kc * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + kc * CHUNK_B) = u;
                if (act1_boards)   // training: ReLU(conv1) for conv2's backward pass, POSITION-contiguous [tile of 16][file][chunk][rank]
                                   // [16 positions][8 halfs]: 4 lines per warp store; a1_boards_kernel (train_kernels.cu) turns it
                                   // into boards on the side stream (16 bytes straight into the boards were 32 lines per store)
                    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(act1_boards) + (pos >> 4) * 65536 + f * 8192 + kc * 2048 +
                                              y * 256 + (int)(pos & 15) * CELL) = u;
                if (dbg_act1 && pos < n)
                    *reinterpret_cast<uint4*>(dbg_act1 + (pos * 64 + y * 8 + f) * 32 + kc * 8) = u;
            }
            TR(2, o, 4);
            fence_proxy_async();
            __syncwarp();
            TR(2, o, 5);
            if (lane == 0) mbar_arrive(bars + B_A1_FULL + slab);
        }
    } else if (warp <= 21) {
        // =============================== epilogue 2: conv2 slot -> global (fc1 A-tile order); eight warps,
        // warps 14-17 take output channels 0-31, warps 18-21 channels 32-63 of every output file
        const int q = warp & 3, hh = (warp - 14) >> 2;
        const int y = 2 * q + (lane >> 4), p = lane & 15;
        float4 b2r[8];    // this half's conv2 bias, kept in registers for the whole kernel
#pragma unroll
        for (int k = 0; k < 8; k++) b2r[k] = reinterpret_cast<const float4*>(smem + OFF_B2)[hh * 8 + k];
        int slot = 0;            // ring position and phase parity of the next output file
        uint32_t phase = 0;
#pragma unroll 1
        for (int i = 0; i < n_local; i++) {
            const long long pos = (cta + (long long)i * n_cta) * TILE_POS + p;
            const int pr = (int)(pos & 127);
            // 16 KB block per (128-position tile, square): [8 channel chunks][16 row groups][8 rows][8 halfs]
            uint8_t* dst = reinterpret_cast<uint8_t*>(act2) + ((pos >> 7) << 20) + ((long long)(y * 8) << 14) +
                           (pr >> 3) * 128 + (pr & 7) * 16 + hh * 4 * 2048;
            unsigned ring_slot = 0, ring_use = 0;
            bool slot_free = true;
            if (PIPE) {
                // producer side of the L2-resident hand-over (pipe_kernel.cu): 128-position tile T lives in
                // ring slot T % R.  The fc side consumes the units of a slot in order, so if the LAST unit
                // of the previous occupant has been loaded the whole slot is free (the common case, one poll);
                // otherwise every unit is waited for just before its first file is written.
                const unsigned T = (unsigned)(pos >> 7);
                ring_use = T / pipe.ring_tiles;
                ring_slot = T % pipe.ring_tiles;
                unsigned seen = 0;
                if (lane == 0) seen = pipe_load(pipe.consumed + ring_slot * PIPE_UNITS + PIPE_UNITS - 1);
                slot_free = __shfl_sync(0xffffffffu, seen, 0) >= ring_use;
                dst = pipe.ring + ((size_t)ring_slot << 20) + ((long long)(y * 8) << 14) + (pr >> 3) * 128 + (pr & 7) * 16 +
                      hh * 4 * 2048;
            }
#pragma unroll 1
            for (int f = 0; f < 8; f++, dst += 1 << 14) {
                const int o = 8 * i + f;
                TR(3, o, 0);
                if (PIPE && !slot_free && f % PIPE_GF == 0) {
                    if (lane == 0) pipe_wait_ge(pipe.consumed + ring_slot * PIPE_UNITS + f / PIPE_GF, ring_use);
                    __syncwarp();
                }
                mbar_wait_warp(lane, bars + B_C2_FULL + slot, phase);
                tc_fence_after();
                TR(3, o, 1);
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_C2 + 64 * slot + 32 * hh, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + B_C2_EMPTY + slot);
                TR(3, o, 2);
                if (++slot == S2) { slot = 0; phase ^= 1; }
#pragma unroll
                for (int kc = 0; kc < 4; kc++) {
                    const float4 ba = b2r[kc * 2], bb = b2r[kc * 2 + 1];
                    uint4 u;
                    u.x = relu_pack_f16(__uint_as_float(v[kc * 8 + 0]) + ba.x, __uint_as_float(v[kc * 8 + 1]) + ba.y);
                    u.y = relu_pack_f16(__uint_as_float(v[kc * 8 + 2]) + ba.z, __uint_as_float(v[kc * 8 + 3]) + ba.w);
                    u.z = relu_pack_f16(__uint_as_float(v[kc * 8 + 4]) + bb.x, __uint_as_float(v[kc * 8 + 5]) + bb.y);
                    u.w = relu_pack_f16(__uint_as_float(v[kc * 8 + 6]) + bb.z, __uint_as_float(v[kc * 8 + 7]) + bb.w);
#ifndef CDQ_EXP_NOSTORE   // A/B experiment only: how much of the kernel time is the act2 write stream
                    *reinterpret_cast<uint4*>(dst + kc * 2048) = u;
#else
                    if (u.x == 0x12345678u) *reinterpret_cast<uint4*>(dst + kc * 2048) = u;
#endif
                }
                TR(3, o, 3);
                if (PIPE && f % PIPE_GF == PIPE_GF - 1) {
                    // hand the finished unit to the signal warp (release at CTA scope, no fence stall here);
                    // the barrier pair is reused every second tile, so wait until its previous use was taken
                    const int ub = (i & 1) * 4 + f / PIPE_GF;
                    mbar_wait_warp(lane, bars + B_UNIT_FREE + ub, ((i >> 1) & 1) ^ 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bars + B_UNIT_DONE + ub);
                }
            }
        }
    }
    else if (PIPE) {
        // =============================== signal warp (pipe mode): when all eight epilogue-2 warps have stored a
        // hand-over unit of a sub-tile, make those stores visible device-wide (the mbarrier acquire at
        // CTA scope followed by a device-scope fence is cumulative) and count them into `ready`; the
        // epilogue warps themselves never wait for a fence
#pragma unroll 1
        for (int i = 0; i < n_local; i++) {
            const unsigned T = (unsigned)(((cta + (long long)i * n_cta) * TILE_POS) >> 7);
            unsigned* ready = pipe.ready + (T % pipe.ring_tiles) * PIPE_UNITS;
#pragma unroll 1
            for (int u = 0; u < PIPE_UNITS; u++) {
                if (lane == 0) {
                    mbar_wait(bars + B_UNIT_DONE + (i & 1) * 4 + u, (i >> 1) & 1);
                    mbar_arrive(bars + B_UNIT_FREE + (i & 1) * 4 + u);
                    __threadfence();
                    atomicAdd(ready + u, 8u);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

} // namespace cs
} // namespace cdq
