// conv_bwd2_kernel.cu -- backward of conv2 (32 -> 64, 3x3) on tcgen05, pipelined over HALF groups.
//
// Operands are zero-bordered boards, so that every 3x3 tap is an operand start address:
//   dz2 boards  [group of 4 positions][8 chunks of 8 out-channels][10 ranks][4 positions][10 files][8 halfs]   (fc1_dgrad_kernel)
//   a1  boards  [group of 4 positions][10 ranks][4 chunks of 8 in-channels][4 positions][10 files][8 halfs]    (the forward's epilogue 1)
//   weight gradient  dW2[co][ci][tap] += sum_cells dz2[cell][co] * a1[cell + tap][ci]      M = 64, N = 3 x 32, K = 16 cells, MN-major
//       The a1 boards are RANK-major: (kernel row dy, channel chunk) is one linear index with a 640-byte stride,
//       so the three taps of a kernel COLUMN are one B operand of N = 96 and a UMMA fills three tap accumulators.
//       A small UMMA costs about 40 + 0.22 N cycles whatever it computes, so what counts is their number:
//       48 per unit instead of 144 (one per tap, N = 32).
//   bias gradient    db2[co] += sum_cells dz2[cell][co]                                    same A against a tile of ones
//   data gradient    dz1[cell][ci] = ReLU'(a1) * sum_{co,tap} dz2[cell - tap][co] W2[co][ci][tap]   M = 128, N = 32, K = 64
// The first kernel (removed) handled a whole group (150 KB of boards) per iteration with one buffer: load, UMMAs and
// epilogue ran one after the other (3.4 + 7.4 + 3 us per group, four rounds of 148 CTAs over 512 groups, the
// last one 46 % full).  Here the unit of work is a group of FOUR positions (75 KB of boards, which the producers of
// the boards lay out so that a unit is two contiguous blocks: two bulk copies; as 120 copies of 640 bytes out of
// 8-position groups the fetch, not the UMMAs, set the pace), so that
//   * two units fit in shared memory: the producer warp fetches unit i+1 while the UMMAs of unit i run;
//   * the data-gradient accumulators are double buffered in TMEM: the epilogue warps drain unit i while the
//     UMMAs of unit i+1 run;
//   * 1 024 units spread over 148 CTAs in 7 rounds of half the size (3.5 group-rounds instead of 4);
//   * units are handed out DYNAMICALLY (one atomicAdd per unit by the producer, the index travels to the other warps
//     through shared memory next to the full barrier): in the training graph this kernel starts while the side
//     stream's fc1 weight gradient still holds some SMs, and with a static round-robin split the CTAs that start
//     late finished late by as much (+7 us on the whole kernel).  The two counters reset themselves.
// Warp roles: 0 producer, 1 UMMA issuer, 2..5 epilogue (TMEM lane quarter = warp % 4).
#include <cstdlib>
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {
namespace cb2 {

constexpr int ROW = 160, RANK = 640, CHUNK = 6400;                  // dz2 unit strides, global = shared memory
constexpr int A_CHUNK = 4 * ROW, A_RANK = 4 * A_CHUNK;               // a1 unit, rank-major: 640 / 2 560
constexpr int U_DZ2 = 8 * CHUNK, U_A1 = 10 * A_RANK, U_BYTES = U_DZ2 + U_A1;   // 51 200 + 25 600 = 76 800
constexpr int W_BYTES = 9 * 64 * 32 * 2;                            // 36 864 : [9 taps][8 chunks of co][4 groups of ci][8 ci][8 co]
constexpr int OFF_W = 2 * U_BYTES;
constexpr int OFF_ONES = OFF_W + W_BYTES;
constexpr int OFF_BAR = OFF_ONES + 256;
constexpr int SMEM = OFF_BAR + 128;
static_assert(SMEM <= 227 * 1024, "conv2 backward shared memory");
constexpr int THREADS = 192;
constexpr int TM_BIAS = 288, TM_DG = 320;       // TMEM columns: taps at 96 dx + 32 dy + ci, bias 288..295, data-gradient 320 + 64 buffer + 32 tile
constexpr int WT_STRIDE = 292;                  // floats per co row of the staged weight-gradient tile
static_assert(64 * WT_STRIDE * 4 <= U_BYTES, "weight-gradient tile is staged over unit buffer 0");

__global__ void __launch_bounds__(THREADS, 1)
conv2_bwd2_kernel(const __half* __restrict__ dz2_boards, const __half* __restrict__ a1_boards,
                  const __half* __restrict__ w2tp, long long n_units, long long n, float inv_scale,
                  float* __restrict__ g_w2, float* __restrict__ g_b2, float* __restrict__ dz1c,
                  unsigned* __restrict__ sched /* [0] next unit, [1] finished CTAs; zero at entry, zero again at exit */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* full = bar_w + 1;          // [2] unit buffer filled (bulk-copy bytes)
    uint64_t* empty = bar_w + 3;         // [2] unit buffer free: UMMAs done (1) + the four epilogue warps have their masks (4)
    uint64_t* dg_full = bar_w + 5;       // [2] data-gradient accumulators of TMEM buffer b complete
    uint64_t* dg_empty = bar_w + 7;      // [2] ... drained by the four epilogue warps
    uint64_t* wg_full = bar_w + 9;       //     every UMMA of this CTA complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 10);
    volatile int* u_slot = reinterpret_cast<volatile int*>(bar_w + 11);   // [2] unit index of buffer b, -1 = no more work
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < 128) reinterpret_cast<uint16_t*>(smem + OFF_ONES)[tid] = 0x3C00;   // fp16 1.0
    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int b = 0; b < 2; b++) { mbar_init(full + b, 1); mbar_init(empty + b, 5); mbar_init(dg_full + b, 1); mbar_init(dg_empty + b, 4); }
        mbar_init(wg_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t s0 = smem_u32(smem);

    if (warp == 0) {
        // ---------------- producer: weights once, then one unit (half group) per iteration
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, W_BYTES);
            bulk_g2s(smem + OFF_W, w2tp, W_BYTES, bar_w);
        }
        for (long long i = 0;; i++) {
            const int b = (int)(i & 1);
            mbar_wait(empty + b, ((i >> 1) & 1) ^ 1);
            int more = 1;
            if (lane == 0) {
                const unsigned u = atomicAdd(sched, 1u);
                if ((long long)u < n_units) {
                    u_slot[b] = (int)u;
                    mbar_arrive_expect_tx(full + b, U_BYTES);
                    uint8_t* sb = smem + b * U_BYTES;
                    bulk_g2s(sb, reinterpret_cast<const uint8_t*>(dz2_boards) + (long long)u * U_DZ2, U_DZ2, full + b);
                    bulk_g2s(sb + U_DZ2, reinterpret_cast<const uint8_t*>(a1_boards) + (long long)u * U_A1, U_A1, full + b);
                } else {
                    u_slot[b] = -1;
                    mbar_arrive(full + b);          // wakes the other warps; nothing is in flight
                    more = 0;
                }
            }
            if (!__shfl_sync(0xffffffffu, more, 0)) break;
        }
    } else if (warp == 1) {
        // ---------------- UMMA issuer
        if (elect_one_sync()) {
            constexpr uint32_t ID_WG = make_idesc_f16_ex(64, 96, true, true);
            constexpr uint32_t ID_B1 = make_idesc_f16_ex(64, 8, true, true);
            constexpr uint32_t ID_DG = make_idesc_f16(128, 32);
            mbar_wait(bar_w, 0);
            // Every operand descriptor is a per-buffer base plus a COMPILE-TIME offset (all loops below are fully
            // unrolled): with run-time descriptor arithmetic the single issuing thread, not the tensor pipe, set
            // the pace (about 50 cycles per UMMA, 232 UMMAs per unit).
            const uint64_t dA_wg0 = make_smem_desc(s0 + 4 * ROW + 16, ROW, CHUNK);             // dz2, MN-major, interior
            const uint64_t dB_wg0 = make_smem_desc(s0 + U_DZ2, ROW, A_CHUNK);                  // a1,  MN-major: rank 0 = dy -1 of the first interior rank, file 0 = dx -1
            const uint64_t dA_dg0 = make_smem_desc(s0, CHUNK, ROW);                            // dz2, K-major
            const uint64_t dB_w = make_smem_desc(s0 + OFF_W, 512, 128);
            const uint64_t d_ones = make_smem_desc(s0 + OFF_ONES, 128, 128);
            for (long long i = 0;; i++) {
                const int b = (int)(i & 1);
                const uint32_t ph = (uint32_t)((i >> 1) & 1);
                mbar_wait(full + b, ph);
                if (u_slot[b] < 0) break;
                mbar_wait(dg_empty + b, ph ^ 1);
                tc_fence_after();
                const uint64_t boff = (uint64_t)(b * (U_BYTES >> 4));
                const uint64_t dA_wg = dA_wg0 + boff, dB_wg = dB_wg0 + boff, dA_dg = dA_dg0 + boff;
                const bool acc0 = i != 0;
                // ---- weight gradient: 16 steps of 16 cells (two rows of 8 files, interior row groups 4..35) x 3 kernel
                // columns; B = the a1 rows of ranks R-1, R, R+1 (N = 3 x 32, 640 bytes between 8-channel groups)
#pragma unroll
                for (int k = 0; k < 16; k++)
#pragma unroll
                    for (int dxi = 0; dxi < 3; dxi++)
                        umma_f16(tmem + dxi * 96, dA_wg + (uint64_t)((k * 2 * ROW) >> 4),
                                 dB_wg + (uint64_t)(((k >> 1) * A_RANK + (k & 1) * 2 * ROW + dxi * 16) >> 4), ID_WG, k != 0 || acc0);
                // ---- bias gradient: same A against ones
#pragma unroll
                for (int k = 0; k < 16; k++)
                    umma_f16(tmem + TM_BIAS, dA_wg + (uint64_t)((k * 2 * ROW) >> 4), d_ones, ID_B1, k != 0 || acc0);
                // ---- data gradient: 2 tiles (4 ranks x 4 positions x 8 files) x 9 flipped taps x 4 K-steps of 16 out-channels
                const uint32_t dg = tmem + TM_DG + 64 * b;
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int u9 = 0; u9 < 9; u9++) {
                        const int dy = u9 / 3 - 1, dx = u9 % 3 - 1;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            umma_f16(dg + 32 * t, dA_dg + (uint64_t)(((4 * t + dy + 1) * RANK + (dx + 1) * 16 + j * 2 * CHUNK) >> 4),
                                     dB_w + (uint64_t)((u9 * 4096 + j * 2 * 512) >> 4), ID_DG, (u9 | j) != 0);
                    }
                umma_commit(dg_full + b);
                umma_commit(empty + b);
            }
            umma_commit(wg_full);
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps: thread = data-gradient tile row (rank in tile, position, file)
        const int q = warp & 3, row = q * 32 + lane;
        const int r_rk = row >> 5, r_pos = (row >> 3) & 3, r_x = row & 7;
        long long i = 0;
        for (;; i++) {
            const int b = (int)(i & 1);
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            mbar_wait(full + b, ph);                 // the boards (for the mask) ...
            const long long u = u_slot[b];
            if (u < 0) break;
            mbar_wait(dg_full + b, ph);              // ... and the accumulators
            tc_fence_after();
            // conv1's ReLU mask of this thread's two rows (32 channels each), then the boards are released
            unsigned relu_mask[2];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int y = 4 * t + r_rk;
                const uint8_t* m = smem + b * U_BYTES + U_DZ2 + (y + 1) * A_RANK + r_pos * ROW + (r_x + 1) * 16;
                unsigned bits = 0;
#pragma unroll
                for (int kc = 0; kc < 4; kc++) {
                    const uint4 mk = *reinterpret_cast<const uint4*>(m + kc * A_CHUNK);
                    const unsigned mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        bits |= (((mw[e >> 1] >> (16 * (e & 1))) & 0xffffu) ? 1u : 0u) << (kc * 8 + e);
                }
                relu_mask[t] = bits;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);
            uint32_t v[2][32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_DG + 64 * b, v[0]);
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_DG + 64 * b + 32, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dg_empty + b);
            const long long pos = u * 4 + r_pos;
            if (pos < n) {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    const int y = 4 * t + r_rk;
                    const unsigned bits = relu_mask[t];
                    float4* o = reinterpret_cast<float4*>(dz1c + (pos * 64 + y * 8 + r_x) * 32);
#pragma unroll
                    for (int q4 = 0; q4 < 8; q4++) {
                        float r[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            r[e] = ((bits >> (q4 * 4 + e)) & 1u) ? __uint_as_float(v[t][q4 * 4 + e]) * inv_scale : 0.f;
                        o[q4] = make_float4(r[0], r[1], r[2], r[3]);
                    }
                }
            }
        }
        // ---- weight/bias gradient: rows co = 16 * quadrant + lane (lanes 0..15 hold data), transposed through shared
        // memory (unit buffer 0 is dead: every UMMA is complete and the producer has nothing in flight) into PyTorch
        // order [co][ci][tap] and added with coalesced 16-byte reductions
        if (i > 0) {
            mbar_wait(wg_full, 0);
            tc_fence_after();
            float* tile = reinterpret_cast<float*>(smem);              // [64 co][WT_STRIDE]
            const int co = q * 16 + lane;
#pragma unroll 1
            for (int tap = 0; tap < 9; tap++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (tap % 3) * 96 + (tap / 3) * 32, v);   // tap = 3 (dy + 1) + (dx + 1)
                tmem_ld_wait();
                if (lane < 16) {
#pragma unroll
                    for (int ci = 0; ci < 32; ci++) tile[co * WT_STRIDE + ci * 9 + tap] = __uint_as_float(v[ci]) * inv_scale;
                }
            }
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_BIAS, v);   // columns 288..319; the first 8 are the bias sums
            tmem_ld_wait();
            if (lane < 16) atomicAdd(g_b2 + co, __uint_as_float(v[0]) * inv_scale);
            asm volatile("bar.sync 1, 128;" ::: "memory");              // the four epilogue warps
            for (int j = row; j < 64 * 72; j += 128) {
                const int r = j / 72, c4 = j - r * 72;
                const float4 t4 = *reinterpret_cast<const float4*>(tile + r * WT_STRIDE + 4 * c4);
                atomicAdd(reinterpret_cast<float4*>(g_w2) + j, t4);     // conv2.weight[co][ci][tap], 288 floats per co
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
    if (tid == 0 && atomicAdd(sched + 1, 1u) == gridDim.x - 1) {     // last CTA out: every producer has seen the end
        sched[0] = 0;
        sched[1] = 0;
    }
}

} // namespace cb2

int launch_conv2_bwd(const __half* dz2_boards, const __half* a1_boards, const __half* w2tp, long long n, float inv_scale,
                     float* g_w2, float* g_b2, float* dz1c, unsigned* sched, cudaStream_t st) {
    once_per_device(ONCE_CONV2_BWD, [] { cudaFuncSetAttribute(cb2::conv2_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cb2::SMEM); });
    const long long units = ((n + 127) >> 7) * 32;      // groups of 4 positions
    const int sms = device_sm_count();
    ProfileScope ps(KK_CONV2_BWD, st);
    cb2::conv2_bwd2_kernel<<<(unsigned)(units < sms ? units : sms), cb2::THREADS, cb2::SMEM, st>>>(
        dz2_boards, a1_boards, w2tp, units, n, inv_scale, g_w2, g_b2, dz1c, sched);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
