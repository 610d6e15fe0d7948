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
//   bias gradient    db2[co] += sum_cells dz2[cell][co]                                    on the CUDA cores, from the unit in shared memory
//   data gradient    dz1[cell][ci] = ReLU'(a1) * sum_{co,tap} dz2[cell - tap][co] W2[co][ci][tap]   M = 128, N = 3 x 32, K = 64
//       The same count argument: one UMMA per (kernel row dy, K step) whose B holds the three taps dx of that row
//       (N = 96) on the A operand of file offset 0 -- 24 per unit instead of 72 (one per tap, N = 32).  Accumulator
//       block dx then holds P_dx[cell] = sum_{dy,co} dz2[cell + (dy, 0)][co] W2f[co][ci][dy][dx], and
//       dz1[cell] = P_-1[cell - 1 file] + P_0[cell] + P_+1[cell + 1 file].  A tile row is (rank, position, file 0..7)
//       with the file fastest, so the file neighbours are the neighbouring LANES of the epilogue warp (two shuffles
//       per channel) and the terms that would come from the zero border are dropped by a lane predicate.
//       The bias gradient used to be 16 more UMMAs per unit (A against a tile of ones, N = 8: 42 cycles each for
//       nothing); the epilogue warps now add the unit's 64 x 320 values while the UMMAs run.
// The first kernel (removed) handled a whole group (150 KB of boards) per iteration with one buffer: load, UMMAs and
// epilogue ran one after the other (3.4 + 7.4 + 3 us per group, four rounds of 148 CTAs over 512 groups, the
// last one 46 % full).  Here the unit of work is a group of FOUR positions (75 KB of boards, which the producers of
// the boards lay out so that a unit is two contiguous blocks: two bulk copies; as 120 copies of 640 bytes out of
// 8-position groups the fetch, not the UMMAs, set the pace), so that
//   * two units fit in shared memory: the producer warp fetches unit i+1 while the UMMAs of unit i run;
//   * the data-gradient accumulators (2 tiles x 96 columns; with the 288 weight-gradient columns TMEM is full) are
//     drained by the epilogue warps while the weight-gradient UMMAs of unit i+1 run: the issuer waits for the drain
//     only before the data-gradient UMMAs of unit i+1;
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

#ifdef CB2_TRACE
// clock64 stamps of CTA 0 (profiles/cb2_trace.py): [role 0 producer / 1 issuer / 2 epilogue][unit slot 0..15][stamp 0..7]
__device__ long long g_cb2_trace[3 * 16 * 8];
#define CB2_STAMP(role, i, k) do { if (blockIdx.x == 0 && (i) < 16) g_cb2_trace[((role) * 16 + (int)(i)) * 8 + (k)] = clock64(); } while (0)
#else
#define CB2_STAMP(role, i, k) do { } while (0)
#endif

constexpr int ROW = 160, RANK = 640, CHUNK = 6400;                  // dz2 unit strides, global = shared memory
constexpr int A_CHUNK = 4 * ROW, A_RANK = 4 * A_CHUNK;               // a1 unit, rank-major: 640 / 2 560
constexpr int U_DZ2 = 8 * CHUNK, U_A1 = 10 * A_RANK, U_BYTES = U_DZ2 + U_A1;   // 51 200 + 25 600 = 76 800
constexpr int W_BYTES = 9 * 64 * 32 * 2;                            // 36 864 : [3 dy][8 chunks of co][3 dx][4 groups of ci][8 ci][8 co]
constexpr int W_LBO = 3 * 4 * 128, W_DY = 8 * W_LBO;                // 1 536 between co chunks, 12 288 between kernel rows
constexpr int OFF_W = 2 * U_BYTES;
constexpr int OFF_OUT = OFF_W + W_BYTES;                            // a unit's data gradient, the way it lies in memory: [4 positions][64 squares][32 ci] fp32
constexpr int OUT_BYTES = 4 * 64 * 32 * 4;
constexpr int OFF_BAR = OFF_OUT + OUT_BYTES;
constexpr int SMEM = OFF_BAR + 128;
static_assert(SMEM <= 227 * 1024, "conv2 backward shared memory");
constexpr int THREADS = 192;
constexpr int TM_DG = 320;                      // TMEM columns: weight-gradient taps at 96 dx + 32 dy + ci, data gradient 320 + 96 tile + 32 dx + ci
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
    uint64_t* dg_full = bar_w + 5;       //     data-gradient accumulators of the current unit complete
    uint64_t* dg_empty = bar_w + 7;      //     ... drained by the four epilogue warps
    uint64_t* wg_full = bar_w + 9;       //     every UMMA of this CTA complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 10);
    volatile int* u_slot = reinterpret_cast<volatile int*>(bar_w + 11);   // [2] unit index of buffer b, -1 = no more work
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) CB2_STAMP(0, 14, 0);               // entry

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int b = 0; b < 2; b++) { mbar_init(full + b, 1); mbar_init(empty + b, 5); }
        mbar_init(dg_full, 1);
        mbar_init(dg_empty, 4);
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
    if (tid == 0) CB2_STAMP(0, 15, 0);               // prologue done

    if (warp == 0) {
        // ---------------- producer: weights once, then one unit (half group) per iteration
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, W_BYTES);
            bulk_g2s(smem + OFF_W, w2tp, W_BYTES, bar_w);
        }
        // The unit index is claimed ONE ITERATION AHEAD: the atomic's round trip (about a microsecond with 148 CTAs on one
        // address) used to sit between a buffer becoming free and its fetch, and two buffers do not hide atomic + fetch.
        unsigned u_next = 0;
        if (lane == 0) u_next = atomicAdd(sched, 1u);
        for (long long i = 0;; i++) {
            const int b = (int)(i & 1);
            mbar_wait(empty + b, ((i >> 1) & 1) ^ 1);
            int more = 1;
            if (lane == 0) {
                CB2_STAMP(0, i, 0);
                const unsigned u = u_next;
                if ((long long)u < n_units) {
                    u_slot[b] = (int)u;
#ifdef CB2_NO_FETCH
                    mbar_arrive(full + b);
#else
                    mbar_arrive_expect_tx(full + b, U_BYTES);
                    uint8_t* sb = smem + b * U_BYTES;
                    bulk_g2s(sb, reinterpret_cast<const uint8_t*>(dz2_boards) + (long long)u * U_DZ2, U_DZ2, full + b);
                    bulk_g2s(sb + U_DZ2, reinterpret_cast<const uint8_t*>(a1_boards) + (long long)u * U_A1, U_A1, full + b);
#endif
                    CB2_STAMP(0, i, 1);
                    u_next = atomicAdd(sched, 1u);
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
            constexpr uint32_t ID_DG = make_idesc_f16(128, 96);
            mbar_wait(bar_w, 0);
            // Every operand descriptor is a per-buffer base plus a COMPILE-TIME offset (all loops below are fully
            // unrolled): with run-time descriptor arithmetic the single issuing thread, not the tensor pipe, set
            // the pace (about 50 cycles per UMMA, 232 UMMAs per unit).
            const uint64_t dA_wg0 = make_smem_desc(s0 + 4 * ROW + 16, ROW, CHUNK);             // dz2, MN-major, interior
            const uint64_t dB_wg0 = make_smem_desc(s0 + U_DZ2, ROW, A_CHUNK);                  // a1,  MN-major: rank 0 = dy -1 of the first interior rank, file 0 = dx -1
            const uint64_t dA_dg0 = make_smem_desc(s0, CHUNK, ROW);                            // dz2, K-major
            const uint64_t dB_w = make_smem_desc(s0 + OFF_W, W_LBO, 128);
            for (long long i = 0;; i++) {
                const int b = (int)(i & 1);
                const uint32_t ph = (uint32_t)((i >> 1) & 1);
                mbar_wait(full + b, ph);
                CB2_STAMP(1, i, 0);
                if (u_slot[b] < 0) break;
                tc_fence_after();
                const uint64_t boff = (uint64_t)(b * (U_BYTES >> 4));
                const uint64_t dA_wg = dA_wg0 + boff, dB_wg = dB_wg0 + boff, dA_dg = dA_dg0 + boff;
                const bool acc0 = i != 0;
                // ---- weight gradient: 16 steps of 16 cells (two rows of 8 files, interior row groups 4..35) x 3 kernel
                // columns; B = the a1 rows of ranks R-1, R, R+1 (N = 3 x 32, 640 bytes between 8-channel groups)
#ifndef CB2_NO_WG
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint64_t a_k = dA_wg + (uint64_t)((k * 2 * ROW) >> 4);
                    const uint64_t b_k = dB_wg + (uint64_t)(((k >> 1) * A_RANK + (k & 1) * 2 * ROW) >> 4);
                    umma_f16(tmem, a_k, b_k, ID_WG, k != 0 || acc0);          // the three taps dx: same A, B one file further
                    umma_f16(tmem + 96, a_k, b_k + 1, ID_WG, k != 0 || acc0);
                    umma_f16(tmem + 192, a_k, b_k + 2, ID_WG, k != 0 || acc0);
                }
#endif
                // ---- data gradient: 2 tiles (4 ranks x 4 positions x 8 files) x 3 kernel rows x 4 K-steps of 16 out-channels,
                // N = the row's three taps; the accumulators of the previous unit have to be drained first
                CB2_STAMP(1, i, 1);
                mbar_wait(dg_empty, (uint32_t)(i & 1) ^ 1);
                CB2_STAMP(1, i, 2);
                tc_fence_after();
#ifndef CB2_NO_DG
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int dyi = 0; dyi < 3; dyi++)
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            umma_f16(tmem + TM_DG + 96 * t, dA_dg + (uint64_t)(((4 * t + dyi) * RANK + 16 + j * 2 * CHUNK) >> 4),
                                     dB_w + (uint64_t)((dyi * W_DY + j * 2 * W_LBO) >> 4), ID_DG, (dyi | j) != 0);
#endif
                umma_commit(dg_full);
                umma_commit(empty + b);
                CB2_STAMP(1, i, 3);
            }
            umma_commit(wg_full);
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps: thread = data-gradient tile row (rank in tile, position, file)
        const int q = warp & 3, row = q * 32 + lane;
        const int r_rk = row >> 5, r_pos = (row >> 3) & 3, r_x = row & 7;
        float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // bias gradient: out-channel chunk row / 16, cells = row % 16 mod 16
        long long i = 0;
        for (;; i++) {
            const int b = (int)(i & 1);
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            mbar_wait(full + b, ph);                 // the boards (for the mask and the bias gradient) ...
            if (row == 0) CB2_STAMP(2, i, 0);
            const long long u = u_slot[b];
            if (u < 0) break;
#ifdef CB2_NO_EPI
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);
            mbar_wait(dg_full, (uint32_t)(i & 1));
            tc_fence_after();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(dg_empty);
            continue;
#endif
            // conv1's ReLU mask of this thread's two rows (32 channels each)
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
            // bias gradient: the 32 interior row groups of an out-channel chunk are 320 contiguous cells (the border
            // files are zero); 16 threads per chunk, fp32 sums
#ifndef CB2_NO_BIAS
            {
                const uint4* src = reinterpret_cast<const uint4*>(smem + b * U_BYTES + (row >> 4) * CHUNK + 4 * ROW);
#pragma unroll 4
                for (int c = row & 15; c < 320; c += 16) {
                    const uint4 x = src[c];
                    const unsigned xw[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&xw[e]));
                        bsum[2 * e] += f.x;
                        bsum[2 * e + 1] += f.y;
                    }
                }
            }
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);   // the boards are released
            if (row == 0) CB2_STAMP(2, i, 1);
            mbar_wait(dg_full, (uint32_t)(i & 1));   // ... and the accumulators
            if (row == 0) CB2_STAMP(2, i, 2);
            tc_fence_after();
            // dz1[file] = P_-1[file - 1] + P_0[file] + P_+1[file + 1]: the file neighbours are the neighbouring lanes
            float o[2][32];
#pragma unroll
            for (int t = 0; t < 2; t++) {
                uint32_t v[3][32];
#pragma unroll
                for (int g = 0; g < 3; g++) tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_DG + 96 * t + 32 * g, v[g]);
                tmem_ld_wait();
                if (t == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dg_empty);
                    if (row == 0) CB2_STAMP(2, i, 3);
                }
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const float lo = __uint_as_float(__shfl_up_sync(0xffffffffu, v[0][c], 1));
                    const float hi = __uint_as_float(__shfl_down_sync(0xffffffffu, v[2][c], 1));
                    o[t][c] = __uint_as_float(v[1][c]) + (r_x > 0 ? lo : 0.f) + (r_x < 7 ? hi : 0.f);
                }
            }
            // ReLU' and the scale, then the thread's two rows of 128 bytes go to the staging tile.  The rows of a quarter
            // warp are the eight files of one board row (1 KB apart in banks = same banks), so step k writes the row's
            // 16-byte chunk k ^ file: the chunks are brought into that order by three conditional swaps on the file bits.
            // The unit's 32 KB are contiguous in dz1c: ONE bulk store per unit instead of 2 048 16-byte stores with 32
            // lines per instruction (which occupied the load/store pipe for a third of the unit's time).
            if (row == 0) CB2_STAMP(2, i, 4);
            if (row == 0) bulk_wait_read0();                             // the previous unit's store has read the tile
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const unsigned bits = relu_mask[t];
                float c[32];
#pragma unroll
                for (int e = 0; e < 32; e++) c[e] = ((bits >> e) & 1u) ? o[t][e] * inv_scale : 0.f;
#pragma unroll
                for (int m = 1; m <= 4; m <<= 1) {
                    const bool sw = (r_x & m) != 0;
                    float d[32];
#pragma unroll
                    for (int e = 0; e < 32; e++) d[e] = sw ? c[e ^ (4 * m)] : c[e];
#pragma unroll
                    for (int e = 0; e < 32; e++) c[e] = d[e];
                }
                uint8_t* orow = smem + OFF_OUT + ((r_pos * 64 + (4 * t + r_rk) * 8 + r_x) << 7);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    *reinterpret_cast<float4*>(orow + ((k ^ r_x) << 4)) = make_float4(c[4 * k], c[4 * k + 1], c[4 * k + 2], c[4 * k + 3]);
            }
            fence_proxy_async();
            asm volatile("bar.sync 2, 128;" ::: "memory");
#ifndef CB2_NO_STORE
            if (row == 0) {
                const long long left = n - u * 4;                        // positions of this unit that exist
                if (left > 0) bulk_s2g(dz1c + u * (4 * 64 * 32), smem + OFF_OUT, (uint32_t)(left < 4 ? left : 4) * (64 * 32 * 4));
                CB2_STAMP(2, i, 5);
            }
#endif
        }
        if (row == 0) bulk_wait0();
        // ---- weight/bias gradient: rows co = 16 * quadrant + lane (lanes 0..15 hold data), transposed through shared
        // memory (unit buffer 0 is dead: every UMMA is complete and the producer has nothing in flight) into PyTorch
        // order [co][ci][tap] and added with coalesced 16-byte reductions
        if (i > 0) {
            if (row == 0) CB2_STAMP(2, 14, 0);
            mbar_wait(wg_full, 0);
            if (row == 0) CB2_STAMP(2, 14, 1);
            tc_fence_after();
            asm volatile("bar.sync 1, 128;" ::: "memory");              // nobody still reads the last unit's boards
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
#pragma unroll
            for (int e = 0; e < 8; e++) {
#pragma unroll
                for (int d = 8; d >= 1; d >>= 1) bsum[e] += __shfl_xor_sync(0xffffffffu, bsum[e], d);
            }
            if ((lane & 15) == 0) {
#pragma unroll
                for (int e = 0; e < 8; e++) atomicAdd(g_b2 + (row >> 4) * 8 + e, bsum[e] * inv_scale);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");              // the four epilogue warps
            for (int j = row; j < 64 * 72; j += 128) {
                const int r = j / 72, c4 = j - r * 72;
                const float4 t4 = *reinterpret_cast<const float4*>(tile + r * WT_STRIDE + 4 * c4);
#ifndef CB2_NO_ATOM
                atomicAdd(reinterpret_cast<float4*>(g_w2) + j, t4);     // conv2.weight[co][ci][tap], 288 floats per co
#else
                if (t4.x == 12345.f) g_w2[j] = 0.f;
#endif
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) CB2_STAMP(0, 15, 1);               // every role done
    if (warp == 1) tmem_dealloc(tmem, 512);
    if (tid == 0 && atomicAdd(sched + 1, 1u) == gridDim.x - 1) {     // last CTA out: every producer has seen the end
        sched[0] = 0;
        sched[1] = 0;
    }
}

} // namespace cb2

#ifdef CB2_TRACE
extern "C" int cdq_debug_cb2_trace(long long* host_out /*[384]*/) {
    return (int)cudaMemcpyFromSymbol(host_out, cb2::g_cb2_trace, sizeof(long long) * 3 * 16 * 8);
}
#endif

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
