// conv_bwd_kernel.cu -- backward of conv2 (32 -> 64, 3x3) on tcgen05, with the same
// boards-in-shared-memory operands as the forward (conv_kernel.cu): per group of 8 positions
//   dz2 boards  [8 chunks of 8 out-channels][10 ranks][8 positions][10 files][8 halfs]  (gradient at
//               conv2's pre-activation, fp16, scaled; written by fc1_dgrad_kernel)
//   a1  boards  [4 chunks of 8 in-channels ][10 ranks][8 positions][10 files][8 halfs]  (ReLU(conv1),
//               written by the forward's epilogue 1)
// both with zero borders, so every 3x3 tap is an operand start address.
//
//   weight gradient  dW2[co][ci][tap] += sum_cells dz2[cell][co] * a1[cell + tap][ci]
//       one UMMA per (tap, 16 cells): M = 64 (co), N = 32 (ci), K = 16 cells; both operands MN-major
//       (8 channels = the 16-byte vector, cells advance by 16 B, 160 B between board rows);
//       accumulators for all 9 taps stay in TMEM over every group the CTA processes.
//   bias gradient    db2[co] += sum_cells dz2[cell][co]  -- the same A against a tile of ones (N = 8)
//   data gradient    dz1[cell][ci] = ReLU'(a1) * sum_{co,tap} dz2[cell - tap][co] * W2[co][ci][tap]
//       the forward's implicit GEMM with K = 64, N = 32 and flipped taps: M = 128 rows per UMMA.
// M = 64 accumulators live in lanes 0..15 of each 32-lane TMEM quadrant (tests/cuda/umma_m64_probe.cu).
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {

constexpr int CB_ROW = 160, CB_RANK = 1280, CB_CHUNK = 12800;
constexpr int CB_DZ2_BYTES = 8 * CB_CHUNK;      // 102 400
constexpr int CB_A1_BYTES = 4 * CB_CHUNK;       //  51 200
constexpr int CB_W_BYTES = 9 * 64 * 32 * 2;     //  36 864 : [9 taps][8 chunks of co][4 groups of ci][8 ci][8 co]
constexpr int CB_OFF_A1 = CB_DZ2_BYTES;
constexpr int CB_OFF_W = CB_OFF_A1 + CB_A1_BYTES;
constexpr int CB_OFF_ONES = CB_OFF_W + CB_W_BYTES;
constexpr int CB_OFF_BAR = CB_OFF_ONES + 256;
constexpr int CB_SMEM = CB_OFF_BAR + 64;
static_assert(CB_SMEM <= 227 * 1024, "conv2 backward shared memory");
constexpr int CB_THREADS = 128;
constexpr int CB_WT_STRIDE = 292;                // floats per co row of the staged weight-gradient tile (288 + 4: bank spread)
static_assert(64 * CB_WT_STRIDE * 4 <= CB_DZ2_BYTES, "weight-gradient tile is staged over the dz2 boards");
constexpr int CB_TM_BIAS = 288, CB_TM_DG = 320;  // TMEM columns: taps 0..287, bias 288..295, data-gradient tiles 320..447

__global__ void __launch_bounds__(CB_THREADS, 1)
conv2_bwd_kernel(const __half* __restrict__ dz2_boards, const __half* __restrict__ a1_boards,
                 const __half* __restrict__ w2tp, long long n_groups, long long n, float inv_scale,
                 float* __restrict__ g_w2, float* __restrict__ g_b2, float* __restrict__ dz1c) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + CB_OFF_BAR);
    uint64_t* bar_load = bar_w + 1;
    uint64_t* bar_mma = bar_w + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < 128) reinterpret_cast<uint16_t*>(smem + CB_OFF_ONES)[tid] = 0x3C00;   // fp16 1.0
    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_load, 1); mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_w, CB_W_BYTES);
        bulk_g2s(smem + CB_OFF_W, w2tp, CB_W_BYTES, bar_w);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    mbar_wait(bar_w, 0);

    const uint32_t s0 = smem_u32(smem);
    constexpr uint32_t ID_WG = make_idesc_f16_ex(64, 32, true, true);
    constexpr uint32_t ID_BIAS = make_idesc_f16_ex(64, 8, true, true);
    constexpr uint32_t ID_DG = make_idesc_f16(128, 32);
    const int row = tid;                                   // data-gradient tile row: (yy, position, file)
    const int r_yy = row >> 6, r_pos = (row >> 3) & 7, r_x = row & 7;

    // The boards of group i+1 are fetched while the data-gradient epilogue of group i runs: the epilogue's only
    // use of the boards -- conv1's ReLU mask -- is pulled into registers first (128 bits per thread).
    auto fetch = [&](long long g) {
        mbar_arrive_expect_tx(bar_load, CB_DZ2_BYTES + CB_A1_BYTES);
        bulk_g2s(smem, reinterpret_cast<const uint8_t*>(dz2_boards) + g * CB_DZ2_BYTES, CB_DZ2_BYTES, bar_load);
        bulk_g2s(smem + CB_OFF_A1, reinterpret_cast<const uint8_t*>(a1_boards) + g * CB_A1_BYTES, CB_A1_BYTES, bar_load);
    };
    if (tid == 0 && (long long)blockIdx.x < n_groups) fetch(blockIdx.x);
    uint32_t iter = 0;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, iter++) {
        mbar_wait(bar_load, iter & 1);
        if (warp == 0) {
            if (elect_one_sync()) {
                tc_fence_after();
                // ---- weight gradient: 9 taps x 32 steps of 16 cells (two board rows); interior rows 8..71
                for (int tap = 0; tap < 9; tap++) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    const uint32_t a_row0 = s0 + 8 * CB_ROW + 16;
                    const uint32_t b_row0 = s0 + CB_OFF_A1 + 8 * CB_ROW + 16 + dy * CB_RANK + dx * 16;
#pragma unroll 4
                    for (int k = 0; k < 32; k++)
                        umma_f16(tmem + tap * 32, make_smem_desc(a_row0 + k * 2 * CB_ROW, CB_ROW, CB_CHUNK),
                                 make_smem_desc(b_row0 + k * 2 * CB_ROW, CB_ROW, CB_CHUNK), ID_WG, (iter | k) != 0);
                }
                // ---- bias gradient: same A against ones
#pragma unroll 4
                for (int k = 0; k < 32; k++)
                    umma_f16(tmem + CB_TM_BIAS, make_smem_desc(s0 + 8 * CB_ROW + 16 + k * 2 * CB_ROW, CB_ROW, CB_CHUNK),
                             make_smem_desc(s0 + CB_OFF_ONES, 128, 128), ID_BIAS, (iter | k) != 0);
                // ---- data gradient: 4 tiles x 9 flipped taps x 4 K-steps of 16 out-channels
                for (int t = 0; t < 4; t++)
                    for (int u = 0; u < 9; u++) {
                        const int dy = u / 3 - 1, dx = u % 3 - 1;
                        const uint32_t a = s0 + (2 * t + dy + 1) * CB_RANK + (dx + 1) * 16;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            umma_f16(tmem + CB_TM_DG + 32 * t, make_smem_desc(a + j * 2 * CB_CHUNK, CB_CHUNK, CB_ROW),
                                     make_smem_desc(s0 + CB_OFF_W + u * 4096 + j * 2 * 512, 512, 128), ID_DG, (u | j) != 0);
                    }
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, iter & 1);
        tc_fence_after();
        // ---- ReLU mask of this thread's four rows (32 channels each) from the a1 boards, then release the boards
        unsigned relu_mask[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int y = 2 * t + r_yy;
            const uint8_t* m = smem + CB_OFF_A1 + (y + 1) * CB_RANK + r_pos * CB_ROW + (r_x + 1) * 16;
            unsigned bits = 0;
#pragma unroll
            for (int kc = 0; kc < 4; kc++) {
                const uint4 mk = *reinterpret_cast<const uint4*>(m + kc * CB_CHUNK);
                const unsigned mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
                for (int e = 0; e < 8; e++)
                    bits |= (((mw[e >> 1] >> (16 * (e & 1))) & 0xffffu) ? 1u : 0u) << (kc * 8 + e);
            }
            relu_mask[t] = bits;
        }
        __syncthreads();   // every thread has its mask: the boards may be overwritten
        if (tid == 0 && g + gridDim.x < n_groups) {
            fence_proxy_async();
            fetch(g + gridDim.x);
        }
        // ---- data-gradient epilogue: mask, unscale, fp32 [position][square][32].  The four tiles are read from
        // TMEM two at a time: one tcgen05.ld latency per pair instead of per tile.
        const long long pos = g * 8 + r_pos;
#pragma unroll
        for (int tp = 0; tp < 4; tp += 2) {
            uint32_t v[2][32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + CB_TM_DG + 32 * tp, v[0]);
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + CB_TM_DG + 32 * (tp + 1), v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int t = tp + h;
                const int y = 2 * t + r_yy;
                const unsigned bits = relu_mask[t];
                if (pos < n) {
                    float4* o = reinterpret_cast<float4*>(dz1c + (pos * 64 + y * 8 + r_x) * 32);
#pragma unroll
                    for (int q4 = 0; q4 < 8; q4++) {
                        float r[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            r[e] = ((bits >> (q4 * 4 + e)) & 1u) ? __uint_as_float(v[h][q4 * 4 + e]) * inv_scale : 0.f;
                        o[q4] = make_float4(r[0], r[1], r[2], r[3]);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();   // the data-gradient accumulators are reused by the next group
    }

    // ---- weight/bias gradient epilogue: rows co = 16*quadrant + lane (lanes 0..15 hold data).  The tile is
    // transposed through shared memory (the boards are dead by now) into PyTorch order [co][ci][tap] and added
    // with coalesced 16-byte reductions: one atomic per element straight from the accumulator rows put 32
    // transactions of every CTA on each 128-byte line of the gradient, and same-line atomics serialise in L2
    // (about 20 us of this kernel at 148 CTAs).
    if (iter > 0) {
        tc_fence_after();
        float* tile = reinterpret_cast<float*>(smem);              // [64 co][CB_WT_STRIDE]
        const int co = warp * 16 + lane;
#pragma unroll 1
        for (int tap = 0; tap < 9; tap++) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + tap * 32, v);
            tmem_ld_wait();
            if (lane < 16) {
#pragma unroll
                for (int ci = 0; ci < 32; ci++) tile[co * CB_WT_STRIDE + ci * 9 + tap] = __uint_as_float(v[ci]) * inv_scale;
            }
        }
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + CB_TM_BIAS, v);   // columns 288..319; the first 8 are the bias sums
        tmem_ld_wait();
        if (lane < 16) atomicAdd(g_b2 + co, __uint_as_float(v[0]) * inv_scale);
        __syncthreads();
        for (int i = tid; i < 64 * 72; i += CB_THREADS) {
            const int r = i / 72, q = i - r * 72;
            const float4 t4 = *reinterpret_cast<const float4*>(tile + r * CB_WT_STRIDE + 4 * q);
            atomicAdd(reinterpret_cast<float4*>(g_w2) + i, t4);     // conv2.weight[co][ci][tap], 288 floats per co
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// first formulation (whole groups, one buffer); CDQ_CONV2_BWD_V1=1 selects it in launch_conv2_bwd (conv_bwd2_kernel.cu)
int launch_conv2_bwd_v1(const __half* dz2_boards, const __half* a1_boards, const __half* w2tp, long long n, float inv_scale,
                     float* g_w2, float* g_b2, float* dz1c, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM);
        attr_set = true;
    }
    const long long groups = ((n + 127) >> 7) * 16;
    const int sms = device_sm_count();
    ProfileScope ps(KK_CONV2_BWD, st);
    conv2_bwd_kernel<<<(unsigned)(groups < sms ? groups : sms), CB_THREADS, CB_SMEM, st>>>(
        dz2_boards, a1_boards, w2tp, groups, n, inv_scale, g_w2, g_b2, dz1c);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
