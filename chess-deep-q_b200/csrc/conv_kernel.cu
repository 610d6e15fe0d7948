// conv_kernel.cu -- K2a: boards -> ReLU(conv2(ReLU(conv1(one-hot planes)))) as fp16 in fc1's
// A-operand tile order (neural_network.py:22-23,30-31).
//
// Implicit GEMM on tcgen05 with no im2col: activations sit in shared memory as zero-bordered
// boards, channel-chunked and rank-major,
//        [K/8 chunks][10 padded ranks][8 positions][10 padded files][8 halfs = 16 B],
// so that a K-major no-swizzle UMMA operand with SBO = 160 B (one padded rank row of one position)
// and LBO = one chunk covers M = 128 rows = 2 ranks x 8 positions x 8 files, and a 3x3 tap (dy,dx)
// is just a different START ADDRESS of the same operand (tests/cuda/umma_probe.cu checks that the
// hardware accepts 16-byte-aligned starts and this SBO).  conv1 reads the one-hot board (12 piece
// channels + a constant-one channel that carries its bias), conv2 reads ReLU(conv1) written back
// to shared memory by an epilogue.
//
// Persistent CTA, one per SM, nine warps, everything double-buffered and handed over by mbarriers:
//   warp 0      MMA issuer:  c1(0), then per group  c1(i+1) ; c2(i)   (36 + 72 UMMAs)
//   warps 1-4   encode planes(i+1) from the packed boards, then epilogue 1 of group i:
//               TMEM -> ReLU -> fp16 -> act1 boards in shared memory
//   warps 5-8   epilogue 2 of group i: TMEM -> +bias -> ReLU -> fp16 -> global (fc1 tile order)
// TMEM: conv1 accumulators 4 tiles x 32 columns, conv2 accumulators 4 tiles x 64 columns.
//
// Algorithmic work: 2 801 664 FLOP/position (dense convention);
// traffic 72 B in + 8 192 B out per position.
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {

typedef unsigned long long u64;

constexpr int GROUP = 8;                          // positions per group
constexpr int CELL = 16;                          // bytes per (square, 8-channel chunk)
constexpr int ROW_BYTES = 10 * CELL;              // one padded rank of one position: 160 B (= SBO)
constexpr int RANK_BYTES = GROUP * ROW_BYTES;     // 1 280 B
constexpr int CHUNK_BYTES = 10 * RANK_BYTES;      // 12 800 B (= LBO)
constexpr int PLANES_BYTES = 2 * CHUNK_BYTES;     // 16 input channels
constexpr int ACT1_BYTES = 4 * CHUNK_BYTES;       // 32 channels
constexpr int W1_TAP_BYTES = 32 * 16 * 2;
constexpr int W2_TAP_BYTES = 64 * 32 * 2;
constexpr int W1_BYTES = 9 * W1_TAP_BYTES;
constexpr int W2_BYTES = 9 * W2_TAP_BYTES;

constexpr int OFF_PLANES = 0;                               // 2 buffers
constexpr int OFF_ACT1 = OFF_PLANES + 2 * PLANES_BYTES;     // 2 buffers
constexpr int OFF_W1 = OFF_ACT1 + 2 * ACT1_BYTES;
constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
constexpr int OFF_B2 = OFF_W2 + W2_BYTES;                   // conv2 bias, 64 floats
constexpr int OFF_REC = OFF_B2 + 256;                       // staged records of one group (576 B)
constexpr int OFF_BAR = OFF_REC + 576;
constexpr int CONV_SMEM = OFF_BAR + 32 * 8;
static_assert(CONV_SMEM <= 227 * 1024, "conv kernel shared memory");
static_assert(OFF_BAR % 8 == 0 && OFF_W1 % 16 == 0, "alignment");

constexpr int CONV_THREADS = 9 * 32;
constexpr int TM_C1 = 0;      // TMEM column of conv1 tile t: TM_C1 + 32 t
constexpr int TM_C2 = 128;    // TMEM column of conv2 tile t: TM_C2 + 64 t

// mbarrier indices
enum { B_W = 0, B_PL_FULL = 1, B_PL_EMPTY = 3, B_C1_FULL = 5, B_C1_EMPTY = 9, B_A1_FULL = 13, B_A1_EMPTY = 15,
       B_C2_FULL = 17, B_C2_EMPTY = 21, B_COUNT = 25 };

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_stack_kernel(const CdqBoard* __restrict__ boards, long long n, const PackedNet net,
                  __half* __restrict__ act2, __half* __restrict__ dbg_act1, __half* __restrict__ act1_boards) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup
    for (int i = tid; i < (2 * PLANES_BYTES + 2 * ACT1_BYTES) / 16; i += CONV_THREADS)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);   // borders stay zero forever
    if (tid < 64) reinterpret_cast<float*>(smem + OFF_B2)[tid] = net.conv2_b[tid];
    if (tid == 0) {
        mbar_init(bars + B_W, 1);
        for (int b = 0; b < 2; b++) {
            mbar_init(bars + B_PL_FULL + b, 4);  mbar_init(bars + B_PL_EMPTY + b, 1);
            mbar_init(bars + B_A1_FULL + b, 4);  mbar_init(bars + B_A1_EMPTY + b, 1);
        }
        for (int t = 0; t < 4; t++) {
            mbar_init(bars + B_C1_FULL + t, 1);  mbar_init(bars + B_C1_EMPTY + t, 4);
            mbar_init(bars + B_C2_FULL + t, 1);  mbar_init(bars + B_C2_EMPTY + t, 4);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bars + B_W, W1_BYTES + W2_BYTES);
        bulk_g2s(smem + OFF_W1, net.w1p, W1_BYTES, bars + B_W);
        bulk_g2s(smem + OFF_W2, net.w2p, W2_BYTES, bars + B_W);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // whole 128-position fc1 tiles are always written (missing boards count as empty boards), so
    // every row of the act2 buffer holds finite values: the training kernels contract over them
    const long long n_groups = ((n + 127) >> 7) * (128 / GROUP);
    // groups of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int n_local = (int)((n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (warp == 0) {
        // =============================== MMA issuer (warp-uniform control, lane 0 issues)
        mbar_wait_warp(lane, bars + B_W, 0);
        const uint32_t smem_a = smem_u32(smem);
        constexpr uint32_t IDESC1 = make_idesc_f16(128, 32);
        constexpr uint32_t IDESC2 = make_idesc_f16(128, 64);
        constexpr uint64_t A_HI = ((uint64_t)(ROW_BYTES >> 4) << 32) | (1ull << 46);   // SBO, version
        constexpr uint64_t A_LBO = (uint64_t)(CHUNK_BYTES >> 4) << 16;
        constexpr uint64_t B_HI = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint64_t w1_desc = B_HI | ((uint64_t)(512 >> 4) << 16) | (uint64_t)(((smem_a + OFF_W1) >> 4) & 0x3fff);
        const uint64_t w2_desc = B_HI | ((uint64_t)(1024 >> 4) << 16) | (uint64_t)(((smem_a + OFF_W2) >> 4) & 0x3fff);

        auto issue_c1 = [&](int j) {
            const int buf = j & 1;
            mbar_wait_warp(lane, bars + B_PL_FULL + buf, (j >> 1) & 1);
            tc_fence_after();
            const uint64_t a_base = A_HI | A_LBO | (uint64_t)(((smem_a + OFF_PLANES + buf * PLANES_BYTES) >> 4) & 0x3fff);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                mbar_wait_warp(lane, bars + B_C1_EMPTY + t, (j & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        const int off = (2 * t + dy + 1) * RANK_BYTES + (dx + 1) * CELL;
                        umma_f16(tmem + TM_C1 + 32 * t, a_base + (uint64_t)(off >> 4),
                                 w1_desc + (uint64_t)((tap * W1_TAP_BYTES) >> 4), IDESC1, tap > 0);
                    }
                    umma_commit(bars + B_C1_FULL + t);
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(bars + B_PL_EMPTY + buf);
            __syncwarp();
        };
        auto issue_c2 = [&](int j) {
            const int buf = j & 1;
            mbar_wait_warp(lane, bars + B_A1_FULL + buf, (j >> 1) & 1);
            tc_fence_after();
            const uint64_t a_base = A_HI | A_LBO | (uint64_t)(((smem_a + OFF_ACT1 + buf * ACT1_BYTES) >> 4) & 0x3fff);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                mbar_wait_warp(lane, bars + B_C2_EMPTY + t, (j & 1) ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
#pragma unroll
                    for (int tap = 0; tap < 9; tap++) {
                        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                        const int off = (2 * t + dy + 1) * RANK_BYTES + (dx + 1) * CELL;
#pragma unroll
                        for (int k = 0; k < 2; k++)
                            umma_f16(tmem + TM_C2 + 64 * t, a_base + (uint64_t)((off + k * 2 * CHUNK_BYTES) >> 4),
                                     w2_desc + (uint64_t)((tap * W2_TAP_BYTES + k * 2 * 1024) >> 4), IDESC2, (tap | k) > 0);
                    }
                    umma_commit(bars + B_C2_FULL + t);
                }
                __syncwarp();
            }
            if (elect_one_sync()) umma_commit(bars + B_A1_EMPTY + buf);
            __syncwarp();
        };
        if (n_local > 0) issue_c1(0);
        for (int i = 0; i < n_local; i++) {
            if (i + 1 < n_local) issue_c1(i + 1);
            issue_c2(i);
        }
    } else if (warp <= 4) {
        // =============================== encoder + epilogue 1 (warps 1..4, 128 threads)
        const int e = tid - 32;                 // 0..127
        const int q = warp & 3;                 // TMEM lane quarter of this warp
        const int row = q * 32 + lane;          // tile row: yy = row/64, position = (row/8)%8, file = row%8
        const int r_yy = row >> 6, r_pos = (row >> 3) & 7, r_x = row & 7;
        u64* srec = reinterpret_cast<u64*>(smem + OFF_REC);
        const u64* gboards = reinterpret_cast<const u64*>(boards);
        // record word prefetch: threads 0..71 each own one u64 of the next group's 8 records
        auto fetch = [&](int j) -> u64 {
            if (e >= 72 || j >= n_local) return 0ull;
            const long long g = blockIdx.x + (long long)j * gridDim.x;
            const long long word = g * (GROUP * 9) + e;
            return word < n * 9 ? __ldg(gboards + word) : 0ull;
        };
        u64 next_word = fetch(0);

        auto encode = [&](int j) {
            const int buf = j & 1;
            mbar_wait_warp(lane, bars + B_PL_EMPTY + buf, ((j >> 1) & 1) ^ 1);
            named_bar_sync(1, 128);             // previous encode finished reading srec
            if (e < 72) srec[e] = next_word;
            named_bar_sync(1, 128);
            next_word = fetch(j + 1);           // in flight during this encode and the next epilogue
            // thread = (half-board e/64, position (e/8)%8, file e%8) -> 4 ranks x 32 B one-hot.  A
            // quarter-warp (8 files of one position) writes 128 contiguous bytes: no bank conflicts.
            const int p = (e >> 3) & 7, y0 = (e >> 6) * 4, xf = e & 7;
            const u64* r = srec + p * 9;
            const int sh = 8 * y0 + xf;
            unsigned tb[6];       // bit 8*k = rank y0+k of this file
#pragma unroll
            for (int t = 0; t < 6; t++) tb[t] = (unsigned)(r[t] >> sh) & 0x01010101u;
            const unsigned wb = (unsigned)(r[6] >> sh) & 0x01010101u, bb = (unsigned)(r[7] >> sh) & 0x01010101u;
            uint8_t* dst0 = smem + OFF_PLANES + buf * PLANES_BYTES + (y0 + 1) * RANK_BYTES + p * ROW_BYTES + (xf + 1) * CELL;
#pragma unroll
            for (int x = 0; x < 4; x++) {   // x = rank offset; bit position 8*x
                int t = -1;
#pragma unroll
                for (int k = 0; k < 6; k++) if ((tb[k] >> (8 * x)) & 1) t = k;
                const bool isw = (wb >> (8 * x)) & 1, isb = (bb >> (8 * x)) & 1;
                const int ch = (t >= 0 && (isw || isb)) ? (isw ? t : t + 6) : -1;   // board_utils.py:17-30
                u64 lo0 = 0, hi0 = 0, lo1 = 0;
                const u64 one = 0x3C00ull;                                          // fp16 1.0
                if (ch >= 0 && ch < 4) lo0 = one << (16 * ch);
                else if (ch >= 4 && ch < 8) hi0 = one << (16 * (ch - 4));
                else if (ch >= 8) lo1 = one << (16 * (ch - 8));
                const u64 hi1 = one;            // channel 12: constant one (carries conv1's bias)
                *reinterpret_cast<uint4*>(dst0 + x * RANK_BYTES) =
                    make_uint4((unsigned)lo0, (unsigned)(lo0 >> 32), (unsigned)hi0, (unsigned)(hi0 >> 32));
                *reinterpret_cast<uint4*>(dst0 + CHUNK_BYTES + x * RANK_BYTES) =
                    make_uint4((unsigned)lo1, (unsigned)(lo1 >> 32), (unsigned)hi1, (unsigned)(hi1 >> 32));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + B_PL_FULL + buf);
        };
        auto epilogue1 = [&](int j) {
            const int buf = j & 1;
            mbar_wait_warp(lane, bars + B_A1_EMPTY + buf, ((j >> 1) & 1) ^ 1);   // conv2 of group j-2 has read this buffer
            const long long pos = (blockIdx.x + (long long)j * gridDim.x) * GROUP + r_pos;
#pragma unroll 1
            for (int t = 0; t < 4; t++) {
                mbar_wait_warp(lane, bars + B_C1_FULL + t, j & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + TM_C1 + 32 * t, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + B_C1_EMPTY + t);
                const int y = 2 * t + r_yy;
                uint8_t* dst = smem + OFF_ACT1 + buf * ACT1_BYTES + (y + 1) * RANK_BYTES + r_pos * ROW_BYTES + (r_x + 1) * CELL;
#pragma unroll
                for (int kc = 0; kc < 4; kc++) {
                    uint4 o;
                    o.x = relu_pack_f16(__uint_as_float(v[kc * 8 + 0]), __uint_as_float(v[kc * 8 + 1]));
                    o.y = relu_pack_f16(__uint_as_float(v[kc * 8 + 2]), __uint_as_float(v[kc * 8 + 3]));
                    o.z = relu_pack_f16(__uint_as_float(v[kc * 8 + 4]), __uint_as_float(v[kc * 8 + 5]));
                    o.w = relu_pack_f16(__uint_as_float(v[kc * 8 + 6]), __uint_as_float(v[kc * 8 + 7]));
                    *reinterpret_cast<uint4*>(dst + kc * CHUNK_BYTES) = o;
                    if (act1_boards)   // training: keep ReLU(conv1), position-contiguous [tile of 16][file][chunk][rank][16 pos][8 halfs]
                        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(act1_boards) + (pos >> 4) * 65536 + r_x * 8192 + kc * 2048 +
                                                  y * 256 + (int)(pos & 15) * CELL) = o;
                    if (dbg_act1 && pos < n)
                        *reinterpret_cast<uint4*>(dbg_act1 + (pos * 64 + y * 8 + r_x) * 32 + kc * 8) = o;
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + B_A1_FULL + buf);
        };
        if (n_local > 0) encode(0);
        for (int i = 0; i < n_local; i++) {
            if (i + 1 < n_local) encode(i + 1);
            epilogue1(i);
        }
    } else {
        // =============================== epilogue 2 (warps 5..8)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int r_yy = row >> 6, r_pos = (row >> 3) & 7, r_x = row & 7;
        float4 b2r[16];   // conv2 bias, kept in registers for the whole kernel
#pragma unroll
        for (int k = 0; k < 16; k++) b2r[k] = reinterpret_cast<const float4*>(smem + OFF_B2)[k];
        for (int i = 0; i < n_local; i++) {
            const long long pos = (blockIdx.x + (long long)i * gridDim.x) * GROUP + r_pos;
            const int pr = (int)(pos & 127);
            // fc1 A-tile order: 16 KB block per (128-position tile, square):
            // [8 channel chunks][16 row groups][8 rows][8 halfs]
            uint8_t* base = reinterpret_cast<uint8_t*>(act2) + ((pos >> 7) << 20) + (pr >> 3) * 128 + (pr & 7) * 16;
#pragma unroll 1
            for (int t = 0; t < 4; t++) {
                mbar_wait_warp(lane, bars + B_C2_FULL + t, i & 1);
                tc_fence_after();
                uint32_t v0[32], v1[32];
                const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + TM_C2 + 64 * t;
                tmem_ld32(ta, v0);
                tmem_ld32(ta + 32, v1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + B_C2_EMPTY + t);
                const int sq = (2 * t + r_yy) * 8 + r_x;
                uint8_t* dst = base + ((long long)sq << 14);
#pragma unroll
                for (int kc = 0; kc < 8; kc++) {
                    const uint32_t* v = kc < 4 ? v0 + kc * 8 : v1 + (kc - 4) * 8;
                    const float4 ba = b2r[kc * 2], bb = b2r[kc * 2 + 1];
                    uint4 o;
                    o.x = relu_pack_f16(__uint_as_float(v[0]) + ba.x, __uint_as_float(v[1]) + ba.y);
                    o.y = relu_pack_f16(__uint_as_float(v[2]) + ba.z, __uint_as_float(v[3]) + ba.w);
                    o.z = relu_pack_f16(__uint_as_float(v[4]) + bb.x, __uint_as_float(v[5]) + bb.y);
                    o.w = relu_pack_f16(__uint_as_float(v[6]) + bb.z, __uint_as_float(v[7]) + bb.w);
                    *reinterpret_cast<uint4*>(dst + kc * 2048) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// first formulation (one MMA per 3x3 tap); kept for A/B runs: CDQ_CONV_V1=1 selects it in launch_conv_stack
int launch_conv_stack_v1(const CdqBoard* boards, long long n, const PackedNet& net, __half* act2, __half* dbg_act1,
                         cudaStream_t st, __half* act1_boards) {
    if (n == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM);
        attr_set = true;
    }
    const long long groups = ((n + 127) >> 7) * (128 / GROUP);
    const int sms = device_sm_count();
    const unsigned grid = (unsigned)(groups < sms ? groups : sms);
    ProfileScope ps(KK_CONV, st);
    conv_stack_kernel<<<grid, CONV_THREADS, CONV_SMEM, st>>>(boards, n, net, act2, dbg_act1, act1_boards);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
