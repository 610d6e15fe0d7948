// cnn_kernels.cu -- K2: ChessQNetwork.forward (neural_network.py:29-37) on packed boards.
//
// Two tcgen05 kernels, fp16 operands / fp32 accumulation in TMEM:
//
//  K2a conv_stack_kernel   boards -> act2 (fp16, fc1's A-operand tile order)
//      A persistent CTA handles groups of 16 positions.  One UMMA tile is M=128 rows =
//      16 positions x 8 files of ONE output rank y.  Activations live in shared memory as
//      zero-bordered 10x10 boards, channel-chunked ([K/8][position][10][10][8 halfs]); a 3x3 tap
//      (dy,dx) is then nothing but a different START ADDRESS of the same K-major no-swizzle
//      operand (SBO = one padded board, LBO = one channel chunk), so the implicit GEMM needs no
//      im2col copy at all.  conv1's input is the one-hot board (12 piece channels + a constant-one
//      channel that carries the bias), conv2's input is ReLU(conv1) written back by the epilogue.
//  K2b fc_head_kernel      act2 -> value = tanh(fc2(relu(fc1)))   (+ optional leaf post-processing)
//      Warp-specialised GEMM, M=128 positions x N=256 x K=4096, operands fetched by 1-D bulk async
//      copies (TMA engine) from pre-tiled global buffers through a 4-stage mbarrier ring; the
//      epilogue applies bias+ReLU and the 256->1 fc2 dot product per TMEM lane (= per position).
//
// Algorithmic work: 4 899 328 FLOP/position; act2 costs 8 KB/position of
// HBM/L2 traffic each way between the two kernels.
#include "umma.cuh"
#include "bitboard.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {

// ------------------------------------------------------------------------------------------
// shared-memory geometry of the conv kernel
constexpr int GROUP = 16;                   // positions per CTA iteration (M = GROUP * 8 = 128)
constexpr int CELL = 16;                    // bytes per (square, 8-channel chunk)
constexpr int BOARD_BYTES = 10 * 10 * CELL; // one zero-bordered board of one chunk: 1600 B  (= SBO)
constexpr int CHUNK_BYTES = GROUP * BOARD_BYTES; // 25 600 B (= LBO)
constexpr int PLANES_BYTES = 2 * CHUNK_BYTES;    // 16 input channels
constexpr int ACT1_BYTES = 4 * CHUNK_BYTES;      // 32 channels
constexpr int W1_TAP_BYTES = 32 * 16 * 2;        // N=32 x K=16 halfs
constexpr int W2_TAP_BYTES = 64 * 32 * 2;        // N=64 x K=32 halfs
constexpr int W1_BYTES = 9 * W1_TAP_BYTES;
constexpr int W2_BYTES = 9 * W2_TAP_BYTES;

constexpr int OFF_PLANES = 0;
constexpr int OFF_ACT1 = OFF_PLANES + PLANES_BYTES;
constexpr int OFF_W1 = OFF_ACT1 + ACT1_BYTES;
constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
constexpr int OFF_B2 = OFF_W2 + W2_BYTES;          // conv2 bias, 64 floats
constexpr int OFF_REC = OFF_B2 + 256;              // 16 staged board records (1152 B)
constexpr int OFF_BAR = OFF_REC + 1152;            // mbarriers + tmem slot
constexpr int CONV_SMEM = OFF_BAR + 64;
static_assert(CONV_SMEM <= 227 * 1024, "conv kernel shared memory");

constexpr int CONV_THREADS = 128;

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_stack_kernel(const CdqBoard* __restrict__ boards, long long n, const PackedNet net,
                  __half* __restrict__ act2, __half* __restrict__ dbg_act1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* planes = smem + OFF_PLANES;
    uint8_t* act1 = smem + OFF_ACT1;
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + OFF_BAR);     // weights landed
    uint64_t* bar_mma = bar_w + 1;                                     // MMA batch complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
    const float* sb2 = reinterpret_cast<const float*>(smem + OFF_B2);
    u64* srec = reinterpret_cast<u64*>(smem + OFF_REC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-time setup: zero the activation boards (borders stay zero for the whole kernel),
    //      fetch the packed weights with bulk copies, allocate TMEM
    for (int i = tid; i < (PLANES_BYTES + ACT1_BYTES) / 16; i += CONV_THREADS)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid < 64) reinterpret_cast<float*>(smem + OFF_B2)[tid] = net.conv2_b[tid];
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar_w, W1_BYTES + W2_BYTES);
        bulk_g2s(smem + OFF_W1, net.w1p, W1_BYTES, bar_w);
        bulk_g2s(smem + OFF_W2, net.w2p, W2_BYTES, bar_w);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    mbar_wait(bar_w, 0);

    const uint32_t planes_a = smem_u32(planes), act1_a = smem_u32(act1);
    const uint32_t w1_a = smem_u32(smem + OFF_W1), w2_a = smem_u32(smem + OFF_W2);
    constexpr uint32_t IDESC1 = make_idesc_f16(128, 32);
    constexpr uint32_t IDESC2 = make_idesc_f16(128, 64);
    uint32_t mma_phase = 0;

    // this thread's TMEM lane = tile row: position p = row / 8, file x = row % 8
    const int row = tid;
    const int p_loc = row >> 3, x_loc = row & 7;
    const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);

    const long long n_groups = (n + GROUP - 1) / GROUP;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long long pos0 = g * GROUP;
        // ---- stage the group's records (missing boards -> empty)
        {
            const long long remaining = n - pos0;
            const int cnt = remaining < GROUP ? (int)remaining : GROUP;
            const u64* grec = reinterpret_cast<const u64*>(boards) + pos0 * 9;
            for (int i = tid; i < GROUP * 9; i += CONV_THREADS) srec[i] = i < cnt * 9 ? __ldg(grec + i) : 0ull;
        }
        __syncthreads();
        // ---- encode: thread = (position tid/8, rank tid%8) writes 8 squares x 32 B one-hot
        {
            const int p = tid >> 3, y = tid & 7;
            const u64* r = srec + p * 9;
            unsigned tb[6];
#pragma unroll
            for (int t = 0; t < 6; t++) tb[t] = (unsigned)(r[t] >> (8 * y)) & 0xffu;
            const unsigned wb = (unsigned)(r[6] >> (8 * y)) & 0xffu;
            const unsigned bb = (unsigned)(r[7] >> (8 * y)) & 0xffu;
            uint8_t* dst0 = planes + p * BOARD_BYTES + ((y + 1) * 10 + 1) * CELL;
#pragma unroll
            for (int x = 0; x < 8; x++) {
                int t = -1;
#pragma unroll
                for (int k = 0; k < 6; k++) if ((tb[k] >> x) & 1) t = k;
                const bool isw = (wb >> x) & 1, isb = (bb >> x) & 1;
                const int ch = (t >= 0 && (isw || isb)) ? (isw ? t : t + 6) : -1; // board_utils.py:17-30
                u64 lo0 = 0, hi0 = 0, lo1 = 0;
                const u64 one = 0x3C00ull; // fp16 1.0
                if (ch >= 0 && ch < 4) lo0 = one << (16 * ch);
                else if (ch >= 4 && ch < 8) hi0 = one << (16 * (ch - 4));
                else if (ch >= 8) lo1 = one << (16 * (ch - 8));
                const u64 hi1 = one; // channel 12: constant one (carries conv1's bias)
                *reinterpret_cast<uint4*>(dst0 + x * CELL) =
                    make_uint4((unsigned)lo0, (unsigned)(lo0 >> 32), (unsigned)hi0, (unsigned)(hi0 >> 32));
                *reinterpret_cast<uint4*>(dst0 + CHUNK_BYTES + x * CELL) =
                    make_uint4((unsigned)lo1, (unsigned)(lo1 >> 32), (unsigned)hi1, (unsigned)(hi1 >> 32));
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();

        // ---- conv1: 8 rank tiles x 9 taps, N = 32, K = 16
        if (tid == 0) {
            tc_fence_after();
            for (int y = 0; y < 8; y++) {
                bool acc = false;
                for (int tap = 0; tap < 9; tap++) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    if (y + dy < 0 || y + dy > 7) continue; // all-zero border rank
                    const uint32_t a = planes_a + ((y + dy + 1) * 10 + (dx + 1)) * CELL;
                    umma_f16(tmem + y * 32, make_smem_desc(a, CHUNK_BYTES, BOARD_BYTES),
                             make_smem_desc(w1_a + tap * W1_TAP_BYTES, 4 * 128, 128), IDESC1, acc);
                    acc = true;
                }
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();

        // ---- epilogue 1: ReLU -> fp16 -> act1 board of this thread's (position, file), rank y
        for (int y = 0; y < 8; y++) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + y * 32, v);
            tmem_ld_wait();
            uint8_t* dst = act1 + p_loc * BOARD_BYTES + ((y + 1) * 10 + (x_loc + 1)) * CELL;
#pragma unroll
            for (int kc = 0; kc < 4; kc++) {
                uint4 o;
                o.x = relu_pack_f16(__uint_as_float(v[kc * 8 + 0]), __uint_as_float(v[kc * 8 + 1]));
                o.y = relu_pack_f16(__uint_as_float(v[kc * 8 + 2]), __uint_as_float(v[kc * 8 + 3]));
                o.z = relu_pack_f16(__uint_as_float(v[kc * 8 + 4]), __uint_as_float(v[kc * 8 + 5]));
                o.w = relu_pack_f16(__uint_as_float(v[kc * 8 + 6]), __uint_as_float(v[kc * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + kc * CHUNK_BYTES) = o;
                if (dbg_act1 && pos0 + p_loc < n)
                    *reinterpret_cast<uint4*>(dbg_act1 + ((pos0 + p_loc) * 64 + y * 8 + x_loc) * 32 + kc * 8) = o;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();

        // ---- conv2: 8 rank tiles x 9 taps x 2 K-steps, N = 64
        if (tid == 0) {
            tc_fence_after();
            for (int y = 0; y < 8; y++) {
                bool acc = false;
                for (int tap = 0; tap < 9; tap++) {
                    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                    if (y + dy < 0 || y + dy > 7) continue;
                    const uint32_t a = act1_a + ((y + dy + 1) * 10 + (dx + 1)) * CELL;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        umma_f16(tmem + y * 64, make_smem_desc(a + j * 2 * CHUNK_BYTES, CHUNK_BYTES, BOARD_BYTES),
                                 make_smem_desc(w2_a + tap * W2_TAP_BYTES + j * 2 * 1024, 8 * 128, 128), IDESC2, acc);
                        acc = true;
                    }
                }
            }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();

        // ---- epilogue 2: bias + ReLU -> fp16 -> act2 in fc1's A-tile order
        //      tile T = position / 128, K-stage = square; inside a 16 KB block:
        //      [8 channel chunks][16 row groups][8 rows][8 halfs]
        {
            const long long pos = pos0 + p_loc;
            const long long tile = pos >> 7;
            const int pr = (int)(pos & 127);
            for (int y = 0; y < 8; y++) {
                const int sq = y * 8 + x_loc;
                uint8_t* dst = reinterpret_cast<uint8_t*>(act2) + ((tile * 64 + sq) << 14) + (pr >> 3) * 128 + (pr & 7) * 16;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_lane + y * 64 + half * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int kc = 0; kc < 4; kc++) {
                        const int c = half * 32 + kc * 8;
                        uint4 o;
                        o.x = relu_pack_f16(__uint_as_float(v[kc * 8 + 0]) + sb2[c + 0], __uint_as_float(v[kc * 8 + 1]) + sb2[c + 1]);
                        o.y = relu_pack_f16(__uint_as_float(v[kc * 8 + 2]) + sb2[c + 2], __uint_as_float(v[kc * 8 + 3]) + sb2[c + 3]);
                        o.z = relu_pack_f16(__uint_as_float(v[kc * 8 + 4]) + sb2[c + 4], __uint_as_float(v[kc * 8 + 5]) + sb2[c + 5]);
                        o.w = relu_pack_f16(__uint_as_float(v[kc * 8 + 6]) + sb2[c + 6], __uint_as_float(v[kc * 8 + 7]) + sb2[c + 7]);
                        *reinterpret_cast<uint4*>(dst + (half * 4 + kc) * 2048) = o;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }

    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// K2b: fc1 (4096 -> 256) + ReLU + fc2 (256 -> 1) + tanh
constexpr int FC_STAGES = 4;
constexpr int FC_A_BYTES = 128 * 64 * 2;  // 16 KB: 128 positions x 64 channels of one square
constexpr int FC_B_BYTES = 256 * 64 * 2;  // 32 KB: 256 outputs  x 64 channels of one square
constexpr int FC_STAGE_BYTES = FC_A_BYTES + FC_B_BYTES;
constexpr int FC_OFF_VEC = FC_STAGES * FC_STAGE_BYTES;   // fc1 bias[256], fc2 weight[256], fc2 bias
constexpr int FC_OFF_BAR = FC_OFF_VEC + 2 * 256 * 4 + 16;
constexpr int FC_SMEM = FC_OFF_BAR + 128;
static_assert(FC_SMEM <= 227 * 1024, "fc kernel shared memory");
constexpr int FC_THREADS = 192; // warp 0: bulk-copy producer, warp 1: MMA issuer, warps 2..5: epilogue

__global__ void __launch_bounds__(FC_THREADS, 1)
fc_head_kernel(const __half* __restrict__ act2, long long n, const PackedNet net,
               float* __restrict__ value, const unsigned* __restrict__ flags, float* __restrict__ leaf) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sb1 = reinterpret_cast<float*>(smem + FC_OFF_VEC);
    float* sw2 = sb1 + 256;
    float* sb2 = sw2 + 256;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + FC_OFF_BAR); // [FC_STAGES]
    uint64_t* empty = full + FC_STAGES;                              // [FC_STAGES]
    uint64_t* acc_full = empty + FC_STAGES;                          // [2]
    uint64_t* acc_empty = acc_full + 2;                              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += FC_THREADS) { sb1[i] = net.fc1_b[i]; sw2[i] = net.fc2_w[i]; }
    if (tid == 0) {
        sb2[0] = net.fc2_b[0];
        for (int s = 0; s < FC_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; s++) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const long long n_tiles = (n + 127) >> 7;
    constexpr uint32_t IDESC = make_idesc_f16(128, 256);

    if (warp == 0) {
        // ---------------- producer: one elected lane streams (A, B) K-stages
        if (lane == 0) {
            uint32_t it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x)
                for (int s = 0; s < 64; s++, it++) {
                    const int st = it % FC_STAGES;
                    const uint32_t ph = (it / FC_STAGES) & 1;
                    mbar_wait(empty + st, ph ^ 1);
                    uint8_t* sa = smem + st * FC_STAGE_BYTES;
                    mbar_arrive_expect_tx(full + st, FC_STAGE_BYTES);
                    bulk_g2s(sa, reinterpret_cast<const uint8_t*>(act2) + ((t * 64 + s) << 14), FC_A_BYTES, full + st);
                    bulk_g2s(sa + FC_A_BYTES, reinterpret_cast<const uint8_t*>(net.fc1p) + ((long long)s << 15), FC_B_BYTES, full + st);
                }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, tcount++) {
                const int ab = tcount & 1;
                mbar_wait(acc_empty + ab, ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int s = 0; s < 64; s++, it++) {
                    const int st = it % FC_STAGES;
                    mbar_wait(full + st, (it / FC_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + st * FC_STAGE_BYTES), b = a + FC_A_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        umma_f16(tmem + ab * 256, make_smem_desc(a + j * 2 * 2048, 2048, 128),
                                 make_smem_desc(b + j * 2 * 4096, 4096, 128), IDESC, (s | j) != 0);
                    umma_commit(empty + st); // frees the stage when these MMAs have read it
                }
                umma_commit(acc_full + ab);
            }
        }
    } else {
        // ---------------- epilogue: TMEM lane quarter = warp % 4, one position per thread
        const int q = warp & 3;
        uint32_t tcount = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, tcount++) {
            const int ab = tcount & 1;
            mbar_wait(acc_full + ab, (tcount >> 1) & 1);
            tc_fence_after();
            float dot = 0.f;
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 256 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float h = fmaxf(__uint_as_float(v[j]) + sb1[c * 32 + j], 0.f);
                    dot = fmaf(h, sw2[c * 32 + j], dot);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + ab);
            const long long pos = (t << 7) + q * 32 + lane;
            if (pos < n) {
                const float v = tanhf(dot + sb2[0]);
                if (value) value[pos] = v;
                if (leaf) {
                    // mcts.py:196-232: terminal short-circuits, else clamp(net, -1, 1)
                    const unsigned f = flags[pos];
                    float lv = fminf(1.f, fmaxf(-1.f, v));
                    if (f & CDQ_F_CHECKMATE) lv = (f & CDQ_F_WHITE_TO_MOVE) ? -1.f : 1.f;
                    else if (f & (CDQ_F_STALEMATE | CDQ_F_INSUFFICIENT | CDQ_F_REP3)) lv = 0.f;
                    leaf[pos] = lv;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// weight repacking: fp32 PyTorch layouts -> fp16 UMMA K-major no-swizzle tiles
__global__ void pack_weights_kernel(CdqWeights w, __half* w1p, __half* w2p, __half* fc1p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N1 = 9 * 32 * 16, N2 = 9 * 64 * 32, N3 = 64ll * 256 * 64;
    if (i < N1) {
        // w1p[tap][kc 2][ng 4][n8][k8]
        const int k8 = i & 7, n8 = (i >> 3) & 7, ng = (i >> 6) & 3, kc = (i >> 8) & 1, tap = (int)(i >> 9);
        const int nn = ng * 8 + n8, k = kc * 8 + k8;
        float v = 0.f;
        if (k < 12) v = w.conv1_w[(nn * 12 + k) * 9 + tap];
        else if (k == 12 && tap == 4) v = w.conv1_b[nn];
        w1p[i] = __float2half_rn(v);
    } else if (i < N1 + N2) {
        // w2p[tap][kc 4][ng 8][n8][k8]
        const long long j = i - N1;
        const int k8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 7, kc = (j >> 9) & 3, tap = (int)(j >> 11);
        const int nn = ng * 8 + n8, k = kc * 8 + k8;
        w2p[j] = __float2half_rn(w.conv2_w[(nn * 32 + k) * 9 + tap]);
    } else if (i < N1 + N2 + N3) {
        // fc1p[square 64][kc 8][ng 32][n8][k8]; reference column = c*64 + square
        const long long j = i - N1 - N2;
        const int k8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 31, kc = (j >> 11) & 7, sq = (int)(j >> 14);
        const int nn = ng * 8 + n8, c = kc * 8 + k8;
        fc1p[j] = __float2half_rn(w.fc1_w[(long long)nn * 4096 + c * 64 + sq]);
    }
}

int launch_pack_weights(const CdqWeights& w, __half* w1p, __half* w2p, __half* fc1p, cudaStream_t st) {
    const long long total = 9 * 32 * 16 + 9 * 64 * 32 + 64ll * 256 * 64;
    ProfileScope ps(KK_PACK, st);
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, w1p, w2p, fc1p);
    count_launch();
    return (int)cudaGetLastError();
}

static int g_sm_count = 0;
static int sm_count() {
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(conv_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM);
        cudaFuncSetAttribute(fc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM);
    }
    return g_sm_count;
}

int launch_conv_stack(const CdqBoard* boards, long long n, const PackedNet& net, __half* act2, __half* dbg_act1,
                      cudaStream_t st) {
    if (n == 0) return 0;
    const long long groups = (n + GROUP - 1) / GROUP;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(groups < sms ? groups : sms);
    ProfileScope ps(KK_CONV, st);
    conv_stack_kernel<<<grid, CONV_THREADS, CONV_SMEM, st>>>(boards, n, net, act2, dbg_act1);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_fc_head(const __half* act2, long long n, const PackedNet& net, float* value, const unsigned* flags,
                   float* leaf, cudaStream_t st) {
    if (n == 0) return 0;
    const long long tiles = (n + 127) / 128;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    ProfileScope ps(KK_FC, st);
    fc_head_kernel<<<grid, FC_THREADS, FC_SMEM, st>>>(act2, n, net, value, flags, leaf);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
