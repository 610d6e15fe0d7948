// fc_kernel.cu -- K2b: fc1 (4096 -> 256) + ReLU + fc2 (256 -> 1) + tanh on the conv stack's output,
// plus the weight repacking kernel (fp32 PyTorch layouts -> fp16 UMMA tiles).
//
// fc_head_kernel is a warp-specialised tcgen05 GEMM: M = 128 positions x N = 256 x K = 4096 per
// tile, both operands fetched by 1-D bulk async copies (TMA engine, UBLKCP) from pre-tiled global
// buffers through a 4-stage mbarrier ring, accumulators double-buffered in TMEM (2 x 256 columns).
// The epilogue applies bias + ReLU and the 256 -> 1 fc2 dot product per TMEM lane (= per position),
// tanh, and optionally the leaf post-processing of mcts.py:196-232.
// Algorithmic work: 2 097 664 FLOP/position; reads 8 KB/position of act2.
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {

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
               float* __restrict__ value, const unsigned* __restrict__ flags, float* __restrict__ leaf,
               float* __restrict__ hidden /* optional [n][256] ReLU(fc1), kept for the backward pass */) {
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
            const long long pos = (t << 7) + q * 32 + lane;
#pragma unroll 1
            for (int c = 0; c < 8; c++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + ab * 256 + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float h = fmaxf(__uint_as_float(v[j]) + sb1[c * 32 + j], 0.f);
                    dot = fmaf(h, sw2[c * 32 + j], dot);
                    v[j] = __float_as_uint(h);
                }
                if (hidden && pos < n) {
                    float4* hp = reinterpret_cast<float4*>(hidden + pos * 256 + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        hp[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + ab);
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
__global__ void pack_weights_kernel(CdqWeights w, __half* w1p, __half* w2p, __half* fc1p, __half* fc1tp, __half* w2tp,
                                    __half* wss, float* conv2_b, float* fc1_b, float* fc2_w, float* fc2_b) {
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
    } else if (i < N1 + N2 + 2 * N3) {
        // fc1tp[chunk 16][stage 4][kc 8][ng 32][n8][j8]: row n' = sq*64 + ch, K = fc1 output j
        const long long j = i - N1 - N2 - N3;
        const int j8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 31, kc = (j >> 11) & 7, sg = (j >> 14) & 3, ck = (int)(j >> 16);
        const int np = ck * 256 + ng * 8 + n8, sq = np >> 6, ch = np & 63, jo = sg * 64 + kc * 8 + j8;
        fc1tp[j] = __float2half_rn(w.fc1_w[(long long)jo * 4096 + ch * 64 + sq]);
    } else if (i < N1 + 2 * N2 + 2 * N3) {
        // w2tp[u 9][kc 8 (co)][ng 4 (ci)][ci8][co8] = conv2.weight[co][ci][8 - u]   (flipped taps)
        const long long j = i - N1 - N2 - 2 * N3;
        const int co8 = j & 7, ci8 = (j >> 3) & 7, ng = (j >> 6) & 3, kc = (j >> 8) & 7, u = (int)(j >> 11);
        const int co = kc * 8 + co8, ci = ng * 8 + ci8;
        w2tp[j] = __float2half_rn(w.conv2_w[(co * 32 + ci) * 9 + (8 - u)]);
    } else if (i < 2 * N1 + 2 * N2 + 2 * N3) {
        // wss, first part: conv1 in the stream kernel's order [dy 3][kc 2][dx = +1, 0, -1][ng 4][n8][k8]
        const long long j = i - N1 - 2 * N2 - 2 * N3;
        const int k8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 3, blk = (int)(j >> 8);
        const int dxb = blk % 3, kc = (blk / 3) % 2, dyi = blk / 6, tap = dyi * 3 + (2 - dxb);
        const int nn = ng * 8 + n8, k = kc * 8 + k8;
        float v = 0.f;
        if (k < 12) v = w.conv1_w[(nn * 12 + k) * 9 + tap];
        else if (k == 12 && tap == 4) v = w.conv1_b[nn];
        wss[j] = __float2half_rn(v);
    } else if (i < 2 * N1 + 3 * N2 + 2 * N3) {
        // wss, second part: conv2 as [dy 3][kc 4][dx = +1, 0, -1][ng 8][n8][k8]
        const long long j = i - 2 * N1 - 2 * N2 - 2 * N3;
        const int k8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 7, blk = (int)(j >> 9);
        const int dxb = blk % 3, kc = (blk / 3) % 4, dyi = blk / 12, tap = dyi * 3 + (2 - dxb);
        const int nn = ng * 8 + n8, k = kc * 8 + k8;
        wss[N1 + j] = __float2half_rn(w.conv2_w[(nn * 32 + k) * 9 + tap]);
    } else {
        // the fp32 vectors the kernels read directly: conv2 bias, fc1 bias, fc2 weight, fc2 bias
        const long long j = i - 2 * N1 - 3 * N2 - 2 * N3;
        if (j < 64) conv2_b[j] = w.conv2_b[j];
        else if (j < 320) fc1_b[j - 64] = w.fc1_b[j - 64];
        else if (j < 576) fc2_w[j - 320] = w.fc2_w[j - 320];
        else if (j == 576) fc2_b[0] = w.fc2_b[0];
    }
}

int launch_pack_weights(const CdqWeights& w, const PackedNet& p, cudaStream_t st) {
    const long long total = 2 * 9 * 32 * 16 + 3 * 9 * 64 * 32 + 2 * 64ll * 256 * 64 + 577;
    ProfileScope ps(KK_PACK, st);
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        w, (__half*)p.w1p, (__half*)p.w2p, (__half*)p.fc1p, (__half*)p.fc1tp, (__half*)p.w2tp, (__half*)p.wss, (float*)p.conv2_b,
        (float*)p.fc1_b, (float*)p.fc2_w, (float*)p.fc2_b);
    count_launch();
    return (int)cudaGetLastError();
}

int device_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

int launch_fc_head(const __half* act2, long long n, const PackedNet& net, float* value, const unsigned* flags,
                   float* leaf, float* hidden, cudaStream_t st) {
    if (n == 0) return 0;
    const long long tiles = (n + 127) / 128;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(fc_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM);
        attr_set = true;
    }
    const int sms = device_sm_count();
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    ProfileScope ps(KK_FC, st);
    fc_head_kernel<<<grid, FC_THREADS, FC_SMEM, st>>>(act2, n, net, value, flags, leaf, hidden);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
