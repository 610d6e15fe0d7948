// fc_kernel.cu -- K2b: fc1 (4096 -> 256) + ReLU + fc2 (256 -> 1) + tanh on the conv stack's output,
// plus the weight repacking kernel (fp32 PyTorch layouts -> fp16 UMMA tiles).
//
// The body of one tile (fc_head.cuh, also the fc role of pipe_kernel.cu) is a warp-specialised tcgen05 GEMM:
// M = 128 positions x N = 256 x K = 4096, both operands fetched by 1-D bulk async copies (TMA engine, UBLKCP)
// from pre-tiled global buffers through an mbarrier ring, accumulators in TMEM.  The epilogue applies bias + ReLU
// and the 256 -> 1 fc2 dot product per TMEM lane (= per position), tanh, and optionally the leaf
// post-processing of mcts.py:196-232.  Algorithmic work: 2 097 664 FLOP/position; reads 8 KB/position of act2.
#include <cstdlib>
#include "fc_head.cuh"

namespace cdq {

// ------------------------------------------------------------------------------------------
// fc_head_pairs_kernel: the same GEMM with TWO position tiles per weight stage.  One tile per stage moves 16 KB of
// A and 32 KB of B per 64 K-elements, i.e. every tile re-reads the whole 2 MB of fc1 weights from L2: 3 MB per
// tile, 24 GB of L2 -> SM traffic per 1M positions, and an SM keeps only its 192 KB ring in flight (about
// 110 GB/s at the L2 latency under load), which is what a tile took.  Here a stage holds A of tiles 2p and 2p+1
// and ONE B (64 KB, three stages), the two 128 x 256 accumulators fill all 512 TMEM columns, and a pair costs
// 4 MB: a third less operand traffic per tile.  The accumulators are single buffered, so the tensor pipe waits
// while the epilogue drains them (16 TMEM loads per pair); the producer keeps prefetching meanwhile.  Same UMMAs
// in the same K order per tile as the single-tile body, so the values are bit-identical to the pipe kernel's.
constexpr int FCP_STAGES = 3;
constexpr int FCP_STAGE_BYTES = 2 * FC_A_BYTES + FC_B_BYTES;
constexpr int FCP_OFF_VEC = FCP_STAGES * FCP_STAGE_BYTES;
constexpr int FCP_OFF_BAR = FCP_OFF_VEC + 2 * 256 * 4 + 16;
constexpr int FCP_SMEM = FCP_OFF_BAR + 128;
static_assert(FCP_SMEM <= 227 * 1024, "fc pairs kernel shared memory");

__global__ void __launch_bounds__(FC_THREADS, 1)
fc_head_pairs_kernel(const __half* __restrict__ act2, long long n, const PackedNet net,
                     float* __restrict__ value, const unsigned* __restrict__ flags, float* __restrict__ leaf,
                     float* __restrict__ hidden) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sb1 = reinterpret_cast<float*>(smem + FCP_OFF_VEC);
    float* sw2 = sb1 + 256;
    float* sb2 = sw2 + 256;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + FCP_OFF_BAR); // [FCP_STAGES]
    uint64_t* empty = full + FCP_STAGES;                              // [FCP_STAGES]
    uint64_t* acc_full = empty + FCP_STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += (int)blockDim.x) { sb1[i] = net.fc1_b[i]; sw2[i] = net.fc2_w[i]; }
    if (tid == 0) {
        sb2[0] = net.fc2_b[0];
        for (int s = 0; s < FCP_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const long long n_tiles = (n + 127) >> 7, n_pairs = (n_tiles + 1) >> 1;
    const int cta = (int)blockIdx.x, n_cta = (int)gridDim.x;
    constexpr uint32_t IDESC = make_idesc_f16(128, 256);

    if (warp == 0) {
        // ---------------- producer
        if (lane == 0) {
            uint32_t it = 0;
            for (long long p = cta; p < n_pairs; p += n_cta) {
                const bool two = 2 * p + 1 < n_tiles;
                const uint8_t* a_tile = reinterpret_cast<const uint8_t*>(act2) + (p << 21);
                for (int s = 0; s < 64; s++, it++) {
                    const int sq = ((s & 7) << 3) | (s >> 3);          // the K order of fc_head.cuh (file-major)
                    const int st = it % FCP_STAGES;
                    mbar_wait(empty + st, ((it / FCP_STAGES) & 1) ^ 1);
                    uint8_t* sa = smem + st * FCP_STAGE_BYTES;
                    mbar_arrive_expect_tx(full + st, (two ? 2 : 1) * FC_A_BYTES + FC_B_BYTES);
                    bulk_g2s(sa, a_tile + ((long long)sq << 14), FC_A_BYTES, full + st);
                    if (two) bulk_g2s(sa + FC_A_BYTES, a_tile + (1ll << 20) + ((long long)sq << 14), FC_A_BYTES, full + st);
                    bulk_g2s(sa + 2 * FC_A_BYTES, reinterpret_cast<const uint8_t*>(net.fc1p) + ((long long)sq << 15), FC_B_BYTES, full + st);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            uint32_t it = 0, pc = 0;
            for (long long p = cta; p < n_pairs; p += n_cta, pc++) {
                const bool two = 2 * p + 1 < n_tiles;
                mbar_wait(acc_empty, (pc & 1) ^ 1);
                tc_fence_after();
                for (int s = 0; s < 64; s++, it++) {
                    const int st = it % FCP_STAGES;
                    mbar_wait(full + st, (it / FCP_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a = smem_u32(smem + st * FCP_STAGE_BYTES), b = a + 2 * FC_A_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        umma_f16(tmem, make_smem_desc(a + j * 2 * 2048, 2048, 128), make_smem_desc(b + j * 2 * 4096, 4096, 128), IDESC,
                                 (s | j) != 0);
                    if (two) {
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            umma_f16(tmem + 256, make_smem_desc(a + FC_A_BYTES + j * 2 * 2048, 2048, 128),
                                     make_smem_desc(b + j * 2 * 4096, 4096, 128), IDESC, (s | j) != 0);
                    }
                    umma_commit(empty + st);
                }
                umma_commit(acc_full);
            }
        }
    } else if (warp < 6) {
        // ---------------- epilogue: TMEM lane quarter = warp % 4, one position per thread and tile
        const int q = warp & 3;
        uint32_t pc = 0;
        for (long long p = cta; p < n_pairs; p += n_cta, pc++) {
            const int n_here = 2 * p + 1 < n_tiles ? 2 : 1;
            mbar_wait(acc_full, pc & 1);
            tc_fence_after();
            float dots[2] = {0.f, 0.f};
#pragma unroll
            for (int tt = 0; tt < 2; tt++) {
                if (tt < n_here) {
                    const long long pos = ((2 * p + tt) << 7) + q * 32 + lane;
                    float dot = 0.f;
#pragma unroll 1
                    for (int c = 0; c < 8; c++) {
                        uint32_t v[32];
                        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + tt * 256 + c * 32, v);
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
                    dots[tt] = dot;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
#pragma unroll
            for (int tt = 0; tt < 2; tt++) {
                const long long pos = ((2 * p + tt) << 7) + q * 32 + lane;
                if (tt < n_here && pos < n) {
                    const float v = tanhf(dots[tt] + sb2[0]);
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
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// weight repacking: fp32 PyTorch layouts -> fp16 UMMA K-major no-swizzle tiles.
// The two fc1 buffers (98 % of the bytes) are written one 16-byte vector per thread: the eight source
// floats of a vector are 64 or 4096 floats apart, so the thread index runs over the SQUARE first and every
// one of the eight loads of a warp is one coalesced 128-byte line (the first version moved one half
// per thread with 256-byte-strided loads: 12 us of the 320 us training step).  The conv tiles are small
// and stay one element per thread.
__device__ __forceinline__ uint4 pack8_f16(const float (&v)[8]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    return make_uint4(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b), *reinterpret_cast<unsigned*>(&c),
                      *reinterpret_cast<unsigned*>(&d));
}
constexpr long long PK_N1 = 9 * 32 * 16, PK_N2 = 9 * 64 * 32, PK_V3 = 64ll * 256 * 64 / 8;
constexpr long long PK_SMALL = 2 * PK_N1 + 3 * PK_N2 + 577;
constexpr long long PK_SMALL_PAD = (PK_SMALL + 255) / 256 * 256;     // the vector parts start on a CTA boundary
__global__ void pack_weights_kernel(CdqWeights w, __half* w1p, __half* w2p, __half* fc1p, __half* fc1tp, __half* w2tp,
                                    __half* wss, float* conv2_b, float* fc1_b, float* fc2_w, float* fc2_b, float* zero_me,
                                    long long first) {
    const long long i = first + (long long)blockIdx.x * blockDim.x + threadIdx.x;   // first: PK_SMALL_PAD when only fc1's tiles are packed
    if (i == 0 && zero_me) *zero_me = 0.f;       // the training step's loss accumulator (saves a fill launch in the graph)
    constexpr long long N1 = PK_N1, N2 = PK_N2;
    if (i >= PK_SMALL_PAD) {
        const long long t = i - PK_SMALL_PAD;
        float v[8];
        if (t < PK_V3) {
            // fc1p[square 64][kc 8][ng 32][n8][k8]; reference column = c*64 + square
            const int sq = t & 63, kc = (t >> 6) & 7, nn = (int)(t >> 9);
            const float* src = w.fc1_w + (long long)nn * 4096 + kc * 512 + sq;
#pragma unroll
            for (int k8 = 0; k8 < 8; k8++) v[k8] = src[k8 * 64];
            *reinterpret_cast<uint4*>(fc1p + ((((long long)sq * 8 + kc) * 32 + (nn >> 3)) * 8 + (nn & 7)) * 8) = pack8_f16(v);
        } else if (t < 2 * PK_V3) {
            // fc1tp[chunk 16][stage 4][kc 8][ng 32][n8][j8], K = fc1 output j.  A chunk of 256 rows is one RANK and one HALF of
            // the channels, row in chunk = file * 32 + channel in half: a 32-column accumulator block of the data gradient
            // (fc_bwd_kernel.cu) is then one square, and the eight blocks of a chunk are the eight files of a board row.
            const long long u = t - PK_V3;
            const int sq = u & 63, ch = (u >> 6) & 63, jg = (int)(u >> 12);          // jg = sg*8 + kc
            const float* src = w.fc1_w + (long long)jg * 8 * 4096 + ch * 64 + sq;
#pragma unroll
            for (int j8 = 0; j8 < 8; j8++) v[j8] = src[(long long)j8 * 4096];
            const int ck = (sq >> 3) * 2 + (ch >> 5), r256 = (sq & 7) * 32 + (ch & 31), ng = r256 >> 3, n8 = r256 & 7;
            *reinterpret_cast<uint4*>(fc1tp + (((((long long)ck * 32 + jg) * 32 + ng) * 8 + n8) * 8)) = pack8_f16(v);
        }
        return;
    }
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
    } else if (i < N1 + 2 * N2) {
        // w2tp[dy 3][kc 8 (co)][dx 3][ng 4 (ci)][ci8][co8] = conv2.weight[co][ci][8 - (3 dy + dx)]   (flipped taps; the
        // three taps of a kernel row are one B operand of N = 96, conv_bwd2_kernel.cu)
        const long long j = i - N1 - N2;
        const int co8 = j & 7, ci8 = (j >> 3) & 7, ng = (j >> 6) & 3, blk = (int)(j >> 8);
        const int dxi = blk % 3, kc = (blk / 3) & 7, u = (blk / 24) * 3 + dxi;
        const int co = kc * 8 + co8, ci = ng * 8 + ci8;
        w2tp[j] = __float2half_rn(w.conv2_w[(co * 32 + ci) * 9 + (8 - u)]);
    } else if (i < 2 * N1 + 2 * N2) {
        // wss, first part: conv1 in the stream kernel's order [dy 3][kc 2][dx = +1, 0, -1][ng 4][n8][k8]
        const long long j = i - N1 - 2 * N2;
        const int k8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 3, blk = (int)(j >> 8);
        const int dxb = blk % 3, kc = (blk / 3) % 2, dyi = blk / 6, tap = dyi * 3 + (2 - dxb);
        const int nn = ng * 8 + n8, k = kc * 8 + k8;
        float v = 0.f;
        if (k < 12) v = w.conv1_w[(nn * 12 + k) * 9 + tap];
        else if (k == 12 && tap == 4) v = w.conv1_b[nn];
        wss[j] = __float2half_rn(v);
    } else if (i < 2 * N1 + 3 * N2) {
        // wss, second part: conv2 as [dy 3][kc 4][dx = +1, 0, -1][ng 8][n8][k8]
        const long long j = i - 2 * N1 - 2 * N2;
        const int k8 = j & 7, n8 = (j >> 3) & 7, ng = (j >> 6) & 7, blk = (int)(j >> 9);
        const int dxb = blk % 3, kc = (blk / 3) % 4, dyi = blk / 12, tap = dyi * 3 + (2 - dxb);
        const int nn = ng * 8 + n8, k = kc * 8 + k8;
        wss[N1 + j] = __float2half_rn(w.conv2_w[(nn * 32 + k) * 9 + tap]);
    } else {
        // the fp32 vectors the kernels read directly: conv2 bias, fc1 bias, fc2 weight, fc2 bias
        const long long j = i - 2 * N1 - 3 * N2;
        if (j < 64) conv2_b[j] = w.conv2_b[j];
        else if (j < 320) fc1_b[j - 64] = w.fc1_b[j - 64];
        else if (j < 576) fc2_w[j - 320] = w.fc2_w[j - 320];
        else if (j == 576) fc2_b[0] = w.fc2_b[0];
    }
}

// part: 0 = everything; 1 = the two fc1 tile sets only (2 x 2 MB, almost all of the work); 2 = everything else (conv tiles, fp32 vectors)
int launch_pack_weights(const CdqWeights& w, const PackedNet& p, cudaStream_t st, float* zero_me, int part) {
    const long long first = part == 1 ? PK_SMALL_PAD : 0;
    const long long total = (part == 2 ? PK_SMALL_PAD : PK_SMALL_PAD + 2 * PK_V3) - first;
    ProfileScope ps(KK_PACK, st);
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        w, (__half*)p.w1p, (__half*)p.w2p, (__half*)p.fc1p, (__half*)p.fc1tp, (__half*)p.w2tp, (__half*)p.wss, (float*)p.conv2_b,
        (float*)p.fc1_b, (float*)p.fc2_w, (float*)p.fc2_b, zero_me, first);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_fc_head(const __half* act2, long long n, const PackedNet& net, float* value, const unsigned* flags,
                   float* leaf, float* hidden, cudaStream_t st) {
    if (n == 0) return 0;
    const long long tiles = (n + 127) / 128;
    if (tiles <= FC_SPLIT_MAX_TILES && fc_split_enabled()) return launch_fc_head_split(act2, n, net, value, flags, leaf, hidden, st);
    once_per_device(ONCE_FC_PAIRS, [] { cudaFuncSetAttribute(fc_head_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FCP_SMEM); });
    const int sms = device_sm_count();
    const long long pairs = (tiles + 1) / 2;
    ProfileScope ps(KK_FC, st);
    fc_head_pairs_kernel<<<(unsigned)(pairs < sms ? pairs : sms), FC_THREADS, FCP_SMEM, st>>>(act2, n, net, value, flags, leaf, hidden);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
