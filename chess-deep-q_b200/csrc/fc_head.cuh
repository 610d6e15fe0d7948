// fc_head.cuh -- body of K2b (fc1 + ReLU + fc2 + tanh, see fc_kernel.cu), shared by the stand-alone
// kernel and by the conv -> fc pipe kernel (pipe_kernel.cu).  Uses warps 0..5 of the CTA.
#pragma once
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"
#include "conv_stream.cuh"   // PipeCtl, pipe_wait_ge

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

template <bool PIPE>
__device__ __forceinline__ void fc_head_body(uint8_t* smem, const __half* __restrict__ act2, long long n, const PackedNet& net,
                                             float* __restrict__ value, const unsigned* __restrict__ flags,
                                             float* __restrict__ leaf, float* __restrict__ hidden, const int cta,
                                             const int n_cta, const cs::PipeCtl& pipe) {
    float* sb1 = reinterpret_cast<float*>(smem + FC_OFF_VEC);
    float* sw2 = sb1 + 256;
    float* sb2 = sw2 + 256;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + FC_OFF_BAR); // [FC_STAGES]
    uint64_t* empty = full + FC_STAGES;                              // [FC_STAGES]
    uint64_t* acc_full = empty + FC_STAGES;                          // [2]
    uint64_t* acc_empty = acc_full + 2;                              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 256; i += (int)blockDim.x) { sb1[i] = net.fc1_b[i]; sw2[i] = net.fc2_w[i]; }
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
            unsigned seen = 0;      // prefetched `ready` counter of the unit that starts next
            for (long long t = cta; t < n_tiles; t += n_cta)
            {
                const uint8_t* a_tile = reinterpret_cast<const uint8_t*>(act2) + (t << 20);
                unsigned slot = 0, use = 0;
                if (PIPE) {
                    slot = (unsigned)(t % pipe.ring_tiles);
                    use = (unsigned)(t / pipe.ring_tiles);
                    a_tile = pipe.ring + ((size_t)slot << 20);
                    if (t == cta) seen = cs::pipe_load(pipe.ready + slot * cs::PIPE_UNITS);
                }
                for (int s = 0; s < 64; s++, it++) {
                    // K order is FILE-major (square = 8 (s % 8) + s / 8): the conv side writes a tile file by
                    // file, so the first stages of a tile are the first data it produces
                    const int sq = ((s & 7) << 3) | (s >> 3);
                    if (PIPE && (s & (8 * cs::PIPE_GF - 1)) == 0) {
                        // consumer side of the L2-resident hand-over: all 64 (warp, sub-tile) pieces of this
                        // unit's current use are stored (`seen` was requested one unit ago, so the check
                        // normally costs nothing); order the bulk (async-proxy) reads after them
                        const int u = s / (8 * cs::PIPE_GF);
                        if (seen < 64u * (use + 1u)) cs::pipe_wait_ge(pipe.ready + slot * cs::PIPE_UNITS + u, 64u * (use + 1u));
                        asm volatile("fence.proxy.async.global;" ::: "memory");
                        // request the next unit's counter now; it is looked at when that unit starts
                        const long long tn = u + 1 < cs::PIPE_UNITS ? t : t + n_cta;
                        const int un = u + 1 < cs::PIPE_UNITS ? u + 1 : 0;
                        seen = tn < n_tiles ? cs::pipe_load(pipe.ready + (unsigned)(tn % pipe.ring_tiles) * cs::PIPE_UNITS + un) : 0u;
                        if (tn != t) seen = seen >= 64u * ((unsigned)(tn / pipe.ring_tiles) + 1u) ? 0xffffffffu : 0u;   // other tile, other use
                    }
                    const int st = it % FC_STAGES;
                    const uint32_t ph = (it / FC_STAGES) & 1;
                    mbar_wait(empty + st, ph ^ 1);
                    uint8_t* sa = smem + st * FC_STAGE_BYTES;
                    mbar_arrive_expect_tx(full + st, FC_STAGE_BYTES);
                    bulk_g2s(sa, a_tile + ((long long)sq << 14), FC_A_BYTES, full + st);
                    bulk_g2s(sa + FC_A_BYTES, reinterpret_cast<const uint8_t*>(net.fc1p) + ((long long)sq << 15), FC_B_BYTES, full + st);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            for (long long t = cta; t < n_tiles; t += n_cta, tcount++) {
                const int ab = tcount & 1;
                mbar_wait(acc_empty + ab, ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int s = 0; s < 64; s++, it++) {
                    const int st = it % FC_STAGES;
                    mbar_wait(full + st, (it / FC_STAGES) & 1);
                    tc_fence_after();
                    // every stage of this unit has landed in shared memory: the conv side may overwrite it
                    if (PIPE && (s & (8 * cs::PIPE_GF - 1)) == 8 * cs::PIPE_GF - 1)
                        atomicAdd(pipe.consumed + (unsigned)(t % pipe.ring_tiles) * cs::PIPE_UNITS + s / (8 * cs::PIPE_GF), 1u);
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
    } else if (warp < 6) {
        // ---------------- epilogue: TMEM lane quarter = warp % 4, one position per thread
        const int q = warp & 3;
        uint32_t tcount = 0;
        for (long long t = cta; t < n_tiles; t += n_cta, tcount++) {
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

} // namespace cdq
