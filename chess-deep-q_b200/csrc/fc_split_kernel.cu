// fc_split_kernel.cu -- K2b for SMALL batches: fc1 (4096 -> 256) + ReLU + fc2 (256 -> 1) + tanh with the K
// dimension of fc1 split over a thread-block CLUSTER of four CTAs.
//
// fc_head_kernel gives every 128-position tile to one CTA, which streams 3 MB (1 MB of act2, 2 MB of fc1
// weights) through its shared memory: about 27 us at the 110 GB/s one SM pulls from L2, whatever the batch
// size.  A replay batch of 4 096 positions is 32 tiles, so 32 of the 148 SMs worked for those 27 us, twice per
// training step, and an MCTS wave of a few thousand leaves paid the same latency.  Here a tile belongs to a
// cluster of four CTAs:
//   * CTA r contracts squares 16 r .. 16 r + 15 (768 KB of operands, 64 UMMAs of M = 128, N = 256, K = 16)
//     into its own TMEM accumulator;
//   * every CTA parks its 128 x 256 fp32 partial sums in its shared memory (the operand ring is dead by then) as
//     [4-column group][row][4 columns]: a warp reads 512 contiguous bytes per group, 16 bytes per lane;
//   * after a cluster barrier CTA r owns fc1 outputs 64 r .. 64 r + 63: it adds the four partial sums through
//     distributed shared memory (always in rank order 0, 1, 2, 3, so the bits do not depend on who adds),
//     applies bias + ReLU, optionally keeps ReLU(fc1) for the backward pass, and reduces its slice of the fc2
//     dot product per position;
//   * the four partial dot products meet in CTA 0's shared memory (remote stores, second cluster barrier), and
//     CTA 0 applies fc2's bias, tanh and the leaf post-processing of mcts.py:196-232.
// No global scratch, no atomics, no second launch.  Used for at most FC_SPLIT_MAX_TILES tiles (two waves of
// 32 clusters); larger batches keep the persistent one-CTA-per-tile kernel, whose fp32 summation order differs
// (one chain of 256 UMMAs instead of four chains of 64), so values from the two kernels agree to fp32 rounding,
// not bit for bit.
#include <cstdlib>
#include "fc_head.cuh"

namespace cdq {

constexpr int FS_SPLIT = 4;                               // CTAs per cluster = K slices
constexpr int FS_SQUARES = 64 / FS_SPLIT;                 // squares (K stages) per CTA
constexpr int FS_STAGES = 4;
constexpr int FS_STAGE_BYTES = FC_A_BYTES + FC_B_BYTES;   // 48 KB
constexpr int FS_PART_BYTES = 256 * 128 * 4;              // partial sums [256 columns][128 rows] fp32, over the ring
static_assert(FS_PART_BYTES <= FS_STAGES * FS_STAGE_BYTES, "partial sums alias the operand ring");
constexpr int FS_OFF_VEC = FS_STAGES * FS_STAGE_BYTES;    // fc1 bias[256], fc2 weight[256], fc2 bias (+ pad)
constexpr int FS_OFF_DOT = FS_OFF_VEC + 2 * 256 * 4 + 16; // [4 ranks][128 rows] partial fc2 dot products (CTA 0's copy is used)
constexpr int FS_OFF_BAR = FS_OFF_DOT + FS_SPLIT * 128 * 4;
constexpr int FS_SMEM = FS_OFF_BAR + 128;
static_assert(FS_SMEM <= 227 * 1024, "fc split kernel shared memory");

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float dsmem_ld(uint32_t addr) {
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 dsmem_ld4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void dsmem_st(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// sums v[0..31] over the 32 lanes of a warp: afterwards lane l holds the total of element l (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const float keep = up ? v[i + s] : v[i], send = up ? v[i] : v[i + s];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// TRAIN: the online network's forward of the replay step.  The loss, tanh' and fc2's backward need nothing but q (known in
// CTA 0 right after the forward), q' of the target network (complete before this launch) and ReLU(fc1), which every CTA still
// holds in registers -- so the head of the backward pass (head_bwd_kernel: 9 us of the step's critical chain plus a 4 MB round
// trip of ReLU(fc1)) is this kernel's epilogue: CTA 0 broadcasts d loss / d z2 per position through distributed shared memory,
// every CTA writes its 64 columns of fc1's pre-activation gradient as the fp16 tiles fc1's backward kernels read and reduces
// its slice of the fc2-weight and fc1-bias gradients.
template <bool TRAIN>
__global__ void __cluster_dims__(FS_SPLIT, 1, 1) __launch_bounds__(FC_THREADS, 1)
fc_head_split_kernel(const __half* __restrict__ act2, long long n, const PackedNet net, float* __restrict__ value,
                     const unsigned* __restrict__ flags, float* __restrict__ leaf, float* __restrict__ hidden, const TrainHead th) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sb1 = reinterpret_cast<float*>(smem + FS_OFF_VEC);
    float* sw2 = sb1 + 256;
    float* sb2 = sw2 + 256;
    float* sdot = reinterpret_cast<float*>(smem + FS_OFF_DOT);
    float* part = reinterpret_cast<float*>(smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + FS_OFF_BAR);   // [FS_STAGES]
    uint64_t* empty = full + FS_STAGES;                                // [FS_STAGES]
    uint64_t* acc_full = empty + FS_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const long long tile = blockIdx.x / FS_SPLIT;
    for (int i = tid; i < 256; i += (int)blockDim.x) { sb1[i] = net.fc1_b[i]; sw2[i] = net.fc2_w[i]; }
    if (tid == 0) {
        sb2[0] = net.fc2_b[0];
        for (int s = 0; s < FS_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t IDESC = make_idesc_f16(128, 256);

    if (warp == 0) {
        // ---------------- producer: this CTA's 16 squares of (A, B) K-stages
        if (lane == 0) {
            const uint8_t* a_tile = reinterpret_cast<const uint8_t*>(act2) + (tile << 20);
            for (int i = 0; i < FS_SQUARES; i++) {
                const int sq = (int)rank * FS_SQUARES + i;
                const int st = i % FS_STAGES;
                mbar_wait(empty + st, ((i / FS_STAGES) & 1) ^ 1);
                uint8_t* sa = smem + st * FS_STAGE_BYTES;
                mbar_arrive_expect_tx(full + st, FS_STAGE_BYTES);
                bulk_g2s(sa, a_tile + ((long long)sq << 14), FC_A_BYTES, full + st);
                bulk_g2s(sa + FC_A_BYTES, reinterpret_cast<const uint8_t*>(net.fc1p) + ((long long)sq << 15), FC_B_BYTES, full + st);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            for (int i = 0; i < FS_SQUARES; i++) {
                const int st = i % FS_STAGES;
                mbar_wait(full + st, (i / FS_STAGES) & 1);
                tc_fence_after();
                const uint32_t a = smem_u32(smem + st * FS_STAGE_BYTES), b = a + FC_A_BYTES;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    umma_f16(tmem, make_smem_desc(a + j * 2 * 2048, 2048, 128), make_smem_desc(b + j * 2 * 4096, 4096, 128), IDESC,
                             (i | j) != 0);
                umma_commit(empty + st);
            }
            umma_commit(acc_full);
        }
        __syncwarp();
    } else {
        // ---------------- epilogue, phase 1: park the partial sums (TMEM lane quarter = warp % 4, row = position)
        const int q = warp & 3, row = q * 32 + lane;
        mbar_wait(acc_full, 0);             // all UMMAs done: the operand ring is free
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 8; c++) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                reinterpret_cast<float4*>(part)[((c * 32 + j) >> 2) * 128 + row] =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
        tc_fence_before();
    }
    cluster_sync_all();                     // every CTA's partial sums are visible cluster-wide

    float hk[TRAIN ? 64 : 1];               // TRAIN: ReLU(fc1) of this CTA's 64 columns stays in registers for the backward head
    if (warp >= 2) {
        // ---------------- phase 2: fc1 outputs 64 rank .. 64 rank + 63 of all 128 positions
        const int q = warp & 3, row = q * 32 + lane;
        const long long pos = (tile << 7) + row;
        uint32_t pbase[FS_SPLIT];
#pragma unroll
        for (int k = 0; k < FS_SPLIT; k++) pbase[k] = dsmem_addr(smem_u32(part), (uint32_t)k) + (uint32_t)row * 16u;
        float dot = 0.f;
#pragma unroll (TRAIN ? 8 : 1)
        for (int c0 = 0; c0 < 64; c0 += 8) {
            float p[FS_SPLIT][8];
#pragma unroll
            for (int k = 0; k < FS_SPLIT; k++)
#pragma unroll
                for (int j = 0; j < 8; j += 4) {
                    const float4 t4 = dsmem_ld4(pbase[k] + (uint32_t)((((int)rank * 64 + c0 + j) >> 2) * 128 * 16));
                    p[k][j] = t4.x; p[k][j + 1] = t4.y; p[k][j + 2] = t4.z; p[k][j + 3] = t4.w;
                }
            float h[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int c = (int)rank * 64 + c0 + j;
                const float z = ((p[0][j] + p[1][j]) + p[2][j]) + p[3][j];       // rank order, whoever adds
                h[j] = fmaxf(z + sb1[c], 0.f);
                dot = fmaf(h[j], sw2[c], dot);
                if (TRAIN) hk[TRAIN ? c0 + j : 0] = h[j];
            }
            if (hidden && pos < n) {
                float4* hp = reinterpret_cast<float4*>(hidden + pos * 256 + (int)rank * 64 + c0);
                hp[0] = make_float4(h[0], h[1], h[2], h[3]);
                hp[1] = make_float4(h[4], h[5], h[6], h[7]);
            }
        }
        dsmem_st(dsmem_addr(smem_u32(sdot), 0u) + (uint32_t)(((int)rank * 128 + row) * 4), dot);
    }
    cluster_sync_all();                     // partial dot products have landed in CTA 0; nobody reads remote memory after this

    if (rank == 0 && warp >= 2) {
        const int q = warp & 3, row = q * 32 + lane;
        const long long pos = (tile << 7) + row;
        float dz2 = 0.f, ls = 0.f;
        if (pos < n) {
            const float v = tanhf((((sdot[row] + sdot[128 + row]) + sdot[256 + row]) + sdot[384 + row]) + sb2[0]);
            if (value) value[pos] = v;
            if (leaf) {
                // mcts.py:196-232: terminal short-circuits, else clamp(net, -1, 1)
                const unsigned f = flags[pos];
                float lv = fminf(1.f, fmaxf(-1.f, v));
                if (f & CDQ_F_CHECKMATE) lv = (f & CDQ_F_WHITE_TO_MOVE) ? -1.f : 1.f;
                else if (f & (CDQ_F_STALEMATE | CDQ_F_INSUFFICIENT | CDQ_F_REP3)) lv = 0.f;
                leaf[pos] = lv;
            }
            if (TRAIN) {
                float dl;                                                          // d loss / d q
                if (th.loss_kind == CDQ_LOSS_UPSTREAM) { dl = th.reward[pos]; ls = dl * v; }
                else {
                    const float y = th.reward[pos] + (1.f - th.done[pos]) * th.gamma * th.qn[pos];     // neural_network.py:238
                    const float d = v - y;
                    if (th.loss_kind == 0) { ls = d * d; dl = 2.f * d; }           // F.mse_loss (neural_network.py:241)
                    else { const float ad = fabsf(d); ls = ad < 1.f ? 0.5f * d * d : ad - 0.5f; dl = fminf(1.f, fmaxf(-1.f, d)); }
                }
                dz2 = dl * th.inv_n * (1.f - v * v);                               // through tanh
            }
        }
        if (TRAIN) {
            // this thread has read its four partial dots: their slots carry d loss / d z2 to the four CTAs
#pragma unroll
            for (int k = 0; k < FS_SPLIT; k++) dsmem_st(dsmem_addr(smem_u32(sdot), (uint32_t)k) + (uint32_t)(row * 4), dz2);
            float gb = dz2;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) { gb += __shfl_xor_sync(0xffffffffu, gb, d); ls += __shfl_xor_sync(0xffffffffu, ls, d); }
            if (lane == 0) { atomicAdd(th.g_fc2_b, gb); atomicAdd(th.loss, ls * th.inv_n); }
        }
    }
    if (TRAIN) {
        cluster_sync_all();                 // d loss / d z2 of the tile's 128 positions is in every CTA's shared memory
        float* red = reinterpret_cast<float*>(smem);          // [4 warps][128]: the partial-sum area is dead after the second barrier
        if (warp >= 2) {
            const int q = warp & 3, row = q * 32 + lane;
            const float dz2 = sdot[row];    // 0 for rows beyond n
            uint8_t* out = reinterpret_cast<uint8_t*>(th.dz1h) + (tile << 16) + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                float gw[32], gb1[32];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    float dz[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int cl = half * 32 + g * 8 + i, c = (int)rank * 64 + cl;
                        const float h = hk[TRAIN ? cl : 0];
                        dz[i] = h > 0.f ? dz2 * sw2[c] : 0.f;                      // through fc2 and ReLU
                        gw[g * 8 + i] = dz2 * h;
                        gb1[g * 8 + i] = dz[i];
                    }
                    __half2 p0 = __floats2half2_rn(dz[0] * th.scale, dz[1] * th.scale), p1 = __floats2half2_rn(dz[2] * th.scale, dz[3] * th.scale);
                    __half2 p2 = __floats2half2_rn(dz[4] * th.scale, dz[5] * th.scale), p3 = __floats2half2_rn(dz[6] * th.scale, dz[7] * th.scale);
                    // dz1h tile: [32 column groups of 8][16 row groups][8 rows][8 halfs]
                    *reinterpret_cast<uint4*>(out + ((int)rank * 8 + half * 4 + g) * 2048) =
                        make_uint4(*reinterpret_cast<unsigned*>(&p0), *reinterpret_cast<unsigned*>(&p1), *reinterpret_cast<unsigned*>(&p2),
                                   *reinterpret_cast<unsigned*>(&p3));
                }
                const float sw = warp_transpose_sum(gw, lane), sb = warp_transpose_sum(gb1, lane);   // lane l: column half * 32 + l
                red[q * 128 + half * 32 + lane] = sw;
                red[q * 128 + 64 + half * 32 + lane] = sb;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");    // the four epilogue warps
            const int t = q * 32 + lane;                     // 0..63: fc2.weight gradient, 64..127: fc1.bias gradient
            const float tot = (red[t] + red[128 + t]) + (red[256 + t] + red[384 + t]);
            atomicAdd((t < 64 ? th.g_fc2_w : th.g_fc1_b) + (int)rank * 64 + (t & 63), tot);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

bool fc_split_enabled() { return option(OPT_FC_SPLIT) != 0; }     // cdq_set_option("fc_split", 0): A/B against the persistent kernel

int launch_fc_head_split(const __half* act2, long long n, const PackedNet& net, float* value, const unsigned* flags,
                         float* leaf, float* hidden, cudaStream_t st, const TrainHead* th) {
    once_per_device(ONCE_FC_SPLIT, [] {
        cudaFuncSetAttribute(fc_head_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM);
        cudaFuncSetAttribute(fc_head_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_SMEM);
    });
    const long long tiles = (n + 127) / 128;
    ProfileScope ps(KK_FC, st);
    if (th) fc_head_split_kernel<true><<<(unsigned)(tiles * FS_SPLIT), FC_THREADS, FS_SMEM, st>>>(act2, n, net, value, flags, leaf, hidden, *th);
    else fc_head_split_kernel<false><<<(unsigned)(tiles * FS_SPLIT), FC_THREADS, FS_SMEM, st>>>(act2, n, net, value, flags, leaf, hidden, TrainHead{});
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
