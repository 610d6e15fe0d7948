// train_kernels.cu -- K3/K4: one DQN replay step (neural_network.py:218-252) on packed boards:
//   q  = net(states)              forward with the tcgen05 kernels, keeping ReLU(conv1), ReLU(conv2)
//   q' = target_net(next_states)  and ReLU(fc1) for the backward pass
//   y  = r + (1 - done) * gamma * q'          (neural_network.py:238)
//   loss = mean((q - y)^2)  (or Huber)        (neural_network.py:241)
//   backward through tanh, fc2, fc1, conv2, conv1; Adam (neural_network.py:65,244-246)
//
// Every GEMM-shaped piece runs on tcgen05: the forward (conv_kernel.cu, fc_kernel.cu), fc1's data and
// weight gradients (fc_bwd_kernel.cu) and conv2's data/weight/bias gradients (conv_bwd2_kernel.cu).
// This file holds the glue: the loss/tanh/fc2 head, conv1's weight gradient (a scatter, because its
// input is one-hot), Adam, and the workspace carve-up.  Gradients flow between kernels as fp16
// scaled by 64*n (a power of two for the usual batch sizes) and are unscaled in fp32 epilogues;
// weight gradients are accumulated in PyTorch layout with atomicAdd.
#include "common.cuh"
#include "cnn.cuh"
#include <cmath>
#include <cstdlib>

namespace cdq {

// ------------------------------------------------------------------------------------------
// loss + backward through tanh, fc2 and fc1's ReLU; one warp per position (grid-stride).  Writes
// dz1h (fc_bwd_kernel.cu): fp16, scaled by `scale`, tiled [128-position tile][32 chunks][16 row
// groups][8 positions][8 halfs]; rows beyond n are written as zeros (the weight gradient contracts
// over whole tiles).  Lane l owns fc1 outputs 8l .. 8l+7.
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ q, const float* __restrict__ qn, const float* __restrict__ reward,
                const float* __restrict__ done, const float* __restrict__ hidden, const float* __restrict__ fc2_w,
                long long n, float inv_n /* 1 / mean_over */, float gamma, int loss_kind, float scale, __half* __restrict__ dz1h,
                float* __restrict__ g_fc2_w, float* __restrict__ g_fc2_b, float* __restrict__ g_fc1_b,
                float* __restrict__ loss) {
    __shared__ float s_gw[8][256];
    __shared__ float s_gb1[8][256];
    __shared__ float s_misc[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gw[8], gb1[8], gb = 0.f, ls = 0.f, w2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { gw[i] = 0.f; gb1[i] = 0.f; w2[i] = fc2_w[lane * 8 + i]; }
    const long long n_pad = ((n + 127) >> 7) << 7;
    for (long long b = (long long)blockIdx.x * 8 + warp; b < n_pad; b += (long long)gridDim.x * 8) {
        uint4 packed = make_uint4(0, 0, 0, 0);
        if (b < n) {
            const float qv = q[b];
            float dl;                                                          // d loss / d q
            if (loss_kind == CDQ_LOSS_UPSTREAM) { dl = reward[b]; ls += dl * qv; }      // caller supplies d loss / d q (autograd)
            else {
                const float y = reward[b] + (1.f - done[b]) * gamma * qn[b];   // neural_network.py:238
                const float d = qv - y;
                if (loss_kind == 0) { ls += d * d; dl = 2.f * d; }             // F.mse_loss (neural_network.py:241)
                else { const float ad = fabsf(d); ls += ad < 1.f ? 0.5f * d * d : ad - 0.5f; dl = fminf(1.f, fmaxf(-1.f, d)); }
            }
            const float dz2 = dl * inv_n * (1.f - qv * qv);                    // through tanh
            gb += dz2;
            const float4 h0 = *reinterpret_cast<const float4*>(hidden + b * 256 + lane * 8);
            const float4 h1 = *reinterpret_cast<const float4*>(hidden + b * 256 + lane * 8 + 4);
            const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            float dz[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                gw[i] = fmaf(dz2, h[i], gw[i]);
                dz[i] = h[i] > 0.f ? dz2 * w2[i] : 0.f;                        // through fc2 and ReLU
                gb1[i] += dz[i];
            }
            __half2 p0 = __floats2half2_rn(dz[0] * scale, dz[1] * scale), p1 = __floats2half2_rn(dz[2] * scale, dz[3] * scale);
            __half2 p2 = __floats2half2_rn(dz[4] * scale, dz[5] * scale), p3 = __floats2half2_rn(dz[6] * scale, dz[7] * scale);
            packed = make_uint4(*reinterpret_cast<unsigned*>(&p0), *reinterpret_cast<unsigned*>(&p1),
                                *reinterpret_cast<unsigned*>(&p2), *reinterpret_cast<unsigned*>(&p3));
        }
        const int pr = (int)(b & 127);
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(dz1h) + (b >> 7) * 65536 + lane * 2048 + (pr >> 3) * 128 + (pr & 7) * 16) = packed;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { s_gw[warp][lane * 8 + i] = gw[i]; s_gb1[warp][lane * 8 + i] = gb1[i]; }
    if (lane == 0) { s_misc[warp][0] = gb; s_misc[warp][1] = ls; }        // identical in all lanes
    __syncthreads();
    float t = 0.f, t1 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) { t += s_gw[w][threadIdx.x]; t1 += s_gb1[w][threadIdx.x]; }
    atomicAdd(g_fc2_w + threadIdx.x, t);
    atomicAdd(g_fc1_b + threadIdx.x, t1);
    if (threadIdx.x == 0) {
        float b0 = 0.f, l0 = 0.f;
        for (int w = 0; w < 8; w++) { b0 += s_misc[w][0]; l0 += s_misc[w][1]; }
        atomicAdd(g_fc2_b, b0);
        atomicAdd(loss, l0 * inv_n);
    }
}

// conv1 weight + bias gradient.  The input is one-hot, so the GEMM collapses to a gather:
// dW1[co][p][tap] += dz1c[b][sq][co] for every piece p standing on square sq + tap.  A board is handled by
// FOUR warps that split the PIECE TYPES (white pawns / black pawns / the other white pieces / the other black pieces:
// 8 + 8 + 8 + 8 pieces on a full board; three warps with 11 + 11 + 10 until round 2c -- the gather runs at the pace of a
// board's busiest warp and at 44 % of the issue slots, so more and shorter warps pay twice), lane = output channel, and a
// lane's accumulators (its types x 9 taps) live in registers.  Per piece a warp pays the bit scan and the
// address arithmetic once and then reads the nine output squares around the piece -- conflict-free
// shared-memory reads (the previous split, one kernel row per warp, made every warp walk ALL pieces: three
// times the per-piece overhead, which was two thirds of the instructions).  The board's 8 KB of gradients
// and its bitboards are brought into shared memory with cp.async, double buffered (the copy of the slot's
// next board runs under the gather of the current one and costs no registers).  Six such board slots per
// CTA, one CTA per SM, each slot synchronised by its own named barrier.  At the end the slots of a CTA are
// summed through shared memory into PyTorch order and added to the gradient with coalesced 16-byte
// reductions: same-line atomics serialise in L2, so what matters is the number of transactions per 128-byte
// line (one scalar atomic per cell and CTA with the channel as the fast index put 32 per CTA on each line
// and cost more than the gather itself).
constexpr int C1W_SLOTS = 6, C1W_SUBS = 4, C1W_WARPS = C1W_SUBS * C1W_SLOTS, C1W_THREADS = 32 * C1W_WARPS;
constexpr int C1W_STAGE_BYTES = C1W_SLOTS * 2 * 8192;             // [slot][2][64 squares][32 channels] fp32
constexpr int C1W_ROWS = 5 * 9 + 1;                               // per warp: up to 5 types x 9 taps, + the bias row
constexpr int C1W_RED_BYTES = C1W_WARPS * C1W_ROWS * 32 * 4;      // reduction buffer, aliases the staging buffers
constexpr int C1W_OFF_BOARD = C1W_RED_BYTES > C1W_STAGE_BYTES ? C1W_RED_BYTES : C1W_STAGE_BYTES;   // [slot][2][8 bitboards]
constexpr int C1W_OFF_OUT = C1W_OFF_BOARD + C1W_SLOTS * 2 * 64;   // [32 co][108] summed over the slots
constexpr int C1W_SMEM = C1W_OFF_OUT + 32 * 108 * 4;
static_assert(C1W_SMEM <= 227 * 1024, "conv1 gather shared memory");
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// piece types (channel = type + 6 * black; pawn, knight, bishop, rook, queen, king: board_utils.py:17-30) of the four warp roles
__host__ __device__ constexpr int c1w_ntypes(int g) { return g < 2 ? 1 : 5; }
__host__ __device__ constexpr int c1w_type(int g, int i) { return g == 0 ? 0 : g == 1 ? 6 : g == 2 ? 1 + i : 7 + i; }
__constant__ int c1w_group_of[12] = {0, 2, 2, 2, 2, 2, 1, 3, 3, 3, 3, 3};
__constant__ int c1w_local_of[12] = {0, 0, 1, 2, 3, 4, 0, 0, 1, 2, 3, 4};

template <int G>
__device__ __forceinline__ void c1w_gather(const CdqBoard* __restrict__ boards, const float* __restrict__ dz1c, long long n,
                                           float* c1w_smem, uint8_t* c1w_raw, int slot, int sub, int lane, int warp) {
    constexpr int NT = c1w_ntypes(G);
    const int t128 = sub * 32 + lane;
    float acc[NT][9];
#pragma unroll
    for (int p = 0; p < NT; p++)
#pragma unroll
        for (int t = 0; t < 9; t++) acc[p][t] = 0.f;
    float bias = 0.f;
    float* stage0 = c1w_smem + slot * 4096;             // two buffers of [64 squares][32 channels]
    unsigned long long* board0 = reinterpret_cast<unsigned long long*>(c1w_raw + C1W_OFF_BOARD) + slot * 16;
    const long long stride = (long long)gridDim.x * C1W_SLOTS;
    auto fetch = [&](long long bb, int buf) {           // this thread's share of board bb: 512 16-byte pieces over 128 threads
        if (bb < n) {
            const float* src = dz1c + bb * 2048;
            float* dst = stage0 + buf * 2048;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int u = i * 128 + t128;
                cp_async16(dst + 4 * u, src + 4 * u);
            }
            if (t128 < 8) cp_async8(board0 + buf * 8 + t128, reinterpret_cast<const unsigned long long*>(boards) + bb * 9 + t128);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    long long b = (long long)blockIdx.x * C1W_SLOTS + slot;
    fetch(b, 0);
    for (int it = 0; b < n; b += stride, it++) {
        const int buf = it & 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory");  // all four warps: copies landed, previous board done
        fetch(b + stride, buf ^ 1);
        const float* stage = stage0 + buf * 2048;
        const unsigned long long* r = board0 + buf * 8;
        {   // bias: every thread sums a quarter of the squares of its channel
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                s0 += stage[(sub + 4 * i) * 32 + lane];
                s1 += stage[(sub + 4 * i + 4) * 32 + lane];
            }
            bias += s0 + s1;
        }
        const unsigned long long white = r[6], black = r[7] & ~white;
#pragma unroll
        for (int pi = 0; pi < NT; pi++) {
            const int p = c1w_type(G, pi);
            unsigned long long m = r[p % 6] & (p < 6 ? white : black);
            while (m) {                                 // warp-uniform
                const int s = __ffsll((long long)m) - 1;
                m &= m - 1;
                // the piece on (y, x) is read through tap (dy, dx) by output square (y - dy, x - dx); tap = 3 (dy + 1) + (dx + 1)
                const int y = s >> 3, x = s & 7;
                const float* c = stage + s * 32 + lane;
                const bool up = y < 7, dn = y > 0, rt = x < 7, lf = x > 0;     // output square one rank up / down, one file right / left exists
                if (up) {                               // dy = -1: output rank y + 1
                    if (rt) acc[pi][0] += c[288];
                    acc[pi][1] += c[256];
                    if (lf) acc[pi][2] += c[224];
                }
                if (rt) acc[pi][3] += c[32];            // dy = 0
                acc[pi][4] += c[0];
                if (lf) acc[pi][5] += c[-32];
                if (dn) {                               // dy = +1: output rank y - 1
                    if (rt) acc[pi][6] += c[-224];
                    acc[pi][7] += c[-256];
                    if (lf) acc[pi][8] += c[-288];
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                                    // all slots done: the staging buffers become the reduction buffer
    float* mine = c1w_smem + warp * (C1W_ROWS * 32);
#pragma unroll
    for (int pi = 0; pi < NT; pi++)
#pragma unroll
        for (int t = 0; t < 9; t++) mine[(pi * 9 + t) * 32 + lane] = acc[pi][t];
    mine[(C1W_ROWS - 1) * 32 + lane] = bias;
}

__global__ void __launch_bounds__(C1W_THREADS, 1)
conv1_wgrad_kernel(const CdqBoard* __restrict__ boards, const float* __restrict__ dz1c, long long n,
                   float* __restrict__ g_w, float* __restrict__ g_b) {
    extern __shared__ __align__(16) uint8_t c1w_raw[];
    float* c1w_smem = reinterpret_cast<float*>(c1w_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp >> 2, sub = warp & 3;
    if (sub == 0) c1w_gather<0>(boards, dz1c, n, c1w_smem, c1w_raw, slot, sub, lane, warp);
    else if (sub == 1) c1w_gather<1>(boards, dz1c, n, c1w_smem, c1w_raw, slot, sub, lane, warp);
    else if (sub == 2) c1w_gather<2>(boards, dz1c, n, c1w_smem, c1w_raw, slot, sub, lane, warp);
    else c1w_gather<3>(boards, dz1c, n, c1w_smem, c1w_raw, slot, sub, lane, warp);
    __syncthreads();
    // ---- sum the slots (channel = fast index: conflict free) into PyTorch order, then coalesced 16-byte reductions
    float* out = reinterpret_cast<float*>(c1w_raw + C1W_OFF_OUT);
    for (int i = threadIdx.x; i < 108 * 32; i += C1W_THREADS) {
        const int co = i & 31, cell = i >> 5, p = cell / 9, tap = cell - p * 9;
        const int g = c1w_group_of[p], row = c1w_local_of[p] * 9 + tap;
        float v = 0.f;
#pragma unroll
        for (int sl = 0; sl < C1W_SLOTS; sl++) v += c1w_smem[(sl * C1W_SUBS + g) * (C1W_ROWS * 32) + row * 32 + co];
        out[co * 108 + cell] = v;                                                 // conv1.weight[co][p][tap]
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 108 / 4; i += C1W_THREADS)
#ifndef CB2_NO_ATOM
        atomicAdd(reinterpret_cast<float4*>(g_w) + i, reinterpret_cast<const float4*>(out)[i]);
#else
        if (out[4 * i] == 12345.f) g_w[i] = 0.f;
#endif
    if (threadIdx.x < 32) {
        float v = 0.f;
        for (int w = 0; w < C1W_WARPS; w++) v += c1w_smem[w * (C1W_ROWS * 32) + (C1W_ROWS - 1) * 32 + threadIdx.x];
        atomicAdd(g_b + threadIdx.x, v);
    }
}

// ReLU(conv1) from the forward's position-contiguous dump [tile of 16][file 8][chunk 4][rank 8][16 positions][8 halfs] to the
// zero-bordered boards conv2's backward fetches with one bulk copy per unit: [unit of 4 positions][10 ranks][4 chunks]
// [4 positions][10 files][8 halfs].  One CTA per (tile, rank, chunk): 2 KB through shared memory, 256-byte runs in, 128-byte
// runs out.  Runs on the side stream under the fc forward.
__global__ void __launch_bounds__(128)
a1_boards_kernel(const uint4* __restrict__ dump, uint4* __restrict__ boards) {
    __shared__ uint4 cell[16][9];                       // [position][file], padded
    const int kc = blockIdx.x & 3, y = (blockIdx.x >> 2) & 7;
    const long long tile = blockIdx.x >> 5;
    const int t = threadIdx.x;
    {
        const int f = t >> 4, p = t & 15;               // 16 consecutive threads = 256 contiguous bytes
        cell[p][f] = dump[tile * 4096 + f * 512 + kc * 128 + y * 16 + p];
    }
    __syncthreads();
    {
        const int p = t >> 3, f = t & 7;                // 8 consecutive threads = the 8 interior files of a board row
        boards[(tile * 4 + (p >> 2)) * 1600 + (y + 1) * 160 + kc * 40 + (p & 3) * 10 + f + 1] = cell[p][f];
    }
}

// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad); grads are pre-multiplied by grad_scale
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    // torch: denom = sqrt(v)/sqrt(bias_correction2) + eps ; p -= (lr / bias_correction1) * m / denom
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
}

int launch_adam(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float b1, float b2,
                float eps, float grad_scale, cudaStream_t st) {
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
    ProfileScope ps(KK_ADAM, st);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
    count_launch();
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Replay minibatch on the device (random.sample + torch.stack of DQNAgent.train, neural_network.py:224-229).  The draw is a
// pseudo-random PERMUTATION of the logical indices [0, len): a four-round Feistel network on b bits (2^b >= len) with
// cycle walking, keyed by (seed, draw counter); its first `batch` images are `batch` distinct, uniformly drawn entries, every
// thread computes its own, and a host restatement reproduces them (tests/test_gpu_train.py).
__host__ __device__ inline unsigned long long replay_mix(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__host__ __device__ inline unsigned replay_permute(unsigned k, unsigned len, unsigned long long key) {
    unsigned half = 1;                                   // bits per Feistel half
    while ((1u << (2 * half)) < len) half++;
    const unsigned mask = (1u << half) - 1u;
    unsigned x = k;
    do {
        unsigned l = x >> half, r = x & mask;
        for (int round = 0; round < 4; round++) {
            const unsigned f = (unsigned)replay_mix(key + ((unsigned long long)round << 32) + r) & mask;
            const unsigned t = l ^ f;
            l = r; r = t;
        }
        x = (l << half) | r;
    } while (x >= len);                                  // cycle walking keeps it a permutation of [0, len)
    return x;
}
__global__ void __launch_bounds__(256)
replay_sample_kernel(const unsigned long long* __restrict__ mem_states, const unsigned long long* __restrict__ mem_next,
                     const float* __restrict__ mem_reward, const float* __restrict__ mem_done, const int* __restrict__ ring,
                     unsigned long long seed, unsigned* __restrict__ counter, long long batch, unsigned long long* __restrict__ states,
                     unsigned long long* __restrict__ next, float* __restrict__ reward, float* __restrict__ done) {
    const int start = ring[0], len = ring[1], cap = ring[2];
    const unsigned long long key = replay_mix(seed ^ replay_mix((unsigned long long)*counter));
    // one thread per (sample, word): 9 words of the state record, 9 of the next state, reward + done ride with word 0
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long k = t / 9;
    const int word = (int)(t - k * 9);
    if (k < batch && len > 0) {
        const int slot = (int)(((long long)start + replay_permute((unsigned)k, (unsigned)len, key)) % cap);
        states[k * 9 + word] = mem_states[(long long)slot * 9 + word];
        next[k * 9 + word] = mem_next[(long long)slot * 9 + word];
        if (word == 0) { reward[k] = mem_reward[slot]; done[k] = mem_done[slot]; }
    }
    // the draw counter advances once per launch, after every CTA has read it: last block out (ring[3] counts arrivals)
    __syncthreads();
    if (threadIdx.x == 0) {
        int* arrivals = const_cast<int*>(ring) + 3;
        if (atomicAdd(arrivals, 1) == (int)gridDim.x - 1) { *arrivals = 0; *counter += 1u; }
    }
}
int launch_replay_sample(const CdqBoard* mem_states, const CdqBoard* mem_next, const float* mem_reward, const float* mem_done,
                         const int* ring, unsigned long long seed, unsigned* counter, long long batch, CdqBoard* states, CdqBoard* next,
                         float* reward, float* done, cudaStream_t st) {
    ProfileScope ps(KK_TRAIN, st);
    replay_sample_kernel<<<(unsigned)((batch * 9 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long*>(mem_states), reinterpret_cast<const unsigned long long*>(mem_next), mem_reward, mem_done,
        ring, seed, counter, batch, reinterpret_cast<unsigned long long*>(states), reinterpret_cast<unsigned long long*>(next), reward, done);
    count_launch();
    return (int)cudaGetLastError();
}
unsigned replay_permute_host(unsigned k, unsigned len, unsigned long long seed, unsigned counter) {
    return replay_permute(k, len, replay_mix(seed ^ replay_mix((unsigned long long)counter)));
}

// ------------------------------------------------------------------------------------------
// workspace carve-up (bytes per position; everything 256-byte aligned per section)
struct TrainWs {
    __half* act2_tiles; __half* act2_target; __half* dz1h; __half* a1_dump; __half* a1_boards; __half* dz2_boards; unsigned* sched; unsigned* relu_mask; float* hidden; float* q; float* qn; float* dz1c;
    size_t zero_from, zero_bytes;   // [a1_boards, dz2_boards]: borders must be zero, cleared every step
    size_t total;
};
static TrainWs carve(void* base, long long n) {
    TrainWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? (uint8_t*)base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    const long long tiles = (n + 127) / 128, groups = tiles * 16;
    w.act2_tiles = (__half*)take((size_t)tiles << 20);
    w.act2_target = (__half*)take((size_t)tiles << 20);
    w.dz1h = (__half*)take((size_t)tiles << 16);
    w.a1_dump = (__half*)take((size_t)tiles * 8 * 65536);      // ReLU(conv1) as the forward writes it (64 KB per 16 positions)
    w.zero_from = off;
    w.a1_boards = (__half*)take((size_t)groups * 51200);
    w.dz2_boards = (__half*)take((size_t)groups * 102400);
    w.sched = (unsigned*)take(256);           // conv2 backward's work counters: zero at entry (they reset themselves)
    w.zero_bytes = off - w.zero_from;
    w.relu_mask = (unsigned*)take((size_t)tiles * 16 * 8 * 128 * 4);      // 1 bit per ReLU(conv2) output
    w.hidden = (float*)take((size_t)n * 256 * 4);
    w.q = (float*)take((size_t)n * 4);
    w.qn = (float*)take((size_t)n * 4);
    w.dz1c = (float*)take((size_t)n * 64 * 32 * 4);
    w.total = off;
    return w;
}
size_t train_workspace_bytes(long long n) { return n > 0 ? carve(nullptr, n).total : 0; }

// Independent pieces of the step run on a side stream (fork/join with events, which a capturing
// stream turns into graph edges): the target network's forward next to the online forward, and
// fc1's weight gradient next to the data-gradient chain.  At batch 4 096 every kernel here fills a
// fraction of the 148 SMs, so the critical path, not the sum, sets the step time.
namespace {
struct SideStream {
    cudaStream_t st = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, mid = nullptr, join2 = nullptr, dgrad_done = nullptr, target_done = nullptr;
    int init() {
        if (st) return 0;
        CDQ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CDQ_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CDQ_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        CDQ_CUDA(cudaEventCreateWithFlags(&mid, cudaEventDisableTiming));
        CDQ_CUDA(cudaEventCreateWithFlags(&join2, cudaEventDisableTiming));
        CDQ_CUDA(cudaEventCreateWithFlags(&dgrad_done, cudaEventDisableTiming));
        CDQ_CUDA(cudaEventCreateWithFlags(&target_done, cudaEventDisableTiming));
        return 0;
    }
};
SideStream g_sides[MAX_DEVICES];      // one per device: the stream and its events belong to the device they were created on
std::mutex g_side_mu[MAX_DEVICES];    // serialises the ENQUEUE of concurrent training calls on one device (they share the events)
} // namespace

int train_fwd_bwd(const CdqNet* net, const CdqNet* target, const CdqWeights& w, const CdqBoard* states,
                  const CdqBoard* next_states, const float* reward, const float* done, long long n, double mean_over,
                  float gamma, int loss_kind_flags, const CdqWeights& g /* gradient outputs, pre-zeroed */, float* loss,
                  void* ws_base, cudaStream_t st, const CdqOptimizer* opt) {
    TrainWs ws = carve(ws_base, n);
    const int loss_kind = loss_kind_flags & 0xff;
    int e;
    const int dev = current_device();
    if (dev < 0) return CDQ_E_NODEVICE;
    std::lock_guard<std::mutex> side_lock(g_side_mu[dev]);
    SideStream& g_side = g_sides[dev];
    if ((e = g_side.init())) return e;
    cudaStream_t side = g_side.st;
    // borders of the board-shaped buffers must be zero; their interiors are rewritten every step
    if (!(loss_kind_flags & CDQ_TRAIN_WS_PERSISTENT))
        CDQ_CUDA(cudaMemsetAsync((uint8_t*)ws_base + ws.zero_from, 0, ws.zero_bytes, st));
    // ---- forward: target value of next_states (inference, side stream) next to q of states (keeps activations)
    CDQ_CUDA(cudaEventRecord(g_side.fork, st));
    CDQ_CUDA(cudaStreamWaitEvent(side, g_side.fork, 0));
    // The two conv launches are persistent, one CTA per SM, and cannot share an SM (170 KB of shared memory each):
    // on the whole machine they would run one after the other, each with its own prologue and a half-empty
    // last round, so each gets half of the SMs and they run side by side.
    // The online launch also writes the ReLU(conv1) boards and sits on the longer chain (repack -> conv -> fc),
    // so it gets the larger share (CDQ_TRAIN_CONV_SPLIT = percent of the SMs for the online network, default 50: measured flat between 50 and 61, profiles/conv_split_sweep.sh; with the backward head fused into the online fc kernel 36 / 42 / 50 % give 165.7 / 162.8 / 161.7 us, profiles/README.md round 2).
    static const int online_pct = [] { const char* e = getenv("CDQ_TRAIN_CONV_SPLIT"); const int v = e ? atoi(e) : 0; return v >= 20 && v <= 80 ? v : 50; }();
    const int online_sms = device_sm_count() * online_pct / 100, target_sms = device_sm_count() - online_sms;
    const bool upstream = loss_kind == CDQ_LOSS_UPSTREAM;     // no target network: d loss / d q comes from the caller
    // cdq_train_step (opt != NULL): `net` is already packed from the current parameters (the previous step's tail did it), so the
    // online forward starts at once; the loss accumulator is cleared on the side stream in front of the target forward
    // Two schedules for the optimizer (option "train_split_optimizer"):
    //   0  the parameters are repacked at the head of the step and ONE optimizer launch closes it;
    //   1  fc1.weight's optimizer launch and repack follow its weight gradient on the side stream, the rest closes the step and
    //      the next step starts with the convolution at once.  The step is bound by SM-time, so 1 only moves work around: on
    //      one GPU 0.1467 against 0.1455 ms (0.1482 against 0.1475 with conv2's backward capped at 168 registers so that an
    //      optimizer CTA fits next to it), on eight GPUs the same span (the second launch keeps its handshake floor).
    const bool split_opt = opt && option(OPT_TRAIN_SPLIT_OPT) != 0;
    if (opt && !split_opt) loss_kind_flags |= CDQ_TRAIN_REPACK;
    if (split_opt) {
        loss_kind_flags &= ~CDQ_TRAIN_REPACK;
        if (loss_kind_flags & CDQ_TRAIN_ZERO_LOSS) CDQ_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), side));
        loss_kind_flags &= ~CDQ_TRAIN_ZERO_LOSS;
    }
    if (!upstream) {
        if ((e = launch_conv_stack(next_states, n, target->p, ws.act2_target, nullptr, side, nullptr, target_sms))) return e;
        if ((e = launch_fc_head(ws.act2_target, n, target->p, ws.qn, nullptr, nullptr, nullptr, side))) return e;
        CDQ_CUDA(cudaEventRecord(g_side.target_done, side));
    }
    // the online network's fp16 tiles are rebuilt from the fp32 parameters next to the target forward
    const bool zero_in_pack = (loss_kind_flags & CDQ_TRAIN_REPACK) && (loss_kind_flags & CDQ_TRAIN_ZERO_LOSS);
    if ((loss_kind_flags & CDQ_TRAIN_ZERO_LOSS) && !zero_in_pack) CDQ_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if ((loss_kind_flags & CDQ_TRAIN_REPACK) && (e = launch_pack_weights(w, net->p, st, zero_in_pack ? loss : nullptr))) return e;
    if ((e = launch_conv_stack(states, n, net->p, ws.act2_tiles, nullptr, st, ws.a1_dump, upstream ? 0 : online_sms))) return e;
    // ReLU(conv2)'s bit mask for the fc1 data gradient: side stream, after the target forward, under the online fc head
    CDQ_CUDA(cudaEventRecord(g_side.mid, st));
    CDQ_CUDA(cudaStreamWaitEvent(side, g_side.mid, 0));
    if ((e = launch_relu_mask(ws.act2_tiles, n, ws.relu_mask, side))) return e;
    CDQ_CUDA(cudaEventRecord(g_side.join, side));
    {   // ReLU(conv1) boards for conv2's backward: after the join the head waits for, needed only by conv2_bwd2
        ProfileScope ps(KK_TRAIN, side);
        const long long tiles16 = ((n + 127) >> 7) * 8;
        a1_boards_kernel<<<(unsigned)(tiles16 * 32), 128, 0, side>>>(reinterpret_cast<const uint4*>(ws.a1_dump), reinterpret_cast<uint4*>(ws.a1_boards));
        count_launch();
    }
    CDQ_CUDA(cudaEventRecord(g_side.join2, side));
    // lift into fp16's normal range; undone in the fp32 epilogues (upstream mode: the caller normalises max|dl| to 1)
    const float scale = upstream ? 64.f : 64.f * (float)mean_over;
    // ---- loss, tanh, fc2, fc1 bias: for batches the cluster split-K fc kernel handles (<= 8 192 positions) this head of the
    //      backward pass is that kernel's epilogue (it needs q' of the target network, so it starts after the target forward);
    //      larger batches keep the persistent fc kernel + head_bwd_kernel
    const bool fused_head = ((n + 127) >> 7) <= FC_SPLIT_MAX_TILES && fc_split_enabled();
    if (fused_head) {
        if (!upstream) CDQ_CUDA(cudaStreamWaitEvent(st, g_side.target_done, 0));
        TrainHead th{ws.qn, reward, done, gamma, upstream ? 1.f : (float)(1.0 / mean_over), scale, loss_kind, ws.dz1h,
                     (float*)g.fc2_w, (float*)g.fc2_b, (float*)g.fc1_b, loss};
        if ((e = launch_fc_head_split(ws.act2_tiles, n, net->p, ws.q, nullptr, nullptr, nullptr, st, &th))) return e;
        CDQ_CUDA(cudaStreamWaitEvent(st, g_side.join, 0));
    } else {
        if ((e = launch_fc_head(ws.act2_tiles, n, net->p, ws.q, nullptr, nullptr, ws.hidden, st))) return e;
        CDQ_CUDA(cudaStreamWaitEvent(st, g_side.join, 0));
        ProfileScope ps(KK_HEAD_BWD, st);
        const long long n_pad = ((n + 127) >> 7) << 7;
        const unsigned blocks = (unsigned)(n_pad / 8 < 1184 ? n_pad / 8 : 1184);
        head_bwd_kernel<<<blocks, 256, 0, st>>>(ws.q, ws.qn, reward, done, ws.hidden, w.fc2_w, n, upstream ? 1.f : (float)(1.0 / mean_over),
                                                gamma, loss_kind, scale, ws.dz1h,
                                                (float*)g.fc2_w, (float*)g.fc2_b, (float*)g.fc1_b, loss);
        count_launch();
    }
    // ---- fc1 weight and data gradients, conv2 weight/bias/data gradients: tcgen05
    CDQ_CUDA(cudaEventRecord(g_side.fork, st));
    CDQ_CUDA(cudaStreamWaitEvent(side, g_side.fork, 0));
    if ((e = launch_fc1_wgrad(ws.dz1h, ws.act2_tiles, n, 1.f / scale, (float*)g.fc1_w, side))) return e;
    if ((e = launch_fc1_dgrad(ws.dz1h, ws.relu_mask, net->p.fc1tp, n, ws.dz2_boards, st))) return e;
    if (split_opt) {
        // fc1.weight (98 % of the parameters): optimizer -- with its reduce-scatter / all-gather under a peer group -- and the
        // repack of fc1's two tile sets follow the weight gradient on the side stream, under conv2's and conv1's backward.
        // The repack overwrites the tiles fc1's data gradient reads, so it waits for that kernel.
        CDQ_CUDA(cudaEventRecord(g_side.dgrad_done, st));
        if ((e = launch_dp_adam(opt->peers, opt->rank, opt->world, opt->d_m, opt->d_v, opt->n_params_padded, opt->d_step, opt->d_grid_bar,
                                opt->lr, opt->beta1, opt->beta2, opt->eps, side, 1))) return e;
        CDQ_CUDA(cudaStreamWaitEvent(side, g_side.dgrad_done, 0));
        if ((e = launch_pack_weights(w, net->p, side, nullptr, 1))) return e;
    }
    CDQ_CUDA(cudaEventRecord(g_side.join, side));
    CDQ_CUDA(cudaStreamWaitEvent(st, g_side.join2, 0));
    if ((e = launch_conv2_bwd(ws.dz2_boards, ws.a1_boards, net->p.w2tp, n, 1.f / scale, (float*)g.conv2_w, (float*)g.conv2_b,
                              ws.dz1c, ws.sched, st))) return e;
    {
        // ---- conv1: weight + bias gradient by scatter over the one-hot input
        ProfileScope ps(KK_CONV1_WGRAD, st);
        once_per_device(ONCE_CONV1_WGRAD, [] { cudaFuncSetAttribute(conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1W_SMEM); });
        const long long want = (n + C1W_SLOTS - 1) / C1W_SLOTS;     // one CTA of 24 warps per SM
        const unsigned blocks = (unsigned)(want < device_sm_count() ? want : device_sm_count());
        conv1_wgrad_kernel<<<blocks, C1W_THREADS, C1W_SMEM, st>>>(states, ws.dz1c, n, (float*)g.conv1_w, (float*)g.conv1_b);
        count_launch();
    }
    if (split_opt) {
        // the remaining 22 497 parameters (conv1, conv2, biases, fc2) and their tiles
        if ((e = launch_dp_adam(opt->peers, opt->rank, opt->world, opt->d_m, opt->d_v, opt->n_params_padded, opt->d_step, opt->d_grid_bar,
                                opt->lr, opt->beta1, opt->beta2, opt->eps, st, 2))) return e;
        if ((e = launch_pack_weights(w, net->p, st, nullptr, 2))) return e;
    }
    CDQ_CUDA(cudaStreamWaitEvent(st, g_side.join, 0));
    if (opt && !split_opt) {
        if ((e = launch_dp_adam(opt->peers, opt->rank, opt->world, opt->d_m, opt->d_v, opt->n_params_padded, opt->d_step, opt->d_grid_bar,
                                opt->lr, opt->beta1, opt->beta2, opt->eps, st, 0))) return e;
    }
    return (int)cudaGetLastError();
}

} // namespace cdq
