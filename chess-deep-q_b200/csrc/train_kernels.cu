// train_kernels.cu -- K3/K4: one DQN replay step (neural_network.py:218-252) on packed boards:
//   q  = net(states)              forward with the tcgen05 kernels, keeping ReLU(conv1), ReLU(conv2)
//   q' = target_net(next_states)  and ReLU(fc1) for the backward pass
//   y  = r + (1 - done) * gamma * q'          (neural_network.py:238)
//   loss = mean((q - y)^2)  (or Huber)        (neural_network.py:241)
//   backward through tanh, fc2, fc1, conv2, conv1; Adam (neural_network.py:65,244-246)
//
// Round-1 state of the backward pass: correct and all on the device, but the GEMMs still run on a
// shared-memory-tiled fp32 SIMT kernel (gemm_kernel below) over im2col buffers; moving them to
// tcgen05 with the board-in-shared-memory operands of conv_kernel.cu is the next step.
// Layout conventions: activations are NHWC-like, [position][square][channel]; weight gradients are
// written in PyTorch order (the epilogues permute).
#include "common.cuh"
#include "cnn.cuh"
#include <cmath>

namespace cdq {

// ------------------------------------------------------------------------------------------
// C[M][N] (+)= A x B with fp32 accumulation.  A is stored [K][M] (A_KM) or [M][K]; B is [K][N].
// blockIdx.z splits K (results combined with atomicAdd when the epilogue allows it).
enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_FC1_DGRAD = 2, EPI_FC1_WGRAD = 3, EPI_CONV2_WGRAD = 4, EPI_CONV1_WGRAD = 5 };

struct GemmAux {
    const __half* mask;   // EPI_FC1_DGRAD: ReLU(conv2) in [position][square][channel] order
};

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }

template <typename TA, typename TB, bool A_KM, int BM, int BN, int EPI>
__global__ void __launch_bounds__(256)
gemm_kernel(const TA* __restrict__ A, const TB* __restrict__ B, float* __restrict__ C, int M, int N, long long K,
            long long lda, long long ldb, long long ldc, long long k_chunk, GemmAux aux) {
    constexpr int BK = 16, TM = BM / 16, TN = BN / 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const long long k_begin = (long long)blockIdx.z * k_chunk;
    const long long k_end = k_begin + k_chunk < K ? k_begin + k_chunk : K;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    for (long long k0 = k_begin; k0 < k_end; k0 += BK) {
        for (int idx = tid; idx < BK * BM; idx += 256) {
            int k, m;
            if (A_KM) { k = idx / BM; m = idx % BM; } else { m = idx / BK; k = idx % BK; }
            float v = 0.f;
            if (m0 + m < M && k0 + k < k_end)
                v = to_f(A_KM ? A[(k0 + k) * lda + m0 + m] : A[(long long)(m0 + m) * lda + k0 + k]);
            As[k][m] = v;
        }
        for (int idx = tid; idx < BK * BN; idx += 256) {
            const int k = idx / BN, nn = idx % BN;
            float v = 0.f;
            if (n0 + nn < N && k0 + k < k_end) v = to_f(B[(k0 + k) * ldb + n0 + nn]);
            Bs[k][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i++) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; j++) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) {
            const int m = m0 + ty + 16 * i, nn = n0 + tx + 16 * j;
            if (m >= M || nn >= N) continue;
            const float v = acc[i][j];
            if (EPI == EPI_STORE) C[(long long)m * ldc + nn] = v;
            else if (EPI == EPI_ATOMIC) atomicAdd(C + (long long)m * ldc + nn, v);
            else if (EPI == EPI_FC1_DGRAD) {
                // nn = c*64 + sq (reference column order) -> [position][sq][c]; ReLU mask of conv2's output
                const long long o = (long long)m * 4096 + (nn & 63) * 64 + (nn >> 6);
                C[o] = __half2float(aux.mask[o]) > 0.f ? v : 0.f;
            } else if (EPI == EPI_FC1_WGRAD) {
                // nn = sq*64 + c -> fc1.weight[m][c*64 + sq]
                atomicAdd(C + (long long)m * 4096 + (nn & 63) * 64 + (nn >> 6), v);
            } else if (EPI == EPI_CONV2_WGRAD) {
                // nn = tap*32 + ci -> conv2.weight[m][ci][tap]
                atomicAdd(C + (long long)m * 288 + (nn & 31) * 9 + (nn >> 5), v);
            } else if (EPI == EPI_CONV1_WGRAD) {
                // nn = tap*16 + p (p < 12) -> conv1.weight[m][p][tap]
                if ((nn & 15) < 12) atomicAdd(C + (long long)m * 108 + (nn & 15) * 9 + (nn >> 4), v);
            }
        }
}

template <typename TA, typename TB, bool A_KM, int BM, int BN, int EPI>
static int run_gemm(const TA* A, const TB* B, float* C, int M, int N, long long K, long long lda, long long ldb,
                    long long ldc, int k_splits, GemmAux aux, cudaStream_t st) {
    long long k_chunk = (K + k_splits - 1) / k_splits;
    k_chunk = (k_chunk + 15) / 16 * 16;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, (unsigned)((K + k_chunk - 1) / k_chunk));
    ProfileScope ps(KK_TRAIN, st);
    gemm_kernel<TA, TB, A_KM, BM, BN, EPI><<<grid, 256, 0, st>>>(A, B, C, M, N, K, lda, ldb, ldc, k_chunk, aux);
    count_launch();
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// loss + backward through tanh, fc2 and fc1's ReLU; one warp per position (grid-stride).  Writes
// dz1h (fc_bwd_kernel.cu): fp16, scaled by `scale`, tiled [128-position tile][32 chunks][16 row
// groups][8 positions][8 halfs]; rows beyond n are written as zeros (the weight gradient contracts
// over whole tiles).  Lane l owns fc1 outputs 8l .. 8l+7.
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ q, const float* __restrict__ qn, const float* __restrict__ reward,
                const float* __restrict__ done, const float* __restrict__ hidden, const float* __restrict__ fc2_w,
                long long n, float gamma, int loss_kind, float scale, __half* __restrict__ dz1h,
                float* __restrict__ g_fc2_w, float* __restrict__ g_fc2_b, float* __restrict__ g_fc1_b,
                float* __restrict__ loss) {
    __shared__ float s_gw[8][256];
    __shared__ float s_gb1[8][256];
    __shared__ float s_misc[8][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gw[8], gb1[8], gb = 0.f, ls = 0.f, w2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { gw[i] = 0.f; gb1[i] = 0.f; w2[i] = fc2_w[lane * 8 + i]; }
    const float inv_n = 1.f / (float)n;
    const long long n_pad = ((n + 127) >> 7) << 7;
    for (long long b = (long long)blockIdx.x * 8 + warp; b < n_pad; b += (long long)gridDim.x * 8) {
        uint4 packed = make_uint4(0, 0, 0, 0);
        if (b < n) {
            const float qv = q[b];
            const float y = reward[b] + (1.f - done[b]) * gamma * qn[b];      // neural_network.py:238
            const float d = qv - y;
            float dl;                                                          // d loss / d q
            if (loss_kind == 0) { ls += d * d; dl = 2.f * d; }                 // F.mse_loss (neural_network.py:241)
            else { const float ad = fabsf(d); ls += ad < 1.f ? 0.5f * d * d : ad - 0.5f; dl = fminf(1.f, fmaxf(-1.f, d)); }
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

// column sums of a [rows][cols] fp32 matrix (bias gradients), cols <= 256
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int cols, float* __restrict__ out) {
    const int c = threadIdx.x % cols, sub = threadIdx.x / cols, nsub = 256 / cols;
    if (sub >= nsub) return;
    float s = 0.f;
    for (long long r = (long long)blockIdx.x * nsub + sub; r < rows; r += (long long)gridDim.x * nsub) s += x[r * cols + c];
    atomicAdd(out + c, s);
}

// im2col of ReLU(conv1): col1[(b,sq)][tap*32 + ci] = a1[b][sq + tap][ci]  (fp16, zero outside)
__global__ void im2col_act1_kernel(const __half* __restrict__ a1, long long n, __half* __restrict__ col) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, sq, tap, quarter)
    if (i >= n * 64 * 9 * 4) return;
    const int qd = (int)(i & 3), tap = (int)((i >> 2) % 9), sq = (int)((i / 36) & 63);
    const long long b = i / (36 * 64);
    const int y = (sq >> 3) + tap / 3 - 1, x = (sq & 7) + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < 8 && x >= 0 && x < 8) v = *reinterpret_cast<const uint4*>(a1 + (b * 64 + y * 8 + x) * 32 + qd * 8);
    *reinterpret_cast<uint4*>(col + (b * 64 + sq) * 288 + tap * 32 + qd * 8) = v;
}

// conv1 weight + bias gradient.  The input is one-hot, so the GEMM collapses to a scatter:
// dW1[co][p][tap] += dz1c[b][sq][co] for the piece p standing on square sq+tap.  One warp per board
// (lane = output channel), per-warp accumulators [9 taps][12 pieces][32] in shared memory (the
// (tap, piece) index is warp-uniform, the lane is the bank: no conflicts, no atomics), one global
// atomicAdd per accumulator cell per CTA at the end.
constexpr int C1W_WARPS = 2;   // 2 x 13.6 KB of accumulators: under the 48 KB default
__global__ void __launch_bounds__(C1W_WARPS * 32)
conv1_wgrad_kernel(const CdqBoard* __restrict__ boards, const float* __restrict__ dz1c, long long n,
                   float* __restrict__ g_w, float* __restrict__ g_b) {
    extern __shared__ float acc_all[];                 // [warps][9*12 + 1][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* acc = acc_all + warp * (109 * 32);
    for (int i = 0; i < 109; i++) acc[i * 32 + lane] = 0.f;
    __syncwarp();
    for (long long b = (long long)blockIdx.x * C1W_WARPS + warp; b < n; b += (long long)gridDim.x * C1W_WARPS) {
        const unsigned long long* r = reinterpret_cast<const unsigned long long*>(boards) + b * 9;
        // lanes 0..31 and a second pass: piece code of every square (12 = empty)
        unsigned long long bb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) bb[k] = __ldg(r + k);
        float bias = 0.f;
        for (int sq = 0; sq < 64; sq++) {
            const float gval = dz1c[(b * 64 + sq) * 32 + lane];
            bias += gval;
#pragma unroll
            for (int tap = 0; tap < 9; tap++) {
                const int y = (sq >> 3) + tap / 3 - 1, x = (sq & 7) + tap % 3 - 1;
                if (y < 0 || y > 7 || x < 0 || x > 7) continue;          // warp-uniform
                const int s = y * 8 + x;
                int t = -1;
#pragma unroll
                for (int k = 0; k < 6; k++) if ((bb[k] >> s) & 1) t = k;
                if (t < 0) continue;                                      // empty square (warp-uniform)
                const int p = ((bb[6] >> s) & 1) ? t : t + 6;
                acc[(tap * 12 + p) * 32 + lane] += gval;
            }
        }
        acc[108 * 32 + lane] += bias;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 109 * 32; i += C1W_WARPS * 32) {
        float v = 0.f;
        for (int w = 0; w < C1W_WARPS; w++) v += acc_all[w * (109 * 32) + i];
        const int cell = i >> 5, co = i & 31;
        if (cell == 108) atomicAdd(g_b + co, v);
        else atomicAdd(g_w + co * 108 + (cell % 12) * 9 + cell / 12, v);   // conv1.weight[co][p][tap]
    }
}

// col2im + ReLU mask: dz1c[b][sq][ci] = a1 > 0 ? sum_tap dcol[b][sq - tap][ci*9 + tap] : 0
__global__ void fold_conv2_dgrad_kernel(const float* __restrict__ dcol, const __half* __restrict__ a1, long long n,
                                        float* __restrict__ dz) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (b, sq, ci)
    if (i >= n * 64 * 32) return;
    const int ci = (int)(i & 31), sq = (int)((i >> 5) & 63);
    const long long b = i >> 11;
    float s = 0.f;
    if (__half2float(a1[i]) > 0.f) {
#pragma unroll
        for (int tap = 0; tap < 9; tap++) {
            const int y = (sq >> 3) - (tap / 3 - 1), x = (sq & 7) - (tap % 3 - 1);   // output square that read us
            if (y >= 0 && y < 8 && x >= 0 && x < 8) s += dcol[(b * 64 + y * 8 + x) * 288 + ci * 9 + tap];
        }
    }
    dz[i] = s;
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
    ProfileScope ps(KK_TRAIN, st);
    adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
    count_launch();
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// workspace carve-up (bytes per position; everything 256-byte aligned per section)
struct TrainWs {
    __half* act2_tiles; __half* act1; __half* dz1h; float* hidden; float* q; float* qn; float* dz2c;
    __half* col1; float* dcol1; float* dz1c;
    size_t total;
};
static TrainWs carve(void* base, long long n) {
    TrainWs w{};
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? (uint8_t*)base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    const long long tiles = (n + 127) / 128;
    w.act2_tiles = (__half*)take((size_t)tiles << 20);
    w.act1 = (__half*)take((size_t)n * 64 * 32 * 2);
    w.dz1h = (__half*)take((size_t)tiles << 16);
    w.hidden = (float*)take((size_t)n * 256 * 4);
    w.q = (float*)take((size_t)n * 4);
    w.qn = (float*)take((size_t)n * 4);
    w.dz2c = (float*)take((size_t)n * 4096 * 4);
    w.col1 = (__half*)take((size_t)n * 64 * 288 * 2);
    w.dcol1 = (float*)take((size_t)n * 64 * 288 * 4);
    w.dz1c = (float*)take((size_t)n * 64 * 32 * 4);
    w.total = off;
    return w;
}
size_t train_workspace_bytes(long long n) { return n > 0 ? carve(nullptr, n).total : 0; }

int train_fwd_bwd(const CdqNet* net, const CdqNet* target, const CdqWeights& w, const CdqBoard* states,
                  const CdqBoard* next_states, const float* reward, const float* done, long long n, float gamma,
                  int loss_kind, const CdqWeights& g /* gradient outputs, pre-zeroed */, float* loss, void* ws_base,
                  cudaStream_t st) {
    TrainWs ws = carve(ws_base, n);
    int e;
    // ---- forward: target value of next_states (inference), then q of states keeping activations
    if ((e = launch_conv_stack(next_states, n, target->p, ws.act2_tiles, nullptr, st))) return e;
    if ((e = launch_fc_head(ws.act2_tiles, n, target->p, ws.qn, nullptr, nullptr, nullptr, st))) return e;
    if ((e = launch_conv_stack(states, n, net->p, ws.act2_tiles, ws.act1, st))) return e;
    if ((e = launch_fc_head(ws.act2_tiles, n, net->p, ws.q, nullptr, nullptr, ws.hidden, st))) return e;
    // ---- loss, tanh, fc2, fc1 bias; then fc1 weight and data gradients on tcgen05 (fc_bwd_kernel.cu)
    const float scale = 64.f * (float)n;      // power-of-two-ish lift into fp16's normal range; undone in the epilogues
    {
        ProfileScope ps(KK_TRAIN, st);
        const long long n_pad = ((n + 127) >> 7) << 7;
        const unsigned blocks = (unsigned)(n_pad / 8 < 1184 ? n_pad / 8 : 1184);
        head_bwd_kernel<<<blocks, 256, 0, st>>>(ws.q, ws.qn, reward, done, ws.hidden, w.fc2_w, n, gamma, loss_kind, scale, ws.dz1h,
                                                (float*)g.fc2_w, (float*)g.fc2_b, (float*)g.fc1_b, loss);
        count_launch();
    }
    if ((e = launch_fc1_wgrad(ws.dz1h, ws.act2_tiles, n, 1.f / scale, (float*)g.fc1_w, st))) return e;
    if ((e = launch_fc1_dgrad(ws.dz1h, ws.act2_tiles, net->p.fc1tp, n, 1.f / scale, ws.dz2c, st))) return e;
    GemmAux aux{nullptr};
    // ---- conv2: bias, weight gradient over im2col(ReLU(conv1)), data gradient via col2im
    const long long rows = n * 64;
    const int ks = (int)(rows / 2048 < 1 ? 1 : (rows / 2048 > 512 ? 512 : rows / 2048));
    {
        ProfileScope ps(KK_TRAIN, st);
        colsum_kernel<<<296, 256, 0, st>>>(ws.dz2c, rows, 64, (float*)g.conv2_b);
        im2col_act1_kernel<<<(unsigned)((rows * 36 + 255) / 256), 256, 0, st>>>(ws.act1, n, ws.col1);
        count_launch(2);
    }
    if ((e = run_gemm<float, __half, true, 64, 128, EPI_CONV2_WGRAD>(ws.dz2c, ws.col1, (float*)g.conv2_w, 64, 288, rows, 64, 288, 288, ks, aux, st)))
        return e;
    if ((e = run_gemm<float, float, false, 128, 128, EPI_STORE>(ws.dz2c, w.conv2_w, ws.dcol1, (int)rows, 288, 64, 64, 288, 288, 1, aux, st)))
        return e;
    {
        ProfileScope ps(KK_TRAIN, st);
        fold_conv2_dgrad_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(ws.dcol1, ws.act1, n, ws.dz1c);
        // ---- conv1: weight + bias gradient by scatter over the one-hot input
        const unsigned blocks = (unsigned)((n + C1W_WARPS - 1) / C1W_WARPS < 1184 ? (n + C1W_WARPS - 1) / C1W_WARPS : 1184);
        conv1_wgrad_kernel<<<blocks, C1W_WARPS * 32, C1W_WARPS * 109 * 32 * 4, st>>>(states, ws.dz1c, n, (float*)g.conv1_w, (float*)g.conv1_b);
        count_launch(2);
    }
    return (int)cudaGetLastError();
}

} // namespace cdq
