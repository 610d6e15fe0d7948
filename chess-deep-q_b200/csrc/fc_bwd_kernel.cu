// fc_bwd_kernel.cu -- backward of fc1 (4096 -> 256) on tcgen05, reusing the forward's tiles.
//
// dz1h: the gradient at fc1's pre-activation, fp16, scaled by a power of two so it stays in fp16's
//       normal range, stored per 128 positions as [32 chunks of 8 outputs][16 row groups][8 positions]
//       [8 halfs] (64 KB).  One 128-byte core matrix of it is simultaneously
//         - a K-major core matrix (rows = positions, K = outputs)   -> A operand of the DATA gradient
//         - an MN-major core matrix (M = outputs, K = positions)    -> A operand of the WEIGHT gradient
// act2: the forward's fc1 A tiles; read MN-major (N = channels of a square, K = positions) they are
//       the B operand of the weight gradient (tests/cuda/umma_mn_probe.cu checks the encoding).
//
//  fc1_dgrad_kernel   (work item = one 256-column chunk of one position tile; a CTA takes a CONTIGUOUS range of
//                     items, so the dz1h tile is fetched once per tile it touches: with one (tile, four chunks)
//                     item per CTA the kernel's time was the latency chain of a single item -- 512 KB of W1^T
//                     through a two-stage ring -- whatever the batch size)
//                     dZ2[pos][sq][ch] = ReLU'(act2) * sum_j dz1[pos][j] W1[j][ch*64+sq], written (still
//                     scaled, fp16) as the zero-bordered board blocks conv_bwd2_kernel.cu consumes
//                     M = 128 positions, N = 256 (the 8 squares of one board rank x 32 channels), K = 256; W1^T tiles
//                     streamed by bulk copies, the masked chunk is staged as the board blocks it becomes and copied out with
//                     coalesced 16-byte stores
//  relu_mask_kernel   one bit per ReLU(conv2) output for the above (side stream, under the fc forward)
//  fc1_wgrad_kernel   dW1[j][ch*64+sq] += sum_pos dz1[pos][j] act2[pos][sq][ch]
//                     M = 2 x 128 outputs, N = 128 (16 squares x 8 channels), K = positions (split over CTAs)
#include <cstdlib>
#include "umma.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {

// ------------------------------------------------------------------------------------------ ReLU bit mask of conv2
// One bit per ReLU(conv2) output (value != 0), in the order the data gradient's epilogue consumes them: word
// [tile][chunk = 2 rank + channel half][block p = file][position] holds the 32 channels of that half of one square,
// bit 8 gg + e = channel chunk gg, channel e.  The data gradient used to read the fp16 activations themselves for the
// mask: 64 KB per chunk and CTA through 16 KB of loads in flight, which at the L2 latency is 7 of the 8 us a chunk
// took.  This kernel reads the 8 KB per position once, coalesced, on the side stream under the fc forward.
__global__ void __launch_bounds__(256)
relu_mask_kernel(const __half* __restrict__ act2, long long n_tiles, unsigned* __restrict__ mask) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (tile, square, half, position), position fastest
    if (t >= n_tiles * 64 * 2 * 128) return;
    const int pos = (int)(t & 127), h = (int)((t >> 7) & 1), sq = (int)((t >> 8) & 63);
    const long long tile = t >> 14;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(act2) + ((tile * 64 + sq) << 14) + h * 4 * 2048 + (pos >> 3) * 128 + (pos & 7) * 16;
    uint4 m[4];
#pragma unroll
    for (int gg = 0; gg < 4; gg++) m[gg] = *reinterpret_cast<const uint4*>(src + gg * 2048);
    unsigned bits = 0;
#pragma unroll
    for (int gg = 0; gg < 4; gg++) {
        const unsigned mw[4] = {m[gg].x, m[gg].y, m[gg].z, m[gg].w};
#pragma unroll
        for (int e = 0; e < 4; e++) {
            bits |= ((mw[e] & 0xffffu) ? 1u : 0u) << (gg * 8 + 2 * e);
            bits |= ((mw[e] >> 16) ? 1u : 0u) << (gg * 8 + 2 * e + 1);
        }
    }
    mask[(((tile * 16 + (sq >> 3) * 2 + h) * 8) + (sq & 7)) * 128 + pos] = bits;
}

int launch_relu_mask(const __half* act2, long long n, unsigned* mask, cudaStream_t st) {
    const long long tiles = (n + 127) >> 7, threads = tiles * 64 * 2 * 128;
    ProfileScope ps(KK_TRAIN, st);
    relu_mask_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(act2, tiles, mask);
    count_launch();
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ dgrad
constexpr int DG_A_BYTES = 128 * 256 * 2;       // 64 KB: dz1h tile
constexpr int DG_B_BYTES = 256 * 64 * 2;        // 32 KB: W1^T, 256 columns x 64 outputs
// What sets the pace is the epilogue (build variants under ncu, profiles/r2c_dgrad_notes.txt: 23.4 us; without the epilogue
// 12.8; with the epilogue but without its bulk copies 15.2; without the W1^T stream 23.0): 128 bulk copies of 640 bytes per
// chunk, at the copy engine's rate for such small copies (a second staging buffer, so that they ran under the conversion of the
// next half, changed nothing).  The chunk is staged in two halves of its channel chunks, each in its own buffer, and the four
// epilogue warps copy a half out themselves with plain 16-byte stores, 512 contiguous bytes per instruction.
constexpr int DG_STAGES = 2;
constexpr int DG_OFF_B = DG_A_BYTES;
constexpr int DG_OFF_OUT = DG_OFF_B + DG_STAGES * DG_B_BYTES;   // two buffers, each half a chunk of the output staged as the 64 board blocks it becomes
constexpr int DG_GROUP_STRIDE = 2 * 640 + 16;                   // [group of 4 positions][2 channel chunks][4 pos][10 files][16 B] (+16: bank spread)
constexpr int DG_HALF_BYTES = 32 * DG_GROUP_STRIDE;
constexpr int DG_OUT_BYTES = 2 * DG_HALF_BYTES;
constexpr int DG_OFF_BAR = DG_OFF_OUT + DG_OUT_BYTES;
constexpr int DG_SMEM = DG_OFF_BAR + 128;
constexpr int DG_THREADS = 192;

__global__ void __launch_bounds__(DG_THREADS, 1)
fc1_dgrad_kernel(const __half* __restrict__ dz1h, const unsigned* __restrict__ relu_mask, const __half* __restrict__ fc1tp,
                 long long n, __half* __restrict__ dz2_boards) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DG_OFF_BAR);   // [DG_STAGES]
    uint64_t* empty = full + DG_STAGES;                                // [DG_STAGES]
    uint64_t* a_full = empty + DG_STAGES;
    uint64_t* a_empty = a_full + 1;
    uint64_t* acc_full = a_empty + 1;                                  // [2]
    uint64_t* acc_empty = acc_full + 2;                                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < DG_STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < 2; s++) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // the staged board blocks include the border files, which stay zero for the whole kernel
    for (int i = tid; i < DG_OUT_BYTES / 16; i += DG_THREADS) reinterpret_cast<uint4*>(smem + DG_OFF_OUT)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    __syncthreads();
    const long long n_chunks = ((n + 127) >> 7) * 16;   // (position tile, one of its 16 chunks of 256 columns = one rank x one half of the channels)
    const long long g_lo = n_chunks * blockIdx.x / gridDim.x, g_hi = n_chunks * (blockIdx.x + 1) / gridDim.x;
    constexpr uint32_t IDESC = make_idesc_f16(128, 256);

    if (warp == 0) {
        if (elect_one_sync()) {
            uint32_t it = 0, a_no = 0;
            for (long long g = g_lo; g < g_hi; g++) {
                const long long tile = g >> 4;
                const int ck = (int)(g & 15);
                if (g == g_lo || ck == 0) {              // first chunk of a tile in this CTA's range: its dz1h tile
                    mbar_wait(a_empty, (a_no & 1) ^ 1);
                    mbar_arrive_expect_tx(a_full, DG_A_BYTES);
                    bulk_g2s(smem, reinterpret_cast<const uint8_t*>(dz1h) + tile * DG_A_BYTES, DG_A_BYTES, a_full);
                    a_no++;
                }
                for (int s = 0; s < 4; s++, it++) {
                    const int st = it % DG_STAGES;
                    mbar_wait(empty + st, ((it / DG_STAGES) & 1) ^ 1);
#ifdef DG_NO_B
                    mbar_arrive(full + st);
#else
                    mbar_arrive_expect_tx(full + st, DG_B_BYTES);
                    bulk_g2s(smem + DG_OFF_B + st * DG_B_BYTES,
                             reinterpret_cast<const uint8_t*>(fc1tp) + ((long long)(ck * 4 + s) << 15), DG_B_BYTES, full + st);
#endif
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            uint32_t it = 0, a_no = 0, chunk_no = 0;
            const uint32_t a0 = smem_u32(smem);
            for (long long g = g_lo; g < g_hi; g++, chunk_no++) {
                const int ck = (int)(g & 15);
                if (g == g_lo || ck == 0) {
                    mbar_wait(a_full, a_no & 1);
                    a_no++;
                }
                const int ab = chunk_no & 1;
                mbar_wait(acc_empty + ab, ((chunk_no >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int s = 0; s < 4; s++, it++) {
                    const int st = it % DG_STAGES;
                    mbar_wait(full + st, (it / DG_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t b0 = a0 + DG_OFF_B + st * DG_B_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        umma_f16(tmem + ab * 256, make_smem_desc(a0 + (s * 8 + 2 * j) * 2048, 2048, 128),
                                 make_smem_desc(b0 + j * 2 * 4096, 4096, 128), IDESC, (s | j) != 0);
                    umma_commit(empty + st);
                }
                umma_commit(acc_full + ab);
                if (g + 1 == g_hi || ck == 15) umma_commit(a_empty);    // last chunk of this tile in the range
            }
        }
    } else {
        const int q = warp & 3;
        const int pr = q * 32 + lane;
        uint32_t chunk_no = 0;
        // conv2's ReLU mask: one 32-bit word per 32-column block (relu_mask_kernel), the eight words of a chunk loaded one
        // chunk ahead; applied as a bitwise AND on the packed fp16 pairs.
        unsigned mcur[8], mnext[8];
        auto mload = [&](unsigned (&dst)[8], long long g) {
            const unsigned* src = relu_mask + g * 8 * 128 + pr;
#pragma unroll
            for (int p = 0; p < 8; p++) dst[p] = src[p * 128];
        };
        if (g_lo < g_hi) mload(mnext, g_lo);
        for (long long gch = g_lo; gch < g_hi; gch++) {
            const long long tile = gch >> 4;
            const int ck = (int)(gch & 15);
            const int ab = chunk_no & 1;
#pragma unroll
            for (int p = 0; p < 8; p++) mcur[p] = mnext[p];
            if (gch + 1 < g_hi) mload(mnext, gch + 1);
            mbar_wait(acc_full + ab, (chunk_no >> 1) & 1);
            tc_fence_after();
            // Two passes over the accumulator, one per HALF of the 32 channels of a block (channel chunks 2 hh, 2 hh + 1).
            // Block p is FILE p of the chunk's rank.  The pieces are staged in shared memory exactly as the board blocks
            // they become -- [group of 4 positions][channel chunk][4 positions][10 files][16 B], 640 contiguous bytes per
            // (group, channel chunk) both here and in global memory -- and leave as 64 bulk copies of 640 bytes per pass.
            // (Written straight from the accumulator rows every warp store put 16 bytes on each of 32 different 128-byte
            // lines: 22 of 39 us; as 64-byte runs through per-thread stores still 17 of 28.)
            const bool live = (tile << 7) + pr < n;         // rows beyond n must be ZERO boards: conv2's backward contracts whole units
#ifdef DG_NO_EPI
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + ab);
            chunk_no++;
            continue;
#endif
#pragma unroll 1
            for (int hh = 0; hh < 2; hh++) {
                // buffer hh is free: every thread finished copying it out before the barrier of the round in between
                // four blocks per TMEM round trip (tcgen05.wait::ld waits for every outstanding load, and a round trip per
                // block -- sixteen per chunk -- was most of the epilogue's time): blocks 4..7 are in flight while 0..3 are
                // converted and staged
                uint32_t vv[2][4][16];
                const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + ab * 256 + hh * 16;
#pragma unroll
                for (int p = 0; p < 4; p++) tmem_ld16(tbase + p * 32, vv[0][p]);
#pragma unroll
                for (int ph4 = 0; ph4 < 2; ph4++) {
                    tmem_ld_wait();
                    if (ph4 == 0) {
#pragma unroll
                        for (int p = 0; p < 4; p++) tmem_ld16(tbase + (4 + p) * 32, vv[1][p]);
                    }
#pragma unroll
                    for (int p4 = 0; p4 < 4; p4++) {
                        const int p = ph4 * 4 + p4;
                        uint32_t (&v)[16] = vv[ph4][p4];
                        uint8_t* sdst = smem + DG_OFF_OUT + hh * DG_HALF_BYTES + (pr >> 2) * DG_GROUP_STRIDE + (pr & 3) * 160 + (p + 1) * 16;
                        const unsigned mbits = live ? mcur[p] >> (hh * 16) : 0u;
#pragma unroll
                        for (int g = 0; g < 2; g++) {
                            unsigned w[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                // the value stays scaled (undone downstream)
                                const __half2 h2 = __floats2half2_rn(__uint_as_float(v[g * 8 + 2 * e]), __uint_as_float(v[g * 8 + 2 * e + 1]));
                                const unsigned two = mbits >> (g * 8 + 2 * e);
                                w[e] = *reinterpret_cast<const unsigned*>(&h2) & ((two & 1u ? 0xffffu : 0u) | (two & 2u ? 0xffff0000u : 0u));
                            }
                            *reinterpret_cast<uint4*>(sdst + g * 640) = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
                if (hh == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + ab);
                }
                // ---- copy-out with plain 16-byte stores, 512 contiguous bytes per warp instruction: a warp takes 16 of the
                // half's 64 blocks (group of 4 positions, channel chunk), 640 bytes = 40 pieces each.  (As 64 bulk copies of
                // 640 bytes per half the copy engine set the kernel's pace: 23.4 us, 16.8 without the copies.)
                asm volatile("bar.sync 1, 128;" ::: "memory");             // the four epilogue warps: half chunk staged
                {
                    const int rank = ck >> 1, chalf = ck & 1;
                    uint8_t* gbase = reinterpret_cast<uint8_t*>(dz2_boards) + (chalf * 4 + hh * 2) * 6400 + (rank + 1) * 640;
                    const uint8_t* sbase = smem + DG_OFF_OUT + hh * DG_HALF_BYTES;
#pragma unroll 4
                    for (int bi = 0; bi < 16; bi++) {
                        const int blk = q * 16 + bi, grp = blk >> 1, gl = blk & 1;
                        if ((tile << 7) + grp * 4 < n) {
                            const uint8_t* src = sbase + grp * DG_GROUP_STRIDE + gl * 640;
                            uint8_t* dst = gbase + (tile * 32 + grp) * 51200 + gl * 6400;
                            reinterpret_cast<uint4*>(dst)[lane] = reinterpret_cast<const uint4*>(src)[lane];
                            if (lane < 8) reinterpret_cast<uint4*>(dst)[32 + lane] = reinterpret_cast<const uint4*>(src)[32 + lane];
                        }
                    }
                }
            }
            chunk_no++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------ wgrad
constexpr int WG_A_BYTES = 128 * 256 * 2;       // 64 KB: dz1h tile
constexpr int WG_B_BYTES = 16 * 2048;           // 32 KB: one 8-channel chunk of 16 consecutive squares
constexpr int WG_STAGE = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_OFF_BAR = 2 * WG_STAGE;
constexpr int WG_SMEM = WG_OFF_BAR + 64;
constexpr int WG_THREADS = 192;

__global__ void __launch_bounds__(WG_THREADS, 1)
fc1_wgrad_kernel(const __half* __restrict__ dz1h, const __half* __restrict__ act2, long long n_tiles, int k_splits,
                 float inv_scale, float* __restrict__ g_fc1_w) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_OFF_BAR);   // [2]
    uint64_t* empty = full + 2;                                        // [2]
    uint64_t* acc_full = empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // N block = squares 16*sg .. 16*sg+15 of channels 8*chunk .. 8*chunk+7: sixteen 2 KB pieces of the act2 tile,
    // 16 KB apart in global memory and back to back in shared memory (the MN-major B operand as before).  For a
    // fixed output j and channel the 16 squares are 64 contiguous bytes of fc1.weight's gradient, so the epilogue
    // adds 16-byte vectors: a quarter of the L2 transactions of two-squares-by-64-channels blocks, whose
    // one-float atomics (32 768 per CTA, each its own sector) took longer than the MMAs.
    const int blk = blockIdx.x / k_splits, split = blockIdx.x % k_splits;
    const int sg = blk & 3, chunk = blk >> 2;
    const long long my_tiles = split < n_tiles ? (n_tiles - split + k_splits - 1) / k_splits : 0;
    constexpr uint32_t IDESC = make_idesc_f16_ex(128, 128, true, true);

    if (warp == 0) {
        if (elect_one_sync()) {
            for (long long i = 0; i < my_tiles; i++) {
                const long long t = split + i * k_splits;
                const int st = (int)(i & 1);
                mbar_wait(empty + st, ((i >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(full + st, WG_STAGE);
                bulk_g2s(smem + st * WG_STAGE, reinterpret_cast<const uint8_t*>(dz1h) + t * WG_A_BYTES, WG_A_BYTES, full + st);
#pragma unroll 4
                for (int sqi = 0; sqi < 16; sqi++)
                    bulk_g2s(smem + st * WG_STAGE + WG_A_BYTES + sqi * 2048,
                             reinterpret_cast<const uint8_t*>(act2) + ((t * 64 + sg * 16 + sqi) << 14) + chunk * 2048, 2048, full + st);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            for (long long i = 0; i < my_tiles; i++) {
                const int st = (int)(i & 1);
                mbar_wait(full + st, (i >> 1) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + st * WG_STAGE), b0 = a0 + WG_A_BYTES;
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int k = 0; k < 8; k++)      // K step = 16 positions = 2 row groups = 256 B
                        umma_f16(tmem + h * 128, make_smem_desc(a0 + h * 32768 + k * 256, 128, 2048),
                                 make_smem_desc(b0 + k * 256, 128, 2048), IDESC, (i | k) != 0);
                umma_commit(empty + st);
            }
            if (my_tiles > 0) umma_commit(acc_full);
        }
    } else if (my_tiles > 0) {
        const int q = warp & 3;
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            const int j = h * 128 + q * 32 + lane;
#pragma unroll 1
            for (int p = 0; p < 4; p++) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 128 + p * 32, v);
                tmem_ld_wait();
                // columns of this block: (square 4p + s, channel e) at 8*s + e; reference column = ch*64 + sq
                float* dst = g_fc1_w + (long long)j * 4096 + chunk * 8 * 64 + sg * 16 + 4 * p;
#pragma unroll
                for (int e = 0; e < 8; e++)
                    atomicAdd(reinterpret_cast<float4*>(dst + e * 64),
                              make_float4(__uint_as_float(v[e]) * inv_scale, __uint_as_float(v[8 + e]) * inv_scale,
                                          __uint_as_float(v[16 + e]) * inv_scale, __uint_as_float(v[24 + e]) * inv_scale));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

int launch_fc1_dgrad(const __half* dz1h, const unsigned* relu_mask, const __half* fc1tp, long long n, __half* dz2_boards,
                     cudaStream_t st) {
    once_per_device(ONCE_FC1_DGRAD, [] {
        cudaFuncSetAttribute(fc1_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM);
        cudaFuncSetAttribute(fc1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    });
    const long long chunks = ((n + 127) >> 7) * 16;     // work items: (tile, 256-column chunk), contiguous ranges per CTA
    const int sms = device_sm_count();
    ProfileScope ps(KK_FC1_DGRAD, st);
    const long long per_cta = (chunks + sms - 1) / sms;      // no CTA takes more than this: use as few CTAs as that allows
    fc1_dgrad_kernel<<<(unsigned)((chunks + per_cta - 1) / per_cta), DG_THREADS, DG_SMEM, st>>>(dz1h, relu_mask, fc1tp, n, dz2_boards);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_fc1_wgrad(const __half* dz1h, const __half* act2, long long n, float inv_scale, float* g_fc1_w, cudaStream_t st) {
    once_per_device(ONCE_FC1_DGRAD, [] {
        cudaFuncSetAttribute(fc1_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM);
        cudaFuncSetAttribute(fc1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    });
    const long long tiles = (n + 127) >> 7;
    // K (positions) split: more splits = more CTAs but one epilogue (32 768 elements of atomics) and one prologue each
    static const int max_splits = [] { const char* e = getenv("CDQ_WGRAD_SPLITS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 8 ? v : 2; }();
    const int k_splits = (int)(tiles < max_splits ? tiles : max_splits);
    ProfileScope ps(KK_FC1_WGRAD, st);
    fc1_wgrad_kernel<<<32 * k_splits, WG_THREADS, WG_SMEM, st>>>(dz1h, act2, tiles, k_splits, inv_scale, g_fc1_w);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
