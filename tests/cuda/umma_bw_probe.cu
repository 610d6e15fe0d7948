// umma_bw_probe.cu -- operand-fetch cost of tcgen05.mma (kind::f16, M=128, K=16) on one SM:
// cycles per MMA as a function of N, of the A start alignment, and of A-collector reuse
// (collector::a::fill / use / lastuse).  Decides the conv kernel's tiling.
//   mode 0  N=64   A start +16 B (misaligned core matrices, SBO 160)   -- the round-1 conv operand
//   mode 1  N=64   A aligned (SBO 128), a different A per MMA
//   mode 2  N=96   A aligned
//   mode 3  N=128  A aligned
//   mode 4  N=192  A aligned
//   mode 5  N=256  A aligned
//   mode 6  3 x N=64 on the SAME A: collector::a::fill, use, lastuse  (D columns 0/64/128)
//   mode 7  3 x N=64 on the SAME A, plain (no collector hint)
// Modes 4, 6 and 7 must produce the same D (checked).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

constexpr int A_TILE = 2048;       // aligned A tile: 16 row groups x 128 B (one K chunk), LBO below
constexpr int A_LBO = 20480, B_LBO = 4096, B_OFF = 160 * 1024;

__global__ void probe(const uint8_t* img, int img_bytes, int mode, int reps, float* out, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint32_t base = smem_u32(smem);
            const int Ns[8] = {64, 64, 96, 128, 192, 256, 64, 64};
            const int N = Ns[mode];
            const uint32_t idesc = make_idesc_f16(128, N);
            const uint64_t b = make_smem_desc(base + B_OFF, B_LBO, 128);
            const long long t0 = clock64();
            for (int r = 0; r < reps; r++) {
#pragma unroll 1
                for (int t = 0; t < 8; t++) {      // 8 different A tiles per repetition
                    const bool acc = r > 0 || t > 0;
                    if (mode == 0) {
                        umma_f16(tmem, make_smem_desc(base + t * 2560 + 176, 12800, 160), b, idesc, acc);
                    } else if (mode <= 5) {
                        umma_f16(tmem, make_smem_desc(base + t * A_TILE, A_LBO, 128), b, idesc, acc);
                    } else {
                        const uint64_t a = make_smem_desc(base + t * A_TILE, A_LBO, 128);
                        const uint64_t b1 = make_smem_desc(base + B_OFF + 1024, B_LBO, 128);
                        const uint64_t b2 = make_smem_desc(base + B_OFF + 2048, B_LBO, 128);
                        if (mode == 6) {
                            umma_f16_ca<0>(tmem, a, b, idesc, acc);
                            umma_f16_ca<1>(tmem + 64, a, b1, idesc, acc);
                            umma_f16_ca<2>(tmem + 128, a, b2, idesc, acc);
                        } else {
                            umma_f16(tmem, a, b, idesc, acc);
                            umma_f16(tmem + 64, a, b1, idesc, acc);
                            umma_f16(tmem + 128, a, b2, idesc, acc);
                        }
                    }
                }
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            cycles[0] = clock64() - t0;
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    for (int col = 0; col < 256; col += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) out[threadIdx.x * 256 + col + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

// sweep: N, number of accumulators in rotation, A alignment (0 = aligned SBO 128, 1 = +176 SBO 160)
__global__ void sweep(const uint8_t* img, int img_bytes, int N, int nacc, int misaligned, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint32_t base = smem_u32(smem);
            const uint32_t idesc = make_idesc_f16(128, N);
            const uint64_t b = make_smem_desc(base + B_OFF, B_LBO, 128);
            uint64_t a[8];
            for (int t = 0; t < 8; t++)
                a[t] = misaligned ? make_smem_desc(base + t * 2560 + 176, 12800, 160) : make_smem_desc(base + t * A_TILE, A_LBO, 128);
            const long long t0 = clock64();
            int acc_i = 0;
            for (int r = 0; r < reps; r++) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    umma_f16(tmem + acc_i * N, a[t], b, idesc, r > 0);
                    acc_i = acc_i + 1 == nacc ? 0 : acc_i + 1;
                }
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            cycles[0] = clock64() - t0;
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// split: per A tile, one N=64 MMA + one N=128 MMA on the SAME A (D columns 0..63, 64..191), with
// (coll=1) or without (coll=0) the A collector; coll=2: single N=192 MMA for reference
template <int coll>
__global__ void split(const uint8_t* img, int img_bytes, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint32_t base = smem_u32(smem);
            const uint32_t i64 = make_idesc_f16(128, 64), i128 = make_idesc_f16(128, 128), i192 = make_idesc_f16(128, 192);
            const uint64_t b = make_smem_desc(base + B_OFF, B_LBO, 128);
            const uint64_t b1 = make_smem_desc(base + B_OFF + 1024, B_LBO, 128);
            uint64_t a[8];
            for (int t = 0; t < 8; t++) a[t] = make_smem_desc(base + t * A_TILE, A_LBO, 128);
            const long long t0 = clock64();
            for (int r = 0; r < reps; r++) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    if (coll == 1) {
                        umma_f16_ca<0>(tmem, a[t], b, i64, r > 0);
                        umma_f16_ca<2>(tmem + 64, a[t], b1, i128, r > 0);
                    } else if (coll == 0) {
                        umma_f16(tmem, a[t], b, i64, r > 0);
                        umma_f16(tmem + 64, a[t], b1, i128, r > 0);
                    } else if (coll == 2) {
                        umma_f16(tmem, a[t], b, i192, r > 0);
                    } else if (coll == 3) {
                        umma_f16_ca<0>(tmem, a[t], b, i64, r > 0);
                        umma_f16_ca<1>(tmem + 64, a[t], b1, i64, r > 0);
                        umma_f16_ca<2>(tmem + 128, a[t], b1, i64, r > 0);
                    } else if (coll == 4) {
                        const uint32_t i32 = make_idesc_f16(128, 32);
                        umma_f16_ca<0>(tmem, a[t], b, i32, r > 0);
                        umma_f16_ca<1>(tmem + 64, a[t], b1, i32, r > 0);
                        umma_f16_ca<2>(tmem + 128, a[t], b1, i32, r > 0);
                    } else {
                        const uint32_t i96 = make_idesc_f16(128, 96);
                        umma_f16(tmem, a[t], b, i96, r > 0);
                    }
                }
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            cycles[0] = clock64() - t0;
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    const int IMG = 200 * 1024;
    std::vector<uint8_t> img(IMG);
    srand(3);
    for (int i = 0; i < IMG / 2; i++) { __half h = __float2half((float)(rand() % 5 - 2)); memcpy(&img[2 * i], &h, 2); }
    uint8_t* d_img; float* d_out; long long* d_cyc;
    cudaMalloc(&d_img, IMG); cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&d_cyc, 16);
    cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    // correctness: N=192 in one MMA == 3 x N=64 with and without the A collector
    std::vector<float> ref(128 * 256), got(128 * 256);
    int rc = 0;
    for (int mode : {4, 6, 7}) {
        cudaMemset(d_out, 0, 128 * 256 * 4);
        probe<<<1, 128, IMG>>>(d_img, IMG, mode, 1, d_out, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
        cudaMemcpy((mode == 4 ? ref : got).data(), d_out, ref.size() * 4, cudaMemcpyDeviceToHost);
        if (mode != 4) {
            int bad = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 192; n++) bad += ref[m * 256 + n] != got[m * 256 + n];
            printf("mode %d vs N=192 single MMA: %s (%d mismatches)\n", mode, bad ? "FAIL" : "ok", bad);
            rc |= bad != 0;
        }
    }
    const char* names[8] = {"N=64  misaligned A (+176, SBO 160)", "N=64  aligned A", "N=96  aligned A", "N=128 aligned A",
                            "N=192 aligned A", "N=256 aligned A", "3 x N=64 same A, collector a fill/use/lastuse",
                            "3 x N=64 same A, plain"};
    const int floor_c[8] = {32, 32, 48, 64, 96, 128, 96, 96};
    for (int mode = 0; mode < 8; mode++) {
        const int reps = 200;
        probe<<<1, 128, IMG>>>(d_img, IMG, mode, reps, d_out, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
        long long c;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-48s: %7.1f cycles per A tile (math floor %d)\n", names[mode], c / (reps * 8.0), floor_c[mode]);
    }
    cudaFuncSetAttribute(sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    printf("\nsweep: cycles per MMA (M=128, K=16), rows N, columns = accumulators in rotation\n");
    for (int mis = 0; mis < 2; mis++) {
        printf("%s A\n   N    floor", mis ? "misaligned (+176 B, SBO 160)" : "aligned (SBO 128)");
        for (int nacc : {1, 2, 3, 4, 6, 8}) printf("  acc=%d ", nacc);
        printf("\n");
        for (int N : {32, 64, 96, 128, 192, 256}) {
            printf("%4d  %6d ", N, N / 2);
            for (int nacc : {1, 2, 3, 4, 6, 8}) {
                if (nacc * N > 512) { printf("      - "); continue; }
                sweep<<<1, 128, IMG>>>(d_img, IMG, N, nacc, mis, 200, d_cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
                long long c;
                cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
                printf(" %7.1f", c / 1600.0);
            }
            printf("\n");
        }
    }
    cudaFuncSetAttribute(split<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    cudaFuncSetAttribute(split<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    cudaFuncSetAttribute(split<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    cudaFuncSetAttribute(split<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    cudaFuncSetAttribute(split<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    cudaFuncSetAttribute(split<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    for (int coll = 0; coll < 6; coll++) {
        if (coll == 0) split<0><<<1, 128, IMG>>>(d_img, IMG, 200, d_cyc);
        else if (coll == 1) split<1><<<1, 128, IMG>>>(d_img, IMG, 200, d_cyc);
        else if (coll == 2) split<2><<<1, 128, IMG>>>(d_img, IMG, 200, d_cyc);
        else if (coll == 3) split<3><<<1, 128, IMG>>>(d_img, IMG, 200, d_cyc);
        else if (coll == 4) split<4><<<1, 128, IMG>>>(d_img, IMG, 200, d_cyc);
        else split<5><<<1, 128, IMG>>>(d_img, IMG, 200, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        long long c;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-52s: %7.1f cycles per A tile\n", coll == 0 ? "N=64 + N=128 on the same A, plain" : coll == 1 ?
               "N=64 + N=128 on the same A, collector fill/lastuse" : coll == 2 ? "N=192 single MMA" : coll == 3 ?
               "3 x N=64 on the same A, collector fill/use/lastuse" : coll == 4 ? "3 x N=32 on the same A, collector fill/use/lastuse" :
               "N=96 single MMA", c / 1600.0);
    }
    return rc;
}
