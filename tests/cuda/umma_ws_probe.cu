// umma_ws_probe.cu -- does tcgen05.mma.ws (weight-stationary, B kept in collector b0) give the
// same D layout/values as the plain form for M=128, N=64, K=16 with fill/use/use/lastuse over four
// different A tiles?  Exact small-integer fp16 data.  Also times both forms (clock64).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

constexpr int A_LBO = 12800, A_SBO = 160, B_LBO = 1024, B_SBO = 128, N = 64;

__global__ void probe(const uint8_t* img, int img_bytes, int b_off, int use_ws, int reps, float* out, long long* cycles) {
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
            const uint32_t idesc = make_idesc_f16(128, N);
            const long long t0 = clock64();
            for (int r = 0; r < reps; r++)
                for (int k = 0; k < 2; k++) {           // two K=16 steps, each with its own B
                    const uint64_t b = make_smem_desc(base + b_off + k * 2 * B_LBO, B_LBO, B_SBO);
                    uint64_t a[4];
                    for (int t = 0; t < 4; t++) a[t] = make_smem_desc(base + t * 2560 + 176 + k * 2 * A_LBO, A_LBO, A_SBO);
                    const bool acc = (k > 0) && (r == 0 || true);
                    if (use_ws) {
                        umma_f16_ws<0>(tmem + 0 * N, a[0], b, idesc, acc && !(r > 0 && k == 0));
                        umma_f16_ws<1>(tmem + 1 * N, a[1], b, idesc, acc);
                        umma_f16_ws<1>(tmem + 2 * N, a[2], b, idesc, acc);
                        umma_f16_ws<2>(tmem + 3 * N, a[3], b, idesc, acc);
                    } else {
                        for (int t = 0; t < 4; t++) umma_f16(tmem + t * N, a[t], b, idesc, acc);
                    }
                }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            cycles[use_ws] = clock64() - t0;
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    for (int col = 0; col < 4 * N; col += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) out[threadIdx.x * 4 * N + col + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }

int main() {
    const int IMG = 200 * 1024, B0 = 160 * 1024;
    std::vector<uint8_t> img(IMG);
    srand(2);
    for (int i = 0; i < IMG / 2; i++) { __half h = __float2half((float)(rand() % 7 - 3)); memcpy(&img[2 * i], &h, 2); }
    uint8_t* d_img; float* d_out; long long* d_cyc;
    cudaMalloc(&d_img, IMG); cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&d_cyc, 16);
    cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    int rc = 0;
    for (int ws = 0; ws < 2; ws++) {
        cudaMemset(d_out, 0, 128 * 256 * 4);
        probe<<<1, 128, IMG>>>(d_img, IMG, B0, ws, 1, d_out, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", ws ? "ws" : "plain", cudaGetErrorString(e)); return 2; }
        std::vector<float> out(128 * 256);
        cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int t = 0; t < 4; t++)
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    float acc = 0;
                    for (int k = 0; k < 32; k++) {
                        uint16_t a, b;
                        memcpy(&a, &img[t * 2560 + 176 + (m / 8) * A_SBO + (m % 8) * 16 + (k / 8) * A_LBO + (k % 8) * 2], 2);
                        memcpy(&b, &img[B0 + (n / 8) * B_SBO + (n % 8) * 16 + (k / 8) * B_LBO + (k % 8) * 2], 2);
                        acc += h2f(a) * h2f(b);
                    }
                    const float got = out[m * 256 + t * N + n];
                    if (acc != got) { if (bad < 4) printf("   t=%d m=%d n=%d want %g got %g\n", t, m, n, acc, got); bad++; }
                }
        printf("%s: %s (%d mismatches)\n", ws ? "tcgen05.mma.ws fill/use/use/lastuse" : "tcgen05.mma plain", bad ? "FAIL" : "ok", bad);
        rc |= bad != 0;
    }
    // timing: 200 repetitions x 8 MMAs (M=128, N=64, K=16; math floor 32 cycles each)
    for (int ws = 0; ws < 2; ws++) {
        probe<<<1, 128, IMG>>>(d_img, IMG, B0, ws, 200, d_out, d_cyc);
        cudaDeviceSynchronize();
        long long c[2];
        cudaMemcpy(c, d_cyc, 16, cudaMemcpyDeviceToHost);
        printf("%s: %.1f cycles per MMA (1600 MMAs)\n", ws ? "ws   " : "plain", c[ws] / 1600.0);
    }
    return rc;
}
