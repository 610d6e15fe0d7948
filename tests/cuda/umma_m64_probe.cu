// umma_m64_probe.cu -- where do the 64 rows of an M=64 (cta_group::1) accumulator live in TMEM?
// MN-major A (M = 64 channels, K = cells) and MN-major B (N = 32), as conv2's weight gradient uses.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

__global__ void probe_kernel(const uint8_t* img, int img_bytes, int a_sbo, int a_lbo, int b_off, int b_sbo, int b_lbo, int N, float* out) {
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
    // pre-fill the accumulator region with a sentinel so untouched lanes are visible
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint32_t base = smem_u32(smem);
            umma_f16(tmem, make_smem_desc(base, a_lbo, a_sbo), make_smem_desc(base + b_off, b_lbo, b_sbo),
                     make_idesc_f16_ex(64, N, true, true), false);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    for (int col = 0; col < N; col += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) out[threadIdx.x * N + col + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }

int main() {
    const int IMG = 128 * 1024, B0 = 64 * 1024, N = 32;
    const int a_sbo = 2048, a_lbo = 128, b_sbo = 2048, b_lbo = 160;
    std::vector<uint8_t> img(IMG);
    srand(4);
    for (int i = 0; i < IMG / 2; i++) { __half h = __float2half((float)(rand() % 7 - 3)); memcpy(&img[2 * i], &h, 2); }
    uint8_t* d_img; float* d_out;
    cudaMalloc(&d_img, IMG); cudaMalloc(&d_out, 128 * N * 4);
    cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, 128 * N * 4);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    probe_kernel<<<1, 128, IMG>>>(d_img, IMG, a_sbo, a_lbo, B0, b_sbo, b_lbo, N, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> out(128 * N), want(64 * N);
    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
    for (int m = 0; m < 64; m++)
        for (int n = 0; n < N; n++) {
            float acc = 0;
            for (int k = 0; k < 16; k++) {
                uint16_t a, b;
                memcpy(&a, &img[(m / 8) * a_sbo + (k / 8) * a_lbo + (k % 8) * 16 + (m % 8) * 2], 2);
                memcpy(&b, &img[B0 + (n / 8) * b_sbo + (k / 8) * b_lbo + (k % 8) * 16 + (n % 8) * 2], 2);
                acc += h2f(a) * h2f(b);
            }
            want[m * N + n] = acc;
        }
    // for every TMEM lane find the row it holds
    int ok_rows = 0;
    for (int lane = 0; lane < 128; lane++) {
        int found = -1;
        for (int m = 0; m < 64 && found < 0; m++) {
            bool eq = true;
            for (int n = 0; n < N; n++) if (out[lane * N + n] != want[m * N + n]) { eq = false; break; }
            if (eq) found = m;
        }
        if (found >= 0) ok_rows++;
        if (lane % 16 == 0 || found < 0) printf("lane %3d -> row %d\n", lane, found);
    }
    printf("lanes holding a correct row: %d (expect 64)\n", ok_rows);
    return ok_rows == 64 ? 0 : 1;
}
