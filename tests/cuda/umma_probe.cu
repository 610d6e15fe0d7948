// umma_probe.cu -- hardware probe for the shared-memory descriptor assumptions the conv kernel
// relies on (K-major, no swizzle): (A) canonical dense layout, (B) SBO that is not a multiple of
// 128 B, (C) operand start addresses that are only 16-byte aligned (the "3x3 tap = start address"
// trick).  Small-integer fp16 data, so the comparison with the host formula is exact.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/umma_probe.bin tests/cuda/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

struct Case { const char* name; int a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, N, ksteps; };

__global__ void probe_kernel(const uint8_t* img, int img_bytes, Case c, float* out) {
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
    if (threadIdx.x == 0) {
        const uint32_t base = smem_u32(smem);
        for (int k = 0; k < c.ksteps; k++)
            umma_f16(tmem, make_smem_desc(base + c.a_off + k * 2 * c.a_lbo, c.a_lbo, c.a_sbo),
                     make_smem_desc(base + c.b_off + k * 2 * c.b_lbo, c.b_lbo, c.b_sbo),
                     make_idesc_f16(128, c.N), k > 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5;
    for (int col = 0; col < c.N; col += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) out[threadIdx.x * c.N + col + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }

int main() {
    const int IMG = 200 * 1024;
    std::vector<uint8_t> img(IMG);
    srand(1);
    for (int i = 0; i < IMG / 2; i++) {
        __half h = __float2half((float)(rand() % 7 - 3));
        memcpy(&img[2 * i], &h, 2);
    }
    uint8_t* d_img; float* d_out;
    cudaMalloc(&d_img, IMG); cudaMalloc(&d_out, 128 * 256 * 4);
    cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    const int B0 = 160 * 1024; // B operand region
    Case cases[] = {
        {"A canonical  SBO=128 LBO=2048           ", 0, 2048, 128, B0, 512, 128, 32, 1},
        {"B padded     SBO=1600 LBO=25600          ", 0, 25600, 1600, B0, 512, 128, 32, 1},
        {"C shifted    start+16  SBO=1600          ", 16, 25600, 1600, B0, 512, 128, 32, 1},
        {"C shifted    start+176 SBO=1600          ", 176, 25600, 1600, B0, 512, 128, 32, 1},
        {"C shifted    start+336 SBO=1600 K=32 N=64", 336, 25600, 1600, B0, 1024, 128, 64, 2},
        {"D fc shape   N=256 K=64 (4 steps)        ", 0, 2048, 128, B0, 4096, 128, 256, 4},
    };
    int bad_total = 0;
    for (auto& c : cases) {
        cudaMemset(d_out, 0, 128 * 256 * 4);
        probe_kernel<<<1, 128, IMG>>>(d_img, IMG, c, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s : CUDA error %s\n", c.name, cudaGetErrorString(e)); return 2; }
        std::vector<float> out(128 * c.N);
        cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < c.N; n++) {
                float acc = 0;
                for (int k = 0; k < 16 * c.ksteps; k++) {
                    uint16_t a, b;
                    memcpy(&a, &img[c.a_off + (m / 8) * c.a_sbo + (m % 8) * 16 + (k / 8) * c.a_lbo + (k % 8) * 2], 2);
                    memcpy(&b, &img[c.b_off + (n / 8) * c.b_sbo + (n % 8) * 16 + (k / 8) * c.b_lbo + (k % 8) * 2], 2);
                    acc += h2f(a) * h2f(b);
                }
                if (acc != out[m * c.N + n]) { if (bad < 3) printf("   m=%d n=%d want %g got %g\n", m, n, acc, out[m * c.N + n]); bad++; }
            }
        printf("%s : %s (%d / %d mismatches)\n", c.name, bad ? "FAIL" : "ok", bad, 128 * c.N);
        bad_total += bad;
    }
    return bad_total ? 1 : 0;
}
