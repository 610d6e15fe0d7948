// umma_cb2_probe.cu -- cycles per tcgen05.mma for the operand shapes of conv2's backward (conv_bwd2_kernel.cu), with
// board rows of 10 files (160 bytes: every 128-byte core matrix of the interior straddles two 128-byte lines) against
// rows of 8 files (128 bytes: aligned).  Decides the dz2 board layout.
//   case 0  weight gradient  M = 64, N = 96, MN-major A and B; A rows of 160 B at +16, B taps at +0 / +16 / +32
//   case 1  weight gradient  same, A rows of 128 B (aligned), B unchanged (rows of 160 B, taps +0 / +16 / +32)
//   case 2  weight gradient  A aligned, B aligned for all three taps (what a layout without shifts would give: the floor)
//   case 3  data gradient    M = 128, N = 96, K-major; A rows of 160 B at +16 (SBO 160)
//   case 4  data gradient    A rows of 128 B (SBO 128)
//   case 5  data gradient    N = 32, A rows of 160 B at +16 (the kernel before the N = 96 formulation)
//   case 6  weight gradient  case 0 with the A operand kept in the collector across the three taps (fill / use / lastuse)
// Values are not checked here (the kernel's parity tests do that); only the pace.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

template <int mode>
__global__ void probe(const uint8_t* img, int img_bytes, int reps, long long* cycles) {
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
            const uint32_t B0 = base + 64 * 1024;
            const uint32_t id_wg = make_idesc_f16_ex(64, 96, true, true);
            const uint32_t id_dg = make_idesc_f16(128, mode == 5 ? 32 : 96);
            const uint64_t bw = make_smem_desc(base + 128 * 1024, 1536, 128);
            const long long t0 = clock64();
            for (int r = 0; r < reps; r++) {
                if (mode <= 2 || mode == 6) {
                    // 16 K steps x 3 taps, like one unit
#pragma unroll
                    for (int k = 0; k < 16; k++)
#pragma unroll
                        for (int dxi = 0; dxi < 3; dxi++) {
                            const uint64_t a = mode == 0 ? make_smem_desc(base + 4 * 160 + 16 + k * 320, 160, 6400)
                                                         : make_smem_desc(base + 4 * 128 + k * 256, 128, 5120);
                            const uint64_t b = mode == 2 ? make_smem_desc(B0 + (k >> 1) * 2048 + (k & 1) * 256 + dxi * 8192, 128, 512)
                                                         : make_smem_desc(B0 + (k >> 1) * 2560 + (k & 1) * 320 + dxi * 16, 160, 640);
                            if (mode != 6) umma_f16(tmem + dxi * 96, a, b, id_wg, (r | k) != 0);
                            else if (dxi == 0) umma_f16_ca<0>(tmem + dxi * 96, a, b, id_wg, (r | k) != 0);
                            else if (dxi == 1) umma_f16_ca<1>(tmem + dxi * 96, a, b, id_wg, (r | k) != 0);
                            else umma_f16_ca<2>(tmem + dxi * 96, a, b, id_wg, (r | k) != 0);
                        }
                } else {
#pragma unroll
                    for (int t = 0; t < 2; t++)
#pragma unroll
                        for (int dyi = 0; dyi < 3; dyi++)
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const uint64_t a = mode == 4 ? make_smem_desc(base + (4 * t + dyi) * 512 + j * 2 * 5120, 5120, 128)
                                                             : make_smem_desc(base + (4 * t + dyi) * 640 + 16 + j * 2 * 6400, 6400, 160);
                                umma_f16(tmem + 320 + 96 * t, a, bw + (uint64_t)((dyi * 12288 + j * 3072) >> 4), id_dg, (dyi | j) != 0);
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
    uint8_t* d_img; long long* d_cyc;
    cudaMalloc(&d_img, IMG); cudaMalloc(&d_cyc, 16);
    cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice);
    const char* names[7] = {"wgrad M64 N96 MN-major, A rows 160 B (+16), B taps +0/+16/+32", "wgrad, A rows 128 B (aligned), B taps +0/+16/+32",
                            "wgrad, A and B aligned", "dgrad M128 N96 K-major, A rows 160 B (+16)", "dgrad, A rows 128 B (aligned)",
                            "dgrad N32, A rows 160 B (+16)", "wgrad as case 0, A kept in the collector across the taps"};
    const int per[7] = {48, 48, 48, 24, 24, 24, 48};
    const int reps = 200;
#define RUN(mode)                                                                                            \
    {                                                                                                        \
        cudaFuncSetAttribute(probe<mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);                 \
        for (int rep = 0; rep < 2; rep++) {                                                                  \
            probe<mode><<<1, 128, IMG>>>(d_img, IMG, reps, d_cyc);                                           \
            cudaError_t e = cudaDeviceSynchronize();                                                         \
            if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; } \
        }                                                                                                    \
        long long c;                                                                                         \
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);                                                    \
        printf("case %d %-62s: %6.1f cycles per MMA\n", mode, names[mode], (double)c / reps / per[mode]);    \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
