// umma_pace_probe.cu -- what does an MMA-issuing warp lose per "step" of six N=192 MMAs (576 cycles
// of tensor work) when the step also contains the bookkeeping of a real pipeline?
//   mode 0  MMAs only, one elected thread, no waits
//   mode 1  + whole-warp elect/syncwarp around every step (the structure of the conv kernel)
//   mode 2  + one mbarrier try_wait per step on an ALREADY COMPLETED phase (whole warp polls)
//   mode 3  + one tcgen05.commit per step to a barrier nobody waits on
//   mode 4  = 2 + 3
//   mode 5  = 4, but the try_wait sits in the MIDDLE of the step (after 4 of the 6 MMAs)
//   mode 6  = 4, but only the elected lane polls the barrier
// Prints cycles per step.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

constexpr int A_TILE = 2048, A_LBO = 20480, B_LBO = 4096, B_OFF = 160 * 1024;

template <int MODE>
__global__ void pace(const uint8_t* img, int img_bytes, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_done, bar_ready, bar_sink;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < img_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(img)[i];
    if (threadIdx.x == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_ready, 1); mbar_init(&bar_sink, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(&bar_ready);      // phase 0 of bar_ready is complete from now on
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        const uint32_t base = smem_u32(smem);
        const uint32_t idesc = make_idesc_f16(128, 192);
        const uint64_t b = make_smem_desc(base + B_OFF, B_LBO, 128);
        uint64_t a[6];
        for (int t = 0; t < 6; t++) a[t] = make_smem_desc(base + t * A_TILE, A_LBO, 128);
        const long long t0 = clock64();
        if (MODE == 0) {
            if (elect_one_sync()) {
                for (int r = 0; r < reps; r++) {
#pragma unroll
                    for (int t = 0; t < 6; t++) umma_f16(tmem, a[t], b, idesc, true);
                }
            }
            __syncwarp();
        } else {
            for (int r = 0; r < reps; r++) {
                if (MODE == 2 || MODE == 4) mbar_wait(&bar_ready, 0);
                if (MODE == 6) { if (lane == 0) mbar_wait(&bar_ready, 0); __syncwarp(); }
                if (elect_one_sync()) {
#pragma unroll
                    for (int t = 0; t < 4; t++) umma_f16(tmem, a[t], b, idesc, true);
                }
                __syncwarp();
                if (MODE == 5) mbar_wait(&bar_ready, 0);
                if (elect_one_sync()) {
#pragma unroll
                    for (int t = 4; t < 6; t++) umma_f16(tmem, a[t], b, idesc, true);
                    if (MODE >= 3) umma_commit(&bar_sink);
                }
                __syncwarp();
            }
        }
        if (elect_one_sync()) umma_commit(&bar_done);
        __syncwarp();
        mbar_wait(&bar_done, 0);
        if (lane == 0) cycles[0] = clock64() - t0;
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
    const char* names[7] = {"MMAs only", "+ elect/syncwarp per step", "+ satisfied try_wait per step", "+ commit per step",
                            "+ try_wait + commit", "+ try_wait (mid-step) + commit", "+ try_wait (one lane polls) + commit"};
    const int reps = 400;
#define RUN(M)                                                                                         \
    {                                                                                                  \
        cudaFuncSetAttribute(pace<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);               \
        pace<M><<<1, 128, IMG>>>(d_img, IMG, reps, d_cyc);                                             \
        cudaError_t e = cudaDeviceSynchronize();                                                       \
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", M, cudaGetErrorString(e)); return 2; } \
        long long c;                                                                                   \
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);                                              \
        printf("mode %d %-40s: %7.1f cycles per step (6 x N=192 MMAs = 576 of math)\n", M, names[M], (double)c / reps); \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    return 0;
}
