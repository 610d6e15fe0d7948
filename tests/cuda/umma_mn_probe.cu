// umma_mn_probe.cu -- MN-major operands, no swizzle: a 128-byte core matrix holds 8 K entries
// (16 B apart), each a vector of 8 consecutive M (or N) elements; SBO = stride between core matrices
// along M/N, LBO = stride between 8-wide K groups.  The same bytes are ALSO a valid K-major tile of
// the transposed problem, which is what lets fc1's weight gradient reuse the forward tiles.
//   D[m][n] = sum_k A(m,k) B(n,k),  M = 128, N = 256 or 128, K = 16 * ksteps
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../chess-deep-q_b200/csrc/umma.cuh"
using namespace cdq;

struct Case { const char* name; int a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, N, ksteps, a_kstep, b_kstep; };

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
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint32_t base = smem_u32(smem);
            for (int k = 0; k < c.ksteps; k++)
                umma_f16(tmem, make_smem_desc(base + c.a_off + k * c.a_kstep, c.a_lbo, c.a_sbo),
                         make_smem_desc(base + c.b_off + k * c.b_kstep, c.b_lbo, c.b_sbo),
                         make_idesc_f16_ex(128, c.N, true, true), k > 0);
            umma_commit(&bar);
        }
        __syncwarp();
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
    const int IMG = 200 * 1024, B0 = 96 * 1024;
    std::vector<uint8_t> img(IMG);
    srand(3);
    for (int i = 0; i < IMG / 2; i++) { __half h = __float2half((float)(rand() % 7 - 3)); memcpy(&img[2 * i], &h, 2); }
    uint8_t* d_img; float* d_out;
    cudaMalloc(&d_img, IMG); cudaMalloc(&d_out, 128 * 256 * 4);
    cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG);
    // fc1 wgrad shapes: A = dz1 tile [32 chunks of j][16 row groups][8 pos][8 j]  (M = j: SBO 2048, K = pos: LBO 128)
    //                   B = act2 blocks [sq][8 chunks of ch][16 row groups][8 pos][8 ch] (N = ch: SBO 2048, K = pos: LBO 128)
    Case cases[] = {
        {"MN-major A,B  N=128 K=16        ", 0, 128, 2048, B0, 128, 2048, 128, 1, 256, 256},
        {"MN-major A,B  N=128 K=128 (8x)  ", 0, 128, 2048, B0, 128, 2048, 128, 8, 256, 256},
        {"MN-major A(+32KB half),B N=256  ", 32768, 128, 2048, B0, 128, 2048, 256, 8, 256, 256},
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
                    const int ks = k / 16, kk = k % 16;
                    uint16_t a, b;
                    // element (mn, k): core matrix (mn/8, kk/8); inside: K entry kk%8 at 16 B, mn%8 at 2 B
                    memcpy(&a, &img[c.a_off + ks * c.a_kstep + (m / 8) * c.a_sbo + (kk / 8) * c.a_lbo + (kk % 8) * 16 + (m % 8) * 2], 2);
                    memcpy(&b, &img[c.b_off + ks * c.b_kstep + (n / 8) * c.b_sbo + (kk / 8) * c.b_lbo + (kk % 8) * 16 + (n % 8) * 2], 2);
                    acc += h2f(a) * h2f(b);
                }
                if (acc != out[m * c.N + n]) { if (bad < 3) printf("   m=%d n=%d want %g got %g\n", m, n, acc, out[m * c.N + n]); bad++; }
            }
        printf("%s : %s (%d / %d mismatches)\n", c.name, bad ? "FAIL" : "ok", bad, 128 * c.N);
        bad_total += bad;
    }
    return bad_total ? 1 : 0;
}
