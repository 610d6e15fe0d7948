// dp_adam_kernel.cu -- the optimizer step of the DQN replay trainer (torch.optim.Adam defaults,
// neural_network.py:65) fused with its data-parallel collective: ONE kernel per rank does
//
//     reduce-scatter of the gradients  ->  Adam on the rank's own slice  ->  all-gather of the parameters
//
// over NVLink peer memory (CUDA IPC mappings of every rank's buffers), instead of
// ncclAllReduce(4.28 MB) followed by a full-size Adam on every rank:
//   * rank r owns chunks [C r / W, C (r+1) / W) of the flat parameter vector (C = chunks of 4 floats);
//   * it sums that slice of every rank's gradient buffer with 16-byte peer loads in rank order
//     0..W-1 (each element is summed by exactly one rank, so all replicas receive identical bits),
//     scales by 1/W, updates ITS slice of exp_avg / exp_avg_sq (optimizer state is sharded,
//     28 -> 28/W bytes of state traffic per parameter) and stores the new parameters into every
//     rank's parameter buffer;
//   * once every peer has signalled that it is done reading, each rank zeroes its own gradient buffer (locally), so the
//     next step needs no zero_grad pass;
//   * cross-GPU ordering uses two epoch flags per (rank, peer) written with system-scope release
//     stores: "my gradients are complete" before the reduction, "my stores have landed" after it.
// With W = 1 the same kernel is the plain fused Adam step with a device-side step counter, which
// is what lets the whole training step live in one CUDA graph (no host-computed bias correction).
#include <atomic>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {

constexpr int DP_THREADS = 512;
// Waiting for a peer is bounded by WALL CLOCK (%globaltimer), not by an iteration count: ranks legitimately arrive
// seconds apart (a checkpoint write on rank 0, a first-step graph capture, a GC pause).  The limit is the option
// "dp_timeout_s" (CDQ_DP_TIMEOUT_S, default 600 s).  A wait that runs out does NOT trap (a trap poisons the CUDA
// context of every rank): the kernel sets grid_bar[2], skips the update of its slice, still completes its own
// arrival protocol, and the host reports CDQ_E_TIMEOUT from cdq_dp_adam_status.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// NVLS: one load that returns the sum over every rank's copy (added inside the NVSwitch), one store that lands in every copy
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// every CTA waits until all W ranks have published `epoch` in this rank's flag row; true = the wait ran out
__device__ __forceinline__ bool wait_peers(const unsigned* my_flags, int world, unsigned epoch, unsigned long long deadline,
                                           unsigned* err) {
    int late = 0;
    if (threadIdx.x < world) {
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(my_flags + threadIdx.x) - epoch) < 0) {
            if ((++spins & 63u) == 0 && globaltimer_ns() > deadline) { late = 1; atomicExch(err, 1u); break; }
        }
    }
    return __syncthreads_or(late) != 0;
}

__global__ void __launch_bounds__(DP_THREADS)
dp_adam_kernel(const CdqDpPeers peers, int rank, int world, float* __restrict__ m, float* __restrict__ v,
               long long chunk_lo, long long chunk_hi, long long chunk_lo2, long long chunk_hi2, int phase, int bump,
               int* __restrict__ d_step, unsigned* __restrict__ grid_bar_base, float lr, float b1, float b2, float eps,
               unsigned long long timeout_ns) {
    // phase: a step may be applied in two launches (fc1.weight early, the rest at the end of the backward pass); each
    // launch of a step uses its own arrival counters (grid_bar + 4 phase) and its own flag row (flags + CDQ_MAX_RANKS phase);
    // only the launch with bump = 1 advances Adam's step counter, both use t = *d_step + 1
    unsigned* __restrict__ grid_bar = grid_bar_base + 4 * phase;
    const int flag_off = CDQ_MAX_RANKS * phase;
    __shared__ float s_bc[2];
    const unsigned long long deadline = globaltimer_ns() + timeout_ns;
    // both counters are read by every thread before block 0 bumps them at the very end
    const int step = *d_step + 1;                      // Adam's t (bias correction); reset by load_state_dict
    const unsigned gen = grid_bar[1] + 1u;             // launch generation of this group: never reset
    const unsigned epoch = 2u * gen;
    if (threadIdx.x == 0) {
        s_bc[0] = (float)(1.0 - pow((double)b1, (double)step));
        s_bc[1] = (float)sqrt(1.0 - pow((double)b2, (double)step));
    }
    // ---- A: this rank's gradients are complete (stream order); tell every peer, wait for all of them
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + flag_off + rank, epoch - 1);
    const bool late = wait_peers(peers.flags[rank] + flag_off, world, epoch - 1, deadline, grid_bar_base + 2);
#ifdef CDQ_DP_TRACE
    unsigned long long* trace = reinterpret_cast<unsigned long long*>(grid_bar_base + 16);
    if (blockIdx.x == 0 && threadIdx.x == 0) { trace[0] = deadline - timeout_ns; trace[1] = globaltimer_ns(); }
#endif
    const float bc1 = s_bc[0], bc2_sqrt = s_bc[1], inv_world = 1.f / (float)world;
    const bool multicast = peers.mc_params != nullptr && peers.mc_grads != nullptr;

    // ---- reduce-scatter + Adam + all-gather on this rank's slice (skipped when a peer's gradients never arrived)
    // the chunks [chunk_lo, chunk_hi) followed by [chunk_lo2, chunk_hi2) form one virtual range that is sliced over the ranks
    const long long n1 = chunk_hi - chunk_lo, n_chunks = n1 + (chunk_hi2 - chunk_lo2);
    const long long lo = n_chunks * rank / world, hi = late ? 0 : n_chunks * (rank + 1) / world;
    for (long long ci = lo + (long long)blockIdx.x * DP_THREADS + threadIdx.x; ci < hi; ci += (long long)gridDim.x * DP_THREADS) {
        const long long c = ci < n1 ? chunk_lo + ci : chunk_lo2 + (ci - n1);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (multicast) g = multimem_ld_reduce_add(peers.mc_grads + 4 * c);
        else
            for (int p = 0; p < world; p++) {
                const float4 x = reinterpret_cast<const float4*>(peers.grads[p])[c];
                g.x += x.x; g.y += x.y; g.z += x.z; g.w += x.w;
            }
        float4 mi = reinterpret_cast<float4*>(m)[c], vi = reinterpret_cast<float4*>(v)[c];
        float4 pi = reinterpret_cast<const float4*>(peers.params[rank])[c];
#define CDQ_ADAM1(F)                                                                    \
        {                                                                               \
            const float gi = g.F * inv_world;                                           \
            mi.F = b1 * mi.F + (1.f - b1) * gi;                                         \
            vi.F = b2 * vi.F + (1.f - b2) * gi * gi;                                    \
            /* torch: denom = sqrt(v)/sqrt(bc2) + eps ; p -= (lr / bc1) * m / denom */  \
            pi.F -= (lr / bc1) * (mi.F / (sqrtf(vi.F) / bc2_sqrt + eps));               \
        }
        CDQ_ADAM1(x) CDQ_ADAM1(y) CDQ_ADAM1(z) CDQ_ADAM1(w)
#undef CDQ_ADAM1
        reinterpret_cast<float4*>(m)[c] = mi;
        reinterpret_cast<float4*>(v)[c] = vi;
        if (multicast) multimem_st(peers.mc_params + 4 * c, pi);
        else
            for (int p = 0; p < world; p++) reinterpret_cast<float4*>(peers.params[p])[c] = pi;
    }

    // ---- B: all CTAs of this rank have stored; publish, then wait until every peer's stores have landed here
#ifdef CDQ_DP_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) trace[2] = globaltimer_ns();
#endif
    __threadfence_system();
    __syncthreads();
#ifdef CDQ_DP_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) trace[3] = globaltimer_ns();
#endif
    if (threadIdx.x == 0) {
        // arrival counter that wraps to 0 with the last CTA of THIS launch (launches of different sizes share it); the last
        // arriver publishes the generation, block 0 waits for it (co-resident cooperative grid: always arrives)
        if (atomicInc(grid_bar, gridDim.x - 1) == gridDim.x - 1) { __threadfence(); atomicExch(grid_bar + 3, gen); }
        if (blockIdx.x == 0) {
            while (*reinterpret_cast<volatile unsigned*>(grid_bar + 3) != gen) { }
            __threadfence_system();
        }
    }
    __syncthreads();
#ifdef CDQ_DP_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) trace[4] = globaltimer_ns();
#endif
    if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(peers.flags[threadIdx.x] + flag_off + rank, epoch);
    wait_peers(peers.flags[rank] + flag_off, world, epoch, deadline, grid_bar_base + 2);
#ifdef CDQ_DP_TRACE
    if (blockIdx.x == 0 && threadIdx.x == 0) trace[5] = globaltimer_ns();
#endif
    // every peer has published epoch, i.e. finished its loop: nobody reads this rank's gradients any more.  They are zeroed
    // HERE, locally (zeroing the consumed slice on every peer from the owner's loop doubled the NVLink store traffic that
    // the system fence above has to drain: profiles/README.md, round 2)
    if (!late) {
        float4* mine = reinterpret_cast<float4*>(peers.grads[rank]);
        for (long long ci = (long long)blockIdx.x * DP_THREADS + threadIdx.x; ci < n_chunks; ci += (long long)gridDim.x * DP_THREADS)
            mine[ci < n1 ? chunk_lo + ci : chunk_lo2 + (ci - n1)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { if (bump) *d_step = step; grid_bar[1] = gen; }
}

// world = 1: no peers, so no flags and no system-scope fences.  One 16-byte chunk per thread (the grid is
// sized so that the whole parameter vector is in flight at once; the looped, 148-CTA peer version spends
// 28 us on 30 MB), loads issued before the bias-correction constants are needed.  The step counter is still
// bumped by block 0 after every CTA has read it, through the same never-reset arrival counter.
__global__ void __launch_bounds__(DP_THREADS, 4)      // 4 x 512 threads per SM (32 registers): the whole vector in flight in one round
adam_single_kernel(float* __restrict__ params, float* __restrict__ grads, float* __restrict__ m, float* __restrict__ v,
                   long long chunk_lo, long long chunk_hi, long long chunk_lo2, long long chunk_hi2, int phase, int bump,
                   int* __restrict__ d_step, unsigned* __restrict__ grid_bar_base, float lr, float b1, float b2, float eps) {
    // the chunks [chunk_lo, chunk_hi) followed by [chunk_lo2, chunk_hi2) (the parameters before and after fc1.weight when a
    // step is applied in two launches); phase / bump as in dp_adam_kernel
    __shared__ float s_bc[2];
    unsigned* __restrict__ grid_bar = grid_bar_base + 4 * phase;
    const long long n1 = chunk_hi - chunk_lo, n_chunks = n1 + (chunk_hi2 - chunk_lo2);
    const int step = *d_step + 1;
    const unsigned gen = grid_bar[1] + 1u;
    if (threadIdx.x == 0) {
        s_bc[0] = (float)(1.0 - pow((double)b1, (double)step));
        s_bc[1] = (float)sqrt(1.0 - pow((double)b2, (double)step));
    }
    for (long long c0 = (long long)blockIdx.x * DP_THREADS; c0 < n_chunks; c0 += (long long)gridDim.x * DP_THREADS) {
        const long long ci = c0 + threadIdx.x;
        const bool live = ci < n_chunks;
        const long long c = ci < n1 ? chunk_lo + ci : chunk_lo2 + (ci - n1);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), mi = g, vi = g, pi = g;
        if (live) {
            g = reinterpret_cast<const float4*>(grads)[c];
            mi = reinterpret_cast<const float4*>(m)[c];
            vi = reinterpret_cast<const float4*>(v)[c];
            pi = reinterpret_cast<const float4*>(params)[c];
        }
        __syncthreads();                                 // s_bc (first round); harmless afterwards
        const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
#define CDQ_ADAM1(F)                                                                    \
        {                                                                               \
            const float gi = g.F;                                                       \
            mi.F = b1 * mi.F + (1.f - b1) * gi;                                         \
            vi.F = b2 * vi.F + (1.f - b2) * gi * gi;                                    \
            pi.F -= (lr / bc1) * (mi.F / (sqrtf(vi.F) / bc2_sqrt + eps));               \
        }
        CDQ_ADAM1(x) CDQ_ADAM1(y) CDQ_ADAM1(z) CDQ_ADAM1(w)
#undef CDQ_ADAM1
        if (live) {
            reinterpret_cast<float4*>(m)[c] = mi;
            reinterpret_cast<float4*>(v)[c] = vi;
            reinterpret_cast<float4*>(params)[c] = pi;
            reinterpret_cast<float4*>(grads)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();                                     // every thread of this CTA has read d_step / grid_bar[1]
    if (threadIdx.x == 0) {
        // last CTA out bumps the counters: every other CTA has arrived, i.e. has read them (no CTA waits for another one)
        if (atomicInc(grid_bar, gridDim.x - 1) == gridDim.x - 1) {
            if (bump) *d_step = step;
            grid_bar[1] = gen;
            grid_bar[3] = gen;
        }
    }
}

// part: 0 = the whole vector in one launch; 1 = fc1.weight only (phase 0, no step bump); 2 = everything else (phase 1, bumps).
// The flat vector is laid out in the reference's parameter order, so fc1.weight is chunks [FC1_LO, FC1_HI).
constexpr long long FC1_LO = (32 * 12 * 9 + 32 + 64 * 32 * 9 + 64) / 4, FC1_HI = FC1_LO + 256 * 4096 / 4;
int launch_dp_adam(const CdqDpPeers& peers, int rank, int world, float* m, float* v, long long n_padded, int* d_step,
                   unsigned* grid_bar, float lr, float b1, float b2, float eps, cudaStream_t st, int part) {
    const long long all_chunks = n_padded / 4;
    const int phase = part == 2 ? 1 : 0, bump = part == 1 ? 0 : 1;
    const int sms = device_sm_count();
    if (world == 1) {
        long long lo = 0, hi = all_chunks, lo2 = 0, hi2 = 0;
        if (part == 1) { lo = FC1_LO; hi = FC1_HI; }
        else if (part == 2) { hi = FC1_LO; lo2 = FC1_HI; hi2 = all_chunks; }
        const long long n_chunks = (hi - lo) + (hi2 - lo2);
        int grid = (int)((n_chunks + DP_THREADS - 1) / DP_THREADS);
        // no CTA waits for another one (the last to arrive bumps the counters), so this is a plain launch: one 16-byte chunk per
        // thread, CTAs start wherever an SM has room -- also next to the CTAs of conv2's backward when fc1.weight's launch
        // runs under it (cdq_train_step, train_split_optimizer)
        if (grid < 1) grid = 1;
        float* params = peers.params[0];
        float* grads = peers.grads[0];
        ProfileScope ps(KK_ADAM, st);
        adam_single_kernel<<<grid, DP_THREADS, 0, st>>>(params, grads, m, v, lo, hi, lo2, hi2, phase, bump, d_step, grid_bar, lr, b1, b2, eps);
        count_launch();
        return (int)cudaGetLastError();
    }
    unsigned long long timeout_ns = (unsigned long long)(option(OPT_DP_TIMEOUT_S) > 0 ? option(OPT_DP_TIMEOUT_S) : 1) * 1000000000ull;
    long long lo = 0, hi = all_chunks, lo2 = 0, hi2 = 0;
    if (part == 1) { lo = FC1_LO; hi = FC1_HI; }
    else if (part == 2) { hi = FC1_LO; lo2 = FC1_HI; hi2 = all_chunks; }
    const long long slice = ((hi - lo) + (hi2 - lo2) + world - 1) / world;
    int grid = (int)((slice + DP_THREADS - 1) / DP_THREADS);
    if (grid > sms) grid = sms;          // all CTAs must be co-resident: block 0 waits for the others
    if (grid < 1) grid = 1;
    ProfileScope ps(KK_ADAM, st);
    // cooperative: the CTAs wait for each other.  (A plain launch -- co-residency would hold by construction, one CTA of 512
    // threads per SM -- gains nothing here: 0.1384 against 0.1372 ms at 2 GPUs; the kernel's start is gated by the peers.)
    void* args[] = {(void*)&peers, &rank, &world, &m, &v, &lo, &hi, &lo2, &hi2, (void*)&phase, (void*)&bump, &d_step, &grid_bar, &lr, &b1, &b2,
                    &eps, &timeout_ns};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)dp_adam_kernel, dim3(grid), dim3(DP_THREADS), args, 0, st);
    count_launch();
    return (int)e;
}

} // namespace cdq

using namespace cdq;

extern "C" {

int cdq_ipc_alloc(size_t bytes, void** d_ptr, void* handle64) {
    if (!d_ptr || !handle64 || bytes == 0) return CDQ_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CDQ_CUDA(cudaMalloc(d_ptr, bytes));
    CDQ_CUDA(cudaMemset(*d_ptr, 0, bytes));
    CDQ_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, *d_ptr));
    return 0;
}
int cdq_ipc_open(const void* handle64, void** d_ptr) {
    if (!d_ptr || !handle64) return CDQ_E_BADARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CDQ_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int cdq_ipc_close(void* d_ptr) { return d_ptr ? (int)cudaIpcCloseMemHandle(d_ptr) : 0; }
int cdq_ipc_free(void* d_ptr) { return d_ptr ? (int)cudaFree(d_ptr) : 0; }

int cdq_dp_adam_step(const CdqDpPeers* peers, int rank, int world, float* d_m, float* d_v, int64_t n_params_padded,
                     int32_t* d_step, uint32_t* d_grid_bar, float lr, float beta1, float beta2, float eps, void* stream) {
    if (!peers || world < 1 || world > CDQ_MAX_RANKS || rank < 0 || rank >= world || !d_m || !d_v || !d_step || !d_grid_bar ||
        n_params_padded <= 0 || (n_params_padded & 3))
        return CDQ_E_BADARG;
    for (int p = 0; p < world; p++)
        if (!peers->params[p] || !peers->grads[p] || !peers->flags[p]) return CDQ_E_BADARG;
    int dev = current_device(), major = 0;
    if (dev < 0) return CDQ_E_NODEVICE;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) return CDQ_E_NODEVICE;
    return launch_dp_adam(*peers, rank, world, d_m, d_v, n_params_padded, d_step, d_grid_bar, lr, beta1, beta2, eps,
                          (cudaStream_t)stream, 0);
}


int cdq_dp_adam_status(const uint32_t* d_grid_bar, void* stream) {
    if (!d_grid_bar) return CDQ_E_BADARG;
    uint32_t err = 0;
    CDQ_CUDA(cudaMemcpyAsync(&err, d_grid_bar + 2, sizeof(err), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CDQ_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return err ? CDQ_E_TIMEOUT : 0;
}

} // extern "C"
