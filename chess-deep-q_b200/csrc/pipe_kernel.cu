// pipe_kernel.cu -- conv stack and fc head as ONE persistent kernel with an L2-resident hand-over.
//
// Run back to back, conv_stream_kernel writes ReLU(conv2) (8 KB per position) to HBM and
// fc_head_kernel reads it back: 16 KB per position of HBM traffic, which makes the fc kernel HBM
// bound (5.8 of 6.5 TB/s) and, because the whole step runs into the board's power cap, costs clock
// for both.  Here the CTAs of one grid split into two roles:
//   CTAs [0, n_conv)       conv_stream_body<PIPE>: produce 16-position sub-tiles into a ring of
//                          1 MiB act2 tiles (128 positions each) that stays in the 126 MB L2
//   CTAs [n_conv, grid)    fc_head_body<PIPE>: consume whole tiles from the ring
// with two counters per hand-over unit (a whole tile by default, CDQ_PIPE_GF files of one otherwise) in global memory:
// `ready` (+1 per epilogue-2 warp and sub-tile after a __threadfence, 64 per use; the fc producer
// lane polls it with ld.acquire.gpu and issues fence.proxy.async.global before its bulk loads) and
// `consumed` (+1 by the fc MMA issuer once every stage of the unit has landed in its shared
// memory; the conv epilogue polls it before it overwrites the unit).  fc1 accumulates its K
// dimension in file-major square order, the order in which the conv side produces a tile.  One launch covers the whole batch, so there are no chunk boundaries.
// Both roles are resident together by construction (grid = SM count, one CTA per SM), and a
// hand-over that never happens traps instead of hanging (CDQ_PIPE_SPIN_LIMIT).
#include <atomic>
#include <cstdlib>
#include <mutex>
#include "conv_stream.cuh"
#include "fc_head.cuh"

namespace cdq {

constexpr int PIPE_SMEM = cs::SMEM > FC_SMEM ? cs::SMEM : FC_SMEM;

// The pipe kernel marks its ring as persisting in L2.  Other entry points (training, the two-kernel
// forward) hand those lines back before they run, otherwise half of the L2 stays reserved.
static std::atomic<bool> g_l2_persisting{false};
static std::atomic<bool> g_l2_limit_set{false};
void release_persisting_l2() {
    if (g_l2_persisting.exchange(false, std::memory_order_relaxed)) {
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);   // give the set-aside back as well
        g_l2_limit_set.store(false, std::memory_order_relaxed);
        cudaGetLastError();
    }
}

__global__ void __launch_bounds__(cs::THREADS_PIPE, 1)
leaf_pipe_kernel(const CdqBoard* __restrict__ boards, long long n, const PackedNet net, int n_conv, const cs::PipeCtl pipe,
                 float* __restrict__ value, const unsigned* __restrict__ flags, float* __restrict__ leaf) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((int)blockIdx.x < n_conv)
        cs::conv_stream_body<true>(smem, boards, n, net, nullptr, nullptr, nullptr, (int)blockIdx.x, n_conv, pipe);
    else
        fc_head_body<true>(smem, nullptr, n, net, value, flags, leaf, nullptr, (int)blockIdx.x - n_conv,
                           (int)gridDim.x - n_conv, pipe);
}

// ring tiles, then the two counter arrays
constexpr unsigned PIPE_RING_TILES_DEFAULT = 64;   // MiB of ring: about half of the 126 MB L2 is the measured optimum
static unsigned pipe_ring_tiles() {
    static const unsigned r = [] {
        const char* e = getenv("CDQ_PIPE_RING");
        const long v = e ? atol(e) : 0;
        return (unsigned)(v >= 16 && v <= 1024 ? v : PIPE_RING_TILES_DEFAULT);
    }();
    return r;
}
size_t pipe_workspace_bytes() { return ((size_t)pipe_ring_tiles() << 20) + 2 * (size_t)pipe_ring_tiles() * cs::PIPE_UNITS * 4 + 256; }

// fraction of the SMs that runs the conv role: the two roles must produce and consume at the same rate
static int pipe_conv_ctas(int sms) {
    static const int pct = [] {
        const char* e = getenv("CDQ_PIPE_CONV_PCT");
        const long v = e ? atol(e) : 0;
        return (int)(v >= 10 && v <= 90 ? v : 66);
    }();
    int c = sms * pct / 100;
    if (c < 1) c = 1;
    if (c > sms - 1) c = sms - 1;
    return c;
}

int launch_leaf_pipe(const CdqBoard* boards, long long n, const PackedNet& net, float* value, const unsigned* flags,
                     float* leaf, void* ws, cudaStream_t st) {
    if (n == 0) return 0;
    once_per_device(ONCE_PIPE, [] { cudaFuncSetAttribute(leaf_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM); });
    const int sms = device_sm_count();
    cs::PipeCtl pipe;
    pipe.ring = (uint8_t*)ws;
    pipe.ring_tiles = pipe_ring_tiles();
    pipe.ready = (unsigned*)((uint8_t*)ws + ((size_t)pipe.ring_tiles << 20));
    pipe.consumed = pipe.ready + pipe.ring_tiles * cs::PIPE_UNITS;
    CDQ_CUDA(cudaMemsetAsync(pipe.ready, 0, 2 * (size_t)pipe.ring_tiles * cs::PIPE_UNITS * 4, st));
    int n_conv = pipe_conv_ctas(sms);
    // Both roles must be resident together.  A cooperative launch starts the grid only when all of its
    // CTAs fit on the device at once; on top of that, pipe kernels of different streams (the host path
    // double-buffers on two) are chained with an event so that two of them never interleave their CTAs.
    static std::mutex mu;
    static cudaEvent_t last = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!last) CDQ_CUDA(cudaEventCreateWithFlags(&last, cudaEventDisableTiming));
    else CDQ_CUDA(cudaStreamWaitEvent(st, last, 0));
    // keep the ring's lines in L2: they are overwritten long before anything needs them in HBM, so
    // every write-back of a ring line is wasted traffic (CDQ_PIPE_PERSIST=0 disables the window)
    static const bool persist = [] { const char* e = getenv("CDQ_PIPE_PERSIST"); return !(e && e[0] == '0'); }();
    if (persist) {
        const size_t ring_bytes = (size_t)pipe.ring_tiles << 20;
        if (!g_l2_limit_set.load(std::memory_order_relaxed)) {
            int dev = 0, max_persist = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ring_bytes < (size_t)max_persist ? ring_bytes : (size_t)max_persist);
            g_l2_limit_set.store(true, std::memory_order_relaxed);
        }
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.base_ptr = pipe.ring;
        attr.accessPolicyWindow.num_bytes = ring_bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaGetLastError();     // a refused window only costs the optimisation
        g_l2_persisting.store(true, std::memory_order_relaxed);
    }
    cudaError_t e;
    {
        ProfileScope ps(KK_PIPE, st);
        void* args[] = {(void*)&boards, (void*)&n, (void*)&net, (void*)&n_conv, (void*)&pipe, (void*)&value, (void*)&flags,
                        (void*)&leaf};
        e = cudaLaunchCooperativeKernel((const void*)leaf_pipe_kernel, dim3(sms), dim3(cs::THREADS_PIPE), args, PIPE_SMEM, st);
        count_launch();
    }
    if (e != cudaSuccess) return (int)e;
    CDQ_CUDA(cudaEventRecord(last, st));
    return 0;
}

} // namespace cdq
