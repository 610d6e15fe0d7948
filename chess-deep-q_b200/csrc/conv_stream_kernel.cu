// conv_stream_kernel.cu -- the stand-alone launch of the file-streamed conv stack (conv_stream.cuh):
// act2 goes to the caller's buffer in fc1's tile order; used for small batches, the training
// forward (keeps ReLU(conv1)) and the debug hook.  Large leaf batches go through pipe_kernel.cu.
#include "conv_stream.cuh"

namespace cdq {
namespace cs {

__global__ void __launch_bounds__(THREADS, 1)
conv_stream_kernel(const CdqBoard* __restrict__ boards, long long n, const PackedNet net,
                   __half* __restrict__ act2, __half* __restrict__ dbg_act1, __half* __restrict__ act1_boards) {
    extern __shared__ __align__(1024) uint8_t smem[];
    conv_stream_body<false>(smem, boards, n, net, act2, dbg_act1, act1_boards, (int)blockIdx.x, (int)gridDim.x, PipeCtl{});
}

} // namespace cs

#ifdef CDQ_TRACE
extern "C" int cdq_debug_set_trace(long long* d_buf) {
    return (int)cudaMemcpyToSymbol(cs::g_trace, &d_buf, sizeof(d_buf));
}
#endif

int launch_conv_stack(const CdqBoard* boards, long long n, const PackedNet& net, __half* act2, __half* dbg_act1,
                      cudaStream_t st, __half* act1_boards, int max_ctas) {
    if (n == 0) return 0;
    once_per_device(ONCE_CONV_STREAM, [] { cudaFuncSetAttribute(cs::conv_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cs::SMEM); });
    const long long tiles = ((n + 127) >> 7) * (128 / cs::TILE_POS);
    const int sms = device_sm_count();
    unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (max_ctas > 0 && grid > (unsigned)max_ctas) grid = (unsigned)max_ctas;   // leave SMs to a concurrent launch
    ProfileScope ps(KK_CONV, st);
    cs::conv_stream_kernel<<<grid, cs::THREADS, cs::SMEM, st>>>(boards, n, net, act2, dbg_act1, act1_boards);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
