// cnn.cuh -- packed network handle shared by cnn_kernels.cu and abi.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../include/cdq.h"

namespace cdq {

// device pointers; passed to kernels by value
struct PackedNet {
    const __half* w1p;   // conv1: [9 taps][2 chunks][4 groups][8 n][8 k]  (k 12 = bias, centre tap)
    const __half* w2p;   // conv2: [9 taps][4 chunks][8 groups][8 n][8 k]
    const __half* fc1p;  // fc1:   [64 squares][8 chunks][32 groups][8 n][8 k]
    const __half* w2tp;  // conv2 flipped for the data gradient: [3 dy][8 chunks of co][3 dx][4 groups][8 ci][8 co]
    const __half* wss;   // conv1 then conv2 in conv_stream_kernel's shared-memory order: [dy][K chunk][dx = +1, 0, -1][n groups][8 n][8 k]
    const __half* fc1tp; // fc1^T: [16 column chunks][4 K stages][8 chunks][32 groups][8 n][8 j]  (data gradient)
    const float* conv2_b;
    const float* fc1_b;
    const float* fc2_w;
    const float* fc2_b;
};

int launch_pack_weights(const CdqWeights& w, const PackedNet& p, cudaStream_t st, float* zero_me = nullptr, int part = 0);   // fp32 parameters -> p
int launch_dp_adam(const CdqDpPeers& peers, int rank, int world, float* m, float* v, long long n_padded, int* d_step,
                   unsigned* grid_bar, float lr, float b1, float b2, float eps, cudaStream_t st, int part);
int launch_relu_mask(const __half* act2, long long n, unsigned* mask, cudaStream_t st);   // 1 bit per ReLU(conv2) output, dgrad's order
int launch_fc1_dgrad(const __half* dz1h, const unsigned* relu_mask, const __half* fc1tp, long long n, __half* dz2_boards,
                     cudaStream_t st);
int launch_conv2_bwd(const __half* dz2_boards, const __half* a1_boards, const __half* w2tp, long long n, float inv_scale,
                     float* g_w2, float* g_b2, float* dz1c, unsigned* sched, cudaStream_t st);
int launch_fc1_wgrad(const __half* dz1h, const __half* act2, long long n, float inv_scale, float* g_fc1_w, cudaStream_t st);
int launch_conv_stack(const CdqBoard* boards, long long n, const PackedNet& net, __half* act2, __half* dbg_act1,
                      cudaStream_t st, __half* act1_boards = nullptr, int max_ctas = 0);
// conv stack + fc head as one persistent kernel with an L2-resident hand-over (pipe_kernel.cu)
size_t pipe_workspace_bytes();
void release_persisting_l2();   // undo the pipe kernel's L2 persistence window before other work
int launch_leaf_pipe(const CdqBoard* boards, long long n, const PackedNet& net, float* value, const unsigned* flags,
                     float* leaf, void* ws, cudaStream_t st);
int launch_fc_head(const __half* act2, long long n, const PackedNet& net, float* value, const unsigned* flags,
                   float* leaf, float* hidden, cudaStream_t st);
// small batches: fc1's K dimension split over a cluster of four CTAs per tile (fc_split_kernel.cu); launch_fc_head
// routes batches of at most FC_SPLIT_MAX_TILES tiles of 128 positions there unless CDQ_FC_SPLIT=0
constexpr long long FC_SPLIT_MAX_TILES = 64;
bool fc_split_enabled();
// the backward head fused into the online network's fc kernel of the replay step (fc_split_kernel.cu, TRAIN)
struct TrainHead {
    const float* qn; const float* reward; const float* done;     // target values, rewards, done flags [n] (qn / done unused in upstream mode)
    float gamma, inv_n, scale; int loss_kind;
    __half* dz1h;                                                // fc1 pre-activation gradient, fp16 tiles scaled by `scale`
    float* g_fc2_w; float* g_fc2_b; float* g_fc1_b; float* loss;  // accumulated
};
int launch_fc_head_split(const __half* act2, long long n, const PackedNet& net, float* value, const unsigned* flags,
                         float* leaf, float* hidden, cudaStream_t st, const TrainHead* train_head = nullptr);

size_t train_workspace_bytes(long long n);
int launch_replay_sample(const CdqBoard* mem_states, const CdqBoard* mem_next, const float* mem_reward, const float* mem_done,
                         const int* ring, unsigned long long seed, unsigned* counter, long long batch, CdqBoard* states, CdqBoard* next,
                         float* reward, float* done, cudaStream_t st);
unsigned replay_permute_host(unsigned k, unsigned len, unsigned long long seed, unsigned counter);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float b1, float b2,
                float eps, float grad_scale, cudaStream_t st);

} // namespace cdq

struct CdqNet;
namespace cdq {
int train_fwd_bwd(const CdqNet* net, const CdqNet* target, const CdqWeights& w, const CdqBoard* states,
                  const CdqBoard* next_states, const float* reward, const float* done, long long n, double mean_over,
                  float gamma, int loss_kind, const CdqWeights& g, float* loss, void* ws_base, cudaStream_t st,
                  const CdqOptimizer* opt = nullptr);
}

// what the opaque C handle points to
struct CdqNet {
    cdq::PackedNet p;
    void* blob;      // one device allocation holding everything above
    int loaded;
};
