// abi.cu -- the extern "C" boundary of libcdq.so (include/cdq.h).  Thin: argument checks, chunking,
// stream plumbing.  All compute is in the kernels; there is no CPU path.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int k) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }

// ---- profiling: event pairs around kernel launches, summed per kernel kind on collect
static std::atomic<int> g_profile{0};
static std::mutex g_prof_mu;
struct ProfRec { int kind; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
ProfileScope::ProfileScope(int kind_, cudaStream_t st_) : kind(kind_), st(st_) {
    if (!g_profile.load(std::memory_order_relaxed)) return;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
}
ProfileScope::~ProfileScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> l(g_prof_mu);
    g_prof.push_back({kind, a, b});
}

// positions per conv->fc pass (act2 = 1 036 MiB): 148 SMs x 7 fc tiles of 128 = 148 SMs x 56 conv tiles of 16,
// so both persistent kernels end their last round with every SM busy
constexpr long long FWD_CHUNK = 148 * 8 * 128;      // 148 SMs x 4 pairs of fc tiles = 148 x 64 conv tiles
static inline long long tiles_of(long long n) { return (n + 127) / 128; }

// With option "pipe" = 1 (CDQ_PIPE=1), batches of at least this many positions go through the conv -> fc pipe kernel
// (one launch, act2 stays in L2: 90 MB instead of 17 GB of DRAM traffic per 1M positions).  It is opt-in: under
// the board's power cap it is only 1-2 % faster than the two stand-alone kernels, and its persisting-L2
// set-aside slows down the chunked host path that runs next to it (profiles/README.md, "pipe kernel").
constexpr long long PIPE_MIN_POSITIONS = 512 * 1024;   // measured break-even is about 256 K (profiles/README.md)

// ---- run-time switches: read from the environment once, changed afterwards only through cdq_set_option
static std::atomic<int> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;
static const char* const OPT_NAME[OPT_COUNT] = {"pipe", "fc_split", "host_ramp", "dp_timeout_s", "train_split_optimizer"};
static const char* const OPT_ENV[OPT_COUNT] = {"CDQ_PIPE", "CDQ_FC_SPLIT", "CDQ_HOST_RAMP", "CDQ_DP_TIMEOUT_S", "CDQ_TRAIN_SPLIT_OPT"};
static const int OPT_DEFAULT[OPT_COUNT] = {0, 1, 1, 600, 0};
static void init_options() {
    std::call_once(g_opt_once, [] {
        for (int k = 0; k < OPT_COUNT; k++) {
            const char* e = getenv(OPT_ENV[k]);
            g_opt[k].store(e && e[0] ? atoi(e) : OPT_DEFAULT[k], std::memory_order_relaxed);
        }
    });
}
int option(int which) {
    init_options();
    return g_opt[which].load(std::memory_order_relaxed);
}

// ---- per-device state
namespace {
struct DeviceState {
    std::mutex mu;
    unsigned once_mask = 0;
    std::atomic<int> sms{0};
    std::atomic<int> arch{0};       // 0 unknown, 1 sm_100, -1 anything else
};
DeviceState g_dev[MAX_DEVICES];
} // namespace
int current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev >= 0 && dev < MAX_DEVICES ? dev : -1;
}
std::mutex& device_mutex(int dev) { return g_dev[dev].mu; }
unsigned& device_once_mask(int dev) { return g_dev[dev].once_mask; }
int device_sm_count() {
    const int dev = current_device();
    if (dev < 0) return 0;
    int sms = g_dev[dev].sms.load(std::memory_order_relaxed);
    if (!sms) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        g_dev[dev].sms.store(sms, std::memory_order_relaxed);
    }
    return sms;
}
// the CURRENT device of the calling thread must be an sm_100 part; checked once per device
static int check_device() {
    const int dev = current_device();
    if (dev < 0) return CDQ_E_NODEVICE;
    int a = g_dev[dev].arch.load(std::memory_order_relaxed);
    if (a == 0) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        a = major == 10 ? 1 : -1;
        g_dev[dev].arch.store(a, std::memory_order_relaxed);
    }
    return a == 1 ? 0 : CDQ_E_NODEVICE;
}
} // namespace cdq

using namespace cdq;

extern "C" {

int cdq_version(void) { return CDQ_VERSION; }

const char* cdq_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case CDQ_E_BADARG: return "cdq: bad argument";
        case CDQ_E_NODEVICE: return "cdq: no sm_100 CUDA device (this library has no CPU fallback)";
        case CDQ_E_WORKSPACE: return "cdq: workspace too small (see cdq_workspace_bytes)";
        case CDQ_E_BADBOARD: return "cdq: malformed board record";
        case CDQ_E_TIMEOUT: return "cdq: a data-parallel peer did not reach the optimizer step within the time limit";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "cdq: unknown error";
    }
}

uint64_t cdq_launch_count(void) { return g_launches.load(); }

int cdq_set_option(const char* name, int value) {
    if (!name) return CDQ_E_BADARG;
    init_options();
    for (int k = 0; k < OPT_COUNT; k++)
        if (!strcmp(name, OPT_NAME[k])) { g_opt[k].store(value); return 0; }
    return CDQ_E_BADARG;
}
int cdq_get_option(const char* name, int* value) {
    if (!name || !value) return CDQ_E_BADARG;
    init_options();
    for (int k = 0; k < OPT_COUNT; k++)
        if (!strcmp(name, OPT_NAME[k])) { *value = g_opt[k].load(); return 0; }
    return CDQ_E_BADARG;
}

void cdq_profile_enable(int on) { g_profile.store(on ? 1 : 0); }

// Waits for all recorded launches, adds their device time (ms) and count per kernel kind
// (0 eval, 1 conv stack, 2 fc head, 3 encode planes, 4 playouts/expand, 5 weight pack, 6 train)
// into ms[8] / launches[8], then clears the records.
int cdq_profile_collect(double* ms, uint64_t* launches) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    for (auto& r : g_prof) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return (int)e;
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        const int k = r.kind >= KK_COUNT ? KK_TRAIN : r.kind;
        if (ms) ms[k] += t;
        if (launches) launches[k] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return 0;
}

// Same with the training kernels kept apart: ms[16] / launches[16], kinds 8 head (loss, tanh, fc2), 9 fc1 weight
// gradient, 10 fc1 data gradient, 11 conv2 backward, 12 conv1 weight gradient, 13 optimizer.
int cdq_profile_collect_ex(double* ms, uint64_t* launches) {
    std::lock_guard<std::mutex> l(g_prof_mu);
    for (auto& r : g_prof) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) return (int)e;
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        if (ms) ms[r.kind] += t;
        if (launches) launches[r.kind] += 1;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    return 0;
}

int cdq_eval_terms(const CdqBoard* d_boards, int64_t n, int ignore_checkmate, int32_t* d_terms, double* d_score,
                   uint32_t* d_flags, void* stream) {
    if (n < 0 || (n > 0 && !d_boards)) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_eval_terms(d_boards, n, ignore_checkmate, d_terms, d_score, d_flags, (cudaStream_t)stream);
}

int cdq_attack_maps(const CdqBoard* d_boards, int64_t n, uint64_t* d_maps, void* stream) {
    if (n < 0 || (n > 0 && (!d_boards || !d_maps))) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_attack_maps(d_boards, n, (unsigned long long*)d_maps, (cudaStream_t)stream);
}

int cdq_encode_planes(const CdqBoard* d_boards, int64_t n, float* d_planes, void* stream) {
    if (n < 0 || (n > 0 && (!d_boards || !d_planes))) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_encode_planes(d_boards, n, d_planes, (cudaStream_t)stream);
}

int cdq_random_playouts(CdqBoard* d_boards, int64_t n, uint64_t seed, int max_plies, void* stream) {
    if (n < 0 || max_plies < 0 || (n > 0 && !d_boards)) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_random_playouts(d_boards, n, seed, max_plies, (cudaStream_t)stream);
}

int cdq_expand(const CdqBoard* d_boards, int64_t n, int max_children, CdqBoard* d_children, int32_t* d_counts,
               void* stream) {
    if (n < 0 || max_children < 0 || (n > 0 && (!d_boards || !d_children || !d_counts))) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_expand(d_boards, n, max_children, d_children, d_counts, nullptr, (cudaStream_t)stream);
}

int cdq_expand_weighted(const CdqBoard* d_boards, int64_t n, int max_children, CdqBoard* d_children, int32_t* d_counts,
                        uint8_t* d_weights, void* stream) {
    if (n < 0 || max_children < 0 || (n > 0 && (!d_boards || !d_children || !d_counts || !d_weights))) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_expand(d_boards, n, max_children, d_children, d_counts, d_weights, (cudaStream_t)stream);
}

int cdq_sample_children(const CdqBoard* d_children, const int32_t* d_counts, const uint8_t* d_weights, int64_t n,
                        int max_children, int width, const uint64_t* d_seeds, CdqBoard* d_selected, int32_t* d_sel_counts,
                        void* stream) {
    if (n < 0 || max_children < 1 || max_children > 224 || width < 1 || width > max_children ||
        (n > 0 && (!d_children || !d_counts || !d_weights || !d_seeds || !d_selected || !d_sel_counts)))
        return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_sample_children(d_children, d_counts, d_weights, n, max_children, width,
                                  (const unsigned long long*)d_seeds, d_selected, d_sel_counts, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ network handle
static const size_t W1P_HALFS = 9 * 32 * 16, W2P_HALFS = 9 * 64 * 32, FC1P_HALFS = 64ull * 256 * 64;

int cdq_net_create(CdqNet** out) {
    if (!out) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    CdqNet* net = new CdqNet();
    const size_t halfs = 2 * W1P_HALFS + 3 * W2P_HALFS + 2 * FC1P_HALFS;
    const size_t bytes = halfs * 2 + (64 + 256 + 256 + 4) * 4;
    cudaError_t e = cudaMalloc(&net->blob, bytes);
    if (e != cudaSuccess) { delete net; return (int)e; }
    __half* h = (__half*)net->blob;
    float* f = (float*)(h + halfs);
    net->p.w1p = h;
    net->p.w2p = h + W1P_HALFS;
    net->p.fc1p = h + W1P_HALFS + W2P_HALFS;
    net->p.fc1tp = h + W1P_HALFS + W2P_HALFS + FC1P_HALFS;
    net->p.w2tp = h + W1P_HALFS + W2P_HALFS + 2 * FC1P_HALFS;
    net->p.wss = h + W1P_HALFS + 2 * W2P_HALFS + 2 * FC1P_HALFS;
    net->p.conv2_b = f;
    net->p.fc1_b = f + 64;
    net->p.fc2_w = f + 64 + 256;
    net->p.fc2_b = f + 64 + 512;
    net->loaded = 0;
    *out = net;
    return 0;
}

int cdq_net_load(CdqNet* net, const CdqWeights* w, void* stream) {
    if (!net || !w || !w->conv1_w || !w->conv1_b || !w->conv2_w || !w->conv2_b || !w->fc1_w || !w->fc1_b ||
        !w->fc2_w || !w->fc2_b)
        return CDQ_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = launch_pack_weights(*w, net->p, st)) return e;
    net->loaded = 1;
    return 0;
}

void cdq_net_destroy(CdqNet* net) {
    if (!net) return;
    cudaFree(net->blob);
    delete net;
}

size_t cdq_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    const long long c = n < FWD_CHUNK ? n : FWD_CHUNK;
    const size_t chunked = (size_t)tiles_of(c) << 20; // one 1 MiB fp16 act2 block per 128 positions
    const size_t piped = n >= PIPE_MIN_POSITIONS ? pipe_workspace_bytes() : 0;   // always: the env switch may change
    return chunked > piped ? chunked : piped;
}

// conv -> fc over chunks of at most FWD_CHUNK positions
static int forward_chunks(const CdqNet* net, const CdqBoard* d_boards, int64_t n, float* d_value,
                          const uint32_t* d_flags, float* d_leaf, void* ws, size_t ws_bytes, __half* dbg_act1,
                          cudaStream_t st) {
    if (!net || !net->loaded || n < 0 || (n > 0 && (!d_boards || !ws))) return CDQ_E_BADARG;
    if (ws_bytes < cdq_workspace_bytes(n)) return CDQ_E_WORKSPACE;
    if ((uintptr_t)ws & 15) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    if (n >= PIPE_MIN_POSITIONS && !dbg_act1 && option(OPT_PIPE) == 1)
        return launch_leaf_pipe(d_boards, n, net->p, d_value, d_flags, d_leaf, ws, st);
    release_persisting_l2();
    for (int64_t off = 0; off < n; off += FWD_CHUNK) {
        const int64_t c = (n - off) < FWD_CHUNK ? (n - off) : FWD_CHUNK;
        if (int e = launch_conv_stack(d_boards + off, c, net->p, (__half*)ws, dbg_act1 ? dbg_act1 + off * 2048 : nullptr, st))
            return e;
        if (int e = launch_fc_head((const __half*)ws, c, net->p, d_value ? d_value + off : nullptr,
                                   d_flags ? d_flags + off : nullptr, d_leaf ? d_leaf + off : nullptr, nullptr, st))
            return e;
    }
    return 0;
}

int cdq_value_fwd(const CdqNet* net, const CdqBoard* d_boards, int64_t n, float* d_value, void* d_workspace,
                  size_t workspace_bytes, void* stream) {
    if (n > 0 && !d_value) return CDQ_E_BADARG;
    return forward_chunks(net, d_boards, n, d_value, nullptr, nullptr, d_workspace, workspace_bytes, nullptr,
                          (cudaStream_t)stream);
}

int cdq_leaf_eval(const CdqNet* net, const CdqBoard* d_boards, int64_t n, float* d_leaf, double* d_score,
                  uint32_t* d_flags, float* d_value_raw, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (n > 0 && (!d_leaf || !d_flags)) return CDQ_E_BADARG;
    if (!net) {
        // no network: _evaluate_position's heuristic branch, tanh(fast_evaluate_position(board) / 10) behind the
        // terminal rules (mcts.py:202-210, 226-229) -- one launch of the evaluation kernel, no workspace
        if (n < 0 || (n > 0 && !d_boards) || d_value_raw) return CDQ_E_BADARG;
        if (int e = check_device()) return e;
        return launch_eval_terms(d_boards, n, 0, nullptr, d_score, d_flags, (cudaStream_t)stream, d_leaf);
    }
    if (int e = cdq_eval_terms(d_boards, n, 0, nullptr, d_score, d_flags, stream)) return e;
    return forward_chunks(net, d_boards, n, d_value_raw, d_flags, d_leaf, d_workspace, workspace_bytes, nullptr,
                          (cudaStream_t)stream);
}

// debug/parity hook: also dumps ReLU(conv1) as fp16 [n][64 squares][32 channels] and act2 tiles
int cdq_debug_conv(const CdqNet* net, const CdqBoard* d_boards, int64_t n, void* d_act1, void* d_act2_tiles,
                   void* stream) {
    if (!net || !net->loaded || n <= 0 || n > FWD_CHUNK || !d_boards || !d_act2_tiles) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_conv_stack(d_boards, n, net->p, (__half*)d_act2_tiles, (__half*)d_act1, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ DQN replay step
size_t cdq_train_workspace_bytes(int64_t n) { return train_workspace_bytes(n); }

int cdq_train_fwd_bwd(const CdqNet* net, const CdqNet* target_net, const CdqWeights* d_weights,
                      const CdqBoard* d_states, const CdqBoard* d_next_states, const float* d_reward,
                      const float* d_done, int64_t n, double mean_over, float gamma, int loss_kind, const CdqWeights* d_grads,
                      float* d_loss, void* d_workspace, size_t workspace_bytes, void* stream) {
    const bool upstream = (loss_kind & 0xff) == CDQ_LOSS_UPSTREAM;      // no target network, no next states, no done flags
    if (!net || !net->loaded || !d_weights || !d_grads || !d_loss || n <= 0 || !d_states || !d_reward || !d_workspace ||
        (!upstream && (!target_net || !target_net->loaded || !d_next_states || !d_done)) ||
        (loss_kind & 0xff) > CDQ_LOSS_UPSTREAM || (loss_kind & ~(0xff | CDQ_TRAIN_WS_PERSISTENT | CDQ_TRAIN_REPACK | CDQ_TRAIN_ZERO_LOSS)))
        return CDQ_E_BADARG;
    if (workspace_bytes < train_workspace_bytes(n)) return CDQ_E_WORKSPACE;
    if (int e = check_device()) return e;
    release_persisting_l2();
    return train_fwd_bwd(net, target_net, *d_weights, d_states, d_next_states, d_reward, d_done, n,
                         mean_over > 0.0 ? mean_over : (double)n, gamma, loss_kind,
                         *d_grads, d_loss, d_workspace, (cudaStream_t)stream);
}

int cdq_train_step(const CdqNet* net, const CdqNet* target_net, const CdqWeights* d_weights, const CdqBoard* d_states,
                   const CdqBoard* d_next_states, const float* d_reward, const float* d_done, int64_t n, double mean_over, float gamma,
                   int loss_kind, const CdqWeights* d_grads, float* d_loss, void* d_workspace, size_t workspace_bytes,
                   const CdqOptimizer* opt, void* stream) {
    if (!net || !target_net || !net->loaded || !target_net->loaded || !d_weights || !d_grads || !d_loss || n <= 0 || !d_states ||
        !d_next_states || !d_reward || !d_done || !d_workspace || !opt || (loss_kind & 0xff) > CDQ_LOSS_HUBER ||
        (loss_kind & ~(0xff | CDQ_TRAIN_WS_PERSISTENT | CDQ_TRAIN_REPACK | CDQ_TRAIN_ZERO_LOSS)))
        return CDQ_E_BADARG;
    if (opt->world < 1 || opt->world > CDQ_MAX_RANKS || opt->rank < 0 || opt->rank >= opt->world || !opt->d_m || !opt->d_v || !opt->d_step ||
        !opt->d_grid_bar || opt->n_params_padded != ((CDQ_NPARAMS + 3) / 4) * 4)
        return CDQ_E_BADARG;
    for (int p = 0; p < opt->world; p++)
        if (!opt->peers.params[p] || !opt->peers.grads[p] || !opt->peers.flags[p]) return CDQ_E_BADARG;
    // the two-launch optimizer addresses fc1.weight by its offset in the flat vector: the views must be that vector's
    if (d_weights->conv1_w != opt->peers.params[opt->rank] || d_weights->fc1_w != opt->peers.params[opt->rank] + 21984 ||
        d_grads->conv1_w != opt->peers.grads[opt->rank] || d_grads->fc1_w != opt->peers.grads[opt->rank] + 21984)
        return CDQ_E_BADARG;
    if (workspace_bytes < train_workspace_bytes(n)) return CDQ_E_WORKSPACE;
    if (int e = check_device()) return e;
    release_persisting_l2();
    return train_fwd_bwd(net, target_net, *d_weights, d_states, d_next_states, d_reward, d_done, n,
                         mean_over > 0.0 ? mean_over : (double)n, gamma, loss_kind, *d_grads, d_loss, d_workspace, (cudaStream_t)stream, opt);
}

int cdq_replay_append(CdqBoard* d_mem_states, CdqBoard* d_mem_next, float* d_mem_reward, float* d_mem_done, int32_t* d_ring,
                      const CdqBoard* h_states, const CdqBoard* h_next, const float* h_reward, const float* h_done, int64_t first_slot,
                      int64_t count, int32_t start, int32_t len, int32_t capacity, void* stream) {
    if (!d_mem_states || !d_mem_next || !d_mem_reward || !d_mem_done || !d_ring || count < 0 || first_slot < 0 ||
        first_slot + count > capacity || len < 0 || len > capacity || start < 0 || start >= (capacity > 0 ? capacity : 1) ||
        (count > 0 && (!h_states || !h_next || !h_reward || !h_done)))
        return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (count > 0) {
        CDQ_CUDA(cudaMemcpyAsync(d_mem_states + first_slot, h_states, count * sizeof(CdqBoard), cudaMemcpyHostToDevice, st));
        CDQ_CUDA(cudaMemcpyAsync(d_mem_next + first_slot, h_next, count * sizeof(CdqBoard), cudaMemcpyHostToDevice, st));
        CDQ_CUDA(cudaMemcpyAsync(d_mem_reward + first_slot, h_reward, count * 4, cudaMemcpyHostToDevice, st));
        CDQ_CUDA(cudaMemcpyAsync(d_mem_done + first_slot, h_done, count * 4, cudaMemcpyHostToDevice, st));
    }
    const int32_t geo[3] = {start, len, capacity};      // pageable source: the runtime stages it before returning
    CDQ_CUDA(cudaMemcpyAsync(d_ring, geo, sizeof(geo), cudaMemcpyHostToDevice, st));
    return 0;
}

int cdq_replay_sample(const CdqBoard* d_mem_states, const CdqBoard* d_mem_next, const float* d_mem_reward, const float* d_mem_done,
                      const int32_t* d_ring, uint64_t seed, uint32_t* d_counter, int64_t batch, CdqBoard* d_states,
                      CdqBoard* d_next_states, float* d_reward, float* d_done, void* stream) {
    if (!d_mem_states || !d_mem_next || !d_mem_reward || !d_mem_done || !d_ring || !d_counter || batch <= 0 || !d_states ||
        !d_next_states || !d_reward || !d_done)
        return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_replay_sample(d_mem_states, d_mem_next, d_mem_reward, d_mem_done, d_ring, seed, d_counter, batch, d_states,
                                d_next_states, d_reward, d_done, (cudaStream_t)stream);
}

uint32_t cdq_replay_permute(uint32_t k, uint32_t len, uint64_t seed, uint32_t counter) {
    return len ? replay_permute_host(k, len, seed, counter) : 0u;
}

int cdq_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, int64_t n_params, int step, float lr,
                  float beta1, float beta2, float eps, float grad_scale, void* stream) {
    if (!d_params || !d_grads || !d_m || !d_v || n_params < 0 || step < 1) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    return launch_adam(d_params, d_grads, d_m, d_v, n_params, step, lr, beta1, beta2, eps, grad_scale, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ host-buffer path (e2e)
// A CdqHostCtx owns two streams and two sets of device buffers on the device that was current when it was
// created; cdq_leaf_eval_host_ctx double-buffers chunks through them.  One call at a time per context (the
// context's mutex serialises callers that share one); independent callers create their own contexts and run
// concurrently.  cdq_leaf_eval_host keeps one lazily created context per device for callers that do not manage one.
struct CdqHostCtx {
    static constexpr long long CHUNK = 2 * FWD_CHUNK;       // two forward chunks
    std::mutex mu;
    int dev = -1;
    cudaStream_t st[2] = {nullptr, nullptr};
    CdqBoard* boards[2] = {nullptr, nullptr};
    float* leaf[2] = {nullptr, nullptr};
    double* score[2] = {nullptr, nullptr};
    uint32_t* flags[2] = {nullptr, nullptr};
    void* ws[2] = {nullptr, nullptr};
    size_t ws_bytes = 0;
    int init() {
        dev = current_device();
        if (dev < 0) return CDQ_E_NODEVICE;
        ws_bytes = cdq_workspace_bytes(CHUNK);
        for (int i = 0; i < 2; i++) {
            CDQ_CUDA(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
            CDQ_CUDA(cudaMalloc(&boards[i], CHUNK * sizeof(CdqBoard)));
            CDQ_CUDA(cudaMalloc(&leaf[i], CHUNK * 4));
            CDQ_CUDA(cudaMalloc(&score[i], CHUNK * 8));
            CDQ_CUDA(cudaMalloc(&flags[i], CHUNK * 4));
            CDQ_CUDA(cudaMalloc(&ws[i], ws_bytes));
        }
        return 0;
    }
    void release() {
        for (int i = 0; i < 2; i++) {
            if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
            cudaFree(boards[i]); cudaFree(leaf[i]); cudaFree(score[i]); cudaFree(flags[i]); cudaFree(ws[i]);
        }
    }
};

int cdq_host_ctx_create(CdqHostCtx** out) {
    if (!out) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    CdqHostCtx* c = new CdqHostCtx();
    if (int e = c->init()) { c->release(); delete c; return e; }
    *out = c;
    return 0;
}

void cdq_host_ctx_destroy(CdqHostCtx* ctx) {
    if (!ctx) return;
    int prev = -1;
    cudaGetDevice(&prev);
    if (ctx->dev >= 0) cudaSetDevice(ctx->dev);
    ctx->release();
    if (prev >= 0) cudaSetDevice(prev);
    delete ctx;
}

int cdq_leaf_eval_host_ctx(CdqHostCtx* ctx, const CdqNet* net, const CdqBoard* h_boards, int64_t n, float* h_leaf,
                           double* h_score, uint32_t* h_flags) {
    if (!ctx || n < 0 || (n > 0 && (!h_boards || !h_leaf))) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    if (current_device() != ctx->dev) return CDQ_E_BADARG;     // the context's buffers live on another device
    std::lock_guard<std::mutex> lock(ctx->mu);
    // Chunks ramp up (1/8, 1/4, 1/2 of the full chunk, then full chunks): the first chunk's host-to-device copy is
    // the one transfer nothing can hide, so it is kept short; from then on copies run under the other stream's kernels.
    const bool ramp = option(OPT_HOST_RAMP) != 0;              // 0: full chunks from the start (A/B)
    int c_idx = 0;
    for (int64_t off = 0; off < n; c_idx++) {
        const int s = c_idx & 1;
        const int64_t cap = c_idx < 3 && ramp ? CdqHostCtx::CHUNK >> (3 - c_idx) : CdqHostCtx::CHUNK;
        const int64_t c = (n - off) < cap ? (n - off) : cap;
        cudaStream_t st = ctx->st[s];
        // the slot's previous use (two chunks ago) is ordered before this one by the stream itself
        CDQ_CUDA(cudaMemcpyAsync(ctx->boards[s], h_boards + off, c * sizeof(CdqBoard), cudaMemcpyHostToDevice, st));
        if (int e = cdq_leaf_eval(net, ctx->boards[s], c, ctx->leaf[s], h_score ? ctx->score[s] : nullptr, ctx->flags[s], nullptr,
                                  ctx->ws[s], ctx->ws_bytes, st))
            return e;
        CDQ_CUDA(cudaMemcpyAsync(h_leaf + off, ctx->leaf[s], c * 4, cudaMemcpyDeviceToHost, st));
        if (h_score) CDQ_CUDA(cudaMemcpyAsync(h_score + off, ctx->score[s], c * 8, cudaMemcpyDeviceToHost, st));
        if (h_flags) CDQ_CUDA(cudaMemcpyAsync(h_flags + off, ctx->flags[s], c * 4, cudaMemcpyDeviceToHost, st));
        off += c;
    }
    CDQ_CUDA(cudaStreamSynchronize(ctx->st[0]));
    CDQ_CUDA(cudaStreamSynchronize(ctx->st[1]));
    return 0;
}

namespace {
std::mutex g_default_ctx_mu;
CdqHostCtx* g_default_ctx[MAX_DEVICES];
} // namespace

int cdq_leaf_eval_host(const CdqNet* net, const CdqBoard* h_boards, int64_t n, float* h_leaf, double* h_score,
                       uint32_t* h_flags) {
    if (n < 0 || (n > 0 && (!h_boards || !h_leaf))) return CDQ_E_BADARG;
    if (int e = check_device()) return e;
    const int dev = current_device();
    CdqHostCtx* ctx;
    {
        std::lock_guard<std::mutex> l(g_default_ctx_mu);
        if (!g_default_ctx[dev])
            if (int e = cdq_host_ctx_create(&g_default_ctx[dev])) return e;
        ctx = g_default_ctx[dev];
    }
    return cdq_leaf_eval_host_ctx(ctx, net, h_boards, n, h_leaf, h_score, h_flags);
}

void cdq_host_release(void) {
    std::lock_guard<std::mutex> l(g_default_ctx_mu);
    for (int d = 0; d < MAX_DEVICES; d++) {
        cdq_host_ctx_destroy(g_default_ctx[d]);
        g_default_ctx[d] = nullptr;
    }
}

} // extern "C"
