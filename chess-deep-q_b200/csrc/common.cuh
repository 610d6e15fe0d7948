// common.cuh -- shared host-side helpers of libcdq.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <mutex>
#include "../../include/cdq.h"

namespace cdq {

// ---- per-device state.  Nothing in the library assumes "the device that was current first": the SM count, the
// one-time cudaFuncSetAttribute calls, the training side stream and the host-path buffers are all keyed by the
// CURRENT device of the calling thread, so one process may drive several GPUs.
constexpr int MAX_DEVICES = 64;
int current_device();                 // cudaGetDevice; -1 without a device
int device_sm_count();                // of the current device (cached per device)
std::mutex& device_mutex(int dev);
unsigned& device_once_mask(int dev);
enum OnceSlot { ONCE_CONV_STREAM, ONCE_FC_PAIRS, ONCE_FC_SPLIT, ONCE_PIPE, ONCE_FC1_WGRAD, ONCE_FC1_DGRAD, ONCE_CONV2_BWD,
                ONCE_CONV1_WGRAD, ONCE_TREE, ONCE_FWD2, ONCE_SPARE0, ONCE_SPARE1 };
template <class F>
inline void once_per_device(int slot, F&& f) {      // e.g. raising a kernel's dynamic shared memory limit
    const int dev = current_device();
    if (dev < 0) return;
    std::lock_guard<std::mutex> l(device_mutex(dev));
    unsigned& m = device_once_mask(dev);
    if (!((m >> slot) & 1u)) { f(); m |= 1u << slot; }
}

// ---- run-time switches (cdq_set_option / cdq_get_option).  Each is initialised ONCE from its environment variable
// when the library is first used; no call path reads the environment afterwards.
enum Option { OPT_PIPE, OPT_FC_SPLIT, OPT_HOST_RAMP, OPT_DP_TIMEOUT_S, OPT_TRAIN_SPLIT_OPT, OPT_COUNT };
int option(int which);

// every kernel launch made by the library bumps this (bench.py reports it as gpu_launches)
void count_launch(int k = 1);

// optional per-kernel timing with CUDA events on the launching stream (cdq_profile_*)
enum KernelKind { KK_EVAL = 0, KK_CONV = 1, KK_FC = 2, KK_ENCODE = 3, KK_PLAYOUT = 4, KK_PACK = 5, KK_TRAIN = 6, KK_PIPE = 7, KK_COUNT = 8,
                  // the training kernels one by one (cdq_profile_collect_ex; cdq_profile_collect folds them into KK_TRAIN)
                  KK_HEAD_BWD = 8, KK_FC1_WGRAD = 9, KK_FC1_DGRAD = 10, KK_CONV2_BWD = 11, KK_CONV1_WGRAD = 12, KK_ADAM = 13,
                  KK_COUNT_EX = 16 };
struct ProfileScope {
    int kind; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    ProfileScope(int kind, cudaStream_t st);
    ~ProfileScope();
};

#define CDQ_CUDA(expr)                          \
    do {                                        \
        cudaError_t _e = (expr);                \
        if (_e != cudaSuccess) return (int)_e;  \
    } while (0)

// launchers implemented in the .cu files
int launch_eval_terms(const CdqBoard* boards, long long n, int ignore_checkmate, int* terms, double* score,
                      unsigned* flags, cudaStream_t st, float* heuristic_leaf = nullptr);
int launch_attack_maps(const CdqBoard* boards, long long n, unsigned long long* maps, cudaStream_t st);
int launch_encode_planes(const CdqBoard* boards, long long n, float* planes, cudaStream_t st);
int launch_random_playouts(CdqBoard* boards, long long n, unsigned long long seed, int max_plies, cudaStream_t st);

int launch_expand(const CdqBoard* boards, long long n, int max_children, CdqBoard* children, int* counts,
                  unsigned char* weights, cudaStream_t st);

int launch_sample_children(const CdqBoard* children, const int* counts, const unsigned char* weights, long long n,
                           int max_children, int width, const unsigned long long* seeds, CdqBoard* selected,
                           int* sel_counts, cudaStream_t st);

} // namespace cdq
