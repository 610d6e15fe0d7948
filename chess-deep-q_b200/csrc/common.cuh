// common.cuh -- shared host-side helpers of libcdq.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/cdq.h"

namespace cdq {

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
                      unsigned* flags, cudaStream_t st);
int launch_attack_maps(const CdqBoard* boards, long long n, unsigned long long* maps, cudaStream_t st);
int launch_encode_planes(const CdqBoard* boards, long long n, float* planes, cudaStream_t st);
int launch_random_playouts(CdqBoard* boards, long long n, unsigned long long seed, int max_plies, cudaStream_t st);

int launch_expand(const CdqBoard* boards, long long n, int max_children, CdqBoard* children, int* counts,
                  unsigned char* weights, cudaStream_t st);

int launch_sample_children(const CdqBoard* children, const int* counts, const unsigned char* weights, long long n,
                           int max_children, int width, const unsigned long long* seeds, CdqBoard* selected,
                           int* sel_counts, cudaStream_t st);

} // namespace cdq
