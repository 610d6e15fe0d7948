// umma.cuh -- thin inline-PTX layer for sm_100a: mbarrier, bulk async copy (TMA engine, 1-D),
// tcgen05 MMA / commit / TMEM alloc / TMEM load, and the shared-memory + instruction descriptors.
//
// Descriptor encodings (matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64); instruction descriptor: c_format [4,6), a_format [7,10),
// b_format [10,13), a_major 15, b_major 16, N>>3 [17,23), M>>4 [24,29)) follow the PTX ISA tables
// for tcgen05.mma.  All operands here are K-major, no swizzle: a "core matrix" is 8 rows x 16 bytes
// stored as 128 contiguous bytes; SBO is the byte stride between 8-row groups (M or N direction),
// LBO the byte stride between 16-byte K chunks.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace cdq {

#ifndef CDQ_MBAR_SPIN_LIMIT
#define CDQ_MBAR_SPIN_LIMIT (1u << 26) // a wait that spins this long traps instead of hanging the GPU
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// true in exactly one lane of a fully converged warp; ptxas then knows the guarded region runs
// in a single thread and emits tcgen05 instructions without a per-lane election loop
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > CDQ_MBAR_SPIN_LIMIT) __trap();
    }
}

// Whole-warp wait.  All lanes poll: measured on B200, letting only lane 0 poll while the others
// park on __syncwarp DOUBLED the conv kernel's time (3.47 -> 7.38 ms per 1M positions), so the
// single-lane form is kept only as an A/B switch.
__device__ __forceinline__ void mbar_wait_warp(int lane, uint64_t* bar, uint32_t parity) {
#ifndef CDQ_ONE_LANE_POLLS
    (void)lane;
    mbar_wait(bar, parity);
#else
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
#endif
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- 1-D bulk copy global -> shared
// shared -> global bulk copy (one bulk async-group per call site: commit, then wait for the READS before the buffer is reused)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address to *smem_slot
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// fp16 x fp16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with per-operand majorness: *_mn = true means the operand is MN-major (8 consecutive M/N
// elements form the 16-byte vector, K advances by 16 bytes inside a 128-byte core matrix)
__host__ __device__ constexpr uint32_t make_idesc_f16_ex(int M, int N, bool a_mn, bool b_mn) {
    return make_idesc_f16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; single thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// A-collector form: consecutive MMAs that read the SAME A operand keep it in the collector
// instead of fetching it from shared memory again.  MODE 0 = fill, 1 = use, 2 = lastuse.
// Measured (tests/cuda/umma_bw_probe.cu): N=64 + N=128 on one A cost 96 cycles with
// fill/lastuse (= one N=192 MMA, the math floor) against 112 without the hint.
template <int MODE>
__device__ __forceinline__ void umma_f16_ca(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
#define CDQ_CA_ASM(M)                                                                                    \
    asm volatile(                                                                                        \
        "{\n\t.reg .pred p;\n\t"                                                                         \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                      \
        "tcgen05.mma.cta_group::1.kind::f16.collector::a::" M " [%0], %1, %2, %3, p;\n\t}"               \
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory")
    if (MODE == 0) CDQ_CA_ASM("fill");
    else if (MODE == 1) CDQ_CA_ASM("use");
    else CDQ_CA_ASM("lastuse");
#undef CDQ_CA_ASM
}
// Weight-stationary form: B is held in collector buffer b0 across MMAs.  MODE 0 = fill (read B
// from shared memory and keep it), 1 = use (reuse the kept B), 2 = lastuse, 3 = discard.
template <int MODE>
__device__ __forceinline__ void umma_f16_ws(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
#define CDQ_WS_ASM(M)                                                                                    \
    asm volatile(                                                                                        \
        "{\n\t.reg .pred p;\n\t"                                                                         \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                      \
        "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::" M " [%0], %1, %2, %3, p;\n\t}"           \
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory")
    if (MODE == 0) CDQ_WS_ASM("fill");
    else if (MODE == 1) CDQ_WS_ASM("use");
    else if (MODE == 2) CDQ_WS_ASM("lastuse");
    else CDQ_WS_ASM("discard");
#undef CDQ_WS_ASM
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// ... 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// relu + round-to-nearest fp16 pack of two fp32 (low half = a)
__device__ __forceinline__ uint32_t relu_pack_f16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

} // namespace cdq
