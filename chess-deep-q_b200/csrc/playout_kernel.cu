// playout_kernel.cu -- synthetic workload generator: n random playouts from the
// start position, one thread per game, entirely on the device.  It reuses the set-wise legal move
// generator of bitboard.cuh with a visitor that SELECTS the r-th legal move instead of counting,
// so the generator that makes the benchmark inputs is the same code K1's mobility term is built
// from (and the tests check every generated position against the CPU oracle).
//
// Recipe per game i: plies = U{0..max_plies}; each ply picks a uniformly random legal move
// (counter-based splitmix64 keyed by (seed, i, ply)); stops early when there is no legal move or
// the material is insufficient.  Reference context: games cap at 200 plies (chess_ai.py:117) and
// the reference seeds everything with 42 (main.py:19-21).
#include "movegen.cuh"
#include "common.cuh"

namespace cdq {

__global__ void __launch_bounds__(128)
random_playouts_kernel(CdqBoard* __restrict__ boards, long long n, u64 seed, int max_plies) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pos p = start_position();
    const u64 key = mix64(seed ^ mix64((u64)i));
    const int plies = (int)(mix64(key) % (u64)(max_plies + 1));
    for (int k = 0; k < plies; k++) {
        const u64 occ = p.co[0] | p.co[1];
        const Props q = make_props(~occ);
        const bool white = p.meta & CDQ_META_TURN_WHITE;
        const u64 att_opp = white ? attack_map<0>(p, q) : attack_map<1>(p, q);
        CountVisitor cv;
        if (white) gen_side<1>(p, q, att_opp, cv); else gen_side<0>(p, q, att_opp, cv);
        if (cv.n == 0 || insufficient_material(p)) break;
        SelectVisitor sv((int)(mix64(key + 1 + (u64)k) % (u64)cv.n), occ, white);
        sv.king_bb = p.K & p.co[white ? 1 : 0];
        if (white) gen_side<1>(p, q, att_opp, sv); else gen_side<0>(p, q, att_opp, sv);
        apply_move(p, sv);
    }
    store_pos(p, reinterpret_cast<u64*>(boards) + i * 9);
}

// All legal successors of each position (side to move), the building block of a batched tree
// expansion (mcts.py:166-194 calls list(board.legal_moves) + push per child).  children is
// [n][max_children]; counts[i] = number of legal moves (children beyond max_children are dropped).
__global__ void __launch_bounds__(128)
expand_kernel(const CdqBoard* __restrict__ boards, long long n, int max_children,
              CdqBoard* __restrict__ children, int* __restrict__ counts, unsigned char* __restrict__ weights) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Pos p = load_pos(reinterpret_cast<const u64*>(boards) + i * 9);
    const bool white = p.meta & CDQ_META_TURN_WHITE;
    const Pos wv = white_view(p);            // one generator instance for both colours (movegen.cuh): its order is the child order
    MoveList L;
    list_moves(wv, L);
    counts[i] = L.n;
    const int m = L.n < max_children ? L.n : max_children;
    for (int r = 0; r < m; r++) {
        const PlainMove mv = pick_from_list(L, r);
        Pos c = p;
        apply_move(c, real_move(mv, white));
        store_pos(c, reinterpret_cast<u64*>(children) + (i * max_children + r) * 9);
        if (weights) weights[i * max_children + r] = (unsigned char)white_view_move_weight(wv, mv, L.att_opp);
    }
}

// select_weighted_moves (evaluation.py:342-366) on the device: `width` successive draws without
// replacement, each with probability proportional to the tactical weight of the remaining moves.
// One warp per node; lane l holds the weights of children l, l+32, ... l+192.  Integer arithmetic
// only -- draw k uses z = mix64(seed + (k+1) * 0x9E3779B97F4A7C15), r = (z * W) >> 64 with W the sum
// of the remaining weights, and takes the first child (in index order) whose running weight sum
// exceeds r -- so a host restatement reproduces the selection exactly (tests/test_gpu_mcts.py).
// Nodes with at most `width` children keep all of them, in order.
__global__ void __launch_bounds__(128)
sample_children_kernel(const CdqBoard* __restrict__ children, const int* __restrict__ counts,
                       const unsigned char* __restrict__ weights, long long n, int max_children, int width,
                       const u64* __restrict__ seeds, CdqBoard* __restrict__ selected, int* __restrict__ sel_counts) {
    const int lane = threadIdx.x & 31;
    const long long node = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (node >= n) return;
    constexpr int NB = 7;                           // 224 >= the 218 legal moves a chess position can have
    const int c = min(counts[node], min(max_children, 32 * NB));
    const u64* src = reinterpret_cast<const u64*>(children + node * max_children);
    u64* dst = reinterpret_cast<u64*>(selected + node * width);
    const int keep = min(c, width);
    if (lane == 0) sel_counts[node] = keep;
    if (c <= width) {
        for (int w = lane; w < c * 9; w += 32) dst[w] = src[w];
        return;
    }
    unsigned wt[NB];
    unsigned total = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        const int idx = lane + 32 * j;
        wt[j] = idx < c ? weights[node * max_children + idx] : 0u;
    }
    const u64 seed = seeds[node];
    for (int k = 0; k < width; k++) {
        // remaining weight per 32-child block j (children are ordered j-major: idx = lane + 32 j)
        unsigned blk[NB];
        total = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) { blk[j] = __reduce_add_sync(0xffffffffu, wt[j]); total += blk[j]; }
        const u64 z = mix64(seed + (u64)(k + 1) * 0x9E3779B97F4A7C15ull);
        unsigned r = (unsigned)__umul64hi(z, (u64)total);
        int j = 0;                                  // block that contains the r-th unit of weight
#pragma unroll
        for (int jj = 0; jj < NB - 1; jj++) if (j == jj && r >= blk[jj]) { r -= blk[jj]; j++; }
        unsigned mine = 0;
#pragma unroll
        for (int jj = 0; jj < NB; jj++) if (jj == j) mine = wt[jj];
        unsigned incl = mine;                       // inclusive prefix sum over lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, incl > r);
        const int pick_lane = __ffs(hit) - 1, pick = pick_lane + 32 * j;
#pragma unroll
        for (int jj = 0; jj < NB; jj++) if (jj == j && lane == pick_lane) wt[jj] = 0;
        if (lane < 9) dst[k * 9 + lane] = src[pick * 9 + lane];
    }
}

int launch_sample_children(const CdqBoard* children, const int* counts, const unsigned char* weights, long long n,
                           int max_children, int width, const unsigned long long* seeds, CdqBoard* selected,
                           int* sel_counts, cudaStream_t st) {
    if (n == 0) return 0;
    ProfileScope ps(KK_PLAYOUT, st);
    sample_children_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(children, counts, weights, n, max_children, width, seeds,
                                                                   selected, sel_counts);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_expand(const CdqBoard* boards, long long n, int max_children, CdqBoard* children, int* counts,
                  unsigned char* weights, cudaStream_t st) {
    if (n == 0) return 0;
    ProfileScope ps(KK_PLAYOUT, st);
    expand_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(boards, n, max_children, children, counts, weights);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_random_playouts(CdqBoard* boards, long long n, unsigned long long seed, int max_plies, cudaStream_t st) {
    if (n == 0) return 0;
    ProfileScope ps(KK_PLAYOUT, st);
    random_playouts_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(boards, n, seed, max_plies);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
