// tree_kernel.cu -- SURVEY.md 8(f1): the caller of the leaf evaluator, ParallelRussianDollMCTS (mcts.py:14-194), with
// its trees RESIDENT IN HBM for thousands of concurrent searches (one per self-play game) and every phase a kernel:
//
//   tree_level_kernel   the descent of one wave for all trees, level by level inside ONE launch: node lookup / expansion (_expand_node, mcts.py:166-194:
//                       legal moves, categorize_moves weights, select_weighted_moves sampling), action selection
//                       (_select_action, mcts.py:34-67: unexplored children first, else the annealed UCB argmax), Board.push,
//                       halfmove clock and repetition bookkeeping (is_game_over / is_repetition, mcts.py:151,209)
//   cdq_leaf_eval       the wave's leaves of ALL trees as one batch (the hot path: K1 + K2)
//   tree_backup_kernel  running-mean backup with the sign flip per ply (mcts.py:158-164)
//   tree_best_kernel    most visited root child (mcts.py:100-116); game_advance_kernel plays it in the game
//
// The launch sequence of a search is static (waves x (descent + leaf evaluation + backup)), so the host enqueues it -- or
// replays it as a CUDA graph -- without reading anything back: one synchronisation per SEARCH, not per level or simulation.
//
// Layout.  A block = one tree, a warp = one simulation of the current wave (the reference's worker thread), lanes = the
// node's children (width <= 32) or the node's legal moves during expansion (lane l owns moves l, l+32, ... of <= 224).
// Per tree: an open-addressing hash table node key -> node index (the reference keys nodes by FEN: position + halfmove clock
// + move number, so the key hashes those), per node a child count and a bit mask of children no simulation has taken yet, per
// edge the move, N and Q (fp64 running mean like the reference's Python floats).  Per simulation: its current record (the
// array of these IS the leaf batch), depth, path, halfmove clock and the transposition keys along its path.
//
// Repetition.  python-chess's is_repetition walks the move stack back to the last irreversible move comparing transposition
// keys (pieces, turn, clean castling rights, ep square only if an en-passant capture is legal).  A game carries the keys of
// its positions since the last irreversible move (g_hist), a simulation the keys along its path and the index from which they
// count (valid_from); is_repetition(3) becomes the record's rep3 bit for the leaf evaluation (mcts.py:209), is_repetition(5)
// and the 75-move rule end the descent like is_game_over() does (mcts.py:151).
//
// Determinism.  Every random draw is a counter-based hash of (search seed, node key | wave, worker, depth), and simulations
// of one wave take their turns in worker order, so a tree's result does not depend on which other trees share the batch and
// a host restatement reproduces it (tests/test_gpu_forest.py).
#include <cmath>
#include <vector>
#include "movegen.cuh"
#include "common.cuh"
#include "cnn.cuh"

namespace cdq {
namespace tree {

constexpr int MAX_DEPTH = 8, MAX_WIDTH = 32, MAX_WORKERS = 32, HIST_MAX = 256, MOVE_BLOCKS = 7;
constexpr unsigned NO_MOVE = 0xffffffffu;
constexpr u64 GOLD = 0x9E3779B97F4A7C15ull;

struct Forest {
    int T, workers, depth, width, max_nodes, ht_size;
    // games
    CdqBoard* g_rec; int* g_hmc; int* g_ply; int* g_active; u64* g_seed; u64* g_hist; int* g_hist_len; u64* g_key;
    unsigned char* g_ep_legal; unsigned* g_status;
    // trees
    u64* n_key; int* n_nchild; unsigned* n_unexp; unsigned* e_move; int* e_n; double* e_q; int* ht; int* n_count;
    long long* t_evaluated;
    // simulations
    CdqBoard* s_rec; int* s_node; int* s_depth; int* s_alive; int* s_hmc; int* s_valid_from; u64* s_keys; int* s_path_node;
    int* s_path_child; unsigned char* s_ep_legal;
    float* leaf; unsigned* flags;
    // results of the last search
    unsigned* best_move; int* best_visits; int* root_nchild;
};

// ---- move codes: from | to << 6 | (promotion piece type + 1) << 12 | special << 16
template <class M>
__device__ __forceinline__ unsigned encode_move(const M& m) {
    return (unsigned)m.from | ((unsigned)m.to << 6) | ((unsigned)(m.promo + 1) << 12) | ((unsigned)m.special << 16);
}
__device__ __forceinline__ PlainMove decode_move(unsigned code) {
    PlainMove m;
    m.found = true;
    m.from = code & 63; m.to = (code >> 6) & 63; m.promo = (int)((code >> 12) & 15) - 1; m.special = (code >> 16) & 7;
    return m;
}

// ---- Board.has_legal_en_passant(): the shared generator instance, nothing picked
__device__ __forceinline__ bool has_legal_ep(const Pos& p) {
    if (((p.meta >> CDQ_META_EP_SHIFT) & CDQ_META_EP_MASK) <= 1) return false;
    MoveList L;
    list_classes(white_view(p), L);          // the target sets only: the enumeration (first[], n) is not needed here
    return L.t[MC_EP] != 0;
}

// Board._transposition_key() as a 64-bit hash: the eight bitboards, turn, clean castling rights, ep square if capturable
__device__ __forceinline__ u64 transposition_key(const Pos& p, bool ep_legal) {
    u64 h = 0x243F6A8885A308D3ull;
    h = mix64(h ^ p.P); h = mix64(h ^ p.N); h = mix64(h ^ p.B); h = mix64(h ^ p.R); h = mix64(h ^ p.Q); h = mix64(h ^ p.K);
    h = mix64(h ^ p.co[1]); h = mix64(h ^ p.co[0]);
    const u64 ep = ep_legal ? ((p.meta >> CDQ_META_EP_SHIFT) & CDQ_META_EP_MASK) : 0ull;
    return mix64(h ^ (p.meta & 0x1full) ^ (ep << 8));
}
// what Board.fen() adds to that: halfmove clock and move number (here: the ply, which fixes turn and move number)
__device__ __forceinline__ u64 node_key(u64 tkey, int hmc, int ply) {
    return mix64(tkey ^ mix64(((u64)(unsigned)hmc << 32) | (u64)(unsigned)ply | 0xA5A5000000000000ull));
}

__device__ __forceinline__ int ht_lookup(const Forest& f, int t, u64 key) {
    const int mask = f.ht_size - 1;
    const int* ht = f.ht + (long long)t * f.ht_size;
    const u64* keys = f.n_key + (long long)t * f.max_nodes;
    int slot = (int)(key & (u64)mask);
    for (int probes = 0; probes < f.ht_size; probes++) {
        const int v = ht[slot];
        if (v == 0) return -1;
        if (keys[v - 1] == key) return v - 1;
        slot = (slot + 1) & mask;
    }
    return -1;
}
__device__ __forceinline__ void ht_insert(const Forest& f, int t, u64 key, int idx) {
    const int mask = f.ht_size - 1;
    int* ht = f.ht + (long long)t * f.ht_size;
    int slot = (int)(key & (u64)mask);
    while (atomicCAS(ht + slot, 0, idx + 1) != 0) slot = (slot + 1) & mask;
}

// number of positions equal to `key` among the positions a walk back from index `cur` may compare with: game history
// [0, hist_len), the search root at hist_len, the simulation's path at hist_len + d; only indices >= valid_from count.
__device__ __forceinline__ int count_repeats(const u64* hist, int hist_len, const u64* path_keys, int valid_from, int cur, u64 key,
                                             int lane) {
    int c = 0;
    for (int i = valid_from + lane; i < cur; i += 32) {
        const u64 k = i < hist_len ? hist[i] : path_keys[i - hist_len];
        c += (k == key);
    }
    return __reduce_add_sync(0xffffffffu, c);
}

// select_weighted_moves (evaluation.py:342-366): `width` successive draws without replacement, each proportional to the weight
// of the remaining moves; draw k takes the first move in index order (index = lane + 32 block) whose running weight sum exceeds
// r = (mix64(seed + (k+1) GOLD) * total) >> 64 -- the arithmetic of sample_children_kernel and of the host restatement.  Every
// lane keeps the INCLUSIVE running sum of its moves; a draw is one ballot, and removing the drawn move is one predicated
// subtraction per block (no prefix scan and no reduction per draw: they were 40 % of the kernel's instructions).
template <int NB>
__device__ __forceinline__ void sample_children(const unsigned (&mv)[MOVE_BLOCKS], const unsigned (&wt)[MOVE_BLOCKS], u64 seed, int width,
                                                unsigned* e_move, int* e_n, double* e_q, int lane) {
    unsigned w[NB], incl[NB], cum[NB];      // w: remaining weights; incl[j]: running sum up to and including move (lane, j); cum[j]: total of blocks <= j
    unsigned base = 0;
#pragma unroll
    for (int j = 0; j < NB; j++) {
        w[j] = wt[j];
        unsigned x = w[j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += up;
        }
        incl[j] = base + x;
        base += __shfl_sync(0xffffffffu, x, 31);
        cum[j] = base;
    }
    unsigned total = base;
    for (int k = 0; k < width; k++) {
        const u64 z = mix64(seed + (u64)(k + 1) * GOLD);
        const unsigned r = (unsigned)__umul64hi(z, (u64)total);
        int j = 0;
#pragma unroll
        for (int jj = 0; jj < NB - 1; jj++) j += (r >= cum[jj]);
        unsigned mine = 0, w_here = 0, m_here = NO_MOVE;
#pragma unroll
        for (int jj = 0; jj < NB; jj++) if (jj == j) { mine = incl[jj]; w_here = w[jj]; m_here = mv[jj]; }
        // the first move of block j whose running sum exceeds r.  A move already drawn (or beyond n) has weight 0, i.e. the
        // running sum of its predecessor, which comes first in the ballot if it exceeds r.
        const int pick_lane = __ffs(__ballot_sync(0xffffffffu, mine > r)) - 1;
        const unsigned drawn = __shfl_sync(0xffffffffu, w_here, pick_lane);
        total -= drawn;
#pragma unroll
        for (int jj = 0; jj < NB; jj++) {
            if (jj > j || (jj == j && lane >= pick_lane)) incl[jj] -= drawn;
            if (jj >= j) cum[jj] -= drawn;
            if (jj == j && lane == pick_lane) w[jj] = 0;
        }
        if (lane == pick_lane) { e_move[k] = m_here; e_n[k] = 0; e_q[k] = 0.0; }
    }
}

// _expand_node (mcts.py:166-194) by one warp: all legal moves of `p` with their tactical weights, lane l owning moves
// l + 32 j, then select_weighted_moves' draw of `width` children without replacement (same integer arithmetic as
// sample_children_kernel).  Returns the number of children stored (0 for a finished game).
//
// The move list lives in SHARED memory, one per warp.  On the stack it was 368 bytes of local memory per LANE for values that
// are the same in every lane: with one expansion per (simulation, level) the tree kernel wrote 742 MB of local memory per
// wave of 1 024 trees (ncu, profiles/r2c_tree_level_summary.txt).  Every lane stores the same value to the same address; the
// __syncwarp()s keep a lane that still reads the previous list out of the way of the next one.  (has_legal_ep, called for
// the few positions with an ep square, keeps its list on the stack.)
__device__ __noinline__ int expand_node(const Forest& f, const Pos& p, int hmc, u64 seed, unsigned* e_move, int* e_n, double* e_q,
                                        int lane, MoveList& L) {
    const bool white = p.meta & CDQ_META_TURN_WHITE;
    const Pos wv = white_view(p);
    __syncwarp();
    list_moves_warp(wv, L, lane);            // one list per warp; lane l picks moves l, l + 32, ...
    // Board(fen).is_game_over(): no legal move (checkmate / stalemate), insufficient material, 75-move rule; a board built
    // from a FEN has no history, so no repetition rule here (mcts.py:172-177)
    if (L.n == 0 || insufficient_material(p) || hmc >= 150) return 0;
    const int n = min(L.n, 32 * MOVE_BLOCKS);
    unsigned mv[MOVE_BLOCKS], wt[MOVE_BLOCKS];
#pragma unroll
    for (int j = 0; j < MOVE_BLOCKS; j++) { mv[j] = NO_MOVE; wt[j] = 0; }
    for (int j = 0; 32 * j < n; j++) {       // warp-uniform trip count
        const PlainMove m = pick_from_list(L, lane + 32 * j);
        if (m.found) {
            const unsigned code = encode_move(real_move(m, white));
            const unsigned weight = (unsigned)white_view_move_weight(wv, m, L.att_opp);
#pragma unroll
            for (int jj = 0; jj < MOVE_BLOCKS; jj++) if (jj == j) { mv[jj] = code; wt[jj] = weight; }
        }
    }
    const int width = f.width;
    if (n <= width) {                       // keep all, in generation order (n <= width <= 32: block 0 only)
        if (lane < n) { e_move[lane] = mv[0]; e_n[lane] = 0; e_q[lane] = 0.0; }
        return n;
    }
    if (n <= 32) sample_children<1>(mv, wt, seed, width, e_move, e_n, e_q, lane);
    else if (n <= 64) sample_children<2>(mv, wt, seed, width, e_move, e_n, e_q, lane);
    else sample_children<MOVE_BLOCKS>(mv, wt, seed, width, e_move, e_n, e_q, lane);
    return width;
}

// Board.push + the bookkeeping is_game_over / is_repetition need, for a game (thread) or a simulation (warp, lane-parallel
// history scan).  Returns the repetition count of the new position (1 = first occurrence).
struct Pushed { Pos c; int hmc; int valid_from; u64 key; bool ep_legal; int reps; };
template <bool WARP>
__device__ __forceinline__ Pushed push_move(const Pos& p, unsigned code, int hmc, bool parent_ep_legal, const u64* hist, int hist_len,
                                            const u64* path_keys, int valid_from, int cur_index, int lane) {
    Pushed o;
    const u64 occ = p.co[0] | p.co[1];
    const bool white = p.meta & CDQ_META_TURN_WHITE;
    const PlainMove m = decode_move(code);
    o.c = p;
    apply_move(o.c, m);
    const u64 fb = 1ull << m.from, tb = 1ull << m.to;
    const bool zeroing = (p.P & fb) || (occ & tb) || m.special == MV_EP;                    // Board.is_zeroing
    o.hmc = zeroing ? 0 : hmc + 1;
    const bool irreversible = zeroing || ((p.meta ^ o.c.meta) & 0x1eull) || parent_ep_legal;   // Board.is_irreversible
    o.ep_legal = has_legal_ep(o.c);
    o.key = transposition_key(o.c, o.ep_legal);
    o.valid_from = irreversible ? cur_index : valid_from;
    if (WARP) o.reps = 1 + count_repeats(hist, hist_len, path_keys, o.valid_from, cur_index, o.key, lane);
    else {
        int c = 0;
        for (int i = o.valid_from; i < cur_index; i++) c += ((i < hist_len ? hist[i] : path_keys[i - hist_len]) == o.key);
        o.reps = 1 + c;
    }
    o.c.meta = (o.c.meta & ~CDQ_META_REP3) | (o.reps >= 3 ? CDQ_META_REP3 : 0ull);
    return o;
}

template <int MAXW>
__global__ void __launch_bounds__(MAXW * 32, MAXW <= 8 ? 2 : 1)
tree_level_kernel(const Forest f, int wave, int wave_size, double ew, double alpha, int max_moves) {
    // per-level exchange between the simulations of the tree, double buffered by level parity: the next level posts into the other
    // buffer, so no barrier is needed at the end of a level (three barriers per level instead of five)
    __shared__ u64 sh_key2[2][MAXW];
    __shared__ int sh_need2[2][MAXW], sh_node2[2][MAXW], sh_new2[2][MAXW], sh_dup2[2][MAXW];
    __shared__ MoveList sh_list[MAXW];                   // the expansion's move list of each simulation (warp)
    const int t = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = f.workers, D = f.depth;
    const long long s = (long long)t * W + w;
    u64* path_keys = f.s_keys + s * (MAX_DEPTH + 1);
    const u64* hist = f.g_hist + (long long)t * HIST_MAX;
    const int hist_len = f.g_hist_len[t];
    const int root_ply = f.g_ply[t];
    const u64 search_seed = mix64(f.g_seed[t] + (u64)(unsigned)root_ply * GOLD);

    // ---- the simulation's state: a fresh copy of the game position (mcts.py:83,121), kept in registers over the levels.
    //      A tree is one CTA and nothing crosses trees, so all levels of a wave run in ONE launch (the first version launched
    //      one kernel per level and carried this state through global memory).
    bool alive = f.g_active[t] != 0 && w < wave_size;
    Pos p = load_pos(reinterpret_cast<const u64*>(f.g_rec + t));
    int hmc = f.g_hmc[t], depth = 0, valid_from = 0;
    bool ep_legal = f.g_ep_legal[t] != 0;
    if (lane == 0) path_keys[0] = f.g_key[t];
    __syncwarp();

  for (int level = 0; level < D; level++) {
    u64* sh_key = sh_key2[level & 1];
    int *sh_need = sh_need2[level & 1], *sh_node = sh_node2[level & 1], *sh_new = sh_new2[level & 1], *sh_dup = sh_dup2[level & 1];

    // ---- phase 1: find the node of the current position, expanding it if this tree has not seen it (mcts.py:128-133)
    u64 key = 0;
    int node = -1;
    if (alive) {
        key = node_key(path_keys[depth], hmc, root_ply + depth);
        node = ht_lookup(f, t, key);
    }
    const bool need = alive && node < 0;
    if (lane == 0) { sh_key[w] = key; sh_need[w] = need; sh_node[w] = node; sh_new[w] = 0; }
    __syncthreads();
    int dup_of = -1, rank = 0;
    if (need) {
        for (int a = 0; a < w; a++) {
            if (!sh_need[a]) continue;
            if (sh_key[a] == key) { dup_of = a; break; }
            bool first = true;
            for (int b = 0; b < a; b++) if (sh_need[b] && sh_key[b] == sh_key[a]) { first = false; break; }
            rank += first;
        }
    }
    if (lane == 0) sh_dup[w] = dup_of;       // a duplicate's node is its original's: sh_node[sh_dup[w]] after the next barrier
    if (need && dup_of < 0) {
        node = f.n_count[t] + rank;
        if (node < f.max_nodes) {
            const long long nb = (long long)t * f.max_nodes + node, eb = nb * f.width;
            const int nchild = expand_node(f, p, hmc, mix64(search_seed ^ key), f.e_move + eb, f.e_n + eb, f.e_q + eb, lane, sh_list[w]);
            if (lane == 0) {
                f.n_key[nb] = key;
                f.n_nchild[nb] = nchild;
                f.n_unexp[nb] = nchild >= 32 ? 0xffffffffu : ((1u << nchild) - 1u);
                __threadfence_block();
                ht_insert(f, t, key, node);
                sh_node[w] = node; sh_new[w] = 1;
            }
        } else {                       // cannot happen with max_nodes = iterations * depth + 1; the simulation just ends here
            node = -1;
            if (lane == 0) sh_node[w] = -1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int added = 0;
        for (int a = 0; a < W; a++) added += sh_new[a];
        f.n_count[t] += added;
    }
    if (dup_of >= 0) node = sh_node[dup_of];
    if (alive && node < 0) alive = false;

    // ---- phase 2: action selection (mcts.py:34-67).  The reference's semantics are sequential -- simulations take their
    //      action in worker order, and a simulation that finds never-taken children claims a random one of them -- but the only
    //      state one simulation changes for another is that mask of never-taken children of THEIR COMMON node (N and Q move at
    //      the back-up, after the wave).  So every simulation replays the claims of the lower-numbered simulations standing on
    //      its node (their draws are counter-based hashes it can recompute) and then takes its own action: all in parallel, no
    //      turn-taking; the highest-numbered simulation on a node stores the mask it leaves behind.
    unsigned move = NO_MOVE;
    unsigned fresh_after = 0;
    bool store_mask = false;
    if (alive) {
        const long long nb = (long long)t * f.max_nodes + node, eb = nb * f.width;
        const int nchild = f.n_nchild[nb];
        if (nchild == 0) alive = false;                  // terminal node: `if not action ... break` (mcts.py:139)
        else {
            auto draw = [&](int worker) {
                return mix64(search_seed ^ mix64(((u64)(unsigned)wave << 32) | ((u64)(unsigned)worker << 8) | (u64)(unsigned)depth |
                                                 0x5E1EC70000000000ull));
            };
            auto kth_bit = [](unsigned m, int k) {
                for (int i = 0; i < k; i++) m &= m - 1;
                return __ffs(m) - 1;
            };
            unsigned fresh = f.n_unexp[nb];
            store_mask = true;
            for (int a = 0; a < W; a++) {
                if (a == w || (sh_dup[a] >= 0 ? sh_node[sh_dup[a]] : sh_node[a]) != node) continue;
                if (a > w) { store_mask = false; continue; }
                if (fresh) fresh &= ~(1u << kth_bit(fresh, (int)__umul64hi(draw(a), (u64)__popc(fresh))));
            }
            const u64 z = draw(w);
            int child;
            if (fresh) {                                 // random.choice(unexplored)
                child = kth_bit(fresh, (int)__umul64hi(z, (u64)__popc(fresh)));
                fresh &= ~(1u << child);
            } else {
                const int nc = lane < nchild ? f.e_n[eb + lane] : 0;
                const int total = __reduce_add_sync(0xffffffffu, nc);
                if (total == 0) child = (int)__umul64hi(z, (u64)nchild);      // random.choice(actions)
                else {
                    // explicit IEEE operations in the reference's order (no FMA contraction), so that the argmax is the
                    // one Python floats give
                    const double beta = max_moves > 0 ? __ddiv_rn((double)(root_ply + depth), (double)max_moves) : 0.5;
                    const double eff = __dmul_rn(ew, __dsub_rn(1.0, __dmul_rn(__dmul_rn(0.5, alpha), beta)));
                    const double nk = __dadd_rn((double)nc, 1e-10);
                    // as written in the reference: the running mean Q is divided by N once more (mcts.py:61-62)
                    double ucb = lane < nchild ? __dadd_rn(__ddiv_rn(f.e_q[eb + lane], nk),
                                                           __dmul_rn(eff, sqrt(__ddiv_rn(log(__dadd_rn((double)total, 1e-10)), nk))))
                                               : -INFINITY;
                    int arg = lane;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {       // np.argmax: the first of equal maxima
                        const double ou = __shfl_xor_sync(0xffffffffu, ucb, d);
                        const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
                        if (ou > ucb || (ou == ucb && oa < arg)) { ucb = ou; arg = oa; }
                    }
                    child = arg;
                }
            }
            fresh_after = fresh;
            move = f.e_move[eb + child];
            if (lane == 0) { f.s_path_node[s * MAX_DEPTH + depth] = node; f.s_path_child[s * MAX_DEPTH + depth] = child; }
        }
    }
    __syncthreads();                                     // every simulation has read its node's mask
    if (alive && store_mask && lane == 0) f.n_unexp[(long long)t * f.max_nodes + node] = fresh_after;

    // ---- phase 3: Board.push, halfmove clock, repetition (all simulations in parallel)
    if (move != NO_MOVE) {
        const Pushed o = push_move<true>(p, move, hmc, ep_legal, hist, hist_len, path_keys, valid_from, hist_len + depth + 1, lane);
        if (lane == 0) path_keys[depth + 1] = o.key;
        __syncwarp();
        p = o.c; hmc = o.hmc; valid_from = o.valid_from; ep_legal = o.ep_legal; depth += 1;
        // fivefold repetition ends the game (is_game_over, mcts.py:151); the 75-move rule and no-legal-move positions
        // end the descent at the next level's expansion, which is the same leaf
        alive = o.reps < 5 && depth < D;
    } else {
        alive = false;                      // finished game, depth limit, or a slot beyond the (partial) wave
    }
  }
    // the leaf: this array IS the batch cdq_leaf_eval evaluates next
    if (lane == 0) {
        store_pos(p, reinterpret_cast<u64*>(f.s_rec + s));
        f.s_depth[s] = depth;
    }
}

// mcts.py:158-164, simulations of the wave in worker order (they update shared edges)
// One WARP per tree, lane = level: the edges of one simulation's path belong to nodes of different plies, so its updates are
// independent and run in parallel; the simulations follow each other (two of them may share an edge, which is then the same
// lane's in program order).  One thread per tree walked 8 x 7 dependent global round trips: 95 us per wave of 1 024 trees.
__global__ void __launch_bounds__(128) tree_backup_kernel(const Forest f, int wave_size) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= f.T || !f.g_active[t]) return;
    for (int w = 0; w < wave_size; w++) {
        const long long s = (long long)t * f.workers + w;
        const int depth = f.s_depth[s];
        if (lane < depth) {
            const int d = lane;
            const double leaf = (double)f.leaf[s];
            const double value = ((depth - 1 - d) & 1) ? -leaf : leaf;          // the sign flips per ply on the way up
            const long long e = ((long long)t * f.max_nodes + f.s_path_node[s * MAX_DEPTH + d]) * f.width + f.s_path_child[s * MAX_DEPTH + d];
            const int n = ++f.e_n[e];
            const double q = f.e_q[e];
            f.e_q[e] = __dadd_rn(q, __ddiv_rn(__dsub_rn(value, q), (double)n));
        }
    }
    if (lane == 0) f.t_evaluated[t] += wave_size;
}

__global__ void __launch_bounds__(128) tree_reset_kernel(const Forest f) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)f.T * f.ht_size) f.ht[i] = 0;
    if (i < f.T) { f.n_count[i] = 0; f.best_move[i] = NO_MOVE; f.best_visits[i] = 0; f.root_nchild[i] = 0; }
    if (i < (long long)f.T * f.workers) { f.s_alive[i] = 0; f.s_depth[i] = 0; }
}

// mcts.py:100-116: the most visited root child (first of equal counts); no visits at all -> a random child
__global__ void __launch_bounds__(128) tree_best_kernel(const Forest f) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.T) return;
    f.best_move[t] = NO_MOVE; f.best_visits[t] = 0; f.root_nchild[t] = 0;
    if (!f.g_active[t] || f.n_count[t] == 0) return;
    const long long nb = (long long)t * f.max_nodes, eb = nb * f.width;       // the root is node 0
    const int nchild = f.n_nchild[nb];
    f.root_nchild[t] = nchild;
    if (nchild == 0) return;
    int best = 0, bv = f.e_n[eb];
    for (int c = 1; c < nchild; c++) if (f.e_n[eb + c] > bv) { bv = f.e_n[eb + c]; best = c; }
    if (bv == 0) {
        const u64 search_seed = mix64(f.g_seed[t] + (u64)(unsigned)f.g_ply[t] * GOLD);
        best = (int)__umul64hi(mix64(search_seed ^ 0xBE57C401CEull), (u64)nchild);
    }
    f.best_move[t] = f.e_move[eb + best];
    f.best_visits[t] = bv;
}

// keys of a game's positions as the host hands them over: history records (oldest first) and the current position
__global__ void __launch_bounds__(128) game_keys_kernel(const Forest f, const CdqBoard* __restrict__ history, const int* __restrict__ hist_len,
                                                        int stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.T) return;
    const Pos p = load_pos(reinterpret_cast<const u64*>(f.g_rec + t));
    const bool ep = has_legal_ep(p);
    f.g_ep_legal[t] = ep;
    f.g_key[t] = transposition_key(p, ep);
    int n = history && hist_len ? hist_len[t] : 0;
    n = n < 0 ? 0 : n > HIST_MAX ? HIST_MAX : n > stride ? stride : n;
    for (int i = 0; i < n; i++) {
        const Pos h = load_pos(reinterpret_cast<const u64*>(history + (long long)t * stride + i));
        f.g_hist[(long long)t * HIST_MAX + i] = transposition_key(h, has_legal_ep(h));
    }
    f.g_hist_len[t] = n;
    f.g_status[t] = 0;
}

// plays the search's move in its game (chess_ai.py:151-152: board.push(move)); status bits: 1 = a move was played,
// 2 = fivefold repetition, 4 = 75-move rule reached (the host combines them with the evaluation kernel's terminal flags)
__global__ void __launch_bounds__(128) game_advance_kernel(const Forest f) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= f.T) return;
    if (!f.g_active[t] || f.best_move[t] == NO_MOVE) { f.g_status[t] &= ~1u; return; }
    const Pos p = load_pos(reinterpret_cast<const u64*>(f.g_rec + t));
    u64* hist = f.g_hist + (long long)t * HIST_MAX;
    int hist_len = f.g_hist_len[t];
    const u64 root_key = f.g_key[t];
    const Pushed o = push_move<false>(p, f.best_move[t], f.g_hmc[t], f.g_ep_legal[t] != 0, hist, hist_len, &root_key, 0, hist_len + 1, 0);
    if (o.valid_from == hist_len + 1) hist_len = 0;                      // irreversible move: nothing before it can repeat
    else if (hist_len < HIST_MAX) hist[hist_len++] = root_key;
    store_pos(o.c, reinterpret_cast<u64*>(f.g_rec + t));
    f.g_hist_len[t] = hist_len;
    f.g_key[t] = o.key; f.g_ep_legal[t] = o.ep_legal; f.g_hmc[t] = o.hmc; f.g_ply[t] += 1;
    f.g_status[t] = 1u | (o.reps >= 5 ? 2u : 0u) | (o.hmc >= 150 ? 4u : 0u);
}

} // namespace tree
} // namespace cdq

using namespace cdq;

template <int MAXW>
static void launch_descent(const tree::Forest& f, int wave, int wave_size, double ew, double alpha, int max_moves, cudaStream_t st) {
    tree::tree_level_kernel<MAXW><<<f.T, f.workers * 32, 0, st>>>(f, wave, wave_size, ew, alpha, max_moves);
}

struct CdqForest {
    tree::Forest f;
    std::vector<void*> allocs;
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int max_iterations = 0;
    int dev = -1;
};

extern "C" {

int cdq_forest_create(const CdqForestConfig* cfg, CdqForest** out) {
    if (!cfg || !out || cfg->n_games < 1 || cfg->workers < 1 || cfg->workers > tree::MAX_WORKERS || cfg->depth < 1 ||
        cfg->depth > tree::MAX_DEPTH || cfg->width < 1 || cfg->width > tree::MAX_WIDTH || cfg->max_iterations < 1)
        return CDQ_E_BADARG;
    const int dev = current_device();
    if (dev < 0) return CDQ_E_NODEVICE;
    CdqForest* F = new CdqForest();
    F->dev = dev;
    F->max_iterations = cfg->max_iterations;
    tree::Forest& f = F->f;
    f.T = cfg->n_games; f.workers = cfg->workers; f.depth = cfg->depth; f.width = cfg->width;
    f.max_nodes = cfg->max_iterations * cfg->depth + 1;        // a simulation expands at most one node per level
    f.ht_size = 64;
    while (f.ht_size < 2 * f.max_nodes) f.ht_size <<= 1;
    int err = 0;
    auto take = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (err) return nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e != cudaSuccess) { err = (int)e; return nullptr; }
        cudaMemset(p, 0, bytes ? bytes : 16);
        F->allocs.push_back(p);
        return p;
    };
    const size_t T = f.T, S = T * f.workers, N = T * f.max_nodes, E = N * f.width;
    f.g_rec = (CdqBoard*)take(T * sizeof(CdqBoard)); f.g_hmc = (int*)take(T * 4); f.g_ply = (int*)take(T * 4);
    f.g_active = (int*)take(T * 4); f.g_seed = (u64*)take(T * 8); f.g_hist = (u64*)take(T * tree::HIST_MAX * 8);
    f.g_hist_len = (int*)take(T * 4); f.g_key = (u64*)take(T * 8); f.g_ep_legal = (unsigned char*)take(T); f.g_status = (unsigned*)take(T * 4);
    f.n_key = (u64*)take(N * 8); f.n_nchild = (int*)take(N * 4); f.n_unexp = (unsigned*)take(N * 4);
    f.e_move = (unsigned*)take(E * 4); f.e_n = (int*)take(E * 4); f.e_q = (double*)take(E * 8);
    f.ht = (int*)take(T * f.ht_size * 4); f.n_count = (int*)take(T * 4); f.t_evaluated = (long long*)take(T * 8);
    f.s_rec = (CdqBoard*)take(S * sizeof(CdqBoard)); f.s_node = (int*)take(S * 4); f.s_depth = (int*)take(S * 4);
    f.s_alive = (int*)take(S * 4); f.s_hmc = (int*)take(S * 4); f.s_valid_from = (int*)take(S * 4);
    f.s_keys = (u64*)take(S * (tree::MAX_DEPTH + 1) * 8); f.s_path_node = (int*)take(S * tree::MAX_DEPTH * 4);
    f.s_path_child = (int*)take(S * tree::MAX_DEPTH * 4); f.s_ep_legal = (unsigned char*)take(S);
    f.leaf = (float*)take(S * 4); f.flags = (unsigned*)take(S * 4);
    f.best_move = (unsigned*)take(T * 4); f.best_visits = (int*)take(T * 4); f.root_nchild = (int*)take(T * 4);
    F->ws_bytes = cdq_workspace_bytes((int64_t)S);
    F->ws = take(F->ws_bytes);
    if (err) { cdq_forest_destroy(F); return err; }
    *out = F;
    return 0;
}

void cdq_forest_destroy(CdqForest* F) {
    if (!F) return;
    int prev = -1;
    cudaGetDevice(&prev);
    if (F->dev >= 0) cudaSetDevice(F->dev);
    for (void* p : F->allocs) cudaFree(p);
    if (prev >= 0) cudaSetDevice(prev);
    delete F;
}

int cdq_forest_set_games(CdqForest* F, const CdqBoard* d_boards, const int32_t* d_halfmove_clock, const int32_t* d_ply,
                         const CdqBoard* d_history, const int32_t* d_history_len, int history_stride, const uint64_t* d_seeds,
                         const int32_t* d_active, void* stream) {
    if (!F || !d_boards || !d_seeds || (d_history && (!d_history_len || history_stride < 1))) return CDQ_E_BADARG;
    if (current_device() != F->dev) return CDQ_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const tree::Forest& f = F->f;
    const size_t T = f.T;
    CDQ_CUDA(cudaMemcpyAsync(f.g_rec, d_boards, T * sizeof(CdqBoard), cudaMemcpyDeviceToDevice, st));
    CDQ_CUDA(cudaMemcpyAsync(f.g_seed, d_seeds, T * 8, cudaMemcpyDeviceToDevice, st));
    if (d_halfmove_clock) CDQ_CUDA(cudaMemcpyAsync(f.g_hmc, d_halfmove_clock, T * 4, cudaMemcpyDeviceToDevice, st));
    else CDQ_CUDA(cudaMemsetAsync(f.g_hmc, 0, T * 4, st));
    if (d_ply) CDQ_CUDA(cudaMemcpyAsync(f.g_ply, d_ply, T * 4, cudaMemcpyDeviceToDevice, st));
    else CDQ_CUDA(cudaMemsetAsync(f.g_ply, 0, T * 4, st));
    if (d_active) CDQ_CUDA(cudaMemcpyAsync(f.g_active, d_active, T * 4, cudaMemcpyDeviceToDevice, st));
    else CDQ_CUDA(cudaMemsetAsync(f.g_active, 1, T * 4, st));           // 0x01010101: non-zero = active
    CDQ_CUDA(cudaMemsetAsync(f.t_evaluated, 0, T * 8, st));
    tree::game_keys_kernel<<<(unsigned)((T + 127) / 128), 128, 0, st>>>(f, d_history, d_history_len, history_stride);
    count_launch();
    return (int)cudaGetLastError();
}

int cdq_forest_set_active(CdqForest* F, const int32_t* d_active, void* stream) {
    if (!F || !d_active) return CDQ_E_BADARG;
    CDQ_CUDA(cudaMemcpyAsync(F->f.g_active, d_active, (size_t)F->f.T * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

int cdq_forest_search(CdqForest* F, const CdqNet* net, int iterations, float exploration_weight, float training_progress,
                      int max_moves, void* stream) {
    if (!F || iterations < 1 || iterations > F->max_iterations || (net && !net->loaded)) return CDQ_E_BADARG;
    if (current_device() != F->dev) return CDQ_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const tree::Forest& f = F->f;
    const long long reset_n = (long long)f.T * (f.ht_size > f.workers ? f.ht_size : f.workers);
    tree::tree_reset_kernel<<<(unsigned)((reset_n + 127) / 128), 128, 0, st>>>(f);
    count_launch();
    const double ew = (double)exploration_weight, alpha = 1.0 + (double)training_progress;
    const long long S = (long long)f.T * f.workers;
    for (int first = 0, wave = 0; first < iterations; first += f.workers, wave++) {
        const int wave_size = iterations - first < f.workers ? iterations - first : f.workers;
        {   // the wave's descent, all levels (selection, expansion, push) in one launch
            ProfileScope ps(KK_PLAYOUT, st);
            if (f.workers <= 8) launch_descent<8>(f, wave, wave_size, ew, alpha, max_moves, st);
            else if (f.workers <= 16) launch_descent<16>(f, wave, wave_size, ew, alpha, max_moves, st);
            else launch_descent<32>(f, wave, wave_size, ew, alpha, max_moves, st);
            count_launch();
        }
        // the wave's leaves of all trees: ONE batched leaf evaluation (K1 + K2; without a network K1's heuristic leaf)
        if (int e = cdq_leaf_eval(net, f.s_rec, S, f.leaf, nullptr, f.flags, nullptr, F->ws, F->ws_bytes, stream)) return e;
        tree::tree_backup_kernel<<<(unsigned)((f.T + 3) / 4), 128, 0, st>>>(f, wave_size);      // a warp per tree
        count_launch();
    }
    tree::tree_best_kernel<<<(unsigned)((f.T + 127) / 128), 128, 0, st>>>(f);
    count_launch();
    return (int)cudaGetLastError();
}

int cdq_forest_advance(CdqForest* F, void* stream) {
    if (!F) return CDQ_E_BADARG;
    if (current_device() != F->dev) return CDQ_E_BADARG;
    tree::game_advance_kernel<<<(unsigned)((F->f.T + 127) / 128), 128, 0, (cudaStream_t)stream>>>(F->f);
    count_launch();
    return (int)cudaGetLastError();
}

int cdq_forest_view(const CdqForest* F, CdqForestView* v) {
    if (!F || !v) return CDQ_E_BADARG;
    const tree::Forest& f = F->f;
    v->n_games = f.T; v->workers = f.workers; v->depth = f.depth; v->width = f.width; v->max_nodes = f.max_nodes;
    v->game_boards = f.g_rec; v->game_halfmove_clock = f.g_hmc; v->game_ply = f.g_ply; v->game_active = f.g_active;
    v->game_status = f.g_status; v->game_history_len = f.g_hist_len;
    v->best_move = f.best_move; v->best_visits = f.best_visits; v->root_children = f.root_nchild;
    v->edge_move = f.e_move; v->edge_visits = f.e_n; v->edge_value = f.e_q; v->node_children = f.n_nchild; v->node_count = f.n_count;
    v->leaf_boards = f.s_rec; v->leaf_values = f.leaf; v->leaf_depth = f.s_depth; v->positions_evaluated = reinterpret_cast<const int64_t*>(f.t_evaluated);
    return 0;
}

} // extern "C"
