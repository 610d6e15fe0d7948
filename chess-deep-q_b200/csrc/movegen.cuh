// movegen.cuh -- move SELECTION and Board.push on top of the set-wise generator of bitboard.cuh, shared by the
// synthetic playout generator, the expansion kernels (playout_kernel.cu) and the device-resident search
// (tree_kernel.cu): a visitor that picks the r-th legal move, the packed-board restatement of Board.push
// (python-chess semantics: raw ep square on every double push, clean castling rights), the counter-based RNG and
// the tactical sampling weight of categorize_moves (evaluation.py:232-339).
#pragma once
#include "bitboard.cuh"

namespace cdq {

__device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ int nth_bit(u64 b, int k) {
    for (int i = 0; i < k; i++) b &= b - 1;
    return lsb_sq(b);
}

enum { MV_NORMAL = 0, MV_CASTLE_K = 1, MV_CASTLE_Q = 2, MV_EP = 3, MV_DOUBLE = 4 };
enum { PT_P = 0, PT_N = 1, PT_B = 2, PT_R = 3, PT_Q = 4, PT_K = 5 };

struct SelectVisitor {
    int r;           // index of the wanted move among the remaining classes
    bool found = false;
    int from = 0, to = 0, promo = -1, special = MV_NORMAL;
    u64 occ;
    bool white;

    __device__ SelectVisitor(int r_, u64 occ_, bool w) : r(r_), occ(occ_), white(w) {}

    __device__ __forceinline__ bool pick(u64 t, int mult, int& sq, int& sub) {
        if (found) return false;
        const int c = popc(t) * mult;
        if (r >= c) { r -= c; return false; }
        sq = nth_bit(t, r / mult);
        sub = r % mult;
        found = true;
        return true;
    }
    template <int D> __device__ __forceinline__ void slide(u64 t) {
        int sq, sub;
        if (!pick(t, 1, sq, sub)) return;
        to = sq;
        const int delta = Dir<D>::left ? Dir<D>::s : -Dir<D>::s;
        int f = sq - delta;
        while (!((occ >> f) & 1)) f -= delta; // origin = first piece behind the target
        from = f;
    }
    template <int J> __device__ __forceinline__ void knight(u64 t) {
        int sq, sub;
        if (!pick(t, 1, sq, sub)) return;
        to = sq; from = sq - knight_delta(J);
    }
    u64 king_bb = 0;
    __device__ __forceinline__ void king(u64 t) {
        int sq, sub;
        if (!pick(t, 1, sq, sub)) return;
        to = sq; from = lsb_sq(king_bb);
    }
    __device__ __forceinline__ void pawn_to(u64 t, int delta, int sp) {
        const u64 last = RANK_1 | RANK_8;
        int sq, sub;
        if (pick(t & ~last, 1, sq, sub)) { to = sq; from = sq - delta; special = sp; return; }
        if (pick(t & last, 4, sq, sub)) { to = sq; from = sq - delta; promo = PT_Q - sub; }
    }
    __device__ __forceinline__ void pawn_push(u64 s, u64 d) {
        pawn_to(s, white ? 8 : -8, MV_NORMAL);
        pawn_to(d, white ? 16 : -16, MV_DOUBLE);
    }
    template <int D> __device__ __forceinline__ void pawn_cap(u64 t) {
        pawn_to(t, Dir<D>::left ? Dir<D>::s : -Dir<D>::s, MV_NORMAL);
    }
    __device__ __forceinline__ void castle(bool k, bool q) {
        int sq, sub;
        const int e = white ? 4 : 60;
        if (pick(k ? 1ull : 0ull, 1, sq, sub)) { from = e; to = e + 2; special = MV_CASTLE_K; return; }
        if (pick(q ? 1ull : 0ull, 1, sq, sub)) { from = e; to = e - 2; special = MV_CASTLE_Q; }
    }
    __device__ __forceinline__ void ep(u64 caps, int epsq) {
        int sq, sub;
        if (pick(caps, 1, sq, sub)) { from = sq; to = epsq; special = MV_EP; }
    }
};

__device__ __forceinline__ u64& type_board(Pos& p, int t) {
    switch (t) {
        case PT_P: return p.P; case PT_N: return p.N; case PT_B: return p.B;
        case PT_R: return p.R; case PT_Q: return p.Q; default: return p.K;
    }
}

// Board.push for standard chess on the packed representation (keeps castling rights clean).
struct PlainMove { int from = 0, to = 0, promo = -1, special = MV_NORMAL; bool found = false; };
template <class M>
__device__ __forceinline__ void apply_move(Pos& p, const M& m) {
    const bool white = p.meta & CDQ_META_TURN_WHITE;
    const int us = white ? 1 : 0, them = us ^ 1;
    const u64 fb = 1ull << m.from, tb = 1ull << m.to;
    int pt = PT_K;
    if (p.P & fb) pt = PT_P; else if (p.N & fb) pt = PT_N; else if (p.B & fb) pt = PT_B;
    else if (p.R & fb) pt = PT_R; else if (p.Q & fb) pt = PT_Q;
    // captured piece (if any) leaves
    p.P &= ~tb; p.N &= ~tb; p.B &= ~tb; p.R &= ~tb; p.Q &= ~tb; p.K &= ~tb;
    p.co[them] &= ~tb;
    // mover
    type_board(p, pt) &= ~fb;
    p.co[us] &= ~fb;
    type_board(p, m.promo >= 0 ? m.promo : pt) |= tb;
    p.co[us] |= tb;
    if (m.special == MV_EP) {
        const u64 cap = white ? (tb >> 8) : (tb << 8);
        p.P &= ~cap; p.co[them] &= ~cap;
    } else if (m.special == MV_CASTLE_K || m.special == MV_CASTLE_Q) {
        const int base = white ? 0 : 56;
        const u64 rf = 1ull << (base + (m.special == MV_CASTLE_K ? 7 : 0));
        const u64 rt = 1ull << (base + (m.special == MV_CASTLE_K ? 5 : 3));
        p.R = (p.R & ~rf) | rt;
        p.co[us] = (p.co[us] & ~rf) | rt;
    }
    // meta: flip turn, clear ep, update rights
    u64 meta = p.meta;
    meta ^= CDQ_META_TURN_WHITE;
    meta &= ~(CDQ_META_EP_MASK << CDQ_META_EP_SHIFT);
    if (m.special == MV_DOUBLE) meta |= (u64)((white ? m.from + 8 : m.from - 8) + 1) << CDQ_META_EP_SHIFT;
    const u64 touched = fb | tb;
    if (touched & (1ull << 7)) meta &= ~CDQ_META_CASTLE_WK;
    if (touched & (1ull << 0)) meta &= ~CDQ_META_CASTLE_WQ;
    if (touched & (1ull << 63)) meta &= ~CDQ_META_CASTLE_BK;
    if (touched & (1ull << 56)) meta &= ~CDQ_META_CASTLE_BQ;
    if (pt == PT_K) meta &= white ? ~(CDQ_META_CASTLE_WK | CDQ_META_CASTLE_WQ) : ~(CDQ_META_CASTLE_BK | CDQ_META_CASTLE_BQ);
    p.meta = meta;
}

__device__ __forceinline__ Pos start_position() {
    Pos p;
    p.P = 0x00ff00000000ff00ull; p.N = 0x4200000000000042ull; p.B = 0x2400000000000024ull;
    p.R = 0x8100000000000081ull; p.Q = 0x0800000000000008ull; p.K = 0x1000000000000010ull;
    p.co[1] = 0xffffull; p.co[0] = 0xffff000000000000ull;
    p.meta = CDQ_META_TURN_WHITE | CDQ_META_CASTLE_WK | CDQ_META_CASTLE_WQ | CDQ_META_CASTLE_BK | CDQ_META_CASTLE_BQ;
    return p;
}

// Sampling weight of a move, categorize_moves (evaluation.py:232-339): first matching category of
// capture of a more valuable piece 10, gives check 9, equal capture 8, moves a threatened piece 7,
// develops a minor piece 6, capture of a less valuable piece 5, to d4/e4/d5/e5 4, castling or a king
// move from its back rank 3, pawn move 2, other 1.  (An en-passant capture lands on an empty square
// and is therefore not a "capture" there.)
__device__ __forceinline__ int piece_value_at(const Pos& p, u64 b) {
    if (p.P & b) return 1;
    if ((p.N | p.B) & b) return 3;
    if (p.R & b) return 5;
    if (p.Q & b) return 9;
    return 0;   // king (PIECE_VALUES, evaluation.py:8-15)
}
__device__ __forceinline__ int move_category_weight(const Pos& p, const Pos& child, const SelectVisitor& m, u64 att_opp, bool white) {
    const u64 fb = 1ull << m.from, tb = 1ull << m.to;
    const u64 occ = p.co[0] | p.co[1];
    if (occ & tb) {
        const int vt = piece_value_at(p, tb), vf = piece_value_at(p, fb);
        return vt > vf ? 10 : (vt == vf ? 8 : 5);
    }
    {   // gives check: the child's side to move is attacked on its king square
        const u64 cocc = child.co[0] | child.co[1];
        const Props cq = make_props(~cocc);
        const u64 att_us = white ? attack_map<1>(child, cq) : attack_map<0>(child, cq);
        if (att_us & child.K & child.co[white ? 0 : 1]) return 9;
    }
    if (att_opp & fb) return 7;
    const int fr = m.from >> 3, ff = m.from & 7;
    if (((p.N | p.B) & fb) && fr == (white ? 0 : 7) && (ff == 1 || ff == 2 || ff == 5 || ff == 6)) return 6;
    if (tb & 0x0000001818000000ull) return 4;
    if (p.K & fb) {
        const int d = ff - (m.to & 7);
        if (d > 1 || d < -1) return 3;
        if (fr == (white ? 0 : 7)) return 3;
    }
    if (p.P & fb) return 2;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// One generator instance for both colours.  The kernels that LIST moves (tree_kernel.cu, expand_kernel) do not
// instantiate gen_side once per colour and visitor (six inlined copies, 220 KB of code, instruction-cache bound): black
// to move is turned into white to move by a vertical flip (byte swap of every bitboard, colours and castling rights
// swapped, ep square mirrored), the r-th move is picked in that view by ONE non-inlined function, and its squares are
// mirrored back.  Legal move sets are invariant under that symmetry; the ORDER in which moves are enumerated (which
// cdq_expand / cdq_sample_children / the forest expose as child indices) is the order of this function.
__device__ __forceinline__ u64 vflip(u64 b) {
    return ((u64)__byte_perm((unsigned)b, 0, 0x0123) << 32) | (u64)__byte_perm((unsigned)(b >> 32), 0, 0x0123);
}
__device__ __forceinline__ Pos white_view(const Pos& p) {      // identity when white is to move
    if (p.meta & CDQ_META_TURN_WHITE) return p;
    Pos f;
    f.P = vflip(p.P); f.N = vflip(p.N); f.B = vflip(p.B); f.R = vflip(p.R); f.Q = vflip(p.Q); f.K = vflip(p.K);
    f.co[1] = vflip(p.co[0]); f.co[0] = vflip(p.co[1]);
    const u64 m = p.meta;
    u64 meta = CDQ_META_TURN_WHITE | ((m & (CDQ_META_CASTLE_BK | CDQ_META_CASTLE_BQ)) >> 2) | ((m & (CDQ_META_CASTLE_WK | CDQ_META_CASTLE_WQ)) << 2);
    const u64 ep = (m >> CDQ_META_EP_SHIFT) & CDQ_META_EP_MASK;
    if (ep) meta |= (((ep - 1) ^ 56) + 1) << CDQ_META_EP_SHIFT;
    f.meta = meta | (m & CDQ_META_REP3);
    return f;
}

// The generator's output for a white view as a list of move CLASSES (target-square sets; classes are disjoint, each
// target is one move, x4 on the last rank): generation is one straight-line pass, picking the r-th move is a short loop
// over 27 sets, so a warp generates once and every lane picks its own moves from the same list.
enum { MC_SLIDE = 0, MC_KNIGHT = 8, MC_KING = 16, MC_PUSH = 17, MC_PUSH_PROMO = 18, MC_DOUBLE = 19, MC_CAP_W = 20, MC_CAP_W_PROMO = 21,
       MC_CAP_E = 22, MC_CAP_E_PROMO = 23, MC_CASTLE_K = 24, MC_CASTLE_Q = 25, MC_EP = 26, MC_COUNT = 27 };
struct MoveList {
    u64 t[MC_COUNT];
    u64 occ, king_bb, att_opp;   // white view
    int first[MC_COUNT + 1];     // first[k] = index of class k's first move in the enumeration; first[MC_COUNT] = n
    int ep_sq, n;                // n = number of legal moves
    bool ep_legal;               // Board.has_legal_en_passant()
};
struct ListVisitor {
    MoveList& L;
    __device__ ListVisitor(MoveList& l) : L(l) {}
    template <int D> __device__ __forceinline__ void slide(u64 t) { L.t[MC_SLIDE + D] = t; }
    template <int J> __device__ __forceinline__ void knight(u64 t) { L.t[MC_KNIGHT + J] = t; }
    __device__ __forceinline__ void king(u64 t) { L.t[MC_KING] = t; }
    __device__ __forceinline__ void pawn_push(u64 s, u64 d) { L.t[MC_PUSH] = s & ~RANK_8; L.t[MC_PUSH_PROMO] = s & RANK_8; L.t[MC_DOUBLE] = d; }
    template <int D> __device__ __forceinline__ void pawn_cap(u64 t) {      // white: D = 6 (north-west) and 4 (north-east)
        L.t[D == 6 ? MC_CAP_W : MC_CAP_E] = t & ~RANK_8;
        L.t[D == 6 ? MC_CAP_W_PROMO : MC_CAP_E_PROMO] = t & RANK_8;
    }
    __device__ __forceinline__ void castle(bool k, bool q) { L.t[MC_CASTLE_K] = k ? 1ull : 0ull; L.t[MC_CASTLE_Q] = q ? 1ull : 0ull; }
    __device__ __forceinline__ void ep(u64 caps, int epsq) { L.t[MC_EP] = caps; L.ep_sq = epsq; }
};
__device__ __forceinline__ int class_mult(int k) { return (k == MC_PUSH_PROMO || k == MC_CAP_W_PROMO || k == MC_CAP_E_PROMO) ? 4 : 1; }

// `w` must be a white view (white_view(p)).  The only place the list kernels expand gen_side: the classes' target sets.
static __device__ __noinline__ void list_classes(const Pos& w, MoveList& L) {
    L.occ = w.co[0] | w.co[1];
    L.king_bb = w.K & w.co[1];
    const Props q = make_props(~L.occ);
    L.att_opp = attack_map<0>(w, q);
    L.ep_sq = 0;
    ListVisitor v(L);
    gen_side<1>(w, q, L.att_opp, v);
}
// ... and the enumeration (first[], n), by one thread
__device__ __forceinline__ void list_moves(const Pos& w, MoveList& L) {
    list_classes(w, L);
    int n = 0;
#pragma unroll 1
    for (int k = 0; k < MC_COUNT; k++) { L.first[k] = n; n += popc(L.t[k]) * class_mult(k); }
    L.first[MC_COUNT] = n;
    L.n = n;
    L.ep_legal = L.t[MC_EP] != 0;
}
// ... or by a whole warp on ONE list in shared memory (every lane passes the same w): lane k counts class k, a shuffle scan
// gives first[] -- 27 dependent load / popc / store rounds otherwise, 7 % of the tree kernel's samples
__device__ __forceinline__ void list_moves_warp(const Pos& w, MoveList& L, int lane) {
    list_classes(w, L);
    __syncwarp();
    int c = lane < MC_COUNT ? popc(L.t[lane]) * class_mult(lane) : 0, incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane < MC_COUNT) L.first[lane] = incl - c;
    if (lane == MC_COUNT) { L.first[MC_COUNT] = incl; L.n = incl; L.ep_legal = L.t[MC_EP] != 0; }
    __syncwarp();
}
// the r-th move of the list (enumeration order = class order, targets ascending, promotions Q R B N), white-view coordinates
static __device__ __noinline__ PlainMove pick_from_list(const MoveList& L, int r) {
    PlainMove m;
    if (r < 0 || r >= L.n) return m;
    int k = 0;
#pragma unroll 1
    while (r >= L.first[k + 1]) k++;         // r < n = first[MC_COUNT]: terminates
    const int mult = class_mult(k), idx = r - L.first[k];
    const int sq = nth_bit(L.t[k], idx / mult);
    m.found = true;
    m.to = sq;
    if (mult == 4) m.promo = PT_Q - (idx % 4);
    if (k < MC_KNIGHT) {                                  // slider: origin = first piece behind the target
        const int d = k - MC_SLIDE;
        const int step = (d >> 1) == 0 ? 8 : (d >> 1) == 1 ? 1 : (d >> 1) == 2 ? 9 : 7;
        const int delta = (d & 1) ? -step : step;
        int f = sq - delta;
        while (!((L.occ >> f) & 1)) f -= delta;
        m.from = f;
    } else if (k < MC_KING) m.from = sq - knight_delta(k - MC_KNIGHT);
    else if (k == MC_KING) m.from = lsb_sq(L.king_bb);
    else if (k == MC_PUSH || k == MC_PUSH_PROMO) m.from = sq - 8;
    else if (k == MC_DOUBLE) { m.from = sq - 16; m.special = MV_DOUBLE; }
    else if (k == MC_CAP_W || k == MC_CAP_W_PROMO) m.from = sq - 7;
    else if (k == MC_CAP_E || k == MC_CAP_E_PROMO) m.from = sq - 9;
    else if (k == MC_CASTLE_K) { m.from = 4; m.to = 6; m.special = MV_CASTLE_K; }
    else if (k == MC_CASTLE_Q) { m.from = 4; m.to = 2; m.special = MV_CASTLE_Q; }
    else { m.from = sq; m.to = L.ep_sq; m.special = MV_EP; }
    return m;
}
// gives-check test of categorize_moves for a white-view child (black to move there): is black's king attacked by white
static __device__ __noinline__ bool white_view_child_in_check(const Pos& c) {
    const Props cq = make_props(~(c.co[0] | c.co[1]));
    return (attack_map<1>(c, cq) & c.K & c.co[0]) != 0;
}
// categorize_moves' weight of move m (white-view coordinates) in white view w, att_opp from pick_move
__device__ __forceinline__ int white_view_move_weight(const Pos& w, const PlainMove& m, u64 att_opp) {
    const u64 fb = 1ull << m.from, tb = 1ull << m.to;
    const u64 occ = w.co[0] | w.co[1];
    if (occ & tb) {
        const int vt = piece_value_at(w, tb), vf = piece_value_at(w, fb);
        return vt > vf ? 10 : (vt == vf ? 8 : 5);
    }
    Pos c = w;
    apply_move(c, m);
    if (white_view_child_in_check(c)) return 9;
    if (att_opp & fb) return 7;
    const int fr = m.from >> 3, ff = m.from & 7;
    if (((w.N | w.B) & fb) && fr == 0 && (ff == 1 || ff == 2 || ff == 5 || ff == 6)) return 6;
    if (tb & 0x0000001818000000ull) return 4;
    if (w.K & fb) {
        const int d = ff - (m.to & 7);
        if (d > 1 || d < -1) return 3;
        if (fr == 0) return 3;
    }
    if (w.P & fb) return 2;
    return 1;
}
// the move in the coordinates of the real position
__device__ __forceinline__ PlainMove real_move(const PlainMove& m, bool white_to_move) {
    PlainMove o;
    o.found = m.found; o.promo = m.promo; o.special = m.special;
    o.from = white_to_move ? m.from : (m.from ^ 56);
    o.to = white_to_move ? m.to : (m.to ^ 56);
    return o;
}

} // namespace cdq
