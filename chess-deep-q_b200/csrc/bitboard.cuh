// bitboard.cuh -- device-side chess bitboard machinery shared by the evaluation kernel (K1) and
// the synthetic playout generator.  One THREAD owns one position and everything is computed
// set-wise on 64-bit boards (Kogge-Stone occluded fills per direction, no per-move loops, no
// tables), so a warp runs 32 positions in lock-step without divergence.
//
// What it must reproduce is python-chess behaviour as used by the reference
// (evaluation.py:78-94,208; mcts.py:202-206): `is_attacked_by` maps, and the NUMBER of moves
// `generate_legal_moves` yields for a side whose turn is forced, including its quirks
// (captures of the enemy king by the side that is not really to move, x4 promotions, the raw
// ep square, castling through `clean_castling_rights`).  SURVEY.md 8a/8c hold the spec.
#pragma once
#include <cstdint>
#include "../../include/cdq.h"

namespace cdq {

typedef unsigned long long u64;

constexpr u64 FILE_A = 0x0101010101010101ull;
constexpr u64 FILE_B = FILE_A << 1;
constexpr u64 FILE_G = FILE_A << 6;
constexpr u64 FILE_H = FILE_A << 7;
constexpr u64 RANK_1 = 0xffull;
constexpr u64 RANK_8 = 0xffull << 56;
constexpr u64 ALL = ~0ull;

// directions: 0 N(+8) 1 S(-8) 2 E(+1) 3 W(-1) 4 NE(+9) 5 SW(-9) 6 NW(+7) 7 SE(-7)
// orientation(d) = d >> 1: 0 file, 1 rank, 2 diagonal, 3 anti-diagonal
template <int D> struct Dir;
template <> struct Dir<0> { static constexpr int s = 8; static constexpr bool left = true;  static constexpr u64 wrap = ALL; };
template <> struct Dir<1> { static constexpr int s = 8; static constexpr bool left = false; static constexpr u64 wrap = ALL; };
template <> struct Dir<2> { static constexpr int s = 1; static constexpr bool left = true;  static constexpr u64 wrap = ~FILE_A; };
template <> struct Dir<3> { static constexpr int s = 1; static constexpr bool left = false; static constexpr u64 wrap = ~FILE_H; };
template <> struct Dir<4> { static constexpr int s = 9; static constexpr bool left = true;  static constexpr u64 wrap = ~FILE_A; };
template <> struct Dir<5> { static constexpr int s = 9; static constexpr bool left = false; static constexpr u64 wrap = ~FILE_H; };
template <> struct Dir<6> { static constexpr int s = 7; static constexpr bool left = true;  static constexpr u64 wrap = ~FILE_H; };
template <> struct Dir<7> { static constexpr int s = 7; static constexpr bool left = false; static constexpr u64 wrap = ~FILE_A; };

template <int D, int K> __device__ __forceinline__ u64 raw_shift(u64 x) {
    return Dir<D>::left ? (x << (Dir<D>::s * K)) : (x >> (Dir<D>::s * K));
}
// one step in direction D, squares that fall off the board are dropped
template <int D> __device__ __forceinline__ u64 step(u64 x) { return raw_shift<D, 1>(x) & Dir<D>::wrap; }

// Kogge-Stone propagators of one direction for a fixed set of empty squares; shared by every
// fill in that direction (attack maps, king scans, move sets of both colours).
struct Prop { u64 p1, p2, p4; };
template <int D> __device__ __forceinline__ Prop make_prop(u64 empty) {
    Prop p;
    p.p1 = empty & Dir<D>::wrap;
    p.p2 = p.p1 & raw_shift<D, 1>(p.p1);
    p.p4 = p.p2 & raw_shift<D, 2>(p.p2);
    return p;
}
// squares attacked in direction D by the sliders in `g`: every empty square up to and including
// the first occupied one.  Rays of different sliders in one direction never overlap.
template <int D> __device__ __forceinline__ u64 ray(u64 g, const Prop& p) {
    g |= p.p1 & raw_shift<D, 1>(g);
    g |= p.p2 & raw_shift<D, 2>(g);
    g |= p.p4 & raw_shift<D, 4>(g);
    return step<D>(g);
}

__device__ __forceinline__ u64 knight_attacks(u64 n) {
    return ((n << 17) & ~FILE_A) | ((n << 15) & ~FILE_H) | ((n << 10) & ~(FILE_A | FILE_B)) |
           ((n << 6) & ~(FILE_G | FILE_H)) | ((n >> 17) & ~FILE_H) | ((n >> 15) & ~FILE_A) |
           ((n >> 10) & ~(FILE_G | FILE_H)) | ((n >> 6) & ~(FILE_A | FILE_B));
}
template <int J> __device__ __forceinline__ u64 knight_step(u64 n) {
    switch (J) {
        case 0: return (n << 17) & ~FILE_A;
        case 1: return (n << 15) & ~FILE_H;
        case 2: return (n << 10) & ~(FILE_A | FILE_B);
        case 3: return (n << 6) & ~(FILE_G | FILE_H);
        case 4: return (n >> 17) & ~FILE_H;
        case 5: return (n >> 15) & ~FILE_A;
        case 6: return (n >> 10) & ~(FILE_G | FILE_H);
        default: return (n >> 6) & ~(FILE_A | FILE_B);
    }
}
__device__ __forceinline__ int knight_delta(int j) {
    const int d[8] = {17, 15, 10, 6, -17, -15, -10, -6};
    return d[j];
}
__device__ __forceinline__ u64 king_attacks(u64 k) {
    u64 h = ((k << 1) & ~FILE_A) | ((k >> 1) & ~FILE_H);
    u64 row = h | k;
    return h | (row << 8) | (row >> 8);
}
__device__ __forceinline__ int popc(u64 x) { return __popcll(x); }
__device__ __forceinline__ int msb_sq(u64 x) { return 63 - __clzll((long long)x); }
__device__ __forceinline__ int lsb_sq(u64 x) { return __ffsll((long long)x) - 1; }

// ---------------------------------------------------------------------------------------------
struct Pos {
    u64 P, N, B, R, Q, K; // by type
    u64 co[2];            // co[1] = white, co[0] = black (colour constants as python-chess bools)
    u64 meta;             // cdq.h layout
};

__device__ __forceinline__ Pos load_pos(const u64* rec) {
    Pos p;
    p.P = rec[0]; p.N = rec[1]; p.B = rec[2]; p.R = rec[3]; p.Q = rec[4]; p.K = rec[5];
    p.co[1] = rec[6]; p.co[0] = rec[7]; p.meta = rec[8];
    return p;
}
__device__ __forceinline__ void store_pos(const Pos& p, u64* rec) {
    rec[0] = p.P; rec[1] = p.N; rec[2] = p.B; rec[3] = p.R; rec[4] = p.Q; rec[5] = p.K;
    rec[6] = p.co[1]; rec[7] = p.co[0]; rec[8] = p.meta;
}

struct Props { Prop d0, d1, d2, d3, d4, d5, d6, d7; };
__device__ __forceinline__ Props make_props(u64 empty) {
    Props q;
    q.d0 = make_prop<0>(empty); q.d1 = make_prop<1>(empty); q.d2 = make_prop<2>(empty);
    q.d3 = make_prop<3>(empty); q.d4 = make_prop<4>(empty); q.d5 = make_prop<5>(empty);
    q.d6 = make_prop<6>(empty); q.d7 = make_prop<7>(empty);
    return q;
}
template <int D> __device__ __forceinline__ const Prop& prop_of(const Props& q) {
    if (D == 0) return q.d0; if (D == 1) return q.d1; if (D == 2) return q.d2; if (D == 3) return q.d3;
    if (D == 4) return q.d4; if (D == 5) return q.d5; if (D == 6) return q.d6; return q.d7;
}

// Board.is_attacked_by(colour, .) for all 64 squares at once (evaluation.py:90-94).
template <int C> __device__ __forceinline__ u64 attack_map(const Pos& p, const Props& q) {
    const u64 own = p.co[C];
    const u64 orth = (p.R | p.Q) & own, diag = (p.B | p.Q) & own;
    u64 a = knight_attacks(p.N & own) | king_attacks(p.K & own);
    const u64 pw = p.P & own;
    a |= C ? (step<6>(pw) | step<4>(pw)) : (step<5>(pw) | step<7>(pw));
    a |= ray<0>(orth, q.d0) | ray<1>(orth, q.d1) | ray<2>(orth, q.d2) | ray<3>(orth, q.d3);
    a |= ray<4>(diag, q.d4) | ray<5>(diag, q.d5) | ray<6>(diag, q.d6) | ray<7>(diag, q.d7);
    return a;
}

// slow single-square ray walk, only used in the rare en-passant tail
__device__ __forceinline__ u64 walk(int sq, int df, int dr, u64 occ) {
    u64 a = 0;
    int f = sq & 7, r = sq >> 3;
    for (int i = 0; i < 7; i++) {
        f += df; r += dr;
        if (f < 0 || f > 7 || r < 0 || r > 7) break;
        u64 b = 1ull << (r * 8 + f);
        a |= b;
        if (occ & b) break;
    }
    return a;
}

// Per-side scan state that the visitor-based generator fills in.
struct SideInfo {
    u64 checkers;   // Board.attackers_mask(not side, king)
    int n_moves;    // filled by CountVisitor users
};

// Enumerates the legal moves of side C (turn forced to C, as evaluation.py:78-84 does) as SETS:
// for every move class it hands the visitor a bitboard of target squares; each target is one
// move (x4 on the last rank for pawns).  The classes are disjoint by construction, so the sum
// of popcounts equals len(list(board.legal_moves)).
//   v.slide<D>(targets)    slider moves in direction D (origin = first piece behind the target)
//   v.knight<J>(targets)   knight moves with offset J
//   v.king(targets)        king steps
//   v.pawn_push(single, dbl), v.pawn_cap<D>(targets)   (D = 6/4 for white, 5/7 for black)
//   v.castle(kingside, queenside)
//   v.ep(capturers, ep_square)
// att_opp = attack_map<!C>.  Returns the checkers of C's king.
template <int C, class V>
__device__ __forceinline__ u64 gen_side(const Pos& p, const Props& q, u64 att_opp, V& v) {
    const u64 own = p.co[C], opp = p.co[!C], occ = own | opp, empty = ~occ;
    const u64 orthS = p.R | p.Q, diagS = p.B | p.Q;
    const u64 king = p.K & own;

    // --- king scans: first and second piece met in each direction from the king
    u64 pin0 = 0, pin1 = 0, pin2 = 0, pin3 = 0; // own pieces pinned along orientation 0..3
    u64 checkers = 0, chk_target = 0, excl = 0;
#define CDQ_SCAN(D, KIND, PIN)                                                      \
    {                                                                               \
        const u64 a1 = ray<D>(king, prop_of<D>(q));                                 \
        const u64 b1 = a1 & occ;                                                    \
        const u64 c = b1 & opp & (KIND);                                            \
        checkers |= c;                                                              \
        if (c) { chk_target |= a1; excl |= (step<D>(king) & ~c) | step<(D ^ 1)>(king); } \
        const u64 pb = b1 & own;                                                    \
        const u64 a2 = ray<D>(pb, prop_of<D>(q));                                   \
        if (a2 & occ & opp & (KIND)) PIN |= pb;                                     \
    }
    CDQ_SCAN(0, orthS, pin0) CDQ_SCAN(1, orthS, pin0)
    CDQ_SCAN(2, orthS, pin1) CDQ_SCAN(3, orthS, pin1)
    CDQ_SCAN(4, diagS, pin2) CDQ_SCAN(5, diagS, pin2)
    CDQ_SCAN(6, diagS, pin3) CDQ_SCAN(7, diagS, pin3)
#undef CDQ_SCAN
    {
        const u64 pawn_chk = (C ? (step<6>(king) | step<4>(king)) : (step<5>(king) | step<7>(king))) & p.P;
        const u64 c = ((knight_attacks(king) & p.N) | pawn_chk | (king_attacks(king) & p.K)) & opp;
        checkers |= c;
        chk_target |= c;
    }
    const int n_chk = popc(checkers);
    const u64 T = n_chk == 0 ? ALL : (n_chk == 1 ? chk_target : 0ull);
    const u64 pinned = pin0 | pin1 | pin2 | pin3;
    const u64 land = ~own & T;

    // --- sliders, per direction; a pinned slider only moves along its pin line
#define CDQ_SLIDE(D, KIND, PIN) \
    v.template slide<D>(ray<D>((KIND) & own & ~(pinned & ~(PIN)), prop_of<D>(q)) & land);
    CDQ_SLIDE(0, orthS, pin0) CDQ_SLIDE(1, orthS, pin0) CDQ_SLIDE(2, orthS, pin1) CDQ_SLIDE(3, orthS, pin1)
    CDQ_SLIDE(4, diagS, pin2) CDQ_SLIDE(5, diagS, pin2) CDQ_SLIDE(6, diagS, pin3) CDQ_SLIDE(7, diagS, pin3)
#undef CDQ_SLIDE
    // --- knights (a pinned knight never moves)
    {
        const u64 n = p.N & own & ~pinned;
        v.template knight<0>(knight_step<0>(n) & land); v.template knight<1>(knight_step<1>(n) & land);
        v.template knight<2>(knight_step<2>(n) & land); v.template knight<3>(knight_step<3>(n) & land);
        v.template knight<4>(knight_step<4>(n) & land); v.template knight<5>(knight_step<5>(n) & land);
        v.template knight<6>(knight_step<6>(n) & land); v.template knight<7>(knight_step<7>(n) & land);
    }
    // --- king steps: not onto own pieces, not onto squares attacked under the CURRENT occupancy,
    //     not along the line of a checking slider (Board._generate_evasions / _is_safe)
    v.king(king_attacks(king) & ~own & ~att_opp & ~excl);
    // --- pawns
    {
        const u64 pw = p.P & own;
        const u64 push_ok = pw & ~(pinned & ~pin0);
        if (C) {
            const u64 single = (push_ok << 8) & empty;
            v.pawn_push(single & T, (single << 8) & empty & (0xffull << 24) & T);
            v.template pawn_cap<6>(step<6>(pw & ~(pinned & ~pin3)) & opp & T);
            v.template pawn_cap<4>(step<4>(pw & ~(pinned & ~pin2)) & opp & T);
        } else {
            const u64 single = (push_ok >> 8) & empty;
            v.pawn_push(single & T, (single >> 8) & empty & (0xffull << 32) & T);
            v.template pawn_cap<5>(step<5>(pw & ~(pinned & ~pin2)) & opp & T);
            v.template pawn_cap<7>(step<7>(pw & ~(pinned & ~pin3)) & opp & T);
        }
    }
    // --- castling (Board.generate_castling_moves with clean rights; never while in check)
    {
        const u64 rights = p.meta >> (C ? 1 : 3);
        const int sh = C ? 0 : 56;
        const u64 e = 1ull << (4 + sh);
        const bool ok = n_chk == 0 && (king & e);
        // king side: f,g empty; e,f,g not attacked.  queen side: b,c,d empty; e,d,c not attacked.
        const u64 rk = p.R & own;
        const bool ks = ok && (rights & 1) && (rk & (0x80ull << sh)) && !(occ & (0x60ull << sh)) &&
                        !(att_opp & (0x70ull << sh));
        const bool qs = ok && (rights & 2) && (rk & (0x01ull << sh)) && !(occ & (0x0eull << sh)) &&
                        !(att_opp & (0x1cull << sh));
        v.castle(ks, qs);
    }
    // --- en passant (rare tail): Board.generate_pseudo_legal_ep + _is_safe(ep) + evasion rule
    {
        const int epf = (int)((p.meta >> CDQ_META_EP_SHIFT) & CDQ_META_EP_MASK);
        if (epf > 1) { // raw ep_square truthy (square 0 is falsy in the library's test)
            const int ep = epf - 1;
            const u64 epbb = 1ull << ep;
            const u64 from_sq = C ? (step<5>(epbb) | step<7>(epbb)) : (step<6>(epbb) | step<4>(epbb));
            u64 caps = p.P & own & from_sq & (C ? (0xffull << 32) : (0xffull << 24));
            const u64 last_double = C ? (epbb >> 8) : (epbb << 8);
            bool allowed = !(epbb & occ);
            if (n_chk == 1) allowed = allowed && ((epbb & T) || last_double == checkers);
            else if (n_chk > 1) allowed = false;
            u64 legal = 0;
            if (allowed && caps) {
                if (!king) legal = caps;
                else {
                    const int ksq = msb_sq(king);
                    const int kf = ksq & 7, kr = ksq >> 3, ef = ep & 7, er = ep >> 3;
                    while (caps) {
                        const u64 c = caps & (0 - caps);
                        caps ^= c;
                        bool ok = true;
                        // pin_mask(capturer) must contain the ep square
                        if (c & pin0) ok = (ef == kf);
                        else if (c & pin1) ok = (er == kr);
                        else if (c & pin2) ok = (er - ef == kr - kf);
                        else if (c & pin3) ok = (er + ef == kr + kf);
                        // _ep_skewered: both pawns leave, capturer lands on ep
                        const u64 o2 = (occ & ~last_double & ~c) | epbb;
                        const u64 h = walk(ksq, 1, 0, o2) | walk(ksq, -1, 0, o2);
                        if (h & opp & orthS) ok = false;
                        const u64 d = walk(ksq, 1, 1, o2) | walk(ksq, 1, -1, o2) | walk(ksq, -1, 1, o2) |
                                      walk(ksq, -1, -1, o2);
                        if (d & opp & diagS) ok = false;
                        if (ok) legal |= c;
                    }
                }
            }
            v.ep(legal, ep);
        } else {
            v.ep(0ull, 0);
        }
    }
    return checkers;
}

// Visitor that only counts (K1).
struct CountVisitor {
    int n = 0;
    template <int D> __device__ __forceinline__ void slide(u64 t) { n += popc(t); }
    template <int J> __device__ __forceinline__ void knight(u64 t) { n += popc(t); }
    __device__ __forceinline__ void king(u64 t) { n += popc(t); }
    __device__ __forceinline__ void pawn_push(u64 s, u64 d) {
        n += popc(s) + 3 * popc(s & (RANK_1 | RANK_8)) + popc(d);
    }
    template <int D> __device__ __forceinline__ void pawn_cap(u64 t) { n += popc(t) + 3 * popc(t & (RANK_1 | RANK_8)); }
    __device__ __forceinline__ void castle(bool k, bool q) { n += (int)k + (int)q; }
    __device__ __forceinline__ void ep(u64 caps, int) { n += popc(caps); }
};

// Board.is_insufficient_material()
__device__ __forceinline__ bool side_insufficient(const Pos& p, int c) {
    const u64 own = p.co[c], opp = p.co[!c];
    if (own & (p.P | p.R | p.Q)) return false;
    if (own & p.N) return popc(own) <= 2 && !(opp & ~p.K & ~p.Q);
    if (own & p.B) {
        const u64 DARK = 0xAA55AA55AA55AA55ull;
        const bool same = !(p.B & DARK) || !(p.B & ~DARK);
        return same && !p.P && !p.N;
    }
    return true;
}
__device__ __forceinline__ bool insufficient_material(const Pos& p) {
    return side_insufficient(p, 1) && side_insufficient(p, 0);
}

} // namespace cdq
