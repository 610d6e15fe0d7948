/*
 * cdq_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of (1) the python-chess behaviour that the reference's leaf
 * evaluation depends on and (2) evaluation.fast_evaluate_position itself, written as the
 * reference does it: per-square `is_attacked_by` loops and per-move legal generation with a
 * python-chess style `is_safe` filter.  The CUDA kernel uses a different (set-wise) method, so
 * agreement between the two is a real check.
 *
 * Third-party dependency holding the arithmetic: python-chess (PyPI `chess`), un-vendored and
 * unpinned in the reference (`chess>=1.9.0`, requirements.txt:4,11); absent from this image.
 * Its published algorithm is restated here from the behavioural spec in SURVEY.md 8c.
 *
 * Pinning: perft known answers (public constants, tests/test_oracle.py) pin the rules; the
 * reference's own evaluation.py run over oracle/pychess_shim (tests/golden/make_golden.py)
 * pins the evaluation restatement.  Real python-chess has never been available, so the
 * python-chess half stays "parity pinned to perft + shim", as DESIGN.md states.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * Reference citations are file:line in /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include "../include/cdq.h"

typedef uint64_t u64;

enum { PAWN = 0, KNIGHT, BISHOP, ROOK, QUEEN, KING, NTYPE };
enum { BLACK = 0, WHITE = 1 };

typedef struct {
    u64 pc[6];  /* by type */
    u64 co[2];  /* co[WHITE], co[BLACK] indexed by colour constant */
    u64 occ;
    u64 castling; /* bitboard of rook squares, like python-chess castling_rights */
    int ep;       /* raw ep square or -1 */
    int turn;     /* WHITE / BLACK */
    int rep3;
} Pos;

#define BB(s) (1ull << (s))
#define RANK_1 0xffull
#define RANK_8 0xff00000000000000ull
static const u64 BB_FILE_A = 0x0101010101010101ull;

static u64 KNIGHT_ATT[64], KING_ATT[64], PAWN_ATT[2][64];
static u64 BETWEEN[64][64], RAY[64][64];
static u64 RANK_MASK[64], FILE_MASK[64], DIAG_MASK[64];
static int tables_ready = 0;

static inline int popc(u64 b) { return __builtin_popcountll(b); }
static inline int lsb(u64 b) { return __builtin_ctzll(b); }
static inline int msb(u64 b) { return 63 - __builtin_clzll(b); }
static inline int sq_file(int s) { return s & 7; }
static inline int sq_rank(int s) { return s >> 3; }

static u64 step_attacks(int sq, const int (*d)[2], int nd) {
    u64 a = 0;
    for (int i = 0; i < nd; i++) {
        int f = sq_file(sq) + d[i][0], r = sq_rank(sq) + d[i][1];
        if (f >= 0 && f < 8 && r >= 0 && r < 8) a |= BB(r * 8 + f);
    }
    return a;
}

/* sliding attack along given directions with occupancy (stops at and includes first blocker) */
static u64 slide(int sq, u64 occ, const int (*d)[2], int nd) {
    u64 a = 0;
    for (int i = 0; i < nd; i++) {
        int f = sq_file(sq), r = sq_rank(sq);
        for (;;) {
            f += d[i][0]; r += d[i][1];
            if (f < 0 || f > 7 || r < 0 || r > 7) break;
            a |= BB(r * 8 + f);
            if (occ & BB(r * 8 + f)) break;
        }
    }
    return a;
}
static const int D_RANK[2][2] = {{1, 0}, {-1, 0}};
static const int D_FILE[2][2] = {{0, 1}, {0, -1}};
static const int D_DIAG[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
static const int D_KNIGHT[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
static const int D_KING[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};

static inline u64 rank_att(int sq, u64 occ) { return slide(sq, occ, D_RANK, 2); }
static inline u64 file_att(int sq, u64 occ) { return slide(sq, occ, D_FILE, 2); }
static inline u64 diag_att(int sq, u64 occ) { return slide(sq, occ, D_DIAG, 4); }

void cdq_oracle_init(void) {
    if (tables_ready) return;
    for (int s = 0; s < 64; s++) {
        KNIGHT_ATT[s] = step_attacks(s, D_KNIGHT, 8);
        KING_ATT[s] = step_attacks(s, D_KING, 8);
        const int wp[2][2] = {{-1, 1}, {1, 1}}, bp[2][2] = {{-1, -1}, {1, -1}};
        PAWN_ATT[WHITE][s] = step_attacks(s, wp, 2);
        PAWN_ATT[BLACK][s] = step_attacks(s, bp, 2);
        RANK_MASK[s] = rank_att(s, 0);
        FILE_MASK[s] = file_att(s, 0);
        DIAG_MASK[s] = diag_att(s, 0);
    }
    for (int a = 0; a < 64; a++)
        for (int b = 0; b < 64; b++) {
            u64 line = 0, btw = 0;
            if (a != b) {
                if (RANK_MASK[a] & BB(b)) {
                    line = RANK_MASK[a] | BB(a);
                    btw = rank_att(a, BB(b)) & rank_att(b, BB(a));
                } else if (FILE_MASK[a] & BB(b)) {
                    line = FILE_MASK[a] | BB(a);
                    btw = file_att(a, BB(b)) & file_att(b, BB(a));
                } else if (DIAG_MASK[a] & BB(b)) {
                    /* only the diagonal that contains both */
                    line = (diag_att(a, 0) & diag_att(b, 0)) | BB(a) | BB(b);
                    btw = diag_att(a, BB(b)) & diag_att(b, BB(a));
                }
            }
            RAY[a][b] = line;      /* python-chess chess.ray(a, b): whole line incl. both, else 0 */
            BETWEEN[a][b] = btw;   /* chess.between(a, b): strictly between */
        }
    tables_ready = 1;
}

/* ---------------------------------------------------------------- wire format */
static void unpack(const CdqBoard* in, Pos* p) {
    for (int t = 0; t < 6; t++) p->pc[t] = in->bb[t];
    p->co[WHITE] = in->bb[6];
    p->co[BLACK] = in->bb[7];
    p->occ = p->co[WHITE] | p->co[BLACK];
    u64 m = in->meta;
    p->turn = (m & CDQ_META_TURN_WHITE) ? WHITE : BLACK;
    p->castling = 0;
    if (m & CDQ_META_CASTLE_WK) p->castling |= BB(7);
    if (m & CDQ_META_CASTLE_WQ) p->castling |= BB(0);
    if (m & CDQ_META_CASTLE_BK) p->castling |= BB(63);
    if (m & CDQ_META_CASTLE_BQ) p->castling |= BB(56);
    int e = (int)((m >> CDQ_META_EP_SHIFT) & CDQ_META_EP_MASK);
    p->ep = e ? e - 1 : -1;
    p->rep3 = (m & CDQ_META_REP3) ? 1 : 0;
}

/* board.clean_castling_rights() for standard chess: the right's rook on its corner, own king on
 * e1/e8 */
static u64 clean_castling(const Pos* p) {
    u64 c = p->castling & p->pc[ROOK];
    u64 w = c & RANK_1 & p->co[WHITE] & (BB(0) | BB(7));
    u64 b = c & RANK_8 & p->co[BLACK] & (BB(56) | BB(63));
    if (!(p->co[WHITE] & p->pc[KING] & BB(4))) w = 0;
    if (!(p->co[BLACK] & p->pc[KING] & BB(60))) b = 0;
    return w | b;
}

static void pack(const Pos* p, CdqBoard* out) {
    for (int t = 0; t < 6; t++) out->bb[t] = p->pc[t];
    out->bb[6] = p->co[WHITE];
    out->bb[7] = p->co[BLACK];
    u64 c = clean_castling(p);
    u64 m = p->turn == WHITE ? CDQ_META_TURN_WHITE : 0;
    if (c & BB(7)) m |= CDQ_META_CASTLE_WK;
    if (c & BB(0)) m |= CDQ_META_CASTLE_WQ;
    if (c & BB(63)) m |= CDQ_META_CASTLE_BK;
    if (c & BB(56)) m |= CDQ_META_CASTLE_BQ;
    if (p->ep >= 0) m |= (u64)(p->ep + 1) << CDQ_META_EP_SHIFT;
    if (p->rep3) m |= CDQ_META_REP3;
    out->meta = m;
}

static void set_startpos(Pos* p) {
    memset(p, 0, sizeof *p);
    p->pc[PAWN] = 0x00ff00000000ff00ull;
    p->pc[KNIGHT] = 0x4200000000000042ull;
    p->pc[BISHOP] = 0x2400000000000024ull;
    p->pc[ROOK] = 0x8100000000000081ull;
    p->pc[QUEEN] = 0x0800000000000008ull;
    p->pc[KING] = 0x1000000000000010ull;
    p->co[WHITE] = 0xffffull;
    p->co[BLACK] = 0xffff000000000000ull;
    p->occ = p->co[WHITE] | p->co[BLACK];
    p->castling = BB(0) | BB(7) | BB(56) | BB(63);
    p->ep = -1;
    p->turn = WHITE;
}

/* FEN (first four fields; counters ignored).  Returns 0 on success. */
static int parse_fen(const char* fen, Pos* p) {
    memset(p, 0, sizeof *p);
    p->ep = -1;
    int r = 7, f = 0;
    const char* c = fen;
    for (; *c && *c != ' '; c++) {
        if (*c == '/') { r--; f = 0; continue; }
        if (isdigit((unsigned char)*c)) { f += *c - '0'; continue; }
        int col = isupper((unsigned char)*c) ? WHITE : BLACK, t;
        switch (tolower((unsigned char)*c)) {
            case 'p': t = PAWN; break;   case 'n': t = KNIGHT; break;
            case 'b': t = BISHOP; break; case 'r': t = ROOK; break;
            case 'q': t = QUEEN; break;  case 'k': t = KING; break;
            default: return -1;
        }
        if (r < 0 || f > 7) return -1;
        p->pc[t] |= BB(r * 8 + f);
        p->co[col] |= BB(r * 8 + f);
        f++;
    }
    p->occ = p->co[WHITE] | p->co[BLACK];
    while (*c == ' ') c++;
    p->turn = (*c == 'b') ? BLACK : WHITE;
    if (*c) c++;
    while (*c == ' ') c++;
    for (; *c && *c != ' '; c++) {
        if (*c == 'K') p->castling |= BB(7);
        else if (*c == 'Q') p->castling |= BB(0);
        else if (*c == 'k') p->castling |= BB(63);
        else if (*c == 'q') p->castling |= BB(56);
    }
    while (*c == ' ') c++;
    if (*c && *c != '-' && c[1]) p->ep = (c[1] - '1') * 8 + (c[0] - 'a');
    p->castling = clean_castling(p);
    return 0;
}

/* ---------------------------------------------------------------- attacks (python-chess) */
/* Board.attackers_mask(color, square, occupied) */
static u64 attackers(const Pos* p, int color, int sq, u64 occ) {
    u64 qr = p->pc[QUEEN] | p->pc[ROOK], qb = p->pc[QUEEN] | p->pc[BISHOP];
    u64 a = (KING_ATT[sq] & p->pc[KING]) | (KNIGHT_ATT[sq] & p->pc[KNIGHT]) |
            (rank_att(sq, occ) & qr) | (file_att(sq, occ) & qr) | (diag_att(sq, occ) & qb) |
            (PAWN_ATT[!color][sq] & p->pc[PAWN]);
    return a & p->co[color];
}
static inline int is_attacked_by(const Pos* p, int color, int sq) {
    return attackers(p, color, sq, p->occ) != 0;
}
/* Board.king(color): msb of the colour's kings or -1 */
static inline int king_sq(const Pos* p, int color) {
    u64 k = p->co[color] & p->pc[KING];
    return k ? msb(k) : -1;
}
static int type_at(const Pos* p, int sq) {
    for (int t = 0; t < 6; t++)
        if (p->pc[t] & BB(sq)) return t;
    return -1;
}
/* Board.attacks_mask(square) */
static u64 attacks_from(const Pos* p, int sq) {
    u64 b = BB(sq);
    if (p->pc[PAWN] & b) return PAWN_ATT[(p->co[WHITE] & b) ? WHITE : BLACK][sq];
    if (p->pc[KNIGHT] & b) return KNIGHT_ATT[sq];
    if (p->pc[KING] & b) return KING_ATT[sq];
    u64 a = 0;
    if ((p->pc[BISHOP] | p->pc[QUEEN]) & b) a |= diag_att(sq, p->occ);
    if ((p->pc[ROOK] | p->pc[QUEEN]) & b) a |= rank_att(sq, p->occ) | file_att(sq, p->occ);
    return a;
}

/* ---------------------------------------------------------------- move generation */
typedef uint16_t Move; /* from | to << 6 | promo_type << 12 (0 = none) */
#define MV(f, t, pr) ((Move)((f) | ((t) << 6) | ((pr) << 12)))
#define MV_FROM(m) ((m) & 63)
#define MV_TO(m) (((m) >> 6) & 63)
#define MV_PROMO(m) ((m) >> 12)
#define MAXMOVES 512

typedef struct { Move m[MAXMOVES]; int n; } MoveList;
static inline void add(MoveList* l, int f, int t, int pr) { l->m[l->n++] = MV(f, t, pr); }
static void add_pawn(MoveList* l, int f, int t) {
    if (sq_rank(t) == 0 || sq_rank(t) == 7) {
        add(l, f, t, QUEEN); add(l, f, t, ROOK); add(l, f, t, BISHOP); add(l, f, t, KNIGHT);
    } else add(l, f, t, 0);
}

/* Board.generate_pseudo_legal_ep */
static void gen_ep(const Pos* p, u64 from_mask, u64 to_mask, MoveList* l) {
    if (p->ep <= 0 || !(BB(p->ep) & to_mask)) return; /* `if not self.ep_square` : 0 is falsy too */
    if (BB(p->ep) & p->occ) return;
    u64 rank = p->turn == WHITE ? (0xffull << 32) : (0xffull << 24);
    u64 cap = p->pc[PAWN] & p->co[p->turn] & from_mask & PAWN_ATT[!p->turn][p->ep] & rank;
    while (cap) { int s = msb(cap); cap ^= BB(s); add(l, s, p->ep, 0); }
}

/* any square of `path` attacked by the opponent under `occ` (Board._attacked_for_king) */
static int attacked_for_king(const Pos* p, u64 path, u64 occ) {
    while (path) { int s = lsb(path); path &= path - 1; if (attackers(p, !p->turn, s, occ)) return 1; }
    return 0;
}

/* Board.generate_castling_moves (standard chess) */
static void gen_castling(const Pos* p, u64 from_mask, u64 to_mask, MoveList* l) {
    u64 backrank = p->turn == WHITE ? RANK_1 : RANK_8;
    u64 king = p->co[p->turn] & p->pc[KING] & backrank & from_mask;
    king &= -king;
    if (!king) return;
    int ks = lsb(king);
    u64 cand = clean_castling(p) & backrank & to_mask;
    while (cand) {
        int rs = msb(cand); cand ^= BB(rs);
        u64 rook = BB(rs);
        int a_side = rook < king;
        int base = p->turn == WHITE ? 0 : 56;
        int kto = base + (a_side ? 2 : 6), rto = base + (a_side ? 3 : 5);
        u64 king_path = BETWEEN[ks][kto], rook_path = BETWEEN[rs][rto];
        if ((p->occ ^ king ^ rook) & (king_path | rook_path | BB(kto) | BB(rto))) continue;
        if (attacked_for_king(p, king_path | king, p->occ ^ king)) continue;
        if (attacked_for_king(p, BB(kto), p->occ ^ king ^ rook ^ BB(rto))) continue;
        add(l, ks, kto, 0);
    }
}

/* Board.generate_pseudo_legal_moves */
static void gen_pseudo(const Pos* p, u64 from_mask, u64 to_mask, MoveList* l) {
    u64 ours = p->co[p->turn];
    u64 non_pawns = ours & ~p->pc[PAWN] & from_mask;
    while (non_pawns) {
        int f = msb(non_pawns); non_pawns ^= BB(f);
        u64 mv = attacks_from(p, f) & ~ours & to_mask;
        while (mv) { int t = msb(mv); mv ^= BB(t); add(l, f, t, 0); }
    }
    if (from_mask & p->pc[KING]) gen_castling(p, from_mask, to_mask, l);
    u64 pawns = p->pc[PAWN] & ours & from_mask;
    if (!pawns) return;
    u64 cap = pawns;
    while (cap) {
        int f = msb(cap); cap ^= BB(f);
        u64 tg = PAWN_ATT[p->turn][f] & p->co[!p->turn] & to_mask;
        while (tg) { int t = msb(tg); tg ^= BB(t); add_pawn(l, f, t); }
    }
    u64 single, dbl;
    if (p->turn == WHITE) {
        single = (pawns << 8) & ~p->occ;
        dbl = (single << 8) & ~p->occ & (0xffull << 24);
    } else {
        single = (pawns >> 8) & ~p->occ;
        dbl = (single >> 8) & ~p->occ & (0xffull << 32);
    }
    single &= to_mask; dbl &= to_mask;
    while (single) {
        int t = msb(single); single ^= BB(t);
        add_pawn(l, t + (p->turn == BLACK ? 8 : -8), t);
    }
    while (dbl) {
        int t = msb(dbl); dbl ^= BB(t);
        add(l, t + (p->turn == BLACK ? 16 : -16), t, 0);
    }
    if (p->ep > 0) gen_ep(p, from_mask, to_mask, l);
}

/* Board._slider_blockers(king) */
static u64 slider_blockers(const Pos* p, int king) {
    u64 rq = p->pc[ROOK] | p->pc[QUEEN], bq = p->pc[BISHOP] | p->pc[QUEEN];
    u64 snipers = ((RANK_MASK[king] | FILE_MASK[king]) & rq) | (DIAG_MASK[king] & bq);
    snipers &= p->co[!p->turn];
    u64 blockers = 0;
    while (snipers) {
        int s = msb(snipers); snipers ^= BB(s);
        u64 b = BETWEEN[king][s] & p->occ;
        if (b && (b & (b - 1)) == 0) blockers |= b;
    }
    return blockers & p->co[p->turn];
}

static int is_ep_move(const Pos* p, Move m) {
    int f = MV_FROM(m), t = MV_TO(m), d = t > f ? t - f : f - t;
    return p->ep == t && (p->pc[PAWN] & BB(f)) && (d == 7 || d == 9) && !(p->occ & BB(t));
}
static int is_castling_move(const Pos* p, Move m) {
    int f = MV_FROM(m), t = MV_TO(m);
    if (!(p->pc[KING] & BB(f))) return 0;
    int d = sq_file(f) - sq_file(t);
    return d > 1 || d < -1 || (p->pc[ROOK] & p->co[p->turn] & BB(t)) != 0;
}
/* Board.pin_mask(color, square) */
static u64 pin_mask(const Pos* p, int color, int sq) {
    int king = king_sq(p, color);
    if (king < 0) return ~0ull;
    u64 sm = BB(sq);
    u64 rays[3] = {FILE_MASK[king], RANK_MASK[king], DIAG_MASK[king]};
    u64 sl[3] = {p->pc[ROOK] | p->pc[QUEEN], p->pc[ROOK] | p->pc[QUEEN], p->pc[BISHOP] | p->pc[QUEEN]};
    for (int i = 0; i < 3; i++) {
        if (rays[i] & sm) {
            u64 sn = rays[i] & sl[i] & p->co[!color];
            while (sn) {
                int s = msb(sn); sn ^= BB(s);
                if ((BETWEEN[s][king] & (p->occ | sm)) == sm) return RAY[king][s];
            }
            break;
        }
    }
    return ~0ull;
}
/* Board._ep_skewered(king, capturer) */
static int ep_skewered(const Pos* p, int king, int capturer) {
    int last_double = p->ep + (p->turn == WHITE ? -8 : 8);
    u64 occ = ((p->occ & ~BB(last_double) & ~BB(capturer)) | BB(p->ep));
    u64 h = p->co[!p->turn] & (p->pc[ROOK] | p->pc[QUEEN]);
    if (rank_att(king, occ) & h) return 1;
    u64 d = p->co[!p->turn] & (p->pc[BISHOP] | p->pc[QUEEN]);
    if (diag_att(king, occ) & d) return 1;
    return 0;
}
/* Board._is_safe(king, blockers, move) */
static int is_safe(const Pos* p, int king, u64 blockers, Move m) {
    int f = MV_FROM(m), t = MV_TO(m);
    if (f == king) {
        if (is_castling_move(p, m)) return 1;
        return !is_attacked_by(p, !p->turn, t);
    } else if (is_ep_move(p, m)) {
        return (pin_mask(p, p->turn, f) & BB(t)) && !ep_skewered(p, king, f);
    } else {
        return !(blockers & BB(f)) || (RAY[f][t] & BB(king));
    }
}
/* Board._generate_evasions(king, checkers) */
static void gen_evasions(const Pos* p, int king, u64 checkers, MoveList* l) {
    u64 sliders = checkers & (p->pc[BISHOP] | p->pc[ROOK] | p->pc[QUEEN]);
    u64 attacked = 0;
    u64 s = sliders;
    while (s) { int c = msb(s); s ^= BB(c); attacked |= RAY[king][c] & ~BB(c); }
    u64 kt = KING_ATT[king] & ~p->co[p->turn] & ~attacked;
    while (kt) { int t = msb(kt); kt ^= BB(t); add(l, king, t, 0); }
    int checker = msb(checkers);
    if (BB(checker) == checkers) {
        u64 target = BETWEEN[king][checker] | checkers;
        gen_pseudo(p, ~p->pc[KING], target, l);
        if (p->ep > 0 && !(BB(p->ep) & target)) {
            int last_double = p->ep + (p->turn == WHITE ? -8 : 8);
            if (last_double == checker) gen_ep(p, ~0ull, ~0ull, l);
        }
    }
}
/* Board.generate_legal_moves */
static void gen_legal(const Pos* p, MoveList* out) {
    out->n = 0;
    u64 king_mask = p->pc[KING] & p->co[p->turn];
    if (!king_mask) { gen_pseudo(p, ~0ull, ~0ull, out); return; }
    int king = msb(king_mask);
    u64 blockers = slider_blockers(p, king);
    u64 checkers = attackers(p, !p->turn, king, p->occ);
    MoveList tmp; tmp.n = 0;
    if (checkers) gen_evasions(p, king, checkers, &tmp);
    else gen_pseudo(p, ~0ull, ~0ull, &tmp);
    for (int i = 0; i < tmp.n; i++)
        if (is_safe(p, king, blockers, tmp.m[i])) out->m[out->n++] = tmp.m[i];
}

/* Board.push (standard chess) */
static void push(Pos* p, Move m) {
    int f = MV_FROM(m), t = MV_TO(m), pr = MV_PROMO(m);
    int us = p->turn, them = !us;
    int ep = p->ep;
    p->ep = -1;
    int castling = is_castling_move(p, m);
    int pt = type_at(p, f);
    int captured = type_at(p, t);
    p->castling &= ~BB(t) & ~BB(f);
    if (pt == KING) p->castling &= ~(us == WHITE ? RANK_1 : RANK_8);
    /* remove mover */
    p->pc[pt] &= ~BB(f); p->co[us] &= ~BB(f);
    if (pt == PAWN) {
        int d = t - f;
        if (d == 16 && sq_rank(f) == 1) p->ep = f + 8;
        else if (d == -16 && sq_rank(f) == 6) p->ep = f - 8;
        else if (t == ep && (d == 7 || d == 9 || d == -7 || d == -9) && captured < 0) {
            int cs = ep + (us == WHITE ? -8 : 8);
            p->pc[PAWN] &= ~BB(cs); p->co[them] &= ~BB(cs);
        }
    }
    if (pr) pt = pr;
    if (castling) {
        int base = us == WHITE ? 0 : 56;
        int a_side = sq_file(t) < sq_file(f);
        int rf = base + (a_side ? 0 : 7), rt = base + (a_side ? 3 : 5), kt = base + (a_side ? 2 : 6);
        p->pc[ROOK] &= ~BB(rf); p->co[us] &= ~BB(rf);
        p->pc[KING] |= BB(kt); p->co[us] |= BB(kt);
        p->pc[ROOK] |= BB(rt); p->co[us] |= BB(rt);
    } else {
        if (captured >= 0) { p->pc[captured] &= ~BB(t); p->co[them] &= ~BB(t); }
        p->pc[pt] |= BB(t); p->co[us] |= BB(t);
    }
    p->occ = p->co[WHITE] | p->co[BLACK];
    p->turn = them;
}

/* Board.has_insufficient_material(color) / is_insufficient_material */
static int side_insufficient(const Pos* p, int color) {
    if (p->co[color] & (p->pc[PAWN] | p->pc[ROOK] | p->pc[QUEEN])) return 0;
    if (p->co[color] & p->pc[KNIGHT])
        return popc(p->co[color]) <= 2 && !(p->co[!color] & ~p->pc[KING] & ~p->pc[QUEEN]);
    if (p->co[color] & p->pc[BISHOP]) {
        const u64 DARK = 0xAA55AA55AA55AA55ull, LIGHT = 0x55AA55AA55AA55AAull;
        int same = !(p->pc[BISHOP] & DARK) || !(p->pc[BISHOP] & LIGHT);
        return same && !p->pc[PAWN] && !p->pc[KNIGHT];
    }
    return 1;
}
static int insufficient(const Pos* p) { return side_insufficient(p, WHITE) && side_insufficient(p, BLACK); }
static int in_check(const Pos* p) {
    int k = king_sq(p, p->turn);
    return k >= 0 && attackers(p, !p->turn, k, p->occ) != 0;
}

/* ---------------------------------------------------------------- evaluation.py restated */
static const int PIECE_VALUE[6] = {1, 3, 3, 5, 9, 0}; /* evaluation.py:8-15 */
#define SQ(f, r) ((r) * 8 + (f))

static int count_in(const int* list, int n, u64 set) {
    int c = 0;
    for (int i = 0; i < n; i++) c += (set >> list[i]) & 1;
    return c;
}

/* Computes all integer terms, the flags word and the fp64 score of one position.
 * Mirrors evaluation.py:46-221 statement by statement (cache omitted: it only memoises). */
static void eval_one(const Pos* p, int ignore_checkmate, int32_t* T, double* score, uint32_t* flags) {
    MoveList ml;
    /* --- terminal tests (evaluation.py:50-57) */
    gen_legal(p, &ml);
    int stm_moves = ml.n;
    int chk = in_check(p);
    int checkmate = chk && stm_moves == 0;
    int stalemate = !chk && stm_moves == 0;
    int insuff = insufficient(p);
    int rep3 = p->rep3;
    int term = checkmate ? CDQ_TERM_CHECKMATE : stalemate ? CDQ_TERM_STALEMATE
             : insuff ? CDQ_TERM_INSUFFICIENT : rep3 ? CDQ_TERM_REP3 : CDQ_TERM_NONE;

    /* --- 1. material (:67-71) */
    int wm = 0, bm = 0;
    for (int t = 0; t < 6; t++) {
        wm += popc(p->pc[t] & p->co[WHITE]) * PIECE_VALUE[t];
        bm += popc(p->pc[t] & p->co[BLACK]) * PIECE_VALUE[t];
    }
    /* --- 2. mobility: legal moves with the turn forced (:78-84) */
    Pos q = *p;
    q.turn = WHITE; gen_legal(&q, &ml); int wmob = ml.n;
    q.turn = BLACK; gen_legal(&q, &ml); int bmob = ml.n;
    /* attacked squares (:90-94) */
    u64 watt = 0, batt = 0;
    for (int s = 0; s < 64; s++) {
        if (is_attacked_by(p, WHITE, s)) watt |= BB(s);
        if (is_attacked_by(p, BLACK, s)) batt |= BB(s);
    }
    /* --- 3. king safety (:100-139).  `if white_king_square:` is false for None AND square 0. */
    int wk = king_sq(p, WHITE), bk = king_sq(p, BLACK);
    int wka = 0, bka = 0;
    if (wk > 0)
        for (int s = 0; s < 64; s++)
            if ((p->co[BLACK] & BB(s)) && is_attacked_by(p, BLACK, wk)) wka++;
    if (bk > 0)
        for (int s = 0; s < 64; s++)
            if ((p->co[WHITE] & BB(s)) && is_attacked_by(p, WHITE, bk)) bka++;
    int w_castled = (wk == SQ(6, 0) || wk == SQ(2, 0));
    int b_castled = (bk == SQ(6, 7) || bk == SQ(2, 7));
    int w_centre = (wk == SQ(3, 3) || wk == SQ(4, 3) || wk == SQ(3, 4) || wk == SQ(4, 4));
    int b_centre = (bk == SQ(3, 3) || bk == SQ(4, 3) || bk == SQ(3, 4) || bk == SQ(4, 4));
    /* --- 4. pawn structure (:142-176) */
    int dbl[2] = {0, 0}, iso[2] = {0, 0};
    for (int c = 0; c < 2; c++) {
        u64 pw = p->pc[PAWN] & p->co[c];
        int per_file[8];
        for (int f = 0; f < 8; f++) per_file[f] = popc(pw & (BB_FILE_A << f));
        for (int f = 0; f < 8; f++) if (per_file[f]) dbl[c] += per_file[f] - 1;
        for (int f = 0; f < 8; f++) {
            int nb = (f > 0 && per_file[f - 1]) || (f < 7 && per_file[f + 1]);
            if (!nb) iso[c] += per_file[f];
        }
    }
    /* --- 5. space (:181-190) */
    static const int CENTRE[4] = {SQ(3, 3), SQ(3, 4), SQ(4, 3), SQ(4, 4)};
    static const int EXT[12] = {SQ(2, 2), SQ(2, 3), SQ(2, 4), SQ(2, 5), SQ(3, 2), SQ(3, 5),
                                SQ(4, 2), SQ(4, 5), SQ(5, 2), SQ(5, 3), SQ(5, 4), SQ(5, 5)};
    int wc = count_in(CENTRE, 4, watt), bc = count_in(CENTRE, 4, batt);
    int we = count_in(EXT, 12, watt), be = count_in(EXT, 12, batt);
    /* --- 6. coordination (:196-202) */
    int wd = 0, bd = 0;
    for (int s = 0; s < 64; s++) {
        if (!(p->occ & BB(s))) continue;
        if ((p->co[WHITE] & BB(s)) && is_attacked_by(p, WHITE, s)) wd++;
        else if ((p->co[BLACK] & BB(s)) && is_attacked_by(p, BLACK, s)) bd++;
    }
    /* --- check bonus (:207-209) */
    int check_sign = chk ? (p->turn == BLACK ? 1 : -1) : 0;

    T[CDQ_T_W_MAT] = wm;   T[CDQ_T_B_MAT] = bm;
    T[CDQ_T_W_MOB] = wmob; T[CDQ_T_B_MOB] = bmob;
    T[CDQ_T_W_ATT] = popc(watt); T[CDQ_T_B_ATT] = popc(batt);
    T[CDQ_T_W_KATT] = wka; T[CDQ_T_B_KATT] = bka;
    T[CDQ_T_W_CASTLED] = w_castled; T[CDQ_T_B_CASTLED] = b_castled;
    T[CDQ_T_W_KCENTRE] = w_centre;  T[CDQ_T_B_KCENTRE] = b_centre;
    T[CDQ_T_W_DBL] = dbl[WHITE]; T[CDQ_T_B_DBL] = dbl[BLACK];
    T[CDQ_T_W_ISO] = iso[WHITE]; T[CDQ_T_B_ISO] = iso[BLACK];
    T[CDQ_T_W_CENTRE] = wc; T[CDQ_T_B_CENTRE] = bc;
    T[CDQ_T_W_EXT] = we;    T[CDQ_T_B_EXT] = be;
    T[CDQ_T_W_DEF] = wd;    T[CDQ_T_B_DEF] = bd;
    T[CDQ_T_CHECK] = check_sign;
    T[CDQ_T_TERMINAL] = term;

    *flags = (uint32_t)term | (chk ? CDQ_F_CHECK : 0) | (checkmate ? CDQ_F_CHECKMATE : 0) |
             (stalemate ? CDQ_F_STALEMATE : 0) | (insuff ? CDQ_F_INSUFFICIENT : 0) |
             (rep3 ? CDQ_F_REP3 : 0) | ((uint32_t)(stm_moves > 255 ? 255 : stm_moves) << CDQ_F_STM_MOVES_SHIFT);

    /* --- score: terminal short-circuit (:50-64) then the weighted sum in Python's
     * left-to-right fp64 order (:96-97,123,132-139,178,190,204,212-221). */
    if (checkmate && !ignore_checkmate) { *score = p->turn == WHITE ? -50.0 : 50.0; return; }
    if (stalemate || insuff || rep3) { *score = 0.0; return; }
    volatile double material = (double)(wm - bm);
    volatile double mobility = (double)(wmob - bmob) * 0.1;
    volatile double control = (double)(popc(watt) - popc(batt)) * 0.05;
    volatile double ks = (double)(bka - wka) * 0.5;
    if (w_castled) ks += 1;
    if (b_castled) ks -= 1;
    if (w_centre) ks -= 2;
    if (b_centre) ks += 2;
    volatile double a = (double)(dbl[BLACK] - dbl[WHITE]) * 0.3;
    volatile double b = (double)(iso[BLACK] - iso[WHITE]) * 0.2;
    volatile double pawn = a + b;
    volatile double c1 = (double)(wc - bc) * 0.5;
    volatile double c2 = (double)(we - be) * 0.1;
    volatile double space = c1 + c2;
    volatile double coord = (double)(wd - bd) * 0.1;
    volatile double e = material * 1.0;
    volatile double t1 = mobility * 0.3; e = e + t1;
    volatile double t2 = control * 0.3;  e = e + t2;
    volatile double t3 = ks * 0.5;       e = e + t3;
    volatile double t4 = pawn * 0.4;     e = e + t4;
    volatile double t5 = space * 0.3;    e = e + t5;
    volatile double t6 = coord * 0.4;    e = e + t6;
    e = e + (double)(check_sign * 50);
    *score = e;
}

/* ---------------------------------------------------------------- exported API (ctypes) */
void cdq_oracle_startpos(CdqBoard* out) { Pos p; cdq_oracle_init(); set_startpos(&p); pack(&p, out); }

int cdq_oracle_from_fen(const char* fen, CdqBoard* out) {
    Pos p; cdq_oracle_init();
    if (parse_fen(fen, &p)) return -1;
    pack(&p, out);
    return 0;
}

static u64 perft_rec(const Pos* p, int depth) {
    MoveList l; gen_legal(p, &l);
    if (depth == 1) return (u64)l.n;
    u64 n = 0;
    for (int i = 0; i < l.n; i++) { Pos q = *p; push(&q, l.m[i]); n += perft_rec(&q, depth - 1); }
    return n;
}
uint64_t cdq_oracle_perft(const CdqBoard* b, int depth) {
    Pos p; cdq_oracle_init(); unpack(b, &p);
    return depth <= 0 ? 1 : perft_rec(&p, depth);
}

/* legal moves of the side to move (or of `force_turn` if 0/1; -1 keeps the record's turn) */
int cdq_oracle_legal_moves(const CdqBoard* b, int force_turn, uint16_t* moves /*[512] or NULL*/) {
    Pos p; cdq_oracle_init(); unpack(b, &p);
    if (force_turn == 0 || force_turn == 1) p.turn = force_turn;
    MoveList l; gen_legal(&p, &l);
    if (moves) memcpy(moves, l.m, sizeof(Move) * l.n);
    return l.n;
}

void cdq_oracle_push(CdqBoard* b, uint16_t move) {
    Pos p; cdq_oracle_init(); unpack(b, &p); push(&p, move); pack(&p, b);
}

/* simple pthread fan-out over [0, n) in contiguous slices */
typedef struct { void (*fn)(int64_t, int64_t, void*); int64_t lo, hi; void* arg; } Slice;
static void* slice_main(void* v) { Slice* s = (Slice*)v; s->fn(s->lo, s->hi, s->arg); return NULL; }
static void parallel_for(int64_t n, int nthreads, void (*fn)(int64_t, int64_t, void*), void* arg) {
    if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads > 256) nthreads = 256;
    if (nthreads <= 1 || n < 2 * nthreads) { fn(0, n, arg); return; }
    pthread_t th[256]; Slice sl[256];
    for (int t = 0; t < nthreads; t++) {
        sl[t].fn = fn; sl[t].arg = arg;
        sl[t].lo = n * t / nthreads; sl[t].hi = n * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, slice_main, &sl[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

typedef struct { const CdqBoard* boards; int ignore; int32_t* terms; double* score; uint32_t* flags; } EvalArgs;
static void eval_slice(int64_t lo, int64_t hi, void* v) {
    EvalArgs* a = (EvalArgs*)v;
    for (int64_t i = lo; i < hi; i++) {
        Pos p; unpack(&a->boards[i], &p);
        int32_t T[CDQ_NTERMS]; double s; uint32_t f;
        eval_one(&p, a->ignore, T, &s, &f);
        if (a->terms) memcpy(a->terms + i * CDQ_NTERMS, T, sizeof T);
        if (a->score) a->score[i] = s;
        if (a->flags) a->flags[i] = f;
    }
}
/* nthreads <= 0: all online cores */
void cdq_oracle_eval(const CdqBoard* boards, int64_t n, int ignore_checkmate, int32_t* terms,
                     double* score, uint32_t* flags, int nthreads) {
    cdq_oracle_init();
    EvalArgs a = {boards, ignore_checkmate, terms, score, flags};
    parallel_for(n, nthreads, eval_slice, &a);
}

/* categorize_moves (evaluation.py:232-339): sampling weight of every legal move, in the order of
 * cdq_oracle_legal_moves. */
int cdq_oracle_categorize(const CdqBoard* b, uint16_t* moves, int* weights) {
    Pos p; cdq_oracle_init(); unpack(b, &p);
    MoveList l; gen_legal(&p, &l);
    for (int i = 0; i < l.n; i++) {
        Move m = l.m[i];
        int f = MV_FROM(m), t = MV_TO(m), w;
        int ft = type_at(&p, f), tt = type_at(&p, t);
        int fr = sq_rank(f), ff = sq_file(f), tf = sq_file(t);
        if (tt >= 0) {
            int vt = PIECE_VALUE[tt], vf = PIECE_VALUE[ft];
            w = vt > vf ? 10 : (vt == vf ? 8 : 5);
        } else {
            Pos q = p; push(&q, m);
            if (in_check(&q)) w = 9;
            else if (is_attacked_by(&p, !p.turn, f)) w = 7;
            else if ((ft == KNIGHT || ft == BISHOP) && fr == (p.turn == WHITE ? 0 : 7) && (ff == 1 || ff == 2 || ff == 5 || ff == 6)) w = 6;
            else if (t == SQ(3, 3) || t == SQ(4, 3) || t == SQ(3, 4) || t == SQ(4, 4)) w = 4;
            else if (ft == KING && (ff - tf > 1 || tf - ff > 1 || fr == (p.turn == WHITE ? 0 : 7))) w = 3;
            else if (ft == PAWN) w = 2;
            else w = 1;
        }
        if (moves) moves[i] = m;
        weights[i] = w;
    }
    return l.n;
}

/* evaluation.py:368-428: per position {attacked by white, attacked by black, legal destination
 * squares of white, of black (turn forced)} */
void cdq_oracle_maps(const CdqBoard* boards, int64_t n, uint64_t* out) {
    cdq_oracle_init();
    for (int64_t i = 0; i < n; i++) {
        Pos p; unpack(&boards[i], &p);
        u64 a[2] = {0, 0}, d[2] = {0, 0};
        for (int s = 0; s < 64; s++) {
            if (is_attacked_by(&p, WHITE, s)) a[0] |= BB(s);
            if (is_attacked_by(&p, BLACK, s)) a[1] |= BB(s);
        }
        for (int c = 0; c < 2; c++) {
            Pos q = p; q.turn = c == 0 ? WHITE : BLACK;
            MoveList l; gen_legal(&q, &l);
            for (int k = 0; k < l.n; k++) d[c] |= BB(MV_TO(l.m[k]));
        }
        out[i * 4 + 0] = a[0]; out[i * 4 + 1] = a[1]; out[i * 4 + 2] = d[0]; out[i * 4 + 3] = d[1];
    }
}

/* board_to_tensor (board_utils.py:12-40) for a batch: planes [n][12][8][8] */
void cdq_oracle_encode_planes(const CdqBoard* boards, int64_t n, float* planes) {
    for (int64_t i = 0; i < n; i++) {
        float* o = planes + i * 768;
        memset(o, 0, 768 * sizeof(float));
        for (int s = 0; s < 64; s++)
            for (int t = 0; t < 6; t++)
                if (boards[i].bb[t] & BB(s)) {
                    if (boards[i].bb[6] & BB(s)) o[t * 64 + s] = 1.0f;
                    else if (boards[i].bb[7] & BB(s)) o[(6 + t) * 64 + s] = 1.0f;
                }
    }
}

/* splitmix64 counter RNG */
static inline u64 mix64(u64 z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* Synthetic random-playout positions: position i = start position after
 * U{0..max_plies} uniformly random legal plies; stops early at a terminal position. */
typedef struct { uint64_t seed; int max_plies; CdqBoard* out; } PlayArgs;
static void playout_slice(int64_t lo, int64_t hi, void* v) {
    PlayArgs* a = (PlayArgs*)v;
    for (int64_t i = lo; i < hi; i++) {
        Pos p; set_startpos(&p);
        u64 key = mix64(a->seed ^ mix64((u64)i));
        int plies = (int)(mix64(key) % (u64)(a->max_plies + 1));
        for (int k = 0; k < plies; k++) {
            MoveList l; gen_legal(&p, &l);
            if (l.n == 0 || insufficient(&p)) break;
            push(&p, l.m[mix64(key + 1 + (u64)k) % (u64)l.n]);
        }
        pack(&p, &a->out[i]);
    }
}
void cdq_oracle_playouts(uint64_t seed, int64_t n, int max_plies, CdqBoard* out) {
    cdq_oracle_init();
    PlayArgs a = {seed, max_plies, out};
    parallel_for(n, 0, playout_slice, &a);
}
