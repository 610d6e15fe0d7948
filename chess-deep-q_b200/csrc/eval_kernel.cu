// eval_kernel.cu -- K1: batched handcrafted evaluation (evaluation.py:30-229) and plane encoding
// (board_utils.py:12-40) on packed boards.
//
// Layout: one thread per position, 128 positions per CTA.  A CTA stages its 128 x 72 B records
// through shared memory with fully coalesced 16-byte loads (9 216 B contiguous), each thread then
// reads its own nine u64 (stride 72 B = 18 words: conflict-free for 64-bit accesses).  All terms
// are computed set-wise in registers (bitboard.cuh); the fp64 score is combined with explicit
// __dmul_rn/__dadd_rn in the reference's left-to-right order so no FMA contraction can change a
// bit.  Algorithmic traffic: 72 B in + 8 B score + 4 B flags = 84 B/position (+96 B when the
// integer terms are dumped for parity checks).
#include "bitboard.cuh"
#include "common.cuh"

namespace cdq {

constexpr int EVAL_THREADS = 128;

struct EvalOut {
    int terms[CDQ_NTERMS];
    double score;
    unsigned flags;
};

__device__ __forceinline__ void pawn_structure(u64 pw, int& dbl, int& iso) {
    u64 f = pw | (pw >> 32);
    f |= f >> 16;
    f |= f >> 8;
    f &= 0xff; // occupied files
    dbl = popc(pw) - popc(f);
    const u64 nb = ((f << 1) | (f >> 1)) & 0xff;
    const u64 lonely = f & ~nb;
    iso = popc(pw & (lonely * FILE_A));
}

__device__ __forceinline__ void evaluate_position(const Pos& p, int ignore_checkmate, EvalOut& o) {
    const u64 W = p.co[1], Bk = p.co[0], occ = W | Bk;
    const Props q = make_props(~occ);
    const u64 att_w = attack_map<1>(p, q);
    const u64 att_b = attack_map<0>(p, q);

    CountVisitor cw, cb;
    const u64 chk_w = gen_side<1>(p, q, att_b, cw); // white to move (forced): evaluation.py:78-80
    const u64 chk_b = gen_side<0>(p, q, att_w, cb); // black to move (forced): evaluation.py:82-84

    const bool white_to_move = p.meta & CDQ_META_TURN_WHITE;
    const bool in_check = (white_to_move ? chk_w : chk_b) != 0;
    const int stm_moves = white_to_move ? cw.n : cb.n;
    const bool checkmate = in_check && stm_moves == 0;
    const bool stalemate = !in_check && stm_moves == 0;
    const bool insuff = insufficient_material(p);
    const bool rep3 = (p.meta & CDQ_META_REP3) != 0;

    // material (evaluation.py:67-70)
    const u64 minor = p.N | p.B;
    const int wm = popc(p.P & W) + 3 * popc(minor & W) + 5 * popc(p.R & W) + 9 * popc(p.Q & W);
    const int bm = popc(p.P & Bk) + 3 * popc(minor & Bk) + 5 * popc(p.R & Bk) + 9 * popc(p.Q & Bk);

    // king safety (evaluation.py:100-139); `if white_king_square:` is false for a king on a1
    const u64 wk = p.K & W, bk = p.K & Bk;
    const int wka = (wk & ~1ull & att_b) ? popc(Bk) : 0;
    const int bka = (bk & ~1ull & att_w) ? popc(W) : 0;
    const u64 CENTRE = 0x0000001818000000ull;
    const u64 EXT = 0x00003c24243c0000ull;
    const int w_castled = (wk & 0x44ull) != 0, b_castled = (bk & (0x44ull << 56)) != 0;
    const int w_kc = (wk & CENTRE) != 0, b_kc = (bk & CENTRE) != 0;

    int wd, wi, bd, bi;
    pawn_structure(p.P & W, wd, wi);
    pawn_structure(p.P & Bk, bd, bi);

    const int watt = popc(att_w), batt = popc(att_b);
    const int wc = popc(att_w & CENTRE), bc = popc(att_b & CENTRE);
    const int we = popc(att_w & EXT), be = popc(att_b & EXT);
    const int wdef = popc(att_w & W), bdef = popc(att_b & Bk);
    const int check_sign = in_check ? (white_to_move ? -1 : 1) : 0;

    const int term = checkmate ? CDQ_TERM_CHECKMATE : stalemate ? CDQ_TERM_STALEMATE
                   : insuff ? CDQ_TERM_INSUFFICIENT : rep3 ? CDQ_TERM_REP3 : CDQ_TERM_NONE;
    o.terms[CDQ_T_W_MAT] = wm;   o.terms[CDQ_T_B_MAT] = bm;
    o.terms[CDQ_T_W_MOB] = cw.n; o.terms[CDQ_T_B_MOB] = cb.n;
    o.terms[CDQ_T_W_ATT] = watt; o.terms[CDQ_T_B_ATT] = batt;
    o.terms[CDQ_T_W_KATT] = wka; o.terms[CDQ_T_B_KATT] = bka;
    o.terms[CDQ_T_W_CASTLED] = w_castled; o.terms[CDQ_T_B_CASTLED] = b_castled;
    o.terms[CDQ_T_W_KCENTRE] = w_kc; o.terms[CDQ_T_B_KCENTRE] = b_kc;
    o.terms[CDQ_T_W_DBL] = wd; o.terms[CDQ_T_B_DBL] = bd;
    o.terms[CDQ_T_W_ISO] = wi; o.terms[CDQ_T_B_ISO] = bi;
    o.terms[CDQ_T_W_CENTRE] = wc; o.terms[CDQ_T_B_CENTRE] = bc;
    o.terms[CDQ_T_W_EXT] = we; o.terms[CDQ_T_B_EXT] = be;
    o.terms[CDQ_T_W_DEF] = wdef; o.terms[CDQ_T_B_DEF] = bdef;
    o.terms[CDQ_T_CHECK] = check_sign;
    o.terms[CDQ_T_TERMINAL] = term;

    unsigned fl = (unsigned)term | (in_check ? CDQ_F_CHECK : 0u) | (checkmate ? CDQ_F_CHECKMATE : 0u) |
                  (stalemate ? CDQ_F_STALEMATE : 0u) | (insuff ? CDQ_F_INSUFFICIENT : 0u) |
                  (rep3 ? CDQ_F_REP3 : 0u) | ((unsigned)min(stm_moves, 255) << CDQ_F_STM_MOVES_SHIFT) |
                  (white_to_move ? CDQ_F_WHITE_TO_MOVE : 0u);
    // format violations: more than one king of a colour, overlapping colours, piece sets that do
    // not partition the occupancy
    const u64 types_or = p.P | p.N | p.B | p.R | p.Q | p.K;
    const int types_cnt = popc(p.P) + popc(p.N) + popc(p.B) + popc(p.R) + popc(p.Q) + popc(p.K);
    if (popc(wk) > 1 || popc(bk) > 1 || (W & Bk) || types_or != occ || types_cnt != popc(occ))
        fl |= CDQ_F_BADBOARD;
    o.flags = fl;

    // score (evaluation.py:50-64, 96-97, 123-139, 178, 190, 204, 212-221), strict IEEE fp64
    double s;
    if (checkmate && !ignore_checkmate) s = white_to_move ? -50.0 : 50.0;
    else if (stalemate || insuff || rep3) s = 0.0;
    else {
        const double material = (double)(wm - bm);
        const double mobility = __dmul_rn((double)(cw.n - cb.n), 0.1);
        const double control = __dmul_rn((double)(watt - batt), 0.05);
        double ks = __dmul_rn((double)(bka - wka), 0.5);
        if (w_castled) ks = __dadd_rn(ks, 1.0);
        if (b_castled) ks = __dadd_rn(ks, -1.0);
        if (w_kc) ks = __dadd_rn(ks, -2.0);
        if (b_kc) ks = __dadd_rn(ks, 2.0);
        const double pawn = __dadd_rn(__dmul_rn((double)(bd - wd), 0.3), __dmul_rn((double)(bi - wi), 0.2));
        const double space = __dadd_rn(__dmul_rn((double)(wc - bc), 0.5), __dmul_rn((double)(we - be), 0.1));
        const double coord = __dmul_rn((double)(wdef - bdef), 0.1);
        s = __dmul_rn(material, 1.0);
        s = __dadd_rn(s, __dmul_rn(mobility, 0.3));
        s = __dadd_rn(s, __dmul_rn(control, 0.3));
        s = __dadd_rn(s, __dmul_rn(ks, 0.5));
        s = __dadd_rn(s, __dmul_rn(pawn, 0.4));
        s = __dadd_rn(s, __dmul_rn(space, 0.3));
        s = __dadd_rn(s, __dmul_rn(coord, 0.4));
        s = __dadd_rn(s, (double)(check_sign * 50));
    }
    o.score = s;
}

// Cooperative load of one CTA's records into shared memory.
template <int THREADS>
__device__ __forceinline__ void stage_records(const CdqBoard* __restrict__ boards, long long n, long long base,
                                              u64* smem /*[THREADS*9]*/) {
    const long long remaining = n - base;
    const int count = remaining < THREADS ? (int)remaining : THREADS; // records in this CTA
    const int words = count * 9;                                      // u64 words
    const u64* g = reinterpret_cast<const u64*>(boards) + base * 9;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const ulonglong2* g2 = reinterpret_cast<const ulonglong2*>(g);
        ulonglong2* s2 = reinterpret_cast<ulonglong2*>(smem);
        const int pairs = words >> 1;
        for (int i = threadIdx.x; i < pairs; i += THREADS) s2[i] = __ldg(g2 + i);
        if ((words & 1) && threadIdx.x == 0) smem[words - 1] = __ldg(g + words - 1);
    } else {
        for (int i = threadIdx.x; i < words; i += THREADS) smem[i] = __ldg(g + i);
    }
    __syncthreads();
}

// HEURISTIC_LEAF instantiates the no-network leaf rule as a second kernel so that the hot instantiation
// (the one cdq_leaf_eval launches in front of the CNN) keeps its register count and code footprint.
template <bool HEURISTIC_LEAF>
__global__ void __launch_bounds__(EVAL_THREADS)
eval_terms_kernel(const CdqBoard* __restrict__ boards, long long n, int ignore_checkmate,
                  int* __restrict__ terms, double* __restrict__ score, unsigned* __restrict__ flags,
                  float* __restrict__ heuristic_leaf) {
    __shared__ __align__(16) u64 srec[EVAL_THREADS * 9];
    const long long base = (long long)blockIdx.x * EVAL_THREADS;
    stage_records<EVAL_THREADS>(boards, n, base, srec);
    const long long i = base + threadIdx.x;
    if (i >= n) return;
    const Pos p = load_pos(srec + threadIdx.x * 9);
    EvalOut o;
    evaluate_position(p, ignore_checkmate, o);
    if (score) score[i] = o.score;
    if (flags) flags[i] = o.flags;
    if (HEURISTIC_LEAF) {
        // _evaluate_position without a network (mcts.py:202-210, 226-229): the terminal rules, else
        // tanh(fast_evaluate_position(board) / 10) -- fp64 like math.tanh, rounded to the fp32 leaf array
        const unsigned term = o.flags & CDQ_F_TERM_MASK;
        float v;
        if (term == CDQ_TERM_CHECKMATE) v = (o.flags & CDQ_F_WHITE_TO_MOVE) ? -1.f : 1.f;
        else if (term != CDQ_TERM_NONE) v = 0.f;
        else v = (float)tanh(o.score / 10.0);
        heuristic_leaf[i] = v;
    }
    if (terms) {
        int4* t4 = reinterpret_cast<int4*>(terms + i * CDQ_NTERMS);
#pragma unroll
        for (int k = 0; k < CDQ_NTERMS / 4; k++)
            t4[k] = make_int4(o.terms[4 * k], o.terms[4 * k + 1], o.terms[4 * k + 2], o.terms[4 * k + 3]);
    }
}

// board_to_tensor for a batch: planes[n][12][8][8] fp32.  One thread per (position, plane): it
// expands one 64-bit plane into 64 floats with sixteen float4 stores (256 contiguous bytes).
__global__ void __launch_bounds__(256)
encode_planes_kernel(const CdqBoard* __restrict__ boards, long long n, float* __restrict__ planes) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 12) return;
    const long long i = t / 12;
    const int ch = (int)(t - i * 12);
    const u64* rec = reinterpret_cast<const u64*>(boards) + i * 9;
    const u64 bits = __ldg(rec + (ch % 6)) & __ldg(rec + (ch < 6 ? 6 : 7));
    float4* out = reinterpret_cast<float4*>(planes + t * 64);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const unsigned nib = (unsigned)(bits >> (4 * k)) & 15u;
        out[k] = make_float4((nib & 1) ? 1.f : 0.f, (nib & 2) ? 1.f : 0.f, (nib & 4) ? 1.f : 0.f,
                             (nib & 8) ? 1.f : 0.f);
    }
}

// the maps behind find_threatened_squares / find_guarded_squares /
// calculate_attack_range / calculate_movement_range (evaluation.py:368-428): per position
// [attacked by white, attacked by black, legal destinations of white, legal destinations of black]
// (turn forced, castling = the king's destination, en passant = the ep square).
struct DestVisitor {
    u64 d = 0;
    bool white;
    template <int D> __device__ __forceinline__ void slide(u64 t) { d |= t; }
    template <int J> __device__ __forceinline__ void knight(u64 t) { d |= t; }
    __device__ __forceinline__ void king(u64 t) { d |= t; }
    __device__ __forceinline__ void pawn_push(u64 s, u64 dd) { d |= s | dd; }
    template <int D> __device__ __forceinline__ void pawn_cap(u64 t) { d |= t; }
    __device__ __forceinline__ void castle(bool k, bool q) {
        const int sh = white ? 0 : 56;
        if (k) d |= 0x40ull << sh;
        if (q) d |= 0x04ull << sh;
    }
    __device__ __forceinline__ void ep(u64 caps, int epsq) { if (caps) d |= 1ull << epsq; }
};

__global__ void __launch_bounds__(EVAL_THREADS)
attack_maps_kernel(const CdqBoard* __restrict__ boards, long long n, u64* __restrict__ maps) {
    __shared__ __align__(16) u64 srec[EVAL_THREADS * 9];
    const long long base = (long long)blockIdx.x * EVAL_THREADS;
    stage_records<EVAL_THREADS>(boards, n, base, srec);
    const long long i = base + threadIdx.x;
    if (i >= n) return;
    const Pos p = load_pos(srec + threadIdx.x * 9);
    const Props q = make_props(~(p.co[0] | p.co[1]));
    const u64 att_w = attack_map<1>(p, q), att_b = attack_map<0>(p, q);
    DestVisitor vw, vb;
    vw.white = true; vb.white = false;
    gen_side<1>(p, q, att_b, vw);
    gen_side<0>(p, q, att_w, vb);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(maps + i * 4);
    o[0] = make_ulonglong2(att_w, att_b);
    o[1] = make_ulonglong2(vw.d, vb.d);
}

int launch_attack_maps(const CdqBoard* boards, long long n, unsigned long long* maps, cudaStream_t st) {
    if (n == 0) return 0;
    ProfileScope ps(KK_EVAL, st);
    attack_maps_kernel<<<(unsigned)((n + EVAL_THREADS - 1) / EVAL_THREADS), EVAL_THREADS, 0, st>>>(boards, n, maps);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_eval_terms(const CdqBoard* boards, long long n, int ignore_checkmate, int* terms, double* score,
                      unsigned* flags, cudaStream_t st, float* heuristic_leaf) {
    if (n == 0) return 0;
    const long long blocks = (n + EVAL_THREADS - 1) / EVAL_THREADS;
    ProfileScope ps(KK_EVAL, st);
    if (heuristic_leaf)
        eval_terms_kernel<true><<<(unsigned)blocks, EVAL_THREADS, 0, st>>>(boards, n, ignore_checkmate, terms, score, flags, heuristic_leaf);
    else
        eval_terms_kernel<false><<<(unsigned)blocks, EVAL_THREADS, 0, st>>>(boards, n, ignore_checkmate, terms, score, flags, nullptr);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_encode_planes(const CdqBoard* boards, long long n, float* planes, cudaStream_t st) {
    if (n == 0) return 0;
    const long long threads = n * 12;
    ProfileScope ps(KK_ENCODE, st);
    encode_planes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(boards, n, planes);
    count_launch();
    return (int)cudaGetLastError();
}

} // namespace cdq
