"""Minimal pure-Python stand-in for the `chess` package (python-chess) -- TEST INFRASTRUCTURE, never shipped.

What it is.  python-chess is the un-vendored, unpinned dependency (`chess>=1.9.0`, requirements.txt:4) that
holds all bitboard arithmetic of the reference's evaluation path; it is not installable in this image (no
network).  This module is a FAITHFUL RESTATEMENT of the part of its API that the reference's board_utils.py /
evaluation.py / mcts.py touch: it follows the library's published algorithm step by step -- `_attackers_mask`,
slider blockers and `pin_mask`, `_generate_evasions`, `_is_safe`, `_ep_skewered`, `generate_castling_moves`
through `clean_castling_rights`, `is_insufficient_material`, `_transposition_key` -- and keeps its method names
where the reference calls them, because the quirks that leak into the reference's scores (captures of the enemy
king when the turn is forced, the raw ep square, promotions counted four times) ARE that algorithm.  It is
therefore not an independent derivation and not a clean-room work: python-chess is GPL-3.0, and this file stays
in oracle/ as a checker -- the product (chess-deep-q_b200/) neither imports nor links it
(tests/test_host_abi.py::test_product_never_imports_the_oracle).

What it pins.  The reference's OWN unmodified evaluation.py is imported over this module to record the golden
vectors (tests/golden/make_golden.py).  So "bit-exact against the reference" means: against the reference's code
running on this restatement of its dependency, with both this module and oracle/cdq_oracle.c pinned to the six
public perft suites.  The C oracle shares the algorithm (per-move generation + `_is_safe`), not the code; the
CUDA kernels use a different, set-wise method.  If the real library is ever importable it is preferred:
make_golden.py tries it first, and tests/test_oracle.py::test_goldens_against_real_python_chess re-derives a
sample of the golden vectors with it.
"""
WHITE, BLACK = True, False
COLORS = [WHITE, BLACK]
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = range(1, 7)
PIECE_TYPES = [PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING]
PIECE_SYMBOLS = [None, "p", "n", "b", "r", "q", "k"]
FILE_NAMES = "abcdefgh"
RANK_NAMES = "12345678"
SQUARES = list(range(64))
SQUARE_NAMES = [f + r for r in RANK_NAMES for f in FILE_NAMES]
for _i, _n in enumerate(SQUARE_NAMES):
    globals()[_n.upper()] = _i
STARTING_FEN = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"

BB_ALL = (1 << 64) - 1
BB_SQUARES = [1 << s for s in SQUARES]
BB_RANKS = [0xFF << (8 * r) for r in range(8)]
BB_FILES = [0x0101010101010101 << f for f in range(8)]
BB_RANK_1, BB_RANK_8 = BB_RANKS[0], BB_RANKS[7]
BB_LIGHT_SQUARES = 0x55AA55AA55AA55AA
BB_DARK_SQUARES = 0xAA55AA55AA55AA55


def square(file_index, rank_index):
    return rank_index * 8 + file_index


def square_file(sq):
    return sq & 7


def square_rank(sq):
    return sq >> 3


def msb(bb):
    return bb.bit_length() - 1


def scan_reversed(bb):
    while bb:
        r = bb.bit_length() - 1
        yield r
        bb ^= 1 << r


def popcount(bb):
    return bin(bb).count("1")


def _steps(sq, deltas):
    out = 0
    f0, r0 = sq & 7, sq >> 3
    for df, dr in deltas:
        f, r = f0 + df, r0 + dr
        if 0 <= f < 8 and 0 <= r < 8:
            out |= 1 << (r * 8 + f)
    return out


def _slide(sq, occ, deltas):
    out = 0
    for df, dr in deltas:
        f, r = sq & 7, sq >> 3
        while True:
            f += df
            r += dr
            if not (0 <= f < 8 and 0 <= r < 8):
                break
            b = 1 << (r * 8 + f)
            out |= b
            if occ & b:
                break
    return out


_RANK_D = [(1, 0), (-1, 0)]
_FILE_D = [(0, 1), (0, -1)]
_DIAG_D = [(1, 1), (1, -1), (-1, 1), (-1, -1)]
BB_KNIGHT_ATTACKS = [_steps(s, [(1, 2), (2, 1), (2, -1), (1, -2), (-1, -2), (-2, -1), (-2, 1), (-1, 2)]) for s in SQUARES]
BB_KING_ATTACKS = [_steps(s, [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)]) for s in SQUARES]
BB_PAWN_ATTACKS = [[_steps(s, [(-1, -1), (1, -1)]) for s in SQUARES],  # [BLACK]
                   [_steps(s, [(-1, 1), (1, 1)]) for s in SQUARES]]    # [WHITE]
_RANK_LINE = [_slide(s, 0, _RANK_D) for s in SQUARES]
_FILE_LINE = [_slide(s, 0, _FILE_D) for s in SQUARES]
_DIAG_LINE = [_slide(s, 0, _DIAG_D) for s in SQUARES]


def _line_deltas(a, b):
    """unit step from a towards b if aligned, else None"""
    df, dr = (b & 7) - (a & 7), (b >> 3) - (a >> 3)
    if a == b:
        return None
    if df == 0 or dr == 0 or abs(df) == abs(dr):
        return ((df > 0) - (df < 0), (dr > 0) - (dr < 0))
    return None


def ray(a, b):
    d = _line_deltas(a, b)
    if d is None:
        return 0
    return _slide(a, 0, [d, (-d[0], -d[1])]) | (1 << a)


def between(a, b):
    d = _line_deltas(a, b)
    if d is None:
        return 0
    return _slide(a, 1 << b, [d]) & ~(1 << b)


class Piece:
    def __init__(self, piece_type, color):
        self.piece_type = piece_type
        self.color = color

    def symbol(self):
        s = PIECE_SYMBOLS[self.piece_type]
        return s.upper() if self.color else s


class Move:
    def __init__(self, from_square, to_square, promotion=None):
        self.from_square = from_square
        self.to_square = to_square
        self.promotion = promotion

    def uci(self):
        s = SQUARE_NAMES[self.from_square] + SQUARE_NAMES[self.to_square]
        return s + (PIECE_SYMBOLS[self.promotion] if self.promotion else "")

    def __eq__(self, o):
        return (self.from_square, self.to_square, self.promotion) == (o.from_square, o.to_square, o.promotion)

    def __hash__(self):
        return hash((self.from_square, self.to_square, self.promotion))

    def __repr__(self):
        return "Move.from_uci(%r)" % self.uci()


class SquareSet:
    def __init__(self, mask=0):
        self.mask = mask

    def __len__(self):
        return popcount(self.mask)

    def __iter__(self):
        bb = self.mask
        while bb:
            low = bb & -bb
            yield low.bit_length() - 1
            bb ^= low

    def __contains__(self, sq):
        return bool(self.mask >> sq & 1)

    def __bool__(self):
        return bool(self.mask)


class LegalMoveGenerator:
    def __init__(self, board):
        self.board = board

    def __iter__(self):
        return self.board.generate_legal_moves()

    def __len__(self):
        return sum(1 for _ in self)

    def count(self):
        return len(self)

    def __bool__(self):
        return any(True for _ in self)

    def __contains__(self, move):
        return any(m == move for m in self)


class Board:
    def __init__(self, fen=STARTING_FEN):
        self.pawns = self.knights = self.bishops = self.rooks = self.queens = self.kings = 0
        self.promoted = 0
        self.occupied_co = [0, 0]
        self.occupied = 0
        self.turn = WHITE
        self.castling_rights = 0
        self.ep_square = None
        self.halfmove_clock = 0
        self.fullmove_number = 1
        self.move_stack = []
        self._keys = []  # position keys before each pushed move, for is_repetition
        if fen is not None:
            self.set_fen(fen)

    # ------------------------------------------------------------------ setup / io
    def set_fen(self, fen):
        parts = fen.split()
        r, f = 7, 0
        for ch in parts[0]:
            if ch == "/":
                r, f = r - 1, 0
            elif ch.isdigit():
                f += int(ch)
            else:
                self._set_piece_at(r * 8 + f, PIECE_SYMBOLS.index(ch.lower()), ch.isupper())
                f += 1
        self.turn = WHITE if len(parts) < 2 or parts[1] == "w" else BLACK
        cr = parts[2] if len(parts) > 2 else "-"
        self.castling_rights = 0
        for ch, sq in (("K", H1), ("Q", A1), ("k", H8), ("q", A8)):
            if ch in cr:
                self.castling_rights |= BB_SQUARES[sq]
        ep = parts[3] if len(parts) > 3 else "-"
        self.ep_square = None if ep == "-" else SQUARE_NAMES.index(ep)
        self.halfmove_clock = int(parts[4]) if len(parts) > 4 else 0
        self.fullmove_number = int(parts[5]) if len(parts) > 5 else 1

    def _set_piece_at(self, sq, piece_type, color):
        self._remove_piece_at(sq)
        m = BB_SQUARES[sq]
        name = ("pawns", "knights", "bishops", "rooks", "queens", "kings")[piece_type - 1]
        setattr(self, name, getattr(self, name) | m)
        self.occupied |= m
        self.occupied_co[color] |= m

    def _remove_piece_at(self, sq):
        pt = self.piece_type_at(sq)
        m = BB_SQUARES[sq]
        if pt is None:
            return None
        name = ("pawns", "knights", "bishops", "rooks", "queens", "kings")[pt - 1]
        setattr(self, name, getattr(self, name) & ~m)
        self.occupied &= ~m
        self.occupied_co[WHITE] &= ~m
        self.occupied_co[BLACK] &= ~m
        return pt

    def piece_type_at(self, sq):
        m = BB_SQUARES[sq]
        if not self.occupied & m:
            return None
        for pt, bb in ((PAWN, self.pawns), (KNIGHT, self.knights), (BISHOP, self.bishops),
                       (ROOK, self.rooks), (QUEEN, self.queens)):
            if bb & m:
                return pt
        return KING

    def piece_at(self, sq):
        pt = self.piece_type_at(sq)
        if pt is None:
            return None
        return Piece(pt, bool(self.occupied_co[WHITE] & BB_SQUARES[sq]))

    def pieces_mask(self, piece_type, color):
        bb = (self.pawns, self.knights, self.bishops, self.rooks, self.queens, self.kings)[piece_type - 1]
        return bb & self.occupied_co[color]

    def pieces(self, piece_type, color):
        return SquareSet(self.pieces_mask(piece_type, color))

    def king(self, color):
        k = self.occupied_co[color] & self.kings & ~self.promoted
        return msb(k) if k else None

    def copy(self, stack=True):
        b = Board(None)
        for a in ("pawns", "knights", "bishops", "rooks", "queens", "kings", "promoted", "occupied",
                  "turn", "castling_rights", "ep_square", "halfmove_clock", "fullmove_number"):
            setattr(b, a, getattr(self, a))
        b.occupied_co = list(self.occupied_co)
        if stack:
            b.move_stack = list(self.move_stack)
            b._keys = list(self._keys)
        return b

    def board_fen(self):
        rows = []
        for r in range(7, -1, -1):
            row, empty = "", 0
            for f in range(8):
                p = self.piece_at(r * 8 + f)
                if p is None:
                    empty += 1
                else:
                    row += (str(empty) if empty else "") + p.symbol()
                    empty = 0
            rows.append(row + (str(empty) if empty else ""))
        return "/".join(rows)

    def has_legal_en_passant(self):
        return self.ep_square is not None and any(self.generate_legal_ep())

    def generate_legal_ep(self):
        for m in self.generate_legal_moves():
            if self.is_en_passant(m):
                yield m

    def castling_xfen(self):
        c = self.clean_castling_rights()
        s = "".join(ch for ch, sq in (("K", H1), ("Q", A1), ("k", H8), ("q", A8)) if c & BB_SQUARES[sq])
        return s or "-"

    def fen(self):
        ep = SQUARE_NAMES[self.ep_square] if self.has_legal_en_passant() else "-"
        return " ".join([self.board_fen(), "w" if self.turn else "b", self.castling_xfen(), ep,
                         str(self.halfmove_clock), str(self.fullmove_number)])

    # ------------------------------------------------------------------ attacks
    def attacks_mask(self, sq):
        m = BB_SQUARES[sq]
        if m & self.pawns:
            return BB_PAWN_ATTACKS[bool(m & self.occupied_co[WHITE])][sq]
        if m & self.knights:
            return BB_KNIGHT_ATTACKS[sq]
        if m & self.kings:
            return BB_KING_ATTACKS[sq]
        a = 0
        if m & (self.bishops | self.queens):
            a |= _slide(sq, self.occupied, _DIAG_D)
        if m & (self.rooks | self.queens):
            a |= _slide(sq, self.occupied, _RANK_D) | _slide(sq, self.occupied, _FILE_D)
        return a

    def _attackers_mask(self, color, sq, occupied):
        qr = self.queens | self.rooks
        qb = self.queens | self.bishops
        a = ((BB_KING_ATTACKS[sq] & self.kings) | (BB_KNIGHT_ATTACKS[sq] & self.knights) |
             (_slide(sq, occupied, _RANK_D) & qr) | (_slide(sq, occupied, _FILE_D) & qr) |
             (_slide(sq, occupied, _DIAG_D) & qb) | (BB_PAWN_ATTACKS[not color][sq] & self.pawns))
        return a & self.occupied_co[color]

    def attackers_mask(self, color, sq):
        return self._attackers_mask(color, sq, self.occupied)

    def is_attacked_by(self, color, sq):
        return bool(self.attackers_mask(color, sq))

    def checkers_mask(self):
        k = self.king(self.turn)
        return 0 if k is None else self.attackers_mask(not self.turn, k)

    def is_check(self):
        return bool(self.checkers_mask())

    def pin_mask(self, color, sq):
        king = self.king(color)
        if king is None:
            return BB_ALL
        sm = BB_SQUARES[sq]
        for line, sliders in ((_FILE_LINE, self.rooks | self.queens), (_RANK_LINE, self.rooks | self.queens),
                              (_DIAG_LINE, self.bishops | self.queens)):
            rays = line[king]
            if rays & sm:
                for sniper in scan_reversed(rays & sliders & self.occupied_co[not color]):
                    if between(sniper, king) & (self.occupied | sm) == sm:
                        return ray(king, sniper)
                break
        return BB_ALL

    # ------------------------------------------------------------------ move generation
    def clean_castling_rights(self):
        c = self.castling_rights & self.rooks
        w = c & BB_RANK_1 & self.occupied_co[WHITE] & (BB_SQUARES[A1] | BB_SQUARES[H1])
        b = c & BB_RANK_8 & self.occupied_co[BLACK] & (BB_SQUARES[A8] | BB_SQUARES[H8])
        if not self.occupied_co[WHITE] & self.kings & ~self.promoted & BB_SQUARES[E1]:
            w = 0
        if not self.occupied_co[BLACK] & self.kings & ~self.promoted & BB_SQUARES[E8]:
            b = 0
        return w | b

    def _attacked_for_king(self, path, occupied):
        return any(self._attackers_mask(not self.turn, sq, occupied) for sq in scan_reversed(path))

    def generate_castling_moves(self, from_mask=BB_ALL, to_mask=BB_ALL):
        backrank = BB_RANK_1 if self.turn == WHITE else BB_RANK_8
        king = self.occupied_co[self.turn] & self.kings & ~self.promoted & backrank & from_mask
        king &= -king
        if not king:
            return
        bb_c, bb_d = BB_FILES[2] & backrank, BB_FILES[3] & backrank
        bb_f, bb_g = BB_FILES[5] & backrank, BB_FILES[6] & backrank
        for candidate in scan_reversed(self.clean_castling_rights() & backrank & to_mask):
            rook = BB_SQUARES[candidate]
            a_side = rook < king
            king_to = bb_c if a_side else bb_g
            rook_to = bb_d if a_side else bb_f
            king_path = between(msb(king), msb(king_to))
            rook_path = between(candidate, msb(rook_to))
            if not ((self.occupied ^ king ^ rook) & (king_path | rook_path | king_to | rook_to) or
                    self._attacked_for_king(king_path | king, self.occupied ^ king) or
                    self._attacked_for_king(king_to, self.occupied ^ king ^ rook ^ rook_to)):
                yield Move(msb(king), msb(king_to))

    def generate_pseudo_legal_ep(self, from_mask=BB_ALL, to_mask=BB_ALL):
        if not self.ep_square or not BB_SQUARES[self.ep_square] & to_mask:
            return
        if BB_SQUARES[self.ep_square] & self.occupied:
            return
        capturers = (self.pawns & self.occupied_co[self.turn] & from_mask &
                     BB_PAWN_ATTACKS[not self.turn][self.ep_square] & BB_RANKS[4 if self.turn else 3])
        for capturer in scan_reversed(capturers):
            yield Move(capturer, self.ep_square)

    def generate_pseudo_legal_moves(self, from_mask=BB_ALL, to_mask=BB_ALL):
        our = self.occupied_co[self.turn]
        for f in scan_reversed(our & ~self.pawns & from_mask):
            for t in scan_reversed(self.attacks_mask(f) & ~our & to_mask):
                yield Move(f, t)
        if from_mask & self.kings:
            yield from self.generate_castling_moves(from_mask, to_mask)
        pawns = self.pawns & our & from_mask
        if not pawns:
            return
        for f in scan_reversed(pawns):
            targets = BB_PAWN_ATTACKS[self.turn][f] & self.occupied_co[not self.turn] & to_mask
            for t in scan_reversed(targets):
                if square_rank(t) in (0, 7):
                    for pr in (QUEEN, ROOK, BISHOP, KNIGHT):
                        yield Move(f, t, pr)
                else:
                    yield Move(f, t)
        if self.turn == WHITE:
            single = pawns << 8 & ~self.occupied
            double = single << 8 & ~self.occupied & (BB_RANKS[2] | BB_RANKS[3])
        else:
            single = pawns >> 8 & ~self.occupied
            double = single >> 8 & ~self.occupied & (BB_RANKS[5] | BB_RANKS[4])
        single &= to_mask
        double &= to_mask
        for t in scan_reversed(single):
            f = t + (8 if self.turn == BLACK else -8)
            if square_rank(t) in (0, 7):
                for pr in (QUEEN, ROOK, BISHOP, KNIGHT):
                    yield Move(f, t, pr)
            else:
                yield Move(f, t)
        for t in scan_reversed(double):
            yield Move(t + (16 if self.turn == BLACK else -16), t)
        if self.ep_square:
            yield from self.generate_pseudo_legal_ep(from_mask, to_mask)

    def _slider_blockers(self, king):
        rq, bq = self.rooks | self.queens, self.bishops | self.queens
        snipers = ((_RANK_LINE[king] | _FILE_LINE[king]) & rq) | (_DIAG_LINE[king] & bq)
        blockers = 0
        for sniper in scan_reversed(snipers & self.occupied_co[not self.turn]):
            b = between(king, sniper) & self.occupied
            if b and BB_SQUARES[msb(b)] == b:
                blockers |= b
        return blockers & self.occupied_co[self.turn]

    def is_en_passant(self, move):
        return (self.ep_square == move.to_square and bool(self.pawns & BB_SQUARES[move.from_square]) and
                abs(move.to_square - move.from_square) in (7, 9) and
                not self.occupied & BB_SQUARES[move.to_square])

    def is_castling(self, move):
        if self.kings & BB_SQUARES[move.from_square]:
            diff = square_file(move.from_square) - square_file(move.to_square)
            return abs(diff) > 1 or bool(self.rooks & self.occupied_co[self.turn] & BB_SQUARES[move.to_square])
        return False

    def _ep_skewered(self, king, capturer):
        last_double = self.ep_square + (-8 if self.turn == WHITE else 8)
        occupancy = (self.occupied & ~BB_SQUARES[last_double] & ~BB_SQUARES[capturer]) | BB_SQUARES[self.ep_square]
        horizontal = self.occupied_co[not self.turn] & (self.rooks | self.queens)
        if _slide(king, occupancy, _RANK_D) & horizontal:
            return True
        diagonal = self.occupied_co[not self.turn] & (self.bishops | self.queens)
        if _slide(king, occupancy, _DIAG_D) & diagonal:
            return True
        return False

    def _is_safe(self, king, blockers, move):
        if move.from_square == king:
            if self.is_castling(move):
                return True
            return not self.is_attacked_by(not self.turn, move.to_square)
        elif self.is_en_passant(move):
            return bool(self.pin_mask(self.turn, move.from_square) & BB_SQUARES[move.to_square] and
                        not self._ep_skewered(king, move.from_square))
        else:
            return bool(not blockers & BB_SQUARES[move.from_square] or
                        ray(move.from_square, move.to_square) & BB_SQUARES[king])

    def _generate_evasions(self, king, checkers, from_mask=BB_ALL, to_mask=BB_ALL):
        sliders = checkers & (self.bishops | self.rooks | self.queens)
        attacked = 0
        for checker in scan_reversed(sliders):
            attacked |= ray(king, checker) & ~BB_SQUARES[checker]
        if BB_SQUARES[king] & from_mask:
            for t in scan_reversed(BB_KING_ATTACKS[king] & ~self.occupied_co[self.turn] & ~attacked & to_mask):
                yield Move(king, t)
        checker = msb(checkers)
        if BB_SQUARES[checker] == checkers:
            target = between(king, checker) | checkers
            yield from self.generate_pseudo_legal_moves(~self.kings & from_mask, target & to_mask)
            if self.ep_square and not BB_SQUARES[self.ep_square] & target:
                last_double = self.ep_square + (-8 if self.turn == WHITE else 8)
                if last_double == checker:
                    yield from self.generate_pseudo_legal_ep(from_mask, to_mask)

    def generate_legal_moves(self, from_mask=BB_ALL, to_mask=BB_ALL):
        king_mask = self.kings & self.occupied_co[self.turn]
        if king_mask:
            king = msb(king_mask)
            blockers = self._slider_blockers(king)
            checkers = self.attackers_mask(not self.turn, king)
            gen = (self._generate_evasions(king, checkers, from_mask, to_mask) if checkers
                   else self.generate_pseudo_legal_moves(from_mask, to_mask))
            for move in gen:
                if self._is_safe(king, blockers, move):
                    yield move
        else:
            yield from self.generate_pseudo_legal_moves(from_mask, to_mask)

    @property
    def legal_moves(self):
        return LegalMoveGenerator(self)

    # ------------------------------------------------------------------ rules
    def is_checkmate(self):
        return self.is_check() and not any(self.generate_legal_moves())

    def is_stalemate(self):
        return (not self.is_check()) and not any(self.generate_legal_moves())

    def has_insufficient_material(self, color):
        if self.occupied_co[color] & (self.pawns | self.rooks | self.queens):
            return False
        if self.occupied_co[color] & self.knights:
            return (popcount(self.occupied_co[color]) <= 2 and
                    not (self.occupied_co[not color] & ~self.kings & ~self.queens))
        if self.occupied_co[color] & self.bishops:
            same = (not self.bishops & BB_DARK_SQUARES) or (not self.bishops & BB_LIGHT_SQUARES)
            return same and not self.pawns and not self.knights
        return True

    def is_insufficient_material(self):
        return all(self.has_insufficient_material(c) for c in COLORS)

    def _transposition_key(self):
        return (self.pawns, self.knights, self.bishops, self.rooks, self.queens, self.kings,
                self.occupied_co[WHITE], self.occupied_co[BLACK], self.turn,
                self.clean_castling_rights(), self.ep_square if self.has_legal_en_passant() else None)

    def is_repetition(self, count=3):
        key = self._transposition_key()
        n = 1
        for k in reversed(self._keys):
            if k == key:
                n += 1
                if n >= count:
                    return True
        return n >= count

    def is_game_over(self):
        return self.is_checkmate() or self.is_stalemate() or self.is_insufficient_material()

    def push(self, move):
        self._keys.append(self._transposition_key())
        self.move_stack.append(move)
        ep_square, self.ep_square = self.ep_square, None
        f, t = move.from_square, move.to_square
        from_bb, to_bb = BB_SQUARES[f], BB_SQUARES[t]
        castling = self.is_castling(move)
        piece_type = self._remove_piece_at(f)
        captured = self.piece_type_at(t)
        self.castling_rights &= ~to_bb & ~from_bb
        if piece_type == KING:
            self.castling_rights &= ~(BB_RANK_1 if self.turn == WHITE else BB_RANK_8)
        if piece_type == PAWN:
            diff = t - f
            if diff == 16 and square_rank(f) == 1:
                self.ep_square = f + 8
            elif diff == -16 and square_rank(f) == 6:
                self.ep_square = f - 8
            elif t == ep_square and abs(diff) in (7, 9) and not captured:
                self._remove_piece_at(ep_square + (-8 if self.turn == WHITE else 8))
        if move.promotion:
            piece_type = move.promotion
        if castling:
            a_side = square_file(t) < square_file(f)
            base = 0 if self.turn == WHITE else 56
            self._remove_piece_at(base + (0 if a_side else 7))
            self._set_piece_at(base + (2 if a_side else 6), KING, self.turn)
            self._set_piece_at(base + (3 if a_side else 5), ROOK, self.turn)
        else:
            self._set_piece_at(t, piece_type, self.turn)
        self.turn = not self.turn
