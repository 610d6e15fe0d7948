"""Board representation: mirrors the reference's board_utils.py (board_to_tensor, :12-40) and the
archived boards_to_tensor_batch (archive/board_utils.py:2-10), on top of the packed wire format of
include/cdq.h (CdqBoard, 72 bytes)."""
import numpy as np

from . import _lib

BOARD_DTYPE = np.dtype([("bb", "<u8", (8,)), ("meta", "<u8")])
assert BOARD_DTYPE.itemsize == 72

_A1, _H1, _A8, _H8 = 0, 7, 56, 63


def pack_board(board):
    """python-chess `Board` (anything exposing its public attributes) -> one CdqBoard record.

    Reads exactly what evaluation.py / board_to_tensor depend on: the eight bitboards, the turn,
    `clean_castling_rights()`, the raw `ep_square` and `is_repetition(3)` (history dependent, so it
    is computed here on the host; evaluation.py:55, mcts.py:209)."""
    rec = np.zeros(1, BOARD_DTYPE)
    bb = rec["bb"][0]
    bb[0], bb[1], bb[2] = board.pawns, board.knights, board.bishops
    bb[3], bb[4], bb[5] = board.rooks, board.queens, board.kings
    bb[6], bb[7] = board.occupied_co[True], board.occupied_co[False]
    c = board.clean_castling_rights()
    meta = int(bool(board.turn))
    for bit, sq in ((1, _H1), (2, _A1), (3, _H8), (4, _A8)):
        if (c >> sq) & 1:
            meta |= 1 << bit
    if board.ep_square is not None:
        meta |= (board.ep_square + 1) << 8
    if board.is_repetition(3):
        meta |= 1 << 16
    rec["meta"][0] = meta
    return rec


def pack_boards(boards):
    if len(boards) == 0:
        return np.zeros(0, BOARD_DTYPE)
    return np.concatenate([pack_board(b) for b in boards])


def as_records(boards):
    """-> contiguous numpy BOARD_DTYPE array (host).  Accepts records, (n,9) integer arrays, one
    board object or a list of board objects."""
    if isinstance(boards, np.ndarray):
        if boards.dtype == BOARD_DTYPE:
            return np.ascontiguousarray(boards)
        if boards.ndim == 2 and boards.shape[1] == 9:
            return np.ascontiguousarray(boards.astype(np.uint64, copy=False)).view(BOARD_DTYPE).reshape(-1)
        raise TypeError("unsupported board array: dtype %s shape %s" % (boards.dtype, boards.shape))
    if hasattr(boards, "occupied_co"):
        return pack_board(boards)
    return pack_boards(list(boards))


def to_device(boards):
    """-> CUDA int64 tensor [n, 9] holding CdqBoard records (bit patterns of the u64 fields)."""
    import torch
    dev = _lib.require_cuda()
    if isinstance(boards, torch.Tensor):
        if boards.dtype != torch.int64 or boards.dim() != 2 or boards.shape[1] != 9:
            raise TypeError("board tensor must be int64 [n, 9]")
        return boards.to(dev).contiguous()
    rec = as_records(boards)
    t = torch.from_numpy(rec.view(np.uint64).reshape(-1, 9).view(np.int64))
    return t.to(dev)


def boards_to_tensor_batch(boards):
    """[B,12,8,8] fp32 one-hot planes on the CUDA device (archive/board_utils.py:2-10)."""
    import torch
    d = to_device(boards)
    n = d.shape[0]
    planes = torch.empty((n, 12, 8, 8), dtype=torch.float32, device=d.device)
    L = _lib.lib()
    _lib.check(L.cdq_encode_planes(d.data_ptr(), n, planes.data_ptr(), _lib.stream_ptr()))
    return planes


def board_to_tensor(board):
    """[12,8,8] fp32 planes of one board (board_utils.py:12-40); channel = piece type (P,N,B,R,Q,K)
    + 6 for black, row = rank, col = file.  Returned on the CUDA device."""
    return boards_to_tensor_batch(board)[0]


def planes_to_packed(planes, turn_white=True):
    """Inverse of the plane encoding for the ChessQNetwork drop-in: one-hot planes [B,12,8,8] ->
    packed records (bitboards only; the network does not read turn/castling/ep)."""
    import torch
    x = planes.reshape(-1, 12, 64)
    weights = (torch.ones(64, dtype=torch.int64, device=x.device) << torch.arange(64, device=x.device))
    bits = ((x > 0.5).to(torch.int64) * weights).sum(2)  # [B,12] two's complement bit patterns
    out = torch.zeros((x.shape[0], 9), dtype=torch.int64, device=x.device)
    out[:, 0:6] = bits[:, 0:6] | bits[:, 6:12]
    white = bits[:, 0]
    black = bits[:, 6]
    for c in range(1, 6):
        white = white | bits[:, c]
        black = black | bits[:, 6 + c]
    out[:, 6] = white
    out[:, 7] = black
    out[:, 8] = 1 if turn_white else 0
    return out


class Move:
    """Stand-in for chess.Move where python-chess is not installed: from_square, to_square,
    promotion (piece type 2..5 or None), uci(); compares equal to anything with the same uci()."""
    __slots__ = ("from_square", "to_square", "promotion")

    def __init__(self, from_square, to_square, promotion=None):
        self.from_square, self.to_square, self.promotion = int(from_square), int(to_square), promotion

    def uci(self):
        name = lambda s: "abcdefgh"[s & 7] + str((s >> 3) + 1)  # noqa: E731
        return name(self.from_square) + name(self.to_square) + ("" if self.promotion is None else " nbrq"[self.promotion])

    def __eq__(self, other):
        return hasattr(other, "uci") and other.uci() == self.uci()

    def __hash__(self):
        return hash(self.uci())

    def __repr__(self):
        return "Move.from_uci(%r)" % self.uci()


def _make_move(board, frm, to, promo):
    """chess.Move when the caller handed us a python-chess board (so `board.push(move)` works), else Move."""
    if hasattr(board, "legal_moves"):
        try:
            import chess
            return chess.Move(frm, to, promotion=promo)
        except ImportError:
            pass
    return Move(frm, to, promo)
