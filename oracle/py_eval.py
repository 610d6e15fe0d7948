"""TEST / BENCH INFRASTRUCTURE -- never imported by the product (chess-deep-q_b200/).

Python-path restatement of the reference's leaf evaluation, used for ONE thing: timing what the
reference's own code path costs on a box's host cores (`evaluation.py` in
CPython over python-chess + `ChessQNetwork` on torch-CPU at batch 1, as mcts.py:216-219 runs it).
python-chess is not installable here, so `chess` is the restated API in oracle/pychess_shim; the
reference's evaluation.py cannot travel to the GPU box, so its work is restated here with the same
call structure -- the FEN cache key, two legal-move generations for the terminal tests, two more on
turn-forced copies, 128 + <=64 + <=32 `is_attacked_by` calls, ~190 `piece_at` calls per position --
so that the timing means what the reference's would.  tests/test_oracle.py pins it bit for bit against
the golden scores recorded from the reference's unmodified evaluation.py (tests/golden/eval_golden.npz).
Parity status: pinned to the reference's own outputs over the restated dependency."""
import os
import sys

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pychess_shim")


def shim_chess():
    """The `chess` module the restatement runs over: real python-chess if it is ever installed, else the shim."""
    try:
        import chess
        if hasattr(chess, "Board"):
            return chess
    except ImportError:
        pass
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    import chess
    return chess


_VALUE = (None, 1, 3, 3, 5, 9, 0)          # PIECE_VALUES, evaluation.py:8-15 (piece types 1..6)


def fast_evaluate_position(board, ignore_checkmate=False):
    """evaluation.py:30-229 without the memo (every timed position is unique, so the cache never hits;
    the key is still built because the reference pays for it, evaluation.py:40)."""
    chess = shim_chess()
    W, B = chess.WHITE, chess.BLACK
    board.fen()                                                            # :40
    if board.is_checkmate() and not ignore_checkmate:                      # :50-53
        return -50.0 if board.turn == W else 50.0
    if board.is_stalemate() or board.is_insufficient_material() or board.is_repetition(3):   # :55-57
        return 0.0
    material = 0                                                           # :67-71
    for pt in range(1, 7):
        material += (len(board.pieces(pt, W)) - len(board.pieces(pt, B))) * _VALUE[pt]
    mob = []                                                               # :78-84
    for colour in (W, B):
        forced = board.copy()
        forced.turn = colour
        mob.append(len(list(forced.legal_moves)))
    att = {W: set(), B: set()}                                             # :87-94
    for sq in chess.SQUARES:
        for colour in (W, B):
            if board.is_attacked_by(colour, sq):
                att[colour].add(sq)
    mobility = (mob[0] - mob[1]) * 0.1                                     # :96-97
    control = (len(att[W]) - len(att[B])) * 0.05
    king = {W: board.king(W), B: board.king(B)}                            # :100-120; a king on a1 (square 0) is falsy
    pressure = {W: 0, B: 0}
    for colour in (W, B):
        if king[colour]:
            for sq in chess.SQUARES:
                pc = board.piece_at(sq)
                if pc and pc.color != colour and board.is_attacked_by(not colour, king[colour]):
                    pressure[colour] += 1
    safety = (pressure[B] - pressure[W]) * 0.5                             # :123-139
    if king[W] in (chess.G1, chess.C1):
        safety += 1
    if king[B] in (chess.G8, chess.C8):
        safety -= 1
    centre4 = (chess.D4, chess.E4, chess.D5, chess.E5)
    if king[W] in centre4:
        safety -= 2
    if king[B] in centre4:
        safety += 2
    doubled, isolated = {}, {}                                             # :142-178
    for colour in (W, B):
        files = [chess.square_file(sq) for sq in board.pieces(chess.PAWN, colour)]
        doubled[colour] = sum(files.count(f) - 1 for f in set(files))
        isolated[colour] = sum(1 for f in files if not any(g in files for g in (f - 1, f + 1) if 0 <= g < 8))
    pawns = (doubled[B] - doubled[W]) * 0.3 + (isolated[B] - isolated[W]) * 0.2
    ring = (chess.C3, chess.C4, chess.C5, chess.C6, chess.D3, chess.D6, chess.E3, chess.E6,   # :181-190
            chess.F3, chess.F4, chess.F5, chess.F6)
    inner = (chess.D4, chess.D5, chess.E4, chess.E5)
    space = ((sum(1 for s in inner if s in att[W]) - sum(1 for s in inner if s in att[B])) * 0.5
             + (sum(1 for s in ring if s in att[W]) - sum(1 for s in ring if s in att[B])) * 0.1)
    defended = {W: 0, B: 0}                                                # :193-204
    for sq in chess.SQUARES:
        pc = board.piece_at(sq)
        if pc and board.is_attacked_by(pc.color, sq):
            defended[pc.color] += 1
    coordination = (defended[W] - defended[B]) * 0.1
    check = 0                                                              # :207-209
    if board.is_check():
        check = 50 if board.turn == B else -50
    return (material * 1.0 + mobility * 0.3 + control * 0.3 + safety * 0.5 + pawns * 0.4 + space * 0.3   # :212-221
            + coordination * 0.4 + check)


def board_to_planes(board):
    """board_utils.py:12-40 as a numpy array [12,8,8] (64 piece_at calls, like the reference)."""
    import numpy as np
    chess = shim_chess()
    t = np.zeros((12, 8, 8), np.float32)
    for sq in chess.SQUARES:
        pc = board.piece_at(sq)
        if pc is not None:
            t[pc.piece_type - 1 + (0 if pc.color == chess.WHITE else 6), sq >> 3, sq & 7] = 1.0
    return t


def unpack_board(rec):
    """CdqBoard record (include/cdq.h) -> history-less `chess.Board`."""
    chess = shim_chess()
    b = chess.Board(None)
    bb = [int(x) for x in rec["bb"]]
    b.pawns, b.knights, b.bishops, b.rooks, b.queens, b.kings = bb[:6]
    b.occupied_co[chess.WHITE], b.occupied_co[chess.BLACK] = bb[6], bb[7]
    b.occupied = bb[6] | bb[7]
    meta = int(rec["meta"])
    b.turn = bool(meta & 1)
    rights = 0
    for bit, sq in ((1, chess.H1), (2, chess.A1), (3, chess.H8), (4, chess.A8)):
        if (meta >> bit) & 1:
            rights |= 1 << sq
    b.castling_rights = rights
    ep = (meta >> 8) & 127
    b.ep_square = ep - 1 if ep else None
    return b


def _worker(args):
    records, with_net, params = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    out = np.empty(len(records), np.float64)
    t = {k: torch.from_numpy(v) for k, v in params.items()} if with_net else None
    F = torch.nn.functional
    for i, rec in enumerate(records):
        board = unpack_board(rec)
        out[i] = fast_evaluate_position(board)
        if with_net:                                     # mcts.py:216-219: batch 1, fp32, no grad
            with torch.no_grad():
                x = torch.from_numpy(board_to_planes(board)).unsqueeze(0)
                a = F.relu(F.conv2d(x, t["conv1.weight"], t["conv1.bias"], padding=1))
                a = F.relu(F.conv2d(a, t["conv2.weight"], t["conv2.bias"], padding=1))
                a = F.relu(F.linear(a.view(-1, 4096), t["fc1.weight"], t["fc1.bias"]))
                torch.tanh(F.linear(a, t["fc2.weight"], t["fc2.bias"])).item()
    return out


def python_path_throughput(records, params=None, processes=1):
    """positions/s of the Python path over `records`: handcrafted score (+ the batch-1 torch-CPU value
    forward when params is given), on `processes` worker processes.  -> (positions/s, seconds, scores)."""
    import time
    import numpy as np
    t0 = time.perf_counter()
    if processes <= 1:
        scores = _worker((records, params is not None, params))
    else:
        import multiprocessing as mp
        chunks = np.array_split(records, processes)
        with mp.get_context("fork").Pool(processes) as pool:
            scores = np.concatenate(pool.map(_worker, [(c, params is not None, params) for c in chunks]))
    sec = time.perf_counter() - t0
    return len(records) / sec, sec, scores
