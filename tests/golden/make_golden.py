"""Generate golden vectors by running the REFERENCE's own, unmodified code in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

What runs:
  * /root/reference/evaluation.py::fast_evaluate_position and board_utils.py::board_to_tensor,
    imported unmodified.  They need the `chess` package (python-chess), which is not installed
    and not installable here, so they run over oracle/pychess_shim (a pure-Python restatement of
    the few python-chess calls they make).  If real python-chess is importable it is used instead
    (recorded in the fixture as `chess_impl`).
  * /root/reference/neural_network.py::ChessQNetwork and DQNAgent.train, imported unmodified,
    on torch-CPU fp32.
Integer evaluation terms are captured from the live frame of fast_evaluate_position (its local
variables at return), so they are the reference's values, not a re-derivation.

/root/reference does not exist on the GPU box, which is why the outputs are committed.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle"))

try:
    import chess  # real python-chess, if it ever becomes available
    CHESS_IMPL = "python-chess " + getattr(chess, "__version__", "?")
except ImportError:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pychess_shim"))
    import chess
    CHESS_IMPL = "oracle/pychess_shim"
sys.path.insert(0, REF)

import torch  # noqa: E402
import oracle as O  # noqa: E402  (positions + parameter generator only)
import board_utils as ref_board_utils  # noqa: E402  reference, unmodified
import evaluation as ref_evaluation  # noqa: E402  reference, unmodified
import neural_network as ref_nn  # noqa: E402  reference, unmodified

SPECIAL_FENS = [
    "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
    "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10",
    "4k3/8/8/8/8/8/4R3/4K3 b - - 0 1",          # flipped side captures the king (x1)
    "4k3/3P4/8/8/8/8/8/4K3 b - - 0 1",          # ... by a promoting pawn (x4)
    "7k/8/8/8/8/8/8/K6r w - - 0 1",             # king on a1 is falsy (evaluation.py:106)
    "7k/8/8/8/8/8/8/1K5r w - - 0 1",
    "rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 3",  # checkmate
    "7k/5Q2/6K1/8/8/8/8/8 b - - 0 1",           # stalemate
    "8/8/4k3/8/8/3K4/8/8 w - - 0 1",            # insufficient material
    "8/8/4k3/8/8/3KB3/8/8 w - - 0 1",           # K+B v K
    "8/8/4k1b1/8/8/3KB3/8/8 w - - 0 1",         # opposite-colour bishops: sufficient
    "8/8/4kn2/8/8/3KN3/8/8 w - - 0 1",          # K+N v K+N: sufficient
    "rnbqkbnr/ppp1pppp/8/8/3pP3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 2",   # en passant available
    "8/8/8/8/k2Pp2Q/8/8/3K4 b - d3 0 1",        # ep capture would expose the king on the rank
    "8/8/8/2k5/3Pp3/8/8/4K2B b - d3 0 1",       # ep capture while the pawn checks
    "r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1",     # all castling rights
    "r3k2r/8/8/8/8/8/6n1/R3K2R w KQkq - 0 1",   # castling through attacked squares
    "4k3/8/8/8/8/8/8/R3K2R b KQ - 0 1",         # castling counted for the side not to move
    "3k4/8/8/8/8/8/3q4/3K4 w - - 0 1",          # queen adjacent check
    "4k3/8/8/8/7b/8/5P2/4K3 w - - 0 1",         # pinned pawn
    "k7/8/8/8/8/8/1q6/K7 w - - 0 1",
    "3rk3/8/8/8/8/8/3R4/3K4 w - - 0 1",         # pinned rook moves along the pin
    "4k3/8/8/1b6/8/8/4N3/5K2 w - - 0 1",
    "8/5k2/8/8/8/8/1p6/K7 b - - 0 1",           # under-promotion counts
]

# knight shuffle: the start position occurs three times -> is_repetition(3)
REP_LINE = ["g1f3", "g8f6", "f3g1", "f6g8", "g1f3", "g8f6", "f3g1", "f6g8"]


def pack_board(board):
    """chess.Board -> one CdqBoard record, reading only public python-chess attributes."""
    rec = np.zeros(1, O.BOARD_DTYPE)
    bb = rec["bb"][0]
    for i, v in enumerate([board.pawns, board.knights, board.bishops, board.rooks, board.queens,
                           board.kings, board.occupied_co[chess.WHITE], board.occupied_co[chess.BLACK]]):
        bb[i] = v
    c = board.clean_castling_rights()
    meta = int(bool(board.turn))
    for bit, sq in ((1, chess.H1), (2, chess.A1), (3, chess.H8), (4, chess.A8)):
        if c >> sq & 1:
            meta |= 1 << bit
    if board.ep_square is not None:
        meta |= (board.ep_square + 1) << 8
    if board.is_repetition(3):
        meta |= 1 << 16
    rec["meta"][0] = meta
    return rec


def unpack_board(rec):
    """CdqBoard record -> chess.Board (history-less)."""
    b = chess.Board(None)
    bb = [int(x) for x in rec["bb"]]
    b.pawns, b.knights, b.bishops, b.rooks, b.queens, b.kings = bb[:6]
    b.occupied_co[chess.WHITE], b.occupied_co[chess.BLACK] = bb[6], bb[7]
    b.occupied = bb[6] | bb[7]
    meta = int(rec["meta"])
    b.turn = bool(meta & 1)
    cr = 0
    for bit, sq in ((1, chess.H1), (2, chess.A1), (3, chess.H8), (4, chess.A8)):
        if meta >> bit & 1:
            cr |= 1 << sq
    b.castling_rights = cr
    ep = (meta >> 8) & 127
    b.ep_square = ep - 1 if ep else None
    return b


_captured = {}


def _profiler(frame, event, arg):
    if event == "return" and frame.f_code is ref_evaluation.fast_evaluate_position.__code__:
        _captured.clear()
        _captured.update(frame.f_locals)


def reference_eval(board, ignore_checkmate=False):
    """Run the reference function; return (score, dict of its local integer terms or None)."""
    with ref_board_utils.CACHE_LOCK:
        ref_board_utils.EVAL_CACHE.clear()
    _captured.clear()
    sys.setprofile(_profiler)
    try:
        score = ref_evaluation.fast_evaluate_position(board, ignore_checkmate)
    finally:
        sys.setprofile(None)
    loc = dict(_captured)
    if "eval_score" not in loc:      # terminal short-circuit: terms were never computed
        return float(score), None
    T = np.full(O.NTERMS, -999, np.int32)
    names = {"w_mat": "white_material", "b_mat": "black_material", "w_mob": "white_mobility",
             "b_mob": "black_mobility", "w_katt": "white_king_attackers", "b_katt": "black_king_attackers",
             "w_castled": "white_king_castled", "b_castled": "black_king_castled",
             "w_kcentre": "white_king_center", "b_kcentre": "black_king_center",
             "w_dbl": "white_doubled", "b_dbl": "black_doubled", "w_iso": "white_isolated",
             "b_iso": "black_isolated", "w_centre": "white_center", "b_centre": "black_center",
             "w_ext": "white_ext_center", "b_ext": "black_ext_center", "w_def": "white_defended",
             "b_def": "black_defended"}
    for k, v in names.items():
        T[O.TERM_NAMES.index(k)] = int(loc[v])
    T[O.TERM_NAMES.index("w_att")] = len(loc["white_attacked"])
    T[O.TERM_NAMES.index("b_att")] = len(loc["black_attacked"])
    T[O.TERM_NAMES.index("check")] = int(loc["check_score"]) // 50
    return float(score), T


def make_eval_golden():
    boards = [pack_board(chess.Board(f)) for f in SPECIAL_FENS]
    rep = chess.Board()
    for u in REP_LINE:
        mv = [m for m in rep.legal_moves if m.uci() == u][0]
        rep.push(mv)
    boards.append(pack_board(rep))
    live = [chess.Board(f) for f in SPECIAL_FENS] + [rep]
    rnd = O.playouts(1024, seed=42, max_plies=199)
    boards = np.concatenate(boards + [rnd])
    live += [unpack_board(r) for r in rnd]

    n = len(boards)
    score = np.zeros(n)
    score_ic = np.zeros(n)
    terms = np.full((n, O.NTERMS), -999, np.int32)
    has_terms = np.zeros(n, bool)
    planes = np.zeros((n, 12, 8, 8), np.float32)
    terminal = np.zeros(n, np.int32)
    leaf_term = np.full(n, np.nan, np.float32)   # mcts.py:202-210 short-circuit value, nan = none
    for i, b in enumerate(live):
        assert (pack_board(b) == boards[i]).all() if i >= len(SPECIAL_FENS) + 1 else True
        score[i], T = reference_eval(b, False)
        score_ic[i], T2 = reference_eval(b, True)
        T = T if T is not None else T2
        if T is not None:
            terms[i] = T
            has_terms[i] = True
        planes[i] = ref_board_utils.board_to_tensor(b).numpy()
        # mcts.py:202-210, same calls in the same order
        if b.is_checkmate():
            terminal[i] = 1
            leaf_term[i] = -1.0 if b.turn == chess.WHITE else 1.0
        elif b.is_stalemate():
            terminal[i], leaf_term[i] = 2, 0.0
        elif b.is_insufficient_material():
            terminal[i], leaf_term[i] = 3, 0.0
        elif b.is_repetition(3):
            terminal[i], leaf_term[i] = 4, 0.0
    # planes are one-hot bitboards: store them packed (12 x u64) to keep the fixture small
    packed_planes = np.zeros((n, 12), np.uint64)
    for c in range(12):
        bits = planes[:, c].reshape(n, 64) > 0.5
        packed_planes[:, c] = (bits.astype(np.uint64) << np.arange(64, dtype=np.uint64)).sum(1, dtype=np.uint64)
    assert set(np.unique(planes)) <= {0.0, 1.0}
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), boards=boards, score=score,
                        score_ignore_checkmate=score_ic, terms=terms, has_terms=has_terms,
                        planes_packed=packed_planes, terminal=terminal, leaf_terminal=leaf_term,
                        n_special=len(SPECIAL_FENS) + 1, chess_impl=CHESS_IMPL)
    print("eval golden:", n, "positions;", int(has_terms.sum()), "with integer terms;", CHESS_IMPL)
    return boards, planes


def make_arbitrary_golden():
    """fast_evaluate_position of the reference on arbitrary (mostly unreachable) positions."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from randboards import random_boards
    boards = random_boards(400, seed=2024).astype(O.BOARD_DTYPE)
    n = len(boards)
    score = np.zeros(n)
    terms = np.full((n, O.NTERMS), -999, np.int32)
    has = np.zeros(n, bool)
    for i in range(n):
        b = unpack_board(boards[i])
        assert (pack_board(b) == boards[i:i + 1]).all(), i      # castling bits are clean, ep preserved
        score[i], T = reference_eval(b, False)
        if T is not None:
            terms[i], has[i] = T, True
    np.savez_compressed(os.path.join(HERE, "arbitrary_golden.npz"), boards=boards, score=score, terms=terms, has_terms=has)
    print("arbitrary golden:", n, "positions;", int(has.sum()), "with integer terms")


def make_policy_golden(boards):
    """categorize_moves (evaluation.py:232-339) of the reference itself: sampling weight per legal move."""
    weights = {"captures_high_value": 10, "checks": 9, "captures_equal_value": 8, "threatened_piece_moves": 7,
               "developing_moves": 6, "captures_low_value": 5, "center_control": 4, "king_safety": 3,
               "pawn_structure": 2, "other_moves": 1}
    idx, mv, wt = [], [], []
    pick = list(range(0, 29)) + list(range(29, 29 + 220))
    for i in pick:
        b = unpack_board(boards[i])
        cats, cw = ref_evaluation.categorize_moves(b)
        assert cw == weights
        for name, moves in cats.items():
            for m in moves:
                idx.append(i)
                mv.append(m.from_square | (m.to_square << 6) | (((m.promotion - 1) if m.promotion else 0) << 12))
                wt.append(cw[name])
    np.savez_compressed(os.path.join(HERE, "policy_golden.npz"), boards=boards, pos=np.array(idx, np.int32),
                        move=np.array(mv, np.uint16), weight=np.array(wt, np.int8), picked=np.array(pick, np.int32))
    print("policy golden:", len(pick), "positions,", len(mv), "moves")


def load_params(module, params):
    module.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in params.items()})


def make_net_golden(boards, planes):
    out = {}
    x = torch.from_numpy(planes)
    # (a) the reference's own init under torch.manual_seed(42) (main.py:19-21): checks the
    #     class/layout, stored as values only
    torch.manual_seed(42)
    net = ref_nn.ChessQNetwork()
    with torch.no_grad():
        out["value_torchseed42"] = net(x).numpy().reshape(-1)
    # first 4096 weights of each tensor identify the init on boxes with the same torch build
    # (b) portable parameter sets from oracle.init_params (numpy RNG), three weight scales
    for tag, scale in (("s1", 1.0), ("s2", 2.0), ("s3", 3.0)):
        p = O.init_params(seed=42, scale=scale)
        load_params(net, p)
        with torch.no_grad():
            out["value_" + tag] = net(x).numpy().reshape(-1)
    np.savez_compressed(os.path.join(HERE, "net_golden.npz"), boards=boards, **out)
    print("net golden: |value| max", {k: float(np.abs(v).max()) for k, v in out.items()})


def make_train_golden(boards, planes):
    """One DQNAgent.train() step (neural_network.py:218-252) at batch 256 with the memory holding
    exactly `batch` transitions, so random.sample is a permutation and the loss is order-free."""
    B = 256
    idx = np.arange(32, 32 + B)
    states = planes[idx]
    nxt = planes[idx + 1]
    _, sc, fl = O.evaluate(boards[idx + 1])
    reward = (0.01 * np.tanh(sc / 10.0)).astype(np.float32)      # chess_ai.py:183
    done = ((fl & 7) != 0).astype(np.float32)
    res = {}
    for tag, scale in (("s1", 1.0), ("s3", 3.0)):
        random.seed(42)
        torch.manual_seed(42)
        agent = ref_nn.DQNAgent(batch_size=B)
        p = O.init_params(seed=42, scale=scale)
        load_params(agent.q_network, p)
        pt = O.init_params(seed=43, scale=scale)
        load_params(agent.target_q_network, pt)
        for i in range(B):
            # store_transition (neural_network.py:212-216) with pre-encoded tensors
            agent.memory.append((torch.from_numpy(states[i]), None, float(reward[i]),
                                 torch.from_numpy(nxt[i]), float(done[i])))
        loss = agent.train()
        res["loss_" + tag] = np.float64(loss)
        sd = agent.q_network.state_dict()
        for k, v in sd.items():
            a = v.numpy().reshape(-1)
            res["w_%s_%s" % (tag, k)] = a[:: max(1, a.size // 4096)][:4096].copy()
            res["wsum_%s_%s" % (tag, k)] = np.float64(a.astype(np.float64).sum())
        # second step from the updated state (exercises Adam's moment state)
        loss2 = agent.train()
        res["loss2_" + tag] = np.float64(loss2)
        res["fc2w2_" + tag] = agent.q_network.state_dict()["fc2.weight"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train_golden.npz"), idx=idx, reward=reward, done=done, **res)
    print("train golden:", {k: float(v) for k, v in res.items() if k.startswith("loss")})


if __name__ == "__main__":
    b, pl = make_eval_golden()
    make_policy_golden(b)
    make_arbitrary_golden()
    make_net_golden(b, pl)
    make_train_golden(b, pl)
