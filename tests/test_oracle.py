"""CPU suite: pins the oracle (oracle/cdq_oracle.c, oracle/oracle.py) against
  * perft known answers (public chessprogramming constants), and
  * the golden vectors produced by the reference's own evaluation.py / board_utils.py /
    neural_network.py (tests/golden/make_golden.py).
"""
import numpy as np
import pytest

PERFT = [
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", [20, 400, 8902, 197281, 4865609]),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -", [48, 2039, 97862, 4085603]),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -", [14, 191, 2812, 43238, 674624]),
    ("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq -", [6, 264, 9467, 422333]),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", [44, 1486, 62379, 2103487]),
    ("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", [46, 2079, 89890, 3894594]),
]


@pytest.mark.parametrize("fen,expected", PERFT)
def test_perft_known_answers(oracle, fen, expected):
    b = oracle.from_fen(fen)
    assert [oracle.perft(b, d + 1) for d in range(len(expected))] == expected


# provisional vectors of SURVEY.md 8c (fen, score)
SURVEY_SCORES = [
    ("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 0.0),
    ("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -", 0.09999999999999998),
    ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -", -0.21500000000000002),
    ("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 0.26),
    ("4k3/8/8/8/8/8/4R3/4K3 b", 56.51),
    ("4k3/3P4/8/8/8/8/8/4K3 b", 51.69),
    ("7k/8/8/8/8/8/8/K6r w", -55.655),
    ("7k/8/8/8/8/8/8/1K5r w", -56.05),
    ("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq -", -50.0),
    ("7k/5Q2/6K1/8/8/8/8/8 b", 0.0),
    ("8/8/4k3/8/8/3K4/8/8 w", 0.0),
    ("rnbqkbnr/ppp1pppp/8/8/3pP3/8/PPPP1PPP/RNBQKBNR b KQkq e3", -0.25),
]


@pytest.mark.parametrize("fen,expected", SURVEY_SCORES)
def test_survey_scores(oracle, fen, expected):
    _, s, _ = oracle.evaluate(oracle.from_fen(fen))
    assert s[0] == expected


def test_survey_quirks(oracle):
    T = oracle.TERM_NAMES.index
    t, _, _ = oracle.evaluate(oracle.from_fen("4k3/8/8/8/8/8/4R3/4K3 b"))
    assert t[0, T("w_mob")] == 17 and t[0, T("b_katt")] == 2      # king-capture counted
    t, _, _ = oracle.evaluate(oracle.from_fen("4k3/3P4/8/8/8/8/8/4K3 b"))
    assert t[0, T("w_mob")] == 13                                  # x4 promotions onto the king
    t, _, _ = oracle.evaluate(oracle.from_fen("7k/8/8/8/8/8/8/K6r w"))
    assert t[0, T("w_katt")] == 0                                  # a1 king is falsy
    t, _, _ = oracle.evaluate(oracle.from_fen("7k/8/8/8/8/8/8/1K5r w"))
    assert t[0, T("w_katt")] == 2
    t, _, _ = oracle.evaluate(oracle.from_fen("rnbqkbnr/ppp1pppp/8/8/3pP3/8/PPPP1PPP/RNBQKBNR b KQkq e3"))
    assert (t[0, T("w_mob")], t[0, T("b_mob")]) == (29, 30)


def test_eval_matches_reference_golden(oracle, eval_golden):
    g = eval_golden
    boards = g["boards"]
    terms, score, flags = oracle.evaluate(boards)
    _, score_ic, _ = oracle.evaluate(boards, ignore_checkmate=True)
    assert np.array_equal(score, g["score"])                    # fp64 bit-exact
    assert np.array_equal(score_ic, g["score_ignore_checkmate"])
    assert np.array_equal(terms[:, 23], g["terminal"])
    assert np.array_equal(flags & 7, g["terminal"].astype(np.uint32))
    has = g["has_terms"]
    assert has.sum() > 1000
    assert np.array_equal(terms[has][:, :23], g["terms"][has][:, :23])
    # both colours' multi-threaded evaluation agree with the single-thread run
    t2, s2, f2 = oracle.evaluate(boards, nthreads=4)
    assert np.array_equal(t2, terms) and np.array_equal(s2, score) and np.array_equal(f2, flags)


def test_planes_match_board_to_tensor(oracle, eval_golden):
    g = eval_golden
    planes = oracle.encode_planes(g["boards"])
    bits = (planes.reshape(len(planes), 12, 64) > 0.5).astype(np.uint64)
    packed = (bits << np.arange(64, dtype=np.uint64)).sum(2, dtype=np.uint64)
    assert np.array_equal(packed, g["planes_packed"])
    assert set(np.unique(planes)) <= {0.0, 1.0}


def test_playouts_are_deterministic_and_match_fixture(oracle, eval_golden):
    n_special = int(eval_golden["n_special"])
    again = oracle.playouts(1024, seed=42, max_plies=199)
    assert np.array_equal(again, eval_golden["boards"][n_special:])


def test_net_oracle_matches_reference_class(oracle, eval_golden, net_golden):
    planes = oracle.encode_planes(net_golden["boards"])
    for tag, scale in (("s1", 1.0), ("s2", 2.0), ("s3", 3.0)):
        v = oracle.net_forward(oracle.init_params(42, scale), planes).reshape(-1)
        assert np.abs(v - net_golden["value_" + tag]).max() < 2e-6


def test_empty_batch(oracle):
    t, s, f = oracle.evaluate(np.zeros(0, oracle.BOARD_DTYPE))
    assert t.shape == (0, 24) and s.shape == (0,) and f.shape == (0,)


def test_move_categories_match_reference_categorize_moves(oracle):
    """oracle.categorize == the reference's own categorize_moves (evaluation.py:232-339) run over the
    shim, move for move (tests/golden/policy_golden.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "policy_golden.npz"))
    boards, pos, move, weight = g["boards"], g["pos"], g["move"], g["weight"]
    checked = 0
    for i in g["picked"]:
        mv, w = oracle.categorize(boards[i:i + 1])
        want = {int(m): int(x) for m, x in zip(move[pos == i], weight[pos == i])}
        got = {int(m): int(x) for m, x in zip(mv, w)}
        assert got == want, i
        checked += len(got)
    assert checked > 6000


def test_eval_matches_reference_on_arbitrary_positions(oracle):
    """Oracle vs the reference's own evaluation.py (over the shim) on 400 arbitrary positions:
    multiple checkers, pins while in check, many sliders, kings in corners, ep squares."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "arbitrary_golden.npz"))
    terms, score, _ = oracle.evaluate(g["boards"])
    assert np.array_equal(score, g["score"])
    has = g["has_terms"]
    assert has.sum() > 200
    assert np.array_equal(terms[has][:, :23], g["terms"][has][:, :23])


def test_python_path_restatement_reproduces_the_reference_scores(oracle, eval_golden):
    """oracle/py_eval.py (the Python-path CPU baseline of bench.py) against the scores recorded from the
    reference's unmodified evaluation.py: bit for bit, with and without ignore_checkmate; its plane
    encoding against the reference's board_to_tensor planes."""
    import py_eval
    g = eval_golden
    idx = list(range(0, 60)) + list(range(60, len(g["boards"]), 23))
    for i in idx:
        if g["boards"][i]["meta"] & (1 << 16):
            continue                                    # rep3 comes from history, which a record does not carry
        b = py_eval.unpack_board(g["boards"][i])
        assert py_eval.fast_evaluate_position(b) == g["score"][i], i
        assert py_eval.fast_evaluate_position(b, ignore_checkmate=True) == g["score_ignore_checkmate"][i], i
    planes = oracle.encode_planes(g["boards"][:40])
    for i in range(40):
        assert np.array_equal(py_eval.board_to_planes(py_eval.unpack_board(g["boards"][i])), planes[i])
    v, sec, scores = py_eval.python_path_throughput(g["boards"][100:132], processes=2)
    assert np.array_equal(scores, g["score"][100:132]) and v > 0


def test_goldens_against_real_python_chess(oracle, eval_golden):
    """Optional: where the REAL python-chess is installed (it is not in this image) and the reference checkout is
    reachable, a sample of the golden vectors is re-derived with the real library under the reference's unmodified
    evaluation.py -- the pin the shim-based goldens cannot give."""
    import importlib.util
    import os
    import sys
    spec = importlib.util.find_spec("chess")
    if spec is None or "pychess_shim" in (spec.origin or "") or not os.path.exists("/root/reference/evaluation.py"):
        pytest.skip("real python-chess (or the reference checkout) is not available")
    import chess
    if not hasattr(chess, "__version__"):
        pytest.skip("`chess` on the path is the restated shim")
    sys.path.insert(0, "/root/reference")
    import evaluation as ref_eval
    import py_eval
    g = eval_golden
    for i in range(0, len(g["boards"]), 7):
        if g["boards"][i]["meta"] & (1 << 16):
            continue
        b = py_eval.unpack_board(g["boards"][i])
        ref_eval.EVAL_CACHE.clear()
        assert ref_eval.fast_evaluate_position(b) == g["score"][i], i
