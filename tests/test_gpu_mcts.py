"""GPU tests of the batched-leaf Russian Doll MCTS driver (BASELINE configs[2]):
leaf values are exact/within tolerance against the oracle, moves are legal, and a self-play game
with GPU batched leaf evaluation runs through legal positions only."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cdq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import chess_deep_q_b200 as m
    return m


def _net(cdq, oracle, scale=2.0):
    p = oracle.init_params(42, scale)
    net = cdq.ChessQNetwork()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return net.cuda(), p


def _successors(oracle, rec):
    return {oracle.push(rec, m)[0].tobytes() for m in oracle.legal_moves(rec)}


def test_search_returns_legal_move_and_exact_leaf_values(cdq, oracle):
    import random
    net, p = _net(cdq, oracle)
    root = oracle.from_fen("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    tree = cdq.BatchedRussianDollMCTS(root, iterations=64, rng=random.Random(1))
    child, move = tree.search(net)
    assert child[0].tobytes() in _successors(oracle, root)
    assert int(tree.N[root[0].tobytes()].sum()) == 64                  # every simulation backed up through the root
    # every expanded node's children are legal successors of that node (sampled, at most width 21)
    for key, kids in list(tree.children.items())[:40]:
        parent = np.frombuffer(key, dtype=oracle.BOARD_DTYPE)
        succ = _successors(oracle, parent)
        assert len(kids) <= 21 and all(k.tobytes() in succ for k in kids)
    # the wave's leaves: values follow mcts.py:196-232 (oracle: exact on terminals, 1e-3 on the network)
    leaves, values = tree.last_leaves
    _, _, flags = oracle.evaluate(leaves)
    v = oracle.net_forward(p, oracle.encode_planes(leaves)).reshape(-1)
    ref = np.clip(v, -1, 1)
    white = (leaves["meta"] & 1).astype(bool)
    mate = (flags & 0x10) != 0
    ref[mate & white], ref[mate & ~white] = -1.0, 1.0
    ref[(flags & 0xE0) != 0] = 0.0
    assert np.abs(values - ref).max() <= 1e-3
    assert np.array_equal(values[mate], ref[mate])
    # the reported move transforms the root into the chosen child
    frm, to, promo = move
    legal = oracle.legal_moves(root)
    enc = frm | (to << 6) | ((promo - 1 if promo else 0) << 12)
    assert enc in set(int(m) for m in legal)
    assert oracle.push(root, enc)[0].tobytes() == child[0].tobytes()
    assert cdq.move_uci(root[0], child[0])[:4] == "abcdefgh"[frm & 7] + str((frm >> 3) + 1) + "abcdefgh"[to & 7] + str((to >> 3) + 1)


def test_search_prefers_mate_in_one_for_white(cdq, oracle):
    import random
    net, _ = _net(cdq, oracle, scale=1.0)
    root = oracle.from_fen("6k1/5ppp/8/8/8/8/8/R5K1 w - - 0 1")       # Ra8#
    tree = cdq.BatchedRussianDollMCTS(root, iterations=200, rng=random.Random(3))
    child, move = tree.search(net)
    assert move[:2] == (0, 56)
    _, _, f = oracle.evaluate(child)
    assert f[0] & 0x10


def test_terminal_root(cdq, oracle):
    net, _ = _net(cdq, oracle)
    mated = oracle.from_fen("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 3")
    child, move = cdq.BatchedRussianDollMCTS(mated, iterations=8).search(net)
    assert child is None and move is None


def test_self_play_game_with_batched_leaf_evaluation(cdq, oracle):
    """BASELINE configs[2] as a parity case: Russian Doll widths 21/13/8/5/3/2/1, batched GPU leaves."""
    net, _ = _net(cdq, oracle)
    positions, scores = cdq.self_play_game(net, max_moves=24, iterations=50, seed=42)
    assert len(positions) >= 2 and len(scores) >= len(positions) - 1
    assert positions[0].tobytes() == oracle.startpos()[0].tobytes()
    for a, b in zip(positions[:-1], positions[1:]):
        assert b.tobytes() in _successors(oracle, np.array([a], dtype=oracle.BOARD_DTYPE))
    _, s_ref, _ = oracle.evaluate(positions[: len(scores)])
    assert np.array_equal(np.array(scores), s_ref)                     # reward-shaping scores are the bit-exact eval
    again, _ = cdq.self_play_game(net, max_moves=24, iterations=50, seed=42)
    assert np.array_equal(again, positions)                           # seeded -> reproducible
