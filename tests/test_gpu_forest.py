"""GPU tests of the device-resident Russian Doll MCTS (csrc/tree_kernel.cu) through the C ABI:
exact agreement with the host restatement (tests/forest_restatement.py) on tree statistics, leaves and repetition
flags; independence of a tree from its batch; legality of every self-play move against the oracle."""
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


def _net(cdq, oracle, scale=2.0, seed=42):
    p = oracle.init_params(seed, scale)
    net = cdq.ChessQNetwork()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return net.cuda()


def _recs(t):
    import chess_deep_q_b200 as cdq
    return t.cpu().numpy().view(np.uint64).reshape(-1).view(cdq.BOARD_DTYPE)


def _compare(cdq, forest, i, want, root):
    c = int(forest.root_children[i].item())
    assert c == len(want["root_children"]), (i, c, len(want["root_children"]))
    moves = forest.edge_move[i, 0, :c].cpu().numpy().view(np.uint32)
    for k in range(c):
        assert cdq.mcts._decode_move(moves[k]) == cdq.mcts.move_from_records(root, want["root_children"][k]), (i, k)
    assert forest.edge_visits[i, 0, :c].cpu().tolist() == want["N"], i
    assert np.array_equal(forest.edge_value[i, 0, :c].cpu().numpy(), np.array(want["Q"])), i      # fp64, same operations
    assert int(forest.node_count[i].item()) == want["node_count"], i
    assert int(forest.positions_evaluated[i].item()) == want["evaluated"], i
    n_last = len(want["leaves"])
    assert _recs(forest.leaf_boards[i, :n_last]).tobytes() == want["leaves"].tobytes(), i          # incl. the rep3 bits
    assert forest.leaf_depth[i, :n_last].cpu().tolist() == want["leaf_depth"], i
    assert np.array_equal(forest.leaf_values[i, :n_last].cpu().numpy(), want["leaf_values"]), i


@pytest.mark.parametrize("workers,widths,iterations", [(4, [6, 4, 3], 18), (8, [25, 15, 10, 7, 5, 3, 2], 24), (3, [32, 2], 7)])
def test_forest_search_equals_the_host_restatement(cdq, oracle, workers, widths, iterations):
    from forest_restatement import Restatement
    net = _net(cdq, oracle)
    roots = np.concatenate([cdq.mcts._startpos(), oracle.playouts(40, seed=77, max_plies=120)[[3, 11, 17, 29, 35]]])
    seeds = [1000 + 7 * i for i in range(len(roots))]
    hmc = [0, 3, 0, 12, 149, 40]                       # 149: one quiet move away from the 75-move rule
    ply = [0, 20, 31, 64, 199, 90]
    forest = cdq.mcts.DeviceForest(len(roots), workers, widths, iterations)
    forest.set_games(roots, seeds, halfmove_clock=hmc, ply=ply)
    forest.search(net, iterations, exploration_weight=1.3, training_progress=0.25, max_moves=200)
    torch.cuda.synchronize()
    rs = Restatement(cdq, net, workers, widths, exploration_weight=1.3, training_progress=0.25, max_moves=200)
    for i in range(len(roots)):
        want = rs.search(roots[i], seeds[i], iterations, hmc=hmc[i], ply=ply[i])
        _compare(cdq, forest, i, want, roots[i])
    # the same trees searched alone: a tree does not depend on its batch
    for i in (1, 4):
        single = cdq.mcts.DeviceForest(1, workers, widths, iterations)
        single.set_games(roots[i:i + 1], seeds[i:i + 1], halfmove_clock=hmc[i:i + 1], ply=ply[i:i + 1])
        single.search(net, iterations, exploration_weight=1.3, training_progress=0.25, max_moves=200)
        c = int(forest.root_children[i].item())
        assert torch.equal(single.edge_visits[0, 0, :c], forest.edge_visits[i, 0, :c])
        assert torch.equal(single.edge_value[0, 0, :c], forest.edge_value[i, 0, :c])
        assert int(single.best_move[0].item()) == int(forest.best_move[i].item())
        single.close()
    forest.close()


def test_forest_repetition_history_reaches_the_leaves(cdq, oracle):
    """A game that has shuffled knights (start position twice on the board, the current position for the second time):
    simulations that shuffle back create a threefold repetition, which must arrive at the leaf evaluation as the
    record's rep3 bit (mcts.py:209 -> 0.0) -- identical to the restatement, which walks the same key history."""
    from forest_restatement import Restatement
    net = _net(cdq, oracle)
    start = cdq.mcts._startpos()

    def play(rec, frm, to):
        kids = Restatement(cdq, net, 1, [1]).successors(rec)[0]
        for k in kids:
            if cdq.mcts.move_from_records(rec, k)[:2] == (frm, to):
                return k.copy()
        raise AssertionError("move not found")
    line = [start[0]]
    for frm, to in [(6, 21), (62, 45), (21, 6), (45, 62), (6, 21), (62, 45), (21, 6)]:      # Nf3 Nf6 Ng1 Ng8 Nf3 Nf6 Ng1
        line.append(play(line[-1], frm, to))
    root, history = line[-1:], np.array(line[:-1], dtype=cdq.BOARD_DTYPE)
    widths, workers, iterations = [25, 25, 25], 8, 64
    forest = cdq.mcts.DeviceForest(1, workers, widths, iterations)
    forest.set_games(np.array(root, dtype=cdq.BOARD_DTYPE), [5], halfmove_clock=[7], ply=[7], history=[history])
    forest.search(net, iterations)
    torch.cuda.synchronize()
    assert int(forest.game_history_len[0].item()) == 7
    rs = Restatement(cdq, net, workers, widths)
    want = rs.search(root[0], 5, iterations, hmc=7, ply=7, history=list(history))
    _compare(cdq, forest, 0, want, root[0])
    # Ng8 from the root repeats the start position for the third time: that child, when visited as a leaf, scored 0.0
    kids = want["root_children"]
    k = [i for i, c in enumerate(kids) if cdq.mcts.move_from_records(root[0], c)[:2] == (45, 62)]
    assert k and want["N"][k[0]] >= 1
    forest.close()


def test_self_play_on_the_device_is_legal_reproducible_and_trains(cdq, oracle):
    """self_play_games: every move played is a legal successor according to the oracle, a game is the same alone and in
    a batch, transitions are stored and the network trains once per stored transition (chess_ai.py:187-192)."""
    net = _net(cdq, oracle)
    kw = dict(max_moves=14, iterations=16, samples_per_level=[8, 5, 3], num_workers=4)
    games = cdq.self_play_games(net, 5, seeds=[7, 8, 9, 10, 11], **kw)
    for g in games:
        pos = g["positions"]
        assert len(pos) == len(g["moves"]) + 1 == len(g["scores"]) + 1 and len(pos) > 4
        for a, b, mv in zip(pos[:-1], pos[1:], g["moves"]):
            succ = {oracle.push(a.reshape(1), m)[0].tobytes() for m in oracle.legal_moves(a.reshape(1))}
            clean = b.copy()
            clean["meta"] &= ~np.uint64(1 << 16)
            assert clean.tobytes() in succ, "illegal move"
            assert cdq.mcts.move_from_records(a, b) == mv
        _, sc, _ = oracle.evaluate(pos[1:])
        assert np.array_equal(np.array(g["scores"]), sc)
    alone = cdq.self_play_games(net, 1, seeds=[9], **kw)[0]
    assert alone["positions"].tobytes() == games[2]["positions"].tobytes()
    # the pipelined loop (which games go on is decided on the device, a ply is booked while the next search runs) and the
    # synchronous one (taken when a callback or an agent needs every ply's result first) play the same games
    seen = []
    sync = cdq.self_play_games(net, 5, seeds=[7, 8, 9, 10, 11], on_ply=lambda ply, active: seen.append(int(active.sum())), **kw)
    assert len(seen) >= 4 and seen[0] <= 5
    for a, b in zip(games, sync):
        assert a["positions"].tobytes() == b["positions"].tobytes() and a["moves"] == b["moves"]
        assert a["scores"] == b["scores"] and a["rewards"] == b["rewards"] and a["result_flags"] == b["result_flags"]
    # games that end at different plies, up to the move limit: a short limit, then a position one move from mate
    short = cdq.self_play_games(net, 3, seeds=[1, 2, 3], max_moves=1, iterations=16, samples_per_level=[8, 5, 3], num_workers=4)
    assert [len(g["moves"]) for g in short] == [1, 1, 1]
    # without a network: the heuristic leaf (mcts.py:226-229)
    h = cdq.self_play_games(None, 2, seeds=[1, 2], **kw)
    assert all(len(g["moves"]) > 4 for g in h)
    # with an agent: transitions + one train() per transition once the memory holds a batch
    agent = cdq.DQNAgent(batch_size=16)
    out = cdq.self_play_games(agent.q_network, 4, seeds=[3, 4, 5, 6], agent=agent, **kw)
    stored = sum(len(g["moves"]) for g in out)
    assert len(agent.memory) == stored
    n_losses = sum(len(g["losses"]) for g in out)
    assert n_losses == stored - 15 and all(np.isfinite(x) for g in out for x in g["losses"])
    mv = agent.select_move(cdq.mcts._startpos(), training_progress=0.0, current_move=0)
    assert mv is not None and len(mv.uci()) in (4, 5)


def test_optimized_chess_ai_trains_through_self_play(cdq, oracle, tmp_path):
    """chess_ai.OptimizedChessAI: games of self-play in lockstep on the device forest with the reference's annealing
    (neural_network.py:150-163), transitions stored, one training step per ply once the memory holds a batch, the
    target network synchronised every five games (chess_ai.py:117-250, 346-413), then a checkpoint."""
    from chess_deep_q_b200.chess_ai import OptimizedChessAI
    torch.manual_seed(1)
    ai = OptimizedChessAI(training_games=6)
    ai.dqn_agent.batch_size = 8
    ai.train(parallel_games=3, max_moves=7)
    assert len(ai.game_history) == 6 and all(g["move_count"] == len(g["moves"]) <= 7 for g in ai.game_history)
    assert all(len(m) in (4, 5) for g in ai.game_history for m in g["moves"])
    plies = sum(g["move_count"] for g in ai.game_history)
    assert len(ai.dqn_agent.memory) == plies and len(ai.loss_history) == plies - 7
    assert all(np.isfinite(x) for x in ai.loss_history) and len(ai.evaluation_history) == 6
    assert ai.game_history[0]["result"] in ("*", "1-0", "0-1", "1/2-1/2")
    # the target network was synchronised after the fifth game (and not after the sixth)
    ai.save_model(str(tmp_path / "m.pth"))
    ck = torch.load(str(tmp_path / "m.pth"), map_location="cpu", weights_only=False)
    assert ck["total_games_trained"] == 6 and ck["training_games"] == 6 and len(ck["move_count_history"]) == 6


def test_maximum_move_count_and_widest_configuration(cdq, oracle):
    """Edge cases of the expansion: the position with the most legal moves known (218) through cdq_expand_weighted
    (all of them, with categorize_moves' weights), cdq_sample_children and a forest search at the limits the ABI
    states (width 32, depth 8, 32 workers), against the oracle and the host restatement."""
    from forest_restatement import Restatement
    L = cdq._lib.lib()
    root = oracle.from_fen("R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - - 0 1")
    assert len(oracle.legal_moves(root)) == 218
    d = cdq.board_utils.to_device(root)
    kids = torch.empty((1, 224, 9), dtype=torch.int64, device="cuda")
    counts = torch.empty(1, dtype=torch.int32, device="cuda")
    wts = torch.empty((1, 224), dtype=torch.uint8, device="cuda")
    st = cdq._lib.stream_ptr()
    cdq._lib.check(L.cdq_expand_weighted(d.data_ptr(), 1, 224, kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), st))
    assert int(counts.item()) == 218
    got = {k.tobytes(): int(w) for k, w in zip(_recs(kids[0, :218]), wts[0, :218].cpu().numpy())}
    mv, w = oracle.categorize(root)
    want = {oracle.push(root, m)[0].tobytes(): int(x) for m, x in zip(mv, w)}
    assert got == want
    # weighted sampling over all 218 children: 32 distinct ones, and exactly the host restatement's choice
    seeds = torch.tensor([12345], dtype=torch.int64, device="cuda")
    sel = torch.empty((1, 32, 9), dtype=torch.int64, device="cuda")
    selc = torch.empty(1, dtype=torch.int32, device="cuda")
    cdq._lib.check(L.cdq_sample_children(kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), 1, 224, 32, seeds.data_ptr(),
                                         sel.data_ptr(), selc.data_ptr(), st))
    assert int(selc.item()) == 32 and len({k.tobytes() for k in _recs(sel[0])}) == 32
    net = _net(cdq, oracle)
    widths, workers, iterations = [32] * 8, 32, 40
    forest = cdq.mcts.DeviceForest(2, workers, widths, iterations)
    roots = np.concatenate([root, cdq.mcts._startpos()])
    forest.set_games(roots, [3, 4], ply=[60, 0])
    forest.search(net, iterations)
    torch.cuda.synchronize()
    rs = Restatement(cdq, net, workers, widths)
    for i in range(2):
        _compare(cdq, forest, i, rs.search(roots[i], [3, 4][i], iterations, ply=[60, 0][i]), roots[i])
    assert int(forest.root_children[0].item()) == 32
    forest.close()


def test_reference_call_surface_with_a_chess_board(cdq, oracle):
    """The import swap of INTEGRATION.md with a python-chess style Board (here the restated `chess` API of oracle/): the
    reference's own calls -- ParallelRussianDollMCTS(board, ...).search(neural_net=, device=, ...) (neural_network.py:166-181)
    and DQNAgent.select_move(board, progress, is_training, current_move, max_moves) (chess_ai.py:51-58) -- return a
    chess.Move that is legal on that board and can be pushed."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "pychess_shim"))
    import chess
    board = chess.Board()
    for uci in ("e2e4", "e7e5", "g1f3", "b8c6"):
        board.push([m for m in board.legal_moves if m.uci() == uci][0])
    net = _net(cdq, oracle)
    tree = cdq.ParallelRussianDollMCTS(board, iterations=24, exploration_weight=1.6, samples_per_level=[25, 15, 10, 7, 5, 3, 2], num_workers=7)
    move = tree.search(neural_net=net, device=torch.device("cuda"), training_progress=0.0, current_move=4, max_moves=200)
    assert isinstance(move, chess.Move) and move.uci() in {m.uci() for m in board.legal_moves}
    board.push(move)
    agent = cdq.DQNAgent()
    mv = agent.select_move(board, 0.0, True, 5, 200)
    assert isinstance(mv, chess.Move) and mv.uci() in {m.uci() for m in board.legal_moves}
    board.push(mv)
    assert agent.get_q_value(board) == pytest.approx(float(agent.q_network.forward_packed(cdq.pack_board(board)).item()))
    assert isinstance(cdq.fast_evaluate_position(board), float)
