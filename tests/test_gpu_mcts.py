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


def _move_code(mv):
    frm, to, promo = mv
    return frm | (to << 6) | ((promo - 1 if promo else 0) << 12)      # the oracle's uint16 move encoding


def test_search_returns_legal_move_and_exact_leaf_values(cdq, oracle):
    net, p = _net(cdq, oracle)
    root = oracle.from_fen("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    tree = cdq.BatchedRussianDollMCTS(root, iterations=64, seed=1)
    child, move = tree.search(net)
    assert child[0].tobytes() in _successors(oracle, root)
    assert int(tree.root_visits.sum()) == 64 and tree.total_positions_evaluated == 64     # every simulation backed up through the root
    # the root's children are legal moves (sampled, at most width 21), all different
    legal = set(int(m) for m in oracle.legal_moves(root))
    codes = [_move_code(cdq.mcts._decode_move(m)) for m in tree.root_moves]
    assert len(codes) == 21 == len(set(codes)) and all(c in legal for c in codes)
    # the reported move transforms the root into the chosen child and is the most visited one
    assert _move_code(move) in legal
    assert oracle.push(root, _move_code(move))[0].tobytes() == child[0].tobytes()
    assert codes[int(np.argmax(tree.root_visits))] == _move_code(move)
    frm, to, _ = move
    assert cdq.move_uci(root[0], child[0])[:4] == "abcdefgh"[frm & 7] + str((frm >> 3) + 1) + "abcdefgh"[to & 7] + str((to >> 3) + 1)
    # the last wave's leaves: values follow mcts.py:196-232 (oracle: exact on terminals, 1e-3 on the network)
    forest = cdq.mcts.DeviceForest(1, 8, [21, 13, 8, 5, 3, 2, 1], 64)
    forest.set_games(root, [1])
    forest.search(net, 64)
    leaves = forest.leaf_boards[0].cpu().numpy().view(np.uint64).reshape(-1).view(cdq.BOARD_DTYPE)
    values = forest.leaf_values[0].cpu().numpy()
    assert (forest.leaf_depth[0].cpu().numpy() >= 1).all()
    _, _, flags = oracle.evaluate(leaves)
    v = oracle.net_forward(p, oracle.encode_planes(leaves)).reshape(-1)
    ref = np.clip(v, -1, 1)
    white = (leaves["meta"] & 1).astype(bool)
    mate = (flags & 0x10) != 0
    ref[mate & white], ref[mate & ~white] = -1.0, 1.0
    ref[(flags & 0xE0) != 0] = 0.0
    assert np.abs(values - ref).max() <= 1e-3
    assert np.array_equal(values[mate], ref[mate])
    forest.close()


def test_mate_in_one_backs_up_the_exact_terminal_value(cdq, oracle):
    """Ra8# for white: the mated leaf is worth exactly +1 (mcts.py:202-203, white-centric) and the
    running mean of that edge stays +1; with the reference's UCB (which divides the mean by N once
    more, mcts.py:61-62) that edge is visited at least as often as the average edge."""
    net, _ = _net(cdq, oracle, scale=1.0)
    root = oracle.from_fen("6k1/5ppp/8/8/8/8/8/R5K1 w - - 0 1")
    tree = cdq.BatchedRussianDollMCTS(root, iterations=200, seed=3)
    tree.search(net)
    mate = [i for i, m in enumerate(tree.root_moves) if cdq.mcts._decode_move(m)[:2] == (0, 56)]
    assert len(mate) == 1
    assert tree.root_values[mate[0]] == 1.0
    assert tree.root_visits[mate[0]] >= tree.root_visits.mean()
    # nodes are keyed like FENs (position + halfmove clock + move number): a line that returns to the root position
    # is another node, so the root sees exactly its 200 simulations
    assert int(tree.root_visits.sum()) == 200


def test_terminal_root(cdq, oracle):
    net, _ = _net(cdq, oracle)
    mated = oracle.from_fen("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 3")
    child, move = cdq.BatchedRussianDollMCTS(mated, iterations=8).search(net)
    assert child is None and move is None


def test_self_play_game_with_batched_leaf_evaluation(cdq, oracle):
    """BASELINE configs[2] as a parity case: Russian Doll widths 21/13/8/5/3/2/1, device-resident search."""
    net, _ = _net(cdq, oracle)
    positions, scores = cdq.self_play_game(net, max_moves=24, iterations=50, seed=42)
    assert len(positions) >= 2 and len(scores) == len(positions) - 1
    assert positions[0].tobytes() == oracle.startpos()[0].tobytes()
    for a, b in zip(positions[:-1], positions[1:]):
        assert b.tobytes() in _successors(oracle, np.array([a], dtype=oracle.BOARD_DTYPE))
    _, s_ref, _ = oracle.evaluate(positions[1:])
    assert np.array_equal(np.array(scores), s_ref)                     # reward-shaping scores are the bit-exact eval (chess_ai.py:161)
    again, _ = cdq.self_play_game(net, max_moves=24, iterations=50, seed=42)
    assert np.array_equal(again, positions)                           # seeded -> reproducible


def test_lockstep_games_equal_single_games(cdq, oracle):
    """self_play_games runs the searches of all games in one device forest (one kernel per level, one leaf
    evaluation per wave for all of them); every game must still be move for move the game that
    self_play_game plays alone with the same seed."""
    net, _ = _net(cdq, oracle)
    seeds = [42, 7, 1234]
    games = cdq.self_play_games(net, n_games=3, max_moves=10, iterations=24, seeds=seeds)
    for seed, g in zip(seeds, games):
        alone, alone_scores = cdq.self_play_game(net, max_moves=10, iterations=24, seed=seed)
        assert np.array_equal(g["positions"], alone), seed
        assert g["scores"] == alone_scores, seed


def test_expand_weights_match_categorize_moves(cdq, oracle, eval_golden):
    """the device move categoriser (sampling weights 10..1) against the oracle, which is
    itself pinned to the reference's categorize_moves (tests/test_oracle.py)."""
    boards = np.concatenate([eval_golden["boards"][:250], oracle.playouts(750, seed=9)])
    d = cdq.board_utils.to_device(boards)
    n, maxc = len(boards), 128
    kids = torch.zeros((n, maxc, 9), dtype=torch.int64, device="cuda")
    counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    wts = torch.zeros((n, maxc), dtype=torch.uint8, device="cuda")
    L = cdq._lib.lib()
    cdq._lib.check(L.cdq_expand_weighted(d.data_ptr(), n, maxc, kids.data_ptr(), counts.data_ptr(), wts.data_ptr(),
                                         cdq._lib.stream_ptr()))
    kh, ch, wh = kids.cpu().numpy().view(np.uint64), counts.cpu().numpy(), wts.cpu().numpy()
    for i in range(n):
        mv, w = oracle.categorize(boards[i:i + 1])
        want = {oracle.push(boards[i:i + 1], m)[0].tobytes(): int(x) for m, x in zip(mv, w)}
        # under-promotions to the same square give distinct children, so the map is one-to-one
        got = {kh[i, k].tobytes(): int(wh[i, k]) for k in range(ch[i])}
        assert got == want, i


def _mix64(z):
    m = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def test_device_child_sampler_matches_host_restatement(cdq, oracle, eval_golden):
    """cdq_sample_children (select_weighted_moves on the device) against a pure-Python
    integer restatement of its documented procedure, for several widths."""
    boards = np.concatenate([eval_golden["boards"][:200], oracle.playouts(600, seed=11)])
    d = cdq.board_utils.to_device(boards)
    n, maxc = len(boards), 128
    kids = torch.zeros((n, maxc, 9), dtype=torch.int64, device="cuda")
    counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    wts = torch.zeros((n, maxc), dtype=torch.uint8, device="cuda")
    L, st = cdq._lib.lib(), cdq._lib.stream_ptr()
    cdq._lib.check(L.cdq_expand_weighted(d.data_ptr(), n, maxc, kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), st))
    kh, ch, wh = kids.cpu().numpy().view(np.uint64), counts.cpu().numpy(), wts.cpu().numpy()
    rng = np.random.default_rng(5)
    seeds = rng.integers(0, 1 << 63, size=n, dtype=np.int64)
    d_seeds = torch.from_numpy(seeds).cuda()
    for width in (1, 5, 21, 40):
        sel = torch.zeros((n, width, 9), dtype=torch.int64, device="cuda")
        sc = torch.zeros(n, dtype=torch.int32, device="cuda")
        cdq._lib.check(L.cdq_sample_children(kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), n, maxc, width,
                                             d_seeds.data_ptr(), sel.data_ptr(), sc.data_ptr(), st))
        sh, sch = sel.cpu().numpy().view(np.uint64), sc.cpu().numpy()
        for i in range(n):
            c = int(ch[i])
            if c <= width:
                want = list(range(c))
            else:
                w, want = [int(x) for x in wh[i, :c]], []
                for k in range(width):
                    z = _mix64((int(seeds[i]) + (k + 1) * 0x9E3779B97F4A7C15) & ((1 << 64) - 1))
                    r, acc = (z * sum(w)) >> 64, 0
                    for j, wj in enumerate(w):
                        acc += wj
                        if acc > r:
                            break
                    want.append(j)
                    w[j] = 0
            assert int(sch[i]) == len(want), (width, i)
            assert np.array_equal(sh[i, :len(want)], kh[i, want]), (width, i)


def test_weighted_child_sampling_prefers_tactical_moves(cdq, oracle):
    """exd5 wins a queen with a pawn (weight 10, evaluation.py:253); with 5 weighted draws out of ~31
    moves it must be sampled far more often than the ~16 % a uniform sampler would give."""
    net, _ = _net(cdq, oracle)
    root = oracle.from_fen("rnb1kbnr/ppp1pppp/8/3q4/4P3/8/PPPP1PPP/RNBQKBNR w KQkq - 0 1")
    mv, w = oracle.categorize(root)
    best = [int(m) for m, x in zip(mv, w) if x == 10]
    assert len(best) == 1
    hits = 0
    for seed in range(40):
        tree = cdq.BatchedRussianDollMCTS(root, iterations=1, samples_per_level=[5], seed=seed)
        tree.search(net)
        codes = [_move_code(cdq.mcts._decode_move(m)) for m in tree.root_moves]
        assert len(codes) == 5 and len(set(codes)) == 5
        hits += best[0] in codes
    expected = 1.0 - np.prod([1.0 - 10.0 / (w.sum() - 2.0 * i) for i in range(5)])    # rough lower estimate
    assert hits >= 0.6 * 40 * expected and hits > 40 * 5 / len(mv) * 1.8
