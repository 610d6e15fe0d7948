"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI of
libcdq.so via the Python host layer and is compared with the CPU oracle and with the golden
vectors produced by the reference's own code (tests/golden/)."""
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


def _np(t):
    return t.detach().cpu().numpy()


def _eval_gpu(cdq, boards, ignore_checkmate=False):
    score, flags, terms = cdq.evaluate_batch(boards, ignore_checkmate, return_terms=True)
    return _np(terms), _np(score), _np(flags).view(np.uint32)


# ------------------------------------------------------------------------------ K1: evaluation
def test_eval_matches_reference_golden(cdq, eval_golden):
    g = eval_golden
    terms, score, flags = _eval_gpu(cdq, g["boards"])
    assert np.array_equal(score, g["score"])                      # fp64, bit-exact
    _, score_ic, _ = _eval_gpu(cdq, g["boards"], ignore_checkmate=True)
    assert np.array_equal(score_ic, g["score_ignore_checkmate"])
    assert np.array_equal(terms[:, 23], g["terminal"])
    has = g["has_terms"]
    assert np.array_equal(terms[has][:, :23], g["terms"][has][:, :23])
    assert not (flags & 0x80000000).any()


@pytest.mark.parametrize("seed,n", [(7, 100_000), (1234, 50_000)])
def test_eval_matches_oracle_on_random_playouts(cdq, oracle, seed, n):
    boards = oracle.playouts(n, seed=seed, max_plies=199)
    t_ref, s_ref, f_ref = oracle.evaluate(boards, nthreads=0)
    terms, score, flags = _eval_gpu(cdq, boards)
    bad = np.nonzero((terms != t_ref).any(1))[0]
    assert bad.size == 0, "first mismatch %d: gpu %s oracle %s" % (bad[0], terms[bad[0]], t_ref[bad[0]])
    assert np.array_equal(score, s_ref)
    assert np.array_equal(flags & 0xFFFF, f_ref & 0xFFFF)
    turn = (boards["meta"] & 1).astype(bool)
    assert np.array_equal((flags & 0x10000) != 0, turn)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 127, 128, 129, 1000])
def test_eval_ragged_sizes_and_unaligned_views(cdq, oracle, n):
    boards = oracle.playouts(max(n, 1) + 3, seed=99)[: n + 3]
    t_ref, s_ref, _ = oracle.evaluate(boards)
    d = cdq.board_utils.to_device(boards)
    # a view starting at record 1 is only 8-byte aligned: exercises the non-16B load path
    for off in (0, 1):
        sub = d[off:off + n]
        score, flags, terms = cdq.evaluate_batch(sub, return_terms=True)
        assert np.array_equal(_np(terms), t_ref[off:off + n])
        assert np.array_equal(_np(score), s_ref[off:off + n])


def test_planes_match_board_to_tensor(cdq, oracle, eval_golden):
    g = eval_golden
    planes = _np(cdq.boards_to_tensor_batch(g["boards"]))
    assert planes.shape == (len(g["boards"]), 12, 8, 8)
    assert np.array_equal(planes, oracle.encode_planes(g["boards"]))
    bits = (planes.reshape(len(planes), 12, 64) > 0.5).astype(np.uint64)
    packed = (bits << np.arange(64, dtype=np.uint64)).sum(2, dtype=np.uint64)
    assert np.array_equal(packed, g["planes_packed"])
    one = _np(cdq.board_to_tensor(g["boards"][:1]))
    assert np.array_equal(one, planes[0])
    # round trip planes -> packed bitboards
    back = _np(cdq.board_utils.planes_to_packed(torch.from_numpy(planes).cuda())).view(np.uint64)
    assert np.array_equal(back[:, :8], g["boards"]["bb"])


def test_bad_board_is_flagged(cdq, oracle):
    b = oracle.startpos()
    b["bb"][0][5] |= np.uint64(1 << 30)   # second white king, not in the occupancy
    _, _, flags = _eval_gpu(cdq, b)
    assert flags[0] & 0x80000000


# ------------------------------------------------------------------------------ device move generator
def _dev_playouts(cdq, n, seed=42, max_plies=199):
    d = torch.empty((n, 9), dtype=torch.int64, device="cuda")
    cdq._lib.check(cdq._lib.lib().cdq_random_playouts(d.data_ptr(), n, seed, max_plies, cdq._lib.stream_ptr()))
    return d


def test_device_playouts_are_valid_positions(cdq, oracle):
    n = 100_000
    d = _dev_playouts(cdq, n)
    boards = _np(d).view(np.uint64).reshape(-1).view(oracle.BOARD_DTYPE)
    t_ref, s_ref, f_ref = oracle.evaluate(boards, nthreads=0)
    terms, score, flags = _eval_gpu(cdq, d)
    assert np.array_equal(terms, t_ref) and np.array_equal(score, s_ref)
    assert not (flags & 0x80000000).any()
    # the side that just moved is never left in check (positions are legal), kings present
    flipped = boards.copy()
    flipped["meta"] ^= np.uint64(1)
    _, _, ff = oracle.evaluate(flipped, nthreads=0)
    assert not (ff & 0x8).any()
    # same recipe as the oracle's generator -> same population statistics
    ref = oracle.playouts(n, seed=42)
    tr, sr, fr = oracle.evaluate(ref, nthreads=0)
    assert abs(terms[:, 2].mean() - tr[:, 2].mean()) < 0.5
    assert abs(((flags & 7) == 1).mean() - ((fr & 7) == 1).mean()) < 0.01
    pc = lambda b: np.array([bin(int(x)).count("1") for x in (b["bb"][:2000, 6] | b["bb"][:2000, 7])]).mean()
    assert abs(pc(boards) - pc(ref)) < 0.5
    # deterministic
    assert torch.equal(d, _dev_playouts(cdq, n))


def test_expand_matches_oracle_successors(cdq, oracle, eval_golden):
    boards = np.concatenate([eval_golden["boards"][:200], oracle.playouts(800, seed=5)])
    d = cdq.board_utils.to_device(boards)
    n, maxc = len(boards), 128
    children = torch.zeros((n, maxc, 9), dtype=torch.int64, device="cuda")
    counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    L = cdq._lib.lib()
    cdq._lib.check(L.cdq_expand(d.data_ptr(), n, maxc, children.data_ptr(), counts.data_ptr(), cdq._lib.stream_ptr()))
    ch = _np(children).view(np.uint64)
    cnt = _np(counts)
    for i in range(n):
        moves = oracle.legal_moves(boards[i:i + 1])
        assert cnt[i] == len(moves), i
        want = sorted(oracle.push(boards[i:i + 1], m)[0].tobytes() for m in moves)
        got = sorted(ch[i, k].tobytes() for k in range(cnt[i]))
        assert want == got, i


# ------------------------------------------------------------------------------ K2: network
def _make_net(cdq, oracle, scale, seed=42):
    net = cdq.ChessQNetwork()
    p = oracle.init_params(seed, scale)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    return net.cuda(), p


def _unpack_act2(buf, n):
    tiles = (n + 127) // 128
    a = buf[: tiles * 64 * 8192].view(np.float16).reshape(tiles, 64, 8, 16, 8, 8)  # [T][sq][kc][rg][r][k]
    a = a.transpose(0, 3, 4, 1, 2, 5).reshape(tiles * 128, 64, 64)                   # [pos][sq][c]
    return a[:n]


@pytest.mark.parametrize("scale", [1.0, 3.0])
def test_conv_stack_matches_oracle(cdq, oracle, net_golden, scale):
    boards = net_golden["boards"][:500]
    n = len(boards)
    net, p = _make_net(cdq, oracle, scale)
    _, a1_ref, a2_ref, _ = oracle.net_forward(p, oracle.encode_planes(boards), return_acts=True)
    d = cdq.board_utils.to_device(boards)
    act1 = torch.zeros((n, 64, 32), dtype=torch.float16, device="cuda")
    tiles = (n + 127) // 128
    act2 = torch.zeros(tiles * 64 * 8192, dtype=torch.float16, device="cuda")
    L = cdq._lib.lib()
    cdq._lib.check(L.cdq_debug_conv(net.net_handle(), d.data_ptr(), n, act1.data_ptr(), act2.data_ptr(),
                                    cdq._lib.stream_ptr()))
    torch.cuda.synchronize()
    a1 = _np(act1).astype(np.float32).transpose(0, 2, 1).reshape(n, 32, 8, 8)
    tol1 = 2e-3 * max(1.0, np.abs(a1_ref).max())
    assert np.abs(a1 - a1_ref).max() < tol1
    a2 = _unpack_act2(_np(act2).view(np.uint16).view(np.float16), n).astype(np.float32)
    a2 = a2.transpose(0, 2, 1).reshape(n, 64, 8, 8)
    tol2 = 4e-3 * max(1.0, np.abs(a2_ref).max())
    assert np.abs(a2 - a2_ref).max() < tol2


@pytest.mark.parametrize("tag,scale", [("s1", 1.0), ("s2", 2.0), ("s3", 3.0)])
def test_value_matches_reference_class(cdq, oracle, net_golden, tag, scale):
    """|value - reference fp32 ChessQNetwork| <= 1e-3 (north_star tolerance), fp16 operands."""
    boards = net_golden["boards"]
    net, _ = _make_net(cdq, oracle, scale)
    with torch.no_grad():
        v = _np(net.forward_packed(boards)).reshape(-1)
        v_planes = _np(net(cdq.boards_to_tensor_batch(boards))).reshape(-1)
    ref = net_golden["value_" + tag]
    assert np.abs(v - ref).max() <= 1e-3, np.abs(v - ref).max()
    assert np.array_equal(v, v_planes)      # planes input takes the same kernels


@pytest.mark.parametrize("n", [1, 15, 16, 17, 127, 128, 129, 300])
def test_value_ragged_batch_sizes(cdq, oracle, n):
    boards = oracle.playouts(n, seed=3)
    net, p = _make_net(cdq, oracle, 2.0)
    ref = oracle.net_forward(p, oracle.encode_planes(boards)).reshape(-1)
    with torch.no_grad():
        v = _np(net.forward_packed(boards)).reshape(-1)
    assert v.shape == (n,)
    assert np.abs(v - ref).max() <= 1e-3


def test_value_bitwise_stable_across_runs_and_offsets(cdq, oracle):
    """The stream conv kernel hands tiles between five warp roles through ~50 mbarriers; a missing
    fence or a slot reused too early would show up as run-to-run or placement-dependent bits.
    Twelve repetitions and evaluations of shifted sub-ranges (other tile / CTA assignment, other
    ring phases) must reproduce every value bit for bit.  Batches of at most 8 192 positions take
    the cluster split-K fc kernel (fc_split_kernel.cu), whose fp32 summation order differs from the
    persistent kernel's: with option "fc_split" = 0 they must reproduce the large batch's bits as well,
    with the split kernel they must agree to fp32 rounding and be bitwise stable among themselves."""
    n = 200_000
    d = _dev_playouts(cdq, n, seed=9)
    net, _ = _make_net(cdq, oracle, 3.0)
    small = [(1, 4095), (3, 16), (5, 1), (777, 8192)]
    with torch.no_grad():
        ref = net.forward_packed(d).reshape(-1).clone()
        for _ in range(12):
            assert torch.equal(net.forward_packed(d).reshape(-1), ref)
        for off, cnt in [(17, 70_001), (12_345, 132_608), (99_999, 100_001), (777, 8193)]:
            v = net.forward_packed(d[off:off + cnt].contiguous()).reshape(-1)
            assert torch.equal(v, ref[off:off + cnt]), (off, cnt)
        old = cdq._lib.set_option("fc_split", 0)
        try:
            for off, cnt in small:
                v = net.forward_packed(d[off:off + cnt].contiguous()).reshape(-1)
                assert torch.equal(v, ref[off:off + cnt]), (off, cnt)
        finally:
            cdq._lib.set_option("fc_split", old)
        split_ref = net.forward_packed(d[:8192].contiguous()).reshape(-1).clone()
        assert (split_ref - ref[:8192]).abs().max().item() <= 5e-6      # |value| up to 0.99 with these weights
        for _ in range(12):
            assert torch.equal(net.forward_packed(d[:8192].contiguous()).reshape(-1), split_ref)
        for off, cnt in [(1, 4095), (3, 16), (5, 1), (130, 8000), (4000, 4192)]:
            v = net.forward_packed(d[off:off + cnt].contiguous()).reshape(-1)
            assert torch.equal(v, split_ref[off:off + cnt]), (off, cnt)


def test_pipe_kernel_equals_two_kernel_path(cdq, oracle):
    """With option "pipe" = 1 (CDQ_PIPE=1) large batches run conv and fc as one persistent kernel with an L2-resident
    hand-over (pipe_kernel.cu); the default is the two stand-alone kernels.  Same tiles, same MMAs, same
    epilogues: leaf values, raw values, scores and flags must agree bit for bit, for sizes around
    the ring (64 tiles of 128 positions) and the role split."""
    net, _ = _make_net(cdq, oracle, 2.0)
    ev = cdq.LeafEvaluator(net)
    for n in (524_288, 540_001, 64 * 128 * 67 + 5):
        d = _dev_playouts(cdq, n, seed=n)
        old = cdq._lib.set_option("pipe", 0)
        try:
            a = ev.evaluate_device(d, want_raw=True)
            cdq._lib.set_option("pipe", 1)
            b = ev.evaluate_device(d, want_raw=True)
            c = ev.evaluate_device(d, want_raw=True)
        finally:
            cdq._lib.set_option("pipe", old)
        for k in ("leaf", "value", "score", "flags"):
            assert torch.equal(a[k], b[k]), (n, k)
            assert torch.equal(b[k], c[k]), (n, k)


def test_leaf_eval_without_a_network_is_the_heuristic_branch(cdq, oracle, eval_golden):
    """mcts.py:226-229: `neural_net is None` -> tanh(fast_evaluate_position(board) / 10) behind the terminal
    rules of :202-210.  Device fp64 tanh against libm's math.tanh: 1e-6 after rounding to the fp32 leaf
    array (SURVEY 8b's tolerance class); terminal values exact; host path == device path."""
    g = eval_golden
    ev = cdq.LeafEvaluator(None)
    out = ev.evaluate_device(g["boards"])
    leaf = _np(out["leaf"]).astype(np.float64)
    assert np.array_equal(_np(out["score"]), g["score"])
    want = oracle.leaf_values(g["boards"], g["terminal"], score=g["score"])
    term = g["terminal"] != 0
    assert np.array_equal(leaf[term], want[term])
    assert np.abs(leaf - want).max() <= 1e-6
    boards = oracle.playouts(5000, seed=4)
    _, sc, fl = oracle.evaluate(boards, nthreads=0)
    out = ev.evaluate_device(boards)
    assert np.abs(_np(out["leaf"]).astype(np.float64) - oracle.leaf_values(boards, fl, score=sc)).max() <= 1e-6
    host = ev.evaluate_host(boards)
    assert np.array_equal(host["leaf"], _np(out["leaf"])) and np.array_equal(host["score"], sc)
    assert cdq.evaluate_position(boards[:1]) == pytest.approx(float(oracle.leaf_values(boards[:1], fl[:1], score=sc[:1])[0]), abs=1e-6)
    with pytest.raises(cdq._lib.CdqError):
        ev.evaluate_device(boards, want_raw=True)


def test_host_contexts_are_independent_and_options_round_trip(cdq, oracle):
    """cdq_host_ctx_*: two explicit contexts used from two threads at once give the results of the
    default per-device context; cdq_set_option / cdq_get_option round-trip and reject unknown names."""
    import ctypes
    import threading
    L = cdq._lib.lib()
    net, _ = _make_net(cdq, oracle, 2.0)
    boards = oracle.playouts(70_000, seed=12)
    ref = cdq.LeafEvaluator(net).evaluate_host(boards)
    handle = net.net_handle()
    torch.cuda.synchronize()
    results, errors = [None, None], []

    def work(i):
        try:
            torch.cuda.set_device(0)
            ctx = ctypes.c_void_p()
            cdq._lib.check(L.cdq_host_ctx_create(ctypes.byref(ctx)))
            leaf = np.empty(len(boards), np.float32)
            score = np.empty(len(boards), np.float64)
            flags = np.empty(len(boards), np.uint32)
            for _ in range(3):
                cdq._lib.check(L.cdq_leaf_eval_host_ctx(ctx, handle, boards.ctypes.data, len(boards), leaf.ctypes.data,
                                                        score.ctypes.data, flags.ctypes.data))
            L.cdq_host_ctx_destroy(ctx)
            results[i] = (leaf, score, flags)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
    for leaf, score, flags in results:
        assert np.array_equal(leaf, ref["leaf"]) and np.array_equal(score, ref["score"]) and np.array_equal(flags, ref["flags"])
    old = cdq._lib.set_option("host_ramp", 0)
    assert old == 1 and cdq._lib.set_option("host_ramp", 1) == 0
    assert L.cdq_set_option(b"no_such_option", 1) == -1
    L.cdq_host_release()
    again = cdq.LeafEvaluator(net).evaluate_host(boards)      # the default context is rebuilt on demand
    assert np.array_equal(again["leaf"], ref["leaf"])


def test_repack_after_weight_update(cdq, oracle):
    boards = oracle.playouts(64, seed=11)
    net, p = _make_net(cdq, oracle, 1.0)
    with torch.no_grad():
        v1 = _np(net.forward_packed(boards))
        net.fc2.bias.add_(0.25)
        v2 = _np(net.forward_packed(boards))
    p2 = dict(p)
    p2["fc2.bias"] = p["fc2.bias"] + np.float32(0.25)
    ref2 = oracle.net_forward(p2, oracle.encode_planes(boards))
    assert np.abs(v2 - ref2).max() <= 1e-3 and np.abs(v2 - v1).max() > 0.05


# ------------------------------------------------------------------------------ leaf evaluation
def _leaf_reference(oracle, boards, p):
    terms, score, flags = oracle.evaluate(boards, nthreads=0)
    v = oracle.net_forward(p, oracle.encode_planes(boards)).reshape(-1)
    leaf = np.clip(v, -1.0, 1.0)
    white = (boards["meta"] & 1).astype(bool)
    mate = (flags & 0x10) != 0
    leaf[mate & white] = -1.0
    leaf[mate & ~white] = 1.0
    leaf[(flags & 0xE0) != 0] = 0.0
    return leaf, score, flags, v


def test_leaf_eval_matches_mcts_semantics(cdq, oracle, eval_golden):
    boards = np.concatenate([eval_golden["boards"], oracle.playouts(3000, seed=21)])
    net, p = _make_net(cdq, oracle, 2.0)
    leaf_ref, score_ref, flags_ref, v_ref = _leaf_reference(oracle, boards, p)
    ev = cdq.LeafEvaluator(net)
    out = ev.evaluate_device(boards, want_raw=True)
    leaf, score, flags = _np(out["leaf"]), _np(out["score"]), _np(out["flags"]).view(np.uint32)
    assert np.array_equal(score, score_ref)
    assert np.array_equal(flags & 0xFFFF, flags_ref & 0xFFFF)
    term = (flags_ref & 0xF0) != 0
    assert term.sum() > 50
    assert np.array_equal(leaf[term], leaf_ref[term])            # exact on terminal leaves
    assert np.abs(leaf - leaf_ref).max() <= 1e-3
    assert np.abs(_np(out["value"]) - v_ref).max() <= 1e-3
    # golden terminal values computed with the reference's own mcts.py:202-210 call order
    g = eval_golden
    lt = g["leaf_terminal"]
    m = ~np.isnan(lt)
    assert np.array_equal(leaf[: len(lt)][m], lt[m])
    # host path (H2D + D2H inside): same numbers
    host = ev.evaluate_host(boards)
    assert np.array_equal(host["leaf"], leaf) and np.array_equal(host["score"], score)
    assert np.array_equal(host["flags"], flags)
    # single-board drop-ins
    assert cdq.evaluate_position(boards[:1], net) == pytest.approx(float(leaf[0]), abs=0)
    assert cdq.fast_evaluate_position(boards[:1]) == score_ref[0]


def test_full_size_properties_1m(cdq, oracle):
    """BASELINE config[1] size (1M positions): size-independent properties, the handcrafted score against the oracle
    on every position, the network on a 20 000-position sample."""
    n = 1 << 20
    d = _dev_playouts(cdq, n, seed=42)
    net, p = _make_net(cdq, oracle, 2.0)
    ev = cdq.LeafEvaluator(net)
    out = ev.evaluate_device(d, want_raw=True)
    torch.cuda.synchronize()
    # (1) idempotence
    out2 = ev.evaluate_device(d, want_raw=True)
    for k in ("leaf", "score", "flags", "value"):
        assert torch.equal(out[k], out2[k]), k
    # (2) permutation equivariance: no cross-position interference, tile/group placement irrelevant
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    outp = ev.evaluate_device(d[perm].contiguous(), want_raw=True)
    for k in ("leaf", "score", "flags", "value"):
        assert torch.equal(out[k][perm], outp[k]), k
    # (3) sampled oracle comparison
    idx = np.random.default_rng(0).choice(n, 20000, replace=False)
    boards = _np(d).view(np.uint64).reshape(-1).view(oracle.BOARD_DTYPE)[idx]
    leaf_ref, score_ref, flags_ref, v_ref = _leaf_reference(oracle, boards, p)
    assert np.array_equal(_np(out["score"])[idx], score_ref)
    assert np.array_equal(_np(out["flags"]).view(np.uint32)[idx] & 0xFFFF, flags_ref & 0xFFFF)
    assert np.abs(_np(out["leaf"])[idx] - leaf_ref).max() <= 1e-3
    # (3b) the handcrafted half against the C oracle on ALL 1 048 576 positions: 24 integer terms, fp64 score, flags
    all_boards = _np(d).view(np.uint64).reshape(-1).view(oracle.BOARD_DTYPE)
    t_ref, s_ref, f_ref = oracle.evaluate(all_boards, nthreads=0)
    terms, score_all, flags_all = _eval_gpu(cdq, d)
    assert np.array_equal(terms, t_ref)
    assert np.array_equal(score_all, s_ref) and np.array_equal(_np(out["score"]), s_ref)
    assert np.array_equal(flags_all & 0xFFFF, f_ref & 0xFFFF)
    # (4) ranges
    assert float(out["leaf"].abs().max()) <= 1.0
    assert bool(torch.isfinite(out["score"]).all())
    # (5) the host-buffer path (cdq_leaf_eval_host: ramped, double-buffered chunks, copies inside) returns the same
    #     bits for the whole batch and for a ragged prefix that ends inside a ramp chunk
    h = _np(d).view(np.uint64).reshape(-1).view(oracle.BOARD_DTYPE)
    for m in (n, 100_003):
        host = ev.evaluate_host(h[:m])
        for k in ("leaf", "score"):
            assert np.array_equal(host[k], _np(out[k])[:m]), (m, k)
        assert np.array_equal(host["flags"], _np(out["flags"]).view(np.uint32)[:m]), m


def test_attack_and_movement_maps(cdq, oracle, eval_golden):
    """find_threatened_squares / find_guarded_squares / calculate_attack_range /
    calculate_movement_range (evaluation.py:368-428) from the eval kernel's bitboards."""
    boards = np.concatenate([eval_golden["boards"][:40], oracle.playouts(20000, seed=17)])
    ref = oracle.attack_maps(boards)
    got = _np(cdq.evaluation.attack_maps(boards)).view(np.uint64)
    assert np.array_equal(got, ref)
    b = boards[5:6]
    white = bool(int(b["meta"][0]) & 1)
    sq = lambda bits: {s for s in range(64) if (int(bits) >> s) & 1}   # noqa: E731
    assert cdq.evaluation.calculate_attack_range(b, True) == sq(ref[5, 0])
    assert cdq.evaluation.calculate_movement_range(b, False) == sq(ref[5, 3])
    assert cdq.evaluation.find_threatened_squares(b) == sq(ref[5, 1 if white else 0])
    assert cdq.evaluation.find_guarded_squares(b) == sq(ref[5, 0 if white else 1])


from randboards import random_boards as _random_boards  # noqa: E402


@pytest.mark.parametrize("seed", [1, 2])
def test_eval_matches_oracle_on_arbitrary_positions(cdq, oracle, seed):
    boards = _random_boards(60_000, seed).astype(oracle.BOARD_DTYPE)
    t_ref, s_ref, f_ref = oracle.evaluate(boards, nthreads=0)
    terms, score, flags = _eval_gpu(cdq, boards)
    bad = np.nonzero((terms != t_ref).any(1))[0]
    assert bad.size == 0, "first of %d mismatches, board %s meta %x: gpu %s oracle %s" % (
        bad.size, [hex(int(x)) for x in boards["bb"][bad[0]]], int(boards["meta"][bad[0]]), terms[bad[0]], t_ref[bad[0]])
    assert np.array_equal(score, s_ref)
    assert np.array_equal(flags & 0xFFFF, f_ref & 0xFFFF)
    assert np.array_equal(_np(cdq.evaluation.attack_maps(boards)).view(np.uint64), oracle.attack_maps(boards[:60_000]))


def test_eval_matches_reference_on_arbitrary_golden(cdq):
    """The reference's own evaluation.py on 400 arbitrary positions (tests/golden/arbitrary_golden.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "arbitrary_golden.npz"))
    terms, score, flags = _eval_gpu(cdq, g["boards"])
    assert np.array_equal(score, g["score"])
    has = g["has_terms"]
    assert np.array_equal(terms[has][:, :23], g["terms"][has][:, :23])
