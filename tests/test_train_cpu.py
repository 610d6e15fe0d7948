"""CPU suite for the training path: pins the oracle's restatement of DQNAgent.train against the
golden step recorded from the reference's own class, and covers the data-parallel host logic with
a world_size-2 gloo run."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(oracle, eval_golden, train_golden):
    idx = train_golden["idx"]
    boards = eval_golden["boards"]
    return (oracle.encode_planes(boards[idx]), oracle.encode_planes(boards[idx + 1]),
            train_golden["reward"], train_golden["done"])


@pytest.mark.parametrize("tag,scale", [("s1", 1.0), ("s3", 3.0)])
def test_oracle_train_step_matches_reference_agent(oracle, eval_golden, train_golden, tag, scale):
    s, ns, r, d = _batch(oracle, eval_golden, train_golden)
    p, pt = oracle.init_params(42, scale), oracle.init_params(43, scale)
    loss, grads = oracle.dqn_step_reference(p, pt, s, ns, r, d)
    assert loss == pytest.approx(float(train_golden["loss_" + tag]), rel=1e-5)
    after = oracle.adam_reference(p, [grads])
    for k, v in after.items():
        a = v.reshape(-1)
        got = a[:: max(1, a.size // 4096)][:4096]
        want = train_golden["w_%s_%s" % (tag, k)]
        # first Adam step moves every weight by lr*sign(g): allow sign flips only where g ~ 0
        bad = np.abs(got - want) > 1e-6
        assert bad.mean() < 2e-3, (k, bad.mean())
    # second step from the updated state
    loss2, grads2 = oracle.dqn_step_reference(after, pt, s, ns, r, d)
    assert loss2 == pytest.approx(float(train_golden["loss2_" + tag]), rel=2e-3)


def test_shard_range_partitions_the_batch():
    import chess_deep_q_b200.train as T
    for n in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            parts = [T.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_flat_layout_matches_reference_parameter_order():
    import chess_deep_q_b200 as cdq
    import chess_deep_q_b200.train as T
    net = cdq.ChessQNetwork()
    assert tuple(n for n, _ in net.named_parameters()) == T.PARAM_ORDER
    assert tuple(tuple(p.shape) for p in net.parameters()) == T.PARAM_SHAPES
    assert T.PARAM_OFFSETS[-1] + T.PARAM_SIZES[-1] == 1071073
    assert all(o % 4 == 0 for o in T.PARAM_OFFSETS)      # 16-byte aligned views


def _dp_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import chess_deep_q_b200.train as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 64
    boards = O.playouts(n + 1, seed=5)
    planes = O.encode_planes(boards)
    _, sc, fl = O.evaluate(boards[1:])
    reward = (0.01 * np.tanh(sc / 10.0)).astype(np.float32)
    done = ((fl & 7) != 0).astype(np.float32)
    p, pt = O.init_params(42, 2.0), O.init_params(43, 2.0)
    lo, hi = T.shard_range(n, rank, world)
    loss, grads = O.dqn_step_reference(p, pt, planes[lo:hi], planes[lo + 1:hi + 1], reward[lo:hi], done[lo:hi])
    flat = torch.cat([torch.from_numpy(grads[k]).reshape(-1) for k in T.PARAM_ORDER])
    scale = T.allreduce_mean_scale(flat)
    flat *= scale
    if rank == 0:
        full_loss, full = O.dqn_step_reference(p, pt, planes[:n], planes[1:n + 1], reward, done)
        ref = torch.cat([torch.from_numpy(full[k]).reshape(-1) for k in T.PARAM_ORDER])
        q.put((float((flat - ref).abs().max()), float(ref.abs().max()), scale))
    dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_equals_full_batch_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, mag, scale = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale == 0.5
    assert err <= 1e-6 * max(1.0, mag) + 1e-7
