"""Multi-rank worker of tests/test_gpu_train.py::test_multi_gpu_fused_dp_step_matches_single_gpu (launched with
torchrun on 2, 4 or 8 GPUs; also run by hand through gpurun, logs under profiles/).  Parity is measured at
GRADIENT level with lr = 0 (train.dp_parity_report), on an even and on an uneven global batch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import chess_deep_q_b200 as cdq  # noqa: E402
import oracle as O  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep, dp_parity_report, shard_range  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    p, pt = O.init_params(42, 2.0), O.init_params(43, 2.0)

    def net(params):
        m = cdq.ChessQNetwork()
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        return m.cuda()

    reports = []
    for n in (4096, 4099, 515):                    # BASELINE configs[3]; uneven shards; ragged tiles on every rank
        boards = O.playouts(n + 1, seed=5 + n)
        _, sc, fl = O.evaluate(boards[1:], nthreads=0)
        r = torch.from_numpy((0.01 * np.tanh(sc / 10.0)).astype(np.float32)).cuda()
        d = torch.from_numpy(((fl & 7) != 0).astype(np.float32)).cuda()
        s, ns = cdq.board_utils.to_device(boards[:n]), cdq.board_utils.to_device(boards[1:])
        rep = dp_parity_report(lambda: (net(p), net(pt)), s, ns, r, d)
        reports.append(rep)
        if rank == 0:
            print("dp_parity " + json.dumps(rep), flush=True)
    # sharded optimizer state gathers to the full torch.optim.Adam state (checkpoint format under a peer group)
    n = 512
    boards = O.playouts(n + 1, seed=5)
    _, sc, fl = O.evaluate(boards[1:], nthreads=0)
    r = torch.from_numpy((0.01 * np.tanh(sc / 10.0)).astype(np.float32)).cuda()
    d = torch.from_numpy(((fl & 7) != 0).astype(np.float32)).cuda()
    s, ns = cdq.board_utils.to_device(boards[:n]), cdq.board_utils.to_device(boards[1:])
    q_dp, t_dp = net(p), net(pt)
    opt_dp = FusedAdam(q_dp, peer_group=True)
    lo, hi = shard_range(n, rank, world)
    step_dp = GraphedDQNStep(q_dp, t_dp, opt_dp, batch=hi - lo, global_batch=n)
    step_dp.load(s[lo:hi], ns[lo:hi], r[lo:hi], d[lo:hi])
    q_1, t_1 = net(p), net(pt)
    opt_1 = FusedAdam(q_1)
    step_1 = GraphedDQNStep(q_1, t_1, opt_1, batch=n)
    step_1.load(s, ns, r, d)
    for _ in range(4):
        step_dp.run()
        step_1.run()
    torch.cuda.synchronize()
    osd = opt_dp.state_dict()
    assert len(osd["state"]) == 8 and int(float(osd["state"][0]["step"])) == 4
    m_full = torch.cat([osd["state"][i]["exp_avg"].reshape(-1) for i in range(8)])
    m_1 = torch.cat([opt_1.state_dict()["state"][i]["exp_avg"].reshape(-1) for i in range(8)])
    assert float((m_full - m_1).abs().max()) <= 1e-3 * float(m_1.abs().max()) + 1e-9
    dist.barrier()
    opt_dp.region.close()
    assert all(rep["ok"] for rep in reports), reports
    if rank == 0:
        print("dp_worker ok: world %d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
