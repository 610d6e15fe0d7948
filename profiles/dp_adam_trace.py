"""Where the fused reduce-scatter + Adam + all-gather kernel spends its time (a -DCDQ_DP_TRACE build stamps %globaltimer in
block 0): wait for the peers' gradients | reduce + Adam + stores | system fence | grid barrier | wait for the peers' stores.
    python chess-deep-q_b200/build.py -DCDQ_DP_TRACE --out=libcdq_trace.so
    CDQ_LIBCDQ=chess-deep-q_b200/libcdq_trace.so python -m torch.distributed.run --nproc-per-node W profiles/dp_adam_trace.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep, shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = 4096
boards = torch.empty((B + 1, 9), dtype=torch.int64, device="cuda")
cdq._lib.check(cdq._lib.lib().cdq_random_playouts(boards.data_ptr(), B + 1, 42, 199, cdq._lib.stream_ptr()))
torch.manual_seed(0)
net, tgt = cdq.ChessQNetwork().cuda(), cdq.ChessQNetwork().cuda()
opt = FusedAdam(net, peer_group=True)
lo, hi = shard_range(B, rank, world)
sc, fl = cdq.evaluate_batch(boards[lo + 1:hi + 1])
step = GraphedDQNStep(net, tgt, opt, batch=hi - lo, global_batch=B)
step.load(boards[lo:hi].contiguous(), boards[lo + 1:hi + 1].contiguous(), (0.01 * torch.tanh(sc / 10)).float(), ((fl & 7) != 0).float())
for _ in range(30):
    step.run()
torch.cuda.synchronize()
rows = []
for _ in range(20):
    dist.barrier()
    step.run()
    torch.cuda.synchronize()
    t = opt._grid_bar[16:28].view(torch.int64).cpu().tolist()
    rows.append([(t[i + 1] - t[i]) / 1e3 for i in range(5)])
med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(5)]
out = [None] * world
dist.all_gather_object(out, med)
if rank == 0:
    print("# world %d: median us per phase of dp_adam_kernel (block 0), per rank" % world)
    print("# rank  wait_grads  reduce+adam+stores  fence_system  grid_barrier  wait_stores   total")
    for r, m in enumerate(out):
        print("  %4d  %10.1f  %18.1f  %12.1f  %12.1f  %11.1f  %6.1f" % (r, m[0], m[1], m[2], m[3], m[4], sum(m)))
dist.barrier()
opt.region.close()
dist.destroy_process_group()
