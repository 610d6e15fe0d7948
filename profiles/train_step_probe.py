"""One DQN replay step (batch 4 096, eager launches) for `ncu --metrics gpu__time_duration.sum`:
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c 40 --csv --log-file out.csv \
        python profiles/train_step_probe.py
Prints nothing of value itself; the launch list is the product."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep  # noqa: E402

n = 4096
boards = torch.empty((n + 1, 9), dtype=torch.int64, device="cuda")
L = cdq._lib.lib()
cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), n + 1, 42, 199, cdq._lib.stream_ptr()))
net, tgt = cdq.ChessQNetwork().cuda(), cdq.ChessQNetwork().cuda()
opt = FusedAdam(net)
s, ns = boards[:n].contiguous(), boards[1:].contiguous()
sc, fl = cdq.evaluate_batch(ns)
rew = (0.01 * torch.tanh(sc / 10)).float()
dn = ((fl & 7) != 0).float()
step = GraphedDQNStep(net, tgt, opt, batch=n)
step.load(s, ns, rew, dn)
for _ in range(int(os.environ.get("STEPS", "4"))):
    step._body()          # eager launches: every kernel is a separate profiler record
torch.cuda.synchronize()
print("loss", float(step.loss.item()))
