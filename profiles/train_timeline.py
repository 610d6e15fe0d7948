"""Timeline of the GRAPHED DQN replay step (batch 4 096) from torch.profiler / CUPTI: start, duration and stream of
every kernel of one replay, relative to the step's first kernel.
    python profiles/train_timeline.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep  # noqa: E402

n = int(os.environ.get("BATCH", "4096"))
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
for _ in range(10):
    step.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        step.run()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name and "emset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)
starts = [i for i, e in enumerate(ev) if "pack_weights" in e.name or "conv_stream" in e.name and (i == 0 or "adam" in ev[i - 1].name)]
# last complete step: from the last kernel following an optimizer kernel
idx = [i for i, e in enumerate(ev) if "adam" in e.name]
lo, hi = idx[-2] + 1, idx[-1] + 1
t0 = ev[lo].time_range.start
for e in ev[lo:hi]:
    print("%8.1f us  +%6.1f us  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name.split("(")[0][:60]))
print("step span %.1f us" % (ev[hi - 1].time_range.end - t0))
