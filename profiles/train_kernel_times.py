"""Warm per-kernel device times of one DQN replay step (batch 4 096): the step's kernels are launched
eagerly with a CUDA-event pair around each (cdq_profile_*), caches warm, no profiler attached.
    python profiles/train_kernel_times.py [steps]
The ncu launch list of the same step (train_step_probe.py) is cold-cache and serialised."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
for _ in range(3):
    step._body()
torch.cuda.synchronize()
L.cdq_profile_enable(1)
for _ in range(steps):
    step._body()
torch.cuda.synchronize()
L.cdq_profile_enable(0)
ms = (ctypes.c_double * 16)()
cnt = (ctypes.c_uint64 * 16)()
cdq._lib.check(L.cdq_profile_collect_ex(ms, cnt))
names = {1: "conv_stream (online + target)", 2: "fc_head (online + target)", 5: "pack_weights", 8: "head_bwd",
         9: "fc1_wgrad", 10: "fc1_dgrad", 11: "conv2_bwd", 12: "conv1_wgrad", 13: "optimizer"}
tot = 0.0
for k, name in names.items():
    if cnt[k]:
        per_step = ms[k] * 1e3 / steps
        tot += per_step
        print("%-32s %7.1f us per step (%d launches per step)" % (name, per_step, cnt[k] // steps))
print("%-32s %7.1f us per step (eager launches, event pairs; side-stream kernels overlap)" % ("sum", tot))
