"""In-graph timeline of one search + advance of N lockstep games on the device forest (torch.profiler / CUPTI): which kernels
a ply's device time is made of, and the ply's wall time in self_play_games."""
import os
import sys
import time
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.mcts import annealed_search_parameters  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iterations, ew, widths = annealed_search_parameters(0.0)
net = cdq.ChessQNetwork().cuda()
forest = cdq.mcts.DeviceForest(n, 8, widths, iterations)
boards = torch.empty((n, 9), dtype=torch.int64, device="cuda")
cdq._lib.check(cdq._lib.lib().cdq_random_playouts(boards.data_ptr(), n, 42, 40, cdq._lib.stream_ptr()))
forest.set_games(boards, list(range(n)))
for _ in range(3):
    forest.search_and_advance(net, iterations, ew)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    forest.search_and_advance(net, iterations, ew)
e1.record()
torch.cuda.synchronize()
print("# %d games: %.3f ms per search + advance (CUDA events, graph replay)" % (n, e0.elapsed_time(e1) / 5))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    forest.search_and_advance(net, iterations, ew)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
for e in evs:
    k = e.name.split("(")[0][-48:]
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
span = evs[-1].time_range.end - evs[0].time_range.start
print("# span %.1f us, kernels:" % span)
for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("  %-50s x%-3d %9.1f us" % (k, c, us))
