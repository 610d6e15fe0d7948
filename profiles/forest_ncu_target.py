"""ncu target: two searches of 1024 lockstep games on the device forest (profiles/README.md, round 2)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.mcts import annealed_search_parameters  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iterations, ew, widths = annealed_search_parameters(0.0)
net = cdq.ChessQNetwork().cuda()
forest = cdq.mcts.DeviceForest(n, 8, widths, iterations)
boards = torch.empty((n, 9), dtype=torch.int64, device="cuda")
cdq._lib.check(cdq._lib.lib().cdq_random_playouts(boards.data_ptr(), n, 42, 60, cdq._lib.stream_ptr()))     # mid-game roots
forest.set_games(boards, list(range(n)))
for _ in range(2):
    forest.search(net, iterations, ew)
torch.cuda.synchronize()
print("ok", int(forest.node_count.sum().item()))
