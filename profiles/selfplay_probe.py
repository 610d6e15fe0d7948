"""BASELINE configs[2]: Russian Doll MCTS self-play with batched GPU leaf evaluation -- wall-clock
throughput of the Python driver for 1 game alone and for G games in lockstep (search_many).
    python profiles/selfplay_probe.py [G] [max_moves] [iterations]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 64
max_moves = int(sys.argv[2]) if len(sys.argv) > 2 else 20
iterations = int(sys.argv[3]) if len(sys.argv) > 3 else 48
torch.manual_seed(42)
net = cdq.ChessQNetwork().cuda()
cdq.self_play_game(net, max_moves=2, iterations=8)          # warm-up
for games in (1, G):
    L = cdq._lib.lib()
    l0 = L.cdq_launch_count()
    t0 = time.perf_counter()
    out = cdq.self_play_games(net, n_games=games, max_moves=max_moves, iterations=iterations)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    plies = sum(len(p) - 1 for p, _ in out)
    leaves = plies * iterations
    print("%4d game(s): %6.2f s, %7.1f plies/s, %9.0f leaf evaluations/s, %d kernel launches"
          % (games, dt, plies / dt, leaves / dt, L.cdq_launch_count() - l0))
