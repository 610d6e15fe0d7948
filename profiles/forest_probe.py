"""BASELINE configs[2] as a throughput line: Russian Doll MCTS self-play (widths 25/15/10/7/5/3/2 at training progress 0,
50 simulations per move, 8 simulations per wave) for N lockstep games on the device forest: leaf evaluations/s and
plies/s, search only and with transitions stored + one training step per transition (chess_ai.py:187-192)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.mcts import annealed_search_parameters  # noqa: E402

rng = np.random.default_rng(42)
net = cdq.ChessQNetwork()
with torch.no_grad():
    for name, p in net.named_parameters():
        fan_in = {"conv1": 108, "conv2": 288, "fc1": 4096, "fc2": 256}[name.split(".")[0]]
        a = rng.uniform(-1, 1, size=tuple(p.shape)).astype(np.float32) / np.sqrt(fan_in)
        p.copy_(torch.from_numpy(a * (2.0 if name.endswith("weight") else 1.0)))
net = net.cuda()
iterations, ew, widths = annealed_search_parameters(0.0)
L = cdq._lib.lib()
rows = []
for n_games in [int(x) for x in (sys.argv[1:] or ["1", "64", "1024", "4096"])]:
    plies = 12 if n_games >= 1024 else 24
    evaluated = []

    def on_ply(ply, active):
        evaluated.append(int(active.sum()))
    cdq.self_play_games(net, n_games, max_moves=2, iterations=iterations, samples_per_level=widths, exploration_weight=ew)   # warm-up
    l0 = L.cdq_launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    games = cdq.self_play_games(net, n_games, max_moves=plies, iterations=iterations, samples_per_level=widths,
                                exploration_weight=ew, on_ply=on_ply)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    moves = sum(len(g["moves"]) for g in games)
    rows.append({"games": n_games, "plies_per_game": plies, "moves_played": moves, "seconds": sec,
                 "leaf_evals_per_s": moves * iterations / sec, "plies_per_s": moves / sec, "kernel_launches": int(L.cdq_launch_count() - l0),
                 "simulations_per_move": iterations, "widths": widths, "workers": 8})
    print(json.dumps(rows[-1]), flush=True)
# with the replay memory and one training step per stored transition (batch 64 like the reference's default)
agent = cdq.DQNAgent(batch_size=64)
agent.q_network.load_state_dict(net.state_dict())
agent.update_target_network()
for n_games in (64, 1024):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    games = cdq.self_play_games(agent.q_network, n_games, max_moves=8, iterations=iterations, samples_per_level=widths,
                                exploration_weight=ew, agent=agent)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    moves = sum(len(g["moves"]) for g in games)
    steps = sum(len(g["losses"]) for g in games)
    print(json.dumps({"games": n_games, "with_training": True, "moves_played": moves, "train_steps": steps, "seconds": sec,
                      "leaf_evals_per_s": moves * iterations / sec, "plies_per_s": moves / sec}), flush=True)
# per-kernel-kind device time of one 1024-game ply (event pairs): tree level kernels vs the leaf evaluation
import ctypes
forest = cdq.mcts.DeviceForest(1024, 8, widths, iterations)
forest.set_games(np.concatenate([cdq.mcts._startpos()] * 1024), list(range(1024)))
for _ in range(2):
    forest.search(net, iterations, ew)
    forest.advance()
torch.cuda.synchronize()
L.cdq_profile_enable(1)
forest.search(net, iterations, ew)
torch.cuda.synchronize()
L.cdq_profile_enable(0)
kms, kcnt = (ctypes.c_double * 8)(), (ctypes.c_uint64 * 8)()
cdq._lib.check(L.cdq_profile_collect(kms, kcnt))
print(json.dumps({"one_search_1024_games_ms": {"eval_terms": kms[0], "conv": kms[1], "fc": kms[2], "tree_level": kms[4]},
                  "launches": {"eval_terms": kcnt[0], "conv": kcnt[1], "fc": kcnt[2], "tree_level": kcnt[4]}}))
