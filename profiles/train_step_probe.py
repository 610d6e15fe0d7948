import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import chess_deep_q_b200 as cdq
from chess_deep_q_b200.train import FusedAdam, dqn_train_step
n = 4096
boards = torch.empty((n + 1, 9), dtype=torch.int64, device='cuda')
L = cdq._lib.lib()
cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), n + 1, 42, 199, cdq._lib.stream_ptr()))
net = cdq.ChessQNetwork().cuda(); tgt = cdq.ChessQNetwork().cuda()
opt = FusedAdam(net)
s, ns = boards[:n].contiguous(), boards[1:].contiguous()
sc, fl = cdq.evaluate_batch(ns)
rew = (0.01 * torch.tanh(sc / 10)).float(); dn = ((fl & 7) != 0).float()
for _ in range(3):
    dqn_train_step(net, tgt, opt, s, ns, rew, dn)
torch.cuda.synchronize()
