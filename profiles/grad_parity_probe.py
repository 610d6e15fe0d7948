"""Per-tensor gradient error of the CUDA training step against (a) the fp16-operand restatement and (b) plain fp32
autograd (oracle.dqn_step_reference), as distributions: which tensors carry which error, relative to what."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import chess_deep_q_b200 as cdq  # noqa: E402
import oracle as O  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, PARAM_ORDER, dqn_train_step  # noqa: E402

O.build()
for n, scale in ((4096, 2.0), (203, 2.0), (17, 3.0)):
    boards = O.playouts(n + 1, seed=1000 + n)
    _, sc, fl = O.evaluate(boards[1:], nthreads=0)
    r = (0.01 * np.tanh(sc / 10.0)).astype(np.float32)
    d = ((fl & 7) != 0).astype(np.float32)
    p, pt = O.init_params(42, scale), O.init_params(43, scale)

    def net(params):
        m = cdq.ChessQNetwork()
        m.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        return m.cuda()
    q_net, t_net = net(p), net(pt)
    opt = FusedAdam(q_net, lr=0.0)
    to_dev = cdq.board_utils.to_device
    loss = float(dqn_train_step(q_net, t_net, opt, to_dev(boards[:n]), to_dev(boards[1:]), torch.from_numpy(r).cuda(),
                                torch.from_numpy(d).cuda(), 0.99).item())
    gv = {k: v.cpu().numpy() for k, v in opt.grad_views().items()}
    pl, npl = O.encode_planes(boards[:n]), O.encode_planes(boards[1:])
    l16, r16 = O.dqn_step_reference(p, pt, pl, npl, r, d, fp16_operands=True)
    l32, r32 = O.dqn_step_reference(p, pt, pl, npl, r, d)
    print("n=%d scale=%g loss gpu %.9g fp16-oracle %.9g fp32 %.9g" % (n, scale, loss, l16, l32))
    for k in PARAM_ORDER:
        g, a, b = gv[k].reshape(-1), r16[k].reshape(-1), r32[k].reshape(-1)
        m = np.abs(a).max()
        e16, e32, e1632 = np.abs(g - a) / m, np.abs(g - b) / m, np.abs(a - b) / m
        big = np.abs(a) > 1e-3 * m
        rel = np.abs(g - a)[big] / np.abs(a)[big]
        print("  %-13s max|g| %.3e  |gpu-fp16o|/max: max %.2e p99 %.2e  |gpu-fp32|/max: max %.2e  |fp16o-fp32|/max: max %.2e  rel(big) max %.2e p99 %.2e (at |ref|/max = %.2e)"
              % (k, m, e16.max(), np.quantile(e16, 0.99), e32.max(), e1632.max(), rel.max(), np.quantile(rel, 0.99),
                 np.abs(a)[big][rel.argmax()] / m))
