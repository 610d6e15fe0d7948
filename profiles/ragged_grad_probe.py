import sys, os, ctypes
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import chess_deep_q_b200 as cdq
import oracle as O
from chess_deep_q_b200 import _lib
from chess_deep_q_b200.train import FusedAdam, PARAM_ORDER, _weights_struct, _scratch
eg = np.load('/root/repo/tests/golden/eval_golden.npz'); tg = np.load('/root/repo/tests/golden/train_golden.npz')
boards = eg['boards'].view(O.BOARD_DTYPE).reshape(-1) if eg['boards'].dtype != O.BOARD_DTYPE else eg['boards']
idx = tg['idx']
for n in [1, 2, 3, 4, 5, 16, 17, 203, 256]:
    s, ns, r, d = boards[idx][:n], boards[idx + 1][:n], tg['reward'][:n], tg['done'][:n]
    p, pt = O.init_params(42, 2.0), O.init_params(43, 2.0)
    def mk(pp):
        net = cdq.ChessQNetwork(); net.load_state_dict({k: torch.from_numpy(v) for k, v in pp.items()}); return net.cuda()
    q_net, t_net = mk(p), mk(pt)
    opt = FusedAdam(q_net, lr=0.0)
    ref_loss, ref = O.dqn_step_reference(p, pt, O.encode_planes(s), O.encode_planes(ns), r, d)
    to_dev = cdq.board_utils.to_device
    sd, nsd = to_dev(s), to_dev(ns)
    rd, dd = torch.from_numpy(r).cuda(), torch.from_numpy(d).cuda()
    params = dict(q_net.named_parameters())
    weights = _weights_struct([params[k] for k in PARAM_ORDER])
    gv = opt.grad_views()
    grads = _weights_struct([gv[k] for k in PARAM_ORDER])
    ws, loss = _scratch.get(n, sd.device)
    opt.zero_grad(); loss.zero_()
    _lib.check(_lib.lib().cdq_train_fwd_bwd(q_net.net_handle(), t_net.net_handle(), ctypes.byref(weights), sd.data_ptr(), nsd.data_ptr(), rd.data_ptr(), dd.data_ptr(), n, 0.0, 0.99, 0, ctypes.byref(grads), loss.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    out = ["n=%d loss %.3e/%.3e" % (n, float(loss.item()), float(ref_loss))]
    for k in PARAM_ORDER:
        g = gv[k].cpu().numpy(); e = np.abs(g - ref[k]).max() / (np.abs(ref[k]).max() + 1e-30)
        out.append("%s %.1e" % (k.replace('.weight', '.w').replace('.bias', '.b'), e))
    print("  ".join(out))
