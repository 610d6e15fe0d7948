"""clock64 stamps of CTA 0 of conv2_bwd2_kernel (build with -DCB2_TRACE --out=libcdq_cb2trace.so, run with
CDQ_LIBCDQ=.../libcdq_cb2trace.so): where a unit's time goes -- producer / UMMA issuer / epilogue, per unit.
The step runs eagerly (kernels back to back, the fc1 weight gradient on its side stream as in the graph)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep  # noqa: E402

B = int(os.environ.get("BATCH", "4096"))
L = cdq._lib.lib()
boards = torch.empty((B + 1, 9), dtype=torch.int64, device="cuda")
cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), B + 1, 42, 199, cdq._lib.stream_ptr()))
torch.manual_seed(0)
net, tgt = cdq.ChessQNetwork().cuda(), cdq.ChessQNetwork().cuda()
tgt.load_state_dict(net.state_dict())
opt = FusedAdam(net)
step = GraphedDQNStep(net, tgt, opt, batch=B)
r = torch.rand(B, device="cuda") * 0.01
d = (torch.rand(B, device="cuda") < 0.05).float()
step.load(boards[:B], boards[1:], r, d)
for _ in range(5):
    step.run()
torch.cuda.synchronize()
raw = ctypes.CDLL(cdq._lib.SO_PATH)
buf = (ctypes.c_longlong * 384)()
assert raw.cdq_debug_cb2_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(3, 16, 8)
t0 = t[0, 14, 0]
rel = lambda x: (x - t0) / 1000.0
print("# kcycles since kernel entry (CTA 0, inside the step's graph replay)")
print("prologue done %.1f, all roles done %.1f" % (rel(t[0, 15, 0]), rel(t[0, 15, 1])))
for i in range(9):
    if t[1, i, 0] == 0:
        break
    print("unit %d | producer: buffer free %.1f, fetch issued %.1f | issuer: boards there %.1f, wgrad issued %.1f, dgrad accumulators free %.1f, "
          "all issued %.1f | epilogue: boards %.1f, masks+bias %.1f, accumulators complete %.1f, drained %.1f, combined %.1f, staged+stored %.1f"
          % (i, rel(t[0, i, 0]), rel(t[0, i, 1]), rel(t[1, i, 0]), rel(t[1, i, 1]), rel(t[1, i, 2]), rel(t[1, i, 3]),
             rel(t[2, i, 0]), rel(t[2, i, 1]), rel(t[2, i, 2]), rel(t[2, i, 3]), rel(t[2, i, 4]), rel(t[2, i, 5])))
print("tail: epilogue leaves the loop %.1f, every UMMA complete %.1f" % (rel(t[2, 14, 0]), rel(t[2, 14, 1])))
