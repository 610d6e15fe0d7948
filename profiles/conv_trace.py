"""Timeline study of conv_stream_kernel: per-role clock64 stamps of CTA 0 (a -DCDQ_TRACE build).

    python chess-deep-q_b200/build.py -DCDQ_TRACE --out=libcdq_trace.so
    CDQ_LIBCDQ=chess-deep-q_b200/libcdq_trace.so python profiles/conv_trace.py

Prints, for 32 consecutive file steps in steady state, the cycles every role spends waiting and
working (tags are defined next to the TR() stamps in csrc/conv_stream_kernel.cu).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import chess_deep_q_b200 as cdq  # noqa: E402

L = cdq._lib.lib()
import ctypes  # noqa: E402

L.cdq_debug_set_trace.restype = ctypes.c_int
L.cdq_debug_set_trace.argtypes = [ctypes.c_void_p]

n = 131072
boards = torch.empty((n, 9), dtype=torch.int64, device="cuda")
cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), n, 42, 199, cdq._lib.stream_ptr()))
net = cdq.ChessQNetwork().cuda()
val = torch.empty(n, dtype=torch.float32, device="cuda")
ws = torch.empty(L.cdq_workspace_bytes(n), dtype=torch.uint8, device="cuda")
trace = torch.zeros(4 * 32 * 8, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2:
        cdq._lib.check(L.cdq_debug_set_trace(trace.data_ptr()))
    cdq._lib.check(L.cdq_value_fwd(net.net_handle(), boards.data_ptr(), n, val.data_ptr(), ws.data_ptr(), ws.numel(),
                                   cdq._lib.stream_ptr()))
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(4, 32, 8)
t0 = t[t > 0].min()
rel = np.where(t > 0, t - t0, -1)
names = ["MMA   c1: 0 step 1 rows0-1.issued 2 next.ready / c2: 3 step 4 rows0-3.issued 5 next.ready 6 rest.issued",
         "ENC   0 wait 1 slab.free 2 stored 3 fenced",
         "EPI1  0 wait 1 c1.full 2 loaded 3 slab.free 4 stored 5 fenced",
         "EPI2  0 wait 1 c2.full 2 loaded 3 stored"]
for r in range(4):
    print(names[r])
    for s in range(32):
        row = rel[r, s]
        if (row >= 0).sum() == 0:
            continue
        print("  step %3d  abs %s" % (s, " ".join("%7d" % x for x in row)))
    starts = rel[r, :, 0]
    ok = starts[starts >= 0]
    if len(ok) > 1:
        print("  mean period %.0f cycles per step" % ((ok[-1] - ok[0]) / (len(ok) - 1)))
