"""Leaf evaluation call site: mirrors ParallelRussianDollMCTS._evaluate_position (mcts.py:196-232)
for batches of leaves.  One call = K1 (score + terminal flags) and K2 (CNN value) on the device."""
import numpy as np
import torch

from . import _lib
from .board_utils import BOARD_DTYPE, as_records, to_device
from .neural_network import _Workspace


class LeafEvaluator:
    """Batched replacement for per-leaf `_evaluate_position`: leaf value = -1/+1 on checkmate (from
    white's point of view, mcts.py:202-203), 0 on stalemate / insufficient material / threefold
    repetition (:205-210), else clamp(net(board), -1, 1) (:213-221)."""

    def __init__(self, neural_net):
        self.net = neural_net
        self._ws = _Workspace()
        self.total_positions_evaluated = 0  # mcts.py:28

    def evaluate_device(self, boards, want_raw=False):
        """boards: device int64 [n,9] (or anything to_device accepts).
        -> dict(leaf fp32 [n], score fp64 [n], flags int32 [n][, value fp32 [n]]) of CUDA tensors."""
        d = to_device(boards)
        n = d.shape[0]
        dev = d.device
        out = {"leaf": torch.empty(n, dtype=torch.float32, device=dev),
               "score": torch.empty(n, dtype=torch.float64, device=dev),
               "flags": torch.empty(n, dtype=torch.int32, device=dev)}
        raw = torch.empty(n, dtype=torch.float32, device=dev) if want_raw else None
        if n:
            ws = self._ws.get(n, dev)
            _lib.check(_lib.lib().cdq_leaf_eval(self.net.net_handle(), d.data_ptr(), n, out["leaf"].data_ptr(),
                                                out["score"].data_ptr(), out["flags"].data_ptr(),
                                                raw.data_ptr() if want_raw else None,
                                                ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        if want_raw:
            out["value"] = raw
        self.total_positions_evaluated += n
        return out

    def evaluate_host(self, boards, out=None):
        """Host records in, host numpy arrays out (copies inside): the end-to-end path.
        `boards` may be a pinned uint8/int64 torch tensor or a numpy record array."""
        if isinstance(boards, torch.Tensor):
            n = boards.numel() * boards.element_size() // 72
            src = boards.data_ptr()
        else:
            rec = as_records(boards)
            n = len(rec)
            src = rec.ctypes.data
        if out is None:
            out = {"leaf": np.empty(n, np.float32), "score": np.empty(n, np.float64), "flags": np.empty(n, np.uint32)}
        ptr = lambda a: a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data  # noqa: E731
        handle = self.net.net_handle()
        torch.cuda.current_stream().synchronize()  # weights may have been repacked on this stream
        _lib.check(_lib.lib().cdq_leaf_eval_host(handle, src, n, ptr(out["leaf"]),
                                                 ptr(out["score"]), ptr(out["flags"])))
        self.total_positions_evaluated += n
        return out


def evaluate_positions(boards, neural_net):
    """Batch form of _evaluate_position: -> fp32 [n] leaf values (CUDA tensor)."""
    return LeafEvaluator(neural_net).evaluate_device(boards)["leaf"]


def evaluate_position(board, neural_net, device=None):
    """Drop-in for ParallelRussianDollMCTS._evaluate_position(board, neural_net, device) -> float."""
    return float(evaluate_positions(board, neural_net)[0].item())
