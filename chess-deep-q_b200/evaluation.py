"""Handcrafted evaluation: mirrors the reference's evaluation.fast_evaluate_position
(evaluation.py:30-229) on the CUDA bitboard kernel (csrc/eval_kernel.cu)."""
from . import _lib
from .board_utils import to_device

# evaluation.py:8-15 (python-chess piece type numbers 1..6)
PIECE_VALUES = {1: 1, 2: 3, 3: 3, 4: 5, 5: 9, 6: 0}

NTERMS = 24
TERM_NAMES = [
    "w_mat", "b_mat", "w_mob", "b_mob", "w_att", "b_att", "w_katt", "b_katt",
    "w_castled", "b_castled", "w_kcentre", "b_kcentre", "w_dbl", "b_dbl", "w_iso", "b_iso",
    "w_centre", "b_centre", "w_ext", "b_ext", "w_def", "b_def", "check", "terminal",
]
F_BADBOARD = 0x80000000


def evaluate_batch(boards, ignore_checkmate=False, return_terms=False):
    """Batch evaluation.  -> (score float64 [n], flags int32 [n]) CUDA tensors, plus the integer
    terms int32 [n,24] when return_terms.  `boards`: anything board_utils.to_device accepts."""
    import torch
    d = to_device(boards)
    n = d.shape[0]
    score = torch.empty(n, dtype=torch.float64, device=d.device)
    flags = torch.empty(n, dtype=torch.int32, device=d.device)
    terms = torch.empty((n, NTERMS), dtype=torch.int32, device=d.device) if return_terms else None
    L = _lib.lib()
    _lib.check(L.cdq_eval_terms(d.data_ptr(), n, int(bool(ignore_checkmate)),
                                terms.data_ptr() if return_terms else None,
                                score.data_ptr(), flags.data_ptr(), _lib.stream_ptr()))
    return (score, flags, terms) if return_terms else (score, flags)


def fast_evaluate_position(board, ignore_checkmate=False):
    """Drop-in for evaluation.fast_evaluate_position(board, ignore_checkmate=False) -> float.
    (The reference's FEN-keyed cache only memoises; it is not reproduced.)"""
    score, flags = evaluate_batch(board, ignore_checkmate)
    if int(flags[0].item()) & F_BADBOARD:
        raise _lib.CdqError("malformed board record")
    return float(score[0].item())
