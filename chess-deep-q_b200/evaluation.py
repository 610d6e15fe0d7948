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


# ------------------------------------------------------------------ attack / movement maps (evaluation.py:368-428)
def attack_maps(boards):
    """-> int64 CUDA tensor [n,4]: bitboards {attacked by white, attacked by black, legal
    destinations of white, legal destinations of black} (bit i = square i)."""
    import torch
    d = to_device(boards)
    out = torch.empty((d.shape[0], 4), dtype=torch.int64, device=d.device)
    _lib.check(_lib.lib().cdq_attack_maps(d.data_ptr(), d.shape[0], out.data_ptr(), _lib.stream_ptr()))
    return out


def _squares(bits):
    bits &= (1 << 64) - 1
    return {s for s in range(64) if (bits >> s) & 1}


def _maps_of(board):
    m = attack_maps(board)[0].tolist()
    return [x & ((1 << 64) - 1) for x in m]


def _turn_of(board):
    from .board_utils import as_records
    return bool(int(as_records(board)[0]["meta"]) & 1)


def calculate_attack_range(board, color):
    """evaluation.py:402-413"""
    return _squares(_maps_of(board)[0 if color else 1])


def calculate_movement_range(board, color):
    """evaluation.py:417-428"""
    return _squares(_maps_of(board)[2 if color else 3])


def find_threatened_squares(board):
    """evaluation.py:368-382: squares attacked by the opponent of the side to move"""
    return calculate_attack_range(board, not _turn_of(board))


def find_guarded_squares(board):
    """evaluation.py:384-398: squares attacked by the side to move"""
    return calculate_attack_range(board, _turn_of(board))
