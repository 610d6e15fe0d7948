"""Arbitrary-position generator shared by the tests and tests/golden/make_golden.py."""
import numpy as np


def random_boards(n, seed):
    """Arbitrary (mostly unreachable) positions: one king per side, 0-30 other pieces anywhere except
    pawns on the back ranks, random turn, castling rights where king and rook stand right, sometimes an
    ep square behind a pawn that could have just double-pushed.  Stresses multiple checkers, many
    sliders, pins while in check, kings in the corner etc."""
    rng = np.random.default_rng(seed)
    out = np.zeros(n, dtype=[("bb", "<u8", (8,)), ("meta", "<u8")])
    for i in range(n):
        k = int(rng.integers(0, 31))
        sqs = rng.permutation(64)[: k + 2]
        bb = [0] * 8
        bb[5] |= (1 << int(sqs[0])) | (1 << int(sqs[1]))
        bb[6] |= 1 << int(sqs[0])
        bb[7] |= 1 << int(sqs[1])
        for s in sqs[2:]:
            s = int(s)
            t = int(rng.integers(0, 5))
            if t == 0 and (s < 8 or s >= 56):
                t = int(rng.integers(1, 5))
            bb[t] |= 1 << s
            bb[6 + int(rng.integers(0, 2))] |= 1 << s
        meta = int(rng.integers(0, 2))
        white, black, rooks, kings = bb[6], bb[7], bb[3], bb[5]
        if kings & white & (1 << 4):
            if rooks & white & (1 << 7) and rng.random() < 0.5: meta |= 2
            if rooks & white & (1 << 0) and rng.random() < 0.5: meta |= 4
        if kings & black & (1 << 60):
            if rooks & black & (1 << 63) and rng.random() < 0.5: meta |= 8
            if rooks & black & (1 << 56) and rng.random() < 0.5: meta |= 16
        if rng.random() < 0.3:
            occ = white | black
            wtm = meta & 1
            # white to move: black just pushed a7-a5 style: black pawn on rank 5, ep square on rank 6 empty
            cands = [s for s in range(32, 40) if (bb[0] & black) >> s & 1 and not occ >> (s + 8) & 1 and not occ >> (s + 16) & 1] if wtm else \
                    [s for s in range(24, 32) if (bb[0] & white) >> s & 1 and not occ >> (s - 8) & 1 and not occ >> (s - 16) & 1]
            if cands:
                s = int(rng.choice(cands))
                meta |= ((s + 8 if wtm else s - 8) + 1) << 8
        out["bb"][i] = np.array(bb, dtype=np.uint64)
        out["meta"][i] = meta
    return out
