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


# ---------------------------------------------------------------------------------------------
# batched-leaf Russian Doll MCTS.  Same tree policy as the reference
# (ParallelRussianDollMCTS, mcts.py:14-194: UCB with the annealed exploration weight, unexplored
# children first, running-mean backup with a sign flip per ply, best move = most visited root
# child, depth limit = len(samples_per_level), children per node = samples_per_level[0]), but the
# simulations advance level by level in waves of `num_workers` (the reference's thread count): every
# level of a wave costs one device expansion (cdq_expand) and the wave's leaves cost one
# cdq_leaf_eval, instead of one batch-1 network call per simulation under the GIL.
# Children are sampled like select_weighted_moves does, with the tactical weights of categorize_moves
# computed on the device (cdq_expand_weighted, row f2).  Differences, stated: nodes are keyed by
# their 72-byte record instead of a FEN; positions inside a search carry no move history (rep3 = 0).
import math
import random as _random


def move_from_records(parent, child):
    """(from_square, to_square, promotion piece type or None) of the move parent -> child, in
    python-chess numbering (promotion 2..5 = N,B,R,Q)."""
    pb, cb = [int(x) for x in parent["bb"]], [int(x) for x in child["bb"]]
    us = 6 if int(parent["meta"]) & 1 else 7
    gone = pb[us] & ~cb[us]
    came = cb[us] & ~pb[us]
    kings = pb[5] & pb[us]
    if bin(gone).count("1") == 2:            # castling: report the king's move
        gone &= kings
        came &= cb[5]
    frm, to = gone.bit_length() - 1, came.bit_length() - 1
    promo = None
    if (pb[0] >> frm) & 1 and not (cb[0] >> to) & 1:
        promo = next(t + 1 for t in (1, 2, 3, 4) if (cb[t] >> to) & 1)
    return frm, to, promo


def move_uci(parent, child):
    f, t, p = move_from_records(parent, child)
    name = lambda s: "abcdefgh"[s & 7] + str((s >> 3) + 1)  # noqa: E731
    return name(f) + name(t) + ("" if p is None else " nbrq"[p])


class BatchedRussianDollMCTS:
    def __init__(self, board, iterations=20, exploration_weight=1.0, samples_per_level=None, num_workers=None,
                 rng=None):
        self.root = as_records(board)[:1].copy()
        self.iterations = iterations
        self.exploration_weight = exploration_weight
        self.samples_per_level = samples_per_level or [21, 13, 8, 5, 3, 2, 1]
        self.num_workers = num_workers or 8       # simulations per wave (mcts.py:20: min(8, cpu_count))
        self.rng = rng or _random.Random()
        self.children = {}          # key -> BOARD_DTYPE array of sampled children
        self.N = {}                 # key -> visit counts per child
        self.Q = {}                 # key -> running-mean values per child
        self.total_positions_evaluated = 0
        self.last_leaves = None     # (records, leaf values) of the last search, for inspection/tests

    @staticmethod
    def _key(rec):
        return rec.tobytes()

    def _expand(self, records):
        """Expand a batch of not-yet-expanded nodes on the device (mcts.py:166-194)."""
        todo = [r for r in records if self._key(r) not in self.children]
        if not todo:
            return
        uniq = {self._key(r): r for r in todo}
        recs = np.array(list(uniq.values()), dtype=BOARD_DTYPE)
        d = to_device(recs)
        n, maxc = len(recs), 128
        kids = torch.empty((n, maxc, 9), dtype=torch.int64, device=d.device)
        counts = torch.empty(n, dtype=torch.int32, device=d.device)
        flags = torch.empty(n, dtype=torch.int32, device=d.device)
        wts = torch.empty((n, maxc), dtype=torch.uint8, device=d.device)
        L = _lib.lib()
        _lib.check(L.cdq_expand_weighted(d.data_ptr(), n, maxc, kids.data_ptr(), counts.data_ptr(), wts.data_ptr(),
                                         _lib.stream_ptr()))
        _lib.check(L.cdq_eval_terms(d.data_ptr(), n, 0, None, None, flags.data_ptr(), _lib.stream_ptr()))
        kids_h = kids.cpu().numpy().view(np.uint64).reshape(n, maxc, 9)
        counts_h, flags_h, wts_h = counts.cpu().numpy(), flags.cpu().numpy(), wts.cpu().numpy()
        width = self.samples_per_level[0]
        for i, key in enumerate(uniq):
            c = min(int(counts_h[i]), maxc)
            if c == 0 or (int(flags_h[i]) & 0x40):          # no legal move / insufficient material: game over
                sel = np.zeros(0, BOARD_DTYPE)
            else:
                idx = list(range(c))
                if c > width:
                    # select_weighted_moves (evaluation.py:342-366): `width` draws without replacement,
                    # probability proportional to the tactical weight of the move
                    pool, w, idx = idx, [float(x) for x in wts_h[i, :c]], []
                    for _ in range(width):
                        r, acc = self.rng.random() * sum(w), 0.0
                        for k, wk in enumerate(w):
                            acc += wk
                            if r < acc:
                                break
                        idx.append(pool.pop(k))
                        w.pop(k)
                sel = np.ascontiguousarray(kids_h[i, idx]).view(BOARD_DTYPE).reshape(-1)
            self.children[key] = sel
            self.N[key] = np.zeros(len(sel), np.int64)
            self.Q[key] = np.zeros(len(sel), np.float64)

    def _select(self, key, pending, training_progress, current_move, max_moves):
        """mcts.py:34-67.  `pending` holds the wave's not-yet-backed-up selections so that
        simultaneous simulations spread over unexplored children like the reference's threads do."""
        n_children = len(self.children[key])
        if n_children == 0:
            return None
        alpha = 1 + training_progress
        beta = current_move / max_moves if max_moves > 0 else 0.5
        eff = self.exploration_weight * (1.0 - 0.5 * alpha * beta)
        visits = self.N[key] + pending.get(key, 0)
        unexplored = np.nonzero(visits == 0)[0]
        if len(unexplored):
            return int(self.rng.choice(list(unexplored)))
        total = int(self.N[key].sum())
        if total == 0:
            return self.rng.randrange(n_children)
        log_total = math.log(total + 1e-10)
        nk = self.N[key] + 1e-10
        # as written in the reference: the running mean Q is divided by N once more (mcts.py:61-62)
        ucb = self.Q[key] / nk + eff * np.sqrt(log_total / nk)
        return int(np.argmax(ucb))

    def search(self, neural_net=None, device=None, training_progress=0.0, current_move=0, max_moves=200,
               evaluator=None):
        """-> (best child record, (from, to, promotion)) or (None, None) when the root has no move."""
        ev = evaluator or LeafEvaluator(neural_net)
        root_key = self._key(self.root[0])
        self._expand([self.root[0]])
        depth_limit = len(self.samples_per_level)
        all_leaves, all_values = [], []
        # The reference runs `num_workers` simulations concurrently (mcts.py:77) and backs each one up
        # as it finishes; here a wave of `num_workers` simulations descends level by level together and
        # is backed up before the next wave selects, so UCB sees the same amount of feedback.
        for first in range(0, self.iterations, self.num_workers):
            wave = min(self.num_workers, self.iterations - first)
            sims = [{"rec": self.root[0], "path": [], "alive": True} for _ in range(wave)]
            pending = {}
            for depth in range(depth_limit):
                live = [s for s in sims if s["alive"]]
                if not live:
                    break
                self._expand([s["rec"] for s in live])
                for s in live:
                    key = self._key(s["rec"])
                    a = self._select(key, pending, training_progress, current_move + depth, max_moves)
                    if a is None:
                        s["alive"] = False
                        continue
                    p = pending.setdefault(key, np.zeros(len(self.children[key]), np.int64))
                    p[a] += 1
                    s["path"].append((key, a))
                    s["rec"] = self.children[key][a]
            leaves = np.array([s["rec"] for s in sims], dtype=BOARD_DTYPE)
            values = ev.evaluate_host(leaves)["leaf"]
            self.total_positions_evaluated += len(leaves)
            all_leaves.append(leaves)
            all_values.append(values.copy())
            for s, v in zip(sims, values):                      # mcts.py:158-164
                value = float(v)
                for key, a in reversed(s["path"]):
                    self.N[key][a] += 1
                    self.Q[key][a] += (value - self.Q[key][a]) / self.N[key][a]
                    value = -value
        self.last_leaves = (np.concatenate(all_leaves), np.concatenate(all_values))
        kids = self.children[root_key]
        if len(kids) == 0:
            return None, None
        counts = self.N[root_key]
        best = int(np.argmax(counts)) if counts.any() else self.rng.randrange(len(kids))
        child = kids[best:best + 1].copy()
        return child, move_from_records(self.root[0], child[0])


def self_play_game(neural_net, max_moves=200, iterations=50, samples_per_level=None, seed=42, start=None,
                   training_progress=0.0):
    """One self-play game in the manner of OptimizedChessAI.self_play_game (chess_ai.py:117-250) with
    the batched search: returns the list of positions (records) and the per-ply handcrafted scores
    that the reference uses for reward shaping (chess_ai.py:161,182)."""
    from .evaluation import evaluate_batch
    rng = _random.Random(seed)
    ev = LeafEvaluator(neural_net)
    if start is None:
        start = np.zeros(1, BOARD_DTYPE)
        start["bb"][0] = [0x00ff00000000ff00, 0x4200000000000042, 0x2400000000000024, 0x8100000000000081,
                          0x0800000000000008, 0x1000000000000010, 0xffff, 0xffff000000000000]
        start["meta"][0] = 0x1f
    positions, scores = [as_records(start)[:1].copy()], []
    widths = samples_per_level or [21, 13, 8, 5, 3, 2, 1]
    for ply in range(max_moves):
        cur = positions[-1]
        score, flags = evaluate_batch(cur)
        scores.append(float(score[0].item()))
        if int(flags[0].item()) & 0x70:          # checkmate / stalemate / insufficient material
            break
        tree = BatchedRussianDollMCTS(cur, iterations=iterations, samples_per_level=widths, rng=rng)
        child, _ = tree.search(neural_net, current_move=ply, max_moves=max_moves, evaluator=ev,
                               training_progress=training_progress)
        if child is None:
            break
        positions.append(child)
    return np.concatenate(positions), scores
