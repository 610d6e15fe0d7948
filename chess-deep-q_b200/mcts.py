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
    repetition (:205-210), else clamp(net(board), -1, 1) (:213-221).  neural_net = None is the
    reference's heuristic branch (:226-229): tanh(fast_evaluate_position(board) / 10) behind the same
    terminal rules, computed by the evaluation kernel alone."""

    def __init__(self, neural_net=None):
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
        if want_raw and self.net is None:
            raise _lib.CdqError("want_raw needs a network: without one the leaf value is tanh(score / 10)")
        raw = torch.empty(n, dtype=torch.float32, device=dev) if want_raw else None
        if n and self.net is None:
            _lib.check(_lib.lib().cdq_leaf_eval(None, d.data_ptr(), n, out["leaf"].data_ptr(), out["score"].data_ptr(),
                                                out["flags"].data_ptr(), None, None, 0, _lib.stream_ptr()))
        elif n:
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
        handle = self.net.net_handle() if self.net is not None else None
        torch.cuda.current_stream().synchronize()  # weights may have been repacked on this stream
        _lib.check(_lib.lib().cdq_leaf_eval_host(handle, src, n, ptr(out["leaf"]),
                                                 ptr(out["score"]), ptr(out["flags"])))
        self.total_positions_evaluated += n
        return out


def evaluate_positions(boards, neural_net=None):
    """Batch form of _evaluate_position: -> fp32 [n] leaf values (CUDA tensor)."""
    return LeafEvaluator(neural_net).evaluate_device(boards)["leaf"]


def evaluate_position(board, neural_net=None, device=None):
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
# Children are sampled like select_weighted_moves does, on the device: tactical weights of
# categorize_moves from cdq_expand_weighted, the weighted draw without replacement from
# cdq_sample_children (row f2).  Differences, stated: nodes are keyed by
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
        self.U = {}                 # key -> indices of the children no simulation has taken yet
        self.total_positions_evaluated = 0
        self.last_leaves = None     # (records, leaf values) of the last search, for inspection/tests

    @staticmethod
    def _key(rec):
        return rec.tobytes()

    @staticmethod
    def _device_expand(recs, seeds, width):
        """Successors (cdq_expand_weighted), terminal flags (cdq_eval_terms) and the weighted sample of
        `width` children per node (cdq_sample_children = select_weighted_moves, evaluation.py:342-366)
        for a batch of nodes of any number of trees: three launches, and only the sampled children
        (width x 72 B per node instead of 128 x 73 B) come back to the host."""
        d = to_device(recs)
        n, maxc = len(recs), 128
        dev = d.device
        kids = torch.empty((n, maxc, 9), dtype=torch.int64, device=dev)
        counts = torch.empty(n, dtype=torch.int32, device=dev)
        flags = torch.empty(n, dtype=torch.int32, device=dev)
        wts = torch.empty((n, maxc), dtype=torch.uint8, device=dev)
        sel = torch.empty((n, width, 9), dtype=torch.int64, device=dev)
        sel_counts = torch.empty(n, dtype=torch.int32, device=dev)
        d_seeds = torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64)).to(dev)
        L, st = _lib.lib(), _lib.stream_ptr()
        _lib.check(L.cdq_expand_weighted(d.data_ptr(), n, maxc, kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), st))
        _lib.check(L.cdq_eval_terms(d.data_ptr(), n, 0, None, None, flags.data_ptr(), st))
        _lib.check(L.cdq_sample_children(kids.data_ptr(), counts.data_ptr(), wts.data_ptr(), n, maxc, width,
                                         d_seeds.data_ptr(), sel.data_ptr(), sel_counts.data_ptr(), st))
        sel_h = sel.cpu().numpy().view(np.uint64).reshape(n, width, 9)
        return sel_h, sel_counts.cpu().numpy(), flags.cpu().numpy()

    @classmethod
    def _expand_many(cls, requests):
        """requests: (tree, record) pairs.  Nodes a tree has not expanded yet are expanded and sampled
        with ONE batch of device calls for all trees; every node gets a 64-bit sampling seed drawn
        from its own tree's RNG, in request order."""
        todo, seen = [], set()
        for tree, rec in requests:
            key = cls._key(rec)
            if key in tree.children or (id(tree), key) in seen:
                continue
            seen.add((id(tree), key))
            todo.append((tree, key, rec))
        if not todo:
            return
        width = todo[0][0].samples_per_level[0]
        assert all(t.samples_per_level[0] == width for t, _, _ in todo), "trees searched together share the child width"
        seeds = [t.rng.getrandbits(64) for t, _, _ in todo]
        sel_h, sel_counts, flags = cls._device_expand(np.array([r for _, _, r in todo], dtype=BOARD_DTYPE), seeds, width)
        recs = sel_h.view(BOARD_DTYPE).reshape(len(todo), width)         # views into one batch array, no copies
        n_all, q_all = np.zeros((len(todo), width), np.int64), np.zeros((len(todo), width), np.float64)
        for i, (tree, key, _) in enumerate(todo):
            c = 0 if int(flags[i]) & 0x40 else int(sel_counts[i])      # insufficient material: game over
            tree.children[key] = recs[i, :c]
            tree.N[key] = n_all[i, :c]
            tree.Q[key] = q_all[i, :c]
            tree.U[key] = list(range(c))

    def _expand(self, records):
        """Expand a batch of not-yet-expanded nodes on the device (mcts.py:166-194)."""
        self._expand_many([(self, r) for r in records])

    def _select(self, key, training_progress, current_move, max_moves):
        """mcts.py:34-67: a random child that no simulation has taken yet, else the UCB argmax.  The
        never-taken children of a node are a list that shrinks as simulations claim them (also within
        a wave, before its back-up), so simultaneous simulations spread over unexplored children like
        the reference's threads do."""
        fresh = self.U[key]
        if fresh:
            k = self.rng.randrange(len(fresh))
            fresh[k], fresh[-1] = fresh[-1], fresh[k]
            return fresh.pop()
        counts = self.N[key]
        if len(counts) == 0:
            return None
        total = int(counts.sum())
        if total == 0:
            return self.rng.randrange(len(counts))
        alpha = 1 + training_progress
        beta = current_move / max_moves if max_moves > 0 else 0.5
        eff = self.exploration_weight * (1.0 - 0.5 * alpha * beta)
        log_total = math.log(total + 1e-10)
        nk = counts + 1e-10
        # as written in the reference: the running mean Q is divided by N once more (mcts.py:61-62)
        ucb = self.Q[key] / nk + eff * np.sqrt(log_total / nk)
        return int(np.argmax(ucb))

    def search(self, neural_net=None, device=None, training_progress=0.0, current_move=0, max_moves=200,
               evaluator=None):
        """-> (best child record, (from, to, promotion)) or (None, None) when the root has no move."""
        ev = evaluator or LeafEvaluator(neural_net)
        return search_many([self], ev, training_progress, [current_move], max_moves)[0]


def search_many(trees, evaluator, training_progress=0.0, current_moves=None, max_moves=200):
    """Run the searches of several trees (e.g. one per concurrent self-play game) in lockstep: every
    level of a wave costs ONE device expansion and every wave ONE leaf evaluation for all trees
    together, so the batch the GPU sees grows with the number of games.  Each tree keeps its own
    statistics and RNG, so its result is what `tree.search()` alone would have produced."""
    current_moves = current_moves or [0] * len(trees)
    BatchedRussianDollMCTS._expand_many([(t, t.root[0]) for t in trees])
    iterations = max(t.iterations for t in trees)
    workers = trees[0].num_workers
    leaves_of, values_of = [[] for _ in trees], [[] for _ in trees]
    # The reference runs `num_workers` simulations concurrently (mcts.py:77) and backs each one up
    # as it finishes; here a wave of `num_workers` simulations descends level by level together and
    # is backed up before the next wave selects, so UCB sees the same amount of feedback.
    for first in range(0, iterations, workers):
        waves = []
        for t in trees:
            wave = max(0, min(t.num_workers, t.iterations - first))
            waves.append([{"rec": t.root[0], "path": [], "alive": True} for _ in range(wave)])
        depth_limit = max(len(t.samples_per_level) for t in trees)
        for depth in range(depth_limit):
            live = [[s for s in sims if s["alive"]] if depth < len(t.samples_per_level) else []
                    for t, sims in zip(trees, waves)]
            if not any(live):
                break
            BatchedRussianDollMCTS._expand_many([(t, s["rec"]) for t, ls in zip(trees, live) for s in ls])
            for t, ls, cm in zip(trees, live, current_moves):
                for s in ls:
                    key = t._key(s["rec"])
                    a = t._select(key, training_progress, cm + depth, max_moves)
                    if a is None:
                        s["alive"] = False
                        continue
                    s["path"].append((key, a))
                    s["rec"] = t.children[key][a]
        batch = [s["rec"] for sims in waves for s in sims]
        if not batch:
            continue
        values = evaluator.evaluate_host(np.array(batch, dtype=BOARD_DTYPE))["leaf"]
        pos = 0
        for ti, (t, sims) in enumerate(zip(trees, waves)):
            if not sims:
                continue
            v = values[pos:pos + len(sims)]
            pos += len(sims)
            t.total_positions_evaluated += len(sims)
            leaves_of[ti].append(np.array([s["rec"] for s in sims], dtype=BOARD_DTYPE))
            values_of[ti].append(v.copy())
            for s, x in zip(sims, v):                           # mcts.py:158-164
                value = float(x)
                for key, a in reversed(s["path"]):
                    t.N[key][a] += 1
                    t.Q[key][a] += (value - t.Q[key][a]) / t.N[key][a]
                    value = -value
    out = []
    for ti, t in enumerate(trees):
        if leaves_of[ti]:
            t.last_leaves = (np.concatenate(leaves_of[ti]), np.concatenate(values_of[ti]))
        root_key = t._key(t.root[0])
        kids = t.children[root_key]
        if len(kids) == 0:
            out.append((None, None))
            continue
        counts = t.N[root_key]
        best = int(np.argmax(counts)) if counts.any() else t.rng.randrange(len(kids))
        child = kids[best:best + 1].copy()
        out.append((child, move_from_records(t.root[0], child[0])))
    return out


def _startpos():
    start = np.zeros(1, BOARD_DTYPE)
    start["bb"][0] = [0x00ff00000000ff00, 0x4200000000000042, 0x2400000000000024, 0x8100000000000081,
                      0x0800000000000008, 0x1000000000000010, 0xffff, 0xffff000000000000]
    start["meta"][0] = 0x1f
    return start


def self_play_games(neural_net, n_games, max_moves=200, iterations=50, samples_per_level=None, seeds=None,
                    training_progress=0.0):
    """`n_games` self-play games (chess_ai.py:117-250) advanced ply by ply in lockstep, their searches
    batched by search_many: one device expansion per tree level and one leaf evaluation per wave for
    ALL games.  Game i uses RNG seed seeds[i] and is move for move the game self_play_game(seed =
    seeds[i]) plays alone.  -> list of (positions, scores)."""
    from .evaluation import evaluate_batch
    seeds = list(seeds) if seeds is not None else list(range(42, 42 + n_games))
    rngs = [_random.Random(s) for s in seeds]
    ev = LeafEvaluator(neural_net)
    widths = samples_per_level or [21, 13, 8, 5, 3, 2, 1]
    positions = [[_startpos()] for _ in range(n_games)]
    scores = [[] for _ in range(n_games)]
    active = list(range(n_games))
    for ply in range(max_moves):
        if not active:
            break
        cur = np.concatenate([positions[g][-1] for g in active])
        score, flags = evaluate_batch(cur)
        score_h, flags_h = score.cpu().numpy(), flags.cpu().numpy()
        still, trees = [], []
        for k, g in enumerate(active):
            scores[g].append(float(score_h[k]))
            if int(flags_h[k]) & 0x70:          # checkmate / stalemate / insufficient material
                continue
            still.append(g)
            trees.append(BatchedRussianDollMCTS(positions[g][-1], iterations=iterations, samples_per_level=widths,
                                                rng=rngs[g]))
        if not trees:
            active = []
            break
        results = search_many(trees, ev, training_progress, [ply] * len(trees), max_moves)
        active = []
        for g, (child, _) in zip(still, results):
            if child is None:
                continue
            positions[g].append(child)
            active.append(g)
    return [(np.concatenate(positions[g]), scores[g]) for g in range(n_games)]


def self_play_game(neural_net, max_moves=200, iterations=50, samples_per_level=None, seed=42, start=None,
                   training_progress=0.0):
    """One self-play game in the manner of OptimizedChessAI.self_play_game (chess_ai.py:117-250) with
    the batched search: returns the list of positions (records) and the per-ply handcrafted scores
    that the reference uses for reward shaping (chess_ai.py:161,182)."""
    from .evaluation import evaluate_batch
    rng = _random.Random(seed)
    ev = LeafEvaluator(neural_net)
    if start is None:
        start = _startpos()
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
