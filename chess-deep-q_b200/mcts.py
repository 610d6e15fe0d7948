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
# the Russian Doll MCTS itself, device resident (csrc/tree_kernel.cu).  Same tree policy as the
# reference (ParallelRussianDollMCTS, mcts.py:14-194: UCB with the annealed exploration weight, unexplored children
# first, running-mean backup with a sign flip per ply, best move = most visited root child, depth limit =
# len(samples_per_level), children per node = samples_per_level[0] sampled like select_weighted_moves from
# categorize_moves' weights), but trees, selection, expansion, Board.push, repetition bookkeeping and backup all live on
# the GPU for any number of concurrent games: the simulations advance level by level in waves of `num_workers` (the
# reference's thread count), a wave's leaves of ALL games are ONE cdq_leaf_eval, and the host synchronises once per
# search.  Differences from the reference, stated: nodes are keyed by a 64-bit hash of what a FEN holds (position,
# halfmove clock, move number) instead of the FEN string; random draws are counter-based hashes, so move-level parity
# with the reference is statistical while leaf values and tree statistics are exact against the host restatement
# (tests/test_gpu_forest.py).
import math
import random as _random


def _decode_move(code):
    """forest move code -> (from_square, to_square, promotion piece type 2..5 or None) in python-chess numbering"""
    code = int(code)
    promo = (code >> 12) & 15
    return code & 63, (code >> 6) & 63, (promo if promo else None)       # (promotion type + 1) - 1 + 1: PT_N = 1 -> KNIGHT = 2


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


def history_records(board):
    """What the device needs of a python-chess board's past for is_repetition (mcts.py:151,209): the positions BEFORE
    the current one back to the last irreversible move, oldest first, as packed records; plus halfmove clock and ply.
    Boards without a move stack (packed records, FEN-built boards) have no history."""
    from .board_utils import pack_board
    hmc = int(getattr(board, "halfmove_clock", 0) or 0)
    ply = int(board.ply()) if hasattr(board, "ply") else 0
    recs = []
    if getattr(board, "move_stack", None) and hasattr(board, "is_irreversible"):
        b = board.copy()
        while b.move_stack and len(recs) < 255:
            mv = b.pop()
            if b.is_irreversible(mv):
                break
            recs.append(pack_board(b))
        recs.reverse()
    return (np.concatenate(recs) if recs else np.zeros(0, BOARD_DTYPE)), hmc, ply


class DeviceForest:
    """Host handle of a CdqForest (include/cdq.h): `n_games` games and their search trees in device memory."""

    HIST_MAX = 256

    def __init__(self, n_games, workers=8, samples_per_level=None, max_iterations=50):
        self.device = _lib.require_cuda()
        widths = list(samples_per_level or [21, 13, 8, 5, 3, 2, 1])
        self.n_games, self.workers, self.depth, self.width = int(n_games), int(workers), len(widths), int(widths[0])
        self.max_iterations = int(max_iterations)
        cfg = _lib.CdqForestConfig(self.n_games, self.workers, self.depth, self.width, self.max_iterations)
        import ctypes
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().cdq_forest_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h
        self._graphs = {}
        v = _lib.CdqForestView()
        _lib.check(_lib.lib().cdq_forest_view(h, ctypes.byref(v)))
        self._view = v
        T, W, N, E = self.n_games, self.workers, v.max_nodes, v.max_nodes * self.width
        self.max_nodes = v.max_nodes
        t = self._tensor
        self.game_boards = t(v.game_boards, (T, 9), torch.int64)
        self.game_halfmove_clock = t(v.game_halfmove_clock, (T,), torch.int32)
        self.game_ply = t(v.game_ply, (T,), torch.int32)
        self.game_status = t(v.game_status, (T,), torch.int32)
        self.game_history_len = t(v.game_history_len, (T,), torch.int32)
        self.best_move = t(v.best_move, (T,), torch.int32)
        self.best_visits = t(v.best_visits, (T,), torch.int32)
        self.root_children = t(v.root_children, (T,), torch.int32)
        self.edge_move = t(v.edge_move, (T, N, self.width), torch.int32)
        self.edge_visits = t(v.edge_visits, (T, N, self.width), torch.int32)
        self.edge_value = t(v.edge_value, (T, N, self.width), torch.float64)
        self.node_children = t(v.node_children, (T, N), torch.int32)
        self.node_count = t(v.node_count, (T,), torch.int32)
        self.leaf_boards = t(v.leaf_boards, (T, W, 9), torch.int64)
        self.leaf_values = t(v.leaf_values, (T, W), torch.float32)
        self.leaf_depth = t(v.leaf_depth, (T, W), torch.int32)
        self.positions_evaluated = t(v.positions_evaluated, (T,), torch.int64)

    def _tensor(self, ptr, shape, dtype):
        from .train import _DevicePtr
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        return torch.as_tensor(_DevicePtr(ptr, max(n, 1)), device=self.device)[:n].view(dtype).view(*shape)

    def close(self):
        if self._h is not None:
            torch.cuda.synchronize()
            self._graphs.clear()
            _lib.lib().cdq_forest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_games(self, boards, seeds, halfmove_clock=None, ply=None, history=None, active=None):
        """boards: n_games records; seeds: n_games uint64; history: list of record arrays (positions before the current
        one since the last irreversible move, oldest first) or None."""
        dev = self.device
        d = to_device(boards)
        assert d.shape[0] == self.n_games
        i32 = lambda a: None if a is None else torch.as_tensor(np.asarray(a, dtype=np.int32)).to(dev)  # noqa: E731
        d_seeds = torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64).copy()).to(dev)
        d_hmc, d_ply, d_active = i32(halfmove_clock), i32(ply), i32(active)
        d_hist = d_len = None
        stride = 0
        if history is not None and any(len(h) for h in history):
            stride = max(1, min(self.HIST_MAX, max(len(h) for h in history)))
            buf = np.zeros((self.n_games, stride), BOARD_DTYPE)
            lens = np.zeros(self.n_games, np.int32)
            for g, h in enumerate(history):
                h = as_records(h)[-stride:] if len(h) else h
                buf[g, :len(h)] = h
                lens[g] = len(h)
            d_hist = torch.from_numpy(buf.view(np.uint64).reshape(self.n_games, stride, 9).view(np.int64)).to(dev)
            d_len = torch.from_numpy(lens).to(dev)
        p = lambda x: None if x is None else x.data_ptr()  # noqa: E731
        _lib.check(_lib.lib().cdq_forest_set_games(self._h, d.data_ptr(), p(d_hmc), p(d_ply), p(d_hist), p(d_len), stride,
                                                   d_seeds.data_ptr(), p(d_active), _lib.stream_ptr()))

    def set_active(self, active):
        a = torch.as_tensor(np.asarray(active, dtype=np.int32)).to(self.device)
        _lib.check(_lib.lib().cdq_forest_set_active(self._h, a.data_ptr(), _lib.stream_ptr()))

    def search(self, neural_net=None, iterations=None, exploration_weight=1.0, training_progress=0.0, max_moves=200):
        """One search per active game; asynchronous (results are in the view tensors once the stream has run)."""
        it = self.max_iterations if iterations is None else int(iterations)
        handle = neural_net.net_handle() if neural_net is not None else None
        _lib.check(_lib.lib().cdq_forest_search(self._h, handle, it, float(exploration_weight), float(training_progress),
                                                int(max_moves), _lib.stream_ptr()))

    def advance(self):
        _lib.check(_lib.lib().cdq_forest_advance(self._h, _lib.stream_ptr()))

    def search_and_advance(self, neural_net=None, iterations=None, exploration_weight=1.0, training_progress=0.0, max_moves=200):
        """search() + advance() as ONE CUDA graph replay.  A search is a static sequence of launches (waves x (levels +
        leaf evaluation + backup)) whose arguments do not depend on the positions, so it is captured once per
        (network, iterations, exploration weight, progress, max_moves) and replayed every ply: for a handful of games
        the ~80 launches of a search cost more host time than device time.  The network's tiles are referenced by
        address, so weight updates (repacked in place by net_handle()) are picked up by the replay."""
        it = self.max_iterations if iterations is None else int(iterations)
        key = (id(neural_net), it, float(exploration_weight), float(training_progress), int(max_moves))
        if neural_net is not None:
            neural_net.net_handle()                  # repack (eagerly, in place) if the parameters changed
        g = self._graphs.get(key)
        if g is None:
            self.search(neural_net, it, exploration_weight, training_progress, max_moves)     # eager once: lazy attributes
            self.advance()
            torch.cuda.synchronize()
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.search(neural_net, it, exploration_weight, training_progress, max_moves)
                self.advance()
            self._graphs[key] = g
            return
        g.replay()


_forest_cache = {}


def _cached_forest(n_games, workers, samples_per_level, iterations):
    key = (torch.cuda.current_device(), n_games, workers, len(samples_per_level), samples_per_level[0], iterations)
    f = _forest_cache.get(key)
    if f is None:
        if len(_forest_cache) >= 4:
            _forest_cache.pop(next(iter(_forest_cache))).close()
        f = _forest_cache[key] = DeviceForest(n_games, workers, samples_per_level, iterations)
    return f


class BatchedRussianDollMCTS:
    """Constructor of ParallelRussianDollMCTS (mcts.py:15-31) plus a seed; search() runs on the device forest."""

    def __init__(self, board, iterations=20, exploration_weight=1.0, samples_per_level=None, num_workers=None,
                 rng=None, seed=None):
        self.root = as_records(board)[:1].copy()
        self.history, self.halfmove_clock, self.ply = (np.zeros(0, BOARD_DTYPE), 0, None)
        if hasattr(board, "occupied_co"):
            self.history, self.halfmove_clock, self.ply = history_records(board)
        self.iterations = iterations
        self.exploration_weight = exploration_weight
        self.samples_per_level = samples_per_level or [21, 13, 8, 5, 3, 2, 1]
        self.num_workers = num_workers or 8       # simulations per wave (mcts.py:20: min(8, cpu_count))
        self.rng = rng or _random.Random()
        self.seed = self.rng.getrandbits(64) if seed is None else int(seed)
        self.total_positions_evaluated = 0
        self.root_moves = self.root_visits = self.root_values = None     # statistics of the root after search()

    def search(self, neural_net=None, device=None, training_progress=0.0, current_move=0, max_moves=200,
               evaluator=None):
        """-> (best child record, (from, to, promotion)) or (None, None) when the root has no move."""
        net = neural_net if neural_net is not None else (evaluator.net if evaluator is not None else None)
        return search_many([self], net, training_progress, [current_move], max_moves)[0]


class ParallelRussianDollMCTS(BatchedRussianDollMCTS):
    """The reference's class name, constructor and return type (mcts.py:14-116), so that
    `from chess_deep_q_b200.mcts import ParallelRussianDollMCTS` is the import swap for neural_network.py:166-181:
    search() returns the move -- a chess.Move when the board is a python-chess board, else board_utils.Move -- or None."""

    def __init__(self, board, iterations=20, exploration_weight=1.0, samples_per_level=None, num_workers=None):
        super().__init__(board, iterations, exploration_weight, samples_per_level, num_workers)
        self.board = board

    def search(self, neural_net=None, device=None, training_progress=0.0, current_move=0, max_moves=200):
        from .board_utils import _make_move
        child, mv = super().search(neural_net, device, training_progress, current_move, max_moves)
        return None if child is None else _make_move(self.board, *mv)


def search_many(trees, neural_net, training_progress=0.0, current_moves=None, max_moves=200):
    """The searches of several trees (e.g. one per concurrent self-play game) as ONE device forest: every level of a
    wave is one kernel for all trees, every wave one cdq_leaf_eval.  A tree's result depends only on its own seed and
    position, so it is what `tree.search()` alone produces.  `neural_net` may be a LeafEvaluator (its net is used)."""
    if isinstance(neural_net, LeafEvaluator):
        neural_net = neural_net.net
    t0 = trees[0]
    assert all(t.num_workers == t0.num_workers and t.iterations == t0.iterations and
               list(t.samples_per_level) == list(t0.samples_per_level) and t.exploration_weight == t0.exploration_weight
               for t in trees), "trees searched together share workers, iterations, widths and exploration weight"
    current_moves = list(current_moves) if current_moves is not None else [0] * len(trees)
    forest = _cached_forest(len(trees), min(32, t0.num_workers), list(t0.samples_per_level), t0.iterations)
    plies = [t.ply if t.ply is not None else cm for t, cm in zip(trees, current_moves)]
    forest.set_games(np.concatenate([t.root for t in trees]), [t.seed for t in trees],
                     halfmove_clock=[t.halfmove_clock for t in trees], ply=plies, history=[t.history for t in trees])
    forest.search(neural_net, t0.iterations, t0.exploration_weight, training_progress, max_moves)
    forest.advance()
    best = forest.best_move.cpu().numpy().view(np.uint32)
    kids = forest.game_boards.cpu().numpy().view(np.uint64).reshape(-1).view(BOARD_DTYPE)
    nroot = forest.root_children.cpu().numpy()
    moves = forest.edge_move[:, 0].cpu().numpy().view(np.uint32)
    visits = forest.edge_visits[:, 0].cpu().numpy()
    values = forest.edge_value[:, 0].cpu().numpy()
    evaluated = forest.positions_evaluated.cpu().numpy()
    out = []
    for i, t in enumerate(trees):
        t.total_positions_evaluated += int(evaluated[i])
        c = int(nroot[i])
        t.root_moves, t.root_visits, t.root_values = moves[i, :c].copy(), visits[i, :c].copy(), values[i, :c].copy()
        if best[i] == 0xffffffff:
            out.append((None, None))
        else:
            out.append((kids[i:i + 1].copy(), _decode_move(best[i])))
    return out


def _startpos():
    start = np.zeros(1, BOARD_DTYPE)
    start["bb"][0] = [0x00ff00000000ff00, 0x4200000000000042, 0x2400000000000024, 0x8100000000000081,
                      0x0800000000000008, 0x1000000000000010, 0xffff, 0xffff000000000000]
    start["meta"][0] = 0x1f
    return start


def annealed_search_parameters(training_progress):
    """DQNAgent._get_mcts_move (neural_network.py:150-163): iterations, exploration weight and per-level widths."""
    f = 1 - 0.3 * training_progress
    return (int(50 + 150 * training_progress), max(1.6 * (1 - 0.5 * training_progress), 0.8),
            [max(21, int(25 * f)), max(13, int(15 * f)), max(8, int(10 * f)), max(5, int(7 * f)),
             max(3, int(5 * f)), max(2, int(3 * f)), max(1, int(2 * f))])


def self_play_games(neural_net, n_games, max_moves=200, iterations=50, samples_per_level=None, seeds=None,
                    training_progress=0.0, exploration_weight=1.0, num_workers=8, agent=None, train_every=1, start=None,
                    on_ply=None):
    """`n_games` self-play games (OptimizedChessAI.self_play_game, chess_ai.py:117-250) advanced ply by ply in lockstep
    on ONE device forest: per ply one search for all games (one host synchronisation), the moves are played on the
    device, and one batched evaluation of the new positions gives the display score, the reward and the terminal flags
    (chess_ai.py:161,174-184).  With `agent` (a DQNAgent) every played move is stored as a transition
    (chess_ai.py:187: state before, move, reward, state after, done) and the network is trained once per stored
    transition as soon as the replay memory holds a batch (chess_ai.py:190-192; train_every = n trains once per n
    transitions).  Game i uses seed seeds[i] and is move for move the game it plays alone.
    -> list of dicts(positions, scores, moves, rewards, result_flags, losses)."""
    from .evaluation import evaluate_batch
    seeds = list(seeds) if seeds is not None else list(range(42, 42 + n_games))
    widths = list(samples_per_level or [21, 13, 8, 5, 3, 2, 1])
    forest = _cached_forest(n_games, num_workers, widths, iterations)
    begin = np.concatenate([_startpos()] * n_games) if start is None else as_records(start)
    forest.set_games(begin, seeds)
    # per-game logs as arrays (the per-ply bookkeeping below is vectorised over the games)
    positions = np.zeros((n_games, max_moves + 1), BOARD_DTYPE)
    positions[:, 0] = begin
    scores = np.zeros((n_games, max_moves), np.float64)
    rewards = np.zeros((n_games, max_moves), np.float64)
    codes = np.zeros((n_games, max_moves), np.uint32)
    length = np.zeros(n_games, np.int64)
    flags_of = np.zeros(n_games, np.int64)
    losses = [[] for _ in range(n_games)]
    _, flags0 = evaluate_batch(begin)
    active = (flags0.cpu().numpy() & 0x70) == 0          # `while not self.board.is_game_over()` (chess_ai.py:141)
    pending = 0
    rows = np.arange(n_games)

    def book(active, rec, status, best, score_h, fl):
        """One ply's bookkeeping for all games (vectorised); -> the games still running."""
        nonlocal pending
        moved = active & ((status & 1) != 0)               # `if move is None: break` (chess_ai.py:148-149)
        g = rows[moved]
        k = length[g]
        positions[g, k + 1] = rec[g]
        scores[g, k] = score_h[g]                           # raw_score (chess_ai.py:161)
        codes[g, k] = best[g]
        mate, draw = (fl[g] & 0x10) != 0, (fl[g] & 0x60) != 0                # chess_ai.py:174-184
        rewards[g, k] = np.where(mate, -1.0, np.where(draw, 0.0, 0.01 * np.tanh(score_h[g] / 10.0)))
        flags_of[g] = fl[g] | (status[g].astype(np.int64) << 24)
        length[g] = k + 1
        if agent is not None:
            for gi, ki, done in zip(g, k, mate | draw):
                agent.store_transition(positions[gi, ki:ki + 1], _decode_move(codes[gi, ki]), float(rewards[gi, ki]), positions[gi, ki + 1:ki + 2], bool(done))
                pending += 1
                if len(agent.memory) >= agent.batch_size and pending >= train_every:
                    losses[gi].append(agent.train())
                    pending = 0
        over = ((fl & 0x70) != 0) | ((status & 6) != 0)     # is_game_over(): mate, stalemate, material, fivefold, 75 moves
        return moved & ~over

    if agent is not None or on_ply is not None:
        # training (or a callback) between the plies: the host needs every ply's result before the next search starts
        for ply in range(max_moves):
            if not active.any():
                break
            forest.set_active(active.astype(np.int32))
            forest.search_and_advance(neural_net, iterations, exploration_weight, training_progress, max_moves)
            score, flags = evaluate_batch(forest.game_boards)                    # same stream: after the moves were played
            rec = forest.game_boards.cpu().numpy().view(np.uint64).reshape(-1).view(BOARD_DTYPE)     # the ply's one synchronisation
            status = forest.game_status.cpu().numpy()
            best = forest.best_move.cpu().numpy().view(np.uint32)
            active = book(active, rec, status, best, score.cpu().numpy(), flags.cpu().numpy().astype(np.int64))
            if on_ply is not None:
                on_ply(ply, active)
    else:
        # Pure self-play: which games go on is decided ON THE DEVICE (the same two masks as in book()), so the next ply's
        # search is enqueued before the host has seen this ply's result; the result travels to pinned host memory behind the
        # search that produced it and is booked while the device runs the next one (1.1 of 6.0 ms per ply of 1 024 games were
        # host bookkeeping and five small synchronous copies).  The search enqueued after the last game ended finds every
        # game inactive and plays nothing.
        dev = forest.game_boards.device
        pin = getattr(forest, "_selfplay_pin", None)       # pinned landing buffers, two plies deep, kept with the (cached) forest
        if pin is None:
            pin = forest._selfplay_pin = [
                dict(rec=torch.empty((n_games, 9), dtype=torch.int64).pin_memory(), status=torch.empty(n_games, dtype=torch.int32).pin_memory(),
                     best=torch.empty(n_games, dtype=torch.int32).pin_memory(), score=torch.empty(n_games, dtype=torch.float64).pin_memory(),
                     flags=torch.empty(n_games, dtype=torch.int32).pin_memory(), done=torch.cuda.Event()) for _ in range(2)]
        active_d = torch.from_numpy(active.astype(np.int32)).to(dev)
        in_flight = None                                   # the ply whose result is on its way to the host
        for ply in range(max_moves + 1 if active.any() else 0):
            launched = None
            if ply < max_moves:
                _lib.check(_lib.lib().cdq_forest_set_active(forest._h, active_d.data_ptr(), _lib.stream_ptr()))
                forest.search_and_advance(neural_net, iterations, exploration_weight, training_progress, max_moves)
                score, flags = evaluate_batch(forest.game_boards)
                buf = pin[ply & 1]
                buf["rec"].copy_(forest.game_boards, non_blocking=True)
                buf["status"].copy_(forest.game_status, non_blocking=True)
                buf["best"].copy_(forest.best_move, non_blocking=True)
                buf["score"].copy_(score, non_blocking=True)
                buf["flags"].copy_(flags, non_blocking=True)
                buf["done"].record()
                moved_d = (active_d != 0) & ((forest.game_status & 1) != 0)
                over_d = ((flags & 0x70) != 0) | ((forest.game_status & 6) != 0)
                active_d = (moved_d & ~over_d).to(torch.int32)
                launched = buf
            if in_flight is not None:
                buf = in_flight
                buf["done"].synchronize()
                active = book(active, buf["rec"].numpy().view(np.uint64).reshape(-1).view(BOARD_DTYPE), buf["status"].numpy(),
                              buf["best"].numpy().view(np.uint32), buf["score"].numpy(), buf["flags"].numpy().astype(np.int64))
            in_flight = launched
            if in_flight is None or (ply > 0 and not active.any()):
                break
        if in_flight is not None:                          # the search enqueued after the last game ended: nothing to book
            in_flight["done"].synchronize()
    return [{"positions": positions[i, :length[i] + 1].copy(), "scores": scores[i, :length[i]].tolist(),
             "moves": [_decode_move(c) for c in codes[i, :length[i]]], "rewards": rewards[i, :length[i]].tolist(),
             "result_flags": int(flags_of[i]), "losses": losses[i]} for i in range(n_games)]


def self_play_game(neural_net, max_moves=200, iterations=50, samples_per_level=None, seed=42, start=None,
                   training_progress=0.0, **kw):
    """One self-play game (chess_ai.py:117-250): -> (positions, per-ply handcrafted scores)."""
    g = self_play_games(neural_net, 1, max_moves, iterations, samples_per_level, [seed], training_progress, start=start, **kw)[0]
    return g["positions"], g["scores"]
