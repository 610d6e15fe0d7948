"""Q-network and DQN agent: mirrors the reference's neural_network.py (ChessQNetwork :16-37,
DQNAgent :39-73,183-252) with the forward pass (and the training step, see train.py) running in
the hand-written CUDA kernels of csrc/cnn_kernels.cu behind the C ABI."""
import ctypes
import random
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .board_utils import BOARD_DTYPE, Move, _make_move, as_records, planes_to_packed, to_device

_PARAM_ORDER = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
                "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")


class _Workspace:
    """Grow-only device scratch for cdq_value_fwd / cdq_leaf_eval."""

    def __init__(self):
        self.buf = None

    def get(self, n, device):
        need = _lib.lib().cdq_workspace_bytes(n)
        if self.buf is None or self.buf.numel() < need or self.buf.device != device:
            self.buf = torch.empty(max(need, 16), dtype=torch.uint8, device=device)
        return self.buf


class ChessQNetwork(nn.Module):
    """Same layers / state_dict keys as the reference (neural_network.py:16-27):
    conv3x3(12->32) ReLU conv3x3(32->64) ReLU flatten fc(4096->256) ReLU fc(256->1) tanh.
    The nn.Conv2d / nn.Linear members only hold the fp32 master parameters."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(12, 32, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.fc1 = nn.Linear(64 * 8 * 8, 256)
        self.fc2 = nn.Linear(256, 1)
        self._handle = None
        self._packed_key = None
        self._ws = _Workspace()

    # ------------------------------------------------------------------ packed-weight handle
    def _params(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in _PARAM_ORDER]

    def net_handle(self):
        """CdqNet* with the current parameters repacked to fp16 UMMA tiles (re-packed whenever a
        parameter was modified in place or replaced)."""
        dev = _lib.require_cuda()
        ps = self._params()
        if any(p.device.type != "cuda" for p in ps):
            raise _lib.CdqError("ChessQNetwork parameters must live on the CUDA device (.to('cuda'))")
        key = tuple((p.data_ptr(), p._version) for p in ps)
        L = _lib.lib()
        if self._handle is None:
            h = ctypes.c_void_p()
            _lib.check(L.cdq_net_create(ctypes.byref(h)))
            self._handle = h
        if key != self._packed_key:
            w = _lib.CdqWeights(*[p.detach().contiguous().data_ptr() for p in ps])
            _lib.check(L.cdq_net_load(self._handle, ctypes.byref(w), _lib.stream_ptr()))
            self._packed_key = key
        return self._handle

    def raw_handle(self):
        """The CdqNet* without the is-it-current check: for callers that re-pack inside their own
        call (cdq_train_fwd_bwd with CDQ_TRAIN_REPACK)."""
        return self._handle if self._handle is not None else self.net_handle()

    def invalidate_packed(self):
        """Force a repack on the next forward (used after the fused optimizer wrote the flat
        parameter buffer directly, which autograd's version counters do not see)."""
        self._packed_key = None

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.lib().cdq_net_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def forward_packed(self, boards):
        """boards: packed records (anything board_utils.to_device accepts) -> [n,1] fp32 values."""
        d = to_device(boards)
        n = d.shape[0]
        out = torch.empty((n, 1), dtype=torch.float32, device=d.device)
        if n == 0:
            return out
        ws = self._ws.get(n, d.device)
        _lib.check(_lib.lib().cdq_value_fwd(self.net_handle(), d.data_ptr(), n, out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return out

    def forward(self, x):
        """Reference signature: x = one-hot planes [B,12,8,8] (board_to_tensor output) -> [B,1].
        Also accepts packed boards (int64 [n,9]).

        The kernels convolve BOARDS, not arbitrary images: planes that are not exactly one-hot
        (values other than 0/1, two pieces on one square) raise instead of being thresholded.
        With grad enabled and trainable parameters the result carries a grad_fn
        (_ValueFunction), so reference-style `loss.backward()` (neural_network.py:244-246) fills
        the parameters' .grad; there is no gradient with respect to the planes (the reference
        never asks for one) and a planes tensor that requires grad is refused."""
        if isinstance(x, torch.Tensor) and x.dtype == torch.int64:
            packed = to_device(x)
        else:
            if x.requires_grad:
                raise _lib.CdqError("ChessQNetwork.forward: no gradient with respect to the input planes "
                                    "(the CUDA path convolves packed boards); detach the input")
            x = x.to(_lib.require_cuda()).reshape(-1, 12, 8, 8)
            one_hot = ((x == 0) | (x == 1)).all() & (x.sum(1) <= 1).all()
            if not bool(one_hot.item()):
                raise ValueError("ChessQNetwork.forward expects one-hot board planes as board_to_tensor produces them "
                                 "(board_utils.py:12-40): every entry 0.0 or 1.0, at most one piece per square")
            packed = planes_to_packed(x)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _ValueFunction.apply(packed, self, *self._params())
        return self.forward_packed(packed)


class _ValueFunction(torch.autograd.Function):
    """Autograd surface of the CUDA forward.  forward = cdq_value_fwd; backward =
    cdq_train_fwd_bwd in its CDQ_LOSS_UPSTREAM mode: the forward is recomputed keeping activations
    and the incoming d loss / d value is pushed through tanh, fc2, fc1, conv2 and conv1 by the
    training kernels.  Gradients travel between those kernels as fp16, so the upstream gradient is
    normalised to max |.| = 1 on the way in and the parameter gradients are rescaled on the way out."""

    @staticmethod
    def forward(ctx, packed, module, *params):
        ctx.module = module
        ctx.save_for_backward(packed)
        with torch.no_grad():
            return module.forward_packed(packed)

    @staticmethod
    def backward(ctx, grad_out):
        from . import train as T
        module = ctx.module
        module._cdq_grads_dirty = True
        (packed,) = ctx.saved_tensors
        n = packed.shape[0]
        dev = packed.device
        params = module._params()
        grads = torch.zeros(T.NPARAMS_PAD, dtype=torch.float32, device=dev)
        if n:
            g = grad_out.reshape(-1).to(torch.float32).contiguous()
            gmax = g.abs().max().clamp_min(1e-30)
            g_unit = (g / gmax).contiguous()
            views = [grads[o:o + s].view(sh) for o, s, sh in zip(T.PARAM_OFFSETS, T.PARAM_SIZES, T.PARAM_SHAPES)]
            ws = torch.zeros(_lib.lib().cdq_train_workspace_bytes(n), dtype=torch.uint8, device=dev)
            scratch_loss = torch.zeros(1, dtype=torch.float32, device=dev)
            w = _lib.CdqWeights(*[p.detach().contiguous().data_ptr() for p in params])
            gw = _lib.CdqWeights(*[v.data_ptr() for v in views])
            _lib.check(_lib.lib().cdq_train_fwd_bwd(
                module.net_handle(), None, ctypes.byref(w), packed.data_ptr(), None, g_unit.data_ptr(), None, n, 0.0, 0.0,
                T.LOSS_UPSTREAM | T.WS_PERSISTENT, ctypes.byref(gw), scratch_loss.data_ptr(), ws.data_ptr(), ws.numel(),
                _lib.stream_ptr()))
            grads = grads * gmax
        out = [grads[o:o + s].view(sh) if p.requires_grad else None
               for p, o, s, sh in zip(params, T.PARAM_OFFSETS, T.PARAM_SIZES, T.PARAM_SHAPES)]
        return (None, None, *out)


class ReplayMemory:
    """`deque(maxlen=10000)` of the reference (neural_network.py:66) as a ring of numpy arrays: a
    transition is two 72-byte records + reward + done (+ the move object) instead of two 3 KB fp32
    tensors, and a minibatch is gathered with four fancy-index copies straight into pinned staging
    buffers.  Supports what the reference does with its deque: len(), append(tuple), iteration,
    indexing, random.sample()."""

    def __init__(self, maxlen=10000):
        self.maxlen = int(maxlen)
        self.states = np.zeros(self.maxlen, BOARD_DTYPE)
        self.next_states = np.zeros(self.maxlen, BOARD_DTYPE)
        self.rewards = np.zeros(self.maxlen, np.float32)
        self.dones = np.zeros(self.maxlen, np.float32)
        self.moves = [None] * self.maxlen
        self._start, self._len = 0, 0
        self.appended = 0          # total number of append() calls: lets a device mirror upload only what is new

    def __len__(self):
        return self._len

    def append(self, item):
        s, move, reward, ns, done = item
        slot = (self._start + self._len) % self.maxlen
        if self._len == self.maxlen:
            self._start = (self._start + 1) % self.maxlen       # deque(maxlen): the oldest entry falls out
        else:
            self._len += 1
        self.states[slot], self.next_states[slot] = s, ns
        self.rewards[slot], self.dones[slot], self.moves[slot] = reward, done, move
        self.appended += 1
        return slot

    def _slot(self, i):
        if i < 0:
            i += self._len
        if not 0 <= i < self._len:
            raise IndexError("replay memory index out of range")
        return (self._start + i) % self.maxlen

    def __getitem__(self, i):
        k = self._slot(i)
        return (self.states[k], self.moves[k], float(self.rewards[k]), self.next_states[k], float(self.dones[k]))

    def __iter__(self):
        return (self[i] for i in range(self._len))

    def clear(self):
        self._start = self._len = 0

    def slots(self, indices):
        return (np.asarray(indices, dtype=np.int64) + self._start) % self.maxlen


class DQNAgent:
    """Reference constructor and attributes (neural_network.py:40-73).  The replay memory stores
    72-byte packed boards instead of 3 KB fp32 planes; train() is ONE CUDA graph
    (train.GraphedDQNStep) fed from pinned staging buffers."""

    def __init__(self, gamma=0.99, epsilon=0.1, epsilon_min=0.1, epsilon_decay=0.995, learning_rate=0.001,
                 batch_size=64, use_aha_learning=False):
        self.gamma = gamma
        self.epsilon = epsilon
        self.epsilon_min = epsilon_min
        self.epsilon_decay = epsilon_decay
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.device = _lib.require_cuda()
        # AHA learning parameters (neural_network.py:52-56)
        self.use_aha_learning = use_aha_learning
        self.aha_budget_per_game = 3
        self.aha_budget_remaining = 3
        self.aha_threshold = -1.5
        self.q_network = ChessQNetwork().to(self.device)
        self.target_q_network = ChessQNetwork().to(self.device)
        self.update_target_network()
        from .train import FusedAdam
        self.optimizer = FusedAdam(self.q_network, lr=learning_rate)
        self.memory = ReplayMemory(maxlen=10000)
        import multiprocessing
        self.num_cpu_cores = max(1, multiprocessing.cpu_count() - 1)     # neural_network.py:69 (simulations per wave)
        self._graph_step = None
        self._dmem = None

    def update_target_network(self):
        """neural_network.py:183-185"""
        self.target_q_network.load_state_dict(self.q_network.state_dict())

    def get_q_value(self, board):
        """neural_network.py:187-191"""
        with torch.no_grad():
            return float(self.q_network.forward_packed(board).item())

    # ------------------------------------------------------------------ move selection (neural_network.py:75-210)
    def _get_mcts_move(self, board, training_progress, current_move, max_moves):
        """neural_network.py:148-181: iterations 50 -> 200 and the per-level widths annealed with the
        training progress, then one Russian Doll search with the Q-network as the leaf evaluator."""
        from .mcts import BatchedRussianDollMCTS
        base_iterations = int(50 + 150 * training_progress)
        exploration_weight = max(1.6 * (1 - 0.5 * training_progress), 0.8)
        f = 1 - 0.3 * training_progress
        samples_per_level = [max(21, int(25 * f)), max(13, int(15 * f)), max(8, int(10 * f)), max(5, int(7 * f)),
                             max(3, int(5 * f)), max(2, int(3 * f)), max(1, int(2 * f))]
        tree = BatchedRussianDollMCTS(board, iterations=base_iterations, exploration_weight=exploration_weight,
                                      samples_per_level=samples_per_level, num_workers=self.num_cpu_cores)
        child, mv = tree.search(neural_net=self.q_network, device=self.device, training_progress=training_progress,
                                current_move=current_move, max_moves=max_moves)
        return None if child is None else _make_move(board, *mv)

    def select_move(self, board, training_progress=0.0, is_training=False, current_move=0, max_moves=200):
        """neural_network.py:193-206"""
        if (self.use_aha_learning and is_training and self.aha_budget_remaining > 0 and current_move > 5):
            return self.select_move_with_aha_learning(board, training_progress, is_training,
                                                      self.aha_budget_remaining, self.aha_threshold)
        return self._get_mcts_move(board, training_progress, current_move, max_moves)

    def select_move_with_aha_learning(self, board, training_progress=0.0, is_training=True, undo_budget=3,
                                      eval_threshold=-1.5):
        """neural_network.py:75-146.  The look-ahead evaluations of the reference (one
        fast_evaluate_position per legal move) are ONE cdq_expand + ONE cdq_eval_terms here; the
        immediate negative-feedback update is the same autograd sequence (zero_grad, forward,
        mse_loss against -1, backward, step) through the CUDA kernels."""
        from .evaluation import evaluate_batch
        from .mcts import move_from_records
        if not is_training or undo_budget <= 0:
            return self._get_mcts_move(board, training_progress, 0, 200)
        rec = as_records(board)[:1]
        white = bool(int(rec["meta"][0]) & 1)
        initial_move = self._get_mcts_move(board, training_progress, 0, 200)
        if initial_move is None:
            return None
        d = to_device(rec)
        kids = torch.empty((1, 218, 9), dtype=torch.int64, device=d.device)
        counts = torch.empty(1, dtype=torch.int32, device=d.device)
        _lib.check(_lib.lib().cdq_expand(d.data_ptr(), 1, 218, kids.data_ptr(), counts.data_ptr(), _lib.stream_ptr()))
        c = int(counts.item())
        both = torch.cat([d, kids[0, :c]])
        score = evaluate_batch(both)[0].cpu().numpy()
        current_eval, child_eval = float(score[0]), score[1:]
        kids_h = kids[0, :c].cpu().numpy().view(np.uint64).reshape(-1).view(BOARD_DTYPE)
        moves = [move_from_records(rec[0], k) for k in kids_h]
        ucis = [Move(*m).uci() for m in moves]
        change = (current_eval - child_eval) if white else (child_eval - current_eval)
        k0 = ucis.index(initial_move.uci())
        if change[k0] >= eval_threshold:
            return initial_move
        print("AHA! Detected potential mistake (eval drop: %.2f)" % change[k0])
        self.optimizer.zero_grad()
        q_value = self.q_network(d)
        loss = torch.nn.functional.mse_loss(q_value, torch.full_like(q_value, -1.0))
        loss.backward()
        self.optimizer.step()
        better = [(k, change[k]) for k in range(c) if k != k0 and change[k] >= eval_threshold]
        if not better:
            print("No better alternatives found, keeping original move")
            return initial_move
        best = max(better, key=lambda kv: kv[1])[0]          # first of the best, like the reference's stable sort
        self.aha_budget_remaining -= 1
        alt = _make_move(board, *moves[best])
        print("AHA moment used! Corrected %s -> %s. Remaining budget: %d" % (initial_move.uci(), alt.uci(),
                                                                             self.aha_budget_remaining))
        return alt

    # ------------------------------------------------------------------ replay (neural_network.py:212-252)
    def store_transition(self, state, move, reward, next_state, done):
        """neural_network.py:212-216; a falsy next_state becomes the all-zero board."""
        s = as_records(state)[0]
        ns = as_records(next_state)[0] if next_state is not None and next_state is not False else np.zeros(1, BOARD_DTYPE)[0]
        self.memory.append((s, move, float(reward), ns, float(done)))

    def _device_memory(self):
        """Device mirror of the replay memory (four arrays of maxlen entries: 1.5 MB; ring geometry and draw counter next to
        them), brought up to date by cdq_replay_append with the transitions stored since the last call (in self-play: one per
        training step, chess_ai.py:187-192).  The minibatch is drawn and gathered from it by cdq_replay_sample inside the
        training step's CUDA graph."""
        mem = self.memory
        if self._dmem is None or self._dmem["maxlen"] != mem.maxlen or self._dmem["memory"] is not mem:
            dev, n = self.device, mem.maxlen
            self._dmem = {"maxlen": n, "memory": mem, "synced": 0,
                          "states": torch.zeros((n, 9), dtype=torch.int64, device=dev),
                          "next_states": torch.zeros((n, 9), dtype=torch.int64, device=dev),
                          "rewards": torch.zeros(n, dtype=torch.float32, device=dev),
                          "dones": torch.zeros(n, dtype=torch.float32, device=dev),
                          "ring": torch.zeros(4, dtype=torch.int32, device=dev),
                          "counter": torch.zeros(1, dtype=torch.int32, device=dev),
                          "seed": random.getrandbits(63),          # random.seed() makes the draws reproducible
                          "loss": torch.empty(1, dtype=torch.float32).pin_memory()}
            self._graph_step = None                               # the graph holds the old mirror's addresses
        d = self._dmem
        new = min(mem.appended - d["synced"], mem._len)
        n = mem.maxlen
        first = (mem._start + mem._len - new) % n                   # slots [first, first + new) modulo maxlen
        head = min(new, n - first)
        L, st = _lib.lib(), _lib.stream_ptr()
        segments = [(lo, cnt) for lo, cnt in ((first, head), (0, new - head)) if cnt > 0] or [(0, 0)]   # (0, 0): geometry only
        for lo, cnt in segments:
            _lib.check(L.cdq_replay_append(d["states"].data_ptr(), d["next_states"].data_ptr(), d["rewards"].data_ptr(),
                                           d["dones"].data_ptr(), d["ring"].data_ptr(), mem.states[lo:].ctypes.data,
                                           mem.next_states[lo:].ctypes.data, mem.rewards[lo:].ctypes.data, mem.dones[lo:].ctypes.data,
                                           lo, cnt, mem._start, mem._len, n, st))
        d["synced"] = mem.appended
        return d

    def sampled_indices(self, draw):
        """Logical indices (0 = oldest entry) of the minibatch of draw number `draw` (0 = this agent's first train() call),
        as cdq_replay_sample computes them on the device: for tests and reproducibility."""
        L, d = _lib.lib(), self._dmem
        return [int(L.cdq_replay_permute(k, len(self.memory), d["seed"], draw)) for k in range(self.batch_size)]

    def train(self):
        """neural_network.py:218-252: sample, q=net(s), q'=target(s'), y=r+(1-done)*gamma*q',
        MSE, Adam step, epsilon decay; returns the loss as a float.  Transitions stored since the last call reach the device
        mirror of the replay memory with one small copy each; drawing the minibatch (uniform, without replacement, like
        random.sample: neural_network.py:224), gathering it, the whole step and the optimizer are ONE CUDA graph replay, and
        the loss comes back with one 4-byte read."""
        if len(self.memory) < self.batch_size:
            return 0
        from .train import GraphedDQNStep
        d = self._device_memory()
        if self._graph_step is None or self._graph_step.batch != self.batch_size or self._graph_step.gamma != float(self.gamma):
            self._graph_step = GraphedDQNStep(self.q_network, self.target_q_network, self.optimizer,
                                              batch=self.batch_size, gamma=self.gamma, replay=d)
        loss = self._graph_step.run()
        st = d
        st["loss"].copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if self.epsilon > self.epsilon_min:
            self.epsilon *= self.epsilon_decay
        return float(st["loss"].item())
