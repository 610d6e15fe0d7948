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
from .board_utils import BOARD_DTYPE, as_records, planes_to_packed, to_device

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
        Also accepts packed boards (int64 [n,9])."""
        if isinstance(x, torch.Tensor) and x.dtype == torch.int64:
            return self.forward_packed(x)
        x = x.to(_lib.require_cuda()).reshape(-1, 12, 8, 8)
        return self.forward_packed(planes_to_packed(x))


class DQNAgent:
    """Reference constructor and attributes (neural_network.py:40-73).  The replay memory stores
    72-byte packed boards instead of 3 KB fp32 planes."""

    def __init__(self, gamma=0.99, epsilon=0.1, epsilon_min=0.1, epsilon_decay=0.995, learning_rate=0.001,
                 batch_size=64, use_aha_learning=False):
        self.gamma = gamma
        self.epsilon = epsilon
        self.epsilon_min = epsilon_min
        self.epsilon_decay = epsilon_decay
        self.learning_rate = learning_rate
        self.batch_size = batch_size
        self.device = _lib.require_cuda()
        self.use_aha_learning = use_aha_learning
        self.q_network = ChessQNetwork().to(self.device)
        self.target_q_network = ChessQNetwork().to(self.device)
        self.update_target_network()
        from .train import FusedAdam
        self.optimizer = FusedAdam(self.q_network, lr=learning_rate)
        self.memory = deque(maxlen=10000)

    def update_target_network(self):
        """neural_network.py:183-185"""
        self.target_q_network.load_state_dict(self.q_network.state_dict())

    def get_q_value(self, board):
        """neural_network.py:187-191"""
        with torch.no_grad():
            return float(self.q_network.forward_packed(board).item())

    def store_transition(self, state, move, reward, next_state, done):
        """neural_network.py:212-216; a falsy next_state becomes the all-zero board."""
        s = as_records(state)[0]
        ns = as_records(next_state)[0] if next_state is not None and next_state is not False else np.zeros(1, BOARD_DTYPE)[0]
        self.memory.append((s, move, float(reward), ns, float(done)))

    def train(self):
        """neural_network.py:218-252: sample, q=net(s), q'=target(s'), y=r+(1-done)*gamma*q',
        MSE, Adam step, epsilon decay; returns the loss as a float."""
        if len(self.memory) < self.batch_size:
            return 0
        minibatch = random.sample(self.memory, self.batch_size)
        states = np.array([m[0] for m in minibatch], dtype=BOARD_DTYPE)
        next_states = np.array([m[3] for m in minibatch], dtype=BOARD_DTYPE)
        rewards = torch.tensor([m[2] for m in minibatch], dtype=torch.float32)
        dones = torch.tensor([m[4] for m in minibatch], dtype=torch.float32)
        from .train import dqn_train_step
        loss = dqn_train_step(self.q_network, self.target_q_network, self.optimizer, to_device(states),
                              to_device(next_states), rewards.to(self.device), dones.to(self.device), self.gamma)
        if self.epsilon > self.epsilon_min:
            self.epsilon *= self.epsilon_decay
        return float(loss.item())
