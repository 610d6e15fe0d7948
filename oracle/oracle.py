"""CPU ORACLE bindings -- test infrastructure, NOT product code.

ctypes wrapper around oracle/_build/libcdq_oracle.so (cdq_oracle.c) plus a numpy/torch-CPU
fp32 restatement of the reference network (neural_network.py:16-37) and DQN step
(neural_network.py:218-252).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcdq_oracle.so")

# numpy view of include/cdq.h::CdqBoard (72 bytes)
BOARD_DTYPE = np.dtype([("bb", "<u8", (8,)), ("meta", "<u8")])
assert BOARD_DTYPE.itemsize == 72
NTERMS = 24

TERM_NAMES = [
    "w_mat", "b_mat", "w_mob", "b_mob", "w_att", "b_att", "w_katt", "b_katt",
    "w_castled", "b_castled", "w_kcentre", "b_kcentre", "w_dbl", "b_dbl", "w_iso", "b_iso",
    "w_centre", "b_centre", "w_ext", "b_ext", "w_def", "b_def", "check", "terminal",
]


def build(force=False):
    """Compile the C oracle (gcc, seconds)."""
    src = os.path.join(_HERE, "cdq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, i64, i32, u64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64
        L.cdq_oracle_startpos.argtypes = [vp]
        L.cdq_oracle_from_fen.argtypes = [ctypes.c_char_p, vp]
        L.cdq_oracle_from_fen.restype = i32
        L.cdq_oracle_perft.argtypes = [vp, i32]
        L.cdq_oracle_perft.restype = u64
        L.cdq_oracle_legal_moves.argtypes = [vp, i32, vp]
        L.cdq_oracle_legal_moves.restype = i32
        L.cdq_oracle_push.argtypes = [vp, ctypes.c_uint16]
        L.cdq_oracle_eval.argtypes = [vp, i64, i32, vp, vp, vp, i32]
        L.cdq_oracle_encode_planes.argtypes = [vp, i64, vp]
        L.cdq_oracle_playouts.argtypes = [u64, i64, i32, vp]
        L.cdq_oracle_maps.argtypes = [vp, i64, vp]
        L.cdq_oracle_categorize.argtypes = [vp, vp, vp]
        L.cdq_oracle_categorize.restype = i32
        L.cdq_oracle_init()
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def startpos():
    b = np.zeros(1, BOARD_DTYPE)
    lib().cdq_oracle_startpos(_ptr(b))
    return b


def from_fen(fen):
    b = np.zeros(1, BOARD_DTYPE)
    if lib().cdq_oracle_from_fen(fen.encode(), _ptr(b)) != 0:
        raise ValueError("bad FEN: %r" % fen)
    return b


def from_fens(fens):
    return np.concatenate([from_fen(f) for f in fens]) if len(fens) else np.zeros(0, BOARD_DTYPE)


def perft(board, depth):
    return int(lib().cdq_oracle_perft(_ptr(np.ascontiguousarray(board)), depth))


def legal_moves(board, force_turn=-1):
    mv = np.zeros(512, np.uint16)
    n = lib().cdq_oracle_legal_moves(_ptr(np.ascontiguousarray(board)), force_turn, _ptr(mv))
    return mv[:n].copy()


def push(board, move):
    b = np.ascontiguousarray(board).copy()
    lib().cdq_oracle_push(_ptr(b), int(move))
    return b


def playouts(n, seed=42, max_plies=199):
    """Synthetic random-playout positions (SURVEY.md 8d recipe)."""
    out = np.zeros(n, BOARD_DTYPE)
    lib().cdq_oracle_playouts(seed, n, max_plies, _ptr(out))
    return out


def evaluate(boards, ignore_checkmate=False, nthreads=1):
    """-> (terms int32 [n,24], score f64 [n], flags u32 [n]) per evaluation.py:30-229."""
    boards = np.ascontiguousarray(boards)
    n = len(boards)
    terms = np.zeros((n, NTERMS), np.int32)
    score = np.zeros(n, np.float64)
    flags = np.zeros(n, np.uint32)
    lib().cdq_oracle_eval(_ptr(boards), n, int(bool(ignore_checkmate)), _ptr(terms), _ptr(score),
                          _ptr(flags), nthreads)
    return terms, score, flags


def encode_planes(boards):
    """board_to_tensor for a batch (board_utils.py:12-40): float32 [n,12,8,8]."""
    boards = np.ascontiguousarray(boards)
    out = np.zeros((len(boards), 12, 8, 8), np.float32)
    lib().cdq_oracle_encode_planes(_ptr(boards), len(boards), _ptr(out))
    return out


def categorize(board):
    """categorize_moves (evaluation.py:232-339) -> (moves uint16 [k], sampling weights int32 [k])."""
    mv = np.zeros(512, np.uint16)
    w = np.zeros(512, np.int32)
    k = lib().cdq_oracle_categorize(_ptr(np.ascontiguousarray(board)), _ptr(mv), _ptr(w))
    return mv[:k].copy(), w[:k].copy()


def attack_maps(boards):
    """evaluation.py:368-428 -> uint64 [n,4]: attacked by white / black, legal destinations of white / black."""
    boards = np.ascontiguousarray(boards)
    out = np.zeros((len(boards), 4), np.uint64)
    lib().cdq_oracle_maps(_ptr(boards), len(boards), _ptr(out))
    return out


def leaf_values(boards, flags, score=None, net_value=None):
    """_evaluate_position (mcts.py:196-232) for a batch: checkmate -> -1.0 if white is to move else +1.0
    (:202-203); stalemate / insufficient material / threefold repetition -> 0.0 (:205-210); else
    clamp(net value, -1, 1) (:213-221) or, without a network, math.tanh(score / 10) (:226-229).
    `flags` as cdq.h (low three bits = CDQ_TERM_*), turn from the boards' meta bit 0."""
    import math
    flags = np.asarray(flags).astype(np.uint32)
    term = flags & 7
    white = (np.asarray(boards)["meta"] & 1).astype(bool)
    if net_value is not None:
        v = np.clip(np.asarray(net_value, np.float32).reshape(-1), -1.0, 1.0).astype(np.float64)
    else:
        v = np.array([math.tanh(float(x) / 10.0) for x in np.asarray(score)], np.float64)
    v = np.where(term == 1, np.where(white, -1.0, 1.0), np.where(term != 0, 0.0, v))
    return v


# ------------------------------------------------------------------ network oracle (fp32, CPU)
PARAM_SHAPES = [
    ("conv1.weight", (32, 12, 3, 3)), ("conv1.bias", (32,)),
    ("conv2.weight", (64, 32, 3, 3)), ("conv2.bias", (64,)),
    ("fc1.weight", (256, 4096)), ("fc1.bias", (256,)),
    ("fc2.weight", (1, 256)), ("fc2.bias", (1,)),
]


def init_params(seed=42, scale=1.0):
    """Parameters distributed like torch's default Conv2d/Linear init (kaiming_uniform a=sqrt(5)
    => U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias), from a numpy RNG so the same
    values exist on any box.  `scale` multiplies the weight matrices ("trained-like" ranges)."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, shape in PARAM_SHAPES:
        layer = name.split(".")[0]
        fan_in = {"conv1": 12 * 9, "conv2": 32 * 9, "fc1": 4096, "fc2": 256}[layer]
        bound = 1.0 / np.sqrt(fan_in)
        a = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        if name.endswith("weight"):
            a = (a * scale).astype(np.float32)
        p[name] = a
    return p


def net_forward(params, planes, return_acts=False):
    """ChessQNetwork.forward (neural_network.py:29-37) in fp32 on torch-CPU."""
    import torch
    import torch.nn.functional as F
    x = torch.as_tensor(planes, dtype=torch.float32)
    t = {k: torch.as_tensor(v) for k, v in params.items()}
    a1 = F.relu(F.conv2d(x, t["conv1.weight"], t["conv1.bias"], padding=1))
    a2 = F.relu(F.conv2d(a1, t["conv2.weight"], t["conv2.bias"], padding=1))
    h = F.relu(F.linear(a2.reshape(-1, 4096), t["fc1.weight"], t["fc1.bias"]))
    out = torch.tanh(F.linear(h, t["fc2.weight"], t["fc2.bias"]))
    if return_acts:
        return out.numpy(), a1.numpy(), a2.numpy(), h.numpy()
    return out.numpy()


def _fp16_emulation():
    """Autograd helpers that restate WHERE the CUDA training path rounds to fp16:
    rq = round an operand (weights, stored activations) to fp16 in the forward pass, identity backward;
    gq(x, scale) = identity forward, the gradient flowing back through x is rounded to fp16 after being
    multiplied by `scale` (gradients travel between the kernels as fp16 scaled by 64 n)."""
    import torch

    class RQ(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.half().float()

        @staticmethod
        def backward(ctx, g):
            return g

    class GQ(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, scale):
            ctx.scale = scale
            return x.view_as(x)

        @staticmethod
        def backward(ctx, g):
            return (g * ctx.scale).half().float() / ctx.scale, None

    return RQ.apply, GQ.apply


def dqn_step_reference(params, target_params, states_planes, next_planes, reward, done, gamma=0.99, huber=False,
                       fp16_operands=False, upstream=None):
    """Restatement of DQNAgent.train()'s maths (neural_network.py:232-245) with fp32 torch-CPU
    autograd: returns (loss float, {name: gradient ndarray}).

    fp16_operands=True restates the CUDA path's arithmetic instead of the reference's: conv / fc1
    weights and conv1's bias, ReLU(conv1) and ReLU(conv2) rounded to fp16 where the kernels store them
    as fp16 operands, fp32 accumulation, and the two inter-kernel gradients (fc1's and conv2's
    pre-activation gradients) rounded to fp16 at scale 64 n.  What is left between this and the kernels
    is fp32 summation order, so gradients can be compared element by element at ~1e-3 relative.
    upstream: instead of the DQN loss, backpropagate sum(q * upstream) (CDQ_LOSS_UPSTREAM)."""
    import torch
    import torch.nn.functional as F
    n = len(states_planes)
    rq, gq = _fp16_emulation()
    ident = lambda x, *a: x  # noqa: E731
    r_, g_ = (rq, gq) if fp16_operands else (ident, ident)
    scale = 64.0 * (1.0 if upstream is not None else float(n))

    def fwd(t, x, train):
        z1 = F.conv2d(x, r_(t["conv1.weight"]), r_(t["conv1.bias"]), padding=1)
        a1 = r_(F.relu(z1))
        z2 = F.conv2d(a1, r_(t["conv2.weight"]), t["conv2.bias"], padding=1)
        if train:
            z2 = g_(z2, scale)
        a2 = r_(F.relu(z2))
        z3 = F.linear(a2.reshape(-1, 4096), r_(t["fc1.weight"]))
        if train:
            z3 = g_(z3, scale)      # fc1's bias gradient is summed from the unrounded values (head_bwd_kernel), conv2's from the fp16 boards
        h = F.relu(z3 + t["fc1.bias"])
        return torch.tanh(F.linear(h, t["fc2.weight"], t["fc2.bias"]))

    t = {k: torch.tensor(v, requires_grad=True) for k, v in params.items()}
    q = fwd(t, torch.as_tensor(states_planes), True).reshape(-1)
    if upstream is not None:
        loss = (q * torch.as_tensor(upstream, dtype=torch.float32)).sum()
    else:
        tt = {k: torch.tensor(v) for k, v in target_params.items()}
        with torch.no_grad():
            qn = fwd(tt, torch.as_tensor(next_planes), False).reshape(-1)
        y = torch.as_tensor(reward) + (1 - torch.as_tensor(done)) * gamma * qn
        loss = F.huber_loss(q, y) if huber else F.mse_loss(q, y)
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy().copy() for k, v in t.items()}


def adam_reference(params, grads_seq, lr=1e-3):
    """torch.optim.Adam (the reference's optimizer, neural_network.py:65) applied to a sequence of
    gradient dicts; returns the updated parameters."""
    import torch
    t = {k: torch.nn.Parameter(torch.tensor(v)) for k, v in params.items()}
    opt = torch.optim.Adam(list(t.values()), lr=lr)
    for grads in grads_seq:
        for k, p in t.items():
            p.grad = torch.tensor(grads[k])
        opt.step()
    return {k: v.detach().numpy().copy() for k, v in t.items()}
