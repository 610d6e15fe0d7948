"""DQN replay training step: mirrors DQNAgent.train (neural_network.py:218-252) on the CUDA
kernels of csrc/train_kernels.cu.  Data-parallel variant: every rank runs the step on its shard
of the batch, gradients are summed with ONE all-reduce of the flat 1 071 073-float buffer over
NCCL and the 1/world scale is folded into the fused Adam kernel."""
import ctypes

import torch

from . import _lib

PARAM_ORDER = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias",
               "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias")
PARAM_SHAPES = ((32, 12, 3, 3), (32,), (64, 32, 3, 3), (64,), (256, 4096), (256,), (1, 256), (1,))
PARAM_SIZES = tuple(int(torch.Size(s).numel()) for s in PARAM_SHAPES)
PARAM_OFFSETS = tuple(sum(PARAM_SIZES[:i]) for i in range(len(PARAM_SIZES)))
NPARAMS = sum(PARAM_SIZES)
assert NPARAMS == 1071073

LOSS_MSE, LOSS_HUBER = 0, 1


def shard_range(n, rank, world):
    """Contiguous slice of a batch of n samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_scale(flat_grad, group=None):
    """Sum the flat gradient over the data-parallel group (one collective) and return the scale
    (1/world) that the optimizer applies; no-op outside torch.distributed."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def flatten_parameters(network):
    """Re-home the network's eight parameters as views into one contiguous fp32 buffer (so the
    gradient all-reduce and the Adam update are one call each).  Returns the flat buffer."""
    params = dict(network.named_parameters())
    first = params[PARAM_ORDER[0]]
    flat = torch.empty(NPARAMS, dtype=torch.float32, device=first.device)
    for name, off, size, shape in zip(PARAM_ORDER, PARAM_OFFSETS, PARAM_SIZES, PARAM_SHAPES):
        p = params[name]
        assert tuple(p.shape) == shape, (name, tuple(p.shape))
        view = flat[off:off + size].view(shape)
        view.copy_(p.data)
        p.data = view
    return flat


class FusedAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8), as the reference constructs it
    (neural_network.py:65), as one fused kernel over the flat parameter buffer.
    state_dict()/load_state_dict() use torch.optim.Adam's format so reference checkpoints
    (chess_ai.py:459,483) round-trip."""

    def __init__(self, network, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        _lib.require_cuda()
        self.network = network
        self.flat = flatten_parameters(network)
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False,
                              "params": list(range(len(PARAM_ORDER)))}]

    def zero_grad(self, set_to_none=False):
        self.grad.zero_()

    def grad_views(self):
        return {n: self.grad[o:o + s].view(sh) for n, o, s, sh in zip(PARAM_ORDER, PARAM_OFFSETS, PARAM_SIZES, PARAM_SHAPES)}

    def step(self, grad_scale=1.0):
        g = self.param_groups[0]
        self.step_count += 1
        _lib.check(_lib.lib().cdq_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), NPARAMS, self.step_count, g["lr"],
                                            g["betas"][0], g["betas"][1], g["eps"], grad_scale, _lib.stream_ptr()))
        self.network.invalidate_packed()   # parameters changed behind autograd's back: repack before next forward

    def state_dict(self):
        state = {}
        if self.step_count:
            for i, (o, s, sh) in enumerate(zip(PARAM_OFFSETS, PARAM_SIZES, PARAM_SHAPES)):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + s].view(sh).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + s].view(sh).clone()}
        return {"state": state, "param_groups": [dict(self.param_groups[0])]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.param_groups[0].update({k: g[k] for k in ("lr", "betas", "eps") if k in g})
        self.param_groups[0]["betas"] = tuple(self.param_groups[0]["betas"])
        self.step_count = 0
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for i, st in sd.get("state", {}).items():
            i = int(i)
            o, s = PARAM_OFFSETS[i], PARAM_SIZES[i]
            self.exp_avg[o:o + s].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + s].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = int(float(st["step"]))


class _TrainScratch:
    def __init__(self):
        self.ws = None
        self.loss = None

    def get(self, n, device):
        need = _lib.lib().cdq_train_workspace_bytes(n)
        if self.ws is None or self.ws.numel() < need or self.ws.device != device:
            self.ws = torch.empty(need, dtype=torch.uint8, device=device)
        if self.loss is None or self.loss.device != device:
            self.loss = torch.zeros(1, dtype=torch.float32, device=device)
        return self.ws, self.loss


_scratch = _TrainScratch()


def _weights_struct(tensors):
    return _lib.CdqWeights(*[t.data_ptr() for t in tensors])


def dqn_train_step(q_network, target_network, optimizer, states, next_states, rewards, dones, gamma=0.99,
                   loss_kind=LOSS_MSE, group=None):
    """One replay step on device tensors: states/next_states packed boards int64 [n,9], rewards and
    dones fp32 [n].  Returns the loss (fp32 CUDA tensor, mean over the GLOBAL batch when
    torch.distributed is initialised)."""
    n = states.shape[0]
    dev = states.device
    params = dict(q_network.named_parameters())
    weights = _weights_struct([params[k] for k in PARAM_ORDER])
    gv = optimizer.grad_views()
    grads = _weights_struct([gv[k] for k in PARAM_ORDER])
    ws, loss = _scratch.get(n, dev)
    optimizer.zero_grad()
    loss.zero_()
    _lib.check(_lib.lib().cdq_train_fwd_bwd(
        q_network.net_handle(), target_network.net_handle(), ctypes.byref(weights), states.data_ptr(),
        next_states.data_ptr(), rewards.data_ptr(), dones.data_ptr(), n, float(gamma), int(loss_kind),
        ctypes.byref(grads), loss.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    scale = allreduce_mean_scale(optimizer.grad, group)
    optimizer.step(grad_scale=scale)
    out = loss.clone()
    if scale != 1.0:
        import torch.distributed as dist
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        out *= scale
    return out
