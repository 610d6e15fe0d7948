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
NPARAMS_PAD = (NPARAMS + 3) // 4 * 4      # flat buffers are padded to whole 16-byte chunks (padding stays 0)

LOSS_MSE, LOSS_HUBER = 0, 1
LOSS_UPSTREAM = 2          # CDQ_LOSS_UPSTREAM: backward of the forward alone, d loss / d q supplied by the caller (autograd)
REPACK = 0x200             # CDQ_TRAIN_REPACK: the call re-packs the online network from the fp32 parameters
ZERO_LOSS = 0x400          # CDQ_TRAIN_ZERO_LOSS: the call zeroes the loss accumulator itself
WS_PERSISTENT = 0x100      # CDQ_TRAIN_WS_PERSISTENT: the workspace is ours, zero-filled once, reused every step


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


class _DevicePtr:
    """Zero-copy torch view of raw device memory (a CUDA IPC region) via __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerRegion:
    """One device region per rank -- [flat parameters | flat gradients | flag rows] -- mapped by every other rank of the
    data-parallel group, for cdq_dp_adam_step / cdq_train_step (the fused reduce-scatter + Adam + all-gather kernel).
    Two ways to get the mapping, both plumbing only:
      * torch.distributed._symmetric_memory (default where it works): peer pointers AND, on NVSwitch machines, an NVLink
        MULTICAST address of the region, with which the kernel reduces in the switch (multimem.ld_reduce) and broadcasts
        with one store (multimem.st);
      * CUDA IPC handles exchanged through the process group (cdq_ipc_*): unicast peer access only.
    CDQ_DP_SYMM=0 forces the second, CDQ_DP_MULTICAST=0 keeps symmetric memory but ignores its multicast address.
    world = 1 needs no mapping at all."""

    FLAG_BYTES = 256

    def __init__(self, group=None):
        import os
        import torch.distributed as dist
        dev = _lib.require_cuda()
        self.rank, self.world = 0, 1
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_RANKS:
            raise _lib.CdqError("cdq_dp_adam_step supports at most %d ranks" % _lib.MAX_RANKS)
        nbytes = 2 * NPARAMS_PAD * 4 + self.FLAG_BYTES
        self._base, self._mapped, self._symm = None, [], None
        self.multicast = False
        bases, mc = None, 0
        if self.world > 1 and os.environ.get("CDQ_DP_SYMM", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm
                raw = symm.empty(nbytes, dtype=torch.uint8, device=dev)
                raw.zero_()
                torch.cuda.synchronize()
                hdl = symm.rendezvous(raw, group if group is not None else dist.group.WORLD)
                bases = [int(p) for p in hdl.buffer_ptrs]
                mc = int(hdl.multicast_ptr) if os.environ.get("CDQ_DP_MULTICAST", "1") != "0" else 0
                self._symm = (raw, hdl)
            except Exception as e:  # noqa: BLE001 -- symmetric memory is an optimisation; CUDA IPC below always works
                self._symm_error = repr(e)
                bases = None
        if bases is None:
            L = _lib.lib()
            base, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            _lib.check(L.cdq_ipc_alloc(nbytes, ctypes.byref(base), handle))
            self._base = base
            raw = torch.as_tensor(_DevicePtr(base.value, nbytes), device=dev)
            bases = [None] * self.world
            bases[self.rank] = base.value
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw), group=group)
                for r, h in enumerate(handles):
                    if r == self.rank:
                        continue
                    ptr = ctypes.c_void_p()
                    _lib.check(L.cdq_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(ptr)))
                    self._mapped.append(ptr)
                    bases[r] = ptr.value
        self.params = raw[:NPARAMS_PAD * 4].view(torch.float32)
        self.grads = raw[NPARAMS_PAD * 4:2 * NPARAMS_PAD * 4].view(torch.float32)
        if self.world > 1:
            dist.barrier(group=group)
        self.peers = _lib.CdqDpPeers()
        for r in range(self.world):
            self.peers.params[r] = bases[r]
            self.peers.grads[r] = bases[r] + NPARAMS_PAD * 4
            self.peers.flags[r] = bases[r] + 2 * NPARAMS_PAD * 4
        if mc:
            self.multicast = True
            self.peers.mc_params = mc
            self.peers.mc_grads = mc + NPARAMS_PAD * 4

    def close(self):
        L = _lib.lib()
        torch.cuda.synchronize()
        for ptr in self._mapped:
            L.cdq_ipc_close(ptr)
        self._mapped = []
        self.params = self.grads = None
        if self._base is not None:
            L.cdq_ipc_free(self._base)
            self._base = None
        self._symm = None


def flatten_parameters(network, flat=None):
    """Re-home the network's eight parameters as views into one contiguous fp32 buffer (so the
    gradient reduction and the Adam update are one call each).  Returns the flat buffer."""
    params = dict(network.named_parameters())
    first = params[PARAM_ORDER[0]]
    if flat is None:
        flat = torch.zeros(NPARAMS_PAD, dtype=torch.float32, device=first.device)
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

    def __init__(self, network, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, peer_group=False):
        """peer_group: False = plain buffers (single GPU, or NCCL all-reduce through step(grad_scale));
        True / a ProcessGroup = parameters and gradients live in a CUDA-IPC PeerRegion so that
        step_fused() runs the one-kernel reduce-scatter + Adam + all-gather over NVLink."""
        dev = next(network.parameters()).device      # state_dict()/load_state_dict() work anywhere; stepping needs the B200
        self.network = network
        self.region = None
        if peer_group is not False:
            self.region = PeerRegion(None if peer_group is True else peer_group)
            self.flat = flatten_parameters(network, self.region.params)
            self.grad = self.region.grads
        else:
            self.flat = flatten_parameters(network)
            self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        self._d_step = torch.zeros(1, dtype=torch.int32, device=dev)       # device copy of step_count (step_fused)
        self._grid_bar = torch.zeros(32, dtype=torch.int32, device=dev)   # per optimizer launch of a step: arrivals, generation, time-out flag, last complete generation
        self._flags = torch.zeros(64, dtype=torch.int32, device=dev)        # flag row of the world = 1 case
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False,
                              "params": list(range(len(PARAM_ORDER)))}]
        self._link_param_grads()

    def _link_param_grads(self):
        """Every parameter's .grad is a view of the flat gradient buffer, so `loss.backward()` through
        ChessQNetwork's autograd surface accumulates where step() reads.  A .grad that was replaced
        (module.zero_grad(set_to_none=True), a fresh tensor assigned by the caller) is copied in and
        re-linked."""
        params = dict(self.network.named_parameters())
        for name, view in self.grad_views().items():
            p = params[name]
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                if p.grad is not None:
                    view.copy_(p.grad)
                p.grad = view

    def zero_grad(self, set_to_none=False):
        self.grad.zero_()
        self.network._cdq_grads_dirty = False

    def grad_views(self):
        return {n: self.grad[o:o + s].view(sh) for n, o, s, sh in zip(PARAM_ORDER, PARAM_OFFSETS, PARAM_SIZES, PARAM_SHAPES)}

    @property
    def world(self):
        return self.region.world if self.region is not None else 1

    def _peers(self):
        if self.region is not None:
            return self.region.peers, self.region.rank, self.region.world
        peers = _lib.CdqDpPeers()
        peers.params[0], peers.grads[0], peers.flags[0] = self.flat.data_ptr(), self.grad.data_ptr(), self._flags.data_ptr()
        return peers, 0, 1

    def optimizer_struct(self):
        """CdqOptimizer for cdq_train_step (the whole replay step as one call, optimizer launches inside the backward pass)."""
        g = self.param_groups[0]
        peers, rank, world = self._peers()
        return _lib.CdqOptimizer(peers, rank, world, self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), NPARAMS_PAD,
                                 self._d_step.data_ptr(), self._grid_bar.data_ptr(), g["lr"], g["betas"][0], g["betas"][1], g["eps"])

    def step_fused(self):
        """cdq_dp_adam_step: gradient reduction over the peer group (if any), Adam with the step
        counter on the device, parameter broadcast and gradient zeroing in one kernel.  Safe to
        capture in a CUDA graph.  Every rank of the group must call it."""
        g = self.param_groups[0]
        peers, rank, world = self._peers()
        self.step_count += 1
        _lib.check(_lib.lib().cdq_dp_adam_step(ctypes.byref(peers), rank, world, self.exp_avg.data_ptr(),
                                               self.exp_avg_sq.data_ptr(), NPARAMS_PAD, self._d_step.data_ptr(),
                                               self._grid_bar.data_ptr(), g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                                               _lib.stream_ptr()))
        self.network.invalidate_packed()
        self.network._cdq_param_epoch = getattr(self.network, "_cdq_param_epoch", 0) + 1

    def step(self, grad_scale=1.0):
        g = self.param_groups[0]
        self._link_param_grads()
        self.network._cdq_grads_dirty = True     # unlike step_fused, this leaves the gradients in place (torch semantics)
        self.step_count += 1
        self._d_step.fill_(self.step_count)
        _lib.check(_lib.lib().cdq_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), NPARAMS, self.step_count, g["lr"],
                                            g["betas"][0], g["betas"][1], g["eps"], grad_scale, _lib.stream_ptr()))
        self.network.invalidate_packed()   # parameters changed behind autograd's back: repack before next forward
        self.network._cdq_param_epoch = getattr(self.network, "_cdq_param_epoch", 0) + 1

    def _gathered_state(self):
        """exp_avg / exp_avg_sq of all parameters.  Under a peer group each rank only maintains its
        own slice (the rest of its arrays is zero), so the full state is the sum over ranks."""
        if self.world == 1:
            return self.exp_avg, self.exp_avg_sq
        import torch.distributed as dist
        m, v = self.exp_avg.clone(), self.exp_avg_sq.clone()
        dist.all_reduce(m)
        dist.all_reduce(v)
        return m, v

    def _owned_mask(self):
        chunks = NPARAMS_PAD // 4
        lo, hi = chunks * self.region.rank // self.region.world * 4, chunks * (self.region.rank + 1) // self.region.world * 4
        mask = torch.zeros(NPARAMS_PAD, dtype=torch.bool, device=self.flat.device)
        mask[lo:hi] = True
        return mask

    def state_dict(self):
        state = {}
        exp_avg, exp_avg_sq = self._gathered_state()
        if self.step_count:
            for i, (o, s, sh) in enumerate(zip(PARAM_OFFSETS, PARAM_SIZES, PARAM_SHAPES)):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": exp_avg[o:o + s].view(sh).clone(),
                            "exp_avg_sq": exp_avg_sq[o:o + s].view(sh).clone()}
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
        if self.world > 1:          # keep only the owned slice of the optimizer state
            mask = self._owned_mask()
            self.exp_avg.mul_(mask)
            self.exp_avg_sq.mul_(mask)
        self._d_step.fill_(self.step_count)


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


def _mean_over(n, global_batch, world):
    """cdq_train_fwd_bwd's normaliser: every rank divides by global_batch / world, so the 1/world average of the
    per-rank sums is the mean over the global batch whatever the shard sizes (0 = this call's own n)."""
    return float(global_batch) / world if global_batch else 0.0


def dqn_train_step(q_network, target_network, optimizer, states, next_states, rewards, dones, gamma=0.99,
                   loss_kind=LOSS_MSE, group=None, global_batch=None):
    """One replay step on device tensors: states/next_states packed boards int64 [n,9], rewards and
    dones fp32 [n].  Returns the loss (fp32 CUDA tensor, mean over the GLOBAL batch when
    torch.distributed is initialised; pass global_batch when the shards are not all the same size)."""
    n = states.shape[0]
    world = 1
    if global_batch:
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
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
        next_states.data_ptr(), rewards.data_ptr(), dones.data_ptr(), n, _mean_over(n, global_batch, world), float(gamma),
        int(loss_kind), ctypes.byref(grads), loss.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    scale = allreduce_mean_scale(optimizer.grad, group)
    optimizer.step(grad_scale=scale)
    out = loss.clone()
    if scale != 1.0:
        import torch.distributed as dist
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        out *= scale
    return out


class GraphedDQNStep:
    """The whole replay step -- weight repack, q = net(s), q' = target(s'), loss, backward, gradient
    reduction over the peer group, Adam, parameter broadcast -- as ONE CUDA graph per rank
    (about 15 kernels; launched eagerly from Python they cost more host time than device time at
    batch 4 096).  Inputs are copied into static device buffers, so shapes are fixed at construction.
    Under a peer group (FusedAdam(peer_group=True)) every rank holds `batch` = its shard; pass
    global_batch when the shards are not all the same size, so that the step is the mean over the
    global batch (neural_network.py:241) and not the mean of the shard means."""

    def __init__(self, q_network, target_network, optimizer, batch, gamma=0.99, loss_kind=LOSS_MSE, global_batch=None,
                 replay=None):
        """replay: a device-resident replay memory (dict of tensors states / next_states / rewards / dones / ring / counter
        and an int seed, see DQNAgent._device_memory): the minibatch is then drawn and gathered by cdq_replay_sample as the
        first node of the graph, so a step needs no input from the host at all."""
        dev = _lib.require_cuda()
        self.replay = replay
        self.q_network, self.target_network, self.optimizer = q_network, target_network, optimizer
        self.batch, self.gamma, self.loss_kind = int(batch), float(gamma), int(loss_kind)
        self.mean_over = _mean_over(self.batch, global_batch, optimizer.world)
        # the four static inputs are views of ONE device buffer, so a host minibatch arrives with one copy
        B = self.batch
        self.inputs = torch.zeros(B * self.INPUT_BYTES_PER_SAMPLE, dtype=torch.uint8, device=dev)
        self.states = self.inputs[:B * 72].view(torch.int64).view(B, 9)
        self.next_states = self.inputs[B * 72:B * 144].view(torch.int64).view(B, 9)
        self.rewards = self.inputs[B * 144:B * 148].view(torch.float32)
        self.dones = self.inputs[B * 148:B * 152].view(torch.float32)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)        # mean over THIS rank's shard
        self.ws = torch.zeros(_lib.lib().cdq_train_workspace_bytes(self.batch), dtype=torch.uint8, device=dev)
        self.graph = None
        self._param_versions = None
        optimizer.zero_grad()
        target_network.net_handle()       # packed once here; update_target_network() repacks in place

    def _body(self):
        params = dict(self.q_network.named_parameters())
        weights = _weights_struct([params[k] for k in PARAM_ORDER])
        gv = self.optimizer.grad_views()
        grads = _weights_struct([gv[k] for k in PARAM_ORDER])
        if self.replay is not None:
            r = self.replay
            _lib.check(_lib.lib().cdq_replay_sample(r["states"].data_ptr(), r["next_states"].data_ptr(), r["rewards"].data_ptr(),
                                                    r["dones"].data_ptr(), r["ring"].data_ptr(), r["seed"], r["counter"].data_ptr(),
                                                    self.batch, self.states.data_ptr(), self.next_states.data_ptr(),
                                                    self.rewards.data_ptr(), self.dones.data_ptr(), _lib.stream_ptr()))
        opt = self.optimizer.optimizer_struct()
        _lib.check(_lib.lib().cdq_train_step(
            self.q_network.raw_handle(), self.target_network.net_handle(), ctypes.byref(weights),
            self.states.data_ptr(), self.next_states.data_ptr(), self.rewards.data_ptr(), self.dones.data_ptr(),
            self.batch, self.mean_over, self.gamma, self.loss_kind | WS_PERSISTENT | ZERO_LOSS, ctypes.byref(grads), self.loss.data_ptr(),
            self.ws.data_ptr(), self.ws.numel(), ctypes.byref(opt), _lib.stream_ptr()))
        self.optimizer.step_count += 1
        self.q_network.invalidate_packed()

    INPUT_BYTES_PER_SAMPLE = 72 + 72 + 4 + 4     # state record, next-state record, reward, done

    def load_packed(self, staged):
        """One host-to-device copy of a whole minibatch: `staged` is a (pinned) uint8 tensor laid out like
        self.inputs -- [B state records | B next-state records | B rewards | B dones]."""
        self.inputs.copy_(staged, non_blocking=True)

    def load(self, states, next_states, rewards, dones):
        self.states.copy_(states, non_blocking=True)
        self.next_states.copy_(next_states, non_blocking=True)
        self.rewards.copy_(rewards, non_blocking=True)
        self.dones.copy_(dones, non_blocking=True)

    def run(self):
        """One step on the current contents of the static buffers; returns the local loss tensor."""
        if getattr(self.q_network, "_cdq_grads_dirty", False):
            self.optimizer.zero_grad()         # an eager backward / step() in between left gradients behind; the step accumulates
        # cdq_train_step expects the online network's tiles to be packed from the current parameters and leaves them packed
        # from the updated ones; parameters changed from outside since the last step (load_state_dict bumps the tensors'
        # version counters, FusedAdam.step / step_fused bump the network's _cdq_param_epoch) are repacked here
        versions = tuple(p._version for p in self.q_network.parameters()) + (getattr(self.q_network, "_cdq_param_epoch", 0),)
        if versions != self._param_versions:
            self.q_network.net_handle()
            self._param_versions = versions
        if self.graph is None:
            self._body()                       # eager once: function attributes, lazy tables
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # Captured on a HIGH-priority stream: the library's side stream (target forward, fc1 weight gradient)
            # has default = lowest priority, so wherever both have CTAs pending the critical chain
            # (fc1 data gradient -> conv2 backward -> conv1 gradient) gets the SMs first and the weight
            # gradient fills what it leaves free.  Kernel nodes inherit the priority of the stream they were captured on.
            with torch.cuda.graph(g, stream=torch.cuda.Stream(priority=-1)):
                self._body()
            self.graph = g
            self.optimizer.step_count -= 1     # capture recorded the launch but did not run it
        else:
            self.target_network.net_handle()   # repacks (eagerly, in place) after update_target_network()
            self.graph.replay()
            self.optimizer.step_count += 1
        self.q_network.invalidate_packed()     # eager forwards after this must repack
        return self.loss

    def __call__(self, states, next_states, rewards, dones):
        self.load(states, next_states, rewards, dones)
        return self.run()

    def global_loss(self):
        """Loss over the global batch: the average over the ranks of the peer group (exact for equal shards, and
        for unequal ones when global_batch was given)."""
        out = self.loss.clone()
        if self.optimizer.world > 1:
            import torch.distributed as dist
            dist.all_reduce(out)
            out /= self.optimizer.world
        return out


def dp_parity_report(make_net, states, next_states, rewards, dones, gamma=0.99, real_steps=2):
    """Data-parallel parity of the replay step, measured at gradient level ("N-GPU ==
    1-GPU to fp32 reduction-order noise").  Call on EVERY rank of an initialised NCCL group with the
    same full batch; make_net() -> (q_network, target_network) builds identical replicas.  Checks

      a. the gradient the fused kernel reduces (read back as Adam's first moment after ONE step with
         lr = 0: exp_avg = (1 - beta1) * g) against the rank-ordered sum of the shard gradients computed
         eagerly on private replicas and gathered with NCCL             -> max_rel_vs_shard_sum (<= 1e-5)
      b. that sum against the gradient of the whole batch on ONE GPU    -> max_rel_vs_full_batch (<= 1e-3)
      c. the loss over `real_steps` real steps (lr = 1e-3) against the single-GPU step (<= 1e-4 relative)
      d. that all replicas hold identical parameter bits afterwards.

    Relative errors are max |a - b| / max |b| per parameter tensor, maximised over the tensors.
    Shards come from shard_range (uneven when the batch does not divide).  Returns a dict (same on all ranks)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    n = states.shape[0]
    lo, hi = shard_range(n, rank, world)
    sl = slice(lo, hi)
    L = _lib.lib()

    def eager_gradient(net, tgt, s, ns, r, d, mean_over):
        opt = FusedAdam(net, lr=0.0)
        opt.zero_grad()
        params = dict(net.named_parameters())
        weights = _weights_struct([params[k] for k in PARAM_ORDER])
        gv = opt.grad_views()
        grads = _weights_struct([gv[k] for k in PARAM_ORDER])
        m = s.shape[0]
        ws = torch.zeros(L.cdq_train_workspace_bytes(m), dtype=torch.uint8, device=s.device)
        loss = torch.zeros(1, dtype=torch.float32, device=s.device)
        _lib.check(L.cdq_train_fwd_bwd(net.net_handle(), tgt.net_handle(), ctypes.byref(weights), s.data_ptr(), ns.data_ptr(),
                                       r.data_ptr(), d.data_ptr(), m, mean_over, float(gamma), LOSS_MSE, ctypes.byref(grads),
                                       loss.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
        return opt.grad.clone(), loss

    def max_rel(a, b):
        worst = 0.0
        for o, sz in zip(PARAM_OFFSETS, PARAM_SIZES):
            ref = b[o:o + sz]
            worst = max(worst, float((a[o:o + sz] - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
        return worst

    s, ns, r, d = states[sl].contiguous(), next_states[sl].contiguous(), rewards[sl].contiguous(), dones[sl].contiguous()
    # --- expected: shard gradients (mean_over = n / world) gathered, summed in rank order, scaled like the kernel
    g_shard, _ = eager_gradient(*make_net(), s, ns, r, d, n / world)
    gathered = [torch.empty_like(g_shard) for _ in range(world)]
    dist.all_gather(gathered, g_shard)
    g_sum = torch.zeros_like(g_shard)
    for g in gathered:
        g_sum += g
    g_sum *= 1.0 / world
    # --- single GPU, whole batch
    g_full, _ = eager_gradient(*make_net(), states, next_states, rewards, dones, 0.0)
    # --- a: ONE fused data-parallel step with lr = 0; the reduced gradient is Adam's first moment / (1 - beta1)
    qa, ta = make_net()
    opt_a = FusedAdam(qa, lr=0.0, peer_group=True)
    step_a = GraphedDQNStep(qa, ta, opt_a, batch=hi - lo, gamma=gamma, global_batch=n)
    step_a.load(s, ns, r, d)
    step_a.run()
    torch.cuda.synchronize()
    m_dp, _ = opt_a._gathered_state()
    expect_m = g_sum * (1.0 - opt_a.param_groups[0]["betas"][0])
    rel_shard = max_rel(m_dp, expect_m)
    rel_full = max_rel(g_sum, g_full)
    status_a = L.cdq_dp_adam_status(opt_a._grid_bar.data_ptr(), _lib.stream_ptr())
    dist.barrier()
    opt_a.region.close()
    # --- c, d: real steps against the single-GPU step
    qb, tb = make_net()
    opt_b = FusedAdam(qb, lr=1e-3, peer_group=True)
    step_b = GraphedDQNStep(qb, tb, opt_b, batch=hi - lo, gamma=gamma, global_batch=n)
    step_b.load(s, ns, r, d)
    q1, t1 = make_net()
    opt_1 = FusedAdam(q1, lr=1e-3)
    step_1 = GraphedDQNStep(q1, t1, opt_1, batch=n, gamma=gamma)
    step_1.load(states, next_states, rewards, dones)
    loss_rel = 0.0
    for _ in range(real_steps):
        step_b.run()
        l_dp = float(step_b.global_loss().item())
        l_1 = float(step_1.run().item())
        loss_rel = max(loss_rel, abs(l_dp - l_1) / max(abs(l_1), 1e-30))
    torch.cuda.synchronize()
    flat = opt_b.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([1.0 if torch.equal(flat, ref) else 0.0], device=flat.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    stats = torch.tensor([rel_shard, rel_full, loss_rel, float(status_a != 0)], dtype=torch.float64, device=flat.device)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    dist.barrier()
    opt_b.region.close()
    rel_shard, rel_full, loss_rel, timed_out = [float(x) for x in stats.tolist()]
    identical = bool(same.item() == 1.0)
    return {"ok": bool(rel_shard <= 1e-5 and rel_full <= 1e-3 and loss_rel <= 1e-4 and identical and not timed_out),
            "world": world, "global_batch": n, "shards": [shard_range(n, k, world)[1] - shard_range(n, k, world)[0] for k in range(world)],
            "max_rel_vs_shard_sum": rel_shard, "max_rel_vs_full_batch": rel_full, "loss_rel": loss_rel,
            "replicas_identical": identical, "bars": {"vs_shard_sum": 1e-5, "vs_full_batch": 1e-3, "loss_rel": 1e-4}}
