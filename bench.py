"""bench.py -- headline benchmark: batched leaf evaluation (handcrafted score + CNN value).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--positions P] [--impl reference]

Metric (BASELINE.json): leaf positions/sec.  One "step" = one pass of the hot path
(cdq_leaf_eval: K1 bitboard eval -> K2a conv stack -> K2b fc head) over one batch of synthetic
random-playout positions.  N=1 workload = BASELINE configs[1]: 1 048 576 positions on one B200.
Under torchrun every rank evaluates its own 1M positions (weak scaling, no data-path collective).

  value       whole-job positions/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e         same metric through LeafEvaluator.evaluate_host: pinned HOST buffers in, host out
  roofline    dominant kernel: algorithmic FLOP per launch / mean launch time (CUDA events on the
              launching stream, recorded in a second pass of the same K steps) vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port (C eval on all cores + torch-CPU fp32 CNN) on a bounded sample
  cpu_baseline_python  the reference's Python path (evaluation.py's work in CPython + batch-1 torch-CPU CNN)
  gpu_baseline  the reference's own module in torch-eager on the same B200 (fp32 and bf16 autocast), forward at
              B = 1 ... 1M and DQNAgent.train()'s maths at B = 4 096, next to this repo's path at the same batch
  dqn         second metric: the replay step as one CUDA graph (device-resident), .e2e = DQNAgent.train() from the
              replay memory, .reference = torch-CPU, .parity = gradient-level data-parallel parity at world > 1
--impl reference times that CPU port as the reference arm (the reference itself is Python over
python-chess, which is not installable here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "leaf positions/sec (handcrafted eval + CNN value)"
UNIT = "positions/s"
FLOP_CONV = 2 * (221184 + 1179648)      # conv1 + conv2, dense convention
FLOP_FC = 2 * (1048576 + 256)
H2D_PER_POS, D2H_PER_POS = 72, 16       # record in; leaf f32 + score f64 + flags u32 out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": p.get("bf16_tflops_sustained", 1368.3), "hbm": p.get("hbm_gbs", 6490.5), "src": "measured"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled through NVML every 10 ms while the timed
    region runs (nvidia-smi as a subprocess needs about a second to produce its first line)."""
    BITS = {"gpu_idle": 0x1, "applications_clocks_setting": 0x2, "sw_power_cap": 0x4, "hw_slowdown": 0x8,
            "sync_boost": 0x10, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80,
            "display_clock_setting": 0x100}

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.err = index, [], False, None, None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons(h)),
                                  nv.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.01)
        except Exception as e:  # noqa: BLE001 -- reported in the JSON line, never fatal for the benchmark
            self.err = repr(e)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2.0)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling failed: %s" % self.err], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        mask = 0
        for r in self.rows:
            mask |= r[1]
        watts = sorted(r[2] for r in self.rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(k for k, bit in self.BITS.items() if mask & bit and k != "gpu_idle"),
                "reasons_mask": mask, "power_w_median": watts[len(watts) // 2], "power_w_max": watts[-1], "samples": len(sm)}


def cpu_port_throughput(n_sample, steps=1, passes=1):
    """The reference path's CPU port: oracle C evaluation (all host threads) + fp32 torch-CPU CNN.
    One step = `passes` passes over n_sample unique positions (the port keeps no cache, so every
    pass is the full work)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1
    boards = O.playouts(n_sample, seed=42)
    params = O.init_params(42, 2.0)
    planes = None
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        for _ in range(passes):
            O.evaluate(boards, nthreads=0)
            planes = O.encode_planes(boards)
            with torch.no_grad():
                for i in range(0, n_sample, 4096):
                    O.net_forward(params, planes[i:i + 4096])
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return n_sample * passes / mean, mean, os.cpu_count(), torch.get_num_threads()


def reference_network_class():
    """The reference's ChessQNetwork re-declared from its four nn layers (neural_network.py:16-37): the same
    module torch would build from the reference file, which cannot travel to the GPU box."""
    import torch.nn as nn
    import torch.nn.functional as F

    class RefChessQNetwork(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(12, 32, kernel_size=3, padding=1)
            self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
            self.fc1 = nn.Linear(64 * 8 * 8, 256)
            self.fc2 = nn.Linear(256, 1)

        def forward(self, x):
            x = F.relu(self.conv1(x))
            x = F.relu(self.conv2(x))
            x = x.view(-1, 64 * 8 * 8)
            x = F.relu(self.fc1(x))
            return self.fc2(x).tanh()

    return RefChessQNetwork


def _event_time_ms(fn, iters, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gpu_baseline(cdq, L, ev, net, boards, rew, dn):
    """The reference's only GPU implementation -- its own module in torch-eager (cuDNN / cuBLAS) -- on THIS B200
    (BASELINE.md 3.3-3.4): forward at B = 1 / 8 / 64 / 512 / 4 096 / 1 M in fp32 (torch defaults:
    TF32 convolutions) and under bf16 autocast, inputs already on the device; the per-leaf call exactly as
    mcts.py:216-219 makes it (host tensor -> .to(device) -> forward -> .item()); and DQNAgent.train()'s maths
    (neural_network.py:232-246: two forwards, MSE, backward, torch.optim.Adam) at B = 4 096.  Next to each,
    this repo's path at the same batch (cdq_leaf_eval = evaluation kernel + CNN; GraphedDQNStep)."""
    import torch
    import torch.nn.functional as F
    Ref = reference_network_class()
    ref = Ref().cuda()
    ref.load_state_dict(net.state_dict())
    tgt = Ref().cuda()
    tgt.load_state_dict(net.state_dict())
    rows = []
    sizes = (1, 8, 64, 512, 4096, min(1 << 20, boards.shape[0]))
    planes_all = cdq.boards_to_tensor_batch(boards[:sizes[-1]])
    for B in sizes:
        x = planes_all[:B]
        iters = 200 if B <= 512 else 50 if B <= 4096 else 3
        row = {"batch": B}
        with torch.no_grad():
            ms = _event_time_ms(lambda: ref(x), iters)
            row["eager_fp32_us"] = ms * 1e3
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ms16 = _event_time_ms(lambda: ref(x), iters)
            row["eager_bf16_autocast_us"] = ms16 * 1e3
        b = boards[:B].contiguous()
        ours = _event_time_ms(lambda: ev.evaluate_device(b), iters)
        row["cdq_leaf_eval_us"] = ours * 1e3
        row["eager_fp32_positions_per_s"] = B / (ms * 1e-3)
        row["eager_bf16_positions_per_s"] = B / (ms16 * 1e-3)
        row["cdq_positions_per_s"] = B / (ours * 1e-3)
        row["speedup_vs_eager_fp32"] = ms / ours
        row["speedup_vs_eager_bf16"] = ms16 / ours
        rows.append(row)
    del planes_all
    # the per-leaf call of the reference's search (mcts.py:216-219), wall clock, one leaf at a time
    host_x = cdq.boards_to_tensor_batch(boards[:1]).cpu()
    with torch.no_grad():
        for _ in range(20):
            ref(host_x.to("cuda")).item()
        t0 = time.perf_counter()
        for _ in range(300):
            ref(host_x.to("cuda")).item()
        per_leaf_us = (time.perf_counter() - t0) / 300 * 1e6
    # DQNAgent.train()'s maths in torch-eager at B = 4 096
    B = 4096
    s_pl = cdq.boards_to_tensor_batch(boards[:B])
    ns_pl = cdq.boards_to_tensor_batch(boards[1:B + 1])
    r, d = rew[:B].contiguous(), dn[:B].contiguous()
    train = {}
    for mode in ("fp32", "bf16_autocast"):
        ref.load_state_dict(net.state_dict())
        opt = torch.optim.Adam(ref.parameters(), lr=1e-3)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
                q = ref(s_pl).squeeze()
                qn = tgt(ns_pl).squeeze().detach()
                loss = F.mse_loss(q.float(), r + (1 - d) * 0.99 * qn.float())
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        ms = _event_time_ms(step, 30, warm=5)
        train[mode] = {"ms_per_step": ms, "samples_per_s": B / (ms * 1e-3)}
    return {"what": "reference ChessQNetwork / DQNAgent.train maths in torch-eager (cuDNN/cuBLAS) on the same B200, "
                    "CUDA-event timed, inputs resident on the device",
            "torch": torch.__version__, "forward": rows,
            "per_leaf_call_as_mcts_py_216_219_us": per_leaf_us,
            "per_leaf_call_positions_per_s": 1e6 / per_leaf_us,
            "dqn_train_b4096": train}


def python_path_baseline(n_single=1024, per_core=1024):
    """the reference's Python path (evaluation.py's work in CPython over the restated
    python-chess API + ChessQNetwork on torch-CPU at batch 1, mcts.py:216-219) on this box's host cores:
    one process, then one process per core.  oracle/py_eval.py is the restatement (pinned bit for bit to the
    reference's own scores in tests/test_oracle.py)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import py_eval
    cores = os.cpu_count() or 1
    params = O.init_params(42, 2.0)
    boards = O.playouts(max(n_single, per_core * cores), seed=42)
    v1, s1, _ = py_eval.python_path_throughput(boards[:n_single], params, processes=1)
    vN, sN, _ = py_eval.python_path_throughput(boards[:per_core * cores], params, processes=cores)
    e1, es1, _ = py_eval.python_path_throughput(boards[:n_single], None, processes=1)
    return {"value": vN, "unit": UNIT, "cores": cores, "kind": "port",
            "what": "Python path: evaluation.py's work in CPython over the restated python-chess API + torch-CPU fp32 CNN at batch 1",
            "sample": "%d positions on %d processes, %.1f s" % (per_core * cores, cores, sN),
            "single_process": {"value": v1, "sample": "%d positions, %.1f s" % (n_single, s1)},
            "single_process_eval_only": {"value": e1, "sample": "%d positions, %.1f s" % (n_single, es1)}}


def dqn_cpu_reference(steps=2, B=4096):
    """DQNAgent.train()'s maths (neural_network.py:232-246) on torch-CPU fp32 with the reference module
    re-declared from its layers: samples/s on this box's host cores."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    Ref = reference_network_class()
    torch.manual_seed(42)
    q, t = Ref(), Ref()
    t.load_state_dict(q.state_dict())
    opt = torch.optim.Adam(q.parameters(), lr=1e-3)
    s = (torch.rand(B, 12, 8, 8) < 0.04).float()
    ns = (torch.rand(B, 12, 8, 8) < 0.04).float()
    r, d = torch.randn(B) * 0.01, (torch.rand(B) < 0.05).float()
    times = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        cur = q(s).squeeze()
        nxt = t(ns).squeeze().detach()
        loss = F.mse_loss(cur, r + (1 - d) * 0.99 * nxt)
        opt.zero_grad()
        loss.backward()
        opt.step()
        loss.item()
        if i:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": B / sec, "unit": "samples/s", "ms_per_step": sec * 1e3, "cores": os.cpu_count(),
            "torch_threads": torch.get_num_threads(), "kind": "reference maths on torch-CPU fp32, batch %d, %d steps" % (B, steps)}


def eval_kernel_alu(positions_per_s):
    """K1's honest roof: the integer (alu) pipe.  The pipe utilisation is a hardware counter, so it comes from the committed
    ncu capture of one 1M-position launch (profiles/r2_eval_kernel.json); this run's own K1 rate is put next to the rate of
    that capture so the fraction can be scaled."""
    path = os.path.join(ROOT, "profiles", "r2_eval_kernel.json")
    if not os.path.exists(path):
        return None
    j = json.load(open(path))
    captured_rate = j["captured_positions"] / (j["time_us"] * 1e-6)
    return {"bound": "alu pipe (LOP3/SHF/IADD3 at one warp instruction per 2 cycles per SMSP)",
            "frac_in_capture": j["alu_pipe_pct_of_peak"] / 100.0, "issue_slots_in_capture": j["issue_slots_pct_of_peak"] / 100.0,
            "capture_positions_per_s": captured_rate, "this_run_positions_per_s": positions_per_s,
            "frac_scaled_to_this_run": (j["alu_pipe_pct_of_peak"] / 100.0) * positions_per_s / captured_rate if positions_per_s else None,
            "source": "profiles/r2_eval_kernel.json (ncu --set full, sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active)"}


def workload_config(n):
    """The workload both arms are quoted on (BASELINE configs[1] at the default size)."""
    return {"workload": "BASELINE configs[1]: %d random-playout positions per GPU per step, "
                        "handcrafted eval + 12-channel CNN value" % n,
            "positions_per_gpu": n, "l2": "per-step working set (act2 %.1f GB) >> 126 MB L2" % (n * 8192 / 1e9),
            "weights": "numpy seed 42, torch-default init ranges, weight matrices x2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = args.ref_sample
    for _ in range(args.warmup):
        cpu_port_throughput(min(n_sample, 8192))
    v, sec, cores, tthreads = cpu_port_throughput(n_sample, steps=args.steps)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 score / fp32 CNN", "data": "synthetic", "impl": "reference",
            "config": dict(workload_config(args.positions or 1 << 20),
                           sample="each step is a bounded sample of that workload: %d positions" % n_sample),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d positions/step: C oracle eval on %d threads + torch-CPU fp32 CNN (%d threads)"
                                       % (n_sample, cores, tthreads)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "cpu_baseline_python": python_path_baseline(), "dqn": {"reference": dqn_cpu_reference()}}
    print(json.dumps(line))


def run_sweep(args, cdq, L, ev, rank, world, local, dist):
    """Per-GPU batch sweep (BASELINE configs[4]): every rank evaluates its own batch, sizes 4K..16M,
    device-resident, CUDA events, max over ranks; reports positions/s and the tensor-roofline
    fraction of the whole step (4 899 328 FLOP/position) for every size."""
    import torch
    peaks = measured_peaks()
    rows = []
    import ctypes
    for n in (4096, 16384, 65536, 262144, 1048576, 4194304, 8388608, 16777216):     # 8 GPUs x 8M = the 64M-position point
        boards = torch.empty((n, 9), dtype=torch.int64, device="cuda")
        cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), n, 42 + rank, 199, cdq._lib.stream_ptr()))
        steps = max(3, min(200, (8 << 20) // n))
        for _ in range(3):
            ev.evaluate_device(boards)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.cdq_profile_enable(1)                 # event pair around every launch: per-kernel time of this rank
        e0.record()
        for _ in range(steps):
            ev.evaluate_device(boards)
        e1.record()
        torch.cuda.synchronize()
        L.cdq_profile_enable(0)
        kms, kcnt = (ctypes.c_double * 8)(), (ctypes.c_uint64 * 8)()
        cdq._lib.check(L.cdq_profile_collect(kms, kcnt))
        ms = e0.elapsed_time(e1)
        if dist:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        pps = world * n * steps / (ms * 1e-3)
        per_s = lambda k: n * steps / (kms[k] * 1e-3) if kms[k] > 0 else None       # noqa: E731  positions/s inside kernel kind k
        rows.append({"positions_per_gpu": n, "steps": steps, "ms_per_step": ms / steps, "positions_per_s": pps,
                     "tensor_frac_whole_step": pps / world * (FLOP_CONV + FLOP_FC) / 1e12 / peaks["tflops"],
                     "kernel_ms_per_step": {"eval_terms": kms[0] / steps, "conv_stream": kms[1] / steps, "fc_head": kms[2] / steps},
                     "roofline_frac": {"conv_stream_tensor": per_s(1) * FLOP_CONV / 1e12 / peaks["tflops"] if per_s(1) else None,
                                       "fc_head_tensor": per_s(2) * FLOP_FC / 1e12 / peaks["tflops"] if per_s(2) else None,
                                       "fc_head_hbm": per_s(2) * 8196 / 1e9 / peaks["hbm"] if per_s(2) else None,
                                       "eval_terms_hbm": per_s(0) * 84 / 1e9 / peaks["hbm"] if per_s(0) else None}})
        del boards
    if rank == 0:
        print(json.dumps({"metric": METRIC, "unit": UNIT, "n_gpus": world, "scaling": "weak", "sweep": rows,
                          "peak_tflops": peaks["tflops"], "data": "synthetic"}))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--positions", type=int, default=1 << 20, help="positions per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=131072)
    ap.add_argument("--cpu-sample", type=int, default=131072)
    ap.add_argument("--cpu-passes", type=int, default=10, help="passes over the CPU sample (about 10-15 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dqn", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE configs[4]: device-resident throughput for 4K..16M positions per GPU (one JSON line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import chess_deep_q_b200 as cdq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = cdq._lib.lib()
    n = args.positions
    W = max(args.warmup, 3)

    # ---- synthetic inputs, generated on the device by the product's own move generator
    boards = torch.empty((n, 9), dtype=torch.int64, device="cuda")
    cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), n, 42 + rank, 199, cdq._lib.stream_ptr()))
    # weights: same distribution as torch's default init, numpy RNG (portable), x2 "trained-like" range
    rng = np.random.default_rng(42)
    net = cdq.ChessQNetwork()
    with torch.no_grad():
        for name, p in net.named_parameters():
            fan_in = {"conv1": 108, "conv2": 288, "fc1": 4096, "fc2": 256}[name.split(".")[0]]
            a = rng.uniform(-1, 1, size=tuple(p.shape)).astype(np.float32) / np.sqrt(fan_in)
            p.copy_(torch.from_numpy(a * (2.0 if name.endswith("weight") else 1.0)))
    net = net.cuda()
    ev = cdq.LeafEvaluator(net)

    def step():
        return ev.evaluate_device(boards)

    if args.sweep:
        return run_sweep(args, cdq, L, ev, rank, world, local, dist)

    for _ in range(W):
        out = step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()

    # ---- timed region: K steps, device-timed, NOTHING else recorded inside it
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.cdq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.cdq_launch_count() - launches0
    # ---- the same K steps once more with an event pair around every launch: per-kernel durations for the
    #      roofline (kept out of the timed region: two event records per launch cost about 1 % there)
    L.cdq_profile_enable(1)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(args.steps):
        out = step()
    p1.record()
    torch.cuda.synchronize()
    L.cdq_profile_enable(0)
    ms_profiled = p0.elapsed_time(p1)
    kms = (cdq._lib.ctypes.c_double * 8)()
    kcnt = (cdq._lib.ctypes.c_uint64 * 8)()
    cdq._lib.check(L.cdq_profile_collect(kms, kcnt))
    clocks = sampler.stop()
    if dist:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- e2e: pinned host buffers through the public host API, copies inside the timed region
    h_boards = torch.empty((n, 9), dtype=torch.int64).pin_memory()
    h_boards.copy_(boards)
    h_out = {"leaf": torch.empty(n, dtype=torch.float32).pin_memory(),
             "score": torch.empty(n, dtype=torch.float64).pin_memory(),
             "flags": torch.empty(n, dtype=torch.int32).pin_memory()}
    for _ in range(2):
        ev.evaluate_host(h_boards, h_out)
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 20))
    for _ in range(e2e_steps):
        ev.evaluate_host(h_boards, h_out)       # synchronises before returning
    e2e_s = time.perf_counter() - t0
    if dist:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = world * n * e2e_steps / e2e_s
    assert torch.equal(h_out["leaf"].cuda(), out["leaf"]), "host path and device path disagree"

    # ---- second metric (BASELINE configs[3]): DQN replay step, global batch 4096 split over the ranks
    #      (strong scaling), one NCCL all-reduce of the flat gradient per step
    dqn = None
    if not args.no_dqn:
        from chess_deep_q_b200.train import FusedAdam, GraphedDQNStep, shard_range
        B = 4096
        lo, hi = shard_range(B, rank, world)
        tgt = cdq.ChessQNetwork().cuda()
        tgt.load_state_dict(net.state_dict())
        # parameters + gradients in CUDA-IPC peer memory: the optimizer step is ONE kernel that
        # reduce-scatters the gradients over NVLink, runs Adam on the rank's slice and all-gathers
        # the parameters; the whole step is one CUDA graph per rank (no NCCL call in the loop)
        opt = FusedAdam(net, lr=1e-3, peer_group=(world > 1))
        s_b, ns_b = boards[lo:hi].contiguous(), boards[lo + 1:hi + 1].contiguous()
        sc, fl = cdq.evaluate_batch(ns_b)
        rew = (0.01 * torch.tanh(sc / 10.0)).float()                  # chess_ai.py:183
        dn = ((fl & 7) != 0).float()
        gstep = GraphedDQNStep(net, tgt, opt, batch=hi - lo, gamma=0.99)
        gstep.load(s_b, ns_b, rew, dn)
        for _ in range(3):
            gstep.run()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        l0 = L.cdq_launch_count()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        dsteps = 50
        for _ in range(dsteps):
            gstep.run()
        t1e.record()
        torch.cuda.synchronize()
        dms = t0e.elapsed_time(t1e)
        if dist:
            t = torch.tensor([dms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dms = float(t.item())
        dloss = gstep.global_loss()
        dqn = {"metric": "DQN samples/sec", "value": B * dsteps / (dms * 1e-3), "unit": "samples/s", "global_batch": B,
               "ms_per_step": dms / dsteps, "scaling": "strong", "loss": float(dloss.item()),
               "tensor_bound_frac": (19154944.0 * B / world) / (dms / dsteps * 1e-3) / 1e12 / measured_peaks()["tflops"],
               "parallelism": ("dp%d: one CUDA graph per rank; gradients reduce-scattered, Adam and parameter all-gather "
                               "in one kernel over NVLink peer memory%s" % (world, " (NVLS: multimem.ld_reduce / multimem.st through the switch's "
                                                                             "multicast mapping)" if opt.region is not None and opt.region.multicast
                                                                      else " (unicast peer loads / stores)")) if world > 1 else
                              "single GPU, whole step as one CUDA graph"}
        if opt.region is not None:
            if dist:
                dist.barrier()
            opt.region.close()
        if world > 1:
            # data-parallel parity AT THIS WORLD SIZE, gradient level (lr = 0), even and uneven global batch
            from chess_deep_q_b200.train import dp_parity_report
            parity = []
            for nb in (4096, 4099):
                fb = torch.empty((nb + 1, 9), dtype=torch.int64, device="cuda")
                cdq._lib.check(L.cdq_random_playouts(fb.data_ptr(), nb + 1, 777, 199, cdq._lib.stream_ptr()))   # same boards on every rank
                fsc, ffl = cdq.evaluate_batch(fb[1:])
                frew, fdn = (0.01 * torch.tanh(fsc / 10.0)).float(), ((ffl & 7) != 0).float()
                base_sd = {k: v.detach().clone() for k, v in tgt.state_dict().items()}

                def make_net():
                    a, b = cdq.ChessQNetwork().cuda(), cdq.ChessQNetwork().cuda()
                    a.load_state_dict(base_sd)
                    b.load_state_dict(base_sd)
                    return a, b
                parity.append(dp_parity_report(make_net, fb[:nb].contiguous(), fb[1:].contiguous(), frew, fdn))
            dqn["parity"] = {"ok": all(r["ok"] for r in parity), "max_rel": max(r["max_rel_vs_shard_sum"] for r in parity),
                             "reports": parity}
        elif n >= 10001:
            # the path a user calls: DQNAgent.train() -- random.sample from the replay memory, gather into pinned
            # staging, H2D, the graphed step, the loss back as a Python float (neural_network.py:218-252)
            import random
            random.seed(42)
            agent = cdq.DQNAgent(batch_size=B)
            agent.q_network.load_state_dict(tgt.state_dict())
            agent.update_target_network()
            hb = boards[:10001].cpu().numpy().view(np.uint64).reshape(-1).view(cdq.BOARD_DTYPE)
            hsc, hfl = cdq.evaluate_batch(boards[1:10001])
            hrew = (0.01 * torch.tanh(hsc / 10.0)).float().cpu().numpy()
            hdn = ((hfl & 7) != 0).cpu().numpy()
            for i in range(10000):
                agent.memory.append((hb[i], None, float(hrew[i]), hb[i + 1], float(hdn[i])))
            for _ in range(3):
                agent.train()
            t0 = time.perf_counter()
            esteps = 50
            for k in range(esteps):
                # the reference stores one transition per ply and trains once per stored transition (chess_ai.py:187-192):
                # every timed step appends one from host memory, which train() copies to the device mirror of the replay memory
                agent.memory.append((hb[k], None, float(hrew[k]), hb[k + 1], float(hdn[k])))
                eloss = agent.train()
            esec = time.perf_counter() - t0
            dqn["e2e"] = {"value": B * esteps / esec, "unit": "samples/s", "ms_per_step": esec / esteps * 1e3,
                          "h2d_bytes_per_step": 152 + 12, "d2h_bytes_per_step": 4, "loss": eloss,
                          "what": "DQNAgent.train() after one store_transition per step: the new transition (152 B) and the ring geometry (12 B) from "
                                  "host memory to the device mirror of the replay memory (cdq_replay_append), then ONE graph replay -- minibatch drawn "
                                  "and gathered on the device (cdq_replay_sample), forward, backward, optimizer -- and the loss back as a float; "
                                  "the initial upload of the 10 000-transition memory (1.5 MB) is outside the timed region"}
            del agent

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tensor bound)
    peaks = measured_peaks()
    kinds = {0: ("eval_terms_kernel", None), 1: ("conv_stream_kernel", FLOP_CONV), 2: ("fc_head_pairs_kernel", FLOP_FC),
             7: ("leaf_pipe_kernel", FLOP_CONV + FLOP_FC)}      # 7 = conv + fc as one kernel (act2 stays in L2)
    dom = max((1, 2, 7), key=lambda k: kms[k])
    per_launch_pos = n * args.steps / max(1, kcnt[dom])
    avg_ms = kms[dom] / max(1, kcnt[dom])
    achieved = per_launch_pos * kinds[dom][1] / (avg_ms * 1e-3) / 1e12
    traffic = None        # DRAM bytes per launch of that kernel from the committed ncu capture, scaled to this launch size
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        ent = tj.get(kinds[dom][0])
        if ent:
            traffic = ent["traffic_bytes"] * per_launch_pos / ent.get("captured_positions", tj["captured_positions"])
    roofline = {"kernel": kinds[dom][0], "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"],
                "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "traffic_note": "DRAM read+write bytes per launch: ncu capture of a 151 552-position launch (profiles/r2_traffic.json) scaled to "
                                "this run's average launch of %d positions; algorithmic %.3f GB" % (per_launch_pos, per_launch_pos * 8264 / 1e9),
                "peak_source": peaks["src"] + " bf16_tflops_sustained (kind::f16 shares the bf16 rate)",
                "kernel_ms_per_step": {kinds[k][0]: kms[k] / args.steps for k in kinds if kms[k] > 0},
                "eval_kernel_hbm": {"achieved_GBps": n * args.steps * 84 / (kms[0] * 1e-3) / 1e9 if kms[0] else None,
                                    "peak_GBps": peaks["hbm"], "note": "84 B/position; integer-issue bound"},
                "eval_kernel_alu": eval_kernel_alu(n * args.steps / (kms[0] * 1e-3) if kms[0] else None)}

    # ---- BASELINE configs[2] as a throughput line: Russian Doll MCTS self-play, device-resident search, 1 024 lockstep games
    selfplay = None
    if not args.no_selfplay and world == 1:
        from chess_deep_q_b200.mcts import annealed_search_parameters
        its, sew, widths = annealed_search_parameters(0.0)
        rows = []
        for games, plies in ((1, 16), (64, 16), (1024, 12)):
            cdq.self_play_games(net, games, max_moves=2, iterations=its, samples_per_level=widths, exploration_weight=sew)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            played = cdq.self_play_games(net, games, max_moves=plies, iterations=its, samples_per_level=widths, exploration_weight=sew)
            torch.cuda.synchronize()
            sec = time.perf_counter() - t0
            moves = sum(len(g["moves"]) for g in played)
            rows.append({"games": games, "plies_per_game": plies, "leaf_evals_per_s": moves * its / sec, "plies_per_s": moves / sec})
        selfplay = {"workload": "BASELINE configs[2]: Russian Doll MCTS self-play, widths %s, %d simulations per move in waves of 8, "
                                "GPU leaf evaluation; search trees resident on the device, one host synchronisation per ply" % (widths, its),
                    "unit": "leaf evaluations/s", "lockstep": rows}

    gpu_base = None
    if not args.no_gpu_baseline and world == 1 and n >= 16384:
        bsc, bfl = cdq.evaluate_batch(boards[1:4097])
        gpu_base = gpu_baseline(cdq, L, ev, net, boards, (0.01 * torch.tanh(bsc / 10.0)).float(), ((bfl & 7) != 0).float())
        if dqn:
            gpu_base["dqn_train_b4096"]["cdq_graphed_step_ms"] = dqn["ms_per_step"]
            for mode in ("fp32", "bf16_autocast"):
                gpu_base["dqn_train_b4096"][mode]["speedup_of_cdq"] = gpu_base["dqn_train_b4096"][mode]["ms_per_step"] / dqn["ms_per_step"]

    cpu = cpu_py = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only
        v, sec, cores, tthreads = cpu_port_throughput(args.cpu_sample, passes=args.cpu_passes)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d passes over %d random-playout positions: C oracle eval on %d threads + torch-CPU fp32 CNN "
                         "(%d threads), %.1f s" % (args.cpu_passes, args.cpu_sample, cores, tthreads, sec)}
        cpu_py = python_path_baseline()
        if dqn:
            dqn["reference"] = dqn_cpu_reference()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 bitboards + f64 score; fp16 operands / fp32 accumulate CNN", "data": "synthetic",
            "config": workload_config(n),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * H2D_PER_POS, "d2h_bytes_per_step": n * D2H_PER_POS},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "cpu_baseline_python": cpu_py, "gpu_baseline": gpu_base, "dqn": dqn, "selfplay": selfplay,
            "profiled_pass": {"ms_per_step": ms_profiled / args.steps,
                              "note": "per-kernel event pairs are recorded in a second pass of the same K steps, not in the timed region"}}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
