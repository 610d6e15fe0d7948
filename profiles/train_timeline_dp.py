"""Data-parallel DQN replay step (global batch 4 096) under torchrun: in-graph timeline of one replay on rank 0
(torch.profiler / CUPTI), the fused reduce-scatter + Adam + all-gather kernel's duration on every rank, and the NVLink
bytes the step moves per GPU (NVML throughput counters over 400 replays) against the 2 (W-1)/W x 4.28 MB it should.
    python -m torch.distributed.run --nproc-per-node W profiles/train_timeline_dp.py"""
import json
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import chess_deep_q_b200 as cdq  # noqa: E402
from chess_deep_q_b200.train import NPARAMS_PAD, FusedAdam, GraphedDQNStep, shard_range  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(os.environ.get("BATCH", "4096"))
boards = torch.empty((B + 1, 9), dtype=torch.int64, device="cuda")
L = cdq._lib.lib()
cdq._lib.check(L.cdq_random_playouts(boards.data_ptr(), B + 1, 42, 199, cdq._lib.stream_ptr()))
torch.manual_seed(0)
net, tgt = cdq.ChessQNetwork().cuda(), cdq.ChessQNetwork().cuda()
tgt.load_state_dict(net.state_dict())
opt = FusedAdam(net, peer_group=(world > 1))
lo, hi = shard_range(B, rank, world)
s, ns = boards[lo:hi].contiguous(), boards[lo + 1:hi + 1].contiguous()
sc, fl = cdq.evaluate_batch(ns)
step = GraphedDQNStep(net, tgt, opt, batch=hi - lo, global_batch=B)
step.load(s, ns, (0.01 * torch.tanh(sc / 10)).float(), ((fl & 7) != 0).float())
for _ in range(20):
    step.run()
torch.cuda.synchronize()


def nvlink_kib():
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        vals = nv.nvmlDeviceGetFieldValues(h, [nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX])
        return [int(v.value.ullVal) for v in vals]
    except Exception as e:  # noqa: BLE001
        return repr(e)


if world > 1:
    dist.barrier()
N = 400
before = nvlink_kib()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    step.run()
e1.record()
torch.cuda.synchronize()
after = nvlink_kib()
ms = e0.elapsed_time(e1) / N
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        step.run()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name and "emset" not in e.name]
ev.sort(key=lambda e: e.time_range.start)
# a step starts at the first conv kernel that follows a weight-pack kernel (the previous step's tail)
starts = [i for i, e in enumerate(ev) if "conv_stream" in e.name and i > 0 and "pack_weights" in ev[i - 1].name]
lo_i, hi_i = starts[-2], starts[-1]
t0 = ev[lo_i].time_range.start
lines = ["%8.1f us  +%6.1f us  %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name.split("(")[0][:60]) for e in ev[lo_i:hi_i]]
span = max(e.time_range.end for e in ev[lo_i:hi_i]) - t0
adam_us = [e.time_range.end - e.time_range.start for e in ev if "adam" in e.name][2:]
mine = {"rank": rank, "ms_per_step": ms, "span_us": span, "optimizer_kernels_us": [round(x, 1) for x in adam_us[-2:]],
        "nvlink_tx_bytes_per_step": (after[0] - before[0]) * 1024 / N if isinstance(after, list) and isinstance(before, list) else after,
        "nvlink_rx_bytes_per_step": (after[1] - before[1]) * 1024 / N if isinstance(after, list) and isinstance(before, list) else after}
allr = [mine]
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, mine)
if rank == 0:
    print("# world %d, global batch %d (%d per rank): %.4f ms per step (CUDA events over %d replays, rank 0), %.2f M samples/s"
          % (world, B, hi - lo, max(r["ms_per_step"] for r in allr), N, B / max(r["ms_per_step"] for r in allr) / 1e3))
    print("# expected NVLink traffic per GPU and step: reduce-scatter reads (W-1)/W x %.2f MB + all-gather / gradient-zeroing writes 2 (W-1)/W x %.2f MB"
          % (NPARAMS_PAD * 4 / 1e6, NPARAMS_PAD * 4 / 1e6))
    for r in allr:
        print("# " + json.dumps(r))
    print("# in-graph timeline of one replay on rank 0:")
    print("\n".join(lines))
    print("step span %.1f us" % span)
if world > 1:
    dist.barrier()
    opt.region.close()
    dist.destroy_process_group()
