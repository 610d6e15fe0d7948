"""Does this box support NVLink multicast through torch's symmetric memory (multimem.ld_reduce / multimem.st)?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = symm.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "multicast_ptr", hex(h.multicast_ptr), "buffers", [hex(p) for p in h.buffer_ptrs], "signal pads", [hex(p) for p in h.signal_pad_ptrs][:2],
      "local ptr", hex(t.data_ptr()), flush=True)
dist.barrier()
dist.destroy_process_group()
