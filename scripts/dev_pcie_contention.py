"""developer probe (torchrun): PCIe copies of the e2e path (8 MB up, 24 MB down per rank)
when every rank copies at once vs. rank 0 alone."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
up_h = torch.empty(8 * 1000 * 1000, dtype=torch.uint8, pin_memory=True)
up_d = torch.empty_like(up_h, device=dev)
dn_d = torch.empty(24 * 1000 * 1000, dtype=torch.uint8, device=dev)
dn_h = torch.empty(24 * 1000 * 1000, dtype=torch.uint8, pin_memory=True)


def timed(fn, active, n=20):
    out = []
    for _ in range(n):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        if active:
            fn(); torch.cuda.synchronize()
        out.append(time.perf_counter() - t0)
    out.sort()
    return 1e3 * out[len(out) // 2]


for name, fn in (("H2D 8 MB", lambda: up_d.copy_(up_h, non_blocking=True)),
                 ("D2H 24 MB", lambda: dn_h.copy_(dn_d, non_blocking=True))):
    t_all = timed(fn, True)
    t_solo = timed(fn, rank == 0)
    ts = [None] * dist.get_world_size()
    dist.all_gather_object(ts, (round(t_all, 3), round(t_solo, 3)))
    if rank == 0:
        print(name, "all ranks at once (ms per rank):", [x[0] for x in ts], " rank 0 alone:", ts[0][1], flush=True)
dist.destroy_process_group()
