"""Raw host-to-device copy rates of one box, per GPU, alone and together (run under torchrun, one rank per GPU):
which GPUs share a host uplink shows as the rate they get when they copy at the same time.
usage: python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py  ->  markdown on rank 0"""
import os, time
import torch, torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
MB = 46.7
host = torch.empty(int(MB * 1e6), dtype=torch.uint8, pin_memory=True).fill_(rank + 1)
dev = torch.empty_like(host, device="cuda")
back = torch.empty(int(4e6), dtype=torch.uint8, pin_memory=True)
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

def rate(active, with_d2h=False, reps=30):
    """GB/s of this rank's H2D copies while the ranks in `active` copy at the same time (0 when this rank idles)"""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if rank in active:
        for _ in range(reps):
            with torch.cuda.stream(s_up):
                dev.copy_(host, non_blocking=True)
            if with_d2h:
                with torch.cuda.stream(s_dn):
                    back.copy_(dev[: back.numel()], non_blocking=True)
        torch.cuda.synchronize()
        r = reps * host.numel() / (time.perf_counter() - t0) / 1e9
    else:
        r = 0.0
    t = torch.tensor([r], device="cuda")
    if world > 1:
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        return [float(x) for x in g]
    return [r]

rate(set(range(world)))
rows = []
for r in range(world):
    rows.append((f"GPU {r} alone", rate({r})))
for r in range(0, world - 1, 2):
    rows.append((f"GPUs {r} and {r + 1}", rate({r, r + 1})))
if world >= 4:
    rows.append(("GPUs 0-3", rate({0, 1, 2, 3})))
    rows.append(("even GPUs", rate(set(range(0, world, 2)))))
rows.append((f"all {world}", rate(set(range(world)))))
rows.append((f"all {world}, with a 4 MB D2H per copy", rate(set(range(world)), True)))
if rank == 0:
    print(f"| copying at the same time (pinned {MB} MB blocks, 30 copies) | " + " | ".join(f"GPU {r}" for r in range(world)) + " | sum GB/s |")
    print("|---|" + "---:|" * (world + 1))
    for name, v in rows:
        print(f"| {name} | " + " | ".join("%.1f" % x if x else "" for x in v) + " | %.1f |" % sum(v))
if world > 1:
    dist.destroy_process_group()
