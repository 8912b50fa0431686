"""Raw host<->device copy ceiling per rank with all ranks copying at once (what bounds bench.py's e2e leg): pinned 256 MiB
buffers, H2D alone, D2H alone, both directions at once on two streams.  Launch with torchrun (one rank per GPU) or plain python."""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
hu, hd = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)
du, dd = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(up, down, reps=10):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                du.copy_(hu, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                hd.copy_(dd, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return (n * reps / dt / 1e9 if up else 0.0), (2 * n * reps / dt / 1e9 if down else 0.0)


for _ in range(2):
    run(True, True, 2)
res = {"ranks": world, "h2d_only_gbs_per_rank": run(True, False)[0], "d2h_only_gbs_per_rank": run(False, True)[1]}
u, d = run(True, True)
res.update({"both_h2d_gbs_per_rank": u, "both_d2h_gbs_per_rank": d,
            "note": "both: 1 byte up per 2 bytes down, the ratio of tws_step_host (4 B/cell up, 8 B/cell down); slowest rank"})
if rank == 0:
    print("PCIE " + json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
