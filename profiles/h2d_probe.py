"""What the host side of a node sustains towards N GPUs at once: every rank copies a pinned 30.4 MB buffer (the decision variables of
1024 quadruped trajectories, one bench step's H2D) to its GPU, first one rank at a time, then all ranks together.  The ratio is the
end-to-end scaling ceiling of any step that uploads its inputs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/h2d_probe.py"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
BYTES, REPS = 30416896, 50
host = torch.empty(BYTES, dtype=torch.uint8).pin_memory()
dev = torch.empty(BYTES, dtype=torch.uint8, device="cuda")


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        dev.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return BYTES * REPS / (e0.elapsed_time(e1) * 1e-3) / 1e9


timed()
solo = torch.zeros(world, device="cuda")
for r in range(world):  # one rank at a time
    barrier()
    if r == rank:
        solo[r] = timed()
barrier()
together = torch.zeros(world, device="cuda")
together[rank] = timed()  # all ranks at once
if world > 1:
    dist.all_reduce(solo)
    dist.all_reduce(together)
if rank == 0:
    try:
        import subprocess

        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    except Exception:
        topo = ""
    print(f"H2D of {BYTES / 1e6:.1f} MB pinned buffers, {world} GPU(s), GB/s per GPU")
    print("  one at a time :", " ".join(f"{v:5.1f}" for v in solo.tolist()))
    print("  all together  :", " ".join(f"{v:5.1f}" for v in together.tolist()), f"  sum {together.sum().item():.0f} GB/s"
          f"  ({together.sum().item() / max(solo.sum().item(), 1e-9):.2f} of the sum of the solo rates)")
    print(topo)
if world > 1:
    dist.destroy_process_group()
