"""Multi-GPU plumbing: MPC instances are independent, so the batch is cut into contiguous shards (one process per
GPU, no data-path collective) and the only exchange per outer iteration is one all-gather of the 32-scalar
per-trajectory summaries.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU
tests)."""
from __future__ import annotations

SUMMARY_SIZE = 32
SUMMARY_FIELDS = {"cost": 24, "barrier": 25, "eq_inf_norm": 26, "ineq_max": 27}


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous shard [start, stop) of rank `rank`: sizes differ by at most one, shards tile [0, total)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_summaries(local, group=None):
    """All-gather of per-trajectory summaries ``local[B_local, 32]`` -> ``[sum(B_local), 32]`` in rank order.
    Equal shard sizes take the single-collective path (all_gather_into_tensor); ragged shards are padded to the widest shard."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0], local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    width = max(sizes)  # ragged shards: pad to the widest, gather once, drop the padding
    padded = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * width:r * width + n] for r, n in enumerate(sizes)], dim=0)


class SummaryExchange:
    """The per-iteration all-gather, overlapped with the next iteration's sweep.

    ``slot()`` hands out one of two local summary buffers; ``post(slot)`` starts the all-gather of that buffer with ``async_op=True``
    (NCCL runs it on its own stream, ordered after the work already queued on the current stream), so the sweep of iteration i + 1
    runs while the summaries of iteration i cross NVLink.  A buffer is reused two iterations later, after ``slot()`` has made the
    current stream wait for the collective that read it.  ``latest()`` waits for the newest posted gather and returns
    ``[world * B_local, 32]`` in rank order.  Equal shard sizes only (the bench's weak-scaling layout)."""

    def __init__(self, batch_local: int, dtype, device, group=None):
        import torch
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.local = [torch.zeros((batch_local, SUMMARY_SIZE), dtype=dtype, device=device) for _ in range(2)]
        self.gathered = [torch.zeros((self.world * batch_local, SUMMARY_SIZE), dtype=dtype, device=device) for _ in range(2)]
        self.pending = [None, None]
        self.turn = 0
        self.newest = None

    def slot(self) -> int:
        k = self.turn
        self.turn ^= 1
        if self.pending[k] is not None:  # the gather posted two iterations ago still owns this buffer
            self.pending[k].wait()
            self.pending[k] = None
        return k

    def post(self, k: int) -> None:
        import torch.distributed as dist

        if self.world == 1:
            self.gathered[k] = self.local[k]
        else:
            self.pending[k] = dist.all_gather_into_tensor(self.gathered[k], self.local[k], group=self.group, async_op=True)
        self.newest = k

    def latest(self):
        if self.newest is None:
            raise RuntimeError("no summaries were posted yet")
        k = self.newest
        if self.pending[k] is not None:
            self.pending[k].wait()
            self.pending[k] = None
        return self.gathered[k]

    def drain(self) -> None:
        for k in (0, 1):
            if self.pending[k] is not None:
                self.pending[k].wait()
                self.pending[k] = None


def fleet_status(summaries) -> dict:
    """What an outer loop looks at after the gather: total objective, worst constraint violations."""
    return {"trajectories": int(summaries.shape[0]), "total_cost": float(summaries[:, SUMMARY_FIELDS["cost"]].sum()),
            "total_barrier": float(summaries[:, SUMMARY_FIELDS["barrier"]].sum()),
            "worst_eq_inf_norm": float(summaries[:, SUMMARY_FIELDS["eq_inf_norm"]].max()),
            "worst_ineq": float(summaries[:, SUMMARY_FIELDS["ineq_max"]].max())}
