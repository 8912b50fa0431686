"""Row-strip decomposition of the grid across GPUs (one process per GPU).

`plan_strips` splits the global rows; `connect_strips` wires the per-rank sims together:
each rank exports the CUDA-IPC handle of its state slab, the handles are all-gathered
with torch.distributed (NCCL or gloo — plumbing only), and each rank connects to the
strip above and below.  After that the step needs NO collective: the CUDA library pushes
edge rows and step flags straight into the neighbours' memory over NVLink.

`exchange_halo_rows` is the host-side statement of the same exchange on plain arrays
(send my outermost rows, receive the neighbour's) over torch.distributed point-to-point;
the world_size-2 gloo tests use it with the oracle as the per-strip stepper to pin the
halo bookkeeping (depth 2k rows per k fused steps) without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

HALO_ROWS = 8            # TWS_HALO_ROWS: rows pushed per exchange (2 x max temporal block)
ROW_ALIGN = 8            # strips are cut on multiples of this many rows


@dataclass(frozen=True)
class StripPlan:
    height: int
    world_size: int
    bounds: Tuple[Tuple[int, int], ...]     # [row_begin, row_end) per rank

    def rows(self, rank: int) -> Tuple[int, int]:
        return self.bounds[rank]

    def up(self, rank: int) -> Optional[int]:
        return rank - 1 if rank > 0 else None

    def down(self, rank: int) -> Optional[int]:
        return rank + 1 if rank + 1 < self.world_size else None


def plan_strips(height: int, world_size: int) -> StripPlan:
    """Contiguous row strips, sizes as equal as ROW_ALIGN allows, every strip >= HALO_ROWS rows."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    if world_size == 1:
        return StripPlan(height, 1, ((0, height),))
    if height < world_size * HALO_ROWS:
        raise ValueError(f"{height} rows cannot be split into {world_size} strips of >= {HALO_ROWS} rows")
    units = height // ROW_ALIGN                      # whole alignment units; the remainder goes to the last strip
    base, extra = divmod(units, world_size)
    bounds: List[Tuple[int, int]] = []
    r = 0
    for i in range(world_size):
        n = (base + (1 if i < extra else 0)) * ROW_ALIGN
        e = height if i == world_size - 1 else r + n
        bounds.append((r, e))
        r = e
    return StripPlan(height, world_size, tuple(bounds))


def connect_strips(sim, plan: StripPlan, rank: int, group=None) -> None:
    """All-gather the halo handles and connect `sim` (rank's strip) to its neighbours."""
    if plan.world_size == 1:
        return
    import torch.distributed as dist
    handles: List[Optional[bytes]] = [None] * plan.world_size
    dist.all_gather_object(handles, sim.halo_export(), group=group)
    up, down = plan.up(rank), plan.down(rank)
    sim.halo_connect(handles[up] if up is not None else None, handles[down] if down is not None else None)
    dist.barrier(group=group)


def exchange_halo_rows(own_top, own_bottom, plan: StripPlan, rank: int, group=None):
    """Send my top rows up / bottom rows down, return (halo_from_up, halo_from_down) tensors
    (None at the global edge).  Tensors are torch tensors of identical shape on all ranks."""
    import torch
    import torch.distributed as dist
    up, down = plan.up(rank), plan.down(rank)
    ops, from_up, from_down = [], None, None
    if up is not None:
        from_up = torch.empty_like(own_top)
        ops += [dist.P2POp(dist.isend, own_top.contiguous(), up, group=group), dist.P2POp(dist.irecv, from_up, up, group=group)]
    if down is not None:
        from_down = torch.empty_like(own_bottom)
        ops += [dist.P2POp(dist.isend, own_bottom.contiguous(), down, group=group), dist.P2POp(dist.irecv, from_down, down, group=group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return from_up, from_down
