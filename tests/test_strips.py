"""Host-side logic of the multi-GPU row-strip decomposition, on CPU:
plan_strips properties, and a world_size-2/3 gloo run in which each rank steps its strip
(+ halo rows) with the ORACLE as the per-strip stepper and exchanges 2k halo rows per k
steps through terrainwatersim_b200.strips.exchange_halo_rows.  The stitched result must be
bit-identical to the whole-grid oracle — this pins the halo depth and ordering rules the
CUDA exchange implements."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.oracle_py import Oracle, dam_break, new_state
from terrainwatersim_b200.strips import HALO_ROWS, exchange_halo_rows, plan_strips


def test_plan_strips_properties():
    for height, n in [(32768, 8), (65536, 8), (8192, 4), (1000, 3), (64, 8), (100, 1), (8200, 7)]:
        plan = plan_strips(height, n)
        assert plan.bounds[0][0] == 0 and plan.bounds[-1][1] == height
        for i, (a, b) in enumerate(plan.bounds):
            assert b - a >= HALO_ROWS
            if i:
                assert a == plan.bounds[i - 1][1]
            if i < n - 1:
                assert b % 8 == 0
        sizes = [b - a for a, b in plan.bounds]
        assert max(sizes) - min(sizes) <= 8 + height % 8
        assert plan.up(0) is None and plan.down(n - 1) is None
    with pytest.raises(ValueError):
        plan_strips(40, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, k, blocks, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle()
        h, d = dam_break(W, H, rim=False)
        h = (h + np.sin(np.arange(H, dtype=np.float32))[:, None] * 3).astype(np.float32)      # break y-uniformity
        consts = o.derive_consts(float(W), W)
        plan = plan_strips(H, world)
        r0, r1 = plan.rows(rank)
        P = 2 * k                                           # halo depth needed for k fused steps
        lo = r0 - (P if plan.up(rank) is not None else 0)
        hi = r1 + (P if plan.down(rank) is not None else 0)
        t, f, v = new_state(h[lo:hi], d[lo:hi])
        top = r0 - lo                                       # index of own row 0 in the local arrays
        for _ in range(blocks):
            # local arrays end at the cut (exterior reads 0 there): cells closer than 2k rows to a cut
            # are wrong after k steps, own rows are not.
            o.step(t, f, v, consts, k)
            own_t, own_f = t[top:top + (r1 - r0)], f[top:top + (r1 - r0)]
            send_top = torch.from_numpy(np.concatenate([own_t[:P, :, 3:4], own_f[:P]], axis=2).copy())
            send_bot = torch.from_numpy(np.concatenate([own_t[-P:, :, 3:4], own_f[-P:]], axis=2).copy())
            from_up, from_down = exchange_halo_rows(send_top, send_bot, plan, rank)
            if from_up is not None:
                t[:P, :, 3] = from_up[..., 0].numpy(); f[:P] = from_up[..., 1:].numpy()
            if from_down is not None:
                t[-P:, :, 3] = from_down[..., 0].numpy(); f[-P:] = from_down[..., 1:].numpy()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), d=t[top:top + (r1 - r0), :, 3], f=f[top:top + (r1 - r0)],
                 v=v[top:top + (r1 - r0)].view(np.uint16))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(2, 1), (2, 4), (3, 2)])
def test_strip_exchange_with_gloo_matches_whole_grid(tmp_path, built, world, k):
    W, H, blocks = 64, 72, 6
    mp.spawn(_worker, args=(world, _free_port(), W, H, k, blocks, str(tmp_path)), nprocs=world, join=True)
    o = Oracle()
    h, d = dam_break(W, H, rim=False)
    h = (h + np.sin(np.arange(H, dtype=np.float32))[:, None] * 3).astype(np.float32)
    t, f, v = new_state(h, d)
    o.step(t, f, v, o.derive_consts(float(W), W), k * blocks)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([p["d"] for p in parts]).view(np.uint32), t[..., 3].view(np.uint32))
    assert np.array_equal(np.concatenate([p["f"] for p in parts]).view(np.uint32), f.view(np.uint32))
    assert np.array_equal(np.concatenate([p["v"] for p in parts]), v.view(np.uint16))
