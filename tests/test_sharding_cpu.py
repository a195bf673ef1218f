"""Row-block sharding: partition arithmetic and the single gather to rank 0, exercised with gloo on CPU
(world_size 2 and 3).  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from linon_b200 import sharding


@pytest.mark.parametrize("H,world,rpb", [(600, 2, 32), (1080, 8, 32), (37, 3, 4), (8192, 8, 64), (5, 4, 16)])
def test_partition_covers_every_row_exactly_once(H, world, rpb):
    seen = np.zeros(H, int)
    for r in range(world):
        rows = sharding.local_rows(H, r, world, rpb)
        assert len(rows) == sharding.local_row_count(H, r, world, rpb)
        assert rows == sorted(rows)
        seen[rows] += 1
    assert (seen == 1).all()
    counts = [sharding.local_row_count(H, r, world, rpb) for r in range(world)]
    assert max(counts) - min(counts) <= rpb  # round-robin keeps the shards within one block of each other


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, rpb, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the "frame": value = f(row, col); every rank fills only its own rows, like the sharded march does
        full_rgba = torch.zeros(H, W, 4, dtype=torch.uint8)
        full_ep = torch.zeros(H, W, 4, dtype=torch.float32)
        full_hit = torch.full((H, W), -7, dtype=torch.int32)
        rows = sharding.local_rows(H, rank, world, rpb)
        for y in rows:
            full_rgba[y] = torch.tensor([y % 251, rank, 3, 255], dtype=torch.uint8)
            full_ep[y, :, 0] = float(y)
            full_ep[y, :, 1] = torch.arange(W, dtype=torch.float32)
            full_hit[y] = y * 3 + 1
        moved = sharding.gather_rows_to_root([full_rgba, full_ep, full_hit], H, rank, world, rpb)
        if rank == 0:
            ys = torch.arange(H)
            assert (full_ep[:, :, 0] == ys[:, None].float()).all()
            assert (full_ep[:, :, 1] == torch.arange(W).float()[None, :]).all()
            assert (full_hit == (ys * 3 + 1)[:, None].int()).all()
            owners = torch.tensor([sharding.owner(b, world) for b in range((H + rpb - 1) // rpb)]).repeat_interleave(rpb)[:H]
            assert (full_rgba[:, :, 1] == owners[:, None].to(torch.uint8)).all()
            expect = sum(24 * W * len(sharding.local_rows(H, r, world, rpb)) for r in range(1, world))
            assert moved == expect
            open(out_path, "w").write("ok")
        else:
            assert moved == 24 * W * len(rows)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W,rpb", [(2, 70, 33, 8), (3, 50, 16, 4)])
def test_gather_rows_to_root_gloo(tmp_path, world, H, W, rpb):
    out = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(world, _free_port(), H, W, rpb, str(out)), nprocs=world, join=True)
    assert out.read_text() == "ok"


class _FakeView:
    """Stands in for MainView: records what setup_peer_stores asks of the C ABI."""

    def __init__(self, rank):
        self.rank = rank
        self.redirected = {}

    def ipc_export(self, buffer):
        return bytes([(buffer * 7 + k) % 256 for k in range(64)])

    def ipc_open(self, handle):
        assert len(handle) == 64
        return 0x1000 + handle[0]

    def redirect_output(self, buffer, ptr):
        self.redirected[buffer] = ptr


def _peer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from linon_b200 import _capi as capi
        mv = _FakeView(rank)
        sharding.setup_peer_stores(mv, rank, world)
        bufs = (capi.BUF_RAYCAST_RGBA8, capi.BUF_ENDPOINT_F32X4, capi.BUF_HIT_ID_I32)
        if rank == 0:
            assert mv.redirected == {}  # the root keeps its own buffers
        else:  # every other rank maps the ROOT's handle of each buffer and redirects exactly the three march outputs
            assert mv.redirected == {b: 0x1000 + (b * 7) % 256 for b in bufs}
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_peer_store_handles_reach_every_rank_gloo(tmp_path):
    world = 3
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == [f"ok{r}" for r in range(world)]
