"""Row-block sharding of a frame across ranks and the single gather to rank 0.

The march is embarrassingly parallel per pixel (src/main_view.wgsl:340-375): rows are cut into blocks of
`rows_per_block` and dealt round-robin (block b -> rank b % world) because cost per row varies (rays that miss
the box run the all-face intersection on every step).  Every rank renders its rows in place inside full-frame
buffers; afterwards ONE grouped exchange moves every non-root block to the same offset on rank 0
(torch.distributed batch_isend_irecv: NCCL send/recv over NVLink on GPUs, gloo on CPU for the host-logic tests).
The overlay needs a halo of `delta` rows of endpoints, so it runs once on rank 0 over the assembled image.
No collective is needed inside the march.
"""
from __future__ import annotations

from typing import List, Tuple


def row_blocks(height: int, rows_per_block: int) -> List[Tuple[int, int]]:
    """[(row_start, row_end)) of every block, in order."""
    return [(y, min(y + rows_per_block, height)) for y in range(0, height, rows_per_block)]


def owner(block: int, world: int) -> int:
    return block % world


def local_row_count(height: int, rank: int, world: int, rows_per_block: int) -> int:
    return sum(e - s for b, (s, e) in enumerate(row_blocks(height, rows_per_block)) if owner(b, world) == rank)


def local_rows(height: int, rank: int, world: int, rows_per_block: int) -> List[int]:
    out: List[int] = []
    for b, (s, e) in enumerate(row_blocks(height, rows_per_block)):
        if owner(b, world) == rank:
            out.extend(range(s, e))
    return out


def gather_rows_to_root(buffers, height: int, rank: int, world: int, rows_per_block: int, group=None, root: int = 0):
    """buffers: list of full-frame torch tensors whose first dimension is the image row ([H, ...], contiguous).

    Every non-root rank sends the row blocks it owns; root receives each block into the same rows of its own
    buffers.  One grouped call (a single NCCL group on GPUs).  Returns the number of payload bytes that crossed
    the interconnect as seen by this rank."""
    import torch.distributed as dist

    if world == 1:
        return 0
    ops = []
    moved = 0
    for b, (s, e) in enumerate(row_blocks(height, rows_per_block)):
        src = owner(b, world)
        if src == root:
            continue
        for t in buffers:
            blk = t[s:e]
            if rank == root:
                ops.append(dist.P2POp(dist.irecv, blk, src, group=group))
                moved += blk.numel() * blk.element_size()
            elif rank == src:
                ops.append(dist.P2POp(dist.isend, blk, root, group=group))
                moved += blk.numel() * blk.element_size()
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return moved


def setup_peer_stores(mv, rank: int, world: int, root: int = 0):
    """Multi-GPU frame without a gather (include/linon_b200.h: linon_ipc_export / linon_ipc_open / linon_redirect_output).

    The root rank exports CUDA IPC handles of its colour / endpoint / hit-id frame buffers, the handles travel through one
    torch.distributed broadcast, and every other rank redirects its march stores into the root's buffers: each shard writes
    its own row blocks over NVLink while it marches, so after `mv.sync()` on every rank and a barrier the root holds the
    assembled frame and runs the overlay.  One process per GPU, `torch.distributed` initialised with the nccl backend."""
    import torch
    import torch.distributed as dist

    from . import _capi as capi

    if world == 1:
        return
    bufs = (capi.BUF_RAYCAST_RGBA8, capi.BUF_ENDPOINT_F32X4, capi.BUF_HIT_ID_I32)
    # the handles ride on the job's backend: a device tensor under nccl, a host tensor under gloo (CPU tests)
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    handles = torch.zeros(64 * len(bufs), dtype=torch.uint8, device=device)
    if rank == root:
        raw = b"".join(mv.ipc_export(b) for b in bufs)
        handles.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(handles, root)
    if rank != root:
        raw = handles.cpu().numpy().tobytes()
        for k, b in enumerate(bufs):
            mv.redirect_output(b, mv.ipc_open(raw[64 * k:64 * k + 64]))
