#!/usr/bin/env python
"""C5 of BASELINE.json: offline render of an arcball orbit, row-block sharded over the ranks, one grouped NCCL
exchange of the RGBA / endpoint / hit-id row blocks to rank 0 per frame, overlay on rank 0.

    python tools/orbit.py --size 8192 --frames 64                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \\
        tools/orbit.py --size 8192 --frames 64 [--out orbit_out --save-every 16]

Orbit: default camera, frame k = k successive horizontal drags of 360/frames degrees
(linon_b200.main_view.orbit_camera, src/arcball.rs:73-88).  Enhance step size (h = 0.001, 5001 RK4 steps per ray),
MirageSphericalSigmoid, Lyapunov overlay on.  Per frame the device time of every rank is taken with CUDA events on
the library's stream; a frame costs the MAX over ranks plus the gather and the overlay on rank 0.  Rank 0 prints one
JSON line (and writes it to gpurun_out/orbit_<size>_<frames>f_<N>gpu.json when that directory exists).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _DevBuf:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3, strides=None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--frames", type=int, default=64, help="frames rendered")
    ap.add_argument("--orbit-frames", type=int, default=64, help="frames of the full 360 degree orbit")
    ap.add_argument("--frame-stride", type=int, default=1, help="render every n-th frame of the orbit")
    ap.add_argument("--rows-per-block", type=int, default=32)
    ap.add_argument("--out", default="")
    ap.add_argument("--save-every", type=int, default=0)
    ap.add_argument("--fn", type=int, default=3)
    ap.add_argument("--quiet-zone", type=int, default=0, choices=[0, 1], help="LINON_OPT_QUIET_ZONE (opt-in)")
    ap.add_argument("--peer-stores", type=int, default=0, choices=[0, 1],
                    help="1: the ranks store their rows straight into rank 0's buffers (CUDA IPC + NVLink), no gather (opt-in)")
    args = ap.parse_args()

    import numpy as np
    import torch

    from linon_b200 import _capi as capi
    from linon_b200 import frame_io, sharding
    from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings, orbit_camera

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W = args.size
    H = args.height or args.size
    fn = PredefinedFunction(args.fn)
    verts, faces, center = default_scene()
    mv = MainView(verts, faces, center, W, H, discrete_gpu=True, device=local_rank)
    mv.reload_shader(fn.to_code())
    settings = make_settings(overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)
    mv.update_settings(settings)
    mv.set_shard(rank, world, args.rows_per_block)
    if args.quiet_zone:
        mv.set_option(capi.OPT_QUIET_ZONE, 1)
    if args.peer_stores and world > 1:
        sharding.setup_peer_stores(mv, rank, world)
    stream = torch.cuda.ExternalStream(mv.stream(), device=torch.device("cuda", local_rank))
    ep = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_ENDPOINT_F32X4), (H, W, 4), "<f4"), device="cuda")
    rc = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_RAYCAST_RGBA8), (H, W, 4), "|u1"), device="cuda")
    hid = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_HIT_ID_I32), (H, W), "<i4"), device="cuda")

    march_flags = capi.RENDER_ENHANCE | capi.RENDER_NO_SAMPLE_RAYS
    if world > 1:
        march_flags |= capi.RENDER_NO_OVERLAY

    def barrier():
        mv.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def render(k):
        cam = orbit_camera(center, W, H, k, n_frames=args.orbit_frames).uniform()
        mv.set_camera_uniform(cam)
        mv.render_flags(march_flags)
        gathered = 0
        if world > 1:
            if args.peer_stores:
                mv.sync()
                dist.barrier()  # every rank's rows are in rank 0's buffers
            else:
                with torch.cuda.stream(stream):
                    gathered = sharding.gather_rows_to_root([rc, ep, hid], H, rank, world, args.rows_per_block)
            if rank == 0:
                mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)
        return cam, gathered

    render(0)  # JIT + first touch, untimed
    barrier()

    frame_ids = list(range(0, args.orbit_frames, args.frame_stride))[: args.frames]
    per_frame, march_ms_all, steps_total, gathered_bytes = [], [], 0, 0
    t_wall0 = time.perf_counter()
    for k in frame_ids:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        cam, moved = render(k)
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        c = mv.counters() if world == 1 else None
        t = mv.timings()
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_frame.append(float(ms.item()))
        gathered_bytes += moved
        if world == 1:
            steps_total += c.accepted_steps + c.rejected_steps
            march_ms_all.append(t.march_ms)
        if rank == 0 and args.out and args.save_every and (k // args.frame_stride) % args.save_every == 0:
            frame_io.write_frame(os.path.join(args.out, f"orbit_{k:04d}"), mv.texture(), cam, settings, fn.to_code(),
                                 capi.RENDER_ENHANCE, counters=c, timings=t,
                                 extra=dict(orbit_frame=k, orbit_frames=args.orbit_frames, n_gpus=world))
    wall = time.perf_counter() - t_wall0
    if world > 1:  # every ray runs exactly 5001 steps in the Enhance variant
        steps_total = len(frame_ids) * W * H * 5001
    if rank == 0:
        total_ms = float(sum(per_frame))
        rays = len(frame_ids) * W * H
        line = dict(config="C5", workload=f"{W}x{H} offline orbit, {len(frame_ids)} of {args.orbit_frames} frames (stride {args.frame_stride}), "
                                          f"{fn.name}, Enhance h=0.001, overlay delta=1, {args.rows_per_block}-row blocks round-robin over {world} GPU(s)"
                                          + (" + NCCL gather to rank 0" if world > 1 else ""),
                    n_gpus=world, frames=len(frame_ids), ms_per_frame=total_ms / len(frame_ids), total_s=total_ms / 1e3,
                    mrays_s=rays / total_ms / 1e3, rk4_gsteps_s=steps_total / total_ms / 1e6,
                    ms_per_frame_min=min(per_frame), ms_per_frame_max=max(per_frame),
                    gather_bytes_per_frame=gathered_bytes / len(frame_ids), wall_s=wall)
        if march_ms_all:
            line["march_ms_per_frame"] = float(np.mean(march_ms_all))
        print(json.dumps(line), flush=True)
        if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
            with open(os.path.join(ROOT, "gpurun_out", f"orbit_{W}x{H}_{len(frame_ids)}f_{world}gpu.json"), "w") as f:
                json.dump(dict(line, per_frame_ms=per_frame), f, indent=1)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
