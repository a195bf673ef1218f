#!/usr/bin/env python
"""bench.py — headline benchmark of the linon ray-march path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU port)

One "step" = one frame of the hot path: [click positions] -> march (+ 8 sample rays on a second stream) -> Lyapunov
overlay.  N = 1 runs BASELINE.json configs[2]: 3840x2160, Enhance step size (h = 0.001, 5001 RK4 steps per ray),
overlay on, the application's default field function (MirageSphericalSigmoid), default camera / mesh / settings.
N > 1 is weak scaling of the same frame: N times the pixels over the same field of view (round(3840 sqrt N) x
round(2160 sqrt N), so the per-ray work mix is unchanged), rows cut into blocks dealt round-robin to the ranks, one
grouped NCCL exchange of the RGBA / endpoint / hit-id row blocks to rank 0, overlay on rank 0.

Prints ONE JSON line (rank 0).  value = Mrays/s of the whole job (rays of all ranks / max-over-ranks device time);
e2e = the same through the public MainView API with host buffers (camera+settings upload, final image read back into
pinned memory) inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FN_ID = 3  # MirageSphericalSigmoid: the function the application starts with (src/main_view.rs:100)
BASE_W, BASE_H = 3840, 2160
ROWS_PER_BLOCK = 32


def workload_size(n_gpus: int):
    s = math.sqrt(n_gpus)
    return int(round(BASE_W * s)), int(round(BASE_H * s))


def workload_name(n_gpus: int, w: int, h: int) -> str:
    base = "C3: Cornell box, MirageSphericalSigmoid, Enhance h=0.001 (5001 RK4 steps/ray), Lyapunov overlay delta=1 scaling=50"
    if n_gpus == 1:
        return f"{base}, {w}x{h}, 1xB200"
    return f"{base}, weak-scaled to {w}x{h} ({n_gpus}x the pixels, same field of view), {ROWS_PER_BLOCK}-row blocks round-robin over {n_gpus} GPUs + gather to rank 0"


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(power) if power else None, reasons=sorted(reasons), samples=len(sm))


def ncu_profile():
    """Executed FP32 FLOPs per RK4 step and issue-slot utilisation of the march kernel from the committed ncu summary."""
    path = os.path.join(ROOT, "profiles", "r1_march_c3.json")
    if not os.path.exists(path):
        return None
    for k in json.load(open(path)):
        if k.get("kernel") == "linon_march_tiles" and k.get("fp32_flops_per_rk4_step"):
            return k
    return None


def overlay_ncu_traffic(W, H):
    """DRAM bytes per overlay launch from the committed ncu capture (profiles/r1_overlay_c3.json, taken at the C3 size);
    None for any other frame size."""
    path = os.path.join(ROOT, "profiles", "r1_overlay_c3.json")
    if (W, H) != (3840, 2160) or not os.path.exists(path):
        return None
    for k in json.load(open(path)):
        if "overlay_kernel" in str(k.get("kernel", "")) and k.get("dram_traffic_bytes"):
            return float(k["dram_traffic_bytes"])
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference():
    """The CPU implementation timed beside the GPU path.  oracle/_ref (the reference's own WGSL, transpiled mechanically
    to C++ by oracle/wgsl_ref and compiled with gcc; built where /root/reference is mounted, the .so travels) when it
    exists -> kind "reference"; otherwise the hand-written restatement oracle/linon_oracle.cpp -> kind "port".  The two
    are bit-identical (tests/test_wgsl_reference_cpu.py)."""
    from oracle import wgsl_reference as wr
    if wr.available():
        wr.build()
        return wr, "reference"
    from oracle import cpu_oracle as co
    co.build()
    return co, "port"


def oracle_sample(w: int, h: int, budget_s: float, threads: int = 0):
    """Times the CPU implementation of the reference's path on a bounded sample of the same frame: every `stride`-th
    row, full width (pixels are independent; cost is per ray), sized by a short calibration run to ~budget_s seconds."""
    from oracle import np_oracle as npo

    impl, kind = cpu_reference()
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, w, h))
    st = dict(npo.default_settings(), overlay_mode=1)
    cores = impl.num_threads() if threads <= 0 else threads
    steps_per_ray = 5001  # h_initial = 0.001: the loop of src/main_view.wgsl:244 runs 5001 times, adaptive sampling cannot trigger (h <= 0.002)

    def run(rows):
        y_step = max(1, h // rows)
        ys = range(y_step // 2, h, y_step)
        t0 = time.perf_counter()
        out = impl.march(verts, faces, cam, st, FN_ID, w, h, h_initial=0.001, adaptive=True, rect=(0, w, y_step // 2, h),
                         y_step=y_step, n_threads=threads)
        dt = time.perf_counter() - t0
        rays = len(ys) * w
        if "accepted" in out:  # the port counts its steps; the transpiled shader has no such output
            assert int(out["accepted"].sum() + out["rejected"].sum()) == rays * steps_per_ray
        return dt, rays, rays * steps_per_ray, len(ys)

    # calibration on a strided part of one row
    xs = max(1, w // (cores * 16))
    t0 = time.perf_counter()
    impl.march(verts, faces, cam, st, FN_ID, w, h, h_initial=0.001, adaptive=True, rect=(0, w, h // 2, h // 2 + 1),
               x_step=xs, n_threads=threads)
    per_ray = (time.perf_counter() - t0) / len(range(0, w, xs))
    rows = max(1, min(h, int(budget_s / (per_ray * w))))
    t, rays, steps, rows = run(rows)
    if t < 0.4 * budget_s and rows < h:
        # the calibration (128 rays, thread start-up included) overestimates the cost per ray: size the sample from the
        # rate just measured and time that one instead
        t, rays, steps, rows = run(max(rows + 1, min(h, int(rows * 0.8 * budget_s / t))))
    what = ("the reference's main_view.wgsl + MirageSphericalSigmoid body transpiled to C++ (oracle/_ref), gcc -O2 -ffp-contract=off"
            if kind == "reference" else "hand-written C++ restatement of the reference WGSL (oracle/linon_oracle.cpp)")
    return dict(seconds=t, rays=rays, steps=steps, rows=rows, cores=cores, kind=kind,
                sample=f"{rows} full rows of the {w}x{h} frame, evenly spaced ({rays} rays, {steps} RK4 steps), {cores} OpenMP threads; {what}")


def run_reference(args):
    """--impl reference: the reference's own implementation of the path on the host cores.  The Rust + wgpu host cannot
    be built or run in this image (no rustc/cargo, no Vulkan ICD / lavapipe); what runs is the reference's
    shader itself: main_view.wgsl with the MirageSphericalSigmoid body of functions.rs spliced in, transpiled
    mechanically to C++ and compiled with gcc (oracle/_ref, kind = "reference"; falls back to the hand-written port when
    that library was never built), all host threads, on a bounded sample of the same frame per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, h = workload_size(args.gpus)
    budget = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        oracle_sample(w, h, budget)
    tot_t = tot_rays = tot_steps = 0
    last = None
    for _ in range(args.steps):
        last = oracle_sample(w, h, budget)
        tot_t += last["seconds"]; tot_rays += last["rays"]; tot_steps += last["steps"]
    mrays = tot_rays / tot_t / 1e6
    line = dict(impl="reference", metric="Mrays/s", value=mrays, unit="Mrays/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=tot_t / args.steps * 1e3 * (w * h) / (tot_rays / args.steps),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                rk4_gsteps_per_s=tot_steps / tot_t / 1e9,
                config=dict(workload=workload_name(args.gpus, w, h), width=w, height=h,
                            note="ms_per_step is the sampled time scaled to the full frame (pixels are independent)"),
                cpu_baseline=dict(value=mrays, unit="Mrays/s", cores=last["cores"], kind=last["kind"], sample=last["sample"]),
                e2e=dict(value=mrays, unit="Mrays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="niklaskorz/linon's host is Rust+wgpu (no rustc/cargo/Vulkan in this image): its WGSL compute shader is run on the host cores through a mechanical WGSL->C++ transpile, not through wgpu on lavapipe")
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
class _DevBuf:
    """Exposes a device pointer of the library to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3,
                                             strides=None)


def run_quiet_zone_probe(args):
    """--probe-quiet-zone (child process of the N = 1 run): the C3 frame with the default path and with
    LINON_OPT_QUIET_ZONE on (include/linon_b200.h) - march time of both from CUDA events on the library's stream, L2 flushed
    between frames, and the two frames compared on the device (hit id, colour, endpoint; pixels the overlay marks as
    high-Lyapunov excluded and counted, like the parity tests).  Prints one JSON object."""
    import numpy as np
    import torch

    from linon_b200 import _capi as capi
    from linon_b200 import flops as fl
    from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings

    dev = args.device
    torch.cuda.set_device(dev)
    W, H = workload_size(1)
    verts, faces, center = default_scene()
    mv = MainView(verts, faces, center, W, H, discrete_gpu=True, device=dev)
    mv.reload_shader(PredefinedFunction(FN_ID).to_code())
    mv.update_settings(make_settings(overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0))
    stream = torch.cuda.ExternalStream(mv.stream(), device=torch.device("cuda", dev))
    ep = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_ENDPOINT_F32X4), (H, W, 4), "<f4"), device="cuda")
    rc = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_RAYCAST_RGBA8), (H, W, 4), "|u1"), device="cuda")
    hid = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_HIT_ID_I32), (H, W), "<i4"), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    frames = 3

    sample_rays_ms = {}  # the sample-ray kernel (second stream) takes the same option

    def measure(quiet):
        mv.set_option(capi.OPT_QUIET_ZONE, quiet)
        for _ in range(2):
            mv.render_flags(capi.RENDER_ENHANCE)
            mv.sync()
        frame_ms, march_ms, sample_ms = [], [], []
        for _ in range(frames):
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            mv.render_flags(capi.RENDER_ENHANCE)
            e1.record(stream)
            mv.sync()
            torch.cuda.synchronize()
            frame_ms.append(e0.elapsed_time(e1))
            march_ms.append(mv.timings().march_ms)
            sample_ms.append(mv.timings().sample_rays_ms)
        c = mv.counters()
        sample_rays_ms[quiet] = float(np.mean(sample_ms))
        return float(np.mean(frame_ms)), float(np.mean(march_ms)), c, int(mv.debug_profile()[15])

    def agreement(c):
        same_hit = hid == ref["hid"]
        use = ~excl
        matched = use & same_hit
        n_use, n_matched = int(use.sum()), max(int(matched.sum()), 1)
        err = torch.linalg.vector_norm(ep[..., :3].double() - ref["ep"][..., :3].double(), dim=-1)
        rel = err / torch.clamp(torch.linalg.vector_norm(ref["ep"][..., :3].double(), dim=-1), min=1e-12)
        cdiff = (rc.to(torch.int16) - ref["rc"].to(torch.int16)).abs().amax(dim=-1)
        agree = dict(pixels=W * H, excluded_high_lyapunov=int(excl.sum()), compared=n_use,
                     hit_id_equal=float((same_hit & use).sum()) / max(n_use, 1),
                     colour_within_1=float(((cdiff <= 1) & matched).sum()) / n_matched,
                     endpoint_within_1e3=float(((rel <= 1e-3) & matched).sum()) / n_matched,
                     endpoint_max_rel=float(rel[matched].max()) if int(matched.sum()) else 0.0,
                     hit_id_equal_all_pixels=float(same_hit.sum()) / (W * H))
        agree["meets_north_star"] = bool(agree["hit_id_equal"] >= 0.999 and agree["colour_within_1"] >= 0.999 and
                                         agree["endpoint_within_1e3"] >= 0.999 and int(c.accepted_steps) == W * H * 5001)
        return agree

    d_ms, d_march, d_c, _ = measure(0)
    ref = dict(ep=ep.clone(), rc=rc.clone(), hid=hid.clone())
    excl = torch.as_tensor(mv.exponents().reshape(H, W), device="cuda") >= 0.8
    ms, m_ms, c, quiet = measure(1)
    flops = fl.fused_march_flops(c, FN_ID, field_weight_is_one=True, quiet_steps=quiet)
    # the main result goes out first: whatever the extra measurements below do, the parent already has this line
    print(json.dumps(dict(
        ms_per_step=ms, march_ms=m_ms, value=W * H / (ms * 1e-3) / 1e6, unit="Mrays/s",
        rk4_gsteps_per_s=c.accepted_steps / (ms * 1e-3) / 1e9, frames=frames,
        default_path_same_process=dict(ms_per_step=d_ms, march_ms=d_march, culling_checks=int(d_c.culling_checks)),
        counters=dict(accepted=int(c.accepted_steps), quiet_steps=quiet, quiet_fraction=quiet / max(int(c.accepted_steps), 1),
                      prehit=int(c.prehit_steps), culling_checks=int(c.culling_checks), face_tests=int(c.face_tests)),
        executed_tflops=flops / (m_ms * 1e-3) / 1e12,
        sample_rays_ms=dict(default=sample_rays_ms.get(0), quiet_zone=sample_rays_ms.get(1)),
        vs_default_path=agreement(c),
        note="LINON_OPT_QUIET_ZONE = 1 (opt-in, off in the headline numbers of this line): fixed-step frames of a predefined mirage "
             "body take straight steps while the whole warp is outside the body's zone of influence, where the reference's own "
             "fp32 evaluation has eta == 1; every Runge-Kutta step is executed and counted")), flush=True)

    # (0) the other opt-in kernel of the path: the direct overlay (LINON_OPT_OVERLAY_KERNEL = 1, no shared-memory stage) against
    # the measured default on this very frame - time of the overlay launch alone, L2 flushed before each, results compared
    try:
        fin = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_FINAL_RGBA8), (H, W, 4), "|u1"), device="cuda")
        exb = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_EXPONENTS_F32), (H, W), "<f4"), device="cuda")

        def overlay_only(variant, reps=10):
            mv.set_option(capi.OPT_OVERLAY_KERNEL, variant)
            t = []
            for i in range(reps + 3):
                flush.fill_(1)
                torch.cuda.synchronize()
                mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)
                mv.sync()
                if i >= 3:
                    t.append(mv.timings().overlay_ms)
            return float(np.mean(t)), float(np.min(t)), fin.clone(), exb.clone()

        s_mean, s_min, s_fin, s_ex = overlay_only(0)
        o_mean, o_min, o_fin, o_ex = overlay_only(1)
        mv.set_option(capi.OPT_OVERLAY_KERNEL, 0)
        hbm = measured_peaks()[0]["hbm_gbs"]
        ob = fl.OVERLAY_BYTES[1] * W * H
        print(json.dumps(dict(extra="overlay_direct", staged_ms=s_mean, staged_ms_min=s_min, direct_ms=o_mean, direct_ms_min=o_min,
                              staged_gbs=ob / (s_mean * 1e-3) / 1e9, direct_gbs=ob / (o_mean * 1e-3) / 1e9,
                              staged_frac_of_hbm_peak=ob / (s_mean * 1e-3) / 1e9 / hbm,
                              direct_frac_of_hbm_peak=ob / (o_mean * 1e-3) / 1e9 / hbm, algorithmic_bytes_per_launch=ob,
                              final_equal=bool(torch.equal(s_fin, o_fin)),
                              exponents_bit_equal=bool(torch.equal(s_ex.view(torch.int32), o_ex.view(torch.int32))),
                              note="LINON_OPT_OVERLAY_KERNEL = 1 (csrc/device/overlay_direct.cuh): four taps per pixel through the "
                                   "read-only path instead of a staged tile; opt-in until this line has been seen")), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(extra="overlay_direct", error=f"{type(e).__name__}: {e}"[:300])), flush=True)

    # Extra measurements of the same opt-in path, each on its own line ({"extra": name, ...}); the parent files them under
    # quiet_zone_probe[name].  (1) the tighter radius of the sigmoid profiles, same C3 frame against the same default frame.
    try:
        mv.set_option(capi.OPT_QUIET_RADIUS, 1500)
        r_ms, r_march, r_c, r_quiet = measure(1)
        print(json.dumps(dict(extra="radius_1500", ms_per_step=r_ms, march_ms=r_march, quiet_steps=r_quiet,
                              quiet_fraction=r_quiet / max(int(r_c.accepted_steps), 1),
                              culling_checks=int(r_c.culling_checks), vs_default_path=agreement(r_c),
                              note="LINON_OPT_QUIET_RADIUS = 1500: zone radius 1.5 max_dist instead of where the fp32 sigmoid "
                                   "saturates (1.92 max_dist)")), flush=True)
        mv.set_option(capi.OPT_QUIET_RADIUS, 0)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(extra="radius_1500", error=f"{type(e).__name__}: {e}"[:300])), flush=True)
    # (2) adaptive default-step frames (C1: 800x600, MirageSpherical) with option level 2 against level 0: march time, the
    # kernel's own step counters, hit ids and colours (endpoints of adaptive frames are not comparable)
    try:
        W1, H1 = 800, 600
        mv1 = MainView(verts, faces, center, W1, H1, discrete_gpu=True, device=dev)
        mv1.reload_shader(PredefinedFunction(1).to_code())
        mv1.update_settings(make_settings())
        rc1 = torch.as_tensor(_DevBuf(mv1.device_ptr(capi.BUF_RAYCAST_RGBA8), (H1, W1, 4), "|u1"), device="cuda")
        hid1 = torch.as_tensor(_DevBuf(mv1.device_ptr(capi.BUF_HIT_ID_I32), (H1, W1), "<i4"), device="cuda")

        def c1(level):
            mv1.set_option(capi.OPT_QUIET_ZONE, level)
            times = []
            for i in range(6):
                mv1.render_flags(0)
                mv1.sync()
                if i >= 2:
                    times.append(mv1.timings().march_ms)
            k = mv1.counters()
            return float(np.mean(times)), int(k.accepted_steps), int(k.rejected_steps), rc1.clone(), hid1.clone()

        t0, a0, j0, rc_0, hid_0 = c1(0)
        t2, a2, j2, rc_2, hid_2 = c1(2)
        same = hid_0 == hid_2
        cd = (rc_0.to(torch.int16) - rc_2.to(torch.int16)).abs().amax(dim=-1)
        print(json.dumps(dict(extra="c1_adaptive_level2", width=W1, height=H1, function="MirageSpherical",
                              march_ms_level0=t0, march_ms_level2=t2,
                              steps_per_ray_level0=dict(accepted=a0 / (W1 * H1), rejected=j0 / (W1 * H1)),
                              steps_per_ray_level2=dict(accepted=a2 / (W1 * H1), rejected=j2 / (W1 * H1)),
                              hit_id_equal=float(same.sum()) / (W1 * H1),
                              colour_within_1_where_hit_equal=float(((cd <= 1) & same).sum()) / max(int(same.sum()), 1))),
              flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps(dict(extra="c1_adaptive_level2", error=f"{type(e).__name__}: {e}"[:300])), flush=True)
    return 0


def parse_probe_output(stdout, failure):
    """First JSON line of the probe child = the main result; later {"extra": name, ...} lines are filed under their name."""
    probe = None
    for ln in (stdout or "").splitlines():
        if not ln.startswith("{"):
            continue
        try:
            obj = json.loads(ln)
        except ValueError:
            continue
        if probe is None and "extra" not in obj:
            probe = obj
        elif probe is not None and "extra" in obj:
            probe[obj.pop("extra")] = obj
    if probe is None:
        return dict(error=failure)
    if failure:
        probe["incomplete"] = failure
    return probe


def run_probe_child(local_rank, timeout=300):
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--probe-quiet-zone", "--device", str(local_rank)],
                           capture_output=True, text=True, timeout=timeout)
        return parse_probe_output(r.stdout, "" if r.returncode == 0 else f"exit {r.returncode}: {r.stderr[-300:]}")
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else e.stdout
        return parse_probe_output(out, f"time limit of {timeout} s reached")
    except Exception as e:  # noqa: BLE001
        return dict(error=f"{type(e).__name__}: {e}"[:400])


def run_ours(args):
    import numpy as np
    import torch

    from linon_b200 import _capi as capi
    from linon_b200 import flops as fl
    from linon_b200 import sharding
    from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings, measure_fp32_peak

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # the contract is ONE JSON line on stdout: NCCL_DEBUG=VERSION (set on the GPU boxes) would print its banner there
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H = workload_size(world)
    verts, faces, center = default_scene()
    mv = MainView(verts, faces, center, W, H, discrete_gpu=True, device=local_rank)
    mv.reload_shader(PredefinedFunction(FN_ID).to_code())
    settings = make_settings(overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)
    mv.update_settings(settings)
    mv.set_shard(rank, world, ROWS_PER_BLOCK)
    if args.quiet_zone:
        mv.set_option(capi.OPT_QUIET_ZONE, 1)
    peer_stores = bool(args.peer_stores) and world > 1
    if peer_stores:
        # no gather: every rank's march stores its rows straight into rank 0's frame buffers over NVLink (CUDA IPC mappings of
        # rank 0's buffers, include/linon_b200.h: linon_ipc_export / linon_ipc_open / linon_redirect_output)
        sharding.setup_peer_stores(mv, rank, world)
    stream = torch.cuda.ExternalStream(mv.stream(), device=torch.device("cuda", local_rank))

    ep = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_ENDPOINT_F32X4), (H, W, 4), "<f4"), device="cuda")
    rc = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_RAYCAST_RGBA8), (H, W, 4), "|u1"), device="cuda")
    hid = torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_HIT_ID_I32), (H, W), "<i4"), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    final_host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True) if rank == 0 else None

    march_flags = capi.RENDER_ENHANCE
    if world > 1:
        march_flags |= capi.RENDER_NO_OVERLAY
        if rank != 0:
            march_flags |= capi.RENDER_NO_SAMPLE_RAYS

    last = {}

    def frame():
        """One step of the hot path on this rank; returns launches of our kernels."""
        mv.render_flags(march_flags)
        n = 1 + (2 if not (march_flags & capi.RENDER_NO_SAMPLE_RAYS) else 0)  # march (+ click positions + sample rays)
        if world > 1:
            mv.sync()
            last["timings"], last["counters"] = mv.timings(), mv.counters()  # of the march render (reset per render)
            last["quiet"] = int(mv.debug_profile()[15]) if args.quiet_zone else 0
            if peer_stores:
                dist.barrier()  # every rank's march has completed (mv.sync above): rank 0's buffers hold the whole frame
            else:
                with torch.cuda.stream(stream):
                    sharding.gather_rows_to_root([rc, ep, hid], H, rank, world, ROWS_PER_BLOCK)
            if rank == 0:
                mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)  # overlay on the assembled image
                n += 1
        else:
            n += 1  # overlay
        return n

    def barrier():
        mv.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # JIT + warm-up (untimed)
    for _ in range(max(args.warmup, 3 if args.steps > 0 else 0)):
        frame()
        barrier()

    # ---- device-timed steps: CUDA events on the stream the kernels are launched on; L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches = 0
    dev_ms = []
    march_ms = []
    overlay_ms = []
    counters = None
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches += frame()
        e1.record(stream)
        barrier()
        dev_ms.append(e0.elapsed_time(e1))
        t = mv.timings()
        overlay_ms.append(t.overlay_ms)
        if world > 1:
            t, counters = last["timings"], last["counters"]
        else:
            counters = mv.counters()
            last["quiet"] = int(mv.debug_profile()[15]) if args.quiet_zone else 0
        march_ms.append(t.march_ms)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = float(sum(dev_ms))
    tm = torch.tensor([total_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())

    # ---- e2e: public API with host buffers (uniform upload + final image read back) inside the timed region ----
    e2e_ms = 0.0
    barrier()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        mv.update_camera()                 # 112-byte camera uniform from the host-side arcball (MainView::update_camera)
        mv.update_settings(settings)       # 24-byte settings uniform (MainView::update_settings)
        frame()
        if rank == 0:
            # result read back into pinned host memory (what a headless caller of the reference-facing API receives)
            mv.read_into_ptr(capi.BUF_FINAL_RGBA8, final_host.data_ptr(), final_host.numel())
        barrier()
        e2e_ms += (time.perf_counter() - t0) * 1e3
    te = torch.tensor([e2e_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())

    # counters of all ranks (rays / steps processed by the whole job in one step)
    cvec = torch.tensor([counters.rays, counters.accepted_steps, counters.rejected_steps, counters.prehit_steps,
                         counters.face_tests, counters.plane_tests, last.get("quiet", 0)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(cvec)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    rays, acc, rej, pre, ftests, ptests, quiet_steps = [float(x) for x in cvec.tolist()]

    steps = max(args.steps, 1)
    ms_per_step = total_ms / steps
    value = rays / (ms_per_step * 1e-3) / 1e6
    gsteps = (acc + rej) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the march): FP32 work actually executed / live launch duration ----
    # Two numerators, the SMALLER one is `achieved`: (i) the hand count of the executed arithmetic
    # (linon_b200/flops.py: fused mirage body, three evaluated stages, only the plane / triangle tests that ran);
    # (ii) FADD + FMUL + 2 FFMA per RK4 step as ncu counted them for this exact kernel variant
    # (profiles/r1_march_c3.json, committed), times the steps of this launch.
    fp32_peak, sms = measure_fp32_peak(local_rank)
    m_ms = float(np.mean(march_ms))
    model_flops = fl.fused_march_flops(counters, FN_ID, field_weight_is_one=True, quiet_steps=last.get("quiet", 0))
    # the committed ncu capture is of the default variant: with the quiet zone on only the hand count applies
    ncu_prof = ncu_profile() if not args.quiet_zone else None
    ncu_flops = ncu_prof["fp32_flops_per_rk4_step"] * (counters.accepted_steps + counters.rejected_steps) if ncu_prof else None
    alg_flops = min(model_flops, ncu_flops) if ncu_flops else model_flops
    achieved = alg_flops / (m_ms * 1e-3) / 1e12
    peaks, peak_src = measured_peaks()
    o_ms = float(np.mean(overlay_ms)) if overlay_ms and overlay_ms[0] > 0 else None
    overlay_roof = None
    if o_ms:
        ob = fl.OVERLAY_BYTES[1] * W * H
        a = ob / (o_ms * 1e-3) / 1e9
        overlay_roof = dict(bound="hbm", achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"],
                            traffic=overlay_ncu_traffic(W, H), kernel="overlay_kernel", algorithmic_bytes_per_launch=ob,
                            peak_source=peak_src)

    line = dict(
        metric="Mrays/s", value=value, unit="Mrays/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        rk4_gsteps_per_s=gsteps,
        config=dict(workload=workload_name(world, W, H), width=W, height=H, rays_per_step=rays,
                    rk4_steps_per_step=acc + rej, field_function="MirageSphericalSigmoid", h_initial=0.001,
                    math_mode="fast (--use_fast_math; whole-step restatement of the predefined mirage body: RK4 stage 1 = identity, scaled coordinates, one reciprocal per stage)",
                    kernel="linon_march_tiles (one thread per pixel, 8x4 tiles, converged fixed-step loop split by warp phase, full-mask votes, two-pass culling checks)", plane_culling=True,
                    parallelism=(f"row-blocks x{world}" + (", peer stores into rank 0 over NVLink (no gather)" if peer_stores else ", grouped NCCL send/recv to rank 0")) if world > 1 else "single GPU",
                    l2="flushed between steps (256 MiB write); outputs 199 MB/frame exceed L2, inputs are 1.5 KB of uniforms"),
        clocks=clocks,
        e2e=dict(value=rays / (e2e_ms / steps * 1e-3) / 1e6, unit="Mrays/s", ms_per_step=e2e_ms / steps,
                 h2d_bytes_per_step=112 + 24, d2h_bytes_per_step=W * H * 4),
        gpu_launches=launches,
        roofline=dict(bound="fp32", achieved=achieved, peak=fp32_peak, unit="TFLOP/s", frac=achieved / fp32_peak,
                      traffic=ncu_prof.get("dram_traffic_bytes") if ncu_prof else None,
                      kernel="linon_march_tiles", algorithmic_flops_per_launch=alg_flops, launch_ms=m_ms,
                      peak_source="FFMA-chain microbenchmark measured in this run (MEASURED_PEAKS.json has no fp32 entry; nominal 74.45)",
                      frac_of_nominal=achieved / 74.45,
                      numerators=dict(model_executed_flops=model_flops, ncu_executed_flops=ncu_flops,
                                      as_written_equivalent_flops=fl.as_written_flops(counters, FN_ID)),
                      issue_slot_utilization_ncu=(ncu_prof.get("sm__inst_issued.avg.pct_of_peak_sustained_active") if ncu_prof else None),
                      algorithmic_bytes_per_launch=fl.MARCH_BYTES_PER_PIXEL * rays,
                      note="issue-slot bound fp32 CUDA-core kernel (no dense contraction; HBM traffic is 24 B per 5001-step ray): "
                           "the binding resource is the scheduler's one instruction per cycle (issue_slot_utilization_ncu, "
                           "from the committed ncu capture of this variant), not the FMA pipe; the numerator is work actually "
                           "executed (smaller of hand count and ncu count), never culled or skipped work; traffic = ncu DRAM bytes"),
        roofline_overlay=overlay_roof,
        counters=dict(accepted=acc, rejected=rej, prehit=pre, face_tests=ftests, plane_tests=ptests,
                      culling_checks=float(counters.culling_checks), quiet_steps=quiet_steps),
        wall_s=t_wall,
    )
    if args.quiet_zone:
        line["config"]["quiet_zone"] = ("LINON_OPT_QUIET_ZONE on: straight steps (velocity = direction) while the whole warp is outside "
                                        "the mirage's zone of influence; every step executed and counted (counters.quiet_steps)")
    if world == 1 and not args.no_cpu_baseline:
        s = oracle_sample(W, H, args.cpu_seconds)
        line["cpu_baseline"] = dict(value=s["rays"] / s["seconds"] / 1e6, unit="Mrays/s", cores=s["cores"], kind=s["kind"],
                                    sample=s["sample"], rk4_gsteps_per_s=s["steps"] / s["seconds"] / 1e9,
                                    note="the reference's WGSL run on the host cores (oracle/_ref transpile, or the bit-identical hand port), not wgpu-on-lavapipe: no Rust/Vulkan in this image")
    if world == 1 and not args.quiet_zone and not args.no_quiet_zone_probe:
        # LAST GPU work of the run: the opt-in quiet-zone path on the same frame, timed and compared with the default path's
        # frame on the device.  A reported probe, not the headline: the option stays off until this has been seen on a B200.
        # In a child process with a time limit, so that nothing the unmeasured path does can take this line down.
        line["quiet_zone_probe"] = run_probe_child(local_rank)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quiet-zone", type=int, default=0, choices=[0, 1],
                    help="1: measure with LINON_OPT_QUIET_ZONE on (straight steps outside the mirage's zone of influence); "
                         "0 (default until the option is validated on the GPU): the default path is measured and the option is "
                         "only probed at the end of the N=1 run (quiet_zone_probe)")
    ap.add_argument("--no-quiet-zone-probe", action="store_true")
    ap.add_argument("--peer-stores", type=int, default=0, choices=[0, 1],
                    help="N > 1: 1 = every rank's march stores straight into rank 0's frame buffers (CUDA IPC + NVLink) instead of "
                         "the NCCL gather after the march (opt-in: not yet run on a multi-GPU box)")
    ap.add_argument("--probe-quiet-zone", action="store_true", help="internal: the child process of the N=1 run")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    if args.probe_quiet_zone:
        return run_quiet_zone_probe(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
