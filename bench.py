#!/usr/bin/env python
"""bench.py — headline benchmark of the linon ray-march path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own WGSL on the host cores

One "step" = one frame of the hot path: [click positions] -> march (+ 8 sample rays on a second stream) -> Lyapunov
overlay.  N = 1 runs BASELINE.json configs[2] (C3): 3840x2160, Enhance step size (h = 0.001, 5001 RK4 steps per ray),
overlay on, the application's default field function (MirageSphericalSigmoid), default camera / mesh / settings.
N > 1 is weak scaling of the same frame: N times the pixels over the same field of view (round(3840 sqrt N) x
round(2160 sqrt N), so the per-ray work mix is unchanged), rows cut into blocks dealt round-robin to the ranks; the
row blocks reach rank 0 either by peer stores during the march or by one grouped NCCL exchange after it, overlay on
rank 0.  Every N also renders BASELINE.json configs[4]'s frame (C5: 8192x8192, Enhance) sharded over the same ranks:
`c5_strong` is the fixed-size record a strong-scaling curve is read from.

Prints ONE JSON line (rank 0).  value = Mrays/s of the whole job (rays of all ranks / max-over-ranks device time);
e2e = the same through the public MainView API with host buffers (camera + settings upload, final image read back into
pinned memory) inside the timed region.

The N = 1 line also carries, from the same process and the same frame:
  parity     the GPU frame graded against the oracle on sampled rows of the FULL C3 frame (oracle/sampled_rows.py: bands
             of three rows, centre rows graded, high-Lyapunov pixels from the oracle's own endpoints excluded and
             counted, the excluded fraction capped) - for the shipped path, for the round-1 path (quiet zone off) and
             for the generic translation of the same WGSL text (LINON_OPT_SPECIALIZE = 0); the oracle itself is
             compared bit for bit with the reference's transpiled WGSL (oracle/_ref) on those rows
  variants   frame times of those two other paths beside the headline
  configs    BASELINE.json configs[0], [1], [3] (C1, C2 over the nine predefined functions, C4): kernel times, the
             kernel's own step counters and the same sampled-row parity (child process)
The oracle is the CHECKER of frames the CUDA path produced and, in `cpu_baseline` / `--impl reference`, the CPU arm;
nothing it computes feeds the product path.
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
C5_W, C5_H = 8192, 8192
ROWS_PER_BLOCK = 32
DEFAULT_QUIET = 2     # the library's default level of LINON_OPT_QUIET_ZONE (include/linon_b200.h)
DEFAULT_QUIET_RADIUS = 1500  # ... and of LINON_OPT_QUIET_RADIUS
STEPS_PER_RAY = 5001  # h_initial = 0.001: the loop of src/main_view.wgsl:244 runs 5001 times; adaptive sampling cannot trigger
MARCH_PROFILE = os.path.join(ROOT, "profiles", "r2_march_c3.json")
MARCH_PROFILE_STEP_BY_STEP = os.path.join(ROOT, "profiles", "r2_march_c3_step_by_step.json")  # LINON_OPT_COLLAPSE_RUNS = 0
OVERLAY_PROFILE = os.path.join(ROOT, "profiles", "r2_overlay_c3.json")


def workload_size(n_gpus: int):
    s = math.sqrt(n_gpus)
    return int(round(BASE_W * s)), int(round(BASE_H * s))


def workload_name(n_gpus: int, w: int, h: int) -> str:
    base = "C3: Cornell box, MirageSphericalSigmoid, Enhance h=0.001 (5001 RK4 steps/ray), Lyapunov overlay delta=1 scaling=50"
    if n_gpus == 1:
        return f"{base}, {w}x{h}, 1xB200"
    return f"{base}, weak-scaled to {w}x{h} ({n_gpus}x the pixels, same field of view), {ROWS_PER_BLOCK}-row blocks round-robin over {n_gpus} GPUs, frame assembled on rank 0"


def workload_config(n_gpus: int, w: int, h: int) -> dict:
    """The `config` object: names the WORKLOAD only, identical in both arms (ours / --impl reference)."""
    return dict(workload=workload_name(n_gpus, w, h), width=w, height=h, rays_per_step=w * h,
                rk4_steps_per_step=w * h * STEPS_PER_RAY, field_function="MirageSphericalSigmoid", h_initial=0.001,
                overlay="mode 1, delta 1, scaling 50")


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
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
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


def static_source_hash() -> str:
    """Identity of the ahead-of-time kernels (overlay, outline search, click positions): sha256 of their sources."""
    import hashlib
    h = hashlib.sha256()
    for rel in ("static_kernels.cu", "overlay_direct.cuh", "overlay_tile.cuh", "outline_keys.cuh"):
        with open(os.path.join(ROOT, "linon_b200", "csrc", "device", rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def nvrtc_version():
    """The NVRTC the library compiled the timed kernels with (opened by absolute path: torch's own copy is not used)."""
    from linon_b200.main_view import nvrtc_info
    path, (major, minor) = nvrtc_info()
    return f"{major}.{minor} ({path})"


def committed_profile(path: str, kernel_substring: str, identity: str | None, key: str = "cubin_hash"):
    """A committed ncu summary (tools/ncu_summary.py) is evidence about ONE binary: it names what it profiled - the FNV-1a
    hash of the NVRTC cubin (`cubin_hash`, march kernels) or the hash of the sources of the ahead-of-time kernels
    (`source_hash`) - and is refused, with the reason, when the kernel running now is another one."""
    if not os.path.exists(path):
        return None, f"{os.path.relpath(path, ROOT)} not committed"
    for k in json.load(open(path)):
        if kernel_substring in str(k.get("kernel", "")):
            if identity is not None and k.get(key) != identity:
                return None, (f"{os.path.relpath(path, ROOT)} profiled {key} {k.get(key)}, the running kernel is "
                              f"{identity}: stale capture refused")
            return k, "ok"
    return None, f"no {kernel_substring} entry in {os.path.relpath(path, ROOT)}"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """Threads the CPU arm uses: every core this process may run on.  torchrun exports OMP_NUM_THREADS=1, which would make
    the OpenMP default one thread - the count is therefore always passed explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference():
    """The CPU implementation timed beside the GPU path.  oracle/_ref (the reference's own WGSL, transpiled mechanically
    to C++ by oracle/wgsl_ref and compiled with gcc; built where /root/reference is mounted, the .so travels) when it
    exists -> kind "reference"; otherwise the hand-written restatement oracle/linon_oracle.cpp -> kind "port".  The two
    are bit-identical (tests/test_wgsl_reference_cpu.py; re-checked on the sampled rows of every N = 1 run)."""
    from oracle import wgsl_reference as wr
    if wr.available():
        wr.build()
        return wr, "reference"
    from oracle import cpu_oracle as co
    co.build()
    return co, "port"


def c3_inputs(w: int, h: int):
    from oracle import np_oracle as npo
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, w, h))
    st = dict(npo.default_settings(), overlay_mode=1)
    return verts, faces, cam, st


def oracle_sample(w: int, h: int, budget_s: float, threads: int = 0, keep_outputs: bool = False):
    """Times the CPU implementation of the reference's path on a bounded sample of the same frame: evenly spaced full rows
    (pixels are independent; cost is per ray), sized by a short calibration run to ~budget_s seconds.  The rows are the
    centre rows of oracle/sampled_rows.band_centres, so the parity block can grade exactly these rows."""
    from oracle import sampled_rows as sr

    impl, kind = cpu_reference()
    verts, faces, cam, st = c3_inputs(w, h)
    cores = threads if threads > 0 else host_threads()

    def run(rows):
        centres, stride = sr.band_centres(h, rows)
        t0 = time.perf_counter()
        out = sr.march_bands(verts, faces, cam, st, FN_ID, w, h, centres, stride, h_initial=0.001, adaptive=True, delta=0,
                             n_threads=cores, impl=impl)
        dt = time.perf_counter() - t0
        rays = len(centres) * w
        if "accepted" in out:  # the port counts its steps; the transpiled shader has no such output
            assert int(out["accepted"].sum() + out["rejected"].sum()) == rays * STEPS_PER_RAY
        return dt, rays, rays * STEPS_PER_RAY, centres, stride, out

    # calibration on a strided part of one row
    xs = max(1, w // (cores * 16))
    t0 = time.perf_counter()
    impl.march(verts, faces, cam, st, FN_ID, w, h, h_initial=0.001, adaptive=True, rect=(0, w, h // 2, h // 2 + 1),
               x_step=xs, n_threads=cores)
    per_ray = (time.perf_counter() - t0) / len(range(0, w, xs))
    rows = max(1, min(h // 8, int(budget_s / (per_ray * w))))
    t, rays, steps, centres, stride, out = run(rows)
    if t < 0.4 * budget_s and len(centres) < h // 8:
        # the calibration (a few rays per thread, thread start-up included) overestimates the cost per ray: size the sample
        # from the rate just measured and time that one instead
        t, rays, steps, centres, stride, out = run(max(len(centres) + 1, min(h // 8, int(len(centres) * 0.8 * budget_s / t))))
    what = ("the reference's main_view.wgsl + MirageSphericalSigmoid body transpiled to C++ (oracle/_ref), gcc -O2 -ffp-contract=off"
            if kind == "reference" else "hand-written C++ restatement of the reference WGSL (oracle/linon_oracle.cpp)")
    res = dict(seconds=t, rays=rays, steps=steps, rows=len(centres), cores=cores, kind=kind, centres=centres, stride=stride,
               sample=f"{len(centres)} full rows of the {w}x{h} frame, evenly spaced ({rays} rays, {steps} RK4 steps), {cores} OpenMP threads; {what}")
    if keep_outputs:
        res["outputs"] = out
    return res


def run_reference(args):
    """--impl reference: the reference's own implementation of the path on the host cores.  The Rust + wgpu host cannot
    be built or run in this image (no rustc/cargo, no Vulkan ICD / lavapipe); what runs is the reference's
    shader itself: main_view.wgsl with the MirageSphericalSigmoid body of functions.rs spliced in, transpiled
    mechanically to C++ and compiled with gcc (oracle/_ref, kind = "reference"; falls back to the hand-written port when
    that library was never built), all host threads (counted from the affinity mask: torchrun's OMP_NUM_THREADS=1 is
    not obeyed), on a bounded sample of the same frame per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, h = workload_size(args.gpus)
    threads = host_threads()
    budget = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        oracle_sample(w, h, budget, threads)
    tot_t = tot_rays = tot_steps = 0
    last = None
    for _ in range(args.steps):
        last = oracle_sample(w, h, budget, threads)
        tot_t += last["seconds"]; tot_rays += last["rays"]; tot_steps += last["steps"]
    mrays = tot_rays / tot_t / 1e6
    line = dict(impl="reference", metric="Mrays/s", value=mrays, unit="Mrays/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=tot_t / args.steps * 1e3 * (w * h) / (tot_rays / args.steps),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                rk4_gsteps_per_s=tot_steps / tot_t / 1e9,
                config=workload_config(args.gpus, w, h),
                cpu_baseline=dict(value=mrays, unit="Mrays/s", cores=last["cores"], kind=last["kind"], sample=last["sample"]),
                e2e=dict(value=mrays, unit="Mrays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                note="ms_per_step is the sampled time scaled to the full frame (pixels are independent).  niklaskorz/linon's host is "
                     "Rust+wgpu (no rustc/cargo/Vulkan in this image): its WGSL compute shader is run on the host cores through a "
                     "mechanical WGSL->C++ transpile, not through wgpu on lavapipe")
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
class _DevBuf:
    """Exposes a device pointer of the library to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3,
                                             strides=None)


def frame_buffers(mv):
    """The march outputs of the view's last frame on the host (hit id, colour, endpoint)."""
    return dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())


def timed_frames(mv, flags, n, warm=1, flush=None):
    """Mean frame time (CUDA events on the library's stream) and the kernel times of n frames after `warm` untimed ones."""
    import numpy as np
    import torch
    stream = torch.cuda.ExternalStream(mv.stream())
    for _ in range(warm):
        mv.render_flags(flags)
        mv.sync()
    frame_ms, march_ms, overlay_ms, sample_ms = [], [], [], []
    for _ in range(n):
        if flush is not None:
            flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mv.render_flags(flags)
        e1.record(stream)
        mv.sync()
        torch.cuda.synchronize()
        t = mv.timings()
        frame_ms.append(e0.elapsed_time(e1)); march_ms.append(t.march_ms); overlay_ms.append(t.overlay_ms)
        sample_ms.append(t.sample_rays_ms)
    return dict(ms_per_step=float(np.mean(frame_ms)), march_ms=float(np.mean(march_ms)), overlay_ms=float(np.mean(overlay_ms)),
                sample_rays_ms=float(np.mean(sample_ms)), frames=n)


def c3_parity_and_variants(mv, W, H, cpu, parity_seconds, flush):
    """N = 1, after the timed loops: (i) the oracle on bands around the rows the cpu_baseline leg just computed with the
    reference's transpiled WGSL; (ii) the frame of the shipped path, of the round-1 path (LINON_OPT_QUIET_ZONE = 0) and of
    the generic translation (LINON_OPT_SPECIALIZE = 0), each rendered here on the device and graded against those rows."""
    import numpy as np

    from linon_b200 import _capi as capi
    from oracle import cpu_oracle as co
    from oracle import sampled_rows as sr

    verts, faces, cam, st = c3_inputs(W, H)
    threads = host_threads()
    # bands: as many as the budget allows (3 rows per graded row), never more than the rows the cpu_baseline leg sampled
    if cpu is not None:
        rate = cpu["rays"] / cpu["seconds"]
        n_bands = max(4, min(cpu["rows"], int(parity_seconds * rate / (3 * W))))
        k = max(1, -(-cpu["rows"] // n_bands))
        centres, stride = cpu["centres"][::k], cpu["stride"] * k
    else:
        centres, stride = sr.band_centres(H, 8)
    t0 = time.perf_counter()
    ref = sr.march_bands(verts, faces, cam, st, FN_ID, W, H, centres, stride, h_initial=0.001, adaptive=True, delta=1, n_threads=threads)
    oracle_s = time.perf_counter() - t0
    graded = sr.graded_mask(W, H, centres)
    high = sr.high_lyapunov_on_centres(ref, W, H, centres, dict(st, central_difference_delta=1, lyapunov_scaling=50.0)) & graded
    rf = dict(ref=ref, graded=graded, high_lyapunov=high, undecided=np.zeros_like(high), excluded=high, centres=centres, stride=stride)
    pinned = None
    if cpu is not None and cpu["kind"] == "reference" and "outputs" in cpu:
        shader = cpu["outputs"]  # the reference's transpiled WGSL on the same rows (the timed cpu_baseline run itself)
        pinned = bool(np.array_equal(shader["rgba8"][graded], ref["rgba8"][graded]) and
                      np.array_equal(shader["endpoint"][graded].view(np.uint32), ref["endpoint"][graded].view(np.uint32)))
    out = dict(
        checker="oracle/linon_oracle.cpp (fp32 restatement of src/main_view.wgsl, hit ids per SURVEY F10) on bands of 3 rows; centre rows graded",
        rows=len(centres), pixels=int(graded.sum()), oracle_seconds=round(oracle_s, 2), oracle_threads=threads,
        oracle_equals_transpiled_reference_wgsl_on_these_rows=pinned,
        hit_ids="compared per surface group (coplanar Cornell faces, SURVEY F11); raw agreement reported beside it",
        max_excluded_fraction=0.02)
    variants = {}
    E = capi.RENDER_ENHANCE
    shipped_quiet, shipped_spec = DEFAULT_QUIET, 1

    def one(name, quiet, spec, frames, radius=DEFAULT_QUIET_RADIUS, collapse=1):
        mv.set_option(capi.OPT_QUIET_ZONE, quiet)
        mv.set_option(capi.OPT_SPECIALIZE, spec)
        mv.set_option(capi.OPT_QUIET_RADIUS, radius)
        mv.set_option(capi.OPT_COLLAPSE_RUNS, collapse)
        t = timed_frames(mv, E, frames, warm=1, flush=flush)
        c = mv.counters()
        assert int(c.accepted_steps) == W * H * STEPS_PER_RAY and int(c.rejected_steps) == 0, (name, c.accepted_steps)
        m = sr.grade(frame_buffers(mv), rf, fixed_step=True, max_excluded_fraction=0.02)
        q = int(mv.debug_profile()[15])
        t.update(mrays_s=W * H / (t["ms_per_step"] * 1e-3) / 1e6, cubin_hash=mv.march_cubin_hash(),
                 quiet_steps=q, culling_checks=int(c.culling_checks), face_tests=int(c.face_tests))
        if quiet and spec:  # executed-work model of the fused path (linon_b200/flops.py), for the roofline of this variant
            from linon_b200 import flops as fl
            t["model_executed_flops"] = fl.fused_march_flops(c, FN_ID, field_weight_is_one=True, quiet_steps=q, collapsed_runs=bool(collapse))
        return t, m

    try:
        t, m = one("shipped", shipped_quiet, shipped_spec, 2)
        out["shipped"] = m
        t, m = one("runs_step_by_step", shipped_quiet, shipped_spec, 2, collapse=0)
        variants["runs_step_by_step"] = dict(t, note="LINON_OPT_COLLAPSE_RUNS = 0: every straight step of a quiet run is executed as one FFMA per "
                                                     "coordinate instead of binade by binade (bit-identical frames); the FP32 roofline fraction "
                                                     "of this variant is reported beside the default's (roofline.step_by_step)")
        out["runs_step_by_step"] = m
        t, m = one("quiet_radius_saturation", shipped_quiet, shipped_spec, 2, radius=0)
        variants["quiet_radius_saturation"] = dict(t, note="LINON_OPT_QUIET_RADIUS = 0: the sigmoid zone ends where the reference's fp32 sigmoid "
                                                           "saturates (1.92 max_dist) and its own evaluation IS the identity, instead of the shipped "
                                                           "1.5 max_dist, beyond which the real-number signal is 50x below the fp32 quantum")
        out["quiet_radius_saturation"] = m
        t, m = one("quiet_zone_off", 0, 1, 2)
        variants["quiet_zone_off"] = dict(t, note="LINON_OPT_QUIET_ZONE = 0: every step evaluates the field (the round-1 headline kernel)")
        out["quiet_zone_off"] = m
        t, m = one("generic_translation", 0, 0, 2)
        variants["generic_translation"] = dict(t, note="LINON_OPT_SPECIALIZE = 0: the WGSL body through the generic translator, four field "
                                                       "evaluations per step - the path of a user body no structural match recognises")
        out["generic_translation"] = m
    finally:
        mv.set_option(capi.OPT_QUIET_ZONE, shipped_quiet)
        mv.set_option(capi.OPT_SPECIALIZE, shipped_spec)
        mv.set_option(capi.OPT_QUIET_RADIUS, DEFAULT_QUIET_RADIUS)
        mv.set_option(capi.OPT_COLLAPSE_RUNS, 1)
    out["green"] = bool(out.get("shipped", {}).get("green")) and pinned is not False
    return out, variants


# ----------------------------------------------------------------------------------------------------------------
def run_configs_child(args):
    """--configs-child: BASELINE.json configs[0], [1], [3] on the device (C1; C2 = the nine predefined functions at
    1920x1080; C4 = the outline workload), each with kernel times, the kernel's own counters and sampled-row parity
    against the oracle.  Prints one JSON object per config line, then {"done": true}."""
    import numpy as np
    import torch

    from linon_b200 import _capi as capi
    from linon_b200.main_view import MainView, PredefinedFunction, default_scene, make_settings
    from oracle import np_oracle as npo
    from oracle import sampled_rows as sr

    torch.cuda.set_device(args.device)
    verts, faces, center = default_scene()
    threads = host_threads()
    PRE = [f for f in PredefinedFunction if f != PredefinedFunction.Custom]

    def measure(mv, flags, reps=3):
        best = None
        for _ in range(reps + 1):
            mv.render_flags(flags)
            mv.sync()
            t = mv.timings()
            if best is None or t.march_ms < best.march_ms:
                best = t
        c = mv.counters()
        W, H = mv.size
        steps = c.accepted_steps + c.rejected_steps
        return dict(march_ms=round(best.march_ms, 4), overlay_ms=round(best.overlay_ms, 4), sample_rays_ms=round(best.sample_rays_ms, 4),
                    outline_search_ms=round(best.outline_search_ms, 4), mrays_s=round(W * H / best.march_ms / 1e3, 2),
                    gsteps_s=round(steps / best.march_ms / 1e6, 2), regs=best.march_regs,
                    accepted_per_ray=round(c.accepted_steps / c.rays, 2), rejected_per_ray=round(c.rejected_steps / c.rays, 3),
                    prehit_fraction=round(c.prehit_steps / max(c.accepted_steps, 1), 4),
                    face_tests_per_prehit_step=round(c.face_tests / max(c.prehit_steps, 1), 4),
                    quiet_steps_fraction=round(int(mv.debug_profile()[15]) / max(c.accepted_steps, 1), 4))

    def graded(mv, fn, W, H, adaptive, n_bands, st=None):
        cam = npo.camera_uniform(npo.default_camera(center, W, H))
        std = dict(npo.default_settings()) if st is None else st
        rf = sr.reference_frame(verts, faces, cam, std, int(fn), W, H, n_bands, h_initial=0.1, adaptive=adaptive, n_threads=threads)
        m = sr.grade(dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture()), rf,
                     fixed_step=not adaptive, max_excluded_fraction=sr.excluded_cap(int(fn), adaptive))
        return {k: m[k] for k in ("rows", "compared", "hit_group_agree", "hit_raw_agree", "colour_within_1", "endpoint_within_1e3",
                                  "excluded_high_lyapunov", "excluded_undecided_by_the_oracle", "excluded_fraction", "green")}

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    # C1: 800x600, first predefined function, default step size
    with MainView(verts, faces, center, 800, 600, device=args.device) as mv:
        fn = PredefinedFunction.MirageSpherical
        mv.reload_shader(fn.to_code())
        mv.update_settings(make_settings())
        for name, flags, adaptive in (("adaptive on (as the application renders)", 0, True), ("adaptive off (51 steps)", capi.RENDER_ADAPTIVE_OFF, False)):
            r = measure(mv, flags)
            r["parity"] = graded(mv, fn, 800, 600, adaptive, 24)
            emit(config="C1", size="800x600", function=fn.name, variant=name, **r)
    # C2: 1920x1080, every predefined function
    with MainView(verts, faces, center, 1920, 1080, device=args.device) as mv:
        mv.update_settings(make_settings())
        for fn in PRE:
            mv.reload_shader(fn.to_code())
            variants = [("adaptive on", 0, True, None), ("adaptive off", capi.RENDER_ADAPTIVE_OFF, False, None)]
            if int(fn) <= 4:  # the mirage family defines a quiet zone: the same frame without it beside the default
                variants.insert(1, ("adaptive on, LINON_OPT_QUIET_ZONE = 0 (every attempt evaluates the field)", 0, True, 0))
            for name, flags, adaptive, quiet in variants:
                mv.set_option(capi.OPT_QUIET_ZONE, DEFAULT_QUIET if quiet is None else quiet)
                r = measure(mv, flags)
                r["parity"] = graded(mv, fn, 1920, 1080, adaptive, 8)
                emit(config="C2", size="1920x1080", function=fn.name, variant=name, **r)
            mv.set_option(capi.OPT_QUIET_ZONE, DEFAULT_QUIET)
        # C4: overlay + sample rays, click mode and outline mode, kernels timed separately
        fn = PredefinedFunction.MirageSphericalSigmoid
        mv.reload_shader(fn.to_code())
        for mouse, label in (((0.5, 0.6), "click (0.5, 0.6)"), ((0.36, 0.55), "click on the mirage rim")):
            mv.update_settings(make_settings(overlay_mode=1, mouse_pos=mouse))
            emit(config="C4", size="1920x1080", function=fn.name, variant=label, **measure(mv, 0))
        # At 1920x1080 neighbouring endpoints of this frame differ so little that with the application's start-up scaling (50)
        # no exponent reaches the 0.8 threshold (measured: max 0.41) and all eight searches fall back to pixel (0, 0); the GUI's
        # slider goes from 1 to 100 (src/application.rs:315-317): its upper end puts the rim of the mirage above the threshold.
        mv.update_settings(make_settings(overlay_mode=1, lyapunov_scaling=100.0))
        r = measure(mv, capi.RENDER_OUTLINE)
        ex = mv.exponents()
        emit(config="C4", size="1920x1080", function=fn.name, variant="outline mode (8 searches over exponents >= 0.8, lyapunov_scaling 100)",
             exponents_above_threshold=int((ex >= 0.8).sum()),
             sample_positions=np.asarray(mv.read(capi.BUF_SAMPLE_POS)).round(5).tolist(), **r)
    emit(done=True)
    return 0


def run_child(flag, local_rank, timeout):
    """Runs `bench.py <flag>` in a child process with a time limit; returns the JSON objects it printed + a failure note."""
    lines, note = [], ""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), flag, "--device", str(local_rank)],
                           capture_output=True, text=True, timeout=timeout)
        stdout = r.stdout
        if r.returncode != 0:
            note = f"exit {r.returncode}: {r.stderr[-300:]}"
    except subprocess.TimeoutExpired as e:
        stdout = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        note = f"time limit of {timeout} s reached"
    except Exception as e:  # noqa: BLE001
        return [], f"{type(e).__name__}: {e}"[:400]
    for ln in (stdout or "").splitlines():
        if ln.startswith("{"):
            try:
                lines.append(json.loads(ln))
            except ValueError:
                pass
    return lines, note


def collect_configs(lines, note):
    """Files the child's lines: {"lines": [...], "complete": bool, "all_green": bool}."""
    done = any(ln.get("done") for ln in lines)
    rows = [ln for ln in lines if "config" in ln]
    graded = [ln["parity"]["green"] for ln in rows if isinstance(ln.get("parity"), dict)]
    out = dict(lines=rows, complete=bool(done), graded=len(graded), all_green=bool(graded) and all(graded))
    if note:
        out["incomplete"] = note
    return out


# ----------------------------------------------------------------------------------------------------------------
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

    verts, faces, center = default_scene()
    settings = make_settings(overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)
    peer_stores = bool(args.peer_stores) and world > 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def make_view(W, H):
        mv = MainView(verts, faces, center, W, H, discrete_gpu=True, device=local_rank)
        mv.reload_shader(PredefinedFunction(FN_ID).to_code())
        mv.update_settings(settings)
        mv.set_shard(rank, world, ROWS_PER_BLOCK)
        if args.quiet_zone is not None:
            mv.set_option(capi.OPT_QUIET_ZONE, args.quiet_zone)
        if args.overlay_kernel is not None:
            mv.set_option(capi.OPT_OVERLAY_KERNEL, args.overlay_kernel)
        if peer_stores:
            # no gather: every rank's march stores its rows straight into rank 0's frame buffers over NVLink (CUDA IPC mappings
            # of rank 0's buffers, include/linon_b200.h: linon_ipc_export / linon_ipc_open / linon_redirect_output)
            sharding.setup_peer_stores(mv, rank, world)
        bufs = [torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_RAYCAST_RGBA8), (H, W, 4), "|u1"), device="cuda"),
                torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_ENDPOINT_F32X4), (H, W, 4), "<f4"), device="cuda"),
                torch.as_tensor(_DevBuf(mv.device_ptr(capi.BUF_HIT_ID_I32), (H, W), "<i4"), device="cuda")]
        return mv, torch.cuda.ExternalStream(mv.stream(), device=torch.device("cuda", local_rank)), bufs

    W, H = workload_size(world)
    mv, stream, bufs = make_view(W, H)
    final_host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True) if rank == 0 else None

    march_flags = capi.RENDER_ENHANCE
    if world > 1:
        march_flags |= capi.RENDER_NO_OVERLAY
        if rank != 0:
            march_flags |= capi.RENDER_NO_SAMPLE_RAYS

    last = {}

    def frame(mv, stream, bufs, H):
        """One step of the hot path on this rank; returns launches of our kernels."""
        mv.render_flags(march_flags)
        n = 1 + (2 if not (march_flags & capi.RENDER_NO_SAMPLE_RAYS) else 0)  # march (+ click positions + sample rays)
        if world > 1:
            mv.sync()
            last["timings"], last["counters"] = mv.timings(), mv.counters()  # of the march render (reset per render)
            last["quiet"] = int(mv.debug_profile()[15])
            if peer_stores:
                dist.barrier()  # every rank's march has completed (mv.sync above): rank 0's buffers hold the whole frame
            else:
                with torch.cuda.stream(stream):
                    sharding.gather_rows_to_root(bufs, H, rank, world, ROWS_PER_BLOCK)
            if rank == 0:
                mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)  # overlay on the assembled image
                n += 1
        else:
            n += 1  # overlay
        return n

    def barrier(mv):
        mv.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # JIT (untimed), then the clock sampler starts: a timed region of K x 13 ms is shorter than nvidia-smi's sampling period,
    # so the warm-up frames - the same kernels under the same load - are sampled with it
    frame(mv, stream, bufs, H)
    barrier(mv)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_sampler = time.perf_counter()
    n_warm = max(args.warmup, 3 if args.steps > 0 else 0)
    while n_warm > 0 or time.perf_counter() - t_sampler < 0.6:  # at least 0.6 s of load before the timed steps
        frame(mv, stream, bufs, H)
        barrier(mv)
        n_warm -= 1

    # ---- device-timed steps: CUDA events on the stream the kernels are launched on; L2 flushed between steps ----
    launches = 0
    dev_ms, march_ms, overlay_ms = [], [], []
    counters = None
    barrier(mv)
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches += frame(mv, stream, bufs, H)
        e1.record(stream)
        barrier(mv)
        dev_ms.append(e0.elapsed_time(e1))
        t = mv.timings()
        overlay_ms.append(t.overlay_ms)
        last["overlay_kernel"] = int(t.overlay_kernel)
        if world > 1:
            t, counters = last["timings"], last["counters"]
        else:
            counters = mv.counters()
            last["quiet"] = int(mv.debug_profile()[15])
        march_ms.append(t.march_ms)
    t_wall = time.perf_counter() - t_wall0
    total_ms = float(sum(dev_ms))
    tm = torch.tensor([total_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())

    # ---- e2e: public API with host buffers (uniform upload + final image read back) inside the timed region ----
    e2e_ms = 0.0
    barrier(mv)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        mv.update_camera()                 # 112-byte camera uniform from the host-side arcball (MainView::update_camera)
        mv.update_settings(settings)       # 24-byte settings uniform (MainView::update_settings)
        frame(mv, stream, bufs, H)
        if rank == 0:
            # result read back into pinned host memory (what a headless caller of the reference-facing API receives)
            mv.read_into_ptr(capi.BUF_FINAL_RGBA8, final_host.data_ptr(), final_host.numel())
        barrier(mv)
        e2e_ms += (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()  # sampled over warm-up, the device-timed steps and the e2e steps: one continuous region under load
    te = torch.tensor([e2e_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    cubin_hash = mv.march_cubin_hash()
    march_regs = int(last["timings"].march_regs) if world > 1 else int(mv.timings().march_regs)

    # counters of all ranks (rays / steps processed by the whole job in one step)
    cvec = torch.tensor([counters.rays, counters.accepted_steps, counters.rejected_steps, counters.prehit_steps,
                         counters.face_tests, counters.plane_tests, last.get("quiet", 0), counters.culling_checks],
                        dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(cvec)
    # this rank's own straight steps of the timed frame: the C5 block below renders another frame through frame(), which
    # overwrites last[...] when world > 1 (round 2, N = 2 / 4 lines: the roofline's hand count mixed the C3 counters with the
    # 8192^2 frame's straight steps and came out negative)
    quiet_local = last.get("quiet", 0)

    # ---- C5 (BASELINE.json configs[4]): the 8192 x 8192 Enhance frame of the orbit, sharded over the same ranks: the
    # fixed-size record of every N (strong scaling is read across the N = 1, 2, 4, 8 lines) ----
    c5 = None
    if args.c5_frames > 0:
        try:
            mv5, stream5, bufs5 = make_view(C5_W, C5_H)
            for _ in range(2):
                frame(mv5, stream5, bufs5, C5_H)
                barrier(mv5)
            ms5 = []
            for _ in range(args.c5_frames):
                flush.fill_(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream5)
                frame(mv5, stream5, bufs5, C5_H)
                e1.record(stream5)
                barrier(mv5)
                ms5.append(e0.elapsed_time(e1))
            t5 = torch.tensor([float(np.mean(ms5))], device="cuda")
            if dist is not None:
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            ms = float(t5.item())
            c5 = dict(workload="C5 frame 0: 8192x8192, Enhance, MirageSphericalSigmoid, overlay on, default camera of the orbit",
                      ms_per_frame=ms, mrays_s=C5_W * C5_H / (ms * 1e-3) / 1e6,
                      rk4_gsteps_per_s=C5_W * C5_H * STEPS_PER_RAY / (ms * 1e-3) / 1e9, frames=args.c5_frames, n_gpus=world,
                      scaling="strong (fixed frame size at every N)", orbit_64_frames_s=64 * ms * 1e-3)
            mv5.close()
        except Exception as e:  # noqa: BLE001
            c5 = dict(error=f"{type(e).__name__}: {e}"[:300])

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    rays, acc, rej, pre, ftests, ptests, quiet_steps, checks = [float(x) for x in cvec.tolist()]

    steps = max(args.steps, 1)
    ms_per_step = total_ms / steps
    value = rays / (ms_per_step * 1e-3) / 1e6
    gsteps = (acc + rej) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the march): FP32 work actually executed / live launch duration ----
    # Two numerators, the SMALLER one is `achieved`: (i) the hand count of the executed arithmetic
    # (linon_b200/flops.py: fused mirage body, three evaluated stages, straight steps as 3 FMA, only the plane / triangle
    # tests that ran); (ii) FADD + FMUL + 2 FFMA of one C3 launch as ncu counted them for THIS binary (committed summary,
    # accepted only when its cubin hash equals the running kernel's and the frame is the C3 frame).
    fp32_peak, sms = measure_fp32_peak(local_rank)
    m_ms = float(np.mean(march_ms))

    model_flops = fl.fused_march_flops(counters, FN_ID, field_weight_is_one=True, quiet_steps=quiet_local, collapsed_runs=True)
    prof, prof_note = committed_profile(MARCH_PROFILE, "linon_march_tiles", cubin_hash)
    ncu_flops = None
    if prof is not None and world == 1 and prof.get("fp32_flops_executed"):
        ncu_flops = float(prof["fp32_flops_executed"])
    alg_flops = min(model_flops, ncu_flops) if ncu_flops else model_flops
    achieved = alg_flops / (m_ms * 1e-3) / 1e12
    peaks, peak_src = measured_peaks()
    o_ms = float(np.mean(overlay_ms)) if overlay_ms and overlay_ms[0] > 0 else None
    overlay_roof = None
    if o_ms:
        ob = fl.OVERLAY_BYTES[1] * W * H
        a = ob / (o_ms * 1e-3) / 1e9
        oprof, onote = committed_profile(OVERLAY_PROFILE, "overlay", static_source_hash(), "source_hash")
        overlay_roof = dict(bound="hbm", achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"],
                            traffic=(oprof.get("dram_traffic_bytes") if (oprof and (W, H) == (BASE_W, BASE_H)) else None),
                            kernel={0: "overlay_kernel (hand-staged tile)", 1: "overlay_direct_kernel (four global taps per pixel)",
                                    2: "overlay_tile_kernel (tile + halo by one TMA load per CTA, cp.async.bulk.tensor.2d + mbarrier)"}.get(last.get("overlay_kernel"), "overlay"),
                            profiled_kernel=(oprof or {}).get("kernel"), algorithmic_bytes_per_launch=ob, launch_ms=o_ms,
                            peak_source=peak_src, profile=onote)

    line = dict(
        metric="Mrays/s", value=value, unit="Mrays/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
        rk4_gsteps_per_s=gsteps,
        config=workload_config(world, W, H),
        implementation=dict(
            math_mode="fast (--use_fast_math; whole-step restatement of the predefined mirage body: RK4 stage 1 = identity, scaled coordinates, one reciprocal per stage)",
            kernel="linon_march_tiles (one thread per pixel, 8x4 tiles, converged fixed-step loop split by warp phase, full-mask votes, two-pass "
                   "culling checks, wall hits of quiet runs resolved at plan time)",
            quiet_zone=("on" if last.get("quiet", 0) else "off") + ": straight steps (velocity = direction) while the whole warp is outside the "
                       "mirage's zone of influence (1.5 max_dist: beyond it the real-number signal is 50x below the fp32 quantum); every RK4 step is "
                       "counted, and the straight steps of a run are taken binade by binade - k steps inside one binade of a coordinate are one "
                       "exact FMA, bit-identical with executing them one by one (LINON_OPT_COLLAPSE_RUNS; variants.runs_step_by_step and "
                       "roofline.step_by_step time and grade the step-by-step kernel in this same run)",
            plane_culling=True, march_cubin_hash=cubin_hash, march_regs=march_regs, nvrtc=nvrtc_version(),
            parallelism=(f"row-blocks x{world}" + (", peer stores into rank 0 over NVLink during the march (no gather)" if peer_stores else ", grouped NCCL send/recv to rank 0 after the march")) if world > 1 else "single GPU",
            l2="flushed between steps (256 MiB write); outputs 199 MB/frame exceed L2, inputs are 1.5 KB of uniforms"),
        clocks=clocks,
        e2e=dict(value=rays / (e2e_ms / steps * 1e-3) / 1e6, unit="Mrays/s", ms_per_step=e2e_ms / steps,
                 h2d_bytes_per_step=112 + 24, d2h_bytes_per_step=W * H * 4),
        gpu_launches=launches,
        roofline=dict(bound="fp32", achieved=achieved, peak=fp32_peak, unit="TFLOP/s", frac=achieved / fp32_peak,
                      traffic=(prof.get("dram_traffic_bytes") if (prof and ncu_flops) else None),
                      kernel="linon_march_tiles", algorithmic_flops_per_launch=alg_flops, launch_ms=m_ms,
                      peak_source="FFMA-chain microbenchmark measured in this run (MEASURED_PEAKS.json has no fp32 entry; nominal 74.45)",
                      frac_of_nominal=achieved / 74.45,
                      numerators=dict(model_executed_flops=model_flops, ncu_executed_flops=ncu_flops,
                                      as_written_equivalent_flops=fl.as_written_flops(counters, FN_ID)),
                      ncu_profile=prof_note,
                      issue_slot_utilization_ncu=(prof.get("sm__inst_issued.avg.pct_of_peak_sustained_active") if prof else None),
                      warp_execution_efficiency_ncu=((prof.get("smsp__thread_inst_executed_per_inst_executed.ratio") or 0) / 32 if prof else None),
                      reference_equivalent=dict(
                          rk4_gsteps_per_s_march_only=(counters.accepted_steps + counters.rejected_steps) / (m_ms * 1e-3) / 1e9,
                          as_written_tflops=fl.as_written_flops(counters, FN_ID) / (m_ms * 1e-3) / 1e12,
                          note="steps of the reference's loop completed per second, and the FLOP/s the reference's as-written "
                               "arithmetic for those steps would need (context only: never a roofline numerator)"),
                      algorithmic_bytes_per_launch=fl.MARCH_BYTES_PER_PIXEL * int(counters.rays),
                      note="issue-slot bound fp32 CUDA-core kernel (no dense contraction; HBM traffic is 24 B per 5001-step ray): "
                           "the numerator is work actually executed (smaller of hand count and, when the committed capture is of "
                           "this very binary, the ncu count), never culled or skipped work - the straight steps of collapsed runs are "
                           "credited with nothing; roofline.step_by_step is the same kernel with those steps executed one by one"),
        roofline_overlay=overlay_roof,
        counters=dict(accepted=acc, rejected=rej, prehit=pre, face_tests=ftests, plane_tests=ptests,
                      culling_checks=checks, quiet_steps=quiet_steps, quiet_fraction=quiet_steps / max(acc, 1.0)),
        c5_strong=c5,
        wall_s=t_wall,
    )
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = oracle_sample(W, H, args.cpu_seconds, keep_outputs=True)
        line["cpu_baseline"] = dict(value=cpu["rays"] / cpu["seconds"] / 1e6, unit="Mrays/s", cores=cpu["cores"], kind=cpu["kind"],
                                    sample=cpu["sample"], rk4_gsteps_per_s=cpu["steps"] / cpu["seconds"] / 1e9,
                                    note="the reference's WGSL run on the host cores (oracle/_ref transpile, or the bit-identical hand port), not wgpu-on-lavapipe: no Rust/Vulkan in this image")
    if world == 1 and not args.no_parity:
        try:
            line["parity"], line["variants"] = c3_parity_and_variants(mv, W, H, cpu, args.parity_seconds, flush)
            sbs = line["variants"].get("runs_step_by_step")
            if sbs and sbs.get("model_executed_flops"):
                # the same kernel with every straight step executed (LINON_OPT_COLLAPSE_RUNS = 0): executed FP32 work / its march time
                sprof, snote = committed_profile(MARCH_PROFILE_STEP_BY_STEP, "linon_march_tiles", sbs["cubin_hash"])
                s_ncu = float(sprof["fp32_flops_executed"]) if sprof and sprof.get("fp32_flops_executed") else None
                s_flops = min(sbs["model_executed_flops"], s_ncu) if s_ncu else sbs["model_executed_flops"]
                s_ach = s_flops / (sbs["march_ms"] * 1e-3) / 1e12
                line["roofline"]["step_by_step"] = dict(
                    achieved=s_ach, frac=s_ach / line["roofline"]["peak"], launch_ms=sbs["march_ms"],
                    numerators=dict(model_executed_flops=sbs["model_executed_flops"], ncu_executed_flops=s_ncu), ncu_profile=snote,
                    frac_on_ncu_count=(s_ncu / (sbs["march_ms"] * 1e-3) / 1e12 / line["roofline"]["peak"]) if s_ncu else None,
                    issue_slot_utilization_ncu=(sprof.get("sm__inst_issued.avg.pct_of_peak_sustained_active") if sprof else None),
                    note="LINON_OPT_COLLAPSE_RUNS = 0: the straight steps of quiet runs executed one by one (3 FMA each) - the kernel the "
                         "north star's FP32 target is about; the shipped default takes those stretches binade by binade, executes less and "
                         "finishes sooner, so its fraction of the FP32 peak is lower at a shorter frame time")
        except Exception as e:  # noqa: BLE001
            line["parity"] = dict(error=f"{type(e).__name__}: {e}"[:400], green=False)
    if world == 1 and not args.no_configs:
        mv.close()
        del flush
        torch.cuda.empty_cache()
        line["configs"] = collect_configs(*run_child("--configs-child", local_rank, 420))
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
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--parity-seconds", type=float, default=30.0, help="budget of the oracle bands of the parity block")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--quiet-zone", type=int, default=None, choices=[0, 1, 2],
                    help="override LINON_OPT_QUIET_ZONE (library default: 2 = straight / identity steps outside the mirage's zone of influence)")
    ap.add_argument("--overlay-kernel", type=int, default=None, choices=[0, 1, 2], help="override LINON_OPT_OVERLAY_KERNEL (default 2: TMA tile kernel)")
    ap.add_argument("--peer-stores", type=int, default=1, choices=[0, 1],
                    help="N > 1: 1 (default since round 2: 12.86 vs 16.08 ms per step at N = 2, bit-equal frames) = every rank's march "
                         "stores straight into rank 0's frame buffers (CUDA IPC + NVLink); 0 = one grouped NCCL send/recv to rank 0 "
                         "after the march")
    ap.add_argument("--c5-frames", type=int, default=2, help="frames of the 8192x8192 C5 record (0 = skip)")
    ap.add_argument("--configs-child", action="store_true", help="internal: the child process of the N=1 run")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    if args.configs_child:
        return run_configs_child(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
