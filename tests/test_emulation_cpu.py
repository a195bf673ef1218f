"""The march module's own device code, executed on the host cores (tests/emu), against the oracle.

tests/emu compiles the exact NVRTC source text of a kernel variant with g++ behind a small CUDA shim (one warp per CTA,
lanes as fibers that meet at the warp collectives) - so the logic of linon_march_tiles (tile mapping, phases, votes,
culling budget, counters, stores) is covered without a GPU:
  * strict-math variants must equal the oracle bit for bit (that also pins the emulation itself);
  * the quiet-zone path (LINON_OPT_QUIET_ZONE, csrc/device/mirage_fused.cuh: straight steps where a fused mirage body is
    the identity) must meet the north-star tolerances against the oracle, execute every step, and spend most of a C3-like
    frame in straight steps.
Fast-math variants differ from the device by the MUFU approximation error only (libm here).
"""
import numpy as np
import pytest

from linon_b200 import _capi as capi
from linon_b200 import parity
from linon_b200.main_view import PredefinedFunction
from oracle import np_oracle as npo

from emu import emu

SETTINGS = dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1, central_difference_delta=1, lyapunov_scaling=50.0)
MIRAGES = [PredefinedFunction.MirageSpherical, PredefinedFunction.MiragePlane, PredefinedFunction.MirageSphericalSigmoid,
           PredefinedFunction.MiragePlaneSigmoid]


def _frame_inputs(W, H, orbit_frame=0):
    verts, faces, center = npo.default_scene()[:3]
    cam = npo.camera_uniform(npo.orbit_camera(center, W, H, orbit_frame))
    return verts, faces, cam


def _oracle_frame(oracle, verts, faces, cam, fn, W, H, h_initial, adaptive):
    ref = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=h_initial, adaptive=adaptive)
    _, exponents = oracle.overlay(ref["rgba8"], ref["endpoint"], SETTINGS)
    return ref, parity.high_lyapunov_mask(exponents)


@pytest.mark.parametrize("fn", [PredefinedFunction.MirageSpherical, PredefinedFunction.TranslationX, PredefinedFunction.LorenzAttractor])
@pytest.mark.parametrize("cull", [0, 1])
def test_emulated_strict_kernel_is_bit_exact_with_the_oracle(product_lib, oracle, fn, cull):
    """Fixed default step, strict math: the device code on the host equals the oracle in every output bit and counter."""
    W, H = 40, 27  # ragged: partial tiles on both edges
    verts, faces, cam = _frame_inputs(W, H)
    lib = emu.build(fn.to_code(), capi.RENDER_ADAPTIVE_OFF, capi.MATH_STRICT, cull=cull)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    ref = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=0.1, adaptive=False)
    assert np.array_equal(got["hit_id"], ref["hit_id"])
    assert np.array_equal(got["rgba8"], ref["rgba8"])
    assert np.array_equal(got["endpoint"].view(np.uint32), ref["endpoint"].view(np.uint32))
    assert int(got["counters"][0]) == W * H * 51 == int(ref["accepted"].sum())
    assert int(got["counters"][1]) == int(ref["prehit"].sum())
    if not cull:
        assert int(got["counters"][3]) == int(ref["prehit"].sum()) * 36  # every face on every pre-hit step


def test_emulated_adaptive_strict_kernel_is_bit_exact_with_the_oracle(product_lib, oracle):
    """Adaptive sampling (per-lane loop, votes among the lanes still marching)."""
    W, H = 32, 20
    fn = PredefinedFunction.MirageSpherical
    verts, faces, cam = _frame_inputs(W, H)
    lib = emu.build(fn.to_code(), 0, capi.MATH_STRICT)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    ref = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=0.1, adaptive=True)
    assert np.array_equal(got["hit_id"], ref["hit_id"])
    assert np.array_equal(got["endpoint"].view(np.uint32), ref["endpoint"].view(np.uint32))
    assert int(got["counters"][0]) == int(ref["accepted"].sum()) and int(got["counters"][2]) == int(ref["rejected"].sum())


@pytest.mark.parametrize("adaptive", [False, True])
@pytest.mark.parametrize("kernel", [capi.KERNEL_TILES, capi.KERNEL_PERSISTENT, capi.KERNEL_PAIRS, capi.KERNEL_REFILL])
def test_every_emulated_march_kernel_is_bit_exact_with_the_oracle(product_lib, oracle, kernel, adaptive):
    """The tile kernel, two rays per thread (f32x2, component-wise here), persistent warps in two phases (device queue between
    them) and in one phase (lane refill from a chunked cursor): the same frame, bit for bit, counters included."""
    W, H = 41, 27
    fn = PredefinedFunction.MirageSpherical
    verts, faces, cam = _frame_inputs(W, H)
    lib = emu.build(fn.to_code(), 0 if adaptive else capi.RENDER_ADAPTIVE_OFF, capi.MATH_STRICT, pairs=True)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H, kernel=kernel, grid=5)
    ref = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=0.1, adaptive=adaptive)
    for k in ("hit_id", "rgba8", "endpoint"):
        assert np.array_equal(got[k].view(np.uint8), ref[k].view(np.uint8)), k
    assert int(got["counters"][0]) == int(ref["accepted"].sum())
    assert int(got["counters"][2]) == int(ref["rejected"].sum())


@pytest.mark.parametrize("fn", [PredefinedFunction.MirageSpherical, PredefinedFunction.RoesslerAttractor])
def test_emulated_sample_ray_kernel_is_bit_exact_with_the_oracle(product_lib, oracle, fn):
    """linon_sample_rays (src/main_view.wgsl:281-314): eight rays, 5000 fixed steps, a sample every 50th step."""
    W, H = 64, 48
    verts, faces, cam = _frame_inputs(W, H)
    pos = oracle.sample_positions(cam, SETTINGS, W, H)
    ref = np.asarray(oracle.sample_rays(cam, SETTINGS, int(fn), W, H, pos), np.float32).reshape(800, 8)
    lib = emu.build(fn.to_code(), 0, capi.MATH_STRICT)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H, sample_pos=pos)["ray_samples"]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def _box_with_sphere(n_lat=12, n_lon=24):
    """The start-up scene plus a latitude / longitude sphere inside it (shared edges, degenerate pole triangles)."""
    verts, faces, _ = npo.default_scene()[:3]
    c, r = np.array([0.0, 0.0, 0.0], np.float32), np.float32(0.4)
    pts = []
    for i in range(n_lat + 1):
        th = np.pi * i / n_lat
        for j in range(n_lon):
            ph = 2.0 * np.pi * j / n_lon
            pts.append(c + r * np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)], np.float32))
    tris = []
    for i in range(n_lat):
        for j in range(n_lon):
            a, b = i * n_lon + j, i * n_lon + (j + 1) % n_lon
            tris += [(a, b, a + n_lon), (b, b + n_lon, a + n_lon)]
    v = np.concatenate([np.asarray(verts, np.float32), np.asarray(pts, np.float32)])
    f = np.concatenate([np.asarray(faces, np.uint32), np.asarray(tris, np.uint32) + len(verts)])
    return v, f


@pytest.mark.parametrize("fn,adaptive", [(PredefinedFunction.TranslationX, False), (PredefinedFunction.MirageSpherical, True)])
def test_emulated_hierarchy_traversal_matches_the_oracles_loop_over_every_face(product_lib, oracle, fn, adaptive):
    """LINON_BVH (meshes beyond the shared-memory stage, src/main.rs:139-157): faces in hierarchy order + a bounding-volume
    hierarchy built by mesh_tables.hpp, exact traversal in intersect_mesh.  612 faces here, forced through that path: hit ids
    (lowest index wins a shared edge), endpoints and step counters equal the oracle's loop over all faces bit for bit."""
    W, H = 48, 36
    verts, faces = _box_with_sphere()
    _, _, cam = _frame_inputs(W, H)
    lib = emu.build(fn.to_code(), 0 if adaptive else capi.RENDER_ADAPTIVE_OFF, capi.MATH_STRICT, bvh=True, n_faces=len(faces))
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    ref = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=0.1, adaptive=adaptive)
    assert np.array_equal(got["hit_id"], ref["hit_id"])
    assert np.array_equal(got["rgba8"], ref["rgba8"])
    assert np.array_equal(got["endpoint"].view(np.uint32), ref["endpoint"].view(np.uint32))
    assert (got["hit_id"] >= 36).mean() > 0.03  # the sphere is really hit
    assert int(got["counters"][0]) == int(ref["accepted"].sum()) and int(got["counters"][1]) == int(ref["prehit"].sum())
    assert int(got["counters"][3]) < 0.1 * int(ref["prehit"].sum()) * len(faces)  # the traversal prunes


PREDEFINED = [f for f in PredefinedFunction if f != PredefinedFunction.Custom]


@pytest.mark.parametrize("fn", PREDEFINED)
def test_emulated_fast_build_meets_the_north_star(product_lib, oracle, fn):
    """The production build (fast math; hand-fused whole-step restatement for the four mirage bodies, generic translation
    for the rest) against the oracle, adaptive off: the algebra of the restatements is checked here with libm's special
    functions, the MUFU approximations on top of it by tests/test_gpu_parity.py on the device."""
    W, H = 64, 48
    verts, faces, cam = _frame_inputs(W, H)
    ref, excl = _oracle_frame(oracle, verts, faces, cam, fn, W, H, 0.1, False)
    lib = emu.build(fn.to_code(), capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    m = parity.compare_frames(got, ref, exclude=excl)
    assert parity.passes_north_star(m), m
    assert int(got["counters"][0]) == W * H * 51


@pytest.mark.parametrize("fn", MIRAGES)
def test_quiet_zone_default_step_frames_meet_the_north_star(product_lib, oracle, fn):
    """Adaptive off, h = 0.1 (51 steps): quiet-zone straight steps against the oracle, all four predefined mirage bodies."""
    W, H = 64, 48
    verts, faces, cam = _frame_inputs(W, H)
    ref, excl = _oracle_frame(oracle, verts, faces, cam, fn, W, H, 0.1, False)
    lib = emu.build(fn.to_code(), capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST, quiet_zone=1)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    m = parity.compare_frames(got, ref, exclude=excl)
    assert parity.passes_north_star(m), m
    c = got["counters"]
    assert int(c[0]) == W * H * 51              # every step is still executed and counted
    assert int(c[15]) > 0.5 * int(c[0]), c      # ... most of them as straight steps


@pytest.mark.parametrize("fn,orbit_frame", [(PredefinedFunction.MirageSphericalSigmoid, 0), (PredefinedFunction.MirageSphericalSigmoid, 24),
                                            (PredefinedFunction.MirageSpherical, 0)])
def test_quiet_zone_enhance_frames_meet_the_north_star(product_lib, oracle, fn, orbit_frame):
    """C3 at reduced resolution (h = 0.001, 5001 steps per ray) with the quiet zone on: north star against the oracle, and no
    worse than the same kernel without it."""
    W, H = 48, 36
    verts, faces, cam = _frame_inputs(W, H, orbit_frame)
    ref, excl = _oracle_frame(oracle, verts, faces, cam, fn, W, H, 0.001, False)
    frames = {}
    for quiet in (0, 1):
        lib = emu.build(fn.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=quiet)
        frames[quiet] = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    m_off = parity.compare_frames(frames[0], ref, exclude=excl)
    m_on = parity.compare_frames(frames[1], ref, exclude=excl)
    assert parity.passes_north_star(m_on), m_on
    assert m_on["hit_group_agree"] >= m_off["hit_group_agree"] - 1e-3 and m_on["endpoint_max_rel"] <= max(2.0 * m_off["endpoint_max_rel"], 5e-4), (m_on, m_off)
    c_on, c_off = frames[1]["counters"], frames[0]["counters"]
    assert int(c_on[0]) == int(c_off[0]) == W * H * 5001
    assert int(c_off[15]) == 0
    assert int(c_on[15]) > 0.9 * int(c_on[0]), c_on      # > 90 % of the steps of this frame are outside the zone
    # on against off: the two builds differ only by rounding noise outside the zone
    m = parity.compare_frames(frames[1], frames[0], exclude=excl)
    assert m["hit_group_agree"] >= 0.999 and m["endpoint_within_1e3"] >= 0.999, m


@pytest.mark.parametrize("fn,flags,size,orbit_frame", [
    (PredefinedFunction.MirageSphericalSigmoid, capi.RENDER_ENHANCE, (40, 28), 0),
    (PredefinedFunction.MiragePlane, capi.RENDER_ENHANCE, (40, 28), 13),
    (PredefinedFunction.MirageSpherical, capi.RENDER_ADAPTIVE_OFF, (96, 72), 40),
    (PredefinedFunction.MiragePlaneSigmoid, capi.RENDER_ADAPTIVE_OFF, (96, 72), 0),
    # Enhance frames from other points of the orbit (round 2: window resolve and collapsed runs act on these)
    (PredefinedFunction.MirageSpherical, capi.RENDER_ENHANCE, (56, 36), 27),
    (PredefinedFunction.MiragePlaneSigmoid, capi.RENDER_ENHANCE, (48, 32), 50)])
def test_quiet_zone_culling_never_changes_a_result(product_lib, fn, flags, size, orbit_frame):
    """Straight runs skip culling checks on the strength of two bounds (the safe-distance budget and the directional bound
    of straight_clear_steps).  With LINON_CULL = 0 every pre-hit step runs the loop over all 36 faces instead: the frames must
    be identical bit for bit, i.e. no skipped check could have found a face."""
    W, H = size
    verts, faces, cam = _frame_inputs(W, H, orbit_frame)
    frames = []
    for cull in (1, 0):
        lib = emu.build(fn.to_code(), flags, capi.MATH_FAST, quiet_zone=1, cull=cull)
        frames.append(emu.march(lib, verts, faces, cam, SETTINGS, W, H))
    a, b = frames
    for k in ("hit_id", "rgba8", "endpoint"):
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k
    assert (a["hit_id"] >= 0).mean() > 0.3
    assert int(a["counters"][15]) == int(b["counters"][15]) > 0          # the same steps are straight steps
    assert int(a["counters"][4]) < 0.25 * int(b["counters"][4])           # ... and most checks are gone


@pytest.mark.parametrize("fn,flags,size,orbit_frame", [
    (PredefinedFunction.MirageSphericalSigmoid, capi.RENDER_ENHANCE, (40, 28), 0),
    (PredefinedFunction.MiragePlaneSigmoid, capi.RENDER_ENHANCE, (48, 32), 50),
    (PredefinedFunction.MirageSpherical, capi.RENDER_ADAPTIVE_OFF, (96, 72), 40)])
def test_results_do_not_depend_on_the_last_bits_of_the_approximate_division(product_lib, fn, flags, size, orbit_frame):
    """The one intrinsic the emulation cannot reproduce: __fdividef is a MUFU.RCP and a multiply on the device (up to 2 ulp off),
    an exact division on the host.  It is used where a result is made exact afterwards or carries a margin - the number of steps
    of a collapsed run per binade segment (corrected by an exact FMA test), the length of a quiet run and the steps a
    safe-distance budget covers (rounded down behind a relative margin of 1e-4).  With every quotient moved by a pseudo-random
    -4 ... +4 ulps (tests/emu/cuda_shim.hpp) the frames are still the frames of the kernel without culling, bit for bit, and
    advance_coord is still the chain of single steps."""
    import ctypes as C
    W, H = size
    verts, faces, cam = _frame_inputs(W, H, orbit_frame)
    noisy = emu.build(fn.to_code(), flags, capi.MATH_FAST, quiet_zone=1, cull=1, extra_defines=("-DLINON_EMU_FDIVIDEF_ULPS=4u",))
    plain = emu.build(fn.to_code(), flags, capi.MATH_FAST, quiet_zone=1, cull=0)
    a, b = emu.march(noisy, verts, faces, cam, SETTINGS, W, H), emu.march(plain, verts, faces, cam, SETTINGS, W, H)
    for k in ("hit_id", "rgba8", "endpoint"):
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k
    assert int(a["counters"][15]) == int(b["counters"][15]) > 0
    if flags == capi.RENDER_ENHANCE:
        rng = np.random.default_rng(5)
        m = 300_000
        p = (rng.uniform(-3, 3, m) * rng.choice([1e-3, 1.0, 1.0, 30.0], m)).astype(np.float32)
        d = rng.uniform(-1.5, 1.5, m).astype(np.float32)
        n = rng.integers(0, 5002, m).astype(np.int32)
        x, y = np.zeros_like(p), np.zeros_like(p)
        noisy.emu_advance_coord(C.c_int(m), C.c_void_p(p.ctypes.data), C.c_void_p(d.ctypes.data), C.c_float(0.001),
                                C.c_void_p(n.ctypes.data), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data))
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("fn", MIRAGES)
def test_quiet_zone_adaptive_frames_meet_the_north_star_on_hit_and_colour(product_lib, oracle, fn):
    """Option level 2: adaptive frames (C1 / C2 as the application renders them).  Graded like the GPU test of the adaptive
    path: hit id and colour, pixels the oracle itself cannot decide (fp32 vs fp64 vs a one-ulp change of field_weight)
    excluded and counted - adaptive step sequences are rounding-noise driven."""
    W, H = 96, 72
    verts, faces, cam = _frame_inputs(W, H)
    ref, excl = _oracle_frame(oracle, verts, faces, cam, fn, W, H, 0.1, True)
    alt64 = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=0.1, adaptive=True, use_fp64=True)
    st_ulp = dict(SETTINGS, field_weight=float(np.nextafter(np.float32(1.0), np.float32(0))))
    alt_ulp = oracle.march(verts, faces, cam, st_ulp, int(fn), W, H, h_initial=0.1, adaptive=True)
    excl = excl | parity.disagreement_mask(alt64, ref) | parity.disagreement_mask(alt_ulp, ref)
    lib = emu.build(fn.to_code(), 0, capi.MATH_FAST, quiet_zone=2)
    got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
    m = parity.compare_frames(got, ref, exclude=excl)
    assert m["hit_group_agree"] >= 0.999 and m["colour_within_1"] >= 0.99, m
    c = got["counters"]
    # a good part of the attempts never evaluate the field (30 % for the spherical sigmoid since the step control has the
    # reference's noise floor: rays no longer crawl at h = 0.001 through the calm tail of the profile, see rk4_attempt)
    assert int(c[15]) > 0.25 * (int(c[0]) + int(c[2])), c
    assert int(c[0]) >= W * H * 51


@pytest.mark.parametrize("field_weight,orbit_frame", [(0.5, 7), (0.05, 50)])
def test_quiet_zone_with_partial_field_weight(product_lib, oracle, field_weight, orbit_frame):
    """field_weight != 1: the direction blend of src/main_view.wgsl:256 stays in the regular steps, straight runs keep the
    direction (a blend of a vector with itself).  With a partial weight the mirage frames are ill-conditioned in the
    reference itself - FMA contraction alone moves 4 % of the endpoints by more than 1e-3 (grazing rays at the
    total-reflection threshold) - so the option is graded against the same build without it, and both on hit id and colour."""
    W, H = 64, 48
    fn = PredefinedFunction.MirageSphericalSigmoid
    st = dict(SETTINGS, field_weight=field_weight)
    verts, faces, cam = _frame_inputs(W, H, orbit_frame)
    ref = oracle.march(verts, faces, cam, st, int(fn), W, H, h_initial=0.1, adaptive=False)
    _, exponents = oracle.overlay(ref["rgba8"], ref["endpoint"], st)
    excl = parity.high_lyapunov_mask(exponents)
    m = {}
    for quiet in (0, 1):
        lib = emu.build(fn.to_code(), capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST, quiet_zone=quiet)
        got = emu.march(lib, verts, faces, cam, st, W, H)
        m[quiet] = parity.compare_frames(got, ref, exclude=excl)
        assert (int(got["counters"][15]) > 0) == bool(quiet)
    assert m[1]["hit_group_agree"] >= 0.999 and m[1]["colour_within_1"] >= 0.99, m[1]
    for k in ("hit_group_agree", "colour_within_1", "endpoint_within_1e3"):
        assert m[1][k] >= m[0][k] - 2e-3, (k, m)


def _quiet_plan(lib, p, d, h, verts=None, faces=None, n_steps_left=5001):
    import ctypes as C
    p = np.ascontiguousarray(p, np.float32)
    d = np.ascontiguousarray(d, np.float32)
    n = p.shape[0]
    nq = np.zeros(n, np.int32)
    nc = np.zeros(n, np.int32) if verts is not None else None
    v = np.ascontiguousarray(verts, np.float32) if verts is not None else None
    f = np.ascontiguousarray(faces, np.uint32) if faces is not None else None
    lib.emu_quiet_plan(C.c_int(n), C.c_void_p(p.ctypes.data), C.c_void_p(d.ctypes.data), C.c_float(h),
                       C.c_void_p(v.ctypes.data if v is not None else 0), C.c_int(0 if v is None else v.shape[0]),
                       C.c_void_p(f.ctypes.data if f is not None else 0), C.c_int(0 if f is None else f.shape[0]),
                       C.c_int(n_steps_left), C.c_void_p(nq.ctypes.data), C.c_void_p(nc.ctypes.data if nc is not None else 0))
    return nq, nc


@pytest.mark.parametrize("fn,centre_or_height,zone", [(PredefinedFunction.MirageSphericalSigmoid, (-0.5, 0.5, -0.5), 1.8866 * 0.25),
                                                       (PredefinedFunction.MirageSpherical, (-0.5, 0.5, -0.5), 0.25),
                                                       (PredefinedFunction.MiragePlaneSigmoid, 0.1, 1.8866 * 0.01),
                                                       (PredefinedFunction.MiragePlane, 0.1, 0.01)])
def test_quiet_steps_never_reach_the_zone_of_influence(product_lib, fn, centre_or_height, zone):
    """quiet_steps (mirage_fused.cuh) on 20 000 random rays: every stage point of the promised steps (the straight line from
    the start to n h dir) stays outside the region where the reference's own evaluation can differ from the identity -
    dist < max_dist for the linear profiles, dist < 1.8866 max_dist (fp32 sigmoid saturation) for the sigmoids - and the
    promise is not needlessly short."""
    rng = np.random.default_rng(7)
    n, h = 20000, np.float32(0.001)
    p = rng.uniform(-1.3, 1.3, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= rng.uniform(0.5, 1.5, (n, 1)).astype(np.float32)       # the direction is not normalised once field_weight != 1
    lib = emu.build(fn.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=1)
    nq, _ = _quiet_plan(lib, p, d, h)
    assert (nq >= 0).all() and (nq > 0).mean() > 0.5
    p64, d64 = p.astype(np.float64), d.astype(np.float64) * float(h)
    spherical = not np.isscalar(centre_or_height)

    def dist_min(steps):  # smallest distance to the zone's centre / plane along p + s d, s in [0, steps]
        if spherical:
            c = p64 - np.asarray(centre_or_height)
            s_star = np.clip(-(c * d64).sum(1) / (d64 * d64).sum(1), 0.0, steps)
            return np.linalg.norm(c + s_star[:, None] * d64, axis=1)
        y0, y1 = p64[:, 1] - centre_or_height, p64[:, 1] + steps * d64[:, 1] - centre_or_height
        return np.where(y0 * y1 <= 0.0, 0.0, np.minimum(np.abs(y0), np.abs(y1)))

    capped = np.minimum(nq, 1 << 20).astype(np.float64)
    assert (dist_min(capped)[nq > 0] >= zone).all()
    # not needlessly short: a finite promise ends within a few steps (and 0.1 % of its length) of the radius the kernel uses
    finite = (nq > 0) & (nq < (1 << 20))
    radius = {1: 1.001 * 0.25, 3: 1.92 * 0.25, 2: 1.001 * 0.01, 4: 1.92 * 0.01}[int(fn)]
    assert (dist_min(capped * 1.002 + 4.0)[finite] <= radius * 1.0001).all()


def test_straight_clear_steps_never_skip_a_face(product_lib, oracle):
    """straight_clear_steps (march_kernels.cuh) on random straight rays inside the start-up scene: on none of the promised steps
    can the reference's own hit test (hit_triangle over all 36 faces, t in (eps, max_dist), fp64 here) find a face."""
    rng = np.random.default_rng(11)
    verts, faces, _ = npo.default_scene()[:3]
    verts, faces = np.asarray(verts, np.float64), np.asarray(faces, np.int64)
    n, h = 4000, np.float32(0.001)
    p = rng.uniform(-0.98, 0.98, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    lib = emu.build(PredefinedFunction.MirageSphericalSigmoid.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=1)
    _, nc = _quiet_plan(lib, p, d, h, verts.astype(np.float32), faces.astype(np.uint32))
    assert (nc >= 0).all() and (nc > 10).mean() > 0.5
    # the promised stretch p .. p + nc h d as ONE segment against every triangle (Moeller-Trumbore, src/main_view.wgsl:150-175)
    v0, e1, e2 = verts[faces[:, 0]], verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]]
    o = p.astype(np.float64)[:, None, :]
    length = np.minimum(nc, 20000).astype(np.float64) * float(h)
    dn = d.astype(np.float64)[:, None, :]
    hh = np.cross(dn, e2[None])
    a = (e1[None] * hh).sum(-1)
    ok = np.abs(a) > 1e-12
    f = np.where(ok, 1.0 / np.where(ok, a, 1.0), 0.0)
    s = o - v0[None]
    u = f * (s * hh).sum(-1)
    q = np.cross(s, e1[None])
    v = f * (dn * q).sum(-1)
    t = f * (e2[None] * q).sum(-1)
    hit = ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 1e-7) & (t < length[:, None])
    assert not hit.any(), int(hit.any(axis=1).sum())
    # ... and the promise is not needlessly short: a few steps later a face is within one step's reach, unless none lies ahead
    t_first = np.where(ok & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 1e-7), t, np.inf).min(axis=1)
    ahead = np.isfinite(t_first) & (nc < 20000)
    assert np.median((t_first[ahead] / float(h) - nc[ahead])) < 40.0


def test_quiet_zone_radius_option(product_lib, oracle):
    """LINON_OPT_QUIET_RADIUS: at 1.5 max_dist the sigmoid's real-number deflection over the excluded shell is 1e-9 rad (below
    fp32 resolution, far below the reference's own rounding noise there): same agreement with the oracle, half the evaluated
    steps.  At 1.0 max_dist the signal is real - that radius is refused by the option and would fail here."""
    W, H = 48, 36
    fn = PredefinedFunction.MirageSphericalSigmoid
    verts, faces, cam = _frame_inputs(W, H)
    ref, excl = _oracle_frame(oracle, verts, faces, cam, fn, W, H, 0.001, False)
    full = {}
    for radius in (0, 1500):
        lib = emu.build(fn.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=1, quiet_radius=radius)
        got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
        m = parity.compare_frames(got, ref, exclude=excl)
        assert parity.passes_north_star(m), (radius, m)
        assert m["endpoint_max_rel"] < 1e-3, (radius, m)
        full[radius] = int(got["counters"][0]) - int(got["counters"][15])
    assert full[1500] < 0.7 * full[0], full


def test_quiet_zone_is_closer_to_the_fp64_reference_run(product_lib, oracle):
    """Outside the zone the reference's fp32 result is `v + rounding noise`; a straight step has none.  Against the oracle's
    fp64 evaluation of the same WGSL the quiet-zone frame must not be farther away than the plain fast build."""
    W, H = 48, 36
    fn = PredefinedFunction.MirageSphericalSigmoid
    verts, faces, cam = _frame_inputs(W, H)
    ref64 = oracle.march(verts, faces, cam, SETTINGS, int(fn), W, H, h_initial=0.001, adaptive=False, use_fp64=True)
    _, exponents = oracle.overlay(ref64["rgba8"], ref64["endpoint"], SETTINGS)
    excl = parity.high_lyapunov_mask(exponents)
    err = {}
    for quiet in (0, 1):
        lib = emu.build(fn.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=quiet)
        got = emu.march(lib, verts, faces, cam, SETTINGS, W, H)
        use = ~excl & (parity.canonical_hits(got["hit_id"]) == parity.canonical_hits(ref64["hit_id"]))
        d = np.linalg.norm(got["endpoint"][use][:, :3].astype(np.float64) - ref64["endpoint"][use][:, :3], axis=-1)
        err[quiet] = float(np.median(d))
    assert err[1] <= 1.5 * err[0] + 1e-7, err


def test_quiet_zone_needs_a_fused_mirage_body(product_lib):
    """Generic translations, strict / fused math and adaptive variants never define the zone: the option compiles to the
    same module as without it."""
    import os
    import tempfile
    from linon_b200.main_view import compile_check

    def module_text(body, flags, math, quiet):
        fd, path = tempfile.mkstemp(suffix=".cu")
        os.close(fd)
        old = {k: os.environ.get(k) for k in ("LINON_DUMP_SOURCE", "LINON_CHECK_QUIET_ZONE")}
        os.environ["LINON_DUMP_SOURCE"] = path
        os.environ["LINON_CHECK_QUIET_ZONE"] = str(quiet)
        try:
            compile_check(body, flags=flags, math_mode=math)
            return open(path).read()
        finally:
            os.unlink(path)
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)

    sig = PredefinedFunction.MirageSphericalSigmoid.to_code()
    assert "-DLINON_QUIET_ZONE=1" in module_text(sig, capi.RENDER_ENHANCE, capi.MATH_FAST, 1).splitlines()[0]
    assert "-DLINON_QUIET_ZONE=0" in module_text(sig, capi.RENDER_ENHANCE, capi.MATH_FAST, 0).splitlines()[0]
    assert "-DLINON_QUIET_ZONE=0" in module_text(sig, capi.RENDER_ENHANCE, capi.MATH_STRICT, 1).splitlines()[0]
    assert "-DLINON_QUIET_ZONE=0" in module_text(sig, 0, capi.MATH_FAST, 1).splitlines()[0]  # adaptive sampling on
    # a body outside the mirage family (here: an expression where the family has a literal; csrc/host/mirage_match.hpp) runs
    # the generic translation, which has no quiet_steps: the macro is switched off in the module
    generic = module_text(sig.replace("let t_env = 15.0;", "let t_env = 15.0 + 0.0;"), capi.RENDER_ENHANCE, capi.MATH_FAST, 1)
    assert "LINON_HAS_QUIET_ZONE" not in generic.split("march_kernels.cuh", 1)[0].replace("defined(LINON_HAS_QUIET_ZONE)", "")


@pytest.mark.parametrize("fn", [PredefinedFunction.MirageSpherical, PredefinedFunction.MiragePlane,
                                PredefinedFunction.MirageSphericalSigmoid, PredefinedFunction.MiragePlaneSigmoid])
def test_emulated_sample_rays_with_the_quiet_zone(product_lib, oracle, fn):
    """linon_sample_rays with LINON_OPT_QUIET_ZONE: the eight rays take straight steps outside the body's zone of influence.
    Same 800 samples as the oracle within the spread of the production build, and no farther from the oracle than twice
    what the same build without the option is (max over the eight rays)."""
    W, H = 64, 48
    verts, faces, cam = _frame_inputs(W, H)
    for mouse in ((0.5, 0.6), (0.36, 0.55), (0.05, 0.95)):  # start-up click, the mirage rim, a corner far from the body
        st = dict(SETTINGS, mouse_x=mouse[0], mouse_y=mouse[1])
        pos = oracle.sample_positions(cam, st, W, H)
        ref = np.asarray(oracle.sample_rays(cam, st, int(fn), W, H, pos), np.float32).reshape(800, 8)
        plain = emu.march(emu.build(fn.to_code(), capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST), verts, faces, cam, st, W, H,
                          sample_pos=pos)["ray_samples"]
        quiet = emu.march(emu.build(fn.to_code(), capi.RENDER_ADAPTIVE_OFF, capi.MATH_FAST, quiet_zone=1), verts, faces, cam, st,
                          W, H, sample_pos=pos)["ray_samples"]
        # rays through the body amplify rounding differences (the plain production build is up to 3e-3 from the oracle here)
        assert np.allclose(quiet, ref, rtol=0, atol=5e-3)
        assert np.array_equal(quiet[:, 3:], ref[:, 3:])  # colours and the 0.5 alpha
        err_q = np.abs(quiet[:, :3].astype(np.float64) - ref[:, :3]).max()
        err_p = np.abs(plain[:, :3].astype(np.float64) - ref[:, :3]).max()
        assert err_q <= max(2.0 * err_p, 2e-5), (err_q, err_p)


def test_collapsed_runs_equal_the_step_by_step_chain(product_lib):
    """advance_coord (march_kernels.cuh, LINON_COLLAPSE_RUNS): n straight steps of a coordinate, p <- fma(d, h, p), taken binade by
    binade - one real step, then k steps as q + k delta while every step starts and ends inside the binade.  Must equal the chain
    bit for bit: random coordinates of a frame, zero crossings, starts on and next to binade boundaries, increments that are exact
    half-ulps (ties), increments below the ulp (the coordinate does not move), tiny / huge / non-finite values."""
    import ctypes as C
    lib = emu.build(PredefinedFunction.MirageSphericalSigmoid.to_code(), capi.RENDER_ENHANCE, capi.MATH_FAST, quiet_zone=1)
    rng = np.random.default_rng(11)
    h = np.float32(0.001)
    P, D, N = [], [], []

    def add(p, d, n):
        P.append(np.asarray(p, np.float32).ravel()); D.append(np.asarray(d, np.float32).ravel()); N.append(np.asarray(n, np.int32).ravel())

    m = 400_000
    add(rng.uniform(-3, 3, m), rng.uniform(-1.5, 1.5, m), rng.integers(0, 5002, m))                       # a frame's coordinates
    add(rng.uniform(-2e-3, 2e-3, m // 4), rng.uniform(-1.5, 1.5, m // 4), rng.integers(0, 3000, m // 4))   # around zero
    pw = np.ldexp(1.0, rng.integers(-12, 3, m // 4)).astype(np.float32) * rng.choice([-1.0, 1.0], m // 4).astype(np.float32)
    off = rng.integers(-3, 4, m // 4).astype(np.int32)
    near = (pw.view(np.int32) + off).view(np.float32)                                                     # on / next to a binade boundary
    add(near, rng.uniform(-1.5, 1.5, m // 4), rng.integers(0, 2000, m // 4))
    # ties: d h exactly an odd multiple of half an ulp of p (h = 2^-10 here would be needed for exact products: use d = c / h
    # rounded - not exact - plus true ties with a power-of-two step in a second call below)
    add(rng.uniform(0.5, 1.0, m // 8), rng.uniform(-1e-4, 1e-4, m // 8), rng.integers(0, 5002, m // 8))    # |a| < ulp: stuck or crawling
    add([0.0, -0.0, 1e-30, -1e-30, 3e38, -3e38, np.inf, -np.inf, np.nan, 1.0, 1.0], [1, 1, 1, -1, 1, 1, 1, 1, 1, 0, np.nan],
        [100, 100, 5000, 5000, 100, 100, 10, 10, 10, 5000, 7])
    p, d, n = np.concatenate(P), np.concatenate(D), np.concatenate(N)
    assert p.size == d.size == n.size
    for hh, pp, dd, nn in ((h, p, d, n),
                           # exact ties: h = 2^-10, d a multiple of 2^-14 -> d h is a multiple of 2^-24 = half an ulp in [0.5, 1)
                           (np.float32(2.0 ** -10), rng.uniform(0.5, 1.0, 200_000).astype(np.float32),
                            (rng.integers(-4000, 4000, 200_000) * 2.0 ** -14).astype(np.float32), rng.integers(0, 3000, 200_000).astype(np.int32))):
        pp, dd, nn = np.ascontiguousarray(pp, np.float32), np.ascontiguousarray(dd, np.float32), np.ascontiguousarray(nn, np.int32)
        a, b = np.zeros_like(pp), np.zeros_like(pp)
        lib.emu_advance_coord(C.c_int(pp.size), C.c_void_p(pp.ctypes.data), C.c_void_p(dd.ctypes.data), C.c_float(float(hh)),
                              C.c_void_p(nn.ctypes.data), C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data))
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        bad = np.nonzero(~same)[0]
        assert bad.size == 0, [(float(pp[i]), float(dd[i]), int(nn[i]), float(a[i]), float(b[i])) for i in bad[:5]]
