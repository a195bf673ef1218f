"""LINON_OPT_QUIET_ZONE on the GPU: straight steps where a fused mirage body is the identity (default since round 2).

The option was written in round 1 without GPU time (validated on the host emulation of the exact NVRTC source,
tests/test_emulation_cpu.py), first ran on hardware in the driver's round-1 GPU tests (all green) and is the shipped path
since round 2: level 2 (fixed-step and adaptive frames) is the library default.  The same
holds for the two other paths tested at the end of the file: peer stores (linon_redirect_output) and the overlay kernels
(LINON_OPT_OVERLAY_KERNEL: the TMA tile kernel is the default, the direct kernel its fallback; all validated bit for bit
against the oracle on the host in tests/test_overlay_emulation_cpu.py).  Full-size grading against the oracle: tests/test_gpu_fullsize.py and bench.py's
`parity` block.
"""
import numpy as np
import pytest

from linon_b200 import _capi as capi
from linon_b200 import parity
from linon_b200.main_view import MainView, PredefinedFunction, make_settings
from test_gpu_parity import oracle_frame, render

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

MIRAGES = [PredefinedFunction.MirageSpherical, PredefinedFunction.MiragePlane, PredefinedFunction.MirageSphericalSigmoid,
           PredefinedFunction.MiragePlaneSigmoid]


@pytest.fixture()
def view(product_lib, scene):
    verts, faces, center = scene
    mv = MainView(verts, faces, center, 160, 120)
    mv.set_option(capi.OPT_QUIET_ZONE, 1)
    yield mv
    mv.close()


def quiet_steps(mv):
    return int(mv.debug_profile()[15])


@pytest.mark.parametrize("fn", MIRAGES)
def test_quiet_zone_fixed_step_frames_meet_the_north_star(view, oracle, scene, fn):
    """Adaptive off (51 steps): hit id, colour and endpoint against the oracle; every step is executed and counted."""
    ref, excl = oracle_frame(oracle, scene, view, fn, adaptive=False)
    got = render(view, fn, flags=capi.RENDER_ADAPTIVE_OFF)
    m = parity.compare_frames(got, ref, exclude=excl)
    assert parity.passes_north_star(m, check_endpoint=True), m
    assert view.counters().accepted_steps == 160 * 120 * 51
    assert quiet_steps(view) > 0.5 * 160 * 120 * 51


@pytest.mark.parametrize("fn", MIRAGES)
@pytest.mark.parametrize("flags", [capi.RENDER_ADAPTIVE_OFF, capi.RENDER_ENHANCE])
def test_quiet_zone_culling_never_changes_a_result(view, fn, flags):
    """Straight runs skip culling checks; with culling off every pre-hit step tests all 36 faces: bit-identical frames."""
    a = render(view, fn, flags=flags, cull=1)
    checks_on = view.counters().culling_checks
    q_on = quiet_steps(view)
    hash_on = view.march_cubin_hash()
    b = render(view, fn, flags=flags, cull=0)
    for k in a:
        same = np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8))
        if not same:  # say what differs: which binaries ran, how many pixels, by how much
            px = (a[k].reshape(a[k].shape[0], a[k].shape[1], -1) != b[k].reshape(b[k].shape[0], b[k].shape[1], -1)).any(axis=-1)
            ys, xs = np.nonzero(px)
            detail = dict(buffer=k, cubin_cull1=hash_on, cubin_cull0=view.march_cubin_hash(), pixels=int(px.sum()),
                          first=[(int(y), int(x), a[k][y, x].tolist(), b[k][y, x].tolist()) for y, x in list(zip(ys, xs))[:4]],
                          quiet_steps=(q_on, quiet_steps(view)))
            assert same, detail
    assert q_on == quiet_steps(view) > 0
    assert checks_on < view.counters().culling_checks


@pytest.mark.parametrize("fn", MIRAGES)
@pytest.mark.parametrize("size", [(160, 120), (333, 197)])
def test_collapsed_runs_equal_the_step_by_step_frames(product_lib, scene, fn, size):
    """LINON_OPT_COLLAPSE_RUNS (default 1): quiet runs of 160 steps or more are taken binade by binade (advance_coord) instead of one
    FFMA per coordinate and step.  Same bits in every buffer and the same step counters as with the option off, on the device
    (the host emulation pins the function itself against the chain on adversarial inputs, tests/test_emulation_cpu.py)."""
    verts, faces, center = scene
    W, H = size
    with MainView(verts, faces, center, W, H) as mv:
        frames = {}
        for collapse in (1, 0):
            mv.set_option(capi.OPT_COLLAPSE_RUNS, collapse)
            frames[collapse] = render(mv, fn, flags=capi.RENDER_ENHANCE)
            c = mv.counters()
            frames[collapse]["counters"] = (int(c.accepted_steps), int(c.prehit_steps), int(c.face_tests), int(c.culling_checks), quiet_steps(mv))
            frames[collapse]["hash"] = mv.march_cubin_hash()
        assert frames[1]["hash"] != frames[0]["hash"]  # two different binaries really ran
        for k in ("hit_id", "rgba8", "endpoint"):
            assert np.array_equal(frames[1][k].view(np.uint8), frames[0][k].view(np.uint8)), k
        assert frames[1]["counters"] == frames[0]["counters"]
        assert frames[1]["counters"][4] > 0.5 * W * H * 5001  # most steps are straight steps (plane mirages: 82 %, spherical: 97 %)


def test_quiet_zone_enhance_frame_meets_the_north_star(product_lib, oracle, scene):
    """C3 at reduced resolution (5001 steps per ray): against the oracle, and against the same kernel without the option."""
    verts, faces, center = scene
    with MainView(verts, faces, center, 64, 48) as mv:
        fn = PredefinedFunction.MirageSphericalSigmoid
        ref, excl = oracle_frame(oracle, scene, mv, fn, h_initial=0.001, adaptive=True)
        frames = {}
        for quiet in (0, 1):
            mv.set_option(capi.OPT_QUIET_ZONE, quiet)
            frames[quiet] = render(mv, fn, flags=capi.RENDER_ENHANCE)
            c = mv.counters()
            assert c.accepted_steps == 64 * 48 * 5001 and c.rejected_steps == 0
            q = quiet_steps(mv)
            assert (q > 0.9 * c.accepted_steps) if quiet else (q == 0), q
        m_on = parity.compare_frames(frames[1], ref, exclude=excl)
        assert parity.passes_north_star(m_on), m_on
        m = parity.compare_frames(frames[1], frames[0], exclude=excl)
        assert m["hit_group_agree"] >= 0.999 and m["endpoint_within_1e3"] >= 0.999 and m["colour_within_1"] >= 0.999, m


@pytest.mark.parametrize("fn", MIRAGES)
def test_quiet_zone_adaptive_frames_meet_the_north_star_on_hit_and_colour(view, oracle, scene, fn):
    """Option level 2 (adaptive frames as well), graded like test_adaptive_frames_meet_the_north_star_on_hit_and_colour."""
    view.set_option(capi.OPT_QUIET_ZONE, 2)
    ref, excl = oracle_frame(oracle, scene, view, fn, adaptive=True)
    got = render(view, fn, flags=0)
    m = parity.compare_frames(got, ref, exclude=excl)
    assert m["hit_group_agree"] >= 0.999 and m["colour_within_1"] >= 0.99, m
    assert quiet_steps(view) > 0


def test_quiet_zone_full_size_c3(product_lib, scene):
    """The C3 frame itself: step counts, share of straight steps, agreement with the default path, and the time of both."""
    verts, faces, center = scene
    W, H = 3840, 2160
    with MainView(verts, faces, center, W, H) as mv:
        fn = PredefinedFunction.MirageSphericalSigmoid
        mv.update_settings(make_settings(overlay_mode=1))
        out, ms = {}, {}
        for quiet in (0, 1):
            mv.set_option(capi.OPT_QUIET_ZONE, quiet)
            for _ in range(2):
                out[quiet] = render(mv, fn, flags=capi.RENDER_ENHANCE, settings=make_settings(overlay_mode=1))
            ms[quiet] = mv.timings().march_ms
            assert mv.counters().accepted_steps == W * H * 5001
        q = quiet_steps(mv)
        excl = mv.exponents() >= 0.8
        m = parity.compare_frames(out[1], out[0], exclude=excl)
        print(f"C3 march: default {ms[0]:.1f} ms, quiet zone {ms[1]:.1f} ms, straight steps {q / (W * H * 5001):.4f}, {m}")
        assert q > 0.9 * W * H * 5001
        assert m["hit_group_agree"] >= 0.999 and m["endpoint_within_1e3"] >= 0.999 and m["colour_within_1"] >= 0.999, m
        assert ms[1] < ms[0]


@pytest.mark.parametrize("flags", [capi.RENDER_ADAPTIVE_OFF, 0])
def test_peer_stores_assemble_the_frame_in_the_root_views_buffers(product_lib, scene, flags):
    """linon_redirect_output: the shards of a frame store their rows straight into the root shard's buffers (what the ranks
    of a multi-GPU job do through CUDA IPC mappings, bench.py --peer-stores): no gather, and the root's buffers equal the
    one-view frame bit for bit.  Same process and device here; the IPC calls themselves need two processes."""
    verts, faces, center = scene
    W, H, world = 96, 80, 3
    fn = PredefinedFunction.MirageSphericalSigmoid
    bufs = (capi.BUF_RAYCAST_RGBA8, capi.BUF_ENDPOINT_F32X4, capi.BUF_HIT_ID_I32)
    with MainView(verts, faces, center, W, H) as whole:
        ref = render(whole, fn, flags=flags | capi.RENDER_NO_OVERLAY | capi.RENDER_NO_SAMPLE_RAYS)
    views = [MainView(verts, faces, center, W, H) for _ in range(world)]
    try:
        for rank, mv in enumerate(views):
            mv.set_shard(rank, world, 16)
            if rank:
                for b in bufs:
                    mv.redirect_output(b, views[0].device_ptr(b))
        for mv in views:
            render(mv, fn, flags=flags | capi.RENDER_NO_OVERLAY | capi.RENDER_NO_SAMPLE_RAYS)
        root = views[0]
        got = dict(hit_id=root.hit_ids(), rgba8=root.ray_casting_texture(), endpoint=root.mapping_texture())
        for k in ref:
            assert np.array_equal(got[k].view(np.uint8), ref[k].view(np.uint8)), k
        views[1].redirect_output(capi.BUF_HIT_ID_I32, None)  # restored: the shard writes its own buffer again
        render(views[1], fn, flags=flags | capi.RENDER_NO_OVERLAY | capi.RENDER_NO_SAMPLE_RAYS)
        own = views[1].hit_ids()
        assert np.array_equal(own[16:32], ref["hit_id"][16:32])
    finally:
        for mv in views:
            mv.close()


@pytest.mark.parametrize("delta", [1, 2, 3, 4, 5, 7, 10])
@pytest.mark.parametrize("size", [(200, 150), (97, 61), (1920, 1080), (3840, 2160)])
def test_the_three_overlay_kernels_agree_bit_for_bit(product_lib, oracle, scene, delta, size):
    """LINON_OPT_OVERLAY_KERNEL: 2 = the TMA tile kernel (default, csrc/device/overlay_tile.cuh), 1 = the direct kernel
    (overlay_direct.cuh), 0 = the round-1 hand-staged kernel.  All three bodies are validated bit for bit against the oracle on
    the host (tests/test_overlay_emulation_cpu.py); here the device runs them on the same frame - ragged sizes, tiles cut by
    the frame's edge and by the 2 delta border, every tile-kernel instantiation the test names - and the results must be the
    same words, within the oracle's tolerance like test_gpu_units.test_overlay_matches_the_oracle."""
    verts, faces, center = scene
    W, H = size
    if (W, H) == (3840, 2160) and delta not in (1, 4, 5):
        pytest.skip("the C3 frame size exercises the 32 x 64 tiles (delta <= 4) and the switch back to 32 x 32 (delta 5)")
    with MainView(verts, faces, center, W, H) as mv:
        mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
        st = make_settings(overlay_mode=1, central_difference_delta=delta, lyapunov_scaling=50.0 if W < 3000 else 100.0)
        mv.update_settings(st)
        mv.disable_adaptive_sampling()
        mv.set_option(capi.OPT_OVERLAY_KERNEL, 0)
        mv.render()
        mv.sync()
        assert mv.timings().overlay_kernel == 0
        staged_final, staged_ex = mv.texture().copy(), mv.exponents().copy()
        raycast, ep = mv.ray_casting_texture(), mv.mapping_texture()
        for kernel in (1, 2):
            mv.set_option(capi.OPT_OVERLAY_KERNEL, kernel)
            mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)
            mv.sync()
            assert mv.timings().overlay_kernel == kernel  # the tile kernel really ran (a driver without TMA descriptors reports 1)
            final, ex = mv.texture(), mv.exponents()
            assert np.array_equal(final, staged_final)
            assert np.array_equal(ex.view(np.uint32), staged_ex.view(np.uint32))
        if ((W, H) != (97, 61) or delta <= 5) and W < 3000:
            assert (ex >= 0.8).any()
        if W * H <= 200 * 150:
            ref_final, ref_ex = oracle.overlay(raycast, ep, dict(field_weight=1.0, mouse_x=0.5, mouse_y=0.6, overlay_mode=1,
                                                                 central_difference_delta=delta, lyapunov_scaling=50.0))
            assert np.allclose(ex, ref_ex, rtol=1e-5, atol=1e-7)
            assert np.abs(final.astype(int) - ref_final.astype(int)).max() <= 1
        # mode 0: exact copy, exponents untouched (src/overlay.wgsl:116-118)
        mv.update_settings(make_settings(overlay_mode=0, central_difference_delta=delta))
        for kernel in (1, 2):
            mv.set_option(capi.OPT_OVERLAY_KERNEL, kernel)
            mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)
            mv.sync()
            assert np.array_equal(mv.texture(), raycast) and np.array_equal(mv.exponents(), ex)


def test_the_tile_overlay_follows_a_resize_and_a_new_delta(product_lib, scene):
    """The tensor map of the tile kernel describes one endpoint buffer, one frame size and one delta: after resize_texture and
    after a change of central_difference_delta the kernel must read the new image with the new halo."""
    verts, faces, center = scene
    with MainView(verts, faces, center, 160, 120) as mv:
        mv.reload_shader(PredefinedFunction.MirageSphericalSigmoid.to_code())
        mv.disable_adaptive_sampling()
        for (W, H), delta in (((160, 120), 1), ((160, 120), 4), ((224, 96), 4), ((224, 96), 2), ((64, 64), 2)):
            if mv.size != (W, H):
                mv.resize_texture(W, H)
            mv.update_settings(make_settings(overlay_mode=1, central_difference_delta=delta, lyapunov_scaling=50.0))
            mv.set_option(capi.OPT_OVERLAY_KERNEL, 2)
            mv.render()
            mv.sync()
            assert mv.timings().overlay_kernel == 2
            final, ex = mv.texture().copy(), mv.exponents().copy()
            mv.set_option(capi.OPT_OVERLAY_KERNEL, 1)
            mv.render_flags(capi.RENDER_NO_MARCH | capi.RENDER_NO_SAMPLE_RAYS)
            mv.sync()
            assert np.array_equal(mv.texture(), final) and np.array_equal(mv.exponents().view(np.uint32), ex.view(np.uint32))


@pytest.mark.parametrize("fn", MIRAGES)
def test_quiet_zone_sample_rays_stay_on_the_default_paths(view, scene, fn):
    """linon_sample_rays takes the same option (per-lane straight steps outside the zone): the 800 samples of a fixed-step
    frame stay on the default build's paths (rays through the body amplify rounding differences: 5e-3 absolute, the spread
    the emulation shows between production builds and the oracle), colours and alphas are identical, and it is not slower."""
    verts, faces, center = scene
    view.reload_shader(fn.to_code())
    view.update_settings(make_settings(overlay_mode=0))
    view.render_flags(capi.RENDER_ADAPTIVE_OFF)
    view.sync()
    quiet, quiet_ms = view.ray_samples().copy(), view.timings().sample_rays_ms
    view.set_option(capi.OPT_QUIET_ZONE, 0)
    view.render_flags(capi.RENDER_ADAPTIVE_OFF)
    view.sync()
    plain, plain_ms = view.ray_samples().copy(), view.timings().sample_rays_ms
    assert np.allclose(quiet[:, :3], plain[:, :3], rtol=0, atol=5e-3)
    assert np.array_equal(quiet[:, 3:], plain[:, 3:])
    assert quiet_ms <= plain_ms * 1.05
