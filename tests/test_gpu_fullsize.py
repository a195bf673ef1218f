"""Parity of the CUDA path with the oracle at BASELINE.json's FULL frame sizes, on sampled rows (-m gpu).

The oracle cannot render 3840x2160 x 5001 steps in a test, but pixels are independent: oracle/sampled_rows.py marches
bands of three rows around evenly spaced centre rows, grades the centre rows (hit id per surface group, colour within
1/255, endpoint within 1e-3 relative on fixed-step frames) and takes the high-Lyapunov exclusion from the ORACLE's own
endpoints of the band.  Every test asserts a CAP on the excluded fraction (high Lyapunov + pixels the oracle itself cannot
decide in adaptive mode): a checker that excludes everything fails.  The caps are the oracle's own fractions at these
sizes with head-room (oracle/sampled_rows.excluded_cap: 3 % fixed-step / 5 % adaptive, 10 % / 28 % for the two plane
mirages whose frames the reference itself paints as diverging; C3: 1 %, measured 0).
"""
import numpy as np
import pytest

from conftest import camera_dict, settings_dict
from linon_b200 import _capi as capi
from linon_b200.main_view import MainView, PredefinedFunction, make_settings
from oracle import sampled_rows as sr

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

PREDEFINED = [f for f in PredefinedFunction if f != PredefinedFunction.Custom]
CAP_ENHANCE = 0.01
DEFAULT_QUIET = 2  # the library's default level of LINON_OPT_QUIET_ZONE (include/linon_b200.h)


def gpu_frame(mv):
    return dict(hit_id=mv.hit_ids(), rgba8=mv.ray_casting_texture(), endpoint=mv.mapping_texture())


@pytest.fixture(scope="module")
def view_1080p(product_lib, scene):
    verts, faces, center = scene
    mv = MainView(verts, faces, center, 1920, 1080)
    yield mv
    mv.close()


@pytest.mark.parametrize("fn", PREDEFINED)
@pytest.mark.parametrize("adaptive,quiet", [(False, None), (True, None), (True, 0), (True, 1)])
def test_c2_full_size_sampled_rows(view_1080p, scene, fn, adaptive, quiet):
    """C2: 1920x1080, every predefined function, default step size - as the application renders it (adaptive on) and with
    adaptive sampling off (the variant whose endpoints are comparable).  quiet: LINON_OPT_QUIET_ZONE (None = the
    library default, 2 = identity steps outside the mirage's zone in adaptive frames too; 0 = every attempt evaluates the
    field; 1 = the zone in fixed-step frames only)."""
    verts, faces, _ = scene
    mv = view_1080p
    W, H = mv.size
    if quiet is not None and int(fn) > 4:
        pytest.skip("the quiet zone is defined by the mirage bodies only")
    mv.set_option(capi.OPT_QUIET_ZONE, DEFAULT_QUIET if quiet is None else quiet)
    mv.reload_shader(fn.to_code())
    st = make_settings()
    mv.update_settings(st)
    mv.render_flags(0 if adaptive else capi.RENDER_ADAPTIVE_OFF)
    mv.sync()
    rf = sr.reference_frame(verts, faces, camera_dict(mv.camera.uniform()), settings_dict(st), int(fn), W, H, n_bands=8,
                            h_initial=0.1, adaptive=adaptive)
    m = sr.grade(gpu_frame(mv), rf, fixed_step=not adaptive, max_excluded_fraction=sr.excluded_cap(int(fn), adaptive))
    assert m["green"], m
    c = mv.counters()
    if not adaptive:
        assert c.accepted_steps == W * H * 51 and c.rejected_steps == 0


def test_c3_full_size_sampled_rows(product_lib, scene):
    """C3: 3840x2160, Enhance (5001 steps per ray), overlay on - the shipped path (quiet zone on) and the round-1 path (every
    step evaluates the field), both against the oracle on the same rows; exponents of the overlay on those rows as well."""
    verts, faces, center = scene
    W, H = 3840, 2160
    fn = PredefinedFunction.MirageSphericalSigmoid
    with MainView(verts, faces, center, W, H) as mv:
        st = make_settings(overlay_mode=1)
        mv.update_settings(st)
        mv.reload_shader(fn.to_code())
        rf = sr.reference_frame(verts, faces, camera_dict(mv.camera.uniform()), settings_dict(st), int(fn), W, H, n_bands=3,
                                h_initial=0.001, adaptive=True)
        for quiet in (1, 0):
            mv.set_option(capi.OPT_QUIET_ZONE, quiet)
            mv.render_flags(capi.RENDER_ENHANCE)
            mv.sync()
            c = mv.counters()
            assert c.accepted_steps == W * H * 5001 and c.rejected_steps == 0
            q = int(mv.debug_profile()[15])
            assert (q > 0.9 * c.accepted_steps) if quiet else (q == 0)
            m = sr.grade(gpu_frame(mv), rf, fixed_step=True, max_excluded_fraction=CAP_ENHANCE)
            assert m["green"], (quiet, m)
            # the overlay of the GPU frame on the graded rows against the oracle's overlay of ITS frame
            ex = mv.exponents()
            for cy in rf["centres"]:
                band = dict(rgba8=rf["ref"]["rgba8"][[cy - 1, cy - 1, cy, cy + 1, cy + 1]],
                            endpoint=rf["ref"]["endpoint"][[cy - 1, cy - 1, cy, cy + 1, cy + 1]])
                from oracle import cpu_oracle as co
                _, ref_ex = co.overlay(band["rgba8"], band["endpoint"], settings_dict(st))
                ok = np.isclose(ex[cy, 2:-2], ref_ex[2, 2:-2], rtol=2e-2, atol=1e-4)  # endpoints agree to 1e-3, lambda amplifies
                assert ok.mean() >= 0.99, (quiet, cy, ok.mean())


def test_c4_full_size_sampled_rows(view_1080p, scene):
    """C4: the outline workload at 1920x1080 - overlay on, click mode and outline mode; the march is graded on sampled rows,
    the eight outline positions against the oracle's serial scans over the GPU's own exponents, the sample rays against the
    oracle's from the same positions."""
    from oracle import cpu_oracle as co
    verts, faces, _ = scene
    mv = view_1080p
    W, H = mv.size
    fn = PredefinedFunction.MirageSphericalSigmoid
    mv.set_option(capi.OPT_QUIET_ZONE, DEFAULT_QUIET)
    mv.reload_shader(fn.to_code())
    st = make_settings(overlay_mode=1)
    mv.update_settings(st)
    mv.render_flags(0)
    mv.sync()
    cam = camera_dict(mv.camera.uniform())
    rf = sr.reference_frame(verts, faces, cam, settings_dict(st), int(fn), W, H, n_bands=8, h_initial=0.1, adaptive=True)
    m = sr.grade(gpu_frame(mv), rf, fixed_step=False, max_excluded_fraction=sr.excluded_cap(int(fn), True))
    assert m["green"], m
    ex = mv.exponents()
    mv.render_flags(capi.RENDER_OUTLINE)  # searches the exponents of the previous overlay pass (src/main_view.rs:838-846)
    mv.sync()
    pos = mv.read(capi.BUF_SAMPLE_POS)
    ref_pos = co.sample_positions(cam, settings_dict(st), W, H, outline=True, exponents=ex)
    assert np.array_equal(pos, ref_pos)
    samples = mv.ray_samples()
    ref_samples = co.sample_rays(cam, settings_dict(st), int(fn), W, H, ref_pos)
    assert np.array_equal(samples[:, 3:], ref_samples[:, 3:])  # w = 1, colours, alpha 0.5
    close = np.isclose(samples[:, :3], ref_samples[:, :3], rtol=0, atol=5e-3).all(axis=1)
    assert close.mean() >= 0.98, close.mean()  # rays through the body amplify rounding differences


def test_c1_full_size_sampled_rows(product_lib, scene):
    """C1: 800x600, first predefined function, default step size, both sampling modes."""
    verts, faces, center = scene
    fn = PredefinedFunction.MirageSpherical
    with MainView(verts, faces, center, 800, 600) as mv:
        mv.reload_shader(fn.to_code())
        st = make_settings()
        mv.update_settings(st)
        for adaptive in (True, False):
            mv.render_flags(0 if adaptive else capi.RENDER_ADAPTIVE_OFF)
            mv.sync()
            rf = sr.reference_frame(verts, faces, camera_dict(mv.camera.uniform()), settings_dict(st), int(fn), 800, 600, n_bands=24,
                                    h_initial=0.1, adaptive=adaptive)
            m = sr.grade(gpu_frame(mv), rf, fixed_step=not adaptive, max_excluded_fraction=sr.excluded_cap(int(fn), adaptive))
            assert m["green"], (adaptive, m)
