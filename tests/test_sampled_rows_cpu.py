"""oracle/sampled_rows.py (the checker of full-size frames) against the oracle's own full-frame run: same pixels, same
exclusion masks.  CPU only."""
import numpy as np
import pytest

from linon_b200 import parity
from oracle import np_oracle as npo
from oracle import sampled_rows as sr


def _scene(W, H):
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    return verts, faces, cam, dict(npo.default_settings())


def test_band_centres_are_evenly_spaced_and_inside_the_frame():
    for H, n in ((2160, 22), (1080, 8), (600, 40), (48, 3), (12, 50)):
        centres, stride = sr.band_centres(H, n)
        assert len(centres) >= 1 and stride >= 3
        assert all(b - a == stride for a, b in zip(centres, centres[1:]))
        assert centres[0] - 1 >= 1 and centres[-1] + 1 <= H - 2
        assert len(set(centres)) == len(centres)


@pytest.mark.parametrize("fn,adaptive", [(3, False), (1, True), (7, False)])
def test_sampled_rows_equal_the_full_frame_oracle(oracle, fn, adaptive):
    W, H = 96, 72
    verts, faces, cam, st = _scene(W, H)
    full = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=adaptive)
    _, exps = oracle.overlay(full["rgba8"], full["endpoint"], dict(st, overlay_mode=1))
    rf = sr.reference_frame(verts, faces, cam, st, fn, W, H, n_bands=7, h_initial=0.1, adaptive=adaptive)
    g = rf["graded"]
    assert g.sum() == 7 * W
    for k in ("rgba8", "endpoint", "hit_id", "accepted", "rejected", "prehit"):
        assert np.array_equal(rf["ref"][k][g], full[k][g]), k
    assert np.array_equal(rf["high_lyapunov"][g], parity.high_lyapunov_mask(exps)[g])
    assert (rf["ref"]["hit_id"][~g & ~np.roll(g, 1, 0) & ~np.roll(g, -1, 0)] == -2).all()  # nothing else was marched
    if adaptive:
        alt64 = oracle.march(verts, faces, cam, st, fn, W, H, h_initial=0.1, adaptive=True, use_fp64=True)
        st_ulp = dict(st, field_weight=float(np.nextafter(np.float32(st["field_weight"]), np.float32(0))))
        alt_ulp = oracle.march(verts, faces, cam, st_ulp, fn, W, H, h_initial=0.1, adaptive=True)
        und = parity.disagreement_mask(alt64, full) | parity.disagreement_mask(alt_ulp, full)
        assert np.array_equal(rf["excluded"][g], (und | parity.high_lyapunov_mask(exps))[g])
    # the oracle graded against itself is green, and the cap on the excluded fraction bites
    m = sr.grade(full, rf, fixed_step=not adaptive, max_excluded_fraction=0.5)
    assert m["green"] and m["hit_group_agree"] == 1.0 and m["endpoint_bit_equal"] == 1.0
    assert not sr.grade(full, dict(rf, excluded=g.copy()), fixed_step=not adaptive, max_excluded_fraction=0.5)["green"]


def test_the_transpiled_reference_shader_takes_the_same_sample(oracle):
    from oracle import wgsl_reference as wr
    if not wr.available():
        pytest.skip("oracle/_ref not built")
    W, H = 64, 48
    verts, faces, cam, st = _scene(W, H)
    centres, stride = sr.band_centres(H, 5)
    a = sr.march_bands(verts, faces, cam, st, 3, W, H, centres, stride, h_initial=0.1, adaptive=False, delta=0)
    b = sr.march_bands(verts, faces, cam, st, 3, W, H, centres, stride, h_initial=0.1, adaptive=False, delta=0, impl=wr)
    g = sr.graded_mask(W, H, centres)
    assert np.array_equal(a["rgba8"][g], b["rgba8"][g])
    assert np.array_equal(a["endpoint"][g].view(np.uint32), b["endpoint"][g].view(np.uint32))
