"""Offline frame format (linon_b200/frame_io.py): index topology of the ray-sample mesh and the frame round trip.  CPU only."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from linon_b200 import frame_io
from linon_b200._capi import Camera, Counters, Settings

REFERENCE = "/root/reference/src/ray_samples.rs"


def test_ray_sample_indices_follow_create_indices():
    idx = frame_io.ray_sample_indices()
    assert idx.shape == (4 * 99 * 4, 3) and idx.dtype == np.uint16  # draw_indexed(0..4752), src/reference_view.rs:388
    assert idx.max() == 799 and idx.min() == 0                      # 8 rays x 100 samples
    # first segment of side (0, 1, 2), src/ray_samples.rs:29-40
    assert idx[:4].tolist() == [[0, 101, 1], [0, 100, 101], [100, 200, 101], [200, 201, 101]]
    # last side closes the ring: rays 6, 7, 0
    assert idx[3 * 396].tolist() == [600, 701, 601] and idx[3 * 396 + 2].tolist() == [700, 0, 701]


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="reference mount not present")
def test_ray_sample_constants_match_the_reference_source():
    src = open(REFERENCE).read()
    assert int(re.search(r"SAMPLES_PER_POINT: u16 = (\d+)", src).group(1)) == frame_io.SAMPLES_PER_POINT
    sides = re.findall(r"push_indices_for_side\(&mut result, (\d), (\d), (\d)\)", src)
    assert sides == [("0", "1", "2"), ("2", "3", "4"), ("4", "5", "6"), ("6", "7", "0")]
    assert "array_stride: 32" in src


def test_frame_round_trip(tmp_path):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (5, 7, 4), dtype=np.uint8)
    ep = rng.normal(size=(5, 7, 4)).astype(np.float32)
    hit = rng.integers(-1, 36, (5, 7)).astype(np.int32)
    cam, st, cnt = Camera(), Settings(), Counters()
    cam.origin[:] = (1.0, 2.0, 3.0, 0.0)
    st.field_weight, st.overlay_mode, st.central_difference_delta, st.lyapunov_scaling = 1.0, 1, 1, 50.0
    cnt.rays, cnt.accepted_steps = 35, 35 * 5001
    side = frame_io.write_frame(str(tmp_path / "f0"), img, cam, st, "return v;", 1, counters=cnt, endpoint=ep, hit_id=hit)
    assert np.array_equal(frame_io.read_ppm(str(tmp_path / "f0.ppm")), img[..., :3])
    back = json.load(open(tmp_path / "f0.json"))
    assert back == side and back["width"] == 7 and back["height"] == 5 and back["counters"]["accepted_steps"] == 35 * 5001
    assert back["camera"]["origin"][:3] == [1.0, 2.0, 3.0] and back["settings"]["overlay_mode"] == 1
    assert np.array_equal(np.fromfile(tmp_path / "f0.endpoint.f32", "<f4").reshape(5, 7, 4), ep)
    assert np.array_equal(np.fromfile(tmp_path / "f0.hit.i32", "<i4").reshape(5, 7), hit)
    assert len(back["field_function_sha256"]) == 64
