"""The committed fixtures (tests/golden/) still equal what the oracle produces today.  CPU only."""
import json
import os

import numpy as np

from oracle import np_oracle as npo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_the_golden_fixtures(oracle):
    idx = json.load(open(os.path.join(GOLDEN, "index.json")))
    W, H = idx["width"], idx["height"]
    verts, faces, center = npo.default_scene()
    cam = npo.camera_uniform(npo.default_camera(center, W, H))
    for k in cam:
        assert np.allclose(cam[k], idx["camera"][k])
    assert len(idx["frames"]) == 13
    for entry in idx["frames"]:
        z = np.load(os.path.join(GOLDEN, entry["file"]))
        out = oracle.march(verts, faces, cam, idx["settings"], entry["function_id"], W, H, h_initial=idx["h_initial"],
                           adaptive=entry["adaptive"])
        assert np.array_equal(out["hit_id"], z["hit_id"].astype(np.int32)), entry
        assert np.array_equal(out["rgba8"], z["rgba8"]), entry
        assert np.array_equal(out["endpoint"].view(np.uint32), z["endpoint"].view(np.uint32)), entry
        assert np.array_equal(out["accepted"], z["accepted"].astype(np.int32)), entry
