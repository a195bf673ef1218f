"""linon_sharded_* through ctypes (include/linon_b200.h): a frame rendered by several shards equals the one-view frame bit for bit.

On a one-GPU box the shards share device 0 (peer-store mode; the multi-shard logic - row dealing, redirected stores,
event ordering, root overlay - is the same); with >= 2 GPUs the shards sit on distinct devices and the NCCL gather
mode (the north star's "single NCCL gather ... to GPU 0") runs as well.  -m gpu."""
import ctypes as C

import numpy as np
import pytest

from linon_b200 import _capi as capi
from linon_b200.main_view import MainView, PredefinedFunction, ShardedFrame, make_settings

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

BUFS = (capi.BUF_HIT_ID_I32, capi.BUF_RAYCAST_RGBA8, capi.BUF_ENDPOINT_F32X4, capi.BUF_FINAL_RGBA8, capi.BUF_EXPONENTS_F32,
        capi.BUF_RAY_SAMPLES)


def n_gpus():
    import torch
    return torch.cuda.device_count()


def whole_frame(scene, W, H, fn, flags, settings):
    verts, faces, center = scene
    with MainView(verts, faces, center, W, H) as mv:
        cam = mv.camera.uniform()
        mv.reload_shader(fn.to_code())
        mv.update_settings(settings)
        mv.render_flags(flags)
        mv.sync()
        return cam, {b: mv.read(b) for b in BUFS}, mv.counters()


def sharded_frame(scene, devices, mode, W, H, fn, flags, settings, cam, rows_per_block=16, frames=2):
    verts, faces, _ = scene
    with ShardedFrame(devices, verts, faces, W, H, rows_per_block=rows_per_block, mode=mode) as sf:
        assert sf.world == len(devices)
        sf.set_camera_uniform(cam)
        sf.reload_shader(fn.to_code())
        sf.update_settings(settings)
        for _ in range(frames):  # back-to-back frames without a host sync in between: the event ordering must hold
            sf.render_flags(flags)
        sf.sync()
        out = {b: sf.read_root(b) for b in BUFS}
        steps = sum(int(sf.counters(r).accepted_steps) for r in range(sf.world))
        rays = sum(int(sf.counters(r).rays) for r in range(sf.world))
        t = sf.timings(0)
        assert t.march_ms > 0 and t.overlay_ms > 0  # the root reports both passes of the frame
        return out, steps, rays


@pytest.mark.parametrize("flags", [capi.RENDER_ADAPTIVE_OFF, 0, capi.RENDER_ENHANCE])
@pytest.mark.parametrize("world", [2, 3])
def test_peer_store_shards_equal_the_whole_frame(product_lib, scene, world, flags):
    W, H = (96, 80) if flags != capi.RENDER_ENHANCE else (64, 40)
    fn = PredefinedFunction.MirageSphericalSigmoid
    st = make_settings(overlay_mode=1)
    cam, ref, c = whole_frame(scene, W, H, fn, flags, st)
    g = n_gpus()
    devices = [r % g for r in range(world)]  # distinct devices where the box has them
    got, steps, rays = sharded_frame(scene, devices, capi.SHARD_PEER_STORES, W, H, fn, flags, st, cam)
    for b in BUFS:
        assert np.array_equal(got[b].view(np.uint8), ref[b].view(np.uint8)), b
    assert steps == int(c.accepted_steps) and rays == W * H


def test_nccl_gather_shards_equal_the_whole_frame(product_lib, scene):
    g = n_gpus()
    if g < 2:
        pytest.skip("the NCCL gather needs one communicator per distinct device (>= 2 GPUs)")
    W, H = 128, 100  # a ragged last block
    fn = PredefinedFunction.MirageSpherical
    st = make_settings(overlay_mode=1, central_difference_delta=2)
    for flags in (capi.RENDER_ADAPTIVE_OFF, 0):
        cam, ref, c = whole_frame(scene, W, H, fn, flags, st)
        got, steps, rays = sharded_frame(scene, list(range(min(g, 4))), capi.SHARD_NCCL_GATHER, W, H, fn, flags, st, cam, rows_per_block=8)
        for b in BUFS:
            assert np.array_equal(got[b].view(np.uint8), ref[b].view(np.uint8)), b
        assert steps == int(c.accepted_steps) and rays == W * H


def test_a_bad_body_changes_no_shard_and_errors_are_reported(product_lib, scene):
    verts, faces, _ = scene
    with ShardedFrame([0, 0], verts, faces, 64, 48) as sf:
        with pytest.raises(Exception) as e:
            sf.reload_shader("return vec3<f32>(1.0, 2.0;")
        assert "shard 0" in str(e.value) or "error" in str(e.value).lower()
        with pytest.raises(Exception):
            sf.render_flags(0)  # no camera yet
    with pytest.raises(Exception):
        ShardedFrame([0, 99], verts, faces, 64, 48)
    with pytest.raises(Exception):
        ShardedFrame([], verts, faces, 64, 48)


def test_resize_rewires_the_peer_stores(product_lib, scene):
    verts, faces, center = scene
    fn = PredefinedFunction.TranslationX
    st = make_settings(overlay_mode=1)
    g = n_gpus()
    with ShardedFrame([0, 1 % g], verts, faces, 64, 48, rows_per_block=8) as sf:
        for (W, H) in ((64, 48), (80, 56)):
            cam, ref, _ = whole_frame(scene, W, H, fn, capi.RENDER_ADAPTIVE_OFF, st)
            sf.resize_texture(W, H)
            sf.set_camera_uniform(cam)
            sf.reload_shader(fn.to_code())
            sf.update_settings(st)
            sf.render_flags(capi.RENDER_ADAPTIVE_OFF)
            sf.sync()
            for b in (capi.BUF_HIT_ID_I32, capi.BUF_ENDPOINT_F32X4, capi.BUF_FINAL_RGBA8):
                assert np.array_equal(sf.read_root(b).view(np.uint8), ref[b].view(np.uint8)), (W, H, b)
