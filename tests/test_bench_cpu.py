"""Host logic of bench.py that needs no GPU: how the parent files what the quiet-zone probe child printed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_probe_output_main_line_then_extras():
    out = "\n".join(["some library chatter", json.dumps(dict(ms_per_step=12.5, vs_default_path=dict(meets_north_star=True))),
                     json.dumps(dict(extra="radius_1500", ms_per_step=9.0)), "{not json",
                     json.dumps(dict(extra="c1_adaptive_level2", error="RuntimeError: x"))])
    p = bench.parse_probe_output(out, "")
    assert p["ms_per_step"] == 12.5 and p["radius_1500"] == dict(ms_per_step=9.0)
    assert p["c1_adaptive_level2"] == dict(error="RuntimeError: x") and "incomplete" not in p


def test_probe_output_keeps_the_main_line_of_a_child_that_died_later():
    p = bench.parse_probe_output(json.dumps(dict(ms_per_step=12.5)) + "\n", "time limit of 300 s reached")
    assert p == dict(ms_per_step=12.5, incomplete="time limit of 300 s reached")
    assert bench.parse_probe_output("", "exit -11: Segmentation fault") == dict(error="exit -11: Segmentation fault")
    assert bench.parse_probe_output(None, "time limit of 300 s reached") == dict(error="time limit of 300 s reached")
    # an extra line without a main line before it is not a result
    assert "error" in bench.parse_probe_output(json.dumps(dict(extra="radius_1500", ms_per_step=9.0)), "exit 1: boom")


def test_reference_arm_on_other_ranks_exits_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_overlay_traffic_comes_from_the_committed_capture_of_that_size_only():
    t = bench.overlay_ncu_traffic(3840, 2160)
    assert t is not None and 0.8 * 28 * 3840 * 2160 < t < 1.2 * 28 * 3840 * 2160  # no wasted re-reads (28 algorithmic B/px)
    assert bench.overlay_ncu_traffic(800, 600) is None
