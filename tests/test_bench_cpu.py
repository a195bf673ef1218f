"""Host logic of bench.py that needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_config_lines_of_the_child_are_filed_and_graded():
    lines = [dict(config="C1", parity=dict(green=True)), dict(config="C2", parity=dict(green=True)), dict(config="C4"), dict(done=True)]
    c = bench.collect_configs(lines, "")
    assert c["complete"] and c["all_green"] and c["graded"] == 2 and len(c["lines"]) == 3 and "incomplete" not in c
    c = bench.collect_configs(lines[:2] + [dict(config="C2", parity=dict(green=False))], "time limit of 420 s reached")
    assert not c["complete"] and not c["all_green"] and c["incomplete"].startswith("time limit")
    assert not bench.collect_configs([], "exit -11: Segmentation fault")["all_green"]


def test_a_stale_ncu_capture_is_refused(tmp_path):
    p = tmp_path / "prof.json"
    p.write_text(json.dumps([dict(kernel="linon_march_tiles", cubin_hash="00000000deadbeef", fp32_flops_executed=1.0)]))
    k, note = bench.committed_profile(str(p), "linon_march_tiles", "00000000deadbeef")
    assert k is not None and note == "ok"
    k, note = bench.committed_profile(str(p), "linon_march_tiles", "0123456789abcdef")
    assert k is None and "stale capture refused" in note
    k, note = bench.committed_profile(str(tmp_path / "missing.json"), "x", None)
    assert k is None and "not committed" in note
    assert len(bench.static_source_hash()) == 16


def test_committed_captures_name_what_they_profiled():
    """Every committed r2 capture bench.py reads carries the identity of the binary it profiled."""
    for path, key in ((bench.MARCH_PROFILE, "cubin_hash"), (bench.OVERLAY_PROFILE, "source_hash")):
        if os.path.exists(path):
            assert any(k.get(key) for k in json.load(open(path))), path


def test_both_arms_name_the_same_workload():
    a = bench.workload_config(1, *bench.workload_size(1))
    assert a["width"] == 3840 and a["height"] == 2160 and a["rk4_steps_per_step"] == 3840 * 2160 * 5001
    for n in (2, 4, 8):
        w, h = bench.workload_size(n)
        assert abs(w * h / (3840 * 2160) - n) < 0.01 * n


def test_cpu_arm_counts_its_threads_from_the_affinity_mask_not_from_omp_num_threads(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")  # what torchrun exports
    assert bench.host_threads() == len(os.sched_getaffinity(0))


def test_reference_arm_on_other_ranks_exits_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_every_core_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1; the reference arm must still use all host threads (round-1 finding)."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["config"] == bench.workload_config(2, *bench.workload_size(2))


def test_the_committed_march_captures_are_of_the_kernels_this_tree_builds(tmp_path):
    """bench.py refuses an ncu capture whose code hash is not the running kernel's.  NVRTC's output for one source text and one
    option list is the same on every host (this container without a GPU and every B200 box of rounds' calls produced the same
    two hashes), so the tie can be checked here: the headline variant (C3: MirageSphericalSigmoid, Enhance, fast math, quiet
    zone, collapsed runs) and its step-by-step twin, compiled by this tree in a fresh process, hash to what
    profiles/r2_march_c3*.json say they profiled.  A change to the march template without a new capture fails here first."""
    code = (
        "import os\n"
        "from linon_b200 import _capi as capi\n"
        "from linon_b200.main_view import PredefinedFunction as P, compile_check\n"
        "for collapse in ('1', '0'):\n"
        "    os.environ.update(LINON_CHECK_QUIET_ZONE='2', LINON_CHECK_COLLAPSE=collapse)\n"
        "    compile_check(P.MirageSphericalSigmoid.to_code(), flags=capi.RENDER_ENHANCE, math_mode=capi.MATH_FAST)\n"
        "    print(open(os.environ['LINON_DUMP_CUBIN_HASH']).read().strip())\n")
    env = dict(os.environ, LINON_DUMP_CUBIN_HASH=str(tmp_path / "h.txt"), PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    built = r.stdout.strip().splitlines()[-2:]
    for path, h in zip((bench.MARCH_PROFILE, bench.MARCH_PROFILE_STEP_BY_STEP), built):
        committed = [k["cubin_hash"] for k in json.load(open(path)) if k.get("kernel") == "linon_march_tiles"]
        assert committed == [h], (path, committed, h)
