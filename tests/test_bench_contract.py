"""bench.py's JSON contract, checked on CPU through the reference arm (the OpenMP oracle port): one JSON line
with the keys the driver reads.  The GPU arm prints the same keys plus roofline / cpu_baseline / clocks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--image", "96", "--ref-volume", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "raycast_throughput" and d["unit"] == "Gsamples/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must still use every core it may run on."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--image", "96", "--ref-volume", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_timed_views_span_the_orbit_whatever_the_step_count():
    """The headline must not depend on --steps: K timed views are always spread over the whole 64-view orbit."""
    sys.path.insert(0, ROOT)
    import bench
    for k in (1, 3, 20, 64, 100):
        v = bench.timed_views(k)
        assert len(v) == k and all(0 <= x < 64 for x in v)
    assert bench.timed_views(64) == list(range(64))
    v20 = bench.timed_views(20)
    assert v20[0] == 0 and v20[-1] >= 57 and len(set(v20)) == 20
    assert max(b - a for a, b in zip(v20, v20[1:])) <= 4
    assert bench.timed_views(130)[64:70] == [0, 1, 2, 3, 4, 5]


def test_roofline_bound_follows_the_ray_density():
    """SURVEY.md §8d / round-1 verdict: the 32-bytes-per-sample figure is a bound only where every sample brings its own
    texels.  Rays one voxel apart or closer share them on chip, the figure can exceed the HBM peak there, and the line must
    then report the texture pipe's roofline as the bound and keep the other figure marked as information."""
    sys.path.insert(0, ROOT)
    import bench
    hbm = 6553.6
    # the headline: 1024^3 at 1024^2, rays two voxels apart, 24.2 M samples in 0.2 ms
    r = bench.ray_roofline(24.2e6, 0.2, 1024, 1024, 1024, hbm, "measured", 886e6, None)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["achieved"] - 24.2e6 * 32 / 0.2e-3 / 1e9) < 1e-6
    assert 0 < r["frac"] < 1 and r["ray_spacing_voxels"] == 2.0 and 0 < r["traffic_frac"] < 1
    assert r["l1tex"]["peak"] == bench.L1TEX_FALLBACK_GSAMPLES and "profiles/" in r["l1tex"]["peak_source"]
    # the fixed 2048^2 frame: one voxel apart, 97 M samples in 0.43 ms -> 7.2 TB/s "algorithmic", not a bound
    live = {"tex3d_linear_l1_gsamples_per_s": 560.0}
    s = bench.ray_roofline(97e6, 0.43, 1024, 2048, 2048, hbm, "measured", None, live)
    assert s["bound"] == "l1tex" and s["peak"] == 560.0 and 0 < s["frac"] < 1 and "run in this process" in s["peak_source"]
    assert s["hbm_algorithmic"]["frac"] > 1 and "not a bound" in s["hbm_algorithmic"]["note"]
    assert "traffic_frac" not in s and s["traffic"] is None
    # N = 8 weak: 4096 x 2048, half a voxel apart
    assert bench.ray_roofline(24.6e6, 0.118, 1024, 4096, 2048, hbm, "measured", 173e6, live)["bound"] == "l1tex"
