"""bench.py contract, the part that runs without a GPU: the reference arm (`--impl reference`) prints ONE JSON line with the keys the
driver reads, times the reference's CPU path on the host cores it is given, and does no work on ranks other than 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_the_contract_line():
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the arm must not inherit it (round 1: a one-core reference arm)
    lines = run(["--impl", "reference", "--workload", "ne27000_explicit", "--steps", "1", "--warmup", "1"], env={"OMP_NUM_THREADS": "1"})
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "time steps/s" and j["unit"] == "steps/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and abs(j["ms_per_step"] * j["value"] - 1e3) < 1e-6 * 1e3
    assert j["config"]["workload"] == "ne27000_explicit" and j["dtype"] == "f64" and j["vs_baseline"] is None
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == j["value"] and cb["sample"]
    assert cb["cores"] == min(8, len(os.sched_getaffinity(0))) if cb["kind"] == "reference" else cb["cores"] >= 1
    assert j["e2e"] == dict(value=j["value"], unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert j["gpu_launches"] == 0
    # the shipped mesh is timed at its own size: nothing extrapolated, so no `consistency` escape hatch
    assert cb["extrapolated"] is False and "consistency" not in j


def test_reference_arm_is_silent_on_other_ranks():
    assert run(["--impl", "reference", "--workload", "ne27000_explicit", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []
