"""bench.py's control flow, rehearsed on CPU: `--dry-run-hostsim` runs the same function the GPU
bench runs (band renderer, pipelined views, statistics pass, shared host image, the JSON line) on
the host-loop test build with gloo, so that a Python-level mistake cannot first show up on the GPU
box. The numbers it prints are marked as a dry run and mean nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks")


def _line(out: str) -> dict:
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_dry_run_single_process(built):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--dry-run-hostsim", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    d = _line(p.stdout)
    for k in REQUIRED:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["data"].startswith("DRY RUN") and d["gpu_launches"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert d["config"]["workload"].startswith("dry-run")


def test_dry_run_four_ranks(built):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "bench.py"), "--gpus", "4", "--dry-run-hostsim", "--steps", "5",
           "--warmup", "3"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _line(p.stdout)
    assert d["n_gpus"] == 4 and len(d["bands"]) == 4 and d["bands"][0][0] == 0 and d["bands"][-1][1] == d["config"]["res_y"]
    assert d["tet_steps_per_view"] > 0 and d["e2e"]["value"] > 0
    assert d["warmup"] >= 6                                  # N > 1 settles the band cuts first
    assert "watchdog" not in p.stderr.lower()


def test_reference_arm_prints_a_line(built):
    """--impl reference: the reference's own CPU path (oracle/_ref when built here, else the port)."""
    # the real arm renders C3 at 1200 x 900 and takes a minute per step: only check that the entry
    # point parses and that a non-zero rank stays silent and exits 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_fallback_after_a_stalled_attempt(built):
    """N > 1 runs in child processes: when the first configuration stalls (here: a deliberate sleep) its
    watchdog ends it on every rank and the conservative configuration runs on a fresh rendezvous."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--dry-run-hostsim", "--steps", "3",
           "--warmup", "3"]
    env = dict(os.environ, C5_BENCH_FAKE_HANG="1", C5_BENCH_STALL_LIMIT="5")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _line(p.stdout)
    assert d["attempts"] == [{"gather": "sendrecv", "lanes": 2, "e2e": "shared-host", "exit_code": 3}]
    assert d["config"]["grazing_kernel"].startswith("after")
    assert d["config"]["views_in_flight"] == 1 and "device-to-host copy on rank 0" in d["e2e"]["api"]
    assert "WATCHDOG" in p.stderr
