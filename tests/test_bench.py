"""bench.py on CPU: what can be checked without a GPU — the reference arm (the one place besides
tests/ that may execute oracle/), the parity record it attaches to a line, and that both arms
describe the workload in the same words. The CUDA arm itself refuses to run without a device."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _line(out: str) -> dict:
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_reference_arm_prints_a_line_on_the_named_workload(built):
    """--impl reference on BASELINE.json configs[0] (the reference's own CPU-runnable case; the
    default workload, configs[2], needs ~8 GB and half a minute per step)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = _line(p.stdout)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["unit"] == "tet-steps/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tet-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench
    from course5_b200 import synth
    mesh, view = synth.make_config("C1")
    assert d["config"] == bench.config_dict("C1", mesh, view)        # the same words as the CUDA arm's line
    assert "600x450" in d["config"]["workload"] and "600x450" in d["cpu_baseline"]["sample"]
    assert d["tet_steps_per_view"] > 6_000_000


def test_reference_arm_other_ranks_stay_silent(built):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("only meaningful without a CUDA device")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "no CPU path" in (p.stderr + p.stdout) and "{" not in p.stdout


def test_parity_record(hostsim_lib, port):
    """The record bench.py attaches to its line: green on a faithful image, red (with counts) on a
    perturbed one. (Here the device functions' host build stands in for the GPU.)"""
    import bench
    from course5_b200 import api, synth
    mesh = synth.kuhn_cube(8, seed=81)
    with api.Context(lib=hostsim_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        ours = ctx.render_raw(api.make_view(120, 90, X=0.4, Y=0.3, lib=hostsim_lib, round_through_float=0))
    want = port.render(mesh.tet_points(), mesh.alpha, mesh.q, res_x=120, res_y=90, X=0.4, Y=0.3)
    rec = bench.parity_record(ours, want, against="the C restatement")
    assert rec["ok"] and rec["pixels_out_of_tolerance"] == 0 and rec["max_rel_err_where_value_ge_1e-3"] < 1e-9
    assert rec["hit_sets_equal"] and rec["per_pixel_steps_equal"] and rec["solid_masks_equal"]
    assert rec["tet_steps"][0] == rec["tet_steps"][1] == want.total_steps
    j, i = np.argwhere(want.steps > 0)[5]
    ours.image[j, i, 1] *= 1.0 + 1e-7
    ours.steps[0, 0] += 1
    rec = bench.parity_record(ours, want, against="the C restatement")
    assert not rec["ok"] and rec["pixels_out_of_tolerance"] == 1 and not rec["per_pixel_steps_equal"]
    assert 0.9e-7 < rec["max_rel_err_where_value_ge_1e-3"] < 1.1e-7 or rec["max_abs_err"] > 0
