"""Parity of the CUDA path (libc5gpu.so through the C ABI, on a B200) with the oracle."""
import numpy as np
import pytest

import render_checks as rc
from cases import GOLDEN_CASES, reference_solids
from course5_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden(gpu_lib, name):
    rc.check_golden(gpu_lib, name)


@pytest.mark.parametrize("flags", [dict(X=0.0, Y=0.0), dict(X=0.4, Y=0.7, I=-0.03, alpha_limit=1.2),
                                   dict(X=0.5, Y=1.9, I=0.3)])
def test_against_port(gpu_lib, port, flags):
    rc.check_against_port(gpu_lib, port, synth.kuhn_cube(10, seed=41), 200, 150, flags)


def test_config_c1_with_reference_solids(gpu_lib, port):
    """BASELINE.json configs[0]: 32^3 lattice, 196 608 tets, 600 x 450, -X 0 -Y 0, plus the solids
    the reference always renders (main.cpp:110-116,127)."""
    mesh, view = synth.make_config("C1")
    solids = reference_solids(view["D"])
    flags = dict(X=view["X"], Y=view["Y"], I=view["I"], alpha_limit=view["alpha_limit"])
    rc.check_against_port(gpu_lib, port, mesh, view["res_x"], view["res_y"], flags, solids=solids)


def test_config_c2b_cavity(gpu_lib, port):
    """configs[1] geometry with its inner sphere removed: rays leave and re-enter the mesh."""
    mesh, view = synth.make_config("C2b", n=40)
    flags = dict(X=0.3, Y=0.4, I=view["I"], alpha_limit=view["alpha_limit"])
    rc.check_against_port(gpu_lib, port, mesh, 800, 600, flags)


def test_grazing_rays(gpu_lib, port):
    """Edge-on jittered walls: rays with one crossing per cell go through the warp-per-ray kernel."""
    rc.check_grazing_rays(gpu_lib, port, n=24, res=(480, 360), debug_key="graze_list", debug_value=128)


def test_grazing_list_overflow(gpu_lib):
    """A 96-cell wall seen edge-on gives rays ~96 crossings: with the collection shrunk to 128
    entries (truncation at > 64) the overflow path runs; the image must not change by a bit."""
    mesh = synth.kuhn_cube(96, seed=49)
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        v = api.make_view(1200, 900, X=0.5, Y=0.0, alpha_limit=3.0, lib=gpu_lib, round_through_float=0)
        full = ctx.render_raw(v)
        ctx.debug_set("graze_list", 128)
        small = ctx.render_raw(v)
    assert full.stats["grazing_rays"] > 1000 and full.stats["walk_errors"] == 0
    assert small.stats["grazing_rays"] == full.stats["grazing_rays"]
    assert np.array_equal(small.image, full.image) and np.array_equal(small.steps, full.steps)
    assert int(full.steps.sum()) == full.stats["tet_steps"]


def test_sweep_views(gpu_lib, port):
    mesh = synth.kuhn_cube(16, seed=45)
    for k in range(6):
        rc.check_against_port(gpu_lib, port, mesh, 300, 226, dict(X=0.4, Y=2.0 * k / 6, I=-0.03))


def test_graded_mesh(gpu_lib, port):
    rc.check_against_port(gpu_lib, port, synth.kuhn_cube(20, seed=43, grade_beta=1.5), 400, 300,
                          dict(X=0.45, Y=1.2))


def test_row_bands(gpu_lib):
    rc.check_row_bands_equal_full_image(gpu_lib)


def test_round_through_float(gpu_lib):
    rc.check_round_through_float(gpu_lib)


def test_out_of_window(gpu_lib):
    rc.check_out_of_window_geometry_is_background(gpu_lib)


def test_topology_errors(gpu_lib):
    rc.check_topology_errors(gpu_lib)


def test_tiny_meshes(gpu_lib, port):
    rc.check_single_tet_and_tiny_meshes(gpu_lib, port)


def test_uniform_medium(gpu_lib):
    rc.check_uniform_medium_kat(gpu_lib, n=24, res=(480, 360))


def test_full_size_properties_config_c3(gpu_lib):
    """BASELINE.json configs[2] at full size (7 986 000 tets, 2400 x 1800, --alpha_limit 3.0 -X 0.5):
    too big for the oracle to finish in seconds, so parity is checked through exact properties."""
    mesh, view = synth.make_config("C3")
    flags = dict(X=view["X"], Y=view["Y"], I=view["I"], alpha_limit=view["alpha_limit"])
    base = rc.check_scaling_properties(gpu_lib, mesh, view["res_x"], view["res_y"], flags)
    assert base.stats["walk_errors"] == 0
    assert base.stats["tet_steps"] > 3e8
    assert np.isfinite(base.image).all()


# ---- oracle/_ref (the unmodified reference) at the sizes the benchmark lines are quoted on ----------
# BASELINE.json configs[1], [2] and [3] as configured: same mesh, same flags, same resolution, with
# the Roche lobe and sphere the reference always renders (main.cpp:110-116,127). The hit/miss set
# is decided here by orientation tests along a walk and there by scanline fills (plane.cpp:57-142):
# near-ties scale with pixels x edges, so this is where they must agree.

def _upload_with_solids(ctx, mesh, D):
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    roche, sphere = reference_solids(D)
    ctx.clear_solids()
    ctx.upload_solids(roche, True)
    ctx.upload_solids(sphere, False)


def test_config_c2_against_reference(gpu_lib, ref, tmp_path):
    """configs[1]: 998 250-tet sphere-in-cube grid, 1200 x 900, --alpha_limit 2.5."""
    mesh, view = synth.make_config("C2")
    assert mesh.n_tets == 998250
    r = rc.reference_process("C2", [{}], tmp_path / "c2.npz")
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        _upload_with_solids(ctx, mesh, view["D"])
        v = api.make_view(view["res_x"], view["res_y"], X=view["X"], Y=view["Y"], I=view["I"],
                          alpha_limit=view["alpha_limit"], lib=gpu_lib, round_through_float=0)
        rc.assert_matches_reference(ctx.render_raw(v), r, 0, what="C2: ")


def test_configs_c3_and_c4_against_reference(gpu_lib, ref, tmp_path):
    """configs[2]: 7 986 000 tets at 2400 x 1800, --alpha_limit 3.0 -X 0.5 (the README example, the
    workload bench.py measures), and five views of configs[3]'s 360-view sweep of the same mesh
    (1200 x 900, -X 0.4 -I -0.03 -D 0.1, -Y 2k/360; utility/rotate_traces.py:8-9,18). The reference
    needs ~30 s and ~8 GB for the first and ~7 s for each of the others."""
    import json
    import os
    import subprocess
    import sys
    mesh, c3 = synth.make_config("C3")
    _, c4 = synth.make_config("C4", n=2)           # flags only
    sweep = [dict(Y=2.0 * k / 360) for k in (0, 75, 150, 225, 300)]
    # both reference runs start now and work on the host cores while the GPU renders
    runner = os.path.join(os.path.dirname(rc.__file__), "ref_runner.py")
    procs = [subprocess.Popen([sys.executable, runner, name, str(tmp_path / f"{name}.npz"), json.dumps(ov)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for name, ov in (("C3", [{}]), ("C4", sweep))]
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        _upload_with_solids(ctx, mesh, c3["D"])
        v = api.make_view(c3["res_x"], c3["res_y"], X=c3["X"], Y=c3["Y"], I=c3["I"], alpha_limit=c3["alpha_limit"],
                          lib=gpu_lib, round_through_float=0)
        img3 = ctx.render_raw(v)
        _upload_with_solids(ctx, mesh, c4["D"])
        imgs4 = []
        for ov in sweep:
            v = api.make_view(c4["res_x"], c4["res_y"], X=c4["X"], Y=ov["Y"], I=c4["I"], alpha_limit=c4["alpha_limit"],
                              lib=gpu_lib, round_through_float=0)
            imgs4.append(ctx.render_raw(v))
    for p, name in zip(procs, ("C3", "C4")):
        _, err = p.communicate(timeout=1500)
        assert p.returncode == 0, f"the reference did not finish on {name}: rc {p.returncode}\n{err[-2000:]}"
    r3 = np.load(str(tmp_path / "C3.npz"))
    assert img3.stats["tet_steps"] > 3e8 and img3.stats["solid_pixels"] > 0
    rc.assert_matches_reference(img3, r3, 0, what="C3 2400x1800: ")
    r4 = np.load(str(tmp_path / "C4.npz"))
    for k, img in enumerate(imgs4):
        assert img.stats["solid_pixels"] > 0
        rc.assert_matches_reference(img, r4, k, what=f"C4 view {k}: ")


def test_render_device_writes_the_band(gpu_lib):
    import torch
    mesh = synth.kuhn_cube(8, seed=46)
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        full, _ = ctx.render(api.make_view(128, 96, X=0.4, Y=0.2, lib=gpu_lib))
        band = torch.zeros((40, 128, 2), dtype=torch.float64, device="cuda:0")
        v = api.make_view(128, 96, X=0.4, Y=0.2, lib=gpu_lib, row_begin=30, row_end=70)
        ctx.render_device(v, band.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert np.array_equal(band.cpu().numpy(), full[30:70])


def test_sibling_context(gpu_lib):
    rc.check_sibling_context(gpu_lib)


def test_two_views_in_flight(gpu_lib):
    """BandRenderer's two lanes (context + sibling, two streams): pipelined views, some of them
    grazing (their queue tags and the side-stream kernel are per context), equal one-at-a-time renders."""
    import torch
    from course5_b200.dist import BandRenderer
    mesh = synth.kuhn_cube(20, seed=54)
    dev = torch.device("cuda", 0)
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        br = BandRenderer(ctx, device=dev, rank=0, world=1)
        views = [api.make_view(400, 300, X=X, Y=Y, lib=gpu_lib) for X, Y in
                 ((0.5, 0.0), (0.4, 0.3), (0.0, 0.5), (0.45, 1.2), (0.5, 1.0), (0.3, 1.7))]
        for batch in (views[:3], views[3:], views[1:4]):
            imgs = [br.render(v, stats=False, pipeline=True)[0] for v in batch]
            br.finish()
            torch.cuda.synchronize(dev)
            for v, img in zip(batch, imgs):
                want, _ = ctx.render(v)
                assert np.array_equal(img.cpu().numpy(), want, equal_nan=True)
        assert br.kernel_launches() > ctx.kernel_launches()      # the sibling did half of the work
        br.close()


def test_course_cli_end_to_end(gpu_lib, port, tmp_path):
    """The drop-in: `course -f grid.vtk -d out.vti ...` against the reference's own file-to-file
    flow (oracle/_ref when present, else the restatement + float cast). .vti values are doubles
    that went through float (plane.cpp:165-166), so the gate is: equal, or one float ulp apart on
    at most 1e-4 of the pixels (a 1e-15 difference can flip the float rounding)."""
    import subprocess
    from course5_b200 import hostlib
    from oracle import refbind
    from parity import float_ulp_distance
    import os
    mesh = synth.kuhn_cube(12, seed=61)
    src = str(tmp_path / "grid.vtk")
    synth.write_legacy_vtk(src, mesh)
    flags = dict(X=0.4, Y=0.3, D=0.1, I=-0.03, alpha_limit=2.0)
    dst = str(tmp_path / "ours.vti")
    cmd = [hostlib.COURSE_EXE, "-f", src, "-d", dst, "-j4", "-x", "320", "-y", "240", "-X", "0.4", "-Y", "0.3",
           "-D", "0.1", "-I", "-0.03", "--alpha_limit", "2.0", "--stats"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.splitlines()
    assert lines[0] == "Defined grid resolution: 320x240"               # main.cpp:85
    assert lines[1] == f"Source file: {src}"
    assert lines[2] == "Number of parallel threads: 4"
    assert lines[3] == "Initial rotate angle of roche lobe: 0.1 Pi"
    assert lines[7] == "Limit alpha value: 2"
    assert lines[8].startswith("Loading data with VTK lib and other preparations completed in ")
    assert lines[9].startswith("Ray-tracing completed in ") and lines[9].endswith(" ms. ")
    assert lines[10] == "Result exported. Calculations completed."
    ours = hostlib.read_vti(dst)
    assert ours.shape == (240, 320, 2)
    if os.path.exists(refbind.REF_SO):
        ref_dst = str(tmp_path / "ref.vti")
        refbind.Ref().run_files(src, ref_dst, res_x=320, res_y=240, threads=4, **flags)
        want = hostlib.read_vti(ref_dst)
    else:
        roche, sphere = reference_solids(0.1)
        img = port.render(mesh.tet_points(), mesh.alpha, mesh.q, res_x=320, res_y=240, X=0.4, Y=0.3, I=-0.03,
                          alpha_limit=2.0, solid_rot=roche, solid_static=sphere)
        want = np.stack([img.tau, img.inten], axis=-1).astype(np.float32).astype(np.float64)
    assert np.array_equal(np.isnan(ours), np.isnan(want))
    ok = ~np.isnan(want)
    ulp = float_ulp_distance(ours[ok], want[ok])
    assert ulp.max() <= 1
    assert (ulp > 0).mean() <= 1e-4


def test_multi_device_context_with_nccl_gather(gpu_lib):
    """c5_create over several devices: bands rendered concurrently, one grouped ncclSend/ncclRecv."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    rc.check_multi_device_context(gpu_lib, tuple(range(min(n, 4))))


@pytest.mark.parametrize("flags", [dict(X=0.0, Y=0.0), dict(X=0.4, Y=0.7, I=-0.03, alpha_limit=1.2)])
def test_fp32_variant(gpu_lib, port, flags):
    rc.check_fp32_variant(gpu_lib, port, synth.kuhn_cube(24, seed=47), 480, 360, flags)


def test_course_sweep_frames_equal_single_runs(gpu_lib, tmp_path):
    """`course --frames N` (mesh resident, rotations change) == N separate process-per-frame runs,
    the way utility/rotate_traces.py drives the reference."""
    import subprocess
    from course5_b200 import hostlib
    mesh = synth.kuhn_cube(8, seed=62)
    src = str(tmp_path / "grid.vtk")
    synth.write_legacy_vtk(src, mesh, binary=True)
    base = [hostlib.COURSE_EXE, "-f", src, "-x", "200", "-y", "150", "-X", "0.4", "-I", "-0.03"]
    p = subprocess.run(base + ["-d", str(tmp_path / "sweep.vti"), "-Y", "0.0", "--frames", "3", "--sweep_y_to", "1.5"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    for k in range(3):
        y = 0.0 + 1.5 * k / 3
        single = str(tmp_path / f"single_{k}.vti")
        q = subprocess.run(base + ["-d", single, "-Y", repr(y)], capture_output=True, text=True)
        assert q.returncode == 0, q.stdout + q.stderr
        a = hostlib.read_vti(str(tmp_path / f"sweep_{k}.vti"))
        b = hostlib.read_vti(single)
        assert np.array_equal(a, b, equal_nan=True)


def test_full_size_properties_config_c5(gpu_lib):
    """BASELINE.json configs[4] geometry at full size (50 192 562 tets, graded, 4800 x 3600, -X 0.4
    -Y 0.3) with a uniform medium: the recurrence telescopes, so per pixel
    I == (q / a^) (1 - exp(-a^ tau / a)) whatever the 2.2 G tet-steps in between did."""
    mesh = synth.kuhn_cube(203, seed=5, grade_beta=1.5, scalars="const")   # a = 1.5, q = 0.75
    assert mesh.n_tets == 50192562
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        info = ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        assert info.n_boundary_faces == 203 * 203 * 12
        del mesh
        for limit in (2.5, 0.9):
            v = api.make_view(4800, 3600, X=0.4, Y=0.3, alpha_limit=limit, lib=gpu_lib, round_through_float=0)
            img = ctx.render_raw(v)
            assert img.stats["walk_errors"] == 0 and img.stats["tet_steps"] > 2_000_000_000
            assert int(img.steps.sum()) == img.stats["tet_steps"]
            hit = img.hit
            a_hat = min(1.5, limit)
            want = 0.75 / a_hat * (1.0 - np.exp(-a_hat * img.tau[hit] / 1.5))
            assert np.allclose(img.inten[hit], want, rtol=1e-10, atol=1e-13)
            assert np.all(img.image[~hit] == 0.0)
            # the cube's silhouette: tau / a is the chord length, at most the space diagonal (+ jitter)
            assert img.tau.max() / 1.5 < 1.75


def test_pinned_output_is_written_in_place(gpu_lib):
    """c5_render into a pinned host buffer (the walk stores over PCIe, no D2H copy) == into a pageable one."""
    import torch
    mesh = synth.kuhn_cube(10, seed=63)
    solids = reference_solids(0.0)
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        ctx.upload_solids(solids[0], True)
        ctx.upload_solids(solids[1], False)
        v = api.make_view(300, 226, X=0.4, Y=0.6, lib=gpu_lib)
        pageable = np.full((226, 300, 2), -7.0)
        ctx.render(v, out=pageable)
        pinned_t = torch.full((226, 300, 2), -7.0, dtype=torch.float64).pin_memory()
        pinned = pinned_t.numpy()
        _, st = ctx.render(v, out=pinned)
        assert np.array_equal(pinned, pageable, equal_nan=True)
        assert np.isnan(pinned).any() and (pinned == 0).any() and st["tet_steps"] > 0
        band = api.make_view(300, 226, X=0.4, Y=0.6, lib=gpu_lib, row_begin=100, row_end=150)
        pinned[:] = -7.0
        ctx.render(band, out=pinned)
        assert np.array_equal(pinned[100:150], pageable[100:150], equal_nan=True)
        assert np.all(pinned[:100] == -7.0) and np.all(pinned[150:] == -7.0)


def test_solid_mask_high_resolution(gpu_lib, port):
    rc.check_solid_mask_high_resolution(gpu_lib, port, res=(2400, 1800))


def test_solid_mask_tile_sizes(gpu_lib, port):
    rc.check_solid_mask_tile_sizes(gpu_lib, port, res=(800, 600))


def test_static_solid_mask_cache(gpu_lib, port):
    rc.check_static_solid_mask_cache(gpu_lib, port)


def test_search_budget(gpu_lib, port):
    rc.check_search_budget(gpu_lib, port, n=20, res=(400, 300))


def test_submit_wait_lanes(gpu_lib):
    """c5_render_submit / c5_render_wait with page-locked outputs: views in flight on lanes, stored in place."""
    import torch
    keep = []

    def pinned(shape):
        t = torch.zeros(shape, dtype=torch.float64).pin_memory()
        keep.append(t)
        return t.numpy()

    rc.check_submit_wait(gpu_lib, pinned=pinned)
    rc.check_submit_wait(gpu_lib)           # pageable outputs: the copy path


def test_misaligned_output_is_rejected_or_copied(gpu_lib):
    import torch
    keep = []

    def device_buffer(nbytes):
        keep.append(torch.zeros(nbytes, dtype=torch.uint8, device="cuda:0"))
        return keep[-1].data_ptr()

    rc.check_output_alignment(gpu_lib, device_buffer=device_buffer)
    # a page-locked host buffer that is only 8-byte aligned takes the copy path too
    mesh = synth.kuhn_cube(5, seed=65)
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        v = api.make_view(64, 48, X=0.4, Y=0.2, lib=gpu_lib)
        want, _ = ctx.render(v)
        t = torch.zeros(64 * 48 * 2 + 1, dtype=torch.float64).pin_memory()
        odd = t.numpy()[1:] if t.data_ptr() % 16 == 0 else t.numpy()[:-1]
        got, _ = ctx.render(v, out=odd.reshape(48, 64, 2))
        assert np.array_equal(got, want)


def test_timeline_of_pipelined_views(gpu_lib):
    """c5_debug_set("timeline") / c5_timeline_read: phase moments of the last views, in order."""
    import torch
    mesh = synth.kuhn_cube(16, seed=66)
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
        ctx.debug_set("timeline", 8)
        origin = torch.cuda.Event(enable_timing=True)
        origin.record()
        v = api.make_view(400, 300, X=0.5, Y=0.0, lib=gpu_lib)
        for _ in range(5):
            ctx.render(v)
        tl = ctx.timeline(origin.cuda_event)
        assert tl.shape == (5, 6)
        assert np.all(np.diff(tl, axis=1) >= 0) and np.all(np.diff(tl[:, 0]) > 0) and tl[0, 0] >= 0
        ctx.debug_set("timeline", 0)
        assert ctx.timeline(origin.cuda_event).shape == (0, 6)


def test_config_c4_sweep_every_30th_frame_against_reference(gpu_lib, ref, tmp_path):
    """configs[3] as specified: the 360-view turn of the 8M-tet grid at 1200 x 900, -X 0.4 -I -0.03
    -D 0.1 (utility/rotate_traces.py:8-9,18), rendered in one go through the pipelined calls
    (course5_b200.sweep.render_sweep: three views in flight); every 30th frame is held against
    oracle/_ref. The .vti values are doubles that went through float (plane.cpp:165-166): equal, or
    one float ulp apart on at most 1e-4 of the pixels; NaN masks and step totals identical."""
    import json
    import os
    import subprocess
    import sys
    from course5_b200 import sweep
    from parity import float_ulp_distance
    mesh, c4 = synth.make_config("C4")
    every = list(range(0, 360, 30))
    runner = os.path.join(os.path.dirname(rc.__file__), "ref_runner.py")
    proc = subprocess.Popen([sys.executable, runner, "C4", str(tmp_path / "c4.npz"),
                             json.dumps([dict(Y=2.0 * k / 360) for k in every])],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    kept, totals = {}, {}
    with api.Context(devices=(0,), lib=gpu_lib) as ctx:
        _upload_with_solids(ctx, mesh, c4["D"])
        views = sweep.sweep_views(c4["res_x"], c4["res_y"], X=c4["X"], I=c4["I"], alpha_limit=c4["alpha_limit"],
                                  frames=360, lib=gpu_lib)

        def consume(k, image, st):
            totals[k] = st["tet_steps"]
            if k in every:
                kept[k] = image.copy()

        steps, stats = sweep.render_sweep(ctx, views, in_flight=3, consume=consume)
    assert len(stats) == 360 and steps == sum(totals.values()) and all(s["walk_errors"] == 0 for s in stats)
    _, err = proc.communicate(timeout=1500)
    assert proc.returncode == 0, f"the reference did not finish: rc {proc.returncode}\n{err[-2000:]}"
    r = np.load(str(tmp_path / "c4.npz"))
    for n, k in enumerate(every):
        want = np.stack([r[f"tau{n}"], r[f"inten{n}"]], axis=-1).astype(np.float32).astype(np.float64)
        got = kept[k]
        assert np.array_equal(np.isnan(got), np.isnan(want)), f"frame {k}: NaN (solid) masks differ"
        assert totals[k] == int(r[f"total_steps{n}"]), f"frame {k}: tet-steps differ"
        ok = ~np.isnan(want)
        ulp = float_ulp_distance(got[ok], want[ok])
        assert ulp.max() <= 1 and (ulp > 0).mean() <= 1e-4, f"frame {k}: {int((ulp > 0).sum())} pixels differ, max {int(ulp.max())} ulp"
        assert np.array_equal(got[..., 0] != 0, want[..., 0] != 0), f"frame {k}: hit/miss sets differ"
