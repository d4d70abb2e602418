"""The `pathtracer` / `raycaster` executables: the reference's command line (pathtracer.h, raycaster.h, config.h via
docopt), stderr report (main.cpp:97,139,168, lib/progress_bar.h, lib/output.h:101-113) and P3 stdout (lib/raster.h)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PT = os.path.join(ROOT, "turner_b200", "pathtracer")
RC = os.path.join(ROOT, "turner_b200", "raycaster")
RT = os.path.join(ROOT, "turner_b200", "raytracer")


def run(exe, *args):
    return subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=600)


@pytest.fixture(scope="module")
def soup(api, scenes, tmp_path_factory):
    p = tmp_path_factory.mktemp("scenes") / "cornell.soup"
    api.save_soup(scenes.fixture("cornell_box"), str(p))
    return str(p)


def test_soup_round_trip(api, scenes, soup):
    a, b = scenes.fixture("cornell_box"), api.load_blend(soup)
    for k in ("vertices", "normals", "diffuse"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
    assert np.float32(a["camera"]["hfov"]) == np.float32(b["camera"]["hfov"])


def test_flag_grammar_of_the_reference_usage(api):
    # tests/test_config.cpp:21-43,76-91 (docopt grammar): long/short, '=', glued short values, options around <filename>
    r = run(PT, "-w", "100", "-a", "0.5", "--background", "0.5 0.5 0.5", "-t", "2", "--inverse-gamma", "1",
            "--no-gamma-correction", "--exposure", "1.5", "missing_file", "-v")
    assert r.returncode == 1  # import failure: message on stdout, exit 1 (main.cpp:104-107)
    assert "cannot open missing_file" in r.stdout
    e = r.stderr
    assert "Filename: missing_file" in e and "Aspect ratio: 0.5" in e and "Image width: 100" in e
    assert "Number of threads: 2" in e and "Inverse gamma: 1" in e and "Exposure: 1.5" in e
    assert "Gamma correction enabled: 0" in e and e.rstrip().endswith("Loading scene...")
    r = run(PT, "-d", "42", "-p", "42", "-m", "42", "file", "-v")
    assert "Max recursion depth: 42" in r.stderr and "Number of pixel samples: 42" in r.stderr
    assert "Number of Monte-Carlo samples: 42" in r.stderr
    r = run(PT, "file", "-p1", "-m1", "--width=33", "--max-depth=5", "-v")  # scripts/render-samples.sh:21 style
    assert "Number of pixel samples: 1" in r.stderr and "Number of Monte-Carlo samples: 1" in r.stderr
    assert "Image width: 33" in r.stderr and "Max recursion depth: 5" in r.stderr
    r = run(PT, "file", "-v")  # defaults come from the USAGE text (pathtracer.h:3-25)
    assert "Image width: 640" in r.stderr and "Number of Monte-Carlo samples: 8" in r.stderr
    assert "Max recursion depth: 3" in r.stderr and "Inverse gamma: 0.454545" in r.stderr
    r = run(RC, "--max-visibility", "4.5", "file", "-v")
    assert "Max visibility: 4.5" in r.stderr
    assert run(RC, "-d", "3", "file").returncode != 0          # raycaster USAGE has no -d
    assert run(PT, "--max-visibility", "2", "file").returncode != 0
    assert run(PT).returncode != 0 and "Usage: pathtracer <filename> [options]" in run(PT).stderr
    assert run(PT, "--help").returncode == 0 and "--monte-carlo-samples" in run(PT, "--help").stdout
    assert run(PT, "file", "--bogus").returncode != 0
    assert run(PT, "file", "-p", "0").returncode == 2          # the reference asserts (config.h:124); here exit code 2
    r = run(RT, "-d", "5", "--shadow", "0.25", "file", "-v")   # tests/test_config.cpp raytracer case
    assert "Max recursion depth: 5" in r.stderr and "Shadow intensity: 0.25" in r.stderr
    assert run(RT, "file", "-p", "2").returncode != 0          # raytracer USAGE has no -p
    assert run(RT, "file", "--shadow", "1.5").returncode == 2


def test_without_gpu_the_cli_fails_loudly(api, soup):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    r = run(PT, soup, "-w", "16")
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
    assert "Loading triangles and building kd-tree..." in r.stderr and "KDTree runtime: " in r.stderr
    # lib/progress_bar.h / tests/test_progress_bar.cpp format (the leading \\r arrives as a newline in text mode)
    assert "Rendering           " + "□" * 20 + "   0.00%" in r.stderr


@pytest.mark.gpu
def test_pathtracer_end_to_end_matches_oracle(api, ob, scenes, soup, tmp_path):
    W, D, M, P = 64, 3, 2, 4
    lin = tmp_path / "lin.f32"
    hits = tmp_path / "hits.u32"
    r = run(PT, soup, "-w", str(W), "-d", str(D), "-m", str(M), "-p", str(P), "--seed", "7", "--dump-linear", str(lin),
            "--dump-hits", str(hits))
    assert r.returncode == 0, r.stderr
    e = r.stderr
    assert "Triangles      : 36" in e and "Kd-Tree Height : 0" in e  # README.md:30-31
    assert "Rays (primary) : %d" % (W * W * P) in e
    assert re.search(r"Rays/sec       : \d+", e) and re.search(r"Rendering time : [\d.e-]+ sec", e)
    assert "■" * 20 + " 100.00%" in e
    head, body = r.stdout.split("\n255\n", 1)
    assert head == "P3\n%d %d" % (W, W)
    px = np.array(body.split(), dtype=np.int64).reshape(W, W, 3)
    sc = scenes.fixture("cornell_box")
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    ocfg = ob.make_cfg(sc, W, D, M, P, rng_mode=1, seed=7, num_threads=4)
    ref, _, ost = o.render(ocfg)
    assert "Rays           : %d" % ost.num_rays in e
    want = np.array(ob.write_p3(ob.tonemap(ref, P)).split("\n255\n", 1)[1].split(), dtype=np.int64).reshape(W, W, 3)
    assert (np.abs(px - want) <= 1).mean() > 0.999 and (px == want).mean() > 0.98
    linear = np.fromfile(lin, np.float32).reshape(W, W, 4)
    assert np.allclose(linear, ref, rtol=2e-4, atol=2e-4)
    ids = np.fromfile(hits, np.uint32)
    dirs = ob.primary_dirs(ocfg).reshape(-1, 3)
    oi, _ = o.intersect(np.tile(np.array(list(ocfg.cam_pos), np.float32), (dirs.shape[0], 1)), dirs, 0)
    assert np.array_equal(ids, oi)


@pytest.mark.gpu
def test_raycaster_end_to_end_exact(api, ob, scenes, soup):
    W = 96
    r = run(RC, soup, "-w", str(W), "--max-visibility", "1.5", "--background", "0.25")
    assert r.returncode == 0, r.stderr
    sc = scenes.fixture("cornell_box")
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    ref, _, ost = o.render(ob.make_cfg(sc, W, integrator=1, max_visibility=1.5, bg=(0.25, 0.25, 0.25, 1), num_threads=4))
    assert r.stdout == ob.write_p3(ob.tonemap(ref, 1))  # one sample per pixel: the P3 text is byte-identical
    assert "Rays           : %d" % ost.num_rays in r.stderr


@pytest.mark.gpu
def test_raytracer_end_to_end_matches_oracle(api, ob, scenes, soup):
    W = 96
    r = run(RT, soup, "-w", str(W), "-d", "4", "--shadow", "0.4")
    assert r.returncode == 0, r.stderr
    sc = scenes.fixture("cornell_box")
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], reflective=sc["reflective"], reflectivity=sc["reflectivity"])
    ref, _, ost = o.render(ob.make_cfg(sc, W, max_depth=4, integrator=2, shadow_intensity=0.4, num_threads=4))
    assert "Rays           : %d" % ost.num_rays in r.stderr
    px = np.array(r.stdout.split("\n255\n", 1)[1].split(), dtype=np.int64)
    want = np.array(ob.write_p3(ob.tonemap(ref, 1)).split("\n255\n", 1)[1].split(), dtype=np.int64)
    assert (np.abs(px - want) <= 1).all() and (px == want).mean() > 0.999


def test_kdtree_cache_flag(api, scenes, soup, tmp_path):
    # main.cpp:142-167 through the CLI: first run builds and writes the named cache, later runs load it; a cache whose
    # tree does not belong to its triangles is refused (the reference would render it). The scene set-up
    # happens before the device is touched, so this runs without a GPU too (the run then stops with exit code 3).
    cache = str(tmp_path / "kdtree.cache")
    r = run(PT, soup, "-w", "16", "--kdtree-cache", cache)
    assert r.returncode in (0, 3) and os.path.exists(cache), r.stderr
    first = open(cache, "rb").read()
    p = api.Scene.load_cache(cache)
    assert p.num_triangles == 36 and p.height == 0
    r = run(PT, soup, "-w", "16", "--kdtree-cache=" + cache)
    assert r.returncode in (0, 3) and "Triangles      : 36" in r.stderr or r.returncode == 3
    assert open(cache, "rb").read() == first  # loaded, not rewritten
    stale = bytearray(first)
    stale[9:13] = np.float32(42.0).tobytes()
    open(cache, "wb").write(stale)
    r = run(PT, soup, "-w", "16", "--kdtree-cache", cache)
    assert r.returncode == 2 and "stale or foreign" in r.stderr


@pytest.mark.gpu
def test_render_from_kdtree_cache_is_identical(api, soup, tmp_path):
    cache = str(tmp_path / "kdtree.cache")
    a = run(RC, soup, "-w", "64", "--kdtree-cache", cache)
    b = run(RC, soup, "-w", "64", "--kdtree-cache", cache)
    c = run(RC, soup, "-w", "64")
    assert a.returncode == b.returncode == c.returncode == 0, a.stderr
    assert a.stdout == b.stdout == c.stdout


@pytest.mark.gpu
def test_device_built_tree_and_multi_gpu_flags(api, soup, tmp_path):
    # --kd-builder=gpu: same primary hits, byte-identical raycaster image (one sample per pixel: no summation order involved);
    # --gpus 2 (where two are present): the sample split + ncclReduce of trn_render_multi gives the single-GPU image
    a = run(RC, soup, "-w", "96")
    b = run(RC, soup, "-w", "96", "--kd-builder=gpu")
    assert a.returncode == b.returncode == 0, b.stderr
    assert a.stdout == b.stdout
    assert "Kd-Tree Height : 0" in a.stderr and "Kd-Tree Height : " in b.stderr
    ha, hb = tmp_path / "a.u32", tmp_path / "b.u32"
    la, lb = tmp_path / "a.f32", tmp_path / "b.f32"
    p1 = run(PT, soup, "-w", "48", "-d", "3", "-m", "2", "-p", "4", "--dump-hits", str(ha), "--dump-linear", str(la))
    p2 = run(PT, soup, "-w", "48", "-d", "3", "-m", "2", "-p", "4", "--kd-builder", "gpu", "--dump-hits", str(hb), "--dump-linear", str(lb))
    assert p1.returncode == p2.returncode == 0, p2.stderr
    assert np.array_equal(np.fromfile(ha, np.uint32), np.fromfile(hb, np.uint32))
    assert np.allclose(np.fromfile(la, np.float32), np.fromfile(lb, np.float32), rtol=2e-4, atol=2e-4)
    assert run(PT, soup, "--kd-builder", "fpga").returncode == 2
    if api.device_count() >= 2:
        p3 = run(PT, soup, "-w", "48", "-d", "3", "-m", "2", "-p", "4", "--gpus", "2", "--dump-linear", str(lb))
        assert p3.returncode == 0, p3.stderr
        assert np.allclose(np.fromfile(la, np.float32), np.fromfile(lb, np.float32), rtol=2e-4, atol=2e-4)
