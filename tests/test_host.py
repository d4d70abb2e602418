"""Host-side logic of the product (no GPU needed): the C-ABI library loads and exports what include/turner_b200.h
declares, the host kd builder equals the oracle's tree node for node, camera / tone map / P3 writer / .blend loader
equal the oracle resp. the committed fixtures, and compute calls fail loudly without a device."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(api):
    header = open(os.path.join(ROOT, "include", "turner_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(trn_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 15
    L = C.CDLL(api.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(api.EXPORTS) == declared


def test_builder_equals_oracle_tree(api, ob, scenes):
    for sc in [scenes.four_triangles(), scenes.unit_cube(), scenes.fixture("cornell_box"), scenes.fixture("furnace_test"),
               scenes.fixture("colored_cube"), scenes.fixture("orthogonal_planes"), scenes.random_soup(100, 1),
               scenes.random_soup(3000, 4), scenes.random_soup(8000, 5, size=0.5), scenes.cubesphere(20)]:
        o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
        p = api.Scene.from_dict(sc)
        assert (p.num_nodes, p.height, p.num_triangles) == (o.num_nodes, o.height, o.num_tris), sc["name"]
        assert np.array_equal(p.nodes(), o.nodes()), sc["name"]
        assert np.array_equal(np.array(p.info.box, np.float32).view(np.uint32), o.box.view(np.uint32))


def test_builder_thread_count_does_not_change_the_tree(api, scenes, monkeypatch):
    sc = scenes.random_soup(20000, 6, size=0.6)
    monkeypatch.setenv("TRN_BUILD_THREADS", "1")
    a = api.Scene.from_dict(sc).nodes()
    monkeypatch.setenv("TRN_BUILD_THREADS", "7")
    b = api.Scene.from_dict(sc).nodes()
    assert np.array_equal(a, b)


def test_tree_golden_hashes(api, scenes):
    # golden: sha1 of the flattened node arrays (reference FlatNode encoding) of the shipped scenes, checked against
    # the reference's own builder when the fixture was made (tests/golden/tree_hashes.json)
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "tree_hashes.json")))
    for name, want in golden.items():
        p = api.Scene.from_dict(scenes.fixture(name))
        assert hashlib.sha1(p.nodes().tobytes()).hexdigest() == want["sha1"], name
        assert p.height == want["height"] and p.num_nodes == want["num_nodes"]


def test_camera_tonemap_p3_equal_oracle(api, ob, scenes):
    for name, width, aspect in [("cornell_box", 320, 1.0), ("furnace_test", 1024, 16 / 9), ("colored_cube", 77, 0.8)]:
        sc = scenes.fixture(name)
        cam, h = api.camera_setup(sc["camera"]["trafo4x4"], sc["camera"]["hfov"], aspect, width)
        pos, rot, dx, dy, oh = ob.camera_setup(sc["camera"]["trafo4x4"], sc["camera"]["hfov"], aspect, width)
        assert h == oh and list(cam.pos) == list(pos) and list(cam.rot) == list(rot)
        assert (cam.delta_x, cam.delta_y) == (np.float32(dx), np.float32(dy))
    rng = np.random.RandomState(0)
    img = rng.uniform(0, 3, (7, 5, 4)).astype(np.float32)
    for kw in [dict(), dict(exposure=0.3), dict(gamma_enabled=False), dict(inverse_gamma=0.7)]:
        a = api.tonemap(img, 4, **kw)
        b = ob.tonemap(img, 4, **kw)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert api.write_p3(a) == ob.write_p3(b)
    white = np.zeros((2, 2, 4), np.float32)
    white[1, 0] = 1
    assert api.write_p3(white) == "P3\n2 2\n255\n  0   0   0   0   0   0\n255 255 255   0   0   0\n"  # tests/test_raster.cpp:44-57


@pytest.mark.skipif(not os.path.isdir("/root/reference/scenes"), reason="reference scenes not present")
def test_blend_loader_equals_fixtures(api, scenes):
    for name in ("cornell_box", "furnace_test", "colored_cube", "orthogonal_planes"):
        a = api.load_blend("/root/reference/scenes/%s.blend" % name)
        b = scenes.fixture(name)
        for k in ("vertices", "normals", "diffuse", "reflective", "reflectivity"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), (name, k)
        assert np.array_equal(np.float32(a["camera"]["trafo4x4"]), np.float32(b["camera"]["trafo4x4"]))
        assert np.float32(a["camera"]["hfov"]) == np.float32(b["camera"]["hfov"])
        if b["light"]:
            assert np.array_equal(np.float32(a["light"]["pos"]), np.float32(b["light"]["pos"]))
            assert np.array_equal(np.float32(a["light"]["color"]), np.float32(b["light"]["color"]))
        else:
            assert a["light"] is None


def test_blend_loader_errors(api, tmp_path):
    with pytest.raises(api.TurnerError) as e:
        api.load_blend(str(tmp_path / "missing.blend"))
    assert e.value.code == -4
    bad = tmp_path / "bad.blend"
    bad.write_bytes(b"not a blend file at all")
    with pytest.raises(api.TurnerError):
        api.load_blend(str(bad))


def test_argument_validation(api, scenes):
    with pytest.raises(api.TurnerError) as e:
        api.Scene(np.zeros((0, 9), np.float32), np.zeros((0, 9), np.float32), np.zeros((0, 4), np.float32))
    assert e.value.code == -1
    bad = np.zeros((1, 9), np.float32)
    bad[0, 0] = np.nan
    with pytest.raises(api.TurnerError):
        api.Scene(bad, np.zeros((1, 9), np.float32), np.zeros((1, 4), np.float32))
    sc = scenes.fixture("cornell_box")
    p = api.Scene.from_dict(sc)
    for kw in [dict(pixel_samples=0), dict(max_depth=0), dict(mc_samples=0), dict(mc_samples=64, max_depth=8)]:
        cam, cfg = api.make_config(sc, 8, **kw)
        with pytest.raises(api.TurnerError) as e:
            p.render(cam, cfg)
        assert e.value.code in (-1, -5)


def test_no_cpu_fallback(api, scenes):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    sc = scenes.fixture("cornell_box")
    p = api.Scene.from_dict(sc)
    cam, cfg = api.make_config(sc, 8)
    for call in (lambda: p.render(cam, cfg), lambda: p.primary_hits(cam, cfg),
                 lambda: p.intersect(np.zeros((1, 3)), np.ones((1, 3)))):
        with pytest.raises(api.TurnerError) as e:
            call()
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_write_p3_nan_inf_negative_pixels(api):
    """a NaN channel used to index the 256-entry digit table out of range (SIGSEGV). NaN and negatives print 0,
    +inf and anything above 255 print 255 (clamp of lib/raster.h:91-96; int(NaN) is undefined in the reference)."""
    img = np.zeros((2, 3, 4), np.float32)
    img[..., 3] = 1.0
    img[0, 0, 0] = np.nan
    img[0, 1, 1] = np.inf
    img[0, 2, 2] = -np.inf
    img[1, 0] = (-3.0, 0.5, 7.0, 1.0)
    img[1, 1] = (0.25, np.nan, 1.0, np.nan)  # NaN alpha poisons every channel
    txt = api.write_p3(img)
    rows = txt.split("\n")
    assert rows[0] == "P3" and rows[1] == "3 2" and rows[2] == "255"
    assert rows[3] == "  0   0   0   0 255   0   0   0   0"
    assert rows[4] == "  0 127 255   0   0   0   0   0   0"


def test_tonemap_threads_do_not_change_the_result(api, monkeypatch):
    rng = np.random.default_rng(3)
    img = rng.random((512, 600, 4), dtype=np.float32) * 40
    monkeypatch.setenv("TRN_HOST_THREADS", "1")
    a = api.tonemap(img, 16, exposure=0.7)
    monkeypatch.setenv("TRN_HOST_THREADS", "7")
    b = api.tonemap(img, 16, exposure=0.7)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_cache_with_a_lying_header_is_refused_without_allocating(api, tmp_path):
    """the u64 triangle count of an untrusted kdtree.cache sized five vectors before any payload was read"""
    import struct
    p = tmp_path / "kdtree.cache"
    p.write_bytes(b"\x01" + struct.pack("<Q", (1 << 30) - 1))  # 9 bytes claiming 2^30-1 triangles
    with pytest.raises(api.TurnerError) as e:
        api.Scene.load_cache(str(p))
    assert e.value.code == -1
    p.write_bytes(b"\x01" + struct.pack("<Q", 5) + b"\0" * 100)  # count larger than the payload
    with pytest.raises(api.TurnerError):
        api.Scene.load_cache(str(p))


def test_tonemap_zero_shortcut_is_bit_identical_to_the_oracle(api, ob):
    rng = np.random.default_rng(11)
    img = rng.random((64, 80, 4), dtype=np.float32) * 8
    img[rng.random((64, 80)) < 0.5] = 0.0          # black background pixels
    img[3, 5] = (-0.0, 0.0, 1.0, 1.0)              # a negative zero is not shortcut
    for gamma in (True, False):
        for exposure in (1.0, 0.0, 2.5):
            a = api.tonemap(img, 4, exposure=exposure, gamma_enabled=gamma)
            b = ob.tonemap(img, 4, exposure=exposure, gamma_enabled=gamma)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (gamma, exposure)


def test_reference_shape_follows_from_the_pair_layout(api, ob, scenes, monkeypatch):
    """a device-built scene derives the reference's FlatNode view from the sibling-pair layout on demand
    (reference_shape_from_pairs); on a host-built tree that derivation must give back the builder's own array"""
    monkeypatch.setenv("TRN_REDERIVE_NODES", "1")
    for sc in [scenes.four_triangles(), scenes.unit_cube(), scenes.fixture("cornell_box"), scenes.fixture("furnace_test"),
               scenes.random_soup(3000, 4), scenes.cubesphere(20), scenes.tiled_box(8)]:
        o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
        p = api.Scene.from_dict(sc)
        assert (p.num_nodes, p.height) == (o.num_nodes, o.height), sc["name"]
        assert np.array_equal(p.nodes(), o.nodes()), sc["name"]


def test_device_builder_needs_a_device(api, scenes):
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    sc = scenes.unit_cube()
    with pytest.raises(api.TurnerError) as e:
        api.Scene.from_dict(sc, builder="gpu")
    assert e.value.code == -2
