"""The CPU oracle against every known answer the reference's own tests hold for the hot path
(SURVEY.md section 4 / 8(c)). Citations: /root/reference/tests/*."""
import ctypes as C

import numpy as np
import pytest


def _scene(ob, verts, normals=None, diffuse=None):
    v = np.asarray(verts, np.float32).reshape(-1, 9)
    n = np.zeros_like(v) if normals is None else normals
    d = np.zeros((v.shape[0], 4), np.float32) if diffuse is None else diffuse
    return ob.OracleScene(v, n, d)


def test_four_separated_triangles(ob, scenes):
    # tests/test_kdtree.cpp:23-63
    sc = scenes.four_triangles()
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    assert o.height == 1 and o.num_nodes == 5
    org = np.zeros((4, 3), np.float32)
    d = np.array([[0.5, 0.5, 1], [2.5, 0.5, 1], [0.5, 2.5, 1], [2.5, 2.5, 1]], np.float32)
    for mode in (0, 1, 2):
        ids, rst = o.intersect(org, d, mode)
        assert ids.tolist() == [0, 1, 2, 3]
        assert rst[:, 0].tolist() == [1, 1, 1, 1]
        assert rst[0, 1:].tolist() == [0.5, 0.5]
        for k in (1, 2, 3):
            assert rst[k, 1:].tolist() == [0, 0.5]


def test_two_overlapping_triangles_height0(ob):
    # tests/test_kdtree.cpp:16-21
    o = _scene(ob, [[-1, -1, 0, 1, -1, 0, 1, 1, 0], [0, 0, 0, 1, 0, 0, 1, 2, 0]])
    assert o.height == 0


def test_cube_seven_nodes(ob, scenes):
    # tests/test_kdtree.cpp:162-186 (same 12 triangles, same order)
    t = [[-1, -1, -1, 1, -1, -1, 1, 1, -1], [-1, -1, -1, -1, 1, -1, 1, 1, -1], [-1, -1, 1, 1, -1, 1, 1, 1, 1],
         [-1, -1, 1, -1, 1, 1, 1, 1, 1], [-1, -1, -1, -1, -1, 1, -1, 1, 1], [-1, -1, -1, -1, 1, -1, -1, 1, 1],
         [1, -1, -1, 1, -1, 1, 1, 1, 1], [1, -1, -1, 1, 1, -1, 1, 1, 1], [-1, 1, -1, -1, 1, 1, 1, 1, 1],
         [-1, 1, -1, 1, 1, -1, 1, 1, 1], [-1, -1, -1, -1, -1, 1, 1, -1, 1], [-1, -1, -1, 1, -1, -1, 1, -1, 1]]
    o = _scene(ob, t)
    assert o.height == 0 and o.num_nodes == 7


def _regular_tri(ax, pos, phi):
    v = np.zeros((3, 3), np.float32)
    for i in range(3):
        v[i, ax] = pos
        v[i, (ax + 1) % 3] = np.cos(phi + 2.0 * np.pi / 3 * i)
        v[i, (ax + 2) % 3] = np.sin(phi + 2.0 * np.pi / 3 * i)
    return v.reshape(-1)


def test_coplanar_stack(ob):
    # tests/test_kdtree.cpp:211-245: answers do not depend on the random phase
    rng = np.random.RandomState(0)
    for ax in range(3):
        tris = [_regular_tri(ax, float(p), rng.uniform(0, 2 * np.pi)) for p in range(10)]
        o = _scene(ob, tris)
        for mode in (0, 1):
            org = np.zeros((1, 3), np.float32)
            d = np.zeros((1, 3), np.float32)
            org[0, ax], d[0, ax] = -100, 1
            ids, rst = o.intersect(org, d, mode)
            assert ids[0] == 0 and int(rst[0, 0]) == 100
            assert abs(rst[0, 1] - 1 / 3) < 1e-5 and abs(rst[0, 2] - 1 / 3) < 1e-5
            org[0, ax], d[0, ax] = 100, -1
            ids, rst = o.intersect(org, d, mode)
            assert ids[0] == 9 and int(rst[0, 0]) == 91
            assert abs(rst[0, 1] - 1 / 3) < 1e-5 and abs(rst[0, 2] - 1 / 3) < 1e-5


def test_all_triangles_in_one_plane(ob):
    # tests/test_kdtree.cpp:188-209
    rng = np.random.RandomState(3)
    for ax in range(3):
        tris = []
        for _ in range(1000):
            v = np.zeros((3, 3), np.float32)
            for i in range(3):
                phi = rng.uniform(0, 2 * np.pi)
                v[i, (ax + 1) % 3], v[i, (ax + 2) % 3] = np.cos(phi), np.sin(phi)
            tris.append(v.reshape(-1))
        o = _scene(ob, tris)
        phi = rng.uniform(0, 2 * np.pi)
        d = np.zeros((1, 3), np.float32)
        d[0, (ax + 1) % 3], d[0, (ax + 2) % 3] = np.cos(phi), np.sin(phi)
        ids, _ = o.intersect(np.ones((1, 3), np.float32), d)
        assert ids[0] == ob.MISS
        ids, _ = o.intersect(-np.ones((1, 3), np.float32), np.ones((1, 3), np.float32))
        assert ids[0] != ob.MISS


def test_ray_box(ob):
    # tests/test_intersection.cpp:15-36
    f = lambda *a: np.array(a, np.float32)
    tmin, tmax = C.c_float(), C.c_float()
    rb = lambda o, d, b: ob.lib().orc_ray_box(f(*o), f(*d), f(*b), C.byref(tmin), C.byref(tmax))
    unit = (-1, -1, -1, 1, 1, 1)
    assert rb((0, 0, 0), (1, 1, 1), unit)
    assert rb((10, 0, 0), (-1, 0, 0), unit)
    assert rb((0, 10, 0), (0, -1, 0), unit)
    assert rb((0, 0, 10), (0, 0, -1), unit)
    assert not rb((0, 0, 0), (1, 0, 0), (-1, -1, 1, 1, 1, 1))
    assert not rb((-2, -2, -2), (-1, 0, 0), (-1, -1, 1, 1, 1, 1))
    assert not rb((-1, 0, 0), (-1, 0, 0), (0, 0, 0, 1, 1, 1))


def test_flat_node_layout(ob):
    # tests/test_kdtree.cpp:261-337: [split f32][right:30][axis:2] / [id_a:32][id_b:30][11]
    o = _scene(ob, [[0, 0, 1, 0, 1, 1, 1, 0, 1], [2, 0, 1, 3, 0, 1, 3, 1, 1], [0, 2, 1, 0, 3, 1, 1, 3, 1],
                    [3, 2, 1, 3, 3, 1, 2, 3, 1]])
    nodes = o.nodes()
    root = int(nodes[0])
    assert root & 3 != 3  # inner
    right = (root & 0xFFFFFFFF) >> 2
    assert 0 < right < len(nodes)
    split = np.array([root >> 32], np.uint32).view(np.float32)[0]
    assert 0 <= split <= 3
    leaves = [int(n) for n in nodes if int(n) & 3 == 3]
    ids = []
    for n in leaves:
        ids.append(n >> 32)
        if n & 0xFFFFFFFF != 0xFFFFFFFF:
            ids.append((n & 0xFFFFFFFF) >> 2)
    assert sorted(ids) == [0, 1, 2, 3]
    assert 0 in [int(n) for n in nodes]  # an all-zero sentinel terminates the even leaf runs


def test_hemisphere_properties(ob):
    # tests/test_sampling.cpp:12-39
    n = 100
    out = np.zeros(4 * n * 100, np.float32)
    ob.lib().orc_hemisphere(n * 100, out)
    s = out.reshape(-1, 4)
    rng = np.random.RandomState(2)
    for i in range(n):
        nrm = rng.uniform(-10, 10, 3)
        nrm = (nrm / np.linalg.norm(nrm)).astype(np.float32)
        for j in range(100):
            loc = s[i * 100 + j]
            v = np.zeros(3, np.float32)
            ob.lib().orc_frame_apply(nrm, np.ascontiguousarray(loc[:3]), v)
            assert abs(np.linalg.norm(v) - 1) < 1e-3
            assert 0 <= loc[3] <= 1
            assert abs(float(v @ nrm) - loc[3]) < 0.01 * max(loc[3], 1e-2) + 1e-4


def test_sample_on_triangle_is_hit(ob):
    # tests/test_sampling.cpp:41-57: a point on the triangle is accepted by the ray/triangle test with r ~ |p|
    rng = np.random.RandomState(5)
    for _ in range(200):
        tri = rng.uniform(-10, 10, 9).astype(np.float32)
        r1, r2 = rng.uniform(0, 1, 2)
        if r1 + r2 > 1:
            continue
        p = tri[:3] + r1 * (tri[3:6] - tri[:3]) + r2 * (tri[6:] - tri[:3])
        rst = np.zeros(3, np.float32)
        hit = ob.lib().orc_ray_triangle(tri, np.zeros(3, np.float32), (p / np.linalg.norm(p)).astype(np.float32), rst)
        if min(r1, r2, 1 - r1 - r2) > 1e-3:
            assert hit and abs(rst[0] - np.linalg.norm(p)) < 1e-3 * np.linalg.norm(p)


def test_triangle_normal(ob):
    # tests/test_triangle.cpp:7-25: unit face normal, perpendicular to both edges
    rng = np.random.RandomState(1)
    for _ in range(50):
        tri = rng.uniform(-10, 10, 9).astype(np.float32)
        o = _scene(ob, [tri])
        f = o.triangle_fields(0)
        u, v, n = f[35:38], f[38:41], f[41:44]
        assert abs(np.linalg.norm(n) - 1) < 1e-5
        assert abs(float(n @ u)) < 1e-3 and abs(float(n @ v)) < 1e-3


def test_p3_golden_text(ob):
    # tests/test_raster.cpp:44-57 (+ the std::endl of main.cpp:242)
    img = np.zeros((2, 2, 4), np.float32)
    img[1, 0] = 1
    assert ob.write_p3(img) == "P3\n2 2\n255\n  0   0   0   0   0   0\n255 255 255   0   0   0\n"


def test_effects(ob):
    # tests/test_effects.cpp: exposure(.,0)==0, monotone; gamma brightens; alpha untouched
    x = np.array([[0.1, 0.5, 0.9, 0.3]], np.float32)
    assert np.all(ob.tonemap(x, 1, exposure=0.0, gamma_enabled=False)[0, :3] == 0)
    lo = ob.tonemap(x, 1, exposure=0.5, gamma_enabled=False)
    hi = ob.tonemap(x, 1, exposure=2.0, gamma_enabled=False)
    assert np.all(lo[0, :3] < hi[0, :3])
    g = ob.tonemap(x, 1, exposure=1.0, gamma_enabled=True)
    ng = ob.tonemap(x, 1, exposure=1.0, gamma_enabled=False)
    assert np.all(g[0, :3] > ng[0, :3])
    assert g[0, 3] == x[0, 3] and lo[0, 3] == x[0, 3]


def test_lambertian_four_channel_product(ob, scenes):
    # tests/test_lambertian.cpp:5-23 pins 4-channel colour products; the oracle's direct term is the same product:
    # one triangle facing a light straight above it, surface (1,0,0,1), light (.7,.5,.7,1), depth cut right after
    tri = np.array([[-5, -5, 0, 5, -5, 0, 0, 5, 0]], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (1, 3))
    dif = np.array([[1, 0, 0, 1]], np.float32)
    o = ob.OracleScene(tri, nrm, dif)
    sc = {"camera": scenes.look_at_camera((0, 0, 4), (0, 0, 0)), "light": {"pos": [0, 0, 1e6], "color": [.7, .5, .7, 1]}}
    cfg = ob.make_cfg(sc, 1, max_depth=1, mc_samples=1, pixel_samples=1)
    # single primary through the middle: direct = max(0, n.l) * light; children at depth 1 see only bg = 0
    img, _, st = o.render(cfg)
    px = img[0, 0]
    assert st.num_prim_rays == 1
    assert abs(px[0] - 0.7 / np.pi) < 1e-4 and px[1] == 0 and px[2] == 0


def test_readme_replay(ob, scenes):
    # README.md:22-36: cornell_box -w 320 --max-depth 3 -m 1 --pixel-samples 8 -> 2632399 rays, 819200 primary
    sc = scenes.fixture("cornell_box")
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    assert o.num_tris == 36 and o.height == 0  # README.md:30-31
    cfg = ob.make_cfg(sc, 320, max_depth=3, mc_samples=1, pixel_samples=8, num_threads=1)
    _, _, st = o.render(cfg)
    assert st.num_prim_rays == 819200
    assert st.num_rays == 2632399


def test_counter_seeded_mode_is_deterministic_and_thread_independent(ob, scenes):
    sc = scenes.fixture("cornell_box")
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    a, _, sa = o.render(ob.make_cfg(sc, 40, 2, 3, 2, num_threads=1, rng_mode=1, seed=9))
    b, _, sb = o.render(ob.make_cfg(sc, 40, 2, 3, 2, num_threads=4, rng_mode=1, seed=9))
    assert np.array_equal(a, b) and sa.num_rays == sb.num_rays
    c, _, _ = o.render(ob.make_cfg(sc, 40, 2, 3, 2, num_threads=1, rng_mode=1, seed=10))
    assert not np.array_equal(a, c)
