"""Pins the CPU oracle bit-for-bit against oracle/_ref: the reference's OWN kdtree.cpp / pathtracer.cpp /
raycaster.cpp compiled from /root/reference by oracle/build_ref.sh. Skipped when the .so is absent."""
import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def small_scenes(scenes):
    return [scenes.four_triangles(), scenes.unit_cube(), scenes.fixture("cornell_box"), scenes.fixture("furnace_test"),
            scenes.fixture("colored_cube"), scenes.fixture("orthogonal_planes"), scenes.random_soup(100, 1),
            scenes.random_soup(2000, 2), scenes.cubesphere(16)]


def test_tree_and_triangle_fields_identical(ob, scenes, ref_ok):
    for sc in small_scenes(scenes):
        o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
        r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"])
        assert (o.num_nodes, o.height, o.num_tris) == (r.num_nodes, r.height, r.num_tris), sc["name"]
        assert np.array_equal(o.nodes(), r.nodes()), sc["name"]
        assert np.array_equal(bits(o.box), bits(r.box))
        for i in range(min(o.num_tris, 64)):
            assert np.array_equal(bits(o.triangle_fields(i)), bits(r.triangle_fields(i))), (sc["name"], i)


def test_traversal_identical_and_equals_brute_force(ob, scenes, ref_ok):
    for sc in small_scenes(scenes):
        o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
        r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"])
        for inside in (False, True):
            ro, rd = scenes.random_rays(sc, 20000, seed=11, inside=inside)
            rd[::7, 0] = 0  # exact zeros in the direction exercise fix_direction (kdtree.cpp:503-511)
            rd[::11, 1] = 0
            i_ref, r_ref = r.intersect(ro, rd)
            for mode in (0, 1, 2):  # reference schedule, early-exit schedule, brute force
                i_o, r_o = o.intersect(ro, rd, mode)
                assert np.array_equal(i_o, i_ref), (sc["name"], inside, mode)
                assert np.array_equal(bits(r_o), bits(r_ref)), (sc["name"], inside, mode)


def test_rng_streams_identical(ob, ref_ok):
    for seed in (4, 42, 1, 2**63 + 12345):
        a = np.zeros(4096, np.uint64)
        b = np.zeros(4096, np.uint64)
        ob.lib().orc_xorshift_u64(seed, a.size, a)
        ob.ref_lib().ref_xorshift_u64(seed, b.size, b)
        assert np.array_equal(a, b)
        fa = np.zeros(4096, np.float32)
        fb = np.zeros(4096, np.float32)
        ob.lib().orc_xorshift_float(seed, fa.size, fa)
        ob.ref_lib().ref_xorshift_float(seed, fb.size, fb)
        assert np.array_equal(bits(fa), bits(fb)) and fa.min() >= 0 and fa.max() < 1
    ha = np.zeros(4 * 5000, np.float32)
    hb = np.zeros(4 * 5000, np.float32)
    ob.lib().orc_hemisphere(5000, ha)
    ob.ref_lib().ref_hemisphere(5000, hb)
    assert np.array_equal(bits(ha), bits(hb))


def test_primary_rays_identical(ob, scenes, ref_ok):
    for name, width, pps, aspect in [("cornell_box", 97, 3, 1.0), ("furnace_test", 64, 2, 1.5), ("colored_cube", 33, 1, 0.75)]:
        sc = scenes.fixture(name)
        cfg = ob.make_cfg(sc, width, pixel_samples=pps, aspect=aspect)
        pos, dirs = ob.ref_primary_dirs(ob.ref_camera(sc, aspect), width, pps)
        assert dirs.shape[0] == cfg.height
        assert np.array_equal(bits(pos), bits(np.array(list(cfg.cam_pos), np.float32)))
        assert np.array_equal(bits(dirs), bits(ob.primary_dirs(cfg)))


@pytest.mark.parametrize("name,width,depth,m,pps,bg", [
    ("cornell_box", 48, 3, 1, 2, (0, 0, 0, 1)),
    ("cornell_box", 40, 3, 4, 2, (0.1, 0.2, 0.3, 1)),
    ("cornell_box", 32, 1, 2, 1, (0, 0, 0, 1)),
    ("furnace_test", 40, 4, 3, 2, (1, 1, 1, 1)),
    ("colored_cube", 40, 2, 3, 2, (0.5, 0.5, 0.5, 1)),
    ("orthogonal_planes", 40, 3, 2, 1, (0, 0, 0, 1)),
])
def test_pathtracer_radiance_identical(ob, scenes, ref_ok, name, width, depth, m, pps, bg):
    sc = scenes.fixture(name)
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"])
    a, asq, st = o.render(ob.make_cfg(sc, width, depth, m, pps, bg=bg), want_sumsq=True)
    b, bsq, fin, rst = r.render(ob.ref_camera(sc), ob.ref_config(sc, width, depth, m, pps, bg=bg), True, True)
    assert (st.num_rays, st.num_prim_rays) == (rst.num_rays, rst.num_prim_rays)
    assert np.array_equal(bits(a), bits(b)) and np.array_equal(bits(asq), bits(bsq))
    tm = ob.tonemap(a, pps)
    assert np.array_equal(bits(tm), bits(fin))
    assert ob.write_p3(tm) == ob.ref_write_p3(fin)


def test_raycaster_identical(ob, scenes, ref_ok):
    for name in ("cornell_box", "colored_cube"):
        sc = scenes.fixture(name)
        o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
        r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind="raycaster")
        a, _, st = o.render(ob.make_cfg(sc, 96, integrator=1, max_visibility=2.0))
        b, _, _, rst = r.render(ob.ref_camera(sc), ob.ref_config(sc, 96, max_visibility=2.0))
        assert np.array_equal(bits(a), bits(b)) and st.num_rays == rst.num_rays


def test_prebuilt_tree_loads_through_reference_serialize_hook(ob, scenes, ref_ok):
    sc = scenes.cubesphere(12)
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"])
    r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], nodes=o.nodes(), box=o.box)
    assert r.height == o.height and r.num_nodes == o.num_nodes
    ro, rd = scenes.random_rays(sc, 5000, seed=3)
    assert np.array_equal(o.intersect(ro, rd)[0], r.intersect(ro, rd)[0])


@pytest.mark.parametrize("name,depth,shadow,bg", [
    ("cornell_box", 3, 0.5, (0, 0, 0, 1)), ("cornell_box", 1, 0.3, (0.2, 0.3, 0.4, 1)), ("cornell_box", 6, 1.0, (0, 0, 0, 1)),
    ("colored_cube", 3, 0.5, (0.1, 0.1, 0.1, 1)), ("orthogonal_planes", 2, 0.0, (0, 0, 0, 1)),
])
def test_raytracer_radiance_identical(ob, scenes, ref_ok, name, depth, shadow, bg):
    # raytracer.cpp:6-67 (Whitted integrator) against the reference's own TU
    sc = scenes.fixture(name)
    kw = dict(reflective=sc["reflective"], reflectivity=sc["reflectivity"])
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], **kw)
    r = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind="raytracer", **kw)
    a, _, st = o.render(ob.make_cfg(sc, 80, max_depth=depth, integrator=2, shadow_intensity=shadow, bg=bg, pixel_samples=2))
    b, _, fin, rst = r.render(ob.ref_camera(sc), ob.ref_config(sc, 80, max_depth=depth, shadow_intensity=shadow, bg=bg,
                                                               pixel_samples=2), want_final=True)
    assert np.array_equal(bits(a), bits(b)) and st.num_rays == rst.num_rays
    assert ob.write_p3(ob.tonemap(a, 2)) == ob.ref_write_p3(fin)
    # a mirror that is a perfect, unshadowed reflector of a mirror: every recursion level is exercised
    if name == "cornell_box":
        sc2 = dict(sc)
        sc2["reflectivity"] = np.full_like(sc["reflectivity"], 0.8)
        o2 = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], reflective=sc["reflective"], reflectivity=sc2["reflectivity"])
        r2 = ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], kind="raytracer", reflective=sc["reflective"],
                         reflectivity=sc2["reflectivity"])
        a2, _, s2 = o2.render(ob.make_cfg(sc, 48, max_depth=depth, integrator=2, shadow_intensity=shadow, bg=bg))
        b2, _, _, rs2 = r2.render(ob.ref_camera(sc), ob.ref_config(sc, 48, max_depth=depth, shadow_intensity=shadow, bg=bg))
        assert np.array_equal(bits(a2), bits(b2)) and s2.num_rays == rs2.num_rays and s2.num_rays > 48 * 48 * min(depth, 3)
