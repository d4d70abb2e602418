"""kdtree.cache (main.cpp:142-167): the product writes and reads the reference's cache layout.

Pinned here with the reference's own code: the file the product writes is read back by the reference's
KDTree::serialize() (oracle/_ref, compiled from /root/reference) and must give the reference the tree it builds itself;
the file the reference's serialize() writes is byte-identical to the product's and loads into the product. Only the
archive's primitive layer (cereal v1.2.2 PortableBinary: 1 flag byte, raw little-endian scalars, u64 size tags) is a
restatement (oracle/ref_shims/cereal/archives/portable_binary.hpp) -- cereal itself is not available here.
"""
import os
import struct

import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _ref_scene(ob, sc):
    """the reference's own build of the scene's triangles (mirror material included where the scene has one)"""
    if sc.get("reflective") is not None:
        return ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"], reflective=sc["reflective"], reflectivity=sc["reflectivity"])
    return ob.RefScene(sc["vertices"], sc["normals"], sc["diffuse"])


def _scenes(scenes):
    return [scenes.fixture("cornell_box"), scenes.fixture("colored_cube"), scenes.four_triangles(), scenes.random_soup(300, 4),
            scenes.cubesphere(8)]


def test_layout_of_the_written_file(api, scenes, tmp_path):
    sc = scenes.fixture("cornell_box")
    p = api.Scene.from_dict(sc)
    path = str(tmp_path / "kdtree.cache")
    p.save_cache(path)
    raw = open(path, "rb").read()
    n = p.num_triangles
    assert raw[0] == 1 and struct.unpack_from("<Q", raw, 1)[0] == n
    assert len(raw) == 1 + 8 + n * 192 + 24 + 8 + p.num_nodes * 8  # sizeof(Triangle) == 192 (component 6)
    t0 = np.frombuffer(raw, np.float32, 48, 9)
    assert np.array_equal(bits(t0[:9]), bits(np.asarray(sc["vertices"], np.float32).reshape(-1, 9)[0]))
    assert np.array_equal(bits(t0[22:26]), bits(np.asarray(sc["diffuse"], np.float32).reshape(-1, 4)[0]))
    assert struct.unpack_from("<Q", raw, 9 + n * 192 + 24)[0] == p.num_nodes
    nodes = np.frombuffer(raw, np.uint64, p.num_nodes, 9 + n * 192 + 24 + 8)
    assert np.array_equal(nodes, p.nodes())


def test_reference_reads_what_the_product_writes(api, ob, scenes, tmp_path):
    if not ob.ref_available():
        pytest.skip("oracle/_ref not built")
    for i, sc in enumerate(_scenes(scenes)):
        p = api.Scene.from_dict(sc)
        path = str(tmp_path / ("p%d.cache" % i))
        p.save_cache(path)
        r = ob.RefScene.from_cache(path)                      # main.cpp:147-152 with the reference's serialize()
        own = _ref_scene(ob, sc)                              # the reference's own build of the same triangles
        assert (r.num_nodes, r.height, r.num_tris) == (own.num_nodes, own.height, own.num_tris), sc["name"]
        assert np.array_equal(r.nodes(), own.nodes()) and np.array_equal(bits(r.box), bits(own.box))
        for t in range(0, r.num_tris, max(1, r.num_tris // 50)):
            assert np.array_equal(bits(r.triangle_fields(t)), bits(own.triangle_fields(t))), (sc["name"], t)
        ro, rd = scenes.random_rays(sc, 2000, seed=3)
        i0, r0 = own.intersect(ro, rd)
        i1, r1 = r.intersect(ro, rd)
        assert np.array_equal(i0, i1) and np.array_equal(bits(r0), bits(r1))


def test_product_reads_what_the_reference_writes(api, ob, scenes, tmp_path):
    if not ob.ref_available():
        pytest.skip("oracle/_ref not built")
    for i, sc in enumerate(_scenes(scenes)):
        own = _ref_scene(ob, sc)
        ref_path, our_path = str(tmp_path / ("r%d.cache" % i)), str(tmp_path / ("o%d.cache" % i))
        own.write_cache(ref_path)                             # main.cpp:158-165 with the reference's serialize()
        p = api.Scene.load_cache(ref_path)
        assert (p.num_nodes, p.height, p.num_triangles) == (own.num_nodes, own.height, own.num_tris), sc["name"]
        assert np.array_equal(p.nodes(), own.nodes())
        p.save_cache(our_path)
        assert open(our_path, "rb").read() == open(ref_path, "rb").read(), sc["name"]  # byte-identical files


def test_stale_or_damaged_cache_is_refused(api, scenes, tmp_path):
    # the reference renders whatever ./kdtree.cache holds; the product refuses a tree that does not belong
    # to the triangles next to it
    p = api.Scene.from_dict(scenes.fixture("cornell_box"))
    path = str(tmp_path / "kdtree.cache")
    p.save_cache(path)
    raw = bytearray(open(path, "rb").read())
    assert api.Scene.load_cache(path).num_nodes == p.num_nodes
    stale = bytearray(raw)
    stale[9:13] = struct.pack("<f", 123.0)  # move a vertex: the cached tree no longer belongs to the triangles
    open(path, "wb").write(stale)
    with pytest.raises(api.TurnerError) as e:
        api.Scene.load_cache(path)
    assert "stale or foreign" in str(e.value)
    open(path, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(api.TurnerError):
        api.Scene.load_cache(path)
    open(path, "wb").write(b"\x00" + raw[1:])
    with pytest.raises(api.TurnerError):
        api.Scene.load_cache(path)
    with pytest.raises(api.TurnerError):
        api.Scene.load_cache(str(tmp_path / "missing.cache"))
