"""Golden vectors produced by the reference's own code (tools/make_golden.py -> tests/golden/reference_outputs.npz):
the CPU oracle must reproduce every one bit for bit. Works without /root/reference and without oracle/_ref."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


BG = {"cornell_box": (0, 0, 0, 1), "colored_cube": (0.1, 0.2, 0.3, 1), "furnace_test": (1, 1, 1, 1)}


@pytest.mark.parametrize("name", sorted(BG))
def test_oracle_reproduces_reference_outputs(ob, scenes, gold, name):
    sc = scenes.fixture(name)
    o = ob.OracleScene(sc["vertices"], sc["normals"], sc["diffuse"], reflective=sc["reflective"], reflectivity=sc["reflectivity"])
    assert np.array_equal(o.nodes(), gold[name + "/nodes"]) and o.height == int(gold[name + "/height"][0])
    img, _, st = o.render(ob.make_cfg(sc, 24, 3, 2, 2, bg=BG[name]))
    assert np.array_equal(bits(img), bits(gold[name + "/pt_sum"]))
    assert [st.num_rays, st.num_prim_rays] == gold[name + "/pt_rays"].tolist()
    assert np.array_equal(bits(ob.tonemap(img, 2)), bits(gold[name + "/pt_final"]))
    cfg = ob.make_cfg(sc, 48, pixel_samples=1)
    assert np.array_equal(bits(ob.primary_dirs(cfg)), bits(gold[name + "/prim_dirs"]))
    assert np.array_equal(bits(np.array(list(cfg.cam_pos), np.float32)), bits(gold[name + "/cam_pos"]))
    dirs = gold[name + "/prim_dirs"].reshape(-1, 3)
    org = np.tile(gold[name + "/cam_pos"], (dirs.shape[0], 1))
    for mode in (0, 1):
        ids, rst = o.intersect(org, dirs, mode)
        assert np.array_equal(ids, gold[name + "/prim_ids"]) and np.array_equal(bits(rst), bits(gold[name + "/prim_rst"]))
    img, _, st = o.render(ob.make_cfg(sc, 48, integrator=1, bg=BG[name], max_visibility=2.0))
    assert np.array_equal(bits(img), bits(gold[name + "/rc_sum"])) and [st.num_rays, st.num_prim_rays] == gold[name + "/rc_rays"].tolist()
    assert ob.write_p3(ob.tonemap(img, 1)) == gold[name + "/rc_p3"].tobytes().decode()
    if sc["light"]:
        img, _, st = o.render(ob.make_cfg(sc, 48, max_depth=4, integrator=2, bg=BG[name], shadow_intensity=0.5, pixel_samples=2))
        assert np.array_equal(bits(img), bits(gold[name + "/rt_sum"])) and [st.num_rays, st.num_prim_rays] == gold[name + "/rt_rays"].tolist()


def test_rng_golden(ob, gold):
    x = np.zeros(64, np.uint64)
    ob.lib().orc_xorshift_u64(42, 64, x)
    assert np.array_equal(x, gold["xorshift64star_u64_seed42"])
    f = np.zeros(64, np.float32)
    ob.lib().orc_xorshift_float(4, 64, f)
    assert np.array_equal(bits(f), bits(gold["xorshift64star_float_seed4"]))
    h = np.zeros(4 * 64, np.float32)
    ob.lib().orc_hemisphere(64, h)
    assert np.array_equal(bits(h.reshape(64, 4)), bits(gold["hemisphere_first64"]))
