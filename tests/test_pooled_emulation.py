"""CPU check of the pooled traversal kernel's algorithm (traverse_pooled.cuh) without a GPU.

tools/pooled_emul.cpp emulates one warp of trace_pooled_kernel<0> lane by lane -- same cycle (speculative walk, leaf
queue, chunk dealing, survivor queue, exact rounds, tie rule), same arithmetic (fmaf pre-filter with the kernel's error
bounds E and F, unfused fp32 exact part) -- on the product's own kd layout, and compares every ray with the plain per-ray
traversal (the schedule the other kernels and the oracle's early-exit mode follow). What this pins on the CPU: the
division-free pre-filter never discards a triangle the exact test accepts, and the pooled cycle loses or reorders no hit.
"""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    obj = os.path.join(ROOT, "turner_b200", "csrc", "kdtree_build.o")
    if not os.path.exists(obj):
        subprocess.run(["make", "-C", os.path.dirname(obj), "kdtree_build.o"], check=True, capture_output=True)
    exe = str(tmp_path_factory.mktemp("emul") / "pooled_emul")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fopenmp", os.path.join(ROOT, "tools", "pooled_emul.cpp"), obj,
                    "-pthread", "-o", exe], check=True, capture_output=True)
    return exe


def _write_mesh(path, sc):
    V = np.ascontiguousarray(sc["vertices"], np.float32).reshape(-1, 9)
    N = np.ascontiguousarray(sc["normals"], np.float32).reshape(-1, 9)
    with open(path, "wb") as f:
        f.write(struct.pack("I", V.shape[0]))
        f.write(V.tobytes())
        f.write(N.tobytes())


@pytest.mark.parametrize("n", [24, 64])
def test_pooled_cycle_matches_per_ray_traversal(emulator, tmp_path, n):
    from turner_b200 import scenes
    mesh = str(tmp_path / "mesh.bin")
    _write_mesh(mesh, scenes.cubesphere(n))
    out = subprocess.run([emulator, mesh, "12"], check=True, capture_output=True, text=True, timeout=600).stdout
    lines = re.findall(r"depth (\d): (\d+) rays, (\d+) hits, completed (\d), id differences (\d+), \(r,s,t\) bit differences (\d+)", out)
    assert len(lines) == 3, out  # primary band, depth-1 and depth-2 child waves
    total = 0
    for depth, rays, hits, completed, id_diff, bit_diff in lines:
        assert completed == "1" and id_diff == "0" and bit_diff == "0", out
        assert int(hits) > 0
        total += int(rays)
    assert total > 20000
    # the shadow wave of every depth through the any-hit mode of the cycle (inclusive tmax, cells beyond the light pruned)
    shadows = re.findall(r"shadow (\d): (\d+) rays, (\d+) occluded, completed (\d), differences (\d+)", out)
    assert len(shadows) == 3, out
    for depth, rays, occluded, completed, diff in shadows:
        assert completed == "1" and diff == "0" and int(rays) > 0, out
    assert sum(int(s[2]) for s in shadows) > 0  # some rays are occluded


def test_pooled_cycle_on_a_device_built_tree(emulator, tmp_path):
    """The production tree of the benchmark is DEVICE-built (kdtree_build_gpu.cu): other shape, empty-space cuts everywhere.
    tests/golden/kd_gpu_cubesphere32.bin.gz is such a tree, written on a B200 by tools/jobs/dump_small.sh (TRN_KD_DUMP) for
    scenes.cubesphere(32). On the CPU this pins (a) that it is a correct tree -- the per-ray traversal finds what the
    exhaustive search over all 12 288 triangles finds, ties included -- and (b) the pooled cycle on it (closest hit and
    any-hit), void leaves and cut chains included."""
    import gzip
    from turner_b200 import scenes
    mesh, tree = str(tmp_path / "mesh.bin"), str(tmp_path / "tree.bin")
    _write_mesh(mesh, scenes.cubesphere(32))
    with gzip.open(os.path.join(ROOT, "tests", "golden", "kd_gpu_cubesphere32.bin.gz")) as f, open(tree, "wb") as g:
        g.write(f.read())
    out = subprocess.run([emulator, mesh, "12", tree, "1"], check=True, capture_output=True, text=True, timeout=600).stdout
    brute = re.findall(r"brute (\d): (\d+) rays, differences (\d+), exact ties resolved differently (\d+)", out)
    assert len(brute) == 3 and all(b[2] == "0" and b[3] == "0" for b in brute), out
    lines = re.findall(r"depth (\d): (\d+) rays, (\d+) hits, completed (\d), id differences (\d+), \(r,s,t\) bit differences (\d+)", out)
    assert len(lines) == 3 and all(l[3] == "1" and l[4] == "0" and l[5] == "0" and int(l[2]) > 0 for l in lines), out
    shadows = re.findall(r"shadow (\d): (\d+) rays, (\d+) occluded, completed (\d), differences (\d+)", out)
    assert len(shadows) == 3 and all(s[3] == "1" and s[4] == "0" for s in shadows), out
    # the walk really met cut-off voids (the host-built tree of this mesh has hardly any)
    voids = [float(v) for v in re.findall(r"waiting at a void ([0-9.]+)", out)]
    assert max(voids) > 0.5, out
