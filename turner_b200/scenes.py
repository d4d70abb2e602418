"""Deterministic synthetic scenes for tests and benchmarks (host side, numpy only).

A scene is a dict:
  vertices [N,9] f32   world-space v0,v1,v2 per triangle (what main.cpp:53-55 produces)
  normals  [N,9] f32   per-vertex normals, un-normalised (main.cpp:57-59)
  diffuse  [N,4] f32   rgba of AI_MATKEY_COLOR_DIFFUSE (main.cpp:42)
  reflective [N,4], reflectivity [N]  optional: AI_MATKEY_COLOR_REFLECTIVE / AI_MATKEY_REFLECTIVITY (main.cpp:44-47;
                       only the raytracer integrator reads them)
  camera   {'trafo4x4': 16 floats row-major (assimp a1..d4), 'hfov': radians}
  light    {'pos': [3], 'color': [4]} or None

Shapes follow SURVEY.md section 8(d): S-cornell / S-furnace come from the fixtures
extracted from the reference's .blend files (tools/blend_extract.py), S-mesh1M is the
displaced cube-sphere (6 faces x n^2 quads x 2, n=288 -> 995,328 triangles).
"""
import json
import math
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_json(path):
    with open(path) as f:
        d = json.load(f)
    return {
        "name": d.get("source", os.path.basename(path)),
        "vertices": np.asarray(d["vertices"], dtype=np.float32).reshape(-1, 9),
        "normals": np.asarray(d["normals"], dtype=np.float32).reshape(-1, 9),
        "diffuse": np.asarray(d["diffuse"], dtype=np.float32).reshape(-1, 4),
        "reflective": np.asarray(d.get("reflective", np.zeros((len(d["diffuse"]), 4))), dtype=np.float32).reshape(-1, 4),
        "reflectivity": np.asarray(d.get("reflectivity", np.zeros(len(d["diffuse"]))), dtype=np.float32).reshape(-1),
        "camera": d["camera"],
        "light": d.get("light"),
    }


def fixture(name):
    """'cornell_box' | 'furnace_test' | 'colored_cube' | 'orthogonal_planes'"""
    return load_json(os.path.join(_GOLDEN, name + ".json"))


def look_at_camera(eye, target, up=(0.0, 1.0, 0.0), lens=35.0, sensor_x=32.0):
    """node transform of a camera that looks down its local -z (assimp convention, lib/types.h:96-98)"""
    eye = np.asarray(eye, np.float64)
    f = np.asarray(target, np.float64) - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, np.asarray(up, np.float64))
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r, u, -f, eye
    hfov = float(np.float32(np.arctan2(float(np.float32(sensor_x)), float(np.float32(2.0) * np.float32(lens)))))
    return {"trafo4x4": [float(np.float32(x)) for x in m.reshape(-1)], "hfov": hfov}


def cubesphere(n=288, rho=(0.7, 0.7, 0.7)):
    """S-mesh1M: displaced cube-sphere, 12*n^2 triangles, smooth vertex normals, no degenerate faces."""
    t = np.linspace(-1.0, 1.0, n + 1)
    a, b = np.meshgrid(t, t, indexing="ij")
    one = np.ones_like(a)
    faces = [(one, a, b), (-one, b, a), (b, one, a), (a, -one, b), (a, b, one), (b, a, -one)]
    V, Nn = [], []
    for fx, fy, fz in faces:
        c = np.stack([fx, fy, fz], -1)
        dirn = c / np.linalg.norm(c, axis=-1, keepdims=True)
        phi = np.arctan2(dirn[..., 1], dirn[..., 0])
        theta = np.arccos(np.clip(dirn[..., 2], -1, 1))
        rad = 1 + 0.05 * np.sin(8 * phi) * np.sin(6 * theta) + 0.01 * np.sin(40 * phi + 3) * np.sin(32 * theta)
        P = dirn * rad[..., None]
        di = np.gradient(P, axis=0)
        dj = np.gradient(P, axis=1)
        nrm = np.cross(di, dj)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        flip = np.sum(nrm * dirn, -1) < 0
        nrm[flip] *= -1
        p00, p10, p11, p01 = P[:-1, :-1], P[1:, :-1], P[1:, 1:], P[:-1, 1:]
        n00, n10, n11, n01 = nrm[:-1, :-1], nrm[1:, :-1], nrm[1:, 1:], nrm[:-1, 1:]
        # orient so the geometric normal points outward
        t1 = np.stack([p00, p10, p11], -2)
        t2 = np.stack([p00, p11, p01], -2)
        m1 = np.stack([n00, n10, n11], -2)
        m2 = np.stack([n00, n11, n01], -2)
        tris = np.stack([t1, t2], 2).reshape(-1, 3, 3)
        nors = np.stack([m1, m2], 2).reshape(-1, 3, 3)
        g = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0])
        inward = np.sum(g * tris[:, 0], -1) < 0
        tris[inward] = tris[inward][:, [0, 2, 1]]
        nors[inward] = nors[inward][:, [0, 2, 1]]
        V.append(tris)
        Nn.append(nors)
    V = np.concatenate(V).astype(np.float32).reshape(-1, 9)
    Nn = np.concatenate(Nn).astype(np.float32).reshape(-1, 9)
    D = np.tile(np.asarray(list(rho) + [1.0], np.float32), (V.shape[0], 1))
    lo, hi = V.reshape(-1, 3).min(0), V.reshape(-1, 3).max(0)
    diag = float(np.linalg.norm(hi - lo))
    eye = np.array([0.35, 0.25, 1.0])
    eye = eye / np.linalg.norm(eye) * 1.3 * diag
    return {
        "name": "cubesphere_n%d" % n,
        "vertices": V, "normals": Nn, "diffuse": D,
        "camera": look_at_camera(eye, (0, 0, 0)),
        "light": {"pos": [3.0, 4.0, 5.0], "color": [1.0, 1.0, 1.0, 1.0]},
    }


def random_soup(n, seed=0, extent=10.0, size=2.0):
    """n random triangles (the shape of tests/test_kdtree.cpp:97-160's stress input)"""
    rng = np.random.RandomState(seed)
    c = rng.uniform(-extent, extent, (n, 1, 3))
    V = (c + rng.uniform(-size, size, (n, 3, 3))).astype(np.float32)
    g = np.cross(V[:, 1] - V[:, 0], V[:, 2] - V[:, 0])
    g /= np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-20)
    Nn = np.repeat(g[:, None, :], 3, 1).astype(np.float32)
    D = np.concatenate([rng.uniform(0.1, 0.9, (n, 3)), np.ones((n, 1))], 1).astype(np.float32)
    return {
        "name": "soup_%d_%d" % (n, seed),
        "vertices": V.reshape(-1, 9), "normals": Nn.reshape(-1, 9), "diffuse": D,
        "camera": look_at_camera((0.0, 0.0, 3.2 * extent), (0, 0, 0)),
        "light": {"pos": [0.0, 2.5 * extent, 0.0], "color": [1.0, 1.0, 1.0, 1.0]},
    }


def four_triangles():
    """tests/test_kdtree.cpp:23-28"""
    V = np.array([[0, 0, 1, 0, 1, 1, 1, 0, 1], [2, 0, 1, 3, 0, 1, 3, 1, 1], [0, 2, 1, 0, 3, 1, 1, 3, 1],
                  [3, 2, 1, 3, 3, 1, 2, 3, 1]], np.float32)
    return {"name": "four_triangles", "vertices": V, "normals": np.zeros((4, 9), np.float32),
            "diffuse": np.zeros((4, 4), np.float32), "camera": look_at_camera((1.5, 1.5, -4), (1.5, 1.5, 1)),
            "light": None}


def unit_cube():
    """12-triangle cube, tests/test_kdtree.cpp:162-186"""
    q = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    f = [(0, 1, 2), (0, 2, 3), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (2, 3, 7), (2, 7, 6), (0, 3, 7), (0, 7, 4),
         (1, 2, 6), (1, 6, 5)]
    V = np.array([[c for k in tri for c in q[k]] for tri in f], np.float32)
    return {"name": "unit_cube", "vertices": V, "normals": np.zeros((12, 9), np.float32),
            "diffuse": np.full((12, 4), 0.5, np.float32), "camera": look_at_camera((3, 2, 5), (0, 0, 0)),
            "light": {"pos": [2.0, 3.0, 4.0], "color": [1.0, 1.0, 1.0, 1.0]}}


def random_rays(scene, n, seed=0, inside=False):
    """rays aimed at the scene: from a shell outside the box (primary-like) or from inside it (incoherent)"""
    rng = np.random.RandomState(seed)
    P = scene["vertices"].reshape(-1, 3)
    lo, hi = P.min(0), P.max(0)
    ctr, ext = (lo + hi) / 2, (hi - lo) / 2 + 1e-3
    if inside:
        o = ctr + rng.uniform(-1, 1, (n, 3)) * ext
        d = rng.normal(size=(n, 3))
    else:
        u = rng.normal(size=(n, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        o = ctr + u * np.linalg.norm(ext) * 2.5
        tgt = ctr + rng.uniform(-1, 1, (n, 3)) * ext
        d = tgt - o
    return o.astype(np.float32), d.astype(np.float32)


def math_pi():
    return math.pi


def tiled_box(n=8, size=1.0):
    """closed axis-aligned box whose six walls are n x n quads (12 n^2 triangles): every triangle is planar in one axis
    and its edges coincide with candidate split planes -- the adversarial case for anything that reasons about cells"""
    t = np.linspace(0.0, size, n + 1)
    V = []
    for ax in range(3):
        for side in (0.0, size):
            for i in range(n):
                for j in range(n):
                    q = []
                    for (a, b) in ((t[i], t[j]), (t[i + 1], t[j]), (t[i + 1], t[j + 1]), (t[i], t[j + 1])):
                        p = [0.0, 0.0, 0.0]
                        p[ax] = side
                        p[(ax + 1) % 3] = a
                        p[(ax + 2) % 3] = b
                        q.append(p)
                    V.append(q[0] + q[1] + q[2])
                    V.append(q[0] + q[2] + q[3])
    V = np.asarray(V, np.float32)
    N = np.zeros_like(V)
    D = np.tile(np.asarray([0.7, 0.7, 0.7, 1.0], np.float32), (V.shape[0], 1))
    return {"name": "tiled_box_%d" % n, "vertices": V, "normals": N, "diffuse": D,
            "camera": look_at_camera((size / 2, size / 2, size * 0.9), (size / 2, size / 2, 0)),
            "light": {"pos": [size / 2, size * 0.9, size / 2], "color": [1.0, 1.0, 1.0, 1.0]}}
