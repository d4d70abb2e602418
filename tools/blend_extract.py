#!/usr/bin/env python3
"""Fixture generator: read an (uncompressed, v27x, 64-bit LE) .blend file and dump
what the reference's loader path (main.cpp:25-82,96-136 over assimp's Blender
importer) would hand to the renderer: world-space triangles (vertex positions,
vertex normals, diffuse rgba), the camera node transform + horizontal FOV, and the
point light. Output is a small JSON fixture under tests/golden/.

    python tools/blend_extract.py /root/reference/scenes/cornell_box.blend tests/golden/cornell_box.json

This is tooling for fixtures (the scenes cannot travel to the GPU box); the
product's loader is turner_b200/csrc/blend_loader.cpp and is tested against these
fixtures. Loader conventions restated from assimp@a5a5343 BlenderLoader.cpp:
one face per MPoly (tri or quad), per-corner vertices, normals = MVert.no/32767
(ConvertMesh), quads split (0,1,2),(0,2,3) (TriangulateProcess, convex case),
one aiMesh per material slot in slot order, objects in Scene.base list order,
camera FOV = atan2(sensor_x, 2*lens) (ConvertCamera), light colour = rgb*energy.
"""
import json
import struct
import sys

import numpy as np


class Blend:
    def __init__(self, path):
        self.data = open(path, "rb").read()
        d = self.data
        assert d[:7] == b"BLENDER", "not an uncompressed .blend"
        assert d[7:8] == b"-" and d[8:9] == b"v", "need 64-bit little-endian"
        self.version = int(d[9:12])
        self.blocks = []  # (code, size, oldptr, sdna, count, offset)
        off = 12
        while off < len(d):
            code = d[off:off + 4]
            size, oldptr, sdna, count = struct.unpack_from("<iQii", d, off + 4)
            self.blocks.append((code, size, oldptr, sdna, count, off + 24))
            if code == b"ENDB":
                break
            off += 24 + size
        self.by_ptr = {b[2]: b for b in self.blocks if b[2]}
        self._parse_dna()

    def _parse_dna(self):
        blk = [b for b in self.blocks if b[0] == b"DNA1"][0]
        d, off = self.data, blk[5]
        assert d[off:off + 8] == b"SDNANAME"
        off += 8
        (n,) = struct.unpack_from("<i", d, off)
        off += 4
        self.names = []
        for _ in range(n):
            e = d.index(b"\0", off)
            self.names.append(d[off:e].decode())
            off = e + 1
        off = (off + 3) & ~3
        assert d[off:off + 4] == b"TYPE"
        off += 4
        (n,) = struct.unpack_from("<i", d, off)
        off += 4
        self.types = []
        for _ in range(n):
            e = d.index(b"\0", off)
            self.types.append(d[off:e].decode())
            off = e + 1
        off = (off + 3) & ~3
        assert d[off:off + 4] == b"TLEN"
        off += 4
        self.tlen = list(struct.unpack_from("<%dh" % len(self.types), d, off))
        off += 2 * len(self.types)
        off = (off + 3) & ~3
        assert d[off:off + 4] == b"STRC"
        off += 4
        (n,) = struct.unpack_from("<i", d, off)
        off += 4
        self.structs = []  # index -> (typename, {fieldname: (offset, typename, rawname)})
        self.struct_by_name = {}
        for _ in range(n):
            t, nf = struct.unpack_from("<hh", d, off)
            off += 4
            fields, fo = {}, 0
            for _ in range(nf):
                ft, fn = struct.unpack_from("<hh", d, off)
                off += 4
                raw = self.names[fn]
                base = raw.lstrip("*(").split("[")[0].rstrip(")")
                mult = 1
                for part in raw.split("[")[1:]:
                    mult *= int(part.rstrip("]"))
                size = 8 * mult if (raw.startswith("*") or raw.startswith("(*")) else self.tlen[ft] * mult
                fields[base] = (fo, self.types[ft], raw)
                fo += size
            self.structs.append((self.types[t], fields))
            self.struct_by_name[self.types[t]] = fields

    def blocks_of(self, code):
        return [b for b in self.blocks if b[0] == code]

    def field(self, blk, struct_name, name, fmt, index=0):
        """read field `name` of element `index` of a block holding `struct_name`s"""
        fo, _, _ = self.struct_by_name[struct_name][name]
        esize = self.tlen[self.types.index(struct_name)]
        return struct.unpack_from("<" + fmt, self.data, blk[5] + index * esize + fo)

    def struct_name_of(self, blk):
        return self.structs[blk[3]][0]


def extract(path):
    bf = Blend(path)
    d = bf.data
    # objects in Scene.base order
    scene = bf.blocks_of(b"SC\0\0")[0]
    base_fo, _, _ = bf.struct_by_name["Scene"]["base"]
    (first,) = struct.unpack_from("<Q", d, scene[5] + base_fo)  # ListBase.first
    objects = []
    ptr = first
    while ptr:
        b = bf.by_ptr[ptr]
        (nxt,) = bf.field(b, "Base", "next", "Q")
        (obp,) = bf.field(b, "Base", "object", "Q")
        objects.append(bf.by_ptr[obp])
        ptr = nxt

    verts, norms, cols, mirs, refls = [], [], [], [], []
    camera, light = None, None
    for ob in objects:
        (otype,) = bf.field(ob, "Object", "type", "h")
        (parent,) = bf.field(ob, "Object", "parent", "Q")
        assert parent == 0, "parented objects not supported"
        obmat = np.array(bf.field(ob, "Object", "obmat", "16f"), dtype=np.float32).reshape(4, 4)
        T = obmat.T.copy()  # blender stores column vectors as rows -> assimp row-major a1..d4
        (datap,) = bf.field(ob, "Object", "data", "Q")
        if otype == 11 and datap:  # camera
            ca = bf.by_ptr[datap]
            (lens,) = bf.field(ca, "Camera", "lens", "f")
            (sensor_x,) = bf.field(ca, "Camera", "sensor_x", "f")
            # evaluated in double, rounded once (reproduces README.md:22-36 ray counts exactly)
            hfov = float(np.float32(np.arctan2(float(np.float32(sensor_x)), float(np.float32(2.0) * np.float32(lens)))))
            camera = {"trafo4x4": [float(x) for x in T.reshape(-1)], "hfov": hfov,
                      "lens": float(lens), "sensor_x": float(sensor_x)}
        elif otype == 10 and datap:  # lamp
            la = bf.by_ptr[datap]
            r, g, b_ = bf.field(la, "Lamp", "r", "f") + bf.field(la, "Lamp", "g", "f") + bf.field(la, "Lamp", "b", "f")
            (energy,) = bf.field(la, "Lamp", "energy", "f")
            pos = T[:3, 3]
            e = np.float32(energy)
            light = {"pos": [float(x) for x in pos],
                     "color": [float(np.float32(r) * e), float(np.float32(g) * e), float(np.float32(b_) * e), 1.0]}
        elif otype == 1 and datap:  # mesh
            me = bf.by_ptr[datap]
            totvert, = bf.field(me, "Mesh", "totvert", "i")
            totpoly, = bf.field(me, "Mesh", "totpoly", "i")
            totcol, = bf.field(me, "Mesh", "totcol", "h")
            mvert = bf.by_ptr[bf.field(me, "Mesh", "mvert", "Q")[0]]
            mpoly = bf.by_ptr[bf.field(me, "Mesh", "mpoly", "Q")[0]]
            mloop = bf.by_ptr[bf.field(me, "Mesh", "mloop", "Q")[0]]
            matp = bf.field(me, "Mesh", "mat", "Q")[0]
            mats = []
            if matp and totcol:
                mb = bf.by_ptr[matp]
                for i in range(totcol):
                    (mp,) = struct.unpack_from("<Q", d, mb[5] + 8 * i)
                    ma = bf.by_ptr[mp]
                    rgb = [bf.field(ma, "Material", c, "f")[0] for c in ("r", "g", "b")]
                    # assimp omits AI_MATKEY_COLOR_DIFFUSE for an all-zero colour -> Get() leaves aiColor4D() = 0,0,0,0
                    dif = rgb + [1.0] if any(rgb) else [0.0, 0.0, 0.0, 0.0]
                    mir = [bf.field(ma, "Material", c, "f")[0] for c in ("mirr", "mirg", "mirb")] + [1.0]
                    (mode,) = bf.field(ma, "Material", "mode", "i")
                    (ray_mirror,) = bf.field(ma, "Material", "ray_mirror", "f")
                    refl = ray_mirror if (mode & 0x40000) else 0.0  # MA_RAYMIRROR -> AI_MATKEY_REFLECTIVITY
                    mats.append((dif, mir, refl))
            if not mats:
                mats = [([0.6, 0.6, 0.6, 1.0], [0.0, 0.0, 0.0, 0.0], 0.0)]  # assimp's default material
            co = [bf.field(mvert, "MVert", "co", "3f", i) for i in range(totvert)]
            no = [bf.field(mvert, "MVert", "no", "3h", i) for i in range(totvert)]
            polys = []
            for i in range(totpoly):
                ls, tl = bf.field(mpoly, "MPoly", "loopstart", "i", i) + bf.field(mpoly, "MPoly", "totloop", "i", i)
                (mat_nr,) = bf.field(mpoly, "MPoly", "mat_nr", "h", i)
                assert tl in (3, 4), "ngons not supported"
                vi = [bf.field(mloop, "MLoop", "v", "i", ls + j)[0] for j in range(tl)]
                polys.append((mat_nr, vi))
            R = T[:3, :3]
            t = T[:3, 3]

            def xf_point(p):
                p = np.array(p, dtype=np.float32)
                return [np.float32(np.float32(np.float32(R[k, 0] * p[0]) + np.float32(R[k, 1] * p[1])) +
                                   np.float32(R[k, 2] * p[2])) + t[k] for k in range(3)]

            def xf_normal(n):
                n = np.array(n, dtype=np.float32) / np.float32(32767.0)
                return [np.float32(np.float32(R[k, 0] * n[0]) + np.float32(R[k, 1] * n[1])) +
                        np.float32(R[k, 2] * n[2]) for k in range(3)]

            used = sorted(set(m for m, _ in polys))
            for slot in used:  # one aiMesh per used material slot
                col = mats[slot] if slot < len(mats) else mats[0]
                for m, vi in polys:
                    if m != slot:
                        continue
                    tris = [(0, 1, 2)] if len(vi) == 3 else [(0, 1, 2), (0, 2, 3)]
                    for a, b, c in tris:
                        verts.append([float(x) for k in (a, b, c) for x in xf_point(co[vi[k]])])
                        norms.append([float(x) for k in (a, b, c) for x in xf_normal(no[vi[k]])])
                        cols.append([float(np.float32(x)) for x in col[0]])
                        mirs.append([float(np.float32(x)) for x in col[1]])
                        refls.append(float(np.float32(col[2])))
    return {"source": path.split("/")[-1], "num_triangles": len(verts), "vertices": verts, "normals": norms,
            "diffuse": cols, "reflective": mirs, "reflectivity": refls, "camera": camera, "light": light}


if __name__ == "__main__":
    out = extract(sys.argv[1])
    with open(sys.argv[2], "w") as f:
        json.dump(out, f)
    print(sys.argv[2], out["num_triangles"], "triangles; camera", out["camera"] is not None, "light", out["light"])
