"""Generates leisure_software_renderer_b200/assets/suzanne.npz from the reference's Suzanne fixture.

Run HERE (the container that has /root/reference); the GPU box only sees the generated file.
The reference loads cpp-folders/src/assets/obj/monkey/monkey.rawobj through Assimp
(resources/loaders/mesh_loader_assimp.hpp:42-101).  Assimp is not available, and the raster path
only observes triangle order and per-corner attributes, so this reader keeps the file's triangle
order and builds one vertex per distinct (v, vt, vn) triple.
"""
import os
import sys

import numpy as np

SRC = "/root/reference/cpp-folders/src/assets/obj/monkey/monkey.rawobj"
DST = os.path.join(os.path.dirname(__file__), "..", "..", "leisure_software_renderer_b200", "assets", "suzanne.npz")


def main():
    v, vt, vn, tris = [], [], [], []
    for line in open(SRC):
        p = line.split()
        if not p:
            continue
        if p[0] == "v":
            v.append([float(x) for x in p[1:4]])
        elif p[0] == "vt":
            vt.append([float(x) for x in p[1:3]])
        elif p[0] == "vn":
            vn.append([float(x) for x in p[1:4]])
        elif p[0] == "f":
            assert len(p) == 4, "triangles only"
            tris.append([tuple(int(i) for i in c.split("/")) for c in p[1:4]])
    remap, pos, nrm, uv, idx = {}, [], [], [], []
    for tri in tris:
        for key in tri:
            if key not in remap:
                remap[key] = len(pos)
                pos.append(v[key[0] - 1]); uv.append(vt[key[1] - 1]); nrm.append(vn[key[2] - 1])
            idx.append(remap[key])
    np.savez_compressed(DST, positions=np.asarray(pos, np.float32), normals=np.asarray(nrm, np.float32),
                        uvs=np.asarray(uv, np.float32), indices=np.asarray(idx, np.uint32))
    print(f"{len(tris)} triangles, {len(pos)} vertices -> {os.path.normpath(DST)}")


if __name__ == "__main__":
    sys.exit(main())
