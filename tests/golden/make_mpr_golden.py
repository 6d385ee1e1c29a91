"""Generates tests/golden/mpr_{60,61,70,128}.npz from the reference's own golden vectors for the narrowphase:
data/unittest_data.zip -> mprPenetrationKernel{60,61,70,128}.bin, serialized launches of mprPenetrationKernel
(b3LauncherCL::serializeArguments, src/Bullet3OpenCL/ParallelPrimitives/b3LauncherCL.cpp:237-269: int numArguments, then per
argument a 32-byte b3KernelArgData {isBuffer, argIndex, sizeInBytes, pad, 16 bytes of data} followed by the buffer bytes), whose
expected contact totals 0 / 1 / 46 / 98 are asserted by test/OpenCL/AllBullet3Kernels/testExecuteBullet3NarrowphaseKernels.cpp:397-413.
The files are 73 MB each because every buffer is serialized at its full capacity; the fixtures keep only what the numPairs pairs
reference (bodies, collidables, hulls, vertices re-indexed), a few KB each.  Run here (needs /root/reference):
    python tests/golden/make_mpr_golden.py"""
import io
import os
import struct
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bullet3_b200 import capi  # noqa: E402

EXPECTED = {"60": 0, "61": 1, "70": 46, "128": 98}


def parse(buf):
    n = struct.unpack_from("i", buf, 0)[0]
    idx, args = 4, []
    for _ in range(n):
        isbuf, _argidx, sz, _pad = struct.unpack_from("iiii", buf, idx)
        data = buf[idx + 16: idx + 32]
        idx += 32
        if isbuf:
            args.append(buf[idx: idx + sz])
            idx += sz
        else:
            args.append(struct.unpack_from("i", data, 0)[0])
    return args


def main():
    z = zipfile.ZipFile("/root/reference/data/unittest_data.zip")
    for tag, want in EXPECTED.items():
        a = parse(z.read("mprPenetrationKernel%s.bin" % tag))
        num_pairs, capacity = a[10], a[9]
        pairs = np.frombuffer(a[0], capi.int4_t)[:num_pairs].copy()
        bodies = np.frombuffer(a[1], capi.rigid_body_t)
        coll = np.frombuffer(a[2], capi.collidable_t)
        convex = np.frombuffer(a[3], capi.convex_t)
        verts = np.frombuffer(a[4], np.dtype(("f4", 4)))
        sep = np.frombuffer(a[5], np.dtype(("f4", 4)))[:num_pairs].copy()
        has = np.frombuffer(a[6], np.int32)[:num_pairs].copy()
        count0 = int(np.frombuffer(a[8], np.int32)[0])
        # keep only what the pairs reference
        used_b = np.unique(np.concatenate([pairs["x"], pairs["y"]]))
        bmap = {int(b): i for i, b in enumerate(used_b)}
        nb = bodies[used_b].copy()
        used_c = np.unique(nb["collidableIdx"])
        cmap = {int(c): i for i, c in enumerate(used_c)}
        nc = coll[used_c].copy()
        hull = nc["shapeType"] == capi.SHAPE_CONVEX_HULL
        used_s = np.unique(nc["shapeIndex"][hull])
        smap = {int(s): i for i, s in enumerate(used_s)}
        ns = convex[used_s].copy()
        nv = []
        off = 0
        for s in ns:
            k = int(s["numVertices"])
            nv.append(verts[int(s["vertexOffset"]): int(s["vertexOffset"]) + k])
            s["vertexOffset"] = off
            off += k
        nv = np.concatenate(nv) if nv else np.zeros((0, 4), np.float32)
        nc["shapeIndex"][hull] = [smap[int(s)] for s in nc["shapeIndex"][hull]]
        nb["collidableIdx"] = [cmap[int(c)] for c in nb["collidableIdx"]]
        pairs["x"] = [bmap[int(b)] for b in pairs["x"]]
        pairs["y"] = [bmap[int(b)] for b in pairs["y"]]
        out = os.path.join(ROOT, "tests", "golden", "mpr_%s.npz" % tag)
        np.savez_compressed(out, pairs=pairs, bodies=nb, collidables=nc, convex=ns, vertices=nv, sep_normals=sep, has_sep_axis=has, count0=count0,
                            capacity=capacity, expected_total=want)
        print(out, "pairs", num_pairs, "bodies", len(nb), "hulls", len(ns), "vertices", len(nv), "count0", count0, "expected", want, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
