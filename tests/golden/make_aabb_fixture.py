"""Generate tests/golden/aabbs64006.npz from the reference's broadphase input dump
data/64006GPUAABBs.txt (64 006 lines "minx miny minz maxx maxy maxz"), the input of
config 2 / PairBench (examples/OpenCL/broadphase/PairBench.cpp:218-320).  The raw
values are stored as float32 (what parseFloat yields); the x0.1 scaling, the
large/small split (extent length > 500) and the handle numbering from 1024 are
applied by tests/pairbench.py exactly as PairBench.cpp:278-305 does.

Run in the build container only (needs /root/reference):
    python tests/golden/make_aabb_fixture.py
"""
import os

import numpy as np

SRC = "/root/reference/data/64006GPUAABBs.txt"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "aabbs64006.npz")

if __name__ == "__main__":
    a = np.loadtxt(SRC, dtype=np.float64).astype(np.float32)
    assert a.shape == (64006, 6), a.shape
    np.savez_compressed(DST, aabbs=a)
    print(DST, os.path.getsize(DST), "bytes")
