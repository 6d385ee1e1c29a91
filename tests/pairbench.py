"""PairBench input recipe (examples/OpenCL/broadphase/PairBench.cpp:218-320) on the
committed fixture tests/golden/aabbs64006.npz."""
import os

import numpy as np

from bullet3_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
TEST_INDEX_OFFSET = 1024  # PairBench.cpp:77


def load_pairbench_aabbs(limit=None):
    """returns (aabbs[aabb_t] in creation order, small_idx, large_idx)"""
    raw = np.load(os.path.join(HERE, "golden", "aabbs64006.npz"))["aabbs"]
    if limit is not None:
        raw = raw[:limit]
    s = np.float32(0.1)
    mn = raw[:, 0:3] * s  # aabbMin *= 0.1 (float)
    mx = raw[:, 3:6] * s
    ext = mx - mn
    length = np.sqrt((ext[:, 0] * ext[:, 0] + ext[:, 1] * ext[:, 1]) + ext[:, 2] * ext[:, 2]).astype(np.float32)
    large = length > np.float32(500)
    aabbs = np.zeros(len(raw), capi.aabb_t)
    aabbs["min"] = mn
    aabbs["max"] = mx
    # small proxies get consecutive handles from 1024; a large proxy takes the current handle
    # without advancing it (PairBench.cpp:291-305)
    handle = TEST_INDEX_OFFSET + np.cumsum(~large) - (~large)
    aabbs["minIndex"] = handle.astype(np.int32)
    aabbs["maxIndex"] = np.arange(len(raw), dtype=np.int32)
    return aabbs, np.nonzero(~large)[0].astype(np.int32), np.nonzero(large)[0].astype(np.int32)
