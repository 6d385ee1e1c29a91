"""max / histogram of contacts per dynamic body on the settled bench scene vs number of batches"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from bullet3_b200 import capi, scenes
w = capi.World(bench.bench_config(capi, 64))
scenes.bench_config4_scene(w, *bench.scene_dims(64))
w.upload(); w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, 250)
w.update_aabbs(); w.find_pairs(); w.compute_contacts(); w.solver_setup()
c = w.contacts(); b = w.bodies()
inv = b["invMass"]
ids = np.concatenate([np.abs(c["bodyA"]), np.abs(c["bodyB"])])
ids = ids[inv[ids] != 0]
deg = np.bincount(ids, minlength=len(b))
print("contacts", len(c), "max degree", deg.max(), "batches", len(w.batches()) - 1)
print("degree histogram", np.bincount(deg)[:40].tolist())
# the same with all manifolds of one body pair merged into one group (one colour per pair instead of per manifold)
a, bb = np.abs(c["bodyA"]).astype(np.int64), np.abs(c["bodyB"]).astype(np.int64)
key = np.minimum(a, bb) * (1 << 32) + np.maximum(a, bb)
uk, cnt = np.unique(key, return_counts=True)
ga, gb = (uk >> 32).astype(np.int64), (uk & 0xFFFFFFFF).astype(np.int64)
gid = np.concatenate([ga, gb])
gid = gid[inv[gid] != 0]
gdeg = np.bincount(gid, minlength=len(b))
print("pair groups", len(uk), "max group degree", gdeg.max(), "group size histogram", np.bincount(cnt)[:40].tolist(), "max group", cnt.max())
print("group degree histogram", np.bincount(gdeg)[:40].tolist())
print("batch sizes", np.diff(w.batches()).tolist())
