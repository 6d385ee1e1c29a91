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
