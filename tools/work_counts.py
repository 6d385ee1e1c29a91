"""work-item funnel of the narrowphase on the settled bench scene"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from bullet3_b200 import capi, scenes

w = capi.World(bench.bench_config(capi, 64))
scenes.bench_config4_scene(w, *bench.scene_dims(64))
w.upload()
w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, 250)
w.update_aabbs()
w.find_pairs()
w.compute_contacts()
c = w.work_counters()
ct = w.contacts()
mesh = int((np.abs(ct["bodyA"]) == 0).sum())
print("pairs %d | compound raw child pairs %d | small items %d | SAT items %d -> overlapping %d -> contacts (non-mesh) %d | triangle raw items %d -> after quick reject %d -> mesh contacts %d" % (
    c[0], c[5], c[16] + c[17], c[8], c[9], len(ct) - mesh, c[6], c[11] + c[19], mesh))
