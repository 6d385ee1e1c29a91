"""which hull pairs make up the warp-per-item SAT work on the settled bench scene: histogram of (edges A, edges B) over the
contacts of non-small pairs (contacts ~ overlapping SAT items)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench
from bullet3_b200 import capi, scenes
w = capi.World(bench.bench_config(capi, 64))
scenes.bench_config4_scene(w, *bench.scene_dims(64))
w.upload(); w.set_solver(capi.SOLVER_PGS, 10)
w.step_n(1 / 60, 250)
w.update_aabbs(); w.find_pairs(); w.compute_contacts()
t = w.tables(); c = w.contacts(); b = w.bodies()
col = t["collidables"]; cv = t["convex"]; ch = t["child_shapes"]
def shape_of(body, child):
    out = np.full(len(body), -1)
    ci = b["collidableIdx"][body]
    hull = col["shapeType"][ci] == capi.SHAPE_CONVEX_HULL
    out[hull] = col["shapeIndex"][ci[hull]]
    comp = (col["shapeType"][ci] == capi.SHAPE_COMPOUND) & (child >= 0)
    out[comp] = col["shapeIndex"][ch["shapeIndex"][child[comp]]]
    return out
sa = shape_of(np.abs(c["bodyA"]), c["childA"]); sb = shape_of(np.abs(c["bodyB"]), c["childB"])
ok = (sa >= 0) & (sb >= 0)
ea, eb = cv["numUniqueEdges"][sa[ok]], cv["numUniqueEdges"][sb[ok]]
va, vb = cv["numVertices"][sa[ok]], cv["numVertices"][sb[ok]]
small = lambda s: (cv["numVertices"][s] <= 8) & (cv["numFaces"][s] <= 6) & (cv["numUniqueEdges"][s] <= 6)
sm_a, sm_b = small(sa[ok]), small(sb[ok])
both = sm_a & sm_b; one = sm_a ^ sm_b; none = ~sm_a & ~sm_b
print("hull-hull contacts %d: small x small %d, small x larger %d, larger x larger %d" % (ok.sum(), both.sum(), one.sum(), none.sum()))
print("edge pairs (sum of eA*eB): small x larger %.3g, larger x larger %.3g" % ((ea * eb)[one].sum(), (ea * eb)[none].sum()))
print("vertices of the larger hulls:", np.unique(cv["numVertices"][np.unique(np.concatenate([sa[ok][~sm_a], sb[ok][~sm_b]]))]).tolist(), "edges", np.unique(cv["numUniqueEdges"][np.unique(np.concatenate([sa[ok][~sm_a], sb[ok][~sm_b]]))]).tolist())
