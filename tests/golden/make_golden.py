"""Generates tests/golden/*.npz from the reference's own code (oracle/_ref/libvvref.so, built by
oracle/Makefile from /root/reference). Run in the build container: python tests/golden/make_golden.py

Each fixture holds the inputs and the reference's state after every phase of the hot path
(vvflow.cpp:246-257): tree (pre-order node table, interaction lists), epsilon(+merge), convective,
diffusive, move_and_clean — plus, on the state after epsilon, the point evaluators (velocity(p), eps2h, h2) and
the SLAE right-hand-side term NodeInfluence. The oracle port and the CUDA path are both tested against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
os.environ["OMP_NUM_THREADS"] = "1"

import cases  # noqa: E402
from oracle import pyref  # noqa: E402


def snapshot(r, with_body):
    d = {}
    d["in48"] = r.get_list48()
    if with_body:
        d["seg_in"] = r.segments()
        d["body_in"] = np.stack([r.body(b) for b in range(r.n_bodies)])
    mn, mx = r.tree_params(8)
    d["tree_params"] = np.array([8, mn, mx])
    r.tree_build()
    dbl, idx, nl = r.tree_export()
    d["tree_dbl"], d["tree_idx"], d["n_leaves"] = dbl, idx, np.array([nl])
    l = r.tree_lists(nl)
    d["near_ptr"], d["near_idx"], d["far_ptr"], d["far_idx"] = l
    if with_body:
        d["lseg_ptr"], d["lseg_idx"] = r.tree_leaf_segments(nl)
    d["after_build"] = r.get_list48()
    d["interactions"] = np.array(r.count_interactions())
    d["merged"] = np.array([r.epsilon(True)])
    d["after_eps"] = r.get_list48()
    # SURVEY 8(f) rows 1 and 4, evaluated on this state (tree built, _1_eps fresh): MConvectiveFast::velocity(p),
    # static MEpsilonFast::eps2h / h2 of findNode(p), MConvectiveFast::NodeInfluence of every segment
    rng = np.random.default_rng(7)
    xy = d["after_eps"][:, :2]
    lo, hi = xy.min(0), xy.max(0)
    pts = np.concatenate([rng.uniform(lo, hi, (400, 2)), xy[rng.integers(0, xy.shape[0], 30)],
                          rng.uniform(lo - 2 * (hi - lo), hi + 2 * (hi - lo), (70, 2))])
    d["pts"] = pts
    d["vel_at_pts"] = r.velocity_at(pts)
    d["eps2h_h2_at_pts"] = r.eps2h_h2_at(pts)
    if with_body:
        d["node_influence"] = r.node_influence()
    r.convective()
    d["after_conv"] = r.get_list48()
    r.diffusive()
    d["after_diff"] = r.get_list48()
    if with_body:
        d["seg_after_diff"] = r.segments()
    r.tree_destroy()
    d["cleaned"] = np.array([r.move_and_clean(True)])
    d["after_move"] = r.get_list48()
    if with_body:
        d["seg_after_move"] = r.segments()
        d["body_after_move"] = np.stack([r.body(b) for b in range(r.n_bodies)])
    return d


def main():
    # 1. body-free mixed-sign cloud (merge stress, BASELINE config 5b shape)
    r = pyref.Ref(re=600, dt=0.05, inf_vx=1.0)
    r.set_list(cases.cloud(3000, "gauss", "mixed", seed=101))
    d = snapshot(r, False)
    d["params"] = np.array([600, 0.05, 1.0, 0.0])
    np.savez_compressed(os.path.join(HERE, "cloud_mixed_3000.npz"), **d)
    print("cloud_mixed_3000: leaves", d["n_leaves"][0], "merged", d["merged"][0])

    # 2. same-sign Lamb-Oseen blob (BASELINE config 2 shape, small)
    r = pyref.Ref(re=1000, dt=0.005, inf_vx=1.0)
    r.set_list(cases.cloud(4000, "gauss", "equal", seed=12345))
    d = snapshot(r, False)
    d["params"] = np.array([1000, 0.005, 1.0, 0.0])
    np.savez_compressed(os.path.join(HERE, "blob_same_4000.npz"), **d)
    print("blob_same_4000: leaves", d["n_leaves"][0], "merged", d["merged"][0])

    # 3. the bundled case example/cyl_re600.lua: cylinder R=0.5, 350 segments, re=600, dt=0.05,
    #    U=(1,0). Step it with the reference's own loop (SLAE included); README.md:117-120 rows are
    #    force_hydro at t = 0, .05, .10, .15. Then snapshot the hot path of step 30.
    r = pyref.Ref(re=600, dt=0.05, inf_vx=1.0)
    r.add_cylinder(0.5, 350)
    r.tree_params(8)
    rows = []
    for step in range(30):
        r.step_pre()
        rows.append(r.body(0)[17:20].copy())
        r.step_hot()
    d = {"force_hydro": np.stack(rows)}
    r.step_pre()
    d.update(snapshot(r, True))
    d["params"] = np.array([600, 0.05, 1.0, 0.0])
    np.savez_compressed(os.path.join(HERE, "cyl_re600_step30.npz"), **d)
    print("cyl_re600_step30: N", d["in48"].shape[0], "leaves", d["n_leaves"][0], "merged", d["merged"][0],
          "cleaned", d["cleaned"][0])
    for k in range(4):
        print("  force_hydro t=%.2f: %+.6e %+.6e %+.6e" % (0.05 * k, *rows[k]))

    # 4. point sinks / sources (Space::SourceList) through process_all_lists and velocity(p): strengths below and above 1
    #    (MConvectiveFast.cpp:165 truncates the strength to an integer in the reference build)
    xyg = cases.cloud(1500, "gauss", "mixed", seed=33)
    sinks = np.array([[0.3, 0.2, 0.2], [-0.5, 0.1, -1.5], [0.0, -0.4, 2.7], [1.1, 0.9, -0.99]])
    r = pyref.Ref(re=600, dt=0.05, inf_vx=1.0, inf_vy=0.25)
    r.set_list(xyg)
    r.set_list(sinks, pyref.SOURCE)
    r.tree_build()
    r.epsilon(False)
    d = {"xyg": xyg, "sinks": sinks, "params": np.array([600, 0.05, 1.0, 0.25])}
    rng = np.random.default_rng(5)
    d["pts"] = np.concatenate([rng.standard_normal((60, 2)), sinks[:, :2] + 1e-3])
    d["vel_at_pts"] = r.velocity_at(d["pts"])
    r.convective()
    d["after_conv"] = r.get_list48()
    np.savez_compressed(os.path.join(HERE, "sinks_1500.npz"), **d)
    print("sinks_1500: |v|max", np.abs(d["after_conv"][:, 3:5]).max())


if __name__ == "__main__":
    main()
