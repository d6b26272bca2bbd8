"""Seeded synthetic inputs shared by the parity tests (test infrastructure)."""
import numpy as np

from vvflow_b200.vvhd import TBody

DBL_MAX = float(np.finfo(np.float64).max)


def cloud(n, kind="gauss", sign="same", seed=1):
    """(n,3) x,y,g records. kind: gauss (Lamb-Oseen blob, BASELINE config 2) | uniform (config 5);
    sign: same (no merges) | mixed (merge stress, config 5b) | equal (g = 1/n)."""
    rng = np.random.default_rng(seed)
    xyg = np.zeros((n, 3))
    xyg[:, :2] = rng.standard_normal((n, 2)) if kind == "gauss" else rng.uniform(0, 1, (n, 2))
    if sign == "equal":
        xyg[:, 2] = 1.0 / max(n, 1)
    elif sign == "same":
        xyg[:, 2] = rng.uniform(0.5, 1, n) / max(n, 1)
    else:
        xyg[:, 2] = rng.uniform(-1, 1, n) / max(n, 1)
    return xyg


def cylinder(R=0.5, nseg=350, cx=0.0, cy=0.0):
    """gen_cylinder = gen_arc_N(c, R, 2pi -> 0, N), utils/vvflow/gen_cylinder.cpp:33-41, gen_body.cpp:74-88"""
    i = np.arange(nseg, dtype=np.float64)
    a = 2 * np.pi + (0 - 2 * np.pi) * i / nseg
    return TBody(np.stack([cx + R * np.cos(a), cy + R * np.sin(a)], axis=1))


def around_cylinder(n, R=0.5, spread=0.3, sign="mixed", seed=2, inside=20):
    """particles in a shell around a cylinder, a few of them inside the body"""
    rng = np.random.default_rng(seed)
    rad = R + np.abs(rng.standard_normal(n)) * spread
    th = rng.uniform(0, 2 * np.pi, n)
    xyg = np.zeros((n, 3))
    xyg[:, 0], xyg[:, 1] = rad * np.cos(th), rad * np.sin(th)
    xyg[:inside, :2] *= 0.5
    xyg[:, 2] = (rng.uniform(-1, 1, n) if sign == "mixed" else rng.uniform(0.5, 1, n)) / n
    return xyg


def tree_params(bodies):
    """vvflow.cpp:200-203"""
    if not bodies:
        return 0.0, DBL_MAX
    b = bodies[0]
    dl = b.slen / (b.size() - 1) if b.size() > 1 else 0.0
    return (dl * 5, dl * 100) if dl > 0 else (0.0, DBL_MAX)


def port_bodies(pyport, bodies):
    """vvhd.TBody list -> oracle pyport.Bodies"""
    if not bodies:
        return None
    rows, props = [], []
    for ib, b in enumerate(bodies):
        n = b.size()
        seg = np.zeros((n, 12))
        seg[:, 0:2], seg[:, 2:4], seg[:, 4:6] = b.r, b.corner, b.dl
        seg[:, 6], seg[:, 7], seg[:, 8], seg[:, 9], seg[:, 10], seg[:, 11] = b.g, 0, 0, b.ieps, b.slip, ib
        rows.append(seg)
        props.append(np.concatenate([b.axis, b.cofm, b.bl, b.tr, [b.disc_r2, float(b.inside_valid)], b.speed_slae]))
    return pyport.Bodies(np.concatenate(rows), np.stack(props))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def relerr(a, b):
    """max |a-b| scaled by the largest magnitude of the reference field (norm-wise relative error).
    NaN matches NaN (the reference itself yields NaN for a lone particle)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    na, nb = np.isnan(a), np.isnan(b)
    if np.any(na != nb):
        return float("inf")
    ok = ~na
    if not np.any(ok):
        return 0.0
    scale = max(float(np.max(np.abs(b[ok]))), 1e-300)
    return float(np.max(np.abs(a[ok] - b[ok]))) / scale


def relerr_elem(a, b, floor=1e-4):
    """element-wise relative error with a floor: max_i |a_i - b_i| / max(|b_i|, floor * ||b||_inf). For (n, 2) arrays an
    element is a ROW (a particle's velocity vector). Returns (error, index of the worst element)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0, -1
    na, nb = np.isnan(a), np.isnan(b)
    if np.any(na != nb):
        return float("inf"), int(np.argmax((na != nb).reshape(a.shape[0], -1).any(axis=1)))
    a, b = np.where(na, 0.0, a), np.where(nb, 0.0, b)
    if a.ndim == 2:
        diff, mag = np.hypot.reduce(a - b, axis=1), np.hypot.reduce(b, axis=1)
    else:
        diff, mag = np.abs(a - b), np.abs(b)
    scale = np.maximum(mag, floor * max(float(mag.max()), 1e-300))
    e = diff / scale
    k = int(np.argmax(e))
    return float(e[k]), k


def check_close(a, b, tol, what):
    """the velocity / integral-sum bar of north_star (1e-10 relative): norm-wise AND element-wise (with a floor of
    1e-4 of the field maximum, so that exact zeros and cancellation residues do not divide by nothing)"""
    nw = relerr(a, b)
    ew, k = relerr_elem(a, b)
    assert nw <= tol and ew <= tol, (f"{what}: norm-wise {nw:.3e}, element-wise {ew:.3e} at element {k}: "
                                     f"got {np.asarray(a)[k] if k >= 0 else None}, want {np.asarray(b)[k] if k >= 0 else None}")
    return max(nw, ew)


# ---- golden fixtures (tests/golden/*.npz, generated from the compiled reference) ------------------
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ("cloud_mixed_3000", "blob_same_4000", "cyl_re600_step30")   # + sinks_1500 (convective only)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def golden_bodies(d):
    """vvhd.TBody list with the exact TBody state stored in a fixture"""
    if "seg_in" not in d:
        return []
    seg, body = d["seg_in"], d["body_in"]
    out = []
    for ib in range(body.shape[0]):
        rows = seg[seg[:, 11] == ib]
        b = TBody.from_oracle(rows, body[ib])
        b.gsum[:] = 0
        b.fric[:] = 0
        out.append(b)
    return out
