"""TEST INFRASTRUCTURE — ctypes driver for oracle/_build/libvvoracle.so (the C restatement in
oracle/port/vvoracle.c). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; nothing in vvflow_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libvvoracle.so")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


class _PList(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, C.c_void_p) for k in ("x", "y", "g", "vx", "vy", "ieps", "orig")]


class _Bodies(C.Structure):
    _fields_ = [("nseg", C.c_int64), ("nbody", C.c_int64)] + [
        (k, C.c_void_p)
        for k in ("rx", "ry", "cx", "cy", "dlx", "dly", "g", "ieps", "slip", "body", "bfirst", "bprop",
                  "fric", "gsum", "fdt_dead", "g_dead")
    ]


class _Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("n_leaves", C.c_int64), ("cap", C.c_int64)] + [
        (k, C.c_void_p)
        for k in ("x", "y", "h", "w", "cmp", "cmm", "vfirst", "vlast", "sfirst", "slast", "ch1", "ch2", "leaf",
                  "depth", "leaf_node", "seg_perm")
    ] + [("nseg", C.c_int64)] + [(k, C.c_void_p) for k in ("near_ptr", "near_idx", "far_ptr", "far_idx")] + [
        ("near_cap", C.c_int64), ("far_cap", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        L.vvo_tree_build.restype = C.POINTER(_Tree)
        L.vvo_tree_build.argtypes = [C.POINTER(_PList), C.POINTER(_Bodies), C.c_int, C.c_double, C.c_double]
        L.vvo_tree_free.argtypes = [C.POINTER(_Tree)]
        L.vvo_find_node.restype = C.c_int64
        L.vvo_find_node.argtypes = [C.POINTER(_Tree), C.c_double, C.c_double]
        L.vvo_epsilon.restype = C.c_int64
        L.vvo_epsilon.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.POINTER(_Bodies), C.c_int]
        L.vvo_convective.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.POINTER(_Bodies), C.c_double,
                                     C.c_double, C.c_double, C.c_void_p, C.c_int64]
        L.vvo_diffusive.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.POINTER(_Bodies), C.c_double]
        L.vvo_eps2h_h2_at.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.POINTER(_Bodies), C.c_void_p, C.c_int64, C.c_void_p]
        L.vvo_node_influence.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.POINTER(_Bodies), C.c_void_p]
        L.vvo_vorticity_raster.argtypes = [C.POINTER(_PList), C.POINTER(_Bodies), C.c_float, C.c_float, C.c_float, C.c_int,
                                           C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.vvo_pressure_raster.argtypes = [C.POINTER(_PList), C.POINTER(_Bodies), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                          C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int64,
                                          C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.vvo_velocity_at.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.POINTER(_Bodies), C.c_double, C.c_double,
                                      C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        L.vvo_move_and_clean.restype = C.c_int64
        L.vvo_move_and_clean.argtypes = [C.POINTER(_PList), C.POINTER(_Bodies), C.c_double, C.c_double, C.c_int,
                                         C.POINTER(C.c_int64)]
        L.vvo_point_invalid.restype = C.c_int64
        L.vvo_point_invalid.argtypes = [C.POINTER(_Bodies), C.c_int64, C.c_double, C.c_double]
        L.vvo_set_leaf_sample.argtypes = [C.c_int64, C.c_int64]
        L.vvo_count_interactions.argtypes = [C.POINTER(_Tree), C.POINTER(_PList), C.c_int64, C.c_int64,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Bodies:
    """Flattened body segments. `seg` = (nseg,12) rows as pyref.Ref.segments():
    r.x r.y corner.x corner.y dl.x dl.y g gsum fric _1_eps slip body; `bprop` = (nbody,13):
    axis(2) cofm(2) bl(2) tr(2) disc_r2 inside_valid speed_slae(3)."""

    def __init__(self, seg, bprop):
        seg = np.asarray(seg, dtype=np.float64).reshape(-1, 12)
        self.nseg = seg.shape[0]
        self.bprop = np.ascontiguousarray(bprop, dtype=np.float64).reshape(-1, 13)
        self.nbody = self.bprop.shape[0]
        cols = {k: np.ascontiguousarray(seg[:, i]) for i, k in
                enumerate(("rx", "ry", "cx", "cy", "dlx", "dly", "g", "gsum", "fric", "ieps"))}
        self.a = cols
        self.a["slip"] = np.ascontiguousarray(seg[:, 10].astype(np.int32))
        self.a["body"] = np.ascontiguousarray(seg[:, 11].astype(np.int32))
        bf = np.zeros(self.nbody + 1, dtype=np.int64)
        for b in range(self.nbody):
            bf[b + 1] = bf[b] + int(np.sum(self.a["body"] == b))
        self.a["bfirst"] = bf
        self.a["fdt_dead"] = np.zeros(3 * max(1, self.nbody))
        self.a["g_dead"] = np.zeros(max(1, self.nbody))
        self.c = _Bodies(self.nseg, self.nbody, *[_ptr(self.a[k]) if k != "bprop" else _ptr(self.bprop) for k in
                                                  ("rx", "ry", "cx", "cy", "dlx", "dly", "g", "ieps", "slip", "body",
                                                   "bfirst", "bprop", "fric", "gsum", "fdt_dead", "g_dead")])

    @staticmethod
    def from_ref(ref):
        """Snapshot the bodies of a pyref.Ref."""
        seg = ref.segments()
        bp = np.zeros((ref.n_bodies, 13))
        for b in range(ref.n_bodies):
            o = ref.body(b)
            bp[b, :10] = o[:10]
            bp[b, 10:13] = o[10:13]
        return Bodies(seg, bp)


class Port:
    """Particle list + (optional) bodies driven through the C restatement."""

    def __init__(self, rec48=None, xyg=None, bodies=None):
        if rec48 is not None:
            r = np.asarray(rec48, dtype=np.float64).reshape(-1, 6)
            cols = [r[:, 0], r[:, 1], r[:, 2], r[:, 3], r[:, 4], r[:, 5]]
        else:
            a = np.asarray(xyg, dtype=np.float64).reshape(-1, 3)
            z = np.zeros(a.shape[0])
            cols = [a[:, 0], a[:, 1], a[:, 2], z, z, z]
        self.x, self.y, self.g, self.vx, self.vy, self.ieps = [np.ascontiguousarray(c).copy() for c in cols]
        self.orig = np.arange(self.x.shape[0], dtype=np.int64)
        self.bodies = bodies
        self.L = lib()
        self.tree = None
        self._mk()

    def _mk(self):
        self.p = _PList(self.x.shape[0], *[_ptr(a) for a in (self.x, self.y, self.g, self.vx, self.vy, self.ieps,
                                                              self.orig)])

    @property
    def n(self):
        return self.p.n

    def _b(self):
        return C.byref(self.bodies.c) if self.bodies is not None else None

    def rec48(self):
        n = self.p.n
        return np.stack([self.x[:n], self.y[:n], self.g[:n], self.vx[:n], self.vy[:n], self.ieps[:n]], axis=1)

    def tree_build(self, far=8, min_node=0.0, max_node=np.finfo(np.float64).max):
        self.tree_destroy()
        self.tree = self.L.vvo_tree_build(C.byref(self.p), self._b(), far, min_node, max_node)

    def tree_destroy(self):
        if self.tree:
            self.L.vvo_tree_free(self.tree)
            self.tree = None

    def __del__(self):
        try:
            self.tree_destroy()
        except Exception:
            pass

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        ct = C.c_double if dtype == np.float64 else C.c_int64
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()

    def tree_export(self):
        """Same layout as pyref.Ref.tree_export(): (dbl (n,10), idx (n,10), n_leaves)."""
        t = self.tree.contents
        n = t.n_nodes
        f = lambda p, k=1: self._arr(p, n * k, np.float64)
        i = lambda p: self._arr(p, n, np.int64)
        dbl = np.zeros((n, 10))
        dbl[:, 0], dbl[:, 1], dbl[:, 2], dbl[:, 3] = f(t.x), f(t.y), f(t.h), f(t.w)
        dbl[:, 4:7] = f(t.cmp, 3).reshape(n, 3)
        dbl[:, 7:10] = f(t.cmm, 3).reshape(n, 3)
        idx = np.zeros((n, 10), dtype=np.int64)
        idx[:, 0], idx[:, 1] = i(t.vfirst), i(t.vlast)
        idx[:, 6] = i(t.slast) - i(t.sfirst)
        idx[:, 7], idx[:, 8], idx[:, 9] = i(t.ch1), i(t.ch2), i(t.leaf)
        return dbl, idx, t.n_leaves

    def tree_lists(self):
        t = self.tree.contents
        nl = t.n_leaves
        nptr = self._arr(t.near_ptr, nl + 1, np.int64)
        fptr = self._arr(t.far_ptr, nl + 1, np.int64)
        return nptr, self._arr(t.near_idx, int(nptr[-1]), np.int64), fptr, self._arr(t.far_idx, int(fptr[-1]), np.int64)

    def tree_leaf_segments(self):
        t = self.tree.contents
        sf = self._arr(t.sfirst, t.n_nodes, np.int64)
        sl = self._arr(t.slast, t.n_nodes, np.int64)
        ln = self._arr(t.leaf_node, t.n_leaves, np.int64)
        perm = self._arr(t.seg_perm, t.nseg, np.int64)
        ptr = np.zeros(t.n_leaves + 1, dtype=np.int64)
        ptr[1:] = np.cumsum(sl[ln] - sf[ln])
        idx = np.concatenate([perm[sf[k]:sl[k]] for k in ln]) if t.n_leaves and t.nseg else np.zeros(0, dtype=np.int64)
        return ptr, idx

    def find_node(self, x, y):
        return self.L.vvo_find_node(self.tree, x, y)

    def epsilon(self, merge):
        return self.L.vvo_epsilon(self.tree, C.byref(self.p), self._b(), int(merge))

    def convective(self, inf_vx=0.0, inf_vy=0.0, dt=0.0, sinks=None):
        s = np.zeros((0, 3)) if sinks is None else np.ascontiguousarray(sinks, dtype=np.float64).reshape(-1, 3)
        self.L.vvo_convective(self.tree, C.byref(self.p), self._b(), inf_vx, inf_vy, dt, _ptr(s), s.shape[0])

    def velocity_at(self, xy, inf_vx=0.0, inf_vy=0.0, dt=0.0, sinks=None):
        """MConvectiveFast::velocity at arbitrary points (needs a built tree and epsilon)"""
        s = np.zeros((0, 3)) if sinks is None else np.ascontiguousarray(sinks, dtype=np.float64).reshape(-1, 3)
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(a)
        self.L.vvo_velocity_at(self.tree, C.byref(self.p), self._b(), inf_vx, inf_vy, dt, _ptr(s), s.shape[0],
                               _ptr(a), a.shape[0], _ptr(out))
        return out

    def eps2h_h2_at(self, xy):
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(a)
        self.L.vvo_eps2h_h2_at(self.tree, C.byref(self.p), self._b(), _ptr(a), a.shape[0], _ptr(out))
        return out

    def node_influence(self):
        out = np.zeros(int(self.bodies.nseg) if self.bodies is not None else 0)
        if out.shape[0]:
            self.L.vvo_node_influence(self.tree, C.byref(self.p), self._b(), _ptr(out))
        return out

    def vorticity_raster(self, xmin, ymin, dxdy, xres, yres, eps_mult, dl):
        """XVorticity::evaluate on this (post-shed) list; the list is permuted by the tree built inside.
        (yres, xres) float64 — the reference stores float32 of the same values."""
        assert self.tree is None or not self.tree, "destroy the tree first"
        out = np.zeros((yres, xres))
        self.L.vvo_vorticity_raster(C.byref(self.p), self._b(), xmin, ymin, dxdy, xres, yres, eps_mult, dl, _ptr(out))
        return out

    def pressure_raster(self, xmin, ymin, dxdy, xres, yres, dl, re, dt, inf_vx=0.0, inf_vy=0.0, sinks=None, ref_speed=None):
        """XPressure::evaluate on this (post-shed) list and the bodies' post-shed gsum; the list is permuted by the tree built
        inside. ref_speed = None is ref_frame 's'. (yres, xres) float64 — the reference stores float32 of the same values."""
        assert self.tree is None or not self.tree, "destroy the tree first"
        out = np.zeros((yres, xres))
        s = np.zeros((0, 3)) if sinks is None else np.ascontiguousarray(sinks, dtype=np.float64).reshape(-1, 3)
        rs = (0.0, 0.0) if ref_speed is None else ref_speed
        self.L.vvo_pressure_raster(C.byref(self.p), self._b(), xmin, ymin, dxdy, xres, yres, dl, re, dt, inf_vx, inf_vy, _ptr(s),
                                   s.shape[0], 0 if ref_speed is None else 1, rs[0], rs[1], _ptr(out))
        return out

    def diffusive(self, re):
        self.L.vvo_diffusive(self.tree, C.byref(self.p), self._b(), re)

    def move_and_clean(self, dt, remove_eps=1e-10, remove=True):
        cl = C.c_int64(0)
        n = self.L.vvo_move_and_clean(C.byref(self.p), self._b(), dt, remove_eps, int(remove), C.byref(cl))
        return n, cl.value

    def sample_leaves(self, stride, phase=0):
        self.L.vvo_set_leaf_sample(stride, phase)

    def count_interactions(self, l0=0, l1=None):
        a, b = C.c_double(), C.c_double()
        if l1 is None:
            l1 = self.tree.contents.n_leaves
        self.L.vvo_count_interactions(self.tree, C.byref(self.p), l0, l1, C.byref(a), C.byref(b))
        return a.value, b.value
