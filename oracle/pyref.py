"""TEST INFRASTRUCTURE — ctypes driver for oracle/_ref/libvvref.so.

libvvref.so is the reference's own hot-path code (libvvhd TSortedTree, MEpsilonFast,
MConvectiveFast, MDiffusiveFast, MFlowmove, TBody, TMatrix, TSpace) compiled unmodified by
oracle/Makefile, with the C entry points of oracle/ref_shim/ref_capi.cpp around it.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing in vvflow_b200/ does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libvvref.so")

VORT, HEAT, STREAK, SOURCE = 0, 1, 2, 3

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(REF_SO)


def _load():
    lib = C.CDLL(REF_SO)
    lib.vvr_create.restype = C.c_void_p
    sig = {
        "vvr_destroy": (None, [C.c_void_p]),
        "vvr_set_params": (None, [C.c_void_p] + [C.c_double] * 5),
        "vvr_time": (C.c_double, [C.c_void_p]),
        "vvr_dt": (C.c_double, [C.c_void_p]),
        "vvr_set_list": (None, [C.c_void_p, C.c_int, _dp, C.c_size_t]),
        "vvr_set_list48": (None, [C.c_void_p, C.c_int, _dp, C.c_size_t]),
        "vvr_list_size": (C.c_size_t, [C.c_void_p, C.c_int]),
        "vvr_get_list48": (None, [C.c_void_p, C.c_int, _dp]),
        "vvr_add_cylinder": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_size_t]),
        "vvr_add_polygon": (C.c_int, [C.c_void_p, _dp, C.c_void_p, C.c_size_t]),
        "vvr_body_set_dynamics": (None, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
        "vvr_body_set_speed_slae": (None, [C.c_void_p, C.c_int, _dp]),
        "vvr_n_bodies": (C.c_size_t, [C.c_void_p]),
        "vvr_n_segments": (C.c_size_t, [C.c_void_p]),
        "vvr_get_segments": (None, [C.c_void_p, _dp]),
        "vvr_set_segments_ggf": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
        "vvr_get_body": (None, [C.c_void_p, C.c_int, _dp]),
        "vvr_tree_default_params": (None, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "vvr_tree_params": (None, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
        "vvr_tree_build": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
        "vvr_tree_destroy": (None, [C.c_void_p]),
        "vvr_sample_leaves": (C.c_size_t, [C.c_void_p, C.c_size_t, C.c_size_t]),
        "vvr_count_interactions": (None, [C.c_void_p] + [C.POINTER(C.c_double)] * 3),
        "vvr_tree_counts": (None, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "vvr_tree_export": (None, [C.c_void_p, _dp, _ip]),
        "vvr_tree_lists": (None, [C.c_void_p, _ip, C.c_void_p, _ip, C.c_void_p]),
        "vvr_tree_leaf_segments": (None, [C.c_void_p, _ip, C.c_void_p]),
        "vvr_find_node": (C.c_int64, [C.c_void_p, C.c_double, C.c_double]),
        "vvr_epsilon": (C.c_int, [C.c_void_p, C.c_int]),
        "vvr_convective": (None, [C.c_void_p]),
        "vvr_velocity_at": (None, [C.c_void_p, _dp, C.c_size_t, _dp]),
        "vvr_eps2h_h2_at": (None, [C.c_void_p, _dp, C.c_size_t, _dp]),
        "vvr_node_influence": (None, [C.c_void_p, _dp]),
        "vvr_vorticity_raster": (None, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_double,
                                        np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")]),
        "vvr_pressure_raster": (None, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                       np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")]),
        "vvr_diffusive": (None, [C.c_void_p, C.c_int, C.c_int]),
        "vvr_move_and_clean": (C.c_size_t, [C.c_void_p, C.c_int]),
        "vvr_calc_circulation": (None, [C.c_void_p]),
        "vvr_vortex_shed": (None, [C.c_void_p]),
        "vvr_calc_forces": (None, [C.c_void_p]),
        "vvr_zero_forces": (None, [C.c_void_p]),
        "vvr_advance_time": (None, [C.c_void_p]),
        "vvr_step_pre": (None, [C.c_void_p]),
        "vvr_step_hot": (None, [C.c_void_p, C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


class Ref:
    """One reference `Space` + tree + the four solver modules."""

    def __init__(self, re=float("inf"), pr=0.0, dt=1.0, inf_vx=0.0, inf_vy=0.0):
        self.L = lib()
        self.h = self.L.vvr_create()
        self.L.vvr_set_params(self.h, re, pr, dt, inf_vx, inf_vy)

    def close(self):
        if self.h:
            self.L.vvr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- particle lists
    def set_list(self, xyg, which=VORT):
        a = np.ascontiguousarray(xyg, dtype=np.float64).reshape(-1, 3)
        self.L.vvr_set_list(self.h, which, a, a.shape[0])

    def set_list48(self, rec, which=VORT):
        a = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 6)
        self.L.vvr_set_list48(self.h, which, a, a.shape[0])

    def get_list48(self, which=VORT):
        n = self.L.vvr_list_size(self.h, which)
        out = np.zeros((n, 6))
        if n:
            self.L.vvr_get_list48(self.h, which, out)
        return out

    # ---- bodies
    def add_cylinder(self, R, N, cx=0.0, cy=0.0):
        return self.L.vvr_add_cylinder(self.h, cx, cy, R, N)

    def add_polygon(self, xy, slip=None):
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        sp = None
        if slip is not None:
            s = np.ascontiguousarray(slip, dtype=np.uint32)
            sp = s.ctypes.data_as(C.c_void_p)
        return self.L.vvr_add_polygon(self.h, a, sp, a.shape[0])

    def body_set_dynamics(self, b, kspring=None, damping=None, density=1.0, speed=None):
        def p(v):
            if v is None:
                return None, None
            arr = np.ascontiguousarray(v, dtype=np.float64)
            return arr, arr.ctypes.data_as(C.c_void_p)
        k, kp = p(kspring); d, dp_ = p(damping); s, sp = p(speed)
        self.L.vvr_body_set_dynamics(self.h, b, kp, dp_, density, sp)

    def body_set_speed_slae(self, b, v3):
        self.L.vvr_body_set_speed_slae(self.h, b, np.ascontiguousarray(v3, dtype=np.float64))

    @property
    def n_bodies(self):
        return self.L.vvr_n_bodies(self.h)

    def segments(self):
        """(nseg, 12): r.x r.y corner.x corner.y dl.x dl.y g gsum fric _1_eps slip body"""
        n = self.L.vvr_n_segments(self.h)
        out = np.zeros((n, 12))
        if n:
            self.L.vvr_get_segments(self.h, out)
        return out

    def set_segments(self, g=None, gsum=None, fric=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (g, gsum, fric)]
        ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
        self.L.vvr_set_segments_ggf(self.h, *ptrs)

    def body(self, b):
        out = np.zeros(32)
        self.L.vvr_get_body(self.h, b, out)
        return out

    # ---- tree
    def tree_default_params(self):
        a, b = C.c_double(), C.c_double()
        self.L.vvr_tree_default_params(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def tree_params(self, far=8, min_node=None, max_node=None):
        if min_node is None or max_node is None:
            min_node, max_node = self.tree_default_params()
        self.L.vvr_tree_params(self.h, far, min_node, max_node)
        return min_node, max_node

    def tree_build(self, v=True, b=True, h=True):
        self.L.vvr_tree_build(self.h, int(v), int(b), int(h))

    def tree_destroy(self):
        self.L.vvr_tree_destroy(self.h)

    def sample_leaves(self, stride, phase=0):
        return self.L.vvr_sample_leaves(self.h, stride, phase)

    def count_interactions(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.vvr_count_interactions(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def tree_export(self):
        nn, nl = C.c_size_t(), C.c_size_t()
        self.L.vvr_tree_counts(self.h, C.byref(nn), C.byref(nl))
        dbl = np.zeros((nn.value, 10))
        idx = np.zeros((nn.value, 10), dtype=np.int64)
        self.L.vvr_tree_export(self.h, dbl, idx)
        return dbl, idx, nl.value

    def tree_lists(self, n_leaves):
        nptr = np.zeros(n_leaves + 1, dtype=np.int64)
        fptr = np.zeros(n_leaves + 1, dtype=np.int64)
        self.L.vvr_tree_lists(self.h, nptr, None, fptr, None)
        nidx = np.zeros(max(1, nptr[-1]), dtype=np.int64)
        fidx = np.zeros(max(1, fptr[-1]), dtype=np.int64)
        self.L.vvr_tree_lists(self.h, nptr, nidx.ctypes.data_as(C.c_void_p), fptr, fidx.ctypes.data_as(C.c_void_p))
        return nptr, nidx[: nptr[-1]], fptr, fidx[: fptr[-1]]

    def tree_leaf_segments(self, n_leaves):
        ptr = np.zeros(n_leaves + 1, dtype=np.int64)
        self.L.vvr_tree_leaf_segments(self.h, ptr, None)
        idx = np.zeros(max(1, ptr[-1]), dtype=np.int64)
        self.L.vvr_tree_leaf_segments(self.h, ptr, idx.ctypes.data_as(C.c_void_p))
        return ptr, idx[: ptr[-1]]

    def find_node(self, x, y):
        return self.L.vvr_find_node(self.h, x, y)

    # ---- phases
    def epsilon(self, merge):
        return self.L.vvr_epsilon(self.h, int(merge))

    def convective(self):
        self.L.vvr_convective(self.h)

    def velocity_at(self, xy):
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(a)
        self.L.vvr_velocity_at(self.h, a, a.shape[0], out)
        return out

    def eps2h_h2_at(self, xy):
        """(MEpsilonFast::eps2h, MEpsilonFast::h2) of findNode(p) at every point"""
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(a)
        self.L.vvr_eps2h_h2_at(self.h, a, a.shape[0], out)
        return out

    def node_influence(self):
        """MConvectiveFast::NodeInfluence(findNode(seg.r), seg) of every segment (SLAE right-hand side, vortex term)"""
        out = np.zeros(self.segments().shape[0])
        self.L.vvr_node_influence(self.h, out)
        return out

    def vorticity_raster(self, xmin, ymin, dxdy, xres, yres, eps_mult):
        """XVorticity(S, ...).evaluate(): (yres, xres) float32 map"""
        out = np.zeros((yres, xres), dtype=np.float32)
        self.L.vvr_vorticity_raster(self.h, xmin, ymin, dxdy, xres, yres, eps_mult, out)
        return out

    def pressure_raster(self, xmin, ymin, dxdy, xres, yres, ref_frame="s"):
        """XPressure(S, ...).evaluate(): (yres, xres) float32 map"""
        out = np.zeros((yres, xres), dtype=np.float32)
        self.L.vvr_pressure_raster(self.h, xmin, ymin, dxdy, xres, yres, ord(ref_frame), out)
        return out

    def diffusive(self, vort=True, heat=False):
        self.L.vvr_diffusive(self.h, int(vort), int(heat))

    def move_and_clean(self, remove=True):
        return self.L.vvr_move_and_clean(self.h, int(remove))

    def step_pre(self):
        self.L.vvr_step_pre(self.h)

    def step_hot(self):
        t = np.zeros(6)
        self.L.vvr_step_hot(self.h, t.ctypes.data_as(C.c_void_p))
        return t

    @property
    def time(self):
        return self.L.vvr_time(self.h)
