"""ctypes binding of include/vvgpu.h (vvflow_b200/lib/libvvgpu.so).

This is the thinnest possible layer: numpy arrays in, numpy arrays out, every call checked.
There is no fallback: if the library is missing or CUDA is unavailable the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VVGPU_LIB") or os.path.join(HERE, "lib", "libvvgpu.so")   # VVGPU_LIB: an experimental build

# every symbol include/vvgpu.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vvgpu_create", "vvgpu_destroy", "vvgpu_strerror", "vvgpu_last_error",
    "vvgpu_set_particles", "vvgpu_set_particles_xyg", "vvgpu_append_particles", "vvgpu_particle_count", "vvgpu_particle_gsum", "vvgpu_get_particles", "vvgpu_get_particles_range",
    "vvgpu_get_permutation", "vvgpu_set_bodies",
    "vvgpu_tree_build", "vvgpu_tree_destroy", "vvgpu_tree_counts", "vvgpu_tree_export", "vvgpu_tree_lists",
    "vvgpu_tree_leaf_segments", "vvgpu_count_interactions",
    "vvgpu_epsilon", "vvgpu_merge_rounds", "vvgpu_convective", "vvgpu_velocity_at", "vvgpu_eps2h_h2_at", "vvgpu_node_influence", "vvgpu_vorticity_raster", "vvgpu_pressure_raster", "vvgpu_diffusive", "vvgpu_move_and_clean",
    "vvgpu_comm_unique_id", "vvgpu_comm_init", "vvgpu_group_create", "vvgpu_comm_info", "vvgpu_sync_ranks", "vvgpu_shard_owner", "vvgpu_shard_block",
    "vvgpu_set_particles_slice", "vvgpu_particle_arrays_dev", "vvgpu_stream",
    "vvgpu_synchronize", "vvgpu_phase_times", "vvgpu_host_syncs", "vvgpu_fp64_peak",
]

SEG_DTYPE = np.dtype([("rx", "f8"), ("ry", "f8"), ("cx", "f8"), ("cy", "f8"), ("dlx", "f8"), ("dly", "f8"),
                      ("g", "f8"), ("ieps", "f8"), ("slip", "i4"), ("body", "i4")], align=True)
BODY_DTYPE = np.dtype([("axis_x", "f8"), ("axis_y", "f8"), ("cofm_x", "f8"), ("cofm_y", "f8"), ("bl_x", "f8"),
                       ("bl_y", "f8"), ("tr_x", "f8"), ("tr_y", "f8"), ("disc_r2", "f8"), ("speed_x", "f8"),
                       ("speed_y", "f8"), ("speed_o", "f8"), ("inside_valid", "i4"), ("first_seg", "i4"),
                       ("n_seg", "i4"), ("_pad", "i4")], align=True)
assert SEG_DTYPE.itemsize == 72 and BODY_DTYPE.itemsize == 112

PHASES = ("build", "lists", "eps", "conv", "diff", "move")

_lib = None


class VVGpuError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(text)
        self.code = code


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VVGpuError(-3, f"{LIB_PATH} is missing: build it with `python -m vvflow_b200.build` "
                             "(the CUDA extension is the only implementation; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
    sz = C.c_size_t
    sig = {
        "vvgpu_create": [C.c_int, C.POINTER(vp)],
        "vvgpu_set_particles": [vp, C.c_int, dp, sz],
        "vvgpu_set_particles_xyg": [vp, C.c_int, dp, sz],
        "vvgpu_append_particles": [vp, C.c_int, dp, sz],
        "vvgpu_particle_count": [vp, C.c_int, C.POINTER(sz)],
        "vvgpu_particle_gsum": [vp, C.c_int, C.POINTER(C.c_double)],
        "vvgpu_get_particles": [vp, C.c_int, dp, sz, C.POINTER(sz)],
        "vvgpu_get_particles_range": [vp, C.c_int, dp, sz, sz],
        "vvgpu_get_permutation": [vp, C.c_int, ip, sz],
        "vvgpu_set_bodies": [vp, vp, sz, vp, sz],
        "vvgpu_tree_build": [vp, C.c_int, C.c_double, C.c_double, C.c_uint],
        "vvgpu_tree_destroy": [vp],
        "vvgpu_tree_counts": [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)],
        "vvgpu_tree_export": [vp, dp, ip, sz],
        "vvgpu_tree_lists": [vp, ip, ip, sz, ip, ip, sz],
        "vvgpu_tree_leaf_segments": [vp, ip, ip, sz],
        "vvgpu_count_interactions": [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "vvgpu_epsilon": [vp, C.c_int, C.POINTER(C.c_int)],
        "vvgpu_merge_rounds": [vp, C.POINTER(C.c_int)],
        "vvgpu_convective": [vp, C.c_double, C.c_double, C.c_double, dp, sz],
        "vvgpu_velocity_at": [vp, dp, sz, C.c_double, C.c_double, C.c_double, dp, sz, dp],
        "vvgpu_eps2h_h2_at": [vp, dp, sz, dp],
        "vvgpu_node_influence": [vp, dp],
        "vvgpu_vorticity_raster": [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_double, C.c_double, dp],
        "vvgpu_pressure_raster": [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                  C.c_double, C.c_double, dp, sz, dp, C.c_int, C.c_double, C.c_double, dp],
        "vvgpu_diffusive": [vp, C.c_double, dp],
        "vvgpu_move_and_clean": [vp, C.c_double, C.c_double, C.c_int, dp, dp, dp, C.POINTER(sz)],
        "vvgpu_comm_unique_id": [vp, sz],
        "vvgpu_comm_init": [vp, C.c_int, C.c_int, vp],
        "vvgpu_group_create": [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)],
        "vvgpu_comm_info": [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "vvgpu_sync_ranks": [vp],
        "vvgpu_shard_owner": [C.c_int, C.c_int],
        "vvgpu_shard_block": [],
        "vvgpu_set_particles_slice": [vp, C.c_int, dp, sz, sz, sz],
        "vvgpu_particle_arrays_dev": [vp, C.c_int, C.POINTER(vp), C.POINTER(sz)],
        "vvgpu_stream": [vp, C.POINTER(vp)],
        "vvgpu_synchronize": [vp],
        "vvgpu_phase_times": [vp, dp, C.POINTER(C.c_uint64)],
        "vvgpu_host_syncs": [vp, C.POINTER(C.c_uint64)],
        "vvgpu_fp64_peak": [vp, C.POINTER(C.c_double)],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = C.c_int
    L.vvgpu_destroy.argtypes = [vp]
    L.vvgpu_destroy.restype = None
    L.vvgpu_strerror.argtypes = [C.c_int]
    L.vvgpu_strerror.restype = C.c_char_p
    L.vvgpu_last_error.argtypes = [vp]
    L.vvgpu_last_error.restype = C.c_char_p
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One vvgpu context = one CUDA device."""

    def __init__(self, device=0, handle=None):
        self.L = load()
        if handle is not None:      # a context made by vvgpu_group_create
            self.h = handle
            return
        h = C.c_void_p()
        rc = self.L.vvgpu_create(device, C.byref(h))
        if rc:
            raise VVGpuError(rc, f"vvgpu_create(device={device}): {self.L.vvgpu_strerror(rc).decode()} "
                                 "(a CUDA device is required; there is no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.vvgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise VVGpuError(rc, f"{self.L.vvgpu_strerror(rc).decode()}: {self.L.vvgpu_last_error(self.h).decode()}")

    # ---- state
    def set_particles(self, rec48):
        a = np.ascontiguousarray(rec48, dtype=np.float64).reshape(-1, 6)
        self._ck(self.L.vvgpu_set_particles(self.h, 0, _p(a), a.shape[0]))

    def append_particles(self, rec48):
        """append (n, 6) TObj records behind the resident particles (newly shed vortices)"""
        a = np.ascontiguousarray(rec48, dtype=np.float64).reshape(-1, 6)
        self._ck(self.L.vvgpu_append_particles(self.h, 0, _p(a), a.shape[0]))

    def set_particles_xyg(self, xyg):
        a = np.ascontiguousarray(xyg, dtype=np.float64).reshape(-1, 3)
        self._ck(self.L.vvgpu_set_particles_xyg(self.h, 0, _p(a), a.shape[0]))

    def set_particles_ptr(self, ptr, n):
        """48-byte records at a raw host address (e.g. a pinned buffer)."""
        self._ck(self.L.vvgpu_set_particles(self.h, 0, C.c_void_p(ptr), n))

    def get_particles_ptr(self, ptr, cap):
        n = C.c_size_t()
        self._ck(self.L.vvgpu_get_particles(self.h, 0, C.c_void_p(ptr), cap, C.byref(n)))
        return n.value

    def gsum(self):
        v = C.c_double()
        self._ck(self.L.vvgpu_particle_gsum(self.h, 0, C.byref(v)))
        return v.value

    def get_particles_range_ptr(self, ptr, first, count):
        self._ck(self.L.vvgpu_get_particles_range(self.h, 0, C.c_void_p(ptr), first, count))

    @property
    def n(self):
        n = C.c_size_t()
        self._ck(self.L.vvgpu_particle_count(self.h, 0, C.byref(n)))
        return n.value

    def get_particles(self):
        out = np.zeros((self.n, 6))
        n = C.c_size_t()
        self._ck(self.L.vvgpu_get_particles(self.h, 0, _p(out), out.shape[0], C.byref(n)))
        return out

    def get_permutation(self):
        out = np.zeros(self.n, dtype=np.int32)
        self._ck(self.L.vvgpu_get_permutation(self.h, 0, _p(out), out.shape[0]))
        return out

    def set_bodies(self, segs, bodies):
        segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
        bodies = np.ascontiguousarray(bodies, dtype=BODY_DTYPE)
        self.nseg, self.nbody = segs.shape[0], bodies.shape[0]
        self._ck(self.L.vvgpu_set_bodies(self.h, _p(segs), segs.shape[0], _p(bodies), bodies.shape[0]))

    # ---- tree
    def tree_build(self, far=8, min_node=0.0, max_node=np.finfo(np.float64).max, mask=3):
        self._ck(self.L.vvgpu_tree_build(self.h, far, min_node, max_node, mask))

    def tree_destroy(self):
        self._ck(self.L.vvgpu_tree_destroy(self.h))

    def tree_counts(self):
        a, b, d = C.c_size_t(), C.c_size_t(), C.c_size_t()
        self._ck(self.L.vvgpu_tree_counts(self.h, C.byref(a), C.byref(b), C.byref(d)))
        return a.value, b.value, d.value

    def tree_export(self):
        nn, nl, _ = self.tree_counts()
        dbl = np.zeros((nn, 10))
        idx = np.zeros((nn, 8), dtype=np.int64)
        self._ck(self.L.vvgpu_tree_export(self.h, _p(dbl), _p(idx), nn))
        return dbl, idx, nl

    def tree_lists(self):
        _, nl, _ = self.tree_counts()
        nptr = np.zeros(nl + 1, dtype=np.int64)
        fptr = np.zeros(nl + 1, dtype=np.int64)
        self._ck(self.L.vvgpu_tree_lists(self.h, _p(nptr), None, 0, _p(fptr), None, 0))
        nidx = np.zeros(max(1, nptr[-1]), dtype=np.int64)
        fidx = np.zeros(max(1, fptr[-1]), dtype=np.int64)
        self._ck(self.L.vvgpu_tree_lists(self.h, _p(nptr), _p(nidx), nidx.shape[0], _p(fptr), _p(fidx), fidx.shape[0]))
        return nptr, nidx[: nptr[-1]], fptr, fidx[: fptr[-1]]

    def tree_leaf_segments(self):
        _, nl, _ = self.tree_counts()
        ptr = np.zeros(nl + 1, dtype=np.int64)
        self._ck(self.L.vvgpu_tree_leaf_segments(self.h, _p(ptr), None, 0))
        idx = np.zeros(max(1, ptr[-1]), dtype=np.int64)
        self._ck(self.L.vvgpu_tree_leaf_segments(self.h, _p(ptr), _p(idx), idx.shape[0]))
        return ptr, idx[: ptr[-1]]

    def count_interactions(self):
        a, b = C.c_double(), C.c_double()
        self._ck(self.L.vvgpu_count_interactions(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- phases
    def epsilon(self, merge):
        m = C.c_int()
        self._ck(self.L.vvgpu_epsilon(self.h, int(merge), C.byref(m)))
        return m.value

    def merge_rounds(self):
        m = C.c_int()
        self._ck(self.L.vvgpu_merge_rounds(self.h, C.byref(m)))
        return m.value

    def convective(self, inf_vx=0.0, inf_vy=0.0, dt=0.0, sinks=None):
        s = None if sinks is None else np.ascontiguousarray(sinks, dtype=np.float64).reshape(-1, 3)
        self._ck(self.L.vvgpu_convective(self.h, inf_vx, inf_vy, dt, _p(s), 0 if s is None else s.shape[0]))

    def velocity_at(self, xy, inf_vx=0.0, inf_vy=0.0, dt=0.0, sinks=None):
        """MConvectiveFast::velocity at arbitrary points: (n, 2) -> (n, 2)"""
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        s = None if sinks is None else np.ascontiguousarray(sinks, dtype=np.float64).reshape(-1, 3)
        out = np.zeros_like(a)
        self._ck(self.L.vvgpu_velocity_at(self.h, _p(a), a.shape[0], inf_vx, inf_vy, dt, _p(s),
                                          0 if s is None else s.shape[0], _p(out)))
        return out

    def eps2h_h2_at(self, xy):
        """(MEpsilonFast::eps2h, MEpsilonFast::h2) of findNode(p): (n, 2) -> (n, 2)"""
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        out = np.zeros_like(a)
        self._ck(self.L.vvgpu_eps2h_h2_at(self.h, _p(a), a.shape[0], _p(out)))
        return out

    def node_influence(self):
        """MConvectiveFast::NodeInfluence(findNode(seg.r), seg) for every segment set with set_bodies"""
        out = np.zeros(max(1, getattr(self, "nseg", 0)))
        self._ck(self.L.vvgpu_node_influence(self.h, _p(out)))
        return out[: getattr(self, "nseg", 0)]

    def vorticity_raster(self, xmin, ymin, dxdy, xres, yres, eps_mult, dl):
        """XVorticity::evaluate on the resident (post-shed) list: (yres, xres) float64"""
        out = np.zeros((yres, xres))
        self._ck(self.L.vvgpu_vorticity_raster(self.h, xmin, ymin, dxdy, xres, yres, eps_mult, dl, _p(out)))
        return out

    def pressure_raster(self, xmin, ymin, dxdy, xres, yres, dl, re, dt, inf_vx, inf_vy, gsum, sinks=None, ref_speed=None):
        """XPressure::evaluate on the resident (post-shed) list: (yres, xres) float64; ref_speed None = ref_frame 's'"""
        out = np.zeros((yres, xres))
        s = None if sinks is None or len(sinks) == 0 else np.ascontiguousarray(sinks, dtype=np.float64).reshape(-1, 3)
        g = None if gsum is None or len(gsum) == 0 else np.ascontiguousarray(gsum, dtype=np.float64)
        rs = (0.0, 0.0) if ref_speed is None else ref_speed
        self._ck(self.L.vvgpu_pressure_raster(self.h, xmin, ymin, dxdy, xres, yres, dl, re, dt, inf_vx, inf_vy, _p(s),
                                              0 if s is None else s.shape[0], _p(g), 0 if ref_speed is None else 1,
                                              rs[0], rs[1], _p(out)))
        return out

    def diffusive(self, re, want_fric=True):
        fric = np.zeros(max(1, getattr(self, "nseg", 0))) if want_fric and getattr(self, "nseg", 0) else None
        self._ck(self.L.vvgpu_diffusive(self.h, re, _p(fric)))
        return fric

    def move_and_clean(self, dt, remove_eps=1e-10, remove=True):
        nb, ns = getattr(self, "nbody", 0), getattr(self, "nseg", 0)
        fdt = np.zeros(3 * max(1, nb)); gd = np.zeros(max(1, nb)); gs = np.zeros(max(1, ns))
        cl = C.c_size_t()
        self._ck(self.L.vvgpu_move_and_clean(self.h, dt, remove_eps, int(remove), _p(fdt), _p(gd), _p(gs), C.byref(cl)))
        return dict(fdt_dead=fdt[: 3 * nb].reshape(-1, 3), g_dead=gd[:nb], gsum=gs[:ns], cleaned=cl.value)

    # ---- multi-GPU / interop
    def comm_init(self, rank, nranks, unique_id):
        """one process per GPU: `unique_id` = the 128 bytes rank 0 got from comm_unique_id()"""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id)) if nranks > 1 else None
        self._ck(self.L.vvgpu_comm_init(self.h, rank, nranks, buf))

    def comm_info(self):
        r, n, k = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.vvgpu_comm_info(self.h, C.byref(r), C.byref(n), C.byref(k)))
        return r.value, n.value, k.value

    def set_particles_slice_ptr(self, ptr, first, count, n_total):
        self._ck(self.L.vvgpu_set_particles_slice(self.h, 0, C.c_void_p(ptr), first, count, n_total))

    def set_particles_slice(self, rec48, first, n_total):
        a = np.ascontiguousarray(rec48, dtype=np.float64).reshape(-1, 6)
        self._ck(self.L.vvgpu_set_particles_slice(self.h, 0, _p(a), first, a.shape[0], n_total))

    def arrays_dev(self):
        arr = (C.c_void_p * 6)()
        n = C.c_size_t()
        self._ck(self.L.vvgpu_particle_arrays_dev(self.h, 0, arr, C.byref(n)))
        return [arr[k] for k in range(6)], n.value

    def synchronize(self):
        self._ck(self.L.vvgpu_synchronize(self.h))

    def stream(self):
        s = C.c_void_p()
        self._ck(self.L.vvgpu_stream(self.h, C.byref(s)))
        return s.value

    def phase_times(self):
        ms = np.zeros(len(PHASES))
        la = C.c_uint64()
        self._ck(self.L.vvgpu_phase_times(self.h, _p(ms), C.byref(la)))
        return dict(zip(PHASES, ms.tolist())), la.value

    def host_syncs(self):
        n = C.c_uint64()
        self._ck(self.L.vvgpu_host_syncs(self.h, C.byref(n)))
        return n.value

    def fp64_peak(self):
        t = C.c_double()
        self._ck(self.L.vvgpu_fp64_peak(self.h, C.byref(t)))
        return t.value


def comm_unique_id():
    """ncclGetUniqueId through the library (rank 0); carry the bytes to the other ranks with any host transport"""
    L = load()
    buf = (C.c_char * 128)()
    rc = L.vvgpu_comm_unique_id(buf, 128)
    if rc:
        raise VVGpuError(rc, "vvgpu_comm_unique_id: libnccl.so.2 is not usable")
    return bytes(buf)


def shard_owner(group, nranks):
    return load().vvgpu_shard_owner(group, nranks)


def shard_block():
    return load().vvgpu_shard_block()


def group_create(devices):
    """one process, several ranks (devices may repeat): a list of Contexts, each to be driven from its own thread"""
    L = load()
    n = len(devices)
    dev = (C.c_int * n)(*devices)
    hs = (C.c_void_p * n)()
    rc = L.vvgpu_group_create(dev, n, hs)
    if rc:
        raise VVGpuError(rc, f"vvgpu_group_create({list(devices)}): {L.vvgpu_strerror(rc).decode()}")
    return [Context(handle=C.c_void_p(hs[k])) for k in range(n)]
