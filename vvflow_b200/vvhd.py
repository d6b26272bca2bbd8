"""Host-side mirror of the reference's class surface for the hot path.

Same class names, method names, argument meaning and error behaviour as libvvhd, so that a
step loop written against the reference (utils/vvflow/vvflow.cpp:198-266) reads the same here:

    tr = TSortedTree(S, 8, min_node_size, max_node_size)     # TSortedTree.hpp:60-92
    convective = MConvectiveFast(S, tr)                      # MConvectiveFast.hpp:8-27
    epsilon = MEpsilonFast(S, tr)                            # MEpsilonFast.hpp:5-31
    diffusive = MDiffusiveFast(S, tr)                        # MDiffusiveFast.hpp:5-18
    flowmove = MFlowmove(S)                                  # MFlowmove.hpp:5-19
    ...
    tr.build(); epsilon.CalcEpsilonFast(True); convective.process_all_lists()
    diffusive.process_vort_list(); tr.destroy(); flowmove.move_and_clean(True)

All arithmetic happens in libvvgpu.so (CUDA); this module only marshals. `Space.VortexList` is an
(n,6) float64 array of TObj records (x y g vx vy _1_eps); between build() and move_and_clean()
the authoritative copy lives on the device and is fetched lazily when the attribute is read.
The C++ twin of this file is vvflow_b200/host/vvgpu_adapter.hpp.
"""
import numpy as np

from . import capi

DBL_MAX = float(np.finfo(np.float64).max)


class TBody:
    """The TBody/TAtt state the hot path reads (libvvhd/headers/TBody.hpp)."""

    def __init__(self, corners, slip=None):
        c = np.asarray(corners, dtype=np.float64).reshape(-1, 2)
        n = c.shape[0]
        nxt = np.roll(c, -1, axis=0)
        self.corner = c
        self.dl = nxt - c                       # doUpdateSegments, TBody.cpp:200-215
        self.r = 0.5 * (nxt + c)
        self.ieps = 3.0 / np.sqrt(self.dl[:, 0] * self.dl[:, 0] + self.dl[:, 1] * self.dl[:, 1])
        self.g = np.zeros(n)
        self.gsum = np.zeros(n)
        self.fric = np.zeros(n)
        self.slip = np.zeros(n, dtype=np.int32) if slip is None else np.asarray(slip, dtype=np.int32)
        self.axis = np.zeros(2)                 # holder.r + dpos.r
        self.speed_slae = np.zeros(3)
        self.fdt_dead = np.zeros(3)
        self.g_dead = 0.0
        self._fill_properties()

    def _fill_properties(self):
        """doFillProperties (TBody.cpp:283-343): slen, area, centre of mass, bounding rect/disc."""
        c, dl, r = self.corner, self.dl, self.r
        n = c.shape[0]
        self.slen = 0.0
        area = 0.0
        for i in range(n):
            self.slen += float(np.hypot(dl[i, 0], dl[i, 1]))
            area += r[i, 1] * dl[i, 0]
        self.area = area
        nxt = np.roll(c, -1, axis=0)
        if area == 0:
            L = np.zeros(2)
            for i in range(n):
                L += r[i] * float(np.hypot(dl[i, 0], dl[i, 1]))
            self.cofm = L * (1.0 / self.slen)
        else:
            S3 = np.zeros(2)
            for i in range(n):
                cross = -c[i, 1] * nxt[i, 0] + c[i, 0] * nxt[i, 1]  # rotl(corner) * next corner
                S3 -= r[i] * cross
            self.cofm = S3 * (1.0 / (3 * area))
        self.bl = c.min(axis=0)
        self.tr = c.max(axis=0)
        self.disc_r2 = float(np.max(np.sum((c - self.cofm) ** 2, axis=1)))
        self.inside_valid = bool(area <= 0)  # isInsideValid(), TBody.hpp:117

    @staticmethod
    def from_oracle(seg_rows, body_row):
        """Build from oracle/pyref.Ref.segments() rows of one body and Ref.body(b) (exact TBody state)."""
        seg_rows = np.asarray(seg_rows)
        b = TBody.__new__(TBody)
        b.r = seg_rows[:, 0:2].copy(); b.corner = seg_rows[:, 2:4].copy(); b.dl = seg_rows[:, 4:6].copy()
        b.g = seg_rows[:, 6].copy(); b.gsum = seg_rows[:, 7].copy(); b.fric = seg_rows[:, 8].copy()
        b.ieps = seg_rows[:, 9].copy(); b.slip = seg_rows[:, 10].astype(np.int32)
        b.axis = np.array(body_row[0:2]); b.cofm = np.array(body_row[2:4])
        b.bl = np.array(body_row[4:6]); b.tr = np.array(body_row[6:8]); b.disc_r2 = float(body_row[8])
        b.inside_valid = bool(body_row[9]); b.speed_slae = np.array(body_row[10:13])
        b.fdt_dead = np.zeros(3); b.g_dead = 0.0
        b.slen = float(np.sum(np.hypot(b.dl[:, 0], b.dl[:, 1])))
        return b

    def size(self):
        return self.r.shape[0]


class Space:
    """Space (libvvhd/headers/TSpace.hpp:19-149): the state container and marshalling boundary."""

    def __init__(self, device=0, ctx=None):
        self.ctx = ctx if ctx is not None else capi.Context(device)
        self._vl = np.zeros((0, 6))
        self._dev_newer = False
        self.BodyList = []
        self.SourceList = np.zeros((0, 3))
        self.re = float("inf")
        self.dt = 1.0
        self.inf_vx = 0.0
        self.inf_vy = 0.0

    @property
    def VortexList(self):
        if self._dev_newer:
            self._vl = self.ctx.get_particles()
            self._dev_newer = False
        return self._vl

    @VortexList.setter
    def VortexList(self, rec):
        a = np.asarray(rec, dtype=np.float64)
        if a.ndim == 2 and a.shape[1] == 3:
            a = np.concatenate([a, np.zeros((a.shape[0], 3))], axis=1)
        self._vl = np.ascontiguousarray(a.reshape(-1, 6))
        self._dev_newer = False

    def average_segment_length(self):  # TSpace.hpp:119-129
        if not self.BodyList:
            return 0.0
        b = self.BodyList[0]
        if b.size() <= 1:
            return 0.0
        return b.slen / (b.size() - 1)

    def _pack_bodies(self):
        nseg = sum(b.size() for b in self.BodyList)
        segs = np.zeros(nseg, dtype=capi.SEG_DTYPE)
        bodies = np.zeros(len(self.BodyList), dtype=capi.BODY_DTYPE)
        k = 0
        for ib, b in enumerate(self.BodyList):
            n = b.size()
            s = segs[k:k + n]
            s["rx"], s["ry"] = b.r[:, 0], b.r[:, 1]
            s["cx"], s["cy"] = b.corner[:, 0], b.corner[:, 1]
            s["dlx"], s["dly"] = b.dl[:, 0], b.dl[:, 1]
            s["g"], s["ieps"], s["slip"], s["body"] = b.g, b.ieps, b.slip, ib
            B = bodies[ib]
            B["axis_x"], B["axis_y"] = b.axis
            B["cofm_x"], B["cofm_y"] = b.cofm
            B["bl_x"], B["bl_y"] = b.bl
            B["tr_x"], B["tr_y"] = b.tr
            B["disc_r2"] = b.disc_r2
            B["speed_x"], B["speed_y"], B["speed_o"] = b.speed_slae
            B["inside_valid"] = int(b.inside_valid)
            B["first_seg"], B["n_seg"] = k, n
            k += n
        return segs, bodies


class TSortedTree:
    """stree (libvvhd/headers/TSortedTree.hpp:60-92)."""

    def __init__(self, S, farCriteria, minNodeSize, maxNodeSize=DBL_MAX):
        self.S, self.farCriteria, self.minNodeSize, self.maxNodeSize = S, farCriteria, minNodeSize, maxNodeSize
        self.built = False

    def build(self, IncludeVortexes=True, IncludeBody=True, IncludeHeat=True):
        if self.built:  # TSortedTree.cpp:234
            import sys
            print("Tree is already built", file=sys.stderr)
            return
        S = self.S
        if not S._dev_newer:
            S.ctx.set_particles(S._vl)
        segs, bodies = S._pack_bodies()
        S.ctx.set_bodies(segs, bodies)
        mask = (1 if IncludeVortexes else 0) | (2 if IncludeBody else 0)
        S.ctx.tree_build(self.farCriteria, self.minNodeSize, self.maxNodeSize, mask)
        S._dev_newer = True  # the list is permuted in place, like the reference's
        self.built = True

    def destroy(self):
        self.S.ctx.tree_destroy()
        self.built = False

    def getBottomNodes(self):
        """Leaf table in bottomNodes order: rows of (x y h w vfirst vlast nseg)."""
        if not self.built:  # TSortedTree.cpp:277-281
            import sys
            print("PANIC in stree::getBottomNodes()! Tree isn't built", file=sys.stderr)
            return np.zeros((0, 7))
        dbl, idx, nl = self.S.ctx.tree_export()
        leaf = idx[:, 5] >= 0
        order = np.argsort(idx[leaf, 5])
        return np.concatenate([dbl[leaf][:, :4], idx[leaf][:, :3].astype(np.float64)], axis=1)[order]

    def findNode(self, p):
        """Pre-order id of the leaf containing p (stree::findNode, TSortedTree.cpp:284-303)."""
        if not self.built:
            raise ValueError("TTree::findNode(): tree is not built")
        dbl, idx, _ = self.S.ctx.tree_export()
        n = 0
        while idx[n, 3] >= 0:
            if dbl[n, 2] < dbl[n, 3]:
                n = idx[n, 3] if p[0] < dbl[n, 0] else idx[n, 4]
            else:
                n = idx[n, 3] if p[1] < dbl[n, 1] else idx[n, 4]
        return int(n)


class _Module:
    def __init__(self, S, tree):
        self.S, self.tree = S, tree

    def _need_tree(self, who):
        if not self.tree.built:
            raise RuntimeError(f"{who}: tree is not built")


class MEpsilonFast(_Module):
    def __init__(self, S, tree):
        super().__init__(S, tree)
        self.merged_ = 0

    def CalcEpsilonFast(self, merge):
        self._need_tree("MEpsilonFast::CalcEpsilonFast")
        self.merged_ = self.S.ctx.epsilon(bool(merge))
        self.S._dev_newer = True

    def Merged(self):
        return self.merged_

    def eps2h(self, p):
        """static MEpsilonFast::eps2h(node, p) with node = findNode(p), MEpsilonFast.cpp:66-93; p is (x, y) or (n, 2)"""
        self._need_tree("TTree::findNode()")
        pts = np.asarray(p, dtype=np.float64)
        r = self.S.ctx.eps2h_h2_at(pts.reshape(-1, 2))[:, 0]
        return r[0] if pts.ndim == 1 else r

    def h2(self, p):
        """static MEpsilonFast::h2(node, p) with node = findNode(p), MEpsilonFast.cpp:95-107"""
        self._need_tree("TTree::findNode()")
        pts = np.asarray(p, dtype=np.float64)
        r = self.S.ctx.eps2h_h2_at(pts.reshape(-1, 2))[:, 1]
        return r[0] if pts.ndim == 1 else r


class MConvectiveFast(_Module):
    def process_all_lists(self):
        self._need_tree("MConvectiveFast::process_all_lists")
        S = self.S
        S.ctx.convective(S.inf_vx, S.inf_vy, S.dt, S.SourceList if len(S.SourceList) else None)
        S._dev_newer = True

    def NodeInfluence(self):
        """double MConvectiveFast::NodeInfluence(const TSortedNode&, const TAtt&) const (MConvectiveFast.cpp:398-418)
        evaluated for Node = findNode(seg.r) of EVERY segment in BodyList order: the free vortices' term that
        fillSlipEquationForSegment (:459-467) subtracts from the right-hand side."""
        self._need_tree("TTree::findNode()")
        return self.S.ctx.node_influence()

    def velocity(self, p):
        """TVec MConvectiveFast::velocity(TVec p) const, MConvectiveFast.cpp:20-34; `p` is (x, y) or an (n, 2)
        array of points (the X* evaluators call it per raster point). Raises like stree::findNode when the tree
        is not built."""
        self._need_tree("TTree::findNode()")
        S = self.S
        pts = np.asarray(p, dtype=np.float64)
        v = S.ctx.velocity_at(pts.reshape(-1, 2), S.inf_vx, S.inf_vy, S.dt, S.SourceList if len(S.SourceList) else None)
        return v[0] if pts.ndim == 1 else v


class MDiffusiveFast(_Module):
    def process_vort_list(self):
        self._need_tree("MDiffusiveFast::process_vort_list")
        S = self.S
        fric = S.ctx.diffusive(S.re)
        if fric is not None:
            k = 0
            for b in S.BodyList:
                b.fric += fric[k:k + b.size()]
                k += b.size()
        S._dev_newer = True


class MFlowmove:
    def __init__(self, S, remove_eps=1e-10):
        self.S, self.remove_eps = S, remove_eps

    def move_and_clean(self, remove, collision=(), dt_eff=None):
        """Particle part of MFlowmove::move_and_clean. `collision` mirrors the reference's
        out-pointer: passing None raises like the reference's std::invalid_argument."""
        if collision is None:
            raise ValueError("MFlowmove::move_and_clean(): invalid collision pointer")
        S = self.S
        if not S._dev_newer:
            S.ctx.set_particles(S._vl)
        out = S.ctx.move_and_clean(S.dt if dt_eff is None else dt_eff, self.remove_eps, bool(remove))
        k = 0
        for ib, b in enumerate(S.BodyList):
            b.fdt_dead += out["fdt_dead"][ib]
            b.g_dead += out["g_dead"][ib]
            b.gsum += out["gsum"][k:k + b.size()]
            k += b.size()
        S._dev_newer = True
        return out["cleaned"]


class XVorticity:
    """XVorticity (libvvhd/headers/XVorticity.hpp, src/XVorticity.cpp): the vorticity raster of vvplot. Like the
    reference it works on a copy of the Space: the attached vortices are shed into the copy (MFlowmove::vortex_shed,
    MFlowmove.cpp:217-235), a tree with minNodeSize = 20 dl is built for it, and the Space itself is left untouched."""

    def __init__(self, S, xmin, ymin, dxdy, xres, yres):
        self.S = S
        self.xmin, self.ymin, self.dxdy = np.float32(xmin), np.float32(ymin), np.float32(dxdy)   # XField keeps floats
        self.xres, self.yres = int(xres), int(yres)
        self.eps_mult = 0.0
        self.map = None

    def evaluate(self, remove_eps=1e-10):
        if self.eps_mult <= 0:
            raise ValueError("XVorticity(): eps_mult must be positive")   # XVorticity.cpp:28-29
        if self.map is not None:
            return
        S = self.S
        own = S.VortexList.copy()                  # pulls the device copy if it is newer
        shed = []
        for b in S.BodyList:                       # vortex_shed on the copy (gsum is not touched here)
            keep = (np.abs(b.g) >= remove_eps) & (b.slip == 0)
            rec = np.zeros((int(keep.sum()), 6))
            rec[:, 0] = b.corner[keep, 0] - b.dl[keep, 1] * 1e-4   # corner + rotl(dl) * 1e-4
            rec[:, 1] = b.corner[keep, 1] + b.dl[keep, 0] * 1e-4
            rec[:, 2] = b.g[keep]
            shed.append(rec)
        S.ctx.set_particles(np.concatenate([own] + shed) if shed else own)
        segs, bodies = S._pack_bodies()
        S.ctx.set_bodies(segs, bodies)
        m = S.ctx.vorticity_raster(float(self.xmin), float(self.ymin), float(self.dxdy), self.xres, self.yres,
                                   float(self.eps_mult), S.average_segment_length())
        S.ctx.set_particles(own)                   # the Space's own list goes back to the device
        S._dev_newer = False
        self.map = m.astype(np.float32)

    def at(self, xi, yj):
        return self.map[yj, xi]


class XPressure:
    """XPressure (libvvhd/headers/XPressure.hpp, src/XPressure.cpp): the pressure raster of vvplot. Like the reference it
    works on a copy of the Space: the attached vortices are shed into the copy (gsum += g, MFlowmove.cpp:217-235), a tree
    with (8, 20 dl, 0.1) is built, one velocity pass runs (epsilon without merging, convective, diffusive) and the
    pressure is summed per raster point; the Space itself is left untouched."""

    def __init__(self, S, xmin, ymin, dxdy, xres, yres):
        self.S = S
        self.xmin, self.ymin, self.dxdy = np.float32(xmin), np.float32(ymin), np.float32(dxdy)   # XField keeps floats
        self.xres, self.yres = int(xres), int(yres)
        self.eps_mult = 0.0
        self.ref_frame = "s"
        self.map = None

    def evaluate(self, remove_eps=1e-10):
        if self.eps_mult <= 0:
            raise ValueError("XPressure(): eps_mult must be positive")     # XPressure.cpp:34-35
        S = self.S
        if self.ref_frame == "s":
            ref = None
        elif self.ref_frame == "o":
            ref = (0.0, 0.0)
        elif self.ref_frame == "f":
            ref = (S.inf_vx, S.inf_vy)
        elif self.ref_frame == "b":
            ref = (float(S.BodyList[0].speed_slae[0]), float(S.BodyList[0].speed_slae[1]))
        else:
            raise ValueError("XPressure(): bad ref_frame")                 # :50-52
        if self.map is not None:
            return
        own = S.VortexList.copy()                  # pulls the device copy if it is newer
        shed, gsum = [], []
        for b in S.BodyList:                       # vortex_shed on the copy
            keep = (np.abs(b.g) >= remove_eps) & (b.slip == 0)
            rec = np.zeros((int(keep.sum()), 6))
            rec[:, 0] = b.corner[keep, 0] - b.dl[keep, 1] * 1e-4   # corner + rotl(dl) * 1e-4
            rec[:, 1] = b.corner[keep, 1] + b.dl[keep, 0] * 1e-4
            rec[:, 2] = b.g[keep]
            shed.append(rec)
            gsum.append(b.gsum + b.g)              # gsum is incremented for slip segments too (:223-227)
        S.ctx.set_particles(np.concatenate([own] + shed) if shed else own)
        segs, bodies = S._pack_bodies()
        S.ctx.set_bodies(segs, bodies)
        m = S.ctx.pressure_raster(float(self.xmin), float(self.ymin), float(self.dxdy), self.xres, self.yres,
                                  S.average_segment_length(), S.re, S.dt, S.inf_vx, S.inf_vy,
                                  np.concatenate(gsum) if gsum else None, S.SourceList if len(S.SourceList) else None, ref)
        S.ctx.set_particles(own)                   # the Space's own list goes back to the device
        S._dev_newer = False
        self.map = m.astype(np.float32)

    def at(self, xi, yj):
        return self.map[yj, xi]

