// SURVEY §8(f) row 4 — velocity at arbitrary points: MConvectiveFast::velocity (libvvhd/src/MConvectiveFast.cpp:20-34),
// used by the reference for sensors and the vvplot field rasters. For a point p:
//   leaf  = stree::findNode(p)                       (TSortedTree.cpp:284-303: descend the split tests)
//   near  = sum over the leaf's NearNodes particles   (near_nodes_influence + biot_savart, :116-137, source eps)
//   far   = the leaf's FarNodes as eps = 0 monopoles AT p (far_nodes_influence, :139-151) — not the Taylor
//           expansion about the leaf centre that the step loop uses
//   + sink_list_influence(p) + body_list_influence(p) + inf_speed.
// One warp per point. The near/far classification of the point's leaf is recomputed by a walk with the
// reference's exact criterion (snode::FindNearNodes, TSortedTree.cpp:199-217): lanes pop up to 32 frontier nodes
// from a per-warp stack; a far node adds its two monopoles in the lane that tested it, a near internal node pushes
// its children, a near leaf is streamed particle by particle. Sums are reduced across lanes at the end, so the
// result agrees with the reference to rounding (order of summation), not bit for bit.
#pragma once
#include "vvgpu_near.cuh"

namespace vv {

constexpr int kPtWarps = 4;
constexpr int kPtStack = 1024;

struct PointArgs {
    TreeDev T;
    Particles P;
    int npts;
    const double* xy;     // npts x 2
    double* out;          // npts x 2
    double farc, inf_vx, inf_vy, eps2_div_srcg;
    const double* sinks;  // (x, y, g) triples
    int nsink;
    BodyFull B;
    int body_flow;        // some body has slip segments or moves
    int* err;
};

__global__ void __launch_bounds__(kPtWarps * 32) k_velocity_at(PointArgs A) {
    __shared__ int stack[kPtWarps][kPtStack];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kPtWarps + warp;
    if (q >= A.npts) return;   // warps are independent: no block-level barrier below
    const TreeDev& T = A.T;
    const double px = A.xy[2 * q], py = A.xy[2 * q + 1];
    // ---- stree::findNode
    int leaf = 0;
    while (T.ch1[leaf] >= 0) {
        const int c = T.ch1[leaf];
        leaf = T.axis[leaf] ? ((px < T.x[leaf]) ? c : c + 1) : ((py < T.y[leaf]) ? c : c + 1);
    }
    const double lcx = T.x[leaf], lcy = T.y[leaf], lh = T.h[leaf], lw = T.w[leaf];
    int* st = stack[warp];
    if (lane == 0) st[0] = 0;
    int size = 1;
    __syncwarp();
    double nx = 0, ny = 0, fx = 0, fy = 0;   // near / far partial sums of this lane
    while (size > 0) {
        const int take = (size > kPtStack - 80) ? 1 : min(size, 32);
        int n = -1;
        if (lane < take) n = st[size - 1 - lane];
        size -= take;
        __syncwarp();
        bool push = false, nearleaf = false;
        int c1 = -1;
        if (n >= 0) {
            const double tx = T.x[n], ty = T.y[n];
            const double nhw = VV_ADD(T.h[n], T.w[n]);
            c1 = T.ch1[n];
            if (is_far(tx, ty, nhw, lcx, lcy, lh, lw, A.farc)) {
                const double* Pm = T.cmp + 3ll * n;
                const double* Mm = T.cmm + 3ll * n;
                double dx = px - Pm[0], dy = py - Pm[1];
                double w = Pm[2] / (dx * dx + dy * dy);
                fx += -dy * w; fy += dx * w;
                dx = px - Mm[0]; dy = py - Mm[1];
                w = Mm[2] / (dx * dx + dy * dy);
                fx += -dy * w; fy += dx * w;
            } else if (c1 >= 0) push = true;
            else nearleaf = true;
        }
        const u32 pb = __ballot_sync(0xffffffffu, push);
        const int npush = 2 * __popc(pb);
        if (size + npush > kPtStack) {
            if (lane == 0) atomicOr(A.err, 2);
            return;
        }
        if (push) {
            const int off = size + 2 * __popc(pb & lanemask_lt());
            st[off] = c1 + 1; st[off + 1] = c1;
        }
        size += npush;
        // near leaves of this batch: every lane takes particles of each in turn
        for (u32 nb = __ballot_sync(0xffffffffu, nearleaf); nb; nb &= nb - 1) {
            const int src = __ffs(nb) - 1;
            const int ln = __shfl_sync(0xffffffffu, n, src);
            const int f = T.first[ln], l = T.last[ln];
            for (int j = f + lane; j < l; j += 32) {
                const double g = A.P.g[j];
                if (g == 0) continue;   // `if (!lobj->g) continue`, :130
                const double dx = px - A.P.x[j], dy = py - A.P.y[j];
                const double e = 1. / A.P.ie[j];
                const double w = g / (dx * dx + dy * dy + e * e);
                nx += -dy * w; ny += dx * w;
            }
        }
        __syncwarp();
    }
    // ---- sinks and bodies, lanes strided over the sources
    double sx = 0, sy = 0;
    for (int k = lane; k < A.nsink; k += 32) {   // sink_list_influence, :153-170
        const double dx = px - A.sinks[3 * k], dy = py - A.sinks[3 * k + 1], sg = A.sinks[3 * k + 2];
        const double w = sg / (dx * dx + dy * dy + A.eps2_div_srcg * fabs(sg));
        sx += dx * w; sy += dy * w;
    }
    double bx = 0, by = 0;
    if (A.body_flow) {
        for (int ib = 0; ib < A.B.nbody; ib++) {
            const double* bp = A.B.bprop + 16 * ib;
            const int f = A.B.bfirst[ib], e = A.B.bfirst[ib + 1];
            if (bp[13] != 0)
                for (int s = f + lane; s < e; s += 32) body_slip_term(A.B, s, px, py, bx, by);
            if (body_moves(bp))
                for (int s = f + lane; s < e; s += 32) body_motion_term(A.B, bp, s, px, py, bx, by);
        }
    }
    double vx = (nx + fx + sx + bx) * k1_2Pi, vy = (ny + fy + sy + by) * k1_2Pi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vx += __shfl_xor_sync(0xffffffffu, vx, o);
        vy += __shfl_xor_sync(0xffffffffu, vy, o);
    }
    if (lane == 0) { A.out[2 * q] = vx + A.inf_vx; A.out[2 * q + 1] = vy + A.inf_vy; }
}

}  // namespace vv
