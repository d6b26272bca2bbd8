// K1, single launch — the whole bit-exact tree build of vvgpu_tree.cuh as ONE persistent cooperative
// kernel (one 1024-thread CTA per SM, grid-wide barriers between the phases of a level).
//
// The level-by-level version needs ~12 launches and one host read-back per tree level (22 levels at
// N = 1M: ~260 launches, 3.0 ms of which well under 1 ms is kernel work — profiles/r1_launches_*).
// Here the host launches once and reads back (nodes, depth, leaves) once; per level the grid passes
// six barriers:
//   P1 decide        box -> centre/size, leaf tests (DivideNode, TSortedTree.cpp:36-56)
//   P2 tile sums     of the three flag arrays (node splits, particle "coord < mid", segment ditto)
//   P3 scans         node ranks -> children allocated; particle / segment prefix sums
//   P4 partition     Hoare partner table (closed form of :81-99), child ranges; stable segment split
//   P5 swap+relabel  in-place swap of (x, y, g, perm) and the particles' new node ids
//   P6 stretch       tight boxes of the new children (Stretch, :101-137)
// followed by the bottom-up (centres of mass, subtree sizes) and top-down (DFS ids) sweeps.
// The arithmetic is the same device code as the per-level kernels (k_tree_* stay for reference and
// for the unit tests of the individual steps).
#pragma once
#include "vvgpu_tree.cuh"

#include <cooperative_groups.h>

namespace vv {
namespace cg = cooperative_groups;

constexpr int kCoopThreads = 1024;
constexpr int kCoopItems = 4;
constexpr int kCoopTile = kCoopThreads * kCoopItems;
constexpr int kCoopMaxDepth = 4096;

struct BuildState {
    int nnodes, depth, nleaves, err;
    int lvl[kCoopMaxDepth + 2];
};

struct CoopArgs {
    TreeDev T;
    BuildParams bp;
    double *px, *py, *pg;
    const double *sx, *sy;
    int n, nseg;
    int *perm, *tmpR;
    int* segperm[2];
    int* snode[2];
    u32 *G, *Gs, *splitflag;      // particle scan (n+1), segment scan (nseg+1), node flags of one level
    u32 *partN, *partP, *partS;   // tile sums
    BuildState* st;
    long long cap;                // node capacity
};

// sum of f over this CTA's tiles -> partial[tile]
template <class F>
__device__ __forceinline__ void coop_tile_sums(F f, long long n, u32* partial, u32* sh) {
    const long long ntiles = (n + kCoopTile - 1) / kCoopTile;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long base = t * kCoopTile + (long long)threadIdx.x * kCoopItems;
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < kCoopItems; k++)
            if (base + k < n) s += f(base + k);
        u32 total;
        block_exclusive_scan<kCoopThreads>(s, &total, sh);
        if (threadIdx.x == 0) partial[t] = total;
    }
}
// block-wide sum of partial[0..upto)
__device__ __forceinline__ u32 coop_prefix(const u32* partial, long long upto, u32* sh) {
    u32 s = 0;
    for (long long k = threadIdx.x; k < upto; k += kCoopThreads) s += partial[k];
    u32 total;
    block_exclusive_scan<kCoopThreads>(s, &total, sh);
    return total;
}
// exclusive scan of f over [0,n): emit(i, exclusive prefix) for every i, emit_total(total) once
template <class F, class E>
__device__ __forceinline__ void coop_scan_apply(F f, long long n, const u32* partial, u32* sh, E emit) {
    const long long ntiles = (n + kCoopTile - 1) / kCoopTile;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const u32 before = coop_prefix(partial, t, sh);
        const long long base = t * kCoopTile + (long long)threadIdx.x * kCoopItems;
        u32 v[kCoopItems];
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < kCoopItems; k++) {
            v[k] = (base + k < n) ? f(base + k) : 0;
            s += v[k];
        }
        u32 total;
        u32 ex = block_exclusive_scan<kCoopThreads>(s, &total, sh) + before;
#pragma unroll
        for (int k = 0; k < kCoopItems; k++) {
            if (base + k < n) emit(base + k, ex, ex + v[k]);
            ex += v[k];
        }
    }
}

// Stretch of freshly created nodes: fold every object into the box of its (pending) node
__device__ __forceinline__ void coop_bbox(const TreeDev& T, const int* snode, const double* px, const double* py, int n,
                                          const double* sx, const double* sy, const int* seg_perm, int nseg,
                                          long long gtid, long long gsize) {
    const long long total = (long long)n + nseg;
    const long long rounds = (total + gsize - 1) / gsize;
    for (long long r = 0; r < rounds; r++) {   // whole warps stay together for the shuffles below
        const long long i = r * gsize + gtid;
        int node = -1;
        double x = 0, y = 0;
        if (i < n) { node = T.pnode[i]; x = px[i]; y = py[i]; }
        else if (i < total) { int k = (int)(i - n); node = snode[k]; int s = seg_perm[k]; x = sx[s]; y = sy[s]; }
        if (node >= 0 && T.status[node] != ST_PENDING) node = -1;
        const int n0 = __shfl_sync(0xffffffffu, node, 0);
        const bool uniform = __all_sync(0xffffffffu, node == n0);
        if (uniform) {
            if (n0 < 0) continue;
            u64 ex = enc_ordered(x), ey = enc_ordered(y);
            u64 mnx = ex, mxx = ex, mny = ey, mxy = ey;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                u64 t;
                t = __shfl_xor_sync(0xffffffffu, mnx, o); mnx = t < mnx ? t : mnx;
                t = __shfl_xor_sync(0xffffffffu, mxx, o); mxx = t > mxx ? t : mxx;
                t = __shfl_xor_sync(0xffffffffu, mny, o); mny = t < mny ? t : mny;
                t = __shfl_xor_sync(0xffffffffu, mxy, o); mxy = t > mxy ? t : mxy;
            }
            if ((threadIdx.x & 31) == 0) {
                u64* bb = T.bb + 4ll * n0;
                atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny);
                atomicMax(bb + 2, mxx); atomicMax(bb + 3, mxy);
            }
        } else if (node >= 0) {
            u64* bb = T.bb + 4ll * node;
            u64 ex = enc_ordered(x), ey = enc_ordered(y);
            atomicMin(bb + 0, ex); atomicMin(bb + 1, ey);
            atomicMax(bb + 2, ex); atomicMax(bb + 3, ey);
        }
    }
}

__global__ void __launch_bounds__(kCoopThreads, 1) k_tree_build_coop(CoopArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 sh[kCoopThreads / 32 + 1];
    TreeDev T = A.T;
    const int n = A.n, nseg = A.nseg;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsize = (long long)gridDim.x * blockDim.x;
    int segcur = 0;
    T.snode = A.snode[0];

    // ---- root
    if (gtid == 0) {
        T.first[0] = 0; T.last[0] = n; T.sfirst[0] = 0; T.slast[0] = nseg;
        T.ch1[0] = -1; T.parent[0] = -1; T.depth[0] = 0; T.status[0] = ST_PENDING;
        bb_reset(T.bb);
        A.st->lvl[0] = 0; A.st->lvl[1] = 1; A.st->err = 0;
    }
    for (long long i = gtid; i < n; i += gsize) { T.pnode[i] = 0; A.perm[i] = (int)i; }
    for (long long k = gtid; k < nseg; k += gsize) { A.snode[0][k] = 0; A.segperm[0][k] = (int)k; }
    grid.sync();
    coop_bbox(T, A.snode[0], A.px, A.py, n, A.sx, A.sy, A.segperm[0], nseg, gtid, gsize);
    grid.sync();

    int a0 = 0, a1 = 1, d = 0;
    bool failed = false;
    for (;; d++) {
        const int na = a1 - a0;
        // ---- P1: DivideNode's tests for the nodes of this level
        for (long long k = gtid; k < na; k += gsize) {
            const int nn = a0 + (int)k;
            const u64* bb = T.bb + 4ll * nn;
            double blx = dec_ordered(bb[0]), bly = dec_ordered(bb[1]), trx = dec_ordered(bb[2]), try_ = dec_ordered(bb[3]);
            double x = VV_MUL(VV_ADD(blx, trx), 0.5);   // :122-125
            double y = VV_MUL(VV_ADD(bly, try_), 0.5);
            double h = VV_SUB(try_, bly);
            double w = VV_SUB(trx, blx);
            T.x[nn] = x; T.y[nn] = y; T.h[nn] = h; T.w[nn] = w;
            bool leaf = false;
            double mx = std_max(h, w), mn = std_min(h, w);
            if (mx < A.bp.max_node && mn <= A.bp.min_node) leaf = true;
            if (!leaf) {
                int m = T.slast[nn] - T.sfirst[nn];
                int nv = T.last[nn] - T.first[nn];
                if (nv > m) m = nv;
                if (mx < A.bp.max_node && m < kTreeMaxList) leaf = true;
            }
            T.status[nn] = leaf ? ST_LEAF : ST_SPLIT;
            T.axis[nn] = (h < w) ? 1 : 0;  // :87
            A.splitflag[k] = leaf ? 0u : 1u;
        }
        grid.sync();
        // ---- P2: tile sums of the three flag arrays
        const int* segin = A.segperm[segcur];
        coop_tile_sums(FlagArray{A.splitflag}, na, A.partN, sh);
        coop_tile_sums(PartFlag{T, A.px, A.py}, n, A.partP, sh);
        if (nseg) coop_tile_sums(SegFlag{T, A.sx, A.sy, segin}, nseg, A.partS, sh);
        grid.sync();
        // ---- P3: scans. Every CTA derives the same split count, so the exit is grid-uniform.
        const u32 nsplit = coop_prefix(A.partN, ((long long)na + kCoopTile - 1) / kCoopTile, sh);
        if (nsplit == 0) break;
        if (d + 1 >= kCoopMaxDepth || (long long)a1 + 2ll * nsplit > A.cap) { failed = true; break; }
        coop_scan_apply(FlagArray{A.splitflag}, na, A.partN, sh, [&](long long k, u32 rank, u32) {
            const int nn = a0 + (int)k;
            if (T.status[nn] != ST_SPLIT) { T.ch1[nn] = -1; return; }
            const int c = a1 + 2 * (int)rank;
            T.ch1[nn] = c;
            for (int q = 0; q < 2; q++) {
                T.parent[c + q] = nn;
                T.depth[c + q] = T.depth[nn] + 1;
                T.status[c + q] = ST_PENDING;
                T.ch1[c + q] = -1;
                T.first[c + q] = T.last[c + q] = T.first[nn];
                T.sfirst[c + q] = T.slast[c + q] = T.sfirst[nn];
                bb_reset(T.bb + 4ll * (c + q));
            }
        });
        coop_scan_apply(PartFlag{T, A.px, A.py}, n, A.partP, sh, [&](long long p, u32 ex, u32 inc) {
            A.G[p] = ex;
            if (p == n - 1) A.G[n] = inc;
        });
        if (nseg) coop_scan_apply(SegFlag{T, A.sx, A.sy, segin}, nseg, A.partS, sh, [&](long long k, u32 ex, u32 inc) {
            A.Gs[k] = ex;
            if (k == nseg - 1) A.Gs[nseg] = inc;
        });
        // leaves of this level keep ch1 = -1 (written above only for tiles that exist: na >= 1)
        grid.sync();
        // ---- P4: partner table of the Hoare partition + child ranges; stable split of the segments
        for (long long p = gtid; p < n; p += gsize) {
            const int node = T.pnode[p];
            if (T.status[node] != ST_SPLIT) continue;
            const int f = T.first[node], l = T.last[node];
            const u32 Gf = A.G[f];
            const int m = (int)(A.G[l] - Gf);
            const int rel = (int)p - f;
            const int le = (int)(A.G[p] - Gf);
            const bool isless = A.G[p + 1] != A.G[p];
            if (p == f) {
                const int c = T.ch1[node];
                T.first[c] = f; T.last[c] = f + m;
                T.first[c + 1] = f + m; T.last[c + 1] = l;
            }
            if (rel >= m && isless) A.tmpR[f + (m - le - 1)] = (int)p;
        }
        if (nseg) {
            int* pout = A.segperm[segcur ^ 1];
            int* nout = A.snode[segcur ^ 1];
            for (long long k = gtid; k < nseg; k += gsize) {
                const int node = T.snode[k];
                if (T.status[node] != ST_SPLIT) { pout[k] = segin[k]; nout[k] = node; continue; }
                const int f = T.sfirst[node], l = T.slast[node];
                const u32 Gf = A.Gs[f];
                const int m = (int)(A.Gs[l] - Gf);
                const int le = (int)(A.Gs[k] - Gf);
                const bool isless = A.Gs[k + 1] != A.Gs[k];
                const int c = T.ch1[node];
                if (k == f) {
                    T.sfirst[c] = f; T.slast[c] = f + m;
                    T.sfirst[c + 1] = f + m; T.slast[c + 1] = l;
                }
                const int dst = isless ? (f + le) : (f + m + ((int)k - f - le));
                pout[dst] = segin[k];
                nout[dst] = isless ? c : c + 1;
            }
        }
        grid.sync();
        // ---- P5: in-place swaps (each pair is touched by exactly one thread) and the new node ids
        for (long long p = gtid; p < n; p += gsize) {
            const int node = T.pnode[p];
            if (T.status[node] != ST_SPLIT) continue;
            const int f = T.first[node], l = T.last[node];
            const u32 Gf = A.G[f];
            const int m = (int)(A.G[l] - Gf);
            const int rel = (int)p - f;
            const int le = (int)(A.G[p] - Gf);
            const bool isless = A.G[p + 1] != A.G[p];
            if (rel < m && !isless) {
                const int q = A.tmpR[f + (rel - le)];
                double t;
                t = A.px[p]; A.px[p] = A.px[q]; A.px[q] = t;
                t = A.py[p]; A.py[p] = A.py[q]; A.py[q] = t;
                t = A.pg[p]; A.pg[p] = A.pg[q]; A.pg[q] = t;
                int ti = A.perm[p]; A.perm[p] = A.perm[q]; A.perm[q] = ti;
            }
            T.pnode[p] = T.ch1[node] + (rel >= m ? 1 : 0);
        }
        if (nseg) { segcur ^= 1; T.snode = A.snode[segcur]; }
        grid.sync();
        // ---- P6: Stretch of the children
        coop_bbox(T, A.snode[segcur], A.px, A.py, n, A.sx, A.sy, A.segperm[segcur], nseg, gtid, gsize);
        a0 = a1; a1 += 2 * (int)nsplit;
        if (gtid == 0) A.st->lvl[d + 2] = a1;
        grid.sync();
    }
    if (failed) {
        if (gtid == 0) A.st->err = 1;
        return;
    }
    const int depth = d;
    grid.sync();   // lvl[] complete and visible
    // ---- bottom-up: subtree sizes, +/- centres of mass (CalculateCMass, :150-197)
    for (int dd = depth; dd >= 0; dd--) {
        const int b0 = A.st->lvl[dd], b1 = A.st->lvl[dd + 1];
        for (long long k = b0 + gtid; k < b1; k += gsize) {
            const int nn = (int)k;
            double* P = T.cmp + 3ll * nn;
            double* M = T.cmm + 3ll * nn;
            const int c = T.ch1[nn];
            if (c < 0) {
                T.nl[nn] = 1; T.nn[nn] = 1;
                double Px = 0, Py = 0, Pg = 0, Mx = 0, My = 0, Mg = 0;
                for (int i = T.first[nn]; i < T.last[nn]; i++) {
                    double g = A.pg[i];
                    if (g > 0) { Px = VV_ADD(Px, VV_MUL(A.px[i], g)); Py = VV_ADD(Py, VV_MUL(A.py[i], g)); Pg = VV_ADD(Pg, g); }
                    else { Mx = VV_ADD(Mx, VV_MUL(A.px[i], g)); My = VV_ADD(My, VV_MUL(A.py[i], g)); Mg = VV_ADD(Mg, g); }
                }
                if (Pg != 0) { double r = 1. / Pg; Px = VV_MUL(Px, r); Py = VV_MUL(Py, r); } else { Px = T.x[nn]; Py = T.y[nn]; }
                if (Mg != 0) { double r = 1. / Mg; Mx = VV_MUL(Mx, r); My = VV_MUL(My, r); } else { Mx = T.x[nn]; My = T.y[nn]; }
                P[0] = Px; P[1] = Py; P[2] = Pg; M[0] = Mx; M[1] = My; M[2] = Mg;
                continue;
            }
            T.nl[nn] = T.nl[c] + T.nl[c + 1];
            T.nn[nn] = 1 + T.nn[c] + T.nn[c + 1];
            for (int s = 0; s < 2; s++) {
                double* cm = s ? M : P;
                const double* Aa = (s ? T.cmm : T.cmp) + 3ll * c;
                const double* Bb = (s ? T.cmm : T.cmp) + 3ll * (c + 1);
                double sumg = VV_ADD(Aa[2], Bb[2]);
                if (sumg != 0) {
                    double r = 1. / sumg;
                    cm[0] = VV_MUL(VV_ADD(VV_MUL(Aa[0], Aa[2]), VV_MUL(Bb[0], Bb[2])), r);
                    cm[1] = VV_MUL(VV_ADD(VV_MUL(Aa[1], Aa[2]), VV_MUL(Bb[1], Bb[2])), r);
                    cm[2] = sumg;
                } else { cm[0] = T.x[nn]; cm[1] = T.y[nn]; cm[2] = 0; }
            }
        }
        grid.sync();
    }
    // ---- top-down: DFS leaf index and pre-order id
    for (int dd = 0; dd <= depth; dd++) {
        const int b0 = A.st->lvl[dd], b1 = A.st->lvl[dd + 1];
        for (long long k = b0 + gtid; k < b1; k += gsize) {
            const int nn = (int)k;
            if (nn == 0) { T.lstart[0] = 0; T.pre[0] = 0; }
            const int c = T.ch1[nn];
            if (c < 0) { T.leaf_node[T.lstart[nn]] = nn; continue; }
            T.lstart[c] = T.lstart[nn];
            T.lstart[c + 1] = T.lstart[nn] + T.nl[c];
            T.pre[c] = T.pre[nn] + 1;
            T.pre[c + 1] = T.pre[nn] + 1 + T.nn[c];
        }
        grid.sync();
    }
    if (gtid == 0) { A.st->nnodes = a1; A.st->depth = depth; A.st->nleaves = T.nl[0]; }
}

}  // namespace vv
