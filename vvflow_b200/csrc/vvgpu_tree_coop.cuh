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
#include <cstdio>

namespace vv {
namespace cg = cooperative_groups;

#ifndef VV_COOP_THREADS
#define VV_COOP_THREADS 1024
#endif
#ifndef VV_COOP_BATCH
#define VV_COOP_BATCH 4
#endif
constexpr int kCoopThreads = VV_COOP_THREADS;
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

// The per-particle phases below walk the arrays with a grid stride. Every item needs a chain of three or
// four DEPENDENT gathers (pnode -> node record -> scan values), ~700 cycles each from L2: done one item at
// a time that chain was the whole cost of a phase (ncu: 11.7 % issue-active, 2.7 ms for 22 levels at N = 1M).
// Items are therefore processed in batches of kCoopBatch with the loads of a stage issued back to back.
constexpr int kCoopBatch = VV_COOP_BATCH;

// Stretch of freshly created nodes: fold every object into the box of its (pending) node.
// Objects of one node are contiguous, so a warp first reduces every RUN of equal node ids among its 32
// consecutive objects (segmented min/max scan) and only the last lane of a run touches the node's box;
// warps that lie entirely inside one node hand their result to warp 0 through shared memory, which reduces
// runs of equal nodes once more. Without this the top levels funnel N/32 atomics into the same four words
// and the deep levels issue four atomics per particle (the phase was 29 % of the build: 0.78 of 2.7 ms).
struct BoxSlot { int node; u64 mnx, mny, mxx, mxy; };

__device__ __forceinline__ void seg_minmax(int node, u64& mnx, u64& mny, u64& mxx, u64& mxy, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int tn = __shfl_up_sync(0xffffffffu, node, o);
        const u64 a = __shfl_up_sync(0xffffffffu, mnx, o), b = __shfl_up_sync(0xffffffffu, mny, o);
        const u64 c = __shfl_up_sync(0xffffffffu, mxx, o), d = __shfl_up_sync(0xffffffffu, mxy, o);
        if (lane >= o && tn == node) {   // runs are contiguous: equal ids o lanes apart lie in one run
            mnx = a < mnx ? a : mnx; mny = b < mny ? b : mny;
            mxx = c > mxx ? c : mxx; mxy = d > mxy ? d : mxy;
        }
    }
}
__device__ __forceinline__ void box_commit(const TreeDev& T, int node, u64 mnx, u64 mny, u64 mxx, u64 mxy) {
    u64* bb = T.bb + 4ll * node;
    atomicMin(bb + 0, mnx); atomicMin(bb + 1, mny);
    atomicMax(bb + 2, mxx); atomicMax(bb + 3, mxy);
}

__device__ __forceinline__ void coop_bbox(const TreeDev& T, const int* snode, const double* px, const double* py, int n,
                                          const double* sx, const double* sy, const int* seg_perm, int nseg,
                                          long long gtid, long long gsize, BoxSlot (*slots)[kCoopThreads / 32]) {
    const long long total = (long long)n + nseg;
    const long long rounds = (total + gsize - 1) / gsize;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long r0 = 0; r0 < rounds; r0 += kCoopBatch) {   // uniform trip count: barriers inside
        int node[kCoopBatch];
        double x[kCoopBatch], y[kCoopBatch];
#pragma unroll
        for (int k = 0; k < kCoopBatch; k++) {
            const long long i = (r0 + k) * gsize + gtid;
            node[k] = -1; x[k] = 0; y[k] = 0;
            if (r0 + k < rounds) {
                if (i < n) { node[k] = T.pnode[i]; x[k] = px[i]; y[k] = py[i]; }
                else if (i < total) { const int q = (int)(i - n); node[k] = snode[q]; const int s = seg_perm[q]; x[k] = sx[s]; y[k] = sy[s]; }
            }
        }
#pragma unroll
        for (int k = 0; k < kCoopBatch; k++)
            if (node[k] >= 0 && T.status[node[k]] != ST_PENDING) node[k] = -1;
#pragma unroll
        for (int k = 0; k < kCoopBatch; k++) {
            u64 mnx = enc_ordered(x[k]), mny = enc_ordered(y[k]), mxx = mnx, mxy = mny;
            // segment positions [n, total) follow the particles: a warp that straddles n holds runs of both kinds
            seg_minmax(node[k], mnx, mny, mxx, mxy, lane);
            const int nxt = __shfl_down_sync(0xffffffffu, node[k], 1);
            const bool runend = (lane == 31) || (nxt != node[k]);
            const int n0 = __shfl_sync(0xffffffffu, node[k], 0);
            const bool uniform = __all_sync(0xffffffffu, node[k] == n0);
            if (uniform) {
                if (lane == 31) { BoxSlot& sl = slots[k][warp]; sl.node = n0; sl.mnx = mnx; sl.mny = mny; sl.mxx = mxx; sl.mxy = mxy; }
            } else {
                if (lane == 31) slots[k][warp].node = -1;
                if (runend && node[k] >= 0) box_commit(T, node[k], mnx, mny, mxx, mxy);
            }
        }
        __syncthreads();
        if (warp < kCoopBatch) {   // warp k folds the uniform warps' results of batch item k
            BoxSlot sl; sl.node = -1; sl.mnx = sl.mny = sl.mxx = sl.mxy = 0;
            if (lane < kCoopThreads / 32) sl = slots[warp][lane];   // one slot per warp of the CTA
            int nd = sl.node;
            u64 mnx = sl.mnx, mny = sl.mny, mxx = sl.mxx, mxy = sl.mxy;
            seg_minmax(nd, mnx, mny, mxx, mxy, lane);
            const int nxt = __shfl_down_sync(0xffffffffu, nd, 1);
            if (((lane == 31) || (nxt != nd)) && nd >= 0) box_commit(T, nd, mnx, mny, mxx, mxy);
        }
        __syncthreads();
    }
}

#ifdef VV_TREE_TIMING
#define VV_TT(k) do { if (gtid == 0) { long long now_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)); tt_[k] += now_ - tlast_; tlast_ = now_; } } while (0)
#else
#define VV_TT(k) do { } while (0)
#endif
__global__ void __launch_bounds__(kCoopThreads, 1) k_tree_build_coop(CoopArgs A) {
#ifdef VV_TREE_TIMING
    long long tt_[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tlast_));
#endif
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 sh[kCoopThreads / 32 + 1];
    __shared__ BoxSlot slots[kCoopBatch][kCoopThreads / 32];
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
    grid.sync(); VV_TT(8);
    coop_bbox(T, A.snode[0], A.px, A.py, n, A.sx, A.sy, A.segperm[0], nseg, gtid, gsize, slots);
    grid.sync(); VV_TT(8);

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
        grid.sync(); VV_TT(0);
        // ---- P2: tile sums of the three flag arrays
        const int* segin = A.segperm[segcur];
        coop_tile_sums(FlagArray{A.splitflag}, na, A.partN, sh);
        coop_tile_sums(PartFlag{T, A.px, A.py}, n, A.partP, sh);
        if (nseg) coop_tile_sums(SegFlag{T, A.sx, A.sy, segin}, nseg, A.partS, sh);
        grid.sync(); VV_TT(1);
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
        grid.sync(); VV_TT(2);
        // ---- P4: partner table of the Hoare partition + child ranges; stable split of the segments
        for (long long p0 = gtid; p0 < n; p0 += (long long)kCoopBatch * gsize) {
            int node[kCoopBatch], f[kCoopBatch], l[kCoopBatch];
            u32 Gp[kCoopBatch], Gp1[kCoopBatch], Gf[kCoopBatch], Gl[kCoopBatch];
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                const long long p = p0 + (long long)k * gsize;
                node[k] = (p < n) ? T.pnode[p] : -1;
                Gp[k] = 0; Gp1[k] = 0;
                if (p < n) { Gp[k] = A.G[p]; Gp1[k] = A.G[p + 1]; }
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++)
                if (node[k] >= 0 && T.status[node[k]] != ST_SPLIT) node[k] = -1;
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                f[k] = 0; l[k] = 0;
                if (node[k] >= 0) { f[k] = T.first[node[k]]; l[k] = T.last[node[k]]; }
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                Gf[k] = 0; Gl[k] = 0;
                if (node[k] >= 0) { Gf[k] = A.G[f[k]]; Gl[k] = A.G[l[k]]; }
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                if (node[k] < 0) continue;
                const long long p = p0 + (long long)k * gsize;
                const int m = (int)(Gl[k] - Gf[k]);
                const int rel = (int)p - f[k];
                const int le = (int)(Gp[k] - Gf[k]);
                const bool isless = Gp1[k] != Gp[k];
                if (p == f[k]) {
                    const int c = T.ch1[node[k]];
                    T.first[c] = f[k]; T.last[c] = f[k] + m;
                    T.first[c + 1] = f[k] + m; T.last[c + 1] = l[k];
                }
                if (rel >= m && isless) A.tmpR[f[k] + (m - le - 1)] = (int)p;
            }
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
        grid.sync(); VV_TT(3);
        // ---- P5: in-place swaps (each pair is touched by exactly one thread) and the new node ids
        for (long long p0 = gtid; p0 < n; p0 += (long long)kCoopBatch * gsize) {
            int node[kCoopBatch], f[kCoopBatch], l[kCoopBatch], c1[kCoopBatch], q[kCoopBatch];
            u32 Gp[kCoopBatch], Gp1[kCoopBatch], Gf[kCoopBatch], Gl[kCoopBatch];
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                const long long p = p0 + (long long)k * gsize;
                node[k] = (p < n) ? T.pnode[p] : -1;
                Gp[k] = 0; Gp1[k] = 0;
                if (p < n) { Gp[k] = A.G[p]; Gp1[k] = A.G[p + 1]; }
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++)
                if (node[k] >= 0 && T.status[node[k]] != ST_SPLIT) node[k] = -1;
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                f[k] = 0; l[k] = 0; c1[k] = 0;
                if (node[k] >= 0) { f[k] = T.first[node[k]]; l[k] = T.last[node[k]]; c1[k] = T.ch1[node[k]]; }
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                Gf[k] = 0; Gl[k] = 0;
                if (node[k] >= 0) { Gf[k] = A.G[f[k]]; Gl[k] = A.G[l[k]]; }
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {   // partner of every not-less element of the left part
                q[k] = -1;
                if (node[k] < 0) continue;
                const long long p = p0 + (long long)k * gsize;
                const int m = (int)(Gl[k] - Gf[k]);
                const int rel = (int)p - f[k];
                const int le = (int)(Gp[k] - Gf[k]);
                const bool isless = Gp1[k] != Gp[k];
                if (rel < m && !isless) q[k] = A.tmpR[f[k] + (rel - le)];
                T.pnode[p] = c1[k] + (rel >= m ? 1 : 0);
            }
            double ax[kCoopBatch], ay[kCoopBatch], ag[kCoopBatch], bx[kCoopBatch], by[kCoopBatch], bg[kCoopBatch];
            int ai[kCoopBatch], bi[kCoopBatch];
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                if (q[k] < 0) continue;
                const long long p = p0 + (long long)k * gsize;
                ax[k] = A.px[p]; ay[k] = A.py[p]; ag[k] = A.pg[p]; ai[k] = A.perm[p];
                bx[k] = A.px[q[k]]; by[k] = A.py[q[k]]; bg[k] = A.pg[q[k]]; bi[k] = A.perm[q[k]];
            }
#pragma unroll
            for (int k = 0; k < kCoopBatch; k++) {
                if (q[k] < 0) continue;
                const long long p = p0 + (long long)k * gsize;
                A.px[p] = bx[k]; A.py[p] = by[k]; A.pg[p] = bg[k]; A.perm[p] = bi[k];
                A.px[q[k]] = ax[k]; A.py[q[k]] = ay[k]; A.pg[q[k]] = ag[k]; A.perm[q[k]] = ai[k];
            }
        }
        if (nseg) { segcur ^= 1; T.snode = A.snode[segcur]; }
        grid.sync(); VV_TT(4);
        // ---- P6: Stretch of the children
        coop_bbox(T, A.snode[segcur], A.px, A.py, n, A.sx, A.sy, A.segperm[segcur], nseg, gtid, gsize, slots);
        a0 = a1; a1 += 2 * (int)nsplit;
        if (gtid == 0) A.st->lvl[d + 2] = a1;
        grid.sync(); VV_TT(5);
    }
    if (failed) {
        if (gtid == 0) A.st->err = 1;
        return;
    }
    const int depth = d;
    grid.sync(); VV_TT(8);   // lvl[] complete and visible
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
        grid.sync(); VV_TT(6);
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
        grid.sync(); VV_TT(7);
    }
    if (gtid == 0) { A.st->nnodes = a1; A.st->depth = depth; A.st->nleaves = T.nl[0]; }
#ifdef VV_TREE_TIMING
    if (gtid == 0) printf("tree timing us: P1 %.0f P2 %.0f P3 %.0f P4 %.0f P5 %.0f P6 %.0f up %.0f down %.0f other %.0f (depth %d)\n",
                          tt_[0] * 1e-3, tt_[1] * 1e-3, tt_[2] * 1e-3, tt_[3] * 1e-3, tt_[4] * 1e-3, tt_[5] * 1e-3, tt_[6] * 1e-3, tt_[7] * 1e-3, tt_[8] * 1e-3, depth);
#endif
}

}  // namespace vv
