// K1 — device rebuild of the reference's adaptive bisection tree, bit-exact.
//
// Replaces stree::build (libvvhd/src/TSortedTree.cpp:232-265): snode::Stretch (:101-137),
// snode::DivideNode (:36-79), both snode::DistributeContent overloads (:81-99,:139-148) and
// snode::CalculateCMass (:150-197). The reference recurses depth-first over heap nodes; here
// the same tree is grown level by level over flat arrays:
//   * node boxes are exact min/max reductions (order-free), so they are identical;
//   * the unstable Hoare partition of :81-99 has a closed form — with m = #(coord < mid), the
//     k-th not-less element of [first, first+m) (ascending) swaps with the k-th less element of
//     [first+m, last) (descending) — evaluated with one global prefix sum per level;
//   * DFS pre-order ids / leaf order (= the order of stree::bottomNodes) are recovered afterwards
//     from subtree sizes (one bottom-up and one top-down sweep over the levels).
// Nodes are stored in creation order: level d occupies [lvl[d], lvl[d+1]); siblings are adjacent
// (ch2 = ch1 + 1).
#pragma once
#include "vvgpu_common.cuh"

namespace vv {

constexpr int kTreeMaxList = 16;  // Tree_MaxListSize, TSortedTree.cpp:10
enum : unsigned char { ST_PENDING = 0, ST_LEAF = 1, ST_SPLIT = 2 };

struct TreeDev {
    // per node
    double *x, *y, *h, *w;
    u64* bb;  // 4 per node: min x, min y, max x, max y (ordered encoding)
    int *first, *last, *sfirst, *slast, *ch1, *parent, *depth;
    unsigned char *status, *axis;  // axis 1: split on x (h < w), 0: split on y
    int *nl, *nn, *lstart, *pre;   // subtree leaves / nodes, first leaf index, pre-order id
    double *cmp, *cmm;             // 3 per node
    // per leaf (DFS order)
    int* leaf_node;
    // per particle position / per segment position
    int* pnode;
    int* snode;
};

struct BuildParams {
    double min_node, max_node;
};

__device__ __forceinline__ void bb_reset(u64* bb) {
    bb[0] = enc_ordered(DBL_MAX);   // bl.x   (Stretch sentinels, TSortedTree.cpp:103-104)
    bb[1] = enc_ordered(DBL_MAX);   // bl.y
    bb[2] = enc_ordered(-DBL_MAX);  // tr.x
    bb[3] = enc_ordered(-DBL_MAX);  // tr.y
}

__global__ void k_tree_init_root(TreeDev T, int n, int nseg) {
    T.first[0] = 0; T.last[0] = n; T.sfirst[0] = 0; T.slast[0] = nseg;
    T.ch1[0] = -1; T.parent[0] = -1; T.depth[0] = 0; T.status[0] = ST_PENDING;
    bb_reset(T.bb);
}

__global__ void k_iota(int* a, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}

// Stretch: fold every object of a freshly created node into its box. Index space = particle
// positions [0,n) followed by segment positions [n, n+nseg).
__global__ void k_tree_bbox(TreeDev T, const double* __restrict__ px, const double* __restrict__ py, int n,
                            const double* __restrict__ sx, const double* __restrict__ sy,
                            const int* __restrict__ seg_perm, int nseg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int node = -1;
    double x = 0, y = 0;
    if (i < n) {
        node = T.pnode[i];
        x = px[i]; y = py[i];
    } else if (i < n + nseg) {
        int k = i - n;
        node = T.snode[k];
        int s = seg_perm[k];
        x = sx[s]; y = sy[s];
    }
    if (node >= 0 && T.status[node] != ST_PENDING) node = -1;
    // whole warp in one node (the common case near the root): reduce first, 4 atomics per warp
    int n0 = __shfl_sync(0xffffffffu, node, 0);
    bool uniform = __all_sync(0xffffffffu, node == n0);
    if (uniform) {
        if (n0 < 0) return;
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

// DivideNode's leaf tests (TSortedTree.cpp:38-56) for the nodes of one level [a0, a1).
__global__ void k_tree_decide(TreeDev T, int a0, int a1, BuildParams bp, u32* splitflag) {
    int n = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a1) return;
    const u64* bb = T.bb + 4ll * n;
    double blx = dec_ordered(bb[0]), bly = dec_ordered(bb[1]), trx = dec_ordered(bb[2]), try_ = dec_ordered(bb[3]);
    double x = VV_MUL(VV_ADD(blx, trx), 0.5);   // :122-125
    double y = VV_MUL(VV_ADD(bly, try_), 0.5);
    double h = VV_SUB(try_, bly);
    double w = VV_SUB(trx, blx);
    T.x[n] = x; T.y[n] = y; T.h[n] = h; T.w[n] = w;
    bool leaf = false;
    double mx = std_max(h, w), mn = std_min(h, w);
    if (mx < bp.max_node && mn <= bp.min_node) leaf = true;
    if (!leaf) {
        int m = T.slast[n] - T.sfirst[n];
        int nv = T.last[n] - T.first[n];
        if (nv > m) m = nv;
        if (mx < bp.max_node && m < kTreeMaxList) leaf = true;
    }
    T.status[n] = leaf ? ST_LEAF : ST_SPLIT;
    T.axis[n] = (h < w) ? 1 : 0;  // :87
    splitflag[n - a0] = leaf ? 0u : 1u;
}

// allocate the two children of every splitting node; children start with empty ranges that the
// partition kernels overwrite when the parent's range is non-empty
__global__ void k_tree_assign(TreeDev T, int a0, int a1, const u32* rank) {
    int n = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a1) return;
    if (T.status[n] != ST_SPLIT) { T.ch1[n] = -1; return; }
    int c = a1 + 2 * (int)rank[n - a0];
    T.ch1[n] = c;
    for (int k = 0; k < 2; k++) {
        T.parent[c + k] = n;
        T.depth[c + k] = T.depth[n] + 1;
        T.status[c + k] = ST_PENDING;
        T.ch1[c + k] = -1;
        T.first[c + k] = T.last[c + k] = T.first[n];
        T.sfirst[c + k] = T.slast[c + k] = T.sfirst[n];
        bb_reset(T.bb + 4ll * (c + k));
    }
}

struct PartFlag {  // 1 iff the object goes to child 1: coord < mid (TSortedTree.cpp:87,91)
    TreeDev T;
    const double *x, *y;
    __device__ __forceinline__ u32 operator()(long long p) const {
        int n = T.pnode[p];
        if (T.status[n] != ST_SPLIT) return 0;
        return T.axis[n] ? (x[p] < T.x[n]) : (y[p] < T.y[n]);
    }
};
struct SegFlag {
    TreeDev T;
    const double *sx, *sy;
    const int* seg_perm;
    __device__ __forceinline__ u32 operator()(long long k) const {
        int n = T.snode[k];
        if (T.status[n] != ST_SPLIT) return 0;
        int s = seg_perm[k];
        return T.axis[n] ? (sx[s] < T.x[n]) : (sy[s] < T.y[n]);
    }
};

// Hoare closed form, step 1: every less element of the right part publishes its position under
// its descending rank; child ranges are written by the thread at the node's first position.
__global__ void k_tree_partition(TreeDev T, int n, const u32* __restrict__ G, int* __restrict__ tmpR) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = T.pnode[p];
    if (T.status[node] != ST_SPLIT) return;
    int f = T.first[node], l = T.last[node];
    u32 Gf = G[f];
    int m = (int)(G[l] - Gf);
    int rel = p - f;
    int le = (int)(G[p] - Gf);
    bool isless = G[p + 1] != G[p];
    if (p == f) {
        int c = T.ch1[node];
        T.first[c] = f; T.last[c] = f + m;
        T.first[c + 1] = f + m; T.last[c + 1] = l;
    }
    if (rel >= m && isless) tmpR[f + (m - le - 1)] = p;
}
// step 2: every not-less element of the left part swaps with its partner, in place
__global__ void k_tree_swap(TreeDev T, int n, const u32* __restrict__ G, const int* __restrict__ tmpR, double* x,
                            double* y, double* g, int* perm) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = T.pnode[p];
    if (T.status[node] != ST_SPLIT) return;
    int f = T.first[node], l = T.last[node];
    u32 Gf = G[f];
    int m = (int)(G[l] - Gf);
    int rel = p - f;
    int le = (int)(G[p] - Gf);
    bool isless = G[p + 1] != G[p];
    if (rel < m && !isless) {
        int q = tmpR[f + (rel - le)];
        double t;
        t = x[p]; x[p] = x[q]; x[q] = t;
        t = y[p]; y[p] = y[q]; y[q] = t;
        t = g[p]; g[p] = g[q]; g[q] = t;
        int ti = perm[p]; perm[p] = perm[q]; perm[q] = ti;
    }
}
// step 3 (separate launch: k_tree_swap reads pnode of other positions' nodes only through
// status/first/last, but keeping the relabel apart keeps every kernel race-free by construction)
__global__ void k_tree_relabel(TreeDev T, int n, const u32* __restrict__ G) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int node = T.pnode[p];
    if (T.status[node] != ST_SPLIT) return;
    int f = T.first[node], l = T.last[node];
    int m = (int)(G[l] - G[f]);
    T.pnode[p] = T.ch1[node] + ((p - f) >= m ? 1 : 0);
}

// stable split of the segment lists (DistributeContent(LList&), TSortedTree.cpp:139-148)
__global__ void k_tree_seg_scatter(TreeDev T, int nseg, const u32* __restrict__ G, const int* __restrict__ perm_in,
                                   int* __restrict__ perm_out, int* __restrict__ snode_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    int node = T.snode[k];
    if (T.status[node] != ST_SPLIT) {
        perm_out[k] = perm_in[k];
        snode_out[k] = node;
        return;
    }
    int f = T.sfirst[node], l = T.slast[node];
    u32 Gf = G[f];
    int m = (int)(G[l] - Gf);
    int le = (int)(G[k] - Gf);
    bool isless = G[k + 1] != G[k];
    int c = T.ch1[node];
    if (k == f) {
        T.sfirst[c] = f; T.slast[c] = f + m;
        T.sfirst[c + 1] = f + m; T.slast[c + 1] = l;
    }
    int dst = isless ? (f + le) : (f + m + (k - f - le));
    perm_out[dst] = perm_in[k];
    snode_out[dst] = isless ? c : c + 1;
}

// Bottom-up sweep over one level: subtree sizes and the +/- centres of mass
// (CalculateCMass / CalculateCMassFromScratch, TSortedTree.cpp:150-197; sums run in array order,
// so they are bit-identical to the reference's).
__global__ void k_tree_up(TreeDev T, int a0, int a1, const double* __restrict__ px, const double* __restrict__ py,
                          const double* __restrict__ pg) {
    int n = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a1) return;
    double* P = T.cmp + 3ll * n;
    double* M = T.cmm + 3ll * n;
    int c = T.ch1[n];
    if (c < 0) {
        T.nl[n] = 1; T.nn[n] = 1;
        double Px = 0, Py = 0, Pg = 0, Mx = 0, My = 0, Mg = 0;
        for (int i = T.first[n]; i < T.last[n]; i++) {
            double g = pg[i];
            if (g > 0) { Px = VV_ADD(Px, VV_MUL(px[i], g)); Py = VV_ADD(Py, VV_MUL(py[i], g)); Pg = VV_ADD(Pg, g); }
            else { Mx = VV_ADD(Mx, VV_MUL(px[i], g)); My = VV_ADD(My, VV_MUL(py[i], g)); Mg = VV_ADD(Mg, g); }
        }
        if (Pg != 0) { double r = 1. / Pg; Px = VV_MUL(Px, r); Py = VV_MUL(Py, r); } else { Px = T.x[n]; Py = T.y[n]; }
        if (Mg != 0) { double r = 1. / Mg; Mx = VV_MUL(Mx, r); My = VV_MUL(My, r); } else { Mx = T.x[n]; My = T.y[n]; }
        P[0] = Px; P[1] = Py; P[2] = Pg; M[0] = Mx; M[1] = My; M[2] = Mg;
        return;
    }
    T.nl[n] = T.nl[c] + T.nl[c + 1];
    T.nn[n] = 1 + T.nn[c] + T.nn[c + 1];
    for (int s = 0; s < 2; s++) {
        double* cm = s ? M : P;
        const double* A = (s ? T.cmm : T.cmp) + 3ll * c;
        const double* B = (s ? T.cmm : T.cmp) + 3ll * (c + 1);
        double sumg = VV_ADD(A[2], B[2]);
        if (sumg != 0) {
            double r = 1. / sumg;
            cm[0] = VV_MUL(VV_ADD(VV_MUL(A[0], A[2]), VV_MUL(B[0], B[2])), r);
            cm[1] = VV_MUL(VV_ADD(VV_MUL(A[1], A[2]), VV_MUL(B[1], B[2])), r);
            cm[2] = sumg;
        } else { cm[0] = T.x[n]; cm[1] = T.y[n]; cm[2] = 0; }
    }
}

// Top-down sweep over one level: DFS leaf index and pre-order id
__global__ void k_tree_down(TreeDev T, int a0, int a1) {
    int n = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a1) return;
    if (n == 0) { T.lstart[0] = 0; T.pre[0] = 0; }
    int c = T.ch1[n];
    if (c < 0) { T.leaf_node[T.lstart[n]] = n; return; }
    T.lstart[c] = T.lstart[n];
    T.lstart[c + 1] = T.lstart[n] + T.nl[c];
    T.pre[c] = T.pre[n] + 1;
    T.pre[c + 1] = T.pre[n] + 1 + T.nn[c];
}

// carry the rest of the 48-byte TObj (v, _1_eps) and the caller's index through the permutation
__global__ void k_tree_gather_rest(int n, const int* __restrict__ perm, const double* __restrict__ vx,
                                   const double* __restrict__ vy, const double* __restrict__ ie,
                                   const int* __restrict__ orig, double* vx2, double* vy2, double* ie2, int* orig2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = perm[i];
    vx2[i] = vx[s]; vy2[i] = vy[s]; ie2[i] = ie[s]; orig2[i] = orig[s];
}

// compact per-leaf records used by every later phase
struct LeafDev {
    int *first, *last;      // particle range
    int *sfirst, *slast;    // range in seg_perm
    double *cx, *cy, *h, *w;
    int* node;
};
__global__ void k_leaf_fill(TreeDev T, LeafDev L, int nleaves) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    int n = T.leaf_node[l];
    L.node[l] = n;
    L.first[l] = T.first[n]; L.last[l] = T.last[n];
    L.sfirst[l] = T.sfirst[n]; L.slast[l] = T.slast[n];
    L.cx[l] = T.x[n]; L.cy[l] = T.y[n]; L.h[l] = T.h[n]; L.w[l] = T.w[n];
}

// export in DFS pre-order for host consumers (vvgpu_tree_export)
__global__ void k_tree_export(TreeDev T, int nnodes, double* dbl, long long* idx) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nnodes) return;
    int o = T.pre[n];
    double* d = dbl + 10ll * o;
    long long* k = idx + 8ll * o;
    d[0] = T.x[n]; d[1] = T.y[n]; d[2] = T.h[n]; d[3] = T.w[n];
    for (int j = 0; j < 3; j++) { d[4 + j] = T.cmp[3ll * n + j]; d[7 + j] = T.cmm[3ll * n + j]; }
    int c = T.ch1[n];
    k[0] = T.first[n]; k[1] = T.last[n];
    k[2] = (c < 0) ? (T.slast[n] - T.sfirst[n]) : 0;  // bllist.clear() on split, :71
    k[3] = (c < 0) ? -1 : T.pre[c];
    k[4] = (c < 0) ? -1 : T.pre[c + 1];
    k[5] = (c < 0) ? T.lstart[n] : -1;
    k[6] = T.sfirst[n];
    k[7] = T.depth[n];
}

}  // namespace vv
