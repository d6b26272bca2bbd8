// K1 — device rebuild of the reference's adaptive bisection tree, bit-exact: shared definitions.
//
// Replaces stree::build (libvvhd/src/TSortedTree.cpp:232-265): snode::Stretch (:101-137),
// snode::DivideNode (:36-79), both snode::DistributeContent overloads (:81-99,:139-148) and
// snode::CalculateCMass (:150-197). The reference recurses depth-first over heap nodes; here the
// same tree is grown breadth-first over flat arrays (vvgpu_tree_build.cuh):
//   * node boxes are exact min/max reductions (order-free), so they are identical;
//   * the unstable Hoare partition of :81-99 has a closed form — with m = #(coord < mid), the
//     k-th not-less element of [first, first+m) (ascending) swaps with the k-th less element of
//     [first+m, last) (descending) — evaluated with one prefix sum per level;
//   * DFS pre-order ids / leaf order (= the order of stree::bottomNodes) are recovered afterwards
//     from subtree sizes (one bottom-up and one top-down sweep).
// Node numbering: the nodes of the top phase in creation (level) order, then one block per
// CTA-built subtree in DFS order of the subtree roots, each block in the subtree's own level
// order. Siblings are always adjacent (ch2 = ch1 + 1).
#pragma once
#include "vvgpu_common.cuh"

namespace vv {

constexpr int kTreeMaxList = 16;  // Tree_MaxListSize, TSortedTree.cpp:10
// ST_SUB: a node that has to split and is small enough to be finished by one CTA in shared memory
enum : unsigned char { ST_PENDING = 0, ST_LEAF = 1, ST_SPLIT = 2, ST_SUB = 3 };

struct TreeDev {
    // per node
    double *x, *y, *h, *w;
    u64* bb;  // 4 per node: min x, min y, max x, max y (ordered encoding); top-phase nodes only
    int *first, *last, *sfirst, *slast, *ch1, *depth;
    unsigned char *status, *axis;  // axis 1: split on x (h < w), 0: split on y
    int *nl, *nn, *lstart, *pre;   // subtree leaves / nodes, first leaf index, pre-order id
    double *cmp, *cmm;             // 3 per node
    // per leaf (DFS order)
    int* leaf_node;
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

// Stretch's last four lines (:122-125) and DivideNode's leaf tests (:38-56) for one node
struct NodeGeom { double x, y, h, w; bool leaf; unsigned char axis; };
__device__ __forceinline__ NodeGeom node_decide(u64 b0, u64 b1, u64 b2, u64 b3, int nv, int ns, const BuildParams& bp) {
    NodeGeom g;
    const double blx = dec_ordered(b0), bly = dec_ordered(b1), trx = dec_ordered(b2), try_ = dec_ordered(b3);
    g.x = VV_MUL(VV_ADD(blx, trx), 0.5);
    g.y = VV_MUL(VV_ADD(bly, try_), 0.5);
    g.h = VV_SUB(try_, bly);
    g.w = VV_SUB(trx, blx);
    const double mx = std_max(g.h, g.w), mn = std_min(g.h, g.w);
    g.leaf = false;
    if (mx < bp.max_node && mn <= bp.min_node) g.leaf = true;
    if (!g.leaf) {
        const int m = nv > ns ? nv : ns;   // MaxListSize over the lists of the node
        if (mx < bp.max_node && m < kTreeMaxList) g.leaf = true;
    }
    g.axis = (g.h < g.w) ? 1 : 0;  // :87 (h == w splits on y)
    return g;
}

// +/- centres of mass of a leaf (CalculateCMassFromScratch, :176-197): sums run in array order, so they are
// bit-identical to the reference's
__device__ __forceinline__ void leaf_cmass(const double* px, const double* py, const double* pg, int first, int last,
                                           double nx, double ny, double* P, double* M) {
    double Px = 0, Py = 0, Pg = 0, Mx = 0, My = 0, Mg = 0;
    for (int i = first; i < last; i++) {
        const double g = pg[i];
        if (g > 0) { Px = VV_ADD(Px, VV_MUL(px[i], g)); Py = VV_ADD(Py, VV_MUL(py[i], g)); Pg = VV_ADD(Pg, g); }
        else { Mx = VV_ADD(Mx, VV_MUL(px[i], g)); My = VV_ADD(My, VV_MUL(py[i], g)); Mg = VV_ADD(Mg, g); }
    }
    if (Pg != 0) { const double r = 1. / Pg; Px = VV_MUL(Px, r); Py = VV_MUL(Py, r); } else { Px = nx; Py = ny; }
    if (Mg != 0) { const double r = 1. / Mg; Mx = VV_MUL(Mx, r); My = VV_MUL(My, r); } else { Mx = nx; My = ny; }
    P[0] = Px; P[1] = Py; P[2] = Pg; M[0] = Mx; M[1] = My; M[2] = Mg;
}
// CalculateCMassFromChilds (:163-170)
__device__ __forceinline__ void child_cmass(const double* A, const double* B, double nx, double ny, double* cm) {
    const double sumg = VV_ADD(A[2], B[2]);
    if (sumg != 0) {
        const double r = 1. / sumg;
        cm[0] = VV_MUL(VV_ADD(VV_MUL(A[0], A[2]), VV_MUL(B[0], B[2])), r);
        cm[1] = VV_MUL(VV_ADD(VV_MUL(A[1], A[2]), VV_MUL(B[1], B[2])), r);
        cm[2] = sumg;
    } else { cm[0] = nx; cm[1] = ny; cm[2] = 0; }
}

// carry the rest of the 48-byte TObj (g, v, _1_eps) and the caller's index through the permutation
__global__ void k_tree_gather_rest(int n, const int* __restrict__ perm, const double* __restrict__ g, const double* __restrict__ vx,
                                   const double* __restrict__ vy, const double* __restrict__ ie,
                                   const int* __restrict__ orig, double* g2, double* vx2, double* vy2, double* ie2, int* orig2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = perm[i];
    g2[i] = g[s]; vx2[i] = vx[s]; vy2[i] = vy[s]; ie2[i] = ie[s]; orig2[i] = orig[s];
}

// a failed build puts (x, y) back into the caller's order
__global__ void k_tree_unpermute(int n, const int* __restrict__ perm, const double* __restrict__ x, const double* __restrict__ y,
                                 double* x2, double* y2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = perm[i];
    x2[s] = x[i]; y2[s] = y[i];
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
