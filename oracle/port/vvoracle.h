/* TEST INFRASTRUCTURE — NOT part of the product path.
 *
 * vvoracle: a plain-C, single-threaded CPU restatement of libvvhd's per-step particle
 * hot path (reference v2.4.0, paths relative to /root/reference):
 *   stree::build                    libvvhd/src/TSortedTree.cpp:232-265
 *   MEpsilonFast::CalcEpsilonFast   libvvhd/src/MEpsilonFast.cpp:11-63
 *   MConvectiveFast::process_all_lists  libvvhd/src/MConvectiveFast.cpp:36-114
 *   MDiffusiveFast::process_vort_list   libvvhd/src/MDiffusiveFast.cpp:8-48
 *   MFlowmove::move_and_clean (particle part)  libvvhd/src/MFlowmove.cpp:107-144,194-199
 *   MConvectiveFast::velocity(p)    libvvhd/src/MConvectiveFast.cpp:20-34,139-151
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it. Parity is PINNED: tests/test_oracle_port.py checks every function here against
 * oracle/_ref/libvvref.so (the reference's own sources compiled unmodified) and against the
 * golden vectors under tests/golden/ generated from that build.
 *
 * Everything is IEEE double, compiled with -ffp-contract=off (the reference is built for
 * baseline x86-64 without FMA, libvvhd/CMakeLists.txt:32).
 */
#ifndef VVORACLE_H
#define VVORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* particle list, SoA; `orig` carries the caller's index through the tree permutation */
typedef struct {
    int64_t n;
    double *x, *y, *g, *vx, *vy, *ieps; /* ieps = TObj::_1_eps */
    int64_t* orig;
} vvo_plist;

/* body segments (TAtt, libvvhd/headers/TBody.hpp:17-58), flattened over all bodies in
 * Space::BodyList order */
typedef struct {
    int64_t nseg, nbody;
    double *rx, *ry, *cx, *cy, *dlx, *dly, *g, *ieps; /* r, corner, dl, g, _1_eps */
    int32_t* slip;
    int32_t* body;      /* body index of each segment */
    int64_t* bfirst;    /* nbody+1: segment range of each body */
    /* per body: axis(2) cofm(2) bl(2) tr(2) disc_r2 inside_valid speed_slae(3) = 13 doubles */
    double* bprop;
    /* outputs accumulated by the phases */
    double* fric;       /* nseg   (MDiffusiveFast.cpp:121-122) */
    double* gsum;       /* nseg   (MFlowmove.cpp:140) */
    double* fdt_dead;   /* 3*nbody (MFlowmove.cpp:138-139) */
    double* g_dead;     /* nbody  (MFlowmove.cpp:141) */
} vvo_bodies;

typedef struct {
    int64_t n_nodes, n_leaves, cap;
    /* nodes in DFS pre-order, child 1 first */
    double *x, *y, *h, *w;
    double *cmp, *cmm;                /* 3 per node: x y g */
    int64_t *vfirst, *vlast, *sfirst, *slast, *ch1, *ch2, *leaf, *depth;
    int64_t* leaf_node;               /* leaf index -> node id */
    int64_t* seg_perm;                /* nseg: segment ids, each leaf owns [sfirst,slast) */
    int64_t nseg;
    /* interaction lists (CSR over leaves): near -> leaf indices, far -> node ids */
    int64_t *near_ptr, *near_idx, *far_ptr, *far_idx;
    int64_t near_cap, far_cap;
} vvo_tree;

vvo_tree* vvo_tree_build(vvo_plist* p, const vvo_bodies* b, int far_criteria, double min_node,
                         double max_node);
void vvo_tree_free(vvo_tree* t);
int64_t vvo_find_node(const vvo_tree* t, double px, double py);

/* returns the number of merges */
int64_t vvo_epsilon(const vvo_tree* t, vvo_plist* p, const vvo_bodies* b, int merge);
void vvo_convective(const vvo_tree* t, vvo_plist* p, const vvo_bodies* b, double inf_vx, double inf_vy,
                    double dt, const double* sinks_xyg, int64_t nsink);
/* MConvectiveFast::velocity (MConvectiveFast.cpp:20-34) at npts points xy -> out (vx, vy pairs); needs ieps */
void vvo_velocity_at(const vvo_tree* t, const vvo_plist* p, const vvo_bodies* b, double inf_vx, double inf_vy, double dt,
                     const double* sinks_xyg, int64_t nsink, const double* xy, int64_t npts, double* out);
/* (MEpsilonFast::eps2h, MEpsilonFast::h2) of findNode(p), MEpsilonFast.cpp:66-107; out = npts pairs */
void vvo_eps2h_h2_at(const vvo_tree* t, const vvo_plist* p, const vvo_bodies* b, const double* xy, int64_t npts,
                     double* out);
/* MConvectiveFast::NodeInfluence(*findNode(seg.r), seg) of every segment, MConvectiveFast.cpp:398-418 -> out[nseg] */
void vvo_node_influence(const vvo_tree* t, const vvo_plist* p, const vvo_bodies* b, double* out);
/* XVorticity::evaluate (XVorticity.cpp:26-97) on the post-shed list p (permuted by the tree built inside) */
void vvo_vorticity_raster(vvo_plist* p, const vvo_bodies* b, float xmin, float ymin, float dxdy, int xres, int yres,
                          double eps_mult, double dl, double* out);
/* XPressure::evaluate (XPressure.cpp:32-146) on the post-shed list p and gsum (permuted by the tree built inside) */
void vvo_pressure_raster(vvo_plist* p, vvo_bodies* b, float xmin, float ymin, float dxdy, int xres, int yres, double dl,
                         double re, double dt, double inf_vx, double inf_vy, const double* sinks, int64_t nsink,
                         int use_ref_speed, double ref_vx, double ref_vy, double* out);
void vvo_diffusive(const vvo_tree* t, vvo_plist* p, vvo_bodies* b, double re);
/* advect, drop |g|<remove_eps, drop in-body (accumulating dead sums), zero v. Compacts p in place,
 * returns the new n; *cleaned = number removed by the in-body test */
int64_t vvo_move_and_clean(vvo_plist* p, vvo_bodies* b, double dt, double remove_eps, int remove,
                           int64_t* cleaned);
/* TBody::isPointInvalid for body `ib`: nearest segment id (global) or -1 */
int64_t vvo_point_invalid(const vvo_bodies* b, int64_t ib, double px, double py);

/* bench helper: restrict epsilon/convective/diffusive to every stride-th leaf (default: all) */
void vvo_set_leaf_sample(int64_t stride, int64_t phase);
/* bench helper: near pairs / far nodes of leaves [l0, l1) */
void vvo_count_interactions(const vvo_tree* t, const vvo_plist* p, int64_t l0, int64_t l1, double* near_pairs,
                            double* far_nodes);

#ifdef __cplusplus
}
#endif
#endif
