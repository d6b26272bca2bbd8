/* vvgpu — C ABI of the B200-native particle hot path for vvflow (libvvhd).
 *
 * The reference has no plugin/FFI interface; the boundary is the C++ class surface that
 * utils/vvflow/vvflow.cpp:201-208,246-257 instantiates and calls once per time step. Each
 * entry point below names the reference member it replaces (paths relative to the reference
 * tree). Adapter classes with the reference's own names (vvflow_b200/host/vvgpu_adapter.hpp)
 * forward to these calls; INTEGRATION.md shows the binding.
 *
 * Conventions: every call returns 0 on success or a negative VVGPU_E* code (vvgpu_strerror
 * gives the text, vvgpu_last_error the detail); no exception crosses the boundary. One
 * context owns one CUDA device and is driven from one host thread. All pointers are HOST
 * pointers unless the name ends in _dev. Plain pointers and sizes only.
 */
#ifndef VVGPU_H
#define VVGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vvgpu_ctx vvgpu_ctx;

enum {
    VVGPU_OK = 0,
    VVGPU_EINVAL = -1,   /* bad argument (the reference throws std::invalid_argument) */
    VVGPU_ESTATE = -2,   /* call out of order, e.g. epsilon before tree_build ("tree is not built") */
    VVGPU_ECUDA = -3,    /* CUDA runtime error; no CPU fallback exists */
    VVGPU_ENOMEM = -4,
    VVGPU_ELIMIT = -5    /* an internal capacity (tree depth, traversal stack) was exceeded */
};

/* TObj, libvvhd/headers/TObj.hpp:10-16 — 48 bytes: r.x r.y g v.x v.y _1_eps */
typedef struct { double x, y, g, vx, vy, ieps; } vvgpu_obj;

/* TAtt, libvvhd/headers/TBody.hpp:17-58 — the fields the hot path reads */
typedef struct {
    double rx, ry;       /* TObj::r  (segment centre) */
    double cx, cy;       /* corner */
    double dlx, dly;     /* dl */
    double g;            /* attached circulation */
    double ieps;         /* TObj::_1_eps = 3/|dl| (TBody.cpp:209) */
    int32_t slip;
    int32_t body;        /* index into the bodies array */
} vvgpu_seg;

/* the TBody state the hot path reads: get_axis(), _cofm, _min_rect_*, _min_disc_r2,
 * isInsideValid(), speed_slae (libvvhd/headers/TBody.hpp, src/TBody.cpp:238-246) */
typedef struct {
    double axis_x, axis_y, cofm_x, cofm_y;
    double bl_x, bl_y, tr_x, tr_y, disc_r2;
    double speed_x, speed_y, speed_o;   /* speed_slae */
    int32_t inside_valid;
    int32_t first_seg, n_seg;           /* this body's range in the segment array */
    int32_t _pad;
} vvgpu_body;

enum { VVGPU_LIST_VORTEX = 0 };  /* heat / streak lists: SURVEY.md §8(f) row 3, not built yet */

/* ---- lifetime ------------------------------------------------------------------------- */
int vvgpu_create(int device, vvgpu_ctx** out);
void vvgpu_destroy(vvgpu_ctx* ctx);
const char* vvgpu_strerror(int code);
const char* vvgpu_last_error(const vvgpu_ctx* ctx);

/* ---- state in / out (Space::VortexList, Space::BodyList; libvvhd/headers/TSpace.hpp:38-44) */
/* (set_particles / get_particles accept a host OR a device address: the copy uses cudaMemcpyDefault) */
int vvgpu_set_particles(vvgpu_ctx* ctx, int list, const vvgpu_obj* objs, size_t n);
/* 24-byte (x,y,g) records as Space::load_list_bin reads them (TSpace.cpp:161-175); v,_1_eps = 0 */
int vvgpu_set_particles_xyg(vvgpu_ctx* ctx, int list, const double* xyg, size_t n);
/* append n records behind the particles that are resident on the device (what
 * MFlowmove::vortex_shed adds per step, MFlowmove.cpp:217-235, without re-uploading the whole list). The caller-order
 * index (vvgpu_get_permutation) of the appended particles continues after the last one handed over so far. */
int vvgpu_append_particles(vvgpu_ctx* ctx, int list, const vvgpu_obj* objs, size_t n);
int vvgpu_particle_count(vvgpu_ctx* ctx, int list, size_t* n);
/* Space::gsum() (libvvhd/headers/TSpace.hpp): total circulation of the resident list — the right-hand side of the
 * SLAE's circulation equation (MConvectiveFast.cpp:511) when the list lives on the device */
int vvgpu_particle_gsum(vvgpu_ctx* ctx, int list, double* sum);
/* current device order (after tree_build: the reference's in-place permuted order) */
int vvgpu_get_particles(vvgpu_ctx* ctx, int list, vvgpu_obj* out, size_t cap, size_t* n);
/* records [first, first + count) of the current device order (a rank of a multi-GPU job brings back its share only) */
int vvgpu_get_particles_range(vvgpu_ctx* ctx, int list, vvgpu_obj* out, size_t first, size_t count);
/* orig[i] = index, in the order of the last set_particles call, of the particle now at i */
int vvgpu_get_permutation(vvgpu_ctx* ctx, int list, int32_t* orig, size_t cap);
int vvgpu_set_bodies(vvgpu_ctx* ctx, const vvgpu_seg* segs, size_t nseg, const vvgpu_body* bodies, size_t nbody);

/* ---- stree (libvvhd/headers/TSortedTree.hpp:60-92) -------------------------------------- */
/* stree::stree + stree::build(includeV, includeB, .) — TSortedTree.cpp:221-265.
 * include_mask bit0 = vortexes, bit1 = bodies. Building a built tree returns VVGPU_ESTATE
 * (the reference prints "Tree is already built" and returns). */
int vvgpu_tree_build(vvgpu_ctx* ctx, int far_criteria, double min_node_size, double max_node_size,
                     unsigned include_mask);
int vvgpu_tree_destroy(vvgpu_ctx* ctx);                 /* stree::destroy, :267-273 */
int vvgpu_tree_counts(vvgpu_ctx* ctx, size_t* n_nodes, size_t* n_leaves, size_t* depth);
/* Nodes in DFS pre-order (child 1 first). dbl: 10 per node (x y h w CMp.x CMp.y CMp.g CMm.x CMm.y CMm.g);
 * idx: 8 per node (vfirst vlast nseg ch1 ch2 leaf_index sfirst depth), -1 where n/a.
 * Serves stree::getBottomNodes()/findNode() to host code (:275-303). */
int vvgpu_tree_export(vvgpu_ctx* ctx, double* dbl, int64_t* idx, size_t cap_nodes);
/* per-leaf interaction lists exactly as snode::FindNearNodes (:199-217) builds them, CSR:
 * near -> leaf indices, far -> pre-order node ids, both in DFS order. Pass idx == NULL to get
 * only the ptr arrays (sizes n_leaves+1). */
int vvgpu_tree_lists(vvgpu_ctx* ctx, int64_t* near_ptr, int64_t* near_idx, size_t near_cap, int64_t* far_ptr,
                     int64_t* far_idx, size_t far_cap);
/* segment ids held by each leaf's bllist (CSR, n_leaves+1 / nseg) */
int vvgpu_tree_leaf_segments(vvgpu_ctx* ctx, int64_t* ptr, int64_t* idx, size_t cap);
/* near pairs (targets g!=0 x sources g!=0 over near leaves, self included) and far-node visits:
 * the "interactions" of the headline metric */
int vvgpu_count_interactions(vvgpu_ctx* ctx, double* near_pairs, double* far_nodes);

/* ---- MEpsilonFast::CalcEpsilonFast(merge), MEpsilonFast.cpp:11-63; Merged() -> *merged ---- */
int vvgpu_epsilon(vvgpu_ctx* ctx, int merge, int* merged);
/* rounds the merge fixed point of the last vvgpu_epsilon(merge=1) took (one host read-back each) */
int vvgpu_merge_rounds(vvgpu_ctx* ctx, int* rounds);
/* ---- MConvectiveFast::process_all_lists, MConvectiveFast.cpp:36-114. inf_* = S->inf_speed(),
 * dt = S->dt (sink epsilon, :160), sinks = Space::SourceList as (x,y,g) triples -------------- */
int vvgpu_convective(vvgpu_ctx* ctx, double inf_vx, double inf_vy, double dt, const double* sinks_xyg,
                     size_t nsink);
/* ---- MConvectiveFast::velocity(TVec p), MConvectiveFast.cpp:20-34 (sensors, vvplot rasters):
 * velocity at npts arbitrary points xy (x, y pairs, host) -> vxy_out (vx, vy pairs, host). Needs a built tree
 * (fails like stree::findNode, TSortedTree.cpp:286-288, otherwise) and the _1_eps of the current particle set
 * (vvgpu_epsilon of this step, or the values passed in with the particles). ---------------------------------- */
int vvgpu_velocity_at(vvgpu_ctx* ctx, const double* xy, size_t npts, double inf_vx, double inf_vy, double dt,
                      const double* sinks_xyg, size_t nsink, double* vxy_out);
/* ---- static MEpsilonFast::eps2h(node, p) and ::h2(node, p), MEpsilonFast.cpp:66-107, with node = findNode(p)
 * (what XVorticity.cpp:52,90 and XStreamfunction.cpp:89 evaluate): per point the squared distance to the
 * second-nearest particle and to the nearest body segment (+inf without bodies) of the leaf's near leaves.
 * eps2h_h2_out receives npts (eps2h, h2) pairs; bit-identical to the reference. ---------------------------- */
int vvgpu_eps2h_h2_at(vvgpu_ctx* ctx, const double* xy, size_t npts, double* eps2h_h2_out);
/* ---- MConvectiveFast::NodeInfluence(*tree->findNode(seg.r), seg), MConvectiveFast.cpp:398-418:
 * the free vortices' term of the slip equation's right-hand side (fillSlipEquationForSegment, :459-467), for every
 * segment passed to vvgpu_set_bodies, in that order. Uses the _1_eps the particles carry (the reference solves the
 * SLAE before CalcEpsilonFast, so these are last step's values that came in with the TObj records). With it
 * calc_circulation needs no CPU tree: see INTEGRATION.md §2b. ------------------------------------------------ */
int vvgpu_node_influence(vvgpu_ctx* ctx, double* out_nseg);
/* ---- XVorticity::evaluate, XVorticity.cpp:26-97 (the vorticity raster of vvplot) on the resident
 * vortex list, which must already hold what MFlowmove::vortex_shed adds (XVorticity.cpp:36; append them with
 * vvgpu_append_particles). Builds its own tree (far criteria 8, minNodeSize 20 dl, :38) on a copy, so the resident
 * list keeps its order, like the reference's Space copy. xmin, ymin, dxdy are floats as in XField; dl =
 * Space::average_segment_length(). out[yj * xres + xi] in double: XField::map stores the same value as float. ---- */
int vvgpu_vorticity_raster(vvgpu_ctx* ctx, float xmin, float ymin, float dxdy, int xres, int yres, double eps_mult,
                           double dl, double* out);
/* ---- XPressure::evaluate, XPressure.cpp:32-146 (the pressure raster of vvplot) on the resident vortex
 * list, which must already hold what MFlowmove::vortex_shed adds (:56; append with vvgpu_append_particles), with
 * gsum_nseg = TAtt::gsum of every segment AFTER that shed (vortex_shed adds g, MFlowmove.cpp:227). Builds its own tree
 * (far criteria 8, minNodeSize 20 dl, maxNodeSize 0.1, :27) and runs one velocity pass (epsilon without merging,
 * convective, diffusive, :61-64) on a copy: the resident list comes back as it was. ref frame 's': use_ref_speed = 0;
 * 'o' / 'f' / 'b': use_ref_speed = 1 with (0,0) / inf_speed / the body's speed (:38-52). out[yj * xres + xi] in
 * double: XField::map stores the same value as float. Single-rank contexts. ---------------------------------------- */
int vvgpu_pressure_raster(vvgpu_ctx* ctx, float xmin, float ymin, float dxdy, int xres, int yres, double dl, double re,
                          double dt, double inf_vx, double inf_vy, const double* sinks_xyg, size_t nsink,
                          const double* gsum_nseg, int use_ref_speed, double ref_vx, double ref_vy, double* out);
/* ---- MDiffusiveFast::process_vort_list, MDiffusiveFast.cpp:8-48. fric_out (nseg, may be NULL)
 * receives the per-segment increments of TAtt::fric (:121-122) ------------------------------- */
int vvgpu_diffusive(vvgpu_ctx* ctx, double re, double* fric_out);
/* ---- MFlowmove::move_and_clean particle part, MFlowmove.cpp:107-144,194-199. dt_eff is the
 * collision-shortened step (:25-57, computed by the caller). Outputs (may be NULL) are the
 * increments to TBody::fdt_dead (x,y,o per body), TBody::g_dead, TAtt::gsum (per segment). ---- */
int vvgpu_move_and_clean(vvgpu_ctx* ctx, double dt_eff, double remove_eps, int remove_in_body,
                         double* fdt_dead_xyo, double* g_dead, double* gsum_delta, size_t* cleaned);

/* ---- multi-GPU (target-sharded, source-replicated) ----------------------------
 * Every rank holds all particles and rebuilds the (deterministic) tree; the leaf groups are dealt block-cyclically
 * over the ranks and each rank computes epsilon / convective / diffusive for the particles of ITS groups. The
 * exchanges happen inside the calls, on the context's stream: _1_eps (and the merge columns, once per round of the
 * merge fixed point) at the end of vvgpu_epsilon, v before vvgpu_tree_destroy / vvgpu_get_particles, TAtt::fric
 * summed over the ranks in vvgpu_diffusive (MDiffusiveFast.cpp:121-122). All ranks make the same calls in the same
 * order with the same arguments; the results are bit-identical on every rank and to the single-GPU results, except
 * fric (a sum over ranks: 1e-10). Two transports:
 *   one process per GPU (torchrun):  rank 0 calls vvgpu_comm_unique_id, the host layer carries the 128 bytes to the
 *     other ranks, every rank calls vvgpu_comm_init (NCCL all-gathers over NVLink; libnccl.so.2 is dlopen'ed);
 *   one process, several GPUs:       vvgpu_group_create makes one context per entry of `devices` (entries may
 *     repeat: several ranks on one device, for tests); each context is then driven from its own host thread and
 *     the ranks exchange by peer-to-peer copies. */
int vvgpu_comm_unique_id(void* id128, size_t cap);
int vvgpu_comm_init(vvgpu_ctx* ctx, int rank, int nranks, const void* id128);
int vvgpu_group_create(const int* devices, int n, vvgpu_ctx** ctxs_out);
int vvgpu_comm_info(vvgpu_ctx* ctx, int* rank, int* nranks, int* kind /* 0 none, 1 NCCL, 2 in-process */);
/* collective: complete the exchanges that are done lazily (v of the other ranks' targets), e.g. before ONE rank reads the
 * particle list while the tree is still built; a no-op when nothing is pending */
int vvgpu_sync_ranks(vvgpu_ctx* ctx);
/* rank that owns leaf group `group` (groups of 32 consecutive leaves, dealt round-robin in pieces of vvgpu_shard_block()
 * consecutive groups); no device needed */
int vvgpu_shard_owner(int group, int nranks);
int vvgpu_shard_block(void);
/* e2e upload with one slice per rank: rank r passes records [n r / P, n (r + 1) / P) of a list of n_total (host or
 * device address); the slices are gathered over the transport */
int vvgpu_set_particles_slice(vvgpu_ctx* ctx, int list, const vvgpu_obj* objs, size_t first, size_t count, size_t n_total);
/* device pointers of the SoA arrays (x y g vx vy ieps) */
int vvgpu_particle_arrays_dev(vvgpu_ctx* ctx, int list, double** arrays6, size_t* n);
int vvgpu_stream(vvgpu_ctx* ctx, void** cuda_stream);
int vvgpu_synchronize(vvgpu_ctx* ctx);

/* ---- measurement ---------------------------------------------------------------------------- */
enum { VVGPU_T_BUILD = 0, VVGPU_T_LISTS, VVGPU_T_EPS, VVGPU_T_CONV, VVGPU_T_DIFF, VVGPU_T_MOVE, VVGPU_T_COUNT };
/* CUDA-event milliseconds of the last call of each phase; launches = kernels launched since the
 * last vvgpu_phase_times call */
int vvgpu_phase_times(vvgpu_ctx* ctx, double* ms, uint64_t* launches);
/* host waits on the context's stream (read-backs) since the last call of this function */
int vvgpu_host_syncs(vvgpu_ctx* ctx, uint64_t* n);
/* FP64 DFMA micro-benchmark: achieved TFLOP/s of a register-resident FMA chain on this device */
int vvgpu_fp64_peak(vvgpu_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
