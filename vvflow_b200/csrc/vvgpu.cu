// libvvgpu.so — context, host-side orchestration and the C ABI declared in include/vvgpu.h.
// Device code lives in the vvgpu_*.cuh headers next to this file. There is no CPU fallback:
// every entry point either runs the CUDA kernels or returns an error.
#include "../../include/vvgpu.h"
#include "vvgpu_conv.cuh"
#include "vvgpu_diff.cuh"
#include "vvgpu_point.cuh"
#include "vvgpu_move.cuh"
#include "vvgpu_tree_build.cuh"
#include "vvgpu_shard.cuh"
#include "vvgpu_comm.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace vv;

namespace {

struct Buf {
    void* p = nullptr;
    size_t bytes = 0;
    template <class T>
    T* get(size_t n, bool* ok) {
        size_t need = n * sizeof(T);
        if (need == 0) need = sizeof(T);
        if (need > bytes) {
            if (p) cudaFree(p);
            p = nullptr;
            // headroom: a reallocation inside a step is a cudaFree + cudaMalloc (an implicit device synchronisation and
            // milliseconds under multi-process load); bookkeeping buffers whose size drifts from step to step (unit
            // tables, heavy-group scratch) get twice what they need, the big per-particle arrays a quarter more
            size_t want = (need < (64u << 20)) ? 2 * need + 4096 : need + need / 4 + 256;
            if (cudaMalloc(&p, want) != cudaSuccess) { bytes = 0; *ok = false; return nullptr; }
            bytes = want;
        }
        return (T*)p;
    }
    // grow, keeping the current contents (device-to-device copy on the given stream)
    template <class T>
    T* get_keep(size_t n, bool* ok, cudaStream_t st) {
        size_t need = n * sizeof(T);
        if (need <= bytes) return (T*)p;
        void* old = p; size_t oldb = bytes;
        p = nullptr; bytes = 0;
        T* q = get<T>(n, ok);
        if (q && old) cudaMemcpyAsync(q, old, oldb, cudaMemcpyDeviceToDevice, st);
        if (old) { cudaStreamSynchronize(st); cudaFree(old); }
        return q;
    }
    template <class T>
    T* as() const { return (T*)p; }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct PSet {  // one SoA particle set
    Buf x, y, g, vx, vy, ie, orig;
    Particles view() const { return Particles{x.as<double>(), y.as<double>(), g.as<double>(), vx.as<double>(), vy.as<double>(), ie.as<double>()}; }
    bool ensure(size_t n) {
        bool ok = true;
        x.get<double>(n, &ok); y.get<double>(n, &ok); g.get<double>(n, &ok);
        vx.get<double>(n, &ok); vy.get<double>(n, &ok); ie.get<double>(n, &ok); orig.get<int>(n, &ok);
        return ok;
    }
    bool ensure_keep(size_t n, cudaStream_t st) {   // grow, keeping the resident particles
        bool ok = true;
        x.get_keep<double>(n, &ok, st); y.get_keep<double>(n, &ok, st); g.get_keep<double>(n, &ok, st);
        vx.get_keep<double>(n, &ok, st); vy.get_keep<double>(n, &ok, st); ie.get_keep<double>(n, &ok, st);
        orig.get_keep<int>(n, &ok, st);
        return ok;
    }
    void release() { x.release(); y.release(); g.release(); vx.release(); vy.release(); ie.release(); orig.release(); }
};

}  // namespace

struct vvgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0, host_syncs = 0;
    cudaEvent_t ev0[VVGPU_T_COUNT], ev1[VVGPU_T_COUNT];
    bool ev_valid[VVGPU_T_COUNT] = {};
    int* h_pinned = nullptr;  // small pinned scratch for read-backs

    // particles (vortex list)
    size_t n = 0;
    size_t orig_next = 0;   // caller-order index of the next appended particle
    PSet ps[2];
    int cur = 0;
    Buf stage;  // AoS staging

    // bodies
    int nseg = 0, nbody = 0;
    Buf s_rx, s_ry, s_cx, s_cy, s_dlx, s_dly, s_g, s_ie, s_slip, s_body, b_first, b_prop;
    bool any_body_flow = false;
    Buf d_fric, d_gsum, d_fdt, d_gdead, d_cleaned;

    // tree
    bool built = false;
    int tn = 0, tnseg = 0;  // objects included in the built tree
    int nnodes = 0, nleaves = 0, depth = 0, ngroups = 0;
    double farc = 8;
    Buf t_x, t_y, t_h, t_w, t_bb, t_first, t_last, t_sfirst, t_slast, t_ch1, t_depth, t_status, t_axis,
        t_nl, t_nn, t_lstart, t_pre, t_cmp, t_cmm, t_leafnode, t_segperm[2], t_perm, t_tmpR;
    const int segcur = 0;   // the final segment order is always in t_segperm[0] (t_segperm[1] is the build's scratch)
    Buf scan_part, scan_out, flags, build_state;
    // tree build (vvgpu_tree_build.cuh)
    Buf b_enc, b_tilepre, b_chunktot, b_sublist, b_scratch, b_arena, b_aux, b_subinfo;
    int coop_grid = 0, sub_grid = 0;
    Buf l_first, l_last, l_sfirst, l_slast, l_cx, l_cy, l_h, l_w, l_node;
    Buf g_leaf, g_mask, g_cursor, slot_base, slot_count, taylor, farcount, d_err;
    long long pool_cap = 0;
    std::vector<int> h_hist;   // nodes per tree depth
    Buf u_group, u_base, u_count, u_first, u_num, u_sbase, u_tmp, near_scratch, src4, src2, lbox, wall_d, wall_key, hv_list,
        hv_inode, hv_imask, hv_icount, hv_tpart, hv_off;
    int nunits = 0;
    size_t nslots = 0;
    bool lists_ready = false;
    // epsilon
    Buf lcrit, lrestr, latt, ie_tmp, dyn, d_changed, d_nmerged, leaf_dirty, leaf_dbox, unit_dirty, group_dirty, tl;
    Buf mA[6], mB[6];
    Buf d_sinks, d_pairs, pt_xy, pt_out, pt_v;
    PSet ps_backup;   // the resident list while a raster evaluator works on its own tree

    // multi-GPU: target sharding + the transport (vvgpu_shard.cuh, vvgpu_comm.h)
    Comm comm;
    int sm_count = 148;
    int tsplit = 1;            // CTAs per work unit of the near-field kernels (NearArgs::tsplit), chosen with the unit table
    int ncta = 0, cta_heavy = 0;   // CTA table of the near-field kernels (k_cta_table): entries, entries of its first class
    bool cta_off = false;      // launch by the tsplit formula instead (incremental merge rounds)
    Buf cta_unit, cta_part, cta_out;
    int ngmine = 0;            // leaf groups this rank owns
    int npieces = 0;           // pieces of kShardBlock groups
    long long xL = 1;          // particles of the rank that owns most
    bool v_dirty = false;      // v of the other ranks' targets not yet gathered
    int merge_rounds = 0;      // rounds of the last merge fixed point
    Buf sh_first, sh_cnt, sh_off, sh_rankcnt, xsend, xrecv;
    Shard shard() const { return Shard{comm.rank, comm.nranks}; }
    ShardTable shard_table() { return ShardTable{sh_first.as<int>(), sh_cnt.as<int>(), sh_off.as<int>()}; }

    TreeDev T() {
        TreeDev t;
        t.x = t_x.as<double>(); t.y = t_y.as<double>(); t.h = t_h.as<double>(); t.w = t_w.as<double>();
        t.bb = t_bb.as<u64>();
        t.first = t_first.as<int>(); t.last = t_last.as<int>(); t.sfirst = t_sfirst.as<int>(); t.slast = t_slast.as<int>();
        t.ch1 = t_ch1.as<int>(); t.depth = t_depth.as<int>();
        t.status = t_status.as<unsigned char>(); t.axis = t_axis.as<unsigned char>();
        t.nl = t_nl.as<int>(); t.nn = t_nn.as<int>(); t.lstart = t_lstart.as<int>(); t.pre = t_pre.as<int>();
        t.cmp = t_cmp.as<double>(); t.cmm = t_cmm.as<double>();
        t.leaf_node = t_leafnode.as<int>();
        return t;
    }
    LeafDev Lv() {
        return LeafDev{l_first.as<int>(), l_last.as<int>(), l_sfirst.as<int>(), l_slast.as<int>(), l_cx.as<double>(),
                       l_cy.as<double>(), l_h.as<double>(), l_w.as<double>(), l_node.as<int>()};
    }
    GroupLists Gv() { return GroupLists{g_leaf.as<int>(), g_mask.as<u32>()}; }
    Units Uv() { return Units{u_group.as<int>(), u_base.as<long long>(), u_count.as<int>(), u_first.as<int>(), u_num.as<int>(), u_sbase.as<u32>()}; }
    NearArgs near_args() {
        NearArgs a;
        a.P = ps[cur].view(); a.L = Lv(); a.G = Gv(); a.nleaves = nleaves;
        a.U = Uv();
        a.scratch = near_scratch.p;
        a.src4 = src4.as<double4>();
        a.lbox = lbox.as<double>();
        a.nseg = tnseg;
        a.u0 = 0;
        a.tsplit = tsplit;
        const bool tab = ncta > 0 && !cta_off;
        a.cta_unit = tab ? cta_unit.as<int>() : nullptr;
        a.cta_part = tab ? cta_part.as<unsigned short>() : nullptr;
        a.cta_heavy = cta_heavy;
        a.seg_perm = t_segperm[segcur].as<int>();
        a.srx = s_rx.as<double>(); a.sry = s_ry.as<double>(); a.sdlx = s_dlx.as<double>(); a.sdly = s_dly.as<double>();
        return a;
    }
};

namespace {

int fail(vvgpu_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    // a failure of ONE rank of an in-process group (device error, out of memory, capacity) must not leave the others
    // waiting for it in the next exchange: release them, every later exchange of the group fails
    if (c && c->comm.kind == Comm::LOCAL && c->comm.grp && (code == VVGPU_ECUDA || code == VVGPU_ENOMEM || code == VVGPU_ELIMIT))
        c->comm.grp->abort();
    return code;
}
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(c, VVGPU_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)
#define CKLAUNCH()                                                                                   \
    do {                                                                                             \
        c->launches++;                                                                               \
        cudaError_t e_ = cudaGetLastError();                                                         \
        if (e_ != cudaSuccess) return fail(c, VVGPU_ECUDA, std::string("launch: ") + cudaGetErrorString(e_) + " at " + std::to_string(__LINE__)); \
    } while (0)
#define NEED(ptr)                                                                                    \
    do {                                                                                             \
        if (!ok) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");                                  \
    } while (0)

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }
// every host wait on the context's stream is counted (vvgpu_host_syncs): a read-back costs a round trip
inline cudaError_t stream_sync(vvgpu_ctx* c) { c->host_syncs++; return cudaStreamSynchronize(c->stream); }

struct PhaseTimer {
    vvgpu_ctx* c; int ph;
    PhaseTimer(vvgpu_ctx* c, int ph): c(c), ph(ph) { cudaEventRecord(c->ev0[ph], c->stream); }
    ~PhaseTimer() { cudaEventRecord(c->ev1[ph], c->stream); c->ev_valid[ph] = true; }
};

// exclusive scan of a flag functor over [0,n) into c->scan_out (n+1 entries)
template <class F>
int scan_flags(vvgpu_ctx* c, F f, long long n, u32* out) {
    bool ok = true;
    int tiles = cdiv(std::max<long long>(n, 1), kScanTile);
    u32* part = c->scan_part.get<u32>(tiles + 1, &ok);
    NEED(part);
    if (n == 0) { CK(cudaMemsetAsync(out, 0, sizeof(u32), c->stream)); return 0; }
    k_scan_reduce<<<tiles, kScanThreads, 0, c->stream>>>(f, n, part); CKLAUNCH();
    k_scan_partials<<<1, 1024, 0, c->stream>>>(part, tiles); CKLAUNCH();
    k_scan_apply<<<tiles, kScanThreads, 0, c->stream>>>(f, n, part, out); CKLAUNCH();
    return 0;
}

int read_u32(vvgpu_ctx* c, const u32* d, u32* h) {
    CK(cudaMemcpyAsync(c->h_pinned, d, sizeof(u32), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    *h = *(u32*)c->h_pinned;
    return 0;
}

int alloc_tree(vvgpu_ctx* c, size_t cap, size_t n, size_t nseg) {
    bool ok = true;
    c->t_x.get<double>(cap, &ok); c->t_y.get<double>(cap, &ok); c->t_h.get<double>(cap, &ok); c->t_w.get<double>(cap, &ok);
    c->t_bb.get<u64>(4 * cap, &ok);
    c->t_first.get<int>(cap, &ok); c->t_last.get<int>(cap, &ok); c->t_sfirst.get<int>(cap, &ok); c->t_slast.get<int>(cap, &ok);
    c->t_ch1.get<int>(cap, &ok); c->t_depth.get<int>(cap, &ok);
    c->t_status.get<unsigned char>(cap, &ok); c->t_axis.get<unsigned char>(cap, &ok);
    c->t_perm.get<int>(n, &ok); c->t_tmpR.get<int>(n, &ok);
    for (int k = 0; k < 2; k++) c->t_segperm[k].get<int>(nseg, &ok);
    c->scan_out.get<u32>(std::max(std::max(n, nseg), cap) + 2, &ok);
    c->flags.get<u32>(std::max(n, cap) + 2, &ok);
    NEED(ok);
    return 0;
}

// K1 (vvgpu_tree_build.cuh): top phase (cooperative) -> CTA-built subtrees -> top sweeps -> relocation; one read-back
int tree_build_impl(vvgpu_ctx* c, int far_criteria, double min_node, double max_node, unsigned mask) {
    const int n = (mask & 1u) ? (int)c->n : 0;
    const int nseg = (mask & 2u) ? c->nseg : 0;
    cudaStream_t st = c->stream;
    PSet& P = c->ps[c->cur];
    const size_t cap = 2 * ((size_t)n + nseg) + 2;
    int rc = alloc_tree(c, cap, n, nseg);
    if (rc) return rc;
    bool ok = true;
    c->t_nl.get<int>(cap, &ok); c->t_nn.get<int>(cap, &ok); c->t_lstart.get<int>(cap, &ok); c->t_pre.get<int>(cap, &ok);
    c->t_cmp.get<double>(3 * cap, &ok); c->t_cmm.get<double>(3 * cap, &ok);
    c->t_leafnode.get<int>(cap, &ok);
    BuildState* bs = c->build_state.get<BuildState>(1, &ok);
    NEED(ok);
    c->tn = n; c->tnseg = nseg; c->farc = (double)far_criteria;
    if (c->coop_grid == 0) {
        int sms = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        CK(cudaFuncSetAttribute(k_tree_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SubSmem)));
        c->coop_grid = std::min(sms, kTopThreads);   // the top phase scans one total per CTA with one block scan
        c->sub_grid = sms;
    }
    // splitting top-phase nodes of one level are disjoint and hold > kSubCap particles or > kSubSegCap segments
    const int maxact = n / kSubCap + nseg / kSubSegCap + 2;
    const size_t ntile_max = (size_t)n / kTopThreads + maxact + 2;
    const int tcmax = (int)((ntile_max + c->coop_grid - 1) / c->coop_grid) + 1;
    const size_t top_smem = top_smem_bytes(maxact, tcmax);
    if (top_smem > 160 * 1024) return fail(c, VVGPU_ELIMIT, "tree build: too many particles for the top-phase tables");
    if (top_smem > 40 * 1024) CK(cudaFuncSetAttribute(k_tree_top, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)top_smem));
    TopArgs A;
    A.T = c->T();
    A.bp = BuildParams{min_node, max_node};
    A.px = P.x.as<double>(); A.py = P.y.as<double>(); A.pg = P.g.as<double>();
    A.perm = c->t_perm.as<int>();
    A.sx = c->s_rx.as<double>(); A.sy = c->s_ry.as<double>();
    A.segperm = c->t_segperm[0].as<int>(); A.segtmp = c->t_segperm[1].as<int>();
    A.n = n; A.nseg = nseg;
    A.enc = c->b_enc.get<u32>((size_t)n + 1, &ok);
    A.tmpR = c->t_tmpR.as<int>();
    A.tilepre = c->b_tilepre.get<int>(ntile_max, &ok);
    A.chunktot = c->b_chunktot.get<int>(c->coop_grid, &ok);
    A.sublist = c->b_sublist.get<int>(((size_t)n + nseg) / 2 + 2, &ok);
    A.st = bs; A.cap = (long long)cap; A.maxact = maxact; A.tcmax = tcmax;
    SubNode* scratch = c->b_scratch.get<SubNode>(cap, &ok);
    unsigned char* arena = c->b_arena.get<unsigned char>((size_t)c->sub_grid * sub_arena_bytes(), &ok);
    int* aux = c->b_aux.get<int>(3 * cap, &ok);
    int* subinfo = c->b_subinfo.get<int>(3 * (((size_t)n + nseg) / 2 + 2), &ok);
    NEED(ok);
    void* args[] = {&A};
    CK(cudaLaunchCooperativeKernel((void*)k_tree_top, dim3(c->coop_grid), dim3(kTopThreads), args, top_smem, st));
    c->launches++;
    SubArgs SA{A.T, A.bp, A.px, A.py, A.pg, A.perm, A.sx, A.sy, A.segperm, scratch, A.sublist, bs, arena};
    k_tree_sub<<<c->sub_grid, kSubThreads, sizeof(SubSmem), st>>>(SA); CKLAUNCH();
    // the rest of each TObj follows the permutation (g included: the build moved only x, y and the caller index)
    PSet& Q = c->ps[c->cur ^ 1];
    if (n) {
        if (!Q.ensure(c->n)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
        k_tree_gather_rest<<<cdiv(n, 256), 256, 0, st>>>(n, c->t_perm.as<int>(), P.g.as<double>(), P.vx.as<double>(), P.vy.as<double>(),
                                                        P.ie.as<double>(), P.orig.as<int>(), Q.g.as<double>(), Q.vx.as<double>(),
                                                        Q.vy.as<double>(), Q.ie.as<double>(), Q.orig.as<int>()); CKLAUNCH();
    }
    const size_t nsubmax = ((size_t)n + nseg) / 2 + 2;
    SweepArgs WA{A.T, A.px, A.py, n ? Q.g.as<double>() : nullptr, scratch, A.sublist, aux, aux + cap, aux + 2 * cap, subinfo,
                 subinfo + nsubmax, subinfo + 2 * nsubmax, bs};
    k_tree_topsweep<<<1, 1024, 0, st>>>(WA); CKLAUNCH();
    k_tree_relocate<<<c->sub_grid * 8, 256, 0, st>>>(A.T, scratch, A.sublist, subinfo, subinfo + nsubmax, subinfo + 2 * nsubmax, bs); CKLAUNCH();
    const int nranks = c->comm.nranks;
    if (nranks > 1) {
        const size_t pmax = ((size_t)n + nseg) / kGroupLeaves + 8;
        c->sh_first.get<int>(pmax, &ok); c->sh_cnt.get<int>(pmax, &ok); c->sh_off.get<int>(pmax, &ok);
        int* rc_dev = c->sh_rankcnt.get<int>(kMaxRanks, &ok);
        NEED(ok);
        k_shard_table<<<1, 1024, 0, st>>>(A.T, bs, nranks, c->shard_table(), rc_dev); CKLAUNCH();
        CK(cudaMemcpyAsync(c->h_pinned + 128, rc_dev, nranks * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaMemcpyAsync(c->h_pinned + 32, bs, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(stream_sync(c));
    const int* hb = c->h_pinned + 32;
    if (hb[3]) {
        // the build moved (x, y) in place: put them back in the caller's order so that the resident list stays valid
        if (n) {
            k_tree_unpermute<<<cdiv(n, 256), 256, 0, st>>>(n, c->t_perm.as<int>(), P.x.as<double>(), P.y.as<double>(), Q.x.as<double>(),
                                                          Q.y.as<double>()); CKLAUNCH();
            std::swap(P.x, Q.x); std::swap(P.y, Q.y);
            CK(stream_sync(c));
        }
        if (hb[3] == 2) return fail(c, VVGPU_ELIMIT, "tree build: a top level is wider than its tables");
        return fail(c, VVGPU_ELIMIT, "tree deeper than 4096 levels or node capacity exceeded (degenerate input)");
    }
    c->nnodes = hb[0]; c->depth = hb[1];
    const u32 nl = (u32)hb[2];
    c->h_hist.resize(c->depth + 2);
    CK(cudaMemcpyAsync(c->h_hist.data(), bs->hist, sizeof(int) * (c->depth + 1), cudaMemcpyDeviceToHost, st));
    CK(stream_sync(c));
    c->nleaves = (int)nl;
    c->ngroups = cdiv(c->nleaves, kGroupLeaves);
    c->npieces = cdiv(c->ngroups, kShardBlock);
    c->ngmine = shard_count(c->ngroups, c->comm.rank, nranks);
    c->xL = 1;
    for (int r = 0; r < nranks && nranks > 1; r++) c->xL = std::max<long long>(c->xL, c->h_pinned[128 + r]);
    c->v_dirty = false;
    TreeDev T = c->T();
    if (n) { std::swap(P.g, Q.g); std::swap(P.vx, Q.vx); std::swap(P.vy, Q.vy); std::swap(P.ie, Q.ie); std::swap(P.orig, Q.orig); }
    // leaf records
    c->l_first.get<int>(nl, &ok); c->l_last.get<int>(nl, &ok); c->l_sfirst.get<int>(nl, &ok); c->l_slast.get<int>(nl, &ok);
    c->l_cx.get<double>(nl, &ok); c->l_cy.get<double>(nl, &ok); c->l_h.get<double>(nl, &ok); c->l_w.get<double>(nl, &ok);
    c->l_node.get<int>(nl, &ok);
    NEED(ok);
    k_leaf_fill<<<cdiv(nl, 128), 128, 0, st>>>(T, c->Lv(), (int)nl); CKLAUNCH();
    c->built = true;
    c->lists_ready = false;
    return 0;
}

size_t trav_smem() { return (size_t)kTravWarps * ((size_t)kTravStack * sizeof(int2) + (4 * 32 + 32 * 6) * sizeof(double)); }

// Interaction lists + Taylor coefficients for the groups this rank owns in ONE tree walk per group, then the
// work-unit table (vvgpu_lists.cuh)
int lists_impl(vvgpu_ctx* c) {
    const int ng = c->ngroups, nl = c->nleaves;
    if (c->lists_ready) return 0;
    cudaStream_t st = c->stream;
    bool ok = true;
    double* taylor = c->taylor.get<double>(4 * (size_t)nl, &ok);
    double* farcount = c->farcount.get<double>(nl, &ok);
    int* derr = c->d_err.get<int>(4, &ok);   // [0] error bits, [1] heavy groups
    int* hvlist = c->hv_list.get<int>(ng + 1, &ok);
    NEED(ok);
    TreeDev T = c->T();
    LeafDev L = c->Lv();
    CK(cudaFuncSetAttribute(k_traverse<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem()));
    // heavy groups are cut at the first tree level that is at least 256 nodes wide
    int cut = c->depth;
    for (int d = 0; d <= c->depth; d++)
        if (c->h_hist[d] >= 256) { cut = d; break; }
    const int item_cap = std::max(2, c->h_hist[cut]);
    const long long nreg_slots = (long long)ng * kGroupSlots;
    for (int attempt = 0;; attempt++) {
        if (attempt > 6) return fail(c, VVGPU_ELIMIT, "interaction lists do not fit the entry pool");
        if (c->pool_cap == 0) c->pool_cap = 48ll * nl + (4ll << 20);
        int* pleaf = c->g_leaf.get<int>((size_t)c->pool_cap, &ok);
        u32* pmask = c->g_mask.get<u32>((size_t)c->pool_cap, &ok);
        unsigned long long* cursor = c->g_cursor.get<unsigned long long>(1, &ok);
        long long* sbase = c->slot_base.get<long long>((size_t)nreg_slots + 1, &ok);
        int* scount = c->slot_count.get<int>((size_t)nreg_slots + 1, &ok);
        NEED(ok);
        CK(cudaMemsetAsync(derr, 0, 4 * sizeof(int), st));
        CK(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(scount, 0, sizeof(int) * (nreg_slots + 1), st));
        // chunk = work unit of the near kernels: long chunks save the per-unit prologue and the parking of partial
        // states (most groups then are ONE unit), short ones keep every SM busy when there are few groups
        const int unit = ng >= 2000 ? kUnitEntries : (ng >= 600 ? std::min(512, kUnitEntries) : 256);
        TravOut O{GroupLists{pleaf, pmask}, cursor, c->pool_cap, sbase, scount, derr, unit};
        TravItems I0{nullptr, nullptr, nullptr, item_cap, cut};
        if (c->ngmine > 0) {
            k_traverse_cta<0><<<c->ngmine, kTcWarps * 32, 0, st>>>(T, L, nl, c->shard(), c->ngmine, c->farc, O, taylor, farcount, hvlist,
                                                                   derr + 1, nullptr, nullptr, 0, I0); CKLAUNCH();
        }
        CK(cudaMemcpyAsync(c->h_pinned + 64, derr, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(stream_sync(c));
        int errbits = c->h_pinned[64];
        const int nheavy = c->h_pinned[65];
        long long nheavy_slots = 0;
        if (!errbits && nheavy) {
            // ---- fringe groups: top walk -> items -> one walk per item
            nheavy_slots = (long long)nheavy * (item_cap + 1) * kItemSlots;
            sbase = c->slot_base.get_keep<long long>((size_t)(nreg_slots + nheavy_slots) + 1, &ok, st);
            scount = c->slot_count.get_keep<int>((size_t)(nreg_slots + nheavy_slots) + 1, &ok, st);
            int* inode = c->hv_inode.get<int>((size_t)nheavy * item_cap, &ok);
            u32* imask = c->hv_imask.get<u32>((size_t)nheavy * item_cap, &ok);
            int* icount = c->hv_icount.get<int>(nheavy, &ok);
            double* tpart = c->hv_tpart.get<double>((size_t)nheavy * (item_cap + 1) * 32 * 5, &ok);
            NEED(ok);
            CK(cudaMemsetAsync(scount + nreg_slots, 0, sizeof(int) * (nheavy_slots + 1), st));
            CK(cudaMemsetAsync(icount, 0, sizeof(int) * nheavy, st));
            TravOut OH{GroupLists{pleaf, pmask}, cursor, c->pool_cap, sbase + nreg_slots, scount + nreg_slots, derr, unit};
            TravItems I{inode, imask, icount, item_cap, cut};
            k_traverse<1><<<cdiv(nheavy, kTravWarps), kTravWarps * 32, trav_smem(), st>>>(
                T, L, nl, 0, 0, c->farc, OH, taylor, farcount, tpart, hvlist, nheavy, nullptr, nullptr, I); CKLAUNCH();
            k_traverse_cta<2><<<nheavy * item_cap, kTcWarps * 32, 0, st>>>(T, L, nl, c->shard(), 0, c->farc, OH, taylor, farcount, nullptr, nullptr,
                                                                          tpart, hvlist, nheavy, I); CKLAUNCH();
            k_heavy_taylor<<<nheavy, 256, 0, st>>>(hvlist, nheavy, nl, tpart, icount, item_cap, taylor, farcount); CKLAUNCH();
            {
                int* hoff = c->hv_off.get<int>(2 * (size_t)nheavy_slots + 2, &ok);
                NEED(ok);
                k_heavy_pack<<<nheavy, 1024, 0, st>>>(nheavy, item_cap, OH, hoff); CKLAUNCH();
            }
            CK(cudaMemcpyAsync(c->h_pinned + 64, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(stream_sync(c));
            errbits = c->h_pinned[64];
        }
        if (errbits & 1) { c->pool_cap *= 2; continue; }   // pool too small: grow and walk again
        if (errbits & 2) return fail(c, VVGPU_ELIMIT, "near/far traversal stack overflow");
        if (errbits & 4) return fail(c, VVGPU_ELIMIT, "a heavy-group item overflowed its chunk slots");
        // ---- unit table: the non-empty slots in slot order
        const long long nslots_total = nreg_slots + nheavy_slots;
        u32* rank = c->scan_out.get<u32>((size_t)nslots_total + 2, &ok);
        int* ufirst = c->u_first.get<int>(ng + 1, &ok);
        int* unum = c->u_num.get<int>(ng + 1, &ok);
        u32* nsl = c->u_tmp.get<u32>((size_t)ng + 2, &ok);
        u32* usb = c->u_sbase.get<u32>(ng + 1, &ok);
        NEED(ok);
        int rc = scan_flags(c, SlotFlag{scount}, nslots_total, rank);
        if (rc) return rc;
        u32 nunits = 0;
        rc = read_u32(c, rank + nslots_total, &nunits);
        if (rc) return rc;
        c->nunits = (int)nunits;
        {   // fewer than ~5 waves of units (3 CTAs per SM): split the target leaves of every unit over 2 or 4 CTAs
            static const int forced = getenv("VV_TSPLIT") ? atoi(getenv("VV_TSPLIT")) : 0;
            const int slots = 3 * c->sm_count;
            c->tsplit = forced ? forced : (c->nunits >= 5 * slots ? 1 : (c->nunits >= 5 * slots / 2 ? 2 : 4));
        }
        int* ugroup = c->u_group.get<int>(std::max<u32>(nunits, 1), &ok);
        long long* ubase = c->u_base.get<long long>(std::max<u32>(nunits, 1), &ok);
        int* ucount = c->u_count.get<int>(std::max<u32>(nunits, 1), &ok);
        NEED(ok);
        k_units_fill<<<cdiv(nslots_total, 256), 256, 0, st>>>(nslots_total, nreg_slots, rank, sbase, scount, hvlist, item_cap,
                                                             ugroup, ubase, ucount); CKLAUNCH();
        k_units_groups<<<cdiv(ng, 128), 128, 0, st>>>(ng, rank, ufirst, unum); CKLAUNCH();
        if (nheavy) { k_units_heavy<<<cdiv(nheavy, 128), 128, 0, st>>>(nheavy, nreg_slots, item_cap, hvlist, rank, ufirst, unum); CKLAUNCH(); }
        k_unit_slots<<<cdiv(ng, 128), 128, 0, st>>>(L, nl, ng, unum, nsl); CKLAUNCH();
        rc = scan_flags(c, FlagArray{nsl}, ng, usb);
        if (rc) return rc;
        // CTA table (vvgpu_lists.cuh): costly units split over several CTAs and started first
        c->ncta = 0; c->cta_heavy = 0;
        u32 nslots = 0;
        static const bool no_table = getenv("VV_NO_CTA_TABLE") != nullptr;
        if (nunits > 0 && !no_table) {
            static const int cta_fmax = getenv("VV_CTA_MAXF") ? atoi(getenv("VV_CTA_MAXF")) : 4;
            static const int cta_div = getenv("VV_CTA_DIV") ? atoi(getenv("VV_CTA_DIV")) : 3;
            const size_t ctacap = (size_t)std::max(cta_fmax, c->tsplit) * nunits;
            int* ctau = c->cta_unit.get<int>(ctacap, &ok);
            unsigned short* ctap = c->cta_part.get<unsigned short>(ctacap, &ok);
            u32* cout = c->cta_out.get<u32>(4, &ok);
            NEED(ok);
            k_cta_table<<<1, 1024, 0, st>>>((int)nunits, L, nl, ugroup, ucount, 3 * c->sm_count, c->tsplit, cta_fmax, cta_div, ctau, ctap, usb + ng, cout); CKLAUNCH();
            CK(cudaMemcpyAsync(c->h_pinned, cout, 3 * sizeof(u32), cudaMemcpyDeviceToHost, st));
            CK(stream_sync(c));
            c->ncta = (int)((u32*)c->h_pinned)[0]; c->cta_heavy = (int)((u32*)c->h_pinned)[1]; nslots = ((u32*)c->h_pinned)[2];
        } else {
            rc = read_u32(c, usb + ng, &nslots);
            if (rc) return rc;
        }
        c->nslots = nslots;
        c->near_scratch.get<unsigned char>((size_t)nslots * sizeof(DiffOp::Part), &ok);
        NEED(ok);
        break;
    }
    c->lists_ready = true;
    return 0;
}

inline int near_grid(vvgpu_ctx* c) { return (c->ncta > 0 && !c->cta_off) ? c->ncta : c->nunits * c->tsplit; }

template <class Op>
int launch_near(vvgpu_ctx* c, Op op, const unsigned char* dyn = nullptr, const double4* tl = nullptr) {
    static_assert(sizeof(typename Op::Part) <= sizeof(DiffOp::Part), "scratch is sized for the largest Part");
    if (c->nunits <= 0) return 0;
    {
        bool ok = true;
        double4* s4 = c->src4.get<double4>((size_t)c->tn + 1, &ok);
        NEED(ok);
        k_pack_src<Op><<<cdiv(c->tn, 256), 256, 0, c->stream>>>(c->tn, c->ps[c->cur].view(), dyn, s4, tl); CKLAUNCH();
    }
    CK(cudaFuncSetAttribute(k_near<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LwSharedT<Op>)));
    k_near<Op><<<near_grid(c), kLwThreads, sizeof(LwSharedT<Op>), c->stream>>>(c->near_args(), op); CKLAUNCH();
    if (c->nslots > 0) {  // some group has more than one unit
        k_near_finalize<Op><<<c->ngmine, 256, 0, c->stream>>>(c->near_args(), op, c->shard(), c->ngmine); CKLAUNCH();
    }
    return 0;
}

// K4: dedicated kernel (vvgpu_conv.cuh); record tn of the packed sources is the dummy (g = 0) that pads the index lists
int launch_conv(vvgpu_ctx* c, ConvOp op) {
    if (c->nunits <= 0) return 0;
    bool ok = true;
    double4* s4 = c->src4.get<double4>((size_t)c->tn + 1, &ok);
    NEED(ok);
    k_pack_src<ConvOp><<<cdiv(c->tn + 1, 256), 256, 0, c->stream>>>(c->tn, c->ps[c->cur].view(), nullptr, s4); CKLAUNCH();
    CK(cudaFuncSetAttribute(k_conv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CvShared)));
    k_conv<<<near_grid(c), kCvThreads, sizeof(CvShared), c->stream>>>(c->near_args(), op, c->tn); CKLAUNCH();
    if (c->nslots > 0) {  // some group has more than one unit
        k_near_finalize<ConvOp><<<c->ngmine, 256, 0, c->stream>>>(c->near_args(), op, c->shard(), c->ngmine); CKLAUNCH();
    }
    return 0;
}

// K5: dedicated kernel (vvgpu_diff.cuh)
int launch_diff(vvgpu_ctx* c, DiffOp op) {
    if (c->nunits <= 0) return 0;
    bool ok = true;
    double4* s4 = c->src4.get<double4>((size_t)c->tn + 1, &ok);
    NEED(ok);
    double2* xy = c->src2.get<double2>(2 * ((size_t)c->tn + 1), &ok);
    NEED(ok);
    double2* xyn = xy + (size_t)c->tn + 1;
    k_pack_diff<<<cdiv(c->tn + 1, 256), 256, 0, c->stream>>>(c->tn, c->ps[c->cur].view(), s4, xy, xyn); CKLAUNCH();
    CK(cudaFuncSetAttribute(k_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfShared)));
    k_diff<<<near_grid(c), kDfThreads, sizeof(DfShared), c->stream>>>(c->near_args(), op, xy, xyn, c->tn); CKLAUNCH();
    if (c->nslots > 0) {  // some group has more than one unit
        k_near_finalize<DiffOp><<<c->ngmine, 256, 0, c->stream>>>(c->near_args(), op, c->shard(), c->ngmine); CKLAUNCH();
    }
    return 0;
}

// ---- multi-GPU exchange (vvgpu_shard.cuh, vvgpu_comm.h) ------------------------------------------------------
int comm_gather(vvgpu_ctx* c, const void* send, void* recv, size_t bytes) {
    const std::string e = comm_allgather(c->comm, c->device, c->stream, send, recv, bytes);
    if (!e.empty()) return fail(c, VVGPU_ECUDA, e);
    return 0;
}
// Every array of X holds, for the particles this rank owns, the result of the phase that just ran; on return it
// holds every rank's. ONE all-gather: owned pieces packed densely, blocks padded to the largest rank's count.
// `scalar` (device int, optional) is summed over the ranks into scalar_out.
int exchange_owned(vvgpu_ctx* c, XArrays X, const int* scalar, int* scalar_out) {
    const int P = c->comm.nranks;
    if (P <= 1) return 0;
    const long long L = (c->xL + 1) & ~1ll;
    const long long stride = xarrays_layout(X, L) + 8;     // bytes per rank; the last 8 carry the scalar
    bool ok = true;
    unsigned char* send = (unsigned char*)c->xsend.get<u64>((size_t)stride / 8, &ok);
    unsigned char* recv = (unsigned char*)c->xrecv.get<u64>((size_t)stride / 8 * P, &ok);
    NEED(ok);
    cudaStream_t st = c->stream;
    const int mine = (c->npieces > c->comm.rank) ? (c->npieces - c->comm.rank + P - 1) / P : 0;
    if (mine > 0) { k_shard_pack<<<mine, 256, 0, st>>>(c->shard_table(), c->npieces, c->shard(), X, send); CKLAUNCH(); }
    CK(cudaMemsetAsync(send + stride - 8, 0, 8, st));
    if (scalar) CK(cudaMemcpyAsync(send + stride - 8, scalar, sizeof(int), cudaMemcpyDeviceToDevice, st));
    int rc = comm_gather(c, send, recv, (size_t)stride);
    if (rc) return rc;
    if (c->npieces > 0) { k_shard_unpack<<<c->npieces, 256, 0, st>>>(c->shard_table(), c->npieces, c->shard(), X, stride, recv); CKLAUNCH(); }
    if (scalar_out) { k_rank_sum_i32<<<1, 32, 0, st>>>((const int*)(recv + stride - 8), stride / 4, P, 1, scalar_out); CKLAUNCH(); }
    return 0;
}
// buf[0..n) summed over the ranks, in rank order (bit-identical on every rank)
int rank_sum_f64(vvgpu_ctx* c, double* buf, int n) {
    const int P = c->comm.nranks;
    if (P <= 1 || n <= 0) return 0;
    bool ok = true;
    double* recv = (double*)c->xrecv.get<u64>((size_t)n * P, &ok);
    NEED(ok);
    int rc = comm_gather(c, buf, recv, (size_t)n * sizeof(double));
    if (rc) return rc;
    k_rank_sum_f64<<<cdiv(n, 128), 128, 0, c->stream>>>(recv, n, P, n, buf); CKLAUNCH();
    return 0;
}
// v of the other ranks' targets (after the velocity phases of a sharded step)
int sync_v(vvgpu_ctx* c) {
    if (!c->v_dirty || c->comm.nranks <= 1) { c->v_dirty = false; return 0; }
    XArrays X{};
    X.n = 2; X.p[0] = c->ps[c->cur].vx.p; X.p[1] = c->ps[c->cur].vy.p; X.wide[0] = X.wide[1] = 1;
    int rc = exchange_owned(c, X, nullptr, nullptr);
    if (rc) return rc;
    c->v_dirty = false;
    return 0;
}

// per-leaf merge criterion / epsilon restriction / nearest segment (MEpsilonFast.cpp:26-47)
int wall_params(vvgpu_ctx* c, int merge, double* lcrit, double* lrestr, int* latt) {
    cudaStream_t st = c->stream;
    const int nl = c->nleaves;
    BodySegs B{c->tnseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_dlx.as<double>(),
               c->s_dly.as<double>(), c->b_first.as<int>()};
    u64 *bd = nullptr, *bk = nullptr;
    if (c->tnseg > 0) {
        bool ok = true;
        bd = c->wall_d.get<u64>(nl, &ok); bk = c->wall_key.get<u64>(nl, &ok);
        NEED(ok);
        const int* sp = c->t_segperm[c->segcur].as<int>();
        Units U = c->Uv();
        k_wall_init<<<cdiv(nl, 256), 256, 0, st>>>(nl, bd, bk); CKLAUNCH();
        if (c->nunits > 0) {   // a rank of a small tree may own no leaf group at all
            k_wall_pass<1><<<c->nunits, 256, 0, st>>>(c->Lv(), c->Gv(), U, c->nunits, sp, B, bd, bk); CKLAUNCH();
            k_wall_pass<2><<<c->nunits, 256, 0, st>>>(c->Lv(), c->Gv(), U, c->nunits, sp, B, bd, bk); CKLAUNCH();
        }
    }
    k_wall_finish<<<cdiv(nl, 128), 128, 0, st>>>(c->Lv(), nl, c->t_segperm[c->segcur].as<int>(), B, merge, bd, bk, lcrit, lrestr, latt); CKLAUNCH();
    return 0;
}

// source-leaf boxes for the exact pruning of EpsOp (covering the tentative merged positions of M)
int eps_boxes(vvgpu_ctx* c, MergeState M, const unsigned char* only = nullptr) {
    bool ok = true;
    double* lb = c->lbox.get<double>(5 * (size_t)c->nleaves, &ok);
    NEED(ok);
    k_leaf_box<<<cdiv(c->nleaves, 128), 128, 0, c->stream>>>(c->Lv(), c->nleaves, c->ps[c->cur].view(), M, lb, only); CKLAUNCH();
    return 0;
}

MergeState mstate(Buf* b) {
    return MergeState{b[0].as<int>(), b[1].as<int>(), b[2].as<int>(), b[3].as<double>(), b[4].as<double>(), b[5].as<double>()};
}

}  // namespace

// =================================================================================== C ABI
extern "C" {

const char* vvgpu_strerror(int code) {
    switch (code) {
        case VVGPU_OK: return "ok";
        case VVGPU_EINVAL: return "invalid argument";
        case VVGPU_ESTATE: return "call out of order (tree not built / already built)";
        case VVGPU_ECUDA: return "CUDA error";
        case VVGPU_ENOMEM: return "out of device memory";
        case VVGPU_ELIMIT: return "internal capacity exceeded";
        default: return "unknown error";
    }
}
const char* vvgpu_last_error(const vvgpu_ctx* c) { return c ? c->err.c_str() : ""; }

int vvgpu_create(int device, vvgpu_ctx** out) {
    if (!out) return VVGPU_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return VVGPU_ECUDA;  // no CPU fallback
    if (device < 0 || device >= ndev) return VVGPU_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return VVGPU_ECUDA;
    vvgpu_ctx* c = new vvgpu_ctx();
    c->device = device;
    if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) c->sm_count = 148;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return VVGPU_ECUDA; }
    for (int k = 0; k < VVGPU_T_COUNT; k++) { cudaEventCreate(&c->ev0[k]); cudaEventCreate(&c->ev1[k]); }
    if (cudaMallocHost((void**)&c->h_pinned, 1024) != cudaSuccess) { delete c; return VVGPU_ECUDA; }
    *out = c;
    return VVGPU_OK;
}

void vvgpu_destroy(vvgpu_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->comm.kind == Comm::NCCL && c->comm.nccl) { if (NcclApi* api = NcclApi::get(nullptr)) api->CommDestroy(c->comm.nccl); }
    if (c->comm.kind == Comm::LOCAL && c->comm.grp) {
        LocalGroup* g = c->comm.grp;
        cudaEventDestroy(g->slot[c->comm.rank].ready); cudaEventDestroy(g->slot[c->comm.rank].done);
        bool last;
        { std::lock_guard<std::mutex> lk(g->mu); last = (--g->refs == 0); }
        if (last) delete g;
    }
    Buf* all[] = {&c->stage, &c->s_rx, &c->s_ry, &c->s_cx, &c->s_cy, &c->s_dlx, &c->s_dly, &c->s_g, &c->s_ie, &c->s_slip,
                  &c->s_body, &c->b_first, &c->b_prop, &c->d_fric, &c->d_gsum, &c->d_fdt, &c->d_gdead, &c->d_cleaned,
                  &c->t_x, &c->t_y, &c->t_h, &c->t_w, &c->t_bb, &c->t_first, &c->t_last, &c->t_sfirst, &c->t_slast,
                  &c->t_ch1, &c->t_depth, &c->t_status, &c->t_axis, &c->t_nl, &c->t_nn, &c->t_lstart,
                  &c->t_pre, &c->t_cmp, &c->t_cmm, &c->t_leafnode,
                  &c->t_segperm[0], &c->t_segperm[1], &c->t_perm, &c->t_tmpR, &c->scan_part, &c->scan_out, &c->flags, &c->build_state, &c->b_enc, &c->b_tilepre, &c->b_chunktot, &c->b_sublist, &c->b_scratch, &c->b_arena, &c->b_aux, &c->b_subinfo,
                  &c->l_first, &c->l_last, &c->l_sfirst, &c->l_slast, &c->l_cx, &c->l_cy, &c->l_h, &c->l_w, &c->l_node,
                  &c->g_leaf, &c->g_mask, &c->g_cursor, &c->slot_base, &c->slot_count, &c->u_base, &c->u_count, &c->u_num, &c->hv_inode, &c->hv_imask, &c->hv_icount, &c->hv_tpart, &c->hv_off, &c->taylor, &c->farcount, &c->d_err, &c->lcrit,
                  &c->lrestr, &c->latt, &c->leaf_dirty, &c->leaf_dbox, &c->unit_dirty, &c->group_dirty, &c->tl, &c->cta_unit, &c->cta_part, &c->cta_out, &c->ie_tmp, &c->dyn, &c->d_changed, &c->d_nmerged, &c->d_sinks, &c->d_pairs,
                  &c->sh_first, &c->sh_cnt, &c->sh_off, &c->sh_rankcnt, &c->xsend, &c->xrecv, &c->u_group, &c->u_first, &c->u_sbase, &c->u_tmp, &c->near_scratch, &c->src4, &c->src2, &c->lbox, &c->wall_d, &c->wall_key, &c->hv_list};
    for (Buf* b : all) b->release();
    for (int k = 0; k < 6; k++) { c->mA[k].release(); c->mB[k].release(); }
    c->ps[0].release(); c->ps[1].release(); c->ps_backup.release();
    c->pt_xy.release(); c->pt_out.release(); c->pt_v.release();
    for (int k = 0; k < VVGPU_T_COUNT; k++) { cudaEventDestroy(c->ev0[k]); cudaEventDestroy(c->ev1[k]); }
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    cudaStreamDestroy(c->stream);
    delete c;
}

static int set_particles_common(vvgpu_ctx* c, int list, const void* src, size_t n, int rec_doubles) {
    if (!c || list != VVGPU_LIST_VORTEX || (!src && n)) return fail(c, VVGPU_EINVAL, "set_particles: bad argument");
    if (n > (size_t)std::numeric_limits<int>::max() / 4) return fail(c, VVGPU_EINVAL, "set_particles: too many particles");
    if (c->built) return fail(c, VVGPU_ESTATE, "set_particles while the tree is built (the tree holds positions into the list)");
    CK(cudaSetDevice(c->device));
    bool ok = true;
    double* st = c->stage.get<double>(n * rec_doubles, &ok);
    if (!ok || !c->ps[c->cur].ensure(n)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
    c->n = n;
    c->orig_next = n;
    if (n) {
        // cudaMemcpyDefault: the records may also sit in device memory (a multi-GPU host layer that uploads one slice
        // per rank and all-gathers over NVLink hands over a device buffer)
        CK(cudaMemcpyAsync(st, src, n * rec_doubles * sizeof(double), cudaMemcpyDefault, c->stream));
        if (rec_doubles == 6) k_unpack48<<<cdiv(n, 256), 256, 0, c->stream>>>((int)n, st, c->ps[c->cur].view(), c->ps[c->cur].orig.as<int>());
        else k_unpack24<<<cdiv(n, 256), 256, 0, c->stream>>>((int)n, st, c->ps[c->cur].view(), c->ps[c->cur].orig.as<int>());
        CKLAUNCH();
        CK(stream_sync(c));  // the caller may reuse `src`
    }
    return 0;
}
int vvgpu_set_particles(vvgpu_ctx* c, int list, const vvgpu_obj* objs, size_t n) {
    return set_particles_common(c, list, objs, n, 6);
}
int vvgpu_set_particles_xyg(vvgpu_ctx* c, int list, const double* xyg, size_t n) {
    return set_particles_common(c, list, xyg, n, 3);
}
int vvgpu_append_particles(vvgpu_ctx* c, int list, const vvgpu_obj* objs, size_t n) {
    if (!c || list != VVGPU_LIST_VORTEX || (!objs && n)) return fail(c, VVGPU_EINVAL, "append_particles: bad argument");
    if (c->built) return fail(c, VVGPU_ESTATE, "append_particles while the tree is built (the tree holds positions into the list)");
    if (n == 0) return 0;
    if (c->n + n > (size_t)std::numeric_limits<int>::max() / 4) return fail(c, VVGPU_EINVAL, "append_particles: too many particles");
    CK(cudaSetDevice(c->device));
    bool ok = true;
    double* st = c->stage.get<double>(n * 6, &ok);
    // grow with headroom: a shedding body appends a few hundred vortices every step
    const size_t want = c->n + n;
    if (!ok || !c->ps[c->cur].ensure_keep(want + want / 8, c->stream)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
    CK(cudaMemcpyAsync(st, objs, n * 48, cudaMemcpyHostToDevice, c->stream));
    k_unpack48_at<<<cdiv(n, 256), 256, 0, c->stream>>>((int)n, st, c->ps[c->cur].view(), c->ps[c->cur].orig.as<int>(), (int)c->n,
                                                       (int)c->orig_next); CKLAUNCH();
    CK(stream_sync(c));  // the caller may reuse `objs`
    c->n += n;
    c->orig_next += n;
    return 0;
}
int vvgpu_particle_count(vvgpu_ctx* c, int list, size_t* n) {
    if (!c || !n || list != VVGPU_LIST_VORTEX) return fail(c, VVGPU_EINVAL, "particle_count: bad argument");
    *n = c->n;
    return 0;
}
int vvgpu_get_particles(vvgpu_ctx* c, int list, vvgpu_obj* out, size_t cap, size_t* n) {
    if (!c || list != VVGPU_LIST_VORTEX) return fail(c, VVGPU_EINVAL, "get_particles: bad argument");
    if (n) *n = c->n;
    if (!out) return 0;
    if (cap < c->n) return fail(c, VVGPU_EINVAL, "get_particles: buffer too small");
    if (!c->n) return 0;
    CK(cudaSetDevice(c->device));
    if (c->built && c->v_dirty) { int rcv = sync_v(c); if (rcv) return rcv; }
    bool ok = true;
    double* st = c->stage.get<double>(c->n * 6, &ok);
    NEED(ok);
    k_pack48<<<cdiv(c->n, 256), 256, 0, c->stream>>>((int)c->n, c->ps[c->cur].view(), st); CKLAUNCH();
    CK(cudaMemcpyAsync(out, st, c->n * 48, cudaMemcpyDefault, c->stream));   // host or device destination
    CK(stream_sync(c));
    return 0;
}
int vvgpu_particle_gsum(vvgpu_ctx* c, int list, double* sum) {
    if (!c || list != VVGPU_LIST_VORTEX || !sum) return fail(c, VVGPU_EINVAL, "particle_gsum: bad argument");
    *sum = 0;
    if (!c->n) return 0;
    CK(cudaSetDevice(c->device));
    bool ok = true;
    double* d = c->d_pairs.get<double>(2, &ok);
    NEED(ok);
    k_sum_g<<<1, 1024, 0, c->stream>>>((int)c->n, c->ps[c->cur].g.as<double>(), d); CKLAUNCH();
    CK(cudaMemcpyAsync(sum, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    return 0;
}
int vvgpu_get_particles_range(vvgpu_ctx* c, int list, vvgpu_obj* out, size_t first, size_t count) {
    if (!c || list != VVGPU_LIST_VORTEX || (!out && count) || first + count > c->n) return fail(c, VVGPU_EINVAL, "get_particles_range: bad argument");
    if (!count) return 0;
    CK(cudaSetDevice(c->device));
    if (c->built && c->v_dirty) { int rcv = sync_v(c); if (rcv) return rcv; }
    bool ok = true;
    double* st = c->stage.get<double>(count * 6, &ok);
    NEED(ok);
    Particles p = c->ps[c->cur].view();
    p.x += first; p.y += first; p.g += first; p.vx += first; p.vy += first; p.ie += first;
    k_pack48<<<cdiv(count, 256), 256, 0, c->stream>>>((int)count, p, st); CKLAUNCH();
    CK(cudaMemcpyAsync(out, st, count * 48, cudaMemcpyDefault, c->stream));
    CK(stream_sync(c));
    return 0;
}
int vvgpu_get_permutation(vvgpu_ctx* c, int list, int32_t* orig, size_t cap) {
    if (!c || list != VVGPU_LIST_VORTEX || !orig || cap < c->n) return fail(c, VVGPU_EINVAL, "get_permutation: bad argument");
    CK(cudaSetDevice(c->device));
    if (c->n) CK(cudaMemcpyAsync(orig, c->ps[c->cur].orig.p, c->n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    return 0;
}

int vvgpu_set_bodies(vvgpu_ctx* c, const vvgpu_seg* segs, size_t nseg, const vvgpu_body* bodies, size_t nbody) {
    if (!c || (nseg && !segs) || (nbody && !bodies)) return fail(c, VVGPU_EINVAL, "set_bodies: bad argument");
    if (c->built) return fail(c, VVGPU_ESTATE, "set_bodies while the tree is built");
    CK(cudaSetDevice(c->device));
    std::vector<double> col[8];
    std::vector<int> slip(nseg), body(nseg), bfirst(nbody + 1, 0);
    for (auto& v : col) v.resize(nseg);
    for (size_t i = 0; i < nseg; i++) {
        const vvgpu_seg& s = segs[i];
        col[0][i] = s.rx; col[1][i] = s.ry; col[2][i] = s.cx; col[3][i] = s.cy; col[4][i] = s.dlx; col[5][i] = s.dly;
        col[6][i] = s.g; col[7][i] = s.ieps; slip[i] = s.slip; body[i] = s.body;
        if (s.body < 0 || (size_t)s.body >= nbody) return fail(c, VVGPU_EINVAL, "set_bodies: segment with bad body index");
    }
    std::vector<double> bprop(16 * std::max<size_t>(nbody, 1), 0.0);
    c->any_body_flow = false;
    for (size_t b = 0; b < nbody; b++) {
        const vvgpu_body& B = bodies[b];
        if (B.first_seg < 0 || B.n_seg < 0 || (size_t)B.first_seg + B.n_seg > nseg ||
            (b && B.first_seg != bfirst[b])) return fail(c, VVGPU_EINVAL, "set_bodies: bodies must tile the segment array in order");
        bfirst[b] = B.first_seg; bfirst[b + 1] = B.first_seg + B.n_seg;
        double* p = &bprop[16 * b];
        p[0] = B.axis_x; p[1] = B.axis_y; p[2] = B.cofm_x; p[3] = B.cofm_y; p[4] = B.bl_x; p[5] = B.bl_y;
        p[6] = B.tr_x; p[7] = B.tr_y; p[8] = B.disc_r2; p[9] = B.speed_x; p[10] = B.speed_y; p[11] = B.speed_o;
        p[12] = B.inside_valid ? 1 : 0;
        bool any_slip = false;
        for (int s = bfirst[b]; s < bfirst[b + 1]; s++) any_slip |= slip[s] != 0;
        p[13] = any_slip ? 1 : 0;
        bool moving = !(std::fabs(B.speed_x) + std::fabs(B.speed_y) + std::fabs(B.speed_o) < 1E-10);
        if (any_slip || moving) c->any_body_flow = true;
    }
    if (nbody && (size_t)bfirst[nbody] != nseg) return fail(c, VVGPU_EINVAL, "set_bodies: bodies must cover all segments");
    bool ok = true;
    Buf* dst[8] = {&c->s_rx, &c->s_ry, &c->s_cx, &c->s_cy, &c->s_dlx, &c->s_dly, &c->s_g, &c->s_ie};
    for (int k = 0; k < 8; k++) {
        double* d = dst[k]->get<double>(nseg, &ok);
        NEED(ok);
        if (nseg) CK(cudaMemcpyAsync(d, col[k].data(), nseg * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    int* d1 = c->s_slip.get<int>(nseg, &ok); int* d2 = c->s_body.get<int>(nseg, &ok);
    int* d3 = c->b_first.get<int>(nbody + 1, &ok); double* d4 = c->b_prop.get<double>(bprop.size(), &ok);
    c->d_fric.get<double>(nseg, &ok); c->d_gsum.get<double>(nseg, &ok);
    c->d_fdt.get<double>(3 * std::max<size_t>(nbody, 1), &ok); c->d_gdead.get<double>(std::max<size_t>(nbody, 1), &ok);
    NEED(ok);
    if (nseg) {
        CK(cudaMemcpyAsync(d1, slip.data(), nseg * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(d2, body.data(), nseg * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaMemcpyAsync(d3, bfirst.data(), (nbody + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d4, bprop.data(), bprop.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(stream_sync(c));
    c->nseg = (int)nseg; c->nbody = (int)nbody;
    return 0;
}

int vvgpu_tree_build(vvgpu_ctx* c, int far_criteria, double min_node, double max_node, unsigned include_mask) {
    if (!c) return VVGPU_EINVAL;
    if (c->built) return fail(c, VVGPU_ESTATE, "Tree is already built");
    CK(cudaSetDevice(c->device));
    int rc;
    {
        PhaseTimer t(c, VVGPU_T_BUILD);
        rc = tree_build_impl(c, far_criteria, min_node, max_node, include_mask);
    }
    if (rc) return rc;
    {
        PhaseTimer t(c, VVGPU_T_LISTS);
        rc = lists_impl(c);
    }
    if (rc) { c->built = false; return rc; }
    return 0;
}

int vvgpu_tree_destroy(vvgpu_ctx* c) {
    if (!c) return VVGPU_EINVAL;
    if (c->built && c->v_dirty) {   // the other ranks' share of v, while the shard table is still this tree's
        CK(cudaSetDevice(c->device));
        int rc = sync_v(c);
        if (rc) return rc;
    }
    c->built = false; c->lists_ready = false;
    c->nnodes = c->nleaves = c->ngroups = 0;
    return 0;
}

int vvgpu_tree_counts(vvgpu_ctx* c, size_t* n_nodes, size_t* n_leaves, size_t* depth) {
    if (!c) return VVGPU_EINVAL;
    if (!c->built) return fail(c, VVGPU_ESTATE, "tree is not built");
    if (n_nodes) *n_nodes = c->nnodes;
    if (n_leaves) *n_leaves = c->nleaves;
    if (depth) *depth = c->depth;
    return 0;
}

int vvgpu_tree_export(vvgpu_ctx* c, double* dbl, int64_t* idx, size_t cap_nodes) {
    if (!c || !dbl || !idx) return fail(c, VVGPU_EINVAL, "tree_export: bad argument");
    if (!c->built) return fail(c, VVGPU_ESTATE, "tree is not built");
    if (cap_nodes < (size_t)c->nnodes) return fail(c, VVGPU_EINVAL, "tree_export: buffer too small");
    CK(cudaSetDevice(c->device));
    bool ok = true;
    const size_t nn = c->nnodes;
    double* st = c->stage.get<double>(nn * 18, &ok);
    NEED(ok);
    long long* si = (long long*)(st + nn * 10);
    k_tree_export<<<cdiv(nn, 128), 128, 0, c->stream>>>(c->T(), (int)nn, st, si); CKLAUNCH();
    CK(cudaMemcpyAsync(dbl, st, nn * 10 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(idx, si, nn * 8 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    return 0;
}

int vvgpu_tree_lists(vvgpu_ctx* c, int64_t* near_ptr, int64_t* near_idx, size_t near_cap, int64_t* far_ptr,
                     int64_t* far_idx, size_t far_cap) {
    if (!c || !near_ptr || !far_ptr) return fail(c, VVGPU_EINVAL, "tree_lists: bad argument");
    if (!c->built) return fail(c, VVGPU_ESTATE, "tree is not built");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int nl = c->nleaves;
    bool ok = true;
    Buf bn, bf, bnp, bfp, bni, bfi;
    u32* ncount = bn.get<u32>(nl + 1, &ok); u32* fcount = bf.get<u32>(nl + 1, &ok);
    u32* nptr = bnp.get<u32>(nl + 2, &ok); u32* fptr = bfp.get<u32>(nl + 2, &ok);
    int* derr = c->d_err.get<int>(1, &ok);
    int rc = 0;
    std::vector<u32> hn(nl + 1), hf(nl + 1);
    auto cleanup = [&]() { bn.release(); bf.release(); bnp.release(); bfp.release(); bni.release(); bfi.release(); };
    if (!ok) { cleanup(); return fail(c, VVGPU_ENOMEM, "cudaMalloc failed"); }
    cudaMemsetAsync(derr, 0, sizeof(int), st);
    k_lists_dfs<false><<<cdiv(nl, 128), 128, 0, st>>>(c->T(), c->Lv(), nl, c->farc, ncount, fcount, nullptr, nullptr, nullptr, nullptr, derr);
    c->launches++;
    rc = scan_flags(c, FlagArray{ncount}, nl, nptr);
    if (!rc) rc = scan_flags(c, FlagArray{fcount}, nl, fptr);
    if (rc) { cleanup(); return rc; }
    cudaMemcpyAsync(hn.data(), nptr, sizeof(u32) * (nl + 1), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(hf.data(), fptr, sizeof(u32) * (nl + 1), cudaMemcpyDeviceToHost, st);
    int herr = 0;
    cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); return fail(c, VVGPU_ECUDA, "tree_lists: sync failed"); }
    if (herr) { cleanup(); return fail(c, VVGPU_ELIMIT, "tree_lists: tree deeper than the export stack"); }
    for (int i = 0; i <= nl; i++) { near_ptr[i] = hn[i]; far_ptr[i] = hf[i]; }
    if (near_idx && far_idx) {
        if (near_cap < hn[nl] || far_cap < hf[nl]) { cleanup(); return fail(c, VVGPU_EINVAL, "tree_lists: buffers too small"); }
        long long* ni = bni.get<long long>(hn[nl], &ok); long long* fi = bfi.get<long long>(hf[nl], &ok);
        if (!ok) { cleanup(); return fail(c, VVGPU_ENOMEM, "cudaMalloc failed"); }
        k_lists_dfs<true><<<cdiv(nl, 128), 128, 0, st>>>(c->T(), c->Lv(), nl, c->farc, ncount, fcount, nptr, fptr, ni, fi, derr);
        c->launches++;
        if (hn[nl]) cudaMemcpyAsync(near_idx, ni, sizeof(long long) * hn[nl], cudaMemcpyDeviceToHost, st);
        if (hf[nl]) cudaMemcpyAsync(far_idx, fi, sizeof(long long) * hf[nl], cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); return fail(c, VVGPU_ECUDA, "tree_lists: sync failed"); }
    }
    cleanup();
    return 0;
}

int vvgpu_tree_leaf_segments(vvgpu_ctx* c, int64_t* ptr, int64_t* idx, size_t cap) {
    if (!c || !ptr) return fail(c, VVGPU_EINVAL, "tree_leaf_segments: bad argument");
    if (!c->built) return fail(c, VVGPU_ESTATE, "tree is not built");
    CK(cudaSetDevice(c->device));
    const int nl = c->nleaves;
    std::vector<int> sf(nl), sl(nl), perm(c->tnseg);
    CK(cudaMemcpyAsync(sf.data(), c->l_sfirst.p, sizeof(int) * nl, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(sl.data(), c->l_slast.p, sizeof(int) * nl, cudaMemcpyDeviceToHost, c->stream));
    if (c->tnseg) CK(cudaMemcpyAsync(perm.data(), c->t_segperm[c->segcur].p, sizeof(int) * c->tnseg, cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    int64_t p = 0;
    for (int l = 0; l < nl; l++) {
        ptr[l] = p;
        for (int k = sf[l]; k < sl[l]; k++) {
            if (idx) { if ((size_t)p >= cap) return fail(c, VVGPU_EINVAL, "tree_leaf_segments: buffer too small"); idx[p] = perm[k]; }
            p++;
        }
    }
    ptr[nl] = p;
    return 0;
}

int vvgpu_count_interactions(vvgpu_ctx* c, double* near_pairs, double* far_nodes) {
    if (!c) return VVGPU_EINVAL;
    if (!c->built) return fail(c, VVGPU_ESTATE, "tree is not built");
    CK(cudaSetDevice(c->device));
    bool ok = true;
    const int nl = c->nleaves, nu = c->nunits;
    double* d = c->d_pairs.get<double>(nu + 1, &ok);
    NEED(ok);
    k_count_pairs<<<std::max(nu, 1), 128, 0, c->stream>>>(c->Lv(), nl, nu, c->u_group.as<int>(), c->u_base.as<long long>(),
                                                         c->u_count.as<int>(), c->Gv(), c->ps[c->cur].g.as<double>(), d); CKLAUNCH();
    std::vector<double> h(nu), f(nl);
    if (nu) CK(cudaMemcpyAsync(h.data(), d, sizeof(double) * nu, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(f.data(), c->farcount.p, sizeof(double) * nl, cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    double s = 0, t = 0;
    for (double v : h) s += v;
    // far counts exist for the leaves of this rank's groups only
    for (int m = 0; m < c->ngmine; m++) {
        const int g = c->shard().group(m);
        for (int l = g * kGroupLeaves; l < std::min(nl, (g + 1) * kGroupLeaves); l++) t += f[l];
    }
    if (c->comm.nranks > 1) {   // totals over the ranks
        double* two = c->d_pairs.get<double>(2, &ok);
        NEED(ok);
        const double hv[2] = {s, t};
        CK(cudaMemcpyAsync(two, hv, sizeof(hv), cudaMemcpyHostToDevice, c->stream));
        int rc = rank_sum_f64(c, two, 2);
        if (rc) return rc;
        double out[2];
        CK(cudaMemcpyAsync(out, two, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
        CK(stream_sync(c));
        s = out[0]; t = out[1];
    }
    if (near_pairs) *near_pairs = s;
    if (far_nodes) *far_nodes = t;
    return 0;
}

int vvgpu_epsilon(vvgpu_ctx* c, int merge, int* merged) {
    if (!c) return VVGPU_EINVAL;
    if (!c->built || !c->lists_ready) return fail(c, VVGPU_ESTATE, "tree is not built");
    CK(cudaSetDevice(c->device));
    PhaseTimer t(c, VVGPU_T_EPS);
    cudaStream_t st = c->stream;
    if (merged) *merged = 0;
    const int n = c->tn, nl = c->nleaves;
    if (n == 0) return 0;
    bool ok = true;
    PSet& P = c->ps[c->cur];
    const bool walls = c->tnseg > 0;
    const bool multi = c->comm.nranks > 1;
    // per-leaf wall parameters: every rank needs them for the leaves of its own targets only, which are exactly the
    // leaves its work units cover
    double *lcrit = nullptr, *lrestr = nullptr;
    if (walls || c->nbody > 0 || merge) {
        lcrit = c->lcrit.get<double>(nl, &ok); lrestr = c->lrestr.get<double>(nl, &ok);
        int* latt = c->latt.get<int>(nl, &ok);
        NEED(ok);
        int rcw = wall_params(c, merge, lcrit, lrestr, latt);
        if (rcw) return rcw;
    }
    int* dchg = c->d_changed.get<int>(2, &ok);
    NEED(ok);
    auto gather_ie = [&](double* ie) -> int {
        XArrays X{};
        X.n = 1; X.p[0] = ie; X.wide[0] = 1;
        return exchange_owned(c, X, nullptr, nullptr);
    };
    if (!merge) {
        int rcb = eps_boxes(c, MergeState{});
        if (rcb) return rcb;
        EpsOp<false> op{MergeState{}, MergeState{}, nullptr, lrestr, nullptr, P.ie.as<double>(), dchg};
        int rc = launch_near(c, op);
        if (!rc && multi) rc = gather_ie(P.ie.as<double>());
        return rc;
    }
    // merging is order-dependent: iterate the tentative solution to its fixed point (see MergeState in
    // vvgpu_near.cuh). Every rank recomputes the decisions of ITS targets; the solution's columns are gathered after
    // every round, so all ranks iterate on the same state and stop in the same round.
    double* ietmp = c->ie_tmp.get<double>(n, &ok);
    unsigned char* dyn = c->dyn.get<unsigned char>(n, &ok);
    for (int k = 0; k < 3; k++) { c->mA[k].get<int>(n, &ok); c->mB[k].get<int>(n, &ok); }
    for (int k = 3; k < 6; k++) { c->mA[k].get<double>(n, &ok); c->mB[k].get<double>(n, &ok); }
    NEED(ok);
    CK(cudaMemcpyAsync(ietmp, P.ie.p, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    bool haveA = false;
    int rounds = 0;
    unsigned char* ldirty = c->leaf_dirty.get<unsigned char>(nl, &ok);
    double* ldbox = c->leaf_dbox.get<double>(4 * (size_t)nl, &ok);
    unsigned char* udirty = c->unit_dirty.get<unsigned char>((size_t)std::max(c->nunits, 1), &ok);
    unsigned char* gdirty = c->group_dirty.get<unsigned char>((size_t)std::max(c->ngroups, 1), &ok);
    double4* tl = c->tl.get<double4>(2 * (size_t)n, &ok);
    NEED(ok);
    u32 prev_changed = 0;
    for (;; rounds++) {
        if (rounds > n + 2) return fail(c, VVGPU_ELIMIT, "merge fixed point did not converge");
        MergeState A = haveA ? mstate(c->mA) : MergeState{};
        MergeState B = mstate(c->mB);
        // Later rounds recompute only the targets that see a changed entry; the others keep last round's outcome. While
        // most entries still change (the first rounds of a wake full of merges) nearly every target sees one, and looking
        // for them costs more than it saves: those rounds recompute everything.
        const bool incremental = haveA && prev_changed <= (u32)nl / 8;
        if (haveA) { k_merge_copy<<<cdiv(n, 256), 256, 0, st>>>(n, A, B); CKLAUNCH(); }
        else { k_merge_clear<<<cdiv(n, 256), 256, 0, st>>>(n, B); CKLAUNCH(); }
        CK(cudaMemsetAsync(dchg, 0, 2 * sizeof(int), st));
        if (incremental && c->nunits > 0) {
            CK(cudaMemsetAsync(gdirty, 0, (size_t)c->ngroups, st));
            k_unit_dirty<<<c->nunits, 128, 0, st>>>(c->nunits, c->Uv(), c->Gv(), ldirty, udirty, gdirty); CKLAUNCH();
        }
        // the boxes of the leaves whose particles kept their entries are those of the last round
        int rc = eps_boxes(c, A, haveA ? ldirty : nullptr);
        if (rc) return rc;
        auto run_round = [&](auto op) -> int {
            op.leaf_dirty = incremental ? ldirty : nullptr;
            op.leaf_dbox = ldbox;
            op.tl = haveA ? tl : nullptr;
            if (incremental && c->nunits > 0) { op.unit_dirty = udirty; op.group_dirty = gdirty; }
            // few units have anything to do in an incremental round, and the ones that always do (the fringe groups of
            // several units, which see the whole tree) are the slowest: spread each over 4 CTAs (same bits, NearArgs::tsplit)
            const int ts = c->tsplit;
            if (incremental) { c->tsplit = 4; c->cta_off = true; }
            const int rcn = launch_near(c, op, haveA ? dyn : nullptr, haveA ? tl : nullptr);
            c->tsplit = ts; c->cta_off = false;
            return rcn;
        };
        rc = haveA ? run_round(EpsOp<false, true>{A, B, lcrit, lrestr, dyn, ietmp, dchg})
                   : run_round(EpsOp<false, false>{A, B, lcrit, lrestr, dyn, ietmp, dchg});
        if (rc) return rc;
        if (multi) {
            // what travels: with whom every particle merges (-1: it does not) and its epsilon. The merged state of an
            // initiator follows from those and the assumed solution A, which every rank holds: recomputed locally.
            XArrays X{};
            X.n = 2;
            X.p[0] = B.part; X.p[1] = ietmp;
            X.wide[0] = 0; X.wide[1] = 1;
            rc = exchange_owned(c, X, dchg, dchg);
            if (rc) return rc;
            k_fill_i32<<<cdiv(n, 256), 256, 0, st>>>(n, B.absby, kNoAbs); CKLAUNCH();
            k_merge_fill<<<cdiv(n, 256), 256, 0, st>>>(n, P.view(), A, B); CKLAUNCH();
        } else {
            // absorbed-by follows from the (init, part) columns
            k_fill_i32<<<cdiv(n, 256), 256, 0, st>>>(n, B.absby, kNoAbs); CKLAUNCH();
            k_merge_absby<<<cdiv(n, 256), 256, 0, st>>>(n, B.init, B.part, B.absby); CKLAUNCH();
        }
        u32 changed = 0;
        rc = read_u32(c, (u32*)dchg, &changed);
        if (rc) return rc;
        if (getenv("VVGPU_DEBUG_MERGE")) {
            u32 redo = 0;
            read_u32(c, (u32*)dchg + 1, &redo);
            fprintf(stderr, "[vvgpu] merge round %d: %u decisions changed, %u target batches of %d leaves recomputed (0 = all)\n", rounds, changed, redo, nl);
        }
        if (!changed) break;  // B reproduces A
        prev_changed = changed;
        static const bool no_prune = getenv("VV_NO_MERGE_PRUNE") != nullptr;
        if (!haveA && !no_prune) {
            // round 0 assumed no merges: cancel the initiators that an earlier one absorbs (k_prune_*, vvgpu_shard.cuh);
            // `dyn` is free until k_merge_dyn below
            for (int sweep = 0; sweep < 3; sweep++) {
                k_prune_valid<<<cdiv(n, 256), 256, 0, st>>>(n, B.init, B.absby, dyn); CKLAUNCH();
                k_fill_i32<<<cdiv(n, 256), 256, 0, st>>>(n, B.absby, kNoAbs); CKLAUNCH();
                k_prune_absby<<<cdiv(n, 256), 256, 0, st>>>(n, dyn, B.part, B.absby); CKLAUNCH();
            }
            k_prune_commit<<<cdiv(n, 256), 256, 0, st>>>(n, dyn, B.init, B.part); CKLAUNCH();
            prev_changed = 0xffffffffu;   // the next round starts from a guess, not from this round's outcome: recompute all
        }
        k_leaf_dirty_clear<<<cdiv(nl, 256), 256, 0, st>>>(nl, ldirty, ldbox); CKLAUNCH();
        k_leaf_dirty<<<cdiv(n, 256), 256, 0, st>>>(c->Lv(), nl, n, P.view(), A, B, ldirty, ldbox); CKLAUNCH();
        for (int k = 0; k < 6; k++) std::swap(c->mA[k], c->mB[k]);
        haveA = true;
        k_merge_dyn<<<cdiv(n, 256), 256, 0, st>>>(n, P.view(), mstate(c->mA), dyn, tl); CKLAUNCH();
    }
    c->merge_rounds = rounds + 1;
    if (!haveA) {  // no merge anywhere: every epsilon is final
        std::swap(c->ie_tmp, P.ie);
        return 0;
    }
    // epsilon of the initiators at their merged position (the recursive epsv call, :169), then commit
    MergeState A = mstate(c->mA);
    EpsOp<true, true> opf{A, MergeState{}, nullptr, lrestr, dyn, ietmp, dchg};
    opf.tl = tl;
    int rc = launch_near(c, opf, dyn, tl);   // (the leaf boxes are those of the last round: the same A)
    if (!rc && multi) rc = gather_ie(ietmp);
    if (rc) return rc;
    std::swap(c->ie_tmp, P.ie);  // absorbed-before-turn particles kept their old value in ie_tmp (never written)
    CK(cudaMemsetAsync(dchg, 0, 2 * sizeof(int), st));
    k_merge_apply<<<cdiv(n, 256), 256, 0, st>>>(n, A, P.x.as<double>(), P.y.as<double>(), P.g.as<double>(), dchg); CKLAUNCH();
    u32 nm = 0;
    rc = read_u32(c, (u32*)dchg, &nm);
    if (rc) return rc;
    if (merged) *merged = (int)nm;
    return 0;
}

int vvgpu_convective(vvgpu_ctx* c, double inf_vx, double inf_vy, double dt, const double* sinks_xyg, size_t nsink) {
    if (!c || (nsink && !sinks_xyg)) return fail(c, VVGPU_EINVAL, "convective: bad argument");
    if (!c->built || !c->lists_ready) return fail(c, VVGPU_ESTATE, "tree is not built");
    CK(cudaSetDevice(c->device));
    PhaseTimer t(c, VVGPU_T_CONV);
    if (c->tn == 0) return 0;
    c->v_dirty = true;
    bool ok = true;
    double* ds = c->d_sinks.get<double>(3 * nsink, &ok);
    NEED(ok);
    if (nsink) CK(cudaMemcpyAsync(ds, sinks_xyg, 3 * nsink * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ConvOp op{inf_vx, inf_vy, dt * k1_Pi, c->taylor.as<double>(), ds, (int)nsink};
    int rc = launch_conv(c, op);
    if (rc) return rc;
    if (c->any_body_flow && c->nbody) {
        BodyFull B{c->nseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(), c->s_cy.as<double>(),
                   c->s_dlx.as<double>(), c->s_dly.as<double>(), c->s_g.as<double>(), c->s_ie.as<double>(),
                   c->s_slip.as<int>(), c->b_first.as<int>(), c->b_prop.as<double>()};
        k_body_influence<<<cdiv(c->tn, 128), 128, 0, c->stream>>>(c->tn, c->ps[c->cur].view(), B); CKLAUNCH();
    }
    if (nsink) CK(stream_sync(c));
    return 0;
}

int vvgpu_velocity_at(vvgpu_ctx* c, const double* xy, size_t npts, double inf_vx, double inf_vy, double dt,
                      const double* sinks_xyg, size_t nsink, double* vxy_out) {
    if (!c || (npts && (!xy || !vxy_out)) || (nsink && !sinks_xyg)) return fail(c, VVGPU_EINVAL, "velocity_at: bad argument");
    if (!c->built) return fail(c, VVGPU_ESTATE, "TTree::findNode(): tree is not built");   // TSortedTree.cpp:286-288
    if (npts == 0) return 0;
    CK(cudaSetDevice(c->device));
    bool ok = true;
    double* dxy = c->pt_xy.get<double>(2 * npts, &ok);
    double* dout = c->pt_out.get<double>(2 * npts, &ok);
    double* ds = c->d_sinks.get<double>(3 * nsink, &ok);
    int* derr = c->d_err.get<int>(4, &ok);
    NEED(ok);
    CK(cudaMemcpyAsync(dxy, xy, 2 * npts * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (nsink) CK(cudaMemcpyAsync(ds, sinks_xyg, 3 * nsink * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(derr, 0, 4 * sizeof(int), c->stream));
    PointArgs A;
    A.T = c->T(); A.P = c->ps[c->cur].view(); A.npts = (int)npts; A.xy = dxy; A.out = dout;
    A.farc = c->farc; A.inf_vx = inf_vx; A.inf_vy = inf_vy; A.eps2_div_srcg = dt * k1_Pi;
    A.sinks = ds; A.nsink = (int)nsink;
    A.B = BodyFull{c->nseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(), c->s_cy.as<double>(),
                   c->s_dlx.as<double>(), c->s_dly.as<double>(), c->s_g.as<double>(), c->s_ie.as<double>(),
                   c->s_slip.as<int>(), c->b_first.as<int>(), c->b_prop.as<double>()};
    A.body_flow = (c->any_body_flow && c->nbody) ? 1 : 0;
    A.err = derr;
    k_velocity_at<<<cdiv(npts, kPtWarps), kPtWarps * 32, 0, c->stream>>>(A); CKLAUNCH();
    CK(cudaMemcpyAsync(vxy_out, dout, 2 * npts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->h_pinned + 64, derr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    if (c->h_pinned[64] & 2) return fail(c, VVGPU_ELIMIT, "velocity_at: traversal stack overflow");
    return 0;
}

int vvgpu_eps2h_h2_at(vvgpu_ctx* c, const double* xy, size_t npts, double* eps2h_h2_out) {
    if (!c || (npts && (!xy || !eps2h_h2_out))) return fail(c, VVGPU_EINVAL, "eps2h_h2_at: bad argument");
    if (!c->built) return fail(c, VVGPU_ESTATE, "TTree::findNode(): tree is not built");
    if (npts == 0) return 0;
    CK(cudaSetDevice(c->device));
    bool ok = true;
    double* dxy = c->pt_xy.get<double>(2 * npts, &ok);
    double* dout = c->pt_out.get<double>(2 * npts, &ok);
    int* derr = c->d_err.get<int>(4, &ok);
    NEED(ok);
    CK(cudaMemcpyAsync(dxy, xy, 2 * npts * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(derr, 0, 4 * sizeof(int), c->stream));
    ScalarArgs A;
    A.T = c->T(); A.P = c->ps[c->cur].view(); A.npts = (int)npts; A.xy = dxy; A.out = dout; A.farc = c->farc;
    A.seg_perm = c->t_segperm[c->segcur].as<int>(); A.srx = c->s_rx.as<double>(); A.sry = c->s_ry.as<double>();
    A.err = derr;
    k_eps2h_h2_at<<<cdiv(npts, kPtWarps), kPtWarps * 32, 0, c->stream>>>(A); CKLAUNCH();
    CK(cudaMemcpyAsync(eps2h_h2_out, dout, 2 * npts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->h_pinned + 64, derr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    if (c->h_pinned[64] & 2) return fail(c, VVGPU_ELIMIT, "eps2h_h2_at: traversal stack overflow");
    return 0;
}

int vvgpu_node_influence(vvgpu_ctx* c, double* out_nseg) {
    if (!c || (c->nseg && !out_nseg)) return fail(c, VVGPU_EINVAL, "node_influence: bad argument");
    if (!c->built) return fail(c, VVGPU_ESTATE, "TTree::findNode(): tree is not built");
    if (c->nseg == 0) return 0;
    CK(cudaSetDevice(c->device));
    bool ok = true;
    double* dout = c->pt_out.get<double>(c->nseg, &ok);
    int* derr = c->d_err.get<int>(4, &ok);
    NEED(ok);
    CK(cudaMemsetAsync(derr, 0, 4 * sizeof(int), c->stream));
    SegInflArgs A{c->T(), c->ps[c->cur].view(), c->nseg, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(),
                  c->s_cy.as<double>(), c->s_dlx.as<double>(), c->s_dly.as<double>(), dout, c->farc, derr};
    k_node_influence<<<cdiv(c->nseg, kPtWarps), kPtWarps * 32, 0, c->stream>>>(A); CKLAUNCH();
    CK(cudaMemcpyAsync(out_nseg, dout, c->nseg * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->h_pinned + 64, derr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(stream_sync(c));
    if (c->h_pinned[64] & 2) return fail(c, VVGPU_ELIMIT, "node_influence: traversal stack overflow");
    return 0;
}

int vvgpu_vorticity_raster(vvgpu_ctx* c, float xmin, float ymin, float dxdy, int xres, int yres, double eps_mult, double dl,
                           double* out) {
    if (!c || xres <= 0 || yres <= 0 || !out) return fail(c, VVGPU_EINVAL, "vorticity_raster: bad argument");
    if (!(eps_mult > 0)) return fail(c, VVGPU_EINVAL, "XVorticity(): eps_mult must be positive");   // XVorticity.cpp:28-29
    if (c->built) return fail(c, VVGPU_ESTATE, "vorticity_raster builds its own tree: destroy the step's tree first");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->n, npts = (size_t)xres * yres;
    bool ok = true;
    // the reference works on a COPY of the Space (:33-38): the resident list must come back in its own order
    PSet& P = c->ps[c->cur];
    if (n && !c->ps_backup.ensure(n)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
    auto copy7 = [&](PSet& dst, PSet& src) {
        if (!n) return;
        Buf* d[6] = {&dst.x, &dst.y, &dst.g, &dst.vx, &dst.vy, &dst.ie};
        Buf* s7[6] = {&src.x, &src.y, &src.g, &src.vx, &src.vy, &src.ie};
        for (int k = 0; k < 6; k++) cudaMemcpyAsync(d[k]->p, s7[k]->p, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(dst.orig.p, src.orig.p, n * sizeof(int), cudaMemcpyDeviceToDevice, st);
    };
    copy7(c->ps_backup, P);
    int rc = tree_build_impl(c, 8, dl * 20, std::numeric_limits<double>::max(), 3u);   // TSortedTree tree(&S, 8, dl*20), :38
    if (!rc) {
        double* dxy = c->pt_xy.get<double>(2 * std::max<size_t>(n, 1), &ok);
        double* de2 = c->pt_out.get<double>(2 * std::max<size_t>(std::max(n, npts), 1), &ok);
        double* dv = c->pt_v.get<double>(2 * std::max<size_t>(n, 1), &ok);
        int* derr = c->d_err.get<int>(4, &ok);
        if (!ok) rc = fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
        if (!rc) {
            cudaMemsetAsync(derr, 0, 4 * sizeof(int), st);
            PSet& PP = c->ps[c->cur];
            if (n) {
                k_interleave_xy<<<cdiv(n, 256), 256, 0, st>>>((int)n, PP.view(), dxy);
                ScalarArgs SA;
                SA.T = c->T(); SA.P = PP.view(); SA.npts = (int)n; SA.xy = dxy; SA.out = de2; SA.farc = c->farc;
                SA.seg_perm = c->t_segperm[c->segcur].as<int>(); SA.srx = c->s_rx.as<double>(); SA.sry = c->s_ry.as<double>();
                SA.err = derr;
                k_eps2h_h2_at<<<cdiv(n, kPtWarps), kPtWarps * 32, 0, st>>>(SA);
                k_vort_prepare<<<cdiv(n, 256), 256, 0, st>>>((int)n, PP.view(), de2, eps_mult, dl, dv, dv + n);
                c->launches += 3;
            }
            RasterArgs RA;
            RA.T = c->T(); RA.P = PP.view(); RA.vx = dv; RA.vy = dv + n;
            RA.seg_perm = c->t_segperm[c->segcur].as<int>(); RA.srx = c->s_rx.as<double>(); RA.sry = c->s_ry.as<double>();
            RA.B = BodyGeom{c->nseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(), c->s_cy.as<double>(),
                            c->b_first.as<int>(), c->b_prop.as<double>()};
            RA.xmin = xmin; RA.ymin = ymin; RA.dxdy = dxdy; RA.xres = xres; RA.yres = yres;
            RA.eps_mult = eps_mult; RA.dl = dl; RA.farc = c->farc; RA.out = de2; RA.err = derr;
            k_vorticity_at<<<cdiv(npts, kPtWarps), kPtWarps * 32, 0, st>>>(RA); c->launches++;
            if (cudaGetLastError() != cudaSuccess) rc = fail(c, VVGPU_ECUDA, "vorticity_raster: launch failed");
            cudaMemcpyAsync(out, de2, npts * sizeof(double), cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(c->h_pinned + 64, derr, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(c, VVGPU_ECUDA, "vorticity_raster: kernel failed");
            if (!rc && (c->h_pinned[64] & 2)) rc = fail(c, VVGPU_ELIMIT, "vorticity_raster: traversal stack overflow");
        }
    }
    // destroy the raster's tree and put the resident list back
    c->built = false; c->lists_ready = false;
    c->nnodes = c->nleaves = c->ngroups = 0;
    copy7(c->ps[c->cur], c->ps_backup);
    if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = fail(c, VVGPU_ECUDA, "vorticity_raster: restore failed");
    return rc;
}

int vvgpu_pressure_raster(vvgpu_ctx* c, float xmin, float ymin, float dxdy, int xres, int yres, double dl, double re, double dt,
                          double inf_vx, double inf_vy, const double* sinks_xyg, size_t nsink, const double* gsum_nseg,
                          int use_ref_speed, double ref_vx, double ref_vy, double* out) {
    if (!c || xres <= 0 || yres <= 0 || !out || (c->nseg && !gsum_nseg) || (nsink && !sinks_xyg))
        return fail(c, VVGPU_EINVAL, "pressure_raster: bad argument");
    if (c->built) return fail(c, VVGPU_ESTATE, "pressure_raster builds its own tree: destroy the step's tree first");
    if (c->comm.nranks > 1) return fail(c, VVGPU_ESTATE, "pressure_raster: single-rank contexts only");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t n = c->n, npts = (size_t)xres * yres;
    bool ok = true;
    // the reference works on a COPY of the Space (:20-30): the resident list must come back as it was
    PSet& P0 = c->ps[c->cur];
    if (n && !c->ps_backup.ensure(n)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
    auto copy7 = [&](PSet& dst, PSet& src) {
        if (!n) return;
        Buf* d[6] = {&dst.x, &dst.y, &dst.g, &dst.vx, &dst.vy, &dst.ie};
        Buf* s7[6] = {&src.x, &src.y, &src.g, &src.vx, &src.vy, &src.ie};
        for (int k = 0; k < 6; k++) cudaMemcpyAsync(d[k]->p, s7[k]->p, n * sizeof(double), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(dst.orig.p, src.orig.p, n * sizeof(int), cudaMemcpyDeviceToDevice, st);
    };
    copy7(c->ps_backup, P0);
    // (process_all_lists ADDS to the v the records carry, MConvectiveFast.cpp:79: the caller's v is kept, as in the reference)
    // tree(&S, 8, dl*20, 0.1), eps.CalcEpsilonFast(false), process_all_lists, process_vort_list (:27, :61-64)
    int rc = vvgpu_tree_build(c, 8, dl * 20, 0.1, 3u);
    if (!rc) rc = vvgpu_epsilon(c, 0, nullptr);
    if (!rc) rc = vvgpu_convective(c, inf_vx, inf_vy, dt, sinks_xyg, nsink);
    if (!rc) rc = vvgpu_diffusive(c, re, nullptr);
    if (!rc) {
        double* dxy = c->pt_xy.get<double>(2 * npts, &ok);
        double* dvel = c->pt_out.get<double>(2 * npts, &ok);
        double* dacc = c->pt_v.get<double>(2 * npts, &ok);
        double* dres = dacc + npts;
        double* dgs = c->d_gsum.get<double>(std::max(c->nseg, 1), &ok);
        double* ds = c->d_sinks.get<double>(3 * nsink, &ok);
        int* derr = c->d_err.get<int>(4, &ok);
        if (!ok) rc = fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
        if (!rc) {
            PSet& PP = c->ps[c->cur];
            cudaMemsetAsync(derr, 0, 4 * sizeof(int), st);
            cudaMemsetAsync(dacc, 0, npts * sizeof(double), st);
            if (c->nseg) cudaMemcpyAsync(dgs, gsum_nseg, c->nseg * sizeof(double), cudaMemcpyHostToDevice, st);
            if (nsink) cudaMemcpyAsync(ds, sinks_xyg, 3 * nsink * sizeof(double), cudaMemcpyHostToDevice, st);
            k_raster_points<<<cdiv(npts, 256), 256, 0, st>>>(xmin, ymin, dxdy, xres, yres, dxy);
            BodyFull BF{c->nseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(), c->s_cy.as<double>(),
                        c->s_dlx.as<double>(), c->s_dly.as<double>(), c->s_g.as<double>(), c->s_ie.as<double>(),
                        c->s_slip.as<int>(), c->b_first.as<int>(), c->b_prop.as<double>()};
            PointArgs A;
            A.T = c->T(); A.P = PP.view(); A.npts = (int)npts; A.xy = dxy; A.out = dvel;
            A.farc = c->farc; A.inf_vx = inf_vx; A.inf_vy = inf_vy; A.eps2_div_srcg = dt * k1_Pi;
            A.sinks = ds; A.nsink = (int)nsink; A.B = BF;
            A.body_flow = (c->any_body_flow && c->nbody) ? 1 : 0;
            A.err = derr;
            k_velocity_at<<<cdiv(npts, kPtWarps), kPtWarps * 32, 0, st>>>(A);
            if (n) {
                dim3 grid(cdiv(npts, kPrPoints), cdiv(n, kPrChunk));
                k_pressure_vortices<<<grid, kPrPoints, 0, st>>>((int)n, PP.view(), (long long)npts, dxy, dacc);
            }
            PressureArgs R;
            R.B = BF;
            R.G = BodyGeom{c->nseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(), c->s_cy.as<double>(),
                           c->b_first.as<int>(), c->b_prop.as<double>()};
            R.gsum = dgs; R.npts = (long long)npts; R.xy = dxy; R.vel = dvel; R.acc = dacc;
            R.dt = dt; R.inf_vx = inf_vx; R.inf_vy = inf_vy; R.ref_vx = ref_vx; R.ref_vy = ref_vy; R.use_ref = use_ref_speed;
            R.out = dres;
            k_pressure_finish<<<cdiv(npts, 128), 128, 0, st>>>(R);
            c->launches += 4;
            if (cudaGetLastError() != cudaSuccess) rc = fail(c, VVGPU_ECUDA, "pressure_raster: launch failed");
            cudaMemcpyAsync(out, dres, npts * sizeof(double), cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(c->h_pinned + 64, derr, sizeof(int), cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(c, VVGPU_ECUDA, "pressure_raster: kernel failed");
            if (!rc && (c->h_pinned[64] & 2)) rc = fail(c, VVGPU_ELIMIT, "pressure_raster: traversal stack overflow");
        }
    }
    // destroy the raster's tree and put the resident list back
    c->built = false; c->lists_ready = false; c->v_dirty = false;
    c->nnodes = c->nleaves = c->ngroups = 0;
    copy7(c->ps[c->cur], c->ps_backup);
    if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = fail(c, VVGPU_ECUDA, "pressure_raster: restore failed");
    return rc;
}

int vvgpu_diffusive(vvgpu_ctx* c, double re, double* fric_out) {
    if (!c) return VVGPU_EINVAL;
    if (!c->built || !c->lists_ready) return fail(c, VVGPU_ESTATE, "tree is not built");
    CK(cudaSetDevice(c->device));
    {
        PhaseTimer t(c, VVGPU_T_DIFF);
        if (c->nseg) CK(cudaMemsetAsync(c->d_fric.p, 0, sizeof(double) * c->nseg, c->stream));
        if (c->tn) {
            bool ok = true;
            double* lb = c->lbox.get<double>(5 * (size_t)c->nleaves, &ok);
            NEED(ok);
            k_leaf_box<<<cdiv(c->nleaves, 128), 128, 0, c->stream>>>(c->Lv(), c->nleaves, c->ps[c->cur].view(), MergeState{}, lb); CKLAUNCH();
            DiffOp op{re, c->d_fric.as<double>()};
            int rc = launch_diff(c, op);
            if (rc) return rc;
            c->v_dirty = true;
        }
        // TAtt::fric gets a term from every nearby vortex (MDiffusiveFast.cpp:121-122): sum the ranks' shares
        if (c->nseg) { int rc = rank_sum_f64(c, c->d_fric.as<double>(), c->nseg); if (rc) return rc; }
    }
    if (fric_out && c->nseg) {
        CK(cudaMemcpyAsync(fric_out, c->d_fric.p, sizeof(double) * c->nseg, cudaMemcpyDeviceToHost, c->stream));
        CK(stream_sync(c));
    }
    return 0;
}

int vvgpu_move_and_clean(vvgpu_ctx* c, double dt_eff, double remove_eps, int remove_in_body, double* fdt_dead_xyo,
                         double* g_dead, double* gsum_delta, size_t* cleaned) {
    if (!c) return VVGPU_EINVAL;
    if (c->built) return fail(c, VVGPU_ESTATE, "move_and_clean while the tree is built (call tree_destroy first, vvflow.cpp:255-257)");
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int n = (int)c->n;
    bool ok = true;
    unsigned long long* dcl = c->d_cleaned.get<unsigned long long>(1, &ok);
    c->d_fdt.get<double>(3 * std::max(c->nbody, 1), &ok); c->d_gdead.get<double>(std::max(c->nbody, 1), &ok);
    c->d_gsum.get<double>(std::max(c->nseg, 1), &ok);
    u32* keep = c->flags.get<u32>(n + 2, &ok);
    u32* scan = c->scan_out.get<u32>(n + 2, &ok);
    PSet& P = c->ps[c->cur];
    PSet& Q = c->ps[c->cur ^ 1];
    if (!ok || !Q.ensure(n)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
    u32 nkeep = 0;
    {
        PhaseTimer t(c, VVGPU_T_MOVE);
        CK(cudaMemsetAsync(dcl, 0, sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(c->d_fdt.p, 0, sizeof(double) * 3 * std::max(c->nbody, 1), st));
        CK(cudaMemsetAsync(c->d_gdead.p, 0, sizeof(double) * std::max(c->nbody, 1), st));
        CK(cudaMemsetAsync(c->d_gsum.p, 0, sizeof(double) * std::max(c->nseg, 1), st));
        if (n) {
            BodyGeom B{c->nseg, c->nbody, c->s_rx.as<double>(), c->s_ry.as<double>(), c->s_cx.as<double>(), c->s_cy.as<double>(),
                       c->b_first.as<int>(), c->b_prop.as<double>()};
            k_move_flag<<<cdiv(n, 256), 256, 0, st>>>(n, P.view(), dt_eff, remove_eps, remove_in_body, B, keep, c->d_fdt.as<double>(),
                                                     c->d_gdead.as<double>(), c->d_gsum.as<double>(), dcl); CKLAUNCH();
            int rc = scan_flags(c, FlagArray{keep}, n, scan);
            if (rc) return rc;
            k_move_compact<<<cdiv(n, 256), 256, 0, st>>>(n, scan, P.view(), P.orig.as<int>(), Q.view(), Q.orig.as<int>()); CKLAUNCH();
            rc = read_u32(c, scan + n, &nkeep);
            if (rc) return rc;
            c->cur ^= 1;
            c->n = nkeep;
        }
    }
    if (fdt_dead_xyo && c->nbody) CK(cudaMemcpyAsync(fdt_dead_xyo, c->d_fdt.p, sizeof(double) * 3 * c->nbody, cudaMemcpyDeviceToHost, st));
    if (g_dead && c->nbody) CK(cudaMemcpyAsync(g_dead, c->d_gdead.p, sizeof(double) * c->nbody, cudaMemcpyDeviceToHost, st));
    if (gsum_delta && c->nseg) CK(cudaMemcpyAsync(gsum_delta, c->d_gsum.p, sizeof(double) * c->nseg, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(c->h_pinned + 16, dcl, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(stream_sync(c));
    if (cleaned) *cleaned = (size_t) * (unsigned long long*)(c->h_pinned + 16);
    return 0;
}

// ---- multi-GPU -----------------------------------------------------------------------------------------------
int vvgpu_comm_unique_id(void* id128, size_t cap) {
    if (!id128 || cap < sizeof(NcclApi::UniqueId)) return VVGPU_EINVAL;
    std::string err;
    NcclApi* api = NcclApi::get(&err);
    if (!api) return VVGPU_ECUDA;
    NcclApi::UniqueId id;
    if (api->GetUniqueId(&id)) return VVGPU_ECUDA;
    memcpy(id128, &id, sizeof(id));
    return 0;
}
static int comm_ready(vvgpu_ctx* c, int rank, int nranks) {
    if (!c || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks) return fail(c, VVGPU_EINVAL, "comm_init: bad rank / nranks");
    if (c->built) return fail(c, VVGPU_ESTATE, "comm_init while the tree is built");
    if (c->comm.kind != Comm::NONE) return fail(c, VVGPU_ESTATE, "comm_init: this context already has a communicator");
    return 0;
}
int vvgpu_comm_init(vvgpu_ctx* c, int rank, int nranks, const void* id128) {
    int rc = comm_ready(c, rank, nranks);
    if (rc) return rc;
    if (nranks == 1) return 0;
    if (!id128) return fail(c, VVGPU_EINVAL, "comm_init: no unique id");
    CK(cudaSetDevice(c->device));
    std::string err;
    NcclApi* api = NcclApi::get(&err);
    if (!api) return fail(c, VVGPU_ECUDA, err);
    NcclApi::UniqueId id;
    memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    const int nrc = api->CommInitRank(&comm, nranks, id, rank);
    if (nrc) return fail(c, VVGPU_ECUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(nrc));
    c->comm.kind = Comm::NCCL; c->comm.rank = rank; c->comm.nranks = nranks; c->comm.nccl = comm;
    return 0;
}
int vvgpu_group_create(const int* devices, int n, vvgpu_ctx** out) {
    if (!devices || !out || n < 1 || n > kMaxRanks) return VVGPU_EINVAL;
    for (int r = 0; r < n; r++) out[r] = nullptr;
    LocalGroup* g = new LocalGroup();
    g->n = n; g->refs = n; g->slot.resize(n);
    for (int r = 0; r < n; r++) {
        int rc = vvgpu_create(devices[r], &out[r]);
        if (rc) {
            for (int q = 0; q < r; q++) { out[q]->comm = Comm{}; vvgpu_destroy(out[q]); out[q] = nullptr; }
            delete g;
            return rc;
        }
        vvgpu_ctx* c = out[r];
        cudaEventCreateWithFlags(&g->slot[r].ready, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g->slot[r].done, cudaEventDisableTiming);
        if (n > 1) { c->comm.kind = Comm::LOCAL; c->comm.rank = r; c->comm.nranks = n; c->comm.grp = g; }
    }
    // direct NVLink copies between the group's devices where the hardware allows them
    for (int r = 0; r < n; r++)
        for (int q = 0; q < n; q++) {
            if (devices[r] == devices[q]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[r], devices[q]) == cudaSuccess && can) {
                cudaSetDevice(devices[r]);
                if (cudaDeviceEnablePeerAccess(devices[q], 0) != cudaSuccess) cudaGetLastError();   // already enabled
            }
        }
    if (n == 1) delete g;
    return 0;
}
int vvgpu_host_syncs(vvgpu_ctx* c, uint64_t* n) {
    if (!c || !n) return VVGPU_EINVAL;
    *n = c->host_syncs;
    c->host_syncs = 0;
    return 0;
}
int vvgpu_merge_rounds(vvgpu_ctx* c, int* rounds) {
    if (!c || !rounds) return VVGPU_EINVAL;
    *rounds = c->merge_rounds;
    return 0;
}
int vvgpu_comm_info(vvgpu_ctx* c, int* rank, int* nranks, int* kind) {
    if (!c) return VVGPU_EINVAL;
    if (rank) *rank = c->comm.rank;
    if (nranks) *nranks = c->comm.nranks;
    if (kind) *kind = (int)c->comm.kind;
    return 0;
}
// groups [first, last) x kShardBlock... : which rank owns leaf group g of ngroups (pure arithmetic, no device needed)
int vvgpu_shard_block(void) { return kShardBlock; }
int vvgpu_shard_owner(int group, int nranks) { return (nranks > 0 && group >= 0) ? (group / kShardBlock) % nranks : -1; }

// Upload of one SLICE per rank: rank r hands over records [first, first + count) of a list of n_total; the slices are
// gathered over the transport, so the full list crosses PCIe once in total instead of once per rank.
int vvgpu_set_particles_slice(vvgpu_ctx* c, int list, const vvgpu_obj* objs, size_t first, size_t count, size_t n_total) {
    if (!c || list != VVGPU_LIST_VORTEX || (!objs && count) || first + count > n_total) return fail(c, VVGPU_EINVAL, "set_particles_slice: bad argument");
    if (n_total > (size_t)std::numeric_limits<int>::max() / 4) return fail(c, VVGPU_EINVAL, "set_particles_slice: too many particles");
    if (c->built) return fail(c, VVGPU_ESTATE, "set_particles while the tree is built (the tree holds positions into the list)");
    const int P = c->comm.nranks;
    const size_t lo = n_total * c->comm.rank / P, hi = n_total * (c->comm.rank + 1) / P;
    if (first != lo || count != hi - lo) return fail(c, VVGPU_EINVAL, "set_particles_slice: rank r owns [n r / P, n (r + 1) / P)");
    CK(cudaSetDevice(c->device));
    const size_t per = (n_total + P - 1) / P;   // padded slice, in records
    bool ok = true;
    double* send = (double*)c->xsend.get<u64>(per * 6 + 8, &ok);
    double* recv = (P > 1) ? (double*)c->xrecv.get<u64>(per * 6 * P + 8, &ok) : send;
    if (!ok || !c->ps[c->cur].ensure(n_total)) return fail(c, VVGPU_ENOMEM, "cudaMalloc failed");
    if (count) CK(cudaMemcpyAsync(send, objs, count * 48, cudaMemcpyDefault, c->stream));
    if (P > 1) { int rc = comm_gather(c, send, recv, per * 48); if (rc) return rc; }
    c->n = n_total; c->orig_next = n_total;
    if (n_total) {
        k_unpack48_slices<<<cdiv(n_total, 256), 256, 0, c->stream>>>((int)n_total, P, (int)per, recv, c->ps[c->cur].view(), c->ps[c->cur].orig.as<int>());
        CKLAUNCH();
    }
    CK(stream_sync(c));  // the caller may reuse `objs`
    return 0;
}
int vvgpu_particle_arrays_dev(vvgpu_ctx* c, int list, double** arrays6, size_t* n) {
    if (!c || list != VVGPU_LIST_VORTEX || !arrays6) return VVGPU_EINVAL;
    Particles p = c->ps[c->cur].view();
    arrays6[0] = p.x; arrays6[1] = p.y; arrays6[2] = p.g; arrays6[3] = p.vx; arrays6[4] = p.vy; arrays6[5] = p.ie;
    if (n) *n = c->n;
    return 0;
}
// pending lazy exchanges (v of the other ranks' targets); collective, a no-op when nothing is pending
int vvgpu_sync_ranks(vvgpu_ctx* c) {
    if (!c) return VVGPU_EINVAL;
    if (!(c->built && c->v_dirty)) return 0;
    CK(cudaSetDevice(c->device));
    return sync_v(c);
}
int vvgpu_stream(vvgpu_ctx* c, void** s) {
    if (!c || !s) return VVGPU_EINVAL;
    *s = (void*)c->stream;
    return 0;
}
int vvgpu_synchronize(vvgpu_ctx* c) {
    if (!c) return VVGPU_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(stream_sync(c));
    return 0;
}

int vvgpu_phase_times(vvgpu_ctx* c, double* ms, uint64_t* launches) {
    if (!c) return VVGPU_EINVAL;
    CK(cudaSetDevice(c->device));
    CK(stream_sync(c));
    for (int k = 0; k < VVGPU_T_COUNT; k++) {
        float f = 0;
        if (c->ev_valid[k]) cudaEventElapsedTime(&f, c->ev0[k], c->ev1[k]);
        if (ms) ms[k] = f;
    }
    if (launches) *launches = c->launches;
    c->launches = 0;
    return 0;
}

int vvgpu_fp64_peak(vvgpu_ctx* c, double* tflops) {
    if (!c || !tflops) return VVGPU_EINVAL;
    CK(cudaSetDevice(c->device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    bool ok = true;
    double* out = c->stage.get<double>((size_t)blocks * threads, &ok);
    NEED(ok);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a, c->stream);
        k_fp64_peak<<<blocks, threads, 0, c->stream>>>(out, iters); c->launches++;
        cudaEventRecord(b, c->stream);
        CK(cudaEventSynchronize(b));
        float msf = 0;
        cudaEventElapsedTime(&msf, a, b);
        double fl = 2.0 * 8 * (double)iters * blocks * threads;
        if (rep) best = std::max(best, fl / (msf * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops = best;
    return 0;
}

}  // extern "C"
