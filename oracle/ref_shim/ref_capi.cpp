// TEST INFRASTRUCTURE — not part of the product path.
//
// C entry points around the reference's OWN classes (compiled unmodified from
// /root/reference, see oracle/Makefile) so that tests and bench.py's
// cpu_baseline / --impl reference legs can drive them through ctypes:
//   stree            libvvhd/headers/TSortedTree.hpp:60-92
//   MEpsilonFast     libvvhd/headers/MEpsilonFast.hpp:5-31
//   MConvectiveFast  libvvhd/headers/MConvectiveFast.hpp:8-27
//   MDiffusiveFast   libvvhd/headers/MDiffusiveFast.hpp:5-18
//   MFlowmove        libvvhd/headers/MFlowmove.hpp:5-19
// The step loop in vvr_step() follows utils/vvflow/vvflow.cpp:198-266 minus
// Stepdata, Space::save and sensors (which need HDF5 / files).
//
// `private` is opened up ONLY in this translation unit so the exporter can read
// stree::rootNode / bottomNodes and TBody::_cofm etc.; access specifiers do not
// change the Itanium-ABI layout, and the reference objects themselves are
// compiled without this define.

#define private public
#include "TSortedTree.hpp"
#include "TBody.hpp"
#undef private
#include "TSpace.hpp"
#include "MEpsilonFast.hpp"
#define private public   /* NodeInfluence (the vortex term of the SLAE right-hand side) is a private member */
#include "MConvectiveFast.hpp"
#undef private
#include "MDiffusiveFast.hpp"
#include "MFlowmove.hpp"
#include "XVorticity.hpp"
#include "XPressure.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

namespace {

struct Ctx {
    Space S;
    std::unique_ptr<stree> tree;
    std::unique_ptr<MEpsilonFast> eps;
    std::unique_ptr<MConvectiveFast> conv;
    std::unique_ptr<MDiffusiveFast> diff;
    std::unique_ptr<MFlowmove> flow;
    const void* collision = nullptr;
    int far_criteria = 8;
    double min_node = 0, max_node = std::numeric_limits<double>::max();
    // full leaf list saved while a sample is active (vvr_sample_leaves)
    std::vector<TSortedNode*> saved_leaves;
    // export caches
    std::vector<const snode*> pre;                       // nodes in DFS pre-order
    std::unordered_map<const snode*, int64_t> pre_id;
    std::unordered_map<const snode*, int64_t> leaf_id;
    std::string err;
};

std::vector<TObj>* list_of(Ctx* c, int list) {
    switch (list) {
        case 0: return &c->S.VortexList;
        case 1: return &c->S.HeatList;
        case 2: return &c->S.StreakList;
        case 3: return &c->S.SourceList;
        default: return nullptr;
    }
}

void make_modules(Ctx* c) {
    c->tree.reset(new stree(&c->S, c->far_criteria, c->min_node, c->max_node));
    c->eps.reset(new MEpsilonFast(&c->S, c->tree.get()));
    c->conv.reset(new MConvectiveFast(&c->S, c->tree.get()));
    c->diff.reset(new MDiffusiveFast(&c->S, c->tree.get()));
    c->flow.reset(new MFlowmove(&c->S));
}

void walk(Ctx* c, const snode* n) {
    c->pre_id[n] = (int64_t)c->pre.size();
    c->pre.push_back(n);
    if (n->ch1) { walk(c, n->ch1); walk(c, n->ch2); }
}

void index_tree(Ctx* c) {
    c->pre.clear(); c->pre_id.clear(); c->leaf_id.clear();
    if (!c->tree || !c->tree->rootNode) return;
    walk(c, c->tree->rootNode);
    const auto& leaves = c->saved_leaves.empty() ? c->tree->bottomNodes : c->saved_leaves;
    for (size_t i = 0; i < leaves.size(); i++) c->leaf_id[leaves[i]] = (int64_t)i;
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void* vvr_create() { Ctx* c = new Ctx(); return c; }
void vvr_destroy(void* h) { delete (Ctx*)h; }

// re, pr, dt (decimal seconds, as the Lua binding does: TTime::makeWithSecondsDecimal),
// constant U_inf. time is reset to 0.
void vvr_set_params(void* h, double re, double pr, double dt, double inf_vx, double inf_vy) {
    Ctx* c = (Ctx*)h;
    c->S.re = re; c->S.pr = pr;
    c->S.dt = TTime::makeWithSecondsDecimal(dt);
    c->S.time = TTime();
    char buf[64];
    snprintf(buf, sizeof buf, "%.17g", inf_vx); c->S.inf_vx = std::string(buf);
    snprintf(buf, sizeof buf, "%.17g", inf_vy); c->S.inf_vy = std::string(buf);
}
double vvr_time(void* h) { return double(((Ctx*)h)->S.time); }
double vvr_dt(void* h) { return double(((Ctx*)h)->S.dt); }

// 24-byte (x,y,g) records, the format Space::load_list_bin reads (TSpace.cpp:161-175)
void vvr_set_list(void* h, int list, const double* xyg, size_t n) {
    auto* li = list_of((Ctx*)h, list);
    li->clear(); li->reserve(n);
    for (size_t i = 0; i < n; i++) li->push_back(TObj(xyg[3 * i], xyg[3 * i + 1], xyg[3 * i + 2]));
}
// 48-byte TObj records (x y g vx vy _1_eps), TObj.hpp:10-16
void vvr_set_list48(void* h, int list, const double* rec, size_t n) {
    auto* li = list_of((Ctx*)h, list);
    li->resize(n);
    static_assert(sizeof(TObj) == 48, "TObj layout");
    if (n) memcpy((void*)li->data(), rec, n * sizeof(TObj));
}
size_t vvr_list_size(void* h, int list) { return list_of((Ctx*)h, list)->size(); }
void vvr_get_list48(void* h, int list, double* out) {
    auto* li = list_of((Ctx*)h, list);
    if (li->size()) memcpy(out, (const void*)li->data(), li->size() * sizeof(TObj));
}

// gen_cylinder{R,N} = gen_arc_N(c=0, R, 2pi -> 0, N)  (utils/vvflow/gen_cylinder.cpp:33-41,
// gen_body.cpp:74-88); returns the body index.
int vvr_add_cylinder(void* h, double cx, double cy, double R, size_t N) {
    Ctx* c = (Ctx*)h;
    std::shared_ptr<TBody> body = std::make_shared<TBody>();
    double a1 = 2 * M_PI, a2 = 0;
    for (size_t i = 0; i < N; i++) {
        double ai = a1 + (a2 - a1) * double(i) / double(N);
        TVec p = TVec(cx, cy) + TVec(R * cos(ai), R * sin(ai));
        body->alist.emplace_back(p, (uint32_t)0);
    }
    body->doUpdateSegments();
    body->doFillProperties();
    c->S.BodyList.push_back(body);
    c->S.EnumerateBodies();
    return (int)c->S.BodyList.size() - 1;
}
// arbitrary closed polygon given by its corners (x,y pairs) and per-corner slip flags (may be NULL)
int vvr_add_polygon(void* h, const double* xy, const uint32_t* slip, size_t N) {
    Ctx* c = (Ctx*)h;
    std::shared_ptr<TBody> body = std::make_shared<TBody>();
    for (size_t i = 0; i < N; i++) body->alist.emplace_back(TVec(xy[2 * i], xy[2 * i + 1]), slip ? slip[i] : 0u);
    body->doUpdateSegments();
    body->doFillProperties();
    c->S.BodyList.push_back(body);
    c->S.EnumerateBodies();
    return (int)c->S.BodyList.size() - 1;
}
// spring/holder setup for moving-body cases: kspring (x,y,o), damping(x,y,o), density,
// holder speed expressions must be constants
void vvr_body_set_dynamics(void* h, int b, const double* kspring3, const double* damping3, double density,
                           const double* speed3) {
    Ctx* c = (Ctx*)h;
    TBody& body = *c->S.BodyList.at(b);
    if (kspring3) body.kspring = TVec3D(kspring3[0], kspring3[1], kspring3[2]);
    if (damping3) body.damping = TVec3D(damping3[0], damping3[1], damping3[2]);
    body.density = density;
    if (speed3) {
        char buf[64];
        snprintf(buf, sizeof buf, "%.17g", speed3[0]); body.speed_x = std::string(buf);
        snprintf(buf, sizeof buf, "%.17g", speed3[1]); body.speed_y = std::string(buf);
        snprintf(buf, sizeof buf, "%.17g", speed3[2]); body.speed_o = std::string(buf);
    }
}
void vvr_body_set_speed_slae(void* h, int b, const double* v3) {
    TBody& body = *((Ctx*)h)->S.BodyList.at(b);
    body.speed_slae = TVec3D(v3[0], v3[1], v3[2]);
}
size_t vvr_n_bodies(void* h) { return ((Ctx*)h)->S.BodyList.size(); }
size_t vvr_n_segments(void* h) { return ((Ctx*)h)->S.total_segment_count(); }

// per segment 12 doubles: r.x r.y corner.x corner.y dl.x dl.y g gsum fric _1_eps slip body
void vvr_get_segments(void* h, double* out) {
    Ctx* c = (Ctx*)h;
    size_t k = 0; int b = 0;
    for (auto& lbody : c->S.BodyList) {
        for (auto& a : lbody->alist) {
            double* o = out + 12 * k++;
            o[0] = a.r.x; o[1] = a.r.y; o[2] = a.corner.x; o[3] = a.corner.y;
            o[4] = a.dl.x; o[5] = a.dl.y; o[6] = a.g; o[7] = a.gsum; o[8] = a.fric;
            o[9] = a._1_eps; o[10] = (double)a.slip; o[11] = (double)b;
        }
        b++;
    }
}
// write back g / gsum / fric of every segment (to inject a state)
void vvr_set_segments_ggf(void* h, const double* g, const double* gsum, const double* fric) {
    Ctx* c = (Ctx*)h; size_t k = 0;
    for (auto& lbody : c->S.BodyList)
        for (auto& a : lbody->alist) {
            if (g) a.g = g[k];
            if (gsum) a.gsum = gsum[k];
            if (fric) a.fric = fric[k];
            k++;
        }
}
// per body 32 doubles:
//  0 axis.x 1 axis.y 2 cofm.x 3 cofm.y 4 bl.x 5 bl.y 6 tr.x 7 tr.y 8 disc_r2 9 inside_valid
//  10-12 speed_slae  13-15 fdt_dead  16 g_dead  17-19 force_hydro  20-22 force_holder
//  23-25 friction  26 n_segments 27 slip 28 area
void vvr_get_body(void* h, int b, double* o) {
    TBody& B = *((Ctx*)h)->S.BodyList.at(b);
    memset(o, 0, 32 * sizeof(double));
    TVec ax = B.get_axis();
    o[0] = ax.x; o[1] = ax.y; o[2] = B._cofm.x; o[3] = B._cofm.y;
    o[4] = B._min_rect_bl.x; o[5] = B._min_rect_bl.y; o[6] = B._min_rect_tr.x; o[7] = B._min_rect_tr.y;
    o[8] = B._min_disc_r2; o[9] = B.isInsideValid() ? 1 : 0;
    o[10] = B.speed_slae.r.x; o[11] = B.speed_slae.r.y; o[12] = B.speed_slae.o;
    o[13] = B.fdt_dead.r.x; o[14] = B.fdt_dead.r.y; o[15] = B.fdt_dead.o; o[16] = B.g_dead;
    o[17] = B.force_hydro.r.x; o[18] = B.force_hydro.r.y; o[19] = B.force_hydro.o;
    o[20] = B.force_holder.r.x; o[21] = B.force_holder.r.y; o[22] = B.force_holder.o;
    o[23] = B.friction.r.x; o[24] = B.friction.r.y; o[25] = B.friction.o;
    o[26] = (double)B.size(); o[27] = B.get_slip() ? 1 : 0; o[28] = B.get_area();
}

// tree parameters exactly as vvflow.cpp:200-203 derives them
void vvr_tree_default_params(void* h, double* min_node, double* max_node) {
    Ctx* c = (Ctx*)h;
    double dl = c->S.average_segment_length();
    *min_node = dl > 0 ? dl * 5 : 0;
    *max_node = dl > 0 ? dl * 100 : std::numeric_limits<double>::max();
}
void vvr_tree_params(void* h, int far_criteria, double min_node, double max_node) {
    Ctx* c = (Ctx*)h;
    c->far_criteria = far_criteria; c->min_node = min_node; c->max_node = max_node;
    c->saved_leaves.clear();
    make_modules(c);
}
void vvr_tree_build(void* h, int v, int b, int hh) {
    Ctx* c = (Ctx*)h;
    if (!c->tree) make_modules(c);
    c->saved_leaves.clear();
    c->tree->build(v, b, hh);
}
void vvr_tree_destroy(void* h) {
    Ctx* c = (Ctx*)h;
    if (!c->tree) return;
    if (!c->saved_leaves.empty()) { c->tree->bottomNodes = c->saved_leaves; c->saved_leaves.clear(); }
    c->tree->destroy();
    c->pre.clear(); c->pre_id.clear(); c->leaf_id.clear();
}
// Keep only every `stride`-th leaf (offset `phase`) in bottomNodes, so that the reference's
// CalcEpsilonFast/process_all_lists/process_vort_list — unmodified — run on a bounded sample
// of leaves (bench.py's cpu_baseline). Returns the number of leaves kept.
size_t vvr_sample_leaves(void* h, size_t stride, size_t phase) {
    Ctx* c = (Ctx*)h;
    if (c->saved_leaves.empty()) c->saved_leaves = c->tree->bottomNodes;
    std::vector<TSortedNode*> keep;
    for (size_t i = phase; i < c->saved_leaves.size(); i += stride) keep.push_back(c->saved_leaves[i]);
    c->tree->bottomNodes = keep;
    return keep.size();
}
// near-pair count (targets with g!=0 x sources with g!=0 over near leaves) and far-node count of the
// CURRENT bottomNodes (sampled or full) — the "interactions" of SURVEY.md §8(d)
void vvr_count_interactions(void* h, double* near_pairs, double* far_nodes, double* n_targets) {
    Ctx* c = (Ctx*)h;
    double np = 0, nf = 0, nt = 0;
    std::unordered_map<const snode*, size_t> nz;
    auto count = [&](const snode* n) {
        auto it = nz.find(n);
        if (it != nz.end()) return it->second;
        size_t k = 0;
        for (TObj* o = n->vRange.first; o < n->vRange.last; o++) k += (o->g != 0);
        nz[n] = k; return k;
    };
    for (auto* leaf : c->tree->bottomNodes) {
        size_t t = count(leaf), s = 0;
        for (auto* nn : *leaf->NearNodes) s += count(nn);
        np += double(t) * double(s); nf += leaf->FarNodes->size(); nt += t;
    }
    *near_pairs = np; *far_nodes = nf; *n_targets = nt;
}

void vvr_tree_counts(void* h, size_t* n_nodes, size_t* n_leaves) {
    Ctx* c = (Ctx*)h;
    index_tree(c);
    *n_nodes = c->pre.size();
    *n_leaves = c->leaf_id.size();
}
// nodes in DFS pre-order (child 1 first).
//  dbl[10*i..]: x y h w CMp.x CMp.y CMp.g CMm.x CMm.y CMm.g
//  idx[10*i..]: vfirst vlast hfirst hlast sfirst slast nseg ch1 ch2 leaf_index   (-1 where n/a)
void vvr_tree_export(void* h, double* dbl, int64_t* idx) {
    Ctx* c = (Ctx*)h;
    if (c->pre.empty()) index_tree(c);
    const TObj* v0 = c->S.VortexList.data();
    const TObj* h0 = c->S.HeatList.data();
    const TObj* s0 = c->S.StreakList.data();
    for (size_t i = 0; i < c->pre.size(); i++) {
        const snode* n = c->pre[i];
        double* d = dbl + 10 * i; int64_t* k = idx + 10 * i;
        d[0] = n->x; d[1] = n->y; d[2] = n->h; d[3] = n->w;
        d[4] = n->CMp.r.x; d[5] = n->CMp.r.y; d[6] = n->CMp.g;
        d[7] = n->CMm.r.x; d[8] = n->CMm.r.y; d[9] = n->CMm.g;
        k[0] = n->vRange.first ? n->vRange.first - v0 : 0; k[1] = n->vRange.first ? n->vRange.last - v0 : 0;
        k[2] = n->hRange.first ? n->hRange.first - h0 : 0; k[3] = n->hRange.first ? n->hRange.last - h0 : 0;
        k[4] = n->sRange.first ? n->sRange.first - s0 : 0; k[5] = n->sRange.first ? n->sRange.last - s0 : 0;
        k[6] = (int64_t)n->bllist.size();
        k[7] = n->ch1 ? c->pre_id[n->ch1] : -1;
        k[8] = n->ch2 ? c->pre_id[n->ch2] : -1;
        auto it = c->leaf_id.find(n);
        k[9] = it == c->leaf_id.end() ? -1 : it->second;
    }
}
// CSR interaction lists of every leaf: near -> leaf indices, far -> pre-order node ids.
// Call with idx pointers NULL to get the totals in ptr[n_leaves].
void vvr_tree_lists(void* h, int64_t* near_ptr, int64_t* near_idx, int64_t* far_ptr, int64_t* far_idx) {
    Ctx* c = (Ctx*)h;
    if (c->pre.empty()) index_tree(c);
    const auto& leaves = c->saved_leaves.empty() ? c->tree->bottomNodes : c->saved_leaves;
    int64_t np = 0, fp = 0;
    for (size_t i = 0; i < leaves.size(); i++) {
        near_ptr[i] = np; far_ptr[i] = fp;
        for (auto* nn : *leaves[i]->NearNodes) { if (near_idx) near_idx[np] = c->leaf_id[nn]; np++; }
        for (auto* fn : *leaves[i]->FarNodes) { if (far_idx) far_idx[fp] = c->pre_id[fn]; fp++; }
    }
    near_ptr[leaves.size()] = np; far_ptr[leaves.size()] = fp;
}
// global segment indices held by each leaf's bllist (CSR)
void vvr_tree_leaf_segments(void* h, int64_t* ptr, int64_t* idx) {
    Ctx* c = (Ctx*)h;
    std::unordered_map<const TObj*, int64_t> seg_id;
    int64_t k = 0;
    for (auto& lbody : c->S.BodyList) for (auto& a : lbody->alist) seg_id[&a] = k++;
    const auto& leaves = c->saved_leaves.empty() ? c->tree->bottomNodes : c->saved_leaves;
    int64_t p = 0;
    for (size_t i = 0; i < leaves.size(); i++) {
        ptr[i] = p;
        for (auto* o : leaves[i]->bllist) { if (idx) idx[p] = seg_id[o]; p++; }
    }
    ptr[leaves.size()] = p;
}
int64_t vvr_find_node(void* h, double x, double y) {
    Ctx* c = (Ctx*)h;
    if (c->pre.empty()) index_tree(c);
    return c->pre_id[c->tree->findNode(TVec(x, y))];
}

int vvr_epsilon(void* h, int merge) { Ctx* c = (Ctx*)h; c->eps->CalcEpsilonFast(merge != 0); return c->eps->Merged(); }
void vvr_convective(void* h) { ((Ctx*)h)->conv->process_all_lists(); }
void vvr_velocity_at(void* h, const double* xy, size_t n, double* out) {
    Ctx* c = (Ctx*)h;
    for (size_t i = 0; i < n; i++) {
        TVec v = c->conv->velocity(TVec(xy[2 * i], xy[2 * i + 1]));
        out[2 * i] = v.x; out[2 * i + 1] = v.y;
    }
}
/* MEpsilonFast::eps2h / h2 (static, MEpsilonFast.cpp:66-107) with node = findNode(p): out = (eps2h, h2) pairs */
void vvr_eps2h_h2_at(void* h, const double* xy, size_t n, double* out) {
    Ctx* c = (Ctx*)h;
    for (size_t i = 0; i < n; i++) {
        TVec p(xy[2 * i], xy[2 * i + 1]);
        const TSortedNode* node = c->tree->findNode(p);
        out[2 * i] = MEpsilonFast::eps2h(*node, p);
        out[2 * i + 1] = MEpsilonFast::h2(*node, p);
    }
}
/* MConvectiveFast::NodeInfluence(*findNode(seg.r), seg) for every segment, in Space::BodyList order
 * (the vortex term of fillSlipEquationForSegment, MConvectiveFast.cpp:459-467) */
void vvr_node_influence(void* h, double* out) {
    Ctx* c = (Ctx*)h;
    size_t k = 0;
    for (auto& lbody : c->S.BodyList)
        for (auto& latt : lbody->alist)
            out[k++] = c->conv->NodeInfluence(*c->tree->findNode(latt.r), latt);
}
/* XVorticity(S, xmin, ymin, dxdy, xres, yres).evaluate() (libvvhd/src/XVorticity.cpp:26-97): the vorticity raster of
 * vvplot. It works on a COPY of the Space (sheds the attached vortices into it, builds its own tree with
 * minNodeSize = 20 dl), so the context's lists are untouched — except TAtt::gsum, which vortex_shed increments through
 * the shared TBody pointers; it is restored here. out[yj * xres + xi], float like XField::map. */
void vvr_vorticity_raster(void* h, float xmin, float ymin, float dxdy, int xres, int yres, double eps_mult, float* out) {
    Ctx* c = (Ctx*)h;
    std::vector<double> gsum;
    for (auto& lbody : c->S.BodyList) for (auto& latt : lbody->alist) gsum.push_back(latt.gsum);
    {
        XVorticity f(c->S, xmin, ymin, dxdy, xres, yres);
        f.eps_mult = eps_mult;
        f.evaluate();
        for (int yj = 0; yj < yres; yj++)
            for (int xi = 0; xi < xres; xi++) out[yj * xres + xi] = f.at(xi, yj);
    }
    size_t k = 0;
    for (auto& lbody : c->S.BodyList) for (auto& latt : lbody->alist) latt.gsum = gsum[k++];
}
/* XPressure(S, ...).evaluate() (libvvhd/src/XPressure.cpp:32-97), the pressure raster of vvplot; like XVorticity it works on a
 * copy of the Space that shares the TBody objects: vortex_shed increments TAtt::gsum and process_vort_list adds to
 * TAtt::fric through them; both are restored here. ref_frame: 's' 'o' 'f' 'b' (:38-52). */
void vvr_pressure_raster(void* h, float xmin, float ymin, float dxdy, int xres, int yres, int ref_frame, float* out) {
    Ctx* c = (Ctx*)h;
    std::vector<double> gsum, fric;
    for (auto& lbody : c->S.BodyList) for (auto& latt : lbody->alist) { gsum.push_back(latt.gsum); fric.push_back(latt.fric); }
    {
        XPressure f(c->S, xmin, ymin, dxdy, xres, yres);
        f.eps_mult = 1;
        f.ref_frame = (char)ref_frame;
        f.evaluate();
        for (int yj = 0; yj < yres; yj++)
            for (int xi = 0; xi < xres; xi++) out[yj * xres + xi] = f.at(xi, yj);
    }
    size_t k = 0;
    for (auto& lbody : c->S.BodyList) for (auto& latt : lbody->alist) { latt.gsum = gsum[k]; latt.fric = fric[k]; k++; }
}
void vvr_diffusive(void* h, int vort, int heat) {
    Ctx* c = (Ctx*)h;
    if (vort) c->diff->process_vort_list();
    if (heat) c->diff->process_heat_list();
}
size_t vvr_move_and_clean(void* h, int remove) {
    Ctx* c = (Ctx*)h; size_t cleaned = 0;
    c->flow->move_and_clean(remove != 0, &c->collision, &cleaned);
    return cleaned;
}
void vvr_calc_circulation(void* h) {
    Ctx* c = (Ctx*)h;
    if (c->collision != nullptr) c->conv->calc_circulation(&c->collision);
    c->conv->calc_circulation(&c->collision);
}
void vvr_vortex_shed(void* h) { ((Ctx*)h)->flow->vortex_shed(); }
void vvr_calc_forces(void* h) { ((Ctx*)h)->S.calc_forces(); }
void vvr_zero_forces(void* h) { ((Ctx*)h)->S.zero_forces(); }
void vvr_advance_time(void* h) { Ctx* c = (Ctx*)h; c->S.time = TTime::add(c->S.time, c->S.dt); }

// First half of one iteration of the loop at vvflow.cpp:214-244: body SLAE, shedding, force
// bookkeeping. After it the Space is exactly the state the hot path (vvflow.cpp:246-257) starts from.
void vvr_step_pre(void* h) {
    Ctx* c = (Ctx*)h;
    if (!c->tree) make_modules(c);
    if (c->S.BodyList.size()) {
        c->tree->build();
        vvr_calc_circulation(h);
        c->tree->destroy();
    }
    c->flow->heat_shed();
    c->flow->vortex_shed();
    c->flow->streak_shed();
    c->S.calc_forces();
    /* stepdata.write() would record force_hydro etc. here */
}
// ... callers read forces between vvr_step_pre and vvr_step_hot ...
// The hot path itself, vvflow.cpp:244-262. times[6]: build eps conv diff destroy move (seconds).
void vvr_step_hot(void* h, double* times) {
    Ctx* c = (Ctx*)h;
    bool is_viscous = (c->S.re != std::numeric_limits<double>::infinity());
    c->S.zero_forces();
    double t0 = now();
    c->tree->build();
    double t1 = now();
    c->eps->CalcEpsilonFast(is_viscous);
    double t2 = now();
    c->conv->process_all_lists();
    double t3 = now();
    if (is_viscous) { c->diff->process_vort_list(); c->diff->process_heat_list(); }
    double t4 = now();
    c->tree->destroy();
    double t5 = now();
    c->flow->move_and_clean(true, &c->collision);
    c->flow->heat_crop();
    double t6 = now();
    c->S.time = TTime::add(c->S.time, c->S.dt);
    if (times) { times[0] = t1 - t0; times[1] = t2 - t1; times[2] = t3 - t2; times[3] = t4 - t3; times[4] = t5 - t4; times[5] = t6 - t5; }
}

}  // extern "C"
