// vvgpu_adapter.hpp — the reference's own class surface for the per-step particle hot path,
// implemented on top of the vvgpu C ABI (include/vvgpu.h -> libvvgpu.so, CUDA sm_100a).
//
// Header-only; compiled INSIDE the reference tree against libvvhd's own headers
// (-I libvvhd/headers) and linked with libvvhd + libvvgpu. Nothing of the reference is copied:
// the classes below have the same names, constructor arguments, method names, defaults and
// error behaviour as
//      stree / TSortedTree   libvvhd/headers/TSortedTree.hpp:60-92
//      MEpsilonFast          libvvhd/headers/MEpsilonFast.hpp:5-31
//      MConvectiveFast       libvvhd/headers/MConvectiveFast.hpp:8-27   (process_all_lists only)
//      MDiffusiveFast        libvvhd/headers/MDiffusiveFast.hpp:5-18
//      MFlowmove             libvvhd/headers/MFlowmove.hpp:5-19
// but live in namespace vvgpu, because the unchanged CPU code (body SLAE, sensors, vvplot's X*
// evaluators) keeps using the reference's ::TSortedTree next to them. The step loop of
// utils/vvflow/vvflow.cpp:246-257 switches over by naming vvgpu:: types for its hot-path block —
// INTEGRATION.md shows the eight-line patch.
//
// Data contract: `Space` stays the owner of all state. build() uploads Space::VortexList
// (48-byte TObj records) and the body segments; the device permutes, merges, sets _1_eps and v;
// move_and_clean() advects/removes on the device and writes the surviving TObj records back
// into Space::VortexList in the reference's order, plus the per-body / per-segment increments
// (fdt_dead, g_dead, gsum, fric). Between build() and move_and_clean() the host copy of
// VortexList is stale unless sync_to_host() is called (e.g. for --sensors or a save).
//
// There is NO CPU fallback: if libvvgpu cannot create a context or a call fails, the adapter
// throws std::runtime_error with vvgpu_last_error().
#pragma once

#include "vvgpu.h"

#include "TSpace.hpp"
#include "TBody.hpp"
#include "TSortedTree.hpp"
#include "MFlowmove.hpp"

#include <cmath>
#include <cstdio>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace vvgpu {

static_assert(sizeof(TObj) == sizeof(vvgpu_obj), "TObj must be the 48-byte record of TObj.hpp:10-16");

// One device context per Space, shared by the five adapter objects of a step loop — or, with VVGPU_DEVICES="0,1,2,3",
// one context per listed device (SURVEY 8(e) from inside the vvflow process: vvgpu_group_create; entries may repeat,
// which puts several ranks on one device). Every rank holds the whole state and makes the same calls: `each` runs a
// call on all ranks, one host thread per rank (the library's exchanges meet there); replicated results are read from
// rank 0.
class Device {
    public:
        explicit Device(int device_index = 0): ctx(nullptr), index(device_index) {
            const char* env = getenv("VVGPU_DEVICES");
            std::vector<int> devs;
            if (env) for (const char* p = env; *p; ) { devs.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p == ',') p++; }
            if (devs.size() > 1) {
                all.resize(devs.size(), nullptr);
                int rc = vvgpu_group_create(devs.data(), (int)devs.size(), all.data());
                if (rc) throw std::runtime_error(std::string("vvgpu_group_create: ") + vvgpu_strerror(rc) +
                                                 " (libvvgpu has no CPU fallback; CUDA devices are required)");
                ctx = all[0];
                return;
            }
            int rc = vvgpu_create(devs.empty() ? device_index : devs[0], &ctx);
            if (rc) throw std::runtime_error(std::string("vvgpu_create: ") + vvgpu_strerror(rc) +
                                             " (libvvgpu has no CPU fallback; a CUDA device is required)");
            all.assign(1, ctx);
        }
        ~Device() { for (vvgpu_ctx* c: all) vvgpu_destroy(c); }
        Device(const Device&) = delete;
        Device& operator=(const Device&) = delete;

        void check(int rc, const char* what) const {
            if (rc) throw std::runtime_error(std::string(what) + ": " + vvgpu_strerror(rc) + ": " + vvgpu_last_error(ctx));
        }
        // f(ctx of rank r) on every rank; rank 0's return code is checked first, then the others'
        template <class F>
        void each(F f, const char* what) const {
            if (all.size() == 1) { check(f(ctx), what); return; }
            std::vector<int> rc(all.size(), 0);
            std::vector<std::thread> th;
            for (size_t r = 1; r < all.size(); r++) th.emplace_back([&, r] { rc[r] = f(all[r]); });
            rc[0] = f(all[0]);
            for (auto& t: th) t.join();
            for (size_t r = 0; r < all.size(); r++)
                if (rc[r]) throw std::runtime_error(std::string(what) + " (rank " + std::to_string(r) + "): " + vvgpu_strerror(rc[r]) +
                                                    ": " + vvgpu_last_error(all[r]));
        }
        static std::shared_ptr<Device>& of(Space* S) {
            // one Device per Space, created on first use (device from VVGPU_DEVICE, default 0; VVGPU_DEVICES for several)
            static std::vector<std::pair<Space*, std::shared_ptr<Device>>> table;
            for (auto& e: table) if (e.first == S) return e.second;
            const char* env = getenv("VVGPU_DEVICE");
            table.emplace_back(S, std::make_shared<Device>(env ? atoi(env) : 0));
            return table.back().second;
        }

        vvgpu_ctx* ctx;                 // rank 0
        std::vector<vvgpu_ctx*> all;    // every rank
        int index;
        bool dev_newer = false;   // the device holds a newer VortexList than the host
        bool resident = false;    // the list LIVES on the device: the host copy is only materialised on request
};

// SURVEY 8(f) row 2. With resident = true the vortex list stays on the device from step to step: move_and_clean() does
// not download it, vortex_shed() appends what the bodies shed (vvgpu_append_particles) and Space::VortexList stays
// empty until sync_to_host() materialises a snapshot (a save, TSpace.cpp:64-83). Host code that READS the list every
// step must not be in the loop then: the SLAE right-hand side comes from MConvectiveFast::NodeInfluence() on the device
// tree (tests/host/dropin_step.cpp `resident` shows the loop).
inline void set_resident(Space* S, bool on) { Device::of(S)->resident = on; }
inline size_t particle_count(Space* S) {
    Device& D = *Device::of(S);
    if (!D.dev_newer) return S->VortexList.size();
    size_t n = 0;
    D.check(vvgpu_particle_count(D.ctx, VVGPU_LIST_VORTEX, &n), "vvgpu_particle_count");
    return n;
}

// Space::gsum() of the resident list (the circulation equation of the SLAE reads it, MConvectiveFast.cpp:511)
inline double gsum(Space* S) {
    Device& D = *Device::of(S);
    if (!D.dev_newer) return S->gsum();
    double v = 0;
    D.check(vvgpu_particle_gsum(D.ctx, VVGPU_LIST_VORTEX, &v), "vvgpu_particle_gsum");
    return v;
}

// Space::VortexList <- device (reference order). Needed only if host code wants to look at the
// list between build() and move_and_clean().
inline void sync_to_host(Space* S) {
    Device& D = *Device::of(S);
    if (!D.dev_newer) return;
    size_t n = 0;
    D.each([](vvgpu_ctx* c) { return vvgpu_sync_ranks(c); }, "vvgpu_sync_ranks");
    D.check(vvgpu_particle_count(D.ctx, VVGPU_LIST_VORTEX, &n), "vvgpu_particle_count");
    S->VortexList.resize(n);
    D.check(vvgpu_get_particles(D.ctx, VVGPU_LIST_VORTEX, reinterpret_cast<vvgpu_obj*>(S->VortexList.data()), n, &n),
            "vvgpu_get_particles");
    if (!D.resident) D.dev_newer = false;   // (resident: the host copy is a read-only snapshot, the device stays the owner)
}

// the TAtt/TBody state the hot path reads; the private bounding rect / disc of TBody
// (TBody.cpp:328-343) are recomputed from the public corners with the same comparisons
inline void upload_bodies(Space* S) {
    Device& D = *Device::of(S);
    std::vector<vvgpu_seg> segs;
    std::vector<vvgpu_body> bodies;
    int ib = 0;
    for (auto& lbody: S->BodyList) {
        vvgpu_body B;
        TVec axis = lbody->get_axis(), cofm = lbody->get_cofm();
        B.axis_x = axis.x; B.axis_y = axis.y; B.cofm_x = cofm.x; B.cofm_y = cofm.y;
        double inf = std::numeric_limits<double>::infinity();
        B.bl_x = B.bl_y = inf; B.tr_x = B.tr_y = -inf; B.disc_r2 = 0;
        B.first_seg = (int32_t)segs.size(); B.n_seg = (int32_t)lbody->size(); B._pad = 0;
        for (auto& latt: lbody->alist) {
            double r2 = (latt.corner - cofm).abs2();
            if (r2 > B.disc_r2) B.disc_r2 = r2;
            if (latt.corner.x > B.tr_x) B.tr_x = latt.corner.x;
            if (latt.corner.y > B.tr_y) B.tr_y = latt.corner.y;
            if (latt.corner.x < B.bl_x) B.bl_x = latt.corner.x;
            if (latt.corner.y < B.bl_y) B.bl_y = latt.corner.y;
            vvgpu_seg s;
            s.rx = latt.r.x; s.ry = latt.r.y; s.cx = latt.corner.x; s.cy = latt.corner.y;
            s.dlx = latt.dl.x; s.dly = latt.dl.y; s.g = latt.g; s.ieps = latt._1_eps;
            s.slip = (int32_t)latt.slip; s.body = ib;
            segs.push_back(s);
        }
        B.speed_x = lbody->speed_slae.r.x; B.speed_y = lbody->speed_slae.r.y; B.speed_o = lbody->speed_slae.o;
        B.inside_valid = lbody->isInsideValid() ? 1 : 0;
        bodies.push_back(B);
        ib++;
    }
    D.each([&](vvgpu_ctx* c) { return vvgpu_set_bodies(c, segs.data(), segs.size(), bodies.data(), bodies.size()); }, "vvgpu_set_bodies");
}

// ------------------------------------------------------------------ stree, TSortedTree.hpp:60-92
class stree {
    public:
        stree(Space* sS, int sFarCriteria, double sMinNodeSize,
              double sMaxNodeSize = std::numeric_limits<double>::max()):
            S(sS), farCriteria(sFarCriteria), minNodeSize(sMinNodeSize), maxNodeSize(sMaxNodeSize), built(false) {}
        stree() = delete;
        stree(const stree&) = delete;
        stree& operator=(const stree&) = delete;

        // stree::build, TSortedTree.cpp:232-265
        void build(bool IncludeVortexes = true, bool IncludeBody = true, bool IncludeHeat = true) {
            if (built) { fprintf(stderr, "Tree is already built\n"); return; }   // :234
            if (IncludeHeat && (!S->HeatList.empty() || !S->StreakList.empty()))
                throw std::runtime_error("vvgpu::stree::build: heat / streak lists are not on the device path yet "
                                         "(SURVEY.md 8f row 3) and there is no CPU fallback");
            Device& D = *Device::of(S);
            if (!D.dev_newer)
                D.each([&](vvgpu_ctx* c) { return vvgpu_set_particles(c, VVGPU_LIST_VORTEX, reinterpret_cast<const vvgpu_obj*>(S->VortexList.data()),
                                                                      S->VortexList.size()); }, "vvgpu_set_particles");
            upload_bodies(S);
            unsigned mask = (IncludeVortexes ? 1u : 0u) | (IncludeBody ? 2u : 0u);
            D.each([&](vvgpu_ctx* c) { return vvgpu_tree_build(c, farCriteria, minNodeSize, maxNodeSize, mask); }, "vvgpu_tree_build");
            D.dev_newer = true;   // the list is permuted in place, like the reference's
            built = true;
        }
        void destroy() {   // :267-273
            Device& D = *Device::of(S);
            D.each([](vvgpu_ctx* c) { return vvgpu_tree_destroy(c); }, "vvgpu_tree_destroy");
            drop_mirror();
            built = false;
        }
        bool isBuilt() const { return built; }

        // stree::getBottomNodes / findNode (TSortedTree.hpp:86-87, TSortedTree.cpp:275-303) with the reference's own
        // signatures: the device tree is mirrored on the host, on first use after build(), as a tree of the reference's
        // ::TSortedNode objects — x y h w, CMp CMm, ch1 ch2, vRange into Space::VortexList (synchronised and in the
        // tree's order), bllist into the bodies' TAtt, NearNodes / FarNodes of every bottom node — so that unchanged
        // host code (X* evaluators, static MEpsilonFast::eps2h / h2, MConvectiveFast::NodeInfluence) can walk it.
        const std::vector< ::TSortedNode*>& getBottomNodes() const {
            if (!built) { fprintf(stderr, "PANIC in stree::getBottomNodes()! Tree isn't built\n"); return bottomNodes; }   // :277-281
            materialize();
            return bottomNodes;
        }
        const ::TSortedNode* findNode(TVec p) const {
            if (!built) throw std::invalid_argument("TTree::findNode(): tree is not built");   // :286-288
            materialize();
            const ::TSortedNode* Node = rootNode;
            while (Node->ch1) {   // :291-301
                if (Node->h < Node->w) Node = (p.x < Node->x) ? Node->ch1 : Node->ch2;
                else Node = (p.y < Node->y) ? Node->ch1 : Node->ch2;
            }
            return Node;
        }
        // the leaf table without the mirror: per leaf x y h w and the [first,last) range of its vortexes
        struct Leaf { double x, y, h, w; size_t vfirst, vlast, nseg; };
        std::vector<Leaf> leafTable() const {
            std::vector<Leaf> out;
            if (!built) { fprintf(stderr, "PANIC in stree::getBottomNodes()! Tree isn't built\n"); return out; }
            Device& D = *Device::of(S);
            size_t nn = 0, nl = 0, depth = 0;
            D.check(vvgpu_tree_counts(D.ctx, &nn, &nl, &depth), "vvgpu_tree_counts");
            std::vector<double> dbl(10 * nn);
            std::vector<int64_t> idx(8 * nn);
            D.check(vvgpu_tree_export(D.ctx, dbl.data(), idx.data(), nn), "vvgpu_tree_export");
            out.resize(nl);
            for (size_t i = 0; i < nn; i++) {
                int64_t li = idx[8 * i + 5];
                if (li < 0) continue;
                Leaf& L = out[(size_t)li];
                L.x = dbl[10 * i]; L.y = dbl[10 * i + 1]; L.h = dbl[10 * i + 2]; L.w = dbl[10 * i + 3];
                L.vfirst = (size_t)idx[8 * i]; L.vlast = (size_t)idx[8 * i + 1]; L.nseg = (size_t)idx[8 * i + 2];
            }
            return out;
        }
        ~stree() { drop_mirror(); }

    private:
        void drop_mirror() const {
            delete rootNode;   // snode::~snode deletes its children and its lists (TSortedTree.cpp:28-34)
            rootNode = nullptr;
            bottomNodes.clear();
        }
        void materialize() const {
            if (rootNode) return;
            Device& D = *Device::of(S);
            sync_to_host(S);   // vRange points into Space::VortexList, which has to be the tree's (permuted) list
            size_t nn = 0, nl = 0, depth = 0;
            D.check(vvgpu_tree_counts(D.ctx, &nn, &nl, &depth), "vvgpu_tree_counts");
            std::vector<double> dbl(10 * nn);
            std::vector<int64_t> idx(8 * nn);
            D.check(vvgpu_tree_export(D.ctx, dbl.data(), idx.data(), nn), "vvgpu_tree_export");
            std::vector< ::TSortedNode*> node(nn, nullptr);   // pre-order ids, children after their parent
            for (size_t i = 0; i < nn; i++) node[i] = new ::TSortedNode(nullptr);
            TObj* v0 = S->VortexList.data();
            bottomNodes.assign(nl, nullptr);
            for (size_t i = 0; i < nn; i++) {
                ::TSortedNode* n = node[i];
                n->x = dbl[10 * i]; n->y = dbl[10 * i + 1]; n->h = dbl[10 * i + 2]; n->w = dbl[10 * i + 3];
                n->CMp.r = TVec(dbl[10 * i + 4], dbl[10 * i + 5]); n->CMp.g = dbl[10 * i + 6];
                n->CMm.r = TVec(dbl[10 * i + 7], dbl[10 * i + 8]); n->CMm.g = dbl[10 * i + 9];
                n->i = (int)idx[8 * i + 7]; n->j = 0;
                if (v0) n->vRange.set(v0 + idx[8 * i], v0 + idx[8 * i + 1]);
                if (idx[8 * i + 3] >= 0) { n->ch1 = node[(size_t)idx[8 * i + 3]]; n->ch2 = node[(size_t)idx[8 * i + 4]]; }
                if (idx[8 * i + 5] >= 0) bottomNodes[(size_t)idx[8 * i + 5]] = n;
            }
            rootNode = node[0];
            // bllist of the bottom nodes: the segments in the order DistributeContent(LList&) left them (:139-148)
            std::vector<TObj*> att;
            for (auto& lbody: S->BodyList) for (auto& latt: lbody->alist) att.push_back(&latt);
            if (!att.empty()) {
                std::vector<int64_t> sp(nl + 1), si(att.size());
                D.check(vvgpu_tree_leaf_segments(D.ctx, sp.data(), si.data(), si.size()), "vvgpu_tree_leaf_segments");
                for (size_t l = 0; l < nl; l++)
                    for (int64_t k = sp[l]; k < sp[l + 1]; k++) bottomNodes[l]->bllist.push_back(att[(size_t)si[k]]);
            }
            // NearNodes / FarNodes exactly as snode::FindNearNodes builds them (:199-217)
            std::vector<int64_t> np(nl + 1), fp(nl + 1);
            D.check(vvgpu_tree_lists(D.ctx, np.data(), nullptr, 0, fp.data(), nullptr, 0), "vvgpu_tree_lists");
            std::vector<int64_t> ni((size_t)np[nl] + 1), fi((size_t)fp[nl] + 1);
            D.check(vvgpu_tree_lists(D.ctx, np.data(), ni.data(), ni.size(), fp.data(), fi.data(), fi.size()), "vvgpu_tree_lists");
            for (size_t l = 0; l < nl; l++) {
                ::TSortedNode* n = bottomNodes[l];
                n->NearNodes = new std::vector< ::TSortedNode*>();
                n->FarNodes = new std::vector< ::TSortedNode*>();
                for (int64_t k = np[l]; k < np[l + 1]; k++) n->NearNodes->push_back(bottomNodes[(size_t)ni[k]]);
                for (int64_t k = fp[l]; k < fp[l + 1]; k++) n->FarNodes->push_back(node[(size_t)fi[k]]);
            }
        }

        Space* S;
        int farCriteria;
        double minNodeSize;
        double maxNodeSize;
        bool built;
        mutable ::TSortedNode* rootNode = nullptr;             // host mirror (see getBottomNodes)
        mutable std::vector< ::TSortedNode*> bottomNodes;
};
typedef stree TSortedTree;

// ------------------------------------------------------- MEpsilonFast, MEpsilonFast.hpp:5-31
class MEpsilonFast {
    public:
        MEpsilonFast(Space* S, const TSortedTree* Tree): S(S), Tree(Tree), merged_(0) {}
        void CalcEpsilonFast(bool merge) {   // MEpsilonFast.cpp:11-63
            if (!Tree->isBuilt()) throw std::runtime_error("MEpsilonFast::CalcEpsilonFast: tree is not built");
            Device& D = *Device::of(S);
            std::vector<int> m(D.all.size(), 0);
            D.each([&](vvgpu_ctx* c) { size_t r = 0; while (D.all[r] != c) r++; return vvgpu_epsilon(c, merge ? 1 : 0, &m[r]); }, "vvgpu_epsilon");
            merged_ = m[0];
        }
        int Merged() { return merged_; }

    private:
        Space* S;
        const TSortedTree* Tree;
        int merged_;
};

// -------------------------------------------------- MConvectiveFast, MConvectiveFast.hpp:8-27
// Only process_all_lists (the per-step velocity pass). calc_circulation (the body SLAE) and
// velocity(p) stay with the reference's ::MConvectiveFast and its CPU tree.
class MConvectiveFast {
    public:
        MConvectiveFast() = delete;
        MConvectiveFast(Space* S, const TSortedTree* tree): S(S), tree(tree) {}
        MConvectiveFast(const MConvectiveFast&) = delete;
        MConvectiveFast& operator=(const MConvectiveFast&) = delete;

        void process_all_lists() {   // MConvectiveFast.cpp:36-114
            if (!tree->isBuilt()) throw std::runtime_error("MConvectiveFast::process_all_lists: tree is not built");
            Device& D = *Device::of(S);
            TVec inf = S->inf_speed();   // evaluated once per step on the host (TEval needs Lua)
            std::vector<double> sinks;
            for (auto& lobj: S->SourceList) { sinks.push_back(lobj.r.x); sinks.push_back(lobj.r.y); sinks.push_back(lobj.g); }
            D.each([&](vvgpu_ctx* c) { return vvgpu_convective(c, inf.x, inf.y, double(S->dt), sinks.data(), S->SourceList.size()); },
                   "vvgpu_convective");
        }

        // MConvectiveFast::NodeInfluence(*tree->findNode(seg.r), seg) (MConvectiveFast.cpp:398-418) for every segment
        // in BodyList order: the term fillSlipEquationForSegment (:459-467) subtracts from the right-hand side.
        // With it the SLAE stage runs on the device tree; INTEGRATION.md §2b shows the three-line patch.
        std::vector<double> NodeInfluence() const {
            if (!tree->isBuilt()) throw std::invalid_argument("TTree::findNode(): tree is not built");
            Device& D = *Device::of(S);
            std::vector<double> rhs(S->total_segment_count());
            D.check(vvgpu_node_influence(D.ctx, rhs.empty() ? nullptr : rhs.data()), "vvgpu_node_influence");
            return rhs;
        }

        TVec velocity(TVec p) const {   // MConvectiveFast.cpp:20-34 (sensors, X* rasters)
            TVec v = TVec(0, 0);
            velocity(&p, 1, &v);
            return v;
        }
        // batched form for the raster evaluators: n points in, n velocities out, one device call
        void velocity(const TVec* p, size_t n, TVec* out) const {
            if (!tree->isBuilt()) throw std::invalid_argument("TTree::findNode(): tree is not built");   // TSortedTree.cpp:286-288
            Device& D = *Device::of(S);
            TVec inf = S->inf_speed();
            std::vector<double> sinks, xy(2 * n), v(2 * n);
            for (auto& lobj: S->SourceList) { sinks.push_back(lobj.r.x); sinks.push_back(lobj.r.y); sinks.push_back(lobj.g); }
            for (size_t i = 0; i < n; i++) { xy[2 * i] = p[i].x; xy[2 * i + 1] = p[i].y; }
            D.check(vvgpu_velocity_at(D.ctx, xy.data(), n, inf.x, inf.y, double(S->dt), sinks.data(), S->SourceList.size(), v.data()),
                    "vvgpu_velocity_at");
            for (size_t i = 0; i < n; i++) out[i] = TVec(v[2 * i], v[2 * i + 1]);
        }

    private:
        Space* S;
        const TSortedTree* tree;
};

// ---------------------------------------------------- MDiffusiveFast, MDiffusiveFast.hpp:5-18
class MDiffusiveFast {
    public:
        MDiffusiveFast(Space* S, const TSortedTree* tree): S(S), tree(tree) {}
        void process_vort_list() {   // MDiffusiveFast.cpp:8-48
            if (!tree->isBuilt()) throw std::runtime_error("MDiffusiveFast::process_vort_list: tree is not built");
            Device& D = *Device::of(S);
            std::vector<double> fric(S->total_segment_count());   // summed over the ranks inside the library: rank 0's copy
            D.each([&](vvgpu_ctx* c) { return vvgpu_diffusive(c, S->re, (c == D.ctx && !fric.empty()) ? fric.data() : nullptr); }, "vvgpu_diffusive");
            size_t k = 0;   // TAtt::fric += ..., MDiffusiveFast.cpp:121-122
            for (auto& lbody: S->BodyList) for (auto& latt: lbody->alist) latt.fric += fric[k++];
        }
        void process_heat_list() {   // MDiffusiveFast.cpp:50-89
            if (!S->HeatList.empty())
                throw std::runtime_error("vvgpu::MDiffusiveFast::process_heat_list: heat particles are not on the device path yet");
        }

    private:
        Space* S;
        const TSortedTree* tree;
};

// ------------------------------------------------------------ MFlowmove, MFlowmove.hpp:5-19
class MFlowmove {
    public:
        MFlowmove(Space* S, double remove_eps = 1E-10): S(S), remove_eps(remove_eps), host(S, remove_eps) {}

        // MFlowmove::move_and_clean, MFlowmove.cpp:11-215. The vortex-particle part (:107-109,
        // :113-117, :124-144, :196) runs on the device. Everything about bodies, heat and streak
        // particles (collision dt :25-57, body motion :61-105, attached-vortex sums :201-214) is the
        // reference's own code, called with the vortex list parked aside.
        void move_and_clean(bool remove, const void** collision, size_t* cleaned_v = NULL) {
            if (collision == nullptr)
                throw std::invalid_argument("MFlowmove::move_and_clean(): invalid collision pointer");   // :20-22
            Device& D = *Device::of(S);
            if (!D.dev_newer)
                D.each([&](vvgpu_ctx* c) { return vvgpu_set_particles(c, VVGPU_LIST_VORTEX, reinterpret_cast<const vvgpu_obj*>(S->VortexList.data()),
                                                                      S->VortexList.size()); }, "vvgpu_set_particles");
            const double current_dt = collision_dt();
            std::vector<TObj> parked;
            parked.swap(S->VortexList);
            host.move_and_clean(remove, collision, nullptr);   // bodies move; collision is reported
            parked.swap(S->VortexList);
            upload_bodies(S);   // isPointInvalid tests against the MOVED bodies (:105 precedes :124)

            const size_t nb = S->BodyList.size(), ns = S->total_segment_count();
            std::vector<double> fdt(3 * nb + 1), gdead(nb + 1), gsum(ns + 1);
            size_t cleaned = 0;   // the move is replicated: every rank removes the same particles; rank 0 reports
            D.each([&](vvgpu_ctx* c) {
                if (c == D.ctx) return vvgpu_move_and_clean(c, current_dt, remove_eps, remove ? 1 : 0, fdt.data(), gdead.data(), gsum.data(), &cleaned);
                return vvgpu_move_and_clean(c, current_dt, remove_eps, remove ? 1 : 0, nullptr, nullptr, nullptr, nullptr);
            }, "vvgpu_move_and_clean");
            size_t ib = 0, k = 0;
            for (auto& lbody: S->BodyList) {
                lbody->fdt_dead.r.x += fdt[3 * ib]; lbody->fdt_dead.r.y += fdt[3 * ib + 1]; lbody->fdt_dead.o += fdt[3 * ib + 2];
                lbody->g_dead += gdead[ib];
                for (auto& latt: lbody->alist) latt.gsum += gsum[k++];
                ib++;
            }
            if (cleaned_v) *cleaned_v = cleaned;
            D.dev_newer = true;
            if (D.resident) S->VortexList.clear();   // the device owns the list; nothing crosses PCIe
            else sync_to_host(S);                    // shedding and the SLAE phase of the next step work on the host list
        }
        // MFlowmove::vortex_shed, MFlowmove.cpp:217-235. Resident list: the host vector is empty, so what the reference's
        // code pushes is exactly what is new; it goes behind the device list and the host vector is emptied again.
        void vortex_shed() {
            Device& D = *Device::of(S);
            if (!(D.resident && D.dev_newer)) { host.vortex_shed(); return; }
            S->VortexList.clear();
            host.vortex_shed();
            D.each([&](vvgpu_ctx* c) { return vvgpu_append_particles(c, VVGPU_LIST_VORTEX, reinterpret_cast<const vvgpu_obj*>(S->VortexList.data()),
                                                                     S->VortexList.size()); }, "vvgpu_append_particles");
            S->VortexList.clear();
        }
        void streak_shed() { host.streak_shed(); }
        void heat_shed() { host.heat_shed(); }
        void heat_crop(double scale = 16) { host.heat_crop(scale); }

    private:
        // the collision-shortened time step of MFlowmove.cpp:16-57 (same tests, same order)
        double collision_dt() const {
            double current_dt = S->dt;
            for (std::shared_ptr<TBody>& lbody: S->BodyList) {
                TVec3D cur = lbody->holder + lbody->dpos;
                TVec3D nxt = cur + double(S->dt) * lbody->speed_slae;
                const double cur_pos[3] = {cur.r.x, cur.r.y, cur.o}, new_pos[3] = {nxt.r.x, nxt.r.y, nxt.o};
                const double speed[3] = {lbody->speed_slae.r.x, lbody->speed_slae.r.y, lbody->speed_slae.o};
                const double mx[3] = {lbody->collision_max.r.x, lbody->collision_max.r.y, lbody->collision_max.o};
                const double mn[3] = {lbody->collision_min.r.x, lbody->collision_min.r.y, lbody->collision_min.o};
                const double ks[3] = {lbody->kspring.r.x, lbody->kspring.r.y, lbody->kspring.o};
                for (int i = 0; i < 3; i++) {
                    double dt_ = std::numeric_limits<double>::quiet_NaN();
                    if (TBody::isrigid(ks[i])) { /* do nothing */ }
                    else if (new_pos[i] > mx[i]) dt_ = (mx[i] - cur_pos[i]) / speed[i];
                    else if (new_pos[i] < mn[i]) dt_ = (mn[i] - cur_pos[i]) / speed[i];
                    if (dt_ < current_dt) current_dt = dt_;
                }
            }
            return current_dt;
        }

        Space* S;
        double remove_eps;
        ::MFlowmove host;
};

}  // namespace vvgpu
