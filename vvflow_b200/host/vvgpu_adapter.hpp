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
#include "MFlowmove.hpp"

#include <cmath>
#include <cstdio>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace vvgpu {

static_assert(sizeof(TObj) == sizeof(vvgpu_obj), "TObj must be the 48-byte record of TObj.hpp:10-16");

// One device context per Space, shared by the five adapter objects of a step loop.
class Device {
    public:
        explicit Device(int device_index = 0): ctx(nullptr), index(device_index) {
            int rc = vvgpu_create(device_index, &ctx);
            if (rc) throw std::runtime_error(std::string("vvgpu_create: ") + vvgpu_strerror(rc) +
                                             " (libvvgpu has no CPU fallback; a CUDA device is required)");
        }
        ~Device() { vvgpu_destroy(ctx); }
        Device(const Device&) = delete;
        Device& operator=(const Device&) = delete;

        void check(int rc, const char* what) const {
            if (rc) throw std::runtime_error(std::string(what) + ": " + vvgpu_strerror(rc) + ": " + vvgpu_last_error(ctx));
        }
        static std::shared_ptr<Device>& of(Space* S) {
            // one context per Space, created on first use (device from VVGPU_DEVICE, default 0)
            static std::vector<std::pair<Space*, std::shared_ptr<Device>>> table;
            for (auto& e: table) if (e.first == S) return e.second;
            const char* env = getenv("VVGPU_DEVICE");
            table.emplace_back(S, std::make_shared<Device>(env ? atoi(env) : 0));
            return table.back().second;
        }

        vvgpu_ctx* ctx;
        int index;
        bool dev_newer = false;   // the device holds a newer VortexList than the host
};

// Space::VortexList <- device (reference order). Needed only if host code wants to look at the
// list between build() and move_and_clean().
inline void sync_to_host(Space* S) {
    Device& D = *Device::of(S);
    if (!D.dev_newer) return;
    size_t n = 0;
    D.check(vvgpu_particle_count(D.ctx, VVGPU_LIST_VORTEX, &n), "vvgpu_particle_count");
    S->VortexList.resize(n);
    D.check(vvgpu_get_particles(D.ctx, VVGPU_LIST_VORTEX, reinterpret_cast<vvgpu_obj*>(S->VortexList.data()), n, &n),
            "vvgpu_get_particles");
    D.dev_newer = false;
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
    D.check(vvgpu_set_bodies(D.ctx, segs.data(), segs.size(), bodies.data(), bodies.size()), "vvgpu_set_bodies");
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
                D.check(vvgpu_set_particles(D.ctx, VVGPU_LIST_VORTEX,
                                            reinterpret_cast<const vvgpu_obj*>(S->VortexList.data()), S->VortexList.size()),
                        "vvgpu_set_particles");
            upload_bodies(S);
            unsigned mask = (IncludeVortexes ? 1u : 0u) | (IncludeBody ? 2u : 0u);
            D.check(vvgpu_tree_build(D.ctx, farCriteria, minNodeSize, maxNodeSize, mask), "vvgpu_tree_build");
            D.dev_newer = true;   // the list is permuted in place, like the reference's
            built = true;
        }
        void destroy() {   // :267-273
            Device& D = *Device::of(S);
            D.check(vvgpu_tree_destroy(D.ctx), "vvgpu_tree_destroy");
            built = false;
        }
        bool isBuilt() const { return built; }

        // leaf table in bottomNodes order (getBottomNodes, :275-282): per leaf x y h w and the
        // [first,last) range of its vortexes in the permuted list
        struct Leaf { double x, y, h, w; size_t vfirst, vlast, nseg; };
        std::vector<Leaf> getBottomNodes() const {
            std::vector<Leaf> out;
            if (!built) { fprintf(stderr, "PANIC in stree::getBottomNodes()! Tree isn't built\n"); return out; }   // :277-281
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

    private:
        Space* S;
        int farCriteria;
        double minNodeSize;
        double maxNodeSize;
        bool built;
};
typedef stree TSortedTree;

// ------------------------------------------------------- MEpsilonFast, MEpsilonFast.hpp:5-31
class MEpsilonFast {
    public:
        MEpsilonFast(Space* S, const TSortedTree* Tree): S(S), Tree(Tree), merged_(0) {}
        void CalcEpsilonFast(bool merge) {   // MEpsilonFast.cpp:11-63
            if (!Tree->isBuilt()) throw std::runtime_error("MEpsilonFast::CalcEpsilonFast: tree is not built");
            Device& D = *Device::of(S);
            D.check(vvgpu_epsilon(D.ctx, merge ? 1 : 0, &merged_), "vvgpu_epsilon");
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
            D.check(vvgpu_convective(D.ctx, inf.x, inf.y, double(S->dt), sinks.data(), S->SourceList.size()),
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
            std::vector<double> fric(S->total_segment_count());
            D.check(vvgpu_diffusive(D.ctx, S->re, fric.empty() ? nullptr : fric.data()), "vvgpu_diffusive");
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
                D.check(vvgpu_set_particles(D.ctx, VVGPU_LIST_VORTEX,
                                            reinterpret_cast<const vvgpu_obj*>(S->VortexList.data()), S->VortexList.size()),
                        "vvgpu_set_particles");
            const double current_dt = collision_dt();
            std::vector<TObj> parked;
            parked.swap(S->VortexList);
            host.move_and_clean(remove, collision, nullptr);   // bodies move; collision is reported
            parked.swap(S->VortexList);
            upload_bodies(S);   // isPointInvalid tests against the MOVED bodies (:105 precedes :124)

            const size_t nb = S->BodyList.size(), ns = S->total_segment_count();
            std::vector<double> fdt(3 * nb + 1), gdead(nb + 1), gsum(ns + 1);
            size_t cleaned = 0;
            D.check(vvgpu_move_and_clean(D.ctx, current_dt, remove_eps, remove ? 1 : 0, fdt.data(), gdead.data(),
                                         gsum.data(), &cleaned), "vvgpu_move_and_clean");
            size_t ib = 0, k = 0;
            for (auto& lbody: S->BodyList) {
                lbody->fdt_dead.r.x += fdt[3 * ib]; lbody->fdt_dead.r.y += fdt[3 * ib + 1]; lbody->fdt_dead.o += fdt[3 * ib + 2];
                lbody->g_dead += gdead[ib];
                for (auto& latt: lbody->alist) latt.gsum += gsum[k++];
                ib++;
            }
            if (cleaned_v) *cleaned_v = cleaned;
            D.dev_newer = true;
            sync_to_host(S);   // shedding and the SLAE phase of the next step work on the host list
        }
        void vortex_shed() { host.vortex_shed(); }
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
