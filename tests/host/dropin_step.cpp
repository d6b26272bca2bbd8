// TEST INFRASTRUCTURE — drop-in check of vvflow_b200/host/vvgpu_adapter.hpp.
//
// Replays the step loop of utils/vvflow/vvflow.cpp:198-266 for example/cyl_re600.lua
// (Re=600, dt=0.05, cylinder R=0.5 N=350, U_inf=(1,0)) with the reference's own classes for
// everything EXCEPT the hot-path block (:246-257), which is written twice:
//   arm "ref":  ::TSortedTree  ::MEpsilonFast  ::MConvectiveFast  ::MDiffusiveFast  ::MFlowmove
//   arm "gpu":  vvgpu::TSortedTree vvgpu::MEpsilonFast ... (libvvgpu.so, CUDA)
// Modes:
//   free N      run the gpu arm alone for N steps and print time + force_hydro (x y o) + particle
//               count per step, exactly what stepdata would record (README.md:115-120 rows)
//   points N    N lock-step steps, then on the next step's tree (after epsilon) compare the adapter's
//               velocity(p) on a raster and NodeInfluence() of every segment with the reference classes
//   lockstep N  per step: clone the reference Space into the gpu Space, run the hot path on both,
//               compare order / merge decisions bit-exactly and every floating output to 1e-10
// Built by oracle/Makefile into oracle/_ref/ (needs the reference headers; nothing is copied),
// linked against oracle/_ref/libvvref.so (reference objects) and vvflow_b200/lib/libvvgpu.so.
#include "vvgpu_adapter.hpp"

#include "TSortedTree.hpp"
#include "MEpsilonFast.hpp"
#define private public   /* `points` mode calls the reference's private NodeInfluence (test infrastructure only) */
#include "MConvectiveFast.hpp"
#undef private
#include "MDiffusiveFast.hpp"
#include "MFlowmove.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

static void make_case(Space& S, double re, double dt, double R, size_t N) {
    S.re = re;
    S.dt = TTime::makeWithSecondsDecimal(dt);
    S.time = TTime();
    S.inf_vx = std::string("1");
    std::shared_ptr<TBody> body = std::make_shared<TBody>();
    double a1 = 2 * M_PI, a2 = 0;   // gen_cylinder -> gen_arc_N, utils/vvflow/gen_body.cpp:74-88
    for (size_t i = 0; i < N; i++) {
        double ai = a1 + (a2 - a1) * double(i) / double(N);
        body->alist.emplace_back(TVec(R * cos(ai), R * sin(ai)), (uint32_t)0);
    }
    body->doUpdateSegments();
    body->doFillProperties();
    S.BodyList.push_back(body);
    S.EnumerateBodies();
}

struct RefArm {
    Space& S;
    TSortedTree tr;
    MConvectiveFast convective;
    MEpsilonFast epsilon;
    MDiffusiveFast diffusive;
    MFlowmove flowmove;
    const void* collision = nullptr;
    RefArm(Space& S, double mn, double mx): S(S), tr(&S, 8, mn, mx), convective(&S, &tr), epsilon(&S, &tr),
        diffusive(&S, &tr), flowmove(&S) {}
    // vvflow.cpp:216-244 (no stepdata / save)
    void pre() {
        if (S.BodyList.size()) {
            tr.build();
            if (collision != nullptr) convective.calc_circulation(&collision);
            convective.calc_circulation(&collision);
            tr.destroy();
        }
        flowmove.heat_shed();
        flowmove.vortex_shed();
        flowmove.streak_shed();
        S.calc_forces();
    }
    int hot(size_t* cleaned) {   // vvflow.cpp:246-257
        bool is_viscous = (S.re != std::numeric_limits<double>::infinity());
        tr.build();
        epsilon.CalcEpsilonFast(is_viscous);
        convective.process_all_lists();
        if (is_viscous) { diffusive.process_vort_list(); diffusive.process_heat_list(); }
        tr.destroy();
        flowmove.move_and_clean(true, &collision, cleaned);
        flowmove.heat_crop();
        return epsilon.Merged();
    }
};

struct GpuArm {
    Space& S;
    vvgpu::TSortedTree tr;
    vvgpu::MConvectiveFast convective;
    vvgpu::MEpsilonFast epsilon;
    vvgpu::MDiffusiveFast diffusive;
    vvgpu::MFlowmove flowmove;
    const void* collision = nullptr;
    GpuArm(Space& S, double mn, double mx): S(S), tr(&S, 8, mn, mx), convective(&S, &tr), epsilon(&S, &tr),
        diffusive(&S, &tr), flowmove(&S) {}
    int hot(size_t* cleaned) {   // the same eight lines with vvgpu:: types
        bool is_viscous = (S.re != std::numeric_limits<double>::infinity());
        tr.build();
        epsilon.CalcEpsilonFast(is_viscous);
        convective.process_all_lists();
        if (is_viscous) { diffusive.process_vort_list(); diffusive.process_heat_list(); }
        tr.destroy();
        flowmove.move_and_clean(true, &collision, cleaned);
        flowmove.heat_crop();
        return epsilon.Merged();
    }
};

static double relerr(double a, double b, double scale) { return fabs(a - b) / (scale > 0 ? scale : 1); }

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "free";
    int nsteps = argc > 2 ? atoi(argv[2]) : 4;
    try {
        Space SA, SB;
        make_case(SA, 600, 0.05, 0.5, 350);
        make_case(SB, 600, 0.05, 0.5, 350);
        double dl = SA.average_segment_length();
        double mn = dl > 0 ? dl * 5 : 0, mx = dl > 0 ? dl * 100 : std::numeric_limits<double>::max();
        if (!strcmp(mode, "free")) {
            RefArm pre(SB, mn, mx);   // SLAE / shedding: the reference's classes on the same Space
            GpuArm hot(SB, mn, mx);
            for (int k = 0; k < nsteps; k++) {
                pre.collision = hot.collision;
                pre.pre();
                TVec3D f = SB.BodyList[0]->force_hydro;
                // stepdata stores float32 (MStepdata.cpp:284); vvxtract prints that with %+.6e
                printf("%+.6e %+.6e %+.6e %+.6e %zu\n", double(SB.time), (double)(float)f.r.x, (double)(float)f.r.y,
                       (double)(float)f.o, SB.VortexList.size());
                SB.zero_forces();
                hot.collision = pre.collision;
                size_t cleaned = 0;
                hot.hot(&cleaned);
                SB.time = TTime::add(SB.time, SB.dt);
            }
            return 0;
        }
        // lockstep
        RefArm A(SA, mn, mx);
        GpuArm B(SB, mn, mx);
        double worst = 0;
        const bool points = !strcmp(mode, "points");
        for (int k = 0; k < nsteps; k++) {
            A.pre();
            SA.zero_forces();
            // clone the state the hot path starts from
            SB.VortexList = SA.VortexList;
            *SB.BodyList[0] = *SA.BodyList[0];
            SB.time = SA.time;
            size_t ca = 0, cb = 0;
            int ma = A.hot(&ca);
            int mb = B.hot(&cb);
            if (ma != mb || ca != cb || SA.VortexList.size() != SB.VortexList.size()) {
                printf("FAIL step %d: merged %d/%d cleaned %zu/%zu survivors %zu/%zu\n", k, ma, mb, ca, cb,
                       SA.VortexList.size(), SB.VortexList.size());
                return 1;
            }
            double vmax = 0, e = 0;
            for (size_t i = 0; i < SA.VortexList.size(); i++) {
                const TObj &a = SA.VortexList[i], &b = SB.VortexList[i];
                if (memcmp(&a.g, &b.g, 8) || memcmp(&a._1_eps, &b._1_eps, 8)) {
                    printf("FAIL step %d: particle %zu g/_1_eps not bit-exact\n", k, i);
                    return 1;
                }
                vmax = fmax(vmax, fmax(fabs(a.r.x), fabs(a.r.y)));
                e = fmax(e, fmax(fabs(a.r.x - b.r.x), fabs(a.r.y - b.r.y)));
            }
            e /= vmax;
            TBody &ba = *SA.BodyList[0], &bb = *SB.BodyList[0];
            double fs = fabs(ba.fdt_dead.r.x) + fabs(ba.fdt_dead.r.y) + fabs(ba.fdt_dead.o);
            e = fmax(e, relerr(ba.fdt_dead.r.x, bb.fdt_dead.r.x, fs));
            e = fmax(e, relerr(ba.fdt_dead.r.y, bb.fdt_dead.r.y, fs));
            e = fmax(e, relerr(ba.fdt_dead.o, bb.fdt_dead.o, fs));
            e = fmax(e, relerr(ba.g_dead, bb.g_dead, fabs(ba.g_dead)));
            double fr = 0, gs = 0, efr = 0, egs = 0;
            for (size_t s = 0; s < ba.alist.size(); s++) {
                fr = fmax(fr, fabs(ba.alist[s].fric)); gs = fmax(gs, fabs(ba.alist[s].gsum));
                efr = fmax(efr, fabs(ba.alist[s].fric - bb.alist[s].fric));
                egs = fmax(egs, fabs(ba.alist[s].gsum - bb.alist[s].gsum));
            }
            e = fmax(e, fmax(fr > 0 ? efr / fr : 0, gs > 0 ? egs / gs : 0));
            worst = fmax(worst, e);
            printf("step %d N=%zu merged=%d cleaned=%zu relerr=%.3e\n", k, SA.VortexList.size(), ma, ca, e);
            SA.time = TTime::add(SA.time, SA.dt);
        }
        if (points) {
            // SURVEY 8(f) rows 1 and 4 through the C++ adapter: same state in both Spaces, trees built, epsilon done
            A.pre();
            SA.zero_forces();
            SB.VortexList = SA.VortexList;
            *SB.BodyList[0] = *SA.BodyList[0];
            SB.time = SA.time;
            A.tr.build(); B.tr.build();
            A.epsilon.CalcEpsilonFast(true); B.epsilon.CalcEpsilonFast(true);
            std::vector<TVec> pts, vb;
            for (int j = 0; j < 40; j++)
                for (int i = 0; i < 60; i++) pts.push_back(TVec(-1.0 + 0.07 * i, -1.2 + 0.06 * j));
            vb.resize(pts.size());
            B.convective.velocity(pts.data(), pts.size(), vb.data());
            double vs = 0, ve = 0;
            for (size_t k = 0; k < pts.size(); k++) {
                TVec va = A.convective.velocity(pts[k]);
                vs = fmax(vs, fmax(fabs(va.x), fabs(va.y)));
                ve = fmax(ve, fmax(fabs(va.x - vb[k].x), fabs(va.y - vb[k].y)));
            }
            TVec one = B.convective.velocity(pts[7]);
            if (one.x != vb[7].x || one.y != vb[7].y) { printf("FAIL velocity(p) one-point form differs from the batched form\n"); return 1; }
            std::vector<double> nb = B.convective.NodeInfluence();
            double ns = 0, ne = 0;
            size_t k = 0;
            for (auto& lbody : SA.BodyList)
                for (auto& latt : lbody->alist) {
                    double na = A.convective.NodeInfluence(*A.tr.findNode(latt.r), latt);
                    ns = fmax(ns, fabs(na));
                    ne = fmax(ne, fabs(na - nb[k++]));
                }
            A.tr.destroy(); B.tr.destroy();
            printf("points: velocity(p) on %zu points relerr=%.3e; NodeInfluence on %zu segments relerr=%.3e\n", pts.size(),
                   ve / vs, nb.size(), ne / ns);
            worst = fmax(worst, fmax(ve / vs, ne / ns));
        }
        printf("%s worst=%.3e\n", worst <= 1e-10 ? "OK" : "FAIL", worst);
        return worst <= 1e-10 ? 0 : 1;
    } catch (const std::exception& e) {
        printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
