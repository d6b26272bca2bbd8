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
//   resident N  SURVEY 8(f) rows 1 + 2 together: the vortex list LIVES on the device (vvgpu::set_resident), shed vortices
//               are appended to it, the SLAE right-hand side takes MConvectiveFast::NodeInfluence() from the device tree
//               (the patch of INTEGRATION.md 2b, applied here between the reference's fill_matrix and its solve), so no
//               CPU particle tree is ever built and nothing but the shed vortices crosses PCIe. Prints the same rows as
//               `free` (so the README rows can be checked), runs an all-reference arm beside it and reports steps/s.
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
#include <omp.h>

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

// The SLAE stage without a CPU particle tree: the reference's own fill_matrix + solve, on a CPU tree that holds the BODY
// segments only (findNode works, NodeInfluence is 0), with the free vortices' term of every slip equation
// (MConvectiveFast.cpp:459-467) taken from the device tree. This is MConvectiveFast::calc_circulation (:235-273) with
// the one-line patch of INTEGRATION.md 2b; the reference's sources stay unmodified (private members reached through the
// `#define private public` above, test infrastructure only).
static void calc_circulation_device_rhs(Space& S, MConvectiveFast& convective, TSortedTree& bodytree, const std::vector<double>& ni,
                                        const void** collision) {
    bool use_inverse = (*collision == nullptr) && convective.can_use_inverse() && S.time > 0;
    Matrix& matrix = convective.matrix;
    matrix.resize(S.total_segment_count() + S.BodyList.size() * 9);
    const bool right_only = matrix.bodyMatrixIsOk() && use_inverse;
    if (!right_only) {
        unsigned eq_no = 0;
        for (auto& libody : S.BodyList) {
            for (auto& latt : libody->alist) matrix.setSolutionForCol(eq_no++, &latt.g);
            matrix.setSolutionForCol(eq_no++, &libody->speed_slae.r.x);
            matrix.setSolutionForCol(eq_no++, &libody->speed_slae.r.y);
            matrix.setSolutionForCol(eq_no++, &libody->speed_slae.o);
            matrix.setSolutionForCol(eq_no++, &libody->force_hydro.r.x);
            matrix.setSolutionForCol(eq_no++, &libody->force_hydro.r.y);
            matrix.setSolutionForCol(eq_no++, &libody->force_hydro.o);
            matrix.setSolutionForCol(eq_no++, &libody->force_holder.r.x);
            matrix.setSolutionForCol(eq_no++, &libody->force_holder.r.y);
            matrix.setSolutionForCol(eq_no++, &libody->force_holder.o);
        }
    }
    bodytree.build(/*IncludeVortexes=*/false, /*IncludeBody=*/true, /*IncludeHeat=*/false);
    convective.fill_matrix(right_only, collision);
    bodytree.destroy();
    unsigned eq_no = 0;
    size_t k = 0;
    int once = 0;
    for (auto& libody : S.BodyList) {   // the slip equations are those of every segment but the body's special one (:896-914)
        TAtt* special = &libody->alist[libody->special_segment_no];
        for (auto& latt : libody->alist) {
            if (&latt != special) *matrix.getRightCol(eq_no) -= ni[k];
            // the circulation equation (:508-512) reads Space::gsum(), the sum over the HOST list, which is empty while
            // the list is resident: the device's sum takes its place
            else if (libody->boundary_condition != bc_t::kutta && !once++) *matrix.getRightCol(eq_no) += S.gsum() - vvgpu::gsum(&S);
            eq_no++; k++;
        }
        eq_no += 9;
    }
    matrix.solveUsingInverseMatrix(use_inverse);
}

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
        if (!strcmp(mode, "resident")) {
            const bool resident = getenv("VV_DROPIN_NORESIDENT") == nullptr;   // (diagnostic: device right-hand side alone)
            vvgpu::set_resident(&SB, resident);
            RefArm A(SA, mn, mx);                         // the all-reference run beside it
            TSortedTree bodytree(&SB, 8, mn, mx);         // never holds a particle
            MConvectiveFast slae(&SB, &bodytree);
            GpuArm B(SB, mn, mx);
            double worst = 0, t_gpu = 0, t_ref = 0;
            for (int k = 0; k < nsteps; k++) {
                double t0 = omp_get_wtime();
                A.pre();
                TVec3D fa = SA.BodyList[0]->force_hydro;
                SA.zero_forces();
                size_t ca = 0;
                A.hot(&ca);
                SA.time = TTime::add(SA.time, SA.dt);
                t_ref += omp_get_wtime() - t0;
                t0 = omp_get_wtime();
                // vvflow.cpp:216-227 on the device tree
                B.tr.build();
                std::vector<double> ni = B.convective.NodeInfluence();
                if (B.collision != nullptr) calc_circulation_device_rhs(SB, slae, bodytree, ni, &B.collision);
                calc_circulation_device_rhs(SB, slae, bodytree, ni, &B.collision);
                B.tr.destroy();
                B.flowmove.heat_shed();
                B.flowmove.vortex_shed();                  // appended behind the device list
                B.flowmove.streak_shed();
                SB.calc_forces();
                TVec3D f = SB.BodyList[0]->force_hydro;
                printf("%+.6e %+.6e %+.6e %+.6e %zu\n", double(SB.time), (double)(float)f.r.x, (double)(float)f.r.y,
                       (double)(float)f.o, vvgpu::particle_count(&SB));
                if (getenv("VV_DROPIN_DEBUG")) {
                    TBody &ba = *SA.BodyList[0], &bb = *SB.BodyList[0];
                    double sga = 0, sgb = 0, fra = 0, frb = 0, gsa = 0, gsb = 0;
                    for (size_t q = 0; q < ba.alist.size(); q++) { sga += ba.alist[q].g; sgb += bb.alist[q].g; fra += ba.alist[q].fric; frb += bb.alist[q].fric; gsa += ba.alist[q].gsum; gsb += bb.alist[q].gsum; }
                    printf("   dbg step %d: sum g %.12e / %.12e  fric %.12e / %.12e  gsum %.12e / %.12e  friction_prev.o %.6e / %.6e\n", k, sga, sgb, fra, frb, gsa, gsb,
                           ba.friction_prev.o, bb.friction_prev.o);
                }
                SB.zero_forces();
                size_t cb = 0;
                B.hot(&cb);
                if (getenv("VV_DROPIN_DEBUG")) {
                    TBody &ba = *SA.BodyList[0], &bb = *SB.BodyList[0];
                    printf("   dbg after hot %d: g_dead %.12e / %.12e fdt_dead.o %.12e / %.12e cleaned %zu/%zu N %zu/%zu\n", k, ba.g_dead, bb.g_dead, ba.fdt_dead.o, bb.fdt_dead.o, ca, cb,
                           SA.VortexList.size(), vvgpu::particle_count(&SB));
                }
                SB.time = TTime::add(SB.time, SB.dt);
                t_gpu += omp_get_wtime() - t0;
                if (resident && !SB.VortexList.empty()) { printf("FAIL step %d: the host list is not empty in resident mode\n", k); return 1; }
                double fs = fabs(fa.r.x) + fabs(fa.r.y) + fabs(fa.o);
                double e = (fabs(fa.r.x - f.r.x) + fabs(fa.r.y - f.r.y) + fabs(fa.o - f.o)) / (fs > 0 ? fs : 1);
                worst = fmax(worst, e);
                if (vvgpu::particle_count(&SB) != SA.VortexList.size()) {
                    printf("FAIL step %d: particle count %zu vs %zu\n", k, vvgpu::particle_count(&SB), SA.VortexList.size());
                    return 1;
                }
            }
            // the resident list against the reference's, at the end (free-running for nsteps: round-off drifts apart slowly)
            vvgpu::sync_to_host(&SB);
            double pe = 0;
            for (size_t i = 0; i < SA.VortexList.size(); i++)
                pe = fmax(pe, fmax(fabs(SA.VortexList[i].r.x - SB.VortexList[i].r.x), fabs(SA.VortexList[i].r.y - SB.VortexList[i].r.y)));
            printf("resident: %d steps, N=%zu, force_hydro relerr vs the reference run %.3e, final positions max diff %.3e, "
                   "steps/s device-resident %.2f vs reference classes %.2f\n", nsteps, SA.VortexList.size(), worst, pe,
                   nsteps / t_gpu, nsteps / t_ref);
            printf("%s worst=%.3e ranks=%zu\n", (worst <= 1e-8 && pe <= 1e-8) ? "OK" : "FAIL", fmax(worst, pe), vvgpu::Device::of(&SB)->all.size());
            return (worst <= 1e-8 && pe <= 1e-8) ? 0 : 1;
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
            // stree::getBottomNodes / findNode with the reference's signatures, on the host mirror of the device tree
            const std::vector<TSortedNode*>& la = A.tr.getBottomNodes();
            const std::vector<TSortedNode*>& lb = B.tr.getBottomNodes();
            const size_t nleaves_checked = la.size();
            if (la.size() != lb.size()) { printf("FAIL getBottomNodes: %zu vs %zu leaves\n", la.size(), lb.size()); return 1; }
            for (size_t l = 0; l < la.size(); l++) {
                const TSortedNode &a = *la[l], &b = *lb[l];
                bool same = a.x == b.x && a.y == b.y && a.h == b.h && a.w == b.w && a.vRange.last - a.vRange.first == b.vRange.last - b.vRange.first &&
                            a.bllist.size() == b.bllist.size() && a.NearNodes->size() == b.NearNodes->size() &&
                            a.FarNodes->size() == b.FarNodes->size() && a.CMp.g == b.CMp.g && a.CMm.r.x == b.CMm.r.x;
                for (size_t q = 0; same && q < a.NearNodes->size(); q++) same = (*a.NearNodes)[q]->x == (*b.NearNodes)[q]->x && (*a.NearNodes)[q]->y == (*b.NearNodes)[q]->y;
                for (size_t q = 0; same && q < a.FarNodes->size(); q++) same = (*a.FarNodes)[q]->x == (*b.FarNodes)[q]->x && (*a.FarNodes)[q]->CMp.g == (*b.FarNodes)[q]->CMp.g;
                for (size_t q = 0; same && q < a.bllist.size(); q++) same = a.bllist[q]->r.x == b.bllist[q]->r.x && a.bllist[q]->r.y == b.bllist[q]->r.y;
                if (same && a.vRange.first < a.vRange.last) same = a.vRange.first->r.x == b.vRange.first->r.x && a.vRange.first->g == b.vRange.first->g;
                if (!same) { printf("FAIL getBottomNodes: leaf %zu differs from the reference's\n", l); return 1; }
            }
            for (size_t q = 0; q < pts.size(); q++) {
                const TSortedNode *a = A.tr.findNode(pts[q]), *b = B.tr.findNode(pts[q]);
                if (a->x != b->x || a->y != b->y || a->h != b->h || a->w != b->w) { printf("FAIL findNode at point %zu\n", q); return 1; }
                // the reference's own static evaluators walk the mirror like their own tree
                if (MEpsilonFast::eps2h(*a, pts[q]) != MEpsilonFast::eps2h(*b, pts[q]) || MEpsilonFast::h2(*a, pts[q]) != MEpsilonFast::h2(*b, pts[q])) {
                    printf("FAIL MEpsilonFast::eps2h / h2 on the mirrored node at point %zu\n", q); return 1;
                }
            }
            A.tr.destroy(); B.tr.destroy();
            bool threw = false;
            try { B.tr.findNode(pts[0]); } catch (const std::invalid_argument&) { threw = true; }
            if (!threw) { printf("FAIL findNode on an unbuilt tree did not throw\n"); return 1; }
            printf("points: velocity(p) on %zu points relerr=%.3e; NodeInfluence on %zu segments relerr=%.3e; getBottomNodes / findNode "
                   "mirror identical on %zu leaves\n", pts.size(), ve / vs, nb.size(), ne / ns, nleaves_checked);
            worst = fmax(worst, fmax(ve / vs, ne / ns));
        }
        printf("%s worst=%.3e ranks=%zu\n", worst <= 1e-10 ? "OK" : "FAIL", worst, vvgpu::Device::of(&SB)->all.size());
        return worst <= 1e-10 ? 0 : 1;
    } catch (const std::exception& e) {
        printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
