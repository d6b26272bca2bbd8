// K5 — the diffusive velocity, MDiffusiveFast::process_vort_list / vortex_influence / segment_influence
// (libvvhd/src/MDiffusiveFast.cpp:8-48, 93-123).
//
// Same work decomposition as the other near-field passes (a CTA owns one work unit, its warps pull target
// leaves, a warp expands the source leaves within the leaf's exact 8-epsilon reach into a flat candidate
// list in shared memory). What is specific to this pass: only ~1/4 of the examined pairs lie inside the
// cut-off `|dr| / eps <= 8` (:101), and those pay an exp(), a square root and a division — ten times the
// cost of the test that rejects the others. The previous sub-lane kernel evaluated that body under
// divergence (ncu: 16 of 32 lanes active, the exp path entered on almost every step), so it cost as
// much as if every examined pair were a hit. Here the two halves are separated:
//   examine  lanes = candidates (two per lane and step), one target at a time (broadcast from shared
//            memory): 4 FP64 instructions + compare + sign test; hits are compacted with a ballot
//            into a small queue of candidate indices;
//   evaluate whenever the queue holds 32 hits, ALL lanes take one and run the exp body densely;
//            the three sums of the current target stay in registers and are reduced across lanes
//            once per (target, chunk).
// exp(x) for x in [-8, 0]: n = rint(x log2 e), r = x - n ln 2 (two-term), degree-11 Taylor polynomial
// (truncation 6e-15 relative), exponent patched in — 19 instructions, no special cases; the library exp
// spent a quarter of the kernel's instructions on loading its constants into uniform registers.
// The cut-off decision itself is replayed exactly (df_hit below).
#pragma once
#include "vvgpu_near.cuh"

namespace vv {

#ifndef VV_DF_MINB
#define VV_DF_MINB 2
#endif
#ifndef VV_DF_WARPS
#define VV_DF_WARPS 8
#endif
constexpr int kDfWarps = VV_DF_WARPS;
constexpr int kDfThreads = kDfWarps * 32;
#ifndef VV_DF_FLUSH
#define VV_DF_FLUSH 768
#endif
constexpr int kDfFlush = VV_DF_FLUSH;                  // examine once the candidate list holds this many
constexpr int kDfPiece = 16;
constexpr int kDfCap = kDfFlush + 32 * kDfPiece + 64;
constexpr int kDfQueue = 96;                   // < 32 carried + 64 pushed per step

struct DfWarp {
    int idx[kDfCap];           // candidate particle indices
    int q[kDfQueue];           // hits of the current target
    double tsv[kMaxT][6];      // x, y, ie, ie2, pos, (pad)
};
struct DfShared {
    int4 ent[kUnitEntries];    // first particle, count, first segment, segment count of the entry's leaf
    u32 emk[kUnitEntries];     // target-leaf mask
    float4 ebox[kUnitEntries];   // source-leaf box, rounded OUTWARD to float: the skip below stays conservative
    int bounds[kGroupLeaves + 1];
    int next;
    int anyseg;
    DfWarp w[kDfWarps];
};

__constant__ double kDfExpC[12] = {
    2.50521083854417187751e-08, 2.75573192239858906526e-07, 2.75573192239858906526e-06, 2.48015873015873015873e-05,
    1.98412698412698412698e-04, 1.38888888888888888889e-03, 8.33333333333333333333e-03, 4.16666666666666666667e-02,
    1.66666666666666666667e-01, 0.5, 1.0, 1.0};   // 1/11! ... 1/0!  (operands straight from the constant bank)
__device__ __forceinline__ double df_exp(double x) {   // x in [-8.1, 0]
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);
    const double n = t - 6755399441055744.0;
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double p = kDfExpC[0];
#pragma unroll
    for (int k = 1; k < 12; k++) p = fma(p, r, kDfExpC[k]);
    const int ni = __double2loint(t);
    return __hiloint2double(__double2hiint(p) + (ni << 20), __double2loint(p));
}

struct DfTarget {   // the target being examined, uniform across the warp
    double x, y, ie, r2;   // r2: padded squared cut-off radius, 64.0001 eps^2
};

// vortex_influence (:93-105) for one candidate that passed the conservative pre-test
__device__ __forceinline__ void df_hit(const DfTarget& T, const double4 s, double& S1, double& S2x, double& S2y) {
    const double dx = VV_SUB(T.x, s.x), dy = VV_SUB(T.y, s.y);
    if (VV_ADD(fabs(dx), fabs(dy)) < 1E-10) return;    // TVec::iszero (:96)
    const double d2 = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
    // |dr| and 1/|dr| from one rsqrt; the exact sqrt of the reference decides only when the cut-off
    // test is within 1e-9 of the boundary
    double rinv = rsqrt(d2);
    double drabs = d2 * rinv;
    double exparg = -VV_MUL(drabs, T.ie);
    if (fabs(exparg + 8.) < 1e-9) {
        drabs = sqrt(d2);
        exparg = -VV_MUL(drabs, T.ie);
        rinv = 1. / drabs;
    }
    if (exparg < -8.) return;
    const double i1tmp = s.z * df_exp(exparg);
    const double q = i1tmp * rinv;
    S2x = fma(dx, q, S2x);
    S2y = fma(dy, q, S2y);
    S1 += i1tmp;
}

// per-phase source views: the full record (x, y, g) for the evaluation and two (x, y) views for the examination,
// one per sign of g, in which every particle of the other sign (or with g == 0) is parked at x = +inf
__global__ void k_pack_diff(int n, Particles P, double4* src4, double2* xy_pos, double2* xy_neg) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > n) return;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double x = inf, y = 0, g = 0;
    if (j < n) { x = P.x[j]; y = P.y[j]; g = P.g[j]; }
    src4[j] = make_double4(g == 0 ? inf : x, y, g, 0.);
    xy_pos[j] = make_double2(g > 0 ? x : inf, y);
    xy_neg[j] = make_double2(g < 0 ? x : inf, y);
}

__global__ void __launch_bounds__(kDfThreads, VV_DF_MINB) k_diff(NearArgs A, DiffOp op, const double2* __restrict__ xy_pos, const double2* __restrict__ xy_neg, int dummy) {
    extern __shared__ __align__(16) unsigned char near_smem[];
    DfShared& S = *reinterpret_cast<DfShared*>(near_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const UnitPart up(A, true);
    const int u = A.u0 + up.b;
    const int g = A.U.group[u];
    const int chunk = u - A.U.first[g];
    const bool multi = A.U.num[g] > 1;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, A.nleaves - l0);
    if (up.lt0 >= nl) return;
    const int lt_end = min(nl, up.lt1);
    if (tid <= nl) S.bounds[tid] = (tid < nl) ? A.L.first[l0 + tid] : A.L.last[l0 + nl - 1];
    if (tid == 0) { S.next = up.lt0; S.anyseg = 0; }
    const long long e0 = A.U.base[u];
    const int ne = A.U.count[u];
    __syncthreads();
    for (int e = tid; e < ne; e += kDfThreads) {
        const int sl = A.G.leaf[e0 + e];
        const int f = A.L.first[sl];
        int4 en = make_int4(f, A.L.last[sl] - f, 0, 0);
        if (A.nseg > 0) {
            en.z = A.L.sfirst[sl]; en.w = A.L.slast[sl] - en.z;
            if (en.w > 0) S.anyseg = 1;
        }
        S.ent[e] = en;
        S.emk[e] = A.G.mask[e0 + e];
        const double* b = A.lbox + 5ll * sl;
        S.ebox[e] = make_float4(__double2float_rd(b[0]), __double2float_ru(b[1]), __double2float_rd(b[2]), __double2float_ru(b[3]));
    }
    __syncthreads();
    const int t0 = S.bounds[0], t1 = S.bounds[nl];
    DfWarp& W = S.w[warp];
    DiffOp::Part* scratch = (DiffOp::Part*)A.scratch;
    const size_t sbase = multi ? ((size_t)A.U.sbase[g] + (size_t)chunk * (t1 - t0)) : 0;
    const u32 lt_mask = lanemask_lt();

    for (;;) {
        int lt = 0;
        if (lane == 0) lt = atomicAdd(&S.next, 1);
        lt = __shfl_sync(kFullMask, lt, 0);
        if (lt >= lt_end) break;
        const int leaf = l0 + lt;
        const int pf = S.bounds[lt], pl = S.bounds[lt + 1];
        for (int tb = pf; tb < pl; tb += kMaxT) {
            const int np = min(kMaxT, pl - tb);
            DiffOp::Tgt tg;
            const int i = tb + lane;
            const bool live = op.init(tg, A, i, leaf, lane < np);
            const u32 lm = __ballot_sync(kFullMask, live);
            const int nt = __popc(lm);
            if (nt == 0) continue;
            const int slot = __popc(lm & lt_mask);
            // box of the live targets and their largest squared reach (exact pruning of source leaves)
            double bx0 = DBL_MAX, bx1 = -DBL_MAX, by0 = DBL_MAX, by1 = -DBL_MAX, R2 = 0;
            if (live) { bx0 = bx1 = tg.x; by0 = by1 = tg.y; R2 = op.reach2(tg); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bx0 = fmin(bx0, __shfl_xor_sync(kFullMask, bx0, o)); bx1 = fmax(bx1, __shfl_xor_sync(kFullMask, bx1, o));
                by0 = fmin(by0, __shfl_xor_sync(kFullMask, by0, o)); by1 = fmax(by1, __shfl_xor_sync(kFullMask, by1, o));
                R2 = fmax(R2, __shfl_xor_sync(kFullMask, R2, o));
            }
            R2 *= 1.000000001;  // the skip stays strictly conservative against rounding in the gap
            if (live) {
                double* ts = W.tsv[slot];
                ts[0] = tg.x; ts[1] = tg.y; ts[2] = tg.ie; ts[3] = 64.0001 / tg.ie2; ts[4] = tg.pos ? 1. : 0.;
                // wall segments of the near leaves (MDiffusiveFast.cpp:26-34)
                if (S.anyseg) {
                    for (int e = 0; e < ne; e++) {
                        if (!((S.emk[e] >> lt) & 1u)) continue;
                        const int4 en = S.ent[e];
                        if (en.w > 0) op.segments(tg, A, en.z, en.z + en.w);
                    }
                }
            }
            __syncwarp();
            // ---- chunks of the candidate list: expand, then examine target by target
            int eb = 0, cnt = 0, f = 0;
            for (;;) {
                int fill = 0;
                bool pending = __any_sync(kFullMask, cnt > 0);
                while (fill < kDfFlush && (pending || eb < ne)) {
                    if (!pending) {
                        const int e = eb + lane;
                        eb += 32;
                        if (e < ne && ((S.emk[e] >> lt) & 1u)) {
                            const int4 en = S.ent[e];
                            f = en.x; cnt = en.y;
                            if (cnt) {
                                const float4 b = S.ebox[e];
                                const double gx = fmax(0., fmax((double)b.x - bx1, bx0 - (double)b.y));
                                const double gy = fmax(0., fmax((double)b.z - by1, by0 - (double)b.w));
                                if (gx * gx + gy * gy > R2) cnt = 0;
                            }
                        }
                        pending = __any_sync(kFullMask, cnt > 0);
                        continue;
                    }
                    const int c = min(cnt, kDfPiece);
                    int inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                        asm volatile("{ .reg .pred p; .reg .s32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, t; }"
                                     : "+r"(inc) : "r"(o));
                    const int tot = __shfl_sync(kFullMask, inc, 31);
                    int* dst = W.idx + fill + inc - c;
#pragma unroll
                    for (int k = 0; k < kDfPiece; k++)
                        if (k < c) dst[k] = f + k;
                    fill += tot; cnt -= c; f += c;
                    pending = __any_sync(kFullMask, cnt > 0);
                }
                __syncwarp();
                const bool final = !pending && eb >= ne;
                if (fill > 0) {
                    // pad to whole steps of 64 with the dummy record (x = +inf: never a hit)
                    const int upto = (fill + 63) & ~63;
                    if (fill + lane < upto) W.idx[fill + lane] = dummy;
                    if (fill + lane + 32 < upto) W.idx[fill + lane + 32] = dummy;
                    __syncwarp();
                    for (int t = 0; t < nt; t++) {
                        const double* ts = W.tsv[t];
                        DfTarget T;
                        T.x = ts[0]; T.y = ts[1]; T.ie = ts[2]; T.r2 = ts[3];
                        // same-sign sources only (:95): the view of the other sign (and of g == 0) holds x = +inf
                        const double2* __restrict__ xy = (ts[4] != 0.) ? xy_pos : xy_neg;
                        double S1 = 0, S2x = 0, S2y = 0;
                        int qn = 0;
                        int2 jj = *reinterpret_cast<const int2*>(&W.idx[2 * lane]);
                        double2 ca = xy[jj.x], cb = xy[jj.y];
                        for (int k0 = 0; k0 < upto; k0 += 64) {
                            // examine two candidates per lane: squared distance against the padded cut-off
                            const int2 jc = jj;
                            const double2 a = ca, b = cb;
                            if (k0 + 64 < upto) {   // next step's candidates, behind this step's work
                                jj = *reinterpret_cast<const int2*>(&W.idx[k0 + 64 + 2 * lane]);
                                ca = xy[jj.x]; cb = xy[jj.y];
                            }
                            const double dxa = T.x - a.x, dya = T.y - a.y;
                            const double dxb = T.x - b.x, dyb = T.y - b.y;
                            const double d2a = fma(dxa, dxa, dya * dya), d2b = fma(dxb, dxb, dyb * dyb);
                            const bool ha = !(d2a > T.r2), hb = !(d2b > T.r2);
                            const u32 ba = __ballot_sync(kFullMask, ha), bb = __ballot_sync(kFullMask, hb);
                            if (ha) W.q[qn + __popc(ba & lt_mask)] = jc.x;
                            qn += __popc(ba);
                            if (hb) W.q[qn + __popc(bb & lt_mask)] = jc.y;
                            qn += __popc(bb);
                            __syncwarp();
                            while (qn >= 32) {   // dense evaluation: every lane takes one hit
                                qn -= 32;
                                df_hit(T, A.src4[W.q[qn + lane]], S1, S2x, S2y);
                            }
                            __syncwarp();
                        }
                        if (lane < qn) df_hit(T, A.src4[W.q[lane]], S1, S2x, S2y);
                        __syncwarp();
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            S1 += __shfl_xor_sync(kFullMask, S1, o);
                            S2x += __shfl_xor_sync(kFullMask, S2x, o);
                            S2y += __shfl_xor_sync(kFullMask, S2y, o);
                        }
                        if (live && slot == t) { tg.S1 += S1; tg.S2x += S2x; tg.S2y += S2y; }
                    }
                }
                if (final) break;
            }
            if (live) {
                if (multi) scratch[sbase + (i - t0)] = op.part(tg);
                else op.finish(tg, A, i, leaf);
            }
            __syncwarp();
        }
    }
}

}  // namespace vv
