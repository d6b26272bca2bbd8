// K3 / K4 / K5 — the three near-field passes over the same pair set:
//   EpsOp   MEpsilonFast::epsv + merge decision        (libvvhd/src/MEpsilonFast.cpp:128-173)
//   ConvOp  MConvectiveFast::near_nodes_influence      (libvvhd/src/MConvectiveFast.cpp:116-137)
//           + the per-particle assembly of process_all_lists (:76-86)
//   DiffOp  MDiffusiveFast::process_vort_list          (libvvhd/src/MDiffusiveFast.cpp:8-48,93-123)
//
// One CTA owns one group of 32 consecutive leaves. Its targets are a contiguous slice of the
// permuted particle array (one target per thread, in chunks of kNearThreads); the group's near
// source leaves are streamed through shared memory in tiles and every thread walks the tile with
// broadcast loads. A source leaf is applied to a target only if the leaf's bit is set in the
// entry's mask, i.e. exactly the (leaf, near leaf) pairs of the reference; a warp skips entries
// none of its leaves see.
#pragma once
#include "vvgpu_lists.cuh"

namespace vv {

constexpr int kNearThreads = 384;   // one target per thread: 32 leaves x ~10.4 particles fit in one pass
constexpr int kNearEB = 64;         // list entries per batch
constexpr int kNearTS = 1024;       // source particles per shared-memory tile
constexpr int kUnitEntries = 512;   // list entries per work unit (bounds the serial path of one thread)

struct Particles {
    double *x, *y, *g, *vx, *vy, *ie;
};

// Work units: a group whose list is longer than kUnitEntries is cut into several units (a few
// fringe leaves see tens of thousands of near leaves). Units of such a group
// park their partial per-target state in `scratch`; k_near_finalize combines them in unit order.
struct Units {
    const int* group;       // unit -> group
    const int* first;       // group -> first unit (ngroups + 1)
    const u32* sbase;       // group -> first scratch slot (ngroups + 1)
};

struct NearArgs {
    Particles P;
    LeafDev L;
    GroupLists G;
    Units U;
    void* scratch;
    const double4* src4;  // per particle, packed by k_pack_src for the running phase
    const double* lbox;   // per leaf 5 doubles: min x, max x, min y, max y of its particles NOW, max eps (k_leaf_box)
    int nleaves;
    int nseg;
    int u0;  // first unit of this launch (shard offset)
    // segments (diffusive / epsilon wall terms)
    const int* seg_perm;
    const double *srx, *sry, *sdlx, *sdly;
};

struct NearShared {
    double2 sxy[kNearTS];
    double2 sab[kNearTS];
    int sj[kNearTS];                 // particle index of each staged source
    int epre[kUnitEntries + 1];      // source prefix per entry of the unit (0 sources once filtered out)
    int epf[kUnitEntries];           // first particle of the entry's leaf
    u32 emk[kUnitEntries];           // target-leaf mask
    int eleaf[kUnitEntries];
    int rstart[kUnitEntries + 1];    // runs of consecutive entries with the same mask (units are mask-sorted)
    u32 rmask[kUnitEntries];
    int wsum[8];
    int nruns;
    double gbox[5];                  // group's target box + cut-off radius (ops with kFilter)
    int bounds[kGroupLeaves + 1];
};

// exclusive prefix sum of v[0..kUnitEntries) in place, by threads 0..127 (4 entries each); returns the
// total in *total. Every thread of the CTA must call it (it contains barriers).
__device__ __forceinline__ void unit_scan(int* v, int* wsum, int* total_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0, inc = 0;
    if (tid < kUnitEntries / 4) {
        a0 = v[4 * tid]; a1 = v[4 * tid + 1]; a2 = v[4 * tid + 2]; a3 = v[4 * tid + 3];
        inc = a0 + a1 + a2 + a3;
        const int s = inc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        inc -= s;  // exclusive within the warp
    }
    __syncthreads();
    if (tid < kUnitEntries / 4) {
        int off = 0;
        for (int w = 0; w < warp; w++) off += wsum[w];
        int ex = off + inc;
        v[4 * tid] = ex; v[4 * tid + 1] = ex + a0; v[4 * tid + 2] = ex + a0 + a1; v[4 * tid + 3] = ex + a0 + a1 + a2;
        if (tid == kUnitEntries / 4 - 1) *total_out = ex + a0 + a1 + a2 + a3;
    }
    __syncthreads();
}

template <class Op>
__global__ void __launch_bounds__(kNearThreads) k_near(NearArgs A, Op op) {
    extern __shared__ __align__(16) unsigned char near_smem[];
    NearShared& S = *reinterpret_cast<NearShared*>(near_smem);
    const int tid = threadIdx.x;
    const int u = A.u0 + blockIdx.x;
    const int g = A.U.group[u];
    const int chunk = u - A.U.first[g];
    const bool multi = (A.U.first[g + 1] - A.U.first[g]) > 1;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, A.nleaves - l0);
    if (tid <= nl) S.bounds[tid] = (tid < nl) ? A.L.first[l0 + tid] : A.L.last[l0 + nl - 1];
    if (Op::kFilter && tid < 32) {  // box of the group's targets and the largest cut-off radius among them
        double x0 = DBL_MAX, x1 = -DBL_MAX, y0 = DBL_MAX, y1 = -DBL_MAX, em = 0;
        if (tid < nl) {
            const double* b = A.lbox + 5ll * (l0 + tid);
            x0 = b[0]; x1 = b[1]; y0 = b[2]; y1 = b[3]; em = b[4];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
            em = fmax(em, __shfl_xor_sync(0xffffffffu, em, o));
        }
        if (tid == 0) { S.gbox[0] = x0; S.gbox[1] = x1; S.gbox[2] = y0; S.gbox[3] = y1; S.gbox[4] = op.reach(em); }
    }
    __syncthreads();
    const int t0 = S.bounds[0], t1 = S.bounds[nl];
    const long long e0 = A.G.ptr[g] + (long long)chunk * kUnitEntries;
    const int ne = (int)(min(A.G.ptr[g + 1], e0 + kUnitEntries) - e0);
    typename Op::Part* scratch = (typename Op::Part*)A.scratch;

    // ---- prologue: the unit's whole entry table, in parallel (one entry per thread)
    for (int e = tid; e < kUnitEntries; e += kNearThreads) {
        int cnt = 0, f = 0, sl = 0;
        u32 mk = 0;
        if (e < ne) {
            sl = A.G.leaf[e0 + e];
            mk = A.G.mask[e0 + e];
            f = A.L.first[sl];
            cnt = A.L.last[sl] - f;
            if (Op::kFilter) {  // gap between the source leaf's box and the group's box
                const double* b = A.lbox + 5ll * sl;
                double gx = fmax(0., fmax(b[0] - S.gbox[1], S.gbox[0] - b[1]));
                double gy = fmax(0., fmax(b[2] - S.gbox[3], S.gbox[2] - b[3]));
                if (gx * gx + gy * gy > S.gbox[4] * S.gbox[4]) cnt = 0;
            }
        }
        S.epre[e] = cnt; S.epf[e] = f; S.emk[e] = mk; S.eleaf[e] = sl;
    }
    __syncthreads();
    // a run opens where the mask changes (only among entries that still have sources)
    for (int e = tid; e < kUnitEntries; e += kNearThreads) {
        int flag = 0;
        if (e < ne && S.epre[e] > 0) {
            int p = e - 1;
            while (p >= 0 && S.epre[p] == 0) p--;   // previous surviving entry
            flag = (p < 0) || (S.emk[p] != S.emk[e]);
        }
        S.rstart[e] = flag;
    }
    __syncthreads();
    int total, nruns;
    unit_scan(S.epre, S.wsum, &S.rstart[kUnitEntries]);   // total sources parked in rstart[kUnitEntries]
    total = S.rstart[kUnitEntries];
    __syncthreads();
    if (tid == 0) S.epre[kUnitEntries] = total;
    // run ids: scan the flags, then every opening entry records its run
    int myflag[2] = {0, 0};
    for (int q = 0, e = tid; e < kUnitEntries; e += kNearThreads, q++) myflag[q] = S.rstart[e];
    __syncthreads();
    unit_scan(S.rstart, S.wsum, &S.nruns);
    nruns = S.nruns;
    int myrid[2];
    for (int q = 0, e = tid; e < kUnitEntries; e += kNearThreads, q++) myrid[q] = S.rstart[e];
    __syncthreads();
    for (int q = 0, e = tid; e < kUnitEntries; e += kNearThreads, q++)
        if (myflag[q]) { S.rstart[myrid[q]] = S.epre[e]; S.rmask[myrid[q]] = S.emk[e]; }
    if (tid == 0) S.rstart[nruns] = total;
    __syncthreads();

    for (int tb = t0; tb < t1; tb += kNearThreads) {
        const int i = tb + tid;
        const bool inrange = i < t1;
        int lt = 0;
        if (inrange) {  // largest k with bounds[k] <= i
            int lo = 0, hi = nl - 1;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (S.bounds[mid] <= i) lo = mid; else hi = mid - 1;
            }
            lt = lo;
        }
        typename Op::Tgt tg;
        const bool live = op.init(tg, A, i, l0 + lt, inrange);
        if (live) op.seed(tg, A, l0 + lt);
        // leaves covered by this warp's live targets
        int ltmin = live ? lt : 64, ltmax = live ? lt : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ltmin = min(ltmin, __shfl_xor_sync(0xffffffffu, ltmin, o));
            ltmax = max(ltmax, __shfl_xor_sync(0xffffffffu, ltmax, o));
        }
        u32 wmask = 0;
        if (ltmax >= 0) {
            u32 hi = (ltmax == 31) ? 0xffffffffu : ((1u << (ltmax + 1)) - 1u);
            wmask = hi & ~((1u << ltmin) - 1u);
        }
        const u32 mybit = 1u << lt;

        for (int s0 = 0; s0 < total; s0 += kNearTS) {
            if (s0 > 0 || tb > t0) __syncthreads();  // the previous tile has been consumed
            const int nsrc = min(kNearTS, total - s0);
            for (int k = tid; k < nsrc; k += kNearThreads) {
                const int F = s0 + k;
                int lo = 0, hi = ne - 1;  // largest e with epre[e] <= F (entries without sources never match)
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    if (S.epre[mid] <= F) lo = mid; else hi = mid - 1;
                }
                const int j = S.epf[lo] + (F - S.epre[lo]);
                const double4 v = A.src4[j];
                S.sxy[k] = make_double2(v.x, v.y);
                S.sab[k] = make_double2(v.z, v.w);
                S.sj[k] = j;
            }
            __syncthreads();
            if (wmask) {
                int r = 0;
                {   // first run that reaches into this tile
                    int lo = 0, hi = nruns - 1;
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if (S.rstart[mid + 1] > s0) hi = mid; else lo = mid + 1;
                    }
                    r = lo;
                }
                for (; r < nruns && S.rstart[r] < s0 + kNearTS; r++) {
                    const u32 m = S.rmask[r];
                    if (!(m & wmask)) continue;
                    const int k0 = max(S.rstart[r], s0) - s0, k1 = min(S.rstart[r + 1], s0 + kNearTS) - s0;
                    if (live && (m & mybit) && k1 > k0) op.run(tg, A, S, k0, k1 - k0);
                }
            }
        }
        if (Op::kSegments && A.nseg > 0 && live) {
            for (int e = 0; e < ne; e++) {
                if (!(S.emk[e] & mybit)) continue;
                const int sl = S.eleaf[e];
                const int sf = A.L.sfirst[sl], se = A.L.slast[sl];
                if (se > sf) op.segments(tg, A, sf, se);
            }
        }
        if (multi) {
            if (live) scratch[(size_t)A.U.sbase[g] + (size_t)chunk * (t1 - t0) + (i - t0)] = op.part(tg);
        } else if (live) op.finish(tg, A, i, l0 + lt);
    }
}

// combine the parked partial states of a multi-unit group, in unit order, and finish
template <class Op>
__global__ void __launch_bounds__(256) k_near_finalize(NearArgs A, Op op, int g0, int g1) {
    const int g = g0 + blockIdx.x;
    if (g >= g1) return;
    const int nu = A.U.first[g + 1] - A.U.first[g];
    if (nu <= 1) return;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, A.nleaves - l0);
    const int t0 = A.L.first[l0], t1 = A.L.last[l0 + nl - 1];
    const typename Op::Part* scratch = (const typename Op::Part*)A.scratch + A.U.sbase[g];
    for (int i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
        int lo = 0, hi = nl - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (A.L.first[l0 + mid] <= i) lo = mid; else hi = mid - 1;
        }
        typename Op::Tgt tg;
        if (!op.init(tg, A, i, l0 + lo, true)) continue;
        op.seed(tg, A, l0 + lo);
        for (int c = 0; c < nu; c++) op.combine(tg, scratch[(size_t)c * (t1 - t0) + (i - t0)]);
        op.finish(tg, A, i, l0 + lo);
    }
}

// unit bookkeeping: units per group, scratch slots per group
__global__ void k_unit_count(LeafDev L, GroupLists G, int nleaves, int ngroups, u32* nunits, u32* nslots) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    long long ne = G.ptr[g + 1] - G.ptr[g];
    u32 nu = (u32)((ne + kUnitEntries - 1) / kUnitEntries);
    if (nu == 0) nu = 1;
    int l0 = g * kGroupLeaves, nl = min(kGroupLeaves, nleaves - l0);
    u32 T = (u32)(L.last[l0 + nl - 1] - L.first[l0]);
    nunits[g] = nu;
    nslots[g] = nu > 1 ? nu * T : 0;
}
__global__ void k_unit_fill(int ngroups, const int* first, int* group) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    for (int u = first[g]; u < first[g + 1]; u++) group[u] = g;
}

// per-phase packed source records (one 32-byte load per staged source instead of four scattered ones)
template <class Op>
__global__ void k_pack_src(int n, Particles P, const unsigned char* dyn, double4* out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = Op::pack(P, j, dyn);
}

// Sort the entries of every unit by (mask, leaf): entries that are seen by the same target leaves
// become one long run, so the per-entry bookkeeping of k_near is paid per run. One CTA per unit,
// bitonic sort of <= kUnitEntries 64-bit keys in shared memory. Deterministic.
__global__ void __launch_bounds__(256) k_sort_units(GroupLists G, Units U, int nunits) {
    __shared__ u64 key[kUnitEntries];
    const int u = blockIdx.x;
    if (u >= nunits) return;
    const int g = U.group[u];
    const int chunk = u - U.first[g];
    const long long e0 = G.ptr[g] + (long long)chunk * kUnitEntries;
    const int ne = (int)min((long long)kUnitEntries, G.ptr[g + 1] - e0);
    for (int k = threadIdx.x; k < kUnitEntries; k += blockDim.x)
        key[k] = (k < ne) ? (((u64)G.mask[e0 + k] << 32) | (u32)G.leaf[e0 + k]) : ~0ull;
    __syncthreads();
    for (int size = 2; size <= kUnitEntries; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < kUnitEntries / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = ((lo & size) == 0);
                u64 a = key[lo], b = key[hi];
                if ((a > b) == up) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int k = threadIdx.x; k < ne; k += blockDim.x) {
        G.leaf[e0 + k] = (int)(u32)(key[k] & 0xffffffffull);
        G.mask[e0 + k] = (u32)(key[k] >> 32);
    }
}

// ------------------------------------------------------------------------------------------ K4
struct ConvOp {
    static constexpr bool kSegments = false;
    static constexpr bool kFilter = false;
    __device__ __forceinline__ double reach(double) const { return 0; }
    double inf_vx, inf_vy, eps2_div_srcg;
    const double* taylor;  // 4 per leaf
    const double* sinks;   // (x,y,g) triples
    int nsink;
    struct Tgt { double x, y, rx, ry; };
    struct Part { double rx, ry; };
    __device__ __forceinline__ Part part(const Tgt& t) const { return Part{t.rx, t.ry}; }
    __device__ __forceinline__ void combine(Tgt& t, const Part& p) const { t.rx += p.rx; t.ry += p.ry; }
    __device__ __forceinline__ void seed(Tgt&, const NearArgs&, int) const {}

    __device__ __forceinline__ bool init(Tgt& t, const NearArgs& A, int i, int leaf, bool inrange) const {
        t.rx = t.ry = 0; t.x = t.y = 0;
        if (!inrange) return false;
        if (A.P.g[i] == 0) return false;  // `if (!lobj->g) continue`, MConvectiveFast.cpp:78
        t.x = A.P.x[i]; t.y = A.P.y[i];
        return true;
    }
    // packed source record: eps^2 = sqr(1./_1_eps) of the SOURCE (:119); a g==0 source is skipped by
    // the reference (:130) and is packed as g = 0, eps^2 = 1 so that it adds exactly nothing
    static __device__ __forceinline__ double4 pack(const Particles& P, int j, const unsigned char*) {
        double g = P.g[j];
        double e = 1. / P.ie[j];
        return (g == 0) ? make_double4(P.x[j], P.y[j], 0., 1.) : make_double4(P.x[j], P.y[j], g, e * e);
    }
    // rotl(dr) * g / (|dr|^2 + eps^2): reciprocal by rcp.approx + one third-order correction
    // (relative error ~ e0^3, e0 <= 2^-20: below 1 ulp; exactness is not required of velocities)
    __device__ __forceinline__ void run(Tgt& t, const NearArgs&, const NearShared& S, int k0, int n) const {
        double rx = t.rx, ry = t.ry;
        const double tx = t.x, ty = t.y;
#pragma unroll 4
        for (int k = k0; k < k0 + n; k++) {
            double2 p = S.sxy[k], q = S.sab[k];
            double dx = tx - p.x, dy = ty - p.y;
            double den = fma(dx, dx, fma(dy, dy, q.y));
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
            double e = fma(-den, r0, 1.0);
            double e2 = fma(e, e, e);
            double gr = q.x * r0;
            double w = fma(gr, e2, gr);
            rx = fma(-dy, w, rx);
            ry = fma(dx, w, ry);
        }
        t.rx = rx; t.ry = ry;
    }
    __device__ __forceinline__ void segments(Tgt&, const NearArgs&, int, int) const {}
    __device__ __forceinline__ void finish(Tgt& t, const NearArgs& A, int i, int leaf) const {
        double vx = inf_vx + t.rx * k1_2Pi, vy = inf_vy + t.ry * k1_2Pi;
        if (nsink) {  // sink_list_influence, :153-170
            double sx = 0, sy = 0;
            for (int k = 0; k < nsink; k++) {
                double dx = t.x - sinks[3 * k], dy = t.y - sinks[3 * k + 1], sg = sinks[3 * k + 2];
                double q = sg / (dx * dx + dy * dy + eps2_div_srcg * fabs(sg));
                sx += dx * q; sy += dy * q;
            }
            vx += sx * k1_2Pi; vy += sy * k1_2Pi;
        }
        const double* T = taylor + 4ll * leaf;  // :84-85
        double dlx = t.x - A.L.cx[leaf], dly = t.y - A.L.cy[leaf];
        vx += T[0] + T[2] * dlx + T[3] * dly;
        vy += T[1] + T[3] * dlx - T[2] * dly;
        A.P.vx[i] += vx;
        A.P.vy[i] += vy;
    }
};

// ------------------------------------------------------------------------------------------ K5
struct DiffOp {
    static constexpr bool kSegments = true;
    // only sources within 8 eps of a target contribute (:101): leaves farther than that from the whole
    // group are never staged. 1e-6 relative slack keeps the skip strictly conservative.
    static constexpr bool kFilter = true;
    __device__ __forceinline__ double reach(double epsmax) const { return 8.000008 * epsmax; }
    double re;
    double* fric;  // per segment, atomically accumulated (MDiffusiveFast.cpp:121-122)
    struct Tgt { double x, y, ie, ie2, lim, g, S1, S2x, S2y, S0, S3x, S3y; bool pos; };
    struct Part { double S1, S2x, S2y, S0, S3x, S3y; };
    __device__ __forceinline__ Part part(const Tgt& t) const { return Part{t.S1, t.S2x, t.S2y, t.S0, t.S3x, t.S3y}; }
    __device__ __forceinline__ void combine(Tgt& t, const Part& p) const {
        t.S1 += p.S1; t.S2x += p.S2x; t.S2y += p.S2y; t.S0 += p.S0; t.S3x += p.S3x; t.S3y += p.S3y;
    }
    __device__ __forceinline__ void seed(Tgt&, const NearArgs&, int) const {}

    __device__ __forceinline__ bool init(Tgt& t, const NearArgs& A, int i, int leaf, bool inrange) const {
        t.S1 = t.S2x = t.S2y = t.S0 = t.S3x = t.S3y = 0;
        t.x = t.y = t.ie = t.ie2 = t.g = 0; t.pos = false; t.lim = 64.0001;
        if (!inrange) return false;
        double g = A.P.g[i];
        if (g == 0) return false;
        t.x = A.P.x[i]; t.y = A.P.y[i]; t.g = g; t.ie = A.P.ie[i]; t.ie2 = t.ie * t.ie; t.pos = g > 0;
        return true;
    }
    // a g == 0 source is parked at x = +inf: its distance is inf and the pre-test below drops it
    static __device__ __forceinline__ double4 pack(const Particles& P, int j, const unsigned char*) {
        double g = P.g[j];
        return make_double4(g == 0 ? __longlong_as_double(0x7ff0000000000000ll) : P.x[j], P.y[j], g, 0.);
    }
    // vortex_influence, :93-105. The cut-off decision `-|dr|*_1_eps < -8` is replayed exactly in
    // hit(); a squared-distance pre-test with a 1e-6 safety margin rejects the ~98 % of pairs that
    // are far outside it, so the common path is 7 FP64 instructions and one rarely-taken branch.
    __device__ __forceinline__ void hit(Tgt& t, double dx, double dy, double d2, double sg) const {
        if (sg == 0 || ((sg > 0) != t.pos)) return;        // same sign only (:95)
        if (VV_ADD(fabs(dx), fabs(dy)) < 1E-10) return;    // TVec::iszero
        // |dr| and 1/|dr| from one rsqrt; the exact sqrt of the reference decides only when the
        // cut-off test is within 1e-9 of the boundary
        double rinv = rsqrt(d2);
        double drabs = d2 * rinv;
        double exparg = -VV_MUL(drabs, t.ie);
        if (fabs(exparg + 8.) < 1e-9) {
            drabs = sqrt(d2);
            exparg = -VV_MUL(drabs, t.ie);
            rinv = 1. / drabs;
        }
        if (exparg < -8.) return;
        double i1tmp = sg * exp(exparg);
        double q = i1tmp * rinv;
        t.S2x = fma(dx, q, t.S2x);
        t.S2y = fma(dy, q, t.S2y);
        t.S1 += i1tmp;
    }
    __device__ __forceinline__ void run(Tgt& t, const NearArgs&, const NearShared& S, int k0, int n) const {
#pragma unroll 4
        for (int k = k0; k < k0 + n; k++) {
            double2 p = S.sxy[k];
            double dx = VV_SUB(t.x, p.x), dy = VV_SUB(t.y, p.y);
            double d2 = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
            if (!(d2 * t.ie2 > t.lim)) hit(t, dx, dy, d2, S.sab[k].x);
        }
    }
    // segment_influence, :107-123
    __device__ __forceinline__ void segments(Tgt& t, const NearArgs& A, int sf, int sl) const {
        for (int k = sf; k < sl; k++) {
            int s = A.seg_perm[k];
            double dx = VV_SUB(t.x, A.srx[s]), dy = VV_SUB(t.y, A.sry[s]);
            double drabs2 = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
            double drabs = sqrt(drabs2);
            double exparg = -VV_MUL(drabs, t.ie);
            if (exparg < -8.) continue;
            double expres = exp(exparg);
            double dSx = -A.sdly[s], dSy = A.sdlx[s];
            t.S3x += dSx * expres; t.S3y += dSy * expres;
            t.S0 += (drabs * t.ie + 1) / drabs2 * (dx * dSx + dy * dSy) * expres;
            atomicAdd(&fric[s], t.ie2 * t.g * expres * sqrt(dSx * dSx + dSy * dSy));
        }
    }
    __device__ __forceinline__ void finish(Tgt& t, const NearArgs& A, int i, int) const {
        double S1 = t.S1;
        if ((sgn(S1) != sgn(t.g)) || (fabs(S1) < fabs(0.1 * t.g))) S1 = 0.1 * t.g;  // :41
        double k2 = t.ie / (re * S1);
        double vx = k2 * t.S2x, vy = k2 * t.S2y;
        double S0 = t.S0;
        if (S0 > kPi) S0 = kPi;
        double k3 = t.ie2 / (re * (k2Pi - S0));
        vx += k3 * t.S3x; vy += k3 * t.S3y;
        A.P.vx[i] += vx;
        A.P.vy[i] += vy;
    }
};

// ------------------------------------------------------------------------------------------ K3
// Merge bookkeeping ("timeline"): the reference processes particles in array order and a merge
// (MergeVortexes, MEpsilonFast.cpp:111-126) changes what LATER particles see. A tentative solution
// says, per particle q: init[q] (q merges when its turn comes), part[q] (with whom), (nx,ny,ng)[q]
// (its state afterwards) and absby[q] (index of the initiator that absorbed q, INT_MAX if none).
// A target i then sees source q as: skipped if absby[q] < i; post-merge state if init[q] && q < i;
// original state otherwise. EpsOp recomputes every particle's outcome under the tentative solution;
// the host iterates until the solution reproduces itself, which is the sequential result (the
// prefix of correct outcomes grows by at least one particle per round).
struct MergeState {
    int* absby;
    int* init;
    int* part;
    double *nx, *ny, *ng;
};
constexpr int kNoAbs = 0x7fffffff;

template <bool FINAL>
struct EpsOp {
    static constexpr bool kSegments = false;
    static constexpr bool kFilter = false;
    __device__ __forceinline__ double reach(double) const { return 0; }
    MergeState A_;      // assumed solution (absby == nullptr: no merges anywhere)
    MergeState B_;      // recomputed solution (decision mode only)
    const double* lcrit;   // per leaf merge_criteria_sq (NaN: never merge)
    const double* lrestr;  // per leaf eps_restriction
    const unsigned char* dyn;  // per particle: has a timeline entry in A_
    double* ie_out;
    int* changed;
    struct Tgt { double x, y, r1, r2; int i, i1, i2; };
    struct Part { double r1, r2; int i1, i2; };
    __device__ __forceinline__ Part part(const Tgt& t) const { return Part{t.r1, t.r2, t.i1, t.i2}; }
    // two smallest in (distance, index) order == the reference's first-seen-wins scan (:143-151)
    __device__ __forceinline__ void consider(Tgt& t, double d, int j) const {
        if (j == t.i1 || j == t.i2) return;  // already recorded (own-leaf seed, or a parked partial)
        if (d < t.r1 || (d == t.r1 && j < t.i1)) {
            t.r2 = t.r1; t.i2 = t.i1; t.r1 = d; t.i1 = j;
        } else if (d < t.r2 || (d == t.r2 && j < t.i2)) {
            t.r2 = d; t.i2 = j;
        }
    }
    __device__ __forceinline__ void combine(Tgt& t, const Part& p) const {
        if (p.i1 >= 0) consider(t, p.r1, p.i1);
        if (p.i2 >= 0) consider(t, p.r2, p.i2);
    }

    __device__ __forceinline__ bool init(Tgt& t, const NearArgs& A, int i, int leaf, bool inrange) const {
        t.r1 = t.r2 = DBL_MAX; t.i1 = t.i2 = -1; t.i = i; t.x = t.y = 0;
        if (!inrange) return false;
        if (A.P.g[i] == 0) return false;  // MEpsilonFast.cpp:51
        if (A_.absby) {
            if (FINAL) {
                if (!A_.init[i]) return false;
                t.x = A_.nx[i]; t.y = A_.ny[i];
                return true;
            }
            if (A_.absby[i] < i) {  // absorbed before its turn: g == 0 by then, _1_eps stays as it was
                if (A_.init[i]) atomicAdd(changed, 1);
                ie_out[i] = A.P.ie[i];
                return false;
            }
        }
        t.x = A.P.x[i]; t.y = A.P.y[i];
        return true;
    }
    // Seed the two-nearest search with a few neighbours from the target's own leaf, so that r2 is
    // already small when the stream starts and almost no source takes the cand() branch. Sources
    // met again in the stream are recognised by index in consider().
    __device__ __forceinline__ void seed(Tgt& t, const NearArgs& A, int leaf) const {
        const int f = max(A.L.first[leaf], t.i - 8), l = min(A.L.last[leaf], t.i + 9);
        for (int j = f; j < l; j++) {
            if (j == t.i || A.P.g[j] == 0) continue;
            double d;
            if (A_.absby && dyn[j]) d = __longlong_as_double(0x7ff8000000000000ll);
            else {
                double dx = VV_SUB(t.x, A.P.x[j]), dy = VV_SUB(t.y, A.P.y[j]);
                d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
            }
            cand(t, A, d, j);
        }
    }
    // A g == 0 source is parked at x = +inf (never a neighbour, :140); a source with a timeline entry
    // at x = NaN, which fails the `d > r2` test below and is re-read from the timeline in cand().
    static __device__ __forceinline__ double4 pack(const Particles& P, int j, const unsigned char* dyn) {
        double x = P.x[j];
        if (P.g[j] == 0) x = __longlong_as_double(0x7ff0000000000000ll);
        else if (dyn && dyn[j]) x = __longlong_as_double(0x7ff8000000000000ll);
        return make_double4(x, P.y[j], 0., 0.);
    }
    // state of source j as target i sees it; false = not a neighbour candidate
    __device__ __forceinline__ bool seen(int j, int i, double& sx, double& sy, double& sg) const {
        int ab = A_.absby[j];
        if (FINAL ? (ab <= i) : (ab < i)) return false;
        if (A_.init[j] && j < i) { sx = A_.nx[j]; sy = A_.ny[j]; sg = A_.ng[j]; }
        return sg != 0;
    }
    __device__ __forceinline__ void cand(Tgt& t, const NearArgs& A, double d, int j) const {
        if (j == t.i) return;  // :140
        if (isnan(d)) {
            double sx = A.P.x[j], sy = A.P.y[j], sg = A.P.g[j];
            if (!A_.absby || !seen(j, t.i, sx, sy, sg)) return;
            double dx = VV_SUB(t.x, sx), dy = VV_SUB(t.y, sy);
            d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
        }
        consider(t, d, j);
    }
    // common path: 5 FP64 + one compare; only a source at least as close as the current second
    // neighbour (or a parked NaN) takes the branch
    __device__ __forceinline__ void run(Tgt& t, const NearArgs& A, const NearShared& S, int k0, int n) const {
#pragma unroll 4
        for (int k = k0; k < k0 + n; k++) {
            double2 p = S.sxy[k];
            double dx = VV_SUB(t.x, p.x), dy = VV_SUB(t.y, p.y);
            double d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
            if (!(d > t.r2)) cand(t, A, d, S.sj[k]);
        }
    }
    __device__ __forceinline__ void segments(Tgt&, const NearArgs&, int, int) const {}
    __device__ __forceinline__ void finish(Tgt& t, const NearArgs& A, int i, int leaf) const {
        const double restr = lrestr ? lrestr[leaf] : 0.;
        double eps;
        if (t.i1 < 0) eps = DBL_MIN;           // :155
        else if (t.i2 < 0) eps = sqrt(t.r1);   // :157
        else {
            eps = sqrt(t.r2);
            const double crit = lcrit ? lcrit[leaf] : __longlong_as_double(0x7ff8000000000000ll);
            if (!FINAL && !isnan(crit)) {
                // neighbour states as seen (for the sign rule and the merged position)
                double x1 = A.P.x[t.i1], y1 = A.P.y[t.i1], g1 = A.P.g[t.i1];
                double x2 = 0, y2 = 0, g2 = A.P.g[t.i2];
                if (A_.absby) { seen(t.i1, i, x1, y1, g1); seen(t.i2, i, x2, y2, g2); }
                const double gi = A.P.g[i];
                if ((t.r1 < crit) || ((sgn(g1) == sgn(g2)) && (sgn(g1) != sgn(gi)))) {  // :162-166
                    // MergeVortexes(lv, lv1), :111-126
                    double nx = t.x, ny = t.y;
                    if (sgn(gi) == sgn(g1)) {
                        double r = 1. / VV_ADD(gi, g1);
                        nx = VV_MUL(VV_ADD(VV_MUL(t.x, gi), VV_MUL(x1, g1)), r);
                        ny = VV_MUL(VV_ADD(VV_MUL(t.y, gi), VV_MUL(y1, g1)), r);
                    } else if (fabs(gi) < fabs(g1)) { nx = x1; ny = y1; }
                    double ng = VV_ADD(gi, g1);
                    bool diff = true;
                    if (A_.absby && A_.init[i] && A_.part[i] == t.i1 &&
                        __double_as_longlong(A_.nx[i]) == __double_as_longlong(nx) &&
                        __double_as_longlong(A_.ny[i]) == __double_as_longlong(ny) &&
                        __double_as_longlong(A_.ng[i]) == __double_as_longlong(ng)) diff = false;
                    B_.init[i] = 1; B_.part[i] = t.i1; B_.nx[i] = nx; B_.ny[i] = ny; B_.ng[i] = ng;
                    atomicMin(&B_.absby[t.i1], i);
                    if (diff) atomicAdd(changed, 1);
                    return;  // its epsilon comes from the FINAL pass at the merged position
                }
            }
        }
        if (!FINAL && A_.absby && A_.init[i]) atomicAdd(changed, 1);  // assumed a merge that does not happen
        ie_out[i] = 1.0 / std_max(eps, restr);  // :52
    }
};

// write the converged merges back into the particle arrays
__global__ void k_merge_apply(int n, MergeState M, double* x, double* y, double* g, int* nmerged) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ini = M.init[i] != 0;
    if (ini) { x[i] = M.nx[i]; y[i] = M.ny[i]; g[i] = M.ng[i]; }
    if (M.absby[i] != kNoAbs) g[i] = 0;
    unsigned b = __ballot_sync(__activemask(), ini);
    if (ini && (threadIdx.x & 31) == (__ffs(b) - 1)) atomicAdd(nmerged, __popc(b));
}
__global__ void k_merge_clear(int n, MergeState M) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    M.absby[i] = kNoAbs; M.init[i] = 0; M.part[i] = -1;
}
__global__ void k_merge_dyn(int n, MergeState M, unsigned char* dyn) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dyn[i] = (M.init[i] != 0 || M.absby[i] != kNoAbs) ? 1 : 0;
}

// current bounding box and largest epsilon of every leaf's particles (after epsilon / merging)
__global__ void k_leaf_box(LeafDev L, int nleaves, Particles P, double* lbox) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    double x0 = DBL_MAX, x1 = -DBL_MAX, y0 = DBL_MAX, y1 = -DBL_MAX, em = 0;
    for (int i = L.first[l]; i < L.last[l]; i++) {
        double x = P.x[i], y = P.y[i];
        x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
        if (P.g[i] != 0) em = fmax(em, 1. / P.ie[i]);
    }
    double* b = lbox + 5ll * l;
    b[0] = x0; b[1] = x1; b[2] = y0; b[3] = y1; b[4] = em;
}

// Per-leaf wall parameters of CalcEpsilonFast (MEpsilonFast.cpp:26-47) with nearestBodySegment
// (:214-253). Only the few list entries whose leaf holds body segments do any work: pass 1 finds
// each target leaf's smallest squared distance to a segment of its near leaves, pass 2 resolves ties
// to the first segment in the reference's scan order (leaf order, then list order; strict `<`, :229),
// k_wall_finish falls back to the global scan with its skip-ahead (:237-251) and derives the criteria.
struct BodySegs {
    int nseg, nbody;
    const double *rx, *ry, *dlx, *dly;
    const int* bfirst;  // nbody + 1
};
template <int PASS>
__global__ void __launch_bounds__(256) k_wall_pass(LeafDev L, GroupLists G, Units U, int nunits,
                                                   const int* __restrict__ seg_perm, BodySegs B, u64* best_d,
                                                   u64* best_key) {
    const int u = blockIdx.x;
    if (u >= nunits) return;
    const int g = U.group[u];
    const int chunk = u - U.first[g];
    const long long e0 = G.ptr[g] + (long long)chunk * kUnitEntries;
    const long long e1 = min(G.ptr[g + 1], e0 + kUnitEntries);
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const int sl = G.leaf[e];
        const int sf = L.sfirst[sl], se = L.slast[sl];
        if (se <= sf) continue;
        for (u32 m = G.mask[e]; m; m &= m - 1) {
            const int l = g * kGroupLeaves + (__ffs(m) - 1);
            const double px = L.cx[l], py = L.cy[l];
            for (int k = sf; k < se; k++) {
                const int s = seg_perm[k];
                double dx = VV_SUB(px, B.rx[s]), dy = VV_SUB(py, B.ry[s]);
                u64 d = (u64)__double_as_longlong(VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy)));  // >= 0: bits are ordered
                if (PASS == 1) atomicMin(&best_d[l], d);
                else if (d == best_d[l]) atomicMin(&best_key[l], ((u64)(u32)sl << 32) | (u32)k);
            }
        }
    }
}
__global__ void k_wall_init(int nleaves, u64* best_d, u64* best_key) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nleaves) { best_d[l] = 0x7fefffffffffffffull /* DBL_MAX */; best_key[l] = ~0ull; }
}
__global__ void k_wall_finish(LeafDev L, int nleaves, const int* __restrict__ seg_perm, BodySegs B, int merge,
                              const u64* best_d, const u64* best_key, double* lcrit, double* lrestr, int* latt) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    const double px = L.cx[l], py = L.cy[l];
    int att = -1;
    // `drabs2 < res` with res starting at DBL_MAX (:229): a segment at distance DBL_MAX never wins
    if (best_key && best_key[l] != ~0ull && best_d[l] < 0x7fefffffffffffffull) att = seg_perm[(int)(best_key[l] & 0xffffffffull)];
    if (att < 0) {
        double res = DBL_MAX;
        for (int ib = 0; ib < B.nbody; ib++) {
            for (int s = B.bfirst[ib]; s < B.bfirst[ib + 1]; s++) {
                double dx = VV_SUB(px, B.rx[s]), dy = VV_SUB(py, B.ry[s]);
                double d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
                if (d < res) { res = d; att = s; }
                else s += 9;  // the reference's skip-ahead (:248)
            }
        }
    }
    double crit, restr = 0;
    if (merge) {
        crit = 0;
        if (att >= 0) {
            double dl2 = VV_ADD(VV_MUL(B.dlx[att], B.dlx[att]), VV_MUL(B.dly[att], B.dly[att]));
            double dx = VV_SUB(px, B.rx[att]), dy = VV_SUB(py, B.ry[att]);
            double dist = sqrt(VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy)));
            crit = VV_MUL(VV_MUL(0.16, dl2), VV_ADD(1., dist));  // :33
        }
    } else crit = __longlong_as_double(0x7ff8000000000000ll);
    if (att >= 0) {
        double dl = sqrt(VV_ADD(VV_MUL(B.dlx[att], B.dlx[att]), VV_MUL(B.dly[att], B.dly[att])));
        restr = VV_MUL(dl, (1.0 / 3.0));  // :45-47
    }
    lcrit[l] = crit; lrestr[l] = restr; latt[l] = att;
}

// MConvectiveFast::body_list_influence (MConvectiveFast.cpp:172-215) for every particle; only
// launched when a body has slip segments or a non-zero speed_slae.
struct BodyFull {
    int nseg, nbody;
    const double *rx, *ry, *cx, *cy, *dlx, *dly, *g, *ie;
    const int* slip;
    const int* bfirst;
    const double* bprop;  // per body 16 doubles (vvgpu_body as doubles, see vvgpu.cu)
};
__device__ __forceinline__ void cplx_mul(double ar, double ai, double br, double bi, double& cr, double& ci) {
    cr = ar * br - ai * bi; ci = ar * bi + ai * br;
}
__device__ __forceinline__ void cplx_div(double ar, double ai, double br, double bi, double& cr, double& ci) {
    double d = br * br + bi * bi;
    cr = (ar * br + ai * bi) / d; ci = (ai * br - ar * bi) / d;
}
// SegmentInfluence_linear_source, :440-457
__device__ __forceinline__ void linear_source(double px, double py, double rx, double ry, double dlx, double dly,
                                              double q1, double q2, double& ox, double& oy) {
    double z1r = px - (rx - dlx * 0.5), z1i = py - (ry - dly * 0.5);
    double z2r = px - (rx + dlx * 0.5), z2i = py - (ry + dly * 0.5);
    // conj(q2*z1 - q1*z2) / conj(dz)
    double ar = q2 * z1r - q1 * z2r, ai = -(q2 * z1i - q1 * z2i);
    double br, bi;
    cplx_div(ar, ai, dlx, -dly, br, bi);
    // log(conj(z2)/conj(z1))
    double cr, ci;
    cplx_div(z2r, -z2i, z1r, -z1i, cr, ci);
    double lr = 0.5 * log(cr * cr + ci * ci), li = atan2(ci, cr);
    double mr, mi;
    cplx_mul(br, bi, lr, li, mr, mi);
    cplx_div((q2 - q1) - mr, -mi, dlx, -dly, ox, oy);
}
__global__ void k_body_influence(int n, Particles P, BodyFull B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (P.g[i] == 0) return;
    const double px = P.x[i], py = P.y[i];
    double rx = 0, ry = 0;
    for (int ib = 0; ib < B.nbody; ib++) {
        const double* bp = B.bprop + 16 * ib;
        const int f = B.bfirst[ib], e = B.bfirst[ib + 1];
        if (bp[13] != 0) {  // any slip segment (TBody::get_slip)
            for (int s = f; s < e; s++) {
                if (!B.slip[s]) continue;
                double dx = px - B.rx[s], dy = py - B.ry[s];
                double ee = 1. / B.ie[s];
                double k = B.g[s] / (dx * dx + dy * dy + ee * ee);
                rx += -dy * k; ry += dx * k;
            }
        }
        const double sx = bp[9], sy = bp[10], so = bp[11], ax = bp[0], ay = bp[1];
        if (!(fabs(sx) + fabs(sy) + fabs(so) < 1E-10)) {
            for (int s = f; s < e; s++) {
                double dx = px - B.rx[s], dy = py - B.ry[s];
                double drabs2 = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
                double dlx = B.dlx[s], dly = B.dly[s];
                if (drabs2 < VV_ADD(VV_MUL(dlx, dlx), VV_MUL(dly, dly))) {
                    double ux = B.cx[s] - ax, uy = B.cy[s] - ay;
                    double v1x = sx - so * uy, v1y = sy + so * ux;
                    double g1 = -(v1x * dlx + v1y * dly), q1 = -(-v1y * dlx + v1x * dly);
                    double wx = B.cx[s] + dlx - ax, wy = B.cy[s] + dly - ay;
                    double v2x = sx - so * wy, v2y = sy + so * wx;
                    double g2 = -(v2x * dlx + v2y * dly), q2 = -(-v2y * dlx + v2x * dly);
                    double ix, iy;
                    linear_source(px, py, B.rx[s], B.ry[s], dlx, dly, g1, g2, ix, iy);
                    rx += -iy; ry += ix;
                    linear_source(px, py, B.rx[s], B.ry[s], dlx, dly, q1, q2, ix, iy);
                    rx += ix; ry += iy;
                } else {
                    double ux = B.rx[s] - ax, uy = B.ry[s] - ay;
                    double vx = sx - so * uy, vy = sy + so * ux;
                    double gg = -(vx * dlx + vy * dly), q = -(-vy * dlx + vx * dly);
                    double r = 1. / drabs2;
                    rx += (dx * q - dy * gg) * r;
                    ry += (dy * q + dx * gg) * r;
                }
            }
        }
    }
    P.vx[i] += rx * k1_2Pi;
    P.vy[i] += ry * k1_2Pi;
}

}  // namespace vv
