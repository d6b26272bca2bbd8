// K3 / K5 — the two near-field passes with short, pruned source lists (K4, the convective pass, has its
// own kernel in vvgpu_conv.cuh and shares the work decomposition, the operators' finish/combine code and
// k_near_finalize with this file):
//   EpsOp   MEpsilonFast::epsv + merge decision        (libvvhd/src/MEpsilonFast.cpp:128-173)
//   DiffOp  MDiffusiveFast::process_vort_list          (libvvhd/src/MDiffusiveFast.cpp:8-48,93-123)
//   ConvOp  per-particle assembly of MConvectiveFast::process_all_lists (:76-86)
//
// Layout ("leaf-warp"): a CTA owns one work unit = one group of 32 consecutive leaves x <= 512
// entries of the group's near list. Its warps pull TARGET LEAVES from a shared counter; a warp
// scans the unit's entry table, keeps the source leaves whose mask bit names its leaf and whose box
// is within the leaf's exact reach, expands them into a flat list of source particle indices in
// shared memory and streams that list: the warp is split into nt x m sub-lanes (m = 32 / nt), each
// target's sources are dealt round-robin to its m sub-lanes and the partial states are merged with a
// segmented shuffle reduction.
// A leaf normally holds < 16 particles (Tree_MaxListSize); leaves made by the min-node-size rule
// can hold more and are processed in chunks of 15 targets.
#pragma once
#include <type_traits>
#include "vvgpu_lists.cuh"

namespace vv {

#ifndef VV_LW_WARPS
#define VV_LW_WARPS 8
#endif
#ifndef VV_EPS_MINB
#define VV_EPS_MINB 3
#endif
constexpr int kLwWarps = VV_LW_WARPS;
constexpr int kLwThreads = kLwWarps * 32;
constexpr int kMaxT = 15;           // targets per pass of a warp
constexpr int kIdxCap = 1024;       // flat source indices buffered per warp
constexpr u32 kFullMask = 0xffffffffu;

struct Particles {
    double *x, *y, *g, *vx, *vy, *ie;
};

// Work units = the chunks the traversal filled (vvgpu_lists.cuh). A group whose list is longer than one
// chunk has several units (a few fringe leaves see tens of thousands of near leaves);
// they park their partial per-target state in `scratch` and k_near_finalize combines them in unit order.
struct Units {
    const int* group;        // unit -> group
    const long long* base;   // unit -> first entry in the pool
    const int* count;        // unit -> entries
    const int* first;        // group -> first unit
    const int* num;          // group -> number of units
    const u32* sbase;        // group -> first scratch slot
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
    const int* cta_unit;             // CTA table (k_cta_table), or nullptr: blockIdx -> unit by the formula with tsplit
    const unsigned short* cta_part;  // first | one-past-last << 8 target leaf of the group
    int cta_heavy;                   // CTAs of the table's first class
    int tsplit;  // CTAs per unit: each takes 32 / tsplit of the group's target leaves against the whole entry table. What a
                 // target sums, and in which order, does not depend on it: few units (a small problem, one rank of many)
                 // still fill the SMs, with identical bits
    // segments (diffusive / epsilon wall terms)
    const int* seg_perm;
    const double *srx, *sry, *sdlx, *sdly;
};

// CTAs start in blockIdx order; the units are in DFS order of their groups, and the expensive ones — the fringe groups,
// whose sparse leaves see long lists and large epsilons — sit at BOTH ends of that order. Starting from both ends
// alternately puts them first instead of into the kernel's tail (with the work of a step split over 8 GPUs one
// late fringe unit was a third of the kernel's duration).
__device__ __forceinline__ int unit_order(int b, int n) { return (b & 1) ? (n - 1 - (b >> 1)) : (b >> 1); }
// CTA -> (unit slot, first and one-past-last target leaf of the group it serves); both_ends: the unit_order heuristic
struct UnitPart {
    int b, lt0, lt1;
    __device__ __forceinline__ UnitPart(const NearArgs& A, bool both_ends) {
        if (A.cta_unit) {
            int pos = (int)blockIdx.x;
            if (both_ends && pos >= A.cta_heavy) pos = A.cta_heavy + unit_order(pos - A.cta_heavy, (int)gridDim.x - A.cta_heavy);
            b = A.cta_unit[pos];
            const int p = A.cta_part[pos];
            lt0 = p & 0xff; lt1 = p >> 8;
        } else {
            const int q = (int)blockIdx.x / A.tsplit;
            const int part = (int)blockIdx.x - q * A.tsplit, per = kGroupLeaves / A.tsplit;
            b = both_ends ? unit_order(q, (int)gridDim.x / A.tsplit) : q;
            lt0 = part * per; lt1 = lt0 + per;
        }
    }
};

template <class Op>
struct LwWarpT {
    static constexpr int kCap = kIdxCap / 2;
    static constexpr int kFlush = kCap / 2;
    int idx[kCap];                                             // flat list of source particle indices
    double tsave[kMaxT][16];                                   // target states handed to the sub-lanes
    int tpart[kMaxT + 1];                                      // particle index of each target slot
};
template <class Op>
struct LwSharedT {
    int4 ent[kUnitEntries];      // first particle, count, first segment, segment count of the entry's leaf
    u32 emk[kUnitEntries];       // target-leaf mask
    double ebox[Op::kFilter ? kUnitEntries : 1][4];   // source-leaf box (ops with kFilter)
    unsigned char edirty[Op::kDirty ? kUnitEntries : 1];   // source leaf holds a particle whose merge entry changed last round
    int bounds[kGroupLeaves + 1];
    int next;                    // next target leaf of the group to hand out
    int anyseg;
    LwWarpT<Op> w[kLwWarps];
};

template <class T>
__device__ __forceinline__ T shfl_down_struct(const T& v, int o) {
    static_assert(sizeof(T) % 4 == 0, "word-sized parts only");
    T r;
    const u32* a = reinterpret_cast<const u32*>(&v);
    u32* b = reinterpret_cast<u32*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 4); k++) b[k] = __shfl_down_sync(kFullMask, a[k], o);
    return r;
}

template <class Op>
__global__ void __launch_bounds__(kLwThreads, Op::kMinBlocks) k_near(NearArgs A, Op op) {
    extern __shared__ __align__(16) unsigned char near_smem[];
    LwSharedT<Op>& S = *reinterpret_cast<LwSharedT<Op>*>(near_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const UnitPart up(A, true);
    const int u = A.u0 + up.b;
    const int g = A.U.group[u];
    const int chunk = u - A.U.first[g];
    const bool multi = A.U.num[g] > 1;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, A.nleaves - l0);
    if (up.lt0 >= nl) return;
    if constexpr (Op::kDirty) {
        // a later round of the merge fixed point, and nothing this unit's targets can see changed: they keep their entries
        if (op.unit_dirty && (multi ? !op.group_dirty[g] : !op.unit_dirty[u])) return;
    }
    const int lt_end = min(nl, up.lt1);
    if (tid <= nl) S.bounds[tid] = (tid < nl) ? A.L.first[l0 + tid] : A.L.last[l0 + nl - 1];
    if (tid == 0) { S.next = up.lt0; S.anyseg = 0; }
    const long long e0 = A.U.base[u];
    const int ne = A.U.count[u];
    __syncthreads();
    // ---- the unit's entry table, once per CTA
    for (int e = tid; e < ne; e += kLwThreads) {
        const int sl = A.G.leaf[e0 + e];
        const int f = A.L.first[sl];
        int4 en = make_int4(f, A.L.last[sl] - f, 0, 0);
        if (Op::kSegments && A.nseg > 0) {
            en.z = A.L.sfirst[sl]; en.w = A.L.slast[sl] - en.z;
            if (en.w > 0) S.anyseg = 1;
        }
        S.ent[e] = en;
        S.emk[e] = A.G.mask[e0 + e];
        if constexpr (Op::kFilter) {
            const double* b = A.lbox + 5ll * sl;
            S.ebox[e][0] = b[0]; S.ebox[e][1] = b[1]; S.ebox[e][2] = b[2]; S.ebox[e][3] = b[3];
        }
        if constexpr (Op::kDirty) {
            static_assert(!Op::kSegments, "ent.z carries the source leaf for the dirty test");
            S.ent[e].z = sl;
            S.edirty[e] = op.leaf_dirty ? op.leaf_dirty[sl] : 1;
        }
    }
    __syncthreads();
    const int t0 = S.bounds[0], t1 = S.bounds[nl];
    LwWarpT<Op>& W = S.w[warp];
    constexpr int kFlush = LwWarpT<Op>::kFlush;
    constexpr int kPiece = kFlush / 32;   // sources appended per entry per round: 32 x kPiece fills half a buffer
    typename Op::Part* scratch = (typename Op::Part*)A.scratch;
    const size_t sbase = multi ? ((size_t)A.U.sbase[g] + (size_t)chunk * (t1 - t0)) : 0;

    for (;;) {
        int lt = 0;
        if (lane == 0) lt = atomicAdd(&S.next, 1);
        lt = __shfl_sync(kFullMask, lt, 0);
        if (lt >= lt_end) break;
        const int leaf = l0 + lt;
        const int pf = S.bounds[lt], pl = S.bounds[lt + 1];
        for (int tb = pf; tb < pl; tb += kMaxT) {
            const int np = min(kMaxT, pl - tb);
            typename Op::Tgt tg;
            const int i = tb + lane;
            const bool live = op.init(tg, A, i, leaf, lane < np);
            if (live) op.seed(tg, A, leaf);
            const u32 lm = __ballot_sync(kFullMask, live);
            const int nt = __popc(lm);
            if (nt == 0) continue;
            const int slot = __popc(lm & lanemask_lt());
            // box of the live targets and their largest squared reach (exact pruning of source leaves)
            double bx0 = DBL_MAX, bx1 = -DBL_MAX, by0 = DBL_MAX, by1 = -DBL_MAX, R2 = 0;
            if constexpr (Op::kFilter) {
                if (live) { bx0 = bx1 = tg.x; by0 = by1 = tg.y; R2 = op.reach2(tg); }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    bx0 = fmin(bx0, __shfl_xor_sync(kFullMask, bx0, o)); bx1 = fmax(bx1, __shfl_xor_sync(kFullMask, bx1, o));
                    by0 = fmin(by0, __shfl_xor_sync(kFullMask, by0, o)); by1 = fmax(by1, __shfl_xor_sync(kFullMask, by1, o));
                    R2 = fmax(R2, __shfl_xor_sync(kFullMask, R2, o));
                }
                R2 *= 1.000000001;  // the skip stays strictly conservative against rounding in the gap
            }
            if constexpr (Op::kDirty) {
                // A later round of the merge fixed point. These targets decide as in the last round — whose outcome
                // the new solution was seeded with — unless a particle whose entry changed lies within their reach at
                // any of the positions it was or is seen at (original, old merged, new merged: the leaf's dirty box).
                // Their own leaf is in the list at distance 0, so a change among the seeds always recomputes.
                if (op.leaf_dirty && !multi) {
                    bool need = false;
                    for (int eb0 = 0; eb0 < ne && !need; eb0 += 32) {
                        const int e = eb0 + lane;
                        bool hit = false;
                        if (e < ne && ((S.emk[e] >> lt) & 1u) && S.edirty[e]) {
                            const double* b = op.leaf_dbox + 4ll * S.ent[e].z;
                            const double gx = fmax(0., fmax(b[0] - bx1, bx0 - b[1]));
                            const double gy = fmax(0., fmax(b[2] - by1, by0 - b[3]));
                            hit = !(gx * gx + gy * gy > R2);
                        }
                        need = __any_sync(kFullMask, hit);
                    }
                    if (!need) continue;
                    if (lane == 0) atomicAdd(op.changed + 1, 1);   // target batches recomputed in this round (diagnostic)
                }
            }
            // sub-lane layout: nt x m lanes, m = 32 / nt
            static_assert(sizeof(typename Op::Tgt) <= sizeof(W.tsave[0]), "tsave slot too small");
            if (live) { *reinterpret_cast<typename Op::Tgt*>(W.tsave[slot]) = tg; W.tpart[slot] = i; }
            const int m = 32 / nt;
            const int myslot = lane / m;
            const int sub = lane - myslot * m;
            const bool active = myslot < nt;
            __syncwarp();
            typename Op::Tgt my = *reinterpret_cast<const typename Op::Tgt*>(W.tsave[active ? myslot : 0]);
            // wall segments of the near leaves (MDiffusiveFast.cpp:26-34): one sub-lane per target
            if (Op::kSegments && S.anyseg && active && sub == 0) {
                for (int e = 0; e < ne; e++) {
                    if (!((S.emk[e] >> lt) & 1u)) continue;
                    const int4 en = S.ent[e];
                    if (en.w > 0) op.segments(my, A, en.z, en.z + en.w);
                }
            }
            // ---- scan the entry table, expand the kept leaves into source indices, stream them
            int fill = 0, eb = 0, cnt = 0, f = 0;
            for (;;) {
                // refill: until the buffer is worth draining or the table is exhausted
                bool pending = __any_sync(kFullMask, cnt > 0);
                while (fill < kFlush && (pending || eb < ne)) {
                    if (!pending) {
                        const int e = eb + lane;
                        eb += 32;
                        if (e < ne && ((S.emk[e] >> lt) & 1u)) {
                            const int4 en = S.ent[e];
                            f = en.x; cnt = en.y;
                            if constexpr (Op::kFilter) if (cnt) {
                                const double* b = S.ebox[e];
                                const double gx = fmax(0., fmax(b[0] - bx1, bx0 - b[1]));
                                const double gy = fmax(0., fmax(b[2] - by1, by0 - b[3]));
                                if (gx * gx + gy * gy > R2) cnt = 0;
                            }
                        }
                        pending = __any_sync(kFullMask, cnt > 0);
                        continue;
                    }
                    const int c = min(cnt, kPiece);
                    int inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFullMask, inc, o);
                        if (lane >= o) inc += t;
                    }
                    const int tot = __shfl_sync(kFullMask, inc, 31);
                    int* dst = W.idx + fill + inc - c;
                    for (int k = 0; k < c; k++) dst[k] = f + k;
                    fill += tot; cnt -= c; f += c;
                    pending = __any_sync(kFullMask, cnt > 0);
                }
                __syncwarp();
                const bool final = !pending && eb >= ne;
                if (active) {   // four loads in flight per lane before the first is used
                    for (int k = sub; k < fill; k += 4 * m) {
                        int j[4];
                        typename Op::Src v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int kk = k + q * m;
                            j[q] = (kk < fill) ? W.idx[kk] : -1;
                        }
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = (j[q] >= 0) ? Op::fetch(A, j[q]) : Op::none();
#pragma unroll
                        for (int q = 0; q < 4; q++) op.use(my, A, v[q], j[q]);
                    }
                }
                __syncwarp();
                fill = 0;
                if (final) break;
            }
            // merge the m sub-lane states of every target into its first sub-lane
            for (int o = 1; o < m; o <<= 1) {
                const typename Op::Part q = shfl_down_struct(op.part(my), o);
                if (active && sub + o < m) op.combine(my, q);
            }
            if (active && sub == 0) {
                const int ip = W.tpart[myslot];
                if (multi) scratch[sbase + (ip - t0)] = op.part(my);
                else op.finish(my, A, ip, leaf);
            }
            __syncwarp();
        }
    }
}

// combine the parked partial states of a multi-unit group, in unit order, and finish
template <class Op>
__global__ void __launch_bounds__(256) k_near_finalize(NearArgs A, Op op, Shard sh, int ngmine) {
    if ((int)blockIdx.x >= ngmine) return;
    const int g = sh.group(blockIdx.x);
    const int nu = A.U.num[g];
    if (nu <= 1) return;
    if (op.skip_group(g)) return;
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
        // combined in unit order; the loads of a batch do not wait for the combines before them (a fringe group parks
        // ~100 partial states per target: one dependent L2 round trip each was the whole duration of this kernel)
        constexpr int kBatch = 8;
        const typename Op::Part* sp = scratch + (i - t0);
        const size_t stride = (size_t)(t1 - t0);
        for (int c = 0; c < nu; c += kBatch) {
            typename Op::Part pb[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; k++)
                if (c + k < nu) pb[k] = sp[(size_t)(c + k) * stride];
#pragma unroll
            for (int k = 0; k < kBatch; k++)
                if (c + k < nu) op.combine(tg, pb[k]);
        }
        op.finish(tg, A, i, l0 + lo);
    }
}

// scratch slots per group: multi-unit groups park one partial state per (unit, particle)
__global__ void k_unit_slots(LeafDev L, int nleaves, int ngroups, const int* __restrict__ unum, u32* nslots) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const int nu = unum[g];
    int l0 = g * kGroupLeaves, nl = min(kGroupLeaves, nleaves - l0);
    u32 T = (u32)(L.last[l0 + nl - 1] - L.first[l0]);
    nslots[g] = nu > 1 ? (u32)nu * T : 0;
}

// per-phase packed source records (one 32-byte load per staged source instead of four scattered ones)
template <class Op>
__global__ void k_pack_src(int n, Particles P, const unsigned char* dyn, double4* out, const double4* tl = nullptr) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = Op::pack(P, j, dyn, tl);
    else if (j == n) out[j] = make_double4(0., 0., 0., 1.);   // dummy record (g = 0): pads K4's index lists
}

// ------------------------------------------------------------------------------------------ K4
struct ConvOp {
    static constexpr bool kSegments = false;
    static constexpr bool kFilter = false;
    static constexpr bool kDirty = false;
    double inf_vx, inf_vy, eps2_div_srcg;
    const double* taylor;  // 4 per leaf
    const double* sinks;   // (x,y,g) triples
    int nsink;
    struct Tgt { double x, y, rx, ry; };
    struct Part { double rx, ry; };
    __device__ __forceinline__ Part part(const Tgt& t) const { return Part{t.rx, t.ry}; }
    __device__ __forceinline__ void combine(Tgt& t, const Part& p) const { t.rx += p.rx; t.ry += p.ry; }
    __device__ __forceinline__ void seed(Tgt&, const NearArgs&, int) const {}
    __device__ __forceinline__ bool skip_group(int) const { return false; }

    __device__ __forceinline__ bool init(Tgt& t, const NearArgs& A, int i, int leaf, bool inrange) const {
        t.rx = t.ry = 0; t.x = t.y = 0;
        if (!inrange) return false;
        if (A.P.g[i] == 0) return false;  // `if (!lobj->g) continue`, MConvectiveFast.cpp:78
        t.x = A.P.x[i]; t.y = A.P.y[i];
        return true;
    }
    // packed source record: eps^2 = sqr(1./_1_eps) of the SOURCE (:119); a g==0 source is skipped by
    // the reference (:130) and is packed as g = 0, eps^2 = 1 so that it adds exactly nothing
    static __device__ __forceinline__ double4 pack(const Particles& P, int j, const unsigned char*, const double4*) {
        double g = P.g[j];
        double e = 1. / P.ie[j];
        // _1_eps == 0 (a list that never went through epsilon): the reference adds g / (dr^2 + inf) = 0; packed like a
        // g == 0 source, because the reciprocal's Newton step would turn the infinite denominator into a NaN
        return (g == 0 || isinf(e)) ? make_double4(P.x[j], P.y[j], 0., 1.) : make_double4(P.x[j], P.y[j], g, e * e);
    }
    // (the pair sum itself lives in vvgpu_conv.cuh)
    __device__ __forceinline__ void take(Tgt& t, double vx, double vy) const { t.rx = vx; t.ry = vy; }
    __device__ __forceinline__ void segments(Tgt&, const NearArgs&, int, int) const {}
    __device__ __forceinline__ void finish(Tgt& t, const NearArgs& A, int i, int leaf) const {
        double vx = inf_vx + t.rx * k1_2Pi, vy = inf_vy + t.ry * k1_2Pi;
        if (nsink) {  // sink_list_influence, :153-170
            double sx = 0, sy = 0;
            for (int k = 0; k < nsink; k++) {
                double dx = t.x - sinks[3 * k], dy = t.y - sinks[3 * k + 1], sg = sinks[3 * k + 2];
                double q = sg / (dx * dx + dy * dy + eps2_div_srcg * sink_abs(sg));
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
    static constexpr bool kDirty = false;
    double re;
    double* fric;  // per segment, atomically accumulated (MDiffusiveFast.cpp:121-122)
    struct Tgt { double x, y, ie, ie2, lim, g, S1, S2x, S2y, S0, S3x, S3y; bool pos; };
    struct Part { double S1, S2x, S2y, S0, S3x, S3y; };
    __device__ __forceinline__ Part part(const Tgt& t) const { return Part{t.S1, t.S2x, t.S2y, t.S0, t.S3x, t.S3y}; }
    __device__ __forceinline__ void combine(Tgt& t, const Part& p) const {
        t.S1 += p.S1; t.S2x += p.S2x; t.S2y += p.S2y; t.S0 += p.S0; t.S3x += p.S3x; t.S3y += p.S3y;
    }
    __device__ __forceinline__ void seed(Tgt&, const NearArgs&, int) const {}
    __device__ __forceinline__ bool skip_group(int) const { return false; }

    __device__ __forceinline__ bool init(Tgt& t, const NearArgs& A, int i, int leaf, bool inrange) const {
        t.S1 = t.S2x = t.S2y = t.S0 = t.S3x = t.S3y = 0;
        t.x = t.y = t.ie = t.ie2 = t.g = 0; t.pos = false; t.lim = 64.0001;
        if (!inrange) return false;
        double g = A.P.g[i];
        if (g == 0) return false;
        t.x = A.P.x[i]; t.y = A.P.y[i]; t.g = g; t.ie = A.P.ie[i]; t.ie2 = t.ie * t.ie; t.pos = g > 0;
        return true;
    }
    // (source views, the examine / evaluate loop and vortex_influence itself live in vvgpu_diff.cuh)
    // squared reach of a target: sources farther than 8 eps never contribute (:101)
    __device__ __forceinline__ double reach2(const Tgt& t) const { double r = 8.000008 / t.ie; return r * r; }
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

// MergeVortexes(lv, lv1), MEpsilonFast.cpp:111-126: the state of initiator i after it merged with i1, i1 taken as i sees
// it under the assumed solution A (post-merge state if i1 merged before i's turn). ONE function for the decision
// kernel and for the ranks that only receive (i, i1) from the owner of i: the bits must be the same everywhere.
__device__ __forceinline__ void merged_state(const Particles& P, const MergeState& A, int i, int i1, double& nx, double& ny,
                                             double& ng) {
    double x1 = P.x[i1], y1 = P.y[i1], g1 = P.g[i1];
    if (A.absby && !(A.absby[i1] < i) && A.init[i1] && i1 < i) { x1 = A.nx[i1]; y1 = A.ny[i1]; g1 = A.ng[i1]; }
    const double tx = P.x[i], ty = P.y[i], gi = P.g[i];
    nx = tx; ny = ty;
    if (sgn(gi) == sgn(g1)) {
        const double r = 1. / VV_ADD(gi, g1);
        nx = VV_MUL(VV_ADD(VV_MUL(tx, gi), VV_MUL(x1, g1)), r);
        ny = VV_MUL(VV_ADD(VV_MUL(ty, gi), VV_MUL(y1, g1)), r);
    } else if (fabs(gi) < fabs(g1)) { nx = x1; ny = y1; }
    ng = VV_ADD(gi, g1);
}
// a merge solution's remaining columns from its `part` column (the other ranks' decisions arrive as (i, part[i]) only)
__global__ void k_merge_fill(int n, Particles P, MergeState A, MergeState B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int i1 = B.part[i];
    if (i1 < 0) { B.init[i] = 0; return; }
    double nx, ny, ng;
    merged_state(P, A, i, i1, nx, ny, ng);
    B.init[i] = 1; B.nx[i] = nx; B.ny[i] = ny; B.ng[i] = ng;
    atomicMin(&B.absby[i1], i);
}

// TL: the round runs against a tentative solution with a compact timeline (every round but the first): sources are
// fetched as 32-byte records that carry the pre-filter of their timeline; without, as 16 bytes (x, y).
template <bool FINAL, bool TL = false>
struct EpsOp {
    static constexpr bool kSegments = false;
    // a source leaf whose box is farther from the leaf's targets than the largest seeded second-neighbour
    // distance cannot hold a closer neighbour of any of them: exact pruning
    static constexpr bool kFilter = true;
    static constexpr int kMinBlocks = VV_EPS_MINB;
    static constexpr bool kDirty = !FINAL;
    MergeState A_;      // assumed solution (absby == nullptr: no merges anywhere)
    MergeState B_;      // recomputed solution (decision mode only)
    const double* lcrit;   // per leaf merge_criteria_sq (NaN: never merge)
    const double* lrestr;  // per leaf eps_restriction
    const unsigned char* dyn;  // per particle: has a timeline entry in A_
    double* ie_out;
    int* changed;
    const unsigned char* leaf_dirty = nullptr;   // per leaf (decision rounds after the first): holds a changed entry
    const double* leaf_dbox = nullptr;           // per leaf 4: box of every position its changed particles were or are seen at
    const unsigned char* unit_dirty = nullptr;   // per work unit: its table names a dirty leaf (k_unit_dirty)
    const unsigned char* group_dirty = nullptr;  // per group: one of its units does (a group of several units is redone as a whole)
    const double4* tl = nullptr;                 // per particle with a timeline entry in A_, two records: (x, y, g, nx), (ny, ng, absby | init, -)
    __device__ __forceinline__ bool skip_group(int g) const { return kDirty && group_dirty && !group_dirty[g]; }
    struct Tgt { double x, y, r1, r2; int i, i1, i2, pad_; };
    __device__ __forceinline__ double reach2(const Tgt& t) const { return t.r2; }
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
                if (B_.init) { B_.init[i] = 0; B_.part[i] = -1; }
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
    // Packed source (x, y, z, w). A g == 0 source is parked at x = +inf (never a neighbour, :140). A source with a
    // timeline entry (w = 1) is seen by a target at its original or at its merged position, delta apart: it can only
    // be a neighbour candidate if d(original) <= (sqrt(r2) + delta)^2 <= 2 r2 + 2 delta^2, and z = 2 delta^2 makes that
    // one FMA in use(); what passes is resolved from the timeline in cand(). (Without the compact timeline such a
    // source sits at x = NaN, which fails every `d > r2` test.)
    static __device__ __forceinline__ double4 pack(const Particles& P, int j, const unsigned char* dyn, const double4* tl) {
        double x = P.x[j], z = 0., w = 0.;
        if (P.g[j] == 0) x = __longlong_as_double(0x7ff0000000000000ll);
        else if (dyn && dyn[j]) {
            if (tl) {
                const double4 a = tl[2ll * j], b = tl[2ll * j + 1];
                if (__double2hiint(b.z)) {
                    const double ex = a.w - a.x, ey = b.x - a.y;
                    z = 2.000000002 * (ex * ex + ey * ey);
                }
                w = 1.;
            } else x = __longlong_as_double(0x7ff8000000000000ll);
        }
        return make_double4(x, P.y[j], z, w);
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
            if (!A_.absby) return;
            double sx, sy, sg;
            if (tl) {
                // every pair with such a source comes here, near or not: its timeline in two 32-byte records instead of
                // nine scattered columns
                const double4 a = tl[2ll * j], b = tl[2ll * j + 1];
                const int ab = __double2loint(b.z), in = __double2hiint(b.z);
                if (FINAL ? (ab <= t.i) : (ab < t.i)) return;
                sx = a.x; sy = a.y; sg = a.z;
                if (in && j < t.i) { sx = a.w; sy = b.x; sg = b.y; }
                if (sg == 0) return;
            } else {
                sx = A.P.x[j]; sy = A.P.y[j]; sg = A.P.g[j];
                if (!seen(j, t.i, sx, sy, sg)) return;
            }
            double dx = VV_SUB(t.x, sx), dy = VV_SUB(t.y, sy);
            d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
        }
        consider(t, d, j);
    }
    // common path: 5 FP64 (+ 1 FMA in rounds with a timeline) + one compare; only a source at least as close as the
    // current second neighbour, or a timeline source that might be, takes the branch
    typedef typename std::conditional<TL, double4, double2>::type Src;
    static __device__ __forceinline__ Src none() {
        if constexpr (TL) return make_double4(__longlong_as_double(0x7ff0000000000000ll), 0., 0., 0.);
        else return make_double2(__longlong_as_double(0x7ff0000000000000ll), 0.);
    }
    static __device__ __forceinline__ Src fetch(const NearArgs& A, int j) {
        if constexpr (TL) return A.src4[j];
        else return *reinterpret_cast<const double2*>(A.src4 + j);
    }
    __device__ __forceinline__ void use(Tgt& t, const NearArgs& A, const Src& p, int j) const {
        double dx = VV_SUB(t.x, p.x), dy = VV_SUB(t.y, p.y);
        double d = VV_ADD(VV_MUL(dx, dx), VV_MUL(dy, dy));
        if constexpr (TL) {
            if (!(d > fma(2.000000002, t.r2, p.z))) {
                if (p.w != 0.) cand(t, A, __longlong_as_double(0x7ff8000000000000ll), j);
                else if (!(d > t.r2)) cand(t, A, d, j);
            }
        } else if (!(d > t.r2)) cand(t, A, d, j);
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
                    double nx, ny, ng;
                    merged_state(A.P, A_, i, t.i1, nx, ny, ng);
                    bool diff = true;
                    if (A_.absby && A_.init[i] && A_.part[i] == t.i1 &&
                        __double_as_longlong(A_.nx[i]) == __double_as_longlong(nx) &&
                        __double_as_longlong(A_.ny[i]) == __double_as_longlong(ny) &&
                        __double_as_longlong(A_.ng[i]) == __double_as_longlong(ng)) diff = false;
                    B_.init[i] = 1; B_.part[i] = t.i1; B_.nx[i] = nx; B_.ny[i] = ny; B_.ng[i] = ng;
                    if (diff) atomicAdd(changed, 1);   // (absorbed-by is rebuilt from (init, part) after the round)
                    return;  // its epsilon comes from the FINAL pass at the merged position
                }
            }
        }
        if (!FINAL && A_.absby && A_.init[i]) atomicAdd(changed, 1);  // assumed a merge that does not happen
        if (!FINAL && B_.init) { B_.init[i] = 0; B_.part[i] = -1; }   // (the new solution may have been seeded with the old one)
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
// the next round's solution starts as a copy of this round's: leaves that are not recomputed keep their entries
__global__ void k_merge_copy(int n, MergeState A, MergeState B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    B.init[i] = A.init[i]; B.part[i] = A.part[i]; B.nx[i] = A.nx[i]; B.ny[i] = A.ny[i]; B.ng[i] = A.ng[i];
}
// per leaf: does it hold a particle whose entry differs between the old solution (A, absent before the first round)
// and the new one (B)? Targets that see no such leaf would decide as before. One thread per PARTICLE (coalesced reads of
// the solution's columns); the few changed ones find their leaf and widen its dirty box (min / max: order-independent).
__device__ __forceinline__ void atomic_min_f64(double* a, double v) {
    unsigned long long old = *reinterpret_cast<unsigned long long*>(a);
    while (v < __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(reinterpret_cast<unsigned long long*>(a), assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__device__ __forceinline__ void atomic_max_f64(double* a, double v) {
    unsigned long long old = *reinterpret_cast<unsigned long long*>(a);
    while (v > __longlong_as_double((long long)old)) {
        const unsigned long long assumed = old;
        old = atomicCAS(reinterpret_cast<unsigned long long*>(a), assumed, (unsigned long long)__double_as_longlong(v));
        if (old == assumed) break;
    }
}
__global__ void k_leaf_dirty_clear(int nleaves, unsigned char* dirty, double* dbox) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    dirty[l] = 0;
    double* b = dbox + 4ll * l;
    b[0] = DBL_MAX; b[1] = -DBL_MAX; b[2] = DBL_MAX; b[3] = -DBL_MAX;
}
__global__ void k_leaf_dirty(LeafDev L, int nleaves, int n, Particles P, MergeState A, MergeState B, unsigned char* dirty, double* dbox) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bi = B.init[i], ba = B.absby[i];
    const int ai = A.absby ? A.init[i] : 0, aa = A.absby ? A.absby[i] : kNoAbs;
    bool d = bi != ai || ba != aa;
    if (!d && bi) d = B.part[i] != A.part[i] || __double_as_longlong(B.nx[i]) != __double_as_longlong(A.nx[i]) ||
                      __double_as_longlong(B.ny[i]) != __double_as_longlong(A.ny[i]) ||
                      __double_as_longlong(B.ng[i]) != __double_as_longlong(A.ng[i]);
    if (!d) return;
    int lo = 0, hi = nleaves - 1;   // the leaf that holds particle i: the last one that starts at or before it and is not empty
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (L.first[mid] <= i) lo = mid; else hi = mid - 1;
    }
    while (lo > 0 && L.last[lo] <= i) lo--;   // (leaves without particles share their start with the next one)
    double x0 = P.x[i], x1 = x0, y0 = P.y[i], y1 = y0;
    if (ai) { const double x = A.nx[i], y = A.ny[i]; x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y); }
    if (bi) { const double x = B.nx[i], y = B.ny[i]; x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y); }
    dirty[lo] = 1;
    double* b = dbox + 4ll * lo;
    atomic_min_f64(b + 0, x0); atomic_max_f64(b + 1, x1); atomic_min_f64(b + 2, y0); atomic_max_f64(b + 3, y1);
}
// per work unit: does its entry table name a dirty leaf? A unit that does not keeps last round's outcome as a whole.
__global__ void __launch_bounds__(128) k_unit_dirty(int nunits, Units U, GroupLists G, const unsigned char* __restrict__ leaf_dirty,
                                                    unsigned char* unit_dirty, unsigned char* group_dirty) {
    const int u = blockIdx.x;
    if (u >= nunits) return;
    const long long e0 = U.base[u];
    const int ne = U.count[u];
    int any = 0;
    for (int e = threadIdx.x; e < ne; e += 128) any |= leaf_dirty[G.leaf[e0 + e]];
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) {
        unit_dirty[u] = any ? 1 : 0;
        if (any) group_dirty[U.group[u]] = 1;   // (zeroed by the caller; every writer stores the same value)
    }
}
__global__ void k_merge_dyn(int n, Particles P, MergeState M, unsigned char* dyn, double4* tl) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int in = M.init[i], ab = M.absby[i];
    const bool d = in != 0 || ab != kNoAbs;
    dyn[i] = d ? 1 : 0;
    if (!d) return;
    double nx = 0, ny = 0, ng = 0;
    if (in) { nx = M.nx[i]; ny = M.ny[i]; ng = M.ng[i]; }
    tl[2ll * i] = make_double4(P.x[i], P.y[i], P.g[i], nx);
    tl[2ll * i + 1] = make_double4(ny, ng, __hiloint2double(in, ab), 0.);
}

// current bounding box and largest epsilon of every leaf's particles (after epsilon / merging). With a
// tentative merge solution M (epsilon rounds), the box also covers the post-merge position of every
// initiator, so that it bounds every state a source of the leaf can be seen in.
__global__ void k_leaf_box(LeafDev L, int nleaves, Particles P, MergeState M, double* lbox, const unsigned char* __restrict__ only = nullptr) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    if (only && !only[l]) return;   // (a later merge round: the entries of this leaf's particles are those of the last round)
    double x0 = DBL_MAX, x1 = -DBL_MAX, y0 = DBL_MAX, y1 = -DBL_MAX, em = 0;
    for (int i = L.first[l]; i < L.last[l]; i++) {
        double x = P.x[i], y = P.y[i];
        x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
        if (M.absby && M.init[i]) {
            x = M.nx[i]; y = M.ny[i];
            x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
        }
        if (P.g[i] != 0 && P.ie[i] != 0) em = fmax(em, 1. / P.ie[i]);
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
    const long long e0 = U.base[u];
    const long long e1 = e0 + U.count[u];
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
// one segment's terms of body_list_influence (:172-215) at the point (px, py), accumulated before the 1/2pi factor
__device__ __forceinline__ void body_slip_term(const BodyFull& B, int s, double px, double py, double& rx, double& ry) {
    if (!B.slip[s]) return;
    double dx = px - B.rx[s], dy = py - B.ry[s];
    double ee = 1. / B.ie[s];
    double k = B.g[s] / (dx * dx + dy * dy + ee * ee);
    rx += -dy * k; ry += dx * k;
}
__device__ __forceinline__ void body_motion_term(const BodyFull& B, const double* bp, int s, double px, double py, double& rx,
                                                 double& ry) {
    const double sx = bp[9], sy = bp[10], so = bp[11], ax = bp[0], ay = bp[1];
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
__device__ __forceinline__ bool body_moves(const double* bp) { return !(fabs(bp[9]) + fabs(bp[10]) + fabs(bp[11]) < 1E-10); }

__global__ void k_body_influence(int n, Particles P, BodyFull B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (P.g[i] == 0) return;
    const double px = P.x[i], py = P.y[i];
    double rx = 0, ry = 0;
    for (int ib = 0; ib < B.nbody; ib++) {
        const double* bp = B.bprop + 16 * ib;
        const int f = B.bfirst[ib], e = B.bfirst[ib + 1];
        if (bp[13] != 0)   // any slip segment (TBody::get_slip)
            for (int s = f; s < e; s++) body_slip_term(B, s, px, py, rx, ry);
        if (body_moves(bp))
            for (int s = f; s < e; s++) body_motion_term(B, bp, s, px, py, rx, ry);
    }
    P.vx[i] += rx * k1_2Pi;
    P.vy[i] += ry * k1_2Pi;
}

}  // namespace vv
