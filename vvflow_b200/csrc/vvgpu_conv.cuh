// K4 — the convective near field, MConvectiveFast::near_nodes_influence + biot_savart
// (libvvhd/src/MConvectiveFast.cpp:116-137) and the per-particle assembly of process_all_lists (:76-86).
//
// Same work decomposition as the other near-field passes (vvgpu_near.cuh): a CTA owns one work unit
// (a group of 32 consecutive leaves x <= 512 entries of its near list), its warps pull target leaves
// from a shared counter, and a warp expands the source leaves of its leaf into a flat list of particle
// indices in shared memory. What differs is the streaming: LANES HOLD SOURCES, the leaf's <= 15 targets
// are broadcast from shared memory and 2 x 15 accumulators stay in registers, so every lane does useful
// FP64 work on every pair and only one cross-lane reduction per leaf is needed.
//
// What the ncu source view of the previous version (k_near<ConvOp>) showed and what this kernel does about it
// (profiles/README.md):
//   * 14 % of the executed instructions were per-iteration bookkeeping (moving the prefetched records
//     into place, re-creating the dummy record, per-lane bounds predicates)  ->  the loop is unrolled
//     twice over ping-pong registers (no moves) and the index list is padded with the index of a dummy
//     record (g = 0) up to a whole iteration, so the loop carries no per-lane predicates;
//   * the unrolled loop must stay small: a build with two loop variants (41 KB of hot code) ran 2 x slower
//     with stall_no_instruction at 55 % of all samples  ->  targets are laid out partial-group-first
//     (CvGroups below), which needs one code body per size only for the partial group (18 KB).
// tools/microbench4.cu bounds this loop shape (two sources per lane, 30 accumulators, 3 CTAs per SM) at
// 1.33 T pairs/s with three target groups and 1.50 T with two; the kernel reaches 1.06-1.14 T in the step.
//
// Two builds of the source streaming (profiles/README.md, round 2):
//   default        per-lane loads through a flat index list in shared memory, one iteration prefetched ahead;
//   -DVV_CV_TMA=1  the source leaves' contiguous ranges staged into shared memory by TMA bulk copies (cp.async.bulk ->
//                  UBLKCP, completion on an mbarrier per stage, two stages per warp), lanes read the stage by position.
//                  Parity-green (tests/test_gpu_parity.py::test_conv_tma_variant) but 29 % SLOWER at N = 1M (3.81 vs
//                  2.95 ms): the ring has to be small to keep three CTAs per SM (168 registers are the occupancy
//                  limit), so per 192 sources a warp pays a batch scan, an mbarrier wait and a proxy fence; executed
//                  instructions +30 %, FP64 pipe 58.9 % -> 45.7 %. Built as lib/libvvgpu_tma.so for that test only.
#pragma once
#include "vvgpu_near.cuh"

#ifndef VV_CV_TMA
#define VV_CV_TMA 0
#endif

namespace vv {

#ifndef VV_CV_MINB
#define VV_CV_MINB 3
#endif
#ifndef VV_CV_FLUSH
#define VV_CV_FLUSH 1024
#endif
#ifndef VV_CV_NS
#define VV_CV_NS 2
#endif
#ifndef VV_CV_WARPS
#define VV_CV_WARPS 4
#endif
#if VV_CV_TMA
constexpr int kCvWarps = VV_CV_WARPS;
constexpr int kCvThreads = kCvWarps * 32;
#ifndef VV_CV_STAGE
#define VV_CV_STAGE 192
#endif
constexpr int kCvStage = VV_CV_STAGE;          // source records per stage of a warp's ring (a multiple of 32 * VV_CV_NS)

// Source tiles staged by TMA bulk copies. A source leaf is a CONTIGUOUS range of the packed 32-byte source records
// and neighbouring leaves of the list are mostly neighbours in memory, so a warp's source stream is a handful of
// contiguous runs: one cp.async.bulk per run copies it straight into the warp's stage of shared memory (completion
// counted in bytes on the stage's mbarrier) and the lanes then read the stage by position. No index list, no
// per-lane global load. Two stages per warp: the copies of the next one fly while this one is summed.
struct CvWarp {
    double4 stage[2][kCvStage];
    double2 txy[kMaxT + 1];    // target positions, broadcast to all lanes
    unsigned long long bar[2]; // mbarrier of each stage
};
struct CvShared {
    int4 ent[kUnitEntries];    // first particle, count of the entry's leaf, target-leaf mask
    int bounds[kGroupLeaves + 1];
    int next;                  // next target leaf of the group to hand out
    CvWarp w[kCvWarps];
};

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, non-tensor form): bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy accesses to shared memory (the lanes' reads of a stage) before async-proxy writes (its refill)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#else
constexpr int kCvWarps = VV_CV_WARPS;
constexpr int kCvThreads = kCvWarps * 32;
constexpr int kCvFlush = VV_CV_FLUSH;          // drain the index buffer once it holds this many sources
constexpr int kCvPiece = 16;                   // sources appended per entry and round
constexpr int kCvCap = kCvFlush + 32 * kCvPiece + 64;

struct CvWarp {
    int idx[kCvCap];           // flat list of source particle indices
    double2 txy[kMaxT + 1];    // target positions, broadcast to all lanes
};
struct CvShared {
    int4 ent[kUnitEntries];    // first particle, count of the entry's leaf, target-leaf mask
    int bounds[kGroupLeaves + 1];
    int next;                  // next target leaf of the group to hand out
    CvWarp w[kCvWarps];
};

#endif
// rotl(dr) * g / (|dr|^2 + eps^2). Reciprocal = rcp.approx.ftz.f64 (MUFU.RCP64H, relative error
// e0 <= 2^-19.9 measured on B200, tools/microbench2.cu) + one Newton step: 1/den = r0 (1 + e) up to
// e0^2 <= 2^-39.8 ~ 1e-12 per pair, two orders below the 1e-10 bar on velocities. 9 FP64 ops / pair.
__device__ __forceinline__ void cv_pair(const double2 p, const double4 s, double& ax, double& ay) {
    const double dx = p.x - s.x, dy = p.y - s.y;
    const double den = fma(dx, dx, fma(dy, dy, s.w));
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(den));
    const double e = fma(-den, r0, 1.0);
    const double gr = s.z * r0;
    const double w = fma(gr, e, gr);
    ax = fma(-dy, w, ax);
    ay = fma(dx, w, ay);
}

template <int NS, int BASE, int N>
__device__ __forceinline__ void cv_group(const double2* txy, const double4& s, const double4& s2, double (&ax)[kMaxT],
                                         double (&ay)[kMaxT]) {
#pragma unroll
    for (int t = 0; t < N; t++) {
        const double2 p = txy[BASE + t];   // one broadcast load serves both sources of the lane
        cv_pair(p, s, ax[BASE + t], ay[BASE + t]);
        if (NS == 2) cv_pair(p, s2, ax[BASE + t], ay[BASE + t]);
    }
}

// The nt <= 15 live targets of a pass sit in slots [0, p) u [5, 5 + 5 * nfull): a partial group of
// p = nt - 5 * nfull in 1..5 targets at base 0 and nfull in 0..2 FULL groups of five at bases 5 and 10. Only
// the partial group needs one code body per size; the instruction footprint of the streaming loop is
// what bounds this kernel's shape (a loop of more than ~30 KB thrashes the instruction cache: ncu
// showed stall_no_instruction at 55 % of all samples for a two-variant build).
struct CvGroups {
    int p, nfull;
    __device__ __forceinline__ void set(int nt) { nfull = (nt - 1) / 5; p = nt - 5 * nfull; }
    __device__ __forceinline__ int slot(int k) const { return (k < p) ? k : (5 + (k - p)); }   // k-th live target -> slot
    __device__ __forceinline__ bool used(int t) const { return (t < 5) ? (t < p) : (t < 5 + 5 * nfull); }
};

template <int NS>
__device__ __forceinline__ void cv_groups(const double2* txy, const double4& s, const double4& s2, const CvGroups& G,
                                          double (&ax)[kMaxT], double (&ay)[kMaxT]) {
    if (G.nfull > 0) {
        cv_group<NS, 5, 5>(txy, s, s2, ax, ay);
        if (G.nfull > 1) cv_group<NS, 10, 5>(txy, s, s2, ax, ay);
    }
    if (G.p >= 4) {
        if (G.p == 5) cv_group<NS, 0, 5>(txy, s, s2, ax, ay);
        else cv_group<NS, 0, 4>(txy, s, s2, ax, ay);
    } else if (G.p == 3) cv_group<NS, 0, 3>(txy, s, s2, ax, ay);
    else if (G.p == 2) cv_group<NS, 0, 2>(txy, s, s2, ax, ay);
    else cv_group<NS, 0, 1>(txy, s, s2, ax, ay);
}

#if VV_CV_TMA
// nit iterations of 32 * NS sources from a stage (padded with dummy records up to a whole iteration)
template <int NS>
__device__ __forceinline__ void cv_stream(const double4* stage, const double2* txy, int nit, const CvGroups& G,
                                          double (&ax)[kMaxT], double (&ay)[kMaxT], int lane) {
    constexpr int STEP = 32 * NS;
    const double4* sp = stage + lane;
    double4 a0 = sp[0], a1 = a0;
    if (NS == 2) a1 = sp[32];
    double4 b0 = a0, b1 = a1;
    int it = 0;
    for (;;) {
        if (it + 1 < nit) {   // the next records behind this iteration's math
            b0 = sp[STEP];
            if (NS == 2) b1 = sp[STEP + 32];
        }
        cv_groups<NS>(txy, a0, a1, G, ax, ay);
        if (++it >= nit) break;
        if (it + 1 < nit) {
            a0 = sp[2 * STEP];
            if (NS == 2) a1 = sp[2 * STEP + 32];
        }
        cv_groups<NS>(txy, b0, b1, G, ax, ay);
        if (++it >= nit) break;
        sp += 2 * STEP;
    }
}

// State of a warp's walk over the unit's entry table for one target leaf: the next batch of 32 entries, and per lane
// what is left of its entry (first record, count).
struct CvScan {
    int eb, f, cnt;
    bool pending;
};
// Fill one stage with up to kCvStage source records of the leaf's near leaves; returns how many. Per batch of 32 entries:
// lanes whose entry names the leaf hold (f, cnt); an exclusive scan of the counts places them behind each other in
// the stage; consecutive kept entries whose ranges touch are ONE run, copied by the run's first lane.
__device__ __forceinline__ int cv_fill(const CvShared& S, CvWarp& W, int st, const double4* __restrict__ src4, int ne, int lt,
                                       CvScan& sc, int lane) {
    int fill = 0;
    while (fill < kCvStage && (sc.pending || sc.eb < ne)) {
        if (!sc.pending) {
            const int e = sc.eb + lane;
            sc.eb += 32;
            sc.cnt = 0;
            if (e < ne) {
                const int4 en = S.ent[e];
                sc.f = en.x;
                sc.cnt = (((u32)en.z >> lt) & 1u) ? en.y : 0;
            }
            sc.pending = __any_sync(kFullMask, sc.cnt > 0);
            continue;
        }
        int inc = sc.cnt;   // inclusive scan; the shuffle's own predicate says whether the source lane exists
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
            asm volatile("{ .reg .pred p; .reg .s32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, t; }"
                         : "+r"(inc) : "r"(o));
        const int start = fill + inc - sc.cnt;                        // where this lane's records go
        const int take = max(0, min(sc.cnt, kCvStage - start));      // ... and how many of them fit
        // runs: a lane continues its predecessor's run if both copy something and the ranges touch in memory
        const int pend = __shfl_up_sync(kFullMask, sc.f + take, 1);
        const int ptake = __shfl_up_sync(kFullMask, take, 1);
        const int pcnt = __shfl_up_sync(kFullMask, sc.cnt, 1);
        const bool cont = lane > 0 && take > 0 && ptake > 0 && ptake == pcnt && pend == sc.f;
        const u32 heads = __ballot_sync(kFullMask, take > 0 && !cont);
        const int total = min(kCvStage, fill + __shfl_sync(kFullMask, inc, 31)) - fill;   // records copied this round
        if (lane == 0 && total > 0) mbar_expect_tx(&W.bar[st], (u32)total * 32u);
        __syncwarp();
        // a run ends before the next head; lanes that copy nothing sit at the end position of the run before them
        const u32 later = heads & ~((2u << lane) - 1u);
        const int last = later ? (__ffs(later) - 2) : 31;
        const int endpos = __shfl_sync(kFullMask, min(start + take, kCvStage), last);
        if (take > 0 && !cont) bulk_g2s(&W.stage[st][start], src4 + sc.f, (u32)(endpos - start) * 32u, &W.bar[st]);
        sc.f += take; sc.cnt -= take;
        fill += total;
        sc.pending = __any_sync(kFullMask, sc.cnt > 0);
    }
    return fill;
}

__global__ void __launch_bounds__(kCvThreads, VV_CV_MINB) k_conv(NearArgs A, ConvOp op, int dummy) {
    extern __shared__ __align__(16) unsigned char near_smem[];
    CvShared& S = *reinterpret_cast<CvShared*>(near_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const UnitPart up(A, false);
    const int u = A.u0 + up.b;   // (uniform cost per entry: DFS order keeps neighbouring units on neighbouring SMs)
    const int g = A.U.group[u];
    const int chunk = u - A.U.first[g];
    const bool multi = A.U.num[g] > 1;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, A.nleaves - l0);
    if (up.lt0 >= nl) return;
    const int lt_end = min(nl, up.lt1);
    if (tid <= nl) S.bounds[tid] = (tid < nl) ? A.L.first[l0 + tid] : A.L.last[l0 + nl - 1];
    if (tid == 0) S.next = up.lt0;
    if (lane == 0) {
        mbar_init(&S.w[warp].bar[0], 1); mbar_init(&S.w[warp].bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    u32 phase = 0;   // parity of each stage's mbarrier
    const long long e0 = A.U.base[u];
    const int ne = A.U.count[u];
    // ---- the unit's entry table, once per CTA
    for (int e = tid; e < ne; e += kCvThreads) {
        const int sl = A.G.leaf[e0 + e];
        const int f = A.L.first[sl];
        S.ent[e] = make_int4(f, A.L.last[sl] - f, (int)A.G.mask[e0 + e], 0);
    }
    __syncthreads();
    const int t0 = S.bounds[0], t1 = S.bounds[nl];
    CvWarp& W = S.w[warp];
    ConvOp::Part* scratch = (ConvOp::Part*)A.scratch;
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
            ConvOp::Tgt tg;
            const int i = tb + lane;
            const bool live = op.init(tg, A, i, leaf, lane < np);
            const u32 lm = __ballot_sync(kFullMask, live);
            const int nt = __popc(lm);
            if (nt == 0) continue;
            CvGroups GR;
            GR.set(nt);
            const int myt = live ? GR.slot(__popc(lm & lanemask_lt())) : -1;
            if (live) W.txy[myt] = make_double2(tg.x, tg.y);
            double ax[kMaxT], ay[kMaxT];
#pragma unroll
            for (int t = 0; t < kMaxT; t++) { ax[t] = 0; ay[t] = 0; }
            constexpr int kNS = VV_CV_NS, step = 32 * kNS;
            static_assert(kCvStage % step == 0, "a stage holds whole iterations");
            __syncwarp();
            // ---- stage the leaf's source runs (TMA), sum a stage while the next one is in flight
            CvScan sc{0, 0, 0, false};
            auto fill_stage = [&](int st) {
                const int fill = cv_fill(S, W, st, A.src4, ne, lt, sc, lane);
                const int upto = (fill + step - 1) / step * step;   // pad the last iteration with dummy records (g = 0)
                for (int k = fill + lane; k < upto; k += 32) W.stage[st][k] = make_double4(0., 0., 0., 1.);
                if (lane == 0) mbar_arrive(&W.bar[st]);
                return upto / step;
            };
            int st = 0;
            int nit = fill_stage(0);
            for (;;) {
                const bool more = sc.pending || sc.eb < ne;
                int nit2 = 0;
                if (more) nit2 = fill_stage(st ^ 1);
                mbar_wait(&W.bar[st], (phase >> st) & 1u);
                phase ^= 1u << st;
                __syncwarp();
                if (nit > 0) cv_stream<kNS>(W.stage[st], W.txy, nit, GR, ax, ay, lane);
                __syncwarp();
                fence_proxy_async();   // this stage's reads are done before a later bulk copy refills it
                if (!more) break;
                st ^= 1; nit = nit2;
            }
            // ---- lane sums -> the lane that owns the target
#pragma unroll
            for (int t = 0; t < kMaxT; t++) {
                if (GR.used(t)) {
                    double vx = ax[t], vy = ay[t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        vx += __shfl_xor_sync(kFullMask, vx, o);
                        vy += __shfl_xor_sync(kFullMask, vy, o);
                    }
                    if (myt == t) op.take(tg, vx, vy);
                }
            }
            if (live) {
                if (multi) scratch[sbase + (i - t0)] = op.part(tg);
                else op.finish(tg, A, i, leaf);
            }
            __syncwarp();
        }
    }
}

#else
// nit iterations of 32 * NS sources; idx[0 .. nit * 32 * NS) are valid indices (padded with the dummy record)
template <int NS>
__device__ __forceinline__ void cv_stream(const double4* __restrict__ src4, const double2* txy, const int* idx, int nit,
                                          const CvGroups& G, double (&ax)[kMaxT], double (&ay)[kMaxT], int lane) {
    constexpr int STEP = 32 * NS;
    const int* ip = idx + lane;
    double4 a0 = src4[ip[0]], a1 = a0;
    if (NS == 2) a1 = src4[ip[32]];
    double4 b0 = a0, b1 = a1;
    int it = 0;
    for (;;) {
        if (it + 1 < nit) {   // prefetch the next records behind this iteration's math
            b0 = src4[ip[STEP]];
            if (NS == 2) b1 = src4[ip[STEP + 32]];
        }
        cv_groups<NS>(txy, a0, a1, G, ax, ay);
        if (++it >= nit) break;
        if (it + 1 < nit) {
            a0 = src4[ip[2 * STEP]];
            if (NS == 2) a1 = src4[ip[2 * STEP + 32]];
        }
        cv_groups<NS>(txy, b0, b1, G, ax, ay);
        if (++it >= nit) break;
        ip += 2 * STEP;
    }
}

__global__ void __launch_bounds__(kCvThreads, VV_CV_MINB) k_conv(NearArgs A, ConvOp op, int dummy) {
    extern __shared__ __align__(16) unsigned char near_smem[];
    CvShared& S = *reinterpret_cast<CvShared*>(near_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const UnitPart up(A, false);
    const int u = A.u0 + up.b;   // (uniform cost per entry: DFS order keeps neighbouring units on neighbouring SMs)
    const int g = A.U.group[u];
    const int chunk = u - A.U.first[g];
    const bool multi = A.U.num[g] > 1;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, A.nleaves - l0);
    if (up.lt0 >= nl) return;
    const int lt_end = min(nl, up.lt1);
    if (tid <= nl) S.bounds[tid] = (tid < nl) ? A.L.first[l0 + tid] : A.L.last[l0 + nl - 1];
    if (tid == 0) S.next = up.lt0;
    const long long e0 = A.U.base[u];
    const int ne = A.U.count[u];
    // ---- the unit's entry table, once per CTA
    for (int e = tid; e < ne; e += kCvThreads) {
        const int sl = A.G.leaf[e0 + e];
        const int f = A.L.first[sl];
        S.ent[e] = make_int4(f, A.L.last[sl] - f, (int)A.G.mask[e0 + e], 0);
    }
    __syncthreads();
    const int t0 = S.bounds[0], t1 = S.bounds[nl];
    CvWarp& W = S.w[warp];
    ConvOp::Part* scratch = (ConvOp::Part*)A.scratch;
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
            ConvOp::Tgt tg;
            const int i = tb + lane;
            const bool live = op.init(tg, A, i, leaf, lane < np);
            const u32 lm = __ballot_sync(kFullMask, live);
            const int nt = __popc(lm);
            if (nt == 0) continue;
            CvGroups GR;
            GR.set(nt);
            const int myt = live ? GR.slot(__popc(lm & lanemask_lt())) : -1;
            if (live) W.txy[myt] = make_double2(tg.x, tg.y);
            double ax[kMaxT], ay[kMaxT];
#pragma unroll
            for (int t = 0; t < kMaxT; t++) { ax[t] = 0; ay[t] = 0; }
            constexpr int kNS = VV_CV_NS, step = 32 * kNS;
            __syncwarp();
            // ---- scan the entry table, expand the kept leaves into source indices, stream them
            int fill = 0, eb = 0, cnt = 0, f = 0;
            for (;;) {
                bool pending = __any_sync(kFullMask, cnt > 0);
                while (fill < kCvFlush && (pending || eb < ne)) {
                    if (!pending) {
                        const int e = eb + lane;
                        eb += 32;
                        if (e < ne) {
                            const int4 en = S.ent[e];
                            f = en.x;
                            cnt = (((u32)en.z >> lt) & 1u) ? en.y : 0;
                        }
                        pending = __any_sync(kFullMask, cnt > 0);
                        continue;
                    }
                    const int c = min(cnt, kCvPiece);
                    int inc = c;   // inclusive scan; the shuffle's own predicate says whether the source lane exists
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                        asm volatile("{ .reg .pred p; .reg .s32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, t; }"
                                     : "+r"(inc) : "r"(o));
                    const int tot = __shfl_sync(kFullMask, inc, 31);
                    int* dst = W.idx + fill + inc - c;
#pragma unroll
                    for (int k = 0; k < kCvPiece; k++)
                        if (k < c) dst[k] = f + k;
                    fill += tot; cnt -= c; f += c;
                    pending = __any_sync(kFullMask, cnt > 0);
                }
                const bool final = !pending && eb >= ne;
                int nit, upto;
                if (final) {   // pad the last iteration with the dummy record
                    nit = (fill + step - 1) / step;
                    upto = nit * step;
                    if (fill + lane < upto) W.idx[fill + lane] = dummy;
                    if (fill + lane + 32 < upto) W.idx[fill + lane + 32] = dummy;
                } else {
                    nit = fill / step;
                    upto = nit * step;
                }
                __syncwarp();
                if (nit > 0) {
                    cv_stream<kNS>(A.src4, W.txy, W.idx, nit, GR, ax, ay, lane);
                }
                if (final) break;
                // the incomplete last iteration goes to the front of the next drain
                const int carry = fill - upto;
                int v = 0, v2 = 0;
                if (lane < carry) v = W.idx[upto + lane];
                if (lane + 32 < carry) v2 = W.idx[upto + lane + 32];
                __syncwarp();
                if (lane < carry) W.idx[lane] = v;
                if (lane + 32 < carry) W.idx[lane + 32] = v2;
                fill = carry;
                __syncwarp();
            }
            // ---- lane sums -> the lane that owns the target
#pragma unroll
            for (int t = 0; t < kMaxT; t++) {
                if (GR.used(t)) {
                    double vx = ax[t], vy = ay[t];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        vx += __shfl_xor_sync(kFullMask, vx, o);
                        vy += __shfl_xor_sync(kFullMask, vy, o);
                    }
                    if (myt == t) op.take(tg, vx, vy);
                }
            }
            if (live) {
                if (multi) scratch[sbase + (i - t0)] = op.part(tg);
                else op.finish(tg, A, i, leaf);
            }
            __syncwarp();
        }
    }
}

#endif

}  // namespace vv
