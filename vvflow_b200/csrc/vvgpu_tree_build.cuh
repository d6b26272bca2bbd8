// K1 — the bit-exact tree build of vvgpu_tree.cuh in four launches and one host read-back:
//
//   k_tree_top      (cooperative, one CTA per SM) grows the TOP of the tree level by level, but only while a node
//                   holds more particles (segments) than one CTA can keep in shared memory. Per level every CTA
//                   redoes the cheap per-node work itself (DivideNode's tests, child allocation: a few hundred
//                   nodes), so a level costs three grid barriers: tile counts | partner table of the Hoare
//                   partition | swaps + Stretch of the children. Only the ranges of the nodes that still split are
//                   touched; settled ranges are never read again.
//   k_tree_sub      every remaining subtree (<= kSubCap particles, <= kSubSegCap segments) is finished by ONE CTA
//                   in shared memory: no grid barrier, coordinates loaded once and written back once, (g, caller
//                   index) gathered once at the end; it also does the subtree's own bottom-up (centres of mass,
//                   sizes) and top-down (DFS ids) sweeps. CTAs pull subtrees from a counter.
//   k_tree_topsweep one CTA: the same two sweeps over the (small) top tree; hands every subtree its node-id base
//                   and DFS offsets.
//   k_tree_relocate copies the subtrees' node records into the flat node arrays at their final ids.
//
// The previous single cooperative kernel passed ~130 grid barriers over ALL particles (2.2 ms at N = 1M, barrier
// stall 59 %, 0.07 of the HBM roofline for its algorithmic bytes); see profiles/ for this build's numbers.
#pragma once
#include "vvgpu_tree.cuh"

#include <cooperative_groups.h>
#include <cstdio>

namespace vv {
namespace cg = cooperative_groups;

#ifndef VV_SUB_CAP
#define VV_SUB_CAP 4096
#endif
constexpr int kSubCap = VV_SUB_CAP;     // particles of a CTA-built subtree
constexpr int kSubSegCap = 1024;        // body segments of a CTA-built subtree
constexpr int kSubSmallNode = 64;       // a child of at most this many particles is boxed by one thread
constexpr int kSubLvl = 1024;           // level width kept in shared memory (wider levels use the CTA's global arena)
constexpr int kSubThreads = 1024;
constexpr int kTopThreads = 1024;       // = tile of the top phase: one element per thread
#ifndef VV_TOP_BATCH
#define VV_TOP_BATCH 4
#endif
constexpr int kTopBatch = VV_TOP_BATCH; // tiles a CTA keeps in flight
constexpr int kMaxDepth = 4096;
constexpr int kLvlCache = 64;
constexpr unsigned short kNone16 = 0xffffu;

struct BuildState {
    int nnodes, depth, nleaves, err;    // err: 1 depth/capacity, 2 top level wider than the tables
    int ntop, nsub, dtop, subcursor;
    int lvl[kMaxDepth + 2];             // top phase: nodes of level d are [lvl[d], lvl[d+1])
    int hist[kMaxDepth + 2];            // nodes per depth (whole tree)
};

// one node of a CTA-built subtree, in the subtree's own numbering (0 = the subtree's root, a top-phase node)
struct __align__(16) SubNode {
    double x, y, h, w;
    double cmp[3], cmm[3];
    int first, last, sfirst, slast;     // global positions
    int ch1;                            // local id of child 1, -1: leaf
    int depth;                          // absolute
    int nl, nn, lstart, pre;            // subtree sizes; leaf index / pre-order id relative to the subtree root
    int axis, pad;
};
static_assert(sizeof(SubNode) == 128, "SubNode is one 128-byte record");

// a node of the level being split (decided) / being formed (box under reduction)
struct __align__(16) LvlNode {
    int first, cnt, sfirst, scnt;       // positions relative to the subtree
    union {
        u64 bb[4];
        struct { double mid; int rank /* among the splitting nodes of the level, -1: leaf */; int axis; double pad_[2]; } d;
    };
};
static_assert(sizeof(LvlNode) == 48, "LvlNode");

__device__ __forceinline__ u64 warp_min_u64(u64 v) {
    const u32 hi = (u32)(v >> 32), lo = (u32)v;
    const u32 mh = __reduce_min_sync(0xffffffffu, hi);
    const u32 ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((u64)mh << 32) | ml;
}
__device__ __forceinline__ u64 warp_max_u64(u64 v) {
    const u32 hi = (u32)(v >> 32), lo = (u32)v;
    const u32 mh = __reduce_max_sync(0xffffffffu, hi);
    const u32 ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return ((u64)mh << 32) | ml;
}
constexpr u64 kU64Max = 0xffffffffffffffffull;

// min/max of (x, y) over runs of equal `node` among the 32 lanes (runs are contiguous); lanes with node < 0 idle
__device__ __forceinline__ void seg_minmax(int node, u64& mnx, u64& mny, u64& mxx, u64& mxy, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int tn = __shfl_up_sync(0xffffffffu, node, o);
        const u64 a = __shfl_up_sync(0xffffffffu, mnx, o), b = __shfl_up_sync(0xffffffffu, mny, o);
        const u64 c = __shfl_up_sync(0xffffffffu, mxx, o), d = __shfl_up_sync(0xffffffffu, mxy, o);
        if (lane >= o && tn == node) {
            mnx = a < mnx ? a : mnx; mny = b < mny ? b : mny;
            mxx = c > mxx ? c : mxx; mxy = d > mxy ? d : mxy;
        }
    }
}

// ================================================================================= top phase
struct TopArgs {
    TreeDev T;
    BuildParams bp;
    double *px, *py, *pg;
    int* perm;
    const double *sx, *sy;
    int *segperm, *segtmp;
    int n, nseg;
    u32* enc;        // per position of a splitting node: 2 * (less elements of the node before it) + (is less)
    int* tmpR;       // partner table of the Hoare partition
    int* tilepre;    // per tile: less elements in the owning CTA's earlier tiles
    int* chunktot;   // per CTA: less elements in its tiles
    int* sublist;    // ST_SUB nodes in creation order
    BuildState* st;
    long long cap;   // node capacity
    int maxact;      // splitting nodes per level the shared tables hold
    int tcmax;       // tiles one CTA may own
};

__host__ __device__ inline size_t top_smem_bytes(int maxact, int tcmax) {
    return (size_t)maxact * (sizeof(double) + 7 * sizeof(int)) + ((size_t)maxact + 1 + tcmax) * sizeof(int);
}

// block-wide min/max of two boxes (left / right child): red[8][32]
__device__ __forceinline__ void top_commit_boxes(u64 (&v)[8], u64 (*red)[32], u64* bbL, u64* bbR, int lane, int warp) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const bool ismin = (q & 2) == 0;   // order: L.mnx L.mny L.mxx L.mxy R.mnx R.mny R.mxx R.mxy
        v[q] = ismin ? warp_min_u64(v[q]) : warp_max_u64(v[q]);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) red[q][warp] = v[q];
    }
    __syncthreads();
    if (warp < 8) {
        const bool ismin = (warp & 2) == 0;
        u64 r = red[warp][lane];   // kTopThreads / 32 == 32 warps
        r = ismin ? warp_min_u64(r) : warp_max_u64(r);
        if (lane == 0) {
            u64* bb = (warp < 4) ? bbL : bbR;
            if (ismin) { if (r != kU64Max) atomicMin(bb + (warp & 3), r); }
            else { if (r != 0) atomicMax(bb + (warp & 3), r); }
        }
    }
    __syncthreads();
}

#ifdef VV_TREE_TIMING
#define VV_TT(k) do { if (gtid == 0) { long long now_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)); tt_[k] += now_ - tlast_; tlast_ = now_; } } while (0)
#else
#define VV_TT(k) do { } while (0)
#endif
__global__ void __launch_bounds__(kTopThreads, 1) k_tree_top(TopArgs A) {
    static_assert(kTopThreads == 1024, "32 warps assumed by the box reduction");
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char top_smem[];
    __shared__ u32 sh[kTopThreads / 32 + 1];
    __shared__ u64 red[8][32];
    __shared__ int s_chunkex[kTopThreads];
    __shared__ u64 sh64[kTopThreads / 32 + 1];
    double* a_mid = (double*)top_smem;
    int* a_first = (int*)(a_mid + A.maxact);
    int* a_cnt = a_first + A.maxact;
    int* a_node = a_cnt + A.maxact;
    int* a_tile0 = a_node + A.maxact;
    int* a_sfirst = a_tile0 + A.maxact;
    int* a_scnt = a_sfirst + A.maxact;
    int* a_axis = a_scnt + A.maxact;
    int* a_gf = a_axis + A.maxact;          // maxact + 1: less elements in all tiles before the node's first
    int* s_mytp = a_gf + A.maxact + 1;      // tcmax: less elements in this CTA's tiles before tile k of its chunk
    TreeDev T = A.T;
    const int n = A.n, nseg = A.nseg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const long long gtid = (long long)blockIdx.x * blockDim.x + tid;
    const long long gsize = (long long)G * blockDim.x;
#ifdef VV_TREE_TIMING
    long long tt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tlast_));
#endif

    // ---- root: ranges, identity permutations, Stretch over everything
    if (gtid == 0) {
        T.first[0] = 0; T.last[0] = n; T.sfirst[0] = 0; T.slast[0] = nseg;
        T.ch1[0] = -1; T.depth[0] = 0; T.status[0] = ST_PENDING;
        bb_reset(T.bb);
        A.st->lvl[0] = 0; A.st->lvl[1] = 1; A.st->err = 0; A.st->subcursor = 0; A.st->depth = 0;
    }
    for (long long i = gtid; i < kMaxDepth + 2; i += gsize) A.st->hist[i] = 0;
    for (long long i = gtid; i < n; i += gsize) A.perm[i] = (int)i;
    for (long long k = gtid; k < nseg; k += gsize) A.segperm[k] = (int)k;
    grid.sync();
    {
        u64 v[8] = {kU64Max, kU64Max, 0, 0, kU64Max, kU64Max, 0, 0};
        for (long long i = gtid; i < (long long)n + nseg; i += gsize) {
            double x, y;
            if (i < n) { x = A.px[i]; y = A.py[i]; } else { x = A.sx[i - n]; y = A.sy[i - n]; }
            const u64 ex = enc_ordered(x), ey = enc_ordered(y);
            v[0] = ex < v[0] ? ex : v[0]; v[1] = ey < v[1] ? ey : v[1];
            v[2] = ex > v[2] ? ex : v[2]; v[3] = ey > v[3] ? ey : v[3];
        }
        top_commit_boxes(v, red, T.bb, T.bb, lane, warp);
    }
    grid.sync();

    int a0 = 0, a1 = 1, d = 0, nsub = 0;
    int failed = 0;
    VV_TT(0);
    for (;; d++) {
        const int na = a1 - a0;
        // ---- P1 (every CTA for itself; CTA 0 writes the node records): DivideNode's tests, child allocation
        int nsplit = 0;
        for (int base = 0; base < na; base += kTopThreads) {
            const int k = base + tid;
            const int nn = a0 + k;
            int st = 0, f = 0, l = 0, sf = 0, sl = 0;
            NodeGeom g{};
            if (k < na) {
                const u64* bb = T.bb + 4ll * nn;
                f = __ldcg(T.first + nn); l = __ldcg(T.last + nn); sf = __ldcg(T.sfirst + nn); sl = __ldcg(T.slast + nn);
                g = node_decide(__ldcg(bb), __ldcg(bb + 1), __ldcg(bb + 2), __ldcg(bb + 3), l - f, sl - sf, A.bp);
                st = g.leaf ? ST_LEAF : ((l - f <= kSubCap && sl - sf <= kSubSegCap) ? ST_SUB : ST_SPLIT);
            }
            u32 tot1, tot2;
            const u32 e1 = block_exclusive_scan<kTopThreads>(st == ST_SPLIT ? 1u : 0u, &tot1, sh);
            const u32 e2 = block_exclusive_scan<kTopThreads>(st == ST_SUB ? 1u : 0u, &tot2, sh);
            if (k < na) {
                const int j = nsplit + (int)e1;
                const long long c = (long long)a1 + 2ll * j;
                if (st == ST_SPLIT && j < A.maxact) {
                    a_first[j] = f; a_cnt[j] = l - f; a_node[j] = nn; a_sfirst[j] = sf; a_scnt[j] = sl - sf;
                    a_mid[j] = g.axis ? g.x : g.y; a_axis[j] = g.axis;
                }
                if (blockIdx.x == 0) {
                    T.x[nn] = g.x; T.y[nn] = g.y; T.h[nn] = g.h; T.w[nn] = g.w;
                    T.status[nn] = (unsigned char)st; T.axis[nn] = g.axis;
                    T.ch1[nn] = (st == ST_SPLIT) ? (int)c : -1;
                    if (st == ST_SUB) A.sublist[nsub + (int)e2] = nn;
                    if (st == ST_SPLIT && c + 1 < A.cap) {
                        for (int q = 0; q < 2; q++) {
                            T.depth[c + q] = d + 1; T.status[c + q] = ST_PENDING; T.ch1[c + q] = -1;
                            T.first[c + q] = T.last[c + q] = f;
                            T.sfirst[c + q] = T.slast[c + q] = sf;
                            bb_reset(T.bb + 4ll * (c + q));
                        }
                    }
                }
            }
            nsplit += (int)tot1; nsub += (int)tot2;
        }
        if (nsplit == 0) break;
        if (d + 1 >= kMaxDepth || (long long)a1 + 2ll * nsplit > A.cap) { failed = 1; break; }
        if (nsplit > A.maxact) { failed = 2; break; }
        __syncthreads();   // the a_* tables are complete
        VV_TT(1);
        // tiles of the splitting nodes' particle ranges
        int NT = 0;
        for (int base = 0; base < nsplit; base += kTopThreads) {
            const int j = base + tid;
            const u32 v = (j < nsplit) ? (u32)((a_cnt[j] + kTopThreads - 1) / kTopThreads) : 0u;
            u32 tot;
            const u32 ex = block_exclusive_scan<kTopThreads>(v, &tot, sh);
            if (j < nsplit) a_tile0[j] = NT + (int)ex;
            NT += (int)tot;
        }
        __syncthreads();
        // every CTA owns TC consecutive tiles (the last ones may own fewer or none): owner(t) = t / TC is one division
        const int TC = (NT + G - 1) / G;
        const int t_lo = min(NT, (int)blockIdx.x * TC), t_hi = min(NT, ((int)blockIdx.x + 1) * TC);
        // node of a tile: the LAST j with a_tile0[j] <= t (nodes without particles own no tile)
        auto find_node = [&](int t) {
            int lo = 0, hi = nsplit - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a_tile0[mid] <= t) lo = mid; else hi = mid - 1; }
            return lo;
        };
        // The three particle passes below walk this CTA's tiles kTopBatch at a time: every pass is a chain of dependent
        // L2 accesses per element (coordinate | scan | partner -> payload), and with one tile in flight per CTA that
        // latency was the whole cost of a level (0.75 ms at N = 1M for ~70 MB of traffic per level).
        // ---- P2: less-counts of my tiles
        {
            int running = 0;
            int j = (t_lo < t_hi) ? find_node(t_lo) : 0;
            for (int t = t_lo; t < t_hi; t += kTopBatch) {
                int jq[kTopBatch];
                double cv[kTopBatch];
                bool ok[kTopBatch];
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    jq[q] = -1; ok[q] = false; cv[q] = 0;
                    if (t + q < t_hi) {
                        while (j + 1 < nsplit && a_tile0[j + 1] <= t + q) j++;
                        jq[q] = j;
                        const int p = a_first[j] + (t + q - a_tile0[j]) * kTopThreads + tid;
                        if (p < a_first[j] + a_cnt[j]) { ok[q] = true; cv[q] = a_axis[j] ? __ldcg(A.px + p) : __ldcg(A.py + p); }
                    }
                }
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    if (jq[q] < 0) continue;
                    const int cnt = __syncthreads_count(ok[q] && cv[q] < a_mid[jq[q]]);
                    if (tid == 0) { A.tilepre[t + q] = running; s_mytp[t + q - t_lo] = running; }
                    running += cnt;
                }
            }
            if (tid == 0) A.chunktot[blockIdx.x] = running;
        }
        VV_TT(2);
        grid.sync();
        VV_TT(3);
        // ---- P3: prefix over the CTAs' totals; P4: partner table, child ranges
        {
            const u32 ctv = (tid < G) ? (u32)__ldcg(A.chunktot + tid) : 0u;
            u32 total;
            const u32 cex = block_exclusive_scan<kTopThreads>(ctv, &total, sh);
            s_chunkex[tid] = (int)cex;
            __syncthreads();
            // less elements before every splitting node (its first tile may belong to another CTA): one load per node, here
            for (int jn = tid; jn <= nsplit; jn += kTopThreads) {
                const int t = (jn < nsplit) ? a_tile0[jn] : NT;
                a_gf[jn] = (t >= NT) ? (int)total : s_chunkex[(unsigned)t / (unsigned)TC] + __ldcg(A.tilepre + t);
            }
            __syncthreads();
            const int mybase = s_chunkex[blockIdx.x];
            int j = (t_lo < t_hi) ? find_node(t_lo) : 0;
            for (int t = t_lo; t < t_hi; t += kTopBatch) {
                int jq[kTopBatch], mq[kTopBatch], gpre[kTopBatch], pq[kTopBatch];
                bool fl[kTopBatch], ok[kTopBatch];
                u64 pk = 0;
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    jq[q] = -1; ok[q] = false; fl[q] = false; mq[q] = 0; gpre[q] = 0; pq[q] = 0;
                    if (t + q < t_hi) {
                        while (j + 1 < nsplit && a_tile0[j + 1] <= t + q) j++;
                        jq[q] = j;
                        const int f = a_first[j], cnt = a_cnt[j], tj0 = a_tile0[j];
                        const int Gf = a_gf[j];
                        mq[q] = a_gf[j + 1] - Gf;
                        gpre[q] = mybase + s_mytp[t + q - t_lo] - Gf;
                        pq[q] = f + (t + q - tj0) * kTopThreads + tid;
                        if (pq[q] < f + cnt) {
                            ok[q] = true;
                            fl[q] = (a_axis[j] ? __ldcg(A.px + pq[q]) : __ldcg(A.py + pq[q])) < a_mid[j];
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) pk |= (u64)(fl[q] ? 1u : 0u) << (16 * q);   // a tile holds <= 1024 < 2^16
                u64 tot;
                const u64 ex = block_exclusive_scan_t<u64, kTopThreads>(pk, &tot, sh64);
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    if (jq[q] < 0) continue;
                    const int jj = jq[q], f = a_first[jj], m = mq[q];
                    if (ok[q]) {
                        const int le = gpre[q] + (int)((ex >> (16 * q)) & 0xffffu), rel = pq[q] - f;
                        A.enc[pq[q]] = (u32)le * 2u + (fl[q] ? 1u : 0u);
                        if (rel >= m && fl[q]) A.tmpR[f + (m - le - 1)] = pq[q];
                    }
                    if (t + q == a_tile0[jj] && tid == 0) {
                        const int c = a1 + 2 * jj;
                        T.first[c] = f; T.last[c] = f + m;
                        T.first[c + 1] = f + m; T.last[c + 1] = f + a_cnt[jj];
                    }
                }
            }
        }
        VV_TT(4);
        grid.sync();
        VV_TT(3);
        // ---- P5: in-place swaps of (x, y, caller index) — each pair is touched by exactly one thread; g follows the
        // caller index once, at the end of the build — + Stretch of the two children, folded over the tiles of a node
        {
            int j = (t_lo < t_hi) ? find_node(t_lo) : 0;
            int accj = -1;
            u64 v[8] = {kU64Max, kU64Max, 0, 0, kU64Max, kU64Max, 0, 0};
            for (int t = t_lo; t < t_hi; t += kTopBatch) {
                int jq[kTopBatch], pq[kTopBatch], qq[kTopBatch], role[kTopBatch];   // role 0 none, 1 swaps, 2 stays left, 3 stays right
                double ax[kTopBatch], ay[kTopBatch], bx[kTopBatch], by[kTopBatch];
                int ai[kTopBatch], bi[kTopBatch];
                u32 e[kTopBatch];
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    jq[q] = -1; role[q] = 0; pq[q] = 0; e[q] = 0;
                    if (t + q < t_hi) {
                        while (j + 1 < nsplit && a_tile0[j + 1] <= t + q) j++;
                        jq[q] = j;
                        pq[q] = a_first[j] + (t + q - a_tile0[j]) * kTopThreads + tid;
                        if (pq[q] < a_first[j] + a_cnt[j]) { role[q] = 4; e[q] = __ldcg(A.enc + pq[q]); }
                    }
                }
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    qq[q] = 0;
                    if (role[q] != 4) continue;
                    const int jj = jq[q], f = a_first[jj], m = a_gf[jj + 1] - a_gf[jj];
                    const bool flag = e[q] & 1u;
                    const int le = (int)(e[q] >> 1), rel = pq[q] - f;
                    if (rel < m && !flag) { role[q] = 1; qq[q] = __ldcg(A.tmpR + f + (rel - le)); }
                    else if (rel < m) role[q] = 2;
                    else if (!flag) role[q] = 3;
                    else role[q] = 0;   // a less element of the right part is moved (and counted) by its partner
                }
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    ax[q] = ay[q] = bx[q] = by[q] = 0; ai[q] = bi[q] = 0;
                    if (role[q] == 0) continue;
                    ax[q] = __ldcg(A.px + pq[q]); ay[q] = __ldcg(A.py + pq[q]);
                    if (role[q] == 1) {
                        ai[q] = __ldcg(A.perm + pq[q]);
                        bx[q] = __ldcg(A.px + qq[q]); by[q] = __ldcg(A.py + qq[q]); bi[q] = __ldcg(A.perm + qq[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < kTopBatch; q++) {
                    if (jq[q] < 0) continue;
                    if (jq[q] != accj) {   // uniform across the CTA
                        if (accj >= 0) {
                            const int c = a1 + 2 * accj;
                            top_commit_boxes(v, red, T.bb + 4ll * c, T.bb + 4ll * (c + 1), lane, warp);
                            v[0] = v[1] = v[4] = v[5] = kU64Max; v[2] = v[3] = v[6] = v[7] = 0;
                        }
                        accj = jq[q];
                    }
                    if (role[q] == 0) continue;
                    u64 lx = 0, ly = 0, rx = 0, ry = 0;
                    bool hasl = false, hasr = false;
                    if (role[q] == 1) {
                        A.px[pq[q]] = bx[q]; A.py[pq[q]] = by[q]; A.perm[pq[q]] = bi[q];
                        A.px[qq[q]] = ax[q]; A.py[qq[q]] = ay[q]; A.perm[qq[q]] = ai[q];
                        lx = enc_ordered(bx[q]); ly = enc_ordered(by[q]); rx = enc_ordered(ax[q]); ry = enc_ordered(ay[q]);
                        hasl = hasr = true;
                    } else if (role[q] == 2) { lx = enc_ordered(ax[q]); ly = enc_ordered(ay[q]); hasl = true; }
                    else { rx = enc_ordered(ax[q]); ry = enc_ordered(ay[q]); hasr = true; }
                    if (hasl) { v[0] = lx < v[0] ? lx : v[0]; v[1] = ly < v[1] ? ly : v[1]; v[2] = lx > v[2] ? lx : v[2]; v[3] = ly > v[3] ? ly : v[3]; }
                    if (hasr) { v[4] = rx < v[4] ? rx : v[4]; v[5] = ry < v[5] ? ry : v[5]; v[6] = rx > v[6] ? rx : v[6]; v[7] = ry > v[7] ? ry : v[7]; }
                }
            }
            if (accj >= 0) {
                const int c = a1 + 2 * accj;
                top_commit_boxes(v, red, T.bb + 4ll * c, T.bb + 4ll * (c + 1), lane, warp);
            }
            // stable split of the segment lists (DistributeContent(LList&), TSortedTree.cpp:139-148): one CTA per node
            if (nseg) {
                for (int j2 = blockIdx.x; j2 < nsplit; j2 += G) {
                    const int sc = a_scnt[j2];
                    if (sc == 0) continue;
                    const int sf = a_sfirst[j2], c = a1 + 2 * j2;
                    const double mid = a_mid[j2];
                    const int ax = a_axis[j2];
                    int ms = 0;
                    for (int base = 0; base < sc; base += kTopThreads) {
                        const int k = base + tid;
                        bool fl = false;
                        if (k < sc) { const int s = A.segperm[sf + k]; fl = (ax ? A.sx[s] : A.sy[s]) < mid; }
                        ms += __syncthreads_count(fl);
                    }
                    u64 v[8] = {kU64Max, kU64Max, 0, 0, kU64Max, kU64Max, 0, 0};
                    int carry = 0;
                    for (int base = 0; base < sc; base += kTopThreads) {
                        const int k = base + tid;
                        bool fl = false;
                        int s = 0;
                        if (k < sc) { s = A.segperm[sf + k]; fl = (ax ? A.sx[s] : A.sy[s]) < mid; }
                        u32 tot;
                        const u32 ex = block_exclusive_scan<kTopThreads>(fl ? 1u : 0u, &tot, sh);
                        if (k < sc) {
                            const int le = carry + (int)ex;
                            A.segtmp[sf + (fl ? le : ms + (k - le))] = s;
                            const u64 ex_ = enc_ordered(A.sx[s]), ey_ = enc_ordered(A.sy[s]);
                            const int o = fl ? 0 : 4;
                            v[o] = v[o + 2] = ex_; v[o + 1] = v[o + 3] = ey_;
                        }
                        carry += (int)tot;
                        // several chunks: fold as we go (the commit is idempotent for identities)
                        if (base + kTopThreads < sc) {
                            top_commit_boxes(v, red, T.bb + 4ll * c, T.bb + 4ll * (c + 1), lane, warp);
                            v[0] = v[1] = v[4] = v[5] = kU64Max; v[2] = v[3] = v[6] = v[7] = 0;
                        }
                    }
                    top_commit_boxes(v, red, T.bb + 4ll * c, T.bb + 4ll * (c + 1), lane, warp);
                    for (int k = tid; k < sc; k += kTopThreads) A.segperm[sf + k] = A.segtmp[sf + k];
                    if (tid == 0) {
                        T.sfirst[c] = sf; T.slast[c] = sf + ms;
                        T.sfirst[c + 1] = sf + ms; T.slast[c + 1] = sf + sc;
                    }
                    __syncthreads();
                }
            }
        }
        a0 = a1; a1 += 2 * nsplit;
        if (gtid == 0) A.st->lvl[d + 2] = a1;
        VV_TT(5);
        grid.sync();
        VV_TT(3);
    }
#ifdef VV_TREE_TIMING
    if (gtid == 0) printf("top timing us: root %.0f P1 %.0f P2 %.0f gridsync(wait of CTA0) %.0f P4 %.0f P5 %.0f levels %d\n", tt_[0] * 1e-3, tt_[1] * 1e-3,
                          tt_[2] * 1e-3, tt_[3] * 1e-3, tt_[4] * 1e-3, tt_[5] * 1e-3, d);
#endif
    if (gtid == 0) {
        A.st->err = failed;
        A.st->ntop = a1; A.st->dtop = d; A.st->nsub = nsub;
    }
}

// exclusive prefix of one value per WARP over the 32 warps of a CTA, ONE barrier: every warp scans all 32 totals itself.
// `sh` (32 words) must not be written again before another barrier has passed.
__device__ __forceinline__ u32 warp_totals_scan(u32 wsum, u32* total, u32* sh, int lane, int warp) {
    if (lane == 0) sh[warp] = wsum;
    __syncthreads();
    const u32 v = sh[lane];
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    return __shfl_sync(0xffffffffu, inc - v, warp);
}

// ================================================================================= CTA-built subtrees
struct SubArgs {
    TreeDev T;
    BuildParams bp;
    double *px, *py;
    const double* pg;             // in the CALLER's order (the build moves x, y and the caller index only)
    int* perm;
    const double *sx, *sy;
    int* segperm;
    SubNode* scratch;             // subtree of root r: records at scratch + 2 * (first[r] + sfirst[r])
    const int* sublist;
    BuildState* st;
    unsigned char* arena;         // per CTA: sub_arena_bytes()
};

// per-node state of the subtree's own sweeps
struct __align__(16) SweepNode {
    double cmp[3], cmm[3];
    double nx, ny;                      // node centre (the centre of mass of an empty sign)
    int nl, nn, ch1, first, cnt, lstart, pre, pad;
};
static_assert(sizeof(SweepNode) == 96, "SweepNode");

constexpr int kSubNodesMax = 2 * (kSubCap + kSubSegCap) + 2;
constexpr size_t kArenaLvl = (size_t)(kMaxDepth + 2) * sizeof(int);
constexpr size_t kArenaTab = (size_t)(kSubCap + kSubSegCap + 2) * sizeof(LvlNode);
constexpr size_t kArenaSweep = (size_t)kSubNodesMax * sizeof(SweepNode);
__host__ __device__ constexpr size_t sub_arena_bytes() { return ((kArenaLvl + 15) / 16 * 16) + 2 * kArenaTab + kArenaSweep; }

struct SubSmem {
    double sx[kSubCap], sy[kSubCap];
    // the four per-position tables of the level loop; afterwards the same bytes hold g (double) per position
    unsigned short sidx[kSubCap], spn[kSubCap], sG[kSubCap + 8], tmpR[kSubCap];
    LvlNode tab[2][kSubLvl];        // afterwards: SweepNode[]
    double ssx[kSubSegCap], ssy[kSubSegCap];
    int sgi[kSubSegCap];            // global segment index of the segment loaded at k
    unsigned short sord[2][kSubSegCap], ssn[2][kSubSegCap], sGs[kSubSegCap + 8];
    int lvl[kLvlCache + 2];
    u32 sh[kSubThreads / 32 + 1];
    u32 shA[32], shB[32], shC[32];   // one-barrier scans (B2 particles, B2 segments, B6)
    int nbig;
    int cur_sub;
};
constexpr int kSweepSmemNodes = (int)(sizeof(LvlNode) * 2 * kSubLvl / sizeof(SweepNode));

__global__ void __launch_bounds__(kSubThreads, 1) k_tree_sub(SubArgs A) {
    extern __shared__ __align__(16) unsigned char sub_smem[];
    SubSmem& S = *reinterpret_cast<SubSmem*>(sub_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kRounds = kSubCap / kSubThreads;   // positions per thread: warp w owns [w * 32 * kRounds, ...)
    static_assert(kSubCap % kSubThreads == 0 && kSubSegCap <= kSubThreads, "position mapping");
    unsigned char* arena = A.arena + (size_t)blockIdx.x * sub_arena_bytes();
    int* g_lvl = (int*)arena;
    LvlNode* g_tab[2] = {(LvlNode*)(arena + (kArenaLvl + 15) / 16 * 16), (LvlNode*)(arena + (kArenaLvl + 15) / 16 * 16 + kArenaTab)};
    SweepNode* g_sweep = (SweepNode*)(arena + (kArenaLvl + 15) / 16 * 16 + 2 * kArenaTab);
    const TreeDev& T = A.T;
    const int nsub = A.st->nsub;

    for (;;) {
        if (tid == 0) S.cur_sub = atomicAdd(&A.st->subcursor, 1);
        __syncthreads();
        const int s = S.cur_sub;
        if (s >= nsub) break;
        const int root = A.sublist[s];
        const int f0 = T.first[root], np = T.last[root] - f0;
        const int sf0 = T.sfirst[root], ns = T.slast[root] - sf0;
        const int rdepth = T.depth[root];
        SubNode* rec = A.scratch + 2ll * ((long long)f0 + sf0);
        auto lvl_set = [&](int d, int v) { if (d < kLvlCache + 2) S.lvl[d] = v; g_lvl[d] = v; };
        auto lvl_get = [&](int d) { return (d < kLvlCache + 2) ? S.lvl[d] : g_lvl[d]; };

        // ---- load the subtree's objects
#pragma unroll
        for (int i = 0; i < kRounds; i++) {
            const int p = (warp * kRounds + i) * 32 + lane;
            if (p < np) { S.sx[p] = A.px[f0 + p]; S.sy[p] = A.py[f0 + p]; S.sidx[p] = (unsigned short)p; S.spn[p] = 0; }
        }
        if (tid < ns) {
            const int gi = A.segperm[sf0 + tid];
            S.sgi[tid] = gi; S.ssx[tid] = A.sx[gi]; S.ssy[tid] = A.sy[gi];
            S.sord[0][tid] = (unsigned short)tid; S.ssn[0][tid] = 0;
        }
        int sb = 0;   // current segment order / node buffer
        LvlNode* cur = S.tab[0];
        int curbuf = 0;
        if (tid == 0) {
            cur[0].first = 0; cur[0].cnt = np; cur[0].sfirst = 0; cur[0].scnt = ns;
            const int ax = T.axis[root];
            cur[0].d.mid = ax ? T.x[root] : T.y[root]; cur[0].d.rank = 0; cur[0].d.axis = ax;
            lvl_set(0, 0); lvl_set(1, 1);
            SubNode& r = rec[0];
            r.first = f0; r.last = f0 + np; r.sfirst = sf0; r.slast = sf0 + ns; r.ch1 = 1; r.depth = rdepth;
            r.x = T.x[root]; r.y = T.y[root]; r.h = T.h[root]; r.w = T.w[root]; r.axis = ax;
        }
        __syncthreads();
        int width = 1, nsplit = 1, total = 1, d = 0;
        int err = 0;
        // ---- level loop: split the nodes of level d into level d + 1
        while (nsplit > 0) {
            const int nw = 2 * nsplit;   // width of the next level
            if (rdepth + d + 1 >= kMaxDepth || total + nw > kSubNodesMax || total + nw > 2 * (np + ns)) { err = 1; break; }
            LvlNode* nxt = (nw <= kSubLvl) ? S.tab[curbuf ^ 1] : g_tab[curbuf ^ 1];
            // children start empty at their parent's first position; boxes at the Stretch sentinels
            for (int k = tid; k < width; k += kSubThreads) {
                const int r = cur[k].d.rank;
                if (r < 0) continue;
                for (int q = 0; q < 2; q++) {
                    LvlNode& c = nxt[2 * r + q];
                    c.first = cur[k].first; c.cnt = 0; c.sfirst = cur[k].sfirst; c.scnt = 0;
                    bb_reset(c.bb);
                }
            }
            // B2: flags "coord < mid" and their exclusive prefix over the positions (ballots; warp totals scanned once)
            u32 bal[kRounds];
            {
                u32 wsum = 0;
#pragma unroll
                for (int i = 0; i < kRounds; i++) {
                    const int p = (warp * kRounds + i) * 32 + lane;
                    bool fl = false;
                    if (p < np) {
                        const unsigned short k = S.spn[p];
                        if (k != kNone16 && cur[k].d.rank >= 0) fl = (cur[k].d.axis ? S.sx[p] : S.sy[p]) < cur[k].d.mid;
                    }
                    bal[i] = __ballot_sync(0xffffffffu, fl);
                    wsum += __popc(bal[i]);
                }
                u32 tot;
                u32 ex = warp_totals_scan(wsum, &tot, S.shA, lane, warp);
#pragma unroll
                for (int i = 0; i < kRounds; i++) {
                    const int p = (warp * kRounds + i) * 32 + lane;
                    if (p <= np) S.sG[p] = (unsigned short)(ex + __popc(bal[i] & lanemask_lt()));
                    ex += __popc(bal[i]);
                }
                if (np == kSubCap && tid == kSubThreads - 1) S.sG[np] = (unsigned short)tot;
            }
            bool sfl = false;
            if (tid == 0) S.nbig = 0;
            if (ns) {
                if (tid < ns) {
                    const unsigned short k = S.ssn[sb][tid];
                    if (k != kNone16 && cur[k].d.rank >= 0) {
                        const int o = S.sord[sb][tid];
                        sfl = (cur[k].d.axis ? S.ssx[o] : S.ssy[o]) < cur[k].d.mid;
                    }
                }
                u32 tot;
                const u32 sbal = __ballot_sync(0xffffffffu, sfl);
                const u32 ex = warp_totals_scan(__popc(sbal), &tot, S.shB, lane, warp) + __popc(sbal & lanemask_lt());
                if (tid <= ns) S.sGs[tid] = (unsigned short)ex;
                if (ns == kSubThreads && tid == kSubThreads - 1) S.sGs[ns] = (unsigned short)tot;
            }
            __syncthreads();
            // B3: partner table; child ranges
#pragma unroll
            for (int i = 0; i < kRounds; i++) {
                const int p = (warp * kRounds + i) * 32 + lane;
                if (p >= np) continue;
                const unsigned short k = S.spn[p];
                if (k == kNone16 || cur[k].d.rank < 0) continue;
                const int f = cur[k].first, cnt = cur[k].cnt;
                const int Gf = S.sG[f], m = S.sG[f + cnt] - Gf, le = S.sG[p] - Gf, rel = p - f;
                const bool fl = (bal[i] >> lane) & 1u;
                if (rel >= m && fl) S.tmpR[f + (m - le - 1)] = (unsigned short)p;
                if (rel == 0) {
                    LvlNode* c = nxt + 2 * cur[k].d.rank;
                    c[0].first = f; c[0].cnt = m; c[1].first = f + m; c[1].cnt = cnt - m;
                }
            }
            if (ns && tid < ns) {
                const unsigned short k = S.ssn[sb][tid];
                unsigned short nk = kNone16, dst = (unsigned short)tid;
                if (k != kNone16 && cur[k].d.rank >= 0) {
                    const int f = cur[k].sfirst, cnt = cur[k].scnt;
                    const int Gf = S.sGs[f], m = S.sGs[f + cnt] - Gf, le = S.sGs[tid] - Gf, rel = tid - f;
                    dst = (unsigned short)(sfl ? f + le : f + m + (rel - le));
                    nk = (unsigned short)(2 * cur[k].d.rank + (sfl ? 0 : 1));
                    if (rel == 0) {
                        LvlNode* c = nxt + 2 * cur[k].d.rank;
                        c[0].sfirst = f; c[0].scnt = m; c[1].sfirst = f + m; c[1].scnt = cnt - m;
                    }
                }
                S.sord[sb ^ 1][dst] = S.sord[sb][tid];
                S.ssn[sb ^ 1][dst] = nk;
            }
            __syncthreads();
            // B4: swaps and the positions' new nodes
#pragma unroll
            for (int i = 0; i < kRounds; i++) {
                const int p = (warp * kRounds + i) * 32 + lane;
                if (p >= np) continue;
                const unsigned short k = S.spn[p];
                if (k == kNone16) continue;
                const int r = cur[k].d.rank;
                if (r < 0) { S.spn[p] = kNone16; continue; }   // a leaf: settled
                const int f = cur[k].first, cnt = cur[k].cnt;
                const int Gf = S.sG[f], m = S.sG[f + cnt] - Gf, le = S.sG[p] - Gf, rel = p - f;
                const bool fl = (bal[i] >> lane) & 1u;
                if (rel < m && !fl) {
                    const int q = S.tmpR[f + (rel - le)];
                    const double tx = S.sx[p], ty = S.sy[p];
                    const unsigned short ti = S.sidx[p];
                    S.sx[p] = S.sx[q]; S.sy[p] = S.sy[q]; S.sidx[p] = S.sidx[q];
                    S.sx[q] = tx; S.sy[q] = ty; S.sidx[q] = ti;
                }
                S.spn[p] = (unsigned short)(2 * r + (rel >= m ? 1 : 0));
            }
            sb ^= 1;
            __syncthreads();
            // B5: Stretch of the children. After the partition a child is a contiguous range: a small child is folded by
            // ONE thread, a big one by one warp (a segmented warp reduction over all positions cost ~200 instructions
            // per element-round and was most of this kernel); segments, few, are added with atomics afterwards.
            {
                unsigned short* big = S.tmpR;   // free again: the partner table was consumed in B4
                for (int k = tid; k < nw; k += kSubThreads) {
                    const int f = nxt[k].first, cnt = nxt[k].cnt;
                    if (cnt > kSubSmallNode) { big[atomicAdd(&S.nbig, 1)] = (unsigned short)k; continue; }
                    if (cnt == 0) continue;
                    u64 mnx = kU64Max, mny = kU64Max, mxx = 0, mxy = 0;
                    for (int i = 0; i < cnt; i++) {
                        const u64 ex = enc_ordered(S.sx[f + i]), ey = enc_ordered(S.sy[f + i]);
                        mnx = ex < mnx ? ex : mnx; mny = ey < mny ? ey : mny; mxx = ex > mxx ? ex : mxx; mxy = ey > mxy ? ey : mxy;
                    }
                    u64* bb = nxt[k].bb;
                    bb[0] = mnx; bb[1] = mny; bb[2] = mxx; bb[3] = mxy;
                }
                __syncthreads();
                const int nbig = S.nbig;
                for (int b = warp; b < nbig; b += kSubThreads / 32) {
                    const int k = big[b];
                    const int f = nxt[k].first, cnt = nxt[k].cnt;
                    u64 mnx = kU64Max, mny = kU64Max, mxx = 0, mxy = 0;
                    for (int i = lane; i < cnt; i += 32) {
                        const u64 ex = enc_ordered(S.sx[f + i]), ey = enc_ordered(S.sy[f + i]);
                        mnx = ex < mnx ? ex : mnx; mny = ey < mny ? ey : mny; mxx = ex > mxx ? ex : mxx; mxy = ey > mxy ? ey : mxy;
                    }
                    mnx = warp_min_u64(mnx); mny = warp_min_u64(mny); mxx = warp_max_u64(mxx); mxy = warp_max_u64(mxy);
                    if (lane == 0) { u64* bb = nxt[k].bb; bb[0] = mnx; bb[1] = mny; bb[2] = mxx; bb[3] = mxy; }
                }
            }
            if (ns) {
                __syncthreads();
                if (tid < ns) {
                    const unsigned short k = S.ssn[sb][tid];
                    if (k != kNone16) {
                        const int o = S.sord[sb][tid];
                        const u64 ex = enc_ordered(S.ssx[o]), ey = enc_ordered(S.ssy[o]);
                        u64* bb = nxt[k].bb;
                        atomicMin(bb + 0, ex); atomicMin(bb + 1, ey); atomicMax(bb + 2, ex); atomicMax(bb + 3, ey);
                    }
                }
            }
            __syncthreads();
            // B6: DivideNode's tests for the new level; node records
            const int base_id = total;           // local id of the new level's first node
            int nsplit2 = 0;
            for (int b0 = 0; b0 < nw; b0 += kSubThreads) {
                const int k = b0 + tid;
                NodeGeom g{};
                bool split = false;
                LvlNode ln{};
                if (k < nw) {
                    ln.first = nxt[k].first; ln.cnt = nxt[k].cnt; ln.sfirst = nxt[k].sfirst; ln.scnt = nxt[k].scnt;
                    volatile u64* bb = nxt[k].bb;
                    g = node_decide(bb[0], bb[1], bb[2], bb[3], ln.cnt, ln.scnt, A.bp);
                    split = !g.leaf;
                }
                if (b0) __syncthreads();   // S.shC of the previous chunk has been read by every warp
                u32 tot;
                const u32 sb2 = __ballot_sync(0xffffffffu, split);
                const u32 ex = warp_totals_scan(__popc(sb2), &tot, S.shC, lane, warp) + __popc(sb2 & lanemask_lt());
                if (k < nw) {
                    const int r = split ? nsplit2 + (int)ex : -1;
                    nxt[k].d.mid = g.axis ? g.x : g.y; nxt[k].d.rank = r; nxt[k].d.axis = g.axis;
                    SubNode& o = rec[base_id + k];
                    o.x = g.x; o.y = g.y; o.h = g.h; o.w = g.w;
                    o.first = f0 + ln.first; o.last = f0 + ln.first + ln.cnt;
                    o.sfirst = sf0 + ln.sfirst; o.slast = sf0 + ln.sfirst + ln.scnt;
                    o.ch1 = split ? base_id + nw + 2 * r : -1;
                    o.depth = rdepth + d + 1; o.axis = g.axis;
                }
                nsplit2 += (int)tot;
            }
            total += nw;
            d++;
            if (tid == 0) lvl_set(d + 1, total);
            cur = nxt; curbuf ^= 1; width = nw; nsplit = nsplit2;
            __syncthreads();
        }
        if (err) {
            if (tid == 0) A.st->err = 1;
            continue;
        }
        const int depth_l = d;   // levels 0 .. depth_l
        // ---- write the particles back: (x, y) from shared memory, the caller index gathered once (g follows it in
        // k_tree_gather_rest; here it is only needed for the centres of mass)
        {
            double gv[kRounds];
            int pv[kRounds];
#pragma unroll
            for (int i = 0; i < kRounds; i++) {
                const int p = (warp * kRounds + i) * 32 + lane;
                gv[i] = 0; pv[i] = 0;
                if (p < np) { pv[i] = A.perm[f0 + S.sidx[p]]; gv[i] = A.pg[pv[i]]; }   // g never moved: read it by caller index
            }
            int sv = 0;
            if (tid < ns) sv = S.sgi[S.sord[sb][tid]];
            __syncthreads();
            double* sg = reinterpret_cast<double*>(S.sidx);
#pragma unroll
            for (int i = 0; i < kRounds; i++) {
                const int p = (warp * kRounds + i) * 32 + lane;
                if (p < np) {
                    A.px[f0 + p] = S.sx[p]; A.py[f0 + p] = S.sy[p]; A.perm[f0 + p] = pv[i];
                    sg[p] = gv[i];
                }
            }
            if (tid < ns) A.segperm[sf0 + tid] = sv;
        }
        // ---- the subtree's own sweeps (CalculateCMass, :150-197; DFS ids)
        SweepNode* sw = (total <= kSweepSmemNodes) ? reinterpret_cast<SweepNode*>(S.tab) : g_sweep;
        for (int i = tid; i < total; i += kSubThreads) {
            const SubNode& o = rec[i];
            sw[i].ch1 = o.ch1; sw[i].first = o.first - f0; sw[i].cnt = o.last - o.first; sw[i].nx = o.x; sw[i].ny = o.y;
        }
        __syncthreads();
        {
            const double* sg = reinterpret_cast<const double*>(S.sidx);
            for (int dd = depth_l; dd >= 0; dd--) {
                const int b0 = lvl_get(dd), b1 = lvl_get(dd + 1);
                for (int i = b0 + tid; i < b1; i += kSubThreads) {
                    SweepNode& o = sw[i];
                    const int c = o.ch1;
                    if (c < 0) {
                        o.nl = 1; o.nn = 1;
                        leaf_cmass(S.sx, S.sy, sg, o.first, o.first + o.cnt, o.nx, o.ny, o.cmp, o.cmm);
                    } else {
                        o.nl = sw[c].nl + sw[c + 1].nl;
                        o.nn = 1 + sw[c].nn + sw[c + 1].nn;
                        child_cmass(sw[c].cmp, sw[c + 1].cmp, o.nx, o.ny, o.cmp);
                        child_cmass(sw[c].cmm, sw[c + 1].cmm, o.nx, o.ny, o.cmm);
                    }
                }
                __syncthreads();
            }
            for (int dd = 0; dd <= depth_l; dd++) {
                const int b0 = lvl_get(dd), b1 = lvl_get(dd + 1);
                for (int i = b0 + tid; i < b1; i += kSubThreads) {
                    SweepNode& o = sw[i];
                    if (i == 0) { o.lstart = 0; o.pre = 0; }
                    const int c = o.ch1;
                    if (c < 0) continue;
                    sw[c].lstart = o.lstart; sw[c + 1].lstart = o.lstart + sw[c].nl;
                    sw[c].pre = o.pre + 1; sw[c + 1].pre = o.pre + 1 + sw[c].nn;
                }
                __syncthreads();
            }
            for (int i = tid; i < total; i += kSubThreads) {
                SubNode& o = rec[i];
                const SweepNode& q = sw[i];
                for (int k = 0; k < 3; k++) { o.cmp[k] = q.cmp[k]; o.cmm[k] = q.cmm[k]; }
                o.nl = q.nl; o.nn = q.nn; o.lstart = q.lstart; o.pre = q.pre;
            }
            for (int dd = 1 + tid; dd <= depth_l; dd += kSubThreads)
                atomicAdd(&A.st->hist[rdepth + dd], lvl_get(dd + 1) - lvl_get(dd));
            if (tid == 0) atomicMax(&A.st->depth, rdepth + depth_l);
        }
        __syncthreads();
    }
}

// ================================================================================= top sweeps
struct SweepArgs {
    TreeDev T;
    const double *px, *py, *pg;
    const SubNode* scratch;
    const int* sublist;
    int *aux_sub, *aux_sn, *aux_nb;   // per top node: subtree index (ST_SUB); block nodes below it; id base of those blocks
    int *sub_base, *sub_lstart, *sub_pre;   // per subtree (sublist order)
    BuildState* st;
};

__global__ void __launch_bounds__(1024, 1) k_tree_topsweep(SweepArgs A) {
    const TreeDev& T = A.T;
    BuildState* st = A.st;
    if (st->err) return;
    const int tid = threadIdx.x;
    const int ntop = st->ntop, dtop = st->dtop, nsub = st->nsub;
    for (int s = tid; s < nsub; s += blockDim.x) A.aux_sub[A.sublist[s]] = s;
    __syncthreads();
    // bottom-up: subtree sizes, +/- centres of mass (CalculateCMass, :150-197)
    for (int dd = dtop; dd >= 0; dd--) {
        const int b0 = st->lvl[dd], b1 = st->lvl[dd + 1];
        for (int nn = b0 + tid; nn < b1; nn += blockDim.x) {
            double* P = T.cmp + 3ll * nn;
            double* M = T.cmm + 3ll * nn;
            const unsigned char s = T.status[nn];
            if (s == ST_LEAF) {
                T.nl[nn] = 1; T.nn[nn] = 1; A.aux_sn[nn] = 0;
                leaf_cmass(A.px, A.py, A.pg, T.first[nn], T.last[nn], T.x[nn], T.y[nn], P, M);
            } else if (s == ST_SUB) {
                const SubNode& r = A.scratch[2ll * ((long long)T.first[nn] + T.sfirst[nn])];
                T.nl[nn] = r.nl; T.nn[nn] = r.nn; A.aux_sn[nn] = r.nn - 1;
                for (int k = 0; k < 3; k++) { P[k] = r.cmp[k]; M[k] = r.cmm[k]; }
            } else {
                const int c = T.ch1[nn];
                T.nl[nn] = T.nl[c] + T.nl[c + 1];
                T.nn[nn] = 1 + T.nn[c] + T.nn[c + 1];
                A.aux_sn[nn] = A.aux_sn[c] + A.aux_sn[c + 1];
                child_cmass(T.cmp + 3ll * c, T.cmp + 3ll * (c + 1), T.x[nn], T.y[nn], P);
                child_cmass(T.cmm + 3ll * c, T.cmm + 3ll * (c + 1), T.x[nn], T.y[nn], M);
            }
        }
        __syncthreads();
    }
    // top-down: DFS leaf index, pre-order id, node-id base of the subtree blocks
    for (int dd = 0; dd <= dtop; dd++) {
        const int b0 = st->lvl[dd], b1 = st->lvl[dd + 1];
        for (int nn = b0 + tid; nn < b1; nn += blockDim.x) {
            if (nn == 0) { T.lstart[0] = 0; T.pre[0] = 0; A.aux_nb[0] = ntop; }
            const int ls = T.lstart[nn], pr = T.pre[nn], nb = A.aux_nb[nn];
            const unsigned char s = T.status[nn];
            if (s == ST_LEAF) {
                T.leaf_node[ls] = nn;
            } else if (s == ST_SUB) {
                const int si = A.aux_sub[nn];
                A.sub_base[si] = nb; A.sub_lstart[si] = ls; A.sub_pre[si] = pr;
                T.ch1[nn] = nb;   // local node 1 of the block
            } else {
                const int c = T.ch1[nn];
                T.lstart[c] = ls; T.pre[c] = pr + 1; A.aux_nb[c] = nb;
                T.lstart[c + 1] = ls + T.nl[c]; T.pre[c + 1] = pr + 1 + T.nn[c]; A.aux_nb[c + 1] = nb + A.aux_sn[c];
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        st->nnodes = ntop + A.aux_sn[0];
        st->nleaves = T.nl[0];
        if (dtop > st->depth) st->depth = dtop;
    }
    for (int dd = tid; dd <= dtop; dd += blockDim.x) atomicAdd(&st->hist[dd], st->lvl[dd + 1] - st->lvl[dd]);
}

// subtree blocks -> flat node arrays at their final ids (one CTA per subtree, grid-strided)
__global__ void __launch_bounds__(256) k_tree_relocate(TreeDev T, const SubNode* scratch, const int* sublist, const int* sub_base,
                                                       const int* sub_lstart, const int* sub_pre, const BuildState* st) {
    if (st->err) return;
    const int nsub = st->nsub;
    for (int s = blockIdx.x; s < nsub; s += gridDim.x) {
        const int root = sublist[s];
        const SubNode* rec = scratch + 2ll * ((long long)T.first[root] + T.sfirst[root]);
        const int cnt = rec[0].nn - 1, base = sub_base[s], ls0 = sub_lstart[s], pr0 = sub_pre[s];
        for (int i = 1 + threadIdx.x; i <= cnt; i += blockDim.x) {
            const SubNode o = rec[i];
            const int id = base + i - 1;
            T.x[id] = o.x; T.y[id] = o.y; T.h[id] = o.h; T.w[id] = o.w;
            for (int k = 0; k < 3; k++) { T.cmp[3ll * id + k] = o.cmp[k]; T.cmm[3ll * id + k] = o.cmm[k]; }
            T.first[id] = o.first; T.last[id] = o.last; T.sfirst[id] = o.sfirst; T.slast[id] = o.slast;
            T.ch1[id] = (o.ch1 < 0) ? -1 : base + o.ch1 - 1;
            T.depth[id] = o.depth; T.axis[id] = (unsigned char)o.axis;
            T.status[id] = (o.ch1 < 0) ? ST_LEAF : ST_SPLIT;
            T.lstart[id] = ls0 + o.lstart; T.pre[id] = pr0 + o.pre;
            if (o.ch1 < 0) T.leaf_node[ls0 + o.lstart] = id;
        }
    }
}

}  // namespace vv
