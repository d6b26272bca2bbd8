// K2 — near/far classification (snode::FindNearNodes, libvvhd/src/TSortedTree.cpp:199-217).
//
// The reference walks the tree once per leaf and stores two pointer vectors per leaf. Here the tree is
// walked ONCE for a GROUP of 32 consecutive leaves (consecutive in DFS order = contiguous in the
// permuted particle array): lanes hold frontier nodes, each lane tests its node against the group's
// leaves with the reference's exact criterion and produces a 32-bit mask of leaves for which the
// node is far / still near. Outputs of the single walk:
//   * the near list of the group: (source leaf, mask of target leaves that see it) entries, appended
//     to an entry pool in chunks (TravOut::unit entries). A chunk IS a work unit of the near-field
//     kernels, so no counting pass is needed; chunks are claimed with one atomic add each (their
//     placement in the pool is arbitrary, their content and their order within the group are
//     deterministic);
//   * per leaf, the four far-field Taylor coefficients of MConvectiveFast.cpp:48-69, accumulated on
//     the fly, so far lists are never stored.
// Who walks: k_traverse_cta<0> — one CTA of kTcWarps warps per group over a shared frontier in
// lock-step (a one-warp walk is ~150 dependent iterations and the kernel ended with its longest one).
// A few fringe groups see (almost) the whole tree; a walk that exceeds its node budget gives up and
// the group is redone in two steps: k_traverse<1> (one warp) walks the top levels and turns every
// still-near node of the cut level into an ITEM (node, mask); k_traverse_cta<2> walks each item's
// subtree. Items own their chunks and Taylor partials, which are combined in item order.
// k_traverse<0> / <2> are the one-warp forms of the same walks; they are not launched any more and
// stay as the plain statement of the algorithm that the CTA form parallelises.
// k_lists_dfs is the literal per-leaf walk, used only to export the reference's own lists
// (vvgpu_tree_lists) to host code and tests.
#pragma once
#include "vvgpu_tree.cuh"

namespace vv {

constexpr int kGroupLeaves = 32;
#ifndef VV_UNIT_ENTRIES
#define VV_UNIT_ENTRIES 768
#endif
constexpr int kUnitEntries = VV_UNIT_ENTRIES;   // list entries per chunk = per work unit of the near-field kernels
constexpr int kTravWarps = 4;       // warps per CTA in k_traverse
#ifndef VV_TRAV_BUDGET
#define VV_TRAV_BUDGET 256
#endif
constexpr int kTravBudget = VV_TRAV_BUDGET;    // warp iterations before a group is declared heavy (the mean is ~150)
constexpr int kTravStack = 1024;    // stack entries per warp (shared memory)
constexpr int kGroupSlots = 16;     // chunks a regular group may fill before it is declared heavy
constexpr int kItemSlots = 64;      // chunks one item of a heavy group may fill

// Target sharding over ranks (vvgpu_shard.cuh): pieces of kShardBlock consecutive groups are dealt round-robin
#ifndef VV_SHARD_BLOCK
#define VV_SHARD_BLOCK 1
#endif
constexpr int kShardBlock = VV_SHARD_BLOCK;   // groups per piece
struct Shard {
    int rank, nranks;
    // m-th group owned by this rank -> global group index (identity for one rank)
    __host__ __device__ __forceinline__ int group(int m) const {
        return ((m / kShardBlock) * nranks + rank) * kShardBlock + (m % kShardBlock);
    }
};

struct GroupLists {   // the entry pool
    int* leaf;        // source leaf index
    u32* mask;        // target leaves (bit k = leaf group*32 + k) that have `leaf` in NearNodes
};

// where a walk puts its chunks: slot s of the walk -> (base, count) of its s-th chunk
struct TravOut {
    GroupLists G;
    unsigned long long* cursor;   // next free pool entry
    long long pool_cap;
    long long* slot_base;
    int* slot_count;
    int* err;                     // bit 0: pool full, bit 1: stack overflow, bit 2: slot overflow in an item
    int unit;                     // entries per chunk (<= kUnitEntries, the capacity of the near kernels' tables):
                                  // small problems use shorter chunks so that there are enough work units to fill the GPU
};

// far iff dr.abs2() > farCriteria*HalfPerim*HalfPerim, HalfPerim = top.h + top.w + h + w
// evaluated left to right (TSortedTree.cpp:201-204)
__device__ __forceinline__ bool is_far(double nx, double ny, double nhw, double cx, double cy, double h, double w,
                                       double farc) {
    double drx = VV_SUB(nx, cx), dry = VV_SUB(ny, cy);
    double hp = VV_ADD(VV_ADD(nhw, h), w);
    double d2 = VV_ADD(VV_MUL(drx, drx), VV_MUL(dry, dry));
    return d2 > VV_MUL(VV_MUL(farc, hp), hp);
}

// one far node's contribution to a leaf's Taylor sums (MConvectiveFast.cpp:50-63); a monopole with
// g == 0 contributes exactly 0 in the reference and is skipped
__device__ __forceinline__ void taylor_add(double cx, double cy, double mx, double my, double mg, double& T1,
                                           double& T2, double& T3, double& T4) {
    if (mg == 0) return;
    double dx = cx - mx, dy = cy - my;
    double a = dx * dx + dy * dy;
    // 1/a by MUFU.RCP64H + two Newton steps (relative error ~2^-79 before rounding): the coefficients only
    // feed velocities (1e-10 bar), and two IEEE divisions per monopole were a third of this kernel
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    double f1 = mg * r;
    double f2 = f1 * r;
    T1 -= f1 * dy;
    T2 += f1 * dx;
    T3 += f2 * dy * dx;
    T4 += f2 * (dy * dy - dx * dx);
}

// MODE 0: a group from the root (budgeted; gives up -> heavy). MODE 1: top of a heavy group (nodes
// shallower than cut_depth; still-near nodes AT cut_depth become items). MODE 2: one item of a heavy group.
struct TravItems {
    int* node;        // [nheavy * item_cap]
    u32* mask;
    int* count;       // [nheavy]
    int item_cap;
    int cut_depth;
};

template <int MODE>
__global__ void __launch_bounds__(kTravWarps * 32)
k_traverse(TreeDev T, LeafDev L, int nleaves, int g0, int g1, double farc, TravOut O, double* taylor, double* farcount,
           double* tpart /* MODE 1,2: per walk 32 x 5 partials */, const int* heavy_list, int nheavy, int* heavy_out,
           int* nheavy_out, TravItems I) {
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * kTravWarps + warp;   // walk index
    constexpr size_t per_warp = (size_t)kTravStack * sizeof(int2) + 4 * 32 * sizeof(double) + 32 * 6 * sizeof(double);
    unsigned char* base = smem_raw + per_warp * warp;
    int2* stack = (int2*)base;
    double* lcx = (double*)(base + (size_t)kTravStack * sizeof(int2));
    double* lcy = lcx + 32;
    double* lh = lcy + 32;
    double* lw = lh + 32;
    double* cm = lw + 32;
    // ---- which walk is this
    int g, slot0, nslots, hidx = 0, seq = 0;
    int start_node = 0;
    u32 start_mask = 0;
    if (MODE == 0) {
        g = g0 + w;
        if (g >= g1) return;   // warps are independent: no block-level barrier below
        slot0 = g * kGroupSlots; nslots = kGroupSlots;
    } else if (MODE == 1) {
        hidx = w;
        if (hidx >= nheavy) return;
        g = heavy_list[hidx];
        seq = 0;
    } else {
        hidx = w / I.item_cap;
        if (hidx >= nheavy) return;
        const int it = w - hidx * I.item_cap;
        if (it >= I.count[hidx]) return;
        g = heavy_list[hidx];
        seq = it + 1;
        start_node = I.node[(long long)hidx * I.item_cap + it];
        start_mask = I.mask[(long long)hidx * I.item_cap + it];
    }
    long long walk = 0;   // index of this walk among the heavy walks (MODE 1, 2)
    if (MODE != 0) {
        walk = (long long)hidx * (I.item_cap + 1) + seq;
        slot0 = (int)(walk * kItemSlots); nslots = kItemSlots;   // (offset by the regular slots, added by the caller in O)
    }
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, nleaves - l0);
    double mycx = 0, mycy = 0;
    if (lane < nl) {
        mycx = L.cx[l0 + lane]; mycy = L.cy[l0 + lane];
        lcx[lane] = mycx; lcy[lane] = mycy; lh[lane] = L.h[l0 + lane]; lw[lane] = L.w[l0 + lane];
    } else { lcx[lane] = 0; lcy[lane] = 0; lh[lane] = 0; lw[lane] = 0; }
    const u32 full = (nl == 32) ? 0xffffffffu : ((1u << nl) - 1u);
    if (lane == 0) stack[0] = (MODE == 2) ? make_int2(start_node, (int)start_mask) : make_int2(0, (int)full);
    int size = 1;
    __syncwarp();
    double T1 = 0, T2 = 0, T3 = 0, T4 = 0;
    double nfar = 0;
    int iters = 0;
    // chunk bookkeeping (uniform across the warp)
    int nchunk = 0, fill = O.unit;   // "full": the first emission claims a chunk
    long long cbase = 0;
    int nitems = 0;
    bool bail = false;
    while (size > 0) {
        if (MODE == 0 && ++iters > kTravBudget) { bail = true; break; }
        // near the capacity fall back to plain depth-first order (growth <= 1 per pop)
        int take = (size > kTravStack - 80) ? 1 : min(size, 32);
        int n = -1;
        u32 em = 0;
        if (lane < take) { int2 e = stack[size - 1 - lane]; n = e.x; em = (u32)e.y; }
        size -= take;
        __syncwarp();
        u32 farm = 0, nearm = 0;
        int c1 = -1;
        if (n >= 0) {
            double nx = T.x[n], ny = T.y[n];
            double nhw = VV_ADD(T.h[n], T.w[n]);
            c1 = T.ch1[n];
            for (u32 mm = em; mm; mm &= mm - 1) {  // only the leaves that still see this subtree
                const int l = __ffs(mm) - 1;
                if (is_far(nx, ny, nhw, lcx[l], lcy[l], lh[l], lw[l], farc)) farm |= (1u << l);
            }
            nearm = em & ~farm;
            if (farm) {
                cm[lane * 6 + 0] = T.cmp[3ll * n + 0]; cm[lane * 6 + 1] = T.cmp[3ll * n + 1];
                cm[lane * 6 + 2] = T.cmp[3ll * n + 2]; cm[lane * 6 + 3] = T.cmm[3ll * n + 0];
                cm[lane * 6 + 4] = T.cmm[3ll * n + 1]; cm[lane * 6 + 5] = T.cmm[3ll * n + 2];
            }
        }
        // descend: both children inherit the still-near leaves (child 1 ends up on top). In the top
        // walk of a heavy group the children at the cut depth become items instead.
        bool push = (n >= 0) && nearm && (c1 >= 0);
        bool cut = false;
        if (MODE == 1) cut = push && (T.depth[n] + 1 >= I.cut_depth);
        const u32 pb = __ballot_sync(0xffffffffu, push && !cut);
        const int npush = 2 * __popc(pb);
        if (size + npush > kTravStack) {
            if (lane == 0) atomicOr(O.err, 2);
            return;
        }
        if (push && !cut) {
            int off = size + 2 * __popc(pb & lanemask_lt());
            stack[off] = make_int2(c1 + 1, (int)nearm);
            stack[off + 1] = make_int2(c1, (int)nearm);
        }
        size += npush;
        if (MODE == 1) {
            const u32 cb = __ballot_sync(0xffffffffu, cut);
            if (cut) {
                const long long at = (long long)hidx * I.item_cap + nitems + 2 * __popc(cb & lanemask_lt());
                I.node[at] = c1; I.mask[at] = nearm;
                I.node[at + 1] = c1 + 1; I.mask[at + 1] = nearm;
            }
            nitems += 2 * __popc(cb);
        }
        // a near leaf: one list entry for the group
        const bool emit = (n >= 0) && nearm && (c1 < 0);
        const u32 eb = __ballot_sync(0xffffffffu, emit);
        if (eb) {
            const int k = __popc(eb);
            int pos = fill + __popc(eb & lanemask_lt());
            long long nbase = cbase;
            if (fill + k > O.unit) {   // (part of) this batch goes to a fresh chunk
                if (nchunk >= nslots) {
                    if (MODE == 0) { bail = true; break; }
                    if (lane == 0) atomicOr(O.err, 4);
                    return;
                }
                unsigned long long nb = 0;
                if (lane == 0) {
                    nb = atomicAdd(O.cursor, (unsigned long long)O.unit);
                    if (nchunk > 0) O.slot_count[slot0 + nchunk - 1] = O.unit;
                    O.slot_base[slot0 + nchunk] = (long long)nb;
                }
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if ((long long)nb + O.unit > O.pool_cap) {
                    if (lane == 0) atomicOr(O.err, 1);
                    return;
                }
                nbase = (long long)nb;
                nchunk++;
            }
            if (emit) {
                const long long at = (pos < O.unit) ? (cbase + pos) : (nbase + (pos - O.unit));
                O.G.leaf[at] = T.lstart[n];
                O.G.mask[at] = nearm;
            }
            if (fill + k > O.unit) { fill = fill + k - O.unit; cbase = nbase; }
            else fill += k;
        }
        // far nodes: transpose (node lane x leaf bit) -> (leaf lane x node bit) and accumulate
        const u32 anyfar = __ballot_sync(0xffffffffu, farm != 0);
        if (anyfar) {
            __syncwarp();
            u32 mine = 0;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                u32 tm = __ballot_sync(0xffffffffu, (farm >> l) & 1u);
                if (lane == l) mine = tm;
            }
            nfar += (double)__popc(mine);
            while (mine) {
                int j = __ffs(mine) - 1;
                mine &= mine - 1;
                const double* c = cm + j * 6;
                taylor_add(mycx, mycy, c[0], c[1], c[2], T1, T2, T3, T4);
                taylor_add(mycx, mycy, c[3], c[4], c[5], T1, T2, T3, T4);
            }
        }
        __syncwarp();
    }
    if (MODE == 0 && bail) {
        // a fringe group that sees most of the tree: a single warp would serialise the whole step behind
        // it. Its chunks are dropped (count 0) and the group is redone as a heavy one.
        if (lane == 0) {
            heavy_out[atomicAdd(nheavy_out, 1)] = g;
            for (int k = 0; k < min(nchunk, nslots); k++) O.slot_count[slot0 + k] = 0;
        }
        return;
    }
    if (lane == 0 && nchunk > 0) O.slot_count[slot0 + nchunk - 1] = fill;
    if (MODE == 0) {
        if (lane < nl) {
            double* t = taylor + 4ll * (l0 + lane);
            t[0] = T1 * k1_2Pi; t[1] = T2 * k1_2Pi; t[2] = T3 * k1_Pi; t[3] = T4 * k1_2Pi;  // :66-69
            if (farcount) farcount[l0 + lane] = nfar;
        }
    } else {
        double* t = tpart + (walk * 32 + lane) * 5;
        t[0] = T1; t[1] = T2; t[2] = T3; t[3] = T4; t[4] = nfar;
        if (MODE == 1 && lane == 0) I.count[hidx] = nitems;
    }
}

// The regular walk, one CTA per group. A single warp's walk is ~150 dependent iterations (up to the budget of
// 256), so a kernel of one-warp walks ends with its LONGEST walk: 1.7 ms at N = 1M for 0.6 ms of FP64 work, and
// no faster when the groups are sharded over GPUs. Here kTcWarps warps pop up to 32 * kTcWarps frontier nodes per
// iteration from one shared stack and advance in lock-step; offsets of the pushed children and of the emitted
// list entries come from a cross-warp prefix of the per-warp counts, so the result is deterministic. Every
// warp keeps Taylor partials for the 32 leaves (lane = leaf); they are summed in warp order at the end.
#ifndef VV_TC_WARPS
#define VV_TC_WARPS 4
#endif
constexpr int kTcWarps = VV_TC_WARPS;
constexpr int kTcStack = 2048;
constexpr int kTcBudgetNodes = kTravBudget * 32;   // the same bound on visited nodes as the one-warp walk

// MODE 0: a group from the root (budgeted; gives up -> heavy). MODE 2: one item of a heavy group (k_traverse<1>
// made the items), blockIdx.x = heavy index * item_cap + item.
template <int MODE>
__global__ void __launch_bounds__(kTcWarps * 32)
k_traverse_cta(TreeDev T, LeafDev L, int nleaves, Shard sh, int ngmine, double farc, TravOut O, double* taylor, double* farcount,
               int* heavy_out, int* nheavy_out, double* tpart, const int* heavy_list, int nheavy, TravItems I) {
    __shared__ int2 stack[kTcStack];
    __shared__ double lcx[32], lcy[32], lh[32], lw[32];
    __shared__ double cmw[kTcWarps][32 * 6];
    __shared__ double tsum[kTcWarps][32][5];
    __shared__ int wpush[kTcWarps], wemit[kTcWarps];
    __shared__ long long s_nbase;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int g, slot0, nslots;
    int start_node = 0;
    u32 start_mask = 0;
    long long walk = 0;
    if (MODE == 0) {
        if ((int)blockIdx.x >= ngmine) return;
        g = sh.group(blockIdx.x);
        slot0 = g * kGroupSlots; nslots = kGroupSlots;
    } else {
        const int hidx = blockIdx.x / I.item_cap;
        const int it = blockIdx.x - hidx * I.item_cap;
        if (hidx >= nheavy || it >= I.count[hidx]) return;
        g = heavy_list[hidx];
        start_node = I.node[(long long)hidx * I.item_cap + it];
        start_mask = I.mask[(long long)hidx * I.item_cap + it];
        walk = (long long)hidx * (I.item_cap + 1) + it + 1;
        slot0 = (int)(walk * kItemSlots); nslots = kItemSlots;   // (offset by the regular slots, added by the caller in O)
    }
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, nleaves - l0);
    if (warp == 0) {
        if (lane < nl) { lcx[lane] = L.cx[l0 + lane]; lcy[lane] = L.cy[l0 + lane]; lh[lane] = L.h[l0 + lane]; lw[lane] = L.w[l0 + lane]; }
        else { lcx[lane] = 0; lcy[lane] = 0; lh[lane] = 0; lw[lane] = 0; }
    }
    const u32 full = (nl == 32) ? 0xffffffffu : ((1u << nl) - 1u);
    if (tid == 0) stack[0] = (MODE == 2) ? make_int2(start_node, (int)start_mask) : make_int2(0, (int)full);
    __syncthreads();
    const double mycx = lcx[lane], mycy = lcy[lane];
    double* cm = cmw[warp];
    int size = 1;
    double T1 = 0, T2 = 0, T3 = 0, T4 = 0, nfar = 0;
    int nchunk = 0, fill = O.unit;   // uniform across the CTA; "full": the first emission claims a chunk
    long long cbase = 0;
    int visited = 0;
    bool bail = false;
    while (size > 0) {
        if (MODE == 0 && visited > kTcBudgetNodes) { bail = true; break; }
        // near the capacity fall back to plain depth-first order (growth <= 1 per pop)
        const int take = (size > kTcStack - 2 * 32 * kTcWarps - 8) ? 1 : min(size, 32 * kTcWarps);
        visited += take;
        const int k = warp * 32 + lane;
        int n = -1;
        u32 em = 0;
        if (k < take) { const int2 e = stack[size - 1 - k]; n = e.x; em = (u32)e.y; }
        const int base = size - take;
        __syncthreads();   // every pop is read before a push may overwrite it
        u32 farm = 0, nearm = 0;
        int c1 = -1;
        if (n >= 0) {
            const double nx = T.x[n], ny = T.y[n];
            const double nhw = VV_ADD(T.h[n], T.w[n]);
            c1 = T.ch1[n];
            for (u32 mm = em; mm; mm &= mm - 1) {  // only the leaves that still see this subtree
                const int l = __ffs(mm) - 1;
                if (is_far(nx, ny, nhw, lcx[l], lcy[l], lh[l], lw[l], farc)) farm |= (1u << l);
            }
            nearm = em & ~farm;
            if (farm) {
                cm[lane * 6 + 0] = T.cmp[3ll * n + 0]; cm[lane * 6 + 1] = T.cmp[3ll * n + 1];
                cm[lane * 6 + 2] = T.cmp[3ll * n + 2]; cm[lane * 6 + 3] = T.cmm[3ll * n + 0];
                cm[lane * 6 + 4] = T.cmm[3ll * n + 1]; cm[lane * 6 + 5] = T.cmm[3ll * n + 2];
            }
        }
        const bool push = (n >= 0) && nearm && (c1 >= 0);
        const bool emit = (n >= 0) && nearm && (c1 < 0);
        const u32 pb = __ballot_sync(0xffffffffu, push);
        const u32 eb = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) { wpush[warp] = 2 * __popc(pb); wemit[warp] = __popc(eb); }
        __syncthreads();
        int pbefore = 0, ptotal = 0, ebefore = 0, etotal = 0;
#pragma unroll
        for (int w = 0; w < kTcWarps; w++) {
            const int a = wpush[w], b = wemit[w];
            if (w < warp) { pbefore += a; ebefore += b; }
            ptotal += a; etotal += b;
        }
        if (base + ptotal > kTcStack) {
            if (tid == 0) atomicOr(O.err, 2);
            return;
        }
        // descend: both children inherit the still-near leaves (child 1 ends up on top)
        if (push) {
            const int off = base + pbefore + 2 * __popc(pb & lanemask_lt());
            stack[off] = make_int2(c1 + 1, (int)nearm);
            stack[off + 1] = make_int2(c1, (int)nearm);
        }
        size = base + ptotal;
        // near leaves: one list entry each for the group
        if (etotal) {
            const int pos = fill + ebefore + __popc(eb & lanemask_lt());
            long long nbase = cbase;
            const bool cross = fill + etotal > O.unit;   // (part of) this batch goes to a fresh chunk
            if (cross) {
                if (nchunk >= nslots) {
                    if (MODE == 0) { bail = true; break; }
                    if (tid == 0) atomicOr(O.err, 4);
                    return;
                }
                if (tid == 0) {
                    const unsigned long long nb = atomicAdd(O.cursor, (unsigned long long)O.unit);
                    if (nchunk > 0) O.slot_count[slot0 + nchunk - 1] = O.unit;
                    O.slot_base[slot0 + nchunk] = (long long)nb;
                    s_nbase = (long long)nb;
                }
                __syncthreads();
                nbase = s_nbase;
                if (nbase + O.unit > O.pool_cap) {
                    if (tid == 0) atomicOr(O.err, 1);
                    return;
                }
                nchunk++;
            }
            if (emit) {
                const long long at = (pos < O.unit) ? (cbase + pos) : (nbase + (pos - O.unit));
                O.G.leaf[at] = T.lstart[n];
                O.G.mask[at] = nearm;
            }
            if (cross) { fill = fill + etotal - O.unit; cbase = nbase; }
            else fill += etotal;
        }
        // far nodes of this warp: transpose (node lane x leaf bit) -> (leaf lane x node bit) and accumulate
        const u32 anyfar = __ballot_sync(0xffffffffu, farm != 0);
        if (anyfar) {
            __syncwarp();
            u32 mine = 0;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                const u32 tm = __ballot_sync(0xffffffffu, (farm >> l) & 1u);
                if (lane == l) mine = tm;
            }
            nfar += (double)__popc(mine);
            while (mine) {
                const int j = __ffs(mine) - 1;
                mine &= mine - 1;
                const double* c = cm + j * 6;
                taylor_add(mycx, mycy, c[0], c[1], c[2], T1, T2, T3, T4);
                taylor_add(mycx, mycy, c[3], c[4], c[5], T1, T2, T3, T4);
            }
        }
        __syncthreads();   // pushes visible, counters and s_nbase reusable
    }
    if (bail) {
        // a fringe group that sees most of the tree: its chunks are dropped (count 0) and the group is redone
        // as a heavy one (per-subtree items)
        if (tid == 0) {
            heavy_out[atomicAdd(nheavy_out, 1)] = g;
            for (int q = 0; q < min(nchunk, nslots); q++) O.slot_count[slot0 + q] = 0;
        }
        return;
    }
    if (tid == 0 && nchunk > 0) O.slot_count[slot0 + nchunk - 1] = fill;
    tsum[warp][lane][0] = T1; tsum[warp][lane][1] = T2; tsum[warp][lane][2] = T3; tsum[warp][lane][3] = T4; tsum[warp][lane][4] = nfar;
    __syncthreads();
    if (warp == 0) {
        double s5[5] = {0, 0, 0, 0, 0};
        for (int w = 0; w < kTcWarps; w++)
            for (int q = 0; q < 5; q++) s5[q] += tsum[w][lane][q];
        if (MODE == 0) {
            if (lane < nl) {
                double* t = taylor + 4ll * (l0 + lane);
                t[0] = s5[0] * k1_2Pi; t[1] = s5[1] * k1_2Pi; t[2] = s5[2] * k1_Pi; t[3] = s5[3] * k1_2Pi;  // :66-69
                if (farcount) farcount[l0 + lane] = s5[4];
            }
        } else {
            double* t = tpart + (walk * 32 + lane) * 5;
            for (int q = 0; q < 5; q++) t[q] = s5[q];
        }
    }
}

// Taylor coefficients of the leaves of heavy groups: partials of the top walk and of the items, in order
__global__ void __launch_bounds__(256) k_heavy_taylor(const int* heavy_list, int nheavy, int nleaves, const double* tpart,
                                                      const int* item_count, int item_cap, double* taylor, double* farcount) {
    __shared__ double part[8][32][5];
    const int hidx = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (hidx >= nheavy) return;
    const int l = heavy_list[hidx] * kGroupLeaves + lane;
    double s[5] = {0, 0, 0, 0, 0};
    const int nw = item_count[hidx] + 1;
    for (int q = warp; q < nw; q += 8) {   // fixed assignment of walks to warps: deterministic
        const double* t = tpart + (((long long)hidx * (item_cap + 1) + q) * 32 + lane) * 5;
        for (int k = 0; k < 5; k++) s[k] += t[k];
    }
    for (int k = 0; k < 5; k++) part[warp][lane][k] = s[k];
    __syncthreads();
    if (warp != 0 || l >= nleaves) return;
    for (int k = 0; k < 5; k++) s[k] = 0;
    for (int w = 0; w < 8; w++)
        for (int k = 0; k < 5; k++) s[k] += part[w][lane][k];
    double* t = taylor + 4ll * l;
    t[0] = s[0] * k1_2Pi; t[1] = s[1] * k1_2Pi; t[2] = s[2] * k1_Pi; t[3] = s[3] * k1_2Pi;
    if (farcount) farcount[l] = s[4];
}

// The walks of a heavy group leave many short chunks (one tail per item). Pack them, in walk order, into
// dense chunks of a fresh pool region, so that the near-field kernels see full work units. One CTA per
// heavy group; `off` is scratch of the same shape as the group's slot range.
__global__ void __launch_bounds__(1024) k_heavy_pack(int nheavy, int item_cap, TravOut O, int* off) {
    const int h = blockIdx.x;
    if (h >= nheavy) return;
    __shared__ u32 sh[1024 / 32 + 1];
    __shared__ long long region;
    __shared__ int total_s, nlive_s;
    const long long per = (long long)(item_cap + 1) * kItemSlots;
    long long* sbase = O.slot_base + h * per;
    int* scount = O.slot_count + h * per;
    int* doff = off + h * per;                              // dense offset of every slot
    int* live = off + (long long)nheavy * per + h * per;    // the slots that hold entries, in order
    // a walk claims its slots in order, so the slots that hold entries are a prefix of its kItemSlots: one scan over
    // the WALKS (a thread sums its walk's few used slots) instead of one over all their mostly empty slots
    const int nwalk = item_cap + 1;
    u32 carry = 0, lcarry = 0;
    for (int b = 0; b < nwalk; b += 1024) {
        const int wk = b + threadIdx.x;
        const int* sc = scount + (long long)wk * kItemSlots;
        u32 v = 0, nlv = 0;
        if (wk < nwalk)
            for (int q = 0; q < kItemSlots; q++) {
                const int cnt = sc[q];
                if (cnt <= 0) break;
                v += (u32)cnt; nlv++;
            }
        u32 total, ltotal;
        const u32 ex = block_exclusive_scan<1024>(v, &total, sh);
        const u32 lex = block_exclusive_scan<1024>(nlv, &ltotal, sh);
        u32 o = carry + ex;
        for (u32 q = 0; q < nlv; q++) {
            const long long sl = (long long)wk * kItemSlots + q;
            doff[sl] = (int)o;
            live[lcarry + lex + q] = (int)sl;
            o += (u32)sc[q];
        }
        carry += total; lcarry += ltotal;
    }
    if (threadIdx.x == 0) {
        nlive_s = (int)lcarry;
        total_s = (int)carry;
        const unsigned long long need = ((unsigned long long)carry + O.unit - 1) / O.unit * O.unit;
        const unsigned long long at = atomicAdd(O.cursor, need);
        region = (long long)at;
        if ((long long)(at + need) > O.pool_cap) { atomicOr(O.err, 1); region = -1; }
    }
    __syncthreads();
    if (region < 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int q = warp; q < nlive_s; q += 32) {
        const long long s = live[q];
        const int n = scount[s];
        const long long src = sbase[s], dst = region + doff[s];
        for (int k = lane; k < n; k += 32) { O.G.leaf[dst + k] = O.G.leaf[src + k]; O.G.mask[dst + k] = O.G.mask[src + k]; }
    }
    __syncthreads();
    const int total = total_s;
    for (long long s = threadIdx.x; s < per; s += 1024) {
        const long long first = s * O.unit;
        const bool used = first < total;
        sbase[s] = used ? region + first : 0;
        scount[s] = used ? (int)min((long long)O.unit, total - first) : 0;
    }
}

// ---- units: the non-empty chunk slots, in slot order (regular groups first, then heavy walks) ----------
struct SlotFlag {
    const int* count;
    __device__ __forceinline__ u32 operator()(long long i) const { return count[i] > 0 ? 1u : 0u; }
};
// slot -> group: regular slots are group-major; heavy slots are (heavy group, walk)-major
__global__ void k_units_fill(long long nslots_total, long long nreg_slots, const u32* __restrict__ rank,
                             const long long* __restrict__ slot_base, const int* __restrict__ slot_count,
                             const int* __restrict__ heavy_list, int item_cap, int* ugroup, long long* ubase, int* ucount) {
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots_total || slot_count[s] <= 0) return;
    int g;
    if (s < nreg_slots) g = (int)(s / kGroupSlots);
    else g = heavy_list[(s - nreg_slots) / ((long long)(item_cap + 1) * kItemSlots)];
    const u32 u = rank[s];
    ugroup[u] = g; ubase[u] = slot_base[s]; ucount[u] = slot_count[s];
}
// group -> its unit range. Regular: from the ranks at its first / one-past-last slot.
__global__ void k_units_groups(int ngroups, const u32* __restrict__ rank, int* ufirst, int* unum) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    const u32 a = rank[(long long)g * kGroupSlots], b = rank[(long long)(g + 1) * kGroupSlots];
    ufirst[g] = (int)a; unum[g] = (int)(b - a);
}
__global__ void k_units_heavy(int nheavy, long long nreg_slots, int item_cap, const int* __restrict__ heavy_list,
                              const u32* __restrict__ rank, int* ufirst, int* unum) {
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nheavy) return;
    const long long per = (long long)(item_cap + 1) * kItemSlots;
    const u32 a = rank[nreg_slots + h * per], b = rank[nreg_slots + (h + 1) * per];
    ufirst[heavy_list[h]] = (int)a; unum[heavy_list[h]] = (int)(b - a);
}

// ---- literal per-leaf walk, for export only ---------------------------------------------------
constexpr int kDfsStack = 256;
template <bool FILL>
__global__ void k_lists_dfs(TreeDev T, LeafDev L, int nleaves, double farc, u32* ncount, u32* fcount,
                            const u32* nptr, const u32* fptr, long long* nidx, long long* fidx, int* err) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nleaves) return;
    double cx = L.cx[l], cy = L.cy[l], h = L.h[l], w = L.w[l];
    int stack[kDfsStack];
    int size = 0;
    stack[size++] = 0;
    u32 nn = 0, nf = 0;
    long long nb = FILL ? nptr[l] : 0, fb = FILL ? fptr[l] : 0;
    while (size > 0) {
        int n = stack[--size];
        if (is_far(T.x[n], T.y[n], VV_ADD(T.h[n], T.w[n]), cx, cy, h, w, farc)) {
            if (FILL) fidx[fb + nf] = T.pre[n];
            nf++;
            continue;
        }
        int c = T.ch1[n];
        if (c >= 0) {
            if (size + 2 > kDfsStack) { atomicExch(err, 1); return; }
            stack[size++] = c + 1;
            stack[size++] = c;
            continue;
        }
        if (FILL) nidx[nb + nn] = T.lstart[n];
        nn++;
    }
    if (!FILL) { ncount[l] = nn; fcount[l] = nf; }
}

// ---- CTA table of the near-field kernels ---------------------------------------------------------------------
// A work unit's cost is (about) its pair count: (#targets of its group) x (#particles of its entries' source leaves).
// Estimated without reading the entries as entries x T x T / 32, T = particles of the group's 32 leaves: the source
// leaves are the neighbourhood of the targets and about as dense. Where the minimum node size stops the bisection (dense
// regions next to a body), leaves hold hundreds of particles and ONE unit can cost a hundred average ones: the kernel then ends with a few
// CTAs on a mostly idle GPU (cylinder wake, N = 1M: SMs busy 42 % of k_conv's duration). So units that cost more than a
// third of an SM slot's fair share are split over 2 or 4 CTAs, each taking 16 or 8 of the unit's 32 target leaves against
// the whole entry table, and their CTAs start first. What a target sums, and in which order, does not depend on it.
// cta -> (unit, first and one-past-last target leaf): the units that are split further than `f0` first, then the others
// in unit order. out[0] = number of CTAs, out[1] = how many of them belong to the first class, out[2] = *carry (a value
// the host wants with the same read-back).
__device__ __forceinline__ unsigned long long unit_cost(const LeafDev& L, int nleaves, int g, int entries) {
    const int l0 = g * kGroupLeaves, l1 = min(l0 + kGroupLeaves, nleaves) - 1;
    const unsigned long long t = (unsigned long long)(L.last[l1] - L.first[l0]);
    return (unsigned long long)entries * t * t;
}
__global__ void __launch_bounds__(1024) k_cta_table(int nunits, LeafDev L, int nleaves, const int* __restrict__ ugroup,
                                                    const int* __restrict__ ucount, int slots, int f0, int fmax, int div,
                                                    int* cta_unit, unsigned short* cta_part, const u32* carry, u32* out) {
    __shared__ u32 sh[1024 / 32 + 1];
    __shared__ unsigned long long tsum[32];
    __shared__ unsigned long long total_s;
    unsigned long long s = 0;
    for (int u = threadIdx.x; u < nunits; u += 1024) s += unit_cost(L, nleaves, ugroup[u], ucount[u]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) tsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 32; k++) t += tsum[k];
        total_s = t;
    }
    __syncthreads();
    const unsigned long long thr = max(1ull, total_s / ((unsigned long long)div * (unsigned long long)slots));
    u32 base = 0;
    for (int pass = 0; pass < 2; pass++) {   // 0: the units split further than f0, 1: the others
        for (int b = 0; b < nunits; b += 1024) {
            const int u = b + threadIdx.x;
            int f = 0;
            if (u < nunits) {
                f = f0;
                const unsigned long long cu = unit_cost(L, nleaves, ugroup[u], ucount[u]);
                while (f < fmax && cu > thr * (unsigned long long)f) f *= 2;
                if ((f > f0) != (pass == 0)) f = 0;
            }
            u32 total;
            const u32 ex = block_exclusive_scan<1024>((u32)f, &total, sh);
            const int per = f ? kGroupLeaves / f : 0;
            for (int q = 0; q < f; q++) {
                cta_unit[base + ex + q] = u;
                cta_part[base + ex + q] = (unsigned short)((q * per) | (((q + 1) * per) << 8));
            }
            base += total;
        }
        if (pass == 0 && threadIdx.x == 0) out[1] = base;
    }
    if (threadIdx.x == 0) { out[0] = base; out[2] = carry ? *carry : 0u; }
}

// near-pair count of one work unit: sum over its entries of (#targets g!=0 in masked leaves) x (#sources g!=0)
__global__ void k_count_pairs(LeafDev L, int nleaves, int nunits, const int* __restrict__ ugroup,
                              const long long* __restrict__ ubase, const int* __restrict__ ucount, GroupLists G,
                              const double* __restrict__ pg, double* out) {
    int u = blockIdx.x;
    if (u >= nunits) return;
    const int g = ugroup[u];
    __shared__ int nz[kGroupLeaves];
    __shared__ double acc[32];
    int l0 = g * kGroupLeaves, nl = min(kGroupLeaves, nleaves - l0);
    if (threadIdx.x < kGroupLeaves) {
        int k = 0;
        if ((int)threadIdx.x < nl)
            for (int i = L.first[l0 + threadIdx.x]; i < L.last[l0 + threadIdx.x]; i++) k += (pg[i] != 0);
        nz[threadIdx.x] = k;
    }
    __syncthreads();
    double s = 0;
    const long long e0 = ubase[u], e1 = e0 + ucount[u];
    for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        int sl = G.leaf[e];
        u32 m = G.mask[e];
        int ns = 0;
        for (int i = L.first[sl]; i < L.last[sl]; i++) ns += (pg[i] != 0);
        int nt = 0;
        while (m) { int b = __ffs(m) - 1; m &= m - 1; nt += nz[b]; }
        s += (double)ns * (double)nt;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) acc[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += acc[k];
        out[u] = t;
    }
}

}  // namespace vv
