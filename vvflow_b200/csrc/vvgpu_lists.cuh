// K2 — near/far classification (snode::FindNearNodes, libvvhd/src/TSortedTree.cpp:199-217).
//
// The reference walks the tree once per leaf and stores two pointer vectors per leaf. Here one
// warp walks the tree once for a GROUP of 32 consecutive leaves (consecutive in DFS order =
// contiguous in the permuted particle array): lanes hold up to 32 frontier nodes, each lane
// tests its node against all 32 leaves with the reference's exact criterion and produces a
// 32-bit mask of leaves for which the node is far / still near. Outputs:
//   * per group, the list of near source leaves with the mask of target leaves that see them
//     (two passes: count, then fill) — the only interaction list that is materialised;
//   * per leaf, the four far-field Taylor coefficients of MConvectiveFast.cpp:48-69, accumulated
//     on the fly, so far lists are never stored.
// k_lists_dfs is the literal per-leaf walk, used only to export the reference's own lists
// (vvgpu_tree_lists) to host code and tests.
#pragma once
#include "vvgpu_tree.cuh"

namespace vv {

constexpr int kGroupLeaves = 32;
constexpr int kTravWarps = 4;  // warps per CTA in k_traverse
constexpr int kTravBudget = 512;  // warp iterations before a group is declared heavy
constexpr int kHeavyThreads = 1024;

struct GroupLists {
    long long* ptr;  // ngroups + 1
    int* leaf;       // source leaf index
    u32* mask;       // target leaves (bit k = leaf group*32 + k) that have `leaf` in NearNodes
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
    double f1 = mg / a;
    double f2 = f1 / a;
    T1 -= f1 * dy;
    T2 += f1 * dx;
    T3 += f2 * dy * dx;
    T4 += f2 * (dy * dy - dx * dx);
}

template <bool FILL>
__global__ void __launch_bounds__(kTravWarps * 32)
k_traverse(TreeDev T, LeafDev L, int nleaves, int ngroups, double farc, GroupLists G, u32* gcount, double* taylor,
           double* farcount, int stack_cap, int* err, int* heavy, int* nheavy, const unsigned char* is_heavy) {
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * kTravWarps + warp;
    // per-warp shared layout: int2 stack[stack_cap]; double lc[4][32]; double cm[32][6]
    size_t per_warp = (size_t)stack_cap * sizeof(int2) + 4 * 32 * sizeof(double) + 32 * 6 * sizeof(double);
    unsigned char* base = smem_raw + per_warp * warp;
    int2* stack = (int2*)base;
    double* lcx = (double*)(base + (size_t)stack_cap * sizeof(int2));
    double* lcy = lcx + 32;
    double* lh = lcy + 32;
    double* lw = lh + 32;
    double* cm = lw + 32;
    if (g >= ngroups) return;  // warps are independent: no block-level barrier below
    if (FILL && is_heavy[g]) return;  // handled by k_traverse_heavy
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, nleaves - l0);
    double mycx = 0, mycy = 0;
    if (lane < nl) {
        mycx = L.cx[l0 + lane]; mycy = L.cy[l0 + lane];
        lcx[lane] = mycx; lcy[lane] = mycy; lh[lane] = L.h[l0 + lane]; lw[lane] = L.w[l0 + lane];
    } else { lcx[lane] = 0; lcy[lane] = 0; lh[lane] = 0; lw[lane] = 0; }
    const u32 full = (nl == 32) ? 0xffffffffu : ((1u << nl) - 1u);
    if (lane == 0) stack[0] = make_int2(0, (int)full);
    int size = 1;
    __syncwarp();
    double T1 = 0, T2 = 0, T3 = 0, T4 = 0;
    u32 cnt = 0;
    double nfar = 0;
    const long long gbase = FILL ? G.ptr[g] : 0;
    int iters = 0;
    while (size > 0) {
        if (!FILL && ++iters > kTravBudget) {
            // a fringe group that sees most of the tree: a single warp would serialise the whole
            // step behind it, so it is handed to k_traverse_heavy (one 1024-thread CTA per group)
            if (lane == 0) { heavy[atomicAdd(nheavy, 1)] = g; gcount[g] = 0; }
            return;
        }
        // near the capacity fall back to plain depth-first order (growth <= 1 per pop)
        int take = (size > stack_cap - 80) ? 1 : min(size, 32);
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
            if (FILL && farm) {
                cm[lane * 6 + 0] = T.cmp[3ll * n + 0]; cm[lane * 6 + 1] = T.cmp[3ll * n + 1];
                cm[lane * 6 + 2] = T.cmp[3ll * n + 2]; cm[lane * 6 + 3] = T.cmm[3ll * n + 0];
                cm[lane * 6 + 4] = T.cmm[3ll * n + 1]; cm[lane * 6 + 5] = T.cmm[3ll * n + 2];
            }
        }
        // descend: both children inherit the still-near leaves (child 1 ends up on top)
        bool push = (n >= 0) && nearm && (c1 >= 0);
        u32 pb = __ballot_sync(0xffffffffu, push);
        int npush = 2 * __popc(pb);
        if (size + npush > stack_cap) {
            if (lane == 0) atomicExch(err, 1);
            return;
        }
        if (push) {
            int off = size + 2 * __popc(pb & lanemask_lt());
            stack[off] = make_int2(c1 + 1, (int)nearm);
            stack[off + 1] = make_int2(c1, (int)nearm);
        }
        size += npush;
        // a near leaf: one list entry for the group
        bool emit = (n >= 0) && nearm && (c1 < 0);
        u32 eb = __ballot_sync(0xffffffffu, emit);
        if (FILL && emit) {
            long long k = gbase + cnt + __popc(eb & lanemask_lt());
            G.leaf[k] = T.lstart[n];
            G.mask[k] = nearm;
        }
        cnt += __popc(eb);
        // far nodes: transpose (node lane x leaf bit) -> (leaf lane x node bit) and accumulate
        u32 anyfar = __ballot_sync(0xffffffffu, farm != 0);
        if (anyfar) {
            __syncwarp();
            u32 mine = 0;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                u32 tm = __ballot_sync(0xffffffffu, (farm >> l) & 1u);
                if (lane == l) mine = tm;
            }
            nfar += (double)__popc(mine);
            if (FILL) {
                while (mine) {
                    int j = __ffs(mine) - 1;
                    mine &= mine - 1;
                    const double* c = cm + j * 6;
                    taylor_add(mycx, mycy, c[0], c[1], c[2], T1, T2, T3, T4);
                    taylor_add(mycx, mycy, c[3], c[4], c[5], T1, T2, T3, T4);
                }
            }
        }
        __syncwarp();
    }
    if (!FILL) {
        if (lane == 0) gcount[g] = cnt;
        if (lane < nl && farcount) farcount[l0 + lane] = nfar;
    } else if (lane < nl) {
        double* t = taylor + 4ll * (l0 + lane);
        t[0] = T1 * k1_2Pi; t[1] = T2 * k1_2Pi; t[2] = T3 * k1_Pi; t[3] = T4 * k1_2Pi;  // :66-69
    }
}

// Cooperative traversal of one HEAVY group by a 1024-thread CTA: the same walk as k_traverse with a
// 1024-wide frontier and the stack in global memory (at most one entry per tree node). Entry order
// and the Taylor sums are deterministic (block-wide prefix sums, warp partials combined in order).
template <bool FILL>
__global__ void __launch_bounds__(kHeavyThreads)
k_traverse_heavy(TreeDev T, LeafDev L, int nleaves, const int* heavy, double farc, GroupLists G, u32* gcount,
                 double* taylor, double* farcount, int2* stacks, long long stack_stride, int* err) {
    __shared__ double lcx[32], lcy[32], lh[32], lw[32];
    __shared__ int wn[32][32];
    __shared__ int wpush[33], wemit[33];
    __shared__ double acc[4][32];
    __shared__ double accn[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = heavy[blockIdx.x];
    int2* stack = stacks + stack_stride * blockIdx.x;
    const int l0 = g * kGroupLeaves;
    const int nl = min(kGroupLeaves, nleaves - l0);
    if (tid < 32) {
        bool in = tid < nl;
        lcx[tid] = in ? L.cx[l0 + tid] : 0; lcy[tid] = in ? L.cy[l0 + tid] : 0;
        lh[tid] = in ? L.h[l0 + tid] : 0; lw[tid] = in ? L.w[l0 + tid] : 0;
        acc[0][tid] = acc[1][tid] = acc[2][tid] = acc[3][tid] = 0; accn[tid] = 0;
    }
    const u32 full = (nl == 32) ? 0xffffffffu : ((1u << nl) - 1u);
    if (tid == 0) stack[0] = make_int2(0, (int)full);
    __syncthreads();
    const double mycx = lcx[lane], mycy = lcy[lane];
    int size = 1;
    double T1 = 0, T2 = 0, T3 = 0, T4 = 0, nfar = 0;
    u32 cnt = 0;
    const long long gbase = FILL ? G.ptr[g] : 0;
    while (size > 0) {
        const int take = min(size, kHeavyThreads);
        int n = -1;
        u32 em = 0;
        if (tid < take) { int2 e = stack[size - 1 - tid]; n = e.x; em = (u32)e.y; }
        size -= take;
        u32 farm = 0, nearm = 0;
        int c1 = -1;
        if (n >= 0) {
            double nx = T.x[n], ny = T.y[n];
            double nhw = VV_ADD(T.h[n], T.w[n]);
            c1 = T.ch1[n];
            for (u32 mm = em; mm; mm &= mm - 1) {
                const int l = __ffs(mm) - 1;
                if (is_far(nx, ny, nhw, lcx[l], lcy[l], lh[l], lw[l], farc)) farm |= (1u << l);
            }
            nearm = em & ~farm;
        }
        wn[warp][lane] = n;
        const bool push = (n >= 0) && nearm && (c1 >= 0);
        const bool emit = (n >= 0) && nearm && (c1 < 0);
        const u32 pb = __ballot_sync(0xffffffffu, push), eb = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) { wpush[warp] = __popc(pb); wemit[warp] = __popc(eb); }
        __syncthreads();  // also orders the stack reads above before the pushes below
        if (warp == 0) {
            int a = wpush[lane], b = wemit[lane], ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { ia += ta; ib += tb; }
            }
            wpush[lane] = ia - a; wemit[lane] = ib - b;
            if (lane == 31) { wpush[32] = ia; wemit[32] = ib; }
        }
        __syncthreads();
        if (size + 2 * wpush[32] > stack_stride) {  // cannot happen: every node is pushed at most once
            if (tid == 0) atomicExch(err, 1);
            return;
        }
        if (push) {
            int off = size + 2 * (wpush[warp] + __popc(pb & lanemask_lt()));
            stack[off] = make_int2(c1 + 1, (int)nearm);
            stack[off + 1] = make_int2(c1, (int)nearm);
        }
        if (FILL && emit) {
            long long k = gbase + cnt + wemit[warp] + __popc(eb & lanemask_lt());
            G.leaf[k] = T.lstart[n];
            G.mask[k] = nearm;
        }
        size += 2 * wpush[32];
        cnt += wemit[32];
        // far nodes of this warp's 32 frontier nodes -> lane = leaf
        if (__ballot_sync(0xffffffffu, farm != 0)) {
            u32 mine = 0;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                u32 tm = __ballot_sync(0xffffffffu, (farm >> l) & 1u);
                if (lane == l) mine = tm;
            }
            nfar += (double)__popc(mine);
            if (FILL) {
                while (mine) {
                    int j = __ffs(mine) - 1;
                    mine &= mine - 1;
                    const int nj = wn[warp][j];
                    taylor_add(mycx, mycy, T.cmp[3ll * nj], T.cmp[3ll * nj + 1], T.cmp[3ll * nj + 2], T1, T2, T3, T4);
                    taylor_add(mycx, mycy, T.cmm[3ll * nj], T.cmm[3ll * nj + 1], T.cmm[3ll * nj + 2], T1, T2, T3, T4);
                }
            }
        }
        __syncthreads();  // pushes visible, wn / wpush / wemit free for the next round
    }
    for (int w = 0; w < kHeavyThreads / 32; w++) {  // warp partials, in warp order
        if (warp == w) { acc[0][lane] += T1; acc[1][lane] += T2; acc[2][lane] += T3; acc[3][lane] += T4; accn[lane] += nfar; }
        __syncthreads();
    }
    if (!FILL) {
        if (tid == 0) gcount[g] = cnt;
        if (tid < nl && farcount) farcount[l0 + tid] = accn[tid];
    } else if (tid < nl) {
        double* t = taylor + 4ll * (l0 + tid);
        t[0] = acc[0][tid] * k1_2Pi; t[1] = acc[1][tid] * k1_2Pi; t[2] = acc[2][tid] * k1_Pi; t[3] = acc[3][tid] * k1_2Pi;
    }
}
__global__ void k_mark_heavy(const int* heavy, int nheavy, unsigned char* is_heavy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nheavy) is_heavy[heavy[i]] = 1;
}

__global__ void k_group_ptr(const u32* gscan, long long* ptr, int ngroups) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= ngroups) ptr[i] = gscan[i];
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

// near-pair count of one group: sum over entries of (#targets g!=0 in masked leaves) x (#sources g!=0)
__global__ void k_count_pairs(LeafDev L, int nleaves, int ngroups, GroupLists G, const double* __restrict__ pg,
                              double* out) {
    int g = blockIdx.x;
    if (g >= ngroups) return;
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
    for (long long e = G.ptr[g] + threadIdx.x; e < G.ptr[g + 1]; e += blockDim.x) {
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
        out[g] = t;
    }
}

}  // namespace vv
