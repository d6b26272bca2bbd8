// Target sharding over ranks: who owns which leaf groups, and the pack / unpack kernels of the
// exchange that follows every sharded phase.
//
// Ownership is BLOCK-CYCLIC over the leaf groups: piece k = groups [k * kShardBlock, (k + 1) * kShardBlock) belongs to
// rank k mod P. Neighbouring groups cost about the same (same local density, same list length) and the few fringe
// groups that see the whole tree are dealt round the ranks, so every phase balances without measuring anything
// (contiguous slices cut by group count left the per-rank K2 times 2x apart: the fringe groups sat in the first and
// last slice). A piece is a contiguous particle range; what a rank owns is therefore a list of ranges, packed densely
// for the all-gather and scattered back by the same table on the other side.
#pragma once
#include "vvgpu_tree_build.cuh"
#include "vvgpu_lists.cuh"

namespace vv {

constexpr int kMaxRanks = 64;

// groups owned by `rank` out of ng
inline int shard_count(int ng, int rank, int nranks) {
    const int npieces = (ng + kShardBlock - 1) / kShardBlock;
    int cnt = 0;
    for (int k = rank; k < npieces; k += nranks) cnt += std::min(kShardBlock, ng - k * kShardBlock);
    return cnt;
}

struct ShardTable {
    int* first;   // per piece: first particle
    int* cnt;     // per piece: particles
    int* off;     // per piece: offset inside its owner's packed block
};

// one CTA; writes the per-rank particle counts to rank_cnt[0 .. nranks)
__global__ void __launch_bounds__(1024) k_shard_table(TreeDev T, const BuildState* st, int nranks, ShardTable S, int* rank_cnt) {
    if (st->err) return;
    const int nl = st->nleaves;
    const int ng = (nl + kGroupLeaves - 1) / kGroupLeaves;
    const int npieces = (ng + kShardBlock - 1) / kShardBlock;
    for (int k = threadIdx.x; k < npieces; k += blockDim.x) {
        const int l0 = k * kShardBlock * kGroupLeaves, l1 = min(nl, (k + 1) * kShardBlock * kGroupLeaves);
        const int pf = T.first[T.leaf_node[l0]], pe = T.last[T.leaf_node[l1 - 1]];
        S.first[k] = pf; S.cnt[k] = pe - pf;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < nranks; r += (int)(blockDim.x >> 5)) {   // one warp per rank: running offsets of its pieces
        int run = 0;
        for (int j0 = 0; r + (long long)j0 * nranks < npieces; j0 += 32) {
            const long long k = r + (long long)(j0 + lane) * nranks;
            const int c = (k < npieces) ? S.cnt[k] : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (k < npieces) S.off[k] = run + inc - c;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) rank_cnt[r] = run;
    }
}

// the arrays of one exchange: 8-byte or 4-byte elements. A rank's block = the arrays one after the other, each L
// elements long (L even, so that every array starts 8-byte aligned), then one 8-byte slot for a scalar.
constexpr int kMaxXArr = 8;
struct XArrays {
    void* p[kMaxXArr];
    int wide[kMaxXArr];       // 1: 8-byte elements, 0: 4-byte
    long long off[kMaxXArr];  // byte offset of the array inside a rank's block (filled by xarrays_layout)
    int n;
};
// byte size of a rank's block (without the scalar slot)
inline long long xarrays_layout(XArrays& X, long long L) {
    long long o = 0;
    for (int a = 0; a < X.n; a++) { X.off[a] = o; o += L * (X.wide[a] ? 8 : 4); }
    return o;
}

// send block <- the pieces this rank owns, packed densely in piece order; one CTA per owned piece
__global__ void __launch_bounds__(256) k_shard_pack(ShardTable S, int npieces, Shard sh, XArrays X, unsigned char* send) {
    const long long k = sh.rank + (long long)blockIdx.x * sh.nranks;
    if (k >= npieces) return;
    const int f = S.first[k], c = S.cnt[k], o = S.off[k];
    for (int a = 0; a < X.n; a++) {
        if (X.wide[a]) {
            u64* dst = (u64*)(send + X.off[a]) + o;
            const u64* src = (const u64*)X.p[a] + f;
            for (int i = threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
        } else {
            u32* dst = (u32*)(send + X.off[a]) + o;
            const u32* src = (const u32*)X.p[a] + f;
            for (int i = threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
        }
    }
}
// the other ranks' pieces back into the arrays; recv = [rank][stride bytes]; one CTA per piece
__global__ void __launch_bounds__(256) k_shard_unpack(ShardTable S, int npieces, Shard sh, XArrays X, long long stride,
                                                      const unsigned char* recv) {
    const int k = blockIdx.x;
    if (k >= npieces) return;
    const int owner = k % sh.nranks;
    if (owner == sh.rank) return;
    const int f = S.first[k], c = S.cnt[k], o = S.off[k];
    const unsigned char* blk = recv + owner * stride;
    for (int a = 0; a < X.n; a++) {
        if (X.wide[a]) {
            const u64* src = (const u64*)(blk + X.off[a]) + o;
            u64* dst = (u64*)X.p[a] + f;
            for (int i = threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
        } else {
            const u32* src = (const u32*)(blk + X.off[a]) + o;
            u32* dst = (u32*)X.p[a] + f;
            for (int i = threadIdx.x; i < c; i += blockDim.x) dst[i] = src[i];
        }
    }
}
// out[i] = sum over ranks (in rank order) of recv[r * stride + i]
__global__ void k_rank_sum_f64(const double* recv, long long stride, int nranks, int n, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0;
    for (int r = 0; r < nranks; r++) s += recv[r * stride + i];
    out[i] = s;
}
__global__ void k_rank_sum_i32(const int* recv, long long stride_ints, int nranks, int n, int* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = 0;
    for (int r = 0; r < nranks; r++) s += recv[r * stride_ints + i];
    out[i] = s;
}

// absorbed-by table of a merge solution from its (init, part) columns: absby[q] = first initiator that merges with q
__global__ void k_merge_absby(int n, const int* __restrict__ init, const int* __restrict__ part, int* absby) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (init[i]) atomicMin(&absby[part[i]], i);
}
// First guess of the merge fixed point, after the round that assumed no merges: an initiator that an EARLIER initiator
// absorbs never gets its turn (MEpsilonFast.cpp:51, g == 0 by then). In a wake full of mutual nearest neighbours that
// is half of round 0's initiators; cancelling them here (a few sweeps of: absorbed-by from the valid initiators, valid =
// not absorbed before its turn) spares the fixed point one full round. Any guess converges to the same solution.
__global__ void k_prune_valid(int n, const int* __restrict__ init0, const int* __restrict__ absby, unsigned char* valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) valid[i] = (init0[i] && !(absby[i] < i)) ? 1 : 0;
}
__global__ void k_prune_absby(int n, const unsigned char* __restrict__ valid, const int* __restrict__ part, int* absby) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && valid[i]) atomicMin(&absby[part[i]], i);
}
__global__ void k_prune_commit(int n, const unsigned char* __restrict__ valid, int* init, int* part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && init[i] && !valid[i]) { init[i] = 0; part[i] = -1; }
}
__global__ void k_fill_i32(int n, int* a, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

}  // namespace vv
