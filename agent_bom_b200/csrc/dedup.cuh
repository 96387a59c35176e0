// dedup.cuh — root-frontier de-duplication for batches of single-source walks (sm_100a).
//
// A finding's blast radius is a function of what it is attached to, not of the
// finding itself: every CVE on the same package version has the same reverse
// reach.  Formally, for a single-source walk from f (source pre-visited, depth
// limit D), let S'(f) be the ordered, de-duplicated list of f's passing
// neighbours other than f — exactly the depth-1 part of the reference's queue.
// Everything the sequential BFS does afterwards depends only on S'(f) and on
// the fact that f is visited.  So two sources with the same S' have the same
// result (same ORDER), provided neither source is itself reached again.
//
// The batch is therefore grouped by S' (64-bit signature sort + exact
// verification of the lists), ONE "canonical" walk per group is seeded with S'
// (depth limit D-1, depths biased by one), every member source is tested
// against the canonical walk's visited set, and members simply point their
// result slice at the group's slice.  A member whose source WAS reached (a
// cycle through the source), a hash collision, or a source with more than
// L1_CAP neighbours falls back to an individual walk — same kernels, same
// order, so the output is bit-identical either way (tests run both paths).
#pragma once
#include "walk.cuh"

namespace abb {

constexpr int L1_CAP = 8;
// Signatures live in the low SIG_BITS bits so the radix sort only passes over those (a 48-bit hash still makes a false match
// between two of ~10^7 frontiers a 10^-1-per-batch event at worst, and every match is verified element-wise anyway);
// bit SIG_BITS-1 marks an ineligible source, whose key is its own query index (sorted behind every eligible one).
constexpr int SIG_BITS = 48;
constexpr unsigned long long SIG_INELIGIBLE = 1ull << (SIG_BITS - 1);

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// S'(f): returns the length (0..L1_CAP) or -1 when f is not eligible (invalid root or too many candidates)
__device__ __forceinline__ int level1_list(const GraphView &g, const abb_walk_spec &sp, int32_t f, int32_t *out) {
    if (f < 0 || f >= g.n) return -1;
    if ((sp.flags & ABB_WALK_REAL_ROOTS) && __ldg(g.ntype + f) == ABB_NODE_GHOST) return -1;
    int n = 0;
    for (int pass = 0; pass < 2; pass++) {
        if (!(sp.direction & (1 << pass))) continue;
        const uint32_t *off = pass ? g.roff : g.foff; const int32_t *nb = pass ? g.rnbr : g.fnbr; const uint8_t *me = pass ? g.rmeta : g.fmeta;
        const uint32_t a = __ldg(off + f), b = __ldg(off + f + 1);
        if (b - a > 4u * L1_CAP) return -1;
        for (uint32_t p = a; p < b; p++) {
            const uint32_t m = __ldg(me + p);
            if (!((sp.rel_mask >> (m & ABB_META_REL_MASK)) & 1u)) continue;
            if ((sp.flags & ABB_WALK_TRAVERSABLE_ONLY) && !(m & ABB_META_TRAVERSABLE)) continue;
            const int32_t v = __ldg(nb + p);
            if (v == f) continue;                       // the source is already visited
            bool dup = false;
            for (int i = 0; i < n; i++) dup |= (out[i] == v);
            if (dup) continue;
            if (n == L1_CAP) return -1;
            out[n++] = v;
        }
    }
    return n;
}

__global__ void dedup_sig_kernel(GraphView g, abb_walk_spec sp, const int32_t *roots, int64_t nq, unsigned long long *sig, int32_t *qidx) {
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int32_t lst[L1_CAP];
    const int n = level1_list(g, sp, __ldg(roots + q), lst);
    unsigned long long h;
    if (n < 0) {
        h = SIG_INELIGIBLE | static_cast<unsigned long long>(q);
    } else {
        h = mix64(0xABB200ull + n);
        for (int i = 0; i < n; i++) h = mix64(h ^ (static_cast<unsigned long long>(static_cast<uint32_t>(lst[i])) + (static_cast<unsigned long long>(i + 1) << 32)));
        h &= SIG_INELIGIBLE - 1ull;
    }
    sig[q] = h;
    if (qidx) qidx[q] = static_cast<int32_t>(q);
}

// per sorted position: group-head flag; ineligible sources go straight to the individual list.
// counters: [0] n_groups (written by the select), [1] n_individual, [2] n_eligible
__global__ void dedup_heads_kernel(const unsigned long long *ssig, const int32_t *sq, int64_t nq, uint8_t *head, int32_t *indiv, unsigned long long *counters) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const unsigned long long k = ssig[i];
    if (k & SIG_INELIGIBLE) {
        head[i] = 0;
        indiv[atomicAdd(counters + 1, 1ull)] = sq[i];
        if (i == 0 || !(ssig[i - 1] & SIG_INELIGIBLE)) counters[2] = static_cast<unsigned long long>(i);
    } else {
        head[i] = (i == 0 || ssig[i - 1] != k) ? 1 : 0;
        if (i == nq - 1) counters[2] = static_cast<unsigned long long>(nq);
    }
}

// per group: frontier length of its leader; member range start
__global__ void dedup_groups_kernel(GraphView g, abb_walk_spec sp, const int32_t *roots, const int32_t *sq, const int64_t *hp, const unsigned long long *counters,
                                    int64_t *glen, int64_t *mem_off) {
    const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t ng = static_cast<int64_t>(counters[0]);
    if (gi == ng) mem_off[ng] = static_cast<int64_t>(counters[2]);
    if (gi >= ng) return;
    int32_t lst[L1_CAP];
    const int64_t pos = hp[gi];
    const int n = level1_list(g, sp, __ldg(roots + sq[pos]), lst);
    glen[gi] = n < 0 ? 0 : n;
    mem_off[gi] = pos;
}

// per group: write the leader's frontier into the canonical root arena
__global__ void dedup_roots_kernel(GraphView g, abb_walk_spec sp, const int32_t *roots, const int32_t *sq, const int64_t *hp, const unsigned long long *counters,
                                   const int64_t *goff, int32_t *arena) {
    const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gi >= static_cast<int64_t>(counters[0])) return;
    int32_t lst[L1_CAP];
    const int n = level1_list(g, sp, __ldg(roots + sq[hp[gi]]), lst);
    const int64_t o = goff[gi];
    for (int i = 0; i < n; i++) arena[o + i] = lst[i];
}

// per group: locality key = its first seed node (builder insertion order clusters an agent's servers / packages / credentials, so groups
// with neighbouring seeds walk the same credential cliques); groups are then WALKED in key order so that those rows stay in L2
__global__ void dedup_locality_keys_kernel(const unsigned long long *counters, const int64_t *goff, const int32_t *arena, uint32_t *key, int32_t *gid, int64_t cap) {
    const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gi >= cap) return;
    const int64_t ng = static_cast<int64_t>(counters[0]);
    gid[gi] = static_cast<int32_t>(gi);
    key[gi] = (gi < ng && goff[gi + 1] > goff[gi]) ? static_cast<uint32_t>(arena[goff[gi]]) : 0xFFFFFFFFu;   // padding sorts last and is never read (count = n_groups)
}

// per eligible member: its group id (for the share pass), its source, and exact verification of its list against
// the group's (signature collisions are sent to the individual list)
__global__ void dedup_members_kernel(GraphView g, abb_walk_spec sp, const int32_t *roots, const int32_t *sq, const uint8_t *head, const int64_t *gid_incl,
                                     const int64_t *goff, const int32_t *arena, int32_t *mem_src, int32_t *mem_state, int32_t *indiv, unsigned long long *counters) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(counters[2])) return;
    const int32_t f = __ldg(roots + sq[i]);
    mem_src[i] = f;
    int state = 0;
    if (!head[i]) {
        const int64_t gi = gid_incl[i] - 1;
        const int64_t o = goff[gi], len = goff[gi + 1] - o;
        int32_t lst[L1_CAP];
        const int n = level1_list(g, sp, f, lst);
        bool same = (n == len);
        for (int k = 0; same && k < n; k++) same = (lst[k] == arena[o + k]);
        if (!same) { state = 1; indiv[atomicAdd(counters + 1, 1ull)] = sq[i]; }
    }
    mem_state[i] = state;
}

// per eligible member: adopt the group's result slice (or queue an individual walk when the source itself was reached)
__global__ void dedup_share_kernel(const int32_t *sq, const int64_t *gid_incl, const int32_t *mem_state, const unsigned long long *counters_ro,
                                   const int64_t *g_start, const int32_t *g_count, const int32_t *g_maxd, const int32_t *g_flags, const uint32_t *g_hist,
                                   int64_t *q_start, int32_t *q_count, int32_t *q_maxd, int32_t *q_flags, uint32_t *q_hist, int32_t *indiv,
                                   unsigned long long *counters) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t ne = static_cast<int64_t>(counters_ro[2]);
    // 32 members per warp iteration; the histogram rows are copied warp-cooperatively
    for (int64_t base = warp * 32; base < ne; base += nwarps * 32) {
        const int64_t i = base + lane;
        int state = 1; int32_t q = 0; int64_t gi = 0;
        if (i < ne) { state = mem_state[i]; q = sq[i]; gi = gid_incl[i] - 1; }
        if (i < ne && state == 2) indiv[atomicAdd(counters + 1, 1ull)] = q;
        if (i < ne && state == 0) {
            q_start[q] = g_start[gi]; q_count[q] = g_count[gi]; q_maxd[q] = g_maxd[gi];
            q_flags[q] = g_flags[gi] & ~ABB_QFLAG_NO_ROOT;     // an empty frontier is a valid source that reaches nothing
        }
        if (q_hist) {
            for (int j = 0; j < 32; j++) {
                const int st = __shfl_sync(FULL, state, j);
                const int64_t ii = base + j;
                if (ii >= ne || st != 0) continue;
                const int32_t qq = __shfl_sync(FULL, q, j);
                const int64_t gg = __shfl_sync(FULL, gi, j);
                if (lane < ABB_N_ENTITY_TYPES) q_hist[static_cast<int64_t>(qq) * ABB_N_ENTITY_TYPES + lane] = g_hist[gg * ABB_N_ENTITY_TYPES + lane];
            }
        }
    }
}

}  // namespace abb
