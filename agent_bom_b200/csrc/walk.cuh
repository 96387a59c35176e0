// walk.cuh — batched, ordered, multi-source frontier BFS over the typed CSR (sm_100a).
//
// One warp owns one query at a time (persistent grid, atomic work counter).
// The warp keeps the query's FIFO queue, visited set and per-entry depth in
// shared memory (tier S) or, for the rare query that outgrows it, in a
// per-warp global scratch slot with a bitmap visited set (tier G).
//
// Exactness: the reference engine is a sequential deque BFS
// (graph/container.py:230-279,367-391,411-436,438-538; dependency_reach.py:169-198).
// The warp reproduces its discovery ORDER, not only its sets:
//   * a level's frontier is consumed in queue order, 32 frontier nodes at a time;
//   * their adjacency rows are flattened (warp prefix sum over degrees) and the
//     flattened candidate list is consumed 32 candidates at a time, i.e. in
//     (queue position, row position) lexicographic order = the reference's scan order;
//   * inside a 32-candidate chunk, __match_any_sync elects the lowest lane among
//     duplicates of a neighbour and ballot/popc ranks give the append positions,
//     so the queue grows exactly as the sequential loop would grow it;
//   * max_edges / max_nodes budgets are applied with warp prefix counts at the
//     exact candidate where the sequential loop would hit them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abb200.h"

namespace abb {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int32_t EMPTY = -1;
constexpr uint32_t NO_TOK = 0xFFFFu;

struct GraphView {
    int32_t n;
    int64_t m;
    const uint32_t *foff; const int32_t *fnbr; const uint8_t *fmeta; const uint32_t *feid;
    const uint32_t *roff; const int32_t *rnbr; const uint8_t *rmeta; const uint32_t *reid;
    const uint8_t *ntype;
    const int32_t *rank;
};

struct WalkArgs {
    GraphView g;
    abb_walk_spec spec;
    abb_walk_io io;
    const int32_t *qlist;        // queries of this launch (NULL = 0..nq-1)
    int64_t nq;
    const unsigned long long *nq_dev;  // when set, the query count is read from device memory (overflow tiers): entries qlist[0 .. *nq_dev)
    const unsigned long long *nq_back_dev;   // second part of a two-ended work list: entries qlist[list_cap-1 .. list_cap-*nq_back_dev], taken after the front part
    const int32_t *qlist2;       // a second work list, taken after the first (the big block tier's hand-offs, joined to S1's for the next tier)
    const unsigned long long *nq2_dev;
    int64_t list_cap;            // capacity of qlist / overflow (a two-ended list grows from both ends)
    unsigned long long *ctl;     // [0] work counter, [1] overflow count (front), [2] fatal flag, [3] overflow count (back)
    int32_t *overflow;           // query ids that outgrew this tier: HEAVY ones (forecast says thousands of nodes) from the front, the others
                                 // from the back when ov_cnt_back is set — the next tier then starts the heavy walks first (or a block tier takes them)
    unsigned long long *ov_cnt_front, *ov_cnt_back;
    // tier G scratch (per warp slot)
    uint32_t *g_bitmap; int32_t *g_queue; int32_t *g_par; int32_t *g_dep;
    int64_t g_words, g_qcap;
    // canonical (root-frontier de-duplicated) mode: a query is a GROUP of sources that share their level-1 frontier,
    // seeded with that frontier; depths are offset by one and every member source is tested against the visited set
    int32_t slice_align;         // result slices start on a multiple of this many ints (32 when the arena is host-mapped: whole 128-byte lines per PCIe write)
    int32_t depth_bias;          // added to emitted depths / max depth (1 in canonical mode)
    int32_t hist_roots;          // count the seeded roots in the histogram too
    const int64_t *mem_off;      // [groups+1] member range of each group in the sorted member arrays (NULL = not canonical)
    const int32_t *mem_src;      // source node of each member
    int32_t *mem_state;          // set to 2 when the member's own source was reached (its result differs: traversed individually)
};

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// ---------------------------------------------------------------- tier work lists
__device__ __forceinline__ int64_t list_count(const WalkArgs &A) {
    if (!A.nq_dev && !A.nq_back_dev) return A.nq;
    int64_t n = 0;
    if (A.nq_dev) n += static_cast<int64_t>(*A.nq_dev);
    if (A.nq_back_dev) n += static_cast<int64_t>(*A.nq_back_dev);
    if (A.nq2_dev) n += static_cast<int64_t>(*A.nq2_dev);
    return n;
}
__device__ __forceinline__ int64_t list_item(const WalkArgs &A, int64_t i) {
    if (!A.qlist) return i;
    if (!A.nq_back_dev && !A.nq2_dev) return A.qlist[i];
    const int64_t front = A.nq_dev ? static_cast<int64_t>(*A.nq_dev) : 0;
    if (i < front) return A.qlist[i];
    const int64_t back = A.nq_back_dev ? static_cast<int64_t>(*A.nq_back_dev) : 0;
    if (i < front + back) return A.qlist[A.list_cap - 1 - (i - front)];
    return A.qlist2[i - front - back];
}
__device__ __forceinline__ void list_append(const WalkArgs &A, int64_t q, bool heavy) {
    if (A.ov_cnt_back && !heavy) {
        const unsigned long long k = atomicAdd(A.ov_cnt_back, 1ull);
        A.overflow[A.list_cap - 1 - static_cast<int64_t>(k)] = static_cast<int32_t>(q);
    } else {
        const unsigned long long k = atomicAdd(A.ov_cnt_front, 1ull);
        A.overflow[k] = static_cast<int32_t>(q);
    }
}

// ---------------------------------------------------------------- tier S store
// Visited set = open addressing in BUCKETS of four 32-bit slots (16 bytes): a probe is ONE ld.shared.v4 plus four compares,
// and with at most Q entries in H = 4·Q slots a second bucket is almost never needed — the revisit-dominated inner loop of
// the walk (>90 % of all scanned entries on clique-shaped estates) has no data-dependent loop in the common case.  Slots of
// a bucket fill in order and nothing is ever deleted during a walk, so an EMPTY last slot ends an unsuccessful probe.
// The table is wiped with vector stores after each walk (H/128 stores per lane) instead of tracking touched slots.
template <int H, int Q, bool PAR>
struct SmemStore {
    static constexpr bool kGlobal = false;
    static constexpr int kBytes = H * 4 + Q * 4 + Q + (PAR ? Q * 4 : 0);
    static constexpr int NB = H / 4;                      // buckets
    int32_t *htab; int32_t *queue; int32_t *par; uint8_t *dep;
    uint32_t hbase;     // shared-window address of htab: probes use ld.shared directly (no generic-address arithmetic in the hot loop)
    __device__ SmemStore(unsigned char *base) {
        htab = reinterpret_cast<int32_t *>(base);
        queue = htab + H;
        par = PAR ? queue + Q : nullptr;
        dep = reinterpret_cast<uint8_t *>(queue + Q + (PAR ? Q : 0));
        hbase = static_cast<uint32_t>(__cvta_generic_to_shared(base));
        asm volatile("mov.u32 %0, %0;" : "+r"(hbase));    // opaque: keep it in a register instead of re-deriving it at every probe
    }
    __device__ void init(int lane) {
        for (int i = lane; i < NB; i += 32) reinterpret_cast<int4 *>(htab)[i] = make_int4(EMPTY, EMPTY, EMPTY, EMPTY);
        __syncwarp();
    }
    __device__ __forceinline__ int qcap() const { return Q; }
    __device__ __forceinline__ int max_level() const { return 255; }
    static_assert((NB & (NB - 1)) == 0 && Q * 4 <= H, "bucket count must be a power of two and the table at most a quarter full");
    static constexpr int log2c(int v) { return v <= 1 ? 0 : 1 + log2c(v >> 1); }
    __device__ __forceinline__ static uint32_t bucket(int32_t k) { return (static_cast<uint32_t>(k) * 0x9E3779B1u) >> (32 - log2c(NB)); }
    // k == EMPTY (-1, the value lanes past the end of a row carry) reads as "present": the first bucket with a free slot matches it
    __device__ __forceinline__ bool contains(int32_t k) const {
        uint32_t b = bucket(k);
        for (;;) {
            int x, y, z, w;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(hbase + b * 16u) : "memory");
            if ((x == k) | (y == k) | (z == k) | (w == k)) return true;
            if (w == EMPTY) return false;                  // bucket not full: k would be in it
            b = (b + 1) & (NB - 1);
        }
    }
    // true when newly inserted (t is unused: the table is wiped as a whole)
    __device__ __forceinline__ bool test_and_set(int32_t k, uint32_t &t) {
        t = 0;
        uint32_t b = bucket(k);
        for (;;) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const int32_t old = atomicCAS(&htab[b * 4 + s], EMPTY, k);
                if (old == EMPTY) return true;
                if (old == k) return false;
            }
            b = (b + 1) & (NB - 1);
        }
    }
    __device__ __forceinline__ int32_t q_get(int i) const { return queue[i]; }
    __device__ __forceinline__ void put(int i, int32_t node, int32_t p, uint32_t, int d) {
        queue[i] = node; dep[i] = static_cast<uint8_t>(d);
        if (PAR) par[i] = p;
    }
    __device__ __forceinline__ int32_t par_get(int i) const { return PAR ? par[i] : -1; }
    __device__ __forceinline__ int dep_get(int i) const { return dep[i]; }
    __device__ __forceinline__ void unset(int32_t, uint32_t) {}      // only called right before clear()
    __device__ void clear(int, int lane) {
        __syncwarp();
        for (int i = lane; i < NB; i += 32) reinterpret_cast<int4 *>(htab)[i] = make_int4(EMPTY, EMPTY, EMPTY, EMPTY);
        __syncwarp();
    }
};

// ---------------------------------------------------------------- tier G store
struct GlobalStore {
    static constexpr bool kGlobal = true;
    uint32_t *bits; int32_t *queue; int32_t *par; int32_t *dep; int cap;   // dep == nullptr when no depth is ever read back
    __device__ void init(int) {}
    __device__ __forceinline__ int qcap() const { return cap; }
    __device__ __forceinline__ int max_level() const { return 0x7FFFFFF0; }
    // read through L2 (ld.cg): the bits are set by L2 atomics, an L1 line could be stale
    __device__ __forceinline__ bool contains(int32_t k) const { return (__ldcg(bits + (k >> 5)) >> (k & 31)) & 1u; }
    __device__ __forceinline__ void probe_first(int32_t k, uint32_t &h, bool &found, bool &ended) const { h = 0; found = contains(k); ended = true; }
    __device__ __forceinline__ bool probe_more(int32_t, uint32_t) const { return false; }
    __device__ __forceinline__ bool test_and_set(int32_t k, uint32_t &t) {
        t = 0;
        uint32_t bit = 1u << (k & 31);
        uint32_t old = atomicOr(&bits[k >> 5], bit);
        return !(old & bit);
    }
    __device__ __forceinline__ int32_t q_get(int i) const { return queue[i]; }
    __device__ __forceinline__ void put(int i, int32_t node, int32_t p, uint32_t, int d) {
        queue[i] = node; if (dep) dep[i] = d; if (par) par[i] = p;
    }
    __device__ __forceinline__ int32_t par_get(int i) const { return par ? par[i] : -1; }
    __device__ __forceinline__ int dep_get(int i) const { return dep ? dep[i] : 0; }
    __device__ __forceinline__ void unset(int32_t k, uint32_t) { atomicAnd(&bits[k >> 5], ~(1u << (k & 31))); }
    __device__ void clear(int count, int lane) {
        __syncwarp();
        for (int i = lane; i < count; i += 32) bits[queue[i] >> 5] = 0u;
        __threadfence_block();
        __syncwarp();
    }
};

// ---------------------------------------------------------------- candidate enumeration
// Flattened candidates of up to 32 frontier nodes.  Lane j describes frontier
// node j: forward row [sF, sF+dF) then reverse row [sR, sR+dR).
struct Frontier32 {
    uint32_t sF, dF, sR, excl;  // excl = exclusive prefix of (dF+dR)
    uint32_t total;
};

__device__ __forceinline__ Frontier32 load_frontier(const GraphView &g, int dir, int32_t u, bool valid) {
    Frontier32 f; f.sF = f.dF = f.sR = 0; uint32_t dR = 0;
    if (valid) {
        if (dir & 1) { uint32_t a = __ldg(g.foff + u), b = __ldg(g.foff + u + 1); f.sF = a; f.dF = b - a; }
        if (dir & 2) { uint32_t a = __ldg(g.roff + u), b = __ldg(g.roff + u + 1); f.sR = a; dR = b - a; }
    }
    uint32_t deg = f.dF + dR, inc = deg;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { uint32_t v = __shfl_up_sync(FULL, inc, s); if (lane >= s) inc += v; }
    f.excl = inc - deg;
    f.total = __shfl_sync(FULL, inc, 31);
    return f;
}

struct Cand { int32_t nbr; uint32_t meta; uint32_t eid; int owner; bool active; };

// candidate ci (clamped) -> owning frontier lane + CSR position
template <bool NEED_META, bool NEED_EID>
__device__ __forceinline__ Cand fetch_cand(const GraphView &g, const Frontier32 &f, uint32_t ci) {
    Cand c; c.active = ci < f.total;
    uint32_t cc = c.active ? ci : (f.total - 1);
    int j = 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        int t = j + s;
        uint32_t v = __shfl_sync(FULL, f.excl, t & 31);
        if (t < 32 && v <= cc) j = t;
    }
    // several lanes may share excl (zero-degree nodes): the LAST lane with excl<=cc owns it,
    // but lanes past the frontier carry excl==total>cc, so j is a real frontier lane.
    uint32_t ex = __shfl_sync(FULL, f.excl, j), sF = __shfl_sync(FULL, f.sF, j), dF = __shfl_sync(FULL, f.dF, j), sR = __shfl_sync(FULL, f.sR, j);
    uint32_t k = cc - ex;
    c.owner = j; c.meta = ABB_META_TRAVERSABLE; c.eid = 0; c.nbr = 0;
    if (c.active) {
        if (k < dF) {
            uint32_t p = sF + k; c.nbr = __ldg(g.fnbr + p);
            if (NEED_META) c.meta = __ldg(g.fmeta + p);
            if (NEED_EID) c.eid = __ldg(g.feid + p);
        } else {
            uint32_t p = sR + (k - dF); c.nbr = __ldg(g.rnbr + p);
            if (NEED_META) c.meta = __ldg(g.rmeta + p);
            if (NEED_EID) c.eid = __ldg(g.reid + p);
        }
    }
    return c;
}

// fast path for a frontier chunk that holds ONE node (the common case on tree-like inventories): no prefix scan, no search
template <bool NEED_META, bool NEED_EID>
__device__ __forceinline__ Cand fetch_single(const GraphView &g, uint32_t sF, uint32_t dF, uint32_t sR, uint32_t total, uint32_t ci) {
    Cand c; c.active = ci < total; c.owner = 0; c.meta = ABB_META_TRAVERSABLE; c.eid = 0; c.nbr = 0;
    if (c.active) {
        if (ci < dF) {
            const uint32_t p = sF + ci; c.nbr = __ldg(g.fnbr + p);
            if (NEED_META) c.meta = __ldg(g.fmeta + p);
            if (NEED_EID) c.eid = __ldg(g.feid + p);
        } else {
            const uint32_t p = sR + (ci - dF); c.nbr = __ldg(g.rnbr + p);
            if (NEED_META) c.meta = __ldg(g.rmeta + p);
            if (NEED_EID) c.eid = __ldg(g.reid + p);
        }
    }
    return c;
}

__device__ __forceinline__ bool cand_passes(const abb_walk_spec &sp, const Cand &c) {
    return c.active && ((sp.rel_mask >> (c.meta & ABB_META_REL_MASK)) & 1u) &&
           (!(sp.flags & ABB_WALK_TRAVERSABLE_ONLY) || (c.meta & ABB_META_TRAVERSABLE));
}

__device__ __forceinline__ bool type_emitted(uint32_t emit_types, uint8_t t) {
    return (emit_types >> (t < 31 ? t : 31)) & 1u;
}


// ---------------------------------------------------------------- row mode
// Lean expansion of ONE frontier node's row(s): candidates are consumed 32 at a time in row order (forward row, then
// reverse row), which is exactly the reference's scan order for that node, with no flattening scan / search.  The
// visited set is PROBED first (plain loads); the ordered insert/append machinery runs only for chunks that hold an
// unvisited candidate — on clique-shaped inventories >90 % of all scanned entries are revisits (80 agents sharing a
// credential each list the other 79), so the common chunk costs one coalesced load, one hash and one probe.
struct RowCand { int32_t nbr; uint32_t meta; uint32_t pos; bool fwd; bool active; };

template <bool NEED_META>
__device__ __forceinline__ RowCand row_load(const GraphView &g, uint32_t sF, uint32_t dF, uint32_t sR, uint32_t tot, uint32_t k) {
    RowCand c; c.active = k < tot; c.nbr = 0; c.meta = ABB_META_TRAVERSABLE; c.pos = 0; c.fwd = true;
    if (c.active) {
        if (k < dF) { c.pos = sF + k; c.nbr = __ldg(g.fnbr + c.pos); if (NEED_META) c.meta = __ldg(g.fmeta + c.pos); }
        else { c.fwd = false; c.pos = sR + (k - dF); c.nbr = __ldg(g.rnbr + c.pos); if (NEED_META) c.meta = __ldg(g.rmeta + c.pos); }
    }
    return c;
}

// Returns false when the store overflowed (the caller re-queues the query; the store has been cleared).
template <class Store, bool NEED_META, class idx_t>
__device__ __forceinline__ bool expand_row(const WalkArgs &A, Store &st, uint32_t sF, uint32_t dF, uint32_t sR, uint32_t tot, int32_t parent_pos, int depth1,
                                           int lane, idx_t &tail, long long &rec_edges) {
    const GraphView &g = A.g;
    const abb_walk_spec &sp = A.spec;
    const uint32_t fl = sp.flags;
    // two chunks in flight ahead of the one being probed
    RowCand n1 = row_load<NEED_META>(g, sF, dF, sR, tot, lane);
    RowCand n2 = row_load<NEED_META>(g, sF, dF, sR, tot, 32 + lane);
    for (uint32_t c0 = 0; c0 < tot; c0 += 32) {
        const RowCand c = n1;
        n1 = n2;
        n2 = row_load<NEED_META>(g, sF, dF, sR, tot, c0 + 64 + lane);
        bool pass = c.active;
        if (NEED_META) pass = pass && ((sp.rel_mask >> (c.meta & ABB_META_REL_MASK)) & 1u) && (!(fl & ABB_WALK_TRAVERSABLE_ONLY) || (c.meta & ABB_META_TRAVERSABLE));
        if (fl & ABB_WALK_EDGES) rec_edges += __popc(__ballot_sync(FULL, pass));
        const bool unv = pass && !st.contains(c.nbr);
        if (__ballot_sync(FULL, unv) == 0u) continue;
        // rare path: some candidate of this chunk is new.  The first occurrence of a neighbour inside the chunk speaks for it:
        // FIRST_PAIR when the chunk lies in one unfiltered row (an earlier occurrence in an earlier chunk would already be visited).
        bool leader;
        if (!NEED_META && sp.direction != ABB_DIR_BOTH) {
            const uint32_t m = unv ? static_cast<uint32_t>(__ldg((c.fwd ? g.fmeta : g.rmeta) + c.pos)) : 0u;
            leader = unv && (m & ABB_META_FIRST_PAIR);
        } else {
            const unsigned mm = __match_any_sync(FULL, unv ? c.nbr : (-2 - lane));
            leader = unv && (__ffs(mm) - 1) == lane;
        }
        uint32_t tok = NO_TOK;
        const bool isnew = leader && st.test_and_set(c.nbr, tok);
        const unsigned nm = __ballot_sync(FULL, isnew);
        const int cnt = __popc(nm);
        if (cnt) {
            if (tail + cnt > st.qcap()) {
                if (isnew) st.unset(c.nbr, tok);
                st.clear(tail, lane);
                return false;
            }
            if (isnew) st.put(tail + __popc(nm & lanemask_lt(lane)), c.nbr, parent_pos, tok, depth1);
            tail += cnt;
        }
        __syncwarp();
    }
    return true;
}


// Lean form of the row mode for the plain single-direction walk (no relationship / traversable filter, no recorded edges): the
// blast-radius walk itself.  Per 32-candidate chunk the common (all visited) case is one coalesced load, one hash, one 16-byte
// shared-memory probe and one vote; lanes past the end of a row carry EMPTY, which the probe reports as present.
template <class Store, class idx_t>
__device__ __forceinline__ bool expand_window_lean(const WalkArgs &A, Store &st, const Frontier32 &f, bool single, uint32_t nrows, int32_t base, int depth1,
                                                   int lane, idx_t &tail) {
    const bool fwd = A.spec.direction == ABB_DIR_FORWARD;
    const int32_t *__restrict__ nbr = fwd ? A.g.fnbr : A.g.rnbr;
    const uint8_t *__restrict__ meta = fwd ? A.g.fmeta : A.g.rmeta;
    // row start / length of frontier node `lane` of the window (uniform when the window holds a single node)
    const uint32_t my_s = fwd ? f.sF : f.sR;
    uint32_t my_d = f.total;
    if (!single) {
        const uint32_t nx = __shfl_down_sync(FULL, f.excl, 1);
        my_d = (lane == 31 ? f.total : nx) - f.excl;
    }
    // insert the new nodes of one 32-candidate chunk in lane order; false = the store overflowed (already cleared)
    auto insert_chunk = [&](int32_t v, bool unv, const uint8_t *mp, int32_t parent) -> bool {
        if (__ballot_sync(FULL, unv) == 0u) return true;
        // One row, unfiltered: the first occurrence of a neighbour in the row carries FIRST_PAIR — a later occurrence (same or later
        // chunk) never speaks, an earlier one has already been inserted — so no match.any and no re-probe are needed.
        const uint32_t m = unv ? static_cast<uint32_t>(__ldg(mp)) : 0u;
        const bool leader = unv && (m & ABB_META_FIRST_PAIR);
        uint32_t tok = NO_TOK;
        const bool isnew = leader && st.test_and_set(v, tok);
        const unsigned nm = __ballot_sync(FULL, isnew);
        const int cnt = __popc(nm);
        if (cnt) {
            if (tail + cnt > st.qcap()) {
                if (isnew) st.unset(v, tok);
                st.clear(tail, lane);
                return false;
            }
            if (isnew) st.put(tail + __popc(nm & lanemask_lt(lane)), v, parent, tok, depth1);
            tail += cnt;
        }
        __syncwarp();
        return true;
    };
    for (uint32_t j = 0; j < nrows; j++) {
        const uint32_t s = single ? my_s : __shfl_sync(FULL, my_s, j);
        const uint32_t tot = single ? my_d : __shfl_sync(FULL, my_d, j);
        if (tot == 0) continue;
        const int32_t *pl = nbr + s + lane;                 // this lane's column of the row, bumped by 64 per step
        const uint32_t rem0 = tot - min(tot, static_cast<uint32_t>(lane));      // entries at or after this lane's column
        // two chunks per step (two independent loads and probes per lane, one vote), the next two already in flight
        int32_t v1 = rem0 > 0u ? __ldg(pl) : EMPTY;
        int32_t v2 = rem0 > 32u ? __ldg(pl + 32) : EMPTY;
        for (uint32_t c0 = 0; c0 < tot; c0 += 64, pl += 64) {
            const int32_t va = v1, vb = v2;
            v1 = (rem0 > c0 + 64u) ? __ldg(pl + 64) : EMPTY;
            v2 = (rem0 > c0 + 96u) ? __ldg(pl + 96) : EMPTY;
            const bool ua = Store::kGlobal ? (va >= 0 && !st.contains(va)) : !st.contains(va);
            const bool ub = Store::kGlobal ? (vb >= 0 && !st.contains(vb)) : !st.contains(vb);
            if (__ballot_sync(FULL, ua | ub) == 0u) continue;
            const uint8_t *mp = meta + (pl - nbr);
            if (!insert_chunk(va, ua, mp, base + static_cast<int32_t>(j))) return false;
            if (!insert_chunk(vb, ub, mp + 32, base + static_cast<int32_t>(j))) return false;
        }
    }
    return true;
}

// ---------------------------------------------------------------- one query
// Returns 1 when done; 0 when the query outgrew the store (the caller re-queues it for the next tier); -1 when a forecast says
// it is a HEAVY one (its roots or one level alone hold more candidates than this tier's queue).
template <class Store, bool NEED_META, bool BUDGET>
__device__ int walk_one(const WalkArgs &A, Store &st, int64_t q, int lane, uint32_t *hist_bins) {
    const GraphView &g = A.g;
    const abb_walk_spec &sp = A.spec;
    const abb_walk_io &io = A.io;
    const uint32_t fl = sp.flags;
    using idx_t = decltype(st.qcap());

    int64_t r0 = q, r1 = q + 1;
    if (io.root_off) { r0 = io.root_off[q]; r1 = io.root_off[q + 1]; }

    idx_t tail = 0;
    long long nvis = 0;
    // ---- seed the queue with the roots, in order (container.py:465-472)
    for (int64_t rb = r0; rb < r1; rb += 32) {
        int64_t ri = rb + lane;
        int32_t r = ri < r1 ? __ldg(io.roots + ri) : -1;
        bool ok = ri < r1 && r >= 0 && r < g.n;
        if (ok && (fl & ABB_WALK_REAL_ROOTS) && __ldg(g.ntype + r) == ABB_NODE_GHOST) ok = false;
        unsigned om = __ballot_sync(FULL, ok);
        int cnt = __popc(om), rk = __popc(om & lanemask_lt(lane));
        if (tail + cnt > st.qcap()) { st.clear(tail, lane); return false; }
        uint32_t tok = NO_TOK;
        if (fl & ABB_WALK_MARK_ROOTS) {
            unsigned mm = __match_any_sync(FULL, ok ? r : (-2 - lane));
            bool isnew = false;
            if (ok && (__ffs(mm) - 1) == lane) isnew = st.test_and_set(r, tok);
            nvis += __popc(__ballot_sync(FULL, isnew));
        }
        if (ok) st.put(tail + rk, r, -1, tok, 0);
        tail += cnt;
        __syncwarp();
    }
    const idx_t n_roots = tail;
    if (!Store::kGlobal && n_roots > 0 && n_roots <= 32 && (sp.max_depth < 0 || sp.max_depth > 0)) {
        // cheap size forecast: if the roots alone have more candidates than this tier's queue holds, the walk will
        // almost surely outgrow it — hand it to the next tier now instead of after filling the queue
        uint32_t deg = 0;
        if (lane < n_roots) {
            const int32_t r = st.q_get(lane);
            if (sp.direction & 1) deg += __ldg(g.foff + r + 1) - __ldg(g.foff + r);
            if (sp.direction & 2) deg += __ldg(g.roff + r + 1) - __ldg(g.roff + r);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) deg += __shfl_xor_sync(FULL, deg, o);
        if (deg > static_cast<uint32_t>(st.qcap())) { st.clear(tail, lane); return -1; }
    }
    int qflags = 0;
    if (n_roots == 0) qflags |= ABB_QFLAG_NO_ROOT;
    const int32_t target = (fl & ABB_WALK_TARGET) ? __ldg(io.targets + q) : -1;

    idx_t lvl_begin = 0, lvl_end = tail, exp_end = 0;
    int depth = 0, maxd = n_roots ? A.depth_bias : 0;
    long long rec_edges = 0;      // passing candidates recorded (== edge_count while under budget)
    bool stop = false;

    while (lvl_begin < lvl_end && (sp.max_depth < 0 || depth < sp.max_depth) && !stop) {
        if (depth + 1 > st.max_level()) { st.clear(tail, lane); return false; }
        if (A.overflow != nullptr && lvl_end - lvl_begin > 1) {
            // level forecast: if this level's candidates exceed several times the room left in this tier's queue the walk
            // will almost surely outgrow it — move it to the next tier now rather than after filling the queue
            unsigned long long cand = 0;
            for (idx_t base = lvl_begin; base < lvl_end; base += 32) {
                const idx_t fi = base + lane;
                if (fi < lvl_end) {
                    const int32_t u = st.q_get(fi);
                    if (sp.direction & 1) cand += __ldg(g.foff + u + 1) - __ldg(g.foff + u);
                    if (sp.direction & 2) cand += __ldg(g.roff + u + 1) - __ldg(g.roff + u);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(FULL, cand, o);
            // (shared-memory tier: only a gross excess counts — on clique-shaped graphs most candidates are revisits, and a wrong
            //  guess costs at most one queue-full of discoveries there)
            const unsigned long long factor = Store::kGlobal ? 3ull : 64ull;
            if (cand > factor * static_cast<unsigned long long>(st.qcap() - tail) + 64ull) { st.clear(tail, lane); return -1; }
        }
        for (idx_t base = lvl_begin; base < lvl_end && !stop; base += 32) {
            const bool single = (lvl_end - base) == 1;
            Frontier32 f; f.sF = f.dF = f.sR = f.excl = f.total = 0;
            if (single) {
                const int32_t u = st.q_get(base);
                uint32_t dR = 0;
                if (sp.direction & 1) { const uint32_t a = __ldg(g.foff + u), b = __ldg(g.foff + u + 1); f.sF = a; f.dF = b - a; }
                if (sp.direction & 2) { const uint32_t a = __ldg(g.roff + u), b = __ldg(g.roff + u + 1); f.sR = a; dR = b - a; }
                f.total = f.dF + dR;
            } else {
                idx_t fi = base + lane;
                bool fvalid = fi < lvl_end;
                int32_t u = fvalid ? st.q_get(fi) : 0;
                f = load_frontier(g, sp.direction, u, fvalid);
            }
            exp_end = (base + 32 < lvl_end) ? base + 32 : lvl_end;
            const uint32_t nrows = static_cast<uint32_t>(exp_end - base);
            if (!BUDGET && !(fl & ABB_WALK_TARGET) && (single || f.total >= 4u * nrows)) {
                // row mode (long rows): one frontier node after the other, in queue order
                if (!NEED_META && sp.direction != ABB_DIR_BOTH && !(fl & ABB_WALK_EDGES)) {
                    if (!expand_window_lean<Store, idx_t>(A, st, f, single, nrows, static_cast<int32_t>(base), depth + 1, lane, tail)) return false;
                    continue;
                }
                if (single) {
                    if (!expand_row<Store, NEED_META, idx_t>(A, st, f.sF, f.dF, f.sR, f.total, static_cast<int32_t>(base), depth + 1, lane, tail, rec_edges)) return false;
                } else {
                    const uint32_t incl = f.excl, tot_all = f.total;
                    for (uint32_t j = 0; j < nrows; j++) {
                        const uint32_t sF = __shfl_sync(FULL, f.sF, j), dF = __shfl_sync(FULL, f.dF, j), sR = __shfl_sync(FULL, f.sR, j);
                        const uint32_t e0 = __shfl_sync(FULL, incl, j), e1 = (j + 1 < 32) ? __shfl_sync(FULL, incl, (j + 1) & 31) : tot_all;
                        const uint32_t tot = ((j + 1 < 32) ? e1 : tot_all) - e0;
                        if (tot == 0) continue;
                        if (!expand_row<Store, NEED_META, idx_t>(A, st, sF, dF, sR, tot, static_cast<int32_t>(base + j), depth + 1, lane, tail, rec_edges)) return false;
                    }
                }
                continue;
            }
            // software pipeline: the next chunk's neighbour loads are in flight while this chunk's visited-set round trip resolves
            Cand nxt;
            if (f.total) nxt = single ? fetch_single<true, false>(g, f.sF, f.dF, f.sR, f.total, lane) : fetch_cand<true, false>(g, f, lane);
            for (uint32_t c0 = 0; c0 < f.total && !stop; c0 += 32) {
                Cand c = nxt;
                if (c0 + 32 < f.total)
                    nxt = single ? fetch_single<true, false>(g, f.sF, f.dF, f.sR, f.total, c0 + 32 + lane) : fetch_cand<true, false>(g, f, c0 + 32 + lane);
                bool pass = cand_passes(sp, c);
                unsigned pm = __ballot_sync(FULL, pass);
                if (BUDGET && sp.max_edges >= 0) {
                    // edge_count += 1; if edge_count > max_edges: truncated, break  (container.py:507-510)
                    long long my_ec = rec_edges + __popc(pm & (lanemask_lt(lane) | (1u << lane)));
                    unsigned over = __ballot_sync(FULL, pass && my_ec > sp.max_edges);
                    if (over) {
                        int fo = __ffs(over) - 1;
                        pass = pass && lane < fo;
                        pm &= lanemask_lt(fo);
                        qflags |= ABB_QFLAG_TRUNCATED;
                        stop = true;
                    }
                }
                if (BUDGET || (fl & ABB_WALK_EDGES)) rec_edges += __popc(pm);
                // The first lane among duplicates of a neighbour inside the chunk speaks for it.  When the whole chunk comes
                // from ONE row of an unfiltered walk, "first of its neighbour in the row" is the precomputed FIRST_PAIR bit:
                // a later duplicate can never discover anything (its first occurrence already did), so no match.any is needed.
                bool leader;
                const int owner0 = __shfl_sync(FULL, c.owner, 0);      // every lane takes part: no shuffle inside a short-circuit
                const bool same_owner = __all_sync(FULL, !c.active || c.owner == owner0);
                const bool one_row = !NEED_META && sp.direction != ABB_DIR_BOTH && (single || same_owner);
                if (one_row) {
                    leader = pass && (c.meta & ABB_META_FIRST_PAIR);
                } else {
                    unsigned mm = __match_any_sync(FULL, pass ? c.nbr : (-2 - lane));
                    leader = pass && (__ffs(mm) - 1) == lane;
                }
                bool isnew = false;
                uint32_t tok = NO_TOK;
                if (BUDGET && sp.max_nodes >= 0) {
                    // if neighbor in visited: continue; if len(visited) >= max_nodes: truncated; continue (container.py:515-519)
                    bool unseen = leader && !st.contains(c.nbr);
                    unsigned um = __ballot_sync(FULL, unseen);
                    long long before = nvis + __popc(um & lanemask_lt(lane));
                    bool allowed = unseen && before < sp.max_nodes;
                    if (um && (nvis + __popc(um)) > sp.max_nodes) qflags |= ABB_QFLAG_TRUNCATED;
                    // a non-leader duplicate of a denied neighbour is denied too (still unvisited, budget still full)
                    if (allowed) isnew = st.test_and_set(c.nbr, tok);
                } else if (leader) {
                    // global tier: a revisit is answered by a plain L2 load instead of an atomic round trip
                    if (!Store::kGlobal || !st.contains(c.nbr)) isnew = st.test_and_set(c.nbr, tok);
                }
                unsigned nm = __ballot_sync(FULL, isnew);
                int cnt = __popc(nm);
                if (cnt) {
                    if (tail + cnt > st.qcap()) {
                        if (isnew) st.unset(c.nbr, tok);   // inserted this chunk but never queued
                        st.clear(tail, lane);
                        return false;
                    }
                    if (isnew) st.put(tail + __popc(nm & lanemask_lt(lane)), c.nbr, static_cast<int32_t>(base + c.owner), tok, depth + 1);
                    tail += cnt;
                    if (BUDGET) nvis += cnt;
                    if ((fl & ABB_WALK_TARGET) && __any_sync(FULL, isnew && c.nbr == target)) { qflags |= ABB_QFLAG_TARGET_FOUND; stop = true; }
                }
                __syncwarp();
            }
        }
        depth++;
        lvl_begin = lvl_end; lvl_end = tail;
        if (lvl_end > lvl_begin) maxd = depth + A.depth_bias;
    }

    // ---- emit the slice: reserve a contiguous range, then copy (optionally type-filtered)
    const idx_t first = (fl & ABB_WALK_OMIT_ROOTS) ? n_roots : 0;
    const bool filtered = sp.emit_types != 0xFFFFFFFFu;
    long long count = static_cast<long long>(tail - first);
    if (filtered) {
        count = 0;
        for (idx_t i = first; i < tail; i += 32) {
            idx_t k = i + lane;
            bool ok = k < tail && type_emitted(sp.emit_types, __ldg(g.ntype + st.q_get(k)));
            count += __popc(__ballot_sync(FULL, ok));
        }
    }
    unsigned long long start = 0;
    const unsigned long long reserve = A.slice_align > 1 ? ((static_cast<unsigned long long>(count) + A.slice_align - 1) / A.slice_align) * A.slice_align
                                                         : static_cast<unsigned long long>(count);
    if (lane == 0) {
        start = atomicAdd(io.totals, reserve);
        if (A.slice_align > 1) atomicAdd(io.totals + 2, static_cast<unsigned long long>(count));   // ints really written (host path only: totals has 3 slots there)
    }
    start = __shfl_sync(FULL, start, 0);
    const bool fits = static_cast<long long>(start + reserve) <= io.node_cap;
    if ((fl & ABB_WALK_HIST) && lane < ABB_N_ENTITY_TYPES) hist_bins[lane] = 0;
    __syncwarp();
    if (fits || (fl & ABB_WALK_HIST)) {
        unsigned long long w = start;
        for (idx_t i = first; i < tail; i += 32) {
            idx_t k = i + lane;
            bool in = k < tail;
            int32_t node = in ? st.q_get(k) : 0;
            uint8_t t = 0;
            if (in && (filtered || (fl & ABB_WALK_HIST))) t = __ldg(g.ntype + node);
            bool ok = in && (!filtered || type_emitted(sp.emit_types, t));
            unsigned okm = __ballot_sync(FULL, ok);
            if (ok && fits) {
                unsigned long long pos = w + __popc(okm & lanemask_lt(lane));
                io.nodes[pos] = node;
                if (fl & ABB_WALK_PARENTS) io.parent[pos] = st.par_get(k);
                if (fl & ABB_WALK_DEPTHS) io.depth[pos] = st.dep_get(k) + A.depth_bias;
            }
            w += __popc(okm);
            if (fl & ABB_WALK_HIST) {
                // roots are never counted (impact_of excludes the source, container.py:265)
                const bool hv = ok && (k >= n_roots || A.hist_roots) && t < ABB_N_ENTITY_TYPES;
                // one lane per distinct type in the chunk adds the whole group's count to that type's bin
                const unsigned tm = __match_any_sync(FULL, hv ? static_cast<int>(t) : (-1 - lane));
                if (hv && (__ffs(tm) - 1) == lane) hist_bins[t] += __popc(tm);
                __syncwarp();
            }
        }
    }
    __syncwarp();
    if ((fl & ABB_WALK_HIST) && lane < ABB_N_ENTITY_TYPES) io.q_hist[q * ABB_N_ENTITY_TYPES + lane] = hist_bins[lane];

    // ---- recorded edges: re-scan the expanded prefix in the same order (container.py:512-513)
    unsigned long long estart = 0;
    if (fl & ABB_WALK_EDGES) {
        if (lane == 0) estart = atomicAdd(io.totals + 1, static_cast<unsigned long long>(rec_edges));
        estart = __shfl_sync(FULL, estart, 0);
        if (static_cast<long long>(estart) + rec_edges <= io.edge_cap) {
            long long written = 0;
            for (idx_t base = 0; base < exp_end && written < rec_edges; base += 32) {
                idx_t fi = base + lane;
                bool fvalid = fi < exp_end && (sp.max_depth < 0 || st.dep_get(fi) < sp.max_depth);
                int32_t u = fvalid ? st.q_get(fi) : 0;
                Frontier32 f = load_frontier(g, sp.direction, u, fvalid);
                for (uint32_t c0 = 0; c0 < f.total && written < rec_edges; c0 += 32) {
                    Cand c = fetch_cand<true, true>(g, f, c0 + lane);
                    bool pass = cand_passes(sp, c);
                    unsigned pm = __ballot_sync(FULL, pass);
                    long long pos = written + __popc(pm & lanemask_lt(lane));
                    if (pass && pos < rec_edges) io.edges[estart + pos] = c.eid;
                    written += __popc(pm);
                }
            }
        }
    }

    // canonical mode: a member whose own source was reached would not list it -> mark it for an individual traversal
    if (A.mem_off) {
        const int64_t m0 = A.mem_off[q], m1 = A.mem_off[q + 1];
        for (int64_t m = m0 + lane; m < m1; m += 32)
            if (A.mem_state[m] == 0 && st.contains(__ldg(A.mem_src + m))) A.mem_state[m] = 2;
    }
    if (lane == 0) {
        io.q_start[q] = static_cast<int64_t>(start);
        io.q_count[q] = static_cast<int32_t>(count);
        io.q_maxd[q] = maxd;
        io.q_flags[q] = qflags;
        if (fl & ABB_WALK_EDGES) { io.q_estart[q] = static_cast<int64_t>(estart); io.q_ecount[q] = rec_edges; }
    }
    st.clear(tail, lane);
    return true;
}

constexpr int WORK_CHUNK = 4;

__device__ __forceinline__ int64_t next_chunk(unsigned long long *ctl, int lane) {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(ctl, static_cast<unsigned long long>(WORK_CHUNK));
    return static_cast<int64_t>(__shfl_sync(FULL, v, 0));
}

// ---------------------------------------------------------------- kernels
// Who takes S1's heavy hand-offs?  The big block tier walks one of them an order of magnitude faster than a single warp, but only one
// per SM at a time; with thousands of them the warp tier's 40 walks per SM win on throughput.  One thread decides from the count.
__global__ void route_heavy_kernel(const unsigned long long *heavy, unsigned long long *to_big, unsigned long long *to_warp, unsigned long long limit) {
    const unsigned long long h = *heavy;
    if (h <= limit) { *to_big = h; *to_warp = 0ull; } else { *to_big = 0ull; *to_warp = h; }
}

template <int H, int Q, bool PAR, bool NEED_META, bool BUDGET, int WARPS, int MIN_BLOCKS>
__global__ void __launch_bounds__(WARPS * 32, MIN_BLOCKS) walk_smem_kernel(const WalkArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_hist[WARPS][ABB_N_ENTITY_TYPES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    using Store = SmemStore<H, Q, PAR>;
    constexpr int kStride = (Store::kBytes + 15) & ~15;
    Store st(smem + warp * kStride);
    const int64_t nq = list_count(A);
    if (nq == 0) return;
    st.init(lane);
    for (;;) {
        int64_t c = next_chunk(A.ctl, lane);
        if (c >= nq) break;
        int64_t ce = c + WORK_CHUNK < nq ? c + WORK_CHUNK : nq;
        for (int64_t i = c; i < ce; i++) {
            const int64_t q = list_item(A, i);
            const int r = walk_one<Store, NEED_META, BUDGET>(A, st, q, lane, s_hist[warp]);
            if (r <= 0 && lane == 0) list_append(A, q, r < 0);
        }
    }
}

template <bool NEED_META, bool BUDGET, int MIN_BLOCKS>
__global__ void __launch_bounds__(128, MIN_BLOCKS) walk_global_kernel(const WalkArgs A) {
    __shared__ uint32_t s_hist[4][ABB_N_ENTITY_TYPES];
    const int lane = threadIdx.x & 31;
    const int64_t slot = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    GlobalStore st;
    st.bits = A.g_bitmap + slot * A.g_words;
    st.queue = A.g_queue + slot * A.g_qcap;
    st.par = (A.g_par && (A.spec.flags & ABB_WALK_PARENTS)) ? A.g_par + slot * A.g_qcap : nullptr;
    st.dep = (A.spec.flags & (ABB_WALK_DEPTHS | ABB_WALK_EDGES)) ? A.g_dep + slot * A.g_qcap : nullptr;
    st.cap = static_cast<int>(A.g_qcap);
    const int64_t nq = list_count(A);
    for (;;) {
        unsigned long long v = 0;
        if (lane == 0) v = atomicAdd(A.ctl, 1ull);
        int64_t i = static_cast<int64_t>(__shfl_sync(FULL, v, 0));
        if (i >= nq) break;
        const int64_t q = list_item(A, i);          // a two-ended list hands out its heavy (front) part first
        if (walk_one<GlobalStore, NEED_META, BUDGET>(A, st, q, lane, s_hist[threadIdx.x >> 5]) <= 0) {
            if (lane == 0) {
                if (A.overflow) list_append(A, q, true);   // G1 -> GX
                else atomicExch(A.ctl + 2, 1ull);   // GX: cannot happen unless the whole-graph scratch is undersized
            }
        }
    }
}

}  // namespace abb
