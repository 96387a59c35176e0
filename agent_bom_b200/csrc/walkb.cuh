// walkb.cuh — block-cooperative tiers of the ordered batched BFS (sm_100a).
//
// walk.cuh gives one WARP one query.  That is the right grain for the millions of tiny blast radii of an estate, but a
// query that reaches thousands of nodes becomes a serial chain of dependent 32-candidate steps (≈2 ms for a 30 K-node
// walk), which is both the latency tail of a batch and the reason the global-bitmap tier sat at a few percent of HBM
// bandwidth.  Here a whole thread BLOCK owns one query:
//
//   * the visited set is an open-addressing hash table in SHARED memory (one 32-bit word per slot: node id in the low
//     `id_bits`, a claim in the bits above; 0 = settled), so a revisit costs one shared-memory probe instead of an L2 atomic;
//   * a level is consumed in queue order in TILES: 32 frontier nodes are flattened exactly as in walk.cuh, the flattened
//     candidate range of the tile is cut into W contiguous pieces (one per warp, ≤ 128 candidates each, held in registers),
//     every warp probes its piece, and an unvisited neighbour is CLAIMED with atomicMin(word, (warp+1) << id_bits | id):
//     the lowest warp — the earliest position in the reference's scan order — wins, duplicates inside one warp's piece
//     are resolved in the warp (first-of-pair bit / match.any for the chunk, "claim is already mine" for a later chunk);
//   * after a block barrier the winners settle their slots, a prefix sum over the per-warp winner counts gives every
//     winner its queue position — the same (queue position, row position) append order as the sequential deque loop
//     (reference graph/container.py:230-279,367-391,411-436,438-538) — and the next tile starts.
//
// Two instantiations: a mid tier (8 warps, queue + table in shared memory, several blocks per SM) and a big tier
// (32 warps, 192 KB table, queue in an L2-resident global scratch slot, one block per SM).  Budgeted / target /
// edge-recording walks keep to the warp tiers.  A query that outgrows a block tier is re-queued for the next tier.
#pragma once
#include "walk.cuh"

namespace abb {

struct BlockTier {
    uint32_t slots;               // hash slots (shared memory)
    int32_t qcap;                 // queue entries
    int32_t id_bits;              // node id bits of a slot word
    int32_t *gq, *gpar, *gdep;    // [grid][qcap] global queue scratch (big tier); null for the shared-memory queue
};

constexpr uint32_t B_EMPTY = 0xFFFFFFFFu;
constexpr int B_CH = 4;           // 32-candidate chunks a warp holds per tile
constexpr int B_CAP = 32 * B_CH;

template <bool QSM, bool PAR>
struct BlockStore {
    uint32_t *tab; uint32_t slots; int idb; uint32_t idmask;
    int32_t *queue; int32_t *par; uint8_t *dep8; int32_t *dep32; int qcap;
    __device__ __forceinline__ uint32_t hash(int32_t v) const { return __umulhi(static_cast<uint32_t>(v) * 0x9E3779B1u, slots); }
    __device__ __forceinline__ uint32_t next(uint32_t h) const { return (h + 1u == slots) ? 0u : h + 1u; }
    __device__ __forceinline__ int32_t q_get(int i) const { return queue[i]; }
    __device__ __forceinline__ void put(int i, int32_t node, int32_t p, int d) {
        queue[i] = node;
        if (QSM) dep8[i] = static_cast<uint8_t>(d); else if (dep32) dep32[i] = d;
        if (PAR && par) par[i] = p;
    }
    __device__ __forceinline__ int32_t par_get(int i) const { return (PAR && par) ? par[i] : -1; }
    __device__ __forceinline__ int dep_get(int i) const { return QSM ? dep8[i] : (dep32 ? dep32[i] : 0); }
    __device__ __forceinline__ int max_level() const { return QSM ? 255 : 0x7FFFFFF0; }
    __device__ __forceinline__ bool contains(int32_t v) const {
        uint32_t h = hash(v);
        for (;;) {
            const uint32_t w = tab[h];
            if (w == B_EMPTY) return false;
            if ((w & idmask) == static_cast<uint32_t>(v)) return true;
            h = next(h);
        }
    }
    // settled insert (seeding; callers make sure no two lanes insert the same id at once); true when newly inserted
    __device__ __forceinline__ bool insert_settled(int32_t v) {
        uint32_t h = hash(v);
        for (;;) {
            uint32_t w = tab[h];
            if (w == B_EMPTY) { w = atomicCAS(&tab[h], B_EMPTY, static_cast<uint32_t>(v)); if (w == B_EMPTY) return true; }
            if ((w & idmask) == static_cast<uint32_t>(v)) return false;
            h = next(h);
        }
    }
};

// candidate ci of the flattened window, active only below `limit` (the end of this warp's piece)
__device__ __forceinline__ Cand fetch_cand_lim(const GraphView &g, const Frontier32 &f, uint32_t ci, uint32_t limit) {
    Cand c; c.active = ci < limit;
    const uint32_t cc = ci < f.total ? ci : (f.total - 1);
    int j = 0;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const int t = j + s;
        const uint32_t v = __shfl_sync(FULL, f.excl, t & 31);
        if (t < 32 && v <= cc) j = t;
    }
    const uint32_t ex = __shfl_sync(FULL, f.excl, j), sF = __shfl_sync(FULL, f.sF, j), dF = __shfl_sync(FULL, f.dF, j), sR = __shfl_sync(FULL, f.sR, j);
    const uint32_t k = cc - ex;
    c.owner = j; c.meta = ABB_META_TRAVERSABLE; c.eid = 0; c.nbr = 0;
    if (c.active) {
        if (k < dF) { const uint32_t p = sF + k; c.nbr = __ldg(g.fnbr + p); c.meta = __ldg(g.fmeta + p); }
        else { const uint32_t p = sR + (k - dF); c.nbr = __ldg(g.rnbr + p); c.meta = __ldg(g.rmeta + p); }
    }
    return c;
}

// sum of one value per warp, published through s_cnt; every thread returns the total.  Two barriers.
template <int W>
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long *s_red, int lane, int warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();                 // s_red may still be read by the previous user
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    unsigned long long x = lane < W ? s_red[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// One query by the whole block.  Returns false (block-uniform) when the query outgrew this tier; the table is clean on return.
template <int W, bool QSM, bool PAR, bool NEED_META>
__device__ bool block_walk_one(const WalkArgs &A, BlockStore<QSM, PAR> &st, int64_t q, uint32_t *s_cnt, unsigned long long *s_red,
                               uint32_t (*s_hist)[ABB_N_ENTITY_TYPES], unsigned long long *s_b) {
    const GraphView &g = A.g;
    const abb_walk_spec &sp = A.spec;
    const abb_walk_io &io = A.io;
    const uint32_t fl = sp.flags;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t myclaim = static_cast<uint32_t>(warp + 1);
    constexpr int NT = W * 32;

    auto wipe = [&]() {
        __syncthreads();
        for (uint32_t i = tid; i < st.slots; i += NT) st.tab[i] = B_EMPTY;
        __syncthreads();
    };

    int64_t r0 = q, r1 = q + 1;
    if (io.root_off) { r0 = io.root_off[q]; r1 = io.root_off[q + 1]; }

    // ---- seed the queue with the roots, in order (container.py:465-472): warp 0, the others wait
    if (warp == 0) {
        int tail0 = 0; bool ovf = false;
        for (int64_t rb = r0; rb < r1 && !ovf; rb += 32) {
            const int64_t ri = rb + lane;
            const int32_t r = ri < r1 ? __ldg(io.roots + ri) : -1;
            bool ok = ri < r1 && r >= 0 && r < g.n;
            if (ok && (fl & ABB_WALK_REAL_ROOTS) && __ldg(g.ntype + r) == ABB_NODE_GHOST) ok = false;
            const unsigned om = __ballot_sync(FULL, ok);
            const int cnt = __popc(om), rk = __popc(om & lanemask_lt(lane));
            if (tail0 + cnt > st.qcap) { ovf = true; break; }
            if (fl & ABB_WALK_MARK_ROOTS) {
                const unsigned mm = __match_any_sync(FULL, ok ? r : (-2 - lane));
                if (ok && (__ffs(mm) - 1) == lane) st.insert_settled(r);
            }
            if (ok) st.put(tail0 + rk, r, -1, 0);
            tail0 += cnt;
            __syncwarp();
        }
        if (lane == 0) { s_b[0] = static_cast<unsigned long long>(tail0); s_b[1] = ovf ? 1ull : 0ull; }
    }
    __syncthreads();
    int tail = static_cast<int>(s_b[0]);
    if (s_b[1]) { wipe(); return false; }
    const int n_roots = tail;
    int qflags = n_roots == 0 ? ABB_QFLAG_NO_ROOT : 0;

    int lvl_begin = 0, lvl_end = tail, depth = 0, maxd = n_roots ? A.depth_bias : 0;

    while (lvl_begin < lvl_end && (sp.max_depth < 0 || depth < sp.max_depth)) {
        if (depth + 1 > st.max_level()) { wipe(); return false; }
        // level forecast (also warms the row offsets): a level whose candidates exceed several times the room left will outgrow the tier
        // (a small frontier cannot: skip the two barriers)
        if (lvl_end - lvl_begin >= 64) {
            unsigned long long cand = 0;
            for (int fi = lvl_begin + tid; fi < lvl_end; fi += NT) {
                const int32_t u = st.q_get(fi);
                if (sp.direction & 1) cand += __ldg(g.foff + u + 1) - __ldg(g.foff + u);
                if (sp.direction & 2) cand += __ldg(g.roff + u + 1) - __ldg(g.roff + u);
            }
            cand = block_sum<W>(cand, s_red, lane, warp);
            if (cand > 8ull * static_cast<unsigned long long>(st.qcap - tail) + 64ull) { wipe(); return false; }
        }
        // One TILE: every warp holds a contiguous piece [my0, myend) of the flattened candidates of window `f` (whose first frontier node
        // sits at queue position `wbase`); pieces are ordered by warp.  Returns false (block-uniform) when the queue would overflow.
        auto tile = [&](const Frontier32 &f, uint32_t my0, uint32_t myend, int wbase) -> bool {
            const int nch = my0 < myend ? static_cast<int>((myend - my0 + 31u) >> 5) : 0;
            // ---- phase 1: fetch (all chunks in flight), probe, claim
            int32_t cv[B_CH]; uint32_t cmeta[B_CH]; uint32_t cslot[B_CH]; int cown[B_CH]; bool cact[B_CH]; bool ccont[B_CH];
#pragma unroll
            for (int c = 0; c < B_CH; c++) {
                cv[c] = 0; cmeta[c] = 0; cslot[c] = 0; cown[c] = 0; cact[c] = false; ccont[c] = false;
                if (c < nch) {
                    const Cand cd = fetch_cand_lim(g, f, my0 + c * 32 + lane, myend);
                    cv[c] = cd.nbr; cmeta[c] = cd.meta; cown[c] = cd.owner; cact[c] = cd.active;
                }
            }
#pragma unroll
            for (int c = 0; c < B_CH; c++) {
                if (c < nch) {
                    const int32_t v = cv[c];
                    bool pass = cact[c];
                    if (NEED_META) pass = pass && ((sp.rel_mask >> (cmeta[c] & ABB_META_REL_MASK)) & 1u) && (!(fl & ABB_WALK_TRAVERSABLE_ONLY) || (cmeta[c] & ABB_META_TRAVERSABLE));
                    // first occurrence of a neighbour inside the chunk speaks for it (see walk.cuh)
                    bool leader;
                    const int owner0 = __shfl_sync(FULL, cown[c], 0);
                    const bool same_owner = __all_sync(FULL, !cact[c] || cown[c] == owner0);
                    if (!NEED_META && sp.direction != ABB_DIR_BOTH && same_owner) {
                        leader = pass && (cmeta[c] & ABB_META_FIRST_PAIR);
                    } else {
                        const unsigned mm = __match_any_sync(FULL, pass ? v : (-2 - lane));
                        leader = pass && (__ffs(mm) - 1) == lane;
                    }
                    if (leader) {
                        const uint32_t mine = (myclaim << st.idb) | static_cast<uint32_t>(v);
                        uint32_t h = st.hash(v);
                        for (;;) {
                            uint32_t w = st.tab[h];
                            if (w == B_EMPTY) {
                                w = atomicCAS(&st.tab[h], B_EMPTY, mine);
                                if (w == B_EMPTY) { ccont[c] = true; cslot[c] = h; break; }
                            }
                            if ((w & st.idmask) == static_cast<uint32_t>(v)) {
                                const uint32_t cl = w >> st.idb;
                                if (cl != 0u && cl != myclaim) {         // claimed in this tile by another warp: the lower warp wins
                                    const uint32_t old = atomicMin(&st.tab[h], mine);
                                    if ((old >> st.idb) > myclaim) { ccont[c] = true; cslot[c] = h; }
                                }
                                break;                                    // settled (visited) or already claimed by an earlier chunk of this warp
                            }
                            h = st.next(h);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- phase 2: winners settle their slot; per-warp winner count
            unsigned wm[B_CH]; int mycnt = 0;
#pragma unroll
            for (int c = 0; c < B_CH; c++) {
                wm[c] = 0u;
                if (c < nch) {
                    bool win = false;
                    if (ccont[c]) {
                        // a losing contender may read the slot while the winner settles it: either value tells it "not mine".  Both
                        // sides go through atomics so that this intended overlap is not a formal data race (compute-sanitizer racecheck)
                        win = (atomicAdd(&st.tab[cslot[c]], 0u) >> st.idb) == myclaim;
                        if (win) atomicExch(&st.tab[cslot[c]], static_cast<uint32_t>(cv[c]));
                    }
                    wm[c] = __ballot_sync(FULL, win);
                    mycnt += __popc(wm[c]);
                }
            }
            if (lane == 0) s_cnt[warp] = static_cast<uint32_t>(mycnt);
            __syncthreads();
            // ---- phase 3: ordered append
            uint32_t x = lane < W ? s_cnt[lane] : 0u, incl = x;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, s); if (lane >= s) incl += y; }
            const int total = static_cast<int>(__shfl_sync(FULL, incl, 31));
            const int mybase = static_cast<int>(__shfl_sync(FULL, incl - x, warp));
            if (total) {
                if (tail + total > st.qcap) return false;
                int run = tail + mybase;
#pragma unroll
                for (int c = 0; c < B_CH; c++) {
                    if (c < nch) {
                        if ((wm[c] >> lane) & 1u) st.put(run + __popc(wm[c] & lanemask_lt(lane)), cv[c], wbase + cown[c], depth + 1);
                        run += __popc(wm[c]);
                    }
                }
                tail += total;
            }
            return true;
        };
        for (int grp = lvl_begin; grp < lvl_end; grp += W * 32) {
            // MULTI-WINDOW tile: warp w flattens ITS OWN 32 frontier nodes; valid when every one of the W windows fits one piece
            // (<= B_CAP candidates) — a frontier of short rows is then consumed W*32 nodes per tile instead of 32.
            const int mywin = grp + warp * 32;
            const int fi_own = mywin + lane;
            const bool own_valid = fi_own < lvl_end;
            const int32_t u_own = own_valid ? st.q_get(fi_own) : 0;
            const Frontier32 f_own = load_frontier(g, sp.direction, u_own, own_valid);
            __syncthreads();                                   // s_red may still be read (forecast / previous group)
            if (lane == 0) s_red[warp] = f_own.total;
            __syncthreads();
            unsigned long long mx = lane < W ? s_red[lane] : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(FULL, mx, o); mx = y > mx ? y : mx; }
            if (mx <= static_cast<unsigned long long>(B_CAP)) {
                if (!tile(f_own, 0u, f_own.total, mywin)) { wipe(); return false; }
                continue;
            }
            // SHARED-WINDOW tiles: the W windows of the group one after the other, each cut into W pieces
            for (int win = grp; win < lvl_end && win < grp + W * 32; win += 32) {
                const int fi = win + lane;
                const bool fvalid = fi < lvl_end;
                const int32_t u = fvalid ? st.q_get(fi) : 0;
                const Frontier32 f = load_frontier(g, sp.direction, u, fvalid);     // identical in every warp
                for (uint32_t t0 = 0; t0 < f.total; t0 += W * B_CAP) {
                    const uint32_t rem = (f.total - t0 < static_cast<uint32_t>(W * B_CAP)) ? f.total - t0 : static_cast<uint32_t>(W * B_CAP);
                    const uint32_t cap_t = (((rem + W - 1) / W) + 31u) & ~31u;       // this tile's piece per warp (multiple of 32, <= B_CAP)
                    const uint32_t my0 = t0 + warp * cap_t;
                    uint32_t myend = my0 + cap_t; if (myend > t0 + rem) myend = t0 + rem;
                    if (!tile(f, my0, myend, win)) { wipe(); return false; }
                }
            }
        }
        __syncthreads();       // this level's appends are the next level's frontier
        depth++;
        lvl_begin = lvl_end; lvl_end = tail;
        if (lvl_end > lvl_begin) maxd = depth + A.depth_bias;
    }

    // ---- emit the slice
    const int first = (fl & ABB_WALK_OMIT_ROOTS) ? n_roots : 0;
    const bool filtered = sp.emit_types != 0xFFFFFFFFu;
    long long count = static_cast<long long>(tail - first);
    if (filtered) {
        unsigned long long c = 0;
        for (int k = first + tid; k < tail; k += NT) c += type_emitted(sp.emit_types, __ldg(g.ntype + st.q_get(k))) ? 1ull : 0ull;
        count = static_cast<long long>(block_sum<W>(c, s_red, lane, warp));
    }
    const unsigned long long reserve = A.slice_align > 1 ? ((static_cast<unsigned long long>(count) + A.slice_align - 1) / A.slice_align) * A.slice_align
                                                         : static_cast<unsigned long long>(count);
    __syncthreads();
    if (tid == 0) {
        s_b[2] = atomicAdd(io.totals, reserve);
        if (A.slice_align > 1) atomicAdd(io.totals + 2, static_cast<unsigned long long>(count));
    }
    if ((fl & ABB_WALK_HIST) && lane < ABB_N_ENTITY_TYPES) s_hist[warp][lane] = 0;
    __syncthreads();
    const unsigned long long start = s_b[2];
    const bool fits = static_cast<long long>(start + reserve) <= io.node_cap;
    if (fits || (fl & ABB_WALK_HIST)) {
        if (!filtered) {
            for (int cb = first + warp * 32; cb < tail; cb += NT) {
                const int k = cb + lane;
                const bool in = k < tail;
                const int32_t node = in ? st.q_get(k) : 0;
                if (in && fits) {
                    const unsigned long long pos = start + static_cast<unsigned long long>(k - first);
                    io.nodes[pos] = node;
                    if (fl & ABB_WALK_PARENTS) io.parent[pos] = st.par_get(k);
                    if (fl & ABB_WALK_DEPTHS) io.depth[pos] = st.dep_get(k) + A.depth_bias;
                }
                if (fl & ABB_WALK_HIST) {
                    const uint8_t t = in ? __ldg(g.ntype + node) : 0;
                    const bool hv = in && (k >= n_roots || A.hist_roots) && t < ABB_N_ENTITY_TYPES;
                    const unsigned tm = __match_any_sync(FULL, hv ? static_cast<int>(t) : (-1 - lane));
                    if (hv && (__ffs(tm) - 1) == lane) s_hist[warp][t] += __popc(tm);
                    __syncwarp();
                }
            }
        } else {
            // type-filtered slice: ordered compaction, W chunks per round
            unsigned long long wrun = start;
            for (int tb = first; tb < tail; tb += NT) {
                const int k = tb + tid;
                const bool in = k < tail;
                const int32_t node = in ? st.q_get(k) : 0;
                const uint8_t t = in ? __ldg(g.ntype + node) : 0;
                const bool ok = in && type_emitted(sp.emit_types, t);
                const unsigned okm = __ballot_sync(FULL, ok);
                __syncthreads();
                if (lane == 0) s_cnt[warp] = __popc(okm);
                __syncthreads();
                uint32_t x = lane < W ? s_cnt[lane] : 0u, incl = x;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) { const uint32_t y = __shfl_up_sync(FULL, incl, s); if (lane >= s) incl += y; }
                const uint32_t total = __shfl_sync(FULL, incl, 31), mybase = __shfl_sync(FULL, incl - x, warp);
                if (ok && fits) {
                    const unsigned long long pos = wrun + mybase + __popc(okm & lanemask_lt(lane));
                    io.nodes[pos] = node;
                    if (fl & ABB_WALK_PARENTS) io.parent[pos] = st.par_get(k);
                    if (fl & ABB_WALK_DEPTHS) io.depth[pos] = st.dep_get(k) + A.depth_bias;
                }
                wrun += total;
                if (fl & ABB_WALK_HIST) {
                    const bool hv = ok && (k >= n_roots || A.hist_roots) && t < ABB_N_ENTITY_TYPES;
                    const unsigned tm = __match_any_sync(FULL, hv ? static_cast<int>(t) : (-1 - lane));
                    if (hv && (__ffs(tm) - 1) == lane) s_hist[warp][t] += __popc(tm);
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();
    if ((fl & ABB_WALK_HIST) && tid < ABB_N_ENTITY_TYPES) {
        uint32_t sum = 0;
#pragma unroll 4
        for (int w = 0; w < W; w++) sum += s_hist[w][tid];
        io.q_hist[q * ABB_N_ENTITY_TYPES + tid] = sum;
    }
    // canonical mode: a member whose own source was reached would not list it -> mark it for an individual traversal
    if (A.mem_off) {
        const int64_t m0 = A.mem_off[q], m1 = A.mem_off[q + 1];
        for (int64_t m = m0 + tid; m < m1; m += NT)
            if (A.mem_state[m] == 0 && st.contains(__ldg(A.mem_src + m))) A.mem_state[m] = 2;
    }
    if (tid == 0) {
        io.q_start[q] = static_cast<int64_t>(start);
        io.q_count[q] = static_cast<int32_t>(count);
        io.q_maxd[q] = maxd;
        io.q_flags[q] = qflags;
    }
    wipe();
    return true;
}

template <int W, bool QSM, bool PAR, bool NEED_META, int MIN_BLOCKS>
__global__ void __launch_bounds__(W * 32, MIN_BLOCKS) walk_block_kernel(const WalkArgs A, const BlockTier T) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_cnt[32];
    __shared__ unsigned long long s_red[32];
    __shared__ uint32_t s_hist[W][ABB_N_ENTITY_TYPES];
    __shared__ unsigned long long s_b[4];
    const int tid = threadIdx.x;
    const int64_t nq = list_count(A);
    if (nq == 0) return;
    BlockStore<QSM, PAR> st;
    st.tab = reinterpret_cast<uint32_t *>(smem);
    st.slots = T.slots; st.idb = T.id_bits; st.idmask = (1u << T.id_bits) - 1u; st.qcap = T.qcap;
    st.dep8 = nullptr; st.dep32 = nullptr; st.par = nullptr;
    if (QSM) {
        st.queue = reinterpret_cast<int32_t *>(st.tab + T.slots);
        int32_t *nxt = st.queue + T.qcap;
        if (PAR) { st.par = nxt; nxt += T.qcap; }
        st.dep8 = reinterpret_cast<uint8_t *>(nxt);
    } else {
        const int64_t o = static_cast<int64_t>(blockIdx.x) * T.qcap;
        st.queue = T.gq + o;
        if (PAR && (A.spec.flags & ABB_WALK_PARENTS)) st.par = T.gpar + o;
        if (A.spec.flags & ABB_WALK_DEPTHS) st.dep32 = T.gdep + o;
    }
    for (uint32_t i = tid; i < T.slots; i += W * 32) st.tab[i] = B_EMPTY;
    __syncthreads();
    unsigned long long nxt = 0;
    if (tid == 0) nxt = atomicAdd(A.ctl, 1ull);
    for (;;) {
        if (tid == 0) s_b[3] = nxt;
        __syncthreads();
        const int64_t i = static_cast<int64_t>(s_b[3]);
        if (i >= nq) break;
        if (tid == 0) nxt = atomicAdd(A.ctl, 1ull);          // the next work item is fetched while this one is walked
        const int64_t q = list_item(A, i);
        const bool ok = block_walk_one<W, QSM, PAR, NEED_META>(A, st, q, s_cnt, s_red, s_hist, s_b);
        if (!ok && tid == 0) {
            if (A.overflow) list_append(A, q, A.ov_cnt_back == nullptr);     // onto the back of a two-ended list when given one
            else atomicExch(A.ctl + 2, 1ull);
        }
    }
}

}  // namespace abb
