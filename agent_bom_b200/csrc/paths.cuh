// paths.cuh — per-finding typed exposure-path pattern walk (sm_100a).
//
// Device restatement of the topology part of the reference's
// _derived_attack_paths (api/routes/graph.py:686-760) and
// _edge_relationships_for_hops (:488-503):
//
//   finding f  <-VULNERABLE_TO- vs            (in-edges of f, graph.edges order, originals only)
//   servers = [vs] if vs is a SERVER else
//             [s for s -DEPENDS_ON-> vs if s is a SERVER]
//   per server: agents = sorted(set(a for a -USES-> server if a is agent/user/service_account)) or [server]
//               ncred / ntool = # EXPOSES_CRED / PROVIDES_TOOL out-edges whose target has a node record
//   row (agent, server, vs, f) + first relationship per hop pair.
//
// One warp per finding, two passes over the same code (count, then fill at the
// scanned offsets) so rows land in the reference's emission order.  Inside a
// finding the reverse row of vs is consumed 32 entries at a time and each lane
// owns one candidate server (its own short reverse row is walked lane-serially),
// so a popular package with hundreds of dependent servers is handled 32 servers
// per step instead of one.  "Originals only" = entries without
// ABB_META_REVERSED_COPY: the reference builds its incoming/outgoing maps from
// graph.edges, which never holds the reversed copies of bidirectional edges (:698-702).
//
// Pair relationships: by_pair.setdefault((a,b)) keeps the first edge between a
// and b.  Entries of adjacency[a] with neighbour b and entries of
// reverse_adjacency[b] with neighbour a correspond one-to-one in order, so the
// answer is the relationship of the first reverse-row entry of b whose
// neighbour is a — and every pair the pattern needs is reached THROUGH such an
// entry, whose ABB_META_FIRST_PAIR bit says whether it is that first one.
#pragma once
#include "walk.cuh"

namespace abb {

constexpr int REL_USES = 1, REL_DEPENDS_ON = 2, REL_PROVIDES_TOOL = 3, REL_EXPOSES_CRED = 4, REL_VULNERABLE_TO = 9;
constexpr int ET_AGENT = 0, ET_SERVER = 1, ET_VULN = 8, ET_MISCONF = 9, ET_USER = 13, ET_SERVICE_ACCOUNT = 17;
constexpr uint32_t LONG_ROW = 64;   // reverse rows of a server longer than this are handled by the whole warp

struct PathsArgs {
    GraphView g;
    abb_paths_io io;
    const int32_t *srv_cred;    // [n_nodes] EXPOSES_CRED fan-out of server nodes (graph-constant table)
    const int32_t *srv_tool;    // [n_nodes] PROVIDES_TOOL fan-out
    // links: one per (finding, vulnerable source) in emission order
    int64_t *link_cnt;          // [n_findings+1] links per finding (count pass), then scanned into link_off
    int64_t *link_off;          // [n_findings+1]
    int32_t *link_vs;           // [n_links]
    int8_t *link_rel;           // [n_links] relationship of the pair (vs, finding)
    int64_t *link_rows;         // [n_links+1] rows per link, scanned into link_roff
    int64_t *link_toff;         // [n_links] template offset of the link's vulnerable source
    int64_t *link_roff;         // [n_links+1]
    const unsigned long long *n_links;   // device scalar
    // templates: the rows of one vulnerable source, shared by every finding attached to it
    uint8_t *need;              // [n_nodes] vs is referenced by some link
    const int32_t *ulist;       // [n_unique] vulnerable sources, node order
    const unsigned long long *n_unique;  // device scalar
    int64_t *t_cnt;             // [n_unique+1] rows per template, scanned into t_off
    int64_t *t_off;             // [n_unique+1]
    int64_t *t_off_node;        // [n_nodes] template offset by vs node
    int32_t *t_cnt_node;        // [n_nodes] template length by vs node
    int4 *t_row;                // [n_template_rows] agent, server, ncred, ntool
    int8_t *t_rel;              // [n_template_rows*2] (agent,server) and (server,vs) relationships
};

// graph-constant table: credential / tool fan-out of every server (originals only, target must have a node record)
__global__ void server_table_kernel(GraphView g, int32_t *srv_cred, int32_t *srv_tool) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t base = warp * 32; base < g.n; base += nwarps * 32) {
        const int64_t u = base + lane;
        int nc = 0, nt = 0;
        if (u < g.n && __ldg(g.ntype + u) == ET_SERVER) {
            for (uint32_t p = __ldg(g.foff + u); p < __ldg(g.foff + u + 1); p++) {
                const uint32_t m = __ldg(g.fmeta + p);
                if (m & ABB_META_REVERSED_COPY) continue;
                const int rel = m & ABB_META_REL_MASK;
                if (rel != REL_EXPOSES_CRED && rel != REL_PROVIDES_TOOL) continue;
                if (__ldg(g.ntype + __ldg(g.fnbr + p)) == ABB_NODE_GHOST) continue;
                if (rel == REL_EXPOSES_CRED) nc++; else nt++;
            }
        }
        if (u < g.n) { srv_cred[u] = nc; srv_tool[u] = nt; }
    }
}

__device__ __forceinline__ bool is_principal(uint8_t t) { return t == ET_AGENT || t == ET_USER || t == ET_SERVICE_ACCOUNT; }

// relationship of the first reverse-row entry of b whose neighbour is a (lane-serial), -1 if none
__device__ __forceinline__ int first_rel_rev(const GraphView &g, int32_t a, int32_t b) {
    for (uint32_t p = __ldg(g.roff + b), e = __ldg(g.roff + b + 1); p < e; p++)
        if (__ldg(g.rnbr + p) == a) return __ldg(g.rmeta + p) & ABB_META_REL_MASK;
    return -1;
}

// next agent of server s in ascending id-string rank above `last` (lane-serial over the reverse row); false when exhausted
__device__ __forceinline__ bool next_agent_serial(const GraphView &g, int32_t s, int32_t &last, int32_t &agent, int &rel_as) {
    int32_t best = 0x7FFFFFFF; uint32_t bm = 0;
    for (uint32_t p = __ldg(g.roff + s), e = __ldg(g.roff + s + 1); p < e; p++) {
        const uint32_t m = __ldg(g.rmeta + p);
        if ((m & ABB_META_REVERSED_COPY) || (m & ABB_META_REL_MASK) != REL_USES) continue;
        const int32_t a = __ldg(g.rnbr + p);
        if (!is_principal(__ldg(g.ntype + a))) continue;
        const int32_t r = __ldg(g.rank + a);
        if (r > last && r < best) { best = r; agent = a; bm = m; }
    }
    if (best == 0x7FFFFFFF) return false;
    last = best;
    rel_as = (bm & ABB_META_FIRST_PAIR) ? REL_USES : first_rel_rev(g, agent, s);
    return true;
}

// same, whole warp over a long reverse row
__device__ __forceinline__ bool next_agent_warp(const GraphView &g, int32_t s, int32_t &last, int32_t &agent, int &rel_as, int lane) {
    int32_t best = 0x7FFFFFFF, bn = -1; uint32_t bm = 0;
    for (uint32_t p = __ldg(g.roff + s), e = __ldg(g.roff + s + 1); p < e; p += 32) {
        const uint32_t k = p + lane;
        if (k >= e) continue;
        const uint32_t m = __ldg(g.rmeta + k);
        if ((m & ABB_META_REVERSED_COPY) || (m & ABB_META_REL_MASK) != REL_USES) continue;
        const int32_t a = __ldg(g.rnbr + k);
        if (!is_principal(__ldg(g.ntype + a))) continue;
        const int32_t r = __ldg(g.rank + a);
        if (r > last && r < best) { best = r; bn = a; bm = m; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int32_t ob = __shfl_xor_sync(FULL, best, o), on = __shfl_xor_sync(FULL, bn, o);
        const uint32_t om = __shfl_xor_sync(FULL, bm, o);
        if (ob < best) { best = ob; bn = on; bm = om; }
    }
    if (bn < 0) return false;
    last = best; agent = bn;
    rel_as = (bm & ABB_META_FIRST_PAIR) ? REL_USES : first_rel_rev(g, agent, s);
    return true;
}

// one template row: (agent, server) and the relationships of the pairs (agent,server), (server,vs)
__device__ __forceinline__ void write_row(const PathsArgs &A, int64_t row, int32_t agent, int32_t srv, int32_t, int32_t, int rel_as, int rel_sv, int) {
    A.t_row[row] = make_int4(agent, srv, __ldg(A.srv_cred + srv), __ldg(A.srv_tool + srv));
    A.t_rel[row * 2] = static_cast<int8_t>(rel_as); A.t_rel[row * 2 + 1] = static_cast<int8_t>(rel_sv);
}

// Rows of up to 32 candidate servers (lane i owns server s, -1 = none) of one (finding, vulnerable source).
// Returns the number of rows; rows are laid out in lane order from row0.
template <bool FILL>
__device__ int64_t server_chunk(const PathsArgs &A, int64_t row0, int32_t s, int rel_sv, int32_t vs, int32_t f, int rel_vf, int lane) {
    const GraphView &g = A.g;
    const bool have = s >= 0;
    const uint32_t len = have ? (__ldg(g.roff + s + 1) - __ldg(g.roff + s)) : 0;
    const bool is_long = have && len > LONG_ROW;
    int cnt = 0;
    if (have && !is_long) {
        int32_t last = -1, a; int r;
        while (next_agent_serial(g, s, last, a, r)) cnt++;
        if (cnt == 0) cnt = 1;                      // no agent uses the server: the path starts at the server itself (:737-738)
    }
    unsigned lm = __ballot_sync(FULL, is_long);
    for (unsigned rem = lm; rem; rem &= rem - 1) {
        const int src = __ffs(rem) - 1;
        const int32_t ls = __shfl_sync(FULL, s, src);
        int32_t last = -1, a; int r; int c = 0;
        while (next_agent_warp(g, ls, last, a, r, lane)) c++;
        if (c == 0) c = 1;
        if (lane == src) cnt = c;
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
    const int total = __shfl_sync(FULL, inc, 31);
    if (FILL) {
        const int64_t my0 = row0 + (inc - cnt);
        if (have && !is_long) {
            int32_t last = -1, a; int r; int k = 0;
            while (next_agent_serial(g, s, last, a, r)) write_row(A, my0 + k++, a, s, vs, f, r, rel_sv, rel_vf);
            if (k == 0) write_row(A, my0, s, s, vs, f, first_rel_rev(g, s, s), rel_sv, rel_vf);
        }
        for (unsigned rem = lm; rem; rem &= rem - 1) {
            const int src = __ffs(rem) - 1;
            const int32_t ls = __shfl_sync(FULL, s, src);
            const int lrel = __shfl_sync(FULL, rel_sv, src);
            const int64_t l0 = __shfl_sync(FULL, my0, src);
            int32_t last = -1, a; int r; int k = 0;
            while (next_agent_warp(g, ls, last, a, r, lane)) { if (lane == 0) write_row(A, l0 + k, a, ls, vs, f, r, lrel, rel_vf); k++; }
            if (k == 0 && lane == 0) write_row(A, l0, ls, ls, vs, f, first_rel_rev(g, ls, ls), lrel, rel_vf);
        }
    }
    return total;
}

// ---- pass A: links of every finding (VULNERABLE_TO originals whose source has a node record), in row order
template <bool FILL>
__global__ void __launch_bounds__(256) links_kernel(const PathsArgs A) {
    const GraphView &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t fi = warp; fi < A.io.n_findings; fi += nwarps) {
        const int32_t f = __ldg(A.io.findings + fi);
        int64_t n = 0;
        const int64_t base = FILL ? A.link_off[fi] : 0;
        bool ok = f >= 0 && f < g.n;
        if (ok) { const uint8_t ft = __ldg(g.ntype + f); ok = ft == ET_VULN || ft == ET_MISCONF; }
        if (ok) {
            const uint32_t s0 = __ldg(g.roff + f), e0 = __ldg(g.roff + f + 1);
            for (uint32_t p = s0; p < e0; p += 32) {
                const uint32_t k = p + lane;
                int32_t vs = -1; int rvf = -1;
                if (k < e0) {
                    const uint32_t m = __ldg(g.rmeta + k);
                    if (!(m & ABB_META_REVERSED_COPY) && (m & ABB_META_REL_MASK) == REL_VULNERABLE_TO) {
                        const int32_t v = __ldg(g.rnbr + k);
                        if (__ldg(g.ntype + v) != ABB_NODE_GHOST) { vs = v; rvf = (m & ABB_META_FIRST_PAIR) ? REL_VULNERABLE_TO : -3; }
                    }
                }
                const unsigned vm = __ballot_sync(FULL, vs >= 0);
                if (FILL && vs >= 0) {
                    if (rvf == -3) rvf = first_rel_rev(g, vs, f);
                    const int64_t o = base + n + __popc(vm & lanemask_lt(lane));
                    A.link_vs[o] = vs; A.link_rel[o] = static_cast<int8_t>(rvf);
                    A.need[vs] = 1;
                }
                n += __popc(vm);
            }
        }
        if (!FILL && lane == 0) A.link_cnt[fi] = n;
    }
}

// ---- pass B: the template of each referenced vulnerable source (count, then fill)
template <bool FILL>
__global__ void __launch_bounds__(256) template_kernel(const PathsArgs A) {
    const GraphView &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t nu = static_cast<int64_t>(*A.n_unique);
    for (int64_t ui = warp; ui < nu; ui += nwarps) {
        const int32_t vs = __ldg(A.ulist + ui);
        const int64_t row = FILL ? A.t_off[ui] : 0;
        int64_t n = 0;
        if (__ldg(g.ntype + vs) == ET_SERVER) {
            n += server_chunk<FILL>(A, row, lane == 0 ? vs : -1, -2, vs, 0, 0, lane);
        } else {
            const uint32_t s2 = __ldg(g.roff + vs), e2 = __ldg(g.roff + vs + 1);
            for (uint32_t p2 = s2; p2 < e2; p2 += 32) {
                const uint32_t k2 = p2 + lane;
                int32_t srv = -1; int rsv = -1;
                if (k2 < e2) {
                    const uint32_t m2 = __ldg(g.rmeta + k2);
                    if (!(m2 & ABB_META_REVERSED_COPY) && (m2 & ABB_META_REL_MASK) == REL_DEPENDS_ON) {
                        const int32_t v2 = __ldg(g.rnbr + k2);
                        if (__ldg(g.ntype + v2) == ET_SERVER) { srv = v2; rsv = (m2 & ABB_META_FIRST_PAIR) ? REL_DEPENDS_ON : -3; }
                    }
                }
                if (FILL && rsv == -3) rsv = first_rel_rev(g, srv, vs);
                if (__any_sync(FULL, srv >= 0)) n += server_chunk<FILL>(A, row + n, srv, rsv, vs, 0, 0, lane);
            }
        }
        if (lane == 0) {
            if (!FILL) { A.t_cnt[ui] = n; A.t_cnt_node[vs] = static_cast<int32_t>(n); }
            else A.t_off_node[vs] = row;
        }
    }
}

__global__ void link_rows_kernel(const PathsArgs A) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < static_cast<int64_t>(*A.n_links)) { const int32_t vs = A.link_vs[i]; A.link_rows[i] = A.t_cnt_node[vs]; A.link_toff[i] = A.t_off_node[vs]; }
}

__global__ void finding_offsets_kernel(const PathsArgs A) {
    const int64_t fi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (fi <= A.io.n_findings) A.io.f_off[fi] = A.link_roff[A.link_off[fi]];
}

// ---- pass C: every link copies its vulnerable source's template, substituting the finding
__device__ __forceinline__ void emit_row(const PathsArgs &A, int64_t out, int64_t t, int32_t vs, int32_t f, int rel_vf) {
    const int4 tr = A.t_row[t];
    const int32_t srv = tr.y;
    const int ras = A.t_rel[t * 2], rsv = A.t_rel[t * 2 + 1];
    reinterpret_cast<int4 *>(A.io.hops)[out] = make_int4(tr.x, srv, (vs != srv) ? vs : -1, f);
    // rels are stored 4 bytes per row (the 4th is padding) so a row's relationships are one aligned word
    const uint32_t packed = (vs != srv) ? ((ras & 0xFF) | ((rsv & 0xFF) << 8) | ((rel_vf & 0xFF) << 16) | (0xFEu << 24))
                                        : ((ras & 0xFF) | ((rel_vf & 0xFF) << 8) | (0xFEu << 16) | (0xFEu << 24));
    reinterpret_cast<uint32_t *>(A.io.rels)[out] = packed;
    A.io.ncred[out] = tr.z;
    A.io.ntool[out] = tr.w;
}

__global__ void __launch_bounds__(256) replicate_kernel(const PathsArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t nf = A.io.n_findings;
    // a warp takes 32 findings; short links are copied lane-serially, long ones by the whole warp
    for (int64_t fb = warp * 32; fb < nf; fb += nwarps * 32) {
        const int64_t fi = fb + lane;
        int64_t l0 = 0, l1 = 0; int32_t f = 0;
        if (fi < nf) { l0 = A.link_off[fi]; l1 = A.link_off[fi + 1]; f = __ldg(A.io.findings + fi); }
        for (int64_t l = l0; l < l1; l++) {
            const int64_t rows = A.link_roff[l + 1] - A.link_roff[l];
            if (rows > 8) continue;                 // long links below
            const int32_t vs = A.link_vs[l]; const int rel = A.link_rel[l];
            const int64_t t0 = A.t_off_node[vs], o0 = A.link_roff[l];
            for (int64_t k = 0; k < rows; k++) emit_row(A, o0 + k, t0 + k, vs, f, rel);
        }
        // long links: lane j announces them one at a time, the warp streams the template 32 rows per step
        for (int j = 0; j < 32; j++) {
            const int64_t jl0 = __shfl_sync(FULL, l0, j), jl1 = __shfl_sync(FULL, l1, j);
            const int32_t jf = __shfl_sync(FULL, f, j);
            for (int64_t l = jl0; l < jl1; l++) {
                const int64_t o0 = A.link_roff[l], rows = A.link_roff[l + 1] - o0;
                if (rows <= 8) continue;
                const int32_t vs = A.link_vs[l]; const int rel = A.link_rel[l];
                const int64_t t0 = A.t_off_node[vs];
                for (int64_t k = lane; k < rows; k += 32) emit_row(A, o0 + k, t0 + k, vs, jf, rel);
            }
        }
    }
}


// ---- ranking (reference api/routes/graph.py:762-786): sort key of every row, then the requested page is gathered.
// key (descending) = (risk rank, #hops, #distinct credential labels, #distinct tool labels); ties keep emission order
// (Python's stable sort with reverse=True) — the radix sort below is stable on the complemented key.
struct RankArgs {
    const int32_t *base_id;       // [n_findings] index of the finding's base risk value
    const uint32_t *risk_rank;    // [n_base * 5 * 15] dense rank of round(min(100, base + min(10, 3c) + min(10, .75t)), 2), c<=4, t<=14
    const int32_t *ncu;           // [n_nodes] distinct credential labels of a server
    const int32_t *ntu;           // [n_nodes] distinct tool labels of a server
    unsigned long long *keys;     // [n_rows] complemented sort key
    uint32_t *rows;               // [n_rows] row index
};

__device__ __forceinline__ unsigned long long rank_key(const PathsArgs &A, const RankArgs &R, int64_t fi, int64_t t, int32_t vs) {
    const int4 tr = A.t_row[t];
    const int nc = tr.z < 4 ? tr.z : 4, nt = tr.w < 14 ? tr.w : 14;
    const unsigned long long rr = R.risk_rank[(static_cast<int64_t>(R.base_id[fi]) * 5 + nc) * 15 + nt];
    const unsigned long long nh = (vs != tr.y) ? 4ull : 3ull;
    const unsigned long long cu = static_cast<unsigned long long>(min(R.ncu[tr.y], 0x3FFFF)), tu = static_cast<unsigned long long>(min(R.ntu[tr.y], 0x3FFFF));
    return ~((rr << 40) | (nh << 36) | (cu << 18) | tu);
}

__global__ void __launch_bounds__(256) rank_keys_kernel(const PathsArgs A, const RankArgs R) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t nf = A.io.n_findings;
    for (int64_t fb = warp * 32; fb < nf; fb += nwarps * 32) {
        const int64_t fi = fb + lane;
        int64_t l0 = 0, l1 = 0;
        if (fi < nf) { l0 = A.link_off[fi]; l1 = A.link_off[fi + 1]; }
        for (int64_t l = l0; l < l1; l++) {
            const int64_t o0 = A.link_roff[l], rows = A.link_roff[l + 1] - o0;
            if (rows > 8) continue;
            const int32_t vs = A.link_vs[l]; const int64_t t0 = A.t_off_node[vs];
            for (int64_t k = 0; k < rows; k++) { R.keys[o0 + k] = rank_key(A, R, fi, t0 + k, vs); R.rows[o0 + k] = static_cast<uint32_t>(o0 + k); }
        }
        for (int j = 0; j < 32; j++) {
            const int64_t jl0 = __shfl_sync(FULL, l0, j), jl1 = __shfl_sync(FULL, l1, j);
            const int64_t jfi = fb + j;
            for (int64_t l = jl0; l < jl1; l++) {
                const int64_t o0 = A.link_roff[l], rows = A.link_roff[l + 1] - o0;
                if (rows <= 8) continue;
                const int32_t vs = A.link_vs[l]; const int64_t t0 = A.t_off_node[vs];
                for (int64_t k = lane; k < rows; k += 32) { R.keys[o0 + k] = rank_key(A, R, jfi, t0 + k, vs); R.rows[o0 + k] = static_cast<uint32_t>(o0 + k); }
            }
        }
    }
}

// the page [first, first+count) of the sorted order -> flat rows (+ the risk rank so the host can print the float)
__global__ void rank_gather_kernel(const PathsArgs A, const RankArgs R, const uint32_t *sorted_rows, int64_t first, int64_t count, uint32_t *out_rank,
                                   int64_t *out_row) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t row = sorted_rows[first + i];
    // finding = last f with f_off[f] <= row; link = last l in the finding with link_roff[l] <= row
    int64_t lo = 0, hi = A.io.n_findings;
    while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (A.io.f_off[mid] <= row) lo = mid; else hi = mid; }
    const int64_t fi = lo;
    int64_t a = A.link_off[fi], b = A.link_off[fi + 1];
    while (b - a > 1) { const int64_t mid = (a + b) >> 1; if (A.link_roff[mid] <= row) a = mid; else b = mid; }
    const int64_t l = a;
    const int32_t vs = A.link_vs[l];
    const int64_t t = A.t_off_node[vs] + (row - A.link_roff[l]);
    emit_row(A, i, t, vs, A.io.findings[fi], A.link_rel[l]);
    const int4 tr = A.t_row[t];
    const int nc = tr.z < 4 ? tr.z : 4, nt = tr.w < 14 ? tr.w : 14;
    out_rank[i] = R.risk_rank[(static_cast<int64_t>(R.base_id[fi]) * 5 + nc) * 15 + nt];
    out_row[i] = row;
}

}  // namespace abb
