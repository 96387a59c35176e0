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
    int64_t *counts;            // [n_findings] per-finding row counts (pass 1 output)
    const int32_t *srv_cred;    // [n_nodes] EXPOSES_CRED fan-out of server nodes (graph-constant table)
    const int32_t *srv_tool;    // [n_nodes] PROVIDES_TOOL fan-out
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

__device__ __forceinline__ void write_row(const PathsArgs &A, int64_t row, int32_t agent, int32_t srv, int32_t vs, int32_t f, int rel_as, int rel_sv, int rel_vf) {
    int32_t *h = A.io.hops + row * 4;
    h[0] = agent; h[1] = srv; h[2] = (vs != srv) ? vs : -1; h[3] = f;
    int8_t *r = A.io.rels + row * 3;
    if (vs != srv) { r[0] = static_cast<int8_t>(rel_as); r[1] = static_cast<int8_t>(rel_sv); r[2] = static_cast<int8_t>(rel_vf); }
    else { r[0] = static_cast<int8_t>(rel_as); r[1] = static_cast<int8_t>(rel_vf); r[2] = -2; }
    A.io.ncred[row] = __ldg(A.srv_cred + srv);
    A.io.ntool[row] = __ldg(A.srv_tool + srv);
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

template <bool FILL>
__global__ void __launch_bounds__(256) paths_kernel(const PathsArgs A) {
    const GraphView &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t fi = warp; fi < A.io.n_findings; fi += nwarps) {
        const int32_t f = __ldg(A.io.findings + fi);
        int64_t row = FILL ? A.io.f_off[fi] : 0, n = 0;
        bool ok = f >= 0 && f < g.n;
        if (ok) { const uint8_t ft = __ldg(g.ntype + f); ok = ft == ET_VULN || ft == ET_MISCONF; }
        if (ok) {
            const uint32_t s0 = __ldg(g.roff + f), e0 = __ldg(g.roff + f + 1);
            for (uint32_t p = s0; p < e0; p += 32) {
                const uint32_t k = p + lane;
                int32_t vs = -1; int vt = 0, rvf = -1;
                if (k < e0) {
                    const uint32_t m = __ldg(g.rmeta + k);
                    if (!(m & ABB_META_REVERSED_COPY) && (m & ABB_META_REL_MASK) == REL_VULNERABLE_TO) {
                        const int32_t v = __ldg(g.rnbr + k);
                        vt = __ldg(g.ntype + v);
                        if (vt != ABB_NODE_GHOST) { vs = v; rvf = (m & ABB_META_FIRST_PAIR) ? REL_VULNERABLE_TO : -3; }
                    }
                }
                unsigned vm = __ballot_sync(FULL, vs >= 0);
                while (vm) {  // vulnerable sources in row order
                    const int src = __ffs(vm) - 1; vm &= vm - 1;
                    const int32_t cvs = __shfl_sync(FULL, vs, src);
                    const int cvt = __shfl_sync(FULL, vt, src);
                    int rel_vf = __shfl_sync(FULL, rvf, src);
                    if (FILL && rel_vf == -3) rel_vf = first_rel_rev(g, cvs, f);
                    if (cvt == ET_SERVER) {
                        n += server_chunk<FILL>(A, row + n, lane == 0 ? cvs : -1, -2, cvs, f, rel_vf, lane);
                    } else {
                        const uint32_t s2 = __ldg(g.roff + cvs), e2 = __ldg(g.roff + cvs + 1);
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
                            if (FILL && rsv == -3) rsv = first_rel_rev(g, srv, cvs);
                            if (__any_sync(FULL, srv >= 0)) n += server_chunk<FILL>(A, row + n, srv, rsv, cvs, f, rel_vf, lane);
                        }
                    }
                }
            }
        }
        if (!FILL && lane == 0) A.counts[fi] = n;
    }
}

}  // namespace abb
