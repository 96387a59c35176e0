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
// One warp per finding; two passes over the same code (count, then fill at the
// scanned offsets) so rows land in the reference's emission order.  "Originals
// only" = entries without ABB_META_REVERSED_COPY: the reference builds its
// incoming/outgoing maps from graph.edges, which never holds the reversed
// copies of bidirectional edges (:698-702).
#pragma once
#include "walk.cuh"

namespace abb {

constexpr int REL_USES = 1, REL_DEPENDS_ON = 2, REL_PROVIDES_TOOL = 3, REL_EXPOSES_CRED = 4, REL_VULNERABLE_TO = 9;
constexpr int ET_AGENT = 0, ET_SERVER = 1, ET_VULN = 8, ET_MISCONF = 9, ET_USER = 13, ET_SERVICE_ACCOUNT = 17;

struct PathsArgs {
    GraphView g;
    abb_paths_io io;
    int64_t *counts;  // [n_findings] per-finding row counts (pass 1 output)
};

// first relationship recorded for the pair (a,b) == first entry of forward row a with neighbour b
__device__ __forceinline__ int first_rel(const GraphView &g, int32_t a, int32_t b, int lane) {
    uint32_t s = __ldg(g.foff + a), e = __ldg(g.foff + a + 1);
    for (uint32_t p = s; p < e; p += 32) {
        uint32_t k = p + lane;
        bool hit = k < e && __ldg(g.fnbr + k) == b;
        unsigned hm = __ballot_sync(FULL, hit);
        if (hm) {
            int src = __ffs(hm) - 1;
            int rel = hit ? (__ldg(g.fmeta + k) & ABB_META_REL_MASK) : 0;
            return __shfl_sync(FULL, rel, src);
        }
    }
    return -1;
}

template <bool FILL>
__device__ __forceinline__ void emit_row(const PathsArgs &A, int64_t row, int32_t agent, int32_t srv, int32_t vs, int32_t f,
                                         int ncred, int ntool, int lane) {
    if (!FILL) return;
    const GraphView &g = A.g;
    int32_t hp[4]; int nh = 0;
    hp[nh++] = agent; hp[nh++] = srv; if (vs != srv) hp[nh++] = vs; hp[nh++] = f;
    int rels[3];
#pragma unroll
    for (int k = 0; k < 3; k++) rels[k] = (k + 1 < nh) ? first_rel(g, hp[k], hp[k + 1], lane) : -2;
    if (lane == 0) {
        A.io.hops[row * 4 + 0] = agent; A.io.hops[row * 4 + 1] = srv; A.io.hops[row * 4 + 2] = (vs != srv) ? vs : -1; A.io.hops[row * 4 + 3] = f;
        A.io.rels[row * 3 + 0] = static_cast<int8_t>(rels[0]); A.io.rels[row * 3 + 1] = static_cast<int8_t>(rels[1]); A.io.rels[row * 3 + 2] = static_cast<int8_t>(rels[2]);
        A.io.ncred[row] = ncred; A.io.ntool[row] = ntool;
    }
}

// rows for one (finding, vulnerable_source, server); returns number of rows
template <bool FILL>
__device__ int64_t rows_for_server(const PathsArgs &A, int64_t row0, int32_t srv, int32_t vs, int32_t f, int lane) {
    const GraphView &g = A.g;
    int ncred = 0, ntool = 0;
    if (FILL) {
        uint32_t s = __ldg(g.foff + srv), e = __ldg(g.foff + srv + 1);
        for (uint32_t p = s; p < e; p += 32) {
            uint32_t k = p + lane;
            int rel = -1;
            if (k < e) {
                uint32_t m = __ldg(g.fmeta + k);
                if (!(m & ABB_META_REVERSED_COPY) && __ldg(g.ntype + __ldg(g.fnbr + k)) != ABB_NODE_GHOST) rel = m & ABB_META_REL_MASK;
            }
            ncred += __popc(__ballot_sync(FULL, rel == REL_EXPOSES_CRED));
            ntool += __popc(__ballot_sync(FULL, rel == REL_PROVIDES_TOOL));
        }
    }
    // agents in ascending id-string rank, duplicates collapsed: repeatedly take the smallest rank above the last one
    uint32_t s = __ldg(g.roff + srv), e = __ldg(g.roff + srv + 1);
    int64_t n = 0;
    int32_t last = -1;
    for (;;) {
        int32_t best = 0x7FFFFFFF, best_node = -1;
        for (uint32_t p = s; p < e; p += 32) {
            uint32_t k = p + lane;
            if (k < e) {
                uint32_t m = __ldg(g.rmeta + k);
                if (!(m & ABB_META_REVERSED_COPY) && (m & ABB_META_REL_MASK) == REL_USES) {
                    int32_t a = __ldg(g.rnbr + k);
                    uint8_t t = __ldg(g.ntype + a);
                    if (t == ET_AGENT || t == ET_USER || t == ET_SERVICE_ACCOUNT) {
                        int32_t r = __ldg(g.rank + a);
                        if (r > last && r < best) { best = r; best_node = a; }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            int32_t ob = __shfl_xor_sync(FULL, best, o), on = __shfl_xor_sync(FULL, best_node, o);
            if (ob < best) { best = ob; best_node = on; }
        }
        if (best_node < 0) break;
        emit_row<FILL>(A, row0 + n, best_node, srv, vs, f, ncred, ntool, lane);
        n++; last = best;
    }
    if (n == 0) {  // no agent uses the server: the path starts at the server itself (:737-738)
        emit_row<FILL>(A, row0, srv, srv, vs, f, ncred, ntool, lane);
        n = 1;
    }
    return n;
}

template <bool FILL>
__global__ void __launch_bounds__(256) paths_kernel(const PathsArgs A) {
    const GraphView &g = A.g;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t fi = warp; fi < A.io.n_findings; fi += nwarps) {
        int32_t f = __ldg(A.io.findings + fi);
        int64_t row = FILL ? A.io.f_off[fi] : 0, n = 0;
        bool ok = f >= 0 && f < g.n;
        if (ok) { uint8_t ft = __ldg(g.ntype + f); ok = ft == ET_VULN || ft == ET_MISCONF; }
        if (ok) {
            uint32_t s = __ldg(g.roff + f), e = __ldg(g.roff + f + 1);
            for (uint32_t p = s; p < e; p += 32) {
                uint32_t k = p + lane;
                int32_t vs = -1; uint8_t vt = 0;
                if (k < e) {
                    uint32_t m = __ldg(g.rmeta + k);
                    if (!(m & ABB_META_REVERSED_COPY) && (m & ABB_META_REL_MASK) == REL_VULNERABLE_TO) {
                        int32_t v = __ldg(g.rnbr + k);
                        vt = __ldg(g.ntype + v);
                        if (vt != ABB_NODE_GHOST) vs = v;
                    }
                }
                unsigned vm = __ballot_sync(FULL, vs >= 0);
                while (vm) {  // matches in row order
                    int src = __ffs(vm) - 1; vm &= vm - 1;
                    int32_t cvs = __shfl_sync(FULL, vs, src);
                    int cvt = __shfl_sync(FULL, static_cast<int>(vt), src);
                    if (cvt == ET_SERVER) {
                        n += rows_for_server<FILL>(A, row + n, cvs, cvs, f, lane);
                    } else {
                        uint32_t s2 = __ldg(g.roff + cvs), e2 = __ldg(g.roff + cvs + 1);
                        for (uint32_t p2 = s2; p2 < e2; p2 += 32) {
                            uint32_t k2 = p2 + lane;
                            int32_t sp = -1;
                            if (k2 < e2) {
                                uint32_t m2 = __ldg(g.rmeta + k2);
                                if (!(m2 & ABB_META_REVERSED_COPY) && (m2 & ABB_META_REL_MASK) == REL_DEPENDS_ON) {
                                    int32_t v2 = __ldg(g.rnbr + k2);
                                    if (__ldg(g.ntype + v2) == ET_SERVER) sp = v2;
                                }
                            }
                            unsigned sm = __ballot_sync(FULL, sp >= 0);
                            while (sm) {
                                int s3 = __ffs(sm) - 1; sm &= sm - 1;
                                int32_t srv = __shfl_sync(FULL, sp, s3);
                                n += rows_for_server<FILL>(A, row + n, srv, cvs, f, lane);
                            }
                        }
                    }
                }
            }
        }
        if (!FILL && lane == 0) A.counts[fi] = n;
    }
}

}  // namespace abb
