// lateral.cuh — capped simple-path search behind find_lateral_paths (sm_100a).
//
// Reference: agent_bom/context_graph.py:397-477.  From a source node a FIFO of *paths* is expanded over the mirrored
// adjacency; a path whose end is a "lateral target" (an agent other than the source's, or a credential / tool owned by
// another agent) is recorded and not extended — unless the same node sequence was recorded already (it can arrive
// twice through parallel edges of different kinds), in which case it IS extended.  The search stops at 100 recorded
// paths or an empty queue; a popped path is not extended while 10 000 or more paths wait in the queue, and paths longer
// than max_depth+1 nodes are dropped.  All of that is order-sensitive, so one warp replays the reference's queue
// exactly for one source: the pop is serial, the expansion of the popped node's adjacency row is lane-parallel with a
// ballot / prefix append that keeps adjacency order.  Sources (the CLI and the REST route call this for every agent)
// are the data-parallel axis: one warp each, a private ring of path records in global memory.
//
// Record layout (int32 words): [0] length, [1] edge kinds (4 bits per hop), [2 .. 2+W) node ids; W = max_depth + 2.
#pragma once
#include <cstdint>

namespace abb {

constexpr int LAT_MAX_PATHS = 100;      // _MAX_PATHS
constexpr int LAT_MAX_QUEUE = 10000;    // _MAX_QUEUE_SIZE
constexpr int LAT_MAX_W = 9;            // max_depth <= 7: 8 hops x 4 bits of edge kind fit one word
constexpr uint8_t LAT_KIND_AGENT = 0, LAT_KIND_CRED = 2, LAT_KIND_TOOL = 3, LAT_KIND_GHOST = 255;

struct LateralArgs {
    int32_t n;                   // nodes
    const int64_t *off;          // [n+1] adjacency rows (graph.adjacency[id], list order)
    const int32_t *nbr;          // entry target
    const uint8_t *ekind;        // entry edge kind (0..15)
    const uint8_t *nkind;        // node kind; 255 = id without a node record
    const int32_t *nkey;         // agents: id of the label; credentials / tools: id of metadata["agent"]; -1 = absent or empty
    int64_t n_sources;
    const int32_t *src;          // source node per search (-1: not in graph.nodes -> no paths)
    const int32_t *src_key;      // id of the source's agent name ("" has an id of its own)
    int32_t max_depth;
    int32_t W;                   // node slots per record
    int32_t ring_cap;            // records per ring (> LAT_MAX_QUEUE + longest row)
    int32_t *ring;               // [n_warps][ring_cap][W+2]
    long long max_pops;          // safety valve per source; exceeding it sets the overflow flag instead of running on
    int32_t *out_count;          // [n_sources]
    int32_t *out_paths;          // [n_sources][LAT_MAX_PATHS][W+2]
    int32_t *out_flags;          // [n_sources] bit0 = safety valve hit
};

__global__ void __launch_bounds__(128) lateral_search_kernel(LateralArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int R = A.W + 2;
    int32_t *ring = A.ring + warp * static_cast<int64_t>(A.ring_cap) * R;
    for (int64_t q = warp; q < A.n_sources; q += nwarps) {
        int32_t *found = A.out_paths + q * static_cast<int64_t>(LAT_MAX_PATHS) * R;
        const int32_t s = A.src[q];
        int npaths = 0, flags = 0;
        if (s >= 0 && s < A.n && A.nkind[s] != LAT_KIND_GHOST) {
            const int32_t skey = A.src_key[q];
            int head = 0, count = 1;
            if (lane == 0) { ring[0] = 1; ring[1] = 0; ring[2] = s; }
            __syncwarp();
            long long pops = 0;
            // every quantity that steers control flow below is read from the same record by all lanes: the warp never diverges
            while (count > 0 && npaths < LAT_MAX_PATHS) {
                if (++pops > A.max_pops) { flags |= 1; break; }
                const int32_t *e = ring + static_cast<int64_t>(head) * R;
                const int len = e[0];
                const uint32_t kinds = static_cast<uint32_t>(e[1]);
                int32_t pn[LAT_MAX_W];                               // the popped path, whole, in every lane (broadcast loads)
#pragma unroll
                for (int k = 0; k < LAT_MAX_W; k++) pn[k] = k < len ? e[2 + k] : -1;
                head = head + 1 == A.ring_cap ? 0 : head + 1;
                count--;
                if (len > A.max_depth + 1) continue;
                const int32_t cur = e[2 + len - 1];
                if (len > 1) {
                    const uint8_t ck = A.nkind[cur];
                    const int32_t key = A.nkey[cur];
                    bool lateral = false;
                    if (ck == LAT_KIND_AGENT) lateral = key != skey;
                    else if (ck == LAT_KIND_CRED || ck == LAT_KIND_TOOL) lateral = key >= 0 && key != skey;
                    if (lateral) {
                        bool dup = false;                               // same node sequence recorded before?
                        for (int base = 0; base < npaths; base += 32) {
                            const int i = base + lane;
                            bool same = false;
                            if (i < npaths) {
                                const int32_t *f = found + static_cast<int64_t>(i) * R;
                                same = f[0] == len;
#pragma unroll
                                for (int k = 0; k < LAT_MAX_W; k++) if (k < len) same = same && f[2 + k] == pn[k];
                            }
                            dup = dup || __any_sync(0xFFFFFFFFu, same);
                        }
                        if (!dup) {
                            int32_t *f = found + static_cast<int64_t>(npaths) * R;
                            if (lane == 0) {
                                f[0] = len; f[1] = static_cast<int32_t>(kinds);
#pragma unroll
                                for (int k = 0; k < LAT_MAX_W; k++) if (k < A.W) f[2 + k] = pn[k];
                            }
                            npaths++;
                            __syncwarp();
                            continue;                                   // a recorded target is not extended
                        }
                    }
                }
                if (count >= LAT_MAX_QUEUE) continue;                   // len(queue) after the pop
                const int64_t a = A.off[cur], b = A.off[cur + 1];
                for (int64_t p0 = a; p0 < b; p0 += 32) {
                    const int64_t p = p0 + lane;
                    int32_t nb = -1; uint32_t ek = 0;
                    bool ok = p < b;
                    if (ok) { nb = A.nbr[p]; ek = A.ekind[p]; }
#pragma unroll
                    for (int k = 0; k < LAT_MAX_W; k++) if (k < len) ok = ok && nb != pn[k];
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, ok);
                    if (ok) {
                        int slot = head + count + __popc(m & ((1u << lane) - 1u));
                        if (slot >= A.ring_cap) slot -= A.ring_cap;
                        int32_t *w = ring + static_cast<int64_t>(slot) * R;
                        w[0] = len + 1;
                        w[1] = static_cast<int32_t>(kinds | (ek << (4 * (len - 1))));
#pragma unroll
                        for (int k = 0; k < LAT_MAX_W; k++) if (k < len) w[2 + k] = pn[k];
                        w[2 + len] = nb;
                    }
                    count += __popc(m);
                }
                __syncwarp();
            }
        }
        if (lane == 0) { A.out_count[q] = npaths; A.out_flags[q] = flags; }
        __syncwarp();
    }
}

}  // namespace abb
