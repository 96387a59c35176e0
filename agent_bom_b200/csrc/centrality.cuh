// centrality.cuh — subtree roll-up for the sampled "bottleneck" score (sm_100a).
//
// Reference: UnifiedGraph.bottleneck_nodes (graph/container.py:548-567) and InMemoryBackend.bottleneck_nodes
// (graph_backend.py:127-155).  From each sampled source an unbounded BFS keeps, per reached node, the path through
// its first discoverer; every node strictly inside such a path gets +1.  In tree terms: every reached node other
// than the source gets the number of its proper descendants in that BFS tree.
//
// The ordered walk already emits the tree: slice entries in queue order (level by level), `parent` = queue position
// of the first discoverer (0 = the source, k > 0 = emitted entry k-1), `depth` non-decreasing along the slice.
// One block per source rolls descendants up from the deepest level to level 2 (entries of one level never depend on
// each other), then adds every entry's count into the per-node 64-bit score.
#pragma once
#include <cstdint>

namespace abb {

__global__ void subtree_rollup_kernel(int64_t nq, const int64_t *q_start, const int32_t *q_count, const int32_t *q_maxd, const int32_t *nodes,
                                      const int32_t *parent, const int32_t *depth, int32_t *desc, unsigned long long *scores) {
    __shared__ int s_lo, s_hi;
    for (int64_t q = blockIdx.x; q < nq; q += gridDim.x) {
        const int64_t base = q_start[q];
        const int n = q_count[q];
        const int32_t *dep = depth + base;
        const int32_t *par = parent + base;
        int32_t *d = desc + base;
        for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = 0;
        if (threadIdx.x == 0) s_hi = n;
        __syncthreads();
        for (int level = q_maxd[q]; level >= 2; level--) {
            if (threadIdx.x == 0) {                       // first entry at this depth (depths are non-decreasing)
                int lo = 0, hi = s_hi;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (dep[mid] < level) lo = mid + 1; else hi = mid; }
                s_lo = lo;
            }
            __syncthreads();
            const int lo = s_lo, hi = s_hi;
            for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&d[par[i] - 1], d[i] + 1);
            __syncthreads();
            if (threadIdx.x == 0) s_hi = lo;
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int c = d[i];
            if (c) atomicAdd(&scores[nodes[base + i]], static_cast<unsigned long long>(c));
        }
        __syncthreads();
    }
}

}  // namespace abb
