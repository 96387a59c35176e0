// Column-compacted entity-type histograms for the host result (abb_walk_host / abb_exposure_host).
//
// The reference's impact_of returns `affected_by_type`, a dict with an entry only for the entity types a blast radius
// contains (graph/container.py:268-276).  The device keeps ABB_N_ENTITY_TYPES = 24 counters per query; on a real estate a
// handful of those columns are ever non-zero (agents, servers, tools, credentials, packages ...), and counts fit 16 bits
// unless a single traversal reaches 65 536 nodes of one type.  What crosses PCIe is therefore the non-zero columns only,
// at the narrowest width that holds the largest count; the dense [n_queries x 24] uint32 table of the ABI is rebuilt on the
// host on first access (abb_walk_result_hist), and abb_walk_result_hist_packed hands out the packed form as is.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/abb200.h"

namespace abb {

struct HistCols { uint8_t col[ABB_N_ENTITY_TYPES]; int k; };

// info[0] |= bit t for every column t with a non-zero count; info[1] = max count
__global__ void hist_columns_kernel(const uint32_t *__restrict__ hist, int64_t n_elems, unsigned long long *info) {
    unsigned mask = 0, mx = 0;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
        const uint32_t v = __ldg(hist + e);
        if (v) { mask |= 1u << static_cast<unsigned>(e % ABB_N_ENTITY_TYPES); mx = max(mx, v); }
    }
    mask = __reduce_or_sync(0xFFFFFFFFu, mask);
    mx = __reduce_max_sync(0xFFFFFFFFu, mx);
    if ((threadIdx.x & 31) == 0 && mask) { atomicOr(info, static_cast<unsigned long long>(mask)); atomicMax(info + 1, static_cast<unsigned long long>(mx)); }
}

// out[q * k + j] = hist[q * 24 + col[j]]
template <class T>
__global__ void hist_pack_kernel(const uint32_t *__restrict__ hist, int64_t nq, HistCols hc, T *__restrict__ out) {
    const int64_t n_out = nq * hc.k;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        const int64_t q = i / hc.k;
        const int j = static_cast<int>(i - q * hc.k);
        out[i] = static_cast<T>(__ldg(hist + q * ABB_N_ENTITY_TYPES + hc.col[j]));
    }
}

// ---- chunked host walks (abb200.cu: walk_host_chunked): sources are partitioned into chunks by frontier signature (all members of a
// frontier group land in one chunk), walked chunk by chunk in that order, and the per-query results are put back in caller order.

// key[q] = chunk of query q; counts[c] += 1
__global__ void chunk_key_kernel(const unsigned long long *__restrict__ sig, int64_t nq, int n_chunks, uint32_t *key, int32_t *idx, unsigned long long *counts) {
    __shared__ unsigned int s_cnt[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q < nq) {
        const uint32_t c = sig ? static_cast<uint32_t>(sig[q] % static_cast<unsigned long long>(n_chunks)) : static_cast<uint32_t>(q * n_chunks / nq);
        key[q] = c;
        idx[q] = static_cast<int32_t>(q);
        atomicAdd(&s_cnt[c], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_chunks && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, static_cast<unsigned long long>(s_cnt[threadIdx.x]));
}

__global__ void gather_i32_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ idx, int32_t *__restrict__ dst, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __ldg(src + __ldg(idx + i));
}

// dst[idx[i] * width + j] = src[i * width + j]
template <class T>
__global__ void scatter_rows_kernel(const T *__restrict__ src, const int32_t *__restrict__ idx, T *__restrict__ dst, int64_t n, int width) {
    const int64_t n_elems = n * width;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
        const int64_t i = e / width;
        const int j = static_cast<int>(e - i * width);
        dst[static_cast<int64_t>(__ldg(idx + i)) * width + j] = src[e];
    }
}

// out[idx[q] * k + j] = hist[q * 24 + col[j]]
template <class T>
__global__ void hist_pack_scatter_kernel(const uint32_t *__restrict__ hist, const int32_t *__restrict__ idx, int64_t nq, HistCols hc, T *__restrict__ out) {
    const int64_t n_out = nq * hc.k;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        const int64_t q = i / hc.k;
        const int j = static_cast<int>(i - q * hc.k);
        out[static_cast<int64_t>(__ldg(idx + q)) * hc.k + j] = static_cast<T>(__ldg(hist + q * ABB_N_ENTITY_TYPES + hc.col[j]));
    }
}

}  // namespace abb
