// union.cuh — per-group union of member item lists + per-group maxima (sm_100a).
//
// Effective-reach scoring (reference effective_reach.py:372-426) reduces, for every
// vulnerability, over the servers that are VULNERABLE_TO it: the strongest tool
// capability, the most visible credential tier, and three de-duplicated, sorted
// label lists (tools, credentials, agents — the agent list's length is the
// "breadth").  In array form that is: groups (vulnerabilities) -> members
// (servers) -> items (label ranks, category in the top bits), plus two byte
// weights per member.  One expansion into 64-bit (group << 32 | item) keys, one
// radix sort, one unique pass and a histogram give every group's sorted set;
// the maxima ride along in the expansion kernel.  HBM-bound integer work: 12 bytes
// written and 12 read per (group, item) pair around the sort.
#pragma once
#include <cstdint>

namespace abb {

// warp per group: pairs this group expands to, and the two maxima over its members
__global__ void union_count_kernel(int64_t n_groups, const int64_t *member_off, const int32_t *members, const int64_t *item_off,
                                   const uint8_t *w0, const uint8_t *w1, int64_t *counts, uint8_t *g_w0, uint8_t *g_w1) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t g = warp; g < n_groups; g += nwarps) {
        const int64_t a = member_off[g], b = member_off[g + 1];
        long long total = 0; unsigned m0 = 0, m1 = 0;
        for (int64_t k = a + lane; k < b; k += 32) {
            const int32_t m = members[k];
            total += item_off[m + 1] - item_off[m];
            if (w0) m0 = max(m0, static_cast<unsigned>(w0[m]));
            if (w1) m1 = max(m1, static_cast<unsigned>(w1[m]));
        }
        for (int s = 16; s; s >>= 1) {
            total += __shfl_xor_sync(0xFFFFFFFFu, total, s);
            m0 = max(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, s));
            m1 = max(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, s));
        }
        if (lane == 0) { counts[g] = total; g_w0[g] = static_cast<uint8_t>(m0); g_w1[g] = static_cast<uint8_t>(m1); }
    }
}

// warp per group: write (group << 32 | item) for every item of every member
__global__ void union_fill_kernel(int64_t n_groups, const int64_t *member_off, const int32_t *members, const int64_t *item_off, const int32_t *items,
                                  const int64_t *pair_off, unsigned long long *keys, int32_t *vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t g = warp; g < n_groups; g += nwarps) {
        const unsigned long long hi = static_cast<unsigned long long>(static_cast<uint32_t>(g)) << 32;
        int64_t w = pair_off[g];
        for (int64_t k = member_off[g]; k < member_off[g + 1]; k++) {
            const int32_t m = members[k];
            const int64_t i0 = item_off[m], n = item_off[m + 1] - i0;
            for (int64_t i = lane; i < n; i += 32) {
                const int32_t it = items[i0 + i];
                keys[w + i] = hi | static_cast<uint32_t>(it);
                vals[w + i] = it;
            }
            w += n;
        }
    }
}

}  // namespace abb
