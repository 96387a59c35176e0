// union_host.inl — host orchestration of abb_group_union_host (included by abb200.cu after reach_host.inl).
// Reference use: effective_reach.py:372-426 (see union.cuh).

struct abb_union_result {       // result arrays live in pinned host blocks of the library's pool: D2H at PCIe rate, no zero-fill of fresh vectors
    HostBlock off, items, w0, w1;
    double ms = 0.0;
};
extern "C" void abb_union_result_free(abb_union_result *r) {
    if (!r) return;
    for (HostBlock *b : {&r->off, &r->items, &r->w0, &r->w1}) b->release();
    delete r;
}
extern "C" const int64_t *abb_union_result_off(const abb_union_result *r) { return r->off.as<int64_t>(); }
extern "C" const int32_t *abb_union_result_items(const abb_union_result *r) { return r->items.as<int32_t>(); }
extern "C" const uint8_t *abb_union_result_w0(const abb_union_result *r) { return r->w0.as<uint8_t>(); }
extern "C" const uint8_t *abb_union_result_w1(const abb_union_result *r) { return r->w1.as<uint8_t>(); }
extern "C" double abb_union_result_ms(const abb_union_result *r) { return r->ms; }

struct StreamGuard {
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~StreamGuard() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (s) cudaStreamDestroy(s); }
};

extern "C" int abb_group_union_host(int device, int64_t n_groups, const int64_t *member_off, const int32_t *members, int64_t n_members,
                                    const int64_t *item_off, const int32_t *items, const uint8_t *w0, const uint8_t *w1, abb_union_result **out) {
    NvtxRange nvtx_("abb_group_union_host");
    if (!out || n_groups < 0 || n_members < 0 || !member_off || !item_off) return fail(ABB_ERR_ARG, "bad arguments");
    if (n_groups >= (1ll << 31)) return fail(ABB_ERR_ARG, "at most 2^31-1 groups");
    const int64_t n_refs = member_off[n_groups], n_items = item_off[n_members];
    if (n_refs < 0 || n_items < 0 || (n_refs && !members) || (n_items && !items)) return fail(ABB_ERR_ARG, "bad offsets");
    for (int64_t k = 0; k < n_refs; k++)
        if (members[k] < 0 || members[k] >= n_members) return fail(ABB_ERR_ARG, "member index %d out of range at %lld", members[k], static_cast<long long>(k));
    DeviceGuard dg(device);
    StreamGuard sg;
    CUDA_TRY(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&sg.e0)); CUDA_TRY(cudaEventCreate(&sg.e1));
    cudaStream_t st = sg.s;

    Tmp d_moff, d_mem, d_ioff, d_items, d_w0, d_w1, counts, poff, gw0, gw1;
    if (int rc = d_moff.alloc(static_cast<size_t>(n_groups + 1) * 8)) return rc;
    if (int rc = d_mem.alloc(static_cast<size_t>(n_refs) * 4)) return rc;
    if (int rc = d_ioff.alloc(static_cast<size_t>(n_members + 1) * 8)) return rc;
    if (int rc = d_items.alloc(static_cast<size_t>(n_items) * 4)) return rc;
    if (int rc = counts.alloc(static_cast<size_t>(n_groups + 2) * 8)) return rc;
    if (int rc = poff.alloc(static_cast<size_t>(n_groups + 2) * 8)) return rc;
    if (int rc = gw0.alloc(static_cast<size_t>(n_groups + 1))) return rc;
    if (int rc = gw1.alloc(static_cast<size_t>(n_groups + 1))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_moff.p, member_off, static_cast<size_t>(n_groups + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_refs) CUDA_TRY(cudaMemcpyAsync(d_mem.p, members, static_cast<size_t>(n_refs) * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_ioff.p, item_off, static_cast<size_t>(n_members + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_items) CUDA_TRY(cudaMemcpyAsync(d_items.p, items, static_cast<size_t>(n_items) * 4, cudaMemcpyHostToDevice, st));
    if (w0) { if (int rc = d_w0.alloc(static_cast<size_t>(n_members + 1))) return rc; if (n_members) CUDA_TRY(cudaMemcpyAsync(d_w0.p, w0, static_cast<size_t>(n_members), cudaMemcpyHostToDevice, st)); }
    if (w1) { if (int rc = d_w1.alloc(static_cast<size_t>(n_members + 1))) return rc; if (n_members) CUDA_TRY(cudaMemcpyAsync(d_w1.p, w1, static_cast<size_t>(n_members), cudaMemcpyHostToDevice, st)); }
    CUDA_TRY(cudaMemsetAsync(counts.p, 0, static_cast<size_t>(n_groups + 2) * 8, st));
    CUDA_TRY(cudaEventRecord(sg.e0, st));
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(std::max<int64_t>(1, (n_groups + 7) / 8), 148ll * 32));
    if (n_groups) {
        union_count_kernel<<<grid, 256, 0, st>>>(n_groups, d_moff.as<int64_t>(), d_mem.as<int32_t>(), d_ioff.as<int64_t>(), w0 ? d_w0.as<uint8_t>() : nullptr,
                                                 w1 ? d_w1.as<uint8_t>() : nullptr, counts.as<int64_t>(), gw0.as<uint8_t>(), gw1.as<uint8_t>());
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if (int rc = exclusive_scan_i64(st, counts.as<int64_t>(), poff.as<int64_t>(), n_groups + 1)) return rc;
    int64_t A0 = 0;
    CUDA_TRY(cudaMemcpy(&A0, poff.as<int64_t>() + n_groups, 8, cudaMemcpyDeviceToHost));
    Tmp k1, v1, k2, v2;
    if (int rc = k1.alloc(static_cast<size_t>(A0) * 8)) return rc;
    if (int rc = v1.alloc(static_cast<size_t>(A0) * 4)) return rc;
    if (int rc = k2.alloc(static_cast<size_t>(A0) * 8)) return rc;
    if (int rc = v2.alloc(static_cast<size_t>(A0) * 4)) return rc;
    if (A0) {
        union_fill_kernel<<<grid, 256, 0, st>>>(n_groups, d_moff.as<int64_t>(), d_mem.as<int32_t>(), d_ioff.as<int64_t>(), d_items.as<int32_t>(),
                                                poff.as<int64_t>(), k1.as<unsigned long long>(), v1.as<int32_t>());
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    int64_t AU = 0;
    int group_bits = 1;
    while ((1ll << group_bits) < n_groups) group_bits++;          // only the key bits that vary take a radix pass
    if (int rc = sort_unique_pairs(st, k1.as<unsigned long long>(), v1.as<int32_t>(), k2.as<unsigned long long>(), v2.as<int32_t>(), A0, &AU, true, 32 + group_bits)) return rc;
    CUDA_TRY(cudaMemsetAsync(counts.p, 0, static_cast<size_t>(n_groups + 2) * 8, st));
    if (AU) { reach_group_counts_kernel<<<nblk(AU, 256), 256, 0, st>>>(AU, k1.as<unsigned long long>(), counts.as<unsigned long long>()); g_launches++; }
    if (int rc = exclusive_scan_i64(st, counts.as<int64_t>(), poff.as<int64_t>(), n_groups + 1)) return rc;
    CUDA_TRY(cudaEventRecord(sg.e1, st));
    CUDA_TRY(cudaStreamSynchronize(st));

    abb_union_result *r = new abb_union_result();
    float ms = 0.f; cudaEventElapsedTime(&ms, sg.e0, sg.e1); r->ms = ms;
    if (!(r->off.alloc(static_cast<size_t>(n_groups + 1) * 8) && r->items.alloc(static_cast<size_t>(AU) * 4) && r->w0.alloc(static_cast<size_t>(n_groups)) &&
          r->w1.alloc(static_cast<size_t>(n_groups)))) { abb_union_result_free(r); return fail(ABB_ERR_NOMEM, "pinned host allocation failed"); }
    cudaError_t e = cudaMemcpyAsync(r->off.p, poff.p, static_cast<size_t>(n_groups + 1) * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && AU) e = cudaMemcpyAsync(r->items.p, v1.p, static_cast<size_t>(AU) * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n_groups) e = cudaMemcpyAsync(r->w0.p, gw0.p, static_cast<size_t>(n_groups), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n_groups) e = cudaMemcpyAsync(r->w1.p, gw1.p, static_cast<size_t>(n_groups), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { abb_union_result_free(r); return fail(ABB_ERR_CUDA, "union D2H: %s", cudaGetErrorString(e)); }
    *out = r;
    return ABB_OK;
}
