// lateral_host.inl — abb_lateral_paths_host (included by abb200.cu).  Reference: context_graph.py:397-477.

struct abb_lateral_result {
    std::vector<int64_t> off;        // [n_sources+1] first record of each source
    std::vector<int32_t> records;    // [off[n]][W+2]: length, edge kinds (4 bits per hop), node ids (-1 padded)
    std::vector<int32_t> flags;      // [n_sources]
    int32_t W = 0;
    double ms = 0.0;
};
extern "C" void abb_lateral_result_free(abb_lateral_result *r) { delete r; }
extern "C" const int64_t *abb_lateral_result_off(const abb_lateral_result *r) { return r->off.data(); }
extern "C" const int32_t *abb_lateral_result_records(const abb_lateral_result *r) { return r->records.data(); }
extern "C" const int32_t *abb_lateral_result_flags(const abb_lateral_result *r) { return r->flags.data(); }
extern "C" int32_t abb_lateral_result_width(const abb_lateral_result *r) { return r->W; }
extern "C" double abb_lateral_result_ms(const abb_lateral_result *r) { return r->ms; }

extern "C" int abb_lateral_paths_host(int device, int32_t n_nodes, const int64_t *adj_off, const int32_t *adj_nbr, const uint8_t *adj_kind,
                                      const uint8_t *node_kind, const int32_t *node_key, int64_t n_sources, const int32_t *sources,
                                      const int32_t *source_key, int32_t max_depth, int64_t max_pops, abb_lateral_result **out) {
    NvtxRange nvtx_("abb_lateral_paths_host");
    if (!out || n_nodes < 0 || n_sources < 0 || !adj_off || (n_nodes && (!node_kind || !node_key)) || (n_sources && (!sources || !source_key)))
        return fail(ABB_ERR_ARG, "bad arguments");
    if (max_depth < 0 || max_depth + 2 > LAT_MAX_W) return fail(ABB_ERR_ARG, "max_depth must be 0..%d", LAT_MAX_W - 2);
    const int64_t n_entries = adj_off[n_nodes];
    if (n_entries < 0 || (n_entries && (!adj_nbr || !adj_kind))) return fail(ABB_ERR_ARG, "bad adjacency");
    int64_t max_row = 0;
    for (int32_t u = 0; u < n_nodes; u++) {
        const int64_t d = adj_off[u + 1] - adj_off[u];
        if (d < 0) return fail(ABB_ERR_ARG, "adjacency offsets must not decrease");
        max_row = std::max(max_row, d);
    }
    for (int64_t p = 0; p < n_entries; p++) {
        if (adj_nbr[p] < 0 || adj_nbr[p] >= n_nodes) return fail(ABB_ERR_ARG, "adjacency target out of range at %lld", static_cast<long long>(p));
        if (adj_kind[p] > 15) return fail(ABB_ERR_ARG, "edge kinds must fit 4 bits");
    }
    DeviceGuard dg(device);
    StreamGuard sg;
    CUDA_TRY(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&sg.e0)); CUDA_TRY(cudaEventCreate(&sg.e1));
    cudaStream_t st = sg.s;
    const int W = max_depth + 2, R = W + 2;
    const int64_t ring_cap = LAT_MAX_QUEUE + max_row + 32;
    if (ring_cap >= (1ll << 30)) return fail(ABB_ERR_ARG, "adjacency row too long for the path ring");
    const int64_t ring_bytes = ring_cap * R * 4;
    int64_t warps = std::min<int64_t>(std::max<int64_t>(n_sources, 1), 148ll * 16);
    warps = std::max<int64_t>(1, std::min<int64_t>(warps, (6ll << 30) / ring_bytes));       // at most 6 GiB of rings
    const int64_t blocks = (warps + 3) / 4;
    warps = blocks * 4;

    Tmp d_off, d_nbr, d_ek, d_nk, d_key, d_src, d_skey, d_ring, d_cnt, d_paths, d_flags;
    if (int rc = d_off.alloc(static_cast<size_t>(n_nodes + 1) * 8)) return rc;
    if (int rc = d_nbr.alloc(static_cast<size_t>(n_entries) * 4)) return rc;
    if (int rc = d_ek.alloc(static_cast<size_t>(n_entries))) return rc;
    if (int rc = d_nk.alloc(static_cast<size_t>(n_nodes))) return rc;
    if (int rc = d_key.alloc(static_cast<size_t>(n_nodes) * 4)) return rc;
    if (int rc = d_src.alloc(static_cast<size_t>(n_sources) * 4)) return rc;
    if (int rc = d_skey.alloc(static_cast<size_t>(n_sources) * 4)) return rc;
    if (int rc = d_ring.alloc(static_cast<size_t>(warps * ring_bytes))) return rc;
    if (int rc = d_cnt.alloc(static_cast<size_t>(n_sources) * 4)) return rc;
    if (int rc = d_flags.alloc(static_cast<size_t>(n_sources) * 4)) return rc;
    if (int rc = d_paths.alloc(static_cast<size_t>(n_sources) * LAT_MAX_PATHS * R * 4)) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_off.p, adj_off, static_cast<size_t>(n_nodes + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_entries) {
        CUDA_TRY(cudaMemcpyAsync(d_nbr.p, adj_nbr, static_cast<size_t>(n_entries) * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_ek.p, adj_kind, static_cast<size_t>(n_entries), cudaMemcpyHostToDevice, st));
    }
    if (n_nodes) {
        CUDA_TRY(cudaMemcpyAsync(d_nk.p, node_kind, static_cast<size_t>(n_nodes), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_key.p, node_key, static_cast<size_t>(n_nodes) * 4, cudaMemcpyHostToDevice, st));
    }
    if (n_sources) {
        CUDA_TRY(cudaMemcpyAsync(d_src.p, sources, static_cast<size_t>(n_sources) * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_skey.p, source_key, static_cast<size_t>(n_sources) * 4, cudaMemcpyHostToDevice, st));
    }
    LateralArgs A{};
    A.n = n_nodes; A.off = d_off.as<int64_t>(); A.nbr = d_nbr.as<int32_t>(); A.ekind = d_ek.as<uint8_t>(); A.nkind = d_nk.as<uint8_t>(); A.nkey = d_key.as<int32_t>();
    A.n_sources = n_sources; A.src = d_src.as<int32_t>(); A.src_key = d_skey.as<int32_t>(); A.max_depth = max_depth; A.W = W;
    A.ring_cap = static_cast<int32_t>(ring_cap); A.ring = d_ring.as<int32_t>(); A.max_pops = max_pops > 0 ? max_pops : 50'000'000ll;
    A.out_count = d_cnt.as<int32_t>(); A.out_paths = d_paths.as<int32_t>(); A.out_flags = d_flags.as<int32_t>();
    CUDA_TRY(cudaEventRecord(sg.e0, st));
    if (n_sources) {
        lateral_search_kernel<<<static_cast<unsigned>(blocks), 128, 0, st>>>(A);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(sg.e1, st));
    std::vector<int32_t> counts(static_cast<size_t>(n_sources));
    abb_lateral_result *r = new abb_lateral_result();
    r->W = W; r->flags.resize(static_cast<size_t>(n_sources)); r->off.assign(static_cast<size_t>(n_sources) + 1, 0);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess && n_sources) e = cudaMemcpy(counts.data(), d_cnt.p, static_cast<size_t>(n_sources) * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && n_sources) e = cudaMemcpy(r->flags.data(), d_flags.p, static_cast<size_t>(n_sources) * 4, cudaMemcpyDeviceToHost);
    std::vector<int32_t> all;
    if (e == cudaSuccess && n_sources) {
        all.resize(static_cast<size_t>(n_sources) * LAT_MAX_PATHS * R);
        e = cudaMemcpy(all.data(), d_paths.p, all.size() * 4, cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess) { delete r; return fail(ABB_ERR_CUDA, "lateral search: %s", cudaGetErrorString(e)); }
    float ms = 0.f; cudaEventElapsedTime(&ms, sg.e0, sg.e1); r->ms = ms;
    for (int64_t q = 0; q < n_sources; q++) r->off[q + 1] = r->off[q] + counts[q];
    r->records.resize(static_cast<size_t>(r->off[n_sources]) * R);
    for (int64_t q = 0; q < n_sources; q++)
        if (counts[q]) memcpy(r->records.data() + r->off[q] * R, all.data() + q * LAT_MAX_PATHS * R, static_cast<size_t>(counts[q]) * R * 4);
    *out = r;
    return ABB_OK;
}
