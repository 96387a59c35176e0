// reach_host.inl — host orchestration of compute_dependency_reach (included by abb200.cu).
// Reference: graph/dependency_reach.py:109-220.

struct Tmp {   // scoped device allocation
    void *p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        return e == cudaSuccess ? ABB_OK : fail(ABB_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct abb_reach_result {
    std::vector<int32_t> pkg_ids, pkg_agents, pkg_minhop, vuln_ids, vuln_pkgs, vuln_agents, vuln_minhop;
    std::vector<int64_t> pkg_off, vuln_poff, vuln_aoff;
};
extern "C" void abb_reach_result_free(abb_reach_result *r) { delete r; }
extern "C" int64_t abb_reach_n_packages(const abb_reach_result *r) { return static_cast<int64_t>(r->pkg_ids.size()); }
extern "C" const int32_t *abb_reach_pkg_ids(const abb_reach_result *r) { return r->pkg_ids.data(); }
extern "C" const int64_t *abb_reach_pkg_off(const abb_reach_result *r) { return r->pkg_off.data(); }
extern "C" const int32_t *abb_reach_pkg_agents(const abb_reach_result *r) { return r->pkg_agents.data(); }
extern "C" const int32_t *abb_reach_pkg_minhop(const abb_reach_result *r) { return r->pkg_minhop.data(); }
extern "C" int64_t abb_reach_n_vulns(const abb_reach_result *r) { return static_cast<int64_t>(r->vuln_ids.size()); }
extern "C" const int32_t *abb_reach_vuln_ids(const abb_reach_result *r) { return r->vuln_ids.data(); }
extern "C" const int64_t *abb_reach_vuln_poff(const abb_reach_result *r) { return r->vuln_poff.data(); }
extern "C" const int32_t *abb_reach_vuln_pkgs(const abb_reach_result *r) { return r->vuln_pkgs.data(); }
extern "C" const int64_t *abb_reach_vuln_aoff(const abb_reach_result *r) { return r->vuln_aoff.data(); }
extern "C" const int32_t *abb_reach_vuln_agents(const abb_reach_result *r) { return r->vuln_agents.data(); }
extern "C" const int32_t *abb_reach_vuln_minhop(const abb_reach_result *r) { return r->vuln_minhop.data(); }

static inline unsigned nblk(int64_t n, int t) { return static_cast<unsigned>(std::max<int64_t>(1, (n + t - 1) / t)); }

// sort (key,val) pairs by key, then drop adjacent duplicate keys; returns the unique count
static int sort_unique_pairs(cudaStream_t st, unsigned long long *k_in, int32_t *v_in, unsigned long long *k_tmp, int32_t *v_tmp, int64_t n,
                             int64_t *n_unique, bool unique, int end_bit = 64) {
    if (n == 0) { *n_unique = 0; return ABB_OK; }
    if (n >= (1ll << 31)) return fail(ABB_ERR_ARG, "reach pipeline limited to 2^31 pairs per batch");
    size_t tb = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_tmp, v_in, v_tmp, static_cast<int>(n), 0, end_bit, st));
    Tmp t; if (int rc = t.alloc(tb)) return rc;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(t.p, tb, k_in, k_tmp, v_in, v_tmp, static_cast<int>(n), 0, end_bit, st));
    g_launches++;
    if (!unique) {
        CUDA_TRY(cudaMemcpyAsync(k_in, k_tmp, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(v_in, v_tmp, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, st));
        *n_unique = n;
        CUDA_TRY(cudaStreamSynchronize(st));
        return ABB_OK;
    }
    Tmp nsel; if (int rc = nsel.alloc(8)) return rc;
    size_t ub = 0;
    CUDA_TRY(cub::DeviceSelect::UniqueByKey(nullptr, ub, k_tmp, v_tmp, k_in, v_in, nsel.as<int64_t>(), static_cast<int>(n), st));
    Tmp u; if (int rc = u.alloc(ub)) return rc;
    CUDA_TRY(cub::DeviceSelect::UniqueByKey(u.p, ub, k_tmp, v_tmp, k_in, v_in, nsel.as<int64_t>(), static_cast<int>(n), st));
    g_launches++;
    int64_t ns = 0;
    CUDA_TRY(cudaMemcpyAsync(&ns, nsel.p, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *n_unique = ns;
    return ABB_OK;
}

static int exclusive_scan_i64(cudaStream_t st, const int64_t *in, int64_t *out, int64_t n) {
    size_t tb = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, static_cast<int>(n), st));
    Tmp t; if (int rc = t.alloc(tb)) return rc;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(t.p, tb, in, out, static_cast<int>(n), st));
    g_launches++;
    CUDA_TRY(cudaStreamSynchronize(st));
    return ABB_OK;
}

extern "C" int abb_dependency_reach_host(abb_graph *g, const int32_t *agents, int64_t n_agents, uint32_t rel_mask, uint32_t vuln_pkg_mask,
                                         abb_reach_result **out) {
    NvtxRange nvtx_("abb_dependency_reach_host");
    if (!g || !out || n_agents < 0 || (n_agents && !agents)) return fail(ABB_ERR_ARG, "bad arguments");
    DeviceGuard dg(g->device);
    std::lock_guard<std::mutex> lk(g->mu);
    cudaStream_t st = g->stream;
    const int64_t n = g->v.n;

    // ---- pass 1: BFS distances from every agent, packages only (dependency_reach.py:121-130)
    abb_walk_spec spec = abb_spec_distances_along(rel_mask, 1u << ET_PACKAGE);
    abb_walk_io io{}; unsigned long long totals[3] = {0, 0, 0}; int64_t h2d = 0;
    if (int rc = walk_device_stage(g, &spec, agents, nullptr, nullptr, n_agents, &io, totals, &h2d)) return rc;
    const int64_t T = static_cast<int64_t>(totals[0]);

    // ---- pass 2a: invert to package -> (agents sorted by id string, min hops)  (:132-145)
    Tmp keys, vals, keys2, vals2, cnt, noff, minhop;
    if (int rc = keys.alloc(static_cast<size_t>(T) * 8)) return rc;
    if (int rc = vals.alloc(static_cast<size_t>(T) * 4)) return rc;
    if (int rc = keys2.alloc(static_cast<size_t>(T) * 8)) return rc;
    if (int rc = vals2.alloc(static_cast<size_t>(T) * 4)) return rc;
    if (int rc = cnt.alloc(static_cast<size_t>(n + 1) * 8)) return rc;
    if (int rc = noff.alloc(static_cast<size_t>(n + 2) * 8)) return rc;
    if (int rc = minhop.alloc(static_cast<size_t>(n + 1) * 4)) return rc;
    CUDA_TRY(cudaMemsetAsync(cnt.p, 0, static_cast<size_t>(n + 1) * 8, st));
    fill_i32_kernel<<<nblk(n, 256), 256, 0, st>>>(minhop.as<int32_t>(), n, 0x7FFFFFFF); g_launches++;
    if (n_agents) {
        reach_invert_kernel<<<nblk(n_agents, 8), 256, 0, st>>>(n_agents, io.roots, io.q_start, io.q_count, io.nodes, io.depth, g->v.rank,
                                                             keys.as<unsigned long long>(), vals.as<int32_t>(), cnt.as<unsigned long long>(),
                                                             minhop.as<int32_t>());
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    int64_t dummy = 0;
    if (int rc = sort_unique_pairs(st, keys.as<unsigned long long>(), vals.as<int32_t>(), keys2.as<unsigned long long>(), vals2.as<int32_t>(), T, &dummy, false)) return rc;
    if (int rc = exclusive_scan_i64(st, cnt.as<int64_t>(), noff.as<int64_t>(), n + 1)) return rc;

    // ---- host: package / vulnerability node lists in node order
    std::vector<uint8_t> ntype(static_cast<size_t>(n));
    if (n) CUDA_TRY(cudaMemcpy(ntype.data(), g->v.ntype, static_cast<size_t>(n), cudaMemcpyDeviceToHost));
    std::unique_ptr<abb_reach_result> r_owner(new abb_reach_result());   // freed on every early return
    abb_reach_result *r = r_owner.get();
    for (int64_t u = 0; u < n; u++) {
        if (ntype[u] == ET_PACKAGE) r->pkg_ids.push_back(static_cast<int32_t>(u));
        else if (ntype[u] == ET_VULN) r->vuln_ids.push_back(static_cast<int32_t>(u));
    }
    const int64_t nv = static_cast<int64_t>(r->vuln_ids.size());

    // ---- pass 2b: vulnerability -> attached packages (unique, sorted by id string)  (:201-220)
    Tmp d_vulns, vcounts, voff;
    if (int rc = d_vulns.alloc(static_cast<size_t>(nv + 1) * 4)) return rc;
    if (int rc = vcounts.alloc(static_cast<size_t>(nv + 2) * 8)) return rc;
    if (int rc = voff.alloc(static_cast<size_t>(nv + 2) * 8)) return rc;
    CUDA_TRY(cudaMemsetAsync(vcounts.p, 0, static_cast<size_t>(nv + 2) * 8, st));
    if (nv) CUDA_TRY(cudaMemcpyAsync(d_vulns.p, r->vuln_ids.data(), static_cast<size_t>(nv) * 4, cudaMemcpyHostToDevice, st));
    if (nv) { reach_vuln_pkgs_kernel<false><<<nblk(nv, 128), 128, 0, st>>>(g->v, nv, d_vulns.as<int32_t>(), vuln_pkg_mask, nullptr, vcounts.as<int64_t>(), nullptr, nullptr); g_launches++; }
    if (int rc = exclusive_scan_i64(st, vcounts.as<int64_t>(), voff.as<int64_t>(), nv + 1)) return rc;
    int64_t P0 = 0;
    CUDA_TRY(cudaMemcpy(&P0, voff.as<int64_t>() + nv, 8, cudaMemcpyDeviceToHost));
    Tmp pk, pv, pk2, pv2;
    if (int rc = pk.alloc(static_cast<size_t>(P0) * 8)) return rc;
    if (int rc = pv.alloc(static_cast<size_t>(P0) * 4)) return rc;
    if (int rc = pk2.alloc(static_cast<size_t>(P0) * 8)) return rc;
    if (int rc = pv2.alloc(static_cast<size_t>(P0) * 4)) return rc;
    if (nv) { reach_vuln_pkgs_kernel<true><<<nblk(nv, 128), 128, 0, st>>>(g->v, nv, d_vulns.as<int32_t>(), vuln_pkg_mask, voff.as<int64_t>(), nullptr, pk.as<unsigned long long>(), pv.as<int32_t>()); g_launches++; }
    int64_t P = 0;
    if (int rc = sort_unique_pairs(st, pk.as<unsigned long long>(), pv.as<int32_t>(), pk2.as<unsigned long long>(), pv2.as<int32_t>(), P0, &P, true)) return rc;
    // per-vulnerability package counts -> vuln_poff
    Tmp gcnt, goff;
    if (int rc = gcnt.alloc(static_cast<size_t>(nv + 2) * 8)) return rc;
    if (int rc = goff.alloc(static_cast<size_t>(nv + 2) * 8)) return rc;
    CUDA_TRY(cudaMemsetAsync(gcnt.p, 0, static_cast<size_t>(nv + 2) * 8, st));
    if (P) { reach_group_counts_kernel<<<nblk(P, 256), 256, 0, st>>>(P, pk.as<unsigned long long>(), gcnt.as<unsigned long long>()); g_launches++; }
    if (int rc = exclusive_scan_i64(st, gcnt.as<int64_t>(), goff.as<int64_t>(), nv + 1)) return rc;
    r->vuln_poff.resize(static_cast<size_t>(nv) + 1);
    CUDA_TRY(cudaMemcpy(r->vuln_poff.data(), goff.p, static_cast<size_t>(nv + 1) * 8, cudaMemcpyDeviceToHost));
    r->vuln_pkgs.resize(static_cast<size_t>(P));
    if (P) CUDA_TRY(cudaMemcpy(r->vuln_pkgs.data(), pv.p, static_cast<size_t>(P) * 4, cudaMemcpyDeviceToHost));

    // ---- pass 2c: vulnerability -> union of agents of its reachable packages, min of their mins  (:147-164)
    Tmp pcounts, poff, vmin;
    if (int rc = pcounts.alloc(static_cast<size_t>(P + 2) * 8)) return rc;
    if (int rc = poff.alloc(static_cast<size_t>(P + 2) * 8)) return rc;
    if (int rc = vmin.alloc(static_cast<size_t>(nv + 1) * 4)) return rc;
    CUDA_TRY(cudaMemsetAsync(pcounts.p, 0, static_cast<size_t>(P + 2) * 8, st));
    fill_i32_kernel<<<nblk(nv, 256), 256, 0, st>>>(vmin.as<int32_t>(), nv, 0x7FFFFFFF); g_launches++;
    if (P) { reach_pair_counts_kernel<<<nblk(P, 256), 256, 0, st>>>(P, pk.as<unsigned long long>(), pv.as<int32_t>(), cnt.as<unsigned long long>(), minhop.as<int32_t>(), pcounts.as<int64_t>(), vmin.as<int32_t>()); g_launches++; }
    if (int rc = exclusive_scan_i64(st, pcounts.as<int64_t>(), poff.as<int64_t>(), P + 1)) return rc;
    int64_t A0 = 0;
    CUDA_TRY(cudaMemcpy(&A0, poff.as<int64_t>() + P, 8, cudaMemcpyDeviceToHost));
    Tmp ak, av, ak2, av2;
    if (int rc = ak.alloc(static_cast<size_t>(A0) * 8)) return rc;
    if (int rc = av.alloc(static_cast<size_t>(A0) * 4)) return rc;
    if (int rc = ak2.alloc(static_cast<size_t>(A0) * 8)) return rc;
    if (int rc = av2.alloc(static_cast<size_t>(A0) * 4)) return rc;
    if (P) { reach_pair_fill_kernel<<<nblk(P, 8), 256, 0, st>>>(P, pk.as<unsigned long long>(), pv.as<int32_t>(), poff.as<int64_t>(), noff.as<int64_t>(), vals.as<int32_t>(), g->v.rank, ak.as<unsigned long long>(), av.as<int32_t>()); g_launches++; }
    int64_t AU = 0;
    if (int rc = sort_unique_pairs(st, ak.as<unsigned long long>(), av.as<int32_t>(), ak2.as<unsigned long long>(), av2.as<int32_t>(), A0, &AU, true)) return rc;
    CUDA_TRY(cudaMemsetAsync(gcnt.p, 0, static_cast<size_t>(nv + 2) * 8, st));
    if (AU) { reach_group_counts_kernel<<<nblk(AU, 256), 256, 0, st>>>(AU, ak.as<unsigned long long>(), gcnt.as<unsigned long long>()); g_launches++; }
    if (int rc = exclusive_scan_i64(st, gcnt.as<int64_t>(), goff.as<int64_t>(), nv + 1)) return rc;
    minhop_finalize_kernel<<<nblk(n, 256), 256, 0, st>>>(minhop.as<int32_t>(), n); g_launches++;
    minhop_finalize_kernel<<<nblk(nv, 256), 256, 0, st>>>(vmin.as<int32_t>(), nv); g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));

    // ---- results to the host
    r->vuln_aoff.resize(static_cast<size_t>(nv) + 1);
    CUDA_TRY(cudaMemcpy(r->vuln_aoff.data(), goff.p, static_cast<size_t>(nv + 1) * 8, cudaMemcpyDeviceToHost));
    r->vuln_agents.resize(static_cast<size_t>(AU));
    if (AU) CUDA_TRY(cudaMemcpy(r->vuln_agents.data(), av.p, static_cast<size_t>(AU) * 4, cudaMemcpyDeviceToHost));
    r->vuln_minhop.resize(static_cast<size_t>(nv));
    if (nv) CUDA_TRY(cudaMemcpy(r->vuln_minhop.data(), vmin.p, static_cast<size_t>(nv) * 4, cudaMemcpyDeviceToHost));
    r->pkg_agents.resize(static_cast<size_t>(T));
    if (T) CUDA_TRY(cudaMemcpy(r->pkg_agents.data(), vals.p, static_cast<size_t>(T) * 4, cudaMemcpyDeviceToHost));
    std::vector<int64_t> h_noff(static_cast<size_t>(n) + 1);
    CUDA_TRY(cudaMemcpy(h_noff.data(), noff.p, static_cast<size_t>(n + 1) * 8, cudaMemcpyDeviceToHost));
    std::vector<int32_t> h_min(static_cast<size_t>(n));
    if (n) CUDA_TRY(cudaMemcpy(h_min.data(), minhop.p, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost));
    const size_t np = r->pkg_ids.size();
    r->pkg_off.resize(np + 1); r->pkg_minhop.resize(np);
    for (size_t i = 0; i < np; i++) { r->pkg_off[i] = h_noff[r->pkg_ids[i]]; r->pkg_minhop[i] = h_min[r->pkg_ids[i]]; }
    r->pkg_off[np] = T;   // only package nodes were emitted, so the sorted pair array is exactly their concatenation
    *out = r_owner.release();
    return ABB_OK;
}
