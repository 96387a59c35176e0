// centrality_host.inl — abb_bottleneck_host (included by abb200.cu).  Reference: graph/container.py:548-567,
// graph_backend.py:127-155.

extern "C" int abb_bottleneck_host(abb_graph *g, const int32_t *sources, int64_t n_sources, uint64_t *scores_out) {
    NvtxRange nvtx_("abb_bottleneck_host");
    if (!g || !scores_out || n_sources < 0 || (n_sources && !sources)) return fail(ABB_ERR_ARG, "bad arguments");
    DeviceGuard dg(g->device);
    std::lock_guard<std::mutex> lk(g->mu);
    cudaStream_t st = g->stream;
    const int64_t n = g->v.n;
    // unbounded forward BFS over every adjacency entry, first-discoverer parents and depths, source not emitted
    abb_walk_spec spec = abb_spec_bfs(0, 0);
    spec.max_depth = -1;
    spec.flags |= ABB_WALK_DEPTHS;
    abb_walk_io io{}; unsigned long long totals[3] = {0, 0, 0}; int64_t h2d = 0;
    if (int rc = walk_device_stage(g, &spec, sources, nullptr, nullptr, n_sources, &io, totals, &h2d)) return rc;
    const int64_t T = static_cast<int64_t>(totals[0]);
    Tmp desc, scores;
    if (int rc = desc.alloc(static_cast<size_t>(T + 1) * 4)) return rc;
    if (int rc = scores.alloc(static_cast<size_t>(n + 1) * 8)) return rc;
    CUDA_TRY(cudaMemsetAsync(scores.p, 0, static_cast<size_t>(n + 1) * 8, st));
    if (n_sources) {
        const unsigned grid = static_cast<unsigned>(std::min<int64_t>(n_sources, 148 * 2));
        subtree_rollup_kernel<<<grid, 1024, 0, st>>>(n_sources, io.q_start, io.q_count, io.q_maxd, io.nodes, io.parent, io.depth, desc.as<int32_t>(),
                                                    scores.as<unsigned long long>());
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    if (n) CUDA_TRY(cudaMemcpyAsync(scores_out, scores.p, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ABB_OK;
}
