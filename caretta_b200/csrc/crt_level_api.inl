// crt_level_api.inl -- all independent nodes of one level of the guide tree in one call ("nodes at
// the same tree depth are independent; batch per tree level").  Included at the end of crt_api.cu; kernels in crt_node.cuh.

extern "C" {

int crt_progressive_level(crt_ctx *c, int32_t n_nodes, int32_t d, const double *tensors, const double *coords, const double *weights,
                          const int64_t *offsets, const double *mult, double gamma_tensor, double gamma_coords, double gamma_weight,
                          double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2, int32_t *aln_len, double *tensors_mean,
                          double *coords_mean, double *weights_mean, double *score, int32_t *status)
{
    if (!c || !tensors || !coords || !weights || !offsets || !mult || !aln1 || !aln2 || !aln_len || !tensors_mean || !coords_mean ||
        !weights_mean)
        return fail(CRT_E_ARG, "null argument");
    if (n_nodes <= 0) return fail(CRT_E_ARG, "n_nodes must be > 0");
    CU(cudaSetDevice(c->device));
    int rc;
    if (!c->node_ctx && (rc = crt_create(c->device, &c->node_ctx))) return rc;
    crt_ctx *nc = c->node_ctx;
    nc->stage1_only = true;
    const bool timeline = getenv("CARETTA_B200_TIMELINE") && atoi(getenv("CARETTA_B200_TIMELINE")) != 0;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    // ---- stage 1 of score_function for every node: the fp64 pair kernels on the packed children, pairs (2k, 2k+1)
    if ((rc = crt_set_chains(nc, coords, tensors, offsets, 2 * n_nodes, d))) return rc;
    crt_params prm{};
    prm.gamma_tensor = gamma_tensor; prm.gamma_coords = gamma_coords; prm.sw_gap = 0.0; prm.precision = CRT_FP64;
    std::vector<int32_t> pi((size_t)n_nodes), pj((size_t)n_nodes), st1((size_t)n_nodes);
    for (int k = 0; k < n_nodes; ++k) { pi[(size_t)k] = 2 * k; pj[(size_t)k] = 2 * k + 1; }
    const double t_chains = now();
    if ((rc = crt_pairwise_list(nc, &prm, pi.data(), pj.data(), n_nodes, nullptr, nullptr, nullptr, nullptr, st1.data(), nullptr, nullptr,
                                nullptr, 0)))
        return rc;
    double pair_ms = nc->elapsed_ms;
    long long launches = nc->launches;
    const double t_pairs = now();
    // ---- score matrices, affine DTW and the intermediate nodes, in chunks of nodes bounded by the workspace budget
    const long long total = offsets[2 * n_nodes];
    std::vector<DpProblem> probs((size_t)n_nodes);
    long long max_cells = 0;
    int max_len = 0;
    for (int k = 0; k < n_nodes; ++k) {
        DpProblem &p = probs[(size_t)k];
        p.n = (int)(offsets[2 * k + 1] - offsets[2 * k]);
        p.m = (int)(offsets[2 * k + 2] - offsets[2 * k + 1]);
        p.aln_off = offsets[2 * k];
        max_cells = std::max(max_cells, (long long)p.n * p.m);
        max_len = std::max(max_len, p.n + p.m);
    }
    const long long cell_budget = std::max<long long>(max_cells, (long long)std::min<size_t>(c->mem_total / 4, (size_t)24 << 30) / 9);
    if ((rc = c->lv_probs.ensure((size_t)n_nodes))) return rc;
    if ((rc = c->lv_mult.ensure((size_t)n_nodes * 2))) return rc;
    if ((rc = c->lv_xf2.ensure((size_t)n_nodes * XF))) return rc;
    if ((rc = c->nd_w.ensure((size_t)total))) return rc;
    if ((rc = c->nd_a1.ensure((size_t)total + 1))) return rc;
    if ((rc = c->nd_a2.ensure((size_t)total + 1))) return rc;
    if ((rc = c->nd_len.ensure((size_t)n_nodes))) return rc;
    if ((rc = c->nd_f.ensure((size_t)n_nodes * 3))) return rc;
    if ((rc = c->nd_score.ensure((size_t)n_nodes))) return rc;
    if ((rc = c->nd_t.ensure((size_t)total * d))) return rc;
    if ((rc = c->nd_c.ensure((size_t)total * 3))) return rc;
    if ((rc = c->nd_wm.ensure((size_t)total))) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->nd_w.p, weights, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->lv_mult.p, mult, sizeof(double) * (size_t)n_nodes * 2, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev0, st));
    int k0 = 0;
    while (k0 < n_nodes) {
        long long cells = 0, rows = 0;
        int k1 = k0;
        while (k1 < n_nodes && k1 - k0 < 65535 && (k1 == k0 || cells + (long long)probs[(size_t)k1].n * probs[(size_t)k1].m <= cell_budget)) {
            DpProblem &p = probs[(size_t)k1];
            p.s_off = cells; p.b_off = cells; p.bnd_off = rows;
            cells += (long long)p.n * p.m;
            rows += p.n;
            ++k1;
        }
        const int nk = k1 - k0;
        if ((rc = c->nd_S.ensure((size_t)cells))) return rc;
        if ((rc = c->nd_B.ensure((size_t)cells))) return rc;
        if ((rc = c->nd_bnd.ensure((size_t)rows * 2 + 2))) return rc;
        DpProblem *dp = c->lv_probs.p + k0;
        CU(cudaMemcpyAsync(dp, probs.data() + k0, sizeof(DpProblem) * (size_t)nk, cudaMemcpyHostToDevice, st));
        long long mc = 0;
        int ml = 0;
        for (int k = k0; k < k1; ++k) {
            mc = std::max(mc, (long long)probs[(size_t)k].n * probs[(size_t)k].m);
            ml = std::max(ml, probs[(size_t)k].n + probs[(size_t)k].m);
        }
        k_level_score<<<dim3((unsigned)((mc + 255) / 256), (unsigned)nk), 256, 0, st>>>(dp, nc->coords.p, c->nd_w.p, nc->xform.p + (size_t)k0 * XF,
                                                                                       c->lv_mult.p + (size_t)k0 * 2, -gamma_coords, -gamma_weight,
                                                                                       c->nd_S.p);
        k_dtw_fill<<<nk, 32, 0, st>>>(dp, nk, c->nd_S.p, c->nd_B.p, c->nd_bnd.p, c->nd_f.p + (size_t)k0 * 3, gap_open, gap_extend);
        k_dtw_trace<<<(nk + 31) / 32, 32, 0, st>>>(dp, nk, c->nd_B.p, c->nd_f.p + (size_t)k0 * 3, c->nd_a1.p, c->nd_a2.p, c->nd_len.p + k0,
                                                   c->nd_score.p + k0);
        k_level_kabsch<<<(nk + 31) / 32, 32, 0, st>>>(dp, nk, nc->coords.p, c->nd_a1.p, c->nd_a2.p, c->nd_len.p + k0, c->lv_xf2.p + (size_t)k0 * XF);
        k_level_mean<<<dim3((unsigned)((ml + 127) / 128), (unsigned)nk), 128, 0, st>>>(dp, nc->tensors.p, nc->coords.p, c->nd_w.p, d, c->nd_a1.p,
                                                                                      c->nd_a2.p, c->nd_len.p + k0, c->lv_xf2.p + (size_t)k0 * XF,
                                                                                      c->nd_t.p, c->nd_c.p, c->nd_wm.p);
        CU(cudaGetLastError());
        launches += 5;
        k0 = k1;
    }
    CU(cudaEventRecord(c->ev1, st));
    const double t_launched = now();
    CU(cudaMemcpyAsync(aln_len, c->nd_len.p, sizeof(int) * (size_t)n_nodes, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(aln1, c->nd_a1.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(aln2, c->nd_a2.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tensors_mean, c->nd_t.p, sizeof(double) * (size_t)total * d, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(coords_mean, c->nd_c.p, sizeof(double) * (size_t)total * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(weights_mean, c->nd_wm.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st));
    if (score) CU(cudaMemcpyAsync(score, c->nd_score.p, sizeof(double) * (size_t)n_nodes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->elapsed_ms = pair_ms + ms;
    if (timeline)
        fprintf(stderr, "[level] nodes %5d residues %8lld: set_chains %7.3f  pair run %7.3f (device %7.3f)  enqueue %7.3f  kernels+D2H %7.3f (device %7.3f) ms\n",
                n_nodes, total, t_chains - t_begin, t_pairs - t_chains, pair_ms, t_launched - t_pairs, now() - t_launched, (double)ms);
    c->launches = launches;
    for (int k = 0; k < n_nodes; ++k)
        if (aln_len[k] < 0 || aln_len[k] > probs[(size_t)k].n + probs[(size_t)k].m) return fail(CRT_E_STATE, "alignment length %d of node %d out of range", aln_len[k], k);
    if (status) std::memcpy(status, st1.data(), sizeof(int32_t) * (size_t)n_nodes);
    return 0;
}

}  // extern "C"
