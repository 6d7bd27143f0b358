// crt_level_api.inl -- all independent nodes of one level of the guide tree in one call ("nodes at
// the same tree depth are independent; batch per tree level").  Included at the end of crt_api.cu; kernels in crt_node.cuh.
//
//   crt_progressive_level   stateless: the children of the level come from and the new nodes go back to host arrays
//   crt_msa_begin / crt_msa_level / crt_msa_lengths / crt_msa_fetch / crt_msa_end
//                           the whole progressive alignment with the sequences resident on the device: a pool holds the leaves
//                           and every intermediate node; a level gathers its children from the pool and appends its nodes to it;
//                           only the alignments (two int32 per column) travel to the host per level.

namespace {

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct LevelParams {
    double gamma_tensor, gamma_coords, gamma_weight, gap_open, gap_extend;
    // the C ABI's sentinels in gamma_coords: CRT_GAMMA_COORDS_FLEXIBLE (-1) = score_function and mean_function flexible=True,
    // CRT_GAMMA_COORDS_FLEXIBLE_SCORE (-2) = score_function flexible=True only (the node still superposes its coordinates)
    bool flexible() const { return gamma_coords < 0.0; }
    bool flexible_mean() const { return gamma_coords < 0.0 && gamma_coords > -1.5; }
};

// The level on a chain set that is already in the node context `nc` (children 2k, 2k+1 = node k, packed offsets `offsets`) with
// the packed consensus weights in c->nd_w.  Node outputs go to (t_out, c_out, w_out) at row out_off[k] (host array; nullptr: at the
// node's own packed offset).  aln1 / aln2 / aln_len / score / status: host outputs (aln arrays packed like the children).
int level_core(crt_ctx *c, crt_ctx *nc, int32_t n_nodes, int32_t d, const int64_t *offsets, const double *mult, const LevelParams &lp,
               const long long *out_off, double *t_out, double *c_out, double *w_out, int32_t *aln1, int32_t *aln2, int32_t *aln_len,
               double *score, int32_t *status, double *pair_ms_out, double *level_ms_out)
{
    int rc;
    crt_params prm{};
    prm.gamma_tensor = lp.gamma_tensor; prm.gamma_coords = lp.gamma_coords; prm.sw_gap = 0.0; prm.precision = CRT_FP64;
    std::vector<int32_t> pi((size_t)n_nodes), pj((size_t)n_nodes), st1((size_t)n_nodes);
    for (int k = 0; k < n_nodes; ++k) { pi[(size_t)k] = 2 * k; pj[(size_t)k] = 2 * k + 1; }
    // ---- stage 1 of score_function for every node: the fp64 pair kernels on the packed children, pairs (2k, 2k+1)
    //      (flexible nodes score on the tensors alone: no stage-1 alignment, no superposition)
    const bool flexible = lp.flexible();
    double pair_ms = 0.0;
    long long launches = 0;
    if (!flexible) {
        if ((rc = crt_pairwise_list(nc, &prm, pi.data(), pj.data(), n_nodes, nullptr, nullptr, nullptr, nullptr, st1.data(), nullptr, nullptr,
                                    nullptr, 0)))
            return rc;
        pair_ms = nc->elapsed_ms;
        launches = nc->launches;
    } else {
        std::fill(st1.begin(), st1.end(), 0);
    }
    // ---- score matrices, affine DTW and the intermediate nodes, in chunks of nodes bounded by the workspace budget
    const long long total = offsets[2 * n_nodes];
    std::vector<DpProblem> probs((size_t)n_nodes);
    long long max_cells = 0;
    for (int k = 0; k < n_nodes; ++k) {
        DpProblem &p = probs[(size_t)k];
        p.n = (int)(offsets[2 * k + 1] - offsets[2 * k]);
        p.m = (int)(offsets[2 * k + 2] - offsets[2 * k + 1]);
        p.aln_off = offsets[2 * k];
        max_cells = std::max(max_cells, (long long)p.n * p.m);
    }
    // score matrix + backtrack bytes: 9 B per cell, bounded by min(memory / 4, 24 GB); CARETTA_B200_LEVEL_CELLS overrides (tests)
    long long cell_budget = (long long)std::min<size_t>(c->mem_total / 4, (size_t)24 << 30) / 9;
    if (const char *e = getenv("CARETTA_B200_LEVEL_CELLS")) { const long long v = atoll(e); if (v > 0) cell_budget = v; }
    cell_budget = std::max<long long>(max_cells, cell_budget);
    if ((rc = c->lv_probs.ensure((size_t)n_nodes))) return rc;
    if ((rc = c->lv_mult.ensure((size_t)n_nodes * 2))) return rc;
    if ((rc = c->lv_xf2.ensure((size_t)n_nodes * XF))) return rc;
    if ((rc = c->nd_a1.ensure((size_t)total + 1))) return rc;
    if ((rc = c->nd_a2.ensure((size_t)total + 1))) return rc;
    if ((rc = c->nd_len.ensure((size_t)n_nodes))) return rc;
    if ((rc = c->nd_f.ensure((size_t)n_nodes * 3))) return rc;
    if ((rc = c->nd_score.ensure((size_t)n_nodes))) return rc;
    cudaStream_t st = c->stream;
    const long long *d_out_off = nullptr;
    if (out_off) {
        if ((rc = c->lv_out_off.ensure((size_t)n_nodes))) return rc;
        CU(cudaMemcpyAsync(c->lv_out_off.p, out_off, sizeof(long long) * (size_t)n_nodes, cudaMemcpyHostToDevice, st));
        d_out_off = c->lv_out_off.p;
    }
    CU(cudaMemcpyAsync(c->lv_mult.p, mult, sizeof(double) * (size_t)n_nodes * 2, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev0, st));
    int k0 = 0;
    while (k0 < n_nodes) {
        long long cells = 0, bytes = 0, rows = 0;
        int k1 = k0;
        while (k1 < n_nodes && k1 - k0 < 65535 && (k1 == k0 || cells + (long long)probs[(size_t)k1].n * probs[(size_t)k1].m <= cell_budget)) {
            DpProblem &p = probs[(size_t)k1];
            p.s_off = cells; p.b_off = bytes; p.bnd_off = rows;
            cells += (long long)p.n * p.m;
            bytes += (long long)p.n * dtw_pitch(p.m);
            rows += p.n;
            ++k1;
        }
        const int nk = k1 - k0;
        if ((rc = c->nd_S.ensure((size_t)cells))) return rc;
        if ((rc = c->nd_B.ensure((size_t)bytes))) return rc;
        if ((rc = c->nd_bnd.ensure((size_t)rows * 2 + 2))) return rc;
        DpProblem *dp = c->lv_probs.p + k0;
        CU(cudaMemcpyAsync(dp, probs.data() + k0, sizeof(DpProblem) * (size_t)nk, cudaMemcpyHostToDevice, st));
        long long mc = 0;
        int ml = 0, mm = 0;
        for (int k = k0; k < k1; ++k) {
            mc = std::max(mc, (long long)probs[(size_t)k].n * probs[(size_t)k].m);
            ml = std::max(ml, probs[(size_t)k].n + probs[(size_t)k].m);
            mm = std::max(mm, probs[(size_t)k].m);
        }
        if (flexible)
            k_level_score_flex<<<dim3((unsigned)((mc + 255) / 256), (unsigned)nk), 256, 0, st>>>(dp, nc->tensors.p, d, c->nd_w.p,
                                                                                                c->lv_mult.p + (size_t)k0 * 2,
                                                                                                -lp.gamma_tensor, -lp.gamma_weight, c->nd_S.p);
        else
            k_level_score<<<dim3((unsigned)((mc + 255) / 256), (unsigned)nk), 256, 0, st>>>(dp, nc->coords.p, c->nd_w.p, nc->xform.p + (size_t)k0 * XF,
                                                                                           c->lv_mult.p + (size_t)k0 * 2, -lp.gamma_coords,
                                                                                           -lp.gamma_weight, c->nd_S.p);
        CU(launch_dtw_fill(dp, nk, mm, c->nd_S.p, c->nd_B.p, c->nd_bnd.p, c->nd_f.p + (size_t)k0 * 3, lp.gap_open, lp.gap_extend, st));
        k_dtw_trace_w<<<nk, 32, 0, st>>>(dp, nk, c->nd_B.p, c->nd_f.p + (size_t)k0 * 3, c->nd_a1.p, c->nd_a2.p, c->nd_len.p + k0,
                                                   c->nd_score.p + k0);
        if (lp.flexible_mean())   // mean_function(flexible=True) makes no coordinates (:359-360): flag 0 = no superposition, raw means (unused)
            CU(cudaMemsetAsync(c->lv_xf2.p + (size_t)k0 * XF, 0, sizeof(double) * (size_t)nk * XF, st));
        else
            k_level_kabsch<<<nk, 32, 0, st>>>(dp, nk, nc->coords.p, c->nd_a1.p, c->nd_a2.p, c->nd_len.p + k0, c->lv_xf2.p + (size_t)k0 * XF);
        k_level_mean<<<dim3((unsigned)((ml + 127) / 128), (unsigned)nk), 128, 0, st>>>(dp, nc->tensors.p, nc->coords.p, c->nd_w.p, d, c->nd_a1.p,
                                                                                      c->nd_a2.p, c->nd_len.p + k0, c->lv_xf2.p + (size_t)k0 * XF,
                                                                                      d_out_off ? d_out_off + k0 : nullptr, t_out, c_out, w_out);
        CU(cudaGetLastError());
        launches += 5;
        k0 = k1;
    }
    CU(cudaEventRecord(c->ev1, st));
    CU(cudaMemcpyAsync(aln_len, c->nd_len.p, sizeof(int) * (size_t)n_nodes, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(aln1, c->nd_a1.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(aln2, c->nd_a2.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, st));
    if (score) CU(cudaMemcpyAsync(score, c->nd_score.p, sizeof(double) * (size_t)n_nodes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->elapsed_ms = pair_ms + ms;
    c->launches = launches;
    if (pair_ms_out) *pair_ms_out = pair_ms;
    if (level_ms_out) *level_ms_out = ms;
    for (int k = 0; k < n_nodes; ++k)
        if (aln_len[k] < 0 || aln_len[k] > probs[(size_t)k].n + probs[(size_t)k].m) return fail(CRT_E_STATE, "alignment length %d of node %d out of range", aln_len[k], k);
    if (status) std::memcpy(status, st1.data(), sizeof(int32_t) * (size_t)n_nodes);
    return 0;
}

bool level_timeline() { return getenv("CARETTA_B200_TIMELINE") && atoi(getenv("CARETTA_B200_TIMELINE")) != 0; }

}  // namespace

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
    const double t_begin = now_ms();
    if ((rc = crt_set_chains(nc, coords, tensors, offsets, 2 * n_nodes, d))) return rc;
    const double t_chains = now_ms();
    const long long total = offsets[2 * n_nodes];
    if ((rc = c->nd_w.ensure((size_t)total))) return rc;
    if ((rc = c->nd_t.ensure((size_t)total * d))) return rc;
    if ((rc = c->nd_c.ensure((size_t)total * 3))) return rc;
    if ((rc = c->nd_wm.ensure((size_t)total))) return rc;
    CU(cudaMemcpyAsync(c->nd_w.p, weights, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, c->stream));
    const LevelParams lp{gamma_tensor, gamma_coords, gamma_weight, gap_open, gap_extend};
    double pair_ms = 0, level_ms = 0;
    if ((rc = level_core(c, nc, n_nodes, d, offsets, mult, lp, nullptr, c->nd_t.p, c->nd_c.p, c->nd_wm.p, aln1, aln2, aln_len, score, status,
                         &pair_ms, &level_ms)))
        return rc;
    const double t_core = now_ms();
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(tensors_mean, c->nd_t.p, sizeof(double) * (size_t)total * d, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(coords_mean, c->nd_c.p, sizeof(double) * (size_t)total * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(weights_mean, c->nd_wm.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (level_timeline())
        fprintf(stderr, "[level] nodes %5d residues %8lld: set_chains %7.3f  pair run + level kernels %7.3f (device %7.3f + %7.3f)  D2H %7.3f ms\n",
                n_nodes, total, t_chains - t_begin, t_core - t_chains, pair_ms, level_ms, now_ms() - t_core);
    return 0;
}

/* ---- device-resident progressive alignment ---- */

int crt_msa_begin(crt_ctx *c, double consensus_weight, int32_t *n_sequences)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (c->N <= 0) return fail(CRT_E_STATE, "crt_set_chains has not been called");
    CU(cudaSetDevice(c->device));
    crt_ctx::MsaPool &P = c->pool;
    P.active = false;
    P.d = c->d;
    const long long total = c->total;
    int rc;
    if ((rc = P.t.ensure((size_t)total * 3 * c->d)) || (rc = P.c.ensure((size_t)total * 3 * 3)) || (rc = P.w.ensure((size_t)total * 3))) return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(P.t.p, c->tensors.p, sizeof(double) * (size_t)total * c->d, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(P.c.p, c->coords.p, sizeof(double) * (size_t)total * 3, cudaMemcpyDeviceToDevice, st));
    k_fill_value<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(P.w.p, total, consensus_weight);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    P.off.assign(c->offsets.begin(), c->offsets.begin() + c->N);
    P.len.resize((size_t)c->N);
    for (int p = 0; p < c->N; ++p) P.len[(size_t)p] = (int)(c->offsets[(size_t)p + 1] - c->offsets[(size_t)p]);
    P.used = total;
    P.active = true;
    P.n_leaves = c->N;
    P.ch1.clear(); P.ch2.clear(); P.al1.clear(); P.al2.clear();
    if (n_sequences) *n_sequences = c->N;
    return 0;
}

int crt_msa_level(crt_ctx *c, int32_t n_nodes, const int32_t *child1, const int32_t *child2, const double *mult, double gamma_tensor,
                  double gamma_coords, double gamma_weight, double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2,
                  int64_t aln_cap, int64_t *aln_off, int32_t *aln_len, double *score, int32_t *status, int32_t *first_new_id)
{
    if (!c || !child1 || !child2 || !mult || !aln1 || !aln2 || !aln_off || !aln_len) return fail(CRT_E_ARG, "null argument");
    crt_ctx::MsaPool &P = c->pool;
    if (!P.active) return fail(CRT_E_STATE, "crt_msa_begin has not been called");
    if (n_nodes <= 0) return fail(CRT_E_ARG, "n_nodes must be > 0");
    const int n_seq = (int)P.len.size(), d = P.d;
    std::vector<int64_t> offsets((size_t)2 * n_nodes + 1, 0);
    for (int k = 0; k < n_nodes; ++k) {
        const int a = child1[k], b = child2[k];
        if (a < 0 || a >= n_seq || b < 0 || b >= n_seq) return fail(CRT_E_ARG, "node %d: child (%d, %d) is not in the pool (%d sequences)", k, a, b, n_seq);
        offsets[(size_t)2 * k + 1] = offsets[(size_t)2 * k] + P.len[(size_t)a];
        offsets[(size_t)2 * k + 2] = offsets[(size_t)2 * k + 1] + P.len[(size_t)b];
    }
    const long long total = offsets[(size_t)2 * n_nodes];
    if (aln_cap < total) return fail(CRT_E_ARG, "alignment buffers hold %lld entries, %lld needed", (long long)aln_cap, total);
    CU(cudaSetDevice(c->device));
    int rc;
    if (!c->node_ctx && (rc = crt_create(c->device, &c->node_ctx))) return rc;
    crt_ctx *nc = c->node_ctx;
    nc->stage1_only = true;
    const double t_begin = now_ms();
    // room for the new nodes at the end of the pool (capacity n + m rows each) -- before any pointer into the pool is used
    const long long need = P.used + total;
    if ((rc = P.t.grow_keep((size_t)need * d, (size_t)P.used * d, c->stream)) || (rc = P.c.grow_keep((size_t)need * 3, (size_t)P.used * 3, c->stream)) ||
        (rc = P.w.grow_keep((size_t)need, (size_t)P.used, c->stream)))
        return rc;
    // gather the children into the node context's chain set (on its stream) and derive its tables
    if ((rc = chains_prepare(nc, offsets.data(), 2 * n_nodes, d))) return rc;
    if ((rc = c->nd_w.ensure((size_t)total))) return rc;
    std::vector<long long> tab((size_t)n_nodes * 6), out_off((size_t)n_nodes);
    int max_len = 1;
    for (int k = 0; k < n_nodes; ++k) {
        const int ch[2] = {child1[k], child2[k]};
        for (int s = 0; s < 2; ++s) {
            long long *t = tab.data() + ((size_t)2 * k + s) * 3;
            t[0] = P.off[(size_t)ch[s]]; t[1] = offsets[(size_t)2 * k + s]; t[2] = P.len[(size_t)ch[s]];
            max_len = std::max(max_len, P.len[(size_t)ch[s]]);
        }
        out_off[(size_t)k] = P.used + offsets[(size_t)2 * k];
    }
    if ((rc = c->lv_tab.ensure(tab.size()))) return rc;
    CU(cudaMemcpyAsync(c->lv_tab.p, tab.data(), sizeof(long long) * tab.size(), cudaMemcpyHostToDevice, nc->stream));
    for (int q0 = 0; q0 < 2 * n_nodes; q0 += 32768) {              // grid.y limit: children in slices
        const int nq = std::min(32768, 2 * n_nodes - q0);
        k_pool_gather<<<dim3((unsigned)(((long long)max_len * d + 255) / 256), (unsigned)nq), 256, 0, nc->stream>>>(
            c->lv_tab.p + (size_t)q0 * 3, P.t.p, P.c.p, P.w.p, d, nc->tensors.p, nc->coords.p, c->nd_w.p);
    }
    CU(cudaGetLastError());
    if ((rc = chains_finish(nc, 2 * n_nodes, d))) return rc;          // synchronises nc->stream: the gathered weights are in place too
    const double t_chains = now_ms();
    const LevelParams lp{gamma_tensor, gamma_coords, gamma_weight, gap_open, gap_extend};
    double pair_ms = 0, level_ms = 0;
    if ((rc = level_core(c, nc, n_nodes, d, offsets.data(), mult, lp, out_off.data(), P.t.p, P.c.p, P.w.p, aln1, aln2, aln_len, score, status,
                         &pair_ms, &level_ms)))
        return rc;
    if (first_new_id) *first_new_id = n_seq;
    for (int k = 0; k < n_nodes; ++k) {
        P.off.push_back(out_off[(size_t)k]);
        P.len.push_back(aln_len[k]);
        aln_off[k] = offsets[(size_t)2 * k];
        P.ch1.push_back(child1[k]); P.ch2.push_back(child2[k]);
        P.al1.emplace_back(aln1 + aln_off[k], aln1 + aln_off[k] + aln_len[k]);
        P.al2.emplace_back(aln2 + aln_off[k], aln2 + aln_off[k] + aln_len[k]);
    }
    aln_off[n_nodes] = total;
    P.used = need;
    if (level_timeline())
        fprintf(stderr, "[msa level] nodes %5d residues %8lld: gather + tables %7.3f  pair run + level kernels + alignments D2H %7.3f (device %7.3f + %7.3f) ms\n",
                n_nodes, total, t_chains - t_begin, now_ms() - t_chains, pair_ms, level_ms);
    return 0;
}

int crt_msa_lengths(crt_ctx *c, int32_t *n_sequences, int32_t *lengths, int32_t cap)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (!c->pool.active) return fail(CRT_E_STATE, "crt_msa_begin has not been called");
    const int n = (int)c->pool.len.size();
    if (n_sequences) *n_sequences = n;
    if (lengths) {
        if (cap < n) return fail(CRT_E_ARG, "lengths holds %d entries, %d needed", cap, n);
        std::memcpy(lengths, c->pool.len.data(), sizeof(int32_t) * (size_t)n);
    }
    return 0;
}

/* the listed sequences of the pool, packed one after the other (rows = sum of their lengths): one gather on the device into a
 * staging buffer, then three copies */
int crt_msa_fetch(crt_ctx *c, const int32_t *ids, int32_t count, double *tensors, double *coords, double *weights)
{
    if (!c || !ids || !tensors || !coords || !weights) return fail(CRT_E_ARG, "null argument");
    crt_ctx::MsaPool &P = c->pool;
    if (!P.active) return fail(CRT_E_STATE, "crt_msa_begin has not been called");
    if (count <= 0) return 0;
    std::vector<long long> tab((size_t)count * 3);
    long long rows = 0;
    int max_len = 1;
    for (int q = 0; q < count; ++q) {
        if (ids[q] < 0 || ids[q] >= (int)P.len.size()) return fail(CRT_E_ARG, "sequence %d is not in the pool", ids[q]);
        tab[(size_t)q * 3] = P.off[(size_t)ids[q]]; tab[(size_t)q * 3 + 1] = rows; tab[(size_t)q * 3 + 2] = P.len[(size_t)ids[q]];
        rows += P.len[(size_t)ids[q]];
        max_len = std::max(max_len, P.len[(size_t)ids[q]]);
    }
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = c->nd_t.ensure((size_t)rows * P.d)) || (rc = c->nd_c.ensure((size_t)rows * 3)) || (rc = c->nd_wm.ensure((size_t)rows)) ||
        (rc = c->lv_tab.ensure(tab.size())))
        return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->lv_tab.p, tab.data(), sizeof(long long) * tab.size(), cudaMemcpyHostToDevice, st));
    const int gy = 32768;                      // grid.y limit 65535: sequences in slices
    for (int q0 = 0; q0 < count; q0 += gy) {
        const int nq = std::min(gy, count - q0);
        k_pool_gather<<<dim3((unsigned)(((long long)max_len * P.d + 255) / 256), (unsigned)nq), 256, 0, st>>>(c->lv_tab.p + (size_t)q0 * 3, P.t.p, P.c.p,
                                                                                                             P.w.p, P.d, c->nd_t.p, c->nd_c.p, c->nd_wm.p);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(tensors, c->nd_t.p, sizeof(double) * (size_t)rows * P.d, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(coords, c->nd_c.p, sizeof(double) * (size_t)rows * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(weights, c->nd_wm.p, sizeof(double) * (size_t)rows, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

/* Index arrays of the sequences under node `root` in the frame of that node (the bookkeeping of multiple_alignment.py:219-232,
 * composed top-down: the map columns of the frame -> columns of a child is one gather through the node's alignment per edge, and at
 * a leaf the map is the index array).  Host-side integer bookkeeping on the alignments crt_msa_level returned; no device work.
 * Rows of `out` [n_under][A] (A = length of `root`) in the reference's dictionary order (first child's sequences first). */
int crt_msa_compose(crt_ctx *c, int32_t root, int32_t *leaf_ids, int32_t leaf_cap, int32_t *n_under, int64_t *out, int64_t out_cap)
{
    if (!c || !leaf_ids || !n_under || !out) return fail(CRT_E_ARG, "null argument");
    crt_ctx::MsaPool &P = c->pool;
    if (!P.active) return fail(CRT_E_STATE, "crt_msa_begin has not been called");
    const int n_seq = (int)P.len.size(), nl = P.n_leaves;
    if (root < 0 || root >= n_seq) return fail(CRT_E_ARG, "sequence %d is not in the pool", root);
    const long long A = P.len[(size_t)root];
    struct Item { int node; std::vector<int32_t> map; };
    std::vector<Item> stack;
    {
        Item it{root, std::vector<int32_t>((size_t)A)};
        for (long long x = 0; x < A; ++x) it.map[(size_t)x] = (int32_t)x;
        stack.push_back(std::move(it));
    }
    int count = 0;
    while (!stack.empty()) {
        Item it = std::move(stack.back());
        stack.pop_back();
        if (it.node < nl) {
            if (count >= leaf_cap || (long long)(count + 1) * A > out_cap) return fail(CRT_E_ARG, "output holds %d rows, more needed", count);
            leaf_ids[count] = it.node;
            int64_t *o = out + (long long)count * A;
            for (long long x = 0; x < A; ++x) o[x] = it.map[(size_t)x];
            ++count;
            continue;
        }
        const size_t q = (size_t)(it.node - nl);
        const std::vector<int32_t> *al[2] = {&P.al1[q], &P.al2[q]};
        const int ch[2] = {P.ch1[q], P.ch2[q]};
        for (int s = 1; s >= 0; --s) {                     // second child first onto the stack: the first child's subtree comes out first
            Item nx{ch[s], std::vector<int32_t>((size_t)A)};
            const int32_t *a = al[s]->data();
            for (long long x = 0; x < A; ++x) {
                const int32_t v = it.map[(size_t)x];
                nx.map[(size_t)x] = v < 0 ? -1 : a[v];
            }
            stack.push_back(std::move(nx));
        }
    }
    *n_under = count;
    return 0;
}

int crt_msa_end(crt_ctx *c)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    c->pool.active = false;
    c->pool.ch1.clear(); c->pool.ch2.clear(); c->pool.al1.clear(); c->pool.al2.clear();
    c->pool.t.release(); c->pool.c.release(); c->pool.w.release();
    c->pool.off.clear(); c->pool.len.clear(); c->pool.used = 0;
    return 0;
}

}  // extern "C"
