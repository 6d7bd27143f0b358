// crt_multi.inl -- multi-GPU inside the C ABI (included by crt_api.cu).
//
// The reference's seam is ONE process calling ONE method (multiple_alignment.py:498-500), so a drop-in that wants every GPU of
// the box cannot ask the caller for torchrun.  crt_multi owns one crt_ctx per device and a NCCL communicator over them
// (ncclCommInitAll, single process); a call
//   * uploads the packed chains to every device from the caller's buffers (one host thread per device),
//   * runs crt_pairwise_shard(rank = device index, world = devices) on every device at once (cost-sharded units, no
//     data-path collective),
//   * packs score | rmsd | tm of the shard into one vector (float32 in the production mode, float64 in the parity mode, so
//     that the float64 results stay bit-identical to the one-GPU run),
//   * does the ONE exchange of the path -- a grouped ncclAllGather of those vectors over NVLink --
//   * and scatters the gathered vectors into the dense symmetric [N,N] float64 matrices on device 0 (k_scatter_ranks), from
//     where they are copied to the caller's arrays once.
// The same packing / scatter entry points (crt_pack_results, crt_scatter_gathered) serve the one-process-per-GPU layout
// (torchrun + torch.distributed in caretta_b200/distributed.py): there the all-gather is torch's, the scatter runs on rank 0 only.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already mapped by torch when there is one, else the system's),
// so the library itself links only libcudart; with one device no NCCL is needed at all.
#include <dlfcn.h>
#include <nccl.h>
#include <thread>

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.handle) return 0;
    const char *names[] = {getenv("CARETTA_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(CRT_E_STATE, "NCCL not found (dlopen libnccl.so.2: %s); set CARETTA_B200_NCCL to its path", dlerror());
    NcclApi a;
    a.handle = h;
#define CRT_SYM(field, name)                                                                 \
    *reinterpret_cast<void **>(&a.field) = dlsym(h, name);                                   \
    if (!a.field) return fail(CRT_E_STATE, "NCCL symbol %s missing", name);
    CRT_SYM(CommInitAll, "ncclCommInitAll") CRT_SYM(CommDestroy, "ncclCommDestroy") CRT_SYM(AllGather, "ncclAllGather")
    CRT_SYM(GroupStart, "ncclGroupStart") CRT_SYM(GroupEnd, "ncclGroupEnd") CRT_SYM(GetErrorString, "ncclGetErrorString")
    CRT_SYM(GetVersion, "ncclGetVersion")
#undef CRT_SYM
    g_nccl = a;
    return 0;
}

#define NC(call)                                                                                              \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess) return fail(CRT_E_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));      \
    } while (0)

template <typename T>
__global__ void k_pack(const double *s, const double *r, const double *t, T *dst, long long np, long long pad)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= pad) return;
    const bool in = q < np;
    dst[q] = in ? (T)s[q] : T(0);
    dst[pad + q] = in ? (T)r[q] : T(0);
    dst[2 * pad + q] = in ? (T)t[q] : T(0);
}

// gathered: [world][3][pad]; pi / pj: the pairs of rank r at [rank_off[r], rank_off[r + 1]) in shard order.
// out[i][j] = out[j][i] = value (multiple_alignment.py:164); TM diagonal 1 (:1019-1024)
template <typename T>
__global__ void k_scatter_ranks(const T *gathered, const int *pi, const int *pj, const long long *rank_off, int world, long long pad,
                                double *S, double *R, double *Tm, int N)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (Tm && q < N) Tm[q * N + q] = 1.0;
    const long long total = rank_off[world];
    if (q >= total) return;
    int r = 0;
    while (r + 1 < world && rank_off[r + 1] <= q) ++r;
    const long long k = q - rank_off[r];
    const T *g = gathered + (long long)r * 3 * pad;
    const long long i = pi[q], j = pj[q];
    const double s = (double)g[k];
    S[i * N + j] = S[j * N + i] = s;
    if (R) { const double v = (double)g[pad + k]; R[i * N + j] = R[j * N + i] = v; }
    if (Tm) { const double v = (double)g[2 * pad + k]; Tm[i * N + j] = Tm[j * N + i] = v; }
}

}  // namespace

struct crt_multi {
    int world = 0;
    std::vector<int> dev;
    std::vector<crt_ctx *> ctx;
    std::vector<ncclComm_t> comm;
    std::vector<void *> send, recv;          // per device: [3 * pad] and [world * 3 * pad] elements of the run's dtype
    size_t send_bytes = 0, recv_bytes = 0;
    double elapsed_ms = 0, gather_ms = 0, wall_ms = 0;
    long long rerun_pairs = 0;
};

extern "C" {

/* score | rmsd | tm of the last crt_pairwise_shard packed into d_dst [3 * pad] (device address) as float32 (is_f64 = 0) or
 * float64 (is_f64 = 1: exact copies); entries beyond the shard's pair count are 0.  On the context's stream, synchronous. */
int crt_pack_results(crt_ctx *c, void *d_dst, int64_t pad, int32_t is_f64)
{
    if (!c || !d_dst) return fail(CRT_E_ARG, "null argument");
    if (pad < c->run_pairs) return fail(CRT_E_ARG, "pad %lld < %lld pairs of the run", (long long)pad, c->run_pairs);
    if (pad == 0) return 0;
    CU(cudaSetDevice(c->device));
    const unsigned grid = (unsigned)((pad + 255) / 256);
    if (is_f64) k_pack<double><<<grid, 256, 0, c->stream>>>(c->score.p, c->rmsd.p, c->tm.p, (double *)d_dst, c->run_pairs, pad);
    else k_pack<float><<<grid, 256, 0, c->stream>>>(c->score.p, c->rmsd.p, c->tm.p, (float *)d_dst, c->run_pairs, pad);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

/* d_gathered: [world][3][pad] (device address on the context's device; what an all-gather of crt_pack_results vectors gives).
 * Scatters every rank's shard into the dense symmetric matrices on the device and copies them to the host arrays (out_rmsd /
 * out_tm may be NULL).  The context must hold the chain set the shards were planned for. */
int crt_scatter_gathered(crt_ctx *c, const void *d_gathered, int32_t world, int64_t pad, int32_t is_f64, double *out_score,
                         double *out_rmsd, double *out_tm)
{
    if (!c || !d_gathered || !out_score) return fail(CRT_E_ARG, "null argument");
    if (c->N <= 0) return fail(CRT_E_STATE, "no chains");
    if (world < 1 || pad < 0) return fail(CRT_E_ARG, "bad world / pad");
    CU(cudaSetDevice(c->device));
    int rc;
    // the layout of all ranks (cached per chain set and world)
    if (c->lay_hash != c->offsets_hash || c->lay_world != world) {
        std::vector<int> pi, pj;
        std::vector<long long> off((size_t)world + 1, 0);
        for (int r = 0; r < world; ++r) {
            std::vector<HostUnit> units;
            build_all_units(c, CRT_FP32, world, units);
            shard_units(units, r, world);
            std::vector<int> a, b;
            long long np = 0;
            assign_pairs(units, &a, &b, &np, nullptr, c);
            if (np > pad) return fail(CRT_E_ARG, "rank %d has %lld pairs > pad %lld", r, np, (long long)pad);
            pi.insert(pi.end(), a.begin(), a.end());
            pj.insert(pj.end(), b.begin(), b.end());
            off[(size_t)r + 1] = off[(size_t)r] + np;
        }
        if ((rc = c->lay_pi.ensure(pi.size() + 1))) return rc;
        if ((rc = c->lay_pj.ensure(pj.size() + 1))) return rc;
        if ((rc = c->lay_off.ensure(off.size()))) return rc;
        CU(cudaMemcpyAsync(c->lay_pi.p, pi.data(), sizeof(int) * pi.size(), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->lay_pj.p, pj.data(), sizeof(int) * pj.size(), cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->lay_off.p, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->lay_hash = c->offsets_hash; c->lay_world = world; c->lay_total = off[(size_t)world];
    }
    const size_t N = (size_t)c->N;
    const int nm = 1 + (out_rmsd ? 1 : 0) + (out_tm ? 1 : 0);
    if ((rc = c->dense.ensure(N * N * (size_t)nm))) return rc;
    double *dS = c->dense.p, *dR = out_rmsd ? dS + N * N : nullptr, *dT = out_tm ? dS + N * N * (size_t)(out_rmsd ? 2 : 1) : nullptr;
    CU(cudaMemsetAsync(c->dense.p, 0, sizeof(double) * N * N * (size_t)nm, c->stream));
    const long long work = std::max<long long>(c->lay_total, (long long)N);
    const unsigned grid = (unsigned)((work + 255) / 256);
    if (is_f64) k_scatter_ranks<double><<<grid, 256, 0, c->stream>>>((const double *)d_gathered, c->lay_pi.p, c->lay_pj.p, c->lay_off.p, world, pad, dS, dR, dT, (int)N);
    else k_scatter_ranks<float><<<grid, 256, 0, c->stream>>>((const float *)d_gathered, c->lay_pi.p, c->lay_pj.p, c->lay_off.p, world, pad, dS, dR, dT, (int)N);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_score, dS, sizeof(double) * N * N, cudaMemcpyDeviceToHost, c->stream));
    if (out_rmsd) CU(cudaMemcpyAsync(out_rmsd, dR, sizeof(double) * N * N, cudaMemcpyDeviceToHost, c->stream));
    if (out_tm) CU(cudaMemcpyAsync(out_tm, dT, sizeof(double) * N * N, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int crt_multi_destroy(crt_multi *m);

/* One context per device and a NCCL communicator over them.  n_dev <= 0: every visible device (dev_ids ignored). */
int crt_multi_create(int32_t n_dev, const int32_t *dev_ids, crt_multi **out)
{
    if (!out) return fail(CRT_E_ARG, "null out pointer");
    *out = nullptr;
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0)
        return fail(CRT_E_CUDA, "no CUDA device: %s (this engine has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    crt_multi *m = new crt_multi();
    if (n_dev <= 0) { for (int d = 0; d < have; ++d) m->dev.push_back(d); }
    else {
        if (!dev_ids) { delete m; return fail(CRT_E_ARG, "null dev_ids"); }
        for (int k = 0; k < n_dev; ++k) {
            if (dev_ids[k] < 0 || dev_ids[k] >= have) { delete m; return fail(CRT_E_ARG, "device %d out of range (%d devices)", dev_ids[k], have); }
            for (int q = 0; q < k; ++q) if (dev_ids[q] == dev_ids[k]) { delete m; return fail(CRT_E_ARG, "device %d listed twice", dev_ids[k]); }
            m->dev.push_back(dev_ids[k]);
        }
    }
    m->world = (int)m->dev.size();
    m->ctx.assign((size_t)m->world, nullptr);
    m->send.assign((size_t)m->world, nullptr);
    m->recv.assign((size_t)m->world, nullptr);
    for (int r = 0; r < m->world; ++r) {
        int rc = crt_create(m->dev[(size_t)r], &m->ctx[(size_t)r]);
        if (rc) { crt_multi_destroy(m); return rc; }
    }
    if (m->world > 1) {
        int rc = load_nccl();
        if (rc) { crt_multi_destroy(m); return rc; }
        m->comm.assign((size_t)m->world, nullptr);
        ncclResult_t r_ = g_nccl.CommInitAll(m->comm.data(), m->world, m->dev.data());
        if (r_ != ncclSuccess) {
            m->comm.clear();
            crt_multi_destroy(m);
            return fail(CRT_E_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r_));
        }
    }
    *out = m;
    return 0;
}

int crt_multi_destroy(crt_multi *m)
{
    if (!m) return 0;
    for (size_t r = 0; r < m->comm.size(); ++r)
        if (m->comm[r]) g_nccl.CommDestroy(m->comm[r]);
    for (int r = 0; r < m->world; ++r) {
        if (!m->ctx[(size_t)r]) continue;
        cudaSetDevice(m->dev[(size_t)r]);
        if (m->send[(size_t)r]) cudaFree(m->send[(size_t)r]);
        if (m->recv[(size_t)r]) cudaFree(m->recv[(size_t)r]);
        crt_destroy(m->ctx[(size_t)r]);
    }
    delete m;
    return 0;
}

int32_t crt_multi_devices(crt_multi *m) { return m ? m->world : 0; }

/* crt_set_chains on every device, from the caller's buffers (one host thread per device). */
int crt_multi_set_chains(crt_multi *m, const double *coords, const double *tensors, const int64_t *offsets, int32_t n_chains, int32_t d)
{
    if (!m) return fail(CRT_E_ARG, "null context");
    std::vector<int> rcs((size_t)m->world, 0);
    std::vector<std::string> errs((size_t)m->world);
    std::vector<std::thread> th;
    for (int r = 0; r < m->world; ++r)
        th.emplace_back([&, r] {
            rcs[(size_t)r] = crt_set_chains(m->ctx[(size_t)r], coords, tensors, offsets, n_chains, d);
            if (rcs[(size_t)r]) errs[(size_t)r] = g_err;
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < m->world; ++r)
        if (rcs[(size_t)r]) return fail(rcs[(size_t)r], "device %d: %s", m->dev[(size_t)r], errs[(size_t)r].c_str());
    return 0;
}

/* make_pairwise_matrix over every device of the set: shards by cost, one grouped ncclAllGather, dense scatter on the first
 * device, one copy to the caller's float64 [N,N] arrays (out_rmsd / out_tm may be NULL).  Bitwise the one-GPU result. */
int crt_multi_pairwise_all(crt_multi *m, const crt_params *prm, double *out_score, double *out_rmsd, double *out_tm)
{
    if (!m || !prm || !out_score) return fail(CRT_E_ARG, "null argument");
    const int W = m->world;
    const auto t0 = std::chrono::steady_clock::now();
    if (W == 1) {
        int rc = crt_pairwise_all(m->ctx[0], prm, out_score, out_rmsd, out_tm);
        m->elapsed_ms = m->ctx[0]->elapsed_ms; m->gather_ms = 0; m->rerun_pairs = m->ctx[0]->rerun_pairs;
        m->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return rc;
    }
    int rc = check_params(m->ctx[0], prm);
    if (rc) return rc;
    const bool f64 = prm->precision == CRT_FP64;
    const size_t esz = f64 ? 8 : 4;
    long long pad = 1;
    for (int r = 0; r < W; ++r) pad = std::max<long long>(pad, crt_plan_shard_size(reinterpret_cast<const int64_t *>(m->ctx[0]->offsets.data()), m->ctx[0]->N, r, W));
    const size_t sb = (size_t)3 * (size_t)pad * esz, rb = sb * (size_t)W;
    for (int r = 0; r < W; ++r) {
        CU(cudaSetDevice(m->dev[(size_t)r]));
        if (sb > m->send_bytes || !m->send[(size_t)r]) { if (m->send[(size_t)r]) cudaFree(m->send[(size_t)r]); m->send[(size_t)r] = nullptr; CU(cudaMalloc(&m->send[(size_t)r], sb + sb / 8)); }
        if (rb > m->recv_bytes || !m->recv[(size_t)r]) { if (m->recv[(size_t)r]) cudaFree(m->recv[(size_t)r]); m->recv[(size_t)r] = nullptr; CU(cudaMalloc(&m->recv[(size_t)r], rb + rb / 8)); }
    }
    if (sb > m->send_bytes) m->send_bytes = sb + sb / 8;
    if (rb > m->recv_bytes) m->recv_bytes = rb + rb / 8;

    std::vector<int> rcs((size_t)W, 0);
    std::vector<std::string> errs((size_t)W);
    std::vector<std::thread> th;
    for (int r = 0; r < W; ++r)
        th.emplace_back([&, r] {
            int q = crt_pairwise_shard(m->ctx[(size_t)r], prm, r, W);
            if (!q) q = crt_pack_results(m->ctx[(size_t)r], m->send[(size_t)r], pad, f64 ? 1 : 0);
            rcs[(size_t)r] = q;
            if (q) errs[(size_t)r] = g_err;
        });
    for (auto &t : th) t.join();
    m->elapsed_ms = 0; m->rerun_pairs = 0;
    for (int r = 0; r < W; ++r) {
        if (rcs[(size_t)r]) return fail(rcs[(size_t)r], "device %d: %s", m->dev[(size_t)r], errs[(size_t)r].c_str());
        m->elapsed_ms = std::max(m->elapsed_ms, m->ctx[(size_t)r]->elapsed_ms);
        m->rerun_pairs += m->ctx[(size_t)r]->rerun_pairs;
    }
    // the one exchange step of the path
    crt_ctx *c0 = m->ctx[0];
    CU(cudaSetDevice(m->dev[0]));
    CU(cudaEventRecord(c0->ev2, c0->stream));
    NC(g_nccl.GroupStart());
    for (int r = 0; r < W; ++r)
        NC(g_nccl.AllGather(m->send[(size_t)r], m->recv[(size_t)r], (size_t)3 * (size_t)pad, f64 ? ncclFloat64 : ncclFloat32, m->comm[(size_t)r],
                            m->ctx[(size_t)r]->stream));
    NC(g_nccl.GroupEnd());
    CU(cudaSetDevice(m->dev[0]));
    CU(cudaEventRecord(c0->ev3, c0->stream));
    rc = crt_scatter_gathered(c0, m->recv[0], W, pad, f64 ? 1 : 0, out_score, out_rmsd, out_tm);
    if (rc) return rc;
    for (int r = 1; r < W; ++r) { CU(cudaSetDevice(m->dev[(size_t)r])); CU(cudaStreamSynchronize(m->ctx[(size_t)r]->stream)); }
    CU(cudaSetDevice(m->dev[0]));
    float gms = 0;
    CU(cudaEventElapsedTime(&gms, c0->ev2, c0->ev3));
    m->gather_ms = gms;
    m->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

/* out3 = {max over devices of the shard's device time, device time of the all-gather on the first device, host wall time of the
 * whole call}, all in ms; *rerun_pairs = pairs the tie detection sent through the float64 kernels (all devices). */
int crt_multi_last_timing(crt_multi *m, double *out3, int64_t *rerun_pairs)
{
    if (!m || !out3) return fail(CRT_E_ARG, "null argument");
    out3[0] = m->elapsed_ms; out3[1] = m->gather_ms; out3[2] = m->wall_ms;
    if (rerun_pairs) *rerun_pairs = m->rerun_pairs;
    return 0;
}

/* The per-device context (e.g. for crt_last_* queries, or to run the consumers of the matrix on that device). */
crt_ctx *crt_multi_ctx(crt_multi *m, int32_t index)
{
    if (!m || index < 0 || index >= m->world) { fail(CRT_E_ARG, "bad device index"); return nullptr; }
    return m->ctx[(size_t)index];
}

}  // extern "C"
