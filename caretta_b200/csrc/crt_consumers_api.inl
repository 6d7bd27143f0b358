// crt_consumers_api.inl -- host side of the alignment consumers (included at the end of crt_api.cu; kernels in
// crt_consumers.cuh).  SURVEY section 8f ranks 3-4: superposition consumers and the text writers.

namespace {

struct AlnDev {
    long long *aln = nullptr;
    unsigned *bitsT = nullptr;
    int *present = nullptr, *bad = nullptr;
    int W = 0;
};

// upload the alignment, build the transposed presence masks; offsets = device chain offsets for the range check (or null)
int upload_alignment(crt_ctx *c, Scratch &sc, const int64_t *aln, int N, int64_t A, const long long *d_offsets, AlnDev &o,
                     std::vector<int> &present)
{
    o.W = (int)((A + 31) / 32);
    CU(sc.alloc(&o.aln, (size_t)N * A));
    CU(sc.alloc(&o.bitsT, (size_t)N * o.W));
    CU(sc.alloc(&o.present, (size_t)N));
    CU(sc.alloc(&o.bad, 1));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(o.aln, aln, sizeof(long long) * (size_t)N * A, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(o.present, 0, sizeof(int) * (size_t)N, st));
    CU(cudaMemsetAsync(o.bad, 0, sizeof(int), st));
    const long long warps = (long long)N * o.W;
    k_aln_bits<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(o.aln, N, A, o.W, d_offsets, o.bitsT, o.present, o.bad);
    CU(cudaGetLastError());
    present.resize((size_t)N);
    int bad = 0;
    CU(cudaMemcpyAsync(present.data(), o.present, sizeof(int) * (size_t)N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&bad, o.bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (bad) return fail(CRT_E_ARG, "alignment holds an index < -1 or beyond the end of its chain");
    return 0;
}

// sequential batches of independent (reference, member) superpositions; coordinates end up in d_out
int superpose_batches(crt_ctx *c, Scratch &sc, const AlnDev &ad, int64_t A, const int32_t *ref, const int32_t *mem, int64_t n_pairs,
                      const int64_t *batch_off, int n_batches, const int *d_cols, int n_cols, int center_ref, int min_common,
                      bool in_place, double *d_out, double *out_rot, double *out_tran, int32_t *out_ncommon)
{
    cudaStream_t st = c->stream;
    int *d_ref = nullptr, *d_mem = nullptr, *d_nc = nullptr;
    double *d_rot = nullptr, *d_tran = nullptr;
    CU(sc.alloc(&d_ref, (size_t)n_pairs));
    CU(sc.alloc(&d_mem, (size_t)n_pairs));
    CU(sc.alloc(&d_nc, (size_t)n_pairs));
    CU(sc.alloc(&d_rot, (size_t)n_pairs * 9));
    CU(sc.alloc(&d_tran, (size_t)n_pairs * 3));
    if (n_pairs > 0) {
        CU(cudaMemcpyAsync(d_ref, ref, sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_mem, mem, sizeof(int) * (size_t)n_pairs, cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(d_nc, 0, sizeof(int) * (size_t)n_pairs, st));
        CU(cudaMemsetAsync(d_rot, 0, sizeof(double) * (size_t)n_pairs * 9, st));
        CU(cudaMemsetAsync(d_tran, 0, sizeof(double) * (size_t)n_pairs * 3, st));
    }
    CU(cudaMemcpyAsync(d_out, c->coords.p, sizeof(double) * (size_t)c->total * 3, cudaMemcpyDeviceToDevice, st));
    CU(cudaEventRecord(c->ev0, st));
    c->launches = 0;
    for (int b = 0; b < n_batches; ++b) {
        const int64_t lo = batch_off[b], n = batch_off[b + 1] - lo;
        if (n <= 0) continue;
        k_superpose<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(in_place ? d_out : c->coords.p, d_out, c->d_offsets.p, ad.aln, A,
                                                                       d_ref + lo, d_mem + lo, (int)n, d_cols, n_cols, center_ref,
                                                                       min_common, d_rot + lo * 9, d_tran + lo * 3, d_nc + lo);
        ++c->launches;
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, st));
    if (n_pairs > 0) {
        if (out_rot) CU(cudaMemcpyAsync(out_rot, d_rot, sizeof(double) * (size_t)n_pairs * 9, cudaMemcpyDeviceToHost, st));
        if (out_tran) CU(cudaMemcpyAsync(out_tran, d_tran, sizeof(double) * (size_t)n_pairs * 3, cudaMemcpyDeviceToHost, st));
        if (out_ncommon) CU(cudaMemcpyAsync(out_ncommon, d_nc, sizeof(int) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

int finish_timed(crt_ctx *c, const char *what)
{
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(CRT_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->elapsed_ms = ms;
    return 0;
}

}  // namespace

extern "C" {

/* make_coverage_gap_distance_matrix, multiple_alignment.py:45-56 */
int crt_coverage_gap_matrix(crt_ctx *c, const int64_t *aln, int32_t N, int64_t A, double *distance, int32_t *aligning)
{
    if (!c || !aln || !distance || !aligning) return fail(CRT_E_ARG, "null argument");
    if (N <= 0 || A <= 0) return fail(CRT_E_ARG, "empty alignment (%d x %lld)", N, (long long)A);
    CU(cudaSetDevice(c->device));
    Scratch sc(c);
    AlnDev ad;
    std::vector<int> present;
    int rc = upload_alignment(c, sc, aln, N, A, nullptr, ad, present);
    if (rc) return rc;
    for (int p = 0; p < N; ++p)
        if (present[(size_t)p] == 0)
            return fail(CRT_E_ARG, "protein %d has no residue in the alignment (the reference divides by zero, multiple_alignment.py:54)", p);
    double *d_dist = nullptr;
    int *d_al = nullptr;
    const size_t NN = (size_t)N * N;
    CU(sc.alloc(&d_dist, NN));
    CU(sc.alloc(&d_al, NN));
    cudaStream_t st = c->stream;
    CU(cudaEventRecord(c->ev0, st));
    k_coverage_gap<<<dim3((unsigned)((N + 255) / 256), (unsigned)N), 256, sizeof(unsigned) * (size_t)ad.W, st>>>(ad.bitsT, ad.present, N, ad.W,
                                                                                                                 d_dist, d_al);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, st));
    c->launches = 1;
    CU(cudaMemcpyAsync(distance, d_dist, sizeof(double) * NN, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(aligning, d_al, sizeof(int) * NN, cudaMemcpyDeviceToHost, st));
    return finish_timed(c, "crt_coverage_gap_matrix");
}

/* superpose (multiple_alignment.py:854-867) = superpose_core (:869-905) or superpose_reference (:908-927) on the chains of the
 * context. */
int crt_superpose(crt_ctx *c, const int64_t *aln, int64_t A, int32_t mode, int32_t reference, const int64_t *core_columns,
                  int64_t n_core_columns, double *out_coords, double *out_rot, double *out_tran, int32_t *out_ncommon,
                  int32_t *out_mode, int32_t *out_reference, int64_t *out_ncore)
{
    if (!c || !aln || !out_coords) return fail(CRT_E_ARG, "null argument");
    if (c->N <= 0) return fail(CRT_E_STATE, "crt_set_chains has not been called");
    if (A <= 0) return fail(CRT_E_ARG, "alignment length must be > 0");
    if (mode < CRT_SUP_AUTO || mode > CRT_SUP_REFERENCE) return fail(CRT_E_ARG, "unknown mode %d", mode);
    const int N = c->N;
    if (reference >= N) return fail(CRT_E_ARG, "reference %d out of range", reference);
    CU(cudaSetDevice(c->device));
    Scratch sc(c);
    AlnDev ad;
    std::vector<int> present;
    int rc = upload_alignment(c, sc, aln, N, A, c->d_offsets.p, ad, present);
    if (rc) return rc;
    // reference_name = sorted(names, key = residues in the alignment, reverse=True)[0]: the FIRST protein with the maximum (:855)
    int r = reference;
    if (r < 0) {
        r = 0;
        for (int p = 1; p < N; ++p)
            if (present[(size_t)p] > present[(size_t)r]) r = p;
    }
    // core columns (:856-862)
    cudaStream_t st = c->stream;
    unsigned *d_core = nullptr;
    CU(sc.alloc(&d_core, (size_t)ad.W));
    k_core_mask<<<ad.W, 256, 0, st>>>(ad.bitsT, N, ad.W, d_core);
    CU(cudaGetLastError());
    std::vector<unsigned> core((size_t)ad.W);
    CU(cudaMemcpyAsync(core.data(), d_core, sizeof(unsigned) * (size_t)ad.W, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<int> cols;
    for (int64_t q = 0; q < A; ++q)
        if (core[(size_t)(q >> 5)] >> (q & 31) & 1u) cols.push_back((int)q);
    if (core_columns) {
        if (mode != CRT_SUP_CORE) return fail(CRT_E_ARG, "core_columns needs mode CRT_SUP_CORE");
        std::vector<int> given;
        for (int64_t k = 0; k < n_core_columns; ++k) {
            const int64_t q = core_columns[k];
            if (q < 0 || q >= A) return fail(CRT_E_ARG, "core column %lld out of range", (long long)q);
            if (!(core[(size_t)(q >> 5)] >> (q & 31) & 1u)) return fail(CRT_E_ARG, "core column %lld holds a gap", (long long)q);
            given.push_back((int)q);
        }
        cols.swap(given);
    }
    const int64_t n_core = (int64_t)cols.size();
    int m = mode;
    if (m == CRT_SUP_AUTO) m = n_core < A / 2 ? CRT_SUP_REFERENCE : CRT_SUP_CORE;          // :864-867
    if (out_mode) *out_mode = m;
    if (out_reference) *out_reference = r;
    if (out_ncore) *out_ncore = n_core;
    if (m == CRT_SUP_CORE && n_core == 0) return fail(CRT_E_ARG, "no core column: every alignment column has a gap");
    std::vector<int32_t> ref((size_t)N, r), mem((size_t)N);
    for (int p = 0; p < N; ++p) mem[(size_t)p] = p;
    double *d_out = nullptr;
    CU(sc.alloc(&d_out, (size_t)c->total * 3));
    if (m == CRT_SUP_CORE) {
        int *d_cols = nullptr;
        CU(sc.alloc(&d_cols, cols.size()));
        CU(cudaMemcpyAsync(d_cols, cols.data(), sizeof(int) * cols.size(), cudaMemcpyHostToDevice, st));
        const int64_t boff[2] = {0, N};
        rc = superpose_batches(c, sc, ad, A, ref.data(), mem.data(), N, boff, 1, d_cols, (int)cols.size(), 1, 1, false, d_out, out_rot,
                               out_tran, out_ncommon);
    } else {
        // the loop of superpose_reference includes the reference itself, whose coordinates are replaced (by themselves up to
        // rounding) before the later members are superposed onto it: three dependent launches keep that order
        const int64_t boff[4] = {0, r, r + 1, N};
        rc = superpose_batches(c, sc, ad, A, ref.data(), mem.data(), N, boff, 3, nullptr, 0, 0, 4, true, d_out, out_rot, out_tran,
                               out_ncommon);
    }
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_coords, d_out, sizeof(double) * (size_t)c->total * 3, cudaMemcpyDeviceToHost, st));
    return finish_timed(c, "crt_superpose");
}

/* superpose_references (multiple_alignment.py:930-950) and the loops of write_superposed_pdbs_reference(s) (:684-737, :787-850) */
int crt_superpose_pairs(crt_ctx *c, const int64_t *aln, int64_t A, const int32_t *ref, const int32_t *mem, int64_t n_pairs,
                        const int64_t *batch_off, int32_t n_batches, double *out_coords, double *out_rot, double *out_tran,
                        int32_t *out_ncommon)
{
    if (!c || !aln || !out_coords || (n_pairs > 0 && (!ref || !mem)) || !batch_off) return fail(CRT_E_ARG, "null argument");
    if (c->N <= 0) return fail(CRT_E_STATE, "crt_set_chains has not been called");
    if (A <= 0 || n_pairs < 0 || n_batches < 0) return fail(CRT_E_ARG, "bad size");
    const int N = c->N;
    if (batch_off[0] != 0 || batch_off[n_batches] != n_pairs) return fail(CRT_E_ARG, "batch_off must run from 0 to n_pairs");
    // pairs of one batch run concurrently and in place: a member must not be the reference (or the member) of another pair there
    std::vector<int> stamp_ref((size_t)N, -1), stamp_mem((size_t)N, -1);
    for (int b = 0; b < n_batches; ++b) {
        const int64_t lo = batch_off[b], hi = batch_off[b + 1];
        if (hi < lo) return fail(CRT_E_ARG, "batch_off must be non-decreasing");
        for (int64_t q = lo; q < hi; ++q) {
            if (ref[q] < 0 || ref[q] >= N || mem[q] < 0 || mem[q] >= N) return fail(CRT_E_ARG, "pair %lld out of range", (long long)q);
            stamp_ref[(size_t)ref[q]] = b;
        }
        for (int64_t q = lo; q < hi; ++q) {
            if (stamp_mem[(size_t)mem[q]] == b) return fail(CRT_E_ARG, "protein %d is a member twice in batch %d", mem[q], b);
            stamp_mem[(size_t)mem[q]] = b;
            if (stamp_ref[(size_t)mem[q]] == b && hi - lo > 1)
                return fail(CRT_E_ARG, "protein %d is both a member and a reference in batch %d (dependent pairs need separate batches)", mem[q], b);
        }
    }
    CU(cudaSetDevice(c->device));
    Scratch sc(c);
    AlnDev ad;
    std::vector<int> present;
    int rc = upload_alignment(c, sc, aln, N, A, c->d_offsets.p, ad, present);
    if (rc) return rc;
    double *d_out = nullptr;
    CU(sc.alloc(&d_out, (size_t)c->total * 3));
    rc = superpose_batches(c, sc, ad, A, ref, mem, n_pairs, batch_off, n_batches, nullptr, 0, 0, 4, true, d_out, out_rot, out_tran, out_ncommon);
    if (rc) return rc;
    CU(cudaMemcpyAsync(out_coords, d_out, sizeof(double) * (size_t)c->total * 3, cudaMemcpyDeviceToHost, c->stream));
    return finish_timed(c, "crt_superpose_pairs");
}

/* helper.write_distance_matrix, helper.py:183-203 */
int crt_format_matrix(crt_ctx *c, const double *matrix, int32_t n_rows, int32_t n_cols, const char *names, const int64_t *name_off,
                      int64_t *out_len)
{
    if (!c || (!matrix && (int64_t)n_rows * n_cols > 0) || !name_off || !out_len) return fail(CRT_E_ARG, "null argument");
    if (n_rows < 0 || n_cols < 0) return fail(CRT_E_ARG, "negative size");
    if (name_off[0] != 0) return fail(CRT_E_ARG, "name_off[0] must be 0");
    for (int i = 0; i < n_rows; ++i)
        if (name_off[i + 1] < name_off[i]) return fail(CRT_E_ARG, "name_off must be non-decreasing");
    if (name_off[n_rows] > 0 && !names) return fail(CRT_E_ARG, "null names");
    CU(cudaSetDevice(c->device));
    char header[32];
    const int hl = snprintf(header, sizeof(header), "%d\n", n_rows);          // f.write(f"{len(names)}\n"), :199
    c->text_len = 0;
    cudaStream_t st = c->stream;
    if (n_rows == 0) {
        int rc = c->text.ensure((size_t)hl);
        if (rc) return rc;
        CU(cudaMemcpyAsync(c->text.p, header, (size_t)hl, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        c->text_len = hl; *out_len = hl;
        return 0;
    }
    Scratch sc(c);
    double *d_M = nullptr;
    char *d_names = nullptr;
    long long *d_noff = nullptr, *d_len = nullptr, *d_off = nullptr;
    const size_t cells = (size_t)n_rows * n_cols;
    CU(sc.alloc(&d_M, cells));
    CU(sc.alloc(&d_names, (size_t)name_off[n_rows]));
    CU(sc.alloc(&d_noff, (size_t)n_rows + 1));
    CU(sc.alloc(&d_len, (size_t)n_rows));
    CU(sc.alloc(&d_off, (size_t)n_rows + 1));
    if (cells) CU(cudaMemcpyAsync(d_M, matrix, sizeof(double) * cells, cudaMemcpyHostToDevice, st));
    if (name_off[n_rows]) CU(cudaMemcpyAsync(d_names, names, (size_t)name_off[n_rows], cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_noff, name_off, sizeof(long long) * ((size_t)n_rows + 1), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev0, st));
    k_fmt_rowlen<<<n_rows, FMT_THREADS, 0, st>>>(d_M, n_cols, d_noff, d_len);
    k_fmt_scan<<<1, 1024, 0, st>>>(d_len, n_rows, hl, d_off);
    CU(cudaGetLastError());
    long long total = 0;
    CU(cudaMemcpyAsync(&total, d_off + n_rows, sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int rc = c->text.ensure((size_t)total + 16);
    if (rc) return rc;
    CU(cudaMemcpyAsync(c->text.p, header, (size_t)hl, cudaMemcpyHostToDevice, st));
    k_fmt_write<<<n_rows, FMT_THREADS, 0, st>>>(d_M, n_cols, d_names, d_noff, d_off, c->text.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, st));
    c->launches = 3;
    rc = finish_timed(c, "crt_format_matrix");
    if (rc) return rc;
    c->text_len = total;
    *out_len = total;
    return 0;
}

/* MultipleAlignment.write_alignment / to_sequence_alignment, multiple_alignment.py:287-309 */
int crt_format_fasta(crt_ctx *c, const int64_t *aln, int32_t N, int64_t A, const char *seqs, const int64_t *seq_off, const char *names,
                     const int64_t *name_off, int64_t *out_len)
{
    if (!c || !seq_off || !name_off || !out_len || (N > 0 && A > 0 && !aln)) return fail(CRT_E_ARG, "null argument");
    if (N < 0 || A < 0) return fail(CRT_E_ARG, "negative size");
    CU(cudaSetDevice(c->device));
    c->text_len = 0;
    if (N == 0) { *out_len = 0; return 0; }
    if (seq_off[0] != 0 || name_off[0] != 0) return fail(CRT_E_ARG, "offsets must start at 0");
    std::vector<long long> rec((size_t)N + 1, 0);
    for (int p = 0; p < N; ++p) {
        if (seq_off[p + 1] < seq_off[p] || name_off[p + 1] < name_off[p]) return fail(CRT_E_ARG, "offsets must be non-decreasing");
        rec[(size_t)p + 1] = rec[(size_t)p] + 1 + (name_off[p + 1] - name_off[p]) + 1 + A + 1;      // f">{name}\n{aligned}\n", :309
    }
    if ((seq_off[N] > 0 && !seqs) || (name_off[N] > 0 && !names)) return fail(CRT_E_ARG, "null text");
    const long long total = rec[(size_t)N];
    int rc = c->text.ensure((size_t)total + 16);
    if (rc) return rc;
    Scratch sc(c);
    long long *d_aln = nullptr, *d_soff = nullptr, *d_noff = nullptr, *d_rec = nullptr;
    char *d_seqs = nullptr, *d_names = nullptr;
    int *d_bad = nullptr;
    CU(sc.alloc(&d_aln, (size_t)N * A));
    CU(sc.alloc(&d_soff, (size_t)N + 1));
    CU(sc.alloc(&d_noff, (size_t)N + 1));
    CU(sc.alloc(&d_rec, (size_t)N + 1));
    CU(sc.alloc(&d_seqs, (size_t)seq_off[N]));
    CU(sc.alloc(&d_names, (size_t)name_off[N]));
    CU(sc.alloc(&d_bad, 1));
    cudaStream_t st = c->stream;
    if (A > 0) CU(cudaMemcpyAsync(d_aln, aln, sizeof(long long) * (size_t)N * A, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_soff, seq_off, sizeof(long long) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_noff, name_off, sizeof(long long) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rec, rec.data(), sizeof(long long) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
    if (seq_off[N]) CU(cudaMemcpyAsync(d_seqs, seqs, (size_t)seq_off[N], cudaMemcpyHostToDevice, st));
    if (name_off[N]) CU(cudaMemcpyAsync(d_names, names, (size_t)name_off[N], cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    CU(cudaEventRecord(c->ev0, st));
    k_fasta<<<N, 256, 0, st>>>(d_aln, A, d_seqs, d_soff, d_names, d_noff, d_rec, c->text.p, d_bad);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, st));
    c->launches = 1;
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    rc = finish_timed(c, "crt_format_fasta");
    if (rc) return rc;
    if (bad) return fail(CRT_E_ARG, "alignment holds an index < -1 or beyond the end of its sequence (the reference raises IndexError)");
    c->text_len = total;
    *out_len = total;
    return 0;
}

/* make_count_matrix, multiple_alignment.py:128-134 */
int crt_count_matrix(crt_ctx *c, const int64_t *indices, const int64_t *offsets, int32_t N, int32_t alphabet_size, double *out)
{
    if (!c || !offsets || !out || (N > 0 && offsets[N] > 0 && !indices)) return fail(CRT_E_ARG, "null argument");
    if (N <= 0 || alphabet_size <= 0) return fail(CRT_E_ARG, "empty count matrix (%d x %d)", N, alphabet_size);
    if (offsets[0] != 0) return fail(CRT_E_ARG, "offsets[0] must be 0");
    for (int p = 0; p < N; ++p)
        if (offsets[p + 1] < offsets[p]) return fail(CRT_E_ARG, "offsets must be non-decreasing");
    CU(cudaSetDevice(c->device));
    Scratch sc(c);
    long long *d_idx = nullptr, *d_off = nullptr;
    double *d_out = nullptr;
    int *d_bad = nullptr;
    const size_t cells = (size_t)N * alphabet_size;
    CU(sc.alloc(&d_idx, (size_t)offsets[N]));
    CU(sc.alloc(&d_off, (size_t)N + 1));
    CU(sc.alloc(&d_out, cells));
    CU(sc.alloc(&d_bad, 1));
    cudaStream_t st = c->stream;
    if (offsets[N]) CU(cudaMemcpyAsync(d_idx, indices, sizeof(long long) * (size_t)offsets[N], cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_off, offsets, sizeof(long long) * ((size_t)N + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_out, 0, sizeof(double) * cells, st));
    CU(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    CU(cudaEventRecord(c->ev0, st));
    k_count_matrix<<<N, 256, 0, st>>>(d_idx, d_off, N, alphabet_size, d_out, d_bad);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, st));
    c->launches = 1;
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out, d_out, sizeof(double) * cells, cudaMemcpyDeviceToHost, st));
    int rc = finish_timed(c, "crt_count_matrix");
    if (rc) return rc;
    if (bad) return fail(CRT_E_ARG, "shapemer index outside [0, %d) (the reference indexes out of bounds)", alphabet_size);
    return 0;
}

/* braycurtis, multiple_alignment.py:137-145 */
int crt_braycurtis(crt_ctx *c, const double *counts_1, int32_t n1, const double *counts_2, int32_t n2, int32_t K, double *out)
{
    if (!c || !counts_1 || !counts_2 || !out) return fail(CRT_E_ARG, "null argument");
    if (n1 <= 0 || n2 <= 0 || K <= 0) return fail(CRT_E_ARG, "empty input (%d, %d, %d)", n1, n2, K);
    CU(cudaSetDevice(c->device));
    Scratch sc(c);
    double *d_a = nullptr, *d_b = nullptr, *d_out = nullptr;
    CU(sc.alloc(&d_a, (size_t)n1 * K));
    const bool same = counts_1 == counts_2 && n1 == n2;
    if (!same) CU(sc.alloc(&d_b, (size_t)n2 * K));
    CU(sc.alloc(&d_out, (size_t)n1 * n2));
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(d_a, counts_1, sizeof(double) * (size_t)n1 * K, cudaMemcpyHostToDevice, st));
    if (!same) CU(cudaMemcpyAsync(d_b, counts_2, sizeof(double) * (size_t)n2 * K, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(c->ev0, st));
    k_braycurtis<<<dim3((unsigned)((n2 + BC_TILE - 1) / BC_TILE), (unsigned)((n1 + BC_TILE - 1) / BC_TILE)), 256, 0, st>>>(d_a, n1, same ? d_a : d_b,
                                                                                                                         n2, K, d_out);
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev1, st));
    c->launches = 1;
    CU(cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n1 * n2, cudaMemcpyDeviceToHost, st));
    return finish_timed(c, "crt_braycurtis");
}

int crt_text_fetch(crt_ctx *c, char *out, int64_t cap)
{
    if (!c || (!out && c->text_len > 0)) return fail(CRT_E_ARG, "null argument");
    if (cap < c->text_len) return fail(CRT_E_ARG, "buffer of %lld bytes for %lld bytes of text", (long long)cap, c->text_len);
    if (c->text_len == 0) return 0;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(out, c->text.p, (size_t)c->text_len, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
