// crt_api.cu -- host side of the C ABI declared in include/caretta_b200.h.
// Plain CUDA runtime; no torch types.  See crt_kernels.cuh for the kernels and DESIGN.md for the data layout.
#include "../../include/caretta_b200.h"
#include "crt_kernels.cuh"
#include "crt_fill_f32.cuh"
#include "crt_fill1_v2.cuh"
#include "crt_fill1_v4.cuh"
#include "crt_fill_tc.cuh"
#include "crt_node_fill.cuh"
#include "crt_fill2_v3.cuh"
#include "crt_dp_batch.cuh"
#include "crt_nj.cuh"
#include "crt_node.cuh"
#include "crt_consumers.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace crt;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? CRT_E_NOMEM : CRT_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                   \
    } while (0)

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;     // elements
    int ensure(size_t n)
    {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e != cudaSuccess) { p = nullptr; return fail(CRT_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e)); }
        cap = want;
        return 0;
    }
    // grow to n elements keeping the first `keep` (device-to-device copy on `st`, which is synchronised before the old block goes)
    int grow_keep(size_t n, size_t keep, cudaStream_t st)
    {
        if (n <= cap) return 0;
        T *old = p;
        size_t want = n + n / 2 + 64;
        T *np_ = nullptr;
        cudaError_t e = cudaMalloc(&np_, want * sizeof(T));
        if (e != cudaSuccess) return fail(CRT_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want * sizeof(T), cudaGetErrorString(e));
        if (old && keep) {
            e = cudaMemcpyAsync(np_, old, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { cudaFree(np_); return fail(CRT_E_CUDA, "pool copy failed: %s", cudaGetErrorString(e)); }
        }
        if (old) cudaFree(old);
        p = np_; cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// result slots whose status carries ST_TIE: list[atomicAdd(count)] = slot (the host sorts the list)
__global__ void k_collect_tie(const int *status, long long n, int *list, int *count)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n && (status[q] & ST_TIE)) list[atomicAdd(count, 1)] = (int)q;
}
__global__ void k_gather_tie(const crt::TieInfo *info, const int *list, int n, crt::TieInfo *out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) out[q] = info[list[q]];
}
__global__ void k_mark_status(int *status, const int *list, int n, int bit)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) status[list[q]] |= bit;
}

// end of an overlapped re-run: the marked slots lose ST_TIE (it only kept the stage-2 fill of the main run off them) and get `bit`
__global__ void k_finish_status(int *status, const int *list, int n, int bit)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) status[list[q]] = (status[list[q]] & ~ST_TIE) | bit;
}

struct HostUnit {
    Unit u;
    int C;          // columns per lane
    int multi;      // n_strips > 1
    double cost;    // G * m
};

struct Batch {
    size_t first, count; int C, multi; size_t tb_n, rows2_n, bnd_n, path_n; int n_dense;
    // tensor-core stage 1 (crt_fill_tc.cuh): the batch's pairs regrouped into rounds; what is left runs on k_fill1_v4 units
    bool tc = false;
    size_t tc_r0 = 0, tc_nr = 0;        // rounds [tc_r0, tc_r0 + tc_nr) of the run's round array
    size_t tc_p0 = 0, tc_np = 0;        // partner records
    size_t lf_u0 = 0, lf_nu = 0;        // left-over stage-1 units
    int lf_dense = 0;                   // pairs of the left-over units
    size_t tc_bnd_n = 0;                // floats of strip boundary values (rounds first, then the left-over units)
    // node contexts: stage 1 from precomputed scores (crt_node_fill.cuh)
    size_t s_n = 0;                     // doubles of score matrices
    long long max_tiles = 0;            // most 16 x 64 score tiles of a unit
    int max_strips = 1;                 // most strips of a unit
    bool single = true;                 // every unit holds one pair
};

}  // namespace

struct crt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    int sm_count = 0, clock_khz = 0;
    size_t mem_total = 0;

    // chains
    int N = 0, d = 0, D = 0;
    long long total = 0;
    std::vector<long long> offsets;      // host copy
    int max_len = 0;
    DevBuf<double> coords, tensors, centroid, rec64, stats;      // stats: [STATS_BLOCKS][32] partial sums, [32] mean
    DevBuf<int> flag;
    DevBuf<long long> d_offsets;
    DevBuf<int> chain_of, meta;
    DevBuf<float> rec32;
    DevBuf<float4> cols2;
    double prep_gamma_t = -1, prep_gamma_c = -1;   // parameters the fp32 records were built with

    // run workspaces: one set per stream so that the batches of a run overlap (the tail of one batch's fill runs
    // beside the next batch's, and the latency-bound traceback hides under the fills)
    struct Workspace {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};      // serial mode: phase boundaries of the last batch
        DevBuf<uint4> tb;
        DevBuf<unsigned char> rows2, bnd, bnd2;      // bnd: stage-1 strip boundaries, bnd2: stage-2 (the two fills of
        DevBuf<short2> path;                         // different batches run concurrently in pipeline mode)
        DevBuf<double> svals;                        // node contexts: precomputed stage-1 scores of the batch
        cudaEvent_t e_f1 = nullptr, e_t = nullptr, e_f2 = nullptr;   // pipeline mode: stage completion of the last batch here
    };
    // pipeline mode: one stream per stage, so that stage 1 of batch k+1, the traceback of batch k and stage 2 of
    // batch k-1 are resident together (the light stage-2 and traceback warps fill the registers stage 1 leaves free)
    cudaStream_t s_f1[2] = {nullptr, nullptr}, s_tr = nullptr, s_f2[2] = {nullptr, nullptr};   // fills alternate between two
                                                                                              // streams: the tail of one launch overlaps the head of the next
    static constexpr int MAX_WS = 4;
    Workspace ws[MAX_WS];
    Workspace ws_rr;                     // the float64 re-run of the marked pairs: own buffers and stream, so that it overlaps the last batches
    cudaEvent_t e_traces = nullptr, e_rr = nullptr;
    int n_streams = 3;
    DevBuf<Unit> d_units, d_units2;      // d_units2: the float64 re-run of the tie pairs
    DevBuf<crt::TcRound> d_tc_rounds;    // tensor-core stage 1: rounds, partner records, left-over units of the run, round counters per batch
    DevBuf<crt::TcPartner> d_tc_partners;
    DevBuf<Unit> d_tc_left;
    DevBuf<int> d_tc_counter;
    long long tc_pairs = 0;              // pairs of the last run whose stage 1 ran on the tensor cores
    DevBuf<int> tie_list;                // [0] = count, [1..] = result slots marked ST_TIE
    DevBuf<crt::TieInfo> tie_info, tie_info_list;     // windowed re-run: per result slot / gathered for the marked slots
    DevBuf<short2> tie_pool;             // saved path prefixes of the marked pairs
    DevBuf<unsigned long long> tie_pool_used;
    long long rerun_pairs = 0;
    double rerun_ms = 0;
    double tb_bytes = 0;                 // traceback words the stage-1 fills of the last run wrote (all batches)
    DevBuf<int> path_len, pair_istar, pair_zflag, ncommon, status;
    DevBuf<int> d_pi, d_pj;              // pair ids of the cached all-vs-all plan, result order (device copy for the dense scatter)
    DevBuf<double> dense;                // [1 or 3][N][N] dense matrices of crt_pairwise_all
    // crt_scatter_gathered: the pairs of every rank's shard (rank-major, shard order), cached per chain set and world
    DevBuf<int> lay_pi, lay_pj;
    DevBuf<long long> lay_off;
    unsigned long long lay_hash = 0;
    int lay_world = 0;
    long long lay_total = 0;
    DevBuf<double> score, score1, rmsd, tm, xform;
    DevBuf<float> f32tmp;
    double phase_ms[4] = {0, 0, 0, 0};      // fill1, trace, rows2, fill2 (only meaningful with one stream)

    // plan cache of the all-vs-all shard: the unit list, batches and device copy depend only on the chain lengths
    struct PlanCache {
        bool valid = false;
        unsigned long long offsets_hash = 0;
        int rank = -1, world = -1, prec = -1, ns = -1;
        size_t budget = 0;
        std::vector<Batch> batches;
        std::vector<Unit> hu;
        std::vector<crt::TcRound> tc_rounds;
        std::vector<crt::TcPartner> tc_partners;
        std::vector<Unit> tc_left;
        long long n_pairs = 0;
        double cells = 0;
    } plan;
    unsigned long long offsets_hash = 0;

    // progressive-alignment nodes (crt_progressive_node): a private context for the stage-1 pair run of the two
    // consensus sequences, and grow-only device buffers
    crt_ctx *node_ctx = nullptr;
    DevBuf<double> nd_w, nd_S, nd_bnd, nd_f, nd_score, nd_xf2, nd_t, nd_c, nd_wm;
    DevBuf<unsigned char> nd_B;
    DevBuf<int> nd_a1, nd_a2, nd_len;

    // crt_msa_*: device-resident sequence pool of the progressive alignment (leaves + every intermediate node)
    struct MsaPool {
        bool active = false;
        int d = 0;
        long long used = 0;                   // rows in use
        DevBuf<double> t, c, w;               // tensors [rows, d], coordinates [rows, 3], consensus weights [rows]
        std::vector<long long> off;           // first row of every sequence
        std::vector<int> len;                 // its length
        // bookkeeping of crt_msa_compose: the children and the two alignments of every node made since crt_msa_begin
        int n_leaves = 0;
        std::vector<int> ch1, ch2;
        std::vector<std::vector<int32_t>> al1, al2;
    } pool;
    DevBuf<long long> lv_tab, lv_out_off;

    DevBuf<char> arena;                   // scratch of the consumer / neighbor-joining calls (crt_consumers_api.inl: Scratch)
    bool coords_only = false;             // chain set made by crt_set_coords: consumers only, no pair runs
    bool stage1_only = false;             // node contexts: the run stops after the traceback / Kabsch (no stage-2 fill)
    bool no_byproducts = false;           // crt_pairwise_all without RMSD / TM matrices: the tracebacks skip the by-product pass
    DevBuf<DpProblem> lv_probs;           // crt_progressive_level: per-node problem records and packed level buffers
    DevBuf<double> lv_mult, lv_xf2;
    DevBuf<long long> lv_off;

    // text produced by crt_format_matrix / crt_format_fasta, fetched with crt_text_fetch
    DevBuf<char> text;
    long long text_len = 0;

    // last run
    long long run_pairs = 0;
    double elapsed_ms = 0, cell_updates = 0;
    long long launches = 0;
    int last_streams = 1;
    std::vector<int> run_pi, run_pj;     // pair ids in result order
};

namespace {

// Scratch memory of one call: bump allocation from an arena the context keeps (cudaMalloc / cudaFree of a few hundred MB per
// call cost up to hundreds of milliseconds, erratically).  What does not fit is allocated for this call only and the arena
// grows to the call's total when the call is over, so the next call of the same size allocates nothing.
struct Scratch {
    crt_ctx *c;
    size_t used = 0, wanted = 0;
    std::vector<void *> extra;
    explicit Scratch(crt_ctx *ctx) : c(ctx) {}
    ~Scratch()
    {
        for (void *p : extra) cudaFree(p);
        if (wanted > c->arena.cap) { cudaStreamSynchronize(c->stream); c->arena.release(); c->arena.ensure(wanted + wanted / 8); }
    }
    template <typename T>
    cudaError_t alloc(T **out, size_t n)
    {
        const size_t bytes = ((n ? n : 1) * sizeof(T) + 255) & ~(size_t)255;
        wanted += bytes;
        if (used + bytes <= c->arena.cap) {
            *out = reinterpret_cast<T *>(c->arena.p + used);
            used += bytes;
            return cudaSuccess;
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) extra.push_back(p);
        *out = static_cast<T *>(p);
        return e;
    }
};

// ---------------------------------------------------------------------------------------------- kernel dispatch
struct Choice { int C; int n_strips; };

// columns per lane: bounded by the registers the stage-1 column vectors need (C * D values per lane)
int cmax_for(int precision, int D)
{
    if (precision == CRT_FP32) return D <= 10 ? 10 : 6;
    return D <= 10 ? 4 : 3;
}

Choice choose_cols(int m, int precision, int D)
{
    static const int allowed64[] = {2, 3, 4, 5, 6, 8, 10};
    static const int allowed32[] = {2, 4, 6, 8, 10, 10, 10};      // the fp32 fills pair columns: even C only
    const int *allowed = precision == CRT_FP32 ? allowed32 : allowed64;
    const int cmax = cmax_for(precision, D);
    int ns = (m + 32 * cmax - 1) / (32 * cmax);
    if (ns < 1) ns = 1;
    int need = (m + 32 * ns - 1) / (32 * ns);
    int C = cmax;
    for (int q = 0; q < 7; ++q)
        if (allowed[q] >= need) { C = allowed[q]; break; }
    if (C > cmax) C = cmax;
    return Choice{C, ns};
}

template <typename P, bool DIFF, bool CODES, int CMAX>
int launch_fill_c(int C, bool multi, const Unit *units, int n, typename P::Args args, FillOut out, cudaStream_t st)
{
    const int grid = (n + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int block = WARPS_PER_CTA * 32;
#define CRT_CASE(CC)                                                                                         \
    case CC:                                                                                                 \
        if constexpr (CC <= CMAX) {                                                                          \
            if (multi) k_fill<P, CC, DIFF, CODES, true><<<grid, block, 0, st>>>(units, n, args, out);         \
            else k_fill<P, CC, DIFF, CODES, false><<<grid, block, 0, st>>>(units, n, args, out);             \
            break;                                                                                           \
        } else return fail(CRT_E_ARG, "no kernel for C=%d (max %d)", C, CMAX);
    switch (C) {
        CRT_CASE(2) CRT_CASE(3) CRT_CASE(4) CRT_CASE(5) CRT_CASE(6) CRT_CASE(8) CRT_CASE(10)
    default: return fail(CRT_E_ARG, "no kernel for C=%d", C);
    }
#undef CRT_CASE
    CU(cudaGetLastError());
    return 0;
}

// stage-1 fp32 kernel: k_fill1_v4 (two columns per packed FMA, fast/checked row groups, tie flags: 3 bits per cell)
template <int D>
int launch_fill1_f32(int C, bool multi, const Unit *units, int n, Fill1Args args, FillOut out, const long long *offsets, TieArgs tie,
                     cudaStream_t st)
{
#define CRT_CASE(CC)                                                                                  \
    case CC:                                                                                          \
        if constexpr (CC * D <= 100) {                                                                \
            if (multi) k_fill1_v4<D, CC, true><<<n, 32, 0, st>>>(units, n, args, out, offsets, tie);   \
            else k_fill1_v4<D, CC, false><<<n, 32, 0, st>>>(units, n, args, out, offsets, tie);       \
            break;                                                                                    \
        } else return fail(CRT_E_ARG, "no fp32 stage-1 kernel for C=%d D=%d", C, D);
    switch (C) {
        CRT_CASE(2) CRT_CASE(4) CRT_CASE(6) CRT_CASE(8) CRT_CASE(10)
    default: return fail(CRT_E_ARG, "no fp32 stage-1 kernel for C=%d", C);
    }
#undef CRT_CASE
    CU(cudaGetLastError());
    return 0;
}

int launch_fill2_f32(int C, bool multi, const Unit *units, int n, Fill2Args args, FillOut out, const long long *offsets, cudaStream_t st)
{
#define CRT_CASE(CC)                                                                              \
    case CC:                                                                                      \
        if (multi) k_fill2_v3<CC, true><<<n, 32, 0, st>>>(units, n, args, out, offsets);           \
        else k_fill2_v3<CC, false><<<n, 32, 0, st>>>(units, n, args, out, offsets);               \
        break;
    switch (C) {
        CRT_CASE(2) CRT_CASE(4) CRT_CASE(6) CRT_CASE(8) CRT_CASE(10)
    default: return fail(CRT_E_ARG, "no fp32 stage-2 kernel for C=%d", C);
    }
#undef CRT_CASE
    CU(cudaGetLastError());
    return 0;
}

int pad_dim(int d)
{
    if (d <= 10) return 10;
    if (d <= 16) return 16;
    return -1;
}

bool env_pipe();
bool env_tc();
int env_tc_min_part();
int env_batches();
int env_streams();
size_t env_budget();

// ---------------------------------------------------------------------------------------------- unit building
constexpr int MAX_PAIRS_PER_UNIT = 32;
// Rows streamed by one unit.  Longer units amortise the 31-step pipeline fill of the systolic array, shorter units give
// the block scheduler more, finer work items (a short run sharded over 8 GPUs must not end on a few long stragglers).
// Default: about 24 units per resident-warp slot of the rank's shard, between one chain and 3072 rows; the rule depends
// on (offsets, world) only, so the host-side planner and every precision enumerate the same units.
// CARETTA_B200_UNIT_ROWS overrides it (tuning).
int env_unit_rows()
{
    const char *e = getenv("CARETTA_B200_UNIT_ROWS");
    int n = e ? atoi(e) : 0;
    return n <= 0 ? 0 : std::min(std::max(n, 64), 1 << 20);
}

constexpr int UNIT_ROWS_MIN = 512, UNIT_ROWS_MAX = 3072, NOMINAL_STRIP_COLS = 320, RESIDENT_SLOTS = 148 * 8, UNITS_PER_SLOT = 24;

// row-steps (rows x nominal strips) one unit should stream, for this shard
double unit_rowsteps_target(const std::vector<long long> &offsets, int N, int world)
{
    double total = 0.0, rows_before = 0.0;
    for (int j = 0; j < N; ++j) {
        const double len = (double)(offsets[j + 1] - offsets[j]);
        total += rows_before * std::ceil(len / NOMINAL_STRIP_COLS);
        rows_before += len;
    }
    return total / std::max(world, 1) / ((double)RESIDENT_SLOTS * UNITS_PER_SLOT);
}

int unit_rows_cap(double target_rowsteps, int m)
{
    if (const int forced = env_unit_rows()) return forced;
    const double strips = std::ceil((double)m / NOMINAL_STRIP_COLS);
    const double r = target_rowsteps / std::max(strips, 1.0);
    return (int)std::min<double>(std::max<double>(r, UNIT_ROWS_MIN), UNIT_ROWS_MAX);
}

// one_variant (the float64 re-run of the fp32 mode): every unit on the widest multi-strip kernel of the precision, whatever its
// length, so that the re-run is ONE batch per stage: its batches run one after the other and each waits for its longest
// single-warp pair (on the first 1000 chains of C4, lengths 50 - 1000, the re-run fell into ~10 batches: 27 ms for 4 353 pairs)
void finish_unit(const crt_ctx *c, HostUnit &h, int precision, bool one_variant = false)
{
    Choice ch = choose_cols(h.u.m, precision, c->D);
    if (one_variant) {
        ch.C = cmax_for(precision, c->D);
        ch.n_strips = std::max(1, (h.u.m + 32 * ch.C - 1) / (32 * ch.C));
    }
    h.C = ch.C;
    h.u.n_strips = ch.n_strips;
    h.multi = one_variant || ch.n_strips > 1;
    h.u.tchunks = (h.u.G + 31 + 3) / 4;
    h.cost = (double)h.u.G * (double)h.u.m;
    (void)c;
}

// all-vs-all units in a deterministic order: column chain j ascending, runs of row chains ascending
void build_all_units(const crt_ctx *c, int precision, int world, std::vector<HostUnit> &out)
{
    const double target = unit_rowsteps_target(c->offsets, c->N, world);
    out.clear();
    for (int j = 1; j < c->N; ++j) {
        const int TARGET_ROWS_PER_UNIT = unit_rows_cap(target, (int)(c->offsets[j + 1] - c->offsets[j]));
        int i = 0;
        while (i < j) {
            HostUnit h{};
            h.u.row_chain0 = i;
            h.u.row_base = c->offsets[i];
            h.u.col_chain = j;
            h.u.col_base = (int)c->offsets[j];
            h.u.m = (int)(c->offsets[j + 1] - c->offsets[j]);
            int cnt = 0, maxn = 0;
            long long G = 0;
            while (i < j && cnt < MAX_PAIRS_PER_UNIT) {
                int n = (int)(c->offsets[i + 1] - c->offsets[i]);
                if (cnt > 0 && G + n > TARGET_ROWS_PER_UNIT) break;
                G += n; maxn = std::max(maxn, n); ++cnt; ++i;
            }
            h.u.G = (int)G;
            h.u.n_pairs = cnt;
            h.u.path_stride = maxn + h.u.m;
            finish_unit(c, h, precision);
            out.push_back(h);
        }
    }
}

// deal units to ranks: sort by cost (stable, descending), snake order
void shard_units(std::vector<HostUnit> &units, int rank, int world)
{
    if (world <= 1) return;
    std::vector<int> idx(units.size());
    for (size_t q = 0; q < idx.size(); ++q) idx[q] = (int)q;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return units[a].cost > units[b].cost; });
    std::vector<HostUnit> mine;
    for (size_t q = 0; q < idx.size(); ++q) {
        int round = (int)(q / world), pos = (int)(q % world);
        int owner = (round & 1) ? world - 1 - pos : pos;
        if (owner == rank) mine.push_back(units[idx[q]]);
    }
    // keep the deterministic (j, i0) order inside the shard
    std::stable_sort(mine.begin(), mine.end(), [](const HostUnit &a, const HostUnit &b) {
        if (a.u.col_chain != b.u.col_chain) return a.u.col_chain < b.u.col_chain;
        return a.u.row_chain0 < b.u.row_chain0;
    });
    units.swap(mine);
}

void assign_pairs(std::vector<HostUnit> &units, std::vector<int> *pi, std::vector<int> *pj, long long *n_pairs, double *cells,
                  const crt_ctx *c)
{
    long long p = 0;
    double cu = 0;
    if (pi) pi->clear();
    if (pj) pj->clear();
    for (auto &h : units) {
        h.u.pair_base = (int)p;
        for (int q = 0; q < h.u.n_pairs; ++q) {
            if (pi) pi->push_back(h.u.row_chain0 + q);
            if (pj) pj->push_back(h.u.col_chain);
        }
        p += h.u.n_pairs;
        cu += 2.0 * (double)h.u.G * (double)h.u.m;
    }
    (void)c;
    if (n_pairs) *n_pairs = p;
    if (cells) *cells = cu;
}

int ensure_prepared(crt_ctx *c, const crt_params *prm);

// ---------------------------------------------------------------------------------------------- the run
// Executes the units (already carrying pair_base) in batches bounded by a workspace budget.
// If paths are requested, they are copied back per batch into host vectors (descending order as walked).
struct PathSink {
    bool want = false;
    std::vector<std::vector<int>> a1, a2;   // per pair, ascending residue order
};

size_t env_budget()
{
    const char *e = getenv("CARETTA_B200_WORKSPACE_MB");
    size_t mb = e ? (size_t)atoll(e) : 8192;      // per workspace set (up to 4 sets in flight); B200 has 180 GB
    if (mb < 16) mb = 16;
    return mb << 20;
}

int env_streams()
{
    const char *e = getenv("CARETTA_B200_STREAMS");
    int n = e ? atoi(e) : 3;
    return std::min(std::max(n, 1), (int)crt_ctx::MAX_WS);
}

// CARETTA_B200_PIPE=0 selects the older "one stream per batch" schedule (A/B runs); default is the stage pipeline
bool env_pipe()
{
    const char *e = getenv("CARETTA_B200_PIPE");
    return e ? atoi(e) != 0 : true;
}

int env_batches()
{
    const char *e = getenv("CARETTA_B200_BATCHES");
    int n = e ? atoi(e) : 8;
    return std::min(std::max(n, 1), 256);
}

// Batches of the stage pipeline for a run of `cells` residue pairs between chains of mean length `mean_len`.  The tracebacks of the
// batches run one after the other on their stream and a traceback launch has a latency floor (one thread walks a pair: ~0.5 ms +
// 1.7 us per path step, whatever the number of pairs), so a SMALL run -- a shard of a strong-scaled job -- cut into 8 batches waits
// for 8 floors: measured on rank 0's shard of C3 over 8 ranks, 8 batches 11.8 ms, 3 batches 8.4 ms (4 ranks: 16.1 -> 15.3 ms; the
// whole of C3 on one GPU: 8 batches 59.4 ms, 6 batches 59.9 ms).  Rule: the floors may take 40 % of the fills' time; 2 <= batches <= 8.
// CARETTA_B200_BATCHES overrides.
int auto_batches(double cells, double mean_len)
{
    if (getenv("CARETTA_B200_BATCHES")) return env_batches();
    const double fill_ms = cells * 1.3e-9, floor_ms = 0.5 + 0.0017 * 2.0 * mean_len;
    const int b = (int)(0.4 * fill_ms / floor_ms);
    return std::min(std::max(b, 2), 8);
}

// tie3: the codes come from k_fill1_v4 (3 bits per cell); otherwise from the float64 k_fill (2 bits per cell)
int launch_trace(int C, const TraceArgs &ta, int nu, int n_dense, cudaStream_t st, bool tie3, bool warp = false)
{
    // warp: a warp per pair, the passes over the path by the warp (k_trace_w).  Chosen by CONTEXT, never by the size of the batch
    // -- tree levels and the float64 re-run of the fp32 mode, whose batches are small and latency-bound --, so that a pair's bits do
    // not depend on how a run is cut into batches or shards (CARETTA_B200_TRACE_WARP=0: a thread per pair everywhere)
    static const bool warp_on = !(getenv("CARETTA_B200_TRACE_WARP") && atoi(getenv("CARETTA_B200_TRACE_WARP")) == 0);
    const bool wp = warp && warp_on;
    const int grid = wp ? (n_dense * 32 + TRACE_THREADS - 1) / TRACE_THREADS : (n_dense + TRACE_THREADS - 1) / TRACE_THREADS;
#define CRT_CASE(CC)                                                                          \
    case CC:                                                                                  \
        if (wp) {                                                                             \
            if (tie3) k_trace_w<CC, true><<<grid, TRACE_THREADS, 0, st>>>(ta, nu, n_dense);    \
            else k_trace_w<CC, false><<<grid, TRACE_THREADS, 0, st>>>(ta, nu, n_dense);       \
        } else if (tie3) k_trace<CC, true><<<grid, TRACE_THREADS, 0, st>>>(ta, nu, n_dense);   \
        else k_trace<CC, false><<<grid, TRACE_THREADS, 0, st>>>(ta, nu, n_dense);             \
        break;
    switch (C) {
        CRT_CASE(2) CRT_CASE(3) CRT_CASE(4) CRT_CASE(6) CRT_CASE(8) CRT_CASE(10)
    default: return fail(CRT_E_ARG, "no trace kernel for C=%d", C);
    }
#undef CRT_CASE
    CU(cudaGetLastError());
    return 0;
}

// Tie detection of the fp32 production mode (crt_fill1_v4.cuh).  CARETTA_B200_TIE_C: candidates closer than C * 2^-53 * H tie in
// the reference's float64 matrix (default 8; 2 catches every such pair of config C3, 1 misses 5 of 814); CARETTA_B200_TIE_EPS: relative closeness fp32 cannot order (default 1e-4);
// CARETTA_B200_TIE_RERUN=0 leaves the marked pairs as fp32 computed them (study runs).
TieArgs env_tie()
{
    const char *ec = getenv("CARETTA_B200_TIE_C"), *ee = getenv("CARETTA_B200_TIE_EPS");
    const double cc = ec ? atof(ec) : 8.0, ep = ee ? atof(ee) : 1e-4;
    return TieArgs{(float)(cc * 1.1102230246251565e-16), (float)ep};
}
bool env_tie_rerun()
{
    const char *e = getenv("CARETTA_B200_TIE_RERUN");
    return e ? atoi(e) != 0 : true;
}


// ---------------------------------------------------------------------------------------------- tensor-core stage 1
// CARETTA_B200_TC=1 moves stage 1 of the fp32 mode to the tensor-core kernel k_fill1_tc wherever a column chain has enough
// partners for a round.  Off by default: the kernel is 1.45 x faster per cell than k_fill1_v4, but at 1000 chains a fifth of
// the lanes of the rounds idle and the run gains 3 %; and its exponents (tf32 hi / lo split) carry about twice the error of the
// fp32 FMA chain, which lets 1 pair of 499 500 (config C3) and 3 of 124 750 (C5) take an unmarked different path
// (profiles/r02_tensor_core_stage1.md).
bool env_tc()
{
    const char *e = getenv("CARETTA_B200_TC");
    return e ? atoi(e) != 0 : false;
}
// a round with fewer partners than this costs more on the tensor-core kernel (a round takes the time of 128 lanes whatever it
// holds) than on the systolic one (measured 1.45 x per cell): its pairs stay on k_fill1_v4
int env_tc_min_part()
{
    const char *e = getenv("CARETTA_B200_TC_MIN");
    const int n = e ? atoi(e) : 72;
    return std::min(std::max(n, 1), TC_LANES);
}

// Regroups the pairs of one batch (units bu[0..count), slot order: column chain ascending) for k_fill1_tc: per column chain the
// partner row chains are sorted by length (stable) and cut into rounds of 128; a tail shorter than env_tc_min_part() becomes
// units of consecutive chains for k_fill1_v4.  Result slots are untouched.  Indices in the records are batch-local.
void plan_tc_batch(const crt_ctx *c, const Unit *bu, size_t count, Batch &b, std::vector<TcRound> &rounds, std::vector<TcPartner> &parts,
                   std::vector<Unit> &left)
{
    struct Cand { int i, n, slot; };
    std::vector<Cand> cand, rest;
    const int min_part = env_tc_min_part();
    b.tc = true;
    b.tc_r0 = rounds.size(); b.tc_p0 = parts.size(); b.lf_u0 = left.size();
    size_t tb_n = 0, path_n = 0, bnd_n = 0;
    int lf_dense = 0;
    size_t k = 0;
    while (k < count) {
        const int j = bu[k].col_chain, m = bu[k].m;
        cand.clear(); rest.clear();
        size_t k2 = k;
        for (; k2 < count && bu[k2].col_chain == j; ++k2)
            for (int q = 0; q < bu[k2].n_pairs; ++q) {
                const int i = bu[k2].row_chain0 + q;
                cand.push_back(Cand{i, (int)(c->offsets[i + 1] - c->offsets[i]), bu[k2].pair_base + q});
            }
        std::stable_sort(cand.begin(), cand.end(), [](const Cand &x, const Cand &y) { return x.n > y.n; });
        const int n_strips = (m + TC_SC - 1) / TC_SC;
        const int strip_w = (((m + n_strips - 1) / n_strips + 15) / 16) * 16;
        int tiles_row = 0;
        for (int s = 0; s < n_strips; ++s) tiles_row += (std::min(strip_w, ((m - s * strip_w + 15) / 16) * 16) + TC_TILE - 1) / TC_TILE;
        size_t pos = 0;
        while (pos < cand.size()) {
            const size_t cnt = std::min<size_t>(TC_LANES, cand.size() - pos);
            if ((int)cnt < min_part) { rest.assign(cand.begin() + pos, cand.end()); break; }
            TcRound R{};
            R.col_base = bu[k].col_base; R.col_chain = j; R.m = m; R.n_strips = n_strips; R.strip_w = strip_w;
            R.part_base = (int)(parts.size() - b.tc_p0); R.n_part = (int)cnt; R.max_rows = cand[pos].n;
            R.bnd_base = (long long)bnd_n;
            if (n_strips > 1) bnd_n += (size_t)R.max_rows * TC_LANES;
            const int ridx = (int)(rounds.size() - b.tc_r0);
            for (size_t q = pos; q < pos + cnt; ++q) {
                TcPartner P{};
                P.tb_base = (long long)tb_n; tb_n += (size_t)cand[q].n * tiles_row;
                P.path_base = (long long)path_n; path_n += (size_t)cand[q].n + m;
                P.row_base = (int)c->offsets[cand[q].i]; P.row_chain = cand[q].i; P.n = cand[q].n; P.slot = cand[q].slot; P.round = ridx;
                parts.push_back(P);
            }
            rounds.push_back(R);
            pos += cnt;
        }
        // what is left: runs of consecutive chains (consecutive slots) become systolic units, as build_all_units makes them
        std::sort(rest.begin(), rest.end(), [](const Cand &x, const Cand &y) { return x.i < y.i; });
        size_t q = 0;
        while (q < rest.size()) {
            HostUnit h{};
            h.u.row_chain0 = rest[q].i; h.u.row_base = c->offsets[rest[q].i];
            h.u.col_chain = j; h.u.col_base = bu[k].col_base; h.u.m = m; h.u.pair_base = rest[q].slot;
            int cnt = 0, maxn = 0;
            long long G = 0;
            while (q < rest.size() && cnt < MAX_PAIRS_PER_UNIT && rest[q].i == h.u.row_chain0 + cnt && rest[q].slot == h.u.pair_base + cnt) {
                if (cnt > 0 && G + rest[q].n > UNIT_ROWS_MAX) break;
                G += rest[q].n; maxn = std::max(maxn, rest[q].n); ++cnt; ++q;
            }
            h.u.G = (int)G; h.u.n_pairs = cnt; h.u.path_stride = maxn + m;
            finish_unit(c, h, CRT_FP32);
            h.u.tb_base = (long long)tb_n; tb_n += (size_t)h.u.n_strips * h.u.tchunks * 32;
            h.u.path_base = (long long)path_n; path_n += (size_t)cnt * h.u.path_stride;
            h.u.bnd_base = (long long)bnd_n; if (h.multi) bnd_n += (size_t)h.u.tchunks * 4 + 8;
            h.u.dense_base = lf_dense; lf_dense += cnt;
            left.push_back(h.u);
        }
        k = k2;
    }
    b.tc_nr = rounds.size() - b.tc_r0; b.tc_np = parts.size() - b.tc_p0; b.lf_nu = left.size() - b.lf_u0; b.lf_dense = lf_dense;
    b.tb_n = tb_n; b.path_n = path_n; b.tc_bnd_n = bnd_n;
}

template <int RS>
int launch_fill1_tc(const TcFill1Args &a, int n_rounds, int sm_count, cudaStream_t st)
{
#ifdef TC_SPLIT3
    constexpr int K = ((6 * RS + 7) / 8) * 8;
#else
    constexpr int K = ((4 * RS + 7) / 8) * 8;
#endif
    constexpr size_t smem = (size_t)(2 * TC_LANES + TC_SC) * K * 4;
    static bool configured = false;
    if (!configured) {
        CU(cudaFuncSetAttribute(k_fill1_tc<RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU(cudaFuncSetAttribute(k_fill1_tc<RS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    const int grid = std::min(n_rounds, 2 * sm_count);
    k_fill1_tc<RS><<<grid, TC_THREADS, smem, st>>>(a, n_rounds);
    CU(cudaGetLastError());
    return 0;
}

int run_units(crt_ctx *c, const crt_params *prm, std::vector<HostUnit> &units, long long n_pairs, PathSink *sink,
              bool use_cached_plan = false, bool store_plan = false)
{
    const int prec = prm->precision;
    const bool f32 = prec == CRT_FP32;
    const size_t tsz = f32 ? 4 : 8;
    const size_t row2sz = f32 ? 16 : 32;
    const int rs32 = ((c->D + 2 + 3) / 4) * 4;
    const bool want_paths = sink && sink->want;
    // Protein.score_function(flexible=True) (multiple_alignment.py:323-326): the score matrix is the tensor Gaussian alone, so the
    // pair score is the stage-1 Smith-Waterman score; no traceback, no superposition, no stage 2
    const bool flexible = (prm->flags & CRT_FLEXIBLE) != 0;
    const int NS = want_paths ? 1 : env_streams();
    const bool pipe = !want_paths && NS > 1 && env_pipe() && !flexible;
    const int NW = pipe ? 4 : NS;                 // workspace sets in flight
    // stage 1 of the fp32 mode on the tensor cores (crt_fill_tc.cuh) wherever a column chain has enough partners for a round
    const bool use_tc = f32 && !flexible && !c->stage1_only && env_tc() && (c->D == 10 || c->D == 16);
    int rc;
    if ((rc = c->score.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->score1.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->rmsd.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->tm.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->ncommon.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->status.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->pair_istar.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->pair_zflag.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->path_len.ensure((size_t)n_pairs + 1))) return rc;
    if ((rc = c->xform.ensure(((size_t)n_pairs + 1) * XF))) return rc;
    if (flexible) {
        if (want_paths) return fail(CRT_E_ARG, "flexible scoring has no alignment paths (the reference only takes the score)");
        CU(cudaMemsetAsync(c->rmsd.p, 0, sizeof(double) * (size_t)n_pairs, c->stream));
        CU(cudaMemsetAsync(c->tm.p, 0, sizeof(double) * (size_t)n_pairs, c->stream));
        CU(cudaMemsetAsync(c->ncommon.p, 0, sizeof(int) * (size_t)n_pairs, c->stream));
        CU(cudaMemsetAsync(c->status.p, 0, sizeof(int) * (size_t)n_pairs, c->stream));
    }
    if (want_paths) { sink->a1.assign((size_t)n_pairs, {}); sink->a2.assign((size_t)n_pairs, {}); }

    // group by kernel variant so that each launch is homogeneous (inside a group the order is kept), then carve batches:
    // same (C, multi), bounded workspace; pipeline mode aims at env_batches() batches, the stream-per-batch mode at two per stream
    auto carve = [&](std::vector<HostUnit> &uv, int uprec, bool main_run, std::vector<Batch> &bout, std::vector<Unit> &hout) {
        const size_t tsz_u = uprec == CRT_FP32 ? 4 : 8, row2sz_u = uprec == CRT_FP32 ? 16 : 32;
        std::stable_sort(uv.begin(), uv.end(), [](const HostUnit &a, const HostUnit &b) {
            if (a.multi != b.multi) return a.multi < b.multi;
            return a.C < b.C;
        });
        size_t total_bytes = 0;
        auto unit_bytes = [&](const HostUnit &h) {
            return (size_t)h.u.n_strips * h.u.tchunks * 32 * 16 + (size_t)h.u.G * row2sz_u + (size_t)h.u.n_pairs * h.u.path_stride * 4 +
                   (h.multi ? ((size_t)h.u.tchunks * 4 + 8) * tsz_u : 0);
        };
        for (auto &h : uv) total_bytes += unit_bytes(h);
        size_t budget = env_budget();
        budget = std::min(budget, std::max<size_t>(c->mem_total / 20, (size_t)64 << 20));
        if (main_run) {
            // 2 % of slack: the units do not cut the total into exactly equal parts, and a ninth batch of a few units would only
            // lengthen the tail of the pipeline
            if (pipe) {
                double cells_all = 0;
                for (auto &h : uv) cells_all += h.cost;
                const size_t nbat = (size_t)auto_batches(cells_all, c->N > 0 ? (double)c->total / c->N : 300.0);
                budget = std::min(budget, std::max<size_t>(total_bytes / nbat + total_bytes / (50 * nbat) + 1, (size_t)64 << 20));
            }
            else if (NS > 1) budget = std::min(budget, std::max<size_t>(total_bytes / (2 * NS) + 1, (size_t)64 << 20));
        }
        hout.resize(uv.size());
        bout.clear();
        size_t pos = 0;
        while (pos < uv.size()) {
            Batch b{pos, 0, uv[pos].C, uv[pos].multi, 0, 0, 0, 0, 0};
            size_t end = pos;
            while (end < uv.size() && uv[end].C == b.C && uv[end].multi == b.multi) {
                HostUnit &h = uv[end];
                const size_t tb_u = (size_t)h.u.n_strips * h.u.tchunks * 32, rows_u = (size_t)h.u.G;
                const size_t path_u = (size_t)h.u.n_pairs * h.u.path_stride, bnd_u = b.multi ? (size_t)h.u.tchunks * 4 + 8 : 0;
                const size_t bytes = (b.tb_n + tb_u) * 16 + (b.rows2_n + rows_u) * row2sz_u + (b.path_n + path_u) * 4 + (b.bnd_n + bnd_u) * tsz_u;
                if (end > pos && bytes > budget) break;
                h.u.tb_base = (long long)b.tb_n; h.u.rows2_base = (long long)b.rows2_n;
                h.u.path_base = (long long)b.path_n; h.u.bnd_base = (long long)b.bnd_n;
                h.u.dense_base = b.n_dense; b.n_dense += h.u.n_pairs;
                h.u.s_base = (long long)b.s_n; b.s_n += (size_t)h.u.G * (size_t)h.u.m;
                b.max_tiles = std::max(b.max_tiles, (long long)((h.u.G + 15) / 16) * ((h.u.m + 63) / 64));
                b.max_strips = std::max(b.max_strips, h.u.n_strips);
                if (h.u.n_pairs != 1) b.single = false;
                b.tb_n += tb_u; b.rows2_n += rows_u; b.path_n += path_u; b.bnd_n += bnd_u;
                hout[end] = h.u;
                ++end;
            }
            b.count = end - pos;
            bout.push_back(b);
            pos = end;
        }
    };
    std::vector<Batch> batches;
    std::vector<Unit> hu;
    std::vector<TcRound> tcr;
    std::vector<TcPartner> tcp;
    std::vector<Unit> tcl;
    if (use_cached_plan) {
        batches = c->plan.batches;
        hu = c->plan.hu;
    } else {
    carve(units, prec, true, batches, hu);
    if (use_tc)
        for (auto &b : batches) plan_tc_batch(c, hu.data() + b.first, b.count, b, tcr, tcp, tcl);
    if (store_plan) { c->plan.batches = batches; c->plan.hu = hu; c->plan.tc_rounds = tcr; c->plan.tc_partners = tcp; c->plan.tc_left = tcl; }
    }
    const bool tc_on = !batches.empty() && batches[0].tc;
    // ---- size the workspaces once (no allocation inside the timed region after the first run of a shape)
    for (int w = 0; w < NW; ++w) {
        crt_ctx::Workspace &ws = c->ws[w];
        size_t tb_n = 0, rows2_n = 0, bnd_n = 0, path_n = 0, tc_bnd_n = 0;
        for (size_t k = w; k < batches.size(); k += NW) {
            tb_n = std::max(tb_n, batches[k].tb_n); rows2_n = std::max(rows2_n, batches[k].rows2_n);
            bnd_n = std::max(bnd_n, batches[k].bnd_n); path_n = std::max(path_n, batches[k].path_n);
            tc_bnd_n = std::max(tc_bnd_n, batches[k].tc_bnd_n);
        }
        if ((rc = ws.tb.ensure(tb_n + 1))) return rc;
        if ((rc = ws.rows2.ensure((rows2_n + 2 * ROW_PAD) * row2sz))) return rc;
        if ((rc = ws.path.ensure(path_n + 1))) return rc;
        if ((rc = ws.bnd.ensure(std::max(bnd_n * tsz, tc_bnd_n * 4) + 16))) return rc;
        if ((rc = ws.bnd2.ensure(bnd_n * tsz + 16))) return rc;
    }
    if ((rc = c->d_units.ensure(hu.size()))) return rc;
    c->tc_pairs = 0;
    if (tc_on) {
        for (auto &b : batches) c->tc_pairs += (long long)b.tc_np;
        if ((rc = c->d_tc_counter.ensure(batches.size()))) return rc;
        CU(cudaMemsetAsync(c->d_tc_counter.p, 0, sizeof(int) * batches.size(), c->stream));
        if (!use_cached_plan) {
            if ((rc = c->d_tc_rounds.ensure(tcr.size() + 1))) return rc;
            if ((rc = c->d_tc_partners.ensure(tcp.size() + 1))) return rc;
            if ((rc = c->d_tc_left.ensure(tcl.size() + 1))) return rc;
            CU(cudaMemcpyAsync(c->d_tc_rounds.p, tcr.data(), sizeof(TcRound) * tcr.size(), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(c->d_tc_partners.p, tcp.data(), sizeof(TcPartner) * tcp.size(), cudaMemcpyHostToDevice, c->stream));
            CU(cudaMemcpyAsync(c->d_tc_left.p, tcl.data(), sizeof(Unit) * tcl.size(), cudaMemcpyHostToDevice, c->stream));
            CU(cudaStreamSynchronize(c->stream));        // the host vectors go out of scope with this call
        }
    }

    c->tb_bytes = 0;
    for (auto &b : batches) c->tb_bytes += (double)b.tb_n * 16.0;
    c->launches = 0;
    c->last_streams = NS;
    for (int k = 0; k < 4; ++k) c->phase_ms[k] = 0;
    CU(cudaEventRecord(c->ev0, c->stream));
    if (!use_cached_plan) CU(cudaMemcpyAsync(c->d_units.p, hu.data(), sizeof(Unit) * hu.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));

    // ---- the three stages of one batch (dunits / f32x: the main run's units in the run's precision, or the float64 re-run
    //      of the pairs the fp32 traceback marked)
    const TieArgs tie = env_tie();
    // windowed re-run (crt_kernels.cuh: TieInfo): the float64 re-run of a marked pair fills only the rows and columns up to the cell
    // where the fp32 walk first met a marked decision and resumes the walk there (CARETTA_B200_RERUN_WINDOW=0: the whole pair)
    const bool window_rr = f32 && !flexible && env_tie_rerun() && !c->stage1_only &&
                           !(getenv("CARETTA_B200_RERUN_WINDOW") && atoi(getenv("CARETTA_B200_RERUN_WINDOW")) == 0);
    const Unit *DU_main = c->d_units.p;
    if (window_rr) {
        // pool: the prefixes of up to ~5 % of the pairs at the longest path (what does not fit is re-run whole)
        const size_t want = std::max<size_t>((size_t)1 << 22, (size_t)((double)n_pairs * 0.05 * 2.0 * (double)c->max_len));
        if ((rc = c->tie_info.ensure((size_t)n_pairs + 1))) return rc;
        if ((rc = c->tie_pool.ensure(want))) return rc;
        if ((rc = c->tie_pool_used.ensure(1))) return rc;
        CU(cudaMemsetAsync(c->tie_pool_used.p, 0, sizeof(unsigned long long), c->stream));
        CU(cudaMemsetAsync(c->tie_info.p, 0, sizeof(TieInfo) * (size_t)n_pairs, c->stream));      // (pairs traced by k_trace_tc: no window)
    }
    // node contexts (stage 1 only, float64): the fill reads precomputed scores (CARETTA_B200_NODE_RING=0: the generic parity kernel)
    // (worth it where the level waits for single warps: up to CARETTA_B200_NODE_RING_MAX units per batch, default 128)
    const int node_ring_max = getenv("CARETTA_B200_NODE_RING_MAX") ? atoi(getenv("CARETTA_B200_NODE_RING_MAX")) : 128;
    const bool node_ring = !f32 && c->stage1_only && !want_paths && !(getenv("CARETTA_B200_NODE_RING") && atoi(getenv("CARETTA_B200_NODE_RING")) == 0);
    // the float64 re-run overlaps the last batches of the stage pipeline: the main run's stage 2 leaves marked pairs alone and the
    // re-run's tracebacks keep ST_TIE up (CARETTA_B200_RERUN_OVERLAP=0: after the pipeline has drained, as in the first version)
    const bool overlap_rr = pipe && f32 && !flexible && env_tie_rerun() && !(getenv("CARETTA_B200_RERUN_OVERLAP") && atoi(getenv("CARETTA_B200_RERUN_OVERLAP")) == 0);
    int status_or_cur = 0;
    bool skip_tie_cur = false;
    auto stage1 = [&](const Batch &b, crt_ctx::Workspace &ws, cudaStream_t st, const Unit *dunits, bool f32x) -> int {
        const Unit *du = dunits + b.first;
        const int nu = (int)b.count;
        FillOut fo{};
        fo.tb = ws.tb.p; fo.pair_istar = c->pair_istar.p; fo.pair_zflag = c->pair_zflag.p;
        fo.pair_score = flexible ? c->score.p : c->score1.p; fo.bnd = ws.bnd.p;
        if (f32x && b.tc) {
            // rounds on the tensor cores, what is left of the batch on the systolic kernel (same result slots)
            if (b.tc_nr > 0) {
                TcFill1Args ta{};
                ta.rec = c->rec32.p + (size_t)ROW_PAD * rs32; ta.rounds = c->d_tc_rounds.p + b.tc_r0; ta.partners = c->d_tc_partners.p + b.tc_p0;
                ta.tb = ws.tb.p; ta.bnd = reinterpret_cast<float *>(ws.bnd.p); ta.pair_istar = c->pair_istar.p; ta.pair_zflag = c->pair_zflag.p;
                ta.pair_score = c->score1.p; ta.counter = c->d_tc_counter.p + (&b - batches.data()); ta.tie = tie;
                int r1 = c->D == 10 ? launch_fill1_tc<12>(ta, (int)b.tc_nr, c->sm_count, st) : launch_fill1_tc<20>(ta, (int)b.tc_nr, c->sm_count, st);
                if (r1) return r1;
            }
            if (b.lf_nu > 0) {
                Fill1Args a{c->rec32.p + (size_t)ROW_PAD * rs32, c->meta.p + ROW_PAD};
                const Unit *lu = c->d_tc_left.p + b.lf_u0;
                if (c->D == 10) return launch_fill1_f32<10>(b.C, b.multi, lu, (int)b.lf_nu, a, fo, c->d_offsets.p, tie, st);
                return launch_fill1_f32<16>(b.C, b.multi, lu, (int)b.lf_nu, a, fo, c->d_offsets.p, tie, st);
            }
            return 0;
        }
        if (f32x) {
            Fill1Args a{c->rec32.p + (size_t)ROW_PAD * rs32, c->meta.p + ROW_PAD};
            if (c->D == 10) return launch_fill1_f32<10>(b.C, b.multi, du, nu, a, fo, c->d_offsets.p, tie, st);
            return launch_fill1_f32<16>(b.C, b.multi, du, nu, a, fo, c->d_offsets.p, tie, st);
        }
        if (node_ring && b.single && b.C >= 2 && b.C <= 4 && nu <= node_ring_max) {
            // progressive-alignment nodes: scores of every unit by a cell-parallel kernel, then the recurrence alone (crt_node_fill.cuh)
            int r1;
            if ((r1 = ws.svals.ensure(b.s_n + 1))) return r1;
            NodeScoreArgs sa{du, c->rec64.p, -prm->gamma_tensor, ws.svals.p, c->D};
            const unsigned gx = (unsigned)b.max_tiles;
            k_pair_scores64<<<dim3(gx, (unsigned)nu), 256, 0, st>>>(sa);
            CU(cudaGetLastError());
            // multi-strip nodes: a warp per strip, in lockstep (k_fill_s64_mw; CARETTA_B200_NODE_MW=0: the strips one after the other)
            const bool node_mw = !(getenv("CARETTA_B200_NODE_MW") && atoi(getenv("CARETTA_B200_NODE_MW")) == 0);
            const int mw_nw = std::min(b.max_strips, NODE_MW_MAX);
#define CRT_NODE_CASE(CC)                                                                        \
            case CC:                                                                             \
                if (b.multi && node_mw && mw_nw > 1) {                                           \
                    const size_t sm = (size_t)mw_nw * (NODE_PF * CC * 32 + 64) * sizeof(double); \
                    if (sm > 48 * 1024)                                                          \
                        CU(cudaFuncSetAttribute(k_fill_s64_mw<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
                    k_fill_s64_mw<CC><<<nu, 32 * mw_nw, sm, st>>>(du, nu, ws.svals.p, fo, mw_nw); \
                } else if (b.multi) k_fill_s64<CC, true><<<nu, 32, 0, st>>>(du, nu, ws.svals.p, fo); \
                else k_fill_s64<CC, false><<<nu, 32, 0, st>>>(du, nu, ws.svals.p, fo);           \
                break;
            switch (b.C) { CRT_NODE_CASE(2) CRT_NODE_CASE(3) CRT_NODE_CASE(4) }
#undef CRT_NODE_CASE
            CU(cudaGetLastError());
            return 0;
        }
        if (c->D == 10) { P1F64<10>::Args a{c->rec64.p, c->meta.p + ROW_PAD, -prm->gamma_tensor}; return launch_fill_c<P1F64<10>, false, true, 4>(b.C, b.multi, du, nu, a, fo, st); }
        P1F64<16>::Args a{c->rec64.p, c->meta.p + ROW_PAD, -prm->gamma_tensor};
        return launch_fill_c<P1F64<16>, false, true, 3>(b.C, b.multi, du, nu, a, fo, st);
    };
    // traceback + Kabsch, then the stage-2 row records
    // f32x: code layout and zero test of the stage-1 fill that ran; rows_f32: format of the stage-2 row records
    auto stage_trace = [&](const Batch &b, crt_ctx::Workspace &ws, cudaStream_t st, cudaEvent_t mid, const Unit *dunits, bool f32x,
                           bool do_trace = true, bool do_rows2 = true, bool rows_f32 = false) -> int {
        const Unit *du = dunits + b.first;
        const int nu = (int)b.count;
        if (flexible) { if (mid) CU(cudaEventRecord(mid, st)); return 0; }
        const bool r32 = f32x || rows_f32;
        const size_t r2sz = r32 ? 16 : 32;
        TraceArgs ta{};
        ta.units = du; ta.tb = ws.tb.p; ta.pair_istar = c->pair_istar.p; ta.pair_zflag = c->pair_zflag.p;
        ta.offsets = c->d_offsets.p; ta.coords = c->coords.p; ta.centroid = c->centroid.p;
        ta.path = ws.path.p; ta.path_len = c->path_len.p; ta.rmsd = c->rmsd.p; ta.tm = c->tm.p;
        ta.ncommon = c->ncommon.p; ta.status = c->status.p; ta.xform = c->xform.p; ta.meta = c->meta.p + ROW_PAD;
        ta.rows2 = ws.rows2.p + (size_t)ROW_PAD * r2sz;
        ta.rec32 = c->rec32.p + (size_t)ROW_PAD * rs32; ta.rs32 = rs32; ta.d32 = c->D;
        ta.rec64 = c->rec64.p; ta.d64 = c->D; ta.neg_gamma_t = -prm->gamma_tensor;
        ta.scale2 = (float)std::sqrt(prm->gamma_coords * 1.4426950408889634);
        ta.precision = f32x ? CRT_FP32 : CRT_FP64;
        ta.rows2_f32 = r32 ? 1 : 0;
        ta.skip_byproducts = (c->stage1_only || c->no_byproducts) ? 1 : 0;
        ta.status_or = status_or_cur;
        if (window_rr) {
            ta.tie_pool = c->tie_pool.p; ta.tie_pool_used = c->tie_pool_used.p; ta.tie_pool_cap = (unsigned long long)c->tie_pool.cap;
            if (f32x && dunits == DU_main) ta.tie_out = c->tie_info.p;                      // the main run's fp32 traceback
            if (!f32x && dunits == c->d_units2.p && do_trace) ta.tie_in = c->tie_info.p;      // the re-run's float64 traceback
        }
        int r2;
        if (do_trace && f32x && b.tc) {
            if (b.tc_np > 0) {
                ta.tc_rounds = c->d_tc_rounds.p + b.tc_r0; ta.tc_partners = c->d_tc_partners.p + b.tc_p0;
                k_trace_tc<<<(unsigned)((b.tc_np + TRACE_THREADS - 1) / TRACE_THREADS), TRACE_THREADS, 0, st>>>(ta, (int)b.tc_np);
                CU(cudaGetLastError());
            }
            if (b.lf_nu > 0) {
                TraceArgs tl = ta;
                tl.units = c->d_tc_left.p + b.lf_u0;
                if ((r2 = launch_trace(b.C, tl, (int)b.lf_nu, b.lf_dense, st, true))) return r2;
            }
        } else
        if (do_trace && (r2 = launch_trace(b.C, ta, nu, b.n_dense, st, f32x, c->stage1_only || dunits == c->d_units2.p))) return r2;
        if (mid) CU(cudaEventRecord(mid, st));
        if (c->stage1_only || !do_rows2) return 0;
        k_rows2<<<nu, 256, 0, st>>>(ta, nu);
        CU(cudaGetLastError());
        return 0;
    };
    auto stage2 = [&](const Batch &b, crt_ctx::Workspace &ws, cudaStream_t st, const Unit *dunits, bool f32x) -> int {
        const Unit *du = dunits + b.first;
        const int nu = (int)b.count;
        FillOut fo{};
        fo.tb = ws.tb.p; fo.pair_istar = c->pair_istar.p; fo.pair_zflag = c->pair_zflag.p;
        fo.pair_score = c->score.p; fo.bnd = pipe ? ws.bnd2.p : ws.bnd.p;
        fo.skip_status = skip_tie_cur ? c->status.p : nullptr;
        if (c->stage1_only || flexible) return 0;
        if (f32x) {
            Fill2Args a{reinterpret_cast<const float4 *>(ws.rows2.p) + ROW_PAD, c->cols2.p};
            return launch_fill2_f32(b.C, b.multi, du, nu, a, fo, c->d_offsets.p, st);
        }
        P2F64::Args a{reinterpret_cast<const double *>(ws.rows2.p) + (size_t)ROW_PAD * 4, c->coords.p, -prm->gamma_coords};
        return launch_fill_c<P2F64, false, false, 4>(b.C, b.multi, du, nu, a, fo, st);
    };
    // the paths of one finished batch, copied back for the caller (tests, progressive-alignment nodes)
    auto fetch_paths = [&](const Batch &b, crt_ctx::Workspace &ws, cudaStream_t st, const std::vector<Unit> &hunits) -> int {
        std::vector<short2> hp(b.path_n + 1);
        std::vector<int> hl((size_t)n_pairs);
        CU(cudaMemcpyAsync(hp.data(), ws.path.p, b.path_n * sizeof(short2), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hl.data(), c->path_len.p, (size_t)n_pairs * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        auto put = [&](int pidx, const short2 *pp) {
            const int len = hl[pidx];
            auto &v1 = sink->a1[pidx];
            auto &v2 = sink->a2[pidx];
            v1.resize(len); v2.resize(len);
            for (int k2 = 0; k2 < len; ++k2) { v1[k2] = pp[len - 1 - k2].x; v2[k2] = pp[len - 1 - k2].y; }
        };
        if (b.tc && &hunits == &hu) {
            // tensor-core batch of the main run: the paths lie where the partner records / left-over units say
            for (size_t k = b.tc_p0; k < b.tc_p0 + b.tc_np; ++k) put(tcp[k].slot, hp.data() + tcp[k].path_base);
            for (size_t k = b.lf_u0; k < b.lf_u0 + b.lf_nu; ++k)
                for (int q = 0; q < tcl[k].n_pairs; ++q) put(tcl[k].pair_base + q, hp.data() + tcl[k].path_base + (size_t)q * tcl[k].path_stride);
            return 0;
        }
        for (size_t k = b.first; k < b.first + b.count; ++k) {
            const Unit &u = hunits[k];
            for (int q = 0; q < u.n_pairs; ++q) put(u.pair_base + q, hp.data() + u.path_base + (size_t)q * u.path_stride);
        }
        return 0;
    };
    const Unit *DU = c->d_units.p;
    auto batch_launches = [&](const Batch &b) -> long long {
        if (flexible) return 1;
        if (!b.tc) return 4;
        return 2 + 2 * ((b.tc_nr > 0 ? 1 : 0) + (b.lf_nu > 0 ? 1 : 0));
    };

    // ---- fp32 production mode: the pairs whose traceback met a decision the reference's float64 DP may take differently
    //      (ST_TIE, crt_fill1_v4.cuh) are computed again by the float64 parity kernels, one pair per unit, into the same slots
    auto rerun = [&](cudaStream_t rs, crt_ctx::Workspace &ws, bool overlapped) -> int {
        if ((rc = c->tie_list.ensure((size_t)n_pairs + 1))) return rc;
        CU(cudaMemsetAsync(c->tie_list.p, 0, sizeof(int), rs));
        k_collect_tie<<<(unsigned)((n_pairs + 255) / 256), 256, 0, rs>>>(c->status.p, n_pairs, c->tie_list.p + 1, c->tie_list.p);
        CU(cudaGetLastError());
        int n_tie = 0;
        CU(cudaMemcpyAsync(&n_tie, c->tie_list.p, sizeof(int), cudaMemcpyDeviceToHost, rs));
        // (pair_base, unit) -> the (i, j) of a result slot: sorted while the device still works on the batches -- the
        // synchronisation below returns when the last traceback is through, and from there on every host millisecond is tail
        std::vector<std::pair<int, int>> by_base(hu.size());
        for (size_t k = 0; k < hu.size(); ++k) by_base[k] = {hu[k].pair_base, (int)k};
        std::sort(by_base.begin(), by_base.end());
        CU(cudaStreamSynchronize(rs));
        if (n_tie > 0) {
            CU(cudaEventRecord(c->ev2, rs));
            std::vector<int> slots((size_t)n_tie);
            std::vector<TieInfo> winfo;
            CU(cudaMemcpyAsync(slots.data(), c->tie_list.p + 1, sizeof(int) * (size_t)n_tie, cudaMemcpyDeviceToHost, rs));
            if (window_rr) {
                if ((rc = c->tie_info_list.ensure((size_t)n_tie))) return rc;
                k_gather_tie<<<(unsigned)((n_tie + 255) / 256), 256, 0, rs>>>(c->tie_info.p, c->tie_list.p + 1, n_tie, c->tie_info_list.p);
                CU(cudaGetLastError());
                winfo.resize((size_t)n_tie);
                CU(cudaMemcpyAsync(winfo.data(), c->tie_info_list.p, sizeof(TieInfo) * (size_t)n_tie, cudaMemcpyDeviceToHost, rs));
            }
            CU(cudaStreamSynchronize(rs));
            {   // ascending slots (and their windows with them)
                std::vector<int> ord((size_t)n_tie);
                for (int q = 0; q < n_tie; ++q) ord[(size_t)q] = q;
                std::sort(ord.begin(), ord.end(), [&](int x, int y) { return slots[(size_t)x] < slots[(size_t)y]; });
                std::vector<int> s2((size_t)n_tie);
                std::vector<TieInfo> w2(winfo.size());
                for (int q = 0; q < n_tie; ++q) { s2[(size_t)q] = slots[(size_t)ord[(size_t)q]]; if (!winfo.empty()) w2[(size_t)q] = winfo[(size_t)ord[(size_t)q]]; }
                slots.swap(s2); winfo.swap(w2);
            }
            std::vector<HostUnit> ru((size_t)n_tie);
            for (int q = 0; q < n_tie; ++q) {
                auto it = std::upper_bound(by_base.begin(), by_base.end(), std::make_pair(slots[q], 0x7fffffff));
                const Unit &src = hu[(size_t)(it - 1)->second];
                const int i = src.row_chain0 + (slots[q] - src.pair_base), j = src.col_chain;
                HostUnit h{};
                h.u.row_chain0 = i; h.u.row_base = c->offsets[i];
                h.u.col_chain = j; h.u.col_base = (int)c->offsets[j]; h.u.m = (int)(c->offsets[j + 1] - c->offsets[j]);
                h.u.G = (int)(c->offsets[i + 1] - c->offsets[i]); h.u.n_pairs = 1; h.u.pair_base = slots[q];
                h.u.path_stride = h.u.G + h.u.m;
                finish_unit(c, h, CRT_FP64, true);
                ru[(size_t)q] = h;
            }
            // stage 1 + traceback in float64 (the reference's decisions), stage 2 by the fp32 kernel on the float64 alignment:
            // two unit lists over the same pairs; the float64 one may be a WINDOW of the pair (rows and columns up to the marked cell)
            std::vector<HostUnit> ru_full = ru;
            if (window_rr)
                for (int q = 0; q < n_tie; ++q) {
                    const TieInfo &w = winfo[(size_t)q];
                    HostUnit &h = ru[(size_t)q];
                    if (w.i > 0 && w.i <= h.u.G && w.j > 0 && w.j <= h.u.m) {
                        h.u.G = w.i; h.u.m = w.j;            // path_stride stays the whole pair's: prefix + the resumed walk
                        finish_unit(c, h, CRT_FP64, true);
                    }
                }
            // stage 1 + traceback in float64 (the reference's decisions), stage 2 by the fp32 kernel on the float64 alignment:
            // two unit lists over the same pairs, each with the columns-per-lane / strips of its precision
            // longest pairs first: a pair is one warp's work from start to end (a 1000 x 1000 pair ~5 ms in float64), so the order of
            // the blocks decides how long the last one runs alone
            std::stable_sort(ru.begin(), ru.end(), [](const HostUnit &x, const HostUnit &y) { return x.cost > y.cost; });
            std::stable_sort(ru_full.begin(), ru_full.end(), [](const HostUnit &x, const HostUnit &y) { return x.cost > y.cost; });
            std::vector<HostUnit> ru32 = ru_full;
            for (auto &h : ru32) finish_unit(c, h, CRT_FP32, true);
            std::vector<Batch> rb, rb32;
            std::vector<Unit> rhu, rhu32;
            carve(ru, CRT_FP64, false, rb, rhu);
            carve(ru32, CRT_FP32, false, rb32, rhu32);
            size_t tb_n = 0, rows2_n = 0, bnd_n = 0, bnd32_n = 0, path_n = 0;
            for (auto &b : rb) { tb_n = std::max(tb_n, b.tb_n); bnd_n = std::max(bnd_n, b.bnd_n); path_n = std::max(path_n, b.path_n); }
            for (auto &b : rb32) { rows2_n = std::max(rows2_n, b.rows2_n); bnd32_n = std::max(bnd32_n, b.bnd_n); }
            if ((rc = ws.tb.ensure(tb_n + 1))) return rc;
            if ((rc = ws.rows2.ensure((rows2_n + 2 * ROW_PAD) * 16))) return rc;
            if ((rc = ws.path.ensure(path_n + 1))) return rc;
            if ((rc = ws.bnd.ensure(std::max(bnd_n * 8, bnd32_n * 4) + 16))) return rc;
            if ((rc = ws.bnd2.ensure(std::max(bnd_n * 8, bnd32_n * 4) + 16))) return rc;
            if ((rc = c->d_units2.ensure(rhu.size() + rhu32.size()))) return rc;
            CU(cudaMemcpyAsync(c->d_units2.p, rhu.data(), sizeof(Unit) * rhu.size(), cudaMemcpyHostToDevice, rs));
            CU(cudaMemcpyAsync(c->d_units2.p + rhu.size(), rhu32.data(), sizeof(Unit) * rhu32.size(), cudaMemcpyHostToDevice, rs));
            status_or_cur = overlapped ? ST_TIE : 0;
            skip_tie_cur = false;
            for (auto &b : rb) {
                if ((rc = stage1(b, ws, rs, c->d_units2.p, false))) return rc;
                if ((rc = stage_trace(b, ws, rs, nullptr, c->d_units2.p, false, true, false))) return rc;
                c->launches += 2;
                if (want_paths && (rc = fetch_paths(b, ws, rs, rhu))) return rc;
            }
            for (auto &b : rb32) {
                if ((rc = stage_trace(b, ws, rs, nullptr, c->d_units2.p + rhu.size(), false, false, true, true))) return rc;
                if ((rc = stage2(b, ws, rs, c->d_units2.p + rhu.size(), true))) return rc;
                c->launches += 2;
            }
            if (!overlapped) {
                k_mark_status<<<(unsigned)((n_tie + 255) / 256), 256, 0, rs>>>(c->status.p, c->tie_list.p + 1, n_tie, CRT_ST_FP64);
                CU(cudaGetLastError());
            }
            c->launches += 2;
            CU(cudaEventRecord(c->ev3, rs));
            c->rerun_pairs = n_tie;
        }
        return 0;
    };
    c->rerun_pairs = 0; c->rerun_ms = 0;

    if (pipe) {
        // ---- stage pipeline: s_f1[] run the stage-1 fills (alternating, so consecutive launches overlap at their
        //      tails), s_tr (high priority) the tracebacks, s_f2[] the stage-2 fills.  Workspace set w = k % NW is reused by batch k + NW:
        //        fill1(k)  needs tb[w] consumed            -> waits for trace(k - NW)
        //        trace(k)  needs fill1(k), rows2/path free  -> waits for fill1(k), fill2(k - NW)
        //        fill2(k)  needs trace(k)
        for (int k = 0; k < 2; ++k) {
            CU(cudaStreamWaitEvent(c->s_f1[k], c->ev1, 0));
            CU(cudaStreamWaitEvent(c->s_f2[k], c->ev1, 0));
        }
        CU(cudaStreamWaitEvent(c->s_tr, c->ev1, 0));
        // CARETTA_B200_TIMELINE=1: bracket every kernel with timing events and print the schedule (debug)
        const bool timeline = getenv("CARETTA_B200_TIMELINE") && atoi(getenv("CARETTA_B200_TIMELINE")) != 0;
        struct Mark { const char *what; size_t batch; cudaEvent_t e0, e1; };
        std::vector<Mark> marks;
        auto mark0 = [&](const char *what, size_t bi, cudaStream_t st) {
            if (!timeline) return;
            Mark mk{what, bi, nullptr, nullptr};
            cudaEventCreate(&mk.e0); cudaEventCreate(&mk.e1);
            cudaEventRecord(mk.e0, st);
            marks.push_back(mk);
        };
        auto mark1 = [&](cudaStream_t st) { if (timeline) cudaEventRecord(marks.back().e1, st); };
        for (size_t bi = 0; bi < batches.size(); ++bi) {
            const Batch &b = batches[bi];
            crt_ctx::Workspace &ws = c->ws[bi % NW];
            cudaStream_t sf1 = c->s_f1[bi & 1], sf2 = c->s_f2[bi & 1];
            if (bi >= (size_t)NW) CU(cudaStreamWaitEvent(sf1, ws.e_t, 0));
            mark0("fill1", bi, sf1);
            if ((rc = stage1(b, ws, sf1, DU, f32))) return rc;
            mark1(sf1);
            CU(cudaEventRecord(ws.e_f1, sf1));
            CU(cudaStreamWaitEvent(c->s_tr, ws.e_f1, 0));
            if (bi >= (size_t)NW) CU(cudaStreamWaitEvent(c->s_tr, ws.e_f2, 0));
            mark0("trace", bi, c->s_tr);
            if ((rc = stage_trace(b, ws, c->s_tr, nullptr, DU, f32))) return rc;
            mark1(c->s_tr);
            CU(cudaEventRecord(ws.e_t, c->s_tr));
            CU(cudaStreamWaitEvent(sf2, ws.e_t, 0));
            mark0("fill2", bi, sf2);
            skip_tie_cur = overlap_rr;
            if ((rc = stage2(b, ws, sf2, DU, f32))) return rc;
            skip_tie_cur = false;
            mark1(sf2);
            CU(cudaEventRecord(ws.e_f2, sf2));
            c->launches += batch_launches(b);
        }
        if (timeline) {
            CU(cudaDeviceSynchronize());
            for (auto &mk : marks) {
                float t0 = 0, t1 = 0;
                cudaEventElapsedTime(&t0, c->ev0, mk.e0); cudaEventElapsedTime(&t1, c->ev0, mk.e1);
                const Batch &bb = batches[mk.batch];
                double cells = 0;
                for (size_t k = bb.first; k < bb.first + bb.count; ++k) cells += (double)hu[k].G * hu[k].m;
                fprintf(stderr, "[timeline] %-6s batch %3zu C=%2d multi=%d units=%6zu Mcells=%9.1f  %9.3f -> %9.3f ms  (%7.3f)\n", mk.what, mk.batch,
                        bb.C, bb.multi, bb.count, cells * 1e-6, t0, t1, t1 - t0);
                cudaEventDestroy(mk.e0); cudaEventDestroy(mk.e1);
            }
        }
        if (overlap_rr) {
            // every traceback of the run is on s_tr: once they are through, the marked slots are known and the float64 re-run starts
            // on its own stream and buffers, beside the stage-2 fills of the last batches (which leave the marked slots alone)
            CU(cudaEventRecord(c->e_traces, c->s_tr));
            CU(cudaStreamWaitEvent(c->ws_rr.stream, c->e_traces, 0));
            if ((rc = rerun(c->ws_rr.stream, c->ws_rr, true))) return rc;
            status_or_cur = 0;
            CU(cudaEventRecord(c->e_rr, c->ws_rr.stream));
            CU(cudaStreamWaitEvent(c->stream, c->e_rr, 0));
        }
        cudaStream_t all5[5] = {c->s_f1[0], c->s_f1[1], c->s_f2[0], c->s_f2[1], c->s_tr};
        for (int k = 0; k < 5; ++k) {
            CU(cudaEventRecord(c->ws[k % crt_ctx::MAX_WS].ev[k / crt_ctx::MAX_WS], all5[k]));
            CU(cudaStreamWaitEvent(c->stream, c->ws[k % crt_ctx::MAX_WS].ev[k / crt_ctx::MAX_WS], 0));
        }
        if (overlap_rr && c->rerun_pairs > 0) {
            const int n_tie = (int)c->rerun_pairs;
            k_finish_status<<<(unsigned)((n_tie + 255) / 256), 256, 0, c->stream>>>(c->status.p, c->tie_list.p + 1, n_tie, CRT_ST_FP64);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(c->stream));
            float rms = 0;
            CU(cudaEventElapsedTime(&rms, c->ev2, c->ev3));
            c->rerun_ms = rms;
        }
    } else {
    for (int w = 0; w < NS; ++w) CU(cudaStreamWaitEvent(c->ws[w].stream, c->ev1, 0));
    for (size_t bi = 0; bi < batches.size(); ++bi) {
        const Batch &b = batches[bi];
        crt_ctx::Workspace &ws = c->ws[bi % NS];
        cudaStream_t st = ws.stream;
        const bool timed = NS == 1;
        if (timed) CU(cudaEventRecord(ws.ev[0], st));
        if ((rc = stage1(b, ws, st, DU, f32))) return rc;
        if (timed) CU(cudaEventRecord(ws.ev[1], st));
        if ((rc = stage_trace(b, ws, st, timed ? ws.ev[2] : nullptr, DU, f32))) return rc;
        if ((rc = stage2(b, ws, st, DU, f32))) return rc;
        if (timed) CU(cudaEventRecord(ws.ev[3], st));
        c->launches += batch_launches(b);

        if (timed) {
            // one stream: phase timing (and, for the tests, the paths) per batch
            CU(cudaStreamSynchronize(st));
            float m01 = 0, m12 = 0, m23 = 0;
            CU(cudaEventElapsedTime(&m01, ws.ev[0], ws.ev[1]));
            CU(cudaEventElapsedTime(&m12, ws.ev[1], ws.ev[2]));
            CU(cudaEventElapsedTime(&m23, ws.ev[2], ws.ev[3]));
            c->phase_ms[0] += m01; c->phase_ms[1] += m12; c->phase_ms[3] += m23;
            if (getenv("CARETTA_B200_TIMELINE") && atoi(getenv("CARETTA_B200_TIMELINE")) != 0) {      // serial schedule: per-variant rates
                double cells = 0;
                for (size_t k = b.first; k < b.first + b.count; ++k) cells += (double)hu[k].G * hu[k].m;
                fprintf(stderr, "[serial] batch %3zu C=%2d multi=%d units=%6zu Mcells=%9.1f  fill1 %8.3f ms (%7.1f Gcell/s)  trace %7.3f ms  rows2 + fill2 %8.3f ms (%7.1f Gcell/s)\n",
                        bi, b.C, b.multi, b.count, cells * 1e-6, m01, cells / (m01 * 1e-3) / 1e9, m12, m23, cells / (m23 * 1e-3) / 1e9);
            }
        }
        if (want_paths && (rc = fetch_paths(b, ws, st, hu))) return rc;
    }
    for (int w = 0; w < NS; ++w) {
        CU(cudaEventRecord(c->ws[w].done, c->ws[w].stream));
        CU(cudaStreamWaitEvent(c->stream, c->ws[w].done, 0));
    }
    }
    if (f32 && !flexible && env_tie_rerun() && !overlap_rr) {
        if ((rc = rerun(c->stream, c->ws[0], false))) return rc;
        if (c->rerun_pairs > 0) {
            CU(cudaStreamSynchronize(c->stream));
            float rms = 0;
            CU(cudaEventElapsedTime(&rms, c->ev2, c->ev3));
            c->rerun_ms = rms;
        }
    }
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->elapsed_ms = ms;
    c->run_pairs = n_pairs;
    return 0;
}

int check_params(const crt_ctx *c, const crt_params *prm)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (!prm) return fail(CRT_E_ARG, "null params");
    if (c->N <= 0) return fail(CRT_E_STATE, "crt_set_chains has not been called");
    if (c->coords_only) return fail(CRT_E_STATE, "the chain set was made by crt_set_coords (coordinates only): pair runs need crt_set_chains");
    if (prm->precision != CRT_FP64 && prm->precision != CRT_FP32) return fail(CRT_E_ARG, "precision must be CRT_FP64 or CRT_FP32");
    if (prm->sw_gap != 0.0) return fail(CRT_E_ARG, "sw_gap != 0 is not on the reference's pair path (multiple_alignment.py:335, :164)");
    if (!(prm->gamma_tensor >= 0) || !(prm->gamma_coords >= 0)) return fail(CRT_E_ARG, "gamma must be >= 0");
    if (prm->flags & ~CRT_FLEXIBLE) return fail(CRT_E_ARG, "unknown flags 0x%x", (unsigned)prm->flags);
    return 0;
}

}  // namespace

// =================================================================================================== C ABI
extern "C" {

const char *crt_last_error(void) { return g_err.c_str(); }
int crt_version(void) { return 100; }

int crt_create(int device, crt_ctx **out)
{
    if (!out) return fail(CRT_E_ARG, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(CRT_E_CUDA, "no CUDA device: %s (this engine has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= n) return fail(CRT_E_ARG, "device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    crt_ctx *c = new crt_ctx();
    c->device = device;
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, device));
    c->sm_count = p.multiProcessorCount;
    c->clock_khz = p.clockRate;
    c->mem_total = p.totalGlobalMem;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreate(&c->ev2));
    CU(cudaEventCreate(&c->ev3));
    for (int w = 0; w < crt_ctx::MAX_WS; ++w) {
        CU(cudaStreamCreateWithFlags(&c->ws[w].stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ws[w].done, cudaEventDisableTiming));
        for (int k = 0; k < 4; ++k) CU(cudaEventCreate(&c->ws[w].ev[k]));
        CU(cudaEventCreateWithFlags(&c->ws[w].e_f1, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ws[w].e_t, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ws[w].e_f2, cudaEventDisableTiming));
    }
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int k = 0; k < 2; ++k) {
        CU(cudaStreamCreateWithPriority(&c->s_f1[k], cudaStreamNonBlocking, prio_lo));
        CU(cudaStreamCreateWithPriority(&c->s_f2[k], cudaStreamNonBlocking, prio_lo));
    }
    CU(cudaStreamCreateWithPriority(&c->s_tr, cudaStreamNonBlocking, prio_hi));     // short latency-bound kernels go first
    CU(cudaStreamCreateWithPriority(&c->ws_rr.stream, cudaStreamNonBlocking, prio_hi));
    CU(cudaEventCreateWithFlags(&c->e_traces, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->e_rr, cudaEventDisableTiming));
    *out = c;
    return 0;
}

int crt_destroy(crt_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->coords.release(); c->tensors.release(); c->centroid.release(); c->rec64.release(); c->d_offsets.release();
    c->chain_of.release(); c->stats.release(); c->flag.release(); c->meta.release(); c->rec32.release(); c->cols2.release(); c->d_units.release();
    for (int w = 0; w < crt_ctx::MAX_WS; ++w) {
        crt_ctx::Workspace &ws = c->ws[w];
        ws.tb.release(); ws.rows2.release(); ws.bnd.release(); ws.bnd2.release(); ws.path.release(); ws.svals.release();
        if (ws.done) cudaEventDestroy(ws.done);
        if (ws.e_f1) cudaEventDestroy(ws.e_f1);
        if (ws.e_t) cudaEventDestroy(ws.e_t);
        if (ws.e_f2) cudaEventDestroy(ws.e_f2);
        for (int k = 0; k < 4; ++k) if (ws.ev[k]) cudaEventDestroy(ws.ev[k]);
        if (ws.stream) cudaStreamDestroy(ws.stream);
    }
    c->xform.release(); c->path_len.release(); c->d_pi.release(); c->d_pj.release(); c->dense.release();
    c->pair_istar.release(); c->pair_zflag.release(); c->ncommon.release(); c->status.release();
    c->score.release(); c->score1.release(); c->rmsd.release(); c->tm.release(); c->f32tmp.release();
    if (c->node_ctx) { crt_destroy(c->node_ctx); c->node_ctx = nullptr; }
    c->nd_w.release(); c->nd_S.release(); c->nd_bnd.release(); c->nd_f.release(); c->nd_score.release(); c->nd_xf2.release();
    c->nd_t.release(); c->nd_c.release(); c->nd_wm.release(); c->nd_B.release(); c->nd_a1.release(); c->nd_a2.release(); c->nd_len.release(); c->text.release();
    c->lv_probs.release(); c->lv_mult.release(); c->lv_xf2.release(); c->lv_off.release();
    c->pool.t.release(); c->pool.c.release(); c->pool.w.release(); c->lv_tab.release(); c->lv_out_off.release(); c->arena.release();
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2); cudaEventDestroy(c->ev3);
    c->d_units2.release(); c->tie_list.release(); c->lay_pi.release(); c->lay_pj.release(); c->lay_off.release();
    for (int k = 0; k < 2; ++k) {
        if (c->s_f1[k]) cudaStreamDestroy(c->s_f1[k]);
        if (c->s_f2[k]) cudaStreamDestroy(c->s_f2[k]);
    }
    if (c->s_tr) cudaStreamDestroy(c->s_tr);
    if (c->ws_rr.stream) cudaStreamDestroy(c->ws_rr.stream);
    c->ws_rr.tb.release(); c->ws_rr.rows2.release(); c->ws_rr.bnd.release(); c->ws_rr.bnd2.release(); c->ws_rr.path.release(); c->ws_rr.svals.release();
    if (c->e_traces) cudaEventDestroy(c->e_traces);
    if (c->e_rr) cudaEventDestroy(c->e_rr);
    c->d_tc_rounds.release(); c->d_tc_partners.release(); c->d_tc_left.release(); c->d_tc_counter.release();
    cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

int crt_device_info(crt_ctx *c, int32_t *sm_count, int32_t *clock_khz, int64_t *mem_bytes)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (sm_count) *sm_count = c->sm_count;
    if (clock_khz) *clock_khz = c->clock_khz;
    if (mem_bytes) *mem_bytes = (int64_t)c->mem_total;
    return 0;
}

}  // extern "C"

namespace {

// crt_set_chains in two halves, so that a chain set can also be assembled on the device (crt_msa_level gathers the children of a
// tree level from the sequence pool): chains_prepare validates the lengths and sizes every buffer, the caller fills
// c->coords / c->tensors (host copy or device gather, on c->stream), chains_finish derives the tables and checks finiteness.
int chains_prepare(crt_ctx *c, const int64_t *offsets, int32_t n_chains, int32_t d)
{
    if (n_chains <= 0) return fail(CRT_E_ARG, "n_chains must be > 0");
    const int D = pad_dim(d);
    if (d <= 0 || D < 0) return fail(CRT_E_ARG, "tensor width d=%d unsupported (1..16)", d);
    if (offsets[0] != 0) return fail(CRT_E_ARG, "offsets[0] must be 0");
    int max_len = 0;
    for (int p = 0; p < n_chains; ++p) {
        long long l = offsets[p + 1] - offsets[p];
        if (l <= 0) return fail(CRT_E_ARG, "chain %d is empty (the reference cannot align an empty chain)", p);
        if (l > 32000) return fail(CRT_E_ARG, "chain %d has %lld residues; the path buffer is 16-bit (max 32000)", p, l);
        max_len = std::max<int>(max_len, (int)l);
    }
    const long long total = offsets[n_chains];
    if (total >= (1ll << 31) - 64) return fail(CRT_E_ARG, "too many residues");
    CU(cudaSetDevice(c->device));
    c->N = 0;
    int rc;
    if ((rc = c->coords.ensure((size_t)total * 3))) return rc;
    if ((rc = c->tensors.ensure((size_t)total * d))) return rc;
    if ((rc = c->d_offsets.ensure((size_t)n_chains + 1))) return rc;
    if ((rc = c->chain_of.ensure((size_t)total))) return rc;
    if ((rc = c->meta.ensure((size_t)total + 2 * ROW_PAD))) return rc;
    if ((rc = c->centroid.ensure((size_t)n_chains * 3))) return rc;
    if ((rc = c->rec64.ensure((size_t)total * D))) return rc;
    if ((rc = c->rec32.ensure(((size_t)total + 2 * ROW_PAD) * (((D + 2 + 3) / 4) * 4)))) return rc;
    CU(cudaMemsetAsync(c->meta.p, 0, sizeof(int) * ((size_t)total + 2 * ROW_PAD), c->stream));
    CU(cudaMemsetAsync(c->rec32.p, 0, sizeof(float) * ((size_t)total + 2 * ROW_PAD) * (((D + 2 + 3) / 4) * 4), c->stream));
    if ((rc = c->cols2.ensure((size_t)total))) return rc;
    if ((rc = c->stats.ensure((size_t)(STATS_BLOCKS + 1) * 32))) return rc;
    if ((rc = c->flag.ensure(1))) return rc;
    c->offsets.assign(offsets, offsets + n_chains + 1);
    c->max_len = max_len;
    CU(cudaMemsetAsync(c->flag.p, 0, sizeof(int), c->stream));
    CU(cudaMemcpyAsync(c->d_offsets.p, c->offsets.data(), sizeof(long long) * ((size_t)n_chains + 1), cudaMemcpyHostToDevice, c->stream));
    return 0;
}

int chains_finish(crt_ctx *c, int32_t n_chains, int32_t d)
{
    const int D = pad_dim(d);
    const long long total = c->offsets[(size_t)n_chains];
    // chain index table, global tensor mean (the Gaussian is translation invariant; centring keeps the fp32 dot-product
    // form accurate) and the finiteness check run on the device; the only host wait is for the 4-byte flag
    {
        PrepArgs a{};
        a.offsets = c->d_offsets.p; a.chain_of = c->chain_of.p; a.n_chains = n_chains; a.total = total;
        k_chain_of<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(a);
        CU(cudaGetLastError());
        k_tensor_stats<<<STATS_BLOCKS, STATS_THREADS, 0, c->stream>>>(c->tensors.p, c->coords.p, total, d, c->stats.p, c->flag.p);
        CU(cudaGetLastError());
        k_tensor_mean<<<1, 32, 0, c->stream>>>(c->stats.p, total, d, c->stats.p + (size_t)STATS_BLOCKS * 32);
        CU(cudaGetLastError());
    }
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, c->flag.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (bad) return fail(CRT_E_ARG, "non-finite value in coords/tensors");
    {
        unsigned long long h = 1469598103934665603ull;
        for (int p = 0; p <= n_chains; ++p) { h ^= (unsigned long long)c->offsets[(size_t)p]; h *= 1099511628211ull; }
        h ^= (unsigned long long)D; h *= 1099511628211ull;
        c->offsets_hash = h;
    }
    c->N = n_chains; c->d = d; c->D = D; c->total = total;
    c->prep_gamma_t = c->prep_gamma_c = -1;
    return 0;
}

}  // namespace

extern "C" {

int crt_set_chains(crt_ctx *c, const double *coords, const double *tensors, const int64_t *offsets, int32_t n_chains, int32_t d)
{
    if (!c || !coords || !tensors || !offsets) return fail(CRT_E_ARG, "null argument");
    int rc = chains_prepare(c, offsets, n_chains, d);
    if (rc) return rc;
    const long long total = offsets[n_chains];
    CU(cudaMemcpyAsync(c->coords.p, coords, sizeof(double) * (size_t)total * 3, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->tensors.p, tensors, sizeof(double) * (size_t)total * d, cudaMemcpyHostToDevice, c->stream));
    c->coords_only = false;
    return chains_finish(c, n_chains, d);
}

/* Coordinates only (the consumers of the alignment -- crt_superpose*, crt_rmsd_cov_tm* -- never read the shape tensors): the chain
 * set gets a one-column zero tensor on the device instead of an upload of [sum L, d] doubles. */
int crt_set_coords(crt_ctx *c, const double *coords, const int64_t *offsets, int32_t n_chains)
{
    if (!c || !coords || !offsets) return fail(CRT_E_ARG, "null argument");
    int rc = chains_prepare(c, offsets, n_chains, 1);
    if (rc) return rc;
    const long long total = offsets[n_chains];
    CU(cudaMemcpyAsync(c->coords.p, coords, sizeof(double) * (size_t)total * 3, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->tensors.p, 0, sizeof(double) * (size_t)total, c->stream));
    c->coords_only = true;
    return chains_finish(c, n_chains, 1);
}

}  // extern "C"

namespace {

int ensure_prepared(crt_ctx *c, const crt_params *prm)
{
    if (!prm) return 0;
    if (c->prep_gamma_t == prm->gamma_tensor && c->prep_gamma_c == prm->gamma_coords) return 0;
    PrepArgs a{};
    a.coords = c->coords.p; a.tensors = c->tensors.p; a.offsets = c->d_offsets.p; a.chain_of = c->chain_of.p;
    a.n_chains = c->N; a.d = c->d; a.total = c->total;
    a.mean = c->stats.p + (size_t)STATS_BLOCKS * 32;
    a.g2 = prm->gamma_tensor * 1.4426950408889634;
    a.scale2 = std::sqrt(prm->gamma_coords * 1.4426950408889634);
    a.rs32 = ((c->D + 2 + 3) / 4) * 4; a.d32 = c->D; a.rec32 = c->rec32.p + (size_t)ROW_PAD * a.rs32;
    a.rec64 = c->rec64.p; a.d64 = c->D;
    a.meta = c->meta.p + ROW_PAD; a.cols2 = c->cols2.p; a.centroid = c->centroid.p;
    k_centroid<<<(c->N + 3) / 4, 128, 0, c->stream>>>(a);
    CU(cudaGetLastError());
    k_prep<<<(unsigned)((c->total + 127) / 128), 128, 0, c->stream>>>(a);
    CU(cudaGetLastError());
    c->prep_gamma_t = prm->gamma_tensor; c->prep_gamma_c = prm->gamma_coords;
    return 0;
}

}  // namespace

extern "C" {

int crt_pairwise_shard(crt_ctx *c, const crt_params *prm, int32_t rank, int32_t world)
{
    int rc = check_params(c, prm);
    if (rc) return rc;
    if (world < 1 || rank < 0 || rank >= world) return fail(CRT_E_ARG, "bad rank/world %d/%d", rank, world);
    CU(cudaSetDevice(c->device));
    if ((rc = ensure_prepared(c, prm))) return rc;
    const int ns = env_streams() + 100 * (env_pipe() ? 1 : 0) + 1000 * env_batches() + 100000 * (env_unit_rows() / 64) + (env_tc() ? 50 + 10000000 * env_tc_min_part() : 0);
    const size_t budget = env_budget();
    crt_ctx::PlanCache &pc = c->plan;
    if (pc.valid && pc.offsets_hash == c->offsets_hash && pc.rank == rank && pc.world == world && pc.prec == prm->precision &&
        pc.ns == ns && pc.budget == budget) {
        std::vector<HostUnit> none;
        c->cell_updates = pc.cells;
        if (pc.n_pairs == 0) { c->run_pairs = 0; c->elapsed_ms = 0; c->launches = 0; return 0; }
        return run_units(c, prm, none, pc.n_pairs, nullptr, true, false);
    }
    std::vector<HostUnit> units;
    build_all_units(c, prm->precision, world, units);
    shard_units(units, rank, world);
    long long np = 0;
    assign_pairs(units, &c->run_pi, &c->run_pj, &np, &c->cell_updates, c);
    pc.valid = false;
    if (np == 0) { c->run_pairs = 0; c->elapsed_ms = 0; c->launches = 0; return 0; }
    if ((rc = c->d_pi.ensure((size_t)np))) return rc;
    if ((rc = c->d_pj.ensure((size_t)np))) return rc;
    CU(cudaMemcpyAsync(c->d_pi.p, c->run_pi.data(), sizeof(int) * (size_t)np, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_pj.p, c->run_pj.data(), sizeof(int) * (size_t)np, cudaMemcpyHostToDevice, c->stream));
    rc = run_units(c, prm, units, np, nullptr, false, true);
    if (rc == 0) {
        pc.valid = true; pc.offsets_hash = c->offsets_hash; pc.rank = rank; pc.world = world; pc.prec = prm->precision;
        pc.ns = ns; pc.budget = budget; pc.n_pairs = np; pc.cells = c->cell_updates;
    }
    return rc;
}

int64_t crt_shard_size(crt_ctx *c, int32_t rank, int32_t world)
{
    if (!c || c->N <= 0) return fail(CRT_E_STATE, "no chains");
    if (world < 1 || rank < 0 || rank >= world) return fail(CRT_E_ARG, "bad rank/world");
    std::vector<HostUnit> units;
    build_all_units(c, CRT_FP32, world, units);     // the enumeration does not depend on the precision
    shard_units(units, rank, world);
    long long np = 0;
    assign_pairs(units, nullptr, nullptr, &np, nullptr, c);
    return np;
}

int crt_shard_pairs(crt_ctx *c, int32_t rank, int32_t world, int32_t *pair_i, int32_t *pair_j)
{
    if (!c || c->N <= 0) return fail(CRT_E_STATE, "no chains");
    if (!pair_i || !pair_j) return fail(CRT_E_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(CRT_E_ARG, "bad rank/world");
    std::vector<HostUnit> units;
    build_all_units(c, CRT_FP32, world, units);
    shard_units(units, rank, world);
    std::vector<int> pi, pj;
    long long np = 0;
    assign_pairs(units, &pi, &pj, &np, nullptr, c);
    std::memcpy(pair_i, pi.data(), sizeof(int) * (size_t)np);
    std::memcpy(pair_j, pj.data(), sizeof(int) * (size_t)np);
    return 0;
}

// Host-only planning (no device needed): the same deterministic enumeration crt_pairwise_shard uses.
static int plan_units(const int64_t *offsets, int32_t n_chains, int32_t rank, int32_t world, std::vector<HostUnit> &units)
{
    if (!offsets || n_chains <= 0) return fail(CRT_E_ARG, "bad offsets / n_chains");
    if (world < 1 || rank < 0 || rank >= world) return fail(CRT_E_ARG, "bad rank/world %d/%d", rank, world);
    crt_ctx tmp;
    tmp.N = n_chains; tmp.D = 10;
    tmp.offsets.assign(offsets, offsets + n_chains + 1);
    build_all_units(&tmp, CRT_FP32, world, units);
    shard_units(units, rank, world);
    return 0;
}

int64_t crt_plan_shard_size(const int64_t *offsets, int32_t n_chains, int32_t rank, int32_t world)
{
    std::vector<HostUnit> units;
    int rc = plan_units(offsets, n_chains, rank, world, units);
    if (rc) return rc;
    long long np = 0;
    assign_pairs(units, nullptr, nullptr, &np, nullptr, nullptr);
    return np;
}

int crt_plan_shard_pairs(const int64_t *offsets, int32_t n_chains, int32_t rank, int32_t world, int32_t *pair_i, int32_t *pair_j)
{
    if (!pair_i || !pair_j) return fail(CRT_E_ARG, "null argument");
    std::vector<HostUnit> units;
    int rc = plan_units(offsets, n_chains, rank, world, units);
    if (rc) return rc;
    std::vector<int> pi, pj;
    long long np = 0;
    assign_pairs(units, &pi, &pj, &np, nullptr, nullptr);
    std::memcpy(pair_i, pi.data(), sizeof(int) * (size_t)np);
    std::memcpy(pair_j, pj.data(), sizeof(int) * (size_t)np);
    return 0;
}

int crt_fetch(crt_ctx *c, double *score, double *rmsd, double *tm, int32_t *ncommon, int32_t *status)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    const size_t n = (size_t)c->run_pairs;
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    if (score) CU(cudaMemcpyAsync(score, c->score.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (rmsd) CU(cudaMemcpyAsync(rmsd, c->rmsd.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (tm) CU(cudaMemcpyAsync(tm, c->tm.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (ncommon) CU(cudaMemcpyAsync(ncommon, c->ncommon.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (status) CU(cudaMemcpyAsync(status, c->status.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"

namespace {
// dense symmetric matrices from the packed per-pair results: out[i][j] = out[j][i] = v; TM diagonal 1
__global__ void k_scatter_dense(const int *pi, const int *pj, const double *s, const double *r, const double *t,
                                double *S, double *R, double *T, long long np, int N)
{
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (T && q < N) T[q * N + q] = 1.0;
    if (q >= np) return;
    const long long i = pi[q], j = pj[q];
    S[i * N + j] = S[j * N + i] = s[q];
    if (R) R[i * N + j] = R[j * N + i] = r[q];
    if (T) T[i * N + j] = T[j * N + i] = t[q];
}

__global__ void k_d2f(const double *a, float *o, long long n)
{
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) o[q] = (float)a[q];
}
}  // namespace

extern "C" {

int crt_fetch_device(crt_ctx *c, void *d_score, void *d_rmsd, void *d_tm, int64_t n)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (n < c->run_pairs) return fail(CRT_E_ARG, "destination holds %lld elements, run produced %lld", (long long)n, c->run_pairs);
    const long long np = c->run_pairs;
    if (np == 0) return 0;
    CU(cudaSetDevice(c->device));
    const unsigned grid = (unsigned)((np + 255) / 256);
    if (d_score) k_d2f<<<grid, 256, 0, c->stream>>>(c->score.p, (float *)d_score, np);
    if (d_rmsd) k_d2f<<<grid, 256, 0, c->stream>>>(c->rmsd.p, (float *)d_rmsd, np);
    if (d_tm) k_d2f<<<grid, 256, 0, c->stream>>>(c->tm.p, (float *)d_tm, np);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

double crt_last_elapsed_ms(crt_ctx *c) { return c ? c->elapsed_ms : -1.0; }
int crt_last_phase_ms(crt_ctx *c, double *out4)
{
    if (!c || !out4) return fail(CRT_E_ARG, "null argument");
    for (int k = 0; k < 4; ++k) out4[k] = c->last_streams == 1 ? c->phase_ms[k] : -1.0;
    return 0;
}
int64_t crt_last_launches(crt_ctx *c) { return c ? c->launches : -1; }
int crt_last_rerun(crt_ctx *c, int64_t *pairs, double *ms)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (pairs) *pairs = c->rerun_pairs;
    if (ms) *ms = c->rerun_ms;
    return 0;
}
int64_t crt_last_tc_pairs(crt_ctx *c) { return c ? c->tc_pairs : -1; }
double crt_last_cell_updates(crt_ctx *c) { return c ? c->cell_updates : -1.0; }
double crt_last_traceback_bytes(crt_ctx *c) { return c ? c->tb_bytes : -1.0; }

int crt_pairwise_all(crt_ctx *c, const crt_params *prm, double *out_score, double *out_rmsd, double *out_tm)
{
    if (!out_score) return fail(CRT_E_ARG, "null out_score");
    if (!c) return fail(CRT_E_ARG, "null context");
    // the reference's make_pairwise_matrix returns the scores alone (multiple_alignment.py:158-170): without RMSD / TM matrices to
    // fill, the tracebacks leave out the by-product pass over the paths (crt_fetch then reports 0 for both)
    c->no_byproducts = !out_rmsd && !out_tm;
    int rc = crt_pairwise_shard(c, prm, 0, 1);
    c->no_byproducts = false;
    if (rc) return rc;
    // symmetric fill on the device (multiple_alignment.py:164), then one copy per requested matrix
    const size_t N = (size_t)c->N, np = (size_t)c->run_pairs;
    const int nm = 1 + (out_rmsd ? 1 : 0) + (out_tm ? 1 : 0);
    if ((rc = c->dense.ensure(N * N * (size_t)nm))) return rc;
    double *dS = c->dense.p, *dR = out_rmsd ? dS + N * N : nullptr, *dT = out_tm ? dS + N * N * (size_t)(out_rmsd ? 2 : 1) : nullptr;
    CU(cudaMemsetAsync(c->dense.p, 0, sizeof(double) * N * N * (size_t)nm, c->stream));
    if (np > 0 || dT) {
        const long long work = (long long)std::max(np, N);
        k_scatter_dense<<<(unsigned)((work + 255) / 256), 256, 0, c->stream>>>(c->d_pi.p, c->d_pj.p, c->score.p, c->rmsd.p, c->tm.p,
                                                                                dS, dR, dT, (long long)np, (int)N);
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(out_score, dS, sizeof(double) * N * N, cudaMemcpyDeviceToHost, c->stream));
    if (out_rmsd) CU(cudaMemcpyAsync(out_rmsd, dR, sizeof(double) * N * N, cudaMemcpyDeviceToHost, c->stream));
    if (out_tm) CU(cudaMemcpyAsync(out_tm, dT, sizeof(double) * N * N, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

/* Page-locked host memory for the caller's input/output arrays (copies to and from it run at full PCIe rate). */
int crt_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(CRT_E_ARG, "null out pointer");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(CRT_E_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return 0;
}

int crt_host_free(void *p)
{
    if (p) cudaFreeHost(p);
    return 0;
}

int crt_pairwise_list(crt_ctx *c, const crt_params *prm, const int32_t *pair_i, const int32_t *pair_j, int64_t n_pairs,
                      double *score, double *rmsd, double *tm, int32_t *ncommon, int32_t *status,
                      int32_t *aln1, int32_t *aln2, int64_t *aln_off, int64_t aln_cap)
{
    int rc = check_params(c, prm);
    if (rc) return rc;
    if (n_pairs < 0 || (n_pairs > 0 && (!pair_i || !pair_j))) return fail(CRT_E_ARG, "bad pair list");
    const bool want_paths = aln1 || aln2 || aln_off;
    if (want_paths && !(aln1 && aln2 && aln_off)) return fail(CRT_E_ARG, "aln1, aln2 and aln_off must be given together");
    if (n_pairs == 0) { if (aln_off) aln_off[0] = 0; c->run_pairs = 0; return 0; }
    for (int64_t q = 0; q < n_pairs; ++q)
        if (pair_i[q] < 0 || pair_i[q] >= c->N || pair_j[q] < 0 || pair_j[q] >= c->N)
            return fail(CRT_E_ARG, "pair %lld = (%d, %d) out of range", (long long)q, pair_i[q], pair_j[q]);
    CU(cudaSetDevice(c->device));
    if ((rc = ensure_prepared(c, prm))) return rc;
    // sort by (column chain, row chain); runs of consecutive row chains become units
    std::vector<long long> order((size_t)n_pairs);
    for (int64_t q = 0; q < n_pairs; ++q) order[(size_t)q] = q;
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) {
        if (pair_j[a] != pair_j[b]) return pair_j[a] < pair_j[b];
        return pair_i[a] < pair_i[b];
    });
    std::vector<HostUnit> units;
    std::vector<long long> slot_of((size_t)n_pairs);      // caller index -> result slot
    long long slot = 0;
    size_t q = 0;
    while (q < order.size()) {
        HostUnit h{};
        const int j = pair_j[order[q]];
        const int i = pair_i[order[q]];
        h.u.row_chain0 = i; h.u.row_base = c->offsets[i];
        h.u.col_chain = j; h.u.col_base = (int)c->offsets[j]; h.u.m = (int)(c->offsets[j + 1] - c->offsets[j]);
        h.u.pair_base = (int)slot;
        int cnt = 0, maxn = 0;
        long long G = 0;
        while (q < order.size() && pair_j[order[q]] == j) {
            const int ii = pair_i[order[q]];
            if (cnt > 0 && ii == i + cnt - 1) { slot_of[(size_t)order[q]] = slot - 1; ++q; continue; }   // duplicate pair
            if (ii != i + cnt || cnt >= MAX_PAIRS_PER_UNIT) break;
            const int n = (int)(c->offsets[ii + 1] - c->offsets[ii]);
            if (cnt > 0 && G + n > (env_unit_rows() ? env_unit_rows() : UNIT_ROWS_MAX)) break;
            G += n; maxn = std::max(maxn, n);
            slot_of[(size_t)order[q]] = slot++;
            ++cnt; ++q;
        }
        h.u.G = (int)G; h.u.n_pairs = cnt; h.u.path_stride = maxn + h.u.m;
        finish_unit(c, h, prm->precision);
        units.push_back(h);
    }
    double cu = 0;
    for (auto &h : units) cu += 2.0 * (double)h.u.G * (double)h.u.m;
    c->cell_updates = cu;
    c->run_pi.clear(); c->run_pj.clear();
    PathSink sink;
    sink.want = want_paths;
    c->plan.valid = false;
    if ((rc = run_units(c, prm, units, slot, &sink))) return rc;
    const size_t np = (size_t)slot;
    std::vector<double> s(np), r(np), t(np);
    std::vector<int> nc(np), st(np);
    if ((rc = crt_fetch(c, s.data(), r.data(), t.data(), nc.data(), st.data()))) return rc;
    int64_t off = 0;
    for (int64_t k = 0; k < n_pairs; ++k) {
        const size_t sl = (size_t)slot_of[(size_t)k];
        if (score) score[k] = s[sl];
        if (rmsd) rmsd[k] = r[sl];
        if (tm) tm[k] = t[sl];
        if (ncommon) ncommon[k] = nc[sl];
        if (status) status[k] = st[sl];
        if (want_paths) {
            const auto &v1 = sink.a1[sl];
            const auto &v2 = sink.a2[sl];
            aln_off[k] = off;
            if (off + (int64_t)v1.size() > aln_cap) return fail(CRT_E_ARG, "aln_cap too small");
            for (size_t x = 0; x < v1.size(); ++x) { aln1[off + x] = v1[x]; aln2[off + x] = v2[x]; }
            off += (int64_t)v1.size();
        }
    }
    if (want_paths) aln_off[n_pairs] = off;
    return 0;
}

}  // extern "C"

namespace {

// Shared host driver of the two DP-in-isolation entry points.
int dp_batch(crt_ctx *c, bool affine, const double *S, const int64_t *shape_off, const int32_t *n, const int32_t *m,
             int32_t n_problems, double p0, double p1, int32_t *aln1, int32_t *aln2, int64_t *aln_off, int64_t aln_cap,
             double *score, int32_t *status)
{
    if (!c) return fail(CRT_E_ARG, "null context");
    if (n_problems < 0 || (n_problems > 0 && (!S || !shape_off || !n || !m || !score))) return fail(CRT_E_ARG, "null argument");
    const bool want_paths = aln1 || aln2 || aln_off;
    if (want_paths && !(aln1 && aln2 && aln_off)) return fail(CRT_E_ARG, "aln1, aln2 and aln_off must be given together");
    if (affine && !want_paths && false) return 0;
    if (n_problems == 0) { if (aln_off) aln_off[0] = 0; return 0; }
    CU(cudaSetDevice(c->device));
    std::vector<DpProblem> probs((size_t)n_problems);
    long long cells = 0, rows = 0, alen = 0, s_end = 0;
    for (int p = 0; p < n_problems; ++p) {
        if (n[p] <= 0 || m[p] <= 0) return fail(CRT_E_ARG, "problem %d has an empty dimension (%d x %d)", p, n[p], m[p]);
        probs[p].s_off = shape_off[p]; probs[p].b_off = cells; probs[p].bnd_off = rows; probs[p].aln_off = alen;
        probs[p].n = n[p]; probs[p].m = m[p];
        // workspace per problem: backtrack bytes [n][dtw_pitch(m)] (affine) or the H matrix [n][m] (Smith-Waterman)
        cells += (long long)n[p] * (affine ? dtw_pitch(m[p]) : m[p]); rows += n[p]; alen += (long long)n[p] + m[p] + 1;
        s_end = std::max<long long>(s_end, shape_off[p] + (long long)n[p] * m[p]);
    }
    DevBuf<double> dS, dW, dBnd, dF, dScore;
    DevBuf<unsigned char> dB;
    DevBuf<DpProblem> dP;
    DevBuf<int> dA1, dA2, dLen, dSt;
    DevBuf<long long> dIdx;
    int rc = 0;
    auto cleanup = [&]() { dS.release(); dW.release(); dBnd.release(); dF.release(); dScore.release(); dB.release(); dP.release();
                           dA1.release(); dA2.release(); dLen.release(); dSt.release(); dIdx.release(); };
#define TRY(x) do { if ((rc = (x))) { cleanup(); return rc; } } while (0)
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(CRT_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); } } while (0)
    TRY(dS.ensure((size_t)s_end)); TRY(dP.ensure((size_t)n_problems)); TRY(dBnd.ensure((size_t)rows * 2 + 2));
    TRY(dScore.ensure((size_t)n_problems)); TRY(dLen.ensure((size_t)n_problems)); TRY(dSt.ensure((size_t)n_problems));
    TRY(dA1.ensure((size_t)alen)); TRY(dA2.ensure((size_t)alen));
    if (affine) { TRY(dB.ensure((size_t)cells)); TRY(dF.ensure((size_t)n_problems * 3)); }
    else { TRY(dW.ensure((size_t)cells)); TRY(dF.ensure((size_t)n_problems)); TRY(dIdx.ensure((size_t)n_problems)); }
    cudaStream_t st = c->stream;
    CUT(cudaMemcpyAsync(dS.p, S, sizeof(double) * (size_t)s_end, cudaMemcpyHostToDevice, st));
    CUT(cudaMemcpyAsync(dP.p, probs.data(), sizeof(DpProblem) * (size_t)n_problems, cudaMemcpyHostToDevice, st));
    const int tb = 64, tg = (n_problems + tb - 1) / tb;
    if (affine) {
        int max_m = 0;
        for (const DpProblem &q : probs) max_m = std::max(max_m, q.m);
        CUT(launch_dtw_fill(dP.p, n_problems, max_m, dS.p, dB.p, dBnd.p, dF.p, p0, p1, st));
        k_dtw_trace_w<<<n_problems, 32, 0, st>>>(dP.p, n_problems, dB.p, dF.p, dA1.p, dA2.p, dLen.p, dScore.p);
    } else {
        k_sw_fill<<<n_problems, 32, 0, st>>>(dP.p, n_problems, dS.p, dW.p, dBnd.p, dF.p, dIdx.p, p0);
        k_sw_trace<<<tg, tb, 0, st>>>(dP.p, n_problems, dS.p, dW.p, dF.p, dIdx.p, dA1.p, dA2.p, dLen.p, dScore.p, dSt.p, p0, want_paths ? 1 : 0);
    }
    CUT(cudaGetLastError());
    std::vector<int> h1, h2, hl((size_t)n_problems), hs((size_t)n_problems);
    CUT(cudaMemcpyAsync(score, dScore.p, sizeof(double) * (size_t)n_problems, cudaMemcpyDeviceToHost, st));
    CUT(cudaMemcpyAsync(hl.data(), dLen.p, sizeof(int) * (size_t)n_problems, cudaMemcpyDeviceToHost, st));
    if (!affine) CUT(cudaMemcpyAsync(hs.data(), dSt.p, sizeof(int) * (size_t)n_problems, cudaMemcpyDeviceToHost, st));
    if (want_paths) {
        h1.resize((size_t)alen); h2.resize((size_t)alen);
        CUT(cudaMemcpyAsync(h1.data(), dA1.p, sizeof(int) * (size_t)alen, cudaMemcpyDeviceToHost, st));
        CUT(cudaMemcpyAsync(h2.data(), dA2.p, sizeof(int) * (size_t)alen, cudaMemcpyDeviceToHost, st));
    }
    CUT(cudaStreamSynchronize(st));
    if (status) for (int p = 0; p < n_problems; ++p) status[p] = affine ? 0 : hs[p];
    if (want_paths) {
        int64_t off = 0;
        for (int p = 0; p < n_problems; ++p) {
            aln_off[p] = off;
            if (off + hl[p] > aln_cap) { cleanup(); return fail(CRT_E_ARG, "aln_cap too small"); }
            std::memcpy(aln1 + off, h1.data() + probs[p].aln_off, sizeof(int) * (size_t)hl[p]);
            std::memcpy(aln2 + off, h2.data() + probs[p].aln_off, sizeof(int) * (size_t)hl[p]);
            off += hl[p];
        }
        aln_off[n_problems] = off;
    }
    cleanup();
#undef TRY
#undef CUT
    return 0;
}

}  // namespace

extern "C" {

int crt_sw_align_batch(crt_ctx *c, const double *S, const int64_t *shape_off, const int32_t *n, const int32_t *m,
                       int32_t n_problems, double gap, int32_t *aln1, int32_t *aln2, int64_t *aln_off, int64_t aln_cap,
                       double *score, int32_t *status)
{
    if (!(gap >= 0.0)) return fail(CRT_E_ARG, "gap must be >= 0");
    return dp_batch(c, false, S, shape_off, n, m, n_problems, gap, 0.0, aln1, aln2, aln_off, aln_cap, score, status);
}

int crt_dtw_align_batch(crt_ctx *c, const double *S, const int64_t *shape_off, const int32_t *n, const int32_t *m,
                        int32_t n_problems, double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2,
                        int64_t *aln_off, int64_t aln_cap, double *score)
{
    return dp_batch(c, true, S, shape_off, n, m, n_problems, gap_open, gap_extend, aln1, aln2, aln_off, aln_cap, score, nullptr);
}

static int rmsd_cov_tm_impl(crt_ctx *c, const int64_t *aln, int64_t A, double *rmsd, double *cov, double *tm, int32_t *n_bad, int superpose);

int crt_rmsd_cov_tm(crt_ctx *c, const int64_t *aln, int64_t A, double *rmsd, double *cov, double *tm, int32_t *n_bad)
{
    return rmsd_cov_tm_impl(c, aln, A, rmsd, cov, tm, n_bad, 1);
}

int crt_rmsd_cov_tm_superposed(crt_ctx *c, const int64_t *aln, int64_t A, double *rmsd, double *cov, double *tm, int32_t *n_bad)
{
    return rmsd_cov_tm_impl(c, aln, A, rmsd, cov, tm, n_bad, 0);
}

static int rmsd_cov_tm_impl(crt_ctx *c, const int64_t *aln, int64_t A, double *rmsd, double *cov, double *tm, int32_t *n_bad, int superpose)
{
    if (!c || !aln || !rmsd || !cov || !tm) return fail(CRT_E_ARG, "null argument");
    if (c->N <= 0) return fail(CRT_E_STATE, "crt_set_chains has not been called");
    if (A <= 0) return fail(CRT_E_ARG, "alignment length must be > 0");
    CU(cudaSetDevice(c->device));
    const int N = c->N;
    for (int p = 0; p < N; ++p) {
        const long long L = c->offsets[p + 1] - c->offsets[p];
        for (int64_t k = 0; k < A; ++k) {
            const long long v = aln[(size_t)p * A + k];
            if (v < -1 || v >= L) return fail(CRT_E_ARG, "aln[%d][%lld] = %lld out of range for a chain of %lld residues", p, (long long)k, v, L);
        }
    }
    // centroids are produced by the prep kernels; make sure they exist (parameters are irrelevant for them)
    crt_params prm{7.0, 0.03, 0.0, CRT_FP32, 0};
    int rc = ensure_prepared(c, c->prep_gamma_t >= 0 ? nullptr : &prm);
    if (rc) return rc;
    struct { long long *p = nullptr; } dAln;
    struct { double *p = nullptr; } dR, dC, dT;
    struct { int *p = nullptr; } dBad;
    Scratch sc(c);
    auto cleanup = [&]() {};
    const size_t NN = (size_t)N * N;
    const int W = (int)((A + 31) / 32);
    unsigned *dBits = nullptr;
    int *dPresent = nullptr;
    CU(sc.alloc(&dBits, (size_t)N * W));
    CU(sc.alloc(&dPresent, (size_t)N));
    CU(sc.alloc(&dAln.p, (size_t)N * A));
    CU(sc.alloc(&dR.p, NN));
    CU(sc.alloc(&dC.p, NN));
    CU(sc.alloc(&dT.p, NN));
    CU(sc.alloc(&dBad.p, 1));
    cudaStream_t st = c->stream;
    cudaMemcpyAsync(dAln.p, aln, sizeof(long long) * (size_t)N * A, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(dBad.p, 0, sizeof(int), st);
    k_fill_diag<<<(unsigned)((NN + 255) / 256), 256, 0, st>>>(dR.p, dC.p, dT.p, N);
    const long long np = (long long)N * (N - 1) / 2;
    cudaEventRecord(c->ev0, st);
    cudaMemsetAsync(dPresent, 0, sizeof(int) * (size_t)N, st);
    k_aln_bits<<<(unsigned)(((long long)N * W * 32 + 255) / 256), 256, 0, st>>>(dAln.p, N, A, W, nullptr, dBits, dPresent, dBad.p);
    if (np > 0)
        k_rmsd_cov_tm<<<(unsigned)((np + 63) / 64), 64, 0, st>>>(dAln.p, N, A, dBits, W, c->coords.p, c->d_offsets.p, c->centroid.p, dR.p, dC.p, dT.p, dBad.p, superpose);
    int bad = 0;
    cudaEventRecord(c->ev1, st);
    cudaMemcpyAsync(rmsd, dR.p, sizeof(double) * NN, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(cov, dC.p, sizeof(double) * NN, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(tm, dT.p, sizeof(double) * NN, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&bad, dBad.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess) return fail(CRT_E_CUDA, "crt_rmsd_cov_tm: %s", cudaGetErrorString(e));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->elapsed_ms = ms;
    if (n_bad) *n_bad = bad;
    return 0;
}

/* One node of progressive_align (multiple_alignment.py:195-234): score_function + weight Gaussian -> dtw_align ->
 * mean_function + get_mean_weights, all on the device in float64.  Inputs are host arrays of the two (consensus)
 * sequences; outputs: the alignment (int32, -1 = gap, length *aln_len <= n + m) and the intermediate node. */
int crt_progressive_node(crt_ctx *c, const double *tensors1, const double *coords1, const double *weights1, int32_t n,
                         const double *tensors2, const double *coords2, const double *weights2, int32_t m, int32_t d,
                         double mult1, double mult2, double gamma_tensor, double gamma_coords, double gamma_weight,
                         double gap_open, double gap_extend, int32_t *aln1, int32_t *aln2, int32_t *aln_len,
                         double *tensors_mean, double *coords_mean, double *weights_mean, double *score, int32_t *status)
{
    if (!c || !tensors1 || !coords1 || !weights1 || !tensors2 || !coords2 || !weights2 || !aln1 || !aln2 || !aln_len ||
        !tensors_mean || !coords_mean || !weights_mean)
        return fail(CRT_E_ARG, "null argument");
    if (n <= 0 || m <= 0) return fail(CRT_E_ARG, "empty sequence (%d x %d)", n, m);
    if (gamma_coords < 0.0) {            // flexible=True: the level path with one node (same kernels, same outputs)
        std::vector<double> pt((size_t)(n + m) * d), pc((size_t)(n + m) * 3), pw((size_t)n + m);
        std::memcpy(pt.data(), tensors1, sizeof(double) * (size_t)n * d);
        std::memcpy(pt.data() + (size_t)n * d, tensors2, sizeof(double) * (size_t)m * d);
        std::memcpy(pc.data(), coords1, sizeof(double) * (size_t)n * 3);
        std::memcpy(pc.data() + (size_t)n * 3, coords2, sizeof(double) * (size_t)m * 3);
        std::memcpy(pw.data(), weights1, sizeof(double) * (size_t)n);
        std::memcpy(pw.data() + n, weights2, sizeof(double) * (size_t)m);
        const int64_t off[3] = {0, n, (int64_t)n + m};
        const double mult[2] = {mult1, mult2};
        return crt_progressive_level(c, 1, d, pt.data(), pc.data(), pw.data(), off, mult, gamma_tensor, gamma_coords, gamma_weight, gap_open,
                                     gap_extend, aln1, aln2, aln_len, tensors_mean, coords_mean, weights_mean, score, status);
    }
    CU(cudaSetDevice(c->device));
    int rc;
    if (!c->node_ctx && (rc = crt_create(c->device, &c->node_ctx))) return rc;
    crt_ctx *nc = c->node_ctx;
    nc->stage1_only = true;
    // ---- stage 1 of score_function on the two sequences: the fp64 pair kernels, pair (0, 1)
    std::vector<double> pc((size_t)(n + m) * 3), pt((size_t)(n + m) * d);
    std::memcpy(pc.data(), coords1, sizeof(double) * (size_t)n * 3);
    std::memcpy(pc.data() + (size_t)n * 3, coords2, sizeof(double) * (size_t)m * 3);
    std::memcpy(pt.data(), tensors1, sizeof(double) * (size_t)n * d);
    std::memcpy(pt.data() + (size_t)n * d, tensors2, sizeof(double) * (size_t)m * d);
    const int64_t off[3] = {0, n, (int64_t)n + m};
    if ((rc = crt_set_chains(nc, pc.data(), pt.data(), off, 2, d))) return rc;
    crt_params prm{};
    prm.gamma_tensor = gamma_tensor; prm.gamma_coords = gamma_coords; prm.sw_gap = 0.0; prm.precision = CRT_FP64;
    const int32_t pi = 0, pj = 1;
    double sc1 = 0, rm = 0, tm = 0;
    int32_t ncm = 0, st1 = 0;
    if ((rc = crt_pairwise_list(nc, &prm, &pi, &pj, 1, &sc1, &rm, &tm, &ncm, &st1, nullptr, nullptr, nullptr, 0))) return rc;
    // ---- score matrix, affine DTW, intermediate node
    const size_t cells = (size_t)n * m, alen = (size_t)n + m + 1;
    if ((rc = c->nd_w.ensure((size_t)n + m))) return rc;
    if ((rc = c->nd_S.ensure(cells))) return rc;
    if ((rc = c->nd_B.ensure((size_t)n * dtw_pitch(m)))) return rc;
    if ((rc = c->nd_bnd.ensure((size_t)n * 2 + 2))) return rc;
    if ((rc = c->nd_f.ensure(3))) return rc;
    if ((rc = c->nd_score.ensure(1))) return rc;
    if ((rc = c->nd_xf2.ensure(XF))) return rc;
    if ((rc = c->nd_a1.ensure(alen))) return rc;
    if ((rc = c->nd_a2.ensure(alen))) return rc;
    if ((rc = c->nd_len.ensure(1))) return rc;
    if ((rc = c->nd_t.ensure(alen * (size_t)d))) return rc;
    if ((rc = c->nd_c.ensure(alen * 3))) return rc;
    if ((rc = c->nd_wm.ensure(alen))) return rc;
    if ((rc = c->d_units.ensure(1))) return rc;                      // one DpProblem record fits in a Unit slot
    static_assert(sizeof(DpProblem) <= sizeof(Unit), "DpProblem must fit in the unit buffer");
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->nd_w.p, weights1, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->nd_w.p + n, weights2, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, st));
    DpProblem pr{};
    pr.s_off = 0; pr.b_off = 0; pr.bnd_off = 0; pr.aln_off = 0; pr.n = n; pr.m = m;
    DpProblem *d_pr = reinterpret_cast<DpProblem *>(c->d_units.p);
    CU(cudaMemcpyAsync(d_pr, &pr, sizeof(pr), cudaMemcpyHostToDevice, st));
    const double *dc1 = nc->coords.p, *dc2 = nc->coords.p + (size_t)n * 3;          // the node context holds the packed chains
    const double *dt1 = nc->tensors.p, *dt2 = nc->tensors.p + (size_t)n * d;
    k_node_score<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(dc1, n, dc2, m, nc->xform.p, c->nd_w.p, c->nd_w.p + n, mult1, mult2,
                                                                 -gamma_coords, -gamma_weight, c->nd_S.p);
    CU(launch_dtw_fill(d_pr, 1, m, c->nd_S.p, c->nd_B.p, c->nd_bnd.p, c->nd_f.p, gap_open, gap_extend, st));
    k_dtw_trace_w<<<1, 32, 0, st>>>(d_pr, 1, c->nd_B.p, c->nd_f.p, c->nd_a1.p, c->nd_a2.p, c->nd_len.p, c->nd_score.p);
    k_node_kabsch<<<1, 32, 0, st>>>(dc1, dc2, c->nd_a1.p, c->nd_a2.p, c->nd_len.p, c->nd_xf2.p);
    k_node_mean<<<(unsigned)((alen + 127) / 128), 128, 0, st>>>(dt1, dc1, c->nd_w.p, dt2, dc2, c->nd_w.p + n, d, c->nd_a1.p, c->nd_a2.p,
                                                                c->nd_len.p, c->nd_xf2.p, c->nd_t.p, c->nd_c.p, c->nd_wm.p);
    CU(cudaGetLastError());
    int len = 0;
    double sc = 0;
    CU(cudaMemcpyAsync(&len, c->nd_len.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&sc, c->nd_score.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (len < 0 || (size_t)len > alen) return fail(CRT_E_STATE, "alignment length %d out of range", len);
    CU(cudaMemcpyAsync(aln1, c->nd_a1.p, sizeof(int) * (size_t)len, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(aln2, c->nd_a2.p, sizeof(int) * (size_t)len, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tensors_mean, c->nd_t.p, sizeof(double) * (size_t)len * d, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(coords_mean, c->nd_c.p, sizeof(double) * (size_t)len * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(weights_mean, c->nd_wm.p, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *aln_len = len;
    if (score) *score = sc;
    if (status) *status = st1;
    return 0;
}

/* Protein.score_function (multiple_alignment.py:321-349): the n x m float64 score matrix of one pair, as the reference returns it. */
int crt_score_matrix(crt_ctx *c, const double *tensors1, const double *coords1, int32_t n, const double *tensors2, const double *coords2,
                     int32_t m, int32_t d, double gamma_tensor, double gamma_coords, int32_t flexible, double *score_matrix, int32_t *status)
{
    if (!c || !tensors1 || !tensors2 || !score_matrix) return fail(CRT_E_ARG, "null argument");
    if (!flexible && (!coords1 || !coords2)) return fail(CRT_E_ARG, "coordinates are needed unless flexible is set");
    if (n <= 0 || m <= 0 || d <= 0) return fail(CRT_E_ARG, "empty sequence (%d x %d, d = %d)", n, m, d);
    if (!(gamma_tensor >= 0) || (!flexible && !(gamma_coords >= 0))) return fail(CRT_E_ARG, "gamma must be >= 0");
    CU(cudaSetDevice(c->device));
    int rc;
    if (!c->node_ctx && (rc = crt_create(c->device, &c->node_ctx))) return rc;
    crt_ctx *nc = c->node_ctx;
    nc->stage1_only = true;
    std::vector<double> pc((size_t)(n + m) * 3, 0.0), pt((size_t)(n + m) * d);
    if (coords1) std::memcpy(pc.data(), coords1, sizeof(double) * (size_t)n * 3);
    if (coords2) std::memcpy(pc.data() + (size_t)n * 3, coords2, sizeof(double) * (size_t)m * 3);
    std::memcpy(pt.data(), tensors1, sizeof(double) * (size_t)n * d);
    std::memcpy(pt.data() + (size_t)n * d, tensors2, sizeof(double) * (size_t)m * d);
    const int64_t off[3] = {0, n, (int64_t)n + m};
    if ((rc = crt_set_chains(nc, pc.data(), pt.data(), off, 2, d))) return rc;
    int32_t st1 = 0;
    if (!flexible) {                       // stage 1 of score_function: the fp64 pair kernels leave the superposition in nc->xform
        crt_params prm{};
        prm.gamma_tensor = gamma_tensor; prm.gamma_coords = gamma_coords; prm.sw_gap = 0.0; prm.precision = CRT_FP64;
        const int32_t pi = 0, pj = 1;
        if ((rc = crt_pairwise_list(nc, &prm, &pi, &pj, 1, nullptr, nullptr, nullptr, nullptr, &st1, nullptr, nullptr, nullptr, 0))) return rc;
    }
    const size_t cells = (size_t)n * m;
    if ((rc = c->nd_S.ensure(cells))) return rc;
    if ((rc = c->nd_w.ensure((size_t)n + m))) return rc;
    if ((rc = c->d_units.ensure(1))) return rc;
    cudaStream_t st = c->stream;
    DpProblem pr{};
    pr.n = n; pr.m = m;
    DpProblem *d_pr = reinterpret_cast<DpProblem *>(c->d_units.p);
    CU(cudaMemcpyAsync(d_pr, &pr, sizeof(pr), cudaMemcpyHostToDevice, st));
    // the weight term is switched off (neg_gamma_w > 0): the weights / multipliers are not read
    if (flexible)
        k_level_score_flex<<<dim3((unsigned)((cells + 255) / 256), 1), 256, 0, st>>>(d_pr, nc->tensors.p, d, c->nd_w.p, c->nd_w.p, -gamma_tensor, 1.0,
                                                                                   c->nd_S.p);
    else
        k_node_score<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(nc->coords.p, n, nc->coords.p + (size_t)n * 3, m, nc->xform.p, c->nd_w.p,
                                                                     c->nd_w.p + n, 0.0, 0.0, -gamma_coords, 1.0, c->nd_S.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(score_matrix, c->nd_S.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (status) *status = st1;
    return 0;
}

namespace {
// alignment columns of mean_function / get_mean_weights: int64 with -1 = gap -> int32; a column must have a residue on one side
int check_columns(const int64_t *aln1, const int64_t *aln2, int64_t len, int n, int m, std::vector<int> &a1, std::vector<int> &a2, int *n_common)
{
    a1.resize((size_t)len); a2.resize((size_t)len);
    int common = 0;
    for (int64_t q = 0; q < len; ++q) {
        const int64_t x = aln1[q], y = aln2[q];
        if (x < -1 || x >= n || y < -1 || y >= m) return fail(CRT_E_ARG, "alignment column %lld: index (%lld, %lld) out of range", (long long)q, (long long)x, (long long)y);
        if (x < 0 && y < 0) return fail(CRT_E_ARG, "alignment column %lld has a gap on both sides", (long long)q);
        a1[(size_t)q] = (int)x; a2[(size_t)q] = (int)y;
        common += x >= 0 && y >= 0;
    }
    if (n_common) *n_common = common;
    return 0;
}
}  // namespace

/* Protein.mean_function (multiple_alignment.py:351-383) for a given alignment. */
int crt_mean_function(crt_ctx *c, const double *tensors1, const double *coords1, int32_t n, const double *tensors2, const double *coords2,
                      int32_t m, int32_t d, const int64_t *aln1, const int64_t *aln2, int64_t len, int32_t flexible, double *tensors_mean,
                      double *coords_mean, int32_t *status)
{
    if (!c || !tensors1 || !tensors2 || !tensors_mean || (len > 0 && (!aln1 || !aln2))) return fail(CRT_E_ARG, "null argument");
    if (!flexible && (!coords1 || !coords2 || !coords_mean)) return fail(CRT_E_ARG, "coordinates are needed unless flexible is set");
    if (n <= 0 || m <= 0 || d <= 0) return fail(CRT_E_ARG, "empty sequence (%d x %d, d = %d)", n, m, d);
    if (len < 0 || len > (int64_t)n + m) return fail(CRT_E_ARG, "alignment length %lld out of range (<= n + m)", (long long)len);
    std::vector<int> a1, a2;
    int common = 0, rc;
    if ((rc = check_columns(aln1, aln2, len, n, m, a1, a2, &common))) return rc;
    if (status) *status = (!flexible && common <= 3) ? CRT_ST_FEW_COMMON : 0;
    if (len == 0) return 0;
    CU(cudaSetDevice(c->device));
    const size_t rows = (size_t)n + m, alen = (size_t)len;
    // the two sequences go to the score-matrix workspace (free here): tensors [n + m, d], then coordinates [n + m, 3]
    if ((rc = c->nd_S.ensure(rows * ((size_t)d + 3))) || (rc = c->nd_w.ensure(rows)) || (rc = c->nd_xf2.ensure(XF)) ||
        (rc = c->nd_a1.ensure(alen + 1)) || (rc = c->nd_a2.ensure(alen + 1)) || (rc = c->nd_len.ensure(1)) || (rc = c->nd_t.ensure(alen * (size_t)d)) ||
        (rc = c->nd_c.ensure(alen * 3)) || (rc = c->nd_wm.ensure(alen)))
        return rc;
    cudaStream_t st = c->stream;
    double *dt1 = c->nd_S.p, *dt2 = dt1 + (size_t)n * d, *dc1 = c->nd_S.p + rows * (size_t)d, *dc2 = dc1 + (size_t)n * 3;
    const int ilen = (int)len;
    CU(cudaMemcpyAsync(dt1, tensors1, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dt2, tensors2, sizeof(double) * (size_t)m * d, cudaMemcpyHostToDevice, st));
    if (flexible) {
        CU(cudaMemsetAsync(dc1, 0, sizeof(double) * rows * 3, st));
    } else {
        CU(cudaMemcpyAsync(dc1, coords1, sizeof(double) * (size_t)n * 3, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(dc2, coords2, sizeof(double) * (size_t)m * 3, cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemsetAsync(c->nd_w.p, 0, sizeof(double) * rows, st));
    CU(cudaMemcpyAsync(c->nd_a1.p, a1.data(), sizeof(int) * alen, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->nd_a2.p, a2.data(), sizeof(int) * alen, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->nd_len.p, &ilen, sizeof(int), cudaMemcpyHostToDevice, st));
    if (flexible)
        CU(cudaMemsetAsync(c->nd_xf2.p, 0, sizeof(double) * XF, st));             // flag 0: no superposition
    else
        k_node_kabsch<<<1, 32, 0, st>>>(dc1, dc2, c->nd_a1.p, c->nd_a2.p, c->nd_len.p, c->nd_xf2.p);
    k_node_mean<<<(unsigned)((alen + 127) / 128), 128, 0, st>>>(dt1, dc1, c->nd_w.p, dt2, dc2, c->nd_w.p + n, d, c->nd_a1.p, c->nd_a2.p, c->nd_len.p,
                                                                c->nd_xf2.p, c->nd_t.p, c->nd_c.p, c->nd_wm.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(tensors_mean, c->nd_t.p, sizeof(double) * alen * d, cudaMemcpyDeviceToHost, st));
    if (!flexible) CU(cudaMemcpyAsync(coords_mean, c->nd_c.p, sizeof(double) * alen * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

/* get_mean_weights (multiple_alignment.py:73-82). */
int crt_mean_weights(crt_ctx *c, const double *weights1, int32_t n, const double *weights2, int32_t m, const int64_t *aln1, const int64_t *aln2,
                     int64_t len, double *weights_mean)
{
    if (!c || !weights1 || !weights2 || (len > 0 && (!aln1 || !aln2 || !weights_mean))) return fail(CRT_E_ARG, "null argument");
    if (n <= 0 || m <= 0) return fail(CRT_E_ARG, "empty sequence (%d x %d)", n, m);
    if (len < 0) return fail(CRT_E_ARG, "negative alignment length");
    std::vector<int> a1, a2;
    int rc;
    if ((rc = check_columns(aln1, aln2, len, n, m, a1, a2, nullptr))) return rc;
    if (len == 0) return 0;
    CU(cudaSetDevice(c->device));
    const size_t alen = (size_t)len;
    if ((rc = c->nd_w.ensure((size_t)n + m)) || (rc = c->nd_a1.ensure(alen + 1)) || (rc = c->nd_a2.ensure(alen + 1)) || (rc = c->nd_wm.ensure(alen)))
        return rc;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->nd_w.p, weights1, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->nd_w.p + n, weights2, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->nd_a1.p, a1.data(), sizeof(int) * alen, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->nd_a2.p, a2.data(), sizeof(int) * alen, cudaMemcpyHostToDevice, st));
    k_mean_weights<<<(unsigned)((alen + 255) / 256), 256, 0, st>>>(c->nd_w.p, c->nd_w.p + n, c->nd_a1.p, c->nd_a2.p, (long long)len, c->nd_wm.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(weights_mean, c->nd_wm.p, sizeof(double) * alen, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

/* neighbor_joining.py:17-99 on the device: guide tree (node_1, node_2) rows + branch lengths, bit-identical to the
 * reference for any float64 input (symmetric or not).  tree: [2N-3][2] uint64, branch_lengths: [2N-3] float64. */
int crt_neighbor_joining(crt_ctx *c, const double *distance_matrix, int32_t N, uint64_t *tree, double *branch_lengths, int64_t *n_rows)
{
    if (!c || !distance_matrix || !tree || !branch_lengths) return fail(CRT_E_ARG, "null argument");
    if (N < 3) return fail(CRT_E_ARG, "neighbor joining needs at least 3 nodes (the reference indexes out of range below that)");
    CU(cudaSetDevice(c->device));
    const size_t NN = (size_t)N * N, rows_max = (size_t)2 * N - 3;
    const int max_part = 1024;
    double *A = nullptr, *B = nullptr, *S0 = nullptr, *S1 = nullptr, *pq = nullptr, *d_bl = nullptr;
    long long *t0 = nullptr, *t1 = nullptr, *plin = nullptr;
    unsigned long long *d_tree = nullptr;
    NjSel *sel = nullptr;
    unsigned *ticket = nullptr;
    Scratch sc(c);                       // arena of the context: no cudaMalloc / cudaFree per call once it has grown
    auto cleanup = [&]() {};
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess && r != cudaSuccess) e = r; return e == cudaSuccess; };
    // ---- in-place path (crt_nj.cuh, second half): reversed storage, joins append, the host compacts every n / 8 joins
    //      from CARETTA_B200_NJ_INPLACE_MIN nodes on (default 512: below that the segments' extra launches cost more than the
    //      moved matrix; 4 = always, 0 = never)
    const int inplace_min = getenv("CARETTA_B200_NJ_INPLACE_MIN") ? atoi(getenv("CARETTA_B200_NJ_INPLACE_MIN")) : 512;
    if (N > 3 && inplace_min > 0 && N >= inplace_min) {
        const int seg_min = 64;                          // joins per segment: n / 8, at least this many
        const int ld = ((N + std::max(N / 8, seg_min) + 2 + 15) / 16) * 16;
        double *in = nullptr, *D0 = nullptr, *D1 = nullptr, *Sa = nullptr, *Sb = nullptr;
        long long *ta = nullptr, *tb = nullptr, *pkey = nullptr;
        int *al_a = nullptr, *al_b = nullptr, *map = nullptr, *cnt = nullptr;
        cudaStream_t st = c->stream;
        const size_t LL = (size_t)ld * ld;
        if (ok(sc.alloc(&D0, LL)) && ok(sc.alloc(&D1, LL)) && ok(sc.alloc(&Sa, (size_t)ld)) && ok(sc.alloc(&Sb, (size_t)ld)) &&
            ok(sc.alloc(&ta, (size_t)ld)) && ok(sc.alloc(&tb, (size_t)ld)) && ok(sc.alloc(&al_a, (size_t)ld)) && ok(sc.alloc(&al_b, (size_t)ld)) &&
            ok(sc.alloc(&map, (size_t)ld)) && ok(sc.alloc(&cnt, 1)) && ok(sc.alloc(&pq, (size_t)max_part)) && ok(sc.alloc(&pkey, (size_t)max_part)) &&
            ok(sc.alloc(&d_tree, rows_max * 2)) && ok(sc.alloc(&d_bl, rows_max)) && ok(sc.alloc(&sel, 1))) {
            in = D1;                                     // the upload lands in the second buffer (N * N <= ld * ld)
            ok(cudaMemcpyAsync(in, distance_matrix, NN * 8, cudaMemcpyHostToDevice, st));
            ok(cudaMemsetAsync(sel, 0, sizeof(NjSel), st));
            ok(cudaFuncSetAttribute(k_nj2_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NJ2_SMEM));
            int per_sm = 0;
            ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nj2_persistent, NJ_REBUILD_THREADS, NJ2_SMEM));
            const int nb_max = std::min(max_part, std::max(1, per_sm) * c->sm_count);
            CU(cudaEventRecord(c->ev0, st));
            k_nj2_load<<<dim3((unsigned)((N + 255) / 256), (unsigned)N), 256, 0, st>>>(in, N, D0, ld, ta, al_a, nullptr);
            k_nj2_rowsums<<<(unsigned)(((size_t)N * 32 + 255) / 256), 256, 0, st>>>(D0, N, ld, Sa);
            int n = N;
            while (n > 3 && e == cudaSuccess) {
                // one segment: `iters` joins on a compact matrix of n nodes (the extent grows to n + iters <= ld), then compaction
                const int iters = std::min(n - 3, std::max(seg_min, n / 8));
                Nj2Args a{D0, ld, Sa, ta, al_a, n, n, iters, N, pq, pkey, sel, d_tree, d_bl};
                // enough CTAs for every row group of the join to run at once (the add chain of a row is n dependent adds), no more
                const int groups = (n + iters + NJ_ROWS) / NJ_ROWS;
                // every row group of the join at once; beyond that the scan wants loads in flight (large n) and the two grid
                // barriers of a join want few CTAs (small n): CARETTA_B200_NJ_NB overrides (experiments)
                int nb = std::min(nb_max, std::max(groups, n >= 2048 ? nb_max : (n >= 1024 ? 2 * c->sm_count : c->sm_count)));
                if (getenv("CARETTA_B200_NJ_NB")) nb = std::max(1, std::min(nb_max, atoi(getenv("CARETTA_B200_NJ_NB"))));
                void *args[] = {&a};
                ok(cudaLaunchCooperativeKernel((const void *)k_nj2_persistent, dim3((unsigned)nb), dim3(NJ_REBUILD_THREADS), args, NJ2_SMEM, st));
                const int W = n + iters;
                n -= iters;
                k_nj2_map<<<1, 1024, 0, st>>>(al_a, W, map, cnt);
                k_nj2_compact<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256, 0, st>>>(D0, ld, map, n, D1, ld, Sa, Sb, ta, tb, al_b);
                std::swap(D0, D1); std::swap(Sa, Sb); std::swap(ta, tb); std::swap(al_a, al_b);
                ok(cudaGetLastError());
            }
            k_nj2_last3<<<1, 32, 0, st>>>(D0, ld, Sa, N, ta, sel, d_tree, d_bl);
            ok(cudaGetLastError());
            CU(cudaEventRecord(c->ev1, st));
            ok(cudaMemcpyAsync(tree, d_tree, rows_max * 16, cudaMemcpyDeviceToHost, st));
            ok(cudaMemcpyAsync(branch_lengths, d_bl, rows_max * 8, cudaMemcpyDeviceToHost, st));
            ok(cudaStreamSynchronize(st));
            float ms = 0;
            if (e == cudaSuccess && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->elapsed_ms = ms;
        }
        if (e != cudaSuccess) return fail(CRT_E_CUDA, "crt_neighbor_joining: %s", cudaGetErrorString(e));
        if (n_rows) *n_rows = (int64_t)rows_max;
        return 0;
    }
    std::vector<long long> ident((size_t)N);
    for (int q = 0; q < N; ++q) ident[(size_t)q] = q;
    cudaStream_t st = c->stream;
    if (ok(sc.alloc(&A, NN)) && ok(sc.alloc(&B, NN)) && ok(sc.alloc(&S0, (size_t)N)) && ok(sc.alloc(&S1, (size_t)N)) &&
        ok(sc.alloc(&pq, (size_t)max_part)) && ok(sc.alloc(&plin, (size_t)max_part)) && ok(sc.alloc(&t0, (size_t)N)) &&
        ok(sc.alloc(&t1, (size_t)N)) && ok(sc.alloc(&d_tree, rows_max * 2)) && ok(sc.alloc(&d_bl, rows_max)) &&
        ok(sc.alloc(&sel, 1)) && ok(sc.alloc(&ticket, 1))) {
        ok(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
        ok(cudaMemcpyAsync(A, distance_matrix, NN * 8, cudaMemcpyHostToDevice, st));
        ok(cudaMemcpyAsync(t0, ident.data(), (size_t)N * 8, cudaMemcpyHostToDevice, st));
        ok(cudaMemsetAsync(sel, 0, sizeof(NjSel), st));
        ok(cudaFuncSetAttribute(k_nj_rebuild, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NJ_REBUILD_SMEM));
        CU(cudaEventRecord(c->ev0, st));
        k_nj_rowsums<<<(unsigned)(((size_t)N * 32 + 255) / 256), 256, 0, st>>>(A, N, S0);
        int n = N;
        // the small end of the run in one cooperative launch (CARETTA_B200_NJ_PERSIST=0: two launches per iteration throughout)
        const int persist_max = getenv("CARETTA_B200_NJ_PERSIST") ? atoi(getenv("CARETTA_B200_NJ_PERSIST")) : 2048;
        while (n > 3 && e == cudaSuccess) {
            if (n <= persist_max) {
                ok(cudaFuncSetAttribute(k_nj_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NJ_REBUILD_SMEM));
                int nb = std::min(c->sm_count, max_part);
                int nn = n, NN_ = N;
                void *args[] = {&A, &B, &S0, &S1, &t0, &t1, &nn, &NN_, &pq, &plin, &sel, &d_tree, &d_bl};
                ok(cudaLaunchCooperativeKernel((const void *)k_nj_persistent, dim3((unsigned)nb), dim3(NJ_REBUILD_THREADS), args, NJ_REBUILD_SMEM, st));
                if ((n - 3) & 1) { std::swap(A, B); std::swap(S0, S1); std::swap(t0, t1); }       // n - 3 iterations ran on the device
                n = 3;
                break;
            }
            const int n_part = std::min(max_part, n);                    // one block per row, rows beyond max_part wrap around
            k_nj_argmin<<<n_part, NJ_ARGMIN_THREADS, 0, st>>>(A, S0, n, pq, plin, N, t0, sel, d_tree, d_bl, ticket);
            k_nj_rebuild<<<(unsigned)((n - 1 + NJ_ROWS - 1) / NJ_ROWS), NJ_REBUILD_THREADS, NJ_REBUILD_SMEM, st>>>(A, n, sel, N, t0, t1, B, S1);
            std::swap(A, B); std::swap(S0, S1); std::swap(t0, t1);
            --n;
            if ((n & 255) == 0) ok(cudaGetLastError());
        }
        k_nj_last3<<<1, 32, 0, st>>>(A, S0, N, t0, sel, d_tree, d_bl);
        ok(cudaGetLastError());
        CU(cudaEventRecord(c->ev1, st));
        ok(cudaMemcpyAsync(tree, d_tree, rows_max * 16, cudaMemcpyDeviceToHost, st));
        ok(cudaMemcpyAsync(branch_lengths, d_bl, rows_max * 8, cudaMemcpyDeviceToHost, st));
        ok(cudaStreamSynchronize(st));
        float ms = 0;
        if (e == cudaSuccess && cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->elapsed_ms = ms;
    }
    cleanup();
    if (e != cudaSuccess) return fail(CRT_E_CUDA, "crt_neighbor_joining: %s", cudaGetErrorString(e));
    if (n_rows) *n_rows = (int64_t)rows_max;
    return 0;
}

int crt_fp32_peak(crt_ctx *c, double *ffma_per_s, double *elapsed_ms)
{
    if (!c || !ffma_per_s) return fail(CRT_E_ARG, "null argument");
    CU(cudaSetDevice(c->device));
    int rc;
    if ((rc = c->f32tmp.ensure(64))) return rc;
    const int iters = 4096, block = 256, grid = c->sm_count * 8;
    k_ffma_peak<<<grid, block, 0, c->stream>>>(c->f32tmp.p, 64);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaEventRecord(c->ev0, c->stream));
        k_ffma_peak<<<grid, block, 0, c->stream>>>(c->f32tmp.p, iters);
        CU(cudaEventRecord(c->ev1, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        best = std::min(best, ms);
    }
    const double ops = (double)grid * block * (double)iters * 16.0 * 8.0;
    *ffma_per_s = ops / (best * 1e-3);
    if (elapsed_ms) *elapsed_ms = best;
    return 0;
}

}  // extern "C"

#include "crt_consumers_api.inl"
#include "crt_level_api.inl"

#include "crt_multi.inl"
