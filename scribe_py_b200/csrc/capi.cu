// capi.cu -- host side of libscribe_b200.so: handle, workspaces, job batching and the extern "C" boundary
// declared in include/scribe_b200.h.  No torch types, no CPU compute path: without a CUDA device every
// entry point fails.
#include "../../include/scribe_b200.h"
#include "kernels.cuh"
#include "sweep_window.cuh"
#include "prep.cuh"
#include <cub/device/device_segmented_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace scribe;

static thread_local std::string g_err;
extern "C" const char* scribe_last_error(void) { return g_err.c_str(); }
extern "C" int scribe_abi_version(void) { return SCRIBE_ABI_VERSION; }

namespace {

struct Fail { int code; std::string msg; };
[[noreturn]] void fail(int code, const std::string& m) { throw Fail{code, m}; }

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    fail(SCRIBE_E_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        CU(cudaMalloc(&p, want)); cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct scribe_handle {
    int device = 0;
    int sm_count = 0, clock_khz = 0;
    size_t mem_bytes = 0;
    char name[128] = {0};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    // resident inputs
    DevBuf E, run_off_d; int64_t n_genes = 0, n_cells = 0; std::vector<int64_t> run_off;
    DevBuf S, Y; int64_t cg = 0, cc = 0; bool permS_valid = false;
    // workspaces
    DevBuf psi; int psi_n = 0;
    DevBuf jobs, partial, cols, specs, u, r, eps, cnt, wsum, out, flag, idx32a, idx32b, neff, dims, pjobs, single;
    DevBuf keys_in, keys_out, vals_in, perm, sort_tmp, seg_off, keyspecs, visited, permS, strips, permE, sortedE, rawjobs;
    DevBuf post[8];                              // CLR / edge ranking / roc scratch
    DevBuf raw0, raw1, stat, nanflag;            // raw layers / row statistics of the on-device normalisation
    DevBuf wcum;                                 // umi: prefix sums of the weights in sweep order
    DevBuf sorted, rankf, crank, sortedS, rankS, permY, sortedY, rankY, rankE;   // rank-space count pass: ascending values + float ranks
    bool sortE_valid = false;             // permE / sortedE describe the uploaded expression matrix
    scribe_stats st{};
    size_t ws_budget = (size_t)3 << 30;   // per-chunk workspace target (bytes)
    int force_q = 0;                      // SCRIBE_B200_Q: override queries per thread (tuning)
    int no_fast_rdi = 0;                  // SCRIBE_B200_NO_FAST_RDI=1: always take the generic job loop of scribe_lagged_jobs
    int sweep_v1 = 0;                     // SCRIBE_B200_SWEEP_V1=1: first sweep form (FP64 scan in both passes) for unweighted jobs too
    int no_strip2 = 0;                    // SCRIBE_B200_NO_STRIP2=1: use the first strip form also for dz = 1 (tuning)
};

namespace {

struct Guard {   // selects the handle's device for the duration of a call
    int prev = -1;
    explicit Guard(scribe_handle* h) { cudaGetDevice(&prev); cudaSetDevice(h->device); }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

void ensure_psi(scribe_handle* h, int n_max) {
    if (n_max + 2 <= h->psi_n) return;
    const int m = std::max(n_max + 2, 4096);
    std::vector<double> t(m);
    for (int i = 0; i < m; i++) t[i] = digamma_f64((double)i);      // psi(0) = -inf like scipy
    h->psi.ensure(sizeof(double) * m);
    CU(cudaMemcpyAsync(h->psi.p, t.data(), sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->st.h2d_bytes += sizeof(double) * m;
    h->psi_n = m;
}

void check_opts(const scribe_opts* o) {
    if (!o) fail(SCRIBE_E_ARG, "opts is NULL");
    if (o->estimator < 0 || o->estimator > 3) fail(SCRIBE_E_ARG, "unknown estimator");
    if (o->k < 1 || o->k > SCRIBE_MAX_K) fail(SCRIBE_E_ARG, "k must be in [1, 63]");
    if (o->estimator >= SCRIBE_EST_UMI) {
        if (o->density != SCRIBE_DENSITY_KDE && o->density != SCRIBE_DENSITY_KNN)
            fail(SCRIBE_E_ARG, "The density estimation method is not recognized");
        if (o->density == SCRIBE_DENSITY_KNN && (o->k_density < 1 || o->k_density > SCRIBE_MAX_K))
            fail(SCRIBE_E_ARG, "k_density must be in [1, 63]");
        if (o->density == SCRIBE_DENSITY_KDE && !(o->bw > 0.0)) fail(SCRIBE_E_ARG, "bw must be positive");
    }
}

// ---- fused-kernel dispatch -------------------------------------------------------------------------
using FusedFn = void (*)(const JobDesc*, int, int, int, int, const double*, double*, PointOut);

template <int DX, int DY, int DZ, int KCAP, bool W>
FusedFn pick_q(int q) {
    switch (q) {
        case 1: return ksg_fused_kernel<DX, DY, DZ, KCAP, 1, W>;
        case 2: return ksg_fused_kernel<DX, DY, DZ, KCAP, 2, W>;
    }
    return nullptr;
}

template <int DX, int DY, int DZ, bool W>
FusedFn pick_kq(int kcap, int q) {
    if (kcap == 6)  return pick_q<DX, DY, DZ, 6, W>(q);
    if (kcap == 16) return pick_q<DX, DY, DZ, 16, W>(q);
    return nullptr;
}

FusedFn pick_fused(int dx, int dy, int dz, int kcap, int q, bool weighted) {
    if (dx != 1 || dy != 1) return nullptr;
    if (!weighted) {
        switch (dz) {
            case 0: return pick_kq<1, 1, 0, false>(kcap, q);
            case 1: return pick_kq<1, 1, 1, false>(kcap, q);
            case 2: return pick_kq<1, 1, 2, false>(kcap, q);
            case 3: return pick_kq<1, 1, 3, false>(kcap, q);
        }
    } else {
        switch (dz) {
            case 1: return pick_kq<1, 1, 1, true>(kcap, q);
            case 2: return pick_kq<1, 1, 2, true>(kcap, q);
            case 3: return pick_kq<1, 1, 3, true>(kcap, q);
        }
    }
    return nullptr;
}

using SweepFn = void (*)(const JobDesc*, int, int, const double*, double*, PointOut, unsigned long long*);

template <int DZ, bool W>
SweepFn pick_sweep_k(int kcap) {
    if (kcap == 6)  return ksg_sweep_kernel<1, 1, DZ, 6, 2, W>;
    if (kcap == 16) return ksg_sweep_kernel<1, 1, DZ, 16, 2, W>;
    return nullptr;
}
SweepFn pick_sweep(int dx, int dy, int dz, int kcap, bool weighted) {
    if (dx != 1 || dy != 1) return nullptr;
    switch (dz) {
        case 1: return weighted ? pick_sweep_k<1, true>(kcap) : pick_sweep_k<1, false>(kcap);
        case 2: return weighted ? pick_sweep_k<2, true>(kcap) : pick_sweep_k<2, false>(kcap);
        case 3: return weighted ? pick_sweep_k<3, true>(kcap) : pick_sweep_k<3, false>(kcap);
    }
    return nullptr;
}
// windowed sweep (sweep_window.cuh)
template <int DZ, bool W>
SweepFn pick_sweepw_k(int kcap) {
    if (kcap == 6)  return ksg_sweepw_kernel<DZ, 6, 2, W>;
    if (kcap == 16) return ksg_sweepw_kernel<DZ, 16, 2, W>;
    return nullptr;
}
SweepFn pick_sweepw(int dx, int dy, int dz, int kcap, bool weighted) {
    if (dx != 1 || dy != 1) return nullptr;
    switch (dz) {
        case 0: return weighted ? pick_sweepw_k<0, true>(kcap) : pick_sweepw_k<0, false>(kcap);   // mi(x, y) / umi(x, y): sorted by y, marginals by rank
        case 1: return weighted ? pick_sweepw_k<1, true>(kcap) : pick_sweepw_k<1, false>(kcap);
        case 2: return weighted ? pick_sweepw_k<2, true>(kcap) : pick_sweepw_k<2, false>(kcap);
        case 3: return weighted ? pick_sweepw_k<3, true>(kcap) : pick_sweepw_k<3, false>(kcap);
    }
    return nullptr;
}
using StripFn = void (*)(const JobDesc*, int, int, const double*, double*, PointOut, unsigned long long*);
using StripSortFn = void (*)(JobDesc*, int, int, double*, int64_t, int64_t);
template <int DZ>
StripFn pick_strip_k(int kcap, bool weighted) {
    if (kcap == 6)  return weighted ? ksg_strip_kernel<DZ, 6, true>  : ksg_strip_kernel<DZ, 6, false>;
    if (kcap == 16) return weighted ? ksg_strip_kernel<DZ, 16, true> : ksg_strip_kernel<DZ, 16, false>;
    return nullptr;
}
StripFn pick_strip(int dx, int dy, int dz, int kcap, bool weighted) {
    if (dx != 1 || dy != 1) return nullptr;
    switch (dz) {
        case 1: return pick_strip_k<1>(kcap, weighted);
        case 2: return pick_strip_k<2>(kcap, weighted);
        case 3: return pick_strip_k<3>(kcap, weighted);
    }
    return nullptr;
}
StripSortFn pick_strip_sort(int D) {
    switch (D) {
        case 3: return strip_sort_kernel<3>;
        case 4: return strip_sort_kernel<4>;
        case 5: return strip_sort_kernel<5>;
    }
    return nullptr;
}
constexpr int SWEEP_Q = 2;
constexpr int SWEEP_MIN_N = 256;      // below this the dense kernel's single tile is at least as good

using KdeSweepFn = void (*)(const JobDesc*, int, const int*, double, double*, unsigned long long*);
KdeSweepFn pick_kde_sweep(int dp) {
    switch (dp) {
        case 2: return kde_sweep_kernel<2, 2>;
        case 3: return kde_sweep_kernel<3, 2>;
        case 4: return kde_sweep_kernel<4, 2>;
    }
    return nullptr;
}

// Will run_jobs take the sorted-sweep kernels for this configuration?  (callers must then provide JobDesc.perm
// or pre-sorted columns)
bool want_sweep(const scribe_opts* o, int dx, int dy, int dz, int n_max) {
    const int kcap = (o->k + 1 <= 6) ? 6 : ((o->k + 1 <= 16) ? 16 : 0);
    if (o->estimator == SCRIBE_EST_MI) {                   // windowed sweep only (sorted by y; JobDesc.srt[0] for the x marginal)
        if (!kcap || !pick_sweepw(dx, dy, 0, kcap, false)) return false;
        return o->path == SCRIBE_PATH_SWEEP || (o->path == SCRIBE_PATH_AUTO && n_max >= SWEEP_MIN_N);
    }
    if (o->estimator == SCRIBE_EST_UMI) {                  // Euclidean variant of the windowed sweep (scalar x and y)
        if (!kcap || !pick_sweepw(dx, dy, 0, kcap, true)) return false;
        return o->path == SCRIBE_PATH_SWEEP || (o->path == SCRIBE_PATH_AUTO && n_max >= SWEEP_MIN_N);
    }
    if (o->estimator != SCRIBE_EST_CMI && o->estimator != SCRIBE_EST_CUMI) return false;
    if (!kcap || !pick_sweep(dx, dy, dz, kcap, false)) return false;
    if (o->path == SCRIBE_PATH_SWEEP || o->path == SCRIBE_PATH_STRIP) return true;
    return o->path == SCRIBE_PATH_AUTO && n_max >= SWEEP_MIN_N;
}

// argsort of n_seg key segments laid out [n_seg][stride] (segment s has len[s] valid items): perm_out[s][r] = index
// of the r-th smallest key of segment s; sorted_out (optional, default h->keys_out) receives the keys in ascending order and
// rank_out (optional) the inverse permutation as float.  One cub::DeviceSegmentedRadixSort over all segments.
void sort_segments(scribe_handle* h, const double* keys_in, int n_seg, int64_t stride, const std::vector<int>& len,
                   int32_t* perm_out, double* sorted_out = nullptr, float* rank_out = nullptr, const std::vector<int>* only = nullptr)
{
    // `only`: sort just these of the n_seg columns (outputs keep the [n_seg][stride] layout; the other columns are left untouched)
    const int64_t total = (int64_t)n_seg * stride;
    if (total >= (int64_t)1 << 31) fail(SCRIBE_E_ARG, "sort too large for one chunk");
    if (!sorted_out) { h->keys_out.ensure(sizeof(double) * total); sorted_out = h->keys_out.as<double>(); }
    h->vals_in.ensure(sizeof(int32_t) * total);
    const int n_sort = only ? (int)only->size() : n_seg;
    std::vector<int> off(2 * (size_t)n_sort);
    for (int i = 0; i < n_sort; i++) {
        const int s = only ? (*only)[i] : i;
        off[i] = (int)(s * stride); off[n_sort + i] = (int)(s * stride) + len[s];
    }
    if (n_sort == 0) return;
    h->seg_off.ensure(sizeof(int) * off.size());
    CU(cudaMemcpyAsync(h->seg_off.p, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, h->stream));
    h->st.h2d_bytes += sizeof(int) * off.size();
    iota_segments_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 65535), 256, 0, h->stream>>>(h->vals_in.as<int32_t>(), stride, n_seg);
    CU(cudaGetLastError());
    size_t tmp = 0;
    const int* b = h->seg_off.as<int>(); const int* e = b + n_sort;
    CU(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tmp, keys_in, sorted_out, h->vals_in.as<int32_t>(), perm_out,
                                                (int)total, n_sort, b, e, 0, 64, h->stream));
    h->sort_tmp.ensure(tmp + 16);
    CU(cub::DeviceSegmentedRadixSort::SortPairs(h->sort_tmp.p, tmp, keys_in, sorted_out, h->vals_in.as<int32_t>(), perm_out,
                                                (int)total, n_sort, b, e, 0, 64, h->stream));
    h->st.kernel_launches += 2;
    if (rank_out) {
        if (only) fail(SCRIBE_E_STATE, "ranks need every column sorted");
        rank_from_perm_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 65535), 256, 0, h->stream>>>(perm_out, b, e, stride, n_seg, rank_out);
        CU(cudaGetLastError());
        h->st.kernel_launches++;
    }
    CU(cudaStreamSynchronize(h->stream));         // `off` must outlive the copy; also surfaces sort errors here
}

// every column of h->cols [n_cols][stride] (column c has len[c] valid items) in ascending order: h->sorted (values),
// h->perm (argsort), h->rankf (inverse argsort as float).  The sweep order of a job is the perm of its key column; the
// rank-space count pass reads sorted/rankf of its other columns.
constexpr bool kRankPass = SWEEPW_RANK != 0;         // the windowed sweep's count pass works on ranks: every column is sorted, not only the keys

void sort_all_columns(scribe_handle* h, int64_t n_cols, const std::vector<int>& len, int64_t stride, const std::vector<int>& key_cols)
{
    h->sorted.ensure(sizeof(double) * n_cols * stride);
    h->perm.ensure(sizeof(int32_t) * n_cols * stride);
    if (kRankPass) {
        h->rankf.ensure(sizeof(float) * n_cols * stride);
        sort_segments(h, h->cols.as<double>(), (int)n_cols, stride, len, h->perm.as<int32_t>(), h->sorted.as<double>(), h->rankf.as<float>());
    } else {                                         // FP64 count pass: only the sweep order of the key columns is needed
        sort_segments(h, h->cols.as<double>(), (int)n_cols, stride, len, h->perm.as<int32_t>(), h->sorted.as<double>(), nullptr, &key_cols);
    }
}

using KdeFn = void (*)(const JobDesc*, int, const int*, double, double*);
KdeFn pick_kde(int dp) {
    switch (dp) {
        case 1: return kde_kernel<1, 2>;
        case 2: return kde_kernel<2, 2>;
        case 3: return kde_kernel<3, 2>;
        case 4: return kde_kernel<4, 2>;
        case 5: return kde_kernel<5, 2>;
        case 6: return kde_kernel<6, 2>;
    }
    return nullptr;
}

struct DebugPtrs { double* eps = nullptr; int32_t* cnt = nullptr; double* wsum = nullptr; double* weights = nullptr; };

// Evaluate `hj.size()` jobs whose columns are already in device memory.  out_dev: one double per job.
// dbg (optional, single-chunk only): host pointers that receive the per-point quantities of job 0..n-1.
void run_jobs(scribe_handle* h, std::vector<JobDesc>& hj, int n_max, int dx, int dy, int dz,
              const scribe_opts* o, double* out_dev, bool n_on_device, const DebugPtrs* dbg, bool* domain_error,
              bool have_order = false, int64_t nj_dev = -1, bool* used_sweep = nullptr)
{
    // nj_dev >= 0: the descriptors of nj_dev jobs (qoff included) were built in h->jobs on the device; `hj` is not used
    const bool host_desc = nj_dev < 0;
    const int64_t nj = host_desc ? (int64_t)hj.size() : nj_dev;
    if (nj == 0) return;
    if (!host_desc && (!n_on_device || dbg)) fail(SCRIBE_E_STATE, "device-built descriptors: internal misuse");
    const int est = o->estimator;
    const int D = dx + dy + dz;
    if (dx < 1 || dy < 1 || dz < 0 || D > SCRIBE_MAX_DIM) fail(SCRIBE_E_ARG, "unsupported dimensions (need dx,dy >= 1, dx+dy+dz <= 16)");
    if ((est == SCRIBE_EST_MI || est == SCRIBE_EST_UMI) && dz != 0) fail(SCRIBE_E_ARG, "mi/umi take no conditioning variable");
    if ((est == SCRIBE_EST_CMI || est == SCRIBE_EST_CUMI) && dz < 1) fail(SCRIBE_E_ARG, "cmi/cumi need a conditioning variable");
    const bool weighted = est >= SCRIBE_EST_UMI;
    const bool euclid = est == SCRIBE_EST_UMI;
    if (weighted && n_on_device && host_desc) fail(SCRIBE_E_ARG, "uniformised estimators are not available on the coupling path");
    if (weighted && !host_desc && o->density != SCRIBE_DENSITY_KDE) fail(SCRIBE_E_STATE, "device-built descriptors need the KDE weights");
    ensure_psi(h, n_max);

    const int k = o->k;
    const int kcap = (k + 1 <= 6) ? 6 : ((k + 1 <= 16) ? 16 : 0);
    int q = (n_max > 2 * TPB) ? 2 : 1;
    if (h->force_q) q = h->force_q;
    FusedFn fused = nullptr;
    SweepFn sweep = nullptr;
    StripFn strip = nullptr;
    bool use_strip2 = false, windowed = false;
    if (have_order && want_sweep(o, dx, dy, dz, n_max)) {
        // Measured on B200 (tools/tune.py matrix, est_ms per batch; first sweep / strip / strip2 / windowed sweep):
        //   cmi   N=1000 2.51/2.60/2.93/2.11   2000 6.41/6.28/6.76/5.19   5000 11.9/15.5/15.3/8.99   10000 15.1/18.1/17.7/11.1   20000 25.4/28.1/27.1/18.4
        //   cumi  N=1000 4.42/3.89/4.23/3.41   2000 11.1/9.05/9.18/7.87   5000 18.3/18.9/17.8/13.0   10000 23.1/22.4/20.8/16.2   20000 39.3/35.0/32.1/27.0
        // The strip forms visit 5-10x fewer candidates, but their lane-private loops run at the pace of the sparsest query
        // of the warp; the windowed sweep (sweep_window.cuh) is the fastest form everywhere, so AUTO takes it for both
        // estimators.  The first sweep and the strip forms stay selectable (SCRIBE_B200_SWEEP_V1=1, SCRIBE_PATH_STRIP).
        const bool auto_strip = o->path == SCRIBE_PATH_AUTO && dz == 1 && weighted && h->sweep_v1;   // the windowed sweep beats both strip forms
        if (o->path == SCRIBE_PATH_STRIP || auto_strip) strip = pick_strip(dx, dy, dz, kcap, weighted);
        use_strip2 = dz == 1 && !h->no_strip2 && (o->path == SCRIBE_PATH_STRIP || n_max >= 4096);
        sweep = (dz == 0) ? nullptr : pick_sweep(dx, dy, dz, kcap, weighted);          // also selects the KDE sweep and the visit counters
        if (!h->sweep_v1 || dz == 0) { sweep = pick_sweepw(dx, dy, dz, kcap, weighted); windowed = true; }
    }
    if (euclid && dbg) { sweep = nullptr; strip = nullptr; }        // per-point outputs of umi come from the generic kernel
    if (!sweep && o->path != SCRIBE_PATH_GENERIC && !euclid && kcap) fused = pick_fused(dx, dy, dz, kcap, q, weighted);
    const bool need_points = (fused == nullptr && sweep == nullptr) || dbg != nullptr;
    const int wpj = (n_max + 32 * SWEEP_Q - 1) / (32 * SWEEP_Q);        // warps per job, sweep kernels
    unsigned long long* visited = nullptr;
    if (sweep) {
        h->visited.ensure(4 * sizeof(unsigned long long));        // kNN / count / KDE visits, umi domain flag
        CU(cudaMemsetAsync(h->visited.p, 0, 4 * sizeof(unsigned long long), h->stream));
        visited = h->visited.as<unsigned long long>();
    }

    if (host_desc) for (int64_t j = 0; j < nj; j++) hj[j].qoff = j * (int64_t)n_max;
    const int bpj_f = (n_max + TPB * q - 1) / (TPB * q);       // CTAs per job, fused
    const int bpj_1 = (n_max + TPB - 1) / TPB;                  // CTAs per job, one query per thread
    const int bpj_kde = (n_max + TPB * 2 - 1) / (TPB * 2);
    if (nj * (int64_t)std::max(bpj_f, bpj_1) > 2000000000LL) fail(SCRIBE_E_ARG, "too many CTAs in one chunk");

    h->jobs.ensure(sizeof(JobDesc) * nj);
    JobDesc* dj = h->jobs.as<JobDesc>();
    if (!n_on_device) {      // coupling: the caller uploaded the descriptors and coupling_prep_kernel wrote n into them
        CU(cudaMemcpyAsync(dj, hj.data(), sizeof(JobDesc) * nj, cudaMemcpyHostToDevice, h->stream));
        h->st.h2d_bytes += sizeof(JobDesc) * nj;
    }

    PointOut po{nullptr, nullptr, nullptr};
    if (need_points) {
        h->eps.ensure(sizeof(double) * nj * n_max);
        h->cnt.ensure(sizeof(int32_t) * 4 * nj * n_max);
        po.eps = h->eps.as<double>(); po.cnt = h->cnt.as<int32_t>();
        if (weighted) { h->wsum.ensure(sizeof(double) * 2 * nj * n_max); po.wsum = h->wsum.as<double>(); }
        if (dbg) {   // make untouched slots recognisable
            CU(cudaMemsetAsync(po.cnt, 0xff, sizeof(int32_t) * 4 * nj * n_max, h->stream));
        }
    }

    CU(cudaEventRecord(h->ev_k0, h->stream));
    // ---- uniformisation weights -------------------------------------------------------------------
    if (weighted) {
        h->u.ensure(sizeof(double) * nj * n_max);
        double* u = h->u.as<double>();
        std::vector<int> pd;                                    // dims of P: x (umi) or x|z (cumi)
        for (int c = 0; c < dx; c++) pd.push_back(c);
        if (est == SCRIBE_EST_CUMI) for (int c = 0; c < dz; c++) pd.push_back(dx + dy + c);
        const int dp = (int)pd.size();
        if (o->density == SCRIBE_DENSITY_KDE) {
            KdeFn kf = pick_kde(dp);
            KdeSweepFn ks = (sweep && est == SCRIBE_EST_CUMI) ? pick_kde_sweep(dp) : nullptr;
            if (!kf) fail(SCRIBE_E_ARG, "KDE supports at most 6 density dimensions");
            h->dims.ensure(sizeof(int) * MAXD);
            CU(cudaMemcpyAsync(h->dims.p, pd.data(), sizeof(int) * dp, cudaMemcpyHostToDevice, h->stream));
            const double scale = std::sqrt(1.4426950408889634 / (2.0 * o->bw * o->bw));
            if (ks) {
                const int bps = (wpj + SWEEP_TPB / 32 - 1) / (SWEEP_TPB / 32);
                job_absmax_kernel<<<(unsigned)nj, 128, 0, h->stream>>>(dj, D);      // bound of the KDE's FP32 box pre-filter
                CU(cudaGetLastError());
                h->st.kernel_launches++;
                ks<<<(unsigned)(nj * bps), SWEEP_TPB, 0, h->stream>>>(dj, wpj, h->dims.as<int>(), scale, u, visited);
            } else {
                kf<<<(unsigned)(nj * bpj_kde), TPB, 0, h->stream>>>(dj, bpj_kde, h->dims.as<int>(), scale, u);
            }
            CU(cudaGetLastError());
            h->st.kernel_launches++; h->st.estimator_launches++;
        } else {
            // sub-jobs over the density subspace, kNN-only generic kernel, then k/N/r^d
            std::vector<JobDesc> sub(hj);
            for (auto& s : sub) { for (int c = 0; c < dp; c++) s.col[c] = hj[&s - sub.data()].col[pd[c]]; s.w = nullptr; }
            h->pjobs.ensure(sizeof(JobDesc) * nj);
            CU(cudaMemcpyAsync(h->pjobs.p, sub.data(), sizeof(JobDesc) * nj, cudaMemcpyHostToDevice, h->stream));
            h->r.ensure(sizeof(double) * nj * n_max);
            PointOut pr{h->r.as<double>(), nullptr, nullptr};
            ksg_generic_kernel<<<(unsigned)(nj * bpj_1), TPB, 0, h->stream>>>(h->pjobs.as<JobDesc>(), bpj_1, dp, 0, 0,
                                                                             o->k_density, 0, 1, pr);
            CU(cudaGetLastError());
            knn_density_kernel<<<(unsigned)(nj * bpj_1), TPB, 0, h->stream>>>(dj, bpj_1, o->k_density, dp, h->r.as<double>(), u);
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(h->stream));               // `sub` must outlive the copy
            h->st.kernel_launches += 2; h->st.estimator_launches++;
        }
        normalize_weights_kernel<<<(unsigned)nj, TPB, 0, h->stream>>>(dj, u);
        CU(cudaGetLastError());
        h->st.kernel_launches++;
        // point every job at its weights
        if (host_desc) {
            for (int64_t j = 0; j < nj; j++) hj[j].w = u + hj[j].qoff;
            CU(cudaMemcpyAsync(dj, hj.data(), sizeof(JobDesc) * nj, cudaMemcpyHostToDevice, h->stream));
            CU(cudaStreamSynchronize(h->stream));
        } else {
            set_weight_ptr_kernel<<<(unsigned)((nj + 255) / 256), 256, 0, h->stream>>>(dj, nj, u);
            CU(cudaGetLastError());
            h->st.kernel_launches++;
        }
    }

    if (euclid && sweep) {                                       // umi on the sweep: prefix sums of the weights in y order
        h->wcum.ensure(sizeof(double) * nj * ((size_t)n_max + 1));
        weight_prefix_kernel<<<(unsigned)nj, 256, 0, h->stream>>>(dj, h->wcum.as<double>(), (int64_t)n_max + 1);
        CU(cudaGetLastError());
        h->st.kernel_launches++;
    }

    // ---- estimator -----------------------------------------------------------------------------------
    if (strip) {
        const int spj = (n_max + STRIP - 1) / STRIP;                 // strips (= warps) per job
        const int64_t npad = (int64_t)spj * STRIP;
        const int64_t blk = strip_block_doubles(D, npad);
        h->strips.ensure(sizeof(double) * (size_t)nj * blk);
        h->partial.ensure(sizeof(double) * nj * spj);
        CU(cudaMemsetAsync(h->partial.p, 0, sizeof(double) * nj * spj, h->stream));
        const int64_t nwarps = nj * (int64_t)spj;
        const int bps = (spj + SWEEP_TPB / 32 - 1) / (SWEEP_TPB / 32);
        if (use_strip2) {                                            // all four counts as range queries on two strip images
            const int64_t blk2 = strip2_block_doubles(npad, weighted);
            h->strips.ensure(sizeof(double) * (size_t)nj * blk2);
            if (weighted) strip_sort2_kernel<true><<<(unsigned)((nwarps + 1) / 2), 64, 0, h->stream>>>(dj, (int)nj, spj, h->strips.as<double>(), blk2, npad);
            else          strip_sort2_kernel<false><<<(unsigned)((nwarps + 1) / 2), 64, 0, h->stream>>>(dj, (int)nj, spj, h->strips.as<double>(), blk2, npad);
            CU(cudaGetLastError());
            StripFn s2 = kcap == 6 ? (weighted ? ksg_strip2_kernel<6, true> : ksg_strip2_kernel<6, false>)
                                   : (weighted ? ksg_strip2_kernel<16, true> : ksg_strip2_kernel<16, false>);
            s2<<<(unsigned)(nj * bps), SWEEP_TPB, 0, h->stream>>>(dj, spj, k, h->psi.as<double>(), h->partial.as<double>(), po, visited);
            CU(cudaGetLastError());
        } else {
            StripSortFn ssort = pick_strip_sort(D);
            ssort<<<(unsigned)((nwarps + 1) / 2), 64, 0, h->stream>>>(dj, (int)nj, spj, h->strips.as<double>(), blk, npad);
            CU(cudaGetLastError());
            strip<<<(unsigned)(nj * bps), SWEEP_TPB, 0, h->stream>>>(dj, spj, k, h->psi.as<double>(), h->partial.as<double>(), po, visited);
            CU(cudaGetLastError());
        }
        finalize_mean_kernel<<<(unsigned)((nj + 255) / 256), 256, 0, h->stream>>>(dj, h->partial.as<double>(), spj, (int)nj, out_dev);
        CU(cudaGetLastError());
        h->st.kernel_launches += 3; h->st.estimator_launches++;
    } else if (sweep) {
        h->partial.ensure(sizeof(double) * nj * wpj);
        CU(cudaMemsetAsync(h->partial.p, 0, sizeof(double) * nj * wpj, h->stream));
        const int bps = (wpj + SWEEP_TPB / 32 - 1) / (SWEEP_TPB / 32);
        if (windowed) {                                              // FP32 pre-filter bound: per-job largest |coordinate|
            job_absmax_kernel<<<(unsigned)nj, 128, 0, h->stream>>>(dj, D);
            CU(cudaGetLastError());
            h->st.kernel_launches++;
        }
        sweep<<<(unsigned)(nj * bps), SWEEP_TPB, 0, h->stream>>>(dj, wpj, k, h->psi.as<double>(), h->partial.as<double>(), po, visited);
        CU(cudaGetLastError());
        finalize_mean_kernel<<<(unsigned)((nj + 255) / 256), 256, 0, h->stream>>>(dj, h->partial.as<double>(), wpj, (int)nj, out_dev);
        CU(cudaGetLastError());
        h->st.kernel_launches += 2; h->st.estimator_launches++;
    } else if (fused) {
        h->partial.ensure(sizeof(double) * nj * bpj_f);
        CU(cudaMemsetAsync(h->partial.p, 0, sizeof(double) * nj * bpj_f, h->stream));
        const int mode = (est == SCRIBE_EST_MI) ? MODE_MI : MODE_CMI;
        const int cap = std::min((n_max + 3) & ~3, FUSED_TILE_MAX);
        const size_t smem = sizeof(double) * (size_t)cap * (D + (weighted ? 1 : 0));
        if (smem > 40 * 1024) CU(cudaFuncSetAttribute((const void*)fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fused<<<(unsigned)(nj * bpj_f), TPB, smem, h->stream>>>(dj, bpj_f, k, mode, cap, h->psi.as<double>(), h->partial.as<double>(), po);
        CU(cudaGetLastError());
        finalize_mean_kernel<<<(unsigned)((nj + 255) / 256), 256, 0, h->stream>>>(dj, h->partial.as<double>(), bpj_f, (int)nj, out_dev);
        CU(cudaGetLastError());
        h->st.kernel_launches += 2; h->st.estimator_launches++;
    } else {
        if (k + 1 > GEN_KCAP) fail(SCRIBE_E_ARG, "k too large");
        h->flag.ensure(sizeof(int));
        CU(cudaMemsetAsync(h->flag.p, 0, sizeof(int), h->stream));
        ksg_generic_kernel<<<(unsigned)(nj * bpj_1), TPB, 0, h->stream>>>(dj, bpj_1, dx, dy, dz, k, euclid ? 1 : 0, 0, po);
        CU(cudaGetLastError());
        ksg_finish_kernel<<<(unsigned)nj, TPB, 0, h->stream>>>(dj, est, h->psi.as<double>(), po, out_dev, h->flag.as<int>());
        CU(cudaGetLastError());
        h->st.kernel_launches += 2; h->st.estimator_launches++;
        if (euclid) {
            int flag = 0;
            CU(cudaMemcpyAsync(&flag, h->flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaStreamSynchronize(h->stream));
            if (flag && domain_error) *domain_error = true;
        }
    }
    CU(cudaEventRecord(h->ev_k1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1));
    h->st.estimator_ms += ms;
    h->st.jobs += nj;
    if (sweep) {
        unsigned long long v[4] = {0, 0, 0, 0};
        CU(cudaMemcpy(v, h->visited.p, sizeof(v), cudaMemcpyDeviceToHost));
        if (euclid) {
            if (v[3] && domain_error) *domain_error = true;
            if (used_sweep) *used_sweep = true;
        }
        h->st.visited_knn += (int64_t)v[0]; h->st.visited_count += (int64_t)v[1]; h->st.visited_kde += (int64_t)v[2];
        if (weighted && v[2] == 0 && host_desc) for (int64_t j = 0; j < nj; j++) h->st.visited_kde += (int64_t)hj[j].n * hj[j].n;
    } else if (!n_on_device) {
        for (int64_t j = 0; j < nj; j++) {
            const int64_t nn = (int64_t)hj[j].n * hj[j].n;
            h->st.visited_knn += nn; h->st.visited_count += nn;
            if (weighted && o->density == SCRIBE_DENSITY_KDE) h->st.visited_kde += nn;
        }
    }

    if (dbg) {
        const int n = hj[0].n;
        if (dbg->eps)  CU(cudaMemcpy(dbg->eps, po.eps, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (dbg->cnt)  CU(cudaMemcpy(dbg->cnt, po.cnt, sizeof(int32_t) * 4 * n, cudaMemcpyDeviceToHost));
        if (dbg->wsum && weighted) CU(cudaMemcpy(dbg->wsum, po.wsum, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
        if (dbg->weights && weighted) CU(cudaMemcpy(dbg->weights, h->u.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    }
}

double vd_host(int d) { return 0.5 * d * std::log(3.14159265358979323846) - std::lgamma(0.5 * d + 1.0); }

void begin_call(scribe_handle* h) {
    std::memset(&h->st, 0, sizeof(h->st));
    CU(cudaEventRecord(h->ev_call0, h->stream));
}
void end_call(scribe_handle* h) {
    CU(cudaEventRecord(h->ev_call1, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev_call0, h->ev_call1));
    h->st.total_ms = ms;
}

// single job from host columns (estimator-level entry points)
void single_job(scribe_handle* h, const double* x, const double* y, const double* z, int64_t n, int dx, int dy, int dz,
                const scribe_opts* o, double* out, const DebugPtrs* dbg)
{
    check_opts(o);
    if (n < 1) fail(SCRIBE_E_ARG, "need at least one sample");
    if (n > 2000000) fail(SCRIBE_E_ARG, "n too large for a single job");
    if (!x || !y || (dz > 0 && !z)) fail(SCRIBE_E_ARG, "NULL input");
    if (o->estimator == SCRIBE_EST_UMI && !(o->k <= n - 1)) fail(SCRIBE_E_LENGTH, "Set k smaller than num. samples - 1");
    const int D = dx + dy + dz;
    if (D > SCRIBE_MAX_DIM || dx < 1 || dy < 1 || dz < 0) fail(SCRIBE_E_ARG, "unsupported dimensions");
    begin_call(h);
    h->single.ensure(sizeof(double) * (size_t)D * n + sizeof(double));
    double* base = h->single.as<double>();
    CU(cudaMemcpyAsync(base, x, sizeof(double) * dx * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(base + (size_t)dx * n, y, sizeof(double) * dy * n, cudaMemcpyHostToDevice, h->stream));
    if (dz) CU(cudaMemcpyAsync(base + (size_t)(dx + dy) * n, z, sizeof(double) * dz * n, cudaMemcpyHostToDevice, h->stream));
    h->st.h2d_bytes += sizeof(double) * D * n;
    std::vector<JobDesc> hj(1);
    std::memset(&hj[0], 0, sizeof(JobDesc));
    for (int c = 0; c < D; c++) hj[0].col[c] = base + (size_t)c * n;
    hj[0].n = (int)n;
    h->out.ensure(sizeof(double));
    bool dom = false;
    bool ordered = false;
    if (want_sweep(o, dx, dy, dz, (int)n)) {                 // all D columns in ascending order: sweep order = the last one (y for mi)
        h->sorted.ensure(sizeof(double) * D * n); h->perm.ensure(sizeof(int32_t) * D * n); h->rankf.ensure(sizeof(float) * D * n);
        sort_segments(h, base, D, n, std::vector<int>((size_t)D, (int)n), h->perm.as<int32_t>(), h->sorted.as<double>(), h->rankf.as<float>());
        hj[0].perm = h->perm.as<int32_t>() + (size_t)(D - 1) * n;
        for (int c = 0; c < D - 1 && c < 4; c++) {
            hj[0].srt[c] = h->sorted.as<double>() + (size_t)c * n; hj[0].rk[c] = h->rankf.as<float>() + (size_t)c * n; hj[0].srt_n[c] = (int)n;
        }
        ordered = true;
    }
    bool umi_mean = false;                                   // the sweep returns the mean of the per-point terms, the generic path their sum
    run_jobs(h, hj, (int)n, dx, dy, dz, o, h->out.as<double>(), false, dbg, &dom, ordered, -1, &umi_mean);
    h->st.point_pairs += n * n;
    double v = 0.0;
    CU(cudaMemcpyAsync(&v, h->out.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    end_call(h);
    h->st.d2h_bytes += sizeof(double);
    if (o->estimator == SCRIBE_EST_UMI) {
        if (dom) fail(SCRIBE_E_DOMAIN, "math domain error");
        // information_estimators.py:264,:274-276
        const double N = (double)n;
        v = digamma_f64((double)o->k) + 2.0 * std::log(N - 1.0) - digamma_f64(N) + vd_host(dx) + vd_host(dy) - vd_host(dx + dy)
            - (umi_mean ? v : v / N);
    }
    if (out) *out = v;
}

struct ColKey {
    int32_t gene, shift, tau, diff; uint64_t scale_bits;
    bool operator==(const ColKey& o) const { return gene == o.gene && shift == o.shift && tau == o.tau && diff == o.diff && scale_bits == o.scale_bits; }
};
struct ColKeyHash {
    size_t operator()(const ColKey& k) const {
        uint64_t h = (uint64_t)(uint32_t)k.gene * 0x9E3779B97F4A7C15ULL;
        h ^= ((uint64_t)(uint32_t)k.shift << 32 | (uint32_t)k.tau) * 0xC2B2AE3D27D4EB4FULL;
        h ^= (k.scale_bits + (uint64_t)k.diff) * 0x165667B19E3779F9ULL;
        return (size_t)(h ^ (h >> 29));
    }
};

static_assert(sizeof(RawJob) == sizeof(scribe_job), "RawJob mirrors scribe_job");

// SCRIBE_B200_TRACE=1: host wall-clock per phase of a call, to stderr (diagnostics only)
struct Trace {
    bool on; std::chrono::steady_clock::time_point t;
    Trace() : on(std::getenv("SCRIBE_B200_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[scribe trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

void lagged_jobs(scribe_handle* h, const scribe_job* jobs, int64_t n_jobs, const scribe_opts* o, int differential,
                 const double* scales, double* out, bool out_is_device)
{
    check_opts(o);
    if (o->estimator != SCRIBE_EST_CMI && o->estimator != SCRIBE_EST_CUMI) fail(SCRIBE_E_ARG, "lagged jobs use cmi or cumi");
    if (h->n_genes == 0) fail(SCRIBE_E_STATE, "no expression matrix uploaded");
    if (n_jobs < 0 || (n_jobs > 0 && (!jobs || !out))) fail(SCRIBE_E_ARG, "NULL jobs/out");
    begin_call(h);
    Trace tr;
    const int n_runs = (int)h->run_off.size() - 1;
    // jobs are processed in chunks of equal conditioning-set size (equal joint dimension)
    int64_t pos = 0;
    std::vector<double> host_out;
    double* out_dev = out;
    if (!out_is_device) { h->out.ensure(sizeof(double) * std::max<int64_t>(n_jobs, 1)); out_dev = h->out.as<double>(); }

    auto n_for_tau = [&](int tau) { int64_t n = 0; for (int r = 0; r < n_runs; r++) { int64_t T = h->run_off[r + 1] - h->run_off[r]; if (T > tau) n += T - tau; } return n; };

    // ---- plain RDI (no conditioning set, no per-job scaling, at most 8 distinct delays, most genes in use): the lagged
    // columns of EVERY gene and delay are laid out by formula, gathered and sorted once per call, and the descriptors are
    // built on the device from the caller's job table -- no per-job host work beyond validation.
    const bool weighted_knn = o->estimator == SCRIBE_EST_CUMI && o->density != SCRIBE_DENSITY_KDE;
    if (n_jobs > 0 && !scales && !weighted_knn && !h->no_fast_rdi) {
        RdiLayout lay; lay.n_delays = 0;
        bool eligible = true;
        const int64_t G = h->n_genes;
        std::vector<int64_t> per_delay(8, 0);
        for (int64_t j = 0; j < n_jobs && eligible; j++) {
            const scribe_job& jb = jobs[j];
            if (jb.n_cond != 0) { eligible = false; break; }
            if (jb.src < 0 || jb.src >= G || jb.dst < 0 || jb.dst >= G) fail(SCRIBE_E_ARG, "gene index out of range");
            if (jb.delay < 1) fail(SCRIBE_E_ARG, "delay must be >= 1");
            int di = -1;
            for (int t = 0; t < lay.n_delays; t++) if (lay.delays[t] == jb.delay) { di = t; break; }
            if (di < 0) {
                if (lay.n_delays == 8) { eligible = false; break; }
                di = lay.n_delays++; lay.delays[di] = jb.delay;
            }
            per_delay[di]++;
        }
        int64_t n_max = 0;
        if (eligible) {
            for (int t = 0; t < lay.n_delays; t++) {
                const int64_t n = n_for_tau(lay.delays[t]);
                if (n < 1) fail(SCRIBE_E_LENGTH, "delay leaves no samples");
                if (n > 2000000) fail(SCRIBE_E_ARG, "too many cells");
                lay.n_of[t] = (int32_t)n; n_max = std::max(n_max, n);
            }
            for (int t = lay.n_delays; t < 8; t++) { lay.delays[t] = 0; lay.n_of[t] = 0; }
        }
        const int64_t ncols = 3 * G * lay.n_delays, nk = G * lay.n_delays;
        if (eligible && (n_jobs < nk || (size_t)ncols * n_max * 8 > h->ws_budget / 2 || ncols * n_max >= ((int64_t)1 << 31))) eligible = false;
        if (eligible) {
            const int64_t stride = n_max;
            std::vector<ColSpec> specs((size_t)ncols);
            for (int t = 0; t < lay.n_delays; t++)
                for (int64_t g = 0; g < G; g++) {
                    const int d = lay.delays[t];
                    ColSpec* sp = &specs[(size_t)(t * G + g) * 3];
                    sp[0] = ColSpec{(int32_t)g, 0, d, 0, 1.0};                                   // x   (cn:147)
                    sp[1] = ColSpec{(int32_t)g, d, d, differential ? 1 : 0, 1.0};              // y   (cn:151-153)
                    sp[2] = ColSpec{(int32_t)g, d - 1, d, 0, 1.0};                              // z   (cn:149)
                }
            h->cols.ensure(sizeof(double) * ncols * stride);
            h->specs.ensure(sizeof(ColSpec) * ncols);
            CU(cudaMemcpyAsync(h->specs.p, specs.data(), sizeof(ColSpec) * ncols, cudaMemcpyHostToDevice, h->stream));
            h->st.h2d_bytes += sizeof(ColSpec) * ncols;
            dim3 grid((unsigned)std::min<int64_t>((stride + TPB - 1) / TPB, 64), 1);
            for (int64_t c0 = 0; c0 < ncols; c0 += 65535) {
                grid.y = (unsigned)std::min<int64_t>(65535, ncols - c0);
                gather_lagged_kernel<<<grid, TPB, 0, h->stream>>>(h->E.as<double>(), h->n_cells, h->run_off_d.as<int64_t>(), n_runs,
                    h->specs.as<ColSpec>() + c0, h->cols.as<double>() + c0 * stride, stride);
                CU(cudaGetLastError());
                h->st.kernel_launches++;
            }
            h->rawjobs.ensure(sizeof(scribe_job) * n_jobs);
            CU(cudaMemcpyAsync(h->rawjobs.p, jobs, sizeof(scribe_job) * n_jobs, cudaMemcpyHostToDevice, h->stream));
            h->st.h2d_bytes += sizeof(scribe_job) * n_jobs;
            tr.lap("fast: validate + gather");
            const bool ordered = want_sweep(o, 1, 1, 1, (int)n_max);
            if (ordered) {                                    // every lagged column in ascending order: sweep order (z), rank windows (x, y)
                std::vector<int> collen((size_t)ncols);
                for (int64_t c = 0; c < ncols; c++) collen[c] = lay.n_of[c / (3 * G)];
                std::vector<int> keys((size_t)nk);
                for (int64_t kq = 0; kq < nk; kq++) keys[kq] = (int)(kq * 3 + 2);
                sort_all_columns(h, ncols, collen, stride, keys);
            }
            tr.lap("fast: key select + argsort");
            for (int t = 0; t < lay.n_delays; t++) h->st.point_pairs += per_delay[t] * (int64_t)lay.n_of[t] * lay.n_of[t];
            const bool strips = o->path == SCRIBE_PATH_STRIP || (o->path == SCRIBE_PATH_AUTO && h->sweep_v1 && o->estimator == SCRIBE_EST_CUMI);
            const size_t per_point = (o->estimator == SCRIBE_EST_CUMI ? 8 : 0) + (o->path == SCRIBE_PATH_GENERIC ? 40 : 0) + (strips ? 72 : 0);
            int64_t chunk = 262144;
            if (per_point) chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, (int64_t)((h->ws_budget / 2) / (per_point * (size_t)n_max))));
            std::vector<JobDesc> none;
            for (int64_t p0 = 0; p0 < n_jobs; p0 += chunk) {
                const int64_t nj = std::min(chunk, n_jobs - p0);
                h->jobs.ensure(sizeof(JobDesc) * nj);
                build_rdi_jobs_kernel<<<(unsigned)((nj + TPB - 1) / TPB), TPB, 0, h->stream>>>(
                    reinterpret_cast<const RawJob*>(h->rawjobs.p) + p0, nj, h->jobs.as<JobDesc>(), h->cols.as<double>(), stride,
                    ordered ? h->perm.as<int32_t>() : nullptr, (ordered && kRankPass) ? h->sorted.as<double>() : nullptr,
                    (ordered && kRankPass) ? h->rankf.as<float>() : nullptr, (int)G, lay, n_max);
                CU(cudaGetLastError());
                h->st.kernel_launches++;
                run_jobs(h, none, (int)n_max, 1, 1, 1, o, out_dev + p0, true, nullptr, nullptr, ordered, nj);
            }
            tr.lap("fast: run_jobs");
            pos = n_jobs;
        }
    }

    while (pos < n_jobs) {
        const int L = jobs[pos].n_cond;
        if (L < 0 || L > SCRIBE_MAX_COND) fail(SCRIBE_E_ARG, "n_cond out of range");
        const int dz = L + 1, D = dz + 2;
        // chunk: same n_cond, bounded workspace
        // distinct lagged columns of the chunk: open-addressing table over `specs` (no per-entry allocation; the loop below is
        // on the critical path of every call -- the GPU is idle while it runs)
        std::vector<int32_t> table(1u << 14, -1);
        std::vector<uint64_t> spec_scale;                                                     // scale bits per spec (part of the key)
        ColKey last_key[2 + SCRIBE_MAX_COND + 1]; int last_id[2 + SCRIBE_MAX_COND + 1];     // per-role one-entry cache: jobs arrive
        for (auto& v : last_id) v = -1;                                                      // target-major, so y/z columns repeat
        std::vector<ColSpec> specs;
        std::vector<int32_t> n_of;        // samples per job
        std::vector<int> colidx;          // D per job
        const int64_t room = std::min<int64_t>(n_jobs - pos, 262144);
        n_of.reserve(room); colidx.reserve(room * D); specs.reserve(4096); spec_scale.reserve(4096);
        auto rehash = [&]() {
            std::vector<int32_t> bigger(table.size() * 4, -1);
            const size_t mask = bigger.size() - 1;
            for (size_t id = 0; id < specs.size(); id++) {
                const ColSpec& sp = specs[id];
                size_t slot = ColKeyHash{}(ColKey{sp.gene, sp.shift, sp.tau, sp.diff, spec_scale[id]}) & mask;
                while (bigger[slot] >= 0) slot = (slot + 1) & mask;
                bigger[slot] = (int32_t)id;
            }
            table.swap(bigger);
        };
        int n_max = 0;
        int64_t end = pos;
        // per-point workspace: cumi weights; generic path eps/counts/sums; strip images only where the strip kernels run
        const bool strips = o->path == SCRIBE_PATH_STRIP || (o->path == SCRIBE_PATH_AUTO && h->sweep_v1 && o->estimator == SCRIBE_EST_CUMI);
        size_t per_point = (o->estimator == SCRIBE_EST_CUMI ? 8 : 0) + (o->path == SCRIBE_PATH_GENERIC ? 40 : 0)
                         + (strips ? (size_t)(std::max(D, 8) * 8 + 8) : 0);
        int tau_cached = -1; int64_t n_cached = 0;
        int64_t pairs = 0;
        while (end < n_jobs && jobs[end].n_cond == L && (end - pos) < 262144) {
            const scribe_job& jb = jobs[end];
            if (jb.src < 0 || jb.src >= h->n_genes || jb.dst < 0 || jb.dst >= h->n_genes) fail(SCRIBE_E_ARG, "gene index out of range");
            if (jb.delay < 1) fail(SCRIBE_E_ARG, "delay must be >= 1");
            int tau = jb.delay;
            for (int c = 0; c < L; c++) {
                if (jb.cond_gene[c] < 0 || jb.cond_gene[c] >= h->n_genes) fail(SCRIBE_E_ARG, "conditioning gene out of range");
                if (jb.cond_delay[c] < 1) fail(SCRIBE_E_ARG, "conditioning delay must be >= 1");
                tau = std::max(tau, jb.cond_delay[c]);
            }
            if (tau != tau_cached) { n_cached = n_for_tau(tau); tau_cached = tau; }
            const int64_t n = n_cached;
            if (n < 1) fail(SCRIBE_E_LENGTH, "delay leaves no samples");
            if (n > 2000000) fail(SCRIBE_E_ARG, "too many cells");
            const int64_t new_nmax = std::max<int64_t>(n_max, n);
            // workspace estimate if this job is added
            const size_t ws = (size_t)(specs.size() + D) * new_nmax * 8 + (size_t)(end - pos + 1) * new_nmax * per_point;
            if (end > pos && ws > h->ws_budget) break;
            n_max = (int)new_nmax;
            const double sx = scales ? scales[3 * end] : 1.0, sy = scales ? scales[3 * end + 1] : 1.0, sz = scales ? scales[3 * end + 2] : 1.0;
            auto get = [&](int role, int gene, int shift, int diff, double sc) {
                ColKey key{gene, shift, tau, diff, 0};
                std::memcpy(&key.scale_bits, &sc, 8);
                if (last_id[role] >= 0 && last_key[role] == key) return last_id[role];
                const size_t mask = table.size() - 1;
                size_t slot = ColKeyHash{}(key) & mask;
                int id = -1;
                while (table[slot] >= 0) {
                    const int c = table[slot];
                    const ColSpec& sp = specs[c];
                    if (sp.gene == gene && sp.shift == shift && sp.tau == tau && sp.diff == diff && spec_scale[c] == key.scale_bits) { id = c; break; }
                    slot = (slot + 1) & mask;
                }
                if (id < 0) {
                    id = (int)specs.size();
                    specs.push_back(ColSpec{gene, shift, tau, diff, sc});
                    spec_scale.push_back(key.scale_bits);
                    table[slot] = id;
                    if (specs.size() * 2 > table.size()) rehash();
                }
                last_key[role] = key; last_id[role] = id;
                return id;
            };
            colidx.push_back(get(0, jb.src, tau - jb.delay, 0, sx));                 // x   (cn:147,:266)
            colidx.push_back(get(1, jb.dst, tau, differential ? 1 : 0, sy));         // y   (cn:151-153,:269-271)
            for (int c = L - 1; c >= 0; c--)                                         // z: cond[L-1] .. cond[0], then dst's past (cn:273-278)
                colidx.push_back(get(2 + c, jb.cond_gene[c], tau - jb.cond_delay[c], 0, sz));
            colidx.push_back(get(2 + SCRIBE_MAX_COND, jb.dst, tau - 1, 0, sz));      //     (cn:149,:267)
            n_of.push_back((int32_t)n);
            pairs += n * n;
            end++;
        }
        h->st.point_pairs += pairs;
        std::vector<JobDesc> hj((size_t)(end - pos));                                // value-initialised: all fields zero
        for (size_t j = 0; j < hj.size(); j++) hj[j].n = n_of[j];
        const int64_t nj = end - pos;
        const int64_t stride = n_max;
        tr.lap("job loop (host)");
        h->cols.ensure(sizeof(double) * specs.size() * stride);
        h->specs.ensure(sizeof(ColSpec) * specs.size());
        CU(cudaMemcpyAsync(h->specs.p, specs.data(), sizeof(ColSpec) * specs.size(), cudaMemcpyHostToDevice, h->stream));
        h->st.h2d_bytes += sizeof(ColSpec) * specs.size();
        {
            dim3 grid((unsigned)std::min<int64_t>((stride + TPB - 1) / TPB, 64), (unsigned)specs.size());
            // gridDim.y limit is 65535
            for (size_t c0 = 0; c0 < specs.size(); c0 += 65535) {
                grid.y = (unsigned)std::min<size_t>(65535, specs.size() - c0);
                gather_lagged_kernel<<<grid, TPB, 0, h->stream>>>(h->E.as<double>(), h->n_cells, h->run_off_d.as<int64_t>(), n_runs,
                    h->specs.as<ColSpec>() + c0, h->cols.as<double>() + c0 * stride, stride);
                CU(cudaGetLastError());
                h->st.kernel_launches++;
            }
        }
        for (int64_t j = 0; j < nj; j++)
            for (int c = 0; c < D; c++) hj[j].col[c] = h->cols.as<double>() + (size_t)colidx[j * D + c] * stride;
        tr.lap("gather launch + col pointers");
        bool ordered = false;
        if (want_sweep(o, 1, 1, dz, n_max)) {
            // every distinct lagged column of the chunk is sorted once and shared by the jobs that read it: the key column
            // (the target's own past, last z column) gives the sweep order, the others their rank windows
            if ((int64_t)specs.size() * stride >= ((int64_t)1 << 31)) fail(SCRIBE_E_ARG, "chunk too large to sort");
            std::vector<int> collen(specs.size(), 0);
            for (int64_t j = 0; j < nj; j++)
                for (int c = 0; c < D; c++) collen[colidx[j * D + c]] = hj[j].n;      // equal tau (part of the column key) => equal n
            std::vector<int> keys; std::vector<char> is_key(specs.size(), 0);
            for (int64_t j = 0; j < nj; j++) { const int kc = colidx[j * D + D - 1]; if (!is_key[kc]) { is_key[kc] = 1; keys.push_back(kc); } }
            sort_all_columns(h, (int64_t)specs.size(), collen, stride, keys);
            for (int64_t j = 0; j < nj; j++) {
                hj[j].perm = h->perm.as<int32_t>() + (size_t)colidx[j * D + D - 1] * stride;
                for (int c = 0; kRankPass && c < D - 1 && c < 4; c++) {
                    hj[j].srt[c] = h->sorted.as<double>() + (size_t)colidx[j * D + c] * stride;
                    hj[j].rk[c] = h->rankf.as<float>() + (size_t)colidx[j * D + c] * stride;
                    hj[j].srt_n[c] = hj[j].n;
                }
            }
            ordered = true;
        }
        tr.lap("key select + argsort");
        run_jobs(h, hj, n_max, 1, 1, dz, o, out_dev + pos, false, nullptr, nullptr, ordered);
        tr.lap("run_jobs");
        pos = end;
    }
    if (!out_is_device && n_jobs > 0) {
        CU(cudaMemcpyAsync(out, out_dev, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost, h->stream));
        h->st.d2h_bytes += sizeof(double) * n_jobs;
    }
    end_call(h);
    tr.lap("result copy + end");
}

void coupling_jobs(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs, int drop_zero,
                   const scribe_opts* o, double* out, bool out_is_device, int32_t* n_eff)
{
    check_opts(o);
    if (o->estimator != SCRIBE_EST_CMI) fail(SCRIBE_E_ARG, "dynamics coupling uses cmi");
    if (h->cg == 0) fail(SCRIBE_E_STATE, "no coupling matrices uploaded");
    if (n_jobs < 0 || (n_jobs > 0 && (!src || !dst || !out))) fail(SCRIBE_E_ARG, "NULL jobs/out");
    for (int64_t j = 0; j < n_jobs; j++)
        if (src[j] < 0 || src[j] >= h->cg || dst[j] < 0 || dst[j] >= h->cg) fail(SCRIBE_E_ARG, "gene index out of range");
    begin_call(h);
    const int64_t N = h->cc;
    // the sweep needs every gene's rows sorted in one segmented pass; a matrix too large for that (genes x cells >= 2^31) takes
    // the dense kernel instead of failing
    const bool ordered = want_sweep(o, 1, 1, 1, (int)N) && h->cg * N < ((int64_t)1 << 31);
    if (ordered && !h->permS_valid) {                    // every gene's S and Y row in ascending order, once per upload
        const std::vector<int> len((size_t)h->cg, (int)N);
        h->permS.ensure(sizeof(int32_t) * h->cg * N); h->sortedS.ensure(sizeof(double) * h->cg * N); h->rankS.ensure(sizeof(float) * h->cg * N);
        h->permY.ensure(sizeof(int32_t) * h->cg * N); h->sortedY.ensure(sizeof(double) * h->cg * N); h->rankY.ensure(sizeof(float) * h->cg * N);
        sort_segments(h, h->S.as<double>(), (int)h->cg, N, len, h->permS.as<int32_t>(), h->sortedS.as<double>(), kRankPass ? h->rankS.as<float>() : nullptr);
        if (kRankPass) sort_segments(h, h->Y.as<double>(), (int)h->cg, N, len, h->permY.as<int32_t>(), h->sortedY.as<double>(), h->rankY.as<float>());
        h->permS_valid = true;
    }
    double* out_dev = out;
    if (!out_is_device) { h->out.ensure(sizeof(double) * std::max<int64_t>(n_jobs, 1)); out_dev = h->out.as<double>(); }
    const bool strips = o->path == SCRIBE_PATH_STRIP;                                                 // 3 compacted columns + order (+ strip images)
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(262144, (int64_t)(h->ws_budget / ((strips ? 13 : 5) * N * 8))));
    h->idx32a.ensure(sizeof(int32_t) * std::max<int64_t>(n_jobs, 1));
    h->idx32b.ensure(sizeof(int32_t) * std::max<int64_t>(n_jobs, 1));
    h->neff.ensure(sizeof(int32_t) * std::max<int64_t>(n_jobs, 1));
    CU(cudaMemcpyAsync(h->idx32a.p, src, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->idx32b.p, dst, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, h->stream));
    h->st.h2d_bytes += 2 * sizeof(int32_t) * n_jobs;
    std::vector<int32_t> neff_host(n_jobs);
    for (int64_t pos = 0; pos < n_jobs; pos += chunk) {
        const int64_t nj = std::min(chunk, n_jobs - pos);
        h->cols.ensure(sizeof(double) * 3 * N * nj);
        const bool ranked = ordered && kRankPass;
        if (ranked) h->crank.ensure(sizeof(float) * 2 * N * nj);
        std::vector<JobDesc> hj(nj);
        for (int64_t j = 0; j < nj; j++) {
            std::memset(&hj[j], 0, sizeof(JobDesc));
            double* b = h->cols.as<double>() + (size_t)j * 3 * N;
            hj[j].col[0] = b; hj[j].col[1] = b + N; hj[j].col[2] = b + 2 * N;
            hj[j].n = (int)N; hj[j].qoff = j * N;
            if (ranked) {                                // ranks of the kept cells among ALL cells of the source's S row / the target's Y row
                float* r = h->crank.as<float>() + (size_t)j * 2 * N;
                hj[j].rk[0] = r; hj[j].rk[1] = r + N;
                hj[j].srt[0] = h->sortedS.as<double>() + (size_t)src[pos + j] * N; hj[j].srt[1] = h->sortedY.as<double>() + (size_t)dst[pos + j] * N;
                hj[j].srt_n[0] = hj[j].srt_n[1] = (int)N;
            }
        }
        h->jobs.ensure(sizeof(JobDesc) * nj);
        CU(cudaMemcpyAsync(h->jobs.p, hj.data(), sizeof(JobDesc) * nj, cudaMemcpyHostToDevice, h->stream));
        h->st.h2d_bytes += sizeof(JobDesc) * nj;
        coupling_prep_kernel<<<(unsigned)nj, TPB, 0, h->stream>>>(h->S.as<double>(), h->Y.as<double>(), N,
            ordered ? h->permS.as<int32_t>() : nullptr, ranked ? h->rankS.as<float>() : nullptr, ranked ? h->rankY.as<float>() : nullptr,
            ranked ? h->crank.as<float>() : nullptr, h->idx32a.as<int32_t>() + pos, h->idx32b.as<int32_t>() + pos, drop_zero, h->cols.as<double>(), N,
            h->jobs.as<JobDesc>(), h->neff.as<int32_t>() + pos);
        CU(cudaGetLastError());
        h->st.kernel_launches++;
        CU(cudaMemcpyAsync(neff_host.data() + pos, h->neff.as<int32_t>() + pos, sizeof(int32_t) * nj, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->st.d2h_bytes += sizeof(int32_t) * nj;
        for (int64_t j = 0; j < nj; j++) h->st.point_pairs += (int64_t)neff_host[pos + j] * neff_host[pos + j];
        run_jobs(h, hj, (int)N, 1, 1, 1, o, out_dev + pos, true, nullptr, nullptr, ordered);
        if (!ordered) for (int64_t j = 0; j < nj; j++) {
            const int64_t nn = (int64_t)neff_host[pos + j] * neff_host[pos + j];
            h->st.visited_knn += nn; h->st.visited_count += nn;
        }
    }
    if (n_eff) std::memcpy(n_eff, neff_host.data(), sizeof(int32_t) * n_jobs);
    if (!out_is_device && n_jobs > 0) {
        CU(cudaMemcpyAsync(out, out_dev, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost, h->stream));
        h->st.d2h_bytes += sizeof(double) * n_jobs;
    }
    end_call(h);
}

void mi_jobs(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs, const scribe_opts* o, double* out)
{
    check_opts(o);
    if (o->estimator != SCRIBE_EST_MI) fail(SCRIBE_E_ARG, "mi jobs use the mi estimator");
    if (h->n_genes == 0) fail(SCRIBE_E_STATE, "no expression matrix uploaded");
    if (n_jobs < 0 || (n_jobs > 0 && (!src || !dst || !out))) fail(SCRIBE_E_ARG, "NULL jobs/out");
    if (h->n_cells > 2000000) fail(SCRIBE_E_ARG, "too many cells");
    begin_call(h);
    h->out.ensure(sizeof(double) * std::max<int64_t>(n_jobs, 1));
    const int64_t N = h->n_cells;
    const bool ordered = want_sweep(o, 1, 1, 0, (int)N) && (int64_t)h->n_genes * N < ((int64_t)1 << 31);
    if (ordered && !h->sortE_valid) {                    // every gene's row: argsort, sorted values and ranks, once per upload
        h->permE.ensure(sizeof(int32_t) * h->n_genes * N); h->sortedE.ensure(sizeof(double) * h->n_genes * N); h->rankE.ensure(sizeof(float) * h->n_genes * N);
        sort_segments(h, h->E.as<double>(), (int)h->n_genes, N, std::vector<int>((size_t)h->n_genes, (int)N), h->permE.as<int32_t>(),
                      h->sortedE.as<double>(), h->rankE.as<float>());
        h->sortE_valid = true;
    }
    for (int64_t pos = 0; pos < n_jobs; pos += 262144) {
        const int64_t nj = std::min<int64_t>(262144, n_jobs - pos);
        std::vector<JobDesc> hj(nj);
        for (int64_t j = 0; j < nj; j++) {
            const int a = src[pos + j], b = dst[pos + j];
            if (a < 0 || a >= h->n_genes || b < 0 || b >= h->n_genes) fail(SCRIBE_E_ARG, "gene index out of range");
            std::memset(&hj[j], 0, sizeof(JobDesc));
            hj[j].col[0] = h->E.as<double>() + (size_t)a * N;
            hj[j].col[1] = h->E.as<double>() + (size_t)b * N;
            hj[j].n = (int)N;
            if (ordered) {
                hj[j].perm = h->permE.as<int32_t>() + (size_t)b * N;
                hj[j].srt[0] = h->sortedE.as<double>() + (size_t)a * N; hj[j].rk[0] = h->rankE.as<float>() + (size_t)a * N; hj[j].srt_n[0] = (int)N;
            }
            h->st.point_pairs += N * N;
        }
        run_jobs(h, hj, (int)N, 1, 1, 0, o, h->out.as<double>() + pos, false, nullptr, nullptr, ordered);
    }
    if (n_jobs > 0) {
        CU(cudaMemcpyAsync(out, h->out.p, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost, h->stream));
        h->st.d2h_bytes += sizeof(double) * n_jobs;
    }
    end_call(h);
}

template <class F> int guarded(scribe_handle* h, F&& f) {
    if (!h) { g_err = "NULL handle"; return SCRIBE_E_ARG; }
    try { Guard g(h); f(); g_err.clear(); return SCRIBE_OK; }
    catch (const Fail& e) { g_err = e.msg; cudaGetLastError(); return e.code; }
    catch (const std::exception& e) { g_err = e.what(); return SCRIBE_E_ARG; }
}

}  // namespace

extern "C" {

int scribe_create(int device, scribe_handle** out) {
    if (!out) { g_err = "NULL out"; return SCRIBE_E_ARG; }
    *out = nullptr;
    try {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            fail(SCRIBE_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); scribe_b200 has no CPU fallback");
        if (device < 0) CU(cudaGetDevice(&device));
        if (device >= count) fail(SCRIBE_E_ARG, "device index out of range");
        scribe_handle* h = new scribe_handle();
        h->device = device;
        int prev = 0; cudaGetDevice(&prev);
        CU(cudaSetDevice(device));
        cudaDeviceProp p;
        CU(cudaGetDeviceProperties(&p, device));
        h->sm_count = p.multiProcessorCount;
        cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device);
        h->mem_bytes = p.totalGlobalMem;
        std::snprintf(h->name, sizeof(h->name), "%.127s", p.name);
        if (p.major != 10 || p.minor != 0) {           // the library carries sm_100a SASS only: fail here, not at the first launch
            delete h; cudaSetDevice(prev);
            fail(SCRIBE_E_CUDA, "scribe_b200 is built for sm_100a (B200) only; found sm_" + std::to_string(p.major) + std::to_string(p.minor));
        }
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&h->ev_call0)); CU(cudaEventCreate(&h->ev_call1));
        CU(cudaEventCreate(&h->ev_k0)); CU(cudaEventCreate(&h->ev_k1));
        h->ws_budget = std::min<size_t>((size_t)8 << 30, p.totalGlobalMem / 8);
        if (const char* e = std::getenv("SCRIBE_B200_NO_STRIP2")) h->no_strip2 = std::atoi(e);
        if (const char* e = std::getenv("SCRIBE_B200_SWEEP_V1")) h->sweep_v1 = std::atoi(e);
        if (const char* e = std::getenv("SCRIBE_B200_NO_FAST_RDI")) h->no_fast_rdi = std::atoi(e);
        if (const char* e = std::getenv("SCRIBE_B200_Q")) { int v = std::atoi(e); if (v == 1 || v == 2) h->force_q = v; }
        cudaSetDevice(prev);
        *out = h;
        g_err.clear();
        return SCRIBE_OK;
    } catch (const Fail& e) { g_err = e.msg; return e.code; }
}

void scribe_destroy(scribe_handle* h) {
    if (!h) return;
    Guard g(h);
    cudaStreamSynchronize(h->stream);
    for (DevBuf* b : {&h->E, &h->run_off_d, &h->S, &h->Y, &h->psi, &h->jobs, &h->partial, &h->cols, &h->specs, &h->u, &h->r,
                      &h->eps, &h->cnt, &h->wsum, &h->out, &h->flag, &h->idx32a, &h->idx32b, &h->neff, &h->dims, &h->pjobs, &h->single,
                      &h->keys_in, &h->keys_out, &h->vals_in, &h->perm, &h->sort_tmp, &h->seg_off, &h->keyspecs, &h->visited, &h->permS, &h->strips, &h->permE, &h->sortedE, &h->rawjobs,
                      &h->raw0, &h->raw1, &h->stat, &h->nanflag, &h->wcum, &h->sorted, &h->rankf, &h->crank, &h->sortedS, &h->rankS, &h->permY, &h->sortedY, &h->rankY, &h->rankE})
        b->release();
    for (DevBuf& b : h->post) b.release();
    cudaEventDestroy(h->ev_call0); cudaEventDestroy(h->ev_call1); cudaEventDestroy(h->ev_k0); cudaEventDestroy(h->ev_k1);
    cudaStreamDestroy(h->stream);
    delete h;
}

int scribe_device_info(scribe_handle* h, int32_t* sm_count, int32_t* clock_khz, int64_t* mem_bytes, char* name, int32_t name_len) {
    if (!h) { g_err = "NULL handle"; return SCRIBE_E_ARG; }
    if (sm_count) *sm_count = h->sm_count;
    if (clock_khz) *clock_khz = h->clock_khz;
    if (mem_bytes) *mem_bytes = (int64_t)h->mem_bytes;
    if (name && name_len > 0) std::snprintf(name, name_len, "%s", h->name);
    return SCRIBE_OK;
}

int scribe_device_index(scribe_handle* h) { return h ? h->device : -1; }

int scribe_get_stats(scribe_handle* h, scribe_stats* out) {
    if (!h || !out) { g_err = "NULL argument"; return SCRIBE_E_ARG; }
    *out = h->st;
    return SCRIBE_OK;
}

int scribe_estimate(scribe_handle* h, const double* x, const double* y, const double* z, int64_t n,
                    int32_t dx, int32_t dy, int32_t dz, const scribe_opts* opts, double* out) {
    return guarded(h, [&] { single_job(h, x, y, z, n, dx, dy, dz, opts, out, nullptr); });
}

int scribe_debug_counts(scribe_handle* h, const double* x, const double* y, const double* z, int64_t n,
                        int32_t dx, int32_t dy, int32_t dz, const scribe_opts* opts,
                        double* eps, int32_t* counts, double* wsum, double* weights) {
    return guarded(h, [&] {
        DebugPtrs d; d.eps = eps; d.cnt = counts; d.wsum = wsum; d.weights = weights;
        double v;
        try { single_job(h, x, y, z, n, dx, dy, dz, opts, &v, &d); }
        catch (const Fail& e) { if (e.code != SCRIBE_E_DOMAIN) throw; }   // counts were still produced
    });
}

int scribe_upload_matrix(scribe_handle* h, const double* E, int64_t n_genes, int64_t n_cells,
                         const int64_t* run_offsets, int32_t n_runs) {
    return guarded(h, [&] {
        if (!E || n_genes < 1 || n_cells < 1) fail(SCRIBE_E_ARG, "empty expression matrix");
        std::vector<int64_t> ro;
        if (run_offsets && n_runs > 0) ro.assign(run_offsets, run_offsets + n_runs + 1);
        else ro = {0, n_cells};
        if (ro.front() != 0 || ro.back() != n_cells) fail(SCRIBE_E_ARG, "run_offsets must span [0, n_cells]");
        for (size_t r = 1; r < ro.size(); r++) if (ro[r] < ro[r - 1]) fail(SCRIBE_E_ARG, "run_offsets must be non-decreasing");
        begin_call(h);
        h->E.ensure(sizeof(double) * n_genes * n_cells);
        CU(cudaMemcpyAsync(h->E.p, E, sizeof(double) * n_genes * n_cells, cudaMemcpyHostToDevice, h->stream));
        h->run_off_d.ensure(sizeof(int64_t) * ro.size());
        CU(cudaMemcpyAsync(h->run_off_d.p, ro.data(), sizeof(int64_t) * ro.size(), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->st.h2d_bytes = sizeof(double) * n_genes * n_cells + sizeof(int64_t) * ro.size();
        h->n_genes = n_genes; h->n_cells = n_cells; h->run_off = ro; h->sortE_valid = false;
        end_call(h);
    });
}

int scribe_lagged_jobs(scribe_handle* h, const scribe_job* jobs, int64_t n_jobs, const scribe_opts* opts,
                       int32_t differential, const double* scales, double* out_host) {
    return guarded(h, [&] { lagged_jobs(h, jobs, n_jobs, opts, differential, scales, out_host, false); });
}
int scribe_lagged_jobs_dev(scribe_handle* h, const scribe_job* jobs, int64_t n_jobs, const scribe_opts* opts,
                           int32_t differential, const double* scales, double* out_dev) {
    return guarded(h, [&] { lagged_jobs(h, jobs, n_jobs, opts, differential, scales, out_dev, true); });
}

int scribe_mi_jobs(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs, const scribe_opts* opts, double* out_host) {
    return guarded(h, [&] { mi_jobs(h, src, dst, n_jobs, opts, out_host); });
}

int scribe_upload_coupling(scribe_handle* h, const double* S, const double* Y, int64_t n_genes, int64_t n_cells) {
    return guarded(h, [&] {
        if (!S || !Y || n_genes < 1 || n_cells < 1) fail(SCRIBE_E_ARG, "empty coupling matrices");
        if (n_cells > 2000000) fail(SCRIBE_E_ARG, "too many cells");
        begin_call(h);
        const size_t bytes = sizeof(double) * n_genes * n_cells;
        h->S.ensure(bytes); h->Y.ensure(bytes);
        CU(cudaMemcpyAsync(h->S.p, S, bytes, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->Y.p, Y, bytes, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->st.h2d_bytes = 2 * bytes;
        h->cg = n_genes; h->cc = n_cells; h->permS_valid = false;
        end_call(h);
    });
}

// ---- consumers of the causal matrix ---------------------------------------------------------------
namespace {
// rank the n x n entries of M (host) by descending score; mode as in edge_keys_kernel.  Leaves the sorted linear indices in
// h->post[3] and returns how many of them are valid edges (finite key).
int64_t rank_edges(scribe_handle* h, const double* M_host, int64_t n, int mode)
{
    const int64_t total = n * n;
    if (total >= ((int64_t)1 << 31)) fail(SCRIBE_E_ARG, "matrix too large");
    h->post[0].ensure(sizeof(double) * total); h->post[1].ensure(sizeof(double) * total); h->post[2].ensure(sizeof(int32_t) * total);
    h->post[3].ensure(sizeof(int32_t) * total); h->post[4].ensure(sizeof(double) * total);
    CU(cudaMemcpyAsync(h->post[0].p, M_host, sizeof(double) * total, cudaMemcpyHostToDevice, h->stream));
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    edge_keys_kernel<<<grid, 256, 0, h->stream>>>(h->post[0].as<double>(), n, mode, h->post[1].as<double>(), h->post[2].as<int32_t>());
    CU(cudaGetLastError());
    size_t tmp = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->post[1].as<double>(), h->post[4].as<double>(), h->post[2].as<int32_t>(),
                                       h->post[3].as<int32_t>(), (int)total, 0, 64, h->stream));
    h->sort_tmp.ensure(tmp + 16);
    CU(cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tmp, h->post[1].as<double>(), h->post[4].as<double>(), h->post[2].as<int32_t>(),
                                       h->post[3].as<int32_t>(), (int)total, 0, 64, h->stream));
    // valid edges = keys below +inf: count them on the host from the sorted keys' tail (binary search over a D2H copy is overkill
    // for n^2 <= a few million: one pass)
    std::vector<double> keys((size_t)total);
    CU(cudaMemcpyAsync(keys.data(), h->post[4].p, sizeof(double) * total, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int64_t valid = std::lower_bound(keys.begin(), keys.end(), INFINITY) - keys.begin();
    h->st.h2d_bytes += sizeof(double) * total; h->st.d2h_bytes += sizeof(double) * total; h->st.kernel_launches += 3;
    return valid;
}
}  // namespace

int scribe_clr(scribe_handle* h, const double* M, int64_t rows, int64_t cols, int32_t both_dims, double* out) {
    return guarded(h, [&] {
        if (!M || !out || rows < 1 || cols < 1) fail(SCRIBE_E_ARG, "empty matrix");
        begin_call(h);
        const int64_t total = rows * cols;
        h->post[0].ensure(sizeof(double) * total); h->post[1].ensure(sizeof(double) * total);
        h->post[2].ensure(sizeof(double) * 2 * rows); h->post[3].ensure(sizeof(double) * 2 * cols);
        CU(cudaMemcpyAsync(h->post[0].p, M, sizeof(double) * total, cudaMemcpyHostToDevice, h->stream));
        double* rs = h->post[2].as<double>(); double* cs = h->post[3].as<double>();
        line_mean_std_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, h->stream>>>(h->post[0].as<double>(), rows, cols, 1, rs, rs + rows);
        CU(cudaGetLastError());
        if (both_dims) {
            line_mean_std_kernel<<<(unsigned)((cols * 32 + 255) / 256), 256, 0, h->stream>>>(h->post[0].as<double>(), rows, cols, 0, cs, cs + cols);
            CU(cudaGetLastError());
        }
        clr_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, h->stream>>>(
            h->post[0].as<double>(), rows, cols, rs, rs + rows, cs, cs + cols, both_dims ? 1 : 0, rows == cols ? 1 : 0, h->post[1].as<double>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, h->post[1].p, sizeof(double) * total, cudaMemcpyDeviceToHost, h->stream));
        h->st.h2d_bytes = sizeof(double) * total; h->st.d2h_bytes = sizeof(double) * total; h->st.kernel_launches = both_dims ? 3 : 2;
        end_call(h);
    });
}

int scribe_top_edges(scribe_handle* h, const double* M, int64_t n, int64_t n_top, int32_t* src, int32_t* dst, double* weight, int64_t* n_found) {
    return guarded(h, [&] {
        if (!M || n < 1 || n_top < 0 || !n_found || (n_top > 0 && (!src || !dst || !weight))) fail(SCRIBE_E_ARG, "bad top-edges request");
        begin_call(h);
        const int64_t valid = rank_edges(h, M, n, 0);
        const int64_t take = std::min(valid, n_top);
        std::vector<int32_t> idx((size_t)take);
        if (take > 0) CU(cudaMemcpyAsync(idx.data(), h->post[3].p, sizeof(int32_t) * take, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        for (int64_t i = 0; i < take; i++) { src[i] = idx[i] / (int32_t)n; dst[i] = idx[i] % (int32_t)n; weight[i] = M[idx[i]]; }
        *n_found = take;
        end_call(h);
    });
}

int scribe_roc(scribe_handle* h, const double* scores, const uint8_t* truth, int64_t n, double* auroc, double* fpr, double* tpr) {
    return guarded(h, [&] {
        if (!scores || !truth || n < 2 || !auroc) fail(SCRIBE_E_ARG, "bad roc request");
        begin_call(h);
        const int64_t m = rank_edges(h, scores, n, 1);                // ranked off-diagonal edges (NaN scores excluded)
        if (m < 1) fail(SCRIBE_E_ARG, "no scored edges");
        h->post[5].ensure((size_t)n * n); h->post[6].ensure(sizeof(int32_t) * 2 * (m + 1)); h->post[7].ensure(sizeof(double) * (3 * (m + 1) + 1));
        CU(cudaMemcpyAsync(h->post[5].p, truth, (size_t)n * n, cudaMemcpyHostToDevice, h->stream));
        int32_t* tp = h->post[6].as<int32_t>(); int32_t* fp = tp + (m + 1);
        double* x = h->post[7].as<double>(); double* y = x + (m + 1); double* terms = y + (m + 1); double* area = terms + m;
        roc_scan_kernel<<<1, 1024, 0, h->stream>>>(h->post[3].as<int32_t>(), h->post[5].as<unsigned char>(), m, tp, fp);
        CU(cudaGetLastError());
        int32_t tot[2];
        CU(cudaMemcpyAsync(&tot[0], tp + m, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(&tot[1], fp + m, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (tot[0] == 0 || tot[1] == 0) fail(SCRIBE_E_DOMAIN, "float division by zero");     // max(true_positive) or max(false_positive) == 0 (:375-376)
        roc_curve_kernel<<<(unsigned)std::min<int64_t>((m + 256) / 256, 148 * 8), 256, 0, h->stream>>>(tp, fp, m, x, y, terms);
        CU(cudaGetLastError());
        sequential_sum_kernel<<<1, 32, 0, h->stream>>>(terms, m, area);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(auroc, area, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (fpr) CU(cudaMemcpyAsync(fpr, x, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, h->stream));
        if (tpr) CU(cudaMemcpyAsync(tpr, y, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, h->stream));
        h->st.kernel_launches += 3; h->st.d2h_bytes += sizeof(double) * (2 * (m + 1) + 1);
        end_call(h);
    });
}

// ---- preprocessing of the resident expression matrix -------------------------------------------------
int scribe_normalize_resident(scribe_handle* h, int32_t mean_order, int32_t avg_order) {
    return guarded(h, [&] {
        if (h->n_genes == 0) fail(SCRIBE_E_STATE, "no expression matrix uploaded");
        if ((mean_order | avg_order) & ~1) fail(SCRIBE_E_ARG, "summation order must be 0 (pairwise) or 1 (sequential)");
        begin_call(h);
        const int n_runs = (int)h->run_off.size() - 1;
        h->raw0.ensure(sizeof(double) * h->n_genes * h->n_cells);
        const int64_t segs = h->n_genes * n_runs;
        normalize_rows_kernel<<<(unsigned)((segs + 63) / 64), 64, 0, h->stream>>>(h->E.as<double>(), h->n_genes, h->n_cells, h->run_off_d.as<int64_t>(),
                                                                                 n_runs, mean_order, avg_order, h->raw0.as<double>());
        CU(cudaGetLastError());
        h->st.kernel_launches = 1; h->sortE_valid = false;
        end_call(h);
    });
}

int scribe_smooth_resident(scribe_handle* h, int32_t window, int64_t* new_n_cells, int64_t* new_run_offsets) {
    return guarded(h, [&] {
        if (h->n_genes == 0) fail(SCRIBE_E_STATE, "no expression matrix uploaded");
        if (window < 1) fail(SCRIBE_E_ARG, "window must be >= 1");
        const int n_runs = (int)h->run_off.size() - 1;
        std::vector<int64_t> off(n_runs + 1, 0);
        for (int r = 0; r < n_runs; r++) off[r + 1] = off[r] + std::max<int64_t>(0, h->run_off[r + 1] - h->run_off[r] - (window - 1));
        const int64_t out_cells = off[n_runs];
        if (out_cells < 1) fail(SCRIBE_E_LENGTH, "window leaves no cells");
        begin_call(h);
        h->raw0.ensure(sizeof(double) * h->n_genes * out_cells); h->raw1.ensure(sizeof(int64_t) * (n_runs + 1));
        CU(cudaMemcpyAsync(h->raw1.p, off.data(), sizeof(int64_t) * (n_runs + 1), cudaMemcpyHostToDevice, h->stream));
        const int64_t segs = h->n_genes * n_runs;
        smooth_rows_kernel<<<(unsigned)((segs + 63) / 64), 64, 0, h->stream>>>(h->E.as<double>(), h->n_genes, h->n_cells, h->run_off_d.as<int64_t>(),
                                                                              n_runs, window, h->raw0.as<double>(), out_cells, h->raw1.as<int64_t>());
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
        std::swap(h->E, h->raw0);                                  // the smoothed matrix becomes the resident one
        CU(cudaMemcpyAsync(h->run_off_d.p, off.data(), sizeof(int64_t) * (n_runs + 1), cudaMemcpyHostToDevice, h->stream));
        h->n_cells = out_cells; h->run_off = off; h->sortE_valid = false;
        if (new_n_cells) *new_n_cells = out_cells;
        if (new_run_offsets) std::memcpy(new_run_offsets, off.data(), sizeof(int64_t) * (n_runs + 1));
        h->st.kernel_launches = 1; h->st.h2d_bytes = 2 * sizeof(int64_t) * (n_runs + 1);
        end_call(h);
    });
}

int scribe_read_matrix(scribe_handle* h, int64_t gene0, int64_t n_genes, double* out) {
    return guarded(h, [&] {
        if (h->n_genes == 0) fail(SCRIBE_E_STATE, "no expression matrix uploaded");
        if (gene0 < 0 || n_genes < 1 || gene0 + n_genes > h->n_genes || !out) fail(SCRIBE_E_ARG, "bad row range");
        begin_call(h);
        CU(cudaMemcpyAsync(out, h->E.as<double>() + (size_t)gene0 * h->n_cells, sizeof(double) * n_genes * h->n_cells, cudaMemcpyDeviceToHost, h->stream));
        h->st.d2h_bytes = sizeof(double) * n_genes * h->n_cells;
        end_call(h);
    });
}

int scribe_knn_regress(scribe_handle* h, const double* px, const double* py, const double* value, int64_t n,
                       const double* qx, const double* qy, int64_t m, int32_t k, double* out) {
    return guarded(h, [&] {
        if (!px || !py || !value || !qx || !qy || !out || n < 1 || m < 1) fail(SCRIBE_E_ARG, "empty kNN regression");
        if (k < 1 || k + 1 > KNN_REG_MAXK) fail(SCRIBE_E_ARG, "k must be in [1, 15]");
        begin_call(h);
        h->post[0].ensure(sizeof(double) * 3 * n); h->post[1].ensure(sizeof(double) * 3 * m);
        double* d = h->post[0].as<double>(); double* qd = h->post[1].as<double>();
        CU(cudaMemcpyAsync(d, px, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(d + n, py, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(d + 2 * n, value, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(qd, qx, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(qd + m, qy, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
        knn_regress_kernel<<<(unsigned)((m * 32 + 127) / 128), 128, 0, h->stream>>>(d, d + n, d + 2 * n, n, qd, qd + m, m, k, qd + 2 * m);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, qd + 2 * m, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
        h->st.h2d_bytes = sizeof(double) * (3 * n + 2 * m); h->st.d2h_bytes = sizeof(double) * m; h->st.kernel_launches = 1;
        end_call(h);
    });
}

int scribe_upload_coupling_raw(scribe_handle* h, const double* t0_layer, const double* t1_layer, int64_t n_cells, int64_t n_genes,
                               int32_t normalize, int32_t y_is_sum, int32_t* t0_has_nan) {
    return guarded(h, [&] {
        if (!t0_layer || !t1_layer || n_genes < 1 || n_cells < 1) fail(SCRIBE_E_ARG, "empty coupling layers");
        if (n_cells > 2000000) fail(SCRIBE_E_ARG, "too many cells");
        begin_call(h);
        const size_t bytes = sizeof(double) * n_genes * n_cells;
        h->raw0.ensure(bytes); h->raw1.ensure(bytes); h->S.ensure(bytes); h->Y.ensure(bytes);
        h->stat.ensure(sizeof(double) * 4 * n_genes); h->nanflag.ensure(2 * sizeof(int));
        CU(cudaMemsetAsync(h->nanflag.p, 0, 2 * sizeof(int), h->stream));
        CU(cudaMemcpyAsync(h->raw0.p, t0_layer, bytes, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->raw1.p, t1_layer, bytes, cudaMemcpyHostToDevice, h->stream));
        // cells x genes -> genes x cells; NaN of the t1 layer -> 0 (Scribe.py:101); a NaN in the t0 layer is reported, not handled
        dim3 grid((unsigned)((n_genes + 31) / 32), (unsigned)((n_cells + 31) / 32));
        transpose_kernel<<<grid, 256, 0, h->stream>>>(h->raw0.as<double>(), n_cells, n_genes, h->S.as<double>(), 0, h->nanflag.as<int>());
        CU(cudaGetLastError());
        transpose_kernel<<<grid, 256, 0, h->stream>>>(h->raw1.as<double>(), n_cells, n_genes, h->Y.as<double>(), 1, h->nanflag.as<int>() + 1);
        CU(cudaGetLastError());
        double* st = h->stat.as<double>();
        if (normalize) {
            row_mean_range_kernel<<<(unsigned)((n_genes + 63) / 64), 64, 0, h->stream>>>(h->S.as<double>(), n_genes, n_cells, st, st + n_genes);
            CU(cudaGetLastError());
            row_mean_range_kernel<<<(unsigned)((n_genes + 63) / 64), 64, 0, h->stream>>>(h->Y.as<double>(), n_genes, n_cells, st + 2 * n_genes, st + 3 * n_genes);
            CU(cudaGetLastError());
        }
        coupling_normalise_kernel<<<(unsigned)std::min<int64_t>((n_genes * n_cells + 255) / 256, 148 * 16), 256, 0, h->stream>>>(
            h->S.as<double>(), h->Y.as<double>(), n_genes, n_cells, st, st + n_genes, st + 2 * n_genes, st + 3 * n_genes, normalize ? 1 : 0, y_is_sum ? 1 : 0);
        CU(cudaGetLastError());
        int flags[2] = {0, 0};
        CU(cudaMemcpyAsync(flags, h->nanflag.p, sizeof(flags), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (t0_has_nan) *t0_has_nan = flags[0];
        h->st.h2d_bytes = 2 * bytes; h->st.d2h_bytes = sizeof(flags); h->st.kernel_launches = normalize ? 5 : 3;
        h->cg = n_genes; h->cc = n_cells; h->permS_valid = false;
        end_call(h);
    });
}

int scribe_read_coupling_row(scribe_handle* h, int32_t which, int64_t gene, double* out_row) {
    return guarded(h, [&] {
        if (h->cg == 0) fail(SCRIBE_E_STATE, "no coupling matrices uploaded");
        if (gene < 0 || gene >= h->cg || !out_row || (which != 0 && which != 1)) fail(SCRIBE_E_ARG, "bad row request");
        const double* src = (which == 0 ? h->S.as<double>() : h->Y.as<double>()) + (size_t)gene * h->cc;
        CU(cudaMemcpyAsync(out_row, src, sizeof(double) * h->cc, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    });
}

int scribe_coupling_jobs(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs, int32_t drop_zero_cells,
                         const scribe_opts* opts, double* out_host, int32_t* n_eff) {
    return guarded(h, [&] { coupling_jobs(h, src, dst, n_jobs, drop_zero_cells, opts, out_host, false, n_eff); });
}
int scribe_coupling_jobs_dev(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs, int32_t drop_zero_cells,
                             const scribe_opts* opts, double* out_dev, int32_t* n_eff) {
    return guarded(h, [&] { coupling_jobs(h, src, dst, n_jobs, drop_zero_cells, opts, out_dev, true, n_eff); });
}

}  // extern "C"
