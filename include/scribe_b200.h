/*
 * scribe_b200.h -- C-ABI of the B200-native replacement for Scribe-py's pairwise-causality hot path.
 *
 * The reference (aristoteleo/Scribe-py) is pure Python and has no FFI; its seams for this path are plain
 * Python callables.  Each entry point below names the reference callable(s) it stands in for
 * (paths relative to the reference tree).  Python binds these with ctypes (scribe_py_b200/_lib.py);
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; scribe_last_error() gives the message
 *     (thread-local).  Python maps codes onto the reference's exception types (see SCRIBE_E_*).
 *   - "host" pointers are ordinary CPU memory owned by the caller; "dev" pointers are CUDA device memory
 *     on the handle's device, owned by the caller.  The library keeps no caller pointer after return.
 *   - point sets are passed dimension-major ("SoA"): x is dx rows of n contiguous doubles.
 *   - calls on one handle are not re-entrant; use one handle per GPU / per thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with SCRIBE_E_CUDA.
 */
#ifndef SCRIBE_B200_H
#define SCRIBE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCRIBE_ABI_VERSION 1
#define SCRIBE_MAX_DIM   16   /* joint dimension dx+dy+dz accepted by the estimators          */
#define SCRIBE_MAX_COND   4   /* conditioning genes per cRDI job (reference: L, causal_network.py:190) */
#define SCRIBE_MAX_K     63   /* k and k_density upper bound                                    */

/* status codes */
#define SCRIBE_OK          0
#define SCRIBE_E_ARG       1  /* bad argument (-> ValueError)                                   */
#define SCRIBE_E_LENGTH    2  /* "Lists should have same length" (-> AssertionError)            */
#define SCRIBE_E_CUDA      3  /* CUDA runtime failure / no device (-> RuntimeError)             */
#define SCRIBE_E_DOMAIN    4  /* umi hit log(0): reference raises ValueError("math domain error")*/
#define SCRIBE_E_STATE     5  /* e.g. jobs submitted before a matrix was uploaded                */

/* estimators: Scribe/information_estimators.py mi :59, cmi :111, umi :209, cumi :326 */
#define SCRIBE_EST_MI    0
#define SCRIBE_EST_CMI   1
#define SCRIBE_EST_UMI   2
#define SCRIBE_EST_CUMI  3

/* density_estimation_method of umi/cumi (information_estimators.py:247-259, :384-394) */
#define SCRIBE_DENSITY_KDE 0
#define SCRIBE_DENSITY_KNN 1

/* kernel path selection (testing / debugging) */
#define SCRIBE_PATH_AUTO    0  /* windowed sorted-sweep kernel for mi/cmi/cumi (scalar x and y, <= 3 conditioning
                                  columns, k <= 15) when N >= 256, else the dense fused kernels, else the generic
                                  kernels                                                         */
#define SCRIBE_PATH_GENERIC 1  /* runtime-dimension kernels (any dx,dy,dz,k within the limits)     */
#define SCRIBE_PATH_DENSE   2  /* dense fused kernels: every query visits every candidate          */
#define SCRIBE_PATH_SWEEP   3  /* sorted-sweep kernels regardless of N (environment SCRIBE_B200_SWEEP_V1=1 selects
                                  the first, all-FP64 form for cmi/cumi)                          */
#define SCRIBE_PATH_STRIP   4  /* strip kernels (x-sorted z-strips + z-sweep) regardless of N       */

typedef struct scribe_handle scribe_handle;

typedef struct scribe_opts {
    int32_t estimator;   /* SCRIBE_EST_*                                                        */
    int32_t k;           /* neighbours (reference default 5; the drivers never override it)     */
    int32_t density;     /* SCRIBE_DENSITY_* (umi / cumi only)                                  */
    int32_t k_density;   /* k of the kNN density estimate                                       */
    int32_t path;        /* SCRIBE_PATH_*                                                       */
    int32_t reserved;
    double  bw;          /* KDE bandwidth in data units (reference default 0.01)                */
} scribe_opts;

/* One lagged-embedding job = one call of cmi/cumi made by causal_model.rdi (causal_network.py:145-176)
 * or causal_model.crdi (causal_network.py:257-283).
 *   n_cond == 0 : RDI(src -> dst, delay):  x = e_src[t], z = e_dst[t+delay-1], y = e_dst[t+delay]
 *   n_cond == L : cRDI: tau = max(delay, cond_delay[]); x = e_src[t+tau-delay], y = e_dst[t+tau],
 *                 z = [ e_cond[L-1][t+tau-cond_delay[L-1]], ..., e_cond[0][..], e_dst[t+tau-1] ]
 * per run, runs concatenated in upload order. */
typedef struct scribe_job {
    int32_t src, dst, delay, n_cond;
    int32_t cond_gene[SCRIBE_MAX_COND];
    int32_t cond_delay[SCRIBE_MAX_COND];
} scribe_job;

typedef struct scribe_stats {
    int64_t kernel_launches;   /* kernels launched by the last compute call                     */
    int64_t estimator_launches;/* ... of which the O(N^2) estimator kernels                      */
    double  estimator_ms;      /* device time of those estimator kernels (CUDA events, own stream)*/
    double  total_ms;          /* device time of the whole call incl. copies                     */
    int64_t point_pairs;       /* sum over jobs of N^2 (dense point pairs the estimator visited) */
    int64_t jobs;              /* jobs evaluated                                                 */
    int64_t h2d_bytes, d2h_bytes;
    /* point-pair visits actually made, per pass: N^2 each for the dense kernels, less for the sorted sweep */
    int64_t visited_knn;       /* pass 1: k-th neighbour search                                   */
    int64_t visited_count;     /* pass 2: ball counts                                             */
    int64_t visited_kde;       /* KDE weights (cumi / umi with density = KDE)                     */
} scribe_stats;

/* ---- lifetime -------------------------------------------------------------------------------- */
int  scribe_abi_version(void);
const char* scribe_last_error(void);
/* device < 0 : current device.  Fails with SCRIBE_E_CUDA when no CUDA device is usable. */
int  scribe_create(int device, scribe_handle** out);
void scribe_destroy(scribe_handle* h);
int  scribe_device_info(scribe_handle* h, int32_t* sm_count, int32_t* clock_khz, int64_t* mem_bytes,
                        char* name, int32_t name_len);
int  scribe_get_stats(scribe_handle* h, scribe_stats* out);
/* CUDA device ordinal the handle computes on (the process group's collectives must use the same device). */
int  scribe_device_index(scribe_handle* h);

/* ---- estimator level: replaces a single call of information_estimators.mi/cmi/umi/cumi -------- */
/* x: dx*n, y: dy*n, z: dz*n doubles (z == NULL and dz == 0 for mi/umi), already normalised by the caller
 * when the reference's normalization=True is wanted (x/np.std(x), information_estimators.py:150-153). */
int scribe_estimate(scribe_handle* h, const double* x, const double* y, const double* z,
                    int64_t n, int32_t dx, int32_t dy, int32_t dz, const scribe_opts* opts, double* out);

/* Debug / bit-exactness entry point: the per-point quantities behind scribe_estimate.
 *   eps[n]        k-th neighbour distance (information_estimators.py:101,:165,:263,:396)
 *   counts[4*n]   ball sizes minus self: cmi/cumi {xyz,xz,yz,z}; mi {xy,x,y,(n-1)}; umi {-,x,y,-}
 *   wsum[2*n]     cumi {W_yz - w_i, W_z - w_i}; umi {W_y - w_i, unused}   (may be NULL)
 *   weights[n]    uniformisation weights w_i (umi/cumi)                   (may be NULL)      */
int scribe_debug_counts(scribe_handle* h, const double* x, const double* y, const double* z,
                        int64_t n, int32_t dx, int32_t dy, int32_t dz, const scribe_opts* opts,
                        double* eps, int32_t* counts, double* wsum, double* weights);

/* ---- driver level: causal_model.rdi / .crdi (causal_network.py:112-289) ----------------------- */
/* E: n_genes rows of n_cells doubles (gene-major), NaN padding already removed; run r occupies
 * columns [run_offsets[r], run_offsets[r+1]) of every row.  Replaces model.expression for the loops
 * at causal_network.py:140-176 / :223-283.  The matrix stays resident until the next upload. */
int scribe_upload_matrix(scribe_handle* h, const double* E, int64_t n_genes, int64_t n_cells,
                         const int64_t* run_offsets, int32_t n_runs);
/* causal_model.normalize (causal_network.py:104-109) in place on the resident matrix: every (gene, run) row -> (v - mean) /
 * std(ddof = 1), NaN -> 0.  mean_order / avg_order select the summation order (0 numpy pairwise, 1 sequential) of the row sum
 * behind the mean and of the one inside the variance, so that the caller can reproduce its pandas build bit for bit. */
int scribe_normalize_resident(scribe_handle* h, int32_t mean_order, int32_t avg_order);
/* causal_model.smooth (causal_network.py:97-101): the resident matrix is replaced by its rolling mean over `window` cells
 * along every (gene, run) row (pandas' Kahan-compensated online mean); every run shrinks by window - 1 cells.
 * new_n_cells / new_run_offsets (n_runs + 1 entries; either may be NULL) receive the new layout. */
int scribe_smooth_resident(scribe_handle* h, int32_t window, int64_t* new_n_cells, int64_t* new_run_offsets);
/* Rows [gene0, gene0 + n_genes) of the resident matrix back to the host (n_genes x n_cells doubles). */
int scribe_read_matrix(scribe_handle* h, int64_t gene0, int64_t n_genes, double* out);
/* Evaluate jobs against the resident matrix.  opts->estimator is CMI (uniformization=False) or CUMI.
 * differential != 0: y = e_dst[p-1] - e_dst[p] (causal_network.py:153,:173).
 * scales: NULL, or 3 doubles per job {sx, sy, sz}; each role is divided by its scale (the caller computes
 * np.std for normalization=True).  out: n_jobs doubles, host (…_jobs) or device (…_jobs_dev). */
int scribe_lagged_jobs(scribe_handle* h, const scribe_job* jobs, int64_t n_jobs, const scribe_opts* opts,
                       int32_t differential, const double* scales, double* out_host);
int scribe_lagged_jobs_dev(scribe_handle* h, const scribe_job* jobs, int64_t n_jobs, const scribe_opts* opts,
                           int32_t differential, const double* scales, double* out_dev);

/* Pairwise mutual-information network over the resident matrix: job j = mi(E[src[j]], E[dst[j]]) on all cells, no lag
 * (other_estimators.mi, Scribe/other_estimators.py:50-83 -- dead code in the reference because its helper is shadowed,
 * :10-11 vs :50).  opts->estimator must be SCRIBE_EST_MI. */
int scribe_mi_jobs(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs,
                   const scribe_opts* opts, double* out_host);

/* ---- driver level: causal_net_dynamics_coupling (Scribe.py:110-135) --------------------------- */
/* S, Y: n_genes rows of n_cells doubles.  S = normalised t0 layer (x and z columns, Scribe.py:119,:123);
 * Y = the y column per gene (S + V when t1_key == 'velocity', else V; Scribe.py:121). */
int scribe_upload_coupling(scribe_handle* h, const double* S, const double* Y, int64_t n_genes, int64_t n_cells);
/* The same from the RAW layers, normalised on the device: t0_layer / t1_layer are the AnnData layers as they are stored,
 * n_cells rows of n_genes doubles (dense, cell-major).  NaN of the t1 layer -> 0 (Scribe.py:101); normalize != 0 applies
 * (layer - mean) / (max - min) per gene (Scribe.py:104-106) with numpy's summation order for the mean; Y = S + V when
 * y_is_sum (t1_key == 'velocity', Scribe.py:121) else V.  *t0_has_nan != 0 on return means the t0 layer holds NaN, which
 * pandas would skip in the mean: the caller must then normalise on the host and use scribe_upload_coupling. */
int scribe_upload_coupling_raw(scribe_handle* h, const double* t0_layer, const double* t1_layer, int64_t n_cells, int64_t n_genes,
                               int32_t normalize, int32_t y_is_sum, int32_t* t0_has_nan);
/* One row of the resident S (which == 0) or Y (which == 1) matrix back to the host (n_cells doubles): lets the caller check
 * the device normalisation against the reference's pandas expression on a few genes. */
int scribe_read_coupling_row(scribe_handle* h, int32_t which, int64_t gene, double* out_row);
/* job j: I(S[src[j]] ; Y[dst[j]] | S[dst[j]]); drop_zero_cells keeps cells with (x + y) + z > 0
 * (Scribe.py:125-128).  n_eff (may be NULL) receives the cells kept per job. */
int scribe_coupling_jobs(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs,
                         int32_t drop_zero_cells, const scribe_opts* opts, double* out_host, int32_t* n_eff);
int scribe_coupling_jobs_dev(scribe_handle* h, const int32_t* src, const int32_t* dst, int64_t n_jobs,
                             int32_t drop_zero_cells, const scribe_opts* opts, double* out_dev, int32_t* n_eff);

/* ---- consumers of the causal matrix --------------------------------------------------- */
/* Context likelihood of relatedness (Scribe.CLR, Scribe/Scribe.py:142-176) of a rows x cols score matrix (row-major, host):
 * z-scores of round(M, 10) against the row (both_dims: and column) mean / std (ddof = 1) of M, negatives clipped to 0,
 * sqrt((z_row^2 + z_col^2) / 2).  A square matrix is z-scored against the statistics of ROW j, as the reference's broadcast
 * does (Scribe.py:151-154). */
int scribe_clr(scribe_handle* h, const double* M, int64_t rows, int64_t cols, int32_t both_dims, double* out);
/* The n_top strongest edges of a square n x n score matrix, strongest first (vis_causal_net, Scribe/plot/networks.py:34-47):
 * an entry smaller than its transpose (the weaker direction of a reciprocal pair) and NaN entries are dropped; ties keep
 * row-major order.  src/dst/weight receive *n_found <= n_top edges. */
int scribe_top_edges(scribe_handle* h, const double* M, int64_t n, int64_t n_top, int32_t* src, int32_t* dst, double* weight,
                     int64_t* n_found);
/* ROC curve and AUROC of an n x n score matrix against a truth mask (causal_model.roc, Scribe/causal_network.py:337-389):
 * off-diagonal edges, source-major, ranked by descending score (stable); truth[i*n+j] != 0 marks a true edge.  fpr/tpr
 * (may be NULL) receive n*(n-1)+1 points when no score is NaN.  SCRIBE_E_DOMAIN when there is no true or no false edge
 * (the reference divides by zero there). */
int scribe_roc(scribe_handle* h, const double* scores, const uint8_t* truth, int64_t n, double* auroc, double* fpr, double* tpr);

/* kNN grid regressor of viz_causality / viz_comb_logic (Scribe/plot/heatmaps.py:405-423, :578-596): for each of the m query
 * points (qx, qy) the k+1 Euclidean-nearest of the n data points (px, py) -- tree.query(q, k + 1) -- the nearest dropped,
 * weights exp(-d / min d) normalised to 1, out[q] = sum w * value[neighbour].  1 <= k <= 15. */
int scribe_knn_regress(scribe_handle* h, const double* px, const double* py, const double* value, int64_t n,
                       const double* qx, const double* qy, int64_t m, int32_t k, double* out);

#ifdef __cplusplus
}
#endif
#endif /* SCRIBE_B200_H */
