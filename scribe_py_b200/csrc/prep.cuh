// prep.cuh -- the callers either side of the estimator kernels, on the device: layer normalisation of
// causal_net_dynamics_coupling (Scribe.py:88-121), z-score / rolling mean of causal_model.normalize / .smooth
// (causal_network.py:97-109), CLR (Scribe.py:142-176), the edge ranking behind roc (causal_network.py:337-389) and the
// top-n edge extraction of vis_causal_net (plot/networks.py:34-47).  All O(G * cells) or O(G^2): HBM-bound, one pass each.
//
// Sums that feed a value the estimators later compare bit-for-bit (the normalised layers, the z-scored matrix) follow the
// summation ORDER of the numpy/pandas expression they replace, so the device result is the same float64 as the reference's.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace scribe {

// numpy's pairwise summation of n doubles, `stride` elements apart (numpy/core/src/umath/loops_utils.h.src,
// DOUBLE_pairwise_sum): blocks of <= 128 elements summed with 8 interleaved accumulators, blocks combined by a binary
// split at floor(n/2) rounded down to a multiple of 8.  np.sum / np.mean over a contiguous axis use exactly this order.
__device__ inline double np_pairwise_leaf(const double* a, int64_t n, int64_t stride)
{
    if (n < 8) {
        double res = 0.;
        for (int64_t i = 0; i < n; i++) res += a[i * stride];
        return res;
    }
    double r[8];
    #pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j * stride];
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
        #pragma unroll
        for (int j = 0; j < 8; j++) r[j] += a[(i + j) * stride];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i * stride];
    return res;
}

__device__ inline double np_pairwise_sum(const double* a, int64_t n, int64_t stride = 1)
{
    struct Frame { int64_t off, n; double left; int stage; };
    Frame st[48];
    int sp = 0;
    st[0] = Frame{0, n, 0.0, 0};
    double ret = 0.0;
    while (sp >= 0) {
        Frame& f = st[sp];
        if (f.n <= 128) { ret = np_pairwise_leaf(a + f.off * stride, f.n, stride); --sp; continue; }
        int64_t n2 = f.n / 2; n2 -= n2 % 8;
        if (f.stage == 0)      { f.stage = 1; st[sp + 1] = Frame{f.off, n2, 0.0, 0}; ++sp; }
        else if (f.stage == 1) { f.left = ret; f.stage = 2; st[sp + 1] = Frame{f.off + n2, f.n - n2, 0.0, 0}; ++sp; }
        else                   { ret = f.left + ret; --sp; }
    }
    return ret;
}

// out[c][r] = in[r][c] (in: rows x cols, row-major), optionally NaN -> 0; *nan_flag is set when a NaN was seen
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ in, int64_t rows, int64_t cols, double* __restrict__ out, int nan_to_zero, int* __restrict__ nan_flag)
{
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8 threads
    bool saw = false;
    for (int k = ty; k < 32; k += 8) {
        const int64_t r = r0 + k, c = c0 + tx;
        double v = 0.0;
        if (r < rows && c < cols) {
            v = in[r * cols + c];
            if (v != v) { saw = true; if (nan_to_zero) v = 0.0; }
        }
        tile[k][tx] = v;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int64_t c = c0 + k, r = r0 + tx;
        if (r < rows && c < cols) out[c * rows + r] = tile[tx][k];
    }
    if (saw && nan_flag) *nan_flag = 1;
}

// per row of M (n_rows x n_cols, row-major, no NaN): mean = np.sum(row) / n_cols in numpy's order, range = max - min
__global__ void row_mean_range_kernel(const double* __restrict__ M, int64_t n_rows, int64_t n_cols, double* __restrict__ mean, double* __restrict__ range)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rows) return;
    const double* a = M + g * n_cols;
    mean[g] = np_pairwise_sum(a, n_cols) / (double)n_cols;
    double mx = a[0], mn = a[0];
    for (int64_t i = 1; i < n_cols; i++) { const double v = a[i]; mx = (v > mx) ? v : mx; mn = (v < mn) ? v : mn; }
    range[g] = mx - mn;
}

// Scribe.py:104-106,:121 per element: s = (s - mean_s) / range_s, v = (v - mean_v) / range_v, y = s + v (velocity) or v.
// S, V gene-major; V is overwritten with y.  normalize == 0 leaves s and v as they are.
__global__ void __launch_bounds__(256)
coupling_normalise_kernel(double* __restrict__ S, double* __restrict__ V, int64_t n_genes, int64_t n_cells, const double* __restrict__ mean_s,
                          const double* __restrict__ range_s, const double* __restrict__ mean_v, const double* __restrict__ range_v,
                          int normalize, int y_is_sum)
{
    const int64_t total = n_genes * n_cells;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = i / n_cells;
        double s = S[i], v = V[i];
        if (normalize) { s = (s - mean_s[g]) / range_s[g]; v = (v - mean_v[g]) / range_v[g]; }
        S[i] = s;
        V[i] = y_is_sum ? s + v : v;
    }
}

// ------------------------------------------------------------------------------------------------
// causal_model.normalize (causal_network.py:104-109) on the resident matrix: every (gene, run) row becomes
// (v - mean) / std(ddof = 1), NaN -> 0.  pandas forms the two statistics with different summation orders depending on its
// version and the frame's memory layout, so both orders are implemented (0 = numpy pairwise, 1 = sequential) and the host
// picks the pair that reproduces pandas on a few probe rows (and falls back to pandas itself when none does).
//   c   = v - sum_{order m}(v) / n                                  (first statement of the reference: expression.sub(mean))
//   avg = sum_{order a}(c) / n ;  std = sqrt( sum_{pairwise}((avg - c)^2) / (n - 1) )   (second statement: the std is taken of the
//   z   = c / std                                                    ALREADY CENTRED frame; pandas nanops.nanvar)
// One thread per (gene, run) segment: the sums are sequential by definition.
// ------------------------------------------------------------------------------------------------
__device__ inline double ordered_sum(const double* a, int64_t n, int order)
{
    if (order == 0) return np_pairwise_sum(a, n);
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += a[i];
    return s;
}

__global__ void normalize_rows_kernel(double* __restrict__ E, int64_t n_genes, int64_t n_cells, const int64_t* __restrict__ run_off, int n_runs,
                                      int mean_order, int avg_order, double* __restrict__ scratch)
{
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_genes * n_runs) return;
    const int64_t g = seg / n_runs; const int r = (int)(seg - g * n_runs);
    const int64_t n = run_off[r + 1] - run_off[r];
    if (n <= 0) return;
    double* v = E + g * n_cells + run_off[r];
    double* sq = scratch + g * n_cells + run_off[r];
    const double mean = ordered_sum(v, n, mean_order) / (double)n;
    for (int64_t i = 0; i < n; i++) v[i] = v[i] - mean;
    const double avg = ordered_sum(v, n, avg_order) / (double)n;
    for (int64_t i = 0; i < n; i++) { const double d = avg - v[i]; sq[i] = d * d; }
    const double sd = sqrt(np_pairwise_sum(sq, n) / (double)(n - 1));
    for (int64_t i = 0; i < n; i++) { const double z = v[i] / sd; v[i] = (z != z) ? 0.0 : z; }
}

// causal_model.smooth (causal_network.py:97-101): rolling mean of `w` cells along every (gene, run) row, pandas' online
// algorithm (pandas/_libs/window/aggregations.pyx roll_mean): Kahan-compensated running sum with separate compensation terms
// for the values entering and leaving the window, leaving before entering; a window of identical values returns that value;
// a window without negatives never returns a negative mean (and vice versa).  The first w - 1 positions of a row have no
// full window: dropna(axis=1, how="all") removes those columns, so every run shrinks by w - 1 cells in the output layout.
__global__ void smooth_rows_kernel(const double* __restrict__ E, int64_t n_genes, int64_t n_cells, const int64_t* __restrict__ run_off, int n_runs,
                                   int w, double* __restrict__ out, int64_t out_cells, const int64_t* __restrict__ out_off)
{
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_genes * n_runs) return;
    const int64_t g = seg / n_runs; const int r = (int)(seg - g * n_runs);
    const int64_t n = run_off[r + 1] - run_off[r];
    if (n < w) return;
    const double* v = E + g * n_cells + run_off[r];
    double* o = out + g * out_cells + out_off[r];
    double sum_x = 0.0, c_add = 0.0, c_rem = 0.0, prev = v[0];
    int64_t nobs = 0, neg = 0, same = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i >= w) {                                            // remove_mean(values[i - w])
            const double val = v[i - w];
            const double y = -val - c_rem, t = sum_x + y;
            c_rem = t - sum_x - y; sum_x = t;
            nobs--; if (signbit(val)) neg--;
        }
        {                                                        // add_mean(values[i])
            const double val = v[i];
            const double y = val - c_add, t = sum_x + y;
            c_add = t - sum_x - y; sum_x = t;
            nobs++; if (signbit(val)) neg++;
            same = (val == prev) ? same + 1 : 1;
            prev = val;
        }
        if (i >= w - 1) {                                        // calc_mean
            double res = sum_x / (double)nobs;
            if (same >= nobs) res = prev;
            else if (neg == 0 && res < 0) res = 0;
            else if (neg == nobs && res > 0) res = 0;
            o[i - (w - 1)] = res;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CLR (Scribe.py:142-176): z-scores of round(M, 10) against the row (and column) mean / std (ddof = 1) of M, negatives
// clipped to 0, combined as sqrt((z_row^2 + z_col^2) / 2).  The reference subtracts the row statistics without a keepdims
// axis when M is square, so for square M element (i, j) is z-scored against the statistics of ROW j (Scribe.py:151-154);
// `square_quirk` reproduces that.
// ------------------------------------------------------------------------------------------------
// mean and std (ddof = 1) of every row (axis = 1) or column (axis = 0) of M (R x C, row-major): one warp per line, lanes
// stride the line, fixed shuffle tree (deterministic; summation order differs from numpy's by O(1e-16) relative)
__global__ void __launch_bounds__(256)
line_mean_std_kernel(const double* __restrict__ M, int64_t R, int64_t C, int axis, double* __restrict__ mean, double* __restrict__ sd)
{
    const int lane = threadIdx.x & 31;
    const int64_t line = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_lines = axis == 1 ? R : C, len = axis == 1 ? C : R;
    if (line >= n_lines) return;
    const int64_t base = axis == 1 ? line * C : line, step = axis == 1 ? 1 : C;
    double s = 0.0;
    for (int64_t i = lane; i < len; i += 32) s += M[base + i * step];
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double mu = s / (double)len;
    double q = 0.0;
    for (int64_t i = lane; i < len; i += 32) { const double d = M[base + i * step] - mu; q += d * d; }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) { mean[line] = mu; sd[line] = sqrt(q / (double)(len - 1)); }
}

__global__ void __launch_bounds__(256)
clr_kernel(const double* __restrict__ M, int64_t R, int64_t C, const double* __restrict__ row_mean, const double* __restrict__ row_sd,
           const double* __restrict__ col_mean, const double* __restrict__ col_sd, int both, int square_quirk, double* __restrict__ out)
{
    const int64_t total = R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C, c = i - r * C;
        const double v = rint(M[i] * 1e10) / 1e10;               // np.round(mat, 10)
        const int64_t rs = square_quirk ? c : r;                 // square input: broadcast of the (n,) row statistics along the last axis
        double zr = (v - row_mean[rs]) / row_sd[rs];
        if (zr < 0.0) zr = 0.0;                                  // NaN stays NaN, as z[z < 0] = 0 leaves it
        double res;
        if (both) {
            double zc = (v - col_mean[c]) / col_sd[c];
            if (zc < 0.0) zc = 0.0;
            res = sqrt((zr * zr + zc * zc) / 2.0);
        } else {
            res = sqrt(zr * zr);
        }
        out[i] = res;
    }
}

// ------------------------------------------------------------------------------------------------
// Edge ranking.  Keys for a stable ascending radix sort = descending score order; NaN scores sort last.
//   mode 0 (top-n edges, plot/networks.py:34-47): every entry of the square matrix, row-major; an entry that is smaller than its
//           transpose (M[i][j] - M[j][i] < 0) is dropped (set NaN by the reference, removed by stack()).
//   mode 1 (roc, causal_network.py:351-359): off-diagonal entries only, source-major; `flag` carries the edge's truth bit.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
edge_keys_kernel(const double* __restrict__ M, int64_t n, int mode, double* __restrict__ keys, int32_t* __restrict__ idx)
{
    const int64_t total = n * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n, c = i - r * n;
        double w = M[i];
        bool drop = w != w;
        if (mode == 0) drop = drop || (w - M[c * n + r] < 0.0);
        else           drop = drop || (r == c);
        keys[i] = drop ? INFINITY : -w;
        idx[i] = (int32_t)i;
    }
}

// after the sort: the first n_valid entries are the ranked edges.  roc: cumulative true / false positives along the ranking
// (tp[0] = fp[0] = 0 as in the reference), one block-wide scan per 1024-edge tile chained through a running carry.
__global__ void __launch_bounds__(1024)
roc_scan_kernel(const int32_t* __restrict__ idx_sorted, const unsigned char* __restrict__ truth, int64_t m, int32_t* __restrict__ tp, int32_t* __restrict__ fp)
{
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { carry = 0; tp[0] = 0; fp[0] = 0; }
    __syncthreads();
    for (int64_t base = 0; base < m; base += 1024) {
        const int64_t i = base + tid;
        const int hit = (i < m) ? (truth[idx_sorted[i]] ? 1 : 0) : 0;
        int v = hit;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) warp_tot[wid] = v;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < wid; w++) off += warp_tot[w];
        if (i < m) { tp[i + 1] = off + v; fp[i + 1] = (int32_t)(i + 1) - (off + v); }
        __syncthreads();
        if (tid == 1023) carry = off + v;
        __syncthreads();
    }
}

// x = fp / max(fp), y = tp / max(tp) (causal_network.py:375-376) and the trapezoid terms |dx * (y_i + y_{i-1}) / 2| (:379-381)
__global__ void __launch_bounds__(256)
roc_curve_kernel(const int32_t* __restrict__ tp, const int32_t* __restrict__ fp, int64_t m, double* __restrict__ x, double* __restrict__ y,
                 double* __restrict__ area_terms)
{
    const double fmax_ = (double)fp[m], tmax_ = (double)tp[m];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= m; i += (int64_t)gridDim.x * blockDim.x) {
        const double xi = (double)fp[i] / fmax_, yi = (double)tp[i] / tmax_;
        x[i] = xi; y[i] = yi;
        if (i > 0) {
            const double xp = (double)fp[i - 1] / fmax_, yp = (double)tp[i - 1] / tmax_;
            area_terms[i - 1] = fabs((xi - xp) * (yi + yp) / 2.0);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// kNN grid regressor of viz_causality / viz_comb_logic (Scribe/plot/heatmaps.py:405-423, :578-596): for every query point
// (a grid_num x grid_num mesh, 625 by default) the k+1 nearest data points in the Euclidean plane (cKDTree.query(q, k+1)),
// the nearest one dropped, weights u = exp(-d / min d) normalised to 1, result = sum w * value[neighbour].
// One warp per query: every lane keeps the k+1 best of its stride of the data points, then the warp merges the 32 lists
// by k+1 rounds of arg-min.  Distances as scipy forms them: sqrt(dx*dx + dy*dy), no FMA.  Ties in distance resolve to the
// lower point index (the reference inherits whatever order the kd-tree traversal gives them).
// ------------------------------------------------------------------------------------------------
constexpr int KNN_REG_MAXK = 16;      // k + 1 <= 16

__global__ void __launch_bounds__(128)
knn_regress_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ value, int64_t n,
                   const double* __restrict__ qx, const double* __restrict__ qy, int64_t m, int k, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= m) return;
    const int kk = k + 1;
    const double x0 = qx[q], y0 = qy[q];
    double bd[KNN_REG_MAXK]; int64_t bi[KNN_REG_MAXK];
    #pragma unroll
    for (int t = 0; t < KNN_REG_MAXK; t++) { bd[t] = INFINITY; bi[t] = -1; }
    for (int64_t j = lane; j < n; j += 32) {
        const double dx = x0 - px[j], dy = y0 - py[j];
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (d2 < bd[KNN_REG_MAXK - 1] || (kk < KNN_REG_MAXK && d2 < bd[kk - 1])) {
            // insertion into the lane's ascending list of kk entries (indices ascend within equal distances: j only grows)
            #pragma unroll
            for (int t = KNN_REG_MAXK - 1; t >= 0; --t) {
                if (t < kk) {
                    const bool here = d2 < bd[t] && (t == 0 || !(d2 < bd[t - 1]));
                    const bool shift = t > 0 && d2 < bd[t - 1];
                    if (shift) { bd[t] = bd[t - 1]; bi[t] = bi[t - 1]; }
                    else if (here) { bd[t] = d2; bi[t] = j; }
                }
            }
        }
    }
    // merge: kk rounds; each round the warp takes the smallest head (ties: lower index) and that lane pops it
    double nd[KNN_REG_MAXK]; double nv[KNN_REG_MAXK];
    int head = 0;
    for (int r = 0; r < kk; r++) {
        double hd = INFINITY; int64_t hi = INT64_MAX;
        #pragma unroll
        for (int t = 0; t < KNN_REG_MAXK; t++) if (t == head) { hd = bd[t]; hi = bi[t] < 0 ? INT64_MAX : bi[t]; }
        if (head >= kk) { hd = INFINITY; hi = INT64_MAX; }
        double md = hd; int64_t mi = hi;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, md, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
        }
        if (hi == mi && mi != INT64_MAX) head++;
        nd[r] = (mi == INT64_MAX) ? INFINITY : sqrt(md);        // cKDTree pads missing neighbours with inf
        nv[r] = (mi == INT64_MAX) ? 0.0 : value[mi];
    }
    if (lane == 0) {
        // heatmaps.py:412-416: neighbours 1..k (the nearest is dropped), u = exp(-d / min d), w = u / sum u, sum(w * value)
        double dmin = INFINITY;
        for (int r = 1; r < kk; r++) dmin = fmin(dmin, nd[r]);
        double su = 0.0, acc = 0.0, u[KNN_REG_MAXK];
        for (int r = 1; r < kk; r++) { u[r] = exp(-nd[r] / dmin); su += u[r]; }
        for (int r = 1; r < kk; r++) acc += (u[r] / su) * nv[r];
        out[q] = acc;
    }
}

// sum of n doubles in index order (one thread: the reference accumulates the AUROC terms sequentially, :380-381)
__global__ void sequential_sum_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int64_t i = 0; i < n; i++) s += v[i];
        *out = s;
    }
}

}  // namespace scribe
