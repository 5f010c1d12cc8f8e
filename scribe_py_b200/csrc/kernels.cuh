// kernels.cuh -- sm_100a device code of the KSG hot path.
//
// Everything that decides a neighbour COUNT computes in FP64 on purpose (two documented exceptions below).  The reference
// (scipy cKDTree, float64) decides ball
// membership with `max_c |a_c - b_c| <= eps_i` where eps_i is itself such a difference, so neighbour
// counts are only reproducible bit-for-bit if the differences and comparisons are done in float64.
// B200 (unlike B300) keeps a real FP64 pipe: measured 64 lanes/clk/SM for DADD/DSETP
// (tools/microbench.cu -> profiles/microbench_r1.jsonl), i.e. half the FP32 FADD rate, which makes the
// exact path only ~1.4x slower than an (inexact) FP32 one -- see DESIGN.md "Why FP64".
// Exceptions, neither of which can change a count: (1) the KDE kernels (kde_kernel, kde_sweep_kernel) form the squared
// distance in FP64 but evaluate 2^x with ex2.approx.ftz.f32 and accumulate each chunk's terms in FP32 -- the uniformisation
// weights are then good to ~2e-6 relative, inside the 1e-5 budget of umi/cumi values (tests hold them to 1e-6); (2) the
// windowed sweep (sweep_window.cuh) pre-filters k-NN candidates in FP32 with a bound that makes the marked set a superset of
// the exact one, every marked candidate being re-tested in FP64.
//
// Layout: a "job" is one estimator evaluation on N points of joint dimension D = dx+dy+dz.  Its columns
// live dimension-major in HBM (D pointers to N contiguous doubles, JobDesc).  A CTA owns a block of
// TPB*Q query points held in registers and streams all N candidates through shared memory in tiles;
// every lane reads the same candidate (LDS broadcast), so the inner loops are pure FP64-pipe work:
//   pass 1 (kNN)   : 3 DADD + 3 chained DSETP per point pair, register top-(k+1) insert off the hot path
//   pass 2 (count) : D DADD + (dz+3) chained DSETP + 4 predicated integer adds per point pair
// followed by the digamma terms and a deterministic block reduction (one partial per CTA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <type_traits>

namespace scribe {

constexpr int TPB  = 256;     // threads per CTA (all kernels)
constexpr int TILE = 512;     // candidates staged per shared-memory tile
constexpr int MAXD = 16;      // == SCRIBE_MAX_DIM
constexpr int GEN_KCAP = 64;  // generic kernel: k+1 <= 64

struct JobDesc {
    const double* col[MAXD];  // x dims, then y dims, then z dims
    const double* w;          // uniformisation weights (cumi/umi) or nullptr; indexed like the columns
    const int32_t* perm;      // sweep kernels: rank -> row index, rows sorted ascending by the LAST column (nullptr = identity)
    const double* strip;      // strip kernels: this job's x-sorted z-strips (see strip_sort_kernel), or nullptr
    int64_t npad;             //   padded length (multiple of STRIP) of each strip column
    int64_t qoff;             // offset of this job's per-point outputs (eps/counts/...) in the workspace
    int32_t n;                // points
    int32_t pad;
    double absmax;            // windowed sweep: largest |coordinate| of the job, written on the device by job_absmax_kernel
    // windowed sweep, rank-space count pass: for every NON-key coordinate c (x, y, then the extra conditioning columns)
    const double* srt[4];     //   the values of that coordinate in ascending order (srt_n[c] of them): the ball |q_c - v| <= eps is one
    const float*  rk[4];      //   rank range of it; rk[c][row] = position of the job's row in that order (an integer < 2^22, as float)
    int32_t srt_n[4];         //   (srt_n may exceed n: dynamics coupling ranks a job's kept cells among ALL cells of the gene)
    const double* wcum;       // umi on the sweep: exclusive prefix sums of the weights in sweep (y) order, n + 1 values
};

struct PointOut {             // optional per-point outputs (debug entry point, generic path, kNN-only mode)
    double*  eps;             // [qoff + i]
    int32_t* cnt;             // [4*(qoff+i) + s]
    double*  wsum;            // [2*(qoff+i) + s]
};

enum : int { MODE_CMI = 0, MODE_MI = 1, MODE_KNN_ONLY = 2 };

// ------------------------------------------------------------------------------------------------
// digamma for real argument, double precision (replaces scipy.special.digamma at
// information_estimators.py:102-107,:169-172,:264,:399-402).  |error| ~1e-15, far inside the 1e-5 budget.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline double digamma_pos(double x) {   // x > 0, finite
    double acc = 0.0;
    while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
    const double xi = 1.0 / x, x2 = xi * xi;
    const double ser = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240 - x2 * (1.0 / 132
                       - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
    return acc + log(x) - 0.5 * xi - ser;
}

__host__ __device__ inline double digamma_f64(double x) {
    if (x != x) return x;
    if (x == 0.0) return copysign(INFINITY, -x);          // scipy: digamma(0.0) = -inf, digamma(-0.0) = +inf
    if (isinf(x)) return x > 0 ? x : NAN;
    if (x < 0.0) {                                        // reflection: psi(x) = psi(1-x) - pi/tan(pi x)
        const double r = x - floor(x);
        if (r == 0.0) return NAN;
        return digamma_pos(1.0 - x) - 3.14159265358979323846 / tan(3.14159265358979323846 * r);
    }
    return digamma_pos(x);
}

__device__ __forceinline__ double psi_int(const double* __restrict__ tab, int n) { return __ldg(tab + n); }

// Deterministic block sum: fixed shuffle tree inside a warp, fixed order across warps.
__device__ __forceinline__ double block_sum(double v, double* smem8) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem8[wid] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
        #pragma unroll
        for (int w = 0; w < TPB / 32; w++) s += smem8[w];
    }
    return s;   // valid in thread 0
}

// ------------------------------------------------------------------------------------------------
// Fused dense kernel (Chebyshev metric): mi / cmi / cumi and kNN-only.
//   replaces tree_xyz.query(...)[0][k] + 3-4 query_ball_point per point + digamma + mean
//   (information_estimators.py:101-109, :165-173, :396-403)
// ------------------------------------------------------------------------------------------------
template <int KCAP>
__device__ __forceinline__ void topk_insert(double (&L)[KCAP], double d) {
    // L ascending; caller guarantees d < L[KCAP-1].  All position tests are independent of one another and every new
    // slot depends only on old slots, so the dependency depth is 2 instead of the 2*(KCAP-1) of a compare-swap chain
    // (the insert path is latency-bound: few lanes are active and the FP64 compare is ~10 cycles).
    bool p[KCAP];
    #pragma unroll
    for (int t = 0; t < KCAP - 1; t++) p[t] = d < L[t];
    p[KCAP - 1] = true;
    #pragma unroll
    for (int t = KCAP - 1; t > 0; --t) L[t] = p[t - 1] ? L[t - 1] : (p[t] ? d : L[t]);
    L[0] = p[0] ? d : L[0];
}

// predicated accumulators: one issue slot each (the compiler otherwise emits add + select pairs)
__device__ __forceinline__ void inc_if(int& n, bool p) {
    asm("{.reg .pred q; setp.ne.b32 q, %1, 0; @q add.s32 %0, %0, 1;}" : "+r"(n) : "r"((int)p));
}
__device__ __forceinline__ void add_if(double& acc, double v, bool p) {
    asm("{.reg .pred q; setp.ne.b32 q, %2, 0; @q add.f64 %0, %0, %1;}" : "+d"(acc) : "d"(v), "r"((int)p));
}

// One (query, candidate) step of the count pass as a single PTX block: the predicates live only inside the block, so the
// compiler cannot spill them to registers when several queries and unrolled candidates are in flight (7 predicate
// registers per thread).  dq* are the coordinate differences q - c; semantics as in the reference's four ball tests
// (information_estimators.py:169-172): nz += [|dz|<=e], nxz += [|dx|<=e & z], nyz += [|dy|<=e & z], nxyz += [all].
template <int DZ, bool WEIGHTED>
__device__ __forceinline__ void count_step(double dx, double dy, const double (&dz)[DZ > 0 ? DZ : 1], double e, double wj,
                                           int& nz, int& nxz, int& nyz, int& nxyz, double& wz, double& wyz)
{
    static_assert(DZ >= 1 && DZ <= 3, "count_step supports 1..3 conditioning dimensions");
    if (DZ == 1) {
        if (WEIGHTED)
            asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az;\n abs.f64 ax,%6; abs.f64 ay,%7; abs.f64 az,%8;\n"
                " setp.le.f64 pz, az, %9;\n setp.le.and.f64 pxz, ax, %9, pz;\n setp.le.and.f64 pyz, ay, %9, pz;\n setp.le.and.f64 pxyz, ay, %9, pxz;\n"
                " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n"
                " @pz add.f64 %4,%4,%10;\n @pyz add.f64 %5,%5,%10;\n}"
                : "+r"(nz), "+r"(nxz), "+r"(nyz), "+r"(nxyz), "+d"(wz), "+d"(wyz) : "d"(dx), "d"(dy), "d"(dz[0]), "d"(e), "d"(wj));
        else
            asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az;\n abs.f64 ax,%4; abs.f64 ay,%5; abs.f64 az,%6;\n"
                " setp.le.f64 pz, az, %7;\n setp.le.and.f64 pxz, ax, %7, pz;\n setp.le.and.f64 pyz, ay, %7, pz;\n setp.le.and.f64 pxyz, ay, %7, pxz;\n"
                " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n}"
                : "+r"(nz), "+r"(nxz), "+r"(nyz), "+r"(nxyz) : "d"(dx), "d"(dy), "d"(dz[0]), "d"(e));
    } else if (DZ == 2) {
        if (WEIGHTED)
            asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az,aw;\n abs.f64 ax,%6; abs.f64 ay,%7; abs.f64 az,%8; abs.f64 aw,%9;\n"
                " setp.le.f64 pz, az, %10;\n setp.le.and.f64 pz, aw, %10, pz;\n setp.le.and.f64 pxz, ax, %10, pz;\n setp.le.and.f64 pyz, ay, %10, pz;\n setp.le.and.f64 pxyz, ay, %10, pxz;\n"
                " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n"
                " @pz add.f64 %4,%4,%11;\n @pyz add.f64 %5,%5,%11;\n}"
                : "+r"(nz), "+r"(nxz), "+r"(nyz), "+r"(nxyz), "+d"(wz), "+d"(wyz) : "d"(dx), "d"(dy), "d"(dz[0]), "d"(dz[DZ > 1 ? 1 : 0]), "d"(e), "d"(wj));
        else
            asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az,aw;\n abs.f64 ax,%4; abs.f64 ay,%5; abs.f64 az,%6; abs.f64 aw,%7;\n"
                " setp.le.f64 pz, az, %8;\n setp.le.and.f64 pz, aw, %8, pz;\n setp.le.and.f64 pxz, ax, %8, pz;\n setp.le.and.f64 pyz, ay, %8, pz;\n setp.le.and.f64 pxyz, ay, %8, pxz;\n"
                " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n}"
                : "+r"(nz), "+r"(nxz), "+r"(nyz), "+r"(nxyz) : "d"(dx), "d"(dy), "d"(dz[0]), "d"(dz[DZ > 1 ? 1 : 0]), "d"(e));
    } else {
        if (WEIGHTED)
            asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az,aw,av;\n abs.f64 ax,%6; abs.f64 ay,%7; abs.f64 az,%8; abs.f64 aw,%9; abs.f64 av,%10;\n"
                " setp.le.f64 pz, az, %11;\n setp.le.and.f64 pz, aw, %11, pz;\n setp.le.and.f64 pz, av, %11, pz;\n setp.le.and.f64 pxz, ax, %11, pz;\n setp.le.and.f64 pyz, ay, %11, pz;\n setp.le.and.f64 pxyz, ay, %11, pxz;\n"
                " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n"
                " @pz add.f64 %4,%4,%12;\n @pyz add.f64 %5,%5,%12;\n}"
                : "+r"(nz), "+r"(nxz), "+r"(nyz), "+r"(nxyz), "+d"(wz), "+d"(wyz)
                : "d"(dx), "d"(dy), "d"(dz[0]), "d"(dz[DZ > 1 ? 1 : 0]), "d"(dz[DZ > 2 ? 2 : 0]), "d"(e), "d"(wj));
        else
            asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az,aw,av;\n abs.f64 ax,%4; abs.f64 ay,%5; abs.f64 az,%6; abs.f64 aw,%7; abs.f64 av,%8;\n"
                " setp.le.f64 pz, az, %9;\n setp.le.and.f64 pz, aw, %9, pz;\n setp.le.and.f64 pz, av, %9, pz;\n setp.le.and.f64 pxz, ax, %9, pz;\n setp.le.and.f64 pyz, ay, %9, pz;\n setp.le.and.f64 pxyz, ay, %9, pxz;\n"
                " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n}"
                : "+r"(nz), "+r"(nxz), "+r"(nyz), "+r"(nxyz)
                : "d"(dx), "d"(dy), "d"(dz[0]), "d"(dz[DZ > 1 ? 1 : 0]), "d"(dz[DZ > 2 ? 2 : 0]), "d"(e));
    }
}

// One (query, candidate) step of the kNN scan as a single PTX block: sets `bit` in `mask` when every coordinate
// difference is strictly below the query's current k-th distance `thr` (strict, like the insert test).
template <int D>
__device__ __forceinline__ void knn_mark(const double (&df)[D], double thr, unsigned bit, unsigned& mask)
{
    static_assert(D >= 3 && D <= 5, "knn_mark supports joint dimension 3..5");
    if (D == 3)
        asm("{\n .reg .pred p;\n .reg .f64 a0,a1,a2;\n abs.f64 a0,%1; abs.f64 a1,%2; abs.f64 a2,%3;\n"
            " setp.lt.f64 p, a0, %4;\n setp.lt.and.f64 p, a1, %4, p;\n setp.lt.and.f64 p, a2, %4, p;\n @p or.b32 %0,%0,%5;\n}"
            : "+r"(mask) : "d"(df[0]), "d"(df[1]), "d"(df[2]), "d"(thr), "r"(bit));
    else if (D == 4)
        asm("{\n .reg .pred p;\n .reg .f64 a0,a1,a2,a3;\n abs.f64 a0,%1; abs.f64 a1,%2; abs.f64 a2,%3; abs.f64 a3,%4;\n"
            " setp.lt.f64 p, a0, %5;\n setp.lt.and.f64 p, a1, %5, p;\n setp.lt.and.f64 p, a2, %5, p;\n setp.lt.and.f64 p, a3, %5, p;\n @p or.b32 %0,%0,%6;\n}"
            : "+r"(mask) : "d"(df[0]), "d"(df[1]), "d"(df[2]), "d"(df[D > 3 ? 3 : 0]), "d"(thr), "r"(bit));
    else
        asm("{\n .reg .pred p;\n .reg .f64 a0,a1,a2,a3,a4;\n abs.f64 a0,%1; abs.f64 a1,%2; abs.f64 a2,%3; abs.f64 a3,%4; abs.f64 a4,%5;\n"
            " setp.lt.f64 p, a0, %6;\n setp.lt.and.f64 p, a1, %6, p;\n setp.lt.and.f64 p, a2, %6, p;\n setp.lt.and.f64 p, a3, %6, p;\n setp.lt.and.f64 p, a4, %6, p;\n"
            " @p or.b32 %0,%0,%7;\n}"
            : "+r"(mask) : "d"(df[0]), "d"(df[1]), "d"(df[2]), "d"(df[D > 3 ? 3 : 0]), "d"(df[D > 4 ? 4 : 0]), "d"(thr), "r"(bit));
}

// Shared-memory tile of candidates: D rows of `cap` doubles (+ one row of weights), dynamic size.
// cap = min(round_up(n_max, 4), FUSED_TILE_MAX): jobs up to FUSED_TILE_MAX points are staged ONCE for both passes.
constexpr int FUSED_TILE_MAX = 2048;

template <int DX, int DY, int DZ, int KCAP, int Q, bool WEIGHTED>
__global__ void __launch_bounds__(TPB)
ksg_fused_kernel(const JobDesc* __restrict__ jobs, int blocks_per_job, int k, int mode, int cap,
                 const double* __restrict__ psi_tab, double* __restrict__ partial, PointOut po)
{
    constexpr int D = DX + DY + DZ;
    extern __shared__ double smem[];
    double* tile  = smem;                 // [D][cap]
    double* wtile = smem + D * cap;       // [cap] (WEIGHTED only)
    __shared__ double red[TPB / 32];

    const int job = blockIdx.x / blocks_per_job;
    const int qb  = blockIdx.x - job * blocks_per_job;
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    const int q0 = qb * (TPB * Q);
    if (q0 >= n) return;                                   // uniform: this CTA has no query points
    const int tid = threadIdx.x;

    double qc[Q][D];
    int    qi[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = q0 + tid * Q + q;
        qi[q] = i;
        const int ii = i < n ? i : n - 1;                  // idle slots shadow the last point (results dropped)
        #pragma unroll
        for (int c = 0; c < D; c++) qc[q][c] = jd.col[c][ii];
    }

    const int ntiles = (n + cap - 1) / cap;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    auto load_tile = [&](int t) {
        const int base = t * cap;
        const int cnt = min(cap, n - base);
        const int cnt4 = (cnt + 3) & ~3;
        __syncthreads();
        for (int i = tid; i < cnt4; i += TPB) {
            #pragma unroll
            for (int c = 0; c < D; c++) tile[c * cap + i] = (i < cnt) ? jd.col[c][base + i] : qnan;   // NaN never compares true
            if (WEIGHTED) wtile[i] = (i < cnt) ? jd.w[base + i] : 0.0;
        }
        __syncthreads();
        return cnt4;
    };

    // ---------------- pass 1: k-th neighbour distance in the joint space ----------------
    // L holds the KCAP smallest values seen; the lowest KCAP-(k+1) slots are pre-filled with -inf sentinels
    // so that L[KCAP-1] is always the (k+1)-th smallest joint distance (self included) = eps_i.
    double L[Q][KCAP];
    #pragma unroll
    for (int q = 0; q < Q; q++)
        #pragma unroll
        for (int t = 0; t < KCAP; t++) L[q][t] = (t < KCAP - 1 - k) ? -INFINITY : INFINITY;

    int cnt4 = 0;
    for (int tt = 0; tt < ntiles; tt++) {
        cnt4 = load_tile(tt);
        #pragma unroll 2
        for (int j = 0; j < cnt4; j += 2) {
            double ad[2][Q][D];
            bool lt[2][Q];
            bool any = false;
            #pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                #pragma unroll
                for (int q = 0; q < Q; q++) {
                    bool l = true;
                    #pragma unroll
                    for (int c = 0; c < D; c++) { ad[jj][q][c] = fabs(qc[q][c] - tile[c * cap + j + jj]); l = l & (ad[jj][q][c] < L[q][KCAP - 1]); }
                    lt[jj][q] = l;
                    any = any | l;
                }
            }
            if (any) {                                      // rare once the threshold has settled
                #pragma unroll
                for (int jj = 0; jj < 2; jj++) {
                    #pragma unroll
                    for (int q = 0; q < Q; q++) {
                        double d = ad[jj][q][0];
                        #pragma unroll
                        for (int c = 1; c < D; c++) d = fmax(d, ad[jj][q][c]);
                        if (lt[jj][q] && d < L[q][KCAP - 1]) topk_insert<KCAP>(L[q], d);   // re-check: the first insert moves the bar
                    }
                }
            }
        }
    }

    double eps[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) eps[q] = L[q][KCAP - 1];   // +inf when n <= k, like cKDTree's padding

    if (mode == MODE_KNN_ONLY) {
        #pragma unroll
        for (int q = 0; q < Q; q++) if (qi[q] < n) po.eps[jd.qoff + qi[q]] = eps[q];
        return;
    }

    // ---------------- pass 2: closed-ball counts in the marginal subspaces ----------------
    int nxyz[Q], nxz[Q], nyz[Q], nz[Q];
    double wyz[Q], wz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) { nxyz[q] = nxz[q] = nyz[q] = nz[q] = 0; wyz[q] = wz[q] = 0.0; }

    for (int tt = 0; tt < ntiles; tt++) {
        if (ntiles > 1) cnt4 = load_tile(tt);              // a single tile is still resident from pass 1
        #pragma unroll 4
        for (int j = 0; j < cnt4; j++) {
            double cc[D];
            #pragma unroll
            for (int c = 0; c < D; c++) cc[c] = tile[c * cap + j];
            double wj = 0.0;
            if (WEIGHTED) wj = wtile[j];
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                bool pz = true, px = true, py = true;
                #pragma unroll
                for (int c = 0; c < DZ; c++) pz = pz & (fabs(qc[q][DX + DY + c] - cc[DX + DY + c]) <= eps[q]);
                #pragma unroll
                for (int c = 0; c < DX; c++) px = px & (fabs(qc[q][c] - cc[c]) <= eps[q]);
                #pragma unroll
                for (int c = 0; c < DY; c++) py = py & (fabs(qc[q][DX + c] - cc[DX + c]) <= eps[q]);
                if (DZ == 0) pz = !(cc[0] != cc[0]);       // mi: every real candidate is in the (empty-z) ball; padding is not
                const bool pxz = px & pz, pyz = py & pz, pxyz = pxz & py;
                inc_if(nz[q], pz); inc_if(nxz[q], pxz); inc_if(nyz[q], pyz); inc_if(nxyz[q], pxyz);
                if (WEIGHTED) { add_if(wz[q], wj, pz); add_if(wyz[q], wj, pyz); }
            }
        }
    }

    // ---------------- digamma terms + reduction ----------------
    double acc = 0.0;
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = qi[q];
        if (i < n) {
            const int a = nxyz[q] - 1, b = nxz[q] - 1, c = nyz[q] - 1, d = nz[q] - 1;   // self removed
            double term;
            if (WEIGHTED) {
                const double wi = jd.w[i];
                const double Wyz = wyz[q] - wi, Wz = wz[q] - wi;
                // information_estimators.py:399-402
                term = wi * psi_int(psi_tab, a) + wi * -psi_int(psi_tab, b) + wi * -digamma_f64(Wyz) + wi * digamma_f64(Wz);
                if (po.wsum) { po.wsum[2 * (jd.qoff + i)] = Wyz; po.wsum[2 * (jd.qoff + i) + 1] = Wz; }
            } else if (mode == MODE_MI) {
                // information_estimators.py:102-107 (here x -> "xz" slot, y -> "yz" slot)
                term = psi_int(psi_tab, n) + psi_int(psi_tab, a) - psi_int(psi_tab, b) - psi_int(psi_tab, c);
            } else {
                // information_estimators.py:169-172
                term = psi_int(psi_tab, a) - psi_int(psi_tab, b) - psi_int(psi_tab, c) + psi_int(psi_tab, d);
            }
            acc += term;
            if (po.eps) po.eps[jd.qoff + i] = eps[q];
            if (po.cnt) { int32_t* o = po.cnt + 4 * (jd.qoff + i); o[0] = a; o[1] = b; o[2] = c; o[3] = d; }
        }
    }
    const double s = block_sum(acc, red);
    if (tid == 0) partial[blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------
// Sorted-sweep kernels.  Every subspace the conditional estimators count in (xyz, xz, yz, z) contains the
// conditioning coordinates, so a candidate with |z_j - z_i| > eps_i can be in none of the balls -- and, for the
// same reason, cannot be among the k nearest joint neighbours once the running k-th distance has dropped below
// |z_j - z_i|.  The points of a job are therefore visited in the order of their last conditioning coordinate
// (JobDesc.perm, one device sort per (target gene, lag) column, shared by every source gene that reads it).
// A warp owns 32*Q consecutive ranks and sweeps outwards from them in chunks of 32 candidates, left and right,
// until the nearest unvisited candidate on that side is too far in z for every one of its queries:
//   pass 1 stops a side when fl(|z_i - z_next|) >= current k-th distance of every query (strict insert test),
//   pass 2 stops a side when fl(|z_i - z_next|) >  eps_i for every query (closed ball),
// both evaluated with the very subtraction the ball test uses, so by monotonicity of rounding the visited set is
// an exact superset of what the dense kernel would have accepted: results are bit-identical to it.
// No block-level synchronisation: warps are independent (chunks are staged in a per-warp shared-memory buffer).
// ------------------------------------------------------------------------------------------------
// Sweep kernels: warps are independent and their windows differ in size, so CTAs are kept small (SWEEP_TPB threads)
// to let a finished warp's slot be refilled early; SWEEP_MIN_CTAS caps registers for 24 resident warps per SM.
#ifndef SWEEP_TPB
#define SWEEP_TPB 64
#endif
#ifndef SWEEP_MIN_CTAS
#define SWEEP_MIN_CTAS (768 / SWEEP_TPB)
#endif
template <int D, bool WEIGHTED>
struct Chunk { double v[D]; double w; };

template <int D, bool WEIGHTED>
__device__ __forceinline__ Chunk<D, WEIGHTED> load_chunk(const JobDesc& jd, int chunk, int lane, int n) {
    Chunk<D, WEIGHTED> c;
    const int rank = chunk * 32 + lane;
    const bool ok = rank < n;
    const int idx = ok ? (jd.perm ? jd.perm[rank] : rank) : 0;
    #pragma unroll
    for (int d = 0; d < D; d++) c.v[d] = ok ? jd.col[d][idx] : __longlong_as_double(0x7ff8000000000000LL);
    c.w = 0.0;
    if (WEIGHTED) c.w = ok ? jd.w[idx] : 0.0;
    return c;
}

template <int DX, int DY, int DZ, int KCAP, int Q, bool WEIGHTED>
__global__ void __launch_bounds__(SWEEP_TPB, SWEEP_MIN_CTAS)
ksg_sweep_kernel(const JobDesc* __restrict__ jobs, int warps_per_job, int k, const double* __restrict__ psi_tab,
                 double* __restrict__ partial, PointOut po, unsigned long long* __restrict__ visited)
{
    static_assert(DX == 1 && DY == 1 && DZ >= 1, "sweep kernels: scalar x and y, at least one conditioning column");
    constexpr int D = DX + DY + DZ;
    constexpr int KEY = D - 1;
    constexpr int W = SWEEP_TPB / 32;
    __shared__ __align__(16) double sbuf[W][D + (WEIGHTED ? 1 : 0)][32];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bpj = (warps_per_job + W - 1) / W;
    const int job = blockIdx.x / bpj;
    const int wj  = (blockIdx.x - job * bpj) * W + wid;          // warp index inside the job
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    const int w0 = wj * 32 * Q;
    if (wj >= warps_per_job || w0 >= n) return;                  // warp-uniform; nothing below synchronises the CTA
    const int c0 = w0 / 32;                                      // first centre chunk
    const int nchunks = (n + 31) / 32;
    double (*buf)[32] = sbuf[wid];

    // queries = the centre chunks themselves (rank = w0 + q*32 + lane); idle slots shadow the last point
    double qc[Q][D], qw[Q];
    int qidx[Q]; bool qok[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int rank = w0 + q * 32 + lane;
        qok[q] = rank < n;
        const int rr = qok[q] ? rank : n - 1;
        qidx[q] = jd.perm ? jd.perm[rr] : rr;
        #pragma unroll
        for (int d = 0; d < D; d++) qc[q][d] = jd.col[d][qidx[q]];
        qw[q] = 0.0;
        if (WEIGHTED) qw[q] = jd.w[qidx[q]];
    }

    auto stage = [&](const Chunk<D, WEIGHTED>& c) {              // registers -> this warp's buffer
        __syncwarp();
        #pragma unroll
        for (int d = 0; d < D; d++) buf[d][lane] = c.v[d];
        if (WEIGHTED) buf[D][lane] = c.w;
        __syncwarp();
    };
    // Sweep shared by both passes: the warp's own chunks first, then outwards, taking whichever side still has a
    // candidate that some query of the warp needs (need(z_next) -> bit q set when this lane's query in slot q needs it),
    // alternating when both sides do.  After regroup() below, slot Q-1 holds the warp's sparsest queries (largest
    // eps), so the far part of the sweep is needed by that slot only: work(false) then skips the other slots.
    auto sweep = [&](auto need, auto work) -> unsigned long long {
        unsigned long long visits = 0;
        int cl = c0 - 1, cr = c0 + Q;
        bool toggle = false;
        Chunk<D, WEIGHTED> Lc, Rc, cur;
        if (cl >= 0) Lc = load_chunk<D, WEIGHTED>(jd, cl, lane, n);
        if (cr < nchunks) Rc = load_chunk<D, WEIGHTED>(jd, cr, lane, n);
        while (true) {
            unsigned al = 0, ar = 0;
            if (cl >= 0)      al = __reduce_or_sync(0xffffffffu, need(__shfl_sync(0xffffffffu, Lc.v[KEY], 31)));
            if (cr < nchunks) ar = __reduce_or_sync(0xffffffffu, need(__shfl_sync(0xffffffffu, Rc.v[KEY], 0)));
            if (!al && !ar) break;
            const bool left = al && (!ar || toggle);
            toggle = !toggle;
            const unsigned act = left ? al : ar;
            if (left) { cur = Lc; --cl; if (cl >= 0) Lc = load_chunk<D, WEIGHTED>(jd, cl, lane, n); }
            else      { cur = Rc; ++cr; if (cr < nchunks) Rc = load_chunk<D, WEIGHTED>(jd, cr, lane, n); }
            stage(cur);
            const bool all = (act & ((1u << (Q - 1)) - 1u)) != 0u || Q == 1;
            work(all);
            visits += all ? Q : 1;
        }
        return visits;
    };
    auto centre = [&](auto work) -> unsigned long long {
        unsigned long long visits = 0;
        for (int c = c0; c < c0 + Q && c < nchunks; c++) { stage(load_chunk<D, WEIGHTED>(jd, c, lane, n)); work(true); visits += Q; }
        return visits;
    };

    // ---------------- pass 1 ----------------
    double L[Q][KCAP];
    #pragma unroll
    for (int q = 0; q < Q; q++)
        #pragma unroll
        for (int t = 0; t < KCAP; t++) L[q][t] = (t < KCAP - 1 - k) ? -INFINITY : INFINITY;

    // Per chunk: (i) a branch-free scan marks, per query, which of the 32 candidates beat the k-th distance the query
    // had when the chunk started (one bit each); (ii) the marked candidates are then inserted in a short per-lane loop.
    // Inside the z-window nearly every warp-step would otherwise contain some lane that inserts (the number of inserts per
    // query, ~ (k+1) ln(window), does not shrink with the window), so batching them per chunk is what keeps the
    // divergent insert path off the scan.
    auto knn_body = [&](auto q_first) {
        constexpr int Q0 = decltype(q_first)::value;
        unsigned mask[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) mask[q] = 0u;
        #pragma unroll 8
        for (int j = 0; j < 32; j++) {
            double cc[D];
            #pragma unroll
            for (int d = 0; d < D; d++) cc[d] = buf[d][j];
            const unsigned bit = 1u << j;
            #pragma unroll
            for (int q = Q0; q < Q; q++) {
                double df[D];
                #pragma unroll
                for (int d = 0; d < D; d++) df[d] = qc[q][d] - cc[d];
                knn_mark<D>(df, L[q][KCAP - 1], bit, mask[q]);
            }
        }
        #pragma unroll
        for (int q = Q0; q < Q; q++) {
            while (mask[q]) {
                const int j = __ffs(mask[q]) - 1;
                mask[q] &= mask[q] - 1;
                double dd = fabs(qc[q][0] - buf[0][j]);
                #pragma unroll
                for (int d = 1; d < D; d++) { const double a = fabs(qc[q][d] - buf[d][j]); dd = (a > dd) ? a : dd; }   // no NaNs here
                if (dd < L[q][KCAP - 1]) topk_insert<KCAP>(L[q], dd);          // the bar may have moved since the scan
            }
        }
    };
    auto knn_chunk = [&](bool all) {
        if (all) knn_body(std::integral_constant<int, 0>{}); else knn_body(std::integral_constant<int, Q - 1>{});
    };

    unsigned long long nvis = centre(knn_chunk);

    // Regroup the warp's 32*Q queries by local sparsity: key = k-th distance among the warp's own points (a close proxy of
    // eps: it orders the queries almost as the final eps does).  Slot q receives the queries of rank [32q, 32q+32), so
    // the lanes of a slot stop needing far chunks at about the same time.  Every per-query value moves with its query.
    if (Q > 1) {
        double* kb = &buf[0][0];                                   // >= 32*Q doubles (D + WEIGHTED >= Q)
        double key[Q]; int pos[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) key[q] = qok[q] ? L[q][KCAP - 1] : INFINITY;
        __syncwarp();
        #pragma unroll
        for (int q = 0; q < Q; q++) kb[q * 32 + lane] = key[q];
        __syncwarp();
        #pragma unroll
        for (int q = 0; q < Q; q++) pos[q] = 0;
        for (int j = 0; j < 32 * Q; j++) {
            const double kj = kb[j];
            #pragma unroll
            for (int q = 0; q < Q; q++) pos[q] += ((kj < key[q]) | ((kj == key[q]) & (j < q * 32 + lane))) ? 1 : 0;
        }
        auto move = [&](double (&v)[Q]) {
            __syncwarp();
            #pragma unroll
            for (int q = 0; q < Q; q++) kb[pos[q]] = v[q];
            __syncwarp();
            #pragma unroll
            for (int q = 0; q < Q; q++) v[q] = kb[q * 32 + lane];
        };
        double t[Q];
        #pragma unroll
        for (int d = 0; d < D; d++) {
            #pragma unroll
            for (int q = 0; q < Q; q++) t[q] = qc[q][d];
            move(t);
            #pragma unroll
            for (int q = 0; q < Q; q++) qc[q][d] = t[q];
        }
        #pragma unroll
        for (int e = 0; e < KCAP; e++) {
            #pragma unroll
            for (int q = 0; q < Q; q++) t[q] = L[q][e];
            move(t);
            #pragma unroll
            for (int q = 0; q < Q; q++) L[q][e] = t[q];
        }
        if (WEIGHTED) move(qw);
        #pragma unroll
        for (int q = 0; q < Q; q++) t[q] = (double)(2 * (int64_t)qidx[q] + (qok[q] ? 1 : 0));     // exact: < 2^32
        move(t);
        #pragma unroll
        for (int q = 0; q < Q; q++) { const int64_t c = (int64_t)t[q]; qidx[q] = (int)(c >> 1); qok[q] = (c & 1) != 0; }
        __syncwarp();
    }

    nvis += sweep(
        [&](double znext) { unsigned f = 0u;
                            #pragma unroll
                            for (int q = 0; q < Q; q++) f |= (fabs(qc[q][KEY] - znext) < L[q][KCAP - 1]) ? (1u << q) : 0u;
                            return f; },
        knn_chunk);

    double eps[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) eps[q] = L[q][KCAP - 1];

    // ---------------- pass 2 ----------------
    int nxyz[Q], nxz[Q], nyz[Q], nz[Q];
    double wyz[Q], wz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) { nxyz[q] = nxz[q] = nyz[q] = nz[q] = 0; wyz[q] = wz[q] = 0.0; }

    auto count_body = [&](auto q_first) {
        constexpr int Q0 = decltype(q_first)::value;
        #pragma unroll 4
        for (int j = 0; j < 32; j++) {
            double cc[D];
            #pragma unroll
            for (int d = 0; d < D; d++) cc[d] = buf[d][j];
            double wj_ = 0.0;
            if (WEIGHTED) wj_ = buf[D][j];
            #pragma unroll
            for (int q = Q0; q < Q; q++) {
                double dzv[DZ];
                #pragma unroll
                for (int d = 0; d < DZ; d++) dzv[d] = qc[q][DX + DY + d] - cc[DX + DY + d];
                count_step<DZ, WEIGHTED>(qc[q][0] - cc[0], qc[q][DX] - cc[DX], dzv, eps[q], wj_,
                                         nz[q], nxz[q], nyz[q], nxyz[q], wz[q], wyz[q]);
            }
        }
    };
    auto count_chunk = [&](bool all) {
        if (all) count_body(std::integral_constant<int, 0>{}); else count_body(std::integral_constant<int, Q - 1>{});
    };
    unsigned long long nvis2 = centre(count_chunk);
    nvis2 += sweep(
        [&](double znext) { unsigned f = 0u;
                            #pragma unroll
                            for (int q = 0; q < Q; q++) f |= (fabs(qc[q][KEY] - znext) <= eps[q]) ? (1u << q) : 0u;
                            return f; },
        count_chunk);

    // ---------------- digamma terms + per-warp partial (fixed shuffle tree) ----------------
    double acc = 0.0;
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        if (qok[q]) {
            const int a = nxyz[q] - 1, b = nxz[q] - 1, c = nyz[q] - 1, d = nz[q] - 1;
            const int64_t o = jd.qoff + qidx[q];
            double term;
            if (WEIGHTED) {
                const double wi = qw[q];
                const double Wyz = wyz[q] - wi, Wz = wz[q] - wi;
                term = wi * psi_int(psi_tab, a) + wi * -psi_int(psi_tab, b) + wi * -digamma_f64(Wyz) + wi * digamma_f64(Wz);
                if (po.wsum) { po.wsum[2 * o] = Wyz; po.wsum[2 * o + 1] = Wz; }
            } else {
                term = psi_int(psi_tab, a) - psi_int(psi_tab, b) - psi_int(psi_tab, c) + psi_int(psi_tab, d);
            }
            acc += term;
            if (po.eps) po.eps[o] = eps[q];
            if (po.cnt) { int32_t* oc = po.cnt + 4 * o; oc[0] = a; oc[1] = b; oc[2] = c; oc[3] = d; }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        partial[(size_t)job * warps_per_job + wj] = acc;
        if (visited) { atomicAdd(visited, nvis * 32ull * 32ull); atomicAdd(visited + 1, nvis2 * 32ull * 32ull); }
    }
}

// ------------------------------------------------------------------------------------------------
// Strip kernels.  On expression data the conditioning coordinate z (the target's past) and y (its present) are strongly
// correlated, so the z-window of a query holds hundreds of candidates while the (x, z) box holds ten to thirty: x, which
// comes from another gene, is what prunes.  Each job therefore gets a small two-level index: its points in z-order
// (JobDesc.perm, shared per target column) are cut into strips of STRIP consecutive z-ranks and every strip is sorted by
// x (strip_sort_kernel, one warp per strip, per job: O(N log STRIP), negligible against O(N * window)).
//   step A (lane-private): each query walks the strips outwards from its own, nearest first, and inside a strip visits
//           only the x-range [x_i - bar, x_i + bar] found by binary search: the k-th neighbour distance (bar = running
//           k-th distance, strict) and then n_xz, n_xyz (bar = eps, closed).  A side is abandoned when the strip's nearest
//           z is already too far.  All tests are the reference's own subtractions, so the pruning is exact.
//   step B (collective, the sweep machinery above): n_z and n_yz -- which no x-range can prune -- over the z-window,
//           2 tests per visited pair instead of 4, one pass instead of two, queries regrouped by their true eps.
// Workspace layout per job (doubles): D columns of npad values in strip order | npad int32 original indices |
// 2 doubles per strip: z of its first and last rank.
// ------------------------------------------------------------------------------------------------
constexpr int STRIP = 64;

__host__ __device__ inline int64_t strip_block_doubles(int D, int64_t npad) { return (int64_t)D * npad + npad / 2 + 2 * (npad / STRIP); }

template <int D>
__global__ void __launch_bounds__(64)
strip_sort_kernel(JobDesc* __restrict__ jobs, int n_jobs, int strips_per_job, double* __restrict__ ws, int64_t block_doubles, int64_t npad)
{
    __shared__ double key[2][STRIP];
    __shared__ int    id[2][STRIP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * 2 + wid;
    const int job = (int)(gw / strips_per_job);
    const int s = (int)(gw - (int64_t)job * strips_per_job);
    if (job >= n_jobs) return;                                       // odd warp count: the last CTA's second warp is idle
    JobDesc& jd = jobs[job];
    const int n = jd.n;
    double* A = ws + (size_t)job * block_doubles;
    if (s == 0 && lane == 0) { jd.strip = A; jd.npad = npad; }
    if ((int64_t)s * STRIP >= npad) return;
    const int base = s * STRIP;
    double* K = key[wid]; int* I = id[wid];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int r = base + h * 32 + lane;
        const int idx = r < n ? (jd.perm ? jd.perm[r] : r) : -1;
        K[h * 32 + lane] = idx >= 0 ? jd.col[0][idx] : INFINITY;       // padding sorts last
        I[h * 32 + lane] = idx;
    }
    __syncwarp();
    // bitonic sort of 64 (key, id) pairs, one compare-exchange per lane and step
    for (int k2 = 2; k2 <= STRIP; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const int i = 2 * j * (lane / j) + (lane % j);
            const int p = i + j;
            const bool up = (i & k2) == 0;
            const double a = K[i], b = K[p];
            const int ia = I[i], ib = I[p];
            const bool sw = up ? (a > b) : (a < b);
            if (sw) { K[i] = b; K[p] = a; I[i] = ib; I[p] = ia; }
            __syncwarp();
        }
    }
    int32_t* SI = reinterpret_cast<int32_t*>(A + (size_t)D * npad);
    double* ED = A + (size_t)D * npad + npad / 2;
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int m = h * 32 + lane;
        const int idx = I[m];
        A[base + m] = K[m];
        #pragma unroll
        for (int c = 1; c < D; c++) A[(size_t)c * npad + base + m] = idx >= 0 ? jd.col[c][idx] : INFINITY;
        SI[base + m] = idx;
    }
    if (lane == 0) {                                                     // z-range of the strip (z-ranks base .. base+cnt-1)
        const int cnt = min(STRIP, n - base);
        double zl = INFINITY, zh = -INFINITY;
        if (cnt > 0) {
            const int i0 = jd.perm ? jd.perm[base] : base, i1 = jd.perm ? jd.perm[base + cnt - 1] : base + cnt - 1;
            zl = jd.col[D - 1][i0]; zh = jd.col[D - 1][i1];
        }
        ED[2 * s] = zl; ED[2 * s + 1] = zh;
    }
}

template <int DZ, int KCAP, bool WEIGHTED>
__global__ void __launch_bounds__(SWEEP_TPB, SWEEP_MIN_CTAS)
ksg_strip_kernel(const JobDesc* __restrict__ jobs, int strips_per_job, int k, const double* __restrict__ psi_tab,
                 double* __restrict__ partial, PointOut po, unsigned long long* __restrict__ visited)
{
    constexpr int D = 2 + DZ, KEY = D - 1, Q = 2, W = SWEEP_TPB / 32;
    static_assert(STRIP == 32 * Q, "a warp owns exactly one strip");
    __shared__ __align__(16) double sbuf[W][2 * D + 1][32];          // step A: [D][STRIP] strip image; step B: [D+1][32] chunk

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bpj = (strips_per_job + W - 1) / W;
    const int job = blockIdx.x / bpj;
    const int s0  = (blockIdx.x - job * bpj) * W + wid;              // this warp's strip
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    if (s0 >= strips_per_job || s0 * STRIP >= n) return;            // warp-uniform; no CTA-level synchronisation below
    const int nstrips = (n + STRIP - 1) / STRIP;
    const int nchunks = (n + 31) / 32;
    const int c0 = s0 * Q;
    const int64_t np = jd.npad;
    const double* __restrict__ A = jd.strip;
    const int32_t* __restrict__ SI = reinterpret_cast<const int32_t*>(A + (size_t)D * np);
    const double* __restrict__ ED = A + (size_t)D * np + np / 2;
    double (*buf)[32] = sbuf[wid];

    // queries = the strip's entries in x-order; padding (x = +inf) sorts last, idle slots shadow the first entry
    double qc[Q][D], qw[Q];
    int qidx[Q]; bool qok[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int r = s0 * STRIP + q * 32 + lane;
        const int idx = SI[r];
        qok[q] = idx >= 0;
        const int rr = qok[q] ? r : s0 * STRIP;
        qidx[q] = SI[rr];
        #pragma unroll
        for (int c = 0; c < D; c++) qc[q][c] = A[(size_t)c * np + rr];
        qw[q] = 0.0;
        if (WEIGHTED) qw[q] = jd.w[qidx[q]];
    }

    // ---------------- step A: walk over the x-sorted strips ----------------
    // The warp visits strips collectively (own strip first, then outwards while some lane still needs the next one);
    // each visited strip is staged in shared memory once, and every lane then searches ITS x-range in it privately
    // (binary search + short scan in shared memory), so the work per strip is ~x-range, not STRIP, per query.
    double (*sst)[STRIP] = reinterpret_cast<double (*)[STRIP]>(&sbuf[wid][0][0]);      // [D][STRIP] view (aliases buf)
    auto stage_strip = [&](int s) {
        const int base = s * STRIP;
        __syncwarp();
        #pragma unroll
        for (int h = 0; h < 2; h++)
            #pragma unroll
            for (int c = 0; c < D; c++) sst[c][h * 32 + lane] = A[(size_t)c * np + base + h * 32 + lane];
        __syncwarp();
    };
    unsigned long long nvis = 0;
    double L[Q][KCAP];
    #pragma unroll
    for (int q = 0; q < Q; q++)
        #pragma unroll
        for (int t = 0; t < KCAP; t++) L[q][t] = (t < KCAP - 1 - k) ? -INFINITY : INFINITY;

    auto knn_in_strip = [&](int s, bool own) {
        const int cnt = min(STRIP, n - s * STRIP);
        const double zedge = own ? 0.0 : (s < s0 ? ED[2 * s + 1] : ED[2 * s]);
        #pragma unroll
        for (int q = 0; q < Q; q++) {
            const double x = qc[q][0];
            auto cand = [&](int m) {
                ++nvis;
                double dd = fabs(x - sst[0][m]);
                bool l = dd < L[q][KCAP - 1];
                #pragma unroll
                for (int c = 1; c < D; c++) { const double a = fabs(qc[q][c] - sst[c][m]); l = l & (a < L[q][KCAP - 1]); dd = (a > dd) ? a : dd; }
                if (l) topk_insert<KCAP>(L[q], dd);
            };
            if (qok[q] && (own || fabs(qc[q][KEY] - zedge) < L[q][KCAP - 1])) {
                int ea = 0, eb = 0;
                if (own) {      // seed with the x-neighbours (close in x, and in z by construction), then the rest of the strip
                    const int p0 = q * 32 + lane;
                    ea = max(0, p0 - (KCAP + 1)); eb = min(cnt, p0 + KCAP + 2);
                    for (int m = ea; m < eb; m++) cand(m);
                }
                int lo = 0, hi = cnt;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (x - sst[0][mid] >= L[q][KCAP - 1]) lo = mid + 1; else hi = mid; }
                for (int m = lo; m < cnt; m++) {
                    if (m >= ea && m < eb) { m = eb - 1; continue; }
                    if (sst[0][m] - x >= L[q][KCAP - 1]) break;
                    cand(m);
                }
            }
        }
    };
    // collective strip order shared by both walks: own strip, then outwards, alternating sides while both are needed
    auto walk = [&](auto need, auto work) {
        stage_strip(s0);
        work(s0, true);
        int sl = s0 - 1, sr = s0 + 1;
        bool toggle = false;
        while (true) {
            bool al = false, ar = false;
            if (sl >= 0)      al = __any_sync(0xffffffffu, need(ED[2 * sl + 1]));
            if (sr < nstrips) ar = __any_sync(0xffffffffu, need(ED[2 * sr]));
            if (!al && !ar) break;
            const bool left = al && (!ar || toggle);
            toggle = !toggle;
            const int s = left ? sl-- : sr++;
            stage_strip(s);
            work(s, false);
        }
    };
    walk([&](double zedge) { bool f = false;
                             #pragma unroll
                             for (int q = 0; q < Q; q++) f = f | (qok[q] & (fabs(qc[q][KEY] - zedge) < L[q][KCAP - 1]));
                             return f; },
         knn_in_strip);

    double eps[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) eps[q] = L[q][KCAP - 1];             // +inf when n <= k, like cKDTree's padding

    // n_xz, n_xyz over the same structure with the closed ball
    int nxz[Q], nxyz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) nxz[q] = nxyz[q] = 0;
    auto cnt_in_strip = [&](int s, bool own) {
        const int cnt = min(STRIP, n - s * STRIP);
        const double zedge = own ? 0.0 : (s < s0 ? ED[2 * s + 1] : ED[2 * s]);
        #pragma unroll
        for (int q = 0; q < Q; q++) {
            const double x = qc[q][0], e = eps[q];
            if (qok[q] && (own || fabs(qc[q][KEY] - zedge) <= e)) {
                int lo = 0, hi = cnt;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (x - sst[0][mid] > e) lo = mid + 1; else hi = mid; }
                for (int m = lo; m < cnt; m++) {
                    if (sst[0][m] - x > e) break;
                    ++nvis;
                    bool pz = true;
                    #pragma unroll
                    for (int c = 0; c < DZ; c++) pz = pz & (fabs(qc[q][2 + c] - sst[2 + c][m]) <= e);
                    const bool py = fabs(qc[q][1] - sst[1][m]) <= e;
                    nxz[q] += pz ? 1 : 0;
                    nxyz[q] += (pz & py) ? 1 : 0;
                }
            }
        }
    };
    walk([&](double zedge) { bool f = false;
                             #pragma unroll
                             for (int q = 0; q < Q; q++) f = f | (qok[q] & (fabs(qc[q][KEY] - zedge) <= eps[q]));
                             return f; },
         cnt_in_strip);

    // ---------------- regroup the warp's queries by eps (slot 1 = the 32 sparsest), as in ksg_sweep_kernel ----------------
    {
        double* kb = &buf[0][0];
        double key[Q]; int pos[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) key[q] = qok[q] ? eps[q] : INFINITY;
        __syncwarp();
        #pragma unroll
        for (int q = 0; q < Q; q++) kb[q * 32 + lane] = key[q];
        __syncwarp();
        #pragma unroll
        for (int q = 0; q < Q; q++) pos[q] = 0;
        for (int j = 0; j < 32 * Q; j++) {
            const double kj = kb[j];
            #pragma unroll
            for (int q = 0; q < Q; q++) pos[q] += ((kj < key[q]) | ((kj == key[q]) & (j < q * 32 + lane))) ? 1 : 0;
        }
        auto move = [&](double (&v)[Q]) {
            __syncwarp();
            #pragma unroll
            for (int q = 0; q < Q; q++) kb[pos[q]] = v[q];
            __syncwarp();
            #pragma unroll
            for (int q = 0; q < Q; q++) v[q] = kb[q * 32 + lane];
        };
        double t[Q];
        #pragma unroll
        for (int c = 1; c < D; c++) {                                   // x is not needed any more
            #pragma unroll
            for (int q = 0; q < Q; q++) t[q] = qc[q][c];
            move(t);
            #pragma unroll
            for (int q = 0; q < Q; q++) qc[q][c] = t[q];
        }
        move(eps);
        if (WEIGHTED) move(qw);
        #pragma unroll
        for (int q = 0; q < Q; q++) t[q] = (double)(2 * (int64_t)qidx[q] + (qok[q] ? 1 : 0));
        move(t);
        #pragma unroll
        for (int q = 0; q < Q; q++) { const int64_t c = (int64_t)t[q]; qidx[q] = (int)(c >> 1); qok[q] = (c & 1) != 0; }
        #pragma unroll
        for (int q = 0; q < Q; q++) t[q] = (double)(((int64_t)nxz[q] << 24) + nxyz[q]);      // both < 2^24, exact in a double
        move(t);
        #pragma unroll
        for (int q = 0; q < Q; q++) { const int64_t c = (int64_t)t[q]; nxz[q] = (int)(c >> 24); nxyz[q] = (int)(c & 0xffffff); }
        __syncwarp();
    }

    // ---------------- step B: n_z, n_yz (and the weighted sums) over the z-window, collectively ----------------
    int nyz[Q], nz[Q];
    double wyz[Q], wz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) { nyz[q] = nz[q] = 0; wyz[q] = wz[q] = 0.0; }

    auto load_yz = [&](int chunk) {                                   // z-ordered chunk: y and z columns (+ weight) only
        Chunk<D, WEIGHTED> c;
        const int rank = chunk * 32 + lane;
        const bool ok = rank < n;
        const int idx = ok ? (jd.perm ? jd.perm[rank] : rank) : 0;
        c.v[0] = 0.0;
        #pragma unroll
        for (int d = 1; d < D; d++) c.v[d] = ok ? jd.col[d][idx] : __longlong_as_double(0x7ff8000000000000LL);
        c.w = 0.0;
        if (WEIGHTED) c.w = ok ? jd.w[idx] : 0.0;
        return c;
    };
    auto stage = [&](const Chunk<D, WEIGHTED>& c) {
        __syncwarp();
        #pragma unroll
        for (int d = 1; d < D; d++) buf[d][lane] = c.v[d];
        if (WEIGHTED) buf[D][lane] = c.w;
        __syncwarp();
    };
    auto yz_body = [&](auto q_first) {
        constexpr int Q0 = decltype(q_first)::value;
        #pragma unroll 4
        for (int j = 0; j < 32; j++) {
            double cc[D];
            #pragma unroll
            for (int d = 1; d < D; d++) cc[d] = buf[d][j];
            double wj_ = 0.0;
            if (WEIGHTED) wj_ = buf[D][j];
            #pragma unroll
            for (int q = Q0; q < Q; q++) {
                bool pz = true;
                #pragma unroll
                for (int d = 0; d < DZ; d++) pz = pz & (fabs(qc[q][2 + d] - cc[2 + d]) <= eps[q]);
                const bool pyz = pz & (fabs(qc[q][1] - cc[1]) <= eps[q]);
                inc_if(nz[q], pz); inc_if(nyz[q], pyz);
                if (WEIGHTED) { add_if(wz[q], wj_, pz); add_if(wyz[q], wj_, pyz); }
            }
        }
    };
    unsigned long long nvis2 = 0;
    {
        for (int c = c0; c < c0 + Q && c < nchunks; c++) { stage(load_yz(c)); yz_body(std::integral_constant<int, 0>{}); nvis2 += Q; }
        int cl = c0 - 1, cr = c0 + Q;
        bool toggle = false;
        Chunk<D, WEIGHTED> Lc, Rc, cur;
        if (cl >= 0) Lc = load_yz(cl);
        if (cr < nchunks) Rc = load_yz(cr);
        auto need = [&](double znext) { unsigned f = 0u;
                                        #pragma unroll
                                        for (int q = 0; q < Q; q++) f |= (fabs(qc[q][KEY] - znext) <= eps[q]) ? (1u << q) : 0u;
                                        return f; };
        while (true) {
            unsigned al = 0, ar = 0;
            if (cl >= 0)      al = __reduce_or_sync(0xffffffffu, need(__shfl_sync(0xffffffffu, Lc.v[KEY], 31)));
            if (cr < nchunks) ar = __reduce_or_sync(0xffffffffu, need(__shfl_sync(0xffffffffu, Rc.v[KEY], 0)));
            if (!al && !ar) break;
            const bool left = al && (!ar || toggle);
            toggle = !toggle;
            const unsigned act = left ? al : ar;
            if (left) { cur = Lc; --cl; if (cl >= 0) Lc = load_yz(cl); }
            else      { cur = Rc; ++cr; if (cr < nchunks) Rc = load_yz(cr); }
            stage(cur);
            if (act & 1u) { yz_body(std::integral_constant<int, 0>{}); nvis2 += Q; }
            else          { yz_body(std::integral_constant<int, Q - 1>{}); nvis2 += 1; }
        }
    }

    // ---------------- digamma terms + per-warp partial ----------------
    double acc = 0.0;
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        if (qok[q]) {
            const int a = nxyz[q] - 1, b = nxz[q] - 1, c = nyz[q] - 1, d = nz[q] - 1;
            const int64_t o = jd.qoff + qidx[q];
            double term;
            if (WEIGHTED) {
                const double wi = qw[q];
                const double Wyz = wyz[q] - wi, Wz = wz[q] - wi;
                term = wi * psi_int(psi_tab, a) + wi * -psi_int(psi_tab, b) + wi * -digamma_f64(Wyz) + wi * digamma_f64(Wz);
                if (po.wsum) { po.wsum[2 * o] = Wyz; po.wsum[2 * o + 1] = Wz; }
            } else {
                term = psi_int(psi_tab, a) - psi_int(psi_tab, b) - psi_int(psi_tab, c) + psi_int(psi_tab, d);
            }
            acc += term;
            if (po.eps) po.eps[o] = eps[q];
            if (po.cnt) { int32_t* oc = po.cnt + 4 * o; oc[0] = a; oc[1] = b; oc[2] = c; oc[3] = d; }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvis += __shfl_down_sync(0xffffffffu, nvis, o);
    if (lane == 0) {
        partial[(size_t)job * strips_per_job + s0] = acc;
        if (visited) { atomicAdd(visited, nvis); atomicAdd(visited + 1, nvis2 * 32ull * 32ull); }
    }
}

// ------------------------------------------------------------------------------------------------
// Strip kernels, second form (one conditioning column, DZ = 1: RDI and dynamics coupling).  Besides the x-sorted image
// of every z-strip, the job also gets a y-sorted image with the weights' prefix sums, so that ALL four ball counts become
// range queries: a strip whose whole z-range lies inside the ball contributes (upper - lower bound) of a binary search in
// y to n_yz, its size to n_z and prefix-sum differences to the weighted sums; only the (at most two, plus the own) strips
// cut by the ball's z-boundary are examined entry by entry.  No candidate-by-candidate z-sweep remains:
// per query the cost is O(strips in the ball * log STRIP + |(x,z) box|) instead of O(z-window).
// Interior test: z-ranks are contiguous, so if fl|z_i - z_first| <= eps and fl|z_i - z_last| <= eps every entry of the
// strip satisfies the reference's test fl|z_i - z_j| <= eps (rounding is monotone on either side of z_i).
// Layout per job (doubles): image X = 3 columns (x | y | z, x-sorted) + npad/2 (original indices) ;
// image Y = y | z (y-sorted) [+ w | exclusive prefix of w, y-order] | z in rank order [+ exclusive prefix of w, rank order] ;
// 2 per strip: z of first / last rank ; [+ 1 per strip: sum w].
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int64_t strip2_block_doubles(int64_t npad, bool weighted) {
    return 3 * npad + npad / 2 + (weighted ? 6 : 3) * npad + (weighted ? 3 : 2) * (npad / STRIP);
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(64)
strip_sort2_kernel(JobDesc* __restrict__ jobs, int n_jobs, int strips_per_job, double* __restrict__ ws, int64_t block_doubles, int64_t npad)
{
    __shared__ double key[2][STRIP];
    __shared__ int    id[2][STRIP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * 2 + wid;
    const int job = (int)(gw / strips_per_job);
    const int s = (int)(gw - (int64_t)job * strips_per_job);
    if (job >= n_jobs) return;
    JobDesc& jd = jobs[job];
    const int n = jd.n;
    double* A = ws + (size_t)job * block_doubles;
    if (s == 0 && lane == 0) { jd.strip = A; jd.npad = npad; }
    const int base = s * STRIP;
    double* K = key[wid]; int* I = id[wid];
    int32_t* SI = reinterpret_cast<int32_t*>(A + 3 * npad);
    double* Y = A + 3 * npad + npad / 2;                              // image Y
    const int ycols = WEIGHTED ? 6 : 3;
    const int zzc = WEIGHTED ? 4 : 2;                                // column of z in rank (= z) order
    double* ED = Y + (size_t)ycols * npad;
    double* TOT = ED + 2 * (npad / STRIP);

    auto bitonic = [&]() {
        for (int k2 = 2; k2 <= STRIP; k2 <<= 1)
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                const int i = 2 * j * (lane / j) + (lane % j);
                const int p = i + j;
                const bool up = (i & k2) == 0;
                const double a = K[i], b = K[p];
                const int ia = I[i], ib = I[p];
                const bool sw = up ? (a > b) : (a < b);
                if (sw) { K[i] = b; K[p] = a; I[i] = ib; I[p] = ia; }
                __syncwarp();
            }
    };
    int myidx[2];
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int r = base + h * 32 + lane;
        myidx[h] = r < n ? (jd.perm ? jd.perm[r] : r) : -1;
    }
    // ---- image X: sort by x
    #pragma unroll
    for (int h = 0; h < 2; h++) { K[h * 32 + lane] = myidx[h] >= 0 ? jd.col[0][myidx[h]] : INFINITY; I[h * 32 + lane] = myidx[h]; }
    __syncwarp();
    bitonic();
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int m = h * 32 + lane;
        const int idx = I[m];
        A[base + m] = K[m];
        A[npad + base + m]     = idx >= 0 ? jd.col[1][idx] : INFINITY;
        A[2 * npad + base + m] = idx >= 0 ? jd.col[2][idx] : INFINITY;
        SI[base + m] = idx;
    }
    __syncwarp();
    // ---- image Y: sort by y
    #pragma unroll
    for (int h = 0; h < 2; h++) { K[h * 32 + lane] = myidx[h] >= 0 ? jd.col[1][myidx[h]] : INFINITY; I[h * 32 + lane] = myidx[h]; }
    __syncwarp();
    bitonic();
    double wv[2] = {0.0, 0.0};
    #pragma unroll
    for (int h = 0; h < 2; h++) {
        const int m = h * 32 + lane;
        const int idx = I[m];
        Y[base + m] = K[m];
        Y[npad + base + m] = idx >= 0 ? jd.col[2][idx] : INFINITY;
        if (WEIGHTED) { wv[h] = idx >= 0 ? jd.w[idx] : 0.0; Y[2 * npad + base + m] = wv[h]; }
    }
    if (WEIGHTED) {                                                  // exclusive prefix of w over the strip (fixed order)
        double inc0 = wv[0], inc1 = wv[1];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t0 = __shfl_up_sync(0xffffffffu, inc0, o), t1 = __shfl_up_sync(0xffffffffu, inc1, o);
            if (lane >= o) { inc0 += t0; inc1 += t1; }
        }
        const double half = __shfl_sync(0xffffffffu, inc0, 31);
        Y[3 * npad + base + lane]      = inc0 - wv[0];
        Y[3 * npad + base + 32 + lane] = half + (inc1 - wv[1]);
        if (lane == 31) TOT[s] = half + inc1;
    }
    {   // z (and the weights' prefix) in rank order: the ball cuts an edge strip at a rank found by binary search
        double wr[2] = {0.0, 0.0};
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            const int idx = myidx[h];
            Y[(size_t)zzc * npad + base + h * 32 + lane] = idx >= 0 ? jd.col[2][idx] : INFINITY;
            if (WEIGHTED) wr[h] = idx >= 0 ? jd.w[idx] : 0.0;
        }
        if (WEIGHTED) {
            double inc0 = wr[0], inc1 = wr[1];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t0 = __shfl_up_sync(0xffffffffu, inc0, o), t1 = __shfl_up_sync(0xffffffffu, inc1, o);
                if (lane >= o) { inc0 += t0; inc1 += t1; }
            }
            const double half = __shfl_sync(0xffffffffu, inc0, 31);
            Y[(size_t)5 * npad + base + lane]      = inc0 - wr[0];
            Y[(size_t)5 * npad + base + 32 + lane] = half + (inc1 - wr[1]);
        }
    }
    if (lane == 0) {
        const int cnt = min(STRIP, n - base);
        double zl = INFINITY, zh = -INFINITY;
        if (cnt > 0) {
            const int i0 = jd.perm ? jd.perm[base] : base, i1 = jd.perm ? jd.perm[base + cnt - 1] : base + cnt - 1;
            zl = jd.col[2][i0]; zh = jd.col[2][i1];
        }
        ED[2 * s] = zl; ED[2 * s + 1] = zh;
    }
}

template <int KCAP, bool WEIGHTED>
__global__ void __launch_bounds__(SWEEP_TPB, SWEEP_MIN_CTAS)
ksg_strip2_kernel(const JobDesc* __restrict__ jobs, int strips_per_job, int k, const double* __restrict__ psi_tab,
                  double* __restrict__ partial, PointOut po, unsigned long long* __restrict__ visited)
{
    constexpr int D = 3, Q = 2, W = SWEEP_TPB / 32;
    constexpr int YC = WEIGHTED ? 6 : 3;
    constexpr int ZZ = WEIGHTED ? 4 : 2;                              // image-Y column: z in rank order
    __shared__ __align__(16) double sbuf[W][6][STRIP];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bpj = (strips_per_job + W - 1) / W;
    const int job = blockIdx.x / bpj;
    const int s0  = (blockIdx.x - job * bpj) * W + wid;
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    if (s0 >= strips_per_job || s0 * STRIP >= n) return;
    const int nstrips = (n + STRIP - 1) / STRIP;
    const int64_t np = jd.npad;
    const double* __restrict__ A = jd.strip;
    const int32_t* __restrict__ SI = reinterpret_cast<const int32_t*>(A + 3 * np);
    const double* __restrict__ Y = A + 3 * np + np / 2;
    const double* __restrict__ ED = Y + (size_t)YC * np;
    const double* __restrict__ TOT = ED + 2 * (np / STRIP);
    double (*sst)[STRIP] = sbuf[wid];

    double qc[Q][D], qw[Q];
    int qidx[Q]; bool qok[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int r = s0 * STRIP + q * 32 + lane;
        qok[q] = SI[r] >= 0;
        const int rr = qok[q] ? r : s0 * STRIP;
        qidx[q] = SI[rr];
        #pragma unroll
        for (int c = 0; c < D; c++) qc[q][c] = A[(size_t)c * np + rr];
        qw[q] = 0.0;
        if (WEIGHTED) qw[q] = jd.w[qidx[q]];
    }

    auto stage = [&](const double* img, int ncols, int s) {
        const int base = s * STRIP;
        __syncwarp();
        for (int c = 0; c < ncols; c++) {
            sst[c][lane]      = img[(size_t)c * np + base + lane];
            sst[c][32 + lane] = img[(size_t)c * np + base + 32 + lane];
        }
        __syncwarp();
    };
    auto walk = [&](const double* img, int ncols, auto need, auto work) {
        stage(img, ncols, s0);
        work(s0, true);
        int sl = s0 - 1, sr = s0 + 1;
        bool toggle = false;
        while (true) {
            bool al = false, ar = false;
            if (sl >= 0)      al = __any_sync(0xffffffffu, need(ED[2 * sl + 1]));
            if (sr < nstrips) ar = __any_sync(0xffffffffu, need(ED[2 * sr]));
            if (!al && !ar) break;
            const bool left = al && (!ar || toggle);
            toggle = !toggle;
            const int s = left ? sl-- : sr++;
            stage(img, ncols, s);
            work(s, false);
        }
    };

    // ---------------- k-th neighbour distance (image X) ----------------
    unsigned long long nvis = 0;
    double L[Q][KCAP];
    #pragma unroll
    for (int q = 0; q < Q; q++)
        #pragma unroll
        for (int t = 0; t < KCAP; t++) L[q][t] = (t < KCAP - 1 - k) ? -INFINITY : INFINITY;
    auto knn_in_strip = [&](int s, bool own) {
        const int cnt = min(STRIP, n - s * STRIP);
        const double zedge = own ? 0.0 : (s < s0 ? ED[2 * s + 1] : ED[2 * s]);
        #pragma unroll
        for (int q = 0; q < Q; q++) {
            const double x = qc[q][0];
            auto cand = [&](int m) {
                ++nvis;
                double dd = fabs(x - sst[0][m]);
                bool l = dd < L[q][KCAP - 1];
                #pragma unroll
                for (int c = 1; c < D; c++) { const double a = fabs(qc[q][c] - sst[c][m]); l = l & (a < L[q][KCAP - 1]); dd = (a > dd) ? a : dd; }
                if (l) topk_insert<KCAP>(L[q], dd);
            };
            if (qok[q] && (own || fabs(qc[q][2] - zedge) < L[q][KCAP - 1])) {
                int ea = 0, eb = 0;
                if (own) {
                    const int p0 = q * 32 + lane;
                    ea = max(0, p0 - (KCAP + 1)); eb = min(cnt, p0 + KCAP + 2);
                    for (int m = ea; m < eb; m++) cand(m);
                }
                int lo = 0, hi = cnt;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (x - sst[0][mid] >= L[q][KCAP - 1]) lo = mid + 1; else hi = mid; }
                for (int m = lo; m < cnt; m++) {
                    if (m >= ea && m < eb) { m = eb - 1; continue; }
                    if (sst[0][m] - x >= L[q][KCAP - 1]) break;
                    cand(m);
                }
            }
        }
    };
    walk(A, 3, [&](double zedge) { bool f = false;
                                   #pragma unroll
                                   for (int q = 0; q < Q; q++) f = f | (qok[q] & (fabs(qc[q][2] - zedge) < L[q][KCAP - 1]));
                                   return f; },
         knn_in_strip);
    double eps[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) eps[q] = L[q][KCAP - 1];

    auto need_eps = [&](double zedge) { bool f = false;
                                        #pragma unroll
                                        for (int q = 0; q < Q; q++) f = f | (qok[q] & (fabs(qc[q][2] - zedge) <= eps[q]));
                                        return f; };

    // ---------------- n_xz, n_xyz (image X, closed ball) ----------------
    int nxz[Q], nxyz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) nxz[q] = nxyz[q] = 0;
    walk(A, 3, need_eps, [&](int s, bool own) {
        const int cnt = min(STRIP, n - s * STRIP);
        const double zedge = own ? 0.0 : (s < s0 ? ED[2 * s + 1] : ED[2 * s]);
        #pragma unroll
        for (int q = 0; q < Q; q++) {
            const double x = qc[q][0], e = eps[q];
            if (qok[q] && (own || fabs(qc[q][2] - zedge) <= e)) {
                int lo = 0, hi = cnt;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (x - sst[0][mid] > e) lo = mid + 1; else hi = mid; }
                for (int m = lo; m < cnt; m++) {
                    if (sst[0][m] - x > e) break;
                    ++nvis;
                    const bool pz = fabs(qc[q][2] - sst[2][m]) <= e;
                    const bool py = fabs(qc[q][1] - sst[1][m]) <= e;
                    nxz[q] += pz ? 1 : 0;
                    nxyz[q] += (pz & py) ? 1 : 0;
                }
            }
        }
    });

    // ---------------- n_z, n_yz and the weighted sums (image Y) ----------------
    int nyz[Q], nz[Q];
    double wyz[Q], wz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) { nyz[q] = nz[q] = 0; wyz[q] = wz[q] = 0.0; }
    walk(Y, YC, need_eps, [&](int s, bool own) {
        const int cnt = min(STRIP, n - s * STRIP);
        const double zlo = ED[2 * s], zhi = ED[2 * s + 1];
        double tot = 0.0;
        if (WEIGHTED) tot = TOT[s];
        #pragma unroll
        for (int q = 0; q < Q; q++) {
            const double yq = qc[q][1], zq = qc[q][2], e = eps[q];
            const bool in_lo = fabs(zq - zlo) <= e, in_hi = fabs(zq - zhi) <= e;
            // the ball's z-range meets this strip iff its nearer end is inside (for the own strip: always)
            const bool touches = own || (s < s0 ? in_hi : in_lo);
            if (qok[q] && touches) {
                int lb = 0, hi = cnt;                                   // y-range [lb, ub): first entry with y_q - y <= e ...
                while (lb < hi) { const int mid = (lb + hi) >> 1; if (yq - sst[0][mid] > e) lb = mid + 1; else hi = mid; }
                int ub = lb; hi = cnt;                                  // ... up to the first entry with y - y_q > e
                while (ub < hi) { const int mid = (ub + hi) >> 1; if (sst[0][mid] - yq > e) hi = mid; else ub = mid + 1; }
                if (in_lo && in_hi) {                                   // whole strip inside the ball in z
                    nz[q] += cnt; nyz[q] += ub - lb;
                    if (WEIGHTED) {
                        wz[q] += tot;
                        // exclusive prefix P[m] = sum_{m' < m} w; P[cnt] = tot
                        const double pu = (ub < cnt) ? sst[3][ub] : tot;
                        const double pl = (lb < cnt) ? sst[3][lb] : tot;
                        wyz[q] += pu - pl;
                    }
                    nvis += 12;
                } else {        // cut by the ball's boundary: n_z from the rank order (z ascending), n_yz over the y-range only
                    int zl = 0; hi = cnt;
                    while (zl < hi) { const int mid = (zl + hi) >> 1; if (zq - sst[ZZ][mid] > e) zl = mid + 1; else hi = mid; }
                    int zu = zl; hi = cnt;
                    while (zu < hi) { const int mid = (zu + hi) >> 1; if (sst[ZZ][mid] - zq > e) hi = mid; else zu = mid + 1; }
                    nz[q] += zu - zl;
                    if (WEIGHTED) {
                        const double pu = (zu < cnt) ? sst[5][zu] : tot, pl = (zl < cnt) ? sst[5][zl] : tot;
                        wz[q] += pu - pl;
                    }
                    for (int m = lb; m < ub; m++) {
                        const bool pyz = fabs(zq - sst[1][m]) <= e;
                        nyz[q] += pyz ? 1 : 0;
                        if (WEIGHTED) wyz[q] += pyz ? sst[2][m] : 0.0;
                    }
                    nvis += 12 + (ub - lb);
                }
            }
        }
    });

    // ---------------- digamma terms + per-warp partial ----------------
    double acc = 0.0;
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        if (qok[q]) {
            const int a = nxyz[q] - 1, b = nxz[q] - 1, c = nyz[q] - 1, d = nz[q] - 1;
            const int64_t o = jd.qoff + qidx[q];
            double term;
            if (WEIGHTED) {
                const double wi = qw[q];
                const double Wyz = wyz[q] - wi, Wz = wz[q] - wi;
                term = wi * psi_int(psi_tab, a) + wi * -psi_int(psi_tab, b) + wi * -digamma_f64(Wyz) + wi * digamma_f64(Wz);
                if (po.wsum) { po.wsum[2 * o] = Wyz; po.wsum[2 * o + 1] = Wz; }
            } else {
                term = psi_int(psi_tab, a) - psi_int(psi_tab, b) - psi_int(psi_tab, c) + psi_int(psi_tab, d);
            }
            acc += term;
            if (po.eps) po.eps[o] = eps[q];
            if (po.cnt) { int32_t* oc = po.cnt + 4 * o; oc[0] = a; oc[1] = b; oc[2] = c; oc[3] = d; }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvis += __shfl_down_sync(0xffffffffu, nvis, o);
    if (lane == 0) {
        partial[(size_t)job * strips_per_job + s0] = acc;
        if (visited) atomicAdd(visited, nvis);
    }
}

// FP32 box test of the KDE pre-filter: set `bit` unless some coordinate difference is certainly above T (chained unordered
// compares and one predicated OR: DP + 1 issue slots)
template <int NP>
__device__ __forceinline__ void box_mark32(const float (&df)[NP], float T, unsigned bit, unsigned& mask)
{
    static_assert(NP >= 1 && NP <= 4, "1..4 density coordinates on the sweep");
    if (NP == 1)
        asm("{\n .reg .pred p;\n .reg .f32 a0;\n abs.f32 a0,%1;\n setp.leu.f32 p, a0, %2;\n @p or.b32 %0,%0,%3;\n}"
            : "+r"(mask) : "f"(df[0]), "f"(T), "r"(bit));
    else if (NP == 2)
        asm("{\n .reg .pred p;\n .reg .f32 a0,a1;\n abs.f32 a0,%1; abs.f32 a1,%2;\n"
            " setp.leu.f32 p, a0, %3;\n setp.leu.and.f32 p, a1, %3, p;\n @p or.b32 %0,%0,%4;\n}"
            : "+r"(mask) : "f"(df[0]), "f"(df[1]), "f"(T), "r"(bit));
    else if (NP == 3)
        asm("{\n .reg .pred p;\n .reg .f32 a0,a1,a2;\n abs.f32 a0,%1; abs.f32 a1,%2; abs.f32 a2,%3;\n"
            " setp.leu.f32 p, a0, %4;\n setp.leu.and.f32 p, a1, %4, p;\n setp.leu.and.f32 p, a2, %4, p;\n @p or.b32 %0,%0,%5;\n}"
            : "+r"(mask) : "f"(df[0]), "f"(df[1]), "f"(df[NP > 2 ? 2 : 0]), "f"(T), "r"(bit));
    else
        asm("{\n .reg .pred p;\n .reg .f32 a0,a1,a2,a3;\n abs.f32 a0,%1; abs.f32 a1,%2; abs.f32 a2,%3; abs.f32 a3,%4;\n"
            " setp.leu.f32 p, a0, %5;\n setp.leu.and.f32 p, a1, %5, p;\n setp.leu.and.f32 p, a2, %5, p;\n setp.leu.and.f32 p, a3, %5, p;\n"
            " @p or.b32 %0,%0,%6;\n}"
            : "+r"(mask) : "f"(df[0]), "f"(df[1]), "f"(df[NP > 2 ? 2 : 0]), "f"(df[NP > 3 ? 3 : 0]), "f"(T), "r"(bit));
}

// Gaussian KDE inverse densities, sorted sweep: terms with |P_i - P_j|^2 * log2(e)/(2 bw^2) > KDE_CUT contribute less
// than 2^-KDE_CUT (= 2^-44) each against a sum that is >= 1 (the self term), so
//   * a side of the sweep is abandoned once the nearest unvisited candidate is that far in the key coordinate for every
//     query of the warp, and
//   * inside the window only candidates that can be that close in EVERY density coordinate reach the exponential: the scan
//     is an FP32 box pre-filter (|d_c| <= sqrt(KDE_CUT), bound widened by the FP32 rounding of the scaled coordinates; one
//     packed subtract + one FSETP per coordinate), and the marked candidates -- ~6 % of the window at bw = 0.01 on z-scored
//     data -- get the exact FP64 squared distance, F2F and MUFU.EX2.  ncu on the first form (every visited pair through the
//     exponential) showed the XU pipe (F2F + MUFU) 86 % busy: the kernel was bound by conversions of negligible terms.
constexpr double KDE_CUT = 44.0;     // N * 2^-44 < 2e-9 relative for N <= 2^15; the estimators' budget is 1e-5 absolute

template <int DP, int Q>
__global__ void __launch_bounds__(SWEEP_TPB)
kde_sweep_kernel(const JobDesc* __restrict__ jobs, int warps_per_job, const int* __restrict__ dims, double scale,
                 double* __restrict__ u_out, unsigned long long* __restrict__ visited)
{
    constexpr int W = SWEEP_TPB / 32;
    constexpr int KEY = DP - 1;                                   // dims[] lists the key column last
    constexpr int DPA = (DP + 1) & ~1;                            // float rows padded to whole pairs
    __shared__ __align__(16) double sbuf[W][32][DPA];             // scaled candidates, FP64 (exact distance of marked candidates)
    __shared__ __align__(16) float  sflt[W][32][DPA];             // the same in FP32 (box pre-filter)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bpj = (warps_per_job + W - 1) / W;
    const int job = blockIdx.x / bpj;
    const int wj  = (blockIdx.x - job * bpj) * W + wid;
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    const int w0 = wj * 32 * Q;
    if (wj >= warps_per_job || w0 >= n) return;
    const int c0 = w0 / 32, nchunks = (n + 31) / 32;
    double (*buf)[DPA] = sbuf[wid];
    float  (*bfl)[DPA] = sflt[wid];
    const double* cp[DP];
    #pragma unroll
    for (int d = 0; d < DP; d++) cp[d] = jd.col[dims[d]];
    const double rcut = sqrt(KDE_CUT);                            // in scaled units
    // FP32 bound of the box filter: the scaled FP64 coordinates are rounded to FP32 (<= 2^-24 relative each) and subtracted
    // in FP32 (<= 2^-24 of the difference); absmax is the job's largest |raw coordinate| (job_absmax_kernel).  A non-finite
    // bound (huge data) marks everything.
    const float tcut = __double2float_ru(rcut * (1.0 + 0x1p-20) + fabs(jd.absmax * scale) * 0x1p-21 + 1e-30);

    auto load = [&](int chunk, double (&v)[DP]) {
        const int rank = chunk * 32 + lane;
        const bool ok = rank < n;
        const int idx = ok ? (jd.perm ? jd.perm[rank] : rank) : 0;
        #pragma unroll
        for (int d = 0; d < DP; d++) v[d] = ok ? cp[d][idx] * scale : INFINITY;     // never inside the box; exp2(-inf) = 0
    };
    auto stage = [&](const double (&v)[DP]) {
        __syncwarp();
        #pragma unroll
        for (int d = 0; d < DPA; d += 2) {
            double2 t; t.x = v[d]; t.y = (d + 1 < DP) ? v[d + 1 < DP ? d + 1 : d] : 0.0;
            *reinterpret_cast<double2*>(&buf[lane][d]) = t;
            float2 f; f.x = __double2float_rn(t.x); f.y = __double2float_rn(t.y);
            *reinterpret_cast<float2*>(&bfl[lane][d]) = f;
        }
        __syncwarp();
    };

    double qc[Q][DP]; int qidx[Q]; bool qok[Q];
    unsigned long long qf2[Q][DPA / 2];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int rank = w0 + q * 32 + lane;
        qok[q] = rank < n;
        const int rr = qok[q] ? rank : n - 1;
        qidx[q] = jd.perm ? jd.perm[rr] : rr;
        #pragma unroll
        for (int d = 0; d < DP; d++) qc[q][d] = cp[d][qidx[q]] * scale;
        #pragma unroll
        for (int d = 0; d < DPA; d += 2) {
            const float a = __double2float_rn(qc[q][d]), b = (d + 1 < DP) ? __double2float_rn(qc[q][d + 1 < DP ? d + 1 : d]) : 0.f;
            asm("mov.b64 %0, {%1,%2};" : "=l"(qf2[q][d / 2]) : "f"(a), "f"(b));
        }
    }
    double S[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) S[q] = 0.0;

    auto kde_chunk = [&]() {
        unsigned mask[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) mask[q] = 0u;
        #pragma unroll 1
        for (int jb = 0; jb < 32; jb += 8) {
            unsigned mb[Q];
            #pragma unroll
            for (int q = 0; q < Q; q++) mb[q] = 0u;
            #pragma unroll
            for (int t = 0; t < 8; t++) {
                const int j = jb + t;
                unsigned long long cf2[DPA / 2];
                if constexpr (DPA == 2) { cf2[0] = *reinterpret_cast<const unsigned long long*>(&bfl[j][0]); }
                else { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&bfl[j][0]); cf2[0] = v.x; cf2[DPA / 2 - 1] = v.y; }
                #pragma unroll
                for (int q = 0; q < Q; q++) {
                    float df[DPA];
                    #pragma unroll
                    for (int d = 0; d < DPA; d += 2) {
                        unsigned long long dd;
                        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dd) : "l"(qf2[q][d / 2]), "l"(cf2[d / 2]));
                        asm("mov.b64 {%0,%1}, %2;" : "=f"(df[d]), "=f"(df[d + 1]) : "l"(dd));
                    }
                    float dn[DP];
                    #pragma unroll
                    for (int d = 0; d < DP; d++) dn[d] = df[d];
                    box_mark32<DP>(dn, tcut, 1u << t, mb[q]);      // unordered-or-less-equal: a NaN difference stays marked
                }
            }
            #pragma unroll
            for (int q = 0; q < Q; q++) mask[q] |= mb[q] << jb;
        }
        float s32[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) s32[q] = 0.f;
        unsigned any = 0u;
        #pragma unroll
        for (int q = 0; q < Q; q++) any |= mask[q];
        while (any) {                                              // the next marked candidate of every slot per trip
            any = 0u;
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                const bool has = mask[q] != 0u;
                const int j = has ? __ffs(mask[q]) - 1 : 0;
                mask[q] &= mask[q] - 1;
                any |= mask[q];
                double d2 = 0.0;
                #pragma unroll
                for (int d = 0; d < DP; d++) { const double df = qc[q][d] - buf[j][d]; d2 = fma(df, df, d2); }
                float e;
                const float a = has ? -(float)d2 : -INFINITY;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
                s32[q] += e;
            }
        }
        #pragma unroll
        for (int q = 0; q < Q; q++) S[q] += (double)s32[q];
    };

    unsigned long long nvis = 0;
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        if (c0 + q < nchunks) {
            double v[DP];
            #pragma unroll
            for (int d = 0; d < DP; d++) v[d] = qok[q] ? qc[q][d] : INFINITY;
            stage(v); kde_chunk(); nvis++;
        }
    }
    int cl = c0 - 1, cr = c0 + Q;
    double Lv[DP], Rv[DP];
    if (cl >= 0) load(cl, Lv);
    if (cr < nchunks) load(cr, Rv);
    while (true) {
        bool nl = false, nr = false;
        if (cl >= 0) {
            const double zl = __shfl_sync(0xffffffffu, Lv[KEY], 31);
            #pragma unroll
            for (int q = 0; q < Q; q++) nl = nl | (fabs(qc[q][KEY] - zl) <= rcut);
            nl = __any_sync(0xffffffffu, nl);
        }
        if (cr < nchunks) {
            const double zr = __shfl_sync(0xffffffffu, Rv[KEY], 0);
            #pragma unroll
            for (int q = 0; q < Q; q++) nr = nr | (fabs(qc[q][KEY] - zr) <= rcut);
            nr = __any_sync(0xffffffffu, nr);
        }
        if (!nl && !nr) break;
        if (nl) { stage(Lv); --cl; if (cl >= 0) load(cl, Lv); kde_chunk(); nvis++; }
        if (nr) { stage(Rv); ++cr; if (cr < nchunks) load(cr, Rv); kde_chunk(); nvis++; }
    }
    #pragma unroll
    for (int q = 0; q < Q; q++) if (qok[q]) u_out[jd.qoff + qidx[q]] = 1.0 / S[q];
    if (lane == 0 && visited) atomicAdd(visited + 2, nvis * 32ull * 32ull * Q);
}

// copy selected columns of a [n_cols][stride] block into a contiguous [n_sel][stride] block (sort keys)
__global__ void __launch_bounds__(TPB)
select_columns_kernel(const double* __restrict__ cols, const int* __restrict__ sel, int64_t stride, double* __restrict__ out)
{
    const double* src = cols + (size_t)sel[blockIdx.y] * stride;
    double* dst = out + (size_t)blockIdx.y * stride;
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < stride; i += (int64_t)gridDim.x * TPB) dst[i] = src[i];
}

__global__ void iota_segments_kernel(int32_t* __restrict__ v, int64_t stride, int n_seg) {
    const int64_t total = stride * n_seg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        v[i] = (int32_t)(i % stride);
}

// rank[s][perm[s][p]] = p for p < len(s): the inverse of an argsort, as float (exact: p < 2^22)
__global__ void rank_from_perm_kernel(const int32_t* __restrict__ perm, const int* __restrict__ seg_begin, const int* __restrict__ seg_end,
                                      int64_t stride, int n_seg, float* __restrict__ rank)
{
    const int64_t total = stride * n_seg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / stride);
        const int p = (int)(i - (int64_t)s * stride);
        if (p < seg_end[s] - seg_begin[s]) rank[(size_t)s * stride + perm[i]] = (float)p;
    }
}

// wcum[job][r] = w[perm[0]] + ... + w[perm[r-1]]  (r = 0..n): prefix sums of a job's weights in sweep order; one CTA per job,
// tiles of 256 ranks scanned in order (deterministic)
__global__ void __launch_bounds__(256)
weight_prefix_kernel(JobDesc* __restrict__ jobs, double* __restrict__ wcum, int64_t stride)
{
    __shared__ double wtot[8];
    __shared__ double carry;
    JobDesc& jd = jobs[blockIdx.x];
    const int n = jd.n;
    double* out = wcum + (size_t)blockIdx.x * stride;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { carry = 0.0; out[0] = 0.0; jd.wcum = out; }
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int r = base + tid;
        double v = (r < n) ? jd.w[jd.perm ? jd.perm[r] : r] : 0.0;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) wtot[wid] = v;
        __syncthreads();
        double off = carry;
        for (int w = 0; w < wid; w++) off += wtot[w];
        if (r < n) out[r + 1] = off + v;
        __syncthreads();
        if (tid == 255) carry = off + v;
        __syncthreads();
    }
}

// out[job] = (sum of that job's CTA partials, in CTA order) / n      (np.mean at :109,:173,:403)
__global__ void finalize_mean_kernel(const JobDesc* __restrict__ jobs, const double* __restrict__ partial,
                                     int blocks_per_job, int n_jobs, double* __restrict__ out)
{
    const int job = blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= n_jobs) return;
    const int n = jobs[job].n;
    double s = 0.0;
    for (int b = 0; b < blocks_per_job; b++) s += partial[(size_t)job * blocks_per_job + b];
    out[job] = (n > 0) ? s / (double)n : NAN;
}

// ------------------------------------------------------------------------------------------------
// Generic kernel: any dx,dy,dz (D <= 16), k+1 <= 64, Chebyshev or Euclidean, optional weights.
// One thread per query point, candidates read through L1/L2 (all lanes hit the same address).
// Writes eps / counts / weighted sums per point; a finish kernel turns them into the estimate.
//   euclid: scipy's p=2 convention -- d2 = sum diff^2 (no FMA contraction), r = sqrt(d2_k), ball test
//   sum_{c in S} diff^2 <= r*r  (information_estimators.py:263,:268,:272-273)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
ksg_generic_kernel(const JobDesc* __restrict__ jobs, int blocks_per_job, int dx, int dy, int dz, int k,
                   int euclid, int knn_only, PointOut po)
{
    const int job = blockIdx.x / blocks_per_job;
    const int qb  = blockIdx.x - job * blocks_per_job;
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    const int i = qb * TPB + threadIdx.x;
    if (i >= n) return;
    const int D = dx + dy + dz;
    double qc[MAXD];
    for (int c = 0; c < D; c++) qc[c] = jd.col[c][i];

    double L[GEN_KCAP];
    const int kc = k + 1;
    for (int t = 0; t < kc; t++) L[t] = INFINITY;
    for (int j = 0; j < n; j++) {
        double d = 0.0;
        for (int c = 0; c < D; c++) {
            const double df = qc[c] - jd.col[c][j];
            d = euclid ? __dadd_rn(d, __dmul_rn(df, df)) : fmax(d, fabs(df));
        }
        if (d < L[kc - 1]) {
            int t = kc - 1;
            while (t > 0 && L[t - 1] > d) { L[t] = L[t - 1]; --t; }
            L[t] = d;
        }
    }
    double eps = L[k];
    double thr = eps;
    if (euclid) { eps = __dsqrt_rn(eps); thr = __dmul_rn(eps, eps); }
    po.eps[jd.qoff + i] = eps;
    if (knn_only) return;

    int nxyz = 0, nxz = 0, nyz = 0, nz = 0;
    double wyz = 0.0, wz = 0.0, wy = 0.0;
    const bool weighted = jd.w != nullptr;
    for (int j = 0; j < n; j++) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int c = 0; c < dx; c++) { const double df = qc[c] - jd.col[c][j]; ax = euclid ? __dadd_rn(ax, __dmul_rn(df, df)) : fmax(ax, fabs(df)); }
        for (int c = dx; c < dx + dy; c++) { const double df = qc[c] - jd.col[c][j]; ay = euclid ? __dadd_rn(ay, __dmul_rn(df, df)) : fmax(ay, fabs(df)); }
        for (int c = dx + dy; c < D; c++) { const double df = qc[c] - jd.col[c][j]; az = euclid ? __dadd_rn(az, __dmul_rn(df, df)) : fmax(az, fabs(df)); }
        const bool px = ax <= thr, py = ay <= thr, pz = az <= thr;      // dz == 0 -> az = 0 -> pz true
        const bool pxz = px & pz, pyz = py & pz, pxyz = pxz & py;
        nz += pz; nxz += pxz; nyz += pyz; nxyz += pxyz;
        if (weighted) { const double wj = jd.w[j]; wz += pz ? wj : 0.0; wyz += pyz ? wj : 0.0; wy += py ? wj : 0.0; }
    }
    int32_t* o = po.cnt + 4 * (jd.qoff + i);
    o[0] = nxyz - 1; o[1] = nxz - 1; o[2] = nyz - 1; o[3] = nz - 1;
    if (po.wsum && weighted) {
        const double wi = jd.w[i];
        po.wsum[2 * (jd.qoff + i)]     = (euclid ? wy : wyz) - wi;      // umi: W_y - w_i ; cumi: W_yz - w_i
        po.wsum[2 * (jd.qoff + i) + 1] = wz - wi;
    }
}

// Finish kernel for the generic path: one CTA per job, deterministic strided accumulation + block sum.
//   est 0 mi, 1 cmi, 2 umi (sum of w_i (log n_x + log W_y); flag set on log(<=0)), 3 cumi
__global__ void __launch_bounds__(TPB)
ksg_finish_kernel(const JobDesc* __restrict__ jobs, int est, const double* __restrict__ psi_tab,
                  PointOut po, double* __restrict__ out, int* __restrict__ domain_flag)
{
    __shared__ double red[TPB / 32];
    const JobDesc& jd = jobs[blockIdx.x];
    const int n = jd.n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += TPB) {
        const int32_t* c = po.cnt + 4 * (jd.qoff + i);
        double term;
        if (est == 0)      term = psi_int(psi_tab, n) + psi_int(psi_tab, c[0]) - psi_int(psi_tab, c[1]) - psi_int(psi_tab, c[2]);
        else if (est == 1) term = psi_int(psi_tab, c[0]) - psi_int(psi_tab, c[1]) - psi_int(psi_tab, c[2]) + psi_int(psi_tab, c[3]);
        else if (est == 3) {
            const double wi = jd.w[i];
            term = wi * psi_int(psi_tab, c[0]) + wi * -psi_int(psi_tab, c[1]) + wi * -digamma_f64(po.wsum[2 * (jd.qoff + i)])
                 + wi * digamma_f64(po.wsum[2 * (jd.qoff + i) + 1]);
        } else {
            const double wi = jd.w[i];
            const int nx = c[1];
            const double Wy = po.wsum[2 * (jd.qoff + i)];
            // math.log raises ValueError for arguments <= 0 (and for NaN? no: log(nan)=nan) -- information_estimators.py:274,:276
            if (nx <= 0 || Wy <= 0.0) *domain_flag = 1;
            term = wi * log((double)nx) + wi * log(Wy);
        }
        acc += term;
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) out[blockIdx.x] = (est == 2) ? s : ((n > 0) ? s / (double)n : NAN);
}

// ------------------------------------------------------------------------------------------------
// Gaussian KDE inverse densities (replaces KernelDensity(bandwidth=bw).fit(P).score_samples(P),
// information_estimators.py:248-251, :385-388).  u_i = 1 / sum_j exp(-|P_i-P_j|^2 / (2 bw^2));
// the normalising constant and 1/N cancel in w = u / mean(u).  Differences and squares in FP64
// (the exponent argument needs ~1e-9 absolute accuracy in d^2 at bw=0.01), 2^x on the MUFU in FP32.
// Coordinates are pre-scaled by sqrt(log2(e)/(2 bw^2)) so the argument is just -(sum of squares).
// ------------------------------------------------------------------------------------------------
template <int DP, int Q>
__global__ void __launch_bounds__(TPB)
kde_kernel(const JobDesc* __restrict__ jobs, int blocks_per_job, const int* __restrict__ dims, double scale,
           double* __restrict__ u_out)
{
    __shared__ double tile[DP][TILE];
    const int job = blockIdx.x / blocks_per_job;
    const int qb  = blockIdx.x - job * blocks_per_job;
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    const int q0 = qb * (TPB * Q);
    if (q0 >= n) return;
    const int tid = threadIdx.x;
    const double* cp[DP];
    #pragma unroll
    for (int c = 0; c < DP; c++) cp[c] = jd.col[dims[c]];

    double qc[Q][DP];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = q0 + q * TPB + tid;
        const int ii = i < n ? i : n - 1;
        #pragma unroll
        for (int c = 0; c < DP; c++) qc[q][c] = cp[c][ii] * scale;
    }
    double S[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) S[q] = 0.0;

    for (int base = 0; base < n; base += TILE) {
        const int cnt = min(TILE, n - base);
        __syncthreads();
        for (int t = tid; t < TILE; t += TPB) {
            #pragma unroll
            for (int c = 0; c < DP; c++) tile[c][t] = (t < cnt) ? cp[c][base + t] * scale : INFINITY;   // exp2(-inf) = 0
        }
        __syncthreads();
        float s32[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) s32[q] = 0.f;
        const int cnt4 = (cnt + 3) & ~3;
        #pragma unroll 4
        for (int j = 0; j < cnt4; j++) {
            double cc[DP];
            #pragma unroll
            for (int c = 0; c < DP; c++) cc[c] = tile[c][j];
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                double d2 = 0.0;
                #pragma unroll
                for (int c = 0; c < DP; c++) { const double df = qc[q][c] - cc[c]; d2 = fma(df, df, d2); }
                float e;
                const float a = -(float)d2;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
                s32[q] += e;
            }
        }
        #pragma unroll
        for (int q = 0; q < Q; q++) S[q] += (double)s32[q];
    }
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int i = q0 + q * TPB + tid;
        if (i < n) u_out[jd.qoff + i] = 1.0 / S[q];
    }
}

// kNN-density inverse densities: u_i = 1/dens_i, dens_i = k_density / N / r_i^dim
// (information_estimators.py:255-256, :391-392).  r_i == 0 -> dens = inf -> u = 0, as in numpy.
__global__ void knn_density_kernel(const JobDesc* __restrict__ jobs, int blocks_per_job, int k_density, int dim,
                                   const double* __restrict__ r, double* __restrict__ u_out)
{
    const int job = blockIdx.x / blocks_per_job;
    const int qb  = blockIdx.x - job * blocks_per_job;
    const JobDesc& jd = jobs[job];
    const int i = qb * TPB + threadIdx.x;
    if (i >= jd.n) return;
    const double dens = (double)k_density / (double)jd.n / pow(r[jd.qoff + i], (double)dim);
    u_out[jd.qoff + i] = 1.0 / dens;
}

// w = u / mean(u), in place; one CTA per job, deterministic.
__global__ void __launch_bounds__(TPB)
normalize_weights_kernel(const JobDesc* __restrict__ jobs, double* __restrict__ u)
{
    __shared__ double red[TPB / 32];
    __shared__ double mean_s;
    const JobDesc& jd = jobs[blockIdx.x];
    const int n = jd.n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += TPB) acc += u[jd.qoff + i];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) mean_s = s / (double)n;
    __syncthreads();
    const double m = mean_s;
    for (int i = threadIdx.x; i < n; i += TPB) u[jd.qoff + i] = u[jd.qoff + i] / m;
}

// ------------------------------------------------------------------------------------------------
// Lagged-embedding materialisation (replaces the per-pair list comprehensions at
// causal_network.py:147-153,:167-173,:266-277).  One output column = (gene, shift, tau, diff, scale):
//   for run r, t in [0, T_r - tau):  v = E[gene][run_off[r] + t + shift]   (diff: E[p-1] - E[p])
// ------------------------------------------------------------------------------------------------
struct ColSpec { int32_t gene, shift, tau, diff; double scale; };

__global__ void __launch_bounds__(TPB)
gather_lagged_kernel(const double* __restrict__ E, int64_t n_cells, const int64_t* __restrict__ run_off, int n_runs,
                     const ColSpec* __restrict__ specs, double* __restrict__ out, int64_t col_stride)
{
    const ColSpec sp = specs[blockIdx.y];
    const double* row = E + (size_t)sp.gene * n_cells;
    double* o = out + (size_t)blockIdx.y * col_stride;
    int64_t lag_base = 0;
    for (int r = 0; r < n_runs; r++) {
        const int64_t T = run_off[r + 1] - run_off[r];
        const int64_t len = T - sp.tau;
        if (len <= 0) continue;
        for (int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x; t < len; t += (int64_t)gridDim.x * TPB) {
            const int64_t p = run_off[r] + t + sp.shift;
            double v = sp.diff ? (row[p - 1] - row[p]) : row[p];
            if (sp.scale != 1.0) v = v / sp.scale;
            o[lag_base + t] = v;
        }
        lag_base += len;
    }
}

// ------------------------------------------------------------------------------------------------
// Descriptors of plain RDI jobs (no conditioning set, no per-job scaling) built on the device from the caller's job
// table.  Column layout of the chunk: column ((di * G + gene) * 3 + role), role 0 = x (shift 0), 1 = y (shift d),
// 2 = z (shift d-1), di = index of the job's delay in `delays`; every column is sorted: perm / sorted / rank are laid out like cols.
// ------------------------------------------------------------------------------------------------
struct RdiLayout { int32_t n_delays; int32_t delays[8]; int32_t n_of[8]; };
struct RawJob { int32_t src, dst, delay, n_cond; int32_t cond_gene[4]; int32_t cond_delay[4]; };   // == scribe_job

__global__ void __launch_bounds__(TPB)
build_rdi_jobs_kernel(const RawJob* __restrict__ raw, int64_t n_jobs, JobDesc* __restrict__ out, const double* __restrict__ cols,
                      int64_t stride, const int32_t* __restrict__ perm, const double* __restrict__ sorted, const float* __restrict__ rank,
                      int n_genes, RdiLayout lay, int64_t n_max)
{
    const int64_t j = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (j >= n_jobs) return;
    const RawJob jb = raw[j];
    int di = 0;
    #pragma unroll
    for (int t = 0; t < 8; t++) if (t < lay.n_delays && lay.delays[t] == jb.delay) di = t;
    JobDesc jd;
    #pragma unroll
    for (int c = 0; c < MAXD; c++) jd.col[c] = nullptr;
    const size_t base = (size_t)di * n_genes;
    jd.col[0] = cols + ((base + jb.src) * 3 + 0) * stride;
    jd.col[1] = cols + ((base + jb.dst) * 3 + 1) * stride;
    jd.col[2] = cols + ((base + jb.dst) * 3 + 2) * stride;
    jd.w = nullptr;
    jd.perm = perm ? perm + ((base + jb.dst) * 3 + 2) * stride : nullptr;          // order of the key: the target's z column
    jd.strip = nullptr; jd.npad = 0;
    jd.qoff = j * n_max;
    jd.n = lay.n_of[di]; jd.pad = 0;
    jd.absmax = 0.0; jd.wcum = nullptr;
    #pragma unroll
    for (int c = 0; c < 4; c++) { jd.srt[c] = nullptr; jd.rk[c] = nullptr; jd.srt_n[c] = 0; }
    if (sorted) {                                    // rank-space count pass: x of the source, y of the target
        jd.srt[0] = sorted + ((base + jb.src) * 3 + 0) * stride; jd.rk[0] = rank + ((base + jb.src) * 3 + 0) * stride;
        jd.srt[1] = sorted + ((base + jb.dst) * 3 + 1) * stride; jd.rk[1] = rank + ((base + jb.dst) * 3 + 1) * stride;
        jd.srt_n[0] = jd.srt_n[1] = lay.n_of[di];
    }
    out[j] = jd;
}

// point every job at its slice of the weight workspace (device-resident descriptors)
__global__ void set_weight_ptr_kernel(JobDesc* __restrict__ jobs, int64_t n_jobs, const double* __restrict__ u)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_jobs) jobs[j].w = u + jobs[j].qoff;
}

// ------------------------------------------------------------------------------------------------
// dynamics-coupling prep (replaces Scribe.py:119-133): per job gather x = S[src], y = Y[dst], z = S[dst],
// keep cells with (x + y) + z > 0 (float64, this association order) and compact them stably -- visiting the cells
// in the target's z-sorted order when permS is given, so the compacted columns come out sorted for the sweep kernels
// (the estimate does not depend on the order of the samples).
// One CTA per job; writes 3 columns of stride `col_stride` and the kept count into jobs[job].n.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB)
coupling_prep_kernel(const double* __restrict__ S, const double* __restrict__ Y, int64_t n_cells,
                     const int32_t* __restrict__ permS,   // [G][n_cells]: cells of each gene sorted by S (or nullptr)
                     const float* __restrict__ rankS,     // [G][n_cells]: rank of every cell in its gene's ascending S row (or nullptr)
                     const float* __restrict__ rankY,     //               ... in its gene's ascending Y row
                     float* __restrict__ crank,           // [jobs][2][col_stride]: ranks of the kept cells (x in S[src], y in Y[dst])
                     const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int drop_zero,
                     double* __restrict__ cols, int64_t col_stride, JobDesc* __restrict__ jobs, int32_t* __restrict__ n_eff)
{
    __shared__ int warp_cnt[TPB / 32];
    __shared__ int base_s;
    const int job = blockIdx.x;
    const double* xs = S + (size_t)src[job] * n_cells;
    const double* zs = S + (size_t)dst[job] * n_cells;
    const double* ys = Y + (size_t)dst[job] * n_cells;
    double* ox = cols + (size_t)job * 3 * col_stride;
    double* oy = ox + col_stride;
    double* oz = oy + col_stride;
    float* orx = crank ? crank + (size_t)job * 2 * col_stride : nullptr;
    float* ory = crank ? orx + col_stride : nullptr;
    const float* rxs = rankS ? rankS + (size_t)src[job] * n_cells : nullptr;
    const float* rys = rankY ? rankY + (size_t)dst[job] * n_cells : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n_cells; c0 += TPB) {
        const int64_t r = c0 + tid;                      // rank in the target's z-sorted order (or plain cell index)
        double x = 0, y = 0, z = 0;
        float rx = 0.f, ry = 0.f;
        bool keep = false;
        if (r < n_cells) {
            const int64_t c = permS ? permS[(size_t)dst[job] * n_cells + r] : r;
            x = xs[c]; y = ys[c]; z = zs[c];
            if (orx) { rx = rxs[c]; ry = rys[c]; }
            keep = drop_zero ? (__dadd_rn(__dadd_rn(x, y), z) > 0.0) : true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; w++) off += warp_cnt[w];
        if (keep) {
            const int pos = off + __popc(m & ((1u << lane) - 1u));
            ox[pos] = x; oy[pos] = y; oz[pos] = z;
            if (orx) { orx[pos] = rx; ory[pos] = ry; }
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < TPB / 32; w++) t += warp_cnt[w]; base_s += t; }
        __syncthreads();
    }
    if (tid == 0) {
        jobs[job].n = base_s;
        if (n_eff) n_eff[job] = base_s;
    }
}

}  // namespace scribe
