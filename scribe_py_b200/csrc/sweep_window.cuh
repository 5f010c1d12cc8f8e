// sweep_window.cuh -- second form of the sorted sweep (cmi, scalar x and y, 1..3 conditioning columns).
//
// Same contract as ksg_sweep_kernel (kernels.cuh): the points of a job are visited in the order of the LAST conditioning
// column ("key"), a warp owns 32*Q consecutive ranks, results are bit-identical to the dense kernel.  Two changes cut the
// FP64-pipe work per visited pair from 6 + 7 instructions to 0 + 4:
//
//   pass 1 (k-th neighbour distance, replaces tree_xyz.query(...)[0][k], information_estimators.py:165):
//     the per-candidate scan is an FP32 *pre-filter* over the non-key coordinates.  A candidate is marked when
//     |fl32(q_d) - fl32(c_d)| <= T for every non-key d, with T = ru32(thr*(1+2^-22) + absmax*2^-22 + 1e-37), thr the
//     query's k-th distance at the start of the chunk and absmax the largest |coordinate| of the job
//     (job_absmax_kernel).  T bounds every rounding the FP32 path can make, so the marked set is a superset of the
//     candidates the FP64 test `max_d |q_d - c_d| < thr` accepts; each marked candidate is then re-tested in FP64,
//     exactly as in the first sweep, before it enters the top-(k+1) list.  Unordered compares mark NaN/overflowed
//     differences too, so a job whose absmax is huge or non-finite degrades to "mark everything", never to a wrong result.
//
//   pass 2 (the four ball counts, replaces 4 x query_ball_point, :169-172):
//     every counted subspace contains the key, and the key is sorted, so the candidates with |key_i - key_j| <= eps_i
//     are one contiguous rank window [lo_i, hi_i], found by two binary searches that use the very subtraction the dense
//     kernel uses (monotone rounding => the same set).  Inside the warp's chunk range only the x and y (and extra z)
//     differences are formed -- 2 DADD + 2 DSETP per pair -- and their outcomes are collected as one bit per candidate;
//     the counts are popc(window & ...) once per chunk.  n_z for a single conditioning column is hi - lo + 1 - 1.
//
// Everything that decides a count is still FP64.
#pragma once
#include "kernels.cuh"

namespace scribe {

#ifndef SWEEPW_UNROLL
#define SWEEPW_UNROLL 16                  // candidates per scan block, unweighted kernels.  Measured on B200 (r2, tools/ab.py) 16 vs 8:
#endif                                    // C2 -2.5 %, C3 -3.7 %, C4 -3.0 %, cmi at N = 20 000 -3.6 %; the weighted (cumi) kernels lose 9 %
#ifndef SWEEPW_UNROLL_W                   // with 16 (code size), so they keep 8
#define SWEEPW_UNROLL_W 8
#endif
#ifndef SWEEPW_MERGED
#define SWEEPW_MERGED 0                   // 1: one marked-candidate loop for all query slots of a lane.  Measured on B200 (r2): +1..3 % time, off
#endif
#ifndef SWEEPW_UNROLL_C
#define SWEEPW_UNROLL_C SWEEPW_UNROLL
#endif
#ifndef SWEEPW_UNROLL_CW
#define SWEEPW_UNROLL_CW SWEEPW_UNROLL_W
#endif
#ifndef SWEEPW_DIAG
#define SWEEPW_DIAG 0                     // 1: diagnostics build (visited_kde reports the count-pass visits inside the lane's own window)
#endif
#ifndef SWEEPW_WFMA
#define SWEEPW_WFMA 0                     // 1: weighted count pass accumulates W_yz with an indicator DFMA instead of a predicated DADD
#endif
#ifndef SWEEPW_CPASYNC
#define SWEEPW_CPASYNC 0                  // 1: pass 2 of unweighted jobs with pre-sorted columns (dynamics coupling) stages its candidate
#endif                                    // chunks with cp.async into a double buffer instead of LDG -> registers -> STS (experiment, r2)
#ifndef SWEEPW_FINE_ALL
#define SWEEPW_FINE_ALL 0                 // 1: insert marked candidates after every scan block in every chunk, not only in the centre chunks
#endif
#ifndef SWEEPW_RANK
#define SWEEPW_RANK 0                     // 1: pass 2 tests integer rank windows on the FP32/ALU pipes instead of FP64 differences (bit-identical;
                                          // measured on B200, r2: +4 % time at N = 2000, -1.5 % on cumi at N = 20 000 -- see DESIGN.md)
#endif
#ifndef SWEEPW_MIN_CTAS
#define SWEEPW_MIN_CTAS 10                // 102 registers: 20 resident warps per SM; measured equal at N = 2000 and 8 % faster at N = 20 000 than 12 CTAs / 80 registers
#endif

// largest |coordinate| over the D columns of each job -> JobDesc.absmax (one CTA per job; O(N*D) reads from L2)
__global__ void __launch_bounds__(128)
job_absmax_kernel(JobDesc* __restrict__ jobs, int D)
{
    __shared__ double red[4];
    JobDesc& jd = jobs[blockIdx.x];
    const int n = jd.n;
    double m = 0.0;
    bool bad = false;
    for (int d = 0; d < D; d++) {
        const double* __restrict__ c = jd.col[d];
        for (int i = threadIdx.x; i < n; i += 128) { const double a = fabs(c[i]); bad |= !(a <= 1e30); m = (a > m) ? a : m; }
    }
    if (bad) m = INFINITY;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_down_sync(0xffffffffu, m, o); m = (t > m) ? t : m; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 4; w++) m = (red[w] > m) ? red[w] : m;
        jd.absmax = m;
    }
}

// FP32 pre-filter step: mark `bit` unless some non-key difference is certainly above T (unordered-or-less-equal).
template <int NP>
__device__ __forceinline__ void knn_mark32(const float (&df)[NP], float T, unsigned bit, unsigned& mask)
{
    static_assert(NP >= 1 && NP <= 4, "1..4 non-key coordinates");
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

// pass-2 step: one bit per candidate and tested coordinate (closed ball, FP64): m |= bit when |d| <= e
__device__ __forceinline__ void ball_bit(double d, double e, unsigned bit, unsigned& m)
{
    asm("{\n .reg .pred p;\n .reg .f64 a;\n abs.f64 a,%1;\n setp.le.f64 p, a, %2;\n @p or.b32 %0,%0,%3;\n}"
        : "+r"(m) : "d"(d), "d"(e), "r"(bit));
}
__device__ __forceinline__ void ball_bit2(double d0, double d1, double e, unsigned bit, unsigned& m)
{
    asm("{\n .reg .pred p;\n .reg .f64 a,b;\n abs.f64 a,%1; abs.f64 b,%2;\n setp.le.f64 p, a, %3;\n setp.le.and.f64 p, b, %3, p;\n"
        " @p or.b32 %0,%0,%4;\n}"
        : "+r"(m) : "d"(d0), "d"(d1), "d"(e), "r"(bit));
}

// weighted pass-2 steps (cumi): `win` holds the candidate's in-window bit at position `bit`.
//   DZ == 1: y-ball inside the window -> by bit and W_yz += w          (W_z comes from prefix sums of the window)
//   DZ  > 1: z-ball (extra conditioning columns) inside the window -> bz bit, W_z += w; then the y-ball inside it
__device__ __forceinline__ void wball_y(double dy, double e, double w, unsigned win, unsigned bit, unsigned& by, double& wyz)
{
#if SWEEPW_WFMA
    // W_yz += w * [inside] as one DFMA with a 0.0 / 1.0 indicator built from the predicate (SEL of the high word): ptxas turns
    // a predicated add.f64 into DADD + 2 FSEL, this form is SEL + DFMA.  fma(w, 1, acc) == acc + w exactly.
    asm("{\n .reg .pred pin,p;\n .reg .b32 t,hi;\n .reg .f64 a,ind;\n and.b32 t,%4,%5;\n setp.ne.u32 pin,t,0;\n abs.f64 a,%2;\n"
        " setp.le.and.f64 p,a,%3,pin;\n selp.b32 hi,0x3ff00000,0,p;\n mov.b64 ind,{%7,hi};\n fma.rn.f64 %1,%6,ind,%1;\n @p or.b32 %0,%0,%5;\n}"
        : "+r"(by), "+d"(wyz) : "d"(dy), "d"(e), "r"(win), "r"(bit), "d"(w), "r"(0));
#else
    asm("{\n .reg .pred pin,p;\n .reg .b32 t;\n .reg .f64 a;\n and.b32 t,%4,%5;\n setp.ne.u32 pin,t,0;\n abs.f64 a,%2;\n"
        " setp.le.and.f64 p,a,%3,pin;\n @p add.f64 %1,%1,%6;\n @p or.b32 %0,%0,%5;\n}"
        : "+r"(by), "+d"(wyz) : "d"(dy), "d"(e), "r"(win), "r"(bit), "d"(w));
#endif
}
__device__ __forceinline__ void wball_zy(double dz0, double dz1, bool two, double dy, double e, double w, unsigned win, unsigned bit,
                                         unsigned& bz, unsigned& by, double& wz, double& wyz)
{
    if (two)
        asm("{\n .reg .pred pz,py;\n .reg .b32 t;\n .reg .f64 a,b,c;\n and.b32 t,%8,%9;\n setp.ne.u32 pz,t,0;\n abs.f64 a,%4; abs.f64 b,%5; abs.f64 c,%6;\n"
            " setp.le.and.f64 pz,a,%7,pz;\n setp.le.and.f64 pz,b,%7,pz;\n setp.le.and.f64 py,c,%7,pz;\n"
            " @pz add.f64 %2,%2,%10;\n @pz or.b32 %0,%0,%9;\n @py add.f64 %3,%3,%10;\n @py or.b32 %1,%1,%9;\n}"
            : "+r"(bz), "+r"(by), "+d"(wz), "+d"(wyz) : "d"(dz0), "d"(dz1), "d"(dy), "d"(e), "r"(win), "r"(bit), "d"(w));
    else
        asm("{\n .reg .pred pz,py;\n .reg .b32 t;\n .reg .f64 a,c;\n and.b32 t,%7,%8;\n setp.ne.u32 pz,t,0;\n abs.f64 a,%4; abs.f64 c,%5;\n"
            " setp.le.and.f64 pz,a,%6,pz;\n setp.le.and.f64 py,c,%6,pz;\n"
            " @pz add.f64 %2,%2,%9;\n @pz or.b32 %0,%0,%8;\n @py add.f64 %3,%3,%9;\n @py or.b32 %1,%1,%8;\n}"
            : "+r"(bz), "+r"(by), "+d"(wz), "+d"(wyz) : "d"(dz0), "d"(dy), "d"(e), "r"(win), "r"(bit), "d"(w));
}

// ---- rank-space pass 2 -------------------------------------------------------------------------------------------------------
// A ball |q_c - v| <= eps in one coordinate is a contiguous range [lo_c, hi_c] of that coordinate's ascending order (found by
// binary search with the dense kernel's own subtraction, so it is the same set).  With centre m = (lo + hi) / 2 and half-width
// h = (hi - lo) / 2 (half-integers, exact in FP32 below 2^22) a candidate of rank r is in the ball iff |r - m| <= h: one packed
// FP32 subtract for two coordinates, one FSETP and one predicated OR per coordinate -- nothing on the FP64 pipe.
__device__ __forceinline__ void rank_bit1(float d0, float h0, unsigned bit, unsigned& b0)
{
    asm("{\n .reg .pred p;\n .reg .f32 a0;\n abs.f32 a0,%1;\n setp.le.f32 p,a0,%2;\n @p or.b32 %0,%0,%3;\n}"
        : "+r"(b0) : "f"(d0), "f"(h0), "r"(bit));
}
__device__ __forceinline__ void rank_bit2(float d0, float d1, float h0, float h1, unsigned bit, unsigned& b0, unsigned& b1)
{
    asm("{\n .reg .pred p,q;\n .reg .f32 a0,a1;\n abs.f32 a0,%2; abs.f32 a1,%3;\n setp.le.f32 p,a0,%4;\n setp.le.f32 q,a1,%5;\n"
        " @p or.b32 %0,%0,%6;\n @q or.b32 %1,%1,%6;\n}"
        : "+r"(b0), "+r"(b1) : "f"(d0), "f"(d1), "f"(h0), "f"(h1), "r"(bit));
}
// both extra conditioning columns inside -> one bit
__device__ __forceinline__ void rank_bit_and2(float d0, float d1, float h0, float h1, unsigned bit, unsigned& b)
{
    asm("{\n .reg .pred p;\n .reg .f32 a0,a1;\n abs.f32 a0,%1; abs.f32 a1,%2;\n setp.le.f32 p,a0,%3;\n setp.le.and.f32 p,a1,%4,p;\n"
        " @p or.b32 %0,%0,%5;\n}"
        : "+r"(b) : "f"(d0), "f"(d1), "f"(h0), "f"(h1), "r"(bit));
}
// weighted (cumi), one conditioning column: y inside the key window -> by bit and W_yz += w
__device__ __forceinline__ void rank_wy(float dy, float hy, double w, unsigned win, unsigned bit, unsigned& by, double& wyz)
{
    asm("{\n .reg .pred pin,p;\n .reg .b32 t;\n .reg .f32 a;\n and.b32 t,%4,%5;\n setp.ne.u32 pin,t,0;\n abs.f32 a,%2;\n"
        " setp.le.and.f32 p,a,%3,pin;\n @p add.f64 %1,%1,%6;\n @p or.b32 %0,%0,%5;\n}"
        : "+r"(by), "+d"(wyz) : "f"(dy), "f"(hy), "r"(win), "r"(bit), "d"(w));
}
// weighted, 2 or 3 conditioning columns: z-ball inside the key window -> bz bit, W_z += w; y inside it -> by bit, W_yz += w
__device__ __forceinline__ void rank_wzy(float dz0, float hz0, float dz1, float hz1, bool two, float dy, float hy, double w, unsigned win,
                                         unsigned bit, unsigned& bz, unsigned& by, double& wz, double& wyz)
{
    if (two)
        asm("{\n .reg .pred pz,py;\n .reg .b32 t;\n .reg .f32 a,b,c;\n and.b32 t,%10,%11;\n setp.ne.u32 pz,t,0;\n abs.f32 a,%4; abs.f32 b,%6; abs.f32 c,%8;\n"
            " setp.le.and.f32 pz,a,%5,pz;\n setp.le.and.f32 pz,b,%7,pz;\n setp.le.and.f32 py,c,%9,pz;\n"
            " @pz add.f64 %2,%2,%12;\n @pz or.b32 %0,%0,%11;\n @py add.f64 %3,%3,%12;\n @py or.b32 %1,%1,%11;\n}"
            : "+r"(bz), "+r"(by), "+d"(wz), "+d"(wyz)
            : "f"(dz0), "f"(hz0), "f"(dz1), "f"(hz1), "f"(dy), "f"(hy), "r"(win), "r"(bit), "d"(w));
    else
        asm("{\n .reg .pred pz,py;\n .reg .b32 t;\n .reg .f32 a,c;\n and.b32 t,%8,%9;\n setp.ne.u32 pz,t,0;\n abs.f32 a,%4; abs.f32 c,%6;\n"
            " setp.le.and.f32 pz,a,%5,pz;\n setp.le.and.f32 py,c,%7,pz;\n"
            " @pz add.f64 %2,%2,%10;\n @pz or.b32 %0,%0,%9;\n @py add.f64 %3,%3,%10;\n @py or.b32 %1,%1,%9;\n}"
            : "+r"(bz), "+r"(by), "+d"(wz), "+d"(wyz)
            : "f"(dz0), "f"(hz0), "f"(dy), "f"(hy), "r"(win), "r"(bit), "d"(w));
}

// |v| without the FP64 pipe (fabs() of a value that feeds a select compiles to DADD -RZ,|v|)
__device__ __forceinline__ double abs_f64(double v)
{
    return __hiloint2double(__double2hiint(v) & 0x7fffffff, __double2loint(v));
}

// bits [a, b] of a 32-bit word, for any integers a, b (empty when the range misses 0..31)
__device__ __forceinline__ unsigned range_mask(int a, int b)
{
    a = max(a, 0); b = min(b, 31);
    return (b >= a) ? ((0xffffffffu >> (31 - b + a)) << a) : 0u;
}

template <int DZ, int KCAP, int Q, bool WEIGHTED>
__global__ void __launch_bounds__(SWEEP_TPB, SWEEPW_MIN_CTAS)
ksg_sweepw_kernel(const JobDesc* __restrict__ jobs, int warps_per_job, int k, const double* __restrict__ psi_tab,
                  double* __restrict__ partial, PointOut po, unsigned long long* __restrict__ visited)
{
    static_assert(DZ >= 0 && DZ <= 3, "0..3 conditioning columns; DZ = 0 is mi(x, y), or umi(x, y) when WEIGHTED: sorted by y");
    // umi (information_estimators.py:209-277): EUCLIDEAN joint k-NN (the list holds squared distances d2 = dx^2 + dy^2, formed
    // without FMA as scipy does), r = sqrt(d2_k), balls by scipy's squared-radius test diff^2 <= fl(r * r) (:263-273).  Both
    // marginal balls are rank ranges of a sorted column, so there is no count scan at all: n_x = range in the sorted x column,
    // W_y = difference of two prefix sums of the weights in y order (JobDesc.wcum).
    constexpr bool UMI = DZ == 0 && WEIGHTED;
    constexpr int D   = 2 + DZ;
    constexpr int KEY = D - 1;
    constexpr int NP  = D - 1;                    // non-key coordinates: x, y, z_0 .. z_{DZ-2}
    constexpr int NPA = (NP + 1) & ~1;            // padded to a multiple of 2 -> 16-byte rows of doubles, 8/16-byte rows of floats
    constexpr int W   = SWEEP_TPB / 32;
    constexpr int UNR = WEIGHTED ? SWEEPW_UNROLL_W : SWEEPW_UNROLL;
    constexpr int UNRC = WEIGHTED ? SWEEPW_UNROLL_CW : SWEEPW_UNROLL_C;        // count-pass scan blocks
    constexpr int KL  = KCAP - 1;                 // list length: the k nearest OTHER points (the query itself is never inserted)
    __shared__ __align__(16) double sd[W][32][NPA];          // candidates, non-key coordinates (FP64)
#if SWEEPW_CPASYNC
    __shared__ __align__(16) double sd2[WEIGHTED ? 1 : W][WEIGHTED ? 1 : 32][NPA];      // second buffer of the cp.async pipeline
#endif
    __shared__ __align__(16) float  sf[W][32][NPA];          // the same rounded to FP32 (pass 1 pre-filter)
    __shared__ __align__(8)  double sk[W][32];               // candidates, key coordinate
    __shared__ __align__(8)  double sw[WEIGHTED ? W : 1][WEIGHTED ? 66 : 1];   // cumi: weights [0..31]; DZ == 1: their exclusive prefix sums at [32..64]

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bpj = (warps_per_job + W - 1) / W;
    const int job = blockIdx.x / bpj;
    const int wj  = (blockIdx.x - job * bpj) * W + wid;
    const JobDesc& jd = jobs[job];
    const int n = jd.n;
    const int w0 = wj * 32 * Q;
    if (wj >= warps_per_job || w0 >= n) return;                  // warp-uniform; nothing below synchronises the CTA
    const int c0 = w0 / 32;
    const int nchunks = (n + 31) / 32;
    double (*bd)[NPA] = sd[wid];
    float  (*bf)[NPA] = sf[wid];
    double* bk = sk[wid];
    double* bw = sw[WEIGHTED ? wid : 0];
    const int32_t* __restrict__ perm = jd.perm;

    double qc[Q][D], qw[Q];
    unsigned long long qf2[Q][NPA / 2];           // FP32 images of the non-key query coordinates, packed in pairs
    int qidx[Q]; bool qok[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        const int rank = w0 + q * 32 + lane;
        qok[q] = rank < n;
        const int rr = qok[q] ? rank : n - 1;
        qidx[q] = perm ? perm[rr] : rr;
        #pragma unroll
        for (int d = 0; d < D; d++) qc[q][d] = jd.col[d][qidx[q]];
        qw[q] = 0.0;
        if (WEIGHTED) qw[q] = jd.w[qidx[q]];
    }

    auto stage = [&](const Chunk<D, WEIGHTED>& c, bool with_f32) {
        __syncwarp();
        #pragma unroll
        for (int d = 0; d < NP; d += 2) {
            double2 v; v.x = c.v[d]; v.y = (d + 1 < NP) ? c.v[d + 1] : 0.0;
            *reinterpret_cast<double2*>(&bd[lane][d]) = v;
        }
        bk[lane] = c.v[KEY];
        if (with_f32) {
            #pragma unroll
            for (int d = 0; d < NP; d += 2) {
                float2 v; v.x = __double2float_rn(c.v[d]); v.y = (d + 1 < NP) ? __double2float_rn(c.v[d + 1]) : 0.f;
                *reinterpret_cast<float2*>(&bf[lane][d]) = v;
            }
        } else if (WEIGHTED) {
            bw[lane] = c.w;
            if (DZ == 1) {                                         // exclusive prefix sums of the chunk's weights: bw[32 + j] = w_0 + .. + w_{j-1}
                double ps = c.w;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, ps, o); if (lane >= o) ps += t; }
                bw[33 + lane] = ps;
                if (lane == 0) bw[32] = 0.0;
            }
        }
        __syncwarp();
    };

    // ---------------- pass 1: k-th neighbour distance ----------------
    double L[Q][KL];
    #pragma unroll
    for (int q = 0; q < Q; q++)
        #pragma unroll
        for (int t = 0; t < KL; t++) L[q][t] = (t < KL - k) ? -INFINITY : INFINITY;
    const double slack = jd.absmax * 0x1p-22 + 1e-37;           // +inf when the job's scale is not FP32-safe: everything is marked

    // FP32 bound of the pre-filter from the current bar (a distance; for umi a squared distance: the bound applies to |dx|)
    auto bound32 = [&](double bar) {
        const double lin = UMI ? __dsqrt_ru(bar) : bar;
        return __double2float_ru(__dadd_ru(__dmul_ru(lin, 1.0 + 0x1p-22), slack));
    };
    // exact joint distance of a query and staged candidate j: Chebyshev maximum, or for umi the squared Euclidean distance
    auto exact_dist = [&](int q, int j) {
        if constexpr (UMI) {
            const double dx = qc[q][0] - bd[j][0], dy = qc[q][KEY] - bk[j];
            return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        } else {
            double dd = abs_f64(qc[q][KEY] - bk[j]);
            #pragma unroll
            for (int d = 0; d < NP; d++) { const double a = abs_f64(qc[q][d] - bd[j][d]); dd = (a > dd) ? a : dd; }
            return dd;
        }
    };

    // selfq: slot whose own points are the staged chunk (its lane-th candidate is the query itself and is skipped), or -1;
    // fine: insert after every block of UNR candidates instead of once per chunk (centre chunks: the bar drops fast there)
    auto knn_body = [&](auto q_first, int selfq, bool fine) {
        constexpr int Q0 = decltype(q_first)::value;
        unsigned mask[Q];
        float T[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) { mask[q] = 0u; T[q] = bound32(L[q][KL - 1]); }
        // blocks of UNR candidates: bits are set with immediates inside a block and shifted into place once per block
        // (a fully unrolled scan does not fit the instruction cache next to the other phases: ncu no_inst 37 %)
        #pragma unroll 1
        for (int jb = 0; jb < 32; jb += UNR) {
            unsigned mb[Q];
            #pragma unroll
            for (int q = 0; q < Q; q++) mb[q] = 0u;
            #pragma unroll
            for (int t = 0; t < UNR; t++) {
                const int j = jb + t;
                unsigned long long cf2[NPA / 2];                     // packed FP32 pairs: one FADD2 forms two differences
                if constexpr (NPA == 2) { cf2[0] = *reinterpret_cast<const unsigned long long*>(&bf[j][0]); }
                else { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&bf[j][0]); cf2[0] = v.x; cf2[NPA / 2 - 1] = v.y; }
                #pragma unroll
                for (int q = Q0; q < Q; q++) {
                    float df[NPA];
                    #pragma unroll
                    for (int d = 0; d < NPA; d += 2) {
                        unsigned long long dd;
                        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dd) : "l"(qf2[q][d / 2]), "l"(cf2[d / 2]));
                        asm("mov.b64 {%0,%1}, %2;" : "=f"(df[d]), "=f"(df[d + 1]) : "l"(dd));
                    }
                    float dn[NP];
                    #pragma unroll
                    for (int d = 0; d < NP; d++) dn[d] = df[d];
                    knn_mark32<NP>(dn, T[q], 1u << t, mb[q]);
                }
            }
            #pragma unroll
            for (int q = Q0; q < Q; q++) mask[q] |= mb[q] << jb;
            if (!fine && jb + UNR < 32) continue;
#if SWEEPW_MERGED
            // One loop for all query slots of the lane: every trip takes the next marked candidate of EACH slot, so the exact
            // re-tests of the slots are independent instruction streams (the loop is latency-bound: LDS -> DADD -> DSETP
            // chains at ~25 % lane efficiency) and the trip count is max over lanes of max over slots, not the sum of the
            // per-slot maxima.
            unsigned any = 0u;
            #pragma unroll
            for (int q = Q0; q < Q; q++) { if (q == selfq) mask[q] &= ~(1u << lane); any |= mask[q]; }
            while (any) {
                any = 0u;
                #pragma unroll
                for (int q = Q0; q < Q; q++) {
                    const bool has = mask[q] != 0u;
                    const int j = has ? __ffs(mask[q]) - 1 : 0;
                    mask[q] &= mask[q] - 1;                                    // 0 stays 0
                    any |= mask[q];
                    const double dd = exact_dist(q, j);
                    if (has && dd < L[q][KL - 1]) topk_insert<KL>(L[q], dd);   // exact test; NaN (padding) never passes
                }
            }
#else
            #pragma unroll
            for (int q = Q0; q < Q; q++) {
                if (q == selfq) mask[q] &= ~(1u << lane);
                while (mask[q]) {
                    const int j = __ffs(mask[q]) - 1;
                    mask[q] &= mask[q] - 1;
                    const double dd = exact_dist(q, j);
                    if (dd < L[q][KL - 1]) topk_insert<KL>(L[q], dd);          // exact test; NaN (padding) never passes
                }
            }
#endif
            if (fine) {
                #pragma unroll
                for (int q = Q0; q < Q; q++) T[q] = bound32(L[q][KL - 1]);
            }
        }
    };
    auto knn_chunk = [&](bool all) {
        if (all) knn_body(std::integral_constant<int, 0>{}, -1, SWEEPW_FINE_ALL != 0); else knn_body(std::integral_constant<int, Q - 1>{}, -1, SWEEPW_FINE_ALL != 0);
    };
    auto set_qf = [&]() {
        #pragma unroll
        for (int q = 0; q < Q; q++)
            #pragma unroll
            for (int d = 0; d < NPA; d += 2) {
                const float a = __double2float_rn(qc[q][d]), b = (d + 1 < NP) ? __double2float_rn(qc[q][d + 1]) : 0.f;
                asm("mov.b64 %0, {%1,%2};" : "=l"(qf2[q][d / 2]) : "f"(a), "f"(b));
            }
    };

    unsigned long long nvis = 0;
    set_qf();
    for (int c = c0; c < c0 + Q && c < nchunks; c++) { stage(load_chunk<D, WEIGHTED>(jd, c, lane, n), true); knn_body(std::integral_constant<int, 0>{}, c - c0, true); nvis += Q; }

    // regroup the warp's queries by the k-th distance seen so far (see ksg_sweep_kernel): slot Q-1 gets the sparsest
    if (Q > 1) {
        double* kb = &bd[0][0];                                    // 32*NPA >= 32*Q doubles
        static_assert(NPA >= Q, "scratch too small for the regroup");
        double key[Q]; int pos[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) key[q] = qok[q] ? L[q][KL - 1] : INFINITY;
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
        for (int e = 0; e < KL; e++) {
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
        set_qf();
    }

    {   // outward sweep: whichever side still has a candidate that some query needs (key distance below its k-th distance)
        int cl = c0 - 1, cr = c0 + Q;
        bool toggle = false;
        Chunk<D, WEIGHTED> Lc, Rc;
        if (cl >= 0) Lc = load_chunk<D, WEIGHTED>(jd, cl, lane, n);
        if (cr < nchunks) Rc = load_chunk<D, WEIGHTED>(jd, cr, lane, n);
        auto need = [&](double znext) {
            unsigned f = 0u;
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                const double dk = qc[q][KEY] - znext;
                f |= ((UMI ? __dmul_rn(dk, dk) : fabs(dk)) < L[q][KL - 1]) ? (1u << q) : 0u;
            }
            return f;
        };
        while (true) {
            unsigned al = 0, ar = 0;
            if (cl >= 0)      al = __reduce_or_sync(0xffffffffu, need(__shfl_sync(0xffffffffu, Lc.v[KEY], 31)));
            if (cr < nchunks) ar = __reduce_or_sync(0xffffffffu, need(__shfl_sync(0xffffffffu, Rc.v[KEY], 0)));
            if (!al && !ar) break;
            const bool left = al && (!ar || toggle);
            toggle = !toggle;
            const unsigned act = left ? al : ar;
            // stage straight from the prefetched registers, then reuse them for the next prefetch (issued before the scan, so
            // its latency still hides behind it): no chunk-to-chunk register copies (they were ~5 % of the instructions)
            // (measured, tools/ab.py r2: -1.5 % on C3, -0.8 % on C2; the weighted kernels lose 3.5 % with it and keep the copy)
            if constexpr (WEIGHTED) {
                Chunk<D, WEIGHTED> cur;
                if (left) { cur = Lc; --cl; if (cl >= 0) Lc = load_chunk<D, WEIGHTED>(jd, cl, lane, n); }
                else      { cur = Rc; ++cr; if (cr < nchunks) Rc = load_chunk<D, WEIGHTED>(jd, cr, lane, n); }
                stage(cur, true);
            } else {
                if (left) { stage(Lc, true); --cl; if (cl >= 0) Lc = load_chunk<D, WEIGHTED>(jd, cl, lane, n); }
                else      { stage(Rc, true); ++cr; if (cr < nchunks) Rc = load_chunk<D, WEIGHTED>(jd, cr, lane, n); }
            }
            const bool all = (act & ((1u << (Q - 1)) - 1u)) != 0u || Q == 1;
            knn_chunk(all);
            nvis += all ? Q : 1;
        }
    }

    double eps[Q], rad[Q];                                         // rad: what a coordinate difference is tested against
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        eps[q] = L[q][KL - 1]; rad[q] = eps[q];
        if (UMI) { eps[q] = __dsqrt_rn(eps[q]); rad[q] = __dmul_rn(eps[q], eps[q]); }      // scipy: r = sqrt(d2_k); ball test diff^2 <= r * r
    }
    auto inball = [&](double diff, int q) { return UMI ? (__dmul_rn(diff, diff) <= rad[q]) : (fabs(diff) <= rad[q]); };

    // ---------------- pass 2: rank windows, then x / y (/ extra z) bits per candidate ----------------
    int lo[Q], hi[Q];
    {
        // lo = first rank r with  key_r > key_i  or  |key_i - key_r| <= eps   (monotone false -> true in r)
        // hi = last rank r with   key_r < key_i  or  |key_i - key_r| <= eps   (monotone true -> false in r)
        int la[Q], lb[Q], ha[Q], hb[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) { la[q] = 0; lb[q] = n; ha[q] = 0; hb[q] = n; }
        const double* __restrict__ kcol = jd.col[KEY];
        const int iters = 33 - __clz(n);
        for (int it = 0; it < iters; it++) {
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                if (la[q] < lb[q]) {
                    const int mid = (la[q] + lb[q]) >> 1;
                    const double zr = kcol[perm ? perm[mid] : mid];
                    const bool A = (zr > qc[q][KEY]) || inball(qc[q][KEY] - zr, q);
                    if (A) lb[q] = mid; else la[q] = mid + 1;
                }
                if (ha[q] < hb[q]) {
                    const int mid = (ha[q] + hb[q]) >> 1;
                    const double zr = kcol[perm ? perm[mid] : mid];
                    const bool B = (zr < qc[q][KEY]) || inball(qc[q][KEY] - zr, q);
                    if (B) ha[q] = mid + 1; else hb[q] = mid;
                }
            }
        }
        #pragma unroll
        for (int q = 0; q < Q; q++) { lo[q] = la[q]; hi[q] = ha[q] - 1; }
    }
    // window of a query in an ascending column `sv` of m values: first / last index whose value is within eps of v
    auto window_in = [&](const double* __restrict__ sv, int m, const double (&v)[Q], int (&wl)[Q], int (&wh)[Q]) {
        int la[Q], lb[Q], ha[Q], hb[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) { la[q] = 0; lb[q] = m; ha[q] = 0; hb[q] = m; }
        const int iters = 33 - __clz(m);
        for (int it = 0; it < iters; it++) {
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                if (la[q] < lb[q]) {
                    const int mid = (la[q] + lb[q]) >> 1;
                    const double xr = sv[mid];
                    const bool A = (xr > v[q]) || inball(v[q] - xr, q);
                    if (A) lb[q] = mid; else la[q] = mid + 1;
                }
                if (ha[q] < hb[q]) {
                    const int mid = (ha[q] + hb[q]) >> 1;
                    const double xr = sv[mid];
                    const bool B = (xr < v[q]) || inball(v[q] - xr, q);
                    if (B) ha[q] = mid + 1; else hb[q] = mid;
                }
            }
        }
        #pragma unroll
        for (int q = 0; q < Q; q++) { wl[q] = la[q]; wh[q] = ha[q] - 1; }
    };
    constexpr bool RANKED = SWEEPW_RANK != 0;
    int nxm[Q];                                                    // mi: marginal count of x over ALL points (incl. self)
    #pragma unroll
    for (int q = 0; q < Q; q++) nxm[q] = 0;
    float hw[Q][NPA];                                              // rank windows of the non-key coordinates: half-widths (centres go to qf2)
    #pragma unroll
    for (int q = 0; q < Q; q++)
        #pragma unroll
        for (int d = 0; d < NPA; d++) hw[q][d] = -1.f;
    if constexpr (RANKED || DZ == 0) {
        float ctr[Q][NPA];
        #pragma unroll
        for (int q = 0; q < Q; q++)
            #pragma unroll
            for (int d = 0; d < NPA; d++) ctr[q][d] = 0.f;
        #pragma unroll
        for (int c = 0; c < (RANKED ? NP : 1); c++) {
            double v[Q]; int wl[Q], wh[Q];
            #pragma unroll
            for (int q = 0; q < Q; q++) v[q] = qc[q][c];
            window_in(jd.srt[c], jd.srt_n[c], v, wl, wh);
            #pragma unroll
            for (int q = 0; q < Q; q++) {
                ctr[q][c] = 0.5f * (float)(wl[q] + wh[q]);         // exact: srt_n <= 2^21
                hw[q][c]  = 0.5f * (float)(wh[q] - wl[q]);
                if (DZ == 0 && c == 0) nxm[q] = wh[q] - wl[q] + 1;
            }
        }
        if constexpr (RANKED) {
            #pragma unroll
            for (int q = 0; q < Q; q++)
                #pragma unroll
                for (int d = 0; d < NPA; d += 2)
                    asm("mov.b64 %0, {%1,%2};" : "=l"(qf2[q][d / 2]) : "f"(ctr[q][d]), "f"(ctr[q][d + 1]));
        }
    }
    int nxyz[Q], nxz[Q], nyz[Q], nz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) { nxyz[q] = nxz[q] = nyz[q] = 0; nz[q] = (DZ <= 1) ? hi[q] - lo[q] + 1 : 0; }

    int clo[Q], chi[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        clo[q] = __reduce_min_sync(0xffffffffu, lo[q] >> 5);
        chi[q] = __reduce_max_sync(0xffffffffu, hi[q] >> 5);
    }
    int cmin = clo[0], cmax = chi[0];
    #pragma unroll
    for (int q = 1; q < Q; q++) { cmin = min(cmin, clo[q]); cmax = max(cmax, chi[q]); }

    double wyz[Q], wz[Q];
    #pragma unroll
    for (int q = 0; q < Q; q++) wyz[q] = wz[q] = 0.0;
#if SWEEPW_DIAG
    unsigned long long useful = 0;
#endif

    // per-chunk epilogue shared by both forms of the scan: masks -> counts
    auto tally = [&](int q, int base, unsigned zw, unsigned mx, unsigned my, unsigned mz) {
#if SWEEPW_DIAG
        useful += __popc(zw);                                      // candidates of this chunk inside the lane's own key window
#endif
        if (DZ == 0) {
            nxz[q] += __popc(zw & mx);                             // n_xy: x-ball inside the y-window
        } else if (!WEIGHTED) {
            unsigned z = zw;
            if (DZ > 1) { z &= mz; nz[q] += __popc(z); }
            const unsigned xw = z & mx;
            nxz[q]  += __popc(xw);
            nyz[q]  += __popc(z & my);
            nxyz[q] += __popc(xw & my);
        } else {                                                   // my (and mz) already carry the window
            const unsigned z = (DZ > 1) ? mz : zw;
            if (DZ > 1) nz[q] += __popc(z);
            nxz[q]  += __popc(z & mx);
            nyz[q]  += __popc(my);
            nxyz[q] += __popc(my & mx);
            if (DZ == 1) {
                const int a = max(lo[q] - base, 0), b = min(hi[q] - base, 31);
                if (b >= a) wz[q] += bw[33 + b] - bw[32 + a];
            }
        }
    };

    // FP64 form (SWEEPW_RANK=0): coordinate differences against eps
    auto count_body = [&](auto q_first, int base) {
        constexpr int Q0 = decltype(q_first)::value;
        unsigned mx[Q], my[Q], mz[Q], zw[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) { mx[q] = my[q] = 0u; mz[q] = 0u; zw[q] = range_mask(lo[q] - base, hi[q] - base); }
        #pragma unroll 1
        for (int jb = 0; jb < 32; jb += UNRC) {
            unsigned bx[Q], by[Q], bz[Q], zb[Q];
            #pragma unroll
            for (int q = 0; q < Q; q++) { bx[q] = by[q] = 0u; bz[q] = 0u; zb[q] = zw[q] >> jb; }
            #pragma unroll
            for (int t = 0; t < UNRC; t++) {
                const int j = jb + t;
                double cc[NPA];
                #pragma unroll
                for (int d = 0; d < NPA; d += 2) { const double2 v = *reinterpret_cast<const double2*>(&bd[j][d]); cc[d] = v.x; cc[d + 1] = v.y; }
                double wj_ = 0.0;
                if (WEIGHTED) wj_ = bw[j];
                const unsigned bit = 1u << t;
                #pragma unroll
                for (int q = Q0; q < Q; q++) {
                    ball_bit(qc[q][0] - cc[0], eps[q], bit, bx[q]);
                    if constexpr (DZ == 0) {
                        // mi: the key is y itself; only x is tested inside the y-window
                    } else if constexpr (!WEIGHTED) {
                        ball_bit(qc[q][1] - cc[1], eps[q], bit, by[q]);
                        if constexpr (DZ == 2) ball_bit(qc[q][2] - cc[2], eps[q], bit, bz[q]);
                        if constexpr (DZ == 3) ball_bit2(qc[q][2] - cc[2], qc[q][3] - cc[3], eps[q], bit, bz[q]);
                    } else if constexpr (DZ == 1) {
                        wball_y(qc[q][1] - cc[1], eps[q], wj_, zb[q], bit, by[q], wyz[q]);
                    } else if constexpr (DZ == 2) {
                        wball_zy(qc[q][2] - cc[2], 0.0, false, qc[q][1] - cc[1], eps[q], wj_, zb[q], bit, bz[q], by[q], wz[q], wyz[q]);
                    } else {
                        wball_zy(qc[q][2] - cc[2], qc[q][3] - cc[3], true, qc[q][1] - cc[1], eps[q], wj_, zb[q], bit, bz[q], by[q], wz[q], wyz[q]);
                    }
                }
            }
            #pragma unroll
            for (int q = Q0; q < Q; q++) { mx[q] |= bx[q] << jb; my[q] |= by[q] << jb; if (DZ > 1) mz[q] |= bz[q] << jb; }
        }
        #pragma unroll
        for (int q = Q0; q < Q; q++) tally(q, base, zw[q], mx[q], my[q], mz[q]);
    };

    // rank form: the staged candidates are their ranks in the non-key columns (bf), the queries their window centres (qf2)
    auto count_body_rank = [&](auto q_first, int base) {
        constexpr int Q0 = decltype(q_first)::value;
        unsigned mx[Q], my[Q], mz[Q], zw[Q];
        #pragma unroll
        for (int q = 0; q < Q; q++) { mx[q] = my[q] = 0u; mz[q] = 0u; zw[q] = range_mask(lo[q] - base, hi[q] - base); }
        #pragma unroll 1
        for (int jb = 0; jb < 32; jb += UNRC) {
            unsigned bx[Q], by[Q], bz[Q], zb[Q];
            #pragma unroll
            for (int q = 0; q < Q; q++) { bx[q] = by[q] = 0u; bz[q] = 0u; zb[q] = zw[q] >> jb; }
            #pragma unroll
            for (int t = 0; t < UNRC; t++) {
                const int j = jb + t;
                unsigned long long cf2[NPA / 2];
                if constexpr (NPA == 2) { cf2[0] = *reinterpret_cast<const unsigned long long*>(&bf[j][0]); }
                else { const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&bf[j][0]); cf2[0] = v.x; cf2[NPA / 2 - 1] = v.y; }
                double wj_ = 0.0;
                if (WEIGHTED) wj_ = bw[j];
                const unsigned bit = 1u << t;
                #pragma unroll
                for (int q = Q0; q < Q; q++) {
                    float df[NPA];
                    #pragma unroll
                    for (int d = 0; d < NPA; d += 2) {
                        unsigned long long dd;
                        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dd) : "l"(qf2[q][d / 2]), "l"(cf2[d / 2]));
                        asm("mov.b64 {%0,%1}, %2;" : "=f"(df[d]), "=f"(df[d + 1]) : "l"(dd));
                    }
                    if constexpr (DZ == 0) {
                        rank_bit1(df[0], hw[q][0], bit, bx[q]);
                    } else if constexpr (!WEIGHTED) {
                        rank_bit2(df[0], df[1], hw[q][0], hw[q][1], bit, bx[q], by[q]);
                        if constexpr (DZ == 2) rank_bit1(df[NP > 2 ? 2 : 0], hw[q][NP > 2 ? 2 : 0], bit, bz[q]);
                        if constexpr (DZ == 3) rank_bit_and2(df[NP > 2 ? 2 : 0], df[NP > 3 ? 3 : 0], hw[q][NP > 2 ? 2 : 0], hw[q][NP > 3 ? 3 : 0], bit, bz[q]);
                    } else if constexpr (DZ == 1) {
                        rank_bit1(df[0], hw[q][0], bit, bx[q]);
                        rank_wy(df[1], hw[q][1], wj_, zb[q], bit, by[q], wyz[q]);
                    } else if constexpr (DZ == 2) {
                        rank_bit1(df[0], hw[q][0], bit, bx[q]);
                        rank_wzy(df[NP > 2 ? 2 : 0], hw[q][NP > 2 ? 2 : 0], 0.f, 0.f, false, df[1], hw[q][1], wj_, zb[q], bit, bz[q], by[q], wz[q], wyz[q]);
                    } else {
                        rank_bit1(df[0], hw[q][0], bit, bx[q]);
                        rank_wzy(df[NP > 2 ? 2 : 0], hw[q][NP > 2 ? 2 : 0], df[NP > 3 ? 3 : 0], hw[q][NP > 3 ? 3 : 0], true, df[1], hw[q][1], wj_, zb[q], bit,
                                 bz[q], by[q], wz[q], wyz[q]);
                    }
                }
            }
            #pragma unroll
            for (int q = Q0; q < Q; q++) { mx[q] |= bx[q] << jb; my[q] |= by[q] << jb; if (DZ > 1) mz[q] |= bz[q] << jb; }
        }
        #pragma unroll
        for (int q = Q0; q < Q; q++) tally(q, base, zw[q], mx[q], my[q], mz[q]);
    };

    // candidates of the rank form: ranks of the NP non-key coordinates (NaN beyond n: never inside a window) and the weight
    struct RChunk { float r[NPA]; double w; };
    auto load_rchunk = [&](int chunk) {
        RChunk c;
        const int rank = chunk * 32 + lane;
        const bool ok = rank < n;
        const int idx = ok ? (perm ? perm[rank] : rank) : 0;
        #pragma unroll
        for (int d = 0; d < NPA; d++) c.r[d] = (ok && d < NP) ? jd.rk[d < NP ? d : 0][idx] : __int_as_float(0x7fc00000);
        c.w = 0.0;
        if (WEIGHTED) c.w = ok ? jd.w[idx] : 0.0;
        return c;
    };
    auto stage_r = [&](const RChunk& c) {
        __syncwarp();
        #pragma unroll
        for (int d = 0; d < NPA; d += 2) { float2 v; v.x = c.r[d]; v.y = c.r[d + 1]; *reinterpret_cast<float2*>(&bf[lane][d]) = v; }
        if (WEIGHTED) {
            bw[lane] = c.w;
            if (DZ == 1) {                                         // exclusive prefix sums of the chunk's weights: bw[32 + j] = w_0 + .. + w_{j-1}
                double ps = c.w;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, ps, o); if (lane >= o) ps += t; }
                bw[33 + lane] = ps;
                if (lane == 0) bw[32] = 0.0;
            }
        }
        __syncwarp();
    };

    unsigned long long nvis2 = 0;
    if constexpr (UMI) {
        // nothing to scan: both marginal balls are rank ranges
    } else if constexpr (RANKED) {
        RChunk nxt = load_rchunk(cmin), cur;
        for (int c = cmin; c <= cmax; c++) {
            cur = nxt;
            if (c < cmax) nxt = load_rchunk(c + 1);
            stage_r(cur);
            bool all = Q == 1;
            #pragma unroll
            for (int q = 0; q < Q - 1; q++) all |= (c >= clo[q]) & (c <= chi[q]);
            if (all) count_body_rank(std::integral_constant<int, 0>{}, c * 32);
            else     count_body_rank(std::integral_constant<int, Q - 1>{}, c * 32);
            nvis2 += all ? Q : 1;
        }
    }
#if SWEEPW_CPASYNC
    else if (!WEIGHTED && perm == nullptr) {
        // cp.async pipeline: chunk c + 1 is copied global -> shared (8 bytes per lane and coordinate, no registers) while
        // chunk c is scanned; lanes beyond n copy nothing (their slots keep stale data: never inside a key window)
        double (*buf[2])[NPA] = { sd[wid], sd2[WEIGHTED ? 0 : wid] };
        auto issue = [&](int c, double (*dst)[NPA]) {
            const int rank = c * 32 + lane;
            if (rank < n) {
                #pragma unroll
                for (int d = 0; d < NP; d++) {
                    const unsigned sa = (unsigned)__cvta_generic_to_shared(&dst[lane][d]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(sa), "l"(jd.col[d] + rank));
                }
            }
            asm volatile("cp.async.commit_group;");
        };
        __syncwarp();
        issue(cmin, buf[0]);
        int cur_b = 0;
        for (int c = cmin; c <= cmax; c++) {
            if (c < cmax) issue(c + 1, buf[cur_b ^ 1]); else asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 1;");
            __syncwarp();
            bd = buf[cur_b];
            bool all = Q == 1;
            #pragma unroll
            for (int q = 0; q < Q - 1; q++) all |= (c >= clo[q]) & (c <= chi[q]);
            if (all) count_body(std::integral_constant<int, 0>{}, c * 32);
            else     count_body(std::integral_constant<int, Q - 1>{}, c * 32);
            nvis2 += all ? Q : 1;
            __syncwarp();
            cur_b ^= 1;
        }
        asm volatile("cp.async.wait_group 0;");
    }
#endif
    else {
        Chunk<D, WEIGHTED> ch = load_chunk<D, WEIGHTED>(jd, cmin, lane, n);
        for (int c = cmin; c <= cmax; c++) {
            if constexpr (WEIGHTED) {
                const Chunk<D, WEIGHTED> cur = ch;
                if (c < cmax) ch = load_chunk<D, WEIGHTED>(jd, c + 1, lane, n);
                stage(cur, false);
            } else {
                stage(ch, false);
                if (c < cmax) ch = load_chunk<D, WEIGHTED>(jd, c + 1, lane, n);
            }
            bool all = Q == 1;
            #pragma unroll
            for (int q = 0; q < Q - 1; q++) all |= (c >= clo[q]) & (c <= chi[q]);
            if (all) count_body(std::integral_constant<int, 0>{}, c * 32);
            else     count_body(std::integral_constant<int, Q - 1>{}, c * 32);
            nvis2 += all ? Q : 1;
        }
    }

    // ---------------- digamma terms + per-warp partial (fixed shuffle tree) ----------------
    double acc = 0.0;
    #pragma unroll
    for (int q = 0; q < Q; q++) {
        if (qok[q]) {
            const int a = nxyz[q] - 1, b = nxz[q] - 1, c = nyz[q] - 1, d = nz[q] - 1;
            const int64_t o = jd.qoff + qidx[q];
            if (UMI) {
                // information_estimators.py:268-276: n_x = |ball_x| - 1, W_y = sum of the weights over ball_y minus w_i;
                // term = w_i * log(n_x) + w_i * log(W_y); math.log raises for arguments <= 0 -> domain flag (visited[3])
                const double wi = qw[q];
                const int nx = nxm[q] - 1;
                const double Wy = (jd.wcum[hi[q] + 1] - jd.wcum[lo[q]]) - wi;
                if ((nx <= 0 || Wy <= 0.0) && visited) atomicOr(reinterpret_cast<unsigned int*>(visited + 3), 1u);
                acc += wi * log((double)nx) + wi * log(Wy);
                if (po.wsum) { po.wsum[2 * o] = Wy; po.wsum[2 * o + 1] = 0.0; }
            } else if (DZ == 0) {
                // information_estimators.py:102-107: psi(N) + psi(n_xy) - psi(n_x) - psi(n_y); slots as in the dense kernel
                acc += psi_int(psi_tab, n) + psi_int(psi_tab, b) - psi_int(psi_tab, nxm[q] - 1) - psi_int(psi_tab, d);
            } else if (WEIGHTED) {
                const double wi = qw[q];
                const double Wyz = wyz[q] - wi, Wz = wz[q] - wi;
                acc += wi * psi_int(psi_tab, a) + wi * -psi_int(psi_tab, b) + wi * -digamma_f64(Wyz) + wi * digamma_f64(Wz);
                if (po.wsum) { po.wsum[2 * o] = Wyz; po.wsum[2 * o + 1] = Wz; }
            } else {
                acc += psi_int(psi_tab, a) - psi_int(psi_tab, b) - psi_int(psi_tab, c) + psi_int(psi_tab, d);
            }
            if (po.eps) po.eps[o] = eps[q];
            if (po.cnt) {
                int32_t* oc = po.cnt + 4 * o;
                if (UMI)          { oc[0] = -1; oc[1] = nxm[q] - 1; oc[2] = d; oc[3] = n - 1; }  // (joint ball not counted), n_x, n_y, (all)
                else if (DZ == 0) { oc[0] = b; oc[1] = nxm[q] - 1; oc[2] = d; oc[3] = n - 1; }  // n_xy, n_x, n_y, (all)
                else         { oc[0] = a; oc[1] = b; oc[2] = c; oc[3] = d; }
            }
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        partial[(size_t)job * warps_per_job + wj] = acc;
        if (visited) { atomicAdd(visited, nvis * 32ull * 32ull); atomicAdd(visited + 1, nvis2 * 32ull * 32ull); }
    }
#if SWEEPW_DIAG
    {   // diagnostics build only: how many of the count-pass visits fall inside the visiting lane's own window (visited[2], unweighted)
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) useful += __shfl_down_sync(0xffffffffu, useful, o);
        if (lane == 0 && visited && !WEIGHTED) atomicAdd(visited + 2, useful);
    }
#endif
}

}  // namespace scribe
