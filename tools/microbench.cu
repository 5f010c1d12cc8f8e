// Pipe-rate microbenchmark for sm_100a: measures warp-instruction issue rates of the
// instruction mixes the KSG kernels are built from, to fix the ALU roofline denominator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

template<int OP> __global__ void __launch_bounds__(256) k_pipe(float* out, float seed, int iters) {
    float a[CH]; double d[CH]; int n[CH];
    #pragma unroll
    for (int c=0;c<CH;c++){ a[c]=seed*(threadIdx.x+c+1); d[c]=a[c]; n[c]=threadIdx.x+c; }
    float b = seed*3.f, e = seed*0.5f; double db=b, de=e;
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int u=0;u<8;u++) {
            #pragma unroll
            for (int c=0;c<CH;c++) {
                if (OP==0) a[c] = a[c] + b;                       // FADD
                if (OP==1) a[c] = fmaf(a[c], b, e);               // FFMA
                if (OP==2) a[c] = fmaxf(a[c], b) ;                // FMNMX (degenerates? b const)
                if (OP==3) { asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(b), "f"(e)); }  // FMNMX3
                if (OP==4) { asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(n[c]) : "f"(a[c]), "f"(b)); } // FSETP + @P IADD
                if (OP==5) { asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; @p add.f32 %0, %0, 0f3F800000;}" : "+f"(a[(c+1)%CH]) : "f"(a[c]), "f"(b)); } // FSETP + @P FADD
                if (OP==6) n[c] = n[c] + (n[c]>>3) ;              // int ops
                if (OP==7) d[c] = d[c] + db;                      // DADD
                if (OP==8) { asm volatile("{.reg .pred p; setp.le.f64 p, %1, %2; @p add.s32 %0, %0, 1;}" : "+r"(n[c]) : "d"(d[c]), "d"(db)); } // DSETP + @P IADD
                if (OP==9) d[c] = fma(d[c], db, de);              // DFMA
                if (OP==10) { asm volatile("max.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)); }  // FMNMX via asm
                if (OP==11) { asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; selp.f32 %0, %0, %2, p;}" : "+f"(a[c]) : "f"(a[c]), "f"(b)); }
            }
        }
    }
    float s=0; 
    #pragma unroll
    for (int c=0;c<CH;c++) s += a[c] + (float)d[c] + n[c];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// ---- prototype inner loops: thread-per-query, Q queries per thread, candidates broadcast from smem ----
// candidates stored AoS float4 (x,y,z,pad). Returns something data-dependent so nothing is elided.
constexpr int TILE = 2048;

template<int Q> __global__ void __launch_bounds__(256) k_knn32(const float4* __restrict__ cand, int ncand, float* out) {
    __shared__ float4 tile[TILE];
    float qx[Q], qy[Q], qz[Q], thr[Q];
    #pragma unroll
    for (int q=0;q<Q;q++){ float4 p = cand[(blockIdx.x*256*Q + q*256 + threadIdx.x) % ncand]; qx[q]=p.x; qy[q]=p.y; qz[q]=p.z; thr[q]=1e-3f; }
    int hits = 0;
    for (int base=0; base<ncand; base+=TILE) {
        __syncthreads();
        for (int t=threadIdx.x; t<TILE; t+=256) tile[t] = cand[base+t];
        __syncthreads();
        #pragma unroll 1
        for (int j=0;j<TILE;j+=4) {
            float m[Q];
            #pragma unroll
            for (int jj=0;jj<4;jj++) {
                float4 c = tile[j+jj];
                #pragma unroll
                for (int q=0;q<Q;q++) {
                    float dx=qx[q]-c.x, dy=qy[q]-c.y, dz=qz[q]-c.z, mm;
                    asm("max.f32 %0, %1, %2, %3;" : "=f"(mm) : "f"(fabsf(dx)), "f"(fabsf(dy)), "f"(fabsf(dz)));
                    m[q] = jj==0 ? mm : fminf(m[q], mm);
                }
            }
            bool any=false;
            #pragma unroll
            for (int q=0;q<Q;q++) any |= (m[q] < thr[q]);
            if (any) {   // rare path stand-in for the top-k insert
                #pragma unroll
                for (int q=0;q<Q;q++) if (m[q]<thr[q]) { thr[q] = fmaxf(m[q], thr[q]*0.999f); hits++; }
            }
        }
    }
    float s=hits;
    #pragma unroll
    for (int q=0;q<Q;q++) s+=thr[q];
    out[blockIdx.x*256+threadIdx.x]=s;
}

// count pass, FP32. ACC=0: predicated IADD, ACC=1: predicated FADD
template<int Q, int ACC> __global__ void __launch_bounds__(256) k_cnt32(const float4* __restrict__ cand, int ncand, float* out) {
    __shared__ float4 tile[TILE];
    float qx[Q], qy[Q], qz[Q], eps[Q];
    float fz[Q], fxz[Q], fyz[Q], fxyz[Q]; int iz[Q], ixz[Q], iyz[Q], ixyz[Q];
    #pragma unroll
    for (int q=0;q<Q;q++){ float4 p = cand[(blockIdx.x*256*Q + q*256 + threadIdx.x) % ncand]; qx[q]=p.x; qy[q]=p.y; qz[q]=p.z; eps[q]=0.05f+0.0001f*q;
        fz[q]=fxz[q]=fyz[q]=fxyz[q]=0.f; iz[q]=ixz[q]=iyz[q]=ixyz[q]=0; }
    for (int base=0; base<ncand; base+=TILE) {
        __syncthreads();
        for (int t=threadIdx.x; t<TILE; t+=256) tile[t] = cand[base+t];
        __syncthreads();
        #pragma unroll 4
        for (int j=0;j<TILE;j++) {
            float4 c = tile[j];
            #pragma unroll
            for (int q=0;q<Q;q++) {
                float dx=qx[q]-c.x, dy=qy[q]-c.y, dz=qz[q]-c.z;
                if (ACC==0) asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f32 ax,ay,az;\n abs.f32 ax,%4; abs.f32 ay,%5; abs.f32 az,%6;\n"
                    " setp.le.f32 pz, az, %7;\n setp.le.and.f32 pxz, ax, %7, pz;\n setp.le.and.f32 pyz, ay, %7, pz;\n setp.le.and.f32 pxyz, ay, %7, pxz;\n"
                    " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n}"
                    : "+r"(iz[q]), "+r"(ixz[q]), "+r"(iyz[q]), "+r"(ixyz[q]) : "f"(dx), "f"(dy), "f"(dz), "f"(eps[q]));
                else asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f32 ax,ay,az;\n abs.f32 ax,%4; abs.f32 ay,%5; abs.f32 az,%6;\n"
                    " setp.le.f32 pz, az, %7;\n setp.le.and.f32 pxz, ax, %7, pz;\n setp.le.and.f32 pyz, ay, %7, pz;\n setp.le.and.f32 pxyz, ay, %7, pxz;\n"
                    " @pz add.f32 %0,%0,0f3F800000;\n @pxz add.f32 %1,%1,0f3F800000;\n @pyz add.f32 %2,%2,0f3F800000;\n @pxyz add.f32 %3,%3,0f3F800000;\n}"
                    : "+f"(fz[q]), "+f"(fxz[q]), "+f"(fyz[q]), "+f"(fxyz[q]) : "f"(dx), "f"(dy), "f"(dz), "f"(eps[q]));
            }
        }
    }
    float s=0;
    #pragma unroll
    for (int q=0;q<Q;q++) s+= fz[q]+fxz[q]*3+fyz[q]*5+fxyz[q]*7 + iz[q]+ixz[q]*3+iyz[q]*5+ixyz[q]*7;
    out[blockIdx.x*256+threadIdx.x]=s;
}

// count pass, FP64 exact (candidates as double SoA-in-AoS: x,y,z,pad)
template<int Q> __global__ void __launch_bounds__(256) k_cnt64(const float4* __restrict__ cand32, int ncand, float* out) {
    __shared__ double tx[TILE/2], ty[TILE/2], tz[TILE/2];
    double qx[Q], qy[Q], qz[Q], eps[Q]; int iz[Q], ixz[Q], iyz[Q], ixyz[Q];
    #pragma unroll
    for (int q=0;q<Q;q++){ float4 p = cand32[(blockIdx.x*256*Q + q*256 + threadIdx.x) % ncand]; qx[q]=p.x; qy[q]=p.y; qz[q]=p.z; eps[q]=0.05+0.0001*q; iz[q]=ixz[q]=iyz[q]=ixyz[q]=0; }
    for (int base=0; base<ncand; base+=TILE/2) {
        __syncthreads();
        for (int t=threadIdx.x; t<TILE/2; t+=256) { float4 c=cand32[base+t]; tx[t]=c.x; ty[t]=c.y; tz[t]=c.z; }
        __syncthreads();
        #pragma unroll 4
        for (int j=0;j<TILE/2;j++) {
            double cx=tx[j], cy=ty[j], cz=tz[j];
            #pragma unroll
            for (int q=0;q<Q;q++) {
                double dx=qx[q]-cx, dy=qy[q]-cy, dz=qz[q]-cz;
                asm("{\n .reg .pred pz,pxz,pyz,pxyz;\n .reg .f64 ax,ay,az;\n abs.f64 ax,%4; abs.f64 ay,%5; abs.f64 az,%6;\n"
                    " setp.le.f64 pz, az, %7;\n setp.le.and.f64 pxz, ax, %7, pz;\n setp.le.and.f64 pyz, ay, %7, pz;\n setp.le.and.f64 pxyz, ay, %7, pxz;\n"
                    " @pz add.s32 %0,%0,1;\n @pxz add.s32 %1,%1,1;\n @pyz add.s32 %2,%2,1;\n @pxyz add.s32 %3,%3,1;\n}"
                    : "+r"(iz[q]), "+r"(ixz[q]), "+r"(iyz[q]), "+r"(ixyz[q]) : "d"(dx), "d"(dy), "d"(dz), "d"(eps[q]));
            }
        }
    }
    float s=0;
    #pragma unroll
    for (int q=0;q<Q;q++) s+= iz[q]+ixz[q]*3+iyz[q]*5+ixyz[q]*7;
    out[blockIdx.x*256+threadIdx.x]=s;
}

template<typename F> float timeit(F f, int reps=5) {
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    float best=1e30f;
    for (int r=0;r<reps;r++){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    int clk_khz=0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"l2_bytes\":%d,\"smem_optin\":%zu}\n", p.name, nsm, clk_khz, p.l2CacheSize, p.sharedMemPerBlockOptin);
    float* out; CK(cudaMalloc(&out, sizeof(float)*nsm*8*256*4));
    const char* names[] = {"FADD","FFMA","FMNMX(c)","FMNMX3","FSETP+@IADD","FSETP+@FADD","IADD+SHF","DADD","DSETP+@IADD","DFMA","FMNMX","FSETP+SEL"};
    int per[] = {1,1,1,1,2,2,2,1,2,1,1,2};
    int blocks = nsm*8;
    #define RUN(OP) { float ms = timeit([&]{ k_pipe<OP><<<blocks,256>>>(out, 1e-9f, ITERS); }); \
        double winst = (double)blocks*8 /*warps*/ * ITERS * 8.0 * CH * per[OP]; \
        double rate = winst / (ms*1e-3) / nsm;  /* warp-instr per second per SM */ \
        printf("{\"op\":\"%s\",\"ms\":%.3f,\"warp_inst_per_s_per_sm\":%.4g,\"per_clk_per_sm_at_max\":%.3f}\n", names[OP], ms, rate, rate/(clk_khz*1e3)); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)

    int ncand = 16384;
    float4* cand; CK(cudaMalloc(&cand, sizeof(float4)*ncand));
    { float4* h=(float4*)malloc(sizeof(float4)*ncand); srand(1); for(int i=0;i<ncand;i++){h[i].x=rand()/(float)RAND_MAX;h[i].y=rand()/(float)RAND_MAX;h[i].z=rand()/(float)RAND_MAX;h[i].w=0;} CK(cudaMemcpy(cand,h,sizeof(float4)*ncand,cudaMemcpyHostToDevice)); free(h);}
    #define RUNP(NAME, KERN, Q, CTAS_PER_SM) { int nb=nsm*CTAS_PER_SM; float ms = timeit([&]{ KERN<<<nb,256>>>(cand, ncand, out); }); \
        double pairs = (double)nb*256*Q*ncand; \
        printf("{\"proto\":\"%s\",\"Q\":%d,\"ctas_per_sm\":%d,\"ms\":%.3f,\"pairs_per_s\":%.4g,\"pairs_per_clk_per_sm\":%.3f}\n", NAME, Q, CTAS_PER_SM, ms, pairs/(ms*1e-3), pairs/(ms*1e-3)/nsm/(clk_khz*1e3)); }
    RUNP("knn32", (k_knn32<1>), 1, 8) RUNP("knn32", (k_knn32<2>), 2, 8) RUNP("knn32", (k_knn32<4>), 4, 4) RUNP("knn32", (k_knn32<8>), 8, 2)
    RUNP("cnt32_iadd", (k_cnt32<1,0>), 1, 8) RUNP("cnt32_iadd", (k_cnt32<2,0>), 2, 8) RUNP("cnt32_iadd", (k_cnt32<4,0>), 4, 4) RUNP("cnt32_iadd", (k_cnt32<8,0>), 8, 2)
    RUNP("cnt32_fadd", (k_cnt32<1,1>), 1, 8) RUNP("cnt32_fadd", (k_cnt32<2,1>), 2, 8) RUNP("cnt32_fadd", (k_cnt32<4,1>), 4, 4) RUNP("cnt32_fadd", (k_cnt32<8,1>), 8, 2)
    RUNP("cnt64", (k_cnt64<1>), 1, 8) RUNP("cnt64", (k_cnt64<2>), 2, 8) RUNP("cnt64", (k_cnt64<4>), 4, 4)
    return 0;
}
