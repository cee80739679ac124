// micro-benchmark: dependent-access latencies and software grid-barrier cost on B200.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/latency scripts/micro/latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__global__ void chase_ldcg(const unsigned* next, int steps, unsigned* out, long long* cyc){
    unsigned i = threadIdx.x + blockIdx.x*blockDim.x;
    long long t0=clock64();
    for(int s=0;s<steps;++s) i = __ldcg(&next[i]);
    long long t1=clock64();
    if(threadIdx.x==0 && blockIdx.x==0){ *out=i; *cyc=t1-t0; }
}
__global__ void chase_atomic(unsigned long long* tab, const unsigned* next, int steps, unsigned* out, long long* cyc){
    unsigned i = threadIdx.x + blockIdx.x*blockDim.x;
    long long t0=clock64();
    for(int s=0;s<steps;++s){ unsigned long long o = atomicAdd(&tab[i], 1ull); i = next[i] + (unsigned)(o & 0); }
    long long t1=clock64();
    if(threadIdx.x==0 && blockIdx.x==0){ *out=i; *cyc=t1-t0; }
}
__global__ void chase_atomic_f64(double* tab, unsigned n, int steps, double* out, long long* cyc){
    unsigned i = (threadIdx.x + blockIdx.x*blockDim.x) % n;
    double acc=0; long long t0=clock64();
    for(int s=0;s<steps;++s){ double o = atomicAdd(&tab[i], 1e-9); acc+=o; i = (i*1664525u + 1013904223u + (unsigned)(o>1e300)) % n; }
    long long t1=clock64();
    if(threadIdx.x==0 && blockIdx.x==0){ *out=acc; *cyc=t1-t0; }
}
__device__ __forceinline__ void bar_arrive_release(unsigned *addr){ asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(addr) : "memory"); }
__device__ __forceinline__ unsigned bar_load_acquire(const unsigned *addr){ unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory"); return v; }
__global__ void barrier_bench(unsigned* bar, int iters, long long* cyc){
    long long t0=clock64();
    for(int g=1; g<=iters; ++g){
        __syncthreads();
        if(threadIdx.x==0){ bar_arrive_release(bar); unsigned target=g*gridDim.x; while(bar_load_acquire(bar)<target){} }
        __syncthreads();
    }
    long long t1=clock64();
    if(threadIdx.x==0 && blockIdx.x==0) *cyc=t1-t0;
}
int main(){
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0)); printf("SM clock attr %d kHz\n", clk);
    unsigned n=1u<<22; // 16 MB index table: L2 resident
    unsigned* h=(unsigned*)malloc(n*4); unsigned x=1; for(unsigned i=0;i<n;++i){ x=x*1664525u+1013904223u; h[i]=x%n; }
    unsigned *next,*out; long long* cyc; unsigned long long* tab; double* ftab; double* fout;
    CK(cudaMalloc(&next,n*4)); CK(cudaMemcpy(next,h,n*4,cudaMemcpyHostToDevice)); CK(cudaMalloc(&out,4)); CK(cudaMalloc(&cyc,8));
    CK(cudaMalloc(&tab,n*8)); CK(cudaMemset(tab,0,n*8)); CK(cudaMalloc(&ftab,n*8)); CK(cudaMemset(ftab,0,n*8)); CK(cudaMalloc(&fout,8));
    long long hc; int steps=2000;
    int cfgs[][2]={{1,32},{1,256},{148,256},{592,256}};
    for(auto&c:cfgs){
        chase_ldcg<<<c[0],c[1]>>>(next,steps,out,cyc); chase_ldcg<<<c[0],c[1]>>>(next,steps,out,cyc); CK(cudaMemcpy(&hc,cyc,8,cudaMemcpyDeviceToHost));
        printf("ld.cg chase    grid %4d x %3d: %.1f cycles/step\n",c[0],c[1],(double)hc/steps);
        chase_atomic<<<c[0],c[1]>>>(tab,next,steps,out,cyc); chase_atomic<<<c[0],c[1]>>>(tab,next,steps,out,cyc); CK(cudaMemcpy(&hc,cyc,8,cudaMemcpyDeviceToHost));
        printf("atom.u64+ld    grid %4d x %3d: %.1f cycles/step (two dependent accesses)\n",c[0],c[1],(double)hc/steps);
        chase_atomic_f64<<<c[0],c[1]>>>(ftab,n,steps,fout,cyc); chase_atomic_f64<<<c[0],c[1]>>>(ftab,n,steps,fout,cyc); CK(cudaMemcpy(&hc,cyc,8,cudaMemcpyDeviceToHost));
        printf("atom.f64 chase grid %4d x %3d: %.1f cycles/step\n",c[0],c[1],(double)hc/steps);
    }
    unsigned* bar; CK(cudaMalloc(&bar,4));
    int grids[]={148,296,592,1184};
    for(int g:grids){
        CK(cudaMemset(bar,0,4)); int iters=2000;
        void* args[]={&bar,&iters,&cyc};
        cudaError_t e=cudaLaunchCooperativeKernel((void*)barrier_bench,dim3(g),dim3(256),args,0,0);
        if(e!=cudaSuccess){printf("grid %d: %s\n",g,cudaGetErrorString(e)); cudaGetLastError(); continue;}
        CK(cudaMemcpy(&hc,cyc,8,cudaMemcpyDeviceToHost));
        printf("grid barrier %4d CTAs: %.1f cycles (%.2f us @1.965GHz)\n",g,(double)hc/iters,(double)hc/iters/1965.0);
    }
    return 0;
}
