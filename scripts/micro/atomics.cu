// micro-benchmark: FP64 atomicAdd throughput on B200 under different address distributions.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/atomics scripts/micro/atomics.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ unsigned hash32(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

// each thread performs `per` atomics; address = table[hash % nhot] with prob hot_frac (x/1024), else random in [0,n)
template<int RET, int F64>
__global__ void k(double* r, unsigned long long* ri, unsigned n, unsigned nhot, unsigned hot1024, int per, double* sink){
    unsigned t = blockIdx.x*blockDim.x+threadIdx.x;
    double acc=0;
    for(int i=0;i<per;++i){
        unsigned h = hash32(t*977u + i*7919u + 13u);
        unsigned h2 = hash32(h+0x9e3779b9u);
        unsigned idx = ((h & 1023u) < hot1024) ? (h2 % nhot) * 37u % n : h2 % n;
        if (F64) { if(RET) acc += atomicAdd(&r[idx], 1e-9); else atomicAdd(&r[idx], 1e-9); }
        else { if(RET) acc += (double)atomicAdd(&ri[idx], 1ull); else atomicAdd(&ri[idx], 1ull); }
    }
    if(RET && acc==12345.678) *sink=acc;
}
template<int RET,int F64>
float run(double* r, unsigned long long* ri, unsigned n, unsigned nhot, unsigned hot1024, int per, int blocks, double* sink){
    cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<RET,F64><<<blocks,256>>>(r,ri,n,nhot,hot1024,per,sink); // warm
    cudaEventRecord(a);
    k<RET,F64><<<blocks,256>>>(r,ri,n,nhot,hot1024,per,sink);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms,a,b); return ms;
}
int main(){
    unsigned n = 1u<<20; // 8 MB of doubles: L2 resident
    double* r; unsigned long long* ri; double* sink;
    CK(cudaMalloc(&r, n*8)); CK(cudaMalloc(&ri, n*8)); CK(cudaMalloc(&sink,8));
    CK(cudaMemset(r,0,n*8)); CK(cudaMemset(ri,0,n*8));
    int blocks=148*4, per=64; double total = (double)blocks*256*per;
    printf("total atomics per launch %.0f (blocks %d x256 x%d)\n", total, blocks, per);
    struct C{unsigned nhot,hot;} cases[]={{1,0},{1,1024},{1,8},{1,32},{1,128},{16,1024},{16,128},{256,1024},{256,256},{4096,1024}};
    for(auto c: cases){
        float a=run<1,1>(r,ri,n,c.nhot,c.hot,per,blocks,sink);
        float b=run<0,1>(r,ri,n,c.nhot,c.hot,per,blocks,sink);
        float d=run<1,0>(r,ri,n,c.nhot,c.hot,per,blocks,sink);
        printf("nhot %5u hotfrac %5.3f : f64 ret %8.3f ms (%7.2f /ns)  f64 red %8.3f ms (%7.2f /ns)  u64 ret %8.3f ms (%7.2f /ns)\n",
            c.nhot, c.hot/1024.0, a, total/a*1e-6, b, total/b*1e-6, d, total/d*1e-6);
    }
    // DRAM-sized table
    unsigned nbig = 1u<<28; double* rb; CK(cudaMalloc(&rb,(size_t)nbig*8)); CK(cudaMemset(rb,0,(size_t)nbig*8));
    float a=run<1,1>(rb,ri,nbig,1,0,per,blocks,sink); float b=run<0,1>(rb,ri,nbig,1,0,per,blocks,sink);
    printf("2 GiB table random: f64 ret %8.3f ms (%7.2f /ns) red %8.3f ms (%7.2f /ns)\n", a,total/a*1e-6,b,total/b*1e-6);
    return 0;
}
