// micro-benchmark: how many bytes must be in flight per SM for random 256-byte row gathers to reach HBM speed?
// 16 lanes x 16 B per row (the multi-source sweep's row piece), CTAS/SM x 256 threads x UNROLL pieces in flight.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/gather_mlp scripts/micro/gather_mlp.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)
__device__ __forceinline__ unsigned hash32(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }
__global__ void fill_idx(unsigned* idx, size_t m, unsigned nrows){
    for(size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x;i<m;i+=(size_t)gridDim.x*blockDim.x) idx[i]=hash32(hash32((unsigned)i*2654435761u+17u)+0x9e3779b9u)%nrows;
}
template<int UNROLL, int REGS_PAD>
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ table, const unsigned* __restrict__ idx, size_t m, double* sink){
    const unsigned g = threadIdx.x & 15;
    const size_t group = ((size_t)blockIdx.x*blockDim.x+threadIdx.x)/16, ngroups=((size_t)gridDim.x*blockDim.x)/16;
    const size_t per=(m+ngroups-1)/ngroups, lo=group*per, hi=min(m,lo+per);
    unsigned acc=0;
    for(size_t i=lo;i<hi;i+=UNROLL){
        unsigned u[UNROLL]; uint4 v[UNROLL];
#pragma unroll
        for(int k=0;k<UNROLL;++k) u[k]= (i+k<hi)? __ldcs(&idx[i+k]) : 0xffffffffu;
#pragma unroll
        for(int k=0;k<UNROLL;++k) v[k]= (u[k]!=0xffffffffu)? table[(size_t)u[k]*16+g] : make_uint4(0,0,0,0);
#pragma unroll
        for(int k=0;k<UNROLL;++k) acc+=v[k].x+v[k].y+v[k].z+v[k].w;
    }
    if(acc==12345678u) *sink=acc;
}
template<int UNROLL>
void run(const uint4* table, const unsigned* idx, size_t m, double* sink, int ctas){
    cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    gather<UNROLL,0><<<148*ctas,256>>>(table,idx,m,sink);
    cudaEventRecord(a);
    gather<UNROLL,0><<<148*ctas,256>>>(table,idx,m,sink);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms,a,b);
    const double inflight = (double)ctas*256*UNROLL*16;
    printf("ctas/SM %d unroll %2d  (%6.1f KB in flight per SM): %8.3f ms  %7.1f GB/s   implied latency %5.2f us\n", ctas, UNROLL, inflight/1024, ms, (double)m*256/ms*1e-6,
           inflight*148/((double)m*256/(ms*1e-3))*1e6);
}
int main(){
    const size_t table_bytes = 786ull<<20;   // the x array of BASELINE configs[3] with 125 sources (bf16 rows of 256 B)
    uint4* table; CK(cudaMalloc(&table, table_bytes)); CK(cudaMemset(table,0,table_bytes));
    double* sink; CK(cudaMalloc(&sink,8));
    const unsigned nrows=(unsigned)(table_bytes/256);
    const size_t m = 1ull<<26;  // 16 GiB gathered
    unsigned* idx; CK(cudaMalloc(&idx,m*4));
    fill_idx<<<148*8,256>>>(idx,m,nrows); CK(cudaDeviceSynchronize());
    for(int ctas: {1,2,3,4,6,8}){ run<2>(table,idx,m,sink,ctas); run<4>(table,idx,m,sink,ctas); run<8>(table,idx,m,sink,ctas); run<16>(table,idx,m,sink,ctas); }
    return 0;
}
