// micro-benchmark: random ROW gathers from a DRAM-resident table on B200, by row size.
// The ceiling of the gather sweeps (csrc/pull.cuh): per out-list entry one row of x (8 B x sources) is read at a random
// position.  Rows of R bytes are read by R/32 adjacent lanes (32 B per lane: 2 x LDG.128), `UNROLL` independent rows in
// flight per lane group; the row indices stream in coalesced (4 B each, like the out-list slots).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/gather scripts/micro/gather.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__device__ __forceinline__ unsigned hash32(unsigned x){ x^=x>>16; x*=0x7feb352dU; x^=x>>15; x*=0x846ca68bU; x^=x>>16; return x; }

__global__ void fill_idx(unsigned* idx, size_t m, unsigned nrows, unsigned hot_rows, unsigned hot1024){
    for(size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x;i<m;i+=(size_t)gridDim.x*blockDim.x){
        unsigned h=hash32((unsigned)i*2654435761u+17u), h2=hash32(h+0x9e3779b9u);
        idx[i] = ((h&1023u)<hot1024) ? h2%hot_rows : h2%nrows;   // a fraction of the gathers goes to a small hot set
    }
}

// G = lanes per row (row bytes = 32 G); each lane group walks its own contiguous run of indices
template<int G, int UNROLL>
__global__ void __launch_bounds__(256) gather(const double4* __restrict__ table, const unsigned* __restrict__ idx, size_t m, double* sink){
    const unsigned lane_in_g = threadIdx.x & (G-1);
    const size_t group = ((size_t)blockIdx.x*blockDim.x+threadIdx.x)/G, ngroups=((size_t)gridDim.x*blockDim.x)/G;
    const size_t per=(m+ngroups-1)/ngroups, lo=group*per, hi=min(m,lo+per);
    double acc=0;
    for(size_t i=lo;i<hi;i+=UNROLL){
        unsigned u[UNROLL];
#pragma unroll
        for(int k=0;k<UNROLL;++k) u[k]= (i+k<hi)? __ldcs(&idx[i+k]) : 0xffffffffu;
        double4 v[UNROLL];
#pragma unroll
        for(int k=0;k<UNROLL;++k) v[k]= (u[k]!=0xffffffffu)? table[(size_t)u[k]*G+lane_in_g] : make_double4(0,0,0,0);
#pragma unroll
        for(int k=0;k<UNROLL;++k) acc+=v[k].x+v[k].y+v[k].z+v[k].w;
    }
    if(acc==12345.678) *sink=acc;
}

template<int G,int UNROLL>
void run(const double4* table, const unsigned* idx, size_t m, double* sink, size_t table_bytes, const char* tag, int ctas_per_sm){
    cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks=148*ctas_per_sm;
    gather<G,UNROLL><<<blocks,256>>>(table,idx,m,sink);
    cudaEventRecord(a);
    gather<G,UNROLL><<<blocks,256>>>(table,idx,m,sink);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms,a,b);
    printf("%-22s row %5d B unroll %d ctas/SM %d: %8.3f ms  %7.1f rows/ns  %7.1f GB/s gathered\n", tag, 32*G, UNROLL, ctas_per_sm, ms, m/ms*1e-6, (double)m*32*G/ms*1e-6);
}

int main(int argc,char**argv){
    const size_t table_bytes = 3ull<<30;   // 3 GiB: the x array of BASELINE configs[3] with 125 sources
    double4* table; CK(cudaMalloc(&table, table_bytes)); CK(cudaMemset(table,0,table_bytes));
    double* sink; CK(cudaMalloc(&sink,8));
    struct Case{ const char* tag; unsigned hot_rows_div, hot1024; } cases[]={{"uniform",1,0},{"50% to hottest 1%",100,512}};
    for(auto c: cases){
        for(int g=1; g<=32; g*=2){
            const unsigned nrows=(unsigned)(table_bytes/(32ull*g));
            const size_t m = (size_t)(24ull<<30)/(32ull*g) > (1ull<<28) ? (1ull<<28) : (size_t)(24ull<<30)/(32ull*g);  // <= 24 GiB gathered
            unsigned* idx; CK(cudaMalloc(&idx,m*4));
            fill_idx<<<148*8,256>>>(idx,m,nrows,nrows/c.hot_rows_div>0?nrows/c.hot_rows_div:1,c.hot1024); CK(cudaDeviceSynchronize());
            switch(g){
                case 1: run<1,4>(table,idx,m,sink,table_bytes,c.tag,8); run<1,8>(table,idx,m,sink,table_bytes,c.tag,8); break;
                case 2: run<2,4>(table,idx,m,sink,table_bytes,c.tag,8); run<2,8>(table,idx,m,sink,table_bytes,c.tag,8); break;
                case 4: run<4,4>(table,idx,m,sink,table_bytes,c.tag,8); run<4,8>(table,idx,m,sink,table_bytes,c.tag,8); break;
                case 8: run<8,4>(table,idx,m,sink,table_bytes,c.tag,4); run<8,4>(table,idx,m,sink,table_bytes,c.tag,8); run<8,8>(table,idx,m,sink,table_bytes,c.tag,8); break;
                case 16: run<16,4>(table,idx,m,sink,table_bytes,c.tag,8); run<16,8>(table,idx,m,sink,table_bytes,c.tag,8); break;
                case 32: run<32,2>(table,idx,m,sink,table_bytes,c.tag,8); run<32,4>(table,idx,m,sink,table_bytes,c.tag,8); run<32,8>(table,idx,m,sink,table_bytes,c.tag,8); break;
            }
            CK(cudaFree(idx));
        }
    }
    return 0;
}
