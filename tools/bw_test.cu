// read-bandwidth ceiling probes for the scan kernel's access pattern (scratch, not product)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../gofeature_b200/csrc/gf_device.cuh"
using namespace gf;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

// A/B/C/E: bulk-copy ring; TILE bytes per stage split in NSPLIT copies; consumers read tile with LDS.128 and sum
template<int TILE, int NSPLIT, bool HINT>
__global__ void __launch_bounds__(256) k_bulk(const float* __restrict__ src, size_t ntiles, int stages, float* out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    float* tiles=(float*)smem; uint64_t* full=(uint64_t*)(smem+(size_t)stages*TILE); uint64_t* empty=full+stages;
    int warp=threadIdx.x>>5, lane=threadIdx.x&31;
    if(threadIdx.x==0){ for(int s=0;s<stages;++s){ mbar_init(&full[s],1); mbar_init(&empty[s],7);} mbar_fence_init(); }
    __syncthreads();
    if(warp==7){
        if(lane==0){
            uint64_t pol=l2_policy_evict_first(); int s=0; uint32_t ph=0;
            for(size_t t=blockIdx.x;t<ntiles;t+=gridDim.x){
                mbar_wait(&empty[s],ph^1);
                mbar_arrive_expect_tx(&full[s],TILE);
                #pragma unroll
                for(int k=0;k<NSPLIT;++k){
                    const char* g=(const char*)src+t*TILE+(size_t)k*(TILE/NSPLIT);
                    char* d=(char*)tiles+(size_t)s*TILE+(size_t)k*(TILE/NSPLIT);
                    if(HINT) bulk_g2s_hint(d,g,TILE/NSPLIT,&full[s],pol); else bulk_g2s(d,g,TILE/NSPLIT,&full[s]);
                }
                if(++s==stages){s=0;ph^=1;}
            }
        }
        return;
    }
    float acc=0.f; int s=0; uint32_t ph=0;
    for(size_t t=blockIdx.x;t<ntiles;t+=gridDim.x){
        mbar_wait(&full[s],ph);
        const float4* tl=(const float4*)((char*)tiles+(size_t)s*TILE);
        #pragma unroll 4
        for(int i=warp*32+lane;i<TILE/16;i+=224){ float4 v=tl[i]; acc+=v.x+v.y+v.z+v.w; }
        __syncwarp(); if(lane==0) mbar_arrive(&empty[s]);
        if(++s==stages){s=0;ph^=1;}
    }
    if(acc==123.456f) out[0]=acc;
}
// D: direct LDG streaming
template<int UNROLL>
__global__ void __launch_bounds__(512) k_ldg(const float4* __restrict__ src, size_t n4, float* out)
{
    size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x, stride=(size_t)gridDim.x*blockDim.x; float acc=0.f;
    for(; i+ (UNROLL-1)*stride<n4; i+=UNROLL*stride){
        float4 v[UNROLL];
        #pragma unroll
        for(int u=0;u<UNROLL;++u) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(v[u].x),"=f"(v[u].y),"=f"(v[u].z),"=f"(v[u].w):"l"(src+i+u*stride));
        #pragma unroll
        for(int u=0;u<UNROLL;++u) acc+=v[u].x+v[u].y+v[u].z+v[u].w;
    }
    if(acc==123.456f) out[0]=acc;
}
template<typename F> float timeit(F f,int reps=20){ cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b); for(int i=0;i<3;++i)f(); cudaEventRecord(a); for(int i=0;i<reps;++i)f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); return ms/reps; }
template<int TILE,int NSPLIT,bool HINT> void run_bulk(const float* d,size_t bytes,int stages,int grid,float* out,const char* name){
    size_t smem=(size_t)stages*TILE+256; CK(cudaFuncSetAttribute(k_bulk<TILE,NSPLIT,HINT>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
    size_t nt=bytes/TILE; float ms=timeit([&]{k_bulk<TILE,NSPLIT,HINT><<<grid,256,smem>>>(d,nt,stages,out);}); CK(cudaGetLastError());
    printf("%-44s tile=%6d split=%d stages=%2d grid=%4d : %.3f ms  %.0f GB/s\n",name,TILE,NSPLIT,stages,grid,ms,nt*(double)TILE/ms/1e6);
}
int main(){
    size_t bytes=(size_t)2048000000; float* d; CK(cudaMalloc(&d,bytes+1024)); CK(cudaMemset(d,0,bytes)); float* out; CK(cudaMalloc(&out,4));
    run_bulk<28672,1,true >(d,bytes,7,148,out,"A bulk 28K x7 hint");
    run_bulk<28672,1,false>(d,bytes,7,148,out,"E bulk 28K x7 nohint");
    run_bulk<28672,4,false>(d,bytes,7,148,out,"B bulk 28K split4 x7");
    run_bulk<14336,1,false>(d,bytes,14,148,out,"bulk 14K x14");
    run_bulk<57344,1,false>(d,bytes,3,148,out,"bulk 56K x3");
    run_bulk<28672,1,false>(d,bytes,3,296,out,"C bulk 28K x3, 2 CTA/SM");
    run_bulk<14336,1,false>(d,bytes,7,296,out,"bulk 14K x7, 2 CTA/SM");
    run_bulk<28672,1,false>(d,bytes,4,148,out,"bulk 28K x4");
    run_bulk<28672,1,false>(d,bytes,2,148,out,"bulk 28K x2");
    run_bulk<8192,1,false>(d,bytes,24,148,out,"bulk 8K x24");
    run_bulk<8192,1,false>(d,bytes,12,296,out,"bulk 8K x12 2CTA/SM");
    run_bulk<4096,1,false>(d,bytes,48,148,out,"bulk 4K x48");
    { size_t n4=bytes/16; float ms=timeit([&]{k_ldg<8><<<148*4,512>>>((const float4*)d,n4,out);}); printf("D ldg.nc v4 unroll8 grid 592x512: %.3f ms %.0f GB/s\n",ms,bytes/ms/1e6);
      ms=timeit([&]{k_ldg<4><<<148*4,512>>>((const float4*)d,n4,out);}); printf("D ldg.nc v4 unroll4 grid 592x512: %.3f ms %.0f GB/s\n",ms,bytes/ms/1e6);
      ms=timeit([&]{k_ldg<8><<<148*2,512>>>((const float4*)d,n4,out);}); printf("D ldg.nc v4 unroll8 grid 296x512: %.3f ms %.0f GB/s\n",ms,bytes/ms/1e6);
      ms=timeit([&]{k_ldg<16><<<148*2,512>>>((const float4*)d,n4,out);}); printf("D ldg.nc v4 unroll16 grid 296x512: %.3f ms %.0f GB/s\n",ms,bytes/ms/1e6); }
    { float* d2; CK(cudaMalloc(&d2,bytes)); float ms=timeit([&]{cudaMemcpyAsync(d2,d,bytes/2,cudaMemcpyDeviceToDevice);}); printf("cudaMemcpy D2D 1GB: %.3f ms  %.0f GB/s (r+w)\n",ms,bytes/ms/1e6); }
    return 0;
}
