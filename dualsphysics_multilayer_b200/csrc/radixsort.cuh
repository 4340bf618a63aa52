// radixsort.cuh -- stable LSD radix sort of (key u32, value u32) pairs restricted to the significant key bits,
// plus the block-scan primitives it needs.  Hand-written for sm_100a (no Thrust/CUB).
//
// Replaces cudiv::Sort (S/JCellDivGpu_ker.cu:116-123 = thrust::sort_by_key / stable_sort_by_key).  It is STABLE, so
// the resulting permutation equals the reference CPU counting sort (JCellDivCpuSingle::MakeSortFull,
// S/JCellDivCpuSingle.cpp:203-214): particles of one box keep their previous relative order.
//
// Per 8-bit pass: (1) per-tile digit histogram -> hist[digit][tile]; (2) exclusive scan of the flattened table;
// (3) scatter with stable in-tile ranks (warp match.any + per-warp digit counters in shared memory).
// HBM traffic per pass: 4 B (hist read) + 8 B (scatter read) + 8 B (scatter write) per element.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "devtypes.h"

namespace dsph {

constexpr int RS_THREADS=256;
constexpr int RS_ITEMS=16;
constexpr int RS_TILE=RS_THREADS*RS_ITEMS;   // 4096 keys per block
constexpr int RS_WARPS=RS_THREADS/32;

__global__ void __launch_bounds__(RS_THREADS) ker_rs_hist(const uint32_t* __restrict__ keys,uint32_t n,int shift,uint32_t ntiles,uint32_t* __restrict__ hist)
{
  __shared__ uint32_t sh[256];
  sh[threadIdx.x]=0;
  __syncthreads();
  const uint32_t base=blockIdx.x*RS_TILE;
  #pragma unroll
  for(int i=0;i<RS_ITEMS;i++){
    const uint32_t p=base+i*RS_THREADS+threadIdx.x;
    if(p<n)atomicAdd(&sh[(keys[p]>>shift)&255u],1u);
  }
  __syncthreads();
  hist[threadIdx.x*ntiles+blockIdx.x]=sh[threadIdx.x];
}

// ---- generic exclusive +scan over u32, three kernels (chunk scan, scan of chunk sums, add) ------------------
constexpr int SC_THREADS=256;
constexpr int SC_ITEMS=8;
constexpr int SC_CHUNK=SC_THREADS*SC_ITEMS;   // 2048

__device__ __forceinline__ uint32_t BlockExclusiveSum(uint32_t v,uint32_t *total){
  __shared__ uint32_t wsum[SC_THREADS/32];
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  uint32_t inc=v;
  #pragma unroll
  for(int o=1;o<32;o<<=1){ const uint32_t t=__shfl_up_sync(0xffffffffu,inc,o); if(lane>=o)inc+=t; }
  if(lane==31)wsum[w]=inc;
  __syncthreads();
  if(w==0){
    uint32_t s=(lane<SC_THREADS/32? wsum[lane]: 0);
    #pragma unroll
    for(int o=1;o<SC_THREADS/32;o<<=1){ const uint32_t t=__shfl_up_sync(0xffffffffu,s,o); if(lane>=o)s+=t; }
    if(lane<SC_THREADS/32)wsum[lane]=s;   // inclusive over warps
  }
  __syncthreads();
  const uint32_t woff=(w? wsum[w-1]: 0);
  if(total)*total=wsum[SC_THREADS/32-1];
  const uint32_t r=woff+inc-v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SC_THREADS) ker_scan_chunks(uint32_t* __restrict__ data,uint32_t n,uint32_t* __restrict__ chunksum)
{
  const uint32_t base=blockIdx.x*SC_CHUNK+threadIdx.x*SC_ITEMS;
  uint32_t v[SC_ITEMS]; uint32_t s=0;
  #pragma unroll
  for(int i=0;i<SC_ITEMS;i++){ v[i]=(base+i<n? data[base+i]: 0); s+=v[i]; }
  uint32_t total;
  uint32_t off=BlockExclusiveSum(s,&total);
  #pragma unroll
  for(int i=0;i<SC_ITEMS;i++){ if(base+i<n)data[base+i]=off; off+=v[i]; }
  if(threadIdx.x==0)chunksum[blockIdx.x]=total;
}
// single block: exclusive scan of m chunk sums in place (loops when m > SC_CHUNK)
__global__ void __launch_bounds__(SC_THREADS) ker_scan_sums(uint32_t* __restrict__ sums,uint32_t m)
{
  uint32_t carry=0;
  for(uint32_t b0=0;b0<m;b0+=SC_CHUNK){
    const uint32_t base=b0+threadIdx.x*SC_ITEMS;
    uint32_t v[SC_ITEMS]; uint32_t s=0;
    #pragma unroll
    for(int i=0;i<SC_ITEMS;i++){ v[i]=(base+i<m? sums[base+i]: 0); s+=v[i]; }
    uint32_t total;
    uint32_t off=BlockExclusiveSum(s,&total)+carry;
    #pragma unroll
    for(int i=0;i<SC_ITEMS;i++){ if(base+i<m)sums[base+i]=off; off+=v[i]; }
    carry+=total;
  }
}
__global__ void __launch_bounds__(SC_THREADS) ker_scan_add(uint32_t* __restrict__ data,uint32_t n,const uint32_t* __restrict__ chunkoff)
{
  const uint32_t base=blockIdx.x*SC_CHUNK+threadIdx.x*SC_ITEMS;
  const uint32_t off=chunkoff[blockIdx.x];
  #pragma unroll
  for(int i=0;i<SC_ITEMS;i++)if(base+i<n)data[base+i]+=off;
}

// ---- scatter with stable ranks ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS) ker_rs_scatter(const uint32_t* __restrict__ keys,const uint32_t* __restrict__ vals,
  uint32_t n,int shift,uint32_t ntiles,const uint32_t* __restrict__ histscan,uint32_t* __restrict__ keysout,uint32_t* __restrict__ valsout,int iotavals)
{
  __shared__ uint32_t wcount[RS_WARPS][256];   // per-warp digit counters -> exclusive offsets across warps
  __shared__ uint32_t gbase[256];
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  for(int i=threadIdx.x;i<RS_WARPS*256;i+=RS_THREADS)(&wcount[0][0])[i]=0;
  gbase[threadIdx.x]=histscan[threadIdx.x*ntiles+blockIdx.x];
  __syncthreads();
  // warp w owns the contiguous sub-tile [wbase, wbase+32*RS_ITEMS); element (i,lane) = wbase+i*32+lane, i.e. original
  // order inside the warp is (i, lane): ranks are assigned in that order -> stable.
  const uint32_t wbase=blockIdx.x*RS_TILE+w*(32*RS_ITEMS);
  uint32_t key[RS_ITEMS]; uint32_t rank[RS_ITEMS];
  const uint32_t ltmask=(1u<<lane)-1u;
  #pragma unroll
  for(int i=0;i<RS_ITEMS;i++){
    const uint32_t p=wbase+i*32+lane;
    const bool ok=(p<n);
    key[i]=(ok? keys[p]: 0xffffffffu);
    const uint32_t d=(ok? ((key[i]>>shift)&255u): 256u);     // 256: invalid lanes form their own group
    const uint32_t peers=__match_any_sync(0xffffffffu,d);
    const int leader=__ffs(peers)-1;
    uint32_t old=0;
    if(ok && lane==leader){ old=wcount[w][d]; wcount[w][d]=old+__popc(peers); }
    old=__shfl_sync(0xffffffffu,old,leader);
    rank[i]=old+__popc(peers&ltmask);
    __syncwarp();
  }
  __syncthreads();
  { // exclusive prefix over warps for digit = threadIdx.x, plus the global base of this tile
    const int d=threadIdx.x; uint32_t s=gbase[d];
    #pragma unroll
    for(int ww=0;ww<RS_WARPS;ww++){ const uint32_t c=wcount[ww][d]; wcount[ww][d]=s; s+=c; }
  }
  __syncthreads();
  #pragma unroll
  for(int i=0;i<RS_ITEMS;i++){
    const uint32_t p=wbase+i*32+lane;
    if(p<n){
      const uint32_t d=(key[i]>>shift)&255u;
      const uint32_t dst=wcount[w][d]+rank[i];
      keysout[dst]=key[i];
      valsout[dst]=(iotavals? p: vals[p]);
    }
  }
}

inline uint32_t RsTiles(uint32_t n){ return (n+RS_TILE-1)/RS_TILE; }
inline uint32_t ScChunks(uint32_t m){ return (m+SC_CHUNK-1)/SC_CHUNK; }

// exclusive scan of data[0..m) in place; tmp1 holds tmp1cap >= ScChunks(m) entries, tmp2 >= ScChunks(ScChunks(m)) entries.
// Returns -1 (nothing launched) when the scratch is too small for m.
inline int DeviceExclusiveScan(uint32_t *data,uint32_t m,uint32_t *tmp1,uint32_t tmp1cap,uint32_t *tmp2,uint32_t tmp2cap,cudaStream_t stm,unsigned long long *launches){
  if(!m)return 0;
  const uint32_t c1=ScChunks(m);
  if(c1>tmp1cap || ScChunks(c1)>tmp2cap)return -1;
  ker_scan_chunks<<<c1,SC_THREADS,0,stm>>>(data,m,tmp1); (*launches)++;
  if(c1>1){
    if(c1<=SC_CHUNK*64){ ker_scan_sums<<<1,SC_THREADS,0,stm>>>(tmp1,c1); (*launches)++; }
    else{
      const uint32_t c2=ScChunks(c1);
      ker_scan_chunks<<<c2,SC_THREADS,0,stm>>>(tmp1,c1,tmp2); (*launches)++;
      ker_scan_sums<<<1,SC_THREADS,0,stm>>>(tmp2,c2); (*launches)++;
      ker_scan_add<<<c2,SC_THREADS,0,stm>>>(tmp1,c1,tmp2); (*launches)++;
    }
    ker_scan_add<<<c1,SC_THREADS,0,stm>>>(data,m,tmp1); (*launches)++;
  }
  return 0;
}

// Sorts keys[0]/vals[0] (n pairs, key < 2^keybits).  Values of the first pass are generated as iota when
// iotavals!=0.  Returns the index (0|1) of the buffers that hold the result.
inline int RadixSortPairs(RadixSortBuffers &b,uint32_t n,int keybits,int iotavals,cudaStream_t stm,unsigned long long *launches){
  if(!n)return 0;
  const uint32_t ntiles=RsTiles(n);
  int cur=0;
  const int npass=(keybits+7)/8;
  for(int pass=0;pass<npass;pass++){
    const int shift=pass*8;
    ker_rs_hist<<<ntiles,RS_THREADS,0,stm>>>(b.keys[cur],n,shift,ntiles,b.hist); (*launches)++;
    DeviceExclusiveScan(b.hist,256u*ntiles,b.chunksum,b.chunksumcap,b.chunksum2,b.chunksum2cap,stm,launches);
    ker_rs_scatter<<<ntiles,RS_THREADS,0,stm>>>(b.keys[cur],b.vals[cur],n,shift,ntiles,b.hist,b.keys[cur^1],b.vals[cur^1],(pass==0 && iotavals)?1:0); (*launches)++;
    cur^=1;
  }
  if(npass==0 && iotavals){ /* nothing to sort: caller fills iota */ }
  return cur;
}

} // namespace dsph
