// periodic.cuh -- periodic duplicates ("ghosts") for sm_100a.
//
// Replaces cusph::PeriodicIgnore / PeriodicMakeList / PeriodicDuplicateVerlet|Symplectic (S/JSphGpu_ker.cu:2445-2712).
// The reference GPU builds the list with atomics (order changes run to run, S/JSphGpu_ker.cu:2493-2504); here the list
// is built in two deterministic passes (count, exclusive scan, write), which reproduces the order of the reference CPU
// (oracle) list: ascending particle, +perinc before -perinc (JSphCpuSingle::PeriodicMakeList S/JSphCpuSingle.cpp:182-213).
#pragma once
#include "devtypes.h"
#include "celldiv.cuh"

namespace dsph {

struct PeriArgs {
  unsigned n, pini;
  double inc[3], mapmin[3], mapmax[3];
  const typecode *code; const double *pos;
};
struct PeriDupArrays {
  unsigned *idp; typecode *code; unsigned *dcell; double *pos; float4 *velrhop;
  float4 *velrhopm1; double *pospre; float4 *velrhoppre; float *spstau;
};

__global__ void ker_peri_ignore(unsigned n,typecode* __restrict__ code){
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<n){
    const unsigned rc=code[p];
    if((rc&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC)code[p]=typecode((rc&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_OUTIGNORE);
  }
}
__device__ __forceinline__ bool PeriInMap(double x,double y,double z,const PeriArgs &a){
  return (a.mapmin[0]<=x && a.mapmin[1]<=y && a.mapmin[2]<=z && x<a.mapmax[0] && y<a.mapmax[1] && z<a.mapmax[2]);
}
__global__ void ker_peri_count(PeriArgs a,unsigned* __restrict__ count){
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=a.n)return;
  const unsigned p=a.pini+i; unsigned c=0;
  if((a.code[p]&DSPH_CODE_MASKSPECIAL)<=DSPH_CODE_PERIODIC){
    const double x=a.pos[3*size_t(p)],y=a.pos[3*size_t(p)+1],z=a.pos[3*size_t(p)+2];
    if(PeriInMap(x+a.inc[0],y+a.inc[1],z+a.inc[2],a))c++;
    if(PeriInMap(x-a.inc[0],y-a.inc[1],z-a.inc[2],a))c++;
  }
  count[i]=c;
}
// JSphCpuSingle::PeriodicDuplicatePos + PeriodicDuplicateVerlet/Symplectic S/JSphCpuSingle.cpp:232-321
__device__ __forceinline__ void PeriDuplicateOne(unsigned pnew,unsigned pcopy,double sx,double sy,double sz,const PeriDupArrays &d,const DevCtes &c){
  const double x=d.pos[3*size_t(pcopy)]+sx, y=d.pos[3*size_t(pcopy)+1]+sy, z=d.pos[3*size_t(pcopy)+2]+sz;
  unsigned cx=unsigned((x-c.domposmin[0])/c.scell_d), cy=unsigned((y-c.domposmin[1])/c.scell_d), cz=unsigned((z-c.domposmin[2])/c.scell_d);
  cx=(cx<=c.domcells[0]? cx: c.domcells[0]); cy=(cy<=c.domcells[1]? cy: c.domcells[1]); cz=(cz<=c.domcells[2]? cz: c.domcells[2]);
  d.pos[3*size_t(pnew)]=x; d.pos[3*size_t(pnew)+1]=y; d.pos[3*size_t(pnew)+2]=z;
  d.dcell[pnew]=DcCell(c.domcellcode,cx,cy,cz);
  d.idp[pnew]=d.idp[pcopy];
  d.code[pnew]=typecode((d.code[pcopy]&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_PERIODIC);
  d.velrhop[pnew]=d.velrhop[pcopy];
  if(d.velrhopm1)d.velrhopm1[pnew]=d.velrhopm1[pcopy];
  if(d.pospre){
    d.pospre[3*size_t(pnew)]=d.pospre[3*size_t(pcopy)]; d.pospre[3*size_t(pnew)+1]=d.pospre[3*size_t(pcopy)+1]; d.pospre[3*size_t(pnew)+2]=d.pospre[3*size_t(pcopy)+2];
    d.velrhoppre[pnew]=d.velrhoppre[pcopy];
  }
  if(d.spstau){
    #pragma unroll
    for(int k=0;k<6;k++)d.spstau[6*size_t(pnew)+k]=d.spstau[6*size_t(pcopy)+k];
  }
}
__global__ void ker_peri_duplicate(PeriArgs a,const unsigned* __restrict__ offset,unsigned npbase,PeriDupArrays d,DevCtes c){
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=a.n)return;
  const unsigned p=a.pini+i;
  if((d.code[p]&DSPH_CODE_MASKSPECIAL)<=DSPH_CODE_PERIODIC){
    const double x=a.pos[3*size_t(p)],y=a.pos[3*size_t(p)+1],z=a.pos[3*size_t(p)+2];
    unsigned o=npbase+offset[i];
    if(PeriInMap(x+a.inc[0],y+a.inc[1],z+a.inc[2],a)){ PeriDuplicateOne(o,p, a.inc[0], a.inc[1], a.inc[2],d,c); o++; }
    if(PeriInMap(x-a.inc[0],y-a.inc[1],z-a.inc[2],a)){ PeriDuplicateOne(o,p,-a.inc[0],-a.inc[1],-a.inc[2],d,c); }
  }
}

} // namespace dsph
