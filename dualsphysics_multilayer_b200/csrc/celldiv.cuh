// celldiv.cuh -- GPU cell-linked list, hand-written for sm_100a.
//
//   ker_div_limits    min/max cell of valid particles           (cudiv::LimitsCell, S/JCellDivGpu_ker.cu:343-481;
//                                                                 CPU: JCellDivCpu::LimitsCellBound/Fluid S/JCellDivCpu.cpp:255-330)
//   ker_div_prepare   cell domain + special boxes ON THE DEVICE  (JCellDivCpuSingle::CalcCellDomain/MergeMapCellBoundFluid/PrepareNct,
//                                                                 S/JCellDivCpuSingle.cpp:47-125)
//   ker_div_presort   sort key per particle                      (KerPreSortFull S/JCellDivGpuSingle_ker.cu:41-72,
//                                                                 CPU PreSortFull S/JCellDivCpuSingle.cpp:134-162)
//   RadixSortPairs    stable sort (radixsort.cuh)                (cudiv::Sort)
//   ker_div_cellmark + suffix-min scan   begin-cell table in the CPU form unsigned[2*nct+6] (S/JCellDivCpu.h:80,141,153)
//   ker_div_counts    sizes of the special boxes                 (S/JCellDivCpuSingle.cpp:326-337)
//   ker_div_reorder   ONE fused gather of every particle array + cell-relative position + pressure
//                     (replaces 6+ KerSortDataParticles launches S/JCellDivGpu_ker.cu:553-642 + cusphs::UpdatePosCell
//                      S/JSphGpuSimple_ker.cu:41-69 + the per-pair pow() for pressure S/JSphGpu_ker.cu:559)
#pragma once
#include "devtypes.h"
#include "radixsort.cuh"
#include <cuda_fp16.h>
#include "../../include/dsph_b200.h"

namespace dsph {

constexpr int DIV_THREADS=256;


// First cell hash of uploaded particles: JSph::LoadDcellParticles (S/JSph.cpp:1690-1711), periodic duplicates of a
// restarted state follow PeriodicDuplicatePos (S/JSphCpuSingle.cpp:232-251).  nout counts valid particles out of the domain.
__global__ void __launch_bounds__(DIV_THREADS) ker_upload_dcell(unsigned n,const double* __restrict__ pos,const typecode* __restrict__ code,
  unsigned* __restrict__ dcell,unsigned* __restrict__ nout,DevCtes c,double rminx,double rminy,double rminz,double rmaxx,double rmaxy,double rmaxz)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  const unsigned sp=code[p]&DSPH_CODE_MASKSPECIAL;
  unsigned cell=0xFFFFFFFFu;
  if(sp<DSPH_CODE_OUTIGNORE){
    const double x=pos[3*size_t(p)],y=pos[3*size_t(p)+1],z=pos[3*size_t(p)+2];
    unsigned cx=unsigned((x-c.domposmin[0])/c.scell_d), cy=unsigned((y-c.domposmin[1])/c.scell_d), cz=unsigned((z-c.domposmin[2])/c.scell_d);
    if(sp==DSPH_CODE_PERIODIC){
      cx=(cx<=c.domcells[0]? cx: c.domcells[0]); cy=(cy<=c.domcells[1]? cy: c.domcells[1]); cz=(cz<=c.domcells[2]? cz: c.domcells[2]);
      cell=DcCell(c.domcellcode,cx,cy,cz);
    }
    else if(x>=rminx && y>=rminy && z>=rminz && x<rmaxx && y<rmaxy && z<rmaxz)cell=DcCell(c.domcellcode,cx,cy,cz);
    else atomicAdd(nout,1u);
  }
  dcell[p]=cell;
}

// limits[0..2]=min (init UINT_MAX), limits[3..5]=max (init 0)
__global__ void __launch_bounds__(DIV_THREADS) ker_div_limits(unsigned n,unsigned pini,const unsigned* __restrict__ dcell,
  const typecode* __restrict__ code,unsigned domcellcode,unsigned* __restrict__ limits)
{
  unsigned mn[3]={0xffffffffu,0xffffffffu,0xffffffffu}, mx[3]={0,0,0};
  for(unsigned i=blockIdx.x*blockDim.x+threadIdx.x;i<n;i+=gridDim.x*blockDim.x){
    const unsigned p=pini+i;
    if((code[p]&DSPH_CODE_MASKSPECIAL)<DSPH_CODE_OUTIGNORE){
      const unsigned rcell=dcell[p];
      const unsigned c[3]={DcCellx(domcellcode,rcell),DcCelly(domcellcode,rcell),DcCellz(domcellcode,rcell)};
      #pragma unroll
      for(int k=0;k<3;k++){ mn[k]=min(mn[k],c[k]); mx[k]=max(mx[k],c[k]); }
    }
  }
  #pragma unroll
  for(int k=0;k<3;k++){
    mn[k]=__reduce_min_sync(0xffffffffu,mn[k]);
    mx[k]=__reduce_max_sync(0xffffffffu,mx[k]);
  }
  if((threadIdx.x&31)==0){
    #pragma unroll
    for(int k=0;k<3;k++){ atomicMin(&limits[k],mn[k]); atomicMax(&limits[3+k],mx[k]); }
  }
}

// the same over one range that mixes boundary and fluid particles (slab runs with periodic copies appended in any order): every
// particle goes to the boundary (limb) or the fluid (limf) limits by the type bits of its code
__global__ void __launch_bounds__(DIV_THREADS) ker_div_limits_bytype(unsigned n,const unsigned* __restrict__ dcell,
  const typecode* __restrict__ code,unsigned domcellcode,unsigned* __restrict__ limb,unsigned* __restrict__ limf)
{
  unsigned mn[2][3]={{0xffffffffu,0xffffffffu,0xffffffffu},{0xffffffffu,0xffffffffu,0xffffffffu}}, mx[2][3]={{0,0,0},{0,0,0}};
  for(unsigned p=blockIdx.x*blockDim.x+threadIdx.x;p<n;p+=gridDim.x*blockDim.x){
    const unsigned rc=code[p];
    if((rc&DSPH_CODE_MASKSPECIAL)<DSPH_CODE_OUTIGNORE){
      const int t=((rc&DSPH_CODE_MASKTYPE)<DSPH_CODE_TYPE_FLOATING? 0: 1);
      const unsigned rcell=dcell[p];
      const unsigned c[3]={DcCellx(domcellcode,rcell),DcCelly(domcellcode,rcell),DcCellz(domcellcode,rcell)};
      #pragma unroll
      for(int k=0;k<3;k++){
        if(t==0){ mn[0][k]=min(mn[0][k],c[k]); mx[0][k]=max(mx[0][k],c[k]); }
        else{     mn[1][k]=min(mn[1][k],c[k]); mx[1][k]=max(mx[1][k],c[k]); }
      }
    }
  }
  #pragma unroll
  for(int t=0;t<2;t++){
    #pragma unroll
    for(int k=0;k<3;k++){
      const unsigned a=__reduce_min_sync(0xffffffffu,mn[t][k]), b=__reduce_max_sync(0xffffffffu,mx[t][k]);
      if((threadIdx.x&31)==0){ unsigned *lim=(t? limf: limb); atomicMin(&lim[k],a); atomicMax(&lim[3+k],b); }
    }
  }
}

// limb/limf: raw limits of boundary / fluid particles.  useblim: recompute the cached boundary limits.
__global__ void ker_div_prepare(const unsigned* __restrict__ limb,const unsigned* __restrict__ limf,int useblim,
  unsigned dcx,unsigned dcy,unsigned dcz,unsigned scelldiv,unsigned nptot,unsigned maxnctt,DivParams* __restrict__ dp)
{
  if(threadIdx.x||blockIdx.x)return;
  const unsigned domcells[3]={dcx,dcy,dcz};
  unsigned celbmin[3],celbmax[3],celfmin[3],celfmax[3];
  if(useblim){
    const bool empty=(limb[0]>limb[3]);
    for(int i=0;i<3;i++){ celbmin[i]=(empty? domcells[i]: limb[i]); celbmax[i]=(empty? 0: limb[3+i]); dp->blimmin[i]=celbmin[i]; dp->blimmax[i]=celbmax[i]; }
  }
  else for(int i=0;i<3;i++){ celbmin[i]=dp->blimmin[i]; celbmax[i]=dp->blimmax[i]; }
  {
    const bool empty=(limf[0]>limf[3]);
    for(int i=0;i<3;i++){ celfmin[i]=(empty? domcells[i]: limf[i]); celfmax[i]=(empty? 0: limf[3+i]); }
  }
  unsigned cmin[3],cmax[3];
  for(int i=0;i<3;i++){
    cmin[i]=max(min(celbmin[i],celfmin[i]),(celfmin[i]>=scelldiv? celfmin[i]-scelldiv: 0u));
    cmax[i]=min(max(celbmax[i],celfmax[i]),celfmax[i]+scelldiv);
    if(cmax[i]>=domcells[i])cmax[i]=domcells[i]-1;
  }
  if(cmin[0]>cmax[0]||cmin[1]>cmax[1]||cmin[2]>cmax[2])for(int i=0;i<3;i++)cmin[i]=cmax[i]=0;
  for(int i=0;i<3;i++){ dp->cellmin[i]=cmin[i]; dp->cellmax[i]=cmax[i]; }
  const unsigned ncx=cmax[0]-cmin[0]+1, ncy=cmax[1]-cmin[1]+1, ncz=cmax[2]-cmin[2]+1;
  const unsigned long long nct64=(unsigned long long)(ncx)*ncy*ncz;
  dp->error=(nct64*2+6>maxnctt? 1u: 0u);
  const unsigned nct=unsigned(nct64);
  dp->ncx=ncx; dp->ncy=ncy; dp->ncz=ncz; dp->nsheet=ncx*ncy; dp->nct=nct; dp->nctt=nct*2+6;
  dp->boxboundignore=nct; dp->boxfluid=nct+1; dp->boxboundout=nct+1+nct; dp->boxfluidout=dp->boxboundout+1;
  dp->boxboundoutignore=dp->boxfluidout+1; dp->boxfluidoutignore=dp->boxboundoutignore+1;
  dp->nptot=nptot; dp->maxcellpop=0;
}

__global__ void __launch_bounds__(DIV_THREADS) ker_div_presort(unsigned n,const unsigned* __restrict__ dcell,const typecode* __restrict__ code,
  unsigned domcellcode,const DivParams* __restrict__ dpp,unsigned* __restrict__ cellpart)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  const DivParams &dp=*dpp;
  const unsigned rcell=dcell[p];
  const unsigned cx=DcCellx(domcellcode,rcell)-dp.cellmin[0];
  const unsigned cy=DcCelly(domcellcode,rcell)-dp.cellmin[1];
  const unsigned cz=DcCellz(domcellcode,rcell)-dp.cellmin[2];
  const unsigned cellsort=cx+cy*dp.ncx+cz*dp.nsheet;
  const unsigned rcode=code[p];
  const unsigned codetype=rcode&DSPH_CODE_MASKTYPE, codeout=rcode&DSPH_CODE_MASKSPECIAL;
  unsigned box;
  if(codetype<DSPH_CODE_TYPE_FLOATING)
    box=(codeout<DSPH_CODE_OUTIGNORE? ((cx<dp.ncx && cy<dp.ncy && cz<dp.ncz)? cellsort: dp.boxboundignore): (codeout==DSPH_CODE_OUTIGNORE? dp.boxboundoutignore: dp.boxboundout));
  else
    box=(codeout<=DSPH_CODE_OUTIGNORE? (codeout<DSPH_CODE_OUTIGNORE? dp.boxfluid+cellsort: dp.boxfluidoutignore): (codetype==DSPH_CODE_TYPE_FLOATING? dp.boxboundout: dp.boxfluidout));
  cellpart[p]=box;
}

// begin[] pre-filled with 0xffffffff; marks first particle of every non-empty box, begin[nctt-1]=n (sentinel)
__global__ void __launch_bounds__(DIV_THREADS) ker_div_cellmark(unsigned n,const unsigned* __restrict__ sortedkeys,const DivParams* __restrict__ dpp,unsigned* __restrict__ begin)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p==0)begin[dpp->nctt-1]=n;
  if(p>=n)return;
  const unsigned k=sortedkeys[p];
  if(p==0 || sortedkeys[p-1]!=k)begin[k]=p;
}

// suffix-min over begin[0..nctt): empty boxes take the begin of the next non-empty box.  Chunked 3-kernel scan with the
// number of elements read from the device (host launches for the maximum).
constexpr int SM_THREADS=256;
constexpr int SM_ITEMS=8;
constexpr int SM_CHUNK=SM_THREADS*SM_ITEMS;
__device__ __forceinline__ unsigned BlockSuffixMinExclusive(unsigned v,unsigned *total){
  // exclusive min over threads with HIGHER threadIdx (suffix), identity 0xffffffff
  __shared__ unsigned wmin[SM_THREADS/32];
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  unsigned inc=v;
  #pragma unroll
  for(int o=1;o<32;o<<=1){ const unsigned t=__shfl_down_sync(0xffffffffu,inc,o); if(lane+o<32)inc=min(inc,t); }
  if(lane==0)wmin[w]=inc;
  __syncthreads();
  if(w==0){
    unsigned s=(lane<SM_THREADS/32? wmin[lane]: 0xffffffffu);
    #pragma unroll
    for(int o=1;o<SM_THREADS/32;o<<=1){ const unsigned t=__shfl_down_sync(0xffffffffu,s,o); if(lane+o<SM_THREADS/32)s=min(s,t); }
    if(lane<SM_THREADS/32)wmin[lane]=s;   // inclusive suffix over warps
  }
  __syncthreads();
  const unsigned wnext=(w+1<SM_THREADS/32? wmin[w+1]: 0xffffffffu);
  unsigned excl=__shfl_down_sync(0xffffffffu,inc,1); if(lane==31)excl=0xffffffffu;
  if(total)*total=wmin[0];
  const unsigned r=min(excl,wnext);
  __syncthreads();
  return r;
}
__global__ void __launch_bounds__(SM_THREADS) ker_div_sufmin_chunks(unsigned* __restrict__ data,const DivParams* __restrict__ dpp,unsigned* __restrict__ chunkmin)
{
  const unsigned n=dpp->nctt;
  const unsigned base=blockIdx.x*SM_CHUNK+threadIdx.x*SM_ITEMS;
  if(blockIdx.x*SM_CHUNK>=n){ if(threadIdx.x==0)chunkmin[blockIdx.x]=0xffffffffu; return; }
  unsigned v[SM_ITEMS]; unsigned s=0xffffffffu;
  #pragma unroll
  for(int i=0;i<SM_ITEMS;i++){ v[i]=(base+i<n? data[base+i]: 0xffffffffu); s=min(s,v[i]); }
  unsigned total;
  unsigned after=BlockSuffixMinExclusive(s,&total);
  #pragma unroll
  for(int i=SM_ITEMS-1;i>=0;i--){ after=min(after,v[i]); if(base+i<n)data[base+i]=after; }
  if(threadIdx.x==0)chunkmin[blockIdx.x]=total;
}
// single block: chunkmin[c] <- min over chunks > c (exclusive suffix), serial over blocks of SM_THREADS from the end
__global__ void __launch_bounds__(SM_THREADS) ker_div_sufmin_sums(unsigned* __restrict__ chunkmin,unsigned m)
{
  __shared__ unsigned carry_s;
  if(threadIdx.x==0)carry_s=0xffffffffu;
  __syncthreads();
  const unsigned nb=(m+SM_THREADS-1)/SM_THREADS;
  for(unsigned b=nb;b-->0;){
    const unsigned i=b*SM_THREADS+threadIdx.x;
    const unsigned v=(i<m? chunkmin[i]: 0xffffffffu);
    unsigned total;
    const unsigned carry=carry_s;
    unsigned ex=BlockSuffixMinExclusive(v,&total);
    ex=min(ex,carry);
    if(i<m)chunkmin[i]=ex;
    __syncthreads();
    if(threadIdx.x==0)carry_s=min(carry,total);
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SM_THREADS) ker_div_sufmin_apply(unsigned* __restrict__ data,const DivParams* __restrict__ dpp,const unsigned* __restrict__ chunkmin)
{
  const unsigned n=dpp->nctt;
  const unsigned base=blockIdx.x*SM_CHUNK+threadIdx.x*SM_ITEMS;
  const unsigned off=chunkmin[blockIdx.x];
  #pragma unroll
  for(int i=0;i<SM_ITEMS;i++)if(base+i<n)data[base+i]=min(data[base+i],off);
}

// largest population of a cell (boundary and fluid cells; the special boxes are never neighbours): the chunked candidate loops
// of the ring force kernels read past a lane's own cell by up to that many entries (the arrays are padded, engine.h POSP_PAD)
__global__ void __launch_bounds__(256) ker_div_maxpop(const unsigned* __restrict__ begin,DivParams* __restrict__ dp)
{
  const unsigned c=blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned nct=dp->nct;
  unsigned v=0;
  if(c<2u*nct+1u && c!=nct)v=begin[c+1]-begin[c];
  v=__reduce_max_sync(0xffffffffu,v);
  if((threadIdx.x&31)==0 && v>dp->maxcellpop)atomicMax(&dp->maxcellpop,v);
}

__global__ void ker_div_counts(const unsigned* __restrict__ begin,unsigned npb12,DivParams* __restrict__ dp)
{
  if(threadIdx.x||blockIdx.x)return;
  dp->npbignore=begin[dp->boxboundignore+1]-begin[dp->boxboundignore];
  dp->npbout=begin[dp->boxboundout+1]-begin[dp->boxboundout];
  dp->npfout=begin[dp->boxfluidout+1]-begin[dp->boxfluidout];
  dp->npboutignore=begin[dp->boxboundoutignore+1]-begin[dp->boxboundoutignore];
  dp->npfoutignore=begin[dp->boxfluidoutignore+1]-begin[dp->boxfluidoutignore];
  dp->npfinal=dp->nptot-dp->npbout-dp->npfout-dp->npboutignore-dp->npfoutignore;
  dp->npbfinal=npb12-dp->npboutignore;
}

// Every particle array is gathered in ONE pass; optional arrays are null.
struct ReorderArrays {
  const unsigned *idp_in; unsigned *idp_out;
  const typecode *code_in; typecode *code_out;
  const unsigned *dcell_in; unsigned *dcell_out;
  const double *pos_in; double *pos_out;            // xyz packed (24 B)
  const float4 *velrhop_in; float4 *velrhop_out;
  const float4 *velrhopm1_in; float4 *velrhopm1_out;   // Verlet
  const double *pospre_in; double *pospre_out;         // Symplectic (mid-step divide)
  const float4 *velrhoppre_in; float4 *velrhoppre_out;
  const float *spstau_in; float *spstau_out;           // 6 floats
  const float4 *boundnormal_in; float4 *boundnormal_out;   // mDBC: normal vectors of the boundary particles (JSphCpuSingle.cpp:464-467)
  float4 *posp;                                        // out: cell-relative xyz + pressure
  uint4 *posh;                                         // out: packed-half culling copy (forces_v10.cuh), one uint4 per 2 particles
};

__device__ __forceinline__ float PressFromRhop(float rhop,const DevCtes &c,bool tail){
  // same rounding points as the reference CPU build (see oracle/sph_oracle.cpp PreInteraction): float product
  // rhop*(1/rhop0), then pow, -1 and *B in double, one rounding to float.
  // tail: the reference BINARY vectorises this loop 4-wide and its scalar epilogue keeps the true division for the last
  // Np-4*((Np-1)/4) particles (1 ulp of the quotient = 0.13 Pa).  That index-dependent artefact is NOT part of the product
  // path (one rule for every particle, so a single domain and x-slabs round alike); it is reproduced only when a test asks for
  // it (dsph_set_refbinary_tail) to compare with dumps of that binary at the 1e-5 level on 2-3 k particle cases.
  const float x=(tail? __fdiv_rn(rhop,c.rhopzero): __fmul_rn(rhop,c.ovrhopzero));
  return float(c.cteb_d*(pow(double(x),c.gamma_d)-1.0));
}

// JSphCpu::ComputePress_NN M/JSphCpu_Tensors.cpp:47-71: EOS of the particle's phase (boundaries: global constants), float
// arithmetic with powf; the power is taken in double and rounded once = the correctly rounded float power of glibc.
__device__ __forceinline__ float PressFromRhopNN(float rhop,unsigned rcode,const DevCtes &c,const DevPhases &ph){
  float rhozero=c.rhopzero, cteb=c.cteb, gamma=c.gamma;
  if((rcode&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID){ const DevPhase &p=ph.ph[rcode&DSPH_CODE_MASKVALUE]; rhozero=p.rho; cteb=p.cteb; gamma=p.gamma; }
  const float rhop_r0=__fdiv_rn(rhop,rhozero);
  const float pw=float(pow(double(rhop_r0),double(gamma)));
  return __fmul_rn(cteb,__fsub_rn(pw,1.0f));
}

template<bool NN> __global__ void __launch_bounds__(DIV_THREADS) ker_div_reorder(unsigned n,const unsigned* __restrict__ sortpart,ReorderArrays a,DevCtes c,DevPhases ph,const DivParams* __restrict__ divp)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  const bool ok=(p<n);
  unsigned npfinal_vec=0xffffffffu;                                // first particle of the reference binary's scalar epilogue (test switch)
  if(c.reftail){ const unsigned npfin=divp->npfinal; npfinal_vec=(npfin? ((npfin-1u)/4u)*4u: 0u); }
  float4 pp=make_float4(0,0,0,0);
  unsigned rcell=0xffffffffu;
  if(ok){
    const unsigned s=sortpart[p];
    a.idp_out[p]=a.idp_in[s];
    const typecode rcode=a.code_in[s];
    a.code_out[p]=rcode;
    rcell=a.dcell_in[s];
    a.dcell_out[p]=rcell;
    const double px=a.pos_in[3*size_t(s)], py=a.pos_in[3*size_t(s)+1], pz=a.pos_in[3*size_t(s)+2];
    a.pos_out[3*size_t(p)]=px; a.pos_out[3*size_t(p)+1]=py; a.pos_out[3*size_t(p)+2]=pz;
    const float4 vr=a.velrhop_in[s];
    a.velrhop_out[p]=vr;
    if(a.velrhopm1_in)a.velrhopm1_out[p]=a.velrhopm1_in[s];
    if(a.pospre_in){
      a.pospre_out[3*size_t(p)]=a.pospre_in[3*size_t(s)]; a.pospre_out[3*size_t(p)+1]=a.pospre_in[3*size_t(s)+1]; a.pospre_out[3*size_t(p)+2]=a.pospre_in[3*size_t(s)+2];
      a.velrhoppre_out[p]=a.velrhoppre_in[s];
    }
    if(a.spstau_in){
      #pragma unroll
      for(int k=0;k<6;k++)a.spstau_out[6*size_t(p)+k]=a.spstau_in[6*size_t(s)+k];
    }
    if(a.boundnormal_in && (rcode&DSPH_CODE_MASKTYPE)!=DSPH_CODE_TYPE_FLUID)a.boundnormal_out[p]=a.boundnormal_in[s];
    // cell-relative position (float) + pressure; excluded particles (dcell=0xffffffff) get zeros
    if(rcell!=0xffffffffu){
      const unsigned cx=DcCellx(c.domcellcode,rcell), cy=DcCelly(c.domcellcode,rcell), cz=DcCellz(c.domcellcode,rcell);
      pp.x=float(px-(c.domposmin[0]+double(cx)*c.scell_d));
      pp.y=float(py-(c.domposmin[1]+double(cy)*c.scell_d));
      pp.z=float(pz-(c.domposmin[2]+double(cz)*c.scell_d));
      pp.w=(NN? PressFromRhopNN(vr.w,rcode,c,ph): PressFromRhop(vr.w,c,p>=npfinal_vec));
    }
    a.posp[p]=pp;
  }
  // packed-half culling copy (forces_v10.cuh): centre-relative coordinates in units of half a cell, n = |rounded xyz|^2;
  // the even lane of each pair of particles writes one uint4 {x0 x1, y0 y1, z0 z1, n0 n1}.  Particles past n or excluded ones
  // are parked far away (never inside anybody's cut-off).
  {
    const bool in=(ok && rcell!=0xffffffffu);
    const __half hx=__float2half_rn(in? pp.x*c.hscale-1.f: 100.f), hy=__float2half_rn(in? pp.y*c.hscale-1.f: 100.f), hz=__float2half_rn(in? pp.z*c.hscale-1.f: 100.f);
    const float fx=__half2float(hx), fy=__half2float(hy), fz=__half2float(hz);
    const __half hn=__float2half_rn(fx*fx+fy*fy+fz*fz);
    const unsigned wx=__half_as_ushort(hx), wy=__half_as_ushort(hy), wz=__half_as_ushort(hz), wn=__half_as_ushort(hn);
    const unsigned lo=wx|(wy<<16), hi=wz|(wn<<16);
    const unsigned lo2=__shfl_down_sync(0xffffffffu,lo,1), hi2=__shfl_down_sync(0xffffffffu,hi,1);   // DIV_THREADS is even: the partner is in the warp
    if(!(p&1u) && p<n+1u){
      uint4 v;
      v.x=(lo&0xffffu)|(lo2<<16); v.y=(lo>>16)|(lo2&0xffff0000u); v.z=(hi&0xffffu)|(hi2<<16); v.w=(hi>>16)|(hi2&0xffff0000u);
      a.posh[p>>1]=v;
    }
  }
}

} // namespace dsph
