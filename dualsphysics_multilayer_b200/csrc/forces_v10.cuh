// forces_v10.cuh -- fluid Interaction_Forces, variant 10 (production): packed-half candidate culling + hit-mask ring.
//
// Why (ncu of variant 9, profiles/r01_forces.md + VERDICT r01): the kernel is instruction-issue bound and a third of its issue
// slots test candidates that are rejected (1078 candidates tested for 173 accepted pairs per fluid particle: ~13 instructions per
// candidate with the per-candidate predicated ring push).  Variant 10 keeps variant 9's structure -- one thread per particle,
// warp-uniform traversal of the 27 neighbour cells in variant 9's row order, accepted pairs deferred in a per-lane ring in
// shared memory, pair body run by (nearly) full warps, bit-identical pair arithmetic -- and replaces the candidate test:
//
//   * a second, 8-byte copy of the cell-relative positions, `posh` (built by the reorder, celldiv.cuh): per PAIR of consecutive
//     particles one uint4 = half2{x0,x1}, half2{y0,y1}, half2{z0,z1}, half2{n0,n1}; coordinates relative to the CENTRE of the
//     particle's cell in units of half a cell (|x|<=1), n = x^2+y^2+z^2 of the rounded coordinates.
//   * |o-p|^2 - |o|^2 = n - 2 o.p : three HFMA2 give that quantity for TWO candidates, one HSET2 compares both with the lane's
//     threshold (2h)^2 + margin - |o|^2, one LOP3 deposits the two result bits in a 16-bit hit mask: 3.5 issue slots per
//     candidate instead of ~13.  The half test is CONSERVATIVE (margin V10_MARGIN covers every rounding of the half pipeline,
//     bound derived in DESIGN.md and checked exhaustively-at-random by tests/test_v10_cull.py): it never drops a true pair; the
//     ~1 % false positives are discarded by the exact fp32 test that the pair body recomputes anyway (same expression and
//     rounding as variants 4-9, so accepted pairs and their order are those of variant 9: results are BIT-IDENTICAL to variant 9).
//   * the ring holds one 8-byte entry per 16-candidate chunk with at least one hit {first index | cell code, hit mask} instead of
//     one 4-byte entry per hit: 16 entries (the same 16 KB per CTA as variant 9) hold ~45 hits instead of 32, which also raises
//     the lane occupancy of the pair rounds.
// Limits as variant 9: ScellDiv=1 (cell = 2h) and np+16 < 2^26 per GPU; otherwise variant 4 runs.
#pragma once
#include "forces_ring.cuh"
#include <cuda_fp16.h>

namespace dsph {

constexpr int V10_Q=16;                                   // ring entries per lane (production; Q is a template parameter for the measured alternatives)
constexpr unsigned V10_STRIDE=FORCES_THREADS*8u;          // bytes between consecutive entries of one lane
constexpr int V10_CHP=8;                                  // uint4 loads (= candidate pairs) per chunk -> 16 candidates
// Conservative margin of the half-precision test in scaled units (cell = 2 units, cut-off^2 = 4): see DESIGN.md "variant 10".
constexpr float V10_MARGIN=0.05f;

template<int Q> __device__ __forceinline__ void V10Push(unsigned qrel,unsigned tail,unsigned e,unsigned m){
  const unsigned a=(tail&((Q-1)*V10_STRIDE))|qrel;
  asm volatile("st.shared.v2.u32 [%0+1024], {%1,%2};" :: "r"(a), "r"(e), "r"(m) : "memory");
}
template<int Q> __device__ __forceinline__ void V10Pop(unsigned qrel,unsigned head,unsigned &e,unsigned &m){
  const unsigned a=(head&((Q-1)*V10_STRIDE))|qrel;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+1024];" : "=r"(e), "=r"(m) : "r"(a) : "memory");
}
// cell-offset table right behind the ring (64 x float4, same contents as variant 9's)
template<int Q> __device__ __forceinline__ float4 V10OffTab(unsigned e){
  float4 r;
  const unsigned a=((e>>22)&0x3f0u);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w) : "r"(a), "n"(V5_SMEM0+Q*V10_STRIDE) : "memory");
  return r;
}
__device__ __forceinline__ __half2 V10AsH2(unsigned v){ return *reinterpret_cast<__half2*>(&v); }

// One pair round: every lane with work takes its next hit (ascending candidate order = FIFO) and runs the pair body.
template<int TKER,int TVISCO,int TDDT,bool SHIFT,int Q> __device__ __forceinline__ void V10Round(
  unsigned q,unsigned &head,unsigned tail,unsigned &cure,unsigned &curm,const float4 &pp1,const FastP1 &fp,const float* taup1,bool boundp2,
  const ForceParams &t,const DevCtes &c,float ks2,PairAcc &a,unsigned &anyb,unsigned &nosh)
{
  if(!curm && head!=tail){ V10Pop<Q>(q,head,cure,curm); head+=V10_STRIDE; }
  if(curm){
    const unsigned b=unsigned(__ffs(int(curm))-1);
    curm&=curm-1u;
    const unsigned e=cure+b;                               // index+b in the low 26 bits (np+16 < 2^26), cell code above
    const unsigned p2=e&V5_IDXMASK;
    const float4 off=V10OffTab<Q>(e);
    const f32x2 oxy=Sub2(Pk(pp1.x,pp1.y),Pk(off.x,off.y)); const float oz=pp1.z-off.z;
    const float4 pp2=__ldg(&t.posp[p2]);
    const float4 vr2=__ldg(&t.velrhop[p2]);
    float drx,dry; const float drz=oz-pp2.z;
    Upk(Sub2(oxy,Pk(pp2.x,pp2.y)),drx,dry);
    const float rr2=drx*drx+dry*dry+drz*drz;               // the expression and rounding of the candidate test of variants 4-9
    unsigned codep2=0; if(SHIFT && boundp2)codep2=t.code[p2];
    const float *taup2=(TVISCO==DSPH_VISCO_LAMINARSPS && !boundp2? t.spstau+6*size_t(p2): nullptr);
    // exact acceptance (cut-off and ALMOSTZERO) without a branch: a rejected pair runs with zero mass and sets no flag
    PairFast<TKER,TVISCO,TDDT,SHIFT,true>(drx,dry,drz,fmaxf(rr2,1e-18f),fp,vr2,pp2.w,boundp2,taup1,taup2,codep2,c,a,anyb,nosh,(rr2>=1e-18f && rr2<=ks2));
  }
}

// resident CTAs per SM the kernel is compiled for: 8 (64 registers, <= 12 bytes of spills, 23.33 ms at DB(20 M); 7 CTAs 23.62 ms, 6 CTAs 24.17 ms) for the
// light instantiations; laminar+SPS and shifting carry 6-10 more accumulators and would spill 30-150 bytes at 64 registers: 6 CTAs (80 registers)
constexpr int V10MinBlocks(int tvisco,bool shift){ return (tvisco==DSPH_VISCO_ARTIFICIAL && !shift)? 8: 6; }

template<int TKER,int TVISCO,int TDDT,bool SHIFT,int MINB=V10MinBlocks(TVISCO,SHIFT),int Q=V10_Q> __global__ void __launch_bounds__(FORCES_THREADS,MINB)
ker_forces_fluid_v10(ForceParams t,DevCtes c)
{
  __shared__ __align__(16) unsigned qsm[Q*FORCES_THREADS*2+64*4];   // ring (8-byte entries) + 64 x float4 cell-offset table
  const unsigned FULL=0xffffffffu;
  if(unsigned(__cvta_generic_to_shared(qsm))!=V5_SMEM0)__trap();         // OR-addressing: the ring must start the CTA's shared window
  if(threadIdx.x<64){
    const unsigned k=threadIdx.x;
    reinterpret_cast<float4*>(qsm+Q*FORCES_THREADS*2)[k]=make_float4((float(int(k&3u))-1.f)*c.scell,(float(int((k>>2)&3u))-1.f)*c.scell,(float(int(k>>4))-1.f)*c.scell,0.f);
  }
  __syncthreads();
  unsigned q=threadIdx.x*8u;
  asm volatile("" : "+r"(q));
  unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  unsigned p1c=min(p1,t.np-1);
  bool active;
  if(t.ownlist){                                           // x-slab halo copies take no lanes: thread -> owned fluid particle
    const unsigned nown=__ldg(t.nown), tid=blockIdx.x*blockDim.x+threadIdx.x;
    if(blockIdx.x*blockDim.x>=nown)return;
    active=(tid<nown);
    p1=p1c=__ldg(t.ownlist+min(tid,nown-1u));
  }
  else active=(p1<t.np && !(t.skipghosts && (t.code[p1c]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC));
  const float4 pp1=t.posp[p1c];
  FastP1 fp; fp.velrhop=t.velrhop[p1c]; fp.press=pp1.w; fp.ovrhop=1.f/fp.velrhop.w;
  const unsigned rcell=t.dcell[p1c];
  const int cx=int(DcCellx(c.domcellcode,rcell))-int(t.cellzero[0]);
  const int cy=int(DcCelly(c.domcellcode,rcell))-int(t.cellzero[1]);
  const int cz=int(DcCellz(c.domcellcode,rcell))-int(t.cellzero[2]);
  float taup1[6]={0,0,0,0,0,0};
  if(TVISCO==DSPH_VISCO_LAMINARSPS){
    #pragma unroll
    for(int k=0;k<6;k++)taup1[k]=t.spstau[6*size_t(p1c)+k];
  }
  const float ks2=c.kernelsize2;
  // particle 1 in the centred, scaled frame of its own cell (cell = [-1,1]^3); a neighbour cell at offset d is 2*d away
  const float hs=c.hscale;                                 // 2/scell
  const float bx=pp1.x*hs-1.f, by=pp1.y*hs-1.f, bz=pp1.z*hs-1.f;
  const float thr0=c.kernelsize2*hs*hs+V10_MARGIN;
  const uint4* __restrict__ posh=t.posh;
  PairAcc a; ZeroAcc(a);
  unsigned anyb=0,nosh=0,head=0,tail=0,cure=0,curm=0;
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const bool boundp2=(pass==1);
    const unsigned cellinit=(boundp2? 0u: t.cellfluid);
    #pragma unroll 1
    for(int r=0;r<9;r++){
      const int dz=int((V5_ROWDZ>>(2*r))&3u)-1;
      const int dy=int((V5_ROWDY>>(2*r))&3u)-1;
      const int z=cz+dz, y=cy+dy;
      const float osz=bz-2.f*float(dz), osy=by-2.f*float(dy);
      const float thr1=thr0-osz*osz-osy*osy;
      const __half2 mz=__float2half2_rn(-2.f*osz), my=__float2half2_rn(-2.f*osy);
      const bool rowok=(active && z>=0 && z<t.ncz && y>=0 && y<t.ncy);
      #pragma unroll 1
      for(int dx=-1;dx<=1;dx++){
        const int x=cx+dx;
        unsigned pini=0,pfin=0;
        if(rowok && x>=0 && x<t.ncx){
          const unsigned v=unsigned(t.nsheet*z+t.ncx*y+x)+cellinit;
          pini=t.begincell[v]; pfin=t.begincell[v+1];
        }
        const float osx=bx-2.f*float(dx);
        const __half2 mx=__float2half2_rn(-2.f*osx);
        const __half2 thr=__half2half2(__float2half_ru(thr1-osx*osx));   // rounded up: conservative
        const unsigned codebits=(unsigned(dx+1)|(unsigned(dy+1)<<2)|(unsigned(dz+1)<<4))<<26;
        const unsigned i0=pini&~1u;                                      // chunks start at an even index (one uint4 = two particles)
        const unsigned nk=(pfin>pini? (pfin-i0+15u)>>4: 0u);             // chunks of 16 candidates of this lane
        const unsigned nkmax=__reduce_max_sync(FULL,nk);
        const unsigned klast=(nk? nk-1u: 0u);
        const unsigned oddstart=pini&1u;
        #pragma unroll 1
        for(unsigned k=0;k<nkmax;k++){
          while(__any_sync(FULL,(tail-head)>=unsigned(Q)*V10_STRIDE))
            V10Round<TKER,TVISCO,TDDT,SHIFT,Q>(q,head,tail,cure,curm,pp1,fp,taup1,boundp2,t,c,ks2,a,anyb,nosh);
          // lanes whose own range is exhausted re-read their last chunk (in bounds) and mask everything out
          const unsigned base=i0+(min(k,klast)<<4);
          const uint4* __restrict__ ph=posh+(base>>1);
          unsigned acc=0;
          #pragma unroll
          for(int u=0;u<V10_CHP;u++){
            const uint4 v=__ldg(ph+u);
            __half2 s=__hfma2(mx,V10AsH2(v.x),V10AsH2(v.w));
            s=__hfma2(my,V10AsH2(v.y),s);
            s=__hfma2(mz,V10AsH2(v.z),s);
            const unsigned m=__hle2_mask(s,thr);                         // 0xffff per half that passes
            acc|=m&((1u<<(2*u))|(1u<<(17+2*u)));                          // even particle -> bit 2u, odd particle -> bit 16+2u+1
          }
          unsigned hits=(acc|(acc>>16))&0xffffu;                         // bit b <-> candidate base+b
          const int cnt=int(pfin)-int(i0+(k<<4));                         // candidates of this lane from this chunk on
          const unsigned nv=unsigned(min(max(cnt,0),16));
          unsigned valid=(1u<<nv)-1u;
          if(k==0)valid&=~oddstart;
          hits&=valid;
          if(hits){ V10Push<Q>(q,tail,base|codebits,hits); tail+=V10_STRIDE; }
        }
      }
    }
    // end of the pass: the boundary pass uses other constants, drain everything
    while(__any_sync(FULL,(head!=tail)||curm))
      V10Round<TKER,TVISCO,TDDT,SHIFT,Q>(q,head,tail,cure,curm,pp1,fp,taup1,boundp2,t,c,ks2,a,anyb,nosh);
    if(pass==0)FluidPassStore<TVISCO,SHIFT>(a);
  }
  float viscmax=0,ace2=0,vel2=0;
  if(active){
    float ar=a.ar, ax=a.ax, ay=a.ay, az=a.az;
    if(c.simulate2d)ay=0;
    if(TDDT!=DSPH_DDT_NONE && !anyb)ar+=a.delta;
    t.acear[p1]=make_float4(ax,ay,az,ar);
    if(SHIFT){ float4 s=a.shift; if(nosh)s.x=FLT_MAX; t.shiftposfs[p1]=s; }
    if(TVISCO==DSPH_VISCO_LAMINARSPS){
      const float g[6]={a.gxx,a.gxy,a.gxz,a.gyy,a.gyz,a.gzz};
      #pragma unroll
      for(int k=0;k<6;k++)t.spsgradvel[6*size_t(p1)+k]=g[k];
    }
    viscmax=a.visc;
    const bool normal=(!c.ghosts || (t.code[p1]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
    if(normal)ace2=ax*ax+ay*ay+az*az;
    vel2=fp.velrhop.x*fp.velrhop.x+fp.velrhop.y*fp.velrhop.y+fp.velrhop.z*fp.velrhop.z;
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
}

} // namespace dsph
