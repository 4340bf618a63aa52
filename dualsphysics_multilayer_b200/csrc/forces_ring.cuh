// forces_ring.cuh -- fluid Interaction_Forces with a deferred pair ring: variant 9 (cross-check of variant 10, fallback for
// np >= 2^26-16) and the traversal skeleton of the multiphase ring kernels (forces_nn.cuh).
//
// ncu on variant 4 (profiles/r01_forces.md): issue bound, and 68 % of the issue slots go to the pair body executed with only
// 8.9 of 32 lanes active (a warp's 32 particles fill most of a 2h cell, so a candidate is inside the cut-off of a few
// lanes only; 620 body executions per warp for 173 accepted pairs per particle).  Here the body is deferred:
//   TEST   the broadcast candidate loop of variant 4 (one thread per particle) only tests the distance and pushes accepted
//          neighbours as 4-byte entries {particle index : 26 bits | neighbour-cell code : 6 bits} into the lane's own ring in
//          shared memory (q[slot][lane]: conflict free; 32 slots = 16 KB per 128-thread block),
//   PAIR   when some lane's ring is nearly full every lane pops one entry, re-reads posp/velrhop of that neighbour (L1/L2
//          hits), rebuilds dr with exactly the operations of the test and runs the division-free pair body.
// The ring is FIFO per lane; control flow is warp uniform (fixed 3x3x3 cell loops, trip counts by redux.sync) so that the
// votes are legal.  Candidate test: chunks of 8 unrolled loads from the padded posp array (no index clamp); full chunks skip
// the range check; the ALMOSTZERO exclusion lives in the pair body as a zero-mass pair.  Ring slots are addressed by one LOP3
// (OR of the lane offset, ring base folded into the ld/st immediate), the neighbour-cell offsets of the pair body come from a
// 1 KB shared table, x,y components use the packed fp32x2 instructions of sm_100 (FADD2/FMUL2/FFMA2, bit-identical to the
// scalar forms), 72 registers / 7 CTAs per SM.  Rows of three neighbour cells are visited so that opposite rows follow each
// other (evens out the arrival of accepted pairs between the lanes).  The history of the measured variants 5-9 (and the
// rejected 2, 3) is in profiles/r01_forces.md; their code was removed from the product in round 2.
// Limits: cells of size 2h (ScellDiv=1) and np < 2^26 per GPU; otherwise variant 4 is used.
#pragma once
#include "forces_fast.cuh"

namespace dsph {

constexpr int V5_CH=8;      // candidates per unrolled chunk (4 was measured: 1.7 % slower)
constexpr int V5_Q=32;      // ring slots per lane
constexpr unsigned V5_IDXMASK=(1u<<26)-1u;
constexpr unsigned V5_STRIDE=FORCES_THREADS*4u;    // bytes between consecutive slots of one lane

// The ring of a CTA is the only static shared array besides the 1 KB offset table behind it, so it starts at V5_SMEM0 (the first
// byte after the 1 KB the system reserves per CTA; checked at kernel entry, trap otherwise; the engine probes it once per context
// and falls back to variant 4 with a line in the log) and is 16 KB-aligned relative to that:
// slot address = ((tail & mask) | lane part) + V5_SMEM0 = ONE LOP3 with the constant folded into the ld/st.shared offset.
constexpr unsigned V5_SMEM0=0x400u;
// row order, two bits per row (value+1): (dz,dy) = (-1,-1),(1,1),(-1,1),(1,-1),(0,0),(-1,0),(1,0),(0,-1),(0,1)
// (measured at DB(20 M): 28.90 ms; z,y ascending 31.75 ms; centre row last 29.11 ms)
constexpr unsigned V5RowPack(int a0,int a1,int a2,int a3,int a4,int a5,int a6,int a7,int a8){
  return unsigned(a0+1)|(unsigned(a1+1)<<2)|(unsigned(a2+1)<<4)|(unsigned(a3+1)<<6)|(unsigned(a4+1)<<8)|(unsigned(a5+1)<<10)|(unsigned(a6+1)<<12)|(unsigned(a7+1)<<14)|(unsigned(a8+1)<<16);
}
constexpr unsigned V5_ROWDZ=V5RowPack(-1,1,-1,1,0,-1,1,0,0);
constexpr unsigned V5_ROWDY=V5RowPack(-1,1,1,-1,0,0,0,-1,1);
__device__ __forceinline__ void V5PushOr(unsigned qrel,unsigned tail,unsigned v){
  const unsigned a=(tail&((V5_Q-1)*V5_STRIDE))|qrel;
  asm volatile("st.shared.u32 [%0+1024], %1;" :: "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned V5PopOr(unsigned qrel,unsigned head){
  const unsigned a=(head&((V5_Q-1)*V5_STRIDE))|qrel;
  unsigned v; asm volatile("ld.shared.u32 %0, [%1+1024];" : "=r"(v) : "r"(a) : "memory"); return v;
}
// entry e = {index:26 | cell code:6}: table slot of the code = (e>>22)&0x3f0, table right behind the ring
__device__ __forceinline__ float4 V5OffTab(unsigned e){
  float4 r;
  const unsigned a=((e>>22)&0x3f0u);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w) : "r"(a), "n"(V5_SMEM0+V5_Q*V5_STRIDE) : "memory");
  return r;
}

// one chunk of V5_CH candidates: vb = first index | cell code (index < 2^26, so vb+u carries index+u), vfin = end | code
template<bool CHECK> __device__ __forceinline__ void V5Chunk(const float4* __restrict__ posp,unsigned vb,unsigned vfin,
  float ox,float oy,float oz,float ks2,unsigned q,unsigned &tail)
{
  const float4* __restrict__ ppb=posp+(vb&V5_IDXMASK);     // posp is padded: a chunk may read past the lane's own range
  const f32x2 oxy=Pk(ox,oy);
  #pragma unroll
  for(int u=0;u<V5_CH;u++){
    const float4 pp2=__ldg(ppb+u);
    float drx,dry; const float drz=oz-pp2.z;
    Upk(Sub2(oxy,Pk(pp2.x,pp2.y)),drx,dry);                // FADD2: same rounding as the scalar form, one issue slot
    const float rr2=drx*drx+dry*dry+drz*drz;
    if((!CHECK || vb+u<vfin) && rr2<=ks2){
      V5PushOr(q,tail,vb+u);
      tail+=V5_STRIDE;
    }
  }
}

// Shared memory of a ring CTA: V5_Q*FORCES_THREADS ring words + 64 float4 of cell offsets; must be the kernel's only static shared array.
struct V8Smem { unsigned w[V5_Q*FORCES_THREADS+64*4]; };

// fills the offset table, checks the ring's place in the shared window and returns the lane's (pinned) ring offset
__device__ __forceinline__ unsigned V8Setup(unsigned* qsm,float scell){
  if(unsigned(__cvta_generic_to_shared(qsm))!=V5_SMEM0)__trap();
  if(threadIdx.x<64){
    const unsigned k=threadIdx.x;
    reinterpret_cast<float4*>(qsm+V5_Q*FORCES_THREADS)[k]=make_float4((float(int(k&3u))-1.f)*scell,(float(int((k>>2)&3u))-1.f)*scell,(float(int(k>>4))-1.f)*scell,0.f);
  }
  __syncthreads();
  unsigned q=threadIdx.x*4u;
  asm volatile("" : "+r"(q));                    // keep it in a register (the compiler re-derived it from SR_TID in every chunk)
  return q;
}

// pair(p2,drx,dry,drz,rr2 clamped,press of p2,boundp2,valid) is called for every accepted neighbour (fluid pass, then boundary pass;
// rows in the order above, cells and particles ascending inside a row)
template<class PAIR> __device__ __forceinline__ void V8Traverse(const ForceParams &t,const DevCtes &c,bool active,const float4 &pp1,unsigned rcell,
  unsigned q,PAIR &&pair)
{
  const unsigned FULL=0xffffffffu;
  const int cx=int(DcCellx(c.domcellcode,rcell))-int(t.cellzero[0]);
  const int cy=int(DcCelly(c.domcellcode,rcell))-int(t.cellzero[1]);
  const int cz=int(DcCellz(c.domcellcode,rcell))-int(t.cellzero[2]);
  const float ks2=c.kernelsize2, scell=c.scell;
  unsigned head=0,tail=0;
  auto round=[&](bool boundp2){
    if(head!=tail){
      const unsigned e=V5PopOr(q,head);
      head+=V5_STRIDE;
      const unsigned p2=e&V5_IDXMASK;
      const float4 off=V5OffTab(e);
      const float4 pp2=__ldg(&t.posp[p2]);
      float drx,dry;
      Upk(Sub2(Sub2(Pk(pp1.x,pp1.y),Pk(off.x,off.y)),Pk(pp2.x,pp2.y)),drx,dry);
      const float drz=(pp1.z-off.z)-pp2.z;
      const float rr2=drx*drx+dry*dry+drz*drz;
      // the ALMOSTZERO exclusion (the particle itself, coincident particles) is applied here, not in the candidate test, and
      // without a branch: such a pair runs with rr2 clamped and zero mass (every contribution is exactly 0, no flag is set)
      pair(p2,drx,dry,drz,fmaxf(rr2,1e-18f),pp2.w,boundp2,rr2>=1e-18f);
    }
  };
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const bool boundp2=(pass==1);
    const unsigned cellinit=(boundp2? 0u: t.cellfluid);
    #pragma unroll 1
    for(int r=0;r<9;r++){
      const int dz=int((V5_ROWDZ>>(2*r))&3u)-1;
      const int dy=int((V5_ROWDY>>(2*r))&3u)-1;
      const int z=cz+dz;
      const float oz=pp1.z-float(dz)*scell;
      const int y=cy+dy;
      const float oy=pp1.y-float(dy)*scell;
      const bool rowok=(active && z>=0 && z<t.ncz && y>=0 && y<t.ncy);
      #pragma unroll 1
      for(int dx=-1;dx<=1;dx++){
        const int x=cx+dx;
        const float ox=pp1.x-float(dx)*scell;
        unsigned pini=0,pfin=0;
        if(rowok && x>=0 && x<t.ncx){
          const unsigned v=unsigned(t.nsheet*z+t.ncx*y+x)+cellinit;
          pini=t.begincell[v]; pfin=t.begincell[v+1];
        }
        // the chunk loop runs on vb = index | cell code (what is pushed), split into the chunks every lane has in full (one
        // CREDUX.MIN per cell instead of a vote per chunk) and the ragged rest
        const unsigned n=pfin-pini;
        const unsigned nmax=__reduce_max_sync(FULL,n);
        const unsigned ifull=(__reduce_min_sync(FULL,n)/V5_CH)*V5_CH;
        const unsigned codebits=(unsigned(dx+1)|(unsigned(dy+1)<<2)|(unsigned(dz+1)<<4))<<26;
        unsigned vb=pini|codebits;
        const unsigned vfin=pfin|codebits;
        unsigned i=0;
        for(;i<ifull;i+=V5_CH,vb+=V5_CH){
          while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))round(boundp2);
          V5Chunk<false>(t.posp,vb,vfin,ox,oy,oz,ks2,q,tail);
        }
        for(;i<nmax;i+=V5_CH,vb+=V5_CH){
          while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))round(boundp2);
          V5Chunk<true>(t.posp,vb,vfin,ox,oy,oz,ks2,q,tail);
        }
      }
    }
    // end of the pass: the boundary pass uses other constants, drain everything
    while(__any_sync(FULL,head!=tail))round(boundp2);
  }
}

// ---- variant 9: single-phase fluid kernel on the skeleton above ----
template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS,6)
ker_forces_fluid_v9(ForceParams t,DevCtes c)
{
  __shared__ __align__(16) V8Smem qsm;
  const unsigned q=V8Setup(qsm.w,c.scell);
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned p1c=min(p1,t.np-1);
  // x-slab halo copies are neighbours only: their forces are computed by the owning GPU (skipghosts: multi-GPU mode)
  const bool active=(p1<t.np && !(t.skipghosts && (t.code[p1c]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC));
  const float4 pp1=t.posp[p1c];
  FastP1 fp; fp.velrhop=t.velrhop[p1c]; fp.press=pp1.w; fp.ovrhop=1.f/fp.velrhop.w;
  float taup1[6]={0,0,0,0,0,0};
  if(TVISCO==DSPH_VISCO_LAMINARSPS){
    #pragma unroll
    for(int k=0;k<6;k++)taup1[k]=t.spstau[6*size_t(p1c)+k];
  }
  PairAcc a; ZeroAcc(a);
  unsigned anyb=0,nosh=0,ffdone=0;
  V8Traverse(t,c,active,pp1,t.dcell[p1c],q,[&](unsigned p2,float drx,float dry,float drz,float rr2,float pressp2,bool boundp2,bool valid){
    if(boundp2 && !ffdone){ ffdone=1u; FluidPassStore<TVISCO,SHIFT>(a); }   // the ring was drained at the end of the fluid-fluid pass
    const float4 vr2=__ldg(&t.velrhop[p2]);
    unsigned codep2=0; if(SHIFT && boundp2)codep2=t.code[p2];
    const float *taup2=(TVISCO==DSPH_VISCO_LAMINARSPS && !boundp2? t.spstau+6*size_t(p2): nullptr);
    PairFast<TKER,TVISCO,TDDT,SHIFT,true>(drx,dry,drz,rr2,fp,vr2,pressp2,boundp2,taup1,taup2,codep2,c,a,anyb,nosh,valid);
  });
  if(!ffdone)FluidPassStore<TVISCO,SHIFT>(a);
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
