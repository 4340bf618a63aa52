// forces_v5.cuh -- fluid Interaction_Forces with a deferred pair ring (variant 5).
//
// ncu on variant 4 (profiles/r01_forces.md): issue bound, and 68 % of the issue slots go to the pair body executed with only
// 8.9 of 32 lanes active (a warp's 32 particles fill most of a 2h cell, so a candidate is inside the cut-off of a few
// lanes only; 620 body executions per warp for 173 accepted pairs per particle).  Here the body is deferred:
//   TEST   the broadcast candidate loop of variant 4 (one thread per particle, oracle pair order) only tests the distance and
//          pushes accepted neighbours as 4-byte entries {particle index : 26 bits | neighbour-cell code : 6 bits} into the
//          lane's own ring in shared memory (q[slot][lane]: conflict free; 32 or 64 slots = 16 or 32 KB per 128-thread block),
//   PAIR   when some lane's ring is nearly full every lane pops one entry, re-reads posp/velrhop of that neighbour (L1/L2
//          hits), rebuilds dr with exactly the operations of the test and runs the division-free pair body.
// The ring evens out the bursts (tools/simq.py: 227 rounds per warp with 24.5 lanes active at 64 slots, 282 at 32 slots, instead of
// 620 immediate executions with 8.9 lanes).  Production is the 32-slot ring (variants 6 and 7): 16 KB of shared memory per CTA lets 8
// CTAs (32 warps) share an SM, which is worth more than the deeper ring once the candidate test is lean (profiles/r01_forces.md).
// The ring is FIFO per lane, so the summation order per particle is the oracle's and the result is bit-identical to
// variant 4.  Control flow is warp uniform (fixed 3x3x3 cell loops, trip counts by redux.sync) so that the votes are legal.
// Candidate test: chunks of 8 unrolled loads from the padded posp array (no index clamp); when every lane has a full chunk the
// range check is skipped; the ALMOSTZERO exclusion lives in the pair body as a zero-mass pair (no branch between the two
// neighbour loads).  Limits: cells of size 2h (ScellDiv=1) and np < 2^26 per GPU; otherwise variant 4 is used.
// Variant 7 (OPT=1, production) is variant 6 with fewer issue slots around the same arithmetic (37.07 -> 34.12 ms at 20 M, bit-identical):
// ring slots addressed by one LOP3 (OR instead of AND+ADD, the ring base folded into the ld/st offset), the pushed word kept as a
// loop variable (index | cell code, +u per candidate), the neighbour-cell offsets of the pair body read from a 1 KB shared table
// (LDS.128 + 3 FADD instead of 13 shift/convert instructions), the lane's ring offset pinned in a register, and the full chunks of
// a cell counted once (CREDUX.MIN) instead of a vote per chunk.
// Variant 8 (OPT=2, production, 31.73 ms) adds the packed fp32x2 instructions of sm_100 (FADD2/FMUL2/FFMA2: two IEEE operations per
// issue slot, bit-identical to the scalar forms) for the x,y components of dr, dv, fr and the ace accumulation in the pair body and
// of dr in the candidate test, selects the artificial-viscosity term of receding pairs to zero instead of branching around it, and is
// compiled for 7 resident CTAs per SM (72 registers, no spills: __launch_bounds__(128,6); 8 CTAs = 64 registers spill 20 bytes and
// measure 32.9 ms, 6 CTAs with 78 registers 33.2 ms).
#pragma once
#include "forces_fast.cuh"

namespace dsph {

constexpr int V5_CH=8;      // candidates per unrolled chunk (4 was measured: 1.7 % slower at Q=32, 9 % slower at Q=64)
constexpr unsigned V5_IDXMASK=(1u<<26)-1u;

// Ring slots are addressed with 32-bit shared-window addresses and explicit ld/st.shared: through a generic pointer the
// compiler rebuilt the shared base (S2R SR_CgaCtaId + 4 integer instructions) for every push and pop (ncu source page).
constexpr unsigned V5_STRIDE=FORCES_THREADS*4u;    // bytes between consecutive slots of one lane
// head and tail count in BYTES of the lane's ring (slot k lives at k*V5_STRIDE), so a slot address is one AND + one ADD
__device__ __forceinline__ void V5Push(unsigned qaddr,unsigned byteoff,unsigned v){
  asm volatile("st.shared.u32 [%0], %1;" :: "r"(qaddr+byteoff), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned V5Pop(unsigned qaddr,unsigned byteoff){
  unsigned v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(qaddr+byteoff) : "memory"); return v;
}

// OPT=1 (variant 7).  The ring of a CTA is the only static shared array besides the 1 KB offset table behind it, so it starts
// at V5_SMEM0 (the first byte after the 1 KB the system reserves per CTA; checked at kernel entry, trap otherwise) and is
// 16 KB-aligned relative to that: slot address = ((tail & mask) | lane part) + V5_SMEM0 = ONE LOP3 with the constant folded
// into the ld/st.shared offset, instead of AND + ADD.
constexpr unsigned V5_SMEM0=0x400u;
// row order of variant 9, two bits per row (value+1): (dz,dy) = (-1,-1),(1,1),(-1,1),(1,-1),(0,0),(-1,0),(1,0),(0,-1),(0,1)
// (measured at DB(20 M): 28.90 ms; centre row last 29.11 ms; best order of a random search in the queue simulation 29.33 ms)
constexpr unsigned V5RowPack(int a0,int a1,int a2,int a3,int a4,int a5,int a6,int a7,int a8){
  return unsigned(a0+1)|(unsigned(a1+1)<<2)|(unsigned(a2+1)<<4)|(unsigned(a3+1)<<6)|(unsigned(a4+1)<<8)|(unsigned(a5+1)<<10)|(unsigned(a6+1)<<12)|(unsigned(a7+1)<<14)|(unsigned(a8+1)<<16);
}
constexpr unsigned V5_ROWDZ=V5RowPack(-1,1,-1,1,0,-1,1,0,0);
constexpr unsigned V5_ROWDY=V5RowPack(-1,1,1,-1,0,0,0,-1,1);
template<int V5_Q> __device__ __forceinline__ void V5PushOr(unsigned qrel,unsigned tail,unsigned v){
  const unsigned a=(tail&((V5_Q-1)*V5_STRIDE))|qrel;
  asm volatile("st.shared.u32 [%0+1024], %1;" :: "r"(a), "r"(v) : "memory");
}
template<int V5_Q> __device__ __forceinline__ unsigned V5PopOr(unsigned qrel,unsigned head){
  const unsigned a=(head&((V5_Q-1)*V5_STRIDE))|qrel;
  unsigned v; asm volatile("ld.shared.u32 %0, [%1+1024];" : "=r"(v) : "r"(a) : "memory"); return v;
}
// entry e = {index:26 | cell code:6}: table slot of the code = ((e>>22)&0x3f0) | window base, table right behind the ring
template<int V5_Q> __device__ __forceinline__ float4 V5OffTab(unsigned e,unsigned ubase){
  float4 r;
  const unsigned a=((e>>22)&0x3f0u); (void)ubase;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w) : "r"(a), "n"(V5_SMEM0+V5_Q*V5_STRIDE) : "memory");
  return r;
}

// one chunk of V5_CH candidates of variant 7: vb = first index | cell code (index < 2^26, so vb+u carries index+u), vfin = end | code
template<int V5_Q,bool CHECK,bool PACKED> __device__ __forceinline__ void V5Chunk(const float4* __restrict__ posp,unsigned vb,unsigned vfin,
  float ox,float oy,float oz,float ks2,unsigned q,unsigned &tail)
{
  const float4* __restrict__ ppb=posp+(vb&V5_IDXMASK);     // posp is padded: a chunk may read up to V5_CH-1 entries past the range
  const f32x2 oxy=Pk(ox,oy);
  #pragma unroll
  for(int u=0;u<V5_CH;u++){
    const float4 pp2=__ldg(ppb+u);
    float drx,dry; const float drz=oz-pp2.z;
    if(PACKED)Upk(Sub2(oxy,Pk(pp2.x,pp2.y)),drx,dry); else{ drx=ox-pp2.x; dry=oy-pp2.y; }   // FADD2: same rounding, one issue slot
    const float rr2=drx*drx+dry*dry+drz*drz;
    if((!CHECK || vb+u<vfin) && rr2<=ks2){
      V5PushOr<V5_Q>(q,tail,vb+u);
      tail+=V5_STRIDE;
    }
  }
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT,int V5_Q,int OPT> __device__ __forceinline__ void V5Round(
  unsigned q,unsigned &head,unsigned tail,const float4 &pp1,const FastP1 &fp,const float* taup1,bool boundp2,
  const ForceParams &t,const DevCtes &c,PairAcc &a,unsigned &anyb,unsigned &nosh)
{
  if(head!=tail){
    const unsigned e=(OPT? V5PopOr<V5_Q>(q,head): V5Pop(q,head&((V5_Q-1)*V5_STRIDE)));
    head+=V5_STRIDE;
    const unsigned p2=e&V5_IDXMASK;
    const unsigned code=e>>26;
    float ox,oy,oz;
    f32x2 oxy=0;
    if(OPT>=2){
      const float4 off=V5OffTab<V5_Q>(e,q&0xff000000u);
      oxy=Sub2(Pk(pp1.x,pp1.y),Pk(off.x,off.y)); oz=pp1.z-off.z; ox=oy=0;
    }
    else if(OPT){
      // cell-origin offsets of the 27 neighbour cells from a 1 KB table in shared memory (one LDS.128 + 3 FADD instead of
      // 13 integer/convert instructions; fd*scell is exact, so pp1 - table entry rounds like the FFMA of the candidate test)
      const float4 off=V5OffTab<V5_Q>(e,q&0xff000000u);
      ox=pp1.x-off.x; oy=pp1.y-off.y; oz=pp1.z-off.z;
    }
    else{
      const float fdx=float(int(code&3u))-1.f, fdy=float(int((code>>2)&3u))-1.f, fdz=float(int(code>>4))-1.f;
      ox=pp1.x-fdx*c.scell; oy=pp1.y-fdy*c.scell; oz=pp1.z-fdz*c.scell;
    }
    const float4 pp2=__ldg(&t.posp[p2]);
    const float4 vr2=__ldg(&t.velrhop[p2]);
    float drx,dry; const float drz=oz-pp2.z;
    if(OPT>=2)Upk(Sub2(oxy,Pk(pp2.x,pp2.y)),drx,dry); else{ drx=ox-pp2.x; dry=oy-pp2.y; }
    const float rr2=drx*drx+dry*dry+drz*drz;
    unsigned codep2=0; if(SHIFT && boundp2)codep2=t.code[p2];
    const float *taup2=(TVISCO==DSPH_VISCO_LAMINARSPS && !boundp2? t.spstau+6*size_t(p2): nullptr);
    // the ALMOSTZERO exclusion (the particle itself, coincident particles) is applied here, not in the candidate test, and
    // without a branch: such a pair runs with rr2 clamped and zero mass (every contribution is exactly 0, no flag is set)
    PairFast<TKER,TVISCO,TDDT,SHIFT,(OPT>=2)>(drx,dry,drz,fmaxf(rr2,1e-18f),fp,vr2,pp2.w,boundp2,taup1,taup2,codep2,c,a,anyb,nosh,rr2>=1e-18f);
  }
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT,int V5_Q,int OPT,int MINB=(V5_Q>=128? 3: (V5_Q>=64? 6: 8))> __global__ void __launch_bounds__(FORCES_THREADS,MINB)
ker_forces_fluid_v5(ForceParams t,DevCtes c)
{
  __shared__ __align__(16) unsigned qsm[V5_Q*FORCES_THREADS+(OPT? 64*4: 0)];   // ring [+ 64 x float4 cell-offset table]
  const unsigned FULL=0xffffffffu;
  unsigned q=unsigned(__cvta_generic_to_shared(qsm));
  if(OPT){
    if(q!=V5_SMEM0)__trap();                      // the OR-addressing needs the ring at the first usable byte of the CTA's window (no cluster launch)
    if(threadIdx.x<64){
      const unsigned k=threadIdx.x;
      float4* tab=reinterpret_cast<float4*>(qsm+V5_Q*FORCES_THREADS);
      tab[k]=make_float4((float(int(k&3u))-1.f)*c.scell,(float(int((k>>2)&3u))-1.f)*c.scell,(float(int(k>>4))-1.f)*c.scell,0.f);
    }
    __syncthreads();
    q=threadIdx.x*4u;                            // lane part; slots are OR-ed in, V5_SMEM0 is an immediate of the ld/st
    asm volatile("" : "+r"(q));                  // keep it in a register (the compiler re-derived it from SR_TID in every chunk)
  }
  else q+=threadIdx.x*4u;
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned p1c=min(p1,t.np-1);
  // x-slab halo copies are neighbours only: their forces are computed by the owning GPU (skipghosts: multi-GPU mode)
  const bool active=(p1<t.np && !(t.skipghosts && (t.code[p1c]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC));
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
  const float ks2=c.kernelsize2, scell=c.scell;
  PairAcc a; ZeroAcc(a);
  unsigned anyb=0,nosh=0,head=0,tail=0;
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const bool boundp2=(pass==1);
    const unsigned cellinit=(boundp2? 0u: t.cellfluid);
    #pragma unroll 1
    for(int r=0;r<9;r++){
      {
        // rows (dz,dy) of three cells: z,y ascending = the oracle's order, or (OPT>=3, variant 9) opposite rows back to back, which
        // evens out the arrival of accepted pairs between the lanes of a warp (tools/simq.py: 284 -> 250 pair rounds per warp)
        const int dz=(OPT>=3? int((V5_ROWDZ>>(2*r))&3u)-1: r/3-1);
        const int dy=(OPT>=3? int((V5_ROWDY>>(2*r))&3u)-1: r%3-1);
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
          const unsigned nmax=__reduce_max_sync(FULL,pfin-pini);
          const unsigned codebits=(unsigned(dx+1)|(unsigned(dy+1)<<2)|(unsigned(dz+1)<<4))<<26;
          if(OPT){
            // variant 7: the chunk loop runs on vb = index | cell code (what is pushed), split into the chunks every lane has in
            // full (one CREDUX.MIN per cell instead of a vote per chunk) and the ragged rest
            const unsigned n=pfin-pini;
            const unsigned ifull=(__reduce_min_sync(FULL,n)/V5_CH)*V5_CH;
            unsigned vb=pini|codebits;
            const unsigned vfin=pfin|codebits;
            unsigned i=0;
            for(;i<ifull;i+=V5_CH,vb+=V5_CH){
              while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))
                V5Round<TKER,TVISCO,TDDT,SHIFT,V5_Q,OPT>(q,head,tail,pp1,fp,taup1,boundp2,t,c,a,anyb,nosh);
              V5Chunk<V5_Q,false,(OPT>=2)>(t.posp,vb,vfin,ox,oy,oz,ks2,q,tail);
            }
            for(;i<nmax;i+=V5_CH,vb+=V5_CH){
              while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))
                V5Round<TKER,TVISCO,TDDT,SHIFT,V5_Q,OPT>(q,head,tail,pp1,fp,taup1,boundp2,t,c,a,anyb,nosh);
              V5Chunk<V5_Q,true,(OPT>=2)>(t.posp,vb,vfin,ox,oy,oz,ks2,q,tail);
            }
          }
          else
          for(unsigned i=0;i<nmax;i+=V5_CH){
            while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))
              V5Round<TKER,TVISCO,TDDT,SHIFT,V5_Q,OPT>(q,head,tail,pp1,fp,taup1,boundp2,t,c,a,anyb,nosh);
            // posp is padded (engine.cu): a chunk may read up to V5_CH-1 entries past the range without a clamp
            const unsigned pb=pini+i;
            const float4* __restrict__ ppb=t.posp+pb;
            if(__all_sync(FULL,pb+V5_CH<=pfin)){      // every lane has a full chunk: no per-candidate range check
              #pragma unroll
              for(int u=0;u<V5_CH;u++){
                const float4 pp2=__ldg(ppb+u);
                const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
                const float rr2=drx*drx+dry*dry+drz*drz;
                if(rr2<=ks2){
                  V5Push(q,tail&((V5_Q-1)*V5_STRIDE),(pb+u)|codebits);
                  tail+=V5_STRIDE;
                }
              }
            }
            else{
              #pragma unroll
              for(int u=0;u<V5_CH;u++){
                const float4 pp2=__ldg(ppb+u);
                const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
                const float rr2=drx*drx+dry*dry+drz*drz;
                if(pb+u<pfin && rr2<=ks2){
                  V5Push(q,tail&((V5_Q-1)*V5_STRIDE),(pb+u)|codebits);
                  tail+=V5_STRIDE;
                }
              }
            }
          }
        }
      }
    }
    // end of the pass: the boundary pass uses other constants, drain everything
    while(__any_sync(FULL,head!=tail))
      V5Round<TKER,TVISCO,TDDT,SHIFT,V5_Q,OPT>(q,head,tail,pp1,fp,taup1,boundp2,t,c,a,anyb,nosh);
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

// ---- the variant-8 skeleton for other pair bodies (multiphase kernels of forces_nn.cuh) ----
// Shared memory of a CTA: V5_Q*FORCES_THREADS ring words + 64 float4 of cell offsets; must be the kernel's only static shared array.
template<int V5_Q> struct V8Smem { unsigned w[V5_Q*FORCES_THREADS+64*4]; };

// fills the offset table, checks the ring's place in the shared window and returns the lane's (pinned) ring offset
template<int V5_Q> __device__ __forceinline__ unsigned V8Setup(unsigned* qsm,float scell){
  if(unsigned(__cvta_generic_to_shared(qsm))!=V5_SMEM0)__trap();
  if(threadIdx.x<64){
    const unsigned k=threadIdx.x;
    reinterpret_cast<float4*>(qsm+V5_Q*FORCES_THREADS)[k]=make_float4((float(int(k&3u))-1.f)*scell,(float(int((k>>2)&3u))-1.f)*scell,(float(int(k>>4))-1.f)*scell,0.f);
  }
  __syncthreads();
  unsigned q=threadIdx.x*4u;
  asm volatile("" : "+r"(q));
  return q;
}

// pair(p2,drx,dry,drz,rr2 clamped,press of p2,boundp2,valid) is called for every accepted neighbour (fluid pass, then boundary pass;
// rows in the order of variant 9, cells and particles ascending inside a row)
template<int V5_Q,class PAIR> __device__ __forceinline__ void V8Traverse(const ForceParams &t,const DevCtes &c,bool active,const float4 &pp1,unsigned rcell,
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
      const unsigned e=V5PopOr<V5_Q>(q,head);
      head+=V5_STRIDE;
      const unsigned p2=e&V5_IDXMASK;
      const float4 off=V5OffTab<V5_Q>(e,0u);
      const float4 pp2=__ldg(&t.posp[p2]);
      float drx,dry;
      Upk(Sub2(Sub2(Pk(pp1.x,pp1.y),Pk(off.x,off.y)),Pk(pp2.x,pp2.y)),drx,dry);
      const float drz=(pp1.z-off.z)-pp2.z;
      const float rr2=drx*drx+dry*dry+drz*drz;
      pair(p2,drx,dry,drz,fmaxf(rr2,1e-18f),pp2.w,boundp2,rr2>=1e-18f);
    }
  };
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const bool boundp2=(pass==1);
    const unsigned cellinit=(boundp2? 0u: t.cellfluid);
    #pragma unroll 1
    for(int r=0;r<9;r++){
      {
        // rows of three cells in the order of variant 9 (opposite rows back to back)
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
          const unsigned n=pfin-pini;
          const unsigned nmax=__reduce_max_sync(FULL,n);
          const unsigned ifull=(__reduce_min_sync(FULL,n)/V5_CH)*V5_CH;
          const unsigned codebits=(unsigned(dx+1)|(unsigned(dy+1)<<2)|(unsigned(dz+1)<<4))<<26;
          unsigned vb=pini|codebits;
          const unsigned vfin=pfin|codebits;
          unsigned i=0;
          for(;i<ifull;i+=V5_CH,vb+=V5_CH){
            while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))round(boundp2);
            V5Chunk<V5_Q,false,true>(t.posp,vb,vfin,ox,oy,oz,ks2,q,tail);
          }
          for(;i<nmax;i+=V5_CH,vb+=V5_CH){
            while(__any_sync(FULL,(tail-head)>unsigned(V5_Q-V5_CH)*V5_STRIDE))round(boundp2);
            V5Chunk<V5_Q,true,true>(t.posp,vb,vfin,ox,oy,oz,ks2,q,tail);
          }
        }
      }
    }
    while(__any_sync(FULL,head!=tail))round(boundp2);
  }
}

// Where does the first static shared array of a kernel start in the CTA's shared window?  Variants 7/8 fold V5_SMEM0 into their
// ld/st.shared offsets; the engine asks once per context and falls back to variant 6 (multiphase: the plain loop) if it differs.
__global__ void ker_v8_probe(unsigned *out){
  __shared__ __align__(16) V8Smem<32> qsm;
  if(threadIdx.x==0)*out=unsigned(__cvta_generic_to_shared(qsm.w));
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> static void LaunchFluidV5(unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm,int qdepth=64,int opt=0){
  const unsigned nb=(npf+FORCES_THREADS-1)/FORCES_THREADS;
  if(c.scelldiv==1 && t.np<=V5_IDXMASK){
    if(qdepth==32 && opt==2)ker_forces_fluid_v5<TKER,TVISCO,TDDT,SHIFT,32,2,6><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    else if(qdepth==32 && opt==3)ker_forces_fluid_v5<TKER,TVISCO,TDDT,SHIFT,32,3,6><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    else if(qdepth==32 && opt)ker_forces_fluid_v5<TKER,TVISCO,TDDT,SHIFT,32,1><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    else if(qdepth==32)ker_forces_fluid_v5<TKER,TVISCO,TDDT,SHIFT,32,0><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    else ker_forces_fluid_v5<TKER,TVISCO,TDDT,SHIFT,64,0><<<nb,FORCES_THREADS,0,stm>>>(t,c);
  }
  else ker_forces_fluid_v4<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c);
}

} // namespace dsph
