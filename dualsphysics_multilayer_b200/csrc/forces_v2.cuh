// forces_v2.cuh -- tuned fluid Interaction_Forces for sm_100a: "test densely, compute densely".
//
// Problem with the plain loop (forces.cuh v1 = the reference's structure, S/JSphGpu_ker.cu:512-745): only ~1 candidate in
// 6 passes the 2h cut-off (162 accepted of 972 checked per fluid particle), so the ~80-instruction pair body
// runs with ~17 % of the lanes of a warp active.  v2 splits the work into two warp-uniform phases:
//   TEST   every lane walks its 27 (125 with cells of size h) neighbour cells; the candidate loads are warp-broadcast
//          (lanes of one cell read the same address); a hit is appended to the lane's own FIFO in shared memory as
//          {dr.xyz, press2} + {vel2.xyz, rho2} (two conflict-free 16 B stores, slot-major layout q[slot][lane]).
//   PAIR   when every lane has a queued hit (or some lane's FIFO is nearly full) ALL lanes pop one entry and run the pair
//          body together -> the expensive code runs with (nearly) full warps and needs no gather loads at all.
// The FIFO keeps the oracle's pair order per particle, so v2 is bit-identical to v1 (tests/test_gpu_parity.py).
// Control flow is uniform per warp (fixed 3x3x3 / 5x5x5 cell loops, trip counts reduced with redux.sync), which is what
// allows the warp votes.  Maxima are reduced with redux.sync + one atomicMax per warp.
#pragma once
#include "forces.cuh"

namespace dsph {

constexpr int V2_Q=8;                    // FIFO slots per lane
constexpr int V2_CHUNK=4;                // candidates tested between two queue checks
constexpr int V2_SMEM=2*V2_Q*FORCES_THREADS*16;   // 32 KB per block

template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void V2Round(
  float4* __restrict__ q0,float4* __restrict__ q1,unsigned &head,unsigned tail,const float4 &velrhop1,float pressp1,
  const float* taup1,float massp2,float visco,const DevCtes &c,PairAcc &a)
{
  if(head!=tail){
    const unsigned slot=(head&(V2_Q-1))*FORCES_THREADS;
    const float4 d=q0[slot];
    const float4 v2=q1[slot];
    head++;
    const float rr2=d.x*d.x+d.y*d.y+d.z*d.z;
    PairFluid<TKER,TVISCO,TDDT,SHIFT,BOUNDP2>(d.x,d.y,d.z,rr2,velrhop1,pressp1,v2,d.w,massp2,visco,taup1,nullptr,0u,c,a);
  }
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void V2Pass(
  bool active,int cx,int cy,int cz,const float4 &pp1,const float4 &velrhop1,const float* taup1,float massp2,float visco,
  float4* __restrict__ q0,float4* __restrict__ q1,const ForceParams &t,const DevCtes &c,PairAcc &a)
{
  const unsigned FULL=0xffffffffu;
  const unsigned cellinit=(BOUNDP2? 0u: t.cellfluid);
  const int sd=c.scelldiv;
  unsigned head=0,tail=0;
  for(int dz=-sd;dz<=sd;dz++){
    const int z=cz+dz;
    const float oz=pp1.z-float(dz)*c.scell;
    for(int dy=-sd;dy<=sd;dy++){
      const int y=cy+dy;
      const float oy=pp1.y-float(dy)*c.scell;
      const bool rowok=(active && z>=0 && z<t.ncz && y>=0 && y<t.ncy);
      for(int dx=-sd;dx<=sd;dx++){
        const int x=cx+dx;
        const float ox=pp1.x-float(dx)*c.scell;
        unsigned pini=0,pfin=0;
        if(rowok && x>=0 && x<t.ncx){
          const unsigned v=unsigned(t.nsheet*z+t.ncx*y+x)+cellinit;
          pini=t.begincell[v]; pfin=t.begincell[v+1];
        }
        const unsigned nmax=__reduce_max_sync(FULL,pfin-pini);
        for(unsigned i=0;i<nmax;i+=V2_CHUNK){
          // make room: every lane needs V2_CHUNK free slots before the next chunk
          while(__any_sync(FULL,(tail-head)>unsigned(V2_Q-V2_CHUNK)))
            V2Round<TKER,TVISCO,TDDT,SHIFT,BOUNDP2>(q0,q1,head,tail,velrhop1,pp1.w,taup1,massp2,visco,c,a);
          #pragma unroll
          for(int u=0;u<V2_CHUNK;u++){
            const unsigned p2=pini+i+u;
            const bool ok=(p2<pfin);
            const unsigned p2c=(ok? p2: 0u);
            const float4 pp2=__ldg(&t.posp[p2c]);
            const float4 vr2=__ldg(&t.velrhop[p2c]);
            const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
            const float rr2=drx*drx+dry*dry+drz*drz;
            if(ok && rr2<=c.kernelsize2 && rr2>=1e-18f){
              const unsigned slot=(tail&(V2_Q-1))*FORCES_THREADS;
              q0[slot]=make_float4(drx,dry,drz,pp2.w);
              q1[slot]=vr2;
              tail++;
            }
          }
          // dense rounds: run the pair body while every lane has work
          while(__all_sync(FULL,head!=tail))
            V2Round<TKER,TVISCO,TDDT,SHIFT,BOUNDP2>(q0,q1,head,tail,velrhop1,pp1.w,taup1,massp2,visco,c,a);
        }
      }
    }
  }
  while(__any_sync(FULL,head!=tail))
    V2Round<TKER,TVISCO,TDDT,SHIFT,BOUNDP2>(q0,q1,head,tail,velrhop1,pp1.w,taup1,massp2,visco,c,a);
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS,4)
ker_forces_fluid_v2(ForceParams t,DevCtes c)
{
  __shared__ float4 qsm[2*V2_Q*FORCES_THREADS];
  float4 *q0=qsm+threadIdx.x;
  float4 *q1=qsm+V2_Q*FORCES_THREADS+threadIdx.x;
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  const bool active=(p1<t.np);
  const unsigned p1c=(active? p1: t.np-1);
  const float4 pp1=t.posp[p1c];
  const float4 velrhop1=t.velrhop[p1c];
  const unsigned rcell=t.dcell[p1c];
  const int cx=int(DcCellx(c.domcellcode,rcell))-int(t.cellzero[0]);
  const int cy=int(DcCelly(c.domcellcode,rcell))-int(t.cellzero[1]);
  const int cz=int(DcCellz(c.domcellcode,rcell))-int(t.cellzero[2]);
  float taup1[6]={0,0,0,0,0,0};
  PairAcc ff; ZeroAcc(ff);
  V2Pass<TKER,TVISCO,TDDT,SHIFT,false>(active,cx,cy,cz,pp1,velrhop1,taup1,c.massfluid,c.visco,q0,q1,t,c,ff);
  PairAcc fb; ZeroAcc(fb);
  if(SHIFT)fb.shift=ff.shift;
  V2Pass<TKER,TVISCO,TDDT,SHIFT,true>(active,cx,cy,cz,pp1,velrhop1,taup1,c.massbound,c.viscobound,q0,q1,t,c,fb);
  float viscmax=0,ace2=0,vel2=0;
  if(active){
    float ar=0,ax=0,ay=0,az=0,delta=0;
    if(SHIFT||ff.ar||ff.ax||ff.ay||ff.az||ff.visc){
      if(TDDT!=DSPH_DDT_NONE)delta=(ff.delta==FLT_MAX? FLT_MAX: 0.f+ff.delta);
      ar+=ff.ar; ax+=ff.ax; ay+=ff.ay; az+=ff.az; viscmax=fmaxf(viscmax,ff.visc);
    }
    if(SHIFT||fb.ar||fb.ax||fb.ay||fb.az||fb.visc){
      if(TDDT!=DSPH_DDT_NONE)delta=(delta==FLT_MAX || fb.delta==FLT_MAX? FLT_MAX: delta+fb.delta);
      ar+=fb.ar; ax+=fb.ax; ay+=fb.ay; az+=fb.az; viscmax=fmaxf(viscmax,fb.visc);
    }
    if(c.simulate2d)ay=0;
    if(TDDT!=DSPH_DDT_NONE && delta!=FLT_MAX)ar+=delta;
    t.acear[p1]=make_float4(ax,ay,az,ar);
    if(SHIFT)t.shiftposfs[p1]=fb.shift;
    const bool normal=(!c.ghosts || (t.code[p1]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
    if(normal)ace2=ax*ax+ay*ay+az*az;
    vel2=velrhop1.x*velrhop1.x+velrhop1.y*velrhop1.y+velrhop1.z*velrhop1.z;
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
}

// v2 covers artificial viscosity without the NoFixed shifting mode (the FIFO does not carry tau / code of p2);
// the other template combinations use the plain kernel.
template<int TKER,int TVISCO,int TDDT,bool SHIFT> static void LaunchFluidV2(unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  const unsigned nb=(npf+FORCES_THREADS-1)/FORCES_THREADS;
  if(TVISCO==DSPH_VISCO_ARTIFICIAL && !(SHIFT && c.shiftmode==DSPH_SHIFT_NOFIXED))
    ker_forces_fluid_v2<TKER,DSPH_VISCO_ARTIFICIAL,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c);
  else
    ker_forces_fluid_v1<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c);
}

} // namespace dsph
