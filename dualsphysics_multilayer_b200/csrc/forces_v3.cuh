// forces_v3.cuh -- warp-cooperative fluid Interaction_Forces for sm_100a (variant 3).
//
// The warp is turned by 90 degrees with respect to the reference structure:
//   * a warp owns 32 consecutive (cell-sorted) fluid particles and processes them in GROUPS of G<=2 particles of the same
//     cell; the LANES are the CANDIDATES: one (y,z) row of the neighbour search is a contiguous particle range (3 cells,
//     5 with cells of size h) read with coalesced 16 B loads (posp + velrhop),
//   * every lane tests its candidate against the G particles; ballot + popc compact the hits into a per-particle ring
//     buffer in shared memory as {dr.xyz, press2} {vel2.xyz, +-rho2} (sign of rho2 = boundary flag),
//   * whenever a ring holds >=32 hits all 32 lanes pop one entry each and run the pair body together: full warps in the
//     expensive code and no gather loads; fluid-fluid and fluid-boundary pairs share the ring,
//   * per-lane partial sums are combined with warp shuffles (Ace/Ar/delta; max for viscdt) and lane 0 stores one float4.
// Pair terms use the division-free body of forces_fast.cuh; the summation is a tree instead of the oracle's sequential
// order.  Laminar+SPS and SHIFT_NoFixed need per-neighbour tau/code and use variant 4.
#pragma once
#include "forces_fast.cuh"

namespace dsph {

constexpr int V3_WARPS=4;
constexpr int V3_THREADS=V3_WARPS*32;
constexpr int V3_QCAP=64;                       // ring entries per particle of the group (<=31 left + <=32 new)
constexpr int V3_G=2;

__device__ __forceinline__ float WarpSum(float v){
  #pragma unroll
  for(int o=16;o;o>>=1)v+=__shfl_xor_sync(0xffffffffu,v,o);
  return v;
}
__device__ __forceinline__ float WarpMaxF(float v){
  #pragma unroll
  for(int o=16;o;o>>=1)v=fmaxf(v,__shfl_xor_sync(0xffffffffu,v,o));
  return v;
}

template<int TKER,int TDDT,bool SHIFT,int G,int SD> __device__ __forceinline__ void V3Group(
  const unsigned (&p1)[G],const float4 (&pp1)[G],const float4 (&vr1)[G],int cx,int cy,int cz,
  float4* __restrict__ q0,float4* __restrict__ q1,int lane,const ForceParams &t,const DevCtes &c,
  float &viscmax,float &ace2max,float &vel2max)
{
  constexpr int NC=2*SD+1;                      // cells per row
  const unsigned FULL=0xffffffffu;
  const unsigned ltmask=(1u<<lane)-1u;
  PairAcc acc[G]; FastP1 fp[G];
  unsigned anyb[G],nosh[G],qhead[G],qcnt[G];
  #pragma unroll
  for(int g=0;g<G;g++){
    ZeroAcc(acc[g]); anyb[g]=nosh[g]=0; qhead[g]=qcnt[g]=0;
    fp[g].velrhop=vr1[g]; fp[g].press=pp1[g].w; fp[g].ovrhop=1.f/vr1[g].w;
  }
  const float scell=c.scell, ks2=c.kernelsize2;
  // x offsets of the (up to NC) cells of a row, applied on the candidate side
  const int xa=max(cx-SD,0), xb=min(cx+SD,t.ncx-1);
  const int ncell=xb-xa+1;
  float xoff[NC];
  #pragma unroll
  for(int k=0;k<NC;k++)xoff[k]=float(xa+k-cx)*scell;
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const unsigned cellinit=(pass==0? t.cellfluid: 0u);
    const float bsign=(pass==0? 1.f: -1.f);
    #pragma unroll 1
    for(int dz=-SD;dz<=SD;dz++){
      const int z=cz+dz;
      if(z<0 || z>=t.ncz)continue;
      #pragma unroll 1
      for(int dy=-SD;dy<=SD;dy++){
        const int y=cy+dy;
        if(y<0 || y>=t.ncy)continue;
        const unsigned* __restrict__ bc=t.begincell+(unsigned(t.nsheet*z+t.ncx*y+xa)+cellinit);
        unsigned b[NC+1];
        #pragma unroll
        for(int k=0;k<=NC;k++)b[k]=bc[min(k,ncell)];
        const unsigned rowb=b[0], rowe=b[NC];
        if(rowe==rowb)continue;
        float oy[G],oz[G];
        #pragma unroll
        for(int g=0;g<G;g++){ oy[g]=pp1[g].y-float(dy)*scell; oz[g]=pp1[g].z-float(dz)*scell; }
        #pragma unroll 1
        for(unsigned pb=rowb;pb<rowe;pb+=32){
          const unsigned p2=pb+lane;
          const bool valid=(p2<rowe);
          const unsigned p2c=(valid? p2: rowb);
          const float4 pp2=__ldg(&t.posp[p2c]);
          float4 vr2=__ldg(&t.velrhop[p2c]);
          vr2.w*=bsign;
          float xs=xoff[0];
          #pragma unroll
          for(int k=1;k<NC;k++)if(p2>=b[k])xs=xoff[k];
          const float x2=pp2.x+xs;
          #pragma unroll
          for(int g=0;g<G;g++){
            const float drx=pp1[g].x-x2, dry=oy[g]-pp2.y, drz=oz[g]-pp2.z;
            const float rr2=drx*drx+dry*dry+drz*drz;
            const bool hit=(valid && rr2<=ks2 && rr2>=1e-18f);
            const unsigned bal=__ballot_sync(FULL,hit);
            if(bal){
              if(hit){
                const unsigned slot=(qhead[g]+qcnt[g]+__popc(bal&ltmask))&(V3_QCAP-1);
                q0[g*V3_QCAP+slot]=make_float4(drx,dry,drz,pp2.w);
                q1[g*V3_QCAP+slot]=vr2;
              }
              qcnt[g]+=__popc(bal);
              if(qcnt[g]>=32){
                __syncwarp();
                const unsigned slot=(qhead[g]+lane)&(V3_QCAP-1);
                const float4 d=q0[g*V3_QCAP+slot];
                float4 w=q1[g*V3_QCAP+slot];
                const bool bnd=(w.w<0.f); w.w=fabsf(w.w);
                PairFast<TKER,DSPH_VISCO_ARTIFICIAL,TDDT,SHIFT>(d.x,d.y,d.z,d.x*d.x+d.y*d.y+d.z*d.z,fp[g],w,d.w,bnd,nullptr,nullptr,0u,c,acc[g],anyb[g],nosh[g]);
                qhead[g]=(qhead[g]+32)&(V3_QCAP-1); qcnt[g]-=32;
                __syncwarp();
              }
            }
          }
        }
      }
    }
  }
  #pragma unroll
  for(int g=0;g<G;g++){
    __syncwarp();
    if(unsigned(lane)<qcnt[g]){
      const unsigned slot=(qhead[g]+lane)&(V3_QCAP-1);
      const float4 d=q0[g*V3_QCAP+slot];
      float4 w=q1[g*V3_QCAP+slot];
      const bool bnd=(w.w<0.f); w.w=fabsf(w.w);
      PairFast<TKER,DSPH_VISCO_ARTIFICIAL,TDDT,SHIFT>(d.x,d.y,d.z,d.x*d.x+d.y*d.y+d.z*d.z,fp[g],w,d.w,bnd,nullptr,nullptr,0u,c,acc[g],anyb[g],nosh[g]);
    }
    __syncwarp();
    float ax=WarpSum(acc[g].ax), ay=WarpSum(acc[g].ay), az=WarpSum(acc[g].az), ar=WarpSum(acc[g].ar);
    const float visc=WarpMaxF(acc[g].visc);
    if(TDDT!=DSPH_DDT_NONE){
      const float delta=WarpSum(acc[g].delta);
      if(!__any_sync(FULL,anyb[g]!=0))ar+=delta;
    }
    if(c.simulate2d)ay=0;
    float4 sh=make_float4(0,0,0,0);
    if(SHIFT){
      sh.x=WarpSum(acc[g].shift.x); sh.y=WarpSum(acc[g].shift.y); sh.z=WarpSum(acc[g].shift.z); sh.w=WarpSum(acc[g].shift.w);
      if(__any_sync(FULL,nosh[g]!=0))sh.x=FLT_MAX;
    }
    if(lane==0){
      t.acear[p1[g]]=make_float4(ax,ay,az,ar);
      if(SHIFT)t.shiftposfs[p1[g]]=sh;
      const bool normal=(!c.ghosts || (t.code[p1[g]]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
      if(normal)ace2max=fmaxf(ace2max,ax*ax+ay*ay+az*az);
      vel2max=fmaxf(vel2max,vr1[g].x*vr1[g].x+vr1[g].y*vr1[g].y+vr1[g].z*vr1[g].z);
      viscmax=fmaxf(viscmax,visc);
    }
  }
}

template<int TKER,int TDDT,bool SHIFT,int SD> __global__ void __launch_bounds__(V3_THREADS,4)
ker_forces_fluid_v3(ForceParams t,DevCtes c)
{
  __shared__ float4 qsm[V3_WARPS*2*V3_G*V3_QCAP];     // 4 warps x 2 arrays x 2 particles x 64 entries x 16 B = 16 KB
  const unsigned FULL=0xffffffffu;
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  float4 *q0=qsm+w*(2*V3_G*V3_QCAP);
  float4 *q1=q0+V3_G*V3_QCAP;
  const unsigned base=t.npb+(blockIdx.x*V3_WARPS+w)*32;
  if(base>=t.np)return;
  const unsigned nmine=min(32u,t.np-base);
  const unsigned pl=base+min(unsigned(lane),nmine-1);
  const float4 ppl=t.posp[pl];
  const float4 vrl=t.velrhop[pl];
  const unsigned cel=t.dcell[pl];
  float viscmax=0,ace2max=0,vel2max=0;
  unsigned i=0;
  while(i<nmine){
    const unsigned c0=__shfl_sync(FULL,cel,i);
    const unsigned c1=__shfl_sync(FULL,cel,min(i+1,31u));
    const int cx=int(DcCellx(c.domcellcode,c0))-int(t.cellzero[0]);
    const int cy=int(DcCelly(c.domcellcode,c0))-int(t.cellzero[1]);
    const int cz=int(DcCellz(c.domcellcode,c0))-int(t.cellzero[2]);
    float4 a0,b0;
    a0.x=__shfl_sync(FULL,ppl.x,i); a0.y=__shfl_sync(FULL,ppl.y,i); a0.z=__shfl_sync(FULL,ppl.z,i); a0.w=__shfl_sync(FULL,ppl.w,i);
    b0.x=__shfl_sync(FULL,vrl.x,i); b0.y=__shfl_sync(FULL,vrl.y,i); b0.z=__shfl_sync(FULL,vrl.z,i); b0.w=__shfl_sync(FULL,vrl.w,i);
    if(i+1<nmine && c1==c0){
      const unsigned j=i+1;
      float4 a1,b1;
      a1.x=__shfl_sync(FULL,ppl.x,j); a1.y=__shfl_sync(FULL,ppl.y,j); a1.z=__shfl_sync(FULL,ppl.z,j); a1.w=__shfl_sync(FULL,ppl.w,j);
      b1.x=__shfl_sync(FULL,vrl.x,j); b1.y=__shfl_sync(FULL,vrl.y,j); b1.z=__shfl_sync(FULL,vrl.z,j); b1.w=__shfl_sync(FULL,vrl.w,j);
      const unsigned pp[2]={base+i,base+j}; const float4 xp[2]={a0,a1}; const float4 xv[2]={b0,b1};
      V3Group<TKER,TDDT,SHIFT,2,SD>(pp,xp,xv,cx,cy,cz,q0,q1,lane,t,c,viscmax,ace2max,vel2max);
      i+=2;
    }
    else{
      const unsigned pp[1]={base+i}; const float4 xp[1]={a0}; const float4 xv[1]={b0};
      V3Group<TKER,TDDT,SHIFT,1,SD>(pp,xp,xv,cx,cy,cz,q0,q1,lane,t,c,viscmax,ace2max,vel2max);
      i+=1;
    }
  }
  if(lane==0){
    if(viscmax>0)atomicMax(&t.fmax->viscdt_bits,__float_as_uint(viscmax));
    if(ace2max>0)atomicMax(&t.fmax->ace2_bits,__float_as_uint(ace2max));
    if(vel2max>0)atomicMax(&t.fmax->vel2_bits,__float_as_uint(vel2max));
  }
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> static void LaunchFluidV3(unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  if(TVISCO==DSPH_VISCO_ARTIFICIAL && !(SHIFT && c.shiftmode==DSPH_SHIFT_NOFIXED)){
    const unsigned nb=(npf+V3_THREADS-1)/V3_THREADS;
    if(c.scelldiv==1)ker_forces_fluid_v3<TKER,TDDT,SHIFT,1><<<nb,V3_THREADS,0,stm>>>(t,c);
    else             ker_forces_fluid_v3<TKER,TDDT,SHIFT,2><<<nb,V3_THREADS,0,stm>>>(t,c);
  }
  else LaunchFluidV4<TKER,TVISCO,TDDT,SHIFT>(npf,t,c,stm);
}

} // namespace dsph
