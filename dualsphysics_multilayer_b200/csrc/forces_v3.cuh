// forces_v3.cuh -- warp-cooperative fluid Interaction_Forces for sm_100a (the production kernel).
//
// ncu on the plain kernel (v1 = the reference's one-thread-per-particle structure) shows it is issue-bound (81 % issue
// active) while only ~1 candidate in 6 passes the 2h cut-off, i.e. the ~100-instruction pair body runs with ~17 % of the
// lanes active.  A per-lane FIFO (v2) does not fix it: hits arrive in bursts that differ between lanes, so the forced
// partial rounds run with ~11 lanes (profiles/r01_forces_v1_v2.md).  v3 turns the warp by 90 degrees:
//
//   * a warp owns 32 consecutive (cell-sorted) fluid particles and processes them in GROUPS of G<=2 particles of the same
//     cell; the lanes of the warp are the CANDIDATES: a (y,z) row of the neighbour search is one contiguous particle range
//     (3 cells, 5 with cells of size h), read with coalesced 16 B loads (posp + velrhop: 2 x 512 B per 32 candidates),
//   * every lane tests its candidate against the G particles (3 FADD + 3 FFMA each); ballot + popc compact the hits into a
//     per-particle ring buffer in shared memory as {dr.xyz, press2} {vel2.xyz, +-rho2} (sign of rho2 = boundary flag),
//   * whenever a ring holds >=32 hits all 32 lanes pop one entry each and run the pair body together: full warps in the
//     expensive code, no gather loads; fluid-fluid and fluid-boundary pairs share the ring (per-lane mass / viscosity /
//     DDT handling selected by the flag), so only one partial round is left per particle,
//   * per-lane partial sums are combined with warp shuffles (Ace/Ar/delta; max for viscdt) and lane 0 stores one float4.
// Pair terms are computed from exactly the same dr values as v1 (same cell-relative subtraction), only the summation order
// differs (tree instead of sequential), so v3 agrees with v1 to summation rounding (~1e-7, tests/test_gpu_parity.py).
// Supported here: Wendland|cubic x DDT 0-3 x artificial viscosity x shifting (modes none, NoBound, Full).  Laminar+SPS and
// SHIFT_NoFixed need per-neighbour tau/code and fall back to v1.
#pragma once
#include "forces.cuh"

namespace dsph {

constexpr int V3_WARPS=4;                       // warps per block
constexpr int V3_THREADS=V3_WARPS*32;
constexpr int V3_QCAP=64;                       // ring entries per particle of the group (>= 31 left + 32 new)
constexpr int V3_G=2;                           // particles per group
constexpr int V3_MAXROW=6;                      // begincell values per row: 2*scelldiv+2 (scelldiv<=2)

struct V3Acc { float ax,ay,az,ar,delta,visc; float4 shift; unsigned anybound,noshift; };

// pair body with per-lane boundary flag (same arithmetic as PairFluid<...,BOUNDP2>)
template<int TKER,int TDDT,bool SHIFT> __device__ __forceinline__ void V3Pair(const float4 d,const float4 v2in,
  const float4 &velrhop1,float pressp1,const DevCtes &c,V3Acc &a)
{
  const bool boundp2=(v2in.w<0.f);
  float4 velrhop2=v2in; velrhop2.w=fabsf(v2in.w);
  const float massp2=(boundp2? c.massbound: c.massfluid);
  const float visco=(boundp2? c.viscobound: c.visco);
  const float drx=d.x, dry=d.y, drz=d.z, pressp2=d.w;
  const float rr2=drx*drx+dry*dry+drz*drz;
  const float fac=KernelFac<TKER>(rr2,c);
  const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
  const float rhopp1=velrhop1.w;
  {
    const float prs=(pressp1+pressp2)/(rhopp1*velrhop2.w)+(TKER==DSPH_KERNEL_CUBIC? CubicTensil(rr2,rhopp1,pressp1,velrhop2.w,pressp2,c): 0.f);
    const float p_vpm=-prs*massp2;
    a.ax+=p_vpm*frx; a.ay+=p_vpm*fry; a.az+=p_vpm*frz;
  }
  const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
  a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhopp1/velrhop2.w);
  const float cbar=c.cs0;
  if(TDDT==DSPH_DDT_DDT){
    if(boundp2)a.anybound=1u;
    else{
      const float rhop1over2=rhopp1/velrhop2.w;
      const float visc_densi=c.ddtkh*cbar*(rhop1over2-1.f)/(rr2+c.eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      a.delta+=visc_densi*dot3*massp2;
    }
  }
  if(TDDT==DSPH_DDT_DDT2 || TDDT==DSPH_DDT_DDT2FULL){
    if(boundp2){ if(TDDT==DSPH_DDT_DDT2)a.anybound=1u; }
    else{
      const float rh=1.f+c.ddtgz*drz;
      const float drhop=c.rhopzero*PowNearOne(rh,c)-c.rhopzero;
      const float visc_densi=c.ddtkh*cbar*((velrhop2.w-rhopp1)-drhop)/(rr2+c.eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      a.delta-=visc_densi*dot3*massp2/velrhop2.w;
    }
  }
  if(SHIFT){
    const float massrhop=massp2/velrhop2.w;
    if(boundp2 && c.shiftmode==DSPH_SHIFT_NOBOUND)a.noshift=1u;
    a.shift.x+=massrhop*frx; a.shift.y+=massrhop*fry; a.shift.z+=massrhop*frz;
    a.shift.w-=massrhop*(drx*frx+dry*fry+drz*frz);
  }
  {
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot/(rr2+c.eta2);
    a.visc=fmaxf(dot_rr2,a.visc);
    if(dot<0){
      const float amubar=c.kernelh*dot_rr2;
      const float robar=(rhopp1+velrhop2.w)*0.5f;
      const float pi_visc=(-visco*cbar*amubar/robar)*massp2;
      a.ax-=pi_visc*frx; a.ay-=pi_visc*fry; a.az-=pi_visc*frz;
    }
  }
}

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

// Processes one group of G particles of the same cell.  P*: uniform (warp-broadcast) particle data.
template<int TKER,int TDDT,bool SHIFT,int G> __device__ __forceinline__ void V3Group(
  const unsigned (&p1)[G],const float4 (&pp1)[G],const float4 (&vr1)[G],int cx,int cy,int cz,
  float4* __restrict__ q0,float4* __restrict__ q1,int lane,const ForceParams &t,const DevCtes &c,
  float &viscmax,float &ace2max,float &vel2max)
{
  const unsigned FULL=0xffffffffu;
  const unsigned ltmask=(1u<<lane)-1u;
  const int sd=c.scelldiv;
  V3Acc acc[G];
  unsigned qhead[G],qcnt[G];
  #pragma unroll
  for(int g=0;g<G;g++){
    acc[g].ax=acc[g].ay=acc[g].az=acc[g].ar=acc[g].delta=acc[g].visc=0; acc[g].shift=make_float4(0,0,0,0); acc[g].anybound=acc[g].noshift=0;
    qhead[g]=0; qcnt[g]=0;
  }
  // x range of the row (cells xa..xb) and the per-cell x offsets of each particle of the group
  const int xa=max(cx-sd,0), xb=min(cx+sd,t.ncx-1);
  const int ncell=xb-xa+1;
  float oxc[G][V3_MAXROW-1];
  #pragma unroll
  for(int g=0;g<G;g++){
    #pragma unroll
    for(int k=0;k<V3_MAXROW-1;k++)oxc[g][k]=pp1[g].x-float(xa+k-cx)*c.scell;
  }
  #pragma unroll 1
  for(int pass=0;pass<2;pass++){
    const unsigned cellinit=(pass==0? t.cellfluid: 0u);
    const float bsign=(pass==0? 1.f: -1.f);
    #pragma unroll 1
    for(int dz=-sd;dz<=sd;dz++){
      const int z=cz+dz;
      if(z<0 || z>=t.ncz)continue;
      #pragma unroll 1
      for(int dy=-sd;dy<=sd;dy++){
        const int y=cy+dy;
        if(y<0 || y>=t.ncy)continue;
        const unsigned v=unsigned(t.nsheet*z+t.ncx*y+xa)+cellinit;
        unsigned b[V3_MAXROW];
        #pragma unroll
        for(int k=0;k<V3_MAXROW;k++)b[k]=(k<=ncell? t.begincell[v+k]: 0xffffffffu);
        const unsigned rowb=b[0];
        unsigned rowe=rowb;
        #pragma unroll
        for(int k=1;k<V3_MAXROW;k++)if(k<=ncell)rowe=b[k];
        if(rowe==rowb)continue;
        float oy[G],oz[G];
        #pragma unroll
        for(int g=0;g<G;g++){ oy[g]=pp1[g].y-float(dy)*c.scell; oz[g]=pp1[g].z-float(dz)*c.scell; }
        #pragma unroll 1
        for(unsigned pb=rowb;pb<rowe;pb+=32){
          const unsigned p2=pb+lane;
          const bool valid=(p2<rowe);
          const unsigned p2c=(valid? p2: rowb);
          const float4 pp2=__ldg(&t.posp[p2c]);
          float4 vr2=__ldg(&t.velrhop[p2c]);
          vr2.w*=bsign;
          #pragma unroll
          for(int g=0;g<G;g++){
            float ox=oxc[g][0];
            #pragma unroll
            for(int k=1;k<V3_MAXROW-1;k++)if(p2>=b[k])ox=oxc[g][k];
            const float drx=ox-pp2.x, dry=oy[g]-pp2.y, drz=oz[g]-pp2.z;
            const float rr2=drx*drx+dry*dry+drz*drz;
            const bool hit=(valid && rr2<=c.kernelsize2 && rr2>=1e-18f);
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
                const float4 w=q1[g*V3_QCAP+slot];
                V3Pair<TKER,TDDT,SHIFT>(d,w,vr1[g],pp1[g].w,c,acc[g]);
                qhead[g]=(qhead[g]+32)&(V3_QCAP-1); qcnt[g]-=32;
                __syncwarp();
              }
            }
          }
        }
      }
    }
  }
  // left-over entries (<32 per particle), then warp reduction and store
  #pragma unroll
  for(int g=0;g<G;g++){
    __syncwarp();
    if(unsigned(lane)<qcnt[g]){
      const unsigned slot=(qhead[g]+lane)&(V3_QCAP-1);
      const float4 d=q0[g*V3_QCAP+slot];
      const float4 w=q1[g*V3_QCAP+slot];
      V3Pair<TKER,TDDT,SHIFT>(d,w,vr1[g],pp1[g].w,c,acc[g]);
    }
    __syncwarp();
    float ax=WarpSum(acc[g].ax), ay=WarpSum(acc[g].ay), az=WarpSum(acc[g].az), ar=WarpSum(acc[g].ar);
    const float visc=WarpMaxF(acc[g].visc);
    if(TDDT!=DSPH_DDT_NONE){
      const float delta=WarpSum(acc[g].delta);
      const bool anyb=(__any_sync(FULL,acc[g].anybound!=0));
      if(!anyb)ar+=delta;
    }
    if(c.simulate2d)ay=0;
    float4 sh=make_float4(0,0,0,0);
    if(SHIFT){
      sh.x=WarpSum(acc[g].shift.x); sh.y=WarpSum(acc[g].shift.y); sh.z=WarpSum(acc[g].shift.z); sh.w=WarpSum(acc[g].shift.w);
      if(__any_sync(FULL,acc[g].noshift!=0))sh.x=FLT_MAX;
    }
    if(lane==0){
      t.acear[p1[g]]=make_float4(ax,ay,az,ar);
      if(SHIFT)t.shiftposfs[p1[g]]=sh;
      const bool normal=(!c.periactive || (t.code[p1[g]]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
      if(normal)ace2max=fmaxf(ace2max,ax*ax+ay*ay+az*az);
      vel2max=fmaxf(vel2max,vr1[g].x*vr1[g].x+vr1[g].y*vr1[g].y+vr1[g].z*vr1[g].z);
      viscmax=fmaxf(viscmax,visc);
    }
  }
}

template<int TKER,int TDDT,bool SHIFT> __global__ void __launch_bounds__(V3_THREADS,4)
ker_forces_fluid_v3(ForceParams t,DevCtes c)
{
  __shared__ float4 qsm[V3_WARPS*2*V3_G*V3_QCAP];     // 4 warps x 2 arrays x 2 particles x 64 entries x 16 B = 16 KB
  const int lane=threadIdx.x&31, w=threadIdx.x>>5;
  float4 *q0=qsm+w*(2*V3_G*V3_QCAP);
  float4 *q1=q0+V3_G*V3_QCAP;
  const unsigned base=t.npb+(blockIdx.x*V3_WARPS+w)*32;
  if(base>=t.np)return;
  const unsigned nmine=min(32u,t.np-base);
  // the warp's 32 particles, one per lane (coalesced), broadcast with shuffles while iterating
  const unsigned pl=base+min(unsigned(lane),nmine-1);
  const float4 ppl=t.posp[pl];
  const float4 vrl=t.velrhop[pl];
  const unsigned cel=t.dcell[pl];
  float viscmax=0,ace2max=0,vel2max=0;
  unsigned i=0;
  while(i<nmine){
    const unsigned FULL=0xffffffffu;
    const unsigned c0=__shfl_sync(FULL,cel,i);
    const unsigned c1=__shfl_sync(FULL,cel,min(i+1,31u));
    const int cx=int(DcCellx(c.domcellcode,c0))-int(t.cellzero[0]);
    const int cy=int(DcCelly(c.domcellcode,c0))-int(t.cellzero[1]);
    const int cz=int(DcCellz(c.domcellcode,c0))-int(t.cellzero[2]);
    float4 a0,b0,a1,b1;
    a0.x=__shfl_sync(FULL,ppl.x,i); a0.y=__shfl_sync(FULL,ppl.y,i); a0.z=__shfl_sync(FULL,ppl.z,i); a0.w=__shfl_sync(FULL,ppl.w,i);
    b0.x=__shfl_sync(FULL,vrl.x,i); b0.y=__shfl_sync(FULL,vrl.y,i); b0.z=__shfl_sync(FULL,vrl.z,i); b0.w=__shfl_sync(FULL,vrl.w,i);
    if(i+1<nmine && c1==c0){
      const unsigned j=i+1;
      a1.x=__shfl_sync(FULL,ppl.x,j); a1.y=__shfl_sync(FULL,ppl.y,j); a1.z=__shfl_sync(FULL,ppl.z,j); a1.w=__shfl_sync(FULL,ppl.w,j);
      b1.x=__shfl_sync(FULL,vrl.x,j); b1.y=__shfl_sync(FULL,vrl.y,j); b1.z=__shfl_sync(FULL,vrl.z,j); b1.w=__shfl_sync(FULL,vrl.w,j);
      const unsigned pp[2]={base+i,base+j}; const float4 xp[2]={a0,a1}; const float4 xv[2]={b0,b1};
      V3Group<TKER,TDDT,SHIFT,2>(pp,xp,xv,cx,cy,cz,q0,q1,lane,t,c,viscmax,ace2max,vel2max);
      i+=2;
    }
    else{
      const unsigned pp[1]={base+i}; const float4 xp[1]={a0}; const float4 xv[1]={b0};
      V3Group<TKER,TDDT,SHIFT,1>(pp,xp,xv,cx,cy,cz,q0,q1,lane,t,c,viscmax,ace2max,vel2max);
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
  if(TVISCO==DSPH_VISCO_ARTIFICIAL && !(SHIFT && c.shiftmode==DSPH_SHIFT_NOFIXED) && c.scelldiv<=2){
    const unsigned nb=(npf+V3_THREADS-1)/V3_THREADS;
    ker_forces_fluid_v3<TKER,TDDT,SHIFT><<<nb,V3_THREADS,0,stm>>>(t,c);
  }
  else{
    const unsigned nb=(npf+FORCES_THREADS-1)/FORCES_THREADS;
    ker_forces_fluid_v1<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c);
  }
}

} // namespace dsph
