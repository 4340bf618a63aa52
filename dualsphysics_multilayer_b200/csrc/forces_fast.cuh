// forces_fast.cuh -- division-free pair body + the production one-thread-per-particle kernel (variant 4).
//
// ncu on the plain kernel (profiles/r01_forces.md): issue-bound (81 % issue active), and two thirds of the issued
// instructions are the eight IEEE divisions + sqrt of the pair body (each a MUFU + Newton + FCHK slow-path sequence,
// ~17 instructions).  The reference's own CUDA build uses -use_fast_math (S/Makefile:159); here the algebra is
// rearranged so that one pair needs exactly 1 MUFU.RSQ + 3 MUFU.RCP and no division sequence at all:
//   fac      = bwen*qq*w^3/rad with qq=rad/h      ->  (bwen/h)*w^3                 (exact algebra, no division)
//   rad      = sqrt(rr2)                           ->  rr2*rsqrt(rr2)
//   x/rho2 (3 uses), x/(rr2+eta2) (2 uses), x/robar ->  one reciprocal each, 1/rho1 hoisted per particle
// Relative error of MUFU.RCP/RSQ is ~1 ulp; the terms are independent per pair, so the effect on Ace/Ar is far below the
// 1e-5 tolerance (measured in tests/test_gpu_parity.py against the reference dumps).
#pragma once
#include "forces.cuh"

namespace dsph {

__device__ __forceinline__ float FastRcp(float x){ float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float FastRsqrt(float x){ float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

struct FastP1 { float4 velrhop; float press, ovrhop; };

// Store condition of the reference (S/JSphCpu.cpp:800-818, ker_forces_fluid_v1 above): what the fluid-fluid pass accumulated -- the DDT term
// and the SPS velocity gradient with it -- is kept only when one of ar, ace, viscdt is non-zero (always with shifting).  The variants with
// ONE accumulator for both passes apply it when the fluid-fluid pass is complete.  It only matters for a fluid at rest with a uniform
// pressure field, where the reference drops the DDT term.
template<int TVISCO,bool SHIFT> __device__ __forceinline__ void FluidPassStore(PairAcc &a){
  if(!SHIFT && a.ar==0 && a.ax==0 && a.ay==0 && a.az==0 && a.visc==0){
    a.delta=0;
    if(TVISCO==DSPH_VISCO_LAMINARSPS){ a.gxx=a.gxy=a.gxz=a.gyy=a.gyz=a.gzz=0; }
  }
}

// packed fp32x2 arithmetic of sm_100 (FADD2/FMUL2/FFMA2: two IEEE fp32 operations per issue slot, same rounding as the scalar forms)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 Pk(float a,float b){ f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void Upk(f32x2 v,float &a,float &b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 Sub2(f32x2 a,f32x2 b){ f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 Mul2(f32x2 a,f32x2 b){ f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 Fma2(f32x2 a,f32x2 b,f32x2 c){ f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// Accumulates one accepted pair.  boundp2 may be a per-lane value (v3) or a literal (v4: constant-folded).
// NOBRANCH (variant 8): the artificial-viscosity term of receding pairs is selected to zero instead of branched around (the
// branch diverges in nearly every round: BSSY + BRA + BSYNC cost more issue slots than the select).
template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool NOBRANCH=false> __device__ __forceinline__ void PairFast(
  float drx,float dry,float drz,float rr2,const FastP1 &p1,const float4 &velrhop2,float pressp2,bool boundp2,
  const float* __restrict__ taup1,const float* __restrict__ taup2,unsigned codep2,const DevCtes &c,PairAcc &a,unsigned &anybound,unsigned &noshiftf,
  bool valid=true)
{
  // valid=false (v5: the particle itself or a coincident one, rr2 clamped by the caller): the pair runs with zero mass and sets
  // no flag, so that every term is exactly 0 and no branch separates the two neighbour loads of the caller
  const float massp2=(valid? (boundp2? c.massbound: c.massfluid): 0.f);
  const bool boundflag=(boundp2 && valid);
  const float visco=(boundp2? c.viscobound: c.visco);
  // (a Newton step on the MUFU.RSQ seed was measured: +2 % time, no change of the Ace/Ar difference to the IEEE kernel,
  //  which is set by the fp32 conditioning of the pressure sum, not by the approximations)
  const float rinv=FastRsqrt(rr2);
  const float rad=rr2*rinv;
  const float qq=rad*c.ovkernelh;
  float fac;
  if(TKER==DSPH_KERNEL_WENDLAND){
    const float wqq1=1.f-0.5f*qq;
    fac=c.bwen_ovh*(wqq1*wqq1*wqq1);
  }
  else{
    if(rad>c.kernelh){ const float wqq1=2.0f-qq; fac=c.cub_c2*(wqq1*wqq1)*rinv; }
    else fac=c.cub_c1_ovh+c.cub_d1_ovh*qq;
  }
  float frx,fry; const float frz=fac*drz;
  if(NOBRANCH)Upk(Mul2(Pk(fac,fac),Pk(drx,dry)),frx,fry); else{ frx=fac*drx; fry=fac*dry; }
  const float rhopp1=p1.velrhop.w;
  const float ov2=FastRcp(velrhop2.w);
  {
    float prs=(p1.press+pressp2)*(p1.ovrhop*ov2);
    if(TKER==DSPH_KERNEL_CUBIC){
      float wab;
      if(rad>c.kernelh){ const float wqq1=2.0f-qq; wab=c.cub_a24*(wqq1*wqq1*wqq1); }
      else wab=c.cub_a2*(1.0f+(0.75f*qq-1.5f)*(qq*qq));
      float fab=wab*c.cub_od_wdeltap; fab*=fab; fab*=fab;
      const float t1=(p1.press*p1.ovrhop*p1.ovrhop)*(p1.press>0? 0.01f: -0.2f);
      const float t2=(pressp2*ov2*ov2)*(pressp2>0? 0.01f: -0.2f);
      prs+=fab*(t1+t2);
    }
    const float p_vpm=-prs*massp2;
    if(NOBRANCH)Upk(Fma2(Pk(p_vpm,p_vpm),Pk(frx,fry),Pk(a.ax,a.ay)),a.ax,a.ay); else{ a.ax+=p_vpm*frx; a.ay+=p_vpm*fry; }
    a.az+=p_vpm*frz;
  }
  float dvx,dvy; const float dvz=p1.velrhop.z-velrhop2.z;
  if(NOBRANCH)Upk(Sub2(Pk(p1.velrhop.x,p1.velrhop.y),Pk(velrhop2.x,velrhop2.y)),dvx,dvy); else{ dvx=p1.velrhop.x-velrhop2.x; dvy=p1.velrhop.y-velrhop2.y; }
  a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhopp1*ov2);
  const float ovrr=FastRcp(rr2+c.eta2);
  const float cbar=c.cs0;
  if(TDDT==DSPH_DDT_DDT){
    if(boundflag && !c.mdbc)anybound=1u;
    if(!boundp2 || c.mdbc){
      // rho1/rho2-1 cancels to ~1e-4: the oracle's value is dominated by the rounding of the IEEE quotient, so this one
      // division is kept correctly rounded (an approximate reciprocal here changed Ar by 6e-5 in the periodic box).
      // (cubic + laminar/SPS: the reference binary multiplies by the rounded reciprocal instead, see forces.cuh)
      // (variant 8: the quotient is rebuilt from the approximate reciprocal with one exact residual, q0=rho1*ov2, q=q0+(rho1-q0*rho2)*ov2:
      //  the correctly rounded quotient except for rare half-ulp cases, 3 instructions instead of the ~10 of the IEEE division sequence)
      float rhop1over2;
      if(TKER==DSPH_KERNEL_CUBIC && TVISCO==DSPH_VISCO_LAMINARSPS)rhop1over2=__fmul_rn(rhopp1,__frcp_rn(velrhop2.w));
      else if(NOBRANCH){ const float q0=rhopp1*ov2; rhop1over2=__fmaf_rn(__fmaf_rn(-q0,velrhop2.w,rhopp1),ov2,q0); }
      else rhop1over2=__fdiv_rn(rhopp1,velrhop2.w);
      const float visc_densi=c.ddtkh*cbar*(rhop1over2-1.f)*ovrr;
      const float dot3=(drx*frx+dry*fry+drz*frz);
      a.delta+=visc_densi*dot3*massp2;
    }
  }
  if(TDDT==DSPH_DDT_DDT2 || TDDT==DSPH_DDT_DDT2FULL){
    if(boundflag && TDDT==DSPH_DDT_DDT2)anybound=1u;
    if(!boundp2){
      const float rh=1.f+c.ddtgz*drz;
      const float drhop=c.rhopzero*(PowNearOne(rh,c)-1.f);   // rho0*(pow-1): the form the reference binary evaluates
      const float visc_densi=c.ddtkh*cbar*((velrhop2.w-rhopp1)-drhop)*ovrr;
      const float dot3=(drx*frx+dry*fry+drz*frz);
      a.delta-=visc_densi*dot3*massp2*ov2;
    }
  }
  if(SHIFT){
    const float massrhop=massp2*ov2;
    if(boundflag && (c.shiftmode==DSPH_SHIFT_NOBOUND || (c.shiftmode==DSPH_SHIFT_NOFIXED && (codep2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FIXED)))noshiftf=1u;
    a.shift.x+=massrhop*frx; a.shift.y+=massrhop*fry; a.shift.z+=massrhop*frz;
    a.shift.w-=massrhop*(drx*frx+dry*fry+drz*frz);
  }
  {
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot*ovrr;
    a.visc=fmaxf((valid? dot_rr2: 0.f),a.visc);      // rejected pair (variant 10: false positive of the half-precision culling): no effect
    if(TVISCO==DSPH_VISCO_ARTIFICIAL && NOBRANCH){
      const float amubar=c.kernelh*dot_rr2;
      const float ovrobar=FastRcp((rhopp1+velrhop2.w)*0.5f);
      float pi_visc=(-visco*cbar*amubar*ovrobar)*massp2;
      pi_visc=(dot<0? -pi_visc: 0.f);                // x + 0*fr == x: identical sums
      Upk(Fma2(Pk(pi_visc,pi_visc),Pk(frx,fry),Pk(a.ax,a.ay)),a.ax,a.ay);
      a.az+=pi_visc*frz;
    }
    else if(TVISCO==DSPH_VISCO_ARTIFICIAL){
      if(dot<0){
        const float amubar=c.kernelh*dot_rr2;
        const float ovrobar=FastRcp((rhopp1+velrhop2.w)*0.5f);
        const float pi_visc=(-visco*cbar*amubar*ovrobar)*massp2;
        a.ax-=pi_visc*frx; a.ay-=pi_visc*fry; a.az-=pi_visc*frz;
      }
    }
    else{
      {
        const float temp=4.f*visco*ovrr*FastRcp(rhopp1+velrhop2.w);
        const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
        a.ax+=vtemp*dvx; a.ay+=vtemp*dvy; a.az+=vtemp*dvz;
      }
      float tau_xx=taup1[0],tau_xy=taup1[1],tau_xz=taup1[2],tau_yy=taup1[3],tau_yz=taup1[4],tau_zz=taup1[5];
      if(!boundp2){ tau_xx+=taup2[0]; tau_xy+=taup2[1]; tau_xz+=taup2[2]; tau_yy+=taup2[3]; tau_yz+=taup2[4]; tau_zz+=taup2[5]; }
      a.ax+=massp2*(tau_xx*frx+tau_xy*fry+tau_xz*frz);
      a.ay+=massp2*(tau_xy*frx+tau_yy*fry+tau_yz*frz);
      a.az+=massp2*(tau_xz*frx+tau_yz*fry+tau_zz*frz);
      const float volp2=-massp2*ov2;
      float dv=dvx*volp2; a.gxx+=dv*frx; a.gxy+=dv*fry; a.gxz+=dv*frz;
            dv=dvy*volp2; a.gxy+=dv*frx; a.gyy+=dv*fry; a.gyz+=dv*frz;
            dv=dvz*volp2; a.gxz+=dv*frx; a.gyz+=dv*fry; a.gzz+=dv*frz;
    }
  }
}

// ---- variant 4: the plain loop structure (oracle pair order) with the division-free body ----
template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS)
ker_forces_fluid_v4(ForceParams t,DevCtes c)
{
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  float viscmax=0,ace2=0,vel2=0;
  if(p1<t.np){
    const float4 pp1=t.posp[p1];
    FastP1 q; q.velrhop=t.velrhop[p1]; q.press=pp1.w; q.ovrhop=1.f/q.velrhop.w;
    float taup1[6]={0,0,0,0,0,0};
    if(TVISCO==DSPH_VISCO_LAMINARSPS){
      #pragma unroll
      for(int k=0;k<6;k++)taup1[k]=t.spstau[6*size_t(p1)+k];
    }
    const NgRange ng=NgInit(t.dcell[p1],t,c);
    PairAcc a; ZeroAcc(a);
    unsigned anybound=0,noshiftf=0;
    #pragma unroll 1
    for(int pass=0;pass<2;pass++){
      const bool boundp2=(pass==1);
      const unsigned cellinit=(boundp2? 0u: t.cellfluid);
      for(int z=ng.zini;z<ng.zfin;z++){
        const float oz=pp1.z-float(z-ng.cz)*c.scell;
        for(int y=ng.yini;y<ng.yfin;y++){
          const float oy=pp1.y-float(y-ng.cy)*c.scell;
          const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+cellinit;
          unsigned pini=t.begincell[v+ng.cxini];
          for(int x=ng.cxini;x<ng.cxfin;x++){
            const unsigned pfin=t.begincell[v+x+1];
            const float ox=pp1.x-float(x-ng.cx)*c.scell;
            for(unsigned p2=pini;p2<pfin;p2++){
              const float4 pp2=__ldg(&t.posp[p2]);
              const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
              const float rr2=drx*drx+dry*dry+drz*drz;
              if(rr2<=c.kernelsize2 && rr2>=1e-18f){
                const float4 velrhop2=__ldg(&t.velrhop[p2]);
                unsigned codep2=0; if(SHIFT && boundp2)codep2=t.code[p2];
                const float *taup2=(TVISCO==DSPH_VISCO_LAMINARSPS && !boundp2? t.spstau+6*size_t(p2): nullptr);
                PairFast<TKER,TVISCO,TDDT,SHIFT>(drx,dry,drz,rr2,q,velrhop2,pp2.w,boundp2,taup1,taup2,codep2,c,a,anybound,noshiftf);
              }
            }
            pini=pfin;
          }
        }
      }
      if(pass==0)FluidPassStore<TVISCO,SHIFT>(a);
    }
    float ar=a.ar, ax=a.ax, ay=a.ay, az=a.az;
    if(c.simulate2d)ay=0;
    if(TDDT!=DSPH_DDT_NONE && !anybound)ar+=a.delta;
    t.acear[p1]=make_float4(ax,ay,az,ar);
    if(SHIFT){ float4 s=a.shift; if(noshiftf)s.x=FLT_MAX; t.shiftposfs[p1]=s; }
    if(TVISCO==DSPH_VISCO_LAMINARSPS){
      const float g[6]={a.gxx,a.gxy,a.gxz,a.gyy,a.gyz,a.gzz};
      #pragma unroll
      for(int k=0;k<6;k++)t.spsgradvel[6*size_t(p1)+k]=g[k];
    }
    viscmax=a.visc;
    const bool normal=(!c.ghosts || (t.code[p1]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
    if(normal)ace2=ax*ax+ay*ay+az*az;
    vel2=q.velrhop.x*q.velrhop.x+q.velrhop.y*q.velrhop.y+q.velrhop.z*q.velrhop.z;
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> static void LaunchFluidV4(unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  ker_forces_fluid_v4<TKER,TVISCO,TDDT,SHIFT><<<(npf+FORCES_THREADS-1)/FORCES_THREADS,FORCES_THREADS,0,stm>>>(t,c);
}

// ---- boundary particles, division-free pair arithmetic shared by the two production kernels below
template<int TKER> __device__ __forceinline__ void BoundPairFast(float drx,float dry,float drz,float rr2,const float4 &velrhop1,const float4 &velrhop2,const DevCtes &c,float &arp1,float &visc){
  const float rinv=FastRsqrt(rr2), rad=rr2*rinv, qq=rad*c.ovkernelh;
  float fac;
  if(TKER==DSPH_KERNEL_WENDLAND){ const float wqq1=1.f-0.5f*qq; fac=c.bwen_ovh*(wqq1*wqq1*wqq1); }
  else{
    if(rad>c.kernelh){ const float wqq1=2.0f-qq; fac=c.cub_c2*(wqq1*wqq1)*rinv; }
    else fac=c.cub_c1_ovh+c.cub_d1_ovh*qq;
  }
  const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
  arp1+=c.massfluid*(dvx*(fac*drx)+dvy*(fac*dry)+dvz*(fac*drz))*(velrhop1.w*FastRcp(velrhop2.w));
  const float dot=drx*dvx+dry*dvy+drz*dvz;
  visc=fmaxf(dot*FastRcp(rr2+c.eta2),visc);
}

// ---- boundary particles, variant 3: one THREAD per boundary particle (the structure of ker_forces_bound_v1) with the division-free
// arithmetic.  Used when there are enough boundary particles to fill the GPU with threads (>= BOUND_WARP_LIMIT): there the warp-per-
// particle kernel below pays its redundant per-lane loop overhead for nothing (DB(20 M), 1.1 M boundary particles: 0.86 ms vs 0.43 ms).
constexpr unsigned BOUND_WARP_LIMIT=262144u;
template<int TKER> __global__ void __launch_bounds__(FORCES_THREADS) ker_forces_bound_v3(ForceParams t,DevCtes c)
{
  const unsigned p1=blockIdx.x*blockDim.x+threadIdx.x;
  float viscmax=0,vel2=0;
  if(p1<t.npb){
    float arp1=0,visc=0;
    const float4 velrhop1=t.velrhop[p1];
    if(p1<t.npbok){
      const float4 pp1=t.posp[p1];
      const NgRange ng=NgInit(t.dcell[p1],t,c);
      for(int z=ng.zini;z<ng.zfin;z++){
        const float oz=pp1.z-float(z-ng.cz)*c.scell;
        for(int y=ng.yini;y<ng.yfin;y++){
          const float oy=pp1.y-float(y-ng.cy)*c.scell;
          const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
          unsigned pini=t.begincell[v+ng.cxini];
          for(int x=ng.cxini;x<ng.cxfin;x++){
            const unsigned pfin=t.begincell[v+x+1];
            const float ox=pp1.x-float(x-ng.cx)*c.scell;
            for(unsigned p2=pini;p2<pfin;p2++){
              const float4 pp2=__ldg(&t.posp[p2]);
              const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
              const float rr2=drx*drx+dry*dry+drz*drz;
              if(rr2<=c.kernelsize2 && rr2>=1e-18f)BoundPairFast<TKER>(drx,dry,drz,rr2,velrhop1,__ldg(&t.velrhop[p2]),c,arp1,visc);
            }
            pini=pfin;
          }
        }
      }
    }
    t.acear[p1]=make_float4(0,0,0,arp1);
    viscmax=visc;
    if(c.dtallparticles)vel2=velrhop1.x*velrhop1.x+velrhop1.y*velrhop1.y+velrhop1.z*velrhop1.z;
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
}

// ---- boundary particles, variant 2 (production below BOUND_WARP_LIMIT boundary particles): density rate + viscdt of S/JSphCpu.cpp:548-625 with ONE WARP per boundary particle and
// the division-free arithmetic of PairFast.  ker_forces_bound_v1 (forces.cuh: one thread per particle, IEEE divisions) stays as the
// cross-check.  Why: a boundary particle has one third of the neighbours of a fluid one and the boundary is a thin sheet, so one thread
// per particle leaves the GPU nearly empty on small and medium cases (C1, 170 k particles: 40 k boundary particles = 2.1 warps per SM
// scheduler, each walking ~500 candidates serially; it was the slowest kernel of the step below ~1 M particles).  With a warp per
// particle the lanes stride over the candidates of each cell (coalesced posp/velrhop reads), the partial sums meet in a shuffle
// reduction, and the warps loop over the particles grid-stride so that the two maxima cost one atomic per warp, not per particle.
template<int TKER> __global__ void __launch_bounds__(FORCES_THREADS) ker_forces_bound_v2(ForceParams t,DevCtes c)
{
  const unsigned lane=threadIdx.x&31u;
  const unsigned nwarps=(gridDim.x*blockDim.x)>>5;
  float viscmax=0,vel2=0;
  for(unsigned p1=(blockIdx.x*blockDim.x+threadIdx.x)>>5;p1<t.npb;p1+=nwarps){
    const float4 velrhop1=t.velrhop[p1];
    float arp1=0,visc=0;
    if(p1<t.npbok){
      const float4 pp1=t.posp[p1];
      const NgRange ng=NgInit(t.dcell[p1],t,c);
      for(int z=ng.zini;z<ng.zfin;z++){
        const float oz=pp1.z-float(z-ng.cz)*c.scell;
        for(int y=ng.yini;y<ng.yfin;y++){
          const float oy=pp1.y-float(y-ng.cy)*c.scell;
          const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
          unsigned pini=t.begincell[v+ng.cxini];
          for(int x=ng.cxini;x<ng.cxfin;x++){
            const unsigned pfin=t.begincell[v+x+1];
            const float ox=pp1.x-float(x-ng.cx)*c.scell;
            for(unsigned p2=pini+lane;p2<pfin;p2+=32u){
              const float4 pp2=__ldg(&t.posp[p2]);
              const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
              const float rr2=drx*drx+dry*dry+drz*drz;
              if(rr2<=c.kernelsize2 && rr2>=1e-18f)BoundPairFast<TKER>(drx,dry,drz,rr2,velrhop1,__ldg(&t.velrhop[p2]),c,arp1,visc);
            }
            pini=pfin;
          }
        }
      }
      #pragma unroll
      for(int o=16;o>0;o>>=1){ arp1+=__shfl_xor_sync(0xffffffffu,arp1,o); visc=fmaxf(visc,__shfl_xor_sync(0xffffffffu,visc,o)); }
    }
    if(!lane)t.acear[p1]=make_float4(0,0,0,arp1);
    viscmax=fmaxf(viscmax,visc);
    if(c.dtallparticles)vel2=fmaxf(vel2,velrhop1.x*velrhop1.x+velrhop1.y*velrhop1.y+velrhop1.z*velrhop1.z);
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
}

} // namespace dsph
