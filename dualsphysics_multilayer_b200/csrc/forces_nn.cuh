// forces_nn.cuh -- Interaction_Forces of the multiphase non-Newtonian ("multilayer") fork for sm_100a.
//
// Physics of one pair follows the reference CPU kernels, which are the oracle:
//   FDA approach  JSphCpu::InteractionForcesFluid_NN_FDA_All          M/JSphCpu_NN_FDA.cpp:129-340   (one pass)
//   SPH approach  InteractionForcesFluid_NN_SPH_PressGrad             M/JSphCpu_NN_SPH.cpp:452-620   (pass 1: pressure, continuity, grad v)
//                 InteractionForcesFluid_NN_SPH_Visco_eta              :174-228                       (per particle: D, eta_eff)
//                 InteractionForcesFluid_NN_SPH_Morris                 :337-446                       (pass 2a: Morris operator)
//                 ..._Visco_Stress_tensor + ..._ConsEq                 :128-169, :234-331             (pass 2b: div(tau))
//   tensors       JSphCpu_Tensors.cpp (GetVelocityGradients_FDA, GetStrainRateTensor[_tsym], GetEta_Effective, GetStressTensor[_sym])
// The reference GPU twin is M/JSphGpu_NN_ker.cu:461-1482 (KerInteractionForcesFluid_NN_FDA / _NN_SPH_*), which reads ten
// phases from __constant__ arrays; here the phases travel as a by-value kernel parameter (DevPhases) and every pair uses
// the cell-relative float positions + the per-particle pressure written by the reorder kernel (celldiv.cuh).
//
// Pair order, the fluid-fluid / fluid-boundary split and the order in which the partial sums are added are the oracle's,
// so Ace/Ar agree with the CPU restatement to rounding of the pair distance (tests/test_gpu_parity.py, nn_* cases).
#pragma once
#include "forces.cuh"
#include "forces_ring.cuh"

namespace dsph {

struct NNAcc { float visc,ar,delta,ax,ay,az; float4 shift; };
__device__ __forceinline__ void NNZero(NNAcc &a){ a.visc=a.ar=a.delta=a.ax=a.ay=a.az=0; a.shift=make_float4(0,0,0,0); }

// neighbours of one particle in the oracle's order; f(p2,drx,dry,drz,rr2,pressp2) for every accepted pair
template<bool BOUNDP2,class F> __device__ __forceinline__ void NNForPairs(const NgRange &ng,const float4 &pp1,const ForceParams &t,const DevCtes &c,F &&f)
{
  const unsigned cellinit=(BOUNDP2? 0u: t.cellfluid);
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
          if(rr2<=c.kernelsize2 && rr2>=1e-18f)f(p2,drx,dry,drz,rr2,pp2.w);
        }
        pini=pfin;
      }
    }
  }
}

// JSphCpu::GetEta_Effective M/JSphCpu_Tensors.cpp:113-141 (P: phase of p1; the other arguments come from the phase of p2)
__device__ __forceinline__ float NNEtaEffective(const DevPhase &P,float tau_yield,float dmagn,float visco,float m_nn,float n_nn){
  if(dmagn<=1e-18f)dmagn=1e-18f;
  float miou_yield=(P.tau_max? P.tau_max/(2.0f*dmagn): tau_yield/(2.0f*dmagn));
  const bool bi_region=(P.tau_max && dmagn<=P.tau_max/(2.f*P.bi_multi*visco));
  if(bi_region)miou_yield=P.bi_multi*visco;
  const float miouPap=miou_yield*(1.f-expf(-m_nn*dmagn));
  const bool capped=(miouPap>m_nn*tau_yield || dmagn==1e-18f);
  const float term1=(P.tau_max? miou_yield: (capped? m_nn*tau_yield: miouPap));
  const float miouHB=visco*powf(dmagn,(n_nn-1.0f));
  const float term2=(bi_region? visco: (capped? visco: miouHB));
  return term1+term2;
}

template<int TKER> __device__ __forceinline__ float NNTensil(float rr2,float rhopp1,float pressp1,float rhopp2,float pressp2,const DevCtes &c){
  if(TKER!=DSPH_KERNEL_CUBIC)return 0.f;
  const float rad=sqrtf(rr2), qq=rad/c.kernelh;
  float wab;
  if(rad>c.kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; wab=c.cub_a24*(wqq2*wqq1); }
  else{ const float wqq2=qq*qq; wab=c.cub_a2*(1.0f+(0.75f*qq-1.5f)*wqq2); }
  float fab=wab*c.cub_od_wdeltap; fab*=fab; fab*=fab;
  const float t1=(pressp1/(rhopp1*rhopp1))*(pressp1>0? 0.01f: -0.2f);
  const float t2=(pressp2/(rhopp2*rhopp2))*(pressp2>0? 0.01f: -0.2f);
  return fab*(t1+t2);
}

// pressure, continuity, DDT and shifting of one pair: M/JSphCpu_NN_FDA.cpp:205-251 == M/JSphCpu_NN_SPH.cpp:528-581
template<int TKER,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void NNPairBasic(
  unsigned pp1,unsigned pp2,unsigned codep2,float massp2,const float4 &velrhop1,float pressp1,float pressp2,const float4 &velrhop2,
  float drx,float dry,float drz,float rr2,float frx,float fry,float frz,float &dvx,float &dvy,float &dvz,
  const DevCtes &c,const DevPhases &ph,NNAcc &a)
{
  const float rhopp1=velrhop1.w;
  {
    const float prs=(pressp1+pressp2)/(rhopp1*velrhop2.w)+NNTensil<TKER>(rr2,rhopp1,pressp1,velrhop2.w,pressp2,c);
    const float p_vpm=-prs*massp2;
    a.ax+=p_vpm*frx; a.ay+=p_vpm*fry; a.az+=p_vpm*frz;
  }
  const float rhop1over2=rhopp1/velrhop2.w;
  dvx=velrhop1.x-velrhop2.x; dvy=velrhop1.y-velrhop2.y; dvz=velrhop1.z-velrhop2.z;
  a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*rhop1over2;
  const float cbar=ph.ph[pp2].cs0;
  if(TDDT==DSPH_DDT_DDT){
    if(a.delta!=FLT_MAX){
      const float visc_densi=c.ddtkh*cbar*(rhop1over2-1.f)/(rr2+c.eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      const float delta=(pp1==pp2? visc_densi*dot3*massp2: 0.f);
      a.delta=(BOUNDP2? FLT_MAX: a.delta+delta);
    }
  }
  if(TDDT==DSPH_DDT_DDT2 || (TDDT==DSPH_DDT_DDT2FULL && !BOUNDP2)){
    if(a.delta!=FLT_MAX){
      const float rh=1.f+c.ddtgz*drz;
      const float drhop=c.rhopzero*(PowNearOne(rh,c)-1.f);
      const float visc_densi=c.ddtkh*cbar*((velrhop2.w-rhopp1)-drhop)/(rr2+c.eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      const float delta=(pp1==pp2? visc_densi*dot3*massp2/velrhop2.w: 0.f);
      a.delta=(BOUNDP2? FLT_MAX: a.delta-delta);
    }
  }
  if(SHIFT){
    if(a.shift.x!=FLT_MAX){
      const bool heavyphase=(ph.ph[pp1].mass>ph.ph[pp2].mass && pp1!=pp2);
      const float massrhop=massp2/velrhop2.w;
      const bool noshift=(BOUNDP2 && (c.shiftmode==DSPH_SHIFT_NOBOUND || (c.shiftmode==DSPH_SHIFT_NOFIXED && (codep2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FIXED)));
      a.shift.x=(noshift? FLT_MAX: (heavyphase? 0.f: a.shift.x+massrhop*frx));
      a.shift.y+=(heavyphase? 0.f: massrhop*fry);
      a.shift.z+=(heavyphase? 0.f: massrhop*frz);
      a.shift.w-=(heavyphase? 0.f: massrhop*(drx*frx+dry*fry+drz*frz));
    }
  }
}

// result of the two passes of one particle, combined as the two CPU calls do (store condition of M/JSphCpu_NN_FDA.cpp:317-323)
struct NNSum { float ar,ax,ay,az,delta,visc; };
template<int TDDT,bool SHIFT> __device__ __forceinline__ bool NNStore(const NNAcc &a,NNSum &s){
  if(SHIFT||a.ar||a.ax||a.ay||a.az||a.visc){
    if(TDDT!=DSPH_DDT_NONE)s.delta=(s.delta==FLT_MAX || a.delta==FLT_MAX? FLT_MAX: s.delta+a.delta);
    s.ar+=a.ar; s.ax+=a.ax; s.ay+=a.ay; s.az+=a.az;
    s.visc=fmaxf(s.visc,a.visc);
    return true;
  }
  return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// FDA approach: everything in one pass per neighbour class
// ---------------------------------------------------------------------------------------------------------------------
template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void NNPassFDA(
  const NgRange &ng,const float4 &pp1,const float4 &velrhop1,unsigned pp1ph,const ForceParams &t,const DevCtes &c,const DevPhases &ph,
  NNAcc &a,float &visceta)
{
  NNForPairs<BOUNDP2>(ng,pp1,t,c,[&](unsigned p2,float drx,float dry,float drz,float rr2,float pressp2){
    const float fac=KernelFac<TKER>(rr2,c);
    const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
    unsigned codep2=0; if(!BOUNDP2 || SHIFT)codep2=t.code[p2];
    const unsigned pp2=(BOUNDP2? pp1ph: (codep2&DSPH_CODE_MASKVALUE));
    const float massp2=(BOUNDP2? c.massbound: ph.ph[pp2].mass);
    const float4 velrhop2=__ldg(&t.velrhop[p2]);
    float dvx,dvy,dvz;
    NNPairBasic<TKER,TDDT,SHIFT,BOUNDP2>(pp1ph,pp2,codep2,massp2,velrhop1,pp1.w,pressp2,velrhop2,drx,dry,drz,rr2,frx,fry,frz,dvx,dvy,dvz,c,ph,a);
    const float cbar=ph.ph[pp2].cs0;
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot/(rr2+c.eta2);
    a.visc=fmaxf(dot_rr2,a.visc);
    const float visco_NN=ph.ph[pp2].visco;
    if(TVISCO==DSPH_VISCO_ARTIFICIAL){
      if(dot<0){
        const float amubar=c.kernelh*dot_rr2;
        const float robar=(velrhop1.w+velrhop2.w)*0.5f;
        const float pi_visc=(-visco_NN*cbar*amubar/robar)*massp2;
        a.ax-=pi_visc*frx; a.ay-=pi_visc*fry; a.az-=pi_visc*frz;
      }
    }
    else{
      if(BOUNDP2){ dvx=2.f*velrhop1.x; dvy=2.f*velrhop1.y; dvz=2.f*velrhop1.z; }      // no-slip on the tensor
      const float a11=dvx*drx/rr2,a12=dvx*dry/rr2,a13=dvx*drz/rr2;
      const float a21=dvy*drx/rr2,a22=dvy*dry/rr2,a23=dvy*drz/rr2;
      const float a31=dvz*drx/rr2,a32=dvz*dry/rr2,a33=dvz*drz/rr2;
      const float div_vel=(a11+a22+a33)/3.f;
      const float d11=a11-div_vel, d12=0.5f*(a12+a21), d13=0.5f*(a13+a31);
      const float d21=0.5f*(a21+a12), d22=a22-div_vel, d23=0.5f*(a23+a32);
      const float d31=0.5f*(a31+a13), d32=0.5f*(a32+a23), d33=a33-div_vel;
      const float II_D_1=d11*d22+d22*d33+d11*d33;
      const float II_D_2=d12*d21+d23*d32+d13*d31;
      const float II_D=II_D_1-II_D_2;
      const float dmagn=sqrtf(II_D*II_D);
      const float eta=NNEtaEffective(ph.ph[pp1ph],ph.ph[pp2].tau_yield,dmagn,visco_NN,ph.ph[pp2].m_nn,ph.ph[pp2].n_nn);
      visceta=fmaxf(eta,visceta);
      if(TVISCO==DSPH_VISCO_LAMINARSPS){
        const float temp=2.0f*eta/((rr2+c.eta2)*velrhop2.w);
        const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
        a.ax+=vtemp*dvx; a.ay+=vtemp*dvy; a.az+=vtemp*dvz;
      }
      else{
        const float t2=2.f*eta;
        const float taux=((t2*d11)*frx+(t2*d12)*fry+(t2*d13)*frz)/(velrhop2.w);
        const float tauy=((t2*d21)*frx+(t2*d22)*fry+(t2*d23)*frz)/(velrhop2.w);
        const float tauz=((t2*d31)*frx+(t2*d32)*fry+(t2*d33)*frz)/(velrhop2.w);
        a.ax+=taux*massp2; a.ay+=tauy*massp2; a.az+=tauz*massp2;
      }
    }
  });
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS)
ker_nn_fda(ForceParams t,DevCtes c,DevPhases ph)
{
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  float viscmax=0,ace2=0,vel2=0,etamax=0;
  if(p1<t.np){
    const unsigned code1=t.code[p1];
    const bool ghost=((code1&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC);
    const float4 velrhop1=t.velrhop[p1];
    if(!(t.skipghosts && ghost)){
      const float4 pp1=t.posp[p1];
      const unsigned pp1ph=code1&DSPH_CODE_MASKVALUE;
      const NgRange ng=NgInit(t.dcell[p1],t,c);
      NNSum s; s.ar=s.ax=s.ay=s.az=s.delta=s.visc=0;
      NNAcc a; NNZero(a);
      float visceta=0;
      NNPassFDA<TKER,TVISCO,TDDT,SHIFT,false>(ng,pp1,velrhop1,pp1ph,t,c,ph,a,visceta);
      if(NNStore<TDDT,SHIFT>(a,s))etamax=fmaxf(etamax,visceta);
      const float4 shiftff=a.shift;
      NNZero(a); if(SHIFT)a.shift=shiftff; visceta=0;
      NNPassFDA<TKER,TVISCO,TDDT,SHIFT,true>(ng,pp1,velrhop1,pp1ph,t,c,ph,a,visceta);
      if(NNStore<TDDT,SHIFT>(a,s))etamax=fmaxf(etamax,visceta);
      if(c.simulate2d)s.ay=0;
      if(TDDT!=DSPH_DDT_NONE && s.delta!=FLT_MAX)s.ar+=s.delta;
      t.acear[p1]=make_float4(s.ax,s.ay,s.az,s.ar);
      if(SHIFT)t.shiftposfs[p1]=a.shift;
      viscmax=s.visc;
      if(!c.ghosts || !ghost)ace2=s.ax*s.ax+s.ay*s.ay+s.az*s.az;
    }
    vel2=velrhop1.x*velrhop1.x+velrhop1.y*velrhop1.y+velrhop1.z*velrhop1.z;
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
  WarpMaxToGlobal(etamax,&t.fmax->visceta_bits);
}

// ---------------------------------------------------------------------------------------------------------------------
// SPH approach, pass 1: pressure, continuity, DDT, shifting + symmetric velocity gradient
// ---------------------------------------------------------------------------------------------------------------------
template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void NNPassPressGrad(
  const NgRange &ng,const float4 &pp1,const float4 &velrhop1,unsigned pp1ph,const ForceParams &t,const DevCtes &c,const DevPhases &ph,
  NNAcc &a,float *g)
{
  NNForPairs<BOUNDP2>(ng,pp1,t,c,[&](unsigned p2,float drx,float dry,float drz,float rr2,float pressp2){
    const float fac=KernelFac<TKER>(rr2,c);
    const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
    unsigned codep2=0; if(!BOUNDP2 || SHIFT)codep2=t.code[p2];
    const unsigned pp2=(BOUNDP2? pp1ph: (codep2&DSPH_CODE_MASKVALUE));
    const float massp2=(BOUNDP2? c.massbound: ph.ph[pp2].mass);
    const float4 velrhop2=__ldg(&t.velrhop[p2]);
    float dvx,dvy,dvz;
    NNPairBasic<TKER,TDDT,SHIFT,BOUNDP2>(pp1ph,pp2,codep2,massp2,velrhop1,pp1.w,pressp2,velrhop2,drx,dry,drz,rr2,frx,fry,frz,dvx,dvy,dvz,c,ph,a);
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot/(rr2+c.eta2);
    a.visc=fmaxf(dot_rr2,a.visc);
    if(TVISCO!=DSPH_VISCO_ARTIFICIAL){
      if(BOUNDP2){ dvx=2.f*velrhop1.x; dvy=2.f*velrhop1.y; dvz=2.f*velrhop1.z; }
      const float volp2=-massp2/velrhop2.w;
      float dv=dvx*volp2; g[0]+=dv*frx; g[1]+=dv*fry; g[2]+=dv*frz;
            dv=dvy*volp2; g[1]+=dv*frx; g[3]+=dv*fry; g[4]+=dv*frz;
            dv=dvz*volp2; g[2]+=dv*frx; g[4]+=dv*fry; g[5]+=dv*frz;
    }
  });
}

// Writes acear (pressure part), shiftposfs, viscdt/vel maxima and, per particle, eta_eff (+ tau for the constitutive eq.)
template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS)
ker_nn_sph_pressgrad(ForceParams t,DevCtes c,DevPhases ph)
{
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  float viscmax=0,vel2=0,etamax=0;
  if(p1<t.np){
    const unsigned code1=t.code[p1];
    const float4 velrhop1=t.velrhop[p1];
    const float4 pp1=t.posp[p1];
    const unsigned pp1ph=code1&DSPH_CODE_MASKVALUE;
    const NgRange ng=NgInit(t.dcell[p1],t,c);
    NNSum s; s.ar=s.ax=s.ay=s.az=s.delta=s.visc=0;
    float gs[6]={0,0,0,0,0,0};
    NNAcc a; NNZero(a);
    float g[6]={0,0,0,0,0,0};
    NNPassPressGrad<TKER,TVISCO,TDDT,SHIFT,false>(ng,pp1,velrhop1,pp1ph,t,c,ph,a,g);
    if(NNStore<TDDT,SHIFT>(a,s)){
      #pragma unroll
      for(int k=0;k<6;k++)gs[k]+=g[k];
    }
    const float4 shiftff=a.shift;
    NNZero(a); if(SHIFT)a.shift=shiftff;
    #pragma unroll
    for(int k=0;k<6;k++)g[k]=0;
    NNPassPressGrad<TKER,TVISCO,TDDT,SHIFT,true>(ng,pp1,velrhop1,pp1ph,t,c,ph,a,g);
    if(NNStore<TDDT,SHIFT>(a,s)){
      #pragma unroll
      for(int k=0;k<6;k++)gs[k]+=g[k];
    }
    if(TDDT!=DSPH_DDT_NONE && s.delta!=FLT_MAX)s.ar+=s.delta;     // Deltac is folded in after all passes; it only touches ar
    t.acear[p1]=make_float4(s.ax,s.ay,s.az,s.ar);
    if(SHIFT)t.shiftposfs[p1]=a.shift;
    // x-slab halo copies (their neighbourhood is complete only for the inner layer) take no part in the maxima
    const bool halo=(t.skipghosts && (code1&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC);
    if(!halo)viscmax=s.visc;
    vel2=velrhop1.x*velrhop1.x+velrhop1.y*velrhop1.y+velrhop1.z*velrhop1.z;
    if(TVISCO!=DSPH_VISCO_ARTIFICIAL){
      // InteractionForcesFluid_NN_SPH_Visco_eta + GetStrainRateTensor_tsym (M/JSphCpu_NN_SPH.cpp:174-228, Tensors:184-207)
      const float div_vel=(gs[0]+gs[3]+gs[5])/3.f;
      const float dxx=gs[0]-div_vel, dxy=0.5f*gs[1], dxz=0.5f*gs[2], dyy=gs[3]-div_vel, dyz=0.5f*gs[4], dzz=gs[5]-div_vel;
      const float II_D_1=dxx*dyy+dyy*dzz+dxx*dzz;
      const float II_D_2=dxy*dxy+dyz*dyz+dxz*dxz;
      const float II_D=-II_D_1+II_D_2;
      const float dmagn=sqrtf(II_D);
      const DevPhase &P=ph.ph[pp1ph];
      const float eta=NNEtaEffective(P,P.tau_yield,dmagn,P.visco,P.m_nn,P.n_nn);
      t.viscoeta[p1]=eta;
      if(!halo)etamax=eta;
      if(TVISCO==DSPH_VISCO_CONSTEQ){          // GetStressTensor_sym (Tensors:211-231)
        const float t2=2.f*eta;
        float *tau=t.nntau+6*size_t(p1);
        tau[0]=t2*dxx; tau[1]=t2*dxy; tau[2]=t2*dxz; tau[3]=t2*dyy; tau[4]=t2*dyz; tau[5]=t2*dzz;
      }
    }
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
  WarpMaxToGlobal(etamax,&t.fmax->visceta_bits);
}

// ---------------------------------------------------------------------------------------------------------------------
// SPH approach, pass 2: viscous forces (artificial / Morris operator / divergence of the stress tensor)
// ---------------------------------------------------------------------------------------------------------------------
template<int TKER,int TVISCO,bool BOUNDP2> __device__ __forceinline__ void NNPassVisco(
  const NgRange &ng,const float4 &pp1,const float4 &velrhop1,unsigned pp1ph,float eta1,const float *tau1,
  const ForceParams &t,const DevCtes &c,const DevPhases &ph,float &ax,float &ay,float &az)
{
  NNForPairs<BOUNDP2>(ng,pp1,t,c,[&](unsigned p2,float drx,float dry,float drz,float rr2,float){
    const float fac=KernelFac<TKER>(rr2,c);
    const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
    const unsigned pp2=(BOUNDP2? pp1ph: (t.code[p2]&DSPH_CODE_MASKVALUE));
    const float massp2=(BOUNDP2? c.massbound: ph.ph[pp2].mass);
    const float4 velrhop2=__ldg(&t.velrhop[p2]);
    if(TVISCO==DSPH_VISCO_CONSTEQ){
      float ts[6];
      #pragma unroll
      for(int k=0;k<6;k++)ts[k]=tau1[k]+(BOUNDP2? tau1[k]: t.nntau[6*size_t(p2)+k]);
      const float taux=(ts[0]*frx+ts[1]*fry+ts[2]*frz)/(velrhop2.w);
      const float tauy=(ts[1]*frx+ts[3]*fry+ts[4]*frz)/(velrhop2.w);
      const float tauz=(ts[2]*frx+ts[4]*fry+ts[5]*frz)/(velrhop2.w);
      ax+=taux*massp2; ay+=tauy*massp2; az+=tauz*massp2;
    }
    else{
      float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
      if(BOUNDP2){ dvx=2.f*velrhop1.x; dvy=2.f*velrhop1.y; dvz=2.f*velrhop1.z; }
      if(TVISCO==DSPH_VISCO_ARTIFICIAL){
        const float dot=drx*dvx+dry*dvy+drz*dvz;
        if(dot<0){
          const float dot_rr2=dot/(rr2+c.eta2);
          const float amubar=c.kernelh*dot_rr2;
          const float robar=(velrhop1.w+velrhop2.w)*0.5f;
          const float pi_visc=(-ph.ph[pp2].visco*ph.ph[pp2].cs0*amubar/robar)*massp2;
          ax-=pi_visc*frx; ay-=pi_visc*fry; az-=pi_visc*frz;
        }
      }
      else{
        const float eta2=(BOUNDP2? eta1: t.viscoeta[p2]);
        const float temp=(eta1+eta2)/((rr2+c.eta2)*velrhop2.w);
        const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
        ax+=vtemp*dvx; ay+=vtemp*dvy; az+=vtemp*dvz;
      }
    }
  });
}

template<int TKER,int TVISCO> __global__ void __launch_bounds__(FORCES_THREADS)
ker_nn_sph_visco(ForceParams t,DevCtes c,DevPhases ph)
{
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  float ace2=0;
  if(p1<t.np){
    const unsigned code1=t.code[p1];
    const float4 velrhop1=t.velrhop[p1];
    const float4 pp1=t.posp[p1];
    const unsigned pp1ph=code1&DSPH_CODE_MASKVALUE;
    const NgRange ng=NgInit(t.dcell[p1],t,c);
    float eta1=0; float tau1[6]={0,0,0,0,0,0};
    if(TVISCO==DSPH_VISCO_LAMINARSPS)eta1=t.viscoeta[p1];
    if(TVISCO==DSPH_VISCO_CONSTEQ){
      #pragma unroll
      for(int k=0;k<6;k++)tau1[k]=t.nntau[6*size_t(p1)+k];
    }
    float4 acc=t.acear[p1];
    float ax=0,ay=0,az=0;
    NNPassVisco<TKER,TVISCO,false>(ng,pp1,velrhop1,pp1ph,eta1,tau1,t,c,ph,ax,ay,az);
    if(ax||ay||az){ acc.x=acc.x+ax; acc.y=acc.y+ay; acc.z=acc.z+az; }
    ax=ay=az=0;
    NNPassVisco<TKER,TVISCO,true>(ng,pp1,velrhop1,pp1ph,eta1,tau1,t,c,ph,ax,ay,az);
    if(ax||ay||az){ acc.x=acc.x+ax; acc.y=acc.y+ay; acc.z=acc.z+az; }
    if(c.simulate2d)acc.y=0;
    t.acear[p1]=acc;
    const bool ghost=((code1&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC);
    if(!c.ghosts || !ghost)ace2=acc.x*acc.x+acc.y*acc.y+acc.z*acc.z;
  }
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
}

// ---------------------------------------------------------------------------------------------------------------------
// FDA approach, production kernel: the deferred-ring structure of forces_ring.cuh (candidate test -> per-lane ring in shared
// memory -> pair rounds at full lane occupancy) with a division-free multiphase pair body.  Same pair order as the plain
// kernel above (which stays as the IEEE cross-check, variant 1); differs from it only by the rounding of MUFU.RCP/RSQ/EX2.
// ---------------------------------------------------------------------------------------------------------------------
struct NNFastP1 { float4 velrhop; float press, ovrhop; unsigned ph; float tau_max, bi_multi, mass; };
struct NNFastAcc { float ax,ay,az,ar,delta,visc,visceta; float4 shift; };

__device__ __forceinline__ float NNEtaFast(float tau_max1,float bi_multi1,float tau_yield,float dmagn,float visco,float m_nn,float n_nn){
  const float dm=fmaxf(dmagn,1e-18f);
  const float ov2d=FastRcp(2.0f*dm);
  float miou_yield=(tau_max1? tau_max1: tau_yield)*ov2d;
  bool bi_region=false;
  if(tau_max1){ bi_region=(dm<=tau_max1*FastRcp(2.f*bi_multi1*visco)); if(bi_region)miou_yield=bi_multi1*visco; }
  const float miouPap=miou_yield*(1.f-__expf(-m_nn*dm));
  const bool capped=(miouPap>m_nn*tau_yield || dm==1e-18f);
  const float term1=(tau_max1? miou_yield: (capped? m_nn*tau_yield: miouPap));
  const float miouHB=(n_nn==1.0f? visco: visco*__powf(dm,(n_nn-1.0f)));
  const float term2=(bi_region? visco: (capped? visco: miouHB));
  return term1+term2;
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> __device__ __forceinline__ void NNPairFast(
  float drx,float dry,float drz,float rr2,const NNFastP1 &p1,const float4 &velrhop2,float pressp2,bool boundp2,unsigned codep2,bool valid,
  const DevCtes &c,const DevPhases &ph,NNFastAcc &a,unsigned &anybound,unsigned &noshiftf)
{
  const unsigned pp2=(boundp2? p1.ph: (codep2&DSPH_CODE_MASKVALUE));
  const DevPhase &P2=ph.ph[pp2];
  const float massp2=(valid? (boundp2? c.massbound: P2.mass): 0.f);
  const bool boundflag=(boundp2 && valid);
  const float rinv=FastRsqrt(rr2);
  const float rad=rr2*rinv;
  const float qq=rad*c.ovkernelh;
  float fac;
  if(TKER==DSPH_KERNEL_WENDLAND){ const float wqq1=1.f-0.5f*qq; fac=c.bwen_ovh*(wqq1*wqq1*wqq1); }
  else{
    if(rad>c.kernelh){ const float wqq1=2.0f-qq; fac=c.cub_c2*(wqq1*wqq1)*rinv; }
    else fac=c.cub_c1_ovh+c.cub_d1_ovh*qq;
  }
  float frx,fry; const float frz=fac*drz;
  Upk(Mul2(Pk(fac,fac),Pk(drx,dry)),frx,fry);          // packed fp32x2 (forces_fast.cuh): same rounding, one issue slot
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
    Upk(Fma2(Pk(p_vpm,p_vpm),Pk(frx,fry),Pk(a.ax,a.ay)),a.ax,a.ay);
    a.az+=p_vpm*frz;
  }
  float dvx,dvy,dvz=p1.velrhop.z-velrhop2.z;
  Upk(Sub2(Pk(p1.velrhop.x,p1.velrhop.y),Pk(velrhop2.x,velrhop2.y)),dvx,dvy);
  a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhopp1*ov2);
  const float ovrr=FastRcp(rr2+c.eta2);
  const float cbar=P2.cs0;
  const float dot3=(drx*frx+dry*fry+drz*frz);
  const bool samephase=(p1.ph==pp2);
  if(TDDT==DSPH_DDT_DDT){
    if(boundflag)anybound=1u;
    if(!boundp2 && samephase){
      const float visc_densi=c.ddtkh*cbar*(__fdiv_rn(rhopp1,velrhop2.w)-1.f)*ovrr;
      a.delta+=visc_densi*dot3*massp2;
    }
  }
  if(TDDT==DSPH_DDT_DDT2 || TDDT==DSPH_DDT_DDT2FULL){
    if(boundflag && TDDT==DSPH_DDT_DDT2)anybound=1u;
    if(!boundp2 && samephase){
      const float rh=1.f+c.ddtgz*drz;
      const float drhop=c.rhopzero*(PowNearOne(rh,c)-1.f);
      const float visc_densi=c.ddtkh*cbar*((velrhop2.w-rhopp1)-drhop)*ovrr;
      a.delta-=visc_densi*dot3*massp2*ov2;
    }
  }
  if(SHIFT){
    if(boundflag && (c.shiftmode==DSPH_SHIFT_NOBOUND || (c.shiftmode==DSPH_SHIFT_NOFIXED && (codep2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FIXED)))noshiftf=1u;
    // a heavier particle next to a lighter phase resets the x accumulator and skips the rest (M/JSphCpu_NN_FDA.cpp:244-250)
    const bool heavyphase=(valid && p1.mass>P2.mass && !samephase);
    const float massrhop=massp2*ov2;
    a.shift.x=(heavyphase? 0.f: a.shift.x+massrhop*frx);
    if(!heavyphase){ a.shift.y+=massrhop*fry; a.shift.z+=massrhop*frz; a.shift.w-=massrhop*dot3; }
  }
  {
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot*ovrr;
    a.visc=fmaxf(dot_rr2,a.visc);
    const float visco_NN=P2.visco;
    if(TVISCO==DSPH_VISCO_ARTIFICIAL){
      if(dot<0){
        const float amubar=c.kernelh*dot_rr2;
        const float ovrobar=FastRcp((rhopp1+velrhop2.w)*0.5f);
        const float pi_visc=(-visco_NN*cbar*amubar*ovrobar)*massp2;
        a.ax-=pi_visc*frx; a.ay-=pi_visc*fry; a.az-=pi_visc*frz;
      }
    }
    else{
      if(boundp2){ dvx=2.f*p1.velrhop.x; dvy=2.f*p1.velrhop.y; dvz=2.f*p1.velrhop.z; }
      const float ovr2=rinv*rinv;
      const float sx=drx*ovr2, sy=dry*ovr2, sz=drz*ovr2;
      const float a11=dvx*sx,a12=dvx*sy,a13=dvx*sz;
      const float a21=dvy*sx,a22=dvy*sy,a23=dvy*sz;
      const float a31=dvz*sx,a32=dvz*sy,a33=dvz*sz;
      const float div_vel=(a11+a22+a33)*(1.f/3.f);
      const float d11=a11-div_vel, d22=a22-div_vel, d33=a33-div_vel;
      const float d12=0.5f*(a12+a21), d13=0.5f*(a13+a31), d23=0.5f*(a23+a32);
      const float II_D=(d11*d22+d22*d33+d11*d33)-(d12*d12+d23*d23+d13*d13);
      const float dmagn=fabsf(II_D);
      const float eta=NNEtaFast(p1.tau_max,p1.bi_multi,P2.tau_yield,dmagn,visco_NN,P2.m_nn,P2.n_nn);
      if(valid)a.visceta=fmaxf(eta,a.visceta);
      if(TVISCO==DSPH_VISCO_LAMINARSPS){
        const float temp=2.0f*eta*ovrr*ov2;
        const float vtemp=massp2*temp*dot3;
        Upk(Fma2(Pk(vtemp,vtemp),Pk(dvx,dvy),Pk(a.ax,a.ay)),a.ax,a.ay);
        a.az+=vtemp*dvz;
      }
      else{
        const float t2=2.f*eta*ov2*massp2;
        a.ax+=t2*(d11*frx+d12*fry+d13*frz);
        a.ay+=t2*(d12*frx+d22*fry+d23*frz);
        a.az+=t2*(d13*frx+d23*fry+d33*frz);
      }
    }
  }
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS,7)
ker_nn_fda_ring(ForceParams t,DevCtes c,DevPhases ph)
{
  __shared__ __align__(16) V8Smem qsm;
  const unsigned q=V8Setup(qsm.w,c.scell);
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned p1c=min(p1,t.np-1);
  const unsigned code1=t.code[p1c];
  const bool ghost=((code1&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC);
  const bool active=(p1<t.np && !(t.skipghosts && ghost));
  const float4 pp1=t.posp[p1c];
  NNFastP1 fp; fp.velrhop=t.velrhop[p1c]; fp.press=pp1.w; fp.ovrhop=1.f/fp.velrhop.w;
  fp.ph=code1&DSPH_CODE_MASKVALUE; if(fp.ph>=unsigned(ph.count))fp.ph=0;
  fp.tau_max=ph.ph[fp.ph].tau_max; fp.bi_multi=ph.ph[fp.ph].bi_multi; fp.mass=ph.ph[fp.ph].mass;
  NNFastAcc a; a.ax=a.ay=a.az=a.ar=a.delta=a.visc=a.visceta=0; a.shift=make_float4(0,0,0,0);
  unsigned anyb=0,nosh=0,ffdone=0;
  // Store condition of the reference (M/JSphCpu_NN_FDA.cpp:317-323, NNStore above): the results of the fluid-fluid pass -- the DDT term and
  // the viscous-dt maximum with them -- are kept only when one of ar, ace, viscdt is non-zero (always with shifting).  It matters for a
  // fluid at rest with uniform pressure (step 0 of a stacked-layer case): there the reference drops the DDT term.  The ring has been
  // drained when the first boundary pair arrives, so the accumulators hold exactly the fluid-fluid sums at that moment.
  auto ffstore=[&]{ ffdone=1u; if(!SHIFT && a.ar==0 && a.ax==0 && a.ay==0 && a.az==0 && a.visc==0){ a.delta=0; a.visceta=0; } };
  V8Traverse(t,c,active,pp1,t.dcell[p1c],q,[&](unsigned p2,float drx,float dry,float drz,float rr2,float pressp2,bool boundp2,bool valid){
    if(boundp2 && !ffdone)ffstore();
    const float4 vr2=__ldg(&t.velrhop[p2]);
    unsigned codep2=0; if(!boundp2 || SHIFT)codep2=t.code[p2];
    NNPairFast<TKER,TVISCO,TDDT,SHIFT>(drx,dry,drz,rr2,fp,vr2,pressp2,boundp2,codep2,valid,c,ph,a,anyb,nosh);
  });
  if(!ffdone)ffstore();
  float viscmax=0,ace2=0,vel2=0,etamax=0;
  if(active){
    float ar=a.ar, ax=a.ax, ay=a.ay, az=a.az;
    if(c.simulate2d)ay=0;
    if(TDDT!=DSPH_DDT_NONE && !anyb)ar+=a.delta;
    t.acear[p1]=make_float4(ax,ay,az,ar);
    if(SHIFT){ float4 s=a.shift; if(nosh)s.x=FLT_MAX; t.shiftposfs[p1]=s; }
    viscmax=a.visc; etamax=a.visceta;
    if(!c.ghosts || !ghost)ace2=ax*ax+ay*ay+az*az;
  }
  if(p1<t.np)vel2=fp.velrhop.x*fp.velrhop.x+fp.velrhop.y*fp.velrhop.y+fp.velrhop.z*fp.velrhop.z;
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
  WarpMaxToGlobal(etamax,&t.fmax->visceta_bits);
}

// ---------------------------------------------------------------------------------------------------------------------
// SPH approach on the ring structure.  RingTraverse is the skeleton of forces_ring.cuh (cell loops, chunked candidate test,
// per-lane ring, rounds at full lane occupancy) with the pair body passed in: pair(p2,drx,dry,drz,rr2,pressp2,boundp2,valid).
// ---------------------------------------------------------------------------------------------------------------------
template<class PAIR> __device__ __forceinline__ void RingTraverse(const ForceParams &t,const DevCtes &c,bool active,const float4 &pp1,unsigned rcell,
  unsigned q,PAIR &&pair)
{
  V8Traverse(t,c,active,pp1,rcell,q,pair);      // forces_ring.cuh: the skeleton of the production single-phase kernel (variant 8)
}

template<int TKER> __device__ __forceinline__ float NNFacFast(float rr2,const DevCtes &c,float &rinv,float &rad,float &qq){
  rinv=FastRsqrt(rr2); rad=rr2*rinv; qq=rad*c.ovkernelh;
  if(TKER==DSPH_KERNEL_WENDLAND){ const float wqq1=1.f-0.5f*qq; return c.bwen_ovh*(wqq1*wqq1*wqq1); }
  if(rad>c.kernelh){ const float wqq1=2.0f-qq; return c.cub_c2*(wqq1*wqq1)*rinv; }
  return c.cub_c1_ovh+c.cub_d1_ovh*qq;
}

// pass 1: pressure, continuity, DDT, shifting, viscdt + symmetric velocity gradient; epilogue: D, eta_eff (+ tau)
template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS,6)
ker_nn_sph_pressgrad_ring(ForceParams t,DevCtes c,DevPhases ph)
{
  __shared__ __align__(16) V8Smem qsm;
  const unsigned q=V8Setup(qsm.w,c.scell);
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned p1c=min(p1,t.np-1);
  const bool active=(p1<t.np);
  const unsigned code1=t.code[p1c];
  const float4 pp1=t.posp[p1c];
  NNFastP1 fp; fp.velrhop=t.velrhop[p1c]; fp.press=pp1.w; fp.ovrhop=1.f/fp.velrhop.w;
  fp.ph=code1&DSPH_CODE_MASKVALUE; if(fp.ph>=unsigned(ph.count))fp.ph=0;
  fp.tau_max=ph.ph[fp.ph].tau_max; fp.bi_multi=ph.ph[fp.ph].bi_multi; fp.mass=ph.ph[fp.ph].mass;
  NNFastAcc a; a.ax=a.ay=a.az=a.ar=a.delta=a.visc=a.visceta=0; a.shift=make_float4(0,0,0,0);
  float g0=0,g1=0,g2=0,g3=0,g4=0,g5=0;
  unsigned anyb=0,nosh=0,ffdone=0;
  // store condition of the fluid-fluid pass (see ker_nn_fda_ring): DDT term and velocity gradient are dropped when ar, ace and viscdt are all zero
  auto ffstore=[&]{ ffdone=1u; if(!SHIFT && a.ar==0 && a.ax==0 && a.ay==0 && a.az==0 && a.visc==0){ a.delta=0; g0=g1=g2=g3=g4=g5=0; } };
  RingTraverse(t,c,active,pp1,t.dcell[p1c],q,[&](unsigned p2,float drx,float dry,float drz,float rr2,float pressp2,bool boundp2,bool valid){
    if(boundp2 && !ffdone)ffstore();
    const float4 velrhop2=__ldg(&t.velrhop[p2]);
    unsigned codep2=0; if(!boundp2 || SHIFT)codep2=t.code[p2];
    const unsigned pp2=(boundp2? fp.ph: (codep2&DSPH_CODE_MASKVALUE));
    const DevPhase &P2=ph.ph[pp2];
    const float massp2=(valid? (boundp2? c.massbound: P2.mass): 0.f);
    const bool boundflag=(boundp2 && valid);
    float rinv,rad,qq; const float fac=NNFacFast<TKER>(rr2,c,rinv,rad,qq);
    const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
    const float rhopp1=fp.velrhop.w;
    const float ov2=FastRcp(velrhop2.w);
    {
      float prs=(fp.press+pressp2)*(fp.ovrhop*ov2);
      if(TKER==DSPH_KERNEL_CUBIC){
        float wab;
        if(rad>c.kernelh){ const float wqq1=2.0f-qq; wab=c.cub_a24*(wqq1*wqq1*wqq1); }
        else wab=c.cub_a2*(1.0f+(0.75f*qq-1.5f)*(qq*qq));
        float fab=wab*c.cub_od_wdeltap; fab*=fab; fab*=fab;
        const float t1=(fp.press*fp.ovrhop*fp.ovrhop)*(fp.press>0? 0.01f: -0.2f);
        const float t2=(pressp2*ov2*ov2)*(pressp2>0? 0.01f: -0.2f);
        prs+=fab*(t1+t2);
      }
      const float p_vpm=-prs*massp2;
      a.ax+=p_vpm*frx; a.ay+=p_vpm*fry; a.az+=p_vpm*frz;
    }
    float dvx=fp.velrhop.x-velrhop2.x, dvy=fp.velrhop.y-velrhop2.y, dvz=fp.velrhop.z-velrhop2.z;
    a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhopp1*ov2);
    const float ovrr=FastRcp(rr2+c.eta2);
    const float cbar=P2.cs0;
    const float dot3=(drx*frx+dry*fry+drz*frz);
    const bool samephase=(fp.ph==pp2);
    if(TDDT==DSPH_DDT_DDT){
      if(boundflag)anyb=1u;
      if(!boundp2 && samephase){
        const float visc_densi=c.ddtkh*cbar*(__fdiv_rn(rhopp1,velrhop2.w)-1.f)*ovrr;
        a.delta+=visc_densi*dot3*massp2;
      }
    }
    if(TDDT==DSPH_DDT_DDT2 || TDDT==DSPH_DDT_DDT2FULL){
      if(boundflag && TDDT==DSPH_DDT_DDT2)anyb=1u;
      if(!boundp2 && samephase){
        const float rh=1.f+c.ddtgz*drz;
        const float drhop=c.rhopzero*(PowNearOne(rh,c)-1.f);
        const float visc_densi=c.ddtkh*cbar*((velrhop2.w-rhopp1)-drhop)*ovrr;
        a.delta-=visc_densi*dot3*massp2*ov2;
      }
    }
    if(SHIFT){
      if(boundflag && (c.shiftmode==DSPH_SHIFT_NOBOUND || (c.shiftmode==DSPH_SHIFT_NOFIXED && (codep2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FIXED)))nosh=1u;
      const bool heavyphase=(valid && fp.mass>P2.mass && !samephase);
      const float massrhop=massp2*ov2;
      a.shift.x=(heavyphase? 0.f: a.shift.x+massrhop*frx);
      if(!heavyphase){ a.shift.y+=massrhop*fry; a.shift.z+=massrhop*frz; a.shift.w-=massrhop*dot3; }
    }
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    a.visc=fmaxf(dot*ovrr,a.visc);
    if(TVISCO!=DSPH_VISCO_ARTIFICIAL){
      if(boundp2){ dvx=2.f*fp.velrhop.x; dvy=2.f*fp.velrhop.y; dvz=2.f*fp.velrhop.z; }
      const float volp2=-massp2*ov2;
      float dv=dvx*volp2; g0+=dv*frx; g1+=dv*fry; g2+=dv*frz;
            dv=dvy*volp2; g1+=dv*frx; g3+=dv*fry; g4+=dv*frz;
            dv=dvz*volp2; g2+=dv*frx; g4+=dv*fry; g5+=dv*frz;
    }
  });
  if(!ffdone)ffstore();
  float viscmax=0,vel2=0,etamax=0;
  if(active){
    float ar=a.ar;
    if(TDDT!=DSPH_DDT_NONE && !anyb)ar+=a.delta;
    t.acear[p1]=make_float4(a.ax,a.ay,a.az,ar);
    if(SHIFT){ float4 s=a.shift; if(nosh)s.x=FLT_MAX; t.shiftposfs[p1]=s; }
    const bool halo=(t.skipghosts && (code1&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC);   // x-slab halo copy: no part in the maxima
    if(!halo)viscmax=a.visc;
    vel2=fp.velrhop.x*fp.velrhop.x+fp.velrhop.y*fp.velrhop.y+fp.velrhop.z*fp.velrhop.z;
    if(TVISCO!=DSPH_VISCO_ARTIFICIAL){
      const float div_vel=(g0+g3+g5)/3.f;
      const float dxx=g0-div_vel, dxy=0.5f*g1, dxz=0.5f*g2, dyy=g3-div_vel, dyz=0.5f*g4, dzz=g5-div_vel;
      const float II_D=-(dxx*dyy+dyy*dzz+dxx*dzz)+(dxy*dxy+dyz*dyz+dxz*dxz);
      const float dmagn=sqrtf(II_D);
      const DevPhase &P=ph.ph[fp.ph];
      const float eta=NNEtaEffective(P,P.tau_yield,dmagn,P.visco,P.m_nn,P.n_nn);     // per particle: IEEE version
      t.viscoeta[p1]=eta; if(!halo)etamax=eta;
      if(TVISCO==DSPH_VISCO_CONSTEQ){
        const float t2=2.f*eta; float *tau=t.nntau+6*size_t(p1);
        tau[0]=t2*dxx; tau[1]=t2*dxy; tau[2]=t2*dxz; tau[3]=t2*dyy; tau[4]=t2*dyz; tau[5]=t2*dzz;
      }
    }
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
  WarpMaxToGlobal(etamax,&t.fmax->visceta_bits);
}

// pass 2: viscous forces (artificial / Morris with eta1+eta2 / divergence of tau1+tau2), added to acear
template<int TKER,int TVISCO> __global__ void __launch_bounds__(FORCES_THREADS,7)
ker_nn_sph_visco_ring(ForceParams t,DevCtes c,DevPhases ph)
{
  __shared__ __align__(16) V8Smem qsm;
  const unsigned q=V8Setup(qsm.w,c.scell);
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned p1c=min(p1,t.np-1);
  const bool active=(p1<t.np);
  const unsigned code1=t.code[p1c];
  const float4 pp1=t.posp[p1c];
  const float4 velrhop1=t.velrhop[p1c];
  unsigned ph1=code1&DSPH_CODE_MASKVALUE; if(ph1>=unsigned(ph.count))ph1=0;
  float eta1=0; float tau1[6]={0,0,0,0,0,0};
  if(TVISCO==DSPH_VISCO_LAMINARSPS)eta1=t.viscoeta[p1c];
  if(TVISCO==DSPH_VISCO_CONSTEQ){
    #pragma unroll
    for(int k=0;k<6;k++)tau1[k]=t.nntau[6*size_t(p1c)+k];
  }
  float ax=0,ay=0,az=0;
  RingTraverse(t,c,active,pp1,t.dcell[p1c],q,[&](unsigned p2,float drx,float dry,float drz,float rr2,float,bool boundp2,bool valid){
    const float4 velrhop2=__ldg(&t.velrhop[p2]);
    const unsigned pp2=(boundp2? ph1: (unsigned(t.code[p2])&DSPH_CODE_MASKVALUE));
    const DevPhase &P2=ph.ph[pp2];
    const float massp2=(valid? (boundp2? c.massbound: P2.mass): 0.f);
    float rinv,rad,qq; const float fac=NNFacFast<TKER>(rr2,c,rinv,rad,qq);
    const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
    const float ov2=FastRcp(velrhop2.w);
    if(TVISCO==DSPH_VISCO_CONSTEQ){
      float ts[6];
      #pragma unroll
      for(int k=0;k<6;k++)ts[k]=tau1[k]+(boundp2? tau1[k]: t.nntau[6*size_t(p2)+k]);
      const float m2=massp2*ov2;
      ax+=m2*(ts[0]*frx+ts[1]*fry+ts[2]*frz);
      ay+=m2*(ts[1]*frx+ts[3]*fry+ts[4]*frz);
      az+=m2*(ts[2]*frx+ts[4]*fry+ts[5]*frz);
    }
    else{
      float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
      if(boundp2){ dvx=2.f*velrhop1.x; dvy=2.f*velrhop1.y; dvz=2.f*velrhop1.z; }
      const float ovrr=FastRcp(rr2+c.eta2);
      if(TVISCO==DSPH_VISCO_ARTIFICIAL){
        const float dot=drx*dvx+dry*dvy+drz*dvz;
        if(dot<0){
          const float amubar=c.kernelh*(dot*ovrr);
          const float ovrobar=FastRcp((velrhop1.w+velrhop2.w)*0.5f);
          const float pi_visc=(-P2.visco*P2.cs0*amubar*ovrobar)*massp2;
          ax-=pi_visc*frx; ay-=pi_visc*fry; az-=pi_visc*frz;
        }
      }
      else{
        const float eta2=(boundp2? eta1: t.viscoeta[p2]);
        const float temp=(eta1+eta2)*ovrr*ov2;
        const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
        ax+=vtemp*dvx; ay+=vtemp*dvy; az+=vtemp*dvz;
      }
    }
  });
  float ace2=0;
  if(active){
    float4 acc=t.acear[p1];
    acc.x+=ax; acc.y+=ay; acc.z+=az;
    if(c.simulate2d)acc.y=0;
    t.acear[p1]=acc;
    const bool ghost=((code1&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC);
    if(!c.ghosts || !ghost)ace2=acc.x*acc.x+acc.y*acc.y+acc.z*acc.z;
  }
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
}

// ---------------------------------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------------------------------
template<int TKER,int TVISCO,int TDDT,bool SHIFT> static unsigned LaunchNN4(int variant,const ForceParams &t,const DevCtes &c,const DevPhases &ph,cudaStream_t stm){
  const unsigned npf=t.np-t.npb, nb=(npf+FORCES_THREADS-1)/FORCES_THREADS;
  if(ph.velgrad==DSPH_VELGRAD_FDA){
    // variant 1: plain IEEE loop (cross-check); otherwise the ring kernel (needs cells of 2h and np < 2^26 like forces_ring.cuh)
    if(variant!=1 && c.scelldiv==1 && t.np<=V5_IDXMASK)ker_nn_fda_ring<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c,ph);
    else ker_nn_fda<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c,ph);
    return 1;
  }
  if(variant!=1 && c.scelldiv==1 && t.np<=V5_IDXMASK){
    ker_nn_sph_pressgrad_ring<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c,ph);
    ker_nn_sph_visco_ring<TKER,TVISCO><<<nb,FORCES_THREADS,0,stm>>>(t,c,ph);
  }
  else{
    ker_nn_sph_pressgrad<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c,ph);
    ker_nn_sph_visco<TKER,TVISCO><<<nb,FORCES_THREADS,0,stm>>>(t,c,ph);
  }
  return 2;
}
template<int TKER,int TVISCO,int TDDT> static unsigned LaunchNN3(int variant,bool shift,const ForceParams &t,const DevCtes &c,const DevPhases &ph,cudaStream_t stm){
  return shift? LaunchNN4<TKER,TVISCO,TDDT,true>(variant,t,c,ph,stm): LaunchNN4<TKER,TVISCO,TDDT,false>(variant,t,c,ph,stm);
}
template<int TKER,int TVISCO> static unsigned LaunchNN2(int variant,int tddt,bool shift,const ForceParams &t,const DevCtes &c,const DevPhases &ph,cudaStream_t stm){
  switch(tddt){
    case DSPH_DDT_NONE:     return LaunchNN3<TKER,TVISCO,DSPH_DDT_NONE>(variant,shift,t,c,ph,stm);
    case DSPH_DDT_DDT:      return LaunchNN3<TKER,TVISCO,DSPH_DDT_DDT>(variant,shift,t,c,ph,stm);
    case DSPH_DDT_DDT2:     return LaunchNN3<TKER,TVISCO,DSPH_DDT_DDT2>(variant,shift,t,c,ph,stm);
    case DSPH_DDT_DDT2FULL: return LaunchNN3<TKER,TVISCO,DSPH_DDT_DDT2FULL>(variant,shift,t,c,ph,stm);
  }
  return 0;
}
template<int TKER> static unsigned LaunchNN1(int variant,int tvisco,int tddt,bool shift,const ForceParams &t,const DevCtes &c,const DevPhases &ph,cudaStream_t stm){
  switch(tvisco){
    case DSPH_VISCO_ARTIFICIAL: return LaunchNN2<TKER,DSPH_VISCO_ARTIFICIAL>(variant,tddt,shift,t,c,ph,stm);
    case DSPH_VISCO_LAMINARSPS: return LaunchNN2<TKER,DSPH_VISCO_LAMINARSPS>(variant,tddt,shift,t,c,ph,stm);
    case DSPH_VISCO_CONSTEQ:    return LaunchNN2<TKER,DSPH_VISCO_CONSTEQ>(variant,tddt,shift,t,c,ph,stm);
  }
  return 0;
}
// Interaction of the fluid particles for the multiphase fork (cusphNN::Interaction_ForcesNN, M/JSphGpu_NN_ker.h:31-41).
// Writes acear (+shiftposfs) for [npb,np) and the four maxima of t.fmax; returns the number of kernels launched.
unsigned LaunchForcesNN(int variant,int tkernel,int tvisco,int tdensity,bool shift,const ForceParams &t,const DevCtes &c,const DevPhases &ph,cudaStream_t stm){
  if(t.np<=t.npb)return 0;
  if(tkernel==DSPH_KERNEL_WENDLAND)return LaunchNN1<DSPH_KERNEL_WENDLAND>(variant,tvisco,tdensity,shift,t,c,ph,stm);
  if(tkernel==DSPH_KERNEL_CUBIC)return LaunchNN1<DSPH_KERNEL_CUBIC>(variant,tvisco,tdensity,shift,t,c,ph,stm);
  return 0;
}

} // namespace dsph
