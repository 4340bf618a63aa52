// forces.cuh -- Interaction_Forces for sm_100a.
//
// Replaces cusph::Interaction_Forces -> KerInteractionForcesFluid/Bound (+Box) (S/JSphGpu_ker.cu:420-885), the four
// cudaMemset + VelMod + max-reductions of JSphGpu::PreInteraction_Forces (S/JSphGpu.cpp:811-853) and the
// ReduMaxFloat/ComputeAceMod passes after it (S/JSphGpuSingle.cpp:480-514).  The physics of one pair follows the
// reference CPU kernel JSphCpu::InteractionForcesFluid / InteractionForcesBound (S/JSphCpu.cpp:548-822), which is the
// oracle; the pair visiting order is the oracle's (z rows, y rows, x cells, ascending particle) and the fluid-fluid
// and fluid-boundary contributions are accumulated separately and added in the oracle's order.
//
// Data layout (HBM, cell-sorted SoA):
//   posp[p]    float4  position relative to the origin of the particle's own cell (x,y,z) + pressure (w)
//   velrhop[p] float4  velocity + density
//   acear[p]   float4  OUT acceleration (x,y,z) + density rate ar (w)          (one 16 B store per particle)
// The neighbour of a particle in cell (cx,cy,cz) is visited cell by cell; for one neighbour cell the offset between
// the two cell origins is a warp-uniform constant, so a pair distance costs 3 FADD + 3 FFMA on cell-relative floats
// (no per-pair bit unpacking as in the reference's PosCell, S/JSphGpu_ker.cu:526-528).
// Pressure is precomputed per particle at reorder time (celldiv.cuh) instead of pow() per pair (S/JSphGpu_ker.cu:559).
// ViscDt / |ace|^2 / |vel|^2 maxima are reduced with warp shuffles and one atomicMax per warp on the float bit
// pattern (all values >= 0), so no reduction passes and no extra arrays.
#pragma once
#include "forceparams.h"
#include <cfloat>

namespace dsph {

__device__ __forceinline__ float PowNearOne(float rh,const DevCtes &c){
  // rh^(1/gamma) for the DDT2/3 hydrostatic term (S/JSphCpu.cpp:729-731).  rh=1+ddtgz*drz is within ~1e-2 of 1, where
  // powf() of the value loses nothing but the float result is dominated by its rounding; the series of
  // (1+t)^(1/gamma)-1 evaluated in float is exact to well below half an ulp of the result, i.e. it returns the
  // correctly rounded float power (what glibc powf returns for the oracle) at a cost of 6 FFMA.
  const float t=rh-1.f;
  if(fabsf(t)<0.03f){
    const float d=t*(c.powc[0]+t*(c.powc[1]+t*(c.powc[2]+t*(c.powc[3]+t*(c.powc[4]+t*c.powc[5])))));
    return 1.f+d;
  }
  return powf(rh,c.ovgamma);
}

template<int TKER> __device__ __forceinline__ float KernelFac(float rr2,const DevCtes &c){
  const float rad=sqrtf(rr2);
  const float qq=rad/c.kernelh;
  if(TKER==DSPH_KERNEL_WENDLAND){
    const float wqq1=1.f-0.5f*qq;
    return c.bwen*qq*wqq1*wqq1*wqq1/rad;
  }
  else{
    if(rad>c.kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; return c.cub_c2*wqq2/rad; }
    const float wqq2=qq*qq; return (c.cub_c1*qq+c.cub_d1*wqq2)/rad;
  }
}
__device__ __forceinline__ float CubicTensil(float rr2,float rhopp1,float pressp1,float rhopp2,float pressp2,const DevCtes &c){
  const float rad=sqrtf(rr2); const float qq=rad/c.kernelh;
  float wab;
  if(rad>c.kernelh){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; wab=c.cub_a24*(wqq2*wqq1); }
  else{ const float wqq2=qq*qq; wab=c.cub_a2*(1.0f+(0.75f*qq-1.5f)*wqq2); }
  float fab=wab*c.cub_od_wdeltap; fab*=fab; fab*=fab;
  const float t1=(pressp1/(rhopp1*rhopp1))*(pressp1>0? 0.01f: -0.2f);
  const float t2=(pressp2/(rhopp2*rhopp2))*(pressp2>0? 0.01f: -0.2f);
  return fab*(t1+t2);
}

struct PairAcc {
  float ax,ay,az,ar,delta,visc;
  float gxx,gxy,gxz,gyy,gyz,gzz;   // velocity gradient (laminar+SPS)
  float4 shift;
};

// One accepted pair; BOUNDP2 selects the fluid-boundary variant (oracle: second call of InteractionForcesFluid).
template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void PairFluid(
  float drx,float dry,float drz,float rr2,const float4 &velrhop1,float pressp1,const float4 &velrhop2,float pressp2,
  float massp2,float visco,const float* __restrict__ taup1,const float* __restrict__ taup2,unsigned codep2,const DevCtes &c,PairAcc &a)
{
  const float fac=KernelFac<TKER>(rr2,c);
  const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
  const float rhopp1=velrhop1.w;
  { //-Momentum
    const float prs=(pressp1+pressp2)/(rhopp1*velrhop2.w)+(TKER==DSPH_KERNEL_CUBIC? CubicTensil(rr2,rhopp1,pressp1,velrhop2.w,pressp2,c): 0.f);
    const float p_vpm=-prs*massp2;
    a.ax+=p_vpm*frx; a.ay+=p_vpm*fry; a.az+=p_vpm*frz;
  }
  const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
  a.ar+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhopp1/velrhop2.w);
  const float cbar=c.cs0;
  if(TDDT==DSPH_DDT_DDT){
    if(a.delta!=FLT_MAX){
      // the reference binary evaluates rho1*(1/rho2) for cubic + laminar/SPS and the true quotient otherwise (oracle/sph_oracle.cpp)
      const float rhop1over2=(TKER==DSPH_KERNEL_CUBIC && TVISCO==DSPH_VISCO_LAMINARSPS? __fmul_rn(rhopp1,__frcp_rn(velrhop2.w)): __fdiv_rn(rhopp1,velrhop2.w));
      const float visc_densi=c.ddtkh*cbar*(rhop1over2-1.f)/(rr2+c.eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      const float delta=visc_densi*dot3*massp2;
      a.delta=(BOUNDP2 && !c.mdbc? FLT_MAX: a.delta+delta);
    }
  }
  if(TDDT==DSPH_DDT_DDT2 || (TDDT==DSPH_DDT_DDT2FULL && !BOUNDP2)){
    if(a.delta!=FLT_MAX){
      const float rh=1.f+c.ddtgz*drz;
      const float drhop=c.rhopzero*(PowNearOne(rh,c)-1.f);   // the reference's fast-math build evaluates rho0*(pow-1)
      const float visc_densi=c.ddtkh*cbar*((velrhop2.w-rhopp1)-drhop)/(rr2+c.eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      const float delta=visc_densi*dot3*massp2/velrhop2.w;
      a.delta=(BOUNDP2? FLT_MAX: a.delta-delta);
    }
  }
  if(SHIFT){
    if(a.shift.x!=FLT_MAX){
      const float massrhop=massp2/velrhop2.w;
      const bool noshift=(BOUNDP2 && (c.shiftmode==DSPH_SHIFT_NOBOUND || (c.shiftmode==DSPH_SHIFT_NOFIXED && (codep2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FIXED)));
      a.shift.x=(noshift? FLT_MAX: a.shift.x+massrhop*frx);
      a.shift.y+=massrhop*fry; a.shift.z+=massrhop*frz;
      a.shift.w-=massrhop*(drx*frx+dry*fry+drz*frz);
    }
  }
  { //-Viscosity
    const float dot=drx*dvx+dry*dvy+drz*dvz;
    const float dot_rr2=dot/(rr2+c.eta2);
    a.visc=fmaxf(dot_rr2,a.visc);
    if(TVISCO==DSPH_VISCO_ARTIFICIAL){
      if(dot<0){
        const float amubar=c.kernelh*dot_rr2;
        const float robar=(rhopp1+velrhop2.w)*0.5f;
        const float pi_visc=(-visco*cbar*amubar/robar)*massp2;
        a.ax-=pi_visc*frx; a.ay-=pi_visc*fry; a.az-=pi_visc*frz;
      }
    }
    else{
      {
        const float robar2=(rhopp1+velrhop2.w);
        const float temp=4.f*visco/((rr2+c.eta2)*robar2);
        const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
        a.ax+=vtemp*dvx; a.ay+=vtemp*dvy; a.az+=vtemp*dvz;
      }
      float tau_xx=taup1[0],tau_xy=taup1[1],tau_xz=taup1[2],tau_yy=taup1[3],tau_yz=taup1[4],tau_zz=taup1[5];
      if(!BOUNDP2){ tau_xx+=taup2[0]; tau_xy+=taup2[1]; tau_xz+=taup2[2]; tau_yy+=taup2[3]; tau_yz+=taup2[4]; tau_zz+=taup2[5]; }
      a.ax+=massp2*(tau_xx*frx+tau_xy*fry+tau_xz*frz);
      a.ay+=massp2*(tau_xy*frx+tau_yy*fry+tau_yz*frz);
      a.az+=massp2*(tau_xz*frx+tau_yz*fry+tau_zz*frz);
      const float volp2=-massp2/velrhop2.w;
      float dv=dvx*volp2; a.gxx+=dv*frx; a.gxy+=dv*fry; a.gxz+=dv*frz;
            dv=dvy*volp2; a.gxy+=dv*frx; a.gyy+=dv*fry; a.gyz+=dv*frz;
            dv=dvz*volp2; a.gxz+=dv*frx; a.gyz+=dv*fry; a.gzz+=dv*frz;
    }
  }
}

__device__ __forceinline__ void WarpMaxToGlobal(float v,unsigned *dst){
  // v>=0: the uint bit pattern preserves the order
  unsigned b=__float_as_uint(v);
  b=__reduce_max_sync(0xffffffffu,b);
  if((threadIdx.x&31)==0 && b)atomicMax(dst,b);
}

// ---- v1: one thread per fluid particle, direct loop (reference structure; the tuned variants must reproduce it) ----
template<int TKER,int TVISCO,int TDDT,bool SHIFT,bool BOUNDP2> __device__ __forceinline__ void LoopFluid(
  const NgRange &ng,const float4 &pp1,const float4 &velrhop1,const float* taup1,float massp2,float visco,
  const ForceParams &t,const DevCtes &c,PairAcc &a)
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
          if(rr2<=c.kernelsize2 && rr2>=1e-18f){
            const float4 velrhop2=__ldg(&t.velrhop[p2]);
            unsigned codep2=0; if(SHIFT && BOUNDP2)codep2=t.code[p2];
            const float *taup2=(TVISCO==DSPH_VISCO_LAMINARSPS && !BOUNDP2? t.spstau+6*size_t(p2): nullptr);
            PairFluid<TKER,TVISCO,TDDT,SHIFT,BOUNDP2>(drx,dry,drz,rr2,velrhop1,pp1.w,velrhop2,pp2.w,massp2,visco,taup1,taup2,codep2,c,a);
          }
        }
        pini=pfin;
      }
    }
  }
}

__device__ __forceinline__ void ZeroAcc(PairAcc &a){
  a.ax=a.ay=a.az=a.ar=a.delta=a.visc=0; a.gxx=a.gxy=a.gxz=a.gyy=a.gyz=a.gzz=0; a.shift=make_float4(0,0,0,0);
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> __global__ void __launch_bounds__(FORCES_THREADS)
ker_forces_fluid_v1(ForceParams t,DevCtes c)
{
  const unsigned p1=t.npb+blockIdx.x*blockDim.x+threadIdx.x;
  float viscmax=0,ace2=0,vel2=0;
  if(p1<t.np){
    const float4 pp1=t.posp[p1];
    const float4 velrhop1=t.velrhop[p1];
    float taup1[6]={0,0,0,0,0,0};
    if(TVISCO==DSPH_VISCO_LAMINARSPS){
      #pragma unroll
      for(int k=0;k<6;k++)taup1[k]=t.spstau[6*size_t(p1)+k];
    }
    const NgRange ng=NgInit(t.dcell[p1],t,c);
    // fluid-fluid, then fluid-boundary, combined exactly as the two CPU passes do (S/JSphCpu.cpp:800-818)
    PairAcc ff; ZeroAcc(ff);
    LoopFluid<TKER,TVISCO,TDDT,SHIFT,false>(ng,pp1,velrhop1,taup1,c.massfluid,c.visco,t,c,ff);
    PairAcc fb; ZeroAcc(fb);
    if(SHIFT)fb.shift=ff.shift;       // shiftposfs[] is carried from the first pass into the second
    LoopFluid<TKER,TVISCO,TDDT,SHIFT,true>(ng,pp1,velrhop1,taup1,c.massbound,c.viscobound,t,c,fb);
    float ar=0,ax=0,ay=0,az=0,delta=0;
    float g[6]={0,0,0,0,0,0};
    if(SHIFT||ff.ar||ff.ax||ff.ay||ff.az||ff.visc){
      if(TDDT!=DSPH_DDT_NONE)delta=(ff.delta==FLT_MAX? FLT_MAX: 0.f+ff.delta);
      ar+=ff.ar; ax+=ff.ax; ay+=ff.ay; az+=ff.az;
      viscmax=fmaxf(viscmax,ff.visc);
      if(TVISCO==DSPH_VISCO_LAMINARSPS){ g[0]+=ff.gxx; g[1]+=ff.gxy; g[2]+=ff.gxz; g[3]+=ff.gyy; g[4]+=ff.gyz; g[5]+=ff.gzz; }
    }
    if(SHIFT||fb.ar||fb.ax||fb.ay||fb.az||fb.visc){
      if(TDDT!=DSPH_DDT_NONE)delta=(delta==FLT_MAX || fb.delta==FLT_MAX? FLT_MAX: delta+fb.delta);
      ar+=fb.ar; ax+=fb.ax; ay+=fb.ay; az+=fb.az;
      viscmax=fmaxf(viscmax,fb.visc);
      if(TVISCO==DSPH_VISCO_LAMINARSPS){ g[0]+=fb.gxx; g[1]+=fb.gxy; g[2]+=fb.gxz; g[3]+=fb.gyy; g[4]+=fb.gyz; g[5]+=fb.gzz; }
    }
    if(c.simulate2d)ay=0;
    if(TDDT!=DSPH_DDT_NONE && delta!=FLT_MAX)ar+=delta;          // S/JSphCpuSingle.cpp:553-559
    t.acear[p1]=make_float4(ax,ay,az,ar);
    if(SHIFT)t.shiftposfs[p1]=fb.shift;
    if(TVISCO==DSPH_VISCO_LAMINARSPS){
      #pragma unroll
      for(int k=0;k<6;k++)t.spsgradvel[6*size_t(p1)+k]=g[k];
    }
    // maxima: AceMax ignores periodic duplicates (S/JSphCpuSingle.cpp:583-636); VelMax takes every fluid particle
    const bool normal=(!c.ghosts || (t.code[p1]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
    if(normal)ace2=ax*ax+ay*ay+az*az;
    vel2=velrhop1.x*velrhop1.x+velrhop1.y*velrhop1.y+velrhop1.z*velrhop1.z;
  }
  WarpMaxToGlobal(viscmax,&t.fmax->viscdt_bits);
  WarpMaxToGlobal(ace2,&t.fmax->ace2_bits);
  WarpMaxToGlobal(vel2,&t.fmax->vel2_bits);
}

// ---- boundary particles: density rate + viscdt (S/JSphCpu.cpp:548-625) ----
template<int TKER> __global__ void __launch_bounds__(FORCES_THREADS) ker_forces_bound_v1(ForceParams t,DevCtes c)
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
              if(rr2<=c.kernelsize2 && rr2>=1e-18f){
                const float fac=KernelFac<TKER>(rr2,c);
                const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
                const float4 velrhop2=__ldg(&t.velrhop[p2]);
                const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
                arp1+=c.massfluid*(dvx*frx+dvy*fry+dvz*frz)*(velrhop1.w/velrhop2.w);
                const float dot=drx*dvx+dry*dvy+drz*dvz;
                const float dot_rr2=dot/(rr2+c.eta2);
                visc=fmaxf(dot_rr2,visc);
              }
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

// ---- SPS stress tensor from the velocity gradient (S/JSphCpu.cpp:930-955) ----
static __global__ void __launch_bounds__(256) ker_sps_tau(unsigned n,unsigned pini,const float4* __restrict__ velrhop,const float* __restrict__ gradvel,float* __restrict__ tau,DevCtes c)
{
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n)return;
  const size_t p=size_t(pini)+i;
  const float gxx=gradvel[6*p],gxy=gradvel[6*p+1],gxz=gradvel[6*p+2],gyy=gradvel[6*p+3],gyz=gradvel[6*p+4],gzz=gradvel[6*p+5];
  const float pow1=gxx*gxx+gyy*gyy+gzz*gzz;
  const float prr=pow1+pow1+gxy*gxy+gxz*gxz+gyz*gyz;
  const float visc_sps=c.spssmag*sqrtf(prr);
  const float div_u=gxx+gyy+gzz;
  const float sps_k=(2.0f/3.0f)*visc_sps*div_u;
  const float sps_blin=c.spsblin*prr;
  const float sumsps=-(sps_k+sps_blin);
  const float twovisc_sps=(visc_sps+visc_sps);
  const float one_rho2=1.0f/velrhop[p].w;
  tau[6*p]=one_rho2*(twovisc_sps*gxx+sumsps); tau[6*p+1]=one_rho2*(visc_sps*gxy); tau[6*p+2]=one_rho2*(visc_sps*gxz);
  tau[6*p+3]=one_rho2*(twovisc_sps*gyy+sumsps); tau[6*p+4]=one_rho2*(visc_sps*gyz); tau[6*p+5]=one_rho2*(twovisc_sps*gzz+sumsps);
}

// ---- pair counters (JDsPips, S/JDsPips_ker.cu:66): real and checked pairs per particle class ----
static __global__ void __launch_bounds__(FORCES_THREADS) ker_pips(ForceParams t,DevCtes c,unsigned pini1,unsigned n1,int fluid)
{
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  const unsigned p1=pini1+i;
  unsigned long long real=0,checked=0;
  if(i<n1){
    const float4 pp1=t.posp[p1];
    const NgRange ng=NgInit(t.dcell[p1],t,c);
    for(int pass=0;pass<(fluid? 2: 1);pass++){
      const unsigned cellinit=(pass==0? t.cellfluid: 0u);
      for(int z=ng.zini;z<ng.zfin;z++)for(int y=ng.yini;y<ng.yfin;y++){
        const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+cellinit;
        unsigned pini=t.begincell[v+ng.cxini];
        for(int x=ng.cxini;x<ng.cxfin;x++){
          const unsigned pfin=t.begincell[v+x+1];
          const float ox=pp1.x-float(x-ng.cx)*c.scell, oy=pp1.y-float(y-ng.cy)*c.scell, oz=pp1.z-float(z-ng.cz)*c.scell;
          checked+=pfin-pini;
          for(unsigned p2=pini;p2<pfin;p2++){
            const float4 pp2=__ldg(&t.posp[p2]);
            const float drx=ox-pp2.x, dry=oy-pp2.y, drz=oz-pp2.z;
            const float rr2=drx*drx+dry*dry+drz*drz;
            if(rr2<=c.kernelsize2 && rr2>=1e-18f)real++;
          }
          pini=pfin;
        }
      }
    }
  }
  for(int o=16;o;o>>=1){ real+=__shfl_down_sync(0xffffffffu,real,o); checked+=__shfl_down_sync(0xffffffffu,checked,o); }
  if((threadIdx.x&31)==0 && checked){
    atomicAdd(fluid? &t.fmax->pips_fluid_real: &t.fmax->pips_bound_real,real);
    atomicAdd(fluid? &t.fmax->pips_fluid_checked: &t.fmax->pips_bound_checked,checked);
  }
}

} // namespace dsph
