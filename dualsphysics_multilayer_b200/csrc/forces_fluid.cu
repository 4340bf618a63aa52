// forces_fluid.cu -- translation unit of the single-phase Interaction_Forces kernels and their dispatch
// (cusph::Interaction_Forces dispatcher, S/JSphGpu_ker.cu:788-885: template selection by kernel, viscosity, DDT, shifting).
#include <algorithm>
#include "forces_api.h"
#include "forces.cuh"
#include "forces_fast.cuh"
#include "forces_ring.cuh"
#include "forces_v10.cuh"

namespace dsph {

static __global__ void ker_ringbase_probe(unsigned *out){
  __shared__ __align__(16) V8Smem qsm;
  if(threadIdx.x==0)*out=unsigned(__cvta_generic_to_shared(qsm.w));
}
void LaunchRingBaseProbe(unsigned *out_device,cudaStream_t stm){ ker_ringbase_probe<<<1,32,0,stm>>>(out_device); }
unsigned RingBaseExpected(){ return V5_SMEM0; }

int EffectiveFluidVariant(int variant,bool ringbase_ok,const ForceParams &t,const DevCtes &c){
  if(variant==1 || variant==4)return variant;
  const bool ringok=(ringbase_ok && c.scelldiv==1 && size_t(t.np)+16u<=size_t(V5_IDXMASK));
  if(!ringok)return 4;
#ifdef DSPH_V10_EXPERIMENTS
  if(variant>=11 && variant<=15)return variant;
#endif
  return (variant==9? 9: 10);
}

template<int TKER,int TVISCO,int TDDT,bool SHIFT> static void LaunchFluid(int variant,unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  const unsigned nb=(npf+FORCES_THREADS-1)/FORCES_THREADS;
  switch(variant){
    case 1:  ker_forces_fluid_v1<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c); break;
    case 4:  ker_forces_fluid_v4<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c); break;
    case 9:  ker_forces_fluid_v9<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c); break;
#ifdef DSPH_V10_EXPERIMENTS   // measured alternatives of the production kernel (profiles/r02_forces.md), production instantiation only
    case 11: if(TKER==2&&TVISCO==1&&TDDT==2&&!SHIFT){ ker_forces_fluid_v10<2,1,2,false,7><<<nb,FORCES_THREADS,0,stm>>>(t,c); break; }
    case 12: if(TKER==2&&TVISCO==1&&TDDT==2&&!SHIFT){ ker_forces_fluid_v10<2,1,2,false,8><<<nb,FORCES_THREADS,0,stm>>>(t,c); break; }
    case 13: if(TKER==2&&TVISCO==1&&TDDT==2&&!SHIFT){ ker_forces_fluid_v10<2,1,2,false,6,32><<<nb,FORCES_THREADS,0,stm>>>(t,c); break; }
    case 14: if(TKER==2&&TVISCO==1&&TDDT==2&&!SHIFT){ ker_forces_fluid_v10<2,1,2,false,8,8><<<nb,FORCES_THREADS,0,stm>>>(t,c); break; }
    case 15: if(TKER==2&&TVISCO==1&&TDDT==2&&!SHIFT){ ker_forces_fluid_v10<2,1,2,false,7,8><<<nb,FORCES_THREADS,0,stm>>>(t,c); break; }
#endif
    default: ker_forces_fluid_v10<TKER,TVISCO,TDDT,SHIFT><<<nb,FORCES_THREADS,0,stm>>>(t,c); break;
  }
}
template<int TKER,int TVISCO,int TDDT> static void LaunchFluid3(int variant,bool shift,unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  if(shift)LaunchFluid<TKER,TVISCO,TDDT,true>(variant,npf,t,c,stm); else LaunchFluid<TKER,TVISCO,TDDT,false>(variant,npf,t,c,stm);
}
template<int TKER,int TVISCO> static bool LaunchFluid2(int variant,int tddt,bool shift,unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  switch(tddt){
    case DSPH_DDT_NONE:     LaunchFluid3<TKER,TVISCO,DSPH_DDT_NONE>(variant,shift,npf,t,c,stm); return true;
    case DSPH_DDT_DDT:      LaunchFluid3<TKER,TVISCO,DSPH_DDT_DDT>(variant,shift,npf,t,c,stm); return true;
    case DSPH_DDT_DDT2:     LaunchFluid3<TKER,TVISCO,DSPH_DDT_DDT2>(variant,shift,npf,t,c,stm); return true;
    case DSPH_DDT_DDT2FULL: LaunchFluid3<TKER,TVISCO,DSPH_DDT_DDT2FULL>(variant,shift,npf,t,c,stm); return true;
  }
  return false;
}
template<int TKER> static bool LaunchFluid1(int variant,int tvisco,int tddt,bool shift,unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  if(tvisco==DSPH_VISCO_ARTIFICIAL)return LaunchFluid2<TKER,DSPH_VISCO_ARTIFICIAL>(variant,tddt,shift,npf,t,c,stm);
  if(tvisco==DSPH_VISCO_LAMINARSPS)return LaunchFluid2<TKER,DSPH_VISCO_LAMINARSPS>(variant,tddt,shift,npf,t,c,stm);
  return false;
}

int LaunchForcesFluid(int variant,bool ringbase_ok,int tkernel,int tvisco,int tddt,bool shift,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  if(t.np<=t.npb)return 0;
  const int v=EffectiveFluidVariant(variant,ringbase_ok,t,c);
  const unsigned npf=t.np-t.npb;
  bool ok=false;
  if(tkernel==DSPH_KERNEL_WENDLAND)ok=LaunchFluid1<DSPH_KERNEL_WENDLAND>(v,tvisco,tddt,shift,npf,t,c,stm);
  else if(tkernel==DSPH_KERNEL_CUBIC)ok=LaunchFluid1<DSPH_KERNEL_CUBIC>(v,tvisco,tddt,shift,npf,t,c,stm);
  return ok? v: 0;
}

void LaunchForcesBound(int variant,int tkernel,int numsm,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  if(!t.npb)return;
  if(variant==1){                                           // IEEE cross-check: one thread per particle
    const unsigned nb=(t.npb+FORCES_THREADS-1)/FORCES_THREADS;
    if(tkernel==DSPH_KERNEL_WENDLAND)ker_forces_bound_v1<DSPH_KERNEL_WENDLAND><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    else ker_forces_bound_v1<DSPH_KERNEL_CUBIC><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    return;
  }
  if(t.npb>=BOUND_WARP_LIMIT){                              // enough particles to fill the GPU with one thread each
    const unsigned nb=(t.npb+FORCES_THREADS-1)/FORCES_THREADS;
    if(tkernel==DSPH_KERNEL_WENDLAND)ker_forces_bound_v3<DSPH_KERNEL_WENDLAND><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    else ker_forces_bound_v3<DSPH_KERNEL_CUBIC><<<nb,FORCES_THREADS,0,stm>>>(t,c);
    return;
  }
  // small and medium cases: one warp per particle, grid-stride; at most 16 CTAs per SM
  const unsigned need=(t.npb+(FORCES_THREADS/32)-1)/(FORCES_THREADS/32);
  const unsigned nb=std::min(need,unsigned(numsm)*16u);
  if(tkernel==DSPH_KERNEL_WENDLAND)ker_forces_bound_v2<DSPH_KERNEL_WENDLAND><<<nb,FORCES_THREADS,0,stm>>>(t,c);
  else ker_forces_bound_v2<DSPH_KERNEL_CUBIC><<<nb,FORCES_THREADS,0,stm>>>(t,c);
}
void LaunchSpsTau(unsigned npf,unsigned npb,const float4 *velrhop,const float *gradvel,float *tau,const DevCtes &c,cudaStream_t stm){
  if(npf)ker_sps_tau<<<(npf+255)/256,256,0,stm>>>(npf,npb,velrhop,gradvel,tau,c);
}
void LaunchPips(const ForceParams &t,const DevCtes &c,unsigned pini,unsigned n,int fluid,cudaStream_t stm){
  if(n)ker_pips<<<(n+FORCES_THREADS-1)/FORCES_THREADS,FORCES_THREADS,0,stm>>>(t,c,pini,n,fluid);
}

} // namespace dsph
