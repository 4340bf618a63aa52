// forces_v2.cuh -- tuned Interaction_Forces variants (see DESIGN.md "force kernel").
#pragma once
#include "forces.cuh"
namespace dsph {
template<int TKER,int TVISCO,int TDDT,bool SHIFT> static void LaunchFluidV2(unsigned npf,const ForceParams &t,const DevCtes &c,cudaStream_t stm){
  ker_forces_fluid_v1<TKER,TVISCO,TDDT,SHIFT><<<(npf+FORCES_THREADS-1)/FORCES_THREADS,FORCES_THREADS,0,stm>>>(t,c);
}
} // namespace dsph
