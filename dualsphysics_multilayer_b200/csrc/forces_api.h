// forces_api.h -- host-side launch functions of the force kernels (defined in forces_fluid.cu / forces_nn.cu, called by engine.cu).
// The kernels live in their own translation units so that the 100+ template instantiations compile in parallel.
#pragma once
#include "forceparams.h"

namespace dsph {

// Single-phase fluid particles [npb,np): variant 1 = plain IEEE loop in the reference's structure and pair order (cross-check),
// 4 = same loop with the division-free pair body (any ScellDiv), 9 = deferred pair ring (forces_ring.cuh), 10 = packed-half
// culling + hit-mask ring (forces_v10.cuh, production).  9/10 need ScellDiv=1, np+16 < 2^26 and the ring at the start of the
// shared window (ringbase_ok), otherwise 4 runs.  Returns the variant that was launched, 0 for invalid options.
int LaunchForcesFluid(int variant,bool ringbase_ok,int tkernel,int tvisco,int tddt,bool shift,const ForceParams &t,const DevCtes &c,cudaStream_t stm);
int EffectiveFluidVariant(int variant,bool ringbase_ok,const ForceParams &t,const DevCtes &c);
// Boundary particles [0,npb): density rate + viscdt (KerInteractionForcesBound S/JSphGpu_ker.cu:420-506)
void LaunchForcesBound(int variant,int tkernel,int numsm,const ForceParams &t,const DevCtes &c,cudaStream_t stm);   // variant 1: IEEE thread-per-particle loop, else warp per particle
// SPS stress tensor from the velocity gradient (ComputeSpsTau S/JSphGpu_ker.cu:1516-1558)
void LaunchSpsTau(unsigned npf,unsigned npb,const float4 *velrhop,const float *gradvel,float *tau,const DevCtes &c,cudaStream_t stm);
// pair counters (JDsPips S/JDsPips_ker.cu:66) accumulated into t.fmax->pips_*
void LaunchPips(const ForceParams &t,const DevCtes &c,unsigned pini,unsigned n,int fluid,cudaStream_t stm);
// where the first static shared array of a kernel starts in the CTA's shared window (written to *out_device)
void LaunchRingBaseProbe(unsigned *out_device,cudaStream_t stm);
unsigned RingBaseExpected();
// multiphase non-Newtonian fork (cusphNN::Interaction_ForcesNN, M/JSphGpu_NN_ker.h:31-41); returns the number of kernels launched
unsigned LaunchForcesNN(int variant,int tkernel,int tvisco,int tdensity,bool shift,const ForceParams &t,const DevCtes &c,const DevPhases &ph,cudaStream_t stm);

} // namespace dsph
