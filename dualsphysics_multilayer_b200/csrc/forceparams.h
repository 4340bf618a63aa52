// forceparams.h -- arguments of the force / probe kernels and the neighbour-cell range helper, shared by the translation
// units that launch them (engine.cu, forces_fluid.cu, forces_nn.cu).
#pragma once
#include "devtypes.h"
#include "../../include/dsph_b200.h"
#include <cuda_runtime.h>

namespace dsph {

constexpr int FORCES_THREADS=128;

struct ForceParams {
  unsigned np, npb, npbok;
  int skipghosts;              // multi-GPU: CODE_PERIODIC fluid particles are halo copies, no forces needed
  const unsigned* __restrict__ ownlist;   // multi-GPU, variant 10: the owned fluid particles in index order (ker_mg_own_write), else null
  const unsigned* __restrict__ nown;      // their number (device)
  int ncx, ncy, ncz, nsheet;
  unsigned cellfluid;          // nct+1: offset of the fluid boxes in begincell
  unsigned cellzero[3];        // CellDomainMin
  const unsigned* __restrict__ begincell;
  const unsigned* __restrict__ dcell;
  const float4* __restrict__ posp;
  const uint4* __restrict__ posh;      // packed-half copy of posp for the candidate culling of forces_v10.cuh (one uint4 per 2 particles)
  const float4* __restrict__ velrhop;
  const typecode* __restrict__ code;
  const float* __restrict__ spstau;     // 6 floats per particle (laminar+SPS)
  float* __restrict__ spsgradvel;       // 6 floats per particle, OUT
  float4* __restrict__ shiftposfs;      // OUT (shifting)
  float4* __restrict__ acear;           // OUT
  ForceMax* __restrict__ fmax;
  float* __restrict__ viscoeta;         // multiphase, SPH velocity gradients: effective viscosity per particle
  float* __restrict__ nntau;            // multiphase, constitutive equation: stress tensor, 6 floats per particle
};

#ifdef __CUDACC__
struct NgRange { int cxini,cxfin,yini,yfin,zini,zfin,cx,cy,cz; };
__device__ __forceinline__ NgRange NgInit(unsigned rcell,const ForceParams &t,const DevCtes &c){
  // nsearch::Init S/JCellSearch_inline.h:33-52
  NgRange r;
  r.cx=int(DcCellx(c.domcellcode,rcell))-int(t.cellzero[0]);
  r.cy=int(DcCelly(c.domcellcode,rcell))-int(t.cellzero[1]);
  r.cz=int(DcCellz(c.domcellcode,rcell))-int(t.cellzero[2]);
  const int sd=c.scelldiv;
  r.cxini=r.cx-min(r.cx,sd); r.cxfin=r.cx+min(t.ncx-r.cx-1,sd)+1;
  r.yini=r.cy-min(r.cy,sd);  r.yfin=r.cy+min(t.ncy-r.cy-1,sd)+1;
  r.zini=r.cz-min(r.cz,sd);  r.zfin=r.cz+min(t.ncz-r.cz-1,sd)+1;
  return r;
}

#endif

} // namespace dsph
