// measure.cuh -- flow probes on the device: pressure, velocity and kernel sum at arbitrary points, interpolated from the fluid
// particles of the current cell list with the SPH kernel and the Shepard correction.  A stand-in for the reference's MeasureTool
// post-processing (`-points <file> -vars:+press,+vel,+kcorr`, R/examples/main/01_DamBreak/xCaseDambreak_linux64_CPU.sh:76-85),
// which is distributed as a binary only; the interpolation is the one of the in-solver velocity gauge
// (JGaugeVelocity::CalculeCpuT, S/JDsGaugeItem.cpp:325-367: sum over fluid neighbours of W_ab * m/rho_b * f_b) divided by the
// kernel sum.  One thread per probe point; positions are the double-precision ones (no cell-relative trick needed here).
#pragma once
#include "forceparams.h"

namespace dsph {

struct MeasureArgs {
  unsigned npoints;
  const double* __restrict__ points;    // xyz packed
  const double* __restrict__ pos;       // xyz packed, cell-sorted
  float* __restrict__ press; float* __restrict__ vel; float* __restrict__ kcorr;   // [npoints], [3*npoints], [npoints]
};

template<int TKER> __device__ __forceinline__ float KernelWab(float rr2,const DevCtes &c){
  const float rad=sqrtf(rr2), qq=rad/c.kernelh;
  if(TKER==DSPH_KERNEL_WENDLAND){ const float wqq=2.f*qq+1.f, wqq1=1.f-0.5f*qq, wqq2=wqq1*wqq1; return c.awen*wqq*wqq2*wqq2; }   // S/FunSphKernel.h:212-222
  if(rad>c.kernelh){ const float wqq1=2.0f-qq; return c.cub_a24*(wqq1*wqq1*wqq1); }                                                // S/FunSphKernel.h:86-95
  return c.cub_a2*(1.0f+(0.75f*qq-1.5f)*(qq*qq));
}

template<int TKER,bool NN> __global__ void ker_measure_points(MeasureArgs m,ForceParams t,DevCtes c,DevPhases ph)
{
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=m.npoints)return;
  const double px=m.points[3*size_t(i)], py=m.points[3*size_t(i)+1], pz=m.points[3*size_t(i)+2];
  double sumw=0,sump=0,svx=0,svy=0,svz=0;
  // cell of the point inside the current cell domain (nsearch::Init for a position, S/JCellSearch_inline.h:54-70)
  const double dx=px-c.domposmin[0], dy=py-c.domposmin[1], dz=pz-c.domposmin[2];
  if(dx>=0 && dy>=0 && dz>=0){
    const int cx=int(dx/c.scell_d)-int(t.cellzero[0]), cy=int(dy/c.scell_d)-int(t.cellzero[1]), cz=int(dz/c.scell_d)-int(t.cellzero[2]);
    const int sd=c.scelldiv;
    if(cx>=-sd && cy>=-sd && cz>=-sd && cx<t.ncx+sd && cy<t.ncy+sd && cz<t.ncz+sd){
      const int xini=max(cx-sd,0), xfin=min(cx+sd,t.ncx-1)+1;
      for(int z=max(cz-sd,0);z<=min(cz+sd,t.ncz-1);z++)for(int y=max(cy-sd,0);y<=min(cy+sd,t.ncy-1);y++){
        if(xini>=xfin)continue;
        const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
        const unsigned pini=t.begincell[v+xini], pfin=t.begincell[v+xfin];
        for(unsigned p2=pini;p2<pfin;p2++){
          const float drx=float(px-m.pos[3*size_t(p2)]), dry=float(py-m.pos[3*size_t(p2)+1]), drz=float(pz-m.pos[3*size_t(p2)+2]);
          const float rr2=drx*drx+dry*dry+drz*drz;
          if(rr2<=c.kernelsize2){
            const float4 vr=t.velrhop[p2];
            float mass=c.massfluid;
            if(NN){ const unsigned k=unsigned(t.code[p2])&DSPH_CODE_MASKVALUE; mass=ph.ph[k<unsigned(ph.count)? k: 0].mass; }
            const double w=double(KernelWab<TKER>(rr2,c)*(mass/vr.w));
            sumw+=w; sump+=w*double(t.posp[p2].w); svx+=w*double(vr.x); svy+=w*double(vr.y); svz+=w*double(vr.z);
          }
        }
      }
    }
  }
  const double ov=(sumw>0? 1.0/sumw: 0.0);
  m.kcorr[i]=float(sumw); m.press[i]=float(sump*ov);
  m.vel[3*size_t(i)]=float(svx*ov); m.vel[3*size_t(i)+1]=float(svy*ov); m.vel[3*size_t(i)+2]=float(svz*ov);
}

} // namespace dsph
