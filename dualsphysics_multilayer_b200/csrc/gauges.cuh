// gauges.cuh -- in-solver gauges on the device: velocity at a point, surface water level along a line, maximum z of
// the fluid around a vertical line, pressure force of the fluid on a boundary block.  What they compute is what the reference's
// gauges compute (CPU: JGaugeVelocity/JGaugeSwl/JGaugeMaxZ/JGaugeForce::CalculeCpuT, S/JDsGaugeItem.cpp:325-367,541-607,736-776,
// 986-1041; its GPU twins: S/JDsGauge_ker.cu:50,168,250,311), on the double-precision positions and the CURRENT densities (the
// gauges run after the step update and before the next divide, S/JSphCpuSingle.cpp:1094-1098: the cell table is the one of the
// mid-step divide, the particles have moved since).
//
// Mapping: the reference runs ONE thread per measuring point (velocity, force: one per boundary particle; maxz: one thread in all;
// swl: one thread per line point).  Here ONE WARP works per point / boundary particle / line point: the lanes stride over the
// candidate particles of each row of cells (coalesced reads of pos / velrhop), partial sums are kept per lane in the precision the
// reference uses (double for velocity and mass, float for the force) and combined with shuffles.  All the gauges of a step are
// evaluated by at most five launches (every velocity point in one, every line point of every swl gauge in one, ...).
#pragma once
#include "forceparams.h"
#include <cfloat>

namespace dsph {

struct GaugeArgs {
  const double* __restrict__ pos;        // xyz packed, cell-sorted (current positions)
  double domposmax[3];
};

__device__ __forceinline__ double WarpSumD(double v){
  #pragma unroll
  for(int o=16;o>0;o>>=1)v+=__shfl_xor_sync(0xffffffffu,v,o);
  return v;
}
__device__ __forceinline__ float WarpSumF(float v){
  #pragma unroll
  for(int o=16;o>0;o>>=1)v+=__shfl_xor_sync(0xffffffffu,v,o);
  return v;
}

// range of cells around a position (nsearch::Init for a position, S/JCellSearch_inline.h:54-70), clipped to the cell domain
struct GaugeCells { int xini,xfin,yini,yfin,zini,zfin; bool any; };
__device__ __forceinline__ GaugeCells GaugeCellRange(double px,double py,double pz,const ForceParams &t,const DevCtes &c){
  GaugeCells g; g.any=false;
  const int cx=int((px-c.domposmin[0])/c.scell_d)-int(t.cellzero[0]);
  const int cy=int((py-c.domposmin[1])/c.scell_d)-int(t.cellzero[1]);
  const int cz=int((pz-c.domposmin[2])/c.scell_d)-int(t.cellzero[2]);
  const int sd=c.scelldiv;
  g.xini=max(cx-sd,0); g.xfin=min(cx+sd,t.ncx-1)+1;
  g.yini=max(cy-sd,0); g.yfin=min(cy+sd,t.ncy-1)+1;
  g.zini=max(cz-sd,0); g.zfin=min(cz+sd,t.ncz-1)+1;
  g.any=(g.xini<g.xfin && g.yini<g.yfin && g.zini<g.zfin);
  return g;
}

__device__ __forceinline__ float GaugeWab(float rr2,const DevCtes &c){            // GetKernelWendland_Wab, S/FunSphKernel.h:206-213
  const float rad=sqrtf(rr2), qq=rad/c.kernelh;
  const float wqq=qq+qq+1.f, wqq1=1.f-0.5f*qq, wqq2=wqq1*wqq1;
  return c.awen*wqq*wqq2*wqq2;
}
__device__ __forceinline__ float GaugeFac(float rr2,const DevCtes &c){            // GetKernelWendland_Fac, S/FunSphKernel.h:217-222
  const float rad=sqrtf(rr2), qq=rad/c.kernelh;
  const float wqq1=1.f-0.5f*qq;
  return c.bwen*qq*wqq1*wqq1*wqq1/rad;
}
// ComputePressMonaghan (S/FunSphEos.h:38-40) with the rounding points of the reference build (-ffast-math: rhop*(1/rhop0) instead of the
// division; power, -1 and *B evaluated in double and rounded once), the same rule as the pressure of the reorder (celldiv.cuh
// PressFromRhop).  It matters: the power is ~1 and 1 is subtracted from it, so one ulp of rho/rho0 is a relative pressure error of
// 2^-23/(rho/rho0-1) ~ 3e-5 for water -- and float(1/1000) is 4.7e-8 too large, a bias every particle shares (measured on the tank
// force of the 3 k dam break, relative to rho0*g*H0*A: 1.7e-5 with the true division, 4e-6 with this rule).
__device__ __forceinline__ float GaugePress(float rhop,const DevCtes &c){
  const float x=__fmul_rn(rhop,c.ovrhopzero);
  return float(c.cteb_d*(pow(double(x),c.gamma_d)-1.0));
}
__device__ __forceinline__ bool GaugeIsFluid(typecode code){ return (code&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID; }

// ---- velocity (mode 0: out[4*i..] = sum w*vx, sum w*vy, sum w*vz, sum w) and mass (mode 1: out[i] = sum w*massf) at points.
// One warp per point.  Points: explicit list (points!=null) or line points p0+dir*cp of the swl gauges (lines[]: 8 doubles per gauge =
// p0, dir, first output index, number of points).
struct GaugeLine { double p0[3], dir[3]; unsigned first, count; };
template<int MODE> __global__ void __launch_bounds__(128) ker_gauge_points(unsigned npoints,const double* __restrict__ points,
  const GaugeLine* __restrict__ lines,unsigned nlines,float* __restrict__ out,GaugeArgs g,ForceParams t,DevCtes c)
{
  const unsigned w=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31;
  if(w>=npoints)return;
  double px,py,pz;
  if(points){ px=points[3*size_t(w)]; py=points[3*size_t(w)+1]; pz=points[3*size_t(w)+2]; }
  else{
    unsigned l=0; while(l+1<nlines && w>=lines[l+1].first)l++;
    const double cp=double(w-lines[l].first);
    px=lines[l].p0[0]+lines[l].dir[0]*cp; py=lines[l].p0[1]+lines[l].dir[1]*cp; pz=lines[l].p0[2]+lines[l].dir[2]*cp;
  }
  double s0=0,s1=0,s2=0,s3=0;
  // velocity gauges return zero for points out of the domain (PointIsOut, S/JDsGaugeItem.h:125); the mass of a line point is computed anyway
  const bool out_=(MODE==0 && (px!=px || py!=py || pz!=pz || px<c.domposmin[0] || py<c.domposmin[1] || pz<c.domposmin[2] || px>=g.domposmax[0] || py>=g.domposmax[1] || pz>=g.domposmax[2]));
  const GaugeCells gc=GaugeCellRange(px,py,pz,t,c);
  if(!out_ && gc.any)for(int z=gc.zini;z<gc.zfin;z++)for(int y=gc.yini;y<gc.yfin;y++){
    const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
    const unsigned pini=t.begincell[v+gc.xini], pfin=t.begincell[v+gc.xfin];
    for(unsigned p2=pini+lane;p2<pfin;p2+=32){
      const float drx=float(px-g.pos[3*size_t(p2)]), dry=float(py-g.pos[3*size_t(p2)+1]), drz=float(pz-g.pos[3*size_t(p2)+2]);
      const float rr2=drx*drx+dry*dry+drz*drz;
      // x-slab runs: every rank sums over the particles it owns (halo copies are owned elsewhere), the partial sums are all-reduced
      if(rr2<=c.kernelsize2 && rr2>=1e-18f && GaugeIsFluid(t.code[p2]) && (!t.skipghosts || (t.code[p2]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL)){
        const float4 vr=t.velrhop[p2];
        float wab=GaugeWab(rr2,c);
        wab*=c.massfluid/vr.w;
        if(MODE==0){ s3+=wab; s0+=wab*vr.x; s1+=wab*vr.y; s2+=wab*vr.z; }
        else s0+=wab*c.massfluid;
      }
    }
  }
  s0=WarpSumD(s0);
  if(MODE==0){ s1=WarpSumD(s1); s2=WarpSumD(s2); s3=WarpSumD(s3); }
  if(!lane){
    if(MODE==0){ out[4*size_t(w)]=float(s0); out[4*size_t(w)+1]=float(s1); out[4*size_t(w)+2]=float(s2); out[4*size_t(w)+3]=float(s3); }
    else out[w]=float(s0);
  }
}

// ---- swl: first change from fluid to empty along the line (JGaugeSwl::CalculeCpuT, S/JDsGaugeItem.cpp:575-600), one thread per gauge
__global__ void ker_gauge_swl_scan(unsigned nlines,const GaugeLine* __restrict__ lines,const float* __restrict__ mass,const float* __restrict__ masslimit,float* __restrict__ res)
{
  const unsigned l=blockIdx.x*blockDim.x+threadIdx.x;
  if(l>=nlines)return;
  const GaugeLine ln=lines[l];
  const float lim=masslimit[l];
  const unsigned pointnp=ln.count-1;                   // points cp = 0..PointNp
  float mpre=0; bool found=false; double sx=0,sy=0,sz=0;
  for(unsigned cp=0;cp<=pointnp && !found;cp++){
    const float m=mass[ln.first+cp];
    if(m>lim)mpre=m;
    if(m<lim && mpre!=0.f){
      const float fxm1=(lim-mpre)/(m-mpre)-1.f;
      const double f=double(cp)+double(fxm1);          // ptpos(cp) + dir*fxm1
      sx=ln.p0[0]+ln.dir[0]*double(cp)+ln.dir[0]*double(fxm1); sy=ln.p0[1]+ln.dir[1]*double(cp)+ln.dir[1]*double(fxm1); sz=ln.p0[2]+ln.dir[2]*double(cp)+ln.dir[2]*double(fxm1);
      (void)f; found=true;
    }
  }
  if(!found){ const double k=(mpre!=0.f? double(pointnp): 0.0); sx=ln.p0[0]+ln.dir[0]*k; sy=ln.p0[1]+ln.dir[1]*k; sz=ln.p0[2]+ln.dir[2]*k; }
  res[3*size_t(l)]=float(sx); res[3*size_t(l)+1]=float(sy); res[3*size_t(l)+2]=float(sz);
}

// ---- maxz: highest fluid particle within distlimit of a vertical line, searched layer of cells by layer of cells from the top; the
// first layer with a hit ends the search (JGaugeMaxZ::CalculeCpu, S/JDsGaugeItem.cpp:736-776).  One warp per gauge.
struct GaugeMaxzDef { double p0[3]; double height; float distlimit; float pad; };
__global__ void __launch_bounds__(128) ker_gauge_maxz(unsigned n,const GaugeMaxzDef* __restrict__ defs,float* __restrict__ res,GaugeArgs g,ForceParams t,DevCtes c)
{
  const unsigned w=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31;
  if(w>=n)return;
  const GaugeMaxzDef d=defs[w];
  const float maxdist2=d.distlimit*d.distlimit;
  // GetInteractionCellsMaxZ (S/JDsGaugeItem.cpp:715-731)
  const int cx=int((d.p0[0]-c.domposmin[0])/c.scell)-int(t.cellzero[0]);
  const int cy=int((d.p0[1]-c.domposmin[1])/c.scell)-int(t.cellzero[1]);
  const int cz=int((d.p0[2]-c.domposmin[2])/c.scell)-int(t.cellzero[2]);
  const int cz2=int((d.p0[2]+d.height-c.domposmin[2])/c.scell)-int(t.cellzero[2]);
  const int ncel=int(ceilf(d.distlimit/c.scell));
  const int cxini=max(cx-ncel,0), cxfin=min(cx+ncel+1,t.ncx), yini=max(cy-ncel,0), yfin=min(cy+ncel+1,t.ncy), zini=max(cz-ncel,0), zfin=min(cz2+ncel+1,t.ncz);
  float zmax=-FLT_MAX; bool hit=false;
  if(cxini<cxfin)for(int z=zfin-1;z>=zini && !hit;z--){
    for(int y=yini;y<yfin;y++){
      const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
      const unsigned pini=t.begincell[v+cxini], pfin=t.begincell[v+cxfin];
      for(unsigned p2=pini+lane;p2<pfin;p2+=32){
        const double pz2=g.pos[3*size_t(p2)+2];
        if(pz2>double(zmax)){
          const float drx=float(d.p0[0]-g.pos[3*size_t(p2)]), dry=float(d.p0[1]-g.pos[3*size_t(p2)+1]);
          if(drx*drx+dry*dry<=maxdist2 && GaugeIsFluid(t.code[p2]))zmax=float(pz2);
        }
      }
    }
    zmax=fmaxf(zmax,__shfl_xor_sync(0xffffffffu,zmax,16)); zmax=fmaxf(zmax,__shfl_xor_sync(0xffffffffu,zmax,8));
    zmax=fmaxf(zmax,__shfl_xor_sync(0xffffffffu,zmax,4));  zmax=fmaxf(zmax,__shfl_xor_sync(0xffffffffu,zmax,2));
    zmax=fmaxf(zmax,__shfl_xor_sync(0xffffffffu,zmax,1));
    hit=(zmax>-FLT_MAX);
  }
  if(!lane)res[w]=fmaxf(zmax,float(c.domposmin[2]));
}

// ---- force: pressure term of the momentum equation on the boundary particles of one mk block, fluid neighbours only, no viscosity
// (JGaugeForce::CalculeCpuT, S/JDsGaugeItem.cpp:986-1041).  One warp per boundary particle; partace indexed by idp-idbegin.
__global__ void __launch_bounds__(128) ker_gauge_force(unsigned npbok,typecode codesel,unsigned idbegin,unsigned count,const unsigned* __restrict__ idp,
  float* __restrict__ partace,GaugeArgs g,ForceParams t,DevCtes c)
{
  const unsigned p1=(blockIdx.x*blockDim.x+threadIdx.x)>>5, lane=threadIdx.x&31;
  if(p1>=npbok)return;
  const typecode code1=t.code[p1];
  if((code1&(DSPH_CODE_MASKTYPE|DSPH_CODE_MASKVALUE))!=codesel || (code1&DSPH_CODE_MASKSPECIAL)!=DSPH_CODE_NORMAL)return;
  const double px=g.pos[3*size_t(p1)], py=g.pos[3*size_t(p1)+1], pz=g.pos[3*size_t(p1)+2];
  const float rhop1=t.velrhop[p1].w;
  const float press1=GaugePress(rhop1,c);
  float ax=0,ay=0,az=0;
  const GaugeCells gc=GaugeCellRange(px,py,pz,t,c);
  if(gc.any)for(int z=gc.zini;z<gc.zfin;z++)for(int y=gc.yini;y<gc.yfin;y++){
    const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
    const unsigned pini=t.begincell[v+gc.xini], pfin=t.begincell[v+gc.xfin];
    for(unsigned p2=pini+lane;p2<pfin;p2+=32){
      const float drx=float(px-g.pos[3*size_t(p2)]), dry=float(py-g.pos[3*size_t(p2)+1]), drz=float(pz-g.pos[3*size_t(p2)+2]);
      const float rr2=drx*drx+dry*dry+drz*drz;
      if(rr2<=c.kernelsize2 && rr2>=1e-18f && GaugeIsFluid(t.code[p2])){
        const float fac=GaugeFac(rr2,c);
        const float rhop2=t.velrhop[p2].w;
        const float press2=GaugePress(rhop2,c);
        const float prs=(press1+press2)/(rhop1*rhop2);
        const float pv=-prs*c.massfluid;
        ax+=pv*(fac*drx); ay+=pv*(fac*dry); az+=pv*(fac*drz);
      }
    }
  }
  ax=WarpSumF(ax); ay=WarpSumF(ay); az=WarpSumF(az);
  const unsigned k=idp[p1]-idbegin;
  if(!lane && k<count){ partace[3*size_t(k)]=ax; partace[3*size_t(k)+1]=ay; partace[3*size_t(k)+2]=az; }
}

// sum of partace over the block in double (one block, fixed order of the partial sums: deterministic)
__global__ void __launch_bounds__(256) ker_gauge_force_sum(unsigned count,const float* __restrict__ partace,float massbound,float* __restrict__ res)
{
  __shared__ double sh[3][8];
  double sx=0,sy=0,sz=0;
  for(unsigned k=threadIdx.x;k<count;k+=256){ sx+=double(partace[3*size_t(k)]); sy+=double(partace[3*size_t(k)+1]); sz+=double(partace[3*size_t(k)+2]); }
  sx=WarpSumD(sx); sy=WarpSumD(sy); sz=WarpSumD(sz);
  if((threadIdx.x&31)==0){ sh[0][threadIdx.x>>5]=sx; sh[1][threadIdx.x>>5]=sy; sh[2][threadIdx.x>>5]=sz; }
  __syncthreads();
  if(threadIdx.x==0){
    double a=0,b=0,d=0;
    for(int i=0;i<8;i++){ a+=sh[0][i]; b+=sh[1][i]; d+=sh[2][i]; }
    res[0]=float(a)*massbound; res[1]=float(b)*massbound; res[2]=float(d)*massbound;
  }
}

} // namespace dsph
