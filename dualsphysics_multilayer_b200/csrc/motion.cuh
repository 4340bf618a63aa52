// motion.cuh -- predefined motion of moving boundary particles.
//
// Replaces cusph::CalcRidp / MoveLinBound / MoveMatBound (S/JSphGpu_ker.cu, called by JSphGpu::RunMotion S/JSphGpu.cpp:1035-1075);
// arithmetic follows the oracle JSphCpu::CalcRidp / MoveLinBound / MoveMatBound (S/JSphCpu.cpp:1660-1730): the displacement goes
// through the same UpdatePos as the fluid update (periodic wrap, out-of-domain marking, new cell), the velocity of a matrix motion
// is float(displacement/dt), and with mDBC the normal is carried as the image of the ghost node minus the image of the particle.
#pragma once
#include "step.cuh"

namespace dsph {

// position in the sorted arrays of every moving particle id [idini,idfin) (UINT_MAX when it is not on this device or excluded);
// only CODE_NORMAL particles count when periodic / halo copies exist (S/JSphCpu.cpp:1667-1677)
__global__ void ker_motion_ridp(unsigned npb,unsigned idini,unsigned idfin,int checknormal,const unsigned* __restrict__ idp,const typecode* __restrict__ code,unsigned* __restrict__ ridp){
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=npb)return;
  const unsigned id=idp[p];
  if(id>=idini && id<idfin && (!checknormal || (code[p]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL))ridp[id-idini]=p;
}

struct MotionArrays { double *pos; unsigned *dcell; float4 *velrhop; typecode *code; float4 *boundnormal; };

// MoveLinBound: the same displacement and velocity for the `count` ids from `ini` (relative to the first moving id)
__global__ void ker_motion_linear(unsigned count,unsigned ini,double mx,double my,double mz,float vx,float vy,float vz,const unsigned* __restrict__ ridp,MotionArrays a,DevCtes c){
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=count)return;
  const unsigned pid=ridp[ini+i];
  if(pid==0xffffffffu)return;
  UpdatePos(a.pos[3*size_t(pid)],a.pos[3*size_t(pid)+1],a.pos[3*size_t(pid)+2],mx,my,mz,false,pid,a.pos,a.dcell,a.code,c);
  const float rho=a.velrhop[pid].w;
  a.velrhop[pid]=make_float4(vx,vy,vz,rho);
}

struct Mat4d { double a[16]; };   // row-major a11..a44 (tmatrix4d)
__device__ __forceinline__ void MatMulPoint(const Mat4d &m,double x,double y,double z,double &rx,double &ry,double &rz){
  // separate roundings (no fused multiply-add) in the association of the reference's -ffast-math CPU binary, (a*x+b*y)+(c*z+d):
  // moved positions equal its dumps bit for bit (tests: mv_* goldens)
  rx=__dadd_rn(__dadd_rn(__dmul_rn(m.a[0],x),__dmul_rn(m.a[1],y)),__dadd_rn(__dmul_rn(m.a[2],z),m.a[3]));
  ry=__dadd_rn(__dadd_rn(__dmul_rn(m.a[4],x),__dmul_rn(m.a[5],y)),__dadd_rn(__dmul_rn(m.a[6],z),m.a[7]));
  rz=__dadd_rn(__dadd_rn(__dmul_rn(m.a[8],x),__dmul_rn(m.a[9],y)),__dadd_rn(__dmul_rn(m.a[10],z),m.a[11]));
}
// MoveMatBound: rigid transformation of the positions, velocity = displacement / dt
__global__ void ker_motion_matrix(unsigned count,unsigned ini,Mat4d m,double dt,const unsigned* __restrict__ ridp,MotionArrays a,DevCtes c){
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=count)return;
  const unsigned pid=ridp[ini+i];
  if(pid==0xffffffffu)return;
  const double px=a.pos[3*size_t(pid)], py=a.pos[3*size_t(pid)+1], pz=a.pos[3*size_t(pid)+2];
  double qx,qy,qz; MatMulPoint(m,px,py,pz,qx,qy,qz);
  if(c.simulate2d)qy=py;
  const double dx=qx-px, dy=qy-py, dz=qz-pz;
  UpdatePos(px,py,pz,dx,dy,dz,false,pid,a.pos,a.dcell,a.code,c);
  const float rho=a.velrhop[pid].w;
  a.velrhop[pid]=make_float4(float(dx/dt),float(dy/dt),float(dz/dt),rho);
  if(a.boundnormal){
    const float4 n=a.boundnormal[pid];
    double gx,gy,gz; MatMulPoint(m,px+double(n.x),py+double(n.y),pz+double(n.z),gx,gy,gz);
    a.boundnormal[pid]=make_float4(float(gx-qx),float(gy-qy),float(gz-qz),n.w);
  }
}

} // namespace dsph
