// step.cuh -- ComputeStep (Verlet, Symplectic predictor/corrector) fused with the position update, out-marking and
// the new cell hash in ONE pass over the particle arrays.
//
// Replaces cusphs::ComputeStepVerlet / ComputeStepSymplecticPre / ComputeStepSymplecticCor (S/JSphGpuSimple_ker.cu:129-445)
// + cusph::ComputeStepPos / ComputeStepPos2 + KerUpdatePos (S/JSphGpu_ker.cu:320-378,1598-1690) + cushift::RunShifting
// (S/JSphShifting_ker.cu:132-170): the reference writes movxy/movz (24 B/particle) only to read them back in a second
// kernel; here the displacement never leaves registers.  Arithmetic (double on float state) follows the oracle,
// JSphCpu::ComputeVerletVarsFluid / ComputeSymplecticPre / ComputeSymplecticCorr / UpdatePos (S/JSphCpu.cpp:1240-1606)
// and JSphShifting::RunCpu (S/JSphShifting.cpp:387-417).
#pragma once
#include "devtypes.h"
#include "celldiv.cuh"
#include <cfloat>

namespace dsph {

constexpr int STEP_THREADS=256;

// JSphCpu::UpdatePos S/JSphCpu.cpp:1240-1293.  Returns the new code; writes pos/dcell.
__device__ __forceinline__ void UpdatePos(double rx,double ry,double rz,double movx,double movy,double movz,bool outrhop,
  unsigned p,double* __restrict__ pos,unsigned* __restrict__ dcell,typecode* __restrict__ code,const DevCtes &c)
{
  const bool outmove=(fabsf(float(movx))>c.movlimit || fabsf(float(movy))>c.movlimit || fabsf(float(movz))>c.movlimit);
  rx+=movx; ry+=movy; rz+=movz;
  double dx=rx-c.maprealposmin[0], dy=ry-c.maprealposmin[1], dz=rz-c.maprealposmin[2];
  bool out=(dx!=dx||dy!=dy||dz!=dz||dx<0||dy<0||dz<0||dx>=c.maprealsize[0]||dy>=c.maprealsize[1]||dz>=c.maprealsize[2]);
  if(c.periactive && out){
    if(c.perix){ if(dx<0){ dx-=c.perixinc[0]; dy-=c.perixinc[1]; dz-=c.perixinc[2]; } if(dx>=c.maprealsize[0]){ dx+=c.perixinc[0]; dy+=c.perixinc[1]; dz+=c.perixinc[2]; } }
    if(c.periy){ if(dy<0){ dx-=c.periyinc[0]; dy-=c.periyinc[1]; dz-=c.periyinc[2]; } if(dy>=c.maprealsize[1]){ dx+=c.periyinc[0]; dy+=c.periyinc[1]; dz+=c.periyinc[2]; } }
    if(c.periz){ if(dz<0){ dx-=c.perizinc[0]; dy-=c.perizinc[1]; dz-=c.perizinc[2]; } if(dz>=c.maprealsize[2]){ dx+=c.perizinc[0]; dy+=c.perizinc[1]; dz+=c.perizinc[2]; } }
    const bool outx=!c.perix && (dx<0||dx>=c.maprealsize[0]);
    const bool outy=!c.periy && (dy<0||dy>=c.maprealsize[1]);
    const bool outz=!c.periz && (dz<0||dz>=c.maprealsize[2]);
    out=(outx||outy||outz);
    rx=dx+c.maprealposmin[0]; ry=dy+c.maprealposmin[1]; rz=dz+c.maprealposmin[2];
  }
  pos[3*size_t(p)]=rx; pos[3*size_t(p)+1]=ry; pos[3*size_t(p)+2]=rz;
  if(outrhop||outmove||out){
    unsigned rcode=code[p]&(~DSPH_CODE_MASKSPECIAL);
    if(out)rcode|=DSPH_CODE_OUTPOS;
    else if(outrhop)rcode|=DSPH_CODE_OUTRHOP;
    else rcode|=DSPH_CODE_OUTMOVE;
    code[p]=typecode(rcode);
    dcell[p]=0xFFFFFFFFu;
  }
  else{
    if(c.periactive){ dx=rx-c.domposmin[0]; dy=ry-c.domposmin[1]; dz=rz-c.domposmin[2]; }
    const unsigned cx=unsigned(dx/c.scell_d), cy=unsigned(dy/c.scell_d), cz=unsigned(dz/c.scell_d);
    dcell[p]=DcCell(c.domcellcode,cx,cy,cz);
  }
}

// JSphShifting::RunCpu for one particle; returns the final shift displacement (0 when cancelled near boundaries).
__device__ __forceinline__ void ShiftDisp(const float4 rs,const float4 velrhop,double coefumagn,double coeftfs,float maxdist,const DevCtes &c,double &sx,double &sy,double &sz)
{
  if(rs.x!=FLT_MAX){
    const double vx=double(velrhop.x),vy=double(velrhop.y),vz=double(velrhop.z);
    double umagn=coefumagn*sqrt(vx*vx+vy*vy+vz*vz);
    if(c.shifttfs){
      if(rs.w<c.shifttfs)umagn=0;
      else umagn*=(double(rs.w)-c.shifttfs)/coeftfs;
    }
    const float dx=float(double(rs.x)*umagn), dy=float(double(rs.y)*umagn), dz=float(double(rs.z)*umagn);
    if(c.nn_on){ // the v5.0 multiphase fork clamps one side only (M/JSphShifting.cpp:412-414)
      sx=double(dx<maxdist? dx: maxdist); sy=double(dy<maxdist? dy: maxdist); sz=double(dz<maxdist? dz: maxdist);
      return;
    }
    sx=double(fabsf(dx)<maxdist? dx: (dx>=0? maxdist: -maxdist));
    sy=double(fabsf(dy)<maxdist? dy: (dy>=0? maxdist: -maxdist));
    sz=double(fabsf(dz)<maxdist? dz: (dz>=0? maxdist: -maxdist));
  }
  else{ sx=sy=sz=0; }
}

struct StepArrays {
  unsigned np,npb;
  double *pos; unsigned *dcell; typecode *code;
  float4 *velrhop;          // in: current; out: new (Symplectic) / becomes M1 (Verlet, pointer swap on the host)
  float4 *velrhopm1;        // Verlet: in M1, out new values
  double *pospre; float4 *velrhoppre;
  const float4 *acear;
  const float4 *shiftposfs; // raw accumulators of the force kernel (null: no shifting)
};

// Symplectic predictor (S/JSphCpu.cpp:1406-1497).  shift is never applied in the predictor (:1408).
__global__ void __launch_bounds__(STEP_THREADS) ker_step_sympre(StepArrays a,double dt,DevCtes c)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.np)return;
  const double dt05=dt*.5;
  const float4 vr=a.velrhop[p];
  const double px=a.pos[3*size_t(p)], py=a.pos[3*size_t(p)+1], pz=a.pos[3*size_t(p)+2];
  a.pospre[3*size_t(p)]=px; a.pospre[3*size_t(p)+1]=py; a.pospre[3*size_t(p)+2]=pz;
  a.velrhoppre[p]=vr;
  const float4 ac=a.acear[p];
  const float rhopnew=float(double(vr.w)+dt05*ac.w);
  if(p<a.npb){
    a.velrhop[p]=make_float4(vr.x,vr.y,vr.z,(rhopnew<c.rhopzero? c.rhopzero: rhopnew));
    return;
  }
  const unsigned rcode=a.code[p];
  const double dx=double(vr.x)*dt05, dy=double(vr.y)*dt05, dz=double(vr.z)*dt05;
  bool outrhop=(rhopnew<c.rhopoutmin || rhopnew>c.rhopoutmax);
  a.velrhop[p]=make_float4(float(double(vr.x)+(double(ac.x)+c.gravity[0])*dt05),
                           float(double(vr.y)+(double(ac.y)+c.gravity[1])*dt05),
                           float(double(vr.z)+(double(ac.z)+c.gravity[2])*dt05),rhopnew);
  const bool isnormal=((rcode&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
  unsigned code2=rcode;
  if(outrhop && isnormal){ code2=(rcode&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_OUTRHOP; a.code[p]=typecode(code2); }
  outrhop=((code2&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_OUTRHOP);
  const bool normal=(!c.ghosts || outrhop || (code2&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
  if(normal && (code2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID)UpdatePos(px,py,pz,dx,dy,dz,outrhop,p,a.pos,a.dcell,a.code,c);
}

// Symplectic corrector (S/JSphCpu.cpp:1503-1606), shifting folded in (S/JSphShifting.cpp:387-417).
__global__ void __launch_bounds__(STEP_THREADS) ker_step_symcor(StepArrays a,double dt,double coefumagn,double coeftfs,float maxdist,DevCtes c)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.np)return;
  const double dt05=dt*.5;
  const float4 vr=a.velrhop[p];           // predicted values
  const float4 vpre=a.velrhoppre[p];
  const float4 ac=a.acear[p];
  const double epsilon_rdot=(-double(ac.w)/double(vr.w))*dt;
  const float rhopnew=float(double(vpre.w)*(2.-epsilon_rdot)/(2.+epsilon_rdot));
  if(p<a.npb){
    a.velrhop[p]=make_float4(0,0,0,(rhopnew<c.rhopzero? c.rhopzero: rhopnew));
    // boundary position is restored from PosPre (predictor copied it unchanged): nothing to do
    return;
  }
  const unsigned rcode=a.code[p];
  const float4 rv=make_float4(float(double(vpre.x)+(double(ac.x)+c.gravity[0])*dt),
                              float(double(vpre.y)+(double(ac.y)+c.gravity[1])*dt),
                              float(double(vpre.z)+(double(ac.z)+c.gravity[2])*dt),rhopnew);
  double dx=(double(vpre.x)+double(rv.x))*dt05, dy=(double(vpre.y)+double(rv.y))*dt05, dz=(double(vpre.z)+double(rv.z))*dt05;
  if(a.shiftposfs){
    double sx,sy,sz; ShiftDisp(a.shiftposfs[p],vr,coefumagn,coeftfs,maxdist,c,sx,sy,sz);
    dx+=sx; dy+=sy; dz+=sz;
  }
  bool outrhop=(rhopnew<c.rhopoutmin || rhopnew>c.rhopoutmax);
  const bool isnormal=((rcode&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
  unsigned code2=rcode;
  if(outrhop && isnormal){ code2=(rcode&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_OUTRHOP; a.code[p]=typecode(code2); }
  a.velrhop[p]=rv;
  outrhop=((code2&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_OUTRHOP);
  const bool normal=(!c.ghosts || outrhop || (code2&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL);
  if(normal && (code2&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID){
    const double px=a.pospre[3*size_t(p)], py=a.pospre[3*size_t(p)+1], pz=a.pospre[3*size_t(p)+2];
    UpdatePos(px,py,pz,dx,dy,dz,outrhop,p,a.pos,a.dcell,a.code,c);
  }
}

// Verlet (S/JSphCpu.cpp:1300-1399).  New values are written to velrhopm1[]; the host swaps the two pointers.
__global__ void __launch_bounds__(STEP_THREADS) ker_step_verlet(StepArrays a,double dt,int euler,double coefumagn,double coeftfs,float maxdist,DevCtes c)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.np)return;
  const float4 vr1=a.velrhop[p];
  const float4 vr2=(euler? vr1: a.velrhopm1[p]);
  const double dt2=(euler? dt: dt+dt);
  const float4 ac=a.acear[p];
  const float rhopnew=float(double(vr2.w)+dt2*ac.w);
  if(p<a.npb){
    a.velrhopm1[p]=make_float4(0,0,0,(rhopnew<c.rhopzero? c.rhopzero: rhopnew));
    return;
  }
  // x-slab halo copies (not the reference's periodic duplicates, which Verlet does update) are never integrated
  if(c.ghosts && !c.periactive && (a.code[p]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC){ a.velrhopm1[p]=vr1; return; }
  const double dt205=0.5*dt*dt;
  const double agx=double(ac.x)+double(c.gravity[0]), agy=double(ac.y)+double(c.gravity[1]), agz=double(ac.z)+double(c.gravity[2]);
  double dx=double(vr1.x)*dt+agx*dt205, dy=double(vr1.y)*dt+agy*dt205, dz=double(vr1.z)*dt+agz*dt205;
  if(a.shiftposfs){
    double sx,sy,sz; ShiftDisp(a.shiftposfs[p],vr1,coefumagn,coeftfs,maxdist,c,sx,sy,sz);
    dx+=sx; dy+=sy; dz+=sz;
  }
  const bool outrhop=(rhopnew<c.rhopoutmin || rhopnew>c.rhopoutmax);
  const float4 rv=make_float4(float(double(vr2.x)+agx*dt2),float(double(vr2.y)+agy*dt2),float(double(vr2.z)+agz*dt2),rhopnew);
  const double px=a.pos[3*size_t(p)], py=a.pos[3*size_t(p)+1], pz=a.pos[3*size_t(p)+2];
  UpdatePos(px,py,pz,dx,dy,dz,outrhop,p,a.pos,a.dcell,a.code,c);
  a.velrhopm1[p]=rv;
}

} // namespace dsph
