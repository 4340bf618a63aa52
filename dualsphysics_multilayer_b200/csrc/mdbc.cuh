// mdbc.cuh -- modified Dynamic Boundary Conditions (mDBC, English et al. 2021): before the forces, the density of every
// boundary particle is replaced by the fluid density extrapolated to its ghost node (particle position + normal vector) with a
// first-order consistent correction, mirrored back along the normal.  What is computed is JSphCpu::InteractionMdbcCorrectionT2
// (S/JSphCpu.cpp:1020-1190; the reference GPU twins are KerInteractionMdbcCorrection_Fast/_Dbl, S/JSphGpu_ker.cu:912-1282), slip mode
// "velocity = 0" (the only one the reference accepts, S/JSph.cpp:788): only velrhop.w of the boundary particle changes.
//
// Mapping: the reference runs one thread per boundary particle over ~170 fluid neighbours of the ghost node, accumulating 16+5
// sums.  Here ONE WARP works per boundary particle: lanes stride over the candidates of each cell (coalesced), keep partial sums in
// the precision the reference CPU code uses (float for the density and its gradient and the kernel sum, double for the moment
// matrix), and combine them with shuffles; lane 0 inverts the 4x4 (3x3 in 2-D) matrix in double by cofactors and writes the new density
// AND the pressure that goes with it (posp.w is the per-particle pressure the force kernels read).
#pragma once
#include "forceparams.h"
#include "celldiv.cuh"
#include <cfloat>

namespace dsph {

struct MdbcArgs {
  unsigned npbok;
  const double* __restrict__ pos;            // xyz packed, cell-sorted
  const float4* __restrict__ boundnormal;    // particle -> ghost node (xyz), cell-sorted with the particles
  float4* velrhop;                           // in/out: .w of the boundary particles
  float4* posp;                              // out: .w = pressure of the corrected density
  float determlimit, threshold;
};

__device__ __forceinline__ double MdbcWarpSum(double v){
  #pragma unroll
  for(int o=16;o>0;o>>=1)v+=__shfl_xor_sync(0xffffffffu,v,o);
  return v;
}
__device__ __forceinline__ float MdbcWarpSumF(float v){
  #pragma unroll
  for(int o=16;o>0;o>>=1)v+=__shfl_xor_sync(0xffffffffu,v,o);
  return v;
}

// kernel value and gradient factor (fsph::GetKernel_WabFac, S/FunSphKernel.h:226-234 Wendland, :97-118 cubic).  Every operation is an
// explicitly rounded one (no FMA contraction): the CPU reference is an x86-64 build without FMA, and the corrected density is compared at
// the level of the last bit of a float near 1000 (see the ordered sums below): the 4x4 system amplifies a last-bit difference of a kernel
// value into ~10 ulp of the corrected density, i.e. 6e-5 of the wall pressure and of Ace next to the wall (measured: 421 of 692 wall
// particles off by up to 10 ulp with q = r*(1/h) instead of the quotient r/h; 7 of 692 off by 1 ulp with the operations below).
template<int TKER> __device__ __forceinline__ float MdbcWabFac(float rr2,const DevCtes &c,float &fac){
  const float rad=__fsqrt_rn(rr2), qq=__fdiv_rn(rad,c.kernelh);
  if(TKER==DSPH_KERNEL_WENDLAND){
    const float wqq1=__fsub_rn(1.f,__fmul_rn(0.5f,qq)), wqq2=__fmul_rn(wqq1,wqq1);
    // the association of the reference binary (objdump of oracle/_ref/obj52/JSphCpu.o, -O3 -ffast-math re-associates the products):
    // fac = ((bwen*q)*(w^2*w))/r,  wab = (awen*(1+2q))*(w^2*w^2)
    fac=__fdiv_rn(__fmul_rn(__fmul_rn(c.bwen,qq),__fmul_rn(wqq1,wqq2)),rad);
    const float wqq=__fadd_rn(__fadd_rn(qq,qq),1.f);
    return __fmul_rn(__fmul_rn(c.awen,wqq),__fmul_rn(wqq2,wqq2));
  }
  if(rad>c.kernelh){
    const float wqq1=__fsub_rn(2.0f,qq), wqq2=__fmul_rn(wqq1,wqq1);
    fac=__fdiv_rn(__fmul_rn(c.cub_c2,wqq2),rad);
    return __fmul_rn(__fmul_rn(c.cub_a24,wqq1),wqq2);            // (a24*w)*w^2 in the reference binary
  }
  const float wqq2=__fmul_rn(qq,qq);
  fac=__fdiv_rn(__fadd_rn(__fmul_rn(c.cub_c1,qq),__fmul_rn(c.cub_d1,wqq2)),rad);
  return __fmul_rn(c.cub_a2,__fadd_rn(1.0f,__fmul_rn(__fsub_rn(__fmul_rn(0.75f,qq),1.5f),wqq2)));
}

// x = A^-1 b by cofactors (the inverse the reference builds with fmath::InverseMatrix4x4 / 3x3, applied to b on the fly)
__device__ inline double MdbcDet3(const double m[3][3]){
  return m[0][0]*(m[1][1]*m[2][2]-m[1][2]*m[2][1])-m[0][1]*(m[1][0]*m[2][2]-m[1][2]*m[2][0])+m[0][2]*(m[1][0]*m[2][1]-m[1][1]*m[2][0]);
}
template<int N> __device__ inline double MdbcDet(const double a[4][4]){
  if(N==3){ double m[3][3]; for(int i=0;i<3;i++)for(int j=0;j<3;j++)m[i][j]=a[i][j]; return MdbcDet3(m); }
  double det=0;
  for(int cidx=0;cidx<4;cidx++){
    double m[3][3];
    for(int i=1;i<4;i++){ int cc=0; for(int j=0;j<4;j++)if(j!=cidx){ m[i-1][cc]=a[i][j]; cc++; } }
    const double d3=MdbcDet3(m);
    det+=((cidx&1)? -a[0][cidx]: a[0][cidx])*d3;
  }
  return det;
}
// Cramer's rule: column k of A replaced by b
template<int N> __device__ inline void MdbcSolve(const double a[4][4],const double b[4],double det,double x[4]){
  for(int k=0;k<N;k++){
    double t[4][4];
    for(int i=0;i<N;i++)for(int j=0;j<N;j++)t[i][j]=(j==k? b[i]: a[i][j]);
    x[k]=MdbcDet<N>(t)/det;
  }
}

template<int TKER,bool SIM2D> __global__ void __launch_bounds__(FORCES_THREADS) ker_mdbc_correction(MdbcArgs m,ForceParams t,DevCtes c)
{
  const unsigned lane=threadIdx.x&31u;
  const unsigned nwarps=(gridDim.x*blockDim.x)>>5;
  for(unsigned p1=(blockIdx.x*blockDim.x+threadIdx.x)>>5;p1<m.npbok;p1+=nwarps){
    const float4 bn=m.boundnormal[p1];
    if(bn.x==0.f && bn.y==0.f && bn.z==0.f)continue;                   // particles without normal keep their DBC density
    //-Ghost node position.
    const double gx=m.pos[3*size_t(p1)]+double(bn.x), gy=m.pos[3*size_t(p1)+1]+double(bn.y), gz=m.pos[3*size_t(p1)+2]+double(bn.z);
    float rhopp1=0,grx=0,gry=0,grz=0,sumwab=0;
    double a[4][4];
    #pragma unroll
    for(int i=0;i<4;i++){
      #pragma unroll
      for(int j=0;j<4;j++)a[i][j]=0;
    }
    //-Cells around the ghost node (nsearch::Init for a position, S/JCellSearch_inline.h:54-70).
    const int cx=int((gx-c.domposmin[0])/c.scell_d)-int(t.cellzero[0]);
    const int cy=int((gy-c.domposmin[1])/c.scell_d)-int(t.cellzero[1]);
    const int cz=int((gz-c.domposmin[2])/c.scell_d)-int(t.cellzero[2]);
    const int sd=c.scelldiv;
    const int xini=max(cx-sd,0), xfin=min(cx+sd,t.ncx-1)+1, yini=max(cy-sd,0), yfin=min(cy+sd,t.ncy-1)+1, zini=max(cz-sd,0), zfin=min(cz+sd,t.ncz-1)+1;
    if(xini<xfin)for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
      const unsigned v=unsigned(t.nsheet*z+t.ncx*y)+t.cellfluid;
      const unsigned pini=t.begincell[v+xini], pfin=t.begincell[v+xfin];
      for(unsigned base=pini;base<pfin;base+=32u){
        const unsigned p2=base+lane;
        float t_rho=0,t_gx=0,t_gy=0,t_gz=0,t_w=0;
        bool hit=false;
        if(p2<pfin){
          const float drx=float(gx-m.pos[3*size_t(p2)]), dry=float(gy-m.pos[3*size_t(p2)+1]), drz=float(gz-m.pos[3*size_t(p2)+2]);
          const float rr2=__fadd_rn(__fadd_rn(__fmul_rn(drx,drx),__fmul_rn(dry,dry)),__fmul_rn(drz,drz));
          if(rr2<=c.kernelsize2 && (t.code[p2]&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID){
            hit=true;
            float fac;
            const float wab=MdbcWabFac<TKER>(rr2,c,fac);
            const float frx=__fmul_rn(fac,drx), fry=__fmul_rn(fac,dry), frz=__fmul_rn(fac,drz);
            const float rhop2=m.velrhop[p2].w;
            const float massp2=c.massfluid;
            const float volp2=__fdiv_rn(massp2,rhop2);
            t_rho=__fmul_rn(massp2,wab); t_gx=__fmul_rn(massp2,frx); t_gy=__fmul_rn(massp2,fry); t_gz=__fmul_rn(massp2,frz);
            const float vwab=__fmul_rn(wab,volp2);
            t_w=vwab;
            const float vfrx=__fmul_rn(frx,volp2), vfry=__fmul_rn(fry,volp2), vfrz=__fmul_rn(frz,volp2);
            if(SIM2D){
              a[0][0]+=vwab; a[0][1]+=__fmul_rn(drx,vwab); a[0][2]+=__fmul_rn(drz,vwab);
              a[1][0]+=vfrx; a[1][1]+=__fmul_rn(drx,vfrx); a[1][2]+=__fmul_rn(drz,vfrx);
              a[2][0]+=vfrz; a[2][1]+=__fmul_rn(drx,vfrz); a[2][2]+=__fmul_rn(drz,vfrz);
            }
            else{
              a[0][0]+=vwab; a[0][1]+=__fmul_rn(drx,vwab); a[0][2]+=__fmul_rn(dry,vwab); a[0][3]+=__fmul_rn(drz,vwab);
              a[1][0]+=vfrx; a[1][1]+=__fmul_rn(drx,vfrx); a[1][2]+=__fmul_rn(dry,vfrx); a[1][3]+=__fmul_rn(drz,vfrx);
              a[2][0]+=vfry; a[2][1]+=__fmul_rn(drx,vfry); a[2][2]+=__fmul_rn(dry,vfry); a[2][3]+=__fmul_rn(drz,vfry);
              a[3][0]+=vfrz; a[3][1]+=__fmul_rn(drx,vfrz); a[3][2]+=__fmul_rn(dry,vfrz); a[3][3]+=__fmul_rn(drz,vfrz);
            }
          }
        }
        // The five FLOAT sums are added in the reference's order (ascending particle index): the corrected density is the small
        // difference rho-rho0 of sums of magnitude rho, so another order of the same float additions moves it by 1e-4 of itself
        // (measured) -- and with it the pressure of every wall particle.  The terms are computed by 32 lanes at once; only the
        // additions are serial (every lane keeps the same accumulators).
        unsigned mask=__ballot_sync(0xffffffffu,hit);
        while(mask){
          const int l=__ffs(int(mask))-1; mask&=mask-1u;
          rhopp1+=__shfl_sync(0xffffffffu,t_rho,l); grx+=__shfl_sync(0xffffffffu,t_gx,l); gry+=__shfl_sync(0xffffffffu,t_gy,l);
          grz+=__shfl_sync(0xffffffffu,t_gz,l); sumwab+=__shfl_sync(0xffffffffu,t_w,l);
        }
      }
    }
    constexpr int N=(SIM2D? 3: 4);
    #pragma unroll
    for(int i=0;i<N;i++){
      #pragma unroll
      for(int j=0;j<N;j++)a[i][j]=MdbcWarpSum(a[i][j]);
    }
    if(lane==0 && (sumwab>=m.threshold || (m.threshold>=2.f && sumwab+2.f>=m.threshold))){
      float rhopfinal=FLT_MAX;
      const double b[4]={double(rhopp1),double(grx),SIM2D? double(grz): double(gry),double(grz)};
      const double determ=MdbcDet<N>(a);
      if(fabs(determ)>=double(m.determlimit)){
        double x[4]; MdbcSolve<N>(a,b,determ,x);
        //-Ghost node density mirrored back to the boundary particle: dpos = -normal.
        const float rhoghost=float(x[0]);
        if(SIM2D){ const float gxx=-float(x[1]), gzz=-float(x[2]); rhopfinal=__fadd_rn(__fadd_rn(rhoghost,__fmul_rn(gxx,-bn.x)),__fmul_rn(gzz,-bn.z)); }
        else{ const float gxx=-float(x[1]), gyy=-float(x[2]), gzz=-float(x[3]); rhopfinal=__fadd_rn(__fadd_rn(__fadd_rn(rhoghost,__fmul_rn(gxx,-bn.x)),__fmul_rn(gyy,-bn.y)),__fmul_rn(gzz,-bn.z)); }
      }
      else if(a[0][0]>0)rhopfinal=float(double(rhopp1)/a[0][0]);          // determinant too small: zeroth order
      rhopfinal=(rhopfinal!=FLT_MAX? rhopfinal: c.rhopzero);
      float4 vr=m.velrhop[p1]; vr.w=rhopfinal; m.velrhop[p1]=vr;
      float4 pp=m.posp[p1]; pp.w=PressFromRhop(rhopfinal,c,false); m.posp[p1]=pp;
    }
  }
}

} // namespace dsph
