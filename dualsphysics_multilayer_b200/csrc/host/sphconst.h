// sphconst.h -- constants the reference derives on the host before the first step.
// Same formulas and the same float/double rounding points as the reference so that cell assignment
// is bit-exact:
//   JSph::ConfigConstants1/2      S/JSph.cpp:1392-1457   (KernelSize, Eta2, Cs0, kernel ctes, DDT, SPS, DtIni/DtMin)
//   JSph::LoadCaseParticles       S/JSph.cpp:2056-2080   (Map_PosMin/Max with periodic margin, Peri*inc)
//   JSph::ConfigCellDivision      S/JSph.cpp:1772-1779   (ScellDiv, Scell, MovLimit, Map_Cells)
//   JSph::SelecDomain             S/JSph.cpp:1794-1829   (Dom*, DomCellCode)
//   JDsDcell::CalcCellCode        S/JDsDcell.cpp:30-68, S/JDsDcellDef.h:33-44
#pragma once
#include "../../../include/dsph_b200.h"
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

namespace dsph {

struct SphConst {
  // kernel
  float kernelsize, kernelsize2, eta2;
  float awen, bwen;
  float cub_a1, cub_a2, cub_aa, cub_a24, cub_c1, cub_d1, cub_c2, cub_od_wdeltap;
  double cs0;
  // DDT / SPS
  float ddtkh, ddtgz, spssmag, spsblin;
  // time step
  double dtini, dtmin;
  // cells
  int scelldiv; float scell, movlimit;
  double mapposmin[3], mapposmax[3], mapsize[3];          // Map_PosMin/Max/Size (with periodic margin)
  double maprealposmin[3], maprealposmax[3], maprealsize[3];
  unsigned mapcells[3];
  unsigned domcells[3]; double domposmin[3], domposmax[3], domrealposmin[3], domrealposmax[3];
  unsigned domcellcode;
  double perixinc[3], periyinc[3], perizinc[3];
  int periactive; // bit0 X, bit1 Y, bit2 Z (DefPeriActive)
};

// ---- DomCellCode helpers (bit layout of S/JDsDcellDef.h:33-44) ----
inline unsigned DcelBits(unsigned v,unsigned minbits){ unsigned n=minbits; for(;v>>n;n++); return n; }
inline unsigned DcelGetCode(unsigned sx,unsigned sy,unsigned sz){
  return ((sx+1)<<25)|(sy<<20)|(sz<<15)|((sy+sz)<<10)|((sx+1+sz)<<5)|(sx+1+sy);
}
// maxbits: 31 in v5.2 (S/JDsDcell.cpp:39-68); the multiphase fork (v5.0) distributes 32 (M/JSph.cpp:1794-1822)
inline unsigned CalcCellCode(unsigned ncx,unsigned ncy,unsigned ncz,unsigned maxbits=31){
  unsigned sx=DcelBits(ncx,2), sy=DcelBits(ncy,2), sz=DcelBits(ncz,2);
  if(sx+sy+sz>maxbits)return 0;
  unsigned rest=maxbits-(sx+sy+sz);
  while(rest){
    if(rest){ sx++; rest--; }
    if(rest){ sy++; rest--; }
    if(rest){ sz++; rest--; }
  }
  return DcelGetCode(sx,sy,sz);
}
inline unsigned DcelCellx(unsigned dcc,unsigned cel){ return cel>>((dcc>>10)&31); }
inline unsigned DcelCelly(unsigned dcc,unsigned cel){ return (cel<<(dcc>>25))>>((dcc>>5)&31); }
inline unsigned DcelCellz(unsigned dcc,unsigned cel){ return (cel<<(dcc&31))>>(dcc&31); }
inline unsigned DcelCell(unsigned dcc,unsigned cx,unsigned cy,unsigned cz){ return (cx<<((dcc>>10)&31))|(cy<<((dcc>>15)&31))|cz; }

inline void DeriveConstants(const dsph_config &c,SphConst &k){
  memset(&k,0,sizeof(k));
  const bool sim2d=c.simulate2d!=0;
  const double h=c.kernelh;                  // KernelH is float (JSph.cpp:574), promoted
  const double kh=h*2.0;                     // kernel factor 2 for cubic and wendland (FunSphKernel.h)
  k.kernelsize=float(kh);
  k.kernelsize2=k.kernelsize*k.kernelsize;
  const double PI=3.14159265358979323846;
  if(sim2d){ k.awen=float(0.557/(h*h)); k.bwen=float(-2.7852/(h*h*h)); }
  else     { k.awen=float(0.41778/(h*h*h)); k.bwen=float(-2.08891/(h*h*h*h)); }
  {
    const double a1=(sim2d? 10./(PI*7.): 1./PI);
    const double a2=(sim2d? a1/(h*h): a1/(h*h*h));
    const double aa=(sim2d? a1/(h*h*h): a1/(h*h*h*h));
    const double deltap=1./1.5;
    const double wdeltap=a2*(1.-1.5*deltap*deltap+0.75*deltap*deltap*deltap);
    k.cub_od_wdeltap=float(1./wdeltap);
    k.cub_a1=float(a1); k.cub_a2=float(a2); k.cub_aa=float(aa); k.cub_a24=float(0.25*a2);
    k.cub_c1=float(-3.*aa); k.cub_d1=float(9.*aa/4.); k.cub_c2=float(-3.*aa/4.);
  }
  k.cs0=sqrt(double(c.gamma)*double(c.cteb)/double(c.rhopzero));
  k.eta2=float((h*0.1)*(h*0.1));
  if(c.tvisco==DSPH_VISCO_LAMINARSPS){
    const double dp_sps=(sim2d? sqrt(c.dp*c.dp*2.)/2.: sqrt(c.dp*c.dp*3.)/3.);
    k.spssmag=float(pow((0.12*dp_sps),2));
    k.spsblin=float((2./3.)*0.0066*dp_sps*dp_sps);
  }
  const float ddtvalue=(c.tdensity==DSPH_DDT_NONE? 0.f: c.ddtvalue);
  k.ddtkh=k.kernelsize*ddtvalue;
  k.ddtgz=float(double(c.rhopzero)*double(fabsf(c.gravity[2]))/double(c.cteb));
  k.dtini=(c.dtini>0? c.dtini: double(c.kernelh)/k.cs0);
  k.dtmin=(c.dtmin>0? c.dtmin: (double(c.kernelh)/k.cs0)*double(c.coefdtmin));
  if(c.nn_on){ float coefdtmin=c.coefdtmin; coefdtmin*=1.0e-5f; k.dtmin=(c.kernelh/k.cs0)*coefdtmin; }   // JSph::ConfigConstantsMP M/JSph.cpp:3240-3241
  // ---- map limits
  for(int i=0;i<3;i++){ k.maprealposmin[i]=c.maprealposmin[i]; k.maprealposmax[i]=c.maprealposmax[i]; k.maprealsize[i]=k.maprealposmax[i]-k.maprealposmin[i]; }
  for(int i=0;i<3;i++){ k.perixinc[i]=c.perixinc[i]; k.periyinc[i]=c.periyinc[i]; k.perizinc[i]=c.perizinc[i]; }
  if(c.perix)k.perixinc[0]=-k.maprealsize[0];
  if(c.periy)k.periyinc[1]=-k.maprealsize[1];
  if(c.periz)k.perizinc[2]=-k.maprealsize[2];
  k.periactive=(c.perix?1:0)|(c.periy?2:0)|(c.periz?4:0);
  const int peri[3]={c.perix,c.periy,c.periz};
  for(int i=0;i<3;i++){
    k.mapposmin[i]=k.maprealposmin[i]; k.mapposmax[i]=k.maprealposmax[i];
    if(peri[i]){ k.mapposmin[i]=k.mapposmin[i]-k.kernelsize; k.mapposmax[i]=k.mapposmax[i]+k.kernelsize; }
    k.mapsize[i]=k.mapposmax[i]-k.mapposmin[i];
  }
  // ---- cells
  if(c.cellmode!=DSPH_CELLMODE_FULL && c.cellmode!=DSPH_CELLMODE_HALF)throw std::runtime_error("The CellMode is invalid.");
  k.scelldiv=(c.cellmode==DSPH_CELLMODE_FULL? 1: 2);
  k.scell=k.kernelsize/k.scelldiv;
  k.movlimit=k.scell*0.9f;
  for(int i=0;i<3;i++)k.mapcells[i]=unsigned(ceil(k.mapsize[i]/k.scell));
  // ---- single domain = whole map (SelecDomain(TUint3(0,0,0),Map_Cells))
  for(int i=0;i<3;i++){
    k.domcells[i]=k.mapcells[i];
    if(k.domcells[i]<1)throw std::runtime_error("The domain of cells is invalid.");
    k.domposmin[i]=k.mapposmin[i]+(0*k.scell);
    k.domposmax[i]=k.mapposmin[i]+(k.mapcells[i]*k.scell);   // unsigned*float -> float, promoted to double like the reference
    if(k.domposmax[i]>k.mapposmax[i])k.domposmax[i]=k.mapposmax[i];
    k.domrealposmin[i]=k.domposmin[i]; k.domrealposmax[i]=k.domposmax[i];
    if(k.domrealposmax[i]>k.maprealposmax[i])k.domrealposmax[i]=k.maprealposmax[i];
    if(k.domrealposmin[i]<k.maprealposmin[i])k.domrealposmin[i]=k.maprealposmin[i];
  }
  k.domcellcode=CalcCellCode(k.domcells[0]+1,k.domcells[1]+1,k.domcells[2]+1,(c.nn_on? 32u: 31u));
  if(!k.domcellcode)throw std::runtime_error("Failed to select a valid CellCode for the number of cells.");
}

} // namespace dsph
