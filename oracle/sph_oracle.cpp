// sph_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-step particle hot
// path (JSphCpuSingle / JSphCpu / JCellDivCpuSingle), written from the algorithm, every function citing
// the reference file:line it follows.  S/ = /root/reference/src/source/.
//
// It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg may load it.  The product (dualsphysics_multilayer_b200/) never links, imports or calls it.
// PARITY PIN: the reference ships no golden vectors; this restatement is pinned against
// dumps of the reference's own JSphCpuSingle produced by oracle/ref_harness.cpp (tests/golden/*.npz,
// generator tests/golden/make_golden.py) -- see tests/test_oracle_vs_reference.py.
//
// The exported C functions mirror include/dsph_b200.h one to one (orc_* instead of dsph_*), so parity tests
// drive both sides with the same calls.  Arithmetic keeps the reference's float/double rounding points.
#include "../include/dsph_b200.h"
#include <cmath>
#include <cfloat>
#include <climits>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <string>
#include <algorithm>
#include <stdexcept>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef uint16_t typecode;
struct f3{ float x,y,z; };
struct f4{ float x,y,z,w; };
struct d3{ double x,y,z; };
struct sym3{ float xx,xy,xz,yy,yz,zz; };

// ---- typecode helpers: S/DualSphDef.h:164-226
inline typecode CodeSpecial(typecode c){ return c&DSPH_CODE_MASKSPECIAL; }
inline typecode CodeType(typecode c){ return c&DSPH_CODE_MASKTYPE; }
inline bool CodeIsNormal(typecode c){ return CodeSpecial(c)==DSPH_CODE_NORMAL; }
inline bool CodeIsPeriodic(typecode c){ return CodeSpecial(c)==DSPH_CODE_PERIODIC; }
inline bool CodeIsFluid(typecode c){ return CodeType(c)==DSPH_CODE_TYPE_FLUID; }
inline bool CodeIsFixed(typecode c){ return CodeType(c)==DSPH_CODE_TYPE_FIXED; }
inline unsigned CodeTypeValue(typecode c){ return c&DSPH_CODE_MASKVALUE; }
inline bool CodeIsOutRhop(typecode c){ return CodeSpecial(c)==DSPH_CODE_OUTRHOP; }
inline typecode CodeSetNormal(typecode c){ return c&(~DSPH_CODE_MASKSPECIAL); }

// ---- dcell bit layout: S/JDsDcellDef.h:33-44
inline unsigned DcBits(unsigned v,unsigned minbits){ unsigned n=minbits; for(;v>>n;n++); return n; }
inline unsigned DcCode(unsigned sx,unsigned sy,unsigned sz){ return ((sx+1)<<25)|(sy<<20)|(sz<<15)|((sy+sz)<<10)|((sx+1+sz)<<5)|(sx+1+sy); }
inline unsigned DcCellx(unsigned dcc,unsigned cel){ return cel>>((dcc>>10)&31); }
inline unsigned DcCelly(unsigned dcc,unsigned cel){ return (cel<<(dcc>>25))>>((dcc>>5)&31); }
inline unsigned DcCellz(unsigned dcc,unsigned cel){ return (cel<<(dcc&31))>>(dcc&31); }
inline unsigned DcCell(unsigned dcc,unsigned cx,unsigned cy,unsigned cz){ return (cx<<((dcc>>10)&31))|(cy<<((dcc>>15)&31))|cz; }
// JDsDcell::CalcCellDistribution/CalcCellCode: S/JDsDcell.cpp:39-68
// (the v5.0 multiphase fork distributes 32 bits instead of 31: JSph::CalcCellDistribution M/JSph.cpp:1794-1822)
unsigned CalcCellCode(unsigned nx,unsigned ny,unsigned nz,unsigned maxbits){
  unsigned sx=DcBits(nx,2),sy=DcBits(ny,2),sz=DcBits(nz,2);
  if(sx+sy+sz>maxbits)return 0;
  unsigned rest=maxbits-(sx+sy+sz);
  while(rest){ if(rest){ sx++; rest--; } if(rest){ sy++; rest--; } if(rest){ sz++; rest--; } }
  return DcCode(sx,sy,sz);
}

struct Oracle {
  dsph_config cfg;
  // ---- derived constants: JSph::ConfigConstants1/2 S/JSph.cpp:1392-1457
  float KernelH,KernelSize,KernelSize2,Eta2,Awen,Bwen;
  float CubA1,CubA2,CubAa,CubA24,CubC1,CubD1,CubC2,CubOdWdeltap;
  float CteB,Gamma,RhopZero,MassFluid,MassBound,Visco,ViscoBoundFactor;
  float DDTkh,DDTgz,SpsSmag,SpsBlin;
  f3 Gravity;
  double Cs0,Dp,CFLnumber,DtIni,DtMin;
  float RhopOutMin,RhopOutMax;
  bool Simulate2D,DDTArray;
  int TKernel,TStep,TVisco,TDensity,ShiftMode,VerletSteps;
  float ShiftCoef,ShiftTFS;
  // ---- domain: JSph::LoadCaseParticles S/JSph.cpp:2056-2080, ConfigCellDivision :1772-1779, SelecDomain :1794-1829
  d3 MapRealPosMin,MapRealPosMax,MapRealSize,Map_PosMin,Map_PosMax,Map_Size;
  d3 DomPosMin,DomPosMax,DomRealPosMin,DomRealPosMax;
  d3 PeriXinc,PeriYinc,PeriZinc;
  bool PeriX,PeriY,PeriZ; int PeriActive;
  unsigned MapCells[3],DomCells[3],DomCellCode;
  int ScellDiv; float Scell,MovLimit;
  // ---- particle state (JSphCpu members S/JSphCpu.h:95-160)
  unsigned Np=0,Npb=0,NpbOk=0,NpbPer=0,NpfPer=0;
  std::vector<unsigned> Idp,Dcell; std::vector<typecode> Code;
  std::vector<d3> Pos,PosPre; std::vector<f4> Velrhop,VelrhopM1,VelrhopPre;
  bool HasPre=false;
  std::vector<f3> Ace; std::vector<float> Ar,Delta,Press; std::vector<f4> ShiftPosfs;
  std::vector<sym3> SpsTau,SpsGradvel;
  double VelMax=0,AceMax=0; float ViscDtMax=0;
  double SymplecticDtPre=0; int VerletStep=0;
  bool BoundChanged=true;
  // ---- mDBC (S/JSph.h:192-199): boundary mode, vectors particle -> ghost node of the boundary particles, sorted with them
  bool Mdbc=false,MdbcCorrector=false; float MdbcThreshold=0; std::vector<f3> BoundNormal;
  // ---- in-solver gauges (JGaugeSystem / JGaugeItem, S/JDsGaugeItem.cpp): definitions, next compute time, last results, run time
  std::vector<dsph_gauge> Gauges; std::vector<double> GaugeNext; std::vector<dsph_gaugeres> GaugeRes; double GaugeTime=0;
  // ---- cell division (JCellDivCpu members S/JCellDivCpu.h:44-150)
  std::vector<unsigned> CellPart,SortPart,PartsInCell,BeginCell;
  unsigned Ncx=0,Ncy=0,Ncz=0,Nsheet=0,Nct=0,Nctt=0;
  unsigned BoxBoundIgnore=0,BoxFluid=0,BoxBoundOut=0,BoxFluidOut=0,BoxBoundOutIgnore=0,BoxFluidOutIgnore=0;
  unsigned CellDomainMin[3]={0,0,0},CellDomainMax[3]={0,0,0};
  bool BoundLimitOk=false,BoundDivideOk=false,DivideFull=false;
  unsigned BoundLimitCellMin[3],BoundLimitCellMax[3],BoundDivideCellMin[3],BoundDivideCellMax[3];
  unsigned Npb1=0,Npf1=0,Npb2=0,Npf2=0,Nptot=0,NpbOut=0,NpfOut=0,NpbOutIgnore=0,NpfOutIgnore=0,NpFinal=0,NpbFinal=0,NpbIgnore=0;
  unsigned long long PipsFluidReal=0,PipsFluidChecked=0,PipsBoundReal=0,PipsBoundChecked=0;
  std::string Error;
  // ---- x-slab decomposition (test twin of csrc/mgpu.cuh; the reference has no multi-GPU code)
  bool Mgpu=false; int MgRank=0,MgNranks=1; unsigned MgXlo=0,MgXhi=0; unsigned MgCounts[8]={0,0,0,0,0,0,0,0};
  std::vector<unsigned char> MgFlags; unsigned MaxBits[4]={0,0,0,0}; float LastAce2=0,LastVel2=0;
  bool Ghosts()const{ return PeriActive!=0 || Mgpu; }
  // ---- multiphase non-Newtonian fork (M/JSph.h:519-531, M/DualSphDef.h StPhaseCte/StPhaseArray, M/JSphCpu.h:363-371)
  struct Phase{ float rho,mass,cs0,cteb,gamma,visco,tau_yield,tau_max,bi_multi,m_nn,n_nn; };
  bool MultiPhase=false; int TVelGrad=0; unsigned PhaseCount=0; float Lamda=0.2f; Phase Ph[DSPH_MAXPHASES];
  std::vector<float> ViscoEta; std::vector<sym3> Dtensor;
  float ViscEtaDtMax=0;

  void Resize(unsigned n){
    Idp.resize(n); Dcell.resize(n); Code.resize(n); Pos.resize(n); Velrhop.resize(n);
    if(TStep==DSPH_STEP_VERLET)VelrhopM1.resize(n);
    if(TStep==DSPH_STEP_SYMPLECTIC){ PosPre.resize(n); VelrhopPre.resize(n); }
    if(TVisco==DSPH_VISCO_LAMINARSPS)SpsTau.resize(n);
  }

  void Configure(const dsph_config &c){
    cfg=c;
    TKernel=c.tkernel; TStep=c.tstep; TVisco=c.tvisco; TDensity=c.tdensity; ShiftMode=c.shiftmode; VerletSteps=c.verletsteps;
    Mdbc=((c.tboundary&0xff)==DSPH_BOUNDARY_MDBC); MdbcCorrector=(Mdbc && (c.tboundary&DSPH_BOUNDARY_MDBCCORRECTOR)!=0);
    ShiftCoef=c.shiftcoef; ShiftTFS=c.shifttfs;
    Simulate2D=c.simulate2d!=0;
    KernelH=c.kernelh; CteB=c.cteb; Gamma=c.gamma; RhopZero=c.rhopzero; MassFluid=c.massfluid; MassBound=c.massbound;
    Gravity={c.gravity[0],c.gravity[1],c.gravity[2]};
    Visco=c.visco; ViscoBoundFactor=c.viscoboundfactor; Dp=c.dp; CFLnumber=c.cflnumber;
    RhopOutMin=c.rhopoutmin; RhopOutMax=c.rhopoutmax;
    if(!(RhopOutMin<RhopOutMax)){ RhopOutMin=-FLT_MAX; RhopOutMax=FLT_MAX; }   // S/JSph.cpp:862-863
    // ConfigConstants1 (S/JSph.cpp:1392-1431)
    const double h=KernelH, kh=h*2.0;
    KernelSize=float(kh); KernelSize2=KernelSize*KernelSize;
    if(Simulate2D){ Awen=float(0.557/(h*h)); Bwen=float(-2.7852/(h*h*h)); }          // S/FunSphKernel.h:199-210
    else          { Awen=float(0.41778/(h*h*h)); Bwen=float(-2.08891/(h*h*h*h)); }
    {                                                                                // S/FunSphKernel.h:50-80
      const double PI=3.14159265358979323846;
      const double a1=(Simulate2D? 10./(PI*7.): 1./PI);
      const double a2=(Simulate2D? a1/(h*h): a1/(h*h*h));
      const double aa=(Simulate2D? a1/(h*h*h): a1/(h*h*h*h));
      const double deltap=1./1.5;
      const double wdeltap=a2*(1.-1.5*deltap*deltap+0.75*deltap*deltap*deltap);
      CubOdWdeltap=float(1./wdeltap); CubA1=float(a1); CubA2=float(a2); CubAa=float(aa); CubA24=float(0.25*a2);
      CubC1=float(-3.*aa); CubD1=float(9.*aa/4.); CubC2=float(-3.*aa/4.);
    }
    Cs0=sqrt(double(Gamma)*double(CteB)/double(RhopZero));
    Eta2=float((h*0.1)*(h*0.1));
    // ConfigConstants2 (S/JSph.cpp:1437-1457)
    SpsSmag=SpsBlin=0;
    if(TVisco==DSPH_VISCO_LAMINARSPS){
      const double dp_sps=(Simulate2D? sqrt(Dp*Dp*2.)/2.: sqrt(Dp*Dp*3.)/3.);
      SpsSmag=float(pow((0.12*dp_sps),2)); SpsBlin=float((2./3.)*0.0066*dp_sps*dp_sps);
    }
    const float ddtvalue=(TDensity==DSPH_DDT_NONE? 0.f: c.ddtvalue);            // S/JSph.cpp:811
    DDTkh=KernelSize*ddtvalue;
    DDTgz=float(double(RhopZero)*double(fabsf(Gravity.z))/double(CteB));
    DDTArray=(TDensity!=DSPH_DDT_NONE);                                         // Cpu: S/JSph.cpp:814
    DtIni=(c.dtini>0? c.dtini: double(KernelH)/Cs0);
    DtMin=(c.dtmin>0? c.dtmin: (double(KernelH)/Cs0)*double(c.coefdtmin));
    // Multiphase: JSph::InitMultiPhase + ConfigConstantsMP M/JSph.cpp:3137-3242 (phases arrive sorted by mkfluid)
    MultiPhase=(c.nn_on!=0); TVelGrad=c.nn_velgrad; PhaseCount=0; Lamda=c.nn_relaxationdt;
    if(MultiPhase){
      if(c.nn_nphases<1 || c.nn_nphases>DSPH_MAXPHASES)throw std::runtime_error("The number of phases is invalid.");
      if(TVelGrad!=DSPH_VELGRAD_FDA && TVelGrad!=DSPH_VELGRAD_SPH)throw std::runtime_error("Velocity gradient type is not valid.");
      PhaseCount=unsigned(c.nn_nphases);
      bool cs0present=true;
      for(unsigned i=0;i<PhaseCount;i++){
        const dsph_phase &s=c.phases[i]; Phase &d=Ph[i];
        d.rho=s.rho; d.cs0=s.cs0; d.gamma=(s.gamma? s.gamma: Gamma);
        d.cteb=(d.cs0? d.cs0*d.cs0*d.rho/d.gamma: 0.f);
        d.visco=s.visco; d.tau_yield=s.tau_yield; d.tau_max=s.tau_max; d.bi_multi=(s.tau_max? s.bi_multi: 0.f); d.m_nn=s.m_nn; d.n_nn=s.n_nn;
        if(!d.cs0)cs0present=false;
      }
      for(unsigned i=0;i<PhaseCount;i++){
        if(!cs0present)Ph[i].cteb=float(Cs0*Cs0*Ph[i].rho/Ph[i].gamma);
        Ph[i].mass=float(Ph[i].rho*(Simulate2D? Dp*Dp: Dp*Dp*Dp));
      }
      float coefdtmin=c.coefdtmin; coefdtmin*=1.0e-5f;
      DtMin=(KernelH/Cs0)*coefdtmin;                                             // overrides any DtMin of the XML
    }
    // LoadCaseParticles (S/JSph.cpp:2056-2080)
    MapRealPosMin={c.maprealposmin[0],c.maprealposmin[1],c.maprealposmin[2]};
    MapRealPosMax={c.maprealposmax[0],c.maprealposmax[1],c.maprealposmax[2]};
    MapRealSize={MapRealPosMax.x-MapRealPosMin.x,MapRealPosMax.y-MapRealPosMin.y,MapRealPosMax.z-MapRealPosMin.z};
    PeriX=c.perix!=0; PeriY=c.periy!=0; PeriZ=c.periz!=0; PeriActive=(PeriX?1:0)|(PeriY?2:0)|(PeriZ?4:0);
    PeriXinc={c.perixinc[0],c.perixinc[1],c.perixinc[2]}; PeriYinc={c.periyinc[0],c.periyinc[1],c.periyinc[2]}; PeriZinc={c.perizinc[0],c.perizinc[1],c.perizinc[2]};
    if(PeriX)PeriXinc.x=-MapRealSize.x;
    if(PeriY)PeriYinc.y=-MapRealSize.y;
    if(PeriZ)PeriZinc.z=-MapRealSize.z;
    Map_PosMin=MapRealPosMin; Map_PosMax=MapRealPosMax;
    if(PeriX){ Map_PosMin.x=Map_PosMin.x-KernelSize; Map_PosMax.x=Map_PosMax.x+KernelSize; }
    if(PeriY){ Map_PosMin.y=Map_PosMin.y-KernelSize; Map_PosMax.y=Map_PosMax.y+KernelSize; }
    if(PeriZ){ Map_PosMin.z=Map_PosMin.z-KernelSize; Map_PosMax.z=Map_PosMax.z+KernelSize; }
    Map_Size={Map_PosMax.x-Map_PosMin.x,Map_PosMax.y-Map_PosMin.y,Map_PosMax.z-Map_PosMin.z};
    // ConfigCellDivision (S/JSph.cpp:1772-1779)
    ScellDiv=(c.cellmode==DSPH_CELLMODE_FULL? 1: 2);
    Scell=KernelSize/ScellDiv; MovLimit=Scell*0.9f;
    MapCells[0]=unsigned(ceil(Map_Size.x/Scell)); MapCells[1]=unsigned(ceil(Map_Size.y/Scell)); MapCells[2]=unsigned(ceil(Map_Size.z/Scell));
    // SelecDomain(0,Map_Cells) (S/JSph.cpp:1794-1829)
    for(int i=0;i<3;i++)DomCells[i]=MapCells[i];
    DomPosMin=Map_PosMin;
    DomPosMax.x=Map_PosMin.x+(MapCells[0]*Scell); DomPosMax.y=Map_PosMin.y+(MapCells[1]*Scell); DomPosMax.z=Map_PosMin.z+(MapCells[2]*Scell);
    if(DomPosMax.x>Map_PosMax.x)DomPosMax.x=Map_PosMax.x;
    if(DomPosMax.y>Map_PosMax.y)DomPosMax.y=Map_PosMax.y;
    if(DomPosMax.z>Map_PosMax.z)DomPosMax.z=Map_PosMax.z;
    DomRealPosMin=DomPosMin; DomRealPosMax=DomPosMax;
    if(DomRealPosMax.x>MapRealPosMax.x)DomRealPosMax.x=MapRealPosMax.x;
    if(DomRealPosMax.y>MapRealPosMax.y)DomRealPosMax.y=MapRealPosMax.y;
    if(DomRealPosMax.z>MapRealPosMax.z)DomRealPosMax.z=MapRealPosMax.z;
    if(DomRealPosMin.x<MapRealPosMin.x)DomRealPosMin.x=MapRealPosMin.x;
    if(DomRealPosMin.y<MapRealPosMin.y)DomRealPosMin.y=MapRealPosMin.y;
    if(DomRealPosMin.z<MapRealPosMin.z)DomRealPosMin.z=MapRealPosMin.z;
    DomCellCode=CalcCellCode(DomCells[0]+1,DomCells[1]+1,DomCells[2]+1,(MultiPhase? 32u: 31u));
    if(!DomCellCode)throw std::runtime_error("Failed to select a valid CellCode.");
    SymplecticDtPre=DtIni;                                                        // JSph::InitRun S/JSph.cpp:2090
    VerletStep=0;
  }

  // ---- JSph::LoadDcellParticles S/JSph.cpp:1690-1711
  void LoadDcellParticles(){
    for(unsigned p=0;p<Np;p++){
      if(CodeSpecial(Code[p])<DSPH_CODE_OUTIGNORE){
        const d3 ps=Pos[p];
        if(CodeIsPeriodic(Code[p])){   // restart of a dumped state that still holds periodic duplicates: PeriodicDuplicatePos rule
          unsigned cx=unsigned((ps.x-DomPosMin.x)/Scell), cy=unsigned((ps.y-DomPosMin.y)/Scell), cz=unsigned((ps.z-DomPosMin.z)/Scell);
          cx=(cx<=DomCells[0]? cx: DomCells[0]); cy=(cy<=DomCells[1]? cy: DomCells[1]); cz=(cz<=DomCells[2]? cz: DomCells[2]);
          Dcell[p]=DcCell(DomCellCode,cx,cy,cz);
          continue;
        }
        const bool in=(ps.x>=DomRealPosMin.x && ps.y>=DomRealPosMin.y && ps.z>=DomRealPosMin.z && ps.x<DomRealPosMax.x && ps.y<DomRealPosMax.y && ps.z<DomRealPosMax.z);
        if(!in)throw std::runtime_error("Found new particles out.");
        const double dx=ps.x-DomPosMin.x, dy=ps.y-DomPosMin.y, dz=ps.z-DomPosMin.z;
        const unsigned cx=unsigned(dx/Scell), cy=unsigned(dy/Scell), cz=unsigned(dz/Scell);
        Dcell[p]=DcCell(DomCellCode,cx,cy,cz);
      }
      else Dcell[p]=0xFFFFFFFFu;
    }
  }

  // ---- JSphCpuSingle::RunPeriodic + PeriodicMakeList + PeriodicDuplicate* S/JSphCpuSingle.cpp:182-431
  void PeriodicDuplicate(unsigned pnew,unsigned rp,const d3 &perinc){
    const unsigned pcopy=(rp&0x7FFFFFFFu); const bool inverse=(rp>=0x80000000u);
    d3 ps=Pos[pcopy];
    ps.x+=(inverse? -perinc.x: perinc.x); ps.y+=(inverse? -perinc.y: perinc.y); ps.z+=(inverse? -perinc.z: perinc.z);
    unsigned cx=unsigned((ps.x-DomPosMin.x)/Scell), cy=unsigned((ps.y-DomPosMin.y)/Scell), cz=unsigned((ps.z-DomPosMin.z)/Scell);
    cx=(cx<=DomCells[0]? cx: DomCells[0]); cy=(cy<=DomCells[1]? cy: DomCells[1]); cz=(cz<=DomCells[2]? cz: DomCells[2]);
    Pos[pnew]=ps; Dcell[pnew]=DcCell(DomCellCode,cx,cy,cz);
    Idp[pnew]=Idp[pcopy]; Code[pnew]=typecode(CodeSetNormal(Code[pcopy])|DSPH_CODE_PERIODIC);
    Velrhop[pnew]=Velrhop[pcopy];
    if(TStep==DSPH_STEP_VERLET)VelrhopM1[pnew]=VelrhopM1[pcopy];
    if(TStep==DSPH_STEP_SYMPLECTIC && HasPre){ PosPre[pnew]=PosPre[pcopy]; VelrhopPre[pnew]=VelrhopPre[pcopy]; }
    if(TVisco==DSPH_VISCO_LAMINARSPS)SpsTau[pnew]=SpsTau[pcopy];
  }
  void RunPeriodic(){
    for(unsigned p=0;p<Np;p++){ const typecode rc=Code[p]; if(CodeIsPeriodic(rc))Code[p]=typecode(CodeSetNormal(rc)|DSPH_CODE_OUTIGNORE); }
    const unsigned npb0=Npb, npf0=Np-Npb;
    NpbPer=NpfPer=0; BoundChanged=true;
    for(unsigned ctype=0;ctype<2;ctype++){
      const unsigned pini=(ctype? npb0: 0), num=(ctype? npf0: npb0);
      for(unsigned cper=0;cper<3;cper++)if((cper==0&&PeriX)||(cper==1&&PeriY)||(cper==2&&PeriZ)){
        const d3 perinc=(cper==0? PeriXinc: (cper==1? PeriYinc: PeriZinc));
        for(unsigned cblock=0;cblock<2;cblock++){
          const unsigned nper=(ctype? NpfPer: NpbPer);
          const unsigned pini2=(cblock? pini: Np-nper), num2=(cblock? num: nper);
          if(!num2)continue;
          std::vector<unsigned> listp;
          for(unsigned p=0;p<num2;p++){
            const unsigned p2=p+pini2;
            if(CodeSpecial(Code[p2])<=DSPH_CODE_PERIODIC){
              const d3 ps=Pos[p2];
              d3 q={ps.x+perinc.x,ps.y+perinc.y,ps.z+perinc.z};
              if(Map_PosMin.x<=q.x && Map_PosMin.y<=q.y && Map_PosMin.z<=q.z && q.x<Map_PosMax.x && q.y<Map_PosMax.y && q.z<Map_PosMax.z)listp.push_back(p2);
              q={ps.x-perinc.x,ps.y-perinc.y,ps.z-perinc.z};
              if(Map_PosMin.x<=q.x && Map_PosMin.y<=q.y && Map_PosMin.z<=q.z && q.x<Map_PosMax.x && q.y<Map_PosMax.y && q.z<Map_PosMax.z)listp.push_back(p2|0x80000000u);
            }
          }
          const unsigned count=unsigned(listp.size());
          if(Idp.size()<size_t(Np)+count)Resize(Np+count+(Np+count)/20+128);
          for(unsigned p=0;p<count;p++)PeriodicDuplicate(Np+p,listp[p],perinc);
          Np+=count;
          if(!ctype)NpbPer+=count; else NpfPer+=count;
        }
      }
    }
  }

  // ---- JCellDivCpu::LimitsCellBound/Fluid + CalcCellDomainBound/Fluid S/JCellDivCpu.cpp:255-350
  void LimitsCell(unsigned n,unsigned pini,unsigned cmin[3],unsigned cmax[3])const{
    cmin[0]=cmin[1]=cmin[2]=UINT_MAX; cmax[0]=cmax[1]=cmax[2]=0;
    for(unsigned p=pini;p<pini+n;p++){
      const unsigned rcell=Dcell[p];
      const unsigned cx=DcCellx(DomCellCode,rcell),cy=DcCelly(DomCellCode,rcell),cz=DcCellz(DomCellCode,rcell);
      if(CodeSpecial(Code[p])<DSPH_CODE_OUTIGNORE){
        if(cmin[0]>cx)cmin[0]=cx; if(cmin[1]>cy)cmin[1]=cy; if(cmin[2]>cz)cmin[2]=cz;
        if(cmax[0]<cx)cmax[0]=cx; if(cmax[1]<cy)cmax[1]=cy; if(cmax[2]<cz)cmax[2]=cz;
      }
    }
  }
  void CalcCellDomainRange(unsigned n,unsigned pini,unsigned n2,unsigned pini2,unsigned cellmin[3],unsigned cellmax[3])const{
    unsigned cmin[3],cmax[3];
    LimitsCell(n,pini,cmin,cmax);
    const bool empty=(cmin[0]>cmax[0]);
    for(int i=0;i<3;i++){ cellmin[i]=(empty? DomCells[i]: cmin[i]); cellmax[i]=(empty? 0: cmax[i]); }
    if(n2){
      LimitsCell(n2,pini2,cmin,cmax);
      const bool empty2=(cmin[0]>cmax[0]);
      for(int i=0;i<3;i++){
        const unsigned a=(empty2? DomCells[i]: cmin[i]);
        // NOTE (reference quirk, S/JCellDivCpu.cpp:295-296): cmin is replaced before cmax is tested, so cmax is only
        // zeroed when the replaced cmin.x still exceeds cmax.x.
        cmin[i]=a;
      }
      const bool empty3=(cmin[0]>cmax[0]);
      for(int i=0;i<3;i++){
        const unsigned b=(empty3? 0: cmax[i]);
        cellmin[i]=std::min(cellmin[i],cmin[i]); cellmax[i]=std::max(cellmax[i],b);
      }
    }
  }
  // ---- JCellDivCpuSingle::CalcCellDomain + MergeMapCellBoundFluid S/JCellDivCpuSingle.cpp:47-100
  void CalcCellDomain(){
    unsigned celbmin[3],celbmax[3];
    if(!BoundLimitOk){
      CalcCellDomainRange(Npb1,0,Npb2,Npb1+Npf1,celbmin,celbmax);
      BoundLimitOk=true; for(int i=0;i<3;i++){ BoundLimitCellMin[i]=celbmin[i]; BoundLimitCellMax[i]=celbmax[i]; }
    }
    else for(int i=0;i<3;i++){ celbmin[i]=BoundLimitCellMin[i]; celbmax[i]=BoundLimitCellMax[i]; }
    unsigned celfmin[3],celfmax[3];
    CalcCellDomainRange(Npf1,Npb1,Npf2,Npb1+Npf1+Npb2,celfmin,celfmax);
    const unsigned sd=unsigned(ScellDiv);
    for(int i=0;i<3;i++){
      CellDomainMin[i]=std::max(std::min(celbmin[i],celfmin[i]),(celfmin[i]>=sd? celfmin[i]-sd: 0u));
      CellDomainMax[i]=std::min(std::max(celbmax[i],celfmax[i]),celfmax[i]+sd);
      if(CellDomainMax[i]>=DomCells[i])CellDomainMax[i]=DomCells[i]-1;
    }
    if(CellDomainMin[0]>CellDomainMax[0]||CellDomainMin[1]>CellDomainMax[1]||CellDomainMin[2]>CellDomainMax[2])
      for(int i=0;i<3;i++)CellDomainMin[i]=CellDomainMax[i]=0;
  }
  // ---- JCellDivCpuSingle::PrepareNct S/JCellDivCpuSingle.cpp:109-125
  void PrepareNct(){
    Ncx=CellDomainMax[0]-CellDomainMin[0]+1; Ncy=CellDomainMax[1]-CellDomainMin[1]+1; Ncz=CellDomainMax[2]-CellDomainMin[2]+1;
    Nsheet=Ncx*Ncy; Nct=Nsheet*Ncz; Nctt=Nct*2+6;   // SizeBeginCell(nct)=nct*2+5+1, S/JCellDivCpu.h:141
    BoxBoundIgnore=Nct; BoxFluid=BoxBoundIgnore+1; BoxBoundOut=BoxFluid+Nct; BoxFluidOut=BoxBoundOut+1;
    BoxBoundOutIgnore=BoxFluidOut+1; BoxFluidOutIgnore=BoxBoundOutIgnore+1;
  }
  // ---- PreSortFull/PreSortFluid/MakeSortFull/MakeSortFluid S/JCellDivCpuSingle.cpp:134-234
  unsigned BoxOf(unsigned p,bool fluidonly)const{
    const unsigned rcell=Dcell[p];
    const unsigned cx=DcCellx(DomCellCode,rcell)-CellDomainMin[0], cy=DcCelly(DomCellCode,rcell)-CellDomainMin[1], cz=DcCellz(DomCellCode,rcell)-CellDomainMin[2];
    const unsigned cellsort=cx+cy*Ncx+cz*Nsheet;
    const typecode rcode=Code[p], codetype=CodeType(rcode), codeout=CodeSpecial(rcode);
    if(!fluidonly && codetype<DSPH_CODE_TYPE_FLOATING)
      return (codeout<DSPH_CODE_OUTIGNORE? ((cx<Ncx && cy<Ncy && cz<Ncz)? cellsort: BoxBoundIgnore): (codeout==DSPH_CODE_OUTIGNORE? BoxBoundOutIgnore: BoxBoundOut));
    return (codeout<=DSPH_CODE_OUTIGNORE? (codeout<DSPH_CODE_OUTIGNORE? BoxFluid+cellsort: BoxFluidOutIgnore): (codetype==DSPH_CODE_TYPE_FLOATING? BoxBoundOut: BoxFluidOut));
  }
  void PreSort(){
    if(DivideFull){
      std::fill(PartsInCell.begin(),PartsInCell.begin()+(Nctt-1),0u);
      for(unsigned p=0;p<Nptot;p++){ const unsigned box=BoxOf(p,false); CellPart[p]=box; PartsInCell[box]++; }
      BeginCell[0]=0;
      for(unsigned box=0;box<Nctt-1;box++)BeginCell[box+1]=BeginCell[box]+PartsInCell[box];
      std::fill(PartsInCell.begin(),PartsInCell.begin()+(Nctt-1),0u);
      for(unsigned p=0;p<Nptot;p++){ const unsigned box=CellPart[p]; SortPart[BeginCell[box]+PartsInCell[box]]=p; PartsInCell[box]++; }
    }
    else{
      std::fill(PartsInCell.begin()+BoxFluid,PartsInCell.begin()+(Nctt-1),0u);
      for(unsigned p=Npb1;p<Npb1+Npf1;p++){ const unsigned box=BoxOf(p,true); CellPart[p]=box; PartsInCell[box]++; }
      for(unsigned box=BoxFluid;box<Nctt-1;box++)BeginCell[box+1]=BeginCell[box]+PartsInCell[box];
      std::fill(PartsInCell.begin()+BoxFluid,PartsInCell.begin()+(Nctt-1),0u);
      for(unsigned p=Npb1;p<Npb1+Npf1;p++){ const unsigned box=CellPart[p]; SortPart[BeginCell[box]+PartsInCell[box]]=p; PartsInCell[box]++; }
    }
  }
  template<class T> void SortArray(std::vector<T> &v)const{                         // S/JCellDivCpu.cpp:352-450
    const unsigned ini=(DivideFull? 0: NpbFinal);
    std::vector<T> tmp(Nptot-ini);
    for(unsigned p=ini;p<Nptot;p++)tmp[p-ini]=v[SortPart[p]];
    std::copy(tmp.begin(),tmp.end(),v.begin()+ini);
  }
  unsigned CellSize(unsigned box)const{ return BeginCell[box+1]-BeginCell[box]; }
  // ---- JCellDivCpuSingle::Divide S/JCellDivCpuSingle.cpp:259-345
  void Divide(bool boundchanged){
    DivideFull=false;
    Npb1=Npb; Npf1=Np-Npb-NpbPer-NpfPer; Npb2=NpbPer; Npf2=NpfPer;
    Nptot=Npb1+Npf1+Npb2+Npf2;
    NpbOut=NpfOut=NpbOutIgnore=NpfOutIgnore=0; NpFinal=NpbFinal=0; NpbIgnore=0;
    if(CellPart.size()<Nptot){ CellPart.resize(Nptot+Nptot/20+128); SortPart.resize(CellPart.size()); BoundDivideOk=false; } // CheckMemoryNp -> AllocMemoryNp resets BoundDivideOk
    if(boundchanged || PeriActive || Mgpu){ BoundLimitOk=BoundDivideOk=false; for(int i=0;i<3;i++)BoundLimitCellMin[i]=BoundLimitCellMax[i]=BoundDivideCellMin[i]=BoundDivideCellMax[i]=0; }
    CalcCellDomain();
    PrepareNct();
    if(BeginCell.size()<Nctt){ BeginCell.assign(Nctt,0u); PartsInCell.assign(Nctt,0u); BoundDivideOk=false; }   // CheckMemoryNct
    bool same=true; for(int i=0;i<3;i++)if(BoundDivideCellMin[i]!=CellDomainMin[i]||BoundDivideCellMax[i]!=CellDomainMax[i])same=false;
    if(!BoundDivideOk || !same){
      DivideFull=true; BoundDivideOk=true;
      for(int i=0;i<3;i++){ BoundDivideCellMin[i]=CellDomainMin[i]; BoundDivideCellMax[i]=CellDomainMax[i]; }
    }
    PreSort();
    // (the reference also reorders CellPart here, S/JCellDivCpuSingle.cpp:256; it is not used afterwards on the CPU)
    NpbIgnore=CellSize(BoxBoundIgnore); NpbOut=CellSize(BoxBoundOut); NpfOut=CellSize(BoxFluidOut);
    NpbOutIgnore=CellSize(BoxBoundOutIgnore); NpfOutIgnore=CellSize(BoxFluidOutIgnore);
    NpFinal=Nptot-NpbOut-NpfOut-NpbOutIgnore-NpfOutIgnore;
    NpbFinal=Npb1+Npb2-NpbOutIgnore;
    if(NpbOut!=0 && DivideFull)NpbFinal=UINT_MAX;
  }
  // ---- JSphCpuSingle::RunCellDivide S/JSphCpuSingle.cpp:437-500
  void RunCellDivide(bool updateperiodic){
    if(updateperiodic && PeriActive && !Mgpu)RunPeriodic();
    Divide(BoundChanged);
    SortArray(Idp); SortArray(Code); SortArray(Dcell); SortArray(Pos); SortArray(Velrhop);
    if(TStep==DSPH_STEP_VERLET)SortArray(VelrhopM1);
    else if(TStep==DSPH_STEP_SYMPLECTIC && HasPre){ SortArray(PosPre); SortArray(VelrhopPre); }
    if(TVisco==DSPH_VISCO_LAMINARSPS)SortArray(SpsTau);
    if(Mdbc && BoundNormal.size()>=Np)SortArray(BoundNormal);                           // S/JSphCpuSingle.cpp:465
    Np=NpFinal; Npb=NpbFinal; NpbOk=Npb-NpbIgnore;
    if(NpbOut)throw std::runtime_error("Boundary particles out of the domain (AbortBoundOut).");
    BoundChanged=false;
  }

  // ---- kernels: S/FunSphKernel.h:86-151 (cubic), :212-222 (wendland)
  inline float FacWendland(float rr2)const{
    const float rad=sqrtf(rr2); const float qq=rad/KernelH; const float wqq1=1.f-0.5f*qq;
    return Bwen*qq*wqq1*wqq1*wqq1/rad;
  }
  inline float WabCubic(float rr2)const{
    const float rad=sqrtf(rr2); const float qq=rad/KernelH;
    if(rad>KernelH){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; return CubA24*(wqq2*wqq1); }
    const float wqq2=qq*qq; return CubA2*(1.0f+(0.75f*qq-1.5f)*wqq2);
  }
  inline float FacCubic(float rr2)const{
    const float rad=sqrtf(rr2); const float qq=rad/KernelH;
    if(rad>KernelH){ const float wqq1=2.0f-qq; const float wqq2=wqq1*wqq1; return CubC2*wqq2/rad; }
    const float wqq2=qq*qq; return (CubC1*qq+CubD1*wqq2)/rad;
  }
  inline float TensilCubic(float rr2,float rhopp1,float pressp1,float rhopp2,float pressp2)const{
    const float wab=WabCubic(rr2); float fab=wab*CubOdWdeltap; fab*=fab; fab*=fab;
    const float t1=(pressp1/(rhopp1*rhopp1))*(pressp1>0? 0.01f: -0.2f);
    const float t2=(pressp2/(rhopp2*rhopp2))*(pressp2>0? 0.01f: -0.2f);
    return fab*(t1+t2);
  }
  inline float Fac(float rr2)const{ return TKernel==DSPH_KERNEL_WENDLAND? FacWendland(rr2): FacCubic(rr2); }

  // ---- neighbour search: nsearch::Init / ParticleRange S/JCellSearch_inline.h:33-83
  struct NgSearch{ int cellinit,cxini,cxfin,yini,yfin,zini,zfin; };
  inline NgSearch NgInit(unsigned rcell,bool boundp2)const{
    const int cx=int(DcCellx(DomCellCode,rcell))-int(CellDomainMin[0]);
    const int cy=int(DcCelly(DomCellCode,rcell))-int(CellDomainMin[1]);
    const int cz=int(DcCellz(DomCellCode,rcell))-int(CellDomainMin[2]);
    const int ncx=int(Ncx),ncy=int(Ncy),ncz=int(Ncz),sd=ScellDiv;
    NgSearch r; r.cellinit=(boundp2? 0: int(Nsheet*Ncz+1));
    r.cxini=cx-(cx<sd? cx: sd); r.cxfin=cx+(ncx-cx-1<sd? ncx-cx-1: sd)+1;
    r.yini=cy-(cy<sd? cy: sd);  r.yfin=cy+(ncy-cy-1<sd? ncy-cy-1: sd)+1;
    r.zini=cz-(cz<sd? cz: sd);  r.zfin=cz+(ncz-cz-1<sd? ncz-cz-1: sd)+1;
    return r;
  }

  // ---- JSphCpu::PreInteractionVars_Forces + PreInteraction_Forces S/JSphCpu.cpp:433-479
  void PreInteraction(){
    Ar.assign(Np,0.f); Ace.assign(Np,f3{0,0,0});
    if(DDTArray)Delta.assign(Np,0.f); else Delta.clear();
    if(ShiftMode)ShiftPosfs.assign(Np,f4{0,0,0,0}); else ShiftPosfs.clear();   // no shifting zones: JSphShifting::InitCpu S/JSphShifting.cpp:345-347
    if(TVisco==DSPH_VISCO_LAMINARSPS)SpsGradvel.assign(Np,sym3{0,0,0,0,0,0});
    Press.resize(Np);
    // fsph::ComputePressMonaghan S/FunSphEos.h:38-40: b*(pow(rhop/rhop0,gamma)-1.0f).  Inside namespace fsph only the C
    // ::pow(double,double) is visible, so the reference evaluates pow, the -1 and the *b in DOUBLE and rounds to float once;
    // its -ffast-math build turns rhop/rhop0 into the float product rhop*(1.0f/rhop0) (verified in the disassembly of
    // JSphCpu::PreInteractionVars_Forces, oracle/_ref/obj52/JSphCpu.o).  Both facts are reproduced here.
    const float ovrhopzero=1.0f/RhopZero;
    // One more fact of the binary, pinned by the pressure dumps of oracle/ref_harness.cpp: the loop is vectorised 4-wide with a
    // scalar epilogue ("peeling for gaps" of the strided .w load) that handles the last Np-4*((Np-1)/4) particles, and the
    // epilogue keeps the true division rhop/rhop0.  One ulp of that quotient is 0.13 Pa, visible in Ace of free-surface particles.
    // The epilogue depends on the particle index and count (non-physical, and it differs between a single domain and x-slabs), so
    // the product keeps ONE rounding rule (the product form).  The emulation is on by default here because it is what pins this
    // restatement to the reference dumps; tests that compare the CUDA path with this checker switch it off (ORC_REFBINARY_TAIL=0).
    const char *envtail=getenv("ORC_REFBINARY_TAIL");
    const bool reftail=!(envtail && envtail[0]=='0');
    const unsigned nvec=(reftail? (Np? ((Np-1)/4)*4: 0): Np);
    if(!MultiPhase)for(unsigned p=0;p<Np;p++){
      const float x=(p<nvec? Velrhop[p].w*ovrhopzero: Velrhop[p].w/RhopZero);
      Press[p]=float(double(CteB)*(pow(double(x),double(Gamma))-1.0));
    }
    else{
      // M/JSphCpu.cpp:516-546 (extra arrays) + JSphCpu::ComputePress_NN M/JSphCpu_Tensors.cpp:47-71 (float pow, per-phase EOS)
      if(TVisco!=DSPH_VISCO_ARTIFICIAL)SpsGradvel.assign(Np,sym3{0,0,0,0,0,0}); else SpsGradvel.clear();
      SpsTau.assign(Np,sym3{0,0,0,0,0,0}); Dtensor.assign(Np,sym3{0,0,0,0,0,0}); ViscoEta.assign(Np,0.f);
      for(unsigned p=0;p<Np;p++){
        float rhozero=RhopZero,cteb=CteB,gamma=Gamma;
        if(CodeIsFluid(Code[p])){ const Phase &ph=Ph[CodeTypeValue(Code[p])]; rhozero=ph.rho; cteb=ph.cteb; gamma=ph.gamma; }
        const float rhop_r0=Velrhop[p].w/rhozero;
        Press[p]=cteb*(powf(rhop_r0,gamma)-1.0f);
      }
      ViscEtaDtMax=0;
    }
    const unsigned pini=(cfg.dtallparticles? 0: Npb);
    float vmax=0;
    for(unsigned p=pini;p<Np;p++){ const f4 v=Velrhop[p]; const float v2=v.x*v.x+v.y*v.y+v.z*v.z; if(vmax<v2)vmax=v2; }
    VelMax=sqrtf(vmax); LastVel2=vmax;                                                        // CalcVelMaxSeq/Omp S/JSphCpu.cpp:485-527 (float sqrt)
    ViscDtMax=0;
  }

  // ---- JSphCpu::InteractionForcesFluid S/JSphCpu.cpp:631-822 (FtMode none, no symmetry)
  void InteractionForcesFluid(bool boundp2,float visco,float &viscdt){
    const int pinit=int(Npb), pfin=int(Np);
    const bool shift=(ShiftMode!=0);
    float vmaxall=viscdt; unsigned long long nreal=0,nchk=0;
    #pragma omp parallel for schedule(guided) reduction(max:vmaxall) reduction(+:nreal,nchk)
    for(int p1=pinit;p1<pfin;p1++){
      if(Mgpu && CodeIsPeriodic(Code[p1]))continue;   // halo copies: forces are computed by the owning rank
      float visc=0,arp1=0,deltap1=0; f3 acep1={0,0,0}; sym3 gradvelp1={0,0,0,0,0,0};
      f4 shiftposfsp1={0,0,0,0}; if(shift)shiftposfsp1=ShiftPosfs[p1];
      const d3 posp1=Pos[p1]; const f3 velp1={Velrhop[p1].x,Velrhop[p1].y,Velrhop[p1].z};
      const float rhopp1=Velrhop[p1].w, pressp1=Press[p1];
      const sym3 taup1=(TVisco==DSPH_VISCO_ARTIFICIAL? gradvelp1: SpsTau[p1]);
      const NgSearch ngs=NgInit(Dcell[p1],boundp2);
      for(int z=ngs.zini;z<ngs.zfin;z++)for(int y=ngs.yini;y<ngs.yfin;y++){
        const int v=int(Nsheet)*z+int(Ncx)*y+ngs.cellinit;
        const unsigned pini=BeginCell[v+ngs.cxini], pfin2=BeginCell[v+ngs.cxfin];
        nchk+=pfin2-pini;
        for(unsigned p2=pini;p2<pfin2;p2++){
          const float drx=float(posp1.x-Pos[p2].x), dry=float(posp1.y-Pos[p2].y), drz=float(posp1.z-Pos[p2].z);
          const float rr2=drx*drx+dry*dry+drz*drz;
          if(rr2<=KernelSize2 && rr2>=1e-18f){
            nreal++;
            const float fac=Fac(rr2);
            const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
            const float massp2=(boundp2? MassBound: MassFluid);
            const f4 velrhop2=Velrhop[p2];
            { //-Momentum
              const float prs=(pressp1+Press[p2])/(rhopp1*velrhop2.w)+(TKernel==DSPH_KERNEL_CUBIC? TensilCubic(rr2,rhopp1,pressp1,velrhop2.w,Press[p2]): 0);
              const float p_vpm=-prs*massp2;
              acep1.x+=p_vpm*frx; acep1.y+=p_vpm*fry; acep1.z+=p_vpm*frz;
            }
            //-Continuity
            const float dvx=velp1.x-velrhop2.x, dvy=velp1.y-velrhop2.y, dvz=velp1.z-velrhop2.z;
            arp1+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(rhopp1/velrhop2.w);
            const float cbar=(float)Cs0;
            if(TDensity==DSPH_DDT_DDT && deltap1!=FLT_MAX){             // Molteni & Colagrossi
              // rho1/rho2-1 cancels to ~1e-4, so one ulp of the quotient is visible in Ar (5e-6 .. 5e-5).  The reference binary
              // (-ffast-math: reciprocal-math) evaluates rho1*(1/rho2) in exactly one template instantiation family, cubic kernel +
              // laminar/SPS (3 divisions by rho2 there), and the true quotient in all others: measured on 16 instantiations
              // (kernel x viscosity x shifting x step algorithm), Ar agrees to 1.5e-7 with the matching form and 5e-6 with the other.
              const float rhop1over2=(TKernel==DSPH_KERNEL_CUBIC && TVisco==DSPH_VISCO_LAMINARSPS? rhopp1*(1.0f/velrhop2.w): rhopp1/velrhop2.w);
              const float visc_densi=DDTkh*cbar*(rhop1over2-1.f)/(rr2+Eta2);
              const float dot3=(drx*frx+dry*fry+drz*frz);
              const float delta=visc_densi*dot3*massp2;
              deltap1=(boundp2 && !Mdbc? FLT_MAX: deltap1+delta);      // boundp2 && TBoundary==BC_DBC (S/JSphCpu.cpp:730)
            }
            if((TDensity==DSPH_DDT_DDT2 || (TDensity==DSPH_DDT_DDT2FULL && !boundp2)) && deltap1!=FLT_MAX){   // Fourtakas
              const float rh=1.f+DDTgz*drz;
              // source: RhopZero*pow(rh,1.f/Gamma)-RhopZero; the reference's -ffast-math build factors it as
              // RhopZero*(pow-1) (pinned by the dumps: Ar agrees to 2e-7 with this form, 3e-6..7e-6 with the literal one)
              const float drhop=RhopZero*(powf(rh,1.f/Gamma)-1.f);
              const float visc_densi=DDTkh*cbar*((velrhop2.w-rhopp1)-drhop)/(rr2+Eta2);
              const float dot3=(drx*frx+dry*fry+drz*frz);
              const float delta=visc_densi*dot3*massp2/velrhop2.w;
              deltap1=(boundp2? FLT_MAX: deltap1-delta);
            }
            if(shift && shiftposfsp1.x!=FLT_MAX){
              const float massrhop=massp2/velrhop2.w;
              const bool noshift=(boundp2 && (ShiftMode==DSPH_SHIFT_NOBOUND || (ShiftMode==DSPH_SHIFT_NOFIXED && CodeIsFixed(Code[p2]))));
              shiftposfsp1.x=(noshift? FLT_MAX: shiftposfsp1.x+massrhop*frx);
              shiftposfsp1.y+=massrhop*fry; shiftposfsp1.z+=massrhop*frz;
              shiftposfsp1.w-=massrhop*(drx*frx+dry*fry+drz*frz);
            }
            { //-Viscosity
              const float dot=drx*dvx+dry*dvy+drz*dvz;
              const float dot_rr2=dot/(rr2+Eta2);
              visc=std::max(dot_rr2,visc);
              if(TVisco==DSPH_VISCO_ARTIFICIAL){
                if(dot<0){
                  const float amubar=KernelH*dot_rr2;
                  const float robar=(rhopp1+velrhop2.w)*0.5f;
                  const float pi_visc=(-visco*cbar*amubar/robar)*massp2;
                  acep1.x-=pi_visc*frx; acep1.y-=pi_visc*fry; acep1.z-=pi_visc*frz;
                }
              }
              else if(TVisco==DSPH_VISCO_LAMINARSPS){
                {
                  const float robar2=(rhopp1+velrhop2.w);
                  const float temp=4.f*visco/((rr2+Eta2)*robar2);
                  const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
                  acep1.x+=vtemp*dvx; acep1.y+=vtemp*dvy; acep1.z+=vtemp*dvz;
                }
                float tau_xx=taup1.xx,tau_xy=taup1.xy,tau_xz=taup1.xz,tau_yy=taup1.yy,tau_yz=taup1.yz,tau_zz=taup1.zz;
                if(!boundp2){
                  const sym3 t2=SpsTau[p2];
                  tau_xx+=t2.xx; tau_xy+=t2.xy; tau_xz+=t2.xz; tau_yy+=t2.yy; tau_yz+=t2.yz; tau_zz+=t2.zz;
                }
                acep1.x+=massp2*(tau_xx*frx+tau_xy*fry+tau_xz*frz);
                acep1.y+=massp2*(tau_xy*frx+tau_yy*fry+tau_yz*frz);
                acep1.z+=massp2*(tau_xz*frx+tau_yz*fry+tau_zz*frz);
                const float volp2=-massp2/velrhop2.w;
                float dv=dvx*volp2; gradvelp1.xx+=dv*frx; gradvelp1.xy+=dv*fry; gradvelp1.xz+=dv*frz;
                      dv=dvy*volp2; gradvelp1.xy+=dv*frx; gradvelp1.yy+=dv*fry; gradvelp1.yz+=dv*frz;
                      dv=dvz*volp2; gradvelp1.xz+=dv*frx; gradvelp1.yz+=dv*fry; gradvelp1.zz+=dv*frz;
              }
            }
          }
        }
      }
      if(shift||arp1||acep1.x||acep1.y||acep1.z||visc){
        if(TDensity!=DSPH_DDT_NONE){
          if(!Delta.empty())Delta[p1]=(Delta[p1]==FLT_MAX || deltap1==FLT_MAX? FLT_MAX: Delta[p1]+deltap1);
          else if(deltap1!=FLT_MAX)arp1+=deltap1;
        }
        Ar[p1]+=arp1;
        Ace[p1].x=Ace[p1].x+acep1.x; Ace[p1].y=Ace[p1].y+acep1.y; Ace[p1].z=Ace[p1].z+acep1.z;
        if(visc>vmaxall)vmaxall=visc;
        if(TVisco==DSPH_VISCO_LAMINARSPS){
          sym3 &g=SpsGradvel[p1];
          g.xx+=gradvelp1.xx; g.xy+=gradvelp1.xy; g.xz+=gradvelp1.xz; g.yy+=gradvelp1.yy; g.yz+=gradvelp1.yz; g.zz+=gradvelp1.zz;
        }
        if(shift)ShiftPosfs[p1]=shiftposfsp1;
      }
    }
    viscdt=vmaxall;
    if(!boundp2){ PipsFluidReal=nreal; PipsFluidChecked=nchk; } else { PipsFluidReal+=nreal; PipsFluidChecked+=nchk; }
  }
  // ---- JSphCpu::InteractionForcesBound S/JSphCpu.cpp:548-625
  void InteractionForcesBound(float &viscdt){
    const int n=int(NpbOk);
    float vmaxall=viscdt; unsigned long long nreal=0,nchk=0;
    #pragma omp parallel for schedule(guided) reduction(max:vmaxall) reduction(+:nreal,nchk)
    for(int p1=0;p1<n;p1++){
      float visc=0,arp1=0;
      const d3 posp1=Pos[p1]; const f4 velrhop1=Velrhop[p1];
      const NgSearch ngs=NgInit(Dcell[p1],false);
      for(int z=ngs.zini;z<ngs.zfin;z++)for(int y=ngs.yini;y<ngs.yfin;y++){
        const int v=int(Nsheet)*z+int(Ncx)*y+ngs.cellinit;
        const unsigned pini=BeginCell[v+ngs.cxini], pfin2=BeginCell[v+ngs.cxfin];
        nchk+=pfin2-pini;
        for(unsigned p2=pini;p2<pfin2;p2++){
          const float drx=float(posp1.x-Pos[p2].x), dry=float(posp1.y-Pos[p2].y), drz=float(posp1.z-Pos[p2].z);
          const float rr2=drx*drx+dry*dry+drz*drz;
          if(rr2<=KernelSize2 && rr2>=1e-18f){
            nreal++;
            const float fac=Fac(rr2);
            const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
            const float massp2=MassFluid;
            const f4 velrhop2=Velrhop[p2];
            const float dvx=velrhop1.x-velrhop2.x, dvy=velrhop1.y-velrhop2.y, dvz=velrhop1.z-velrhop2.z;
            arp1+=massp2*(dvx*frx+dvy*fry+dvz*frz)*(velrhop1.w/velrhop2.w);
            const float dot=drx*dvx+dry*dvy+drz*dvz;
            const float dot_rr2=dot/(rr2+Eta2);
            visc=std::max(dot_rr2,visc);
          }
        }
      }
      if(arp1||visc){ Ar[p1]+=arp1; if(visc>vmaxall)vmaxall=visc; }
    }
    viscdt=vmaxall; PipsBoundReal=nreal; PipsBoundChecked=nchk;
  }
  // ---- JSphCpu::ComputeSpsTau S/JSphCpu.cpp:930-955
  void ComputeSpsTau(){
    for(unsigned p=Npb;p<Np;p++){
      const sym3 g=SpsGradvel[p];
      const float pow1=g.xx*g.xx+g.yy*g.yy+g.zz*g.zz;
      const float prr=pow1+pow1+g.xy*g.xy+g.xz*g.xz+g.yz*g.yz;
      const float visc_sps=SpsSmag*sqrtf(prr);
      const float div_u=g.xx+g.yy+g.zz;
      const float sps_k=(2.0f/3.0f)*visc_sps*div_u;
      const float sps_blin=SpsBlin*prr;
      const float sumsps=-(sps_k+sps_blin);
      const float twovisc_sps=(visc_sps+visc_sps);
      const float one_rho2=1.0f/Velrhop[p].w;
      sym3 &t=SpsTau[p];
      t.xx=one_rho2*(twovisc_sps*g.xx+sumsps); t.xy=one_rho2*(visc_sps*g.xy); t.xz=one_rho2*(visc_sps*g.xz);
      t.yy=one_rho2*(twovisc_sps*g.yy+sumsps); t.yz=one_rho2*(visc_sps*g.yz); t.zz=one_rho2*(twovisc_sps*g.zz+sumsps);
    }
  }
  // =====================================================================================================================
  // Multiphase non-Newtonian formulation (src_mphase fork, CPU path).  FtMode none, no symmetry, DBC boundaries.
  // =====================================================================================================================
  // neighbours of p1 in cell order (nsearch::Init/ParticleRange M/JCellSearch_inline.h): f(p2,drx,dry,drz,rr2) for every accepted pair
  template<class F> inline void ForPairs(unsigned p1,bool boundp2,F &&f)const{
    const d3 posp1=Pos[p1];
    const NgSearch ngs=NgInit(Dcell[p1],boundp2);
    for(int z=ngs.zini;z<ngs.zfin;z++)for(int y=ngs.yini;y<ngs.yfin;y++){
      const int v=int(Nsheet)*z+int(Ncx)*y+ngs.cellinit;
      const unsigned pini=BeginCell[v+ngs.cxini], pfin2=BeginCell[v+ngs.cxfin];
      for(unsigned p2=pini;p2<pfin2;p2++){
        const float drx=float(posp1.x-Pos[p2].x), dry=float(posp1.y-Pos[p2].y), drz=float(posp1.z-Pos[p2].z);
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=KernelSize2 && rr2>=1e-18f)f(p2,drx,dry,drz,rr2);
      }
    }
  }
  // ---- JSphCpu::GetEta_Effective M/JSphCpu_Tensors.cpp:113-141
  inline float EtaEffective(unsigned ppx,float tau_yield,float dmagn,float visco,float m_nn,float n_nn)const{
    const Phase &P=Ph[ppx];
    if(dmagn<=1e-18f)dmagn=1e-18f;
    float miou_yield=(P.tau_max? P.tau_max/(2.0f*dmagn): tau_yield/(2.0f*dmagn));
    const bool bi_region=P.tau_max && dmagn<=P.tau_max/(2.f*P.bi_multi*visco);
    if(bi_region)miou_yield=P.bi_multi*visco;
    const float miouPap=miou_yield*(1.f-expf(-m_nn*dmagn));
    const bool capped=(miouPap>m_nn*tau_yield || dmagn==1e-18f);
    const float term1=(P.tau_max? miou_yield: (capped? m_nn*tau_yield: miouPap));
    const float miouHB=visco*powf(dmagn,(n_nn-1.0f));
    const float term2=(bi_region? visco: (capped? visco: miouHB));
    return term1+term2;
  }
  // ---- shared by FDA and SPH-PressGrad: pressure, continuity, DDT and shifting of one pair
  //      M/JSphCpu_NN_FDA.cpp:205-251 == M/JSphCpu_NN_SPH.cpp:528-581
  struct NNAcc{ float visc,arp1,deltap1; f3 acep1; f4 shiftposfsp1; };
  inline void NNPairBasic(bool boundp2,bool shift,unsigned pp1,unsigned pp2,unsigned p2,float massp2,float rhopp1,float pressp1,const f3 &velp1,
    float drx,float dry,float drz,float rr2,float frx,float fry,float frz,const f4 &velrhop2,float &dvx,float &dvy,float &dvz,NNAcc &a)const
  {
    {
      const float prs=(pressp1+Press[p2])/(rhopp1*velrhop2.w)+(TKernel==DSPH_KERNEL_CUBIC? TensilCubic(rr2,rhopp1,pressp1,velrhop2.w,Press[p2]): 0);
      const float p_vpm=-prs*massp2;
      a.acep1.x+=p_vpm*frx; a.acep1.y+=p_vpm*fry; a.acep1.z+=p_vpm*frz;
    }
    const float rhop1over2=rhopp1/velrhop2.w;
    dvx=velp1.x-velrhop2.x; dvy=velp1.y-velrhop2.y; dvz=velp1.z-velrhop2.z;
    a.arp1+=massp2*(dvx*frx+dvy*fry+dvz*frz)*rhop1over2;
    const float cbar=Ph[pp2].cs0;                                   // max(Cs0[pp2],Cs0[pp2])
    if(TDensity==DSPH_DDT_DDT && a.deltap1!=FLT_MAX){
      const float visc_densi=DDTkh*cbar*(rhop1over2-1.f)/(rr2+Eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      const float delta=(pp1==pp2? visc_densi*dot3*massp2: 0);
      a.deltap1=(boundp2? FLT_MAX: a.deltap1+delta);                // TBoundary==BC_DBC
    }
    if((TDensity==DSPH_DDT_DDT2 || (TDensity==DSPH_DDT_DDT2FULL && !boundp2)) && a.deltap1!=FLT_MAX){
      const float rh=1.f+DDTgz*drz;
      const float drhop=RhopZero*(powf(rh,1.f/Gamma)-1.f);          // fast-math factoring of RhopZero*pow-RhopZero, as above
      const float visc_densi=DDTkh*cbar*((velrhop2.w-rhopp1)-drhop)/(rr2+Eta2);
      const float dot3=(drx*frx+dry*fry+drz*frz);
      const float delta=(pp1==pp2? visc_densi*dot3*massp2/velrhop2.w: 0);
      a.deltap1=(boundp2? FLT_MAX: a.deltap1-delta);
    }
    if(shift && a.shiftposfsp1.x!=FLT_MAX){
      const bool heavyphase=(Ph[pp1].mass>Ph[pp2].mass && pp1!=pp2);
      const float massrhop=massp2/velrhop2.w;
      const bool noshift=(boundp2 && (ShiftMode==DSPH_SHIFT_NOBOUND || (ShiftMode==DSPH_SHIFT_NOFIXED && CodeIsFixed(Code[p2]))));
      a.shiftposfsp1.x=(noshift? FLT_MAX: (heavyphase? 0: a.shiftposfsp1.x+massrhop*frx));
      a.shiftposfsp1.y+=(heavyphase? 0: massrhop*fry);
      a.shiftposfsp1.z+=(heavyphase? 0: massrhop*frz);
      a.shiftposfsp1.w-=(heavyphase? 0: massrhop*(drx*frx+dry*fry+drz*frz));
    }
  }
  inline void NNStore(unsigned p1,bool shift,NNAcc &a,float &vmaxall){
    if(shift||a.arp1||a.acep1.x||a.acep1.y||a.acep1.z||a.visc){
      if(TDensity!=DSPH_DDT_NONE){
        if(!Delta.empty())Delta[p1]=(Delta[p1]==FLT_MAX || a.deltap1==FLT_MAX? FLT_MAX: Delta[p1]+a.deltap1);
        else if(a.deltap1!=FLT_MAX)a.arp1+=a.deltap1;
      }
      Ar[p1]+=a.arp1;
      Ace[p1].x=Ace[p1].x+a.acep1.x; Ace[p1].y=Ace[p1].y+a.acep1.y; Ace[p1].z=Ace[p1].z+a.acep1.z;
      if(a.visc>vmaxall)vmaxall=a.visc;
      if(shift)ShiftPosfs[p1]=a.shiftposfsp1;
    }
  }
  // ---- JSphCpu::InteractionForcesFluid_NN_FDA_All M/JSphCpu_NN_FDA.cpp:129-340
  void NNForcesFluidFDA(bool boundp2,float &viscdt,float &viscetadt){
    const int pinit=int(Npb), pfin=int(Np);
    const bool shift=(ShiftMode!=0);
    float vmaxall=viscdt, etamaxall=viscetadt;
    #pragma omp parallel for schedule(guided) reduction(max:vmaxall,etamaxall)
    for(int p1=pinit;p1<pfin;p1++){
      if(Mgpu && CodeIsPeriodic(Code[p1]))continue;
      NNAcc a; a.visc=0; a.arp1=0; a.deltap1=0; a.acep1={0,0,0}; a.shiftposfsp1={0,0,0,0};
      if(shift)a.shiftposfsp1=ShiftPosfs[p1];
      const f3 velp1={Velrhop[p1].x,Velrhop[p1].y,Velrhop[p1].z};
      const float rhopp1=Velrhop[p1].w, pressp1=Press[p1];
      const unsigned pp1=CodeTypeValue(Code[p1]);
      float visco_etap1=0, visceta=0;
      ForPairs(unsigned(p1),boundp2,[&](unsigned p2,float drx,float dry,float drz,float rr2){
        const float fac=Fac(rr2);
        const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
        const unsigned pp2=(boundp2? pp1: CodeTypeValue(Code[p2]));
        const float massp2=(boundp2? MassBound: Ph[pp2].mass);
        const f4 velrhop2=Velrhop[p2];
        float dvx,dvy,dvz;
        NNPairBasic(boundp2,shift,pp1,pp2,p2,massp2,rhopp1,pressp1,velp1,drx,dry,drz,rr2,frx,fry,frz,velrhop2,dvx,dvy,dvz,a);
        const float cbar=Ph[pp2].cs0;
        const float dot=drx*dvx+dry*dvy+drz*dvz;
        const float dot_rr2=dot/(rr2+Eta2);
        a.visc=std::max(dot_rr2,a.visc);
        const float visco_NN=Ph[pp2].visco;
        if(TVisco==DSPH_VISCO_ARTIFICIAL){
          if(dot<0){
            const float amubar=KernelH*dot_rr2;
            const float robar=(rhopp1+velrhop2.w)*0.5f;
            const float pi_visc=(-visco_NN*cbar*amubar/robar)*massp2;
            a.acep1.x-=pi_visc*frx; a.acep1.y-=pi_visc*fry; a.acep1.z-=pi_visc*frz;
          }
        }
        else{ // VISCO_LaminarSPS || VISCO_ConstEq
          if(boundp2){ dvx=2.f*velp1.x; dvy=2.f*velp1.y; dvz=2.f*velp1.z; }        // no-slip
          // GetVelocityGradients_FDA M/JSphCpu_Tensors.cpp:80-88
          const float a11=dvx*drx/rr2,a12=dvx*dry/rr2,a13=dvx*drz/rr2;
          const float a21=dvy*drx/rr2,a22=dvy*dry/rr2,a23=dvy*drz/rr2;
          const float a31=dvz*drx/rr2,a32=dvz*dry/rr2,a33=dvz*drz/rr2;
          const float div_vel=(a11+a22+a33)/3.f;
          // GetStrainRateTensor M/JSphCpu_Tensors.cpp:92-111
          const float d11=a11-div_vel, d12=0.5f*(a12+a21), d13=0.5f*(a13+a31);
          const float d21=0.5f*(a21+a12), d22=a22-div_vel, d23=0.5f*(a23+a32);
          const float d31=0.5f*(a31+a13), d32=0.5f*(a32+a23), d33=a33-div_vel;
          const float II_D_1=d11*d22+d22*d33+d11*d33;
          const float II_D_2=d12*d21+d23*d32+d13*d31;
          const float II_D=II_D_1-II_D_2;
          const float dmagn=sqrtf(II_D*II_D);
          visco_etap1=EtaEffective(pp1,Ph[pp2].tau_yield,dmagn,visco_NN,Ph[pp2].m_nn,Ph[pp2].n_nn);
          visceta=std::max(visco_etap1,visceta);
          if(TVisco==DSPH_VISCO_LAMINARSPS){                                        // Morris operator
            const float temp=2.0f*(visco_etap1)/((rr2+Eta2)*velrhop2.w);
            const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
            a.acep1.x+=vtemp*dvx; a.acep1.y+=vtemp*dvy; a.acep1.z+=vtemp*dvz;
          }
          else{                                                                     // GetStressTensor :143-164
            const float t2=2.f*visco_etap1;
            const float taux=((t2*d11)*frx+(t2*d12)*fry+(t2*d13)*frz)/(velrhop2.w);
            const float tauy=((t2*d21)*frx+(t2*d22)*fry+(t2*d23)*frz)/(velrhop2.w);
            const float tauz=((t2*d31)*frx+(t2*d32)*fry+(t2*d33)*frz)/(velrhop2.w);
            a.acep1.x+=taux*massp2; a.acep1.y+=tauy*massp2; a.acep1.z+=tauz*massp2;
          }
        }
      });
      const bool stored=(shift||a.arp1||a.acep1.x||a.acep1.y||a.acep1.z||a.visc);
      NNStore(unsigned(p1),shift,a,vmaxall);
      if(stored){
        if(visceta>etamaxall)etamaxall=visceta;
        if(visco_etap1)ViscoEta[p1]=visco_etap1;
      }
    }
    viscdt=vmaxall; viscetadt=etamaxall;
  }
  // ---- JSphCpu::InteractionForcesFluid_NN_SPH_PressGrad M/JSphCpu_NN_SPH.cpp:452-620
  void NNForcesFluidSPHPressGrad(bool boundp2,float &viscdt){
    const int pinit=int(Npb), pfin=int(Np);
    const bool shift=(ShiftMode!=0);
    float vmaxall=viscdt;
    #pragma omp parallel for schedule(guided) reduction(max:vmaxall)
    for(int p1=pinit;p1<pfin;p1++){
      NNAcc a; a.visc=0; a.arp1=0; a.deltap1=0; a.acep1={0,0,0}; a.shiftposfsp1={0,0,0,0};
      sym3 g={0,0,0,0,0,0};
      if(shift)a.shiftposfsp1=ShiftPosfs[p1];
      const f3 velp1={Velrhop[p1].x,Velrhop[p1].y,Velrhop[p1].z};
      const float rhopp1=Velrhop[p1].w, pressp1=Press[p1];
      const unsigned pp1=CodeTypeValue(Code[p1]);
      ForPairs(unsigned(p1),boundp2,[&](unsigned p2,float drx,float dry,float drz,float rr2){
        const float fac=Fac(rr2);
        const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
        const unsigned pp2=(boundp2? pp1: CodeTypeValue(Code[p2]));
        const float massp2=(boundp2? MassBound: Ph[pp2].mass);
        const f4 velrhop2=Velrhop[p2];
        float dvx,dvy,dvz;
        NNPairBasic(boundp2,shift,pp1,pp2,p2,massp2,rhopp1,pressp1,velp1,drx,dry,drz,rr2,frx,fry,frz,velrhop2,dvx,dvy,dvz,a);
        const float dot=drx*dvx+dry*dvy+drz*dvz;
        const float dot_rr2=dot/(rr2+Eta2);
        a.visc=std::max(dot_rr2,a.visc);
        if(TVisco!=DSPH_VISCO_ARTIFICIAL){
          if(boundp2){ dvx=2.f*velp1.x; dvy=2.f*velp1.y; dvz=2.f*velp1.z; }
          const float volp2=-massp2/velrhop2.w;                                   // GetVelocityGradients_SPH_tsym M/JSphCpu_Tensors.cpp:171-180
          float dv=dvx*volp2; g.xx+=dv*frx; g.xy+=dv*fry; g.xz+=dv*frz;
                dv=dvy*volp2; g.xy+=dv*frx; g.yy+=dv*fry; g.yz+=dv*frz;
                dv=dvz*volp2; g.xz+=dv*frx; g.yz+=dv*fry; g.zz+=dv*frz;
        }
      });
      const bool stored=(shift||a.arp1||a.acep1.x||a.acep1.y||a.acep1.z||a.visc);
      float vdummy=0;
      NNStore(unsigned(p1),shift,a,(Mgpu && CodeIsPeriodic(Code[p1])? vdummy: vmaxall));   // halo copies do not take part in the maxima
      if(stored && TVisco!=DSPH_VISCO_ARTIFICIAL){
        sym3 &t=SpsGradvel[p1];
        t.xx+=g.xx; t.xy+=g.xy; t.xz+=g.xz; t.yy+=g.yy; t.yz+=g.yz; t.zz+=g.zz;
      }
    }
    viscdt=vmaxall;
  }
  // ---- JSphCpu::InteractionForcesFluid_NN_SPH_Visco_eta M/JSphCpu_NN_SPH.cpp:174-228 (+ GetStrainRateTensor_tsym Tensors:184-207)
  void NNViscoEtaSPH(float &viscetadt){
    float etamaxall=viscetadt;
    #pragma omp parallel for schedule(guided) reduction(max:etamaxall)
    for(int p1=int(Npb);p1<int(Np);p1++){
      const sym3 gv=SpsGradvel[p1];
      const unsigned pp1=CodeTypeValue(Code[p1]);
      const float div_vel=(gv.xx+gv.yy+gv.zz)/3.f;
      sym3 D; D.xx=gv.xx-div_vel; D.xy=0.5f*(gv.xy); D.xz=0.5f*(gv.xz); D.yy=gv.yy-div_vel; D.yz=0.5f*(gv.yz); D.zz=gv.zz-div_vel;
      const float II_D_1=D.xx*D.yy+D.yy*D.zz+D.xx*D.zz;
      const float II_D_2=D.xy*D.xy+D.yz*D.yz+D.xz*D.xz;
      const float II_D=-II_D_1+II_D_2;
      const float dmagn=sqrtf(II_D);
      const float eta=EtaEffective(pp1,Ph[pp1].tau_yield,dmagn,Ph[pp1].visco,Ph[pp1].m_nn,Ph[pp1].n_nn);
      if(eta)ViscoEta[p1]=eta;
      if(TVisco!=DSPH_VISCO_ARTIFICIAL)Dtensor[p1]=D;
      if(eta>etamaxall && !(Mgpu && CodeIsPeriodic(Code[p1])))etamaxall=eta;
    }
    viscetadt=etamaxall;
  }
  // ---- JSphCpu::InteractionForcesFluid_NN_SPH_Morris M/JSphCpu_NN_SPH.cpp:337-446
  void NNForcesFluidSPHMorris(bool boundp2){
    #pragma omp parallel for schedule(guided)
    for(int p1=int(Npb);p1<int(Np);p1++){
      f3 acep1={0,0,0};
      const f3 velp1={Velrhop[p1].x,Velrhop[p1].y,Velrhop[p1].z};
      const float rhopp1=Velrhop[p1].w;
      const float visco_etap1=ViscoEta[p1];
      const unsigned pp1=CodeTypeValue(Code[p1]);
      ForPairs(unsigned(p1),boundp2,[&](unsigned p2,float drx,float dry,float drz,float rr2){
        const float fac=Fac(rr2);
        const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
        const unsigned pp2=(boundp2? pp1: CodeTypeValue(Code[p2]));
        const float massp2=(boundp2? MassBound: Ph[pp2].mass);
        const f4 velrhop2=Velrhop[p2];
        float dvx=velp1.x-velrhop2.x,dvy=velp1.y-velrhop2.y,dvz=velp1.z-velrhop2.z;
        if(boundp2){ dvx=2.f*velp1.x; dvy=2.f*velp1.y; dvz=2.f*velp1.z; }
        const float cbar=Ph[pp2].cs0;
        const float dot=drx*dvx+dry*dvy+drz*dvz;
        const float dot_rr2=dot/(rr2+Eta2);
        const float visco_NN=Ph[pp2].visco;
        if(TVisco==DSPH_VISCO_ARTIFICIAL){
          if(dot<0){
            const float amubar=KernelH*dot_rr2;
            const float robar=(rhopp1+velrhop2.w)*0.5f;
            const float pi_visc=(-visco_NN*cbar*amubar/robar)*massp2;
            acep1.x-=pi_visc*frx; acep1.y-=pi_visc*fry; acep1.z-=pi_visc*frz;
          }
        }
        else if(TVisco==DSPH_VISCO_LAMINARSPS){
          float visco_etap2=ViscoEta[p2];
          if(boundp2)visco_etap2=visco_etap1;
          const float temp=(visco_etap1+visco_etap2)/((rr2+Eta2)*velrhop2.w);
          const float vtemp=massp2*temp*(drx*frx+dry*fry+drz*frz);
          acep1.x+=vtemp*dvx; acep1.y+=vtemp*dvy; acep1.z+=vtemp*dvz;
        }
      });
      if(acep1.x||acep1.y||acep1.z){ Ace[p1].x=Ace[p1].x+acep1.x; Ace[p1].y=Ace[p1].y+acep1.y; Ace[p1].z=Ace[p1].z+acep1.z; }
    }
  }
  // ---- JSphCpu::InteractionForcesFluid_NN_SPH_Visco_Stress_tensor M/JSphCpu_NN_SPH.cpp:128-169 (+ GetStressTensor_sym Tensors:211-231)
  void NNStressTensorSPH(){
    for(unsigned p1=Npb;p1<Np;p1++){
      const float t2=2.f*ViscoEta[p1]; const sym3 D=Dtensor[p1];
      if(TVisco!=DSPH_VISCO_ARTIFICIAL)SpsTau[p1]=sym3{t2*D.xx,t2*D.xy,t2*D.xz,t2*D.yy,t2*D.yz,t2*D.zz};
    }
  }
  // ---- JSphCpu::InteractionForcesFluid_NN_SPH_ConsEq M/JSphCpu_NN_SPH.cpp:234-331
  void NNForcesFluidSPHConsEq(bool boundp2){
    #pragma omp parallel for schedule(guided)
    for(int p1=int(Npb);p1<int(Np);p1++){
      f3 acep1={0,0,0};
      const sym3 tau1=SpsTau[p1];
      const unsigned pp1=CodeTypeValue(Code[p1]);
      ForPairs(unsigned(p1),boundp2,[&](unsigned p2,float drx,float dry,float drz,float rr2){
        const float fac=Fac(rr2);
        const float frx=fac*drx,fry=fac*dry,frz=fac*drz;
        const unsigned pp2=(boundp2? pp1: CodeTypeValue(Code[p2]));
        const float massp2=(boundp2? MassBound: Ph[pp2].mass);
        const f4 velrhop2=Velrhop[p2];
        sym3 tau2=SpsTau[p2]; if(boundp2)tau2=tau1;
        const sym3 ts={tau1.xx+tau2.xx,tau1.xy+tau2.xy,tau1.xz+tau2.xz,tau1.yy+tau2.yy,tau1.yz+tau2.yz,tau1.zz+tau2.zz};
        const float taux=(ts.xx*frx+ts.xy*fry+ts.xz*frz)/(velrhop2.w);
        const float tauy=(ts.xy*frx+ts.yy*fry+ts.yz*frz)/(velrhop2.w);
        const float tauz=(ts.xz*frx+ts.yz*fry+ts.zz*frz)/(velrhop2.w);
        acep1.x+=taux*massp2; acep1.y+=tauy*massp2; acep1.z+=tauz*massp2;
      });
      if(acep1.x||acep1.y||acep1.z){ Ace[p1].x=Ace[p1].x+acep1.x; Ace[p1].y=Ace[p1].y+acep1.y; Ace[p1].z=Ace[p1].z+acep1.z; }
    }
  }
  // ---- Interaction_ForcesCpuT_NN_FDA M/JSphCpu_NN_FDA.cpp:346-379, Interaction_ForcesCpuT_NN_SPH M/JSphCpu_NN_SPH.cpp:654-710
  //      (InteractionForcesBound_NN_* are the classic boundary pass with MassFluid: M/JSphCpu_NN_FDA.cpp:46-124)
  void InteractionForcesNN(float &viscdt,float &viscetadt){
    if(Np>Npb){
      if(TVelGrad==DSPH_VELGRAD_FDA){
        NNForcesFluidFDA(false,viscdt,viscetadt);
        NNForcesFluidFDA(true,viscdt,viscetadt);
      }
      else{
        NNForcesFluidSPHPressGrad(false,viscdt);
        NNForcesFluidSPHPressGrad(true,viscdt);
        if(TVisco!=DSPH_VISCO_ARTIFICIAL)NNViscoEtaSPH(viscetadt);
        if(TVisco!=DSPH_VISCO_CONSTEQ){ NNForcesFluidSPHMorris(false); NNForcesFluidSPHMorris(true); }
        else{ NNStressTensorSPH(); NNForcesFluidSPHConsEq(false); NNForcesFluidSPHConsEq(true); }
      }
    }
  }

  // ---- JSphCpuSingle::Interaction_Forces S/JSphCpuSingle.cpp:524-567 + Interaction_ForcesCpuT S/JSphCpu.cpp:960-987
  // ---- gauges: JGaugeVelocity / JGaugeSwl / JGaugeMaxZ / JGaugeForce::CalculeCpu (S/JDsGaugeItem.cpp:325-367, 542-600, 780-822, 986-1041)
  inline float WabWendland(float rr2)const{                            // S/FunSphKernel.h:206-213
    const float rad=sqrtf(rr2), qq=rad/KernelH; const float wqq=qq+qq+1.f, wqq1=1.f-0.5f*qq, wqq2=wqq1*wqq1;
    return Awen*wqq*wqq2*wqq2;
  }
  inline float Wab(float rr2)const{ return TKernel==DSPH_KERNEL_WENDLAND? WabWendland(rr2): WabCubic(rr2); }
  // fluid particle range of one (y,z) row of cells around a position (nsearch::Init(pos) + ParticleRange, S/JCellSearch_inline.h:54-83)
  struct PosRange{ int cxini,cxfin,yini,yfin,zini,zfin; };
  PosRange PosCells(const d3 &ps)const{
    const int cx=int((ps.x-DomPosMin.x)/Scell)-int(CellDomainMin[0]);
    const int cy=int((ps.y-DomPosMin.y)/Scell)-int(CellDomainMin[1]);
    const int cz=int((ps.z-DomPosMin.z)/Scell)-int(CellDomainMin[2]);
    const int ncx=int(Ncx),ncy=int(Ncy),ncz=int(Ncz),sd=ScellDiv;
    PosRange r; r.cxini=std::max(cx-sd,0); r.cxfin=std::min(cx+sd,ncx-1)+1; r.yini=std::max(cy-sd,0); r.yfin=std::min(cy+sd,ncy-1)+1; r.zini=std::max(cz-sd,0); r.zfin=std::min(cz+sd,ncz-1)+1;
    return r;
  }
  template<class F> void ForFluidNeighbours(const d3 &ps,F f)const{
    const PosRange r=PosCells(ps);
    const int cellfluid=int(Nsheet*Ncz+1);
    if(r.cxini<r.cxfin)for(int z=r.zini;z<r.zfin;z++)for(int y=r.yini;y<r.yfin;y++){
      const int v=int(Nsheet)*z+int(Ncx)*y+cellfluid;
      for(unsigned p2=BeginCell[v+r.cxini];p2<BeginCell[v+r.cxfin];p2++){
        const float drx=float(ps.x-Pos[p2].x), dry=float(ps.y-Pos[p2].y), drz=float(ps.z-Pos[p2].z);
        const float rr2=drx*drx+dry*dry+drz*drz;
        if(rr2<=KernelSize2 && rr2>=1e-18f && CodeIsFluid(Code[p2]))f(p2,drx,dry,drz,rr2);
      }
    }
  }
  float GaugeMass(const d3 &ps)const{                                   // JGaugeSwl::CalculeMassCpu :542-567
    double summass=0;
    ForFluidNeighbours(ps,[&](unsigned p2,float,float,float,float rr2){ float wab=Wab(rr2); wab*=MassFluid/Velrhop[p2].w; summass+=wab*MassFluid; });
    return float(summass);
  }
  // the pressure of the gauges: fsph::ComputePress with the rounding of the reference build (see PreInteraction: product form, double pow)
  float GaugePress(float rhop)const{ const float x=rhop*(1.0f/RhopZero); return float(double(CteB)*(pow(double(x),double(Gamma))-1.0)); }
  void GaugesConfig(unsigned n,const dsph_gauge *g){
    Gauges.assign(g,g+n); GaugeNext.assign(n,0.0); GaugeRes.assign(n,dsph_gaugeres());
    for(auto &q:Gauges){
      if(q.type<DSPH_GAUGE_VEL || q.type>DSPH_GAUGE_FORCE)throw std::runtime_error("Gauge type is invalid.");
      if(q.type==DSPH_GAUGE_SWL && !(q.masslimit>0))q.masslimit=MassFluid*(Simulate2D? 0.4f: 0.5f);   // JGaugeSwl defaults, S/JDsGaugeSystem.cpp:240
    }
  }
  unsigned GaugesRun(double timestep,bool all){
    unsigned ndue=0;
    for(auto &r:GaugeRes)r.computed=0;
    for(size_t i=0;i<Gauges.size();i++){
      const dsph_gauge &q=Gauges[i];
      if(!(all || (timestep>=GaugeNext[i] && q.computestart<=timestep && timestep<=q.computeend)))continue;   // JGaugeItem::Update S/JDsGaugeItem.h:164
      ndue++;
      float res[3]={0,0,0};
      if(q.type==DSPH_GAUGE_VEL){
        const d3 pt={q.point0[0],q.point0[1],q.point0[2]};
        const bool ptout=(pt.x!=pt.x||pt.y!=pt.y||pt.z!=pt.z||pt.x<DomPosMin.x||pt.y<DomPosMin.y||pt.z<DomPosMin.z||pt.x>=DomPosMax.x||pt.y>=DomPosMax.y||pt.z>=DomPosMax.z);
        double sv[3]={0,0,0};
        if(!ptout)ForFluidNeighbours(pt,[&](unsigned p2,float,float,float,float rr2){
          float wab=Wab(rr2); const f4 v2=Velrhop[p2]; wab*=MassFluid/v2.w;
          sv[0]+=wab*v2.x; sv[1]+=wab*v2.y; sv[2]+=wab*v2.z;
        });
        res[0]=float(sv[0]); res[1]=float(sv[1]); res[2]=float(sv[2]);
      }
      else if(q.type==DSPH_GAUGE_SWL){
        // JGaugeSwl::SetPoints :447-465
        const double d[3]={q.point2[0]-q.point0[0],q.point2[1]-q.point0[1],q.point2[2]-q.point0[2]};
        const double dis=sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]);
        unsigned pointnp=0; d3 dir={0,0,0};
        if(dis>0 && q.pointdp>0){
          pointnp=unsigned(dis/q.pointdp);
          if(dis-(q.pointdp*pointnp)>=q.pointdp*0.1)pointnp++;
          if(pointnp<1)pointnp++;
          const double dp=dis/pointnp;
          dir={d[0]/dis*dp,d[1]/dis*dp,d[2]/dis*dp};
        }
        // first change from fluid to empty along the line (:575-600)
        const d3 p0={q.point0[0],q.point0[1],q.point0[2]};
        d3 ptsurf={DBL_MAX,DBL_MAX,DBL_MAX}; float mpre=0; d3 ptpos=p0;
        for(unsigned cp=0;cp<=pointnp;cp++){
          const float mass=GaugeMass(ptpos);
          if(mass>q.masslimit)mpre=mass;
          if(mass<q.masslimit && mpre){
            const float fxm1=(q.masslimit-mpre)/(mass-mpre)-1;
            ptsurf={ptpos.x+dir.x*double(fxm1),ptpos.y+dir.y*double(fxm1),ptpos.z+dir.z*double(fxm1)};
            cp=pointnp+1;
          }
          ptpos={ptpos.x+dir.x,ptpos.y+dir.y,ptpos.z+dir.z};
        }
        if(ptsurf.x==DBL_MAX){ const double kk=(mpre? double(pointnp): 0.0); ptsurf={p0.x+dir.x*kk,p0.y+dir.y*kk,p0.z+dir.z*kk}; }
        res[0]=float(ptsurf.x); res[1]=float(ptsurf.y); res[2]=float(ptsurf.z);
      }
      else if(q.type==DSPH_GAUGE_MAXZ){
        // GetInteractionCellsMaxZ :760-775 + CalculeCpu :780-822
        const float maxdist2=q.distlimit*q.distlimit;
        const int cx=int((q.point0[0]-DomPosMin.x)/Scell)-int(CellDomainMin[0]);
        const int cy=int((q.point0[1]-DomPosMin.y)/Scell)-int(CellDomainMin[1]);
        const int cz=int((q.point0[2]-DomPosMin.z)/Scell)-int(CellDomainMin[2]);
        const int cz2=int((q.point0[2]+q.height-DomPosMin.z)/Scell)-int(CellDomainMin[2]);
        const int ncel=int(ceil(q.distlimit/Scell));
        const int cxini=std::max(cx-ncel,0), cxfin=std::min(cx+ncel+1,int(Ncx)), yini=std::max(cy-ncel,0), yfin=std::min(cy+ncel+1,int(Ncy)), zini=std::max(cz-ncel,0), zfin=std::min(cz2+ncel+1,int(Ncz));
        const int cellfluid=int(Nsheet*Ncz+1);
        unsigned pmax=UINT_MAX; float zmax=-FLT_MAX;
        if(cxini<cxfin)for(int z=zfin-1;z>=zini && pmax==UINT_MAX;z--)for(int y=yini;y<yfin;y++){
          const int ymod=int(Nsheet)*z+cellfluid+int(Ncx)*y;
          for(unsigned p2=BeginCell[cxini+ymod];p2<BeginCell[cxfin+ymod];p2++)if(Pos[p2].z>zmax){
            const float drx=float(q.point0[0]-Pos[p2].x), dry=float(q.point0[1]-Pos[p2].y);
            if(drx*drx+dry*dry<=maxdist2 && CodeIsFluid(Code[p2])){ zmax=float(Pos[p2].z); pmax=p2; }
          }
        }
        res[2]=std::max(zmax,float(DomPosMin.z));
      }
      else{
        // force on the boundary particles of one mk block from the pressure term of the momentum equation, fluid neighbours only (:986-1041)
        std::vector<f3> partace(q.count,f3{0,0,0});
        for(unsigned p1=0;p1<NpbOk;p1++)if((Code[p1]&(DSPH_CODE_MASKTYPE|DSPH_CODE_MASKVALUE))==q.code && CodeIsNormal(Code[p1])){
          const f4 v1=Velrhop[p1]; const float press1=GaugePress(v1.w);
          f3 ace={0,0,0};
          ForFluidNeighbours(Pos[p1],[&](unsigned p2,float drx,float dry,float drz,float rr2){
            const float fac=Fac(rr2);
            const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
            const f4 v2=Velrhop[p2]; const float press2=GaugePress(v2.w);
            const float prs=(press1+press2)/(v1.w*v2.w)+(TKernel==DSPH_KERNEL_CUBIC? TensilCubic(rr2,v1.w,press1,v2.w,press2): 0.f);
            const float pv=-prs*MassFluid;
            ace.x+=pv*frx; ace.y+=pv*fry; ace.z+=pv*frz;
          });
          const unsigned kk=Idp[p1]-q.idbegin; if(kk<q.count)partace[kk]=ace;
        }
        f3 sum={0,0,0}; for(auto &a:partace){ sum.x+=a.x; sum.y+=a.y; sum.z+=a.z; }
        res[0]=sum.x*MassBound; res[1]=sum.y*MassBound; res[2]=sum.z*MassBound;
      }
      dsph_gaugeres &r=GaugeRes[i]; r.timestep=timestep; r.value[0]=res[0]; r.value[1]=res[1]; r.value[2]=res[2]; r.computed=1;
      const double cdt=q.computedt;                                          // JGaugeItem::SetTimeStep :204-211
      if(cdt){ const unsigned nt=unsigned(timestep/cdt); GaugeNext[i]=cdt*nt; if(GaugeNext[i]<=timestep)GaugeNext[i]=cdt*(nt+1); }
    }
    return ndue;
  }

  // ---- fsph::GetKernel_WabFac S/FunSphKernel.h:97-118 (cubic), :226-234 (Wendland) in the associations of the reference's -O3 -ffast-math
  //      binary (read from the disassembly of oracle/_ref/obj52/JSphCpu.o): fac=((bwen*q)*(w^2*w))/r, wab=(awen*(1+2q))*(w^2*w^2); cubic far
  //      branch wab=(a24*w)*w^2.  The corrected density is compared at the last bit of a float near 1000, and the 4x4 system amplifies a
  //      last-bit difference of a kernel value into ~10 ulp of it.
  inline float WabFac(float rr2,float &fac)const{
    const float rad=sqrtf(rr2), qq=rad/KernelH;
    if(TKernel==DSPH_KERNEL_WENDLAND){
      const float wqq1=1.f-0.5f*qq, wqq2=wqq1*wqq1;
      fac=((Bwen*qq)*(wqq1*wqq2))/rad;
      const float wqq=(qq+qq)+1.f;
      return (Awen*wqq)*(wqq2*wqq2);
    }
    if(rad>KernelH){ const float wqq1=2.0f-qq, wqq2=wqq1*wqq1; fac=(CubC2*wqq2)/rad; return (CubA24*wqq1)*wqq2; }
    const float wqq2=qq*qq;
    fac=(CubC1*qq+CubD1*wqq2)/rad;
    return CubA2*(1.0f+(0.75f*qq-1.5f)*wqq2);
  }
  // ---- JSphCpu::InteractionMdbcCorrectionT2 S/JSphCpu.cpp:1020-1190, slip mode velocity=0 (the only one the reference accepts,
  //      S/JSph.cpp:788): the fluid density extrapolated to the ghost node of every boundary particle with a first-order consistent
  //      correction (moment matrix 4x4, 3x3 in 2-D, inverted by cofactors like fmath::InverseMatrix4x4 / 3x3), mirrored back.
  void MdbcCorrection(){
    const float determlimit=1e-3f;                                                       // S/JSphCpu.cpp:1197
    const int n=(Simulate2D? 3: 4);
    for(unsigned p1=0;p1<NpbOk;p1++){
      const f3 bn=BoundNormal[p1];
      if(bn.x==0 && bn.y==0 && bn.z==0)continue;
      float rhopfinal=FLT_MAX, sumwab=0;
      const d3 g={Pos[p1].x+double(bn.x),Pos[p1].y+double(bn.y),Pos[p1].z+double(bn.z)};
      float rhopp1=0; f3 grad={0,0,0};
      double a[4][4]; for(int i=0;i<4;i++)for(int j=0;j<4;j++)a[i][j]=0;
      // nsearch::Init for a position (S/JCellSearch_inline.h:54-70)
      const int cx=int((g.x-DomPosMin.x)/Scell)-int(CellDomainMin[0]);
      const int cy=int((g.y-DomPosMin.y)/Scell)-int(CellDomainMin[1]);
      const int cz=int((g.z-DomPosMin.z)/Scell)-int(CellDomainMin[2]);
      const int ncx=int(Ncx),ncy=int(Ncy),ncz=int(Ncz),sd=ScellDiv;
      const int cxini=std::max(cx-sd,0), cxfin=std::min(cx+sd,ncx-1)+1, yini=std::max(cy-sd,0), yfin=std::min(cy+sd,ncy-1)+1, zini=std::max(cz-sd,0), zfin=std::min(cz+sd,ncz-1)+1;
      const int cellfluid=int(Nsheet*Ncz+1);
      if(cxini<cxfin)for(int z=zini;z<zfin;z++)for(int y=yini;y<yfin;y++){
        const int v=int(Nsheet)*z+ncx*y+cellfluid;
        const unsigned pini=BeginCell[v+cxini], pfin=BeginCell[v+cxfin];
        for(unsigned p2=pini;p2<pfin;p2++){
          const float drx=float(g.x-Pos[p2].x), dry=float(g.y-Pos[p2].y), drz=float(g.z-Pos[p2].z);
          const float rr2=(drx*drx+dry*dry+drz*drz);
          if(rr2<=KernelSize2 && CodeIsFluid(Code[p2])){
            float fac; const float wab=WabFac(rr2,fac);
            const float frx=fac*drx, fry=fac*dry, frz=fac*drz;
            const float massp2=MassFluid, volp2=massp2/Velrhop[p2].w;
            rhopp1+=massp2*wab; grad.x+=massp2*frx; grad.y+=massp2*fry; grad.z+=massp2*frz;
            const float vwab=wab*volp2; sumwab+=vwab;
            const float vfrx=frx*volp2, vfry=fry*volp2, vfrz=frz*volp2;
            if(Simulate2D){
              a[0][0]+=vwab; a[0][1]+=drx*vwab; a[0][2]+=drz*vwab;
              a[1][0]+=vfrx; a[1][1]+=drx*vfrx; a[1][2]+=drz*vfrx;
              a[2][0]+=vfrz; a[2][1]+=drx*vfrz; a[2][2]+=drz*vfrz;
            }
            else{
              a[0][0]+=vwab; a[0][1]+=drx*vwab; a[0][2]+=dry*vwab; a[0][3]+=drz*vwab;
              a[1][0]+=vfrx; a[1][1]+=drx*vfrx; a[1][2]+=dry*vfrx; a[1][3]+=drz*vfrx;
              a[2][0]+=vfry; a[2][1]+=drx*vfry; a[2][2]+=dry*vfry; a[2][3]+=drz*vfry;
              a[3][0]+=vfrz; a[3][1]+=drx*vfrz; a[3][2]+=dry*vfrz; a[3][3]+=drz*vfrz;
            }
          }
        }
      }
      if(sumwab>=MdbcThreshold || (MdbcThreshold>=2 && sumwab+2>=MdbcThreshold)){
        // minors and determinant of the n x n block; inverse(r,c) = (-1)^(r+c) * minor(c,r) / det
        auto det3=[](const double m[3][3]){ return m[0][0]*(m[1][1]*m[2][2]-m[1][2]*m[2][1])-m[0][1]*(m[1][0]*m[2][2]-m[1][2]*m[2][0])+m[0][2]*(m[1][0]*m[2][1]-m[1][1]*m[2][0]); };
        auto minor=[&](int ri,int ci){                       // determinant of the block without row ri and column ci
          if(n==3){ int r[2],c[2],k=0; for(int i=0;i<3;i++)if(i!=ri)r[k++]=i; k=0; for(int j=0;j<3;j++)if(j!=ci)c[k++]=j; return a[r[0]][c[0]]*a[r[1]][c[1]]-a[r[0]][c[1]]*a[r[1]][c[0]]; }
          double m[3][3]; int rr=0;
          for(int i=0;i<4;i++)if(i!=ri){ int cc=0; for(int j=0;j<4;j++)if(j!=ci){ m[rr][cc]=a[i][j]; cc++; } rr++; }
          return det3(m);
        };
        double determ=0;
        for(int j=0;j<n;j++)determ+=((j&1)? -a[0][j]: a[0][j])*minor(0,j);
        if(fabs(determ)>=double(determlimit)){
          const double b[4]={double(rhopp1),double(grad.x),Simulate2D? double(grad.z): double(grad.y),double(grad.z)};
          double x[4]={0,0,0,0};
          for(int r=0;r<n;r++){ double sum=0; for(int c=0;c<n;c++)sum+=((((r+c)&1)? -minor(c,r): minor(c,r))/determ)*b[c]; x[r]=sum; }
          const float rhoghost=float(x[0]);
          if(Simulate2D){ const float grx=-float(x[1]), grz=-float(x[2]); rhopfinal=(rhoghost+grx*(-bn.x)+grz*(-bn.z)); }
          else{ const float grx=-float(x[1]), gry=-float(x[2]), grz=-float(x[3]); rhopfinal=(rhoghost+grx*(-bn.x)+gry*(-bn.y)+grz*(-bn.z)); }
        }
        else if(a[0][0]>0)rhopfinal=float(double(rhopp1)/a[0][0]);
        rhopfinal=(rhopfinal!=FLT_MAX? rhopfinal: RhopZero);
        Velrhop[p1].w=rhopfinal;
      }
    }
  }
  void InteractionForces(int interstep=-1){
    // boundary correction for mDBC before every force evaluation, the corrector one only with MDBCCorrector (S/JSphCpuSingle.cpp:525)
    if(Mdbc && (MdbcCorrector || interstep!=DSPH_INTERSTEP_SYMCORRECTOR)){
      if(BoundNormal.size()<Npb)throw std::runtime_error("mDBC needs the normal vectors of the boundary particles (orc_particles_set_boundnormals).");
      MdbcCorrection();
    }
    PreInteraction();
    float viscdt=0;
    if(MultiPhase){ float viscetadt=0; InteractionForcesNN(viscdt,viscetadt); ViscEtaDtMax=viscetadt; }   // M/JSphCpuSingle.cpp:594-636
    else if(Np>Npb){
      InteractionForcesFluid(false,Visco,viscdt);
      InteractionForcesFluid(true,Visco*ViscoBoundFactor,viscdt);
      if(TVisco==DSPH_VISCO_LAMINARSPS)ComputeSpsTau();
    }
    if(NpbOk)InteractionForcesBound(viscdt);
    if(Simulate2D)for(unsigned p=Npb;p<Np;p++)Ace[p].y=0;
    if(!Delta.empty())for(unsigned p=Npb;p<Np;p++)if(Delta[p]!=FLT_MAX)Ar[p]+=Delta[p];
    ViscDtMax=viscdt;
    // ComputeAceMax S/JSphCpuSingle.cpp:583-636 (periodic particles ignored when PeriActive)
    float amax=0;
    for(unsigned p=Npb;p<Np;p++){
      const bool use=(!Ghosts() || CodeIsNormal(Code[p]));
      const f3 a=(use? Ace[p]: f3{0,0,0});
      const float a2=a.x*a.x+a.y*a.y+a.z*a.z; if(amax<a2)amax=a2;
    }
    AceMax=sqrt(double(amax)); LastAce2=amax;
  }
  // ---- JSphCpu::DtVariable S/JSphCpu.cpp:1614-1639
  double DtVariable(double *pdt1,double *pdt2,double *pdt3=nullptr)const{
    const double dt1=(AceMax? (sqrt(double(KernelH)/AceMax)): DBL_MAX);
    const double dt2=double(KernelH)/(std::max(Cs0,VelMax*10.)+double(KernelH)*ViscDtMax);
    // viscous time step of the multiphase fork: M/JSphCpu.cpp:1687-1689
    const double dt3=(MultiPhase && ViscEtaDtMax*Lamda>0? double(KernelH*KernelH)/double(ViscEtaDtMax*Lamda): DBL_MAX);
    double dt=CFLnumber*std::min(dt3,std::min(dt1,dt2));
    if(pdt3)*pdt3=dt3;
    if(cfg.dtfixed_on)dt=cfg.dtfixed;
    if(dt!=dt || std::isinf(dt))throw std::runtime_error("The computed Dt is NaN or infinity.");
    if(dt<DtMin)dt=DtMin;
    if(pdt1)*pdt1=dt1; if(pdt2)*pdt2=dt2;
    return dt;
  }
  // ---- JSphShifting::RunCpu S/JSphShifting.cpp:387-417
  void RunShifting(double dt){
    const double coefumagn=dt*ShiftCoef*KernelH;
    const double coeftfs=(Simulate2D? 2.0: 3.0)-ShiftTFS;
    const float maxdist=float(Dp*0.1);
    for(unsigned p=Npb;p<Np;p++){
      const f4 rs=ShiftPosfs[p];
      if(rs.x!=FLT_MAX){
        const double vx=double(Velrhop[p].x),vy=double(Velrhop[p].y),vz=double(Velrhop[p].z);
        double umagn=coefumagn*sqrt(vx*vx+vy*vy+vz*vz);
        if(ShiftTFS){
          if(rs.w<ShiftTFS)umagn=0;
          else umagn*=(double(rs.w)-ShiftTFS)/coeftfs;
        }
        const float sx=float(double(rs.x)*umagn), sy=float(double(rs.y)*umagn), sz=float(double(rs.z)*umagn);
        if(MultiPhase){   // the v5.0 fork clamps one side only: M/JSphShifting.cpp:412-414
          ShiftPosfs[p].x=(sx<maxdist? sx: maxdist); ShiftPosfs[p].y=(sy<maxdist? sy: maxdist); ShiftPosfs[p].z=(sz<maxdist? sz: maxdist);
          continue;
        }
        ShiftPosfs[p].x=(fabsf(sx)<maxdist? sx: (sx>=0? maxdist: -maxdist));
        ShiftPosfs[p].y=(fabsf(sy)<maxdist? sy: (sy>=0? maxdist: -maxdist));
        ShiftPosfs[p].z=(fabsf(sz)<maxdist? sz: (sz>=0? maxdist: -maxdist));
      }
      else ShiftPosfs[p]=f4{0,0,0,0};
    }
  }
  // ---- JSphCpu::UpdatePos S/JSphCpu.cpp:1240-1293
  void UpdatePos(d3 rpos,double movx,double movy,double movz,bool outrhop,unsigned p){
    const bool outmove=(fabsf(float(movx))>MovLimit || fabsf(float(movy))>MovLimit || fabsf(float(movz))>MovLimit);
    rpos.x+=movx; rpos.y+=movy; rpos.z+=movz;
    double dx=rpos.x-MapRealPosMin.x, dy=rpos.y-MapRealPosMin.y, dz=rpos.z-MapRealPosMin.z;
    bool out=(dx!=dx||dy!=dy||dz!=dz||dx<0||dy<0||dz<0||dx>=MapRealSize.x||dy>=MapRealSize.y||dz>=MapRealSize.z);
    if(PeriActive && out){
      if(PeriX){ if(dx<0){ dx-=PeriXinc.x; dy-=PeriXinc.y; dz-=PeriXinc.z; } if(dx>=MapRealSize.x){ dx+=PeriXinc.x; dy+=PeriXinc.y; dz+=PeriXinc.z; } }
      if(PeriY){ if(dy<0){ dx-=PeriYinc.x; dy-=PeriYinc.y; dz-=PeriYinc.z; } if(dy>=MapRealSize.y){ dx+=PeriYinc.x; dy+=PeriYinc.y; dz+=PeriYinc.z; } }
      if(PeriZ){ if(dz<0){ dx-=PeriZinc.x; dy-=PeriZinc.y; dz-=PeriZinc.z; } if(dz>=MapRealSize.z){ dx+=PeriZinc.x; dy+=PeriZinc.y; dz+=PeriZinc.z; } }
      const bool outx=!PeriX && (dx<0||dx>=MapRealSize.x), outy=!PeriY && (dy<0||dy>=MapRealSize.y), outz=!PeriZ && (dz<0||dz>=MapRealSize.z);
      out=(outx||outy||outz);
      rpos={dx+MapRealPosMin.x,dy+MapRealPosMin.y,dz+MapRealPosMin.z};
    }
    Pos[p]=rpos;
    if(outrhop||outmove||out){
      typecode rcode=Code[p];
      if(out)rcode=typecode(CodeSetNormal(rcode)|DSPH_CODE_OUTPOS);
      else if(outrhop)rcode=typecode(CodeSetNormal(rcode)|DSPH_CODE_OUTRHOP);
      else rcode=typecode(CodeSetNormal(rcode)|DSPH_CODE_OUTMOVE);
      Code[p]=rcode; Dcell[p]=0xFFFFFFFFu;
    }
    else{
      if(PeriActive){ dx=rpos.x-DomPosMin.x; dy=rpos.y-DomPosMin.y; dz=rpos.z-DomPosMin.z; }
      const unsigned cx=unsigned(dx/Scell),cy=unsigned(dy/Scell),cz=unsigned(dz/Scell);
      Dcell[p]=DcCell(DomCellCode,cx,cy,cz);
    }
  }
  // ---- JSphCpu::ComputeVerlet + ComputeVerletVarsFluid + ComputeVelrhopBound S/JSphCpu.cpp:1300-1399
  void ComputeVerlet(double dt){
    const bool shift=(ShiftMode!=0);
    VerletStep++;
    const bool euler=!(VerletStep<VerletSteps);
    const double dt2=(euler? dt: dt+dt);
    const std::vector<f4> &vr2=(euler? Velrhop: VelrhopM1);
    std::vector<f4> vnew(Np);
    const double dt205=0.5*dt*dt;
    const d3 gravity={double(Gravity.x),double(Gravity.y),double(Gravity.z)};
    for(unsigned p=Npb;p<Np;p++){
      if(Mgpu && CodeIsPeriodic(Code[p])){ vnew[p]=Velrhop[p]; continue; }
      const float rhopnew=float(double(vr2[p].w)+dt2*Ar[p]);
      const d3 acegr={double(Ace[p].x)+gravity.x,double(Ace[p].y)+gravity.y,double(Ace[p].z)+gravity.z};
      double dx=double(Velrhop[p].x)*dt+acegr.x*dt205, dy=double(Velrhop[p].y)*dt+acegr.y*dt205, dz=double(Velrhop[p].z)*dt+acegr.z*dt205;
      if(shift){ dx+=double(ShiftPosfs[p].x); dy+=double(ShiftPosfs[p].y); dz+=double(ShiftPosfs[p].z); }
      const bool outrhop=(rhopnew<RhopOutMin || rhopnew>RhopOutMax);
      const f4 rv={float(double(vr2[p].x)+acegr.x*dt2),float(double(vr2[p].y)+acegr.y*dt2),float(double(vr2[p].z)+acegr.z*dt2),rhopnew};
      UpdatePos(Pos[p],dx,dy,dz,outrhop,p);
      vnew[p]=rv;
    }
    for(unsigned p=0;p<Npb;p++){
      const float rhopnew=float(double(vr2[p].w)+dt2*Ar[p]);
      vnew[p]=f4{0,0,0,(rhopnew<RhopZero? RhopZero: rhopnew)};
    }
    // new values computed in VelrhopM1 then swapped (S/JSphCpu.cpp:1385-1398)
    std::vector<f4> old=Velrhop;
    for(unsigned p=0;p<Np;p++)Velrhop[p]=vnew[p];
    for(unsigned p=0;p<Np;p++)VelrhopM1[p]=old[p];
    if(euler)VerletStep=0;
  }
  // ---- JSphCpu::ComputeSymplecticPre S/JSphCpu.cpp:1406-1497
  void ComputeSymplecticPre(double dt){
    const double dt05=dt*.5;
    for(unsigned p=0;p<Np;p++){ PosPre[p]=Pos[p]; VelrhopPre[p]=Velrhop[p]; }
    HasPre=true;
    for(unsigned p=0;p<Npb;p++){
      const f4 vr=VelrhopPre[p];
      const float rhopnew=float(double(vr.w)+dt05*Ar[p]);
      Velrhop[p]=f4{vr.x,vr.y,vr.z,(rhopnew<RhopZero? RhopZero: rhopnew)};
    }
    std::vector<d3> mov(Np);
    for(unsigned p=Npb;p<Np;p++){
      const typecode rcode=Code[p];
      const float rhopnew=float(double(VelrhopPre[p].w)+dt05*Ar[p]);
      const double dx=double(VelrhopPre[p].x)*dt05, dy=double(VelrhopPre[p].y)*dt05, dz=double(VelrhopPre[p].z)*dt05;   // predictor never shifts (S/JSphCpu.cpp:1408)
      const bool outrhop=(rhopnew<RhopOutMin || rhopnew>RhopOutMax);
      const f4 rv={float(double(VelrhopPre[p].x)+(double(Ace[p].x)+Gravity.x)*dt05),
                   float(double(VelrhopPre[p].y)+(double(Ace[p].y)+Gravity.y)*dt05),
                   float(double(VelrhopPre[p].z)+(double(Ace[p].z)+Gravity.z)*dt05),rhopnew};
      mov[p]={dx,dy,dz}; Velrhop[p]=rv;
      if(outrhop && CodeIsNormal(rcode))Code[p]=typecode(CodeSetNormal(rcode)|DSPH_CODE_OUTRHOP);
    }
    for(unsigned p=Npb;p<Np;p++){
      const typecode rcode=Code[p];
      const bool outrhop=CodeIsOutRhop(rcode);
      const bool normal=(!Ghosts() || outrhop || CodeIsNormal(rcode));
      if(normal){
        if(CodeIsFluid(rcode))UpdatePos(PosPre[p],mov[p].x,mov[p].y,mov[p].z,outrhop,p);
        else Pos[p]=PosPre[p];
      }
    }
    for(unsigned p=0;p<Npb;p++)Pos[p]=PosPre[p];
  }
  // ---- JSphCpu::ComputeSymplecticCorr S/JSphCpu.cpp:1503-1606
  void ComputeSymplecticCorr(double dt){
    const bool shift=(ShiftMode!=0);
    const double dt05=dt*.5;
    for(unsigned p=0;p<Npb;p++){
      const double epsilon_rdot=(-double(Ar[p])/double(Velrhop[p].w))*dt;
      const float rhopnew=float(double(VelrhopPre[p].w)*(2.-epsilon_rdot)/(2.+epsilon_rdot));
      Velrhop[p]=f4{0,0,0,(rhopnew<RhopZero? RhopZero: rhopnew)};
    }
    std::vector<d3> mov(Np);
    for(unsigned p=Npb;p<Np;p++){
      const typecode rcode=Code[p];
      const double epsilon_rdot=(-double(Ar[p])/double(Velrhop[p].w))*dt;
      const float rhopnew=float(double(VelrhopPre[p].w)*(2.-epsilon_rdot)/(2.+epsilon_rdot));
      const f4 rv={float(double(VelrhopPre[p].x)+(double(Ace[p].x)+Gravity.x)*dt),
                   float(double(VelrhopPre[p].y)+(double(Ace[p].y)+Gravity.y)*dt),
                   float(double(VelrhopPre[p].z)+(double(Ace[p].z)+Gravity.z)*dt),rhopnew};
      double dx=(double(VelrhopPre[p].x)+double(rv.x))*dt05, dy=(double(VelrhopPre[p].y)+double(rv.y))*dt05, dz=(double(VelrhopPre[p].z)+double(rv.z))*dt05;
      if(shift){ dx+=double(ShiftPosfs[p].x); dy+=double(ShiftPosfs[p].y); dz+=double(ShiftPosfs[p].z); }
      const bool outrhop=(rhopnew<RhopOutMin || rhopnew>RhopOutMax);
      mov[p]={dx,dy,dz};
      if(outrhop && CodeIsNormal(rcode))Code[p]=typecode(CodeSetNormal(rcode)|DSPH_CODE_OUTRHOP);
      Velrhop[p]=rv;
    }
    for(unsigned p=Npb;p<Np;p++){
      const typecode rcode=Code[p];
      const bool outrhop=CodeIsOutRhop(rcode);
      const bool normal=(!Ghosts() || outrhop || CodeIsNormal(rcode));
      if(normal){
        if(CodeIsFluid(rcode))UpdatePos(PosPre[p],mov[p].x,mov[p].y,mov[p].z,outrhop,p);
        else Pos[p]=PosPre[p];
      }
    }
    HasPre=false;
  }

  // ---- x-slab decomposition: same classification / record layout / append order as csrc/mgpu.cuh + Engine::Mg*
  struct MgRecord{ double pos[3]; f4 velrhop; double pospre[3]; f4 aux; unsigned idp,code; };
  void MgClassify(unsigned p,int &tl,int &tr)const{
    tl=tr=-1;
    const typecode rc=Code[p];
    if(!CodeIsNormal(rc) || Dcell[p]==0xFFFFFFFFu)return;
    const unsigned cx=DcCellx(DomCellCode,Dcell[p]); const bool fluid=CodeIsFluid(rc); const unsigned sd=MgHalo();
    // kinds: 0 boundary ghost, 1 boundary migrant (only after the cuts moved: re-balancing), 2 fluid migrant, 3 fluid ghost
    if(MgRank>0){ if(cx<MgXlo)tl=(fluid? 2: 1); else if(cx<MgXlo+sd)tl=(fluid? 3: 0); }
    if(MgRank<MgNranks-1){ if(cx>=MgXhi)tr=(fluid? 2: 1); else if(cx+sd>=MgXhi)tr=(fluid? 3: 0); }
  }
  void MgCount(unsigned counts[8]){
    for(unsigned p=0;p<Np;p++){ const typecode rc=Code[p]; if(CodeIsPeriodic(rc))Code[p]=typecode(CodeSetNormal(rc)|DSPH_CODE_OUTIGNORE); }
    MgFlags.assign(Np,0);
    for(int g=0;g<8;g++)MgCounts[g]=0;
    for(unsigned p=0;p<Np;p++){
      int tl,tr; MgClassify(p,tl,tr); unsigned char f=0;
      if(tl>=0){ f|=(unsigned char)(1u<<tl); MgCounts[tl]++; }
      if(tr>=0){ f|=(unsigned char)(1u<<(4+tr)); MgCounts[4+tr]++; }
      MgFlags[p]=f;
    }
    for(int g=0;g<8;g++)counts[g]=MgCounts[g];
  }
  // owned (normal) particles per cell column, fluid and boundary separately: input of the slab re-balancing
  void MgColumnCounts(unsigned *fluid,unsigned *bound,unsigned ncol)const{
    for(unsigned c=0;c<ncol;c++)fluid[c]=bound[c]=0;
    for(unsigned p=0;p<Np;p++){
      const typecode rc=Code[p];
      if(!CodeIsNormal(rc) || Dcell[p]==0xFFFFFFFFu)continue;
      const unsigned cx=DcCellx(DomCellCode,Dcell[p]);
      if(cx<ncol){ if(CodeIsFluid(rc))fluid[cx]++; else bound[cx]++; }
    }
  }
  // halo width in cells: 2h, or 4h for the SPH velocity-gradient approach of the multiphase fork, whose second pair pass reads
  // eta_eff / tau of the neighbours: with a double halo the first pass is complete for the inner ghost layer (no second exchange)
  unsigned MgHalo()const{ return unsigned(ScellDiv)*(MultiPhase && TVelGrad==DSPH_VELGRAD_SPH && TVisco!=DSPH_VISCO_ARTIFICIAL? 2u: 1u); }
  bool MgSps()const{ return TVisco==DSPH_VISCO_LAMINARSPS && !MultiPhase; }     // SpsTau is particle state: it travels with the record
  unsigned MgRecordBytes()const{ return unsigned(sizeof(MgRecord))+(MgSps()? 24u: 0u); }
  void MgPack(void *sendleft,void *sendright){
    char *out[2]={(char*)sendleft,(char*)sendright};
    const unsigned sd=MgHalo(), rb=MgRecordBytes();
    for(int d=0;d<2;d++){
      unsigned base=0;
      for(int kd=0;kd<4;kd++){
        const int g=4*d+kd; unsigned n=0; const bool migrant=(kd==1 || kd==2);
        for(unsigned p=0;p<Np && MgCounts[g];p++)if((MgFlags[p]>>g)&1u){
          MgRecord r; memset(&r,0,sizeof(r));
          r.pos[0]=Pos[p].x; r.pos[1]=Pos[p].y; r.pos[2]=Pos[p].z; r.velrhop=Velrhop[p];
          if(TStep==DSPH_STEP_SYMPLECTIC && HasPre){ r.pospre[0]=PosPre[p].x; r.pospre[1]=PosPre[p].y; r.pospre[2]=PosPre[p].z; r.aux=VelrhopPre[p]; }
          if(TStep==DSPH_STEP_VERLET)r.aux=VelrhopM1[p];
          r.idp=Idp[p];
          const typecode rc=Code[p];
          r.code=(migrant? rc: typecode(CodeSetNormal(rc)|DSPH_CODE_PERIODIC));
          char *dstp=out[d]+size_t(base+n)*rb;
          memcpy(dstp,&r,sizeof(r));
          if(MgSps())memcpy(dstp+sizeof(r),&SpsTau[p],24);
          n++;
          if(migrant){
            const unsigned cx=DcCellx(DomCellCode,Dcell[p]);
            const bool band=(d==0? (cx+sd>=MgXlo): (cx<MgXhi+sd));
            Code[p]=typecode(CodeSetNormal(rc)|(band? DSPH_CODE_PERIODIC: DSPH_CODE_OUTIGNORE));
          }
        }
        base+=MgCounts[g];
      }
    }
  }
  void MgUnpack(const unsigned counts[8],const void *recvleft,const void *recvright){
    const unsigned bgl=counts[0]+counts[1], fl=counts[2]+counts[3], bgr=counts[4]+counts[5], fr=counts[6]+counts[7];   // boundary | fluid per side
    const unsigned total=bgl+fl+bgr+fr;
    if(Idp.size()<size_t(Np)+total)Resize(Np+total+(Np+total)/20+128);
    unsigned dst=Np;
    const unsigned rb=MgRecordBytes();
    auto up=[&](unsigned n,const char *src){
      for(unsigned i=0;i<n;i++){
        MgRecord r; memcpy(&r,src+size_t(i)*rb,sizeof(r)); const unsigned p=dst+i;
        if(MgSps())memcpy(&SpsTau[p],src+size_t(i)*rb+sizeof(r),24);
        Pos[p]={r.pos[0],r.pos[1],r.pos[2]}; Velrhop[p]=r.velrhop;
        if(TStep==DSPH_STEP_SYMPLECTIC && HasPre){ PosPre[p]={r.pospre[0],r.pospre[1],r.pospre[2]}; VelrhopPre[p]=r.aux; }
        if(TStep==DSPH_STEP_VERLET)VelrhopM1[p]=r.aux;
        Idp[p]=r.idp; Code[p]=typecode(r.code);
        const unsigned cx=unsigned((r.pos[0]-DomPosMin.x)/Scell), cy=unsigned((r.pos[1]-DomPosMin.y)/Scell), cz=unsigned((r.pos[2]-DomPosMin.z)/Scell);
        Dcell[p]=DcCell(DomCellCode,cx,cy,cz);
      }
      dst+=n;
    };
    up(bgl,(const char*)recvleft); up(bgr,(const char*)recvright);
    up(fl,(const char*)recvleft+size_t(bgl)*rb); up(fr,(const char*)recvright+size_t(bgr)*rb);
    NpbPer=bgl+bgr; NpfPer=fl+fr; Np+=total; BoundChanged=true;
  }
};

thread_local std::string g_err;
template<class F> int Guard(F f){
  try{ f(); return DSPH_OK; }
  catch(const std::exception &e){ g_err=e.what(); return DSPH_ERR_STATE; }
}

} // namespace

extern "C" {

struct orc_state { Oracle o; };

const char* orc_last_error(void){ return g_err.c_str(); }

int orc_create(orc_state **out,const dsph_config *cfg){
  return Guard([&]{ orc_state *s=new orc_state(); try{ s->o.Configure(*cfg); }catch(...){ delete s; throw; } *out=s; });
}
int orc_destroy(orc_state *s){ delete s; return DSPH_OK; }

int orc_particles_upload(orc_state *s,uint32_t np,uint32_t npb,const uint32_t *idp,const uint16_t *code,const double *pos,const float *velrhop){
  return Guard([&]{
    Oracle &o=s->o;
    o.Np=np; o.Npb=npb; o.NpbOk=npb; o.NpbPer=o.NpfPer=0;
    o.Resize(np+np/10+1024);
    for(unsigned p=0;p<np;p++){
      o.Idp[p]=idp[p];
      o.Code[p]=(code? code[p]: typecode(idp[p]<o.cfg.casenfixed? DSPH_CODE_TYPE_FIXED: DSPH_CODE_TYPE_FLUID));
      o.Pos[p]={pos[3*p],pos[3*p+1],pos[3*p+2]};
      o.Velrhop[p]={velrhop[4*p],velrhop[4*p+1],velrhop[4*p+2],velrhop[4*p+3]};
    }
    if(o.TStep==DSPH_STEP_VERLET)for(unsigned p=0;p<np;p++)o.VelrhopM1[p]=o.Velrhop[p];     // InitRunCpu S/JSphCpu.cpp:422
    if(o.TVisco==DSPH_VISCO_LAMINARSPS)for(unsigned p=0;p<np;p++)o.SpsTau[p]=sym3{0,0,0,0,0,0};
    o.HasPre=false; o.BoundChanged=true; o.BoundLimitOk=o.BoundDivideOk=false; o.BoundNormal.clear();
    o.SymplecticDtPre=o.DtIni; o.VerletStep=0;
    o.LoadDcellParticles();
  });
}
// vectors particle -> ghost node (xyz) of the first `count` particles of the current order (boundary particles first); zero for the rest
int orc_particles_set_boundnormals(orc_state *s,const float *n,uint32_t count){
  return Guard([&]{
    Oracle &o=s->o;
    if(!o.Mdbc)throw std::runtime_error("Normal vectors are only used with cfg.tboundary = DSPH_BOUNDARY_MDBC.");
    if(count>o.Np)throw std::runtime_error("More normal vectors than particles.");
    o.BoundNormal.assign(o.Np,f3{0,0,0});
    for(uint32_t i=0;i<count;i++)o.BoundNormal[i]=f3{n[3*size_t(i)],n[3*size_t(i)+1],n[3*size_t(i)+2]};
  });
}
int orc_particles_get_boundnormals(orc_state *s,float *n,uint32_t count){
  return Guard([&]{
    Oracle &o=s->o;
    if(count>o.BoundNormal.size())throw std::runtime_error("More normal vectors requested than stored.");
    for(uint32_t i=0;i<count;i++){ n[3*size_t(i)]=o.BoundNormal[i].x; n[3*size_t(i)+1]=o.BoundNormal[i].y; n[3*size_t(i)+2]=o.BoundNormal[i].z; }
  });
}
int orc_gauges_config(orc_state *s,uint32_t n,const dsph_gauge *g){ return Guard([&]{ s->o.GaugesConfig(n,g); }); }
int orc_gauges_run(orc_state *s,double timestep,int force_all,dsph_gaugeres *res){
  return Guard([&]{ s->o.GaugesRun(timestep,force_all!=0); if(res)for(size_t i=0;i<s->o.GaugeRes.size();i++)res[i]=s->o.GaugeRes[i]; });
}
int orc_gauges_results(orc_state *s,dsph_gaugeres *res){ return Guard([&]{ for(size_t i=0;i<s->o.GaugeRes.size();i++)res[i]=s->o.GaugeRes[i]; }); }
int orc_gauges_set_time(orc_state *s,double timestep){ s->o.GaugeTime=timestep; return DSPH_OK; }
int orc_particles_set_velrhopm1(orc_state *s,const float *v){
  return Guard([&]{ Oracle &o=s->o; if(o.TStep!=DSPH_STEP_VERLET)throw std::runtime_error("VelrhopM1 is only used by Verlet."); for(unsigned p=0;p<o.Np;p++)o.VelrhopM1[p]={v[4*p],v[4*p+1],v[4*p+2],v[4*p+3]}; });
}
int orc_particles_set_spstau(orc_state *s,const float *t){
  return Guard([&]{ Oracle &o=s->o; if(o.TVisco!=DSPH_VISCO_LAMINARSPS)throw std::runtime_error("SpsTau is only used by Laminar+SPS."); for(unsigned p=0;p<o.Np;p++)o.SpsTau[p]={t[6*p],t[6*p+1],t[6*p+2],t[6*p+3],t[6*p+4],t[6*p+5]}; });
}
int orc_particles_download(orc_state *s,uint32_t first,uint32_t count,uint32_t *idp,uint16_t *code,uint32_t *dcell,double *pos,float *velrhop){
  return Guard([&]{
    Oracle &o=s->o; if(first+count>o.Np)throw std::runtime_error("range out of particles");
    for(unsigned i=0;i<count;i++){
      const unsigned p=first+i;
      if(idp)idp[i]=o.Idp[p]; if(code)code[i]=o.Code[p]; if(dcell)dcell[i]=o.Dcell[p];
      if(pos){ pos[3*i]=o.Pos[p].x; pos[3*i+1]=o.Pos[p].y; pos[3*i+2]=o.Pos[p].z; }
      if(velrhop){ velrhop[4*i]=o.Velrhop[p].x; velrhop[4*i+1]=o.Velrhop[p].y; velrhop[4*i+2]=o.Velrhop[p].z; velrhop[4*i+3]=o.Velrhop[p].w; }
    }
  });
}
uint32_t orc_np(const orc_state *s){ return s->o.Np; }
uint32_t orc_npb(const orc_state *s){ return s->o.Npb; }

static void FillDivInfo(const Oracle &o,dsph_divinfo *info){
  if(!info)return;
  memset(info,0,sizeof(*info));
  info->np=o.Np; info->npb=o.Npb; info->npbok=o.NpbOk; info->npbper=o.NpbPer; info->npfper=o.NpfPer;
  info->npbout=o.NpbOut; info->npfout=o.NpfOut; info->npbignore=o.NpbIgnore; info->npboutignore=o.NpbOutIgnore; info->npfoutignore=o.NpfOutIgnore;
  info->ncx=o.Ncx; info->ncy=o.Ncy; info->ncz=o.Ncz; info->nct=o.Nct; info->nctt=o.Nctt;
  for(int i=0;i<3;i++){ info->cellmin[i]=o.CellDomainMin[i]; info->cellmax[i]=o.CellDomainMax[i]; info->domcells[i]=o.DomCells[i]; }
  info->domcellcode=o.DomCellCode; info->dividefull=o.DivideFull; info->scell=o.Scell;
  info->domposmin[0]=o.DomPosMin.x; info->domposmin[1]=o.DomPosMin.y; info->domposmin[2]=o.DomPosMin.z;
}
int orc_celldiv_run(orc_state *s,int updateperiodic,dsph_divinfo *info){
  return Guard([&]{ s->o.RunCellDivide(updateperiodic!=0); FillDivInfo(s->o,info); });
}
int orc_celldiv_get_begincell(orc_state *s,uint32_t *begincell,uint32_t capacity){
  return Guard([&]{ const Oracle &o=s->o; if(capacity<o.Nctt)throw std::runtime_error("capacity too small"); memcpy(begincell,o.BeginCell.data(),size_t(o.Nctt)*4); });
}
static void FillForceRes(const Oracle &o,dsph_forceres *res){
  if(!res)return;
  memset(res,0,sizeof(*res));
  res->viscdtmax=o.ViscDtMax; res->acemax=float(o.AceMax); res->velmax=float(o.VelMax);
  res->viscetadtmax=o.ViscEtaDtMax;
  res->dt=o.DtVariable(&res->dt1,&res->dt2,&res->dt3);
}
int orc_forces_run(orc_state *s,int interstep,dsph_forceres *res){
  return Guard([&]{ s->o.InteractionForces(interstep); FillForceRes(s->o,res); });
}
int orc_forces_download(orc_state *s,uint32_t first,uint32_t count,float *ace,float *ar){
  return Guard([&]{
    const Oracle &o=s->o; if(first+count>o.Np || o.Ace.size()<o.Np)throw std::runtime_error("forces are not available");
    for(unsigned i=0;i<count;i++){ const unsigned p=first+i; if(ace){ ace[3*i]=o.Ace[p].x; ace[3*i+1]=o.Ace[p].y; ace[3*i+2]=o.Ace[p].z; } if(ar)ar[i]=o.Ar[p]; }
  });
}
int orc_step_verlet(orc_state *s,double dt){ return Guard([&]{ Oracle &o=s->o; if(o.ShiftMode)o.RunShifting(dt); o.ComputeVerlet(dt); }); }
int orc_step_sympre(orc_state *s,double dt){ return Guard([&]{ Oracle &o=s->o; if(o.ShiftMode)o.RunShifting(dt*.5); o.ComputeSymplecticPre(dt); }); }
int orc_step_symcor(orc_state *s,double dt){ return Guard([&]{ Oracle &o=s->o; if(o.ShiftMode)o.RunShifting(dt); o.ComputeSymplecticCorr(dt); }); }

// JSphCpuSingle::ComputeStep_Sym/_Ver + closing RunCellDivide S/JSphCpuSingle.cpp:674-721,1093-1098
int orc_step(orc_state *s,double *stepdt,dsph_forceres *res){
  return Guard([&]{
    Oracle &o=s->o; double dt;
    if(o.TStep==DSPH_STEP_SYMPLECTIC){
      dt=o.SymplecticDtPre;
      o.InteractionForces(DSPH_INTERSTEP_SYMPREDICTOR);
      const double ddt_p=o.DtVariable(nullptr,nullptr);
      if(o.ShiftMode)o.RunShifting(dt*.5);
      o.ComputeSymplecticPre(dt);
      o.RunCellDivide(true);
      o.InteractionForces(DSPH_INTERSTEP_SYMCORRECTOR);
      const double ddt_c=o.DtVariable(nullptr,nullptr);
      FillForceRes(o,res);
      if(o.ShiftMode)o.RunShifting(dt);
      o.ComputeSymplecticCorr(dt);
      o.SymplecticDtPre=std::min(ddt_p,ddt_c);
    }
    else{
      o.InteractionForces(DSPH_INTERSTEP_VERLET);
      dt=o.DtVariable(nullptr,nullptr);
      FillForceRes(o,res);
      if(o.ShiftMode)o.RunShifting(dt);
      o.ComputeVerlet(dt);
    }
    if(!o.Gauges.empty())o.GaugesRun(o.GaugeTime+dt,false);        // RunGaugeSystem(TimeStep+stepdt), S/JSphCpuSingle.cpp:1095
    o.GaugeTime+=dt;
    o.RunCellDivide(true);
    if(stepdt)*stepdt=dt;
  });
}
double orc_symplectic_dtpre(const orc_state *s){ return s->o.SymplecticDtPre; }
// JSphCpu::RunMotion S/JSphCpu.cpp:1762-1805 for the motion data of one step: CalcRidp (:1660-1687), MoveLinBound (:1692-1703),
// MoveMatBound (:1709-1730, with the mDBC normals)
int orc_motion_run(orc_state *s,const dsph_motiondata *objs,uint32_t nobj,double stepdt){
  return Guard([&]{
    Oracle &o=s->o;
    const unsigned idini=o.cfg.casenfixed, idfin=o.cfg.casenfixed+o.cfg.casenmoving;
    if(idfin==idini)throw std::runtime_error("The case has no moving boundary particles.");
    bool any=false; for(uint32_t i=0;i<nobj;i++)if(objs[i].type!=DSPH_MOTION_NONE)any=true;
    if(!any)return;
    std::vector<unsigned> ridp(idfin-idini,0xffffffffu);
    for(unsigned p=0;p<o.Npb;p++){
      const unsigned id=o.Idp[p];
      if(idini<=id && id<idfin && (!o.PeriActive || CodeSpecial(o.Code[p])==DSPH_CODE_NORMAL))ridp[id-idini]=p;
    }
    o.BoundChanged=true;
    for(uint32_t i=0;i<nobj;i++){
      const dsph_motiondata &m=objs[i];
      const unsigned ini=m.idbegin-idini, fin=ini+m.count;
      if(m.type==DSPH_MOTION_LINEAR){
        for(unsigned id=ini;id<fin;id++){
          const unsigned pid=ridp[id];
          if(pid==0xffffffffu)continue;
          o.UpdatePos(o.Pos[pid],m.linmov[0],m.linmov[1],m.linmov[2],false,pid);
          o.Velrhop[pid].x=float(m.linvel[0]); o.Velrhop[pid].y=float(m.linvel[1]); o.Velrhop[pid].z=float(m.linvel[2]);
        }
      }
      else if(m.type==DSPH_MOTION_MATRIX){
        const double *a=m.matmov;
        for(unsigned id=ini;id<fin;id++){
          const unsigned pid=ridp[id];
          if(pid==0xffffffffu)continue;
          const d3 ps=o.Pos[pid];
          // association of the reference binary (-ffast-math pairs the four terms; pinned by the mv_* dumps)
          d3 ps2; ps2.x=(a[0]*ps.x+a[1]*ps.y)+(a[2]*ps.z+a[3]); ps2.y=(a[4]*ps.x+a[5]*ps.y)+(a[6]*ps.z+a[7]); ps2.z=(a[8]*ps.x+a[9]*ps.y)+(a[10]*ps.z+a[11]);
          if(o.Simulate2D)ps2.y=ps.y;
          const double dx=ps2.x-ps.x, dy=ps2.y-ps.y, dz=ps2.z-ps.z;
          o.UpdatePos(ps,dx,dy,dz,false,pid);
          o.Velrhop[pid].x=float(dx/stepdt); o.Velrhop[pid].y=float(dy/stepdt); o.Velrhop[pid].z=float(dz/stepdt);
          if(o.Mdbc && pid<o.BoundNormal.size()){                          // the image of the ghost node minus the image of the particle
            const f3 bn=o.BoundNormal[pid];
            const d3 gs={ps.x+double(bn.x),ps.y+double(bn.y),ps.z+double(bn.z)};
            d3 gs2; gs2.x=(a[0]*gs.x+a[1]*gs.y)+(a[2]*gs.z+a[3]); gs2.y=(a[4]*gs.x+a[5]*gs.y)+(a[6]*gs.z+a[7]); gs2.z=(a[8]*gs.x+a[9]*gs.y)+(a[10]*gs.z+a[11]);
            o.BoundNormal[pid]=f3{float(gs2.x-ps2.x),float(gs2.y-ps2.y),float(gs2.z-ps2.z)};
          }
        }
      }
    }
  });
}
int orc_set_symplectic_dtpre(orc_state *s,double dt){ s->o.SymplecticDtPre=dt; return DSPH_OK; }
int orc_set_verletstep(orc_state *s,int v){ s->o.VerletStep=v; return DSPH_OK; }
int orc_pips(orc_state *s,uint64_t *fr,uint64_t *fc,uint64_t *br,uint64_t *bc){
  const Oracle &o=s->o; if(fr)*fr=o.PipsFluidReal; if(fc)*fc=o.PipsFluidChecked; if(br)*br=o.PipsBoundReal; if(bc)*bc=o.PipsBoundChecked; return DSPH_OK;
}
int orc_num_threads(void){
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- x-slab decomposition twin (tests/test_mgpu_gloo.py drives two ranks over gloo with this CPU engine) ----
int orc_mgpu_config(orc_state *s,int rank,int nranks,uint32_t xlo,uint32_t xhi){
  return Guard([&]{ Oracle &o=s->o; if(o.PeriActive)throw std::runtime_error("x-slab decomposition cannot be combined with periodic boundaries yet.");
    if(nranks<1||rank<0||rank>=nranks||xlo>=xhi||xhi>o.DomCells[0])throw std::runtime_error("Invalid slab configuration.");
    o.Mgpu=(nranks>1); o.MgRank=rank; o.MgNranks=nranks; o.MgXlo=xlo; o.MgXhi=xhi; });
}
uint32_t orc_mgpu_record_bytes(const orc_state *s){ static_assert(sizeof(Oracle::MgRecord)==88,"wire format"); return s? s->o.MgRecordBytes(): 88u; }
int orc_mgpu_count(orc_state *s,uint32_t counts[8]){ return Guard([&]{ unsigned c[8]; s->o.MgCount(c); for(int i=0;i<8;i++)counts[i]=c[i]; }); }
int orc_mgpu_column_counts(orc_state *s,uint32_t *fluid,uint32_t *bound,uint32_t ncol){ return Guard([&]{ s->o.MgColumnCounts(fluid,bound,ncol); }); }
int orc_mgpu_pack(orc_state *s,void *sl,void *sr){ return Guard([&]{ s->o.MgPack(sl,sr); }); }
int orc_mgpu_unpack(orc_state *s,const uint32_t counts[8],const void *rl,const void *rr){ return Guard([&]{ unsigned c[8]; for(int i=0;i<8;i++)c[i]=counts[i]; s->o.MgUnpack(c,rl,rr); }); }
int orc_forces_launch(orc_state *s,int interstep){
  (void)interstep;
  return Guard([&]{ Oracle &o=s->o; o.InteractionForces(interstep);
    float f; f=o.ViscDtMax; memcpy(&o.MaxBits[0],&f,4); f=o.LastAce2; memcpy(&o.MaxBits[1],&f,4); f=o.LastVel2; memcpy(&o.MaxBits[2],&f,4); o.MaxBits[3]=0; });
}
void* orc_forces_maxima_ptr(orc_state *s){ return (void*)s->o.MaxBits; }
int orc_forces_finish(orc_state *s,dsph_forceres *res){
  return Guard([&]{ Oracle &o=s->o; float f; memcpy(&f,&o.MaxBits[0],4); o.ViscDtMax=f; memcpy(&f,&o.MaxBits[1],4); o.AceMax=sqrt(double(f)); memcpy(&f,&o.MaxBits[2],4); o.VelMax=sqrtf(f); FillForceRes(o,res); });
}
int orc_domain_info(const dsph_config *cfg,double *scell,double domposmin[3],uint32_t domcells[3]){
  return Guard([&]{ Oracle o; o.Configure(*cfg); if(scell)*scell=o.Scell; if(domposmin){ domposmin[0]=o.DomPosMin.x; domposmin[1]=o.DomPosMin.y; domposmin[2]=o.DomPosMin.z; } if(domcells)for(int i=0;i<3;i++)domcells[i]=o.DomCells[i]; });
}

} // extern "C"
