// JSphGpuSingle_b200.cpp -- the reference-side binding of libdsph_b200.so: the file a DualSPHysics maintainer adds to
// src/source to run the per-step hot path through the C ABI of include/dsph_b200.h.
//
// It is compiled AGAINST THE REFERENCE HEADERS (-I<reference>/src/source) and linked with the reference's own, unmodified objects
// (JSph, JSphGpu, JSphGpuSingle, JCellDivGpu... and main.cpp) into oracle/_ref/DualSPHysics5.2_b200dropin by `make -C oracle
// gpu52_b200`.  Nothing of the reference is copied or edited: the recipe compiles JSphGpuSingle.cpp, JSphGpu.cpp and JCellDivGpu.cpp
// as they are (position independent, so that their member functions stay interposable), marks the definitions of the SEVEN member
// functions below as weak (objcopy --weaken-symbol), and the strong definitions of this file take their place at link time.  Every
// other line of the program -- command line, XML and bi4 input, JSph::Run's time loop, DtVariable, PART / Run.out / Run.csv output --
// is the reference's.
//
//   JSphGpuSingle::RunCellDivide        S/JSphGpuSingle.cpp:331-417   -> dsph_celldiv_run (+ context creation and upload at the first call)
//   JSphGpuSingle::Interaction_Forces   S/JSphGpuSingle.cpp:435-486   -> dsph_forces_run
//   JSphGpuSingle::ComputeStep_Sym      S/JSphGpuSingle.cpp:574-596   -> dsph_forces_run / dsph_step_sympre / dsph_step_symcor (same call sequence)
//   JSphGpuSingle::ComputeStep_Ver      S/JSphGpuSingle.cpp:548-560   -> dsph_forces_run / dsph_step_verlet
//   JSphGpuSingle::RunGaugeSystem       S/JSphGpuSingle.cpp:776-786   -> (cases with <gauges> are refused: the reference's gauge objects read its own arrays)
//   JSphGpu::ParticlesDataDown          S/JSphGpu.cpp:556-590         -> dsph_particles_download (what SaveData stores in the PART files)
//   JSphGpu::RunMotion                  S/JSphGpu.cpp:1035-1075       -> dsph_motion_run with the StMotionData of the reference's JDsMotion
//   JCellDivGpu::GetDomainLimits        S/JCellDivGpu.cpp             -> cell domain of dsph_celldiv_last_info (PART header DomainMin/Max)
//
// Particle state lives in the dsph context from the first RunCellDivide on (the reference's own device arrays, allocated by its
// ConfigDomain, are left untouched).  Scope = the scope of the library: DBC and mDBC boundaries, moving boundaries (JSphGpu::RunMotion -> dsph_motion_run), no floating bodies, no inlet,
// no damping / relaxation zones (refused loudly below).
#include "JSphGpuSingle.h"
#include "JCellDivGpuSingle.h"
#include "JSphShifting.h"
#include "JDsGaugeSystem.h"
#include "JDsFixedDt.h"
#include "JDsMotion.h"
#include "JDsPartsOut.h"
#include "JPartsLoad4.h"
#include "JDsTimersGpu.h"
#include "JDsGpuInfo.h"
#include "JLog2.h"
#include "Functions.h"
#include "dsph_b200.h"
#include <cstring>
#include <vector>
#include <string>
#include <algorithm>

using namespace std;

namespace {
dsph_ctx *Ctx=nullptr;             // one solver instance per process, like the reference's main()
dsph_divinfo LastDiv;
std::vector<dsph_typecode> CodeBuf;

void Check(int rc,const char *what){
  if(rc!=DSPH_OK)throw std::string("libdsph_b200: ")+what+": "+dsph_last_error();
}
}

//==============================================================================
/// Neighbour list.  First call (from the reference's ConfigDomain): creates the context from the solver's configuration and uploads the
/// initial particles, which ConfigDomain has just prepared in Idp / AuxPos / Velrhop / Code.
//==============================================================================
void JSphGpuSingle::RunCellDivide(bool updateperiodic){
  if(!Ctx){
    if(CaseNfloat)Run_Exceptioon("libdsph_b200 binding: floating bodies are outside the scope of the library.");
    if(CaseNmoving && (WaveGen || MLPistons))Run_Exceptioon("libdsph_b200 binding: wave generation and multi-layer pistons are outside the scope of the library.");
    if(TBoundary!=BC_DBC && TBoundary!=BC_MDBC)Run_Exceptioon("libdsph_b200 binding: DBC or mDBC boundaries.");
    if(TBoundary==BC_MDBC && (SlipMode!=SLIP_Vel0 || UseNormalsFt))Run_Exceptioon("libdsph_b200 binding: mDBC with slip mode velocity=0 and without floating normals.");
    if(InOut || Damping || RelaxZones || MLPistons || AccInput)Run_Exceptioon("libdsph_b200 binding: inlet/outlet, damping, relaxation zones, pistons and external accelerations are outside the scope of the library.");
    dsph_config c; memset(&c,0,sizeof(c));
    c.tkernel=int(TKernel); c.tstep=int(TStep); c.tvisco=int(TVisco); c.tdensity=int(TDensity);
    c.shiftmode=(Shifting? int(Shifting->GetShiftMode()): 0); c.cellmode=int(CellMode); c.verletsteps=VerletSteps; c.simulate2d=(Simulate2D? 1: 0);
    c.perix=(PeriX? 1: 0); c.periy=(PeriY? 1: 0); c.periz=(PeriZ? 1: 0);
    c.dtallparticles=(DtAllParticles? 1: 0); c.dtfixed_on=0; c.stable=(Stable? 1: 0);
    if(FixedDt)Run_Exceptioon("libdsph_b200 binding: DtFixed from a file is not bound.");
    c.kernelh=KernelH; c.cteb=CteB; c.gamma=Gamma; c.rhopzero=RhopZero; c.massfluid=MassFluid; c.massbound=MassBound;
    c.gravity[0]=Gravity.x; c.gravity[1]=Gravity.y; c.gravity[2]=Gravity.z;
    c.visco=Visco; c.viscoboundfactor=ViscoBoundFactor; c.ddtvalue=DDTValue;
    c.shiftcoef=(Shifting? Shifting->GetShiftCoef(): 0); c.shifttfs=(Shifting? Shifting->GetShiftTFS(): 0);
    c.coefdtmin=CoefDtMin; c.rhopoutmin=(RhopOut? RhopOutMin: 0.f); c.rhopoutmax=(RhopOut? RhopOutMax: 0.f);
    c.dp=Dp; c.cflnumber=CFLnumber; c.dtini=DtIni; c.dtmin=DtMin; c.simulate2dposy=Simulate2DPosY;
    c.timemax=TimeMax; c.timeout=TimePart;
    c.maprealposmin[0]=MapRealPosMin.x; c.maprealposmin[1]=MapRealPosMin.y; c.maprealposmin[2]=MapRealPosMin.z;
    c.maprealposmax[0]=MapRealPosMax.x; c.maprealposmax[1]=MapRealPosMax.y; c.maprealposmax[2]=MapRealPosMax.z;
    c.perixinc[0]=PeriXinc.x; c.perixinc[1]=PeriXinc.y; c.perixinc[2]=PeriXinc.z;
    c.periyinc[0]=PeriYinc.x; c.periyinc[1]=PeriYinc.y; c.periyinc[2]=PeriYinc.z;
    c.perizinc[0]=PeriZinc.x; c.perizinc[1]=PeriZinc.y; c.perizinc[2]=PeriZinc.z;
    c.casenp=CaseNp; c.casenfixed=CaseNfixed; c.casenmoving=CaseNmoving; c.casenfloat=CaseNfloat; c.casenfluid=CaseNfluid;
    c.tboundary=(TBoundary==BC_MDBC? (DSPH_BOUNDARY_MDBC|(MdbcCorrector? DSPH_BOUNDARY_MDBCCORRECTOR: 0)): 0);   // 0 = DBC (the value the library stores for it)
    c.partsoutmax_on=0;                                    // JSph::Run checks Np<NpMinimum itself
    const unsigned capacity=Np+Np/10+1024;
    Check(dsph_ctx_create(&Ctx,GpuInfo->GetGpuId(),&c,capacity),"dsph_ctx_create");
    // initial state as ConfigDomain left it on the host (S/JSphGpuSingle.cpp:158-195): ids, positions, velocity+density, codes
    Check(dsph_particles_upload(Ctx,Np,Npb,Idp,Code,&AuxPos[0].x,&Velrhop[0].x),"dsph_particles_upload");
    if(TBoundary==BC_MDBC){
      // mDBC: the vectors particle -> ghost node that ConfigBoundNormals prepared and ParticlesDataUp left in BoundNormalg (S/JSphGpuSingle.cpp:169-195)
      std::vector<float> nrm(size_t(Npb)*3);
      if(Npb && cudaMemcpy(nrm.data(),BoundNormalg,sizeof(float)*3*Npb,cudaMemcpyDeviceToHost)!=cudaSuccess)Run_Exceptioon("libdsph_b200 binding: cannot read the boundary normals.");
      Check(dsph_particles_set_boundnormals(Ctx,nrm.data(),Npb),"dsph_particles_set_boundnormals");
      Check(dsph_set_mdbc_threshold(Ctx,MdbcThreshold),"dsph_set_mdbc_threshold");
    }
    if(TStep==STEP_Symplectic && SymplecticDtPre)Check(dsph_set_symplectic_dtpre(Ctx,SymplecticDtPre),"dsph_set_symplectic_dtpre");
    Check(dsph_excluded_keep(Ctx,1),"dsph_excluded_keep");
    Log->Print("**libdsph_b200: the cell list, Interaction_Forces and ComputeStep run through the C ABI of dsph_b200.h");
  }
  Timersg->TmStart(TMG_NlSortData,false);
  Check(dsph_celldiv_run(Ctx,updateperiodic? 1: 0,&LastDiv),"dsph_celldiv_run");
  Np=LastDiv.np; Npb=LastDiv.npb; NpbOk=LastDiv.npbok; NpbPer=LastDiv.npbper; NpfPer=LastDiv.npfper;
  Timersg->TmStop(TMG_NlSortData,false);
  //-Control of excluded particles (S/JSphGpuSingle.cpp:405-412).
  if(LastDiv.npfout){
    const unsigned n=LastDiv.npfout;
    if(n>CpuParticlesSize)Run_Exceptioon("libdsph_b200 binding: more excluded particles than the host arrays hold.");
    std::vector<float> vr(size_t(n)*4); unsigned got=0;
    Check(dsph_excluded_take(Ctx,n,&got,Idp,Code,&AuxPos[0].x,vr.data()),"dsph_excluded_take");
    for(unsigned p=0;p<n;p++){ AuxVel[p]=TFloat3(vr[4*p],vr[4*p+1],vr[4*p+2]); AuxRhop[p]=vr[4*p+3]; }
    AddParticlesOut(n,Idp,AuxPos,AuxVel,AuxRhop,Code);
  }
  BoundChanged=false;
}

//==============================================================================
/// Interaction_Forces: the maxima go where JSphGpu::DtVariable reads them (S/JSphGpu.cpp:984-1009).
//==============================================================================
void JSphGpuSingle::Interaction_Forces(TpInterStep interstep){
  InterStep=interstep;
  Timersg->TmStart(TMG_CfForces,true);
  dsph_forceres r;
  Check(dsph_forces_run(Ctx,int(interstep),&r),"dsph_forces_run");
  ViscDtMax=r.viscdtmax; AceMax=r.acemax; VelMax=r.velmax;
  Timersg->TmStop(TMG_CfForces,true);
}

//==============================================================================
/// ComputeStep_Sym: the reference's sequence with the library's updates; dt comes from the reference's own DtVariable.
//==============================================================================
double JSphGpuSingle::ComputeStep_Sym(){
  const double dt=SymplecticDtPre;
  if(CaseNmoving)CalcMotion(dt);               // the reference's JDsMotion (S/JSphGpuSingle.cpp:571)
  //-Predictor
  Interaction_Forces(INTERSTEP_SymPredictor);
  const double ddt_p=DtVariable(false);
  Timersg->TmStart(TMG_SuComputeStep,false);
  Check(dsph_step_sympre(Ctx,dt),"dsph_step_sympre");      // shifting + ComputeSymplecticPre
  Timersg->TmStop(TMG_SuComputeStep,false);
  //-Corrector
  RunCellDivide(true);
  Interaction_Forces(INTERSTEP_SymCorrector);
  const double ddt_c=DtVariable(true);
  Timersg->TmStart(TMG_SuComputeStep,false);
  Check(dsph_step_symcor(Ctx,dt),"dsph_step_symcor");      // shifting + ComputeSymplecticCorr
  Timersg->TmStop(TMG_SuComputeStep,false);
  SymplecticDtPre=min(ddt_p,ddt_c);
  Check(dsph_set_symplectic_dtpre(Ctx,SymplecticDtPre),"dsph_set_symplectic_dtpre");
  return(dt);
}

//==============================================================================
/// ComputeStep_Ver.
//==============================================================================
double JSphGpuSingle::ComputeStep_Ver(){
  Interaction_Forces(INTERSTEP_Verlet);
  const double dt=DtVariable(true);
  if(CaseNmoving)CalcMotion(dt);               // S/JSphGpuSingle.cpp:551
  Timersg->TmStart(TMG_SuComputeStep,false);
  Check(dsph_step_verlet(Ctx,dt),"dsph_step_verlet");
  Timersg->TmStop(TMG_SuComputeStep,false);
  return(dt);
}

//==============================================================================
/// Predefined motion of the moving boundaries (S/JSphGpu.cpp:1035-1075): the reference's JDsMotion has computed the motion data of the
/// step (JSph::CalcMotion, called by ComputeStep_Sym / _Ver below like in the reference); the library applies it to its particles.
//==============================================================================
void JSphGpu::RunMotion(double stepdt){
  BoundChanged=false;
  if(!Ctx || !DsMotion->GetActiveMotion())return;
  const unsigned nref=DsMotion->GetNumObjects();
  std::vector<dsph_motiondata> objs(nref);
  for(unsigned ref=0;ref<nref;ref++){
    const StMotionData& m=DsMotion->GetMotionData(ref);
    dsph_motiondata &d=objs[ref]; memset(&d,0,sizeof(d));
    d.idbegin=m.idbegin; d.count=m.count;
    d.type=(m.type==MOTT_Linear? DSPH_MOTION_LINEAR: (m.type==MOTT_Matrix? DSPH_MOTION_MATRIX: DSPH_MOTION_NONE));
    d.linmov[0]=m.linmov.x; d.linmov[1]=m.linmov.y; d.linmov[2]=m.linmov.z;
    d.linvel[0]=m.linvel.x; d.linvel[1]=m.linvel.y; d.linvel[2]=m.linvel.z;
    const tmatrix4d &t=m.matmov;
    const double mt[16]={t.a11,t.a12,t.a13,t.a14,t.a21,t.a22,t.a23,t.a24,t.a31,t.a32,t.a33,t.a34,t.a41,t.a42,t.a43,t.a44};
    memcpy(d.matmov,mt,sizeof(mt));
  }
  if(dsph_motion_run(Ctx,objs.data(),nref,stepdt)!=DSPH_OK)Run_Exceptioon(string("libdsph_b200: ")+dsph_last_error());
  BoundChanged=true;
}

//==============================================================================
/// The reference's gauge objects read the reference's own device arrays and cell data, which are not maintained here.
//==============================================================================
void JSphGpuSingle::RunGaugeSystem(double timestep,bool saveinput){
  if(GaugeSystem->GetCount())Run_Exceptioon("libdsph_b200 binding: cases with <gauges> run through dsph_gauges_* (bin/dsph_b200), not through this binding.");
}

//==============================================================================
/// What SaveData stores: ids, positions, velocity and density of the particles of the context (S/JSphGpu.cpp:556-590).
//==============================================================================
unsigned JSphGpu::ParticlesDataDown(unsigned n,unsigned pini,bool code,bool onlynormal){
  if(!Ctx)return(0);
  if(n>CpuParticlesSize)Run_Exceptioon("libdsph_b200 binding: the host arrays are smaller than the number of particles.");
  if(CodeBuf.size()<n)CodeBuf.resize(n);
  if(dsph_particles_download(Ctx,pini,n,Idp,CodeBuf.data(),NULL,&AuxPos[0].x,&Velrhop[0].x)!=DSPH_OK)Run_Exceptioon(string("libdsph_b200: ")+dsph_last_error());
  unsigned num=n;
  if(code || onlynormal)for(unsigned p=0;p<n;p++)Code[p]=CodeBuf[p];
  if(onlynormal){
    unsigned ndel=0;
    for(unsigned p=0;p<n;p++){
      const bool normal=CODE_IsNormal(Code[p]);
      if(ndel && normal){ Idp[p-ndel]=Idp[p]; AuxPos[p-ndel]=AuxPos[p]; Velrhop[p-ndel]=Velrhop[p]; Code[p-ndel]=Code[p]; }
      if(!normal)ndel++;
    }
    num-=ndel;
  }
  for(unsigned p=0;p<n;p++){
    Posxy[p]=TDouble2(AuxPos[p].x,AuxPos[p].y); Posz[p]=AuxPos[p].z;
    AuxVel[p]=TFloat3(Velrhop[p].x,Velrhop[p].y,Velrhop[p].z);
    AuxRhop[p]=Velrhop[p].w;
  }
  return(num);
}

//==============================================================================
/// Limits of the current cell domain (DomainMin / DomainMax of the PART header).
//==============================================================================
tdouble3 JCellDivGpu::GetDomainLimits(bool limitmin,unsigned slicecellmin)const{
  if(!Ctx)return(TDouble3(0));
  const double sc=LastDiv.scell;
  if(limitmin)return(TDouble3(LastDiv.domposmin[0]+sc*LastDiv.cellmin[0],LastDiv.domposmin[1]+sc*LastDiv.cellmin[1],LastDiv.domposmin[2]+sc*LastDiv.cellmin[2]));
  return(TDouble3(LastDiv.domposmin[0]+sc*(LastDiv.cellmax[0]+1),LastDiv.domposmin[1]+sc*(LastDiv.cellmax[1]+1),LastDiv.domposmin[2]+sc*(LastDiv.cellmax[2]+1)));
}
