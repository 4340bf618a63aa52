// ref_harness.cpp -- TEST INFRASTRUCTURE (never part of the product).
//
// Compiled by oracle/Makefile AGAINST THE REFERENCE SOURCES (with -I<reference source dir>) and linked
// with the reference's own objects.  It subclasses the reference CPU solver JSphCpuSingle (all members
// used are `protected`, so no reference file is modified or copied) and replays
// JSphCpuSingle::Run (S/JSphCpuSingle.cpp:1049-1122) while dumping, in binary, the intermediate
// quantities the parity tests need: cell assignment + sorted order + begin-cell table after
// RunCellDivide, Ace/Ar/ViscDtMax/AceMax/VelMax/dt after Interaction_Forces, and the state after
// ComputeSymplecticPre/Corr or ComputeVerlet.  The step sequence below follows
// JSphCpuSingle::ComputeStep_Sym/_Ver (S/JSphCpuSingle.cpp:674-721) call by call.
//
//   ref_harness52 <case> <dirout> -cpu [reference options...] -hdir:<dumpdir> -hsteps:0,1,50 -hnsteps:51
//
#include "JSphCpuSingle.h"
#include "JSphCfgRun.h"
#include "JCellDivCpuSingle.h"
#include "JAppInfo.h"
#include "JLog2.h"
#include "JException.h"
#ifndef HARNESS_NN
#include "JDsTimersCpu.h"
#endif
#include "JArraysCpu.h"
#include "JSphShifting.h"
#include "JDsOutputTime.h"
#ifndef HARNESS_NN
#include "JDsGaugeSystem.h"
#include "JDsMotion.h"
#include "JDsGaugeItem.h"
#endif
#include "Functions.h"
#include "ref_dump.h"
#include <set>
#include <string>
#include <vector>
#include <cstdlib>

JAppInfo AppInfo("DualSPHysics-refharness","v0","","","");

class RefHarness : public JSphCpuSingle {
 public:
  std::string DumpDir;
  std::set<unsigned> DumpSteps;
  unsigned NstepsRun=0;

  std::string DumpName(unsigned step,const char *stage)const{
    return DumpDir+"/"+fun::PrintStr("s%04u_%s.dmp",step,stage);
  }
  //-State of particles in current order (after a divide: cell-sorted).
  void DumpState(unsigned step,const char *stage,bool withcells){
    RefDumpWriter w(DumpName(step,stage));
    w.U32("np",Np); w.U32("npb",Npb); w.U32("npbok",NpbOk); w.U32("npbper",NpbPer); w.U32("npfper",NpfPer);
    w.U32("nstep",Nstep); w.F64("timestep",TimeStep); w.F64("symplecticdtpre",SymplecticDtPre);
    w.U32("verletstep",VerletStep);
    w.Add("idp",2,1,Np,Idpc); w.Add("code",1,1,Np,Codec); w.Add("dcell",2,1,Np,Dcellc);
    w.Add("pos",4,3,Np,Posc); w.Add("velrhop",3,4,Np,Velrhopc);
    if(TStep==STEP_Verlet && VelrhopM1c)w.Add("velrhopm1",3,4,Np,VelrhopM1c);
    if(PosPrec)w.Add("pospre",4,3,Np,PosPrec);
    if(VelrhopPrec)w.Add("velrhoppre",3,4,Np,VelrhopPrec);
    if(SpsTauc)w.Add("spstau",3,6,Np,SpsTauc);
#ifndef HARNESS_NN
    if(UseNormals && BoundNormalc)w.Add("boundnormal",3,3,Np,BoundNormalc);   // mDBC: particle -> ghost node, in the current particle order
#endif
    if(withcells){
      const tuint3 cmin=CellDivSingle->GetCellDomainMin(), cmax=CellDivSingle->GetCellDomainMax();
      const unsigned cm[6]={cmin.x,cmin.y,cmin.z,cmax.x,cmax.y,cmax.z};
      w.Add("celldomain",2,6,1,cm);
      const unsigned nct=CellDivSingle->GetNct();
      const unsigned nc[4]={CellDivSingle->GetNcx(),CellDivSingle->GetNcy(),CellDivSingle->GetNcz(),nct};
      w.Add("ncells",2,4,1,nc);
      w.Add("begincell",2,1,size_t(nct)*2+6,CellDivSingle->GetBeginCell());
      const unsigned outs[6]={CellDivSingle->GetNpbOut(),CellDivSingle->GetNpfOut(),CellDivSingle->GetNpbIgnore(),
                              CellDivSingle->GetNpbOutIgnore(),CellDivSingle->GetNpfOutIgnore(),0};
      w.Add("outcounts",2,6,1,outs);
      w.U32("domcellcode",DomCellCode);
      const unsigned dc[3]={DomCells.x,DomCells.y,DomCells.z}; w.Add("domcells",2,3,1,dc);
      const double dpm[3]={DomPosMin.x,DomPosMin.y,DomPosMin.z}; w.Add("domposmin",4,3,1,dpm);
      w.F32("scell",Scell);
    }
  }
  void DumpForces(unsigned step,const char *stage,double dt){
    RefDumpWriter w(DumpName(step,stage));
    w.U32("np",Np); w.U32("npb",Npb); w.U32("npbok",NpbOk);
    w.Add("ace",3,3,Np,Acec); w.Add("ar",3,1,Np,Arc);
#ifndef HARNESS_NN
    if(TBoundary==BC_MDBC)w.Add("velrhop",3,4,Np,Velrhopc);                   // boundary densities after MdbcBoundCorrection
#endif
    if(ShiftPosfsc)w.Add("shiftposfs",3,4,Np,ShiftPosfsc);
    if(SpsTauc)w.Add("spstau",3,6,Np,SpsTauc);
    if(Pressc)w.Add("press",3,1,Np,Pressc);
#ifdef HARNESS_NN
    if(Visco_etac)w.Add("visco_eta",3,1,Np,Visco_etac);
    if(SpsGradvelc)w.Add("gradvel",3,6,Np,SpsGradvelc);
    if(D_tensorc)w.Add("d_tensor",3,6,Np,D_tensorc);
    w.F32("viscetadtmax",ViscEtaDtMax);
#endif
    w.F32("viscdtmax",ViscDtMax); w.F64("acemax",AceMax); w.F64("velmax",VelMax); w.F64("dt",dt);
  }
  void DumpConstants(){
    RefDumpWriter w(DumpDir+"/constants.dmp");
    w.F32("kernelh",KernelH); w.F32("kernelsize",KernelSize); w.F32("kernelsize2",KernelSize2);
    w.F32("awen",KWend.awen); w.F32("bwen",KWend.bwen); w.F32("eta2",Eta2); w.F64("cs0",Cs0);
    w.F32("cteb",CteB); w.F32("gamma",Gamma); w.F32("rhopzero",RhopZero); w.F32("massfluid",MassFluid); w.F32("massbound",MassBound);
    w.F32("ddtkh",DDTkh); w.F32("ddtgz",DDTgz); w.F32("visco",Visco); w.F32("viscoboundfactor",ViscoBoundFactor);
    w.F64("dp",Dp); w.F64("cflnumber",CFLnumber); w.F64("dtini",DtIni); w.F64("dtmin",DtMin);
    w.F32("scell",Scell); w.F32("movlimit",MovLimit); w.U32("scelldiv",unsigned(ScellDiv));
    const double m[18]={Map_PosMin.x,Map_PosMin.y,Map_PosMin.z,Map_PosMax.x,Map_PosMax.y,Map_PosMax.z,
      MapRealPosMin.x,MapRealPosMin.y,MapRealPosMin.z,MapRealPosMax.x,MapRealPosMax.y,MapRealPosMax.z,
      DomPosMin.x,DomPosMin.y,DomPosMin.z,DomPosMax.x,DomPosMax.y,DomPosMax.z};
    w.Add("maplimits",4,18,1,m);
    const unsigned mc[3]={Map_Cells.x,Map_Cells.y,Map_Cells.z}; w.Add("mapcells",2,3,1,mc);
    w.U32("domcellcode",DomCellCode);
    w.U32("tstep",unsigned(TStep)); w.U32("tkernel",unsigned(TKernel)); w.U32("tvisco",unsigned(TVisco));
    w.U32("tdensity",unsigned(TDensity)); w.U32("shiftmode",unsigned(ShiftingMode)); w.U32("periactive",unsigned(PeriActive));
    w.U32("casenp",CaseNp); w.U32("casenfixed",CaseNfixed); w.U32("casenfluid",CaseNfluid);
    const double pi[9]={PeriXinc.x,PeriXinc.y,PeriXinc.z,PeriYinc.x,PeriYinc.y,PeriYinc.z,PeriZinc.x,PeriZinc.y,PeriZinc.z};
    w.Add("periinc",4,9,1,pi);
#ifdef HARNESS_NN
    w.U32("multiphase",MultiPhase? 1u: 0u); w.U32("tvelgrad",unsigned(TVelGrad)); w.U32("phasecount",PhaseCount); w.F32("lamda",lamda);
    if(MultiPhase){
      std::vector<float> ph;
      for(unsigned c=0;c<PhaseCount;c++){
        const StPhaseArray &a=PhaseArray[c]; const StPhaseCte &k=PhaseCte[c];
        const float v[12]={float(k.mkfluid),a.rho,a.mass,a.Cs0,a.CteB,a.Gamma,k.visco,k.tau_yield,k.tau_max,k.Bi_multi,k.m_NN,k.n_NN};
        ph.insert(ph.end(),v,v+12);
      }
      w.Add("phases",3,12,PhaseCount,ph.data());
    }
#endif
  }

#ifndef HARNESS_NN
  //-Last results of the reference's own gauges (JGaugeSystem, configured from <special><gauges> of the case): type and 9 floats per gauge
  // (vel: point,vel | swl: point0,point2,posswl | maxz: point0,(0,0,zmax) | force: force), computed by RunGaugeSystem where
  // JSphCpuSingle::Run computes them (S/JSphCpuSingle.cpp:1070,1095).
  void DumpGauges(unsigned step,const char *stage,double timestep){
    const unsigned ng=GaugeSystem->GetCount();
    RefDumpWriter w(DumpName(step,stage));
    w.U32("count",ng); w.F64("timestep",timestep);
    std::vector<unsigned> type(ng); std::vector<float> val(size_t(ng)*9,0.f); std::vector<double> tim(ng,0.0);
    for(unsigned c=0;c<ng;c++){
      JGaugeItem *g=GaugeSystem->GetGauge(c);
      type[c]=unsigned(g->Type);
      float *v=&val[size_t(c)*9];
      if(g->Type==JGaugeItem::GAUGE_Vel){ const JGaugeVelocity::StGaugeVelRes &r=((JGaugeVelocity*)g)->GetResult(); tim[c]=r.timestep; v[0]=r.point.x; v[1]=r.point.y; v[2]=r.point.z; v[3]=r.vel.x; v[4]=r.vel.y; v[5]=r.vel.z; }
      else if(g->Type==JGaugeItem::GAUGE_Swl){ const JGaugeSwl::StGaugeSwlRes &r=((JGaugeSwl*)g)->GetResult(); tim[c]=r.timestep; v[0]=r.point0.x; v[1]=r.point0.y; v[2]=r.point0.z; v[3]=r.point2.x; v[4]=r.point2.y; v[5]=r.point2.z; v[6]=r.posswl.x; v[7]=r.posswl.y; v[8]=r.posswl.z; }
      else if(g->Type==JGaugeItem::GAUGE_MaxZ){ const JGaugeMaxZ::StGaugeMaxzRes &r=((JGaugeMaxZ*)g)->GetResult(); tim[c]=r.timestep; v[0]=r.point0.x; v[1]=r.point0.y; v[2]=r.point0.z; v[5]=r.zmax; }
      else if(g->Type==JGaugeItem::GAUGE_Force){ const JGaugeForce::StGaugeForceRes &r=((JGaugeForce*)g)->GetResult(); tim[c]=r.timestep; v[0]=r.force.x; v[1]=r.force.y; v[2]=r.force.z; }
    }
    w.Add("type",2,1,ng,type.data()); w.Add("values",3,9,ng,val.data()); w.Add("time",4,1,ng,tim.data());
  }
#endif

#ifndef HARNESS_NN
  std::vector<double> MotionLogT,MotionLogV;
  //-What JDsMotion computed for this step (StMotionData of every moving object): type, linmov, linvel, matmov.
  void DumpMotion(unsigned step){
    RefDumpWriter w(DumpName(step,"motion"));
    const unsigned nref=DsMotion->GetNumObjects();
    w.U32("nobj",nref);
    std::vector<unsigned> head; std::vector<double> val;
    for(unsigned ref=0;ref<nref;ref++){
      const StMotionData& m=DsMotion->GetMotionData(ref);
      head.push_back(m.idbegin); head.push_back(m.count); head.push_back(unsigned(m.type));
      const double lin[6]={m.linmov.x,m.linmov.y,m.linmov.z,m.linvel.x,m.linvel.y,m.linvel.z};
      val.insert(val.end(),lin,lin+6);
      const tmatrix4d &t=m.matmov;
      const double mt[16]={t.a11,t.a12,t.a13,t.a14,t.a21,t.a22,t.a23,t.a24,t.a31,t.a32,t.a33,t.a34,t.a41,t.a42,t.a43,t.a44};
      val.insert(val.end(),mt,mt+16);
    }
    w.Add("head",2,1,head.size(),head.data()); w.Add("val",4,1,val.size(),val.data());
  }
#endif

  //-Same call sequence as JSphCpuSingle::ComputeStep_Sym, with dumps in between.
  double StepSymDump(unsigned step){
    const double dt=SymplecticDtPre;
#ifndef HARNESS_NN
    if(CaseNmoving){ CalcMotion(dt); DumpMotion(step); }
#endif
    Interaction_Forces(INTERSTEP_SymPredictor);
    const double ddt_p=DtVariable(false);
    DumpForces(step,"fpre",ddt_p);
    if(Shifting)RunShifting(dt*.5);
    ComputeSymplecticPre(dt);
    PosInteraction_Forces();
    DumpState(step,"upre",false);
    RunCellDivide(true);
    DumpState(step,"div1",true);
    Interaction_Forces(INTERSTEP_SymCorrector);
    const double ddt_c=DtVariable(true);
    DumpForces(step,"fcor",ddt_c);
    if(Shifting)RunShifting(dt);
    ComputeSymplecticCorr(dt);
    PosInteraction_Forces();
    SymplecticDtPre=std::min(ddt_p,ddt_c);
    DumpState(step,"ucor",false);
    return(dt);
  }
  //-Same call sequence as JSphCpuSingle::ComputeStep_Ver, with dumps in between.
  double StepVerDump(unsigned step){
    Interaction_Forces(INTERSTEP_Verlet);
    const double dt=DtVariable(true);
    DumpForces(step,"fver",dt);
#ifndef HARNESS_NN
    if(CaseNmoving){ CalcMotion(dt); DumpMotion(step); }
#endif
    if(Shifting)RunShifting(dt);
    ComputeVerlet(dt);
    PosInteraction_Forces();
    DumpState(step,"uver",false);
    return(dt);
  }

  //-Replays JSphCpuSingle::Run (initialisation and main loop) without PART output bookkeeping.
  void RunDump(std::string appname,const JSphCfgRun *cfg,JLog2 *log){
#ifdef HARNESS_NN
    // M/JSphCpuSingle.cpp:1085-1115
    AppName=appname; Log=log;
    TmcCreation(Timers,cfg->SvTimers);
    TmcStart(Timers,TMC_Init);
    LoadConfig(const_cast<JSphCfgRun*>(cfg));
    LoadCaseParticles();
    VisuConfig();
    ConfigDomain();
    ConfigRunMode(cfg);
    VisuParticleSummary();
    InitRunCpu();
    FreePartsInit();
    UpdateMaxValues();
    SaveData();
    TmcResetValues(Timers);
    TmcStop(Timers,TMC_Init);
#else
    AppName=appname; Log=log; CfgRun=cfg;
    Timersc->Config(cfg->SvTimers);
    Timersc->TmStart(TMC_Init);
    LoadConfig(cfg);
    LoadCaseParticles();
    VisuConfig();
    ConfigDomain();
    ConfigRunMode();
    VisuParticleSummary();
    InitRunCpu();
    RunGaugeSystem(TimeStep,true);
    if(GaugeSystem->GetCount())DumpGauges(0,"ginit",TimeStep);
    FreePartsInit();
    UpdateMaxValues();
    SaveData();
    Timersc->ResetTimes();
    Timersc->TmStop(TMC_Init);
#endif
    PartNstep=-1; Part++;
    DumpConstants();
    TimerSim.Start(); TimerPart.Start();
    while(TimeStep<TimeMax && Nstep<int(NstepsRun)){
      const unsigned step=unsigned(Nstep);
      const bool dump=(DumpSteps.count(step)!=0);
      InterStep=(TStep==STEP_Symplectic? INTERSTEP_SymPredictor: INTERSTEP_Verlet);
      double stepdt;
      if(dump){
        DumpState(step,"div0",true);
        stepdt=(TStep==STEP_Symplectic? StepSymDump(step): StepVerDump(step));
      }
      else stepdt=ComputeStep();
#ifndef HARNESS_NN
      if(GaugeSystem->GetCount()){ RunGaugeSystem(TimeStep+stepdt); if(dump)DumpGauges(step,"gauges",TimeStep+stepdt); }
#endif
#ifndef HARNESS_NN
      if(CaseNmoving){
        //-Log of what JDsMotion computed for every step (the evaluation is stateful: a checker needs all the steps).
        MotionLogT.push_back(TimeStep); MotionLogT.push_back(stepdt);
        for(unsigned ref=0;ref<DsMotion->GetNumObjects();ref++){
          const StMotionData& m=DsMotion->GetMotionData(ref);
          const tmatrix4d &t=m.matmov;
          const double v[23]={double(m.type),m.linmov.x,m.linmov.y,m.linmov.z,m.linvel.x,m.linvel.y,m.linvel.z,
            t.a11,t.a12,t.a13,t.a14,t.a21,t.a22,t.a23,t.a24,t.a31,t.a32,t.a33,t.a34,t.a41,t.a42,t.a43,t.a44};
          MotionLogV.insert(MotionLogV.end(),v,v+23);
        }
        RunMotion(stepdt);
      }
#endif
      RunCellDivide(true);
#ifndef HARNESS_NN
      if(dump && CaseNmoving)DumpState(step,"dmot",true);   //-sorted state after RunMotion + RunCellDivide
#endif
      TimeStep+=stepdt;
      LastDt=stepdt;
      if(dump){
        RefDumpWriter w(DumpName(step,"end"));
        w.F64("stepdt",stepdt); w.F64("timestep",TimeStep); w.F64("symplecticdtpre",SymplecticDtPre);
        w.U32("np",Np); w.U32("npb",Npb); w.U32("npbok",NpbOk);
      }
      Nstep++;
    }
    DumpState(unsigned(Nstep),"final",true);
#ifndef HARNESS_NN
    if(CaseNmoving){
      RefDumpWriter w(DumpName(unsigned(Nstep),"motionlog"));
      w.U32("nobj",DsMotion->GetNumObjects());
      w.Add("t_dt",4,2,MotionLogT.size()/2,MotionLogT.data());
      w.Add("val",4,23,MotionLogV.size()/23,MotionLogV.data());
    }
#endif
    TimerSim.Stop(); TimerTot.Stop();
    log->Printf("refharness: finished %d steps, TimeStep=%.9g",Nstep,TimeStep);
  }
};

int main(int argc,char **argv){
  std::vector<char*> args; RefHarness *sph=new RefHarness();
  std::string dumpdir="dump";
  for(int i=0;i<argc;i++){
    std::string a=argv[i];
    if(a.compare(0,6,"-hdir:")==0)dumpdir=a.substr(6);
    else if(a.compare(0,8,"-hsteps:")==0){
      std::string s=a.substr(8); size_t p=0;
      while(p<s.size()){ size_t q=s.find(',',p); if(q==std::string::npos)q=s.size(); sph->DumpSteps.insert(unsigned(atoi(s.substr(p,q-p).c_str()))); p=q+1; }
    }
    else if(a.compare(0,9,"-hnsteps:")==0)sph->NstepsRun=unsigned(atoi(a.substr(9).c_str()));
    else args.push_back(argv[i]);
  }
  sph->DumpDir=dumpdir;
  fun::MkdirPath(dumpdir);
  int errcode=1;
  AppInfo.ConfigRunPaths(argv[0]);
  JLog2 *log=NULL;
  JSphCfgRun cfg;
  try{
    cfg.LoadArgv(int(args.size()),args.data());
    AppInfo.ConfigOutput(cfg.CreateDirs,cfg.CsvSepComa,cfg.DirOut,cfg.DirDataOut);
    AppInfo.LogInit(AppInfo.GetDirOut()+"/Run.out");
    log=AppInfo.LogPtr();
    cfg.Cpu=true;
    sph->RunDump(AppInfo.GetFullName(),&cfg,log);
    errcode=0;
  }
  catch(const char *cad){ printf("\n*** Exception(chr): %s\n",cad); }
  catch(const std::string &e){ printf("\n*** Exception(str): %s\n",e.c_str()); }
  catch(const JException &e){ printf("\n*** JException: %s\n",e.what()); }
  catch(const std::exception &e){ printf("\n*** Exception(exc): %s\n",e.what()); }
  printf("refharness finished (code=%d)\n",errcode);
  return(errcode);
}
