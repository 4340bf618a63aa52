// driver.cpp -- see driver.h.  Host C++ only; every particle operation goes through Engine (CUDA).
// Loop structure, PART cadence and the values stored in PART headers follow JSphGpuSingle::Run
// (S/JSphGpuSingle.cpp:808-883), JSph::SavePartData (S/JSph.cpp:2581-2711) and JDsOutputTime::GetNextTime
// (S/JDsOutputTime.cpp:157-186).
#include "driver.h"
#include "casefile.h"
#include "bi4.h"
#include "../engine.h"
#include <chrono>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <memory>
#include <stdexcept>
#include <sys/stat.h>

namespace dsph {

namespace {
struct Log {
  FILE *f=nullptr; bool echo=true;
  void Open(const std::string &file){ f=fopen(file.c_str(),"w"); }
  void Printf(const char *fmt,...){
    char tx[2048]; va_list ap; va_start(ap,fmt); vsnprintf(tx,sizeof(tx),fmt,ap); va_end(ap);
    if(echo){ printf("%s\n",tx); fflush(stdout); }
    if(f){ fprintf(f,"%s\n",tx); fflush(f); }
  }
  ~Log(){ if(f)fclose(f); }
};
double NextTime(double t,double tout){
  unsigned nt=unsigned(t/tout);
  double next=tout*nt;
  for(;next<=t;nt++)next=tout*nt;
  return next;
}
}

// PART output: the state is copied off the device by the time loop (one D2H per array), everything else --
// dropping the halo/periodic copies, building the bi4 container and writing it -- runs on a writer thread while the
// simulation goes on.  One job in flight: a new PART waits for the previous file to be complete.
struct PartJob {
  std::string file; PartHead h; PartData p;
  std::vector<unsigned> idp; std::vector<typecode> code; std::vector<double> pos; std::vector<float> vr;
};
class PartWriter {
  std::thread th; std::mutex mx; std::condition_variable cv;
  std::unique_ptr<PartJob> job; bool stop=false; std::string error;
  static void Write(PartJob &j){
    PartData &p=j.p; const size_t np=j.idp.size();
    p.idp.reserve(np); p.pos.reserve(np*3); p.vel.reserve(np*3); p.rhop.reserve(np);
    for(size_t q=0;q<np;q++)if((j.code[q]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL){   // periodic / halo copies are not saved (S/JSphGpuSingle.cpp:897)
      p.idp.push_back(j.idp[q]);
      p.pos.push_back(j.pos[3*q]); p.pos.push_back(j.pos[3*q+1]); p.pos.push_back(j.pos[3*q+2]);
      p.vel.push_back(j.vr[4*q]); p.vel.push_back(j.vr[4*q+1]); p.vel.push_back(j.vr[4*q+2]);
      p.rhop.push_back(j.vr[4*q+3]);
    }
    p.npok=unsigned(p.idp.size());
    SavePartFile(j.file,j.h,p,true);
  }
  void Loop(){
    for(;;){
      std::unique_ptr<PartJob> j;
      { std::unique_lock<std::mutex> lk(mx); cv.wait(lk,[&]{ return job || stop; }); if(!job && stop)return; j=std::move(job); }
      try{ Write(*j); }catch(const std::exception &e){ std::lock_guard<std::mutex> lk(mx); error=e.what(); }
      { std::lock_guard<std::mutex> lk(mx); busy=false; }
      cv.notify_all();
    }
  }
  bool busy=false;
 public:
  PartWriter(){ th=std::thread([this]{ Loop(); }); }
  ~PartWriter(){ try{ Finish(); }catch(...){} { std::lock_guard<std::mutex> lk(mx); stop=true; } cv.notify_all(); if(th.joinable())th.join(); }
  void Submit(std::unique_ptr<PartJob> j){
    std::unique_lock<std::mutex> lk(mx);
    cv.wait(lk,[&]{ return !busy; });
    if(!error.empty())throw std::runtime_error("PART output failed: "+error);
    job=std::move(j); busy=true;
    lk.unlock(); cv.notify_all();
  }
  void Finish(){ std::unique_lock<std::mutex> lk(mx); cv.wait(lk,[&]{ return !busy; }); if(!error.empty()){ std::string e=error; error.clear(); throw std::runtime_error("PART output failed: "+e); } }
};

static void SavePart(Engine &eng,const CaseData &cd,const std::string &dirout,unsigned cpart,double timestep,unsigned step,double runtime,unsigned long long nout,Log &log,PartWriter &writer){
  const unsigned np=eng.np;
  std::unique_ptr<PartJob> j(new PartJob);
  j->idp.resize(np); j->code.resize(np); j->pos.resize(size_t(np)*3); j->vr.resize(size_t(np)*4);
  eng.Download(0,np,j->idp.data(),j->code.data(),nullptr,j->pos.data(),j->vr.data());
  PartHead &h=j->h; PartData &p=j->p;
  const dsph_config &c=cd.cfg;
  h.casename=cd.name; h.data2d=c.simulate2d!=0; h.data2dposy=c.simulate2dposy;
  h.casenp=c.casenp; h.casenfixed=c.casenfixed; h.casenmoving=c.casenmoving; h.casenfloat=c.casenfloat; h.casenfluid=c.casenfluid;
  for(int i=0;i<3;i++){ h.caseposmin[i]=cd.caseposmin[i]; h.caseposmax[i]=cd.caseposmax[i]; h.mapposmin[i]=eng.k.maprealposmin[i]; h.mapposmax[i]=eng.k.maprealposmax[i];
    h.perixinc[i]=eng.k.perixinc[i]; h.periyinc[i]=eng.k.periyinc[i]; h.perizinc[i]=eng.k.perizinc[i]; }
  h.perimode=eng.k.periactive;
  h.dp=c.dp; h.h=c.kernelh; h.b=c.cteb; h.rhop0=c.rhopzero; h.gamma=c.gamma; h.massbound=c.massbound; h.massfluid=c.massfluid;
  p.cpart=cpart; p.timestep=timestep; p.step=step; p.runtime=runtime; p.nout=unsigned(nout);
  p.symplecticdtpre=(c.tstep==DSPH_STEP_SYMPLECTIC? eng.symplecticdtpre: 0);
  const double sc=eng.k.scell;
  for(int i=0;i<3;i++){ p.domainmin[i]=eng.k.domposmin[i]+sc*eng.dp.cellmin[i]; p.domainmax[i]=eng.k.domposmin[i]+sc*(eng.dp.cellmax[i]+1); }
  j->file=dirout+"/"+PartFileName(cpart);
  const unsigned nsave=np-eng.npbper-eng.npfper;
  writer.Submit(std::move(j));
  log.Printf("Part_%04u  %12.6f  %12u  %10u particles stored",cpart,timestep,step,nsave);
}

// Probe points: <dir>/<casename>_Points.txt, one "x y z" per line (separators blank , ;  comments #), the file format of the
// reference examples' CaseDambreak_Points*.txt lists for MeasureTool.  Results go to <dirout>/MeasurePoints.csv, one row per
// PART and point: pressure, velocity and kernel sum from dsph_measure_points' kernel (measure.cuh).
static std::vector<double> LoadPoints(const std::string &file){
  std::vector<double> pts;
  FILE *f=fopen(file.c_str(),"r");
  if(!f)return pts;
  char line[512];
  while(fgets(line,sizeof(line),f)){
    for(char *q=line;*q;q++)if(*q==','||*q==';'||*q=='\t')*q=' ';
    if(line[0]=='#')continue;
    double x,y,z;
    if(sscanf(line,"%lf %lf %lf",&x,&y,&z)==3){ pts.push_back(x); pts.push_back(y); pts.push_back(z); }
  }
  fclose(f);
  return pts;
}
static void SaveProbes(Engine &eng,const std::vector<double> &pts,const std::string &dirout,unsigned cpart,double timestep){
  const unsigned n=unsigned(pts.size()/3);
  if(!n)return;
  std::vector<float> press(n),vel(size_t(n)*3),kc(n);
  eng.MeasurePoints(n,pts.data(),press.data(),vel.data(),kc.data());
  const std::string file=dirout+"/MeasurePoints.csv";
  FILE *f=fopen(file.c_str(),cpart? "a": "w");
  if(!f)return;
  if(!cpart)fprintf(f,"Part;Time [s];Point;PosX [m];PosY [m];PosZ [m];Press [Pa];VelX [m/s];VelY [m/s];VelZ [m/s];Kcorr\n");
  for(unsigned i=0;i<n;i++)fprintf(f,"%u;%.9g;%u;%.9g;%.9g;%.9g;%.7g;%.7g;%.7g;%.7g;%.7g\n",cpart,timestep,i,pts[3*i],pts[3*i+1],pts[3*i+2],
    press[i],vel[3*i],vel[3*i+1],vel[3*i+2],kc[i]);
  fclose(f);
}

RunSummary RunCase(const RunOptions &opt){
  RunSummary sum;
  mkdir(opt.dirout.c_str(),0755);
  Log log; log.echo=opt.verbose; log.Open(opt.dirout+"/Run.out");
  log.Printf("dsph_b200 (B200-native DualSPHysics hot path)  case=%s/%s",opt.dir.c_str(),opt.casename.c_str());
  CaseData cd; LoadCase(cd,opt.dir,opt.casename);
  dsph_config &c=cd.cfg;
  if(opt.timemax>0)c.timemax=opt.timemax;
  if(opt.timeout>0)c.timeout=opt.timeout;
  double timestep=0; unsigned part=0; double restartdtpre=0;
  if(opt.partbegin>=0){
    // restart: particle state, time and the Symplectic dt of the next step come from a PART file of a previous run
    // (JPartsLoad4::LoadParticles with partbegin + JSph::RestartCheckData, S/JSph.cpp:2005-2040); excluded particles stay excluded
    PartHead ph; PartData pd;
    const std::string dirb=(opt.partbegindir.empty()? opt.dirout: opt.partbegindir);
    LoadPartFile(dirb+"/"+PartFileName(unsigned(opt.partbegin)),ph,pd,unsigned(opt.partbegin));
    if(ph.casenp!=c.casenp || ph.casenfixed!=c.casenfixed || ph.casenfluid!=c.casenfluid)throw std::runtime_error("The PART file of -partbegin does not belong to this case.");
    cd.idp=pd.idp; cd.pos=pd.pos; cd.velrhop.resize(size_t(pd.npok)*4);
    for(unsigned i=0;i<pd.npok;i++){ cd.velrhop[4*i]=pd.vel[3*i]; cd.velrhop[4*i+1]=pd.vel[3*i+1]; cd.velrhop[4*i+2]=pd.vel[3*i+2]; cd.velrhop[4*i+3]=pd.rhop[i]; }
    timestep=pd.timestep; part=unsigned(opt.partbegin); restartdtpre=pd.symplecticdtpre;
    log.Printf("Restart from %s/%s: TimeStep=%.9g  np=%u",dirb.c_str(),PartFileName(part).c_str(),timestep,pd.npok);
  }
  const unsigned np0=cd.np();
  // JSph::CheckRhopLimits (S/JSph.cpp:2021-2030): the initial density of every fluid particle must be inside RhopOut
  if(c.rhopoutmin<c.rhopoutmax)for(unsigned i=0;i<np0;i++)if(cd.idp[i]>=c.casenfixed+c.casenmoving+c.casenfloat){
    const float r=cd.velrhop[4*size_t(i)+3];
    if(r<c.rhopoutmin || c.rhopoutmax<r)throw std::runtime_error("Initial fluid density is out of limits. *** To change the limits modify the value of 'RhopOutMin' and 'RhopOutMax' in the XML file (parameters section).");
  }
  unsigned capacity=std::max(np0,c.casenp)+c.casenp/20+1024;
  if(c.perix||c.periy||c.periz)capacity=unsigned(np0*1.7)+65536;
  Engine eng(opt.device,c,capacity);
  if(opt.forcevariant)eng.forcevariant=opt.forcevariant;
  log.Printf("Particles: np=%u nbound=%u nfluid=%u  dp=%g h=%g cs0=%g  cells(map)=%ux%ux%u",np0,c.casenfixed,c.casenfluid,c.dp,double(c.kernelh),eng.k.cs0,eng.k.mapcells[0],eng.k.mapcells[1],eng.k.mapcells[2]);
  std::vector<typecode> code(np0); ComputeCodes(cd,np0,cd.idp.data(),code.data());
  unsigned npb0=0; for(unsigned i=0;i<np0;i++)if(cd.idp[i]<c.casenfixed+c.casenmoving)npb0++;     // boundary ids come first (CheckSortParticles)
  eng.Upload(np0,npb0,cd.idp.data(),code.data(),cd.pos.data(),cd.velrhop.data());
  if(restartdtpre>0)eng.symplecticdtpre=restartdtpre;
  dsph_divinfo di; eng.CellDivide(true,&di);
  unsigned long long nstep=0;
  const auto t0=std::chrono::steady_clock::now();
  auto Elapsed=[&]{ return std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); };
  const std::vector<double> probes=LoadPoints(opt.dir+"/"+opt.casename+"_Points.txt");
  if(!probes.empty())log.Printf("Probe points: %u (%s_Points.txt) -> MeasurePoints.csv",unsigned(probes.size()/3),opt.casename.c_str());
  PartWriter writer;
  if(opt.partbegin<0){
    if(opt.savepart)SavePart(eng,cd,opt.dirout,part,timestep,0,0,0,log,writer);
    SaveProbes(eng,probes,opt.dirout,part,timestep);
  }
  part++;
  double timepartnext=NextTime(timestep,c.timeout);
  eng.GetTimers(nullptr,true);
  const auto tsim0=std::chrono::steady_clock::now();
  unsigned long long partsteps=0;
  while(timestep<c.timemax){
    double dt; dsph_forceres fr;
    eng.Step(&dt,&fr);
    partsteps+=eng.np;
    timestep+=dt; nstep++;
    const unsigned long long nout=eng.npout_total;
    const bool partoutstop=(c.partsoutmax_on && double(nout)>double(c.partsoutmax)*double(c.casenfluid)) || !eng.np;
    if(timestep>=timepartnext || partoutstop){
      if(opt.savepart)SavePart(eng,cd,opt.dirout,part,timestep,unsigned(nstep),Elapsed(),nout,log,writer);
      else log.Printf("Part_%04u  %12.6f  %12llu  dt=%.6g  np=%u",part,timestep,nstep,dt,eng.np);
      SaveProbes(eng,probes,opt.dirout,part,timestep);
      part++;
      timepartnext=NextTime(timestep,c.timeout);
      if(partoutstop){ log.Printf("Particles OUT limit reached..."); break; }
    }
    if(opt.nsteps>0 && nstep>=(unsigned long long)opt.nsteps)break;
  }
  const double tsim=std::chrono::duration<double>(std::chrono::steady_clock::now()-tsim0).count();
  writer.Finish();                                   // the last PART file must be complete before the run reports success
  dsph_timers tm; eng.GetTimers(&tm,false);
  sum.np0=np0; sum.steps=nstep; sum.simtime=timestep; sum.runtime_s=tsim; sum.parts=part; sum.npout=eng.npout_total;
  sum.mpartsteps_s=(tsim>0? double(partsteps)/tsim/1e6: 0);
  log.Printf("[Simulation finished]  steps=%llu  TimeStep=%.9g  runtime=%.4f s  %.3f Mparticle-steps/s  excluded=%llu",nstep,timestep,tsim,sum.mpartsteps_s,sum.npout);
  log.Printf("[GPU Timers] (device ms)");
  log.Printf("NL-Limits........: %.3f",tm.nl_limits); log.Printf("NL-PreSort.......: %.3f",tm.nl_presort);
  log.Printf("NL-RadixSort.....: %.3f",tm.nl_radixsort); log.Printf("NL-CellBegin.....: %.3f",tm.nl_cellbegin);
  log.Printf("NL-SortData......: %.3f",tm.nl_sortdata); log.Printf("NL-OutCheck......: %.3f",tm.nl_outcheck);
  log.Printf("CF-PreForces.....: %.3f",tm.cf_preforces); log.Printf("CF-Forces........: %.3f",tm.cf_forces);
  log.Printf("SU-ComputeStep...: %.3f",tm.su_computestep); log.Printf("SU-Periodic......: %.3f",tm.su_periodic);
  log.Printf("Kernel launches..: %llu",(unsigned long long)tm.kernel_launches);
  { // Run.csv one-line summary (S/JSph.cpp:2962-3006)
    FILE *f=fopen((opt.dirout+"/Run.csv").c_str(),"w");
    if(f){
      fprintf(f,"#RunName;Np;Steps;TSimul;TimeStep;MPartStepsPerSec;NL-Limits;NL-PreSort;NL-RadixSort;NL-CellBegin;NL-SortData;CF-Forces;SU-ComputeStep\n");
      fprintf(f,"%s;%u;%llu;%.6f;%.9g;%.4f;%.3f;%.3f;%.3f;%.3f;%.3f;%.3f;%.3f\n",opt.casename.c_str(),np0,nstep,tsim,timestep,sum.mpartsteps_s,tm.nl_limits,tm.nl_presort,tm.nl_radixsort,tm.nl_cellbegin,tm.nl_sortdata,tm.cf_forces,tm.su_computestep);
      fclose(f);
    }
  }
  return sum;
}

} // namespace dsph
