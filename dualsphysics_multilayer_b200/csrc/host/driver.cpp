// driver.cpp -- see driver.h.  Host C++ only; every particle operation goes through Engine (CUDA).
// Loop structure, PART cadence and the values stored in PART headers follow JSphGpuSingle::Run
// (S/JSphGpuSingle.cpp:808-883), JSph::SavePartData (S/JSph.cpp:2581-2711) and JDsOutputTime::GetNextTime
// (S/JDsOutputTime.cpp:157-186).
#include "driver.h"
#include "casefile.h"
#include "bi4.h"
#include "partset.h"
#include "../engine.h"
#include <chrono>
#include <cstdio>
#include <cstdarg>
#include <cmath>
#include <algorithm>
#include <vector>
#include <deque>
#include <atomic>
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

// PART output (S/JSphGpuSingle.cpp:889-935 SaveData, S/JSph.cpp:2581-2711 SavePartData).  The time loop only enqueues
// a device-side copy of the state (Engine::SnapshotBegin, ~0.3 ms per 20 M particles); the D2H transfer into one of two pinned host
// buffers runs on a copy stream and -- dropping the halo/periodic copies, building the bi4 containers, writing Part_NNNN.bi4 (or the
// pieces), PartInfo.ibi4 and PartOut_BBB.obi4 -- on a writer thread, all while the simulation goes on.  A new PART only waits when
// both host buffers are still in use.
struct PartJob {
  int slot=-1; PartMeta meta; PartInfoPlus info;
  std::vector<unsigned> outidp; std::vector<typecode> outcode; std::vector<double> outpos; std::vector<float> outvelrhop;
  float rhopmin=0,rhopmax=0;
};
class PartWriter {
  Engine &eng; PartSet &pset;
  Engine::HostSnap snap[2]; bool inuse[2]={false,false};
  std::thread th; std::mutex mx; std::condition_variable cv;
  std::deque<std::unique_ptr<PartJob>> queue; bool stop=false; std::string error;
  std::vector<float> rhop;
  void Write(PartJob &j){
    Engine::HostSnap &hs=snap[j.slot];
    eng.SnapshotWait(hs);
    // in-place compaction of the pinned buffers: periodic / halo copies are not saved (S/JSphGpuSingle.cpp:897)
    const size_t np=hs.np; size_t n=0;
    rhop.resize(np);
    float *vel=hs.velrhop;                          // xyzw -> packed xyz, moving towards lower addresses only
    for(size_t q=0;q<np;q++)if((hs.code[q]&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_NORMAL){
      hs.idp[n]=hs.idp[q];
      hs.pos[3*n]=hs.pos[3*q]; hs.pos[3*n+1]=hs.pos[3*q+1]; hs.pos[3*n+2]=hs.pos[3*q+2];
      const float vx=hs.velrhop[4*q],vy=hs.velrhop[4*q+1],vz=hs.velrhop[4*q+2],r=hs.velrhop[4*q+3];
      vel[3*n]=vx; vel[3*n+1]=vy; vel[3*n+2]=vz; rhop[n]=r;
      n++;
    }
    pset.SavePart(j.meta,unsigned(n),hs.idp,hs.pos,vel,rhop.data(),&j.info);
    // excluded particles since the previous PART
    const unsigned nout=unsigned(j.outidp.size());
    std::vector<float> ovel(size_t(nout)*3),orhop(nout); std::vector<unsigned char> motive(nout);
    for(unsigned q=0;q<nout;q++){
      ovel[3*q]=j.outvelrhop[4*q]; ovel[3*q+1]=j.outvelrhop[4*q+1]; ovel[3*q+2]=j.outvelrhop[4*q+2]; orhop[q]=j.outvelrhop[4*q+3];
      const unsigned sp=j.outcode[q]&DSPH_CODE_MASKSPECIAL;
      motive[q]=(sp==DSPH_CODE_OUTPOS? 1: (sp==DSPH_CODE_OUTRHOP? 2: (sp==DSPH_CODE_OUTMOVE? 3: 0)));
    }
    pset.SaveOut(j.meta.cpart,j.meta.timestep,nout,j.outidp.data(),j.outpos.data(),ovel.data(),orhop.data(),motive.data(),j.rhopmin,j.rhopmax);
  }
  void Loop(){
    for(;;){
      std::unique_ptr<PartJob> j;
      { std::unique_lock<std::mutex> lk(mx); cv.wait(lk,[&]{ return !queue.empty() || stop; }); if(queue.empty())return; j=std::move(queue.front()); queue.pop_front(); }
      try{ Write(*j); }catch(const std::exception &e){ std::lock_guard<std::mutex> lk(mx); if(error.empty())error=e.what(); }
      { std::lock_guard<std::mutex> lk(mx); inuse[j->slot]=false; }
      cv.notify_all();
    }
  }
 public:
  double waited_s=0;      // time the simulation loop spent waiting for a free host buffer
  PartWriter(Engine &e,PartSet &ps,unsigned capacity):eng(e),pset(ps){
    for(int i=0;i<2;i++)eng.HostSnapAlloc(snap[i],capacity);
    th=std::thread([this]{ Loop(); });
  }
  ~PartWriter(){ try{ Finish(); }catch(...){} { std::lock_guard<std::mutex> lk(mx); stop=true; } cv.notify_all(); if(th.joinable())th.join(); for(int i=0;i<2;i++)eng.HostSnapFree(snap[i]); }
  // called by the time loop: takes a free host buffer, enqueues the snapshot copies and hands the rest to the writer thread
  void Submit(std::unique_ptr<PartJob> j){
    std::unique_lock<std::mutex> lk(mx);
    const auto t0=std::chrono::steady_clock::now();
    cv.wait(lk,[&]{ return !inuse[0] || !inuse[1]; });
    waited_s+=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
    if(!error.empty())throw std::runtime_error("PART output failed: "+error);
    j->slot=(!inuse[0]? 0: 1); inuse[j->slot]=true;
    lk.unlock();
    eng.SnapshotBegin(snap[j->slot]);
    lk.lock(); queue.push_back(std::move(j)); lk.unlock();
    cv.notify_all();
  }
  void Finish(){ std::unique_lock<std::mutex> lk(mx); cv.wait(lk,[&]{ return !inuse[0] && !inuse[1] && queue.empty(); }); if(!error.empty()){ std::string e=error; error.clear(); throw std::runtime_error("PART output failed: "+e); } }
};

struct PartStats { double dtmin=1e300, dtmax=0, timem1=0; unsigned long long stepm1=0; };

static void SavePart(Engine &eng,const CaseData &cd,unsigned cpart,double timestep,unsigned long long step,double runtime,unsigned long long nout,PartStats &st,Log &log,PartWriter &writer){
  std::unique_ptr<PartJob> j(new PartJob);
  const dsph_config &c=cd.cfg;
  PartMeta &m=j->meta;
  m.cpart=cpart; m.timestep=timestep; m.step=unsigned(step); m.runtime=runtime; m.nout=unsigned(nout);
  m.symplecticdtpre=(c.tstep==DSPH_STEP_SYMPLECTIC? eng.symplecticdtpre: 0);
  const double sc=eng.k.scell;
  for(int i=0;i<3;i++){ m.domainmin[i]=eng.k.domposmin[i]+sc*eng.dp.cellmin[i]; m.domainmax[i]=eng.k.domposmin[i]+sc*(eng.dp.cellmax[i]+1); }
  PartInfoPlus &f=j->info; f.on=true;
  f.dtmean=(step>st.stepm1? (timestep-st.timem1)/double(step-st.stepm1): 0); f.dtmin=(step? st.dtmin: 0); f.dtmax=(step? st.dtmax: 0);
  f.timesim=runtime; f.nct=eng.dp.nct; f.npbin=eng.npbok; f.npbout=eng.npb-eng.npbok; f.npf=eng.np-eng.npb; f.npbper=eng.npbper; f.npfper=eng.npfper;
  f.npalloc=(long long)eng.cap; f.npused=(long long)eng.np; f.nctalloc=(long long)eng.maxnctt; f.nctused=(long long)eng.dp.nctt;
  st.dtmin=1e300; st.dtmax=0; st.timem1=timestep; st.stepm1=step;
  j->outidp.swap(eng.outidp); j->outcode.swap(eng.outcode); j->outpos.swap(eng.outpos); j->outvelrhop.swap(eng.outvelrhop);
  j->rhopmin=(c.rhopoutmin<c.rhopoutmax? c.rhopoutmin: 0); j->rhopmax=(c.rhopoutmin<c.rhopoutmax? c.rhopoutmax: 0);
  const unsigned nsave=eng.np-eng.npbper-eng.npfper;
  writer.Submit(std::move(j));
  log.Printf("Part_%04u  %12.6f  %12llu  %10u particles stored",cpart,timestep,step,nsave);
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

// Gauge results -> Gauges<Type>_<name>.csv in the output directory, columns and number format of JGauge*::SaveResults
// (S/JDsGaugeItem.cpp:276-298,489-512,694-714,936-957); rows are buffered and written at every PART like JGaugeSystem::SaveResults(Part).
struct GaugeOutput {
  std::vector<dsph_gauge> defs; std::vector<double> outnext; std::vector<std::string> buf; std::vector<bool> started; std::string dir;
  void Config(const std::vector<dsph_gauge> &g,const std::string &dirout){ defs=g; outnext.assign(g.size(),0.0); buf.assign(g.size(),std::string()); started.assign(g.size(),false); dir=dirout; }
  void Collect(const std::vector<dsph_gaugeres> &res){
    for(size_t i=0;i<defs.size() && i<res.size();i++){
      const dsph_gauge &g=defs[i]; const dsph_gaugeres &r=res[i];
      if(!r.computed || !g.output)continue;
      const double t=r.timestep;
      if(!(t>=outnext[i] && g.outputstart<=t && t<=g.outputend))continue;              // JGaugeItem::Output (S/JDsGaugeItem.h:165)
      char tx[512];
      const float *v=r.value;
      if(g.type==DSPH_GAUGE_VEL)snprintf(tx,sizeof(tx),"%g;%g;%g;%g;%g;%g;%g\n",t,v[0],v[1],v[2],float(g.point0[0]),float(g.point0[1]),float(g.point0[2]));
      else if(g.type==DSPH_GAUGE_SWL)snprintf(tx,sizeof(tx),"%g;%g;%g;%g;%g;%g;%g;%g;%g;%g\n",t,v[0],v[1],v[2],float(g.point0[0]),float(g.point0[1]),float(g.point0[2]),float(g.point2[0]),float(g.point2[1]),float(g.point2[2]));
      else if(g.type==DSPH_GAUGE_MAXZ)snprintf(tx,sizeof(tx),"%g;%g;%g;%g;%g\n",t,v[2],float(g.point0[0]),float(g.point0[1]),float(g.point0[2]));
      else snprintf(tx,sizeof(tx),"%g;%g;%g;%g;%g\n",t,std::sqrt(v[0]*v[0]+v[1]*v[1]+v[2]*v[2]),v[0],v[1],v[2]);
      buf[i]+=tx;
      if(g.outputdt){ const unsigned nt=unsigned(t/g.outputdt); outnext[i]=g.outputdt*nt; if(outnext[i]<=t)outnext[i]=g.outputdt*(nt+1); }
    }
  }
  void Save(){
    static const char *tn[5]={"","Vel","SWL","MaxZ","Force"};
    static const char *hd[5]={"","time [s];velx [m/s];vely [m/s];velz [m/s];posx [m];posy [m];posz [m]","time [s];swlx [m];swly [m];swlz [m];pos0x [m];pos0y [m];pos0z [m];pos2x [m];pos2y [m];pos2z [m]",
                              "time [s];zmax [m];posx [m];posy [m];posz [m]","time [s];force [N];forcex [N];forcey [N];forcez [N]"};
    for(size_t i=0;i<defs.size();i++)if(!buf[i].empty()){
      const std::string file=dir+"/Gauges"+tn[defs[i].type]+"_"+defs[i].name+".csv";
      FILE *f=fopen(file.c_str(),started[i]? "a": "w");
      if(!f)continue;
      if(!started[i])fprintf(f,"%s\n",hd[defs[i].type]);
      fputs(buf[i].c_str(),f); fclose(f);
      started[i]=true; buf[i].clear();
    }
  }
};

// ------------------------------------------------------------------------------------------------
// x-slab runs (-mgpu:N): one host thread per GPU of this process, each with its own Engine; the exchange and the dt all-reduce go
// through the library's NCCL transport (mgpu_comm.cu), so the threads run in lockstep without any host-side message passing.
// ------------------------------------------------------------------------------------------------
struct SlabGroup {
  int nranks=1;
  unsigned char uid[128]={0};
  std::vector<unsigned> cuts;
  std::mutex mx; std::condition_variable cv; int arrived=0; unsigned long long generation=0;
  std::vector<Engine*> engines; std::atomic<bool> failed{false}; std::string error;
  std::vector<unsigned long long> nout, partsteps; std::vector<unsigned> np0;
  explicit SlabGroup(int n):nranks(n),engines(n,nullptr),nout(n,0),partsteps(n,0),np0(n,0){}
  void Barrier(){
    std::unique_lock<std::mutex> lk(mx);
    const unsigned long long g=generation;
    if(++arrived==nranks){ arrived=0; generation++; cv.notify_all(); }
    else cv.wait(lk,[&]{ return generation!=g || failed.load(); });
    if(failed.load())throw std::runtime_error("another rank of the slab run failed");
  }
  void Fail(int rank,const std::string &msg){
    { std::lock_guard<std::mutex> lk(mx); if(error.empty()){ char tx[32]; snprintf(tx,sizeof(tx),"[rank %d] ",rank); error=tx+msg; } }
    failed.store(true);
    for(int r=0;r<nranks;r++)if(r!=rank && engines[r])engines[r]->MgCommAbort();
    cv.notify_all();
  }
};

static RunSummary RunRank(const RunOptions &opt,CaseData &cd,int rank,SlabGroup *grp,double timestep,unsigned part,double restartdtpre){
  RunSummary sum;
  const int nranks=(grp? grp->nranks: 1);
  const bool slabs=(nranks>1);
  Log log; log.echo=(opt.verbose && rank==0);
  { char nm[32]; if(rank)snprintf(nm,sizeof(nm),"/Run_p%02d.out",rank); else snprintf(nm,sizeof(nm),"/Run.out"); log.Open(opt.dirout+nm); }
  dsph_config &c=cd.cfg;
  const unsigned npcase=cd.np();
  // ---- particles of this rank
  std::vector<unsigned> sel;                 // indices into the case arrays (slab runs)
  unsigned np0=npcase;
  if(slabs){
    SphConst k; DeriveConstants(c,k);
    const double sc=double(k.scell);
    for(unsigned i=0;i<npcase;i++){
      const unsigned cx=unsigned((cd.pos[3*size_t(i)]-k.domposmin[0])/sc);       // the arithmetic of the cell hash (celldiv.cuh)
      if(cx>=grp->cuts[rank] && cx<grp->cuts[rank+1])sel.push_back(i);
    }
    np0=unsigned(sel.size());
  }
  if(grp)grp->np0[rank]=np0;
  unsigned capacity=std::max(np0,slabs? 0u: c.casenp)+np0/20+1024;
  if(c.perix||c.periy||c.periz)capacity=unsigned(np0*1.7)+65536;
  if(slabs)capacity=unsigned(np0*((c.perix||c.periy||c.periz)? 2.0: 1.5))+200000;   // room for halo copies, periodic copies and for the re-balancing
  Engine eng(opt.device+rank,c,capacity);
  if(grp)grp->engines[rank]=&eng;
  if(opt.forcevariant)eng.forcevariant=opt.forcevariant;
  eng.keepout=opt.savepart;
  log.Printf("dsph_b200 (B200-native DualSPHysics hot path)  case=%s/%s%s",opt.dir.c_str(),opt.casename.c_str(),slabs? "  [x-slab run]": "");
  if(timestep>0)log.Printf("Restart from Part_%04u: TimeStep=%.9g",part,timestep);
  log.Printf("Particles: np=%u nbound=%u nfluid=%u  dp=%g h=%g cs0=%g  cells(map)=%ux%ux%u",npcase,c.casenfixed,c.casenfluid,c.dp,double(c.kernelh),eng.k.cs0,eng.k.mapcells[0],eng.k.mapcells[1],eng.k.mapcells[2]);
  if(!eng.ringbase_ok)log.Printf("*** WARNING: the shared-memory window of this device/driver does not start where the ring force kernels expect it (probe at context creation): "
                                 "the plain kernels (variant 4) run instead, about 1.5x slower in Interaction_Forces.");
  {
    std::vector<typecode> code(np0); std::vector<unsigned> idp; std::vector<double> pos; std::vector<float> vr;
    const unsigned *pidp=cd.idp.data(); const double *ppos=cd.pos.data(); const float *pvr=cd.velrhop.data();
    if(slabs){
      idp.resize(np0); pos.resize(size_t(np0)*3); vr.resize(size_t(np0)*4);
      for(unsigned q=0;q<np0;q++){ const size_t i=sel[q]; idp[q]=cd.idp[i]; memcpy(&pos[3*size_t(q)],&cd.pos[3*i],24); memcpy(&vr[4*size_t(q)],&cd.velrhop[4*i],16); }
      pidp=idp.data(); ppos=pos.data(); pvr=vr.data();
    }
    ComputeCodes(cd,np0,pidp,code.data());
    unsigned npb0=0; for(unsigned i=0;i<np0;i++)if(pidp[i]<c.casenfixed+c.casenmoving)npb0++;     // boundary ids come first (CheckSortParticles)
    eng.Upload(np0,npb0,pidp,code.data(),ppos,pvr);
    if(eng.mdbc){
      // JSph::LoadBoundNormals + ConfigBoundNormals (S/JSph.cpp:1262-1340): vector particle -> boundary limit from the normals file,
      // doubled = particle -> ghost node
      if(opt.partbegin>=0)throw std::runtime_error("Restart with mDBC needs the extra data of the PART (not written by this driver).");
      if(slabs)throw std::runtime_error("mDBC cannot be combined with -mgpu yet (the halo records carry no normals).");
      std::vector<float> nrm(size_t(npb0)*3,0.f);
      for(unsigned q=0;q<npb0;q++){ const size_t id=pidp[q]; if(3*id+2<cd.normals.size())for(int i=0;i<3;i++)nrm[3*size_t(q)+i]=float(cd.normals[3*id+i])*2.f; }
      eng.SetBoundNormals(nrm.data(),npb0);
      log.Printf("Boundary: mDBC (slip mode velocity=0%s), %u normal vectors",eng.mdbccorrector? ", corrector": "",npb0);
    }
  }
  if(restartdtpre>0)eng.symplecticdtpre=restartdtpre;
  dsph_divinfo di;
  if(slabs){
    eng.MgCommInit(grp->uid,rank,nranks);
    eng.MgConfigCuts(rank,nranks,grp->cuts.data());
    log.Printf("Slab %d of %d: cell columns [%u,%u)  np=%u  (NCCL %d)",rank,nranks,grp->cuts[rank],grp->cuts[rank+1],np0,MgNcclVersion());
    eng.MgCellDivide(&di);
  }
  else eng.CellDivide(true,&di);
  // ---- moving boundaries: every rank evaluates the motion laws of Case.xml itself (its own copy of the stateful objects) and moves the
  //      moving particles it owns; a restart replays the laws up to the restart time (JMotion::ResetTime, S/JMotion.cpp:432-440)
  MotionSet motion;
  eng.gaugetime=timestep;
  if(cd.cfg.casenmoving){
    motion=cd.motion; motion.Reset();
    if(timestep>0){ std::vector<dsph_motiondata> skip; motion.Process(0,timestep,skip); }
    eng.motionset=&motion;
    if(!rank)log.Printf("Motion: %u moving object(s), %u particles (<motion> of Case.xml)",unsigned(motion.objs.size()),cd.cfg.casenmoving);
  }
  // ---- gauges: first evaluation on the initial state (RunGaugeSystem(TimeStep,true), S/JSphCpuSingle.cpp:1070)
  GaugeOutput gout;
  if(!cd.gauges.empty()){
    if(slabs && (c.perix||c.periy||c.periz))log.Printf("*** WARNING: gauges are not evaluated in x-slab runs of periodic cases.");
    else{
      eng.GaugesConfig(unsigned(cd.gauges.size()),cd.gauges.data());
      if(!rank)gout.Config(cd.gauges,opt.dirout);                 // slab runs: every rank evaluates (collective), rank 0 writes the files
      eng.gaugetime=timestep;
      eng.GaugesRun(timestep,false,nullptr); if(!rank)gout.Collect(eng.gaugeres);
      log.Printf("Gauges: %u (<special><gauges>) -> Gauges<Type>_<name>.csv",unsigned(cd.gauges.size()));
    }
  }
  // ---- output set
  PartSetHead head;
  MakePartSetHead(head,cd,eng.k.maprealposmin,eng.k.maprealposmax,eng.k.periactive,eng.k.perixinc,eng.k.periyinc,eng.k.perizinc);
  head.firstpart=(opt.partbegin>=0? part: 0);
  if(slabs){ char tx[16]; unsigned v=0; for(int i=0;i<128;i++)v=v*131u+grp->uid[i]; snprintf(tx,sizeof(tx),"%08x",v); head.runcode=tx; }   // the same run code on every piece
  PartSet pset(opt.dirout,head,unsigned(rank),unsigned(nranks),slabs? 1: 0,opt.piecemax? opt.piecemax: PartSet::PIECE_MAX_DEFAULT,opt.partbegin>=0);
  if(opt.savepart && rank==0)pset.SaveHead();
  if(pset.Npiece()>1 && !slabs)log.Printf("PART files are written in %u pieces (4 GB limit of the bi4 container).",pset.Npiece());
  std::unique_ptr<PartWriter> writer;
  if(opt.savepart)writer.reset(new PartWriter(eng,pset,capacity));
  unsigned long long nstep=0;
  const auto t0=std::chrono::steady_clock::now();
  auto Elapsed=[&]{ return std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); };
  std::vector<double> probes;
  if(!slabs)probes=LoadPoints(opt.dir+"/"+opt.casename+"_Points.txt");
  if(!probes.empty())log.Printf("Probe points: %u (%s_Points.txt) -> MeasurePoints.csv",unsigned(probes.size()/3),opt.casename.c_str());
  PartStats pst; pst.timem1=timestep;
  if(opt.partbegin<0){
    if(opt.savepart)SavePart(eng,cd,part,timestep,0,0,0,pst,log,*writer);
    SaveProbes(eng,probes,opt.dirout,part,timestep);
    gout.Save();
  }
  part++;
  double timepartnext=NextTime(timestep,c.timeout);
  eng.GetTimers(nullptr,true);
  const auto tsim0=std::chrono::steady_clock::now();
  unsigned long long partsteps=0;
  const int rebalance_every=(opt.rebalancesteps>0? opt.rebalancesteps: 50);
  unsigned nrebalanced=0;
  while(timestep<c.timemax){
    double dt; dsph_forceres fr;
    if(slabs)eng.MgStep(&dt,&fr); else eng.Step(&dt,&fr);
    partsteps+=eng.np-eng.npbper-eng.npfper;
    timestep+=dt; nstep++;
    if(!gout.defs.empty())gout.Collect(eng.gaugeres);
    pst.dtmin=std::min(pst.dtmin,dt); pst.dtmax=std::max(pst.dtmax,dt);
    if(slabs && nstep%rebalance_every==0){
      // collective: the same decision on every rank (the histogram and the costs are all-reduced)
      if(eng.MgRebalance(opt.boundweight>0? opt.boundweight: 0.05,eng.MgMeasuredCost(),nullptr)){ eng.MgCellDivide(nullptr); nrebalanced++; }
    }
    unsigned long long nout=eng.npout_total;
    const bool dopart=(timestep>=timepartnext);          // dt is the same on every rank: the same decision everywhere
    bool partoutstop=false;
    if(!slabs)partoutstop=(c.partsoutmax_on && double(nout)>double(c.partsoutmax)*double(c.casenfluid)) || !eng.np;
    else if(dopart){                                      // slab runs check the limit of excluded particles at PART times, on the global count
      grp->nout[rank]=nout; grp->Barrier();
      unsigned long long tot=0; for(int r=0;r<nranks;r++)tot+=grp->nout[r];
      grp->Barrier();
      partoutstop=(c.partsoutmax_on && double(tot)>double(c.partsoutmax)*double(c.casenfluid));
    }
    if(dopart || partoutstop){
      if(opt.savepart)SavePart(eng,cd,part,timestep,nstep,Elapsed(),nout,pst,log,*writer);
      else log.Printf("Part_%04u  %12.6f  %12llu  dt=%.6g  np=%u",part,timestep,nstep,dt,eng.np);
      SaveProbes(eng,probes,opt.dirout,part,timestep);
      gout.Save();
      part++;
      timepartnext=NextTime(timestep,c.timeout);
      if(partoutstop){ log.Printf("Particles OUT limit reached..."); break; }
    }
    if(opt.nsteps>0 && nstep>=(unsigned long long)opt.nsteps)break;
  }
  eng.FlushStepTimer();
  gout.Save();
  const double tsim=std::chrono::duration<double>(std::chrono::steady_clock::now()-tsim0).count();
  double twrite=0;
  if(writer){ const auto tw0=std::chrono::steady_clock::now(); writer->Finish(); twrite=std::chrono::duration<double>(std::chrono::steady_clock::now()-tw0).count(); }   // the last PART must be complete before the run reports success
  dsph_timers tm; eng.GetTimers(&tm,false);
  sum.np0=np0; sum.steps=nstep; sum.simtime=timestep; sum.runtime_s=tsim; sum.parts=part; sum.npout=eng.npout_total;
  sum.mpartsteps_s=(tsim>0? double(partsteps)/tsim/1e6: 0);
  if(grp){
    grp->partsteps[rank]=partsteps; grp->nout[rank]=eng.npout_total; grp->Barrier();
    unsigned long long ps=0,no=0; unsigned n0=0; for(int r=0;r<nranks;r++){ ps+=grp->partsteps[r]; no+=grp->nout[r]; n0+=grp->np0[r]; }
    sum.np0=n0; sum.npout=no; sum.mpartsteps_s=(tsim>0? double(ps)/tsim/1e6: 0);
    grp->Barrier();
  }
  log.Printf("[Simulation finished]  steps=%llu  TimeStep=%.9g  runtime=%.4f s  %.3f Mparticle-steps/s  excluded=%llu",nstep,timestep,tsim,sum.mpartsteps_s,sum.npout);
  if(writer)log.Printf("PART pipeline: the time loop waited %.4f s for a free host buffer; %.4f s to finish the last files after the loop",writer->waited_s,twrite);
  if(slabs)log.Printf("Slab re-balancing: every %d steps, cuts moved %u times",rebalance_every,nrebalanced);
  log.Printf("[GPU Timers] (device ms)");
  log.Printf("NL-Limits........: %.3f",tm.nl_limits); log.Printf("NL-PreSort.......: %.3f",tm.nl_presort);
  log.Printf("NL-RadixSort.....: %.3f",tm.nl_radixsort); log.Printf("NL-CellBegin.....: %.3f",tm.nl_cellbegin);
  log.Printf("NL-SortData......: %.3f",tm.nl_sortdata); log.Printf("NL-OutCheck......: %.3f",tm.nl_outcheck);
  log.Printf("CF-PreForces.....: %.3f",tm.cf_preforces); log.Printf("CF-Forces........: %.3f",tm.cf_forces);
  log.Printf("SU-ComputeStep...: %.3f",tm.su_computestep); log.Printf("SU-Periodic......: %.3f",tm.su_periodic);
  log.Printf("Kernel launches..: %llu",(unsigned long long)tm.kernel_launches);
  { // Run.csv one-line summary (S/JSph.cpp:2962-3006)
    char nm[32]; if(rank)snprintf(nm,sizeof(nm),"/Run_p%02d.csv",rank); else snprintf(nm,sizeof(nm),"/Run.csv");
    FILE *f=fopen((opt.dirout+nm).c_str(),"w");
    if(f){
      fprintf(f,"#RunName;Np;Steps;TSimul;TimeStep;MPartStepsPerSec;NL-Limits;NL-PreSort;NL-RadixSort;NL-CellBegin;NL-SortData;CF-Forces;SU-ComputeStep\n");
      fprintf(f,"%s;%u;%llu;%.6f;%.9g;%.4f;%.3f;%.3f;%.3f;%.3f;%.3f;%.3f;%.3f\n",opt.casename.c_str(),sum.np0,nstep,tsim,timestep,sum.mpartsteps_s,tm.nl_limits,tm.nl_presort,tm.nl_radixsort,tm.nl_cellbegin,tm.nl_sortdata,tm.cf_forces,tm.su_computestep);
      fclose(f);
    }
  }
  writer.reset();
  if(grp)grp->engines[rank]=nullptr;
  return sum;
}

RunSummary RunCase(const RunOptions &opt){
  mkdir(opt.dirout.c_str(),0755);
  CaseData cd; LoadCase(cd,opt.dir,opt.casename);
  dsph_config &c=cd.cfg;
  if(opt.timemax>0)c.timemax=opt.timemax;
  if(opt.timeout>0)c.timeout=opt.timeout;
  double timestep=0; unsigned part=0; double restartdtpre=0;
  if(opt.partbegin>=0){
    // restart: particle state, time and the Symplectic dt of the next step come from the PART file(s) of a previous run
    // (JPartsLoad4::LoadParticles with partbegin + JSph::RestartCheckData, S/JSph.cpp:2005-2040); excluded particles stay excluded
    PartHead ph; PartData pd;
    const std::string dirb=(opt.partbegindir.empty()? opt.dirout: opt.partbegindir);
    LoadPartPieces(dirb,unsigned(opt.partbegin),ph,pd);
    if(ph.casenp!=c.casenp || ph.casenfixed!=c.casenfixed || ph.casenfluid!=c.casenfluid)throw std::runtime_error("The PART file of -partbegin does not belong to this case.");
    cd.idp=pd.idp; cd.pos=pd.pos; cd.velrhop.resize(size_t(pd.npok)*4);
    for(unsigned i=0;i<pd.npok;i++){ cd.velrhop[4*size_t(i)]=pd.vel[3*size_t(i)]; cd.velrhop[4*size_t(i)+1]=pd.vel[3*size_t(i)+1]; cd.velrhop[4*size_t(i)+2]=pd.vel[3*size_t(i)+2]; cd.velrhop[4*size_t(i)+3]=pd.rhop[i]; }
    timestep=pd.timestep; part=unsigned(opt.partbegin); restartdtpre=pd.symplecticdtpre;
  }
  const unsigned np0=cd.np();
  // JSph::CheckRhopLimits (S/JSph.cpp:2021-2030): the initial density of every fluid particle must be inside RhopOut
  if(c.rhopoutmin<c.rhopoutmax)for(unsigned i=0;i<np0;i++)if(cd.idp[i]>=c.casenfixed+c.casenmoving+c.casenfloat){
    const float r=cd.velrhop[4*size_t(i)+3];
    if(r<c.rhopoutmin || c.rhopoutmax<r)throw std::runtime_error("Initial fluid density is out of limits. *** To change the limits modify the value of 'RhopOutMin' and 'RhopOutMax' in the XML file (parameters section).");
  }
  if(opt.mgpu<=1)return RunRank(opt,cd,0,nullptr,timestep,part,restartdtpre);
  // ---- slab run: cuts from the per-column histogram of the work (fluid 1, boundary weight), one thread per GPU
  SlabGroup grp(opt.mgpu);
  {
    SphConst k; DeriveConstants(c,k);
    const double sc=double(k.scell), bw=(opt.boundweight>0? opt.boundweight: 0.05);
    std::vector<double> col(k.domcells[0],0.0);
    const unsigned nbound=c.casenfixed+c.casenmoving+c.casenfloat;
    for(unsigned i=0;i<np0;i++){
      const unsigned cx=unsigned((cd.pos[3*size_t(i)]-k.domposmin[0])/sc);
      if(cx<col.size())col[cx]+=(cd.idp[i]<nbound? bw: 1.0);
    }
    const unsigned halo=unsigned(k.scelldiv)*(c.nn_on && c.nn_velgrad==DSPH_VELGRAD_SPH && c.tvisco!=DSPH_VISCO_ARTIFICIAL? 2u: 1u);
    const unsigned minwidth=2*halo+1;
    if(unsigned(opt.mgpu)*minwidth>col.size())throw std::runtime_error("Too many GPUs for the cell columns of the domain.");
    grp.cuts=MgBalancedCuts(col,opt.mgpu,minwidth);
  }
  MgUniqueId(grp.uid);
  std::vector<RunSummary> sums(opt.mgpu);
  std::vector<std::thread> ths;
  for(int r=0;r<opt.mgpu;r++)ths.emplace_back([&,r]{
    try{ sums[r]=RunRank(opt,cd,r,&grp,timestep,part,restartdtpre); }
    catch(const std::exception &e){ grp.Fail(r,e.what()); }
  });
  for(auto &t:ths)t.join();
  if(grp.failed.load())throw std::runtime_error(grp.error.empty()? std::string("slab run failed"): grp.error);
  return sums[0];
}

} // namespace dsph
