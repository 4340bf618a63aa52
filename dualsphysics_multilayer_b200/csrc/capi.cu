// capi.cu -- extern "C" surface of include/dsph_b200.h (thin: argument checks + exception -> status code).
#include "../../include/dsph_b200.h"
#include "engine.h"
#include "host/casefile.h"
#include "host/driver.h"
#include "host/bi4.h"
#include "host/partset.h"
#include <vector>
#include <algorithm>
#include <cstring>
#include <string>
#include <new>

using namespace dsph;

struct dsph_ctx { Engine *eng; };

static thread_local std::string g_lasterror;

template<class F> static int Guard(F f){
  try{ f(); return DSPH_OK; }
  catch(const EngineError &e){ g_lasterror=e.what(); return e.code; }
  catch(const std::bad_alloc &){ g_lasterror="Out of host memory."; return DSPH_ERR_STATE; }
  catch(const std::exception &e){ g_lasterror=e.what(); return DSPH_ERR_STATE; }
  catch(...){ g_lasterror="Unknown exception."; return DSPH_ERR_STATE; }
}
#define NEEDCTX(c) if(!(c)||!(c)->eng){ g_lasterror="Invalid context."; return DSPH_ERR_ARG; }

extern "C" {

const char* dsph_version(void){ return "dsph_b200 0.1 (sm_100a)"; }
const char* dsph_last_error(void){ return g_lasterror.c_str(); }
int dsph_device_count(void){ int n=0; if(cudaGetDeviceCount(&n)!=cudaSuccess){ cudaGetLastError(); return 0; } return n; }

int dsph_ctx_create(dsph_ctx **out,int device,const dsph_config *cfg,uint32_t np_capacity){
  if(!out||!cfg){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  *out=nullptr;
  return Guard([&]{
    unsigned capacity=np_capacity;
    if(!capacity){
      capacity=cfg->casenp+cfg->casenp/20+1024;                    // reference: Np*(1+over)+128 (S/JSphGpu.cpp:359-429)
      if(cfg->perix||cfg->periy||cfg->periz)capacity=unsigned(cfg->casenp*1.7)+65536;
    }
    Engine *e=new Engine(device,*cfg,capacity);
    dsph_ctx *c=new dsph_ctx; c->eng=e; *out=c;
  });
}
int dsph_ctx_destroy(dsph_ctx *ctx){
  if(!ctx)return DSPH_OK;
  return Guard([&]{ delete ctx->eng; ctx->eng=nullptr; delete ctx; });
}
int dsph_constants_upload(dsph_ctx *ctx,const dsph_config *cfg){
  NEEDCTX(ctx); if(!cfg){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ ctx->eng->Reconfigure(*cfg); });
}
int dsph_particles_upload(dsph_ctx *ctx,uint32_t np,uint32_t npb,const uint32_t *idp,const dsph_typecode *code,const double *pos,const float *velrhop){
  NEEDCTX(ctx); if(!idp||!pos||!velrhop){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ ctx->eng->Upload(np,npb,idp,code,pos,velrhop); });
}
int dsph_particles_set_velrhopm1(dsph_ctx *ctx,const float *v){
  NEEDCTX(ctx); if(!v){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ ctx->eng->SetVelrhopM1(v); });
}
int dsph_particles_set_spstau(dsph_ctx *ctx,const float *t){
  NEEDCTX(ctx); if(!t){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ ctx->eng->SetSpsTau(t); });
}
int dsph_particles_download(dsph_ctx *ctx,uint32_t first,uint32_t count,uint32_t *idp,dsph_typecode *code,uint32_t *dcell,double *pos,float *velrhop){
  NEEDCTX(ctx);
  return Guard([&]{ ctx->eng->Download(first,count,idp,code,dcell,pos,velrhop); });
}
uint32_t dsph_np(const dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->np: 0; }
uint32_t dsph_npb(const dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->npb: 0; }

int dsph_celldiv_run(dsph_ctx *ctx,int updateperiodic,dsph_divinfo *info){
  NEEDCTX(ctx);
  return Guard([&]{ ctx->eng->CellDivide(updateperiodic!=0,info); ctx->eng->FlushStepTimer(); });
}
int dsph_celldiv_get_begincell(dsph_ctx *ctx,uint32_t *begincell,uint32_t capacity){
  NEEDCTX(ctx); if(!begincell){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ ctx->eng->GetBeginCell(begincell,capacity); });
}
int dsph_forces_run(dsph_ctx *ctx,int interstep,dsph_forceres *res){
  NEEDCTX(ctx);
  return Guard([&]{ ctx->eng->Forces(interstep,res); });
}
int dsph_forces_download(dsph_ctx *ctx,uint32_t first,uint32_t count,float *ace,float *ar){
  NEEDCTX(ctx);
  return Guard([&]{ ctx->eng->ForcesDownload(first,count,ace,ar); });
}
int dsph_step_verlet(dsph_ctx *ctx,double dt){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->StepVerlet(dt); }); }
int dsph_step_sympre(dsph_ctx *ctx,double dt){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->StepSymPre(dt); }); }
int dsph_step_symcor(dsph_ctx *ctx,double dt){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->StepSymCor(dt); }); }
int dsph_step(dsph_ctx *ctx,double *stepdt,dsph_forceres *res){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->Step(stepdt,res); }); }
double dsph_symplectic_dtpre(const dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->symplecticdtpre: 0; }
int dsph_set_symplectic_dtpre(dsph_ctx *ctx,double dt){ NEEDCTX(ctx); ctx->eng->symplecticdtpre=dt; return DSPH_OK; }
int dsph_set_verletstep(dsph_ctx *ctx,int v){ NEEDCTX(ctx); ctx->eng->verletstep=v; return DSPH_OK; }
int dsph_set_force_variant(dsph_ctx *ctx,int v){ NEEDCTX(ctx); ctx->eng->forcevariant=v; return DSPH_OK; }
int dsph_set_refbinary_tail(dsph_ctx *ctx,int on){ NEEDCTX(ctx); ctx->eng->reftail=(on!=0); ctx->eng->dc.reftail=(on? 1: 0); ctx->eng->divok=false; return DSPH_OK; }
int dsph_get_force_variant(const dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->EffectiveForceVariant(): 0; }
int dsph_get_timers(dsph_ctx *ctx,dsph_timers *t,int reset){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->GetTimers(t,reset!=0); }); }
double dsph_last_force_ms(const dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->lastforcems: 0; }
int dsph_pips(dsph_ctx *ctx,uint64_t *fr,uint64_t *fc,uint64_t *br,uint64_t *bc){
  NEEDCTX(ctx);
  return Guard([&]{ unsigned long long a,b,c,d; ctx->eng->Pips(&a,&b,&c,&d); if(fr)*fr=a; if(fc)*fc=b; if(br)*br=c; if(bc)*bc=d; });
}
int dsph_measure_points(dsph_ctx *ctx,uint32_t npoints,const double *points_xyz,float *press,float *vel_xyz,float *kcorr){
  NEEDCTX(ctx); if(npoints && !points_xyz){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ ctx->eng->MeasurePoints(npoints,points_xyz,press,vel_xyz,kcorr); });
}
// CUDA-event stopwatch on the context's stream (bench.py: torch events do not see this stream)
int dsph_timer_start(dsph_ctx *ctx){ NEEDCTX(ctx); return Guard([&]{ if(cudaEventRecord(ctx->eng->ev[14],ctx->eng->stm)!=cudaSuccess)throw EngineError(DSPH_ERR_CUDA,"cudaEventRecord"); }); }
int dsph_timer_stop(dsph_ctx *ctx,double *ms){
  NEEDCTX(ctx);
  return Guard([&]{
    Engine *e=ctx->eng; float f=0;
    if(cudaEventRecord(e->ev[15],e->stm)!=cudaSuccess || cudaEventSynchronize(e->ev[15])!=cudaSuccess || cudaEventElapsedTime(&f,e->ev[14],e->ev[15])!=cudaSuccess)throw EngineError(DSPH_ERR_CUDA,"cuda event timing failed");
    if(ms)*ms=f;
  });
}
// accumulated device time and launch count of the fluid force kernel (the dominant kernel) since the last reset
int dsph_fluid_kernel_stats(dsph_ctx *ctx,double *ms_sum,uint64_t *nlaunch,int reset){
  NEEDCTX(ctx); Engine *e=ctx->eng;
  if(ms_sum)*ms_sum=e->fluidms_sum; if(nlaunch)*nlaunch=e->fluid_launches;
  if(reset){ e->fluidms_sum=0; e->fluid_launches=0; }
  return DSPH_OK;
}
void* dsph_cuda_stream(dsph_ctx *ctx){ return (ctx&&ctx->eng)? (void*)ctx->eng->stm: nullptr; }

// ---- x-slab multi-GPU ----
int dsph_mgpu_config(dsph_ctx *ctx,int rank,int nranks,uint32_t xlo,uint32_t xhi){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgConfig(rank,nranks,xlo,xhi); }); }
uint32_t dsph_mgpu_record_bytes(const dsph_ctx *ctx){ return (ctx && ctx->eng)? ctx->eng->MgRecordBytes(): 88u; }
int dsph_mgpu_count(dsph_ctx *ctx,uint32_t counts[8]){ NEEDCTX(ctx); return Guard([&]{ unsigned c[8]; ctx->eng->MgCount(c); for(int i=0;i<8;i++)counts[i]=c[i]; }); }
int dsph_mgpu_column_counts(dsph_ctx *ctx,uint32_t *fluid,uint32_t *bound,uint32_t ncol){ NEEDCTX(ctx); if(!fluid||!bound){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ ctx->eng->MgColumnCounts(fluid,bound,ncol); }); }
int dsph_mgpu_pack(dsph_ctx *ctx,void *sl,void *sr){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgPack(sl,sr); }); }
int dsph_mgpu_unpack(dsph_ctx *ctx,const uint32_t counts[8],const void *rl,const void *rr){ NEEDCTX(ctx); return Guard([&]{ unsigned c[8]; for(int i=0;i<8;i++)c[i]=counts[i]; ctx->eng->MgUnpack(c,rl,rr); }); }
// native NCCL transport (mgpu_comm.cu)
int dsph_mgpu_unique_id(void *id128){ if(!id128){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ dsph::MgUniqueId(id128); }); }
int dsph_mgpu_nccl_version(void){ return dsph::MgNcclVersion(); }
int dsph_mgpu_comm_init(dsph_ctx *ctx,const void *id128,int rank,int nranks){ NEEDCTX(ctx); if(!id128){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ ctx->eng->MgCommInit(id128,rank,nranks); }); }
int dsph_mgpu_comm_attach(dsph_ctx *ctx,void *ncclcomm,int rank,int nranks){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgCommAttach(ncclcomm,rank,nranks); }); }
int dsph_mgpu_config_cuts(dsph_ctx *ctx,int rank,int nranks,const uint32_t *cuts){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgConfigCuts(rank,nranks,cuts); }); }
uint32_t dsph_mgpu_halo_cells(const dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->MgHalo(): 0u; }
int dsph_mgpu_exchange(dsph_ctx *ctx){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgExchange(); }); }
int dsph_motion_run(dsph_ctx *ctx,const dsph_motiondata *objs,uint32_t nobj,double stepdt){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MotionRun(objs,nobj,stepdt); }); }
struct dsph_motionset { MotionSet ms; };
int dsph_case_motion_open(const char *dir,const char *casename,dsph_motionset **ms,uint32_t *nobj){
  if(!dir||!casename||!ms){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  *ms=nullptr;
  const int r=Guard([&]{ dsph_motionset *m=new dsph_motionset; try{ LoadCaseMotion(m->ms,dir,casename); }catch(...){ delete m; throw; } m->ms.Reset(); *ms=m; if(nobj)*nobj=uint32_t(m->ms.objs.size()); });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_case_motion_process(dsph_motionset *ms,double timestep,double dt,dsph_motiondata *out,int *active){
  if(!ms||(!out && !ms->ms.objs.empty())){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ std::vector<dsph_motiondata> v; const bool a=ms->ms.Process(timestep,dt,v); for(size_t i=0;i<v.size();i++)out[i]=v[i]; if(active)*active=(a? 1: 0); });
}
int dsph_case_motion_close(dsph_motionset *ms){ delete ms; return DSPH_OK; }
int dsph_motion_attach(dsph_ctx *ctx,dsph_motionset *ms){ NEEDCTX(ctx); ctx->eng->motionset=(ms? &ms->ms: nullptr); return DSPH_OK; }
int dsph_particles_set_boundnormals(dsph_ctx *ctx,const float *n,uint32_t count){ NEEDCTX(ctx); if(!n && count){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ ctx->eng->SetBoundNormals(n,count); }); }
int dsph_particles_get_boundnormals(dsph_ctx *ctx,float *n,uint32_t count){ NEEDCTX(ctx); if(!n && count){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ ctx->eng->GetBoundNormals(n,count); }); }
int dsph_set_mdbc_threshold(dsph_ctx *ctx,float v){ NEEDCTX(ctx); ctx->eng->mdbcthreshold=v; return DSPH_OK; }
int dsph_case_normals(const char *dir,const char *casename,double *out,uint32_t max,uint32_t *n){
  if(!dir||!casename||!n){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{
    std::vector<double> v; LoadCaseNormals(v,dir,casename);
    *n=uint32_t(v.size()/3);
    if(out)memcpy(out,v.data(),std::min<size_t>(v.size(),size_t(max)*3)*8);
  });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_excluded_keep(dsph_ctx *ctx,int keep){ NEEDCTX(ctx); ctx->eng->keepout=(keep!=0); return DSPH_OK; }
int dsph_excluded_take(dsph_ctx *ctx,uint32_t max,uint32_t *n,uint32_t *idp,dsph_typecode *code,double *pos,float *velrhop){
  NEEDCTX(ctx);
  return Guard([&]{
    Engine *e=ctx->eng;
    const size_t have=e->outidp.size(), m=std::min<size_t>(have,max);
    if(n)*n=uint32_t(have);
    if(idp && m)memcpy(idp,e->outidp.data(),m*4);
    if(code && m)memcpy(code,e->outcode.data(),m*sizeof(typecode));
    if(pos && m)memcpy(pos,e->outpos.data(),m*24);
    if(velrhop && m)memcpy(velrhop,e->outvelrhop.data(),m*16);
    e->outidp.clear(); e->outcode.clear(); e->outpos.clear(); e->outvelrhop.clear();
  });
}
int dsph_gauges_config(dsph_ctx *ctx,uint32_t n,const dsph_gauge *g){ NEEDCTX(ctx); if(n && !g){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ ctx->eng->GaugesConfig(n,g); }); }
int dsph_gauges_run(dsph_ctx *ctx,double timestep,int force_all,dsph_gaugeres *res){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->GaugesRun(timestep,force_all!=0,res); }); }
int dsph_gauges_results(dsph_ctx *ctx,dsph_gaugeres *res){ NEEDCTX(ctx); if(!res){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ if(!ctx->eng->gaugeres.empty())memcpy(res,ctx->eng->gaugeres.data(),ctx->eng->gaugeres.size()*sizeof(dsph_gaugeres)); }); }
int dsph_gauges_set_time(dsph_ctx *ctx,double timestep){ NEEDCTX(ctx); ctx->eng->gaugetime=timestep; return DSPH_OK; }
int dsph_case_gauges(const char *dir,const char *casename,dsph_gauge *out,uint32_t max,uint32_t *n){
  if(!dir||!casename||!n){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{
    CaseData cd; LoadCase(cd,dir,casename);
    *n=uint32_t(cd.gauges.size());
    for(uint32_t i=0;i<*n && i<max && out;i++)out[i]=cd.gauges[i];
  });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_celldiv_last_info(dsph_ctx *ctx,dsph_divinfo *info){ NEEDCTX(ctx); if(!info){ g_lasterror="Null argument."; return DSPH_ERR_ARG; } return Guard([&]{ ctx->eng->LastDivInfo(info); }); }
int dsph_mgpu_sent_records(dsph_ctx *ctx,uint64_t *sent){ NEEDCTX(ctx); if(sent)*sent=ctx->eng->mgsentrecords; return DSPH_OK; }
int dsph_mgpu_transport(dsph_ctx *ctx,int *direct,uint64_t *fallbacks){ NEEDCTX(ctx); if(direct)*direct=ctx->eng->mgdirect; if(fallbacks)*fallbacks=ctx->eng->mgdfallbacks; return DSPH_OK; }
int dsph_mgpu_celldiv(dsph_ctx *ctx,dsph_divinfo *info){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgCellDivide(info); ctx->eng->FlushStepTimer(); }); }
int dsph_mgpu_forces(dsph_ctx *ctx,int interstep,dsph_forceres *res){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgForces(interstep,res); }); }
int dsph_mgpu_step(dsph_ctx *ctx,double *stepdt,dsph_forceres *res){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->MgStep(stepdt,res); }); }
int dsph_mgpu_rebalance(dsph_ctx *ctx,double bound_weight,double rank_cost,int *changed,uint32_t *cuts_out){
  NEEDCTX(ctx);
  return Guard([&]{ const bool ch=ctx->eng->MgRebalance(bound_weight,rank_cost,cuts_out); if(changed)*changed=(ch? 1: 0); });
}
int dsph_mgpu_balanced_cuts(const double *column_work,uint32_t ncol,int nranks,uint32_t minwidth,uint32_t *cuts_out){
  if(!column_work || !cuts_out || nranks<1 || !minwidth){ g_lasterror="Invalid argument."; return DSPH_ERR_ARG; }
  if((unsigned long long)nranks*minwidth>ncol){ g_lasterror="Too many ranks for the cell columns (minimum slab width)."; return DSPH_ERR_ARG; }
  return Guard([&]{ const std::vector<unsigned> c=MgBalancedCuts(std::vector<double>(column_work,column_work+ncol),nranks,minwidth); for(int r=0;r<=nranks;r++)cuts_out[r]=c[r]; });
}
double dsph_mgpu_measured_cost(dsph_ctx *ctx){ return (ctx&&ctx->eng)? ctx->eng->MgMeasuredCost(): 0; }
int dsph_mgpu_profile(dsph_ctx *ctx,int on,double phases_ms[6],uint64_t *sent_records){
  NEEDCTX(ctx); Engine *e=ctx->eng;
  if(phases_ms)for(int i=0;i<6;i++)phases_ms[i]=e->mgphase_ms[i];
  if(sent_records)*sent_records=e->mgsentrecords;
  for(int i=0;i<6;i++)e->mgphase_ms[i]=0;
  e->mgprofile=(on!=0);
  return DSPH_OK;
}
int dsph_ringbase_ok(const dsph_ctx *ctx){ return (ctx&&ctx->eng && ctx->eng->ringbase_ok)? 1: 0; }
int dsph_forces_launch(dsph_ctx *ctx,int interstep){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->ForcesLaunch(interstep); }); }
void* dsph_forces_maxima_ptr(dsph_ctx *ctx){ return (ctx&&ctx->eng)? (void*)ctx->eng->fmax: nullptr; }
int dsph_forces_finish(dsph_ctx *ctx,dsph_forceres *res){ NEEDCTX(ctx); return Guard([&]{ ctx->eng->ForcesFinish(res); }); }
int dsph_domain_info(const dsph_config *cfg,double *scell,double domposmin[3],uint32_t domcells[3]){
  if(!cfg){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ SphConst k; DeriveConstants(*cfg,k); if(scell)*scell=k.scell; for(int i=0;i<3;i++){ if(domposmin)domposmin[i]=k.domposmin[i]; if(domcells)domcells[i]=k.domcells[i]; } });
}

// ---- case tooling (host only) ----
static void CaseToC(const CaseData &cd,dsph_case *out){
  memset(out,0,sizeof(*out));
  out->cfg=cd.cfg; out->np=cd.np(); out->npb=cd.npb();
  out->idp=new uint32_t[cd.idp.size()]; out->pos_xyz=new double[cd.pos.size()]; out->velrhop_xyzw=new float[cd.velrhop.size()];
  memcpy(out->idp,cd.idp.data(),cd.idp.size()*4); memcpy(out->pos_xyz,cd.pos.data(),cd.pos.size()*8); memcpy(out->velrhop_xyzw,cd.velrhop.data(),cd.velrhop.size()*4);
  memcpy(out->caseposmin,cd.caseposmin,24); memcpy(out->caseposmax,cd.caseposmax,24);
  strncpy(out->name,cd.name.c_str(),sizeof(out->name)-1);
  for(auto &b:cd.blocks)if(b.type==3 && out->nfluidblocks<DSPH_MAXPHASES)out->fluidblock_count[out->nfluidblocks++]=b.count;
}
int dsph_case_make(dsph_case *out,const char *kind,uint64_t np_target,uint32_t seed){
  if(!out||!kind){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ CaseData cd; MakeCase(cd,kind,np_target,seed); CaseToC(cd,out); });
}
int dsph_case_make_slab(dsph_case *out,const char *kind,uint64_t np_target,uint32_t seed,uint32_t cx_lo,uint32_t cx_hi){
  if(!out||!kind){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{
    CaseData cd; MakeCaseSlab(cd,kind,np_target,seed,cx_lo,cx_hi); CaseToC(cd,out);
    unsigned nb=0; for(unsigned id:cd.idp)if(id<cd.cfg.casenfixed)nb++;
    out->npb=nb;   // boundary particles of the slab (cfg.casenfixed keeps the count of the full case)
  });
}
int dsph_case_free(dsph_case *c){
  if(!c)return DSPH_OK;
  delete[] c->idp; delete[] c->pos_xyz; delete[] c->velrhop_xyzw; c->idp=nullptr; c->pos_xyz=nullptr; c->velrhop_xyzw=nullptr;
  return DSPH_OK;
}
int dsph_case_load(dsph_case *out,const char *dir,const char *casename){
  if(!out||!dir||!casename){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{ CaseData cd; LoadCase(cd,dir,casename); CaseToC(cd,out); });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_case_save(const dsph_case *c,const char *dir,const char *casename){
  if(!c||!dir||!casename){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{
    CaseData cd; cd.cfg=c->cfg; cd.name=casename;
    cd.idp.assign(c->idp,c->idp+c->np); cd.pos.assign(c->pos_xyz,c->pos_xyz+size_t(c->np)*3); cd.velrhop.assign(c->velrhop_xyzw,c->velrhop_xyzw+size_t(c->np)*4);
    memcpy(cd.caseposmin,c->caseposmin,24); memcpy(cd.caseposmax,c->caseposmax,24);
    // mk blocks: one fixed block + one fluid block per phase (single block when not multiphase)
    if(c->cfg.casenfixed)cd.blocks.push_back({0,0,cd.mkboundfirst,0,c->cfg.casenfixed});
    if(c->nfluidblocks>1){ unsigned beg=c->cfg.casenfixed; for(unsigned b=0;b<c->nfluidblocks;b++){ cd.blocks.push_back({3,b,cd.mkfluidfirst+b,beg,c->fluidblock_count[b]}); beg+=c->fluidblock_count[b]; } }
    else cd.blocks.push_back({3,0,cd.mkfluidfirst,c->cfg.casenfixed,c->cfg.casenfluid});
    // keep the simulation domain exactly as configured
    for(int i=0;i<3;i++){ cd.posmin[i].mode=3; cd.posmin[i].value=c->cfg.maprealposmin[i]; cd.posmax[i].mode=3; cd.posmax[i].value=c->cfg.maprealposmax[i]; }
    SaveCase(cd,dir,casename);
  });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_part_load(dsph_case *out,const char *file,double *timestep,uint32_t *step){
  if(!out||!file){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{
    PartHead h; PartData p; LoadPartFile(file,h,p,0);
    CaseData cd; ConfigDefaults(cd.cfg);
    cd.cfg.casenp=unsigned(h.casenp); cd.cfg.casenfixed=unsigned(h.casenfixed); cd.cfg.casenmoving=unsigned(h.casenmoving);
    cd.cfg.casenfloat=unsigned(h.casenfloat); cd.cfg.casenfluid=unsigned(h.casenfluid);
    cd.cfg.dp=h.dp; cd.cfg.kernelh=float(h.h); cd.cfg.cteb=float(h.b); cd.cfg.rhopzero=float(h.rhop0); cd.cfg.gamma=float(h.gamma);
    cd.cfg.massbound=float(h.massbound); cd.cfg.massfluid=float(h.massfluid); cd.cfg.simulate2d=(h.data2d? 1: 0);
    memcpy(cd.caseposmin,h.caseposmin,24); memcpy(cd.caseposmax,h.caseposmax,24);
    memcpy(cd.cfg.maprealposmin,h.mapposmin,24); memcpy(cd.cfg.maprealposmax,h.mapposmax,24);
    cd.name=h.casename;
    cd.idp=p.idp; cd.pos=p.pos; cd.velrhop.resize(size_t(p.npok)*4);
    for(unsigned i=0;i<p.npok;i++){ cd.velrhop[4*i]=p.vel[3*i]; cd.velrhop[4*i+1]=p.vel[3*i+1]; cd.velrhop[4*i+2]=p.vel[3*i+2]; cd.velrhop[4*i+3]=p.rhop[i]; }
    CaseToC(cd,out);
    unsigned nb=0; for(unsigned id:cd.idp)if(id<cd.cfg.casenfixed+cd.cfg.casenmoving)nb++;
    out->npb=nb;
    if(timestep)*timestep=p.timestep;
    if(step)*step=p.step;
  });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
static void CaseFromC(const dsph_case *c,CaseData &cd){
  cd.cfg=c->cfg; cd.name=c->name;
  cd.idp.assign(c->idp,c->idp+c->np); cd.pos.assign(c->pos_xyz,c->pos_xyz+size_t(c->np)*3); cd.velrhop.assign(c->velrhop_xyzw,c->velrhop_xyzw+size_t(c->np)*4);
  memcpy(cd.caseposmin,c->caseposmin,24); memcpy(cd.caseposmax,c->caseposmax,24);
  if(c->cfg.casenfixed)cd.blocks.push_back({0,0,cd.mkboundfirst,0,c->cfg.casenfixed});
  if(c->nfluidblocks>1){ unsigned beg=c->cfg.casenfixed; for(unsigned b=0;b<c->nfluidblocks;b++){ cd.blocks.push_back({3,b,cd.mkfluidfirst+b,beg,c->fluidblock_count[b]}); beg+=c->fluidblock_count[b]; } }
  else cd.blocks.push_back({3,0,cd.mkfluidfirst,c->cfg.casenfixed,c->cfg.casenfluid});
}
int dsph_partset_save(const dsph_case *c,const char *dirout,uint32_t cpart,double timestep,uint32_t step,double symplecticdtpre,uint32_t piecemax,
                      uint32_t nout,const uint32_t *out_idp,const double *out_pos,const float *out_velrhop,const uint8_t *out_motive){
  if(!c||!dirout||(nout && (!out_idp||!out_pos||!out_velrhop||!out_motive))){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{
    CaseData cd; CaseFromC(c,cd);
    SphConst k; DeriveConstants(cd.cfg,k);
    PartSetHead head; MakePartSetHead(head,cd,k.maprealposmin,k.maprealposmax,k.periactive,k.perixinc,k.periyinc,k.perizinc);
    PartSet ps(dirout,head,0,1,0,piecemax? piecemax: PartSet::PIECE_MAX_DEFAULT,true);
    ps.SaveHead();
    PartMeta m; m.cpart=cpart; m.timestep=timestep; m.step=step; m.nout=nout; m.symplecticdtpre=symplecticdtpre;
    const unsigned n=c->np;
    for(int i=0;i<3;i++){ m.domainmin[i]=k.domposmin[i]; m.domainmax[i]=k.domposmax[i]; }
    std::vector<float> vel(size_t(n)*3),rhop(n);
    for(unsigned q=0;q<n;q++){ vel[3*size_t(q)]=c->velrhop_xyzw[4*size_t(q)]; vel[3*size_t(q)+1]=c->velrhop_xyzw[4*size_t(q)+1]; vel[3*size_t(q)+2]=c->velrhop_xyzw[4*size_t(q)+2]; rhop[q]=c->velrhop_xyzw[4*size_t(q)+3]; }
    PartInfoPlus info; info.on=true; info.npf=n-c->npb; info.npbin=c->npb; info.npused=n; info.npalloc=n;
    ps.SavePart(m,n,c->idp,c->pos_xyz,vel.data(),rhop.data(),&info);
    std::vector<float> ovel(size_t(nout)*3),orhop(nout);
    for(unsigned q=0;q<nout;q++){ ovel[3*size_t(q)]=out_velrhop[4*size_t(q)]; ovel[3*size_t(q)+1]=out_velrhop[4*size_t(q)+1]; ovel[3*size_t(q)+2]=out_velrhop[4*size_t(q)+2]; orhop[q]=out_velrhop[4*size_t(q)+3]; }
    const bool ro=(cd.cfg.rhopoutmin<cd.cfg.rhopoutmax);
    ps.SaveOut(cpart,timestep,nout,out_idp,out_pos,ovel.data(),orhop.data(),out_motive,ro? cd.cfg.rhopoutmin: 0.f,ro? cd.cfg.rhopoutmax: 0.f);
  });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_partset_load(dsph_case *out,const char *dir,uint32_t cpart,double *timestep,uint32_t *step,uint32_t *npiece){
  if(!out||!dir){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  const int r=Guard([&]{
    PartHead h; PartData p; LoadPartPieces(dir,cpart,h,p);
    CaseData cd; ConfigDefaults(cd.cfg);
    cd.cfg.casenp=unsigned(h.casenp); cd.cfg.casenfixed=unsigned(h.casenfixed); cd.cfg.casenmoving=unsigned(h.casenmoving);
    cd.cfg.casenfloat=unsigned(h.casenfloat); cd.cfg.casenfluid=unsigned(h.casenfluid);
    cd.cfg.dp=h.dp; cd.cfg.kernelh=float(h.h); cd.cfg.cteb=float(h.b); cd.cfg.rhopzero=float(h.rhop0); cd.cfg.gamma=float(h.gamma);
    cd.cfg.massbound=float(h.massbound); cd.cfg.massfluid=float(h.massfluid); cd.cfg.simulate2d=(h.data2d? 1: 0);
    memcpy(cd.caseposmin,h.caseposmin,24); memcpy(cd.caseposmax,h.caseposmax,24);
    memcpy(cd.cfg.maprealposmin,h.mapposmin,24); memcpy(cd.cfg.maprealposmax,h.mapposmax,24);
    cd.name=h.casename;
    cd.idp=p.idp; cd.pos=p.pos; cd.velrhop.resize(size_t(p.npok)*4);
    for(unsigned i=0;i<p.npok;i++){ cd.velrhop[4*size_t(i)]=p.vel[3*size_t(i)]; cd.velrhop[4*size_t(i)+1]=p.vel[3*size_t(i)+1]; cd.velrhop[4*size_t(i)+2]=p.vel[3*size_t(i)+2]; cd.velrhop[4*size_t(i)+3]=p.rhop[i]; }
    CaseToC(cd,out);
    unsigned nb=0; for(unsigned id:cd.idp)if(id<cd.cfg.casenfixed+cd.cfg.casenmoving)nb++;
    out->npb=nb;
    if(timestep)*timestep=p.timestep;
    if(step)*step=p.step;
    if(npiece)*npiece=h.npiece;
  });
  return r==DSPH_OK? r: DSPH_ERR_IO;
}
int dsph_run_case_ex(const dsph_runopts *r){
  if(!r||!r->dir||!r->casename||!r->dirout){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{
    RunOptions o; o.dir=r->dir; o.casename=r->casename; o.dirout=r->dirout; o.device=r->device; o.nsteps=r->nsteps; o.savepart=(r->savepart!=0);
    o.timemax=r->timemax; o.timeout=r->timeout; o.forcevariant=r->forcevariant;
    o.partbegin=(r->restart? r->partbegin: -1); if(r->partbegindir)o.partbegindir=r->partbegindir;
    o.mgpu=r->mgpu; o.rebalancesteps=r->rebalancesteps; o.boundweight=r->boundweight; o.piecemax=r->piecemax;
    RunCase(o);
  });
}
int dsph_run_case(const char *dir,const char *casename,const char *dirout,int device,int nsteps,int savepart){
  if(!dir||!casename||!dirout){ g_lasterror="Null argument."; return DSPH_ERR_ARG; }
  return Guard([&]{ RunOptions o; o.dir=dir; o.casename=casename; o.dirout=dirout; o.device=device; o.nsteps=nsteps; o.savepart=(savepart!=0); RunCase(o); });
}

} // extern "C"
