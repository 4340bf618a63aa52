// engine.cu -- host orchestration of the B200 hot path: the JSphGpuSingle / JCellDivGpuSingle work-alike behind the
// C ABI of include/dsph_b200.h.  One context = one GPU = one CUDA stream; every kernel is launched on that stream and
// timed with CUDA events on it.  Sequencing follows JSphGpuSingle::RunCellDivide / Interaction_Forces /
// ComputeStep_Sym / ComputeStep_Ver (S/JSphGpuSingle.cpp:331-417,435-486,548-596).
#include "engine.h"
#include <nvtx3/nvToolsExt.h>
#include "mdbc.cuh"
#include "radixsort.cuh"
#include "celldiv.cuh"
#include "forces_api.h"
#include "step.cuh"
#include "motion.cuh"
#include "periodic.cuh"
#include "mgpu.cuh"
#include "measure.cuh"
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <algorithm>

namespace dsph {

// NVTX ranges named like the reference's GPU timers (S/JDsTimersGpu.h:64-88: NL-Limits ... SU-ComputeStep): host-side ranges around the
// enqueue of each phase, projected onto the stream by Nsight Systems / Compute.  nvtx3 is header-only and resolves the tool's
// injection library at run time: no link dependency, a few ns per call without a tool attached.
static inline void NvtxBegin(const char *name){ nvtxRangePushA(name); }
static inline void NvtxEnd(){ nvtxRangePop(); }
static inline void NvtxNext(const char *name){ nvtxRangePop(); nvtxRangePushA(name); }


#define CUCHECK(x) do{ cudaError_t e_=(x); if(e_!=cudaSuccess)throw EngineError(DSPH_ERR_CUDA,std::string(#x)+": "+cudaGetErrorString(e_)); }while(0)

template<class T> static void DevAlloc(T *&p,size_t n){ if(p){ cudaFree(p); p=nullptr; } if(n)CUCHECK(cudaMalloc((void**)&p,n*sizeof(T))); }
template<class T> static void DevFree(T *&p){ if(p){ cudaFree(p); p=nullptr; } }

// Small host<->device transfers (cell domain, maxima, counters: 4..120 bytes) are done by a one-warp kernel through the
// mapped pinned block instead of cudaMemcpyAsync: a DMA engine serves its queue in order, so a 4-byte copy of one context
// would wait behind a 0.9 GB particle upload/download of another context (measured: the end-to-end pipeline of bench.py
// lost 25 ms per step to that).  The kernel runs on the context's stream like any other kernel.
__global__ void ker_small_copy(unsigned *dst,const unsigned *src,unsigned nwords){
  for(unsigned i=threadIdx.x;i<nwords;i+=blockDim.x)dst[i]=src[i];
  __threadfence_system();
}
static void SmallCopy(void *dst,const void *src,size_t bytes,cudaStream_t stm){
  ker_small_copy<<<1,32,0,stm>>>((unsigned*)dst,(const unsigned*)src,unsigned(bytes/4));
}

void Engine::SmallCopyAsync(void *dst,const void *src,size_t bytes){ SmallCopy(dst,src,bytes,stm); }

static int KeyBits(unsigned long long maxkey){ int b=1; while((maxkey>>b)!=0)b++; return b; }

Engine::Engine(int device_,const dsph_config &c,unsigned capacity):device(device_)
{
  int ndev=0;
  if(cudaGetDeviceCount(&ndev)!=cudaSuccess || ndev<=0)throw EngineError(DSPH_ERR_NOGPU,"No CUDA device is available: the B200 hot path has no CPU fallback.");
  if(device<0 || device>=ndev)throw EngineError(DSPH_ERR_ARG,"Invalid CUDA device index.");
  CUCHECK(cudaSetDevice(device));
  { int v=0; if(cudaDeviceGetAttribute(&v,cudaDevAttrMultiProcessorCount,device)==cudaSuccess && v>0)numsm=v; }
  CUCHECK(cudaStreamCreateWithFlags(&stm,cudaStreamNonBlocking));
  for(int i=0;i<NEVENTS;i++)CUCHECK(cudaEventCreate(&ev[i]));
  CUCHECK(cudaHostAlloc((void**)&hpin,sizeof(HostPinned),cudaHostAllocMapped|cudaHostAllocPortable));
  CUCHECK(cudaHostGetDevicePointer((void**)&hpind,hpin,0));
  memset(hpin,0,sizeof(HostPinned));
  Configure(c);
  Allocate(capacity);
  // the ring kernels (variants 9, 10, multiphase) address their ring relative to the start of the CTA's shared window:
  // check the assumption once per context; without it the plain kernels run (variant 4 / multiphase variant 1), a perf
  // cliff that the driver reports in Run.out (dsph_ringbase_ok)
  LaunchRingBaseProbe(&hpind->probe,stm);
  CUCHECK(cudaStreamSynchronize(stm));
  ringbase_ok=(hpin->probe==RingBaseExpected());
}

Engine::~Engine(){
  cudaSetDevice(device);
  MgCommDestroy();
  GaugesFree();
  if(measuredev)cudaFree(measuredev);
  if(mgown)cudaFree(mgown);
  if(ridpmove)cudaFree(ridpmove);
  if(stmcopy){ cudaStreamSynchronize(stmcopy); cudaStreamDestroy(stmcopy); cudaEventDestroy(evsnap); cudaEventDestroy(evsnapfree); }
  if(snapdev)cudaFree(snapdev);
  FreeAll();
  for(int i=0;i<NEVENTS;i++)if(ev[i])cudaEventDestroy(ev[i]);
  if(hpin)cudaFreeHost(hpin);
  if(stm)cudaStreamDestroy(stm);
}

void Engine::Configure(const dsph_config &c){
  cfg=c;
  DeriveConstants(cfg,k);
  if(cfg.casenfloat)throw EngineError(DSPH_ERR_ARG,"Floating bodies are outside the hot-path scope (moving boundaries are supported: dsph_motion_run).");
  if(cfg.tvisco==DSPH_VISCO_CONSTEQ && !cfg.nn_on)throw EngineError(DSPH_ERR_ARG,"ViscoTreatment 'Constitutive eq.' not valid for Single-phase classic formulation.");
  sps=(cfg.tvisco==DSPH_VISCO_LAMINARSPS && !cfg.nn_on);
  memset(&dc,0,sizeof(dc)); memset(&dph,0,sizeof(dph));
  if(cfg.nn_on){
    // JSph::InitMultiPhase + ConfigConstantsMP (M/JSph.cpp:3137-3242): per-phase B and particle mass, same rounding points
    if(cfg.nn_nphases<1 || cfg.nn_nphases>DSPH_MAXPHASES)throw EngineError(DSPH_ERR_ARG,"The number of phases is invalid.");
    if(cfg.nn_velgrad!=DSPH_VELGRAD_FDA && cfg.nn_velgrad!=DSPH_VELGRAD_SPH)throw EngineError(DSPH_ERR_ARG,"Velocity gradient type is not valid.");
    dph.count=cfg.nn_nphases; dph.velgrad=cfg.nn_velgrad; dph.lamda=cfg.nn_relaxationdt;
    bool cs0present=true;
    for(int i=0;i<dph.count;i++){
      const dsph_phase &s=cfg.phases[i]; DevPhase &d=dph.ph[i];
      d.rho=s.rho; d.cs0=s.cs0; d.gamma=(s.gamma? s.gamma: cfg.gamma);
      d.cteb=(d.cs0? d.cs0*d.cs0*d.rho/d.gamma: 0.f);
      d.visco=s.visco; d.tau_yield=s.tau_yield; d.tau_max=s.tau_max; d.bi_multi=(s.tau_max? s.bi_multi: 0.f); d.m_nn=s.m_nn; d.n_nn=s.n_nn;
      if(!d.cs0)cs0present=false;
    }
    for(int i=0;i<dph.count;i++){
      DevPhase &d=dph.ph[i];
      if(!cs0present)d.cteb=float(k.cs0*k.cs0*d.rho/d.gamma);
      d.mass=float(d.rho*(cfg.simulate2d? cfg.dp*cfg.dp: cfg.dp*cfg.dp*cfg.dp));
    }
  }
  dc.nn_on=cfg.nn_on;
  dc.kernelh=cfg.kernelh; dc.ovkernelh=1.f/cfg.kernelh; dc.kernelsize2=k.kernelsize2; dc.eta2=k.eta2; dc.awen=k.awen; dc.bwen=k.bwen;
  dc.cub_a2=k.cub_a2; dc.cub_a24=k.cub_a24; dc.cub_c1=k.cub_c1; dc.cub_d1=k.cub_d1; dc.cub_c2=k.cub_c2; dc.cub_od_wdeltap=k.cub_od_wdeltap;
  dc.bwen_ovh=float(double(k.bwen)/double(cfg.kernelh)); dc.cub_c1_ovh=float(double(k.cub_c1)/double(cfg.kernelh)); dc.cub_d1_ovh=float(double(k.cub_d1)/double(cfg.kernelh));
  dc.cteb=cfg.cteb; dc.gamma=cfg.gamma; dc.ovgamma=1.f/cfg.gamma; dc.rhopzero=cfg.rhopzero; dc.ovrhopzero=1.0f/cfg.rhopzero;
  dc.massfluid=cfg.massfluid; dc.massbound=cfg.massbound; dc.cteb_d=double(cfg.cteb); dc.gamma_d=double(cfg.gamma);
  dc.cs0=float(k.cs0); dc.visco=cfg.visco; dc.viscobound=cfg.visco*cfg.viscoboundfactor; dc.ddtkh=k.ddtkh; dc.ddtgz=k.ddtgz;
  dc.spssmag=k.spssmag; dc.spsblin=k.spsblin;
  { // series of (1+t)^g-1 with g=1/gamma (float, like the reference's 1.f/Gamma), coefficients in double
    const double g=double(dc.ovgamma); double cf=1;
    for(int i=0;i<6;i++){ cf*=(g-i)/double(i+1); dc.powc[i]=float(cf); }
  }
  dc.scell=k.scell; dc.scelldiv=k.scelldiv; dc.scell_d=double(k.scell); dc.hscale=2.f/k.scell;
  for(int i=0;i<3;i++){
    dc.domposmin[i]=k.domposmin[i]; dc.domcells[i]=k.domcells[i];
    dc.maprealposmin[i]=k.maprealposmin[i]; dc.maprealsize[i]=k.maprealsize[i];
    dc.perixinc[i]=k.perixinc[i]; dc.periyinc[i]=k.periyinc[i]; dc.perizinc[i]=k.perizinc[i];
    dc.gravity[i]=cfg.gravity[i];
  }
  dc.domcellcode=k.domcellcode;
  dc.perix=cfg.perix; dc.periy=cfg.periy; dc.periz=cfg.periz; dc.periactive=k.periactive; dc.ghosts=(k.periactive||mgpu)? 1: 0;
  dc.movlimit=k.movlimit;
  dc.rhopoutmin=cfg.rhopoutmin; dc.rhopoutmax=cfg.rhopoutmax;
  if(!(dc.rhopoutmin<dc.rhopoutmax)){ dc.rhopoutmin=-FLT_MAX; dc.rhopoutmax=FLT_MAX; }
  dc.simulate2d=cfg.simulate2d; dc.dtallparticles=cfg.dtallparticles;
  dc.shiftmode=cfg.shiftmode; dc.shifttfs=cfg.shifttfs;
  dc.reftail=(reftail? 1: 0);
  mdbc=((cfg.tboundary&0xff)==DSPH_BOUNDARY_MDBC); mdbccorrector=(mdbc && (cfg.tboundary&DSPH_BOUNDARY_MDBCCORRECTOR)!=0);
  dc.mdbc=(mdbc? 1: 0);
  if(mdbc && cfg.nn_on)throw EngineError(DSPH_ERR_ARG,"mDBC belongs to the v5.2 formulation: it cannot be combined with the multiphase fork.");
  if(mdbc && (cfg.perix||cfg.periy||cfg.periz))throw EngineError(DSPH_ERR_ARG,"mDBC with periodic boundaries is not implemented (the duplicates would need their normals).");
  symplecticdtpre=k.dtini; verletstep=0;
}

// dsph_constants_upload on a live context (the reference re-uploads StCteInteraction when the DDT ramp changes a constant,
// S/JSphGpuSingle.cpp:520-539).  Everything that decides which arrays exist or how cells are laid out must stay as it was
// at creation: those fields are compared and the call is rejected with DSPH_ERR_ARG, the context stays untouched.
void Engine::Reconfigure(const dsph_config &c){
  SphConst k2; DeriveConstants(c,k2);
  bool same=(c.tboundary==cfg.tboundary && c.tstep==cfg.tstep && c.shiftmode==cfg.shiftmode && c.tvisco==cfg.tvisco && c.tkernel==cfg.tkernel && c.nn_on==cfg.nn_on &&
    c.nn_velgrad==cfg.nn_velgrad && c.nn_nphases==cfg.nn_nphases && c.perix==cfg.perix && c.periy==cfg.periy && c.periz==cfg.periz &&
    c.simulate2d==cfg.simulate2d && c.cellmode==cfg.cellmode && c.casenfixed==cfg.casenfixed && c.casenmoving==cfg.casenmoving &&
    c.casenfloat==cfg.casenfloat && k2.scell==k.scell && k2.scelldiv==k.scelldiv && k2.domcellcode==k.domcellcode && k2.periactive==k.periactive);
  for(int i=0;i<3 && same;i++)same=(k2.domcells[i]==k.domcells[i] && k2.domposmin[i]==k.domposmin[i] && k2.maprealposmin[i]==k.maprealposmin[i] && k2.maprealsize[i]==k.maprealsize[i]);
  if(!same)throw EngineError(DSPH_ERR_ARG,"dsph_constants_upload: the formulation (step algorithm, viscosity, shifting, kernel, phases, periodicity) and the cell domain are fixed at dsph_ctx_create; create a new context to change them.");
  const double dtpre=symplecticdtpre; const int vs=verletstep;
  Configure(c);
  symplecticdtpre=dtpre; verletstep=vs;
  forcesok=false;    // forces computed with the previous constants must not feed an update
}

void Engine::FreeAll(){
  for(int b=0;b<2;b++){
    DevFree(idp[b]); DevFree(code[b]); DevFree(dcell[b]); DevFree(pos[b]); DevFree(velrhop[b]);
    DevFree(velrhopm1[b]); DevFree(pospre[b]); DevFree(velrhoppre[b]); DevFree(spstau[b]); DevFree(boundnormal[b]);
    DevFree(rs.keys[b]); DevFree(rs.vals[b]);
  }
  DevFree(posp); DevFree(posh); DevFree(acear); DevFree(shiftposfs); DevFree(spsgradvel); DevFree(viscoeta); DevFree(nntau);
  DevFree(rs.hist); DevFree(rs.chunksum); DevFree(rs.chunksum2);
  DevFree(begincell); DevFree(sufmin); DevFree(divparams); DevFree(limits); DevFree(fmax); DevFree(perilist); DevFree(pericount);
}

void Engine::Allocate(unsigned capacity){
  cap=std::max(capacity,1024u);
  const bool verlet=(cfg.tstep==DSPH_STEP_VERLET), sympl=(cfg.tstep==DSPH_STEP_SYMPLECTIC);
  if(cfg.nn_on && cfg.nn_velgrad==DSPH_VELGRAD_SPH && cfg.tvisco!=DSPH_VISCO_ARTIFICIAL){
    DevAlloc(viscoeta,cap);
    if(cfg.tvisco==DSPH_VISCO_CONSTEQ)DevAlloc(nntau,size_t(cap)*6);
  }
  for(int b=0;b<2;b++){
    DevAlloc(idp[b],cap); DevAlloc(code[b],cap); DevAlloc(dcell[b],cap); DevAlloc(pos[b],size_t(cap)*3); DevAlloc(velrhop[b],cap);
    if(verlet)DevAlloc(velrhopm1[b],cap);
    if(sympl){ DevAlloc(pospre[b],size_t(cap)*3); DevAlloc(velrhoppre[b],cap); }
    if(sps)DevAlloc(spstau[b],size_t(cap)*6);
    if(mdbc){ DevAlloc(boundnormal[b],cap); CUCHECK(cudaMemsetAsync(boundnormal[b],0,size_t(cap)*sizeof(float4),stm)); }
    DevAlloc(rs.keys[b],cap); DevAlloc(rs.vals[b],cap);
  }
  // padded: the chunked candidate loops of the ring kernels read past a lane's own cell range (forces_ring.cuh: up to the
  // largest cell population among the lanes of a warp + 7 entries; forces_v10.cuh: 16 entries); POSP_PAD bounds the cell
  // population the ring kernels accept (checked per divide: dp.maxcellpop), denser states run variant 4
  DevAlloc(posp,size_t(cap)+POSP_PAD);
  CUCHECK(cudaMemsetAsync(posp,0,(size_t(cap)+POSP_PAD)*sizeof(float4),stm));
  DevAlloc(posh,(size_t(cap)+POSP_PAD)/2+1);
  CUCHECK(cudaMemsetAsync(posh,0,((size_t(cap)+POSP_PAD)/2+1)*sizeof(uint4),stm));
  DevAlloc(acear,cap);
  if(cfg.shiftmode)DevAlloc(shiftposfs,cap);
  if(sps)DevAlloc(spsgradvel,size_t(cap)*6);
  const unsigned ntiles=RsTiles(cap);
  rs.capacity=cap; rs.histcap=256u*ntiles;
  DevAlloc(rs.hist,rs.histcap);
  // scan scratch: used for the digit histograms (histcap entries) AND for the per-particle periodic counts (up to cap entries)
  rs.chunksumcap=ScChunks(std::max(rs.histcap,cap))+1; rs.chunksum2cap=ScChunks(rs.chunksumcap)+1;
  DevAlloc(rs.chunksum,rs.chunksumcap);
  DevAlloc(rs.chunksum2,rs.chunksum2cap);
  const unsigned long long nctmax=(unsigned long long)k.domcells[0]*k.domcells[1]*k.domcells[2];
  if(nctmax*2+6>=0x7fffffffull)throw EngineError(DSPH_ERR_ARG,"The number of cells is too big.");
  maxnctt=unsigned(nctmax*2+6);
  keybits=KeyBits(maxnctt);
  UpdateCellBound();
  DevAlloc(begincell,size_t(maxnctt)+8);
  DevAlloc(sufmin,size_t((maxnctt+SM_CHUNK-1)/SM_CHUNK)+8);
  DevAlloc(divparams,1); DevAlloc(limits,12); DevAlloc(fmax,1);
  CUCHECK(cudaMemsetAsync(divparams,0,sizeof(DivParams),stm));
  if(k.periactive){ DevAlloc(perilist,size_t(cap)); DevAlloc(pericount,4); }
  cur=0;
}

void Engine::Upload(unsigned np_,unsigned npb_,const unsigned *hidp,const typecode *hcode,const double *hpos,const float *hvelrhop){
  CUCHECK(cudaSetDevice(device));
  if(np_>cap)throw EngineError(DSPH_ERR_ARG,"The number of particles exceeds the capacity of the context.");
  if(npb_>np_)throw EngineError(DSPH_ERR_ARG,"npb > np.");
  np=np_; npb=npb_; npbok=npb_; npbper=npfper=0; haspre=false; boundchanged=true; blimok=false; divok=false; celltableok=false;
  std::vector<typecode> tmpcode;
  if(!hcode){
    tmpcode.resize(np);
    for(unsigned p=0;p<np;p++)tmpcode[p]=typecode(hidp[p]<cfg.casenfixed? DSPH_CODE_TYPE_FIXED: (hidp[p]<cfg.casenfixed+cfg.casenmoving? DSPH_CODE_TYPE_MOVING: DSPH_CODE_TYPE_FLUID));
    hcode=tmpcode.data();
  }
  CUCHECK(cudaMemcpyAsync(idp[cur],hidp,size_t(np)*4,cudaMemcpyHostToDevice,stm));
  CUCHECK(cudaMemcpyAsync(code[cur],hcode,size_t(np)*2,cudaMemcpyHostToDevice,stm));
  CUCHECK(cudaMemcpyAsync(pos[cur],hpos,size_t(np)*24,cudaMemcpyHostToDevice,stm));
  CUCHECK(cudaMemcpyAsync(velrhop[cur],hvelrhop,size_t(np)*16,cudaMemcpyHostToDevice,stm));
  if(cfg.tstep==DSPH_STEP_VERLET)CUCHECK(cudaMemcpyAsync(velrhopm1[cur],hvelrhop,size_t(np)*16,cudaMemcpyHostToDevice,stm));
  if(sps)CUCHECK(cudaMemsetAsync(spstau[cur],0,size_t(np)*24,stm));
  // the upload itself is copies only (pure DMA: it overlaps the kernels of other contexts); the first cell hash runs with
  // the first consumer of dcell (EnsureDcell)
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
  dcellpending=true;
  symplecticdtpre=k.dtini; verletstep=0; forcesok=false;
}

// JSph::LoadDcellParticles (S/JSph.cpp:1690-1711): first cell hash, same double arithmetic, on the device
void Engine::EnsureDcell(){
  if(!dcellpending)return;
  dcellpending=false;
  CUCHECK(cudaMemsetAsync(limits,0,4,stm));
  if(np){ ker_upload_dcell<<<(np+255)/256,256,0,stm>>>(np,pos[cur],code[cur],dcell[cur],limits,dc,
      k.domrealposmin[0],k.domrealposmin[1],k.domrealposmin[2],k.domrealposmax[0],k.domrealposmax[1],k.domrealposmax[2]); launches++; }
  SmallCopy(&hpind->scratch[0],limits,4,stm);
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
  if(hpin->scratch[0])throw EngineError(DSPH_ERR_PHYSICS,"Found new particles out.");
}

void Engine::SetVelrhopM1(const float *h){
  if(cfg.tstep!=DSPH_STEP_VERLET)throw EngineError(DSPH_ERR_STATE,"VelrhopM1 is only used by Verlet.");
  CUCHECK(cudaMemcpyAsync(velrhopm1[cur],h,size_t(np)*16,cudaMemcpyHostToDevice,stm));
  CUCHECK(cudaStreamSynchronize(stm));
}

void Engine::SetSpsTau(const float *h){
  if(!sps)throw EngineError(DSPH_ERR_STATE,"SpsTau is only used by Laminar+SPS.");
  CUCHECK(cudaMemcpyAsync(spstau[cur],h,size_t(np)*24,cudaMemcpyHostToDevice,stm));
  CUCHECK(cudaStreamSynchronize(stm));
}

// normal vectors particle -> ghost node of the first `count` particles of the uploaded order (boundary particles come first)
void Engine::SetBoundNormals(const float *h,unsigned count){
  if(!mdbc)throw EngineError(DSPH_ERR_STATE,"Normal vectors are only used with cfg.tboundary = DSPH_BOUNDARY_MDBC.");
  if(count>np)throw EngineError(DSPH_ERR_ARG,"More normal vectors than particles.");
  std::vector<float4> tmp(count);
  for(unsigned i=0;i<count;i++)tmp[i]=make_float4(h[3*size_t(i)],h[3*size_t(i)+1],h[3*size_t(i)+2],0.f);
  CUCHECK(cudaMemsetAsync(boundnormal[cur],0,size_t(cap)*sizeof(float4),stm));
  if(count)CUCHECK(cudaMemcpyAsync(boundnormal[cur],tmp.data(),size_t(count)*sizeof(float4),cudaMemcpyHostToDevice,stm));
  CUCHECK(cudaStreamSynchronize(stm));
}
void Engine::GetBoundNormals(float *h,unsigned count){
  if(!mdbc)throw EngineError(DSPH_ERR_STATE,"Normal vectors are only used with cfg.tboundary = DSPH_BOUNDARY_MDBC.");
  if(count>np)throw EngineError(DSPH_ERR_ARG,"More normal vectors than particles.");
  std::vector<float4> tmp(count);
  if(count)CUCHECK(cudaMemcpyAsync(tmp.data(),boundnormal[cur],size_t(count)*sizeof(float4),cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaStreamSynchronize(stm));
  for(unsigned i=0;i<count;i++){ h[3*size_t(i)]=tmp[i].x; h[3*size_t(i)+1]=tmp[i].y; h[3*size_t(i)+2]=tmp[i].z; }
}
void Engine::MdbcCorrection(){
  if(!npbok)return;
  MdbcArgs m; m.npbok=npbok; m.pos=pos[cur]; m.boundnormal=boundnormal[cur]; m.velrhop=velrhop[cur]; m.posp=posp;
  m.determlimit=1e-3f; m.threshold=mdbcthreshold;                      // S/JSphCpu.cpp:1197
  const ForceParams t=MakeForceParams();
  const unsigned need=(npbok+(FORCES_THREADS/32)-1)/(FORCES_THREADS/32), nb=std::min(need,unsigned(numsm)*16u);
  const bool wend=(cfg.tkernel==DSPH_KERNEL_WENDLAND);
  if(cfg.simulate2d){ if(wend)ker_mdbc_correction<DSPH_KERNEL_WENDLAND,true><<<nb,FORCES_THREADS,0,stm>>>(m,t,dc); else ker_mdbc_correction<DSPH_KERNEL_CUBIC,true><<<nb,FORCES_THREADS,0,stm>>>(m,t,dc); }
  else{ if(wend)ker_mdbc_correction<DSPH_KERNEL_WENDLAND,false><<<nb,FORCES_THREADS,0,stm>>>(m,t,dc); else ker_mdbc_correction<DSPH_KERNEL_CUBIC,false><<<nb,FORCES_THREADS,0,stm>>>(m,t,dc); }
  launches++;
}

void Engine::Download(unsigned first,unsigned count,unsigned *hidp,typecode *hcode,unsigned *hdcell,double *hpos,float *hvelrhop){
  if(first+count>np)throw EngineError(DSPH_ERR_ARG,"Range out of particles.");
  if(hdcell)EnsureDcell();
  if(hidp)CUCHECK(cudaMemcpyAsync(hidp,idp[cur]+first,size_t(count)*4,cudaMemcpyDeviceToHost,stm));
  if(hcode)CUCHECK(cudaMemcpyAsync(hcode,code[cur]+first,size_t(count)*2,cudaMemcpyDeviceToHost,stm));
  if(hdcell)CUCHECK(cudaMemcpyAsync(hdcell,dcell[cur]+first,size_t(count)*4,cudaMemcpyDeviceToHost,stm));
  if(hpos)CUCHECK(cudaMemcpyAsync(hpos,pos[cur]+size_t(first)*3,size_t(count)*24,cudaMemcpyDeviceToHost,stm));
  if(hvelrhop)CUCHECK(cudaMemcpyAsync(hvelrhop,velrhop[cur]+first,size_t(count)*16,cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaStreamSynchronize(stm));
}

// ------------------------------------------------------------------------------------------------
// PART snapshots
// ------------------------------------------------------------------------------------------------
void Engine::HostSnapAlloc(HostSnap &hs,unsigned capacity){
  CUCHECK(cudaSetDevice(device));
  HostSnapFree(hs);
  CUCHECK(cudaHostAlloc((void**)&hs.idp,size_t(capacity)*4,cudaHostAllocDefault));
  CUCHECK(cudaHostAlloc((void**)&hs.code,size_t(capacity)*sizeof(typecode),cudaHostAllocDefault));
  CUCHECK(cudaHostAlloc((void**)&hs.pos,size_t(capacity)*24,cudaHostAllocDefault));
  CUCHECK(cudaHostAlloc((void**)&hs.velrhop,size_t(capacity)*16,cudaHostAllocDefault));
  CUCHECK(cudaEventCreateWithFlags(&hs.done,cudaEventDisableTiming));
  hs.capacity=capacity; hs.np=0;
}
void Engine::HostSnapFree(HostSnap &hs){
  if(hs.done){ cudaEventSynchronize(hs.done); cudaEventDestroy(hs.done); hs.done=nullptr; }
  if(hs.idp)cudaFreeHost(hs.idp); if(hs.code)cudaFreeHost(hs.code); if(hs.pos)cudaFreeHost(hs.pos); if(hs.velrhop)cudaFreeHost(hs.velrhop);
  hs.idp=nullptr; hs.code=nullptr; hs.pos=nullptr; hs.velrhop=nullptr; hs.capacity=0;
}
void Engine::SnapshotBegin(HostSnap &hs){
  CUCHECK(cudaSetDevice(device));
  if(np>hs.capacity)throw EngineError(DSPH_ERR_ARG,"Snapshot buffers are smaller than the number of particles.");
  if(!stmcopy){
    CUCHECK(cudaStreamCreateWithFlags(&stmcopy,cudaStreamNonBlocking));
    CUCHECK(cudaEventCreateWithFlags(&evsnap,cudaEventDisableTiming));
    CUCHECK(cudaEventCreateWithFlags(&evsnapfree,cudaEventDisableTiming));
  }
  if(snapcap<cap){                                   // 46 bytes per particle: idp 4, code 2, pos 24, velrhop 16
    if(snapdev){ CUCHECK(cudaStreamSynchronize(stmcopy)); cudaFree(snapdev); snapdev=nullptr; }
    CUCHECK(cudaMalloc((void**)&snapdev,size_t(cap)*48));
    snapcap=cap; snapbusy=false;
  }
  char *dpos=snapdev, *dvr=snapdev+size_t(snapcap)*24, *did=dvr+size_t(snapcap)*16, *dcode=did+size_t(snapcap)*4;
  const size_t n=np;
  if(snapbusy)CUCHECK(cudaStreamWaitEvent(stm,evsnapfree,0));          // the previous snapshot has left the staging area
  CUCHECK(cudaMemcpyAsync(dpos,pos[cur],n*24,cudaMemcpyDeviceToDevice,stm));
  CUCHECK(cudaMemcpyAsync(dvr,velrhop[cur],n*16,cudaMemcpyDeviceToDevice,stm));
  CUCHECK(cudaMemcpyAsync(did,idp[cur],n*4,cudaMemcpyDeviceToDevice,stm));
  CUCHECK(cudaMemcpyAsync(dcode,code[cur],n*sizeof(typecode),cudaMemcpyDeviceToDevice,stm));
  CUCHECK(cudaEventRecord(evsnap,stm));
  CUCHECK(cudaStreamWaitEvent(stmcopy,evsnap,0));
  CUCHECK(cudaMemcpyAsync(hs.pos,dpos,n*24,cudaMemcpyDeviceToHost,stmcopy));
  CUCHECK(cudaMemcpyAsync(hs.velrhop,dvr,n*16,cudaMemcpyDeviceToHost,stmcopy));
  CUCHECK(cudaMemcpyAsync(hs.idp,did,n*4,cudaMemcpyDeviceToHost,stmcopy));
  CUCHECK(cudaMemcpyAsync(hs.code,dcode,n*sizeof(typecode),cudaMemcpyDeviceToHost,stmcopy));
  CUCHECK(cudaEventRecord(evsnapfree,stmcopy));
  CUCHECK(cudaEventRecord(hs.done,stmcopy));
  hs.np=np; snapbusy=true;
}
void Engine::SnapshotWait(const HostSnap &hs){
  CUCHECK(cudaSetDevice(device));
  if(hs.done)CUCHECK(cudaEventSynchronize(hs.done));
}

// ------------------------------------------------------------------------------------------------
// Periodic duplicates (JSphCpuSingle::RunPeriodic order, S/JSphCpuSingle.cpp:361-431)
// ------------------------------------------------------------------------------------------------
void Engine::RunPeriodic(){
  const int B=256;
  // mark current periodic particles to be ignored
  ker_peri_ignore<<<(np+B-1)/B,B,0,stm>>>(np,code[cur]); launches++;
  const unsigned npb0=npb, npf0=np-npb;
  npbper=npfper=0; boundchanged=true; celltableok=false;
  for(unsigned ctype=0;ctype<2;ctype++){
    const unsigned pini=(ctype? npb0: 0), num=(ctype? npf0: npb0);
    for(unsigned cper=0;cper<3;cper++)if((cper==0&&cfg.perix)||(cper==1&&cfg.periy)||(cper==2&&cfg.periz)){
      const double *inc=(cper==0? k.perixinc: (cper==1? k.periyinc: k.perizinc));
      for(unsigned cblock=0;cblock<2;cblock++){
        const unsigned nper=(ctype? npfper: npbper);
        const unsigned pini2=(cblock? pini: np-nper), num2=(cblock? num: nper);
        if(!num2)continue;
        // ordered list (ascending particle, +inc before -inc) = the CPU list; built with a two-pass count/scan/write
        PeriArgs pa; pa.n=num2; pa.pini=pini2;
        for(int i=0;i<3;i++){ pa.inc[i]=inc[i]; pa.mapmin[i]=k.mapposmin[i]; pa.mapmax[i]=k.mapposmax[i]; }
        pa.code=code[cur]; pa.pos=pos[cur];
        unsigned *flags=rs.keys[0];   // scratch: count per particle (0..2), scanned in place
        ker_peri_count<<<(num2+B-1)/B,B,0,stm>>>(pa,flags); launches++;
        // total = last exclusive offset + last count: keep the last count
        SmallCopy(&hpind->scratch[0],flags+(num2-1),4,stm);
        if(DeviceExclusiveScan(flags,num2,rs.chunksum,rs.chunksumcap,rs.chunksum2,rs.chunksum2cap,stm,&launches)<0)
          throw EngineError(DSPH_ERR_STATE,"Scan scratch too small for the periodic pass.");
        SmallCopy(&hpind->scratch[1],flags+(num2-1),4,stm);
        CUCHECK(cudaStreamSynchronize(stm));
        const unsigned count=hpin->scratch[0]+hpin->scratch[1];
        if(!count)continue;
        if(size_t(np)+count>cap)throw EngineError(DSPH_ERR_STATE,"Periodic particles exceed the capacity of the context (create it with a larger np_capacity).");
        PeriDupArrays d;
        d.idp=idp[cur]; d.code=code[cur]; d.dcell=dcell[cur]; d.pos=pos[cur]; d.velrhop=velrhop[cur];
        d.velrhopm1=(cfg.tstep==DSPH_STEP_VERLET? velrhopm1[cur]: nullptr);
        d.pospre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? pospre[cur]: nullptr);
        d.velrhoppre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? velrhoppre[cur]: nullptr);
        d.spstau=(sps? spstau[cur]: nullptr);
        ker_peri_duplicate<<<(num2+B-1)/B,B,0,stm>>>(pa,flags,np,d,dc); launches++;
        np+=count;
        if(!ctype)npbper+=count; else npfper+=count;
      }
    }
  }
}

// Slab runs: periodic copies along y / z are made AFTER the slab exchange, from every live particle of the rank (owned, freshly arrived
// migrants and x-halo copies alike, so the copies at the edges of the cut faces exist too).  The stale copies were retired by the
// classification of the exchange.  Sources of an axis = the particle blocks of the type + the copies earlier axes appended; the
// order of the copies differs from the single-GPU list (S/JSphCpuSingle.cpp:182-213), which slab runs do not promise to reproduce.
void Engine::RunPeriodicSlab(unsigned npb_owned,unsigned np_owned,unsigned nbrecv,unsigned nfrecv){
  const int B=256;
  struct Range{ unsigned ini,num; };
  std::vector<Range> src[2];
  src[0].push_back({0u,npb_owned}); src[0].push_back({np_owned,nbrecv});
  src[1].push_back({npb_owned,np_owned-npb_owned}); src[1].push_back({np_owned+nbrecv,nfrecv});
  for(unsigned cper=1;cper<3;cper++)if((cper==1&&cfg.periy)||(cper==2&&cfg.periz)){
    const double *inc=(cper==1? k.periyinc: k.perizinc);
    std::vector<Range> added[2];
    for(unsigned ctype=0;ctype<2;ctype++)for(const Range &r:src[ctype]){
      if(!r.num)continue;
      PeriArgs pa; pa.n=r.num; pa.pini=r.ini;
      for(int i=0;i<3;i++){ pa.inc[i]=inc[i]; pa.mapmin[i]=k.mapposmin[i]; pa.mapmax[i]=k.mapposmax[i]; }
      pa.code=code[cur]; pa.pos=pos[cur];
      unsigned *flags=rs.keys[0];
      ker_peri_count<<<(r.num+B-1)/B,B,0,stm>>>(pa,flags); launches++;
      SmallCopy(&hpind->scratch[0],flags+(r.num-1),4,stm);
      if(DeviceExclusiveScan(flags,r.num,rs.chunksum,rs.chunksumcap,rs.chunksum2,rs.chunksum2cap,stm,&launches)<0)
        throw EngineError(DSPH_ERR_STATE,"Scan scratch too small for the periodic pass.");
      SmallCopy(&hpind->scratch[1],flags+(r.num-1),4,stm);
      CUCHECK(cudaStreamSynchronize(stm));
      const unsigned count=hpin->scratch[0]+hpin->scratch[1];
      if(!count)continue;
      if(size_t(np)+count>cap)throw EngineError(DSPH_ERR_STATE,"Periodic particles exceed the capacity of the context (create it with a larger np_capacity).");
      PeriDupArrays d;
      d.idp=idp[cur]; d.code=code[cur]; d.dcell=dcell[cur]; d.pos=pos[cur]; d.velrhop=velrhop[cur];
      d.velrhopm1=(cfg.tstep==DSPH_STEP_VERLET? velrhopm1[cur]: nullptr);
      d.pospre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? pospre[cur]: nullptr);
      d.velrhoppre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? velrhoppre[cur]: nullptr);
      d.spstau=(sps? spstau[cur]: nullptr);
      ker_peri_duplicate<<<(r.num+B-1)/B,B,0,stm>>>(pa,flags,np,d,dc); launches++;
      added[ctype].push_back({np,count});
      np+=count;
      if(!ctype)npbper+=count; else npfper+=count;
    }
    for(unsigned ctype=0;ctype<2;ctype++)for(const Range &r:added[ctype])src[ctype].push_back(r);
  }
  boundchanged=true; celltableok=false;
}

// ------------------------------------------------------------------------------------------------
// RunCellDivide
// ------------------------------------------------------------------------------------------------
void Engine::CellDivide(bool updateperiodic,dsph_divinfo *info){
  CUCHECK(cudaSetDevice(device));
  EnsureDcell();
  const int B=DIV_THREADS;
  NvtxBegin("SU-Periodic");
  CUCHECK(cudaEventRecord(ev[0],stm));
  if(updateperiodic && k.periactive && !mgpu)RunPeriodic();
  NvtxNext("NL-Limits");
  CUCHECK(cudaEventRecord(ev[1],stm));
  const unsigned npb1=npb, npf1=np-npb-npbper-npfper, npb2=npbper, npf2=npfper;
  const unsigned nptot=np;
  if(boundchanged || k.periactive || mgpu){ blimok=false; divok=false; }
  // ---- limits (one reduction kernel per particle range, results stay on the device)
  {
    unsigned init[12]={0xffffffffu,0xffffffffu,0xffffffffu,0,0,0, 0xffffffffu,0xffffffffu,0xffffffffu,0,0,0};
    memcpy(hpin->limitsinit,init,sizeof(init));
    SmallCopy(limits,hpind->limitsinit,sizeof(init),stm);
    const int nsm=numsm*4;
    auto lim=[&](unsigned n,unsigned pini,unsigned *dst){ if(n){ ker_div_limits<<<std::min<unsigned>((n+B-1)/B,nsm),B,0,stm>>>(n,pini,dcell[cur],code[cur],k.domcellcode,dst); launches++; } };
    if(mgpu && k.periactive){      // slab + periodic: boundary and fluid copies are appended in any order behind the owned particles
      if(nptot){ ker_div_limits_bytype<<<std::min<unsigned>((nptot+B-1)/B,nsm),B,0,stm>>>(nptot,dcell[cur],code[cur],k.domcellcode,limits,limits+6); launches++; }
    }
    else{
      if(!blimok){ lim(npb1,0,limits); lim(npb2,npb1+npf1,limits); }
      lim(npf1,npb1,limits+6); lim(npf2,npb1+npf1+npb2,limits+6);
    }
    ker_div_prepare<<<1,1,0,stm>>>(limits,limits+6,blimok? 0: 1,k.domcells[0],k.domcells[1],k.domcells[2],unsigned(k.scelldiv),nptot,effnctt,divparams); launches++;
    blimok=true;
  }
  NvtxNext("NL-PreSort");
  CUCHECK(cudaEventRecord(ev[2],stm));
  // ---- sort keys
  if(nptot){ ker_div_presort<<<(nptot+B-1)/B,B,0,stm>>>(nptot,dcell[cur],code[cur],k.domcellcode,divparams,rs.keys[0]); launches++; }
  NvtxNext("NL-RadixSort");
  CUCHECK(cudaEventRecord(ev[3],stm));
  const int sidx=RadixSortPairs(rs,nptot,effkeybits,1,stm,&launches);
  NvtxNext("NL-CellBegin");
  CUCHECK(cudaEventRecord(ev[4],stm));
  // ---- begin-cell table
  CUCHECK(cudaMemsetAsync(begincell,0xff,size_t(effnctt)*4,stm));
  {
    const unsigned nb=std::max(1u,(nptot+B-1)/B);
    ker_div_cellmark<<<nb,B,0,stm>>>(nptot,rs.keys[sidx],divparams,begincell); launches++;
    const unsigned nchunks=(effnctt+SM_CHUNK-1)/SM_CHUNK;
    ker_div_sufmin_chunks<<<nchunks,SM_THREADS,0,stm>>>(begincell,divparams,sufmin); launches++;
    if(nchunks>1){
      ker_div_sufmin_sums<<<1,SM_THREADS,0,stm>>>(sufmin,nchunks); launches++;
      ker_div_sufmin_apply<<<nchunks,SM_THREADS,0,stm>>>(begincell,divparams,sufmin); launches++;
    }
    ker_div_maxpop<<<(effnctt+255)/256,256,0,stm>>>(begincell,divparams); launches++;
    ker_div_counts<<<1,1,0,stm>>>(begincell,npb1+npb2,divparams); launches++;
  }
  NvtxNext("NL-SortData");
  CUCHECK(cudaEventRecord(ev[5],stm));
  // ---- fused reorder of every array (+ cell-relative position and pressure)
  if(nptot){
    const int o=cur^1;
    ReorderArrays a; memset(&a,0,sizeof(a));
    a.idp_in=idp[cur]; a.idp_out=idp[o]; a.code_in=code[cur]; a.code_out=code[o]; a.dcell_in=dcell[cur]; a.dcell_out=dcell[o];
    a.pos_in=pos[cur]; a.pos_out=pos[o]; a.velrhop_in=velrhop[cur]; a.velrhop_out=velrhop[o];
    if(cfg.tstep==DSPH_STEP_VERLET){ a.velrhopm1_in=velrhopm1[cur]; a.velrhopm1_out=velrhopm1[o]; }
    if(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre){ a.pospre_in=pospre[cur]; a.pospre_out=pospre[o]; a.velrhoppre_in=velrhoppre[cur]; a.velrhoppre_out=velrhoppre[o]; }
    if(sps){ a.spstau_in=spstau[cur]; a.spstau_out=spstau[o]; }
    if(mdbc){ a.boundnormal_in=boundnormal[cur]; a.boundnormal_out=boundnormal[o]; }
    a.posp=posp; a.posh=posh;
    if(cfg.nn_on)ker_div_reorder<true><<<(nptot+B-1)/B,B,0,stm>>>(nptot,rs.vals[sidx],a,dc,dph,divparams);
    else ker_div_reorder<false><<<(nptot+B-1)/B,B,0,stm>>>(nptot,rs.vals[sidx],a,dc,dph,divparams);
    launches++;
    cur=o;
  }
  NvtxNext("NL-OutCheck");
  CUCHECK(cudaEventRecord(ev[6],stm));
  // ---- the only device->host transfer of the divide: cell domain + counts (one small async copy + sync)
  SmallCopy(&hpind->dp,divparams,sizeof(DivParams),stm);
  CUCHECK(cudaEventRecord(ev[7],stm));
  NvtxEnd();
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
  dp=hpin->dp;
  if(dp.error)throw EngineError(DSPH_ERR_STATE,(effnctt<maxnctt? "Particles outside the slab + halo columns of this rank (a divide without the slab exchange?).": "The number of cells is too big for the allocated cell table."));
  // DivideFull bookkeeping (S/JCellDivCpuSingle.cpp:299-305); this implementation always sorts every particle,
  // which yields the same order as the reference's fluid-only sort because the sort is stable.
  bool same=divok;
  for(int i=0;i<3;i++)if(lastcellmin[i]!=dp.cellmin[i] || lastcellmax[i]!=dp.cellmax[i])same=false;
  dividefull=!same;
  for(int i=0;i<3;i++){ lastcellmin[i]=dp.cellmin[i]; lastcellmax[i]=dp.cellmax[i]; }
  divok=true; celltableok=true; mgownok=false;
  npfout_last=dp.npfout; npout_total+=dp.npfout;
  np=dp.npfinal; npb=dp.npbfinal; npbok=npb-dp.npbignore;
  boundchanged=false; forcesok=false;
  if(keepout && dp.npfout){
    // excluded fluid particles sit right behind the kept ones in the sorted arrays (boxes: ..., BoundOut, FluidOut, ...)
    const size_t o=outidp.size(), n=dp.npfout, first=size_t(dp.npfinal)+dp.npbout;
    outidp.resize(o+n); outcode.resize(o+n); outpos.resize((o+n)*3); outvelrhop.resize((o+n)*4);
    CUCHECK(cudaMemcpyAsync(outidp.data()+o,idp[cur]+first,n*4,cudaMemcpyDeviceToHost,stm));
    CUCHECK(cudaMemcpyAsync(outcode.data()+o,code[cur]+first,n*sizeof(typecode),cudaMemcpyDeviceToHost,stm));
    CUCHECK(cudaMemcpyAsync(outpos.data()+o*3,pos[cur]+first*3,n*24,cudaMemcpyDeviceToHost,stm));
    CUCHECK(cudaMemcpyAsync(outvelrhop.data()+o*4,velrhop[cur]+first,n*16,cudaMemcpyDeviceToHost,stm));
    CUCHECK(cudaStreamSynchronize(stm));
  }
  if(dp.npbout)throw EngineError(DSPH_ERR_PHYSICS,"Boundary particles were excluded from the domain (AbortBoundOut, S/JSph.cpp:2477).");
  AccumTimers(true);
  if(info)LastDivInfo(info);
}

void Engine::LastDivInfo(dsph_divinfo *info)const{
  {
    memset(info,0,sizeof(*info));
    info->np=np; info->npb=npb; info->npbok=npbok; info->npbper=npbper; info->npfper=npfper;
    info->npbout=dp.npbout; info->npfout=dp.npfout; info->npbignore=dp.npbignore; info->npboutignore=dp.npboutignore; info->npfoutignore=dp.npfoutignore;
    info->ncx=dp.ncx; info->ncy=dp.ncy; info->ncz=dp.ncz; info->nct=dp.nct; info->nctt=dp.nctt;
    for(int i=0;i<3;i++){ info->cellmin[i]=dp.cellmin[i]; info->cellmax[i]=dp.cellmax[i]; info->domcells[i]=k.domcells[i]; info->domposmin[i]=k.domposmin[i]; }
    info->domcellcode=k.domcellcode; info->dividefull=dividefull? 1: 0; info->scell=k.scell;
  }
}

void Engine::AccumTimers(bool divide){
  float ms=0;
  if(divide){
    cudaEventElapsedTime(&ms,ev[0],ev[1]); tm.su_periodic+=ms;
    cudaEventElapsedTime(&ms,ev[1],ev[2]); tm.nl_limits+=ms;
    cudaEventElapsedTime(&ms,ev[2],ev[3]); tm.nl_presort+=ms;
    cudaEventElapsedTime(&ms,ev[3],ev[4]); tm.nl_radixsort+=ms;
    cudaEventElapsedTime(&ms,ev[4],ev[5]); tm.nl_cellbegin+=ms;
    cudaEventElapsedTime(&ms,ev[5],ev[6]); tm.nl_sortdata+=ms;
    cudaEventElapsedTime(&ms,ev[6],ev[7]); tm.nl_outcheck+=ms;
    cudaEventElapsedTime(&ms,ev[0],ev[7]); tm.total+=ms;
  }
}

void Engine::GetBeginCell(unsigned *h,unsigned capacity){
  if(capacity<dp.nctt)throw EngineError(DSPH_ERR_ARG,"begincell capacity too small.");
  CUCHECK(cudaMemcpyAsync(h,begincell,size_t(dp.nctt)*4,cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaStreamSynchronize(stm));
}

// ------------------------------------------------------------------------------------------------
// Interaction_Forces
// ------------------------------------------------------------------------------------------------
ForceParams Engine::MakeForceParams(){
  ForceParams t; memset(&t,0,sizeof(t));
  t.np=np; t.npb=npb; t.npbok=npbok; t.skipghosts=(mgpu? 1: 0);
  t.ncx=int(dp.ncx); t.ncy=int(dp.ncy); t.ncz=int(dp.ncz); t.nsheet=int(dp.nsheet);
  t.cellfluid=dp.nct+1;
  for(int i=0;i<3;i++)t.cellzero[i]=dp.cellmin[i];
  t.begincell=begincell; t.dcell=dcell[cur]; t.posp=posp; t.posh=posh; t.velrhop=velrhop[cur]; t.code=code[cur];
  t.spstau=spstau[cur]; t.spsgradvel=spsgradvel; t.shiftposfs=shiftposfs; t.acear=acear; t.fmax=fmax;
  t.viscoeta=viscoeta; t.nntau=nntau;
  return t;
}

int Engine::EffectiveForceVariant()const{
  if(cfg.nn_on)return (ringbase_ok && RingVariant()!=4 && RingVariant()!=1 && k.scelldiv==1? RingVariant(): 1);
  ForceParams t; memset(&t,0,sizeof(t)); t.np=np;
  return EffectiveFluidVariant(RingVariant(),ringbase_ok,t,dc);
}

void Engine::Forces(int interstep,dsph_forceres *res){
  ForcesLaunch(interstep);
  ForcesFinish(res);
}

// kernels only: the three maxima stay on the device (multi-GPU: all-reduce(max) on them before ForcesFinish)
void Engine::ForcesLaunch(int interstep){
  (void)interstep;
  CUCHECK(cudaSetDevice(device));
  if(!divok)throw EngineError(DSPH_ERR_STATE,"dsph_forces_run needs a cell division first (dsph_celldiv_run).");
  NvtxBegin("CF-PreForces");
  CUCHECK(cudaEventRecord(ev[8],stm));
  CUCHECK(cudaMemsetAsync(fmax,0,sizeof(ForceMax),stm));
  // boundary correction for mDBC (S/JSphCpuSingle.cpp:525: before every force evaluation, the corrector one only with MDBCCorrector)
  if(mdbc && (mdbccorrector || interstep!=DSPH_INTERSTEP_SYMCORRECTOR))MdbcCorrection();
  NvtxNext("CF-Forces");
  CUCHECK(cudaEventRecord(ev[9],stm));
  ForceParams t=MakeForceParams();
  const unsigned npf=np-npb;
  const bool shift=(cfg.shiftmode!=0);
  if(mgpu && npf && !cfg.nn_on && EffectiveForceVariant()==10){ MgBuildOwnList(); t.ownlist=mgown; t.nown=mgscratch[1]+8; }
  if(npf && cfg.nn_on){
    const unsigned nl=LaunchForcesNN((ringbase_ok && RingVariant()!=4? RingVariant(): 1),cfg.tkernel,cfg.tvisco,cfg.tdensity,shift,t,dc,dph,stm);
    if(!nl)throw EngineError(DSPH_ERR_ARG,"Kernel / viscosity / DDT choice is not valid for the multiphase formulation.");
    launches+=nl;
  }
  else if(npf){
    if(cfg.tkernel!=DSPH_KERNEL_WENDLAND && cfg.tkernel!=DSPH_KERNEL_CUBIC)throw EngineError(DSPH_ERR_ARG,"Kernel choice is not valid.");
    if(cfg.tvisco!=DSPH_VISCO_ARTIFICIAL && cfg.tvisco!=DSPH_VISCO_LAMINARSPS)throw EngineError(DSPH_ERR_ARG,"Viscosity treatment is not valid.");
    if(!LaunchForcesFluid(RingVariant(),ringbase_ok,cfg.tkernel,cfg.tvisco,cfg.tdensity,shift,t,dc,stm))throw EngineError(DSPH_ERR_ARG,"Density Diffusion Term mode is not valid.");
    launches++;
    if(sps){ LaunchSpsTau(npf,npb,velrhop[cur],spsgradvel,spstau[cur],dc,stm); launches++; }
  }
  CUCHECK(cudaEventRecord(ev[13],stm));
  if(npb){ LaunchForcesBound(forcevariant,cfg.tkernel,numsm,t,dc,stm); launches++; }
  CUCHECK(cudaEventRecord(ev[10],stm));
  NvtxEnd();
  forceslaunched=true;
}

void Engine::ForcesFinish(dsph_forceres *res){
  if(!forceslaunched)throw EngineError(DSPH_ERR_STATE,"dsph_forces_finish needs dsph_forces_launch first.");
  forceslaunched=false;
  SmallCopy(&hpind->fmax,fmax,sizeof(ForceMax),stm);
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
  { float ms=0; cudaEventElapsedTime(&ms,ev[8],ev[9]); tm.cf_preforces+=ms; cudaEventElapsedTime(&ms,ev[9],ev[10]); tm.cf_forces+=ms; tm.total+=ms; lastforcems=ms;
    cudaEventElapsedTime(&ms,ev[9],ev[13]); lastfluidms=ms; fluidms_sum+=ms; fluid_launches++; }
  const ForceMax &m=hpin->fmax;
  float f; unsigned b;
  b=m.viscdt_bits; memcpy(&f,&b,4); viscdtmax=f;
  b=m.ace2_bits; memcpy(&f,&b,4); acemax=sqrt(double(f));
  b=m.vel2_bits; memcpy(&f,&b,4); velmax=double(sqrtf(f));
  b=m.visceta_bits; memcpy(&f,&b,4); viscetadtmax=f;
  forcesok=true;
  if(res)FillForceRes(res);
}

double Engine::DtVariable(double *pdt1,double *pdt2,double *pdt3)const{
  // JSphGpu::DtVariable S/JSphGpu.cpp:984-1009 (== JSphCpu::DtVariable S/JSphCpu.cpp:1614-1639)
  const double dt1=(acemax? (sqrt(double(cfg.kernelh)/acemax)): DBL_MAX);
  const double dt2=double(cfg.kernelh)/(std::max(k.cs0,velmax*10.)+double(cfg.kernelh)*viscdtmax);
  // viscous time step of the multiphase fork (M/JSphCpu.cpp:1687-1689, M/JSphGpu.cpp:1075-1078)
  const double dt3=(cfg.nn_on && viscetadtmax*dph.lamda>0? double(cfg.kernelh*cfg.kernelh)/double(viscetadtmax*dph.lamda): DBL_MAX);
  if(pdt3)*pdt3=dt3;
  double dt=cfg.cflnumber*std::min(dt3,std::min(dt1,dt2));
  if(cfg.dtfixed_on)dt=cfg.dtfixed;
  if(dt!=dt || std::isinf(dt))throw EngineError(DSPH_ERR_PHYSICS,"The computed Dt is NaN or infinity.");
  if(dt<k.dtmin)dt=k.dtmin;
  if(pdt1)*pdt1=dt1; if(pdt2)*pdt2=dt2;
  return dt;
}
void Engine::FillForceRes(dsph_forceres *res)const{
  memset(res,0,sizeof(*res));
  res->viscdtmax=viscdtmax; res->acemax=float(acemax); res->velmax=float(velmax);
  res->viscetadtmax=viscetadtmax;
  res->dt=DtVariable(&res->dt1,&res->dt2,&res->dt3);
}

void Engine::ForcesDownload(unsigned first,unsigned count,float *ace,float *ar){
  if(!forcesok)throw EngineError(DSPH_ERR_STATE,"Forces are not available.");
  if(first+count>np)throw EngineError(DSPH_ERR_ARG,"Range out of particles.");
  std::vector<float> h(size_t(count)*4);
  CUCHECK(cudaMemcpyAsync(h.data(),acear+first,size_t(count)*16,cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaStreamSynchronize(stm));
  for(unsigned i=0;i<count;i++){
    if(ace){ ace[3*size_t(i)]=h[4*size_t(i)]; ace[3*size_t(i)+1]=h[4*size_t(i)+1]; ace[3*size_t(i)+2]=h[4*size_t(i)+2]; }
    if(ar)ar[i]=h[4*size_t(i)+3];
  }
}

// Probes: pressure / velocity / kernel sum at n points from the current cell list (measure.cuh)
void Engine::MeasurePoints(unsigned n,const double *hpoints,float *hpress,float *hvel,float *hkcorr){
  if(!divok)throw EngineError(DSPH_ERR_STATE,"dsph_measure_points needs a cell division first (dsph_celldiv_run).");
  if(!n)return;
  // persistent scratch (grown on demand): no allocation per call
  const size_t need=size_t(n)*24+size_t(n)*20;
  if(measurecap<need){ if(measuredev)cudaFree(measuredev); measuredev=nullptr; measurecap=0; CUCHECK(cudaMalloc((void**)&measuredev,need+4096)); measurecap=need+4096; }
  double *dpts=(double*)measuredev; float *dout=(float*)(measuredev+size_t(n)*24);
  CUCHECK(cudaMemcpyAsync(dpts,hpoints,size_t(n)*24,cudaMemcpyHostToDevice,stm));
  MeasureArgs m; m.npoints=n; m.points=dpts; m.pos=pos[cur]; m.press=dout; m.vel=dout+n; m.kcorr=dout+size_t(n)*4;
  const ForceParams t=MakeForceParams();
  const unsigned nb=(n+127)/128;
  const bool wend=(cfg.tkernel==DSPH_KERNEL_WENDLAND);
  if(cfg.nn_on){ if(wend)ker_measure_points<DSPH_KERNEL_WENDLAND,true><<<nb,128,0,stm>>>(m,t,dc,dph); else ker_measure_points<DSPH_KERNEL_CUBIC,true><<<nb,128,0,stm>>>(m,t,dc,dph); }
  else{ if(wend)ker_measure_points<DSPH_KERNEL_WENDLAND,false><<<nb,128,0,stm>>>(m,t,dc,dph); else ker_measure_points<DSPH_KERNEL_CUBIC,false><<<nb,128,0,stm>>>(m,t,dc,dph); }
  launches++;
  if(hpress)CUCHECK(cudaMemcpyAsync(hpress,dout,size_t(n)*4,cudaMemcpyDeviceToHost,stm));
  if(hvel)CUCHECK(cudaMemcpyAsync(hvel,dout+n,size_t(n)*12,cudaMemcpyDeviceToHost,stm));
  if(hkcorr)CUCHECK(cudaMemcpyAsync(hkcorr,dout+size_t(n)*4,size_t(n)*4,cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
}

void Engine::Pips(unsigned long long *fr,unsigned long long *fc,unsigned long long *br,unsigned long long *bc){
  if(!divok)throw EngineError(DSPH_ERR_STATE,"dsph_pips needs a cell division first.");
  CUCHECK(cudaMemsetAsync(fmax,0,sizeof(ForceMax),stm));
  const ForceParams t=MakeForceParams();
  if(np>npb){ LaunchPips(t,dc,npb,np-npb,1,stm); launches++; }
  if(npbok){ LaunchPips(t,dc,0,npbok,0,stm); launches++; }
  SmallCopy(&hpind->fmax,fmax,sizeof(ForceMax),stm);
  CUCHECK(cudaStreamSynchronize(stm));
  if(fr)*fr=hpin->fmax.pips_fluid_real; if(fc)*fc=hpin->fmax.pips_fluid_checked;
  if(br)*br=hpin->fmax.pips_bound_real; if(bc)*bc=hpin->fmax.pips_bound_checked;
  forcesok=false;
}

// ------------------------------------------------------------------------------------------------
// ComputeStep
// ------------------------------------------------------------------------------------------------
StepArrays Engine::MakeStepArrays(){
  StepArrays a; memset(&a,0,sizeof(a));
  a.np=np; a.npb=npb; a.pos=pos[cur]; a.dcell=dcell[cur]; a.code=code[cur]; a.velrhop=velrhop[cur];
  a.velrhopm1=velrhopm1[cur]; a.pospre=pospre[cur]; a.velrhoppre=velrhoppre[cur]; a.acear=acear;
  a.shiftposfs=(cfg.shiftmode? shiftposfs: nullptr);
  return a;
}
void Engine::ShiftCoefs(double dt,double &coefumagn,double &coeftfs,float &maxdist)const{
  coefumagn=dt*cfg.shiftcoef*cfg.kernelh;
  coeftfs=(cfg.simulate2d? 2.0: 3.0)-cfg.shifttfs;
  maxdist=float(cfg.dp*0.1);
}
void Engine::StepSymPre(double dt){
  if(!forcesok)throw EngineError(DSPH_ERR_STATE,"dsph_step_sympre needs dsph_forces_run first.");
  if(cfg.tstep!=DSPH_STEP_SYMPLECTIC)throw EngineError(DSPH_ERR_STATE,"The context is not configured for Symplectic.");
  NvtxBegin("SU-ComputeStep");
  CUCHECK(cudaEventRecord(ev[11],stm));
  StepArrays a=MakeStepArrays();
  ker_step_sympre<<<(np+STEP_THREADS-1)/STEP_THREADS,STEP_THREADS,0,stm>>>(a,dt,dc); launches++;
  CUCHECK(cudaEventRecord(ev[12],stm));
  NvtxEnd();
  haspre=true; forcesok=false; divok=false; steppending=true;
}
void Engine::StepSymCor(double dt){
  if(!forcesok)throw EngineError(DSPH_ERR_STATE,"dsph_step_symcor needs dsph_forces_run first.");
  if(!haspre)throw EngineError(DSPH_ERR_STATE,"dsph_step_symcor needs dsph_step_sympre first.");
  NvtxBegin("SU-ComputeStep");
  CUCHECK(cudaEventRecord(ev[11],stm));
  StepArrays a=MakeStepArrays();
  double cu,ct; float md; ShiftCoefs(dt,cu,ct,md);
  ker_step_symcor<<<(np+STEP_THREADS-1)/STEP_THREADS,STEP_THREADS,0,stm>>>(a,dt,cu,ct,md,dc); launches++;
  CUCHECK(cudaEventRecord(ev[12],stm));
  NvtxEnd();
  haspre=false; forcesok=false; divok=false; steppending=true;
}
void Engine::StepVerlet(double dt){
  if(!forcesok)throw EngineError(DSPH_ERR_STATE,"dsph_step_verlet needs dsph_forces_run first.");
  if(cfg.tstep!=DSPH_STEP_VERLET)throw EngineError(DSPH_ERR_STATE,"The context is not configured for Verlet.");
  NvtxBegin("SU-ComputeStep");
  CUCHECK(cudaEventRecord(ev[11],stm));
  verletstep++;
  const int euler=(verletstep<cfg.verletsteps? 0: 1);
  StepArrays a=MakeStepArrays();
  double cu,ct; float md; ShiftCoefs(dt,cu,ct,md);
  ker_step_verlet<<<(np+STEP_THREADS-1)/STEP_THREADS,STEP_THREADS,0,stm>>>(a,dt,euler,cu,ct,md,dc); launches++;
  std::swap(velrhop[cur],velrhopm1[cur]);
  if(euler)verletstep=0;
  CUCHECK(cudaEventRecord(ev[12],stm));
  NvtxEnd();
  forcesok=false; divok=false; steppending=true;
}
void Engine::FlushStepTimer(){
  if(steppending){
    CUCHECK(cudaEventSynchronize(ev[12]));
    float ms=0; cudaEventElapsedTime(&ms,ev[11],ev[12]); tm.su_computestep+=ms; tm.total+=ms; steppending=false;
  }
}

void Engine::Step(double *stepdt,dsph_forceres *res){
  double dt;
  bool moving=false;
  if(cfg.tstep==DSPH_STEP_SYMPLECTIC){
    dt=symplecticdtpre;
    if(motionset)moving=motionset->Process(gaugetime,dt,motiondata);          // CalcMotion(dt), S/JSphCpuSingle.cpp:697
    Forces(DSPH_INTERSTEP_SYMPREDICTOR,nullptr);
    const double ddt_p=DtVariable(nullptr,nullptr);
    StepSymPre(dt);
    CellDivide(true,nullptr); FlushStepTimer();
    Forces(DSPH_INTERSTEP_SYMCORRECTOR,res);
    const double ddt_c=DtVariable(nullptr,nullptr);
    StepSymCor(dt);
    symplecticdtpre=std::min(ddt_p,ddt_c);
  }
  else{
    Forces(DSPH_INTERSTEP_VERLET,res);
    dt=DtVariable(nullptr,nullptr);
    if(motionset)moving=motionset->Process(gaugetime,dt,motiondata);          // S/JSphCpuSingle.cpp:677
    StepVerlet(dt);
  }
  if(!gauges.empty())GaugesRun(gaugetime+dt,false,nullptr);                    // RunGaugeSystem(TimeStep+stepdt), S/JSphCpuSingle.cpp:1095
  if(moving)MotionRun(motiondata.data(),unsigned(motiondata.size()),dt);       // RunMotion(stepdt), :1096
  gaugetime+=dt;
  CellDivide(true,nullptr); FlushStepTimer();
  if(stepdt)*stepdt=dt;
}

// ------------------------------------------------------------------------------------------------
// x-slab multi-GPU: pack / unpack of migrants and halo copies; transport is done by the caller (NCCL)
// ------------------------------------------------------------------------------------------------
// Cells a divide can meet: after the slab exchange every live particle of a rank lies in its columns + one halo width per side,
// so the sort needs the key bits of that block only (N=8, DB(160 M): 3 radix passes instead of 4) and the cell table is cleared,
// scanned and reduced over that block.  ker_div_prepare raises DivParams.error when the particles need more.
void Engine::UpdateCellBound(){
  effnctt=maxnctt;
  if(mgpu && maxnctt){
    const unsigned long long cols=std::min<unsigned long long>(k.domcells[0],(unsigned long long)(mgxhi-mgxlo)+2ull*MgHalo()+2ull);
    const unsigned long long n=cols*k.domcells[1]*k.domcells[2]*2ull+6ull;
    effnctt=unsigned(std::min<unsigned long long>(n,maxnctt));
  }
  effkeybits=KeyBits(effnctt);
}
void Engine::MgConfig(int rank,int nranks,unsigned xlo,unsigned xhi){
  if(cfg.perix && nranks>1 && (k.perixinc[1]!=0 || k.perixinc[2]!=0))throw EngineError(DSPH_ERR_ARG,"x-slabs with a sheared periodic face along x are not supported.");
  if(nranks<1 || rank<0 || rank>=nranks || xlo>=xhi || xhi>k.domcells[0])throw EngineError(DSPH_ERR_ARG,"Invalid slab configuration.");
  if(nranks>1 && mdbc)throw EngineError(DSPH_ERR_ARG,"mDBC cannot be combined with x-slabs (the halo records carry no normal vectors).");
  if(nranks>1 && xhi-xlo<2u*MgHalo()+1u)throw EngineError(DSPH_ERR_ARG,"The slab is narrower than 2 halo widths + 1 cell.");
  mgpu=(nranks>1); mgrank=rank; mgnranks=nranks; mgxlo=xlo; mgxhi=xhi;
  mgring=(mgpu && cfg.perix!=0);     // periodic along x: rank 0 and rank nranks-1 are neighbours through the periodic face
  dc.ghosts=(k.periactive||mgpu)? 1: 0;
  UpdateCellBound();
  if(!mgpu)for(int i=0;i<3;i++)DevFree(mgscratch[i]);
  else if(!mgscratch[0]){                                            // (re-balancing calls MgConfig again: keep the scratch)
    DevAlloc(mgscratch[0],MG_CLASSES*((size_t(cap)+MG_BLOCK-1)/MG_BLOCK)+8);   // per-block class counts -> offsets
    DevAlloc(mgscratch[1],size_t(2*k.domcells[0]+16));               // class totals; column histograms (fluid, boundary)
    DevAlloc(mgscratch[2],size_t(cap)/4+2);                          // one flag byte per particle
  }
}
// owned fluid particles of the current divide in index order (mgpu.cuh); their number stays on the device (mgscratch[1][8])
void Engine::MgBuildOwnList(){
  if(mgownok)return;
  const unsigned npf=np-npb;
  if(mgowncap<npf){ if(mgown)cudaFree(mgown); mgown=nullptr; mgowncap=0; CUCHECK(cudaMalloc((void**)&mgown,size_t(cap)*sizeof(unsigned))); mgowncap=cap; }
  const unsigned nb=(npf+MG_BLOCK-1)/MG_BLOCK;
  ker_mg_own_count<<<nb,MG_BLOCK,0,stm>>>(np,npb,code[cur],mgscratch[0]);
  ker_mg_scanblocks<<<1,1024,0,stm>>>(nb,mgscratch[0],mgscratch[1]+8);
  ker_mg_own_write<<<nb,MG_BLOCK,0,stm>>>(np,npb,code[cur],mgscratch[0],mgown);
  launches+=3;
  mgownok=true;
}
void Engine::MgColumnCounts(unsigned *fluid,unsigned *bound,unsigned ncol){
  if(!mgpu)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_column_counts needs dsph_mgpu_config with more than one rank.");
  if(ncol!=k.domcells[0])throw EngineError(DSPH_ERR_ARG,"ncol must be the number of cell columns of the domain (dsph_domain_info).");
  EnsureDcell();
  unsigned *d=mgscratch[1]+16;
  CUCHECK(cudaMemsetAsync(d,0,size_t(2*ncol)*4,stm));
  if(np){ ker_mg_column_counts<<<(np+255)/256,256,0,stm>>>(np,dcell[cur],code[cur],k.domcellcode,ncol,d,d+ncol); launches++; }
  CUCHECK(cudaMemcpyAsync(fluid,d,size_t(ncol)*4,cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaMemcpyAsync(bound,d+ncol,size_t(ncol)*4,cudaMemcpyDeviceToHost,stm));
  CUCHECK(cudaStreamSynchronize(stm));
}
// slab geometry of the kernels (ring of ranks when the case is periodic along x)
void Engine::MgFillArgs(MgArgs &a){
  memset(&a,0,sizeof(a));
  a.np=np; a.npb=npb; a.xlo=mgxlo; a.xhi=mgxhi; a.sd=MgHalo();
  a.domcellcode=k.domcellcode; a.dcell=dcell[cur]; a.code=code[cur]; a.pos=pos[cur];
  a.ring=(mgring? 1: 0);
  a.hasleft=(mgring || mgrank>0); a.hasright=(mgring || mgrank<mgnranks-1);
  if(mgring){
    a.wrapleft=(mgrank==0); a.wrapright=(mgrank==mgnranks-1);
    const double sc=dc.scell_d;
    a.bandwidth=double(MgHalo())*sc;
    a.xs0=std::max(k.domposmin[0]+double(mgxlo)*sc,k.maprealposmin[0]);
    a.xs1=std::min(k.domposmin[0]+double(mgxhi)*sc,k.maprealposmin[0]+k.maprealsize[0]);
    a.xlen=k.maprealsize[0];
    for(int i=0;i<3;i++){ a.shiftleft[i]=(a.wrapleft? -k.perixinc[i]: 0.0); a.shiftright[i]=(a.wrapright? k.perixinc[i]: 0.0); }   // perixinc.x = -length
  }
}
// classification + per-class totals on the stream, no host synchronisation: totals land in mgscratch[1][0..7]
void Engine::MgClassifyAsync(){
  if(!mgpu)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_count needs dsph_mgpu_config with more than one rank.");
  EnsureDcell();
  const unsigned nb=(np+MG_BLOCK-1)/MG_BLOCK;
  if(!np){ CUCHECK(cudaMemsetAsync(mgscratch[1],0,MG_CLASSES*sizeof(unsigned),stm)); return; }
  MgArgs a; MgFillArgs(a);
  // one pass: retire the halo copies of the previous divide (dropped by the next sort), classify, count per block
  ker_mg_classify<<<nb,MG_BLOCK,0,stm>>>(a,(unsigned char*)mgscratch[2],mgscratch[0]); launches++;
  ker_mg_scanblocks<<<MG_CLASSES,1024,0,stm>>>(nb,mgscratch[0],mgscratch[1]); launches++;
}
void Engine::MgCount(unsigned counts[8]){
  for(int g=0;g<MG_CLASSES;g++)mgcounts[g]=counts[g]=0;
  MgClassifyAsync();
  mgcounted=true;
  if(!np)return;
  SmallCopy(hpind->mgtot,mgscratch[1],MG_CLASSES*sizeof(unsigned),stm);
  CUCHECK(cudaStreamSynchronize(stm));
  for(int g=0;g<MG_CLASSES;g++)mgcounts[g]=counts[g]=hpin->mgtot[g];
}
// pack kernel on the stream (counts of the classification in mgcounts[])
void Engine::MgPackAsync(void *sendleft,void *sendright){
  if(!mgcounted)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_pack needs dsph_mgpu_count first.");
  mgcounted=false;
  const unsigned nb=(np+MG_BLOCK-1)/MG_BLOCK;
  MgArgs a; MgFillArgs(a);
  const unsigned *counts=mgcounts;
  void *sendbuf[2]={sendleft,sendright};
  MgPackArrays sp; sp.idp=idp[cur]; sp.code=code[cur]; sp.pos=pos[cur]; sp.velrhop=velrhop[cur];
  sp.dcell=dcell[cur]; for(int i=0;i<3;i++)sp.domposmin[i]=dc.domposmin[i]; sp.scell=dc.scell_d;
  sp.pospre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? pospre[cur]: nullptr);
  sp.aux=(cfg.tstep==DSPH_STEP_SYMPLECTIC? (haspre? velrhoppre[cur]: nullptr): velrhopm1[cur]);
  sp.spstau=(sps? spstau[cur]: nullptr); sp.recbytes=MgRecordBytes();
  MgBases bases; unsigned total=0;
  for(int d=0;d<2;d++){
    unsigned base=0;
    for(int kd=0;kd<4;kd++){ bases.v[4*d+kd]=base; base+=counts[4*d+kd]; }
    if(base && !sendbuf[d])throw EngineError(DSPH_ERR_ARG,"Null send buffer.");
    total+=base;
  }
  if(total){ ker_mg_pack<false><<<nb,MG_BLOCK,0,stm>>>(a,(const unsigned char*)mgscratch[2],mgscratch[0],bases,sp,(char*)sendleft,(char*)sendright,MgDirect()); launches++; }
}
// direct transport: the pack kernel writes into the neighbours' receive buffers (peer memory) with the class bases taken from the
// device-side totals; token = 10 words {to left: 4 totals, used; to right: 4 totals, used}
void Engine::MgPackDirectAsync(void *peerleft,void *peerright,unsigned capleft,unsigned capright,unsigned *token){
  const unsigned nb=std::max(1u,(np+MG_BLOCK-1)/MG_BLOCK);
  MgArgs a; MgFillArgs(a);
  MgPackArrays sp; sp.idp=idp[cur]; sp.code=code[cur]; sp.pos=pos[cur]; sp.velrhop=velrhop[cur];
  sp.dcell=dcell[cur]; for(int i=0;i<3;i++)sp.domposmin[i]=dc.domposmin[i]; sp.scell=dc.scell_d;
  sp.pospre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? pospre[cur]: nullptr);
  sp.aux=(cfg.tstep==DSPH_STEP_SYMPLECTIC? (haspre? velrhoppre[cur]: nullptr): velrhopm1[cur]);
  sp.spstau=(sps? spstau[cur]: nullptr); sp.recbytes=MgRecordBytes();
  MgDirect dr; dr.totals=mgscratch[1]; dr.capleft=capleft; dr.capright=capright; dr.token=token;
  MgBases bases; memset(&bases,0,sizeof(bases));
  ker_mg_pack<true><<<nb,MG_BLOCK,0,stm>>>(a,(const unsigned char*)mgscratch[2],mgscratch[0],bases,sp,(char*)peerleft,(char*)peerright,dr); launches++;
}
void Engine::MgPushAsync(const void *srcleft,void *peerleft,const void *srcright,void *peerright,const unsigned *token){
  ker_mg_push<<<numsm*8,256,0,stm>>>((const uint4*)srcleft,(uint4*)peerleft,(const uint4*)srcright,(uint4*)peerright,token,MgRecordBytes()); launches++;
}
void Engine::MgPack(void *sendleft,void *sendright){
  MgPackAsync(sendleft,sendright);
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
}
void Engine::MgUnpackAsync(const unsigned counts[8],const void *recvleft,const void *recvright){
  if(!mgpu)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_unpack needs dsph_mgpu_config with more than one rank.");
  // per side: boundary block (ghosts, migrants) then fluid block (migrants, ghosts)
  const unsigned bgl=counts[0]+counts[1], fl=counts[2]+counts[3], bgr=counts[4]+counts[5], fr=counts[6]+counts[7];
  const unsigned total=bgl+fl+bgr+fr;
  if(size_t(np)+total>cap)throw EngineError(DSPH_ERR_STATE,"Received particles exceed the capacity of the context.");
  MgUnpackArrays d; d.idp=idp[cur]; d.code=code[cur]; d.dcell=dcell[cur]; d.pos=pos[cur]; d.velrhop=velrhop[cur];
  d.pospre=(cfg.tstep==DSPH_STEP_SYMPLECTIC && haspre? pospre[cur]: nullptr);
  d.aux=(cfg.tstep==DSPH_STEP_SYMPLECTIC? (haspre? velrhoppre[cur]: nullptr): velrhopm1[cur]);
  d.spstau=(sps? spstau[cur]: nullptr); d.recbytes=MgRecordBytes();
  const int B=256;
  unsigned dst=np;
  auto up=[&](unsigned n,const char *src){ if(n){ ker_mg_unpack<<<(n+B-1)/B,B,0,stm>>>(n,src,dst,d,dc); launches++; dst+=n; } };
  up(bgl,(const char*)recvleft);
  up(bgr,(const char*)recvright);
  up(fl,(const char*)recvleft+size_t(bgl)*d.recbytes);
  up(fr,(const char*)recvright+size_t(bgr)*d.recbytes);
  npbper=bgl+bgr; npfper=fl+fr; np+=total;
  boundchanged=true; celltableok=false;
}
void Engine::MgUnpack(const unsigned counts[8],const void *recvleft,const void *recvright){
  MgUnpackAsync(counts,recvleft,recvright);
  CUCHECK(cudaStreamSynchronize(stm));
  CUCHECK(cudaGetLastError());
}
// device ms of this rank's own kernels since the last call (the event timers of the context; waiting for neighbours is not in them)
double Engine::MgMeasuredCost(){
  FlushStepTimer();
  const double c=tm.nl_limits+tm.nl_presort+tm.nl_radixsort+tm.nl_cellbegin+tm.nl_sortdata+tm.nl_outcheck+tm.cf_preforces+tm.cf_forces+tm.su_computestep+tm.su_periodic+tm.su_shifting;
  const double r=c-mgcostmark; mgcostmark=c;
  return r;
}

void Engine::GetTimers(dsph_timers *t,bool reset){
  FlushStepTimer();
  tm.kernel_launches=launches;
  if(t)*t=tm;
  if(reset){ memset(&tm,0,sizeof(tm)); launches=0; mgcostmark=0; }
}

// ------------------------------------------------------------------------------------------------
// moving boundaries (kernels: motion.cuh): JSphGpu::RunMotion (S/JSphGpu.cpp:1035-1075)
// ------------------------------------------------------------------------------------------------
// The host (the reference's JDsMotion, or host/motion.cpp of the stand-alone driver) evaluated the motion laws for [TimeStep,TimeStep+dt);
// this applies them to the particles, after ComputeStep and before the divide like JSphGpuSingle::Run (S/JSphGpuSingle.cpp:859).
void Engine::MotionRun(const dsph_motiondata *objs,unsigned nobj,double stepdt){
  CUCHECK(cudaSetDevice(device));
  if(!cfg.casenmoving)throw EngineError(DSPH_ERR_STATE,"The case has no moving boundary particles.");
  if(nobj && !objs)throw EngineError(DSPH_ERR_ARG,"Null argument.");
  bool any=false;
  for(unsigned i=0;i<nobj;i++){
    const dsph_motiondata &m=objs[i];
    if(m.type==DSPH_MOTION_NONE)continue;
    if(m.type!=DSPH_MOTION_LINEAR && m.type!=DSPH_MOTION_MATRIX)throw EngineError(DSPH_ERR_ARG,"Motion type is invalid.");
    if(m.idbegin<cfg.casenfixed || size_t(m.idbegin)+m.count>size_t(cfg.casenfixed)+cfg.casenmoving)throw EngineError(DSPH_ERR_ARG,"The ids of a moving object are outside the moving range of the case.");
    if(m.type==DSPH_MOTION_MATRIX && !(stepdt>0))throw EngineError(DSPH_ERR_ARG,"A matrix motion needs the dt of the step.");
    any=true;
  }
  if(!any)return;
  NvtxBegin("SU-Motion");
  EnsureDcell();
  if(!ridpmove)CUCHECK(cudaMalloc((void**)&ridpmove,size_t(cfg.casenmoving)*sizeof(unsigned)));
  CUCHECK(cudaMemsetAsync(ridpmove,0xff,size_t(cfg.casenmoving)*sizeof(unsigned),stm));
  if(npb){ ker_motion_ridp<<<(npb+255)/256,256,0,stm>>>(npb,cfg.casenfixed,cfg.casenfixed+cfg.casenmoving,(k.periactive||mgpu)? 1: 0,idp[cur],code[cur],ridpmove); launches++; }
  MotionArrays a; a.pos=pos[cur]; a.dcell=dcell[cur]; a.velrhop=velrhop[cur]; a.code=code[cur]; a.boundnormal=(mdbc? boundnormal[cur]: nullptr);
  for(unsigned i=0;i<nobj;i++){
    const dsph_motiondata &m=objs[i];
    if(m.type==DSPH_MOTION_NONE || !m.count)continue;
    const unsigned ini=m.idbegin-cfg.casenfixed;
    if(m.type==DSPH_MOTION_LINEAR)
      ker_motion_linear<<<(m.count+255)/256,256,0,stm>>>(m.count,ini,m.linmov[0],m.linmov[1],m.linmov[2],float(m.linvel[0]),float(m.linvel[1]),float(m.linvel[2]),ridpmove,a,dc);
    else{
      Mat4d mt; memcpy(mt.a,m.matmov,sizeof(mt.a));
      ker_motion_matrix<<<(m.count+255)/256,256,0,stm>>>(m.count,ini,mt,stepdt,ridpmove,a,dc);
    }
    launches++;
  }
  NvtxEnd();
  boundchanged=true; divok=false; celltableok=false;
}

} // namespace dsph
