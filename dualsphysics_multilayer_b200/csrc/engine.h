// engine.h -- see engine.cu.
#pragma once
#include "../../include/dsph_b200.h"
#include "devtypes.h"
#include "host/sphconst.h"
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "host/motion.h"
#include <stdexcept>

namespace dsph {

struct EngineError : public std::runtime_error {
  int code;
  EngineError(int c,const std::string &m):std::runtime_error(m),code(c){}
};

struct ForceParams; struct StepArrays; struct MgArgs;

struct HostPinned {
  DivParams dp;
  ForceMax fmax;
  unsigned limitsinit[12];
  unsigned scratch[8];
  unsigned mgtot[16];   // class totals of the slab exchange: [0..7] own (to left, to right), [8..15] announced by the neighbours
  unsigned mgtoken[20]; // direct transport: own tokens [0..9], the neighbours' [10..14] from left, [15..19] from right
  unsigned probe;
};

class Engine {
 public:
  static const int NEVENTS=16;
  int device;
  cudaStream_t stm=nullptr;
  cudaEvent_t ev[NEVENTS]={};
  HostPinned *hpin=nullptr, *hpind=nullptr;   // mapped pinned block: host view and device alias
  dsph_config cfg; SphConst k; DevCtes dc;
  unsigned cap=0, np=0, npb=0, npbok=0, npbper=0, npfper=0;
  // particle arrays (double buffered for the reorder; cur = live buffer)
  int cur=0;
  unsigned *idp[2]={}, *dcell[2]={}; typecode *code[2]={};
  double *pos[2]={}; float4 *velrhop[2]={}, *velrhopm1[2]={}, *velrhoppre[2]={}; double *pospre[2]={};
  float *spstau[2]={};
  // mDBC (mdbc.cuh): normal vectors particle -> ghost node of the boundary particles, sorted with the particles
  bool mdbc=false, mdbccorrector=false; float mdbcthreshold=0; float4 *boundnormal[2]={};
  void SetBoundNormals(const float *h,unsigned count); void GetBoundNormals(float *h,unsigned count);
  void MdbcCorrection();
  static const unsigned POSP_PAD=512;
  float4 *posp=nullptr; uint4 *posh=nullptr; float4 *acear=nullptr, *shiftposfs=nullptr; float *spsgradvel=nullptr;
  // multiphase non-Newtonian fork (src_mphase): phases + scratch of the SPH velocity-gradient approach
  DevPhases dph; float *viscoeta=nullptr, *nntau=nullptr; float viscetadtmax=0;
  bool reftail=false;   // test switch, see celldiv.cuh PressFromRhop
  bool sps=false;   // Laminar+SPS of the single-phase solver: SpsTau is particle state (sorted, duplicated, exchanged)
  // cell division
  RadixSortBuffers rs;
  unsigned *begincell=nullptr, *sufmin=nullptr, *limits=nullptr, *perilist=nullptr, *pericount=nullptr;
  DivParams *divparams=nullptr; DivParams dp={};
  unsigned maxnctt=0; int keybits=0;          // capacity of the cell table (whole domain)
  unsigned effnctt=0; int effkeybits=0;       // bound of the cells a divide can meet: the domain, or this rank's slab + halo (UpdateCellBound)
  void UpdateCellBound();
  bool boundchanged=true, blimok=false, divok=false, dividefull=false, haspre=false, forcesok=false, steppending=false;
  unsigned lastcellmin[3]={0,0,0}, lastcellmax[3]={0,0,0};
  unsigned npfout_last=0; unsigned long long npout_total=0;
  // forces
  ForceMax *fmax=nullptr;
  float viscdtmax=0; double acemax=0, velmax=0; float lastforcems=0, lastfluidms=0; double fluidms_sum=0; unsigned long long fluid_launches=0;
  bool ringbase_ok=true;  // the shared-window assumption of the ring kernels holds on this device/driver (probed at creation)
  // fluid force kernel (forces_api.h): 10 = packed-half culling + hit-mask ring (production), 9 = deferred pair ring with fp32 candidate
  // test (bit-identical results to 10), 4 = plain loop with the division-free body (any cell mode), 1 = IEEE loop in the reference's
  // structure and pair order (cross-check).  Multiphase: 1 = plain IEEE loops, anything else = ring kernels.
  int forcevariant=10;
  // cells denser than the padding of posp/posh allows (POSP_PAD) must not run the unclamped chunk loops of the ring kernels
  int RingVariant()const{ return (dp.maxcellpop+64u<=POSP_PAD || forcevariant==1)? forcevariant: 4; }
  int EffectiveForceVariant()const;
  // time stepping
  double symplecticdtpre=0; int verletstep=0;
  dsph_timers tm={}; unsigned long long launches=0;

  Engine(int device,const dsph_config &c,unsigned capacity);
  ~Engine();
  void Configure(const dsph_config &c);
  void Reconfigure(const dsph_config &c);   // dsph_constants_upload on a live context: constants only, layout-changing fields are rejected
  void Allocate(unsigned capacity);
  void FreeAll();
  void Upload(unsigned np,unsigned npb,const unsigned *idp,const typecode *code,const double *pos,const float *velrhop);
  bool dcellpending=false;   // set by Upload: the first cell hash runs with the first consumer (EnsureDcell)
  void EnsureDcell();
  void SetVelrhopM1(const float *h);
  void SetSpsTau(const float *h);
  void Download(unsigned first,unsigned count,unsigned *idp,typecode *code,unsigned *dcell,double *pos,float *velrhop);
  // PART snapshots (SURVEY.md 8f-2, S/JSphGpuSingle.cpp:889-935 SaveData): a device-side copy of the live state on the solver's
  // stream (the only thing the time loop waits for), then D2H into pinned host buffers on a copy stream while the next steps run
  struct HostSnap { unsigned *idp=nullptr; typecode *code=nullptr; double *pos=nullptr; float *velrhop=nullptr; unsigned capacity=0, np=0; cudaEvent_t done=nullptr; };
  cudaStream_t stmcopy=nullptr; cudaEvent_t evsnap=nullptr, evsnapfree=nullptr; char *snapdev=nullptr; unsigned snapcap=0; bool snapbusy=false;
  void HostSnapAlloc(HostSnap &hs,unsigned capacity);
  void HostSnapFree(HostSnap &hs);
  void SnapshotBegin(HostSnap &hs);                 // returns at once; hs.done fires when the host buffers are complete
  void SnapshotWait(const HostSnap &hs);            // host wait for hs.done (any thread)
  // excluded fluid particles of every divide are kept on the host until the caller takes them (JDsPartsOut, S/JSph.cpp:2467-2471)
  bool keepout=false; std::vector<unsigned> outidp; std::vector<typecode> outcode; std::vector<double> outpos; std::vector<float> outvelrhop;
  void RunPeriodic();
  void RunPeriodicSlab(unsigned npb_owned,unsigned np_owned,unsigned nbrecv,unsigned nfrecv);
  void CellDivide(bool updateperiodic,dsph_divinfo *info);
  void LastDivInfo(dsph_divinfo *info)const;
  void GetBeginCell(unsigned *h,unsigned capacity);
  ForceParams MakeForceParams();
  void Forces(int interstep,dsph_forceres *res);
  void ForcesLaunch(int interstep);
  void ForcesFinish(dsph_forceres *res);
  // x-slab multi-GPU
  bool mgpu=false, mgring=false; int mgrank=0, mgnranks=1; unsigned mgxlo=0, mgxhi=0;
  void MgFillArgs(MgArgs &a);
  int MgLeft()const{ return mgring? (mgrank+mgnranks-1)%mgnranks: (mgrank>0? mgrank-1: -1); }
  int MgRight()const{ return mgring? (mgrank+1)%mgnranks: (mgrank<mgnranks-1? mgrank+1: -1); }
  unsigned *mgscratch[3]={}; unsigned mgcounts[8]={}; bool mgcounted=false;
  unsigned *mgown=nullptr; unsigned mgowncap=0; bool mgownok=false;   // owned fluid particles of the current divide (force kernel variant 10)
  void MgBuildOwnList();
  void MgAllReduceFloats(float *dev,size_t n,bool max);
  // moving boundaries (motion.cuh, engine.cu)
  unsigned *ridpmove=nullptr;
  MotionSet *motionset=nullptr; std::vector<dsph_motiondata> motiondata;   // attached motion laws (dsph_motion_attach): evaluated inside Step / MgStep
  void MotionRun(const dsph_motiondata *objs,unsigned nobj,double stepdt);
  bool forceslaunched=false;
  void MgConfig(int rank,int nranks,unsigned xlo,unsigned xhi);
  // halo width in cells: 2h; 4h for the SPH velocity-gradient approach of the multiphase fork (its second pair pass reads eta_eff / tau of
  // the neighbours: with a double halo the first pass is complete for the inner ghost layer, so no second exchange is needed)
  unsigned MgHalo()const{ return unsigned(k.scelldiv)*(cfg.nn_on && cfg.nn_velgrad==DSPH_VELGRAD_SPH && cfg.tvisco!=DSPH_VISCO_ARTIFICIAL? 2u: 1u); }
  unsigned MgRecordBytes()const{ return sps? 112u: 88u; }   // pos 24, velrhop 16, pospre 24, velrhoppre|velrhopm1 16, idp 4, code 4 [+ SpsTau 24]
  void MgCount(unsigned counts[8]);
  void MgPushAsync(const void *srcleft,void *peerleft,const void *srcright,void *peerright,const unsigned *token);
  void MgPackDirectAsync(void *peerleft,void *peerright,unsigned capleft,unsigned capright,unsigned *token);
  void MgClassifyAsync(); void MgPackAsync(void *sendleft,void *sendright); void MgUnpackAsync(const unsigned counts[8],const void *recvleft,const void *recvright);
  double mgcostmark=0; double MgMeasuredCost();
  // native transport (mgpu_comm.cu): NCCL communicator of the slab group, exchange buffers, whole-step sequencing
  void *mgcomm=nullptr; bool mgcomm_own=false;
  char *mgsend[2]={nullptr,nullptr}, *mgrecv[2]={nullptr,nullptr}; size_t mgsendcap[2]={0,0}, mgrecvcap[2]={0,0};
  unsigned *mgcntdev=nullptr; double *mghistdev=nullptr; unsigned mghistcap=0;
  // direct transport: receive buffers of this rank [from left|right][parity], the neighbours' buffers this rank writes to [left|right][parity]
  int mgdirect=-1;                          // -1 not negotiated yet, 0 NCCL records, 1 peer memory
  char *mgdrecv[2][2]={{nullptr,nullptr},{nullptr,nullptr}}; unsigned mgdcaprec=0;
  char *mgdsend[2]={nullptr,nullptr};       // local staging of the packed records (left, right), mgdcaprec records each
  char *mgdpeer[2][2]={{nullptr,nullptr},{nullptr,nullptr}}; unsigned mgdpeercap[2]={0,0}; bool mgdpeeripc[2]={false,false};
  int mgdparity=0; unsigned long long mgdfallbacks=0;
  void MgDirectSetup(); void MgDirectFree(); void MgExchangeDirect();
  std::vector<unsigned> mgcuts;             // cell columns of all the cuts (nranks+1), needed by the re-balancing
  unsigned long long mgsentrecords=0; double mgphase_ms[6]={0,0,0,0,0,0}; bool mgprofile=false;
  void SmallCopyAsync(void *dst,const void *src,size_t bytes);   // 4..128-byte transfers through a one-warp kernel (see engine.cu)
  void MgCommInit(const void *uid128,int rank,int nranks);
  void MgCommAttach(void *ncclcomm,int rank,int nranks);
  void MgCommDestroy();
  void MgCommAbort();
  void MgConfigCuts(int rank,int nranks,const unsigned *cuts);
  void MgExchange();
  void MgCellDivide(dsph_divinfo *info);
  void MgForces(int interstep,dsph_forceres *res);
  void MgStep(double *stepdt,dsph_forceres *res);
  bool MgRebalance(double bound_weight,double rank_cost,unsigned *cuts_out);
  void MgColumnCounts(unsigned *fluid,unsigned *bound,unsigned ncol);
  void MgPack(void *sendleft,void *sendright);
  void MgUnpack(const unsigned counts[8],const void *recvleft,const void *recvright);
  void ForcesDownload(unsigned first,unsigned count,float *ace,float *ar);
  void MeasurePoints(unsigned n,const double *points,float *press,float *vel,float *kcorr);
  char *measuredev=nullptr; size_t measurecap=0;
  int numsm=148;                // multiprocessors of the device (grid sizing of the grid-stride kernels)
  // in-solver gauges (gauges.cu)
  std::vector<dsph_gauge> gauges; std::vector<double> gaugenext; std::vector<dsph_gaugeres> gaugeres;
  double gaugetime=0;        // TimeStep of the run as dsph_step / dsph_mgpu_step advance it (gauges and attached motion laws use it)
  bool celltableok=false;
  char *gaugedev=nullptr; size_t gaugedevcap=0; float *gaugehost=nullptr; size_t gaugehostcap=0; float *gaugepartace=nullptr; unsigned gaugepartacecap=0;
  void GaugesConfig(unsigned n,const dsph_gauge *g);
  unsigned GaugesRun(double timestep,bool all,dsph_gaugeres *res);
  void GaugesFree();
  void Pips(unsigned long long *fr,unsigned long long *fc,unsigned long long *br,unsigned long long *bc);
  double DtVariable(double *dt1,double *dt2,double *dt3=nullptr)const;
  void FillForceRes(dsph_forceres *res)const;
  StepArrays MakeStepArrays();
  void ShiftCoefs(double dt,double &coefumagn,double &coeftfs,float &maxdist)const;
  void StepSymPre(double dt);
  void StepSymCor(double dt);
  void StepVerlet(double dt);
  void Step(double *stepdt,dsph_forceres *res);
  void FlushStepTimer();
  void AccumTimers(bool divide);
  void GetTimers(dsph_timers *t,bool reset);
};

std::vector<unsigned> MgBalancedCuts(const std::vector<double> &columnweight,int nranks,unsigned minwidth);   // nranks+1 cut columns
void MgUniqueId(void *id128);   // ncclGetUniqueId (mgpu_comm.cu)
int MgNcclVersion();

} // namespace dsph
