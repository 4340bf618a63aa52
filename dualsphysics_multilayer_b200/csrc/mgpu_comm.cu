// mgpu_comm.cu -- native transport of the x-slab decomposition: the halo / migrant exchange,
// the all-reduce of the dt maxima, the whole-step sequencing and the re-balancing of the cuts run inside the library on the
// context's own CUDA stream with NCCL send/recv/all-reduce, one process per GPU.  The reference is single-GPU (vestigial hooks
// S/DualSphDef.h:494-499, S/JDsDcellDef.h:30-31); the sequence is JSphGpuSingle::ComputeStep_Sym/_Ver (S/JSphGpuSingle.cpp:548-596)
// with an exchange in front of every RunCellDivide and an all-reduce(max) between the force kernels and DtVariable.
//
// Per exchange (one neighbour-list build):
//   classify + scan (2 kernels)  ->  ncclSend/Recv of the 4+4 class totals with the <= 2 neighbours (device to device)
//   -> ONE host synchronisation (own + announced totals through the mapped pinned block)
//   -> pack kernel -> ncclSend/Recv of the records -> unpack kernels, all enqueued on the stream without further host waits.
// (round 1 drove the same kernels from Python through torch.distributed: 4 stream synchronisations, two P2P rounds with host
//  waits and ~4.7 ms per step at 8 GPUs.)
//
// NCCL is resolved at run time (dlopen: the copy already loaded in the process, e.g. by torch, else the system libnccl.so.2), so the
// library has no link-time dependency on it and single-GPU users never load it.
#include "engine.h"
#include <nccl.h>
#include <dlfcn.h>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <unistd.h>

namespace dsph {

namespace {
struct NcclApi {
  void *handle=nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)=nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*,int,ncclUniqueId,int)=nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)=nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t)=nullptr;
  ncclResult_t (*Send)(const void*,size_t,ncclDataType_t,int,ncclComm_t,cudaStream_t)=nullptr;
  ncclResult_t (*Recv)(void*,size_t,ncclDataType_t,int,ncclComm_t,cudaStream_t)=nullptr;
  ncclResult_t (*AllReduce)(const void*,void*,size_t,ncclDataType_t,ncclRedOp_t,ncclComm_t,cudaStream_t)=nullptr;
  ncclResult_t (*GroupStart)()=nullptr;
  ncclResult_t (*GroupEnd)()=nullptr;
  const char* (*GetErrorString)(ncclResult_t)=nullptr;
  ncclResult_t (*GetVersion)(int*)=nullptr;
  std::string error;
  bool Load(){
    if(handle)return true;
    const char *names[]={"libnccl.so.2","libnccl.so"};
    for(const char *n:names){ handle=dlopen(n,RTLD_NOW|RTLD_NOLOAD|RTLD_GLOBAL); if(handle)break; }      // a copy already in the process (torch's)
    if(!handle)for(const char *n:names){ handle=dlopen(n,RTLD_NOW|RTLD_GLOBAL); if(handle)break; }
    if(!handle){ error=std::string("NCCL library not found (libnccl.so.2): ")+(dlerror()? dlerror(): ""); return false; }
    bool ok=true;
    auto sym=[&](const char *name){ void *p=dlsym(handle,name); if(!p){ ok=false; error=std::string("NCCL symbol missing: ")+name; } return p; };
    GetUniqueId=(decltype(GetUniqueId))sym("ncclGetUniqueId");
    CommInitRank=(decltype(CommInitRank))sym("ncclCommInitRank");
    CommDestroy=(decltype(CommDestroy))sym("ncclCommDestroy");
    CommAbort=(decltype(CommAbort))sym("ncclCommAbort");
    Send=(decltype(Send))sym("ncclSend"); Recv=(decltype(Recv))sym("ncclRecv");
    AllReduce=(decltype(AllReduce))sym("ncclAllReduce");
    GroupStart=(decltype(GroupStart))sym("ncclGroupStart"); GroupEnd=(decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString=(decltype(GetErrorString))sym("ncclGetErrorString");
    GetVersion=(decltype(GetVersion))sym("ncclGetVersion");
    if(!ok){ handle=nullptr; return false; }
    return true;
  }
};
NcclApi g_nccl;
NcclApi& Nccl(){ if(!g_nccl.Load())throw EngineError(DSPH_ERR_STATE,g_nccl.error); return g_nccl; }

#define NCCLCHECK(x) do{ ncclResult_t r_=(x); if(r_!=ncclSuccess)throw EngineError(DSPH_ERR_CUDA,std::string(#x)+": "+Nccl().GetErrorString(r_)); }while(0)
#define CUCHECK2(x) do{ cudaError_t e_=(x); if(e_!=cudaSuccess)throw EngineError(DSPH_ERR_CUDA,std::string(#x)+": "+cudaGetErrorString(e_)); }while(0)

void GrowBuffer(char *&p,size_t &cap,size_t need){
  if(need<=cap)return;
  if(p){ cudaFree(p); p=nullptr; cap=0; }
  const size_t n=need+need/4+65536;
  CUCHECK2(cudaMalloc((void**)&p,n));
  cap=n;
}
struct Stopwatch {
  std::chrono::steady_clock::time_point t0; bool on; cudaStream_t stm;
  Stopwatch(bool on_,cudaStream_t s):on(on_),stm(s){ if(on){ cudaStreamSynchronize(stm); t0=std::chrono::steady_clock::now(); } }
  void Lap(double &acc){ if(on){ cudaStreamSynchronize(stm); const auto t=std::chrono::steady_clock::now(); acc+=std::chrono::duration<double,std::milli>(t-t0).count(); t0=t; } }
};
} // namespace

void MgUniqueId(void *id128){
  ncclUniqueId id; NCCLCHECK(Nccl().GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId)==128,"ncclUniqueId is 128 bytes");
  memcpy(id128,&id,sizeof(id));
}
int MgNcclVersion(){ int v=0; if(g_nccl.Load() && g_nccl.GetVersion)g_nccl.GetVersion(&v); return v; }

void Engine::MgCommInit(const void *uid128,int rank,int nranks){
  if(nranks<1 || rank<0 || rank>=nranks)throw EngineError(DSPH_ERR_ARG,"Invalid rank / number of ranks.");
  MgCommDestroy();
  CUCHECK2(cudaSetDevice(device));
  ncclUniqueId id; memcpy(&id,uid128,sizeof(id));
  ncclComm_t comm=nullptr;
  NCCLCHECK(Nccl().CommInitRank(&comm,nranks,id,rank));
  mgcomm=comm; mgcomm_own=true;
  if(!mgcntdev)CUCHECK2(cudaMalloc((void**)&mgcntdev,32*sizeof(unsigned)));
  CUCHECK2(cudaMemsetAsync(mgcntdev,0,32*sizeof(unsigned),stm));
}
void Engine::MgCommAttach(void *ncclcomm,int rank,int nranks){
  if(!ncclcomm || nranks<1 || rank<0 || rank>=nranks)throw EngineError(DSPH_ERR_ARG,"Invalid communicator / rank.");
  Nccl();
  MgCommDestroy();
  CUCHECK2(cudaSetDevice(device));
  mgcomm=ncclcomm; mgcomm_own=false;
  if(!mgcntdev)CUCHECK2(cudaMalloc((void**)&mgcntdev,32*sizeof(unsigned)));
  CUCHECK2(cudaMemsetAsync(mgcntdev,0,32*sizeof(unsigned),stm));
}
// Called by ANOTHER thread when a rank of the group failed: pending and future collectives of this communicator return with an error
// instead of waiting for the failed rank for ever.
void Engine::MgCommAbort(){
  if(mgcomm && g_nccl.handle && g_nccl.CommAbort){ void *c=mgcomm; mgcomm=nullptr; g_nccl.CommAbort((ncclComm_t)c); }
}
void Engine::MgCommDestroy(){
  if(mgcomm && mgcomm_own && g_nccl.handle){ cudaStreamSynchronize(stm); g_nccl.CommDestroy((ncclComm_t)mgcomm); }
  mgcomm=nullptr; mgcomm_own=false;
  MgDirectFree();
  for(int d=0;d<2;d++){ if(mgsend[d]){ cudaFree(mgsend[d]); mgsend[d]=nullptr; mgsendcap[d]=0; } if(mgrecv[d]){ cudaFree(mgrecv[d]); mgrecv[d]=nullptr; mgrecvcap[d]=0; } }
  if(mgcntdev){ cudaFree(mgcntdev); mgcntdev=nullptr; }
  if(mghistdev){ cudaFree(mghistdev); mghistdev=nullptr; mghistcap=0; }
}

void Engine::MgConfigCuts(int rank,int nranks,const unsigned *cuts){
  if(nranks<1 || rank<0 || rank>=nranks || !cuts)throw EngineError(DSPH_ERR_ARG,"Invalid slab configuration.");
  for(int r=0;r<nranks;r++)if(cuts[r]>=cuts[r+1])throw EngineError(DSPH_ERR_ARG,"The cuts must be increasing cell columns.");
  MgConfig(rank,nranks,cuts[rank],cuts[rank+1]);
  mgcuts.assign(cuts,cuts+nranks+1);
}

// ------------------------------------------------------------------------------------------------
// direct transport: records are written by the pack kernel straight into the neighbour's receive buffer over NVLink
// ------------------------------------------------------------------------------------------------
// Negotiated once per communicator (collective, first exchange).  Every rank allocates two receive buffers per side (alternating
// per exchange), tells its <= 2 neighbours where they are {host, pid, device, pointers, CUDA IPC handles, capacity} with ncclSend /
// ncclRecv, and maps the neighbours' buffers: the pointer itself + cudaDeviceEnablePeerAccess when the neighbour is a thread of this
// process (bin/dsph_b200 -mgpu:N), cudaIpcOpenMemHandle when it is another process (torchrun).  An all-reduce(min) of "mapping
// worked" decides for the whole group; DSPH_MGPU_DIRECT=0 keeps the ncclSend/ncclRecv transport of the records.
// Why two buffers: a rank may pack exchange k+1 while its neighbour still unpacks exchange k; the token of exchange k+1 that the
// neighbour sends after that unpack (stream order) is received before this rank packs exchange k+2 into the first buffer again.
namespace {
struct MgHello {
  unsigned long long host; int pid; int device; unsigned caprec; unsigned recbytes; unsigned valid; unsigned pad;
  unsigned long long ptr[2][2];
  cudaIpcMemHandle_t handle[2][2];
};
unsigned long long HostHash(){
  char name[256]; memset(name,0,sizeof(name)); gethostname(name,sizeof(name)-1);
  unsigned long long h=1469598103934665603ull; for(const char *c=name;*c;c++){ h^=(unsigned char)*c; h*=1099511628211ull; }
  return h;
}
}
void Engine::MgDirectFree(){
  for(int d=0;d<2;d++)for(int q=0;q<2;q++){
    if(mgdpeer[d][q] && mgdpeeripc[d])cudaIpcCloseMemHandle(mgdpeer[d][q]);
    mgdpeer[d][q]=nullptr;
    if(mgdrecv[d][q]){ cudaFree(mgdrecv[d][q]); mgdrecv[d][q]=nullptr; }
  }
  for(int d=0;d<2;d++)if(mgdsend[d]){ cudaFree(mgdsend[d]); mgdsend[d]=nullptr; }
  mgdpeeripc[0]=mgdpeeripc[1]=false; mgdpeercap[0]=mgdpeercap[1]=0; mgdcaprec=0; mgdirect=-1; mgdparity=0;
}
void Engine::MgDirectSetup(){
  if(mgdirect>=0)return;
  NcclApi &N=Nccl();
  ncclComm_t comm=(ncclComm_t)mgcomm;
  const int left=MgLeft(), right=MgRight();
  const char *env=getenv("DSPH_MGPU_DIRECT");
  unsigned ok=(env && atoi(env)==0? 0u: 1u);
  MgHello me; memset(&me,0,sizeof(me));
  me.host=HostHash(); me.pid=int(getpid()); me.device=device; me.recbytes=MgRecordBytes();
  if(ok){
    const char *ecap=getenv("DSPH_MGPU_DIRECT_CAP");        // records per buffer (tests: a small value forces the fall-back path)
    mgdcaprec=(ecap && atoi(ecap)>0? unsigned(atoi(ecap)): cap/4u+1024u);
    for(int d=0;d<2 && ok;d++)for(int q=0;q<2 && ok;q++){
      if(cudaMalloc((void**)&mgdrecv[d][q],size_t(mgdcaprec)*me.recbytes+16)!=cudaSuccess){ cudaGetLastError(); mgdrecv[d][q]=nullptr; ok=0; break; }
      me.ptr[d][q]=(unsigned long long)(uintptr_t)mgdrecv[d][q];
      if(cudaIpcGetMemHandle(&me.handle[d][q],mgdrecv[d][q])!=cudaSuccess){ cudaGetLastError(); ok=0; }
    }
    for(int d=0;d<2 && ok;d++)if(cudaMalloc((void**)&mgdsend[d],size_t(mgdcaprec)*me.recbytes+16)!=cudaSuccess){ cudaGetLastError(); mgdsend[d]=nullptr; ok=0; }
    me.caprec=mgdcaprec;
  }
  me.valid=ok;
  // hello to / from the neighbours (device staging: [0] mine, [1] left's, [2] right's)
  MgHello *dev=nullptr; CUCHECK2(cudaMalloc((void**)&dev,3*sizeof(MgHello)));
  CUCHECK2(cudaMemsetAsync(dev,0,3*sizeof(MgHello),stm));
  CUCHECK2(cudaMemcpyAsync(dev,&me,sizeof(me),cudaMemcpyHostToDevice,stm));
  NCCLCHECK(N.GroupStart());
  if(left>=0)NCCLCHECK(N.Send(dev,sizeof(MgHello),ncclUint8,left,comm,stm));
  if(right>=0)NCCLCHECK(N.Send(dev,sizeof(MgHello),ncclUint8,right,comm,stm));
  if(right>=0)NCCLCHECK(N.Recv(dev+2,sizeof(MgHello),ncclUint8,right,comm,stm));
  if(left>=0)NCCLCHECK(N.Recv(dev+1,sizeof(MgHello),ncclUint8,left,comm,stm));
  NCCLCHECK(N.GroupEnd());
  MgHello nb[2];
  CUCHECK2(cudaMemcpyAsync(nb,dev+1,2*sizeof(MgHello),cudaMemcpyDeviceToHost,stm));
  CUCHECK2(cudaStreamSynchronize(stm));
  // map: this rank is the RIGHT neighbour of its left neighbour (writes its "from right" buffers) and vice versa
  for(int d=0;d<2;d++){
    const int peer=(d==0? left: right);
    if(peer<0)continue;
    const MgHello &h=nb[d];
    const int theirside=(d==0? 1: 0);
    if(!ok || !h.valid || h.host!=me.host || h.recbytes!=me.recbytes){ ok=0; continue; }
    mgdpeercap[d]=std::min(h.caprec,mgdcaprec);      // the records are staged in this rank's buffer of mgdcaprec records first
    if(h.pid==me.pid){
      if(h.device!=device){
        int can=0; cudaDeviceCanAccessPeer(&can,device,h.device);
        if(!can){ ok=0; continue; }
        const cudaError_t e=cudaDeviceEnablePeerAccess(h.device,0);
        if(e!=cudaSuccess && e!=cudaErrorPeerAccessAlreadyEnabled){ ok=0; cudaGetLastError(); continue; }
        cudaGetLastError();
      }
      for(int q=0;q<2;q++)mgdpeer[d][q]=(char*)(uintptr_t)h.ptr[theirside][q];
      mgdpeeripc[d]=false;
    }
    else{
      mgdpeeripc[d]=true;
      for(int q=0;q<2;q++){
        void *p=nullptr;
        if(cudaIpcOpenMemHandle(&p,h.handle[theirside][q],cudaIpcMemLazyEnablePeerAccess)!=cudaSuccess){ cudaGetLastError(); ok=0; p=nullptr; }
        mgdpeer[d][q]=(char*)p;
      }
    }
  }
  // the whole group uses the same transport
  unsigned *flag=(unsigned*)dev;
  CUCHECK2(cudaMemcpyAsync(flag,&ok,sizeof(unsigned),cudaMemcpyHostToDevice,stm));
  NCCLCHECK(N.AllReduce(flag,flag,1,ncclUint32,ncclMin,comm,stm));
  unsigned all=0;
  CUCHECK2(cudaMemcpyAsync(&all,flag,sizeof(unsigned),cudaMemcpyDeviceToHost,stm));
  CUCHECK2(cudaStreamSynchronize(stm));
  cudaFree(dev);
  if(!all){ const int keep=0; MgDirectFree(); mgdirect=keep; return; }
  mgdirect=1; mgdparity=0;
}

// ------------------------------------------------------------------------------------------------
// halo + migrant exchange, one host synchronisation
// ------------------------------------------------------------------------------------------------
// Peer-memory transport (default, MgDirectSetup):
//   classify + scan (2 kernels) -> pack kernel (bases from the device-side totals) -> push kernel: coalesced copy into the neighbours' buffers
//   -> ncclSend/Recv of the 20-byte tokens {4 totals, used} = "records landed" -> ONE host synchronisation -> unpack kernels.
void Engine::MgExchangeDirect(){
  NcclApi &N=Nccl();
  ncclComm_t comm=(ncclComm_t)mgcomm;
  const int left=MgLeft(), right=MgRight();
  const int q=mgdparity; mgdparity^=1;
  Stopwatch sw(mgprofile,stm);
  MgClassifyAsync();
  sw.Lap(mgphase_ms[0]);
  MgPackDirectAsync(mgdsend[0],mgdsend[1],(left>=0? mgdpeercap[0]: 0xffffffffu),(right>=0? mgdpeercap[1]: 0xffffffffu),mgcntdev);
  sw.Lap(mgphase_ms[2]);
  MgPushAsync(mgdsend[0],mgdpeer[0][q],mgdsend[1],mgdpeer[1][q],mgcntdev);
  sw.Lap(mgphase_ms[3]);
  NCCLCHECK(N.GroupStart());
  // order: sends (left, right), then receives (right, left).  In a ring of TWO ranks both neighbours are the same peer, and NCCL matches
  // the operations with one peer in the order they were posted: what goes to the peer's right side must meet its "from right" receive.
  if(left>=0)NCCLCHECK(N.Send(mgcntdev,5,ncclUint32,left,comm,stm));
  if(right>=0)NCCLCHECK(N.Send(mgcntdev+5,5,ncclUint32,right,comm,stm));
  if(right>=0)NCCLCHECK(N.Recv(mgcntdev+21,5,ncclUint32,right,comm,stm));
  if(left>=0)NCCLCHECK(N.Recv(mgcntdev+16,5,ncclUint32,left,comm,stm));
  NCCLCHECK(N.GroupEnd());
  SmallCopyAsync(hpind->mgtoken,mgcntdev,10*sizeof(unsigned));
  SmallCopyAsync(hpind->mgtoken+10,mgcntdev+16,10*sizeof(unsigned));
  CUCHECK2(cudaStreamSynchronize(stm));                           // the only host wait of the exchange
  unsigned rcv[8]; bool rdirect[2];
  for(int g=0;g<4;g++){ mgcounts[g]=hpin->mgtoken[g]; mgcounts[4+g]=hpin->mgtoken[5+g]; rcv[g]=hpin->mgtoken[10+g]; rcv[4+g]=hpin->mgtoken[15+g]; }
  const bool mydirect=(hpin->mgtoken[4]!=0);
  rdirect[0]=(left<0 || hpin->mgtoken[14]!=0); rdirect[1]=(right<0 || hpin->mgtoken[19]!=0);
  if(left<0)for(int g=0;g<4;g++)rcv[g]=0;
  if(right<0)for(int g=0;g<4;g++)rcv[4+g]=0;
  sw.Lap(mgphase_ms[1]);
  const size_t rb=MgRecordBytes();
  const size_t nsend[2]={size_t(mgcounts[0])+mgcounts[1]+mgcounts[2]+mgcounts[3],size_t(mgcounts[4])+mgcounts[5]+mgcounts[6]+mgcounts[7]};
  const size_t nrecv[2]={size_t(rcv[0])+rcv[1]+rcv[2]+rcv[3],size_t(rcv[4])+rcv[5]+rcv[6]+rcv[7]};
  const void *src[2]={mgdrecv[0][q],mgdrecv[1][q]};
  if(!mydirect || !rdirect[0] || !rdirect[1]){
    // fall-back of this exchange for the sides whose records did not fit: local pack + ncclSend / ncclRecv
    mgdfallbacks++;
    if(!mydirect){
      for(int d=0;d<2;d++)GrowBuffer(mgsend[d],mgsendcap[d],std::max<size_t>(nsend[d],1)*rb);
      mgcounted=true;
      MgPackAsync(mgsend[0],mgsend[1]);
    }
    for(int d=0;d<2;d++)if(!rdirect[d]){ GrowBuffer(mgrecv[d],mgrecvcap[d],std::max<size_t>(nrecv[d],1)*rb); src[d]=mgrecv[d]; }
    NCCLCHECK(N.GroupStart());
    if(left>=0 && !mydirect && nsend[0])NCCLCHECK(N.Send(mgsend[0],nsend[0]*rb,ncclUint8,left,comm,stm));
    if(right>=0 && !mydirect && nsend[1])NCCLCHECK(N.Send(mgsend[1],nsend[1]*rb,ncclUint8,right,comm,stm));
    if(right>=0 && !rdirect[1] && nrecv[1])NCCLCHECK(N.Recv(mgrecv[1],nrecv[1]*rb,ncclUint8,right,comm,stm));
    if(left>=0 && !rdirect[0] && nrecv[0])NCCLCHECK(N.Recv(mgrecv[0],nrecv[0]*rb,ncclUint8,left,comm,stm));
    NCCLCHECK(N.GroupEnd());
  }
  mgcounted=false;
  sw.Lap(mgphase_ms[3]);
  MgUnpackAsync(rcv,src[0],src[1]);
  sw.Lap(mgphase_ms[4]);
  mgsentrecords+=nsend[0]+nsend[1];
}

void Engine::MgExchange(){
  if(!mgpu)return;
  if(!mgcomm)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_exchange needs dsph_mgpu_comm_init (or dsph_mgpu_comm_attach) first.");
  CUCHECK2(cudaSetDevice(device));
  MgDirectSetup();
  if(mgdirect==1){ MgExchangeDirect(); return; }
  NcclApi &N=Nccl();
  ncclComm_t comm=(ncclComm_t)mgcomm;
  const int left=MgLeft(), right=MgRight();
  Stopwatch sw(mgprofile,stm);
  // 1. classify, class totals -> mgscratch[1][0..7] (to left: 0..3, to right: 4..7)
  MgClassifyAsync();
  sw.Lap(mgphase_ms[0]);
  // 2. totals to the neighbours, theirs into mgcntdev[8..15] (from left: 8..11, from right: 12..15; zero without neighbour)
  NCCLCHECK(N.GroupStart());
  if(left>=0)NCCLCHECK(N.Send(mgscratch[1],4,ncclUint32,left,comm,stm));           // sends (left, right), receives (right, left): see MgExchangeDirect
  if(right>=0)NCCLCHECK(N.Send(mgscratch[1]+4,4,ncclUint32,right,comm,stm));
  if(right>=0)NCCLCHECK(N.Recv(mgcntdev+12,4,ncclUint32,right,comm,stm));
  if(left>=0)NCCLCHECK(N.Recv(mgcntdev+8,4,ncclUint32,left,comm,stm));
  NCCLCHECK(N.GroupEnd());
  SmallCopyAsync(hpind->mgtot,mgscratch[1],8*sizeof(unsigned));
  SmallCopyAsync(hpind->mgtot+8,mgcntdev+8,8*sizeof(unsigned));
  CUCHECK2(cudaStreamSynchronize(stm));                           // the only host wait of the exchange
  unsigned rcv[8];
  for(int g=0;g<8;g++){ mgcounts[g]=hpin->mgtot[g]; rcv[g]=hpin->mgtot[8+g]; }
  if(!np)for(int g=0;g<8;g++)mgcounts[g]=0;
  mgcounted=true;
  sw.Lap(mgphase_ms[1]);
  const size_t rb=MgRecordBytes();
  const size_t nsend[2]={size_t(mgcounts[0])+mgcounts[1]+mgcounts[2]+mgcounts[3],size_t(mgcounts[4])+mgcounts[5]+mgcounts[6]+mgcounts[7]};
  const size_t nrecv[2]={size_t(rcv[0])+rcv[1]+rcv[2]+rcv[3],size_t(rcv[4])+rcv[5]+rcv[6]+rcv[7]};
  for(int d=0;d<2;d++){ GrowBuffer(mgsend[d],mgsendcap[d],std::max<size_t>(nsend[d],1)*rb); GrowBuffer(mgrecv[d],mgrecvcap[d],std::max<size_t>(nrecv[d],1)*rb); }
  // 3. pack -> records to the neighbours -> unpack, all on the stream
  MgPackAsync(mgsend[0],mgsend[1]);
  sw.Lap(mgphase_ms[2]);
  NCCLCHECK(N.GroupStart());
  if(left>=0 && nsend[0])NCCLCHECK(N.Send(mgsend[0],nsend[0]*rb,ncclUint8,left,comm,stm));
  if(right>=0 && nsend[1])NCCLCHECK(N.Send(mgsend[1],nsend[1]*rb,ncclUint8,right,comm,stm));
  if(right>=0 && nrecv[1])NCCLCHECK(N.Recv(mgrecv[1],nrecv[1]*rb,ncclUint8,right,comm,stm));
  if(left>=0 && nrecv[0])NCCLCHECK(N.Recv(mgrecv[0],nrecv[0]*rb,ncclUint8,left,comm,stm));
  NCCLCHECK(N.GroupEnd());
  sw.Lap(mgphase_ms[3]);
  MgUnpackAsync(rcv,mgrecv[0],mgrecv[1]);
  sw.Lap(mgphase_ms[4]);
  mgsentrecords+=nsend[0]+nsend[1];
}

void Engine::MgCellDivide(dsph_divinfo *info){
  const unsigned npb0=npb, np0=np;
  MgExchange();
  // periodic copies along y / z of everything the rank holds now (received particles sit behind the np0 it had: boundary block, fluid block)
  if(mgpu && (cfg.periy || cfg.periz))RunPeriodicSlab(npb0,np0,npbper,npfper);
  CellDivide(false,info);
}

// force kernels, all-reduce(max) of the maxima (uint bit patterns of non-negative floats: integer max = float max), dt
void Engine::MgForces(int interstep,dsph_forceres *res){
  ForcesLaunch(interstep);
  if(mgpu){
    if(!mgcomm)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_step needs dsph_mgpu_comm_init first.");
    Stopwatch sw(mgprofile,stm);
    NCCLCHECK(Nccl().AllReduce(fmax,fmax,4,ncclUint32,ncclMax,(ncclComm_t)mgcomm,stm));
    sw.Lap(mgphase_ms[5]);
  }
  ForcesFinish(res);
}

// in-place all-reduce of a small float block on the stream (gauges of slab runs): sum, or max
void Engine::MgAllReduceFloats(float *dev,size_t n,bool max){
  if(!mgpu || !n)return;
  if(!mgcomm)throw EngineError(DSPH_ERR_STATE,"Gauges of an x-slab run need dsph_mgpu_comm_init first (their evaluation is collective).");
  NCCLCHECK(Nccl().AllReduce(dev,dev,n,ncclFloat,max? ncclMax: ncclSum,(ncclComm_t)mgcomm,stm));
}

void Engine::MgStep(double *stepdt,dsph_forceres *res){
  double dt;
  bool moving=false;
  if(cfg.tstep==DSPH_STEP_SYMPLECTIC){
    dt=symplecticdtpre;
    if(motionset)moving=motionset->Process(gaugetime,dt,motiondata);
    MgForces(DSPH_INTERSTEP_SYMPREDICTOR,nullptr);
    const double ddt_p=DtVariable(nullptr,nullptr);
    StepSymPre(dt);
    MgCellDivide(nullptr); FlushStepTimer();
    MgForces(DSPH_INTERSTEP_SYMCORRECTOR,res);
    const double ddt_c=DtVariable(nullptr,nullptr);
    StepSymCor(dt);
    symplecticdtpre=std::min(ddt_p,ddt_c);
  }
  else{
    MgForces(DSPH_INTERSTEP_VERLET,res);
    dt=DtVariable(nullptr,nullptr);
    if(motionset)moving=motionset->Process(gaugetime,dt,motiondata);
    StepVerlet(dt);
  }
  if(moving)MotionRun(motiondata.data(),unsigned(motiondata.size()),dt);   // every rank moves the moving particles it owns
  gaugetime+=dt;
  MgCellDivide(nullptr); FlushStepTimer();
  // Gauges (collective: partial sums of every rank, all-reduced).  The single-GPU step evaluates them before the closing divide, like
  // the reference (S/JSphCpuSingle.cpp:1095); here the halo copies still hold the state of the mid-step exchange at that point, so the
  // evaluation waits for the exchange of the closing divide: same particles, same states, the neighbour list of the new divide.
  if(!gauges.empty())GaugesRun(gaugetime,false,nullptr);
  if(stepdt)*stepdt=dt;
}

// ------------------------------------------------------------------------------------------------
// re-balancing (collective): new cuts from the global per-column histogram of owned particles (fluid 1, boundary bound_weight),
// scaled per slab to the measured cost of its rank when rank_cost>0 on every rank.  A cut moves by less than the width of the
// narrower of its two slabs, so a particle changes owner by at most one rank; the next exchange migrates fluid AND boundary
// particles that are outside the new range.  Returns true when the cuts changed (the caller then runs dsph_mgpu_celldiv).
// ------------------------------------------------------------------------------------------------
// Cuts that minimise the work of the busiest slab (whole cell columns; every slab at least minwidth columns wide): dynamic programme
// over (slab, last column), ties broken towards the smaller sum of squared slab work.  With ~18 fluid columns per slab (DB(160 M) on
// 8 GPUs) one column is 5.5 % of a slab, and cutting at the equal-share prefix positions (the first version) left one slab a whole
// column above the others.  ncol^2 * nranks operations on the host (576 columns, 8 slabs: 2.7 M).
std::vector<unsigned> MgBalancedCuts(const std::vector<double> &col,int nranks,unsigned minwidth){
  const unsigned ncol=unsigned(col.size());
  std::vector<unsigned> cuts;
  if(nranks<1 || !minwidth || (unsigned long long)nranks*minwidth>ncol){    // not admissible: the callers report it
    for(int r=0;r<=std::max(nranks,1);r++)cuts.push_back(std::min<unsigned long long>(ncol,(unsigned long long)r*minwidth));
    cuts.back()=ncol;
    return cuts;
  }
  std::vector<double> cum(ncol+1,0.0);
  for(unsigned i=0;i<ncol;i++)cum[i+1]=cum[i]+col[i];
  const double INF=1e300;
  const size_t W=size_t(ncol)+1;
  std::vector<double> fmax(size_t(nranks+1)*W,INF), fsq(size_t(nranks+1)*W,INF);
  std::vector<unsigned> from(size_t(nranks+1)*W,0u);
  fmax[0]=0; fsq[0]=0;
  for(int r=1;r<=nranks;r++){
    const unsigned emin=unsigned(r)*minwidth, emax=ncol-unsigned(nranks-r)*minwidth;
    for(unsigned e=(r==nranks? ncol: emin);e<=emax;e++){
      double bm=INF,bs=INF; unsigned bfrom=0;
      const unsigned smin=unsigned(r-1)*minwidth;
      for(unsigned st=smin;st+minwidth<=e;st++){
        const double pm=fmax[size_t(r-1)*W+st];
        if(pm>=INF)continue;
        const double load=cum[e]-cum[st];
        const double m=std::max(pm,load), q=fsq[size_t(r-1)*W+st]+load*load;
        if(m<bm || (m==bm && q<bs)){ bm=m; bs=q; bfrom=st; }
      }
      fmax[size_t(r)*W+e]=bm; fsq[size_t(r)*W+e]=bs; from[size_t(r)*W+e]=bfrom;
    }
  }
  cuts.assign(size_t(nranks)+1,0u);
  unsigned e=ncol;
  for(int r=nranks;r>=1;r--){ cuts[r]=e; e=from[size_t(r)*W+e]; }
  cuts[0]=0;
  return cuts;
}

bool Engine::MgRebalance(double bound_weight,double rank_cost,unsigned *cuts_out){
  if(!mgpu)return false;
  if(!mgcomm || int(mgcuts.size())!=mgnranks+1)throw EngineError(DSPH_ERR_STATE,"dsph_mgpu_rebalance needs dsph_mgpu_comm_init and dsph_mgpu_config_cuts first.");
  CUCHECK2(cudaSetDevice(device));
  const unsigned ncol=k.domcells[0];
  const unsigned minwidth=2u*MgHalo()+1u;
  if(unsigned(mgnranks)*minwidth>ncol)throw EngineError(DSPH_ERR_ARG,"Too many ranks for the cell columns of the domain.");
  std::vector<unsigned> fl(ncol),bd(ncol);
  MgColumnCounts(fl.data(),bd.data(),ncol);
  const unsigned nval=ncol+unsigned(mgnranks);
  if(mghistcap<nval){ if(mghistdev)cudaFree(mghistdev); CUCHECK2(cudaMalloc((void**)&mghistdev,size_t(nval)*sizeof(double))); mghistcap=nval; }
  std::vector<double> h(nval,0.0);
  for(unsigned i=0;i<ncol;i++)h[i]=double(fl[i])+bound_weight*double(bd[i]);
  h[ncol+mgrank]=rank_cost;
  CUCHECK2(cudaMemcpyAsync(mghistdev,h.data(),size_t(nval)*sizeof(double),cudaMemcpyHostToDevice,stm));
  NCCLCHECK(Nccl().AllReduce(mghistdev,mghistdev,nval,ncclDouble,ncclSum,(ncclComm_t)mgcomm,stm));
  CUCHECK2(cudaMemcpyAsync(h.data(),mghistdev,size_t(nval)*sizeof(double),cudaMemcpyDeviceToHost,stm));
  CUCHECK2(cudaStreamSynchronize(stm));
  std::vector<double> hist(h.begin(),h.begin()+ncol);
  bool costs_ok=true;
  for(int r=0;r<mgnranks;r++)if(!(h[ncol+r]>0))costs_ok=false;
  if(costs_ok)for(int r=0;r<mgnranks;r++){
    double tot=0; for(unsigned i=mgcuts[r];i<mgcuts[r+1];i++)tot+=hist[i];
    if(tot>0){ const double f=h[ncol+r]/tot; for(unsigned i=mgcuts[r];i<mgcuts[r+1];i++)hist[i]*=f; }
  }
  const std::vector<unsigned> target=MgBalancedCuts(hist,mgnranks,minwidth);
  std::vector<unsigned> nw(mgcuts);
  for(int r=1;r<mgnranks;r++){
    const long long room=(long long)std::min(mgcuts[r]-mgcuts[r-1],mgcuts[r+1]-mgcuts[r])-1;
    const long long lo=(long long)mgcuts[r]-room, hi=(long long)mgcuts[r]+room;
    nw[r]=unsigned(std::min(std::max((long long)target[r],lo),hi));
  }
  for(int r=1;r<mgnranks;r++)nw[r]=std::max(nw[r],nw[r-1]+minwidth);        // keep every slab at least minwidth wide
  for(int r=mgnranks-1;r>0;r--)nw[r]=std::min(nw[r],nw[r+1]-minwidth);
  if(cuts_out)for(int r=0;r<=mgnranks;r++)cuts_out[r]=nw[r];
  if(nw==mgcuts)return false;
  MgConfigCuts(mgrank,mgnranks,nw.data());
  return true;
}

} // namespace dsph
