// mgpu.cuh -- x-slab domain decomposition helpers: classification of the particles of one slab into
// keep / migrate / halo, deterministic packing into fixed-size records, and unpacking of received records as owned
// particles (migrants) or ghosts.  Ghosts use the reference's CODE_PERIODIC mechanism (S/JSphGpu_ker.cu:2614): they are
// neighbours only, are ignored by AceMax / PART output and are dropped by the next divide (CODE_OUTIGNORE).
// The reference has no multi-GPU code (only vestigial hooks: DCEL_CodeDomLeft/Right S/JDsDcellDef.h:30-31, MGDIV_X
// S/DualSphDef.h:494-499); this is new work.  Transport is NCCL send/recv through torch.distributed (mgpu.py).
#pragma once
#include "devtypes.h"
#include "celldiv.cuh"

namespace dsph {

// one record = 88 bytes, 8-byte aligned (plain float arrays: a float4 member would force 16-byte alignment and padding)
struct MgRecord {
  double pos[3];
  float velrhop[4];
  double pospre[3];
  float aux[4];        // velrhoppre (Symplectic, mid-step) or velrhopm1 (Verlet)
  unsigned idp, code;
};
static_assert(sizeof(MgRecord)==88,"MgRecord layout is part of the wire format (oracle/sph_oracle.cpp uses the same)");

struct MgArgs {
  unsigned np, npb;
  unsigned xlo, xhi, sd;      // owned cell columns [xlo,xhi), halo width in cells
  int hasleft, hasright;
  unsigned domcellcode;
  const unsigned *dcell; typecode *code;
  // periodic along x: the ranks form a ring; the left link of the first rank and the right link of the last rank cross the periodic face
  int ring; int wrapleft, wrapright;
  double xs0, xs1, xlen;      // physical x range of the slab inside the real domain, length of the period
  double bandwidth;           // halo width in space (sd cells)
  double shiftleft[3], shiftright[3];   // added to the position of halo copies sent through a wrapping link
  const double *pos;
};

// classes per particle and direction; kind: 0 boundary ghost, 1 boundary migrant (only after the cuts moved: re-balancing),
// 2 fluid migrant, 3 fluid ghost.  Class g = 4*direction + kind (direction 0 = left neighbour, 1 = right neighbour).
constexpr int MG_CLASSES=8;
__device__ __forceinline__ void MgClassify(unsigned p,const MgArgs &a,int &toleft,int &toright){
  toleft=toright=-1;
  const unsigned rc=a.code[p];
  if((rc&DSPH_CODE_MASKSPECIAL)!=DSPH_CODE_NORMAL)return;
  const unsigned cel=a.dcell[p];
  if(cel==0xFFFFFFFFu)return;
  const unsigned cx=DcCellx(a.domcellcode,cel);
  const bool fluid=((rc&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID);
  if(a.ring){
    // Ring of slabs.  A particle that left through the periodic face was wrapped to the other end of the domain by the update
    // (UpdatePos), so "left of the slab" / "right of the slab" is decided by the shorter way round; a step moves a particle by less than
    // a cell, so that way is always the short one.
    if(cx<a.xlo || cx>=a.xhi){
      const double x=a.pos[3*size_t(p)];
      double dr=x-a.xs1; if(dr<0)dr+=a.xlen;
      double dl=a.xs0-x; if(dl<0)dl+=a.xlen;
      if(dr<dl)toright=(fluid? 2: 1); else toleft=(fluid? 2: 1);
    }
    else{
      // halo bands: whole cell columns at the cuts between slabs; at the periodic face the band is measured in space, because the
      // shifted copy must fall inside the margin of the map (the face may cut the last real column anywhere)
      if(a.wrapleft){ if(a.pos[3*size_t(p)]<a.xs0+a.bandwidth)toleft=(fluid? 3: 0); }
      else if(cx<a.xlo+a.sd)toleft=(fluid? 3: 0);
      if(a.wrapright){ if(a.pos[3*size_t(p)]>=a.xs1-a.bandwidth)toright=(fluid? 3: 0); }
      else if(cx+a.sd>=a.xhi)toright=(fluid? 3: 0);
    }
    return;
  }
  if(a.hasleft){
    if(cx<a.xlo)toleft=(fluid? 2: 1);
    else if(cx<a.xlo+a.sd)toleft=(fluid? 3: 0);
  }
  if(a.hasright){
    if(cx>=a.xhi)toright=(fluid? 2: 1);
    else if(cx+a.sd>=a.xhi)toright=(fluid? 3: 0);
  }
}

// Deterministic (index-ordered) compaction in two passes over one byte per particle:
//  ker_mg_classify  flags[p] = 6 one-bit classes, per-block class counts (also retires the ghosts of the previous divide)
//  ker_mg_scanblocks  exclusive scan of the per-block counts of each class (one CTA per class), class totals
//  ker_mg_pack        rank inside the block by ballot/popc + block offset -> record slot
#define MG_BLOCK 256
__global__ void __launch_bounds__(MG_BLOCK) ker_mg_classify(MgArgs a,unsigned char* __restrict__ flags,unsigned* __restrict__ blockcnt){
  __shared__ unsigned wcnt[MG_CLASSES][MG_BLOCK/32];
  const unsigned p=blockIdx.x*MG_BLOCK+threadIdx.x;
  unsigned f=0;
  if(p<a.np){
    const unsigned rc=a.code[p];
    if((rc&DSPH_CODE_MASKSPECIAL)==DSPH_CODE_PERIODIC)a.code[p]=typecode((rc&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_OUTIGNORE);
    int tl,tr; MgClassify(p,a,tl,tr);
    if(tl>=0)f|=1u<<tl;
    if(tr>=0)f|=1u<<(4+tr);
    flags[p]=(unsigned char)f;
  }
  const unsigned w=threadIdx.x>>5, lane=threadIdx.x&31;
  #pragma unroll
  for(int g=0;g<MG_CLASSES;g++){
    const unsigned bal=__ballot_sync(0xffffffffu,(f>>g)&1u);
    if(lane==0)wcnt[g][w]=__popc(bal);
  }
  __syncthreads();
  if(threadIdx.x<MG_CLASSES){
    unsigned t=0;
    #pragma unroll
    for(int i=0;i<MG_BLOCK/32;i++)t+=wcnt[threadIdx.x][i];
    blockcnt[size_t(threadIdx.x)*gridDim.x+blockIdx.x]=t;
  }
}
// grid = MG_CLASSES CTAs of 1024 threads: in-place exclusive scan of blockcnt[g*nb .. g*nb+nb), total[g].  Every thread owns
// MG_SCAN_ITEMS consecutive counts, so one sweep of the CTA covers 16 384 block counts (22 M particles = 86 k counts: 6 sweeps; the
// first version, one count per thread and sweep, spent 0.1 ms per exchange in its 84 serial sweeps).
constexpr int MG_SCAN_ITEMS=16;
__global__ void __launch_bounds__(1024) ker_mg_scanblocks(unsigned nb,unsigned* __restrict__ blockcnt,unsigned* __restrict__ total){
  __shared__ unsigned wsum[32];
  __shared__ unsigned carry;
  unsigned *v=blockcnt+size_t(blockIdx.x)*nb;
  const unsigned w=threadIdx.x>>5, lane=threadIdx.x&31;
  if(threadIdx.x==0)carry=0;
  __syncthreads();
  for(unsigned base=0;base<nb;base+=1024u*MG_SCAN_ITEMS){
    const unsigned i0=base+threadIdx.x*MG_SCAN_ITEMS;
    unsigned x[MG_SCAN_ITEMS]; unsigned sum=0;
    #pragma unroll
    for(int j=0;j<MG_SCAN_ITEMS;j++){ x[j]=(i0+j<nb? v[i0+j]: 0u); sum+=x[j]; }
    unsigned inc=sum;
    #pragma unroll
    for(int d=1;d<32;d<<=1){ const unsigned y=__shfl_up_sync(0xffffffffu,inc,d); if(lane>=unsigned(d))inc+=y; }
    if(lane==31)wsum[w]=inc;
    __syncthreads();
    if(w==0){
      unsigned t=wsum[lane];
      #pragma unroll
      for(int d=1;d<32;d<<=1){ const unsigned y=__shfl_up_sync(0xffffffffu,t,d); if(lane>=unsigned(d))t+=y; }
      wsum[lane]=t;
    }
    __syncthreads();
    const unsigned c0=carry;
    unsigned excl=c0+(w? wsum[w-1]: 0u)+inc-sum;
    #pragma unroll
    for(int j=0;j<MG_SCAN_ITEMS;j++){ if(i0+j<nb)v[i0+j]=excl; excl+=x[j]; }
    __syncthreads();
    if(threadIdx.x==1023)carry=c0+wsum[31];
    __syncthreads();
  }
  if(threadIdx.x==0)total[blockIdx.x]=carry;
}

// List of the OWNED fluid particles of the sorted state (index order kept): the force kernels take their particle from this list so
// that halo copies, which are interleaved with the owned particles at both ends of every cell row, occupy no lanes.
//  ker_mg_own_count -> ker_mg_scanblocks (1 class) -> ker_mg_own_write
__global__ void __launch_bounds__(MG_BLOCK) ker_mg_own_count(unsigned np,unsigned npb,const typecode* __restrict__ code,unsigned* __restrict__ blockcnt){
  __shared__ unsigned wcnt[MG_BLOCK/32];
  const unsigned p=npb+blockIdx.x*MG_BLOCK+threadIdx.x;
  const bool own=(p<np && (code[p]&DSPH_CODE_MASKSPECIAL)!=DSPH_CODE_PERIODIC);
  const unsigned bal=__ballot_sync(0xffffffffu,own);
  if((threadIdx.x&31)==0)wcnt[threadIdx.x>>5]=__popc(bal);
  __syncthreads();
  if(threadIdx.x==0){
    unsigned t=0;
    #pragma unroll
    for(int i=0;i<MG_BLOCK/32;i++)t+=wcnt[i];
    blockcnt[blockIdx.x]=t;
  }
}
__global__ void __launch_bounds__(MG_BLOCK) ker_mg_own_write(unsigned np,unsigned npb,const typecode* __restrict__ code,const unsigned* __restrict__ blockoff,unsigned* __restrict__ ownlist){
  __shared__ unsigned wcnt[MG_BLOCK/32];
  const unsigned p=npb+blockIdx.x*MG_BLOCK+threadIdx.x;
  const bool own=(p<np && (code[p]&DSPH_CODE_MASKSPECIAL)!=DSPH_CODE_PERIODIC);
  const unsigned bal=__ballot_sync(0xffffffffu,own);
  const unsigned w=threadIdx.x>>5, lane=threadIdx.x&31;
  if(lane==0)wcnt[w]=__popc(bal);
  __syncthreads();
  unsigned base=blockoff[blockIdx.x];
  for(unsigned i=0;i<w;i++)base+=wcnt[i];
  if(own)ownlist[base+__popc(bal&((1u<<lane)-1u))]=p;
}

// owned (normal) particles per cell column, fluid and boundary separately: input of the slab re-balancing
__global__ void ker_mg_column_counts(unsigned n,const unsigned* __restrict__ dcell,const typecode* __restrict__ code,unsigned domcellcode,
  unsigned ncol,unsigned* __restrict__ fluid,unsigned* __restrict__ bound)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=n)return;
  const unsigned rc=code[p], cel=dcell[p];
  if((rc&DSPH_CODE_MASKSPECIAL)!=DSPH_CODE_NORMAL || cel==0xFFFFFFFFu)return;
  const unsigned cx=DcCellx(domcellcode,cel);
  if(cx<ncol)atomicAdd(((rc&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID? fluid: bound)+cx,1u);
}

struct MgPackArrays {
  const unsigned *idp; typecode *code; double *pos; const float4 *velrhop;
  unsigned *dcell; double domposmin[3]; double scell;    // ring: the local copy of a migrant that crossed the periodic face is moved back beyond the face
  const double *pospre; const float4 *aux;
  const float *spstau;      // laminar+SPS: 6 floats per particle appended to the record (record = 88 + 24 bytes)
  unsigned recbytes;        // 88 or 112
};

struct MgBases { unsigned v[MG_CLASSES]; };   // first record slot of each class inside its send buffer

// writes the records of all classes; migrants are converted locally: inside the halo band of the
// neighbour they stay as ghosts (CODE_PERIODIC), otherwise they are dropped by the next sort (CODE_OUTIGNORE)
// DIRECT (transport through peer memory, mgpu_comm.cu): outleft / outright are the receive buffers of the neighbours, mapped into
// this GPU's address space (NVLink peer access / CUDA IPC); the class bases come from the class totals the scan left on the device,
// so no host round trip separates classification and packing.  Thread 0 of block 0 writes the token {4 totals, used} per side that
// travels to the neighbour after this kernel (NCCL, 20 bytes): its arrival tells the neighbour that the records have landed.  When a
// side's records would not fit the neighbour's buffer (capleft / capright records) the kernel changes nothing and sets used=0: both
// ends then fall back to packing into a local buffer and ncclSend/ncclRecv for this exchange.
struct MgDirect { const unsigned *totals; unsigned capleft, capright; unsigned *token; };
template<bool DIRECT> __global__ void __launch_bounds__(MG_BLOCK) ker_mg_pack(MgArgs a,const unsigned char* __restrict__ flags,const unsigned* __restrict__ blockoff,
  MgBases bases,MgPackArrays s,char* __restrict__ outleft,char* __restrict__ outright,MgDirect dr)
{
  __shared__ unsigned wcnt[MG_CLASSES][MG_BLOCK/32];
  if(DIRECT){
    unsigned tot[MG_CLASSES];
    #pragma unroll
    for(int g=0;g<MG_CLASSES;g++)tot[g]=dr.totals[g];
    const unsigned nl=tot[0]+tot[1]+tot[2]+tot[3], nr=tot[4]+tot[5]+tot[6]+tot[7];
    const bool fits=(nl<=dr.capleft && nr<=dr.capright);
    if(blockIdx.x==0 && threadIdx.x==0){
      #pragma unroll
      for(int g=0;g<4;g++){ dr.token[g]=tot[g]; dr.token[5+g]=tot[4+g]; }
      dr.token[4]=dr.token[9]=(fits? 1u: 0u);
    }
    if(!fits)return;
    bases.v[0]=0; bases.v[1]=tot[0]; bases.v[2]=tot[0]+tot[1]; bases.v[3]=tot[0]+tot[1]+tot[2];
    bases.v[4]=0; bases.v[5]=tot[4]; bases.v[6]=tot[4]+tot[5]; bases.v[7]=tot[4]+tot[5]+tot[6];
  }
  const unsigned p=blockIdx.x*MG_BLOCK+threadIdx.x;
  const unsigned f=(p<a.np? flags[p]: 0u);
  const unsigned w=threadIdx.x>>5, lane=threadIdx.x&31;
  unsigned rank[MG_CLASSES];
  #pragma unroll
  for(int g=0;g<MG_CLASSES;g++){
    const unsigned bal=__ballot_sync(0xffffffffu,(f>>g)&1u);
    rank[g]=__popc(bal&((1u<<lane)-1u));
    if(lane==0)wcnt[g][w]=__popc(bal);
  }
  __syncthreads();
  if(!f)return;
  MgRecord r;
  r.pos[0]=s.pos[3*size_t(p)]; r.pos[1]=s.pos[3*size_t(p)+1]; r.pos[2]=s.pos[3*size_t(p)+2];
  { const float4 v=s.velrhop[p]; r.velrhop[0]=v.x; r.velrhop[1]=v.y; r.velrhop[2]=v.z; r.velrhop[3]=v.w; }
  if(s.pospre){ r.pospre[0]=s.pospre[3*size_t(p)]; r.pospre[1]=s.pospre[3*size_t(p)+1]; r.pospre[2]=s.pospre[3*size_t(p)+2]; }
  else{ r.pospre[0]=r.pospre[1]=r.pospre[2]=0; }
  { const float4 v=(s.aux? s.aux[p]: make_float4(0,0,0,0)); r.aux[0]=v.x; r.aux[1]=v.y; r.aux[2]=v.z; r.aux[3]=v.w; }
  r.idp=s.idp[p];
  const unsigned rc=s.code[p];
  #pragma unroll
  for(int g=0;g<MG_CLASSES;g++)if((f>>g)&1u){
    unsigned slot=bases.v[g]+blockoff[size_t(g)*gridDim.x+blockIdx.x]+rank[g];
    for(unsigned i=0;i<w;i++)slot+=wcnt[g][i];
    const int kind=g&3;
    const bool migrant=(kind==1 || kind==2);
    r.code=(migrant? rc: ((rc&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_PERIODIC));   // migrants stay normal, halo copies are ghosts
    char *dst=(g<4? outleft: outright)+size_t(slot)*s.recbytes;
    const bool wrap=(a.ring && (g<4? a.wrapleft: a.wrapright));
    MgRecord rw=r;
    if(wrap && !migrant){          // halo copy through the periodic face: it appears beyond the face on the other side
      const double *sh=(g<4? a.shiftleft: a.shiftright);
      rw.pos[0]+=sh[0]; rw.pos[1]+=sh[1]; rw.pos[2]+=sh[2];
    }
    *reinterpret_cast<MgRecord*>(dst)=rw;
    if(s.spstau){
      float *tau=reinterpret_cast<float*>(dst+sizeof(MgRecord));
      #pragma unroll
      for(int k=0;k<6;k++)tau[k]=s.spstau[6*size_t(p)+k];
    }
    if(migrant){ // local copy becomes a ghost when it is still inside the neighbour's halo band, else it is dropped
      const unsigned cx=DcCellx(a.domcellcode,a.dcell[p]);
      bool band=(g<4? (cx+a.sd>=a.xlo): (cx<a.xhi+a.sd));
      if(wrap){
        // the update wrapped it to the other end of the domain (its position on the new owner); here it stays as the halo copy just
        // beyond the face it crossed: the opposite of the shift that halo copies get on this link
        const double *sh=(g<4? a.shiftleft: a.shiftright);
        const double x=r.pos[0]-sh[0], y=r.pos[1]-sh[1], z=r.pos[2]-sh[2];
        s.pos[3*size_t(p)]=x; s.pos[3*size_t(p)+1]=y; s.pos[3*size_t(p)+2]=z;
        s.dcell[p]=DcCell(a.domcellcode,unsigned((x-s.domposmin[0])/s.scell),unsigned((y-s.domposmin[1])/s.scell),unsigned((z-s.domposmin[2])/s.scell));
        band=true;
      }
      s.code[p]=typecode((rc&(~DSPH_CODE_MASKSPECIAL))|(band? DSPH_CODE_PERIODIC: DSPH_CODE_OUTIGNORE));
    }
  }
}

// direct transport, second half: coalesced 16-byte copies of the packed records of both sides into the neighbours' receive buffers
// (peer memory).  The pack kernel writes 8-byte pieces at an 88-byte stride per lane, which is fine for local HBM but was 5x slower
// than the link when aimed at peer memory (measured at 8 GPUs: 1.0 ms per exchange against 0.17 ms local pack + this kernel).
// token: the words the pack kernel left {to left: 4 totals, used; to right: 4 totals, used}.
__global__ void __launch_bounds__(256) ker_mg_push(const uint4* __restrict__ srcleft,uint4* __restrict__ dstleft,const uint4* __restrict__ srcright,uint4* __restrict__ dstright,
  const unsigned* __restrict__ token,unsigned recbytes)
{
  if(!token[4])return;
  const size_t nl=(size_t(token[0])+token[1]+token[2]+token[3])*recbytes, nr=(size_t(token[5])+token[6]+token[7]+token[8])*recbytes;
  const size_t ql=(nl+15)>>4, qr=(nr+15)>>4;
  const size_t stride=size_t(gridDim.x)*blockDim.x;
  for(size_t i=size_t(blockIdx.x)*blockDim.x+threadIdx.x;i<ql+qr;i+=stride){
    if(i<ql)dstleft[i]=srcleft[i]; else dstright[i-ql]=srcright[i-ql];
  }
}

struct MgUnpackArrays {
  unsigned *idp; typecode *code; unsigned *dcell; double *pos; float4 *velrhop; double *pospre; float4 *aux;
  float *spstau; unsigned recbytes;
};
__global__ void ker_mg_unpack(unsigned n,const char* __restrict__ in,unsigned dstbase,MgUnpackArrays d,DevCtes c){
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n)return;
  const char *src=in+size_t(i)*d.recbytes;
  const MgRecord r=*reinterpret_cast<const MgRecord*>(src);
  if(d.spstau){
    const float *tau=reinterpret_cast<const float*>(src+sizeof(MgRecord));
    #pragma unroll
    for(int k=0;k<6;k++)d.spstau[6*size_t(dstbase+i)+k]=tau[k];
  }
  const unsigned p=dstbase+i;
  d.pos[3*size_t(p)]=r.pos[0]; d.pos[3*size_t(p)+1]=r.pos[1]; d.pos[3*size_t(p)+2]=r.pos[2];
  d.velrhop[p]=make_float4(r.velrhop[0],r.velrhop[1],r.velrhop[2],r.velrhop[3]);
  if(d.pospre){ d.pospre[3*size_t(p)]=r.pospre[0]; d.pospre[3*size_t(p)+1]=r.pospre[1]; d.pospre[3*size_t(p)+2]=r.pospre[2]; }
  if(d.aux)d.aux[p]=make_float4(r.aux[0],r.aux[1],r.aux[2],r.aux[3]);
  d.idp[p]=r.idp; d.code[p]=typecode(r.code);
  unsigned cx=unsigned((r.pos[0]-c.domposmin[0])/c.scell_d), cy=unsigned((r.pos[1]-c.domposmin[1])/c.scell_d), cz=unsigned((r.pos[2]-c.domposmin[2])/c.scell_d);
  d.dcell[p]=DcCell(c.domcellcode,cx,cy,cz);
}

} // namespace dsph
