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
};

// classes per particle and direction: bit0 send-left, bit1 send-right; kind: 0 bound ghost, 1 fluid migrant, 2 fluid ghost
__device__ __forceinline__ void MgClassify(unsigned p,const MgArgs &a,int &toleft,int &toright,bool &migl,bool &migr){
  toleft=toright=-1; migl=migr=false;
  const unsigned rc=a.code[p];
  if((rc&DSPH_CODE_MASKSPECIAL)!=DSPH_CODE_NORMAL)return;
  const unsigned cel=a.dcell[p];
  if(cel==0xFFFFFFFFu)return;
  const unsigned cx=DcCellx(a.domcellcode,cel);
  const bool fluid=((rc&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID);
  if(a.hasleft){
    if(fluid && cx<a.xlo){ toleft=1; migl=true; }
    else if(cx<a.xlo+a.sd && cx>=a.xlo)toleft=(fluid? 2: 0);
  }
  if(a.hasright){
    if(fluid && cx>=a.xhi){ toright=1; migr=true; }
    else if(cx+a.sd>=a.xhi && cx<a.xhi)toright=(fluid? 2: 0);
  }
}

// counts[dir*3+kind] per particle are scanned separately: 6 count arrays of np entries would be heavy; instead each
// particle writes a packed 6 x 1-bit word and six scans run over extracted bits (np is per-GPU, the pass is cheap).
__global__ void ker_mg_flags(MgArgs a,unsigned* __restrict__ flags){
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.np)return;
  int tl,tr; bool ml,mr; MgClassify(p,a,tl,tr,ml,mr);
  unsigned f=0;
  if(tl>=0)f|=1u<<tl;
  if(tr>=0)f|=1u<<(3+tr);
  flags[p]=f;
}
__global__ void ker_mg_bit(unsigned n,const unsigned* __restrict__ flags,int bit,unsigned* __restrict__ out){
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p<n)out[p]=(flags[p]>>bit)&1u;
}

struct MgPackArrays {
  const unsigned *idp; typecode *code; const double *pos; const float4 *velrhop;
  const double *pospre; const float4 *aux;
};

// writes the records of one (direction,kind) group; migrants are converted locally: inside the halo band of the
// neighbour they stay as ghosts (CODE_PERIODIC), otherwise they are dropped by the next sort (CODE_OUTIGNORE)
__global__ void ker_mg_pack(MgArgs a,const unsigned* __restrict__ flags,int bit,const unsigned* __restrict__ offset,unsigned base,
  MgPackArrays s,MgRecord* __restrict__ out)
{
  const unsigned p=blockIdx.x*blockDim.x+threadIdx.x;
  if(p>=a.np)return;
  if(!((flags[p]>>bit)&1u))return;
  MgRecord r;
  r.pos[0]=s.pos[3*size_t(p)]; r.pos[1]=s.pos[3*size_t(p)+1]; r.pos[2]=s.pos[3*size_t(p)+2];
  { const float4 v=s.velrhop[p]; r.velrhop[0]=v.x; r.velrhop[1]=v.y; r.velrhop[2]=v.z; r.velrhop[3]=v.w; }
  if(s.pospre){ r.pospre[0]=s.pospre[3*size_t(p)]; r.pospre[1]=s.pospre[3*size_t(p)+1]; r.pospre[2]=s.pospre[3*size_t(p)+2]; }
  else{ r.pospre[0]=r.pospre[1]=r.pospre[2]=0; }
  { const float4 v=(s.aux? s.aux[p]: make_float4(0,0,0,0)); r.aux[0]=v.x; r.aux[1]=v.y; r.aux[2]=v.z; r.aux[3]=v.w; }
  r.idp=s.idp[p];
  const unsigned rc=s.code[p];
  const int kind=bit%3;
  r.code=(kind==1? rc: ((rc&(~DSPH_CODE_MASKSPECIAL))|DSPH_CODE_PERIODIC));   // migrants stay normal, halo copies are ghosts
  out[base+offset[p]]=r;
  if(kind==1){ // migrant: local copy becomes a ghost when it is still inside the neighbour's halo band, else it is dropped
    const unsigned cx=DcCellx(a.domcellcode,a.dcell[p]);
    const bool band=(bit<3? (cx+a.sd>=a.xlo): (cx<a.xhi+a.sd));
    s.code[p]=typecode((rc&(~DSPH_CODE_MASKSPECIAL))|(band? DSPH_CODE_PERIODIC: DSPH_CODE_OUTIGNORE));
  }
}

struct MgUnpackArrays {
  unsigned *idp; typecode *code; unsigned *dcell; double *pos; float4 *velrhop; double *pospre; float4 *aux;
};
__global__ void ker_mg_unpack(unsigned n,const MgRecord* __restrict__ in,unsigned dstbase,MgUnpackArrays d,DevCtes c){
  const unsigned i=blockIdx.x*blockDim.x+threadIdx.x;
  if(i>=n)return;
  const MgRecord r=in[i];
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
