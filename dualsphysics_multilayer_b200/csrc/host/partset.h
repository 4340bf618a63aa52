// partset.h -- the set of output files one run leaves in its output directory, in the reference's formats:
//   Part_Head.ibi4                  constants of the run                       S/JPartDataHead.cpp:208-272
//   Part_NNNN.bi4                   particles of one PART                      S/JPartDataBi4.cpp:98-102,300-371,437-439
//   Part_pXX_NNNN.bi4               the same in several pieces: one per slab of a multi-GPU run (AxisDiv = DIV_X,
//                                   S/JPartDataBi4.h:81,122-133), or count-split on one GPU when a single file would pass the
//                                   4 GB limit of the container (32-bit sizes, S/JBinaryData.cpp:1223-1235)
//   PartInfo.ibi4 / PartInfo_pXX    one record per PART without the arrays     S/JPartDataBi4.cpp:445-450, S/JSph.cpp:2596-2630
//   PartOut_BBB.obi4                excluded particles with the motive         S/JPartOutBi4Save.cpp:103-231, S/JSph.cpp:2693-2696
// Host C++ only.
#pragma once
#include "bi4.h"
#include "casefile.h"
#include <string>
#include <vector>

namespace dsph {

struct PartSetHead : public PartHead {
  float gravity[3]={0,0,0};
  unsigned viscotype=0; float viscovalue=0, viscoboundfactor=1;   // JPartDataHead::TpVisco: 1 artificial, 2 laminar+SPS
  std::vector<MkBlock> blocks;
  unsigned firstpart=0;
};
// fills the head from the case constants (what JSph::ConfigSaveData hands to JPartDataHead, S/JSph.cpp:2303-2350)
void MakePartSetHead(PartSetHead &h,const CaseData &cd,const double maprealposmin[3],const double maprealposmax[3],int periactive,
                     const double perixinc[3],const double periyinc[3],const double perizinc[3]);

// run-dependent values of one PART (S/JSph.cpp:2596-2630): the SDAT_Info block
struct PartInfoPlus {
  bool on=false;
  double dtmean=0,dtmin=0,dtmax=0,timesim=0;
  unsigned nct=0,npbin=0,npbout=0,npf=0,npbper=0,npfper=0,newnp=0;
  long long cpualloc=0,nctalloc=0,nctused=0,npalloc=0,npused=0;
};

struct PartMeta {
  unsigned cpart=0,step=0,nout=0; double timestep=0,runtime=0,symplecticdtpre=0;
  double domainmin[3]={0,0,0},domainmax[3]={0,0,0};
  unsigned long long nptotal=0;
};

class PartSet {
  std::string dir; PartSetHead head;
  unsigned piece=0,npiece=1; bool countsplit=false;
  // PartOut state (JPartOutBi4Save: a new block file when the current one holds >= 0.5 M and would pass 2.5 M particles)
  unsigned outblock=0,outblocknout=0; bool outinitial=false;
  Bi4Item MakeRoot(unsigned pc,unsigned npc)const;
  Bi4Item MakeOutRoot()const;
 public:
  static const unsigned PIECE_MAX_DEFAULT=80000000u;   // particles per piece: 48 B each -> 3.84 GB per file
  // piece/npiece given (npiece>1): this object writes ONE piece of a spatial decomposition (axisdiv 1=x).
  // npiece==1: one file per PART, or ceil(casenp/piecemax) count-split pieces when casenp>piecemax.
  // append: continue PartInfo / PartOut files that already exist in dirout (restart, or repeated host-side calls); otherwise they
  // are started anew
  PartSet(const std::string &dirout,const PartSetHead &h,unsigned piece,unsigned npiece,int axisdiv,unsigned piecemax=PIECE_MAX_DEFAULT,bool append=false);
  unsigned Npiece()const{ return npiece; }
  void SaveHead()const;   // Part_Head.ibi4 (call on piece 0 only)
  // particle arrays of this piece (halo / periodic copies already removed); vel packed xyz
  void SavePart(const PartMeta &m,unsigned npok,const unsigned *idp,const double *posxyz,const float *velxyz,const float *rhop,const PartInfoPlus *info);
  // excluded particles since the last PART; motive 1 position, 2 density, 3 velocity (S/JSph.cpp:2496-2501)
  void SaveOut(unsigned cpart,double timestep,unsigned nout,const unsigned *idp,const double *posxyz,const float *velxyz,const float *rhop,const unsigned char *motive,
               float rhopmin,float rhopmax);
};

// Loads every piece of PART `cpart` of `dir` (Part_NNNN.bi4 or Part_p00_NNNN.bi4 ...) and concatenates them in piece order.
void LoadPartPieces(const std::string &dir,unsigned cpart,PartHead &head,PartData &part);

} // namespace dsph
