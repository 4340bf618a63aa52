// partset.cpp -- see partset.h.  Value names and their order are those the reference writes (cited per function), produced
// with this repository's own bi4 writer (bi4.cpp).
#include "partset.h"
#include <cstdio>
#include <cstring>
#include <ctime>
#include <stdexcept>
#include <algorithm>
#include <sys/stat.h>

namespace dsph {

static std::string NowText(){
  time_t t=time(nullptr); struct tm tmv; localtime_r(&t,&tmv);
  char tx[64]; strftime(tx,sizeof(tx),"%d-%m-%Y %H:%M:%S",&tmv); return tx;
}
static bool FileExists(const std::string &f){ struct stat st; return stat(f.c_str(),&st)==0; }
static std::string Pxx(unsigned piece){ char tx[16]; snprintf(tx,sizeof(tx),"_p%02u",piece); return tx; }

void MakePartSetHead(PartSetHead &h,const CaseData &cd,const double mapmin[3],const double mapmax[3],int periactive,
                     const double perixinc[3],const double periyinc[3],const double perizinc[3]){
  const dsph_config &c=cd.cfg;
  h.casename=cd.name; h.data2d=c.simulate2d!=0; h.data2dposy=c.simulate2dposy;
  h.casenp=c.casenp; h.casenfixed=c.casenfixed; h.casenmoving=c.casenmoving; h.casenfloat=c.casenfloat; h.casenfluid=c.casenfluid;
  for(int i=0;i<3;i++){
    h.caseposmin[i]=cd.caseposmin[i]; h.caseposmax[i]=cd.caseposmax[i]; h.mapposmin[i]=mapmin[i]; h.mapposmax[i]=mapmax[i];
    h.perixinc[i]=perixinc[i]; h.periyinc[i]=periyinc[i]; h.perizinc[i]=perizinc[i]; h.gravity[i]=c.gravity[i];
  }
  h.perimode=periactive;
  h.dp=c.dp; h.h=c.kernelh; h.b=c.cteb; h.rhop0=c.rhopzero; h.gamma=c.gamma; h.massbound=c.massbound; h.massfluid=c.massfluid;
  h.viscotype=(c.tvisco==DSPH_VISCO_ARTIFICIAL? 1u: 2u); h.viscovalue=c.visco; h.viscoboundfactor=c.viscoboundfactor;
  h.blocks=cd.blocks;
}

PartSet::PartSet(const std::string &dirout,const PartSetHead &h,unsigned pc,unsigned npc,int axisdiv,unsigned piecemax,bool append)
  :dir(dirout),head(h),piece(pc),npiece(std::max(1u,npc))
{
  if(!dir.empty() && dir.back()!='/')dir+='/';
  if(npiece==1 && piecemax && head.casenp>piecemax){
    npiece=unsigned((head.casenp+piecemax-1)/piecemax); countsplit=true; piece=0;
  }
  head.axisdiv=(countsplit? 0: axisdiv);
  head.npiece=npiece; head.piece=piece;
  if(append){                                  // continue the list files of a previous call / run (restart): keep PartOut's current block
    const unsigned npc=(countsplit? 1u: npiece);
    for(;;){ char tx[32]; snprintf(tx,sizeof(tx),"_%03u.obi4",outblock+1); if(FileExists(dir+"PartOut"+(npc>1? Pxx(piece): std::string())+tx))outblock++; else break; }
    char tx[32]; snprintf(tx,sizeof(tx),"_%03u.obi4",outblock);
    outinitial=FileExists(dir+"PartOut"+(npc>1? Pxx(piece): std::string())+tx);
  }
  else{                                        // a new run: the list files start empty
    const unsigned npc=(countsplit? 1u: npiece);
    for(unsigned pc=0;pc<npiece;pc++)if(countsplit || pc==piece)remove((dir+"PartInfo"+(npiece>1? Pxx(pc): std::string())+".ibi4").c_str());
    for(unsigned b=0;;b++){ char tx[32]; snprintf(tx,sizeof(tx),"_%03u.obi4",b); const std::string f=dir+"PartOut"+(npc>1? Pxx(piece): std::string())+tx; if(!FileExists(f))break; remove(f.c_str()); }
  }
  if(head.runcode=="00000000"){ char tx[32]; snprintf(tx,sizeof(tx),"%08x",unsigned(time(nullptr))^0x5bd1e995u); head.runcode=tx; }
}

// Part_Head.ibi4 -- JPartDataHead::SaveFile (S/JPartDataHead.cpp:208-272)
void PartSet::SaveHead()const{
  Bi4Item b; b.name="JPartDataHead";
  b.SetUint("FmtVersion",180324u);
  b.SetText("AppName",head.appname); b.SetText("Date",NowText()); b.SetText("RunCode",head.runcode); b.SetText("CaseName",head.casename);
  b.SetBool("Data2d",head.data2d); b.SetDouble("Data2dPosY",head.data2dposy);
  b.SetUint("Npiece",npiece); b.SetUint("FirstPart",head.firstpart);
  b.SetDouble3("CasePosMin",head.caseposmin); b.SetDouble3("CasePosMax",head.caseposmax);
  b.SetBool("NpDynamic",head.npdynamic); b.SetBool("ReuseIds",head.reuseids);
  b.SetDouble3("MapPosMin",head.mapposmin); b.SetDouble3("MapPosMax",head.mapposmax);
  b.SetInt("PeriMode",head.perimode);
  b.SetDouble3("PeriXinc",head.perixinc); b.SetDouble3("PeriYinc",head.periyinc); b.SetDouble3("PeriZinc",head.perizinc);
  b.SetUint("ViscoType",head.viscotype); b.SetFloat("ViscoValue",head.viscovalue); b.SetFloat("ViscoBoundFactor",head.viscoboundfactor);
  b.SetBool("Symmetry",head.symmetry); b.SetBool("Splitting",head.splitting);
  b.SetDouble("Dp",head.dp); b.SetDouble("H",head.h); b.SetDouble("B",head.b); b.SetDouble("Gamma",head.gamma); b.SetDouble("RhopZero",head.rhop0);
  b.SetDouble("MassBound",head.massbound); b.SetDouble("MassFluid",head.massfluid); b.SetFloat3("Gravity",head.gravity);
  b.SetUllong("CaseNp",head.casenp); b.SetUllong("CaseNfixed",head.casenfixed); b.SetUllong("CaseNmoving",head.casenmoving);
  b.SetUllong("CaseNfloat",head.casenfloat); b.SetUllong("CaseNfluid",head.casenfluid);
  Bi4Item &mk=b.AddItem("MkBlocks");
  mk.SetUint("Count",unsigned(head.blocks.size()));
  static const char *tname[4]={"Fixed","Moving","Floating","Fluid"};
  for(size_t c=0;c<head.blocks.size();c++){
    const MkBlock &m=head.blocks[c];
    char nm[32]; snprintf(nm,sizeof(nm),"MkBlock_%03u",unsigned(c));
    Bi4Item &it=mk.AddItem(nm);
    it.SetText("Type",tname[std::min(std::max(m.type,0),3)]); it.SetUint("Mk",m.mk); it.SetUint("MkType",m.mktype); it.SetUint("Count",m.count);
  }
  Bi4Save(dir+"Part_Head.ibi4",b,true);
}

// constant values of every PART file of the run -- JPartDataBi4::Config (S/JPartDataBi4.cpp:162-177) + ConfigSimDiv
Bi4Item PartSet::MakeRoot(unsigned pc,unsigned npc)const{
  const PartSetHead &h=head;
  Bi4Item root; root.name="JPartDataBi4";
  root.SetUint("Piece",pc); root.SetUint("Npiece",npc);
  root.SetText("RunCode",h.runcode); root.SetText("Date",NowText());
  root.SetText("AppName",h.appname); root.SetText("CaseName",h.casename);
  root.SetBool("Data2d",h.data2d); root.SetDouble("Data2dPosY",h.data2dposy);
  root.SetDouble3("MapPosMin",h.mapposmin); root.SetDouble3("MapPosMax",h.mapposmax);
  root.SetInt("PeriMode",h.perimode);
  root.SetDouble3("PeriXinc",h.perixinc); root.SetDouble3("PeriYinc",h.periyinc); root.SetDouble3("PeriZinc",h.perizinc);
  root.SetInt("AxisDiv",h.axisdiv);
  root.SetUllong("CaseNp",h.casenp); root.SetUllong("CaseNfixed",h.casenfixed); root.SetUllong("CaseNmoving",h.casenmoving);
  root.SetUllong("CaseNfloat",h.casenfloat); root.SetUllong("CaseNfluid",h.casenfluid);
  root.SetDouble3("CasePosMin",h.caseposmin); root.SetDouble3("CasePosMax",h.caseposmax);
  root.SetBool("NpDynamic",h.npdynamic); root.SetBool("ReuseIds",h.reuseids);
  root.SetDouble("Dp",h.dp); root.SetDouble("H",h.h); root.SetDouble("B",h.b); root.SetDouble("Rhop0",h.rhop0);
  root.SetDouble("Gamma",h.gamma); root.SetDouble("MassBound",h.massbound); root.SetDouble("MassFluid",h.massfluid);
  root.SetBool("Symmetry",h.symmetry); root.SetBool("Splitting",h.splitting);
  return root;
}

static void FillPartItem(Bi4Item &part,const PartMeta &m,unsigned npok,const PartInfoPlus *info){
  part.SetUint("Cpart",m.cpart); part.SetDouble("TimeStep",m.timestep);
  part.SetUint("Npok",npok); part.SetUint("Nout",m.nout); part.SetUint("Step",m.step);
  part.SetDouble("RunTime",m.runtime);
  part.SetDouble3("DomainMin",m.domainmin); part.SetDouble3("DomainMax",m.domainmax);
  if(m.nptotal)part.SetUllong("NpTotal",m.nptotal);
  if(m.symplecticdtpre)part.SetDouble("SymplecticDtPre",m.symplecticdtpre);
  if(info && info->on){                                        // S/JSph.cpp:2598-2617
    part.SetDouble("dtmean",info->dtmean); part.SetDouble("dtmin",info->dtmin); part.SetDouble("dtmax",info->dtmax);
    part.SetDouble("timesim",info->timesim);
    part.SetUint("nct",info->nct); part.SetUint("npbin",info->npbin); part.SetUint("npbout",info->npbout); part.SetUint("npf",info->npf);
    part.SetUint("npbper",info->npbper); part.SetUint("npfper",info->npfper); part.SetUint("newnp",info->newnp);
    part.SetLlong("cpualloc",info->cpualloc);
    part.SetLlong("nctalloc",info->nctalloc); part.SetLlong("nctused",info->nctused);
    part.SetLlong("npalloc",info->npalloc); part.SetLlong("npused",info->npused);
  }
}

void PartSet::SavePart(const PartMeta &m,unsigned npok,const unsigned *idp,const double *pos,const float *vel,const float *rhop,const PartInfoPlus *info){
  char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",m.cpart);
  const unsigned nfiles=(countsplit? npiece: 1u);
  const unsigned per=(countsplit? (npok+npiece-1)/npiece: npok);           // equal pieces (the last one takes the remainder)
  for(unsigned f=0;f<nfiles;f++){
    const unsigned pc=(countsplit? f: piece);
    const unsigned first=std::min(npok,f*per), n=std::min(per,npok-first);
    Bi4Item root=MakeRoot(pc,npiece);
    Bi4Item &part=root.AddItem(nm);
    FillPartItem(part,m,n,info);
    part.AddArrayRef("Idp",BI4_UINT,n,idp+first);
    part.AddArrayRef("Posd",BI4_DOUBLE3,n,pos+size_t(first)*3);
    part.AddArrayRef("Vel",BI4_FLOAT3,n,vel+size_t(first)*3);
    part.AddArrayRef("Rhop",BI4_FLOAT,n,rhop+first);
    Bi4Save(dir+PartFileName(m.cpart,pc,npiece),root,true);
    // PartInfo[_pXX].ibi4 -- JPartDataBi4::SaveFileInfo (S/JPartDataBi4.cpp:445-450): parent with its items hidden (written once),
    // then this PART's item with its arrays hidden
    if(info && info->on){
      Bi4Item parent=MakeRoot(pc,npiece);
      for(auto &a:part.arrays)a.hide=true;
      part.hide=true;
      Bi4AppendList(dir+"PartInfo"+(npiece>1? Pxx(pc): std::string())+".ibi4","JPartDataBi4_Info",parent,part,false);
    }
  }
}

// JPartOutBi4Save::ConfigBasic/ConfigParticles/ConfigLimits (S/JPartOutBi4Save.cpp:120-163)
Bi4Item PartSet::MakeOutRoot()const{
  Bi4Item d; d.name="JPartOutBi4";
  d.SetUint("Piece",piece); d.SetUint("Npiece",countsplit? 1u: npiece);
  d.SetText("RunCode",head.runcode); d.SetText("Date",NowText()); d.SetText("AppName",head.appname);
  d.SetBool("Data2d",head.data2d); d.SetUint("FmtVersion",180320u); d.SetUint("Block",outblock);
  return d;
}

void PartSet::SaveOut(unsigned cpart,double timestep,unsigned nout,const unsigned *idp,const double *pos,const float *vel,const float *rhop,const unsigned char *motive,
                      float rhopmin,float rhopmax){
  const unsigned npc=(countsplit? 1u: npiece);
  auto fname=[&]{ char tx[32]; snprintf(tx,sizeof(tx),"_%03u.obi4",outblock); return dir+"PartOut"+(npc>1? Pxx(piece): std::string())+tx; };
  auto initial=[&]{
    Bi4Item d=MakeOutRoot();
    d.SetDouble3("MapPosMin",head.mapposmin); d.SetDouble3("MapPosMax",head.mapposmax);
    d.SetFloat("RhopMin",rhopmin); d.SetFloat("RhopMax",rhopmax);
    d.SetUllong("CaseNp",head.casenp); d.SetUllong("CaseNfixed",head.casenfixed); d.SetUllong("CaseNmoving",head.casenmoving);
    d.SetUllong("CaseNfloat",head.casenfloat); d.SetUllong("CaseNfluid",head.casenfluid);
    d.SetUint("Block",outblock);
    Bi4Item &p=d.AddItem("Part"); p.hide=true;
    Bi4Save(fname(),d,false);                                   // SaveInitial: header + Data without the (hidden) Part item
    outinitial=true;
    return d;
  };
  if(!outinitial)initial();
  if(!nout)return;
  if(outblocknout>=500000u && outblocknout+nout>2500000u){ outblocknout=0; outblock++; initial(); }   // BLOCKNOUTMIN / BLOCKNOUTMAX
  Bi4Item parent=MakeOutRoot();
  Bi4Item part; char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",cpart); part.name=nm; part.hide=true;
  part.SetUint("Cpart",cpart); part.SetDouble("TimeStep",timestep); part.SetUint("Nout",nout);
  part.AddArrayRef("Idp",BI4_UINT,nout,idp);
  part.AddArrayRef("Posd",BI4_DOUBLE3,nout,pos);
  part.AddArrayRef("Vel",BI4_FLOAT3,nout,vel);
  part.AddArrayRef("Rhop",BI4_FLOAT,nout,rhop);
  part.AddArrayRef("Motive",BI4_UCHAR,nout,motive);
  Bi4AppendList(fname(),"JPartOutBi4",parent,part,true);
  outblocknout+=nout;
}

void LoadPartPieces(const std::string &dir0,unsigned cpart,PartHead &head,PartData &part){
  std::string dir=dir0; if(!dir.empty() && dir.back()!='/')dir+='/';
  if(FileExists(dir+PartFileName(cpart,0,1))){ LoadPartFile(dir+PartFileName(cpart,0,1),head,part,cpart); return; }
  if(!FileExists(dir+PartFileName(cpart,0,2)))throw std::runtime_error("bi4: file of the particles was not found: "+dir+PartFileName(cpart,0,1));
  LoadPartFile(dir+PartFileName(cpart,0,2),head,part,cpart);
  const unsigned npc=head.npiece;
  for(unsigned pc=1;pc<npc;pc++){
    PartHead h2; PartData p2;
    LoadPartFile(dir+PartFileName(cpart,pc,npc),h2,p2,cpart);
    part.idp.insert(part.idp.end(),p2.idp.begin(),p2.idp.end());
    part.pos.insert(part.pos.end(),p2.pos.begin(),p2.pos.end());
    part.vel.insert(part.vel.end(),p2.vel.begin(),p2.vel.end());
    part.rhop.insert(part.rhop.end(),p2.rhop.begin(),p2.rhop.end());
    part.npok+=p2.npok;
    for(int i=0;i<3;i++){ part.domainmin[i]=std::min(part.domainmin[i],p2.domainmin[i]); part.domainmax[i]=std::max(part.domainmax[i],p2.domainmax[i]); }
  }
  // a spatial decomposition stores every slab in its own order: boundary particles must come first for the solver (CheckSortParticles)
  if(head.axisdiv!=0 && npc>1){
    const unsigned nb=unsigned(head.casenfixed+head.casenmoving);
    std::vector<unsigned> order(part.npok);
    unsigned k=0;
    for(unsigned p=0;p<part.npok;p++)if(part.idp[p]<nb)order[k++]=p;
    for(unsigned p=0;p<part.npok;p++)if(part.idp[p]>=nb)order[k++]=p;
    PartData q=part;
    for(unsigned p=0;p<part.npok;p++){
      const unsigned s=order[p];
      q.idp[p]=part.idp[s]; q.rhop[p]=part.rhop[s];
      for(int i=0;i<3;i++){ q.pos[size_t(p)*3+i]=part.pos[size_t(s)*3+i]; q.vel[size_t(p)*3+i]=part.vel[size_t(s)*3+i]; }
    }
    part=std::move(q);
  }
}

} // namespace dsph
