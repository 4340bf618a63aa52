// bi4.cpp -- see bi4.h.  Independent implementation of the JBinaryData on-disk format.
#include "bi4.h"
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <ctime>

namespace dsph {

static const char *kCodeItem="\nITEM\n", *kCodeValues="\nVALUES", *kCodeArray="\nARRAY";

size_t Bi4TypeSize(int t){
  switch(t){
    case BI4_BOOL: return 4;           // bool is stored as int (S/JBinaryData.h:231)
    case BI4_CHAR: case BI4_UCHAR: return 1;
    case BI4_SHORT: case BI4_USHORT: return 2;
    case BI4_INT: case BI4_UINT: case BI4_FLOAT: return 4;
    case BI4_LLONG: case BI4_ULLONG: case BI4_DOUBLE: return 8;
    case BI4_INT3: case BI4_UINT3: case BI4_FLOAT3: return 12;
    case BI4_DOUBLE3: return 24;
    default: return 0;
  }
}
// Array elements of type bool are 1 byte in the reference (sizeof(bool)); only values use int.
static size_t ArrayElemSize(int t){ return t==BI4_BOOL? 1: Bi4TypeSize(t); }

// ------------------------------------------------------------------ byte sink / source
struct Sink {
  std::vector<unsigned char> b;
  void raw(const void *p,size_t n){ const unsigned char *c=(const unsigned char*)p; b.insert(b.end(),c,c+n); }
  void u32(unsigned v){ raw(&v,4); }
  void i32(int v){ raw(&v,4); }
  void boolean(bool v){ i32(v?1:0); }
  void str(const std::string &s){ u32((unsigned)s.size()); raw(s.data(),s.size()); }
};
struct Source {
  const unsigned char *p; size_t n, c=0;
  void need(size_t k){ if(c+k>n)throw std::runtime_error("bi4: overflow in reading data"); }
  void raw(void *d,size_t k){ need(k); memcpy(d,p+c,k); c+=k; }
  unsigned u32(){ unsigned v; raw(&v,4); return v; }
  int i32(){ int v; raw(&v,4); return v; }
  bool boolean(){ return i32()!=0; }
  std::string str(){ unsigned l=u32(); need(l); std::string s((const char*)p+c,l); c+=l; return s; }
};

// ------------------------------------------------------------------ Bi4Item members
Bi4Value& Bi4Item::Slot(const std::string &n,int type){
  for(auto &v:values)if(v.name==n){ v.type=type; v.text.clear(); memset(v.raw,0,sizeof(v.raw)); return v; }
  values.emplace_back(); values.back().name=n; values.back().type=type; return values.back();
}
void Bi4Item::SetText(const std::string &n,const std::string &v){ Slot(n,BI4_TEXT).text=v; }
void Bi4Item::SetBool(const std::string &n,bool v){ int i=v?1:0; memcpy(Slot(n,BI4_BOOL).raw,&i,4); }
void Bi4Item::SetInt(const std::string &n,int v){ memcpy(Slot(n,BI4_INT).raw,&v,4); }
void Bi4Item::SetUint(const std::string &n,unsigned v){ memcpy(Slot(n,BI4_UINT).raw,&v,4); }
void Bi4Item::SetUllong(const std::string &n,unsigned long long v){ memcpy(Slot(n,BI4_ULLONG).raw,&v,8); }
void Bi4Item::SetLlong(const std::string &n,long long v){ memcpy(Slot(n,BI4_LLONG).raw,&v,8); }
void Bi4Item::SetFloat(const std::string &n,float v){ memcpy(Slot(n,BI4_FLOAT).raw,&v,4); }
void Bi4Item::SetDouble(const std::string &n,double v){ memcpy(Slot(n,BI4_DOUBLE).raw,&v,8); }
void Bi4Item::SetDouble3(const std::string &n,const double v[3]){ memcpy(Slot(n,BI4_DOUBLE3).raw,v,24); }
void Bi4Item::SetFloat3(const std::string &n,const float v[3]){ memcpy(Slot(n,BI4_FLOAT3).raw,v,12); }
void Bi4Item::SetUint3(const std::string &n,const unsigned v[3]){ memcpy(Slot(n,BI4_UINT3).raw,v,12); }

const Bi4Value* Bi4Item::Find(const std::string &n)const{
  for(auto &v:values)if(v.name==n)return &v;
  return nullptr;
}
static const Bi4Value* Need(const Bi4Item *it,const std::string &n,bool optional){
  const Bi4Value *v=it->Find(n);
  if(!v && !optional)throw std::runtime_error("bi4: value '"+n+"' not found in item '"+it->name+"'");
  return v;
}
std::string Bi4Item::GetText(const std::string &n,bool o,const std::string &def)const{ auto v=Need(this,n,o); return v? v->text: def; }
bool Bi4Item::GetBool(const std::string &n,bool o,bool def)const{ auto v=Need(this,n,o); if(!v)return def; int i; memcpy(&i,v->raw,4); return i!=0; }
int Bi4Item::GetInt(const std::string &n,bool o,int def)const{ auto v=Need(this,n,o); if(!v)return def; int i; memcpy(&i,v->raw,4); return i; }
unsigned Bi4Item::GetUint(const std::string &n,bool o,unsigned def)const{ auto v=Need(this,n,o); if(!v)return def; unsigned i; memcpy(&i,v->raw,4); return i; }
unsigned long long Bi4Item::GetUllong(const std::string &n,bool o,unsigned long long def)const{
  auto v=Need(this,n,o); if(!v)return def;
  if(v->type==BI4_UINT||v->type==BI4_INT){ unsigned i; memcpy(&i,v->raw,4); return i; }
  unsigned long long i; memcpy(&i,v->raw,8); return i;
}
double Bi4Item::GetDouble(const std::string &n,bool o,double def)const{
  auto v=Need(this,n,o); if(!v)return def;
  if(v->type==BI4_FLOAT){ float f; memcpy(&f,v->raw,4); return f; }
  double d; memcpy(&d,v->raw,8); return d;
}
void Bi4Item::GetDouble3(const std::string &n,double out[3],bool o)const{
  auto v=Need(this,n,o); if(!v){ out[0]=out[1]=out[2]=0; return; }
  if(v->type==BI4_FLOAT3){ float f[3]; memcpy(f,v->raw,12); out[0]=f[0]; out[1]=f[1]; out[2]=f[2]; }
  else memcpy(out,v->raw,24);
}
Bi4Array& Bi4Item::AddArray(const std::string &n,int type,unsigned count,const void *data){
  arrays.emplace_back(); Bi4Array &a=arrays.back();
  a.name=n; a.type=type; a.count=count;
  const size_t sz=ArrayElemSize(type)*size_t(count);
  a.data.resize(sz); if(sz)memcpy(a.data.data(),data,sz);
  return a;
}
Bi4Array& Bi4Item::AddArrayRef(const std::string &n,int type,unsigned count,const void *data){
  arrays.emplace_back(); Bi4Array &a=arrays.back();
  a.name=n; a.type=type; a.count=count; a.ext=data; a.extsize=ArrayElemSize(type)*size_t(count);
  return a;
}
const Bi4Array* Bi4Item::FindArray(const std::string &n)const{
  for(auto &a:arrays)if(a.name==n)return &a;
  return nullptr;
}
Bi4Item& Bi4Item::AddItem(const std::string &n){
  items.emplace_back(new Bi4Item()); items.back()->name=n; return *items.back();
}
const Bi4Item* Bi4Item::FindItem(const std::string &n)const{
  for(auto &i:items)if(i->name==n)return i.get();
  return nullptr;
}

// ------------------------------------------------------------------ serialisation
static void PutValues(Sink &s,const Bi4Item &it){
  s.str(kCodeValues); s.u32((unsigned)it.values.size());
  for(auto &v:it.values){
    s.str(v.name); s.i32(v.type);
    if(v.type==BI4_TEXT)s.str(v.text);
    else s.raw(v.raw,Bi4TypeSize(v.type));
  }
}
// `all` = false leaves hidden arrays / items (and the values of items with hidevalues) out of the file, and the counts of the
// item record are those of the visible ones (JBinaryData::InItemBase, S/JBinaryData.cpp:853-865).
static void PutItemDef(Sink &s,const Bi4Item &it,unsigned sizevalues,bool all){
  unsigned na=0,ni=0;
  for(auto &a:it.arrays)if(all||!a.hide)na++;
  for(auto &c:it.items)if(all||!c->hide)ni++;
  s.str(kCodeItem); s.str(it.name); s.boolean(it.hide); s.boolean(it.hidevalues);
  s.str(it.fmtfloat); s.str(it.fmtdouble);
  s.u32(na); s.u32(ni); s.u32((all||!it.hidevalues)? sizevalues: 0u);
}
static void PutArrayDef(Sink &s,const Bi4Array &a){
  s.str(kCodeArray); s.str(a.name); s.boolean(a.hide); s.i32(a.type); s.u32(a.count); s.u32((unsigned)a.DataSize());
}
static void WriteItem(FILE *f,const Bi4Item &it,bool all){
  Sink vals; PutValues(vals,it);
  Sink def; PutItemDef(def,it,(unsigned)vals.b.size(),all);
  unsigned sz=(unsigned)def.b.size();
  fwrite(&sz,4,1,f); fwrite(def.b.data(),1,def.b.size(),f);
  if(all||!it.hidevalues)fwrite(vals.b.data(),1,vals.b.size(),f);
  for(auto &a:it.arrays)if(all||!a.hide){
    if(a.type==BI4_TEXT)throw std::runtime_error("bi4: text arrays are not supported");
    if(a.DataSize()>=(size_t(1)<<32))throw std::runtime_error("bi4: array '"+a.name+"' is larger than 4 GB (use several pieces)");
    Sink ad; PutArrayDef(ad,a);
    unsigned asz=(unsigned)ad.b.size();
    fwrite(&asz,4,1,f); fwrite(ad.b.data(),1,ad.b.size(),f);
    if(a.DataSize())fwrite(a.DataPtr(),1,a.DataSize(),f);
  }
  for(auto &c:it.items)if(all||!c->hide)WriteItem(f,*c,all);
}

static void MakeHead(const std::string &filecode,unsigned char head[64]){
  memset(head,0,64);
  std::string t="#FileJBD "+filecode;
  size_t st=t.size()<58? t.size(): 58;
  memcpy(head,t.data(),st);
  for(size_t c=st;c<58;c++)head[c]=' ';
  head[58]='\n'; head[60]=0; // byteorder 0 = little endian
}

void Bi4Save(const std::string &file,const Bi4Item &root,bool all){
  FILE *f=fopen(file.c_str(),"wb");
  if(!f)throw std::runtime_error("bi4: cannot open file for writing: "+file);
  unsigned char head[64]; MakeHead(root.name,head);
  fwrite(head,1,64,f);
  WriteItem(f,root,all);
  const bool bad=ferror(f)!=0;
  const long long fsize=ftell(f);
  fclose(f);
  if(bad)throw std::runtime_error("bi4: file writing failure: "+file);
  if(fsize>=(1ll<<32))throw std::runtime_error("bi4: file is larger than 4 GB (use several pieces): "+file);
}

// JBinaryData::SaveFileListApp (S/JBinaryData.cpp:1555-1571): a list file is the 64-byte header with `filecode`, the parent item
// (written once, when the file is created) and then one item record per call, appended.
void Bi4AppendList(const std::string &file,const std::string &filecode,const Bi4Item &parent,const Bi4Item &item,bool all){
  FILE *f=fopen(file.c_str(),"rb");
  long long fsize=0;
  if(f){
    unsigned char got[64]; const size_t n=fread(got,1,64,f);
    fseek(f,0,SEEK_END); fsize=ftell(f); fclose(f);
    if(fsize){
      unsigned char head[64]; MakeHead(filecode,head);
      if(n<64 || memcmp(got,head,58)!=0 || got[60]!=0)throw std::runtime_error("bi4: the file code is invalid: "+file);
    }
  }
  f=fopen(file.c_str(),fsize? "ab": "wb");
  if(!f)throw std::runtime_error("bi4: cannot open the file: "+file);
  if(!fsize){
    unsigned char head[64]; MakeHead(filecode,head);
    fwrite(head,1,64,f);
    WriteItem(f,parent,all);
  }
  WriteItem(f,item,all);
  const bool bad=ferror(f)!=0;
  const long long fend=ftell(f);
  fclose(f);
  if(bad)throw std::runtime_error("bi4: file writing failure: "+file);
  if(fend>=(1ll<<32))throw std::runtime_error("bi4: file is larger than 4 GB: "+file);
}

static void ReadItem(Source &s,Bi4Item &it){
  const unsigned sizedef=s.u32(); (void)sizedef;
  if(s.str()!=kCodeItem)throw std::runtime_error("bi4: item validation code is invalid");
  it.name=s.str(); it.hide=s.boolean(); it.hidevalues=s.boolean();
  it.fmtfloat=s.str(); it.fmtdouble=s.str();
  const unsigned narrays=s.u32(), nitems=s.u32(), sizevalues=s.u32();
  if(sizevalues){
    if(s.str()!=kCodeValues)throw std::runtime_error("bi4: values validation code is invalid");
    const unsigned num=s.u32();
    for(unsigned c=0;c<num;c++){
      Bi4Value v; v.name=s.str(); v.type=s.i32();
      if(v.type==BI4_TEXT)v.text=s.str();
      else{
        const size_t k=Bi4TypeSize(v.type);
        if(!k)throw std::runtime_error("bi4: type of value invalid");
        s.raw(v.raw,k);
      }
      it.values.push_back(v);
    }
  }
  for(unsigned c=0;c<narrays;c++){
    const unsigned adef=s.u32(); (void)adef;
    if(s.str()!=kCodeArray)throw std::runtime_error("bi4: array validation code is invalid");
    Bi4Array a; a.name=s.str(); a.hide=s.boolean(); a.type=s.i32(); a.count=s.u32();
    const unsigned sizedata=s.u32();
    if(a.type!=BI4_TEXT && sizedata!=ArrayElemSize(a.type)*size_t(a.count))throw std::runtime_error("bi4: size of array data is invalid");
    s.need(sizedata);
    a.data.assign(s.p+s.c,s.p+s.c+sizedata); s.c+=sizedata;
    it.arrays.push_back(std::move(a));
  }
  for(unsigned c=0;c<nitems;c++){
    it.items.emplace_back(new Bi4Item());
    ReadItem(s,*it.items.back());
  }
}

std::unique_ptr<Bi4Item> Bi4Load(const std::string &file,const std::string &filecode){
  FILE *f=fopen(file.c_str(),"rb");
  if(!f)throw std::runtime_error("bi4: cannot open the file: "+file);
  fseek(f,0,SEEK_END); const long long fsize=ftell(f); fseek(f,0,SEEK_SET);
  if(fsize<64){ fclose(f); throw std::runtime_error("bi4: the format file JBinaryData is invalid: "+file); }
  std::vector<unsigned char> buf((size_t)fsize);
  if(fread(buf.data(),1,(size_t)fsize,f)!=(size_t)fsize){ fclose(f); throw std::runtime_error("bi4: read failure: "+file); }
  fclose(f);
  unsigned char head[64]; MakeHead(filecode,head);
  size_t c=0; for(;c<60 && buf[c]==head[c];c++);
  if(c<9)throw std::runtime_error("bi4: the format file JBinaryData is invalid: "+file);
  if(!filecode.empty() && c<50)throw std::runtime_error("bi4: the file code is invalid: "+file);
  if(buf[60]!=0)throw std::runtime_error("bi4: the byte-order in file is invalid: "+file);
  Source s{buf.data()+64,(size_t)fsize-64};
  std::unique_ptr<Bi4Item> root(new Bi4Item());
  ReadItem(s,*root);
  return root;
}

// ------------------------------------------------------------------ PART helpers
std::string PartFileName(unsigned cpart,unsigned piece,unsigned npiece){
  char tx[64];
  if(npiece>1)snprintf(tx,sizeof(tx),"Part_p%02u_%04u.bi4",piece,cpart);
  else snprintf(tx,sizeof(tx),"Part_%04u.bi4",cpart);
  return tx;
}

static std::string DateTime(){
  time_t t=time(nullptr); struct tm tmv; localtime_r(&t,&tmv);
  char tx[64]; strftime(tx,sizeof(tx),"%d-%m-%Y %H:%M:%S",&tmv); return tx;
}

void SavePartFile(const std::string &file,const PartHead &h,const PartData &p,bool posdouble){
  Bi4Item root; root.name="JPartDataBi4";
  // order of values as produced by ConfigBasic/ConfigParticles/ConfigCtes/ConfigSim* (S/JPartDataBi4.cpp:197-290)
  root.SetUint("Piece",h.piece); root.SetUint("Npiece",h.npiece);
  root.SetText("RunCode",h.runcode); root.SetText("Date",DateTime());
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
  char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",p.cpart);
  Bi4Item &part=root.AddItem(nm);
  part.SetUint("Cpart",p.cpart); part.SetDouble("TimeStep",p.timestep);
  part.SetUint("Npok",p.npok); part.SetUint("Nout",p.nout); part.SetUint("Step",p.step);
  part.SetDouble("RunTime",p.runtime);
  part.SetDouble3("DomainMin",p.domainmin); part.SetDouble3("DomainMax",p.domainmax);
  if(p.nptotal)part.SetUllong("NpTotal",p.nptotal);
  if(p.idmax)part.SetUllong("IdMax",p.idmax);
  if(p.symplecticdtpre)part.SetDouble("SymplecticDtPre",p.symplecticdtpre);
  if(p.idp.size()!=p.npok || p.pos.size()!=size_t(p.npok)*3 || p.vel.size()!=size_t(p.npok)*3 || p.rhop.size()!=p.npok)
    throw std::runtime_error("bi4: PART arrays do not match Npok");
  part.AddArray("Idp",BI4_UINT,p.npok,p.idp.data());
  if(posdouble)part.AddArray("Posd",BI4_DOUBLE3,p.npok,p.pos.data());
  else{
    std::vector<float> pf(p.pos.size()); for(size_t c=0;c<pf.size();c++)pf[c]=float(p.pos[c]);
    part.AddArray("Pos",BI4_FLOAT3,p.npok,pf.data());
  }
  part.AddArray("Vel",BI4_FLOAT3,p.npok,p.vel.data());
  part.AddArray("Rhop",BI4_FLOAT,p.npok,p.rhop.data());
  Bi4Save(file,root);
}

void LoadPartFile(const std::string &file,PartHead &h,PartData &p,unsigned cpart){
  std::unique_ptr<Bi4Item> root=Bi4Load(file,"JPartDataBi4");
  h.piece=root->GetUint("Piece"); h.npiece=root->GetUint("Npiece");
  h.runcode=root->GetText("RunCode",true); h.appname=root->GetText("AppName",true); h.casename=root->GetText("CaseName",true);
  h.data2d=root->GetBool("Data2d"); h.data2dposy=root->GetDouble("Data2dPosY",true,0);
  root->GetDouble3("MapPosMin",h.mapposmin,true); root->GetDouble3("MapPosMax",h.mapposmax,true);
  h.perimode=root->GetInt("PeriMode",true,0);
  root->GetDouble3("PeriXinc",h.perixinc,true); root->GetDouble3("PeriYinc",h.periyinc,true); root->GetDouble3("PeriZinc",h.perizinc,true);
  h.casenp=root->GetUllong("CaseNp"); h.casenfixed=root->GetUllong("CaseNfixed"); h.casenmoving=root->GetUllong("CaseNmoving");
  h.casenfloat=root->GetUllong("CaseNfloat"); h.casenfluid=root->GetUllong("CaseNfluid");
  root->GetDouble3("CasePosMin",h.caseposmin,true); root->GetDouble3("CasePosMax",h.caseposmax,true);
  h.npdynamic=root->GetBool("NpDynamic",true,false); h.reuseids=root->GetBool("ReuseIds",true,false);
  h.dp=root->GetDouble("Dp"); h.h=root->GetDouble("H"); h.b=root->GetDouble("B"); h.rhop0=root->GetDouble("Rhop0");
  h.gamma=root->GetDouble("Gamma"); h.massbound=root->GetDouble("MassBound"); h.massfluid=root->GetDouble("MassFluid");
  char nm[32]; snprintf(nm,sizeof(nm),"PART_%04u",cpart);
  const Bi4Item *part=root->FindItem(nm);
  if(!part && cpart==0){ // number unknown to the caller: one file holds one PART item (JPartDataBi4::LoadFilePart)
    for(const auto &it:root->items)if(it->name.compare(0,5,"PART_")==0){ part=it.get(); break; }
  }
  if(!part)throw std::runtime_error("bi4: PART data is invalid: "+file);
  p.cpart=part->GetUint("Cpart"); p.timestep=part->GetDouble("TimeStep");
  p.npok=part->GetUint("Npok"); p.nout=part->GetUint("Nout",true,0); p.step=part->GetUint("Step",true,0);
  p.runtime=part->GetDouble("RunTime",true,0);
  part->GetDouble3("DomainMin",p.domainmin,true); part->GetDouble3("DomainMax",p.domainmax,true);
  p.nptotal=part->GetUllong("NpTotal",true,0); p.idmax=part->GetUllong("IdMax",true,0);
  p.symplecticdtpre=part->GetDouble("SymplecticDtPre",true,0);
  const Bi4Array *aid=part->FindArray("Idp");
  if(!aid || aid->count!=p.npok)throw std::runtime_error("bi4: only Idp (32 bits) is valid");
  p.idp.resize(p.npok); if(p.npok)memcpy(p.idp.data(),aid->data.data(),size_t(p.npok)*4);
  p.pos.resize(size_t(p.npok)*3);
  if(const Bi4Array *ap=part->FindArray("Posd")){
    if(ap->count!=p.npok)throw std::runtime_error("bi4: Posd size is invalid");
    if(p.npok)memcpy(p.pos.data(),ap->data.data(),size_t(p.npok)*24); p.possingle=false;
  }
  else if(const Bi4Array *ap2=part->FindArray("Pos")){
    if(ap2->count!=p.npok)throw std::runtime_error("bi4: Pos size is invalid");
    const float *pf=(const float*)ap2->data.data();
    for(size_t c=0;c<size_t(p.npok)*3;c++)p.pos[c]=pf[c];
    p.possingle=true;
  }
  else throw std::runtime_error("bi4: the position of particles is missing");
  const Bi4Array *av=part->FindArray("Vel"), *ar=part->FindArray("Rhop");
  if(!av || !ar || av->count!=p.npok || ar->count!=p.npok)throw std::runtime_error("bi4: Vel/Rhop arrays are invalid");
  p.vel.resize(size_t(p.npok)*3); p.rhop.resize(p.npok);
  if(p.npok){ memcpy(p.vel.data(),av->data.data(),size_t(p.npok)*12); memcpy(p.rhop.data(),ar->data.data(),size_t(p.npok)*4); }
}

} // namespace dsph
