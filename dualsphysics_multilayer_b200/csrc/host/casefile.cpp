// casefile.cpp -- see casefile.h.
#include "casefile.h"
#include <memory>
#include "bi4.h"
#include "xmlmini.h"
#include "sphconst.h"
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <random>
#include <stdexcept>
#include <algorithm>

namespace dsph {

void ConfigDefaults(dsph_config &c){
  memset(&c,0,sizeof(c));
  // defaults of JSph::LoadConfigParameters (S/JSph.cpp:588-757) and JSph::InitVars
  c.tkernel=DSPH_KERNEL_WENDLAND; c.tstep=DSPH_STEP_VERLET; c.tvisco=DSPH_VISCO_ARTIFICIAL;
  c.tdensity=DSPH_DDT_NONE; c.shiftmode=DSPH_SHIFT_NONE; c.cellmode=DSPH_CELLMODE_FULL;
  c.verletsteps=40; c.viscoboundfactor=1.f; c.ddtvalue=0.1f; c.shiftcoef=-2.f; c.shifttfs=0.f;
  c.coefdtmin=0.05f; c.rhopoutmin=700.f; c.rhopoutmax=1300.f; c.partsoutmax=1.f; c.partsoutmax_on=1;
  c.cflnumber=0.2; c.nn_relaxationdt=0.2f;
}

// ------------------------------------------------------------------------------------------------
// JPartsLoad4::CalculeLimits (S/JPartsLoad4.cpp:339-347) + JSph::ResizeMapLimits (S/JSph.cpp:1354-1387)
// with the <simulationdomain> grammar of S/JCaseEParms.cpp:283-320 / JSph::ConfigDomainResize :407-427.
// ------------------------------------------------------------------------------------------------
void ComputeMapRealPos(CaseData &cd){
  dsph_config &c=cd.cfg;
  const double border=double(c.kernelh)*0.05;   // BORDER_MAP, S/DualSphDef.h:130
  const double borderperi=c.dp/2.;
  const int peri[3]={c.perix,c.periy,c.periz};
  double mn[3],mx[3];
  for(int i=0;i<3;i++){
    const double b=(peri[i]? borderperi: border);
    mn[i]=cd.caseposmin[i]-b; mx[i]=cd.caseposmax[i]+b;
  }
  double dmin[3],dmax[3];
  for(int i=0;i<3;i++){
    DomainLimit lmin=cd.posmin[i], lmax=cd.posmax[i];
    if(c.simulate2d && i==1){ lmin=DomainLimit(); lmax=DomainLimit(); }
    const double dif=mx[i]-mn[i];
    const double prcmin=(lmin.mode==2? lmin.value/100: 0), prcmax=(lmax.mode==2? lmax.value/100: 0);
    const double valmin=(lmin.mode==1? lmin.value: 0),     valmax=(lmax.mode==1? lmax.value: 0);
    dmin[i]=mn[i]-dif*prcmin; dmax[i]=mx[i]+dif*prcmax;
    dmin[i]=dmin[i]-valmin;   dmax[i]=dmax[i]+valmax;
    if(lmin.mode==3)dmin[i]=lmin.value;
    if(lmax.mode==3)dmax[i]=lmax.value;
    if(dmin[i]>mn[i] || dmax[i]<mx[i])throw std::runtime_error("Domain limits are not valid.");
    if(peri[i]){ dmin[i]=mn[i]; dmax[i]=mx[i]; }   // periodic axes keep the particle limits
    c.maprealposmin[i]=dmin[i]; c.maprealposmax[i]=dmax[i];
  }
}

void ComputeCodes(const CaseData &cd,unsigned np,const unsigned *idp,dsph_typecode *code){
  // JSphMk: code = type bits | rank of the block among the blocks of the same type (S/JSphMk.cpp:200-230)
  struct Rng{ unsigned begin,end; dsph_typecode code; };
  std::vector<Rng> r; unsigned rank[4]={0,0,0,0};
  static const dsph_typecode tp[4]={DSPH_CODE_TYPE_FIXED,DSPH_CODE_TYPE_MOVING,DSPH_CODE_TYPE_FLOATING,DSPH_CODE_TYPE_FLUID};
  for(auto &b:cd.blocks){ r.push_back({b.begin,b.begin+b.count,dsph_typecode(tp[b.type]|rank[b.type])}); rank[b.type]++; }
  for(unsigned p=0;p<np;p++){
    const unsigned id=idp[p]; dsph_typecode cv=0xffff;
    for(auto &q:r)if(id>=q.begin && id<q.end){ cv=q.code; break; }
    if(cv==0xffff)throw std::runtime_error("Particle id without mk block.");
    code[p]=cv;
  }
}

// ------------------------------------------------------------------------------------------------
static std::string D(double v){ char t[64]; snprintf(t,sizeof(t),"%.17g",v); return t; }
static std::string LimStr(const DomainLimit &l,bool ismin){
  char t[96];
  switch(l.mode){
    case 1: snprintf(t,sizeof(t),"default %c %.17g",ismin?'-':'+',l.value); return t;
    case 2: snprintf(t,sizeof(t),"default %c %.17g%%",ismin?'-':'+',l.value); return t;
    case 3: return D(l.value);
    default: return "default";
  }
}
static DomainLimit ParseLim(std::string v){
  std::string s; for(char ch:v)if(ch!=' '&&ch!='\t')s+=char(tolower(ch));
  DomainLimit l;
  if(s=="default")return l;
  if(s.compare(0,7,"default")==0){
    const bool prc=(!s.empty() && s.back()=='%');
    std::string num=s.substr(8,s.size()-8-(prc?1:0));
    l.mode=(prc? 2: 1); l.value=atof(num.c_str());
    return l;
  }
  l.mode=3; l.value=atof(s.c_str());
  return l;
}

void SaveCase(const CaseData &cd,const std::string &dir,const std::string &casename){
  const dsph_config &c=cd.cfg;
  const std::string fx=dir+"/"+casename+".xml";
  FILE *f=fopen(fx.c_str(),"w");
  if(!f)throw std::runtime_error("Cannot open file for writing: "+fx);
  auto P=[&](const char *k,const std::string &v){ fprintf(f,"      <parameter key=\"%s\" value=\"%s\" />\n",k,v.c_str()); };
  auto I=[&](int v){ return std::to_string(v); };
  fprintf(f,"<?xml version=\"1.0\" encoding=\"UTF-8\" ?>\n<case app=\"dsph_b200 mkcase\">\n  <execution>\n");
  if(c.nn_on){
    fprintf(f,"    <special>\n      <nnphases>\n");
    for(int p=0;p<c.nn_nphases;p++){
      const dsph_phase &ph=c.phases[p];
      fprintf(f,"        <phase mkfluid=\"%u\">\n",ph.mkfluid);
      fprintf(f,"          <rhop value=\"%s\" />\n",D(ph.rho).c_str());
      if(ph.cs0)fprintf(f,"          <csound value=\"%s\" />\n",D(ph.cs0).c_str());
      if(ph.gamma)fprintf(f,"          <gamma value=\"%s\" />\n",D(ph.gamma).c_str());
      fprintf(f,"          <visco value=\"%s\" />\n",D(ph.visco).c_str());
      fprintf(f,"          <tau_yield value=\"%s\" />\n",D(ph.tau_yield).c_str());
      if(ph.tau_max){ fprintf(f,"          <tau_max value=\"%s\" />\n",D(ph.tau_max).c_str()); fprintf(f,"          <Bi_multi value=\"%s\" />\n",D(ph.bi_multi).c_str()); }
      fprintf(f,"          <HBP_m value=\"%s\" />\n",D(ph.m_nn).c_str());
      fprintf(f,"          <HBP_n value=\"%s\" />\n",D(ph.n_nn).c_str());
      fprintf(f,"          <phasetype value=\"%u\" />\n",ph.phasetype);
      fprintf(f,"        </phase>\n");
    }
    fprintf(f,"      </nnphases>\n    </special>\n");
  }
  fprintf(f,"    <parameters>\n");
  P("SavePosDouble","1");
  P("StepAlgorithm",I(c.tstep)); P("VerletSteps",I(c.verletsteps)); P("Kernel",I(c.tkernel));
  if(c.nn_on){
    P("RheologyTreatment","2"); P("VelocityGradientType",I(c.nn_velgrad)); P("RelaxationDt",D(c.nn_relaxationdt));
  }
  P("ViscoTreatment",I(c.tvisco)); P("Visco",D(c.visco)); P("ViscoBoundFactor",D(c.viscoboundfactor));
  if((c.tboundary&0xff)==DSPH_BOUNDARY_MDBC){ P("Boundary","2"); P("SlipMode","1"); P("MDBCCorrector",(c.tboundary&DSPH_BOUNDARY_MDBCCORRECTOR)? "1": "0"); }
  else P("Boundary","1");
  P("DensityDT",I(c.tdensity)); P("DensityDTvalue",D(c.ddtvalue));
  P("Shifting",I(c.shiftmode)); P("ShiftCoef",D(c.shiftcoef)); P("ShiftTFS",D(c.shifttfs));
  P("RigidAlgorithm","1"); P("FtPause","0.0");
  P("CoefDtMin",D(c.coefdtmin)); P("DtIni",D(c.dtini)); P("DtMin",D(c.dtmin));
  P("DtFixed",D(c.dtfixed_on? c.dtfixed: 0)); P("DtFixedFile","NONE"); P("DtAllParticles",I(c.dtallparticles));
  P("TimeMax",D(c.timemax)); P("TimeOut",D(c.timeout));
  P("PartsOutMax",D(c.partsoutmax)); P("RhopOutMin",D(c.rhopoutmin)); P("RhopOutMax",D(c.rhopoutmax));
  if(c.perix){ P("XPeriodicIncY",D(c.perixinc[1])); if(c.perixinc[2]!=0)P("XPeriodicIncZ",D(c.perixinc[2])); }
  if(c.periy){ P("YPeriodicIncX",D(c.periyinc[0])); if(c.periyinc[2]!=0)P("YPeriodicIncZ",D(c.periyinc[2])); }
  if(c.periz){ P("ZPeriodicIncX",D(c.perizinc[0])); if(c.perizinc[1]!=0)P("ZPeriodicIncY",D(c.perizinc[1])); }
  fprintf(f,"      <simulationdomain>\n        <posmin x=\"%s\" y=\"%s\" z=\"%s\" />\n        <posmax x=\"%s\" y=\"%s\" z=\"%s\" />\n      </simulationdomain>\n",
    LimStr(cd.posmin[0],true).c_str(),LimStr(cd.posmin[1],true).c_str(),LimStr(cd.posmin[2],true).c_str(),
    LimStr(cd.posmax[0],false).c_str(),LimStr(cd.posmax[1],false).c_str(),LimStr(cd.posmax[2],false).c_str());
  fprintf(f,"    </parameters>\n");
  const unsigned np=cd.np();
  fprintf(f,"    <particles np=\"%u\" nb=\"%u\" nbf=\"%u\" mkboundfirst=\"%u\" mkfluidfirst=\"%u\">\n",
    np,np-c.casenfluid,c.casenfixed,cd.mkboundfirst,cd.mkfluidfirst);
  static const char *tn[4]={"fixed","moving","floating","fluid"};
  for(auto &b:cd.blocks)
  {
    if(b.type==1)fprintf(f,"      <moving mkbound=\"%u\" mk=\"%u\" begin=\"%u\" count=\"%u\" refmotion=\"%u\" />\n",b.mktype,b.mk,b.begin,b.count,b.refmotion);
    else fprintf(f,"      <%s %s=\"%u\" mk=\"%u\" begin=\"%u\" count=\"%u\" />\n",tn[b.type],(b.type==3? "mkfluid":"mkbound"),b.mktype,b.mk,b.begin,b.count);
  }
  fprintf(f,"    </particles>\n    <constants>\n");
  fprintf(f,"      <data2d value=\"%s\" />\n",c.simulate2d? "true":"false");
  if(c.simulate2d)fprintf(f,"      <data2dposy value=\"%s\" />\n",D(c.simulate2dposy).c_str());
  fprintf(f,"      <gravity x=\"%s\" y=\"%s\" z=\"%s\" />\n",D(c.gravity[0]).c_str(),D(c.gravity[1]).c_str(),D(c.gravity[2]).c_str());
  fprintf(f,"      <cflnumber value=\"%s\" />\n      <gamma value=\"%s\" />\n      <rhop0 value=\"%s\" />\n",D(c.cflnumber).c_str(),D(c.gamma).c_str(),D(c.rhopzero).c_str());
  fprintf(f,"      <dp value=\"%s\" />\n      <h value=\"%s\" />\n      <b value=\"%s\" />\n",D(c.dp).c_str(),D(c.kernelh).c_str(),D(c.cteb).c_str());
  fprintf(f,"      <massbound value=\"%s\" />\n      <massfluid value=\"%s\" />\n",D(c.massbound).c_str(),D(c.massfluid).c_str());
  fprintf(f,"    </constants>\n");
  if(c.casenmoving)cd.motion.SaveXml(f,"    ");
  fprintf(f,"  </execution>\n</case>\n");
  fclose(f);
  if((c.tboundary&0xff)==DSPH_BOUNDARY_MDBC){
    // <name>_Normals.nbi4: the "final normals" file of GenCase >= 5.0.268 (JPartNormalData::SaveFile, S/JPartNormalData.cpp:179-207)
    const unsigned nbound=c.casenfixed+c.casenmoving+c.casenfloat;
    if(cd.normals.size()!=size_t(nbound)*3)throw std::runtime_error("mDBC needs one normal vector per boundary particle (CaseData::normals).");
    Bi4Item b; b.name="JPartNormalData";
    b.SetUint("FmtVersion",220911u); b.SetText("AppName","dsph_b200 mkcase"); b.SetText("Date","-"); b.SetText("CaseName",casename);
    b.SetBool("Data2d",c.simulate2d!=0); b.SetDouble("Data2dPosY",c.simulate2dposy); b.SetDouble("Dp",c.dp); b.SetDouble("H",c.kernelh);
    b.SetDouble("Dist",c.dp*0.5); b.SetText("PartNormalsName","tank"); b.SetUint("Nbound",nbound); b.SetUint("CountNormals",0u);
    b.AddArray("PartNormals",BI4_DOUBLE3,nbound,cd.normals.data());
    Bi4Save(dir+"/"+casename+"_Normals.nbi4",b,true);
  }

  PartHead h; PartData p;
  h.casename=casename; h.data2d=c.simulate2d!=0; h.data2dposy=c.simulate2dposy;
  h.casenp=np; h.casenfixed=c.casenfixed; h.casenmoving=c.casenmoving; h.casenfloat=c.casenfloat; h.casenfluid=c.casenfluid;
  memcpy(h.caseposmin,cd.caseposmin,24); memcpy(h.caseposmax,cd.caseposmax,24);
  h.dp=c.dp; h.h=c.kernelh; h.b=c.cteb; h.rhop0=c.rhopzero; h.gamma=c.gamma; h.massbound=c.massbound; h.massfluid=c.massfluid;
  p.cpart=0; p.npok=np; p.idp=cd.idp; p.pos=cd.pos;
  p.vel.resize(size_t(np)*3); p.rhop.resize(np);
  for(unsigned i=0;i<np;i++){ p.vel[3*i]=cd.velrhop[4*i]; p.vel[3*i+1]=cd.velrhop[4*i+1]; p.vel[3*i+2]=cd.velrhop[4*i+2]; p.rhop[i]=cd.velrhop[4*i+3]; }
  memcpy(p.domainmin,cd.caseposmin,24); memcpy(p.domainmax,cd.caseposmax,24);
  SavePartFile(dir+"/"+casename+".bi4",h,p,true);
}

// only the moving blocks of <particles> and <motion> of Case.xml (no particle file needed)
void LoadCaseMotion(MotionSet &ms,const std::string &dir,const std::string &casename){
  XmlDoc doc=XmlDoc::LoadFile(dir+"/"+casename+".xml");
  const XmlNode *ex=doc.root->Path("case.execution");
  if(!ex)throw std::runtime_error("Cannot find the element 'case.execution'.");
  ms=MotionSet();
  if(const XmlNode *ct=ex->Child("constants"))if(const XmlNode *e=ct->Child("data2d"))ms.simulate2d=e->Bool("value");
  const XmlNode *pt=ex->Child("particles");
  if(!pt)throw std::runtime_error("Cannot find the element 'case.execution.particles'.");
  unsigned nmov=0;
  for(auto &e:pt->children)if(e->name=="moving"){
    MotionObj o; o.ref=int(e->Int("refmotion",true,nmov)); o.mkbound=unsigned(e->Int("mkbound")); o.idbegin=unsigned(e->Int("begin")); o.count=unsigned(e->Int("count"));
    ms.objs.push_back(o); nmov++;
  }
  if(nmov)ms.LoadXml(ex->Child("motion"),dir+"/");
}

void LoadCase(CaseData &cd,const std::string &dir,const std::string &casename){
  XmlDoc doc=XmlDoc::LoadFile(dir+"/"+casename+".xml");
  const XmlNode *ex=doc.root->Path("case.execution");
  if(!ex)throw std::runtime_error("Case configuration <case><execution> was not found.");
  dsph_config &c=cd.cfg; ConfigDefaults(c);
  cd.name=casename;
  // ---- constants (JCaseCtes::ReadXmlRun)
  const XmlNode *ct=ex->Child("constants");
  if(!ct)throw std::runtime_error("The item is not found 'case.execution.constants'.");
  auto CV=[&](const char *n,bool opt=false,double def=0){ const XmlNode *e=ct->Child(n); if(!e){ if(opt)return def; throw std::runtime_error(std::string("xml: <constants><")+n+"> is missing"); } return e->Dbl("value"); };
  { const XmlNode *e=ct->Child("data2d"); c.simulate2d=(e && e->Bool("value"))? 1: 0; }
  c.simulate2dposy=(c.simulate2d? CV("data2dposy"): 0);
  { const XmlNode *g=ct->Child("gravity"); if(!g)throw std::runtime_error("xml: <gravity> is missing");
    c.gravity[0]=float(g->Dbl("x")); c.gravity[1]=float(g->Dbl("y")); c.gravity[2]=float(g->Dbl("z")); }
  c.cflnumber=CV("cflnumber"); c.gamma=float(CV("gamma")); c.rhopzero=float(CV("rhop0")); c.dp=CV("dp");
  c.kernelh=float(CV("h")); c.cteb=float(CV("b")); c.massbound=float(CV("massbound")); c.massfluid=float(CV("massfluid"));
  // ---- parameters (JCaseEParms + JSph::LoadConfigParameters)
  const XmlNode *pr=ex->Child("parameters");
  if(!pr)throw std::runtime_error("The item is not found 'case.execution.parameters'.");
  auto PX=[&](const char *k)->const XmlNode*{ for(auto &e:pr->children)if(e->name=="parameter" && e->Str("key")==k)return e.get(); return nullptr; };
  auto PD=[&](const char *k,bool opt,double def){ const XmlNode *e=PX(k); if(!e){ if(opt)return def; throw std::runtime_error(std::string("xml: parameter '")+k+"' is missing"); } return e->Dbl("value"); };
  auto PI=[&](const char *k,int def){ return int(PD(k,true,def)); };
  c.tkernel=PI("Kernel",2); c.tstep=PI("StepAlgorithm",1); c.verletsteps=PI("VerletSteps",40);
  c.tvisco=PI("ViscoTreatment",1); c.visco=float(PD("Visco",false,0)); c.viscoboundfactor=float(PD("ViscoBoundFactor",true,1));
  { // boundary method (S/JSph.cpp:626-641)
    const int tb=PI("Boundary",1);
    if(tb!=1 && tb!=2)throw std::runtime_error("Boundary Condition method is not valid.");
    c.tboundary=(tb==2? DSPH_BOUNDARY_MDBC: 0);      // 0 and DSPH_BOUNDARY_DBC both mean DBC; 0 keeps the configuration bytes of DBC cases as they were
    if(tb==2){
      if(PI("SlipMode",1)!=1)throw std::runtime_error("Only the slip mode velocity=0 is allowed with mDBC conditions.");   // S/JSph.cpp:788
      if(PI("MDBCCorrector",0)!=0)c.tboundary|=DSPH_BOUNDARY_MDBCCORRECTOR;
    }
  }
  c.tdensity=PI("DensityDT",0); c.ddtvalue=float(PD("DensityDTvalue",true,0.1));
  c.shiftmode=PI("Shifting",0);
  if(c.shiftmode){ c.shiftcoef=float(PD("ShiftCoef",true,-2)); if(c.shiftcoef==0)c.shiftmode=0; else c.shifttfs=float(PD("ShiftTFS",true,0)); }
  c.timemax=PD("TimeMax",false,0); c.timeout=PD("TimeOut",false,0);
  c.dtini=std::max(0.0,PD("DtIni",true,0)); c.dtmin=std::max(0.0,PD("DtMin",true,0));
  c.coefdtmin=float(PD("CoefDtMin",true,0.05)); c.dtallparticles=(PI("DtAllParticles",0)==1);
  c.dtfixed=std::max(0.0,PD("DtFixed",true,0)); c.dtfixed_on=(c.dtfixed>0);
  if(PX("RhopOutMin"))c.rhopoutmin=float(PD("RhopOutMin",true,700));
  if(PX("RhopOutMax"))c.rhopoutmax=float(PD("RhopOutMax",true,1300));
  c.partsoutmax=float(PD("PartsOutMax",true,1));
  if(PX("XPeriodicIncY")){ c.perixinc[1]=PD("XPeriodicIncY",true,0); c.perix=1; }
  if(PX("XPeriodicIncZ")){ c.perixinc[2]=PD("XPeriodicIncZ",true,0); c.perix=1; }
  if(PX("YPeriodicIncX")){ c.periyinc[0]=PD("YPeriodicIncX",true,0); c.periy=1; }
  if(PX("YPeriodicIncZ")){ c.periyinc[2]=PD("YPeriodicIncZ",true,0); c.periy=1; }
  if(PX("ZPeriodicIncX")){ c.perizinc[0]=PD("ZPeriodicIncX",true,0); c.periz=1; }
  if(PX("ZPeriodicIncY")){ c.perizinc[1]=PD("ZPeriodicIncY",true,0); c.periz=1; }
  auto zero3=[](double *v){ v[0]=v[1]=v[2]=0; };
  if(PX("XYPeriodic")){ c.perix=1; c.periy=1; c.periz=0; zero3(c.perixinc); zero3(c.periyinc); zero3(c.perizinc); }
  if(PX("XZPeriodic")){ c.perix=1; c.periy=0; c.periz=1; zero3(c.perixinc); zero3(c.periyinc); zero3(c.perizinc); }
  if(PX("YZPeriodic")){ c.perix=0; c.periy=1; c.periz=1; zero3(c.perixinc); zero3(c.periyinc); zero3(c.perizinc); }
  if(const XmlNode *sd=pr->Child("simulationdomain")){
    const XmlNode *a=sd->Child("posmin"), *b=sd->Child("posmax");
    const char *ax[3]={"x","y","z"};
    for(int i=0;i<3;i++){
      if(a)cd.posmin[i]=ParseLim(a->Str(ax[i],true,"default"));
      if(b)cd.posmax[i]=ParseLim(b->Str(ax[i],true,"default"));
    }
  }
  // ---- multiphase (M/JSph.cpp:608-628,3137-3214)
  if(PX("RheologyTreatment")){
    const int rheo=PI("RheologyTreatment",1);
    if(rheo==2){
      c.nn_on=1; c.nn_velgrad=PI("VelocityGradientType",1); c.nn_relaxationdt=float(PD("RelaxationDt",true,0.2));
    }
  }
  // ---- particles (JCaseParts::ReadXml)
  const XmlNode *pt=ex->Child("particles");
  if(!pt)throw std::runtime_error("Cannot find the element 'case.execution.particles'.");
  cd.mkboundfirst=unsigned(pt->Int("mkboundfirst")); cd.mkfluidfirst=unsigned(pt->Int("mkfluidfirst"));
  cd.blocks.clear();
  unsigned cnt[4]={0,0,0,0}, total=0;
  for(auto &e:pt->children){
    int tp=-1;
    if(e->name=="fixed")tp=0; else if(e->name=="moving")tp=1; else if(e->name=="floating")tp=2; else if(e->name=="fluid")tp=3;
    if(tp<0)continue;
    MkBlock b; b.type=tp; b.mktype=unsigned(e->Int(tp==3? "mkfluid":"mkbound")); b.mk=unsigned(e->Int("mk",true,0));
    b.begin=unsigned(e->Int("begin")); b.count=unsigned(e->Int("count"));
    if(tp==1)b.refmotion=unsigned(e->Int("refmotion",true,0xffffffffll));
    cd.blocks.push_back(b); cnt[tp]+=b.count; total+=b.count;
  }
  if(total!=unsigned(pt->Int("np")) || unsigned(pt->Int("nb"))!=total-cnt[3] || unsigned(pt->Int("nbf"))!=cnt[0])
    throw std::runtime_error("The amount of particles does not match the header.");
  if(cnt[2])throw std::runtime_error("Floating bodies are outside the hot-path scope (SURVEY.md 2.2 row 18).");
  // moving objects in the order of their blocks (JDsMotion::ConfigObjects) + their motion laws
  cd.motion=MotionSet(); cd.motion.simulate2d=(c.simulate2d!=0);
  { unsigned nmov=0;
    for(auto &b:cd.blocks)if(b.type==1){
      if(b.refmotion==0xffffffffu)b.refmotion=nmov;
      MotionObj o; o.ref=int(b.refmotion); o.mkbound=b.mktype; o.idbegin=b.begin; o.count=b.count;
      cd.motion.objs.push_back(o); nmov++;
    }
    if(nmov)cd.motion.LoadXml(ex->Child("motion"),dir+"/");
  }
  c.casenp=total; c.casenfixed=cnt[0]; c.casenmoving=cnt[1]; c.casenfloat=cnt[2]; c.casenfluid=cnt[3];
  if(c.nn_on){
    const XmlNode *nn=ex->Path("special.nnphases");
    if(!nn)throw std::runtime_error("The item is not found 'case.execution.special.nnphases'.");
    auto phs=nn->Children("phase");
    if(phs.empty() || phs.size()>DSPH_MAXPHASES)throw std::runtime_error("The number of phases is invalid.");
    c.nn_nphases=int(phs.size());
    for(size_t i=0;i<phs.size();i++){
      dsph_phase &ph=c.phases[i]; const XmlNode *e=phs[i];
      auto V=[&](const char *n,bool opt){ const XmlNode *q=e->Child(n); if(!q){ if(opt)return 0.0; throw std::runtime_error(std::string("xml: <phase><")+n+"> is missing"); } return q->Dbl("value"); };
      ph.mkfluid=unsigned(e->Int("mkfluid"));
      ph.rho=float(V("rhop",false)); ph.cs0=float(V("csound",true)); ph.gamma=float(V("gamma",true));
      ph.visco=float(V("visco",false)); ph.tau_yield=float(V("tau_yield",false)); ph.tau_max=float(V("tau_max",true));
      if(ph.tau_max)ph.bi_multi=float(V("Bi_multi",false));
      ph.n_nn=float(V("HBP_n",false)); ph.m_nn=float(V("HBP_m",false)); ph.phasetype=unsigned(V("phasetype",false));
    }
    std::sort(c.phases,c.phases+c.nn_nphases,[](const dsph_phase &a,const dsph_phase &b){ return a.mkfluid<b.mkfluid; });
  }
  // ---- gauges (JGaugeSystem::ReadXml / ReadXmlCommon, S/JDsGaugeSystem.cpp:186-275); defaults of JGaugeSystem::ResetCfgDefault (:84-93)
  cd.gauges.clear();
  if(const XmlNode *gs=ex->Path("special.gauges")){
    struct Common { double computedt,computestart,computeend,outputdt,outputstart,outputend; bool output; };
    auto readcommon=[&](const XmlNode *e,Common cm){
      if(e){
        auto V=[&](const char *n,const char *a,double def){ const XmlNode *q=e->Child(n); return (q && q->Has(a))? q->Dbl(a): def; };
        cm.computedt=V("computedt","value",cm.computedt); cm.computestart=V("computetime","start",cm.computestart); cm.computeend=V("computetime","end",cm.computeend);
        if(const XmlNode *q=e->Child("output"))if(q->Has("value"))cm.output=q->Bool("value");
        cm.outputdt=V("outputdt","value",cm.outputdt); cm.outputstart=V("outputtime","start",cm.outputstart); cm.outputend=V("outputtime","end",cm.outputend);
      }
      return cm;
    };
    Common def{c.timeout,0,c.timemax,c.timeout,0,c.timemax,false};
    def=readcommon(gs->Child("default"),def);
    auto P3=[&](const XmlNode *e,const char *n,double *v){ const XmlNode *q=e->Child(n); if(!q)throw std::runtime_error(std::string("xml: <gauges> element <")+n+"> is missing"); v[0]=q->Dbl("x"); v[1]=q->Dbl("y"); v[2]=q->Dbl("z"); };
    for(auto &ep:gs->children){
      const XmlNode *e=ep.get();
      if(e->name.empty() || e->name[0]=='_' || e->name=="default")continue;
      if(e->Has("active") && !e->Bool("active"))continue;
      const Common cm=readcommon(e,def);
      dsph_gauge g; memset(&g,0,sizeof(g));
      const std::string name=e->Str("name");
      for(auto &o:cd.gauges)if(name==o.name)throw std::runtime_error("The gauge name '"+name+"' already exists.");
      strncpy(g.name,name.c_str(),sizeof(g.name)-1);
      g.computedt=cm.computedt; g.computestart=cm.computestart; g.computeend=cm.computeend;
      g.output=cm.output? 1: 0; g.outputdt=cm.outputdt; g.outputstart=cm.outputstart; g.outputend=cm.outputend;
      if(e->name=="velocity"){ g.type=DSPH_GAUGE_VEL; P3(e,"point",g.point0); }
      else if(e->name=="swl"){
        g.type=DSPH_GAUGE_SWL;
        g.masslimit=c.massfluid*(c.simulate2d? 0.4f: 0.5f);
        if(const XmlNode *q=e->Child("masslimit")){ if(q->Has("value"))g.masslimit=float(q->Dbl("value")); else if(q->Has("coef"))g.masslimit=c.massfluid*float(q->Dbl("coef")); }
        if(!(g.masslimit>0))throw std::runtime_error("The masslimit of gauge '"+name+"' is invalid.");
        if(const XmlNode *q=e->Child("pointdp")){ if(q->Has("coefdp"))g.pointdp=c.dp*float(q->Dbl("coefdp")); else g.pointdp=float(q->Dbl("value")); }
        if(!(g.pointdp>0))throw std::runtime_error("The pointdp of gauge '"+name+"' is invalid.");
        P3(e,"point0",g.point0); P3(e,"point2",g.point2);
      }
      else if(e->name=="maxz"){
        g.type=DSPH_GAUGE_MAXZ; P3(e,"point0",g.point0);
        { const XmlNode *q=e->Child("height"); if(!q)throw std::runtime_error("xml: <maxz><height> is missing"); g.height=float(q->Dbl("value")); }
        if(const XmlNode *q=e->Child("distlimit")){
          if(q->Has("coefdp"))g.distlimit=float(c.dp*float(q->Dbl("coefdp")));
          else if(q->Has("coefh"))g.distlimit=float(c.kernelh*float(q->Dbl("coefh")));
          else g.distlimit=float(q->Dbl("value"));
        }
        if(!(g.distlimit>0))throw std::runtime_error("The distlimit of gauge '"+name+"' is invalid.");
      }
      else if(e->name=="force"){
        g.type=DSPH_GAUGE_FORCE;
        const XmlNode *q=e->Child("target"); if(!q)throw std::runtime_error("xml: <force><target> is missing");
        g.mkbound=unsigned(q->Int("mkbound"));
        const MkBlock *mb=nullptr; unsigned rank=0,rk[4]={0,0,0,0};
        for(auto &b:cd.blocks){ if(b.type<3 && b.mktype==g.mkbound && !mb){ mb=&b; rank=rk[b.type]; } rk[b.type]++; }
        if(!mb)throw std::runtime_error("Error loading boundary objects. Mkbound of gauge '"+name+"' is unknown.");
        if(mb->type!=0 && mb->type!=1)throw std::runtime_error("Type of boundary particles of gauge '"+name+"' is invalid. Only fixed or moving particles are allowed.");
        g.idbegin=mb->begin; g.count=mb->count;
        g.code=dsph_typecode((mb->type==0? DSPH_CODE_TYPE_FIXED: DSPH_CODE_TYPE_MOVING)|rank);      // JSphMk: type | index of the block within its type
      }
      else throw std::runtime_error("Gauge type '"+e->name+"' is invalid.");
      cd.gauges.push_back(g);
    }
  }
  // ---- particle data (JPartsLoad4::LoadParticles)
  PartHead h; PartData p;
  LoadPartFile(dir+"/"+casename+".bi4",h,p,0);
  if(h.casenp!=c.casenp || h.casenfixed!=c.casenfixed || h.casenfluid!=c.casenfluid)
    throw std::runtime_error("Particle number does not match the configuration of the case.");
  if(h.data2d!=(c.simulate2d!=0))throw std::runtime_error("Data file does not match the dimension of the case (2D/3D).");
  memcpy(cd.caseposmin,h.caseposmin,24); memcpy(cd.caseposmax,h.caseposmax,24);
  cd.idp=p.idp; cd.pos=p.pos; cd.velrhop.resize(size_t(p.npok)*4);
  for(unsigned i=0;i<p.npok;i++){ cd.velrhop[4*i]=p.vel[3*i]; cd.velrhop[4*i+1]=p.vel[3*i+1]; cd.velrhop[4*i+2]=p.vel[3*i+2]; cd.velrhop[4*i+3]=p.rhop[i]; }
  { // CheckSortParticles: boundary ids first and contiguous (S/JPartsLoad4.cpp:120-128)
    const unsigned nbound=c.casenfixed+c.casenmoving;
    if(nbound){ unsigned last=0; for(unsigned q=0;q<p.npok;q++)if(cd.idp[q]<nbound && q>last)last=q; if(last+1!=nbound)throw std::runtime_error("Order of boundary (fixed and moving) particles is invalid."); }
  }
  if(c.nn_on){ // JSph::LoadMultiphaseData M/JSph.cpp:3247-3263: the fluid starts at the rest density of its phase
    std::vector<dsph_typecode> code(p.npok);
    ComputeCodes(cd,p.npok,cd.idp.data(),code.data());
    for(unsigned i=0;i<p.npok;i++)if((code[i]&DSPH_CODE_MASKTYPE)==DSPH_CODE_TYPE_FLUID){
      const unsigned cp=code[i]&DSPH_CODE_MASKVALUE;
      if(cp>=unsigned(c.nn_nphases))throw std::runtime_error("Fluid particle without phase information...");
      cd.velrhop[4*i+3]=c.phases[cp].rho;
    }
  }
  cd.normals.clear();
  if((c.tboundary&0xff)==DSPH_BOUNDARY_MDBC){
    LoadCaseNormals(cd.normals,dir,casename);
    const unsigned nbound=c.casenfixed+c.casenmoving+c.casenfloat;
    if(cd.normals.empty())throw std::runtime_error("No valid normal vectors for using mDBC (file "+casename+"_Normals.nbi4).");
    if(cd.normals.size()!=size_t(nbound)*3)throw std::runtime_error("The number of final normals does not match boundary particles.");
  }
  const bool mapsizeloaded=(h.mapposmin[0]!=h.mapposmax[0]||h.mapposmin[1]!=h.mapposmax[1]||h.mapposmin[2]!=h.mapposmax[2]);
  if(mapsizeloaded){ memcpy(c.maprealposmin,h.mapposmin,24); memcpy(c.maprealposmax,h.mapposmax,24); }
  else ComputeMapRealPos(cd);
}

// ------------------------------------------------------------------------------------------------
// Synthetic cases.  GenCase conventions: h=coefh*sqrt(3)*dp, cs0=coefsound*sqrt(g*hswl),
// B=cs0^2*rho0/gamma, massfluid=massbound=rho0*dp^3, hydrostatic initial density
// (R/examples/main/01_DamBreak/CaseDambreak_Def.xml:5-15).
// ------------------------------------------------------------------------------------------------
static void DamBreakCounts(double dp,double a,unsigned n[6]){
  n[0]=unsigned(llround(4.0*a/dp));    // tank interior lattice intervals x
  n[1]=unsigned(llround(1.675*a/dp));  // y
  n[2]=unsigned(llround(a/dp));        // z
  n[3]=unsigned(llround(a/dp));        // water column x
  n[4]=n[1]-1;                         // water fills the whole width
  n[5]=unsigned(llround(0.75*a/dp));   // water height
  if(n[3]>n[0]-1)n[3]=n[0]-1;
  if(n[5]>n[2]-1)n[5]=n[2]-1;
}
static unsigned long long DamBreakNp(const unsigned n[6],bool twod=false){
  const unsigned long long NX=n[0],NY=n[1],NZ=n[2];
  if(twod)return (NX+1)+2*NZ+(unsigned long long)n[3]*n[5];   // x-z plane: bed + two side walls + water column
  const unsigned long long bottom=(NX+1)*(NY+1);
  const unsigned long long sides=((NX+1)*(NY+1)-(NX-1)*(NY-1))*NZ;
  return bottom+sides+(unsigned long long)n[3]*n[4]*n[5];
}

// Optional x-slab filter (multi-GPU weak scaling): only particles whose cell column lies in [cxlo,cxhi) are stored; ids,
// counts and constants are those of the full case, so every rank can generate just its own slab.
struct SlabFilter{ bool on=false; unsigned cxlo=0,cxhi=0; };
static thread_local SlabFilter g_slab;

// twod: the x-z section at y=0 as a 2-D simulation (data2d: h=coefh*sqrt(2)*dp, mass=rho0*dp^2, no walls in y)
// nlayers: 1 = single-phase dam break; 2 = two stacked layers (phases 2/3 of the fork's example); 3 = the three stacked layers of
// R/examples/mphase_nnewtonian/01_WetDambreak/CaseWetDambreak2DNN_Def.xml:72-97 (rho 2000/1500/1000, shear-thickening Bingham-like bottom layer)
// moving: 0 none; a one-particle-thick plate right of the water column as moving object 0 with 1 = gate lifted at constant speed, then at
// rest (mvrect -> wait), 2 = flap oscillating about its lower edge (mvrotsinu), 3 = piston oscillating along x (mvrectsinu)
static void MakeDamBreak(CaseData &cd,unsigned long long nptarget,int nlayers,bool twod=false,int moving=0){
  const bool twolayer=(nlayers>=2);
  dsph_config &c=cd.cfg; ConfigDefaults(c);
  const double a=0.4;
  // bisection on dp for the particle count
  double lo=1e-5,hi=0.2; unsigned n[6];
  for(int it=0;it<200;it++){
    const double mid=sqrt(lo*hi); DamBreakCounts(mid,a,n);
    if(n[0]<4||(!twod && n[1]<3)||n[2]<3){ hi=mid; continue; }
    if(DamBreakNp(n,twod)>nptarget)lo=mid; else hi=mid;
  }
  const double dp=hi; DamBreakCounts(dp,a,n);
  const unsigned NX=n[0],NY=(twod? 0u: n[1]),NZ=n[2],FX=n[3],FY=(twod? 0u: n[4]),FZ=n[5];
  const unsigned JF0=(twod? 0u: 1u);       // first y index of the fluid block (2-D: the single layer j=0)
  const double g=9.81, rho0=1000, gamma=7, coefsound=20, coefh=1.0;
  const double hswl=FZ*dp;
  const double cs0=coefsound*sqrt(g*hswl), B=cs0*cs0*rho0/gamma;
  c.dp=dp; c.kernelh=float(coefh*sqrt((twod? 2.0: 3.0)*dp*dp)); c.cteb=float(B); c.gamma=float(gamma); c.rhopzero=float(rho0);
  c.massfluid=float(twod? rho0*dp*dp: rho0*dp*dp*dp); c.massbound=c.massfluid;
  c.simulate2d=(twod? 1: 0); c.simulate2dposy=0;
  c.gravity[0]=0; c.gravity[1]=0; c.gravity[2]=float(-g);
  c.cflnumber=0.2; c.tstep=DSPH_STEP_SYMPLECTIC; c.tkernel=DSPH_KERNEL_WENDLAND;
  c.tvisco=DSPH_VISCO_ARTIFICIAL; c.visco=0.1f; c.viscoboundfactor=1.f;
  c.tdensity=DSPH_DDT_DDT2; c.ddtvalue=0.1f; c.shiftmode=0;
  c.rhopoutmin=700; c.rhopoutmax=1300; c.timemax=1.0; c.timeout=0.05;
  cd.posmax[2].mode=2; cd.posmax[2].value=50;   // <posmax z="default + 50%"> like CaseDambreak_Def.xml
  // particles: boundary first (ids 0..nb-1) then fluid; lattice index*dp, walls at index 0 and N
  std::vector<double> &pos=cd.pos; std::vector<float> &vr=cd.velrhop; std::vector<unsigned> &idp=cd.idp;
  pos.clear(); vr.clear(); idp.clear();
  const double zsurf=FZ*dp;
  auto rhohyd=[&](double z){ const double dz=zsurf-z; if(dz<=0)return float(rho0); return float(rho0*pow(1.0+rho0*g*dz/B,1.0/gamma)); };
  cd.caseposmin[0]=cd.caseposmin[1]=cd.caseposmin[2]=0;
  cd.caseposmax[0]=NX*dp; cd.caseposmax[1]=NY*dp; cd.caseposmax[2]=NZ*dp;
  ComputeMapRealPos(cd);
  SphConst kc; DeriveConstants(c,kc);
  unsigned nextid=0;
  auto add=[&](unsigned i,unsigned j,unsigned k,float rho){
    const unsigned id=nextid++;
    if(g_slab.on){ const unsigned cx=unsigned((i*dp-kc.domposmin[0])/kc.scell); if(cx<g_slab.cxlo || cx>=g_slab.cxhi)return; }
    pos.push_back(i*dp); pos.push_back(j*dp); pos.push_back(k*dp); vr.push_back(0); vr.push_back(0); vr.push_back(0); vr.push_back(rho); idp.push_back(id);
  };
  for(unsigned k=0;k<=NZ;k++)for(unsigned j=0;j<=NY;j++)for(unsigned i=0;i<=NX;i++){
    const bool wall=(k==0||i==0||i==NX||(!twod && (j==0||j==NY)));
    if(wall)add(i,j,k,rhohyd(k*dp));
  }
  const unsigned nbfixed=nextid;
  const unsigned GX=FX+1, GZ=std::min(NZ-1,FZ+2);               // plate: column of lattice nodes at i=GX, k=1..GZ
  if(moving && GX<NX){
    for(unsigned k=1;k<=GZ;k++)for(unsigned j=(twod? 0u: 1u);j<=(twod? 0u: NY-1);j++)add(GX,j,k,rhohyd(k*dp));
  }
  const unsigned nb=nextid;
  unsigned nfl[3]={0,0,0};
  if(!twolayer){
    for(unsigned k=1;k<=FZ;k++)for(unsigned j=JF0;j<=FY;j++)for(unsigned i=1;i<=FX;i++)add(i,j,k,rhohyd(k*dp));
    nfl[0]=nextid-nb;
  }
  else{
    // stacked layers, heaviest at the bottom (fluid density is reset to the phase density at load, M/JSph.cpp:3247)
    static const float layerrho2[2]={1500.f,1000.f}, layerrho3[3]={2000.f,1500.f,1000.f};
    unsigned k0=1;
    for(int l=0;l<nlayers;l++){
      const unsigned k1=(l==nlayers-1? FZ: (FZ*unsigned(l+1))/unsigned(nlayers));
      const unsigned before=nextid;
      for(unsigned k=k0;k<=k1;k++)for(unsigned j=JF0;j<=FY;j++)for(unsigned i=1;i<=FX;i++)add(i,j,k,nlayers==2? layerrho2[l]: layerrho3[l]);
      nfl[l]=nextid-before; k0=k1+1;
    }
  }
  const unsigned nf0=nfl[0], nf1=nfl[1], nf2=nfl[2];
  const unsigned np=nextid;
  c.casenp=np; c.casenfixed=nbfixed; c.casenmoving=nb-nbfixed; c.casenfluid=np-nb;
  cd.mkboundfirst=11; cd.mkfluidfirst=1;
  cd.blocks.clear();
  cd.blocks.push_back({0,0,cd.mkboundfirst+0,0,nbfixed});
  cd.motion=MotionSet(); cd.motion.simulate2d=twod;
  if(nb>nbfixed){
    MkBlock mb{1,1,cd.mkboundfirst+1,nbfixed,nb-nbfixed}; mb.refmotion=0;
    cd.blocks.push_back(mb);
    MotionObj o; o.ref=0; o.mkbound=1; o.idbegin=nbfixed; o.count=nb-nbfixed;
    MotionEvent ev; ev.movid=1; ev.start=0; ev.finish=-1; o.events.push_back(ev);
    MotionMov m; m.id=1;
    if(moving==1){
      m.type=MotionMov::Rect; m.time=0.004; m.nextid=2; m.vel[2]=1.5; o.movs.push_back(m);
      MotionMov w; w.id=2; w.type=MotionMov::Wait; w.time=100; o.movs.push_back(w);
    }
    else if(moving==2){
      m.type=MotionMov::RotSinu; m.time=100; m.freq[0]=2; m.ampl[0]=10; m.phase[0]=0;
      m.axisp1[0]=GX*dp; m.axisp1[1]=0; m.axisp1[2]=dp; m.axisp2[0]=GX*dp; m.axisp2[1]=1; m.axisp2[2]=dp;
      o.movs.push_back(m);
    }
    else{
      m.type=MotionMov::RectSinu; m.time=100; m.freq[0]=2; m.ampl[0]=0.25*a; m.phase[0]=0; o.movs.push_back(m);
    }
    for(auto &q:o.movs)for(int i=0;i<3;i++){ q.iaxisp1[i]=q.axisp1[i]; q.iaxisp2[i]=q.axisp2[i]; q.iref[i]=q.ref[i]; }
    cd.motion.objs.push_back(o);
  }
  cd.blocks.push_back({3,0,cd.mkfluidfirst+0,nb,nf0});
  if(twolayer)cd.blocks.push_back({3,1,cd.mkfluidfirst+1,nb+nf0,nf1});
  if(nlayers==3)cd.blocks.push_back({3,2,cd.mkfluidfirst+2,nb+nf0+nf1,nf2});
  cd.caseposmin[0]=cd.caseposmin[1]=cd.caseposmin[2]=0;
  cd.caseposmax[0]=NX*dp; cd.caseposmax[1]=NY*dp; cd.caseposmax[2]=NZ*dp;
  if(twolayer){
    // phases 1/2 of R/examples/mphase_nnewtonian/01_WetDambreak/CaseWetDambreak2DNN_Def.xml (SURVEY.md 8(d) NN(N))
    c.nn_on=1; c.nn_velgrad=DSPH_VELGRAD_FDA; c.nn_nphases=2; c.nn_relaxationdt=0.2f;
    c.tvisco=DSPH_VISCO_LAMINARSPS; c.visco=0.001f; c.tdensity=DSPH_DDT_DDT2FULL; c.cflnumber=0.1;
    c.shiftmode=DSPH_SHIFT_FULL; c.shiftcoef=-10.f; c.shifttfs=2.75f;
    c.rhopoutmin=700; c.rhopoutmax=2300;
    int q=0;
    if(nlayers==3){ dsph_phase &pb=c.phases[q]; pb.mkfluid=unsigned(q); pb.rho=2000; pb.visco=0.2f; pb.tau_yield=0.0001f; pb.m_nn=100; pb.n_nn=1.5f; pb.gamma=7; pb.cs0=float(cs0); q++; c.rhopoutmin=500; c.rhopoutmax=3000; }
    dsph_phase &p0=c.phases[q]; p0.mkfluid=unsigned(q); p0.rho=1500; p0.visco=0.1f; p0.tau_yield=0.001f; p0.m_nn=10; p0.n_nn=1; p0.gamma=7; p0.cs0=float(cs0); q++;
    dsph_phase &p1=c.phases[q]; p1.mkfluid=unsigned(q); p1.rho=1000; p1.visco=0.05f; p1.tau_yield=0.0005f; p1.m_nn=0; p1.n_nn=1; p1.gamma=7; p1.cs0=float(cs0); q++;
    c.nn_nphases=q;
  }
  ComputeMapRealPos(cd);
}

static void MakePeriodicBox(CaseData &cd,unsigned long long nptarget,unsigned seed){
  dsph_config &c=cd.cfg; ConfigDefaults(c);
  unsigned n=unsigned(llround(cbrt(double(nptarget)))); if(n<8)n=8;
  const double L=1.0, dp=L/n, rho0=1000, gamma=7, cs0=20, coefh=1.0;
  c.dp=dp; c.kernelh=float(coefh*sqrt(3.0*dp*dp)); c.gamma=float(gamma); c.rhopzero=float(rho0);
  c.cteb=float(cs0*cs0*rho0/gamma); c.massfluid=float(rho0*dp*dp*dp); c.massbound=c.massfluid;
  c.gravity[0]=c.gravity[1]=c.gravity[2]=0;
  c.cflnumber=0.2; c.tstep=DSPH_STEP_SYMPLECTIC; c.tkernel=DSPH_KERNEL_WENDLAND; c.tvisco=DSPH_VISCO_ARTIFICIAL; c.visco=0.1f;
  c.tdensity=DSPH_DDT_DDT; c.ddtvalue=0.1f; c.timemax=0.2; c.timeout=0.05;
  c.perix=c.periy=c.periz=1;
  std::mt19937 rng(seed? seed: 12345u);
  std::uniform_real_distribution<double> U(-1.0,1.0);
  const double A=0.1*cs0/10, k=2*3.14159265358979323846/L;
  cd.pos.clear(); cd.velrhop.clear(); cd.idp.clear();
  for(int i=0;i<3;i++){ cd.caseposmin[i]=0.5*dp; cd.caseposmax[i]=(n-0.5)*dp; } // lattice extremes, not jittered positions
  ComputeMapRealPos(cd);
  SphConst kc; DeriveConstants(c,kc);
  unsigned nextid=0;
  for(unsigned kz=0;kz<n;kz++)for(unsigned j=0;j<n;j++)for(unsigned i=0;i<n;i++){
    const double x=(i+0.5)*dp+0.05*dp*U(rng), y=(j+0.5)*dp+0.05*dp*U(rng), z=(kz+0.5)*dp+0.05*dp*U(rng);
    const unsigned id=nextid++;
    if(g_slab.on){ const unsigned cx=unsigned((x-kc.domposmin[0])/kc.scell); if(cx<g_slab.cxlo || cx>=g_slab.cxhi)continue; }   // (the random stream is consumed for every particle: the field does not depend on the slabs)
    cd.pos.push_back(x); cd.pos.push_back(y); cd.pos.push_back(z);
    cd.velrhop.push_back(float( A*sin(k*x)*cos(k*y)*cos(k*z)));
    cd.velrhop.push_back(float(-A*cos(k*x)*sin(k*y)*cos(k*z)));
    cd.velrhop.push_back(0.f);
    cd.velrhop.push_back(float(rho0));
    cd.idp.push_back(id);
  }
  const unsigned np=nextid;
  c.casenp=np; c.casenfixed=0; c.casenfluid=np;
  cd.blocks.clear(); cd.blocks.push_back({3,0,cd.mkfluidfirst+0,0,np});
  for(int i=0;i<3;i++){ cd.caseposmin[i]=0.5*dp; cd.caseposmax[i]=(n-0.5)*dp; } // lattice extremes, not jittered positions
  ComputeMapRealPos(cd);
}

void LoadCaseNormals(std::vector<double> &normals,const std::string &dir,const std::string &casename){
  normals.clear();
  const std::string file=dir+"/"+casename+"_Normals.nbi4";
  FILE *f=fopen(file.c_str(),"rb");
  if(!f)return;
  fclose(f);
  std::unique_ptr<Bi4Item> b=Bi4Load(file,"JPartNormalData");
  const unsigned nbound=b->GetUint("Nbound");
  const Bi4Array *a=b->FindArray("PartNormals");
  if(!a || a->type!=BI4_DOUBLE3 || a->count!=nbound)return;
  normals.resize(size_t(nbound)*3);
  if(nbound)memcpy(normals.data(),a->data.data(),size_t(nbound)*24);
}

void MakeTankNormals(CaseData &cd){
  const dsph_config &c=cd.cfg;
  const unsigned nbound=c.casenfixed+c.casenmoving;
  const double dp=c.dp, hd=0.5*dp;
  const bool twod=(c.simulate2d!=0);
  const long long NX=llround(cd.caseposmax[0]/dp), NY=llround(cd.caseposmax[1]/dp);
  cd.normals.assign(size_t(nbound)*3,0.0);
  for(size_t q=0;q<cd.idp.size();q++){
    const unsigned id=cd.idp[q];
    if(id>=nbound)continue;
    const long long i=llround(cd.pos[3*q]/dp), j=llround(cd.pos[3*q+1]/dp), k=llround(cd.pos[3*q+2]/dp);
    double *n=&cd.normals[size_t(id)*3];
    if(id>=c.casenfixed){ n[0]=-hd; continue; }          // moving plate: the water is on its -x side
    if(i==0)n[0]=hd; else if(i==NX)n[0]=-hd;
    if(!twod){ if(j==0)n[1]=hd; else if(j==NY)n[1]=-hd; }
    if(k==0)n[2]=hd;
  }
}

void MakeCaseSlab(CaseData &cd,const std::string &kind,unsigned long long nptarget,unsigned seed,unsigned cxlo,unsigned cxhi){
  if(kind!="dambreak" && kind!="nn2layer" && kind!="periodicbox")throw std::runtime_error("Slab generation is available for dambreak, nn2layer and periodicbox.");
  g_slab.on=true; g_slab.cxlo=cxlo; g_slab.cxhi=cxhi;
  try{ MakeCase(cd,kind,nptarget,seed); }catch(...){ g_slab.on=false; throw; }
  g_slab.on=false;
}

void MakeCase(CaseData &cd,const std::string &kind,unsigned long long nptarget,unsigned seed){
  cd=CaseData();
  if(kind=="dambreak"){ MakeDamBreak(cd,nptarget,1); cd.name="DamBreak"; }
  else if(kind=="nn2layer"){ MakeDamBreak(cd,nptarget,2); cd.name="DamBreakNN"; }
  else if(kind=="nn3layer"){ MakeDamBreak(cd,nptarget,3); cd.name="DamBreakNN3"; }
  else if(kind=="nn3layer2d"){ MakeDamBreak(cd,nptarget,3,true); cd.name="DamBreakNN3_2D"; }
  else if(kind=="dambreak2d"){ MakeDamBreak(cd,nptarget,1,true); cd.name="DamBreak2D"; }
  else if(kind=="movgate"){ MakeDamBreak(cd,nptarget,1,false,1); cd.name="DamBreakGate"; }
  else if(kind=="movflap"){ MakeDamBreak(cd,nptarget,1,false,2); cd.name="DamBreakFlap"; }
  else if(kind=="movpiston"){ MakeDamBreak(cd,nptarget,1,false,3); cd.name="DamBreakPiston"; }
  else if(kind=="movflap2d"){ MakeDamBreak(cd,nptarget,1,true,2); cd.name="DamBreakFlap2D"; }
  else if(kind=="nn2layer2d"){ MakeDamBreak(cd,nptarget,2,true); cd.name="DamBreakNN2D"; }
  else if(kind=="periodicbox"){ MakePeriodicBox(cd,nptarget,seed); cd.name="PeriodicBox"; }
  else throw std::runtime_error("Unknown case kind: "+kind+" (dambreak | periodicbox | nn2layer | nn3layer | dambreak2d | nn2layer2d | nn3layer2d)");
}

} // namespace dsph
