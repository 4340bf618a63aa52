// motion.cpp -- see motion.h.  Restates the time stepping of the reference's motion objects for one level of objects:
// event activation (S/JMotion.cpp:446-470), per-step displacement of every active movement with the split of a step between chained
// movements (S/JMotionObj.cpp:389-560), movement hand-over (JMotionMovActive::NextMov S/JMotionObj.cpp:186-208), composition into a
// translation or a 4x4 matrix (S/JMotionPos.cpp), rotation about an axis (JMatrix4::MatrixRot S/JMatrix4.h:332-366) and the
// per-object result velocity = displacement / dt (JMotionListData::Sp_Movedt S/JMotionList.cpp:43-60).
#include "motion.h"
#include "xmlmini.h"
#include <cmath>
#include <cstring>
#include <cstdio>
#include <stdexcept>

namespace dsph {

namespace {
const double kPi=3.14159265358979323846;
const double kToRad=kPi/180.0, kToDeg=180.0/kPi;

struct Mat4 {
  double a[16];
  Mat4(){ Identity(); }
  void Identity(){ for(int i=0;i<16;i++)a[i]=(i%5==0? 1.0: 0.0); }
  double& operator()(int r,int c){ return a[4*r+c]; }
  double operator()(int r,int c)const{ return a[4*r+c]; }
  void MulRight(const Mat4 &m){               // this = this * m
    Mat4 r;
    for(int i=0;i<4;i++)for(int j=0;j<4;j++){ double s=0; for(int k=0;k<4;k++)s+=(*this)(i,k)*m(k,j); r(i,j)=s; }
    *this=r;
  }
  static Mat4 Move(const double d[3]){ Mat4 m; m(0,3)=d[0]; m(1,3)=d[1]; m(2,3)=d[2]; return m; }
  // rotation of `ang` degrees about the axis p1->p2: move p1 to the origin, turn the axis onto z (about x, then about y), rotate, undo
  static Mat4 Rot(double ang,const double p1[3],const double p2[3]){
    const double rad=ang*kToRad;
    const double v[3]={p2[0]-p1[0],p2[1]-p1[1],p2[2]-p1[2]};
    const double L=std::sqrt(v[0]*v[0]+v[1]*v[1]+v[2]*v[2]);
    const double L1=std::sqrt(v[1]*v[1]+v[2]*v[2]);
    Mat4 t; t(0,3)=p1[0]; t(1,3)=p1[1]; t(2,3)=p1[2];
    Mat4 tm1; tm1(0,3)=-p1[0]; tm1(1,3)=-p1[1]; tm1(2,3)=-p1[2];
    Mat4 rx,rxm1;
    if(L1==0){ rx(1,1)=0; rx(1,2)=1; rx(2,1)=-1; rx(2,2)=0;  rxm1(1,1)=0; rxm1(1,2)=-1; rxm1(2,1)=1; rxm1(2,2)=0; }
    else{ rx(1,1)=v[2]/L1; rx(1,2)=v[1]/L1; rx(2,1)=-v[1]/L1; rx(2,2)=v[2]/L1;  rxm1(1,1)=v[2]/L1; rxm1(1,2)=-v[1]/L1; rxm1(2,1)=v[1]/L1; rxm1(2,2)=v[2]/L1; }
    Mat4 ry,rym1;
    ry(0,0)=L1/L; ry(0,2)=v[0]/L; ry(2,0)=-v[0]/L; ry(2,2)=L1/L;
    rym1(0,0)=L1/L; rym1(0,2)=-v[0]/L; rym1(2,0)=v[0]/L; rym1(2,2)=L1/L;
    const double cs=std::cos(rad), sn=std::sin(rad);
    Mat4 rz; rz(0,0)=cs; rz(0,1)=sn; rz(1,0)=-sn; rz(1,1)=cs;
    t.MulRight(rx); t.MulRight(ry); t.MulRight(rz); t.MulRight(rym1); t.MulRight(rxm1); t.MulRight(tm1);
    return t;
  }
};

// JMotionPos: a translation until the first rotation, then a matrix
struct ModPos {
  bool simple=true; double s[3]={0,0,0}; Mat4 m;
  void Move(const double d[3]){ if(simple){ s[0]+=d[0]; s[1]+=d[1]; s[2]+=d[2]; } else m.MulRight(Mat4::Move(d)); }
  void Rotate(double ang,const double p1[3],const double p2[3]){ if(simple){ m=Mat4::Move(s); simple=false; } m.MulRight(Mat4::Rot(ang,p1,p2)); }
  // JMotionPos::MoveMix / PointMove (S/JMotionPos.cpp:82-107): the motion of a parent object applied to a child, to a point
  void MoveMix(const ModPos &o){ if(simple && !o.simple){ m=Mat4::Move(s); simple=false; } if(o.simple)Move(o.s); else m.MulRight(o.m); }
  void PointMove(double p[3])const{
    if(simple){ p[0]+=s[0]; p[1]+=s[1]; p[2]+=s[2]; return; }
    const double x=p[0],y=p[1],z=p[2];
    for(int i=0;i<3;i++)p[i]=m(i,0)*x+m(i,1)*y+m(i,2)*z+m(i,3);
  }
  void Store(MotionObj &o)const{ o.mpsimple=simple; for(int i=0;i<3;i++)o.mps[i]=s[i]; memcpy(o.mpm,m.a,sizeof(o.mpm)); }
  static ModPos Of(const MotionObj &o){ ModPos r; r.simple=o.mpsimple; for(int i=0;i<3;i++)r.s[i]=o.mps[i]; memcpy(r.m.a,o.mpm,sizeof(o.mpm)); return r; }
};

void ConfigActive(MotionActive &a){
  for(int i=0;i<3;i++){ a.vel[i]=0; a.phase[i]=0; } a.velang=0; a.phaseuni=0;
  const MotionMov &m=*a.mov;
  switch(m.type){
    case MotionMov::Rect: case MotionMov::RectAce: for(int i=0;i<3;i++)a.vel[i]=m.vel[i]; break;
    case MotionMov::Rot: case MotionMov::RotAce: a.velang=m.velang; break;
    case MotionMov::RectSinu: for(int i=0;i<3;i++)a.phase[i]=m.phase[i]; break;
    case MotionMov::RotSinu: a.phaseuni=m.phase[0]; break;
    case MotionMov::Cir: case MotionMov::CirAce: a.velang=m.velang; break;      // (the travelling point keeps its state in the movement, like the
    case MotionMov::CirSinu: a.phaseuni=m.phase[0]; break;                      //  reference's JMotionAxis: it is not reset by an activation)
    case MotionMov::RectFile: a.dfindex=0; for(int i=0;i<3;i++)a.dflastpos[i]=m.fvals[i]; break;       // JMotionMovActive::DfConfig S/JMotionObj.cpp:96-116
    case MotionMov::RotFile: a.dfindex=0; a.dflastang=m.fvals[0]; break;
    default: break;
  }
}
// position in times[] of the closest value below t, as the reference finds it (JMotionMovActive::BinarySearch S/JMotionObj.cpp:122-145)
unsigned DfBinarySearch(const std::vector<double> &times,double t){
  const unsigned size=unsigned(times.size());
  unsigned ret=0;
  if(size>1){
    int ccen=0,cmin=0,cmax=int(size)-1;
    while(cmin<=cmax){
      ccen=((cmax-cmin)/2)+cmin;
      if(times[ccen]==t)cmin=cmax+1;
      else if(t<times[ccen])cmax=ccen-1;
      else cmin=ccen+1;
    }
    if(ccen && ccen<int(size) && t<times[ccen])ccen--;
    ret=unsigned(ccen);
    while(ret && t==times[ret-1])ret--;
  }
  return ret;
}
// linear interpolation of the data file at t; the interval index only moves forward (DfGetNewPos / DfGetNewAng S/JMotionObj.cpp:150-183)
void DfQuery(const MotionMov &m,MotionActive &a,double t,int ncomp,double *out){
  const unsigned count=unsigned(m.ftimes.size());
  if(a.dfindex==0)a.dfindex=DfBinarySearch(m.ftimes,t);
  while(a.dfindex<count && t>m.ftimes[a.dfindex])a.dfindex++;
  if(a.dfindex>=count){ for(int c=0;c<ncomp;c++)out[c]=m.fvals[size_t(count-1)*ncomp+c]; return; }
  const unsigned i1=a.dfindex, i0=(i1? i1-1: 0);
  const double tfactor=(i1? (t-m.ftimes[i0])/(m.ftimes[i1]-m.ftimes[i0]): 0.0);
  for(int c=0;c<ncomp;c++){ const double v0=m.fvals[size_t(i0)*ncomp+c], v1=m.fvals[size_t(i1)*ncomp+c]; out[c]=v0+tfactor*(v1-v0); }
}
// rows of a text data file: lines starting with # are remarks, values separated by blanks, tabs, ; or , (JReadDatafile)
std::vector<std::vector<double>> ReadDataFile(const std::string &file){
  FILE *f=fopen(file.c_str(),"r");
  if(!f)throw std::runtime_error("motion: cannot open the data file "+file);
  std::vector<std::vector<double>> rows; char line[4096];
  while(fgets(line,sizeof(line),f)){
    char *p=line; while(*p==' '||*p=='\t')p++;
    if(*p=='#'||*p=='\n'||*p=='\r'||*p==0)continue;
    std::vector<double> v;
    while(*p){
      while(*p==' '||*p=='\t'||*p==';'||*p==','||*p=='\n'||*p=='\r')p++;
      if(!*p)break;
      char *e=nullptr; const double x=strtod(p,&e);
      if(e==p)break;
      v.push_back(x); p=e;
    }
    if(!v.empty())rows.push_back(v);
  }
  fclose(f);
  return rows;
}
const MotionMov* FindMov(const MotionObj &o,int id){ for(auto &m:o.movs)if(m.id==id)return &m; return nullptr; }

void NextMov(const MotionObj &o,MotionActive &a){
  const MotionMov *nx=FindMov(o,a.mov->nextid);
  if(!nx)return;
  if(!a.flash)a.start+=a.mov->time;
  a.mov=nx;
  a.flash=(nx->time<0);
  a.finish=(a.flash? a.start: a.start+nx->time);
  if(a.eventfinish>=0 && a.eventfinish<a.finish)a.finish=a.eventfinish;
  double velp[3]={a.vel[0],a.vel[1],a.vel[2]}, php[3]={a.phase[0],a.phase[1],a.phase[2]}; const double velangp=a.velang, phunip=a.phaseuni;
  ConfigActive(a);
  switch(nx->type){
    case MotionMov::RectAce:  if(nx->velprev)for(int i=0;i<3;i++)a.vel[i]=velp[i]; break;
    case MotionMov::RotAce: case MotionMov::CirAce: if(nx->velprev)a.velang=velangp; break;
    case MotionMov::RectSinu: if(nx->phaseprev)for(int i=0;i<3;i++)a.phase[i]=php[i]; break;
    case MotionMov::RotSinu: case MotionMov::CirSinu: if(nx->phaseprev)a.phaseuni=phunip; break;
    default: break;
  }
}
} // namespace

void MotionSet::Reset(){
  lasttime=0;
  for(auto *list:{&objs,&groups})for(auto &o:*list){
    o.active.clear(); o.moving=false;
    for(auto &e:o.events)e.done=false;
    for(auto &m:o.movs)for(int i=0;i<3;i++){ m.axisp1[i]=m.iaxisp1[i]; m.axisp2[i]=m.iaxisp2[i]; m.ref[i]=m.iref[i]; }     // JMotionAxis::ResetTime
  }
}

bool MotionSet::Process(double timestep,double dt,std::vector<dsph_motiondata> &out){
  if(lasttime>timestep)throw std::runtime_error("The previous timestep+dt is higher than the requested timestep. It is invalid for simple mode.");
  out.assign(objs.size(),dsph_motiondata());
  for(size_t io=0;io<objs.size();io++){ dsph_motiondata &d=out[io]; memset(&d,0,sizeof(d)); d.idbegin=objs[io].idbegin; d.count=objs[io].count; d.type=DSPH_MOTION_NONE; }
  bool anymove=false;
  if(roots.empty())for(size_t io=0;io<objs.size();io++)roots.push_back({0,int(io)});   // a set filled in memory (mkcase): every object is a first-level one
  for(auto &r:roots)ProcessNode(r.first,r.second,timestep,dt,out,anymove);
  lasttime=timestep+dt;
  return anymove;
}

// JMotionObj::ProcesTime (S/JMotionObj.cpp:368-575) of one object, then of its children
void MotionSet::ProcessNode(int kind,int idx,double timestep,double dt,std::vector<dsph_motiondata> &out,bool &anymove){
  MotionObj &o=(kind==0? objs[idx]: groups[idx]);
  const double tstepfin=timestep+dt;
  const MotionObj *parent=(o.pkind<0? nullptr: (o.pkind==0? &objs[o.pidx]: &groups[o.pidx]));
  {
    // the motion of the parent in this step carries the axes (and the travelling points of circular movements) of this object along
    if(parent && parent->moving){
      const ModPos pm=ModPos::Of(*parent);
      for(auto &m:o.movs){ pm.PointMove(m.axisp1); pm.PointMove(m.axisp2); if(m.type>=MotionMov::Cir && m.type<=MotionMov::CirSinu)pm.PointMove(m.ref); }
    }
    // events that start before the end of this step
    for(auto &e:o.events)if(!e.done && e.start<tstepfin){
      e.done=true;
      MotionActive a; a.mov=FindMov(o,e.movid);
      if(!a.mov)throw std::runtime_error("motion: <begin> refers to a movement that does not exist.");
      a.start=e.start; a.eventfinish=e.finish; a.flash=(a.mov->time<0);
      a.finish=(a.flash? a.start: a.start+a.mov->time);
      if(a.eventfinish>=0 && a.eventfinish<a.finish)a.finish=a.eventfinish;
      ConfigActive(a);
      o.active.push_back(a);
    }
    ModPos mp; bool modif=false;
    for(size_t ca=0;ca<o.active.size();ca++){
      MotionActive &am=o.active[ca];
      if(am.del){ o.active.erase(o.active.begin()+ca); ca--; continue; }
      bool rep; double dt2=dt, timestep2=timestep;
      do{
        rep=false;
        const MotionMov &mv=*am.mov;
        double dtmov=(tstepfin>am.finish? am.finish-timestep2: dt2);
        const double dtover=dt2-dtmov;
        if(timestep2<am.start)dtmov-=(am.start-timestep2);
        if(dtmov>0 || am.flash){
          const double t=(am.flash? -mv.time: dtmov);
          switch(mv.type){
            case MotionMov::Rect:{ const double m3[3]={mv.vel[0]*t,mv.vel[1]*t,mv.vel[2]*t}; mp.Move(m3); modif=true; }break;
            case MotionMov::RectAce:{
              const double at[3]={mv.ace[0]*t,mv.ace[1]*t,mv.ace[2]*t};
              const double m3[3]={am.vel[0]*t+0.5f*at[0]*t,am.vel[1]*t+0.5f*at[1]*t,am.vel[2]*t+0.5f*at[2]*t};
              mp.Move(m3); for(int i=0;i<3;i++)am.vel[i]+=at[i]; modif=true;
            }break;
            case MotionMov::Rot:{ mp.Rotate(mv.velang*t,mv.axisp1,mv.axisp2); modif=true; }break;
            case MotionMov::RotAce:{ const double at=mv.aceang*t; mp.Rotate(am.velang*t+0.5f*at*t,mv.axisp1,mv.axisp2); am.velang+=at; modif=true; }break;
            case MotionMov::RectSinu:{
              double p1[3]={0,0,0},p2[3]={0,0,0};
              for(int i=0;i<3;i++)if(mv.ampl[i]){ p1[i]=mv.ampl[i]*std::sin(am.phase[i]); am.phase[i]+=double(mv.freq[i]*(kPi*2.0)*t); p2[i]=mv.ampl[i]*std::sin(am.phase[i]); }
              const double m3[3]={p2[0]-p1[0],p2[1]-p1[1],p2[2]-p1[2]};
              mp.Move(m3); modif=true;
            }break;
            case MotionMov::RotSinu:{
              double ph=am.phaseuni; double ang=mv.ampl[0]*std::sin(ph);
              ph+=double(mv.freq[0]*(kPi+kPi)*t);
              ang=mv.ampl[0]*std::sin(ph)-ang;
              mp.Rotate(ang,mv.axisp1,mv.axisp2); am.phaseuni=ph; modif=true;
            }break;
            // circular: the reference point travels round the axis and the object is translated with it (no rotation of the object)
            case MotionMov::Cir: case MotionMov::CirAce: case MotionMov::CirSinu:{
              double ang;
              if(mv.type==MotionMov::Cir)ang=mv.velang*t;
              else if(mv.type==MotionMov::CirAce){ const double at=mv.aceang*t; ang=am.velang*t+0.5f*at*t; am.velang+=at; }
              else{ double ph=am.phaseuni; const double a0=mv.ampl[0]*std::sin(ph); ph+=double(mv.freq[0]*(kPi+kPi)*t); ang=mv.ampl[0]*std::sin(ph)-a0; am.phaseuni=ph; }
              const Mat4 m=Mat4::Rot(ang,mv.axisp1,mv.axisp2);
              double *p1=const_cast<MotionMov&>(mv).ref; double p2[3];
              for(int i=0;i<3;i++)p2[i]=m(i,0)*p1[0]+m(i,1)*p1[1]+m(i,2)*p1[2]+m(i,3);
              const double m3[3]={p2[0]-p1[0],p2[1]-p1[1],p2[2]-p1[2]};
              mp.Move(m3); for(int i=0;i<3;i++)p1[i]=p2[i]; modif=true;
            }break;
            // file-driven: the time of the data is the time since the start of the movement measured from the START of the step
            // (S/JMotionObj.cpp:524-545 uses timestep, not the split time), plus the first time of the file, plus the part of dt that is used
            case MotionMov::RectFile:{
              double tq=timestep-am.start; if(tq<0)tq=0;
              tq+=mv.ftimes[0]; tq+=t;
              double np3[3]; DfQuery(mv,am,tq,3,np3);
              if(mv.fieldx<0)np3[0]=0; if(mv.fieldy<0)np3[1]=0; if(mv.fieldz<0)np3[2]=0;
              const double m3[3]={np3[0]-am.dflastpos[0],np3[1]-am.dflastpos[1],np3[2]-am.dflastpos[2]};
              mp.Move(m3); for(int i=0;i<3;i++)am.dflastpos[i]=np3[i]; modif=true;
            }break;
            case MotionMov::RotFile:{
              double tq=timestep-am.start; if(tq<0)tq=0;
              tq+=mv.ftimes[0]; tq+=t;
              double na; DfQuery(mv,am,tq,1,&na);
              mp.Rotate(na-am.dflastang,mv.axisp1,mv.axisp2); am.dflastang=na; modif=true;
            }break;
            default: break;
          }
        }
        if((dtover>0 || dtmov==0) && FindMov(o,mv.nextid)){
          NextMov(o,am);
          dt2=dtover; timestep2=am.start;
          if(timestep2<=am.finish || am.flash)rep=true;
        }
      }while(rep);
      if(tstepfin>am.finish)am.del=true;
    }
    if(parent && parent->moving){                                    // S/JMotionObj.cpp:562-565
      const ModPos pm=ModPos::Of(*parent);
      if(modif)mp.MoveMix(pm); else mp=pm;
      modif=true;
    }
    o.moving=modif; mp.Store(o);
    if(modif && kind==0){
      anymove=true;
      dsph_motiondata &d=out[idx];
      if(mp.simple){
        d.type=DSPH_MOTION_LINEAR;
        for(int i=0;i<3;i++){ d.linmov[i]=mp.s[i]; d.linvel[i]=mp.s[i]/dt; }
        if(simulate2d)d.linmov[1]=d.linvel[1]=0;
      }
      else{ d.type=DSPH_MOTION_MATRIX; memcpy(d.matmov,mp.m.a,sizeof(d.matmov)); }
    }
  }
  for(auto &c:o.children)ProcessNode(c.first,c.second,timestep,dt,out,anymove);
}

// ------------------------------------------------------------------------------------------------
static void Read3(const XmlNode *e,const char *name,double v[3]){
  const XmlNode *q=e->Child(name);
  if(!q)throw std::runtime_error(std::string("xml: <")+e->name+"><"+name+"> is missing");
  v[0]=q->Dbl("x"); v[1]=q->Dbl("y"); v[2]=q->Dbl("z");
}
static double Read1(const XmlNode *e,const char *name,const char *att){
  const XmlNode *q=e->Child(name);
  if(!q)throw std::runtime_error(std::string("xml: <")+e->name+"><"+name+"> is missing");
  return q->Dbl(att);
}

void MotionSet::LoadXml(const XmlNode *motion,const std::string &dirdata){
  // objs[] already holds one entry per moving block (ref, mkbound, idbegin, count); this adds the object tree, movements and events
  for(auto &o:objs){ o.movs.clear(); o.events.clear(); o.active.clear(); o.children.clear(); o.pkind=o.pidx=-1; }
  groups.clear(); roots.clear();
  if(motion)for(auto &ep:motion->children){
    const XmlNode *ob=ep.get();
    if(ob->name.empty() || ob->name[0]=='_')continue;
    if(ob->name!="objreal" && ob->name!="obj")throw std::runtime_error("motion: <"+ob->name+"> is not expected at the first level of <motion>.");
    LoadObj(ob,-1,-1,dirdata);
  }
  for(auto &o:objs)if(o.pkind<0){ bool isroot=false; for(auto &r:roots)if(r.first==0 && &objs[r.second]==&o)isroot=true; if(!isroot)roots.push_back({0,int(&o-&objs[0])}); }   // moving blocks without motion
  // JMotion processes its events from the latest to the earliest start (S/JMotion.cpp:446-460): order of activation = order of start
  for(auto *list:{&objs,&groups})for(auto &o:*list)for(size_t i=1;i<o.events.size();i++)for(size_t j=i;j>0 && o.events[j].start<o.events[j-1].start;j--)std::swap(o.events[j],o.events[j-1]);
  Reset();
}

// one <obj> / <objreal> element: its movements, its events and, recursively, the objects inside it (JMotion::ReadXml S/JMotion.cpp:557-700)
void MotionSet::LoadObj(const XmlNode *ob,int pkind,int pidx,const std::string &dirdata){
  int kind,idx;
  if(ob->name=="objreal"){
    const int ref=int(ob->Int("ref"));
    idx=-1; for(size_t q=0;q<objs.size();q++)if(objs[q].ref==ref)idx=int(q);
    if(idx<0)throw std::runtime_error("The number of mobile objects do not match the predefined motions in XML file.");
    kind=0;
  }
  else{ groups.push_back(MotionObj()); kind=1; idx=int(groups.size())-1; }
  { MotionObj &o=(kind==0? objs[idx]: groups[idx]); o.pkind=pkind; o.pidx=pidx; }
  if(pkind<0)roots.push_back({kind,idx});
  else (pkind==0? objs[pidx]: groups[pidx]).children.push_back({kind,idx});
  {
    for(auto &mp:ob->children){
      const XmlNode *e=mp.get();
      const std::string &n=e->name;
      if(n.empty() || n[0]=='_')continue;
      if(n=="obj" || n=="objreal"){ LoadObj(e,kind,idx,dirdata); continue; }
      MotionObj *o=(kind==0? &objs[idx]: &groups[idx]);          // (groups may have grown: fetched again for every element)
      if(n=="begin"){ MotionEvent ev; ev.movid=int(e->Int("mov")); ev.start=double(float(e->Dbl("start"))); ev.finish=double(float(e->Dbl("finish",true,-1)));   /* JXml::GetAttributeFloat: single precision */ o->events.push_back(ev); continue; }
      MotionMov m; m.id=int(e->Int("id")); m.time=double(float(e->Dbl("duration"))); m.nextid=int(e->Int("next",true,0));   // duration: read in single precision by the reference (S/JMotion.cpp:573)
      const std::string units=e->Str("anglesunits",true,"degrees");
      if(units!="degrees" && units!="radians")throw std::runtime_error("xml: anglesunits must be degrees or radians.");
      const bool deg=(units=="degrees");
      if(n=="wait")m.type=MotionMov::Wait;
      else if(n=="mvrect"){ m.type=MotionMov::Rect; Read3(e,"vel",m.vel); }
      else if(n=="mvrectace"){ m.type=MotionMov::RectAce; Read3(e,"ace",m.ace); m.velprev=e->Child("velini")==nullptr; if(!m.velprev)Read3(e,"velini",m.vel); }
      else if(n=="mvrot"){ m.type=MotionMov::Rot; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2); m.velang=Read1(e,"vel","ang"); if(!deg)m.velang*=kToDeg; }
      else if(n=="mvrotace"){
        m.type=MotionMov::RotAce; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2); m.aceang=Read1(e,"ace","ang");
        m.velprev=e->Child("velini")==nullptr; if(!m.velprev)m.velang=Read1(e,"velini","ang");
        if(!deg){ m.aceang*=kToDeg; m.velang*=kToDeg; }
      }
      else if(n=="mvrectsinu"){
        m.type=MotionMov::RectSinu; Read3(e,"freq",m.freq); Read3(e,"ampl",m.ampl);
        m.phaseprev=e->Child("phase")==nullptr; if(!m.phaseprev){ Read3(e,"phase",m.phase); if(deg)for(int i=0;i<3;i++)m.phase[i]*=kToRad; }
      }
      else if(n=="mvrotsinu"){
        m.type=MotionMov::RotSinu; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2); m.freq[0]=Read1(e,"freq","v"); m.ampl[0]=Read1(e,"ampl","v");
        m.phaseprev=e->Child("phase")==nullptr; if(!m.phaseprev)m.phase[0]=Read1(e,"phase","v");
        if(!deg)m.ampl[0]*=kToDeg;
        if(deg)m.phase[0]*=kToRad;
      }
      else if(n=="mvcir"){ m.type=MotionMov::Cir; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2); Read3(e,"ref",m.ref); m.velang=Read1(e,"vel","ang"); if(!deg)m.velang*=kToDeg; }
      else if(n=="mvcirace"){
        m.type=MotionMov::CirAce; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2); Read3(e,"ref",m.ref); m.aceang=Read1(e,"ace","ang");
        m.velprev=e->Child("velini")==nullptr; if(!m.velprev)m.velang=Read1(e,"velini","ang");
        if(!deg){ m.aceang*=kToDeg; m.velang*=kToDeg; }
      }
      else if(n=="mvcirsinu"){
        m.type=MotionMov::CirSinu; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2); Read3(e,"ref",m.ref); m.freq[0]=Read1(e,"freq","v"); m.ampl[0]=Read1(e,"ampl","v");
        m.phaseprev=e->Child("phase")==nullptr; if(!m.phaseprev)m.phase[0]=Read1(e,"phase","v");
        if(!deg)m.ampl[0]*=kToDeg;
        if(deg)m.phase[0]*=kToRad;
      }
      else if(n=="mvrectfile" || n=="mvfile" || n=="mvpredef" || n=="mvrotfile"){
        const XmlNode *ef=e->Child("file");
        if(!ef)throw std::runtime_error("xml: <"+n+"><file> is missing");
        std::string file=ef->Str("name");
        if(file.find('/')==std::string::npos && file.find('\\')==std::string::npos)file=dirdata+file;       // only a name: next to the case (S/JMotionMov.cpp:290)
        const std::vector<std::vector<double>> rows=ReadDataFile(file);
        if(rows.size()<2)throw std::runtime_error("motion: the data file "+file+" cannot have less than two rows.");
        if(n=="mvrotfile"){
          m.type=MotionMov::RotFile; Read3(e,"axisp1",m.axisp1); Read3(e,"axisp2",m.axisp2);
          for(auto &r:rows){ if(r.size()<2)throw std::runtime_error("motion: a row of "+file+" has less than two values."); m.ftimes.push_back(r[0]); m.fvals.push_back(deg? r[1]: r[1]*kToDeg); }
        }
        else{
          m.type=MotionMov::RectFile;
          const int fields=int(ef->Int("fields")), ft=int(ef->Int("fieldtime"));
          m.fieldx=int(ef->Int("fieldx",true,-1)); m.fieldy=int(ef->Int("fieldy",true,-1)); m.fieldz=int(ef->Int("fieldz",true,-1));
          for(auto &r:rows){
            if(int(r.size())<fields || ft<0 || ft>=fields)throw std::runtime_error("motion: a row of "+file+" has less values than fields.");
            m.ftimes.push_back(r[ft]);
            m.fvals.push_back(m.fieldx>=0 && m.fieldx<fields? r[m.fieldx]: 0.0); m.fvals.push_back(m.fieldy>=0 && m.fieldy<fields? r[m.fieldy]: 0.0); m.fvals.push_back(m.fieldz>=0 && m.fieldz<fields? r[m.fieldz]: 0.0);
          }
        }
      }
      else throw std::runtime_error("motion: <"+n+"> is not supported by the stand-alone driver (wait, mvrect, mvrectace, mvrot, mvrotace, mvcir, mvcirace, mvrectsinu, mvrotsinu, mvcirsinu, mvrectfile, mvrotfile are).");
      for(int i=0;i<3;i++){ m.iaxisp1[i]=m.axisp1[i]; m.iaxisp2[i]=m.axisp2[i]; m.iref[i]=m.ref[i]; }
      o->movs.push_back(m);
    }
  }
}

void MotionSet::SaveXml(FILE *f,const char *ind)const{
  auto D=[](double v){ char t[64]; snprintf(t,sizeof(t),"%.17g",v); return std::string(t); };
  fprintf(f,"%s<motion>\n",ind);
  for(auto &o:objs){
    fprintf(f,"%s  <objreal ref=\"%d\">\n",ind,o.ref);
    for(auto &e:o.events){
      if(e.finish>=0)fprintf(f,"%s    <begin mov=\"%d\" start=\"%s\" finish=\"%s\" />\n",ind,e.movid,D(e.start).c_str(),D(e.finish).c_str());
      else fprintf(f,"%s    <begin mov=\"%d\" start=\"%s\" />\n",ind,e.movid,D(e.start).c_str());
    }
    for(auto &m:o.movs){
      static const char *tn[]={"wait","mvrect","mvrectace","mvrot","mvrotace","mvrectsinu","mvrotsinu"};
      if(m.type>MotionMov::RotSinu)throw std::runtime_error("motion: circular and file-driven movements are read, not written, by this library.");
      std::string head=std::string(ind)+"    <"+tn[m.type]+" id=\""+std::to_string(m.id)+"\" duration=\""+D(m.time)+"\"";
      if(m.nextid)head+=" next=\""+std::to_string(m.nextid)+"\"";
      if(m.type>=MotionMov::Rot)head+=" anglesunits=\"radians\"";
      auto P3=[&](const char *n,const double *v){ fprintf(f,"%s      <%s x=\"%s\" y=\"%s\" z=\"%s\" />\n",ind,n,D(v[0]).c_str(),D(v[1]).c_str(),D(v[2]).c_str()); };
      if(m.type==MotionMov::Wait){ fprintf(f,"%s />\n",head.c_str()); continue; }
      fprintf(f,"%s>\n",head.c_str());
      switch(m.type){
        case MotionMov::Rect: P3("vel",m.vel); break;
        case MotionMov::RectAce: P3("ace",m.ace); if(!m.velprev)P3("velini",m.vel); break;
        case MotionMov::Rot: P3("axisp1",m.axisp1); P3("axisp2",m.axisp2); fprintf(f,"%s      <vel ang=\"%s\" />\n",ind,D(m.velang*kToRad).c_str()); break;
        case MotionMov::RotAce: P3("axisp1",m.axisp1); P3("axisp2",m.axisp2); fprintf(f,"%s      <ace ang=\"%s\" />\n",ind,D(m.aceang*kToRad).c_str());
          if(!m.velprev)fprintf(f,"%s      <velini ang=\"%s\" />\n",ind,D(m.velang*kToRad).c_str()); break;
        case MotionMov::RectSinu: P3("freq",m.freq); P3("ampl",m.ampl); if(!m.phaseprev)P3("phase",m.phase); break;
        case MotionMov::RotSinu: P3("axisp1",m.axisp1); P3("axisp2",m.axisp2); fprintf(f,"%s      <freq v=\"%s\" />\n%s      <ampl v=\"%s\" />\n",ind,D(m.freq[0]).c_str(),ind,D(m.ampl[0]*kToRad).c_str());
          if(!m.phaseprev)fprintf(f,"%s      <phase v=\"%s\" />\n",ind,D(m.phase[0]).c_str()); break;
        default: break;
      }
      fprintf(f,"%s    </%s>\n",ind,tn[m.type]);
    }
    fprintf(f,"%s  </objreal>\n",ind);
  }
  fprintf(f,"%s</motion>\n",ind);
}

} // namespace dsph
