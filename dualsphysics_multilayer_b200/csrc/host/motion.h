// motion.h -- predefined motion laws of the moving boundary objects of a case: what the reference's JDsMotion / JMotion evaluate per
// step in "simple" mode (S/JDsMotion.cpp:121-140, S/JMotion.cpp:347-366, S/JMotionObj.cpp:363-575).  Host code of the stand-alone
// driver; a host application that keeps JDsMotion passes its own StMotionData to dsph_motion_run instead.
// Supported: <objreal ref> with <begin mov start [finish]> events and chains (next) of wait, mvrect, mvrectace, mvrot, mvrotace, mvcir,
// mvcirace, mvrectsinu, mvrotsinu, mvcirsinu, mvrectfile (= mvfile, mvpredef) and mvrotfile, in nested <obj> / <objreal> objects.
#pragma once
#include "../../../include/dsph_b200.h"
#include <cstdio>
#include <string>
#include <vector>
#include <utility>

namespace dsph {

struct XmlNode;

struct MotionMov {
  enum Type { Wait=0, Rect, RectAce, Rot, RotAce, RectSinu, RotSinu, Cir, CirAce, CirSinu, RectFile, RotFile };
  Type type=Wait; int id=0, nextid=0; double time=0;
  double vel[3]={0,0,0}, ace[3]={0,0,0}; bool velprev=false;          // rectilinear (velini in vel)
  double axisp1[3]={0,0,0}, axisp2[3]={0,0,0}; double velang=0, aceang=0;   // rotation, degrees (velini in velang)
  double freq[3]={0,0,0}, ampl[3]={0,0,0}, phase[3]={0,0,0}; bool phaseprev=false;   // sinusoidal: phase in radians; rotsinu / cirsinu use [0], ampl in degrees
  double ref[3]={0,0,0};                                                             // circular: the point that travels round the axis (current position)
  double iaxisp1[3]={0,0,0}, iaxisp2[3]={0,0,0}, iref[3]={0,0,0};                    // initial axis / point: the motion of a parent object carries them along
  std::vector<double> ftimes, fvals;                                                 // file-driven: times; positions xyz (rect) or angles in degrees (rot)
  int fieldx=-1, fieldy=-1, fieldz=-1;
};
struct MotionEvent { int movid; double start, finish; bool done=false; };
struct MotionActive {           // JMotionMovActive
  const MotionMov *mov=nullptr; double start=0, finish=0, eventfinish=-1; bool flash=false, del=false;
  double vel[3]={0,0,0}, velang=0, phase[3]={0,0,0}, phaseuni=0;
  unsigned dfindex=0; double dflastpos[3]={0,0,0}, dflastang=0;   // file-driven: interval of the last query, last position / angle
};
struct MotionObj {
  int ref=-1; unsigned mkbound=0, idbegin=0, count=0;           // ref >= 0: a moving block of <particles>; -1: an <obj> that only groups others
  std::vector<MotionMov> movs; std::vector<MotionEvent> events; std::vector<MotionActive> active;
  // object tree (<obj> / <objreal> may contain others; a child moves with its parent): kind 0 = objs[], 1 = groups[]
  int pkind=-1, pidx=-1; std::vector<std::pair<int,int>> children;
  // result of the current step (JMotionObj::ModPos, Moving)
  bool moving=false, mpsimple=true; double mps[3]={0,0,0}, mpm[16]={1,0,0,0, 0,1,0,0, 0,0,1,0, 0,0,0,1};
};

class MotionSet {
 public:
  std::vector<MotionObj> objs;      // the moving blocks, in the order of <particles>
  std::vector<MotionObj> groups;    // <obj> elements (no particles of their own)
  std::vector<std::pair<int,int>> roots;   // first-level objects in the order of the XML
  double lasttime=0;
  bool simulate2d=false;
  // <case><execution><special><motion> + the moving blocks of <particles>; throws on unsupported content
  void LoadXml(const XmlNode *motion,const std::string &dirdata="");
  void SaveXml(FILE *f,const char *indent)const;
  // = JDsMotion::ProcesTime(MOMT_Simple,timestep,dt) followed by GetMotionData of every object; returns true when any object moves
  bool Process(double timestep,double dt,std::vector<dsph_motiondata> &out);
  void Reset();
 private:
  void ProcessNode(int kind,int idx,double timestep,double dt,std::vector<dsph_motiondata> &out,bool &anymove);
  void LoadObj(const XmlNode *ob,int pkind,int pidx,const std::string &dirdata);
};

} // namespace dsph
