// casefile.h -- Case.xml + Case.bi4 input (GenCase output format) and the synthetic case generator.
// XML keys follow S/JSph.cpp:551-757 (LoadKernelSelection/LoadConfigCtes/LoadConfigParameters),
// S/JCaseCtes.cpp:201-215, S/JCaseParts.cpp:484-508,60-64, S/JCaseEParms.cpp:286-375 and, for the
// multiphase fork, M/JSph.cpp:608-628,3137-3214.
#pragma once
#include "../../../include/dsph_b200.h"
#include "motion.h"
#include <string>
#include <vector>

namespace dsph {

struct MkBlock { int type; /*0 fixed,1 moving,2 floating,3 fluid*/ unsigned mktype, mk, begin, count; unsigned refmotion=0; };

struct DomainLimit { int mode=0; /*0 default,1 default+-value,2 default+-prc,3 fixed*/ double value=0; };

struct CaseData {
  dsph_config cfg;
  std::string name;
  std::vector<MkBlock> blocks;
  unsigned mkboundfirst=11, mkfluidfirst=1;
  DomainLimit posmin[3], posmax[3];        // <simulationdomain>
  double caseposmin[3]={0,0,0}, caseposmax[3]={0,0,0};
  std::vector<unsigned> idp;
  std::vector<double> pos;                 // xyz packed
  std::vector<float> velrhop;              // xyzw packed
  std::vector<dsph_gauge> gauges;          // <special><gauges> (S/JDsGaugeSystem.cpp:186-275)
  MotionSet motion;                        // <execution><motion> of the moving blocks (S/JDsMotion.cpp:94-106)
  std::vector<double> normals;             // mDBC: vector particle -> boundary limit per boundary particle, id order, xyz packed (<name>_Normals.nbi4)
  unsigned np()const{ return (unsigned)idp.size(); }
  unsigned npb()const{ return cfg.casenfixed+cfg.casenmoving; }
};

void ConfigDefaults(dsph_config &c);
// MapRealPosMin/Max from CasePosMin/Max (JPartsLoad4::CalculeLimits + JSph::ResizeMapLimits)
void ComputeMapRealPos(CaseData &cd);
// particle code from id (JSph::LoadCodeParticles -> JSphMk::GetCodeById)
void ComputeCodes(const CaseData &cd,unsigned np,const unsigned *idp,dsph_typecode *code);

// mDBC normal vectors of the synthetic tank cases: every wall particle points half a particle spacing into the tank along each wall it lies on
void MakeTankNormals(CaseData &cd);
// <dir>/<casename>_Normals.nbi4 (JPartNormalData::LoadFile/SaveFile, S/JPartNormalData.cpp:179-262); empty when the file does not exist
void LoadCaseNormals(std::vector<double> &normals,const std::string &dir,const std::string &casename);
void SaveCase(const CaseData &cd,const std::string &dir,const std::string &casename);
void LoadCase(CaseData &cd,const std::string &dir,const std::string &casename);
void LoadCaseMotion(MotionSet &ms,const std::string &dir,const std::string &casename);

// kind: "dambreak", "periodicbox", "nn2layer" (DB(N), PB(N), NN(N))
void MakeCase(CaseData &cd,const std::string &kind,unsigned long long nptarget,unsigned seed);
// same case, but only the particles of the cell columns [cxlo,cxhi) are generated (ids/counts of the full case)
void MakeCaseSlab(CaseData &cd,const std::string &kind,unsigned long long nptarget,unsigned seed,unsigned cxlo,unsigned cxhi);

} // namespace dsph
