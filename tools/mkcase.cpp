// mkcase -- writes <dir>/<name>.xml + <dir>/<name>.bi4 for the synthetic cases of SURVEY.md 8(d).
// Stand-in for GenCase (a missing blob in the reference, R/.MISSING_LARGE_BLOBS); the files load both
// in the reference solver (oracle/_ref) and in dsph_b200.
//   mkcase <dambreak|periodicbox|nn2layer> <np_target> <dir> <name> [key=value ...]
// keys: tstep=1|2 tkernel=1|2 tvisco=1|2 visco=v ddt=0..3 ddtvalue=v shifting=0..3 shiftcoef=v shifttfs=v
//       timemax=t timeout=t cellmode=1|2 seed=n boundary=1|2 (2 = mDBC: writes <name>_Normals.nbi4) mdbccorrector=0|1
#include "../dualsphysics_multilayer_b200/csrc/host/casefile.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <sys/stat.h>
using namespace dsph;
int main(int argc,char **argv){
  if(argc<5){ fprintf(stderr,"usage: mkcase <dambreak|periodicbox|nn2layer> <np_target> <dir> <name> [key=value ...]\n"); return 2; }
  try{
    unsigned seed=12345;
    for(int i=5;i<argc;i++)if(!strncmp(argv[i],"seed=",5))seed=unsigned(atoi(argv[i]+5));
    CaseData cd; MakeCase(cd,argv[1],strtoull(argv[2],nullptr,10),seed);
    dsph_config &c=cd.cfg;
    for(int i=5;i<argc;i++){
      std::string kv=argv[i]; size_t e=kv.find('='); if(e==std::string::npos)throw std::runtime_error("bad option "+kv);
      const std::string k=kv.substr(0,e); const double v=atof(kv.c_str()+e+1);
      if(k=="tstep")c.tstep=int(v); else if(k=="tkernel")c.tkernel=int(v); else if(k=="tvisco")c.tvisco=int(v);
      else if(k=="visco")c.visco=float(v); else if(k=="ddt")c.tdensity=int(v); else if(k=="ddtvalue")c.ddtvalue=float(v);
      else if(k=="shifting")c.shiftmode=int(v); else if(k=="shiftcoef")c.shiftcoef=float(v); else if(k=="shifttfs")c.shifttfs=float(v);
      else if(k=="timemax")c.timemax=v; else if(k=="timeout")c.timeout=v; else if(k=="cellmode")c.cellmode=int(v);
      else if(k=="verletsteps")c.verletsteps=int(v); else if(k=="velgrad")c.nn_velgrad=int(v);
      else if(k=="boundary"){ c.tboundary=(c.tboundary&~0xff)|int(v); } else if(k=="mdbccorrector"){ if(int(v))c.tboundary|=DSPH_BOUNDARY_MDBCCORRECTOR; }
      else if(k=="seed"){} else throw std::runtime_error("unknown option "+k);
    }
    if((c.tboundary&0xff)==DSPH_BOUNDARY_MDBC)MakeTankNormals(cd);      // <name>_Normals.nbi4 for the reference and for dsph_b200
    mkdir(argv[3],0755);
    SaveCase(cd,argv[3],argv[4]);
    printf("%s: np=%u nb=%u nf=%u dp=%.9g h=%.9g\n",argv[4],cd.np(),c.casenfixed,c.casenfluid,c.dp,double(c.kernelh));
  }
  catch(const std::exception &e){ fprintf(stderr,"mkcase: %s\n",e.what()); return 1; }
  return 0;
}
