// dsph_b200 -- command line driver (a subset of the reference's DualSPHysics5.2 options, R/doc/help/DualSPHysics5.2_Help.out):
//   dsph_b200 <casedir/casename> [dirout] [-gpu[:id]] [-mgpu:N] [-nsteps:N] [-sv:none] [-tmax:T] [-tout:T] [-forcevariant:N] [-partbegin:N dir]
//             [-rebalance:K] [-piecemax:N]
//   -mgpu:N  x-slab run on the GPUs id..id+N-1 (one host thread, one NCCL rank and one PART piece per GPU)
#include "../../include/dsph_b200.h"
#include "host/driver.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
int main(int argc,char **argv){
  if(argc<2){ printf("usage: dsph_b200 <casedir/casename> [dirout] [-gpu[:id]] [-mgpu:N] [-nsteps:N] [-sv:none] [-tmax:T] [-tout:T] [-partbegin:N dir]\n"); return 1; }
  std::string casepath=argv[1], dirout="out", partbegindir; int device=0,nsteps=0,savepart=1,partbegin=-1,forcevariant=0,mgpu=0,rebalance=0; unsigned piecemax=0; double tmax=-1,tout=-1;
  int narg=0;
  for(int i=2;i<argc;i++){
    std::string a=argv[i];
    if(a[0]!='-'){ if(narg++==0)dirout=a; continue; }
    if(a.compare(0,6,"-tmax:")==0){ tmax=atof(a.c_str()+6); continue; }
    if(a.compare(0,6,"-tout:")==0){ tout=atof(a.c_str()+6); continue; }
    if(a.compare(0,14,"-forcevariant:")==0){ forcevariant=atoi(a.c_str()+14); continue; }
    if(a.compare(0,11,"-partbegin:")==0){ partbegin=atoi(a.c_str()+11); if(i+1<argc && argv[i+1][0]!='-')partbegindir=argv[++i]; continue; }
    if(a.compare(0,6,"-mgpu:")==0){ mgpu=atoi(a.c_str()+6); continue; }
    if(a.compare(0,11,"-rebalance:")==0){ rebalance=atoi(a.c_str()+11); continue; }
    if(a.compare(0,10,"-piecemax:")==0){ piecemax=unsigned(atoll(a.c_str()+10)); continue; }
    if(a.compare(0,5,"-gpu:")==0)device=atoi(a.c_str()+5);
    else if(a=="-gpu"){}
    else if(a=="-cpu"){ fprintf(stderr,"dsph_b200: there is no CPU path (use the reference CPU solver).\n"); return 1; }
    else if(a.compare(0,8,"-nsteps:")==0)nsteps=atoi(a.c_str()+8);
    else if(a=="-sv:none")savepart=0;
    else if(a.compare(0,4,"-sv:")==0){}
    else { fprintf(stderr,"dsph_b200: option ignored: %s\n",a.c_str()); }
  }
  std::string dir=".",name=casepath; size_t p=casepath.find_last_of('/');
  if(p!=std::string::npos){ dir=casepath.substr(0,p); name=casepath.substr(p+1); }
  dsph_runopts ro; memset(&ro,0,sizeof(ro));
  ro.dir=dir.c_str(); ro.casename=name.c_str(); ro.dirout=dirout.c_str(); ro.device=device; ro.nsteps=nsteps; ro.savepart=savepart;
  ro.timemax=tmax; ro.timeout=tout; ro.forcevariant=forcevariant; ro.restart=(partbegin>=0? 1: 0); ro.partbegin=(partbegin>=0? partbegin: 0); ro.partbegindir=(partbegindir.empty()? nullptr: partbegindir.c_str());
  ro.mgpu=mgpu; ro.rebalancesteps=rebalance; ro.piecemax=piecemax;
  const int rc=dsph_run_case_ex(&ro);
  if(rc!=DSPH_OK){ fprintf(stderr,"\n*** Exception: %s\nFinished execution (code=1).\n",dsph_last_error()); return 1; }
  printf("\nFinished execution (code=0).\n");
  return 0;
}
