// gauges.cu -- Engine side of the in-solver gauges (kernels: gauges.cuh).  All the gauges due at a step are evaluated together:
// definitions go up in one small H2D copy, results come back in one D2H copy, one stream synchronisation per evaluation.
#include "engine.h"
#include "gauges.cuh"
#include <cstring>
#include <cfloat>
#include <cmath>

namespace dsph {

#define CUCHECKG(x) do{ cudaError_t e_=(x); if(e_!=cudaSuccess)throw EngineError(DSPH_ERR_CUDA,std::string(#x)+": "+cudaGetErrorString(e_)); }while(0)

void Engine::GaugesFree(){
  if(gaugedev){ cudaFree(gaugedev); gaugedev=nullptr; gaugedevcap=0; }
  if(gaugehost){ cudaFreeHost(gaugehost); gaugehost=nullptr; gaugehostcap=0; }
  if(gaugepartace){ cudaFree(gaugepartace); gaugepartace=nullptr; gaugepartacecap=0; }
}

void Engine::GaugesConfig(unsigned n,const dsph_gauge *g){
  gauges.assign(g,g+n);
  gaugenext.assign(n,0.0);
  gaugeres.assign(n,dsph_gaugeres());
  for(auto &q:gauges){
    if(q.type<DSPH_GAUGE_VEL || q.type>DSPH_GAUGE_FORCE)throw EngineError(DSPH_ERR_ARG,"Gauge type is invalid.");
    if(q.type==DSPH_GAUGE_SWL && !(q.pointdp>0))throw EngineError(DSPH_ERR_ARG,"The pointdp of a swl gauge is invalid.");
    if(q.type==DSPH_GAUGE_SWL && !(q.masslimit>0))q.masslimit=cfg.massfluid*(cfg.simulate2d? 0.4f: 0.5f);
    if(q.type==DSPH_GAUGE_MAXZ && !(q.distlimit>0))throw EngineError(DSPH_ERR_ARG,"The distlimit of a maxz gauge is invalid.");
  }
  if(n && mgpu && k.periactive)throw EngineError(DSPH_ERR_STATE,"Gauges are not evaluated in x-slab runs of periodic cases (the sums of a rank leave out the periodic copies it holds).");
}

// JGaugeSwl::SetPoints (S/JDsGaugeItem.cpp:447-465)
static void SwlLine(const dsph_gauge &q,GaugeLine &ln,unsigned first){
  const double d[3]={q.point2[0]-q.point0[0],q.point2[1]-q.point0[1],q.point2[2]-q.point0[2]};
  const double dis=std::sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]);
  unsigned pointnp=0; double dir[3]={0,0,0};
  if(dis>0 && q.pointdp>0){
    pointnp=unsigned(dis/q.pointdp);
    if(dis-(q.pointdp*pointnp)>=q.pointdp*0.1)pointnp++;
    if(pointnp<1)pointnp++;
    const double dp=dis/pointnp;
    for(int i=0;i<3;i++)dir[i]=d[i]/dis*dp;
  }
  for(int i=0;i<3;i++){ ln.p0[i]=q.point0[i]; ln.dir[i]=dir[i]; }
  ln.first=first; ln.count=pointnp+1;
}

unsigned Engine::GaugesRun(double timestep,bool all,dsph_gaugeres *res){
  const unsigned ng=unsigned(gauges.size());
  for(auto &r:gaugeres)r.computed=0;
  if(!ng){ return 0; }
  if(!celltableok)throw EngineError(DSPH_ERR_STATE,"Gauges need the cell table of a divide (dsph_celldiv_run) for the current particle order.");
  CUCHECKG(cudaSetDevice(device));
  // which gauges are due: JGaugeItem::Update (S/JDsGaugeItem.h:164)
  std::vector<unsigned> vel,swl,mxz,frc;
  for(unsigned i=0;i<ng;i++){
    const dsph_gauge &q=gauges[i];
    if(all || (timestep>=gaugenext[i] && q.computestart<=timestep && timestep<=q.computeend)){
      (q.type==DSPH_GAUGE_VEL? vel: q.type==DSPH_GAUGE_SWL? swl: q.type==DSPH_GAUGE_MAXZ? mxz: frc).push_back(i);
    }
  }
  const unsigned ndue=unsigned(vel.size()+swl.size()+mxz.size()+frc.size());
  if(!ndue){ if(res)memcpy(res,gaugeres.data(),ng*sizeof(dsph_gaugeres)); return 0; }
  // ---- device input block: [vel points | swl lines | swl mass limits | maxz defs], output block: [vel 4 | mass | swl 3 | maxz 1 | force 3]
  std::vector<GaugeLine> lines(swl.size()); unsigned nlinepts=0;
  for(size_t q=0;q<swl.size();q++){ SwlLine(gauges[swl[q]],lines[q],nlinepts); nlinepts+=lines[q].count; }
  const size_t o_pts=0, o_lines=o_pts+vel.size()*24, o_lim=o_lines+lines.size()*sizeof(GaugeLine), o_mz=((o_lim+swl.size()*4+7)/8)*8, in_bytes=o_mz+mxz.size()*sizeof(GaugeMaxzDef);
  const size_t r_vel=0, r_mass=r_vel+vel.size()*4, r_swl=r_mass+nlinepts, r_mz=r_swl+swl.size()*3, r_frc=r_mz+mxz.size(), out_floats=r_frc+frc.size()*3;
  const size_t o_out=((in_bytes+15)/16)*16, total=o_out+out_floats*4;
  if(gaugedevcap<total){ if(gaugedev)cudaFree(gaugedev); CUCHECKG(cudaMalloc((void**)&gaugedev,total+4096)); gaugedevcap=total+4096; }
  if(gaugehostcap<total){ if(gaugehost)cudaFreeHost(gaugehost); CUCHECKG(cudaHostAlloc((void**)&gaugehost,total+4096,cudaHostAllocDefault)); gaugehostcap=total+4096; }
  char *hb=(char*)gaugehost;
  for(size_t q=0;q<vel.size();q++)memcpy(hb+o_pts+q*24,gauges[vel[q]].point0,24);
  if(!lines.empty())memcpy(hb+o_lines,lines.data(),lines.size()*sizeof(GaugeLine));
  for(size_t q=0;q<swl.size();q++)memcpy(hb+o_lim+q*4,&gauges[swl[q]].masslimit,4);
  for(size_t q=0;q<mxz.size();q++){ GaugeMaxzDef d; memcpy(d.p0,gauges[mxz[q]].point0,24); d.height=gauges[mxz[q]].height; d.distlimit=gauges[mxz[q]].distlimit; d.pad=0; memcpy(hb+o_mz+q*sizeof(GaugeMaxzDef),&d,sizeof(d)); }
  if(in_bytes)CUCHECKG(cudaMemcpyAsync(gaugedev,hb,in_bytes,cudaMemcpyHostToDevice,stm));
  float *dout=(float*)(gaugedev+o_out);
  GaugeArgs ga; ga.pos=pos[cur]; for(int i=0;i<3;i++)ga.domposmax[i]=k.domposmax[i];
  const ForceParams t=MakeForceParams();
  if(!vel.empty()){ const unsigned n=unsigned(vel.size()); ker_gauge_points<0><<<(n*32+127)/128,128,0,stm>>>(n,(const double*)(gaugedev+o_pts),nullptr,0,dout+r_vel,ga,t,dc); launches++; }
  if(mgpu && !vel.empty())MgAllReduceFloats(dout+r_vel,vel.size()*4,false);   // x-slab runs: partial sums over the owned particles of every rank
  if(!swl.empty()){
    // (slab runs: the masses at the line points are partial sums over owned particles: reduced below, before the scan)
    ker_gauge_points<1><<<(nlinepts*32+127)/128,128,0,stm>>>(nlinepts,nullptr,(const GaugeLine*)(gaugedev+o_lines),unsigned(lines.size()),dout+r_mass,ga,t,dc); launches++;
    if(mgpu)MgAllReduceFloats(dout+r_mass,nlinepts,false);
    ker_gauge_swl_scan<<<(unsigned(swl.size())+63)/64,64,0,stm>>>(unsigned(swl.size()),(const GaugeLine*)(gaugedev+o_lines),dout+r_mass,(const float*)(gaugedev+o_lim),dout+r_swl); launches++;
  }
  if(!mxz.empty()){
    const unsigned n=unsigned(mxz.size()); ker_gauge_maxz<<<(n*32+127)/128,128,0,stm>>>(n,(const GaugeMaxzDef*)(gaugedev+o_mz),dout+r_mz,ga,t,dc); launches++;
    if(mgpu)MgAllReduceFloats(dout+r_mz,mxz.size(),true);          // the highest particle any rank found
  }
  for(size_t kf=0;kf<frc.size();kf++){
    const dsph_gauge &q=gauges[frc[kf]];
    if(gaugepartacecap<q.count){ if(gaugepartace)cudaFree(gaugepartace); CUCHECKG(cudaMalloc((void**)&gaugepartace,size_t(q.count)*12+16)); gaugepartacecap=q.count; }
    CUCHECKG(cudaMemsetAsync(gaugepartace,0,size_t(q.count)*12,stm));
    if(npbok){ ker_gauge_force<<<(npbok*32u+127u)/128u,128,0,stm>>>(npbok,q.code,q.idbegin,q.count,idp[cur],gaugepartace,ga,t,dc); launches++; }
    ker_gauge_force_sum<<<1,256,0,stm>>>(q.count,gaugepartace,cfg.massbound,dout+r_frc+3*kf); launches++;
  }
  if(mgpu && !frc.empty())MgAllReduceFloats(dout+r_frc,frc.size()*3,false);   // every rank summed over the boundary particles it owns
  CUCHECKG(cudaMemcpyAsync(hb+o_out,dout,out_floats*4,cudaMemcpyDeviceToHost,stm));
  CUCHECKG(cudaStreamSynchronize(stm));
  CUCHECKG(cudaGetLastError());
  const float *ho=(const float*)(hb+o_out);
  auto done=[&](unsigned i,float x,float y,float z){
    dsph_gaugeres &r=gaugeres[i]; r.timestep=timestep; r.value[0]=x; r.value[1]=y; r.value[2]=z; r.computed=1;
    const double cdt=gauges[i].computedt;                                  // JGaugeItem::SetTimeStep (S/JDsGaugeItem.cpp:204-211)
    if(cdt){ const unsigned nt=unsigned(timestep/cdt); gaugenext[i]=cdt*nt; if(gaugenext[i]<=timestep)gaugenext[i]=cdt*(nt+1); }
  };
  for(size_t kq=0;kq<vel.size();kq++)done(vel[kq],ho[r_vel+4*kq],ho[r_vel+4*kq+1],ho[r_vel+4*kq+2]);
  for(size_t kq=0;kq<swl.size();kq++)done(swl[kq],ho[r_swl+3*kq],ho[r_swl+3*kq+1],ho[r_swl+3*kq+2]);
  for(size_t kq=0;kq<mxz.size();kq++)done(mxz[kq],0.f,0.f,ho[r_mz+kq]);
  for(size_t kq=0;kq<frc.size();kq++)done(frc[kq],ho[r_frc+3*kq],ho[r_frc+3*kq+1],ho[r_frc+3*kq+2]);
  if(res)memcpy(res,gaugeres.data(),ng*sizeof(dsph_gaugeres));
  return ndue;
}

} // namespace dsph
