// driver.h -- JSph / JSphGpuSingle::Run work-alike (S/JSphGpuSingle.cpp:808-883): case loading, time loop, PART output.
#pragma once
#include <string>
namespace dsph {
struct RunOptions {
  std::string dir=".", casename, dirout="out";
  int device=0;
  int nsteps=0;          // >0: stop after nsteps (reference option -nsteps)
  bool savepart=true;    // false: reference option -sv:none
  bool verbose=true;
  double timemax=-1;     // >0 overrides TimeMax (reference option -tmax)
  double timeout=-1;     // >0 overrides TimeOut (reference option -tout)
  int forcevariant=0;    // 0: default
  int partbegin=-1;      // >=0: restart from <partbegindir>/Part_NNNN.bi4 (reference option -partbegin:N dir, S/JSph.cpp:2005-2040)
  std::string partbegindir;
  int mgpu=0;            // >1: x-slab run on the GPUs device..device+mgpu-1 of this process, one host thread each (-mgpu:N)
  int rebalancesteps=0;  // slab runs: steps between re-balancing attempts (0: 50)
  double boundweight=0;  // slab runs: cost of a boundary particle relative to a fluid one (0: 0.05)
  unsigned piecemax=0;   // particles per PART file piece (0: 80 M, the 4 GB limit of the container)
};
struct RunSummary { unsigned np0=0; unsigned long long steps=0; double simtime=0, runtime_s=0, mpartsteps_s=0; unsigned parts=0; unsigned long long npout=0; };
RunSummary RunCase(const RunOptions &opt);
}
