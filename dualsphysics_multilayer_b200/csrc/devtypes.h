// devtypes.h -- PODs shared by the host engine and the CUDA kernels (sm_100a).
#pragma once
#include <cstdint>

namespace dsph {

typedef uint16_t typecode;

// Constants of the force / step kernels, passed BY VALUE as a kernel parameter (constant bank, warp-broadcast
// reads).  Replaces the reference's __constant__ StCteInteraction (S/JSphGpu_ker.h:39-74) uploaded by
// cusph::CteInteractionUp: no symbol upload, no per-TU constant copies, safe with several contexts.
struct DevCtes {
  // kernel (S/FunSphKernel.h:50-80,199-210)
  float kernelh, ovkernelh, kernelsize2, eta2, awen, bwen;
  float cub_a2, cub_a24, cub_c1, cub_d1, cub_c2, cub_od_wdeltap;
  float bwen_ovh, cub_c1_ovh, cub_d1_ovh;   // kernel constants divided by h (division-free gradient, forces_v3.cuh)
  // EOS / masses
  float cteb, gamma, ovgamma, rhopzero, ovrhopzero, massfluid, massbound;
  double cteb_d, gamma_d;
  // viscosity, DDT, SPS
  float cs0, visco, viscobound, ddtkh, ddtgz, spssmag, spsblin;
  float powc[6];          // series of (1+t)^(1/gamma)-1 = t*(c0+t*(c1+...)) used by DDT2/3 (see PowNearOne)
  // cells
  float scell; int scelldiv;
  float hscale;           // 2/scell: scale of the half-precision culling frame of forces_v10.cuh (cell = [-1,1]^3)
  double scell_d;
  double domposmin[3];
  unsigned domcellcode;
  unsigned domcells[3];
  // step update (S/JSphCpu.cpp:1240-1293 UpdatePos)
  double maprealposmin[3], maprealsize[3];
  double perixinc[3], periyinc[3], perizinc[3];
  int perix, periy, periz, periactive;
  int ghosts;             // periactive || x-slab multi-GPU: particles with CODE_PERIODIC are neighbour-only copies
  float movlimit, rhopoutmin, rhopoutmax;
  float gravity[3];
  int simulate2d, dtallparticles;
  // shifting (S/JSphShifting.cpp:387-417)
  int shiftmode; float shifttfs;
  int nn_on;              // multiphase non-Newtonian fork: per-phase EOS, v5.0 shifting clamp
  int mdbc;               // TBoundary = mDBC: boundary neighbours take part in DDT1 instead of switching it off (S/JSphCpu.cpp:730)
  int reftail;            // test switch (dsph_set_refbinary_tail): reproduce the pressure rounding of the reference BINARY's loop epilogue
};

#ifdef __CUDACC__
// cell code of a particle (S/JDsDcellDef.h:33-44): bit fields chosen by JDsDcell::CalcCellCode
__device__ __forceinline__ unsigned DcCellx(unsigned dcc,unsigned cel){ return cel>>((dcc>>10)&31); }
__device__ __forceinline__ unsigned DcCelly(unsigned dcc,unsigned cel){ return (cel<<(dcc>>25))>>((dcc>>5)&31); }
__device__ __forceinline__ unsigned DcCellz(unsigned dcc,unsigned cel){ return (cel<<(dcc&31))>>(dcc&31); }
__device__ __forceinline__ unsigned DcCell(unsigned dcc,unsigned cx,unsigned cy,unsigned cz){ return (cx<<((dcc>>10)&31))|(cy<<((dcc>>15)&31))|cz; }
#endif

// Phases of the multiphase non-Newtonian fork, passed by value to the kernels that need them (constant bank, indexed
// by the type value of the particle code).  Replaces the __constant__ PHASECTE/PHASEARRAY[10] of M/JSphGpu_NN_ker.cu:44-47.
struct DevPhase { float rho, mass, cs0, cteb, gamma, visco, tau_yield, tau_max, bi_multi, m_nn, n_nn, pad; };
struct DevPhases { int count; int velgrad; float lamda; float pad; DevPhase ph[10]; };

// Cell domain of the current divide; written by ker_div_prepare ON THE DEVICE (no host round trip between the
// min/max reduction and the key build), copied to the host once per divide together with the counts.
// Fields follow JCellDivCpuSingle::PrepareNct (S/JCellDivCpuSingle.cpp:109-125).
struct DivParams {
  unsigned cellmin[3], cellmax[3];
  unsigned ncx, ncy, ncz, nsheet, nct, nctt;
  unsigned boxboundignore, boxfluid, boxboundout, boxfluidout, boxboundoutignore, boxfluidoutignore;
  unsigned npbignore, npbout, npfout, npboutignore, npfoutignore;
  unsigned nptot, npfinal, npbfinal;
  unsigned blimmin[3], blimmax[3];   // cached boundary cell limits (BoundLimitCellMin/Max)
  unsigned error;                    // !=0: nct overflow
  unsigned maxcellpop;               // largest population of a (boundary or fluid) cell of this divide
};

// Maxima of one force evaluation, as order-preserving uint bit patterns of non-negative floats.
struct ForceMax {
  unsigned viscdt_bits, ace2_bits, vel2_bits, visceta_bits;
  unsigned long long pips_fluid_real, pips_fluid_checked, pips_bound_real, pips_bound_checked;
};

// ping-pong buffers of the radix sort (radixsort.cuh)
struct RadixSortBuffers {
  uint32_t *keys[2]={nullptr,nullptr}, *vals[2]={nullptr,nullptr};
  uint32_t *hist=nullptr, *chunksum=nullptr, *chunksum2=nullptr;
  uint32_t capacity=0, histcap=0, chunksumcap=0, chunksum2cap=0;   // chunksum*: scan scratch (entries), sized for max(histogram, particle count)
};

} // namespace dsph
