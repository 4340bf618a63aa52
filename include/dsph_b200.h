/* dsph_b200.h -- C ABI of the B200-native per-step particle hot path of DualSPHysics
 * (cell-linked list -> Interaction_Forces -> ComputeStep).
 *
 * The reference (wjin33/DualSPHysics_MultiLayer) has no FFI: its GPU seam is the set of free host
 * functions in the *_ker.h headers (namespaces cudiv::, cusph::, cusphs::, cushift::, cusphNN::) that
 * JSphGpuSingle / JSphGpu / JCellDivGpuSingle call with raw device pointers.  This header is the
 * drop-in replacement for that seam: plain C, opaque context, plain pointers and sizes, int status.
 * Each entry point cites the reference interface it replaces
 *   S/ = src/source/   M/ = src_mphase/DSPH_v5.0_NNewtonian/source/
 * INTEGRATION.md shows the binding a maintainer adds to JSphGpuSingle.cpp.
 *
 * Conventions
 *   - every function returns DSPH_OK (0) or a negative error; dsph_last_error() gives the text
 *     (the reference polls cudaGetLastError() and throws JException, S/RunExceptionGpuDef.h:20-34).
 *   - host arrays are caller-owned; device memory is owned by the context (the reference's
 *     JArraysGpu pool + JCellDivGpu buffers, S/JArraysGpu.h:73-165, S/JCellDivGpu.cpp:110-161).
 *   - particle order on the device is the cell-sorted order produced by dsph_celldiv_run().
 *   - there is NO CPU fallback: every compute entry point fails with DSPH_ERR_NOGPU without a device.
 */
#ifndef DSPH_B200_H
#define DSPH_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define DSPH_OK            0
#define DSPH_ERR_ARG      (-1)
#define DSPH_ERR_CUDA     (-2)
#define DSPH_ERR_NOGPU    (-3)
#define DSPH_ERR_STATE    (-4)
#define DSPH_ERR_IO       (-5)
#define DSPH_ERR_PHYSICS  (-6)   /* NaN dt, boundary particle out, too many particles out ... */

/* ---- particle code (typecode, 16 bit) : S/DualSphDef.h:164-226 -------------------------------- */
typedef uint16_t dsph_typecode;
#define DSPH_CODE_MASKSPECIAL   0xe000u
#define DSPH_CODE_NORMAL        0x0000u
#define DSPH_CODE_PERIODIC      0x2000u
#define DSPH_CODE_OUTIGNORE     0x4000u
#define DSPH_CODE_OUTMOVE       0x6000u
#define DSPH_CODE_OUTPOS        0x8000u
#define DSPH_CODE_OUTRHOP       0xA000u
#define DSPH_CODE_MASKTYPE      0x1800u
#define DSPH_CODE_TYPE_FIXED    0x0000u
#define DSPH_CODE_TYPE_MOVING   0x0800u
#define DSPH_CODE_TYPE_FLOATING 0x1000u
#define DSPH_CODE_TYPE_FLUID    0x1800u
#define DSPH_CODE_MASKVALUE     0x07ffu

/* ---- option enums (values of the reference's XML keys / enums) ------------------------------- */
enum { DSPH_KERNEL_CUBIC=1, DSPH_KERNEL_WENDLAND=2 };                  /* TpKernel, key Kernel            */
enum { DSPH_STEP_VERLET=1, DSPH_STEP_SYMPLECTIC=2 };                   /* TpStep, key StepAlgorithm       */
enum { DSPH_VISCO_ARTIFICIAL=1, DSPH_VISCO_LAMINARSPS=2, DSPH_VISCO_CONSTEQ=3 }; /* TpVisco               */
enum { DSPH_DDT_NONE=0, DSPH_DDT_DDT=1, DSPH_DDT_DDT2=2, DSPH_DDT_DDT2FULL=3 };   /* TpDensity, DensityDT */
enum { DSPH_SHIFT_NONE=0, DSPH_SHIFT_NOBOUND=1, DSPH_SHIFT_NOFIXED=2, DSPH_SHIFT_FULL=3 }; /* TpShifting  */
enum { DSPH_CELLMODE_FULL=1, DSPH_CELLMODE_HALF=2 };                   /* cells of 2h or h (ScellDiv 1|2) */
enum { DSPH_VELGRAD_NONE=0, DSPH_VELGRAD_FDA=1, DSPH_VELGRAD_SPH=2 };  /* M/DualSphDef.h TpVelGrad        */
enum { DSPH_INTERSTEP_VERLET=1, DSPH_INTERSTEP_SYMPREDICTOR=2, DSPH_INTERSTEP_SYMCORRECTOR=3 };

#define DSPH_MAXPHASES 10   /* M/JSphGpu_NN_ker.cu:44-47: __constant__ arrays of 10 phases */

/* One non-Newtonian phase (M/DualSphDef.h StPhaseCte + StPhaseArray; XML <special><nnphases>). */
typedef struct dsph_phase {
  uint32_t mkfluid;       /* mk of the phase = CODE_GetTypeValue(code) of its particles          */
  float rho;              /* rest density                                                        */
  float mass;             /* particle mass = rho*dp^3 (dp^2 in 2-D)                               */
  float cs0;              /* speed of sound                                                      */
  float cteb;             /* B = cs0^2*rho/gamma                                                  */
  float gamma;            /* polytropic index                                                    */
  float visco;            /* kinematic viscosity / consistency index                             */
  float tau_yield;        /* yield stress                                                        */
  float tau_max;          /* bi-viscosity cap (0: Papanastasiou regularisation)                  */
  float bi_multi;         /* bi-viscosity multiplier                                              */
  float m_nn;             /* Papanastasiou exponential growth m                                  */
  float n_nn;             /* Herschel-Bulkley power index n                                      */
  uint32_t phasetype;     /* 0 liquid, 1 gas ...                                                 */
  uint32_t reserved;
} dsph_phase;

/* Everything the hot path needs from <execution><constants>/<parameters> and the case domain.
 * Mirrors what JSph::LoadCaseConfig + ConfigConstants1/2 + LoadCaseParticles + ConfigCellDivision
 * hand to the device side (StCteInteraction S/JSphGpu_ker.h:39-74, StCteSph S/DualSphDef.h:379-410).
 * Derived constants (KernelSize, Eta2, Cs0, awen, bwen, DDTkh, DDTgz, Scell, MovLimit ...) are computed
 * inside the library exactly as S/JSph.cpp:1392-1457,1772-1779 do. */
typedef struct dsph_config {
  /* formulation */
  int32_t tkernel, tstep, tvisco, tdensity, shiftmode, cellmode, verletsteps, simulate2d;
  int32_t perix, periy, periz;            /* periodic axes (JSph.cpp:716-730)                    */
  int32_t dtallparticles, dtfixed_on, stable;
  int32_t nn_on;                          /* 1: multiphase non-Newtonian formulation (src_mphase) */
  int32_t nn_velgrad, nn_nphases;
  int32_t tboundary;                      /* low byte: TpBoundary 1 DBC (0 = DBC too), 2 mDBC (S/JSph.cpp:626-641); bit 8: MDBCCorrector */
  /* constants (float where the reference stores float, JSph.cpp:567-586) */
  float kernelh, cteb, gamma, rhopzero, massfluid, massbound;
  float gravity[3];
  float visco, viscoboundfactor, ddtvalue, shiftcoef, shifttfs;
  float coefdtmin, rhopoutmin, rhopoutmax, nn_relaxationdt;
  double dp, cflnumber, dtini, dtmin, dtfixed, simulate2dposy;
  double timemax, timeout;
  /* domain: MapRealPosMin/Max after ResizeMapLimits (JSph.cpp:2056-2066); periodic inc vectors */
  double maprealposmin[3], maprealposmax[3];
  double perixinc[3], periyinc[3], perizinc[3];
  /* case counts (JSph.cpp:952-960): ids [0,casenfixed) fixed ... fluid last */
  uint32_t casenp, casenfixed, casenmoving, casenfloat, casenfluid;
  uint32_t partsoutmax_on; float partsoutmax;
  dsph_phase phases[DSPH_MAXPHASES];
} dsph_config;

/* Result of one neighbour-list build (JCellDivGpuSingle::Divide + MakeDivDataGpu,
 * S/JCellDivGpuSingle.cpp:152-235, S/JCellDivDataGpu.h:28-68). */
typedef struct dsph_divinfo {
  uint32_t np, npb, npbok;                 /* after removing excluded particles                  */
  uint32_t npbper, npfper;                 /* periodic duplicates included in np                  */
  uint32_t npbout, npfout, npbignore, npboutignore, npfoutignore;
  uint32_t ncx, ncy, ncz, nct;             /* cells of the current cell domain                   */
  uint32_t cellmin[3], cellmax[3];         /* CellDomainMin/Max                                   */
  uint32_t domcells[3], domcellcode;       /* DomCells, DomCellCode (JSph.cpp:1794-1829)          */
  uint32_t dividefull;                     /* 1 if every particle was re-sorted                   */
  uint32_t nctt;                           /* length of the begin-cell table = 2*nct+6            */
  double scell;                            /* cell size                                           */
  double domposmin[3];
} dsph_divinfo;

/* Result of one force evaluation (ViscDtMax/AceMax/VelMax of S/JSphGpuSingle.cpp:435-486,
 * S/JSphGpu.cpp:846-849; ViscEtaDtMax of M/JSphGpuSingle.cpp:497). */
typedef struct dsph_forceres {
  float viscdtmax;      /* max(dot(dr,dv)/(rr2+eta2))                                             */
  float acemax;         /* sqrt(max |ace|^2) over normal fluid particles                          */
  float velmax;         /* sqrt(max |vel|^2) over fluid particles                                 */
  float viscetadtmax;   /* NN only                                                                */
  double dt1, dt2, dt3, dt;  /* DtVariable(), S/JSphGpu.cpp:984-1009                              */
} dsph_forceres;

/* Timers of the last step, milliseconds of device time (names of S/JDsTimersGpu.h:64-88). */
typedef struct dsph_timers {
  double nl_limits, nl_presort, nl_radixsort, nl_cellbegin, nl_sortdata, nl_outcheck;
  double cf_preforces, cf_forces, su_computestep, su_periodic, su_shifting, total;
  uint64_t kernel_launches;
} dsph_timers;

typedef struct dsph_ctx dsph_ctx;

/* ---- library --------------------------------------------------------------------------------- */
const char* dsph_version(void);
const char* dsph_last_error(void);
int dsph_device_count(void);

/* ---- context : replaces JSphGpuSingle ctor / SelecDevice / Alloc*Memory
 *      (S/JSphGpu.cpp:622-630, S/JSphGpu.cpp:246-349) ---------------------------------------- */
int dsph_ctx_create(dsph_ctx **out,int device,const dsph_config *cfg,uint32_t np_capacity);
int dsph_ctx_destroy(dsph_ctx *ctx);
/* replaces cusph::CteInteractionUp (S/JSphGpu_ker.h:191) + cusphNN::CteInteractionUp_NN; may be called
 * again when a constant changes (DDT ramp, S/JSphGpuSingle.cpp:520-539). */
int dsph_constants_upload(dsph_ctx *ctx,const dsph_config *cfg);

/* ---- particle state : replaces JSphGpu::ParticlesDataUp / ParticlesDataDown
 *      (S/JSphGpu.cpp:563-616) and LoadCodeParticles/LoadDcellParticles (S/JSph.cpp:1257,1690) --- */
int dsph_particles_upload(dsph_ctx *ctx,uint32_t np,uint32_t npb,const uint32_t *idp,
                          const dsph_typecode *code /*may be NULL: derived from idp*/,
                          const double *pos_xyz,const float *velrhop_xyzw);
/* optional restart of a mid-state: Verlet VelrhopM1 (NULL -> copy of velrhop, S/JSphGpu.cpp:800) */
int dsph_particles_set_velrhopm1(dsph_ctx *ctx,const float *velrhopm1_xyzw);
/* optional restart of a mid-state: SPS stress tensor (xx,xy,xz,yy,yz,zz) per particle, laminar+SPS only
 * (NULL state = zeros, JSphGpu::InitRunGpu S/JSphGpu.cpp:801) */
int dsph_particles_set_spstau(dsph_ctx *ctx,const float *tau_6);
/* downloads in the current (cell-sorted) order; any pointer may be NULL */
int dsph_particles_download(dsph_ctx *ctx,uint32_t first,uint32_t count,uint32_t *idp,dsph_typecode *code,
                            uint32_t *dcell,double *pos_xyz,float *velrhop_xyzw);
uint32_t dsph_np(const dsph_ctx *ctx);
uint32_t dsph_npb(const dsph_ctx *ctx);

/* ---- neighbour list : replaces JSphGpuSingle::RunCellDivide = RunPeriodic + JCellDivGpuSingle::Divide
 *      + cudiv::PreSortFull/PreSortFluid/Sort/CalcBeginEndCell/SortDataParticles + cusphs::UpdatePosCell
 *      (S/JSphGpuSingle.cpp:249-417, S/JCellDivGpu_ker.h:40-59, S/JCellDivGpuSingle_ker.h) ---------- */
int dsph_celldiv_run(dsph_ctx *ctx,int updateperiodic,dsph_divinfo *info);
/* begin-cell table in the CPU form unsigned[2*nct+6] (S/JCellDivCpu.h:80,141,153); test/inspection */
int dsph_celldiv_get_begincell(dsph_ctx *ctx,uint32_t *begincell,uint32_t capacity);

/* ---- forces : replaces JSphGpu::PreInteraction_Forces + cusph::Interaction_Forces
 *      (+cusphNN::Interaction_ForcesNN) + ReduMaxFloat/ComputeAceMax + JSphGpu::DtVariable
 *      (S/JSphGpu.cpp:811-854, S/JSphGpu_ker.h:200, S/JSphGpuSingle.cpp:435-514, S/JSphGpu.cpp:984) -- */
int dsph_forces_run(dsph_ctx *ctx,int interstep,dsph_forceres *res);
/* ace (xyz) and ar of the last evaluation, sorted order; test/inspection */
int dsph_forces_download(dsph_ctx *ctx,uint32_t first,uint32_t count,float *ace_xyz,float *ar);

/* ---- step update : replaces JSphGpu::ComputeVerlet / ComputeSymplecticPre / ComputeSymplecticCorr
 *      = cusphs::ComputeStepVerlet|SymplecticPre|SymplecticCor + cusph::ComputeStepPos|Pos2
 *      + JSphShifting::RunGpu (S/JSphGpu.cpp:874-978, S/JSphGpuSimple_ker.cu:129-445,
 *      S/JSphGpu_ker.cu:1598-1690) ------------------------------------------------------------- */
int dsph_step_verlet(dsph_ctx *ctx,double dt);
int dsph_step_sympre(dsph_ctx *ctx,double dt);
int dsph_step_symcor(dsph_ctx *ctx,double dt);

/* ---- whole time step : replaces JSphGpuSingle::ComputeStep_Sym / ComputeStep_Ver + the
 *      RunCellDivide(true) that closes the step (S/JSphGpuSingle.cpp:548-596,853-861).
 *      Returns the dt used; *res (optional) holds the maxima of the last force evaluation. ------- */
int dsph_step(dsph_ctx *ctx,double *stepdt,dsph_forceres *res);
double dsph_symplectic_dtpre(const dsph_ctx *ctx);
int dsph_set_symplectic_dtpre(dsph_ctx *ctx,double dt);
int dsph_get_timers(dsph_ctx *ctx,dsph_timers *t,int reset);
/* CUDA-event stopwatch on the context's own stream (JTimerCuda S/JTimerCuda.h) and statistics of the dominant
 * kernel (fluid Interaction_Forces): accumulated device ms and number of launches since the last reset. */
int dsph_timer_start(dsph_ctx *ctx);
int dsph_timer_stop(dsph_ctx *ctx,double *ms);
int dsph_fluid_kernel_stats(dsph_ctx *ctx,double *ms_sum,uint64_t *nlaunch,int reset);
/* selects the fluid force-kernel implementation (profiles/r01_forces.md): 1 plain one-thread-per-particle IEEE loop (the reference's
 * structure and pair order), 2 lane FIFO, 3 warp-cooperative, 4 division-free body, 5/6 deferred pair ring of 64/32 slots, 7 = 6 with
 * OR-addressed ring slots and table-driven cell offsets, 8 = 7 + packed fp32x2 arithmetic (1,2 and 4-8: the oracle's summation order;
 * 4-8 bit-identical to each other), 9 = 8 with opposite rows of neighbour cells visited back to back (default; rounding-level
 * difference to 8).  Multiphase runs: 1 = plain loop, anything else = ring kernels. */
int dsph_set_force_variant(dsph_ctx *ctx,int variant);
int dsph_get_force_variant(const dsph_ctx *ctx);
/* TEST SWITCH, off by default.  The reference CPU binary (g++ -O3 -ffast-math) evaluates the pressure of its last Np-4*((Np-1)/4)
 * particles with rhop/rhop0 and of all others with rhop*(1/rhop0) (scalar epilogue of a vectorised loop, JSphCpu::PreInteractionVars_Forces
 * S/JSphCpu.cpp:430-470).  on=1 reproduces that index-dependent rounding so that dumps of that binary can be compared at the 1e-5
 * level on cases of a few thousand particles; the product path keeps one rounding rule for every particle. */
int dsph_set_refbinary_tail(dsph_ctx *ctx,int on);   /* the variant the next dsph_forces_run will launch (after device-capability fallbacks) */
int dsph_set_verletstep(dsph_ctx *ctx,int verletstep);
double dsph_last_force_ms(const dsph_ctx *ctx);
/* pair counters of the last force evaluation (JDsPips, S/JDsPips_ker.cu:66): real & checked pairs
 * for fluid and boundary particles; used by the roofline byte model. */
int dsph_pips(dsph_ctx *ctx,uint64_t *fluid_real,uint64_t *fluid_checked,uint64_t *bound_real,uint64_t *bound_checked);

/* ---- probes: pressure, velocity and kernel sum at arbitrary points, Shepard-corrected SPH interpolation
 *      over the fluid particles of the current cell list (needs dsph_celldiv_run).  Stand-in for the reference's MeasureTool
 *      post-processing (-points <file> -vars:+press,+vel,+kcorr; binary only in the reference) with the interpolation of the
 *      in-solver velocity gauge (JGaugeVelocity::CalculeCpuT S/JDsGaugeItem.cpp:325-367) divided by the kernel sum.
 *      Any output pointer may be NULL.  Points outside the cell domain give zeros. ----------------------------------- */
int dsph_measure_points(dsph_ctx *ctx,uint32_t npoints,const double *points_xyz,float *press,float *vel_xyz,float *kcorr);

/* ---- x-slab multi-GPU (new: the reference has no multi-GPU code, only the hooks DCEL_CodeDomLeft/Right
 *      S/JDsDcellDef.h:30-31 and MGDIV_X S/DualSphDef.h:494-499).  One context per GPU owns the cell columns
 *      [xcell_lo,xcell_hi) of the domain.  Before every dsph_celldiv_run the caller runs
 *        dsph_mgpu_count -> exchange counts -> dsph_mgpu_pack -> NCCL send/recv of the records -> dsph_mgpu_unpack
 *      which moves particles that left the slab to the neighbour (migrants) and appends copies of the neighbour's
 *      particles within 2h of the cut (ghosts: CODE_PERIODIC, neighbours only, dropped by the next divide), and after
 *      dsph_forces_launch an all-reduce(max) over dsph_forces_maxima_ptr (4 x uint32 bit patterns of non-negative floats)
 *      gives every rank the global dt in dsph_forces_finish.  Record = 88 or 112 bytes (dsph_mgpu_record_bytes).
 *      counts[8] = {boundary ghosts, boundary migrants, fluid migrants, fluid ghosts} for the left neighbour, then for
 *      the right one; a direction buffer holds its four groups back to back.  Buffers are caller-owned device memory.
 *      Re-balancing: dsph_mgpu_column_counts gives the owned particles per cell column; after an all-reduce the caller
 *      picks new cuts (each cut moves by less than the width of its two slabs) and calls dsph_mgpu_config again: the next
 *      count/pack/unpack migrates the fluid AND boundary particles that are outside the new range. ------------------ */
int dsph_mgpu_config(dsph_ctx *ctx,int rank,int nranks,uint32_t xcell_lo,uint32_t xcell_hi);
/* bytes of one record for this context: 88 = pos 24 + velrhop 16 + pospre 24 + velrhoppre|velrhopm1 16 + idp 4 + code 4,
 * + 24 (SpsTau) when the context runs laminar+SPS viscosity (single-phase), whose stress tensor is particle state */
uint32_t dsph_mgpu_record_bytes(const dsph_ctx *ctx);
int dsph_mgpu_count(dsph_ctx *ctx,uint32_t counts[8]);
int dsph_mgpu_pack(dsph_ctx *ctx,void *send_left,void *send_right);
int dsph_mgpu_unpack(dsph_ctx *ctx,const uint32_t counts[8],const void *recv_left,const void *recv_right);
/* owned (CODE_NORMAL) particles per cell column of the domain, fluid and boundary separately; ncol = domcells[0] */
int dsph_mgpu_column_counts(dsph_ctx *ctx,uint32_t *fluid,uint32_t *bound,uint32_t ncol);

/* ---- 5b. native transport (NCCL inside the library, no Python / torch on the path).  SURVEY.md 8b asked for
 *      dsph_ctx_create(device, ncclComm*), dsph_halo_exchange, dsph_migrate: here the communicator is attached to a context after
 *      creation (dsph_mgpu_comm_attach: a ncclComm_t the host application owns; or dsph_mgpu_comm_init: the library creates it from
 *      a 128-byte ncclUniqueId that rank 0 obtained with dsph_mgpu_unique_id and distributed by any means), halo exchange and
 *      migration are one call (dsph_mgpu_exchange: classify -> class totals to the neighbours -> ONE host wait -> pack ->
 *      ncclSend/ncclRecv of the records -> unpack, all on the context's stream), and
 *      Default transport of the records = PEER MEMORY: at the first exchange the ranks trade the addresses / CUDA IPC handles of
 *      their receive buffers and from then on the records go from the pack kernel (local staging) into the neighbour's buffer over NVLink with one coalesced copy kernel
 *      (class bases from the device-side totals: no host wait between classification and packing); a 20-byte ncclSend/ncclRecv
 *      token {totals, used} after the kernel announces "records landed", then the ONE host wait, then unpack.  Records that do
 *      not fit the neighbour's buffer fall back to ncclSend/ncclRecv for that exchange; DSPH_MGPU_DIRECT=0 (environment) or GPUs
 *      without peer access keep the ncclSend/ncclRecv transport altogether (dsph_mgpu_transport tells which is active).
 *        dsph_mgpu_celldiv = dsph_mgpu_exchange + periodic copies along y / z + dsph_celldiv_run(updateperiodic=0)
 *      Periodic cases: along y and z every rank copies what it holds after the exchange (owned particles, arrivals and x-halo copies
 *      alike); along x the ranks form a ring -- rank 0 and rank nranks-1 are neighbours through the periodic face, halo copies cross it
 *      shifted by the period, and a particle that leaves through the face (wrapped by the update) migrates the short way round.  The
 *      cuts still cover all the cell columns of the domain (the margin columns outside the real domain belong to the end ranks).
 *        dsph_mgpu_forces  = dsph_forces_launch + ncclAllReduce(max) of the four maxima + dsph_forces_finish
 *        dsph_mgpu_step    = dsph_step with those two in place of RunCellDivide / Interaction_Forces (S/JSphGpuSingle.cpp:548-596)
 *      dsph_mgpu_rebalance (collective) moves the cuts to the measured load: per-column histogram of owned particles (fluid 1,
 *      boundary bound_weight) all-reduced, scaled per slab by rank_cost (e.g. dsph_mgpu_measured_cost: device ms of this rank's own
 *      kernels since the last call; <=0 on any rank = histogram only); *changed=1 -> call dsph_mgpu_celldiv next.  cuts_out (optional)
 *      receives the nranks+1 cut columns.  NCCL is loaded with dlopen at the first of these calls. ------- */
int dsph_mgpu_unique_id(void *id128);
int dsph_mgpu_nccl_version(void);
int dsph_mgpu_comm_init(dsph_ctx *ctx,const void *id128,int rank,int nranks);
int dsph_mgpu_comm_attach(dsph_ctx *ctx,void *ncclcomm,int rank,int nranks);
int dsph_mgpu_config_cuts(dsph_ctx *ctx,int rank,int nranks,const uint32_t *cuts);   /* = dsph_mgpu_config with all nranks+1 cut columns known */
uint32_t dsph_mgpu_halo_cells(const dsph_ctx *ctx);                                    /* halo width in cells; a slab must be >= 2*halo+1 columns wide */
int dsph_mgpu_exchange(dsph_ctx *ctx);
int dsph_mgpu_celldiv(dsph_ctx *ctx,dsph_divinfo *info);
int dsph_mgpu_forces(dsph_ctx *ctx,int interstep,dsph_forceres *res);
int dsph_mgpu_step(dsph_ctx *ctx,double *stepdt,dsph_forceres *res);
int dsph_mgpu_rebalance(dsph_ctx *ctx,double bound_weight,double rank_cost,int *changed,uint32_t *cuts_out);
double dsph_mgpu_measured_cost(dsph_ctx *ctx);
/* host only (no device, no context): nranks+1 cut columns that minimise the work of the busiest slab for the given work per cell
 * column, every slab at least minwidth columns wide (what dsph_mgpu_rebalance uses).  DSPH_ERR_ARG when nranks*minwidth > ncol. */
int dsph_mgpu_balanced_cuts(const double *column_work,uint32_t ncol,int nranks,uint32_t minwidth,uint32_t *cuts_out);
/* per-phase host timers of the exchange (adds stream synchronisations while on): classify, totals exchange + host wait, pack,
 * records send/recv, unpack, dt all-reduce [ms accumulated since the last call]; records sent so far */
int dsph_mgpu_profile(dsph_ctx *ctx,int on,double phases_ms[6],uint64_t *sent_records);
int dsph_mgpu_sent_records(dsph_ctx *ctx,uint64_t *sent_records);
/* *direct: 1 = records travel through peer memory, 0 = ncclSend/ncclRecv, -1 = not negotiated yet (before the first exchange);
 * *fallbacks: exchanges of this rank that used ncclSend/ncclRecv for a side because the records did not fit the peer buffer */
int dsph_mgpu_transport(dsph_ctx *ctx,int *direct,uint64_t *fallbacks);
/* the dsph_divinfo of the most recent divide (dsph_step / dsph_mgpu_step run their own divides) */
int dsph_celldiv_last_info(dsph_ctx *ctx,dsph_divinfo *info);
/* 1 when the ring kernels (variants 9, 10, multiphase ring) can run on this device/driver (shared-window probe at creation);
 * 0 = the plain kernels run instead, a large performance cliff that the driver reports in Run.out */
int dsph_ringbase_ok(const dsph_ctx *ctx);
int dsph_forces_launch(dsph_ctx *ctx,int interstep);
void* dsph_forces_maxima_ptr(dsph_ctx *ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on: lets the caller order its collective after
 * dsph_forces_launch and before dsph_forces_finish with stream waits instead of a device-wide synchronisation */
void* dsph_cuda_stream(dsph_ctx *ctx);
int dsph_forces_finish(dsph_ctx *ctx,dsph_forceres *res);
/* domain constants needed to place the cuts: cell size, origin and number of cell columns (JSph::ConfigCellDivision) */
int dsph_domain_info(const dsph_config *cfg,double *scell,double domposmin[3],uint32_t domcells[3]);

/* ---- mDBC boundary correction (JSphCpu::InteractionMdbcCorrectionT2 S/JSphCpu.cpp:1020-1190, GPU twin
 *      KerInteractionMdbcCorrection_Fast/_Dbl S/JSphGpu_ker.cu:912-1282): with cfg.tboundary = 2 every force evaluation (the corrector
 *      one only with MDBCCorrector) first extrapolates the fluid density to the ghost node of every boundary particle (position + its
 *      normal vector, first-order corrected by the 4x4 / 3x3 moment matrix) and mirrors it back (slip mode velocity = 0, the only one
 *      the reference accepts, S/JSph.cpp:788).  The normal vectors are per-particle data like the velocities: upload them (xyz per
 *      particle of the last dsph_particles_upload, boundary particles first; fluid entries ignored) in the units of
 *      JSph::ConfigBoundNormals (S/JSph.cpp:1300-1340: particle -> ghost node = 2 x the GenCase vector particle -> boundary limit). */
#define DSPH_BOUNDARY_DBC 1
#define DSPH_BOUNDARY_MDBC 2
#define DSPH_BOUNDARY_MDBCCORRECTOR 0x100
int dsph_particles_set_boundnormals(dsph_ctx *ctx,const float *normal_xyz,uint32_t count);
int dsph_particles_get_boundnormals(dsph_ctx *ctx,float *normal_xyz,uint32_t count);   /* current (sorted) order; a matrix motion rotates them */
int dsph_set_mdbc_threshold(dsph_ctx *ctx,float threshold);      /* command-line option -mdbc_threshold of the reference (default 0) */
/* <dir>/<casename>_Normals.nbi4 (JPartNormalData, S/JPartNormalData.cpp:222-262): vectors particle -> boundary limit of the
 * boundary particles in id order; *n = number stored in the file (0 when there is none), at most max vectors are written */
int dsph_case_normals(const char *dir,const char *casename,double *normal_xyz,uint32_t max,uint32_t *n);

/* ---- moving boundaries: JSphGpu::RunMotion (S/JSphGpu.cpp:1035-1075 -> cusph::CalcRidp / MoveLinBound / MoveMatBound;
 *      CPU twin S/JSphCpu.cpp:1660-1730,1762-1805).  dsph_motiondata = StMotionData (S/JDsMotionDef.h:33-44): what the host's motion
 *      objects (the reference's JDsMotion::ProcesTime + GetMotionData, S/JSph.cpp:2308-2322, called at the start of ComputeStep_Sym or
 *      after DtVariable in ComputeStep_Ver) computed for [TimeStep,TimeStep+dt).  dsph_motion_run applies it where JSphGpuSingle::Run
 *      calls RunMotion (S/JSphGpuSingle.cpp:859: after ComputeStep and the gauges, before RunCellDivide): positions through the same
 *      UpdatePos as the fluid (periodic wrap, new cell), velocity = linvel or displacement/stepdt, mDBC normals rotated with a matrix
 *      motion; the next divide re-sorts the boundary (BoundChanged).  cfg.casenmoving ids follow the cfg.casenfixed fixed ones.
 *      Floating bodies (RunFloating) are not built: contexts with cfg.casenfloat != 0 are refused. ----------------------------------- */
#define DSPH_MOTION_NONE   0
#define DSPH_MOTION_LINEAR 1
#define DSPH_MOTION_MATRIX 2
typedef struct dsph_motiondata {
  uint32_t idbegin, count;    /* ids of the object's particles */
  int32_t type, reserved;
  double linmov[3], linvel[3];
  double matmov[16];          /* tmatrix4d a11..a44, row major */
} dsph_motiondata;
int dsph_motion_run(dsph_ctx *ctx,const dsph_motiondata *objs,uint32_t nobj,double stepdt);
/* host only: the motion laws of <case><execution><motion> (objects = the <moving> blocks of <particles>) evaluated like JDsMotion in
 * simple mode, for hosts without the reference's JMotion library (bin/dsph_b200): wait, mvrect, mvrectace, mvrot, mvrotace, mvcir, mvcirace,
 * mvrectsinu, mvrotsinu, mvcirsinu, mvrectfile (= mvfile, mvpredef) and mvrotfile (data files next to the case) with next-chains, begin
 * events and nested <obj> / <objreal> objects (a child moves with its parent); other elements are refused.  *nobj = moving objects; out[] needs nobj entries. */
typedef struct dsph_motionset dsph_motionset;
int dsph_case_motion_open(const char *dir,const char *casename,dsph_motionset **ms,uint32_t *nobj);
int dsph_case_motion_process(dsph_motionset *ms,double timestep,double dt,dsph_motiondata *out,int *active);
int dsph_case_motion_close(dsph_motionset *ms);
/* dsph_step / dsph_mgpu_step evaluate the attached laws themselves (CalcMotion at the start of a Symplectic step or after DtVariable of
 * a Verlet step, RunMotion before the closing divide) with the run time they advance (start value: dsph_gauges_set_time).  The set
 * must outlive the context or be detached with ms = NULL.  Hosts that call the stages one by one use dsph_motion_run. */
int dsph_motion_attach(dsph_ctx *ctx,dsph_motionset *ms);

/* ---- excluded particles (JSph::AddParticlesOut <- JSphGpuSingle::RunCellDivide, S/JSphGpuSingle.cpp:405-412): with keep = 1 every
 *      divide copies the fluid particles it excludes (out of the domain, RhopOut, speed limit) to the host; dsph_excluded_take hands
 *      them over (at most max; *n = how many were pending) and empties the store.  Arrays may be NULL to discard. ------------------- */
int dsph_excluded_keep(dsph_ctx *ctx,int keep);
int dsph_excluded_take(dsph_ctx *ctx,uint32_t max,uint32_t *n,uint32_t *idp,dsph_typecode *code,double *pos_xyz,float *velrhop_xyzw);

/* ---- in-solver gauges: what JGaugeVelocity / JGaugeSwl / JGaugeMaxZ / JGaugeForce compute (CPU:
 *      S/JDsGaugeItem.cpp:325-367,575-600,736-776,986-1041; GPU: S/JDsGauge_ker.cu:50,168,250,311), evaluated on the device for all the
 *      gauges of a step at once.  Definitions come from <special><gauges> of Case.xml (S/JDsGaugeSystem.cpp:186-275; dsph_case_gauges)
 *      or from the caller.  Timing follows JGaugeItem (S/JDsGaugeItem.h:164-165, S/JDsGaugeItem.cpp:204-211): a gauge is computed at
 *      a step when timestep >= its next compute time inside [computestart, computeend].  dsph_step runs the configured gauges where
 *      the reference runs them: after the step update, BEFORE the closing divide (S/JSphCpuSingle.cpp:1094-1098), with
 *      timestep = (time set by dsph_gauges_set_time) + dt. ----------------------------------------------------------------------  *      In x-slab runs dsph_gauges_run is COLLECTIVE (every rank calls it at the same instants): each rank sums over the particles it owns
 *      and the partial results are all-reduced over the communicator of the context; dsph_mgpu_step evaluates the gauges after the exchange
 *      of its closing divide (fresh halo copies).  Not available for periodic cases in slab runs. */
enum { DSPH_GAUGE_VEL=1, DSPH_GAUGE_SWL=2, DSPH_GAUGE_MAXZ=3, DSPH_GAUGE_FORCE=4 };
typedef struct dsph_gauge {
  int32_t type; int32_t output;                 /* output: results go to Gauges<Type>_<name>.csv (driver) */
  char name[32];
  double computestart, computeend, computedt;   /* computedt = 0: every step */
  double outputstart, outputend, outputdt;
  double point0[3], point2[3];                  /* vel: point0 | swl: point0 -> point2 | maxz: point0 */
  double pointdp; float masslimit;              /* swl: spacing of the line points, mass threshold (default 0.5*massfluid, 0.4 in 2-D) */
  float distlimit; double height;               /* maxz */
  uint32_t mkbound, idbegin, count; dsph_typecode code; uint16_t pad_;   /* force: boundary mk block (id range and particle code) */
} dsph_gauge;
typedef struct dsph_gaugeres {
  double timestep;      /* time of the last evaluation */
  float value[3];       /* vel: velocity | swl: surface point | maxz: (0,0,zmax) | force: force [N] */
  int32_t computed;     /* 1 when the last dsph_gauges_run / dsph_step evaluated this gauge */
} dsph_gaugeres;
int dsph_gauges_config(dsph_ctx *ctx,uint32_t n,const dsph_gauge *gauges);
/* evaluates the gauges that are due at `timestep` (all of them when force_all != 0) on the current state and cell table;
 * results (n entries, every gauge: the last result of those not evaluated now) may be NULL */
int dsph_gauges_run(dsph_ctx *ctx,double timestep,int force_all,dsph_gaugeres *results);
int dsph_gauges_results(dsph_ctx *ctx,dsph_gaugeres *results);
int dsph_gauges_set_time(dsph_ctx *ctx,double timestep);   /* simulation time at the START of the next dsph_step */
/* gauge definitions of <dir>/<casename>.xml; returns the number found in *n (at most max are written) */
int dsph_case_gauges(const char *dir,const char *casename,dsph_gauge *out,uint32_t max,uint32_t *n);

/* ---- case tooling (host only, no GPU): GenCase is a missing blob in the reference, these build the
 *      synthetic benchmark cases of SURVEY.md 8(d) and read/write Case.xml + Case.bi4 ------------- */
typedef struct dsph_case {
  dsph_config cfg;
  uint32_t np, npb;
  uint32_t *idp; double *pos_xyz; float *velrhop_xyzw;   /* owned by the library: dsph_case_free */
  double caseposmin[3], caseposmax[3];
  char name[64];
  uint32_t nfluidblocks;                       /* fluid mk blocks (= phases in the multiphase fork), in id order */
  uint32_t fluidblock_count[DSPH_MAXPHASES];
} dsph_case;
/* kind: "dambreak" (DB(N)), "periodicbox" (PB(N)), "nn2layer" (NN(N)); np_target approximate */
int dsph_case_make(dsph_case *out,const char *kind,uint64_t np_target,uint32_t seed);
/* same case, but only the particles of the cell columns [cx_lo,cx_hi) are generated (ids, counts and constants are
 * those of the full case): every rank of a multi-GPU run generates its own x-slab */
int dsph_case_make_slab(dsph_case *out,const char *kind,uint64_t np_target,uint32_t seed,uint32_t cx_lo,uint32_t cx_hi);
int dsph_case_free(dsph_case *c);
int dsph_case_save(const dsph_case *c,const char *dir,const char *casename); /* Case.xml + Case.bi4 */
int dsph_case_load(dsph_case *out,const char *dir,const char *casename);
/* one PART file (Part_NNNN.bi4 written by this driver or by the reference, JPartDataBi4 S/JPartDataBi4.cpp): particles in
 * file order (excluded particles are not in it), head constants into cfg; free with dsph_case_free */
int dsph_part_load(dsph_case *out,const char *file,double *timestep,uint32_t *step);

/* ---- driver: the JSph/JSphGpuSingle::Run work-alike (S/JSphGpuSingle.cpp:808-883): loads
 *      <dir>/<casename>.xml+.bi4, runs to TimeMax (or nsteps>0), writes Part_NNNN.bi4, Run.out, Run.csv
 *      into dirout. ---------------------------------------------------------------------------- */
int dsph_run_case(const char *dir,const char *casename,const char *dirout,int device,int nsteps,int savepart);
/* the same with the remaining command-line options of the reference that concern this path (R/doc/help/DualSPHysics5.2_Help.out):
 * -tmax, -tout, -partbegin:N dir (restart from dir/Part_NNNN.bi4: particle state, time and the Symplectic dt of the next
 * step, S/JSph.cpp:2005-2040).  Zero-initialise, then set what is needed: timemax/timeout <= 0 keep the XML values,
 * restart = 0 starts from Case.bi4 (partbegin is ignored), partbegindir NULL = dirout. */
typedef struct dsph_runopts {
  const char *dir, *casename, *dirout;
  int32_t device, nsteps, savepart, forcevariant;
  double timemax, timeout;
  int32_t restart, partbegin;
  const char *partbegindir;
  /* x-slab run: mgpu > 1 uses the GPUs device..device+mgpu-1 of this process, one host thread and one PART piece
   * (Part_pXX_NNNN.bi4, AxisDiv = X; S/JPartDataBi4.h:81,122-133) per GPU; exchange and dt all-reduce through the library's NCCL
   * transport (section 5b).  rebalancesteps: steps between re-balancing attempts (0 = 50); boundweight: 0 = 0.05.
   * piecemax: particles per PART file piece of a single-GPU run (0 = 80 M: the 4 GB limit of the container, S/JBinaryData.cpp:1223-1235) */
  int32_t mgpu, rebalancesteps;
  double boundweight;
  uint32_t piecemax, reserved_;
} dsph_runopts;
int dsph_run_case_ex(const dsph_runopts *opts);

/* ---- PART output set, host only (no device needed): what dsph_run_case writes for every PART, callable on host arrays.
 *      Part_Head.ibi4 (S/JPartDataHead.cpp:208-272), Part_NNNN.bi4 or Part_pXX_NNNN.bi4 pieces of at most piecemax particles
 *      (S/JPartDataBi4.cpp:98-102,437-439), PartInfo.ibi4 (S/JPartDataBi4.cpp:445-450) and, when nout > 0, PartOut_000.obi4
 *      (S/JPartOutBi4Save.cpp:212-231).  Particles in `c` must have boundary ids first.  out_* may be NULL when nout = 0;
 *      out_motive: 1 position, 2 density, 3 velocity. */
int dsph_partset_save(const dsph_case *c,const char *dirout,uint32_t cpart,double timestep,uint32_t step,double symplecticdtpre,uint32_t piecemax,
                      uint32_t nout,const uint32_t *out_idp,const double *out_pos_xyz,const float *out_velrhop_xyzw,const uint8_t *out_motive);
/* every piece of PART cpart of dir, concatenated (boundary particles first); free with dsph_case_free */
int dsph_partset_load(dsph_case *out,const char *dir,uint32_t cpart,double *timestep,uint32_t *step,uint32_t *npiece);

#ifdef __cplusplus
}
#endif
#endif /* DSPH_B200_H */
