/*
 * xsb200.h -- C ABI of libxsb200.so: the B200-native (sm_100a) implementation of exaStamp's short-range
 * force hot path (chunk_neighbors + pair / EAM / SNAP force operators + ghost exchange).
 *
 * exaStamp exposes no C ABI: a force field is an onika::scg::OperatorNode subclass registered under a YAML
 * name (src/potential/pair_potential_template/pair_potential_impl.hxx:510-513).  Each entry point below is
 * what the `execute()` of one of those operators needs once its slots have been resolved; the reference
 * interface it replaces is cited per function (paths relative to the exaStamp source tree).  INTEGRATION.md
 * shows the ~30-line OperatorNode shim that forwards slots to these calls.
 *
 * Conventions
 *  - plain pointers and sizes only; every call returns XSB_OK (0) or an xsb_status error code and never
 *    aborts; xsb_last_error() gives the message.  One context per GPU, not thread-safe per context.
 *  - all calls enqueue work on the context's CUDA stream and return; xsb_sync() waits.  Downloads sync.
 *  - particle data: flat SoA sorted by cell (cells IJK row-major, i fastest, ghost layers included), flat
 *    index = cell_particle_offset[cell] + p  -- the layout of exanb::Grid::cell_particle_offset_data().
 *    Positions are in grid space; physical = xform * r (exanb::Domain::xform).  Quantities are in exaStamp
 *    internal units (angstrom, Da, ps, e, K : include/exaStamp/unit_system.h:28-36).
 *  - force operators ACCUMULATE (+=) into fx,fy,fz,ep,virial like the reference functors, so several can be
 *    chained under `compute_force: [...]`; xsb_zero_force_energy() is the `zero_force_energy` operator.
 *  - there is no CPU fallback: every compute entry point fails with XSB_ERR_CUDA when no sm_100 device is
 *    usable.
 */
#ifndef XSB200_H
#define XSB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xsb_ctx xsb_ctx;

typedef enum xsb_status {
  XSB_OK = 0, XSB_ERR_INVALID = 1, XSB_ERR_CUDA = 2, XSB_ERR_STATE = 3, XSB_ERR_NCCL = 4, XSB_ERR_IO = 5,
  XSB_ERR_UNSUPPORTED = 6, XSB_ERR_OVERFLOW = 7
} xsb_status;

/* exanb::Grid + exanb::Domain as seen by the force operators (accessors listed in SURVEY.md 8a row a1) */
typedef struct xsb_grid_desc {
  int32_t dims[3];            /* grid.dimension(): cells per axis INCLUDING ghost layers            */
  int32_t ghost_layers;       /* grid.ghost_layers()                                                */
  double  cell_size;          /* domain.cell_size(), grid space                                     */
  double  origin[3];          /* grid-space corner of cell (0,0,0) (grid.origin() + offset*cell)    */
  double  xform[9];           /* domain.xform(), row-major; physical = xform * grid-space           */
  int32_t xform_is_identity;  /* domain.xform_is_identity()                                         */
  int32_t pad_;
} xsb_grid_desc;

/* per-particle fields (include/exaStamp/fields.h:32-37,64-70 + exanb rx,ry,rz,fx..,vx..,id,type) */
typedef enum xsb_field {
  XSB_F_RX = 0, XSB_F_RY, XSB_F_RZ, XSB_F_FX, XSB_F_FY, XSB_F_FZ, XSB_F_EP,     /* double               */
  XSB_F_VX, XSB_F_VY, XSB_F_VZ,                                                  /* double               */
  XSB_F_VIRIAL,                                                                  /* 9 doubles/atom Mat3d */
  XSB_F_RHO_DEMB,                                                                /* double (eam_buffer.h:37-49 field "rho_dEmb") */
  XSB_F_TYPE,                                                                    /* uint8                */
  XSB_F_ID,                                                                      /* uint64               */
  XSB_F_COUNT_
} xsb_field;

/* operator flags (slots `ghost`, grid-has-virial, trigger_thermo_state, ...) */
enum {
  XSB_FLAG_GHOST   = 1,   /* slot `ghost`: also compute central atoms of ghost cells                    */
  XSB_FLAG_ENERGY  = 2,   /* accumulate ep                                                              */
  XSB_FLAG_VIRIAL  = 4,   /* accumulate per-atom virial (grid field set has field::virial)              */
  XSB_FLAG_MIXED   = 8    /* FP32 pair math, FP64 accumulation (north_star "mixed mode", tol 1e-5)      */
};

/* eam_alloy_force phase slots (src/potential/eam_potential_template/eam_potential_multimat.cu:101-105) */
enum {
  XSB_EAM_RHO = 1, XSB_EAM_RHO2EMB = 2, XSB_EAM_GHOST = 4, XSB_EAM_FORCE = 8,
  XSB_EAM_EFLAG = 16  /* trigger_thermo_state: ep (+ virial when XSB_FLAG_VIRIAL given in flags) */
};

/* pair potentials behind <pot>_compute_force / <pot>_multi_force (src/potential/pair_potentials/) */
typedef enum xsb_pair_pot {
  XSB_POT_LJ = 0,          /* params: epsilon, sigma                  (lennard_jones/.../lennard_jones.h:40-50)     */
  XSB_POT_ZBL = 1,         /* params: r1, rc, z_a, z_b                (zbl/potential.h:36-57,180-303; z from species) */
  XSB_POT_EXP6 = 2,        /* params: A, B, C, D                      (exp6/.../exp6.h:66-84)                        */
  XSB_POT_BUCKINGHAM = 3,  /* params: A, Rho, C                       (buckingham/buckingham.h:41-52)                */
  XSB_POT_YUKAWA = 4,      /* params: A, kappa                        (yukawa/.../yukawa.h:39-48)                    */
  XSB_POT_RELAX = 5,       /* params: r1, rc  (overlap relaxation ramp, relax/potential.h:43-51)                    */
  XSB_POT_ZERO = 6         /* no parameters: e = de = 0               (zero/potential.h:49-54)                       */
} xsb_pair_pot;

/* ---------------------------------------------------------------------------------------------------- */
/* context                                                                                              */
int         xsb_create(int device, xsb_ctx** out);         /* onika init_cuda / ParallelExecutionContext  */
void        xsb_destroy(xsb_ctx* ctx);
const char* xsb_last_error(const xsb_ctx* ctx);
int         xsb_sync(xsb_ctx* ctx);
const char* xsb_version(void);
uint64_t    xsb_kernel_launch_count(const xsb_ctx* ctx);   /* CUDA kernels launched by this context so far */

/* per-operator device time, measured with CUDA events on the context's own stream (onika's              */
/* `profiling.exectime`, data/config/config_defaults.msp:21-23).  tag = xsb_prof_tag.                      */
typedef enum xsb_prof_tag {
  XSB_PROF_NBR_BUILD = 0, XSB_PROF_PAIR, XSB_PROF_EAM_RHO, XSB_PROF_EAM_RHO2EMB, XSB_PROF_EAM_FORCE,
  XSB_PROF_GHOST, XSB_PROF_INTEGRATE, XSB_PROF_SNAP, XSB_PROF_MOVE, XSB_PROF_COUNT_
} xsb_prof_tag;
int xsb_profile_enable(xsb_ctx* ctx, int on);               /* also resets the accumulators */
int xsb_profile_read(xsb_ctx* ctx, int tag, double* ms_total, uint64_t* intervals);   /* syncs the stream */
/* stopwatch on the context's stream: record slot 0 (start) and 1 (stop), then read the device time      */
int xsb_timer_record(xsb_ctx* ctx, int slot);
int xsb_timer_elapsed_ms(xsb_ctx* ctx, double* ms);
/* Recorded steps.  Small systems (configs[0]: 131 k atoms, 0.1 ms of kernels per step) are bound by the launch path: the   */
/* operator calls of a regular step between two neighbour-list rebuilds -- xsb_verlet_boundary_async, xsb_ghost_update,     */
/* xsb_zero_force_energy, the force operators -- are recorded once (capture_begin ... calls ... capture_end) as a CUDA graph */
/* and re-issued with ONE launch per step (xsb_step_replay; xsb_displ_poll works as after the direct call).  No reference    */
/* counterpart: onika enqueues every operator's kernel separately (`ParallelExecutionContext`).  A recorded step freezes     */
/* what the calls pass by value (counts, pointers, cell matrix, flags): after xsb_particles_rebin / xsb_particles_set_cells /*/
/* xsb_chunk_neighbors_build / xsb_grid_set[_xform] a replay fails with XSB_ERR_STATE -- record again.  One rank only; entry  */
/* points that wait for the device or read results back fail while recording (the recording is then lost).                    */
int xsb_step_capture_begin(xsb_ctx* ctx);
int xsb_step_capture_end(xsb_ctx* ctx, int* step_id);
int xsb_step_replay(xsb_ctx* ctx, int step_id);
int xsb_step_release(xsb_ctx* ctx, int step_id);
/* roofline denominators measured on this device ("peaks must be measured on the box"):     */
/* a DFMA loop (FP64 pipe, TFLOP/s), an FFMA loop (FP32 pipe, TFLOP/s) and a 1 GiB copy (HBM, GB/s       */
/* read+write); best of 5 after a warm-up, CUDA events on the context's stream.  Any output may be null. */
int xsb_measure_peaks(xsb_ctx* ctx, double* fp64_tflops, double* fp32_tflops, double* hbm_gbs);

/* ---------------------------------------------------------------------------------------------------- */
/* a1  Grid / GridCellParticles                                                                         */
int      xsb_grid_set(xsb_ctx* ctx, const xsb_grid_desc* grid);
/* Domain::xform() changes every step under NPT / deformation (`domain->set_xform(newXForm)`,                    */
/* src/parrinellorahman/update_xform_parrinellorahman.cpp:164; SURVEY.md 8a row a1, BASELINE configs[4]) while   */
/* the grid, the particles and the neighbour list stay: only the 3x3 matrix (row-major) is replaced.  Like in  */
/* the reference, rebuilding chunk_neighbors when the deformation has eaten the skin is the caller's decision. */
int      xsb_grid_set_xform(xsb_ctx* ctx, const double xform[9]);
/* cell_particle_offset: host array of ncells+1 entries (exanb::Grid::cell_particle_offset_data()).     */
int      xsb_particles_set_cells(xsb_ctx* ctx, const uint64_t* cell_particle_offset);
uint64_t xsb_num_particles(const xsb_ctx* ctx);
uint64_t xsb_num_cells(const xsb_ctx* ctx);
uint64_t xsb_num_own_particles(const xsb_ctx* ctx);       /* particles in non-ghost cells              */
int      xsb_cell_offsets_download(xsb_ctx* ctx, uint64_t* cell_particle_offset);   /* ncells+1 entries */
/* bins n unsorted particles (HOST arrays; v*, type, id may be NULL) into the own cells of the grid:    */
/* stable sort by cell, ghost cells left empty (lattice / init_rcb_grid + particle insertion of the      */
/* decks).  Follow with xsb_ghost_comm_scheme.                                                            */
int      xsb_particles_assign(xsb_ctx* ctx, uint64_t n, const double* rx, const double* ry, const double* rz,
                              const double* vx, const double* vy, const double* vz, const uint8_t* type, const uint64_t* id);
/* Particles found outside the own cells by the last xsb_particles_assign / xsb_particles_rebin on a non-periodic (or  */
/* brick) axis.  Less than one cell outside: clamped into the border cell and counted here (the reference's             */
/* move_particles keeps such "otb" particles aside, ext exaNBody); further out the call fails with XSB_ERR_INVALID.      */
int      xsb_out_of_domain_count(xsb_ctx* ctx, uint64_t* clamped);
/* whole-array copies between host buffers and the context's device SoA (N or 9N elements)              */
int      xsb_field_upload(xsb_ctx* ctx, int field, const void* host_src);
int      xsb_field_download(xsb_ctx* ctx, int field, void* host_dst);
/* Asynchronous transfers for a host application that owns the particle arrays in PINNED host memory (the plugin    */
/* use case: positions in, forces out, every step).  Scalar double fields only (r, f, ep, v, rho_dEmb).             */
/* own_only = 1: the host arrays hold the xsb_num_own_particles() particles of the non-ghost cells in flat order      */
/* (ghost images are produced on the device by xsb_ghost_update); 0: all xsb_num_particles() slots.                   */
/*  upload:   the H2D copy runs on a copy stream concurrently with the work already enqueued on the context's         */
/*            stream (which still sees the old values); work enqueued after the call sees the new ones.               */
/*  download: snapshots the fields at this point of the context's stream, the D2H copy runs on a second copy stream   */
/*            concurrently with later work; host arrays are valid after xsb_copy_wait().                              */
/* The host arrays of a call must stay untouched until the next xsb_copy_wait() / xsb_sync() resp. until a later      */
/* upload call returns.                                                                                               */
int      xsb_fields_upload_async(xsb_ctx* ctx, int nfields, const int* fields, const void* const* host_src, int own_only);
int      xsb_fields_download_async(xsb_ctx* ctx, int nfields, const int* fields, void* const* host_dst, int own_only);
int      xsb_copy_wait(xsb_ctx* ctx);
/* device pointer of a field (zero-copy for callers that already live on the GPU, e.g. managed grids)   */
void*    xsb_field_device_ptr(xsb_ctx* ctx, int field);
/* zero_force_energy operator (src/compute/zero_force_energy.cu:98-135): fx,fy,fz,ep,(virial) = 0       */
int      xsb_zero_force_energy(xsb_ctx* ctx, int ghost);

/* ---------------------------------------------------------------------------------------------------- */
/* a2  chunk_neighbors operator (config: data/config/config_move_particles.msp:54-61; entry point        */
/*     chunk_neighbors_execute, src/particle_species/type_pair_rcut_neighbors.cpp:136)                    */
typedef struct xsb_chunk_neighbors_config {
  int32_t chunk_size;              /* power of two, 1..32 (only affects the exported stream)            */
  int32_t build_particle_offset;   /* emit the per-particle offset table in the exported stream         */
  int32_t subcell_compaction;      /* accepted, no effect (host-side AMR detail of the reference)       */
  int32_t free_scratch_memory;     /* release build scratch after each build                            */
  double  stream_prealloc_factor;  /* growth factor of the device list allocation (>=1)                 */
} xsb_chunk_neighbors_config;
/* builds, for every particle of every cell (ghost cells included), the list of (cell_b,p_b) with        */
/* 0 < |xform*(r_b-r_a)|^2 < nbh_dist_lab^2, in canonical order (ascending cell_b, then p_b).            */
int xsb_chunk_neighbors_build(xsb_ctx* ctx, double nbh_dist_lab, const xsb_chunk_neighbors_config* cfg);
int xsb_chunk_neighbors_stats(xsb_ctx* ctx, uint64_t* total_neighbors, uint32_t* max_neighbors);
/* exanb::GridChunkNeighbors in the reference uint16 per-cell stream format (decoder:                    */
/* src/rigidmol/compute_pair_rigidmol.h:154-234).  stream_off: ncells+1 offsets in uint16 units.         */
int xsb_chunk_neighbors_export_size(xsb_ctx* ctx, uint64_t* total_u16);
int xsb_chunk_neighbors_export(xsb_ctx* ctx, uint64_t* stream_off, uint16_t* data);
/* flat CSR view of the same list (device resident): counts[N] (u32), offsets[N+1] (u64), idx (u32)     */
int xsb_chunk_neighbors_download_flat(xsb_ctx* ctx, uint32_t* counts, uint64_t* offsets, uint32_t* idx);

/* ---------------------------------------------------------------------------------------------------- */
/* a3-a6  <pot>_compute_force (pair_potential_impl.hxx:39-500) and <pot>_multi_force                     */
/*        (pair_potential_force_op_multiparam.h:57-223).  ecut = e(rcut) is computed inside              */
/*        (energy_cutoff, pair_potential_impl.hxx:488-498).                                              */
int xsb_pair_force(xsb_ctx* ctx, int pot, const double* params, int nparams, double rcut, int flags);
/* pair_params: one row per unique_pair_id(type_a,type_b) = hi*(hi+1)/2+lo : {params..., rcut}           */
int xsb_pair_multi_force(xsb_ctx* ctx, int pot, int n_types, const double* pair_params, int nparams,
                         double rcut_max, int flags);

/* ---------------------------------------------------------------------------------------------------- */
/* a7  johnson_force / johnson_emb / johnson_force_reuse_emb (eam_potential.cu:69-176, johnson.h:56-166)  */
/*     params19: re fe rhoe alpha beta A B kappa lambda Fn0..Fn3 F0..F3 Fo eta.                          */
/*     phases: bit0 emb pass, bit1 emb pass covers ghost cells (ComputeGhostEmb), bit2 force pass.        */
int xsb_eam_johnson_force(xsb_ctx* ctx, const double* params19, double rcut, int phases, int flags);
/* The other analytic single-species models of the same operator template (eam_potential_template, one plugin per model:   */
/* `<name>_force`, `<name>_emb`, `<name>_force_reuse_emb`, `<name>_init`): parameters in the reference struct's order.       */
typedef enum xsb_eam_model {
  XSB_EAM_JOHNSON = 0,      /* 19 scalars (eam_potentials/johnson/johnson.h:29-50)                                            */
  XSB_EAM_SUTTON_CHEN = 1,  /* c, epsilon, a0, n, m (eam_potentials/sutton_chen/sutton_chen.h:24-62)                          */
  XSB_EAM_VNIITF = 2        /* rmax, rmin, rt0, Ecoh, E0, beta, A, Z, n, alpha, D, eta, mu (eam_potentials/vniitf/vniitf.h:31-125) */
} xsb_eam_model;
/* flags: XSB_FLAG_VIRIAL, XSB_FLAG_MIXED (FP32 rho(r) / phi(r), FP64 embedding function, distances and sums; 1e-5).      */
int xsb_eam_analytic_force(xsb_ctx* ctx, int model, const double* params, int nparams, double rcut, int phases, int flags);

/* a8  eam_alloy_force (eam_potential_multimat.cu:65-259, eam_alloy.h:37-313).                           */
typedef struct xsb_eam_alloy_tables {
  int32_t nelements, nr, nrho, pad_;
  double  rdr, rdrho, rc, rhomax;
  double  conversion_z2r, conversion_frho;   /* 1 eV.ang and 1 eV in internal units (eam_alloy.h:154-155) */
  const double* frho;   /* [nelements][nrho+1][8]  7-coefficient rows padded to 8 (eam_alloy.h:37-42)    */
  const double* rhor;   /* [nelements][nr+1][8]                                                          */
  const double* z2r;    /* [nelements(nelements+1)/2][nr+1][8]                                           */
} xsb_eam_alloy_tables;
/* reads a setfl file and builds the spline tables (eam_alloy.cpp:66-278); free with xsb_eam_alloy_free  */
int  xsb_eam_alloy_read(const char* path, xsb_eam_alloy_tables* out, char* names, size_t names_len);
void xsb_eam_alloy_free(xsb_eam_alloy_tables* t);
int  xsb_eam_alloy_set(xsb_ctx* ctx, const xsb_eam_alloy_tables* t);   /* uploads tables (host pointers)  */
int  xsb_eam_alloy_force(xsb_ctx* ctx, double rcut, int phases, int flags);
/* Inner skin of the rho phase (angstrom; 0 = off, the default unless XSB_INNER_SKIN is set).  The rho phase leaves the   */
/* in-range sub-list of the step for the force phase; with an inner skin that list keeps the pairs up to rcut + skin and   */
/* serves the rho phases of the following steps too (dense re-evaluation instead of re-filtering the whole neighbour list) */
/* for as long as no atom has moved further than skin / 2 -- accounted on the device by xsb_verlet_boundary[_async]; any   */
/* other way of moving particles (uploads, xsb_push_f_v_r, re-binning, a new xform) makes the next rho phase re-filter.   */
/* Results are those of the plain path (same pairs, same arithmetic per pair; summation order differs).                   */
int  xsb_eam_inner_skin(xsb_ctx* ctx, double skin);
/* Operator chains `compute_force: [eam_alloy_force, lj_multi_force]` (data/regression decks of configs[4]; onika runs the    */
/* operators of a batch one after the other, eam_potential_multimat.cu:196-230 then pair_potential_impl.hxx:476-484): the      */
/* force phase of xsb_eam_alloy_force is enqueued by the NEXT entry point called on the context.  When that entry is           */
/* xsb_pair_force / xsb_pair_multi_force with a Lennard-Jones potential, a cut-off <= the EAM one, no ghost flag and the same  */
/* energy / virial / mixed flags, both potentials are evaluated in ONE pass over the in-range pairs (same pairs, the pair      */
/* force is added to the EAM pair force before the multiplication with dr); in every other case the EAM phase is enqueued      */
/* first and the next operator runs as usual.  Every entry point (xsb_sync and xsb_field_device_ptr included) does this,       */
/* so the deferral cannot be observed.  XSB_NO_CHAIN_FUSION=1 (environment, read by xsb_create) turns it off.                  */
int  xsb_chain_stats(xsb_ctx* ctx, uint64_t* fused_pair_operators);
int  xsb_eam_sublist_stats(xsb_ctx* ctx, uint64_t* refiltered, uint64_t* reused);

/* ---------------------------------------------------------------------------------------------------- */
/* a9  snap_force (src/potential/snap/snap_force.cu:26-36 -> ext md::SnapForceGeneric; call sequence            */
/*     src/potential/snaplmp/snap_force_op.h:177-337; SNA constructor arguments snaplmp.cpp:205-216).           */
/*     LAMMPS conventions: ncoeff = number of bispectrum components for twojmax (55 for 8, 30 for 6),           */
/*     beta[nelements][ncoeff+1] with beta0 first, ALREADY in internal energy units (the reference scales the   */
/*     eV coefficients by EXASTAMP_CONST_QUANTITY(1 eV), snap_force_op.h:77); cutoff of a pair =                */
/*     (radelem[i]+radelem[j])*rcutfac; neighbour weight wjelem[j].                                             */
typedef struct xsb_snap_params {
  int32_t twojmax, switchflag, bzeroflag, nelements;
  int32_t quadraticflag, chemflag, switchinnerflag, pad_;   /* must be 0 (variants not implemented)            */
  double  rfac0, rmin0, rcutfac;
  const double* radelem;    /* [nelements] */
  const double* wjelem;     /* [nelements] */
  const double* beta;       /* [nelements][ncoeff+1] */
} xsb_snap_params;
int    xsb_snap_ncoeff(int twojmax);                        /* -1 if twojmax is outside 0..8                    */
int    xsb_snap_set(xsb_ctx* ctx, const xsb_snap_params* p);
double xsb_snap_rcut_max(xsb_ctx* ctx);                     /* rcut_max output slot: 2 max(radelem) rcutfac     */
/* f_i += fij, f_j -= fij (Newton-on like the reference, forces of ghost neighbours land on the ghost copies:   */
/* follow with xsb_ghost_reduce_add of fx,fy,fz = update_force_energy_from_ghost); flags: GHOST, ENERGY, VIRIAL, MIXED */
/* XSB_FLAG_MIXED: the reference's SNAP_FP32_MATH build (snap_force.cu:25-29): FP32 bispectrum arithmetic, FP64 positions,  */
/* forces and energies (tolerance 1e-5).                                                                                   */
/* Neighbours inside the SNAP cutoff are processed 64 at a time (the reference has no cap; BCC/FCC metals at the shipped */
/* rcutfac have 14-42); beyond 192 per atom (2J >= 7 pipeline) the call returns XSB_ERR_OVERFLOW (it syncs the stream   */
/* to find out).                                                                                                        */
int    xsb_snap_force(xsb_ctx* ctx, int flags);
int    xsb_snap_overflow(xsb_ctx* ctx, int* flag);          /* 1: a call since the last read exceeded the cap    */

/* ---------------------------------------------------------------------------------------------------- */
/* a10 ghost operators.  Single rank: ghosts are periodic images inside the same context.                */
typedef struct xsb_domain_desc {
  int32_t global_cells[3];   /* own (non-ghost) cells of the whole domain                               */
  int32_t periodic[3];
  int32_t rank_dims[3];      /* brick decomposition, product = number of ranks                          */
  int32_t rank_coord[3];     /* this rank's brick                                                        */
  double  box[3];            /* domain extent in grid space (global_cells * cell_size)                   */
} xsb_domain_desc;
/* NCCL communicator from a 128-byte ncclUniqueId created by rank 0 (xsb_comm_unique_id)                  */
int xsb_comm_unique_id(void* id128);
int xsb_comm_init(xsb_ctx* ctx, int nranks, int rank, const void* id128);
int xsb_comm_allreduce_max(xsb_ctx* ctx, double* inout_host);   /* MPI_Allreduce(MAX) of particle_displ_over */
/* ghost_comm_scheme + ghost_update_all_no_fv (config_move_particles.msp:82-87).  Precondition: the grid is   */
/* this rank's brick + ghost layers and the OWN cells hold their particles (ghost cells are ignored).        */
/* Effect: peers exchange per-cell counts, the SoA is re-laid out with every ghost cell sized for its        */
/* images, and r (shifted by the periodic box), v, type, id are copied owner -> ghost.  The particle count    */
/* and cell offsets change: re-read them with xsb_num_particles / xsb_cell_offsets_download.                  */
int xsb_ghost_comm_scheme(xsb_ctx* ctx, const xsb_domain_desc* dom);
/* the exchange plan behind xsb_ghost_comm_scheme as a pure host function (no context, no GPU): receive list of the   */
/* brick at rank_coord, 6 int32 per ghost cell { ghost_cell, owner_rank, owner_cell, wrap_x, wrap_y, wrap_z }, sorted  */
/* by owner rank; out6 may be NULL to query *count.  A rank's send list to peer q = the entries of q's plan it owns.   */
int xsb_ghost_plan(const xsb_domain_desc* dom, int ghost_layers, const int32_t* rank_coord, int32_t* out6, uint64_t capacity, uint64_t* count);
/* Transport of the exchanges below on more than one rank: by default every rank maps its peers' receive buffers (CUDA  */
/* IPC, one node) and the pack kernel of an exchange stores straight into them over NVLink, a release flag per source    */
/* tells the receiver's unpack kernel when a segment has landed (no NCCL call, no staging copy).  When the mapping is     */
/* not possible (or XSB_GHOST_NCCL is set) the exchange is a grouped ncclSend/ncclRecv.  out: "p2p", "nccl: <why>", "self" */
int xsb_ghost_transport(xsb_ctx* ctx, char* out, size_t len);
/* ghost_update_r / ghost_update_opt: owner -> ghost copy of the fields in field_mask (bit = xsb_field)   */
int xsb_ghost_update(xsb_ctx* ctx, uint32_t field_mask);
/* update_force_energy_from_ghost / update_virial_force_energy_from_ghost (src/mpi/update_from_ghosts.cu:29,43):  */
/* ghost -> owner add of the real-valued fields in field_mask (XSB_F_VIRIAL: all 9 components); <= 16 words/atom   */
int xsb_ghost_reduce_add(xsb_ctx* ctx, uint32_t field_mask);

/* move_particles + migrate_cell_particles (config_move_particles.msp:89-96,121-125): wrap into the periodic box,   */
/* send every own particle whose cell now belongs to another brick to that rank (NCCL P2P, any distance), re-bin     */
/* into cells; ghost cells are emptied (call xsb_ghost_comm_scheme next).  Collective over the communicator.         */
int xsb_particles_rebin(xsb_ctx* ctx, const xsb_domain_desc* dom);
int xsb_migration_stats(xsb_ctx* ctx, uint64_t* sent, uint64_t* received);   /* of the last xsb_particles_rebin */

/* ---------------------------------------------------------------------------------------------------- */
/* "next" rows: the per-particle operators either side of the force path               */
/* push_f_v_r: r += v dt + a dt^2/2 (a in fx,fy,fz after force_to_accel; grid-space r, INV_XFORM);       */
/* push_f_v: v += a dt  (config_numerical_schemes.msp:23-52 calls it with dt/2)                          */
int xsb_push_f_v_r(xsb_ctx* ctx, double dt);
int xsb_push_f_v(xsb_ctx* ctx, double dt);
/* force_to_accel (src/compute/force_to_accel.cu:82-101): f /= mass[type]                               */
int xsb_force_to_accel(xsb_ctx* ctx, int n_types, const double* mass);
/* backup_r + particle_displ_over (config_move_particles.msp:19-23): result = max |r - r_backup| > thr   */
int xsb_backup_r(xsb_ctx* ctx);
int xsb_particle_displ_over(xsb_ctx* ctx, double threshold, int* result, double* max_displ);
/* The operators on both sides of a step boundary of the velocity-Verlet scheme in one pass over the own atoms  */
/* (config_numerical_schemes.msp:23-52): force_to_accel, push_f_v(dt/2) | push_f_v_r(dt), push_f_v(dt/2),       */
/* particle_displ_over(threshold).  Same per-atom arithmetic as the five separate calls.                        */
int xsb_verlet_boundary(xsb_ctx* ctx, int n_types, const double* mass, double dt, double threshold, int* result, double* max_displ);

/* The same pass without the host read-back of xsb_verlet_boundary: max |r - r_backup| and the largest displacement of     */
/* this step are all-reduced over the ranks on the stream into a ring of 8 pinned slots.  xsb_displ_poll(lag) returns the   */
/* pair recorded `lag` calls earlier (0 = the call just made, which waits for it).  A driver that rebuilds when             */
/* max_displ(lag 1) + 2 * max_step_displ(lag 1) > threshold keeps every list valid without ever stalling on the GPU.        */
int xsb_verlet_boundary_async(xsb_ctx* ctx, int n_types, const double* mass, double dt);
int xsb_displ_poll(xsb_ctx* ctx, int lag, double* max_displ, double* max_step_displ);

/* simulation_thermodynamic_state (src/thermo_state/simulation_thermodynamic_state.cpp:81-230): sums over
 * the particles of own cells, all-reduced over ranks, in the reference's 27-double layout: virial[9] (zeros when no
 * operator produced it), ke_tensor[9] = 1/2 sum m v(x)v, momentum[3] = sum m v, kinetic_energy[3] = 1/2 sum m v_a^2,
 * potential_energy = sum ep, mass, particle count.  mass: n_types entries indexed by field::type.                  */
int xsb_thermo_state(xsb_ctx* ctx, int n_types, const double* mass, double* out27);

#ifdef __cplusplus
}
#endif
#endif /* XSB200_H */
