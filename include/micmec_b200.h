/*
 * micmec_b200.h - C ABI of libmicmec_b200.so, the B200 (sm_100a) implementation of MicMec's
 * force-evaluation + integration hot path.
 *
 * The reference (molmod/micmec) has NO FFI on this path: the boundary is the pure-Python plugin API
 *   ForcePart.compute(gpos=None, vtens=None) -> float          micmec/pes/mmff.py:87-149
 *   MicMecForceField.update_pos / update_rvecs                 micmec/pes/mmff.py:191-197
 *   ForcePartMechanical(system)                                micmec/pes/mmff.py:204-297
 *   VerletIntegrator.propagate + VerletHook init/pre/post      micmec/sampling/verlet.py:119-166
 *   NHChain.__call__ / MTKBarostat.baro / TBCombination        micmec/sampling/nvt.py:410-451, npt.py:99-148, 653-736
 *   Domain (the reference's only native code)                  micmec/pes/ext.pyx:36-123, micmec/pes/domain.c:13-71
 * Each entry point below names the reference interface it stands in for.  The Python mirror of the plugin API
 * (micmec_b200/pes/mmff.py, micmec_b200/sampling/*.py) binds these symbols with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only.  All floating point data are IEEE double.  Arrays are C-ordered
 * exactly like the reference's NumPy arrays (pos [nnodes][3], rvecs [3][3] rows a,b,c, vtens [3][3]).  Functions
 * return MM_OK (0) or a negative error code; mm_last_error() returns the message of the calling thread's last
 * failure.  A handle is bound to one CUDA device and one stream and is not re-entrant (like the reference's
 * ForcePart, which caches epot_cells/gpos_cells/verts_cells).  There is NO CPU fallback: every compute entry
 * point fails with MM_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MICMEC_B200_H
#define MICMEC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_OK 0
#define MM_ERR_INVALID (-1)   /* bad argument / unsupported topology (ValueError in the Python mirror) */
#define MM_ERR_CUDA (-2)      /* CUDA runtime failure, or no usable device */
#define MM_ERR_NAN (-3)       /* energy / gpos / vtens is NaN (mmff.py:135-147 raise ValueError) */
#define MM_ERR_STATE (-4)     /* call sequence error (e.g. MD not initialised) */

#define MM_MODEL_ORIGINAL 0   /* micmec/pes/nanocell_original.py - what mmff.py:27 imports */
#define MM_MODEL_DEFAULT 1    /* micmec/pes/nanocell.py - eight corner matrices */

#define MM_MAX_TYPES 8
#define MM_MAX_STATES 16      /* total over all types */
#define MM_MAX_CHAIN 8

/* flags for pointer arguments */
#define MM_HOST 0
#define MM_DEVICE 1

typedef struct mm_handle mm_handle;
typedef struct mm_md mm_md;

/* Static description of a system: what ForcePartMechanical.__init__ reads from `system`
 * (mmff.py:207-246) - topology, the minimum-image table restricted to the pairs that are looked up, and the
 * per-type parameters.  All pointers are HOST pointers and are copied. */
typedef struct {
    int64_t nnodes;
    int64_t ncells;
    const int64_t *surrounding_nodes; /* [ncells][8]  system.surrounding_nodes (micmec/system.py:60-63) */
    const int64_t *surrounding_cells; /* [nnodes][8]  system.surrounding_cells, -1 = missing (system.py:53-59) */
    const int8_t *shift;              /* [ncells][8][3] = mic[v0, vk, a] (mmff.py:259-286, used at :347-371) */
    const int32_t *cell_type;         /* [ncells] compact type index 0..ntypes-1 (system.types, mmff.py:373-376) */
    int32_t ntypes;
    const int32_t *type_nstates;      /* [ntypes] metastable states per type */
    const double *h0;                 /* [nstates_total][3][3]        typeN/cell        (mmff.py:229) */
    const double *elasticity;         /* [nstates_total][3][3][3][3]  typeN/elasticity  (mmff.py:231) */
    const double *free_energy;        /* [nstates_total]              typeN/free_energy (mmff.py:227) */
    const double *effective_temp;     /* [ntypes]                     typeN/effective_temp (mmff.py:225) */
    double boltzmann;                 /* molmod.boltzmann (mmff.py:33) */
    int32_t model;                    /* MM_MODEL_* */
    int32_t device;                   /* CUDA device ordinal */
    /* Structured fast path: non-zero nx,ny,nz declare a FULL periodic grid whose node / cell ids follow the
     * reference enumeration id = (k*ny + l)*nz + m (micmec/utils.py:113-137,150-161).  The index arrays above may
     * then be NULL (they are implied) - this is how the 64^3 / 256^3 grids are built without O(N^6) host work. */
    int32_t nx, ny, nz;
    /* z-slab decomposition over the GPUs of one box (one process per GPU): with slab_count > 1 this handle owns the
     * nz planes [slab_rank * nz, (slab_rank + 1) * nz) of a grid that is slab_count * nz planes tall; nnodes / ncells
     * are the LOCAL counts, nnodes_global the total.  Arrays passed to / returned by this handle are the local slab in
     * reference order (id = (k*ny + l)*nz + m_local).  Needs mm_comm_init before the first compute. */
    int32_t slab_rank, slab_count;
    int64_t nnodes_global;
    /* Batch of independent replicas (no communication, no reference counterpart): with nreplicas > 1 the arrays above
     * describe nreplicas systems of nnodes / nreplicas nodes and ncells / nreplicas cells each, concatenated (index
     * arrays already offset).  Every replica has its own domain vectors (mm_set_rvecs_batch); mm_compute returns the
     * summed energy / virial and mm_get_replica_results the per-replica values. */
    int64_t nreplicas;
} mm_desc;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int mm_create(const mm_desc *desc, mm_handle **out); /* ForcePartMechanical.__init__  mmff.py:207-246 */
int mm_destroy(mm_handle *h);
const char *mm_last_error(void);
int mm_version(void);
/* 1 when a CUDA device of compute capability 10.x is present and the library's kernels load on it */
int mm_device_ok(int device);

/* ---- geometry: MicMecForceField.update_pos / update_rvecs  (mmff.py:191-197) ----------------------------- */
int mm_set_pos(mm_handle *h, const double *pos, int where);     /* [nnodes][3], MM_HOST or MM_DEVICE */
int mm_set_rvecs(mm_handle *h, const double *rvecs9_host);      /* [3][3] rows a,b,c (zeros for nvec = 0) */

/* ---- ForcePart.compute(gpos, vtens)  (mmff.py:87-149, 288-323) ------------------------------------------- */
/* energy_host receives the energy.  gpos / vtens9 may be NULL (any combination, cf. stress_strain.py:85).
 * Results OVERWRITE the output arrays; the "add into the caller's array" step of mmff.py:142,148 is done by the
 * Python mirror.  `where` tells whether gpos is a host or a device pointer; vtens9 and energy are host. */
int mm_compute(mm_handle *h, double *energy_host, double *gpos, int where, double *vtens9_host);
/* per-cell caches of the last compute: ForcePartMechanical.epot_cells / gpos_cells (mmff.py:290-292) */
int mm_get_cell_cache(mm_handle *h, double *epot_cells_host, double *gpos_cells_host /* [ncells][8][3] */);
/* number of kernels this handle (and MD objects on it) launched so far */
int64_t mm_launch_count(const mm_handle *h);
/* raw device pointers for zero-copy interop with torch tensors: 0 pos, 1 gpos */
void *mm_device_ptr(mm_handle *h, int which);
/* use the given cudaStream_t for all subsequent work of this handle (default: a private stream) */
int mm_set_stream(mm_handle *h, void *cuda_stream);
int mm_synchronize(mm_handle *h);
/* options: "scatter" 0 = node-centric gather (deterministic) / 1 = cell-centric warp-aggregated atomic scatter for
 * the gpos accumulation of the structured path; "profile" 1 = bracket every force kernel with CUDA events */
int mm_set_option(mm_handle *h, const char *name, int64_t value);
/* read back a setting or a derived launch parameter: "structured" (1 = the structured-grid kernels are active),
 * "chunk" (planes per block along z; the shortest class with the two-class schedule), "rows_per_thread", "tile_rows",
 * "blocks", "mass_uniform", "march2", "tail", "wrap_on_load", "plan_efficiency_permille"; -1 for unknown names */
int64_t mm_get_option(const mm_handle *h, const char *name);
/* The block schedule of the marching kernel for ntx x nty tiles of `planes` owned planes on nsm SMs, without a device (no
 * counterpart in the reference: introspection for tests and tuning).  items: [capacity][4] = tile x, tile y, first owned
 * plane (from 1), one past the last, in dispatch order (may be NULL to query *nitems); cost / ideal: simulated makespan and
 * perfect balance in plane iterations.  uniform_chunk > 0: equal chunks of that length instead of the two-class plan. */
int mm_plan_schedule(int ntx, int nty, int planes, int nsm, int images_on_load, int uniform_chunk, int32_t *items, int64_t capacity,
                     int64_t *nitems, double *cost, double *ideal);
/* with "profile" on: launches timed since the last call and their summed device time (ms), separately for
 * [0] force-only kernels and [1] fused kick-drift-force-kick kernels; synchronises the stream and resets the counters */
int mm_profile(mm_handle *h, int64_t nlaunch[2], double total_ms[2]);

/* ---- replica batches ---------------------------------------------------------------------------------------- */
int mm_set_rvecs_batch(mm_handle *h, const double *rvecs_host /* [nreplicas][3][3] */);
int mm_get_replica_results(mm_handle *h, double *energies_host /* [nreplicas] */, double *vtens_host /* [nreplicas][3][3] or NULL */);

/* ---- batched dense symmetric eigen-decomposition --------------------------------------------------------------- */
/* numpy.linalg.eigh as QNOptimizer uses it on its Hessian model (micmec/sampling/opt.py:196-197 get_spectrum, called at
 * :334-336), for `batch` independent n x n matrices at once (1 <= n <= 96; one thread block per matrix, two-sided cyclic
 * Jacobi in shared memory).  mats [batch][n][n] row-major (symmetrised on load); evals [batch][n] ascending; evecs
 * [batch][n][n] with eigenvector i in COLUMN i, as numpy returns them.  `where` applies to all three arrays.
 * max_sweeps_out (may be NULL) receives the largest number of Jacobi sweeps any matrix needed (30 = not converged). */
int mm_batched_eigh(int device, int64_t batch, int32_t n, const double *mats, int where, double *evals, double *evecs,
                    int32_t *max_sweeps_out);

/* ---- device-resident lockstep quasi-Newton optimiser of a replica batch (config 5) ------------------------------------ */
/* QNOptimizer over a CartesianDOF (kind 0) or a StrainCellDOF (kind 1) - micmec/sampling/opt.py:258-443, dof.py:75-193,
 * 522-697 - one instance per replica of the batch handle `h` (created with nreplicas > 1; per-replica domain vectors set with
 * mm_set_rvecs_batch), all replicas advancing in lockstep: SR1 Hessian models, their spectra (Jacobi), the ridge search of
 * solve_trust_radius, accept / shrink, the DOF mappings and the convergence criteria run on the device; a sweep costs one
 * batched force evaluation.  x0 [R][ndof] (host): positions (kind 0) or [6 strain variables, fractional coordinates]
 * (kind 1).  kind 1 also takes rvecs0 [R][9], jac [R][9][6] = d rvecs / d strain and proj [R][9][9], the projector on the
 * range of jac (dof.py:582-697).  Thresholds and radii: QNOptimizer(trust_radius, small_radius, too_small_radius),
 * DOF(gpos_rms, dpos_rms, grvecs_rms, drvecs_rms). */
typedef struct mm_qn mm_qn;
int mm_qn_create(mm_handle *h, int kind, const double *x0_host, const double *rvecs0_host, const double *jac_host,
                 const double *proj_host, double gpos_rms, double dpos_rms, double grvecs_rms, double drvecs_rms,
                 double trust_radius, double small_radius, double too_small_radius, mm_qn **out);
int mm_qn_destroy(mm_qn *q);
/* nsweeps sweeps; *nlive_out = replicas that have neither converged nor failed (trust radius underflow) yet */
int mm_qn_sweep(mm_qn *q, int nsweeps, int *nlive_out);
/* state read-back (host arrays, any may be NULL): x, g [R][ndof]; f, radius, conv_val [R]; int32 [R]: accepted steps,
 * converged, failed, number of unmet criteria; *evaluations = batched force calls so far */
int mm_qn_get(mm_qn *q, double *x, double *f, double *g, double *radius, double *conv_val, int32_t *iterations,
              int32_t *converged, int32_t *failed, int32_t *conv_count, int64_t *evaluations);

/* ---- multi-GPU (no reference counterpart: the reference is single-process) ---------------------------------- */
/* NCCL bootstrap: rank 0 creates a 128-byte unique id, the host side broadcasts it (torch.distributed), every rank
 * calls mm_comm_init.  nccl_path: the libnccl.so.2 to dlopen (NULL: the one already loaded in the process).
 * Afterwards halo planes travel with ncclSend/ncclRecv and the <= 16 reduced doubles with ncclAllReduce, all on the
 * handle's stream, inside mm_compute / mm_md_run. */
int mm_comm_unique_id(const char *nccl_path, char *out128);
int mm_comm_init(mm_handle *h, const char *nccl_path, const char *id128);
int mm_comm_destroy(mm_handle *h);
/* exchange mode the ranks agreed on: -1 no communicator, 0 NCCL send/recv + all-reduce, 1 peer inboxes over NVLink
 * (CUDA IPC), 2 fused halo (the marching kernel stores its boundary planes into the neighbours' halo planes) */
int mm_comm_mode(const mm_handle *h);

/* ---- Domain  (micmec/pes/ext.pyx:36-123 + micmec/pes/domain.c:13-71) ------------------------------------ */
/* rvecs [nvec][3]; writes volume (domain.c:23-48) and the reciprocal vectors gvecs [nvec][3] (ext.pyx:64-71) */
int mm_domain(const double *rvecs, int nvec, double *volume, double *gvecs);

/* ---- device-resident MD: VerletIntegrator + NHCThermostat + MTKBarostat + TBCombination ------------------ */
typedef struct {
    double timestep;          /* verlet.py:71-83 */
    double ndof;              /* <= 0: decide like the reference (verlet.py:131-132, sampling/utils.py:322-343) */
    /* Nose-Hoover chain (nvt.py:361-408).  has_thermo = 0 disables it. */
    int32_t has_thermo;
    int32_t chain_length;
    double thermo_temp, thermo_timecon;
    /* MTK barostat (npt.py:513-614).  has_baro = 0 disables it. */
    int32_t has_baro;
    int32_t anisotropic, vol_constraint;
    double baro_temp, baro_press, baro_timecon;
    /* Langevin thermostat on the device (nvt.py:165-218; instead of the Nose-Hoover chain, without a barostat).  The noise
     * comes from a counter-based generator (Philox4x32-10 keyed by seed, node id and half-step number): the same node gets
     * the same kick on every GPU and in every kernel layout, but the stream is not NumPy's - parity is statistical. */
    int32_t has_langevin;
    int32_t thermo_kind;  /* with has_thermo: 0 Nose-Hoover chain, 1 Berendsen weak coupling (nvt.py:110-162), 2 stochastic velocity
                           * rescaling (CSVR, nvt.py:221-274; noise from langevin_seed) - both use thermo_temp, thermo_timecon */
    double langevin_temp, langevin_timecon;
    uint64_t langevin_seed;
    double time0;      /* VerletIntegrator(time0=...), verlet.py:96: simulation time at initialisation (restarts) */
    int64_t counter0;  /* VerletIntegrator(counter0=...), iterative.py: step counter at initialisation */
} mm_md_desc;

int mm_md_create(mm_handle *h, const mm_md_desc *desc, mm_md **out);
int mm_md_destroy(mm_md *md);
/* Upload the dynamic state.  masses [nnodes]; chain_pos/chain_vel [chain_length] (NULL -> zeros / keep);
 * vel_press [3][3] (isotropic: [0]).  `where` applies to pos/vel/masses.  Performs the reference's initialisation
 * (verlet.py:119-137, nvt.py:507-523 without the RNG, npt.py:579-614): first force evaluation, chain and barostat
 * masses, ndof, properties at counter0. */
int mm_md_init(mm_md *md, const double *pos, const double *vel, const double *masses, int where,
               const double *rvecs9, const double *chain_pos, const double *chain_vel, const double *vel_press9);
/* overwrite positions and velocities of a running integrator (same geometry bookkeeping as a hook that assigns
 * iterative.pos / iterative.vel); gradients on the device are NOT recomputed */
int mm_md_set_state(mm_md *md, const double *pos, const double *vel, int where);
/* nsteps x VerletIntegrator.propagate (verlet.py:140-166) entirely on the device, no host round trip */
int mm_md_run(mm_md *md, int64_t nsteps);
/* state read-back (host or device destination for the fields, host for the rest); any pointer may be NULL */
int mm_md_get_state(mm_md *md, double *pos, double *vel, double *gpos, int where, double *rvecs9,
                    double *chain_pos, double *chain_vel, double *vel_press9);
/* scalars of verlet.py:171-190 at the current step */
#define MM_S_EPOT 0
#define MM_S_EKIN 1
#define MM_S_TEMP 2
#define MM_S_ETOT 3
#define MM_S_ECONS 4
#define MM_S_CONS_ERR 5
#define MM_S_PRESS 6
#define MM_S_RMSD_GPOS 7
#define MM_S_RMSD_DELTA 8
#define MM_S_TIME 9
#define MM_S_COUNTER 10
#define MM_S_VOLUME 11
#define MM_S_NDOF 12
#define MM_S_ECONS_CORR 13
#define MM_S_VTENS 16  /* 9 values */
#define MM_S_PTENS 25  /* 9 values */
#define MM_S_NFORCE 34 /* force evaluations so far */
#define MM_S_CE_N 35   /* ConsErrTracker (verlet.py:275-307): counter, ekin_m, ekin_s, econs_m, econs_s */
#define MM_S_COUNT 40
int mm_md_scalars(mm_md *md, double *out /* [MM_S_COUNT] */);

#ifdef __cplusplus
}
#endif
#endif /* MICMEC_B200_H */
