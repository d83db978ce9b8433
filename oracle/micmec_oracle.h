/*
 * micmec_oracle.h - CPU ORACLE for the MicMec force + MD hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference algorithm, used as the
 * checker in tests/, in __graft_entry__.smoke() and as bench.py's `cpu_baseline` / `--impl reference`
 * leg.  The product (micmec_b200/) never includes, links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against golden
 * vectors recorded from the unmodified reference (tests/golden/make_golden.py) to <= 1e-12.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 */
#ifndef MICMEC_ORACLE_H
#define MICMEC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODEL_ORIGINAL = 0, ORC_MODEL_DEFAULT = 1 };

/* Static topology + parameters of a system ("construct a new MMFF if anything else changes",
 * micmec/pes/mmff.py:92-93).  All pointers are borrowed. */
typedef struct {
    int64_t nnodes;
    int64_t ncells;
    const int64_t *surrounding_nodes; /* [ncells][8]   micmec/system.py:60-63 */
    const int64_t *surrounding_cells; /* [nnodes][8], -1 = missing   micmec/system.py:53-59 */
    const int8_t *shift;              /* [ncells][8][3] = mic[v0, vk, a]   micmec/pes/mmff.py:259-286 */
    const int32_t *cell_type;         /* [ncells] compact type index */
    int32_t ntypes;
    const int32_t *type_nstates;      /* [ntypes] */
    const int32_t *type_offset;       /* [ntypes] first state of the type in the arrays below */
    const double *h0;                 /* [nstates_total][3][3]   typeN/cell */
    const double *C;                  /* [nstates_total][3][3][3][3]   typeN/elasticity */
    const double *efree;              /* [nstates_total]   typeN/free_energy */
    const double *temp_eff;           /* [ntypes]   typeN/effective_temp */
    double boltzmann;
    int32_t model;                    /* ORC_MODEL_* */
    int32_t nthreads;                 /* OpenMP threads for the cell / node loops (>=1) */
} orc_system;

/* One metastable state of one cell: energy (without free energy) and 8x3 gradient.
 * original: micmec/pes/nanocell_original.py:46-84, 87-132;  default: micmec/pes/nanocell.py:40-77, 80-132 */
void orc_cell_state(int model, const double verts[24], const double h0[9], const double C[81],
                    double *energy, double g[24]);

/* micmec/pes/mmff.py:326-403 (deformation): per-cell energy, gradient and unwrapped vertices. */
void orc_deformation(const orc_system *sys, const double *pos, const double *rvecs,
                     double *epot_cells, double *gpos_cells, double *verts_cells);

/* micmec/pes/mmff.py:288-323: energy (returned), optional gpos [nnodes][3] and vtens [3][3] (overwritten). */
double orc_compute(const orc_system *sys, const double *pos, const double *rvecs,
                   double *gpos_or_null, double *vtens_or_null, double *work);

/* size (in doubles) of the `work` scratch orc_compute needs */
int64_t orc_work_size(const orc_system *sys);

/* micmec/pes/domain.c:42-48 */
double orc_volume(const double *rvecs);

/* ------------------------------------------------------------------ MD ---------------------------------- */

#define ORC_MAX_CHAIN 16

/* micmec/sampling/nvt.py:361-458 (NHChain) */
typedef struct {
    int32_t length;
    double timestep, temp, timecon, ndof;
    double pos[ORC_MAX_CHAIN], vel[ORC_MAX_CHAIN], masses[ORC_MAX_CHAIN];
} orc_chain;

/* micmec/sampling/npt.py:513-757 (MTKBarostat), without baro_thermo */
typedef struct {
    double temp, press, timecon, timestep, mass_press;
    int32_t anisotropic, vol_constraint, dim, baro_ndof;
    double vel_press[9]; /* isotropic: only [0] is used */
} orc_baro;

typedef struct {
    int64_t nnodes;
    double *pos, *vel, *gpos, *masses; /* [nnodes][3] / [nnodes] */
    double *posold, *delta;            /* [nnodes][3] */
    double rvecs[9], vtens[9], ptens[9];
    double timestep, time;
    double ndof;
    int64_t counter;
    double epot, ekin, temp, etot, econs, cons_err, press, rmsd_gpos, rmsd_delta;
    /* ConsErrTracker, micmec/sampling/verlet.py:275-307 */
    int64_t ce_counter;
    double ce_ekin_m, ce_ekin_s, ce_econs_m, ce_econs_s;
    int32_t has_thermo, has_baro;
    orc_chain chain;
    orc_baro baro;
    double econs_correction;
    double *work;
    int64_t nforce; /* number of force evaluations so far */
} orc_md;

void orc_chain_set_ndof(orc_chain *chain, double ndof, double boltzmann); /* nvt.py:393-400 (masses only) */
void orc_chain_call(orc_chain *chain, double boltzmann, double *ekin, double *vel, int64_t nvel,
                    int has_g1, double g1_add); /* nvt.py:410-451 */
double orc_chain_econs(const orc_chain *chain, double boltzmann); /* nvt.py:453-458 */

/* verlet.py:119-137 + hook init (nvt.py:507-523, npt.py:579-614); velocities / chain.vel / vel_press are
 * taken from the struct as given (RNG stays with the caller).  Momenta are NOT cleaned here. */
void orc_md_initialize(const orc_system *sys, orc_md *md);
/* verlet.py:140-166 with TBCombination.pre/post (npt.py:99-148) */
void orc_md_step(const orc_system *sys, orc_md *md);
void orc_md_properties(const orc_system *sys, orc_md *md); /* verlet.py:171-190 */

/* symmetric 3x3: out = Q diag(f(w)) Q^T with f = exp(scale*w); used for npt.py:686-690, 710-721 */
void orc_sym_expm(const double a[9], double scale, double out[9]);

#ifdef __cplusplus
}
#endif
#endif
