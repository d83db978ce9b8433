/*
 * micmec_oracle.c - CPU ORACLE (test infrastructure only, see micmec_oracle.h).
 *
 * A deliberately literal restatement of the reference's NumPy code: the einsum index strings of the
 * reference are kept as explicit loops over the same stencil tables (multiplicator, cell_{x,y,z}derivs),
 * NOT the closed forms the CUDA kernels use - so that CUDA-vs-oracle parity is a real check of those
 * closed forms.  Citations are relative to /root/reference.
 */
#include "micmec_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* micmec/utils.py:32-41 */
static const int NEIGHBOR_CELLS[8][3] = {
    {0, 0, 0}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-1, -1, 0}, {-1, 0, -1}, {0, -1, -1}, {-1, -1, -1}};
/* micmec/utils.py:43-52 */
static const int NEIGHBOR_NODES[8][3] = {
    {0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};

static int tables_ready = 0;
static double MULT[8][3][8];         /* micmec/pes/nanocell_utils.py:32-75 */
static double DERIV[3][8][8][3][3];  /* cell_{x,y,z}derivs[v][a][k][m], nanocell_utils.py:78-110 */

static int node_index(int dx, int dy, int dz) {
    for (int v = 0; v < 8; v++)
        if (NEIGHBOR_NODES[v][0] == dx && NEIGHBOR_NODES[v][1] == dy && NEIGHBOR_NODES[v][2] == dz) return v;
    return -1;
}

static void build_tables(void) {
    if (tables_ready) return;
    /* multiplicator[a][i][:]: edge i of representation a runs from vertex a along +axis i (if a sits at
     * offset 0 on that axis) or arrives at vertex a (if it sits at offset 1); always oriented +axis.
     * This reproduces the literal table nanocell_utils.py:32-75 (checked against it in the golden test). */
    memset(MULT, 0, sizeof(MULT));
    for (int a = 0; a < 8; a++) {
        for (int i = 0; i < 3; i++) {
            int d[3] = {NEIGHBOR_NODES[a][0], NEIGHBOR_NODES[a][1], NEIGHBOR_NODES[a][2]};
            int lo[3] = {d[0], d[1], d[2]}, hi[3] = {d[0], d[1], d[2]};
            lo[i] = 0;
            hi[i] = 1;
            MULT[a][i][node_index(lo[0], lo[1], lo[2])] = -1.0;
            MULT[a][i][node_index(hi[0], hi[1], hi[2])] = 1.0;
        }
    }
    /* nanocell_utils.py:78-110, including the final transposes (:104-106) */
    memset(DERIV, 0, sizeof(DERIV));
    for (int v = 0; v < 8; v++) {
        for (int a = 0; a < 8; a++) {
            double deriv[3];
            int dist_vec[3], dist = 0;
            for (int n = 0; n < 3; n++) {
                deriv[n] = (NEIGHBOR_CELLS[v][n] == -1) ? 1.0 : -1.0;
                dist_vec[n] = abs(NEIGHBOR_CELLS[v][n] - NEIGHBOR_CELLS[a][n]);
                dist += dist_vec[n];
            }
            for (int c = 0; c < 3; c++) { /* c = 0,1,2 -> xderiv, yderiv, zderiv */
                double m[3][3];
                memset(m, 0, sizeof(m));
                if (dist == 0) {
                    for (int n = 0; n < 3; n++) m[n][c] = deriv[n];
                } else if (dist == 1) {
                    for (int n = 0; n < 3; n++)
                        if (dist_vec[n] == 1) m[n][c] = deriv[n];
                }
                for (int k = 0; k < 3; k++)
                    for (int mm = 0; mm < 3; mm++) DERIV[c][v][a][k][mm] = m[mm][k]; /* .T */
            }
        }
    }
    tables_ready = 1;
}

/* exported for the golden test: lets Python compare the tables with the reference's arrays */
void orc_tables(double *mult_out, double *deriv_out) {
    build_tables();
    memcpy(mult_out, MULT, sizeof(MULT));
    memcpy(deriv_out, DERIV, sizeof(DERIV));
}

static double det3(const double *a) {
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
           a[2] * (a[3] * a[7] - a[4] * a[6]);
}

static void inv3(const double *a, double *inv) {
    double d = det3(a);
    inv[0] = (a[4] * a[8] - a[5] * a[7]) / d;
    inv[1] = (a[2] * a[7] - a[1] * a[8]) / d;
    inv[2] = (a[1] * a[5] - a[2] * a[4]) / d;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) / d;
    inv[4] = (a[0] * a[8] - a[2] * a[6]) / d;
    inv[5] = (a[2] * a[3] - a[0] * a[5]) / d;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) / d;
    inv[7] = (a[1] * a[6] - a[0] * a[7]) / d;
    inv[8] = (a[0] * a[4] - a[1] * a[3]) / d;
}

/* domain.c:42-48 */
double orc_volume(const double *r) {
    return fabs(r[0] * (r[4] * r[8] - r[5] * r[7]) + r[1] * (r[5] * r[6] - r[3] * r[8]) +
                r[2] * (r[3] * r[7] - r[4] * r[6]));
}

#define C4(C, i, j, k, l) ((C)[(((i) * 3 + (j)) * 3 + (k)) * 3 + (l)])

/* strain = 0.5 (M_^T M_ - I), with M_[i][k] = sum_j H[j][i] h0inv[k][j]  ("ji,kj->ik")
 * nanocell_original.py:76-79 / nanocell.py:68-70 */
static void strain_of(const double H[3][3], const double *h0inv, double M_[3][3], double eps[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            double s = 0.0;
            for (int j = 0; j < 3; j++) s += H[j][i] * h0inv[k * 3 + j];
            M_[i][k] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            double s = 0.0;
            for (int j = 0; j < 3; j++) s += M_[j][i] * M_[j][k]; /* "ji,jk->ik" */
            eps[i][k] = 0.5 * (s - (i == k ? 1.0 : 0.0));
        }
}

static void stress_of(const double *C, const double eps[3][3], double sig[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) s += C4(C, i, j, k, l) * eps[k][l]; /* "ijkl,kl->ij" */
            sig[i][j] = s;
        }
}

/* g_c = sum_ij sym(M_^T . dH . h0inv^T)[j][i] * stress[i][j] for the derivative stencil dH = scale*DERIV[c][v][a]
 * nanocell_original.py:114-130 / nanocell.py:108-130 */
static double grad_component(const double M_[3][3], const double dH[3][3], double scale, const double *h0inv,
                             const double sig[3][3]) {
    double mat[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++)
                for (int m = 0; m < 3; m++) s += M_[k][i] * (scale * dH[k][m]) * h0inv[j * 3 + m]; /* "ki,km,jm->ij" */
            mat[i][j] = s;
        }
    double g = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) g += 0.5 * (mat[i][j] + mat[j][i]) * sig[i][j]; /* "...ji,ij" on the sym. part */
    return g;
}

void orc_cell_state(int model, const double verts[24], const double h0[9], const double C[81], double *energy,
                    double g[24]) {
    build_tables();
    double h0inv[9];
    inv3(h0, h0inv);
    const double h0det = det3(h0);
    /* matrices[a][i][j] = sum_v multiplicator[a][i][v] verts[v][j]   (nanocell.py:63-65) */
    double mats[8][3][3];
    for (int a = 0; a < 8; a++)
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0.0;
                for (int v = 0; v < 8; v++) s += MULT[a][i][v] * verts[v * 3 + j];
                mats[a][i][j] = s;
            }
    if (model == ORC_MODEL_ORIGINAL) {
        double H[3][3], M_[3][3], eps[3][3], sig[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0.0;
                for (int a = 0; a < 8; a++) s += mats[a][i][j];
                H[i][j] = 0.125 * s; /* nanocell_original.py:75 */
            }
        strain_of(H, h0inv, M_, eps);
        stress_of(C, eps, sig);
        double dens = 0.0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) dens += eps[i][j] * sig[i][j];
        *energy = 0.5 * dens * h0det; /* :81-82 */
        for (int v = 0; v < 8; v++)
            for (int c = 0; c < 3; c++) /* cell_xderiv[v] = 0.25*cell_xderivs[v][v]  (:41-43) */
                g[v * 3 + c] = h0det * grad_component(M_, DERIV[c][v][v], 0.25, h0inv, sig);
    } else {
        double M_[8][3][3], eps[8][3][3], sig[8][3][3];
        double dens = 0.0;
        for (int a = 0; a < 8; a++) {
            strain_of(mats[a], h0inv, M_[a], eps[a]);
            stress_of(C, eps[a], sig[a]);
            double d = 0.0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) d += eps[a][i][j] * sig[a][i][j];
            dens += 0.5 * d;
        }
        *energy = 0.125 * dens * h0det; /* nanocell.py:74-75 */
        for (int v = 0; v < 8; v++)
            for (int c = 0; c < 3; c++) {
                double s = 0.0;
                for (int a = 0; a < 8; a++) s += grad_component(M_[a], DERIV[c][v][a], 1.0, h0inv, sig[a]);
                g[v * 3 + c] = h0det * 0.125 * s; /* nanocell.py:124-132 */
            }
    }
}

/* mmff.py:326-403 */
void orc_deformation(const orc_system *sys, const double *pos, const double *rvecs, double *epot_cells,
                     double *gpos_cells, double *verts_cells) {
    build_tables();
    const int64_t ncells = sys->ncells;
#pragma omp parallel for schedule(static) num_threads(sys->nthreads > 0 ? sys->nthreads : 1)
    for (int64_t c = 0; c < ncells; c++) {
        const int64_t *vidx = sys->surrounding_nodes + 8 * c;
        double *verts = verts_cells + 24 * c;
        const double *r0 = pos + 3 * vidx[0];
        for (int k = 0; k < 8; k++) { /* mmff.py:347-371: r_k = r0 + (pos[k] - pos[0] + sum_a rvecs[a] mic[0,k,a]) */
            const double *rk = pos + 3 * vidx[k];
            const int8_t *s = sys->shift + (c * 8 + k) * 3;
            for (int d = 0; d < 3; d++) {
                if (k == 0) {
                    verts[d] = r0[d];
                } else {
                    double dvec = rk[d] - r0[d];
                    dvec += rvecs[0 + d] * s[0] + rvecs[3 + d] * s[1] + rvecs[6 + d] * s[2];
                    verts[k * 3 + d] = r0[d] + dvec;
                }
            }
        }
        const int t = sys->cell_type[c];
        const int ns = sys->type_nstates[t];
        const int off = sys->type_offset[t];
        const double kT = sys->boltzmann * sys->temp_eff[t];
        double e_states[ORC_MAX_CHAIN], g_states[ORC_MAX_CHAIN][24];
        double emin = INFINITY;
        for (int s = 0; s < ns; s++) { /* mmff.py:377-385 */
            double e;
            orc_cell_state(sys->model, verts, sys->h0 + 9 * (off + s), sys->C + 81 * (off + s), &e, g_states[s]);
            e_states[s] = e + sys->efree[off + s];
            if (e_states[s] < emin) emin = e_states[s];
        }
        double w[ORC_MAX_CHAIN], wsum = 0.0; /* mmff.py:386-398 */
        for (int s = 0; s < ns; s++) {
            w[s] = exp(-(e_states[s] - emin) / kT);
            wsum += w[s];
        }
        epot_cells[c] = emin - sys->temp_eff[t] * sys->boltzmann * log(wsum);
        double *g = gpos_cells + 24 * c;
        for (int i = 0; i < 24; i++) g[i] = 0.0;
        for (int s = 0; s < ns; s++) {
            const double wn = w[s] / wsum;
            for (int i = 0; i < 24; i++) g[i] += wn * g_states[s][i];
        }
    }
}

int64_t orc_work_size(const orc_system *sys) { return sys->ncells * (1 + 24 + 24); }

/* mmff.py:288-323 */
double orc_compute(const orc_system *sys, const double *pos, const double *rvecs, double *gpos, double *vtens,
                   double *work) {
    double *epot_cells = work;
    double *gpos_cells = work + sys->ncells;
    double *verts_cells = gpos_cells + 24 * sys->ncells;
    orc_deformation(sys, pos, rvecs, epot_cells, gpos_cells, verts_cells);
    if (gpos) { /* mmff.py:303-318, fixed order j = 0..7 */
#pragma omp parallel for schedule(static) num_threads(sys->nthreads > 0 ? sys->nthreads : 1)
        for (int64_t n = 0; n < sys->nnodes; n++) {
            double acc[3] = {0.0, 0.0, 0.0};
            for (int j = 0; j < 8; j++) {
                const int64_t c = sys->surrounding_cells[8 * n + j];
                if (c < 0) continue;
                for (int d = 0; d < 3; d++) acc[d] += gpos_cells[(c * 8 + j) * 3 + d];
            }
            for (int d = 0; d < 3; d++) gpos[3 * n + d] = acc[d];
        }
    }
    if (vtens) { /* mmff.py:320-323: einsum("ijk,ijl->kl", gpos_cells, verts_cells) */
        for (int i = 0; i < 9; i++) vtens[i] = 0.0;
        for (int64_t c = 0; c < sys->ncells; c++)
            for (int v = 0; v < 8; v++)
                for (int k = 0; k < 3; k++)
                    for (int l = 0; l < 3; l++)
                        vtens[k * 3 + l] += gpos_cells[(c * 8 + v) * 3 + k] * verts_cells[(c * 8 + v) * 3 + l];
    }
    double e = 0.0; /* mmff.py:299-301 */
    for (int64_t c = 0; c < sys->ncells; c++) e += epot_cells[c];
    return e;
}

/* ---------------------------------------------------------------------------------------------- MD -------- */

/* nvt.py:393-400 (the random velocities of :399-408 stay with the caller) */
void orc_chain_set_ndof(orc_chain *ch, double ndof, double boltzmann) {
    ch->ndof = ndof;
    const double afreq = 2.0 * M_PI / ch->timecon;
    for (int k = 0; k < ch->length; k++) ch->masses[k] = boltzmann * ch->temp / (afreq * afreq);
    ch->masses[0] *= ndof;
}

static void chain_bead(orc_chain *ch, double kb, int k, double ekin, int has_g1, double g1) {
    double g;
    if (k == 0) { /* nvt.py:413-420 */
        g = 2.0 * ekin - ch->ndof * ch->temp * kb;
        if (has_g1) g += g1;
    } else {
        g = ch->masses[k - 1] * ch->vel[k - 1] * ch->vel[k - 1] - ch->temp * kb;
    }
    g /= ch->masses[k];
    if (k == ch->length - 1) {
        ch->vel[k] += g * ch->timestep / 4.0;
    } else {
        ch->vel[k] *= exp(-ch->vel[k + 1] * ch->timestep / 8.0);
        ch->vel[k] += g * ch->timestep / 4.0;
        ch->vel[k] *= exp(-ch->vel[k + 1] * ch->timestep / 8.0);
    }
}

/* nvt.py:410-451 */
void orc_chain_call(orc_chain *ch, double kb, double *ekin, double *vel, int64_t nvel, int has_g1, double g1) {
    for (int k = ch->length - 1; k >= 0; k--) chain_bead(ch, kb, k, *ekin, has_g1, g1);
    for (int k = 0; k < ch->length; k++) ch->pos[k] += ch->vel[k] * ch->timestep / 2.0;
    const double factor = exp(-ch->vel[0] * ch->timestep / 2.0);
    for (int64_t i = 0; i < nvel; i++) vel[i] *= factor;
    *ekin *= factor * factor;
    for (int k = 0; k < ch->length; k++) chain_bead(ch, kb, k, *ekin, has_g1, g1);
}

/* nvt.py:453-458 */
double orc_chain_econs(const orc_chain *ch, double kb) {
    const double kt = kb * ch->temp;
    double s = 0.0, p = 0.0;
    for (int k = 0; k < ch->length; k++) s += ch->vel[k] * ch->vel[k] * ch->masses[k];
    for (int k = 1; k < ch->length; k++) p += ch->pos[k];
    return 0.5 * s + kt * (ch->ndof * ch->pos[0] + p);
}

/* cyclic Jacobi for a symmetric 3x3 (stands in for numpy.linalg.eigh at npt.py:686, 712-716; eigenvector
 * sign/order conventions cancel in Q f(D) Q^T) */
static void jacobi3(const double a_in[9], double w[3], double Q[9]) {
    double a[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = (i >= j) ? a_in[i * 3 + j] : a_in[j * 3 + i]; /* lower triangle, as LAPACK */
    double q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-40 * (diag + 1e-300)) break;
        for (int p = 0; p < 2; p++)
            for (int r = p + 1; r < 3; r++) {
                if (a[p][r] == 0.0) continue;
                double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = a[k][p], akr = a[k][r];
                    a[k][p] = c * akp - s * akr;
                    a[k][r] = s * akp + c * akr;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = a[p][k], ark = a[r][k];
                    a[p][k] = c * apk - s * ark;
                    a[r][k] = s * apk + c * ark;
                }
                for (int k = 0; k < 3; k++) {
                    double qkp = q[k][p], qkr = q[k][r];
                    q[k][p] = c * qkp - s * qkr;
                    q[k][r] = s * qkp + c * qkr;
                }
            }
    }
    for (int i = 0; i < 3; i++) {
        w[i] = a[i][i];
        for (int j = 0; j < 3; j++) Q[i * 3 + j] = q[i][j];
    }
}

void orc_sym_expm(const double a[9], double scale, double out[9]) {
    double w[3], Q[9];
    jacobi3(a, w, Q);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += Q[i * 3 + k] * exp(scale * w[k]) * Q[j * 3 + k];
            out[i * 3 + j] = s;
        }
}

static double compute_ekin(const orc_md *md) { /* verlet.py:168-169 */
    double e = 0.0;
    for (int64_t n = 0; n < md->nnodes; n++)
        for (int d = 0; d < 3; d++) e += 0.5 * (md->vel[3 * n + d] * md->vel[3 * n + d] * md->masses[n]);
    return e;
}

static void mvv(const orc_md *md, double out[9]) { /* np.dot(vel.T * masses, vel) */
    for (int i = 0; i < 9; i++) out[i] = 0.0;
    for (int64_t n = 0; n < md->nnodes; n++)
        for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++) out[k * 3 + l] += md->vel[3 * n + k] * md->masses[n] * md->vel[3 * n + l];
}

static double ekin_baro(const orc_baro *b) { /* npt.py:748-757 */
    if (b->anisotropic) {
        double tr = 0.0;
        for (int i = 0; i < 9; i++) tr += b->vel_press[i] * b->vel_press[i];
        return 0.5 * b->mass_press * tr;
    }
    return 0.5 * b->mass_press * b->vel_press[0] * b->vel_press[0];
}

static void force_call(const orc_system *sys, orc_md *md, int with_vtens) {
    for (int i = 0; i < 9; i++) md->vtens[i] = 0.0;
    md->epot = orc_compute(sys, md->pos, md->rvecs, md->gpos, with_vtens ? md->vtens : NULL, md->work);
    md->nforce++;
}

/* npt.py:654-682 (update_baro_vel) */
static void update_baro_vel(const orc_system *sys, orc_md *md, int has_cv0, double chainvel0) {
    orc_baro *b = &md->baro;
    const int nvp = b->anisotropic ? 9 : 1;
    if (has_cv0)
        for (int i = 0; i < nvp; i++) b->vel_press[i] *= exp(-b->timestep * chainvel0 / 8.0);
    double pv[9], G[9];
    mvv(md, pv);
    for (int i = 0; i < 9; i++) pv[i] -= md->vtens[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) G[i * 3 + j] = 0.5 * (pv[j * 3 + i] + pv[i * 3 + j]);
    const double iso = 2.0 * md->ekin / md->ndof - b->press * orc_volume(md->rvecs);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) G[i * 3 + j] = (G[i * 3 + j] + (i == j ? iso : 0.0)) / b->mass_press;
    if (!b->anisotropic) {
        b->vel_press[0] += (G[0] + G[4] + G[8]) * b->timestep / 4.0;
    } else {
        if (b->vol_constraint) {
            const double tr = (G[0] + G[4] + G[8]) / b->dim;
            G[0] -= tr;
            G[4] -= tr;
            G[8] -= tr;
        }
        for (int i = 0; i < 9; i++) b->vel_press[i] += G[i] * b->timestep / 4.0;
    }
    if (has_cv0)
        for (int i = 0; i < nvp; i++) b->vel_press[i] *= exp(-b->timestep * chainvel0 / 8.0);
}

static void rmul3(double *x, int64_t n, const double R[9]) { /* rows of x times R */
    for (int64_t i = 0; i < n; i++) {
        double a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
        for (int j = 0; j < 3; j++) x[3 * i + j] = a * R[j] + b * R[3 + j] + c * R[6 + j];
    }
}

/* npt.py:653-736 */
static void baro_call(const orc_system *sys, orc_md *md, int has_cv0, double chainvel0) {
    orc_baro *b = &md->baro;
    update_baro_vel(sys, md, has_cv0, chainvel0);
    if (b->anisotropic) {
        double R[9];
        orc_sym_expm(b->vel_press, b->timestep / 2.0, R);
        rmul3(md->pos, md->nnodes, R);
        rmul3(md->rvecs, 3, R);
    } else {
        const double c = exp(b->vel_press[0] * b->timestep / 2.0);
        for (int64_t i = 0; i < 3 * md->nnodes; i++) md->pos[i] *= c;
        for (int i = 0; i < 9; i++) md->rvecs[i] *= c;
    }
    force_call(sys, md, 1);
    if (b->anisotropic) {
        double A[9], R[9];
        memcpy(A, b->vel_press, sizeof(A));
        if (!b->vol_constraint) {
            const double tr = (A[0] + A[4] + A[8]) / md->ndof;
            A[0] += tr;
            A[4] += tr;
            A[8] += tr;
        }
        orc_sym_expm(A, -b->timestep / 2.0, R);
        rmul3(md->vel, md->nnodes, R);
    } else {
        const double c = exp(-((1.0 + 3.0 / md->ndof) * b->vel_press[0]) * b->timestep / 2.0);
        for (int64_t i = 0; i < 3 * md->nnodes; i++) md->vel[i] *= c;
    }
    md->ekin = compute_ekin(md);
    update_baro_vel(sys, md, has_cv0, chainvel0);
}

/* verlet.py:171-190 */
void orc_md_properties(const orc_system *sys, orc_md *md) {
    double sg = 0.0, sd = 0.0;
    for (int64_t i = 0; i < 3 * md->nnodes; i++) {
        sg += md->gpos[i] * md->gpos[i];
        sd += md->delta[i] * md->delta[i];
    }
    md->rmsd_gpos = sqrt(sg / (3.0 * md->nnodes));
    md->rmsd_delta = sqrt(sd / (3.0 * md->nnodes));
    md->ekin = compute_ekin(md);
    md->temp = (md->ekin / md->ndof) * (2.0 / sys->boltzmann);
    md->etot = md->ekin + md->epot;
    md->econs = md->etot + md->econs_correction;
    /* ConsErrTracker.update / get, verlet.py:289-307 */
    if (md->ce_counter == 0) {
        md->ce_ekin_m = md->ekin;
        md->ce_econs_m = md->econs;
    } else {
        double t = md->ekin - md->ce_ekin_m;
        md->ce_ekin_m += t / (md->ce_counter + 1);
        md->ce_ekin_s += t * (md->ekin - md->ce_ekin_m);
        t = md->econs - md->ce_econs_m;
        md->ce_econs_m += t / (md->ce_counter + 1);
        md->ce_econs_s += t * (md->econs - md->ce_econs_m);
    }
    md->ce_counter++;
    md->cons_err = (md->ce_counter > 1) ? sqrt(md->ce_econs_s / md->ce_ekin_s) : 0.0;
    double pv[9];
    mvv(md, pv);
    const double vol = orc_volume(md->rvecs);
    for (int i = 0; i < 9; i++) md->ptens[i] = (pv[i] - md->vtens[i]) / vol;
    md->press = (md->ptens[0] + md->ptens[4] + md->ptens[8]) / 3.0;
}

/* verlet.py:119-137, nvt.py:507-523, npt.py:579-614 */
void orc_md_initialize(const orc_system *sys, orc_md *md) {
    const int64_t n3 = 3 * md->nnodes;
    for (int64_t i = 0; i < n3; i++) md->delta[i] = 0.0;
    md->nforce = 0;
    force_call(sys, md, 0); /* verlet.py:124 passes no vtens: it stays zero */
    memcpy(md->posold, md->pos, sizeof(double) * n3);
    if ((md->has_thermo || md->has_baro) && md->ndof <= 0.0) md->ndof = (double)(3 * md->nnodes - 3); /* utils.py:340-343 */
    if (md->has_thermo) {
        md->chain.timestep = md->timestep;
        orc_chain_set_ndof(&md->chain, md->ndof, sys->boltzmann);
    }
    if (md->has_baro) {
        orc_baro *b = &md->baro;
        b->timestep = md->timestep;
        const double angfreq = 2.0 * M_PI / b->timecon;
        b->mass_press = (md->ndof + b->dim * b->dim) * sys->boltzmann * b->temp / (angfreq * angfreq);
        if (b->vol_constraint) {
            const double tr = (b->vel_press[0] + b->vel_press[4] + b->vel_press[8]) / 3.0;
            b->vel_press[0] -= tr;
            b->vel_press[4] -= tr;
            b->vel_press[8] -= tr;
        }
        force_call(sys, md, 1); /* npt.py:612-614 */
    }
    if (md->ndof <= 0.0) md->ndof = (double)n3; /* verlet.py:131-132 */
    md->econs_correction = 0.0;
    md->ce_counter = 0;
    md->ce_ekin_m = md->ce_ekin_s = md->ce_econs_m = md->ce_econs_s = 0.0;
    orc_md_properties(sys, md);
}

/* verlet.py:140-166 */
void orc_md_step(const orc_system *sys, orc_md *md) {
    const double kb = sys->boltzmann;
    const int64_t n3 = 3 * md->nnodes;
    /* "pre": TBCombination.pre, npt.py:99-115 (barostat first, then thermostat) */
    if (md->has_baro) baro_call(sys, md, md->has_thermo, md->has_thermo ? md->chain.vel[0] : 0.0);
    if (md->has_thermo) {
        const double g1 = md->has_baro ? 2.0 * ekin_baro(&md->baro) - md->baro.baro_ndof * md->baro.temp * kb : 0.0;
        orc_chain_call(&md->chain, kb, &md->ekin, md->vel, n3, md->has_baro, g1);
    }
    /* verlet.py:144-154 */
    for (int64_t n = 0; n < md->nnodes; n++)
        for (int d = 0; d < 3; d++) {
            const double acc = -md->gpos[3 * n + d] / md->masses[n];
            md->vel[3 * n + d] += 0.5 * acc * md->timestep;
            md->pos[3 * n + d] += md->timestep * md->vel[3 * n + d];
        }
    force_call(sys, md, 1);
    for (int64_t n = 0; n < md->nnodes; n++)
        for (int d = 0; d < 3; d++) {
            const double acc = -md->gpos[3 * n + d] / md->masses[n];
            md->vel[3 * n + d] += 0.5 * acc * md->timestep;
        }
    md->ekin = compute_ekin(md);
    /* "post": TBCombination.post, npt.py:117-148 (thermostat first, then barostat) */
    double corr = 0.0;
    if (md->has_thermo) {
        const double g1 = md->has_baro ? 2.0 * ekin_baro(&md->baro) - md->baro.baro_ndof * md->baro.temp * kb : 0.0;
        orc_chain_call(&md->chain, kb, &md->ekin, md->vel, n3, md->has_baro, g1);
        corr += orc_chain_econs(&md->chain, kb); /* nvt.py:532 */
    }
    if (md->has_baro) {
        baro_call(sys, md, md->has_thermo, md->has_thermo ? md->chain.vel[0] : 0.0);
        corr += ekin_baro(&md->baro); /* npt.py:645-648 */
        if (!md->baro.vol_constraint) corr += md->baro.press * orc_volume(md->rvecs);
        if (md->has_thermo) corr += md->baro.baro_ndof * kb * md->chain.temp * md->chain.pos[0]; /* npt.py:144-148 */
    }
    if (md->has_thermo || md->has_baro) md->econs_correction = corr;
    /* verlet.py:158-166 */
    for (int64_t i = 0; i < n3; i++) {
        md->delta[i] = md->pos[i] - md->posold[i];
        md->posold[i] = md->pos[i];
    }
    md->time += md->timestep;
    orc_md_properties(sys, md);
    md->counter++;
}
