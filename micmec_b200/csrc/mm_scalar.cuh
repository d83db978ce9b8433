// mm_scalar.cuh - the O(1) algebra of the thermostat / barostat / conserved quantity, as device functions shared by the
// single-block scalar kernel (mm_md.cu: k_scalar) and by the tail of the marching kernel (mm_march2.cuh), whose last
// block runs it right after reducing the block partials.
//
// Stands in for (relative to /root/reference)
//   NHChain.__call__ / set_ndof / get_econs_correction              micmec/sampling/nvt.py:393-458
//   MTKBarostat.init/pre/post/baro/add_press_cont                   micmec/sampling/npt.py:579-757
//   TBCombination.pre/post                                          micmec/sampling/npt.py:99-148
//   VerletIntegrator.compute_properties, ConsErrTracker             micmec/sampling/verlet.py:171-190, 275-307
#pragma once
#include <cuda_runtime.h>

#include <cmath>

#include "../../include/micmec_b200.h"
#include "mm_structured.cuh"

namespace mm {

struct MDState {
    double boltzmann;
    double timestep, time, ndof;
    long long counter, nforce;
    // last force evaluation / kinetic moments
    double epot, vir[6];
    double ekin, mvv[6];  // mvv = sum m v (x) v (00,11,22,12,02,01) of the TRUE velocities
    double sum_g2, sum_d2;
    double rvecs[9];
    // pending transforms
    double Mvel[9];  // v_true = v_stored . Mvel
    double Rpos[9];  // to be applied by k_apply_pos
    // structured path: positions stay in the frame of their last write; x_true = (x_stored + shifts(rv_stored)).Rpend
    double Rpend[9], rv_stored[9];
    // Nose-Hoover chain (thermo_kind 0) or Berendsen weak coupling (thermo_kind 1: ch_temp, ch_timecon and be_corr only)
    int has_thermo, chain_len;
    double ch_temp, ch_timecon;
    double ch_pos[MM_MAX_CHAIN], ch_vel[MM_MAX_CHAIN], ch_mass[MM_MAX_CHAIN];
    // MTK barostat
    int has_baro, aniso, volc, baro_ndof;
    double b_temp, b_press, b_timecon, mass_press;
    double vp[9];
    // conserved quantity
    double econs_corr;
    long long ce_n;
    double ce_ekin_m, ce_ekin_s, ce_econs_m, ce_econs_s;
    // Langevin thermostat (nvt.py:165-218): accumulated econs_correction = sum of (ekin before - ekin after) of every call
    int has_langevin, thermo_kind;
    double be_corr;  // Berendsen: accumulated econs_correction (nvt.py:153-162)
    double lg_temp, lg_timecon, lg_corr;
    unsigned long long lg_seed;
    // properties (verlet.py:171-190)
    double temp, etot, econs, cons_err, press, ptens[9], rmsd_gpos, rmsd_delta, volume;
};

enum : unsigned {
    OP_RESET_MVEL = 1u << 0,
    OP_TAKE_FORCE = 1u << 1,
    OP_TAKE_KIN = 1u << 2,
    OP_BARO_B = 1u << 3,
    OP_THERMO = 1u << 4,
    OP_BARO_A = 1u << 5,
    OP_ECONS = 1u << 6,
    OP_PROPS = 1u << 7,
    OP_ADVANCE = 1u << 8,
    OP_ZERO_VIR = 1u << 9,
    OP_TAKE_DELTA = 1u << 10,
    OP_SETUP = 1u << 11,
    OP_POS_WRITTEN = 1u << 12,
    OP_NEXT_THERMO = 1u << 13,  // the NEXT step's thermostat "pre" call, merged into this step's last scalar launch
    OP_NEXT_BARO_A = 1u << 14,  // the NEXT step's first barostat half, likewise
    OP_LANG_A = 1u << 15,       // Langevin thermostat: book the first half-step of a k_langevin launch (a "post", or a lone "pre")
    OP_LANG_B = 1u << 16,       // ... and its second half-step (the NEXT step's "pre", merged into the same launch)
};

// ---- counter-based random numbers (Philox4x32-10, Salmon et al., SC'11): the device-resident stochastic hooks -------------
static __device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4]) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
        c1 = (unsigned)p1;
        c3 = (unsigned)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// two independent standard normal deviates (Box-Muller on 64-bit uniforms) of stream (seed, c0..c3)
static __device__ __forceinline__ void philox_normal2(unsigned long long seed, unsigned c0, unsigned c1, unsigned c2, unsigned c3, double &z0,
                                                      double &z1) {
    unsigned r[4];
    philox4x32_10(c0, c1, c2, c3, (unsigned)seed, (unsigned)(seed >> 32), r);
    const double u1 = ((double)(((unsigned long long)r[0] << 32) | r[1]) + 0.5) * 5.421010862427522e-20;  // 2^-64
    const double u2 = ((double)(((unsigned long long)r[2] << 32) | r[3]) + 0.5) * 5.421010862427522e-20;
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}
// one uniform deviate in (0, 1) of stream (seed, c0..c3)
static __device__ __forceinline__ double philox_uniform(unsigned long long seed, unsigned c0, unsigned c1, unsigned c2, unsigned c3) {
    unsigned r[4];
    philox4x32_10(c0, c1, c2, c3, (unsigned)seed, (unsigned)(seed >> 32), r);
    return ((double)(((unsigned long long)r[0] << 32) | r[1]) + 0.5) * 5.421010862427522e-20;
}

// ------------------------------------------------------------------------------------------- 3x3 helpers ----
static __device__ __forceinline__ void mat_mul(const double *a, const double *b, double *c) {  // c = a b (c may alias neither)
    // unrolled: nine independent chains of three - this code runs on ONE thread between two marching launches, its
    // latency is on the critical path of every step
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

static __device__ __forceinline__ double volume_of(const double *r) {  // domain.c:42-48
    return fabs(r[0] * (r[4] * r[8] - r[5] * r[7]) + r[1] * (r[5] * r[6] - r[3] * r[8]) + r[2] * (r[3] * r[7] - r[4] * r[6]));
}

// out = Q exp(scale * w) Q^T for the symmetric matrix whose lower triangle is a (numpy.linalg.eigh reads the
// lower triangle; npt.py:686-690, 710-721).  Cyclic Jacobi, one thread.
static __device__ void sym_expm(const double *a_in, double scale, double *out) {
    double a[3][3], q[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) a[i][j] = (i >= j) ? a_in[i * 3 + j] : a_in[j * 3 + i];
    {
        // In MD the argument X = scale * A is tiny (barostat velocity x half a time step ~ 1e-6): the power series
        // reaches 1e-17 in a few terms of independent 3x3 products, whereas the Jacobi sweeps below are a serial chain
        // of divisions and square roots (~10 us on one thread).  Same result to rounding; Jacobi remains for large X.
        double X[9], nrm = 0.0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                X[i * 3 + j] = scale * a[i][j];
                nrm += X[i * 3 + j] * X[i * 3 + j];
            }
        if (nrm < 1e-4) {  // ||X||_F < 1e-2: 9 terms give < 1e-2^10 / 10! = 3e-27 relative truncation
            double term[9], sum[9], t[9];
            for (int i = 0; i < 9; i++) {
                term[i] = X[i];
                sum[i] = ((i % 4 == 0) ? 1.0 : 0.0) + X[i];
            }
#pragma unroll
            for (int k = 2; k <= 9; k++) {
                mat_mul(term, X, t);
                const double inv = 1.0 / (double)k;  // compile-time constant once unrolled
#pragma unroll
                for (int i = 0; i < 9; i++) {
                    term[i] = t[i] * inv;
                    sum[i] += term[i];
                }
            }
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) out[i * 3 + j] = 0.5 * (sum[i * 3 + j] + sum[j * 3 + i]);
            return;
        }
    }
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-40 * (diag + 1e-300)) break;
        for (int p = 0; p < 2; p++)
            for (int r = p + 1; r < 3; r++) {
                if (a[p][r] == 0.0) continue;
                const double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {
                    const double akp = a[k][p], akr = a[k][r];
                    a[k][p] = c * akp - s * akr;
                    a[k][r] = s * akp + c * akr;
                }
                for (int k = 0; k < 3; k++) {
                    const double apk = a[p][k], ark = a[r][k];
                    a[p][k] = c * apk - s * ark;
                    a[r][k] = s * apk + c * ark;
                }
                for (int k = 0; k < 3; k++) {
                    const double qkp = q[k][p], qkr = q[k][r];
                    q[k][p] = c * qkp - s * qkr;
                    q[k][r] = s * qkp + c * qkr;
                }
            }
    }
    double f[3];
    for (int k = 0; k < 3; k++) f[k] = exp(scale * a[k][k]);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) out[i * 3 + j] = q[i][0] * f[0] * q[j][0] + q[i][1] * f[1] * q[j][1] + q[i][2] * f[2] * q[j][2];
}

static __device__ __forceinline__ void sym6_to_full(const double *s, double *m) {
    m[0] = s[0]; m[4] = s[1]; m[8] = s[2];
    m[5] = m[7] = s[3]; m[2] = m[6] = s[4]; m[1] = m[3] = s[5];
}

static __device__ __forceinline__ void full_to_sym6(const double *m, double *s) {
    s[0] = m[0]; s[1] = m[4]; s[2] = m[8];
    s[3] = 0.5 * (m[5] + m[7]); s[4] = 0.5 * (m[2] + m[6]); s[5] = 0.5 * (m[1] + m[3]);
}

// ------------------------------------------------------------------------------------ scalar sub-steps -------
static __device__ double ekin_baro(const MDState &s) {  // npt.py:748-757
    if (s.aniso) {
        double tr = 0.0;
        for (int i = 0; i < 9; i++) tr += s.vp[i] * s.vp[i];
        return 0.5 * s.mass_press * tr;
    }
    return 0.5 * s.mass_press * s.vp[0] * s.vp[0];
}

static __device__ void chain_bead(MDState &s, int k, double ekin, bool has_g1, double g1) {  // nvt.py:411-435
    const double kb = s.boltzmann;
    double g;
    if (k == 0) {
        g = 2.0 * ekin - s.ndof * s.ch_temp * kb;
        if (has_g1) g += g1;
    } else {
        g = s.ch_mass[k - 1] * s.ch_vel[k - 1] * s.ch_vel[k - 1] - s.ch_temp * kb;
    }
    g /= s.ch_mass[k];
    if (k == s.chain_len - 1) {
        s.ch_vel[k] += g * s.timestep / 4.0;
    } else {
        s.ch_vel[k] *= exp(-s.ch_vel[k + 1] * s.timestep / 8.0);
        s.ch_vel[k] += g * s.timestep / 4.0;
        s.ch_vel[k] *= exp(-s.ch_vel[k + 1] * s.timestep / 8.0);
    }
}

// NHChain.__call__ (nvt.py:410-451): the velocity scaling goes into the pending matrix and the moments
static __device__ void thermo_call(MDState &s) {
    if (s.thermo_kind == 1) {  // BerendsenThermostat.pre (nvt.py:146-162): one global velocity scale, deferred like the chain's
        const double temp_now = 2.0 * s.ekin / (s.boltzmann * s.ndof);
        const double scale = sqrt(1.0 + s.timestep / s.ch_timecon * (s.ch_temp / temp_now - 1.0));
        for (int i = 0; i < 9; i++) s.Mvel[i] *= scale;
        for (int i = 0; i < 6; i++) s.mvv[i] *= scale * scale;
        s.be_corr += (1.0 - scale * scale) * s.ekin;
        s.ekin *= scale * scale;
        return;
    }
    if (s.thermo_kind == 2) {  // CSVRThermostat.pre (nvt.py:259-271; Bussi, Donadio, Parrinello 2007): one stochastic velocity scale
        const double c = exp(-s.timestep / s.ch_timecon);
        const unsigned step = (unsigned)s.counter, step_hi = (unsigned)((unsigned long long)s.counter >> 32);
        double R, z;
        philox_normal2(s.lg_seed, step, step_hi, 0x43535652u, 0u, R, z);
        // S = sum of ndof - 1 squared normal deviates = 2 Gamma((ndof - 1) / 2)  (Marsaglia & Tsang 2000; the reference draws
        // the ndof - 1 deviates themselves)
        const double shape = 0.5 * (s.ndof - 1.0), d = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
        double S = 2.0 * d;
        for (unsigned it = 1; it < 1000u; it++) {
            double x, x2;
            philox_normal2(s.lg_seed, step, step_hi, 0x43535652u, it, x, x2);
            double v = 1.0 + cc * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            const double u = philox_uniform(s.lg_seed, step, step_hi, 0x43535653u, it);
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
                S = 2.0 * d * v;
                break;
            }
        }
        const double kin = 0.5 * s.ndof * s.boltzmann * s.ch_temp;
        const double fact = (1.0 - c) * kin / s.ndof / s.ekin;
        const double arg = R + sqrt(c / fact);
        const double alpha = (arg >= 0.0 ? 1.0 : -1.0) * sqrt(c + (S + R * R) * fact + 2.0 * R * sqrt(c * fact));
        for (int i = 0; i < 9; i++) s.Mvel[i] *= alpha;
        for (int i = 0; i < 6; i++) s.mvv[i] *= alpha * alpha;
        s.be_corr += (1.0 - alpha * alpha) * s.ekin;
        s.ekin *= alpha * alpha;
        return;
    }
    const bool has_g1 = s.has_baro != 0;  // TBCombination.pre/post, npt.py:107-113, 118-125
    const double g1 = has_g1 ? 2.0 * ekin_baro(s) - s.baro_ndof * s.b_temp * s.boltzmann : 0.0;  // npt.py:738-746
    double ekin = s.ekin;
    for (int k = s.chain_len - 1; k >= 0; k--) chain_bead(s, k, ekin, has_g1, g1);
    for (int k = 0; k < s.chain_len; k++) s.ch_pos[k] += s.ch_vel[k] * s.timestep / 2.0;
    const double factor = exp(-s.ch_vel[0] * s.timestep / 2.0);
    for (int i = 0; i < 9; i++) s.Mvel[i] *= factor;
    for (int i = 0; i < 6; i++) s.mvv[i] *= factor * factor;
    ekin *= factor * factor;
    for (int k = 0; k < s.chain_len; k++) chain_bead(s, k, ekin, has_g1, g1);
    s.ekin = ekin;
}

// update_baro_vel, npt.py:654-682
static __device__ void update_baro_vel(MDState &s) {
    const bool has_cv0 = s.has_thermo != 0;  // TBCombination hands chain.vel[0] to the barostat, npt.py:102-106
    const double damp = has_cv0 ? exp(-s.timestep * s.ch_vel[0] / 8.0) : 1.0;
    const int nvp = s.aniso ? 9 : 1;
    if (has_cv0)
        for (int i = 0; i < nvp; i++) s.vp[i] *= damp;
    double G[9], pv[6];
    for (int i = 0; i < 6; i++) pv[i] = s.mvv[i] - s.vir[i];  // both symmetric here: 0.5 (pv^T + pv) is a no-op
    sym6_to_full(pv, G);
    const double iso = 2.0 * s.ekin / s.ndof - s.b_press * volume_of(s.rvecs);
#pragma unroll
    for (int i = 0; i < 9; i++) G[i] = (G[i] + ((i % 4 == 0) ? iso : 0.0)) / s.mass_press;  // nine independent divisions
    if (!s.aniso) {
        s.vp[0] += (G[0] + G[4] + G[8]) * s.timestep / 4.0;
    } else {
        if (s.volc) {
            const double tr = (G[0] + G[4] + G[8]) / 3.0;
            G[0] -= tr; G[4] -= tr; G[8] -= tr;
        }
        for (int i = 0; i < 9; i++) s.vp[i] += G[i] * s.timestep / 4.0;
    }
    if (has_cv0)
        for (int i = 0; i < nvp; i++) s.vp[i] *= damp;
}

// first half of MTKBarostat.baro (npt.py:683-700): barostat velocity, position/cell rotation
static __device__ void baro_a(MDState &s) {
    update_baro_vel(s);
    if (s.aniso) {
        sym_expm(s.vp, s.timestep / 2.0, s.Rpos);
    } else {
        const double c = exp(s.vp[0] * s.timestep / 2.0);
        for (int i = 0; i < 9; i++) s.Rpos[i] = (i % 4 == 0) ? c : 0.0;
    }
    double nr[9];
    mat_mul(s.rvecs, s.Rpos, nr);
    for (int i = 0; i < 9; i++) s.rvecs[i] = nr[i];
    mat_mul(s.Rpend, s.Rpos, nr);
    for (int i = 0; i < 9; i++) s.Rpend[i] = nr[i];
}

// second half of MTKBarostat.baro (npt.py:708-736): velocity rotation (deferred), kinetic energy, barostat velocity
static __device__ void baro_b(MDState &s) {
    double R[9];
    if (s.aniso) {
        double A[9];
        for (int i = 0; i < 9; i++) A[i] = s.vp[i];
        if (!s.volc) {
            const double tr = (A[0] + A[4] + A[8]) / s.ndof;
            A[0] += tr; A[4] += tr; A[8] += tr;
        }
        sym_expm(A, -s.timestep / 2.0, R);
    } else {
        const double c = exp(-((1.0 + 3.0 / s.ndof) * s.vp[0]) * s.timestep / 2.0);
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? c : 0.0;
    }
    double t[9], m[9], mt[9];
    mat_mul(s.Mvel, R, t);
    for (int i = 0; i < 9; i++) s.Mvel[i] = t[i];
    // sum m (vR)(x)(vR) = R^T (sum m v(x)v) R
    sym6_to_full(s.mvv, m);
    mat_mul(m, R, mt);
    double Rt[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rt[i * 3 + j] = R[j * 3 + i];
    mat_mul(Rt, mt, m);
    full_to_sym6(m, s.mvv);
    s.ekin = 0.5 * (s.mvv[0] + s.mvv[1] + s.mvv[2]);
    update_baro_vel(s);
}

static __device__ void econs_update(MDState &s) {
    double corr = 0.0;
    const double kb = s.boltzmann;
    if (s.has_thermo && s.thermo_kind != 0) corr += s.be_corr;  // Berendsen / CSVR: accumulated (1 - scale^2) ekin
    if (s.has_thermo && s.thermo_kind == 0) {  // nvt.py:453-458
        const double kt = kb * s.ch_temp;
        double a = 0.0, p = 0.0;
        for (int k = 0; k < s.chain_len; k++) a += s.ch_vel[k] * s.ch_vel[k] * s.ch_mass[k];
        for (int k = 1; k < s.chain_len; k++) p += s.ch_pos[k];
        corr += 0.5 * a + kt * (s.ndof * s.ch_pos[0] + p);
    }
    if (s.has_baro) {  // npt.py:644-651, 134-148
        corr += ekin_baro(s);
        if (!s.volc) corr += s.b_press * volume_of(s.rvecs);
        if (s.has_thermo) corr += s.baro_ndof * kb * s.ch_temp * s.ch_pos[0];
    }
    if (s.has_langevin) corr += s.lg_corr;
    s.econs_corr = corr;
}

static __device__ void properties(MDState &s, double n3) {  // verlet.py:171-190
    s.rmsd_gpos = sqrt(s.sum_g2 / n3);
    s.rmsd_delta = sqrt(s.sum_d2 / n3);
    s.temp = (s.ekin / s.ndof) * (2.0 / s.boltzmann);
    s.etot = s.ekin + s.epot;
    s.econs = s.etot + s.econs_corr;
    if (s.ce_n == 0) {  // verlet.py:289-307
        s.ce_ekin_m = s.ekin;
        s.ce_econs_m = s.econs;
    } else {
        double t = s.ekin - s.ce_ekin_m;
        s.ce_ekin_m += t / (double)(s.ce_n + 1);
        s.ce_ekin_s += t * (s.ekin - s.ce_ekin_m);
        t = s.econs - s.ce_econs_m;
        s.ce_econs_m += t / (double)(s.ce_n + 1);
        s.ce_econs_s += t * (s.econs - s.ce_econs_m);
    }
    s.ce_n++;
    s.cons_err = (s.ce_n > 1) ? sqrt(s.ce_econs_s / s.ce_ekin_s) : 0.0;
    s.volume = volume_of(s.rvecs);
    double m[9], v[9];
    sym6_to_full(s.mvv, m);
    sym6_to_full(s.vir, v);
    if (s.volume > 0.0) {  // verlet.py:185: only for periodic systems
#pragma unroll
        for (int i = 0; i < 9; i++) s.ptens[i] = (m[i] - v[i]) / s.volume;
        s.press = (s.ptens[0] + s.ptens[4] + s.ptens[8]) / 3.0;
    }
}

// lg: sums of one k_langevin launch - [0] sum m v^2 before, [1..6] sum m v (x) v after its first half-step, [7] sum m v^2 after
// its second half-step (nvt.py:199-218: econs_correction += ekin0 - ekin1 around every call)
static __device__ __noinline__ void scalar_ops(MDState &s, double *rvecs_dev, StepConsts *sc, unsigned ops, const double *fr, const double *kn,
                           const double *dl, int nbn, double n3, const double *lg = nullptr) {
    if (ops & OP_RESET_MVEL)
        for (int i = 0; i < 9; i++) s.Mvel[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (ops & OP_POS_WRITTEN)  // the stored positions are the true ones again
        for (int i = 0; i < 9; i++) {
            s.Rpend[i] = (i % 4 == 0) ? 1.0 : 0.0;
            s.rv_stored[i] = s.rvecs[i];
        }
    if (ops & OP_TAKE_FORCE) {
        s.epot = fr[0];
        for (int k = 0; k < 6; k++) s.vir[k] = fr[1 + k];
        if (nbn > 0) s.sum_g2 = kn[6];
        s.nforce++;
    }
    if (ops & OP_ZERO_VIR)  // verlet.py:124 passes no vtens to the first compute: it stays zero
        for (int k = 0; k < 6; k++) s.vir[k] = 0.0;
    if (ops & OP_TAKE_KIN) {
        for (int k = 0; k < 6; k++) s.mvv[k] = kn[k];
        s.ekin = 0.5 * (kn[0] + kn[1] + kn[2]);
    }
    if (ops & OP_TAKE_DELTA) s.sum_d2 = dl[0];
    if (ops & OP_SETUP) {  // nvt.py:393-400, npt.py:591-596, verlet.py:131-132, sampling/utils.py:340-343
        if ((s.has_thermo || s.has_baro) && s.ndof <= 0.0) s.ndof = n3 - 3.0;
        if (s.ndof <= 0.0) s.ndof = n3;
        if (s.has_thermo && s.thermo_kind == 0) {
            const double afreq = 2.0 * M_PI / s.ch_timecon;
            for (int k = 0; k < s.chain_len; k++) s.ch_mass[k] = s.boltzmann * s.ch_temp / (afreq * afreq);
            s.ch_mass[0] *= s.ndof;
        }
        if (s.has_baro) {
            const double angfreq = 2.0 * M_PI / s.b_timecon;
            s.mass_press = (s.ndof + 9.0) * s.boltzmann * s.b_temp / (angfreq * angfreq);
            if (s.volc) {  // npt.py:606-608
                const double tr = (s.vp[0] + s.vp[4] + s.vp[8]) / 3.0;
                s.vp[0] -= tr; s.vp[4] -= tr; s.vp[8] -= tr;
            }
        }
    }
    if (ops & OP_BARO_B) baro_b(s);
    if (ops & OP_THERMO) thermo_call(s);
    if (ops & OP_BARO_A) {
        baro_a(s);
        for (int i = 0; i < 9; i++) rvecs_dev[i] = s.rvecs[i];
    }
    if ((ops & OP_LANG_A) && lg) {
        const double after = 0.5 * (lg[1] + lg[2] + lg[3]);
        s.lg_corr += 0.5 * lg[0] - after;
        for (int k = 0; k < 6; k++) s.mvv[k] = lg[1 + k];
        s.ekin = after;
    }
    if (ops & OP_ECONS) econs_update(s);
    if (ops & OP_ADVANCE) {
        s.time += s.timestep;
        s.counter++;
    }
    if (ops & OP_PROPS) properties(s, n3);
    if ((ops & OP_LANG_B) && lg) {
        s.lg_corr += 0.5 * (lg[1] + lg[2] + lg[3]) - 0.5 * lg[7];
        s.ekin = 0.5 * lg[7];
    }
    if (ops & OP_NEXT_THERMO) thermo_call(s);
    if (ops & OP_NEXT_BARO_A) {
        baro_a(s);
        for (int i = 0; i < 9; i++) rvecs_dev[i] = s.rvecs[i];
    }
    if (sc) {
        for (int i = 0; i < 9; i++) {
            sc->Rpend[i] = s.Rpend[i];
            sc->Mvel[i] = s.Mvel[i];
            sc->rv[i] = s.rv_stored[i];
        }
        sc->dt = s.timestep;
    }
}

}  // namespace mm
