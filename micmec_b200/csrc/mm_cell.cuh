// mm_cell.cuh - per-cell elastic energy / gradient in closed form (fp64), shared by every force kernel.
//
// Reference math (relative to /root/reference):
//   original model  micmec/pes/nanocell_original.py:69-84 (energy), :105-132 (gradient)
//   default model   micmec/pes/nanocell.py:63-77 (energy), :98-132 (gradient)
//   state mixing    micmec/pes/mmff.py:377-398
// The reference evaluates these with ~20 einsum calls over constant +-1 stencil tables
// (micmec/pes/nanocell_utils.py:32-110).  Here the stencils are folded away analytically:
//
//   H   (3x3, rows = cell edge vectors)           original: mean of the 4 edges per axis; default: the 3 edges at corner a
//   G   = h0^-1 H                                  (reference forms M_ = G^T)
//   eps = 1/2 (G G^T - I)                          symmetric, 6 unique entries (Voigt order 00,11,22,12,02,01)
//   s   = A eps                                    s = sym(C:eps); A is the EXACT 6x6 fold of the full 81-entry tensor:
//                                                  A[I][J] = 1/2 (C_ijkl + C_jikl) (+ the l<->k partner when k != l).
//                                                  A is NOT symmetrised: the fixtures' tensors lack major symmetry and
//                                                  the reference contracts the symmetrised strain derivative with C:eps.
//   E   = 1/2 V0 eps:s
//   D   = V0 h0^-T s G                             dE/dH as the reference defines it; g_v = sum_i W[v][i] D[i][:]
//   vir = V0 G^T s G = D^T H                       = sum_v g_v (x) r_v (mmff.py:320-323) because sum_v g_v = 0
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/micmec_b200.h"

namespace mm {

// One metastable state, pre-reduced on the host in fp64 (mm_api.cu: build_params).
struct StateP {
    double hi[9];   // h0^-1
    double A[36];   // folded elasticity, row-major [I][J]
    double v0;      // det(h0)  (np.linalg.det, sign kept)
    double efree;   // typeN/free_energy[s]
};

struct KParams {
    int32_t ntypes;
    int32_t model;
    int32_t nstates[MM_MAX_TYPES];
    int32_t offset[MM_MAX_TYPES];
    double kT[MM_MAX_TYPES];       // boltzmann * effective_temp
    StateP st[MM_MAX_STATES];
};

// energy (without efree), T = s G and G for one state and one edge matrix H
struct StateOut {
    double e;
    double D[9];    // weight * V0 h0^-T s G
    double vir[6];  // weight * V0 G^T s G   (00,11,22,12,02,01)
};

__device__ __forceinline__ void state_eval(const double H[9], const StateP &P, double weight, StateOut &o) {
    double G[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            G[i * 3 + j] = fma(P.hi[i * 3 + 2], H[6 + j], fma(P.hi[i * 3 + 1], H[3 + j], P.hi[i * 3] * H[j]));
    double eps[6];
    eps[0] = 0.5 * (fma(G[2], G[2], fma(G[1], G[1], G[0] * G[0])) - 1.0);
    eps[1] = 0.5 * (fma(G[5], G[5], fma(G[4], G[4], G[3] * G[3])) - 1.0);
    eps[2] = 0.5 * (fma(G[8], G[8], fma(G[7], G[7], G[6] * G[6])) - 1.0);
    eps[3] = 0.5 * fma(G[5], G[8], fma(G[4], G[7], G[3] * G[6]));
    eps[4] = 0.5 * fma(G[2], G[8], fma(G[1], G[7], G[0] * G[6]));
    eps[5] = 0.5 * fma(G[2], G[5], fma(G[1], G[4], G[0] * G[3]));
    double s[6];
#pragma unroll
    for (int I = 0; I < 6; I++) {
        double acc = P.A[I * 6] * eps[0];
#pragma unroll
        for (int J = 1; J < 6; J++) acc = fma(P.A[I * 6 + J], eps[J], acc);
        s[I] = acc;
    }
    const double dens = fma(2.0, fma(eps[5], s[5], fma(eps[4], s[4], eps[3] * s[3])),
                            fma(eps[2], s[2], fma(eps[1], s[1], eps[0] * s[0])));
    const double wv = weight * P.v0;
    o.e = 0.5 * wv * dens;
    // T = s G with s = [[s0 s5 s4],[s5 s1 s3],[s4 s3 s2]]
    double T[9];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        T[j] = fma(s[4], G[6 + j], fma(s[5], G[3 + j], s[0] * G[j]));
        T[3 + j] = fma(s[3], G[6 + j], fma(s[1], G[3 + j], s[5] * G[j]));
        T[6 + j] = fma(s[2], G[6 + j], fma(s[3], G[3 + j], s[4] * G[j]));
    }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            o.D[i * 3 + j] = wv * fma(P.hi[6 + i], T[6 + j], fma(P.hi[3 + i], T[3 + j], P.hi[i] * T[j]));
    // vir = wv G^T T (symmetric)
    o.vir[0] = wv * fma(G[6], T[6], fma(G[3], T[3], G[0] * T[0]));
    o.vir[1] = wv * fma(G[7], T[7], fma(G[4], T[4], G[1] * T[1]));
    o.vir[2] = wv * fma(G[8], T[8], fma(G[5], T[5], G[2] * T[2]));
    o.vir[3] = wv * fma(G[7], T[8], fma(G[4], T[5], G[1] * T[2]));
    o.vir[4] = wv * fma(G[6], T[8], fma(G[3], T[5], G[0] * T[2]));
    o.vir[5] = wv * fma(G[6], T[7], fma(G[3], T[4], G[0] * T[1]));
}

// vertex offsets, micmec/utils.py:43-52:  v0 000, v1 100, v2 010, v3 001, v4 110, v5 101, v6 011, v7 111
__device__ __forceinline__ constexpr int vbit(int v, int axis) {
    return axis == 0 ? (v == 1 || v == 4 || v == 5 || v == 7)
         : axis == 1 ? (v == 2 || v == 4 || v == 6 || v == 7)
                     : (v == 3 || v == 5 || v == 6 || v == 7);
}
// vertex index from offset bits
__device__ __forceinline__ constexpr int vidx(int dx, int dy, int dz) {
    return (dx == 0 && dy == 0 && dz == 0) ? 0 : (dx == 1 && dy == 0 && dz == 0) ? 1
         : (dx == 0 && dy == 1 && dz == 0) ? 2 : (dx == 0 && dy == 0 && dz == 1) ? 3
         : (dx == 1 && dy == 1 && dz == 0) ? 4 : (dx == 1 && dy == 0 && dz == 1) ? 5
         : (dx == 0 && dy == 1 && dz == 1) ? 6 : 7;
}

// One cell, one state, either model.  R = unwrapped vertices [8][3].  Returns energy (with efree) and the
// 8x3 gradient g, plus the cell virial (6 unique).
template <int MODEL>
__device__ __forceinline__ void cell_state(const double R[24], const StateP &P, double &e, double g[24], double vir[6]) {
    if (MODEL == MM_MODEL_ORIGINAL) {
        // H[i][:] = 1/4 sum_v (2 d_vi - 1) r_v      (nanocell_original.py:71-75 with the multiplicator table)
        double H[9];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double r0 = R[c], r1 = R[3 + c], r2 = R[6 + c], r3 = R[9 + c];
            const double r4 = R[12 + c], r5 = R[15 + c], r6 = R[18 + c], r7 = R[21 + c];
            H[c] = 0.25 * ((r1 - r0) + (r4 - r2) + (r5 - r3) + (r7 - r6));
            H[3 + c] = 0.25 * ((r2 - r0) + (r4 - r1) + (r6 - r3) + (r7 - r5));
            H[6 + c] = 0.25 * ((r3 - r0) + (r5 - r1) + (r6 - r2) + (r7 - r4));
        }
        StateOut o;
        state_eval(H, P, 1.0, o);
        e = o.e + P.efree;
#pragma unroll
        for (int v = 0; v < 8; v++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double a = vbit(v, 0) ? o.D[c] : -o.D[c];
                const double b = vbit(v, 1) ? o.D[3 + c] : -o.D[3 + c];
                const double d = vbit(v, 2) ? o.D[6 + c] : -o.D[6 + c];
                g[v * 3 + c] = 0.25 * ((a + b) + d);
            }
#pragma unroll
        for (int k = 0; k < 6; k++) vir[k] = o.vir[k];
    } else {
        double esum = 0.0;
#pragma unroll
        for (int k = 0; k < 24; k++) g[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) vir[k] = 0.0;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            // corner a: edge i joins the two vertices that differ from a only in bit i, oriented +axis
            const int ax = vbit(a, 0), ay = vbit(a, 1), az = vbit(a, 2);
            const int lo[3] = {vidx(0, ay, az), vidx(ax, 0, az), vidx(ax, ay, 0)};
            const int hi[3] = {vidx(1, ay, az), vidx(ax, 1, az), vidx(ax, ay, 1)};
            double H[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < 3; c++) H[i * 3 + c] = R[hi[i] * 3 + c] - R[lo[i] * 3 + c];
            StateOut o;
            state_eval(H, P, 0.125, o);
            esum += o.e;
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    g[hi[i] * 3 + c] += o.D[i * 3 + c];
                    g[lo[i] * 3 + c] -= o.D[i * 3 + c];
                }
#pragma unroll
            for (int k = 0; k < 6; k++) vir[k] += o.vir[k];
        }
        e = esum + P.efree;
    }
}

// All states of a cell + Boltzmann mixing (mmff.py:377-398).  Single-state types skip exp/log entirely, which is
// exact: w = exp(0) = 1, log(1) = 0.
template <int MODEL>
__device__ __forceinline__ void cell_eval(const double R[24], const KParams &kp, int type, double &e, double g[24],
                                          double vir[6]) {
    const int ns = kp.nstates[type];
    const int off = kp.offset[type];
    if (ns == 1) {
        cell_state<MODEL>(R, kp.st[off], e, g, vir);
        return;
    }
    // two passes: energies first (to find the minimum), then weights.  Recomputing the gradient of each state in the
    // second pass would double the work, so keep a running weighted sum with online rescaling instead.
    const double kT = kp.kT[type];
    double emin = 0.0, wsum = 0.0;
#pragma unroll 1
    for (int s = 0; s < ns; s++) {
        double es, gs[24], vs[6];
        cell_state<MODEL>(R, kp.st[off + s], es, gs, vs);
        if (s == 0) {
            emin = es;
            wsum = 1.0;
#pragma unroll
            for (int k = 0; k < 24; k++) g[k] = gs[k];
#pragma unroll
            for (int k = 0; k < 6; k++) vir[k] = vs[k];
        } else if (es < emin) {
            // new minimum: rescale what was accumulated so far by exp(-(emin_old - es)/kT)
            const double f = exp(-(emin - es) / kT);
            wsum = fma(wsum, f, 1.0);
#pragma unroll
            for (int k = 0; k < 24; k++) g[k] = fma(g[k], f, gs[k]);
#pragma unroll
            for (int k = 0; k < 6; k++) vir[k] = fma(vir[k], f, vs[k]);
            emin = es;
        } else {
            const double w = exp(-(es - emin) / kT);
            wsum += w;
#pragma unroll
            for (int k = 0; k < 24; k++) g[k] = fma(w, gs[k], g[k]);
#pragma unroll
            for (int k = 0; k < 6; k++) vir[k] = fma(w, vs[k], vir[k]);
        }
    }
    const double inv = 1.0 / wsum;
#pragma unroll
    for (int k = 0; k < 24; k++) g[k] *= inv;
#pragma unroll
    for (int k = 0; k < 6; k++) vir[k] *= inv;
    e = emin - kT * log(wsum);
}

}  // namespace mm
