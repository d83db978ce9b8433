// mm_structured.cuh - data structures of the structured-grid (full periodic nx x ny x nz) fast path
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/micmec_b200.h"

namespace mm {

// Per-state constants of the marching kernel (mm_march.cuh: metric formulation), folded on the host in fold_sparams:
//   Hs = 4 H is what the separable stencil produces;  d = Hs Hs^T - c0;  Sq = Bq d;  E = 1/4 d:Sq;  D' = D / 4 = Sq Hs
struct SState {
    double Bq[36];   // V0/512 K' A K  (K, K': the congruences with h0^-1 in Voigt form)
    double c0[6];    // 16 h0 h0^T, Voigt order 00 11 22 12 02 01
    double efree;
};

struct SParams {
    int32_t ntypes;
    int32_t nstates[MM_MAX_TYPES];
    int32_t offset[MM_MAX_TYPES];
    double kT[MM_MAX_TYPES];
    SState st[MM_MAX_STATES];
};

// Scalars every structured kernel reads from device memory.  Written by the MD scalar kernel (or by the host for a
// plain compute()).  Positions are stored in the frame of their last write; pending cell rotations of the barostat
// and pending velocity scalings / rotations are applied on load:
//     x_true = (x_stored + periodic shifts built from rv) . Rpend         v_true = v_stored . Mvel
struct StepConsts {
    double Rpend[9];
    double Mvel[9];
    double rv[9];   // domain vectors consistent with the STORED positions
    double dt;
    double pad[3];
};

// SoA planes in z-major order with one halo plane below and above the nzl owned planes:
//   index(k, l, p) = (p * ny + l) * nx + k,   p = 0 (lower halo), 1..nzl (owned), nzl + 1 (upper halo)
// so that a z-slab of a multi-GPU decomposition is one contiguous range and its halos are whole planes.
struct SGrid {
    int active = 0;
    int nx = 0, ny = 0, nzl = 0;
    int64_t plane = 0, npad = 0;
    double *x[2][3] = {{nullptr}};  // ping-pong: kernels that move nodes read one set and write the other
    double *v[2][3] = {{nullptr}};
    double *g[2][3] = {{nullptr}};
    double *m = nullptr, *minv = nullptr;
    uint8_t *type = nullptr;
    int cx = 0, cv = 0, cg = 0;     // which copy is current
    StepConsts *d_sc = nullptr;
    SParams *d_sp = nullptr;        // device copy of sp
    StepConsts *h_sc = nullptr;     // pinned staging
    double *d_partials = nullptr;   // [nblocks][kRedSlots]
    int nblocks = 0, nblocks_alloc = 0;
    int tile_rows = 8;              // TY of the marching kernel (warps per block)
    int variant = 0;                // tuning bits of the marching kernel (k_march VAR)
    int pf_dist = 3;                // L2 prefetch distance in planes
    int chunk = 32;                 // owned planes per block along z
    SParams sp;
};

struct MarchArgs {
    int nx, ny, nzl, chunk;
    int pf_dist;  // L2 prefetch distance in planes (k_march VAR & 8)
    const double *x[3];
    double *xo[3];
    const double *v[3];
    double *vo[3];
    const double *g[3];
    double *go[3];
    const double *m, *minv;
    const uint8_t *type;
    const StepConsts *sc;
    double *partials;
};

}  // namespace mm
