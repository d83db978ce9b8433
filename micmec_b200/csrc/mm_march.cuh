// mm_march.cuh - the marching structured-grid kernel (see mm_structured.cu for the design notes)
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "mm_reduce.cuh"
#include "mm_structured.cuh"

namespace mm {

constexpr int TX = 32;  // lanes along x: one warp per tile row

// ---------------------------------------------------------------------------------------------------------------
// One state of one cell from Hs = 4 H (rows = summed edge vectors).  The reference chain (nanocell_original.py:69-132)
//     G = h0^-1 H,  eps = 1/2 (G G^T - I),  s = sym(C:eps),  E = 1/2 V0 eps:s,  D = V0 h0^-T s G
// is evaluated through the metric of the edge vectors, c = H H^T (6 unique entries), which removes h0^-1 from the
// per-cell work:  G G^T - I = h0^-1 (c - c0) h0^-T with c0 = h0 h0^T, so with the two congruences folded into the
// elasticity block on the host (fold_sparams),
//     d  = Hs Hs^T - 16 c0                      (18 FMA; the rest value enters the FMA chain first, as -1 did before)
//     Sq = Bq d                                 (36 FMA)    Bq = V0/512 K' A K,  Sq = V0/16 h0^-T s h0^-1
//     E  = 1/4 d:Sq,   D' = D/4 = Sq Hs         (7 + 27)    vir = D'^T Hs  (18, includes V0)
// 88 FP64 instructions per cell-state instead of 142, and a multi-state cell mixes the 6 entries of Sq, not D.
__device__ __forceinline__ void smetric(const double Hs[9], double c[6]) {  // Voigt order 00 11 22 12 02 01
    c[0] = fma(Hs[2], Hs[2], fma(Hs[1], Hs[1], Hs[0] * Hs[0]));
    c[1] = fma(Hs[5], Hs[5], fma(Hs[4], Hs[4], Hs[3] * Hs[3]));
    c[2] = fma(Hs[8], Hs[8], fma(Hs[7], Hs[7], Hs[6] * Hs[6]));
    c[3] = fma(Hs[5], Hs[8], fma(Hs[4], Hs[7], Hs[3] * Hs[6]));
    c[4] = fma(Hs[2], Hs[8], fma(Hs[1], Hs[7], Hs[0] * Hs[6]));
    c[5] = fma(Hs[2], Hs[5], fma(Hs[1], Hs[4], Hs[0] * Hs[3]));
}

// d = c - c0 of one state -> Sq and the elastic energy (without efree)
__device__ __forceinline__ void sstate_eval(const double d[6], const SState &P, double &e, double Sq[6]) {
#pragma unroll
    for (int I = 0; I < 6; I++) {
        // two chains of three: halves the dependent-FMA depth of the longest chain in the plane loop
        const double lo = fma(P.Bq[I * 6 + 2], d[2], fma(P.Bq[I * 6 + 1], d[1], P.Bq[I * 6] * d[0]));
        Sq[I] = fma(P.Bq[I * 6 + 5], d[5], fma(P.Bq[I * 6 + 4], d[4], fma(P.Bq[I * 6 + 3], d[3], lo)));
    }
    const double dens = fma(2.0, fma(d[5], Sq[5], fma(d[4], Sq[4], d[3] * Sq[3])), fma(d[2], Sq[2], fma(d[1], Sq[1], d[0] * Sq[0])));
    e = 0.25 * dens;
}

// named barriers (PTX bar.arrive / bar.sync with a thread count): the producer-consumer handshake between two warps
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_wait(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// all states of a cell, Boltzmann-mixed (mmff.py:377-398); the mixing is linear in the gradient, hence in Sq.
template <bool SINGLE, bool WANT_VIR>
__device__ __forceinline__ void scell_eval(const double Hs[9], const SParams &kp, int type, double &e, double D[9],
                                           double vir[6]) {
    double Sq[6];
    if (SINGLE) {
        const SState &P = kp.st[0];
        double d[6];
        d[0] = fma(Hs[2], Hs[2], fma(Hs[1], Hs[1], fma(Hs[0], Hs[0], -P.c0[0])));
        d[1] = fma(Hs[5], Hs[5], fma(Hs[4], Hs[4], fma(Hs[3], Hs[3], -P.c0[1])));
        d[2] = fma(Hs[8], Hs[8], fma(Hs[7], Hs[7], fma(Hs[6], Hs[6], -P.c0[2])));
        d[3] = fma(Hs[5], Hs[8], fma(Hs[4], Hs[7], fma(Hs[3], Hs[6], -P.c0[3])));
        d[4] = fma(Hs[2], Hs[8], fma(Hs[1], Hs[7], fma(Hs[0], Hs[6], -P.c0[4])));
        d[5] = fma(Hs[2], Hs[5], fma(Hs[1], Hs[4], fma(Hs[0], Hs[3], -P.c0[5])));
        sstate_eval(d, P, e, Sq);
        e += P.efree;
    } else {
        const int ns = kp.nstates[type], off = kp.offset[type];
        double c[6], d[6];
        smetric(Hs, c);
#pragma unroll
        for (int k = 0; k < 6; k++) d[k] = c[k] - kp.st[off].c0[k];
        sstate_eval(d, kp.st[off], e, Sq);
        e += kp.st[off].efree;
        if (ns > 1) {
            const double kT = kp.kT[type];
            double emin = e, wsum = 1.0;
#pragma unroll 1
            for (int s = 1; s < ns; s++) {
                double es, Ss[6];
#pragma unroll
                for (int k = 0; k < 6; k++) d[k] = c[k] - kp.st[off + s].c0[k];
                sstate_eval(d, kp.st[off + s], es, Ss);
                es += kp.st[off + s].efree;
                double fo, fn;  // factors for the old accumulation and for the new state
                if (es < emin) {
                    fo = exp(-(emin - es) / kT);
                    fn = 1.0;
                    emin = es;
                } else {
                    fo = 1.0;
                    fn = exp(-(es - emin) / kT);
                }
                wsum = fma(wsum, fo, fn);
#pragma unroll
                for (int k = 0; k < 6; k++) Sq[k] = fma(Sq[k], fo, fn * Ss[k]);
            }
            const double inv = 1.0 / wsum;
#pragma unroll
            for (int k = 0; k < 6; k++) Sq[k] *= inv;
            e = emin - kT * log(wsum);
        }
    }
    // D' = Sq Hs with Sq = [[0 5 4], [5 1 3], [4 3 2]]
#pragma unroll
    for (int j = 0; j < 3; j++) {
        D[j] = fma(Sq[4], Hs[6 + j], fma(Sq[5], Hs[3 + j], Sq[0] * Hs[j]));
        D[3 + j] = fma(Sq[3], Hs[6 + j], fma(Sq[1], Hs[3 + j], Sq[5] * Hs[j]));
        D[6 + j] = fma(Sq[2], Hs[6 + j], fma(Sq[3], Hs[3 + j], Sq[4] * Hs[j]));
    }
    if (!WANT_VIR) return;
    // cell virial sum_v g_v (x) r_v = D^T H = D'^T Hs (mmff.py:320-323; symmetric because Sq is)
    vir[0] = fma(D[6], Hs[6], fma(D[3], Hs[3], D[0] * Hs[0]));
    vir[1] = fma(D[7], Hs[7], fma(D[4], Hs[4], D[1] * Hs[1]));
    vir[2] = fma(D[8], Hs[8], fma(D[5], Hs[5], D[2] * Hs[2]));
    vir[3] = fma(D[7], Hs[8], fma(D[4], Hs[5], D[1] * Hs[2]));
    vir[4] = fma(D[6], Hs[8], fma(D[3], Hs[5], D[0] * Hs[2]));
    vir[5] = fma(D[6], Hs[7], fma(D[3], Hs[4], D[0] * Hs[1]));
}

// Template parameters
//   STEP    0 force only, 1 fused kick-drift-force-kick
//   SINGLE  one cell type with one metastable state (no type lookups, virial scaled once per block)
//   ROT     FORCE only: 1 = positions are rotated on load (x_true = (x + shift) . Rpend); 2 = and written back
//   VM      STEP only: pending velocity transform  0 none, 1 scalar (Mvel[0]), 2 full 3x3
//   LEAN    no virial, kinetic-energy diagonal only (NVE / NVT steps whose pressure nobody looks at)
//   PSYNC   neighbouring rows synchronise pairwise through named barriers instead of two block-wide barriers per plane
//   TY      tile rows (warps per block); the tile owns (TX-2) x (TY-2) node columns
template <int STEP, bool SINGLE, int ROT, int VM, bool LEAN, bool PSYNC, int TY>
__global__ void __launch_bounds__(TX *TY, 1)
k_march(const __grid_constant__ SParams kp, const __grid_constant__ MarchArgs a, const int write_g) {
    constexpr int OX = TX - 2, OY = TY - 2;
    __shared__ double sf[6][TY][TX];  // forward exchange along y: (px, dx) of the row above
    __shared__ double sb[9][TY][TX];  // backward exchange along y: z-combined D rows of the row below

    const int lane = threadIdx.x, row = threadIdx.y;
    const int nx = a.nx, ny = a.ny;
    const int k = blockIdx.x * OX + lane - 1, l = blockIdx.y * OY + row - 1;
    // periodic images along x and y (floor division also handles grids narrower than a tile)
    const int qx = (k >= 0) ? k / nx : -((-k + nx - 1) / nx);
    const int qy = (l >= 0) ? l / ny : -((-l + ny - 1) / ny);
    const int kk = k - qx * nx, ll = l - qy * ny;
    const bool own_xy = lane >= 1 && lane <= OX && row >= 1 && row <= OY && k < nx && l < ny;
    const StepConsts &sc = *a.sc;
    const double shx = qx * sc.rv[0] + qy * sc.rv[3];
    const double shy = qx * sc.rv[1] + qy * sc.rv[4];
    const double shz = qx * sc.rv[2] + qy * sc.rv[5];
    double R[9], M[9];
    if (ROT) {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = sc.Rpend[i];
    }
    if (STEP && VM) {
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = sc.Mvel[i];
    }
    const double dt = sc.dt, hdt = 0.5 * sc.dt;
    const int rowp = (row + 1 < TY) ? row + 1 : row;
    const int rowm = (row > 0) ? row - 1 : row;

    const int64_t plane = (int64_t)nx * ny;
    const int c0 = 1 + blockIdx.z * a.chunk;
    const int c1 = min(c0 + a.chunk, a.nzl + 1);
    int64_t idx = ((int64_t)(c0 - 1) * ny + ll) * nx + kk;  // node (kk, ll) in array plane p

    double acc[14];
#pragma unroll
    for (int i = 0; i < 14; i++) acc[i] = 0.0;

    // carried from plane to plane
    double fpxy[3], fdxy[3], fpyd[3];   // forward: xy-combined sums / differences of the previous plane
    double Dp[9];                       // D' of the previous cell layer
    double vh[3] = {0, 0, 0}, mprev = 0.0, hminv_prev = 0.0;  // STEP: half-kicked velocity / mass of the previous plane

    // software pipeline: raw loads of the NEXT plane are issued before the arithmetic of the current one
    double nx_[3], nv_[3], ng_[3], nm_ = 0.0, nminv_ = 0.0;
    auto issue_loads = [&](int64_t at) {
#pragma unroll
        for (int d = 0; d < 3; d++) nx_[d] = a.x[d][at];
        if (STEP) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                nv_[d] = a.v[d][at];
                ng_[d] = a.g[d][at];
            }
            nm_ = a.m[at];
            nminv_ = a.minv[at];
        }
    };
    issue_loads(idx);

#pragma unroll 1
    for (int p = c0 - 1; p <= c1; p++, idx += plane) {
        const double cx0 = nx_[0], cx1 = nx_[1], cx2 = nx_[2];
        double cv[3], cg[3];
        const double cm = nm_, cminv = nminv_;
        if (STEP) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                cv[d] = nv_[d];
                cg[d] = ng_[d];
            }
        }
        if (p < c1) issue_loads(idx + plane);

        // ---- node (lane, row, p): true position (and, in STEP mode, kick + drift: verlet.py:144-146) -------------
        const double xs = cx0 + shx, ys = cx1 + shy, zs = cx2 + shz;
        double r[3];
        if (ROT) {
#pragma unroll
            for (int j = 0; j < 3; j++) r[j] = fma(zs, R[6 + j], fma(ys, R[3 + j], xs * R[j]));
        } else {
            r[0] = xs;
            r[1] = ys;
            r[2] = zs;
        }
        double vcur[3] = {0, 0, 0};
        const double hminv = hdt * cminv;
        if (STEP) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                double vt;
                if (VM == 2) vt = fma(cv[2], M[6 + j], fma(cv[1], M[3 + j], cv[0] * M[j]));
                else if (VM == 1) vt = cv[j] * M[0];
                else vt = cv[j];
                vcur[j] = fma(-hminv, cg[j], vt);
                r[j] = fma(dt, vcur[j], r[j]);
            }
        }
        if ((STEP || ROT == 2) && own_xy && p >= c0 && p < c1) {  // owned columns carry no periodic shift
#pragma unroll
            for (int j = 0; j < 3; j++) a.xo[j][idx] = r[j];
        }

        // ---- forward butterfly: x by shuffle, y through shared memory, z in registers ------------------------------
        double px[3], dx[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double rn = __shfl_down_sync(0xffffffffu, r[j], 1);
            px[j] = rn + r[j];
            dx[j] = rn - r[j];
            sf[j][row][lane] = px[j];
            sf[3 + j][row][lane] = dx[j];
        }
        if (PSYNC) {  // row r only needs row r+1: producer arrives, consumer waits (barrier ids 1 .. TY-1)
            if (row > 0) bar_arrive(row, 2 * TX);
            if (row + 1 < TY) bar_wait(row + 1, 2 * TX);
        } else {
            __syncthreads();
        }
        double pxy[3], dxy[3], pyd[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double pxn = sf[j][rowp][lane], dxn = sf[3 + j][rowp][lane];
            pxy[j] = px[j] + pxn;
            dxy[j] = dx[j] + dxn;
            pyd[j] = pxn - px[j];
        }

        const bool have_cell = p >= c0;       // cell layer p-1 (planes p-1 and p)
        const bool have_node = p >= c0 + 1;   // node plane p-1 (cell layers p-2 and p-1)
        double D[9];
        if (have_cell) {
            double Hs[9];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Hs[j] = fdxy[j] + dxy[j];        // 4 * (mean x edge)
                Hs[3 + j] = fpyd[j] + pyd[j];    // 4 * (mean y edge)
                Hs[6 + j] = pxy[j] - fpxy[j];    // 4 * (mean z edge)
            }
            const int type = SINGLE ? 0 : (int)a.type[idx - plane];
            double e, vir[6];
            scell_eval<SINGLE, !LEAN>(Hs, kp, type, e, D, vir);
            if (own_xy && have_node) {  // the warm-up layer c0-1 belongs to the chunk below
                acc[0] += e;
                if (!LEAN) {
#pragma unroll
                    for (int q = 0; q < 6; q++) acc[1 + q] += vir[q];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            fpxy[j] = pxy[j];
            fdxy[j] = dxy[j];
            fpyd[j] = pyd[j];
        }

        // ---- backward butterfly: z in registers, y through shared memory, x by shuffle ---------------------------
        double P[9];  // z-combined rows of this thread's cell column: kept in registers across the barrier
        if (have_node) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                P[j] = Dp[j] + D[j];
                P[3 + j] = Dp[3 + j] + D[3 + j];
                P[6 + j] = Dp[6 + j] - D[6 + j];
                sb[j][row][lane] = P[j];
                sb[3 + j][row][lane] = P[3 + j];
                sb[6 + j][row][lane] = P[6 + j];
            }
        }
        if (PSYNC) {  // row r only needs row r-1 (barrier ids TY .. 2 TY - 2); these two handshakes per plane also order
                      // the reuse of sf / sb between planes
            if (row + 1 < TY) bar_arrive(TY + row, 2 * TX);
            if (row > 0) bar_wait(TY + row - 1, 2 * TX);
        } else {
            __syncthreads();
        }
        if (have_node) {
            double g[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double q0 = sb[j][rowm][lane] + P[j];
                const double q1 = sb[3 + j][rowm][lane] - P[3 + j];
                const double q2 = sb[6 + j][rowm][lane] + P[6 + j];
                const double s12 = q1 + q2;
                const double q0m = __shfl_up_sync(0xffffffffu, q0, 1), s12m = __shfl_up_sync(0xffffffffu, s12, 1);
                g[j] = (q0m - q0) + (s12m + s12);
            }
            if (own_xy) {
                const int64_t at = idx - plane;
                if (STEP) {  // second kick (verlet.py:152-153) + kinetic moments of the new velocities
                    double vn[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        vn[j] = fma(-hminv_prev, g[j], vh[j]);
                        a.vo[j][at] = vn[j];
                    }
                    const double mx = mprev * vn[0], my = mprev * vn[1], mz = mprev * vn[2];
                    acc[7] = fma(mx, vn[0], acc[7]);
                    acc[8] = fma(my, vn[1], acc[8]);
                    acc[9] = fma(mz, vn[2], acc[9]);
                    if (!LEAN) {
                        acc[10] = fma(my, vn[2], acc[10]);
                        acc[11] = fma(mx, vn[2], acc[11]);
                        acc[12] = fma(mx, vn[1], acc[12]);
                    }
                }
                if (write_g) {
#pragma unroll
                    for (int j = 0; j < 3; j++) a.go[j][at] = g[j];
                }
                if (!LEAN) acc[13] += fma(g[0], g[0], fma(g[1], g[1], g[2] * g[2]));
            }
        }
        if (have_cell) {
#pragma unroll
            for (int q = 0; q < 9; q++) Dp[q] = D[q];
        }
        if (STEP) {
#pragma unroll
            for (int j = 0; j < 3; j++) vh[j] = vcur[j];
            mprev = cm;
            hminv_prev = hminv;
        }
    }

    // block reduction: warp shuffles, then one warp over the per-warp sums
    __shared__ double red[TY][14];
#pragma unroll
    for (int q = 0; q < 14; q++) {
        const double s = warp_sum(acc[q]);
        if (lane == 0) red[row][q] = s;
    }
    __syncthreads();
    if (row == 0 && lane < 14) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < TY; w++) s += red[w][lane];
        const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        a.partials[(size_t)bid * kRedSlots + lane] = s;
    }
}

}  // namespace mm
