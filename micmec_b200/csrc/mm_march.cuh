// mm_march.cuh - the marching structured-grid kernel (see mm_structured.cu for the design notes)
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "mm_reduce.cuh"
#include "mm_structured.cuh"

namespace mm {

constexpr int TX = 32;  // lanes along x: one warp per tile row

// ---------------------------------------------------------------------------------------------------------------
// one state of one cell from Hs = 4 H (rows = summed edge vectors); constants pre-scaled (SState)
// Ahg: the state's folded elasticity (36 doubles) in GLOBAL memory.  FP64 instructions on sm_100 take constants only
// from (uniform) registers; 56 constant doubles per state overflow the uniform register file and ptxas then spills
// uniform registers through vector registers inside the plane loop (~85 instructions per plane).  Reading the 6x6
// block with explicit, non-hoistable 16-byte read-only loads (L1 broadcast) keeps it out of the register files.
__device__ __forceinline__ void ld2(const double *p, double &a, double &b) {
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
}

template <bool AHG, bool WANT_VIR>
__device__ __forceinline__ void sstate_eval(const double Hs[9], const SState &P, const double *__restrict__ Ahg, double &e,
                                            double D[9], double vir[6]) {
    double G[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            G[i * 3 + j] = fma(P.hiq[i * 3 + 2], Hs[6 + j], fma(P.hiq[i * 3 + 1], Hs[3 + j], P.hiq[i * 3] * Hs[j]));
    double u[6];  // u = G G^T - I = 2 eps
    u[0] = fma(G[2], G[2], fma(G[1], G[1], fma(G[0], G[0], -1.0)));
    u[1] = fma(G[5], G[5], fma(G[4], G[4], fma(G[3], G[3], -1.0)));
    u[2] = fma(G[8], G[8], fma(G[7], G[7], fma(G[6], G[6], -1.0)));
    u[3] = fma(G[5], G[8], fma(G[4], G[7], G[3] * G[6]));
    u[4] = fma(G[2], G[8], fma(G[1], G[7], G[0] * G[6]));
    u[5] = fma(G[2], G[5], fma(G[1], G[4], G[0] * G[3]));
    double s[6];
#pragma unroll
    for (int I = 0; I < 6; I++) {
        if (AHG) {
            double a0, a1, a2, a3, a4, a5;
            ld2(Ahg + I * 6, a0, a1);
            ld2(Ahg + I * 6 + 2, a2, a3);
            ld2(Ahg + I * 6 + 4, a4, a5);
            s[I] = fma(a5, u[5], fma(a4, u[4], fma(a3, u[3], fma(a2, u[2], fma(a1, u[1], a0 * u[0])))));
        } else {
            double acc = P.Ah[I * 6] * u[0];
#pragma unroll
            for (int J = 1; J < 6; J++) acc = fma(P.Ah[I * 6 + J], u[J], acc);
            s[I] = acc;
        }
    }
    const double dens = fma(2.0, fma(u[5], s[5], fma(u[4], s[4], u[3] * s[3])), fma(u[2], s[2], fma(u[1], s[1], u[0] * s[0])));
    e = P.v0q * dens;
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
            D[i * 3 + j] = fma(P.hitv[i * 3 + 2], T[6 + j], fma(P.hitv[i * 3 + 1], T[3 + j], P.hitv[i * 3] * T[j]));
    if (!WANT_VIR) return;
    // G^T T (symmetric); the V0 factor is applied by the caller
    vir[0] = fma(G[6], T[6], fma(G[3], T[3], G[0] * T[0]));
    vir[1] = fma(G[7], T[7], fma(G[4], T[4], G[1] * T[1]));
    vir[2] = fma(G[8], T[8], fma(G[5], T[5], G[2] * T[2]));
    vir[3] = fma(G[7], T[8], fma(G[4], T[5], G[1] * T[2]));
    vir[4] = fma(G[6], T[8], fma(G[3], T[5], G[0] * T[2]));
    vir[5] = fma(G[6], T[7], fma(G[3], T[4], G[0] * T[1]));
}

// all states of a cell, Boltzmann-mixed (mmff.py:377-398); the mixing is linear in the gradient, hence in D.
// SINGLE (one type, one state): vir is returned WITHOUT the V0 factor (applied once per block by the caller).
template <bool SINGLE, bool AHG, bool WANT_VIR>
__device__ __forceinline__ void scell_eval(const double Hs[9], const SParams &kp, const SParams *__restrict__ kpg, int type,
                                           double &e, double D[9], double vir[6]) {
    if (SINGLE) {
        sstate_eval<AHG, WANT_VIR>(Hs, kp.st[0], kpg->st[0].Ah, e, D, vir);
        e += kp.st[0].efree;
        return;
    }
    const int ns = kp.nstates[type], off = kp.offset[type];
    sstate_eval<AHG, true>(Hs, kp.st[off], kpg->st[off].Ah, e, D, vir);
    e += kp.st[off].efree;
#pragma unroll
    for (int k = 0; k < 6; k++) vir[k] *= kp.st[off].v0;
    if (ns == 1) return;
    const double kT = kp.kT[type];
    double emin = e, wsum = 1.0;
#pragma unroll 1
    for (int s = 1; s < ns; s++) {
        double es, Ds[9], vs[6];
        sstate_eval<AHG, true>(Hs, kp.st[off + s], kpg->st[off + s].Ah, es, Ds, vs);
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
        const double fv = fn * kp.st[off + s].v0;
#pragma unroll
        for (int k = 0; k < 9; k++) D[k] = fma(D[k], fo, fn * Ds[k]);
#pragma unroll
        for (int k = 0; k < 6; k++) vir[k] = fma(vir[k], fo, fv * vs[k]);
    }
    const double inv = 1.0 / wsum;
#pragma unroll
    for (int k = 0; k < 9; k++) D[k] *= inv;
#pragma unroll
    for (int k = 0; k < 6; k++) vir[k] *= inv;
    e = emin - kT * log(wsum);
}

// Template parameters
//   STEP    0 force only, 1 fused kick-drift-force-kick
//   SINGLE  one cell type with one metastable state (no type lookups, virial scaled once per block)
//   ROT     FORCE only: 1 = positions are rotated on load (x_true = (x + shift) . Rpend); 2 = and written back
//   VM      STEP only: pending velocity transform  0 none, 1 scalar (Mvel[0]), 2 full 3x3
//   LEAN    no virial, kinetic-energy diagonal only (NVE / NVT steps whose pressure nobody looks at)
//   AHG     elasticity block read from global memory inside the loop instead of uniform registers
//   TY      tile rows (warps per block); the tile owns (TX-2) x (TY-2) node columns
template <int STEP, bool SINGLE, int ROT, int VM, bool LEAN, bool AHG, int TY>
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
        __syncthreads();
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
            scell_eval<SINGLE, AHG, !LEAN>(Hs, kp, a.spd, type, e, D, vir);
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
        __syncthreads();
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

    if (SINGLE && !LEAN) {  // the V0 factor of the virial is common to every cell of the block
#pragma unroll
        for (int q = 0; q < 6; q++) acc[1 + q] *= kp.st[0].v0;
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
