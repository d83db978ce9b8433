// mm_structured.cu - the performance path: fused force (+ Verlet) kernel for full periodic grids.
//
// Same arithmetic as deformation()/_compute_gpos/_compute_vtens (micmec/pes/mmff.py:288-403) with the `original`
// per-cell model (micmec/pes/nanocell_original.py) and as VerletIntegrator.propagate (micmec/sampling/verlet.py:144-154),
// restructured for the machine:
//
//  * nodes live in SoA planes, z-major, x fastest (mm_structured.cuh) - every global access of a warp is one
//    contiguous 256-byte row segment;
//  * a thread block owns an (x, y) tile and MARCHES along z.  Thread (lane, row) is the node column (x0+lane, y0+row)
//    and, at the same time, the cell column whose origin vertex is that node;
//  * the +-1 stencils of the reference (multiplicator / cell_derivs, micmec/pes/nanocell_utils.py) are separable, so
//    both the 8-vertex -> edge-matrix reduction and the 8-cell -> node gradient gather are done as three 1-D
//    butterflies: along x with warp shuffles, along y through shared memory, along z in registers carried from one
//    plane to the next.  24 + 30 additions per cell instead of 51 + 72, and no per-cell gradient ever reaches HBM;
//  * STEP mode fuses kick-drift-force-kick: positions, velocities and old gradients of the tile (and its one-node
//    apron) are loaded once, the drift is applied on the fly, forces are evaluated at the new positions, the second
//    kick is applied and x, v (and optionally g) are written once.  Pending barostat rotations and thermostat
//    scalings (StepConsts) are applied on load, so there is no separate "scale velocities" / "rotate positions" pass;
//  * energy, virial (6) and the kinetic second moments (6) are reduced with warp shuffles into one partial per block.
//
// Apron cells are recomputed by neighbouring blocks (tile 32 x 8 threads -> 30 x 6 owned nodes): the price of never
// materialising per-cell data.  DESIGN.md discusses the trade-off and the measured numbers.
#include <vector>

#include "mm_internal.h"
#include "mm_reduce.cuh"

namespace mm {

constexpr int TX = 32, TY = 8;
constexpr int OX = TX - 2, OY = TY - 2;

// ---------------------------------------------------------------------------------------------------------------
// one state of one cell from Hs = 4 H (rows = summed edge vectors)
__device__ __forceinline__ void sstate_eval(const double Hs[9], const SState &P, double &e, double D[9], double vir[6]) {
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
        double acc = P.Ah[I * 6] * u[0];
#pragma unroll
        for (int J = 1; J < 6; J++) acc = fma(P.Ah[I * 6 + J], u[J], acc);
        s[I] = acc;
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
    vir[0] = P.v0 * fma(G[6], T[6], fma(G[3], T[3], G[0] * T[0]));
    vir[1] = P.v0 * fma(G[7], T[7], fma(G[4], T[4], G[1] * T[1]));
    vir[2] = P.v0 * fma(G[8], T[8], fma(G[5], T[5], G[2] * T[2]));
    vir[3] = P.v0 * fma(G[7], T[8], fma(G[4], T[5], G[1] * T[2]));
    vir[4] = P.v0 * fma(G[6], T[8], fma(G[3], T[5], G[0] * T[2]));
    vir[5] = P.v0 * fma(G[6], T[7], fma(G[3], T[4], G[0] * T[1]));
}

// all states of a cell, Boltzmann-mixed (mmff.py:377-398); the mixing is linear in the gradient, hence in D
template <bool SINGLE>
__device__ __forceinline__ void scell_eval(const double Hs[9], const SParams &kp, int type, double &e, double D[9], double vir[6]) {
    if (SINGLE) {
        sstate_eval(Hs, kp.st[0], e, D, vir);
        e += kp.st[0].efree;
        return;
    }
    const int ns = kp.nstates[type], off = kp.offset[type];
    sstate_eval(Hs, kp.st[off], e, D, vir);
    e += kp.st[off].efree;
    if (ns == 1) return;
    const double kT = kp.kT[type];
    double emin = e, wsum = 1.0;
#pragma unroll 1
    for (int s = 1; s < ns; s++) {
        double es, Ds[9], vs[6];
        sstate_eval(Hs, kp.st[off + s], es, Ds, vs);
        es += kp.st[off + s].efree;
        double fo, fn;  // factors for the old accumulation and the new state
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
        for (int k = 0; k < 9; k++) D[k] = fma(D[k], fo, fn * Ds[k]);
#pragma unroll
        for (int k = 0; k < 6; k++) vir[k] = fma(vir[k], fo, fn * vs[k]);
    }
    const double inv = 1.0 / wsum;
#pragma unroll
    for (int k = 0; k < 9; k++) D[k] *= inv;
#pragma unroll
    for (int k = 0; k < 6; k++) vir[k] *= inv;
    e = emin - kT * log(wsum);
}

__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }

// ---------------------------------------------------------------------------------------------------------------
template <bool STEP, bool WRITE_G, bool SINGLE>
__global__ void __launch_bounds__(TX *TY)
k_march(const __grid_constant__ SParams kp, const __grid_constant__ MarchArgs a) {
    __shared__ double sf[6][TY][TX];  // forward exchange along y: (px, dx) of the row above
    __shared__ double sb[9][TY][TX];  // backward exchange along y: z-combined D rows of the row below

    const int lane = threadIdx.x, row = threadIdx.y;
    const int nx = a.nx, ny = a.ny;
    const int k = blockIdx.x * OX + lane - 1, l = blockIdx.y * OY + row - 1;
    // periodic images along x and y (floor division handles grids narrower than a tile)
    const int qx = (k >= 0) ? k / nx : -((-k + nx - 1) / nx);
    const int qy = (l >= 0) ? l / ny : -((-l + ny - 1) / ny);
    const int kk = k - qx * nx, ll = l - qy * ny;
    const bool own_xy = lane >= 1 && lane <= OX && row >= 1 && row <= OY && k < nx && l < ny;
    const StepConsts &sc = *a.sc;
    const double shx = qx * sc.rv[0] + qy * sc.rv[3];
    const double shy = qx * sc.rv[1] + qy * sc.rv[4];
    const double shz = qx * sc.rv[2] + qy * sc.rv[5];
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = sc.Rpend[i];
    const double dt = sc.dt, hdt = 0.5 * sc.dt;

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

    for (int p = c0 - 1; p <= c1; p++, idx += plane) {
        double cx[3] = {nx_[0], nx_[1], nx_[2]};
        double cv[3], cg[3], cm = nm_, cminv = nminv_;
        if (STEP) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                cv[d] = nv_[d];
                cg[d] = ng_[d];
            }
        }
        if (p < c1) issue_loads(idx + plane);

        // ---- node (lane, row, p): true position (and, in STEP mode, kick + drift: verlet.py:144-146) -------------
        const double xs = cx[0] + shx, ys = cx[1] + shy, zs = cx[2] + shz;
        double r[3];
#pragma unroll
        for (int j = 0; j < 3; j++) r[j] = fma(zs, R[6 + j], fma(ys, R[3 + j], xs * R[j]));
        double vcur[3] = {0, 0, 0};
        const double hminv = hdt * cminv;
        if (STEP) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double vt = fma(cv[2], sc.Mvel[6 + j], fma(cv[1], sc.Mvel[3 + j], cv[0] * sc.Mvel[j]));
                vcur[j] = fma(-hminv, cg[j], vt);
                r[j] = fma(dt, vcur[j], r[j]);
            }
            if (own_xy && p >= c0 && p < c1) {
#pragma unroll
                for (int j = 0; j < 3; j++) a.xo[j][idx] = r[j];
            }
        }

        // ---- forward butterfly: x by shuffle, y through shared memory, z in registers ------------------------------
        double px[3], dx[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double rn = shfl_down1(r[j]);
            px[j] = rn + r[j];
            dx[j] = rn - r[j];
            sf[j][row][lane] = px[j];
            sf[3 + j][row][lane] = dx[j];
        }
        __syncthreads();
        const int rowp = (row + 1 < TY) ? row + 1 : row;
        double pxy[3], dxy[3], pyd[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double pxn = sf[j][rowp][lane], dxn = sf[3 + j][rowp][lane];
            pxy[j] = px[j] + pxn;
            dxy[j] = dx[j] + dxn;
            pyd[j] = pxn - px[j];
        }

        double g[3] = {0, 0, 0};
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
            scell_eval<SINGLE>(Hs, kp, type, e, D, vir);
            if (own_xy && have_node) {  // the warm-up layer c0-1 belongs to the chunk below
                acc[0] += e;
#pragma unroll
                for (int q = 0; q < 6; q++) acc[1 + q] += vir[q];
            }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            fpxy[j] = pxy[j];
            fdxy[j] = dxy[j];
            fpyd[j] = pyd[j];
        }

        // ---- backward butterfly: z in registers, y through shared memory, x by shuffle ---------------------------
        if (have_node) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                sb[j][row][lane] = Dp[j] + D[j];
                sb[3 + j][row][lane] = Dp[3 + j] + D[3 + j];
                sb[6 + j][row][lane] = Dp[6 + j] - D[6 + j];
            }
        }
        __syncthreads();
        if (have_node) {
            const int rowm = (row > 0) ? row - 1 : row;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double q0 = sb[j][rowm][lane] + sb[j][row][lane];
                const double q1 = sb[3 + j][rowm][lane] - sb[3 + j][row][lane];
                const double q2 = sb[6 + j][rowm][lane] + sb[6 + j][row][lane];
                const double s12 = q1 + q2;
                const double q0m = shfl_up1(q0), s12m = shfl_up1(s12);
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
                    acc[7] = fma(mprev * vn[0], vn[0], acc[7]);
                    acc[8] = fma(mprev * vn[1], vn[1], acc[8]);
                    acc[9] = fma(mprev * vn[2], vn[2], acc[9]);
                    acc[10] = fma(mprev * vn[1], vn[2], acc[10]);
                    acc[11] = fma(mprev * vn[0], vn[2], acc[11]);
                    acc[12] = fma(mprev * vn[0], vn[1], acc[12]);
                }
                if (WRITE_G) {
#pragma unroll
                    for (int j = 0; j < 3; j++) a.go[j][at] = g[j];
                }
                acc[13] += fma(g[0], g[0], fma(g[1], g[1], g[2] * g[2]));
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
    const int bid = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    // block_sum_store expects a 1-D thread index
    {
        __shared__ double red[TY][14];
        const int warp = row;
#pragma unroll
        for (int q = 0; q < 14; q++) {
            const double s = warp_sum(acc[q]);
            if (lane == 0) red[warp][q] = s;
        }
        __syncthreads();
        if (warp == 0 && lane < 14) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < TY; w++) s += red[w][lane];
            a.partials[(size_t)bid * kRedSlots + lane] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// halo planes: p = 0 <- p = nzl (minus c), p = nzl + 1 <- p = 1 (plus c); c = third domain vector of the stored frame
struct HaloArgs {
    double *f[9];
    int nfields;
    int npos;  // the first npos fields are position components 0, 1, 2
};

__global__ void __launch_bounds__(256)
k_halo(const __grid_constant__ HaloArgs h, int64_t plane, int nzl, const StepConsts *sc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
        for (int f = 0; f < h.nfields; f++) {
            const double c = (f < h.npos) ? sc->rv[6 + f] : 0.0;
            h.f[f][i] = h.f[f][(int64_t)nzl * plane + i] - c;
            h.f[f][(int64_t)(nzl + 1) * plane + i] = h.f[f][plane + i] + c;
        }
    }
}

__global__ void __launch_bounds__(256)
k_halo_u8(uint8_t *f, int64_t plane, int nzl) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
        f[i] = f[(int64_t)nzl * plane + i];
        f[(int64_t)(nzl + 1) * plane + i] = f[plane + i];
    }
}

// reference order (AoS, id = (k*ny + l)*nz + m) <-> z-major SoA planes (owned planes only)
__global__ void __launch_bounds__(256)
k_aos_to_soa(const double *__restrict__ aos, double *s0, double *s1, double *s2, int nx, int ny, int nz) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nx), l = (int)((i / nx) % ny), m = (int)(i / ((int64_t)nx * ny));
        const int64_t src = ((int64_t)k * ny + l) * nz + m;
        const int64_t dst = i + (int64_t)nx * ny;  // skip the lower halo plane
        s0[dst] = aos[3 * src];
        s1[dst] = aos[3 * src + 1];
        s2[dst] = aos[3 * src + 2];
    }
}

__global__ void __launch_bounds__(256)
k_soa_to_aos(const double *__restrict__ s0, const double *__restrict__ s1, const double *__restrict__ s2, double *aos,
             int nx, int ny, int nz) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nx), l = (int)((i / nx) % ny), m = (int)(i / ((int64_t)nx * ny));
        const int64_t dst = ((int64_t)k * ny + l) * nz + m;
        const int64_t src = i + (int64_t)nx * ny;
        aos[3 * dst] = s0[src];
        aos[3 * dst + 1] = s1[src];
        aos[3 * dst + 2] = s2[src];
    }
}

__global__ void __launch_bounds__(256)
k_mass_to_soa(const double *__restrict__ masses, double *m, double *minv, int nx, int ny, int nz) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nx), l = (int)((i / nx) % ny), mm_ = (int)(i / ((int64_t)nx * ny));
        const double v = masses[((int64_t)k * ny + l) * nz + mm_];
        m[i + (int64_t)nx * ny] = v;
        minv[i + (int64_t)nx * ny] = 1.0 / v;
    }
}

__global__ void __launch_bounds__(256)
k_type_to_soa(const uint8_t *__restrict__ info, uint8_t *type, int nx, int ny, int nz) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nx), l = (int)((i / nx) % ny), m = (int)(i / ((int64_t)nx * ny));
        type[i + (int64_t)nx * ny] = info[((int64_t)k * ny + l) * nz + m] & 15u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
static void fold_sparams(const KParams &kp, SParams &sp) {
    sp.ntypes = kp.ntypes;
    for (int t = 0; t < MM_MAX_TYPES; t++) {
        sp.nstates[t] = kp.nstates[t];
        sp.offset[t] = kp.offset[t];
        sp.kT[t] = kp.kT[t];
    }
    for (int s = 0; s < MM_MAX_STATES; s++) {
        const StateP &P = kp.st[s];
        SState &S = sp.st[s];
        for (int i = 0; i < 9; i++) S.hiq[i] = 0.25 * P.hi[i];
        for (int i = 0; i < 36; i++) S.Ah[i] = 0.5 * P.A[i];
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 3; k++) S.hitv[i * 3 + k] = 0.25 * P.v0 * P.hi[k * 3 + i];
        S.v0 = P.v0;
        S.v0q = 0.25 * P.v0;
        S.efree = P.efree;
    }
}

bool sg_eligible(const mm_handle *h) {
    return h->structured && h->model == MM_MODEL_ORIGINAL && h->nx >= 2 && h->ny >= 2 && h->nz >= 2;
}

int sg_blocks(const mm_handle *h, dim3 &grid) {
    const SGrid &g = h->sg;
    grid = dim3((g.nx + OX - 1) / OX, (g.ny + OY - 1) / OY, (g.nzl + g.chunk - 1) / g.chunk);
    return (int)(grid.x * grid.y * grid.z);
}

int sg_setup(mm_handle *h) {
    SGrid &g = h->sg;
    g.nx = h->nx;
    g.ny = h->ny;
    g.nzl = h->nz;
    // The reference enumerates id = (k*ny + l)*nz + m with k along the FIRST domain vector.  Device planes are indexed
    // (x = k fastest, y = l, z = m slowest): the marching direction is the reference's third axis, the lanes run along
    // its first axis.
    g.plane = (int64_t)g.nx * g.ny;
    g.npad = g.plane * (g.nzl + 2);
    fold_sparams(h->kp, g.sp);
    const size_t bytes = sizeof(double) * g.npad;
    for (int c = 0; c < 2; c++)
        for (int d = 0; d < 3; d++) {
            MM_CUDA(cudaMalloc(&g.x[c][d], bytes));
            MM_CUDA(cudaMalloc(&g.v[c][d], bytes));
            MM_CUDA(cudaMalloc(&g.g[c][d], bytes));
            MM_CUDA(cudaMemsetAsync(g.x[c][d], 0, bytes, h->stream));
            MM_CUDA(cudaMemsetAsync(g.v[c][d], 0, bytes, h->stream));
            MM_CUDA(cudaMemsetAsync(g.g[c][d], 0, bytes, h->stream));
        }
    MM_CUDA(cudaMalloc(&g.m, bytes));
    MM_CUDA(cudaMalloc(&g.minv, bytes));
    MM_CUDA(cudaMemsetAsync(g.m, 0, bytes, h->stream));
    MM_CUDA(cudaMemsetAsync(g.minv, 0, bytes, h->stream));
    MM_CUDA(cudaMalloc(&g.type, g.npad));
    MM_CUDA(cudaMalloc(&g.d_sc, sizeof(StepConsts)));
    MM_CUDA(cudaHostAlloc(&g.h_sc, sizeof(StepConsts), cudaHostAllocDefault));
    // chunk length along z: enough blocks for >= 4 waves of one block per SM, but no shorter than 8 planes
    dim3 grid;
    g.chunk = 32;
    while (g.chunk > 8 && sg_blocks(h, grid) < 4 * h->num_sms) g.chunk /= 2;
    g.nblocks = sg_blocks(h, grid);
    MM_CUDA(cudaMalloc(&g.d_partials, sizeof(double) * (size_t)g.nblocks * kRedSlots));
    k_type_to_soa<<<grid_for(h, h->ncells, 256), 256, 0, h->stream>>>(h->d_cell_info, g.type, g.nx, g.ny, g.nzl);
    k_halo_u8<<<grid_for(h, g.plane, 256), 256, 0, h->stream>>>(g.type, g.plane, g.nzl);
    MM_CUDA(cudaGetLastError());
    g.active = 1;
    return MM_OK;
}

void sg_free(mm_handle *h) {
    SGrid &g = h->sg;
    for (int c = 0; c < 2; c++)
        for (int d = 0; d < 3; d++) {
            cudaFree(g.x[c][d]);
            cudaFree(g.v[c][d]);
            cudaFree(g.g[c][d]);
        }
    cudaFree(g.m);
    cudaFree(g.minv);
    cudaFree(g.type);
    cudaFree(g.d_sc);
    cudaFree(g.d_partials);
    if (g.h_sc) cudaFreeHost(g.h_sc);
    g.active = 0;
}

// host-side write of the step constants (plain compute(): identity transforms, current domain vectors)
int sg_write_consts(mm_handle *h, const double *rvecs9, double dt) {
    SGrid &g = h->sg;
    MM_CUDA(cudaStreamSynchronize(h->stream));
    StepConsts &sc = *g.h_sc;
    for (int i = 0; i < 9; i++) {
        sc.Rpend[i] = sc.Mvel[i] = (i % 4 == 0) ? 1.0 : 0.0;
        sc.rv[i] = rvecs9[i];
    }
    sc.dt = dt;
    MM_CUDA(cudaMemcpyAsync(g.d_sc, g.h_sc, sizeof(StepConsts), cudaMemcpyHostToDevice, h->stream));
    return MM_OK;
}

int sg_halo(mm_handle *h, bool pos, bool vel, bool grad) {
    SGrid &g = h->sg;
    HaloArgs ha;
    ha.nfields = 0;
    ha.npos = 0;
    if (pos) {
        for (int d = 0; d < 3; d++) ha.f[ha.nfields++] = g.x[g.cx][d];
        ha.npos = 3;
    }
    if (vel)
        for (int d = 0; d < 3; d++) ha.f[ha.nfields++] = g.v[g.cv][d];
    if (grad)
        for (int d = 0; d < 3; d++) ha.f[ha.nfields++] = g.g[g.cg][d];
    if (ha.nfields == 0) return MM_OK;
    k_halo<<<grid_for(h, g.plane, 256), 256, 0, h->stream>>>(ha, g.plane, g.nzl, g.d_sc);
    h->launches++;
    return MM_OK;
}

int sg_halo_mass(mm_handle *h) {
    SGrid &g = h->sg;
    HaloArgs ha;
    ha.nfields = 2;
    ha.npos = 0;
    ha.f[0] = g.m;
    ha.f[1] = g.minv;
    k_halo<<<grid_for(h, g.plane, 256), 256, 0, h->stream>>>(ha, g.plane, g.nzl, g.d_sc);
    h->launches++;
    return MM_OK;
}

int sg_pos_from_aos(mm_handle *h, const double *d_aos) {
    SGrid &g = h->sg;
    k_aos_to_soa<<<grid_for(h, h->nnodes, 256), 256, 0, h->stream>>>(d_aos, g.x[g.cx][0], g.x[g.cx][1], g.x[g.cx][2], g.nx, g.ny, g.nzl);
    h->launches++;
    return sg_halo(h, true, false, false);
}

int sg_vel_from_aos(mm_handle *h, const double *d_aos) {
    SGrid &g = h->sg;
    k_aos_to_soa<<<grid_for(h, h->nnodes, 256), 256, 0, h->stream>>>(d_aos, g.v[g.cv][0], g.v[g.cv][1], g.v[g.cv][2], g.nx, g.ny, g.nzl);
    h->launches++;
    return sg_halo(h, false, true, false);
}

int sg_mass_from_aos(mm_handle *h, const double *d_masses) {
    SGrid &g = h->sg;
    k_mass_to_soa<<<grid_for(h, h->nnodes, 256), 256, 0, h->stream>>>(d_masses, g.m, g.minv, g.nx, g.ny, g.nzl);
    h->launches++;
    return sg_halo_mass(h);
}

int sg_to_aos(mm_handle *h, int which, double *d_aos) {  // 0 pos, 1 vel, 2 gpos
    SGrid &g = h->sg;
    double **src = which == 0 ? g.x[g.cx] : which == 1 ? g.v[g.cv] : g.g[g.cg];
    k_soa_to_aos<<<grid_for(h, h->nnodes, 256), 256, 0, h->stream>>>(src[0], src[1], src[2], d_aos, g.nx, g.ny, g.nzl);
    h->launches++;
    return MM_OK;
}

static void fill_args(mm_handle *h, MarchArgs &a) {
    SGrid &g = h->sg;
    a.nx = g.nx;
    a.ny = g.ny;
    a.nzl = g.nzl;
    a.chunk = g.chunk;
    for (int d = 0; d < 3; d++) {
        a.x[d] = g.x[g.cx][d];
        a.xo[d] = g.x[g.cx ^ 1][d];
        a.v[d] = g.v[g.cv][d];
        a.vo[d] = g.v[g.cv ^ 1][d];
        a.g[d] = g.g[g.cg][d];
        a.go[d] = g.g[g.cg ^ 1][d];
    }
    a.m = g.m;
    a.minv = g.minv;
    a.type = g.type;
    a.sc = g.d_sc;
    a.partials = g.d_partials;
}

// Force evaluation at the stored positions (x Rpend).  write_g: store the node gradient into the CURRENT g set.
int sg_force(mm_handle *h, bool write_g) {
    SGrid &g = h->sg;
    MarchArgs a;
    fill_args(h, a);
    for (int d = 0; d < 3; d++) a.go[d] = g.g[g.cg][d];  // FORCE mode never reads g: write in place
    dim3 grid;
    sg_blocks(h, grid);
    const dim3 block(TX, TY);
    const bool single = g.sp.ntypes == 1 && g.sp.nstates[0] == 1;
    prof_begin(h);
    if (single) {
        if (write_g) k_march<false, true, true><<<grid, block, 0, h->stream>>>(g.sp, a);
        else k_march<false, false, true><<<grid, block, 0, h->stream>>>(g.sp, a);
    } else {
        if (write_g) k_march<false, true, false><<<grid, block, 0, h->stream>>>(g.sp, a);
        else k_march<false, false, false><<<grid, block, 0, h->stream>>>(g.sp, a);
    }
    prof_end(h);
    h->launches++;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// Fused kick-drift-force-kick.  Reads the current x, v, g sets, writes the other x and v sets (and g when write_g)
// and flips them.  Halo planes of what was written are refreshed afterwards by the caller (sg_halo).
int sg_step(mm_handle *h, bool write_g) {
    SGrid &g = h->sg;
    MarchArgs a;
    fill_args(h, a);
    dim3 grid;
    sg_blocks(h, grid);
    const dim3 block(TX, TY);
    const bool single = g.sp.ntypes == 1 && g.sp.nstates[0] == 1;
    prof_begin(h);
    if (single) {
        if (write_g) k_march<true, true, true><<<grid, block, 0, h->stream>>>(g.sp, a);
        else k_march<true, false, true><<<grid, block, 0, h->stream>>>(g.sp, a);
    } else {
        if (write_g) k_march<true, true, false><<<grid, block, 0, h->stream>>>(g.sp, a);
        else k_march<true, false, false><<<grid, block, 0, h->stream>>>(g.sp, a);
    }
    prof_end(h);
    h->launches++;
    g.cx ^= 1;
    g.cv ^= 1;
    if (write_g) g.cg ^= 1;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

}  // namespace mm
