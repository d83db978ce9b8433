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
// Apron cells are recomputed by neighbouring blocks (tile 32 x TY threads -> 30 x (TY-2) owned nodes): the price of
// never materialising per-cell data.  DESIGN.md discusses the trade-off and the measured numbers.  The kernel itself
// lives in mm_march.cuh; this file holds the layout conversions, halo planes and the launch logic.
#include <algorithm>
#include <cstring>
#include <functional>
#include <queue>
#include <vector>

#include "mm_internal.h"
#include "mm_march.cuh"
#include "mm_march2.cuh"
#include "mm_reduce.cuh"

namespace mm {

// ---------------------------------------------------------------------------------------------------------------
// Ghost nodes.  Every array is padded by one node on each side of all three axes (mm_structured.cuh): a tile of the
// marching kernel then never crosses a periodic boundary - the ghosts hold the periodic images, positions already shifted
// by the domain vectors of the STORED frame (a, b along x / y here; c along z below).
//   pass 1 (k_halo_xy): x / y ghosts of the owned planes 1 .. nzl, corners included (source = wrapped node, shift qx a + qy b)
//   pass 2 (k_halo):    whole padded planes  p = 0 <- p = nzl (minus c),  p = nzl + 1 <- p = 1 (plus c)
struct HaloArgs {
    double *f[9];
    int nfields;
    int npos;  // the first npos fields are position components 0, 1, 2
    const double *rv;  // the nine domain-vector components of the stored frame (StepConsts::rv unless the caller knows better)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
k_halo_xy(const __grid_constant__ HaloArgs h, int nx, int ny, int nxp, int64_t plane, int nzl, const StepConsts *sc) {
    const int per_plane = 2 * (nx + 2) + 2 * ny;  // two full ghost rows (with corners) + two ghost columns
    const int64_t total = (int64_t)per_plane * nzl;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = 1 + (int)(i / per_plane);
        int j = (int)(i % per_plane), k, l;
        if (j < 2 * (nx + 2)) {
            l = (j < nx + 2) ? -1 : ny;
            k = (j % (nx + 2)) - 1;
        } else {
            j -= 2 * (nx + 2);
            k = (j < ny) ? -1 : nx;
            l = j % ny;
        }
        const int qx = (k < 0) ? -1 : (k >= nx ? 1 : 0), qy = (l < 0) ? -1 : (l >= ny ? 1 : 0);
        const int64_t dst = ((int64_t)p * (ny + 2) + l + 1) * nxp + k + kGhostX;
        const int64_t src = ((int64_t)p * (ny + 2) + (l - qy * ny) + 1) * nxp + (k - qx * nx) + kGhostX;
        for (int f = 0; f < h.nfields; f++) {
            const double shift = (f < h.npos) ? qx * h.rv[f] + qy * h.rv[3 + f] : 0.0;
            h.f[f][dst] = h.f[f][src] + shift;
        }
    }
}

// Fused-halo variant (the marching kernel has already stored the boundary planes into the neighbours' halo planes).
// Slabs: block 0 first announces the delivery - the marching kernel has completed (stream order), a system-scope fence
// makes its peer stores visible before the two counters move - then every block fills the x / y ghosts of the owned planes,
// waits until both neighbours have announced this exchange too (epoch = exchanges completed, advanced by the last block),
// and fills the ghosts of the two halo planes.  Their nodes were written by another GPU: read them past the L1.
__global__ void __launch_bounds__(256)
k_halo_xy_fused(const __grid_constant__ HaloArgs h, int nx, int ny, int nxp, int nzl, const StepConsts *sc,
                const unsigned long long *flags, unsigned long long *epoch, unsigned int *done_blocks, unsigned long long *flag_lo,
                unsigned long long *flag_hi) {
    const int per_plane = 2 * (nx + 2) + 2 * ny;
    auto ghost = [&](int64_t i, int p, bool remote) {
        int j = (int)(i % per_plane), k, l;
        if (j < 2 * (nx + 2)) {
            l = (j < nx + 2) ? -1 : ny;
            k = (j % (nx + 2)) - 1;
        } else {
            j -= 2 * (nx + 2);
            k = (j < ny) ? -1 : nx;
            l = j % ny;
        }
        const int qx = (k < 0) ? -1 : (k >= nx ? 1 : 0), qy = (l < 0) ? -1 : (l >= ny ? 1 : 0);
        const int64_t dst = ((int64_t)p * (ny + 2) + l + 1) * nxp + k + kGhostX;
        const int64_t src = ((int64_t)p * (ny + 2) + (l - qy * ny) + 1) * nxp + (k - qx * nx) + kGhostX;
        // all loads first: the fields may alias as far as the compiler knows, and nine dependent load -> store round trips
        // are most of this kernel's time
        double val[9];
#pragma unroll
        for (int f = 0; f < 9; f++)
            if (f < h.nfields) val[f] = remote ? __ldcg(h.f[f] + src) : h.f[f][src];
#pragma unroll
        for (int f = 0; f < 9; f++)
            if (f < h.nfields) h.f[f][dst] = val[f] + ((f < h.npos) ? qx * h.rv[f] + qy * h.rv[3 + f] : 0.0);
    };
    // exchanges completed so far: read by every block before the last one to finish can advance it
    const unsigned long long want = flags ? ld_acquire_sys_u64(epoch) + 1 : 0ull;
    if (flags && blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        atomicAdd_system(flag_lo, 1ull);
        atomicAdd_system(flag_hi, 1ull);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t i = first; i < (int64_t)per_plane * nzl; i += stride) ghost(i, 1 + (int)(i / per_plane), false);
    if (flags) {
        if (threadIdx.x == 0) {
            while (ld_acquire_sys_u64(flags) < want) {
            }
            while (ld_acquire_sys_u64(flags + 1) < want) {
            }
        }
        __syncthreads();
    }
    for (int64_t i = first; i < 2 * (int64_t)per_plane; i += stride) ghost(i, (i / per_plane) ? nzl + 1 : 0, flags != nullptr);
    if (flags) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(done_blocks, 1u) == gridDim.x - 1) {
                *done_blocks = 0;
                *epoch = want;
                __threadfence();
            }
        }
    }
}

// k_march2 takes the x / y images on load: after a launch that delivered its boundary planes to the neighbour slabs only
// the hand-over remains (announce, wait for both neighbours' announcements) - unless the launch's tail already did it.
__global__ void k_halo_handshake(const unsigned long long *flags, unsigned long long *epoch, unsigned long long *flag_lo,
                                 unsigned long long *flag_hi) {
    if (threadIdx.x == 0) {
        const unsigned long long want = *epoch + 1;
        __threadfence_system();
        atomicAdd_system(flag_lo, 1ull);
        atomicAdd_system(flag_hi, 1ull);
        while (ld_acquire_sys_u64(flags) < want) {
        }
        while (ld_acquire_sys_u64(flags + 1) < want) {
        }
        *epoch = want;
    }
}

__global__ void __launch_bounds__(256)
k_halo(const __grid_constant__ HaloArgs h, int64_t plane, int nzl, const StepConsts *sc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
        for (int f = 0; f < h.nfields; f++) {
            const double c = (f < h.npos) ? h.rv[6 + f] : 0.0;
            h.f[f][i] = h.f[f][(int64_t)nzl * plane + i] - c;
            h.f[f][(int64_t)(nzl + 1) * plane + i] = h.f[f][plane + i] + c;
        }
    }
}

// cell types: the same two passes on bytes
__global__ void __launch_bounds__(256)
k_halo_xy_u8(uint8_t *f, int nx, int ny, int nxp, int nzl) {
    const int per_plane = 2 * (nx + 2) + 2 * ny;
    const int64_t total = (int64_t)per_plane * nzl;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = 1 + (int)(i / per_plane);
        int j = (int)(i % per_plane), k, l;
        if (j < 2 * (nx + 2)) {
            l = (j < nx + 2) ? -1 : ny;
            k = (j % (nx + 2)) - 1;
        } else {
            j -= 2 * (nx + 2);
            k = (j < ny) ? -1 : nx;
            l = j % ny;
        }
        const int ks = (k < 0) ? nx - 1 : (k >= nx ? 0 : k), ls = (l < 0) ? ny - 1 : (l >= ny ? 0 : l);
        f[((int64_t)p * (ny + 2) + l + 1) * nxp + k + kGhostX] = f[((int64_t)p * (ny + 2) + ls + 1) * nxp + ks + kGhostX];
    }
}

__global__ void __launch_bounds__(256)
k_halo_u8(uint8_t *f, int64_t plane, int nzl) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
        f[i] = f[(int64_t)nzl * plane + i];
        f[(int64_t)(nzl + 1) * plane + i] = f[plane + i];
    }
}

// reference order (AoS, id = (k*ny + l)*nz + m) <-> z-major SoA planes (owned planes only).
// Tiled through shared memory: a block handles a 32 (k) x 32 (m) tile of one l row, so that the SoA side is accessed
// along k (its unit stride) and the AoS side along m (96 consecutive doubles per k) - both sides fully coalesced.
constexpr int TT = 32;

__global__ void __launch_bounds__(TT * 8)
k_aos_to_soa(const double *__restrict__ aos, double *s0, double *s1, double *s2, int nx, int ny, int nz, int nxp) {
    __shared__ double tile[3][TT][TT + 1];  // [d][k][m]
    const int k0 = blockIdx.x * TT, l = blockIdx.y, m0 = blockIdx.z * TT;
    const int mt = min(TT, nz - m0), kt = min(TT, nx - k0);
    const int tid = threadIdx.y * TT + threadIdx.x;
    for (int kk = 0; kk < kt; kk++) {  // one k row of the tile: 3*mt consecutive doubles of the AoS array
        const int64_t base = (((int64_t)(k0 + kk) * ny + l) * nz + m0) * 3;
        for (int j = tid; j < 3 * mt; j += TT * 8) tile[j % 3][kk][j / 3] = aos[base + j];
    }
    __syncthreads();
    double *dst[3] = {s0, s1, s2};
    const int kk = threadIdx.x;
    for (int mm = threadIdx.y; mm < mt; mm += 8) {
        if (kk < kt) {
            const int64_t at = ((int64_t)(m0 + mm + 1) * (ny + 2) + l + 1) * nxp + k0 + kk + kGhostX;  // skip the ghosts
#pragma unroll
            for (int d = 0; d < 3; d++) dst[d][at] = tile[d][kk][mm];
        }
    }
}

__global__ void __launch_bounds__(TT * 8)
k_soa_to_aos(const double *__restrict__ s0, const double *__restrict__ s1, const double *__restrict__ s2, double *aos,
             int nx, int ny, int nz, int nxp) {
    __shared__ double tile[3][TT][TT + 1];  // [d][k][m]
    const int k0 = blockIdx.x * TT, l = blockIdx.y, m0 = blockIdx.z * TT;
    const int mt = min(TT, nz - m0), kt = min(TT, nx - k0);
    const int tid = threadIdx.y * TT + threadIdx.x;
    const double *src[3] = {s0, s1, s2};
    const int kk = threadIdx.x;
    for (int mm = threadIdx.y; mm < mt; mm += 8) {
        if (kk < kt) {
            const int64_t at = ((int64_t)(m0 + mm + 1) * (ny + 2) + l + 1) * nxp + k0 + kk + kGhostX;
#pragma unroll
            for (int d = 0; d < 3; d++) tile[d][kk][mm] = src[d][at];
        }
    }
    __syncthreads();
    for (int k2 = 0; k2 < kt; k2++) {
        const int64_t base = (((int64_t)(k0 + k2) * ny + l) * nz + m0) * 3;
        for (int j = tid; j < 3 * mt; j += TT * 8) aos[base + j] = tile[j % 3][k2][j / 3];
    }
}

__global__ void __launch_bounds__(256)
k_mass_to_soa(const double *__restrict__ masses, double *m, double *minv, int nx, int ny, int nz, int nxp) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nx), l = (int)((i / nx) % ny), mm_ = (int)(i / ((int64_t)nx * ny));
        const double v = masses[((int64_t)k * ny + l) * nz + mm_];
        const int64_t at = ((int64_t)(mm_ + 1) * (ny + 2) + l + 1) * nxp + k + kGhostX;
        m[at] = v;
        minv[at] = 1.0 / v;
    }
}

__global__ void __launch_bounds__(256)
k_type_to_soa(const uint8_t *__restrict__ info, uint8_t *type, int nx, int ny, int nz, int nxp) {
    const int64_t n = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nx), l = (int)((i / nx) % ny), m = (int)(i / ((int64_t)nx * ny));
        type[((int64_t)(m + 1) * (ny + 2) + l + 1) * nxp + k + kGhostX] = info[((int64_t)k * ny + l) * nz + m] & 15u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// model: the averaged `original` cell works on Hs = 4 H (what the separable stencil produces), the eight-corner `default`
// cell on Hs = H_a with weight 1/8: Bq = w V0 / (2 alpha^4) K' A K and c0 = alpha^2 h0 h0^T with (alpha, w) = (4, 1) / (1, 1/8)
static void fold_sparams(const KParams &kp, SParams &sp, int model) {
    const long double bq_div = model == MM_MODEL_DEFAULT ? 16.0L : 512.0L, c0_mul = model == MM_MODEL_DEFAULT ? 1.0L : 16.0L;
    sp.ntypes = kp.ntypes;
    for (int t = 0; t < MM_MAX_TYPES; t++) {
        sp.nstates[t] = kp.nstates[t];
        sp.offset[t] = kp.offset[t];
        sp.kT[t] = kp.kT[t];
    }
    static const int vi[6] = {0, 1, 2, 1, 0, 0}, vj[6] = {0, 1, 2, 2, 2, 1};
    for (int s = 0; s < MM_MAX_STATES; s++) {
        const StateP &P = kp.st[s];
        SState &S = sp.st[s];
        S.efree = P.efree;
        if (P.v0 == 0.0) {  // unused slot
            for (int i = 0; i < 36; i++) S.Bq[i] = 0.0;
            for (int i = 0; i < 6; i++) S.c0[i] = 0.0;
            continue;
        }
        // K: Voigt form of d -> hi d hi^T (shear columns doubled);  Kp: Voigt form of s -> hi^T s hi
        long double K[6][6], Kp[6][6], AK[6][6];
        for (int I = 0; I < 6; I++)
            for (int J = 0; J < 6; J++) {
                const int a = vi[I], b = vj[I], k = vi[J], l = vj[J];
                K[I][J] = (long double)P.hi[a * 3 + k] * P.hi[b * 3 + l];
                if (k != l) K[I][J] += (long double)P.hi[a * 3 + l] * P.hi[b * 3 + k];
                // Kp[J'][I']: J' = (a, b) is the output entry, I' = (k, l) the input entry
                Kp[I][J] = (long double)P.hi[k * 3 + a] * P.hi[l * 3 + b];
                if (k != l) Kp[I][J] += (long double)P.hi[l * 3 + a] * P.hi[k * 3 + b];
            }
        for (int I = 0; I < 6; I++)
            for (int J = 0; J < 6; J++) {
                long double acc = 0.0L;
                for (int M = 0; M < 6; M++) acc += (long double)P.A[I * 6 + M] * K[M][J];
                AK[I][J] = acc;
            }
        for (int I = 0; I < 6; I++)
            for (int J = 0; J < 6; J++) {
                long double acc = 0.0L;
                for (int M = 0; M < 6; M++) acc += Kp[I][M] * AK[M][J];
                // eps = 1/2 K (c - c0), c = Hs Hs^T / 16, Sq = V0/16 hi^T s hi  ->  1/2 * 1/16 * 1/16
                S.Bq[I * 6 + J] = (double)(acc * (long double)P.v0 / bq_div);
            }
        // c0 = h0 h0^T; h0 itself is not kept in StateP: invert hi (3x3 adjugate)
        const double *m = P.hi;
        const long double det = (long double)m[0] * ((long double)m[4] * m[8] - (long double)m[5] * m[7]) -
                                (long double)m[1] * ((long double)m[3] * m[8] - (long double)m[5] * m[6]) +
                                (long double)m[2] * ((long double)m[3] * m[7] - (long double)m[4] * m[6]);
        long double h0[9];
        h0[0] = ((long double)m[4] * m[8] - (long double)m[5] * m[7]) / det;
        h0[1] = ((long double)m[2] * m[7] - (long double)m[1] * m[8]) / det;
        h0[2] = ((long double)m[1] * m[5] - (long double)m[2] * m[4]) / det;
        h0[3] = ((long double)m[5] * m[6] - (long double)m[3] * m[8]) / det;
        h0[4] = ((long double)m[0] * m[8] - (long double)m[2] * m[6]) / det;
        h0[5] = ((long double)m[2] * m[3] - (long double)m[0] * m[5]) / det;
        h0[6] = ((long double)m[3] * m[7] - (long double)m[4] * m[6]) / det;
        h0[7] = ((long double)m[1] * m[6] - (long double)m[0] * m[7]) / det;
        h0[8] = ((long double)m[0] * m[4] - (long double)m[1] * m[3]) / det;
        for (int I = 0; I < 6; I++) {
            const int a = vi[I], b = vj[I];
            long double acc = 0.0L;
            for (int j = 0; j < 3; j++) acc += h0[a * 3 + j] * h0[b * 3 + j];
            S.c0[I] = (double)(c0_mul * acc);
        }
    }
}

// TMA descriptors of the padded SoA arrays (rank 3: x fastest, then y, then z; box = one tile of one plane).  The encoder
// lives in the driver library; it is reached through the runtime (cudaGetDriverEntryPoint), so nothing links to libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int sg_encode_maps(mm_handle *h, int rows) {
    SGrid &g = h->sg;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            return MM_ERR_CUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)g.nxp, (cuuint64_t)(g.ny + 2), (cuuint64_t)(g.nzl + 3)};
    const cuuint64_t strides[2] = {(cuuint64_t)g.nxp * 8, (cuuint64_t)g.plane * 8};
    const cuuint32_t box[3] = {(cuuint32_t)kBoxW, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    auto one = [&](CUtensorMap *m, double *p) {
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    bool ok = true;
    for (int c = 0; c < 2; c++)
        for (int d = 0; d < 3; d++) ok = ok && one(&g.tm_x[c][d], g.x[c][d]) && one(&g.tm_v[c][d], g.v[c][d]) && one(&g.tm_g[c][d], g.g[c][d]);
    ok = ok && one(&g.tm_m, g.m) && one(&g.tm_minv, g.minv);
    // k_march2: image columns of the edge tiles (2 columns x rows, starting on an even column)
    const cuuint32_t boxw[3] = {2, (cuuint32_t)rows, 1};
    auto onew = [&](CUtensorMap *m, double *p) {
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p, dims, strides, boxw, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    for (int c = 0; c < 2; c++)
        for (int d = 0; d < 3; d++) ok = ok && onew(&g.tw_x[c][d], g.x[c][d]) && onew(&g.tw_v[c][d], g.v[c][d]) && onew(&g.tw_g[c][d], g.g[c][d]);
    g.tma_rows = ok ? rows : 0;
    return ok ? MM_OK : MM_ERR_CUDA;
}

// index of an array inside the block: quantity 0 x / 1 v / 2 g, copy c, component d; 18 = m, 19 = 1/m
static inline int sg_array(int quantity, int c, int d) { return (quantity * 2 + c) * 3 + d; }

bool sg_eligible(const mm_handle *h) {
    if (!h->structured || h->nx < 2 || h->ny < 2 || h->nz < 2) return false;
    // the `default` (eight-corner) model exists on k_march2 only: one cell type with one metastable state
    return h->model == MM_MODEL_ORIGINAL || (h->kp.ntypes == 1 && h->kp.nstates[0] == 1);
}

static int sg_tile_rows_total(const SGrid &g) { return g.march2 ? g.rpt * g.tile_rows : g.tile_rows; }

int sg_blocks(const mm_handle *h, dim3 &grid) {
    const SGrid &g = h->sg;
    if (g.march2 && g.nitems > 0) {  // one block per work item (sg_plan_items)
        grid = dim3((unsigned)g.nitems, 1, 1);
        return g.nitems;
    }
    const int ox = TX - 2, oy = sg_tile_rows_total(g) - 2;
    grid = dim3((g.nx + ox - 1) / ox, (g.ny + oy - 1) / oy, (g.nzl + g.chunk - 1) / g.chunk);
    return (int)(grid.x * grid.y * grid.z);
}

// Chunk length along z.  Every block marches chunk + 2 planes (one warm-up plane below, one plane above), one block per
// SM: with T tiles and C = ceil(nzl / chunk) chunks the launch takes about ceil(T C / SMs) rounds of chunk + 2 plane
// iterations.  Pick the chunk that minimises rounds * (chunk + 2); ties go to the longer chunk (fewer warm-up planes).
static int sg_pick_chunk(const mm_handle *h) {
    const SGrid &g = h->sg;
    const int ox = TX - 2, oy = sg_tile_rows_total(g) - 2;
    const int64_t tiles = (int64_t)((g.nx + ox - 1) / ox) * ((g.ny + oy - 1) / oy);
    int best = g.nzl;
    int64_t best_cost = -1;
    for (int nchunks = 1; nchunks <= g.nzl; nchunks++) {
        const int chunk = (g.nzl + nchunks - 1) / nchunks;
        if (chunk < 4 && nchunks > 1) break;
        if ((g.nzl + chunk - 1) / chunk != nchunks) continue;
        const int64_t rounds = (tiles * nchunks + h->num_sms - 1) / h->num_sms;
        const int64_t cost = rounds * (chunk + 2);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = chunk;
        }
    }
    return best;
}

// ---- work items of k_march2 ----------------------------------------------------------------------------------------------
// A uniform chunk length leaves the last round of the SMs partly empty (256^3 on 8 GPUs: 171 tiles x 4 chunks = 684 blocks of
// 10 plane iterations = 4.6 rounds on 148 SMs, 50 iterations against 39 with perfect balance).  The planner splits the
// tiles into two classes - class A: the first TA tiles, nA chunks each (long blocks, dispatched first); class B: the
// rest, nB >= nA chunks each (short blocks that fill the tail) - and picks (TA, nA, nB) by simulating the block scheduler
// (blocks go, in index order, to the SM that becomes free first; a block costs (planes + 2) x weight plane iterations; edge
// tiles weigh more when they fetch the periodic images themselves).
static double plan_makespan(const std::vector<double> &costs, int nsm) {
    std::priority_queue<double, std::vector<double>, std::greater<double>> sms;
    for (int i = 0; i < nsm; i++) sms.push(0.0);
    double last = 0.0;
    for (double c : costs) {
        const double t = sms.top() + c;
        sms.pop();
        sms.push(t);
        last = std::max(last, t);
    }
    return last;
}

// Pure host part (also exported as mm_plan_schedule for the CPU tests): ntx x nty tiles, P owned planes, S SMs.
// uniform_chunk > 0 asks for equal chunks of that length for every tile.  Returns the items (tile x, tile y, first owned
// plane, one past the last; planes count from 1) in dispatch order.
void sg_plan_schedule(int ntx, int nty, int P, int S, int images_on_load, int uniform_chunk, std::vector<int4> &items,
                      double *cost, double *ideal, int *chunk_b) {
    const int T = ntx * nty;
    struct Tile {
        int bx, by;
        double w;
    };
    std::vector<Tile> tiles;
    // Without images on load all tiles weigh the same and keep the natural order (x fastest: blocks that run side by side
    // are neighbours in the grid and share their apron rows in L2).  With images on load the heavier edge tiles go last,
    // into class B, where their extra cost is spread over short chunks (N = 2: 0.789 ms against 0.817 in natural order).
    for (int pass = 0; pass < (images_on_load ? 2 : 1); pass++)
        for (int by = 0; by < nty; by++)
            for (int bx = 0; bx < ntx; bx++) {
                const bool edge = bx == 0 || by == 0 || bx == ntx - 1 || by == nty - 1;
                if (images_on_load && edge != (pass == 1)) continue;
                tiles.push_back({bx, by, (edge && images_on_load) ? 1.2 : 1.0});
            }
    auto build = [&](int TA, int nA, int nB, std::vector<int4> *out, std::vector<double> &costs) {
        costs.clear();
        if (out) out->clear();
        for (int cls = 0; cls < 2; cls++) {
            const int t0 = cls == 0 ? 0 : TA, t1 = cls == 0 ? TA : T, n = cls == 0 ? nA : nB;
            for (int c = 0; c < n; c++)  // chunk-major: the n pieces of a tile are spread over the dispatch order
                for (int t = t0; t < t1; t++) {
                    const int lo = 1 + (int)((int64_t)P * c / n), hi = 1 + (int)((int64_t)P * (c + 1) / n);
                    if (hi <= lo) continue;
                    costs.push_back((hi - lo + 2) * tiles[t].w);
                    if (out) out->push_back(make_int4(tiles[t].bx, tiles[t].by, lo, hi));
                }
        }
    };
    int bTA = 0, bnA = 1, bnB = 1;
    double best = -1.0;
    std::vector<double> costs;
    if (uniform_chunk > 0) {
        bnA = bnB = (P + uniform_chunk - 1) / uniform_chunk;
        bTA = 0;
    } else {
        const int nmax = std::max(1, P / 2);
        static const int cand[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 64};
        for (int nA : cand) {
            if (nA > nmax || nA > 16) break;
            for (int nB : cand) {
                if (nB < nA) continue;
                if (nB > nmax || (int64_t)T * nB > 16384) break;
                for (int TA = 0; TA <= T; TA = (TA == T) ? T + 1 : std::min(T, TA + S)) {
                    if (nA == nB && TA != 0) continue;  // the uniform schedule once
                    build(TA, nA, nB, nullptr, costs);
                    const double m = plan_makespan(costs, S);
                    if (best < 0.0 || m < best * (1.0 - 1e-9) || (m <= best * (1.0 + 1e-9) && (int64_t)TA * nA + (int64_t)(T - TA) * nB < (int64_t)bTA * bnA + (int64_t)(T - bTA) * bnB)) {
                        best = m;
                        bTA = TA;
                        bnA = nA;
                        bnB = nB;
                    }
                }
            }
        }
    }
    build(bTA, bnA, bnB, &items, costs);
    if (cost) *cost = plan_makespan(costs, S);
    double total = 0.0;
    for (const Tile &t : tiles) total += (double)P * t.w;
    if (ideal) *ideal = total / S;
    if (chunk_b) *chunk_b = (P + bnB - 1) / bnB;
}

static int sg_plan_items(mm_handle *h, int chunk_override) {
    SGrid &g = h->sg;
    const int ox = TX - 2, oy = sg_tile_rows_total(g) - 2;
    const int ntx = (g.nx + ox - 1) / ox, nty = (g.ny + oy - 1) / oy;
    // uniform chunks (tuning / A-B): the given length or the cost model's
    const int uniform = chunk_override > 0 ? chunk_override : (g.plan_two_class ? 0 : sg_pick_chunk(h));
    std::vector<int4> items;
    sg_plan_schedule(ntx, nty, g.nzl, h->num_sms, g.wrap_on_load, uniform, items, &g.plan_cost, &g.plan_ideal, &g.chunk);
    g.ntx = ntx;
    g.nty = nty;
    g.nitems = (int)items.size();
    if (g.nitems > g.nitems_alloc) {
        if (g.d_items) cudaFree(g.d_items);
        g.d_items = nullptr;
        MM_CUDA(cudaMalloc(&g.d_items, sizeof(int4) * (size_t)g.nitems));
        g.nitems_alloc = g.nitems;
    }
    MM_CUDA(cudaMemcpyAsync(g.d_items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));  // `items` is a local
    return MM_OK;
}

// (re)derive everything that depends on the tile shape: kernel family, tensor maps, chunk, number of block partials
int sg_retile(mm_handle *h, int chunk_override) {
    SGrid &g = h->sg;
    const bool single = g.sp.ntypes == 1 && g.sp.nstates[0] == 1;
    const bool want2 = g.march2_wanted && single && g.mass_uniform;
    if (want2) {
        g.march2 = 1;
        if (g.tma_rows != g.rpt * g.tile_rows && sg_encode_maps(h, g.rpt * g.tile_rows) != MM_OK) g.march2 = 0;
    } else {
        g.march2 = 0;
    }
    if (!g.march2 && g.tma_rows != g.tile_rows) sg_encode_maps(h, g.tile_rows);
    // Periodic images on load + tail launch (nothing between two marching launches but one single-block kernel) pay off
    // when slabs have to synchronise anyway and on small grids, where a step is a handful of short launches; from ~160^3
    // nodes per GPU on, the ghost-fill launch is cheaper than the image copies of the edge tiles (measured on one GPU,
    // profiles/r02/z: 64^3 NPT 0.092 vs 0.099 ms, 128^3 0.316 vs 0.337, 160^3 0.460 vs 0.440, 256^3 1.437 vs 1.406)
    g.wrap_on_load = g.wrap_wanted >= 0 ? g.wrap_wanted : ((h->slab_count > 1 || h->nnodes <= 128 * 128 * 128) ? 1 : 0);
    g.nitems = 0;
    g.chunk = chunk_override > 0 ? chunk_override : sg_pick_chunk(h);
    if (g.march2) {
        const int rc = sg_plan_items(h, chunk_override);
        if (rc != MM_OK) return rc;
    }
    dim3 grid;
    const int nb = sg_blocks(h, grid);
    if (nb > g.nblocks_alloc) {
        if (g.d_partials) cudaFree(g.d_partials);
        g.d_partials = nullptr;
        MM_CUDA(cudaMalloc(&g.d_partials, sizeof(double) * (size_t)nb * kRedSlots));
        g.nblocks_alloc = nb;
    }
    g.nblocks = nb;
    return MM_OK;
}

int sg_setup(mm_handle *h) {
    SGrid &g = h->sg;
    g.nx = h->nx;
    g.ny = h->ny;
    g.nzl = h->nz;
    // The reference enumerates id = (k*ny + l)*nz + m with k along the FIRST domain vector.  Device planes are indexed
    // (x = k fastest, y = l, z = m slowest): the marching direction is the reference's third axis, the lanes run along
    // its first axis.
    g.nxp = (g.nx + kGhostX + 1 + 1) & ~1;  // ghost nodes on both sides (mm_structured.cuh), even pitch: 16-byte rows for TMA
    g.plane = (int64_t)g.nxp * (g.ny + 2);
    g.npad = g.plane * (g.nzl + 3);  // two halo planes + one spare plane for the prefetch of the marching kernel
    fold_sparams(h->kp, g.sp, h->model);
    if (h->model == MM_MODEL_DEFAULT) {  // the eight-corner cell keeps one node row per thread (register budget)
        g.rpt = 1;
        g.tile_rows = 8;
    }
    // all node arrays in one allocation: array a at block + a * stride (order: sg_array)
    // (the stride is skewed by an odd number of 256-byte lines: with a power-of-two-ish stride the same node of all 20
    // arrays falls on the same memory channel and the 20 streams of the marching kernel queue up there)
    g.stride = ((g.npad + 31) & ~(int64_t)31) + 32 * 37;
    const size_t bytes = sizeof(double) * (size_t)g.stride * 20;
    MM_CUDA(cudaMalloc(&g.block, bytes));
    MM_CUDA(cudaMemsetAsync(g.block, 0, bytes, h->stream));
    for (int c = 0; c < 2; c++)
        for (int d = 0; d < 3; d++) {
            g.x[c][d] = g.block + (size_t)g.stride * sg_array(0, c, d);
            g.v[c][d] = g.block + (size_t)g.stride * sg_array(1, c, d);
            g.g[c][d] = g.block + (size_t)g.stride * sg_array(2, c, d);
        }
    g.m = g.block + (size_t)g.stride * 18;
    g.minv = g.block + (size_t)g.stride * 19;
    // single slab: the "neighbours" are the periodic images of the slab itself
    for (int s = 0; s < 2; s++) {
        g.nb_block[s] = g.block;
        g.nb_stride[s] = g.stride;
        g.nb_nzl[s] = g.nzl;
        g.nb_flag[s] = nullptr;
    }
    g.fused = (h->slab_count <= 1) ? 1 : 0;  // slabs switch it on once the peers are mapped (mm_comm.cu)
    MM_CUDA(cudaMalloc(&g.type, g.npad));
    MM_CUDA(cudaMalloc(&g.d_sc, 512));  // StepConsts (248 B) + scratch doubles from byte 256 on (k_mass_range)
    MM_CUDA(cudaMalloc(&g.d_sp, sizeof(SParams)));
    MM_CUDA(cudaMalloc(&g.d_tail_counter, 64));
    MM_CUDA(cudaMemsetAsync(g.d_tail_counter, 0, 64, h->stream));
    MM_CUDA(cudaMemcpyAsync(g.d_sp, &g.sp, sizeof(SParams), cudaMemcpyHostToDevice, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    MM_CUDA(cudaHostAlloc(&g.h_sc, sizeof(StepConsts), cudaHostAllocDefault));
    {
        const int rc = sg_retile(h, 0);
        if (rc != MM_OK) return rc;
    }
    MM_CUDA(cudaMemsetAsync(g.type, 0, g.npad, h->stream));
    k_type_to_soa<<<grid_for(h, h->ncells, 256), 256, 0, h->stream>>>(h->d_cell_info, g.type, g.nx, g.ny, g.nzl, g.nxp);
    k_halo_xy_u8<<<grid_for(h, (int64_t)(2 * (g.nx + 2) + 2 * g.ny) * g.nzl, 256), 256, 0, h->stream>>>(g.type, g.nx, g.ny, g.nxp, g.nzl);
    if (h->slab_count <= 1)  // slabs receive the neighbours' boundary types in mm_comm_init
        k_halo_u8<<<grid_for(h, g.plane, 256), 256, 0, h->stream>>>(g.type, g.plane, g.nzl);
    MM_CUDA(cudaGetLastError());
    g.active = 1;
    return MM_OK;
}

void sg_free(mm_handle *h) {
    SGrid &g = h->sg;
    cudaFree(g.block);
    g.block = nullptr;
    cudaFree(g.type);
    cudaFree(g.d_sc);
    cudaFree(g.d_sp);
    cudaFree(g.d_tail_counter);
    g.d_tail_counter = nullptr;
    cudaFree(g.d_items);
    g.d_items = nullptr;
    g.nitems = g.nitems_alloc = 0;
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

// rv_src: device pointer to the nine domain-vector components the position ghosts are shifted by (default: StepConsts::rv);
// st: launch on this stream instead of the handle's (single slab only: the MD step runs the ghost fill beside its scalar kernel)
int sg_halo(mm_handle *h, bool pos, bool vel, bool grad, const double *rv_src, cudaStream_t st) {
    SGrid &g = h->sg;
    cudaStream_t stream = st ? st : h->stream;
    HaloArgs ha;
    ha.nfields = 0;
    ha.npos = 0;
    ha.rv = rv_src ? rv_src : reinterpret_cast<const double *>(g.d_sc) + offsetof(StepConsts, rv) / sizeof(double);
    if (pos) {
        for (int d = 0; d < 3; d++) ha.f[ha.nfields++] = g.x[g.cx][d];
        ha.npos = 3;
    }
    if (vel)
        for (int d = 0; d < 3; d++) ha.f[ha.nfields++] = g.v[g.cv][d];
    if (grad)
        for (int d = 0; d < 3; d++) ha.f[ha.nfields++] = g.g[g.cg][d];
    if (ha.nfields == 0) return MM_OK;
    const int mask = (pos ? 1 : 0) | (vel ? 2 : 0) | (grad ? 4 : 0);
    if (g.march2 && g.wrap_on_load) {  // x / y images are taken on load (mm_march2.cuh): only the z halo planes have to be current
        const bool delivered = g.fused && g.fused_mask == mask;
        g.fused_mask = 0;
        if (delivered) {
            if (h->slab_count > 1 && !g.tail_done) {
                k_halo_handshake<<<1, 32, 0, stream>>>(g.halo_flags, g.halo_epoch, g.nb_flag[0], g.nb_flag[1]);
                h->launches++;
            }
            g.tail_done = 0;
            return MM_OK;
        }
        if (h->slab_count > 1) return comm_halo(h, ha.f, ha.nfields, ha.npos);
        k_halo<<<grid_for(h, g.plane, 256), 256, 0, stream>>>(ha, g.plane, g.nzl, g.d_sc);
        h->launches++;
        return MM_OK;
    }
    if (g.fused && g.fused_mask == mask) {  // the marching kernel delivered the boundary planes itself
        g.fused_mask = 0;
        k_halo_xy_fused<<<grid_for(h, (int64_t)(2 * (g.nx + 2) + 2 * g.ny) * (g.nzl + 2), 256), 256, 0, stream>>>(
            ha, g.nx, g.ny, g.nxp, g.nzl, g.d_sc, g.halo_flags, g.halo_epoch, g.halo_done, g.nb_flag[0], g.nb_flag[1]);
        h->launches++;
        return MM_OK;
    }
    g.fused_mask = 0;
    k_halo_xy<<<grid_for(h, (int64_t)(2 * (g.nx + 2) + 2 * g.ny) * g.nzl, 256), 256, 0, stream>>>(ha, g.nx, g.ny, g.nxp, g.plane, g.nzl, g.d_sc);
    h->launches++;
    if (h->slab_count > 1) return comm_halo(h, ha.f, ha.nfields, ha.npos);  // (padded) planes travel between the slabs
    k_halo<<<grid_for(h, g.plane, 256), 256, 0, stream>>>(ha, g.plane, g.nzl, g.d_sc);
    h->launches++;
    return MM_OK;
}

int sg_halo_mass(mm_handle *h) {
    SGrid &g = h->sg;
    HaloArgs ha;
    ha.nfields = 2;
    ha.npos = 0;
    ha.rv = nullptr;  // no position fields: never read
    ha.f[0] = g.m;
    ha.f[1] = g.minv;
    k_halo_xy<<<grid_for(h, (int64_t)(2 * (g.nx + 2) + 2 * g.ny) * g.nzl, 256), 256, 0, h->stream>>>(ha, g.nx, g.ny, g.nxp, g.plane, g.nzl, g.d_sc);
    h->launches++;
    if (h->slab_count > 1) return comm_halo(h, ha.f, 2, 0);
    k_halo<<<grid_for(h, g.plane, 256), 256, 0, h->stream>>>(ha, g.plane, g.nzl, g.d_sc);
    h->launches++;
    return MM_OK;
}

static dim3 transpose_grid(const SGrid &g) { return dim3((g.nx + TT - 1) / TT, g.ny, (g.nzl + TT - 1) / TT); }

int sg_pos_from_aos(mm_handle *h, const double *d_aos) {
    SGrid &g = h->sg;
    k_aos_to_soa<<<transpose_grid(g), dim3(TT, 8), 0, h->stream>>>(d_aos, g.x[g.cx][0], g.x[g.cx][1], g.x[g.cx][2], g.nx, g.ny, g.nzl, g.nxp);
    h->launches++;
    return sg_halo(h, true, false, false);
}

int sg_vel_from_aos(mm_handle *h, const double *d_aos) {
    SGrid &g = h->sg;
    k_aos_to_soa<<<transpose_grid(g), dim3(TT, 8), 0, h->stream>>>(d_aos, g.v[g.cv][0], g.v[g.cv][1], g.v[g.cv][2], g.nx, g.ny, g.nzl, g.nxp);
    h->launches++;
    return sg_halo(h, false, true, false);
}

// min / max of the node masses -> out[0], out[1] (one block; the masses are uploaded once per integrator)
__global__ void __launch_bounds__(256) k_mass_range(const double *__restrict__ masses, int64_t n, double *out) {
    __shared__ double lo[256], hi[256];
    double a = masses[0], b = masses[0];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        a = fmin(a, masses[i]);
        b = fmax(b, masses[i]);
    }
    lo[threadIdx.x] = a;
    hi[threadIdx.x] = b;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            lo[threadIdx.x] = fmin(lo[threadIdx.x], lo[threadIdx.x + s]);
            hi[threadIdx.x] = fmax(hi[threadIdx.x], hi[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = lo[0];
        out[1] = hi[0];
    }
}

int sg_mass_from_aos(mm_handle *h, const double *d_masses) {
    SGrid &g = h->sg;
    {
        // one node mass for the whole grid (micmec/utils.py:217 gives every node of a one-type grid the type mass): the
        // fast path takes it as a kernel argument; anything else keeps the mass arrays and runs on k_march
        double range[2] = {0.0, 0.0};
        k_mass_range<<<1, 256, 0, h->stream>>>(d_masses, h->nnodes, reinterpret_cast<double *>(g.d_sc) + 32);
        h->launches++;
        MM_CUDA(cudaMemcpyAsync(range, reinterpret_cast<double *>(g.d_sc) + 32, sizeof(range), cudaMemcpyDeviceToHost, h->stream));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        const int uniform = range[0] == range[1] ? 1 : 0;
        g.mass = range[0];
        if (uniform != g.mass_uniform) {
            g.mass_uniform = uniform;
            const int rc = sg_retile(h, 0);
            if (rc != MM_OK) return rc;
        }
    }
    k_mass_to_soa<<<grid_for(h, h->nnodes, 256), 256, 0, h->stream>>>(d_masses, g.m, g.minv, g.nx, g.ny, g.nzl, g.nxp);
    h->launches++;
    return sg_halo_mass(h);
}

int sg_to_aos(mm_handle *h, int which, double *d_aos) {  // 0 pos, 1 vel, 2 gpos
    SGrid &g = h->sg;
    double **src = which == 0 ? g.x[g.cx] : which == 1 ? g.v[g.cv] : g.g[g.cg];
    k_soa_to_aos<<<transpose_grid(g), dim3(TT, 8), 0, h->stream>>>(src[0], src[1], src[2], d_aos, g.nx, g.ny, g.nzl, g.nxp);
    h->launches++;
    return MM_OK;
}

static void fill_args(mm_handle *h, MarchArgs &a) {
    SGrid &g = h->sg;
    a.nx = g.nx;
    a.ny = g.ny;
    a.nzl = g.nzl;
    a.nxp = g.nxp;
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
    a.mass = g.mass;
    a.items = g.d_items;
    a.ntx = g.ntx;
    a.nty = g.nty;
    a.sc = g.d_sc;
    a.spg = g.d_sp;
    a.partials = g.d_partials;
    a.fused = 0;
    a.wrap_lo = a.wrap_hi = 0.0;
    for (int f = 0; f < 9; f++) a.halo_lo[f] = a.halo_hi[f] = nullptr;
    memset(&a.tail, 0, sizeof(a.tail));
}

// The tail of a k_march2 launch can take over the reduction, the exchange between the slabs and the scalar algebra when
// the boundary planes travel inside the marching kernel (one slab, or slabs in fused peer mode)
bool sg_tail_ok(const mm_handle *h) {
    const SGrid &g = h->sg;
    if (!g.march2 || !g.tail_wanted || !g.fused || !g.wrap_on_load) return false;
    return h->slab_count <= 1 || h->peer_mode != 0;
}

static void fill_tail(mm_handle *h, MarchArgs &a, const SgTail *tail) {
    SGrid &g = h->sg;
    if (!tail || !sg_tail_ok(h)) return;
    a.tail.enabled = 1;
    a.tail.ops = tail->ops;
    a.tail.counter = g.d_tail_counter;
    a.tail.state = tail->state;
    a.tail.rvecs_dev = h->d_rvecs;
    a.tail.sc_out = g.d_sc;
    a.tail.n3 = 3.0 * (double)h->nnodes_global;
    a.tail.nranks = h->slab_count > 1 ? h->slab_count : 1;
    a.tail.rank = h->slab_rank;
    a.tail.bases = h->slab_count > 1 ? h->d_peer_base : nullptr;
    a.tail.ctl = h->d_peer_ctl;
}

// The same tail as ONE single-block launch behind the marching kernel (the default: measured against the in-kernel tail
// in profiles/r02 - the last block of a marching launch runs the cold scalar code on an SM whose instruction caches hold the
// plane loop, and every boundary block has to drain its peer stores before taking its ticket; as a separate launch the
// tail costs one launch gap instead).  It replaces k_peer_allreduce + k_scalar + the halo hand-over of round 1.
__global__ void __launch_bounds__(256) k_tail(const __grid_constant__ TailArgs t, const double *partials, const int nblocks) {
    if (t.nranks > 1) __threadfence_system();  // stream order: the marching kernel's peer stores precede this launch
    march_tail(t, partials, nblocks);
}

// run the tail of the marching launch that was just issued with `a` (in-kernel tails have already done it)
static int finish_tail(mm_handle *h, MarchArgs &a) {
    if (!a.tail.enabled || h->sg.tail_in_kernel) return MM_OK;
    k_tail<<<1, 256, 0, h->stream>>>(a.tail, a.partials, h->sg.nblocks);
    h->launches++;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// Fused halo: targets of the boundary planes of what a launch writes (array indices ax / av / ag inside the block)
static void fill_fused(mm_handle *h, MarchArgs &a, int cx_out, int cv_out, int cg_out) {
    SGrid &g = h->sg;
    if (!g.fused) return;
    a.fused = 1;
    const int P = h->slab_count, r = h->slab_rank;
    for (int f = 0; f < 9; f++) {
        const int arr = f < 3 ? sg_array(0, cx_out, f) : f < 6 ? sg_array(1, cv_out, f - 3) : sg_array(2, cg_out, f - 6);
        // own plane 1 -> plane nzl' + 1 of the slab below; own plane nzl -> plane 0 of the slab above (pointers biased by
        // the own plane index, so that the kernel adds the same element index as for its own store)
        a.halo_lo[f] = g.nb_block[0] + (size_t)g.nb_stride[0] * arr + (int64_t)(g.nb_nzl[0] + 1 - 1) * g.plane;
        a.halo_hi[f] = g.nb_block[1] + (size_t)g.nb_stride[1] * arr - (int64_t)g.nzl * g.plane;
    }
    a.wrap_lo = (r == 0) ? 1.0 : 0.0;       // below rank 0 sits the last slab: its upper halo is my plane 1 PLUS c
    a.wrap_hi = (r == P - 1) ? -1.0 : 0.0;  // above the last rank sits slab 0: its lower halo is my plane nzl MINUS c
}

template <int STEP, bool SINGLE, int ROT, int VM, bool LEAN, int VAR, int TY>
static int launch_one(mm_handle *h, const MarchArgs &a, int write_g) {
    dim3 grid;
    sg_blocks(h, grid);
    prof_begin(h, STEP);
    // staged variant: kStages planes of the tile in dynamic shared memory (opt-in above 48 KB, once per instantiation)
    constexpr size_t dyn = (VAR & 2) ? sizeof(double) * (kStages * (STEP ? 11 : 3) * TY * kBoxW) : 0;
    if (dyn > 0) {
        static bool configured[64] = {false};
        if (!configured[h->device & 63]) {
            MM_CUDA(cudaFuncSetAttribute(k_march<STEP, SINGLE, ROT, VM, LEAN, VAR, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            configured[h->device & 63] = true;
        }
    }
    TmaMaps maps;
    if (VAR & 2) {
        const SGrid &g = h->sg;
        for (int d = 0; d < 3; d++) {
            maps.in[d] = g.tm_x[g.cx][d];
            maps.in[3 + d] = g.tm_v[g.cv][d];
            maps.in[6 + d] = g.tm_g[g.cg][d];
        }
        maps.in[9] = g.tm_m;
        maps.in[10] = g.tm_minv;
    } else {
        memset(&maps, 0, sizeof(maps));
    }
    k_march<STEP, SINGLE, ROT, VM, LEAN, VAR, TY><<<grid, dim3(TX, TY), dyn, h->stream>>>(h->sg.sp, a, maps, write_g);
    prof_end(h);
    h->launches++;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

template <bool SINGLE, int VAR>
static int launch_sel(mm_handle *h, const MarchArgs &a, bool step, int rot, int vm, bool lean, int write_g) {
    constexpr int TY = 8;
    if (step) {
        if (lean) {
            if (vm == 0) return launch_one<1, SINGLE, 0, 0, true, VAR, TY>(h, a, write_g);
            return launch_one<1, SINGLE, 0, 1, true, VAR, TY>(h, a, write_g);
        }
        if (vm == 0) return launch_one<1, SINGLE, 0, 0, false, VAR, TY>(h, a, write_g);
        if (vm == 1) return launch_one<1, SINGLE, 0, 1, false, VAR, TY>(h, a, write_g);
        return launch_one<1, SINGLE, 0, 2, false, VAR, TY>(h, a, write_g);
    }
    if (rot == 0) return launch_one<0, SINGLE, 0, 0, false, VAR, TY>(h, a, write_g);
    if (rot == 1) return launch_one<0, SINGLE, 1, 0, false, VAR, TY>(h, a, write_g);
    return launch_one<0, SINGLE, 2, 0, false, VAR, TY>(h, a, write_g);
}

template <int MODEL, int MODE, int ROT, int VM, bool LEAN, int RPT, int TY, int PIN, bool WRAP, int UNR>
static int launch_one2(mm_handle *h, const MarchArgs &a, int write_g) {
    using Cfg = March2Cfg<RPT, TY>;
    dim3 grid;
    sg_blocks(h, grid);
    prof_begin(h, MODE == M2_STEP ? 1 : 0);
    constexpr size_t dyn = Cfg::dyn_bytes(MODE == M2_STEP ? 9 : 3);
    static bool configured[64] = {false};
    if (!configured[h->device & 63]) {
        MM_CUDA(cudaFuncSetAttribute(k_march2<MODEL, MODE, ROT, VM, LEAN, RPT, TY, PIN, WRAP, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        configured[h->device & 63] = true;
    }
    TmaMaps maps;
    const SGrid &g = h->sg;
    for (int d = 0; d < 3; d++) {
        maps.in[d] = g.tm_x[g.cx][d];
        maps.in[3 + d] = g.tm_v[g.cv][d];
        maps.in[6 + d] = g.tm_g[g.cg][d];
    }
    for (int d = 0; d < 3; d++) {
        maps.xw[d] = g.tw_x[g.cx][d];
        maps.xw[3 + d] = g.tw_v[g.cv][d];
        maps.xw[6 + d] = g.tw_g[g.cg][d];
    }
    maps.in[9] = g.tm_m;
    maps.in[10] = g.tm_minv;
    k_march2<MODEL, MODE, ROT, VM, LEAN, RPT, TY, PIN, WRAP, UNR><<<grid, dim3(TX, TY), dyn, h->stream>>>(g.sp.st[0], a, maps, write_g);
    prof_end(h);
    h->launches++;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

// mode: M2_FORCE / M2_STEP / M2_VIRIAL (energy + virial only; rot 1).  Rows of Bq in per-thread registers (PIN, measured
// in profiles/r02/d, g): none for the fused step (its register budget is exhausted: 588 -> 60 bytes of spills), three for the
// force-only launches of the averaged model.
template <int MODEL, int RPT, int TY, bool WRAP>
static int launch_sel2(mm_handle *h, const MarchArgs &a, int mode, int rot, int vm, bool lean, int write_g) {
    constexpr int PS = 0, PF = MODEL == MM_MODEL_DEFAULT ? 0 : 3;
    if (mode == M2_STEP) {
        if (lean) {
            if (vm == 0) return launch_one2<MODEL, M2_STEP, 0, 0, true, RPT, TY, PS, WRAP, 1>(h, a, write_g);
            return launch_one2<MODEL, M2_STEP, 0, 1, true, RPT, TY, PS, WRAP, 1>(h, a, write_g);
        }
        if (vm == 0) return launch_one2<MODEL, M2_STEP, 0, 0, false, RPT, TY, PS, WRAP, 1>(h, a, write_g);
        if (vm == 1) return launch_one2<MODEL, M2_STEP, 0, 1, false, RPT, TY, PS, WRAP, 1>(h, a, write_g);
        return launch_one2<MODEL, M2_STEP, 0, 2, false, RPT, TY, PS, WRAP, 1>(h, a, write_g);
    }
    if (mode == M2_VIRIAL) {
        if (rot == 0) return launch_one2<MODEL, M2_VIRIAL, 0, 0, false, RPT, TY, PF, WRAP, 1>(h, a, 0);
        return launch_one2<MODEL, M2_VIRIAL, 1, 0, false, RPT, TY, PF, WRAP, 1>(h, a, 0);
    }
    if (rot == 0) return launch_one2<MODEL, M2_FORCE, 0, 0, false, RPT, TY, PF, WRAP, 1>(h, a, write_g);
    if (rot == 1) return launch_one2<MODEL, M2_FORCE, 1, 0, false, RPT, TY, PF, WRAP, 1>(h, a, write_g);
    return launch_one2<MODEL, M2_FORCE, 2, 0, false, RPT, TY, PF, WRAP, 1>(h, a, write_g);
}

// tile of the averaged model: 2 node rows per thread x 8 warps (30 x 14 owned nodes); profiles/r02/c has the alternatives
// (1 x 8: +10 %, 3 x 8: register spills, +40 %; 2 x 4 with two blocks per SM: +8 %).  The eight-corner model keeps one row.
bool sg_march2_config_ok(int rpt, int ty) { return (rpt == 2 || rpt == 1) && ty == 8; }

static int launch_march(mm_handle *h, const MarchArgs &a, int mode, int rot, int vm, bool lean, int write_g) {
    const SGrid &g = h->sg;
    if (g.march2 && h->model == MM_MODEL_DEFAULT) {
        if (g.wrap_on_load) return launch_sel2<MM_MODEL_DEFAULT, 1, 8, true>(h, a, mode, rot, vm, lean, write_g);
        return launch_sel2<MM_MODEL_DEFAULT, 1, 8, false>(h, a, mode, rot, vm, lean, write_g);
    }
    if (h->model != MM_MODEL_ORIGINAL) {
        set_error("the structured-grid kernels evaluate the 'default' model for one cell type, one state and one node mass only");
        return MM_ERR_INVALID;
    }
    if (g.march2) {
        if (g.wrap_on_load) return launch_sel2<MM_MODEL_ORIGINAL, 2, 8, true>(h, a, mode, rot, vm, lean, write_g);
        return launch_sel2<MM_MODEL_ORIGINAL, 2, 8, false>(h, a, mode, rot, vm, lean, write_g);
    }
    const bool step = mode == M2_STEP;
    const bool single = g.sp.ntypes == 1 && g.sp.nstates[0] == 1;
    // general path: several cell types / metastable states / node masses (register-prefetch loads, mass arrays)
    // (the tuning variants of k_march measured in round 1 - TMA staging, pairwise barriers, two planes per trip - are not
    // instantiated in the product library any more: one-type grids run on k_march2)
    if (single) return launch_sel<true, 0>(h, a, step, rot, vm, lean, write_g);
    return launch_sel<false, 0>(h, a, step, rot, vm, false, write_g);
}

// Force evaluation at the stored positions.  rot: 0 = positions are true as stored, 1 = apply the pending rotation
// Rpend on load, 2 = apply it and write the rotated positions to the other x set (which becomes current).
// write_g: store the node gradient into the CURRENT g set (FORCE mode never reads g).
int sg_force(mm_handle *h, bool write_g, int rot, bool virial_only, const SgTail *tail) {
    SGrid &g = h->sg;
    MarchArgs a;
    fill_args(h, a);
    fill_tail(h, a, tail);
    for (int d = 0; d < 3; d++) a.go[d] = g.g[g.cg][d];
    fill_fused(h, a, g.cx ^ 1, g.cv ^ 1, g.cg);
    // virial_only: energy and virial of a geometry whose gradient nobody reads (k_march2 skips the gather; k_march
    // computes it anyway and drops it)
    const int mode = (virial_only && !write_g && rot != 2 && g.march2) ? M2_VIRIAL : M2_FORCE;
    MarchArgs ak = a;
    if (!g.tail_in_kernel) ak.tail.enabled = 0;
    int rc = launch_march(h, ak, mode, rot, 0, false, write_g ? 1 : 0);
    if (rc == MM_OK) rc = finish_tail(h, a);
    g.fused_mask = g.fused ? ((rot == 2 ? 1 : 0) | (write_g ? 4 : 0)) : 0;
    g.tail_done = a.tail.enabled;
    if (rot == 2) g.cx ^= 1;
    return rc;
}

// Fused kick-drift-force-kick.  Reads the current x, v, g sets, writes the other x and v sets (and g when write_g)
// and flips them.  vm: pending velocity transform 0 none / 1 scalar / 2 matrix.  The halo planes of what was
// written are refreshed afterwards by the caller (sg_halo).
int sg_step(mm_handle *h, bool write_g, int vm, bool lean, const SgTail *tail) {
    SGrid &g = h->sg;
    MarchArgs a;
    fill_args(h, a);
    fill_tail(h, a, tail);
    fill_fused(h, a, g.cx ^ 1, g.cv ^ 1, g.cg ^ 1);
    MarchArgs ak = a;
    if (!g.tail_in_kernel) ak.tail.enabled = 0;
    int rc = launch_march(h, ak, M2_STEP, 0, vm, lean && vm != 2, write_g ? 1 : 0);
    if (rc == MM_OK) rc = finish_tail(h, a);
    g.fused_mask = g.fused ? (3 | (write_g ? 4 : 0)) : 0;
    g.tail_done = a.tail.enabled;
    g.cx ^= 1;
    g.cv ^= 1;
    if (write_g) g.cg ^= 1;
    return rc;
}

int sg_set_chunk(mm_handle *h, int chunk) {  // tuning: planes per block along z (0 = cost model)
    if (chunk < 0) return MM_ERR_INVALID;
    return sg_retile(h, chunk);
}

int sg_set_tile_rows(mm_handle *h, int rows) {  // warps per block
    SGrid &g = h->sg;
    if (!(rows == 8 || (g.march2_wanted && sg_march2_config_ok(g.rpt, rows)))) return MM_ERR_INVALID;
    g.tile_rows = rows;
    return sg_retile(h, 0);
}

int sg_set_rpt(mm_handle *h, int rpt) {  // node rows per thread of k_march2: fixed per model in the product library
    return rpt == h->sg.rpt ? MM_OK : MM_ERR_INVALID;
}

int sg_set_march2(mm_handle *h, int on) {
    SGrid &g = h->sg;
    g.march2_wanted = on ? 1 : 0;
    if (!on) g.tile_rows = 8;
    return sg_retile(h, 0);
}

}  // namespace mm
