// mm_march.cuh - the marching structured-grid kernel (see mm_structured.cu for the design notes)
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <type_traits>

#include "mm_reduce.cuh"
#include "mm_structured.cuh"

namespace mm {

constexpr int TX = 32;  // lanes along x: one warp per tile row
constexpr int kStages = 4;   // planes in flight in the staged (bulk-copy) variant
// MM_ABLATE (profiles/ablation.sh only, never in the product build): remove one ingredient of k_march to time the rest.
// 1 no barriers, 2 no shared-memory exchange (and no barriers), 4 no cell arithmetic, 8 no shuffles, 16 no global stores,
// 32 no global loads after the first plane.  Results are meaningless; only the launch time is looked at.
#ifndef MM_ABLATE
#define MM_ABLATE 0
#endif
// Constants of the SINGLE fast path (36 Bq + 6 c0 doubles = 84 uniform registers) do not fit the uniform register
// file next to the loop bookkeeping: ptxas then parks the overflow in vector registers and copies it back with two
// R2UR per double before EVERY use (ncu source page of variant 14: 87 R2UR + 15 MOV.SPILL of the 740 instructions of
// the two-plane loop body).  MM_PIN makes the last MM_PIN rows of Bq (and the matching c0 entries) ordinary per-thread
// register values instead, so that DFMA takes them as vector-register operands.  They are loaded from a copy of the
// constants in GLOBAL memory: anything ptxas can trace back to the parameter bank is "uniform" again and ends up in the
// same spill cycle (an opaque empty asm does not help: it leaves no instruction in the PTX).
#ifndef MM_PIN
#define MM_PIN 3
#endif
// resident blocks per SM the FORCE-only instantiations are compiled for (register budget 64K / (256 * blocks))
#ifndef MM_FORCE_BLOCKS
#define MM_FORCE_BLOCKS 1
#endif

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

// register-resident copy of the rows 6 - MM_PIN .. 5 of Bq and of the same entries of c0 (SINGLE path only)
struct PinnedConsts {
    double Bq[MM_PIN > 0 ? MM_PIN * 6 : 1];
    double c0[MM_PIN > 0 ? MM_PIN : 1];
};
__device__ __forceinline__ double ld_global_f64(const double *p) {
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void pin_consts(const SState *P, PinnedConsts &pc) {
#pragma unroll
    for (int i = 0; i < MM_PIN * 6; i++) pc.Bq[i] = ld_global_f64(&P->Bq[(6 - MM_PIN) * 6 + i]);
#pragma unroll
    for (int i = 0; i < MM_PIN; i++) pc.c0[i] = ld_global_f64(&P->c0[6 - MM_PIN + i]);
}
// sstate_eval with the pinned rows taken from registers
__device__ __forceinline__ void sstate_eval_pinned(const double d[6], const SState &P, const PinnedConsts &pc, double &e, double Sq[6]) {
#pragma unroll
    for (int I = 0; I < 6; I++) {
        const bool pin = I >= 6 - MM_PIN;
        const double *B = pin ? pc.Bq + (I - (6 - MM_PIN)) * 6 : P.Bq + I * 6;
        const double lo = fma(B[2], d[2], fma(B[1], d[1], B[0] * d[0]));
        Sq[I] = fma(B[5], d[5], fma(B[4], d[4], fma(B[3], d[3], lo)));
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

// Predicated store without a branch.  The plane loop must stay free of divergent branches: ptxas only keeps the
// constant loads (and the loop bookkeeping) on the uniform datapath where it can prove that the warp is converged.
__device__ __forceinline__ void st_if(bool pred, double *addr, double v) {
    if (MM_ABLATE & 16) pred = pred && v == 1.2345e301;  // ablation build: keep the value alive, never store
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %0, 0; @q st.global.f64 [%1], %2; }" ::"r"((int)pred), "l"(addr), "d"(v) : "memory");
}

// ---- bulk asynchronous copies global -> shared memory, completion on an mbarrier (PTX cp.async.bulk / mbarrier) ----------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
// one box of a rank-3 tensor map -> shared memory; completion (box bytes, zero fill included) on the mbarrier
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *map, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Delivery of one boundary-plane triple into a neighbour's halo plane (fused halo exchange).  Deliberately NOT inlined:
// it runs for 2 of the planes of a slab, and as inlined predicated stores it costs every plane 18 address computations
// and 18 predicated-off STG (measured: +10 % on the whole step).
__device__ __noinline__ void deliver3(double *d0, double *d1, double *d2, unsigned at, bool pred, double v0, double v1, double v2) {
    if (pred) {
        d0[at] = v0;
        d1[at] = v1;
        d2[at] = v2;
    }
}

// all states of a cell, Boltzmann-mixed (mmff.py:377-398); the mixing is linear in the gradient, hence in Sq.
template <bool SINGLE, bool WANT_VIR>
__device__ __forceinline__ void scell_eval(const double Hs[9], const SParams &kp, const PinnedConsts &pc, int type, double &e,
                                           double D[9], double vir[6]) {
    double Sq[6];
    if (SINGLE) {
        const SState &P = kp.st[0];
        double c0[6];
#pragma unroll
        for (int q = 0; q < 6; q++) c0[q] = (q >= 6 - MM_PIN) ? pc.c0[q - (6 - MM_PIN)] : P.c0[q];
        double d[6];
        d[0] = fma(Hs[2], Hs[2], fma(Hs[1], Hs[1], fma(Hs[0], Hs[0], -c0[0])));
        d[1] = fma(Hs[5], Hs[5], fma(Hs[4], Hs[4], fma(Hs[3], Hs[3], -c0[1])));
        d[2] = fma(Hs[8], Hs[8], fma(Hs[7], Hs[7], fma(Hs[6], Hs[6], -c0[2])));
        d[3] = fma(Hs[5], Hs[8], fma(Hs[4], Hs[7], fma(Hs[3], Hs[6], -c0[3])));
        d[4] = fma(Hs[2], Hs[8], fma(Hs[1], Hs[7], fma(Hs[0], Hs[6], -c0[4])));
        d[5] = fma(Hs[2], Hs[5], fma(Hs[1], Hs[4], fma(Hs[0], Hs[3], -c0[5])));
        sstate_eval_pinned(d, P, pc, e, Sq);  // efree: added once per owned column after the plane loop
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
//   SINGLE  one cell type with one metastable state (constants from the __constant__ slot, no type lookups)
//   ROT     FORCE only: 1 = positions are rotated on load (x_true = (x + shift) . Rpend); 2 = and written back
//   VM      STEP only: pending velocity transform  0 none, 1 scalar (Mvel[0]), 2 full 3x3
//   LEAN    no virial, kinetic-energy diagonal only (NVE / NVT steps whose pressure nobody looks at)
//   VAR     tuning bits (measured in profiles/): 1 = neighbouring rows synchronise pairwise through named barriers instead
//           of two block-wide barriers per plane; 2 = TMA loads (kStages planes in flight); 4 = two planes per trip;
//           8 = one synchronisation per plane (the gather of node plane p-2 is delayed by an iteration, see PIPE below)
//   TY      tile rows (warps per block); the tile owns (TX-2) x (TY-2) node columns
//
// Loads (VAR & 2).  The ablation in profiles/ shows where the time of the register-prefetch kernel goes: removing the
// loads saves 0.35 ms of 0.90, removing the stores 0.22 ms, removing the cell arithmetic nothing - with one block of
// eight warps per SM the waits do not overlap with anything.  The staged variant keeps kStages planes of the tile in
// flight with TMA: one tensor-map box (32 x 8 nodes of a padded plane, 2 KB) per field and plane, issued by one thread
// each, completion counted in bytes on an mbarrier per stage; the threads then read their node with LDS.  (A first
// version with one cp.async.bulk per field ROW - 88 copies of 256 B per plane - was slower than the register prefetch.)
//
// Shared-memory / shuffle traffic per plane and thread (the LSU pipe is as scarce as the FP64 pipe here: measured
// 1.0 / 2.0 / 2.5 cycles per warp instruction and SM for SHFL / STS.64 / LDS.64, profiles/microbench):
//   forward   y first, on the raw position (3 STS + 3 LDS), then x on the y-sum / y-difference (6 doubles by shuffle)
//   backward  x first, then y, adding rows of equal sign before each exchange: 6 doubles by shuffle, 3 STS + 3 LDS
template <int STEP, bool SINGLE, int ROT, int VM, bool LEAN, int VAR, int TY>
__global__ void __launch_bounds__(TX *TY, STEP ? 1 : MM_FORCE_BLOCKS)
k_march(const __grid_constant__ SParams kp, const __grid_constant__ MarchArgs a, const __grid_constant__ TmaMaps maps,
        const int write_g) {
    constexpr int OX = TX - 2, OY = TY - 2;
    constexpr bool PSYNC = (VAR & 1) != 0;   // pairwise named-barrier handshakes instead of block barriers
    constexpr bool TMA = (VAR & 2) != 0;     // node data staged by bulk async copies (NST planes in flight), see below
    constexpr int UNR = (VAR & 4) ? 2 : 1;   // planes per loop trip
    // PIPE: ONE block barrier per plane.  The gather of node plane p-2 (its second kick and stores) is delayed by an
    // iteration: its shared-memory input was published before the barrier of iteration p, so the region between two
    // barriers holds the exchange reads and shuffles of two planes, the stores of plane p-2 and the cell arithmetic of
    // layer p-1 as independent instruction streams.  sf / sb are double-buffered by plane parity.
    constexpr bool PIPE = (VAR & 8) != 0;
    constexpr int NB = PIPE ? 2 : 1;
    constexpr int NF = STEP ? 11 : 3;        // staged fields per node: x (3) [, v (3), g (3), m, 1/m]
    __shared__ double sf[NB][3][TY][TX];  // forward exchange along y: position of the row above
    __shared__ double sb[NB][3][TY][TX];  // backward exchange along y: x-combined gradient part of the row below
    __shared__ __align__(8) unsigned long long s_full[kStages];
    extern __shared__ __align__(128) double s_stage[];  // TMA: [kStages][NF][TY][kBoxW]

    const int lane = threadIdx.x, row = threadIdx.y;
    const int nx = a.nx, ny = a.ny, nxp = a.nxp;
    // thread = node column (k, l) = cell column with that origin vertex; k = -1 and l = -1 are the ghost column / row
    // of the padded planes, so no tile ever wraps.  Threads beyond the last ghost are clamped onto it (never owned).
    const int k = blockIdx.x * OX + lane - 1, l = blockIdx.y * OY + row - 1;
    const int kc = min(k, nx), lc = min(l, ny);
    const bool own_xy = lane >= 1 && lane <= OX && row >= 1 && row <= OY && k < nx && l < ny;
    const StepConsts &sc = *a.sc;
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
    // third domain vector in the frame of the positions this launch WRITES (rotated with them when ROT): the shift of the
    // boundary planes delivered across the periodic end of the slab ring
    double cw[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
        cw[j] = ROT ? fma(sc.rv[8], R[6 + j], fma(sc.rv[7], R[3 + j], sc.rv[6] * R[j])) : sc.rv[6 + j];
    const int rowp = (row + 1 < TY) ? row + 1 : row;
    const int rowm = (row > 0) ? row - 1 : row;

    // 32-bit element indices (the host refuses grids beyond 2^31 padded nodes per array): one IMAD.WIDE per address
    const unsigned plane = (unsigned)nxp * (unsigned)(ny + 2);
    const int c0 = 1 + blockIdx.z * a.chunk;
    const int c1 = min(c0 + a.chunk, a.nzl + 1);
    unsigned idx = ((unsigned)(c0 - 1) * (ny + 2) + lc + 1) * nxp + kc + kGhostX;  // node (k, l) in array plane p

    double acc[14];
#pragma unroll
    for (int i = 0; i < 14; i++) acc[i] = 0.0;
    PinnedConsts pc;
    if (SINGLE && MM_PIN > 0) pin_consts(&a.spg->st[0], pc);

    // carried from plane to plane
    double fpxy[3] = {0, 0, 0}, fdxy[3] = {0, 0, 0}, fpyd[3] = {0, 0, 0};  // forward: xy-combined sums / differences
    double Dp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};                            // D' of the previous cell layer
    double vh[3] = {0, 0, 0}, mprev = 0.0, hminv_prev = 0.0;  // STEP: half-kicked velocity / mass of the previous plane
    double vh2[3] = {0, 0, 0}, m2 = 0.0, hm2 = 0.0;           // PIPE: ... of the plane before that
    double gdc[3] = {0, 0, 0};                                 // PIPE: own-row gradient part of node plane p-2

    // software pipeline: raw loads of the NEXT plane are issued before the arithmetic of the current one (the arrays
    // carry one spare plane, so the prefetch of the last iteration stays inside the allocation)
    double nx_[3], nv_[3], ng_[3], nm_ = 0.0, nminv_ = 0.0;
    auto issue_loads = [&](unsigned at) {
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
    // staged variant: thread f < NF issues the box of field f (TX x TY nodes of one padded plane) for every staged plane
    const int tid = row * TX + lane;
    unsigned it = 0;  // planes done: plane p = c0 - 1 + it lives in stage it % kStages, barrier phase (it / kStages) & 1
    constexpr unsigned kStageDoubles = NF * TY * kBoxW;
    if (TMA) {
        if (tid == 0) {
            for (int st = 0; st < kStages; st++) mbar_init(smem_u32(&s_full[st]), 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    // plane q -> stage st.  ONE thread (lane 0 of the top apron row, whose own cells nobody uses) arms the barrier with the
    // bytes of all NF boxes and issues them back to back; with one box per thread the compiler serialises the issuing
    // lanes in an elected-lane loop that showed up with 6 % of the stall samples.
    auto stage_issue = [&](const int q, const unsigned st) {
        if (tid == (TY - 1) * TX) {
            const unsigned bar = smem_u32(&s_full[st]);
            mbar_arrive_expect(bar, NF * TY * kBoxW * 8u);
            // node k0 - 1 of the tile is padded column 30 bx + 1; the box starts one column earlier (even)
#pragma unroll
            for (int f = 0; f < NF; f++)
                tma_load_3d(smem_u32(s_stage) + (st * kStageDoubles + f * (TY * kBoxW)) * 8u, &maps.in[f], (int)(blockIdx.x * OX),
                            (int)(blockIdx.y * OY), q, bar);
        }
    };
    if (TMA) {
        for (int st = 0; st < kStages; st++)
            if (c0 - 1 + st <= c1) stage_issue(c0 - 1 + st, st);
    } else {
        issue_loads(idx);
    }

    // Completion of node plane q: gradient = row-below part (shared memory, published before the last barrier) + own part,
    // second kick (verlet.py:152-153), kinetic moments, stores (and delivery to the neighbour slab when q is a boundary plane)
    auto finish_node = [&](auto halo_tag, const int sbpar, const unsigned at, const int q, const double (&vhq)[3], const double hmq,
                           const double mq, const double (&gdq)[3]) {
        constexpr bool HALO = decltype(halo_tag)::value;
        double g[3];
#pragma unroll
        for (int j = 0; j < 3; j++) g[j] = ((MM_ABLATE & 2) ? gdq[j] * 1.25 : sb[sbpar][j][rowm][lane]) + gdq[j];
        if (STEP) {
            double vn[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                vn[j] = fma(-hmq, g[j], vhq[j]);
                st_if(own_xy, a.vo[j] + at, vn[j]);
            }
            if (HALO && a.fused && q == 1) deliver3(a.halo_lo[3], a.halo_lo[4], a.halo_lo[5], at, own_xy, vn[0], vn[1], vn[2]);
            if (HALO && a.fused && q == a.nzl) deliver3(a.halo_hi[3], a.halo_hi[4], a.halo_hi[5], at, own_xy, vn[0], vn[1], vn[2]);
            const double mo = own_xy ? mq : 0.0;
            const double mx = mo * vn[0], my = mo * vn[1], mz = mo * vn[2];
            acc[7] = fma(mx, vn[0], acc[7]);
            acc[8] = fma(my, vn[1], acc[8]);
            acc[9] = fma(mz, vn[2], acc[9]);
            if (!LEAN) {
                acc[10] = fma(my, vn[2], acc[10]);
                acc[11] = fma(mx, vn[2], acc[11]);
                acc[12] = fma(mx, vn[1], acc[12]);
            }
        }
        const bool pg = own_xy && write_g;
#pragma unroll
        for (int j = 0; j < 3; j++) st_if(pg, a.go[j] + at, g[j]);
        if (HALO && a.fused && q == 1) deliver3(a.halo_lo[6], a.halo_lo[7], a.halo_lo[8], at, pg, g[0], g[1], g[2]);
        if (HALO && a.fused && q == a.nzl) deliver3(a.halo_hi[6], a.halo_hi[7], a.halo_hi[8], at, pg, g[0], g[1], g[2]);
        if (!LEAN) acc[13] += own_xy ? fma(g[0], g[0], fma(g[1], g[1], g[2] * g[2])) : 0.0;
    };

    // One plane.  CELL: cell layer p-1 (planes p-1 and p) exists; NODE: node plane p-1 (cell layers p-2, p-1) is
    // completed.  The first two planes of a chunk are peeled (CELL / NODE false), so that the steady-state loop has no
    // uniform branches and its constant loads stay on the uniform datapath.
    auto plane_body = [&](auto cell_tag, auto node_tag, auto gath_tag, auto halo_tag, const int p) {
        constexpr bool CELL = decltype(cell_tag)::value, NODE = decltype(node_tag)::value;
        constexpr bool GATH = decltype(gath_tag)::value;  // PIPE: node plane p-2 is completed in this iteration
        constexpr bool HALO = decltype(halo_tag)::value;  // this plane may be a boundary plane of the slab (fused halo)
        const int par = PIPE ? (p & 1) : 0;
        double cx0, cx1, cx2, cv[3] = {0, 0, 0}, cg[3] = {0, 0, 0}, cm = 0.0, cminv = 0.0;
        const unsigned st = it % kStages;
        if (TMA) {
            if (!(MM_ABLATE & 32) || it < kStages) mbar_wait(smem_u32(&s_full[st]), (it / kStages) & 1u);
            // the box starts at padded column 30 bx, padded row 6 by: thread (lane, row) reads slot (lane + 1, row); slots
            // beyond the array were filled with zeros (never owned)
            const double *sp = s_stage + st * kStageDoubles + row * kBoxW + lane + 1;
            cx0 = sp[0];
            cx1 = sp[TY * kBoxW];
            cx2 = sp[2 * TY * kBoxW];
            if (STEP) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    cv[d] = sp[(3 + d) * TY * kBoxW];
                    cg[d] = sp[(6 + d) * TY * kBoxW];
                }
                cm = sp[9 * TY * kBoxW];
                cminv = sp[10 * TY * kBoxW];
            }
        } else {
            cx0 = nx_[0];
            cx1 = nx_[1];
            cx2 = nx_[2];
            cm = nm_;
            cminv = nminv_;
            if (STEP) {
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    cv[d] = nv_[d];
                    cg[d] = ng_[d];
                }
            }
            if (!(MM_ABLATE & 32)) issue_loads(idx + plane);
        }

        // ---- node (lane, row, p): true position (and, in STEP mode, kick + drift: verlet.py:144-146) -------------
        const double xs = cx0, ys = cx1, zs = cx2;  // ghosts already carry their periodic shift
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
        if ((STEP || ROT == 2) && CELL) {
            {
                const bool px = own_xy && p < c1;
#pragma unroll
                for (int j = 0; j < 3; j++) st_if(px, a.xo[j] + idx, r[j]);
                if (HALO && a.fused && p == 1)
                    deliver3(a.halo_lo[0], a.halo_lo[1], a.halo_lo[2], idx, px, fma(a.wrap_lo, cw[0], r[0]),
                             fma(a.wrap_lo, cw[1], r[1]), fma(a.wrap_lo, cw[2], r[2]));
                if (HALO && a.fused && p == a.nzl)
                    deliver3(a.halo_hi[0], a.halo_hi[1], a.halo_hi[2], idx, px, fma(a.wrap_hi, cw[0], r[0]),
                             fma(a.wrap_hi, cw[1], r[1]), fma(a.wrap_hi, cw[2], r[2]));
            }
        }

        // ---- forward butterfly: y through shared memory, x by shuffle, z in registers -------------------------------
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (!(MM_ABLATE & 2)) sf[par][j][row][lane] = r[j];
        if (MM_ABLATE & 3) {
        } else if (PSYNC && !PIPE) {  // row r only needs row r+1: producer arrives, consumer waits (barrier ids 1 .. TY-1)
            if (row > 0) bar_arrive(row, 2 * TX);
            if (row + 1 < TY) bar_wait(row + 1, 2 * TX);
        } else {
            __syncthreads();
        }
        if (TMA && !(MM_ABLATE & 32) && p + kStages <= c1) stage_issue(p + kStages, st);  // all threads have read stage st
        double pxy[3], dxy[3], pyd[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double rn = (MM_ABLATE & 2) ? r[j] * 1.25 : sf[par][j][rowp][lane];
            const double py = rn + r[j], dy = rn - r[j];
            const double pyn = (MM_ABLATE & 8) ? py * 1.5 : __shfl_down_sync(0xffffffffu, py, 1);
            const double dyn = (MM_ABLATE & 8) ? dy * 0.75 : __shfl_down_sync(0xffffffffu, dy, 1);
            pxy[j] = py + pyn;   // sum over the four nodes of the cell face in this plane
            dxy[j] = pyn - py;   // x difference of the y sums
            pyd[j] = dy + dyn;   // y difference of the x sums
        }

        if (PIPE && GATH) finish_node(halo_tag, par ^ 1, idx - 2 * plane, p - 2, vh2, hm2, m2, gdc);

        double D[9];
        if (CELL) {
            double Hs[9];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                Hs[j] = fdxy[j] + dxy[j];        // 4 * (mean x edge)
                Hs[3 + j] = fpyd[j] + pyd[j];    // 4 * (mean y edge)
                Hs[6 + j] = pxy[j] - fpxy[j];    // 4 * (mean z edge)
            }
            const int type = SINGLE ? 0 : (int)a.type[idx - plane];
            double e, vir[6];
            if (MM_ABLATE & 4) {
                e = Hs[0];
#pragma unroll
                for (int q = 0; q < 9; q++) D[q] = Hs[q];
#pragma unroll
                for (int q = 0; q < 6; q++) vir[q] = Hs[q];
            } else {
                scell_eval<SINGLE, !LEAN>(Hs, kp, pc, type, e, D, vir);
            }
            if (NODE) {  // the warm-up layer c0-1 belongs to the chunk below
                acc[0] += own_xy ? e : 0.0;
                if (!LEAN) {
#pragma unroll
                    for (int q = 0; q < 6; q++) acc[1 + q] += own_xy ? vir[q] : 0.0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            fpxy[j] = pxy[j];
            fdxy[j] = dxy[j];
            fpyd[j] = pyd[j];
        }

        // ---- backward butterfly: z in registers, x by shuffle, y through shared memory -----------------------------
        // With p0, p1, p2 the z-combined rows of a cell column, the node gathers  (+ lane-1, - lane) of p0, (+ row-1, - row)
        // of p1 and all four p2.  Adding what has equal sign BEFORE each exchange leaves 2 doubles per component for the
        // shuffle (p0 + p2 and p1) and 1 for shared memory (ya + yb):
        //   ya = (p0 + p2)[lane-1] + (p2 - p0),  yb = p1[lane-1] + p1,   g = (ya + yb)[row-1] + (ya - yb)
        double gd[3];  // ya - yb: the own row's part of the gradient
        if (NODE) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double p0 = Dp[j] + D[j], p1 = Dp[3 + j] + D[3 + j], p2 = Dp[6 + j] - D[6 + j];
                const double s02m = (MM_ABLATE & 8) ? (p0 + p2) * 1.5 : __shfl_up_sync(0xffffffffu, p0 + p2, 1);
                const double p1m = (MM_ABLATE & 8) ? p1 * 0.75 : __shfl_up_sync(0xffffffffu, p1, 1);
                const double ya = s02m + (p2 - p0), yb = p1m + p1;
                if (!(MM_ABLATE & 2)) sb[par][j][row][lane] = ya + yb;
                gd[j] = ya - yb;
            }
        }
        if (!PIPE) {
            if (MM_ABLATE & 3) {
            } else if (PSYNC) {  // row r only needs row r-1 (barrier ids TY .. 2 TY - 2); these two handshakes per plane also
                          // order the reuse of sf / sb between planes
                if (row + 1 < TY) bar_arrive(TY + row, 2 * TX);
                if (row > 0) bar_wait(TY + row - 1, 2 * TX);
            } else {
                __syncthreads();
            }
            if (NODE) finish_node(halo_tag, 0, idx - plane, p - 1, vh, hminv_prev, mprev, gd);
        } else {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                gdc[j] = gd[j];
                vh2[j] = vh[j];
            }
            m2 = mprev;
            hm2 = hminv_prev;
        }
        if (CELL) {
#pragma unroll
            for (int q = 0; q < 9; q++) Dp[q] = D[q];
        }
        if (STEP) {
#pragma unroll
            for (int j = 0; j < 3; j++) vh[j] = vcur[j];
            mprev = cm;
            hminv_prev = hminv;
        }
        idx += plane;
        it++;
    };

    // The boundary planes of a slab (1, nzl: positions; one iteration later: velocities, gradients) can only come up in the
    // first three and the last two iterations of a chunk: only those bodies carry the delivery code of the fused halo.
    constexpr std::true_type T{};
    constexpr std::false_type F{};
    plane_body(F, F, F, T, c0 - 1);
    plane_body(T, F, F, T, c0);
    if (!PIPE) {
        plane_body(T, T, T, T, c0 + 1);
#pragma unroll UNR
        for (int p = c0 + 2; p <= c1 - 2; p++) plane_body(T, T, T, F, p);
#pragma unroll 1
        for (int p = max(c0 + 2, c1 - 1); p <= c1; p++) plane_body(T, T, T, T, p);
    } else {
        plane_body(T, T, F, T, c0 + 1);
        if (c0 + 2 <= c1) plane_body(T, T, T, T, c0 + 2);  // completes node plane c0 (the lower boundary plane of the first chunk)
#pragma unroll UNR
        for (int p = c0 + 3; p <= c1 - 2; p++) plane_body(T, T, T, F, p);
#pragma unroll 1
        for (int p = max(c0 + 3, c1 - 1); p <= c1; p++) plane_body(T, T, T, T, p);
        // drain: node plane c1-1 (its row-below part was published in iteration c1; after the rotation vh2 / gdc are its data)
        __syncthreads();
        finish_node(T, c1 & 1, idx - 2 * plane, c1 - 1, vh2, hm2, m2, gdc);
    }

    if (SINGLE && own_xy) acc[0] = fma((double)(c1 - c0), kp.st[0].efree, acc[0]);
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
