// mm_md.cu - device-resident velocity Verlet + Nose-Hoover chain + MTK barostat.
//
// Stands in for (relative to /root/reference)
//   VerletIntegrator.initialize / propagate / compute_properties   micmec/sampling/verlet.py:119-190
//   ConsErrTracker                                                  micmec/sampling/verlet.py:275-307
//   NHChain.__call__ / set_ndof / get_econs_correction              micmec/sampling/nvt.py:393-458
//   NHCThermostat.init/pre/post                                     micmec/sampling/nvt.py:507-532
//   MTKBarostat.init/pre/post/baro/add_press_cont                   micmec/sampling/npt.py:579-757
//   TBCombination.pre/post                                          micmec/sampling/npt.py:99-148
//
// Design: the whole step runs on one CUDA stream without any host round trip.  All O(1) algebra of the
// thermostat / barostat (chain sweep, 3x3 eigen-decompositions, conserved-quantity bookkeeping) lives in a
// single-block "scalar" kernel that reads the block partials of the node / cell kernels and updates an MDState
// record in device memory.  Global velocity scalings and rotations are NOT applied to the velocity array when
// the hooks ask for them: they are accumulated in a pending 3x3 matrix (MDState::Mvel) and in the algebraically
// transformed second-moment tensor sum m v(x)v, and are applied by the next kernel that touches the
// velocities anyway (the kick).  That removes every velocity-only pass of the reference
// (vel *= factor, vel = vel @ rot_mat, _compute_ekin).
#include <cmath>
#include <cstdlib>
#include <new>
#include <vector>

#include "mm_internal.h"
#include "mm_reduce.cuh"
#include "mm_scalar.cuh"

namespace mm {

// One block.  Sums the block partials it is told to consume, then thread 0 runs the requested sub-steps in the
// canonical order RESET_MVEL, TAKE_FORCE, TAKE_KIN, BARO_B, THERMO, BARO_A, ECONS, PROPS (see md_step below).
__global__ void __launch_bounds__(256)
k_scalar(MDState *st, double *rvecs_dev, StepConsts *sc, unsigned ops, const double *pc, int nbc, const double *pn, int nbn,
         const double *pd, int nbd, double n3, const double *pl, int nbl) {
    double fr[7] = {0, 0, 0, 0, 0, 0, 0}, kn[7] = {0, 0, 0, 0, 0, 0, 0}, dl[1] = {0}, lg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if ((ops & (OP_LANG_A | OP_LANG_B)) && nbl > 0) partials_sum<8>(pl, nbl, kRedSlots, lg);
    if (pn == pc + 7 && nbn == nbc && nbc > 1 && (ops & (OP_TAKE_KIN | OP_TAKE_FORCE))) {
        double all[16];  // structured path: both halves live in the same 16-slot records
        partials_sum16(pc, nbc, all);
        for (int k = 0; k < 7; k++) {
            fr[k] = all[k];
            kn[k] = all[7 + k];
        }
    } else {
        if (ops & OP_TAKE_FORCE) partials_sum<7>(pc, nbc, kRedSlots, fr);
        if ((ops & (OP_TAKE_KIN | OP_TAKE_FORCE)) && nbn > 0) partials_sum<7>(pn, nbn, kRedSlots, kn);
    }
    if ((ops & OP_TAKE_DELTA) && nbd > 0) partials_sum<1>(pd, nbd, kRedSlots, dl);
    // Work on a shared-memory copy of the state: the serial algebra below touches a few hundred fields, and as
    // dependent global-memory accesses they cost ~30 us per launch; from shared memory it is a few.
    __shared__ MDState sm_state;
    static_assert(sizeof(MDState) % sizeof(double) == 0, "MDState is copied as doubles");
    constexpr int kWords = sizeof(MDState) / sizeof(double);
    for (int i = threadIdx.x; i < kWords; i += blockDim.x)
        reinterpret_cast<double *>(&sm_state)[i] = reinterpret_cast<const double *>(st)[i];
    __syncthreads();
    if (threadIdx.x == 0) scalar_ops(sm_state, rvecs_dev, sc, ops, fr, kn, dl, nbn, n3, lg);
    __syncthreads();
    for (int i = threadIdx.x; i < kWords; i += blockDim.x)
        reinterpret_cast<double *>(st)[i] = reinterpret_cast<const double *>(&sm_state)[i];
}

// --------------------------------------------------------------------------------------- node kernels --------
constexpr int kNodeThreads = 256;
constexpr int kNodeThreadsL = 256;

// ---- Langevin thermostat on the device (nvt.py:165-218) ----------------------------------------------------------------
// v <- c1 v + c2 xi with c1 = exp(-dt / (2 timecon)), c2 = sqrt((1 - c1^2) kB T / m), xi ~ N(0, 1) per component.
// xi comes from Philox4x32-10 (Salmon et al., SC'11) keyed by the seed and counted by (reference node id, half-step
// number, draw): a node gets the same kick wherever a copy of it lives - ghost nodes, halo planes of a z-slab, the AoS
// arrays of the indexed kernels - so no exchange follows the update, and structured and indexed runs of one seed agree.
// three standard normal deviates of node `gid` at half-step `phase` (Box-Muller on 64-bit uniforms)
__device__ __forceinline__ void langevin_noise(unsigned long long seed, long long gid, long long phase, double (&xi)[3]) {
    double z[4];
#pragma unroll
    for (int draw = 0; draw < 2; draw++) {
        unsigned r[4];
        philox4x32_10((unsigned)gid, (unsigned)((unsigned long long)gid >> 32), (unsigned)phase,
                      (unsigned)(((unsigned long long)phase >> 32) << 1) | (unsigned)draw, (unsigned)seed, (unsigned)(seed >> 32), r);
        const double u1 = ((double)(((unsigned long long)r[0] << 32) | r[1]) + 0.5) * 5.421010862427522e-20;  // 2^-64
        const double u2 = ((double)(((unsigned long long)r[2] << 32) | r[3]) + 0.5) * 5.421010862427522e-20;
        const double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        z[2 * draw] = rad * cs;
        z[2 * draw + 1] = rad * sn;
    }
    xi[0] = z[0];
    xi[1] = z[1];
    xi[2] = z[2];
}

struct LangArgs {
    double *v[3];        // structured: SoA planes (all padded nodes are updated, copies included); indexed: v[0] = vel [n][3]
    const double *m;     // node masses in the same layout
    int structured;
    int nx, ny, nzl, nxp, z0, nz_global;
    long long plane, n;  // padded plane / number of entries to visit
    double kT;           // kB T
    int nphase;          // 1: one half-step; 2: a "post" followed by the next step's "pre"
    int phase_offset;    // half-step number = 2 * (steps completed) + offset: 0 for a lone "pre", 1 for a "post"
};

__global__ void __launch_bounds__(kNodeThreadsL)
k_langevin(const __grid_constant__ LangArgs a, const MDState *__restrict__ st, double *__restrict__ partials) {
    const double c1 = exp(-st->timestep / st->lg_timecon / 2.0);
    const double var = (1.0 - c1 * c1) * a.kT;
    const long long phase0 = 2 * st->counter + a.phase_offset;
    const unsigned long long seed = st->lg_seed;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
        long long gid;
        bool owned = true;
        double vx, vy, vz, mass;
        if (a.structured) {
            const int p = (int)(i / a.plane);
            const long long rem = i - (long long)p * a.plane;
            const int lrow = (int)(rem / a.nxp), kcol = (int)(rem - (long long)lrow * a.nxp);
            const int k = kcol - kGhostX, l = lrow - 1;
            if (k < -1 || k > a.nx || l > a.ny) continue;  // pad columns
            const int kw = (k + a.nx) % a.nx, lw = (l + a.ny) % a.ny, pz = (a.z0 + p - 1 + a.nz_global) % a.nz_global;
            gid = ((long long)kw * a.ny + lw) * a.nz_global + pz;  // reference id (micmec/utils.py:226)
            owned = k >= 0 && k < a.nx && l >= 0 && l < a.ny && p >= 1 && p <= a.nzl;
            vx = a.v[0][i];
            vy = a.v[1][i];
            vz = a.v[2][i];
            mass = a.m[i];
        } else {
            gid = i;
            vx = a.v[0][3 * i];
            vy = a.v[0][3 * i + 1];
            vz = a.v[0][3 * i + 2];
            mass = a.m[i];
        }
        if (!(mass > 0.0)) continue;  // unused padded entries
        const double c2 = sqrt(var / mass), w = owned ? mass : 0.0;
        acc[0] += w * (vx * vx + vy * vy + vz * vz);
        double xi[3];
        langevin_noise(seed, gid, phase0, xi);
        vx = fma(c2, xi[0], c1 * vx);
        vy = fma(c2, xi[1], c1 * vy);
        vz = fma(c2, xi[2], c1 * vz);
        acc[1] += w * vx * vx;
        acc[2] += w * vy * vy;
        acc[3] += w * vz * vz;
        acc[4] += w * vy * vz;
        acc[5] += w * vx * vz;
        acc[6] += w * vx * vy;
        if (a.nphase == 2) {
            langevin_noise(seed, gid, phase0 + 1, xi);
            vx = fma(c2, xi[0], c1 * vx);
            vy = fma(c2, xi[1], c1 * vy);
            vz = fma(c2, xi[2], c1 * vz);
            acc[7] += w * (vx * vx + vy * vy + vz * vz);
        }
        if (a.structured) {
            a.v[0][i] = vx;
            a.v[1][i] = vy;
            a.v[2][i] = vz;
        } else {
            a.v[0][3 * i] = vx;
            a.v[0][3 * i + 1] = vy;
            a.v[0][3 * i + 2] = vz;
        }
    }
    block_sum_store<8>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

// v <- v.Mvel - (dt/2) g/m ; x <- x + dt v      (hook scalings + verlet.py:144-146); optional posold snapshot
__global__ void __launch_bounds__(kNodeThreads)
k_kick_drift(const MDState *__restrict__ st, double *__restrict__ pos, double *__restrict__ vel,
             const double *__restrict__ gpos, const double *__restrict__ masses, double *__restrict__ posold, int64_t n) {
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = st->Mvel[i];
    const double dt = st->timestep;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const double im = 1.0 / masses[i];
        double w[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double vt = vx * M[j] + vy * M[3 + j] + vz * M[6 + j];
            const double acc = -gpos[3 * i + j] * im;
            w[j] = vt + 0.5 * acc * dt;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double x = pos[3 * i + j];
            if (posold) posold[3 * i + j] = x;
            pos[3 * i + j] = x + dt * w[j];
            vel[3 * i + j] = w[j];
        }
    }
}

// gather the node gradient (mmff.py:303-318), second kick (verlet.py:152-153), kinetic moments + sum g^2
__global__ void __launch_bounds__(kNodeThreads)
k_gather_kick2(const MDState *__restrict__ st, const int32_t *__restrict__ node_cells, const double *__restrict__ gcell,
               int64_t nnodes, int64_t ncells, double *__restrict__ gpos, double *__restrict__ vel,
               const double *__restrict__ masses, double *__restrict__ partials) {
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const double dt = st->timestep;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t c = node_cells[(int64_t)j * nnodes + n];
            if (c >= 0) {
                g[0] += gcell[(int64_t)(3 * j) * ncells + c];
                g[1] += gcell[(int64_t)(3 * j + 1) * ncells + c];
                g[2] += gcell[(int64_t)(3 * j + 2) * ncells + c];
            }
        }
        const double m = masses[n], im = 1.0 / m;
        double v[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            gpos[3 * n + j] = g[j];
            const double a = -g[j] * im;
            v[j] = vel[3 * n + j] + 0.5 * a * dt;
            vel[3 * n + j] = v[j];
        }
        acc[0] += m * v[0] * v[0];
        acc[1] += m * v[1] * v[1];
        acc[2] += m * v[2] * v[2];
        acc[3] += m * v[1] * v[2];
        acc[4] += m * v[0] * v[2];
        acc[5] += m * v[0] * v[1];
        acc[6] += g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    }
    block_sum_store<7>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

// gradient gather without kick (barostat force calls, npt.py:702-707) + sum g^2 in slot 6
__global__ void __launch_bounds__(kNodeThreads)
k_gather_g2(const int32_t *__restrict__ node_cells, const double *__restrict__ gcell, int64_t nnodes, int64_t ncells,
            double *__restrict__ gpos, double *__restrict__ partials) {
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        double g[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t c = node_cells[(int64_t)j * nnodes + n];
            if (c >= 0) {
                g[0] += gcell[(int64_t)(3 * j) * ncells + c];
                g[1] += gcell[(int64_t)(3 * j + 1) * ncells + c];
                g[2] += gcell[(int64_t)(3 * j + 2) * ncells + c];
            }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) gpos[3 * n + j] = g[j];
        acc[6] += g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
    }
    block_sum_store<7>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

// kinetic moments of the stored velocities (initialisation)
__global__ void __launch_bounds__(kNodeThreads)
k_moments(const double *__restrict__ vel, const double *__restrict__ masses, const double *__restrict__ gpos, int64_t n,
          double *__restrict__ partials) {
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double m = masses[i], vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        acc[0] += m * vx * vx;
        acc[1] += m * vy * vy;
        acc[2] += m * vz * vz;
        acc[3] += m * vy * vz;
        acc[4] += m * vx * vz;
        acc[5] += m * vx * vy;
        if (gpos) acc[6] += gpos[3 * i] * gpos[3 * i] + gpos[3 * i + 1] * gpos[3 * i + 1] + gpos[3 * i + 2] * gpos[3 * i + 2];
    }
    block_sum_store<7>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

// x <- x.Rpos (npt.py:689-698); optional posold snapshot of the un-rotated positions
__global__ void __launch_bounds__(kNodeThreads)
k_apply_pos(const MDState *__restrict__ st, double *__restrict__ pos, double *__restrict__ posold, int64_t n) {
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = st->Rpos[i];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
        if (posold) {
            posold[3 * i] = x;
            posold[3 * i + 1] = y;
            posold[3 * i + 2] = z;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) pos[3 * i + j] = x * R[j] + y * R[3 + j] + z * R[6 + j];
    }
}

// v <- v.Mvel : make the stored velocities the true ones (before a state read-back)
__global__ void __launch_bounds__(kNodeThreads)
k_flush_vel(const MDState *__restrict__ st, double *__restrict__ vel, int64_t n) {
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = st->Mvel[i];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = vel[3 * i], y = vel[3 * i + 1], z = vel[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 3; j++) vel[3 * i + j] = x * M[j] + y * M[3 + j] + z * M[6 + j];
    }
}

// sum (pos - posold)^2  (verlet.py:158-161, 173)
__global__ void __launch_bounds__(kNodeThreads)
k_delta(const double *__restrict__ pos, const double *__restrict__ posold, int64_t n3, double *__restrict__ partials) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = pos[i] - posold[i];
        acc[0] += d * d;
    }
    block_sum_store<1>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

}  // namespace mm

using namespace mm;

struct mm_md {
    mm_handle *h = nullptr;
    mm_md_desc desc;
    MDState *d_state = nullptr;
    MDState *h_state = nullptr;  // pinned mirror
    double *d_vel = nullptr, *d_masses = nullptr, *d_posold = nullptr;
    double *d_pkin = nullptr;    // partials of node kernels [kMaxRedBlocks][kRedSlots]
    double *d_pdelta = nullptr;  // partials of k_delta
    double *d_plang = nullptr;   // partials of k_langevin
    int nblang = 0;              // ... pending for the next scalar launch
    bool initialised = false;
    bool structured = false;  // state lives in the SoA planes of h->sg
    // CUDA graphs of TWO consecutive lean steps (two, because the ping-pong buffers return to the same parity after
    // two steps), one per buffer-parity state.  A 64^3 NVE step is 26 us of kernel time but 4 launches with
    // multi-kilobyte parameter blocks; replaying a captured pair removes the per-launch host cost.
    cudaGraphExec_t gexec[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t glaunches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t steps_done = 0;
    int use_graphs = 1;
    // Lean steps on the ghost-node path (one GPU, grids above 128^3 nodes): the x / y ghost fill that follows a marching launch
    // and the scalar kernel that follows it too do not depend on each other and can run side by side (ghost fill on a side
    // stream, fork / join by events, captured into the step graphs).  Measured (profiles/r02, calls ag / ah): 1.3547-1.3615 ms
    // per NPT step at 256^3 against 1.3568-1.3589 in stream order - the two kernels are ~12-15 us each and the extra graph
    // edges cost what the overlap saves.  Off; MICMEC_B200_HALO_OVERLAP=1 turns it on.
    int overlap_halo = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace mm {

static int scalar_launch(mm_md *md, unsigned ops, int nbc, int nbn, int nbd) {
    mm_handle *h = md->h;
    const int nbl = md->nblang;  // sums of a k_langevin launch waiting to be booked (OP_LANG_A / OP_LANG_B)
    md->nblang = 0;
    const bool sg = md->structured;
    // structured kernels leave all 14 sums of a block in one partial: energy + virial in slots 0-6, moments + g^2 in 7-13
    const double *pc = sg ? h->sg.d_partials : h->d_partials;
    const double *pn = sg ? h->sg.d_partials + 7 : md->d_pkin;
    const double *pd = md->d_pdelta;
    if (h->slab_count > 1 && (nbc > 0 || nbn > 0 || nbd > 0)) {
        // z-slabs: sum the local partials, all-reduce the 16 doubles, and hand the scalar kernel ONE "block" of sums.
        // Every rank then runs the same scalar algebra on bit-identical inputs.
        const int rc = comm_reduce_partials(h, pc, nbc, pn, nbn, pd, nbd);
        if (rc != MM_OK) return rc;
        pc = h->d_red;
        pn = h->d_red + 7;
        pd = h->d_red + 14;
        nbc = nbc > 0 ? 1 : 0;
        nbn = nbn > 0 ? 1 : 0;
        nbd = nbd > 0 ? 1 : 0;
    }
    k_scalar<<<1, 256, 0, h->stream>>>(md->d_state, h->d_rvecs, sg ? h->sg.d_sc : nullptr, ops, pc, nbc, pn, nbn, pd, nbd,
                                       3.0 * (double)h->nnodes_global, md->d_plang, nbl);
    h->launches++;
    return MM_OK;
}

// Langevin half-step(s) on the current velocities (nvt.py:199-218); the sums wait in d_plang for the next scalar launch
static int lang_launch(mm_md *md, int nphase, int phase_offset) {
    mm_handle *h = md->h;
    LangArgs a;
    memset(&a, 0, sizeof(a));
    a.nphase = nphase;
    a.phase_offset = phase_offset;
    a.kT = h->boltzmann * md->desc.langevin_temp;
    if (md->structured) {
        const SGrid &g = h->sg;
        a.structured = 1;
        for (int d = 0; d < 3; d++) a.v[d] = g.v[g.cv][d];
        a.m = g.m;
        a.nx = g.nx;
        a.ny = g.ny;
        a.nzl = g.nzl;
        a.nxp = g.nxp;
        a.plane = g.plane;
        a.z0 = h->slab_rank * g.nzl;
        a.nz_global = g.nzl * (h->slab_count > 1 ? h->slab_count : 1);
        a.n = g.plane * (g.nzl + 2);
    } else {
        a.v[0] = md->d_vel;
        a.m = md->d_masses;
        a.n = h->nnodes;
    }
    const int gl = grid_for(h, a.n, kNodeThreadsL);
    k_langevin<<<gl, kNodeThreadsL, 0, h->stream>>>(a, md->d_state, md->d_plang);
    h->launches++;
    md->nblang = gl;
    return MM_OK;
}

// one barostat force call: x <- x.Rpos, cells, gather (npt.py:683-707)
static int baro_force(mm_md *md, bool snapshot, int &nbc, int &nbn) {
    mm_handle *h = md->h;
    const int gn = grid_for(h, h->nnodes, kNodeThreads);
    k_apply_pos<<<gn, kNodeThreads, 0, h->stream>>>(md->d_state, h->d_pos, snapshot ? md->d_posold : nullptr, h->nnodes);
    h->launches++;
    nbc = cells_launch(h);
    k_gather_g2<<<gn, kNodeThreads, 0, h->stream>>>(h->d_node_cells, h->d_gcell, h->nnodes, h->ncells, h->d_gpos, md->d_pkin);
    h->launches++;
    nbn = gn;
    return MM_OK;
}

// VerletIntegrator.propagate (verlet.py:140-166).  `full` additionally produces rmsd_delta for this step.
// own_pre: launch this step's first scalar "pre" call (false when the previous step merged it into its last launch);
// merge_next: append the NEXT step's first "pre" call to this step's last scalar launch (only between lean steps).
static int md_step(mm_md *md, bool full, bool own_pre, bool merge_next) {
    mm_handle *h = md->h;
    const bool thermo = md->desc.has_thermo != 0, baro = md->desc.has_baro != 0, lang = md->desc.has_langevin != 0;
    const unsigned post_thermo = (thermo && md->desc.thermo_kind == 0) ? OP_THERMO : 0u;  // Berendsen acts in "pre" only
    const int gn = grid_for(h, h->nnodes, kNodeThreads);
    int nbc = 0, nbn = 0;
    // ---- "pre" hooks: TBCombination.pre = barostat, then thermostat (npt.py:99-115) ----
    if (baro) {
        if (own_pre) scalar_launch(md, OP_BARO_A, 0, 0, 0);
        baro_force(md, full, nbc, nbn);
        scalar_launch(md, OP_TAKE_FORCE | OP_BARO_B | (thermo ? OP_THERMO : 0u), nbc, nbn, 0);
    } else if (thermo) {
        if (own_pre) scalar_launch(md, OP_THERMO, 0, 0, 0);
    } else if (lang && own_pre) {
        lang_launch(md, 1, 0);
        scalar_launch(md, OP_LANG_A, 0, 0, 0);
    }
    const unsigned next_op = !merge_next ? 0u : baro ? OP_NEXT_BARO_A : thermo ? OP_NEXT_THERMO : lang ? OP_LANG_B : 0u;
    // ---- velocity Verlet (verlet.py:144-154) ----
    k_kick_drift<<<gn, kNodeThreads, 0, h->stream>>>(md->d_state, h->d_pos, md->d_vel, h->d_gpos, md->d_masses,
                                                     (full && !baro) ? md->d_posold : nullptr, h->nnodes);
    h->launches++;
    nbc = cells_launch(h);
    k_gather_kick2<<<gn, kNodeThreads, 0, h->stream>>>(md->d_state, h->d_node_cells, h->d_gcell, h->nnodes, h->ncells,
                                                       h->d_gpos, md->d_vel, md->d_masses, md->d_pkin);
    h->launches++;
    // ---- "post" hooks: thermostat, then barostat (npt.py:117-148), then verlet.py:158-166 ----
    unsigned ops = OP_RESET_MVEL | OP_TAKE_FORCE | OP_TAKE_KIN | post_thermo;
    int nbd = 0;
    if (!baro) {
        if (full) {
            nbd = grid_for(h, 3 * h->nnodes, kNodeThreads);
            k_delta<<<nbd, kNodeThreads, 0, h->stream>>>(h->d_pos, md->d_posold, 3 * h->nnodes, md->d_pdelta);
            h->launches++;
            ops |= OP_TAKE_DELTA;
        }
        if (lang) {  // "post" (and, between lean steps, the next step's "pre" in the same pass)
            lang_launch(md, merge_next ? 2 : 1, 1);
            ops |= OP_LANG_A;
        }
        scalar_launch(md, ops | OP_ECONS | OP_ADVANCE | OP_PROPS | next_op, nbc, gn, nbd);
    } else {
        scalar_launch(md, ops | OP_BARO_A, nbc, gn, 0);
        baro_force(md, false, nbc, nbn);
        ops = OP_TAKE_FORCE | OP_BARO_B | OP_ECONS | OP_ADVANCE | OP_PROPS | next_op;
        if (full) {
            nbd = grid_for(h, 3 * h->nnodes, kNodeThreads);
            k_delta<<<nbd, kNodeThreads, 0, h->stream>>>(h->d_pos, md->d_posold, 3 * h->nnodes, md->d_pdelta);
            h->launches++;
            ops |= OP_TAKE_DELTA;
        }
        scalar_launch(md, ops, nbc, nbn, nbd);
    }
    return MM_OK;
}

// arr[n][3] <- arr[n][3] . M   with M (3x3) in device memory
__global__ void __launch_bounds__(kNodeThreads)
k_apply_mat9(const double *__restrict__ M9, double *__restrict__ arr, int64_t n) {
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) M[i] = M9[i];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = arr[3 * i], y = arr[3 * i + 1], z = arr[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 3; j++) arr[3 * i + j] = x * M[j] + y * M[3 + j] + z * M[6 + j];
    }
}

// true positions / velocities of the structured state as reference-ordered AoS arrays (h->d_pos, md->d_vel)
static void sg_export(mm_md *md, bool pos, bool vel, double *pos_dst) {
    mm_handle *h = md->h;
    const int gn = grid_for(h, h->nnodes, kNodeThreads);
    const double *sc = reinterpret_cast<const double *>(h->sg.d_sc);
    if (pos) {
        sg_to_aos(h, 0, pos_dst);
        k_apply_mat9<<<gn, kNodeThreads, 0, h->stream>>>(sc, pos_dst, h->nnodes);  // StepConsts::Rpend
        h->launches++;
    }
    if (vel) {
        sg_to_aos(h, 1, md->d_vel);
        k_apply_mat9<<<gn, kNodeThreads, 0, h->stream>>>(sc + 9, md->d_vel, h->nnodes);  // StepConsts::Mvel
        h->launches++;
    }
}

// The same step on the structured-grid kernels (mm_structured.cu): one fused kick-drift-force-kick launch, plus one
// force-only launch per barostat call.  Pending rotations / scalings are consumed by the kernels on load.
static int md_step_structured(mm_md *md, bool full, bool own_pre, bool merge_next) {
    mm_handle *h = md->h;
    const bool thermo = md->desc.has_thermo != 0, baro = md->desc.has_baro != 0, lang = md->desc.has_langevin != 0;
    const unsigned post_thermo = (thermo && md->desc.thermo_kind == 0) ? OP_THERMO : 0u;  // Berendsen acts in "pre" only
    const int nb = h->sg.nblocks;
    const unsigned next_op = !merge_next ? 0u : baro ? OP_NEXT_BARO_A : thermo ? OP_NEXT_THERMO : lang ? OP_LANG_B : 0u;
    if (!full && sg_tail_ok(h) && !lang) {
        // Lean step on the fast path: every marching launch ends with its own reduction, slab exchange and scalar algebra
        // (tail of k_march2), and takes the periodic images on load - no launch between two marching kernels.
        SgTail t;
        t.state = md->d_state;
        if (baro) {
            if (own_pre) scalar_launch(md, OP_BARO_A, 0, 0, 0);
            t.ops = OP_POS_WRITTEN | OP_TAKE_FORCE | OP_BARO_B | (thermo ? OP_THERMO : 0u);
            sg_force(h, true, 2, false, &t);  // npt.py:683-707: rotate, evaluate, write x and g for the step below
            t.ops = OP_RESET_MVEL | OP_POS_WRITTEN | OP_TAKE_FORCE | OP_TAKE_KIN | post_thermo | OP_BARO_A;
            sg_step(h, false, 2, false, &t);
            t.ops = OP_TAKE_FORCE | OP_BARO_B | OP_ECONS | OP_ADVANCE | OP_PROPS | next_op;
            sg_force(h, false, 1, true, &t);  // npt.py:683-707 again: energy + virial of the rotated geometry
        } else {
            if (thermo && own_pre) scalar_launch(md, OP_THERMO, 0, 0, 0);
            t.ops = OP_RESET_MVEL | OP_POS_WRITTEN | OP_TAKE_FORCE | OP_TAKE_KIN | post_thermo | OP_ECONS | OP_ADVANCE |
                    OP_PROPS | next_op;
            sg_step(h, true, thermo ? 1 : 0, true, &t);
        }
        h->sg.fused_mask = 0;
        h->sg.tail_done = 0;
        return MM_OK;
    }
    if (full) sg_export(md, true, false, md->d_posold);  // posold of verlet.py:158-161 (a full step always has own_pre)
    // Ghost fill beside the scalar kernel (lean steps, one slab).  The ghost positions are shifted by the domain vectors of the
    // frame the positions were stored in; where the scalar kernel running alongside is about to change StepConsts::rv, the
    // caller names a source that is already final (see the two call sites).
    const bool beside = md->overlap_halo && !full && !lang && h->slab_count <= 1 && md->side != nullptr;
    auto halo_beside = [&](bool pos, bool vel, bool grad, const double *rv_src, unsigned ops, int nbd_) {
        MM_CUDA(cudaEventRecord(md->ev_fork, h->stream));
        MM_CUDA(cudaStreamWaitEvent(md->side, md->ev_fork, 0));
        int rc = sg_halo(h, pos, vel, grad, rv_src, md->side);
        if (rc != MM_OK) return rc;
        MM_CUDA(cudaEventRecord(md->ev_join, md->side));
        rc = scalar_launch(md, ops, nb, nb, nbd_);
        MM_CUDA(cudaStreamWaitEvent(h->stream, md->ev_join, 0));
        return rc;
    };
    if (baro) {
        if (own_pre) scalar_launch(md, OP_BARO_A, 0, 0, 0);
        // npt.py:683-707: rotate the positions (all pending rotations at once) and evaluate; the rotated positions are
        // written back so that the fused step below needs no rotation; the gradient is what its first kick uses
        sg_force(h, true, 2);
        const unsigned ops_b = OP_POS_WRITTEN | OP_TAKE_FORCE | OP_BARO_B | (thermo ? OP_THERMO : 0u);
        if (beside) {
            // the rotated positions are in the true frame: their domain vectors are the handle's device copy, final since the
            // barostat's first half (OP_BARO_A / OP_NEXT_BARO_A) - exactly what OP_POS_WRITTEN is about to put into StepConsts
            const int rc = halo_beside(true, false, true, h->d_rvecs, ops_b, 0);
            if (rc != MM_OK) return rc;
        } else {
            scalar_launch(md, ops_b, nb, nb, 0);
            sg_halo(h, true, false, true);  // after OP_POS_WRITTEN: the halo shift uses the new stored frame
        }
    } else if (thermo) {
        if (own_pre) scalar_launch(md, OP_THERMO, 0, 0, 0);
    } else if (lang && own_pre) {
        lang_launch(md, 1, 0);
        scalar_launch(md, OP_LANG_A, 0, 0, 0);
    }
    // without a barostat the gradient written here feeds the next step's first kick
    sg_step(h, !baro, baro ? 2 : (thermo ? 1 : 0), !full);
    unsigned ops = OP_RESET_MVEL | OP_POS_WRITTEN | OP_TAKE_FORCE | OP_TAKE_KIN | post_thermo;
    int nbd = 0;
    if (!baro && beside) {
        // rv_stored is unchanged without a barostat: the scalar kernel rewrites StepConsts::rv with the same values
        const int rc = halo_beside(true, true, true, nullptr, ops | OP_ECONS | OP_ADVANCE | OP_PROPS | next_op, 0);
        if (rc != MM_OK) return rc;
    } else if (!baro) {
        sg_halo(h, true, true, true);  // rv_stored is unchanged without a barostat: safe before the scalar kernel
        if (full) {
            sg_export(md, true, false, h->d_pos);
            nbd = grid_for(h, 3 * h->nnodes, kNodeThreads);
            k_delta<<<nbd, kNodeThreads, 0, h->stream>>>(h->d_pos, md->d_posold, 3 * h->nnodes, md->d_pdelta);
            h->launches++;
            ops |= OP_TAKE_DELTA;
        }
        if (lang) {  // every copy of a node (ghosts, halo planes) gets the same kick: no exchange afterwards
            lang_launch(md, merge_next ? 2 : 1, 1);
            ops |= OP_LANG_A;
        }
        scalar_launch(md, ops | OP_ECONS | OP_ADVANCE | OP_PROPS | next_op, nb, nb, nbd);
    } else {
        if (beside) {
            // the drifted positions stay in the frame of the launch before (true frame, cell unchanged by the Verlet step):
            // OP_POS_WRITTEN rewrites StepConsts::rv with the values it holds; OP_BARO_A then moves on the handle's copy only
            const int rc = halo_beside(true, true, false, nullptr, ops | OP_BARO_A, 0);
            if (rc != MM_OK) return rc;
        } else {
            scalar_launch(md, ops | OP_BARO_A, nb, nb, 0);
            sg_halo(h, true, true, false);  // after OP_POS_WRITTEN: the halo shift must use the new stored frame
        }
        // npt.py:683-707 again; this rotation stays pending until the next step.  The gradient of this call is only
        // read by compute_properties (rmsd_gpos) and by trajectory output: lean steps evaluate energy + virial alone
        sg_force(h, full, 1, !full);
        ops = OP_TAKE_FORCE | OP_BARO_B | OP_ECONS | OP_ADVANCE | OP_PROPS | next_op;
        if (full) {
            sg_export(md, true, false, h->d_pos);
            nbd = grid_for(h, 3 * h->nnodes, kNodeThreads);
            k_delta<<<nbd, kNodeThreads, 0, h->stream>>>(h->d_pos, md->d_posold, 3 * h->nnodes, md->d_pdelta);
            h->launches++;
            ops |= OP_TAKE_DELTA;
        }
        scalar_launch(md, ops, nb, nb, nbd);
    }
    return MM_OK;
}

}  // namespace mm

extern "C" {

int mm_md_destroy(mm_md *md) {
    if (!md) return MM_OK;
    cudaSetDevice(md->h->device);
    cudaFree(md->d_state);
    cudaFree(md->d_vel);
    cudaFree(md->d_masses);
    cudaFree(md->d_posold);
    cudaFree(md->d_pkin);
    cudaFree(md->d_pdelta);
    cudaFree(md->d_plang);
    if (md->h_state) cudaFreeHost(md->h_state);
    for (int i = 0; i < 8; i++)
        if (md->gexec[i]) cudaGraphExecDestroy(md->gexec[i]);
    if (md->side) cudaStreamDestroy(md->side);
    if (md->ev_fork) cudaEventDestroy(md->ev_fork);
    if (md->ev_join) cudaEventDestroy(md->ev_join);
    delete md;
    return MM_OK;
}

int mm_md_create(mm_handle *h, const mm_md_desc *desc, mm_md **out) {
    if (!h || !desc || !out) {
        set_error("mm_md_create: null argument");
        return MM_ERR_INVALID;
    }
    *out = nullptr;
    if (desc->has_thermo && desc->thermo_kind != 0 && (desc->has_baro || desc->thermo_kind < 0 || desc->thermo_kind > 2)) {
        set_error("mm_md_create: the device Berendsen / CSVR thermostats run without a barostat");
        return MM_ERR_INVALID;
    }
    if (desc->has_thermo && desc->thermo_kind == 0 && (desc->chain_length < 1 || desc->chain_length > MM_MAX_CHAIN)) {
        set_error("mm_md_create: unsupported Nose-Hoover chain length");
        return MM_ERR_INVALID;
    }
    if (desc->has_baro && !desc->anisotropic && desc->vol_constraint) {
        set_error("Isotropic barostat called with a volume constraint.");  // sampling/utils.py:499-500
        return MM_ERR_INVALID;
    }
    if (!(desc->timestep > 0.0)) {
        set_error("mm_md_create: timestep must be positive");
        return MM_ERR_INVALID;
    }
    if (desc->has_langevin && (desc->has_thermo || desc->has_baro || h->slab_count > 1 || !(desc->langevin_timecon > 0.0))) {
        set_error("mm_md_create: the device Langevin thermostat runs alone (no chain, no barostat) on one GPU");
        return MM_ERR_INVALID;
    }
    mm_md *md = new (std::nothrow) mm_md();
    if (!md) return MM_ERR_INVALID;
    md->h = h;
    md->desc = *desc;
    const int64_t nn = h->nnodes;
#define MM_TRY(call)                                \
    do {                                            \
        cudaError_t err__ = (call);                 \
        if (err__ != cudaSuccess) {                 \
            int rc__ = mm::cuda_fail(err__, #call); \
            mm_md_destroy(md);                      \
            return rc__;                            \
        }                                           \
    } while (0)
    MM_TRY(cudaSetDevice(h->device));
    MM_TRY(cudaMalloc(&md->d_state, sizeof(MDState)));
    MM_TRY(cudaMalloc(&md->d_vel, sizeof(double) * 3 * nn));
    MM_TRY(cudaMalloc(&md->d_masses, sizeof(double) * nn));
    MM_TRY(cudaMalloc(&md->d_posold, sizeof(double) * 3 * nn));
    MM_TRY(cudaMalloc(&md->d_pkin, sizeof(double) * kMaxRedBlocks * kRedSlots));
    MM_TRY(cudaMalloc(&md->d_pdelta, sizeof(double) * kMaxRedBlocks * kRedSlots));
    MM_TRY(cudaMalloc(&md->d_plang, sizeof(double) * kMaxRedBlocks * kRedSlots));
    MM_TRY(cudaHostAlloc(&md->h_state, sizeof(MDState), cudaHostAllocDefault));
    MM_TRY(cudaStreamCreateWithFlags(&md->side, cudaStreamNonBlocking));
    MM_TRY(cudaEventCreateWithFlags(&md->ev_fork, cudaEventDisableTiming));
    MM_TRY(cudaEventCreateWithFlags(&md->ev_join, cudaEventDisableTiming));
    if (const char *e = getenv("MICMEC_B200_HALO_OVERLAP")) md->overlap_halo = atoi(e) != 0;  // A/B switch (profiles/r02)
#undef MM_TRY
    *out = md;
    return MM_OK;
}

int mm_md_init(mm_md *md, const double *pos, const double *vel, const double *masses, int where, const double *rvecs9,
               const double *chain_pos, const double *chain_vel, const double *vel_press9) {
    if (!md || !pos || !vel || !masses || !rvecs9) {
        set_error("mm_md_init: null argument");
        return MM_ERR_INVALID;
    }
    mm_handle *h = md->h;
    const int64_t nn = h->nnodes;
    MM_CUDA(cudaSetDevice(h->device));
    const cudaMemcpyKind kind = where == MM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int rc = mm_set_pos(h, pos, where);
    if (rc != MM_OK) return rc;
    rc = mm_set_rvecs(h, rvecs9);
    if (rc != MM_OK) return rc;
    MM_CUDA(cudaMemcpyAsync(md->d_vel, vel, sizeof(double) * 3 * nn, kind, h->stream));
    MM_CUDA(cudaMemcpyAsync(md->d_masses, masses, sizeof(double) * nn, kind, h->stream));
    MM_CUDA(cudaMemsetAsync(md->d_posold, 0, sizeof(double) * 3 * nn, h->stream));
    MDState &s = *md->h_state;
    memset(&s, 0, sizeof(s));
    const mm_md_desc &d = md->desc;
    s.boltzmann = h->boltzmann;
    s.timestep = d.timestep;
    s.ndof = d.ndof;
    s.time = d.time0;  // verlet.py:96 / iterative.py: a restarted run continues its clock and its step counter
    s.counter = d.counter0;
    for (int i = 0; i < 9; i++) {
        s.rvecs[i] = rvecs9[i];
        s.Mvel[i] = s.Rpos[i] = s.Rpend[i] = (i % 4 == 0) ? 1.0 : 0.0;
        s.rv_stored[i] = rvecs9[i];
    }
    md->structured = h->sg.active != 0;
    s.has_thermo = d.has_thermo;
    s.chain_len = d.has_thermo ? d.chain_length : 0;
    s.ch_temp = d.thermo_temp;
    s.ch_timecon = d.thermo_timecon;
    for (int k = 0; k < s.chain_len; k++) {
        s.ch_pos[k] = chain_pos ? chain_pos[k] : 0.0;
        s.ch_vel[k] = chain_vel ? chain_vel[k] : 0.0;
    }
    s.has_langevin = d.has_langevin;
    s.thermo_kind = d.has_thermo ? d.thermo_kind : 0;
    if (s.thermo_kind == 2) s.lg_seed = d.langevin_seed;  // the stochastic velocity rescaling draws from the same generator
    s.lg_temp = d.langevin_temp;
    s.lg_timecon = d.langevin_timecon;
    s.lg_seed = d.langevin_seed;
    s.has_baro = d.has_baro;
    s.aniso = d.anisotropic;
    s.volc = d.vol_constraint;
    s.baro_ndof = (d.anisotropic ? 6 : 1) - (d.vol_constraint ? 1 : 0);  // sampling/utils.py:478-501
    s.b_temp = d.baro_temp;
    s.b_press = d.baro_press;
    s.b_timecon = d.baro_timecon;
    if (d.has_baro && vel_press9)
        for (int i = 0; i < (d.anisotropic ? 9 : 1); i++) s.vp[i] = vel_press9[i];
    MM_CUDA(cudaMemcpyAsync(md->d_state, md->h_state, sizeof(MDState), cudaMemcpyHostToDevice, h->stream));
    // verlet.py:119-137: first force evaluation (no vtens unless a barostat re-evaluates, npt.py:612-614)
    const int gn = grid_for(h, nn, kNodeThreads);
    if (md->structured) {
        rc = sg_write_consts(h, rvecs9, d.timestep);
        if (rc != MM_OK) return rc;
        h->sg.cx = h->sg.cv = h->sg.cg = 0;
        sg_pos_from_aos(h, h->d_pos);
        sg_vel_from_aos(h, md->d_vel);
        sg_mass_from_aos(h, md->d_masses);
        sg_force(h, true, 0);
        sg_halo(h, false, false, true);
        const int nb = h->sg.nblocks;
        scalar_launch(md, OP_TAKE_FORCE | (d.has_baro ? 0u : OP_ZERO_VIR) | OP_SETUP, nb, nb, 0);
        k_moments<<<gn, kNodeThreads, 0, h->stream>>>(md->d_vel, md->d_masses, nullptr, nn, md->d_pkin);
        h->launches++;
        // the moments come from the generic partial buffer this once: temporarily read them from there
        md->structured = false;
        scalar_launch(md, OP_TAKE_KIN | OP_PROPS, 0, gn, 0);
        md->structured = true;
        scalar_launch(md, 0u, 0, 0, 0);  // publish the StepConsts (dt, frames) for the structured kernels
        MM_CUDA(cudaGetLastError());
        MM_CUDA(cudaStreamSynchronize(h->stream));
        md->initialised = true;
        return MM_OK;
    }
    rc = ensure_generic(h);
    if (rc != MM_OK) return rc;
    const int nbc = cells_launch(h);
    k_gather_g2<<<gn, kNodeThreads, 0, h->stream>>>(h->d_node_cells, h->d_gcell, nn, h->ncells, h->d_gpos, md->d_pkin);
    h->launches++;
    scalar_launch(md, OP_TAKE_FORCE | (d.has_baro ? 0u : OP_ZERO_VIR) | OP_SETUP, nbc, gn, 0);
    // (with a barostat the reference evaluates the same geometry a second time, now with vtens, npt.py:612-614;
    //  the virial of the first evaluation is kept here instead)
    k_moments<<<gn, kNodeThreads, 0, h->stream>>>(md->d_vel, md->d_masses, h->d_gpos, nn, md->d_pkin);
    h->launches++;
    scalar_launch(md, OP_TAKE_KIN | OP_PROPS, 0, gn, 0);
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaStreamSynchronize(h->stream));
    md->initialised = true;
    return MM_OK;
}

int mm_md_set_state(mm_md *md, const double *pos, const double *vel, int where) {
    if (!md || !md->initialised) {
        set_error("mm_md_set_state: integrator not initialised");
        return MM_ERR_STATE;
    }
    mm_handle *h = md->h;
    MM_CUDA(cudaSetDevice(h->device));
    const cudaMemcpyKind kind = where == MM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (pos) MM_CUDA(cudaMemcpyAsync(h->d_pos, pos, sizeof(double) * 3 * h->nnodes, kind, h->stream));
    if (vel) {
        // the uploaded velocities are the true ones: drop any pending transform
        scalar_launch(md, OP_RESET_MVEL, 0, 0, 0);
        MM_CUDA(cudaMemcpyAsync(md->d_vel, vel, sizeof(double) * 3 * h->nnodes, kind, h->stream));
    }
    if (md->structured) {
        if (pos) {
            scalar_launch(md, OP_POS_WRITTEN, 0, 0, 0);  // uploaded positions are in the true frame
            sg_pos_from_aos(h, h->d_pos);
        }
        if (vel) sg_vel_from_aos(h, md->d_vel);
    }
    return MM_OK;
}

int mm_md_run(mm_md *md, int64_t nsteps) {
    if (!md || !md->initialised) {
        set_error("mm_md_run: integrator not initialised");
        return MM_ERR_STATE;
    }
    mm_handle *h = md->h;
    MM_CUDA(cudaSetDevice(h->device));
    auto one_step = [&](bool full, bool own_pre, bool merge_next) {
        return md->structured ? md_step_structured(md, full, own_pre, merge_next) : md_step(md, full, own_pre, merge_next);
    };
    // Step roles inside one call: L0 (own "pre" launch), middle lean steps (their "pre" was merged into the previous
    // step's last scalar launch and they merge the next one's), the last lean step (does not merge: the final step takes
    // its posold snapshot first), and the final FULL step (own "pre"; also produces rmsd_delta / gpos).
    const int64_t lean = nsteps > 0 ? nsteps - 1 : 0;
    // the first steps after initialisation run directly (lazy allocations happen there), as do profiled runs
    const bool graphs = md->use_graphs && !h->profile && md->steps_done >= 2;
    int64_t i = 0;
    if (lean >= 1) {
        one_step(false, true, lean >= 2);
        i = 1;
    }
    while (i + 1 < lean) {  // middle steps i .. lean-2
        const int64_t left = lean - 1 - i;
        if (!graphs || left < 2) {
            one_step(false, false, true);
            i++;
            continue;
        }
        const int key = md->structured ? (h->sg.cx | (h->sg.cv << 1) | (h->sg.cg << 2)) : 0;
        if (!md->gexec[key]) {  // capture two middle steps from the current buffer parity
            cudaGraph_t graph = nullptr;
            const int64_t before = h->launches;
            MM_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            int rc = one_step(false, false, true);
            if (rc == MM_OK) rc = one_step(false, false, true);
            const cudaError_t err = cudaStreamEndCapture(h->stream, &graph);
            md->glaunches[key] = h->launches - before;
            h->launches = before;
            if (rc != MM_OK || err != cudaSuccess || !graph) {
                cudaGetLastError();
                if (graph) cudaGraphDestroy(graph);
                md->use_graphs = 0;  // fall back to direct launches for good
                set_error("CUDA graph capture of the MD step failed");
                return rc != MM_OK ? rc : MM_ERR_CUDA;
            }
            const cudaError_t ierr = cudaGraphInstantiate(&md->gexec[key], graph, 0);
            cudaGraphDestroy(graph);
            if (ierr != cudaSuccess) return cuda_fail(ierr, "cudaGraphInstantiate");
            // capturing does not execute; the buffer-parity flags were flipped twice by the captured calls = unchanged
        }
        MM_CUDA(cudaGraphLaunch(md->gexec[key], h->stream));
        h->launches += md->glaunches[key];
        i += 2;
    }
    if (lean >= 2) one_step(false, false, false);  // last lean step
    if (nsteps > 0) one_step(true, true, false);
    md->steps_done += nsteps;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

int mm_md_get_state(mm_md *md, double *pos, double *vel, double *gpos, int where, double *rvecs9, double *chain_pos,
                    double *chain_vel, double *vel_press9) {
    if (!md || !md->initialised) {
        set_error("mm_md_get_state: integrator not initialised");
        return MM_ERR_STATE;
    }
    mm_handle *h = md->h;
    const int64_t nn = h->nnodes;
    MM_CUDA(cudaSetDevice(h->device));
    const cudaMemcpyKind kind = where == MM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (md->structured) {
        // export copies of the SoA state in reference order; the device state (and its pending transforms) is untouched
        sg_export(md, pos != nullptr, vel != nullptr, h->d_pos);
        if (gpos) sg_to_aos(h, 2, h->d_gpos);
        if (vel) MM_CUDA(cudaMemcpyAsync(vel, md->d_vel, sizeof(double) * 3 * nn, kind, h->stream));
    } else if (vel) {
        const int gn = grid_for(h, nn, kNodeThreads);
        k_flush_vel<<<gn, kNodeThreads, 0, h->stream>>>(md->d_state, md->d_vel, nn);
        h->launches++;
        scalar_launch(md, OP_RESET_MVEL, 0, 0, 0);
        MM_CUDA(cudaMemcpyAsync(vel, md->d_vel, sizeof(double) * 3 * nn, kind, h->stream));
    }
    if (pos) MM_CUDA(cudaMemcpyAsync(pos, h->d_pos, sizeof(double) * 3 * nn, kind, h->stream));
    if (gpos) MM_CUDA(cudaMemcpyAsync(gpos, h->d_gpos, sizeof(double) * 3 * nn, kind, h->stream));
    MM_CUDA(cudaMemcpyAsync(md->h_state, md->d_state, sizeof(MDState), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    const MDState &s = *md->h_state;
    memcpy(h->rvecs, s.rvecs, sizeof(double) * 9);
    if (rvecs9) memcpy(rvecs9, s.rvecs, sizeof(double) * 9);
    if (chain_pos) memcpy(chain_pos, s.ch_pos, sizeof(double) * s.chain_len);
    if (chain_vel) memcpy(chain_vel, s.ch_vel, sizeof(double) * s.chain_len);
    if (vel_press9) memcpy(vel_press9, s.vp, sizeof(double) * 9);
    return MM_OK;
}

int mm_md_scalars(mm_md *md, double *out) {
    if (!md || !md->initialised || !out) {
        set_error("mm_md_scalars: integrator not initialised");
        return MM_ERR_STATE;
    }
    mm_handle *h = md->h;
    MM_CUDA(cudaSetDevice(h->device));
    MM_CUDA(cudaMemcpyAsync(md->h_state, md->d_state, sizeof(MDState), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    const MDState &s = *md->h_state;
    for (int i = 0; i < MM_S_COUNT; i++) out[i] = 0.0;
    out[MM_S_EPOT] = s.epot;
    out[MM_S_EKIN] = s.ekin;
    out[MM_S_TEMP] = s.temp;
    out[MM_S_ETOT] = s.etot;
    out[MM_S_ECONS] = s.econs;
    out[MM_S_CONS_ERR] = s.cons_err;
    out[MM_S_PRESS] = s.press;
    out[MM_S_RMSD_GPOS] = s.rmsd_gpos;
    out[MM_S_RMSD_DELTA] = s.rmsd_delta;
    out[MM_S_TIME] = s.time;
    out[MM_S_COUNTER] = (double)s.counter;
    out[MM_S_VOLUME] = s.volume;
    out[MM_S_NDOF] = s.ndof;
    out[MM_S_ECONS_CORR] = s.econs_corr;
    double v[9];
    v[0] = s.vir[0]; v[4] = s.vir[1]; v[8] = s.vir[2];
    v[5] = v[7] = s.vir[3]; v[2] = v[6] = s.vir[4]; v[1] = v[3] = s.vir[5];
    for (int i = 0; i < 9; i++) {
        out[MM_S_VTENS + i] = v[i];
        out[MM_S_PTENS + i] = s.ptens[i];
    }
    out[MM_S_NFORCE] = (double)s.nforce;
    out[MM_S_CE_N] = (double)s.ce_n;
    out[MM_S_CE_N + 1] = s.ce_ekin_m;
    out[MM_S_CE_N + 2] = s.ce_ekin_s;
    out[MM_S_CE_N + 3] = s.ce_econs_m;
    out[MM_S_CE_N + 4] = s.ce_econs_s;
    // the reference raises from inside compute() (mmff.py:135-147); here the flag surfaces with the scalars
    if (std::isnan(s.epot) || std::isnan(s.ekin)) {
        set_error("The energy is not-a-number (``nan``).");
        return MM_ERR_NAN;
    }
    return MM_OK;
}

}  // extern "C"
