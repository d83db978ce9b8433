// mm_march2.cuh - the marching kernel of the one-type / one-state / uniform-mass fast path (the benchmark grids).
//
// Same arithmetic as k_march (mm_march.cuh; reference: micmec/pes/mmff.py:288-403 with nanocell_original.py:46-132, and
// micmec/sampling/verlet.py:144-154), restructured after the round-1 profiles (profiles/r02/a_ncu_march_npt_head.csv):
// the round-1 kernel ran the FP64 pipe at 46-56 % with two warps per scheduler that march in lockstep around one block
// barrier per plane, recomputed 1.47x cells (a 32 x 8 tile owns 30 x 6 columns), staged m and 1/m as full arrays, and
// needed a ghost-fill launch, an all-reduce launch and a scalar launch between any two marching launches.
//
//   * RPT rows per thread.  A block is still TY warps, but every thread owns RPT consecutive node rows: the tile is
//     32 x (RPT TY) nodes and owns 30 x (RPT TY - 2).  RPT = 2, TY = 8: 30 x 14 of 32 x 16, recomputation 1.22x instead
//     of 1.47x.  The y butterflies between a thread's own rows stay in registers: only the first row of a warp goes
//     through shared memory forward and only its last row backward (3 STS + 3 LDS per plane and THREAD each way, i.e.
//     half of the round-1 traffic per node), one block barrier serves twice as many nodes, and the two cells of a
//     thread are independent instruction streams for the FP64 pipe.
//   * Uniform mass: a grid of one cell type has one node mass (micmec/utils.py:217).  It is a kernel argument; the
//     staged fields of a STEP launch drop from 11 to 9 and 16 B/node of DRAM traffic disappear.  Grids with mixed
//     masses run on k_march, which keeps the mass arrays.
//   * MODE 2 (virial only): the barostat's second force call of a step that nobody observes (npt.py:699-707 followed by
//     the next step's call, which overwrites the gradient) only needs the energy and the virial: no backward butterfly,
//     no gather, nothing written.
//   * Node rows r >= 1 of a thread complete in the iteration that evaluates their upper cell layer; only row 0, whose
//     lower-row part comes from the warp below through shared memory, completes one iteration later (one barrier per
//     plane, shared-memory exchange buffers double-buffered by plane parity, as in k_march's PIPE variant).
//   * Periodic images along x and y are taken ON LOAD: an edge tile fetches the image column (a 2 x rows tensor box), the
//     image row (one 272-byte bulk copy) and the image corner of every staged field next to its main box, and the
//     threads of the apron read their node from there and add the domain-vector shift.  The x / y ghost nodes of the
//     padded arrays are therefore never read by this kernel, and nothing has to refresh them between two launches.
//   * Tail: the last block to finish sums the block partials in a fixed order, exchanges them with the other z-slabs
//     (one-shot mailbox all-gather over NVLink peer memory; receiving every rank's row also proves that the
//     neighbours' boundary planes, stored into this rank's halo planes by their marching blocks, have landed) and runs
//     the Nose-Hoover / MTK algebra of mm_scalar.cuh.  A lean NPT step is exactly three launches.
#pragma once
#include "mm_march.cuh"
#include "mm_peer.cuh"
#include "mm_scalar.cuh"

namespace mm {

constexpr int M2_FORCE = 0, M2_STEP = 1, M2_VIRIAL = 2;

template <int RPT, int TY>
struct March2Cfg {
    static constexpr int NR = RPT * TY;  // node rows of the tile
    static constexpr int OX = TX - 2, OY = NR - 2;
    // one field of one staged plane: main box [NR][34] | image columns [2 sides][NR][2] | image rows [2 sides][34] |
    // image corners [2][2][2]; rounded to 128 bytes (destination alignment of the tensor boxes)
    static constexpr int XW = NR * kBoxW, YW = XW + 4 * NR, CW = YW + 2 * kBoxW;
    static constexpr int FS = (CW + 8 + 15) & ~15;
    static constexpr int kStaticBytes = 2 * 2 * 3 * TY * TX * 8 + 6144;  // sf, sb, reduction scratch, tail state
    __host__ __device__ static constexpr int stages(int nf) { return (4 * nf * FS * 8 + kStaticBytes <= 227 * 1024) ? 4 : 3; }
    __host__ __device__ static constexpr size_t dyn_bytes(int nf) { return (size_t)stages(nf) * nf * FS * 8; }
};

// plain bulk copy global -> shared memory (16-byte aligned, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}

// Image copies of an edge tile for one staged plane (one thread; see k_march2).  Deliberately out of line and not a
// template: inlined into the plane loop (once per peeled body) this code tripled the loop's footprint in the instruction
// cache - 31 KB against a 32 KB L1.5 - and the kernel ran 60 % slower (profiles/r02: stall_no_instruction 0.16 -> 0.99).
// hmask: bit 0 image column k = -1, bit 1 column k = nx, bit 2 image row l = -1, bit 3 row l = ny.
static __device__ __noinline__ void issue_edge_copies(const TmaMaps *maps, const MarchArgs *a, const unsigned field_smem, const int f,
                                                      const unsigned xw_bytes, const unsigned yw_bytes, const unsigned cw_bytes,
                                                      const unsigned xw_side_bytes, const int q, const unsigned plane, const int bx0,
                                                      const int by0, const unsigned bar, const int hmask) {
    const int nx = a->nx, ny = a->ny, nxp = a->nxp;
    // source columns / rows of the images: node nx-1 (padded column nx+1, inside an even-aligned pair) for k = -1, node 0
    // (padded column 2) for k = nx; node row ny-1 (padded row ny) for l = -1, node row 0 (padded row 1) for l = ny
    const double *src = (f < 3 ? a->x[f] : f < 6 ? a->v[f - 3] : a->g[f - 6]) + (size_t)q * plane;
    for (int sx = 0; sx < 2; sx++) {
        if (!(hmask & (1 << sx))) continue;
        const int colx = sx == 0 ? ((nx + 1) & ~1) : kGhostX;
        tma_load_3d(field_smem + xw_bytes + sx * xw_side_bytes, &maps->xw[f], colx, by0, q, bar);
    }
    for (int sy = 0; sy < 2; sy++) {
        if (!(hmask & (4 << sy))) continue;
        const int rowy = sy == 0 ? ny : 1;
        bulk_load_1d(field_smem + yw_bytes + sy * kBoxW * 8u, src + (size_t)rowy * nxp + bx0, kBoxW * 8u, bar);
        for (int sx = 0; sx < 2; sx++) {
            if (!(hmask & (1 << sx))) continue;
            const int colx = sx == 0 ? ((nx + 1) & ~1) : kGhostX;
            bulk_load_1d(field_smem + cw_bytes + (sx * 2 + sy) * 16u, src + (size_t)rowy * nxp + colx, 16u, bar);
        }
    }
}

// ---- tail of a launch (see the header): executed by the last block to finish --------------------------------------
static __device__ __noinline__ void march_tail(const TailArgs &t, const double *partials, const int nblocks) {
    __shared__ double grp[16][17];  // at most 256 threads
    __shared__ double sums[16];
    __shared__ MDState sm_state;
    const int nthreads = blockDim.x * blockDim.y, tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int slot = tid & 15, group = tid >> 4, ngroups = nthreads >> 4;
    // thread (group, slot) adds slot `slot` of the blocks group, group + ngroups, ...; the groups are then added in order
    double acc = 0.0;
    int b = group;
    for (; b + 7 * ngroups < nblocks; b += 8 * ngroups) {  // eight independent loads in flight (L2 latency, not bandwidth)
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcg(partials + (size_t)(b + u * ngroups) * 16 + slot);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
    }
    for (; b < nblocks; b += ngroups) acc += __ldcg(partials + (size_t)b * 16 + slot);
    grp[group][slot] = acc;
    __syncthreads();
    if (tid < 16) {
        double s = 0.0;
        for (int q = 0; q < ngroups; q++) s += grp[q][tid];
        sums[tid] = (tid < 14) ? s : 0.0;
    }
    __syncthreads();
    if (t.nranks > 1) {  // one-shot all-gather through the mailboxes, rank-ordered sum (as k_peer_allreduce, mm_comm.cu)
        PeerCtl *ctl = reinterpret_cast<PeerCtl *>(t.ctl);
        const unsigned long long E = ctl->red_epoch + 1;
        const size_t par = (size_t)(E & 1ull) * t.nranks * 16;
        if (tid < 16 * t.nranks) {  // my row on every rank (my own included)
            double *mail = reinterpret_cast<double *>(t.bases[tid / 16] + kPeerFlagBytes);
            mail[par + (size_t)t.rank * 16 + (tid % 16)] = sums[tid % 16];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < t.nranks) {
            st_release_sys(reinterpret_cast<unsigned long long *>(t.bases[tid]) + 2 + t.rank, E);
            const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(t.bases[t.rank]) + 2 + tid;
            while (ld_acquire_sys(flag) < E) {
            }
        }
        __syncthreads();
        if (tid < 16) {
            const double *mail = reinterpret_cast<const double *>(t.bases[t.rank] + kPeerFlagBytes) + par;
            double s = 0.0;
            for (int r = 0; r < t.nranks; r++) s += __ldcg(mail + (size_t)r * 16 + tid);
            sums[tid] = s;
        }
        if (tid == 0) ctl->red_epoch = E;
        __syncthreads();
    }
    constexpr int kWords = sizeof(MDState) / sizeof(double);
    for (int i = tid; i < kWords; i += nthreads)
        reinterpret_cast<double *>(&sm_state)[i] = reinterpret_cast<const double *>(t.state)[i];
    __syncthreads();
    if (tid == 0) {
        double fr[7], kn[7], dl[1] = {0.0};
        for (int q = 0; q < 7; q++) {
            fr[q] = sums[q];
            kn[q] = sums[7 + q];
        }
        scalar_ops(sm_state, t.rvecs_dev, t.sc_out, t.ops, fr, kn, dl, 1, t.n3);
        *t.counter = 0u;
    }
    __syncthreads();
    for (int i = tid; i < kWords; i += nthreads)
        reinterpret_cast<double *>(t.state)[i] = reinterpret_cast<const double *>(&sm_state)[i];
}

template <int MODEL, int MODE, int ROT, int VM, bool LEAN, int RPT, int TY, int PIN, bool WRAP, int UNR>
__global__ void __launch_bounds__(TX *TY, 1)
k_march2(const __grid_constant__ SState Pc, const __grid_constant__ MarchArgs a, const __grid_constant__ TmaMaps maps,
         const int write_g) {
    using Cfg = March2Cfg<RPT, TY>;
    constexpr bool STEP = MODE == M2_STEP, VIRIAL = MODE == M2_VIRIAL;
    constexpr int NR = Cfg::NR, OX = Cfg::OX, OY = Cfg::OY, FS = Cfg::FS;
    constexpr int NF = STEP ? 9 : 3;  // staged fields per node: x (3) [, v (3), g (3)]
    constexpr int NST = Cfg::stages(NF);
    constexpr bool WANT_VIR = !LEAN;
    __shared__ double sf[2][3][TY][TX];  // forward exchange along y: position of the first row of the warp above
    __shared__ double sb[2][3][TY][TX];  // backward exchange along y: x-combined gradient part of the last row of the warp below
    __shared__ double s_shift[9][3];     // periodic shift of an image node: (sx + 1) * 3 + (sy + 1) -> sx a + sy b
    __shared__ __align__(8) unsigned long long s_full[NST];
    __shared__ int s_last;
    extern __shared__ __align__(128) double s_stage[];  // [NST][NF][FS]

    const int lane = threadIdx.x, w = threadIdx.y;
    const int nx = a.nx, ny = a.ny, nxp = a.nxp;
    // thread = RPT node columns (k, l0 + r) = cell columns with those origin vertices.  k = -1 / k = nx and l = -1 / l = ny
    // are the periodic images of the last / first column and row: edge tiles stage them separately (see stage_issue) and
    // `off` points the thread at the right copy.  Rows or columns beyond are never owned; what they read is finite.
    // work item of this block: tile (bx, by) and the plane range [c0, c1) it owns.  Items come from a table built on the host
    // (sg_plan_items): tiles do not all get the same number of chunks - whole columns first, the remainder of the tiles cut
    // into short pieces that fill the last round of the SMs
    const int4 item = a.items[blockIdx.x];
    const int bx = item.x, by = item.y;
    const int k = bx * OX + lane - 1, l0 = by * OY + w * RPT - 1;
    // WRAP = false: the images are read from the ghost nodes of the padded arrays instead (kept current by sg_halo)
    const bool hx0 = WRAP && bx == 0, hx1 = WRAP && bx == a.ntx - 1;
    const bool hy0 = WRAP && by == 0, hy1 = WRAP && by == a.nty - 1;
    const bool edge = hx0 || hx1 || hy0 || hy1;
    const int hmask = (hx0 ? 1 : 0) | (hx1 ? 2 : 0) | (hy0 ? 4 : 0) | (hy1 ? 8 : 0);
    const int ex0 = (nx + 1) & 1;  // element of node nx-1 inside its (even-aligned) 2-column box; node 0 is element 0 of its box
    bool own[RPT];
    int off[RPT], code[RPT];
    {
        const int wx = (k == -1) ? 0 : (k == nx ? 1 : -1);
#pragma unroll
        for (int r = 0; r < RPT; r++) {
            const int row = w * RPT + r, l = l0 + r;
            own[r] = lane >= 1 && lane <= OX && row >= 1 && row <= OY && k < nx && l < ny;
            const int wy = (l == -1) ? 0 : (l == ny ? 1 : -1);
            const int ex = wx == 0 ? ex0 : 0;
            if (!WRAP || (wx < 0 && wy < 0)) off[r] = row * kBoxW + lane + 1;
            else if (wy < 0) off[r] = Cfg::XW + wx * (2 * NR) + row * 2 + ex;
            else if (wx < 0) off[r] = Cfg::YW + wy * kBoxW + lane + 1;
            else off[r] = Cfg::CW + (wx * 2 + wy) * 2 + ex;
            // image of the LAST column / row (k = -1, l = -1) sits one domain vector BELOW its source, and vice versa
            code[r] = ((wx == 0 ? -1 : wx == 1 ? 1 : 0) + 1) * 3 + (wy == 0 ? -1 : wy == 1 ? 1 : 0) + 1;
        }
    }
    // 1.0 for owned nodes: the accumulators take e, vir, ... of apron nodes with weight zero (one DFMA instead of two
    // FSEL + DADD per term; apron data are real nodes or the zero fill of the tensor maps, hence finite)
    double ownf[RPT];
#pragma unroll
    for (int r = 0; r < RPT; r++) ownf[r] = own[r] ? 1.0 : 0.0;
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
    const double dt = sc.dt, mass = a.mass, hminv = 0.5 * sc.dt / a.mass;
    double cw[3];  // third domain vector in the frame of the positions this launch WRITES (see k_march)
#pragma unroll
    for (int j = 0; j < 3; j++)
        cw[j] = ROT ? fma(sc.rv[8], R[6 + j], fma(sc.rv[7], R[3 + j], sc.rv[6] * R[j])) : sc.rv[6 + j];
    const int wp = (w + 1 < TY) ? w + 1 : w;  // the top warp's last row has no row above: its cells are never used
    const int wm = (w > 0) ? w - 1 : w;

    const unsigned plane = (unsigned)nxp * (unsigned)(ny + 2);
    const int c0 = item.z, c1 = item.w;
    // element index of node (k, l0) in array plane p (rows r follow at + r nxp); only dereferenced where owned
    unsigned idx = ((unsigned)(c0 - 1) * (ny + 2) + (unsigned)(l0 + 1)) * nxp + (unsigned)(k + kGhostX);

    double acc[14];
#pragma unroll
    for (int i = 0; i < 14; i++) acc[i] = 0.0;
    // PIN rows of Bq (and entries of c0) as per-thread register values loaded from global memory, the rest through the
    // constant bank / uniform registers (see MM_PIN in mm_march.cuh); per instantiation, because the register budget differs
    double pB[PIN > 0 ? PIN * 6 : 1], pc0[PIN > 0 ? PIN : 1];
    if (PIN > 0) {
#pragma unroll
        for (int i = 0; i < PIN * 6; i++) pB[i] = ld_global_f64(&a.spg->st[0].Bq[(6 - PIN) * 6 + i]);
#pragma unroll
        for (int i = 0; i < PIN; i++) pc0[i] = ld_global_f64(&a.spg->st[0].c0[6 - PIN + i]);
    }

    // carried from plane to plane, per row
    // `original` model (separable stencils): forward xy-combined sums / differences of the previous plane, D' of the previous
    // cell layer.  `default` model (eight corner matrices, nanocell.py:63-132): the four in-plane vertices of the previous
    // plane (00, 10, 01, 11) and the gradients of the previous layer's upper (dz = 1) vertices.
    constexpr bool DEF = MODEL == MM_MODEL_DEFAULT;
    double fpxy[RPT][3], fdxy[RPT][3], fpyd[RPT][3];
    double Dp[RPT][9];
    double rp[DEF ? RPT : 1][4][3], Gp[DEF ? RPT : 1][4][3];
    if (DEF) {
#pragma unroll
        for (int r = 0; r < RPT; r++)
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int j = 0; j < 3; j++) rp[r][c][j] = Gp[r][c][j] = 0.0;
    }
    double vh[RPT][3];                                // STEP: half-kicked velocity of the previous plane
    double vh2[3] = {0, 0, 0}, gdc[3] = {0, 0, 0};    // row 0: ... of the plane before that / its own-row gradient part
#pragma unroll
    for (int r = 0; r < RPT; r++) {
#pragma unroll
        for (int j = 0; j < 3; j++) fpxy[r][j] = fdxy[r][j] = fpyd[r][j] = vh[r][j] = 0.0;
#pragma unroll
        for (int q = 0; q < 9; q++) Dp[r][q] = 0.0;
    }

    const int tid = w * TX + lane;
    unsigned it = 0;  // planes done: plane p = c0 - 1 + it lives in stage it % NST, barrier phase (it / NST) & 1
    constexpr unsigned kStageDoubles = NF * FS;
    if (tid == 0) {
        for (int st = 0; st < NST; st++) mbar_init(smem_u32(&s_full[st]), TY);  // one arrival per warp
        mbar_fence_init();
    }
    if (tid < 27) {
        const int c = tid / 3, j = tid % 3;
        s_shift[c][j] = (double)(c / 3 - 1) * sc.rv[j] + (double)(c % 3 - 1) * sc.rv[3 + j];
    }
    __syncthreads();
    // plane q -> stage st.  Issuing a bulk copy costs the issuing warp on the order of a hundred cycles; with ONE thread
    // issuing all of them the whole block waited for that warp at the next barrier (profiles/r02: barrier stall 1.76 per
    // issue slot with nine main boxes and up to eighteen image copies per plane).  So lane 0 of EVERY warp arms the
    // barrier with the bytes of the fields f = w, w + TY, ... and issues their main box and, in edge tiles, their image
    // column boxes / image row copies / image corners.
    const unsigned field_bytes = 8u * (NR * kBoxW + (hx0 ? 2 * NR : 0) + (hx1 ? 2 * NR : 0) + (hy0 ? kBoxW : 0) + (hy1 ? kBoxW : 0) +
                                       2 * ((hx0 && hy0) + (hx0 && hy1) + (hx1 && hy0) + (hx1 && hy1)));
    const unsigned my_bytes = field_bytes * (unsigned)((NF - w + TY - 1) / TY);  // fields w, w + TY, ... < NF
    auto stage_issue = [&](const int q, const unsigned st) {
        if (lane == 0) {
            const unsigned bar = smem_u32(&s_full[st]);
            mbar_arrive_expect(bar, w < NF ? my_bytes : 0u);
            const int bx0 = bx * OX, by0 = by * OY;
            for (int f = w; f < NF; f += TY) {
                const unsigned fb = smem_u32(s_stage) + (st * kStageDoubles + f * FS) * 8u;
                tma_load_3d(fb, &maps.in[f], bx0, by0, q, bar);
                if (WRAP && edge)
                    issue_edge_copies(&maps, &a, fb, f, Cfg::XW * 8u, Cfg::YW * 8u, Cfg::CW * 8u, 2 * NR * 8u, q, plane, bx0, by0, bar, hmask);
            }
        }
    };
    for (int st = 0; st < NST; st++)
        if (c0 - 1 + st <= c1) stage_issue(c0 - 1 + st, st);

    // Completion of a node of plane q: gradient = lower-row part + own-row part, second kick (verlet.py:152-153), kinetic
    // moments, stores (and delivery to the neighbour slab when q is a boundary plane)
    auto finish_node = [&](auto halo_tag, const bool ownq, const double ownw, const unsigned at, const int q, const double (&glo)[3],
                           const double (&gdq)[3], const double (&vhq)[3]) {
        constexpr bool HALO = decltype(halo_tag)::value;
        double g[3];
#pragma unroll
        for (int j = 0; j < 3; j++) g[j] = glo[j] + gdq[j];
        if (STEP) {
            double vn[3];
#pragma unroll
            for (int j = 0; j < 3; j++) {
                vn[j] = fma(-hminv, g[j], vhq[j]);
                st_if(ownq, a.vo[j] + at, vn[j]);
            }
            if (HALO && a.fused && q == 1) deliver3(a.halo_lo[3], a.halo_lo[4], a.halo_lo[5], at, ownq, vn[0], vn[1], vn[2]);
            if (HALO && a.fused && q == a.nzl) deliver3(a.halo_hi[3], a.halo_hi[4], a.halo_hi[5], at, ownq, vn[0], vn[1], vn[2]);
            const double mo = ownw * mass;
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
        const bool pg = ownq && write_g;
#pragma unroll
        for (int j = 0; j < 3; j++) st_if(pg, a.go[j] + at, g[j]);
        if (HALO && a.fused && q == 1) deliver3(a.halo_lo[6], a.halo_lo[7], a.halo_lo[8], at, pg, g[0], g[1], g[2]);
        if (HALO && a.fused && q == a.nzl) deliver3(a.halo_hi[6], a.halo_hi[7], a.halo_hi[8], at, pg, g[0], g[1], g[2]);
        if (!LEAN) acc[13] = fma(ownw, fma(g[0], g[0], fma(g[1], g[1], g[2] * g[2])), acc[13]);
    };

    // One plane.  CELL: cell layer p-1 (planes p-1 and p) exists; NODE: node plane p-1 (cell layers p-2, p-1) is completed
    // (rows r >= 1 now, row 0 in the next iteration); GATH: row 0 of node plane p-2 is completed in this iteration.
    auto plane_body = [&](auto cell_tag, auto node_tag, auto gath_tag, auto halo_tag, const int p) {
        constexpr bool CELL = decltype(cell_tag)::value, NODE = decltype(node_tag)::value && !VIRIAL;
        constexpr bool GATH = decltype(gath_tag)::value && !VIRIAL;
        constexpr bool HALO = decltype(halo_tag)::value;
        const int par = p & 1;
        const unsigned st = it % NST;
        mbar_wait(smem_u32(&s_full[st]), (it / NST) & 1u);
        const double *sp = s_stage + st * kStageDoubles;
        double r[RPT][3], vcur[RPT][3];
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            const double *sq = sp + off[q];
            double xs = sq[0], ys = sq[FS], zs = sq[2 * FS];
            if (WRAP && edge) {  // block-uniform: only edge tiles hold image nodes
                const double *sh = s_shift[code[q]];
                xs += sh[0];
                ys += sh[1];
                zs += sh[2];
            }
            if (ROT) {
#pragma unroll
                for (int j = 0; j < 3; j++) r[q][j] = fma(zs, R[6 + j], fma(ys, R[3 + j], xs * R[j]));
            } else {
                r[q][0] = xs;
                r[q][1] = ys;
                r[q][2] = zs;
            }
            if (STEP) {
                double cv[3], cg[3];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    cv[d] = sq[(3 + d) * FS];
                    cg[d] = sq[(6 + d) * FS];
                }
#pragma unroll
                for (int j = 0; j < 3; j++) {  // kick + drift (hook scalings + verlet.py:144-146)
                    double vt;
                    if (VM == 2) vt = fma(cv[2], M[6 + j], fma(cv[1], M[3 + j], cv[0] * M[j]));
                    else if (VM == 1) vt = cv[j] * M[0];
                    else vt = cv[j];
                    vcur[q][j] = fma(-hminv, cg[j], vt);
                    r[q][j] = fma(dt, vcur[q][j], r[q][j]);
                }
            }
            if ((STEP || ROT == 2) && CELL) {
                const bool px = own[q] && p < c1;
                const unsigned at = idx + q * nxp;
#pragma unroll
                for (int j = 0; j < 3; j++) st_if(px, a.xo[j] + at, r[q][j]);
                if (HALO && a.fused && p == 1)
                    deliver3(a.halo_lo[0], a.halo_lo[1], a.halo_lo[2], at, px, fma(a.wrap_lo, cw[0], r[q][0]),
                             fma(a.wrap_lo, cw[1], r[q][1]), fma(a.wrap_lo, cw[2], r[q][2]));
                if (HALO && a.fused && p == a.nzl)
                    deliver3(a.halo_hi[0], a.halo_hi[1], a.halo_hi[2], at, px, fma(a.wrap_hi, cw[0], r[q][0]),
                             fma(a.wrap_hi, cw[1], r[q][1]), fma(a.wrap_hi, cw[2], r[q][2]));
            }
        }

        // ---- forward butterfly: y in registers / through shared memory, x by shuffle, z in registers --------------------
#pragma unroll
        for (int j = 0; j < 3; j++) sf[par][j][w][lane] = r[0][j];
        __syncthreads();
        if (p + NST <= c1) stage_issue(p + NST, st);  // all threads have read stage st
        double pxy[RPT][3], dxy[RPT][3], pyd[RPT][3];
        double cur[DEF ? RPT : 1][4][3];  // default model: the four in-plane vertices of this plane
#pragma unroll
        for (int q = 0; q < RPT; q++) {
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const double rn = (q + 1 < RPT) ? r[(q + 1 < RPT) ? q + 1 : q][j] : sf[par][j][wp][lane];
                if (DEF) {
                    cur[q][0][j] = r[q][j];
                    cur[q][2][j] = rn;
                    cur[q][1][j] = __shfl_down_sync(0xffffffffu, r[q][j], 1);
                    cur[q][3][j] = __shfl_down_sync(0xffffffffu, rn, 1);
                    continue;
                }
                const double py = rn + r[q][j], dy = rn - r[q][j];
                const double pyn = __shfl_down_sync(0xffffffffu, py, 1);
                const double dyn = __shfl_down_sync(0xffffffffu, dy, 1);
                pxy[q][j] = py + pyn;  // sum over the four nodes of the cell face in this plane
                dxy[q][j] = pyn - py;  // x difference of the y sums
                pyd[q][j] = dy + dyn;  // y difference of the x sums
            }
        }

        if (GATH) {  // row 0 of node plane p-2: its lower-row part was published before this iteration's barrier
            double glo[3];
#pragma unroll
            for (int j = 0; j < 3; j++) glo[j] = sb[par ^ 1][j][wm][lane];
            finish_node(halo_tag, own[0], ownf[0], idx - 2 * plane, p - 2, glo, gdc, vh2);
        }

        double yp_prev[3] = {0, 0, 0};  // (ya + yb) of the row below inside this thread
#pragma unroll
        for (int q = 0; q < RPT; q++) {
            if (DEF) {
                // ---- `default` model: eight corner evaluations per cell (nanocell.py:63-132) --------------
                // vertex (dx, dy, dz): dz = 0 -> rp[dx + 2 dy] (plane p-1), dz = 1 -> cur[dx + 2 dy] (plane p).  Corner a uses the
                // three cell edges that meet in it, oriented +axis.  Constants are those of the metric form with Hs = H
                // (fold_sparams: Bq x 32, c0 / 16 with respect to the averaged model).
                double Gv[8][3];  // gradient of vertex dx + 2 dy + 4 dz
                if (CELL) {
#pragma unroll
                    for (int v = 0; v < 8; v++)
#pragma unroll
                        for (int j = 0; j < 3; j++) Gv[v][j] = 0.0;
                    double esum = 0.0, vsum[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
                    for (int ca = 0; ca < 8; ca++) {
                        const int ax = ca & 1, ay = (ca >> 1) & 1, az = ca >> 2;
                        // vertex position by (dx, dy, dz)
                        auto V = [&](int dx, int dy, int dz, int j) -> double { return dz ? cur[q][dx + 2 * dy][j] : rp[q][dx + 2 * dy][j]; };
                        double Hs[9];
#pragma unroll
                        for (int j = 0; j < 3; j++) {
                            Hs[j] = V(1, ay, az, j) - V(0, ay, az, j);
                            Hs[3 + j] = V(ax, 1, az, j) - V(ax, 0, az, j);
                            Hs[6 + j] = V(ax, ay, 1, j) - V(ax, ay, 0, j);
                        }
                        double d[6], Sq[6], D[9];
                        d[0] = fma(Hs[2], Hs[2], fma(Hs[1], Hs[1], fma(Hs[0], Hs[0], -Pc.c0[0])));
                        d[1] = fma(Hs[5], Hs[5], fma(Hs[4], Hs[4], fma(Hs[3], Hs[3], -Pc.c0[1])));
                        d[2] = fma(Hs[8], Hs[8], fma(Hs[7], Hs[7], fma(Hs[6], Hs[6], -Pc.c0[2])));
                        d[3] = fma(Hs[5], Hs[8], fma(Hs[4], Hs[7], fma(Hs[3], Hs[6], -Pc.c0[3])));
                        d[4] = fma(Hs[2], Hs[8], fma(Hs[1], Hs[7], fma(Hs[0], Hs[6], -Pc.c0[4])));
                        d[5] = fma(Hs[2], Hs[5], fma(Hs[1], Hs[4], fma(Hs[0], Hs[3], -Pc.c0[5])));
#pragma unroll
                        for (int I = 0; I < 6; I++) {
                            const double *B = Pc.Bq + I * 6;
                            const double lo = fma(B[2], d[2], fma(B[1], d[1], B[0] * d[0]));
                            Sq[I] = fma(B[5], d[5], fma(B[4], d[4], fma(B[3], d[3], lo)));
                        }
                        esum += 0.25 * fma(2.0, fma(d[5], Sq[5], fma(d[4], Sq[4], d[3] * Sq[3])), fma(d[2], Sq[2], fma(d[1], Sq[1], d[0] * Sq[0])));
#pragma unroll
                        for (int j = 0; j < 3; j++) {  // D_a = Sq H_a (= 1/8 V0 h0^-T sym(sigma_a) G_a, nanocell.py:108-130)
                            D[j] = fma(Sq[4], Hs[6 + j], fma(Sq[5], Hs[3 + j], Sq[0] * Hs[j]));
                            D[3 + j] = fma(Sq[3], Hs[6 + j], fma(Sq[1], Hs[3 + j], Sq[5] * Hs[j]));
                            D[6 + j] = fma(Sq[2], Hs[6 + j], fma(Sq[3], Hs[3 + j], Sq[4] * Hs[j]));
                        }
                        if (!VIRIAL) {
#pragma unroll
                            for (int j = 0; j < 3; j++) {  // the edge's upper end gains D_a[i], its lower end loses it
                                Gv[1 + 2 * ay + 4 * az][j] += D[j];
                                Gv[0 + 2 * ay + 4 * az][j] -= D[j];
                                Gv[ax + 2 + 4 * az][j] += D[3 + j];
                                Gv[ax + 0 + 4 * az][j] -= D[3 + j];
                                Gv[ax + 2 * ay + 4][j] += D[6 + j];
                                Gv[ax + 2 * ay + 0][j] -= D[6 + j];
                            }
                        }
                        if (WANT_VIR && decltype(node_tag)::value) {  // sum_a D_a^T H_a
                            vsum[0] += fma(D[6], Hs[6], fma(D[3], Hs[3], D[0] * Hs[0]));
                            vsum[1] += fma(D[7], Hs[7], fma(D[4], Hs[4], D[1] * Hs[1]));
                            vsum[2] += fma(D[8], Hs[8], fma(D[5], Hs[5], D[2] * Hs[2]));
                            vsum[3] += fma(D[7], Hs[8], fma(D[4], Hs[5], D[1] * Hs[2]));
                            vsum[4] += fma(D[6], Hs[8], fma(D[3], Hs[5], D[0] * Hs[2]));
                            vsum[5] += fma(D[6], Hs[7], fma(D[3], Hs[4], D[0] * Hs[1]));
                        }
                    }
                    if (decltype(node_tag)::value) {  // the warm-up layer c0-1 belongs to the chunk below
                        acc[0] = fma(ownf[q], esum, acc[0]);
                        if (WANT_VIR) {
#pragma unroll
                            for (int u = 0; u < 6; u++) acc[1 + u] = fma(ownf[q], vsum[u], acc[1 + u]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int j = 0; j < 3; j++) rp[q][c][j] = cur[q][c][j];
                // backward: node (lane, row, p-1) gathers vertex (dx, dy, dz) of cell (lane - dx, row - dy, layer p-1-dz):
                // z in registers (Gp), x by shuffle, y in registers / through shared memory
                if (NODE) {
                    double gd[3], yp[3];
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        const double z00 = Gp[q][0][j] + Gv[0][j], z10 = Gp[q][1][j] + Gv[1][j];
                        const double z01 = Gp[q][2][j] + Gv[2][j], z11 = Gp[q][3][j] + Gv[3][j];
                        gd[j] = z00 + __shfl_up_sync(0xffffffffu, z10, 1);  // cells of this row (dy = 0)
                        yp[j] = z01 + __shfl_up_sync(0xffffffffu, z11, 1);  // what the row above gathers (dy = 1)
                    }
                    if (q == 0) {
#pragma unroll
                        for (int j = 0; j < 3; j++) {
                            gdc[j] = gd[j];
                            vh2[j] = vh[0][j];
                        }
                    } else {
                        finish_node(halo_tag, own[q], ownf[q], idx - plane + q * nxp, p - 1, yp_prev, gd, vh[q]);
                    }
                    if (q == RPT - 1) {
#pragma unroll
                        for (int j = 0; j < 3; j++) sb[par][j][w][lane] = yp[j];
                    }
#pragma unroll
                    for (int j = 0; j < 3; j++) yp_prev[j] = yp[j];
                }
                if (CELL && !VIRIAL) {
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int j = 0; j < 3; j++) Gp[q][c][j] = Gv[4 + c][j];
                }
                if (STEP) {
#pragma unroll
                    for (int j = 0; j < 3; j++) vh[q][j] = vcur[q][j];
                }
                continue;
            }
            double D[9];
            if (CELL) {
                double Hs[9];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    Hs[j] = fdxy[q][j] + dxy[q][j];      // 4 * (mean x edge)
                    Hs[3 + j] = fpyd[q][j] + pyd[q][j];  // 4 * (mean y edge)
                    Hs[6 + j] = pxy[q][j] - fpxy[q][j];  // 4 * (mean z edge)
                }
                double e, vir[6], Sq[6], d[6], c0v[6];
#pragma unroll
                for (int u = 0; u < 6; u++) c0v[u] = (u >= 6 - PIN) ? pc0[(u >= 6 - PIN) ? u - (6 - PIN) : 0] : Pc.c0[u];
                d[0] = fma(Hs[2], Hs[2], fma(Hs[1], Hs[1], fma(Hs[0], Hs[0], -c0v[0])));
                d[1] = fma(Hs[5], Hs[5], fma(Hs[4], Hs[4], fma(Hs[3], Hs[3], -c0v[1])));
                d[2] = fma(Hs[8], Hs[8], fma(Hs[7], Hs[7], fma(Hs[6], Hs[6], -c0v[2])));
                d[3] = fma(Hs[5], Hs[8], fma(Hs[4], Hs[7], fma(Hs[3], Hs[6], -c0v[3])));
                d[4] = fma(Hs[2], Hs[8], fma(Hs[1], Hs[7], fma(Hs[0], Hs[6], -c0v[4])));
                d[5] = fma(Hs[2], Hs[5], fma(Hs[1], Hs[4], fma(Hs[0], Hs[3], -c0v[5])));
#pragma unroll
                for (int I = 0; I < 6; I++) {  // Sq = Bq d: two chains of three (halves the dependent-FMA depth)
                    const bool pin = I >= 6 - PIN;
                    const double *B = pin ? pB + (pin ? I - (6 - PIN) : 0) * 6 : Pc.Bq + I * 6;
                    const double lo = fma(B[2], d[2], fma(B[1], d[1], B[0] * d[0]));
                    Sq[I] = fma(B[5], d[5], fma(B[4], d[4], fma(B[3], d[3], lo)));
                }
                // elastic energy 1/4 d:Sq (efree: added once per owned column after the plane loop)
                e = 0.25 * fma(2.0, fma(d[5], Sq[5], fma(d[4], Sq[4], d[3] * Sq[3])), fma(d[2], Sq[2], fma(d[1], Sq[1], d[0] * Sq[0])));
#pragma unroll
                for (int j = 0; j < 3; j++) {  // D' = Sq Hs with Sq = [[0 5 4], [5 1 3], [4 3 2]]
                    D[j] = fma(Sq[4], Hs[6 + j], fma(Sq[5], Hs[3 + j], Sq[0] * Hs[j]));
                    D[3 + j] = fma(Sq[3], Hs[6 + j], fma(Sq[1], Hs[3 + j], Sq[5] * Hs[j]));
                    D[6 + j] = fma(Sq[2], Hs[6 + j], fma(Sq[3], Hs[3 + j], Sq[4] * Hs[j]));
                }
                if (decltype(node_tag)::value) {  // the warm-up layer c0-1 belongs to the chunk below
                    acc[0] = fma(ownf[q], e, acc[0]);
                    if (WANT_VIR) {  // cell virial D'^T Hs (mmff.py:320-323; symmetric because Sq is)
                        vir[0] = fma(D[6], Hs[6], fma(D[3], Hs[3], D[0] * Hs[0]));
                        vir[1] = fma(D[7], Hs[7], fma(D[4], Hs[4], D[1] * Hs[1]));
                        vir[2] = fma(D[8], Hs[8], fma(D[5], Hs[5], D[2] * Hs[2]));
                        vir[3] = fma(D[7], Hs[8], fma(D[4], Hs[5], D[1] * Hs[2]));
                        vir[4] = fma(D[6], Hs[8], fma(D[3], Hs[5], D[0] * Hs[2]));
                        vir[5] = fma(D[6], Hs[7], fma(D[3], Hs[4], D[0] * Hs[1]));
#pragma unroll
                        for (int u = 0; u < 6; u++) acc[1 + u] = fma(ownf[q], vir[u], acc[1 + u]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 3; j++) {
                fpxy[q][j] = pxy[q][j];
                fdxy[q][j] = dxy[q][j];
                fpyd[q][j] = pyd[q][j];
            }
            // ---- backward butterfly of node plane p-1 (see k_march): z in registers, x by shuffle, y in registers /
            // through shared memory ----------------------------------------------------------------------------------------
            if (NODE) {
                double gd[3], yp[3];
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    const double p0 = Dp[q][j] + D[j], p1 = Dp[q][3 + j] + D[3 + j], p2 = Dp[q][6 + j] - D[6 + j];
                    const double s02m = __shfl_up_sync(0xffffffffu, p0 + p2, 1);
                    const double p1m = __shfl_up_sync(0xffffffffu, p1, 1);
                    const double ya = s02m + (p2 - p0), yb = p1m + p1;
                    yp[j] = ya + yb;
                    gd[j] = ya - yb;
                }
                if (q == 0) {
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        gdc[j] = gd[j];
                        vh2[j] = vh[0][j];
                    }
                } else {
                    finish_node(halo_tag, own[q], ownf[q], idx - plane + q * nxp, p - 1, yp_prev, gd, vh[q]);
                }
                if (q == RPT - 1) {
#pragma unroll
                    for (int j = 0; j < 3; j++) sb[par][j][w][lane] = yp[j];
                }
#pragma unroll
                for (int j = 0; j < 3; j++) yp_prev[j] = yp[j];
            }
            if (CELL && !VIRIAL) {
#pragma unroll
                for (int u = 0; u < 9; u++) Dp[q][u] = D[u];
            }
            if (STEP) {
#pragma unroll
                for (int j = 0; j < 3; j++) vh[q][j] = vcur[q][j];
            }
        }
        idx += plane;
        it++;
    };

    // The boundary planes of a slab (1, nzl: positions; one or two iterations later: velocities, gradients) can only come
    // up in the first four and the last two iterations of a chunk: only those bodies carry the delivery code.
    constexpr std::true_type T{};
    constexpr std::false_type F{};
    plane_body(F, F, F, T, c0 - 1);
    plane_body(T, F, F, T, c0);
    plane_body(T, T, F, T, c0 + 1);
    if (c0 + 2 <= c1) plane_body(T, T, T, T, c0 + 2);  // completes row 0 of node plane c0 (the lower boundary plane of the first chunk)
    // UNR = 2: two planes per trip - the carried values (48 doubles) rotate through register renaming instead of ~90 moves
#pragma unroll UNR
    for (int p = c0 + 3; p <= c1 - 2; p++) plane_body(T, T, T, F, p);
#pragma unroll 1
    for (int p = max(c0 + 3, c1 - 1); p <= c1; p++) plane_body(T, T, T, T, p);
    if (!VIRIAL) {
        // drain: row 0 of node plane c1-1 (its lower-row part was published in iteration c1)
        __syncthreads();
        double glo[3];
#pragma unroll
        for (int j = 0; j < 3; j++) glo[j] = sb[c1 & 1][j][wm][lane];
        finish_node(T, own[0], ownf[0], idx - 2 * plane, c1 - 1, glo, gdc, vh2);
    }

    {
        int nown = 0;
#pragma unroll
        for (int r = 0; r < RPT; r++) nown += own[r] ? 1 : 0;
        acc[0] = fma((double)((c1 - c0) * nown), Pc.efree, acc[0]);
    }
    // block reduction: warp shuffles, then one warp over the per-warp sums
    __shared__ double red[TY][14];
#pragma unroll
    for (int q = 0; q < 14; q++) {
        const double s = warp_sum(acc[q]);
        if (lane == 0) red[w][q] = s;
    }
    __syncthreads();
    if (w == 0 && lane < 14) {
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < TY; u++) s += red[u][lane];
        const int bid = blockIdx.x;
        a.partials[(size_t)bid * kRedSlots + lane] = s;
    }
    if (a.tail.enabled) {
        // Ticket.  Only the block partial has to be visible to the last block (warp 0 wrote it: warp 0 fences); the node
        // arrays are consumed by the NEXT launch.  A fence by every thread would wait for the block's whole store queue to
        // drain at the end of every block (measured: +10 us per launch).  The exception are the boundary planes a slab stores
        // into its neighbours' memory: they must have landed before the tail's exchange tells the neighbours so.
        const int nblocks = gridDim.x;
        if (a.fused && a.tail.nranks > 1 && (c0 == 1 || c1 == a.nzl + 1)) {
            __threadfence_system();
            __syncthreads();
        }
        if (w == 0) {
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const unsigned ticket = atomicAdd(a.tail.counter, 1u);
                s_last = ticket == (unsigned)(nblocks - 1);
                __threadfence();
            }
        }
        __syncthreads();
        if (s_last) march_tail(a.tail, a.partials, nblocks);
    }
}

}  // namespace mm
