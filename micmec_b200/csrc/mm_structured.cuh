// mm_structured.cuh - data structures of the structured-grid (full periodic nx x ny x nz) fast path
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/micmec_b200.h"

namespace mm {

constexpr int kGhostX = 2;  // padded column of node k = 0
constexpr int kBoxW = 34;   // width of the TMA load box of a 32-lane tile (starts one node before the tile: even column)

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

// SoA planes in z-major order, padded by one ghost node on each side of every axis (periodic images, kept current by
// sg_halo; positions of ghosts already carry the domain-vector shift):
//   index(k, l, p) = (p * (ny + 2) + l + 1) * nxp + k + kGhostX,   k = -1 .. nx, l = -1 .. ny, p = 0 (lower halo), 1..nzl
//   (owned), nzl + 1 (upper halo); nxp = nx + kGhostX + 1 rounded up to an even pitch.  kGhostX = 2 (one unused column before
//   the x ghost): TMA boxes of 8-byte elements must start on 16 bytes, i.e. on an EVEN column (measured:
//   profiles/microbench/tma_store_probe.cu), and with two columns in front both the load box of a tile (node k0 - 2, 34
//   wide) and its store box (first owned node k0, 30 wide) do.  A z-slab of a multi-GPU decomposition is one contiguous
// range of planes and its halos are whole (padded) planes; a tile of the marching kernel never wraps.
struct SGrid {
    int active = 0;
    int nx = 0, ny = 0, nzl = 0;
    int nxp = 0;                    // row pitch in nodes
    int64_t plane = 0, npad = 0;    // padded plane (nxp * (ny + 2)) and array size in nodes
    double *block = nullptr;        // ONE allocation behind the 20 arrays below (array a at block + a * stride): the peers
    int64_t stride = 0;             // of a z-slab map it with a single CUDA IPC handle (array order: see sg_array)
    double *x[2][3] = {{nullptr}};  // ping-pong: kernels that move nodes read one set and write the other
    double *v[2][3] = {{nullptr}};
    double *g[2][3] = {{nullptr}};
    double *m = nullptr, *minv = nullptr;
    uint8_t *type = nullptr;
    int cx = 0, cv = 0, cg = 0;     // which copy is current
    CUtensorMap tm_x[2][3], tm_v[2][3], tm_g[2][3], tm_m, tm_minv;  // load descriptors of the arrays above
    CUtensorMap tw_x[2][3], tw_v[2][3], tw_g[2][3];                 // ... with the 2-column box of k_march2's image columns
    unsigned int *d_tail_counter = nullptr;                         // k_march2 tail: blocks finished
    int4 *d_items = nullptr;                                        // k_march2 work items (sg_plan_items)
    int nitems = 0, nitems_alloc = 0, ntx = 0, nty = 0;
    int plan_two_class = 1;                                         // option "plan": 0 = uniform chunks for every tile
    double plan_cost = 0.0, plan_ideal = 0.0;                       // simulated makespan / perfect balance, in plane iterations
    int tma_ok = 0;                 // descriptors encoded (driver entry point available, pitch constraints met)
    // fused halo (see MarchArgs): where the z neighbours' copies of this slab's arrays live (own block when there is one
    // slab), their plane counts, and what the last marching launch already delivered
    double *nb_block[2] = {nullptr, nullptr};  // [0] below (rank - 1), [1] above (rank + 1)
    int64_t nb_stride[2] = {0, 0};
    int nb_nzl[2] = {0, 0};
    unsigned long long *nb_flag[2] = {nullptr, nullptr};  // the neighbours' arrival counters for planes coming from this rank
    unsigned long long *halo_flags = nullptr;  // own arrival counters [from below, from above] (slabs; null for one slab)
    unsigned long long *halo_epoch = nullptr;  // fused exchanges completed (device counter, advanced by k_halo_xy_fused)
    unsigned int *halo_done = nullptr;         // its block counter
    int fused = 0;                  // k_march also writes the boundary planes into the neighbours' halo planes
    int fused_mask = 0;             // fields delivered by the last launch: 1 positions, 2 velocities, 4 gradients
    int tail_done = 0;              // ... and its tail already exchanged with the other slabs (k_march2)
    StepConsts *d_sc = nullptr;
    SParams *d_sp = nullptr;        // device copy of sp
    StepConsts *h_sc = nullptr;     // pinned staging
    double *d_partials = nullptr;   // [nblocks][kRedSlots]
    int nblocks = 0, nblocks_alloc = 0;
    int tile_rows = 8;              // TY of the marching kernel (warps per block)
    // fast path (mm_march2.cuh): one type, one state, uniform node mass, tensor maps available
    int march2 = 0;                 // the handle's launches go to k_march2
    int march2_wanted = 1;          // option "march2": 0 keeps everything on k_march
    int wrap_on_load = -1;          // k_march2 takes the x / y periodic images on load instead of from ghost nodes; -1 = decide
                                    // from the decomposition (sg_retile: slabs yes, a single GPU no - measured in profiles/r02)
    int wrap_wanted = -1;           // option "wrap_on_load"
    int tail_in_kernel = 0;         // the tail runs inside the marching launch (last block) instead of as its own launch
    int tail_wanted = 1;            // option "tail": 0 keeps the reduction / exchange / scalar algebra in their own launches
    int rpt = 2;                    // node rows per thread of k_march2 (tile = 32 x rpt * tile_rows nodes)
    int tma_rows = 0;               // box rows the tensor maps were encoded for
    int mass_uniform = 1;           // all node masses equal (checked when the masses are uploaded)
    double mass = 1.0;
    int variant = 14;               // tuning bits of the marching kernel (k_march VAR); default: TMA loads, one barrier
                                    // per plane, two planes per loop trip
    int chunk = 32;                 // owned planes per block along z
    SParams sp;
};

// Tensor maps (TMA descriptors) of the arrays one launch of the staged kernel variant reads: box = kBoxW x TY x 1 nodes of
// a padded plane; coordinates outside the array read as zero (the partial tiles at the upper x / y edge)
struct alignas(64) TmaMaps {
    CUtensorMap in[11];  // x0 x1 x2 v0 v1 v2 g0 g1 g2 m 1/m
    CUtensorMap xw[9];   // k_march2: the same nine arrays with a box of 2 x rows nodes (the periodic image column of an edge tile)
};

// Tail of a k_march2 launch: the last block to finish sums the block partials in a fixed order, exchanges the 16 sums with
// the other z-slabs through the peer mailboxes (mm_comm.cu; the exchange also tells every rank that its neighbours'
// boundary planes have landed) and runs the thermostat / barostat algebra (mm_scalar.cuh) - what used to be the
// k_peer_allreduce, k_scalar and k_halo_xy_fused launches between two marching kernels.
struct TailArgs {
    int enabled;
    unsigned ops;            // OP_* bits of mm_scalar.cuh
    unsigned int *counter;   // blocks finished (zeroed by the tail)
    void *state;             // MDState*
    double *rvecs_dev;       // the handle's device copy of the domain vectors (the barostat updates it)
    StepConsts *sc_out;
    double n3;               // 3 x global number of nodes
    int nranks, rank;
    char *const *bases;      // peer control blocks of all slabs (null for one slab)
    void *ctl;               // PeerCtl*
};

struct MarchArgs {
    int nx, ny, nzl, chunk;
    int nxp;      // row pitch of the padded planes
    const double *x[3];
    double *xo[3];
    const double *v[3];
    double *vo[3];
    const double *g[3];
    double *go[3];
    const double *m, *minv;
    const uint8_t *type;
    // Fused halo exchange: the two boundary planes of everything a launch writes are ALSO stored into the halo planes
    // of the z neighbours (the own ones when the grid is a single slab), through peer mappings over NVLink, as they are
    // produced.  Pointers are biased so that element idx of the own plane lands on the same node of the target plane;
    // wrap = -1 / +1 where the neighbour is the periodic image across the ring of slabs (positions get -c / +c).  The
    // kernel that follows on the stream (k_halo_xy_fused) announces the delivery to the neighbours and waits for theirs.
    double *halo_lo[9];   // own plane 1   -> plane nzl' + 1 of the neighbour below   (x0 x1 x2 v0 v1 v2 g0 g1 g2 as written)
    double *halo_hi[9];   // own plane nzl -> plane 0 of the neighbour above
    double wrap_lo, wrap_hi;
    int fused;
    double mass;  // k_march2: the uniform node mass
    const int4 *items;  // k_march2: work items (tile x, tile y, first owned plane, one past the last), one per block
    int ntx, nty;       // ... and the number of tiles along x / y
    TailArgs tail;
    const StepConsts *sc;
    const SParams *spg;  // the constants once more in global memory (pinned register copies are loaded from here)
    double *partials;
};

}  // namespace mm
