// mm_internal.h - host-side structures shared by the translation units of libmicmec_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/micmec_b200.h"
#include "mm_cell.cuh"
#include "mm_structured.cuh"

namespace mm {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t err, const char *what);

#define MM_CUDA(call)                                            \
    do {                                                         \
        cudaError_t err__ = (call);                              \
        if (err__ != cudaSuccess) return mm::cuda_fail(err__, #call); \
    } while (0)

constexpr int kRedSlots = 16;        // doubles per block partial
constexpr int kMaxRedBlocks = 2048;  // upper bound on reducing grids

// Results of the last force evaluation, device resident (and mirrored to pinned host memory on request)
struct ForceResult {
    double epot;
    double vir[6];   // 00,11,22,12,02,01
    double sum_g2;   // sum of squared gradient components (rmsd_gpos), only when requested
    double pad[8];
};

}  // namespace mm

struct mm_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t nnodes = 0, ncells = 0;
    int model = 0;
    double boltzmann = 0.0;
    int structured = 0;  // full periodic grid with implied topology
    int nx = 0, ny = 0, nz = 0;
    int num_sms = 148;
    mm::KParams kp;
    // indexed topology (SoA: [8][ncells] / [8][nnodes]); int32 on device
    int32_t *d_cell_nodes = nullptr;
    int32_t *d_node_cells = nullptr;
    uint8_t *d_cell_info = nullptr;  // bits 0-3 type, bits 4-6 periodic wrap flags along a,b,c
    // geometry + outputs
    double *d_pos = nullptr;         // [nnodes][3]
    double *d_gpos = nullptr;        // [nnodes][3]
    double *d_gcell = nullptr;       // [24][ncells] per-cell gradients (SoA)
    double *d_ecell = nullptr;       // [ncells]
    double *d_partials = nullptr;    // [kMaxRedBlocks][kRedSlots]
    mm::ForceResult *d_result = nullptr;
    mm::ForceResult *h_result = nullptr;  // pinned
    double *h_stage = nullptr;            // pinned staging for pos / gpos host transfers
    size_t h_stage_bytes = 0;
    double rvecs[9] = {0};
    double *d_rvecs = nullptr;  // [9] device copy read by kernels (device-resident MD updates it in place)
    int64_t launches = 0;
    int scatter_mode = 0;
    int profile = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;  // one pair per timed force-kernel launch
    std::vector<int> prof_kinds;
    bool pos_valid = false;
    int want_structured = 1;  // use the structured-grid kernels when the system allows it
    // z-slab decomposition (mm_comm.cu): this handle owns nz planes of a grid that is slab_count slabs tall
    int slab_rank = 0, slab_count = 1;
    int64_t nnodes_global = 0;
    void *comm = nullptr;     // ncclComm_t
    double *d_red = nullptr;  // [32] reduction buffer all-reduced across the slabs
    // batch of independent replicas (config 5): per-replica domain vectors, per-cell virial, per-replica results
    int64_t nreplicas = 1;
    double *d_rvecs_batch = nullptr;  // [nreplicas][9]
    double *d_vcell = nullptr;        // [6][ncells]
    double *d_rep = nullptr;          // [nreplicas][8]: energy, virial(6), pad
    double *d_halo = nullptr; // packed halo messages: send up / send down / recv from down / recv from up, 9 planes each
    // peer mode (mm_comm.cu): halo planes and the 16-double sums are stored directly into the neighbours' memory over
    // NVLink (CUDA IPC mappings of one block per rank) and handed over with system-scope flags; NCCL only bootstraps
    int peer_mode = 0;
    char *d_peer = nullptr;         // this rank's peer-visible block: flags, mailboxes, inboxes
    char *peer_base[16] = {nullptr};  // mapped base of every rank's block (own entry = d_peer)
    char **d_peer_base = nullptr;   // device copy of peer_base
    void *d_peer_ctl = nullptr;     // local epoch counters (PeerCtl)
    int64_t peer_plane = 0;         // padded plane (nodes) the inboxes were sized for
    void *peer_soa[2] = {nullptr, nullptr};  // IPC mappings of the neighbours' node-array blocks (fused halo, sg.nb_block)
    mm::SGrid sg;
};

namespace mm {

// ---- mm_force.cu ------------------------------------------------------------------------------------------
// Evaluate energy / per-cell gradients / virial partials at the positions in h->d_pos with the cell in
// h->d_rvecs; gather node gradients into `gpos_out` (device, [nnodes][3]) when non-null; leave the reduced
// ForceResult in h->d_result.  want_g2 additionally reduces sum(gpos^2).
int force_evaluate(mm_handle *h, double *gpos_out, bool want_g2);
int grid_for(const mm_handle *h, int64_t n, int threads);
// cell kernel only; returns the number of blocks whose partials (energy + virial) were written to h->d_partials
int cells_launch(mm_handle *h);
// index arrays / per-cell buffers of the indexed kernels (built lazily for structured grids; mm_api.cu)
int ensure_generic(mm_handle *h);
// event bracket around the dominant kernel when h->profile is on
void prof_begin(mm_handle *h, int kind);  // kind 0: force-only kernel, 1: fused step kernel
void prof_end(mm_handle *h);
void final_launch(mm_handle *h, const double *pc, int nbc, const double *pn, int nbn);

// ---- mm_structured.cu ---------------------------------------------------------------------------------------
bool sg_eligible(const mm_handle *h);
int sg_setup(mm_handle *h);
void sg_free(mm_handle *h);
int sg_write_consts(mm_handle *h, const double *rvecs9, double dt);
int sg_halo(mm_handle *h, bool pos, bool vel, bool grad, const double *rv_src = nullptr, cudaStream_t st = nullptr);
int sg_pos_from_aos(mm_handle *h, const double *d_aos);
int sg_vel_from_aos(mm_handle *h, const double *d_aos);
int sg_mass_from_aos(mm_handle *h, const double *d_masses);
int sg_to_aos(mm_handle *h, int which, double *d_aos);
// what the tail of a k_march2 launch does after the reduction (mm_march2.cuh): OP_* bits of mm_scalar.cuh on the MD state
struct SgTail {
    unsigned ops;
    void *state;
};
bool sg_tail_ok(const mm_handle *h);
int sg_force(mm_handle *h, bool write_g, int rot, bool virial_only = false, const SgTail *tail = nullptr);
int sg_step(mm_handle *h, bool write_g, int vm, bool lean, const SgTail *tail = nullptr);
int sg_set_tile_rows(mm_handle *h, int rows);
int sg_set_chunk(mm_handle *h, int chunk);
int sg_set_rpt(mm_handle *h, int rpt);
int sg_set_march2(mm_handle *h, int on);
int sg_retile(mm_handle *h, int chunk_override);
void sg_plan_schedule(int ntx, int nty, int P, int S, int images_on_load, int uniform_chunk, std::vector<int4> &items,
                      double *cost, double *ideal, int *chunk_b);

// ---- mm_comm.cu ---------------------------------------------------------------------------------------------
int comm_halo(mm_handle *h, double **fields, int nfields, int npos);
int comm_allreduce(mm_handle *h, double *buf, int count);
int comm_reduce_partials(mm_handle *h, const double *pc, int nbc, const double *pn, int nbn, const double *pd, int nbd);
void comm_peer_free(mm_handle *h);

}  // namespace mm
