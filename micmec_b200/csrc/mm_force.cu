// mm_force.cu - force / energy / virial kernels for an indexed (arbitrary) MicMec topology.
//
// Stands in for ForcePartMechanical._internal_compute + deformation() + _compute_gpos/_compute_vtens/_compute_epot
// (micmec/pes/mmff.py:288-403).  Two kernels per evaluation:
//   k_cells   one thread per cell (grid-stride): gather the 8 vertices, unwrap with the minimum-image flags
//             (mmff.py:347-371), evaluate every metastable state + Boltzmann mixing (mm_cell.cuh), write the
//             per-cell gradient (SoA [24][ncells]) and energy, block-reduce energy + virial with warp shuffles
//   k_gather  one thread per node: sum the <= 8 incident per-cell gradients in the reference's fixed order
//             j = 0..7 (mmff.py:303-318) -> deterministic, bit-reproducible; optional sum(g^2)
//   k_final   one block: add the per-block partials in a fixed order
// HBM traffic is dominated by the 24-double per-cell gradient (written once, read once); the structured-grid path
// in mm_structured.cu avoids materialising it.
#include "mm_internal.h"
#include "mm_reduce.cuh"

namespace mm {

constexpr int kThreads = 128;

// BATCH: the system is a batch of independent replicas of `cpr` cells each; replica r has its own domain vectors
// rvecs[r][9] and its per-cell virial is kept (vcell, SoA [6][ncells]) for the per-replica reduction.
template <int MODEL, bool BATCH>
__global__ void __launch_bounds__(kThreads)
k_cells(const __grid_constant__ KParams kp, const int32_t *__restrict__ cell_nodes,
        const uint8_t *__restrict__ cell_info, const double *__restrict__ pos, const double *__restrict__ rvecs,
        int64_t ncells, double *__restrict__ gcell, double *__restrict__ ecell, double *__restrict__ partials,
        int64_t cpr, double *__restrict__ vcell) {
    double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double rv[9];
    if (!BATCH) {
#pragma unroll
        for (int i = 0; i < 9; i++) rv[i] = rvecs[i];
    }
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (int64_t)gridDim.x * blockDim.x) {
        const unsigned info = cell_info[c];
        const int type = info & 15;
        if (BATCH) {
            const double *rr = rvecs + 9 * (c / cpr);
#pragma unroll
            for (int i = 0; i < 9; i++) rv[i] = rr[i];
        }
        double R[24];
        const int64_t n0 = cell_nodes[c];
        const double r0x = pos[3 * n0], r0y = pos[3 * n0 + 1], r0z = pos[3 * n0 + 2];
        R[0] = r0x;
        R[1] = r0y;
        R[2] = r0z;
#pragma unroll
        for (int v = 1; v < 8; v++) {
            const int64_t n = cell_nodes[(int64_t)v * ncells + c];
            // r_v = r_0 + (pos[v] - pos[0] + sum_a rvecs[a] mic[0, v, a])      mmff.py:347-371
            const double sa = (vbit(v, 0) && (info & 16)) ? 1.0 : 0.0;
            const double sb = (vbit(v, 1) && (info & 32)) ? 1.0 : 0.0;
            const double sc = (vbit(v, 2) && (info & 64)) ? 1.0 : 0.0;
#pragma unroll
            for (int d = 0; d < 3; d++) {
                double dvec = pos[3 * n + d] - R[d];
                dvec += rv[d] * sa + rv[3 + d] * sb + rv[6 + d] * sc;
                R[v * 3 + d] = R[d] + dvec;
            }
        }
        double e, g[24], vir[6];
        cell_eval<MODEL>(R, kp, type, e, g, vir);
        ecell[c] = e;
#pragma unroll
        for (int k = 0; k < 24; k++) gcell[(int64_t)k * ncells + c] = g[k];
        acc[0] += e;
#pragma unroll
        for (int k = 0; k < 6; k++) acc[1 + k] += vir[k];
        if (BATCH) {
#pragma unroll
            for (int k = 0; k < 6; k++) vcell[(int64_t)k * ncells + c] = vir[k];
        }
    }
    block_sum_store<7>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

// per-replica energy and virial: one thread per replica sums its cells in index order (deterministic)
__global__ void __launch_bounds__(128)
k_replica_reduce(const double *__restrict__ ecell, const double *__restrict__ vcell, int64_t ncells, int64_t cpr,
                 int64_t nrep, double *__restrict__ out /* [nrep][8] */) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrep; r += (int64_t)gridDim.x * blockDim.x) {
        double s[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int64_t c = r * cpr; c < (r + 1) * cpr; c++) {
            s[0] += ecell[c];
#pragma unroll
            for (int k = 0; k < 6; k++) s[1 + k] += vcell[(int64_t)k * ncells + c];
        }
#pragma unroll
        for (int k = 0; k < 7; k++) out[r * 8 + k] = s[k];
    }
}

// ---- the alternative gpos accumulation: cell-centric scatter with warp-aggregated atomics -------------------------------
// Same per-cell work as k_cells, but instead of writing 24 gradient doubles per cell for a later node-centric
// gather, every cell adds its 8 vertex contributions straight into gpos with fp64 atomics (RED.ADD.F64 resolved in
// L2).  Consecutive cells of the reference enumeration are z-neighbours and share four vertices: lane i's
// contributions to its upper-z vertices (3, 5, 6, 7) are handed to lane i+1 with shuffles whenever that lane's
// lower-z vertices (0, 1, 2, 4) are the same nodes, which halves the atomics.  Summation order is not reproducible.
template <int MODEL>
__global__ void __launch_bounds__(kThreads)
k_cells_scatter(const __grid_constant__ KParams kp, const int32_t *__restrict__ cell_nodes,
                const uint8_t *__restrict__ cell_info, const double *__restrict__ pos, const double *__restrict__ rvecs,
                int64_t ncells, double *__restrict__ gpos, double *__restrict__ ecell, double *__restrict__ partials) {
    double acc[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double rv[9] = {rvecs[0], rvecs[1], rvecs[2], rvecs[3], rvecs[4], rvecs[5], rvecs[6], rvecs[7], rvecs[8]};
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nloop = (ncells + stride - 1) / stride * stride;  // whole warps stay in the loop (shuffles)
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nloop; c += stride) {
        const bool live = c < ncells;
        int32_t nd[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        double g[24];
#pragma unroll
        for (int k = 0; k < 24; k++) g[k] = 0.0;
        if (live) {
            const unsigned info = cell_info[c];
            double R[24];
#pragma unroll
            for (int v = 0; v < 8; v++) nd[v] = cell_nodes[(int64_t)v * ncells + c];
            R[0] = pos[3 * (int64_t)nd[0]];
            R[1] = pos[3 * (int64_t)nd[0] + 1];
            R[2] = pos[3 * (int64_t)nd[0] + 2];
#pragma unroll
            for (int v = 1; v < 8; v++) {
                const double sa = (vbit(v, 0) && (info & 16)) ? 1.0 : 0.0;
                const double sb = (vbit(v, 1) && (info & 32)) ? 1.0 : 0.0;
                const double sc = (vbit(v, 2) && (info & 64)) ? 1.0 : 0.0;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    double dvec = pos[3 * (int64_t)nd[v] + d] - R[d];
                    dvec += rv[d] * sa + rv[3 + d] * sb + rv[6 + d] * sc;
                    R[v * 3 + d] = R[d] + dvec;
                }
            }
            double e, vir[6];
            cell_eval<MODEL>(R, kp, info & 15, e, g, vir);
            ecell[c] = e;
            acc[0] += e;
#pragma unroll
            for (int k = 0; k < 6; k++) acc[1 + k] += vir[k];
        }
        // warp aggregation along the enumeration: (upper-z vertex of lane i) == (lower-z vertex of lane i+1)?
        constexpr int UP[4] = {3, 5, 6, 7}, LO[4] = {0, 1, 2, 4};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int lo_next = __shfl_down_sync(0xffffffffu, nd[LO[q]], 1);
            const bool give = lane < 31 && nd[UP[q]] >= 0 && nd[UP[q]] == lo_next;   // I hand my contribution over
            const bool take = __shfl_up_sync(0xffffffffu, (int)give, 1) && lane > 0;  // my lower neighbour hands me its own
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double from_below = __shfl_up_sync(0xffffffffu, g[UP[q] * 3 + d], 1);
                if (take) g[LO[q] * 3 + d] += from_below;
                if (give) g[UP[q] * 3 + d] = 0.0;
            }
            if (give) nd[UP[q]] = -1;
        }
#pragma unroll
        for (int v = 0; v < 8; v++)
            if (nd[v] >= 0) {
#pragma unroll
                for (int d = 0; d < 3; d++) atomicAdd(&gpos[3 * (int64_t)nd[v] + d], g[v * 3 + d]);
            }
    }
    block_sum_store<7>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

__global__ void __launch_bounds__(256)
k_sumsq(const double *__restrict__ a, int64_t n, double *__restrict__ partials) {
    double acc[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc[0] = fma(a[i], a[i], acc[0]);
    block_sum_store<1>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

template <bool WANT_G2>
__global__ void __launch_bounds__(kThreads)
k_gather(const int32_t *__restrict__ node_cells, const double *__restrict__ gcell, int64_t nnodes, int64_t ncells,
         double *__restrict__ gpos, double *__restrict__ partials) {
    double acc[1] = {0.0};
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (int64_t)gridDim.x * blockDim.x) {
        double gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int64_t c = node_cells[(int64_t)j * nnodes + n];
            if (c >= 0) {
                gx += gcell[(int64_t)(3 * j) * ncells + c];
                gy += gcell[(int64_t)(3 * j + 1) * ncells + c];
                gz += gcell[(int64_t)(3 * j + 2) * ncells + c];
            }
        }
        gpos[3 * n] = gx;
        gpos[3 * n + 1] = gy;
        gpos[3 * n + 2] = gz;
        if (WANT_G2) acc[0] += gx * gx + gy * gy + gz * gz;
    }
    if (WANT_G2) block_sum_store<1>(acc, partials + (size_t)blockIdx.x * kRedSlots);
}

__global__ void __launch_bounds__(256)
k_final(const double *__restrict__ pc, int nbc, const double *__restrict__ pn, int nbn, ForceResult *res) {
    double r[7];
    partials_sum<7>(pc, nbc, kRedSlots, r);
    double g2[1] = {0.0};
    if (nbn > 0) partials_sum<1>(pn, nbn, kRedSlots, g2);
    if (threadIdx.x == 0) {
        res->epot = r[0];
#pragma unroll
        for (int k = 0; k < 6; k++) res->vir[k] = r[1 + k];
        res->sum_g2 = g2[0];
    }
}

int grid_for(const mm_handle *h, int64_t n, int threads) {
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)h->num_sms * 8;  // a multiple of the SM count; grid-stride loops cover the rest
    if (blocks > cap) blocks = cap;
    if (blocks > kMaxRedBlocks) blocks = kMaxRedBlocks;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// cell kernel only; returns the number of blocks whose partials (energy + virial) now sit in h->d_partials
void prof_begin(mm_handle *h, int kind) {
    if (!h->profile) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    h->prof_events.emplace_back(a, b);
    h->prof_kinds.push_back(kind);
    cudaEventRecord(a, h->stream);
}

void prof_end(mm_handle *h) {
    if (!h->profile) return;
    cudaEventRecord(h->prof_events.back().second, h->stream);
}

int cells_launch(mm_handle *h) {
    const int gc = grid_for(h, h->ncells, kThreads);
    const bool batch = h->nreplicas > 1;
    const double *rv = batch ? h->d_rvecs_batch : h->d_rvecs;
    const int64_t cpr = batch ? h->ncells / h->nreplicas : h->ncells;
    prof_begin(h, 0);
#define MM_LAUNCH_CELLS(MODEL, BATCH)                                                                                  \
    k_cells<MODEL, BATCH><<<gc, kThreads, 0, h->stream>>>(h->kp, h->d_cell_nodes, h->d_cell_info, h->d_pos, rv, h->ncells, \
                                                          h->d_gcell, h->d_ecell, h->d_partials, cpr, h->d_vcell)
    if (h->model == MM_MODEL_ORIGINAL) {
        if (batch) MM_LAUNCH_CELLS(MM_MODEL_ORIGINAL, true);
        else MM_LAUNCH_CELLS(MM_MODEL_ORIGINAL, false);
    } else {
        if (batch) MM_LAUNCH_CELLS(MM_MODEL_DEFAULT, true);
        else MM_LAUNCH_CELLS(MM_MODEL_DEFAULT, false);
    }
#undef MM_LAUNCH_CELLS
    prof_end(h);
    h->launches++;
    if (batch) {
        k_replica_reduce<<<grid_for(h, h->nreplicas, 128), 128, 0, h->stream>>>(h->d_ecell, h->d_vcell, h->ncells, cpr,
                                                                                 h->nreplicas, h->d_rep);
        h->launches++;
    }
    return gc;
}

void final_launch(mm_handle *h, const double *pc, int nbc, const double *pn, int nbn) {
    k_final<<<1, 256, 0, h->stream>>>(pc, nbc, pn, nbn, h->d_result);
    h->launches++;
}

int force_evaluate(mm_handle *h, double *gpos_out, bool want_g2) {
    if (h->sg.active) {  // structured-grid kernels: AoS -> SoA planes, marching force kernel, SoA -> AoS
        int rc = sg_write_consts(h, h->rvecs, 0.0);
        if (rc != MM_OK) return rc;
        sg_pos_from_aos(h, h->d_pos);
        rc = sg_force(h, gpos_out != nullptr, 0);
        if (rc != MM_OK) return rc;
        if (gpos_out) sg_to_aos(h, 2, gpos_out);
        final_launch(h, h->sg.d_partials, h->sg.nblocks, h->sg.d_partials + 13, gpos_out ? h->sg.nblocks : 0);
        MM_CUDA(cudaGetLastError());
        return MM_OK;
    }
    {
        const int rc = ensure_generic(h);
        if (rc != MM_OK) return rc;
    }
    double *pn = h->d_partials + (size_t)kMaxRedBlocks * kRedSlots;
    if (h->scatter_mode && gpos_out) {  // cell-centric atomic scatter instead of per-cell gradients + node gather
        const int gcs = grid_for(h, h->ncells, kThreads);
        MM_CUDA(cudaMemsetAsync(gpos_out, 0, sizeof(double) * 3 * h->nnodes, h->stream));
        prof_begin(h, 0);
        if (h->model == MM_MODEL_ORIGINAL)
            k_cells_scatter<MM_MODEL_ORIGINAL><<<gcs, kThreads, 0, h->stream>>>(
                h->kp, h->d_cell_nodes, h->d_cell_info, h->d_pos, h->d_rvecs, h->ncells, gpos_out, h->d_ecell, h->d_partials);
        else
            k_cells_scatter<MM_MODEL_DEFAULT><<<gcs, kThreads, 0, h->stream>>>(
                h->kp, h->d_cell_nodes, h->d_cell_info, h->d_pos, h->d_rvecs, h->ncells, gpos_out, h->d_ecell, h->d_partials);
        prof_end(h);
        int gs = 0;
        if (want_g2) {
            gs = grid_for(h, 3 * h->nnodes, 256);
            k_sumsq<<<gs, 256, 0, h->stream>>>(gpos_out, 3 * h->nnodes, pn);
        }
        k_final<<<1, 256, 0, h->stream>>>(h->d_partials, gcs, pn, gs, h->d_result);
        h->launches += 3 + (want_g2 ? 1 : 0);
        MM_CUDA(cudaGetLastError());
        return MM_OK;
    }
    const int gc = cells_launch(h);
    int gn = 0;
    if (gpos_out) {
        gn = grid_for(h, h->nnodes, kThreads);
        if (want_g2)
            k_gather<true><<<gn, kThreads, 0, h->stream>>>(h->d_node_cells, h->d_gcell, h->nnodes, h->ncells, gpos_out, pn);
        else
            k_gather<false><<<gn, kThreads, 0, h->stream>>>(h->d_node_cells, h->d_gcell, h->nnodes, h->ncells, gpos_out, pn);
        h->launches++;
    }
    k_final<<<1, 256, 0, h->stream>>>(h->d_partials, gc, pn, (gpos_out && want_g2) ? gn : 0, h->d_result);
    h->launches++;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

}  // namespace mm
