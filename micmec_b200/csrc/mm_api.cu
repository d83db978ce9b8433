// mm_api.cu - C ABI of libmicmec_b200.so: handle lifetime, parameter folding, topology upload, compute().
// See include/micmec_b200.h for the reference interface each entry point replaces.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "mm_internal.h"

namespace mm {

static thread_local std::string g_error;

void set_error(const std::string &msg) { g_error = msg; }

int cuda_fail(cudaError_t err, const char *what) {
    g_error = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(err);
    cudaGetLastError();  // clear the sticky flag where possible
    return MM_ERR_CUDA;
}

static int invalid(const std::string &msg) {
    g_error = msg;
    return MM_ERR_INVALID;
}

static double det3(const double *a) {
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}

// Fold typeN/cell, typeN/elasticity, typeN/free_energy of one state into the kernel constants (mm_cell.cuh).
static void fold_state(const double *h0, const double *C, double efree, StateP &P) {
    const double d = det3(h0);
    P.v0 = d;
    P.efree = efree;
    P.hi[0] = (h0[4] * h0[8] - h0[5] * h0[7]) / d;
    P.hi[1] = (h0[2] * h0[7] - h0[1] * h0[8]) / d;
    P.hi[2] = (h0[1] * h0[5] - h0[2] * h0[4]) / d;
    P.hi[3] = (h0[5] * h0[6] - h0[3] * h0[8]) / d;
    P.hi[4] = (h0[0] * h0[8] - h0[2] * h0[6]) / d;
    P.hi[5] = (h0[2] * h0[3] - h0[0] * h0[5]) / d;
    P.hi[6] = (h0[3] * h0[7] - h0[4] * h0[6]) / d;
    P.hi[7] = (h0[1] * h0[6] - h0[0] * h0[7]) / d;
    P.hi[8] = (h0[0] * h0[4] - h0[1] * h0[3]) / d;
    static const int vi[6] = {0, 1, 2, 1, 0, 0}, vj[6] = {0, 1, 2, 2, 2, 1};
    auto c4 = [&](int i, int j, int k, int l) { return C[((i * 3 + j) * 3 + k) * 3 + l]; };
    for (int I = 0; I < 6; I++)
        for (int J = 0; J < 6; J++) {
            const int i = vi[I], j = vj[I], k = vi[J], l = vj[J];
            double a = 0.5 * (c4(i, j, k, l) + c4(j, i, k, l));
            if (k != l) a += 0.5 * (c4(i, j, l, k) + c4(j, i, l, k));
            P.A[I * 6 + J] = a;
        }
}

__global__ void k_grid_topology(int nx, int ny, int nz, int32_t *cell_nodes, int32_t *node_cells, uint8_t *cell_info,
                                int64_t n) {
    // full periodic grid, reference enumeration id = (k*ny + l)*nz + m  (micmec/utils.py:113-161, 205-216)
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(id % nz);
        const int l = (int)((id / nz) % ny);
        const int k = (int)(id / ((int64_t)nz * ny));
        unsigned info = cell_info[id] & 15u;
        if (k == nx - 1) info |= 16u;
        if (l == ny - 1) info |= 32u;
        if (m == nz - 1) info |= 64u;
        cell_info[id] = (uint8_t)info;
        for (int v = 0; v < 8; v++) {
            const int dx = vbit(v, 0), dy = vbit(v, 1), dz = vbit(v, 2);
            const int kn = (k + dx) % nx, ln = (l + dy) % ny, mn = (m + dz) % nz;
            cell_nodes[(int64_t)v * n + id] = (int32_t)(((int64_t)kn * ny + ln) * nz + mn);
            // surrounding_cells[node][v] = the cell for which the node is vertex v: offset -neighbor_nodes[v]
            const int kc = (k - dx + nx) % nx, lc = (l - dy + ny) % ny, mc = (m - dz + nz) % nz;
            node_cells[(int64_t)v * n + id] = (int32_t)(((int64_t)kc * ny + lc) * nz + mc);
        }
    }
}

// Index arrays and per-cell buffers of the indexed kernels, built on first use for structured grids (they cost
// 256 B per node, which the structured path never touches).
int ensure_generic(mm_handle *h) {
    if (h->d_cell_nodes) return MM_OK;
    if (h->slab_count > 1) {
        set_error("the indexed-topology kernels are not available on a z-slab of a decomposed grid");
        return MM_ERR_STATE;
    }
    const int64_t nc = h->ncells, nn = h->nnodes;
    MM_CUDA(cudaMalloc(&h->d_cell_nodes, sizeof(int32_t) * 8 * nc));
    MM_CUDA(cudaMalloc(&h->d_node_cells, sizeof(int32_t) * 8 * nn));
    MM_CUDA(cudaMalloc(&h->d_gcell, sizeof(double) * 24 * nc));
    MM_CUDA(cudaMalloc(&h->d_ecell, sizeof(double) * nc));
    k_grid_topology<<<grid_for(h, nc, 256), 256, 0, h->stream>>>(h->nx, h->ny, h->nz, h->d_cell_nodes, h->d_node_cells,
                                                                  h->d_cell_info, nc);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

}  // namespace mm

using namespace mm;

extern "C" {

const char *mm_last_error(void) { return g_error.c_str(); }

int mm_version(void) { return 100; }

int mm_device_ok(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        set_error("no usable CUDA device");
        return 0;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
    if (prop.major != 10) {
        set_error("libmicmec_b200 is built for sm_100a only; found compute capability " + std::to_string(prop.major) +
                  "." + std::to_string(prop.minor));
        return 0;
    }
    return 1;
}

int mm_destroy(mm_handle *h) {
    if (!h) return MM_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_cell_nodes);
    cudaFree(h->d_node_cells);
    cudaFree(h->d_cell_info);
    cudaFree(h->d_pos);
    cudaFree(h->d_gpos);
    cudaFree(h->d_gcell);
    cudaFree(h->d_ecell);
    cudaFree(h->d_partials);
    cudaFree(h->d_result);
    cudaFree(h->d_rvecs);
    cudaFree(h->d_red);
    cudaFree(h->d_halo);
    comm_peer_free(h);
    cudaFree(h->d_rvecs_batch);
    cudaFree(h->d_vcell);
    cudaFree(h->d_rep);
    mm_comm_destroy(h);
    if (h->sg.active || h->sg.d_sc) sg_free(h);
    if (h->h_result) cudaFreeHost(h->h_result);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MM_OK;
}

int mm_create(const mm_desc *desc, mm_handle **out) {
    if (!desc || !out) return invalid("mm_create: null argument");
    *out = nullptr;
    if (!mm_device_ok(desc->device)) return MM_ERR_CUDA;
    const bool structured = desc->nx > 0 && desc->ny > 0 && desc->nz > 0;
    if (desc->ncells <= 0 || desc->nnodes <= 0) return invalid("mm_create: empty system");
    if (desc->nnodes >= (int64_t)1 << 31 || desc->ncells >= (int64_t)1 << 31)
        return invalid("mm_create: more than 2^31 nodes or cells per device");
    if (desc->ntypes < 1 || desc->ntypes > MM_MAX_TYPES) return invalid("mm_create: unsupported number of cell types");
    if (desc->model != MM_MODEL_ORIGINAL && desc->model != MM_MODEL_DEFAULT) return invalid("mm_create: unknown model");
    if (structured) {
        if ((int64_t)desc->nx * desc->ny * desc->nz != desc->ncells || desc->ncells != desc->nnodes)
            return invalid("mm_create: structured grid needs ncells == nnodes == nx*ny*nz");
        if (desc->nx < 2 || desc->ny < 2 || desc->nz < 2)
            return invalid("mm_create: a periodic axis needs at least 2 cells (1-wide axes are degenerate in the reference)");
        if (desc->slab_count > 1 && (desc->slab_rank < 0 || desc->slab_rank >= desc->slab_count))
            return invalid("mm_create: slab_rank out of range");
        if (desc->slab_count > 1 && desc->model != MM_MODEL_ORIGINAL && (desc->ntypes != 1 || !desc->type_nstates || desc->type_nstates[0] != 1))
            return invalid("mm_create: the slab decomposition runs on the structured-grid kernels ('default' model: one cell type with one state)");
    } else if (desc->slab_count > 1) {
        return invalid("mm_create: the slab decomposition needs a structured grid (nx, ny, nz)");
    } else if (!desc->surrounding_nodes || !desc->surrounding_cells || !desc->shift) {
        return invalid("mm_create: index arrays are required for a non-structured system");
    }
    if (!desc->type_nstates || !desc->h0 || !desc->elasticity || !desc->free_energy || !desc->effective_temp)
        return invalid("mm_create: missing parameter arrays");

    mm_handle *h = new (std::nothrow) mm_handle();
    if (!h) return invalid("mm_create: out of host memory");
    h->device = desc->device;
    h->nnodes = desc->nnodes;
    h->ncells = desc->ncells;
    h->model = desc->model;
    h->boltzmann = desc->boltzmann;
    h->structured = structured ? 1 : 0;
    h->nx = desc->nx;
    h->ny = desc->ny;
    h->nz = desc->nz;
    h->slab_count = desc->slab_count > 1 ? desc->slab_count : 1;
    h->slab_rank = desc->slab_count > 1 ? desc->slab_rank : 0;
    h->nnodes_global = desc->nnodes_global > 0 ? desc->nnodes_global : desc->nnodes * h->slab_count;
    h->nreplicas = desc->nreplicas > 1 ? desc->nreplicas : 1;
    if (h->nreplicas > 1 && (structured || desc->nnodes % h->nreplicas != 0 || desc->ncells % h->nreplicas != 0)) {
        delete h;
        return invalid("mm_create: a replica batch needs indexed topology and equal-sized replicas");
    }

    // ---- parameters (mmff.py:219-231) -----------------------------------------------------------------------
    memset(&h->kp, 0, sizeof(h->kp));
    h->kp.ntypes = desc->ntypes;
    h->kp.model = desc->model;
    int total = 0;
    for (int t = 0; t < desc->ntypes; t++) {
        const int ns = desc->type_nstates[t];
        if (ns < 1 || total + ns > MM_MAX_STATES) {
            delete h;
            return invalid("mm_create: unsupported number of metastable states");
        }
        h->kp.nstates[t] = ns;
        h->kp.offset[t] = total;
        h->kp.kT[t] = desc->boltzmann * desc->effective_temp[t];
        for (int s = 0; s < ns; s++) {
            const int idx = total + s;
            if (det3(desc->h0 + 9 * idx) == 0.0) {
                delete h;
                return invalid("mm_create: singular equilibrium cell matrix");
            }
            fold_state(desc->h0 + 9 * idx, desc->elasticity + 81 * idx, desc->free_energy[idx], h->kp.st[idx]);
        }
        total += ns;
    }

    // ---- topology -------------------------------------------------------------------------------------------
    const int64_t nc = h->ncells, nn = h->nnodes;
    std::vector<uint8_t> info((size_t)nc, 0);
    for (int64_t c = 0; c < nc; c++) {
        const int t = desc->cell_type ? desc->cell_type[c] : 0;
        if (t < 0 || t >= desc->ntypes) {
            delete h;
            return invalid("mm_create: cell type index out of range");
        }
        info[(size_t)c] = (uint8_t)t;
    }
    std::vector<int32_t> cn, ncell;
    if (!structured) {
        cn.resize((size_t)nc * 8);
        ncell.resize((size_t)nn * 8);
        for (int64_t c = 0; c < nc; c++) {
            unsigned wrap = 0;
            for (int v = 0; v < 8; v++) {
                const int64_t n = desc->surrounding_nodes[c * 8 + v];
                if (n < 0 || n >= nn) {
                    delete h;
                    return invalid("mm_create: surrounding_nodes entry out of range");
                }
                cn[(size_t)v * nc + c] = (int32_t)n;
                for (int a = 0; a < 3; a++) {
                    const int s = desc->shift[(c * 8 + v) * 3 + a];
                    const int bit = (a == 0) ? (v == 1 || v == 4 || v == 5 || v == 7)
                                  : (a == 1) ? (v == 2 || v == 4 || v == 6 || v == 7)
                                             : (v == 3 || v == 5 || v == 6 || v == 7);
                    if (s != 0 && !(s == 1 && bit)) {
                        delete h;
                        return invalid("mm_create: minimum-image table is not of the periodic-grid form "
                                       "shift[c][v][a] = d_va * wrap_a(c)");
                    }
                    if (s == 1) wrap |= 1u << a;
                }
            }
            // every vertex with offset bit a must carry the flag once any does
            for (int v = 0; v < 8; v++)
                for (int a = 0; a < 3; a++) {
                    const int bit = (a == 0) ? (v == 1 || v == 4 || v == 5 || v == 7)
                                  : (a == 1) ? (v == 2 || v == 4 || v == 6 || v == 7)
                                             : (v == 3 || v == 5 || v == 6 || v == 7);
                    if (bit && ((wrap >> a) & 1u) != (unsigned)desc->shift[(c * 8 + v) * 3 + a]) {
                        delete h;
                        return invalid("mm_create: inconsistent minimum-image flags within a cell");
                    }
                }
            info[(size_t)c] |= (uint8_t)(wrap << 4);
        }
        for (int64_t n = 0; n < nn; n++)
            for (int v = 0; v < 8; v++) {
                const int64_t c = desc->surrounding_cells[n * 8 + v];
                if (c >= nc) {
                    delete h;
                    return invalid("mm_create: surrounding_cells entry out of range");
                }
                ncell[(size_t)v * nn + n] = c < 0 ? -1 : (int32_t)c;
            }
    }

#define MM_TRY(call)                                  \
    do {                                              \
        cudaError_t err__ = (call);                   \
        if (err__ != cudaSuccess) {                   \
            int rc__ = mm::cuda_fail(err__, #call);   \
            mm_destroy(h);                            \
            return rc__;                              \
        }                                             \
    } while (0)

    MM_TRY(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    MM_TRY(cudaGetDeviceProperties(&prop, h->device));
    h->num_sms = prop.multiProcessorCount;
    MM_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    if (!structured) {  // structured grids build these on first use of the indexed kernels (ensure_generic)
        MM_TRY(cudaMalloc(&h->d_cell_nodes, sizeof(int32_t) * 8 * nc));
        MM_TRY(cudaMalloc(&h->d_node_cells, sizeof(int32_t) * 8 * nn));
        MM_TRY(cudaMalloc(&h->d_gcell, sizeof(double) * 24 * nc));
        MM_TRY(cudaMalloc(&h->d_ecell, sizeof(double) * nc));
    }
    MM_TRY(cudaMalloc(&h->d_cell_info, nc));
    MM_TRY(cudaMalloc(&h->d_pos, sizeof(double) * 3 * nn));
    MM_TRY(cudaMalloc(&h->d_gpos, sizeof(double) * 3 * nn));
    MM_TRY(cudaMalloc(&h->d_partials, sizeof(double) * 2 * kMaxRedBlocks * kRedSlots));
    MM_TRY(cudaMalloc(&h->d_result, sizeof(ForceResult)));
    MM_TRY(cudaMalloc(&h->d_rvecs, sizeof(double) * 9));
    if (h->nreplicas > 1) {
        MM_TRY(cudaMalloc(&h->d_rvecs_batch, sizeof(double) * 9 * h->nreplicas));
        MM_TRY(cudaMalloc(&h->d_vcell, sizeof(double) * 6 * nc));
        MM_TRY(cudaMalloc(&h->d_rep, sizeof(double) * 8 * h->nreplicas));
        MM_TRY(cudaMemsetAsync(h->d_rvecs_batch, 0, sizeof(double) * 9 * h->nreplicas, h->stream));
    }
    MM_TRY(cudaHostAlloc(&h->h_result, sizeof(ForceResult), cudaHostAllocDefault));
    MM_TRY(cudaMemsetAsync(h->d_rvecs, 0, sizeof(double) * 9, h->stream));
    MM_TRY(cudaMemsetAsync(h->d_gpos, 0, sizeof(double) * 3 * nn, h->stream));
    MM_TRY(cudaMemcpyAsync(h->d_cell_info, info.data(), nc, cudaMemcpyHostToDevice, h->stream));
    if (!structured) {
        MM_TRY(cudaMemcpyAsync(h->d_cell_nodes, cn.data(), sizeof(int32_t) * 8 * nc, cudaMemcpyHostToDevice, h->stream));
        MM_TRY(cudaMemcpyAsync(h->d_node_cells, ncell.data(), sizeof(int32_t) * 8 * nn, cudaMemcpyHostToDevice, h->stream));
    }
    MM_TRY(cudaStreamSynchronize(h->stream));
#undef MM_TRY
    // small grids are launch-bound either way: keep them on the simple path (slabs always use the structured kernels)
    if (sg_eligible(h) && (h->nnodes >= 4096 || h->slab_count > 1)) {
        const int rc = sg_setup(h);
        if (rc != MM_OK) {
            mm_destroy(h);
            return rc;
        }
    }
    *out = h;
    return MM_OK;
}

int mm_set_stream(mm_handle *h, void *cuda_stream) {
    if (!h) return invalid("mm_set_stream: null handle");
    MM_CUDA(cudaSetDevice(h->device));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return MM_OK;
}

int mm_synchronize(mm_handle *h) {
    if (!h) return invalid("mm_synchronize: null handle");
    MM_CUDA(cudaSetDevice(h->device));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    return MM_OK;
}

int mm_set_option(mm_handle *h, const char *name, int64_t value) {
    if (!h || !name) return invalid("mm_set_option: null argument");
    if (strcmp(name, "scatter") == 0) {
        h->scatter_mode = value ? 1 : 0;
        return MM_OK;
    }
    if (strcmp(name, "structured") == 0) {
        // 1: use the structured-grid kernels (only full periodic grids with the `original` model qualify)
        h->want_structured = value ? 1 : 0;
        if (!value) {
            h->sg.active = 0;
            return MM_OK;
        }
        if (!sg_eligible(h)) return MM_OK;
        MM_CUDA(cudaSetDevice(h->device));
        if (!h->sg.d_sc) return sg_setup(h);
        h->sg.active = 1;
        return MM_OK;
    }
    if (strcmp(name, "variant") == 0) {
        h->sg.variant = (int)value & 15;
        return MM_OK;
    }
    if (strcmp(name, "chunk") == 0) {
        if (!h->sg.d_sc) return MM_OK;
        MM_CUDA(cudaSetDevice(h->device));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        if (sg_set_chunk(h, (int)value) != MM_OK) return invalid("mm_set_option: chunk must be positive");
        return MM_OK;
    }
    if (strcmp(name, "tile_rows") == 0) {
        if (!h->sg.d_sc) return MM_OK;
        MM_CUDA(cudaSetDevice(h->device));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        if (sg_set_tile_rows(h, (int)value) != MM_OK) return invalid("mm_set_option: tile_rows not compiled into this library");
        return MM_OK;
    }
    if (strcmp(name, "rows_per_thread") == 0 || strcmp(name, "march2") == 0) {
        if (!h->sg.d_sc) return MM_OK;
        MM_CUDA(cudaSetDevice(h->device));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        const int rc = name[0] == 'r' ? sg_set_rpt(h, (int)value) : sg_set_march2(h, (int)value);
        if (rc != MM_OK) return invalid("mm_set_option: rows_per_thread / march2 configuration not compiled into this library");
        return MM_OK;
    }
    if (strcmp(name, "profile") == 0) {
        h->profile = value ? 1 : 0;
        return MM_OK;
    }
    if (strcmp(name, "tail") == 0) {
        h->sg.tail_wanted = value ? 1 : 0;
        return MM_OK;
    }
    if (strcmp(name, "wrap_on_load") == 0) {
        h->sg.wrap_wanted = value < 0 ? -1 : (value ? 1 : 0);
        if (!h->sg.d_sc) return MM_OK;
        MM_CUDA(cudaSetDevice(h->device));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        return sg_retile(h, 0);
    }
    if (strcmp(name, "plan") == 0) {
        h->sg.plan_two_class = value ? 1 : 0;
        if (!h->sg.d_sc) return MM_OK;
        MM_CUDA(cudaSetDevice(h->device));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        return sg_retile(h, 0);
    }
    if (strcmp(name, "tail_in_kernel") == 0) {
        h->sg.tail_in_kernel = value ? 1 : 0;
        return MM_OK;
    }
    return invalid(std::string("mm_set_option: unknown option ") + name);
}

int64_t mm_get_option(const mm_handle *h, const char *name) {
    if (!h || !name) return -1;
    if (strcmp(name, "structured") == 0) return h->sg.active ? 1 : 0;
    if (strcmp(name, "chunk") == 0) return h->sg.chunk;
    if (strcmp(name, "tile_rows") == 0) return h->sg.tile_rows;
    if (strcmp(name, "blocks") == 0) return h->sg.nblocks;
    if (strcmp(name, "variant") == 0) return h->sg.variant;
    if (strcmp(name, "rows_per_thread") == 0) return h->sg.march2 ? h->sg.rpt : 1;
    if (strcmp(name, "march2") == 0) return h->sg.march2;
    if (strcmp(name, "mass_uniform") == 0) return h->sg.mass_uniform;
    if (strcmp(name, "plan_efficiency_permille") == 0)  // perfect balance / simulated makespan of the block schedule
        return h->sg.plan_cost > 0.0 ? (int64_t)(1000.0 * h->sg.plan_ideal / h->sg.plan_cost) : -1;
    if (strcmp(name, "tail") == 0) return sg_tail_ok(h) ? (h->sg.tail_in_kernel ? 2 : 1) : 0;
    if (strcmp(name, "wrap_on_load") == 0) return h->sg.wrap_on_load;
    return -1;
}

int mm_plan_schedule(int ntx, int nty, int planes, int nsm, int images_on_load, int uniform_chunk, int32_t *items, int64_t capacity,
                     int64_t *nitems, double *cost, double *ideal) {
    if (ntx < 1 || nty < 1 || planes < 1 || nsm < 1 || !nitems) return invalid("mm_plan_schedule: bad argument");
    std::vector<int4> v;
    sg_plan_schedule(ntx, nty, planes, nsm, images_on_load, uniform_chunk, v, cost, ideal, nullptr);
    *nitems = (int64_t)v.size();
    if (items) {
        if (capacity < (int64_t)v.size()) return invalid("mm_plan_schedule: items buffer too small");
        for (size_t i = 0; i < v.size(); i++) {
            items[4 * i + 0] = v[i].x;
            items[4 * i + 1] = v[i].y;
            items[4 * i + 2] = v[i].z;
            items[4 * i + 3] = v[i].w;
        }
    }
    return MM_OK;
}

int mm_profile(mm_handle *h, int64_t *nlaunch, double *total_ms) {
    if (!h || !nlaunch || !total_ms) return invalid("mm_profile: null argument");
    MM_CUDA(cudaSetDevice(h->device));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    nlaunch[0] = nlaunch[1] = 0;
    total_ms[0] = total_ms[1] = 0.0;
    for (size_t i = 0; i < h->prof_events.size(); i++) {
        auto &ev = h->prof_events[i];
        float ms = 0.0f;
        MM_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        const int kind = h->prof_kinds[i] ? 1 : 0;
        nlaunch[kind]++;
        total_ms[kind] += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    h->prof_events.clear();
    h->prof_kinds.clear();
    return MM_OK;
}

int64_t mm_launch_count(const mm_handle *h) { return h ? h->launches : 0; }

void *mm_device_ptr(mm_handle *h, int which) {
    if (!h) return nullptr;
    return which == 0 ? (void *)h->d_pos : which == 1 ? (void *)h->d_gpos : nullptr;
}

int mm_set_pos(mm_handle *h, const double *pos, int where) {
    if (!h || !pos) return invalid("mm_set_pos: null argument");
    MM_CUDA(cudaSetDevice(h->device));
    if (pos != h->d_pos)
        MM_CUDA(cudaMemcpyAsync(h->d_pos, pos, sizeof(double) * 3 * h->nnodes,
                                where == MM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    h->pos_valid = true;
    return MM_OK;
}

int mm_set_rvecs(mm_handle *h, const double *rvecs9) {
    if (!h || !rvecs9) return invalid("mm_set_rvecs: null argument");
    MM_CUDA(cudaSetDevice(h->device));
    // the H2D copy below is asynchronous: keep the source alive in the handle
    MM_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(h->rvecs, rvecs9, sizeof(double) * 9);
    MM_CUDA(cudaMemcpyAsync(h->d_rvecs, h->rvecs, sizeof(double) * 9, cudaMemcpyHostToDevice, h->stream));
    return MM_OK;
}

int mm_compute(mm_handle *h, double *energy_host, double *gpos, int where, double *vtens9) {
    if (!h || !energy_host) return invalid("mm_compute: null argument");
    if (!h->pos_valid) {
        set_error("mm_compute: positions were never set (call mm_set_pos first)");
        return MM_ERR_STATE;
    }
    MM_CUDA(cudaSetDevice(h->device));
    int rc = force_evaluate(h, gpos ? h->d_gpos : nullptr, gpos != nullptr);
    if (rc != MM_OK) return rc;
    if (h->slab_count > 1) {  // energy, virial and sum g^2 of the whole grid
        rc = comm_allreduce(h, reinterpret_cast<double *>(h->d_result), 8);
        if (rc != MM_OK) return rc;
    }
    MM_CUDA(cudaMemcpyAsync(h->h_result, h->d_result, sizeof(ForceResult), cudaMemcpyDeviceToHost, h->stream));
    if (gpos && gpos != h->d_gpos)
        MM_CUDA(cudaMemcpyAsync(gpos, h->d_gpos, sizeof(double) * 3 * h->nnodes,
                                where == MM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    const ForceResult &r = *h->h_result;
    *energy_host = r.epot;
    if (vtens9) {
        vtens9[0] = r.vir[0];
        vtens9[4] = r.vir[1];
        vtens9[8] = r.vir[2];
        vtens9[5] = vtens9[7] = r.vir[3];
        vtens9[2] = vtens9[6] = r.vir[4];
        vtens9[1] = vtens9[3] = r.vir[5];
    }
    // mmff.py:135-147
    if (std::isnan(r.epot)) {
        set_error("The energy is not-a-number (``nan``).");
        return MM_ERR_NAN;
    }
    if (gpos && std::isnan(r.sum_g2)) {
        set_error("Some ``gpos`` element(s) is/are not-a-number (``nan``).");
        return MM_ERR_NAN;
    }
    if (vtens9)
        for (int k = 0; k < 6; k++)
            if (std::isnan(r.vir[k])) {
                set_error("Some ``vtens`` element(s) is/are not-a-number (``nan``).");
                return MM_ERR_NAN;
            }
    return MM_OK;
}

int mm_set_rvecs_batch(mm_handle *h, const double *rvecs) {
    if (!h || !rvecs) return invalid("mm_set_rvecs_batch: null argument");
    if (h->nreplicas <= 1) return invalid("mm_set_rvecs_batch: the handle is not a replica batch");
    MM_CUDA(cudaSetDevice(h->device));
    MM_CUDA(cudaMemcpyAsync(h->d_rvecs_batch, rvecs, sizeof(double) * 9 * h->nreplicas, cudaMemcpyHostToDevice, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));  // the source is the caller's pageable array
    return MM_OK;
}

int mm_get_replica_results(mm_handle *h, double *energies, double *vtens) {
    if (!h || !energies) return invalid("mm_get_replica_results: null argument");
    if (h->nreplicas <= 1) return invalid("mm_get_replica_results: the handle is not a replica batch");
    MM_CUDA(cudaSetDevice(h->device));
    std::vector<double> tmp((size_t)h->nreplicas * 8);
    MM_CUDA(cudaMemcpyAsync(tmp.data(), h->d_rep, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    for (int64_t r = 0; r < h->nreplicas; r++) {
        const double *s = tmp.data() + r * 8;
        energies[r] = s[0];
        if (vtens) {
            double *v = vtens + r * 9;
            v[0] = s[1]; v[4] = s[2]; v[8] = s[3];
            v[5] = v[7] = s[4]; v[2] = v[6] = s[5]; v[1] = v[3] = s[6];
        }
    }
    return MM_OK;
}

int mm_get_cell_cache(mm_handle *h, double *epot_cells, double *gpos_cells) {
    if (!h) return invalid("mm_get_cell_cache: null handle");
    MM_CUDA(cudaSetDevice(h->device));
    const int64_t nc = h->ncells;
    if (h->sg.active) {  // the structured kernels never materialise per-cell data: produce it on demand
        const int rc = ensure_generic(h);
        if (rc != MM_OK) return rc;
        cells_launch(h);
    }
    if (epot_cells)
        MM_CUDA(cudaMemcpyAsync(epot_cells, h->d_ecell, sizeof(double) * nc, cudaMemcpyDeviceToHost, h->stream));
    if (gpos_cells) {
        std::vector<double> tmp((size_t)nc * 24);
        MM_CUDA(cudaMemcpyAsync(tmp.data(), h->d_gcell, sizeof(double) * 24 * nc, cudaMemcpyDeviceToHost, h->stream));
        MM_CUDA(cudaStreamSynchronize(h->stream));
        for (int64_t c = 0; c < nc; c++)
            for (int k = 0; k < 24; k++) gpos_cells[c * 24 + k] = tmp[(size_t)k * nc + c];
    }
    MM_CUDA(cudaStreamSynchronize(h->stream));
    return MM_OK;
}

// ---- Domain (micmec/pes/ext.pyx:49-71 + micmec/pes/domain.c:13-50) ------------------------------------------
int mm_domain(const double *r, int nvec, double *volume, double *gvecs) {
    if (nvec < 0 || nvec > 3 || (nvec > 0 && !r)) return invalid("mm_domain: rvecs must have at most three rows");
    double vol = 0.0;
    if (nvec == 1) {
        const double n2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        vol = std::sqrt(n2);
        if (gvecs)
            for (int i = 0; i < 3; i++) gvecs[i] = r[i] / n2;
    } else if (nvec == 2) {
        const double aa = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        const double bb = r[3] * r[3] + r[4] * r[4] + r[5] * r[5];
        const double ab = r[0] * r[3] + r[1] * r[4] + r[2] * r[5];
        const double det = aa * bb - ab * ab;
        vol = det > 0 ? std::sqrt(det) : 0.0;
        if (gvecs)
            for (int i = 0; i < 3; i++) {  // rows of the pseudo-inverse transpose: (R R^T)^-1 R
                gvecs[i] = (bb * r[i] - ab * r[3 + i]) / det;
                gvecs[3 + i] = (aa * r[3 + i] - ab * r[i]) / det;
            }
    } else if (nvec == 3) {
        const double d = r[0] * (r[4] * r[8] - r[5] * r[7]) + r[1] * (r[5] * r[6] - r[3] * r[8]) +
                         r[2] * (r[3] * r[7] - r[4] * r[6]);
        vol = std::fabs(d);
        if (gvecs) {  // inverse transpose: g_i . r_j = delta_ij
            gvecs[0] = (r[4] * r[8] - r[5] * r[7]) / d;
            gvecs[1] = (r[5] * r[6] - r[3] * r[8]) / d;
            gvecs[2] = (r[3] * r[7] - r[4] * r[6]) / d;
            gvecs[3] = (r[2] * r[7] - r[1] * r[8]) / d;
            gvecs[4] = (r[0] * r[8] - r[2] * r[6]) / d;
            gvecs[5] = (r[1] * r[6] - r[0] * r[7]) / d;
            gvecs[6] = (r[1] * r[5] - r[2] * r[4]) / d;
            gvecs[7] = (r[2] * r[3] - r[0] * r[5]) / d;
            gvecs[8] = (r[0] * r[4] - r[1] * r[3]) / d;
        }
    }
    if (volume) *volume = vol;
    return MM_OK;
}

}  // extern "C"
