// mm_comm.cu - z-slab decomposition across the GPUs of one box: NCCL plumbing.
//
// The reference has no parallelism.  The path shards as a nearest-neighbour stencil: a rank owns a
// contiguous range of z planes, needs ONE node plane from each z neighbour per force evaluation (positions; for the
// fused step also velocities and old gradients) and one all-reduce of <= 16 doubles (energy, virial, kinetic
// moments).  Thermostat / barostat algebra is replicated: every rank runs the same scalar kernel on bit-identical
// reduced inputs.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already has - the one PyTorch loaded), so the
// library neither links against NCCL nor needs it on a single GPU.
#include <dlfcn.h>

#include <cstring>

#include "mm_internal.h"
#include "mm_reduce.cuh"

namespace mm {

// minimal NCCL surface (matches nccl.h 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load(const char *path) {
    if (g_nccl.lib) return MM_OK;
    const char *names[] = {path, "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        set_error(std::string("cannot load NCCL: ") + dlerror());
        return MM_ERR_CUDA;
    }
#define MM_SYM(field, name)                                         \
    *(void **)(&g_nccl.field) = dlsym(lib, name);                   \
    if (!g_nccl.field) {                                            \
        set_error(std::string("NCCL symbol missing: ") + name);     \
        return MM_ERR_CUDA;                                         \
    }
    MM_SYM(GetUniqueId, "ncclGetUniqueId")
    MM_SYM(CommInitRank, "ncclCommInitRank")
    MM_SYM(CommDestroy, "ncclCommDestroy")
    MM_SYM(Send, "ncclSend")
    MM_SYM(Recv, "ncclRecv")
    MM_SYM(AllReduce, "ncclAllReduce")
    MM_SYM(GroupStart, "ncclGroupStart")
    MM_SYM(GroupEnd, "ncclGroupEnd")
    MM_SYM(GetErrorString, "ncclGetErrorString")
#undef MM_SYM
    g_nccl.lib = lib;
    return MM_OK;
}

static int nccl_fail(int rc, const char *what) {
    set_error(std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    return MM_ERR_CUDA;
}

#define MM_NCCL(call)                                        \
    do {                                                     \
        int rc__ = (call);                                   \
        if (rc__ != ncclSuccess) return nccl_fail(rc__, #call); \
    } while (0)

// ---- halo planes over NCCL ----------------------------------------------------------------------------------------
// up = rank + 1 (receives my top owned plane as its lower halo), down = rank - 1 (receives my bottom owned plane as
// its upper halo); periodic in the rank index.  The boundary planes of all fields of one exchange are packed into
// ONE message per direction (a group of 2 sends + 2 receives instead of 4 per field).
struct PackArgs {
    double *f[9];
    int nfields;
    int npos;  // the first npos fields are position components
};

__global__ void __launch_bounds__(256)
k_pack(const __grid_constant__ PackArgs a, int64_t plane, int nzl, double *up, double *down) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x)
        for (int f = 0; f < a.nfields; f++) {
            up[(int64_t)f * plane + i] = a.f[f][(int64_t)nzl * plane + i];  // top owned plane
            down[(int64_t)f * plane + i] = a.f[f][plane + i];               // bottom owned plane
        }
}

// lower halo <- message from below (minus c across the periodic wrap), upper halo <- message from above (plus c)
__global__ void __launch_bounds__(256)
k_unpack(const __grid_constant__ PackArgs a, int64_t plane, int nzl, const double *from_down, const double *from_up,
         const StepConsts *sc, double sign_lo, double sign_hi) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x)
        for (int f = 0; f < a.nfields; f++) {
            const double c = (f < a.npos) ? sc->rv[6 + f] : 0.0;
            a.f[f][i] = from_down[(int64_t)f * plane + i] + sign_lo * c;
            a.f[f][(int64_t)(nzl + 1) * plane + i] = from_up[(int64_t)f * plane + i] + sign_hi * c;
        }
}

int comm_halo(mm_handle *h, double **fields, int nfields, int npos) {
    SGrid &g = h->sg;
    if (!h->comm) {
        set_error("this handle is a z-slab of a decomposed grid: call mm_comm_init first");
        return MM_ERR_STATE;
    }
    const int P = h->slab_count, r = h->slab_rank;
    const int up = (r + 1) % P, down = (r + P - 1) % P;
    ncclComm_t comm = (ncclComm_t)h->comm;
    const size_t n = (size_t)g.plane;
    if (!h->d_halo) MM_CUDA(cudaMalloc(&h->d_halo, sizeof(double) * 4 * 9 * n));
    double *s_up = h->d_halo, *s_down = s_up + 9 * n, *r_down = s_down + 9 * n, *r_up = r_down + 9 * n;
    PackArgs a;
    a.nfields = nfields;
    a.npos = npos;
    for (int f = 0; f < nfields; f++) a.f[f] = fields[f];
    const int grid = grid_for(h, g.plane, 256);
    k_pack<<<grid, 256, 0, h->stream>>>(a, g.plane, g.nzl, s_up, s_down);
    MM_NCCL(g_nccl.GroupStart());
    MM_NCCL(g_nccl.Send(s_up, n * nfields, ncclFloat64, up, comm, h->stream));
    MM_NCCL(g_nccl.Recv(r_down, n * nfields, ncclFloat64, down, comm, h->stream));
    MM_NCCL(g_nccl.Send(s_down, n * nfields, ncclFloat64, down, comm, h->stream));
    MM_NCCL(g_nccl.Recv(r_up, n * nfields, ncclFloat64, up, comm, h->stream));
    MM_NCCL(g_nccl.GroupEnd());
    // periodic wrap of the rank ring: -c below rank 0, +c above rank P-1 (c of the stored frame)
    k_unpack<<<grid, 256, 0, h->stream>>>(a, g.plane, g.nzl, r_down, r_up, g.d_sc, r == 0 ? -1.0 : 0.0, r == P - 1 ? 1.0 : 0.0);
    h->launches += 3;
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}

int comm_halo_u8(mm_handle *h, uint8_t *field) {
    SGrid &g = h->sg;
    const int P = h->slab_count, r = h->slab_rank;
    const int up = (r + 1) % P, down = (r + P - 1) % P;
    ncclComm_t comm = (ncclComm_t)h->comm;
    const size_t n = (size_t)g.plane;
    MM_NCCL(g_nccl.GroupStart());
    MM_NCCL(g_nccl.Send(field + (size_t)g.nzl * n, n, ncclUint8, up, comm, h->stream));
    MM_NCCL(g_nccl.Recv(field, n, ncclUint8, down, comm, h->stream));
    MM_NCCL(g_nccl.Send(field + n, n, ncclUint8, down, comm, h->stream));
    MM_NCCL(g_nccl.Recv(field + (size_t)(g.nzl + 1) * n, n, ncclUint8, up, comm, h->stream));
    MM_NCCL(g_nccl.GroupEnd());
    return MM_OK;
}

int comm_allreduce(mm_handle *h, double *buf, int count) {
    MM_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream));
    h->launches++;
    return MM_OK;
}

// out[0..6] = sum pc[.][0..6], out[7..13] = sum pn[.][0..6], out[14] = sum pd[.][0]   (one block, fixed order)
__global__ void __launch_bounds__(256)
k_sum_partials(const double *pc, int nbc, const double *pn, int nbn, const double *pd, int nbd, double *out) {
    double a[7] = {0, 0, 0, 0, 0, 0, 0}, b[7] = {0, 0, 0, 0, 0, 0, 0}, c[1] = {0};
    if (nbc > 0) partials_sum<7>(pc, nbc, kRedSlots, a);
    if (nbn > 0) partials_sum<7>(pn, nbn, kRedSlots, b);
    if (nbd > 0) partials_sum<1>(pd, nbd, kRedSlots, c);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 7; k++) {
            out[k] = a[k];
            out[7 + k] = b[k];
        }
        out[14] = c[0];
        out[15] = 0.0;
    }
}

int comm_reduce_partials(mm_handle *h, const double *pc, int nbc, const double *pn, int nbn, const double *pd, int nbd) {
    k_sum_partials<<<1, 256, 0, h->stream>>>(pc, nbc, pn, nbn, pd, nbd, h->d_red);
    h->launches++;
    return comm_allreduce(h, h->d_red, 16);
}

}  // namespace mm

using namespace mm;

extern "C" {

int mm_comm_unique_id(const char *nccl_path, char *out128) {
    if (!out128) {
        set_error("mm_comm_unique_id: null argument");
        return MM_ERR_INVALID;
    }
    int rc = nccl_load(nccl_path);
    if (rc != MM_OK) return rc;
    ncclUniqueId id;
    MM_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return MM_OK;
}

int mm_comm_init(mm_handle *h, const char *nccl_path, const char *id128) {
    if (!h || !id128) {
        set_error("mm_comm_init: null argument");
        return MM_ERR_INVALID;
    }
    if (h->slab_count <= 1) return MM_OK;
    int rc = nccl_load(nccl_path);
    if (rc != MM_OK) return rc;
    MM_CUDA(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t comm = nullptr;
    MM_NCCL(g_nccl.CommInitRank(&comm, h->slab_count, id, h->slab_rank));
    h->comm = comm;
    if (!h->d_red) MM_CUDA(cudaMalloc(&h->d_red, sizeof(double) * 32));
    // static per-cell data of the neighbours' boundary planes
    rc = comm_halo_u8(h, h->sg.type);
    if (rc != MM_OK) return rc;
    MM_CUDA(cudaStreamSynchronize(h->stream));
    return MM_OK;
}

int mm_comm_destroy(mm_handle *h) {
    if (h && h->comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
    }
    return MM_OK;
}

}  // extern "C"
