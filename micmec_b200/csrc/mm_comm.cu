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

#include <cstdlib>
#include <cstring>
#include <vector>

#include "mm_internal.h"
#include "mm_peer.cuh"
#include "mm_reduce.cuh"

namespace mm {

// minimal NCCL surface (matches nccl.h 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint8 = 1, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMin = 3 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    MM_SYM(AllGather, "ncclAllGather")
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


// ---- peer mode: direct stores into the neighbours' memory -----------------------------------------------------------
// One peer-visible block per rank (cudaMalloc, exported with CUDA IPC, mapped by every other rank of the box):
//   flags    [0] blocks that have delivered "from below", [1] "from above" (monotonic counters), [2 + r] reduction epoch of rank r
//   mailbox  [parity][rank][16]   partial sums of every rank (each rank writes its row on EVERY rank: one-shot all-gather)
//   inbox    [parity][from below / from above][9 fields][plane]
// A halo exchange is two launches with no host or NCCL involvement: k_peer_push stores the two boundary planes of every
// field straight into the neighbours' inboxes over NVLink, fences at system scope and counts its blocks on the
// neighbours' flags; k_peer_unpack waits for both counters and copies the inbox into the halo planes (adding the
// periodic shift at the ends of the rank ring).  The all-reduce is ONE launch: sum the local partials, store the 16
// doubles into all mailboxes, release-store the epoch, wait for all ranks' epochs and add the rows in rank order - every
// rank gets bit-identical sums.  Epochs live in device memory (PeerCtl), so captured CUDA graphs replay correctly;
// inboxes / mailboxes alternate by epoch parity: a rank can be at most one exchange ahead of a neighbour.
__global__ void __launch_bounds__(256)
k_peer_push(const __grid_constant__ PackArgs a, int64_t plane, int nzl, char *up_base, char *down_base, size_t inbox_off,
            const PeerCtl *ctl) {
    const unsigned long long E = ctl->halo_epoch + 1;
    const size_t par = (size_t)(E & 1ull) * 2 * 9 * plane;
    double *to_up = reinterpret_cast<double *>(up_base + inbox_off) + par;                 // its "from below" slot
    double *to_down = reinterpret_cast<double *>(down_base + inbox_off) + par + 9 * plane;  // its "from above" slot
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x)
        for (int f = 0; f < a.nfields; f++) {
            to_up[(int64_t)f * plane + i] = a.f[f][(int64_t)nzl * plane + i];  // top owned plane
            to_down[(int64_t)f * plane + i] = a.f[f][plane + i];               // bottom owned plane
        }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd_system(reinterpret_cast<unsigned long long *>(up_base), 1ull);
        atomicAdd_system(reinterpret_cast<unsigned long long *>(down_base) + 1, 1ull);
    }
}

__global__ void __launch_bounds__(256)
k_peer_unpack(const __grid_constant__ PackArgs a, int64_t plane, int nzl, char *own_base, size_t inbox_off, PeerCtl *ctl,
              unsigned long long push_blocks, const StepConsts *sc, double sign_lo, double sign_hi) {
    const unsigned long long E = ctl->halo_epoch + 1;
    if (threadIdx.x == 0) {
        const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(own_base);
        while (ld_acquire_sys(flags) < E * push_blocks) {
        }
        while (ld_acquire_sys(flags + 1) < E * push_blocks) {
        }
    }
    __syncthreads();
    const double *from_down = reinterpret_cast<const double *>(own_base + inbox_off) + (size_t)(E & 1ull) * 2 * 9 * plane;
    const double *from_up = from_down + 9 * plane;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x)
        for (int f = 0; f < a.nfields; f++) {
            const double c = (f < a.npos) ? sc->rv[6 + f] : 0.0;
            // the inbox was written by another GPU: read it past the (non-coherent) L1
            a.f[f][i] = __ldcg(from_down + (int64_t)f * plane + i) + sign_lo * c;
            a.f[f][(int64_t)(nzl + 1) * plane + i] = __ldcg(from_up + (int64_t)f * plane + i) + sign_hi * c;
        }
    __syncthreads();
    if (threadIdx.x == 0) {  // the last block to finish publishes the new epoch (every block has read the old one)
        __threadfence();
        if (atomicAdd(&ctl->halo_done, 1u) == gridDim.x - 1) {
            ctl->halo_done = 0;
            ctl->halo_epoch = E;
            __threadfence();
        }
    }
}

// sum of the local partials (as k_sum_partials) + all-gather through the mailboxes + rank-ordered sum -> out[16]
__global__ void __launch_bounds__(256)
k_peer_allreduce(const double *pc, int nbc, const double *pn, int nbn, const double *pd, int nbd, char *const *bases, int P,
                 int rank, PeerCtl *ctl, double *out) {
    __shared__ double mine[16];
    double a[7] = {0, 0, 0, 0, 0, 0, 0}, b[7] = {0, 0, 0, 0, 0, 0, 0}, c[1] = {0};
    if (pn == pc + 7 && nbn == nbc && nbc > 1) {  // structured path: one coalesced pass over the 16-slot records
        double all[16];
        partials_sum16(pc, nbc, all);
        for (int k = 0; k < 7; k++) {
            a[k] = all[k];
            b[k] = all[7 + k];
        }
    } else {
        if (nbc > 0) partials_sum<7>(pc, nbc, kRedSlots, a);
        if (nbn > 0) partials_sum<7>(pn, nbn, kRedSlots, b);
    }
    if (nbd > 0) partials_sum<1>(pd, nbd, kRedSlots, c);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 7; k++) {
            mine[k] = a[k];
            mine[7 + k] = b[k];
        }
        mine[14] = c[0];
        mine[15] = 0.0;
    }
    __syncthreads();
    const unsigned long long E = ctl->red_epoch + 1;
    const size_t par = (size_t)(E & 1ull) * P * 16;
    const int t = threadIdx.x;
    if (t < 16 * P) {  // my row on every rank (my own included)
        double *mail = reinterpret_cast<double *>(bases[t / 16] + kPeerFlagBytes);
        mail[par + (size_t)rank * 16 + (t % 16)] = mine[t % 16];
    }
    __threadfence_system();
    __syncthreads();
    if (t < P) st_release_sys(reinterpret_cast<unsigned long long *>(bases[t]) + 2 + rank, E);
    if (t < P) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(bases[rank]) + 2 + t;
        while (ld_acquire_sys(flag) < E) {
        }
    }
    __syncthreads();
    if (t < 16) {
        const double *mail = reinterpret_cast<const double *>(bases[rank] + kPeerFlagBytes) + par;
        double s = 0.0;
        for (int r = 0; r < P; r++) s += __ldcg(mail + (size_t)r * 16 + t);
        out[t] = s;
    }
    if (t == 0) ctl->red_epoch = E;
}

void comm_peer_free(mm_handle *h) {
    for (int s = 0; s < 2; s++) {
        if (h->peer_soa[s] && (s == 0 || h->peer_soa[1] != h->peer_soa[0])) cudaIpcCloseMemHandle(h->peer_soa[s]);
    }
    h->peer_soa[0] = h->peer_soa[1] = nullptr;
    if (h->slab_count > 1) h->sg.fused = 0;
    for (int r = 0; r < 16; r++)
        if (h->peer_base[r] && h->peer_base[r] != h->d_peer) cudaIpcCloseMemHandle(h->peer_base[r]);
    for (int r = 0; r < 16; r++) h->peer_base[r] = nullptr;
    cudaFree(h->d_peer);
    cudaFree(h->d_peer_base);
    cudaFree(h->d_peer_ctl);
    h->d_peer = nullptr;
    h->d_peer_base = nullptr;
    h->d_peer_ctl = nullptr;
    h->peer_mode = 0;
}

// Fused halo (mm_structured.cuh: MarchArgs): map the node-array blocks of the two z neighbours, so that the marching
// kernel can store its boundary planes straight into their halo planes.  Collective; all ranks agree on the outcome.
struct SoaInfo {
    cudaIpcMemHandle_t hdl;
    int64_t stride;
    int32_t nzl, pad;
};

static int peer_setup_fused(mm_handle *h) {
    const int P = h->slab_count, r = h->slab_rank;
    const int up = (r + 1) % P, down = (r + P - 1) % P;
    ncclComm_t comm = (ncclComm_t)h->comm;
    SGrid &g = h->sg;
    const char *env = getenv("MICMEC_B200_FUSED");
    int ok = (env && atoi(env) == 0) ? 0 : 1;
    SoaInfo mine;
    memset(&mine, 0, sizeof(mine));
    if (ok && cudaIpcGetMemHandle(&mine.hdl, g.block) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    mine.stride = g.stride;
    mine.nzl = g.nzl;
    unsigned char *d_info = nullptr;
    int *d_ok = nullptr;
    MM_CUDA(cudaMalloc(&d_info, sizeof(SoaInfo) * (size_t)P));
    MM_CUDA(cudaMalloc(&d_ok, sizeof(int)));
    MM_CUDA(cudaMemcpyAsync(d_info + sizeof(SoaInfo) * (size_t)r, &mine, sizeof(SoaInfo), cudaMemcpyHostToDevice, h->stream));
    MM_NCCL(g_nccl.AllGather(d_info + sizeof(SoaInfo) * (size_t)r, d_info, sizeof(SoaInfo), ncclUint8, comm, h->stream));
    std::vector<SoaInfo> all(P);
    MM_CUDA(cudaMemcpyAsync(all.data(), d_info, sizeof(SoaInfo) * (size_t)P, cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    if (ok) {
        void *pd = nullptr, *pu = nullptr;
        if (cudaIpcOpenMemHandle(&pd, all[down].hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        } else if (up == down) {
            pu = pd;
        } else if (cudaIpcOpenMemHandle(&pu, all[up].hdl, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            cudaIpcCloseMemHandle(pd);
            pd = nullptr;
            ok = 0;
        }
        h->peer_soa[0] = pd;
        h->peer_soa[1] = pu;
    }
    MM_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    MM_NCCL(g_nccl.AllReduce(d_ok, d_ok, 1, 2 /* ncclInt32 */, ncclMin, comm, h->stream));
    int all_ok = 0;
    MM_CUDA(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_info);
    cudaFree(d_ok);
    if (!all_ok) {
        for (int s = 0; s < 2; s++)
            if (h->peer_soa[s] && (s == 0 || h->peer_soa[1] != h->peer_soa[0])) cudaIpcCloseMemHandle(h->peer_soa[s]);
        h->peer_soa[0] = h->peer_soa[1] = nullptr;
        return MM_OK;  // peer mode 1 (inbox copies) stays on
    }
    const int nb[2] = {down, up};
    for (int s = 0; s < 2; s++) {
        g.nb_block[s] = (double *)h->peer_soa[s];
        g.nb_stride[s] = all[nb[s]].stride;
        g.nb_nzl[s] = all[nb[s]].nzl;
    }
    // my lower boundary plane lands in the upper halo of the rank below: its "from above" counter; and vice versa
    g.nb_flag[0] = reinterpret_cast<unsigned long long *>(h->peer_base[down]) + 33;
    g.nb_flag[1] = reinterpret_cast<unsigned long long *>(h->peer_base[up]) + 32;
    g.halo_flags = reinterpret_cast<unsigned long long *>(h->d_peer) + 32;  // [0] from below, [1] from above
    g.halo_epoch = &reinterpret_cast<PeerCtl *>(h->d_peer_ctl)->fused_epoch;
    g.halo_done = &reinterpret_cast<PeerCtl *>(h->d_peer_ctl)->fused_done;
    g.fused = 1;
    return MM_OK;
}

// Collective over the slab communicator: allocate and exchange the peer blocks.  Every rank ends with the same
// peer_mode (the success flags are min-reduced), otherwise a rank waiting on a flag nobody writes would hang.
static int peer_setup(mm_handle *h) {
    const int P = h->slab_count, r = h->slab_rank;
    const char *env = getenv("MICMEC_B200_PEER");
    int ok = (env && atoi(env) == 0) ? 0 : 1;
    if (P > 16 || !h->sg.active) ok = 0;
    ncclComm_t comm = (ncclComm_t)h->comm;
    const int64_t plane = h->sg.plane;
    const size_t bytes = peer_inbox_off(P) + sizeof(double) * 2 * 2 * 9 * (size_t)plane;
    unsigned char *d_hdl = nullptr;
    int *d_ok = nullptr;
    MM_CUDA(cudaMalloc(&d_hdl, 64 * (size_t)P));
    MM_CUDA(cudaMalloc(&d_ok, sizeof(int)));
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) {
        if (cudaMalloc(&h->d_peer, bytes) != cudaSuccess || cudaMalloc(&h->d_peer_ctl, sizeof(PeerCtl)) != cudaSuccess ||
            cudaMalloc(&h->d_peer_base, sizeof(char *) * 16) != cudaSuccess || cudaIpcGetMemHandle(&mine, h->d_peer) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        } else {
            MM_CUDA(cudaMemsetAsync(h->d_peer, 0, bytes, h->stream));
            MM_CUDA(cudaMemsetAsync(h->d_peer_ctl, 0, sizeof(PeerCtl), h->stream));
        }
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    MM_CUDA(cudaMemcpyAsync(d_hdl + 64 * (size_t)r, &mine, 64, cudaMemcpyHostToDevice, h->stream));
    MM_NCCL(g_nccl.AllGather(d_hdl + 64 * (size_t)r, d_hdl, 64, ncclUint8, comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all(P);
    MM_CUDA(cudaMemcpyAsync(all.data(), d_hdl, 64 * (size_t)P, cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    if (ok) {
        for (int q = 0; q < P && ok; q++) {
            if (q == r) {
                h->peer_base[q] = h->d_peer;
                continue;
            }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            } else {
                h->peer_base[q] = (char *)p;
            }
        }
    }
    // agree: min over the ranks (also orders "my block is zeroed" before anybody's first push)
    MM_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    MM_NCCL(g_nccl.AllReduce(d_ok, d_ok, 1, 2 /* ncclInt32 */, ncclMin, comm, h->stream));
    int all_ok = 0;
    MM_CUDA(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(d_hdl);
    cudaFree(d_ok);
    if (!all_ok) {
        comm_peer_free(h);
        return MM_OK;
    }
    MM_CUDA(cudaMemcpyAsync(h->d_peer_base, h->peer_base, sizeof(char *) * 16, cudaMemcpyHostToDevice, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    h->peer_plane = plane;
    h->peer_mode = 1;
    return peer_setup_fused(h);
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
    if (h->peer_mode) {
        const size_t off = peer_inbox_off(P);
        k_peer_push<<<grid, 256, 0, h->stream>>>(a, g.plane, g.nzl, h->peer_base[up], h->peer_base[down], off, (const PeerCtl *)h->d_peer_ctl);
        k_peer_unpack<<<grid, 256, 0, h->stream>>>(a, g.plane, g.nzl, h->d_peer, off, (PeerCtl *)h->d_peer_ctl, (unsigned long long)grid,
                                                   g.d_sc, r == 0 ? -1.0 : 0.0, r == P - 1 ? 1.0 : 0.0);
        h->launches += 2;
        MM_CUDA(cudaGetLastError());
        return MM_OK;
    }
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
    if (pn == pc + 7 && nbn == nbc && nbc > 1) {  // structured path: one coalesced pass over the 16-slot records
        double all[16];
        partials_sum16(pc, nbc, all);
        for (int k = 0; k < 7; k++) {
            a[k] = all[k];
            b[k] = all[7 + k];
        }
    } else {
        if (nbc > 0) partials_sum<7>(pc, nbc, kRedSlots, a);
        if (nbn > 0) partials_sum<7>(pn, nbn, kRedSlots, b);
    }
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
    if (h->peer_mode) {
        k_peer_allreduce<<<1, 256, 0, h->stream>>>(pc, nbc, pn, nbn, pd, nbd, h->d_peer_base, h->slab_count, h->slab_rank,
                                                   (PeerCtl *)h->d_peer_ctl, h->d_red);
        h->launches++;
        MM_CUDA(cudaGetLastError());
        return MM_OK;
    }
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
    return peer_setup(h);
}

int mm_comm_mode(const mm_handle *h) {
    if (!h || !h->comm) return -1;
    if (!h->peer_mode) return 0;
    return h->sg.fused ? 2 : 1;
}

int mm_comm_destroy(mm_handle *h) {
    if (h && h->comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
    }
    return MM_OK;
}

}  // extern "C"
