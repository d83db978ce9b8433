// mm_eigh.cu - batched symmetric eigen-decomposition of many small dense matrices (n <= 96), one thread block per matrix.
//
// Why it exists: the quasi-Newton optimisers of the reference (micmec/sampling/opt.py:196-197, 334-336) diagonalise their
// Hessian model (ndof x ndof, 81 or 87 for the 27-node systems of BASELINE.json config 5) at every step.  With 10 240
// replicas optimised in lockstep (micmec_b200/sampling/batchopt.py) that is 10 240 independent 81 x 81 problems per sweep;
// LAPACK on the host needs ~0.7 ms each and the library route (cuSOLVER syevd looped over the batch) is no faster.
//
// Method: two-sided cyclic Jacobi with round-robin ("chess tournament") ordering.  The matrix A and the accumulated
// rotations V live in shared memory (row pitch odd: column walks are bank-conflict free).  A round consists of n'/2
// DISJOINT index pairs, so all its rotations commute: (1) one thread per pair computes (c, s) from a_pp, a_qq, a_pq;
// (2) columns p, q of A and of V are rotated, one work item per (pair, row); (3) rows p, q of A are rotated, one work item
// per (pair, column).  n' - 1 rounds make a sweep; sweeps repeat until the off-diagonal norm is below 1e-15 of the
// Frobenius norm (typically 7-9 sweeps).  Every phase is a strided loop over independent work items followed by a block
// barrier, so the same source runs serially on the host (MM_EIGH_HOST, tests/eigh_host_check.cpp) to check the arithmetic.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#ifndef MM_EIGH_HOST
#include <cuda_runtime.h>

#include "../../include/micmec_b200.h"
#include "mm_internal.h"
#define EIGH_DEVICE __device__ __forceinline__
#define EIGH_SYNC() __syncthreads()
#define EIGH_TID ((int)threadIdx.x)
#define EIGH_NT ((int)blockDim.x)
#else
#define EIGH_DEVICE static inline
#define EIGH_SYNC() ((void)0)
#define EIGH_TID 0
#define EIGH_NT 1
#endif

namespace mm {

constexpr int kEighMaxN = 96;
constexpr int kEighMaxSweeps = 30;

// partner table of the round-robin schedule: in round r (0 .. m-2) of m players (m even), player m-1 stays, the others
// rotate; pair k joins the players at circle positions k and m-1-k.
EIGH_DEVICE void rr_pair(int m, int round, int k, int &p, int &q) {
    auto player = [&](int pos) { return pos == m - 1 ? m - 1 : (pos + round) % (m - 1); };
    int a = player(k), b = player(m - 1 - k);
    p = a < b ? a : b;
    q = a < b ? b : a;
}

// One matrix.  A, V: [n][ld] in (shared) memory, ld >= n; cs: [m/2][2] rotation parameters; pq: [m/2][2] pair indices;
// red: scratch of at least EIGH_NT doubles; out_w [n] and out_v [n][n] (row-major, eigenvector i in COLUMN i) receive the
// result in ascending order.  Returns the number of sweeps (every thread returns the same value).
EIGH_DEVICE int jacobi_eigh(int n, int ld, double *A, double *V, double *cs, int *pq, double *red, double *out_w, double *out_v) {
    const int tid = EIGH_TID, nt = EIGH_NT;
    const int m = (n + 1) & ~1;  // even number of players; index n (if any) is a bye
    const int half = m / 2;
    for (int w = tid; w < n * n; w += nt) V[(w / n) * ld + (w % n)] = (w / n == w % n) ? 1.0 : 0.0;
    EIGH_SYNC();
    int sweep = 0;
    for (; sweep < kEighMaxSweeps; sweep++) {
        // convergence: off-diagonal vs total Frobenius norm (strided partial sums, then a serial sum by every thread)
        double off = 0.0, tot = 0.0;
        for (int w = tid; w < n * n; w += nt) {
            const int i = w / n, j = w % n;
            const double a = A[i * ld + j];
            tot += a * a;
            if (i != j) off += a * a;
        }
        red[2 * tid] = off;
        red[2 * tid + 1] = tot;
        EIGH_SYNC();
        off = tot = 0.0;
        for (int t = 0; t < nt; t++) {
            off += red[2 * t];
            tot += red[2 * t + 1];
        }
        EIGH_SYNC();
        if (off <= 1e-30 * tot || tot == 0.0) break;
        for (int round = 0; round < m - 1; round++) {
            // (1) rotation parameters of the disjoint pairs of this round
            for (int k = tid; k < half; k += nt) {
                int p, q;
                rr_pair(m, round, k, p, q);
                double c = 1.0, s = 0.0;
                if (q < n) {
                    const double apq = A[p * ld + q];
                    if (apq != 0.0) {
                        const double tau = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                } else {
                    q = -1;  // bye
                }
                pq[2 * k] = p;
                pq[2 * k + 1] = q;
                cs[2 * k] = c;
                cs[2 * k + 1] = s;
            }
            EIGH_SYNC();
            // (2) columns p, q of A and V:  x_p' = c x_p - s x_q,  x_q' = s x_p + c x_q
            for (int w = tid; w < half * n; w += nt) {
                const int k = w / n, i = w % n;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (q < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double ap = A[i * ld + p], aq = A[i * ld + q];
                A[i * ld + p] = c * ap - s * aq;
                A[i * ld + q] = s * ap + c * aq;
                const double vp = V[i * ld + p], vq = V[i * ld + q];
                V[i * ld + p] = c * vp - s * vq;
                V[i * ld + q] = s * vp + c * vq;
            }
            EIGH_SYNC();
            // (3) rows p, q of A
            for (int w = tid; w < half * n; w += nt) {
                const int k = w / n, j = w % n;
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (q < 0) continue;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double ap = A[p * ld + j], aq = A[q * ld + j];
                A[p * ld + j] = c * ap - s * aq;
                A[q * ld + j] = s * ap + c * aq;
            }
            EIGH_SYNC();
            // the rotated pair is diagonal by construction: remove the rounding residue so that it cannot feed back
            for (int k = tid; k < half; k += nt) {
                const int p = pq[2 * k], q = pq[2 * k + 1];
                if (q >= 0 && cs[2 * k + 1] != 0.0) A[p * ld + q] = A[q * ld + p] = 0.0;
            }
            EIGH_SYNC();
        }
    }
    // ascending order: rank of every eigenvalue (ties broken by index), then a scatter of values and vector columns
    for (int i = tid; i < n; i += nt) {
        const double d = A[i * ld + i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const double e = A[j * ld + j];
            rank += (e < d || (e == d && j < i)) ? 1 : 0;
        }
        pq[i] = rank;  // n <= 2 * half entries available
        out_w[rank] = d;
    }
    EIGH_SYNC();
    for (int w = tid; w < n * n; w += nt) {
        const int r = w / n, i = w % n;
        out_v[r * n + pq[i]] = V[r * ld + i];
    }
    EIGH_SYNC();
    return sweep;
}

#ifndef MM_EIGH_HOST
// mask (may be null): matrices whose entry is zero are skipped (the lockstep optimiser only needs the spectra of the
// replicas whose Hessian model has just changed)
__global__ void __launch_bounds__(256) k_batched_eigh(const double *__restrict__ mats, int n, int ld, double *__restrict__ evals,
                                                      double *__restrict__ evecs, int *__restrict__ sweeps, const int *__restrict__ mask) {
    if (mask && !mask[blockIdx.x]) return;
    extern __shared__ __align__(16) double sm[];
    double *A = sm;
    double *V = A + (size_t)n * ld;
    double *cs = V + (size_t)n * ld;
    double *red = cs + (kEighMaxN + 2);
    int *pq = reinterpret_cast<int *>(red + 2 * 256);
    const size_t b = blockIdx.x;
    const double *src = mats + b * (size_t)n * n;
    // symmetrise on load (LAPACK's eigh reads one triangle; the models are symmetric up to rounding)
    for (int w = threadIdx.x; w < n * n; w += blockDim.x) {
        const int i = w / n, j = w % n;
        A[i * ld + j] = 0.5 * (src[i * n + j] + src[j * n + i]);
    }
    __syncthreads();
    const int nsweep = jacobi_eigh(n, ld, A, V, cs, pq, red, evals + b * n, evecs + b * (size_t)n * n);
    if (threadIdx.x == 0 && sweeps) sweeps[b] = nsweep;
}
#endif

}  // namespace mm

#ifndef MM_EIGH_HOST
namespace mm {
// device arrays in, device arrays out, on the caller's stream (mm_qn.cu)
int eigh_launch_device(int device, int64_t batch, int n, const double *d_mats, double *d_evals, double *d_evecs, int *d_sweeps,
                       const int *d_mask, cudaStream_t stream) {
    const int ld = (n & 1) ? n : n + 1;
    const size_t smem = sizeof(double) * (2 * (size_t)n * ld + (kEighMaxN + 2) + 2 * 256) + sizeof(int) * (kEighMaxN + 2);
    static bool configured[64] = {false};
    if (!configured[device & 63]) {
        MM_CUDA(cudaFuncSetAttribute(k_batched_eigh, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[device & 63] = true;
    }
    k_batched_eigh<<<(unsigned)batch, 256, smem, stream>>>(d_mats, n, ld, d_evals, d_evecs, d_sweeps, d_mask);
    MM_CUDA(cudaGetLastError());
    return MM_OK;
}
}  // namespace mm

extern "C" int mm_batched_eigh(int device, int64_t batch, int32_t n, const double *mats, int where, double *evals, double *evecs,
                               int32_t *max_sweeps_out) {
    using namespace mm;
    if (batch < 0 || n < 1 || n > kEighMaxN || !mats || !evals || !evecs) {
        set_error("mm_batched_eigh: need 1 <= n <= 96 and non-null arrays");
        return MM_ERR_INVALID;
    }
    if (batch == 0) return MM_OK;
    MM_CUDA(cudaSetDevice(device));
    const int ld = (n & 1) ? n : n + 1;  // odd pitch
    const size_t smem = sizeof(double) * (2 * (size_t)n * ld + (kEighMaxN + 2) + 2 * 256) + sizeof(int) * (kEighMaxN + 2);
    static bool configured[64] = {false};
    if (!configured[device & 63]) {
        MM_CUDA(cudaFuncSetAttribute(k_batched_eigh, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[device & 63] = true;
    }
    const size_t nmat = (size_t)batch * n * n, nval = (size_t)batch * n;
    const double *d_m = mats;
    double *d_w = evals, *d_v = evecs, *own = nullptr;
    int *d_sweeps = nullptr;
    cudaStream_t stream = nullptr;
    MM_CUDA(cudaMalloc(&d_sweeps, sizeof(int) * (size_t)batch));
    if (where == MM_HOST) {
        MM_CUDA(cudaMalloc(&own, sizeof(double) * (2 * nmat + nval)));
        MM_CUDA(cudaMemcpyAsync(own, mats, sizeof(double) * nmat, cudaMemcpyHostToDevice, stream));
        d_m = own;
        d_v = own + nmat;
        d_w = own + 2 * nmat;
    }
    k_batched_eigh<<<(unsigned)batch, 256, smem, stream>>>(d_m, n, ld, d_w, d_v, d_sweeps, nullptr);
    cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess && where == MM_HOST) {
        err = cudaMemcpyAsync(evals, d_w, sizeof(double) * nval, cudaMemcpyDeviceToHost, stream);
        if (err == cudaSuccess) err = cudaMemcpyAsync(evecs, d_v, sizeof(double) * nmat, cudaMemcpyDeviceToHost, stream);
    }
    if (err == cudaSuccess && max_sweeps_out) {
        // the largest sweep count of the batch: kEighMaxSweeps means some matrix did not converge
        int *h = new int[batch];
        err = cudaMemcpyAsync(h, d_sweeps, sizeof(int) * (size_t)batch, cudaMemcpyDeviceToHost, stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
        int worst = 0;
        for (int64_t i = 0; i < batch; i++) worst = h[i] > worst ? h[i] : worst;
        *max_sweeps_out = worst;
        delete[] h;
    }
    if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
    cudaFree(own);
    cudaFree(d_sweeps);
    if (err != cudaSuccess) {
        set_error(std::string("mm_batched_eigh: ") + cudaGetErrorString(err));
        return MM_ERR_CUDA;
    }
    return MM_OK;
}
#endif
