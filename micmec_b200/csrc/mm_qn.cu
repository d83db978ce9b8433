// mm_qn.cu - device-resident lockstep quasi-Newton optimisation of a batch of independent replicas (BASELINE.json config 5;
// SURVEY.md 8(f) rank 1).
//
// Stands in for (relative to /root/reference), per replica:
//   QNOptimizer.initialize / propagate / make_step                  micmec/sampling/opt.py:317-393
//   SR1HessianModel.update, HessianModel.get_spectrum               micmec/sampling/opt.py:196-197, 239-255
//   solve_trust_radius                                              micmec/sampling/opt.py:396-443
//   CartesianDOF.fun / check_convergence                            micmec/sampling/dof.py:132-193
//
// Round 1 ran this state machine in batched NumPy on the host (sampling/batchopt.py) with the forces and the spectra on
// the device: 97 replicas/s, bound by the [R, n, n] host arrays and two 0.5 GB PCIe transfers per sweep.  Here the whole
// sweep stays on the device: Hessian models, spectra, trial points, trust radii and convergence state live in HBM, one
// block per replica, and the host only reads one counter per sweep.
//   k_qn_refresh   SR1 update of the models that accepted a step (or reset to identity when the update is unsafe)
//   k_batched_eigh spectra of the models that changed (mm_eigh.cu, masked)
//   k_qn_step      gradient in the eigenbasis, ridge search (bracketing + false position, as the reference), trial point
//   force kernels  the batch handle's indexed kernels on the trial points (mm_force.cu)
//   k_qn_accept    accept / shrink, trust radius, convergence criteria, bookkeeping
#include <cmath>
#include <new>

#include "mm_internal.h"

namespace mm {
int eigh_launch_device(int device, int64_t batch, int n, const double *d_mats, double *d_evals, double *d_evecs, int *d_sweeps,
                       const int *d_mask, cudaStream_t stream);

constexpr int kQnThreads = 128;
constexpr int kQnMaxN = 96;

struct QnArrays {
    int n, nnodes;          // degrees of freedom (3 nnodes, + 6 cell variables first for the strain DOF) and nodes per replica
    int strain;             // 0: CartesianDOF; 1: StrainCellDOF (dof.py:522-697): x = [6 strain variables, fractional coordinates]
    double *rvecs0, *jac, *proj;        // strain: [R][9] reference cell, [R][9][6] d rvecs / d strain, [R][9][9] projector on its range
    double *gtrial, *grv, *last_rv;     // [R][n] gradient with respect to x at the trial point; [R][9] projected cell gradient, last cell
    double *pos_trial, *rv_trial;       // the batch handle's position / domain-vector arrays (Cartesian trial geometry)
    double grvecs_rms, drvecs_rms;
    double *H, *V, *w;      // [R][n][n] model, eigenvectors (columns), [R][n] eigenvalues
    double *x, *xold, *g, *gold, *ge, *trial, *last;  // [R][n]
    double *f, *fold, *radius, *rnorm, *conv_val;     // [R]
    int *fresh, *live, *failed, *converged, *started, *iters, *conv_count, *need;  // [R]
    double initial_radius, small_radius, too_small_radius, gpos_rms, dpos_rms;
    int *nlive;             // [1]
};

__device__ __forceinline__ double block_sum(double v, double *scratch) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < kQnThreads / 32; i++) s += scratch[i];
    return s;
}
__device__ __forceinline__ double block_max(double v, double *scratch) {
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, off));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double s = scratch[0];
    for (int i = 1; i < kQnThreads / 32; i++) s = fmax(s, scratch[i]);
    return s;
}

// SR1 update of the replicas that accepted a step in the last sweep (opt.py:322-331, 239-255); marks the models whose
// spectrum has to be recomputed
__global__ void __launch_bounds__(kQnThreads) k_qn_refresh(const QnArrays a) {
    __shared__ double dx[kQnMaxN], resid[kQnMaxN], scratch[8];
    const int r = blockIdx.x, n = a.n, t = threadIdx.x;
    const size_t o = (size_t)r * n;
    double *H = a.H + o * n;
    const bool upd = a.fresh[r] && a.iters[r] > 0;  // block-uniform
    if (upd) {
        for (int i = t; i < n; i += kQnThreads) dx[i] = a.x[o + i] - a.xold[o + i];
        __syncthreads();
        for (int i = t; i < n; i += kQnThreads) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s = fma(H[(size_t)i * n + j], dx[j], s);
            resid[i] = (a.g[o + i] - a.gold[o + i]) - s;
        }
        __syncthreads();
        double pd = 0.0, pxx = 0.0, prr = 0.0;
        for (int i = t; i < n; i += kQnThreads) {
            pd += resid[i] * dx[i];
            pxx += dx[i] * dx[i];
            prr += resid[i] * resid[i];
        }
        const double denom = block_sum(pd, scratch), nx = sqrt(block_sum(pxx, scratch)), nr = sqrt(block_sum(prr, scratch));
        if (fabs(denom) > 1e-5 * nx * nr) {
            const double coef = 1.0 / denom;
            for (int w = t; w < n * n; w += kQnThreads) H[w] += resid[w / n] * (coef * resid[w % n]);
        } else {  // a failed update poisons the model: identity and the initial radius again (opt.py:325-329)
            for (int w = t; w < n * n; w += kQnThreads) H[w] = (w / n == w % n) ? 1.0 : 0.0;
            if (t == 0) a.radius[r] = a.initial_radius;
        }
        for (int i = t; i < n; i += kQnThreads) {
            a.xold[o + i] = a.x[o + i];
            a.gold[o + i] = a.g[o + i];
        }
        if (t == 0) a.fold[r] = a.f[r];
    }
    if (t == 0) a.need[r] = (a.fresh[r] && a.live[r]) ? 1 : 0;
}

// |grad / (evals + ridge)| - radius for one replica, evaluated by one warp
__device__ __forceinline__ double qn_excess(const double *ge, const double *w, int n, double ridge, double radius) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) {
        const double q = ge[i] / (w[i] + ridge);
        s = fma(q, q, s);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return sqrt(s) - radius;
}

// Cartesian geometry of x (dof.py:553-581): rvecs = A(s) rvecs0 with A_ii = s_i, A_ij = s_k / 2; pos = frac rvecs
__global__ void __launch_bounds__(kQnThreads) k_qn_geometry(const QnArrays a, const double *xsrc) {
    __shared__ double rv[9];
    const int r = blockIdx.x, t = threadIdx.x;
    const size_t o = (size_t)r * a.n;
    if (!a.strain) {
        for (int i = t; i < a.n; i += kQnThreads) a.pos_trial[(size_t)r * 3 * a.nnodes + i] = xsrc[o + i];
        return;
    }
    if (t < 9) {
        const double *s6 = xsrc + o;  // 00 11 22 12 20 01 (dof.py:522-531)
        const double A[9] = {s6[0], 0.5 * s6[5], 0.5 * s6[4], 0.5 * s6[5], s6[1], 0.5 * s6[3], 0.5 * s6[4], 0.5 * s6[3], s6[2]};
        const int i = t / 3, j = t % 3;
        const double *r0 = a.rvecs0 + (size_t)r * 9;
        const double v = A[i * 3] * r0[j] + A[i * 3 + 1] * r0[3 + j] + A[i * 3 + 2] * r0[6 + j];
        rv[t] = v;
        a.rv_trial[(size_t)r * 9 + t] = v;
    }
    __syncthreads();
    for (int w = t; w < 3 * a.nnodes; w += kQnThreads) {
        const int v = w / 3, j = w % 3;
        const double *fr = xsrc + o + 6 + 3 * v;
        a.pos_trial[(size_t)r * 3 * a.nnodes + w] = fr[0] * rv[j] + fr[1] * rv[3 + j] + fr[2] * rv[6 + j];
    }
}

// gradient with respect to x from the Cartesian gradient and the virial (dof.py:342-385, 582-697):
// grvecs = rvecs^-T vtens;  gx[:6] = grvecs . jac;  gx[6:] = gpos rvecs^T;  projected cell gradient for the criteria
__global__ void __launch_bounds__(kQnThreads) k_qn_gradient(const QnArrays a, const double *rep, const double *gpos) {
    __shared__ double rv[9], grv[9];
    const int r = blockIdx.x, t = threadIdx.x;
    const size_t o = (size_t)r * a.n, op = (size_t)r * 3 * a.nnodes;
    if (!a.strain) {
        for (int i = t; i < a.n; i += kQnThreads) a.gtrial[o + i] = gpos[op + i];
        return;
    }
    if (t < 9) rv[t] = a.rv_trial[(size_t)r * 9 + t];
    __syncthreads();
    if (t == 0) {
        const double *q = rep + (size_t)r * 8;
        const double vt[9] = {q[1], q[6], q[5], q[6], q[2], q[4], q[5], q[4], q[3]};  // 00 11 22 12 02 01 -> full
        const double det = rv[0] * (rv[4] * rv[8] - rv[5] * rv[7]) - rv[1] * (rv[3] * rv[8] - rv[5] * rv[6]) + rv[2] * (rv[3] * rv[7] - rv[4] * rv[6]);
        double inv[9];  // rvecs^-1
        inv[0] = (rv[4] * rv[8] - rv[5] * rv[7]) / det;
        inv[1] = (rv[2] * rv[7] - rv[1] * rv[8]) / det;
        inv[2] = (rv[1] * rv[5] - rv[2] * rv[4]) / det;
        inv[3] = (rv[5] * rv[6] - rv[3] * rv[8]) / det;
        inv[4] = (rv[0] * rv[8] - rv[2] * rv[6]) / det;
        inv[5] = (rv[2] * rv[3] - rv[0] * rv[5]) / det;
        inv[6] = (rv[3] * rv[7] - rv[4] * rv[6]) / det;
        inv[7] = (rv[1] * rv[6] - rv[0] * rv[7]) / det;
        inv[8] = (rv[0] * rv[4] - rv[1] * rv[3]) / det;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)  // gvecs = inv^T: grvecs[i][j] = sum_k inv[k][i] vt[k][j]
                grv[i * 3 + j] = inv[i] * vt[j] + inv[3 + i] * vt[3 + j] + inv[6 + i] * vt[6 + j];
    }
    __syncthreads();
    if (t < 6) {
        const double *jac = a.jac + (size_t)r * 54;
        double sgx = 0.0;
        for (int i = 0; i < 9; i++) sgx += grv[i] * jac[i * 6 + t];
        a.gtrial[o + t] = sgx;
    }
    if (t >= 32 && t < 41) {
        const double *pj = a.proj + (size_t)r * 81 + (t - 32) * 9;
        double sp = 0.0;
        for (int i = 0; i < 9; i++) sp += pj[i] * grv[i];
        a.grv[(size_t)r * 9 + (t - 32)] = sp;
    }
    for (int w = t; w < 3 * a.nnodes; w += kQnThreads) {
        const int v = w / 3, j = w % 3;
        const double *g3 = gpos + op + 3 * v;
        a.gtrial[o + 6 + w] = g3[0] * rv[j * 3] + g3[1] * rv[j * 3 + 1] + g3[2] * rv[j * 3 + 2];
    }
}

// trial point of every live replica: gradient in the eigenbasis of its model, ridge search, step (opt.py:334-356, 396-443)
__global__ void __launch_bounds__(kQnThreads) k_qn_step(const QnArrays a) {
    __shared__ double ge[kQnMaxN], w[kQnMaxN], delta[kQnMaxN], scratch[8];
    __shared__ double s_ridge;
    __shared__ int s_newton;
    const int r = blockIdx.x, n = a.n, t = threadIdx.x;
    const size_t o = (size_t)r * n;
    const double *V = a.V + o * n;
    if (!a.live[r]) {  // frozen replicas are still evaluated, at their accepted point
        for (int i = t; i < n; i += kQnThreads) a.trial[o + i] = a.x[o + i];
        if (t == 0) a.rnorm[r] = 0.0;
        return;
    }
    if (a.need[r]) {  // new spectrum: project the gradient (evecs in columns)
        for (int i = t; i < n; i += kQnThreads) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s = fma(V[(size_t)j * n + i], a.gold[o + j], s);
            a.ge[o + i] = s;
        }
        if (t == 0) a.fresh[r] = 0;
    }
    __syncthreads();
    for (int i = t; i < n; i += kQnThreads) {
        ge[i] = a.ge[o + i];
        w[i] = a.w[o + i];
    }
    __syncthreads();
    const double radius = a.radius[r];
    if (t < 32) {  // one warp runs the scalar iteration; every excess() is a warp reduction over the spectrum
        double emin = w[0], emax = w[0];
        for (int i = 1; i < n; i++) {
            emin = fmin(emin, w[i]);
            emax = fmax(emax, w[i]);
        }
        bool newton = false;
        if (emin > 0.0) newton = qn_excess(ge, w, n, 0.0, radius) <= 0.0;  // the Newton step fits inside the radius
        double ridge = 0.0;
        if (!newton) {
            const double ridge_min = -emin;
            // bracket from below: just above the pole, halving the offset until the step is too long
            double alpha = fmin(1e1, fabs(emax)), a0 = 0.0, a1 = 0.0;
            for (int it = 0; it < 20000; it++) {
                a0 = ridge_min + alpha;
                a1 = qn_excess(ge, w, n, a0, radius);
                if (-a1 < 0.0) break;
                alpha *= 0.5;
            }
            // bracket from above: doubling the offset until the step is short enough
            double b0 = 0.0, b1 = 0.0;
            alpha = fmax(1e-5, fabs(ridge_min));
            for (int it = 0; it < 20000; it++) {
                b0 = ridge_min + alpha;
                b1 = qn_excess(ge, w, n, b0, radius);
                if (b1 < 0.0) break;
                alpha *= 2.0;
            }
            for (int it = 0; it < 20000; it++) {  // false position (opt.py:428-441)
                ridge = (a1 * a0 - b1 * b0) / (a1 - b1);
                const double val = qn_excess(ge, w, n, ridge, radius);
                if (val > 0.0 && a1 > 0.0) {
                    a0 = ridge;
                    a1 = val;
                } else {
                    b0 = ridge;
                    b1 = val;
                }
                if (!(fabs(val) > radius * 1e-5)) break;
            }
        }
        if (t == 0) {
            s_ridge = ridge;
            s_newton = newton ? 1 : 0;
        }
    }
    __syncthreads();
    double pn = 0.0;
    for (int i = t; i < n; i += kQnThreads) {
        const double d = -ge[i] / (w[i] + s_ridge);
        delta[i] = d;
        pn = fma(d, d, pn);
    }
    const double rn = sqrt(block_sum(pn, scratch));
    if (t == 0) a.rnorm[r] = rn;
    __syncthreads();
    for (int i = t; i < n; i += kQnThreads) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s = fma(V[(size_t)i * n + j], delta[j], s);
        a.trial[o + i] = a.xold[o + i] + s;
    }
}

// accept / shrink (opt.py:358-393), convergence criteria of CartesianDOF (dof.py:155-193)
__global__ void __launch_bounds__(kQnThreads) k_qn_accept(const QnArrays a, const double *rep, const double *gpos) {
    __shared__ double scratch[8];
    const int r = blockIdx.x, n = a.n, t = threadIdx.x;
    const size_t o = (size_t)r * n, op = (size_t)r * 3 * a.nnodes;
    if (!a.live[r]) return;
    const double f = rep[(size_t)r * 8];
    double pg = 0.0, pgo = 0.0;
    for (int i = t; i < n; i += kQnThreads) {
        pg = fma(a.gtrial[o + i], a.gtrial[o + i], pg);
        pgo = fma(a.gold[o + i], a.gold[o + i], pgo);
    }
    const double gn = sqrt(block_sum(pg, scratch)), gon = sqrt(block_sum(pgo, scratch));
    const double radius = a.radius[r];
    bool shrink = (f - a.fold[r] > 0.0) || (radius < a.small_radius && gn - gon > 0.0);
    if (!(f == f)) shrink = true;  // a NaN energy is never accepted
    if (shrink) {
        if (t == 0) {  // halve until strictly inside the step that was just tried (opt.py:376-382)
            double tr = radius * 0.5;
            const double rn = a.rnorm[r];
            while (tr >= rn && tr > 0.0) tr *= 0.5;
            a.radius[r] = tr;
            if (tr < a.too_small_radius) {
                a.failed[r] = 1;
                a.live[r] = 0;
            }
        }
        return;
    }
    // accepted
    const bool started = a.started[r] != 0;
    double gmax2 = 0.0, gsum2 = 0.0, dmax2 = 0.0, dsum2 = 0.0;
    for (int v = t; v < a.nnodes; v += kQnThreads) {
        double g2 = 0.0, d2 = 0.0;
        for (int c = 0; c < 3; c++) {
            const double gc = gpos[op + 3 * v + c], dc = a.pos_trial[op + 3 * v + c] - a.last[op + 3 * v + c];
            g2 = fma(gc, gc, g2);
            d2 = fma(dc, dc, d2);
        }
        gmax2 = fmax(gmax2, g2);
        dmax2 = fmax(dmax2, d2);
        gsum2 += g2;
        dsum2 += d2;
    }
    gmax2 = block_max(gmax2, scratch);
    dmax2 = block_max(dmax2, scratch);
    gsum2 = block_sum(gsum2, scratch);
    dsum2 = block_sum(dsum2, scratch);
    for (int i = t; i < n; i += kQnThreads) {
        a.x[o + i] = a.trial[o + i];
        a.g[o + i] = a.gtrial[o + i];
    }
    for (int i = t; i < 3 * a.nnodes; i += kQnThreads) a.last[op + i] = a.pos_trial[op + i];
    if (t == 0) {
        a.f[r] = f;
        if (radius < a.initial_radius) a.radius[r] = radius * 2.0;
        double cell[4] = {0.0, 0.0, 0.0, 0.0};  // cell criteria (dof.py:382-449): rows of the 3 x 3 arrays
        if (a.strain) {
            double cmax2 = 0.0, csum2 = 0.0, emax2 = 0.0, esum2 = 0.0;
            for (int i = 0; i < 3; i++) {
                double c2 = 0.0, e2 = 0.0;
                for (int j = 0; j < 3; j++) {
                    const double gc = a.grv[(size_t)r * 9 + i * 3 + j];
                    const double dc = a.rv_trial[(size_t)r * 9 + i * 3 + j] - a.last_rv[(size_t)r * 9 + i * 3 + j];
                    c2 += gc * gc;
                    e2 += dc * dc;
                }
                cmax2 = fmax(cmax2, c2);
                emax2 = fmax(emax2, e2);
                csum2 += c2;
                esum2 += e2;
            }
            cell[0] = sqrt(csum2 / 3.0) / a.grvecs_rms;
            cell[1] = sqrt(cmax2) / (3.0 * a.grvecs_rms);
            cell[2] = sqrt(esum2 / 3.0) / a.drvecs_rms;
            cell[3] = sqrt(emax2) / (3.0 * a.drvecs_rms);
            for (int i = 0; i < 9; i++) a.last_rv[(size_t)r * 9 + i] = a.rv_trial[(size_t)r * 9 + i];
        }
        if (started) {
            const double ratios[8] = {sqrt(gsum2 / a.nnodes) / a.gpos_rms, sqrt(gmax2) / (3.0 * a.gpos_rms),
                                      sqrt(dsum2 / a.nnodes) / a.dpos_rms, sqrt(dmax2) / (3.0 * a.dpos_rms),
                                      cell[0], cell[1], cell[2], cell[3]};
            double worst = 0.0;
            int count = 0;
            for (int q = 0; q < (a.strain ? 8 : 4); q++) {
                worst = fmax(worst, ratios[q]);
                count += ratios[q] >= 1.0 ? 1 : 0;
            }
            a.conv_val[r] = worst;
            a.conv_count[r] = count;
            if (count == 0) {
                a.converged[r] = 1;
                a.live[r] = 0;
            }
        }
        a.started[r] = 1;
        a.iters[r] += 1;
        a.fresh[r] = 1;
    }
}

__global__ void k_qn_count(const int *live, int nrep, int *out) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int c = 0;
    for (int r = threadIdx.x; r < nrep; r += blockDim.x) c += live[r] ? 1 : 0;
    atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) *out = s;
}

__global__ void k_qn_init(const QnArrays a, int nrep, double radius0) {
    const int r = blockIdx.x, n = a.n;
    const size_t o = (size_t)r * n;
    for (int w = threadIdx.x; w < n * n; w += blockDim.x) a.H[o * n + w] = (w / n == w % n) ? 1.0 : 0.0;
    if (threadIdx.x == 0) {
        a.radius[r] = radius0;
        a.fresh[r] = 1;
        a.live[r] = 1;
        a.failed[r] = a.converged[r] = a.started[r] = a.iters[r] = 0;
        a.conv_val[r] = 2.0;
        a.conv_count[r] = -1;
        a.rnorm[r] = 0.0;
    }
}

// x0 evaluated: x = xold = x0, f, g (QNOptimizer.initialize, opt.py:317-320)
__global__ void k_qn_take_first(const QnArrays a, const double *rep) {
    const int r = blockIdx.x, n = a.n;
    const size_t o = (size_t)r * n, op = (size_t)r * 3 * a.nnodes;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        a.xold[o + i] = a.x[o + i];
        a.g[o + i] = a.gold[o + i] = a.gtrial[o + i];
    }
    for (int i = threadIdx.x; i < 3 * a.nnodes; i += blockDim.x) a.last[op + i] = a.pos_trial[op + i];
    if (a.strain && threadIdx.x < 9) a.last_rv[(size_t)r * 9 + threadIdx.x] = a.rv_trial[(size_t)r * 9 + threadIdx.x];
    if (threadIdx.x == 0) a.f[r] = a.fold[r] = rep[(size_t)r * 8];
}

}  // namespace mm

using namespace mm;

struct mm_qn {
    mm_handle *h = nullptr;
    int64_t nrep = 0;
    QnArrays a;
    double *block = nullptr;
    int *iblock = nullptr;
    int *d_sweeps = nullptr;
    int *h_nlive = nullptr;  // pinned
    int64_t evaluations = 0;
};


extern "C" {

int mm_qn_destroy(mm_qn *q) {
    if (!q) return MM_OK;
    cudaSetDevice(q->h->device);
    cudaFree(q->block);
    cudaFree(q->iblock);
    cudaFree(q->d_sweeps);
    if (q->h_nlive) cudaFreeHost(q->h_nlive);
    delete q;
    return MM_OK;
}

// trial (or initial) point xsrc -> Cartesian geometry -> batched force evaluation -> gradient with respect to x
static int qn_evaluate(mm_qn *q, const double *xsrc) {
    mm_handle *h = q->h;
    const unsigned R = (unsigned)q->nrep;
    k_qn_geometry<<<R, kQnThreads, 0, h->stream>>>(q->a, xsrc);
    h->pos_valid = true;
    const int rc = force_evaluate(h, h->d_gpos, false);
    if (rc != MM_OK) return rc;
    k_qn_gradient<<<R, kQnThreads, 0, h->stream>>>(q->a, h->d_rep, h->d_gpos);
    h->launches += 2;
    q->evaluations++;
    return MM_OK;
}

int mm_qn_create(mm_handle *h, int kind, const double *x0_host, const double *rvecs0_host, const double *jac_host,
                 const double *proj_host, double gpos_rms, double dpos_rms, double grvecs_rms, double drvecs_rms,
                 double trust_radius, double small_radius, double too_small_radius, mm_qn **out) {
    if (!h || !x0_host || !out || (kind != 0 && kind != 1) || (kind == 1 && (!rvecs0_host || !jac_host || !proj_host))) {
        set_error("mm_qn_create: null argument or unknown kind of degrees of freedom");
        return MM_ERR_INVALID;
    }
    *out = nullptr;
    if (h->nreplicas < 1 || h->nnodes % h->nreplicas != 0 || !h->d_rep || !h->d_rvecs_batch) {
        set_error("mm_qn_create: the handle is not a replica batch");
        return MM_ERR_INVALID;
    }
    const int64_t R = h->nreplicas, nn = h->nnodes / R;
    const int n = (int)(3 * nn) + (kind == 1 ? 6 : 0);
    if (n < 1 || n > kQnMaxN) {
        set_error("mm_qn_create: at most 96 degrees of freedom per replica (32 nodes, 30 with the cell variables)");
        return MM_ERR_INVALID;
    }
    if (!(gpos_rms > 0.0) || !(dpos_rms > 0.0) || (kind == 1 && (!(grvecs_rms > 0.0) || !(drvecs_rms > 0.0)))) {
        set_error("mm_qn_create: the device optimiser needs all convergence thresholds of its degrees of freedom");
        return MM_ERR_INVALID;
    }
    mm_qn *q = new (std::nothrow) mm_qn();
    if (!q) return MM_ERR_INVALID;
    q->h = h;
    q->nrep = R;
    MM_CUDA(cudaSetDevice(h->device));
    const size_t nn2 = (size_t)R * n * n, n1 = (size_t)R * n, p1 = (size_t)R * 3 * nn;
    const size_t ndbl = 2 * nn2 + 8 * n1 + p1 + 5 * (size_t)R + (size_t)R * (9 + 54 + 81 + 9 + 9);
    if (cudaMalloc(&q->block, sizeof(double) * ndbl) != cudaSuccess || cudaMalloc(&q->iblock, sizeof(int) * (8 * (size_t)R + 4)) != cudaSuccess ||
        cudaMalloc(&q->d_sweeps, sizeof(int) * (size_t)R) != cudaSuccess || cudaHostAlloc(&q->h_nlive, sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        mm_qn_destroy(q);
        set_error("mm_qn_create: out of device memory");
        return MM_ERR_CUDA;
    }
    QnArrays &a = q->a;
    a.n = n;
    a.nnodes = (int)nn;
    a.strain = kind;
    double *p = q->block;
    a.H = p; p += nn2;
    a.V = p; p += nn2;
    a.w = p; p += n1;
    a.x = p; p += n1;
    a.xold = p; p += n1;
    a.g = p; p += n1;
    a.gold = p; p += n1;
    a.ge = p; p += n1;
    a.trial = p; p += n1;
    a.gtrial = p; p += n1;
    a.last = p; p += p1;
    a.f = p; p += R;
    a.fold = p; p += R;
    a.radius = p; p += R;
    a.rnorm = p; p += R;
    a.conv_val = p; p += R;
    a.rvecs0 = p; p += 9 * R;
    a.jac = p; p += 54 * R;
    a.proj = p; p += 81 * R;
    a.grv = p; p += 9 * R;
    a.last_rv = p; p += 9 * R;
    a.pos_trial = h->d_pos;
    a.rv_trial = h->d_rvecs_batch;
    int *ip = q->iblock;
    a.fresh = ip; ip += R;
    a.live = ip; ip += R;
    a.failed = ip; ip += R;
    a.converged = ip; ip += R;
    a.started = ip; ip += R;
    a.iters = ip; ip += R;
    a.conv_count = ip; ip += R;
    a.need = ip; ip += R;
    a.nlive = ip;
    a.initial_radius = trust_radius;
    a.small_radius = small_radius;
    a.too_small_radius = too_small_radius;
    a.gpos_rms = gpos_rms;
    a.dpos_rms = dpos_rms;
    a.grvecs_rms = grvecs_rms;
    a.drvecs_rms = drvecs_rms;
    MM_CUDA(cudaMemsetAsync(q->block, 0, sizeof(double) * ndbl, h->stream));
    k_qn_init<<<(unsigned)R, kQnThreads, 0, h->stream>>>(a, (int)R, trust_radius);
    MM_CUDA(cudaMemcpyAsync(a.x, x0_host, sizeof(double) * n1, cudaMemcpyHostToDevice, h->stream));
    if (kind == 1) {
        MM_CUDA(cudaMemcpyAsync(a.rvecs0, rvecs0_host, sizeof(double) * 9 * R, cudaMemcpyHostToDevice, h->stream));
        MM_CUDA(cudaMemcpyAsync(a.jac, jac_host, sizeof(double) * 54 * R, cudaMemcpyHostToDevice, h->stream));
        MM_CUDA(cudaMemcpyAsync(a.proj, proj_host, sizeof(double) * 81 * R, cudaMemcpyHostToDevice, h->stream));
    }
    // QNOptimizer.initialize: evaluate x0
    const int rc = qn_evaluate(q, a.x);
    if (rc != MM_OK) {
        mm_qn_destroy(q);
        return rc;
    }
    k_qn_take_first<<<(unsigned)R, kQnThreads, 0, h->stream>>>(a, h->d_rep);
    h->launches += 2;
    MM_CUDA(cudaGetLastError());
    MM_CUDA(cudaStreamSynchronize(h->stream));
    *out = q;
    return MM_OK;
}

/* nsweeps lockstep sweeps (one batched force evaluation each); *nlive_out = replicas still moving afterwards */
int mm_qn_sweep(mm_qn *q, int nsweeps, int *nlive_out) {
    if (!q) {
        set_error("mm_qn_sweep: null optimiser");
        return MM_ERR_INVALID;
    }
    mm_handle *h = q->h;
    MM_CUDA(cudaSetDevice(h->device));
    const unsigned R = (unsigned)q->nrep;
    for (int s = 0; s < nsweeps; s++) {
        k_qn_refresh<<<R, kQnThreads, 0, h->stream>>>(q->a);
        int rc = eigh_launch_device(h->device, q->nrep, q->a.n, q->a.H, q->a.w, q->a.V, q->d_sweeps, q->a.need, h->stream);
        if (rc != MM_OK) return rc;
        k_qn_step<<<R, kQnThreads, 0, h->stream>>>(q->a);
        rc = qn_evaluate(q, q->a.trial);
        if (rc != MM_OK) return rc;
        k_qn_accept<<<R, kQnThreads, 0, h->stream>>>(q->a, h->d_rep, h->d_gpos);
        h->launches += 4;
    }
    k_qn_count<<<1, 256, 0, h->stream>>>(q->a.live, (int)R, q->a.nlive);
    MM_CUDA(cudaMemcpyAsync(q->h_nlive, q->a.nlive, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    MM_CUDA(cudaGetLastError());
    if (nlive_out) *nlive_out = *q->h_nlive;
    return MM_OK;
}

/* results (host arrays, any may be NULL): x [R][n], f [R], g [R][n], trust radius [R], conv_val [R]; int32 [R] each:
 * iterations, converged, failed, conv_count */
int mm_qn_get(mm_qn *q, double *x, double *f, double *g, double *radius, double *conv_val, int32_t *iterations, int32_t *converged,
              int32_t *failed, int32_t *conv_count, int64_t *evaluations) {
    if (!q) {
        set_error("mm_qn_get: null optimiser");
        return MM_ERR_INVALID;
    }
    mm_handle *h = q->h;
    MM_CUDA(cudaSetDevice(h->device));
    const size_t R = (size_t)q->nrep, n1 = R * q->a.n;
    auto get = [&](void *dst, const void *src, size_t bytes) { return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess; };
    MM_CUDA(get(x, q->a.x, sizeof(double) * n1));
    MM_CUDA(get(f, q->a.f, sizeof(double) * R));
    MM_CUDA(get(g, q->a.g, sizeof(double) * n1));
    MM_CUDA(get(radius, q->a.radius, sizeof(double) * R));
    MM_CUDA(get(conv_val, q->a.conv_val, sizeof(double) * R));
    MM_CUDA(get(iterations, q->a.iters, sizeof(int) * R));
    MM_CUDA(get(converged, q->a.converged, sizeof(int) * R));
    MM_CUDA(get(failed, q->a.failed, sizeof(int) * R));
    MM_CUDA(get(conv_count, q->a.conv_count, sizeof(int) * R));
    MM_CUDA(cudaStreamSynchronize(h->stream));
    if (evaluations) *evaluations = q->evaluations;
    return MM_OK;
}

}  // extern "C"
