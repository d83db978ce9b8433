// Microbenchmarks behind the k_march design decisions (B200, sm_100a): FP64 latency / issue rate per scheduler,
// shuffle / shared-memory instruction rates per SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe fp64_mio_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <int MODE>  // 0 SHFL.32, 1 LDS.64, 2 STS.64, 3 LDS.64 dependent-address chain (latency), 4 SHFL dependent (latency)
__global__ void k_mio(double *out, int iters) {
    __shared__ double sm[8][16][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = 0; i < 16; i++) sm[w % 8][i][lane] = lane + i;
    __syncthreads();
    double acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    int ia = lane;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                int v = __double2loint(acc[i]);
                v = __shfl_down_sync(0xffffffffu, v, 1);
                acc[i] = __hiloint2double(__double2hiint(acc[i]), v);
            }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] += sm[w % 8][i + (it & 1)][lane];
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) sm[w % 8][i + (it & 1)][lane] = acc[i];
        } else if (MODE == 3) {
            ia = (int)sm[w % 8][ia & 15][ia & 31] & 31;
        } else {
            ia = __shfl_down_sync(0xffffffffu, ia, 1);
        }
    }
    long long t1 = clock64();
    double s = ia;
    for (int i = 0; i < 8; i++) s += acc[i];
    if (MODE == 2) s += sm[w % 8][3][lane];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <typename F>
static double run(F f, int threads, double *d) {
    f(threads);
    cudaDeviceSynchronize();
    f(threads);
    cudaDeviceSynchronize();
    double c;
    cudaMemcpy(&c, d + 148 * threads, 8, cudaMemcpyDeviceToHost);
    return c;
}

int main() {
    double *d;
    cudaMalloc(&d, 8 * (148 * 1024 + 16));
    const int iters = 2000;
    printf("DFMA: cycles per warp-instruction per scheduler (ILP chains per thread x warps per scheduler)\n");
    for (int wps = 1; wps <= 4; wps *= 2) {
        const int threads = 128 * wps;
#define ROW(ILP) { double c = run([&](int t) { k_dfma<ILP><<<148, t>>>(d, iters, 1.0000001, 1e-9); }, threads, d); \
        printf("  warps/sched %d ILP %d: %.2f cycles per dependent step, %.3f cycles/instr/sched\n", wps, ILP, c / (iters * 8.0), c / (iters * 8.0 * ILP * wps)); }
        ROW(1) ROW(2) ROW(3) ROW(4) ROW(6) ROW(8)
    }
    printf("MIO: 8 warps per SM, cycles per warp-instruction per SM\n");
    const char *names[5] = {"SHFL.32 throughput", "LDS.64 throughput", "STS.64 throughput", "LDS dependent latency", "SHFL dependent latency"};
    for (int m = 0; m < 5; m++) {
        double c = 0;
        if (m == 0) c = run([&](int t) { k_mio<0><<<148, t>>>(d, iters); }, 256, d);
        if (m == 1) c = run([&](int t) { k_mio<1><<<148, t>>>(d, iters); }, 256, d);
        if (m == 2) c = run([&](int t) { k_mio<2><<<148, t>>>(d, iters); }, 256, d);
        if (m == 3) c = run([&](int t) { k_mio<3><<<148, t>>>(d, iters); }, 32, d);
        if (m == 4) c = run([&](int t) { k_mio<4><<<148, t>>>(d, iters); }, 32, d);
        if (m < 3) printf("  %s: %.2f cycles per warp-instr per SM (8 warps x 8 instr per iteration)\n", names[m], c / (iters * 64.0));
        else printf("  %s: %.1f cycles\n", names[m], c / iters);
    }
    return 0;
}
