// Which TMA store boxes work for fp64 planes?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_store_probe.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int n) {
    extern __shared__ __align__(128) double s[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = 1000.0 + i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                         reinterpret_cast<unsigned long long>(&map)), "r"((unsigned)__cvta_generic_to_shared(s)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int nxp = 258, nyp = 258, nz = 4;
    double *d; cudaMalloc(&d, 8ull * nxp * nyp * nz); 
    const cuuint64_t dims[3] = {nxp, nyp, nz}, strides[2] = {nxp * 8ull, 8ull * nxp * nyp};
    const cuuint32_t es[3] = {1, 1, 1};
    struct { int bx, by, c0, c1; } cases[] = {{32, 8, 0, 0}, {32, 8, 1, 1}, {30, 6, 0, 0}, {30, 6, 1, 1}, {30, 6, 2, 1}, {30, 6, 31, 7}};
    for (auto &c : cases) {
        CUtensorMap m; const cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaMemset(d, 0, 8ull * nxp * nyp * nz);
        k<<<1, 128, 8 * c.bx * c.by>>>(m, c.c0, c.c1, 1, c.bx * c.by);
        cudaError_t e = cudaDeviceSynchronize();
        double h[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) {
            cudaMemcpy(&h[0], d + (1ull * nyp + c.c1) * nxp + c.c0, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(&h[1], d + (1ull * nyp + c.c1) * nxp + c.c0 + c.bx - 1, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(&h[2], d + (1ull * nyp + c.c1 + 1) * nxp + c.c0, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(&h[3], d + (1ull * nyp + c.c1) * nxp + c.c0 + c.bx, 8, cudaMemcpyDeviceToHost);
        }
        printf("box %dx%d at (%d,%d): encode %d, run: %s; first %.0f last-in-row %.0f next-row %.0f beyond %.0f\n", c.bx, c.by, c.c0, c.c1, (int)r,
               cudaGetErrorString(e), h[0], h[1], h[2], h[3]);
        if (e != cudaSuccess) { printf("(context lost, stopping)\n"); return 0; }
    }
    return 0;
}
