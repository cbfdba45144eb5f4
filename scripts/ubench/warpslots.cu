// Micro-benchmark: which hardware warp slots (%warpid) do the warps of co-resident CTAs get, and how does the
// FP64 pipe time of a warp depend on its slot?  (developer tool, not part of the library)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, long long *cyc, double a, double b, int n) {
    extern __shared__ double sm[];
    unsigned smid, wslot;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wslot));
    double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
        int w = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
        out[3 * w] = smid; out[3 * w + 1] = wslot; cyc[w] = t1 - t0;
    }
    if (x0 + x1 + x2 + x3 == 1.2345) sm[0] = x0;
}
int main(int argc, char **argv) {
    int threads = argc > 1 ? atoi(argv[1]) : 192, ctas_per_sm = argc > 2 ? atoi(argv[2]) : 2;
    int *o; long long *c;
    int nsm = 148, grid = nsm * ctas_per_sm, nw = grid * threads / 32;
    cudaMalloc(&o, nw * 12); cudaMalloc(&c, nw * 8);
    size_t smem = (220 * 1024) / ctas_per_sm - 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) k<<<grid, threads, smem>>>(o, c, 1.000001, 0.999999, 4096);
    cudaDeviceSynchronize();
    int *ho = new int[nw * 3]; long long *hc = new long long[nw];
    cudaMemcpy(ho, o, nw * 12, cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, nw * 8, cudaMemcpyDeviceToHost);
    printf("threads %d, %d CTAs/SM: warps on SM 0 and SM 77 (cta, warp, hw slot, slot%%4, cycles per 4 DFMA)\n", threads, ctas_per_sm);
    for (int w = 0; w < nw; ++w)
        if (ho[3 * w] == 0 || ho[3 * w] == 77)
            printf("  sm %3d cta %4d warp %d slot %2d sched %d  %.2f\n", ho[3 * w], w / (threads / 32), w % (threads / 32), ho[3 * w + 1], ho[3 * w + 1] % 4, hc[w] / 4096.0);
    return 0;
}
