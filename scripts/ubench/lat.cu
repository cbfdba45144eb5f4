// Micro-benchmark: dependent-chain latencies of fp64 ops on one warp (developer tool, not part of the library).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b, int n) {
    __shared__ double sm[64];
    sm[threadIdx.x] = a + threadIdx.x;
    __syncthreads();
    double x = a + threadIdx.x * 1e-3;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    double y = x;
    for (int i = 0; i < n; ++i) y = 1.0 / (y + a);
    long long t2 = clock64();
    double z = y;
    for (int i = 0; i < n; ++i) z = sqrt(z + a);
    long long t3 = clock64();
    double w = z;
    for (int i = 0; i < n; ++i) w = rsqrt(w + a);
    long long t4 = clock64();
    double u = w;
    for (int i = 0; i < n; ++i) u = a / (u + b);
    long long t5 = clock64();
    // 4 independent chains (ILP 4)
    double p0 = u, p1 = u + 1, p2 = u + 2, p3 = u + 3;
    for (int i = 0; i < n; ++i) { p0 = fma(p0, b, a); p1 = fma(p1, b, a); p2 = fma(p2, b, a); p3 = fma(p3, b, a); }
    long long t6 = clock64();
    int idx = threadIdx.x;
    double q = 0;
    for (int i = 0; i < n; ++i) { q += sm[idx]; idx = (idx + (int)q) & 31; }
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6;
    }
    out[threadIdx.x] = x + y + z + w + u + p0 + p1 + p2 + p3 + q;
}
int main() {
    double *o; long long *c, h[7];
    cudaMalloc(&o, 1024); cudaMalloc(&c, 64);
    const int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k<<<1, 32>>>(o, c, 1.000001, 0.999999, n);
        cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
    }
    const char *nm[7] = {"dfma dependent", "1.0/x", "sqrt", "rsqrt", "a/x", "dfma x4 independent (per 4)", "LDS dependent (+dadd+iadd)"};
    for (int i = 0; i < 7; ++i) printf("%-30s %.1f cycles/iter\n", nm[i], (double)h[i] / n);
    return 0;
}
