// Latency probe for the sequential FP64 sum of k_finish's re-score: how fast can ONE warp run a chain of dependent
// double adds when (a) the addends are already in registers, (b) they are read from shared memory, (c) they are
// selected from four register weights by 2-bit codes.   nvcc -gencode arch=compute_100a,code=sm_100a -o dadd_probe dadd_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_chain_regs(double* out, long long* cyc, double a, double b, int n)
{
    double t = out[threadIdx.x];
    const long long c0 = clock64();
    for (int i = 0; i < n; i += 8) {
        t += a; t += b; t += a; t += b; t += a; t += b; t += a; t += b;
    }
    const long long c1 = clock64();
    out[threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[0] = c1 - c0;
}

__global__ void k_chain_smem(double* out, long long* cyc, const double* w, int n)
{
    __shared__ double s[32][33];
    for (int i = threadIdx.x; i < 32 * 32; i += 32) s[i / 32][i % 32] = w[i % 4];
    __syncwarp();
    double t = out[threadIdx.x];
    const long long c0 = clock64();
    for (int i = 0; i < n; i += 32) {
#pragma unroll
        for (int u = 0; u < 32; u++) t += s[u][threadIdx.x];
    }
    const long long c1 = clock64();
    out[threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[1] = c1 - c0;
}

__global__ void k_chain_select(double* out, long long* cyc, const double* w, const unsigned long long* packs, int n)
{
    const double w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    double t = out[threadIdx.x];
    const long long c0 = clock64();
    for (int i = 0; i < n; i += 32) {
        const unsigned long long pack = packs[(i / 32) * 32 + threadIdx.x];
#pragma unroll
        for (int u = 0; u < 32; u++) {
            const uint32_t c = uint32_t(pack >> (2 * u)) & 3u;
            const double lo = (c & 1u) ? w1 : w0, hi = (c & 1u) ? w3 : w2;
            t += (c & 2u) ? hi : lo;
        }
    }
    const long long c1 = clock64();
    out[threadIdx.x] = t;
    if (threadIdx.x == 0) cyc[2] = c1 - c0;
}

int main()
{
    const int n = 2048;
    double *out, *w; long long* cyc; unsigned long long* packs;
    cudaMallocManaged(&out, 32 * sizeof(double));
    cudaMallocManaged(&w, 4 * sizeof(double));
    cudaMallocManaged(&cyc, 4 * sizeof(long long));
    cudaMallocManaged(&packs, (n / 32) * 32 * sizeof(unsigned long long));
    w[0] = 2; w[1] = -1.5; w[2] = -1.1; w[3] = -1.3;
    for (int i = 0; i < (n / 32) * 32; i++) packs[i] = 0x9E3779B97F4A7C15ull * (i + 1);
    for (int rep = 0; rep < 2; rep++) {
        for (int i = 0; i < 32; i++) out[i] = 0.1 * i;
        k_chain_regs<<<1, 32>>>(out, cyc, 1.1, -1.3, n);
        k_chain_smem<<<1, 32>>>(out, cyc, w, n);
        k_chain_select<<<1, 32>>>(out, cyc, w, packs, n);
        cudaDeviceSynchronize();
        printf("steps %d: regs %.1f cyc/step, smem %.1f cyc/step, select %.1f cyc/step  (%s)\n", n, double(cyc[0]) / n, double(cyc[1]) / n,
               double(cyc[2]) / n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
