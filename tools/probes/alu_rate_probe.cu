// Integer-pipe rate probe: how many LOP3 / IADD3 / SHF / IMAD lane-operations per clock does one SM sustain?
// (SURVEY.md 8(d) asks for the sm_100 INT32 rate to be measured rather than assumed: 64 or 128 lanes/clk/SM.)
// One block per SM, 32 warps per block, 8 independent dependency chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate_probe alu_rate_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) k_rate(uint32_t* out, long long* cyc, int iters, uint32_t seed)
{
    uint32_t a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = seed + threadIdx.x * 8 + k;
    const uint32_t b = seed * 2654435761u + 12345u, c = seed ^ 0x5bd1e995u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (OP == 0) a[k] = (a[k] & a[(k + 1) & 7]) ^ a[(k + 3) & 7];        // one LOP3 (three live registers: cannot be folded)
                else if (OP == 1) a[k] = a[k] + a[(k + 1) & 7] + a[(k + 3) & 7];     // one IADD3
                else if (OP == 2) a[k] = __funnelshift_r(a[k], b, 7);                // one SHF
                else a[k] = a[k] * b + c;                                            // one IMAD (FMA pipe)
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 1024);
    cudaMallocManaged(&cyc, sizeof(long long) * sms);
    const int iters = 4096;
    const char* names[4] = { "LOP3", "IADD3", "SHF", "IMAD" };
    for (int op = 0; op < 4; op++) {
        for (int rep = 0; rep < 2; rep++) {
            if (op == 0) k_rate<0><<<sms, 1024>>>(out, cyc, iters, 7u + rep);
            if (op == 1) k_rate<1><<<sms, 1024>>>(out, cyc, iters, 7u + rep);
            if (op == 2) k_rate<2><<<sms, 1024>>>(out, cyc, iters, 7u + rep);
            if (op == 3) k_rate<3><<<sms, 1024>>>(out, cyc, iters, 7u + rep);
            cudaDeviceSynchronize();
        }
        long long worst = 0;
        for (int i = 0; i < sms; i++) worst = cyc[i] > worst ? cyc[i] : worst;
        const double lane_ops = double(iters) * 64.0 * 1024.0;      // per SM
        printf("%-5s %.1f lane-ops/clk/SM (%d SMs, %lld cycles, %s)\n", names[op], lane_ops / double(worst), sms, worst,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
