// Launch-overhead probe: what does an event bracket around ONE launch cost beyond the time the blocks themselves run?
// Each block spins for `spin_ns` (globaltimer) and records its start and end; the host compares
//   events   : cudaEventElapsedTime around the launch (what bench.py's `value` and psa_batch_run report)
//   span     : last block end - first block start (globaltimer)
//   busy     : the spin itself
// for several launch shapes: block size, dynamic shared memory, parameter bytes, ordinary / cooperative launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_probe launch_probe.cu
#include <algorithm>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda.h>

struct Params { unsigned char bytes[480]; };

__global__ void k_spin(unsigned long long* t, unsigned long long spin_ns, Params p)
{
    extern __shared__ unsigned char smem[];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x == 0) smem[0] = p.bytes[blockIdx.x % 480];
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
    __syncthreads();
    if (threadIdx.x == 0) { t[2 * blockIdx.x] = t0; t[2 * blockIdx.x + 1] = t1 + smem[0] * 0; }
}

__global__ void k_small(unsigned long long* t, unsigned long long spin_ns)
{
    extern __shared__ unsigned char smem[];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
    __syncthreads();
    if (threadIdx.x == 0) { t[2 * blockIdx.x] = t0; t[2 * blockIdx.x + 1] = t1; }
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long* d_t;
    cudaMalloc(&d_t, sizeof(unsigned long long) * 2 * 1024);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    Params p{};
    std::vector<unsigned long long> h(2 * 1024);
    unsigned char* flush;
    cudaMalloc(&flush, 512u << 20);
    // gate: 0 none (the GPU is idle when the host starts to enqueue: host launch latency sits inside the bracket),
    //       1 a 60 us spin kernel in front of the first event, 2 a stream wait on a page-locked word the host writes after
    //       it has enqueued event + launch + event (cuStreamWaitValue32)
    typedef CUresult (*WaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
    WaitFn wait_value = nullptr;
    {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult qr{};
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fp, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) wait_value = (WaitFn)fp;
    }
    volatile unsigned int* h_gate = nullptr;
    cudaHostAlloc((void**)&h_gate, 64, cudaHostAllocMapped);
    *h_gate = 0;
    unsigned int gate_tag = 0;
    struct Shape { const char* name; int threads; size_t smem; bool big_params; bool coop; bool do_flush; int gate; };
    const Shape shapes[] = {
        { "32 thr, 0 smem, 16 B params", 32, 0, false, false, false, 0 },
        { "640 thr, 0 smem, 16 B params", 640, 0, false, false, false, 0 },
        { "640 thr, 210 KB smem, 16 B params", 640, 210 * 1024, false, false, false, 0 },
        { "640 thr, 210 KB smem, 496 B params", 640, 210 * 1024, true, false, false, 0 },
        { "640 thr, 210 KB smem, 496 B params, after a 512 MB memset", 640, 210 * 1024, true, false, true, 0 },
        { "256 thr, 60 KB smem, 496 B params, cooperative", 256, 60 * 1024, true, true, false, 0 },
        { "256 thr, 60 KB smem, 496 B params, ordinary", 256, 60 * 1024, true, false, false, 0 },
        { "640 thr, 210 KB smem, 496 B params, behind a spin kernel", 640, 210 * 1024, true, false, false, 1 },
        { "640 thr, 210 KB smem, 496 B params, behind a host-released gate", 640, 210 * 1024, true, false, false, 2 },
        { "256 thr, 60 KB smem, cooperative, behind a host-released gate", 256, 60 * 1024, true, true, false, 2 },
        { "32 thr, 0 smem, 16 B params, behind a host-released gate", 32, 0, false, false, false, 2 },
    };
    const unsigned long long spin_ns = 20000;
    for (const Shape& s : shapes) {
        std::vector<float> ev, span;
        for (int rep = 0; rep < 25; rep++) {
            if (s.do_flush) cudaMemsetAsync(flush, rep, 512u << 20, st);
            cudaStreamSynchronize(st);
            if (s.gate == 1) k_small<<<1, 32, 0, st>>>(d_t + 2 * 512, 60000);
            if (s.gate == 2 && wait_value) {
                void* dv = nullptr;
                cudaHostGetDevicePointer(&dv, (void*)h_gate, 0);
                wait_value((CUstream)st, (CUdeviceptr)dv, ++gate_tag, CU_STREAM_WAIT_VALUE_EQ);
            }
            cudaEventRecord(e0, st);
            if (s.big_params) {
                if (s.coop) {
                    unsigned long long sp = spin_ns;
                    void* args[] = { (void*)&d_t, (void*)&sp, (void*)&p };
                    cudaLaunchCooperativeKernel((const void*)k_spin, dim3(sms), dim3(s.threads), args, s.smem, st);
                } else
                    k_spin<<<sms, s.threads, s.smem, st>>>(d_t, spin_ns, p);
            } else
                k_small<<<sms, s.threads, s.smem, st>>>(d_t, spin_ns);
            cudaEventRecord(e1, st);
            if (s.gate == 2) *h_gate = gate_tag;
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(h.data(), d_t, sizeof(unsigned long long) * 2 * sms, cudaMemcpyDeviceToHost);
            unsigned long long lo = ~0ull, hi = 0;
            for (int b = 0; b < sms; b++) { lo = std::min(lo, h[2 * b]); hi = std::max(hi, h[2 * b + 1]); }
            if (rep >= 5) { ev.push_back(ms * 1e3f); span.push_back(float(hi - lo) * 1e-3f); }
        }
        std::sort(ev.begin(), ev.end()); std::sort(span.begin(), span.end());
        std::printf("%-62s events %6.2f us  blocks' span %6.2f us  (spin %.1f us)  overhead %5.2f us  err=%s\n", s.name, ev[ev.size() / 2],
                    span[span.size() / 2], spin_ns * 1e-3, ev[ev.size() / 2] - span[span.size() / 2], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
