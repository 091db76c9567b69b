// psa_scan.cu -- bit-sliced scan engine (placeholder until the kernel lands; engine 2 is not selectable yet)
#include "psa_kernels.cuh"

namespace psa {

int scan_chunk_steps(int, int64_t) { return 512; }
size_t scan_smem_bytes(int, int) { return 0; }
void launch_profile(const DeviceTable&, const BatchGeom&, const BatchPtrs&, int, int, cudaStream_t) {}
void launch_scan(const DeviceTable&, const BatchGeom&, const BatchPtrs&, int, int64_t, cudaStream_t) {}

} // namespace psa
