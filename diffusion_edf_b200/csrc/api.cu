// Library-level entry points: build identification and the L2 weight prefetch.
#include "common.cuh"
#include "../../include/dedf.h"

namespace dedf {

// One CTA per tensor (grid-stride over the table): every thread issues L2 prefetches for 128-byte lines.
__global__ void __launch_bounds__(256) prefetch_l2_kernel(const void* const* __restrict__ ptrs, const long long* __restrict__ bytes, int n) {
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        const char* p = static_cast<const char*>(ptrs[t]);
        const long long nb = bytes[t];
        for (long long o = (long long)threadIdx.x * 128; o < nb; o += 256ll * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
    }
}

}  // namespace dedf

extern "C" int dedf_build_arch(void) { return 100; }

extern "C" int dedf_prefetch_l2(const void* const* ptrs_dev, const long long* bytes_dev, int n, cudaStream_t stream) {
    if (n <= 0) return dedf::DEDF_OK;
    if (!ptrs_dev || !bytes_dev) return dedf::DEDF_ERR_ARG;
    dedf::prefetch_l2_kernel<<<dedf::grid_for(n, 1, dedf::kNumSMs * 4), 256, 0, stream>>>(ptrs_dev, bytes_dev, n);
    if (cudaGetLastError() != cudaSuccess) return dedf::DEDF_ERR_LAUNCH;
    return dedf::DEDF_OK;
}
