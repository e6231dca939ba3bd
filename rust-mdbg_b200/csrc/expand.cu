// expand.cu -- device side of the 4:1 upload: bit planes (pack_host.cc) -> ASCII bases in HBM.
// One thread per 32-base word: 8 bytes read, 32 bytes written with two 128-bit stores; pure streaming
// (roofline: HBM, 1.25 B per base).  The kernels downstream see exactly the bytes the caller passed
// to mdbg_push_reads (tiles with a byte outside ACGT are never packed: they are copied as ASCII).
#include <cuda_runtime.h>
#include <stdint.h>

#include "expand_math.h"
#include "mdbg_kernels.h"

namespace mdbg {

__global__ void expand_planes_kernel(const uint2* __restrict__ planes, uint8_t* __restrict__ bases,
                                     uint64_t w_begin, uint64_t w_end) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w = w_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < w_end; w += stride) {
        const uint2 p = __ldg(planes + w);
        uint4 lo, hi;
        lo.x = expand4(p.x, p.y); lo.y = expand4(p.x >> 4, p.y >> 4);
        lo.z = expand4(p.x >> 8, p.y >> 8); lo.w = expand4(p.x >> 12, p.y >> 12);
        hi.x = expand4(p.x >> 16, p.y >> 16); hi.y = expand4(p.x >> 20, p.y >> 20);
        hi.z = expand4(p.x >> 24, p.y >> 24); hi.w = expand4(p.x >> 28, p.y >> 28);
        uint4* dst = reinterpret_cast<uint4*>(bases + w * 32);
        dst[0] = lo;
        dst[1] = hi;
    }
}

// words [w_begin, w_end) of the batch; bases must be 16-byte aligned and hold 32 * w_end bytes
cudaError_t expand_planes(const uint32_t* planes, uint8_t* bases, uint64_t w_begin, uint64_t w_end, int num_sms,
                          cudaStream_t st, uint64_t* launches) {
    if (w_end <= w_begin) return cudaSuccess;
    const uint64_t n = w_end - w_begin;
    const uint64_t want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < (uint64_t)num_sms * 8 ? want : (uint64_t)num_sms * 8);
    expand_planes_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint2*>(planes), bases, w_begin, w_end);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace mdbg
