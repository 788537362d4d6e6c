// binning.cu — K3 / K4: the GPU replacement of the reference's tile scheduler
// (renderer.go:380-406: per-tile local buffers flushed under a mutex).
//
// K2 counted, per device tile, how many emitted triangles touch it.  K3 turns
// the counts into list offsets with an exclusive prefix sum (one block per
// frame, shared-memory staging) and re-zeroes the counters for the next draw;
// K4 scatters the triangle slots into the lists with one atomic cursor per
// tile.  List order is arbitrary: the raster kernel resolves visibility with
// a (depth, submission order) key, so no ordering is needed here.

#include "gr_types.cuh"
#include "kernels.h"

namespace gr {

__global__ void __launch_bounds__(1024) bin_scan_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.x;
    const int nTiles = a.ntx * a.nty;
    uint32_t *cnt = a.tileCount + (size_t)frame * nTiles;
    uint32_t *off = a.tileOff + (size_t)frame * (nTiles + 1);
    uint32_t *cur = a.cursor + (size_t)frame * nTiles;

    __shared__ uint32_t warpSum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nTiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t v = t < nTiles ? cnt[t] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += x;
        }
        if (lane == 31) warpSum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = warpSum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += x;
            }
            warpSum[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t c = carry;
        const uint32_t excl = c + (wid ? warpSum[wid - 1] : 0u) + inc - v;
        if (t < nTiles) {
            off[t] = excl;
            cnt[t] = 0;
            cur[t] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warpSum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nTiles] = carry;
}

__global__ void __launch_bounds__(256) bin_fill_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.y;
    const int nTiles = a.ntx * a.nty;
    const uint32_t n = a.counters[frame].triCount;
    const TriRec *rec = a.rec + (size_t)frame * a.recCap;
    const uint32_t *off = a.tileOff + (size_t)frame * (nTiles + 1);
    uint32_t *cur = a.cursor + (size_t)frame * nTiles;
    uint32_t *list = a.binList + (size_t)frame * a.recCap * kMaxBinsPerTri;

    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // the bbox lives in the record's 4th 16-byte quarter
        const int4 q = __ldg(reinterpret_cast<const int4 *>(rec + i) + 3);
        const int bx0 = (int16_t)(q.x & 0xffff), by0 = (int16_t)(q.x >> 16);
        const int bx1 = (int16_t)(q.y & 0xffff), by1 = (int16_t)(q.y >> 16);
        const int tx0 = bx0 / kTile, tx1 = bx1 / kTile, ty0 = by0 / kTile, ty1 = by1 / kTile;
        if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > kMaxBinsPerTri) continue;  // in bigList
        for (int ty = ty0; ty <= ty1; ty++)
            for (int tx = tx0; tx <= tx1; tx++) {
                const int t = ty * a.ntx + tx;
                const uint32_t pos = atomicAdd(&cur[t], 1u);
                list[off[t] + pos] = i;
            }
    }
}

void launch_bin_scan(const DrawArgs &a, int nframes, cudaStream_t s) {
    bin_scan_kernel<<<nframes, 1024, 0, s>>>(a);
}

void launch_bin_fill(const DrawArgs &a, int nframes, uint32_t maxTris, cudaStream_t s) {
    if (maxTris == 0) return;
    uint32_t blocks = (maxTris + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bin_fill_kernel<<<dim3(blocks, nframes), 256, 0, s>>>(a);
}

}  // namespace gr
