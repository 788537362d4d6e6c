// binning.cu — K3 / K4: the GPU replacement of the reference's tile scheduler
// (renderer.go:380-406: per-tile local buffers flushed under a mutex).
//
// K2 counted, per device tile, the triangles whose first tile it is (A,
// positions already handed out, warp-aggregated) and the other triangle/tile
// pairs (B).  K3 turns A+B into list offsets with an exclusive prefix sum (one
// block per frame, warp shuffles + shared-memory staging) and re-zeroes the
// counters for the next draw.  K4 walks the records warp segment by warp
// segment and writes each triangle's slot to list[off + binPos] for its first
// tile (no atomic) and through an atomic cursor for the remaining tiles of the
// few triangles that straddle tile borders.  List order is arbitrary: the
// raster kernel resolves visibility with a (depth, submission order) key.

#include "gr_types.cuh"
#include "kernels.h"

namespace gr {

__global__ void __launch_bounds__(1024) bin_scan_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.x;
    const int nTiles = a.ntx * a.nty;
    uint32_t *cntA = a.tileCount + (size_t)frame * 2 * nTiles;
    uint32_t *cntB = cntA + nTiles;
    uint32_t *off = a.tileOff + (size_t)frame * (nTiles + 1);
    uint32_t *offB = a.tileOffB + (size_t)frame * nTiles;
    uint32_t *cur = a.cursor + (size_t)frame * nTiles;

    __shared__ uint32_t warpSum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nTiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t va = t < nTiles ? cntA[t] : 0u;
        const uint32_t v = va + (t < nTiles ? cntB[t] : 0u);
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
            offB[t] = excl + va;
            cntA[t] = 0;
            cntB[t] = 0;
            cur[t] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + warpSum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) off[nTiles] = carry;
}

__global__ void __launch_bounds__(kFaceBlock) bin_fill_kernel(const __grid_constant__ DrawArgs a) {
    const int frame = blockIdx.y;
    const int fb = blockIdx.x;
    const int o = a.fblkObj[fb];
    const DrawObj ob = a.objs[o];
    const FrameObj &fo = a.frameObjs[(size_t)frame * a.nobj + o];
    if (fo.visibility == GRB_BOX_OUTSIDE) return;
    const bool clips = (a.options & GRB_OPT_FRUSTUM_CLIPPING) && fo.visibility != GRB_BOX_INSIDE;
    const uint32_t perWarp = clips ? kWarpSlotsClip : kWarpSlots;
    const unsigned lane = threadIdx.x & 31u, warpInBlock = threadIdx.x >> 5;
    const uint32_t cnt = a.warpCount[((size_t)frame * a.nFaceBlocks + fb) * kWarpsPerFaceBlock + warpInBlock];
    const uint32_t slot0 = fo.slotBase + ((uint32_t)(fb - ob.faceBlockBase) * kWarpsPerFaceBlock + warpInBlock) * perWarp;

    const int nTiles = a.ntx * a.nty;
    const TriRec *rec = a.rec + (size_t)frame * a.recCap;
    const uint32_t *off = a.tileOff + (size_t)frame * (nTiles + 1);
    const uint32_t *offB = a.tileOffB + (size_t)frame * nTiles;
    uint32_t *cur = a.cursor + (size_t)frame * nTiles;
    uint32_t *list = a.binList + (size_t)frame * a.recCap * kMaxBinsPerTri;

    for (uint32_t k = lane; k < cnt; k += 32) {
        const uint32_t slot = slot0 + k;
        // bbox, texture id and binPos live in the record's 4th 16-byte quarter
        const int4 q = __ldg(reinterpret_cast<const int4 *>(rec + slot) + 3);
        const int bx0 = (int16_t)(q.x & 0xffff), by0 = (int16_t)(q.x >> 16);
        const int bx1 = (int16_t)(q.y & 0xffff), by1 = (int16_t)(q.y >> 16);
        if (bx1 < bx0) continue;  // survived the cull but draws nothing (setup.cu)
        const int tx0 = bx0 / kTile, tx1 = bx1 / kTile, ty0 = by0 / kTile, ty1 = by1 / kTile;
        if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > kMaxBinsPerTri) continue;  // in bigList
        const int t0 = ty0 * a.ntx + tx0;
        list[off[t0] + (uint32_t)q.w] = slot;
        for (int ty = ty0; ty <= ty1; ty++)
            for (int tx = tx0; tx <= tx1; tx++) {
                const int t = ty * a.ntx + tx;
                if (t == t0) continue;
                list[offB[t] + atomicAdd(&cur[t], 1u)] = slot;
            }
    }
}

void launch_bin_scan(const DrawArgs &a, int nframes, cudaStream_t s) {
    bin_scan_kernel<<<nframes, 1024, 0, s>>>(a);
}

void launch_bin_fill(const DrawArgs &a, int nframes, cudaStream_t s) {
    if (a.nFaceBlocks == 0) return;
    bin_fill_kernel<<<dim3(a.nFaceBlocks, nframes), kFaceBlock, 0, s>>>(a);
}

}  // namespace gr
