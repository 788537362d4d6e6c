// present.cu — mirrors of the framebuffer (host memory, or another GPU's framebuffer), kept in sync tile by tile.
//
// The reference's FrameBuffer lives in host memory (`Pixels`, `ZBuffer`, rasterizer.go:7-13) and every
// Draw starts by clearing all of it (rasterizer.go:36-52).  Here the framebuffer lives in HBM, and
// copying all of it to the caller after every frame (7.4 MB at 1280x720) makes the path PCIe-bound
// while most of a frame is the cleared background the host already holds from the frame before.
// A mirror is a pinned, device-mapped host plane (colour or depth) plus one byte per 32x32 tile on
// the device: "this tile of the HOST copy is not the cleared background".  The raster kernel leaves one
// byte per tile of every frame it renders: "this tile of the DEVICE frame is not the cleared background"
// (DrawArgs::tileBusy).  Updating a mirror from a frame writes exactly the tiles that are busy now or
// were busy in the host copy — straight into host memory with 128-bit stores over PCIe, no staging
// buffer and no host-side scatter — and leaves the rest alone: background is a function of (x, y)
// only (Clear + DotGrid), so the host plane ends up byte for byte what a full copy would have produced.
// Measured on the box (scripts/probes/pcie_probe.cu): 46 GB/s for C3's ~130 busy tiles per frame against
// 57 GB/s for the full-frame DMA that moves 7x the bytes; a compacted DMA plus a host scatter loop
// stops at 5 GB/s per host core.

#include <algorithm>
#include <cstdlib>

#include "gr_types.cuh"
#include "kernels.h"

namespace gr {

// A small fixed grid walks the frames' tile rows in groups of 4 horizontally adjacent tiles (128 pixels): lane l of a
// warp owns the 4 pixels 4l .. 4l+3 of a row of the group, so that a warp's stores are one contiguous 512-byte run when
// neighbouring tiles are written together — the usual case inside an object — instead of four separate 128-byte tile rows
// (measured on the box, scripts/probes/pcie_probe.cu: 52 GB/s against 46.7 GB/s over PCIe; both planes move as 128-bit
// accesses).  Lanes of tiles that are skipped are predicated off.  The grid is kept small on purpose: the kernel is
// bound by PCIe (or NVLink), runs on the high-priority copy stream beside the render kernels of the next batch, and
// one block per tile would fill every SM's thread slots with blocks that do nothing but wait for the bus — the update
// would then serialise with rendering instead of hiding behind it.
constexpr int kMirrorBlocksPerSM = 3;   // measured 1..16 on C3 beside rendering (scripts/mirror_tune.py): 3 is best by a few per cent
constexpr int kGroupTiles = 4;

__global__ void __launch_bounds__(256) mirror_update_kernel(const MirrorArgs m, const int nframes) {
    const int nTiles = m.ntx * m.nty;
    const int groupsX = (m.ntx + kGroupTiles - 1) / kGroupTiles, perFrame = groupsX * m.tileRows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int idx = blockIdx.x; idx < nframes * perFrame; idx += gridDim.x) {
        const int frame = idx / perFrame, g = idx - frame * perFrame;
        const int ty = m.tileRow0 + g / groupsX, tx = (g % groupsX) * kGroupTiles + (lane >> 3);   // this lane's tile
        const bool inRow = tx < m.ntx;
        const size_t flag = (size_t)frame * nTiles + (size_t)ty * m.ntx + tx;
        const uint8_t busy = !inRow ? 0 : (m.full ? 1 : m.tileBusy[flag]);
        const bool doC = inRow && m.dirtyColor && (busy | m.dirtyColor[flag]);
        const bool doZ = inRow && m.dirtyDepth && (busy | m.dirtyDepth[flag]);
        // block-uniform: every tile of the group is background in the mirror and stays background
        if (!__syncthreads_or(doC || doZ)) continue;
        const int gx = tx * kTile + (lane & 7) * 4;
        if (gx < m.width && (doC || doZ)) {
#pragma unroll
            for (int r = 0; r < kTile / 8; r++) {
                const int gy = ty * kTile + warp + 8 * r;
                if (gy >= m.height) break;
                const size_t pix = ((size_t)frame * m.height + gy) * m.width + gx;
                if ((m.width & 3) == 0) {
                    if (doC) *reinterpret_cast<uint4 *>(m.hostColor + pix) = *reinterpret_cast<const uint4 *>(m.color + pix);
                    if (doZ) *reinterpret_cast<float4 *>(m.hostDepth + pix) = *reinterpret_cast<const float4 *>(m.depth + pix);
                } else {
                    for (int k = 0; k < 4 && gx + k < m.width; k++) {
                        if (doC) m.hostColor[pix + k] = m.color[pix + k];
                        if (doZ) m.hostDepth[pix + k] = m.depth[pix + k];
                    }
                }
            }
        }
        __syncthreads();   // every warp has read the flags of the group's tiles
        if (warp == 0 && (lane & 7) == 0 && (doC || doZ)) {
            // one lane per tile; nobody else touches these flags in this launch
            if (doC) m.dirtyColor[flag] = busy;
            if (doZ) m.dirtyDepth[flag] = busy;
            if (m.targetBusy) m.targetBusy[flag] = busy;
            if (doC && m.tilesWrittenColor) atomicAdd(m.tilesWrittenColor, 1ull);
            if (doZ && m.tilesWrittenDepth) atomicAdd(m.tilesWrittenDepth, 1ull);
        }
    }
}

void launch_mirror_update(const MirrorArgs &m, int nframes, cudaStream_t s) {
    if (nframes <= 0 || m.ntx <= 0 || m.tileRows <= 0) return;
    static const int sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    const long long total = (long long)nframes * ((m.ntx + kGroupTiles - 1) / kGroupTiles) * m.tileRows;
    static const int perSM = [] {   // tuning knob (scripts/mirror_tune.py)
        const char *e = getenv("GRB_MIRROR_BLOCKS_PER_SM");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : kMirrorBlocksPerSM;
    }();
    const int grid = (int)std::min<long long>(total, (long long)sms * perSM);
    mirror_update_kernel<<<grid, 256, 0, s>>>(m, nframes);
}

// ---- cross-process hand-off flags of a shared framebuffer (sort-first strips, parallel.py) -------------
//
// The same kernel is the exchange step of the sort-first strips: the mirror of a rank's local framebuffer is then
// rank 0's framebuffer (peer memory mapped through CUDA IPC), the launch covers the rank's rows only, and the
// tiles cross NVLink instead of PCIe — on the copy stream, while the render stream is already setting up the next
// frames, so that rank 0's NVLink ingress (where all strips converge) is busy all the time and not only at the
// end of every raster kernel.  Afterwards the rank raises a flag that lives next to that framebuffer;
// rank 0's stream waits on the flags on the device, the host is not involved.  A signal is queued behind
// the kernels whose writes it publishes; the fence orders those (complete at the kernel boundary) before
// the flag at system scope.  The wait gives up after `timeoutNs` and counts the failure instead of
// hanging the GPU when a peer has died.
__global__ void signal_kernel(volatile uint32_t *flag, uint32_t value) {
    __threadfence_system();
    *flag = value;
    __threadfence_system();
}

__global__ void wait_signals_kernel(volatile uint32_t *flags, int stride, int n, uint32_t value, unsigned long long timeoutNs,
                                    uint32_t *timeouts) {
    const int i = threadIdx.x;
    if (i >= n) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    // flags only grow; signed distance so that a counter that wrapped still compares
    while ((int32_t)(flags[(size_t)i * stride] - value) < 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeoutNs) {
            atomicAdd(timeouts, 1u);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

void launch_signal(uint32_t *flag, uint32_t value, cudaStream_t s) { signal_kernel<<<1, 1, 0, s>>>(flag, value); }

void launch_wait_signals(uint32_t *flags, int strideWords, int n, uint32_t value, unsigned long long timeoutNs, uint32_t *timeouts,
                         cudaStream_t s) {
    if (n <= 0) return;
    wait_signals_kernel<<<1, 64, 0, s>>>(flags, strideWords, n, value, timeoutNs, timeouts);
}

}  // namespace gr
