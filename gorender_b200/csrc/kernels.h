// kernels.h — host-callable launchers of the five per-draw kernels.
#pragma once

#include <cuda_runtime.h>

#include "gr_types.cuh"

namespace gr {

// K1: clip-space vertices of every visible object of every frame.
void launch_transform(const DrawArgs &a, int nframes, cudaStream_t s);
// K2: cull / light / clip / project / snap / emit + tile counts.
void launch_setup(const DrawArgs &a, int nframes, bool anyPlain, bool anyClip, cudaStream_t s);
// K3: exclusive scan of the per-tile counts (one block per frame).
void launch_bin_scan(const DrawArgs &a, int nframes, cudaStream_t s);
// K4: scatter triangle slots into the per-tile lists.
void launch_bin_fill(const DrawArgs &a, int nframes, cudaStream_t s);
// K5: per-tile coverage + z resolve in shared memory, shade, coalesced write-back.
void launch_raster(const DrawArgs &a, int nframes, cudaStream_t s);
// matrixMultiplyVec4Batch over a device array.
void launch_matvec_batch(const float m[16], float4 *vecs, long long n, cudaStream_t s);

}  // namespace gr
