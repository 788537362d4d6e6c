// kernels.h — host-callable launchers of the per-draw kernels.
#pragma once

#include <cuda_runtime.h>

#include "gr_types.cuh"

namespace gr {

// K1: clip-space vertices of every visible object of every frame (stage capture / seam only).
void launch_transform(const DrawArgs &a, int nframes, cudaStream_t s);
// K2 pre-pass (strip draws, objects reaching outside the frustum): the face blocks / warps whose bounds can reach the draw's rows.
void launch_reject(const DrawArgs &a, int nframes, cudaStream_t s);
// K2: transform / cull / light / clip / project / snap / emit records + per-tile descriptor lists.
void launch_setup(const DrawArgs &a, int nframes, bool anyPlain, bool anyClip, cudaStream_t s);
// K3: per-tile coverage + z resolve in shared memory, shade, coalesced write-back.
void launch_raster(const DrawArgs &a, int nframes, cudaStream_t s);
// Upload time: index check + face-corner expansion (+ NewMesh's face normals), boundingBox keys.
void launch_mesh_prepare(const MeshPrepArgs &p, cudaStream_t s);
void launch_bbox(const float4 *verts, int nv, uint32_t *out7, cudaStream_t s);
void launch_warp_bounds(const float4 *const cv[3], int nf, float4 *warpLo, float4 *warpHi, cudaStream_t s);
// matrixMultiplyVec4Batch over a device array.
void launch_matvec_batch(const float m[16], float4 *vecs, long long n, cudaStream_t s);

// Host mirrors: copy the tiles that are (or were) busy straight into mapped host memory.
void launch_mirror_update(const MirrorArgs &m, int nframes, cudaStream_t s);

// Hand-off flags of a framebuffer shared across processes (sort-first strips).
void launch_signal(uint32_t *flag, uint32_t value, cudaStream_t s);
void launch_wait_signals(uint32_t *flags, int strideWords, int n, uint32_t value, unsigned long long timeoutNs, uint32_t *timeouts,
                         cudaStream_t s);

}  // namespace gr
