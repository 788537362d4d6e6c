// gr_types.cuh — device-side data layout shared by the kernels and capi.cu.
//
// HBM layout.  Scene data is uploaded once (MeshDev,
// TexDev).  Everything a draw produces lives in a per-context workspace,
// struct-of-arrays over the frames of a batch:
//
//   tv        [frames][totalVerts]        float4   clip-space vertices (K1; only when stage capture is on)
//   rec       [frames][recCap]            PackedRec emitted triangles, 48 B   (K2 -> K3)
//   uv        [frames][recCap]            TriUV    24 B, textured faces only  (K2 -> K3)
//   warpCount [frames][nFaceBlocks*8]     u32      slots used per warp (stage capture only)
//   descCount [frames][nTiles]            u32      descriptors appended per tile (K2 -> K3; K3 re-zeroes)
//   desc      [frames][nTiles][descCap]   TileDesc per-tile triangle lists, 8 B per (warp, tile) group
//   overflow  [frames][overflowCap]       OverflowDesc  descriptors beyond descCap (rare; bounded pool)
//   bigList   [frames][recCap]            u32      triangles spanning > kMaxBinsPerTri tiles, and the
//                                                  triangles of descriptors that found the pool full
//   blockList [frames][nFaceBlocks]       u32      face blocks kept by reject_kernel (strip draws / partly visible objects)
//   counters  [frames]                    FrameCounters
//   ovl       [frames][H*W]               u64      overlay events (ShowEdges / ShowVertices only), see OverlayKey
#pragma once

#include <stdint.h>

#include "../../include/gorender_b200.h"
#include "gr_math.cuh"

namespace gr {

constexpr int kTile = GRB_TILE;            // raster tile edge (pixels)
constexpr int kTilePix = kTile * kTile;
#ifndef GRB_FACE_BLOCK
#define GRB_FACE_BLOCK 256
#endif
constexpr int kFaceBlock = GRB_FACE_BLOCK; // faces per setup block (32, 64, 128 or 256)
constexpr int kWarpsPerFaceBlock = kFaceBlock / 32;
constexpr int kWarpSlots = 32;             // rec slots reserved per warp, objects that do not clip
constexpr int kWarpSlotsClip = 256;        // ... objects that clip (<= 7 triangles per face)
constexpr int kMaxFan = 7;                 // clipping.go:10: <= 9 vertices -> <= 7 triangles
constexpr int kMaxBinsPerTri = 16;         // more tiles than this -> bigList
constexpr int kDescCap = 512;              // descriptors a tile holds in place; the rest go to `overflow`
constexpr int kCoordLimit = 16383;         // |snapped coord| bound of the int32 edge-function domain

struct MeshDev {
    const float4 *verts;
    const float4 *vnormals;
    const float4 *fnormals;
    // Face-corner expansion built once at upload: corner k of face f at cv[k][f] (object space)
    // and, for meshes with vertex normals, cn[k][f].  The per-frame kernels stream these with
    // fully coalesced 128-bit loads instead of gathering through the index arrays.
    const float4 *cv[3];
    const float4 *cn[3];
    // object-space bounds of every 32 consecutive faces (mesh.cu, warp_bounds_kernel)
    const float4 *warpLo, *warpHi;
    const int32_t *vidx;     // 3 per face
    const int32_t *nidx;     // 3 per face
    const float2 *uvs;       // 3 per face
    const int32_t *tex;      // per face, -1 = nil
    int32_t nv, nvn, nf, pad;
};

// upload-time preparation of one mesh (mesh.cu)
struct MeshPrepArgs {
    const float4 *verts, *vnormals;
    const int32_t *vidx, *nidx;
    int32_t nv, nvn, nf;
    float4 *cv[3], *cn[3];
    float4 *fnormalsOut;     // null: face normals were supplied by the caller
    int *error;              // bit 0: vertex index out of range, bit 1: normal index out of range
};

struct TexDev {
    const uchar4 *pixels;
    int32_t type, width, height;
    float widthF, heightF, scale;
    uchar4 color;
    int32_t pad;
};

// per (frame, object)
struct FrameObj {
    float mvp[16];
    float world[16];
    int32_t visibility;      // GRB_BOX_*
    uint32_t slotBase;       // first rec slot of this object in this frame
    int32_t pad[2];
};

// per object of the draw list (same for every frame of a batch)
struct DrawObj {
    int32_t mesh;
    int32_t vertBase;        // offset of the object's vertices in tv[frame]
    int32_t faceBlockBase;   // first face-block (submission order)
    int32_t vertBlockBase;   // first vertex-block
};

// Storage form of a record, 48 B: everything the coverage / depth phase of the raster kernel needs
// (snapped vertices as int16 — the parity domain is |coord| <= 16383 —, the three w and the raster
// bbox) in the first 32 bytes, what only the shading of a winning fragment needs in the last 16.
// Every byte is written (no partially written sectors, which cost a DRAM fill each), and the
// gather of a tile's ~600 records per frame — the raster kernel's main read — moves 32 B per triangle.
struct __align__(16) PackedRec {
    int16_t x0, y0, x1, y1, x2, y2;
    float w0, w1, w2;
    int16_t bx0, by0, bx1, by1;       // == int4 #1 .z/.w
    float i0, i1, i2;
    int32_t tex;                      // (the record's slot is its submission-order key: no order field)
};
static_assert(sizeof(PackedRec) == 48, "PackedRec must be 48 bytes");

struct __align__(16) TriRec {   // working form, == grb_triangle_rec
    int32_t x0, y0, x1, y1;
    int32_t x2, y2;
    float w0, w1;
    float w2, i0, i1, i2;
    int16_t bx0, by0, bx1, by1;
    int32_t tex;
    uint32_t order;          // the record's slot == submission-order key
};

// One entry of a tile's triangle list: the record slots base + {set bits of mask}.  A warp of
// the setup kernel appends one per (32-slot segment, tile) group, so a list entry stands for ~8
// triangles of C3 and costs one atomic.
struct __align__(8) TileDesc {
    uint32_t base, mask;
};
struct __align__(16) OverflowDesc {
    uint32_t tile, base, mask, pad;
};
static_assert(sizeof(TriRec) == 64, "TriRec must be 64 bytes");
static_assert(sizeof(grb_triangle_rec) == 64, "ABI record must be 64 bytes");

struct __align__(8) TriUV {
    float u0, v0, u1, v1, u2, v2;
};

struct __align__(16) FrameCounters {
    uint32_t triCount;
    uint32_t bigCount;
    uint32_t outOfDomain;
    uint32_t overflowCount;
    unsigned long long tpf;
    uint32_t listFallbacks;  // triangles sent to bigList because tile list and overflow pool were full
    uint32_t pad2;
};

// Overlays (drawProjection's ShowEdges / ShowVertices branches, renderer.go:191-216) are written
// with FrameBuffer.Pixel — no depth test, not clipped to the tile — by every tile pass that lists
// the triangle, so what a pixel finally shows is decided by the LAST write in the reference's
// serial order: tile pass, then list position, then face < edges/centre mark < vertex marks.  Every
// overlay pixel is recorded as max(event key) per pixel and the raster kernel's shading phase
// compares it with the key of the pixel's last face write (its owning tile pass, winning slot).
//   key = (tile pass + 1) << 34 | (record slot + 1) << 2 | kind      (0 = no overlay)
constexpr int kOvlTileShift = 34, kOvlSlotShift = 2;
constexpr unsigned long long kOvlKindEdge = 0ull, kOvlKindVertex = 1ull;
constexpr uint32_t kOptOverlayKeys = GRB_OPT_SHOW_EDGES | GRB_OPT_SHOW_VERTICES;
constexpr uint32_t kOptPostPass = kOptOverlayKeys | GRB_OPT_CROSSHAIR | GRB_OPT_FOG;

struct RefTiles {            // the reference's tile grid (renderer.go:50-76)
    int32_t ntx, nty;        // numTilesX, numTilesY (1 or 4)
    int32_t tw, th;          // tileWidth, tileHeight
    // calculateTileBoundaries per column / row, as floats: start and (clamped) end; NaN beyond ntx / nty
    float sx[4], ex[4], sy[4], ey[4];
};

struct DrawArgs {
    // scene
    const MeshDev *meshes;
    const TexDev *textures;
    int32_t ntex;
    // draw list
    const DrawObj *objs;
    const int32_t *vblkObj;     // vertex-block -> object
    const int32_t *fblkObj;     // face-block   -> object
    const FrameObj *frameObjs;  // [frames][nobj]
    int32_t nobj, nVertBlocks, nFaceBlocks, totalVerts;
    // workspace (per-frame strides in elements)
    float4 *tv;
    PackedRec *rec;
    TriUV *uv;
    uint32_t *warpCount;        // null unless stage capture is on
    uint32_t *descCount;
    TileDesc *desc;
    OverflowDesc *overflow;
    uint32_t overflowCap;       // entries of the overflow pool per frame
    uint32_t descCap;
    uint32_t *bigList;
    FrameCounters *counters;
    uint32_t recCap;
    unsigned long long *ovl;    // null unless ShowEdges / ShowVertices
    // target
    uchar4 *color;              // [frames][H][W]
    float *depth;               // [frames][H][W]
    uint8_t *tileBusy;          // [frames][nTiles] 0: the tile holds the cleared background only (host mirrors skip it)
    // One-frame draws with host mirrors attached (grb_draw_present): the raster kernel's write-back also stores the
    // tile into the mirrors' host planes (MIRROR instantiation), so that the PCIe transfer of a tile overlaps the
    // rasterisation of the others instead of following the whole frame.  Null: a separate mirror update follows.
    uchar4 *mirColor;
    float *mirDepth;
    uint8_t *mirDirtyColor, *mirDirtyDepth;           // [nTiles] of the mirror frame written
    unsigned long long *mirWrittenColor, *mirWrittenDepth;
    int32_t width, height;
    int32_t ntx, nty;           // device tiles
    int32_t tileRowBegin, tileRowEnd;  // strip, in tile rows
    // face blocks (low 24 bits) and their surviving warps (high 8) kept by reject_kernel, [frames][nFaceBlocks], and
    // their number per frame; null: every block is set up (whole-frame draws of objects inside the frustum)
    uint32_t *blockList;
    uint32_t *blockCount;
    // params
    Mat4 screen;
    int32_t screenNoZ;          // screen.m[2] == 0 && screen.m[6] == 0 (NewScreenMatrix): see to_screen
    FmaConsts fma;              // run-time -0 / 1 for the packed FFMA2 products (gr_math.cuh)
    float lx, ly, lz;
    uint32_t options;
    float zNear, zFar;
    RefTiles ref;
    float fogStart, fogEnd;     // FrameBuffer.Fog arguments (rasterizer.go:193), GRB_OPT_FOG only
    uchar4 fogColor;
};

// One update of a mirror pair (present.cu); every pointer is already offset to its first frame.
struct MirrorArgs {
    const uchar4 *color;        // device frames
    const float *depth;
    const uint8_t *tileBusy;    // [frames][nTiles], written by the raster kernel
    uchar4 *hostColor;          // device-visible address of the mirror planes: pinned host memory, or another GPU's
                                // framebuffer (null: plane not mirrored)
    float *hostDepth;
    uint8_t *dirtyColor;        // [frames][nTiles] device flags: the host tile is not the cleared background
    uint8_t *dirtyDepth;
    unsigned long long *tilesWrittenColor, *tilesWrittenDepth;   // statistics, one counter per mirror (may be null)
    uint8_t *targetBusy;        // the mirror is another framebuffer (a strip pushed to its owner): that framebuffer's
                                // own per-tile flags, kept in step so that ITS mirrors see these tiles (else null)
    int32_t width, height, ntx, nty;
    int32_t tileRow0, tileRows; // tile rows covered by the launch (a strip), default all
    int32_t full;               // ignore tileBusy: copy every tile (framebuffers this library does not own)
};

}  // namespace gr
