// capi.cu — the C ABI (include/gorender_b200.h): context, asset upload,
// workspace management and the per-draw orchestration that replaces the body
// of (*Renderer).Draw (renderer.go:443-483).
//
// Host work per draw is what the reference also does once per object per
// frame and is not per-vertex: BoxVisibility of the 8 bounding-box corners
// (renderer.go:268-275, clipping.go:131-154).  Everything per-vertex,
// per-face and per-pixel runs in the kernels of kernels.h (two per batch of
// frames: setup and raster).  There is no CPU rendering path in this library.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <limits>
#include <string>
#include <vector>

#include "gr_types.cuh"
#include "kernels.h"

using namespace gr;

namespace {

std::string g_createError;

struct MeshHost {
    bool live = false;
    MeshDev dev{};
    float bbox[32];
    std::vector<void *> allocs;
};

struct TexHost {
    TexDev dev{};
    void *pixels = nullptr;
};

template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
};

constexpr int kStagingRing = 4;
constexpr int kFlagStride = 32;   // words between two hand-off flags: one 128-byte line each

struct IpcHandle {       // what grb_framebuffer_ipc_export writes into the caller's GRB_IPC_HANDLE_BYTES
    cudaIpcMemHandle_t color, depth, flags, busy;
    int32_t width, height, frames, magic;
};
static_assert(sizeof(IpcHandle) <= GRB_IPC_HANDLE_BYTES, "IPC handle does not fit");

}  // namespace

struct grb_framebuffer {
    grb_context *ctx;
    int32_t width, height, frames;
    uchar4 *color;
    float *depth;
    bool owned;
    // [frames][nTiles] one byte per device tile, written by the raster kernel: the tile holds something
    // other than the cleared background (1 until a frame has been drawn: contents unknown)
    uint8_t *tileBusy = nullptr;
    // hand-off flags (GRB_SIGNAL_SLOTS words, kFlagStride apart) of a framebuffer shared across processes
    uint32_t *flags = nullptr;
    bool ipcOpened = false;       // colour / depth / flags are another process's memory (cudaIpcOpenMemHandle)
    // read-backs run on the context's copy stream so that the D2H of one
    // framebuffer overlaps the rendering of another
    cudaEvent_t drawDone = nullptr, readDone = nullptr;
    bool pendingRead = false;
};

struct grb_mirror {
    grb_context *ctx;
    int32_t width, height, frames, plane;
    void *devPtr;                 // device-visible address of the plane: pinned host memory, or a framebuffer's plane
    grb_framebuffer *target = nullptr;   // the mirror is this framebuffer's plane (a strip pushed to the frame's owner)
    uint8_t *dirty = nullptr;     // [frames][nTiles] device flags: the HOST tile is not the cleared background
    unsigned long long *tilesWritten = nullptr;   // device counter
    int64_t tilesFull = 0;        // tiles full copies would have moved
    cudaEvent_t done = nullptr;
    bool pending = false;
};

struct GraphEntry;

struct grb_context {
    int device = 0;
    int smCount = 148;
    cudaStream_t ownStream = nullptr, stream = nullptr, copyStream = nullptr;
    mutable std::string err;

    std::vector<MeshHost> meshes;
    DevBuf<MeshDev> dMeshes;
    bool meshesDirty = true;
    std::vector<TexHost> textures;
    DevBuf<TexDev> dTextures;
    bool texturesDirty = true;

    // draw-list plan (same meshes => same tables)
    std::vector<int32_t> planMeshes;
    std::vector<DrawObj> planObjs;
    DevBuf<DrawObj> dObjs;
    DevBuf<int32_t> dVblk, dFblk;
    int32_t nVertBlocks = 0, nFaceBlocks = 0, totalVerts = 0;
    bool planValid = false;

    // per-(frame,object) matrices: pinned staging ring + device copy
    FrameObj *hFrameObjs[kStagingRing] = {};
    size_t hFrameObjsCap[kStagingRing] = {};
    cudaEvent_t stagingDone[kStagingRing] = {};
    int stagingNext = 0;
    DevBuf<FrameObj> dFrameObjs;

    // staging + read-back of the one-call Draw (grb_draw_present): synchronous, so one buffer each
    FrameObj *hGraphObjs = nullptr;
    size_t hGraphObjsCap = 0;
    FrameCounters *hCounters = nullptr;
    size_t hCountersCap = 0;
    std::vector<GraphEntry *> graphs;   // cached CUDA graphs of one-frame draws (most recent first)
    int64_t graphReplays = 0, graphCaptures = 0;

    // workspace
    uint64_t workspaceLimit = 0;        // bytes; 0 = a quarter of the device's memory
    uint64_t deviceMemory = 0;
    uint32_t overflowCap = 0;           // per frame
    uint32_t overflowCapForced = 0;     // grb_debug_set_overflow_cap (tests of the fallback path)
    uint32_t *dTimeouts = nullptr;      // device counter of signal waits that gave up
    bool descDirty = false;             // a draw's setup ran but its raster did not: descCount is not all zero
    DevBuf<float4> tv;
    DevBuf<PackedRec> rec;
    DevBuf<TriUV> uv;
    DevBuf<uint32_t> warpCount, descCount, bigList, blockList, blockCount;
    DevBuf<TileDesc> desc;
    DevBuf<OverflowDesc> overflow;
    DevBuf<FrameCounters> counters;
    DevBuf<unsigned long long> ovl;  // per-pixel overlay event keys (ShowEdges / ShowVertices draws only)
    uint32_t recCap = 0;  // per frame

    // last draw (for stats / debug read-backs)
    int32_t lastFrames = 0, lastNobj = 0, lastNTiles = 0;
    uint32_t lastOptions = 0;
    std::vector<int32_t> lastVisibility;

    // seam scratch
    DevBuf<float4> seam;

    // also run K1 so that grb_debug_read_transformed has something to read
    bool stageCapture = false;

    // timing
    bool timing = false;
    cudaEvent_t tev[6] = {};
    double accMs[5] = {0, 0, 0, 0, 0};
    int64_t accLaunches = 0;
    int64_t totalLaunches = 0;
};

namespace {

int32_t fail(const grb_context *ctx, int32_t code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_createError = msg;
    return code;
}

#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            const int32_t code__ = (e__ == cudaErrorMemoryAllocation) ? GRB_ERR_OOM : GRB_ERR_CUDA;     \
            return fail(ctx, code__, std::string(#call) + ": " + cudaGetErrorString(e__));              \
        }                                                                                               \
    } while (0)

void free_graph(GraphEntry *g);

// Grow a device buffer to at least n elements.  Old contents are dropped;
// zero-filled when `zero`.  Synchronises the stream before freeing.
template <typename T> int32_t ensure(grb_context *ctx, DevBuf<T> &b, size_t n, bool zero) {
    if (n <= b.cap && b.p) return GRB_OK;
    if (n == 0) n = 1;
    if (b.p) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t grow = n + n / 4;
    cudaError_t e = cudaMalloc(&b.p, grow * sizeof(T));
    size_t got = grow;
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, n * sizeof(T));
        got = n;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return fail(ctx, GRB_ERR_OOM, "cudaMalloc of " + std::to_string(n * sizeof(T)) + " bytes failed");
    }
    b.cap = got;
    if (zero) CK(ctx, cudaMemsetAsync(b.p, 0, got * sizeof(T), ctx->stream));
    return GRB_OK;
}

template <typename T> int32_t upload(grb_context *ctx, const T *src, size_t n, T **out, std::vector<void *> &allocs) {
    *out = nullptr;
    if (n == 0 || src == nullptr) return GRB_OK;
    void *p = nullptr;
    CK(ctx, cudaMalloc(&p, n * sizeof(T)));
    allocs.push_back(p);
    CK(ctx, cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<T *>(p);
    return GRB_OK;
}

// Frustum.BoxVisibility (clipping.go:131-154) with the planes of
// NewFrustum (clipping.go:93-126); Plane.DistanceToVertex (clipping.go:69-71).
int box_visibility(const float4 corners[8], float zn, float zf) {
    const float4 P[6] = {{-1, 0, 0, 1}, {1, 0, 0, 1}, {0, -1, 0, 1}, {0, 1, 0, 1}, {0, 0, zn, 1}, {0, 0, zf, 1}};
    const float4 N[6] = {{1, 0, 0, 1}, {-1, 0, 0, 1}, {0, 1, 0, 1}, {0, -1, 0, 1}, {0, 0, -1, 0}, {0, 0, 1, 0}};
    for (int i = 0; i < 6; i++) {
        int outside = 0;
        for (int c = 0; c < 8; c++)
            if (fsub(dot4(N[i], corners[c]), dot4(N[i], P[i])) > 0.0f) outside++;
        if (outside == 8) return GRB_BOX_OUTSIDE;
        if (outside > 0) return GRB_BOX_INTERSECT;
    }
    return GRB_BOX_INSIDE;
}

// Rows [y0, y1) of a strip draw against the projected corners of an object's bounding box: true when nothing inside
// the box can reach them (every corner in front of the eye — clip w < 0, SURVEY H5 — and the screen-space row range,
// two pixels of margin, clamped into the frame like the triangles' own rows, outside the strip).
bool box_misses_rows(const float4 corners[8], const float screen[16], int height, int y0, int y1) {
    float lo = std::numeric_limits<float>::infinity(), hi = -lo;
    for (int c = 0; c < 8; c++) {
        const float4 v = corners[c];
        if (!(v.w < -1e-6f)) return false;
        const float sy = screen[5] * (v.y / v.w) + screen[7];
        if (!(std::fabs(sy) < 1e9f)) return false;
        lo = std::min(lo, sy);
        hi = std::max(hi, sy);
    }
    const float margin = 2.0f;
    if (hi < -margin || lo > (float)height + margin) return false;   // off screen: BoxVisibility's business
    const int rlo = std::min(std::max((int)std::floor(lo - margin), 0), height - 1);
    const int rhi = std::min(std::max((int)std::ceil(hi + margin), 0), height - 1);
    return rhi < y0 || rlo >= y1;
}

int32_t sync_tables(grb_context *ctx) {
    if (ctx->meshesDirty) {
        std::vector<MeshDev> t(ctx->meshes.size());
        for (size_t i = 0; i < t.size(); i++) t[i] = ctx->meshes[i].dev;
        if (int32_t r = ensure(ctx, ctx->dMeshes, t.size(), false)) return r;
        if (!t.empty())
            CK(ctx, cudaMemcpyAsync(ctx->dMeshes.p, t.data(), t.size() * sizeof(MeshDev), cudaMemcpyHostToDevice,
                                    ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->meshesDirty = false;
    }
    if (ctx->texturesDirty) {
        std::vector<TexDev> t(ctx->textures.size());
        for (size_t i = 0; i < t.size(); i++) t[i] = ctx->textures[i].dev;
        if (int32_t r = ensure(ctx, ctx->dTextures, t.size(), false)) return r;
        if (!t.empty())
            CK(ctx, cudaMemcpyAsync(ctx->dTextures.p, t.data(), t.size() * sizeof(TexDev), cudaMemcpyHostToDevice,
                                    ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->texturesDirty = false;
    }
    return GRB_OK;
}

int32_t build_plan(grb_context *ctx, const grb_object *objects, int32_t nobj) {
    bool same = ctx->planValid && (int32_t)ctx->planMeshes.size() == nobj;
    for (int32_t i = 0; same && i < nobj; i++) same = ctx->planMeshes[i] == objects[i].mesh;
    if (same) return GRB_OK;

    ctx->planValid = false;
    ctx->planMeshes.resize(nobj);
    ctx->planObjs.resize(nobj);
    std::vector<int32_t> vblk, fblk;
    int64_t verts = 0;
    for (int32_t i = 0; i < nobj; i++) {
        const MeshDev &m = ctx->meshes[objects[i].mesh].dev;
        DrawObj &o = ctx->planObjs[i];
        ctx->planMeshes[i] = objects[i].mesh;
        o.mesh = objects[i].mesh;
        o.vertBase = (int32_t)verts;
        o.vertBlockBase = (int32_t)vblk.size();
        o.faceBlockBase = (int32_t)fblk.size();
        verts += m.nv;
        vblk.insert(vblk.end(), (m.nv + 255) / 256, i);
        fblk.insert(fblk.end(), (m.nf + kFaceBlock - 1) / kFaceBlock, i);
    }
    if (verts > INT32_MAX) return fail(ctx, GRB_ERR_INVALID, "draw list has more than 2^31 vertices");
    if ((int64_t)fblk.size() * kWarpsPerFaceBlock * kWarpSlotsClip >= (int64_t)UINT32_MAX)
        return fail(ctx, GRB_ERR_INVALID, "draw list has too many faces for 32-bit record slots");
    ctx->totalVerts = (int32_t)verts;
    ctx->nVertBlocks = (int32_t)vblk.size();
    ctx->nFaceBlocks = (int32_t)fblk.size();
    if (int32_t r = ensure(ctx, ctx->dObjs, (size_t)nobj, false)) return r;
    if (int32_t r = ensure(ctx, ctx->dVblk, vblk.size(), false)) return r;
    if (int32_t r = ensure(ctx, ctx->dFblk, fblk.size(), false)) return r;
    // tables may still be in use by an in-flight draw
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (nobj) CK(ctx, cudaMemcpy(ctx->dObjs.p, ctx->planObjs.data(), nobj * sizeof(DrawObj), cudaMemcpyHostToDevice));
    if (!vblk.empty())
        CK(ctx, cudaMemcpy(ctx->dVblk.p, vblk.data(), vblk.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (!fblk.empty())
        CK(ctx, cudaMemcpy(ctx->dFblk.p, fblk.data(), fblk.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    ctx->planValid = true;
    return GRB_OK;
}

int32_t set_device(const grb_context *ctx) {
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, GRB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return GRB_OK;
}

// Everything one draw call needs, worked out on the host before anything is queued.  Plain data, zero
// filled before use: two jobs that compare equal byte for byte queue exactly the same work, which is what
// the CUDA-graph cache of grb_draw_present keys on.
struct DrawJob {
    DrawArgs a;              // pointers of the batch's first frame; enqueue_draw offsets them per launch
    int32_t nframes, nobj, chunk, nTiles;
    int32_t anyPlain, anyClip, overlayKeys, ring;
    size_t nfo, npix;
    FrameObj *hfo;           // pinned staging of the per-(frame, object) matrices, filled
};

uint64_t workspace_limit(const grb_context *ctx) {
    if (ctx->workspaceLimit) return ctx->workspaceLimit;
    return ctx->deviceMemory ? ctx->deviceMemory / 4 : (uint64_t)32 << 30;
}

// Host part of a draw: validation, BoxVisibility per (frame, object) (renderer.go:268-275), the matrices into
// pinned staging, workspace sizing.  `oneCall`: the synchronous one-call path (its own staging buffer).
int32_t prepare_draw(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, const grb_object *objects,
                     int32_t nobj, const grb_draw_params *prm, bool oneCall, DrawJob &job) {
    std::memset(&job, 0, sizeof job);
    if (!ctx || !fb || !prm) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (fb->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "framebuffer belongs to another context");
    if (nframes <= 0 || frame0 < 0 || frame0 + nframes > fb->frames)
        return fail(ctx, GRB_ERR_INVALID, "frame range outside the framebuffer");
    if (nobj < 0 || (nobj > 0 && !objects)) return fail(ctx, GRB_ERR_INVALID, "bad object list");
    if (prm->ref_tiles != 1 && prm->ref_tiles != 16)
        return fail(ctx, GRB_ERR_INVALID, "ref_tiles must be 16 (parallel) or 1 (serial): renderer.go:144,151");
    if (fb->width > kCoordLimit || fb->height > kCoordLimit)
        return fail(ctx, GRB_ERR_INVALID, "framebuffer larger than 16383 pixels on a side");
    for (int32_t i = 0; i < nobj; i++) {
        const int32_t m = objects[i].mesh;
        if (m < 0 || m >= (int32_t)ctx->meshes.size() || !ctx->meshes[m].live)
            return fail(ctx, GRB_ERR_INVALID, "object " + std::to_string(i) + " refers to an unknown mesh");
        for (int32_t f = 1; f < nframes; f++)
            if (objects[(size_t)f * nobj + i].mesh != m)
                return fail(ctx, GRB_ERR_INVALID, "all frames of a batch must draw the same meshes");
    }
    if (int32_t r = set_device(ctx)) return r;

    const int ntx = (fb->width + kTile - 1) / kTile, nty = (fb->height + kTile - 1) / kTile;
    int rowBegin = 0, rowEnd = nty;
    if (!(prm->row_begin == 0 && prm->row_end == 0)) {
        if (prm->row_begin < 0 || prm->row_end <= prm->row_begin || prm->row_begin % kTile ||
            (prm->row_end % kTile && prm->row_end != fb->height) || prm->row_end > fb->height)
            return fail(ctx, GRB_ERR_INVALID, "row range must be tile-aligned and inside the frame");
        rowBegin = prm->row_begin / kTile;
        rowEnd = (prm->row_end + kTile - 1) / kTile;
    }

    if (int32_t r = sync_tables(ctx)) return r;
    if (int32_t r = build_plan(ctx, objects, nobj)) return r;

    // ---- per (frame, object): matrices + BoxVisibility (renderer.go:268-275)
    const size_t nfo = (size_t)nframes * std::max(nobj, 1);
    FrameObj *hfo;
    if (oneCall) {
        // the previous one-call draw has completed (it synchronises): the buffer is free
        if (ctx->hGraphObjsCap < nfo) {
            if (ctx->hGraphObjs) CK(ctx, cudaFreeHost(ctx->hGraphObjs));
            ctx->hGraphObjs = nullptr;
            ctx->hGraphObjsCap = 0;
            for (GraphEntry *g : ctx->graphs) free_graph(g);   // they copy from the old buffer
            ctx->graphs.clear();
            CK(ctx, cudaHostAlloc((void **)&ctx->hGraphObjs, (nfo + 16) * sizeof(FrameObj), cudaHostAllocDefault));
            ctx->hGraphObjsCap = nfo + 16;
        }
        hfo = ctx->hGraphObjs;
        job.ring = -1;
    } else {
        const int ring = ctx->stagingNext;
        ctx->stagingNext = (ring + 1) % kStagingRing;
        if (ctx->hFrameObjsCap[ring] < nfo) {
            if (ctx->hFrameObjs[ring]) {
                CK(ctx, cudaEventSynchronize(ctx->stagingDone[ring]));
                CK(ctx, cudaFreeHost(ctx->hFrameObjs[ring]));
                ctx->hFrameObjs[ring] = nullptr;
            }
            CK(ctx, cudaHostAlloc((void **)&ctx->hFrameObjs[ring], nfo * sizeof(FrameObj), cudaHostAllocDefault));
            ctx->hFrameObjsCap[ring] = nfo;
        } else {
            CK(ctx, cudaEventSynchronize(ctx->stagingDone[ring]));
        }
        hfo = ctx->hFrameObjs[ring];
        job.ring = ring;
    }

    const bool optClip = prm->options & GRB_OPT_FRUSTUM_CLIPPING;
    const bool viewportShape = prm->screen[2] == 0.0f && prm->screen[6] == 0.0f && prm->screen[1] == 0.0f && prm->screen[4] == 0.0f;
    const bool stripReject = viewportShape && !(prm->options & kOptOverlayKeys) && (rowBegin != 0 || rowEnd != nty);
    bool anyPlain = false, anyClip = false;
    uint64_t recNeed = 1, facesMax = 1;
    ctx->lastVisibility.assign((size_t)nframes * nobj, GRB_BOX_OUTSIDE);
    for (int32_t f = 0; f < nframes; f++) {
        uint64_t need = 0, faces = 0;  // record slots of this frame: static per-warp segments (setup.cu)
        for (int32_t i = 0; i < nobj; i++) {
            const grb_object &src = objects[(size_t)f * nobj + i];
            FrameObj &dst = hfo[(size_t)f * nobj + i];
            std::memcpy(dst.mvp, src.mvp, 64);
            std::memcpy(dst.world, src.world, 64);
            const MeshHost &mh = ctx->meshes[src.mesh];
            float4 corners[8];
            for (int c = 0; c < 8; c++) {
                const float4 p = make_float4(mh.bbox[4 * c], mh.bbox[4 * c + 1], mh.bbox[4 * c + 2], mh.bbox[4 * c + 3]);
                corners[c] = mat_vec(src.mvp, p);  // matrixMultiplyVec4Batch(&mvpMatrix, bbox[:])
            }
            int vis = box_visibility(corners, prm->z_near, prm->z_far);
            // sort-first strips: an object whose projected bounding box misses this draw's rows altogether is skipped
            // like an invisible one (same conservative rule as the per-block test of the setup kernel; its triangles are
            // drawn and counted by the ranks that own those rows)
            if (stripReject && vis != GRB_BOX_OUTSIDE && box_misses_rows(corners, prm->screen, fb->height, rowBegin * kTile, rowEnd * kTile))
                vis = GRB_BOX_OUTSIDE;
            dst.visibility = vis;
            dst.slotBase = (uint32_t)need;
            dst.pad[0] = dst.pad[1] = 0;
            ctx->lastVisibility[(size_t)f * nobj + i] = vis;
            if (vis == GRB_BOX_OUTSIDE) continue;
            const bool clips = optClip && vis != GRB_BOX_INSIDE;
            (clips ? anyClip : anyPlain) = true;
            const uint64_t blocks = (mh.dev.nf + kFaceBlock - 1) / kFaceBlock;
            need += blocks * kWarpsPerFaceBlock * (clips ? kWarpSlotsClip : kWarpSlots);
            faces += (uint64_t)mh.dev.nf;
        }
        recNeed = std::max(recNeed, need);
        facesMax = std::max(facesMax, faces);
    }
    if (recNeed >= UINT32_MAX) return fail(ctx, GRB_ERR_INVALID, "too many triangles per frame");

    // ---- workspace.  Records have static slots (exact worst case); tile lists hold kDescCap descriptors in
    // place and share a bounded overflow pool per frame; what fits nowhere goes to the frame-wide list (setup.cu).
    const int nTiles = ntx * nty;
    const uint32_t recCap = std::max<uint32_t>(ctx->recCap, (uint32_t)recNeed);
    const uint64_t ovWorst = (uint64_t)recCap * kMaxBinsPerTri;
    const uint32_t ovCap = ctx->overflowCapForced
                               ? ctx->overflowCapForced
                               : std::max<uint32_t>(ctx->overflowCap, (uint32_t)std::min<uint64_t>(ovWorst, std::max<uint64_t>(8192, 2 * facesMax)));
    const bool overlayKeys = (prm->options & kOptOverlayKeys) != 0;
    const size_t npix = (size_t)fb->width * fb->height;
    const size_t nWarps = (size_t)std::max(ctx->nFaceBlocks, 1) * kWarpsPerFaceBlock;
    // a batch whose workspace would not fit the limit is rendered in several launches of `chunk` frames
    uint64_t perFrame = (uint64_t)recCap * (sizeof(PackedRec) + sizeof(TriUV) + 4) + (uint64_t)ovCap * sizeof(OverflowDesc) +
                        (uint64_t)nTiles * (kDescCap * sizeof(TileDesc) + 4);
    if (overlayKeys) perFrame += npix * 8;
    size_t chunk = (size_t)std::min<uint64_t>((uint64_t)nframes, std::max<uint64_t>(1, workspace_limit(ctx) / perFrame));
    chunk = std::min<size_t>(chunk, 32768);   // grid.y / grid.z limit of the launches
    if (ctx->stageCapture) {
        if (nframes > 32768) return fail(ctx, GRB_ERR_INVALID, "stage capture draws at most 32768 frames per call");
        chunk = (size_t)nframes;              // the stage read-backs address the whole batch
    }
    const size_t F = (size_t)nframes;
    if (int32_t r = ensure(ctx, ctx->dFrameObjs, nfo, false)) return r;
    if (ctx->stageCapture)
        if (int32_t r = ensure(ctx, ctx->tv, chunk * std::max(ctx->totalVerts, 1), false)) return r;
    if (int32_t r = ensure(ctx, ctx->rec, chunk * recCap, false)) return r;
    if (int32_t r = ensure(ctx, ctx->uv, chunk * recCap, false)) return r;
    if (int32_t r = ensure(ctx, ctx->bigList, chunk * recCap, false)) return r;
    if (int32_t r = ensure(ctx, ctx->overflow, chunk * ovCap, false)) return r;
    if (int32_t r = ensure(ctx, ctx->desc, chunk * nTiles * kDescCap, false)) return r;
    if (ctx->stageCapture)
        if (int32_t r = ensure(ctx, ctx->warpCount, chunk * nWarps, false)) return r;
    if (int32_t r = ensure(ctx, ctx->counters, F, false)) return r;
    if (overlayKeys)
        if (int32_t r = ensure(ctx, ctx->ovl, chunk * npix, false)) return r;
    ctx->recCap = recCap;
    if (!ctx->overflowCapForced) ctx->overflowCap = ovCap;
    // per-tile descriptor counters must be zero on entry.  The raster kernel re-zeroes what it consumed, but the
    // per-frame stride depends on nTiles (a strip draw leaves the other rows' tiles untouched, at zero), so they
    // are cleared when the geometry changes — and after a draw whose setup ran but whose raster did not
    if (ctx->descCount.cap < chunk * nTiles || ctx->lastNTiles != nTiles || ctx->descDirty) {
        if (int32_t r = ensure(ctx, ctx->descCount, chunk * nTiles, false)) return r;
        CK(ctx, cudaMemsetAsync(ctx->descCount.p, 0, ctx->descCount.cap * sizeof(uint32_t), ctx->stream));
        ctx->descDirty = false;
    }

    DrawArgs &a = job.a;
    a.meshes = ctx->dMeshes.p;
    a.textures = ctx->dTextures.p;
    a.ntex = (int32_t)ctx->textures.size();
    a.objs = ctx->dObjs.p;
    a.vblkObj = ctx->dVblk.p;
    a.fblkObj = ctx->dFblk.p;
    a.frameObjs = ctx->dFrameObjs.p;
    a.nobj = nobj;
    a.nVertBlocks = ctx->nVertBlocks;
    a.nFaceBlocks = ctx->nFaceBlocks;
    a.totalVerts = ctx->totalVerts;
    a.tv = ctx->tv.p;
    a.rec = ctx->rec.p;
    a.uv = ctx->uv.p;
    a.warpCount = ctx->stageCapture ? ctx->warpCount.p : nullptr;
    a.descCount = ctx->descCount.p;
    a.desc = ctx->desc.p;
    a.overflow = ctx->overflow.p;
    a.overflowCap = ovCap;
    a.descCap = kDescCap;
    a.bigList = ctx->bigList.p;
    a.counters = ctx->counters.p;
    a.recCap = recCap;
    a.ovl = overlayKeys ? ctx->ovl.p : nullptr;
    a.fogStart = prm->fog_start;
    a.fogEnd = prm->fog_end;
    a.fogColor = make_uchar4(prm->fog_color[0], prm->fog_color[1], prm->fog_color[2], prm->fog_color[3]);
    a.color = fb->color + (size_t)frame0 * npix;
    a.depth = fb->depth + (size_t)frame0 * npix;
    a.tileBusy = fb->tileBusy ? fb->tileBusy + (size_t)frame0 * nTiles : nullptr;
    a.width = fb->width;
    a.height = fb->height;
    a.ntx = ntx;
    a.nty = nty;
    a.tileRowBegin = rowBegin;
    a.tileRowEnd = rowEnd;
    std::memcpy(a.screen.m, prm->screen, 64);
    a.fma.negZero = make_float2(-0.0f, -0.0f);
    a.fma.one = make_float2(1.0f, 1.0f);
    a.lx = prm->light[0]; a.ly = prm->light[1]; a.lz = prm->light[2];
    a.options = prm->options;
    a.zNear = prm->z_near;
    a.zFar = prm->z_far;
    a.screenNoZ = prm->screen[2] == 0.0f && prm->screen[6] == 0.0f;
    // block-level rejection (setup.cu) needs a NewScreenMatrix-shaped viewport (x from x/w, y from y/w only); it pays
    // when a strip is drawn or when some object reaches outside the frustum
    const bool rejectBlocks = viewportShape && !overlayKeys && !ctx->stageCapture && (rowBegin != 0 || rowEnd != nty || anyClip) &&
                              ctx->nFaceBlocks < (1 << 24);
    if (rejectBlocks) {
        if (int32_t r = ensure(ctx, ctx->blockList, chunk * (size_t)std::max(ctx->nFaceBlocks, 1), false)) return r;
        if (int32_t r = ensure(ctx, ctx->blockCount, F, false)) return r;
        a.blockList = ctx->blockList.p;
        a.blockCount = ctx->blockCount.p;
    }
    if (prm->ref_tiles == 1) {
        a.ref.ntx = a.ref.nty = 1;
        a.ref.tw = fb->width;
        a.ref.th = fb->height;
    } else {
        // calculateTileBoundaries (renderer.go:56-59) for numTiles = 16
        const int n = 16, rtx = 4, rty = (n + rtx - 1) / rtx;
        a.ref.ntx = rtx;
        a.ref.nty = rty;
        a.ref.tw = (fb->width + rtx - 1) / rtx;
        a.ref.th = (fb->height + rty - 1) / rty;
    }
    {   // renderer.go:62-73 per column and per row (numTiles == 1: the whole frame, :51-53)
        const float Wf = (float)fb->width, Hf = (float)fb->height, nan = std::numeric_limits<float>::quiet_NaN();
        for (int c = 0; c < 4; c++) {
            a.ref.sx[c] = a.ref.ex[c] = a.ref.sy[c] = a.ref.ey[c] = nan;
            if (c < a.ref.ntx) {
                a.ref.sx[c] = (float)(c * a.ref.tw);
                a.ref.ex[c] = a.ref.sx[c] + (float)a.ref.tw;
                if (a.ref.ntx == 1 || a.ref.ex[c] > Wf) a.ref.ex[c] = Wf;
            }
            if (c < a.ref.nty) {
                a.ref.sy[c] = (float)(c * a.ref.th);
                a.ref.ey[c] = a.ref.sy[c] + (float)a.ref.th;
                if (a.ref.nty == 1 || a.ref.ey[c] > Hf) a.ref.ey[c] = Hf;
            }
        }
    }
    job.nframes = nframes;
    job.nobj = nobj;
    job.chunk = (int32_t)chunk;
    job.nTiles = nTiles;
    job.anyPlain = anyPlain;
    job.anyClip = anyClip;
    job.overlayKeys = overlayKeys;
    job.nfo = nfo;
    job.npix = npix;
    job.hfo = hfo;

    ctx->lastFrames = nframes;
    ctx->lastNobj = nobj;
    ctx->lastNTiles = nTiles;
    ctx->lastOptions = prm->options;
    return GRB_OK;
}

int launches_of(const grb_context *ctx, const DrawJob &job) {
    int per = 1;  // raster
    if (job.anyPlain || job.anyClip) per += (ctx->stageCapture ? 1 : 0) + (job.anyPlain ? 1 : 0) + (job.anyClip ? 1 : 0) + (job.a.blockList ? 1 : 0);
    return per * ((job.nframes + job.chunk - 1) / job.chunk);
}

// Device part: queue the matrix upload, the counter reset and the kernels of every launch of the batch on the
// render stream.  `capture`: the stream is being captured into a CUDA graph (no event work, no timing).
int32_t enqueue_draw(grb_context *ctx, const DrawJob &job, bool capture) {
    cudaStream_t s = ctx->stream;
    const size_t F = (size_t)job.nframes;
    CK(ctx, cudaMemcpyAsync(ctx->dFrameObjs.p, job.hfo, job.nfo * sizeof(FrameObj), cudaMemcpyHostToDevice, s));
    if (job.ring >= 0 && !capture) CK(ctx, cudaEventRecord(ctx->stagingDone[job.ring], s));
    CK(ctx, cudaMemsetAsync(ctx->counters.p, 0, F * sizeof(FrameCounters), s));
    if (job.a.blockList) CK(ctx, cudaMemsetAsync(job.a.blockCount, 0, F * sizeof(uint32_t), s));
    const bool anyVisible = job.anyPlain || job.anyClip;
    const bool tm = ctx->timing && !capture;
    ctx->descDirty = true;
    for (size_t c0 = 0; c0 < F; c0 += (size_t)job.chunk) {
        const int nf = (int)std::min<size_t>((size_t)job.chunk, F - c0);
        DrawArgs a = job.a;
        a.frameObjs += c0 * (size_t)job.nobj;
        a.counters += c0;
        if (a.blockCount) a.blockCount += c0;
        a.color += c0 * job.npix;
        a.depth += c0 * job.npix;
        if (a.tileBusy) a.tileBusy += c0 * (size_t)job.nTiles;
        if (job.overlayKeys) CK(ctx, cudaMemsetAsync(ctx->ovl.p, 0, (size_t)nf * job.npix * sizeof(unsigned long long), s));
        if (tm) CK(ctx, cudaEventRecord(ctx->tev[0], s));
        // K1 only feeds the stage read-back (grb_debug_read_transformed); K2 transforms on its own
        if (anyVisible && ctx->stageCapture) launch_transform(a, nf, s);
        if (tm) CK(ctx, cudaEventRecord(ctx->tev[1], s));
        if (anyVisible) launch_reject(a, nf, s);
        if (anyVisible) launch_setup(a, nf, job.anyPlain, job.anyClip, s);
        if (tm) CK(ctx, cudaEventRecord(ctx->tev[2], s));
        // (slots 2 and 3 of the timing array were the separate bin-scan / bin-fill kernels; binning
        // now happens inside the setup kernel)
        if (tm) CK(ctx, cudaEventRecord(ctx->tev[3], s));
        if (tm) CK(ctx, cudaEventRecord(ctx->tev[4], s));
        launch_raster(a, nf, s);
        if (tm) CK(ctx, cudaEventRecord(ctx->tev[5], s));
        CK(ctx, cudaGetLastError());
        if (tm) {
            CK(ctx, cudaEventSynchronize(ctx->tev[5]));
            for (int k = 0; k < 5; k++) {
                float ms = 0;
                CK(ctx, cudaEventElapsedTime(&ms, ctx->tev[k], ctx->tev[k + 1]));
                ctx->accMs[k] += ms;
            }
        }
    }
    ctx->descDirty = false;   // every setup launch was followed by its raster launch
    if (tm) ctx->accLaunches += launches_of(ctx, job);
    return GRB_OK;
}

int32_t draw_impl(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, const grb_object *objects,
                  int32_t nobj, const grb_draw_params *prm) {
    DrawJob job;
    if (int32_t r = prepare_draw(ctx, fb, frame0, nframes, objects, nobj, prm, false, job)) return r;
    if (fb->pendingRead) {  // do not overwrite frames a read-back is still copying
        CK(ctx, cudaStreamWaitEvent(ctx->stream, fb->readDone, 0));
        fb->pendingRead = false;
    }
    if (int32_t r = enqueue_draw(ctx, job, false)) return r;
    CK(ctx, cudaEventRecord(fb->drawDone, ctx->stream));
    ctx->totalLaunches += launches_of(ctx, job);
    return GRB_OK;
}

// ---- host mirrors (present.cu)

int32_t mirror_args(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, grb_mirror *color,
                    int32_t cf0, grb_mirror *depth, int32_t df0, int32_t rowBeginPx, int32_t rowEndPx, MirrorArgs &m) {
    std::memset(&m, 0, sizeof m);
    if (!ctx || !fb) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (fb->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "framebuffer belongs to another context");
    if (nframes <= 0 || frame0 < 0 || frame0 + nframes > fb->frames)
        return fail(ctx, GRB_ERR_INVALID, "frame range outside the framebuffer");
    const int ntx = (fb->width + kTile - 1) / kTile, nty = (fb->height + kTile - 1) / kTile;
    const size_t nTiles = (size_t)ntx * nty, npix = (size_t)fb->width * fb->height;
    struct { grb_mirror *mr; int32_t f0; int32_t plane; const char *what; } side[2] = {
        {color, cf0, GRB_PLANE_COLOR, "colour"}, {depth, df0, GRB_PLANE_DEPTH, "depth"}};
    for (auto &sd : side) {
        if (!sd.mr) continue;
        if (sd.mr->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "mirror belongs to another context");
        if (sd.mr->plane != sd.plane) return fail(ctx, GRB_ERR_INVALID, std::string("wrong plane type for the ") + sd.what + " mirror");
        if (sd.mr->width != fb->width || sd.mr->height != fb->height)
            return fail(ctx, GRB_ERR_INVALID, "mirror and framebuffer sizes differ");
        if (sd.f0 < 0 || sd.f0 + nframes > sd.mr->frames) return fail(ctx, GRB_ERR_INVALID, "frame range outside the mirror");
    }
    m.color = fb->color + (size_t)frame0 * npix;
    m.depth = fb->depth + (size_t)frame0 * npix;
    m.tileBusy = fb->tileBusy ? fb->tileBusy + (size_t)frame0 * nTiles : nullptr;
    m.full = ((fb->owned || fb->ipcOpened) && fb->tileBusy) ? 0 : 1;   // wrapped memory may have been written by its owner
    if (color) {
        m.hostColor = static_cast<uchar4 *>(color->devPtr) + (size_t)cf0 * npix;
        m.dirtyColor = color->dirty + (size_t)cf0 * nTiles;
        m.tilesWrittenColor = color->tilesWritten;
    }
    if (depth) {
        m.hostDepth = static_cast<float *>(depth->devPtr) + (size_t)df0 * npix;
        m.dirtyDepth = depth->dirty + (size_t)df0 * nTiles;
        m.tilesWrittenDepth = depth->tilesWritten;
    }
    m.width = fb->width;
    m.height = fb->height;
    m.ntx = ntx;
    m.nty = nty;
    m.tileRow0 = 0;
    m.tileRows = nty;
    if (!(rowBeginPx == 0 && rowEndPx == 0)) {
        if (rowBeginPx < 0 || rowEndPx <= rowBeginPx || rowBeginPx % kTile || (rowEndPx % kTile && rowEndPx != fb->height) || rowEndPx > fb->height)
            return fail(ctx, GRB_ERR_INVALID, "row range must be tile-aligned and inside the frame");
        m.tileRow0 = rowBeginPx / kTile;
        m.tileRows = (rowEndPx + kTile - 1) / kTile - m.tileRow0;
    }
    // a mirror that is another framebuffer's plane: keep that framebuffer's own tile flags in step
    grb_framebuffer *target = color ? color->target : (depth ? depth->target : nullptr);
    if (target) {
        if ((color && color->target != target) || (depth && depth->target != target))
            return fail(ctx, GRB_ERR_INVALID, "colour and depth mirrors belong to different framebuffers");
        const int32_t tf0 = color ? cf0 : df0;
        if (color && depth && cf0 != df0) return fail(ctx, GRB_ERR_INVALID, "framebuffer mirrors need the same frame offset for both planes");
        if (target->tileBusy) m.targetBusy = target->tileBusy + (size_t)tf0 * nTiles;
    }
    return GRB_OK;
}

// ---- cached CUDA graphs of one-frame draws

struct GraphKey {
    DrawJob job;
    MirrorArgs m;
    int32_t hasMirror, stageCapture;
    void *countersHost;
};

}  // namespace

struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
};

namespace {

void free_graph(GraphEntry *g) {
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    delete g;
}

constexpr size_t kGraphCache = 8;

}  // namespace

extern "C" {

int32_t grb_abi_version(void) { return GRB_ABI_VERSION; }

const char *grb_last_error(const grb_context *ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int32_t grb_context_create(int32_t device, grb_context **out) {
    if (!out) return fail(nullptr, GRB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, GRB_ERR_CUDA,
                    std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (this library has no CPU path)");
    }
    if (device < 0 || device >= count) return fail(nullptr, GRB_ERR_INVALID, "device ordinal out of range");
    grb_context *ctx = new grb_context;
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, GRB_ERR_CUDA, std::string("context init: ") + cudaGetErrorString(e));
    }
    ctx->stream = ctx->ownStream;
    cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device);
    {
        size_t freeB = 0, totalB = 0;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) ctx->deviceMemory = totalB;
        else cudaGetLastError();
    }
    // read-backs and host-mirror updates get the highest priority: their blocks are few and short, and should
    // not queue behind the render kernels of the next batch for SM slots
    int prLo = 0, prHi = 0;
    cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    if ((e = cudaStreamCreateWithPriority(&ctx->copyStream, cudaStreamNonBlocking, prHi)) != cudaSuccess) {
        cudaStreamDestroy(ctx->ownStream);
        delete ctx;
        return fail(nullptr, GRB_ERR_CUDA, std::string("context init: ") + cudaGetErrorString(e));
    }
    for (int i = 0; i < kStagingRing; i++) cudaEventCreateWithFlags(&ctx->stagingDone[i], cudaEventDisableTiming);
    for (int i = 0; i < 6; i++) cudaEventCreate(&ctx->tev[i]);
    *out = ctx;
    return GRB_OK;
}

int32_t grb_context_destroy(grb_context *ctx) {
    if (!ctx) return GRB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copyStream);
    for (auto &m : ctx->meshes)
        for (void *p : m.allocs) cudaFree(p);
    for (auto &t : ctx->textures)
        if (t.pixels) cudaFree(t.pixels);
    void *bufs[] = {ctx->dMeshes.p, ctx->dTextures.p, ctx->dObjs.p, ctx->dVblk.p, ctx->dFblk.p, ctx->dFrameObjs.p,
                    ctx->tv.p, ctx->rec.p, ctx->uv.p, ctx->warpCount.p, ctx->descCount.p, ctx->desc.p, ctx->overflow.p,
                    ctx->bigList.p, ctx->counters.p, ctx->seam.p, ctx->ovl.p, ctx->blockList.p, ctx->blockCount.p};
    for (void *p : bufs)
        if (p) cudaFree(p);
    for (int i = 0; i < kStagingRing; i++) {
        if (ctx->hFrameObjs[i]) cudaFreeHost(ctx->hFrameObjs[i]);
        if (ctx->stagingDone[i]) cudaEventDestroy(ctx->stagingDone[i]);
    }
    for (GraphEntry *g : ctx->graphs) free_graph(g);
    if (ctx->hGraphObjs) cudaFreeHost(ctx->hGraphObjs);
    if (ctx->hCounters) cudaFreeHost(ctx->hCounters);
    if (ctx->dTimeouts) cudaFree(ctx->dTimeouts);
    for (int i = 0; i < 6; i++)
        if (ctx->tev[i]) cudaEventDestroy(ctx->tev[i]);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    delete ctx;
    return GRB_OK;
}

int32_t grb_context_set_stream(grb_context *ctx, void *cuda_stream) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    if (int32_t r = set_device(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->ownStream;
    return GRB_OK;
}

int32_t grb_context_synchronize(grb_context *ctx) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    if (int32_t r = set_device(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->copyStream));
    return GRB_OK;
}

int32_t grb_context_set_kernel_timing(grb_context *ctx, int32_t enable) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    ctx->timing = enable != 0;
    return GRB_OK;
}

int32_t grb_context_set_stage_capture(grb_context *ctx, int32_t enable) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    ctx->stageCapture = enable != 0;
    return GRB_OK;
}

int32_t grb_kernel_times(grb_context *ctx, double out_ms[5], int64_t *out_launches) {
    if (!ctx || !out_ms) return fail(ctx, GRB_ERR_INVALID, "null argument");
    for (int k = 0; k < 5; k++) { out_ms[k] = ctx->accMs[k]; ctx->accMs[k] = 0; }
    if (out_launches) *out_launches = ctx->accLaunches;
    ctx->accLaunches = 0;
    return GRB_OK;
}

int64_t grb_launch_count(const grb_context *ctx) { return ctx ? ctx->totalLaunches : 0; }

int32_t grb_context_set_workspace_limit(grb_context *ctx, uint64_t bytes) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    ctx->workspaceLimit = bytes;
    return GRB_OK;
}

int32_t grb_context_trim(grb_context *ctx) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    if (int32_t r = set_device(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->copyStream));
    for (GraphEntry *g : ctx->graphs) free_graph(g);   // they hold the workspace addresses
    ctx->graphs.clear();
    auto drop = [](auto &b) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    };
    drop(ctx->tv); drop(ctx->rec); drop(ctx->uv); drop(ctx->warpCount); drop(ctx->descCount); drop(ctx->bigList);
    drop(ctx->desc); drop(ctx->overflow); drop(ctx->ovl); drop(ctx->seam); drop(ctx->blockList);
    ctx->recCap = 0;
    ctx->overflowCap = 0;
    ctx->lastNTiles = 0;
    ctx->lastFrames = 0;
    return GRB_OK;
}

void *grb_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void grb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int32_t grb_host_register(void *p, uint64_t bytes) {
    if (!p || !bytes) return GRB_ERR_INVALID;
    if (cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return GRB_ERR_CUDA;
    }
    return GRB_OK;
}

int32_t grb_host_unregister(void *p) {
    if (!p) return GRB_ERR_INVALID;
    if (cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return GRB_ERR_CUDA;
    }
    return GRB_OK;
}

int32_t grb_texture_upload(grb_context *ctx, int32_t type, int32_t width, int32_t height, float scale,
                           const uint8_t color[4], const uint8_t *pixels, int32_t *out_id) {
    if (!ctx || !out_id) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (int32_t r = set_device(ctx)) return r;
    TexHost t;
    t.dev.type = type;
    t.dev.scale = scale;
    t.dev.color = color ? make_uchar4(color[0], color[1], color[2], color[3]) : make_uchar4(0, 0, 0, 0);
    if (type == GRB_TEX_SOLID) {
        t.dev.width = t.dev.height = 0;
    } else if (type == GRB_TEX_IMAGE || type == GRB_TEX_IMAGE_FAST) {
        if (width <= 0 || height <= 0 || !pixels) return fail(ctx, GRB_ERR_INVALID, "image texture needs pixels");
        if (type == GRB_TEX_IMAGE_FAST && ((width & (width - 1)) || (height & (height - 1))))
            return fail(ctx, GRB_ERR_INVALID, "TextureTypeImageFast needs power-of-two sizes (texture.go:40-43)");
        t.dev.width = width;
        t.dev.height = height;
        const size_t bytes = (size_t)width * height * 4;
        CK(ctx, cudaMalloc(&t.pixels, bytes));
        CK(ctx, cudaMemcpy(t.pixels, pixels, bytes, cudaMemcpyHostToDevice));
        t.dev.pixels = static_cast<const uchar4 *>(t.pixels);
    } else {
        return fail(ctx, GRB_ERR_INVALID, "unknown texture type");
    }
    t.dev.widthF = (float)t.dev.width;   // texture.go:50-51
    t.dev.heightF = (float)t.dev.height;
    ctx->textures.push_back(t);
    ctx->texturesDirty = true;
    *out_id = (int32_t)ctx->textures.size() - 1;
    return GRB_OK;
}

int32_t grb_texture_set_scale(grb_context *ctx, int32_t id, float scale) {
    if (!ctx || id < 0 || id >= (int32_t)ctx->textures.size()) return fail(ctx, GRB_ERR_INVALID, "bad texture id");
    ctx->textures[id].dev.scale = scale;
    ctx->texturesDirty = true;
    return GRB_OK;
}

}  // extern "C"

namespace {

template <typename T> int32_t dev_alloc(grb_context *ctx, size_t n, T **out, std::vector<void *> &allocs) {
    *out = nullptr;
    if (n == 0) return GRB_OK;
    void *p = nullptr;
    CK(ctx, cudaMalloc(&p, n * sizeof(T)));
    allocs.push_back(p);
    *out = static_cast<T *>(p);
    return GRB_OK;
}

// Shared body of grb_mesh_upload / grb_mesh_new.  The source arrays go up as they are; everything
// derived from them — the index check, the face-corner expansion the frame kernels stream and,
// with `derive`, NewMesh's face normals and bounding box — is produced on the device (mesh.cu).
int32_t mesh_upload_impl(grb_context *ctx, const grb_mesh_desc *d, bool derive, int32_t *out_id) {
    if (!ctx || !d || !out_id) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (d->nv <= 0 || d->nf < 0 || d->nvn < 0 || !d->vertices || (d->nf > 0 && (!d->vidx || (!derive && !d->fnormals))))
        return fail(ctx, GRB_ERR_INVALID, "mesh needs vertices, face normals and vertex indices");
    if (d->nvn > 0 && (!d->vnormals || !d->nidx))
        return fail(ctx, GRB_ERR_INVALID, "mesh with vertex normals needs normal indices");
    if (int32_t r = set_device(ctx)) return r;
    MeshHost m;
    m.live = true;
    m.dev.nv = d->nv; m.dev.nvn = d->nvn; m.dev.nf = d->nf;
    std::memcpy(m.bbox, d->bbox, sizeof(m.bbox));
    int32_t r;
    float4 *f4;
    int32_t *i32;
    float2 *f2;
    float4 *warpLo = nullptr, *warpHi = nullptr;
    uint32_t *scratch = nullptr;   // [0..6] bbox keys + NaN bits, [7] index-check flags
    const uint32_t scratchInit[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u};
    uint32_t scratchHost[8];
    MeshPrepArgs pa{};
    cudaStream_t s = ctx->stream;
    if ((r = upload(ctx, reinterpret_cast<const float4 *>(d->vertices), d->nv, &f4, m.allocs))) goto bad;
    m.dev.verts = f4;
    if ((r = upload(ctx, reinterpret_cast<const float4 *>(d->vnormals), d->nvn, &f4, m.allocs))) goto bad;
    m.dev.vnormals = f4;
    if (derive) {
        if ((r = dev_alloc(ctx, (size_t)d->nf, &f4, m.allocs))) goto bad;
    } else {
        if ((r = upload(ctx, reinterpret_cast<const float4 *>(d->fnormals), d->nf, &f4, m.allocs))) goto bad;
    }
    m.dev.fnormals = f4;
    pa.fnormalsOut = derive ? f4 : nullptr;
    if ((r = upload(ctx, d->vidx, (size_t)d->nf * 3, &i32, m.allocs))) goto bad;
    m.dev.vidx = i32;
    if ((r = upload(ctx, d->nvn > 0 ? d->nidx : nullptr, (size_t)d->nf * 3, &i32, m.allocs))) goto bad;
    m.dev.nidx = i32;
    if ((r = upload(ctx, reinterpret_cast<const float2 *>(d->uvs), (size_t)d->nf * 3, &f2, m.allocs))) goto bad;
    m.dev.uvs = f2;
    if ((r = upload(ctx, d->tex, d->nf, &i32, m.allocs))) goto bad;
    m.dev.tex = i32;
    for (int k = 0; k < 3; k++) {   // face-corner expansion (gr_types.cuh): corner k of face f at cv[k][f]
        if ((r = dev_alloc(ctx, (size_t)d->nf, &f4, m.allocs))) goto bad;
        m.dev.cv[k] = f4;
        pa.cv[k] = f4;
        m.dev.cn[k] = nullptr;
        if (d->nvn > 0) {
            if ((r = dev_alloc(ctx, (size_t)d->nf, &f4, m.allocs))) goto bad;
            m.dev.cn[k] = f4;
        }
        pa.cn[k] = const_cast<float4 *>(m.dev.cn[k]);
    }
    {
        const size_t nw = ((size_t)d->nf + 31) / 32;
        if ((r = dev_alloc(ctx, nw, &f4, m.allocs))) goto bad;
        m.dev.warpLo = f4;
        warpLo = f4;
        if ((r = dev_alloc(ctx, nw, &f4, m.allocs))) goto bad;
        m.dev.warpHi = f4;
        warpHi = f4;
    }
    if ((r = dev_alloc(ctx, (size_t)8, &scratch, m.allocs))) goto bad;
    pa.verts = m.dev.verts; pa.vnormals = m.dev.vnormals; pa.vidx = m.dev.vidx; pa.nidx = m.dev.nidx;
    pa.nv = d->nv; pa.nvn = d->nvn; pa.nf = d->nf;
    pa.error = reinterpret_cast<int *>(scratch + 7);
    {
        // cudaMemcpy from pageable memory may return once the data is staged: make sure the uploads above
        // have landed before a kernel on the context's (non-default) stream reads them
        cudaError_t e = cudaStreamSynchronize(cudaStreamLegacy);
        if (e == cudaSuccess) e = cudaMemcpyAsync(scratch, scratchInit, sizeof(scratchInit), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) {
            launch_mesh_prepare(pa, s);
            launch_warp_bounds(m.dev.cv, d->nf, warpLo, warpHi, s);
            if (derive) launch_bbox(m.dev.verts, d->nv, scratch, s);
            ctx->totalLaunches += (d->nf > 0 ? 2 : 0) + (derive ? 1 : 0);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(scratchHost, scratch, sizeof(scratchHost), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            r = fail(ctx, GRB_ERR_CUDA, std::string("mesh preparation: ") + cudaGetErrorString(e));
            goto bad;
        }
    }
    // the reference would panic on an out-of-range index (renderer.go:318-320, 328-330)
    if (scratchHost[7] & 1u) { r = fail(ctx, GRB_ERR_INVALID, "vertex index out of range"); goto bad; }
    if (scratchHost[7] & 2u) {
        r = fail(ctx, GRB_ERR_INVALID, "normal index out of range (the reference panics: renderer.go:328-330)");
        goto bad;
    }
    if (derive) {
        // boundingBox (mesh.go:28-51): corners in min/max order x, then y, then z
        float lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            auto from_key = [](uint32_t key) {
                const uint32_t b = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
                float f;
                std::memcpy(&f, &b, 4);
                return f;
            };
            lo[k] = from_key(scratchHost[k]);
            hi[k] = from_key(scratchHost[3 + k]);
            if (scratchHost[6] & (1u << k)) lo[k] = hi[k] = std::numeric_limits<float>::quiet_NaN();  // Go's min/max propagate NaN
        }
        int c = 0;
        for (int ix = 0; ix < 2; ix++)
            for (int iy = 0; iy < 2; iy++)
                for (int iz = 0; iz < 2; iz++, c++) {
                    m.bbox[4 * c] = ix ? hi[0] : lo[0];
                    m.bbox[4 * c + 1] = iy ? hi[1] : lo[1];
                    m.bbox[4 * c + 2] = iz ? hi[2] : lo[2];
                    m.bbox[4 * c + 3] = 1.0f;
                }
    }
    ctx->meshes.push_back(std::move(m));
    ctx->meshesDirty = true;
    ctx->planValid = false;
    *out_id = (int32_t)ctx->meshes.size() - 1;
    return GRB_OK;
bad:
    for (void *p : m.allocs) cudaFree(p);
    return r;
}

}  // namespace

extern "C" {

int32_t grb_mesh_upload(grb_context *ctx, const grb_mesh_desc *d, int32_t *out_id) {
    return mesh_upload_impl(ctx, d, false, out_id);
}

int32_t grb_mesh_new(grb_context *ctx, const grb_mesh_desc *d, int32_t *out_id) {
    return mesh_upload_impl(ctx, d, true, out_id);
}

int32_t grb_mesh_read_derived(grb_context *ctx, int32_t id, float *fnormals, float bbox[32]) {
    if (!ctx || id < 0 || id >= (int32_t)ctx->meshes.size() || !ctx->meshes[id].live)
        return fail(ctx, GRB_ERR_INVALID, "bad mesh id");
    if (int32_t r = set_device(ctx)) return r;
    const MeshHost &m = ctx->meshes[id];
    if (bbox) std::memcpy(bbox, m.bbox, sizeof(m.bbox));
    if (fnormals && m.dev.nf > 0) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaMemcpy(fnormals, m.dev.fnormals, (size_t)m.dev.nf * sizeof(float4), cudaMemcpyDeviceToHost));
    }
    return GRB_OK;
}

int32_t grb_mesh_free(grb_context *ctx, int32_t id) {
    if (!ctx || id < 0 || id >= (int32_t)ctx->meshes.size() || !ctx->meshes[id].live)
        return fail(ctx, GRB_ERR_INVALID, "bad mesh id");
    if (int32_t r = set_device(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (void *p : ctx->meshes[id].allocs) cudaFree(p);
    ctx->meshes[id] = MeshHost{};
    ctx->meshesDirty = true;
    ctx->planValid = false;
    return GRB_OK;
}

}  // extern "C"

namespace {
// one byte per (frame, tile), 1 = contents unknown / not the cleared background
int32_t alloc_tile_flags(grb_context *ctx, grb_framebuffer *fb) {
    const size_t n = (size_t)((fb->width + kTile - 1) / kTile) * ((fb->height + kTile - 1) / kTile) * fb->frames;
    CK(ctx, cudaMalloc((void **)&fb->tileBusy, n));
    CK(ctx, cudaMemsetAsync(fb->tileBusy, 1, n, ctx->stream));
    return GRB_OK;
}
}  // namespace

extern "C" {

int32_t grb_framebuffer_create(grb_context *ctx, int32_t width, int32_t height, int32_t frames, grb_framebuffer **out) {
    if (!ctx || !out) return fail(ctx, GRB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (width <= 0 || height <= 0 || frames <= 0) return fail(ctx, GRB_ERR_INVALID, "bad framebuffer size");
    if (int32_t r = set_device(ctx)) return r;
    grb_framebuffer *fb = new grb_framebuffer{ctx, width, height, frames, nullptr, nullptr, true};
    const size_t px = (size_t)width * height * frames;
    cudaError_t e = cudaMalloc((void **)&fb->color, px * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&fb->depth, px * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (fb->color) cudaFree(fb->color);
        delete fb;
        return fail(ctx, GRB_ERR_OOM, std::string("framebuffer allocation: ") + cudaGetErrorString(e));
    }
    if (int32_t r = alloc_tile_flags(ctx, fb)) {
        cudaFree(fb->color);
        cudaFree(fb->depth);
        delete fb;
        return r;
    }
    cudaEventCreateWithFlags(&fb->drawDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&fb->readDone, cudaEventDisableTiming);
    *out = fb;
    return GRB_OK;
}

int32_t grb_framebuffer_wrap(grb_context *ctx, int32_t width, int32_t height, int32_t frames, void *device_color,
                             void *device_depth, grb_framebuffer **out) {
    if (!ctx || !out || !device_color || !device_depth) return fail(ctx, GRB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (width <= 0 || height <= 0 || frames <= 0) return fail(ctx, GRB_ERR_INVALID, "bad framebuffer size");
    if (((uintptr_t)device_color & 15) || ((uintptr_t)device_depth & 15))
        return fail(ctx, GRB_ERR_INVALID, "wrapped framebuffer memory must be 16-byte aligned");
    if (int32_t r = set_device(ctx)) return r;
    grb_framebuffer *fb = new grb_framebuffer{ctx, width, height, frames, static_cast<uchar4 *>(device_color),
                                              static_cast<float *>(device_depth), false};
    if (int32_t r = alloc_tile_flags(ctx, fb)) {
        delete fb;
        return r;
    }
    cudaEventCreateWithFlags(&fb->drawDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&fb->readDone, cudaEventDisableTiming);
    *out = fb;
    return GRB_OK;
}

int32_t grb_framebuffer_destroy(grb_framebuffer *fb) {
    if (!fb) return GRB_OK;
    cudaSetDevice(fb->ctx->device);
    cudaStreamSynchronize(fb->ctx->stream);
    cudaStreamSynchronize(fb->ctx->copyStream);
    if (fb->owned) {
        cudaFree(fb->color);
        cudaFree(fb->depth);
    }
    if (fb->ipcOpened) {
        cudaIpcCloseMemHandle(fb->color);
        cudaIpcCloseMemHandle(fb->depth);
        cudaIpcCloseMemHandle(fb->flags);
        cudaIpcCloseMemHandle(fb->tileBusy);
    } else {
        if (fb->tileBusy) cudaFree(fb->tileBusy);
        if (fb->flags) cudaFree(fb->flags);
    }
    // cached graphs may hold this framebuffer's addresses
    for (GraphEntry *g : fb->ctx->graphs) free_graph(g);
    fb->ctx->graphs.clear();
    if (fb->drawDone) cudaEventDestroy(fb->drawDone);
    if (fb->readDone) cudaEventDestroy(fb->readDone);
    delete fb;
    return GRB_OK;
}

int32_t grb_framebuffer_ipc_export(grb_framebuffer *fb, uint8_t handle[GRB_IPC_HANDLE_BYTES]) {
    if (!fb || !handle) return GRB_ERR_INVALID;
    grb_context *ctx = fb->ctx;
    if (!fb->owned || fb->ipcOpened) return fail(ctx, GRB_ERR_INVALID, "only a framebuffer created by grb_framebuffer_create can be exported");
    if (int32_t r = set_device(ctx)) return r;
    if (!fb->flags) {
        CK(ctx, cudaMalloc((void **)&fb->flags, GRB_SIGNAL_SLOTS * kFlagStride * sizeof(uint32_t)));
        CK(ctx, cudaMemset(fb->flags, 0, GRB_SIGNAL_SLOTS * kFlagStride * sizeof(uint32_t)));
    }
    IpcHandle h;
    std::memset(&h, 0, sizeof h);
    CK(ctx, cudaIpcGetMemHandle(&h.color, fb->color));
    CK(ctx, cudaIpcGetMemHandle(&h.depth, fb->depth));
    CK(ctx, cudaIpcGetMemHandle(&h.flags, fb->flags));
    CK(ctx, cudaIpcGetMemHandle(&h.busy, fb->tileBusy));
    h.width = fb->width; h.height = fb->height; h.frames = fb->frames; h.magic = 0x47524231;
    std::memset(handle, 0, GRB_IPC_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof h);
    return GRB_OK;
}

int32_t grb_framebuffer_ipc_open(grb_context *ctx, const uint8_t handle[GRB_IPC_HANDLE_BYTES], grb_framebuffer **out) {
    if (!ctx || !handle || !out) return fail(ctx, GRB_ERR_INVALID, "null argument");
    *out = nullptr;
    IpcHandle h;
    std::memcpy(&h, handle, sizeof h);
    if (h.magic != 0x47524231 || h.width <= 0 || h.height <= 0 || h.frames <= 0)
        return fail(ctx, GRB_ERR_INVALID, "not a framebuffer handle of this library");
    if (int32_t r = set_device(ctx)) return r;
    void *c = nullptr, *d = nullptr, *f = nullptr, *b = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&c, h.color, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&d, h.depth, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&f, h.flags, cudaIpcMemLazyEnablePeerAccess);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&b, h.busy, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (c) cudaIpcCloseMemHandle(c);
        if (d) cudaIpcCloseMemHandle(d);
        if (f) cudaIpcCloseMemHandle(f);
        return fail(ctx, GRB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e) +
                                           " (the exporting process must be another process on the same node)");
    }
    grb_framebuffer *fb = new grb_framebuffer{ctx, h.width, h.height, h.frames, static_cast<uchar4 *>(c), static_cast<float *>(d), false};
    fb->flags = static_cast<uint32_t *>(f);
    fb->tileBusy = static_cast<uint8_t *>(b);     // the owner's per-tile flags: its host mirrors see every rank's rows
    fb->ipcOpened = true;
    cudaEventCreateWithFlags(&fb->drawDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&fb->readDone, cudaEventDisableTiming);
    *out = fb;
    return GRB_OK;
}

int32_t grb_framebuffer_signal(grb_context *ctx, grb_framebuffer *fb, int32_t slot, uint32_t value, int32_t on_copy_stream) {
    if (!ctx || !fb) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (fb->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "framebuffer belongs to another context");
    if (!fb->flags) return fail(ctx, GRB_ERR_STATE, "the framebuffer is not shared (grb_framebuffer_ipc_export / _open)");
    if (slot < 0 || slot >= GRB_SIGNAL_SLOTS) return fail(ctx, GRB_ERR_INVALID, "signal slot out of range");
    if (int32_t r = set_device(ctx)) return r;
    launch_signal(fb->flags + (size_t)slot * kFlagStride, value, on_copy_stream ? ctx->copyStream : ctx->stream);
    CK(ctx, cudaGetLastError());
    ctx->totalLaunches++;
    return GRB_OK;
}

int32_t grb_framebuffer_wait_signals(grb_context *ctx, grb_framebuffer *fb, int32_t slot0, int32_t nslots, uint32_t value,
                                     int32_t timeout_ms, int32_t on_copy_stream) {
    if (!ctx || !fb) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (fb->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "framebuffer belongs to another context");
    if (!fb->flags) return fail(ctx, GRB_ERR_STATE, "the framebuffer is not shared (grb_framebuffer_ipc_export / _open)");
    if (slot0 < 0 || nslots < 0 || slot0 + nslots > GRB_SIGNAL_SLOTS) return fail(ctx, GRB_ERR_INVALID, "signal slots out of range");
    if (nslots == 0) return GRB_OK;
    if (int32_t r = set_device(ctx)) return r;
    if (!ctx->dTimeouts) {
        CK(ctx, cudaMalloc((void **)&ctx->dTimeouts, sizeof(uint32_t)));
        CK(ctx, cudaMemset(ctx->dTimeouts, 0, sizeof(uint32_t)));
    }
    const unsigned long long ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 5000) * 1000000ull;
    launch_wait_signals(fb->flags + (size_t)slot0 * kFlagStride, kFlagStride, nslots, value, ns, ctx->dTimeouts,
                        on_copy_stream ? ctx->copyStream : ctx->stream);
    CK(ctx, cudaGetLastError());
    ctx->totalLaunches++;
    return GRB_OK;
}

int64_t grb_context_signal_timeouts(grb_context *ctx) {
    if (!ctx || !ctx->dTimeouts) return 0;
    if (set_device(ctx)) return -1;
    uint32_t n = 0;
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->copyStream) != cudaSuccess ||
        cudaMemcpy(&n, ctx->dTimeouts, sizeof n, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (int64_t)n;
}

int32_t grb_framebuffer_read_tile_flags(grb_framebuffer *fb, int32_t frame, uint8_t *out) {
    if (!fb || !out) return GRB_ERR_INVALID;
    grb_context *ctx = fb->ctx;
    if (frame < 0 || frame >= fb->frames || !fb->tileBusy) return fail(ctx, GRB_ERR_INVALID, "no such frame");
    if (int32_t r = set_device(ctx)) return r;
    const size_t n = (size_t)((fb->width + kTile - 1) / kTile) * ((fb->height + kTile - 1) / kTile);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(out, fb->tileBusy + (size_t)frame * n, n, cudaMemcpyDeviceToHost));
    return GRB_OK;
}

int32_t grb_framebuffer_device_ptrs(const grb_framebuffer *fb, void **color, void **depth) {
    if (!fb) return GRB_ERR_INVALID;
    if (color) *color = fb->color;
    if (depth) *depth = fb->depth;
    return GRB_OK;
}

int32_t grb_draw_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                       const grb_object *objects, int32_t nobj, const grb_draw_params *params) {
    return draw_impl(ctx, fb, frame0, nframes, objects, nobj, params);
}

int32_t grb_frame_stats_read(grb_context *ctx, int32_t nframes, grb_frame_stats *stats) {
    if (!ctx || !stats || nframes <= 0 || nframes > ctx->lastFrames)
        return fail(ctx, GRB_ERR_INVALID, "no such frames in the last draw");
    if (int32_t r = set_device(ctx)) return r;
    std::vector<FrameCounters> c(nframes);
    CK(ctx, cudaMemcpyAsync(c.data(), ctx->counters.p, nframes * sizeof(FrameCounters), cudaMemcpyDeviceToHost,
                            ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int32_t f = 0; f < nframes; f++) {
        stats[f].tpf = (int64_t)c[f].tpf;
        stats[f].triangles = (int32_t)c[f].triCount;
        stats[f].big_triangles = (int32_t)c[f].bigCount;
        stats[f].out_of_domain = (int32_t)c[f].outOfDomain;
        stats[f].list_fallbacks = (int32_t)c[f].listFallbacks;
    }
    return GRB_OK;
}

int32_t grb_draw(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, const grb_object *objects,
                 int32_t nobj, const grb_draw_params *params, grb_frame_stats *stats) {
    if (int32_t r = draw_impl(ctx, fb, frame0, nframes, objects, nobj, params)) return r;
    if (stats) return grb_frame_stats_read(ctx, nframes, stats);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

int32_t grb_read_frames_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                              uint8_t *pixels, float *zbuffer) {
    if (!ctx || !fb) return fail(ctx, GRB_ERR_INVALID, "null argument");
    if (fb->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "framebuffer belongs to another context");
    if (nframes <= 0 || frame0 < 0 || frame0 + nframes > fb->frames)
        return fail(ctx, GRB_ERR_INVALID, "frame range outside the framebuffer");
    if (int32_t r = set_device(ctx)) return r;
    const size_t px = (size_t)fb->width * fb->height;
    cudaStream_t cs = ctx->copyStream;
    // everything queued on the render stream so far (draws, or the caller's own work on a
    // wrapped framebuffer) completes before the copy starts
    CK(ctx, cudaEventRecord(fb->drawDone, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(cs, fb->drawDone, 0));
    if (pixels) CK(ctx, cudaMemcpyAsync(pixels, fb->color + frame0 * px, nframes * px * 4, cudaMemcpyDeviceToHost, cs));
    if (zbuffer) CK(ctx, cudaMemcpyAsync(zbuffer, fb->depth + frame0 * px, nframes * px * 4, cudaMemcpyDeviceToHost, cs));
    CK(ctx, cudaEventRecord(fb->readDone, cs));
    fb->pendingRead = true;
    return GRB_OK;
}

int32_t grb_read_frames(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, uint8_t *pixels,
                        float *zbuffer) {
    if (int32_t r = grb_read_frames_async(ctx, fb, frame0, nframes, pixels, zbuffer)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->copyStream));
    fb->pendingRead = false;
    return GRB_OK;
}

int32_t grb_framebuffer_wait(grb_framebuffer *fb) {
    if (!fb) return GRB_ERR_INVALID;
    grb_context *ctx = fb->ctx;
    if (int32_t r = set_device(ctx)) return r;
    if (fb->pendingRead) {
        CK(ctx, cudaEventSynchronize(fb->readDone));
        fb->pendingRead = false;
    }
    return GRB_OK;
}

int32_t grb_mirror_create(grb_context *ctx, int32_t width, int32_t height, int32_t frames, int32_t plane, void *host_plane,
                          grb_mirror **out) {
    if (!ctx || !out || !host_plane) return fail(ctx, GRB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (width <= 0 || height <= 0 || frames <= 0) return fail(ctx, GRB_ERR_INVALID, "bad mirror size");
    if (plane != GRB_PLANE_COLOR && plane != GRB_PLANE_DEPTH) return fail(ctx, GRB_ERR_INVALID, "unknown plane type");
    if ((uintptr_t)host_plane & 15) return fail(ctx, GRB_ERR_INVALID, "mirror memory must be 16-byte aligned");
    if (int32_t r = set_device(ctx)) return r;
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, host_plane, 0) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, GRB_ERR_INVALID, "mirror memory must come from grb_host_alloc or be pinned with grb_host_register");
    }
    grb_mirror *m = new grb_mirror{ctx, width, height, frames, plane, dp};
    const size_t n = (size_t)((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile) * frames;
    cudaError_t e = cudaMalloc((void **)&m->dirty, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->tilesWritten, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(m->dirty, 1, n, ctx->stream);   // host contents unknown: first update writes every tile
    if (e == cudaSuccess) e = cudaMemsetAsync(m->tilesWritten, 0, sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (m->dirty) cudaFree(m->dirty);
        if (m->tilesWritten) cudaFree(m->tilesWritten);
        delete m;
        return fail(ctx, GRB_ERR_OOM, std::string("mirror allocation: ") + cudaGetErrorString(e));
    }
    *out = m;
    return GRB_OK;
}

int32_t grb_mirror_create_on_framebuffer(grb_context *ctx, grb_framebuffer *target, int32_t plane, grb_mirror **out) {
    if (!ctx || !out || !target) return fail(ctx, GRB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (target->ctx != ctx) return fail(ctx, GRB_ERR_INVALID, "framebuffer belongs to another context");
    if (plane != GRB_PLANE_COLOR && plane != GRB_PLANE_DEPTH) return fail(ctx, GRB_ERR_INVALID, "unknown plane type");
    if (int32_t r = set_device(ctx)) return r;
    grb_mirror *m = new grb_mirror{ctx, target->width, target->height, target->frames, plane,
                                   plane == GRB_PLANE_COLOR ? (void *)target->color : (void *)target->depth};
    m->target = target;
    const size_t n = (size_t)((m->width + kTile - 1) / kTile) * ((m->height + kTile - 1) / kTile) * m->frames;
    cudaError_t e = cudaMalloc((void **)&m->dirty, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->tilesWritten, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(m->dirty, 1, n, ctx->stream);   // target contents unknown: first update writes every tile
    if (e == cudaSuccess) e = cudaMemsetAsync(m->tilesWritten, 0, sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (m->dirty) cudaFree(m->dirty);
        if (m->tilesWritten) cudaFree(m->tilesWritten);
        delete m;
        return fail(ctx, GRB_ERR_OOM, std::string("mirror allocation: ") + cudaGetErrorString(e));
    }
    *out = m;
    return GRB_OK;
}

int32_t grb_mirror_destroy(grb_mirror *m) {
    if (!m) return GRB_OK;
    grb_context *ctx = m->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copyStream);
    for (GraphEntry *g : ctx->graphs) free_graph(g);   // cached graphs may write into this mirror
    ctx->graphs.clear();
    cudaFree(m->dirty);
    cudaFree(m->tilesWritten);
    if (m->done) cudaEventDestroy(m->done);
    delete m;
    return GRB_OK;
}

int32_t grb_mirror_invalidate(grb_mirror *m) {
    if (!m) return GRB_ERR_INVALID;
    grb_context *ctx = m->ctx;
    if (int32_t r = set_device(ctx)) return r;
    const size_t n = (size_t)((m->width + kTile - 1) / kTile) * ((m->height + kTile - 1) / kTile) * m->frames;
    // ordered after updates already queued on either stream
    CK(ctx, cudaStreamSynchronize(ctx->copyStream));
    CK(ctx, cudaMemsetAsync(m->dirty, 1, n, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

int32_t grb_mirror_update_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, grb_mirror *color,
                                int32_t color_frame0, grb_mirror *depth, int32_t depth_frame0) {
    return grb_mirror_update_rows_async(ctx, fb, frame0, nframes, color, color_frame0, depth, depth_frame0, 0, 0);
}

int32_t grb_mirror_update_rows_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, grb_mirror *color,
                                     int32_t color_frame0, grb_mirror *depth, int32_t depth_frame0, int32_t row_begin,
                                     int32_t row_end) {
    MirrorArgs m;
    if (int32_t r = mirror_args(ctx, fb, frame0, nframes, color, color_frame0, depth, depth_frame0, row_begin, row_end, m)) return r;
    if (!color && !depth) return GRB_OK;
    if (int32_t r = set_device(ctx)) return r;
    cudaStream_t cs = ctx->copyStream;
    CK(ctx, cudaEventRecord(fb->drawDone, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(cs, fb->drawDone, 0));
    for (int32_t f0 = 0; f0 < nframes; f0 += 32768) {   // grid.y limit
        MirrorArgs mm = m;
        const size_t npix = (size_t)fb->width * fb->height, nTiles = (size_t)m.ntx * m.nty;
        const int32_t nf = std::min(nframes - f0, 32768);
        mm.color += f0 * npix; mm.depth += f0 * npix;
        if (mm.tileBusy) mm.tileBusy += f0 * nTiles;
        if (mm.hostColor) { mm.hostColor += f0 * npix; mm.dirtyColor += f0 * nTiles; }
        if (mm.hostDepth) { mm.hostDepth += f0 * npix; mm.dirtyDepth += f0 * nTiles; }
        if (mm.targetBusy) mm.targetBusy += f0 * nTiles;
        launch_mirror_update(mm, nf, cs);
    }
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaEventRecord(fb->readDone, cs));
    fb->pendingRead = true;
    for (grb_mirror *mr : {color, depth})
        if (mr) {
            CK(ctx, cudaEventRecord(mr->done, cs));
            mr->pending = true;
            mr->tilesFull += (int64_t)m.ntx * m.tileRows * nframes;
        }
    ctx->totalLaunches += (nframes + 32767) / 32768;
    return GRB_OK;
}

int32_t grb_mirror_wait(grb_mirror *m) {
    if (!m) return GRB_ERR_INVALID;
    grb_context *ctx = m->ctx;
    if (int32_t r = set_device(ctx)) return r;
    if (m->pending) {
        CK(ctx, cudaEventSynchronize(m->done));
        m->pending = false;
    }
    return GRB_OK;
}

int32_t grb_mirror_stats(grb_mirror *m, int64_t *tiles_written, int64_t *tiles_full) {
    if (!m) return GRB_ERR_INVALID;
    grb_context *ctx = m->ctx;
    if (int32_t r = set_device(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->copyStream));
    unsigned long long w = 0;
    CK(ctx, cudaMemcpy(&w, m->tilesWritten, sizeof w, cudaMemcpyDeviceToHost));
    if (tiles_written) *tiles_written = (int64_t)w;
    if (tiles_full) *tiles_full = m->tilesFull;
    return GRB_OK;
}

int32_t grb_draw_present(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes, const grb_object *objects,
                         int32_t nobj, const grb_draw_params *params, grb_mirror *color, int32_t color_frame0,
                         grb_mirror *depth, int32_t depth_frame0, grb_frame_stats *stats) {
    GraphKey key;
    std::memset(&key, 0, sizeof key);
    if (int32_t r = prepare_draw(ctx, fb, frame0, nframes, objects, nobj, params, true, key.job)) return r;
    key.hasMirror = (color || depth) ? 1 : 0;
    if (key.hasMirror)
        if (int32_t r = mirror_args(ctx, fb, frame0, nframes, color, color_frame0, depth, depth_frame0, 0, 0, key.m)) return r;
    if (nframes > 32768 && key.hasMirror) return fail(ctx, GRB_ERR_INVALID, "at most 32768 frames per grb_draw_present call");
    if (ctx->hCountersCap < (size_t)nframes) {
        if (ctx->hCounters) CK(ctx, cudaFreeHost(ctx->hCounters));
        ctx->hCounters = nullptr;
        ctx->hCountersCap = 0;
        for (GraphEntry *g : ctx->graphs) free_graph(g);   // they copy into the old buffer
        ctx->graphs.clear();
        CK(ctx, cudaHostAlloc((void **)&ctx->hCounters, ((size_t)nframes + 16) * sizeof(FrameCounters), cudaHostAllocDefault));
        ctx->hCountersCap = (size_t)nframes + 16;
    }
    key.countersHost = ctx->hCounters;
    key.stageCapture = ctx->stageCapture ? 1 : 0;
    // One whole frame with mirrors attached: the raster kernel's write-back stores the tiles into the host planes itself
    // (MIRROR instantiation, raster.cu), so that a tile crosses PCIe while the others are still being rasterised; no
    // separate mirror update follows.  (The rule is the mirrors' own; the kernel knows which tiles it leaves busy.)
    const bool fused = key.hasMirror && nframes == 1 && key.job.a.tileRowBegin == 0 && key.job.a.tileRowEnd == key.job.a.nty &&
                       key.m.targetBusy == nullptr;
    if (fused) {
        DrawArgs &a = key.job.a;
        a.mirColor = key.m.hostColor;
        a.mirDepth = key.m.hostDepth;
        a.mirDirtyColor = key.m.dirtyColor;
        a.mirDirtyDepth = key.m.dirtyDepth;
        a.mirWrittenColor = key.m.tilesWrittenColor;
        a.mirWrittenDepth = key.m.tilesWrittenDepth;
    }
    cudaStream_t s = ctx->stream;
    if (fb->pendingRead) {  // a read-back / mirror update of these frames on the copy stream
        CK(ctx, cudaStreamWaitEvent(s, fb->readDone, 0));
        fb->pendingRead = false;
    }
    for (grb_mirror *mr : {color, depth})
        if (mr && mr->pending) {   // an update of the same mirror still running on the copy stream
            CK(ctx, cudaStreamWaitEvent(s, mr->done, 0));
            mr->pending = false;
        }
    auto enqueue_all = [&](bool capture) -> int32_t {
        if (int32_t r = enqueue_draw(ctx, key.job, capture)) return r;
        if (key.hasMirror && !fused) launch_mirror_update(key.m, nframes, s);
        CK(ctx, cudaMemcpyAsync(ctx->hCounters, ctx->counters.p, (size_t)nframes * sizeof(FrameCounters), cudaMemcpyDeviceToHost, s));
        CK(ctx, cudaGetLastError());
        return GRB_OK;
    };
    const int launches = launches_of(ctx, key.job) + ((key.hasMirror && !fused) ? 1 : 0);
    const bool graphable = nframes == 1 && !ctx->timing;
    if (graphable) {
        GraphEntry *hit = nullptr;
        for (size_t i = 0; i < ctx->graphs.size(); i++)
            if (std::memcmp(&ctx->graphs[i]->key, &key, sizeof key) == 0) {
                hit = ctx->graphs[i];
                ctx->graphs.erase(ctx->graphs.begin() + i);
                break;
            }
        if (!hit) {
            cudaGraph_t graph = nullptr;
            CK(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            const int32_t r = enqueue_all(true);
            const cudaError_t e = cudaStreamEndCapture(s, &graph);
            ctx->descDirty = false;   // nothing has run yet
            if (r != GRB_OK) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                return r;
            }
            if (e != cudaSuccess) return fail(ctx, GRB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
            hit = new GraphEntry;
            std::memcpy(&hit->key, &key, sizeof key);   // byte image, padding included: the cache compares with memcmp
            hit->launches = launches;
            const cudaError_t ei = cudaGraphInstantiate(&hit->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) {
                delete hit;
                return fail(ctx, GRB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei));
            }
            ctx->graphCaptures++;
        } else {
            ctx->graphReplays++;
        }
        ctx->graphs.insert(ctx->graphs.begin(), hit);
        while (ctx->graphs.size() > kGraphCache) {
            free_graph(ctx->graphs.back());
            ctx->graphs.pop_back();
        }
        ctx->descDirty = true;
        CK(ctx, cudaGraphLaunch(hit->exec, s));
    } else {
        if (int32_t r = enqueue_all(false)) return r;
    }
    CK(ctx, cudaStreamSynchronize(s));
    ctx->descDirty = false;
    ctx->totalLaunches += launches;
    for (grb_mirror *mr : {color, depth})
        if (mr) mr->tilesFull += (int64_t)key.m.ntx * key.m.nty * nframes;
    if (stats)
        for (int32_t f = 0; f < nframes; f++) {
            const FrameCounters &c = ctx->hCounters[f];
            stats[f].tpf = (int64_t)c.tpf;
            stats[f].triangles = (int32_t)c.triCount;
            stats[f].big_triangles = (int32_t)c.bigCount;
            stats[f].out_of_domain = (int32_t)c.outOfDomain;
            stats[f].list_fallbacks = (int32_t)c.listFallbacks;
        }
    return GRB_OK;
}

int64_t grb_graph_replays(const grb_context *ctx) { return ctx ? ctx->graphReplays : 0; }

int32_t grb_debug_set_overflow_cap(grb_context *ctx, uint32_t entries) {
    if (!ctx) return fail(ctx, GRB_ERR_INVALID, "null context");
    if (int32_t r = set_device(ctx)) return r;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->overflow.p) cudaFree(ctx->overflow.p);   // re-grown at the next draw for the new stride
    ctx->overflow.p = nullptr;
    ctx->overflow.cap = 0;
    ctx->overflowCap = 0;
    ctx->overflowCapForced = entries;
    return GRB_OK;
}

int32_t grb_matrix_multiply_vec4_batch_device(grb_context *ctx, const float m[16], void *device_vecs, int64_t n) {
    if (!ctx || !m || (n > 0 && !device_vecs) || n < 0) return fail(ctx, GRB_ERR_INVALID, "bad argument");
    if (int32_t r = set_device(ctx)) return r;
    launch_matvec_batch(m, static_cast<float4 *>(device_vecs), n, ctx->stream);
    if (n > 0) ctx->totalLaunches++;
    CK(ctx, cudaGetLastError());
    return GRB_OK;
}

int32_t grb_matrix_multiply_vec4_batch(grb_context *ctx, const float m[16], float *vecs, int64_t n) {
    if (!ctx || !m || (n > 0 && !vecs) || n < 0) return fail(ctx, GRB_ERR_INVALID, "bad argument");
    if (n == 0) return GRB_OK;  // asm_amd64.s:15-17: empty slice is a no-op
    if (int32_t r = set_device(ctx)) return r;
    if (int32_t r = ensure(ctx, ctx->seam, (size_t)n, false)) return r;
    CK(ctx, cudaMemcpyAsync(ctx->seam.p, vecs, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    if (int32_t r = grb_matrix_multiply_vec4_batch_device(ctx, m, ctx->seam.p, n)) return r;
    CK(ctx, cudaMemcpyAsync(vecs, ctx->seam.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

int32_t grb_debug_read_transformed(grb_context *ctx, int32_t frame, float *out, int64_t capacity_vec4, int64_t *out_n) {
    if (!ctx || frame < 0 || frame >= ctx->lastFrames) return fail(ctx, GRB_ERR_INVALID, "no such frame in the last draw");
    if (!ctx->stageCapture) return fail(ctx, GRB_ERR_STATE, "enable grb_context_set_stage_capture before the draw");
    if (int32_t r = set_device(ctx)) return r;
    const int64_t n = ctx->totalVerts;
    if (out_n) *out_n = n;
    if (!out) return GRB_OK;
    if (capacity_vec4 < n) return fail(ctx, GRB_ERR_INVALID, "output too small");
    if (n) CK(ctx, cudaMemcpyAsync(out, ctx->tv.p + (size_t)frame * n, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return GRB_OK;
}

int32_t grb_debug_read_triangles(grb_context *ctx, int32_t frame, grb_triangle_rec *out, float *out_uvs,
                                 int64_t capacity, int64_t *out_n) {
    if (!ctx || frame < 0 || frame >= ctx->lastFrames) return fail(ctx, GRB_ERR_INVALID, "no such frame in the last draw");
    if (!ctx->stageCapture) return fail(ctx, GRB_ERR_STATE, "enable grb_context_set_stage_capture before the draw");
    if (int32_t r = set_device(ctx)) return r;
    FrameCounters c;
    CK(ctx, cudaMemcpyAsync(&c, ctx->counters.p + frame, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t n = c.triCount;
    if (out_n) *out_n = n;
    if (!out) return GRB_OK;
    if (capacity < n) return fail(ctx, GRB_ERR_INVALID, "output too small");
    if (n == 0) return GRB_OK;
    // records sit in per-warp segments in submission order (setup.cu); walk them with the
    // per-warp slot counts
    const int32_t nobj = ctx->lastNobj;
    const size_t nWarps = (size_t)ctx->nFaceBlocks * kWarpsPerFaceBlock;
    std::vector<uint32_t> wc(nWarps);
    std::vector<FrameObj> fo(std::max(nobj, 1));
    CK(ctx, cudaMemcpy(wc.data(), ctx->warpCount.p + (size_t)frame * nWarps, nWarps * 4, cudaMemcpyDeviceToHost));
    if (nobj) CK(ctx, cudaMemcpy(fo.data(), ctx->dFrameObjs.p + (size_t)frame * nobj, nobj * sizeof(FrameObj), cudaMemcpyDeviceToHost));
    const bool optClip = ctx->lastOptions & GRB_OPT_FRUSTUM_CLIPPING;
    int64_t k = 0;
    std::vector<PackedRec> seg;
    std::vector<TriUV> seguv;
    for (int32_t i = 0; i < nobj; i++) {
        if (fo[i].visibility == GRB_BOX_OUTSIDE) continue;
        const bool clips = optClip && fo[i].visibility != GRB_BOX_INSIDE;
        const uint32_t perWarp = clips ? kWarpSlotsClip : kWarpSlots;
        const DrawObj &ob = ctx->planObjs[i];
        const int32_t nf = ctx->meshes[ob.mesh].dev.nf;
        const int32_t blocks = (nf + kFaceBlock - 1) / kFaceBlock;
        for (int64_t w = 0; w < (int64_t)blocks * kWarpsPerFaceBlock; w++) {
            const uint32_t cnt = wc[(size_t)ob.faceBlockBase * kWarpsPerFaceBlock + w];
            if (cnt == 0) continue;
            const size_t slot = (size_t)frame * ctx->recCap + fo[i].slotBase + (size_t)w * perWarp;
            seg.resize(cnt);
            seguv.resize(cnt);
            CK(ctx, cudaMemcpy(seg.data(), ctx->rec.p + slot, cnt * sizeof(PackedRec), cudaMemcpyDeviceToHost));
            CK(ctx, cudaMemcpy(seguv.data(), ctx->uv.p + slot, cnt * sizeof(TriUV), cudaMemcpyDeviceToHost));
            for (uint32_t j = 0; j < cnt; j++) {
                if (seg[j].bx1 < seg[j].bx0) continue;  // slot of a face that survived the cull but draws nothing
                if (k >= n) return fail(ctx, GRB_ERR_STATE, "record walk disagrees with the triangle counter");
                const PackedRec &pr = seg[j];   // storage form -> the ABI's record
                out[k] = grb_triangle_rec{pr.x0, pr.y0, pr.x1, pr.y1, pr.x2, pr.y2, pr.w0, pr.w1, pr.w2, pr.i0, pr.i1, pr.i2,
                                          pr.bx0, pr.by0, pr.bx1, pr.by1, pr.tex,
                                          (uint32_t)(fo[i].slotBase + (size_t)w * perWarp + j)};  // slot == order key
                if (out_uvs) {
                    if (seg[j].tex >= 0) std::memcpy(out_uvs + 6 * k, &seguv[j], 24);
                    else std::memset(out_uvs + 6 * k, 0, 24);
                }
                k++;
            }
        }
    }
    if (k != n) return fail(ctx, GRB_ERR_STATE, "record walk disagrees with the triangle counter");
    return GRB_OK;
}

int32_t grb_debug_read_visibility(grb_context *ctx, int32_t frame, int32_t *out, int32_t capacity) {
    if (!ctx || !out || frame < 0 || frame >= ctx->lastFrames) return fail(ctx, GRB_ERR_INVALID, "no such frame in the last draw");
    const int32_t n = std::min(capacity, ctx->lastNobj);
    for (int32_t i = 0; i < n; i++) out[i] = ctx->lastVisibility[(size_t)frame * ctx->lastNobj + i];
    return GRB_OK;
}

}  // extern "C"
