/*
 * gorender_b200.h — C ABI of the B200-native gorender hot path.
 *
 * The reference (maxpoletaev/gorender) has no plugin / FFI layer: everything
 * is `package main`.  The boundary this library sits behind is therefore the
 * reference's Go API for the per-frame path — `NewFrameBuffer`
 * (rasterizer.go:15), `NewRenderer` (renderer.go:114), `(*Renderer).Draw`
 * (renderer.go:443), plus its one build-tag seam `matrixMultiplyVec4Batch`
 * (asm_amd64.go:8 / asm_purego.go:9).  Each entry point below names the
 * reference interface it replaces.  INTEGRATION.md shows the cgo binding.
 *
 * Conventions
 *   - plain pointers and sizes only; all matrices are row-major float[16]
 *     (reference `Matrix [4][4]float32`, matrix.go:3); vectors are xyzw
 *     float[4] (`Vec4`, vector.go:87-89); colours are RGBA8 (`color.RGBA`);
 *   - every call returns GRB_OK (0) or an error code; `grb_last_error`
 *     returns the message.  The reference's `Draw` cannot fail
 *     (renderer.go:443), so a Go shim turns non-zero into panic;
 *   - entries may be called from any OS thread (each sets its device), one
 *     in-flight call per context;
 *   - host pointers are only read/written during the call; nothing
 *     caller-owned is retained (cgo pointer rules);
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     fails with GRB_ERR_CUDA.
 */
#ifndef GORENDER_B200_H
#define GORENDER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRB_ABI_VERSION 3

enum {
    GRB_OK = 0,
    GRB_ERR_INVALID = 1,   /* bad argument / handle / index            */
    GRB_ERR_CUDA = 2,      /* CUDA runtime error (message has details) */
    GRB_ERR_OOM = 3,       /* device or pinned-host allocation failed  */
    GRB_ERR_STATE = 4      /* call not valid in the current state      */
};

/* Renderer option bits == the reference's Renderer bool fields
 * (renderer.go:90-97; defaults renderer.go:130-137). */
enum {
    GRB_OPT_FRUSTUM_CLIPPING = 1u << 0,
    GRB_OPT_SHOW_FACES       = 1u << 1,
    GRB_OPT_BACKFACE_CULLING = 1u << 2,
    GRB_OPT_LIGHTING         = 1u << 3,
    GRB_OPT_FLAT_SHADING     = 1u << 4,
    GRB_OPT_SHOW_TEXTURES    = 1u << 5,
    /* Overlays of drawProjection (renderer.go:191-216): wireframe edges + face-centre marks
     * (FrameBuffer.Line / Rect, rasterizer.go:54-79) and 3x3 vertex marks. */
    GRB_OPT_SHOW_EDGES       = 1u << 6,
    GRB_OPT_SHOW_VERTICES    = 1u << 7,
    /* Post passes of Draw (renderer.go:476-480).  The reference compiles them in or out
     * (`!demoMode`, main.go:22; the Fog call is commented out); here they are option bits. */
    GRB_OPT_CROSSHAIR        = 1u << 8,   /* FrameBuffer.CrossHair({255,255,0,255}), rasterizer.go:209-217 */
    GRB_OPT_FOG              = 1u << 9,   /* FrameBuffer.Fog(fog_start, fog_end, fog_color), rasterizer.go:193-207 */
    /* Affine texture mapping.  NOT a reference code path: README.md:46 lists it and BASELINE.json's north_star names it,
     * but rasterizer.go:158-159 (perspective-correct) is the only interpolation the reference has, so there is nothing
     * to be identical to — this mode is UNPINNED by construction.  Definition used here: the reference's expression with
     * the 1/w factors and the division by zRec removed, i.e. linear in screen space with the reference's own (biased)
     * barycentric weights, u = -((alpha*u0 + beta*u1) + gamma*u2), v likewise; the sign keeps Texture.Sample's
     * convention (the reference's u, v come out negated).  Off by default; depth is unaffected. */
    GRB_OPT_AFFINE_TEXTURES  = 1u << 10,
    GRB_OPT_DEFAULT = GRB_OPT_FRUSTUM_CLIPPING | GRB_OPT_SHOW_FACES |
                      GRB_OPT_BACKFACE_CULLING | GRB_OPT_LIGHTING |
                      GRB_OPT_SHOW_TEXTURES
};

/* TextureType (texture.go:11-15). */
enum { GRB_TEX_SOLID = 0, GRB_TEX_IMAGE = 1, GRB_TEX_IMAGE_FAST = 2 };

/* BoxVisibility (clipping.go:23-27). */
enum { GRB_BOX_OUTSIDE = 0, GRB_BOX_INTERSECT = 1, GRB_BOX_INSIDE = 2 };

typedef struct grb_context grb_context;
typedef struct grb_framebuffer grb_framebuffer;

/* Flattened `Mesh` (mesh.go:19-26) + `[]Face` (mesh.go:12-17): a Go `Face`
 * holds a Go pointer and 64-bit ints and cannot cross cgo, so the shim
 * flattens it once at upload. */
typedef struct grb_mesh_desc {
    int32_t nv, nvn, nf;
    const float *vertices;   /* nv  * 4  Mesh.Vertices      (x,y,z,1)            */
    const float *vnormals;   /* nvn * 4  Mesh.VertexNormals (x,y,z,1), may be NULL */
    const float *fnormals;   /* nf  * 4  Mesh.FaceNormals   (x,y,z,1)            */
    const int32_t *vidx;     /* nf * 3   Face.VertexIndices                      */
    const int32_t *nidx;     /* nf * 3   Face.NormalIndices (NULL if nvn == 0)   */
    const float *uvs;        /* nf * 6   Face.UVs (u0,v0,u1,v1,u2,v2), may be NULL */
    const int32_t *tex;      /* nf       texture id per face, -1 = nil; may be NULL */
    float bbox[32];          /* Mesh.BoundingBox: 8 corners * xyzw (mesh.go:41-50) */
} grb_mesh_desc;

/* One object of one frame: mesh handle + the finished matrices the unchanged
 * host code computes (renderer.go:255-262), so the host libm never has to be
 * matched on the device. */
typedef struct grb_object {
    int32_t mesh;
    float world[16];         /* NewWorldMatrix(Scale, Rotation, Translation) */
    float mvp[16];           /* ((I * perspective) * view) * world           */
} grb_object;

typedef struct grb_draw_params {
    float screen[16];        /* NewScreenMatrix(width, height)  (renderer.go:264) */
    float light[3];          /* normalize(-1, 1, 1)             (renderer.go:265) */
    uint32_t options;        /* GRB_OPT_*                                         */
    float z_near, z_far;     /* frustum planes (renderer.go:121-122: 0, 50)       */
    int32_t ref_tiles;       /* the reference's numTiles: 16 (parallel) or 1; decides
                                TPF and the tile-list membership rule (renderer.go:226-244) */
    int32_t row_begin;       /* sort-first strip: rasterise rows [row_begin,row_end); */
    int32_t row_end;         /*   both multiples of GRB_TILE; 0,0 = whole frame.  Geometry that cannot reach the rows is
                                skipped (object, then per 32 faces), and a triangle's TPF is counted by the strip that owns
                                its top row: the TPFs of the strips of a frame add up to the frame's                      */
    float fog_start, fog_end;/* GRB_OPT_FOG: Fog(fogStart, fogEnd, c); renderer.go:479 has 0.100, 0.033 */
    uint8_t fog_color[4];    /*   ... and {100,100,100,255}                                */
} grb_draw_params;

#define GRB_TILE 32          /* device raster tile edge, pixels */

/* Device record of one emitted screen-space triangle (the reference's
 * `Triangle`, renderer.go:29-34, after the integer snap of renderer.go:182-184). */
typedef struct grb_triangle_rec {
    int32_t x0, y0, x1, y1, x2, y2;   /* int(a.X), int(a.Y) ...            */
    float w0, w1, w2;                 /* clip-space w (the rasteriser's z) */
    float i0, i1, i2;                 /* vertex intensities                */
    int16_t bx0, by0, bx1, by1;       /* inclusive raster bbox after the tile-list rule */
    int32_t tex;                      /* texture id, -1 = face colour      */
    uint32_t order;                   /* record slot == submission-order key      */
} grb_triangle_rec;

/* Per-frame statistics. */
typedef struct grb_frame_stats {
    int64_t tpf;             /* Renderer.TPF (renderer.go:436-441)                  */
    int32_t triangles;       /* emitted triangles that reached the rasteriser       */
    int32_t big_triangles;   /* of those, spanning more than 16 device tiles (frame-wide list) */
    int32_t out_of_domain;   /* triangles dropped: snapped |coord| > 16383 or NaN   */
    int32_t list_fallbacks;  /* triangles that went to the frame-wide list because their tile's list and the
                                overflow pool were full (slower, never wrong)                             */
} grb_frame_stats;

/* ---- context ------------------------------------------------------------ */

int32_t grb_abi_version(void);
/* Message of the last failing call on `ctx` (or of the last failing
 * grb_context_create when ctx == NULL).  Never NULL. */
const char *grb_last_error(const grb_context *ctx);

int32_t grb_context_create(int32_t device, grb_context **out);
int32_t grb_context_destroy(grb_context *ctx);
/* Run all work of `ctx` on this cudaStream_t (NULL = the context's own stream). */
int32_t grb_context_set_stream(grb_context *ctx, void *cuda_stream);
int32_t grb_context_synchronize(grb_context *ctx);
/* When enabled, every draw records CUDA events between its kernels and
 * accumulates per-kernel time (ms) retrievable with grb_kernel_times. */
int32_t grb_context_set_kernel_timing(grb_context *ctx, int32_t enable);
/* When enabled, draws also run the standalone per-vertex transform kernel and keep its output
 * (Object.TransformedVertices, mesh.go:76) for grb_debug_read_transformed.  Off by default: the
 * frame path fuses the transform into the setup kernel and never materialises that array. */
int32_t grb_context_set_stage_capture(grb_context *ctx, int32_t enable);
/* out_ms[0..4] = transform (stage capture only), setup (incl. binning), 0, 0, raster; out_launches =
 * number of kernel launches accumulated.  Resets the accumulators. */
int32_t grb_kernel_times(grb_context *ctx, double out_ms[5], int64_t *out_launches);
/* Total kernel launches issued by this context since creation. */
int64_t grb_launch_count(const grb_context *ctx);

/* Per-draw workspace (records, tile lists) is sized per frame of a batch; a batch whose workspace would
 * exceed `bytes` is rendered in several launches of fewer frames instead of failing (default: a quarter of
 * the device's memory).  grb_context_trim frees the workspace (it is re-grown by the next draw). */
int32_t grb_context_set_workspace_limit(grb_context *ctx, uint64_t bytes);
int32_t grb_context_trim(grb_context *ctx);

/* Pinned, device-mapped host memory for read-back targets and mirrors (cudaHostAlloc). */
void *grb_host_alloc(uint64_t bytes);
void grb_host_free(void *p);
/* Pin caller-owned host memory in place (cudaHostRegister) so that it can back a mirror or an async
 * read-back; e.g. the Go slices behind FrameBuffer.Pixels (Go's collector does not move heap objects). */
int32_t grb_host_register(void *p, uint64_t bytes);
int32_t grb_host_unregister(void *p);

/* ---- assets (replaces nothing on the hot path: one-time upload of what
 *      LoadObjFile / NewImageTexture produced; texture.go:19-63, mesh.go:53-69) */

int32_t grb_texture_upload(grb_context *ctx, int32_t type, int32_t width, int32_t height,
                           float scale, const uint8_t color[4], const uint8_t *pixels,
                           int32_t *out_id);
/* Texture.SetScale (texture.go:65-67). */
int32_t grb_texture_set_scale(grb_context *ctx, int32_t id, float scale);
int32_t grb_mesh_upload(grb_context *ctx, const grb_mesh_desc *desc, int32_t *out_id);
/* NewMesh (mesh.go:53-69) on the device: like grb_mesh_upload, but Mesh.FaceNormals (mesh.go:54-60)
 * and Mesh.BoundingBox (mesh.go:28-51) are computed on the GPU from the uploaded vertices and
 * indices; desc->fnormals and desc->bbox are ignored.  grb_mesh_read_derived copies them back for
 * the caller's Mesh fields (fnormals: nf*4 floats; bbox: 8 corners * xyzw); either may be NULL. */
int32_t grb_mesh_new(grb_context *ctx, const grb_mesh_desc *desc, int32_t *out_id);
int32_t grb_mesh_read_derived(grb_context *ctx, int32_t id, float *fnormals, float bbox[32]);
int32_t grb_mesh_free(grb_context *ctx, int32_t id);

/* ---- LoadObjFile's parsing (obj.go:196-309) as native host code: one pass over the file in memory,
 *      strtof-grade float32, the reference's four face syntaxes, per-object index offsets and the
 *      `v//vn` quirk (obj.go:77-89).  No CUDA involved (works without a device).  Textures are
 *      not decoded here (image.Decode stays the caller's, texture.go:91-103): faces carry an index
 *      into the file's table of texture sources — path "" is the default solid {255,0,255,255}
 *      texture of materials without map_Kd (obj.go:208), -1 the nil texture.  grb_obj_mesh fills a
 *      grb_mesh_desc whose arrays stay valid until grb_obj_free; fnormals is NULL and bbox unset:
 *      hand it to grb_mesh_new (NewMesh on the device) or derive them on the host. */
typedef struct grb_obj grb_obj;
int32_t grb_obj_parse(const char *filename, int32_t single_mesh, grb_obj **out, char *err, int32_t err_cap);
int32_t grb_obj_num_meshes(const grb_obj *obj);
int32_t grb_obj_num_textures(const grb_obj *obj);
const char *grb_obj_texture_path(const grb_obj *obj, int32_t i);
int32_t grb_obj_mesh(const grb_obj *obj, int32_t i, grb_mesh_desc *desc);
void grb_obj_free(grb_obj *obj);

/* ---- framebuffer: NewFrameBuffer (rasterizer.go:15-23) ------------------
 * `frames` device frames of RGBA8 colour + f32 depth each (frames > 1 for
 * frame-parallel batches).  Clear + DotGrid (rasterizer.go:36-52) are
 * generated inside the raster kernel; there is no separate clear pass. */
int32_t grb_framebuffer_create(grb_context *ctx, int32_t width, int32_t height, int32_t frames,
                               grb_framebuffer **out);
/* Same, over caller-owned device memory (e.g. torch tensors): colour
 * frames*H*W*4 bytes, depth frames*H*W floats. */
int32_t grb_framebuffer_wrap(grb_context *ctx, int32_t width, int32_t height, int32_t frames,
                             void *device_color, void *device_depth, grb_framebuffer **out);
int32_t grb_framebuffer_destroy(grb_framebuffer *fb);
int32_t grb_framebuffer_device_ptrs(const grb_framebuffer *fb, void **color, void **depth);
/* One byte per GRB_TILE x GRB_TILE tile of `frame`, row-major: 0 = the tile holds only the cleared background
 * (written by the raster kernel; 1 for tiles no draw has touched yet).  Synchronises the render stream. */
int32_t grb_framebuffer_read_tile_flags(grb_framebuffer *fb, int32_t frame, uint8_t *out);

/* ---- a framebuffer shared by the processes of one node (one per GPU): the exchange step of the sort-first
 *      screen strips of a single frame.  The owner (rank 0) exports a handle (cudaIpc*), the other ranks open it.
 *      A rank renders its rows (row_begin / row_end) into a framebuffer of its own and pushes them into the owner's
 *      with a mirror whose plane IS the opened framebuffer (grb_mirror_create_on_framebuffer +
 *      grb_mirror_update_rows_async): only tiles that are busy, or were busy in the owner's copy, cross NVLink, on
 *      the copy stream, while the render stream sets up the next frames.  (Drawing straight into the opened
 *      framebuffer also works — the raster kernel's stores then go to peer memory — but every rank's stores then
 *      converge on the owner's NVLink ingress at the same moment, the end of its raster kernel.)  Hand-off is on
 *      the device: 64 flag words live next to the framebuffer; grb_framebuffer_signal raises flag `slot` to `value`
 *      behind everything queued on the caller's render stream (or copy stream), grb_framebuffer_wait_signals makes
 *      that stream wait until flags [slot0, slot0 + nslots) have all reached `value` (flags only grow).  A wait
 *      that sees no progress for `timeout_ms` gives up and is counted (grb_context_signal_timeouts) instead of
 *      hanging the GPU. */
#define GRB_IPC_HANDLE_BYTES 320
#define GRB_SIGNAL_SLOTS 64
int32_t grb_framebuffer_ipc_export(grb_framebuffer *fb, uint8_t handle[GRB_IPC_HANDLE_BYTES]);
int32_t grb_framebuffer_ipc_open(grb_context *ctx, const uint8_t handle[GRB_IPC_HANDLE_BYTES], grb_framebuffer **out);
int32_t grb_framebuffer_signal(grb_context *ctx, grb_framebuffer *fb, int32_t slot, uint32_t value, int32_t on_copy_stream);
int32_t grb_framebuffer_wait_signals(grb_context *ctx, grb_framebuffer *fb, int32_t slot0, int32_t nslots, uint32_t value,
                                     int32_t timeout_ms, int32_t on_copy_stream);
int64_t grb_context_signal_timeouts(grb_context *ctx);

/* ---- the hot path: (*Renderer).Draw (renderer.go:443-483) ----------------
 * Renders `nframes` frames into fb frames [frame0, frame0+nframes).  Every
 * frame draws the same `nobj` meshes; objects[f*nobj + i] carries frame f's
 * matrices for object i (frame-parallel pose batches: SURVEY.md §8e).
 * Asynchronous on the context stream. */
int32_t grb_draw_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                       const grb_object *objects, int32_t nobj, const grb_draw_params *params);
/* Synchronous single-call form: draw, wait, return per-frame stats
 * (stats may be NULL; else nframes entries). */
int32_t grb_draw(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                 const grb_object *objects, int32_t nobj, const grb_draw_params *params,
                 grb_frame_stats *stats);
/* Stats of the frames of the most recent draw (synchronises). */
int32_t grb_frame_stats_read(grb_context *ctx, int32_t nframes, grb_frame_stats *stats);

/* FrameBuffer.Pixels / FrameBuffer.ZBuffer (rasterizer.go:7-13): copy frames
 * [frame0, frame0+nframes) to host.  Either pointer may be NULL.  The async
 * form needs pinned memory (grb_host_alloc) to overlap with rendering. */
int32_t grb_read_frames(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                        uint8_t *pixels, float *zbuffer);
/* Copies run on the context's own copy stream, ordered after everything
 * queued so far on the render stream; a later draw into the same framebuffer
 * waits for them.  With two framebuffers, the read-back of one overlaps the
 * rendering of the other.  grb_context_synchronize waits for both streams. */
int32_t grb_read_frames_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                              uint8_t *pixels, float *zbuffer);

/* Blocks until the most recent grb_read_frames_async of `fb` has landed in host memory (returns
 * at once if there is none).  Together with grb_draw_async this is the streaming form of the
 * reference's render / present double buffer (main.go:198-227, rasterizer.go:32-34): draw frame
 * i+1 into a second framebuffer while frame i is still crossing PCIe. */
int32_t grb_framebuffer_wait(grb_framebuffer *fb);

/* ---- host mirrors: FrameBuffer.Pixels / Pixels2 / ZBuffer (rasterizer.go:7-13) as host memory kept in
 *      sync tile by tile.  The reference clears the whole host framebuffer at the start of every Draw
 *      (rasterizer.go:36-52) and leaves most of it at that background; a mirror remembers, per 32x32
 *      tile, whether the HOST plane holds anything else, the raster kernel records the same per tile of
 *      every DEVICE frame, and an update writes only the tiles that are busy now or were busy in the
 *      host copy — straight into the pinned plane over PCIe (no staging, no host scatter).  After the
 *      update the plane is byte for byte what grb_read_frames would have delivered.
 *      `host_plane`: frames * height * width * 4 bytes of memory from grb_host_alloc or pinned with
 *      grb_host_register; its previous contents are unknown to the mirror (the first update of each frame
 *      writes every tile).  Code that writes into the plane itself must call grb_mirror_invalidate. */
enum { GRB_PLANE_COLOR = 0, GRB_PLANE_DEPTH = 1 };
typedef struct grb_mirror grb_mirror;
int32_t grb_mirror_create(grb_context *ctx, int32_t width, int32_t height, int32_t frames, int32_t plane,
                          void *host_plane, grb_mirror **out);
/* A mirror whose plane is a plane of `target` (normally a framebuffer opened with grb_framebuffer_ipc_open):
 * updating it pushes tiles into that framebuffer, and keeps its per-tile background flags in step.  Destroy the
 * mirror before `target`. */
int32_t grb_mirror_create_on_framebuffer(grb_context *ctx, grb_framebuffer *target, int32_t plane, grb_mirror **out);
int32_t grb_mirror_destroy(grb_mirror *m);
int32_t grb_mirror_invalidate(grb_mirror *m);
/* Bring mirror frames [color_frame0, +nframes) / [depth_frame0, +nframes) up to date with device frames
 * [frame0, +nframes) of `fb`.  Either mirror may be NULL.  Runs on the context's copy stream after
 * everything queued on the render stream; a later draw into the same framebuffer waits for it. */
int32_t grb_mirror_update_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                                grb_mirror *color, int32_t color_frame0, grb_mirror *depth, int32_t depth_frame0);
/* The same for the tile rows [row_begin, row_end) only (both multiples of GRB_TILE; 0,0 = the whole frame): the
 * rows a strip draw has rendered. */
int32_t grb_mirror_update_rows_async(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                                     grb_mirror *color, int32_t color_frame0, grb_mirror *depth, int32_t depth_frame0,
                                     int32_t row_begin, int32_t row_end);
/* Blocks until the most recent update of `m` has landed in host memory. */
int32_t grb_mirror_wait(grb_mirror *m);
/* Tiles written into the host plane / tiles a full copy would have moved, since creation (synchronises). */
int32_t grb_mirror_stats(grb_mirror *m, int64_t *tiles_written, int64_t *tiles_full);

/* The literal (*Renderer).Draw (renderer.go:443-483) for a caller whose FrameBuffer is host memory:
 * draw `nframes` frames, update the mirrors, return the stats — one call, one synchronisation.  A
 * one-frame call replays a CUDA graph of the whole sequence (matrix upload, counters, setup, raster,
 * stats read-back) captured the first time this combination of scene, framebuffer, options and mirrors
 * is seen, and its raster kernel stores each tile it writes into the mirrors' host planes as well (same
 * rule as a mirror update), so that tiles cross PCIe while the rest of the frame is still being
 * rasterised.  Mirrors and stats may be NULL. */
int32_t grb_draw_present(grb_context *ctx, grb_framebuffer *fb, int32_t frame0, int32_t nframes,
                         const grb_object *objects, int32_t nobj, const grb_draw_params *params,
                         grb_mirror *color, int32_t color_frame0, grb_mirror *depth, int32_t depth_frame0,
                         grb_frame_stats *stats);

/* One-frame grb_draw_present calls served by replaying a cached CUDA graph, since context creation. */
int64_t grb_graph_replays(const grb_context *ctx);

/* ---- the build-tag seam: matrixMultiplyVec4Batch (asm_amd64.go:8-11,
 *      asm_amd64.s:7-50, asm_purego.go:9-19).  In place on n host Vec4s. */
int32_t grb_matrix_multiply_vec4_batch(grb_context *ctx, const float m[16], float *vecs, int64_t n);
/* Same on device memory already resident in HBM (n Vec4s), async. */
int32_t grb_matrix_multiply_vec4_batch_device(grb_context *ctx, const float m[16], void *device_vecs, int64_t n);

/* ---- stage read-backs for parity tests (Object.TransformedVertices,
 *      mesh.go:76; the emitted `Triangle`s, renderer.go:372-377) ---------- */
int32_t grb_debug_read_transformed(grb_context *ctx, int32_t frame, float *out, int64_t capacity_vec4,
                                   int64_t *out_n);
int32_t grb_debug_read_triangles(grb_context *ctx, int32_t frame, grb_triangle_rec *out, float *out_uvs /* 6 per tri, may be NULL */,
                                 int64_t capacity, int64_t *out_n);
/* Tests of the list-fallback path: force the per-frame overflow pool to `entries` descriptors (0 = the
 * automatic size, twice the faces of the frame).  Results never depend on it; speed does. */
int32_t grb_debug_set_overflow_cap(grb_context *ctx, uint32_t entries);
/* BoxVisibility (clipping.go:131-154) of each object of `frame` in the last draw. */
int32_t grb_debug_read_visibility(grb_context *ctx, int32_t frame, int32_t *out, int32_t capacity);

#ifdef __cplusplus
}
#endif
#endif
