/*
 * gorender_oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * The oracle is a C++ restatement of the reference renderer's per-frame hot
 * path (reference: maxpoletaev/gorender, renderer.go / clipping.go /
 * rasterizer.go / texture.go / matrix.go / vector.go / mesh.go).  Every
 * function in gorender_oracle.cpp cites the reference file:line it follows.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, no known-answer
 * tests and no fixtures for this path (asm_test.go holds benchmark inputs
 * only), and it is written in Go, for which no toolchain exists in this
 * image, so it cannot be run here either.  The oracle is exact by
 * construction (IEEE-754 binary32 + - * / sqrt and integer arithmetic in the
 * reference's order, built with -ffp-contract=off), but nothing produced by
 * the Go binary anchors it.  What does: a second restatement of the same Go functions in plain
 * Python (oracle/py_project.py, oracle/py_raster.py) that this one must equal bit for bit on small
 * scenes (tests/test_py_raster.py), the reference's C prototype of the batch transform compiled
 * unmodified (oracle/_ref), the counts of SURVEY.md §8c and the committed SHA-256 pins.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (gorender_b200/) never does.
 */
#ifndef GORENDER_ORACLE_H
#define GORENDER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Renderer option bits == reference Renderer bool fields (renderer.go:90-97). */
enum {
    ORC_OPT_FRUSTUM_CLIPPING = 1u << 0,
    ORC_OPT_SHOW_FACES       = 1u << 1,
    ORC_OPT_BACKFACE_CULLING = 1u << 2,
    ORC_OPT_LIGHTING         = 1u << 3,
    ORC_OPT_FLAT_SHADING     = 1u << 4,
    ORC_OPT_SHOW_TEXTURES    = 1u << 5,
    ORC_OPT_SHOW_EDGES       = 1u << 6,   /* renderer.go:191-210 */
    ORC_OPT_SHOW_VERTICES    = 1u << 7,   /* renderer.go:212-216 */
    ORC_OPT_CROSSHAIR        = 1u << 8,   /* renderer.go:476-478 (`!demoMode`, a constant in main.go:22) */
    ORC_OPT_FOG              = 1u << 9,   /* renderer.go:479 (commented out in the reference) */
    /* NOT in the reference: README.md:46 lists "Affine texture mapping" but no code path exists.  The mode
     * restated here is this repository's own definition (include/gorender_b200.h, GRB_OPT_AFFINE_TEXTURES): rasterizer.go:158-159
     * with the 1/w factors and the division by zRec removed.  There is nothing of the reference to pin it against. */
    ORC_OPT_AFFINE_TEXTURES  = 1u << 10,
    ORC_OPT_DEFAULT = ORC_OPT_FRUSTUM_CLIPPING | ORC_OPT_SHOW_FACES |
                      ORC_OPT_BACKFACE_CULLING | ORC_OPT_LIGHTING |
                      ORC_OPT_SHOW_TEXTURES /* renderer.go:130-137 */
};

/* texture.go:11-15 */
enum { ORC_TEX_SOLID = 0, ORC_TEX_IMAGE = 1, ORC_TEX_IMAGE_FAST = 2 };

/* clipping.go:23-27 */
enum { ORC_BOX_OUTSIDE = 0, ORC_BOX_INTERSECT = 1, ORC_BOX_INSIDE = 2 };

typedef struct orc_texture {
    int32_t type;
    int32_t width, height;
    float scale;
    uint8_t color[4];       /* solid colour (RGBA) */
    const uint8_t *pixels;  /* width*height RGBA8, premultiplied, row-major */
} orc_texture;

/* Flattened mesh.go Mesh/Face (mesh.go:12-26). */
typedef struct orc_mesh {
    int32_t nv, nvn, nf;
    const float *vertices;   /* nv  * 4 (x,y,z,1)  */
    const float *vnormals;   /* nvn * 4 (x,y,z,1)  */
    const float *fnormals;   /* nf  * 4 (x,y,z,1)  */
    const int32_t *vidx;     /* nf * 3 */
    const int32_t *nidx;     /* nf * 3 */
    const float *uvs;        /* nf * 6 (u0,v0,u1,v1,u2,v2) */
    const int32_t *tex;      /* nf, index into textures, -1 = nil */
    float bbox[32];          /* 8 corners * (x,y,z,1), mesh.go:41-50 order */
} orc_mesh;

typedef struct orc_object {
    int32_t mesh;
    float world[16];         /* row-major, NewWorldMatrix  (matrix.go:82)  */
    float mvp[16];           /* row-major, ((I*P)*V)*W     (renderer.go:259-262) */
} orc_object;

/* One emitted triangle, reference `Triangle` (renderer.go:29-34) + provenance. */
typedef struct orc_triangle {
    float points[12];        /* 3 * (x_screen, y_screen, z_screen, w_clip) */
    float uvs[6];
    float intensity[3];
    int32_t tex;
    int32_t object, face, fan;
} orc_triangle;

typedef struct orc_renderer orc_renderer;

/* num_tiles: 16 (reference parallel=true, renderer.go:151 on <=16 CPUs) or 1
 * (parallel=false).  threads: 0 = the serial branch (renderer.go:466-474),
 * which is the canonical order; >0 = the reference's goroutine structure
 * (one projection task per object, one raster task per tile) on a pool of
 * `threads` OS threads — used only for the timed CPU baseline. */
orc_renderer *orc_renderer_create(int32_t width, int32_t height, int32_t num_tiles, int32_t threads);
void orc_renderer_destroy(orc_renderer *r);

/* Keep a copy of every emitted triangle (submission order; serial mode only). */
void orc_renderer_record_triangles(orc_renderer *r, int32_t enable);

/* Arguments of FrameBuffer.Fog (rasterizer.go:193-207) used when ORC_OPT_FOG is set; the
 * defaults are the ones of the commented-out call at renderer.go:479. */
void orc_renderer_set_fog(orc_renderer *r, float fog_start, float fog_end, const uint8_t color[4]);

/* Renderer.Draw (renderer.go:443-483).  Returns 0, or -1 on bad arguments. */
int32_t orc_renderer_draw(orc_renderer *r,
                          const orc_mesh *meshes, int32_t nmesh,
                          const orc_texture *textures, int32_t ntex,
                          const orc_object *objects, int32_t nobj,
                          const float screen[16], const float light[3],
                          uint32_t options);

/* Timed CPU baseline: nframes consecutive Draw calls (frame f draws
 * objects[f*nobj ...]); returns the seconds spent, measured in C. */
double orc_renderer_draw_sequence(orc_renderer *r,
                                  const orc_mesh *meshes, int32_t nmesh,
                                  const orc_texture *textures, int32_t ntex,
                                  const orc_object *objects, int32_t nobj, int32_t nframes,
                                  const float screen[16], const float light[3],
                                  uint32_t options);

const uint8_t *orc_renderer_pixels(const orc_renderer *r);   /* W*H*4 */
const float *orc_renderer_zbuffer(const orc_renderer *r);    /* W*H   */
int64_t orc_renderer_tpf(const orc_renderer *r);
int64_t orc_renderer_pixel_writes(const orc_renderer *r);    /* z-test passes, serial mode */
int64_t orc_renderer_num_triangles(const orc_renderer *r);
const orc_triangle *orc_renderer_triangles(const orc_renderer *r);
/* per-object BoxVisibility of the last draw; n = min(nobj, cap) entries copied */
int32_t orc_renderer_visibility(const orc_renderer *r, int32_t *out, int32_t cap);

/* --- stage-level entry points (each one a reference function) --- */

/* matrixMultiplyVec4Batch: scalar twin (asm_purego.go:9-19) and SSE twin
 * (asm_amd64.go:8-11 + asm_amd64.s:7-50).  In place, n vec4s. */
void orc_matvec4_batch_scalar(const float m[16], float *vecs, int64_t n);
void orc_matvec4_batch_sse(const float m[16], float *vecs, int64_t n);

/* Frustum.BoxVisibility (clipping.go:131-154) on 8 clip-space corners. */
int32_t orc_box_visibility(const float bbox_clip[32], float z_near, float z_far);

/* Frustum.ClipTriangle (clipping.go:167-236).  Returns the number of output
 * triangles (0..7); out arrays sized 9*12 / 9*6 / 9*3 floats. */
int32_t orc_clip_triangle(const float pts[12], const float uvs[6], const float intens[3],
                          float z_near, float z_far,
                          float *pts_out, float *uvs_out, float *intens_out);

/* Host-side matrix constructors (matrix.go) — used to cross-check the product's
 * host layer, never by it. */
void orc_world_matrix(const float scale[3], const float rot[3], const float trans[3], float out[16]);
void orc_view_matrix(const float eye[3], const float dir[3], const float up[3], float out[16]);
void orc_perspective_matrix(float fov, float aspect, float z_near, float z_far, float out[16]);
void orc_screen_matrix(int32_t width, int32_t height, float out[16]);
void orc_matrix_multiply(const float a[16], const float b[16], float out[16]);
/* mvp = ((I*P)*V)*W  (renderer.go:259-262) */
void orc_mvp_matrix(const float persp[16], const float view[16], const float world[16], float out[16]);
/* normalize(-1,1,1) (renderer.go:265) */
void orc_light_direction(float out[3]);

/* NewMesh face normals + boundingBox (mesh.go:28-69). */
void orc_face_normals(const float *vertices, const int32_t *vidx, int32_t nf, float *out);
void orc_bounding_box(const float *vertices, int32_t nv, float out[32]);

/* Texture.Sample (texture.go:69-89) -> RGBA. */
void orc_texture_sample(const orc_texture *t, float u, float v, uint8_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
