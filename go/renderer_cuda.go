//go:build cuda

// renderer_cuda.go — the cgo shim that puts gorender's per-frame hot path on a B200.
//
// Drop this file (and gorender_b200.h + libgorender_b200.so) next to the reference's sources,
// add `//go:build !cuda` to the top of a file holding the reference's own
// `func (r *Renderer) Draw` (renderer.go:443-483), and build with
//
//	CGO_ENABLED=1 go build -tags cuda -o gorender
//
// Nothing else changes: NewFrameBuffer, NewRenderer, the option fields, Camera, Object, Mesh,
// LoadObjFile, LoadSceneFile and Texture keep their reference definitions, and matrix.go keeps
// building the per-object matrices on the host (so Go's math.Sin/Cos/Tan are the ones used).
// Draw keeps its signature — no return value, results are side effects on fb.Pixels,
// fb.ZBuffer and r.TPF — and, like the reference, cannot fail: a non-zero status panics.
//
// NOTE: the Go toolchain is not available where this repository is developed; this file is
// written against include/gorender_b200.h and mirrors, call for call, the Python binding
// (gorender_b200/renderer.py) that the GPU parity tests exercise.
package main

/*
#cgo CFLAGS: -I${SRCDIR}
#cgo LDFLAGS: -L${SRCDIR} -lgorender_b200 -Wl,-rpath,${SRCDIR}
#include <stdlib.h>
#include "gorender_b200.h"
*/
import "C"

import (
	"fmt"
	"sync"
	"unsafe"
)

// cudaRenderer is the per-Renderer device state (one grb_context == one GPU).
type cudaRenderer struct {
	ctx      *C.grb_context
	fb       *C.grb_framebuffer
	meshes   map[*Mesh]C.int32_t
	textures map[*Texture]C.int32_t
	objs     []C.grb_object
	// Host mirrors of FrameBuffer.Pixels / Pixels2 / ZBuffer, keyed by the address of the slice's first
	// element: SwapBuffers (rasterizer.go:32-34) only swaps the two slices, so the mirror of whatever
	// fb.Pixels currently is can be looked up without touching FrameBuffer.  The slices are pinned in
	// place once (grb_host_register; Go's collector does not move heap objects), after which a Draw moves
	// only the 32x32 tiles that changed instead of 7.4 MB per 720p frame.
	mirrors map[unsafe.Pointer]*C.grb_mirror
	pins    runtimePinner // the mirrored slices stay pinned (runtime.Pinner) for as long as the library holds their address
	pending *C.grb_mirror // colour mirror of a DrawAsync still in flight
}

var (
	cudaStates   = map[*Renderer]*cudaRenderer{}
	cudaStatesMu sync.Mutex
)

func cudaCheck(ctx *C.grb_context, rc C.int32_t, what string) {
	if rc != C.GRB_OK {
		panic(fmt.Sprintf("gorender_b200: %s: %s", what, C.GoString(C.grb_last_error(ctx))))
	}
}

func (r *Renderer) cuda() *cudaRenderer {
	cudaStatesMu.Lock()
	defer cudaStatesMu.Unlock()
	if s, ok := cudaStates[r]; ok {
		return s
	}
	s := &cudaRenderer{meshes: map[*Mesh]C.int32_t{}, textures: map[*Texture]C.int32_t{}, mirrors: map[unsafe.Pointer]*C.grb_mirror{}}
	cudaCheck(nil, C.grb_context_create(0, &s.ctx), "grb_context_create")
	cudaCheck(s.ctx, C.grb_framebuffer_create(s.ctx, C.int32_t(r.fb.Width), C.int32_t(r.fb.Height), 1, &s.fb),
		"grb_framebuffer_create")
	cudaStates[r] = s
	return s
}

// textureID uploads a Texture once (texture.go:19-26 fields) and returns its device id; nil -> -1.
func (s *cudaRenderer) textureID(t *Texture) C.int32_t {
	if t == nil {
		return -1
	}
	if id, ok := s.textures[t]; ok {
		return id
	}
	var id C.int32_t
	col := [4]C.uint8_t{C.uint8_t(t.color.R), C.uint8_t(t.color.G), C.uint8_t(t.color.B), C.uint8_t(t.color.A)}
	var px *C.uint8_t
	if len(t.pixels) > 0 {
		px = (*C.uint8_t)(unsafe.Pointer(&t.pixels[0])) // color.RGBA is 4 x uint8, R,G,B,A
	}
	cudaCheck(s.ctx, C.grb_texture_upload(s.ctx, C.int32_t(t.typ), C.int32_t(t.width), C.int32_t(t.height),
		C.float(t.scale), &col[0], px, &id), "grb_texture_upload")
	s.textures[t] = id
	return id
}

// meshID flattens a Mesh (mesh.go:12-26: Face holds a Go pointer and 64-bit ints, so it cannot
// cross cgo as is) and uploads it once.
func (s *cudaRenderer) meshID(m *Mesh) C.int32_t {
	if id, ok := s.meshes[m]; ok {
		return id
	}
	nf := len(m.Faces)
	vidx := make([]int32, 3*nf)
	nidx := make([]int32, 3*nf)
	uvs := make([]float32, 6*nf)
	tex := make([]int32, nf)
	for i := range m.Faces {
		f := &m.Faces[i]
		for k := 0; k < 3; k++ {
			vidx[3*i+k] = int32(f.VertexIndices[k])
			nidx[3*i+k] = int32(f.NormalIndices[k])
			uvs[6*i+2*k] = f.UVs[k].U
			uvs[6*i+2*k+1] = f.UVs[k].V
		}
		tex[i] = int32(s.textureID(f.Texture))
	}
	var d C.grb_mesh_desc
	d.nv, d.nvn, d.nf = C.int32_t(len(m.Vertices)), C.int32_t(len(m.VertexNormals)), C.int32_t(nf)
	// The descriptor's pointers refer to Go memory: pin them for the duration of the call.
	var pin runtimePinner
	defer pin.Unpin()
	d.vertices = (*C.float)(pin.ptr(unsafe.Pointer(&m.Vertices[0])))
	if len(m.VertexNormals) > 0 {
		d.vnormals = (*C.float)(pin.ptr(unsafe.Pointer(&m.VertexNormals[0])))
		d.nidx = (*C.int32_t)(pin.ptr(unsafe.Pointer(&nidx[0])))
	}
	if nf > 0 {
		d.fnormals = (*C.float)(pin.ptr(unsafe.Pointer(&m.FaceNormals[0])))
		d.vidx = (*C.int32_t)(pin.ptr(unsafe.Pointer(&vidx[0])))
		d.uvs = (*C.float)(pin.ptr(unsafe.Pointer(&uvs[0])))
		d.tex = (*C.int32_t)(pin.ptr(unsafe.Pointer(&tex[0])))
	}
	for c := 0; c < 8; c++ {
		d.bbox[4*c+0] = C.float(m.BoundingBox[c].X)
		d.bbox[4*c+1] = C.float(m.BoundingBox[c].Y)
		d.bbox[4*c+2] = C.float(m.BoundingBox[c].Z)
		d.bbox[4*c+3] = C.float(m.BoundingBox[c].W)
	}
	var id C.int32_t
	cudaCheck(s.ctx, C.grb_mesh_upload(s.ctx, &d, &id), "grb_mesh_upload")
	s.meshes[m] = id
	return id
}

func matrixToC(dst *[16]C.float, m *Matrix) {
	for i := 0; i < 4; i++ {
		for j := 0; j < 4; j++ {
			dst[4*i+j] = C.float(m[i][j])
		}
	}
}

func (r *Renderer) optionBits() C.uint32_t {
	var o C.uint32_t
	if r.FrustumClipping {
		o |= C.GRB_OPT_FRUSTUM_CLIPPING
	}
	if r.ShowFaces {
		o |= C.GRB_OPT_SHOW_FACES
	}
	if r.BackfaceCulling {
		o |= C.GRB_OPT_BACKFACE_CULLING
	}
	if r.Lighting {
		o |= C.GRB_OPT_LIGHTING
	}
	if r.FlatShading {
		o |= C.GRB_OPT_FLAT_SHADING
	}
	if r.ShowTextures {
		o |= C.GRB_OPT_SHOW_TEXTURES
	}
	if r.ShowEdges {
		o |= C.GRB_OPT_SHOW_EDGES
	}
	if r.ShowVertices {
		o |= C.GRB_OPT_SHOW_VERTICES
	}
	if !demoMode { // renderer.go:476-480
		o |= C.GRB_OPT_CROSSHAIR
	}
	return o
}

// mirror returns the host mirror laid over a FrameBuffer plane (created and pinned on first use).
func (s *cudaRenderer) mirror(r *Renderer, p unsafe.Pointer, bytes int, plane C.int32_t) *C.grb_mirror {
	if m, ok := s.mirrors[p]; ok {
		return m
	}
	// the library keeps this address beyond the call (the GPU writes tiles into the slice on every Draw): cgo allows that
	// only for pinned Go memory, so the slice is pinned for the renderer's lifetime, and page-locked for the device
	s.pins.ptr(p)
	if rc := C.grb_host_register(p, C.uint64_t(bytes)); rc != C.GRB_OK {
		panic("gorender_b200: grb_host_register failed (cannot pin the framebuffer slice)")
	}
	var m *C.grb_mirror
	cudaCheck(s.ctx, C.grb_mirror_create(s.ctx, C.int32_t(r.fb.Width), C.int32_t(r.fb.Height), 1, plane, p, &m), "grb_mirror_create")
	s.mirrors[p] = m
	return m
}

// pack fills s.objs and the draw parameters: the host work of renderer.go:255-265, unchanged.
func (r *Renderer) pack(s *cudaRenderer, objects []*Object, camera *Camera, p *C.grb_draw_params) (*C.grb_object, C.int32_t) {
	viewMatrix := NewViewMatrix(camera.Position, camera.Direction, camera.Up)
	perspectiveMatrix := NewPerspectiveMatrix(r.fovY, r.aspectX, r.zNear, r.zFar)
	if cap(s.objs) < len(objects) {
		s.objs = make([]C.grb_object, len(objects))
	}
	objs := s.objs[:len(objects)]
	for i, object := range objects {
		worldMatrix := NewWorldMatrix(object.Scale, object.Rotation, object.Translation)
		mvpMatrix := NewIdentityMatrix()
		mvpMatrix = mvpMatrix.Multiply(perspectiveMatrix)
		mvpMatrix = mvpMatrix.Multiply(viewMatrix)
		mvpMatrix = mvpMatrix.Multiply(worldMatrix)
		objs[i].mesh = s.meshID(object.Mesh)
		matrixToC(&objs[i].world, &worldMatrix)
		matrixToC(&objs[i].mvp, &mvpMatrix)
	}
	screenMatrix := NewScreenMatrix(r.fb.Width, r.fb.Height)
	matrixToC(&p.screen, &screenMatrix)
	light := Vec3{X: -1, Y: 1, Z: 1}.Normalize()
	p.light[0], p.light[1], p.light[2] = C.float(light.X), C.float(light.Y), C.float(light.Z)
	p.options = r.optionBits()
	p.z_near, p.z_far = C.float(r.zNear), C.float(r.zFar)
	// 16 (parallel) or 1.  The unmodified NewRenderer sets max(NumCPU, 16) and indexes [16]-arrays with it: on
	// hosts with more than 16 CPUs it panics before this shim is ever reached (renderer.go:151,160; SURVEY H1) —
	// run under `taskset -c 0-15` there.
	p.ref_tiles = C.int32_t(min(r.numTiles, 16))
	p.row_begin, p.row_end = 0, 0
	// GRB_OPT_FOG is never set (the Fog call is commented out at renderer.go:479); its arguments are the ones written there
	p.fog_start, p.fog_end = 0.100, 0.033
	p.fog_color = [4]C.uint8_t{100, 100, 100, 255}
	if len(objs) == 0 {
		return nil, 0
	}
	return &objs[0], C.int32_t(len(objs)) // grb_object holds no Go pointers: legal to pass
}

// Draw replaces renderer.go:443-483.  Host work per object is exactly the reference's
// renderer.go:255-265 (matrix constructors); everything else happens on the GPU.  One C call, one
// synchronisation: draw, bring fb.Pixels and fb.ZBuffer up to date (only the tiles that changed cross PCIe),
// return TPF; from the second frame of a scene on it is the replay of a CUDA graph.
func (r *Renderer) Draw(objects []*Object, camera *Camera) {
	s := r.cuda()
	r.WaitFrame()
	var p C.grb_draw_params
	objPtr, n := r.pack(s, objects, camera, &p)
	color := s.mirror(r, unsafe.Pointer(&r.fb.Pixels[0]), 4*len(r.fb.Pixels), C.GRB_PLANE_COLOR)
	depth := s.mirror(r, unsafe.Pointer(&r.fb.ZBuffer[0]), 4*len(r.fb.ZBuffer), C.GRB_PLANE_DEPTH)
	var stats C.grb_frame_stats
	cudaCheck(s.ctx, C.grb_draw_present(s.ctx, s.fb, 0, 1, objPtr, n, &p, color, 0, depth, 0, &stats), "grb_draw_present")
	r.TPF = int(stats.tpf)
}

// DrawAsync is the streaming form for the reference's render / present loop (main.go:198-227): it queues the
// draw and the update of fb.Pixels and returns; the frame has landed after WaitFrame (call it after
// fb.SwapBuffers(), before reading fb.Pixels2).  fb.ZBuffer and r.TPF are not updated.
func (r *Renderer) DrawAsync(objects []*Object, camera *Camera) {
	s := r.cuda()
	r.WaitFrame()
	var p C.grb_draw_params
	objPtr, n := r.pack(s, objects, camera, &p)
	color := s.mirror(r, unsafe.Pointer(&r.fb.Pixels[0]), 4*len(r.fb.Pixels), C.GRB_PLANE_COLOR)
	cudaCheck(s.ctx, C.grb_draw_async(s.ctx, s.fb, 0, 1, objPtr, n, &p), "grb_draw_async")
	cudaCheck(s.ctx, C.grb_mirror_update_async(s.ctx, s.fb, 0, 1, color, 0, nil, 0), "grb_mirror_update_async")
	s.pending = color
}

// WaitFrame blocks until the frame queued by the last DrawAsync is in host memory.
func (r *Renderer) WaitFrame() {
	s := r.cuda()
	if s.pending != nil {
		cudaCheck(s.ctx, C.grb_mirror_wait(s.pending), "grb_mirror_wait")
		s.pending = nil
	}
}

// matrixMultiplyVec4BatchCUDA is the third implementation behind the reference's build-tag seam
// (asm_amd64.go:8 / asm_purego.go:9), for callers that want the batch transform alone.
func matrixMultiplyVec4BatchCUDA(r *Renderer, m *Matrix, vecs []Vec4) {
	if len(vecs) == 0 {
		return
	}
	s := r.cuda()
	var mm [16]C.float
	matrixToC(&mm, m)
	cudaCheck(s.ctx, C.grb_matrix_multiply_vec4_batch(s.ctx, &mm[0], (*C.float)(unsafe.Pointer(&vecs[0])),
		C.int64_t(len(vecs))), "grb_matrix_multiply_vec4_batch")
}
