//go:build paritydump

// parity_dump.go — renders the pinned parity scenes of this repository with the UNMODIFIED reference renderer and
// dumps what it produces (Pixels, ZBuffer, TPF), so that the day a Go toolchain is available the CPU oracle
// (oracle/gorender_oracle.cpp) — and through it every bit-exact claim of the CUDA path — is pinned against the
// real thing instead of against a restatement.  Nothing here touches CUDA.
//
// How to run (on a machine with Go >= 1.22; see INTEGRATION.md section 6):
//
//	python scripts/export_go_scenes.py /tmp/go_scenes                 # in this repository: scenes as raw arrays
//	mkdir /tmp/refbuild && cp /path/to/gorender/*.go /tmp/refbuild/   # the reference's sources ...
//	rm /tmp/refbuild/main.go                                          # ... minus the raylib window
//	cp go/parity/*.go /tmp/refbuild/ && cd /tmp/refbuild
//	go mod init gorender && CGO_ENABLED=0 go build -tags "paritydump purego" -o parity_dump .
//	taskset -c 0-15 ./parity_dump /tmp/go_scenes /tmp/go_dump          # > 16 CPUs panic in NewRenderer (renderer.go:151,160)
//	python scripts/compare_go_dump.py /tmp/go_dump                     # back in this repository
//
// `purego` selects asm_purego.go (same arithmetic as the SSE assembly, asm_amd64.s:33-40); build once more with
// `-tags paritydump` on amd64 to cover the assembly.  The serial branch of Draw (renderer.go:466-474) is used —
// `parallel` is false below — because it is the reference's only deterministic order (SURVEY.md H12); scenes that
// ask for the 16-tile rule get numTiles = 16 and recomputed tile bounds, which the serial branch honours
// (renderer.go:470-472 loops over numTiles).
package main

import (
	"crypto/sha256"
	"encoding/binary"
	"encoding/hex"
	"encoding/json"
	"fmt"
	"image/color"
	"log"
	"math"
	"os"
	"path/filepath"
)

// the two compile-time switches renderer.go reads (renderer.go:150,451,476); they live in main.go in the reference
const (
	parallel = false
	demoMode = true
)

type sceneObject struct {
	Mesh        int        `json:"mesh"`
	Translation [3]float32 `json:"translation"`
	Rotation    [3]float32 `json:"rotation"`
	Scale       [3]float32 `json:"scale"`
}

type sceneDef struct {
	Name     string          `json:"name"`
	Width    int             `json:"width"`
	Height   int             `json:"height"`
	NumTiles uint            `json:"num_tiles"`
	Options  map[string]bool `json:"options"`
	FogStart float32         `json:"fog_start"`
	FogEnd   float32         `json:"fog_end"`
	FogColor [4]uint8        `json:"fog_color"`
	Camera   struct {
		Position  [3]float32 `json:"position"`
		Direction [3]float32 `json:"direction"`
		Up        [3]float32 `json:"up"`
	} `json:"camera"`
	Meshes  []string      `json:"meshes"`
	Objects []sceneObject `json:"objects"`
}

type reader struct {
	b   []byte
	off int
}

func (r *reader) i32() int32 {
	v := int32(binary.LittleEndian.Uint32(r.b[r.off:]))
	r.off += 4
	return v
}
func (r *reader) f32() float32 {
	v := math.Float32frombits(binary.LittleEndian.Uint32(r.b[r.off:]))
	r.off += 4
	return v
}
func (r *reader) u8() uint8 { v := r.b[r.off]; r.off++; return v }

// loadMesh reads what scripts/export_go_scenes.py wrote and hands it to the reference's own NewMesh
// (mesh.go:53-69), which derives the face normals and the bounding box itself.
func loadMesh(filename string) *Mesh {
	b, err := os.ReadFile(filename)
	if err != nil {
		log.Fatal(err)
	}
	r := &reader{b: b}
	nv, nvn, nf, ntex := int(r.i32()), int(r.i32()), int(r.i32()), int(r.i32())
	vertices := make([]Vec4, nv)
	for i := range vertices {
		vertices[i] = Vec4{r.f32(), r.f32(), r.f32(), r.f32()}
	}
	normals := make([]Vec4, nvn)
	for i := range normals {
		normals[i] = Vec4{r.f32(), r.f32(), r.f32(), r.f32()}
	}
	faces := make([]Face, nf)
	for i := range faces {
		for k := 0; k < 3; k++ {
			faces[i].VertexIndices[k] = int(r.i32())
		}
	}
	for i := range faces {
		for k := 0; k < 3; k++ {
			faces[i].NormalIndices[k] = int(r.i32())
		}
	}
	for i := range faces {
		for k := 0; k < 3; k++ {
			faces[i].UVs[k] = UV{r.f32(), r.f32()}
		}
	}
	texIdx := make([]int, nf)
	for i := range texIdx {
		texIdx[i] = int(r.i32())
	}
	textures := make([]*Texture, ntex)
	for t := range textures {
		typ, w, h := TextureType(r.i32()), int(r.i32()), int(r.i32())
		scale := r.f32()
		col := color.RGBA{r.u8(), r.u8(), r.u8(), r.u8()}
		tex := &Texture{typ: typ, width: w, height: h, widthF: float32(w), heightF: float32(h), scale: scale, color: col}
		if typ != TextureTypeSolidColor {
			tex.pixels = make([]color.RGBA, w*h) // premultiplied, as NewImageTexture leaves them (texture.go:57)
			for i := range tex.pixels {
				tex.pixels[i] = color.RGBA{r.u8(), r.u8(), r.u8(), r.u8()}
			}
		}
		textures[t] = tex
	}
	for i := range faces {
		if texIdx[i] >= 0 {
			faces[i].Texture = textures[texIdx[i]]
		}
	}
	if r.off != len(b) {
		log.Fatalf("%s: %d trailing bytes", filename, len(b)-r.off)
	}
	return NewMesh(vertices, normals, faces)
}

func main() {
	if len(os.Args) != 3 {
		log.Fatalf("usage: parity_dump <scene dir> <output dir>")
	}
	in, out := os.Args[1], os.Args[2]
	if err := os.MkdirAll(out, 0o755); err != nil {
		log.Fatal(err)
	}
	raw, err := os.ReadFile(filepath.Join(in, "scenes.json"))
	if err != nil {
		log.Fatal(err)
	}
	var scenes []sceneDef
	if err := json.Unmarshal(raw, &scenes); err != nil {
		log.Fatal(err)
	}
	results := map[string]map[string]interface{}{}
	meshCache := map[string]*Mesh{}
	for _, sc := range scenes {
		fb := NewFrameBuffer(sc.Width, sc.Height)
		r := NewRenderer(fb)
		if sc.NumTiles > 1 { // the 16-tile membership rule and TPF (renderer.go:151), in the serial order
			r.numTiles = sc.NumTiles
			for i := uint(0); i < r.numTiles; i++ {
				start, end := calculateTileBoundaries(i, r.numTiles, fb.Width, fb.Height)
				r.tileBounds[i] = [2]Vec2{start, end}
			}
		}
		for k, v := range sc.Options {
			switch k {
			case "FrustumClipping":
				r.FrustumClipping = v
			case "ShowVertices":
				r.ShowVertices = v
			case "ShowEdges":
				r.ShowEdges = v
			case "ShowFaces":
				r.ShowFaces = v
			case "BackfaceCulling":
				r.BackfaceCulling = v
			case "Lighting":
				r.Lighting = v
			case "FlatShading":
				r.FlatShading = v
			case "ShowTextures":
				r.ShowTextures = v
			case "CrossHair", "Fog": // post passes, applied below
			default:
				log.Fatalf("%s: unknown option %s", sc.Name, k)
			}
		}
		var objects []*Object
		for _, so := range sc.Objects {
			file := sc.Meshes[so.Mesh]
			mesh, ok := meshCache[file]
			if !ok {
				mesh = loadMesh(filepath.Join(in, file))
				meshCache[file] = mesh
			}
			o := NewObject(mesh)
			o.Translation = Vec3{so.Translation[0], so.Translation[1], so.Translation[2]}
			o.Rotation = Vec3{so.Rotation[0], so.Rotation[1], so.Rotation[2]}
			o.Scale = Vec3{so.Scale[0], so.Scale[1], so.Scale[2]}
			objects = append(objects, o)
		}
		cam := &Camera{
			Position:  Vec3{sc.Camera.Position[0], sc.Camera.Position[1], sc.Camera.Position[2]},
			Direction: Vec3{sc.Camera.Direction[0], sc.Camera.Direction[1], sc.Camera.Direction[2]},
			Up:        Vec3{sc.Camera.Up[0], sc.Camera.Up[1], sc.Camera.Up[2]},
		}
		r.Draw(objects, cam)
		// the post passes of Draw, in its order (renderer.go:476-480; compiled out by demoMode, so called here)
		if sc.Options["CrossHair"] {
			fb.CrossHair(color.RGBA{255, 255, 0, 255})
		}
		if sc.Options["Fog"] {
			fb.Fog(sc.FogStart, sc.FogEnd, color.RGBA{sc.FogColor[0], sc.FogColor[1], sc.FogColor[2], sc.FogColor[3]})
		}
		px := make([]byte, 4*len(fb.Pixels))
		for i, c := range fb.Pixels {
			px[4*i], px[4*i+1], px[4*i+2], px[4*i+3] = c.R, c.G, c.B, c.A
		}
		zb := make([]byte, 4*len(fb.ZBuffer))
		for i, z := range fb.ZBuffer {
			binary.LittleEndian.PutUint32(zb[4*i:], math.Float32bits(z))
		}
		if err := os.WriteFile(filepath.Join(out, sc.Name+".pixels"), px, 0o644); err != nil {
			log.Fatal(err)
		}
		if err := os.WriteFile(filepath.Join(out, sc.Name+".zbuffer"), zb, 0o644); err != nil {
			log.Fatal(err)
		}
		hp, hz := sha256.Sum256(px), sha256.Sum256(zb)
		results[sc.Name] = map[string]interface{}{
			"width": sc.Width, "height": sc.Height, "tpf": r.TPF,
			"pixels_sha256": hex.EncodeToString(hp[:]), "zbuffer_sha256": hex.EncodeToString(hz[:]),
		}
		fmt.Printf("%-32s %dx%d tpf=%d\n", sc.Name, sc.Width, sc.Height, r.TPF)
	}
	js, _ := json.MarshalIndent(results, "", " ")
	if err := os.WriteFile(filepath.Join(out, "go_reference_outputs.json"), js, 0o644); err != nil {
		log.Fatal(err)
	}
}
