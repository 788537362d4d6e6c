//go:build cuda && headless

// main_headless.go — the headless driver that replaces main.go's raylib window
// (main.go:145-352) when building with `-tags "cuda headless"`: loads an .obj or .json scene,
// renders the demo spin into the offscreen framebuffer and prints FPS / TPF, the same numbers
// the reference's HUD shows (main.go:224-226, 311-314).  It also supplies the two compile-time
// constants renderer.go reads (renderer.go:150,451,476), which live in main.go in the reference.
package main

import (
	"flag"
	"fmt"
	"log"
	"path"
	"time"
)

const (
	parallel = true
	demoMode = true
)

func main() {
	width := flag.Int("w", 1280, "framebuffer width")
	height := flag.Int("h", 720, "framebuffer height")
	frames := flag.Int("frames", 1000, "frames to render")
	flag.Parse()
	if flag.NArg() == 0 {
		log.Fatalf("usage: gorender [options] filename.obj|scene.json")
	}
	filename := flag.Arg(0)

	fb := NewFrameBuffer(*width, *height)
	renderer := NewRenderer(fb)

	var scene *Scene
	switch path.Ext(filename) {
	case ".obj":
		meshes, err := LoadMeshFile(filename, false)
		if err != nil {
			log.Fatalf("failed to load mesh file: %s", err)
		}
		scene = &Scene{}
		for i := range meshes {
			scene.Objects = append(scene.Objects, NewObject(meshes[i]))
		}
	case ".json":
		var err error
		if scene, err = LoadSceneFile(filename); err != nil {
			log.Fatalf("failed to load scene file: %s", err)
		}
	default:
		log.Fatalf("unsupported file format: %s", path.Ext(filename))
	}

	camera := &Camera{Direction: Vec3{0, 0, -1}, Position: Vec3{0, 0, 5}, Up: Vec3{0, 1, 0}}
	start := time.Now()
	for i := 0; i < *frames; i++ {
		renderer.Draw(scene.Objects, camera)
		fb.SwapBuffers()
		for _, obj := range scene.Objects {
			obj.Rotation.Y += 0.01
		}
	}
	sec := time.Since(start).Seconds()
	fps := float64(*frames) / sec
	fmt.Printf("objects=%d vertices=%d triangles=%d frames=%d fps=%.1f tpf=%d ktps=%.0f\n",
		scene.NumObjects(), scene.NumVertices(), scene.NumTriangles(), *frames, fps, renderer.TPF,
		float64(renderer.TPF)*fps/1000)
}
