// gorender_headless — the headless driver that stands in for main.go's raylib window
// (main.go:145-352): loads an .obj, renders the demo spin (Rotation.Y += 0.01 per frame,
// main.go:229-233) through the C ABI into the offscreen framebuffer and prints the HUD numbers
// (FPS, TPF, k tps: main.go:224-226).  Options:
//   -w W -h H        framebuffer size (default 1280x720)
//   -frames N        frames to render (default 100)
//   -start K         spin frames to skip before the first rendered one
//   -ppm FILE        write the last frame as a binary PPM
//   -raw FILE        write the last frame's RGBA8 pixels followed by the f32 z-buffer
//   -async           streaming loop: DrawAsync + SwapBuffers, the read-back of frame i overlaps frame i+1
//   -matrices        print world / mvp of the last frame as hex words (host-math cross-check, no GPU needed)
//   -texdump         print size, type and an FNV-1a hash of the premultiplied texels of a PNG (no GPU needed)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gorender_host.hpp"

using namespace gorender;

static void printMatrix(const char *name, const Matrix &m) {
    std::printf("%s", name);
    for (int i = 0; i < 16; i++) {
        uint32_t u;
        std::memcpy(&u, m.data() + i, 4);
        std::printf(" %08x", u);
    }
    std::printf("\n");
}

int main(int argc, char **argv) {
    int width = 1280, height = 720, frames = 100, start = 0;
    bool matricesOnly = false, texDump = false, async = false, native = false;
    // the option hot-keys of main.go:255-272 as flags
    bool edges = false, vertices = false, noFaces = false, crossHair = false, noCull = false, noLight = false, flat = false;
    std::string ppm, raw, file;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "-w") width = std::atoi(next());
        else if (a == "-h") height = std::atoi(next());
        else if (a == "-frames") frames = std::atoi(next());
        else if (a == "-start") start = std::atoi(next());
        else if (a == "-ppm") ppm = next();
        else if (a == "-raw") raw = next();
        else if (a == "-matrices") matricesOnly = true;
        else if (a == "-texdump") texDump = true;
        else if (a == "-async") async = true;
        else if (a == "-native") native = true;   // native OBJ parser + NewMesh on the device
        else if (a == "-edges") edges = true;
        else if (a == "-vertices") vertices = true;
        else if (a == "-nofaces") noFaces = true;
        else if (a == "-crosshair") crossHair = true;
        else if (a == "-nocull") noCull = true;
        else if (a == "-nolight") noLight = true;
        else if (a == "-flat") flat = true;
        else file = a;
    }
    if (file.empty()) {
        std::fprintf(stderr, "usage: %s [options] filename.obj\n", argv[0]);  // main.go:58-60
        return 2;
    }
    try {
        if (texDump) {
            TexturePtr t = LoadTextureFile(file);
            uint64_t h = 1469598103934665603ull;
            for (uint8_t b : t->pixels) { h ^= b; h *= 1099511628211ull; }
            std::printf("%d %d %d %016llx\n", t->width, t->height, (int)t->typ, (unsigned long long)h);
            return 0;
        }
        Scene scene;
        if (!native || matricesOnly)
            for (auto &m : LoadMeshFile(file, false)) scene.Objects.push_back(NewObject(m));  // main.go:155-165
        Camera camera{{0, 0, 5}, {0, 0, -1}, {0, 1, 0}};                                   // main.go:192-196
        for (int k = 0; k < start; k++)
            for (auto &o : scene.Objects) o->Rotation.Y += 0.01f;

        if (matricesOnly) {
            struct StubFB { int Width, Height; };
            for (int f = 0; f + 1 < frames; f++)
                for (auto &o : scene.Objects) o->Rotation.Y += 0.01f;
            // same arithmetic as Renderer::objectMatrices, without a device
            float aspectX = (float)width / (float)height, fovY = (float)(45 * (M_PI / 180));
            for (auto &o : scene.Objects) {
                Matrix world = NewWorldMatrix(o->Scale, o->Rotation, o->Translation);
                Matrix view = NewViewMatrix(camera.Position, camera.Direction, camera.Up);
                Matrix persp = NewPerspectiveMatrix(fovY, aspectX, 0.0f, 50.0f);
                Matrix mvp = Multiply(Multiply(Multiply(NewIdentityMatrix(), persp), view), world);
                printMatrix("world", world);
                printMatrix("mvp", mvp);
            }
            std::printf("vertices=%d triangles=%d\n", scene.NumVertices(), scene.NumTriangles());
            return 0;
        }

        Device dev(0);
        if (native) {   // the same scene through grb_obj_parse + grb_mesh_new
            for (auto &m : dev.LoadObjFileNative(file, false)) scene.Objects.push_back(NewObject(m));
            for (int k = 0; k < start; k++)
                for (auto &o : scene.Objects) o->Rotation.Y += 0.01f;
        }
        FrameBuffer fb(dev, width, height);
        Renderer renderer(fb);
        renderer.ShowEdges = edges;
        renderer.ShowVertices = vertices;
        renderer.ShowFaces = !noFaces;
        renderer.CrossHair = crossHair;
        renderer.BackfaceCulling = !noCull;
        renderer.Lighting = !noLight;
        renderer.FlatShading = flat;
        const auto t0 = std::chrono::steady_clock::now();
        if (async) {
            // main.go:198-227 with the present step replaced by "the frame is in Pixels2"
            for (int f = 0; f < frames; f++) {
                renderer.DrawAsync(scene.Objects, camera);   // frame f -> Pixels (kernels, then PCIe)
                if (f > 0) fb.WaitFront();                   // frame f-1 has landed in Pixels2: present it here
                fb.SwapBuffers();                            // frame f is now heading for Pixels2
                if (f + 1 < frames)
                    for (auto &o : scene.Objects) o->Rotation.Y += 0.01f;
            }
            fb.WaitFront();
            fb.SwapBuffers();                                // leave the last frame in Pixels, like Draw does
            renderer.Draw(scene.Objects, camera);            // one synchronous draw for TPF / depth of the last frame
        } else {
            for (int f = 0; f < frames; f++) {
                renderer.Draw(scene.Objects, camera);
                if (f + 1 < frames) {
                    fb.SwapBuffers();
                    for (auto &o : scene.Objects) o->Rotation.Y += 0.01f;
                }
            }
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double fps = frames / sec;
        std::printf("objects=%d vertices=%d triangles=%d frames=%d fps=%.1f tpf=%d ktps=%.0f\n", scene.NumObjects(),
                    scene.NumVertices(), scene.NumTriangles(), frames, fps, renderer.TPF, renderer.TPF * fps / 1000);
        if (!ppm.empty()) {
            FILE *out = std::fopen(ppm.c_str(), "wb");
            if (!out) throw std::runtime_error("cannot write " + ppm);
            std::fprintf(out, "P6\n%d %d\n255\n", width, height);
            for (size_t i = 0; i < (size_t)width * height; i++) std::fwrite(&fb.Pixels[4 * i], 1, 3, out);
            std::fclose(out);
        }
        if (!raw.empty()) {
            FILE *out = std::fopen(raw.c_str(), "wb");
            if (!out) throw std::runtime_error("cannot write " + raw);
            std::fwrite(fb.Pixels.data(), 1, fb.Pixels.size(), out);
            std::fwrite(fb.ZBuffer.data(), 4, fb.ZBuffer.size(), out);
            std::fclose(out);
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "gorender_headless: %s\n", e.what());  // log.Fatalf
        return 1;
    }
    return 0;
}
