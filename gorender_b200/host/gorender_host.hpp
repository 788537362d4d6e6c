// gorender_host.hpp — C++ host side above the C ABI (include/gorender_b200.h).
//
// The reference is compiled Go and its toolchain is absent here, so this is the
// compiled-language mirror of what stays on the host in the reference: the
// float32 vector / matrix library (vector.go, matrix.go, math32.go), the mesh
// containers and their load-time precompute (mesh.go), the OBJ / MTL loader
// (obj.go), PNG textures (texture.go:28-63, zlib inflate instead of Go's
// image/png) and the Renderer / FrameBuffer / Camera API (renderer.go,
// rasterizer.go) whose Draw runs on the GPU through the ABI.  Same names,
// argument meaning and error behaviour (errors that are Go `error` returns are
// C++ exceptions; what panics in Go throws).
//
// Build with -ffp-contract=off: the matrices must be the ones Go's amd64 build
// produces (no fused multiply-add).
#pragma once

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gorender_b200.h"

namespace gorender {

// ------------------------------------------------------------------ vector.go / math32.go

struct Vec3 {
    float X = 0, Y = 0, Z = 0;
};
struct Vec4 {
    float X = 0, Y = 0, Z = 0, W = 0;
};
struct UV {
    float U = 0, V = 0;
};

inline float sqrt32(float x) { return (float)std::sqrt((double)x); }  // math32.go:11-13
inline float sin32(float x) { return (float)std::sin((double)x); }    // math32.go:15-17
inline float cos32(float x) { return (float)std::cos((double)x); }    // math32.go:19-21
inline float tan32(float x) { return (float)std::tan((double)x); }    // math32.go:23-25

inline Vec3 Sub(Vec3 a, Vec3 b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z}; }  // vector.go:51-53
inline Vec3 CrossProduct(Vec3 a, Vec3 b) {                                      // vector.go:67-72
    float x = a.Y * b.Z - a.Z * b.Y;
    float y = a.Z * b.X - a.X * b.Z;
    float z = a.X * b.Y - a.Y * b.X;
    return {x, y, z};
}
inline float DotProduct(Vec3 a, Vec3 b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }  // vector.go:74-76
inline float Length(Vec3 a) { return sqrt32(a.X * a.X + a.Y * a.Y + a.Z * a.Z); }      // vector.go:63-65
inline Vec3 Normalize(Vec3 a) {                                                        // vector.go:78-80
    float n = Length(a);
    return {a.X / n, a.Y / n, a.Z / n};
}
inline Vec3 ToRadians(Vec3 a) {  // vector.go:82-85
    float f = (float)M_PI / 180;
    return {a.X * f, a.Y * f, a.Z * f};
}

// ------------------------------------------------------------------ matrix.go

struct Matrix {
    float m[4][4];
    float *data() { return &m[0][0]; }
    const float *data() const { return &m[0][0]; }
};

inline Matrix NewIdentityMatrix() {  // matrix.go:5-12
    Matrix r{};
    for (int i = 0; i < 4; i++) r.m[i][i] = 1;
    return r;
}
inline Matrix Multiply(const Matrix &a, const Matrix &b) {  // matrix.go:155-165
    Matrix r{};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
inline Matrix NewScaleMatrix(float x, float y, float z) {  // matrix.go:14-21
    Matrix r = NewIdentityMatrix();
    r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z;
    return r;
}
inline Matrix NewTranslationMatrix(float x, float y, float z) {  // matrix.go:23-30
    Matrix r = NewIdentityMatrix();
    r.m[0][3] = x; r.m[1][3] = y; r.m[2][3] = z;
    return r;
}
inline Matrix NewRotationXMatrix(float a) {  // matrix.go:32-44
    if (a == 0) return NewIdentityMatrix();
    float s = sin32(a), c = cos32(a);
    Matrix r = NewIdentityMatrix();
    r.m[1][1] = c; r.m[1][2] = -s; r.m[2][1] = s; r.m[2][2] = c;
    return r;
}
inline Matrix NewRotationYMatrix(float a) {  // matrix.go:46-58
    if (a == 0) return NewIdentityMatrix();
    float s = sin32(a), c = cos32(a);
    Matrix r = NewIdentityMatrix();
    r.m[0][0] = c; r.m[0][2] = s; r.m[2][0] = -s; r.m[2][2] = c;
    return r;
}
inline Matrix NewRotationZMatrix(float a) {  // matrix.go:60-72
    if (a == 0) return NewIdentityMatrix();
    float s = sin32(a), c = cos32(a);
    Matrix r = NewIdentityMatrix();
    r.m[0][0] = c; r.m[0][1] = -s; r.m[1][0] = s; r.m[1][1] = c;
    return r;
}
inline Matrix NewRotationMatrix(float x, float y, float z) {  // matrix.go:74-80
    Matrix m = NewIdentityMatrix();
    m = Multiply(m, NewRotationXMatrix(x));
    m = Multiply(m, NewRotationYMatrix(y));
    m = Multiply(m, NewRotationZMatrix(z));
    return m;
}
inline Matrix NewWorldMatrix(Vec3 scale, Vec3 rotation, Vec3 translation) {  // matrix.go:82-88
    Matrix m = NewIdentityMatrix();
    m = Multiply(NewScaleMatrix(scale.X, scale.Y, scale.Z), m);
    m = Multiply(NewRotationMatrix(rotation.X, rotation.Y, rotation.Z), m);
    m = Multiply(NewTranslationMatrix(translation.X, translation.Y, translation.Z), m);
    return m;
}
inline Matrix NewPerspectiveMatrix(float fov, float aspect, float zNear, float zFar) {  // matrix.go:92-106
    float tanHalfFov = tan32(fov / 2.0f);
    float m00 = 1 / (aspect * tanHalfFov);
    float m11 = 1 / tanHalfFov;
    float m22 = (zFar + zNear) / (zNear - zFar);
    float m23 = (2 * zFar * zNear) / (zNear - zFar);
    Matrix r{};
    r.m[0][0] = m00; r.m[1][1] = m11; r.m[2][2] = -m22; r.m[2][3] = -m23; r.m[3][2] = -1;
    return r;
}
inline Matrix NewScreenMatrix(int width, int height) {  // matrix.go:108-118
    float hw = (float)width / 2, hh = (float)height / 2;
    Matrix r{};
    r.m[0][0] = hw; r.m[0][3] = hw; r.m[1][1] = hh; r.m[1][3] = hh;
    r.m[2][2] = 0.5f; r.m[2][3] = 0.5f; r.m[3][3] = 1;
    return r;
}
inline Matrix NewViewMatrix(Vec3 eye, Vec3 direction, Vec3 up) {  // matrix.go:133-144
    Vec3 z = Normalize(direction);
    Vec3 x = Normalize(CrossProduct(up, z));
    Vec3 y = Normalize(CrossProduct(z, x));
    Matrix r{};
    r.m[0][0] = x.X; r.m[0][1] = x.Y; r.m[0][2] = x.Z; r.m[0][3] = -DotProduct(x, eye);
    r.m[1][0] = y.X; r.m[1][1] = y.Y; r.m[1][2] = y.Z; r.m[1][3] = -DotProduct(y, eye);
    r.m[2][0] = z.X; r.m[2][1] = z.Y; r.m[2][2] = z.Z; r.m[2][3] = -DotProduct(z, eye);
    r.m[3][3] = 1;
    return r;
}

// ------------------------------------------------------------------ texture.go

enum TextureType { TextureTypeSolidColor = 0, TextureTypeImage = 1, TextureTypeImageFast = 2 };

struct Texture {  // texture.go:19-26
    int width = 0, height = 0;
    float scale = 1.0f;
    uint8_t color[4] = {0, 0, 0, 0};
    std::vector<uint8_t> pixels;  // RGBA8, premultiplied
    TextureType typ = TextureTypeSolidColor;
    void SetScale(float s) { scale = s; }  // texture.go:65-67
};
using TexturePtr = std::shared_ptr<Texture>;

inline bool isPowerOfTwo(int n) { return (n & (n - 1)) == 0; }  // utils.go:5-7

inline TexturePtr NewColorTexture(uint8_t r, uint8_t g, uint8_t b, uint8_t a) {  // texture.go:28-33
    auto t = std::make_shared<Texture>();
    t->color[0] = r; t->color[1] = g; t->color[2] = b; t->color[3] = a;
    return t;
}

// NewImageTexture (texture.go:35-63) from non-premultiplied RGBA8: color.RGBAModel.Convert of
// an NRGBA pixel is ((c * 0x101) * a / 0xff) >> 8 with alpha kept.
inline TexturePtr NewImageTexture(int width, int height, const std::vector<uint8_t> &nrgba) {
    auto t = std::make_shared<Texture>();
    t->width = width; t->height = height; t->scale = 1.0f;
    t->typ = (isPowerOfTwo(width) && isPowerOfTwo(height)) ? TextureTypeImageFast : TextureTypeImage;
    t->pixels.resize((size_t)width * height * 4);
    for (size_t i = 0; i < (size_t)width * height; i++) {
        uint32_t a = nrgba[4 * i + 3];
        for (int c = 0; c < 3; c++) t->pixels[4 * i + c] = (uint8_t)((((uint32_t)nrgba[4 * i + c] * 0x101u) * a / 0xffu) >> 8);
        t->pixels[4 * i + 3] = (uint8_t)a;
    }
    return t;
}

// Minimal PNG reader: greyscale / palette at 1-8 bits, RGB / grey+alpha / RGBA at 8 bits, non-interlaced.
inline void DecodePNG(const std::string &filename, int &width, int &height, std::vector<uint8_t> &rgba) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) throw std::runtime_error("open " + filename + ": no such file or directory");
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8)) throw std::runtime_error("image: unknown format");
    auto be32 = [&](size_t o) { return ((uint32_t)buf[o] << 24) | (buf[o + 1] << 16) | (buf[o + 2] << 8) | buf[o + 3]; };
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    for (size_t o = 8; o + 12 <= buf.size();) {
        uint32_t len = be32(o);
        std::string type(buf.begin() + o + 4, buf.begin() + o + 8);
        const uint8_t *d = buf.data() + o + 8;
        if (o + 12 + len > buf.size()) throw std::runtime_error("png: truncated chunk");
        if (type == "IHDR") {
            width = (int)be32(o + 8); height = (int)be32(o + 12);
            depth = d[8]; ctype = d[9]; interlace = d[12];
        } else if (type == "PLTE") plte.assign(d, d + len);
        else if (type == "tRNS") trns.assign(d, d + len);
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        o += 12 + len;
    }
    const bool subByte = (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4);
    if ((depth != 8 && !subByte) || interlace != 0)
        throw std::runtime_error("png: only non-interlaced images of up to 8 bits per sample are supported");
    int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw std::runtime_error("png: unsupported colour type");
    size_t stride = subByte ? ((size_t)width * depth + 7) / 8 : (size_t)width * ch;
    std::vector<uint8_t> raw((stride + 1) * height);
    uLongf rawLen = raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK || rawLen != raw.size())
        throw std::runtime_error("png: bad IDAT stream");
    std::vector<uint8_t> img(stride * height);
    for (int y = 0; y < height; y++) {  // undo the scanline filters
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t *src = &raw[(stride + 1) * y + 1];
        uint8_t *dst = &img[stride * y];
        const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            int a = x >= (size_t)ch ? dst[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int pred = 0;
            switch (ft) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) / 2; break;
            case 4: { int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                      pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: throw std::runtime_error("png: bad filter");
            }
            dst[x] = (uint8_t)(src[x] + pred);
        }
    }
    if (subByte) {  // unpack to one sample per byte; grey samples scale to 0..255
        std::vector<uint8_t> wide((size_t)width * height);
        const int maxv = (1 << depth) - 1;
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {
                const int bit = x * depth;
                int v = (img[stride * y + bit / 8] >> (8 - depth - bit % 8)) & maxv;
                if (ctype == 0) v = v * 255 / maxv;
                wide[(size_t)y * width + x] = (uint8_t)v;
            }
        img.swap(wide);
    }
    rgba.resize((size_t)width * height * 4);
    for (size_t i = 0; i < (size_t)width * height; i++) {
        uint8_t r, g, b, a = 255;
        const uint8_t *p = &img[i * ch];
        switch (ctype) {
        case 0: r = g = b = p[0]; break;
        case 2: r = p[0]; g = p[1]; b = p[2]; break;
        case 3: if ((size_t)p[0] * 3 + 2 >= plte.size()) throw std::runtime_error("png: palette index out of range");
                r = plte[p[0] * 3]; g = plte[p[0] * 3 + 1]; b = plte[p[0] * 3 + 2];
                if (p[0] < trns.size()) a = trns[p[0]];
                break;
        case 4: r = g = b = p[0]; a = p[1]; break;
        default: r = p[0]; g = p[1]; b = p[2]; a = p[3]; break;
        }
        rgba[4 * i] = r; rgba[4 * i + 1] = g; rgba[4 * i + 2] = b; rgba[4 * i + 3] = a;
    }
}

inline TexturePtr LoadTextureFile(const std::string &filename) {  // texture.go:91-103 (PNG only)
    int w = 0, h = 0;
    std::vector<uint8_t> rgba;
    DecodePNG(filename, w, h, rgba);
    return NewImageTexture(w, h, rgba);
}

// ------------------------------------------------------------------ mesh.go

struct Face {  // mesh.go:12-17
    int VertexIndices[3] = {0, 0, 0};
    int NormalIndices[3] = {0, 0, 0};
    UV UVs[3];
    TexturePtr Texture;
};

struct Mesh {  // mesh.go:19-26
    std::string Name;
    std::vector<Vec4> Vertices, VertexNormals, FaceNormals;
    Vec4 BoundingBox[8];
    std::vector<Face> Faces;
};
using MeshPtr = std::shared_ptr<Mesh>;

inline void boundingBox(const std::vector<Vec4> &v, Vec4 out[8]) {  // mesh.go:28-51
    float minX = v[0].X, minY = v[0].Y, minZ = v[0].Z, maxX = minX, maxY = minY, maxZ = minZ;
    for (const Vec4 &p : v) {
        minX = std::min(minX, p.X); minY = std::min(minY, p.Y); minZ = std::min(minZ, p.Z);
        maxX = std::max(maxX, p.X); maxY = std::max(maxY, p.Y); maxZ = std::max(maxZ, p.Z);
    }
    const Vec4 c[8] = {{minX, minY, minZ, 1}, {minX, minY, maxZ, 1}, {minX, maxY, minZ, 1}, {minX, maxY, maxZ, 1},
                       {maxX, minY, minZ, 1}, {maxX, minY, maxZ, 1}, {maxX, maxY, minZ, 1}, {maxX, maxY, maxZ, 1}};
    std::copy(c, c + 8, out);
}

inline MeshPtr NewMesh(std::vector<Vec4> vertices, std::vector<Vec4> vertexNormals, std::vector<Face> faces) {  // mesh.go:53-69
    auto m = std::make_shared<Mesh>();
    m->FaceNormals.resize(faces.size());
    for (size_t i = 0; i < faces.size(); i++) {
        for (int k = 0; k < 3; k++)
            if (faces[i].VertexIndices[k] < 0 || faces[i].VertexIndices[k] >= (int)vertices.size())
                throw std::out_of_range("index out of range");  // Go: runtime panic
        const Vec4 &a = vertices[faces[i].VertexIndices[0]], &b = vertices[faces[i].VertexIndices[1]], &c = vertices[faces[i].VertexIndices[2]];
        Vec3 v0{a.X, a.Y, a.Z}, v1{b.X, b.Y, b.Z}, v2{c.X, c.Y, c.Z};
        Vec3 n = Normalize(CrossProduct(Sub(v1, v0), Sub(v2, v0)));
        m->FaceNormals[i] = {n.X, n.Y, n.Z, 1};
    }
    m->Faces = std::move(faces);
    m->Vertices = std::move(vertices);
    m->VertexNormals = std::move(vertexNormals);
    boundingBox(m->Vertices, m->BoundingBox);
    return m;
}

struct Object {  // mesh.go:71-79 (the per-frame scratch slices live in HBM)
    MeshPtr mesh;
    Vec3 Rotation, Translation, Scale{1, 1, 1};
};
using ObjectPtr = std::shared_ptr<Object>;
inline ObjectPtr NewObject(MeshPtr mesh) {  // mesh.go:81-89
    auto o = std::make_shared<Object>();
    o->mesh = std::move(mesh);
    return o;
}

// ------------------------------------------------------------------ obj.go

namespace detail {
inline std::string trim(const std::string &s) {
    size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
    return b == std::string::npos ? "" : s.substr(b, e - b + 1);
}
inline bool hasPrefix(const std::string &s, const char *p) { return s.rfind(p, 0) == 0; }
inline int count(const std::string &s, const std::string &sub) {
    int n = 0;
    for (size_t p = s.find(sub); p != std::string::npos; p = s.find(sub, p + sub.size())) n++;
    return n;
}
inline std::string dirOf(const std::string &f) {
    size_t p = f.find_last_of('/');
    return p == std::string::npos ? "." : f.substr(0, p);
}
// `%f` into a float32 (strconv-grade: strtof, not double-then-narrow)
inline bool scanFloats(const std::string &line, const char *prefix, float *out, int n) {
    const char *p = line.c_str() + std::strlen(prefix);
    for (int i = 0; i < n; i++) {
        char *end = nullptr;
        out[i] = std::strtof(p, &end);
        if (end == p) return false;
        p = end;
    }
    return true;
}
}  // namespace detail

struct ObjMaterial {  // obj.go:14-17
    std::string Name, MapKd;
};

inline std::vector<ObjMaterial> parseMtlLibFile(const std::string &filename) {  // obj.go:153-192
    std::ifstream f(filename);
    if (!f) throw std::runtime_error("open " + filename + ": no such file or directory");
    std::vector<ObjMaterial> mats;
    bool have = false;
    ObjMaterial cur;
    std::string line;
    while (std::getline(f, line)) {
        line = detail::trim(line);
        if (line.empty()) continue;
        if (detail::hasPrefix(line, "newmtl ")) {
            if (have) mats.push_back(cur);
            cur = ObjMaterial{line.substr(7), ""};
            have = true;
        } else if (detail::hasPrefix(line, "map_Kd ")) {
            if (!have) throw std::runtime_error("map_Kd before newmtl");  // Go: nil pointer panic
            cur.MapKd = line.substr(7);
        }
    }
    if (have) mats.push_back(cur);
    return mats;
}

// LoadObjFile (obj.go:196-309) incl. the index offsets of multi-object files (obj.go:31-40) and
// the `v//vn` quirk (obj.go:77-89: the third normal index lands in vn1, vn2 stays 0).
inline std::vector<MeshPtr> LoadObjFile(const std::string &filename, bool singleMesh) {
    std::ifstream f(filename);
    if (!f) throw std::runtime_error("open " + filename + ": no such file or directory");
    const std::string dirname = detail::dirOf(filename);
    TexturePtr defaultTexture = NewColorTexture(255, 0, 255, 255);  // obj.go:208
    TexturePtr currentTexture;
    std::vector<Vec4> vertices, normals;
    std::vector<UV> tverts;
    std::vector<Face> faces;
    int vOff = 0, vtOff = 0, vnOff = 0;
    std::map<std::string, TexturePtr> textures, textureFiles;
    std::vector<MeshPtr> meshes;
    auto flush = [&]() {
        meshes.push_back(NewMesh(vertices, normals, faces));
        vOff += (int)vertices.size(); vtOff += (int)tverts.size(); vnOff += (int)normals.size();  // ObjContext.Clear
        vertices.clear(); normals.clear(); tverts.clear(); faces.clear();
    };
    auto uvAt = [&](int idx) -> UV {
        if (idx < 0 || idx >= (int)tverts.size()) throw std::out_of_range("index out of range");
        return tverts[idx];
    };
    std::string raw;
    while (std::getline(f, raw)) {
        std::string line = detail::trim(raw);
        if (line.empty()) continue;
        if (detail::hasPrefix(line, "mtllib ")) {
            std::vector<ObjMaterial> mats;
            try { mats = parseMtlLibFile(dirname + "/" + line.substr(7)); }
            catch (const std::exception &e) { throw std::runtime_error(std::string("failed to parse material library: ") + e.what()); }
            for (const ObjMaterial &m : mats) {
                if (m.MapKd.empty()) textures[m.Name] = defaultTexture;
                else if (textureFiles.count(m.MapKd)) textures[m.Name] = textureFiles[m.MapKd];
                else {
                    std::string path = m.MapKd[0] == '/' ? m.MapKd : dirname + "/" + m.MapKd;
                    TexturePtr t;
                    try { t = LoadTextureFile(path); }
                    catch (const std::exception &e) { throw std::runtime_error(std::string("failed to load texture: ") + e.what()); }
                    textureFiles[m.MapKd] = t;
                    textures[m.Name] = t;
                }
            }
        } else if (detail::hasPrefix(line, "o ")) {
            if (!vertices.empty() && !singleMesh) flush();
        } else if (detail::hasPrefix(line, "v ")) {
            float v[3];
            if (!detail::scanFloats(line, "v ", v, 3)) throw std::runtime_error("unexpected EOF");
            vertices.push_back({v[0], v[1], v[2], 1});
        } else if (detail::hasPrefix(line, "vt ")) {
            float v[2];
            if (!detail::scanFloats(line, "vt ", v, 2)) throw std::runtime_error("unexpected EOF");
            tverts.push_back({v[0], v[1]});
        } else if (detail::hasPrefix(line, "vn ")) {
            float v[3];
            if (!detail::scanFloats(line, "vn ", v, 3)) throw std::runtime_error("unexpected EOF");
            normals.push_back({v[0], v[1], v[2], 1});
        } else if (detail::hasPrefix(line, "usemtl ")) {
            auto it = textures.find(line.substr(7));
            currentTexture = it == textures.end() ? nullptr : it->second;
        } else if (detail::hasPrefix(line, "f ")) {
            if (detail::count(line, " ") != 3) throw std::runtime_error("mesh is not triangulated");
            Face face;
            int v[3] = {0, 0, 0}, vt[3] = {0, 0, 0}, vn[3] = {0, 0, 0};
            if (detail::count(line, "//") == 3) {
                int n0 = 0, n1 = 0, n2 = 0;
                if (std::sscanf(line.c_str(), "f %d//%d %d//%d %d//%d", &v[0], &n0, &v[1], &n1, &v[2], &n2) != 6)
                    throw std::runtime_error("input does not match format");
                vn[0] = n0; vn[1] = n2; vn[2] = 0;  // Sscanf(..., &vn0, ..., &vn1, ..., &vn1)
                for (int k = 0; k < 3; k++) { face.VertexIndices[k] = v[k] - vOff - 1; face.NormalIndices[k] = vn[k] - vnOff - 1; }
            } else if (detail::count(line, "/") == 3) {
                if (std::sscanf(line.c_str(), "f %d/%d %d/%d %d/%d", &v[0], &vt[0], &v[1], &vt[1], &v[2], &vt[2]) != 6)
                    throw std::runtime_error("input does not match format");
                for (int k = 0; k < 3; k++) { face.VertexIndices[k] = v[k] - vOff - 1; face.UVs[k] = uvAt(vt[k] - vtOff - 1); }
            } else if (detail::count(line, "/") == 6) {
                if (std::sscanf(line.c_str(), "f %d/%d/%d %d/%d/%d %d/%d/%d", &v[0], &vt[0], &vn[0], &v[1], &vt[1], &vn[1], &v[2], &vt[2], &vn[2]) != 9)
                    throw std::runtime_error("input does not match format");
                for (int k = 0; k < 3; k++) {
                    face.VertexIndices[k] = v[k] - vOff - 1;
                    face.UVs[k] = uvAt(vt[k] - vtOff - 1);
                    face.NormalIndices[k] = vn[k] - vnOff - 1;
                }
            } else {
                if (std::sscanf(line.c_str(), "f %d %d %d", &v[0], &v[1], &v[2]) != 3) throw std::runtime_error("input does not match format");
                for (int k = 0; k < 3; k++) face.VertexIndices[k] = v[k] - vOff - 1;
            }
            face.Texture = currentTexture;
            faces.push_back(face);
        }
    }
    if (!vertices.empty()) flush();
    if (meshes.empty()) throw std::runtime_error("obj file does not have any vertices data");
    return meshes;
}

inline std::vector<MeshPtr> LoadMeshFile(const std::string &filename, bool singleMesh) {  // mesh.go:91-103
    size_t dot = filename.find_last_of('.');
    std::string ext = dot == std::string::npos ? "" : filename.substr(dot);
    if (ext == ".obj") return LoadObjFile(filename, singleMesh);
    throw std::runtime_error("unsupported mesh format: " + ext);
}

struct Scene {  // scene.go:32-54
    std::vector<ObjectPtr> Objects;
    int NumObjects() const { return (int)Objects.size(); }
    int NumVertices() const { int n = 0; for (auto &o : Objects) n += (int)o->mesh->Vertices.size(); return n; }
    int NumTriangles() const { int n = 0; for (auto &o : Objects) n += (int)o->mesh->Faces.size(); return n; }
};

// ------------------------------------------------------------------ rasterizer.go / renderer.go

struct Camera {  // renderer.go:22-26
    Vec3 Position, Direction, Up;
};

class Device {  // one grb_context == one GPU; owns uploaded assets
public:
    explicit Device(int ordinal = 0) {
        if (grb_context_create(ordinal, &ctx_) != GRB_OK) throw std::runtime_error(std::string("gorender_b200: ") + grb_last_error(nullptr));
    }
    ~Device() { grb_context_destroy(ctx_); }
    Device(const Device &) = delete;
    grb_context *ctx() const { return ctx_; }
    void check(int32_t rc, const char *what) const {
        if (rc != GRB_OK) throw std::runtime_error(std::string("gorender_b200: ") + what + ": " + grb_last_error(ctx_));
    }
    int32_t textureID(const TexturePtr &t) {
        if (!t) return -1;
        auto it = textures_.find(t.get());
        if (it != textures_.end()) return it->second;
        int32_t id = -1;
        check(grb_texture_upload(ctx_, t->typ, t->width, t->height, t->scale, t->color, t->pixels.empty() ? nullptr : t->pixels.data(), &id),
              "grb_texture_upload");
        keepT_.push_back(t);
        return textures_[t.get()] = id;
    }
    int32_t meshID(const MeshPtr &m) {  // flattens Faces (mesh.go:12-17) once
        auto it = meshes_.find(m.get());
        if (it != meshes_.end()) return it->second;
        return uploadMesh(m, false);
    }
    // NewMesh (mesh.go:53-69) with the face normals and the bounding box computed by the GPU while
    // the mesh is uploaded (grb_mesh_new), then copied back into the Mesh fields.
    MeshPtr NewMesh(std::vector<Vec4> vertices, std::vector<Vec4> vertexNormals, std::vector<Face> faces) {
        auto m = std::make_shared<Mesh>();
        m->Faces = std::move(faces);
        m->Vertices = std::move(vertices);
        m->VertexNormals = std::move(vertexNormals);
        m->FaceNormals.resize(m->Faces.size());
        const int32_t id = uploadMesh(m, true);
        check(grb_mesh_read_derived(ctx_, id, m->FaceNormals.empty() ? nullptr : &m->FaceNormals[0].X, &m->BoundingBox[0].X),
              "grb_mesh_read_derived");
        return m;
    }

    // LoadObjFile (obj.go:196-309) with the parsing done by the library's native parser (grb_obj_parse) and
    // NewMesh done on the device: file -> GPU without a per-line Sscanf or a per-face host loop.  Textures are
    // decoded here, as in the reference's host code.  Same meshes, bit for bit, as LoadObjFile + NewMesh.
    std::vector<MeshPtr> LoadObjFileNative(const std::string &filename, bool singleMesh) {
        grb_obj *obj = nullptr;
        char err[512] = {0};
        if (grb_obj_parse(filename.c_str(), singleMesh ? 1 : 0, &obj, err, (int32_t)sizeof(err)) != GRB_OK)
            throw std::runtime_error(err);
        struct Guard { grb_obj *o; ~Guard() { grb_obj_free(o); } } guard{obj};
        std::vector<TexturePtr> sources;
        for (int32_t i = 0; i < grb_obj_num_textures(obj); i++) {
            const std::string path = grb_obj_texture_path(obj, i);
            if (path.empty()) sources.push_back(NewColorTexture(255, 0, 255, 255));   // obj.go:208
            else {
                try { sources.push_back(LoadTextureFile(path)); }
                catch (const std::exception &e) { throw std::runtime_error(std::string("failed to load texture: ") + e.what()); }
            }
        }
        std::vector<MeshPtr> meshes;
        for (int32_t i = 0; i < grb_obj_num_meshes(obj); i++) {
            grb_mesh_desc d{};
            check(grb_obj_mesh(obj, i, &d), "grb_obj_mesh");
            std::vector<Vec4> vertices(d.nv), normals(d.nvn);
            for (int32_t k = 0; k < d.nv; k++) vertices[k] = {d.vertices[4 * k], d.vertices[4 * k + 1], d.vertices[4 * k + 2], d.vertices[4 * k + 3]};
            for (int32_t k = 0; k < d.nvn; k++) normals[k] = {d.vnormals[4 * k], d.vnormals[4 * k + 1], d.vnormals[4 * k + 2], d.vnormals[4 * k + 3]};
            std::vector<Face> faces(d.nf);
            for (int32_t f = 0; f < d.nf; f++) {
                for (int k = 0; k < 3; k++) {
                    faces[f].VertexIndices[k] = d.vidx[3 * f + k];
                    faces[f].NormalIndices[k] = d.nidx[3 * f + k];
                    faces[f].UVs[k] = {d.uvs[6 * f + 2 * k], d.uvs[6 * f + 2 * k + 1]};
                }
                if (d.tex[f] >= 0) faces[f].Texture = sources[d.tex[f]];
            }
            meshes.push_back(NewMesh(std::move(vertices), std::move(normals), std::move(faces)));
        }
        return meshes;
    }

private:
    int32_t uploadMesh(const MeshPtr &m, bool derive) {
        const size_t nf = m->Faces.size();
        std::vector<int32_t> vidx(3 * nf), nidx(3 * nf), tex(nf);
        std::vector<float> uvs(6 * nf);
        for (size_t i = 0; i < nf; i++) {
            const Face &f = m->Faces[i];
            for (int k = 0; k < 3; k++) {
                vidx[3 * i + k] = f.VertexIndices[k];
                nidx[3 * i + k] = f.NormalIndices[k];
                uvs[6 * i + 2 * k] = f.UVs[k].U;
                uvs[6 * i + 2 * k + 1] = f.UVs[k].V;
            }
            tex[i] = textureID(f.Texture);
        }
        grb_mesh_desc d{};
        d.nv = (int32_t)m->Vertices.size(); d.nvn = (int32_t)m->VertexNormals.size(); d.nf = (int32_t)nf;
        d.vertices = &m->Vertices[0].X;
        d.vnormals = d.nvn ? &m->VertexNormals[0].X : nullptr;
        d.fnormals = nf ? &m->FaceNormals[0].X : nullptr;
        d.vidx = nf ? vidx.data() : nullptr;
        d.nidx = (nf && d.nvn) ? nidx.data() : nullptr;
        d.uvs = nf ? uvs.data() : nullptr;
        d.tex = nf ? tex.data() : nullptr;
        std::memcpy(d.bbox, m->BoundingBox, sizeof(d.bbox));
        int32_t id = -1;
        if (derive) check(grb_mesh_new(ctx_, &d, &id), "grb_mesh_new");
        else check(grb_mesh_upload(ctx_, &d, &id), "grb_mesh_upload");
        keepM_.push_back(m);
        return meshes_[m.get()] = id;
    }

    grb_context *ctx_ = nullptr;
    std::map<const Texture *, int32_t> textures_;
    std::map<const Mesh *, int32_t> meshes_;
    std::vector<TexturePtr> keepT_;
    std::vector<MeshPtr> keepM_;
};

// Pinned (page-locked) host array: the read-back of a frame runs at PCIe speed only into pinned
// memory, so FrameBuffer keeps Pixels / Pixels2 / ZBuffer there (grb_host_alloc).
template <typename T> class PinnedArray {
public:
    explicit PinnedArray(size_t n) : n_(n), p_(static_cast<T *>(grb_host_alloc(n * sizeof(T)))) {
        if (!p_) throw std::bad_alloc();
        std::memset(p_, 0, n * sizeof(T));
    }
    ~PinnedArray() { grb_host_free(p_); }
    PinnedArray(const PinnedArray &) = delete;
    PinnedArray &operator=(const PinnedArray &) = delete;
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
    void swap(PinnedArray &o) { std::swap(p_, o.p_); std::swap(n_, o.n_); }

private:
    size_t n_;
    T *p_;
};

struct FrameBuffer {  // rasterizer.go:7-23
    int Width, Height;
    PinnedArray<float> ZBuffer;
    PinnedArray<uint8_t> Pixels, Pixels2;  // RGBA8; Pixels is what Draw fills, Pixels2 what is presented
    Device *dev;
    // Two device framebuffers: `handle` belongs to Pixels, `handle2` to Pixels2, and they swap
    // together, so a frame still crossing PCIe into Pixels2 never blocks the draw into Pixels
    // (the reference overlaps render and present the same way, main.go:198-227).
    grb_framebuffer *handle = nullptr, *handle2 = nullptr;
    // Host mirrors: the three host planes are kept exact tile by tile — a Draw moves only the tiles that are
    // busy now or were busy in the plane, not 7.4 MB per 720p frame (grb_mirror_*, include/gorender_b200.h).
    grb_mirror *mirror = nullptr, *mirror2 = nullptr, *mirrorZ = nullptr;
    FrameBuffer(Device &d, int width, int height)
        : Width(width), Height(height), ZBuffer((size_t)width * height), Pixels((size_t)width * height * 4),
          Pixels2((size_t)width * height * 4), dev(&d) {
        d.check(grb_framebuffer_create(d.ctx(), width, height, 1, &handle), "grb_framebuffer_create");
        d.check(grb_framebuffer_create(d.ctx(), width, height, 1, &handle2), "grb_framebuffer_create");
        d.check(grb_mirror_create(d.ctx(), width, height, 1, GRB_PLANE_COLOR, Pixels.data(), &mirror), "grb_mirror_create");
        d.check(grb_mirror_create(d.ctx(), width, height, 1, GRB_PLANE_COLOR, Pixels2.data(), &mirror2), "grb_mirror_create");
        d.check(grb_mirror_create(d.ctx(), width, height, 1, GRB_PLANE_DEPTH, ZBuffer.data(), &mirrorZ), "grb_mirror_create");
    }
    ~FrameBuffer() {
        grb_mirror_destroy(mirror);
        grb_mirror_destroy(mirror2);
        grb_mirror_destroy(mirrorZ);
        grb_framebuffer_destroy(handle);
        grb_framebuffer_destroy(handle2);
    }
    FrameBuffer(const FrameBuffer &) = delete;
    // rasterizer.go:32-34.  Does not wait: after DrawAsync + SwapBuffers the frame is on its way
    // into Pixels2; WaitFront() blocks until it has landed.
    void SwapBuffers() {
        Pixels.swap(Pixels2);
        std::swap(handle, handle2);
        std::swap(mirror, mirror2);
    }
    void WaitFront() { dev->check(grb_mirror_wait(mirror2), "grb_mirror_wait"); }
};

class Renderer {  // renderer.go:83-164
public:
    bool FrustumClipping = true, ShowVertices = false, ShowEdges = false, ShowFaces = true, BackfaceCulling = true,
         Lighting = true, FlatShading = false, ShowTextures = true;  // renderer.go:130-137
    // renderer.go:476-480: `if !demoMode { CrossHair; // Fog(0.100, 0.033, {100,100,100,255}) }` —
    // compile-time switches in the reference (main.go:22), run-time fields here
    bool CrossHair = false, Fog = false;
    float FogStart = 0.100f, FogEnd = 0.033f;
    uint8_t FogColor[4] = {100, 100, 100, 255};
    int TPF = 0;

    explicit Renderer(FrameBuffer &fb, bool parallel = true) : fb_(fb) {
        aspectX_ = (float)fb.Width / (float)fb.Height;   // renderer.go:115-122
        fovY_ = (float)(45 * (M_PI / 180));
        zNear_ = 0.0f; zFar_ = 50.0f;
        numTiles_ = parallel ? 16 : 1;                   // renderer.go:144,151
    }

    uint32_t options() const {
        uint32_t o = 0;
        if (FrustumClipping) o |= GRB_OPT_FRUSTUM_CLIPPING;
        if (ShowFaces) o |= GRB_OPT_SHOW_FACES;
        if (BackfaceCulling) o |= GRB_OPT_BACKFACE_CULLING;
        if (Lighting) o |= GRB_OPT_LIGHTING;
        if (FlatShading) o |= GRB_OPT_FLAT_SHADING;
        if (ShowTextures) o |= GRB_OPT_SHOW_TEXTURES;
        if (ShowEdges) o |= GRB_OPT_SHOW_EDGES;
        if (ShowVertices) o |= GRB_OPT_SHOW_VERTICES;
        if (CrossHair) o |= GRB_OPT_CROSSHAIR;
        if (Fog) o |= GRB_OPT_FOG;
        return o;
    }

    // renderer.go:255-262
    void objectMatrices(const Object &o, const Camera &cam, Matrix &world, Matrix &mvp) const {
        world = NewWorldMatrix(o.Scale, o.Rotation, o.Translation);
        Matrix view = NewViewMatrix(cam.Position, cam.Direction, cam.Up);
        Matrix persp = NewPerspectiveMatrix(fovY_, aspectX_, zNear_, zFar_);
        mvp = NewIdentityMatrix();
        mvp = Multiply(mvp, persp);
        mvp = Multiply(mvp, view);
        mvp = Multiply(mvp, world);
    }

    // Streaming form: queue the draw and the read-back of pixels into fb.Pixels and return.  The
    // frame is complete after fb.SwapBuffers(); fb.WaitFront().  Depth stays on the device and TPF
    // is not updated (use Draw for those).
    void DrawAsync(const std::vector<ObjectPtr> &objects, const Camera &camera) {
        Device &dev = *fb_.dev;
        grb_draw_params p{};
        pack(objects, camera, p);
        dev.check(grb_draw_async(dev.ctx(), fb_.handle, 0, 1, objs_.empty() ? nullptr : objs_.data(), (int32_t)objs_.size(), &p), "grb_draw_async");
        dev.check(grb_mirror_update_async(dev.ctx(), fb_.handle, 0, 1, fb_.mirror, 0, nullptr, 0), "grb_mirror_update_async");
    }

    // renderer.go:443-483: side effects on fb.Pixels, fb.ZBuffer, TPF; throws where Go would panic
    void Draw(const std::vector<ObjectPtr> &objects, const Camera &camera) {
        Device &dev = *fb_.dev;
        grb_draw_params p{};
        pack(objects, camera, p);
        grb_frame_stats st{};
        // one call, one synchronisation: draw + host planes + stats (a CUDA graph replay after the first frame)
        dev.check(grb_draw_present(dev.ctx(), fb_.handle, 0, 1, objs_.empty() ? nullptr : objs_.data(), (int32_t)objs_.size(), &p,
                                   fb_.mirror, 0, fb_.mirrorZ, 0, &st),
                  "grb_draw_present");
        TPF = (int)st.tpf;
    }

private:
    void pack(const std::vector<ObjectPtr> &objects, const Camera &camera, grb_draw_params &p) {
        Device &dev = *fb_.dev;
        objs_.resize(objects.size());
        for (size_t i = 0; i < objects.size(); i++) {
            Matrix world, mvp;
            objectMatrices(*objects[i], camera, world, mvp);
            objs_[i].mesh = dev.meshID(objects[i]->mesh);
            std::memcpy(objs_[i].world, world.data(), 64);
            std::memcpy(objs_[i].mvp, mvp.data(), 64);
        }
        Matrix screen = NewScreenMatrix(fb_.Width, fb_.Height);     // renderer.go:264
        std::memcpy(p.screen, screen.data(), 64);
        Vec3 light = Normalize(Vec3{-1, 1, 1});                     // renderer.go:265
        p.light[0] = light.X; p.light[1] = light.Y; p.light[2] = light.Z;
        p.options = options();
        p.z_near = zNear_; p.z_far = zFar_;
        p.ref_tiles = numTiles_;
        p.fog_start = FogStart; p.fog_end = FogEnd;
        std::memcpy(p.fog_color, FogColor, 4);
    }

    FrameBuffer &fb_;
    float aspectX_, fovY_, zNear_, zFar_;
    int numTiles_;
    std::vector<grb_object> objs_;
};

}  // namespace gorender
