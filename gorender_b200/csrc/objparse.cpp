// objparse.cpp — LoadObjFile's parsing (obj.go:196-309) as native host code behind the C ABI
// (`fmt.Sscanf` per line is what makes 200 k – 2 M face files slow to load).
//
// One pass over the whole file held in memory, hand-rolled tokenising, and `%f` into a float32 as a
// correctly rounded decimal -> binary32 conversion like Go's strconv (an exact fast path, strtof
// for everything it does not cover; never a blind double-then-narrow).
// Follows obj.go line by line: TrimSpace, prefix dispatch (`mtllib `, `o `, `v `, `vt `, `vn `,
// `usemtl `, `f `), the four face syntaxes chosen by counting `/` (obj.go:60-151), triangulated
// faces only, the per-object index offsets of multi-object files (obj.go:31-40), and the `v//vn`
// quirk (obj.go:77-89: the third normal index is scanned into vn1, so NormalIndices[2] stays -1 -
// offset).  What is NOT done here, on purpose: decoding textures (image.Decode stays the caller's,
// texture.go:91-103) and NewMesh's derived arrays (grb_mesh_new computes them on the GPU).  A face
// carries an index into the file's table of texture sources instead of a *Texture.

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gorender_b200.h"

namespace {

struct ObjMesh {
    std::vector<float> vertices, vnormals, uvs;   // xyzw, xyzw, 6 per face
    std::vector<int32_t> vidx, nidx, tex;
};

struct TextureSource {
    std::string path;   // "" = the default solid magenta texture (obj.go:208)
};

}  // namespace

struct grb_obj {
    std::vector<ObjMesh> meshes;
    std::vector<TextureSource> sources;
};

namespace {

struct Span {
    const char *b, *e;
    size_t size() const { return (size_t)(e - b); }
    bool has_prefix(const char *p) const {
        const size_t n = std::strlen(p);
        return size() >= n && std::memcmp(b, p, n) == 0;
    }
    std::string str(size_t skip = 0) const { return std::string(b + skip, e); }
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

inline Span trim(Span s) {   // strings.TrimSpace
    while (s.b < s.e && is_space(*s.b)) s.b++;
    while (s.e > s.b && is_space(s.e[-1])) s.e--;
    return s;
}

inline int count_char(Span s, char c) {
    int n = 0;
    for (const char *p = s.b; p < s.e; p++) n += *p == c;
    return n;
}
inline int count_double_slash(Span s) {   // strings.Count(line, "//"), non-overlapping
    int n = 0;
    for (const char *p = s.b; p + 1 < s.e; p++)
        if (p[0] == '/' && p[1] == '/') { n++; p++; }
    return n;
}

bool read_file(const std::string &name, std::string &out, std::string &err) {
    FILE *f = std::fopen(name.c_str(), "rb");
    if (!f) {
        err = "open " + name + ": " + (errno == ENOENT ? "no such file or directory" : std::strerror(errno));
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? std::fread(&out[0], 1, (size_t)n, f) : 0;
    std::fclose(f);
    out.resize(got);
    return true;
}

// Decimal -> float32, correctly rounded.  Fast path (Clinger): at most 15 significant digits and
// |exponent| <= 22 make `mantissa * 10^e` or `mantissa / 10^e` ONE correctly rounded double
// operation on exact operands; narrowing that double to float equals rounding the decimal directly
// unless the double sits exactly on a float32 rounding midpoint (then the decimal's own low digits
// decide) — those, and anything unusual (more digits, hex, inf/nan, subnormal or overflowing
// results), go to strtof.  Returns the end of the token, or nullptr if there is no number.
const char *parse_f32(const char *p, const char *e, float &out) {
    static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char *q = p;
    bool neg = false;
    if (q < e && (*q == '+' || *q == '-')) neg = *q++ == '-';
    uint64_t mant = 0;
    int digits = 0, exp10 = 0;
    bool any = false, simple = true;
    while (q < e && *q >= '0' && *q <= '9') {
        any = true;
        if (mant != 0 || *q != '0') { mant = mant * 10 + (uint64_t)(*q - '0'); digits++; }
        if (digits > 15) simple = false;
        q++;
    }
    if (q < e && *q == '.') {
        q++;
        while (q < e && *q >= '0' && *q <= '9') {
            any = true;
            if (mant != 0 || *q != '0') { mant = mant * 10 + (uint64_t)(*q - '0'); digits++; }
            if (digits > 15) simple = false;
            exp10--;
            q++;
        }
    }
    if (any && simple && q < e && (*q == 'e' || *q == 'E')) {
        const char *r = q + 1;
        bool eneg = false;
        if (r < e && (*r == '+' || *r == '-')) eneg = *r++ == '-';
        if (r < e && *r >= '0' && *r <= '9') {
            int x = 0;
            while (r < e && *r >= '0' && *r <= '9' && x < 10000) x = x * 10 + (*r++ - '0');
            exp10 += eneg ? -x : x;
            q = r;
        }
    }
    // what follows must end the token: otherwise (hex floats, "inf", "1_000", ...) let strtof decide
    const bool ends = q >= e || *q == ' ' || *q == '\t' || *q == '\r' || *q == '\n';
    if (any && simple && ends && exp10 >= -22 && exp10 <= 22) {
        const double d = exp10 < 0 ? (double)mant / kPow10[-exp10] : (double)mant * kPow10[exp10];
        uint64_t bits;
        std::memcpy(&bits, &d, 8);
        const bool midpoint = (bits & ((1ull << 29) - 1)) == (1ull << 28);
        if (mant == 0 || (!midpoint && d > 1e-30 && d < 1e30)) {
            out = (float)(neg ? -d : d);
            return q;
        }
    }
    char *end = nullptr;
    out = std::strtof(p, &end);   // the buffer is NUL-terminated and strtof stops at the line's '\n'
    if (end == p || end > e) return nullptr;
    return end;
}

// `%f` x n after the prefix: whitespace-separated tokens.
bool scan_floats(Span line, size_t prefix, float *out, int n) {
    const char *p = line.b + prefix;
    for (int i = 0; i < n; i++) {
        while (p < line.e && (*p == ' ' || *p == '\t')) p++;
        if (p >= line.e) return false;
        p = parse_f32(p, line.e, out[i]);
        if (p == nullptr) return false;
    }
    return true;
}

// `%d`: optional sign, decimal digits.
bool scan_int(const char *&p, const char *e, long &v) {
    const char *q = p;
    bool neg = false;
    if (q < e && (*q == '+' || *q == '-')) neg = *q++ == '-';
    if (q >= e || *q < '0' || *q > '9') return false;
    // Go's Sscanf("%d") into an int fails with "value out of range" beyond int64; an OBJ index that large cannot refer
    // to anything anyway, so saturate well inside `long` (no signed overflow) and let the range checks reject it
    long x = 0;
    while (q < e && *q >= '0' && *q <= '9') {
        if (x < (1L << 52)) x = x * 10 + (*q - '0');
        q++;
    }
    v = neg ? -x : x;
    p = q;
    return true;
}
bool expect(const char *&p, const char *e, char c) {
    if (p < e && *p == c) { p++; return true; }
    return false;
}

std::string dir_of(const std::string &f) {   // path.Dir
    const size_t p = f.find_last_of('/');
    if (p == std::string::npos) return ".";
    return p == 0 ? "/" : f.substr(0, p);
}

struct Material { std::string name, map_kd; };

// parseMtlLibFile (obj.go:153-192)
bool parse_mtl(const std::string &filename, std::vector<Material> &mats, std::string &err) {
    std::string text;
    if (!read_file(filename, text, err)) return false;
    bool have = false;
    Material cur;
    const char *p = text.data(), *end = p + text.size();
    while (p < end) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        Span line = trim({p, nl ? nl : end});
        p = nl ? nl + 1 : end;
        if (line.size() == 0) continue;
        if (line.has_prefix("newmtl ")) {
            if (have) mats.push_back(cur);
            cur = {line.str(7), ""};
            have = true;
        } else if (line.has_prefix("map_Kd ")) {
            if (!have) { err = "map_Kd before newmtl"; return false; }   // Go: nil pointer dereference
            cur.map_kd = line.str(7);
        }
    }
    if (have) mats.push_back(cur);
    return true;
}

// ---- the parser: one serial pass that classifies the lines and runs the statements whose effect depends on
// file order (`mtllib`, `usemtl`, `o`), then the bulk — the floats of every `v` / `vt` / `vn` line and the
// indices of every `f` line — parsed by a few threads.  An `o` statement starts a new mesh when the current one
// has vertices (obj.go:257-262), and the index offsets of a mesh (obj.go:31-40) are simply the numbers of `v`,
// `vt` and `vn` lines before it, so no parsed value is needed to place any other.  The first error in file
// order wins, as it would in the reference's line-by-line loop.

struct Segment {             // one mesh: the lines between two effective `o` statements
    size_t v0, vt0, vn0, f0;   // first v / vt / vn / f line (global numbering) == ObjContext offsets
    size_t v1, vt1, vn1, f1;   // one past the last
};

struct FaceLine {
    Span line;
    size_t ordinal;          // statement number in the file (for "first error wins")
    int32_t tex, seg;
};
struct FloatLine {
    Span line;
    size_t ordinal;
};

struct ErrorSlot {           // smallest ordinal wins
    std::mutex mu;
    size_t ordinal = SIZE_MAX;
    std::string msg;
    void report(size_t ord, const std::string &m) {
        std::lock_guard<std::mutex> g(mu);
        if (ord < ordinal) { ordinal = ord; msg = m; }
    }
};

template <typename Fn> void parallel_for(size_t n, Fn fn) {
    const size_t kMinPerThread = 20000;
    size_t threads = std::min<size_t>(std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())), n / kMinPerThread);
    if (threads <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; t++) pool.emplace_back([=] { fn(n * t / threads, n * (t + 1) / threads); });
    for (auto &th : pool) th.join();
}

// parseFace (obj.go:60-151) for one line of segment `sg`; tv = all parsed texture vertices (u, v), global numbering.
// Returns an empty string or the error.
std::string parse_face(Span line, const Segment &sg, const float *tv, int32_t *vidx, int32_t *nidx, float *uvs) {
    if (count_char(line, ' ') != 3) return "mesh is not triangulated";
    const long vOff = (long)sg.v0, vtOff = (long)sg.vt0, vnOff = (long)sg.vn0;
    const int slashes = count_char(line, '/');
    long v[3] = {0, 0, 0}, vt[3] = {0, 0, 0}, vn[3] = {0, 0, 0};
    const char *p = line.b + 2, *e = line.e;
    bool ok = true, has_vt = false;
    if (count_double_slash(line) == 3) {
        long n[3] = {0, 0, 0};
        for (int k = 0; k < 3 && ok; k++) {
            if (k) ok = expect(p, e, ' ');
            ok = ok && scan_int(p, e, v[k]) && expect(p, e, '/') && expect(p, e, '/') && scan_int(p, e, n[k]);
        }
        // Sscanf targets are (&vn0, &vn1, &vn1): vn1 takes the third value, vn2 stays 0
        vn[0] = n[0]; vn[1] = n[2]; vn[2] = 0;
        for (int k = 0; k < 3; k++) vn[k] = vn[k] - vnOff - 1;
    } else if (slashes == 3) {
        has_vt = true;
        for (int k = 0; k < 3 && ok; k++) {
            if (k) ok = expect(p, e, ' ');
            ok = ok && scan_int(p, e, v[k]) && expect(p, e, '/') && scan_int(p, e, vt[k]);
        }
    } else if (slashes == 6) {
        has_vt = true;
        for (int k = 0; k < 3 && ok; k++) {
            if (k) ok = expect(p, e, ' ');
            ok = ok && scan_int(p, e, v[k]) && expect(p, e, '/') && scan_int(p, e, vt[k]) && expect(p, e, '/') &&
                 scan_int(p, e, vn[k]);
        }
        for (int k = 0; k < 3; k++) vn[k] = vn[k] - vnOff - 1;
    } else {
        for (int k = 0; k < 3 && ok; k++) {
            if (k) ok = expect(p, e, ' ');
            ok = ok && scan_int(p, e, v[k]);
        }
    }
    if (!ok) return "unexpected input in face statement";   // the Sscanf error
    for (int k = 0; k < 6; k++) uvs[k] = 0.0f;
    if (has_vt)
        for (int k = 0; k < 3; k++) {
            // c.TextureVertices holds the texture vertices read SO FAR in this mesh: a forward reference panics too
            const long idx = vt[k] - vtOff - 1;
            if (idx < 0 || (size_t)idx >= sg.vt1 - sg.vt0) return "index out of range";   // Go: runtime panic
            uvs[2 * k] = tv[2 * (sg.vt0 + (size_t)idx)];
            uvs[2 * k + 1] = tv[2 * (sg.vt0 + (size_t)idx) + 1];
        }
    for (int k = 0; k < 3; k++) {
        // the indices are checked against the mesh when it is uploaded (the reference panics on use, renderer.go:318-330):
        // keep anything that does not fit the ABI's int32 out of range instead of letting it wrap into a valid index
        const long vi = v[k] - vOff - 1, ni = vn[k];
        vidx[k] = (vi < INT32_MIN || vi > INT32_MAX) ? INT32_MAX : (int32_t)vi;
        nidx[k] = (ni < INT32_MIN || ni > INT32_MAX) ? INT32_MAX : (int32_t)ni;
    }
    return "";
}

struct Parser {
    grb_obj *out;
    bool single;
    std::string dirname, err;
    std::map<std::string, int> materialTex;    // c.Textures: material name -> texture source
    std::map<std::string, int> fileTex;        // textureFiles: map_Kd -> texture source
    int defaultTex = -1, currentTex = -1;

    bool fail(const std::string &m) { err = m; return false; }

    int source(const std::string &path) {
        out->sources.push_back({path});
        return (int)out->sources.size() - 1;
    }

    bool mtllib(const std::string &name) {   // obj.go:222-255
        std::vector<Material> mats;
        std::string e;
        if (!parse_mtl(dirname + "/" + name, mats, e)) return fail("failed to parse material library: " + e);
        for (const Material &m : mats) {
            if (m.map_kd.empty()) {
                if (defaultTex < 0) defaultTex = source("");
                materialTex[m.name] = defaultTex;
            } else {
                auto it = fileTex.find(m.map_kd);
                if (it == fileTex.end()) {
                    std::string path = m.map_kd[0] == '/' ? m.map_kd : dirname + "/" + m.map_kd;
                    it = fileTex.emplace(m.map_kd, source(path)).first;
                }
                materialTex[m.name] = it->second;
            }
        }
        return true;
    }

    bool run(const std::string &filename) {
        std::string text;
        if (!read_file(filename, text, err)) return false;
        dirname = dir_of(filename);

        // ---- pass 1 (serial): classify, run the order-dependent statements
        std::vector<FloatLine> vLines, vtLines, vnLines;
        std::vector<FaceLine> fLines;
        std::vector<Segment> segs;
        Segment cur{0, 0, 0, 0, 0, 0, 0, 0};
        ErrorSlot first;
        size_t ordinal = 0;
        const char *p = text.data(), *end = p + text.size();
        while (p < end) {
            const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
            Span line = trim({p, nl ? nl : end});
            p = nl ? nl + 1 : end;
            if (line.size() == 0) continue;
            ordinal++;
            if (line.b[0] == 'v' && line.has_prefix("v ")) {
                vLines.push_back({line, ordinal});
            } else if (line.b[0] == 'f' && line.has_prefix("f ")) {
                fLines.push_back({line, ordinal, currentTex, (int32_t)segs.size()});
            } else if (line.has_prefix("vt ")) {
                vtLines.push_back({line, ordinal});
            } else if (line.has_prefix("vn ")) {
                vnLines.push_back({line, ordinal});
            } else if (line.has_prefix("mtllib ")) {
                if (!mtllib(line.str(7))) {   // the reference stops here; earlier lines may still fail first
                    first.report(ordinal, err);
                    break;
                }
            } else if (line.has_prefix("o ")) {
                if (vLines.size() != cur.v0 && !single) {   // NewMesh + ObjContext.Clear (obj.go:31-40, 257-262)
                    cur.v1 = vLines.size(); cur.vt1 = vtLines.size(); cur.vn1 = vnLines.size(); cur.f1 = fLines.size();
                    segs.push_back(cur);
                    cur = {cur.v1, cur.vt1, cur.vn1, cur.f1, 0, 0, 0, 0};
                }
            } else if (line.has_prefix("usemtl ")) {
                auto it = materialTex.find(line.str(7));
                currentTex = it == materialTex.end() ? -1 : it->second;   // unknown name -> nil texture
            }
        }
        cur.v1 = vLines.size(); cur.vt1 = vtLines.size(); cur.vn1 = vnLines.size(); cur.f1 = fLines.size();
        const bool tail = cur.v1 != cur.v0;     // obj.go:296-299: the last mesh needs vertices
        if (tail) segs.push_back(cur);

        // ---- pass 2 (parallel): the floats
        std::vector<float> V(4 * vLines.size()), VN(4 * vnLines.size()), VT(2 * vtLines.size());
        auto floats = [&](const std::vector<FloatLine> &lines, size_t prefix, int n, int stride, float w, std::vector<float> &dst) {
            parallel_for(lines.size(), [&, prefix, n, stride, w](size_t b, size_t e) {
                float f[3];
                for (size_t i = b; i < e; i++) {
                    if (!scan_floats(lines[i].line, prefix, f, n)) { first.report(lines[i].ordinal, "unexpected EOF"); continue; }
                    for (int k = 0; k < n; k++) dst[stride * i + k] = f[k];
                    if (stride == 4) dst[4 * i + 3] = w;   // Vec4{x, y, z, 1} (obj.go:45, 57)
                }
            });
        };
        floats(vLines, 2, 3, 4, 1.0f, V);
        floats(vtLines, 3, 2, 2, 0.0f, VT);
        floats(vnLines, 3, 3, 4, 1.0f, VN);

        // ---- pass 3 (parallel): the faces (they read the texture vertices parsed above)
        const size_t nf = fLines.size();
        std::vector<int32_t> VI(3 * nf), NI(3 * nf);
        std::vector<float> UV(6 * nf);
        parallel_for(nf, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; i++) {
                const FaceLine &fl = fLines[i];
                // (a face after the last mesh with vertices lands in no mesh, but is parsed — and can fail — all the same)
                Segment sg = (size_t)fl.seg < segs.size() ? segs[fl.seg] : cur;
                // only the texture vertices that precede the face in the file exist yet (obj.go:107-109)
                size_t lo = sg.vt0, hi = sg.vt1;
                while (lo < hi) {   // number of vt lines of this mesh before this face
                    const size_t mid = (lo + hi) / 2;
                    if (vtLines[mid].ordinal < fl.ordinal) lo = mid + 1; else hi = mid;
                }
                sg.vt1 = lo;
                const std::string m = parse_face(fl.line, sg, VT.data(), &VI[3 * i], &NI[3 * i], &UV[6 * i]);
                if (!m.empty()) first.report(fl.ordinal, m);
            }
        });
        if (first.ordinal != SIZE_MAX) return fail(first.msg);

        // ---- meshes
        for (size_t k = 0; k < segs.size(); k++) {
            const Segment &sg = segs[k];
            ObjMesh m;
            m.vertices.assign(V.begin() + 4 * sg.v0, V.begin() + 4 * sg.v1);
            m.vnormals.assign(VN.begin() + 4 * sg.vn0, VN.begin() + 4 * sg.vn1);
            m.vidx.assign(VI.begin() + 3 * sg.f0, VI.begin() + 3 * sg.f1);
            m.nidx.assign(NI.begin() + 3 * sg.f0, NI.begin() + 3 * sg.f1);
            m.uvs.assign(UV.begin() + 6 * sg.f0, UV.begin() + 6 * sg.f1);
            m.tex.resize(sg.f1 - sg.f0);
            for (size_t i = sg.f0; i < sg.f1; i++) m.tex[i - sg.f0] = fLines[i].tex;
            out->meshes.push_back(std::move(m));
        }
        if (out->meshes.empty()) return fail("obj file does not have any vertices data");
        return true;
    }
};

void put_error(char *err, int32_t cap, const std::string &m) {
    if (err && cap > 0) std::snprintf(err, (size_t)cap, "%s", m.c_str());
}

}  // namespace

extern "C" {

int32_t grb_obj_parse(const char *filename, int32_t single_mesh, grb_obj **out, char *err, int32_t err_cap) {
    if (!filename || !out) {
        put_error(err, err_cap, "null argument");
        return GRB_ERR_INVALID;
    }
    *out = nullptr;
    grb_obj *o = new grb_obj;
    Parser ps;
    ps.out = o;
    ps.single = single_mesh != 0;
    if (!ps.run(filename)) {
        put_error(err, err_cap, ps.err);
        delete o;
        return GRB_ERR_INVALID;
    }
    *out = o;
    return GRB_OK;
}

int32_t grb_obj_num_meshes(const grb_obj *o) { return o ? (int32_t)o->meshes.size() : 0; }
int32_t grb_obj_num_textures(const grb_obj *o) { return o ? (int32_t)o->sources.size() : 0; }

const char *grb_obj_texture_path(const grb_obj *o, int32_t i) {
    if (!o || i < 0 || i >= (int32_t)o->sources.size()) return nullptr;
    return o->sources[i].path.c_str();
}

int32_t grb_obj_mesh(const grb_obj *o, int32_t i, grb_mesh_desc *d) {
    if (!o || !d || i < 0 || i >= (int32_t)o->meshes.size()) return GRB_ERR_INVALID;
    const ObjMesh &m = o->meshes[i];
    std::memset(d, 0, sizeof(*d));
    d->nv = (int32_t)(m.vertices.size() / 4);
    d->nvn = (int32_t)(m.vnormals.size() / 4);
    d->nf = (int32_t)m.tex.size();
    d->vertices = m.vertices.data();
    d->vnormals = d->nvn ? m.vnormals.data() : nullptr;
    d->fnormals = nullptr;   // NewMesh's job: grb_mesh_new derives them (and bbox) on the device
    d->vidx = d->nf ? m.vidx.data() : nullptr;
    d->nidx = d->nf ? m.nidx.data() : nullptr;
    d->uvs = d->nf ? m.uvs.data() : nullptr;
    d->tex = d->nf ? m.tex.data() : nullptr;
    return GRB_OK;
}

void grb_obj_free(grb_obj *o) { delete o; }

}  // extern "C"
