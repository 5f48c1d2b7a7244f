// rb_scene.cpp -- .oct loader, modifier resolution and flattening to
// device-ready tables.  See rb_scene.hpp for the reference files restated.
#include "rb_scene.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <charconv>
#include <fstream>
#include <functional>
#include <memory>
#include <sstream>

namespace rb {

static const double FTINY = 1e-6;
static const double PI = 3.14159265358979323846;

// ---------------------------------------------------------------- types ----
struct TypeInfo { const char* name; int ot; };
static const TypeInfo kTypes[] = {
    {"polygon", OT_POLYGON}, {"cone", OT_CONE}, {"sphere", OT_SPHERE},
    {"ring", OT_RING}, {"cylinder", OT_CYLINDER}, {"cup", OT_CUP},
    {"bubble", OT_BUBBLE}, {"tube", OT_TUBE}, {"source", OT_SOURCE},
    {"instance", OT_INSTANCE}, {"mesh", OT_MESH}, {"alias", OT_ALIAS},
    {"plastic", OT_PLASTIC}, {"metal", OT_METAL}, {"glass", OT_GLASS},
    {"trans", OT_TRANS}, {"glow", OT_GLOW}, {"light", OT_LIGHT},
    {"illum", OT_ILLUM}, {"spotlight", OT_SPOTLIGHT},
    {"dielectric", OT_DIELECTRIC}, {"interface", OT_INTERFACE},
    {"mist", OT_MIST}, {"aBSDF", OT_ABSDF}, {"trans2", OT_TRANS2},
    {"antimatter", OT_ANTIMATTER},
    // other materials (src/radiance/common/otypes.h:127-186, T_M entries)
    {"plastic2", OT_PLASTIC2}, {"metal2", OT_METAL2},
    {"plasfunc", OT_OTHER_MATERIAL}, {"metfunc", OT_OTHER_MATERIAL},
    {"mirror", OT_OTHER_MATERIAL}, {"transfunc", OT_OTHER_MATERIAL},
    {"BRTDfunc", OT_OTHER_MATERIAL}, {"BSDF", OT_BSDF},
    {"WGMDfunc", OT_OTHER_MATERIAL}, {"plasdata", OT_OTHER_MATERIAL},
    {"metdata", OT_OTHER_MATERIAL}, {"transdata", OT_OTHER_MATERIAL},
    {"prism1", OT_OTHER_MATERIAL}, {"prism2", OT_OTHER_MATERIAL},
    {"ashik2", OT_OTHER_MATERIAL},
    // patterns
    {"brightfunc", OT_PATTERN}, {"brightdata", OT_PATTERN},
    {"brighttext", OT_PATTERN}, {"colorpict", OT_PATTERN},
    {"colorfunc", OT_PATTERN}, {"colordata", OT_PATTERN},
    {"colortext", OT_PATTERN}, {"spectrum", OT_PATTERN},
    {"specfile", OT_PATTERN}, {"specfunc", OT_PATTERN},
    {"specdata", OT_PATTERN}, {"specpict", OT_PATTERN},
    // textures, mixtures
    {"texfunc", OT_TEXTURE}, {"texdata", OT_TEXTURE},
    {"mixfunc", OT_MIXTURE}, {"mixdata", OT_MIXTURE},
    {"mixtext", OT_MIXTURE}, {"mixpict", OT_MIXTURE},
};

int ot_from_name(const std::string& s) {
    for (const auto& t : kTypes)
        if (s == t.name) return t.ot;
    return OT_OTHER;
}
bool ot_is_surface(int t) {
    return t == OT_POLYGON || t == OT_CONE || t == OT_SPHERE || t == OT_RING ||
           t == OT_CYLINDER || t == OT_CUP || t == OT_BUBBLE || t == OT_TUBE ||
           t == OT_SOURCE;
}
bool ot_is_volume(int t) { return t == OT_INSTANCE || t == OT_MESH; }
bool ot_is_material(int t) {
    switch (t) {
    case OT_PLASTIC: case OT_METAL: case OT_GLASS: case OT_TRANS: case OT_GLOW:
    case OT_LIGHT: case OT_ILLUM: case OT_SPOTLIGHT: case OT_DIELECTRIC:
    case OT_INTERFACE: case OT_MIST: case OT_ABSDF: case OT_TRANS2:
    case OT_ANTIMATTER: case OT_OTHER_MATERIAL: case OT_PLASTIC2: case OT_METAL2: case OT_BSDF:
        return true;
    }
    return false;
}
bool ot_is_light(int t) {
    return t == OT_GLOW || t == OT_LIGHT || t == OT_ILLUM || t == OT_SPOTLIGHT;
}
bool ot_is_modifier(int t) { return !ot_is_surface(t) && !ot_is_volume(t); }

// ------------------------------------------------------ portable binary ----
namespace {
struct Rd {
    const unsigned char* p;
    const unsigned char* e;
    bool bad = false;
    int getc_() {
        if (p >= e) { bad = true; return -1; }
        return *p++;
    }
    // portio.c:116-133 getint(): big-endian, sign-extended from first byte
    long getint(int siz) {
        int c = getc_();
        if (c < 0) return -1;
        long r = c;
        if (c & 0x80) r |= -256L;
        while (--siz > 0) {
            c = getc_();
            if (c < 0) return -1;
            r = (long)((unsigned long)r << 8);
            r |= c;
        }
        return r;
    }
    // portio.c:136-152 getflt(): 4-byte mantissa + 1-byte exponent
    double getflt() {
        long l = getint(4);
        if (bad) return 0;
        if (l == 0) { getc_(); return 0.0; }
        double d = (l + .5 - (l < 0)) * (1. / 0x7fffffff);
        return ldexp(d, (int)getint(1));
    }
    bool getstr(std::string& s) {
        s.clear();
        for (;;) {
            int c = getc_();
            if (c < 0) return false;
            if (c == 0) return true;
            s.push_back((char)c);
        }
    }
};
}  // namespace

// --------------------------------------------------------------- loader ----
// sceneio.c:90-109 readscene() + :20-87 getobj(): type-name table then objects
static bool read_frozen_scene(Rd& rd, int objsize, std::vector<Object>& objs, std::string& err) {
    std::string s;
    std::vector<int> tmap; std::vector<std::string> tnames;
    for (;;) {
        if (!rd.getstr(s)) { err = "truncated scene"; return false; }
        if (s.empty()) break;
        tnames.push_back(s);
        tmap.push_back(ot_from_name(s));
    }
    for (;;) {
        long ti = rd.getint(1);
        if (rd.bad) { err = "unexpected EOF in scene"; return false; }
        if (ti == -1) break;
        if (ti < 0 || ti >= (long)tmap.size()) { err = "reference to unknown type"; return false; }
        Object o;
        o.otype = tmap[ti]; o.tname = tnames[ti];
        o.omod = (int)rd.getint(objsize);
        rd.getstr(o.name);
        long ns = rd.getint(2);
        for (long i = 0; i < ns; i++) { rd.getstr(s); o.sargs.push_back(s); }
        long nf = rd.getint(2);
        o.fargs.resize(nf > 0 ? nf : 0);
        for (long i = 0; i < nf; i++) o.fargs[i] = rd.getflt();
        if (rd.bad) { err = "unexpected EOF in scene"; return false; }
        objs.push_back(std::move(o));
    }
    return true;
}

static int read_tree(Rd& rd, Scene& sc, int objsize, int depth, std::string& err) {
    if (depth > sc.maxdepth) sc.maxdepth = depth;
    int c = rd.getc_();
    switch (c) {
    case 0:  // OT_EMPTY
        return -1;
    case 1: {  // OT_FULL: readoct.c:153-168 getfullnode()
        long n = rd.getint(objsize);
        if (rd.bad || n < 0 || n > 8191) { err = "bad set in octree"; return -1; }
        int off = (int)sc.leafpool.size();
        sc.leafpool.push_back((int)n);
        for (long i = 0; i < n; i++) sc.leafpool.push_back((int)rd.getint(objsize));
        if (rd.bad) { err = "truncated octree"; return -1; }
        return -(off) - 2;
    }
    case 2: {  // OT_TREE: readoct.c:195-218 gettree()
        if (depth > 60) { err = "octree too deep"; return -1; }
        int idx = (int)(sc.nodes.size() / 8);
        sc.nodes.resize(sc.nodes.size() + 8, -1);
        for (int i = 0; i < 8; i++) {
            int k = read_tree(rd, sc, objsize, depth + 1, err);
            if (!err.empty()) return -1;
            sc.nodes[(size_t)idx * 8 + i] = k;
        }
        return idx;
    }
    default:
        err = (c < 0) ? "truncated octree" : "damaged octree";
        return -1;
    }
}

bool Scene::load_octree(const std::string& path) {
    error.clear();
    std::ifstream f(path, std::ios::binary);
    if (!f) { error = "cannot open octree file \"" + path + "\""; return false; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const unsigned char* b = (const unsigned char*)data.data();
    const unsigned char* e = b + data.size();
    // ---- info header (common/header.c: lines until an empty line) ----
    const unsigned char* p = b;
    bool gotfmt = false, first = true;
    header.clear();
    for (;;) {
        const unsigned char* nl = (const unsigned char*)memchr(p, '\n', e - p);
        if (!nl) { error = "(" + path + "): not an octree"; return false; }
        std::string line((const char*)p, nl - p);
        p = nl + 1;
        if (line.empty()) break;
        if (first) {
            first = false;
            if (line.compare(0, 2, "#?") != 0) { error = "(" + path + "): not an octree"; return false; }
            header.push_back(line);
            continue;
        }
        if (line.compare(0, 7, "FORMAT=") == 0) {
            gotfmt = (line.find("Radiance_octree") != std::string::npos);
            continue;
        }
        header.push_back(line);
    }
    if (!gotfmt) { error = "(" + path + "): not an octree"; return false; }
    Rd rd{p, e};
    const int OCTMAGIC = 4 * 8 + 251;
    int objsize = (int)rd.getint(2) - OCTMAGIC;
    if (objsize <= 0 || objsize > 8) { error = "(" + path + "): incompatible octree format"; return false; }
    std::string s;
    for (int i = 0; i < 3; i++) { rd.getstr(s); cuorg[i] = atof(s.c_str()); }
    rd.getstr(s); cusize = atof(s.c_str());
    srcfiles.clear();
    for (;;) {
        if (!rd.getstr(s)) { error = "(" + path + "): truncated octree"; return false; }
        if (s.empty()) break;
        srcfiles.push_back(s);
    }
    frozen = srcfiles.empty();
    long nobj = rd.getint(objsize);
    if (rd.bad || nobj < 0) { error = "(" + path + "): truncated octree"; return false; }
    nodes.clear(); leafpool.clear(); maxdepth = 0;
    std::string terr;
    root = read_tree(rd, *this, objsize, 0, terr);
    if (!terr.empty()) { error = "(" + path + "): " + terr; return false; }
    objs.clear();
    if (frozen) {
        std::string rerr;
        if (!read_frozen_scene(rd, objsize, objs, rerr)) { error = "(" + path + "): " + rerr; return false; }
        if ((long)objs.size() != nobj) {
            error = "(" + path + "): bad object count in frozen octree"; return false;
        }
    } else {
        // octree refers to scene files: parse them as text (readoct.c:90-100)
        std::string dir;
        size_t sl = path.rfind('/');
        if (sl != std::string::npos) dir = path.substr(0, sl + 1);
        for (const auto& fn : srcfiles) {
            std::string full = (fn[0] == '/' || fn[0] == '!') ? fn : fn;
            if (!read_rad_text(full)) {
                if (dir.empty() || !read_rad_text(dir + fn)) return false;
                error.clear();
            }
        }
        if ((long)objs.size() != nobj) {
            error = "(" + path + "): bad object count; octree stale?"; return false;
        }
    }
    index_modifiers();
    {
        std::string dir;
        size_t sl = path.rfind('/');
        if (sl != std::string::npos) dir = path.substr(0, sl + 1);
        basedir = dir;
        if (!expand_volumes(dir)) return false;
    }
    return true;
}

// getpath(fname, getrlibpath(), R_OK) (common/getpath.c): absolute or explicitly
// relative names as they are, else the RAYPATH directories; as a last resort the
// directory of the referring octree.
std::string find_radiance_file(const std::string& name, const std::string& basedir) {
    auto readable = [](const std::string& p) { std::ifstream f(p); return (bool)f; };
    if (name.empty()) return "";
    if (name[0] == '/' || name[0] == '.') return readable(name) ? name : (readable(basedir + name) ? basedir + name : "");
    const char* rp = getenv("RAYPATH");
    std::string path = rp ? rp : ".:/usr/local/lib/ray";
    size_t i = 0;
    while (i <= path.size()) {
        size_t j = path.find(':', i);
        if (j == std::string::npos) j = path.size();
        std::string d = path.substr(i, j - i);
        if (d.empty()) d = ".";
        if (readable(d + "/" + name)) return d + "/" + name;
        i = j + 1;
    }
    if (readable(basedir + name)) return basedir + name;
    return "";
}

namespace {
struct Xf { double m[4][4]; double sca; };
static void xf_ident(double m[4][4]) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = (i == j); }
static void xf_mul(double a[4][4], double b[4][4], double c[4][4]) {
    double t[4][4];
    for (int i = 4; i--;) for (int j = 4; j--;)
        t[i][j] = b[i][0] * c[0][j] + b[i][1] * c[1][j] + b[i][2] * c[2][j] + b[i][3] * c[3][j];
    memcpy(a, t, sizeof(t));
}
// common/xf.c:38-139 xf(): -t -rx -ry -rz -s -mx -my -mz -i
static bool parse_xf(const std::vector<std::string>& av, size_t i0, Xf& ret, std::string& err) {
    const double D2R = 3.14159265358979323846 / 180.;
    xf_ident(ret.m); ret.sca = 1.0;
    double xfmat[4][4], m4[4][4], xfsca = 1.0;
    int icnt = 1;
    xf_ident(xfmat);
    size_t i = i0;
    auto num = [&](size_t k, double& v) { if (k >= av.size()) return false; char* e; v = strtod(av[k].c_str(), &e); return e != av[k].c_str() && !*e; };
    for (; i < av.size() && av[i][0] == '-'; i++) {
        xf_ident(m4);
        const std::string& a = av[i];
        double v[3];
        if (a == "-t") { if (!num(i + 1, v[0]) || !num(i + 2, v[1]) || !num(i + 3, v[2])) break; m4[3][0] = v[0]; m4[3][1] = v[1]; m4[3][2] = v[2]; i += 3; }
        else if (a == "-rx") { if (!num(i + 1, v[0])) break; double d = D2R * v[0]; m4[1][1] = m4[2][2] = cos(d); m4[2][1] = -(m4[1][2] = sin(d)); i++; }
        else if (a == "-ry") { if (!num(i + 1, v[0])) break; double d = D2R * v[0]; m4[0][0] = m4[2][2] = cos(d); m4[0][2] = -(m4[2][0] = sin(d)); i++; }
        else if (a == "-rz") { if (!num(i + 1, v[0])) break; double d = D2R * v[0]; m4[0][0] = m4[1][1] = cos(d); m4[1][0] = -(m4[0][1] = sin(d)); i++; }
        else if (a == "-s") { if (!num(i + 1, v[0]) || v[0] == 0.0) break; xfsca *= m4[0][0] = m4[1][1] = m4[2][2] = v[0]; i++; }
        else if (a == "-mx") { xfsca *= m4[0][0] = -1.0; }
        else if (a == "-my") { xfsca *= m4[1][1] = -1.0; }
        else if (a == "-mz") { xfsca *= m4[2][2] = -1.0; }
        else if (a == "-i") {
            if (!num(i + 1, v[0])) break;
            while (icnt-- > 0) { xf_mul(ret.m, ret.m, xfmat); ret.sca *= xfsca; }
            icnt = (int)v[0]; xf_ident(xfmat); xfsca = 1.0; i++;
            continue;
        } else break;
        xf_mul(xfmat, xfmat, m4);
    }
    while (icnt-- > 0) { xf_mul(ret.m, ret.m, xfmat); ret.sca *= xfsca; }
    if (i != av.size()) { err = "bad transform"; return false; }
    return true;
}
static void xf_point(double r[3], const double p[3], const Xf& x) {
    double t[3];
    for (int j = 0; j < 3; j++) t[j] = p[0] * x.m[0][j] + p[1] * x.m[1][j] + p[2] * x.m[2][j] + x.m[3][j];
    r[0] = t[0]; r[1] = t[1]; r[2] = t[2];
}
static void xf_vector(double r[3], const double p[3], const Xf& x) {
    double t[3];
    for (int j = 0; j < 3; j++) t[j] = p[0] * x.m[0][j] + p[1] * x.m[1][j] + p[2] * x.m[2][j];
    r[0] = t[0]; r[1] = t[1]; r[2] = t[2];
}
}  // namespace


namespace {
// A compiled triangle mesh (.rtm): common/readmesh.c:124-305, common/mesh.h.
struct MeshFile {
    std::vector<Object> mats;                 // mesh-local materials (frozen scene)
    struct Tri { double v[3][3]; int mat; bool smooth; double n[3][3]; };  // mesh-space vertices, local material
    std::vector<Tri> tris;                    // index (-1 void), vertex normals when all three vertices carry one
};

// common/dircode.c:56-79 decodedir(): the 4-byte direction code of a vertex normal
static void decode_dir(double dv[3], int32_t dc) {
    const double DCSCALE = 11584.5;
    enum { FXNEG = 01, FYNEG = 02, FZNEG = 04, F1X = 010, F2Z = 020, F1SFT = 5, F2SFT = 18, FMASK = 0x1fff };
    if (!dc) { dv[0] = dv[1] = dv[2] = 0.; return; }
    const double d1 = (dc >> F1SFT & FMASK) * (1. / DCSCALE);
    const double d2 = (dc >> F2SFT & FMASK) * (1. / DCSCALE);
    const double der = sqrt(1. - d1 * d1 - d2 * d2);
    if (dc & F1X) {
        dv[0] = d1;
        if (dc & F2Z) { dv[1] = der; dv[2] = d2; } else { dv[1] = d2; dv[2] = der; }
    } else {
        dv[1] = d1;
        if (dc & F2Z) { dv[0] = der; dv[2] = d2; } else { dv[0] = d2; dv[2] = der; }
    }
    if (dc & FXNEG) dv[0] = -dv[0];
    if (dc & FYNEG) dv[1] = -dv[1];
    if (dc & FZNEG) dv[2] = -dv[2];
}

static bool load_rtm(const std::string& path, MeshFile& mf, std::string& err) {
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open mesh file \"" + path + "\""; return false; }
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const unsigned char* b = (const unsigned char*)data.data();
    const unsigned char* e = b + data.size();
    const unsigned char* p = b;
    bool gotfmt = false;
    for (;;) {
        const unsigned char* nl = (const unsigned char*)memchr(p, '\n', e - p);
        if (!nl) { err = "(" + path + "): not a mesh"; return false; }
        std::string line((const char*)p, nl - p);
        p = nl + 1;
        if (line.empty()) break;
        if (line.compare(0, 7, "FORMAT=") == 0) gotfmt = line.find("Radiance_tmesh") != std::string::npos;
    }
    if (!gotfmt) { err = "(" + path + "): not a mesh"; return false; }
    Rd rd{p, e};
    const int MESHMAGIC = 1 * 8 + 311;
    int objsize = (int)rd.getint(2) - MESHMAGIC;
    if (objsize <= 0 || objsize > 8) { err = "(" + path + "): incompatible mesh format"; return false; }
    double cuorg[3], cusize;
    std::string s;
    for (int i = 0; i < 3; i++) { rd.getstr(s); cuorg[i] = atof(s.c_str()); }
    rd.getstr(s); cusize = atof(s.c_str());
    for (int i = 0; i < 4; i++) rd.getflt();                       // uv limits
    Scene dummy; std::string terr;
    read_tree(rd, dummy, objsize, 0, terr);                         // the mesh's own octree is not needed
    if (!terr.empty()) { err = "(" + path + "): " + terr; return false; }
    if (!read_frozen_scene(rd, objsize, mf.mats, err)) { err = "(" + path + "): " + err; return false; }
    long npatches = rd.getint(4);
    if (rd.bad || npatches < 0) { err = "(" + path + "): truncated mesh"; return false; }
    struct Patch { std::vector<uint32_t> xyz; std::vector<int32_t> norm; int nverts; };
    std::vector<Patch> patches(npatches);
    struct RawTri { long v[3]; int mat; };
    std::vector<RawTri> raw;
    for (long pn = 0; pn < npatches; pn++) {                       // readmesh.c:124-230 getpatch()
        int flags = (int)rd.getint(1);
        if (!(flags & 1) || (flags & ~7)) { err = "(" + path + "): bad patch flags"; return false; }
        int nv = (int)rd.getint(2);
        if (nv <= 0 || nv > 256) { err = "(" + path + "): bad number of patch vertices"; return false; }
        patches[pn].nverts = nv;
        patches[pn].xyz.resize((size_t)nv * 3);
        for (int i = 0; i < nv * 3; i++) patches[pn].xyz[i] = (uint32_t)rd.getint(4);
        if (flags & 2) { patches[pn].norm.resize(nv); for (int i = 0; i < nv; i++) patches[pn].norm[i] = (int32_t)rd.getint(4); }
        if (flags & 4) for (int i = 0; i < nv * 2; i++) rd.getint(4);
        int nt = (int)rd.getint(2);
        if (nt < 0 || nt > 512) { err = "(" + path + "): bad number of local triangles"; return false; }
        size_t t0 = raw.size();
        for (int i = 0; i < nt; i++) {
            RawTri t;
            t.v[0] = pn << 8 | rd.getint(1) & 0xff; t.v[1] = pn << 8 | rd.getint(1) & 0xff; t.v[2] = pn << 8 | rd.getint(1) & 0xff;
            t.mat = -1;
            raw.push_back(t);
        }
        if (rd.getint(2) > 1) for (int i = 0; i < nt; i++) raw[t0 + i].mat = (short)rd.getint(2);
        else { int sole = (short)rd.getint(2); for (int i = 0; i < nt; i++) raw[t0 + i].mat = sole; }
        int nj1 = (int)rd.getint(2);
        if (nj1 < 0 || nj1 > 256) { err = "(" + path + "): bad number of joiner triangles"; return false; }
        for (int i = 0; i < nj1; i++) {
            RawTri t;
            t.v[0] = rd.getint(4); t.v[1] = pn << 8 | rd.getint(1) & 0xff; t.v[2] = pn << 8 | rd.getint(1) & 0xff;
            t.mat = (short)rd.getint(2);
            raw.push_back(t);
        }
        int nj2 = (int)rd.getint(2);
        if (nj2 < 0 || nj2 > 256) { err = "(" + path + "): bad number of double joiner triangles"; return false; }
        for (int i = 0; i < nj2; i++) {
            RawTri t;
            t.v[0] = rd.getint(4); t.v[1] = rd.getint(4); t.v[2] = pn << 8 | rd.getint(1) & 0xff;
            t.mat = (short)rd.getint(2);
            raw.push_back(t);
        }
        if (rd.bad) { err = "(" + path + "): truncated mesh"; return false; }
    }
    // common/mesh.c:233-262 getmeshvert(): cuorg + (q + .5) * cusize / 2^32
    const double vres = (1. / 4294967296.) * cusize;
    mf.tris.reserve(raw.size());
    for (const RawTri& t : raw) {
        MeshFile::Tri o;
        o.mat = t.mat;
        o.smooth = true;                          // mesh.c:316: flags of the three vertices ANDed
        for (int k = 0; k < 3; k++) {
            long pn = t.v[k] >> 8; int vid = (int)(t.v[k] & 0xff);
            if (pn < 0 || pn >= npatches || vid >= patches[pn].nverts) { err = "(" + path + "): bad mesh vertex reference"; return false; }
            for (int i = 0; i < 3; i++) o.v[k][i] = cuorg[i] + (patches[pn].xyz[(size_t)vid * 3 + i] + .5) * vres;
            const int32_t dc = patches[pn].norm.empty() ? 0 : patches[pn].norm[vid];      // mesh.c:258-261
            if (dc) decode_dir(o.n[k], dc);
            else { o.smooth = false; o.n[k][0] = o.n[k][1] = o.n[k][2] = 0.; }
        }
        mf.tris.push_back(o);
    }
    return true;
}
}  // namespace

// Instances (rt/o_instance.c:16-71, common/instance.c:24-101) are FLATTENED at
// load: every surface of the instanced octree becomes a world-space copy in the
// parent's object table (named and modified as the reference would report the
// hit: the instance's own name/modifier when it has a modifier, else the inner
// surface's), and the parent's octree is re-built over the flat list.  The walk
// then needs no nested traversal.  Differences: distances are computed on
// transformed geometry instead of a transformed ray (last-bit differences), and
// coincident inner/outer surfaces are arbitrated by rayreject() instead of
// "first found wins".
bool Scene::expand_volumes(const std::string& basedir, int depth) {
    const size_t n0 = objs.size();
    bool any = false;
    for (size_t i = 0; i < n0; i++) if (ot_is_volume(objs[i].otype)) any = true;
    if (!any) return true;
    if (depth > 8) { error = "instances nested too deeply"; return false; }
    struct Nested { std::shared_ptr<Scene> sc; std::vector<int> remap; };
    struct MeshNested { std::shared_ptr<MeshFile> mf; std::vector<int> remap; };
    std::unordered_map<std::string, MeshNested> meshes;
    std::unordered_map<std::string, Nested> cache;
    for (size_t i = 0; i < n0; i++) {
        if (objs[i].otype == OT_MESH) {
            // rt/o_mesh.c:146-217: triangles become world-space polygons; a hit reports the
            // pseudo object "M-Tri" with the mesh-local material, or the mesh object itself
            // when it has its own modifier (or the triangle has none)
            const Object msh = objs[i];
            if (msh.sargs.empty()) { error = "bad # of arguments for mesh \"" + msh.name + "\""; return false; }
            Xf x; std::string xe;
            if (!parse_xf(msh.sargs, 1, x, xe)) { error = xe + " for mesh \"" + msh.name + "\""; return false; }
            auto mit = meshes.find(msh.sargs[0]);
            if (mit == meshes.end()) {
                std::string path = find_radiance_file(msh.sargs[0], basedir);
                if (path.empty()) { error = "cannot find mesh file \"" + msh.sargs[0] + "\""; return false; }
                MeshNested mn;
                mn.mf = std::make_shared<MeshFile>();
                if (!load_rtm(path, *mn.mf, error)) return false;
                mn.remap.assign(mn.mf->mats.size(), -1);
                for (size_t j = 0; j < mn.mf->mats.size(); j++) {
                    Object c = mn.mf->mats[j];
                    c.omod = c.omod >= 0 ? mn.remap[c.omod] : -1;
                    mn.remap[j] = (int)objs.size();
                    objs.push_back(std::move(c));
                }
                mit = meshes.emplace(msh.sargs[0], std::move(mn)).first;
            }
            const MeshFile& mf = *mit->second.mf;
            const bool mirrored = x.sca < 0;
            for (const auto& t : mf.tris) {
                Object c;
                c.otype = OT_POLYGON; c.tname = "polygon";
                if (msh.omod < 0 && t.mat >= 0 && t.mat < (int)mit->second.remap.size()) { c.name = "M-Tri"; c.omod = mit->second.remap[t.mat]; }
                else { c.name = msh.name; c.omod = msh.omod; c.volume_obj = true; }
                c.fargs.resize(9);
                for (int k = 0; k < 3; k++) xf_point(&c.fargs[3 * (mirrored ? 2 - k : k)], t.v[k], x);
                if (t.smooth) {              // o_mesh.c:201-209: the interpolated normal goes through the forward
                    c.vnorm.resize(9);       // matrix (multv3) and is normalised after that, so the vertex normals may
                    for (int k = 0; k < 3; k++) xf_vector(&c.vnorm[3 * (mirrored ? 2 - k : k)], t.n[k], x);   // be moved first
                }
                objs.push_back(std::move(c));
            }
            objs[i].expanded = true;
            nexpanded++;
            continue;
        }
        if (objs[i].otype != OT_INSTANCE) continue;
        const Object inst = objs[i];           // copy: objs grows below
        if (inst.sargs.empty()) { error = "bad # of arguments for instance \"" + inst.name + "\""; return false; }
        Xf x;
        std::string xe;
        if (!parse_xf(inst.sargs, 1, x, xe)) { error = xe + " for instance \"" + inst.name + "\""; return false; }
        const bool mirrored = x.sca < 0;
        const double sca = fabs(x.sca);
        auto it = cache.find(inst.sargs[0]);
        if (it == cache.end()) {
            std::string path = find_radiance_file(inst.sargs[0], basedir);
            if (path.empty()) { error = "cannot find octree file \"" + inst.sargs[0] + "\""; return false; }
            Nested nn;
            nn.sc = std::make_shared<Scene>();
            if (!nn.sc->load_octree(path)) { error = nn.sc->error; return false; }
            // the nested scene's modifiers join our object table once per file
            nn.remap.assign(nn.sc->objs.size(), -1);
            for (size_t j = 0; j < nn.sc->objs.size(); j++) {
                const Object& o = nn.sc->objs[j];
                if (o.expanded || !ot_is_modifier(o.otype)) continue;
                Object c = o;
                c.omod = o.omod >= 0 ? nn.remap[o.omod] : -1;
                nn.remap[j] = (int)objs.size();
                objs.push_back(std::move(c));
            }
            it = cache.emplace(inst.sargs[0], std::move(nn)).first;
        }
        const Scene& ns = *it->second.sc;
        const std::vector<int>& remap = it->second.remap;
        for (size_t j = 0; j < ns.objs.size(); j++) {
            const Object& o = ns.objs[j];
            if (o.expanded) continue;
            if (!ot_is_surface(o.otype) || o.otype == OT_SOURCE) continue;
            Object c = o;
            if (inst.omod >= 0) { c.name = inst.name; c.omod = inst.omod; c.volume_obj = true; }      // o_instance.c:41-43
            else c.omod = o.omod >= 0 ? remap[o.omod] : -1;
            std::vector<double>& a = c.fargs;
            if (o.otype == OT_POLYGON) {
                int nv = (int)a.size() / 3;
                for (int v = 0; v < nv; v++) xf_point(&a[3 * v], &a[3 * v], x);
                if (mirrored)                       // keep the normal = M n (xform reverses vertex order too)
                    for (int v = 0; v < nv / 2; v++)
                        for (int k = 0; k < 3; k++) std::swap(a[3 * v + k], a[3 * (nv - 1 - v) + k]);
            } else if (o.otype == OT_SPHERE || o.otype == OT_BUBBLE) {
                if (a.size() == 4) { xf_point(&a[0], &a[0], x); a[3] *= sca; }
            } else if (o.otype == OT_RING) {
                if (a.size() == 8) { xf_point(&a[0], &a[0], x); xf_vector(&a[3], &a[3], x); a[6] *= sca; a[7] *= sca; }
            } else {
                if (a.size() >= 7) { xf_point(&a[0], &a[0], x); xf_point(&a[3], &a[3], x); a[6] *= sca; if (a.size() == 8) a[7] *= sca; }
            }
            objs.push_back(std::move(c));
        }
        objs[i].expanded = true;
        nexpanded++;
    }
    index_modifiers();
    // Instances and meshes are FLATTENED (one world-space copy of every inner surface per instance): memory grows
    // with instances x inner surfaces where the reference's nested traversal (o_instance.c:16-71) does not.  A forest
    // of heavy instances is refused by name instead of exhausting the host.
    {
        const char* e = getenv("RB_MAX_EXPANDED_SURFACES");
        const size_t limit = e ? (size_t)atoll(e) : (size_t)200000000;
        if (objs.size() > limit) {
            error = "instance / mesh expansion gives " + std::to_string(objs.size()) + " surfaces (limit " + std::to_string(limit) +
                    ", RB_MAX_EXPANDED_SURFACES): nested instance traversal is not built, instances are flattened at load";
            return false;
        }
    }
    std::string err;
    if (!rebuild_octree(*this, 6, 16384, err)) { error = err; return false; }
    return true;
}

// Plain-text scene parser (common/readobj.c:37-203 getobject, readfargs.c).
// "!command" lines need external generators and are rejected explicitly.
bool Scene::read_rad_text(const std::string& path) {
    if (!path.empty() && path[0] == '!') {
        error = "scene input from command \"" + path + "\" is not supported (freeze the octree with oconv -f)";
        return false;
    }
    std::string text;
    {
        FILE* fp = fopen(path.c_str(), "rb");
        if (!fp) { error = "cannot open scene file \"" + path + "\""; return false; }
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
        fclose(fp);
    }
    // one pass over the buffer: whitespace-separated words, `#` comments run to the end of the line,
    // a line that starts with `!` is a command (refused)
    const char* p = text.c_str();
    const char* const end = p + text.size();
    bool bol = true, cmdline = false;
    auto word = [&](const char*& b, const char*& e) -> bool {
        for (;;) {
            while (p < end && isspace((unsigned char)*p)) { if (*p == '\n') bol = true; p++; }
            if (p < end && *p == '#') { while (p < end && *p != '\n') p++; continue; }
            break;
        }
        if (p >= end) return false;
        if (bol && *p == '!') { cmdline = true; return false; }
        bol = false;
        b = p;
        while (p < end && !isspace((unsigned char)*p) && *p != '#') p++;
        e = p;
        return true;
    };
    auto integer = [&](long& v) -> bool {
        const char *b, *e;
        if (!word(b, e)) return false;
        char* ep;
        v = strtol(b, &ep, 10);
        return ep == e;
    };
    auto fail = [&](const std::string& msg) {
        error = cmdline ? "(" + path + "): \"!command\" lines are not supported (freeze the octree with oconv -f)" : msg;
        return false;
    };
    const char *b, *e;
    while (word(b, e)) {
        const std::string mod(b, e);
        const char *tb, *te, *nb, *ne;
        if (!word(tb, te) || !word(nb, ne)) return fail("(" + path + "): unexpected EOF");
        Object o;
        o.tname.assign(tb, te); o.otype = ot_from_name(o.tname); o.name.assign(nb, ne);
        const std::string &typ = o.tname, &name = o.name;
        if (mod == "void") o.omod = -1;
        else {
            auto mt = modtab.find(mod);
            o.omod = (mt == modtab.end()) ? -1 : mt->second;
            if (o.omod < 0) return fail("(" + path + "): undefined modifier \"" + mod + "\" for " + typ + " \"" + name + "\"");
        }
        if (o.otype == OT_ALIAS) {       // alias: "mod alias name target"
            if (!word(b, e)) return fail("(" + path + "): bad alias");
            o.sargs.emplace_back(b, e);
            modtab[o.name] = (int)objs.size();
            objs.push_back(std::move(o));
            continue;
        }
        long n;
        if (!integer(n) || n < 0) return fail("(" + path + "): bad arguments for " + typ + " \"" + name + "\"");
        for (long i = 0; i < n; i++) { if (!word(b, e)) break; o.sargs.emplace_back(b, e); }
        if (!integer(n) || n != 0) return fail("(" + path + "): bad integer arguments for \"" + name + "\"");
        if (!integer(n) || n < 0) return fail("(" + path + "): bad real arguments for \"" + name + "\"");
        o.fargs.resize(n);
        for (long i = 0; i < n; i++) {
            const char* ep = nullptr;
            if (word(b, e)) {
                auto r = std::from_chars(b, e, o.fargs[i]);           // correctly rounded like strtod, several times faster
                ep = r.ptr;
                if (r.ec != std::errc() || ep == b) { char* sp; o.fargs[i] = strtod(b, &sp); ep = sp; }   // "+1", "inf", hex ...
            }
            if (ep == nullptr || ep == b) return fail("(" + path + "): bad real argument for \"" + name + "\"");
        }
        if (ot_is_modifier(o.otype)) modtab[o.name] = (int)objs.size();
        objs.push_back(std::move(o));
    }
    if (cmdline) {
        error = "(" + path + "): \"!command\" lines are not supported (freeze the octree with oconv -f)";
        return false;
    }
    return true;
}

void Scene::index_modifiers() {
    modtab.clear();
    for (int i = 0; i < (int)objs.size(); i++)
        if (ot_is_modifier(objs[i].otype)) modtab[objs[i].name] = i;   // last wins
}

int Scene::lastmod(int before, const std::string& name) const {
    auto it = modtab.find(name);
    int i = (it == modtab.end()) ? -1 : it->second;
    if (before < 0 || i < before) return i;
    for (i = before; i-- > 0;)
        if (ot_is_modifier(objs[i].otype) && objs[i].name == name) return i;
    return -1;
}

// initotypes.c:112-145
int Scene::findmaterial(int oi) const {
    int obj = -1;
    int guard = 0;
    while (!ot_is_material(objs[oi].otype)) {
        if (++guard > 10000) return -1;
        const Object* o = &objs[oi];
        if (o->otype == OT_ALIAS && !o->sargs.empty()) {
            int ao = oi;
            if (obj < 0) obj = oi;
            do {
                if (objs[ao].sargs.empty()) obj = objs[ao].omod;
                else obj = lastmod(obj, objs[ao].sargs[0]);
                if (obj < 0) return -1;
                ao = obj;
            } while (objs[ao].otype == OT_ALIAS && ++guard < 10000);
            if (ot_is_material(objs[ao].otype)) return ao;
        }
        if (o->omod < 0) {
            if (o->otype == OT_MIXTURE) break;
            return -1;
        }
        obj = o->omod;
        oi = obj;
    }
    return oi;
}

// ------------------------------------------------------------- flatten ----
static double vnormalize(double v[3]) {     // common/fvect.c:130-157
    double d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (d == 0.0) return 0.0;
    double len;
    if ((d <= 1.0 + FTINY) & (d >= 1.0 - FTINY)) {
        len = 0.5 + 0.5 * d;
        d = 2.0 - len;
    } else {
        len = sqrt(d);
        d = 1.0 / len;
    }
    v[0] *= d; v[1] *= d; v[2] *= d;
    return len;
}
static void vcross(double r[3], const double a[3], const double b[3]) {
    double t[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    r[0] = t[0]; r[1] = t[1]; r[2] = t[2];
}
static double vdot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// deterministic getperpendicular (fvect.c:159-196 with randomize=0)
// "cannot hit source center" (srcsupp.c:109-110): is the vertex centroid inside the
// polygon?  Load-time input validation only, so the winding number of the centroid
// in the projection plane is enough (the reference uses inface(), face.c:121-162).
static bool inface_host(const double p[3], int hdr0, const double* g, int nv) {
    const int ax = (hdr0 >> 10) & 3;
    const int xi = (ax + 1) % 3, yi = (xi + 1) % 3;
    const double x = p[xi], y = p[yi];
    const double* vp = g + 6;
    double wn = 0.0;
    for (int n = 0; n < nv; n++) {
        const int n1 = (n + 1) % nv;
        const double ax0 = vp[2 * n] - x, ay0 = vp[2 * n + 1] - y;
        const double ax1 = vp[2 * n1] - x, ay1 = vp[2 * n1 + 1] - y;
        wn += atan2(ax0 * ay1 - ay0 * ax1, ax0 * ax1 + ay0 * ay1);
    }
    return fabs(wn) > PI;
}

static bool getperp(double vp[3], const double v[3]) {
    double v1[3] = {0, 0, 0};
    int i;
    for (i = 3; i--;)
        if ((-0.6 < v[i]) & (v[i] < 0.6)) break;
    if (i < 0) return false;
    v1[i] = 1.0;
    vcross(vp, v1, v);
    return vnormalize(vp) > 0.0;
}

static void mat4_ident(double m[4][4]) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = (i == j);
}
static void mat4_mul(double a[4][4], double b[4][4], double c[4][4]) {   // a = b*c
    double t[4][4];
    for (int i = 4; i--;)
        for (int j = 4; j--;)
            t[i][j] = b[i][0] * c[0][j] + b[i][1] * c[1][j] + b[i][2] * c[2][j] + b[i][3] * c[3][j];
    memcpy(a, t, sizeof(t));
}

// Every record is preceded by a 16-byte copy of the object header, so that the
// walk reaches header + plane with one dependent load from the leaf entry.
static size_t geom_alloc(FlatScene& fs, size_t n) {
    if (fs.geom.size() & 1) fs.geom.push_back(0.0);   // keep 16-byte alignment
    size_t off = fs.geom.size() + 2;
    fs.geom.resize(off + n, 0.0);
    return off;
}

// common/tmesh.c:45-93 comp_baryc(): out = {axis, tm[0][0..2], tm[1][0..2]}; false for a degenerate triangle
static bool comp_baryc(double out[7], const double* v1, const double* v2, const double* v3) {
    double va[3], vab[3], vcb[3];
    for (int k = 0; k < 3; k++) { vab[k] = v1[k] - v2[k]; vcb[k] = v3[k] - v2[k]; }
    vcross(va, vab, vcb);
    int ax = (va[1] * va[1] > va[0] * va[0]);
    if (va[2] * va[2] > va[ax] * va[ax]) ax = 2;
    const int ax0 = (ax + 1) % 3, ax1 = (ax + 2) % 3;
    out[0] = ax;
    for (int i = 0; i < 2; i++) {
        vab[0] = v1[ax0] - v2[ax0]; vcb[0] = v3[ax0] - v2[ax0];
        vab[1] = v1[ax1] - v2[ax1]; vcb[1] = v3[ax1] - v2[ax1];
        double d = vcb[0] * vcb[0] + vcb[1] * vcb[1];
        if (d <= FTINY * FTINY) return false;
        d = (vcb[0] * vab[0] + vcb[1] * vab[1]) / d;
        va[0] = vab[0] - vcb[0] * d;
        va[1] = vab[1] - vcb[1] * d;
        d = va[0] * va[0] + va[1] * va[1];
        if (d <= FTINY * FTINY) return false;
        d = 1.0 / d;
        out[1 + 3 * i] = va[0] *= d;
        out[2 + 3 * i] = va[1] *= d;
        out[3 + 3 * i] = -(v2[ax0] * va[0] + v2[ax1] * va[1]);
        const double* vt = v1; v1 = v2; v2 = v3; v3 = vt;
    }
    return true;
}

// common/face.c:35-106 getface() -> plane + 2-D projected vertices
static void flatten_face(const Object& o, FlatScene& fs, int32_t hdr[4], std::string& warn) {
    int nf = (int)o.fargs.size();
    if (nf < 9 || nf % 3) { hdr[0] = PK_UNSUPPORTED; warn = "bad # arguments for polygon \"" + o.name + "\""; return; }
    const double* va = o.fargs.data();
    int nv = nf / 3;
    auto V = [&](int i) { return va + 3 * i; };
    auto dist2 = [&](const double* a, const double* b) {
        double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
        return d0 * d0 + d1 * d1 + d2 * d2;
    };
    if (nv > 3 && dist2(V(0), V(nv - 1)) <= FTINY * FTINY) nv--;
    double norm[3] = {0, 0, 0}, v1[3], v2[3], v3[3];
    for (int k = 0; k < 3; k++) v1[k] = V(1)[k] - V(0)[k];
    for (int i = 2; i < nv; i++) {
        for (int k = 0; k < 3; k++) v2[k] = V(i)[k] - V(0)[k];
        vcross(v3, v1, v2);
        norm[0] += v3[0]; norm[1] += v3[1]; norm[2] += v3[2];
        for (int k = 0; k < 3; k++) v1[k] = v2[k];
    }
    double area = vnormalize(norm);
    double offset = 0.0;
    int ax = 0;
    if (area == 0.0) {
        warn = "zero area for polygon \"" + o.name + "\"";
        norm[0] = norm[1] = norm[2] = 0.0;
    } else {
        offset = vdot(norm, V(0));
        for (int i = 1; i < nv; i++) offset += vdot(norm, V(i));
        offset /= (double)nv;
        ax = (fabs(norm[1]) > fabs(norm[0]));
        if (fabs(norm[2]) > fabs(norm[ax])) ax = 2;
    }
    if (nv > 65535) { hdr[0] = PK_UNSUPPORTED; warn = "too many vertices"; return; }
    int xi = (ax + 1) % 3, yi = (xi + 1) % 3;
    // record: plane (4 doubles), 2-D bounding box xmin xmax ymin ymax as 4 floats rounded
    // OUTWARD (2 doubles' worth; a conservative reject only), vertices (2 doubles each)
    // a mesh triangle with vertex normals (o_mesh.c:193-209) also carries, after its vertices, the barycentric
    // coordinate matrix of common/tmesh.c:45-93 comp_baryc() (axis, tm[2][3]) and its three normals: 16 doubles
    double bary[7];
    const bool smooth = o.vnorm.size() == 9 && nv == 3 && area != 0.0 && comp_baryc(bary, V(0), V(1), V(2));
    size_t off = geom_alloc(fs, 6 + 2 * (size_t)nv + (smooth ? 16 : 0));
    double* g = &fs.geom[off];
    g[0] = norm[0]; g[1] = norm[1]; g[2] = norm[2]; g[3] = offset;
    if (smooth) {
        for (int k = 0; k < 7; k++) g[12 + k] = bary[k];
        for (int k = 0; k < 9; k++) g[19 + k] = o.vnorm[k];
    }
    double bb[4] = {1e300, -1e300, 1e300, -1e300};
    for (int i = 0; i < nv; i++) {
        double x = V(i)[xi], y = V(i)[yi];
        g[6 + 2 * i] = x; g[7 + 2 * i] = y;
        bb[0] = std::min(bb[0], x); bb[1] = std::max(bb[1], x);
        bb[2] = std::min(bb[2], y); bb[3] = std::max(bb[3], y);
    }
    {
        float fb[4];
        for (int k = 0; k < 4; k++) {
            float f = (float)bb[k];
            if (!(k & 1) && (double)f > bb[k]) f = nextafterf(f, -INFINITY);      // lower bounds round down
            if ((k & 1) && (double)f < bb[k]) f = nextafterf(f, INFINITY);        // upper bounds round up
            fb[k] = f;
        }
        memcpy(&g[4], fb, 16);
    }
    // exact axis-aligned rectangle in the projection plane? (fast inside test)
    int rect = 0;
    if (nv == 4 && area != 0.0) {
        auto X = [&](int i) { return V(i)[xi]; };
        auto Y = [&](int i) { return V(i)[yi]; };
        bool a = X(0) == X(1) && Y(1) == Y(2) && X(2) == X(3) && Y(3) == Y(0);
        bool b = Y(0) == Y(1) && X(1) == X(2) && Y(2) == Y(3) && X(3) == X(0);
        rect = (a || b) && bb[1] - bb[0] > 4 * FTINY && bb[3] - bb[2] > 4 * FTINY;
    }
    hdr[0] = PK_FACE | (ax << 10) | (rect << 12) | (smooth ? PX_SMOOTH : 0) | (nv << 16);
    hdr[3] = (int32_t)off;
}

// rt/sphere.c:27-36
static void flatten_sphere(const Object& o, FlatScene& fs, int32_t hdr[4], std::string& warn) {
    if (o.fargs.size() != 4) { hdr[0] = PK_UNSUPPORTED; warn = "bad # arguments for sphere \"" + o.name + "\""; return; }
    int kind = (o.otype == OT_SPHERE) ? PK_SPHERE : PK_BUBBLE;
    double r = o.fargs[3];
    if (r < -FTINY) { kind = (kind == PK_SPHERE) ? PK_BUBBLE : PK_SPHERE; r = -r; }
    else if (r <= FTINY) { hdr[0] = PK_UNSUPPORTED; warn = "zero radius for sphere \"" + o.name + "\""; return; }
    size_t off = geom_alloc(fs, 4);
    double* g = &fs.geom[off];
    g[0] = o.fargs[0]; g[1] = o.fargs[1]; g[2] = o.fargs[2]; g[3] = r;
    hdr[0] = kind; hdr[3] = (int32_t)off;
}

// common/cone.c:44-153 getcone() + :171-218 conexform()
static void flatten_cone(const Object& o, FlatScene& fs, int32_t hdr[4], std::string& warn) {
    int ot = o.otype;
    std::vector<double> ca = o.fargs;
    int p0, p1, r0, r1;
    if (ot == OT_CYLINDER || ot == OT_TUBE) {
        if (ca.size() != 7) goto argerr;
        if (ca[6] < -FTINY) { ot = (ot == OT_CYLINDER) ? OT_TUBE : OT_CYLINDER; ca[6] = -ca[6]; }
        else if (ca[6] <= FTINY) goto raderr;
        p0 = 0; p1 = 3; r0 = r1 = 6;
    } else {
        if (ca.size() != 8) goto argerr;
        int sgn0 = ca[6] < -FTINY ? -1 : ca[6] > FTINY ? 1 : 0;
        int sgn1 = ca[7] < -FTINY ? -1 : ca[7] > FTINY ? 1 : 0;
        if (sgn0 + sgn1 == 0) goto raderr;
        if ((sgn0 < 0) | (sgn1 < 0)) {
            if (ot == OT_RING) goto raderr;
            ot = (ot == OT_CONE) ? OT_CUP : OT_CONE;
        }
        ca[6] = ca[6] * sgn0; ca[7] = ca[7] * sgn1;
        if (ca[7] - ca[6] > FTINY) {
            if (ot == OT_RING) p0 = p1 = 0; else { p0 = 0; p1 = 3; }
            r0 = 6; r1 = 7;
        } else if (ca[6] - ca[7] > FTINY) {
            if (ot == OT_RING) p0 = p1 = 0; else { p0 = 3; p1 = 0; }
            r0 = 7; r1 = 6;
        } else {
            if (ot == OT_RING) goto raderr;
            ot = (ot == OT_CONE) ? OT_CYLINDER : OT_TUBE;
            p0 = 0; p1 = 3; r0 = r1 = 6;
        }
    }
    {
        double ad[3], al, sl;
        if (ot == OT_RING) { ad[0] = ca[3]; ad[1] = ca[4]; ad[2] = ca[5]; }
        else for (int k = 0; k < 3; k++) ad[k] = ca[p1 + k] - ca[p0 + k];
        al = vnormalize(ad);
        if (al == 0.0) { hdr[0] = PK_UNSUPPORTED; warn = "unknown orientation for \"" + o.name + "\""; return; }
        if (ot == OT_RING) { al = 0.0; sl = ca[r1] - ca[r0]; }
        else if (ot == OT_CONE || ot == OT_CUP) { sl = ca[7] - ca[6]; sl = sqrt(sl * sl + al * al); }
        else sl = al;
        // conexform
        double tm[4][4], m4[4][4], d;
        mat4_ident(tm);
        if (r0 == r1) d = 0.0; else d = ca[r0] / (ca[r1] - ca[r0]);
        for (int i = 0; i < 3; i++) tm[3][i] = d * (ca[p1 + i] - ca[p0 + i]) - ca[p0 + i];
        mat4_ident(m4);
        d = ad[1] * ad[1] + ad[2] * ad[2];
        if (d <= FTINY * FTINY) {
            m4[0][0] = 0.0; m4[0][2] = ad[0]; m4[2][0] = -ad[0]; m4[2][2] = 0.0;
        } else {
            d = sqrt(d);
            m4[0][0] = d; m4[1][0] = -ad[0] * ad[1] / d; m4[2][0] = -ad[0] * ad[2] / d;
            m4[1][1] = ad[2] / d; m4[2][1] = -ad[1] / d;
            m4[0][2] = ad[0]; m4[1][2] = ad[1]; m4[2][2] = ad[2];
        }
        mat4_mul(tm, tm, m4);
        if ((p0 != p1) & (r0 != r1)) {
            mat4_ident(m4);
            m4[2][2] = (ca[r1] - ca[r0]) / al;
            mat4_mul(tm, tm, m4);
        }
        size_t off = geom_alloc(fs, 24);
        double* g = &fs.geom[off];
        g[0] = ad[0]; g[1] = ad[1]; g[2] = ad[2]; g[3] = al;
        g[4] = ca[p0]; g[5] = ca[p0 + 1]; g[6] = ca[p0 + 2]; g[7] = sl;
        g[8] = ca[r0]; g[9] = ca[r1];
        {   // bounding sphere of the surface as 4 floats in the two spare words (centre, radius rounded up):
            // the pair loop drops rays that cannot touch it before the out-of-line cone test
            const double hl = 0.5 * al, rm = std::max(fabs(ca[r0]), fabs(ca[r1]));
            float bs[4];
            double cmax = 0;
            for (int k = 0; k < 3; k++) { const double c = ca[p0 + k] + ad[k] * hl; bs[k] = (float)c; cmax = std::max(cmax, fabs(c)); }
            const double R = sqrt(hl * hl + rm * rm);
            bs[3] = (float)((R + 2e-7 * cmax + 1e-5) * (1.0 + 2e-7) + 1e-5);      // float rounding of centre and radius, FTINY slack
            memcpy(&g[10], bs, sizeof(bs));
        }
        for (int i = 0; i < 4; i++) for (int j = 0; j < 3; j++) g[12 + i * 3 + j] = tm[i][j];
        int kind = ot == OT_CONE ? PK_CONE : ot == OT_CUP ? PK_CUP : ot == OT_CYLINDER ? PK_CYL
                 : ot == OT_TUBE ? PK_TUBE : PK_RING;
        hdr[0] = kind; hdr[3] = (int32_t)off;
        return;
    }
argerr:
    hdr[0] = PK_UNSUPPORTED; warn = "bad # arguments for \"" + o.name + "\""; return;
raderr:
    hdr[0] = PK_UNSUPPORTED; warn = "illegal radii for \"" + o.name + "\""; return;
}

// initotypes.c:45-61,70 + otspecial.h:13-20 istransp()
static bool mat_is_transp(const Object& m) {
    switch (m.otype) {
    case OT_TRANS: case OT_TRANS2: case OT_DIELECTRIC: case OT_INTERFACE:
    case OT_MIST: case OT_GLASS: case OT_ABSDF:
        return true;
    }
    if (m.tname == "WGMDfunc" && m.sargs.size() > 5 && m.sargs[5] != "0") return true;
    if (m.tname == "BRTDfunc" && m.sargs.size() > 5 &&
        (m.sargs[3] != "0" || m.sargs[4] != "0" || m.sargs[5] != "0")) return true;
    return false;
}

bool flatten_scene(const Scene& sc, FlatScene& fs, std::string& err) {
    const int n = (int)sc.objs.size();
    fs = FlatScene();
    fs.nodes = sc.nodes;
    fs.leafpool = sc.leafpool;
    fs.objhdr.assign((size_t)n * 4, 0);
    std::vector<int> matslot(n, -2);    // object -> material slot (-2 unknown)
    std::unordered_map<std::string, int> bsdf_index;     // BSDF file name -> slot in fs.bsdfs (loaded once, like SDcacheFile)
    bool load_failed = false;
    auto note_unsupported = [&](const std::string& s) {
        if (fs.unsupported_note.empty()) fs.unsupported_note = s;
    };
    // material slot for a material object (by object index)
    std::function<int(int)> slot_of;   // forward decl via std::function
    slot_of = [&](int mi) -> int {
        if (mi < 0) return -1;
        if (matslot[mi] != -2) return matslot[mi];
        const Object& m = sc.objs[mi];
        MatRec r; memset(&r, 0, sizeof(r));
        r.obj = mi; r.alt = -1;
        r.nargs = (int)std::min<size_t>(m.fargs.size(), 8);
        for (int i = 0; i < r.nargs; i++) r.a[i] = (float)m.fargs[i];
        auto need = [&](size_t k) {
            if (m.fargs.size() != k) {
                r.kind = MK_UNSUPPORTED;
                note_unsupported("bad number of arguments for " + m.tname + " \"" + m.name + "\"");
                return false;
            }
            return true;
        };
        switch (m.otype) {
        case OT_PLASTIC: r.kind = MK_PLASTIC; need(5); break;
        case OT_METAL:   r.kind = MK_METAL; need(5); break;
        case OT_TRANS:   r.kind = MK_TRANS; need(7); break;
        case OT_GLASS:
            r.kind = MK_GLASS;
            if (m.fargs.size() != 3 && m.fargs.size() != 4) {
                r.kind = MK_UNSUPPORTED; note_unsupported("bad arguments for glass \"" + m.name + "\"");
            }
            break;
        case OT_PLASTIC2: case OT_METAL2: case OT_TRANS2: {
            // aniso.c:185-326.  The orientation vector's three expressions must be numeric constants (they
            // usually are); the function file's transform is applied here (getacoords(), multv3 by fxp->xfm).
            r.kind = m.otype == OT_PLASTIC2 ? MK_PLASTIC2 : m.otype == OT_METAL2 ? MK_METAL2 : MK_TRANS2;
            if (!need(m.otype == OT_TRANS2 ? 8 : 6)) break;
            if (m.sargs.size() < 4) {
                r.kind = MK_UNSUPPORTED; note_unsupported("bad arguments for " + m.tname + " \"" + m.name + "\"");
                break;
            }
            if (!(m.fargs[4] > 1e-6) || !(m.fargs[5] > 1e-6)) {
                r.kind = MK_UNSUPPORTED; note_unsupported("roughness too small for " + m.tname + " \"" + m.name + "\"");
                break;
            }
            double u[3];
            bool konst = true;
            for (int k = 0; k < 3; k++) {
                const char* b = m.sargs[k].c_str(); char* e = nullptr;
                u[k] = strtod(b, &e);
                if (e == b || *e) konst = false;
            }
            if (!konst) {
                r.kind = MK_UNSUPPORTED;
                note_unsupported("orientation vector of " + m.tname + " \"" + m.name + "\" is not a numeric constant (.cal expressions are not built)");
                break;
            }
            Xf x; std::string xe;
            if (!parse_xf(m.sargs, 4, x, xe)) {
                r.kind = MK_UNSUPPORTED; note_unsupported(xe + " for " + m.tname + " \"" + m.name + "\"");
                break;
            }
            for (int k = 0; k < 3; k++) r.u[k] = u[0] * x.m[0][k] + u[1] * x.m[1][k] + u[2] * x.m[2][k];   // multv3
            break;
        }
        case OT_BSDF: case OT_ABSDF: {
            // m_bsdf.c:58-71,641-654,704-716.  BSDF: thick file ux uy uz funcfile [xf]; aBSDF: file ux uy uz funcfile [xf];
            // 0, 3, 6 or 9 reals (diffuse colours added to the file's front / back reflection and its transmission).
            // Thickness and up vector must be numeric constants; the function file's transform is applied here.
            const int ht = m.otype == OT_BSDF ? 1 : 0;
            r.kind = ht ? MK_BSDF : MK_ABSDF;
            if (ht && !m.sargs.empty() && m.sargs[0] != "0") r.flags |= 8;       // isBSDFproxy() (rt/otspecial.h:34-35)
            if ((int)m.sargs.size() < ht + 5 || m.fargs.size() > 9 || m.fargs.size() % 3) {
                r.kind = MK_UNSUPPORTED; note_unsupported("bad # arguments for " + m.tname + " \"" + m.name + "\"");
                break;
            }
            r.nargs = (int)m.fargs.size();
            if (m.fargs.size() == 9) { const float f8 = (float)m.fargs[8]; memcpy(&r.pad[1], &f8, 4); }
            double cv[4] = {0, 0, 0, 0};         // thick, ux, uy, uz
            bool konst = true;
            for (int k = 0; k < 4; k++) {
                if (k == 0 && !ht) continue;
                const char* b = m.sargs[k == 0 ? 0 : ht + k].c_str(); char* e = nullptr;
                cv[k] = strtod(b, &e);
                if (e == b || *e) konst = false;
            }
            if (!konst) {
                r.kind = MK_UNSUPPORTED;
                note_unsupported("thickness / up vector of " + m.tname + " \"" + m.name + "\" is not a numeric constant (.cal expressions are not built)");
                break;
            }
            Xf x; std::string xe;
            if (!parse_xf(m.sargs, ht + 5, x, xe)) {
                r.kind = MK_UNSUPPORTED; note_unsupported(xe + " for " + m.tname + " \"" + m.name + "\"");
                break;
            }
            for (int k = 0; k < 3; k++) r.u[k] = cv[1] * x.m[0][k] + cv[2] * x.m[1][k] + cv[3] * x.m[2][k];   // multv3
            double thick = cv[0];
            if ((-1e-6 <= thick) & (thick <= 1e-6)) thick = 0;
            r.pad2 = thick * x.sca;
            const std::string& fname = m.sargs[ht];
            auto it = bsdf_index.find(fname);
            if (it == bsdf_index.end()) {
                const std::string path = find_radiance_file(fname, sc.basedir);
                int bi = -1; std::string be;
                if (path.empty()) be = "cannot find BSDF file \"" + fname + "\"";
                else load_klems_bsdf(path, fs, bi, be);
                if (bi < 0) { err = be + " (" + m.tname + " \"" + m.name + "\")"; load_failed = true; }
                it = bsdf_index.emplace(fname, bi).first;
            }
            if (it->second < 0) { r.kind = MK_UNSUPPORTED; break; }
            r.pad[0] = it->second;
            break;
        }
        case OT_DIELECTRIC: r.kind = MK_DIELECTRIC; need(5); break;      // dielectric.c (built without DISPERSE)
        case OT_INTERFACE: r.kind = MK_INTERFACE; need(8); break;
        case OT_LIGHT: r.kind = MK_LIGHT; need(3); break;
        case OT_GLOW:  r.kind = MK_GLOW; need(4); break;
        case OT_ILLUM: r.kind = MK_ILLUM; need(3); break;
        case OT_SPOTLIGHT: r.kind = MK_SPOT; need(7); break;
        default:
            r.kind = MK_UNSUPPORTED;
            note_unsupported("unsupported material type " + m.tname + " \"" + m.name + "\"");
        }
        // patterns under the material (raytexture(r, m->omod)): the two sky brightness
        // functions are built as native code, anything else poisons the material
        r.pat = -1;
        int lastpat = -1;
        for (int q = m.omod, g = 0; q >= 0 && g < 10000; q = sc.objs[q].omod, g++) {
            const Object& po = sc.objs[q];
            int t = po.otype;
            if (t == OT_ALIAS && po.sargs.empty()) continue;
            PatRec pr; memset(&pr, 0, sizeof(pr));
            pr.next = -1;
            if (po.tname == "brightfunc" && po.sargs.size() >= 2) {
                if (po.sargs[1] == "skybright.cal" && po.sargs[0] == "skybr" && po.fargs.size() >= 7) pr.kind = PAT_SKYBRIGHT;
                else if (po.sargs[1] == "perezlum.cal" && po.sargs[0] == "skybright" && po.fargs.size() >= 10) pr.kind = PAT_PEREZLUM;
            }
            Xf x; std::string xe;
            if (pr.kind && !parse_xf(po.sargs, 2, x, xe)) pr.kind = 0;
            if (pr.kind) {
                for (size_t k = 0; k < 10 && k < po.fargs.size(); k++) pr.a[k] = po.fargs[k];
                // backward transform = inverse of the forward 3x3, divided by the backward scale (= 1 / forward scale)
                const double (*f)[4] = x.m;
                double det = f[0][0] * (f[1][1] * f[2][2] - f[1][2] * f[2][1]) - f[0][1] * (f[1][0] * f[2][2] - f[1][2] * f[2][0]) +
                             f[0][2] * (f[1][0] * f[2][1] - f[1][1] * f[2][0]);
                if (det == 0.0) pr.kind = 0;
                else {
                    double inv[3][3];
                    inv[0][0] = (f[1][1] * f[2][2] - f[1][2] * f[2][1]) / det; inv[0][1] = (f[0][2] * f[2][1] - f[0][1] * f[2][2]) / det; inv[0][2] = (f[0][1] * f[1][2] - f[0][2] * f[1][1]) / det;
                    inv[1][0] = (f[1][2] * f[2][0] - f[1][0] * f[2][2]) / det; inv[1][1] = (f[0][0] * f[2][2] - f[0][2] * f[2][0]) / det; inv[1][2] = (f[0][2] * f[1][0] - f[0][0] * f[1][2]) / det;
                    inv[2][0] = (f[1][0] * f[2][1] - f[1][1] * f[2][0]) / det; inv[2][1] = (f[0][1] * f[2][0] - f[0][0] * f[2][1]) / det; inv[2][2] = (f[0][0] * f[1][1] - f[0][1] * f[1][0]) / det;
                    const double bsca = 1.0 / fabs(x.sca);
                    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) pr.xb[a * 3 + b] = inv[a][b] / bsca;
                }
            }
            if (!pr.kind) {
                r.flags |= 1;
                note_unsupported("unsupported modifier type " + po.tname + " \"" + po.name +
                                 "\" (under " + m.tname + " \"" + m.name + "\")");
                continue;
            }
            int pi = (int)fs.pats.size();
            fs.pats.push_back(pr);
            if (lastpat < 0) r.pat = pi; else fs.pats[lastpat].next = pi;
            lastpat = pi;
        }
        int slot = (int)fs.mats.size();
        matslot[mi] = slot;
        fs.mats.push_back(r);
        if (m.otype == OT_ILLUM && !m.sargs.empty() && m.sargs[0] != "void") {
            int alt = sc.lastmod(mi, m.sargs[0]);
            int am = alt >= 0 ? sc.findmaterial(alt) : -1;
            int as = slot_of(am);
            fs.mats[slot].alt = as;
            if (alt >= 0 && am != alt) {
                fs.mats[slot].flags |= 1;
                note_unsupported("illum \"" + m.name + "\" alternate is not a plain material");
            }
        }
        return slot;
    };

    for (int i = 0; i < n; i++) {
        const Object& o = sc.objs[i];
        int32_t* hdr = &fs.objhdr[(size_t)i * 4];
        hdr[0] = PK_NONE; hdr[1] = o.omod; hdr[2] = -1; hdr[3] = 0;
        if (ot_is_volume(o.otype) && o.expanded) continue;
        if (ot_is_volume(o.otype)) {
            hdr[0] = PK_UNSUPPORTED;
            hdr[3] = (int32_t)geom_alloc(fs, 0);
            fs.nsurf_unsupported++;
            note_unsupported("unsupported object type " + o.tname + " \"" + o.name + "\"");
            continue;
        }
        if (!ot_is_surface(o.otype)) continue;
        std::string warn;
        switch (o.otype) {
        case OT_POLYGON: flatten_face(o, fs, hdr, warn); break;
        case OT_SPHERE: case OT_BUBBLE: flatten_sphere(o, fs, hdr, warn); break;
        case OT_CONE: case OT_CUP: case OT_CYLINDER: case OT_TUBE: case OT_RING:
            flatten_cone(o, fs, hdr, warn); break;
        case OT_SOURCE: hdr[0] = PK_NONE; break;       // never in the octree
        }
        if (hdr[3] == 0) hdr[3] = (int32_t)geom_alloc(fs, 0);   // header-only record
        if (!warn.empty()) fs.warnings.push_back(warn);
        if ((hdr[0] & 0xff) == PK_UNSUPPORTED) { fs.nsurf_unsupported++; note_unsupported(warn); }
        int flags = 0;
        if (o.omod >= 0 && sc.objs[o.omod].name == "Phong") hdr[0] |= PX_PHONG;
        if (o.volume_obj) hdr[0] |= PX_NOTFLAT;
        if (o.omod >= 0) {
            int mi = sc.findmaterial(i);
            if (mi >= 0) {
                flags |= PF_HASMAT;
                if (mat_is_transp(sc.objs[mi])) flags |= PF_TRANSP;
                hdr[2] = slot_of(mi);
                // the chain between the surface and its material must be plain
                // (aliases only); patterns in between poison the material use
                for (int q = o.omod, g = 0; q >= 0 && q != mi && g < 10000; g++) {
                    int t = sc.objs[q].otype;
                    if (t != OT_ALIAS) {
                        fs.mats[hdr[2]].flags |= 1;
                        note_unsupported("unsupported modifier type " + sc.objs[q].tname + " \"" +
                                         sc.objs[q].name + "\"");
                        q = sc.objs[q].omod;
                    } else if (sc.objs[q].sargs.empty()) q = sc.objs[q].omod;
                    else break;   // alias jumps: resolved by findmaterial
                }
            }
        }
        hdr[0] |= flags << 8;
    }
    // header copies in front of the geometry records + (id, record) leaf entries
    for (int i = 0; i < n; i++) {
        const int32_t* hdr = &fs.objhdr[(size_t)i * 4];
        if (hdr[3] >= 2) memcpy(&fs.geom[hdr[3] - 2], hdr, 16);
    }
    if (fs.geom.size() & 1) fs.geom.push_back(0.0);
    {
        std::vector<int> newoff(sc.leafpool.size(), -1);
        fs.leaf2.clear();
        for (size_t p = 0; p < sc.leafpool.size();) {
            int cnt = sc.leafpool[p];
            newoff[p] = (int)(fs.leaf2.size() / 2);
            fs.leaf2.push_back(cnt); fs.leaf2.push_back(0);
            for (int k = 1; k <= cnt; k++) {
                int id = sc.leafpool[p + k];
                if (id < 0 || id >= n) { err = "octree refers to object " + std::to_string(id) + " outside the scene"; return false; }
                // entry = object id | hot bits << 25 (kind, projection axis, rectangle flag), record offset:
                // what the walker needs to pick the test without touching the object header
                const int h0 = fs.objhdr[(size_t)id * 4];
                const int hot = (h0 & 0xf) | (((h0 >> 10) & 3) << 4) | (((h0 >> 12) & 1) << 6);
                fs.leaf2.push_back(id | (hot << 25));
                fs.leaf2.push_back(fs.objhdr[(size_t)id * 4 + 3]);
            }
            p += cnt + 1;
        }
        // the walker knows how many surfaces wait in a leaf (and whether any needs the slow sphere / cone
        // test) without a dependent read of the set's count word
        if (n >= (1 << 25)) { err = "scene has too many objects for this engine (2^25)"; return false; }
        if (fs.leaf2.size() / 2 >= ((size_t)1 << 26)) { err = "octree has too many leaf-set entries for this engine"; return false; }
        // full-leaf word = -(set offset << 4 | has-curved-surface << 3 | min(count, 7)) - 2
        auto leafword = [&](int w) {
            const size_t p = (size_t)(-w - 2);
            const int cnt = sc.leafpool[p];
            int curved = 0;
            for (int k = 1; k <= cnt; k++) {
                const int kind = fs.objhdr[(size_t)sc.leafpool[p + k] * 4] & 0xff;
                if (kind != PK_FACE && kind != PK_NONE) curved = 1;
            }
            return -((newoff[p] << 4) | (curved << 3) | std::min(cnt, 7)) - 2;
        };
        for (auto& w : fs.nodes) if (w < -1) w = leafword(w);
        fs.root = sc.root < -1 ? leafword(sc.root) : sc.root;
        for (int k = 0; k < 8; k++) fs.geom.push_back(0.0);       // records are read 80 bytes at a time
    }

    // ---- sources: rt/source.c:46-142 marksources(), srcsupp.c:155-179 ----
    for (int i = 0; i < n; i++) {
        const Object& o = sc.objs[i];
        if (!ot_is_surface(o.otype) || o.omod < 0) continue;
        int mi = sc.findmaterial(i);
        if (mi < 0) continue;
        const Object& m = sc.objs[mi];
        if (m.otype == OT_ANTIMATTER) { note_unsupported("unsupported material type antimatter \"" + m.name + "\""); continue; }
        if (!ot_is_light(m.otype)) continue;
        size_t want = m.otype == OT_GLOW ? 4 : m.otype == OT_SPOTLIGHT ? 7 : 3;
        if (m.fargs.size() != want) { err = "bad # arguments for " + m.tname + " \"" + m.name + "\""; return false; }
        if (m.fargs[0] <= FTINY && m.fargs[1] <= FTINY && m.fargs[2] <= FTINY) continue;
        if (m.otype == OT_GLOW && o.otype != OT_SOURCE && m.fargs[3] <= FTINY) continue;
        SrcRec s; memset(&s, 0, sizeof(s));
        s.so = i; s.mat = slot_of(mi);
        s.val[0] = (float)m.fargs[0]; s.val[1] = (float)m.fargs[1]; s.val[2] = (float)m.fargs[2];
        if (o.otype == OT_SOURCE) {          // ssetsrc()
            if (o.fargs.size() != 4) { err = "bad arguments for source \"" + o.name + "\""; return false; }
            s.flags |= SF_DISTANT | SF_CIRC;
            s.sloc[0] = o.fargs[0]; s.sloc[1] = o.fargs[1]; s.sloc[2] = o.fargs[2];
            if (vnormalize(s.sloc) == 0.0) { err = "zero direction for source \"" + o.name + "\""; return false; }
            double theta = PI / 180.0 / 2.0 * o.fargs[3];
            if (theta <= FTINY) { err = "zero size for source \"" + o.name + "\""; return false; }
            s.ss2 = 2.0 * PI * (1.0 - cos(theta));
            s.srad = sqrt(s.ss2 / PI);
            // setflatss() with deterministic perpendicular (rand_samp differs
            // only in the orientation of the jitter frame)
            double snorm[3] = {s.sloc[0], s.sloc[1], s.sloc[2]};
            getperp(s.ss[0], snorm);
            double mult = .5 * sqrt(s.ss2);
            for (int k = 0; k < 3; k++) s.ss[0][k] *= mult;
            vcross(s.ss[1], snorm, s.ss[0]);
            s.ss[2][0] = s.ss[2][1] = s.ss[2][2] = 0.0;
        } else {
            // local emitters: srcsupp.c fsetsrc / sphsetsrc / rsetsrc / cylsetsrc.  The
            // perpendiculars are the deterministic ones (-u-); with -u+ the reference
            // only randomises the orientation of the partition / jitter frame.
            const int32_t* hdr = &fs.objhdr[(size_t)i * 4];
            const int kind = hdr[0] & 0xff;
            const double* g = &fs.geom[hdr[3]];
            auto bad = [&](const char* what) { err = std::string(what) + " \"" + o.name + "\""; return false; };
            if (kind == PK_FACE) {                                   // fsetsrc(), srcsupp.c:91-152
                int nv = (hdr[0] >> 16) & 0xffff;
                const double* va = o.fargs.data();
                double norm[3] = {g[0], g[1], g[2]};
                double area = 0.0;
                {   // getface(): area = |sum of fan cross products| / 2 (face.c:66-82)
                    double nsum[3] = {0, 0, 0}, v1[3], v2[3], v3[3];
                    for (int k = 0; k < 3; k++) v1[k] = va[3 + k] - va[k];
                    for (int q = 2; q < nv; q++) {
                        for (int k = 0; k < 3; k++) v2[k] = va[3 * q + k] - va[k];
                        vcross(v3, v1, v2);
                        for (int k = 0; k < 3; k++) { nsum[k] += v3[k]; v1[k] = v2[k]; }
                    }
                    area = 0.5 * vnormalize(nsum);
                }
                if (area == 0.0) return bad("zero source area for");
                for (int j = 0; j < 3; j++) {
                    s.sloc[j] = 0.0;
                    for (int q = 0; q < nv; q++) s.sloc[j] += va[3 * q + j];
                    s.sloc[j] /= (double)nv;
                }
                if (!inface_host(s.sloc, hdr[0], g, nv)) return bad("cannot hit source center of");
                s.flags |= SF_FLAT;
                for (int j = 0; j < 3; j++) s.ss[2][j] = norm[j];
                s.ss2 = area;
                double r2 = 0.0;
                for (int q = 0; q < nv; q++) {
                    double d = 0; for (int j = 0; j < 3; j++) d += (va[3 * q + j] - s.sloc[j]) * (va[3 * q + j] - s.sloc[j]);
                    if (d > r2) r2 = d;
                }
                s.srad = sqrt(r2);
                if (nv == 4) {                                       // parallelogram case
                    for (int j = 0; j < 3; j++) { s.ss[0][j] = .5 * (va[3 + j] - va[j]); s.ss[1][j] = .5 * (va[9 + j] - va[j]); }
                } else if (nv == 3) {                                // triangle case
                    auto d2line = [&](const double* pp, const double* e1, const double* e2) {
                        double d = 0, d1 = 0, dd = 0;
                        for (int k = 0; k < 3; k++) { d += (e1[k] - e2[k]) * (e1[k] - e2[k]); d1 += (e1[k] - pp[k]) * (e1[k] - pp[k]); dd += (e2[k] - pp[k]) * (e2[k] - pp[k]); }
                        double d2 = d + d1 - dd;
                        return d1 - 0.25 * d2 * d2 / d;
                    };
                    int near0 = 2;
                    double dmin = d2line(s.sloc, va + 6, va);
                    for (int q = 0; q < 2; q++) {
                        double d2 = d2line(s.sloc, va + 3 * q, va + 3 * (q + 1));
                        if (d2 >= dmin) continue;
                        near0 = q; dmin = d2;
                    }
                    int i2 = (near0 + 1) % 3;
                    for (int j = 0; j < 3; j++) s.ss[0][j] = va[3 * i2 + j] - va[3 * near0 + j];
                    vnormalize(s.ss[0]);
                    dmin = sqrt(dmin);
                    for (int j = 0; j < 3; j++) s.ss[0][j] *= dmin;
                    vcross(s.ss[1], norm, s.ss[0]);
                } else {                                             // setflatss(): hope for convex
                    getperp(s.ss[0], s.ss[2]);
                    double mult = .5 * sqrt(s.ss2);
                    for (int k = 0; k < 3; k++) s.ss[0][k] *= mult;
                    vcross(s.ss[1], s.ss[2], s.ss[0]);
                }
            } else if (kind == PK_SPHERE) {                          // sphsetsrc(), srcsupp.c:182-203
                if (o.fargs[3] <= FTINY) return bad("illegal source radius for");
                s.flags |= SF_CIRC;
                for (int k = 0; k < 3; k++) s.sloc[k] = o.fargs[k];
                s.srad = (float)o.fargs[3];
                s.ss2 = PI * s.srad * s.srad;
                for (int k = 0; k < 3; k++) s.ss[k][k] = 0.7236 * o.fargs[3];
            } else if (kind == PK_RING) {                            // rsetsrc(), srcsupp.c:206-232
                if (g[9] <= FTINY) return bad("illegal source radius for");
                if (g[8] > 0.0) return bad("cannot hit source center of");
                for (int k = 0; k < 3; k++) { s.sloc[k] = g[4 + k]; s.ss[2][k] = g[k]; }
                s.flags |= SF_FLAT | SF_CIRC;
                s.srad = (float)g[9];
                s.ss2 = PI * s.srad * s.srad;
                getperp(s.ss[0], s.ss[2]);
                double mult = .5 * sqrt((double)(float)s.ss2);
                for (int k = 0; k < 3; k++) s.ss[0][k] *= mult;
                vcross(s.ss[1], s.ss[2], s.ss[0]);
            } else if (kind == PK_CYL) {                             // cylsetsrc(), srcsupp.c:235-268
                const double al = g[3], r0 = g[8];
                const double ad[3] = {g[0], g[1], g[2]};
                if (r0 <= FTINY) return bad("illegal source radius for");
                s.flags |= SF_CYL;
                for (int k = 0; k < 3; k++) s.sloc[k] = .5 * (o.fargs[3 + k] + o.fargs[k]);
                s.srad = (float)(.5 * al);
                s.ss2 = 2. * r0 * al;
                for (int k = 0; k < 3; k++) s.ss[0][k] = .5 * al * ad[k];
                getperp(s.ss[2], ad);
                for (int k = 0; k < 3; k++) s.ss[2][k] *= .8559 * r0;
                vcross(s.ss[1], s.ss[2], ad);
            } else
                return bad("illegal material (this surface type cannot be a light source) on");
            s.srad = (double)(float)s.srad;
        }
        s.ss2 = (double)(float)s.ss2;            // SRCREC.ss2 / .srad are floats (source.h:60-61)
        if (m.otype == OT_GLOW) {
            s.flags |= SF_PROX;
            s.prox = (double)(float)m.fargs[3];
            if (s.flags & SF_DISTANT) s.flags |= SF_SKIP;
        } else if (m.otype == OT_SPOTLIGHT) {    // makespot(), srcsupp.c:271-291
            if (m.fargs[3] <= FTINY) { err = "zero angle for spotlight \"" + m.name + "\""; return false; }
            s.flags |= SF_SPOT;
            s.spot_siz = (float)(2.0 * PI * (1.0 - cos(PI / 180.0 / 2.0 * m.fargs[3])));
            for (int k = 0; k < 3; k++) s.spot_aim[k] = m.fargs[4 + k];
            s.spot_flen = (float)vnormalize(s.spot_aim);
            if (s.spot_flen == 0.0f) { err = "zero focus vector for spotlight \"" + m.name + "\""; return false; }
            if (s.flags & SF_FLAT) {             // checkspot(), srcsupp.c:443-459
                double d = vdot(s.spot_aim, s.ss[2]);
                if (!(d > FTINY)) {
                    double d1 = 1. - s.spot_siz / (2. * PI);
                    if (!(1. - FTINY - d * d < d1 * d1)) {
                        s.flags |= SF_SKIP;
                        fs.warnings.push_back("invalid spotlight direction for \"" + o.name + "\"");
                    }
                }
            }
        }
        fs.srcs.push_back(s);
    }
    if (load_failed) return false;               // a BSDF file that a used material names could not be loaded (err says which)
    return true;
}

}  // namespace rb
