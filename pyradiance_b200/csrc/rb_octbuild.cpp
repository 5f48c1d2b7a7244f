// rb_octbuild.cpp -- own octree builder: text scene -> frozen Radiance .oct
// (built early because the synthetic benchmark scenes must be
// producible without the reference's oconv).
//
// File format written here is the one readoct()/readscene() accept
// (src/radiance/ot/writeoct.c:28-69,117-133; common/sceneio.c:112-159;
// common/portio.c:20-69).  Subdivision rule follows ot/oconv.c:215-320:
// a cube is split while it holds more than `objlim` surfaces and its children
// would not be smaller than cusize/resolution; identical sibling leaves are
// merged like common/octree.c:73-91 combine().  The cube/surface overlap test
// is a conservative one of our own (bounding box + plane/sphere distance), so
// a leaf may list a surface the reference's exact test would skip -- the tree
// is valid for any Radiance reader, results of ray queries do not change.
#include "../../include/rb200.h"
#include "rb_scene.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace rb {
namespace {

const double FTINY = 1e-6;
const int MAXSET = 8191;

struct Prim {
    int obj;
    int kind;                 // 0 polygon, 1 sphere, 2 other (bbox only)
    double lo[3], hi[3];
    double n[3], off;         // polygon plane
    double c[3], r;           // sphere
};

struct Builder {
    std::vector<Prim> prims;
    std::vector<int> nodes;
    std::vector<std::vector<int>> sets;      // leaf sets by id
    std::map<std::vector<int>, int> setid;   // identical sets share an id (fullnode())
    int objlim = 6;
    double mincusize = 0;
    std::string err;

    bool overlaps(const Prim& p, const double org[3], double size) const {
        double lo[3], hi[3];
        for (int k = 0; k < 3; k++) { lo[k] = org[k] - FTINY; hi[k] = org[k] + size + FTINY; }
        for (int k = 0; k < 3; k++) if (p.hi[k] < lo[k] || p.lo[k] > hi[k]) return false;
        if (p.kind == 0) {      // plane must cut the cube (ot/o_face.c:86-99)
            double v1[3], v2[3];
            for (int j = 0; j < 3; j++) {
                if (p.n[j] > 0.0) { v1[j] = lo[j]; v2[j] = hi[j]; } else { v1[j] = hi[j]; v2[j] = lo[j]; }
            }
            double d1 = v1[0] * p.n[0] + v1[1] * p.n[1] + v1[2] * p.n[2] - p.off;
            double d2 = v2[0] * p.n[0] + v2[1] * p.n[1] + v2[2] * p.n[2] - p.off;
            if (d1 > FTINY || d2 < -FTINY) return false;
        } else if (p.kind == 1) {
            double dmin = 0, dmax = 0;
            for (int k = 0; k < 3; k++) {
                double a = p.c[k] - lo[k], b = p.c[k] - hi[k];
                if (a < 0) dmin += a * a; else if (b > 0) dmin += b * b;
                double m = std::max(fabs(a), fabs(b));
                dmax += m * m;
            }
            if (dmin > (p.r + FTINY) * (p.r + FTINY)) return false;     // cube outside
            if (dmax < (p.r - FTINY) * (p.r - FTINY)) return false;     // cube strictly inside
        }
        return true;
    }
    int leaf(const std::vector<int>& ids) {
        std::vector<int> s;
        for (int i : ids) s.push_back(prims[i].obj);
        std::sort(s.begin(), s.end());
        auto it = setid.find(s);
        if (it != setid.end()) return -it->second - 2;
        int id = (int)sets.size();
        sets.push_back(s);
        setid[s] = id;
        return -id - 2;
    }
    // returns tree word: >=0 node, -1 empty, <=-2 leaf id
    int build(const std::vector<int>& cand, const double org[3], double size, int depth) {
        std::vector<int> in;
        for (int i : cand) if (overlaps(prims[i], org, size)) in.push_back(i);
        if (in.empty()) return -1;
        double half = size * 0.5;
        bool toosmall = half < ((int)in.size() < MAXSET ? mincusize : mincusize / 256.0);
        if ((int)in.size() <= objlim || toosmall || depth >= 20) {
            if ((int)in.size() > MAXSET) { err = "set overflow in octree build"; return -1; }
            return leaf(in);
        }
        int kids[8];
        for (int i = 0; i < 8; i++) {
            double ko[3];
            for (int j = 0; j < 3; j++) ko[j] = org[j] + (((1 << j) & i) ? half : 0.0);
            kids[i] = build(in, ko, half, depth + 1);
            if (!err.empty()) return -1;
        }
        bool same = kids[0] < 0;
        for (int i = 1; i < 8 && same; i++) same = (kids[i] == kids[0]);
        if (same) return kids[0];                 // combine()
        int idx = (int)(nodes.size() / 8);
        nodes.insert(nodes.end(), kids, kids + 8);
        return idx;
    }
};

struct Out {
    std::vector<unsigned char> b;
    void putc_(int c) { b.push_back((unsigned char)c); }
    void putstr(const std::string& s) { for (char c : s) putc_(c); putc_(0); }
    void putint(long i, int siz) {                       // portio.c:34-47
        int bits = siz << 3;
        while ((bits -= 8) > 0) putc_((int)((i >> bits) & 0xff));
        putc_((int)(i & 0xff));
    }
    void putflt(double f) {                              // portio.c:50-69
        int e;
        long m = (long)(frexp(f, &e) * 0x7fffffff);
        if (e > 127) { m = m > 0 ? (long)0x7fffffff : -(long)0x7fffffff; e = 127; }
        else if (e < -128) { m = 0; e = 0; }
        putint(m, 4);
        putint(e, 1);
    }
};

// the reference's type table order (common/otypes.h:127-186); frozen scenes
// carry their own table, so only the names matter
const char* kTypeNames[] = {
    "polygon", "cone", "sphere", "texfunc", "ring", "cylinder", "instance", "cup", "bubble", "tube", "mesh",
    "alias", "plastic", "metal", "glass", "trans", "dielectric", "plastic2", "metal2", "trans2", "interface",
    "plasfunc", "metfunc", "brightfunc", "brightdata", "brighttext", "colorpict", "glow", "source", "light",
    "illum", "spotlight", "mist", "mirror", "transfunc", "BRTDfunc", "BSDF", "aBSDF", "WGMDfunc", "plasdata",
    "metdata", "transdata", "colorfunc", "antimatter", "colordata", "colortext", "texdata", "mixfunc",
    "mixdata", "mixtext", "mixpict", "prism1", "prism2", "ashik2", "spectrum", "specfile", "specfunc",
    "specdata", "specpict"};

void puttree(Out& o, const Builder& B, int w) {
    if (w >= 0) {
        o.putc_(2);
        for (int i = 0; i < 8; i++) puttree(o, B, B.nodes[(size_t)w * 8 + i]);
    } else if (w == -1) o.putc_(0);
    else {
        const std::vector<int>& s = B.sets[-w - 2];
        o.putc_(1);
        o.putint((long)s.size(), 4);
        for (int id : s) o.putint(id, 4);
    }
}

}  // namespace

// Collect the surfaces of `sc` and build the tree.  With keep_cube the scene's
// own cube is kept (re-build after instance / mesh expansion), else the cube of
// ot/oconv.c:122-139 is computed.  Result in B; cube in cuorg/cusize.
static bool build_tree(const Scene& sc, Builder& B, int objlim, int maxres, bool keep_cube, double cuorg[3],
                       double& cusize, char sbuf[4][64], int& root, std::string& err) {
    B.objlim = objlim > 0 ? objlim : 6;
    double bbmin[3] = {1e10, 1e10, 1e10}, bbmax[3] = {-1e10, -1e10, -1e10};
    for (int i = 0; i < (int)sc.objs.size(); i++) {
        const Object& o = sc.objs[i];
        if (!ot_is_surface(o.otype) || o.otype == OT_SOURCE) continue;
        Prim p; memset(&p, 0, sizeof(p));
        p.obj = i; p.kind = 2;
        for (int k = 0; k < 3; k++) { p.lo[k] = 1e10; p.hi[k] = -1e10; }
        const std::vector<double>& a = o.fargs;
        auto grow = [&](double x, double y, double z) {
            double v[3] = {x, y, z};
            for (int k = 0; k < 3; k++) { p.lo[k] = std::min(p.lo[k], v[k]); p.hi[k] = std::max(p.hi[k], v[k]); }
        };
        if (o.otype == OT_POLYGON) {
            if (a.size() < 9 || a.size() % 3) { err = "bad polygon \"" + o.name + "\""; return false; }
            int nv = (int)a.size() / 3;
            for (int v = 0; v < nv; v++) grow(a[3 * v], a[3 * v + 1], a[3 * v + 2]);
            double n[3] = {0, 0, 0}, v1[3], v2[3];
            for (int k = 0; k < 3; k++) v1[k] = a[3 + k] - a[k];
            for (int v = 2; v < nv; v++) {
                for (int k = 0; k < 3; k++) v2[k] = a[3 * v + k] - a[k];
                n[0] += v1[1] * v2[2] - v1[2] * v2[1]; n[1] += v1[2] * v2[0] - v1[0] * v2[2]; n[2] += v1[0] * v2[1] - v1[1] * v2[0];
                for (int k = 0; k < 3; k++) v1[k] = v2[k];
            }
            double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (len == 0.0) continue;            // zero area: never in the tree (ot/o_face.c:54-55)
            double off = 0;
            for (int k = 0; k < 3; k++) n[k] /= len;
            for (int v = 0; v < nv; v++) off += n[0] * a[3 * v] + n[1] * a[3 * v + 1] + n[2] * a[3 * v + 2];
            off /= nv;
            p.kind = 0; p.off = off;
            for (int k = 0; k < 3; k++) p.n[k] = n[k];
        } else if (o.otype == OT_SPHERE || o.otype == OT_BUBBLE) {
            if (a.size() != 4) { err = "bad sphere \"" + o.name + "\""; return false; }
            double r = fabs(a[3]);
            grow(a[0] - r, a[1] - r, a[2] - r); grow(a[0] + r, a[1] + r, a[2] + r);
            p.kind = 1; p.r = r;
            for (int k = 0; k < 3; k++) p.c[k] = a[k];
        } else if (o.otype == OT_RING) {
            if (a.size() != 8) { err = "bad ring \"" + o.name + "\""; return false; }
            double r = std::max(fabs(a[6]), fabs(a[7]));
            grow(a[0] - r, a[1] - r, a[2] - r); grow(a[0] + r, a[1] + r, a[2] + r);
        } else {   // cone, cup, cylinder, tube
            size_t need = (o.otype == OT_CYLINDER || o.otype == OT_TUBE) ? 7 : 8;
            if (a.size() != need) { err = "bad arguments for \"" + o.name + "\""; return false; }
            double r0 = fabs(a[6]), r1 = need == 8 ? fabs(a[7]) : r0;
            grow(a[0] - r0, a[1] - r0, a[2] - r0); grow(a[0] + r0, a[1] + r0, a[2] + r0);
            grow(a[3] - r1, a[4] - r1, a[5] - r1); grow(a[3] + r1, a[4] + r1, a[5] + r1);
        }
        for (int k = 0; k < 3; k++) { bbmin[k] = std::min(bbmin[k], p.lo[k]); bbmax[k] = std::max(bbmax[k], p.hi[k]); }
        B.prims.push_back(p);
    }
    const double OMARGIN = 10 * FTINY;
    if (keep_cube) {
        for (int k = 0; k < 3; k++) cuorg[k] = sc.cuorg[k];
        cusize = sc.cusize;
        for (int k = 0; k < 3 && !B.prims.empty(); k++)
            if (bbmin[k] < cuorg[k] - OMARGIN || bbmax[k] > cuorg[k] + cusize + OMARGIN) {
                err = "boundary does not encompass scene (instance or mesh sticks out of the parent octree's cube)";
                return false;
            }
    } else {
        // ot/oconv.c:122-139: cube centred on the bounding box, with margin
        cuorg[0] = cuorg[1] = cuorg[2] = 0; cusize = 0;
        if (!B.prims.empty()) {
            for (int k = 0; k < 3; k++) { bbmin[k] -= OMARGIN; bbmax[k] += OMARGIN; }
            for (int k = 0; k < 3; k++) cusize = std::max(cusize, bbmax[k] - bbmin[k]);
            for (int k = 0; k < 3; k++) cuorg[k] = (bbmax[k] + bbmin[k] - cusize) * .5;
        }
        // the reader parses the "%.12g" strings: build with exactly those values
        for (int k = 0; k < 3; k++) { snprintf(sbuf[k], 64, "%.12g", cuorg[k]); cuorg[k] = atof(sbuf[k]); }
        snprintf(sbuf[3], 64, "%.12g", cusize); cusize = atof(sbuf[3]);
    }
    B.mincusize = cusize / (maxres > 0 ? maxres : 16384) - FTINY;
    std::vector<int> all(B.prims.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = (int)i;
    root = B.build(all, cuorg, cusize, 0);
    if (!B.err.empty()) { err = B.err; return false; }
    return true;
}

// Re-build sc's octree in memory over its current surface list (used after
// instances / meshes have been expanded into world-space surfaces).
bool rebuild_octree(Scene& sc, int objlim, int maxres, std::string& err) {
    Builder B;
    double cuorg[3], cusize; char sbuf[4][64]; int root;
    if (!build_tree(sc, B, objlim, maxres, true, cuorg, cusize, sbuf, root, err)) return false;
    sc.nodes = B.nodes;
    sc.leafpool.clear();
    std::vector<int> off(B.sets.size());
    for (size_t i = 0; i < B.sets.size(); i++) {
        off[i] = (int)sc.leafpool.size();
        sc.leafpool.push_back((int)B.sets[i].size());
        for (int id : B.sets[i]) sc.leafpool.push_back(id);
    }
    auto remap = [&](int w) { return w < -1 ? -off[-w - 2] - 2 : w; };
    for (auto& w : sc.nodes) w = remap(w);
    sc.root = remap(root);
    // depth of the new tree
    sc.maxdepth = 0;
    std::vector<std::pair<int, int>> st;
    if (sc.root >= 0) st.push_back({sc.root, 1});
    while (!st.empty()) {
        auto [nd, d] = st.back(); st.pop_back();
        sc.maxdepth = std::max(sc.maxdepth, d);
        for (int k = 0; k < 8; k++) { int w = sc.nodes[(size_t)nd * 8 + k]; if (w >= 0) st.push_back({w, d + 1}); }
    }
    return true;
}

bool build_octree_file(const Scene& sc, const std::string& cmdline, const std::string& oct_path, int objlim,
                       int maxres, std::string& err) {
    Builder B;
    double cuorg[3], cusize; char sbuf[4][64]; int root;
    if (!build_tree(sc, B, objlim, maxres, false, cuorg, cusize, sbuf, root, err)) return false;

    Out o;
    std::string hdr = "#?RADIANCE\n" + cmdline + "\nFORMAT=Radiance_octree\n\n";
    for (char c : hdr) o.putc_(c);
    o.putint(4 * 8 + 251 + 4, 2);
    for (int k = 0; k < 4; k++) o.putstr(sbuf[k]);
    o.putstr("");
    o.putint((long)sc.objs.size(), 4);
    puttree(o, B, root);
    // frozen scene (sceneio.c:112-159)
    std::map<std::string, int> tindex;
    int nt = (int)(sizeof(kTypeNames) / sizeof(kTypeNames[0]));
    for (int i = 0; i < nt; i++) { o.putstr(kTypeNames[i]); tindex[kTypeNames[i]] = i; }
    o.putstr("");
    for (const Object& ob : sc.objs) {
        auto it = tindex.find(ob.tname);
        if (it == tindex.end()) { err = "unknown object type \"" + ob.tname + "\""; return false; }
        o.putint(it->second, 1);
        o.putint(ob.omod, 4);
        o.putstr(ob.name);
        o.putint((long)ob.sargs.size(), 2);
        for (const auto& s : ob.sargs) o.putstr(s);
        o.putint((long)ob.fargs.size(), 2);
        for (double f : ob.fargs) o.putflt(f);
    }
    o.putint(-1, 1);
    FILE* fp = fopen(oct_path.c_str(), "wb");
    if (!fp) { err = "cannot write \"" + oct_path + "\""; return false; }
    size_t nw = fwrite(o.b.data(), 1, o.b.size(), fp);
    fclose(fp);
    if (nw != o.b.size()) { err = "write error on \"" + oct_path + "\""; return false; }
    return true;
}

}  // namespace rb

// oconv -f [-i octree] file ...: text scenes (and optionally the objects of an existing octree,
// instances and meshes already expanded by the loader) into ONE frozen octree (ot/oconv.c:215-320).
extern "C" int rb_oconv_files(const char* const* rad_paths, int npaths, const char* include_octree, const char* oct_path,
                              int objlim, int maxres, char* errbuf, size_t errlen) {
    rb::Scene sc;
    std::string err, cmd = "rb_oconv -f";
    bool ok = true;
    if (include_octree && *include_octree) {
        ok = sc.load_octree(include_octree);
        if (!ok) err = sc.error;
        cmd += std::string(" -i ") + include_octree;
    }
    for (int k = 0; ok && k < npaths; k++) {
        ok = sc.read_rad_text(rad_paths[k]);
        if (!ok) err = sc.error;
        cmd += std::string(" ") + rad_paths[k];
    }
    for (size_t i = 0; ok && i < sc.objs.size(); i++)
        if (rb::ot_is_volume(sc.objs[i].otype) && !sc.objs[i].expanded) {
            ok = false;
            err = "rb_oconv: " + sc.objs[i].tname + " \"" + sc.objs[i].name +
                  "\" cannot be placed by this builder (use the reference oconv for scenes with instances / meshes)";
        }
    if (ok) ok = rb::build_octree_file(sc, cmd, oct_path, objlim, maxres, err);
    if (!ok && errbuf && errlen) { strncpy(errbuf, err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
    return ok ? 0 : -1;
}

extern "C" int rb_oconv(const char* rad_path, const char* oct_path, int objlim, int maxres, char* errbuf,
                        size_t errlen) {
    return rb_oconv_files(&rad_path, 1, nullptr, oct_path, objlim, maxres, errbuf, errlen);
}
