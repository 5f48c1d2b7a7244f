// rb_octbuild.cpp -- own octree builder: text scene -> frozen Radiance .oct
// (built early because the synthetic benchmark scenes must be
// producible without the reference's oconv).
//
// File format written here is the one readoct()/readscene() accept
// (src/radiance/ot/writeoct.c:28-69,117-133; common/sceneio.c:112-159;
// common/portio.c:20-69).  Subdivision rule follows ot/oconv.c:215-320:
// a cube is split while it holds more than `objlim` surfaces and its children
// would not be smaller than cusize/resolution; identical sibling leaves are
// merged like common/octree.c:73-91 combine().  The cube/surface overlap tests
// restate the reference's (ot/o_face.c, ot/sphere.c, ot/o_cone.c with
// common/plocate.c, clip.c) and the scene cube follows ot/bbox.c, so the file
// equals `oconv -f`'s byte for byte after the header's command line.
#include "../../include/rb200.h"
#include "rb_scene.hpp"
#include "rb_octtests.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace rb {
namespace {

using namespace octt;
}  // namespace
// rb_octbuild_gpu.cu
bool octbuild_device_available();
bool build_tree_device(const std::vector<octt::Prim>& prims, int objlim, double mincusize, const double cuorg[3], double cusize,
                       std::vector<int>& nodes, std::vector<int>& pool, int& root, std::string& err);
namespace {

// nodes and leaf sets of a (sub)tree; words: >= 0 node index, -1 empty, <= -2 leaf = -(offset into pool) - 2,
// where pool holds [count, object ids ascending ...] per leaf
struct SubTree {
    std::vector<int> nodes, pool;
    int leaf(const std::vector<int>& s) {
        const int off = (int)pool.size();
        pool.push_back((int)s.size());
        pool.insert(pool.end(), s.begin(), s.end());
        return -off - 2;
    }
    bool same_leaf(int a, int b) const {
        if (a == b) return true;
        const int *pa = &pool[-a - 2], *pb = &pool[-b - 2];
        return pa[0] == pb[0] && std::equal(pa + 1, pa + 1 + pa[0], pb + 1);
    }
    // append `src` (built independently) and return its root word in this tree's numbering
    int absorb(const SubTree& src, int root) {
        const int nbase = (int)(nodes.size() / 8), pbase = (int)pool.size();
        auto fix = [&](int w) { return w >= 0 ? w + nbase : w == -1 ? -1 : w - pbase; };
        pool.insert(pool.end(), src.pool.begin(), src.pool.end());
        nodes.reserve(nodes.size() + src.nodes.size());
        for (int w : src.nodes) nodes.push_back(fix(w));
        return fix(root);
    }
};

struct Builder {
    std::vector<Prim> prims;
    SubTree top;
    int objlim = 6;
    double mincusize = 0;
    int par_depth = 2;                       // children of cubes above this depth are built by their own threads
    std::string err;
    std::mutex errmu;

    // the reference's own cube tests (ot/o_face.c, ot/sphere.c, ot/o_cone.c), after a bounding-box reject
    // whose margin (4 FTINY) is wider than any of theirs
    bool overlaps(const Prim& p, const double org[3], double size) const { return octt::overlaps(p, org, size, mincusize); }
    bool failed() { std::lock_guard<std::mutex> g(errmu); return !err.empty(); }
    // returns the tree word of the cube; the order of nodes and set ids is that of a serial depth-first build
    int build(SubTree& T, const std::vector<int>& cand, const double org[3], double size, int depth) {
        std::vector<int> in;
        for (int i : cand) if (overlaps(prims[i], org, size)) in.push_back(i);
        if (in.empty()) return -1;
        const double half = size * 0.5;
        const bool toosmall = half < ((int)in.size() < MAXSET ? mincusize : mincusize / 256.0);
        if ((int)in.size() <= objlim || toosmall || depth >= 20) {
            if ((int)in.size() > MAXSET) {
                std::lock_guard<std::mutex> g(errmu);
                err = "set overflow in octree build";
                return -1;
            }
            std::vector<int> s;
            s.reserve(in.size());
            for (int i : in) s.push_back(prims[i].obj);
            std::sort(s.begin(), s.end());
            return T.leaf(s);
        }
        int kids[8];
        double ko[8][3];
        for (int i = 0; i < 8; i++)
            for (int j = 0; j < 3; j++) ko[i][j] = org[j] + (((1 << j) & i) ? half : 0.0);
        if (depth < par_depth && in.size() >= 4096) {
            SubTree sub[8];
            std::thread th[8];
            for (int i = 0; i < 8; i++)
                th[i] = std::thread([&, i] { kids[i] = build(sub[i], in, ko[i], half, depth + 1); });
            for (int i = 0; i < 8; i++) th[i].join();
            if (failed()) return -1;
            for (int i = 0; i < 8; i++) kids[i] = T.absorb(sub[i], kids[i]);
        } else {
            for (int i = 0; i < 8; i++) {
                kids[i] = build(T, in, ko[i], half, depth + 1);
                if (failed()) return -1;
            }
        }
        bool same = kids[0] < 0;                  // combine(): eight equal leaves (or eight empties) become one
        for (int i = 1; i < 8 && same; i++)
            same = kids[0] == -1 ? kids[i] == -1 : (kids[i] < -1 && T.same_leaf(kids[i], kids[0]));
        if (same) return kids[0];
        const int idx = (int)(T.nodes.size() / 8);
        T.nodes.insert(T.nodes.end(), kids, kids + 8);
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
        for (int i = 0; i < 8; i++) puttree(o, B, B.top.nodes[(size_t)w * 8 + i]);
    } else if (w == -1) o.putc_(0);
    else {
        const int* s = &B.top.pool[-w - 2];
        o.putc_(1);
        o.putint(s[0], 4);
        for (int k = 1; k <= s[0]; k++) o.putint(s[k], 4);
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
        auto sbox = [&](const double lo[3], const double hi[3]) {
            for (int k = 0; k < 3; k++) { bbmin[k] = std::min(bbmin[k], lo[k]); bbmax[k] = std::max(bbmax[k], hi[k]); }
        };
        auto grow = [&](double x, double y, double z) {
            double v[3] = {x, y, z};
            for (int k = 0; k < 3; k++) { p.lo[k] = std::min(p.lo[k], v[k]); p.hi[k] = std::max(p.hi[k], v[k]); }
        };
        if (o.otype == OT_POLYGON) {                     // common/face.c:35-106 getface()
            if (a.size() < 9 || a.size() % 3) { err = "bad polygon \"" + o.name + "\""; return false; }
            int nv = (int)a.size() / 3;
            for (int v = 0; v < nv; v++) grow(a[3 * v], a[3 * v + 1], a[3 * v + 2]);
            if (nv > 3 && dist2(&a[0], &a[3 * (nv - 1)]) <= FTINY * FTINY) nv--;      // closing vertex repeats the first
            double n[3] = {0, 0, 0}, v1[3], v2[3];
            for (int k = 0; k < 3; k++) v1[k] = a[3 + k] - a[k];
            for (int v = 2; v < nv; v++) {
                for (int k = 0; k < 3; k++) v2[k] = a[3 * v + k] - a[k];
                n[0] += v1[1] * v2[2] - v1[2] * v2[1]; n[1] += v1[2] * v2[0] - v1[0] * v2[2]; n[2] += v1[0] * v2[1] - v1[1] * v2[0];
                for (int k = 0; k < 3; k++) v1[k] = v2[k];
            }
            sbox(p.lo, p.hi);                        // ot/bbox.c:52-57: every vertex, zero-area faces too
            if (normalize3(n) == 0.0) continue;      // zero area: never in the tree (ot/o_face.c:54-55)
            double off = n[0] * a[0] + n[1] * a[1] + n[2] * a[2];
            for (int v = 1; v < nv; v++) off += n[0] * a[3 * v] + n[1] * a[3 * v + 1] + n[2] * a[3 * v + 2];
            off /= (double)nv;
            p.kind = 0; p.off = off; p.va = a.data(); p.nv = nv;
            for (int k = 0; k < 3; k++) p.n[k] = n[k];
            p.ax = fabs(n[0]) > fabs(n[1]) ? 0 : 1;
            if (fabs(n[2]) > fabs(n[p.ax])) p.ax = 2;
        } else if (o.otype == OT_SPHERE || o.otype == OT_BUBBLE) {
            if (a.size() != 4) { err = "bad sphere \"" + o.name + "\""; return false; }
            double r = fabs(a[3]);
            grow(a[0] - r, a[1] - r, a[2] - r); grow(a[0] + r, a[1] + r, a[2] + r);
            sbox(p.lo, p.hi);                        // ot/bbox.c:40-51
            if (r <= FTINY) continue;                // "zero radius": O_MISS (ot/sphere.c:73-76)
            p.kind = 1; p.r = r;
            for (int k = 0; k < 3; k++) p.c[k] = a[k];
        } else {                                     // ring, cone, cup, cylinder, tube: common/cone.c:44-153 getcone()
            const bool cyl = o.otype == OT_CYLINDER || o.otype == OT_TUBE, ring = o.otype == OT_RING;
            if (a.size() != (cyl ? 7u : 8u)) { err = "bad arguments for \"" + o.name + "\""; return false; }
            if (ring) {
                double r = std::max(fabs(a[6]), fabs(a[7]));
                grow(a[0] - r, a[1] - r, a[2] - r); grow(a[0] + r, a[1] + r, a[2] + r);
            } else {
                double r0 = fabs(a[6]), r1 = cyl ? r0 : fabs(a[7]);
                grow(a[0] - r0, a[1] - r0, a[2] - r0); grow(a[0] + r0, a[1] + r0, a[2] + r0);
                grow(a[3] - r1, a[4] - r1, a[5] - r1); grow(a[3] + r1, a[4] + r1, a[5] + r1);
            }
            int ip0, ip1; double r0, r1; bool degenerate = false;
            if (cyl) {
                if (fabs(a[6]) <= FTINY) degenerate = true;        // "illegal radii": getcone() returns NULL
                ip0 = 0; ip1 = 3; r0 = r1 = fabs(a[6]);
            } else {
                const int s0 = a[6] < -FTINY ? -1 : a[6] > FTINY ? 1 : 0, s1 = a[7] < -FTINY ? -1 : a[7] > FTINY ? 1 : 0;
                if (s0 + s1 == 0) degenerate = true;
                if (((s0 < 0) | (s1 < 0)) && ring) degenerate = true;
                const double c6 = a[6] * s0, c7 = a[7] * s1;
                if (c7 - c6 > FTINY) { ip0 = 0; ip1 = ring ? 0 : 3; r0 = c6; r1 = c7; }
                else if (c6 - c7 > FTINY) { ip0 = ring ? 0 : 3; ip1 = 0; r0 = c7; r1 = c6; }
                else { if (ring) degenerate = true; ip0 = 0; ip1 = 3; r0 = r1 = c6; }
            }
            if (!degenerate) {
                if (ring) { p.ad[0] = a[3]; p.ad[1] = a[4]; p.ad[2] = a[5]; }
                else for (int k = 0; k < 3; k++) p.ad[k] = a[ip1 + k] - a[ip0 + k];
                if (normalize3(p.ad) == 0.0) degenerate = true;    // "zero orientation"
            }
            if (degenerate) continue;                // getcone() == NULL: neither in the bounding box nor in the tree
            p.kind = 2; p.r0 = r0; p.r1 = r1;
            for (int k = 0; k < 3; k++) { p.p0[k] = a[ip0 + k]; p.p1[k] = a[ip1 + k]; }
            double cl[3], ch[3];                     // ot/bbox.c:58-69,116-137 circle2bbox() of the end circles
            for (int k = 0; k < 3; k++) { cl[k] = 1e10; ch[k] = -1e10; }
            for (int e = ring ? 1 : 0; e < 2; e++) {
                const double* c = e ? p.p1 : p.p0; const double rad = e ? r1 : r0;
                for (int k = 0; k < 3; k++) {
                    const double rr = sqrt(1. - p.ad[k] * p.ad[k]);
                    ch[k] = std::max(ch[k], c[k] + rr * rad); cl[k] = std::min(cl[k], c[k] - rr * rad);
                }
            }
            sbox(cl, ch);
        }
        B.prims.push_back(p);
    }
    const double OMARGIN = 10 * FTINY;
    if (keep_cube) {
        for (int k = 0; k < 3; k++) cuorg[k] = sc.cuorg[k];
        cusize = sc.cusize;
        for (int k = 0; k < 3 && bbmin[0] <= bbmax[0]; k++)
            if (bbmin[k] < cuorg[k] - OMARGIN || bbmax[k] > cuorg[k] + cusize + OMARGIN) {
                err = "boundary does not encompass scene (instance or mesh sticks out of the parent octree's cube)";
                return false;
            }
    } else {
        // ot/oconv.c:122-139: cube centred on the bounding box, with margin
        cuorg[0] = cuorg[1] = cuorg[2] = 0; cusize = 0;
        if (bbmin[0] <= bbmax[0]) {
            for (int k = 0; k < 3; k++) { bbmin[k] -= OMARGIN; bbmax[k] += OMARGIN; }
            for (int k = 0; k < 3; k++) cusize = std::max(cusize, bbmax[k] - bbmin[k]);
            for (int k = 0; k < 3; k++) cuorg[k] = (bbmax[k] + bbmin[k] - cusize) * .5;
        }
        // the reader parses the "%.12g" strings: build with exactly those values
        for (int k = 0; k < 3; k++) { snprintf(sbuf[k], 64, "%.12g", cuorg[k]); cuorg[k] = atof(sbuf[k]); }
        snprintf(sbuf[3], 64, "%.12g", cusize); cusize = atof(sbuf[3]);
    }
    B.mincusize = cusize / (maxres > 0 ? maxres : 16384) - FTINY;
    // With a CUDA device the overlap tests of every level run there (rb_octbuild_gpu.cu: same tests, same tree);
    // without one -- the CPU-only tools and tests -- the threaded host builder below does the same work.
    if (B.prims.size() >= 20000 && octbuild_device_available()) {
        std::string derr;
        if (build_tree_device(B.prims, B.objlim, B.mincusize, cuorg, cusize, B.top.nodes, B.top.pool, root, derr)) return true;
        if (getenv("RB_OCTBUILD_DEVICE_STRICT")) { err = "device octree build failed: " + derr; return false; }
        B.top.nodes.clear(); B.top.pool.clear();          // (out of device memory ...): build on the host
    }
    std::vector<int> all(B.prims.size());
    for (size_t i = 0; i < all.size(); i++) all[i] = (int)i;
    root = B.build(B.top, all, cuorg, cusize, 0);
    if (!B.err.empty()) { err = B.err; return false; }
    return true;
}

// Re-build sc's octree in memory over its current surface list (used after
// instances / meshes have been expanded into world-space surfaces).
bool rebuild_octree(Scene& sc, int objlim, int maxres, std::string& err) {
    Builder B;
    double cuorg[3], cusize; char sbuf[4][64]; int root;
    if (!build_tree(sc, B, objlim, maxres, true, cuorg, cusize, sbuf, root, err)) return false;
    sc.nodes = std::move(B.top.nodes);
    sc.leafpool = std::move(B.top.pool);          // same layout: [count, ids...], leaf word = -offset - 2
    sc.root = root;
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
    auto t0 = std::chrono::steady_clock::now();
    if (!build_tree(sc, B, objlim, maxres, false, cuorg, cusize, sbuf, root, err)) return false;
    if (getenv("RB_OCONV_TIMING"))
        fprintf(stderr, "rb_oconv: tree %.3f s (%zu nodes, %zu leaf-pool words)\n",
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), B.top.nodes.size() / 8, B.top.pool.size());

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
    const bool timing = getenv("RB_OCONV_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "rb_oconv: %s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
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
    lap("read scene");
    for (size_t i = 0; ok && i < sc.objs.size(); i++)
        if (rb::ot_is_volume(sc.objs[i].otype) && !sc.objs[i].expanded) {
            ok = false;
            err = "rb_oconv: " + sc.objs[i].tname + " \"" + sc.objs[i].name +
                  "\" cannot be placed by this builder (use the reference oconv for scenes with instances / meshes)";
        }
    if (ok) ok = rb::build_octree_file(sc, cmd, oct_path, objlim, maxres, err);
    lap("build + write");
    if (!ok && errbuf && errlen) { strncpy(errbuf, err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
    return ok ? 0 : -1;
}

extern "C" int rb_oconv(const char* rad_path, const char* oct_path, int objlim, int maxres, char* errbuf,
                        size_t errlen) {
    return rb_oconv_files(&rad_path, 1, nullptr, oct_path, objlim, maxres, errbuf, errlen);
}
