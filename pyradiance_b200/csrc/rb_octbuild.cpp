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

const double FTINY = 1e-6;
const int MAXSET = 8191;

struct Prim {
    int obj;
    int kind;                 // 0 polygon, 1 sphere, 2 cone family
    double lo[3], hi[3];      // bounding box (quick reject only, with a margin wider than the exact tests')
    // polygon (common/face.c:35-106 getface)
    const double* va; int nv, ax;
    double n[3], off;
    // sphere
    double c[3], r;
    // cone family (common/cone.c:44-153 getcone): end points, radii, axis
    double p0[3], p1[3], r0, r1, ad[3];
};

// common/plocate.c:18-36, common/plocate.h
const int XPOS = 03, YPOS = 014, ZPOS = 060, BELOW = 025, ABOVE = 052;
inline int plocate(const double p[3], const double mn[3], const double mx[3]) {
    int loc = 0;
    if (p[0] < mn[0] - FTINY) loc |= XPOS & BELOW; else if (p[0] > mx[0] + FTINY) loc |= XPOS & ABOVE;
    if (p[1] < mn[1] - FTINY) loc |= YPOS & BELOW; else if (p[1] > mx[1] + FTINY) loc |= YPOS & ABOVE;
    if (p[2] < mn[2] - FTINY) loc |= ZPOS & BELOW; else if (p[2] > mx[2] + FTINY) loc |= ZPOS & ABOVE;
    return loc;
}

// common/clip.c:17-72: Cohen-Sutherland, at most 6 chops; modifies the end points
bool clip(double* ep1, double* ep2, const double mn[3], const double mx[3]) {
    int itlim = 6;
    int loc1 = plocate(ep1, mn, mx), loc2 = plocate(ep2, mn, mx);
    bool accept;
    while (!((accept = !(loc1 | loc2)) || (loc1 & loc2))) {
        if (itlim-- <= 0) return false;
        if (!loc1) { std::swap(ep1, ep2); std::swap(loc1, loc2); }
        for (int i = 0; i < 3; i++) {
            double d;
            const int pos = 3 << (i << 1);
            if (loc1 & pos & BELOW) { d = (mn[i] - ep1[i]) / (ep2[i] - ep1[i]); ep1[i] = mn[i]; }
            else if (loc1 & pos & ABOVE) { d = (mx[i] - ep1[i]) / (ep2[i] - ep1[i]); ep1[i] = mx[i]; }
            else continue;
            for (int j = 0; j < 3; j++) if (j != i) ep1[j] += (ep2[j] - ep1[j]) * d;
            break;
        }
        loc1 = plocate(ep1, mn, mx);
    }
    return accept;
}

inline bool fabseq(double a, double b) { return fabs(a - b) <= FTINY; }

// common/face.c:121-162 inface(): crossing count along +x in the projection that drops axis ax
bool inface(const double p[3], const Prim& f) {
    int xi = f.ax + 1; if (xi >= 3) xi -= 3;
    int yi = xi + 1; if (yi >= 3) yi -= 3;
    const double x = p[xi], y = p[yi];
    int n = f.nv;
    const double* p0 = f.va + 3 * (n - 1);
    const double* p1 = f.va;
    int ncross = 0;
    while (n--) {
        if (fabseq(p0[yi], y) && fabseq(p1[yi], y) && ((p0[xi] > x) ^ (p1[xi] > x))) return true;
        if ((p0[yi] > y) ^ (p1[yi] > y)) {
            const int tst = (p0[xi] > x) + (p1[xi] > x);
            if (tst == 2) ncross++;
            else if (tst) {
                const double prodA = (p0[yi] - y) * (p1[xi] - x);
                const double prodB = (p0[xi] - x) * (p1[yi] - y);
                if (fabseq(prodA, prodB)) return true;
                ncross += (p1[yi] > p0[yi]) ^ (prodA > prodB);
            } else if (fabseq(p0[xi], x) && fabseq(p1[xi], x)) return true;
        }
        p0 = p1;
        p1 += 3;
    }
    return ncross & 1;
}

inline double dist2(const double a[3], const double b[3]) {
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// common/fvect.c:76-99
double dist2lseg(const double p[3], const double ep1[3], const double ep2[3]) {
    const double d = dist2(ep1, ep2), d1 = dist2(ep1, p);
    double d2 = dist2(ep2, p);
    if (d2 > d1) { if (d2 - d1 > d) return d1; }
    else if (d1 - d2 > d) return d2;
    d2 = d + d1 - d2;
    return d1 - 0.25 * d2 * d2 / d;
}

const double ROOT3 = 1.732050808;

// common/fvect.c:131-156 normalize(), first-order shortcut included
double normalize3(double v[3]) {
    double d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], len;
    if (d == 0.0) return 0.0;
    if ((d <= 1.0 + FTINY) & (d >= 1.0 - FTINY)) { len = 0.5 + 0.5 * d; d = 2.0 - len; }
    else { len = sqrt(d); d = 1.0 / len; }
    v[0] *= d; v[1] *= d; v[2] *= d;
    return len;
}

// ot/o_face.c:40-99
bool face_in_cube(const Prim& f, const double org[3], double size) {
    double cumin[3], cumax[3], v1[3], v2[3];
    for (int j = 0; j < 3; j++) cumax[j] = (cumin[j] = org[j] - FTINY) + size + 2.0 * FTINY;
    int vloc = ABOVE | BELOW;
    for (int i = 0; i < f.nv; i++) {
        const int j = plocate(f.va + 3 * i, cumin, cumax);
        if (j) vloc &= j; else return true;               // vertex inside
    }
    if (vloc) return false;                               // all to one side
    for (int i = 0; i < f.nv; i++) {                      // edges
        const int j = i + 1 >= f.nv ? 0 : i + 1;
        for (int k = 0; k < 3; k++) { v1[k] = f.va[3 * i + k]; v2[k] = f.va[3 * j + k]; }
        if (clip(v1, v2, cumin, cumax)) return true;
    }
    for (int j = 0; j < 3; j++) {                         // does the cube cut the plane?
        if (f.n[j] > 0.0) { v1[j] = cumin[j]; v2[j] = cumax[j]; } else { v1[j] = cumax[j]; v2[j] = cumin[j]; }
    }
    const double d1 = v1[0] * f.n[0] + v1[1] * f.n[1] + v1[2] * f.n[2] - f.off;
    if (d1 > FTINY) return false;
    const double d2 = v2[0] * f.n[0] + v2[1] * f.n[1] + v2[2] * f.n[2] - f.off;
    if (d2 < -FTINY) return false;
    for (int j = 0; j < 3; j++) v1[j] = (v1[j] * d2 - v2[j] * d1) / (d2 - d1);
    return inface(v1, f);                                 // the cube diagonal passes through the face
}

// ot/sphere.c:52-109
bool sphere_in_cube(const Prim& s, const double org[3], double size) {
    double v1[3];
    const double rad = s.r;
    const double d1 = ROOT3 / 2.0 * size;                 // bounding radius of the cube
    double d2 = size * 0.5;
    for (int i = 0; i < 3; i++) v1[i] = org[i] + d2 - s.c[i];
    d2 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    if (d2 > (rad + d1 + FTINY) * (rad + d1 + FTINY)) return false;
    if (d1 < rad) {
        if (d2 < (rad - d1 - FTINY) * (rad - d1 - FTINY)) return false;     // cube inside the sphere
        if (d2 < (rad + FTINY) * (rad + FTINY)) return true;                // cube centre inside
    }
    for (int i = 0; i < 3; i++) {
        if (s.c[i] < org[i]) v1[i] = org[i] - s.c[i];
        else if (s.c[i] > org[i] + size) v1[i] = s.c[i] - (org[i] + size);
        else v1[i] = 0;
    }
    return v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2] <= (rad + FTINY) * (rad + FTINY);
}

// ot/o_cone.c:37-126 (STRICT): nearest generator segment against the cube's bounding sphere, then the
// line clipper, then the eight sub-cubes down to the minimum cube size
bool cone_in_cube(const Prim& co, const double org[3], double size, double mincusize) {
    double ep0[3], ep1[3], p[3], v[3];
    double r = size * 0.5;
    for (int i = 0; i < 3; i++) p[i] = org[i] + r;
    r *= ROOT3;
    for (int i = 0; i < 3; i++) v[i] = p[i] - co.p0[i];
    const double d = v[0] * co.ad[0] + v[1] * co.ad[1] + v[2] * co.ad[2];
    for (int i = 0; i < 3; i++) v[i] -= d * co.ad[i];
    if (normalize3(v) != 0.0) {                           // findcseg() found the segment
        for (int i = 0; i < 3; i++) { ep0[i] = co.r0 * v[i] + co.p0[i]; ep1[i] = co.r1 * v[i] + co.p1[i]; }
        if (dist2lseg(p, ep0, ep1) > (r + FTINY) * (r + FTINY)) return false;
        double cumin[3], cumax[3];
        for (int i = 0; i < 3; i++) cumax[i] = (cumin[i] = org[i]) + size;
        if (clip(ep0, ep1, cumin, cumax)) return true;
    }
    const double half = size * 0.5;
    if (half < mincusize) return true;                    // cube too small
    for (int j = 0; j < 8; j++) {
        double ko[3];
        for (int i = 0; i < 3; i++) { ko[i] = org[i]; if ((1 << i) & j) ko[i] += half; }
        if (cone_in_cube(co, ko, half, mincusize)) return true;
    }
    return false;
}

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
    bool overlaps(const Prim& p, const double org[3], double size) const {
        const double mg = p.kind == 2 ? size + 4 * FTINY : 4 * FTINY;   // o_cone() accepts by the cube's bounding sphere
        for (int k = 0; k < 3; k++)
            if (p.hi[k] < org[k] - mg || p.lo[k] > org[k] + size + mg) return false;
        if (p.kind == 0) return face_in_cube(p, org, size);
        if (p.kind == 1) return sphere_in_cube(p, org, size);
        return cone_in_cube(p, org, size, mincusize);
    }
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
