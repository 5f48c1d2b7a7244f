// rb_octtests.hpp -- the reference's cube / surface overlap tests, shared by the host octree builder
// (rb_octbuild.cpp) and the device one (rb_octbuild_gpu.cu): ot/o_face.c:40-99, ot/sphere.c:52-109,
// ot/o_cone.c:37-126 (STRICT), common/plocate.c:18-36, common/clip.c:17-72, common/face.c:121-162,
// common/fvect.c:76-99,131-156.  Plain IEEE double arithmetic (no contraction: the whole library is built
// with -fmad=false), so both builders take the same decisions.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define RB_HD __host__ __device__
#else
#define RB_HD
#endif

namespace rb {
namespace octt {

constexpr double FTINY = 1e-6;
constexpr int MAXSET = 8191;

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
constexpr int XPOS = 03, YPOS = 014, ZPOS = 060, BELOW = 025, ABOVE = 052;
RB_HD inline int plocate(const double p[3], const double mn[3], const double mx[3]) {
    int loc = 0;
    if (p[0] < mn[0] - FTINY) loc |= XPOS & BELOW; else if (p[0] > mx[0] + FTINY) loc |= XPOS & ABOVE;
    if (p[1] < mn[1] - FTINY) loc |= YPOS & BELOW; else if (p[1] > mx[1] + FTINY) loc |= YPOS & ABOVE;
    if (p[2] < mn[2] - FTINY) loc |= ZPOS & BELOW; else if (p[2] > mx[2] + FTINY) loc |= ZPOS & ABOVE;
    return loc;
}

// common/clip.c:17-72: Cohen-Sutherland, at most 6 chops; modifies the end points
RB_HD inline bool clip(double* ep1, double* ep2, const double mn[3], const double mx[3]) {
    int itlim = 6;
    int loc1 = plocate(ep1, mn, mx), loc2 = plocate(ep2, mn, mx);
    bool accept;
    while (!((accept = !(loc1 | loc2)) || (loc1 & loc2))) {
        if (itlim-- <= 0) return false;
        if (!loc1) { double* tp = ep1; ep1 = ep2; ep2 = tp; const int tl = loc1; loc1 = loc2; loc2 = tl; }
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

RB_HD inline bool fabseq(double a, double b) { return fabs(a - b) <= FTINY; }

// common/face.c:121-162 inface(): crossing count along +x in the projection that drops axis ax
RB_HD inline bool inface(const double p[3], const Prim& f) {
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

RB_HD inline double dist2(const double a[3], const double b[3]) {
    const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// common/fvect.c:76-99
RB_HD inline double dist2lseg(const double p[3], const double ep1[3], const double ep2[3]) {
    const double d = dist2(ep1, ep2), d1 = dist2(ep1, p);
    double d2 = dist2(ep2, p);
    if (d2 > d1) { if (d2 - d1 > d) return d1; }
    else if (d1 - d2 > d) return d2;
    d2 = d + d1 - d2;
    return d1 - 0.25 * d2 * d2 / d;
}

constexpr double ROOT3 = 1.732050808;

// common/fvect.c:131-156 normalize(), first-order shortcut included
RB_HD inline double normalize3(double v[3]) {
    double d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], len;
    if (d == 0.0) return 0.0;
    if ((d <= 1.0 + FTINY) & (d >= 1.0 - FTINY)) { len = 0.5 + 0.5 * d; d = 2.0 - len; }
    else { len = sqrt(d); d = 1.0 / len; }
    v[0] *= d; v[1] *= d; v[2] *= d;
    return len;
}

// ot/o_face.c:40-99
RB_HD inline bool face_in_cube(const Prim& f, const double org[3], double size) {
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
RB_HD inline bool sphere_in_cube(const Prim& s, const double org[3], double size) {
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
RB_HD inline bool cone_in_cube(const Prim& co, const double org[3], double size, double mincusize) {
    // (the reference recurses; an explicit stack gives the same OR over the same sub-cubes and runs on the device
    //  without a call stack: depth <= log2(cube / minimum cube) + 1, 14 for the default resolution)
    struct Frame { double o[3], s; int next; };
    Frame st[32];
    int sp = 0;
    st[0].o[0] = org[0]; st[0].o[1] = org[1]; st[0].o[2] = org[2]; st[0].s = size; st[0].next = -1;
    while (sp >= 0) {
        Frame& f = st[sp];
        if (f.next < 0) {
            double ep0[3], ep1[3], p[3], v[3];
            double r = f.s * 0.5;
            for (int i = 0; i < 3; i++) p[i] = f.o[i] + r;
            r *= ROOT3;
            for (int i = 0; i < 3; i++) v[i] = p[i] - co.p0[i];
            const double d = v[0] * co.ad[0] + v[1] * co.ad[1] + v[2] * co.ad[2];
            for (int i = 0; i < 3; i++) v[i] -= d * co.ad[i];
            bool out = false;
            if (normalize3(v) != 0.0) {                           // findcseg() found the segment
                for (int i = 0; i < 3; i++) { ep0[i] = co.r0 * v[i] + co.p0[i]; ep1[i] = co.r1 * v[i] + co.p1[i]; }
                if (dist2lseg(p, ep0, ep1) > (r + FTINY) * (r + FTINY)) out = true;
                else {
                    double cumin[3], cumax[3];
                    for (int i = 0; i < 3; i++) cumax[i] = (cumin[i] = f.o[i]) + f.s;
                    if (clip(ep0, ep1, cumin, cumax)) return true;
                }
            }
            if (out) { sp--; continue; }
            if (f.s * 0.5 < mincusize || sp >= 30) return true;  // cube too small
            f.next = 0;
        }
        if (f.next >= 8) { sp--; continue; }
        const int j = f.next++;
        const double half = f.s * 0.5;
        Frame& k = st[sp + 1];
        for (int i = 0; i < 3; i++) { k.o[i] = f.o[i]; if ((1 << i) & j) k.o[i] += half; }
        k.s = half; k.next = -1;
        sp++;
    }
    return false;
}

// the reference's own cube tests after a bounding-box reject whose margin (4 FTINY) is wider than any of theirs
RB_HD inline bool overlaps(const Prim& p, const double org[3], double size, double mincusize) {
    const double mg = p.kind == 2 ? size + 4 * FTINY : 4 * FTINY;   // o_cone() accepts by the cube's bounding sphere
    for (int k = 0; k < 3; k++)
        if (p.hi[k] < org[k] - mg || p.lo[k] > org[k] + size + mg) return false;
    if (p.kind == 0) return face_in_cube(p, org, size);
    if (p.kind == 1) return sphere_in_cube(p, org, size);
    return cone_in_cube(p, org, size, mincusize);
}

}  // namespace octt
}  // namespace rb
