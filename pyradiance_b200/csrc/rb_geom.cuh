// rb_geom.cuh -- octree walk and surface intersection on the device.
//
// Restates, as an iterative integer-coordinate walk with a shared-memory node
// stack, the reference's recursive octant DDA and its intersectors:
//   localhit/raymove/checkhit   src/radiance/rt/raytrace.c:595-760
//   rayhit + rayreject          src/radiance/rt/raytrace.c:535-591
//   incube                      src/radiance/common/octree.c:115-126
//   o_face + inface             src/radiance/rt/o_face.c:16-62, common/face.c:121-162
//   o_sphere                    src/radiance/rt/sphere.c:16-83
//   o_cone + quadratic          src/radiance/rt/o_cone.c:17-146, common/zeroes.c:19-54
// All geometry is IEEE double (RREAL); this translation unit is compiled with
// -fmad=false so that products and sums round exactly like the reference's
// non-fused x86-64 code.  Differences from the reference, by design:
//   * cube origins come from integer cell coordinates (one fused rounding)
//     instead of a chain of += cusize (differs by ulps of the cube origin);
//   * objects already tested in an earlier leaf are tested again instead of
//     being filtered through a per-ray checked set: rayreject() makes the
//     re-test a no-op (o == r->ro, or t > rot + FTINY, or the same pairwise
//     tie decision).
#pragma once
#include "rb_device.cuh"

namespace rb {

struct Hit {
    int robj;           // -1 none
    double rot, rod;
};

__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// raytrace.c:535-575 -- returns true if candidate (t, rod) on object `id`
// must be rejected in favour of the current hit.
__device__ __forceinline__ bool rayreject(const DScene& S, int id, int4 hnew, const Hit& h,
                                          bool aft, double t, double rod) {
    if ((t <= RB_FTINY) | (t > h.rot + RB_FTINY)) return true;
    if (t < h.rot - RB_FTINY) return false;
    // coincident point, so decide...
    if (id == h.robj) return true;
    if (h.robj < 0) return aft ? true /* Aftplane has no material: see below */ : false;
    int4 hold = __ldg(&S.objhdr[h.robj]);
    int fnew = hnew.x >> 8, fold = hold.x >> 8;
    bool mnew = fnew & PF_HASMAT, mray = fold & PF_HASMAT;
    if (!mnew) {
        if (mray) return true;
    } else if (!mray) {
        return false;
    } else if (fnew & PF_TRANSP) {
        if (!(fold & PF_TRANSP)) return true;
    } else if (fold & PF_TRANSP) {
        return false;
    }
    if (rod <= 0) {
        if (h.rod > 0) return true;
    } else if (h.rod <= 0) {
        return false;
    }
    return hold.y >= hnew.y;     // later modifier definition wins tie
}

// common/face.c:121-162 inface() on the pre-projected 2-D vertices
__device__ __forceinline__ bool inface2d(const double* __restrict__ vp, int nv, double x, double y) {
#define FABSEQ(a, b) (fabs((a) - (b)) <= RB_FTINY)
    const double2* v = reinterpret_cast<const double2*>(vp);
    double2 p0 = __ldg(&v[nv - 1]);
    int ncross = 0;
    for (int n = 0; n < nv; n++) {
        double2 p1 = __ldg(&v[n]);
        if (FABSEQ(p0.y, y) && FABSEQ(p1.y, y) && ((p0.x > x) ^ (p1.x > x))) return true;
        if ((p0.y > y) ^ (p1.y > y)) {
            int tst = (p0.x > x) + (p1.x > x);
            if (tst == 2) ncross++;
            else if (tst) {
                double prodA = (p0.y - y) * (p1.x - x);
                double prodB = (p0.x - x) * (p1.y - y);
                if (FABSEQ(prodA, prodB)) return true;
                ncross += (p1.y > p0.y) ^ (prodA > prodB);
            } else if (FABSEQ(p0.x, x) && FABSEQ(p1.x, x)) return true;
        }
        p0 = p1;
    }
    return ncross & 1;
#undef FABSEQ
}

// common/zeroes.c:19-54
__device__ __forceinline__ int quadratic(double r[2], double a, double b, double c) {
    int first;
    if (a < -RB_FTINY) first = 1;
    else if (a > RB_FTINY) first = 0;
    else if (fabs(b) > RB_FTINY) { r[0] = -c / b; return 1; }
    else return 0;
    b *= 0.5;
    double disc = b * b - a * c;
    if (disc < -RB_FTINY * RB_FTINY) return 0;
    if (disc <= RB_FTINY * RB_FTINY) { r[0] = -b / a; return 1; }
    disc = sqrt(disc);
    r[first] = (-b - disc) / a;
    r[1 - first] = (-b + disc) / a;
    return 2;
}

// Test one object; on acceptance updates h and returns true.
__device__ __noinline__ bool hit_object(const DScene& S, int id, const double org[3],
                                        const double dir[3], Hit& h, bool aft,
                                        unsigned* errflag, unsigned* errobj) {
    int4 hd = __ldg(&S.objhdr[id]);
    int kind = hd.x & 0xff;
    const double* g = S.geom + hd.w;
    if (kind == PK_FACE) {
        const double2* g2 = reinterpret_cast<const double2*>(g);
        double2 n01 = __ldg(&g2[0]), n2o = __ldg(&g2[1]);
        double rdot = -(dir[0] * n01.x + dir[1] * n01.y + dir[2] * n2o.x);
        if ((rdot <= RB_FTINY) & (rdot >= -RB_FTINY)) return false;
        double t = ((org[0] * n01.x + org[1] * n01.y + org[2] * n2o.x) - n2o.y) / rdot;
        if (rayreject(S, id, hd, h, aft, t, rdot)) return false;
        int ax = (hd.x >> 10) & 3;
        int xi = ax + 1; if (xi >= 3) xi -= 3;
        int yi = xi + 1; if (yi >= 3) yi -= 3;
        double p[3] = {org[0] + t * dir[0], org[1] + t * dir[1], org[2] + t * dir[2]};
        double x = xi == 0 ? p[0] : xi == 1 ? p[1] : p[2];
        double y = yi == 0 ? p[0] : yi == 1 ? p[1] : p[2];
        if (!inface2d(g + 8, (hd.x >> 16) & 0xffff, x, y)) return false;
        h.robj = id; h.rot = t; h.rod = rdot;
        return true;
    }
    if (kind == PK_SPHERE || kind == PK_BUBBLE) {
        double a = 0, b = 0, c = 0, root[2];
        for (int i = 0; i < 3; i++) {
            a += dir[i] * dir[i];
            double t = org[i] - g[i];
            b += 2.0 * dir[i] * t;
            c += t * t;
        }
        c -= g[3] * g[3];
        int nroots = quadratic(root, a, b, c);
        int i; double t = 0;
        for (i = 0; i < nroots; i++)
            if ((t = root[i]) > RB_FTINY) break;
        if (i >= nroots) return false;
        double rodc = 1 - 2 * ((i > 0) ^ (kind == PK_BUBBLE));
        if (rayreject(S, id, hd, h, aft, t, rodc)) return false;
        // rod proper = -dir . ron (sphere.c:78)
        double ar = g[3] * (1 - 2 * (kind == PK_BUBBLE));
        double rod = 0;
        for (int k = 0; k < 3; k++) {
            double rp = org[k] + dir[k] * t;
            rod += dir[k] * ((rp - g[k]) / ar);
        }
        h.robj = id; h.rot = t; h.rod = -rod;
        return true;
    }
    if (kind >= PK_CONE && kind <= PK_RING) {
        const double* ad = g; double al = g[3];
        const double* p0 = g + 4;
        double r0 = g[8], r1 = g[9];
        const double* tm = g + 12;       // tm[i][j] at tm[i*3+j], i = 0..3
        double rox[3], rdx[3];
        for (int j = 0; j < 3; j++) {
            rdx[j] = dir[0] * tm[0 + j] + dir[1] * tm[3 + j] + dir[2] * tm[6 + j];
            rox[j] = org[0] * tm[0 + j] + org[1] * tm[3 + j] + org[2] * tm[6 + j];
            rox[j] += tm[9 + j];
        }
        double a, b, c, root[2];
        if (kind == PK_CONE || kind == PK_CUP) {
            a = rdx[0] * rdx[0] + rdx[1] * rdx[1] - rdx[2] * rdx[2];
            b = 2.0 * (rdx[0] * rox[0] + rdx[1] * rox[1] - rdx[2] * rox[2]);
            c = rox[0] * rox[0] + rox[1] * rox[1] - rox[2] * rox[2];
        } else if (kind == PK_CYL || kind == PK_TUBE) {
            a = rdx[0] * rdx[0] + rdx[1] * rdx[1];
            b = 2.0 * (rdx[0] * rox[0] + rdx[1] * rox[1]);
            c = rox[0] * rox[0] + rox[1] * rox[1] - r0 * r0;
        } else {  // ring
            if ((rdx[2] <= RB_FTINY) & (rdx[2] >= -RB_FTINY)) return false;
            root[0] = -rox[2] / rdx[2];
            if (rayreject(S, id, hd, h, aft, root[0], -rdx[2])) return false;
            b = root[0] * rdx[0] + rox[0];
            c = root[0] * rdx[1] + rox[1];
            a = b * b + c * c;
            if (a > r1 * r1 || a < r0 * r0) return false;
            h.robj = id; h.rot = root[0]; h.rod = -rdx[2];
            return true;
        }
        int nroots = quadratic(root, a, b, c);
        for (int rn = 0; rn < nroots; rn++) {
            if (root[rn] <= RB_FTINY) continue;
            if (root[rn] > h.rot + RB_FTINY) break;
            double px[3], dx[3];
            for (int k = 0; k < 3; k++) { px[k] = org[k] + root[rn] * dir[k]; dx[k] = px[k] - p0[k]; }
            b = dot3(dx, ad);
            if (b < 0.0) continue;
            if (b > al) continue;
            double rodc = 1 - 2 * ((rn > 0) ^ ((kind == PK_CUP) | (kind == PK_TUBE)));
            if (rayreject(S, id, hd, h, aft, root[rn], rodc)) break;
            // normal (o_cone.c:114-137) only to get rod
            double sl = g[7], ron[3];
            if (kind == PK_CYL) a = r0;
            else if (kind == PK_TUBE) a = -r0;
            else {
                c = r1 - r0;
                a = r0 + b * c / al;
                if (kind == PK_CUP) { c = -c; a = -a; }
            }
            for (int k = 0; k < 3; k++) ron[k] = (dx[k] - b * ad[k]) / a;
            if ((kind == PK_CONE) | (kind == PK_CUP))
                for (int k = 0; k < 3; k++) ron[k] = (al * ron[k] - c * ad[k]) / sl;
            a = dot3(ron, ron);
            if ((a > 1. + RB_FTINY) | (a < 1. - RB_FTINY)) {
                c = 1. / (.5 + .5 * a);
                ron[0] *= c; ron[1] *= c; ron[2] *= c;
            }
            h.robj = id; h.rot = root[rn]; h.rod = -dot3(dir, ron);
            return true;
        }
        return false;
    }
    if (kind == PK_UNSUPPORTED) {
        atomicOr(errflag, RB_ERR_UNSUP_PRIM);
        *errobj = (unsigned)id;
    }
    return false;
}

// Surface normal and hit point of an accepted hit (recomputed once per ray
// instead of being carried through the walk).
__device__ __noinline__ void hit_frame(const DScene& S, const Hit& h, const double org[3],
                                       const double dir[3], double rop[3], double ron[3]) {
    for (int k = 0; k < 3; k++) rop[k] = org[k] + h.rot * dir[k];
    int4 hd = __ldg(&S.objhdr[h.robj]);
    int kind = hd.x & 0xff;
    const double* g = S.geom + hd.w;
    if (kind == PK_FACE) { ron[0] = g[0]; ron[1] = g[1]; ron[2] = g[2]; return; }
    if (kind == PK_SPHERE || kind == PK_BUBBLE) {
        double ar = g[3] * (1 - 2 * (kind == PK_BUBBLE));
        for (int k = 0; k < 3; k++) {
            rop[k] = org[k] + dir[k] * h.rot;
            ron[k] = (rop[k] - g[k]) / ar;
        }
        return;
    }
    if (kind == PK_RING) { ron[0] = g[0]; ron[1] = g[1]; ron[2] = g[2]; return; }
    const double* ad = g; double al = g[3];
    const double* p0 = g + 4;
    double r0 = g[8], r1 = g[9], sl = g[7];
    double dx[3], a, b, c = 0;
    for (int k = 0; k < 3; k++) dx[k] = rop[k] - p0[k];
    b = dot3(dx, ad);
    if (kind == PK_CYL) a = r0;
    else if (kind == PK_TUBE) a = -r0;
    else {
        c = r1 - r0;
        a = r0 + b * c / al;
        if (kind == PK_CUP) { c = -c; a = -a; }
    }
    for (int k = 0; k < 3; k++) ron[k] = (dx[k] - b * ad[k]) / a;
    if ((kind == PK_CONE) | (kind == PK_CUP))
        for (int k = 0; k < 3; k++) ron[k] = (al * ron[k] - c * ad[k]) / sl;
    a = dot3(ron, ron);
    if ((a > 1. + RB_FTINY) | (a < 1. - RB_FTINY)) {
        c = 1. / (.5 + .5 * a);
        ron[0] *= c; ron[1] *= c; ron[2] *= c;
    }
}

struct WalkStats { unsigned nodes, leafents, prims; };

// The polygon test of o_face(), inlined in the walk (it is >95 % of all tests).
__device__ __forceinline__ void hit_face(const DScene& S, int id, int4 hd, const double* __restrict__ g,
                                         const double org[3], const double dir[3], Hit& h, bool aft) {
    const double2* g2 = reinterpret_cast<const double2*>(g);
    double2 n01 = __ldg(&g2[0]), n2o = __ldg(&g2[1]);
    double rdot = -(dir[0] * n01.x + dir[1] * n01.y + dir[2] * n2o.x);
    if ((rdot <= RB_FTINY) & (rdot >= -RB_FTINY)) return;
    double t = ((org[0] * n01.x + org[1] * n01.y + org[2] * n2o.x) - n2o.y) / rdot;
    if (rayreject(S, id, hd, h, aft, t, rdot)) return;
    int ax = (hd.x >> 10) & 3;
    double p0 = org[0] + t * dir[0], p1 = org[1] + t * dir[1], p2 = org[2] + t * dir[2];
    double x = ax == 0 ? p1 : ax == 1 ? p2 : p0;      // xi = (ax+1)%3
    double y = ax == 0 ? p2 : ax == 1 ? p0 : p1;      // yi = (ax+2)%3
    // 2-D bounding box first: outside by more than FTINY can never be "in"
    // (no edge straddles y / all straddling edges on one side, and none of
    // inface()'s three FABSEQ cases can fire); well inside an exact axis-aligned
    // rectangle is always "in".  Only the FTINY border zone runs the edge loop.
    double2 bx = __ldg(&g2[2]), by = __ldg(&g2[3]);
    if ((x < bx.x - RB_FTINY) | (x > bx.y + RB_FTINY) | (y < by.x - RB_FTINY) | (y > by.y + RB_FTINY)) return;
    bool in = ((hd.x >> 12) & 1) && (x > bx.x + RB_FTINY) & (x < bx.y - RB_FTINY) & (y > by.x + RB_FTINY) &
                                        (y < by.y - RB_FTINY);
    if (!in && !inface2d(g + 8, (hd.x >> 16) & 0xffff, x, y)) return;
    h.robj = id; h.rot = t; h.rod = rdot;
}

// sourcehit() lives in rb_shade.cuh; declared here for the retire step
__device__ __forceinline__ int sourcehit(const DScene& S, const double dir[3], int rsrc, int crtype);

struct TraceIO {
    const QRay* __restrict__ qin;
    unsigned nin;
    HitRec* __restrict__ hits;
    unsigned* next;              // global fetch counter
};
#ifndef RB_MAILBOX
#define RB_MAILBOX 0
#endif
#ifndef RB_FETCH_MIN
#define RB_FETCH_MIN 6           // refill a warp when this many lanes are idle
#endif

// walk_rays(): persistent-thread localhit().  Every warp keeps pulling rays
// from the queue: a lane whose ray is finished retires it (writes its HitRec,
// running sourcehit() for misses) and, as soon as RB_FETCH_MIN lanes of the
// warp are idle, the warp reserves that many queue slots with one atomicAdd and
// the idle lanes start new rays.  Ray lengths are heavy-tailed (most rays end
// in nearby clutter, a few cross the whole room), so without the refill a warp
// ran at < 5 live lanes on average (profiles/r1_notes.md).
// The walk itself is organised in warp-synchronous phases -- walk to the next
// full leaf / test the leaf's surfaces / accept-or-continue -- with a
// __syncwarp() after each, so that lanes re-converge every phase.  `stk` is
// this thread's column of the shared-memory node stack.  MUST be called by all
// 32 lanes of a warp together.
__device__ __forceinline__ void walk_rays(const DScene& S, const TraceIO io, volatile int* stk, int stride,
                                          WalkStats& ws, unsigned& nretired, unsigned* errflag,
                                          unsigned* errobj) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const double cs = S.cusize;
    const int2* __restrict__ pool = reinterpret_cast<const int2*>(S.leafpool);
    double org[3] = {0, 0, 0}, dir[3] = {0, 0, 1}, pos[3] = {0, 0, 0}, rmax = 0, size = cs;
    Hit h; h.robj = -1; h.rot = RB_FHUGE; h.rod = 1.0;
    int dirf = 0, w = -1, L = 0, rsrc = -1, crtype = 0;
    unsigned ix = 0, iy = 0, iz = 0, ridx = 0;
    bool aft = false, need_adv = false, done = true, result = false, have = false, exhausted = false;
#if RB_MAILBOX
    int4 mb = make_int4(-1, -1, -1, -1);
#endif
    for (;;) {
        // ---- retire finished rays ----
        if (have & done) {
            HitRec o;
            o.rot = h.rot; o.rod = h.rod; o.robj = h.robj; o.local = 1;
            if (!result) {
                o.rot = RB_FHUGE; o.rod = 1.0; o.robj = -1; o.local = 0;
                if (!(rmax > RB_FTINY)) {            // aft-clipped rays never see sources
                    int sn = sourcehit(S, dir, rsrc, crtype);
                    if (sn >= 0) o.robj = S.srcs[sn].so;
                }
            }
            io.hits[ridx] = o;
            have = false;
            nretired++;
        }
        // ---- refill idle lanes ----
        unsigned idle = __ballot_sync(FULL, !have);
        if (idle == FULL && exhausted) break;
        if (!exhausted && (__popc(idle) >= RB_FETCH_MIN)) {
            int n = __popc(idle);
            int leader = __ffs(idle) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(io.next, (unsigned)n);
            base = __shfl_sync(FULL, base, leader);
            if (base + n >= io.nin) exhausted = true;
            unsigned my = base + __popc(idle & ((1u << lane) - 1));
            if (!have && my < io.nin) {
                const double2* q2 = reinterpret_cast<const double2*>(&io.qin[my]);
                double2 a = __ldg(&q2[0]), b = __ldg(&q2[1]), c = __ldg(&q2[2]), d = __ldg(&q2[3]);
                org[0] = a.x; org[1] = a.y; org[2] = b.x; dir[0] = b.y; dir[1] = c.x; dir[2] = c.y; rmax = d.x;
                crtype = io.qin[my].info & 0x3ff; rsrc = io.qin[my].rsrc;
                ridx = my; have = true;
                // ---- localhit() prologue (raytrace.c:604-651) ----
                dirf = 0;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    pos[i] = org[i];
                    if (dir[i] > 1e-7) dirf |= 1 << i;
                    else if (dir[i] < -1e-7) dirf |= 0x10 << i;
                }
                h.robj = -1; h.rot = RB_FHUGE; h.rod = 1.0;
                done = !dirf; result = false; aft = false; need_adv = false;
                if (!done && rmax > RB_FTINY) { aft = true; h.rot = rmax; }
                if (!done) {
                    bool in = !(S.cuorg[0] > pos[0] || pos[0] >= S.cuorg[0] + cs ||
                                S.cuorg[1] > pos[1] || pos[1] >= S.cuorg[1] + cs ||
                                S.cuorg[2] > pos[2] || pos[2] >= S.cuorg[2] + cs);
                    if (!in) {                        // find global cube entrance point
                        double t = 0.0;
#pragma unroll
                        for (int i = 0; i < 3; i++) {
                            double dt;
                            if (dirf & (1 << i)) dt = S.cuorg[i];
                            else if (dirf & (0x10 << i)) dt = S.cuorg[i] + cs;
                            else continue;
                            dt = (dt - org[i]) / dir[i];
                            if (dt > t) t = dt;
                        }
                        t += RB_FTINY;
                        if (t >= h.rot) done = true;
                        else {
#pragma unroll
                            for (int i = 0; i < 3; i++) pos[i] = pos[i] + dir[i] * t;
                            in = !(S.cuorg[0] > pos[0] || pos[0] >= S.cuorg[0] + cs ||
                                   S.cuorg[1] > pos[1] || pos[1] >= S.cuorg[1] + cs ||
                                   S.cuorg[2] > pos[2] || pos[2] >= S.cuorg[2] + cs);
                            if (!in) done = true;
                        }
                    }
                }
                w = S.root; L = 0; ix = iy = iz = 0; size = cs;
#if RB_MAILBOX
                mb = make_int4(-1, -1, -1, -1);
#endif
            }
        }
        __syncwarp();

        // ---- phase A: walk (descend / skip empty cubes) until standing in a
        //      fresh full leaf.  One loop, two short bodies, so lanes re-join
        //      every iteration (raymove, raytrace.c:668-738) ----
        if (!done) {
            for (;;) {
                if (w >= 0) {                         // descend one level
                    stk[L * stride] = w;
                    double half = size * 0.5;
                    double lox = fma((double)ix, size, S.cuorg[0]);
                    double loy = fma((double)iy, size, S.cuorg[1]);
                    double loz = fma((double)iz, size, S.cuorg[2]);
                    int br = 0;
                    ix <<= 1; iy <<= 1; iz <<= 1;
                    if (pos[0] >= lox + half) { br |= 1; ix |= 1; }
                    if (pos[1] >= loy + half) { br |= 2; iy |= 1; }
                    if (pos[2] >= loz + half) { br |= 4; iz |= 1; }
                    w = __ldg(&S.nodes[(size_t)w * 8 + br]);
                    ws.nodes++;
                    size = half; L++;
                    continue;
                }
                if ((w < -1) & !need_adv) break;      // arrived at a full leaf
                double lox = fma((double)ix, size, S.cuorg[0]);
                double loy = fma((double)iy, size, S.cuorg[1]);
                double loz = fma((double)iz, size, S.cuorg[2]);
                double hix = lox + size, hiy = loy + size, hiz = loz + size;
                if ((w == -1) & aft & (h.robj < 0)) { // aft-plane point in an empty leaf (:709-710)
                    double px = org[0] + h.rot * dir[0];
                    double py = org[1] + h.rot * dir[1];
                    double pz = org[2] + h.rot * dir[2];
                    if (!(lox > px || px >= hix || loy > py || py >= hiy || loz > pz || pz >= hiz)) {
                        done = true; result = false;
                        break;
                    }
                }
                // advance to next cube (raytrace.c:712-738)
                int ax = 0;
                double t;
                if (dirf & 0x11) {
                    double dt = (dirf & 1) ? hix : lox;
                    t = (dt - pos[0]) / dir[0];
                    ax = 0;
                } else t = RB_FHUGE;
                if (dirf & 0x22) {
                    double dt = (dirf & 2) ? hiy : loy;
                    dt = (dt - pos[1]) / dir[1];
                    if (dt < t) { t = dt; ax = 1; }
                }
                if (dirf & 0x44) {
                    double dt = (dirf & 4) ? hiz : loz;
                    dt = (dt - pos[2]) / dir[2];
                    if (dt < t) { t = dt; ax = 2; }
                }
#pragma unroll
                for (int i = 0; i < 3; i++) pos[i] = pos[i] + dir[i] * t;
                // step to the neighbour, ascending on overflow (raytrace.c:688-706):
                // climb while the cell coordinate along ax cannot move that way
                bool positive = dirf & (1 << ax);
                unsigned ia = ax == 0 ? ix : ax == 1 ? iy : iz;
                unsigned blocked = positive ? ia : ~ia;            // trailing ones = levels to climb
                int up = (~blocked) ? __ffs(~blocked) - 1 : 32;    // number of trailing one bits
                if (up >= L) { done = true; result = (h.robj >= 0); break; }   // left the scene cube
                ix >>= up; iy >>= up; iz >>= up; L -= up;
                size = ldexp(size, up);
                if (ax == 0) ix ^= 1; else if (ax == 1) iy ^= 1; else iz ^= 1;
                int br = (ix & 1) | ((iy & 1) << 1) | ((iz & 1) << 2);
                w = __ldg(&S.nodes[(size_t)stk[(L - 1) * stride] * 8 + br]);
                ws.nodes++;
                need_adv = false;
            }
        }
        __syncwarp();
        // ---- phase B: test the leaf's surfaces, highest index first (rayhit) ----
        {
            int cnt = 0;
            const int2* set = pool;
            if (!done) {
                set = pool + (-w - 2);
                cnt = __ldg(&set[0]).x;
                ws.leafents += cnt + 1;
                ws.prims += cnt;
            }
            for (int k = cnt; k > 0; k--) {
                int2 ent = __ldg(&set[k]);
#if RB_MAILBOX
                // mailbox: surfaces spanning several leaves were already tested for
                // this ray; the re-test is a no-op (see header), so skip it
                if ((ent.x == mb.x) | (ent.x == mb.y) | (ent.x == mb.z) | (ent.x == mb.w)) { ws.prims--; continue; }
                mb.w = mb.z; mb.z = mb.y; mb.y = mb.x; mb.x = ent.x;
#endif
                const double* g = S.geom + ent.y;
                int4 hd = __ldg(reinterpret_cast<const int4*>(g - 2));
                if ((hd.x & 0xff) == PK_FACE) hit_face(S, ent.x, hd, g, org, dir, h, aft);
                else hit_object(S, ent.x, org, dir, h, aft, errflag, errobj);
            }
        }
        __syncwarp();
        // ---- phase C: checkhit (raytrace.c:756-759): hit OK if in current cube ----
        if (!done) {
            need_adv = true;
            if (h.robj >= 0) {
                double lox = fma((double)ix, size, S.cuorg[0]);
                double loy = fma((double)iy, size, S.cuorg[1]);
                double loz = fma((double)iz, size, S.cuorg[2]);
                double px = org[0] + h.rot * dir[0];
                double py = org[1] + h.rot * dir[1];
                double pz = org[2] + h.rot * dir[2];
                if (!(lox > px || px >= lox + size || loy > py || py >= loy + size || loz > pz || pz >= loz + size)) {
                    done = true; result = true;
                }
            }
        }
    }
}

}  // namespace rb
