// rb_geom.cuh -- octree walk and surface intersection on the device.
//
// Restates the reference's recursive octant DDA and its intersectors:
//   localhit/raymove/checkhit   src/radiance/rt/raytrace.c:595-760
//   rayhit + rayreject          src/radiance/rt/raytrace.c:535-591
//   incube                      src/radiance/common/octree.c:115-126
//   o_face + inface             src/radiance/rt/o_face.c:16-62, common/face.c:121-162
//   o_sphere                    src/radiance/rt/sphere.c:16-83
//   o_cone + quadratic          src/radiance/rt/o_cone.c:17-146, common/zeroes.c:19-54
// as a persistent-thread, WARP-COOPERATIVE kernel body with shared-memory ray
// staging (walk_rays below).
//
// All geometry is IEEE double (RREAL); this translation unit is compiled with
// -fmad=false so that products and sums round exactly like the reference's
// non-fused x86-64 code.  Differences from the reference, by design:
//   * cube origins come from integer cell coordinates (one fused rounding)
//     instead of a chain of += cusize (differs by ulps of the cube origin);
//   * objects already tested in an earlier leaf are tested again instead of
//     being filtered through a per-ray checked set: rayreject() makes the
//     re-test a no-op (o == r->ro, or t > rot + FTINY, or the same pairwise
//     tie decision);
//   * every surface of a leaf first yields at most one CANDIDATE (t, facing,
//     inside) that does not depend on the ray's current best hit -- computed by
//     whichever lane of the warp is free -- and the ray's own lane then applies
//     rayreject() to the candidates in the reference's order (descending object
//     index).  Since inface()/the end-cap and radius checks do not depend on
//     the current hit, this gives exactly the sequential result.
#pragma once
#include "rb_device.cuh"

namespace rb {

#ifndef RB_OPR
#define RB_OPR 6                 // surfaces per ray per round handed to the warp
#endif
#ifndef RB_DITERS
#define RB_DITERS 32             // descend steps per round
#endif
#ifndef RB_FETCH_MIN
#define RB_FETCH_MIN 10          // refill a warp when this many lanes are idle (6 / 10 / 4: 1521 / 1548 / 1476 Mrays/s)
#endif
#ifndef RB_CSTEPS
#define RB_CSTEPS 1              // cubes a lane may step through per round while it lands in empty leaves
#endif
#ifndef RB_STEP_BRANCHLESS
#define RB_STEP_BRANCHLESS 1
#endif
#ifndef RB_SPHERE_INLINE
#define RB_SPHERE_INLINE 1
#endif
#ifndef RB_WALK_STATS
#define RB_WALK_STATS 0          // count node / leaf-entry / surface-test visits (developer builds)
#endif
#if RB_WALK_STATS
#define RB_STAT(x) x
#else
#define RB_STAT(x)
#endif
#ifndef RB_PAIR_CAP
#define RB_PAIR_CAP 192          // (ray, surface) pairs a warp takes per pass; what does not fit waits for the next pass
#endif
#define RB_PAIRS (32 * RB_OPR < RB_PAIR_CAP ? 32 * RB_OPR : RB_PAIR_CAP)
#ifndef RB_SLOW_MIN
#define RB_SLOW_MIN 0            // lanes in curved-surface leaves that gather before their round runs (0: off; measured slower)
#endif
#ifndef RB_CURVED_PASS
#define RB_CURVED_PASS 0         // (measured: 2 % SLOWER -- the kernel is bound by the length of a round's dependent chain, which
                                 // a second loop lengthens, not by the instructions the grouping saves)
                                 // 1: the pair loop tests polygons only and lists the other pairs; one more loop then runs the
                                 // in-line sphere / cylinder tests for all of them together (instead of in every iteration of
                                 // the pair loop, for the two or three lanes that happen to hold one)
#endif
#ifndef RB_WAIT_MIN
#define RB_WAIT_MIN 0            // (measured: 4 / 6 / 8 lanes, 8 / 16 rounds all 4-5 % SLOWER than 0) > 0: a ray whose leaf holds a curved surface that SURVIVES the in-line miss tests waits in that
                                 // leaf (WF_WAIT) until this many lanes of its warp wait, then the leaf is redone in a round that
                                 // runs the exact out-of-line pass for all of them (see walk_rays)
#endif
#ifndef RB_WAIT_ROUNDS
#define RB_WAIT_ROUNDS 8         // ... or until this many rounds have passed (power of two)
#endif
#ifndef RB_PREFETCH
#define RB_PREFETCH 0            // bit 0: prefetch the next leaf's set entries at the end of the step; bit 1: the next node's words
#endif
#ifndef RB_CONE_BSPHERE
#define RB_CONE_BSPHERE 0        // 1: cone-family pairs are first tested against the bounding sphere the loader left in the record (measured: no gain)
#endif
#ifndef RB_STEP_RCP
#define RB_STEP_RCP 0            // 1: the step to the next cube multiplies by 1/dir (kept per ray) instead of dividing
#endif
#ifndef RB_INTWALK
#define RB_INTWALK 1             // 1: integer cell coordinates + top-level cell table (see walk_rays); 0: the round-1 walk
#endif
#if RB_INTWALK
#undef RB_STEP_RCP
#define RB_STEP_RCP 1            // the integer walk steps with the ray's reciprocal direction
#endif
#ifndef RB_PAIR_ILP
#define RB_PAIR_ILP 1            // (ray, surface) pairs a lane has in flight in the pair loop
#endif

__device__ __forceinline__ double dot3(const double a[3], const double b[3]) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
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
__device__ __forceinline__ int quadratic(double& r0, double& r1, double a, double b, double c) {
    int first;
    r0 = r1 = 0.0;
    if (a < -RB_FTINY) first = 1;
    else if (a > RB_FTINY) first = 0;
    else if (fabs(b) > RB_FTINY) { r0 = -c / b; return 1; }
    else return 0;
    b *= 0.5;
    double disc = b * b - a * c;
    if (disc < -RB_FTINY * RB_FTINY) return 0;
    if (disc <= RB_FTINY * RB_FTINY) { r0 = -b / a; return 1; }
    disc = sqrt(disc);
    const double lo = (-b - disc) / a, hi = (-b + disc) / a;
    if (first) { r1 = lo; r0 = hi; } else { r0 = lo; r1 = hi; }
    return 2;
}

// Candidate of a polygon: o_face() up to, but not including, rayreject().
// `tmax` is a conservative upper bound (current rot + a few FTINY).  `hot` =
// kind | projection axis << 4 | exact-rectangle flag << 6 (from the leaf entry);
// n01/n2o = the plane, box = the polygon's 2-D bounds as floats rounded outward.
__device__ __forceinline__ bool cand_face(int hot, const double* __restrict__ g, double2 n01, double2 n2o, float4 box,
                                          const double org[3], const double dir[3], double tmax, double& t,
                                          bool& front) {
    double rdot = -(dir[0] * n01.x + dir[1] * n01.y + dir[2] * n2o.x);
    if ((rdot <= RB_FTINY) & (rdot >= -RB_FTINY)) return false;
    t = ((org[0] * n01.x + org[1] * n01.y + org[2] * n2o.x) - n2o.y) / rdot;
    if ((t <= RB_FTINY) | (t > tmax)) return false;
    front = rdot > 0;
    int ax = (hot >> 4) & 3;
    double p0 = org[0] + t * dir[0], p1 = org[1] + t * dir[1], p2 = org[2] + t * dir[2];
    double x = ax == 0 ? p1 : ax == 1 ? p2 : p0;      // xi = (ax+1)%3
    double y = ax == 0 ? p2 : ax == 1 ? p0 : p1;      // yi = (ax+2)%3
    // 2-D bounding box first: outside the (outward-rounded, hence larger) box by
    // more than FTINY can never be "in" (no edge straddles y / all straddling
    // edges on one side, and none of inface()'s three FABSEQ cases can fire).
    const double lx = box.x, hx = box.y, ly = box.z, hy = box.w;
    if ((x < lx - RB_FTINY) | (x > hx + RB_FTINY) | (y < ly - RB_FTINY) | (y > hy + RB_FTINY)) return false;
    // Well inside an exact axis-aligned rectangle is always "in".  The float
    // bounds are within one float ulp (<= 1.2e-7 |v|) of the exact ones, so the
    // margin below implies "inside the exact rectangle by more than FTINY".
    // Only the border zone runs the edge loop.
    if (hot & 0x40) {
        const double ulp = 1.2e-7;
        if ((x > lx + (RB_FTINY + 1e-30 + ulp * fabs(lx))) & (x < hx - (RB_FTINY + 1e-30 + ulp * fabs(hx))) &
            (y > ly + (RB_FTINY + 1e-30 + ulp * fabs(ly))) & (y < hy - (RB_FTINY + 1e-30 + ulp * fabs(hy))))
            return true;
    }
    const int nv = (__ldg(reinterpret_cast<const int*>(g - 2)) >> 16) & 0xffff;
    return inface2d(g + 6, nv, x, y);
}

// Candidate of a sphere / cone-family surface (rare kinds, kept out of line).
// Returns the single root the reference could accept: the first root > FTINY
// (sphere.c:60-66) / the first root > FTINY within the end caps (o_cone.c:98-110)
// / the ring's plane hit within its radii (o_cone.c:71-83); as +t for a
// front-facing candidate, -t for a back-facing one, 0 for none.  Everything
// travels by value: reference arguments of a __noinline__ function live in
// local memory, which showed up as 200 M L1 sectors per launch.
__device__ __noinline__ double cand_other(int kind, const double* __restrict__ g, double ox, double oy, double oz,
                                          double dx, double dy, double dz, double tmax) {
    const double org[3] = {ox, oy, oz}, dir[3] = {dx, dy, dz};
    if (kind == PK_SPHERE || kind == PK_BUBBLE) {
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            a += dir[i] * dir[i];
            double d = org[i] - g[i];
            b += 2.0 * dir[i] * d;
            c += d * d;
        }
        c -= g[3] * g[3];
        double r0, r1;
        const int nroots = quadratic(r0, r1, a, b, c);
        double t; int i;
        if (nroots >= 1 && r0 > RB_FTINY) { t = r0; i = 0; }
        else if (nroots >= 2 && r1 > RB_FTINY) { t = r1; i = 1; }
        else return 0.0;
        if (t > tmax) return 0.0;
        const bool front = (1 - 2 * ((i > 0) ^ (kind == PK_BUBBLE))) > 0;
        return front ? t : -t;
    }
    if (kind >= PK_CONE && kind <= PK_RING) {
        const double* ad = g; double al = g[3];
        const double* p0 = g + 4;
        double r0 = g[8], r1 = g[9];
        const double* tm = g + 12;       // tm[i][j] at tm[i*3+j], i = 0..3
        double rox[3], rdx[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            rdx[j] = dir[0] * tm[0 + j] + dir[1] * tm[3 + j] + dir[2] * tm[6 + j];
            rox[j] = org[0] * tm[0 + j] + org[1] * tm[3 + j] + org[2] * tm[6 + j];
            rox[j] += tm[9 + j];
        }
        double a, b, c;
        if (kind == PK_CONE || kind == PK_CUP) {
            a = rdx[0] * rdx[0] + rdx[1] * rdx[1] - rdx[2] * rdx[2];
            b = 2.0 * (rdx[0] * rox[0] + rdx[1] * rox[1] - rdx[2] * rox[2]);
            c = rox[0] * rox[0] + rox[1] * rox[1] - rox[2] * rox[2];
        } else if (kind == PK_CYL || kind == PK_TUBE) {
            a = rdx[0] * rdx[0] + rdx[1] * rdx[1];
            b = 2.0 * (rdx[0] * rox[0] + rdx[1] * rox[1]);
            c = rox[0] * rox[0] + rox[1] * rox[1] - r0 * r0;
        } else {  // ring
            if ((rdx[2] <= RB_FTINY) & (rdx[2] >= -RB_FTINY)) return 0.0;
            const double t = -rox[2] / rdx[2];
            if ((t <= RB_FTINY) | (t > tmax)) return 0.0;
            b = t * rdx[0] + rox[0];
            c = t * rdx[1] + rox[1];
            a = b * b + c * c;
            if (a > r1 * r1 || a < r0 * r0) return 0.0;
            return (-rdx[2] > 0) ? t : -t;
        }
        double q0, q1;
        const int nroots = quadratic(q0, q1, a, b, c);
#pragma unroll
        for (int rn = 0; rn < 2; rn++) {
            if (rn >= nroots) break;
            const double root = rn ? q1 : q0;
            if (root <= RB_FTINY) continue;
            if (root > tmax) break;
            double dx3[3];
#pragma unroll
            for (int k = 0; k < 3; k++) dx3[k] = (org[k] + root * dir[k]) - p0[k];
            b = dot3(dx3, ad);
            if (b < 0.0) continue;
            if (b > al) continue;
            const bool front = (1 - 2 * ((rn > 0) ^ ((kind == PK_CUP) | (kind == PK_TUBE)))) > 0;
            return front ? root : -root;
        }
    }
    return 0.0;
}

// Hit point, surface normal and rod = -rdir.ron of an accepted hit, computed
// once per ray by the shading kernel with the reference's expressions
// (o_face.c:52-56, sphere.c:72-78, o_cone.c:84-138).
__device__ __forceinline__ void hit_frame(const DScene& S, int robj, double rot, const double org[3],
                                       const double dir[3], double rop[3], double ron[3], double& rod) {
    for (int k = 0; k < 3; k++) rop[k] = org[k] + rot * dir[k];
    int4 hd = __ldg(&S.objhdr[robj]);
    int kind = hd.x & 0xff;
    const double* g = S.geom + hd.w;
    if (kind == PK_FACE) {
        ron[0] = g[0]; ron[1] = g[1]; ron[2] = g[2];
        rod = -(dir[0] * ron[0] + dir[1] * ron[1] + dir[2] * ron[2]);
        return;
    }
    if (kind == PK_SPHERE || kind == PK_BUBBLE) {
        double ar = g[3] * (1 - 2 * (kind == PK_BUBBLE));
        for (int k = 0; k < 3; k++) {
            rop[k] = org[k] + dir[k] * rot;
            ron[k] = (rop[k] - g[k]) / ar;
        }
        rod = -dot3(dir, ron);
        return;
    }
    if (kind == PK_RING) {
        const double* tm = g + 12;
        ron[0] = g[0]; ron[1] = g[1]; ron[2] = g[2];
        rod = -(dir[0] * tm[2] + dir[1] * tm[5] + dir[2] * tm[8]);      // -rdx[2]
        return;
    }
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
    rod = -dot3(dir, ron);
}

struct WalkStats { unsigned nodes, leafents, prims; };
#if RB_WALK_STATS
__device__ unsigned long long g_dbg[16];     // developer counters: cylinder pairs, after the in-line miss test, candidates; sphere pairs, candidates
#define RB_DBG(i) atomicAdd(&g_dbg[i], 1ULL)
#define RB_DBG2(i, v) atomicAdd(&g_dbg[i], (unsigned long long)(v))
#else
#define RB_DBG(i)
#define RB_DBG2(i, v)
#endif

// sourcehit() lives in rb_shade.cuh; declared here for the retire step
__device__ __forceinline__ int sourcehit(const DScene& S, const double dir[3], int rsrc, int crtype);

struct TraceIO {
    const QRay* __restrict__ qin;
    unsigned nin;
    HitRec* __restrict__ hits;
    unsigned* next;              // global fetch counter
    int anyhit;                  // 1: shadow rays towards distant sources end at ANY surface that is opaque to shadow rays
};

// Shared memory of one CTA of k_trace (per-thread columns, SoA).  Everything a
// ray owns between rounds lives here, not in registers: the kernel is bound by
// dependent-load latency, throughput scales with the warps resident per SM
// (profiles/r2_notes.md), and registers are what limits them.  The ancestor
// stack [depth + 1][NT] is dynamic shared memory sized by the octree's depth.
template <int NT>
struct WalkSmem {
    double ray[6][NT];           // staged ray: origin, direction
#if RB_STEP_RCP
    double inv[3][NT];           // 1 / direction (1 where the ray does not move along the axis)
#endif
    double pos[3][NT];           // current position along the ray (raymove's pos)
    double rot[NT];              // current best distance
    unsigned cell[3][NT];        // integer walk: the position as depth-bit integers (cube at level L = cell >> (depth - L));
                                 // round-1 walk: integer coordinates of the current cube at its level
    int robj[NT];                // current best object << 1 | front-facing, or -1
    unsigned ridx[NT];           // queue slot of the ray
    int lvl[NT];                 // level of the current cube | direction flags << 8
    unsigned pair[NT / 32][RB_PAIRS];   // (ray, surface) pairs of this round: leaf-set entry << 5 | owner lane
    int ndef[NT / 32];           // deferred (non-polygon) pairs of this round
    unsigned short defer[NT / 32][RB_PAIRS];   // their pair indices
#if RB_CURVED_PASS
    int ncur[NT / 32];           // non-polygon pairs of this pass, waiting for the curved-surface loop
    unsigned short cur[NT / 32][RB_PAIRS];
#endif
    double ct[NT / 32][RB_PAIRS];   // candidate distance per (ray, surface) pair (written for real candidates only)
    int cid[NT / 32][RB_PAIRS];     // candidate: object id << 1 | front
    unsigned cmask[NT];             // per owner: bit j set = its j-th pair of this pass produced a candidate
    unsigned char e0[NT];           // per owner: index of its first pair in the pass
};
static_assert(RB_OPR <= 7 && RB_PAIRS <= 255, "pair bookkeeping: 3-bit counts, byte offsets");

// 1 / x to double precision without the IEEE division sequence: the hardware's 20-bit estimate and two Newton
// steps.  Only used where the last bit does not matter (the walker's plane distances).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// One (ray, surface) pair: polygon candidates are computed here, the rare
// other kinds are queued for the warp's second pass.
#define RB_ENT_ID(e) ((e) & 0x1ffffff)
#define RB_ENT_HOT(e) ((int)((unsigned)(e) >> 25))
// PART: 0 = every kind (one loop), 1 = polygons, the rest is listed for the curved-surface loop, 2 = that loop
template <int NT, int PART = 0>
__device__ __forceinline__ void pair_test(const DScene& S, WalkSmem<NT>& sm, unsigned wid, int p, unsigned own,
                                          int2 ent, const double* __restrict__ g, double2 n01, double2 n2o,
                                          float4 box, unsigned* errflag, unsigned* errobj, bool exact = true) {
    // `exact` false (RB_WAIT_MIN): the round does not run the out-of-line pass; a pair that needs it marks its owner
    // (bit 31 of the candidate mask), who then waits in the leaf and redoes it in a round that does
#define RB_DEFER(p_) do { if (exact) sm.defer[wid][atomicAdd(&sm.ndef[wid], 1)] = (unsigned short)(p_); \
                          else atomicOr(&sm.cmask[own], 0x80000000u); } while (0)
    const int hot = RB_ENT_HOT(ent.x);
    const int kind = hot & 0xf;
#if RB_CURVED_PASS
    if (PART == 1 && kind != PK_FACE) {
        if (kind != PK_NONE) sm.cur[wid][atomicAdd(&sm.ncur[wid], 1)] = (unsigned short)p;
        return;
    }
#endif
    if (PART != 2 && kind == PK_FACE) {
        const double org[3] = {sm.ray[0][own], sm.ray[1][own], sm.ray[2][own]};
        const double rd[3] = {sm.ray[3][own], sm.ray[4][own], sm.ray[5][own]};
        const double tmax = sm.rot[own] + 8 * RB_FTINY;      // ties may raise rot by < FTINY each
        double t = 0;
        bool fr = true;
        if (cand_face(hot, g, n01, n2o, box, org, rd, tmax, t, fr)) {
            sm.ct[wid][p] = t;
            sm.cid[wid][p] = (RB_ENT_ID(ent.x) << 1) | (int)fr;
            atomicOr(&sm.cmask[own], 1u << (p - sm.e0[own]));
        }
#if RB_SPHERE_INLINE
    } else if ((kind == PK_SPHERE) | (kind == PK_BUBBLE)) {
        // o_sphere (sphere.c:16-83) on the record words the pair loop already holds (centre, radius):
        // short enough to stay in line, which keeps the out-of-line second pass for the cone family
        const double org[3] = {sm.ray[0][own], sm.ray[1][own], sm.ray[2][own]};
        const double rd[3] = {sm.ray[3][own], sm.ray[4][own], sm.ray[5][own]};
        const double ctr[3] = {n01.x, n01.y, n2o.x};
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            a += rd[i] * rd[i];
            const double d = org[i] - ctr[i];
            b += 2.0 * rd[i] * d;
            c += d * d;
        }
        c -= n2o.y * n2o.y;
#if RB_SPHERE_INLINE == 2
        // only the discriminant in line (the same expression quadratic() forms, with its own tolerance widened a
        // thousandfold): a ray that cannot have a root is done here; the square root and the two divisions run in
        // the out-of-line pass, together with the cone family's, so that one pass of heavy code serves both kinds
        {
            const double hb = b * 0.5, disc = hb * hb - a * c;
            if (disc < -1e-9 * (1.0 + hb * hb)) return;
            // both roots behind the origin (the centre lies behind and the origin is outside): no candidate either
            if ((hb > 0.0) & (c > 1e-9 * (1.0 + c))) return;
            RB_DEFER(p);
            return;
        }
#endif
        double r0, r1;
        const int nroots = quadratic(r0, r1, a, b, c);
        RB_DBG(3); if (nroots) RB_DBG(4);
        double t = 0.0; int i = 0;
        if (nroots >= 1 && r0 > RB_FTINY) { t = r0; i = 0; }
        else if (nroots >= 2 && r1 > RB_FTINY) { t = r1; i = 1; }
        if ((t > 0.0) & (t <= sm.rot[own] + 8 * RB_FTINY)) {
            const bool fr = !((i > 0) ^ (kind == PK_BUBBLE));
            sm.ct[wid][p] = t;
            sm.cid[wid][p] = (RB_ENT_ID(ent.x) << 1) | (int)fr;
            atomicOr(&sm.cmask[own], 1u << (p - sm.e0[own]));
        }
#endif
    } else {                         // rare kinds: second pass, again with all lanes
        if (kind == PK_UNSUPPORTED) { atomicOr(errflag, RB_ERR_UNSUP_PRIM); *errobj = (unsigned)RB_ENT_ID(ent.x); }
        else if (kind != PK_NONE) {
            if ((kind == PK_CYL) | (kind == PK_TUBE)) {
                RB_DBG(0);
                // Most rays that cross a leaf holding a thin cylinder pass it by, and the out-of-line pass below costs
                // ~400 instructions at 2-3 live lanes.  A ray misses the INFINITE cylinder iff its distance from the axis
                // exceeds the radius: with n = dir x axis and e = org - p0, (e.n)^2 > |n|^2 r^2 -- which is o_cone()'s own
                // discriminant, -(b^2 - a c) > 0, written in world space (the cylinder's transform is a rotation and a
                // translation).  Decided here only when the margin is a million times the rounding error of either
                // form; everything closer, and every ray parallel to the axis, goes to the exact pass.
                const double2* g2 = reinterpret_cast<const double2*>(g);
                const double2 p01 = __ldg(&g2[2]), p2s = __ldg(&g2[3]), r01 = __ldg(&g2[4]);
                const double dx = sm.ray[3][own], dy = sm.ray[4][own], dz = sm.ray[5][own];
                const double ex = sm.ray[0][own] - p01.x, ey = sm.ray[1][own] - p01.y, ez = sm.ray[2][own] - p2s.x;
                const double ax_ = n01.x, ay_ = n01.y, az_ = n2o.x;
                const double nx = dy * az_ - dz * ay_, ny = dz * ax_ - dx * az_, nz = dx * ay_ - dy * ax_;
                const double nn = nx * nx + ny * ny + nz * nz;
                const double tp = ex * nx + ey * ny + ez * nz;
                const double ee = ex * ex + ey * ey + ez * ez;
                if (tp * tp > nn * (r01.x * r01.x) * (1.0 + 1e-6) + 1e-9 * (ee * nn) + 1e-12) return;
                RB_DBG(1);
                // ... and the ends: both roots lie within r / |n| of the ray's closest approach to the axis (t_c), so
                // their heights along the axis lie within |d.axis| r / |n| of the height there.  With everything scaled
                // by |n|^2 = nn (and sqrt(nn) <= 1 bounding the half width from above): a ray whose whole root interval
                // is below the base or above the top (o_cone()'s `0 <= b <= al` test) misses.
                {
                    const double al = n2o.y;                                    // g[3]: length of the axis
                    const double da = dx * ax_ + dy * ay_ + dz * az_;           // dir . axis
                    const double ea = ex * ax_ + ey * ay_ + ez * az_;           // height of the origin above the base
                    const double ed = ex * dx + ey * dy + ez * dz;
                    const double zc = ea * nn - (ed - ea * da) * da;            // height at t_c, times nn
                    const double hw = fabs(da) * r01.x * (1.0 + 1e-6) + 1e-9 * (ee + al + 1.0) + 1e-12;
                    if ((zc + hw < 0.0) | (zc - hw > al * nn)) return;
                }
                RB_DBG(5);
            }
#if RB_CONE_BSPHERE
            // the loader left a bounding sphere (4 floats, rounded outward) in the record's spare words: a ray
            // that passes it by, points away from it, or ends before it cannot have a candidate there
            const float4 bs = __ldg(reinterpret_cast<const float4*>(g + 10));
            const double ox = (double)bs.x - sm.ray[0][own], oy = (double)bs.y - sm.ray[1][own], oz = (double)bs.z - sm.ray[2][own];
            const double dx = sm.ray[3][own], dy = sm.ray[4][own], dz = sm.ray[5][own];
            const double R = (double)bs.w, oo = ox * ox + oy * oy + oz * oz, b = ox * dx + oy * dy + oz * dz;
            const double dd = dx * dx + dy * dy + dz * dz;
            const bool miss = (oo - b * b / dd > R * R * (1.0 + 1e-9)) | ((b < 0.0) & (oo > R * R)) |
                              (b - R * 1.000001 > (sm.rot[own] + 8 * RB_FTINY) * 1.000001);
            if (!miss)
#endif
            RB_DEFER(p);
        }
    }
#undef RB_DEFER
}

// cusize * 2^-L, exactly what L halvings of the root cube size give
__device__ __forceinline__ double cube_size(double cs, int L) {
    return cs * __hiloint2double((1023 - L) << 20, 0);
}

// walk_rays(): persistent-thread, warp-cooperative localhit().
//  * Every warp keeps pulling rays from the queue: a lane whose ray is finished
//    retires it (HitRec; sourcehit() for misses) and, when RB_FETCH_MIN lanes
//    are idle, the warp reserves that many queue slots with one atomicAdd.
//    Ray lengths are heavy-tailed, so without the refill a warp ran at < 5 live
//    lanes on average (profiles/r1_notes.md).
//  * Phase A (per lane): descend to the leaf that holds the ray's position
//    (raymove, raytrace.c:668-687); 1-3 cheap iterations.
//  * Phase B (whole warp): the surfaces waiting in the lanes' leaves form
//    (ray, surface) pairs; the 32 lanes take pairs round-robin, read the ray
//    from shared memory and compute the pair's candidate.  Leaves hold 0..6
//    surfaces and only ~1 lane in 6 stands in a full leaf at a time, so testing
//    them lane-by-lane left ~10 of 32 lanes busy; pairs keep all lanes busy.
//    The owner then applies rayreject() to its candidates in descending index
//    order.  Spheres / cones (rare) are deferred to a second cooperative pass.
//  * Phase C (per lane, all lanes together): checkhit()'s "hit point in this
//    cube" test (raytrace.c:756-759), else the step to the neighbour cube
//    (raytrace.c:688-738).  Every live lane does exactly one step per round,
//    so the expensive step code (3 divisions) runs with the warp nearly full;
//    in the first version it ran with ~5 of 32 lanes.
// Registers carry only what a phase is working on: between phases a lane's
// state is {flags, node word, level} plus the shared-memory columns above.
// MUST be called by all threads of the CTA.
enum : unsigned { WF_HAVE = 1, WF_DONE = 2, WF_RESULT = 4, WF_AFT = 8, WF_EXHAUSTED = 16, WF_WAIT = 32, WF_ANYHIT = 64 };

template <int NT>
__device__ __forceinline__ void walk_rays(const DScene& S, const TraceIO io, WalkSmem<NT>& sm, int* __restrict__ stk,
                                          WalkStats& ws, unsigned* errflag, unsigned* errobj) {
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const double cs = S.cusize;
    const int2* __restrict__ pool = reinterpret_cast<const int2*>(S.leafpool);
    unsigned fl = WF_DONE;       // WF_* flags of this lane's ray
#if RB_SLOW_MIN > 0 || RB_WAIT_MIN > 0
    unsigned round = 0;
#endif
    // cube the ray stands in: set by the refill or by phase C, consumed by phase A
    // (registers only between those two; parked in shared memory across phase B)
    unsigned ix = 0, iy = 0, iz = 0;
    int w = -1, Ld = 0;          // node word; level | direction flags << 8
    double px = 0, py = 0, pz = 0;
    for (;;) {
        // ---- retire finished rays ----
        if ((fl & (WF_HAVE | WF_DONE)) == (WF_HAVE | WF_DONE)) {
            const unsigned ridx = sm.ridx[tid];
            const int ro = sm.robj[tid];
            HitRec o;
            o.rot = sm.rot[tid]; o.rod = (ro & 1) ? 1.0 : -1.0; o.robj = ro >> 1; o.local = 1;
            if (!(fl & WF_RESULT)) {
                o.rot = RB_FHUGE; o.rod = 1.0; o.robj = -1; o.local = 0;
                const QRay& q = io.qin[ridx];
                if (!(q.rmax > RB_FTINY)) {          // aft-clipped rays never see sources
                    const double dir[3] = {sm.ray[3][tid], sm.ray[4][tid], sm.ray[5][tid]};
                    int sn = sourcehit(S, dir, q.rsrc, q.info & 0x3ff);
                    if (sn >= 0) o.robj = S.srcs[sn].so;
                }
            }
            io.hits[ridx] = o;
            fl &= ~WF_HAVE;
        }
        // ---- refill idle lanes ----
        const unsigned idle = __ballot_sync(FULL, !(fl & WF_HAVE));
        if (idle == FULL && (fl & WF_EXHAUSTED)) break;
        if (!(fl & WF_EXHAUSTED) && (__popc(idle) >= RB_FETCH_MIN)) {
            const int n = __popc(idle);
            const int leader = __ffs(idle) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(io.next, (unsigned)n);
            base = __shfl_sync(FULL, base, leader);
            const bool last = base + n >= io.nin;
            const unsigned my = base + __popc(idle & ((1u << lane) - 1));
            if (!(fl & WF_HAVE) && my < io.nin) {
                const double2* q2 = reinterpret_cast<const double2*>(&io.qin[my]);
                const double2 a = __ldg(&q2[0]), b = __ldg(&q2[1]), c = __ldg(&q2[2]), d = __ldg(&q2[3]);
                const double org[3] = {a.x, a.y, b.x};
                const double dir[3] = {b.y, c.x, c.y};
                const double rmax = d.x;
                sm.ray[0][tid] = org[0]; sm.ray[1][tid] = org[1]; sm.ray[2][tid] = org[2];
                sm.ray[3][tid] = dir[0]; sm.ray[4][tid] = dir[1]; sm.ray[5][tid] = dir[2];
#if RB_STEP_RCP
#pragma unroll
                for (int i = 0; i < 3; i++) sm.inv[i][tid] = (fabs(dir[i]) > 1e-7) ? fast_rcp(dir[i]) : 1.0;
#endif
                sm.ridx[tid] = my;
                // ---- localhit() prologue (raytrace.c:604-651) ----
                int dirf = 0;
                double pos[3];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    pos[i] = org[i];
                    if (dir[i] > 1e-7) dirf |= 1 << i;
                    else if (dir[i] < -1e-7) dirf |= 0x10 << i;
                }
                double rot = RB_FHUGE;
                bool done = !dirf;
                fl = WF_HAVE;          // (queue-exhausted bit is set again below)
                if (!done && rmax > RB_FTINY) { fl |= WF_AFT; rot = rmax; }
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
                        if (t >= rot) done = true;
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
                if (done) fl |= WF_DONE;
                // A shadow ray towards a DISTANT source adds nothing as soon as any surface stands in its way whose material
                // returns at once for shadow rays (plastic, metal and the anisotropic kinds: normal.c:190-191,
                // aniso.c:203-204): which of several blockers is the nearest does not matter -- a transparent surface in
                // front only scales what is already zero, and trace_contrib() counts a shadow ray on its source alone
                // (rcontrib.c:283-286).  Such a ray ends at the first candidate of that kind instead of walking on to the
                // leaf that holds the hit point.  (Local sources have an aft plane and keep the nearest-hit rule.)
                if (io.anyhit && !done && !(fl & WF_AFT)) {
                    const int4 t4 = __ldg(reinterpret_cast<const int4*>(&io.qin[my]) + 4);      // bytes 64..79: coef[2], rweight, row, info
                    if (t4.w & RT_SHADOW) {
                        const int rsrc = __ldg(&io.qin[my].rsrc);
                        if (rsrc >= 0 && (S.srcs[rsrc].flags & SF_DISTANT)) fl |= WF_ANYHIT;
                    }
                }
                px = pos[0]; py = pos[1]; pz = pos[2];
                sm.pos[0][tid] = px; sm.pos[1][tid] = py; sm.pos[2][tid] = pz;
                sm.rot[tid] = rot;
                sm.robj[tid] = -1;
#if RB_INTWALK
                // integer position at full depth and the cube that holds it, from the level-K cell table
                ix = __double2uint_rd((px - S.cuorg[0]) * S.inv_cell);
                iy = __double2uint_rd((py - S.cuorg[1]) * S.inv_cell);
                iz = __double2uint_rd((pz - S.cuorg[2]) * S.inv_cell);
                const unsigned top = (1u << S.maxdepth) - 1u;
                ix = min(ix, top); iy = min(iy, top); iz = min(iz, top);
                if (done) { w = -1; Ld = dirf << 8; }
                else {
                    const int sh = S.maxdepth - S.topk;
                    const int2 e = __ldg(&S.top[(ix >> sh) | ((iy >> sh) << S.topk) | ((iz >> sh) << (2 * S.topk))]);
                    w = e.x; Ld = (dirf << 8) | e.y;
                    RB_STAT(ws.nodes++;)
                }
#else
                ix = iy = iz = 0;
                w = S.root; Ld = dirf << 8;
#endif
            }
            if (last) fl |= WF_EXHAUSTED;
        }
        __syncwarp();
        bool act = (fl & (WF_HAVE | WF_DONE)) == WF_HAVE;
#if RB_PREFETCH & 4
        // second level: the set entries asked for at the end of the last step name the records the pair
        // loop will read; ask for those too (the entries themselves are L1 hits by now)
        if (act && w < -1) {
            const unsigned u = (unsigned)(-w - 2);
            const int so = (int)(u >> 4), k = min((int)(u & 7), 6);
            for (int j = 1; j <= k; j++)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(S.geom + __ldg(&pool[so + j]).y));
        }
#endif
        // ---- phase A: descend towards a leaf (raymove, raytrace.c:668-687); at most
        //      RB_DITERS levels per round ----
#if RB_INTWALK
        {
            int L = Ld & 0xff;
            const int D = S.maxdepth, K = S.topk;
#pragma unroll 1
            for (int it = 0; it < RB_DITERS; it++) {
                const bool d = act & (w >= 0);
                if (!__any_sync(FULL, d)) break;
                if (d) {                 // the child is named by bit D-1-L of the three coordinates
                    stk[(L - K) * NT + tid] = w;
                    const int b = D - 1 - L;
                    const int br = ((ix >> b) & 1) | (((iy >> b) & 1) << 1) | (((iz >> b) & 1) << 2);
                    w = __ldg(&S.nodes[(size_t)w * 8 + br]);
                    RB_STAT(ws.nodes++;)
                    L++;
                }
            }
            Ld = (Ld & ~0xff) | L;
            sm.cell[0][tid] = ix; sm.cell[1][tid] = iy; sm.cell[2][tid] = iz;
            sm.lvl[tid] = Ld;
        }
#else
        {
            int L = Ld & 0xff;
            double size = cube_size(cs, L);
#pragma unroll 1
            for (int it = 0; it < RB_DITERS; it++) {
                const bool d = act & (w >= 0);
                if (!__any_sync(FULL, d)) break;
                if (d) {
                    stk[L * NT + tid] = w;
                    const double half = size * 0.5;
                    const double lox = fma((double)ix, size, S.cuorg[0]);
                    const double loy = fma((double)iy, size, S.cuorg[1]);
                    const double loz = fma((double)iz, size, S.cuorg[2]);
                    int br = 0;
                    ix <<= 1; iy <<= 1; iz <<= 1;
                    if (px >= lox + half) { br |= 1; ix |= 1; }
                    if (py >= loy + half) { br |= 2; iy |= 1; }
                    if (pz >= loz + half) { br |= 4; iz |= 1; }
                    w = __ldg(&S.nodes[(size_t)w * 8 + br]);
                    RB_STAT(ws.nodes++;)
                    size = half; L++;
                }
            }
            Ld = (Ld & ~0xff) | L;
            sm.cell[0][tid] = ix; sm.cell[1][tid] = iy; sm.cell[2][tid] = iz;
            sm.lvl[tid] = Ld;
        }
#endif
        act &= (w < 0);                          // still inside the tree: continue next round
        const bool full = act & (w < -1);
        int kleft = 0, setoff = 0;
        bool slow = false;
        if (full) {
            const unsigned u = (unsigned)(-w - 2);       // set offset << 4 | curved << 3 | min(count, 7)
            setoff = (int)(u >> 4);
            slow = (u >> 3) & 1;
            kleft = (int)(u & 7);
            if (kleft == 7) kleft = __ldg(&pool[setoff]).x;
        }
#if RB_SLOW_MIN > 0
        // Leaves holding a sphere / cone need the second (out-of-line, ~400 instruction) pass, which ran
        // with 3-4 live lanes when every round paid for it.  Such lanes now wait in their leaf until
        // RB_SLOW_MIN of them have gathered (or nothing else is runnable, or every 8th round), so the
        // pass runs less often and fuller.  Only the schedule changes, not what any ray computes.
        {
            const unsigned mslow = __ballot_sync(FULL, slow);
            const unsigned mfast = __ballot_sync(FULL, act & !slow);
            const bool go = (__popc(mslow) >= RB_SLOW_MIN) | (mfast == 0) | ((++round & 7) == 0);
            if (slow & !go) { act = false; kleft = 0; }
        }
#endif
#if RB_WAIT_MIN > 0
        // A ray that met a curved surface it could not dismiss in line waits in its leaf (WF_WAIT: no pairs, no step)
        // until RB_WAIT_MIN lanes of the warp wait, nothing else is runnable, or RB_WAIT_ROUNDS rounds have passed; that
        // round runs the exact out-of-line pass (~450 instructions) for all of them at once instead of once per
        // round for a lane or two.  The waiting ray redoes its whole leaf then (rayreject() makes re-tests no-ops,
        // as for surfaces that straddle leaves), so what a ray computes does not change, only when.
        bool exact;
        {
            const bool waiting = act && (fl & WF_WAIT);
            const unsigned mwait = __ballot_sync(FULL, waiting);
            const unsigned mrun = __ballot_sync(FULL, act && !(fl & WF_WAIT));
            exact = (__popc(mwait) >= RB_WAIT_MIN) | (mrun == 0) | ((++round & (RB_WAIT_ROUNDS - 1)) == 0);
            if (waiting & !exact) { act = false; kleft = 0; }
            if (exact) fl &= ~WF_WAIT;
        }
#else
        const bool exact = true;
#endif
        RB_STAT(if (act & full) { ws.leafents += kleft + 1; ws.prims += kleft; })
        // ---- phase B: the warp tests the leaves' surfaces as (ray, surface) pairs ----
        for (;;) {
            if (!__any_sync(FULL, kleft > 0)) break;
            const int m0 = min(kleft, RB_OPR);
            // exclusive prefix sum of the 3-bit counts: one ballot per bit instead of a five-step shuffle scan
            const unsigned b0 = __ballot_sync(FULL, m0 & 1), b1 = __ballot_sync(FULL, m0 & 2), b2 = __ballot_sync(FULL, m0 & 4);
            const unsigned lt = (1u << lane) - 1u;
            const int e0 = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
            const int total = min(__popc(b0) + 2 * __popc(b1) + 4 * __popc(b2), RB_PAIRS);
            if (total == 0) break;
            const unsigned wbase = wid * 32;
            const int m = min(m0, max(0, RB_PAIRS - e0));     // the pair table holds RB_PAIRS per pass
            for (int j = 0; j < m; j++)               // owners publish their pairs, descending set index
                sm.pair[wid][e0 + j] = ((unsigned)(setoff + kleft - j) << 5) | lane;
            sm.cmask[tid] = 0u;
            sm.e0[tid] = (unsigned char)min(e0, 255);
            if (lane == 0) { sm.ndef[wid] = 0;
#if RB_CURVED_PASS
                sm.ncur[wid] = 0;
#endif
            }
            __syncwarp();
#if RB_PAIR_ILP == 2
            // two pairs per lane and pass: both records are in flight together
            for (int p = lane; p < total; p += 64) {
                const int pb = p + 32;
                const bool two = pb < total;
                const unsigned prA = sm.pair[wid][p], prB = sm.pair[wid][two ? pb : p];
                const unsigned ownA = wbase + (prA & 31), ownB = wbase + (prB & 31);
                const int2 entA = __ldg(&pool[prA >> 5]);
                const int2 entB = __ldg(&pool[prB >> 5]);
                const double* gA = S.geom + entA.y;
                const double* gB = S.geom + entB.y;
                const double2* gA2 = reinterpret_cast<const double2*>(gA);
                const double2* gB2 = reinterpret_cast<const double2*>(gB);
                const double2 a0 = __ldg(&gA2[0]), a1 = __ldg(&gA2[1]);
                const float4 a2 = __ldg(reinterpret_cast<const float4*>(gA + 4));
                const double2 b0 = __ldg(&gB2[0]), b1 = __ldg(&gB2[1]);
                const float4 b2 = __ldg(reinterpret_cast<const float4*>(gB + 4));
                pair_test(S, sm, wid, p, ownA, entA, gA, a0, a1, a2, errflag, errobj);
                if (two) pair_test(S, sm, wid, pb, ownB, entB, gB, b0, b1, b2, errflag, errobj);
            }
#else
            for (int p = lane; p < total; p += 32) {
                const unsigned pr = sm.pair[wid][p];
                const unsigned own = wbase + (pr & 31);
                const int2 ent = __ldg(&pool[pr >> 5]);
                const double* g = S.geom + ent.y;
                // plane and 2-D box of the record in one go (one latency): 48 contiguous bytes
                const double2* g2 = reinterpret_cast<const double2*>(g);
                const double2 n01 = __ldg(&g2[0]), n2o = __ldg(&g2[1]);
                const float4 box = __ldg(reinterpret_cast<const float4*>(g + 4));
                pair_test<NT, RB_CURVED_PASS ? 1 : 0>(S, sm, wid, p, own, ent, g, n01, n2o, box, errflag, errobj, exact);
            }
#if RB_CURVED_PASS
            // the pairs that are not polygons, together: the in-line sphere test and the cylinder's miss tests run once
            // per pass for all of them, not once per iteration of the loop above for a lane or two
            __syncwarp();
            {
                const int ncur = sm.ncur[wid];
                for (int q = lane; q < ncur; q += 32) {
                    const int p = sm.cur[wid][q];
                    const unsigned pr = sm.pair[wid][p];
                    const unsigned own = wbase + (pr & 31);
                    const int2 ent = __ldg(&pool[pr >> 5]);
                    const double* g = S.geom + ent.y;
                    const double2* g2 = reinterpret_cast<const double2*>(g);
                    const double2 n01 = __ldg(&g2[0]), n2o = __ldg(&g2[1]);
                    pair_test<NT, 2>(S, sm, wid, p, own, ent, g, n01, n2o, make_float4(0.f, 0.f, 0.f, 0.f), errflag, errobj, exact);
                }
            }
#endif
#endif
            __syncwarp();
            const int ndef = sm.ndef[wid];
            for (int q = lane; q < ndef; q += 32) {
                const int p = sm.defer[wid][q];
                const unsigned pr = sm.pair[wid][p];
                const unsigned own = wbase + (pr & 31);
                const int2 ent = __ldg(&pool[pr >> 5]);
                const double* g = S.geom + ent.y;
                const int kind = RB_ENT_HOT(ent.x) & 0xf;
                const double t = cand_other(kind, g, sm.ray[0][own], sm.ray[1][own], sm.ray[2][own], sm.ray[3][own],
                                            sm.ray[4][own], sm.ray[5][own], sm.rot[own] + 8 * RB_FTINY);
                if (t != 0.0) {
                    RB_DBG(2);
                    sm.ct[wid][p] = fabs(t);
                    sm.cid[wid][p] = (RB_ENT_ID(ent.x) << 1) | (int)(t > 0.0);
                    atomicOr(&sm.cmask[own], 1u << (p - sm.e0[own]));
                }
            }
            __syncwarp();
            // owners apply rayreject() in the reference's order (raytrace.c:535-575)
            unsigned cm = m > 0 ? sm.cmask[tid] : 0u;
#if RB_WAIT_MIN > 0
            if (cm & 0x80000000u) {                   // a surviving curved pair in a round without the exact pass: wait here
                fl |= WF_WAIT; act = false; cm = 0u; kleft = m;       // (kleft = m: nothing more to publish this round)
            }
#endif
            if (cm) {
                double rot = sm.rot[tid];
                int ro = sm.robj[tid];
                for (; cm; cm &= cm - 1u) {           // candidates in pair order = descending object index
                    const int j = __ffs(cm) - 1;
                    const int c = sm.cid[wid][e0 + j];
                    const double t = sm.ct[wid][e0 + j];
                    if ((t <= RB_FTINY) | (t > rot + RB_FTINY)) continue;
                    if (!(t < rot - RB_FTINY)) {              // coincident point, so decide...
                        const int id = c >> 1, robj = ro >> 1;
                        if (id == robj) continue;
                        if (ro < 0) { if (fl & WF_AFT) continue; }
                        else {
                            const bool fr = c & 1, front = ro & 1;
                            const int4 hnew = __ldg(&S.objhdr[id]), hold = __ldg(&S.objhdr[robj]);
                            const int fnew = hnew.x >> 8, fold = hold.x >> 8;
                            const bool mnew = fnew & PF_HASMAT, mray = fold & PF_HASMAT;
                            bool rej = false, dec = false;
                            if (!mnew) { if (mray) { rej = true; dec = true; } }
                            else if (!mray) { dec = true; }
                            else if (fnew & PF_TRANSP) { if (!(fold & PF_TRANSP)) { rej = true; dec = true; } }
                            else if (fold & PF_TRANSP) { dec = true; }
                            if (!dec) {
                                if (!fr) { if (front) { rej = true; dec = true; } }
                                else if (!front) { dec = true; }
                            }
                            if (!dec) rej = hold.y >= hnew.y;  // later modifier definition wins tie
                            if (rej) continue;
                        }
                    }
                    ro = c; rot = t;
                }
                sm.rot[tid] = rot;
                sm.robj[tid] = ro;
                if ((fl & WF_ANYHIT) && ro >= 0) {
                    const int ms = __ldg(&S.objhdr[ro >> 1]).z;
                    if (ms >= 0) {
                        const MatRec& mr = S.mats[ms];
                        const int mk = mr.kind;
                        if (((mk == MK_PLASTIC) | (mk == MK_METAL) | (mk >= MK_PLASTIC2 && mk <= MK_TRANS2)) && mr.flags == 0) {
                            fl |= WF_DONE | WF_RESULT;        // blocked: retire with this hit
                            act = false; kleft = m;
                        }
                    }
                }
            }
            kleft -= m;
            __syncwarp();
        }
        // ---- phase C: accept (checkhit / aft plane), else step to the neighbour cube; a lane
        //      that lands in an EMPTY leaf has nothing to do in phases A and B, so it keeps
        //      stepping (up to RB_CSTEPS cubes per round) instead of idling through a round ----
        if (act) {
            ix = sm.cell[0][tid]; iy = sm.cell[1][tid]; iz = sm.cell[2][tid];
            Ld = sm.lvl[tid];
            int L = Ld & 0xff;
            const int dirf = Ld >> 8;
            const double dir[3] = {sm.ray[3][tid], sm.ray[4][tid], sm.ray[5][tid]};
            const int ro = sm.robj[tid];
            double pos[3] = {sm.pos[0][tid], sm.pos[1][tid], sm.pos[2][tid]};
            bool done = false, fullc = full;
#pragma unroll 1
            for (int cstep = 0;; cstep++) {
                const double size = cube_size(cs, L);
#if RB_INTWALK
                const int sh = S.maxdepth - L;
                const double lox = fma((double)(ix >> sh), size, S.cuorg[0]);
                const double loy = fma((double)(iy >> sh), size, S.cuorg[1]);
                const double loz = fma((double)(iz >> sh), size, S.cuorg[2]);
#else
                const double lox = fma((double)ix, size, S.cuorg[0]);
                const double loy = fma((double)iy, size, S.cuorg[1]);
                const double loz = fma((double)iz, size, S.cuorg[2]);
#endif
                const double hix = lox + size, hiy = loy + size, hiz = loz + size;
                if (fullc ? (ro >= 0) : ((fl & WF_AFT) && ro < 0)) {
                    // checkhit (raytrace.c:756-759) / aft-plane point in an empty leaf (:709-710)
                    const double rot = sm.rot[tid];
                    const double hx = sm.ray[0][tid] + rot * dir[0];
                    const double hy = sm.ray[1][tid] + rot * dir[1];
                    const double hz = sm.ray[2][tid] + rot * dir[2];
                    if (!(lox > hx || hx >= hix || loy > hy || hy >= hiy || loz > hz || hz >= hiz)) {
                        done = true;
                        if (fullc) fl |= WF_RESULT;
                        break;
                    }
                }
                // advance to next cube (raytrace.c:712-738)
                RB_DBG(8); if (fullc) RB_DBG(9); RB_DBG2(10, L);
#if RB_STEP_RCP
                // raymove() walks the ray with t = (plane - pos) / dir; here the quotient is a product with the
                // ray's reciprocal direction.  pos may differ from the reference's in the last bit, which only
                // decides the cube when pos already sits within an ulp of a cube face; hit distances do not
                // come from pos (they are computed from the ray's origin), so results stay what they were.
                const double tx = (((dirf & 1) ? hix : lox) - pos[0]) * sm.inv[0][tid];
                const double ty = (((dirf & 2) ? hiy : loy) - pos[1]) * sm.inv[1][tid];
                const double tz = (((dirf & 4) ? hiz : loz) - pos[2]) * sm.inv[2][tid];
                int ax = 0;
                double t = (dirf & 0x11) ? tx : RB_FHUGE;
                if ((dirf & 0x22) && ty < t) { t = ty; ax = 1; }
                if ((dirf & 0x44) && tz < t) { t = tz; ax = 2; }
#elif RB_STEP_BRANCHLESS
                // the three plane distances are independent: computed unconditionally (with a harmless
                // denominator on an axis the ray does not move along) so their division sequences interleave
                const double tx = (((dirf & 1) ? hix : lox) - pos[0]) / ((dirf & 0x11) ? dir[0] : 1.0);
                const double ty = (((dirf & 2) ? hiy : loy) - pos[1]) / ((dirf & 0x22) ? dir[1] : 1.0);
                const double tz = (((dirf & 4) ? hiz : loz) - pos[2]) / ((dirf & 0x44) ? dir[2] : 1.0);
                int ax = 0;
                double t = (dirf & 0x11) ? tx : RB_FHUGE;
                if ((dirf & 0x22) && ty < t) { t = ty; ax = 1; }
                if ((dirf & 0x44) && tz < t) { t = tz; ax = 2; }
#else
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
#endif
                pos[0] = pos[0] + dir[0] * t; pos[1] = pos[1] + dir[1] * t; pos[2] = pos[2] + dir[2] * t;
#if RB_INTWALK
                // step to the neighbour (raytrace.c:688-706) on the integers: along the exit axis the position becomes
                // the first / last depth-level cell of the neighbour cube (the carry of the increment IS the climb: the
                // highest bit that changes names the common ancestor); along the other two axes it is the cell the new
                // position falls in, kept inside the cube just left (the position can only be an ulp outside it)
                {
                    const int D = S.maxdepth, K = S.topk;
                    const bool positive = dirf & (1 << ax);
                    const unsigned ia = ax == 0 ? ix : ax == 1 ? iy : iz;
                    unsigned c = ia >> sh, ipn;
                    bool left;
                    if (positive) { c++; left = (c >> L) != 0; ipn = c << sh; }
                    else { left = c == 0; ipn = (c << sh) - 1u; }
                    if (left) { done = true; if (ro >= 0) fl |= WF_RESULT; break; }    // left the scene cube
                    const int La = D - 1 - (31 - __clz((int)(ipn ^ ia)));             // level of the common ancestor
                    const unsigned cm = (1u << sh) - 1u;
                    const unsigned qx = __double2uint_rd((pos[0] - S.cuorg[0]) * S.inv_cell);
                    const unsigned qy = __double2uint_rd((pos[1] - S.cuorg[1]) * S.inv_cell);
                    const unsigned qz = __double2uint_rd((pos[2] - S.cuorg[2]) * S.inv_cell);
                    ix = ax == 0 ? ipn : min(max(qx, ix & ~cm), ix | cm);
                    iy = ax == 1 ? ipn : min(max(qy, iy & ~cm), iy | cm);
                    iz = ax == 2 ? ipn : min(max(qz, iz & ~cm), iz | cm);
                    if (La >= K) {
                        const int b = D - 1 - La;
                        const int br = ((ix >> b) & 1) | (((iy >> b) & 1) << 1) | (((iz >> b) & 1) << 2);
                        w = __ldg(&S.nodes[(size_t)stk[(La - K) * NT + tid] * 8 + br]);
                        L = La + 1;
                    } else {
                        const int s2 = D - K;
                        const int2 e = __ldg(&S.top[(ix >> s2) | ((iy >> s2) << K) | ((iz >> s2) << (2 * K))]);
                        w = e.x; L = e.y;
                    }
                    RB_STAT(ws.nodes++;)
                }
#else
                // step to the neighbour, ascending on overflow (raytrace.c:688-706):
                // climb while the cell coordinate along ax cannot move that way
                const bool positive = dirf & (1 << ax);
                const unsigned ia = ax == 0 ? ix : ax == 1 ? iy : iz;
                const unsigned blocked = positive ? ia : ~ia;          // trailing ones = levels to climb
                const int up = (~blocked) ? __ffs(~blocked) - 1 : 32;  // number of trailing one bits
                if (up >= L) { done = true; if (ro >= 0) fl |= WF_RESULT; break; }    // left the scene cube
                ix >>= up; iy >>= up; iz >>= up; L -= up;
                if (ax == 0) ix ^= 1; else if (ax == 1) iy ^= 1; else iz ^= 1;
                const int br = (ix & 1) | ((iy & 1) << 1) | ((iz & 1) << 2);
                w = __ldg(&S.nodes[(size_t)stk[(L - 1) * NT + tid] * 8 + br]);
                RB_STAT(ws.nodes++;)
#endif
                if ((w != -1) | (cstep + 1 >= RB_CSTEPS)) break;
                fullc = false;
            }
#if RB_PREFETCH
            // the word just read names what the next round will read first: say so now, a round's worth of
            // work ahead of the dependent load
            if (!done) {
                if ((RB_PREFETCH & 1) && w < -1)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(pool + ((unsigned)(-w - 2) >> 4)));
                if ((RB_PREFETCH & 2) && w >= 0)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.nodes + (size_t)w * 8));
            }
#endif
            px = pos[0]; py = pos[1]; pz = pos[2];
            sm.pos[0][tid] = px; sm.pos[1][tid] = py; sm.pos[2][tid] = pz;
            Ld = (Ld & ~0xff) | L;
            if (done) fl |= WF_DONE;
        }
    }
}

}  // namespace rb
