// rb_shade.cuh -- material shading, ray spawning and contribution accumulation
// on the device, written as a FORWARD expansion of the reference's recursive
// ray tree: every spawned ray carries the product of rcoef from the primary
// ray down to itself, so the value/coefficients the reference gathers on the
// way back up its recursion (rt/raytrace.c:407-442 raycontrib, and the
// saddscolor() chains in the materials) are added where a ray ends instead.
//
// Restated reference functions:
//   rayorigin                 src/radiance/rt/raytrace.c:39-135
//   raytrans / rayshade       src/radiance/rt/raytrace.c:182-256
//   sourcehit                 src/radiance/rt/source.c:316-377
//   m_light (+macros)         src/radiance/rt/source.c:678-793
//   m_normal + dirnorm        src/radiance/rt/normal.c:71-360 (gaussamp :363-495)
//   m_glass                   src/radiance/rt/glass.c:46-165
//   multambient (aa=0)        src/radiance/rt/ambient.c:229-297
//   samp_hemi / ambsample     src/radiance/rt/ambcomp.c:350-422,177-248
//   direct / srcray / nextssamp  src/radiance/rt/source.c:219-257,398-556, srcsamp.c:36-144
//   trace_contrib             src/radiance/rt/rcontrib.c:272-317
//   square2disk               src/radiance/common/disk2square.c:43-79
//   getperpendicular          src/radiance/common/fvect.c:159-196
#pragma once
#include <cooperative_groups.h>
#include "rb_device.cuh"
#include "rb_bins.cuh"
#include "rb_geom.cuh"
#include "rb_bsdf.cuh"

namespace rb {
namespace cg = cooperative_groups;

// ---------------------------------------------------------------- RNG ------
// Counter-based: a 64-bit path key identifies a ray in the tree (derived from
// the global record index, so results do not depend on batch or GPU count);
// dimension d of a ray's random vector is mix64(key + d*C).
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double rnd01(unsigned long long key, unsigned dim) {
    return (double)(mix64(key + (unsigned long long)dim * 0xd1342543de82ef95ULL) >> 11) *
           (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ unsigned long long child_key(unsigned long long key, unsigned long long idx) {
    return mix64(key ^ mix64(idx + 0x632be59bd9b4e019ULL));
}

// ------------------------------------------------------------- context -----
struct DirectJob;
struct WaveArgs {
    DScene S;
    DParams P;
    const DBinSpec* bins;
    int nbinspecs;
    const QRay* qin; unsigned nin;
    HitRec* hits;           // [qcap] k_trace -> k_shade
    QRay* qout; unsigned qcap;
    QHemi* hout; unsigned hcap;
    double* acc;            // [rows][ncols][3] contribution accumulators (or null)
    int ncols;
    double* vacc;           // [rows][3] value accumulators (or null)
    unsigned row0;          // first row of this batch
    RayResult* res;         // per-row primary-hit report (or null)
    DCounters* C;
    int inline_hemi_max;    // hemispheres with n*n <= this are expanded in-thread
    DirectJob* dout; unsigned dcap;   // parked direct() calculations (null: sources are walked in-thread)
    unsigned* slow;         // [qcap] queue slots k_shade_fast leaves to the general k_shade (null: no split)
    unsigned* mid;          // [qcap] queue slots it leaves to k_shade_mid (glass, trans, spotlights)
    unsigned* lean;         // [qcap] queue slots it leaves to k_shade_lean
    unsigned* spec;         // [qcap] queue slots it leaves to k_shade_spec
    int nodirect;           // direct() has no source to sample in this scene (every source is a glow that is skipped)
    int anyhit;             // k_trace: shadow rays towards distant sources end at any surface opaque to shadow rays
};

struct RayCtx {             // the ray being shaded (a subset of RAY, rt/ray.h:48-83)
    double org[3], dir[3], rmax;
    double rot, rod, rop[3], ron[3];
    float coef[3];          // cumulative coefficient incl. own rcoef
    float rweight;
    unsigned row;
    int crtype, rlvl, rdepth, rsrc;
    int robj;               // object hit (-1: none / fake irradiance hit)
    bool flat;              // isflat(ro->otype)
    unsigned char xfl;      // 1: smooth mesh triangle (vertex normals), 2: its modifier is named "Phong"
    unsigned long long key;
    unsigned nchild;        // children spawned so far (for key derivation)
    unsigned med;           // medium the ray travelled in (QRay.med), possibly replaced by the dielectric it hit from inside
    float re;               // min(cext) * rot of that medium: rayorigin()'s extinction estimate for the children
};

__device__ __forceinline__ float max3(const float c[3]) { return fmaxf(c[0], fmaxf(c[1], c[2])); }

__device__ __forceinline__ unsigned pack_info(int crtype, int rlvl, int rdepth) {
    return (unsigned)(crtype & 0x3ff) | ((unsigned)(rlvl & 0x3f) << 10) | ((unsigned)(rdepth & 0x3f) << 16);
}

// warp-aggregated slot reservation
__device__ __forceinline__ unsigned reserve_slot(unsigned* ctr) {
    cg::coalesced_group g = cg::coalesced_threads();
    unsigned base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(ctr, g.size());
    base = g.shfl(base, 0);
    return base + g.thread_rank();
}

// rayorigin()'s extinction estimate of a child's weight (raytrace.c:96-107), out of line: rays in an absorbing
// medium are rare and k_shade is instruction-fetch sensitive (call sites test re > 0.1 first)
__device__ __noinline__ float ext_weight(float rweight, float re) {
    return re > 92.f ? 0.f : (float)((double)rweight * exp(-(double)re));
}

// ---- absorbing media (dielectric.c sets RAY.cext; raytrace.c:259-295 rayparticipate applies it) ----
// A ray names its medium by QRay.med / RayCtx.med: 0 = none (the global -me medium is not built and is
// rejected), else (material slot + 1) << 1 | side, side 0 = reals 0..2 of that dielectric / interface
// (its inside), side 1 = reals 4..6 of an interface (its outside).  Extinction per unit length is
// -mylog(transmission per unit length), dielectric.c:55-66.
__device__ __forceinline__ float mylogf(float x) { return x < 1e-40f ? -100.f : x >= 1.f ? 0.f : (float)log((double)x); }
__device__ __forceinline__ unsigned medium_id(int slot, int side) { return ((unsigned)(slot + 1) << 1) | (unsigned)side; }
__device__ __noinline__ void medium_cext(const DScene& S, unsigned med, float cext[3]) {
    const MatRec& m = S.mats[(med >> 1) - 1];
    const int o = (med & 1) ? 4 : 0;
    for (int k = 0; k < 3; k++) cext[k] = -mylogf(m.a[o + k]);
}
// rayparticipate() of a non-scattering medium for the ray being shaded, in the forward form: everything the
// ray and its descendants will add is worth exp(-cext * rot) of it, so its cumulative coefficient is scaled
// before anything is spawned.  rcontrib's coefficients carry the same factor: raycontrib() multiplies the
// product of rcoef by exp(-sum of cext * rot) over the chain, the contributing ray included
// (raytrace.c:407-442).  `over` >= 0: the medium is replaced first, the way
// m_dielectric() writes r->cext of a ray that arrives from inside (or at an interface from outside).
// Also leaves min(cext) * rot for rayorigin()'s weight estimate (raytrace.c:96-107).
__device__ __noinline__ void ray_medium(const WaveArgs& A, RayCtx& r, int over) {
    if (over >= 0) r.med = (unsigned)over;
    else if (r.robj >= 0) {               // about to be replaced by the dielectric this ray hits? then leave it to m_dielectric
        const int ms = __ldg(&A.S.objhdr[r.robj]).z;
        if (ms >= 0) {
            const int k = A.S.mats[ms].kind;
            if ((k == MK_DIELECTRIC && r.rod < 0.0) || k == MK_INTERFACE) return;
        }
    }
    if (!r.med) { r.re = 0.f; return; }
    float cext[3];
    medium_cext(A.S, r.med, cext);
    r.re = (float)((double)fminf(cext[0], fminf(cext[1], cext[2])) * r.rot);
    if (fmaxf(cext[0], fmaxf(cext[1], cext[2])) <= (float)(1. / RB_FHUGE)) return;
    for (int k = 0; k < 3; k++) {
        const double e = r.rot * (double)cext[k];
        r.coef[k] *= (float)(e <= RB_FTINY ? 1. : e > 92. ? 0. : exp(-e));
    }
}

// raytrace.c:39-135.  `rc` is the child's coefficient w.r.t. the parent (may be
// rescaled by Russian roulette); has_rc=false stands for rc==NULL.
__device__ __forceinline__ bool rayorigin(const DParams& P, RayCtx& par, int rt, float rc[3], bool has_rc, QRay& q) {
    float rw = 1.0f;
    if (has_rc) { rw = max3(rc); if (rw > 1.0f) rw = 1.0f; }
    else rc[0] = rc[1] = rc[2] = 1.0f;
    if (par.rot >= RB_FHUGE * .99) return false;        // illegal continuation
    int rlvl = par.rlvl, rsrc = par.rsrc;
    double rmax;
    if (rt & RT_RAYREFL) {
        rlvl++;
        if (rsrc >= 0) rsrc = -1;
        rmax = 0.0;
    } else
        rmax = (par.rmax > RB_FTINY) * (par.rmax - par.rot);
    int crtype = par.crtype | rt;
    float rweight = par.rweight * rw;
    if (par.re > 0.1f) rweight = ext_weight(rweight, par.re);   // estimate extinction
    unsigned long long key = child_key(par.key, par.nchild++);
    if (rweight <= 0.0f) return false;
    if (!(crtype & RT_SHADOW)) {
        if ((P.maxdepth <= 0) & has_rc) {               // Russian roulette
            if ((P.maxdepth < 0) & (rlvl > -P.maxdepth)) return false;
            if (rweight < P.minweight) {
                if (rnd01(key, 7) > (double)(rweight / P.minweight)) return false;
                float s = P.minweight / rweight;
                rc[0] *= s; rc[1] *= s; rc[2] *= s;
                rweight = P.minweight;
            }
        } else if (!((rweight >= P.minweight) & (rlvl <= abs(P.maxdepth))))
            return false;
    }
    q.org[0] = par.rop[0]; q.org[1] = par.rop[1]; q.org[2] = par.rop[2];
    q.rmax = rmax;
    q.coef[0] = par.coef[0] * rc[0]; q.coef[1] = par.coef[1] * rc[1]; q.coef[2] = par.coef[2] * rc[2];
    q.rweight = rweight;
    q.row = par.row;
    q.info = pack_info(crtype, rlvl, par.rdepth);
    q.rsrc = rsrc;
    q.key_lo = (unsigned)key; q.key_hi = (unsigned)(key >> 32);
    q.med = par.med;
    return true;
}

__device__ __forceinline__ void push_ray(const WaveArgs& A, const QRay& q) {
    unsigned slot = reserve_slot(&A.C->nq_out);
    if (slot >= A.qcap) { A.C->overflow = 1; return; }
    A.qout[slot] = q;
}

// fvect.c:130-157
__device__ __forceinline__ double normalize3(double v[3]) {
    double d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (d == 0.0) return 0.0;
    double len;
    if ((d <= 1.0 + RB_FTINY) & (d >= 1.0 - RB_FTINY)) { len = 0.5 + 0.5 * d; d = 2.0 - len; }
    else { len = sqrt(d); d = 1.0 / len; }
    v[0] *= d; v[1] *= d; v[2] *= d;
    return len;
}

// disk2square.c:43-79
__device__ __forceinline__ void square2disk(double ds[2], double seedx, double seedy) {
    double phi, r;
    double a = 2. * seedx - 1;
    double b = 2. * seedy - 1;
    const double PI4 = RB_PI / 4.;
    if (a > -b) {
        if (a > b) { r = a; phi = PI4 * (b / a); }
        else { r = b; phi = PI4 * (2. - (a / b)); }
    } else {
        if (a < b) { r = -a; phi = PI4 * (4. + (b / a)); }
        else { r = -b; phi = (b != 0.) ? PI4 * (6. - (a / b)) : 0.; }
    }
    r *= 0.9999999999999;
    double s, c;
    sincos(phi, &s, &c);
    ds[0] = r * c; ds[1] = r * s;
}

// fvect.c:159-196 with randomize=1 (random numbers from the path key)
__device__ __forceinline__ bool getperpendicular_rand(double vp[3], const double v[3],
                                                       unsigned long long key) {
    double v1[3];
    v1[0] = 0.5 - rnd01(key, 11); v1[1] = 0.5 - rnd01(key, 12); v1[2] = 0.5 - rnd01(key, 13);
    int ord[3];
    switch ((int)(6 * rnd01(key, 14))) {
    case 0: ord[0] = 0; ord[1] = 1; ord[2] = 2; break;
    case 1: ord[0] = 0; ord[1] = 2; ord[2] = 1; break;
    case 2: ord[0] = 1; ord[1] = 0; ord[2] = 2; break;
    case 3: ord[0] = 1; ord[1] = 2; ord[2] = 0; break;
    case 4: ord[0] = 2; ord[1] = 0; ord[2] = 1; break;
    default: ord[0] = 2; ord[1] = 1; ord[2] = 0; break;
    }
    int i;
    for (i = 3; i--;) {
        double c = ord[i] == 0 ? v[0] : ord[i] == 1 ? v[1] : v[2];
        if ((-0.6 < c) & (c < 0.6)) break;
    }
    if (i < 0) return false;
    if (ord[i] == 0) v1[0] = 1.0; else if (ord[i] == 1) v1[1] = 1.0; else v1[2] = 1.0;
    vp[0] = v1[1] * v[2] - v1[2] * v[1];
    vp[1] = v1[2] * v[0] - v1[0] * v[2];
    vp[2] = v1[0] * v[1] - v1[1] * v[0];
    return normalize3(vp) > 0.0;
}

// One ambient division sample (ambcomp.c:177-248 ambsample, value-only part).
// par describes the ray whose hit spawns the hemisphere.
__device__ __forceinline__ bool ambsample(const DParams& P, RayCtx& par, int atyp, const float acoef[3],
                                          const double onrm[3], const double ux[3], const double uy[3],
                                          int n, int i, int j, QRay& q) {
    float rc[3] = {acoef[0], acoef[1], acoef[2]};
    if (!rayorigin(P, par, atyp, rc, true, q)) return false;
    unsigned long long key = ((unsigned long long)q.key_hi << 32) | q.key_lo;
    double ss0 = rnd01(key, 1), ss1 = rnd01(key, 2), spt[2];
    square2disk(spt, (j + ss1) / n, (i + ss0) / n);
    double zd = sqrt(1. - spt[0] * spt[0] - spt[1] * spt[1]);
    for (int k = 0; k < 3; k++) q.dir[k] = spt[0] * ux[k] + spt[1] * uy[k] + zd * onrm[k];
    normalize3(q.dir);                 // checknorm() is normalize() under -ffast-math (rt/ray.h:270-274)
    unsigned info = q.info;            // ambient children are one bounce deeper
    int rdepth = ((info >> 16) & 0x3f) + 1;
    q.info = (info & 0xffc0ffffu) | ((unsigned)(rdepth & 0x3f) << 16);
    return true;
}

// ------------------------------------------------------ accumulation -------
__device__ __forceinline__ void add_value(const WaveArgs& A, unsigned row, const float c[3],
                                          float r, float g, float b) {
    if (!A.vacc) return;
    double* v = A.vacc + (size_t)(row - A.row0) * 3;
    atomicAdd(v + 0, (double)(c[0] * r));
    atomicAdd(v + 1, (double)(c[1] * g));
    atomicAdd(v + 2, (double)(c[2] * b));
}

// Bin function + accumulation of one contribution.  Out of line on purpose: one
// ray in a hundred gets here, and the bin functions (asin/acos/atan2 in double)
// are 30 % of k_shade's SASS -- kept out of the hot path's instruction stream.
// (Scalars by value: reference arguments of a real call live in local memory.
// Moving m_glass / gaussamp / raytrans out of line the same way made k_shade 70 %
// SLOWER -- they take the whole RayCtx by reference.)
// Accumulation is WARP-AGGREGATED: the lanes that arrive here together and add to the same coefficient (same record,
// same bin: neighbouring rays of one sensor usually do) are found with __match_any_sync, their values are summed
// by the group's first lane in double, and that lane alone issues the three atomics.  A sun-coefficient matrix adds
// hundreds of millions of contributions, many of a warp's to one (row, sun) cell.
__device__ __noinline__ void add_contrib(const DBinSpec* b, double* row, DCounters* C, double dx, double dy,
                                         double dz, float c0, float c1, float c2) {
    const double D[3] = {dx, dy, dz};
    double bval = rb_eval_bin(*b, D);
    if (bval <= -.5) return;
    int bn = (int)(bval + .5);
    if (bn >= b->nbins) { atomicAdd(&C->badbin, 1u); return; }
    double* d = row + (size_t)(b->col0 + bn) * 3;
    const unsigned active = __activemask();
    const unsigned grp = __match_any_sync(active, (unsigned long long)(size_t)d);
    const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
    double s0 = 0., s1 = 0., s2 = 0.;
    for (unsigned m = grp; m; m &= m - 1) {           // every member runs the same loop: shuffles stay convergent within the group
        const int j = __ffs(m) - 1;
        s0 += (double)__shfl_sync(grp, c0, j);
        s1 += (double)__shfl_sync(grp, c1, j);
        s2 += (double)__shfl_sync(grp, c2, j);
    }
    if (lane == leader) {
        atomicAdd(d + 0, s0);
        atomicAdd(d + 1, s1);
        atomicAdd(d + 2, s2);
        atomicAdd(&C->contribs, (unsigned long long)__popc(grp));
    }
}

// rcontrib.c:272-317.  `rcoef_ok`: the ray's own coefficient was not zeroed by
// its material; rcol = the ray's returned radiance (emitters only).
__device__ __forceinline__ void trace_contrib(const WaveArgs& A, const RayCtx& r, bool rcoef_zeroed,
                                              const float rcol[3], bool have_rcol) {
    if (!A.acc || r.robj < 0) return;
    int4 hd = __ldg(&A.S.objhdr[r.robj]);
    if (hd.y < 0) return;                                   // void modifier
    if (r.rsrc >= 0 && A.S.srcs[r.rsrc].so != r.robj) return;   // shadow ray not on source
    int slot = __ldg(&A.S.otrack[r.robj]);
    if (slot < 0) return;
    if (rcoef_zeroed) return;
    float c[3] = {r.coef[0], r.coef[1], r.coef[2]};
    if (A.P.contrib) {
        // -V+ multiplies by the radiance the ray RETURNS (rcontrib.c:296-301): known here for emitters
        // only -- a forward wavefront has no returned value for a surface that reflects or transmits
        if (!have_rcol) { atomicOr(&A.C->errflag, RB_ERR_CONTRIB_VALUE); A.C->errobj = (unsigned)r.robj; return; }
        c[0] *= rcol[0]; c[1] *= rcol[1]; c[2] *= rcol[2];
        // reference tests rcoef*rcol of the ray itself; the chain product has the same zero set
        if (!(c[0] > 0.f || c[1] > 0.f || c[2] > 0.f)) return;
    }
    add_contrib(A.bins + slot, A.acc + (size_t)(r.row - A.row0) * A.ncols * 3, A.C, r.dir[0], r.dir[1], r.dir[2],
                c[0], c[1], c[2]);
}

// ----------------------------------------------------------- sources -------
// source.c:316-377.  Returns the source index whose object becomes r->ro, or -1.
__device__ __forceinline__ int sourcehit(const DScene& S, const double dir[3], int rsrc, int crtype) {
    int glowsrc = -1, transrc = -1;
    int first = 0, last = S.nsrcs - 1;
    if (rsrc >= 0) first = last = rsrc;
    for (int i = first; i <= last; i++) {
        const SrcRec& s = S.srcs[i];
        if (!(s.flags & SF_DISTANT)) continue;
        if (2. * RB_PI * (1. - (s.sloc[0] * dir[0] + s.sloc[1] * dir[1] + s.sloc[2] * dir[2])) > s.ss2) continue;
        if (i == rsrc) return i;
        if (s.flags & SF_SKIP) { if (glowsrc < 0) glowsrc = i; continue; }
        if (s.flags & 0x100) { if (transrc < 0) transrc = i; continue; }   // transparent illum
        return i;
    }
    if (transrc >= 0 && (crtype & (RT_AMBIENT | RT_SPECULAR))) return -1;
    return glowsrc;
}

// ------------------------------------------------------------ shading ------
struct NormDat {           // normal.c:53-67 NORMDAT
    int specfl;
    float mcolor[3], scolor[3];
    double prdir[3];
    double alpha2, rdiff, rspec, trans, tdiff, tspec;
    double pnorm[3], pdot;
    // plastic2 / metal2 / trans2 only (aniso.c:46-62 ANISODAT), valid when specfl & SP_ANISO
    double u[3], v[3], u_alpha, v_alpha;
    // BSDF / aBSDF only (m_bsdf.c:80-96 BSDFDAT), valid when specfl & SP_BSDF: mcolor = rdiff, scolor = tdiff,
    // prdir = vray, u / v / w = the rows of toloc, pnorm as handed to direct() (turned towards the hit side)
    double w[3];
    float cthru[3], cthru_surr[3];
    int bsdf, dmode;       // dmode: 0 dir_bsdf, 1 dir_brdf, 2 dir_btdf
};
struct DirectJob { RayCtx r; NormDat nd; };
enum : int { SP_REFL = 01, SP_TRAN = 02, SP_PURE = 04, SP_FLAT = 010, SP_RBLT = 020, SP_TBLT = 040, SP_ANISO = 0100,
             SP_BSDF = 0200 };
__device__ void dir_bsdf(float scval[3], const WaveArgs& A, const NormDat& np, const RayCtx& r, const double ldir[3], double omega,
                         unsigned long long jkey);

// aniso.c:64-182 diraniso(): source coefficient of the anisotropic Gaussian (Ward / Geisler-Moroder-Duer)
__device__ __forceinline__ void diraniso(float scval[3], const NormDat& np, const RayCtx& r, const double ldir[3],
                                         double omega, double dstrsrc) {
    scval[0] = scval[1] = scval[2] = 0.f;
    const double ldot = dot3(np.pnorm, ldir);
    if (ldot < 0.0 ? np.trans <= RB_FTINY : np.trans >= 1.0 - RB_FTINY) return;      // wrong side
    if ((ldot > RB_FTINY) & (np.rdiff > RB_FTINY)) {
        const double w = ldot * omega * np.rdiff * (1.0 / RB_PI);
        for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * w);
    }
    if ((ldot < -RB_FTINY) & (np.tdiff > RB_FTINY)) {
        const double w = -ldot * omega * np.tdiff * (1.0 / RB_PI);
        for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * w);
    }
    const double ua2 = np.u_alpha * np.u_alpha, va2 = np.v_alpha * np.v_alpha;
    if ((ldot > RB_FTINY) && (np.specfl & SP_REFL)) {
        double au2 = (np.specfl & SP_FLAT) ? (1. - dstrsrc) * omega * (0.25 / RB_PI) : 0.0;   // source width if flat
        double av2 = au2;
        au2 += ua2; av2 += va2;
        const double h[3] = {ldir[0] - r.dir[0], ldir[1] - r.dir[1], ldir[2] - r.dir[2]};     // half vector
        double e1 = dot3(np.u, h); e1 *= e1 / au2;
        double e2 = dot3(np.v, h); e2 *= e2 / av2;
        double nh = dot3(np.pnorm, h); nh *= nh;
        e1 = (e1 + e2) / nh;
        double w = exp(-e1) * dot3(h, h) / (RB_PI * nh * nh * sqrt(au2 * av2));
        if (w > RB_FTINY) {
            w *= ldot * omega;
            for (int k = 0; k < 3; k++) scval[k] += (float)(np.scolor[k] * w);
        }
    }
    if ((ldot < -RB_FTINY) && (np.specfl & SP_TRAN)) {
        double au2 = omega * (1.0 / RB_PI), av2 = au2;
        au2 += ua2; av2 += va2;
        const double h[3] = {ldir[0] - np.prdir[0], ldir[1] - np.prdir[1], ldir[2] - np.prdir[2]};
        double w = dot3(h, h);
        if (w > RB_FTINY * RB_FTINY) { const double e = dot3(h, np.pnorm); w = 1.0 - e * e / w; }
        if (w > RB_FTINY * RB_FTINY) {
            double e1 = dot3(h, np.u); e1 *= e1 / au2;
            double e2 = dot3(h, np.v); e2 *= e2 / av2;
            w = exp(-((e1 + e2) / w));
        } else
            w = 1.0;
        w *= (1.0 / RB_PI) * sqrt(-ldot / (np.pdot * au2 * av2));
        if (w > RB_FTINY) {
            w *= np.tspec * omega;
            for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * w);
        }
    }
}

// normal.c:71-173
// `jkey`: random key of the shadow ray being set up (the BSDF materials jitter their evaluation, m_bsdf.c:309-329)
// FAST (k_shade_fast): isotropic materials only -- the other two are not in that kernel's code at all.
template <bool FAST = false>
__device__ __forceinline__ void dirnorm(float scval[3], const WaveArgs& A, const NormDat& np, const RayCtx& r, const double ldir[3],
                        double omega, double dstrsrc, unsigned long long jkey) {
    if (!FAST) {
        if (np.specfl & SP_BSDF) { dir_bsdf(scval, A, np, r, ldir, omega, jkey); return; }
        if (np.specfl & SP_ANISO) { diraniso(scval, np, r, ldir, omega, dstrsrc); return; }
    }
    scval[0] = scval[1] = scval[2] = 0.f;
    double ldot = dot3(np.pnorm, ldir);
    if (ldot < 0.0 ? np.trans <= RB_FTINY : np.trans >= 1.0 - RB_FTINY) return;
    double lrdiff = np.rdiff, ltdiff = np.tdiff;
    if ((np.specfl & SP_PURE) && np.rspec >= 0.017999 && ((lrdiff > RB_FTINY) | (ltdiff > RB_FTINY))) {
        double dtmp = 1. - (exp(-5.85 * fabs(ldot)) - 0.00202943064);
        lrdiff *= dtmp; ltdiff *= dtmp;
    }
    if ((ldot > RB_FTINY) & (lrdiff > RB_FTINY)) {
        double dtmp = ldot * omega * lrdiff * (1.0 / RB_PI);
        for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * dtmp);
    }
    if ((ldot < -RB_FTINY) & (ltdiff > RB_FTINY)) {
        double dtmp = -ldot * omega * ltdiff * (1.0 / RB_PI);
        for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * dtmp);
    }
    if ((ldot > RB_FTINY) & ((np.specfl & (SP_REFL | SP_PURE)) == SP_REFL)) {
        double dtmp = np.alpha2;
        if (np.specfl & SP_FLAT) dtmp += (1. - dstrsrc) * omega * (0.25 / RB_PI);
        double vtmp[3] = {ldir[0] - r.dir[0], ldir[1] - r.dir[1], ldir[2] - r.dir[2]};
        double d2 = dot3(vtmp, np.pnorm);
        d2 *= d2;
        double d3 = dot3(vtmp, vtmp);
        double d4 = (d3 - d2) / d2;
        dtmp = exp(-d4 / dtmp) * d3 / (RB_PI * d2 * d2 * dtmp);
        if (dtmp > RB_FTINY) {
            dtmp *= ldot * omega;
            for (int k = 0; k < 3; k++) scval[k] += (float)(np.scolor[k] * dtmp);
        }
    }
    if ((ldot < -RB_FTINY) & ((np.specfl & (SP_TRAN | SP_PURE)) == SP_TRAN)) {
        double dtmp = np.alpha2 + omega * (1.0 / RB_PI);
        dtmp = exp((2. * dot3(np.prdir, ldir) - 2.) / dtmp) / (RB_PI * dtmp);
        if (dtmp > RB_FTINY) {
            dtmp *= np.tspec * omega * sqrt(-ldot / np.pdot);
            for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * dtmp);
        }
    }
}

// source.c:398-556 direct(), with every source tested (the reference's -dt 0
// behaviour, which rcontrib forces: rcmain.c:164-171).  One source -> one
// shadow ray; its random key is child (nchild0 + sn) of the shaded ray, so the
// serial and the warp-cooperative forms below emit identical rays.
template <bool FAST = false>
__device__ __forceinline__ void direct_one(const WaveArgs& A, const RayCtx& r, const NormDat& nd, int sn,
                                           unsigned nchild0) {
    const SrcRec& s = A.S.srcs[sn];
    if (s.flags & SF_SKIP) return;                   // srcskip()
    if (!(s.flags & SF_DISTANT)) return;
    if (r.med) {                                     // srcvalue() -> rayparticipate() over FHUGE: nothing is left of a
        float cext[3];                               // distant source in an absorbing medium, so it is never tested
        medium_cext(A.S, r.med, cext);
        if (fminf(cext[0], fminf(cext[1], cext[2])) * RB_FHUGE > 92.) return;
    }
    // srcray(): rayorigin(sr, SHADOW, r, NULL) never fails for weight > 0
    unsigned long long key = child_key(r.key, nchild0 + (unsigned)sn);
    double vpos[3] = {0, 0, 0};
    if (A.P.dstrsrc > RB_FTINY) {                    // srcsamp.c:67-79 jitter
        for (int k = 0; k < 3; k++) vpos[k] = A.P.dstrsrc * (1. - 2. * rnd01(key, 20 + k));
    }
    if ((s.flags & SF_CIRC) && (A.P.dstrsrc > 0.7)) {   // srcsamp.c:83-107
        double d = 1.12837917;
        double t0 = d * sqrt(1.0 - 0.5 * vpos[1] * vpos[1]);
        double t1 = d * sqrt(1.0 - 0.5 * vpos[0] * vpos[0]);
        vpos[0] *= t0; vpos[1] *= t1; vpos[2] *= 0.0;
    }
    double ldir[3];
    for (int k = 0; k < 3; k++)
        ldir[k] = s.sloc[k] + vpos[0] * s.ss[0][k] + vpos[1] * s.ss[1][k] + vpos[2] * s.ss[2][k];
    if (normalize3(ldir) == 0.0) return;
    double dom = s.ss2;                              // nopart: whole source
    float scval[3];
    dirnorm<FAST>(scval, A, nd, r, ldir, dom, A.P.dstrsrc, key);
    if (!(max3(scval) > 0.f)) return;
    // shadow test ray: TSHADOW if through the surface (ray.h:87 thrudir)
    bool thru = (r.rod > 0) ^ (dot3(r.ron, ldir) > 0);
    int rt = thru ? RT_TSHADOW : RT_RSHADOW;
    if (r.rot >= RB_FHUGE * .99 || !(r.rweight > 0.f)) return;      // rayorigin() refusals
    QRay q;
    q.org[0] = r.rop[0]; q.org[1] = r.rop[1]; q.org[2] = r.rop[2];
    q.dir[0] = ldir[0]; q.dir[1] = ldir[1]; q.dir[2] = ldir[2];
    const bool refl = (rt & RT_RAYREFL) != 0;        // RSHADOW starts a new path segment, TSHADOW continues
    q.rmax = refl ? 0.0 : (r.rmax > RB_FTINY) * (r.rmax - r.rot);
    q.coef[0] = r.coef[0] * scval[0]; q.coef[1] = r.coef[1] * scval[1]; q.coef[2] = r.coef[2] * scval[2];
    q.rweight = r.re > 0.1f ? ext_weight(r.rweight, r.re) : r.rweight;
    q.row = r.row;
    q.info = pack_info(r.crtype | rt, r.rlvl + (refl ? 1 : 0), r.rdepth);
    q.rsrc = sn;
    q.key_lo = (unsigned)key; q.key_hi = (unsigned)(key >> 32);
    q.med = r.med;
    push_ray(A, q);
}

template <bool FAST = false>
__device__ __forceinline__ void direct(const WaveArgs& A, RayCtx& r, const NormDat& nd) {
    const int ns = A.S.nsrcs;
    for (int sn = 0; sn < ns; sn++) direct_one<FAST>(A, r, nd, sn, r.nchild);
    r.nchild += (unsigned)ns;
}

// ---- local light sources: srcsamp.c source partitioning + srcray() + srcvalue() ----
// srcsupp.c:294-322 spotout() for the shadow ray (org, dir)
__device__ __forceinline__ bool spotout(const SrcRec& s, const double org[3], const double dir[3]) {
    if (s.spot_flen < -(float)RB_FTINY) {            // distant-type spot
        double vd[3] = {s.spot_aim[0] - org[0], s.spot_aim[1] - org[1], s.spot_aim[2] - org[2]};
        double d = dot3(dir, vd);
        d = dot3(vd, vd) - d * d;
        return RB_PI * d > (double)s.spot_siz;
    }
    return (double)s.spot_siz < 2.0 * RB_PI * (1.0 + dot3(s.spot_aim, dir));
}

// One partition of a local source (integer centre ct / size sz in 1/64 units, srcsamp.c:57-144):
// sample -> solid angle -> proximity / spot tests -> coefficient -> aiming test against the
// source surface itself (srcvalue(), source.c:260-300) -> shadow ray.
__device__ __noinline__ void local_sample(const WaveArgs& A, const RayCtx& r, const NormDat& nd, int sn,
                                          unsigned long long key, int ctU, int ctV, int szU, int szV, bool many) {
    const SrcRec& s = A.S.srcs[sn];
    const double dj = A.P.dstrsrc;
    const double inv = 1.0 / 64.0;
    double vpos[3] = {0, 0, 0};
    const int sz[3] = {szU, szV, 64};
    if (dj > RB_FTINY) {
        vpos[0] = rnd01(key, 20); vpos[1] = rnd01(key, 21);
        vpos[2] = (s.flags & SF_FLAT) ? 0.5 : rnd01(key, 22);
        for (int i = 0; i < 3; i++) vpos[i] = dj * (1. - 2. * vpos[i]) * (double)sz[i] * inv;
    }
    vpos[0] += ctU * inv; vpos[1] += ctV * inv;
    if ((s.flags & SF_CIRC) && (many | (dj > 0.7))) {
        double trim[3];
        if (s.flags & (SF_FLAT | SF_DISTANT)) {
            const double d = 1.12837917;
            trim[0] = d * sqrt(1.0 - 0.5 * vpos[1] * vpos[1]);
            trim[1] = d * sqrt(1.0 - 0.5 * vpos[0] * vpos[0]);
            trim[2] = 0.0;
        } else {
            trim[2] = trim[0] = vpos[0] * vpos[0];
            double d = vpos[1] * vpos[1];
            if (d > trim[2]) trim[2] = d;
            trim[0] += d;
            d = vpos[2] * vpos[2];
            if (d > trim[2]) trim[2] = d;
            trim[0] += d;
            if (trim[0] > RB_FTINY * RB_FTINY) { d = 1.0 / 0.7236; trim[2] = trim[1] = trim[0] = d * sqrt(trim[2] / trim[0]); }
            else trim[2] = trim[1] = trim[0] = 0.0;
        }
        for (int i = 0; i < 3; i++) vpos[i] *= trim[i];
    }
    double ldir[3];
    for (int i = 0; i < 3; i++)
        ldir[i] = (s.sloc[i] + vpos[0] * s.ss[0][i] + vpos[1] * s.ss[1][i] + vpos[2] * s.ss[2][i]) - r.rop[i];
    double d = normalize3(ldir);
    if (d == 0.0) return;                                // at source!
    double dom;
    if (s.flags & SF_FLAT) dom = -dot3(s.ss[2], ldir) * (szU * szV * (inv * inv));
    else if (s.flags & SF_CYL) {
        double dd = dot3(ldir, s.ss[0]);
        dd *= dd / dot3(s.ss[0], s.ss[0]);
        dom = sqrt(1. - dd) * (szU * inv);
    } else dom = szU * szV * 64.0 * (inv * inv * inv);
    if (dom <= 1e-4) return;                             // behind source?
    dom *= s.ss2 / (d * d);
    if ((s.flags & SF_PROX) && d > s.prox) return;
    if (s.flags & SF_SPOT) {
        if (spotout(s, r.rop, ldir)) return;
        dom *= d * d; d += (double)s.spot_flen; dom /= d * d;
    }
    float scval[3];
    dirnorm(scval, A, nd, r, ldir, dom, dj, key);
    if (!(max3(scval) > 0.f)) return;
    // srcvalue(): the sample must hit the source surface, on its emitting side
    {
        const int4 hd = __ldg(&A.S.objhdr[s.so]);
        const int kind = hd.x & 0xff;
        const double* g = A.S.geom + hd.w;
        bool front = false, hit;
        if (kind == PK_FACE) {
            const double2* g2 = reinterpret_cast<const double2*>(g);
            const int hot = kind | (((hd.x >> 10) & 3) << 4) | (((hd.x >> 12) & 1) << 6);
            double t;
            hit = cand_face(hot, g, __ldg(&g2[0]), __ldg(&g2[1]), __ldg(reinterpret_cast<const float4*>(g + 4)),
                            r.rop, ldir, RB_FHUGE, t, front);
        } else {
            const double t = cand_other(kind, g, r.rop[0], r.rop[1], r.rop[2], ldir[0], ldir[1], ldir[2], RB_FHUGE);
            hit = t != 0.0; front = t > 0.0;
        }
        if (!(hit && front)) return;
    }
    if (r.rot >= RB_FHUGE * .99 || !(r.rweight > 0.f)) return;      // rayorigin() refusals
    const bool thru = (r.rod > 0) ^ (dot3(r.ron, ldir) > 0);
    const int rt = thru ? RT_TSHADOW : RT_RSHADOW;
    const bool refl = (rt & RT_RAYREFL) != 0;
    QRay q;
    q.org[0] = r.rop[0]; q.org[1] = r.rop[1]; q.org[2] = r.rop[2];
    q.dir[0] = ldir[0]; q.dir[1] = ldir[1]; q.dir[2] = ldir[2];
    q.rmax = refl ? 0.0 : (r.rmax > RB_FTINY) * (r.rmax - r.rot);
    q.coef[0] = r.coef[0] * scval[0]; q.coef[1] = r.coef[1] * scval[1]; q.coef[2] = r.coef[2] * scval[2];
    q.rweight = r.re > 0.1f ? ext_weight(r.rweight, r.re) : r.rweight;
    q.row = r.row;
    q.info = pack_info(r.crtype | rt, r.rlvl + (refl ? 1 : 0), r.rdepth);
    q.rsrc = sn;
    q.key_lo = (unsigned)key; q.key_hi = (unsigned)(key >> 32);
    q.med = r.med;
    push_ray(A, q);
}

// direct() for one LOCAL source: srcskip(), then nopart / flatpart / cylpart as an explicit-stack
// pre-order walk (lower half first, like flt_partit / cyl_partit), one sample per leaf.
__device__ void direct_local(const WaveArgs& A, const RayCtx& r, const NormDat& nd, int sn, unsigned nchild0) {
    const SrcRec& s = A.S.srcs[sn];
    if (s.flags & SF_SKIP) return;
    const double* ro = r.rop;
    auto dist2 = [](const double* a, const double* b) {
        return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
    };
    if (s.flags & SF_PROX) {                             // srcsamp.c:28-31
        const double lim = s.prox + s.srad;
        if (dist2(ro, s.sloc) > lim * lim) return;
    }
    const unsigned long long base = (unsigned long long)nchild0 + (unsigned)A.S.nsrcs + (unsigned long long)sn * 64u;
    const double ds = A.P.srcsizerat;
    const bool flat = (s.flags & SF_FLAT) != 0, cyl = (s.flags & SF_CYL) != 0;
    if (ds <= RB_FTINY || !(flat | cyl)) {               // nopart()
        local_sample(A, r, nd, sn, child_key(r.key, base), 0, 0, 64, 64, false);
        return;
    }
    double du2, dv2 = 0.0;
    if (cyl) {                                           // cylpart(), srcsamp.c:229-264
        const double rad2 = 1.365 * dot3(s.ss[1], s.ss[1]);
        const double v[3] = {ro[0] - s.sloc[0], ro[1] - s.sloc[1], ro[2] - s.sloc[2]};
        double d2 = dot3(v, s.ss[0]);
        double safedist2 = dot3(s.ss[0], s.ss[0]);
        d2 *= d2 / safedist2;
        const double dist2cent = dot3(v, v);
        d2 = dist2cent - d2;
        if (d2 <= rad2) return;                          // point inside extended cylinder
        safedist2 *= 4. * (double)r.rweight * (double)r.rweight / (ds * ds);
        if (d2 <= 4. * rad2 || dist2cent >= safedist2) {
            local_sample(A, r, nd, sn, child_key(r.key, base), 0, 0, 64, 64, false);
            return;
        }
        du2 = safedist2;
    } else {                                             // flatpart(), srcsamp.c:322-352
        const double v[3] = {ro[0] - s.sloc[0], ro[1] - s.sloc[1], ro[2] - s.sloc[2]};
        if (dot3(v, s.ss[2]) <= 0.) return;              // behind source
        dv2 = 2. * (double)r.rweight / ds;
        dv2 *= dv2;
        du2 = dv2 * dot3(s.ss[0], s.ss[0]);
        dv2 *= dot3(s.ss[1], s.ss[1]);
    }
    // explicit stack: centre point, integer centre/size, remaining budget mp, halvings of U and V
    double scx[8], scy[8], scz[8];
    short sct[8][2], ssz[8][2];
    unsigned char smp[8], sa[8], sb[8];
    int top = 0, leaf = 0;
    scx[0] = s.sloc[0]; scy[0] = s.sloc[1]; scz[0] = s.sloc[2];
    sct[0][0] = sct[0][1] = 0; ssz[0][0] = ssz[0][1] = 64; smp[0] = 64; sa[0] = sb[0] = 0;
    bool many = false;
    while (top >= 0) {
        const double c[3] = {scx[top], scy[top], scz[top]};
        const int ctU = sct[top][0], ctV = sct[top][1], szU = ssz[top][0], szV = ssz[top][1];
        const int mp = smp[top], a = sa[top], b = sb[top];
        top--;
        const double lu2 = du2 * __hiloint2double((1023 - 2 * a) << 20, 0);      // du2 * 0.25^a, exact
        const double lv2 = dv2 * __hiloint2double((1023 - 2 * b) << 20, 0);
        const double d2 = dist2(ro, c);
        const bool isleaf = cyl ? (mp < 2 || d2 >= lu2) : (mp < 2 || (d2 >= lu2 && d2 >= lv2));
        if (isleaf) {
            local_sample(A, r, nd, sn, child_key(r.key, base + (unsigned)leaf), ctU, ctV, szU, szV, many);
            leaf++;
            continue;
        }
        many = true;                                     // (the root split: np > 1 for every leaf)
        const bool inU = cyl || (lu2 > lv2);
        const double h = __hiloint2double((1023 - ((inU ? a : b) + 1)) << 20, 0);   // 0.5^(n+1), exact
        const double* ax = inU ? s.ss[0] : s.ss[1];
        const double nx = h * ax[0], ny = h * ax[1], nz = h * ax[2];
        const int hs = (inU ? szU : szV) >> 1;
        // push upper first so that the lower half is walked first
        for (int up = 1; up >= 0; up--) {
            top++;
            const double sg = up ? 1.0 : -1.0;
            scx[top] = c[0] + sg * nx; scy[top] = c[1] + sg * ny; scz[top] = c[2] + sg * nz;
            sct[top][0] = (short)(inU ? ctU + (up ? hs : -hs) : ctU);
            sct[top][1] = (short)(inU ? ctV : ctV + (up ? hs : -hs));
            ssz[top][0] = (short)(inU ? hs : szU); ssz[top][1] = (short)(inU ? szV : hs);
            smp[top] = (unsigned char)(mp / 2);
            sa[top] = (unsigned char)(inU ? a + 1 : a); sb[top] = (unsigned char)(inU ? b : b + 1);
        }
    }
}

// Scenes with many sources (a 5-phase sun matrix has thousands): the shading
// thread parks its state in a job queue and k_direct walks the source list
// with one CTA per job, writing its shadow rays to consecutive queue slots.
#ifndef RB_COOP_SRC_MIN
#define RB_COOP_SRC_MIN 64
#endif

template <bool FAST = false>
__device__ __forceinline__ void direct_or_park(const WaveArgs& A, RayCtx& r, const NormDat& nd) {
    if (A.dout) {
        unsigned slot = reserve_slot(&A.C->nd_out);
        if (slot >= A.dcap) { A.C->overflow = 1; return; }
        A.dout[slot].r = r;
        A.dout[slot].nd = nd;
        return;
    }
    if (A.nodirect) return;                          // every source is a glow that direct() skips (srcskip)
    direct<FAST>(A, r, nd);
}

// ambient.c:229-297 (aa = 0 branch) + ambcomp.c:350-422
__device__ __forceinline__ void multambient(const WaveArgs& A, RayCtx& r, const float aval[3], const double nrm[3]) {
    const DParams& P = A.P;
    bool dumb = (P.ambdiv <= 0) | (r.rdepth >= P.ambounce);
    float d = max3(aval);
    if (!dumb && d <= (float)RB_FTINY) dumb = true;      // samp_hemi insignificance -> !ok
    if (dumb) {                                          // dumbamb: global ambient value
        if (A.vacc && (P.ambval[0] > 0.f || P.ambval[1] > 0.f || P.ambval[2] > 0.f)) {
            float c[3] = {r.coef[0] * aval[0], r.coef[1] * aval[1], r.coef[2] * aval[2]};
            add_value(A, r.row, c, P.ambval[0], P.ambval[1], P.ambval[2]);
        }
        return;
    }
    double rdot = dot3(nrm, r.ron);
    int sgn = 1 - 2 * (rdot < 0);
    double wt = (double)r.rweight * sgn;
    bool backside = (wt < 0);
    if (backside) wt = -wt;
    double dd = (double)d;
    dd *= 0.8 * (double)r.rweight / ((double)P.ambdiv * (double)P.minweight + 1e-20);
    if (wt > dd) wt = dd;
    int n = (int)(sqrt(P.ambdiv * wt) + 0.5);
    if (n < 1) n = 1;
    float sc = (float)(1.0 / ((double)n * n));
    float acoef[3] = {aval[0] * sc, aval[1] * sc, aval[2] * sc};
    int atyp = backside ? RT_TAMBIENT : RT_RAMBIENT;
    double onrm[3] = {r.ron[0], r.ron[1], r.ron[2]};
    if (backside) { onrm[0] = -onrm[0]; onrm[1] = -onrm[1]; onrm[2] = -onrm[2]; }
    unsigned long long hkey = child_key(r.key, r.nchild++);
    if (n * n > A.inline_hemi_max) {                    // defer: expanded by k_expand
        unsigned slot = reserve_slot(&A.C->nh_out);
        if (slot >= A.hcap) { A.C->overflow = 1; return; }
        atomicAdd(&A.C->hemi_rays, (unsigned)(n * n));
        QHemi h;
        for (int k = 0; k < 3; k++) { h.rop[k] = r.rop[k]; h.onrm[k] = onrm[k]; h.acoef[k] = acoef[k]; h.ccoef[k] = r.coef[k]; }
        h.rweight = r.re > 0.1f ? ext_weight(r.rweight, r.re) : r.rweight;
        h.n = n; h.row = r.row;
        h.info = pack_info(r.crtype, r.rlvl, r.rdepth);
        h.key_lo = (unsigned)hkey; h.key_hi = (unsigned)(hkey >> 32);
        h.atype = atyp | (int)(r.med << 10); h.rsrc = r.rsrc;
        h.rmax_rem = (r.rmax > RB_FTINY) * (r.rmax - r.rot);
        A.hout[slot] = h;
        return;
    }
    double ux[3], uy[3];
    if (!getperpendicular_rand(ux, onrm, hkey)) return;
    uy[0] = onrm[1] * ux[2] - onrm[2] * ux[1];
    uy[1] = onrm[2] * ux[0] - onrm[0] * ux[2];
    uy[2] = onrm[0] * ux[1] - onrm[1] * ux[0];
    // the samples are children of the hemisphere's key, numbered from 0: the ray stands in for its own copy
    // (rayorigin() reads the parent and only counts its children)
    const unsigned long long rkey = r.key;
    const unsigned rn = r.nchild;
    r.key = hkey; r.nchild = 0;
    for (int i = n; i--;)
        for (int j = n; j--;) {
            QRay q;
            if (ambsample(P, r, atyp, acoef, onrm, ux, uy, n, i, j, q)) push_ray(A, q);
        }
    r.key = rkey; r.nchild = rn;
}

// raytrace.c:196-207 raytrans(): continue the ray unchanged
__device__ __forceinline__ void raytrans(const WaveArgs& A, RayCtx& r) {
    QRay q; float rc[3];
    if (!rayorigin(A.P, r, RT_TRANS, rc, false, q)) return;
    q.dir[0] = r.dir[0]; q.dir[1] = r.dir[1]; q.dir[2] = r.dir[2];
    push_ray(A, q);
}

// normal.c:176-360.  a[] = material reals; mkind = MK_PLASTIC / MK_METAL / MK_TRANS.
// RAY.pert of a smooth mesh triangle (o_mesh.c:193-209): barycentric weights of the hit point
// (tmesh.c:96-112 eval_baryc on the matrix the loader stored behind the vertices), interpolated vertex normal,
// normalised, minus the face normal.  Computed where a material needs it instead of being carried by every ray;
// `flipped` says flipsurface() (raytrace.c) has already reversed r.ron, in which case it reverses pert too.
__device__ __noinline__ bool smooth_pert(const DScene& S, int robj, const double rop[3], const double ron[3], bool flipped,
                                         double pert[3]) {
    const double* g = S.geom + __ldg(&S.objhdr[robj]).w;
    int i = (int)g[12] + 1;
    if (i >= 3) i -= 3;
    const double u = rop[i];
    if (++i >= 3) i -= 3;
    const double v = rop[i];
    double wt[3];
    wt[0] = u * g[13] + v * g[14] + g[15];
    wt[1] = u * g[16] + v * g[17] + g[18];
    wt[2] = 1. - wt[1] - wt[0];
    for (int k = 0; k < 3; k++) pert[k] = wt[0] * g[19 + k] + wt[1] * g[22 + k] + wt[2] * g[25 + k];
    const double sgn = flipped ? -1.0 : 1.0;
    if (normalize3(pert) != 0.0)
        for (int k = 0; k < 3; k++) pert[k] = pert[k] - sgn * ron[k];
    if (flipped)
        for (int k = 0; k < 3; k++) pert[k] = -pert[k];
    return dot3(pert, pert) > RB_FTINY * RB_FTINY;
}
__device__ __forceinline__ bool ray_pert(const DScene& S, const RayCtx& r, bool flipped, double pert[3]) {
    if (!(r.xfl & 1) || r.robj < 0) return false;
    return smooth_pert(S, r.robj, r.rop, r.ron, flipped, pert);
}

// raytrace.c:445-478 raynormal()
__device__ __forceinline__ double raynormal(double norm[3], const RayCtx& r, const double pert[3]) {
    for (int i = 0; i < 3; i++) norm[i] = r.ron[i] + pert[i];
    if (normalize3(norm) == 0.0) {
        for (int i = 0; i < 3; i++) norm[i] = r.ron[i];
        return r.rod;
    }
    double newdot = -dot3(norm, r.dir);
    if ((newdot > 0.0) ^ (r.rod > 0.0)) {
        for (int i = 0; i < 3; i++) norm[i] += 2.0 * newdot * r.dir[i];
        newdot = -newdot;
    }
    return newdot;
}

// FAST (k_shade_mid / k_shade_fast): the caller guarantees a PURE-specular material (roughness^2 <= FTINY) on a surface
// without vertex normals, so the sampled-highlight code and the normal perturbation are compiled out; LEAN
// (k_shade_fast) also guarantees that the material is not `trans`, and that code goes too.  Everything an
// instantiation does execute is the same code, on the same values, as the general one.
// NODIRECT (k_shade_spec): the scene has no source for direct() to sample, so the call and what only it needs go.
template <bool FAST = false, bool LEAN = false, bool NODIRECT = false>
__device__ __forceinline__ void m_normal(const WaveArgs& A, RayCtx& r, int mkind_, const float* a) {
    const DParams& P = A.P;
    const int mkind = LEAN ? (mkind_ == MK_METAL ? MK_METAL : MK_PLASTIC) : mkind_;
    if ((r.crtype & RT_SHADOW) && mkind != MK_TRANS) return;      // easy shadow test
    bool flipped = false;
    if (r.rod < 0.0) {
        if (!P.backvis) { raytrans(A, r); return; }
        r.rod = -r.rod; r.ron[0] = -r.ron[0]; r.ron[1] = -r.ron[1]; r.ron[2] = -r.ron[2];   // flipsurface
        flipped = true;
    }
    NormDat nd;
    nd.mcolor[0] = a[0]; nd.mcolor[1] = a[1]; nd.mcolor[2] = a[2];
    nd.specfl = 0;
    nd.alpha2 = a[4];
    if ((nd.alpha2 *= nd.alpha2) <= RB_FTINY) nd.specfl |= SP_PURE;
    nd.pnorm[0] = r.ron[0]; nd.pnorm[1] = r.ron[1]; nd.pnorm[2] = r.ron[2];
    nd.pdot = r.rod;
    double pert[3] = {0., 0., 0.};
    const bool hastexture = FAST ? false : ray_pert(A.S, r, flipped, pert);       // normal.c:221-226
    if (hastexture) nd.pdot = raynormal(nd.pnorm, r, pert);
    if (!hastexture && r.robj >= 0 && r.flat) nd.specfl |= SP_FLAT;
    if (nd.pdot < .001) nd.pdot = .001;
    nd.rspec = a[3];
    double fest = 0.;
    if ((nd.specfl & SP_PURE) && nd.rspec >= 0.017999) {
        fest = exp(-5.85 * nd.pdot) - 0.00202943064;
        nd.rspec += fest * (1. - nd.rspec);
    }
    if (!LEAN && mkind == MK_TRANS) {
        nd.trans = a[5] * (1.0 - nd.rspec);
        nd.tspec = nd.trans * a[6];
        nd.tdiff = nd.trans - nd.tspec;
        if (nd.tspec > RB_FTINY) {
            nd.specfl |= SP_TRAN;
            if (!(nd.specfl & SP_PURE) && P.specthresh >= nd.tspec - RB_FTINY) nd.specfl |= SP_TBLT;
            nd.prdir[0] = r.dir[0]; nd.prdir[1] = r.dir[1]; nd.prdir[2] = r.dir[2];
            if (hastexture && !(r.crtype & (RT_SHADOW | RT_AMBIENT)) && !(r.xfl & 2)) {      // normal.c:251-262
                double pd[3] = {r.dir[0] - pert[0], r.dir[1] - pert[1], r.dir[2] - pert[2]};
                if (dot3(pd, r.ron) < -RB_FTINY) {
                    normalize3(pd);
                    nd.prdir[0] = pd[0]; nd.prdir[1] = pd[1]; nd.prdir[2] = pd[2];
                }
            }
        }
    } else
        nd.tdiff = nd.tspec = nd.trans = 0.0;
    nd.rdiff = 1.0 - nd.trans - nd.rspec;
    // transmitted ray
    if (!LEAN && (nd.specfl & (SP_TRAN | SP_PURE | SP_TBLT)) == (SP_TRAN | SP_PURE)) {
        float rc[3] = {(float)(nd.mcolor[0] * nd.tspec), (float)(nd.mcolor[1] * nd.tspec), (float)(nd.mcolor[2] * nd.tspec)};
        QRay q;
        if (rayorigin(P, r, RT_TRANS, rc, true, q)) {
            q.dir[0] = nd.prdir[0]; q.dir[1] = nd.prdir[1]; q.dir[2] = nd.prdir[2];
            push_ray(A, q);
        }
    }
    if (r.crtype & RT_SHADOW) return;
    nd.scolor[0] = nd.scolor[1] = nd.scolor[2] = 0.f;
    if (nd.rspec > RB_FTINY) {
        nd.specfl |= SP_REFL;
        if (mkind != MK_METAL) nd.scolor[0] = nd.scolor[1] = nd.scolor[2] = (float)nd.rspec;
        else if (fest > RB_FTINY) {
            double d = a[3] * (1. - fest);
            for (int k = 0; k < 3; k++) nd.scolor[k] = (float)(fest + nd.mcolor[k] * d);
        } else
            for (int k = 0; k < 3; k++) nd.scolor[k] = (float)(nd.mcolor[k] * nd.rspec);
        if (!(nd.specfl & SP_PURE) && P.specthresh >= nd.rspec - RB_FTINY) nd.specfl |= SP_RBLT;
    }
    // reflected ray
    if ((nd.specfl & (SP_REFL | SP_PURE | SP_RBLT)) == (SP_REFL | SP_PURE)) {
        float rc[3] = {nd.scolor[0], nd.scolor[1], nd.scolor[2]};
        QRay q;
        if (rayorigin(P, r, RT_REFLECTED, rc, true, q)) {
            for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + nd.pnorm[k] * (2. * nd.pdot);
            if (hastexture && dot3(q.dir, r.ron) <= RB_FTINY)          // penetration? (normal.c:314-316)
                for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + r.ron[k] * (2. * r.rod);
            normalize3(q.dir);
            push_ray(A, q);
        }
    }
    if ((nd.specfl & SP_PURE) && nd.rdiff <= RB_FTINY && nd.tdiff <= RB_FTINY) return;
    if (!FAST && !(nd.specfl & SP_PURE)) {
        // gaussamp(), normal.c:363-495, single-sample form (-ss <= 1.5)
        unsigned long long gkey = child_key(r.key, r.nchild++);
        double u[3], v[3];
        if (getperpendicular_rand(u, nd.pnorm, gkey)) {
            v[0] = nd.pnorm[1] * u[2] - nd.pnorm[2] * u[1];
            v[1] = nd.pnorm[2] * u[0] - nd.pnorm[0] * u[2];
            v[2] = nd.pnorm[0] * u[1] - nd.pnorm[1] * u[0];
            if ((nd.specfl & (SP_REFL | SP_RBLT)) == SP_REFL) {
                float rc[3] = {nd.scolor[0], nd.scolor[1], nd.scolor[2]};
                QRay q;
                if (rayorigin(P, r, RT_RSPECULAR, rc, true, q)) {
                    for (int ntr = 0; ntr < 10; ntr++) {
                        double rv0 = rnd01(gkey, 30 + 2 * ntr), rv1 = rnd01(gkey, 31 + 2 * ntr);
                        double s, c; sincos(2.0 * RB_PI * rv0, &s, &c);
                        if ((0. <= P.specjitter) & (P.specjitter < 1.)) rv1 = 1.0 - P.specjitter * rv1;
                        double d = (rv1 <= RB_FTINY) ? 1.0 : sqrt(nd.alpha2 * -log(rv1));
                        double h[3];
                        for (int k = 0; k < 3; k++) h[k] = nd.pnorm[k] + d * (c * u[k] + s * v[k]);
                        d = -2.0 * dot3(h, r.dir) / (1.0 + d * d);
                        for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + h[k] * d;
                        if (dot3(q.dir, r.ron) <= RB_FTINY) continue;
                        normalize3(q.dir);
                        push_ray(A, q);
                        break;
                    }
                }
            }
            if ((nd.specfl & (SP_TRAN | SP_TBLT)) == SP_TRAN) {
                float rc[3] = {(float)(nd.mcolor[0] * nd.tspec), (float)(nd.mcolor[1] * nd.tspec), (float)(nd.mcolor[2] * nd.tspec)};
                QRay q;
                if (rayorigin(P, r, RT_TSPECULAR, rc, true, q)) {
                    for (int ntr = 0; ntr < 10; ntr++) {
                        double rv0 = rnd01(gkey, 60 + 2 * ntr), rv1 = rnd01(gkey, 61 + 2 * ntr);
                        double s, c; sincos(2.0 * RB_PI * rv0, &s, &c);
                        if ((0. <= P.specjitter) & (P.specjitter < 1.)) rv1 = 1.0 - P.specjitter * rv1;
                        double d = (rv1 <= RB_FTINY) ? 1.0 : sqrt(nd.alpha2 * -log(rv1));
                        for (int k = 0; k < 3; k++) q.dir[k] = nd.prdir[k] + d * (c * u[k] + s * v[k]);
                        if (dot3(q.dir, r.ron) >= -RB_FTINY) continue;
                        normalize3(q.dir);
                        push_ray(A, q);
                        break;
                    }
                }
            }
        }
    }
    if (nd.rdiff > RB_FTINY) {
        float sct[3];
        for (int k = 0; k < 3; k++) sct[k] = (float)(nd.mcolor[k] * nd.rdiff);
        if (nd.specfl & SP_RBLT) for (int k = 0; k < 3; k++) sct[k] += nd.scolor[k];
        multambient(A, r, sct, nd.pnorm);
    }
    if (!LEAN && nd.tdiff > RB_FTINY) {
        float sct[3];
        double f = (nd.specfl & SP_TBLT) ? nd.trans : nd.tdiff;
        for (int k = 0; k < 3; k++) sct[k] = (float)(nd.mcolor[k] * f);
        double bn[3] = {-nd.pnorm[0], -nd.pnorm[1], -nd.pnorm[2]};
        multambient(A, r, sct, bn);
    }
    if (!NODIRECT) direct_or_park<FAST>(A, r, nd);
}

// fvect.c:159-196 getperpendicular() with randomize = 0
__device__ __forceinline__ bool getperpendicular0(double vp[3], const double v[3]) {
    int i;
    for (i = 3; i--;)
        if ((-0.6 < v[i]) & (v[i] < 0.6)) break;
    if (i < 0) return false;
    const double v1[3] = {i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0};
    vp[0] = v1[1] * v[2] - v1[2] * v[1];
    vp[1] = v1[2] * v[0] - v1[0] * v[2];
    vp[2] = v1[0] * v[1] - v1[1] * v[0];
    return normalize3(vp) > 0.0;
}

// aniso.c:185-297 m_aniso(), :299-326 getacoords(), :329-470 agaussamp() in its single-sample form (-ss <= 1.5).
// a[] = material reals (6, or 8 for trans2), uvec = orientation vector with the function transform applied by the
// loader.  Out of line: these materials are rare, and the isotropic path's code stays what it was.
__device__ __noinline__ void m_aniso(const WaveArgs& A, RayCtx& r, int mkind, const float* a, const double* uvec) {
    const DParams& P = A.P;
    if (r.crtype & RT_SHADOW) return;                    // easy shadow test (trans2 included)
    bool flipped = false;
    if (r.rod < 0.0) {
        if (!P.backvis) { raytrans(A, r); return; }
        r.rod = -r.rod; r.ron[0] = -r.ron[0]; r.ron[1] = -r.ron[1]; r.ron[2] = -r.ron[2];   // flipsurface
        flipped = true;
    }
    NormDat nd;
    nd.specfl = SP_ANISO;
    nd.alpha2 = 0.0;
    for (int k = 0; k < 3; k++) { nd.mcolor[k] = a[k]; nd.scolor[k] = 0.f; nd.pnorm[k] = r.ron[k]; nd.prdir[k] = r.dir[k]; nd.u[k] = uvec[k]; }
    nd.u_alpha = a[4]; nd.v_alpha = a[5];
    nd.pdot = r.rod;
    double pert[3];
    const bool hastexture = ray_pert(A.S, r, flipped, pert);
    if (hastexture) nd.pdot = raynormal(nd.pnorm, r, pert);
    if (nd.pdot < .001) nd.pdot = .001;
    if ((nd.rspec = a[3]) > RB_FTINY) {
        nd.specfl |= SP_REFL;
        for (int k = 0; k < 3; k++) nd.scolor[k] = (float)((mkind == MK_METAL2 ? nd.mcolor[k] : 1.f) * nd.rspec);
        if (P.specthresh >= nd.rspec - RB_FTINY) nd.specfl |= SP_RBLT;
    }
    if (mkind == MK_TRANS2) {
        nd.trans = a[6] * (1.0 - nd.rspec);
        nd.tspec = nd.trans * a[7];
        nd.tdiff = nd.trans - nd.tspec;
        if (nd.tspec > RB_FTINY) {
            nd.specfl |= SP_TRAN;
            if (P.specthresh >= nd.tspec - RB_FTINY) nd.specfl |= SP_TBLT;
            if (hastexture && !(r.xfl & 2)) {            // aniso.c:241-252: bent transmission, not under "Phong"
                double pd[3] = {r.dir[0] - pert[0], r.dir[1] - pert[1], r.dir[2] - pert[2]};
                if (dot3(pd, r.ron) < -RB_FTINY) {
                    normalize3(pd);
                    nd.prdir[0] = pd[0]; nd.prdir[1] = pd[1]; nd.prdir[2] = pd[2];
                }
            }
        }
    } else
        nd.tdiff = nd.tspec = nd.trans = 0.0;
    nd.rdiff = 1.0 - nd.trans - nd.rspec;
    if (!hastexture && r.robj >= 0 && r.flat) nd.specfl |= SP_FLAT;
    // getacoords(): v = n x u, u = v x n; an orientation along the normal falls back to an isotropic lobe
    nd.v[0] = nd.pnorm[1] * nd.u[2] - nd.pnorm[2] * nd.u[1];
    nd.v[1] = nd.pnorm[2] * nd.u[0] - nd.pnorm[0] * nd.u[2];
    nd.v[2] = nd.pnorm[0] * nd.u[1] - nd.pnorm[1] * nd.u[0];
    if (normalize3(nd.v) == 0.0) {
        getperpendicular0(nd.u, nd.pnorm);
        nd.v[0] = nd.pnorm[1] * nd.u[2] - nd.pnorm[2] * nd.u[1];
        nd.v[1] = nd.pnorm[2] * nd.u[0] - nd.pnorm[0] * nd.u[2];
        nd.v[2] = nd.pnorm[0] * nd.u[1] - nd.pnorm[1] * nd.u[0];
        nd.u_alpha = nd.v_alpha = sqrt(0.5 * (nd.u_alpha * nd.u_alpha + nd.v_alpha * nd.v_alpha));
    } else {
        const double ux = nd.v[1] * nd.pnorm[2] - nd.v[2] * nd.pnorm[1];
        const double uy = nd.v[2] * nd.pnorm[0] - nd.v[0] * nd.pnorm[2];
        const double uz = nd.v[0] * nd.pnorm[1] - nd.v[1] * nd.pnorm[0];
        nd.u[0] = ux; nd.u[1] = uy; nd.u[2] = uz;
    }
    if (nd.specfl & (SP_REFL | SP_TRAN)) {
        const unsigned long long gkey = child_key(r.key, r.nchild++);
        const double ua2 = nd.u_alpha * nd.u_alpha, va2 = nd.v_alpha * nd.v_alpha;
        for (int side = 0; side < 2; side++) {           // 0: reflected lobe, 1: transmitted lobe
            if (side == 0 ? (nd.specfl & (SP_REFL | SP_RBLT)) != SP_REFL : (nd.specfl & (SP_TRAN | SP_TBLT)) != SP_TRAN) continue;
            float rc[3];
            for (int k = 0; k < 3; k++) rc[k] = side == 0 ? nd.scolor[k] : (float)(nd.mcolor[k] * nd.tspec);
            QRay q;
            if (!rayorigin(P, r, side == 0 ? RT_RSPECULAR : RT_TSPECULAR, rc, true, q)) continue;
            for (int ntr = 0; ntr < 10; ntr++) {
                const unsigned dim = (side == 0 ? 30 : 60) + 2 * ntr;
                double rv1 = rnd01(gkey, dim + 1);
                double sinp, cosp;
                sincos(2.0 * RB_PI * rnd01(gkey, dim), &sinp, &cosp);
                cosp *= nd.u_alpha; sinp *= nd.v_alpha;
                double d = 1. / sqrt(cosp * cosp + sinp * sinp);
                cosp *= d; sinp *= d;
                if ((0. <= P.specjitter) & (P.specjitter < 1.)) rv1 = 1.0 - P.specjitter * rv1;
                d = (rv1 <= RB_FTINY) ? 1.0 : sqrt(-log(rv1) / (cosp * cosp / ua2 + sinp * sinp / va2));
                if (side == 0) {
                    double h[3];
                    for (int k = 0; k < 3; k++) h[k] = nd.pnorm[k] + d * (cosp * nd.u[k] + sinp * nd.v[k]);
                    d = -2.0 * dot3(h, r.dir) / (1.0 + d * d);
                    for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + h[k] * d;
                    if (dot3(q.dir, r.ron) <= RB_FTINY) continue;      // sample rejection test
                } else {
                    for (int k = 0; k < 3; k++) q.dir[k] = nd.prdir[k] + d * (cosp * nd.u[k] + sinp * nd.v[k]);
                    if (dot3(q.dir, r.ron) >= -RB_FTINY) continue;
                }
                normalize3(q.dir);
                push_ray(A, q);
                break;
            }
        }
    }
    if (nd.rdiff > RB_FTINY) {
        float sct[3];
        for (int k = 0; k < 3; k++) sct[k] = (float)(nd.mcolor[k] * nd.rdiff);
        if (nd.specfl & SP_RBLT) for (int k = 0; k < 3; k++) sct[k] += nd.scolor[k];
        multambient(A, r, sct, nd.pnorm);
    }
    if (nd.tdiff > RB_FTINY) {
        float sct[3];
        const double f = (nd.specfl & SP_TBLT) ? nd.trans : nd.tdiff;
        for (int k = 0; k < 3; k++) sct[k] = (float)(nd.mcolor[k] * f);
        const double bn[3] = {-nd.pnorm[0], -nd.pnorm[1], -nd.pnorm[2]};
        multambient(A, r, sct, bn);
    }
    direct_or_park(A, r, nd);
}

// dielectric.c:69-251 m_dielectric() (built without DISPERSE, like the reference): Fresnel reflection and
// refraction at a dielectric / interface, a[] = reals (5 / 8).  The refracted ray leaves in the medium on
// the other side; the ray itself is charged with the medium it came through (ray_medium).
__device__ __noinline__ void m_dielectric(const WaveArgs& A, RayCtx& r, int mkind, int slot, const float* a) {
    const DParams& P = A.P;
    const bool iface = mkind == MK_INTERFACE;
    double dnorm[3] = {r.ron[0], r.ron[1], r.ron[2]};
    double cos1 = r.rod;
    double pert[3];
    int hastexture = ray_pert(A.S, r, false, pert) ? 1 : 0;
    if (hastexture) cos1 = raynormal(dnorm, r, pert);
    double nratio = iface ? (double)a[3] / (double)a[7] : (double)a[3] + (double)a[4] / 500.;   // Hartmann, mean lambda
    unsigned mtrans;                                 // medium of the refracted ray
    if (cos1 < 0.0) {                                // inside
        hastexture = -hastexture;
        cos1 = -cos1;
        dnorm[0] = -dnorm[0]; dnorm[1] = -dnorm[1]; dnorm[2] = -dnorm[2];
        ray_medium(A, r, (int)medium_id(slot, 0));
        mtrans = iface ? medium_id(slot, 1) : 0u;
    } else {                                         // outside
        nratio = 1.0 / nratio;
        mtrans = medium_id(slot, 0);
        if (iface) ray_medium(A, r, (int)medium_id(slot, 1));     // (a plain dielectric hit from outside: k_shade has charged r.med)
    }
    double d2 = 1.0 - nratio * nratio * (1.0 - cos1 * cos1);       // cos theta2 squared
    double refl;
    if (d2 < RB_FTINY) refl = 1.0;                   // total reflection
    else {
        const double cos2 = sqrt(d2);
        double d1 = cos1;
        d2 = nratio * cos2;
        d1 = (d1 - d2) / (d1 + d2);
        refl = d1 * d1;
        d1 = 1.0 / cos1;
        d2 = nratio / cos2;
        d1 = (d1 - d2) / (d1 + d2);
        refl += d1 * d1;
        refl *= 0.5;
        const double trans = (1.0 - refl) * nratio * nratio;       // solid angle ratio
        float rc[3] = {(float)trans, (float)trans, (float)trans};
        QRay q;
        if (rayorigin(P, r, RT_REFRACTED, rc, true, q)) {
            d1 = nratio * cos1 - cos2;
            for (int k = 0; k < 3; k++) q.dir[k] = nratio * r.dir[k] + d1 * dnorm[k];
            if (hastexture && dot3(q.dir, r.ron) * hastexture >= -RB_FTINY) {     // accidental reflection: ignore texture
                d1 *= (double)hastexture;
                for (int k = 0; k < 3; k++) q.dir[k] = nratio * r.dir[k] + d1 * r.ron[k];
            }
            normalize3(q.dir);
            q.med = mtrans;
            push_ray(A, q);
        }
    }
    if (!(r.crtype & RT_SHADOW)) {
        float rc[3] = {(float)refl, (float)refl, (float)refl};
        QRay q;
        if (rayorigin(P, r, RT_REFLECTED, rc, true, q)) {
            for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + dnorm[k] * (2. * cos1);
            if (hastexture && dot3(q.dir, r.ron) * hastexture <= RB_FTINY)        // accidental penetration
                for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + r.ron[k] * (2. * r.rod);
            normalize3(q.dir);
            push_ray(A, q);
        }
    }
}

// ---- BSDF / aBSDF (rt/m_bsdf.c) on Klems-matrix data ----
// m_bsdf.c:116-219 compute_through(): the "through" (unscattered) component an aBSDF lets view and shadow rays see.
__device__ __noinline__ void bsdf_compute_through(const BsdfRef& B, const NormDat& nd, bool hitfront, float cthru[3], float cthru_surr[3]) {
    const float dir2check[29][2] = {
        {0, 0}, {-0.6f, 0}, {0, 0.6f}, {0, -0.6f}, {0.6f, 0}, {-0.6f, 0.6f}, {-0.6f, -0.6f}, {0.6f, 0.6f}, {0.6f, -0.6f},
        {-1.2f, 0}, {0, 1.2f}, {0, -1.2f}, {1.2f, 0}, {-1.2f, 1.2f}, {-1.2f, -1.2f}, {1.2f, 1.2f}, {1.2f, -1.2f}, {-1.8f, 0},
        {0, 1.8f}, {0, -1.8f}, {1.8f, 0}, {-1.8f, 1.8f}, {-1.8f, -1.8f}, {1.8f, 1.8f}, {1.8f, -1.8f}, {-2.4f, 0}, {0, 2.4f},
        {0, -2.4f}, {2.4f, 0}};
    const BsdfRec& R = *B.rec;
    const int tk = bsdf_tcomp(R, hitfront);
    if (tk < 0) return;                              // no specular transmission
    const double minProjSA = R.c[tk].minProjSA;
    const double srchrad = sqrt(minProjSA);          // evaluate peak
    const double* vray = nd.prdir;
    double vy[29]; signed char ord[29];
    auto tdir_of = [&](int i, double t[3]) {
        t[0] = -vray[0] + (double)dir2check[i][0] * srchrad;
        t[1] = -vray[1] + (double)dir2check[i][1] * srchrad;
        t[2] = -vray[2];
        bsdf_normalize(t);
    };
    for (int i = 0; i < 29; i++) {
        double t[3];
        tdir_of(i, t);
        vy[i] = sd_eval(B, vray, t);
        int j = i;                                   // near-peak values in descending order (stable)
        while (j > 0 && vy[ord[j - 1]] < vy[i]) { ord[j] = ord[j - 1]; j--; }
        ord[j] = (signed char)i;
    }
    if (vy[ord[0]] <= RB_FTINY) return;              // zero BTDF here
    float vpeak = 0.f, vsurr = 0.f;                  // grayscale data: one channel, expanded at the end
    double vypeak = 0, tomsum = 0, tomsurr = 0;
    int ns = 0;
    for (int i = 0; i < 29; i++) {                   // combine top unique values
        const double y = vy[ord[i]];
        if (i && y == vy[ord[i - 1]]) continue;      // assume duplicate sample
        double t[3], tomega[2];
        tdir_of(ord[i], t);
        sd_size(B, tomega, vray, t, SDQ_MIN);
        float vcol = (float)y;
        vcol = (float)(vcol * tomega[0]);
        if (tomega[0] > 1.5 * minProjSA || vypeak > 8. * y * ns) {      // not part of peak?
            if (!i) return;                          // abort
            vsurr += vcol;
            tomsurr += tomega[0];
            continue;
        }
        vpeak += vcol;
        tomsum += tomega[0];
        vypeak += y;
        ++ns;
    }
    if (tomsurr < 0.2 * tomsum) return;              // insufficient surround?
    vsurr = (float)(vsurr * (1. / tomsurr));         // surround is avg. BTDF
    float btdiff = (float)(vray[2] > 0 ? R.lamb[2] : R.lamb[3]);      // get diffuse BTDF
    btdiff = (float)(btdiff * (1. / RB_PI));
    if ((vpeak -= (float)(tomsum * btdiff)) < 0) vpeak = 0;            // remove diffuse contrib.
    if ((vsurr -= btdiff) < 0) vsurr = 0;
    if (vpeak < .0005f) return;                      // < 0.05% specular?
    bsdf_gray(cthru_surr, vsurr);
    bsdf_gray(cthru, vpeak);
}

// m_bsdf.c:222-233 bsdf_jitter()
__device__ __forceinline__ void bsdf_jitter(double vres[3], const double vray[3], double sr_psa, double specjitter,
                                            unsigned long long key, unsigned dim) {
    vres[0] = vray[0]; vres[1] = vray[1]; vres[2] = vray[2];
    if (specjitter < 1.) sr_psa *= specjitter;
    if (sr_psa <= RB_FTINY) return;
    vres[0] += sr_psa * (.5 - rnd01(key, dim));
    vres[1] += sr_psa * (.5 - rnd01(key, dim + 1));
    bsdf_normalize(vres);
}

// m_bsdf.c:236-342 direct_specular_OK(): the BSDF's non-diffuse value towards a light source
__device__ __noinline__ bool bsdf_direct_specular(float scval[3], const BsdfRef& B, const NormDat& np, const RayCtx& r,
                                                  const double ldir[3], double omega, double specjitter, unsigned long long jkey) {
    scval[0] = scval[1] = scval[2] = 0.f;
    const double toloc[3][3] = {{np.u[0], np.u[1], np.u[2]}, {np.v[0], np.v[1], np.v[2]}, {np.w[0], np.w[1], np.w[2]}};
    const double* vray = np.prdir;
    const BsdfRec& R = *B.rec;
    double vsrc[3];
    if (!sd_map_dir(vsrc, toloc, ldir)) return false;
    if (((vsrc[2] > 0) ^ (vray[2] > 0)) && max3(np.cthru) > (float)RB_FTINY) {      // check indirect over-counting
        const double dx = vsrc[0] + vray[0], dy = vsrc[1] + vray[1];
        const int tk = bsdf_tcomp(R, r.rod > 0);
        const double mpsa = R.c[tk].minProjSA;
        const double tomega = omega * fabs(vsrc[2]);
        if (dx * dx + dy * dy <= (2.5 * 4. / RB_PI) * (tomega + mpsa + 2. * sqrt(tomega * mpsa))) {
            if (max3(np.cthru_surr) <= (float)RB_FTINY) return false;
            scval[0] = np.cthru_surr[0]; scval[1] = np.cthru_surr[1]; scval[2] = np.cthru_surr[2];
            return true;                             // return non-zero surround BTDF
        }
    }
    double svY;                                      // will discount diffuse portion
    const bool anyt = R.c[BC_TF].present | R.c[BC_TB].present;
    switch ((vsrc[2] > 0) << 1 | (vray[2] > 0)) {
    case 3: if (!R.c[BC_RF].present) return false; svY = R.lamb[0]; break;
    case 0: if (!R.c[BC_RB].present) return false; svY = R.lamb[1]; break;
    case 1: if (!anyt) return false; svY = R.lamb[2]; break;
    default: if (!anyt) return false; svY = R.lamb[3]; break;
    }
    double diffY = 0;
    float cdiff[3] = {0.f, 0.f, 0.f};
    if (svY > RB_FTINY) { diffY = svY *= 1. / RB_PI; bsdf_gray(cdiff, svY); }
    double tomega[2];
    sd_size(B, tomega, vray, vsrc, SDQ_MIN);
    int nsamp = 1, scnt = 0;
    const double tsr = sqrt(tomega[0]);
    if (tsr > 0) {                                   // check if sampling BSDF
        nsamp = (int)(4. * specjitter * r.rweight + .5);
        nsamp += !nsamp;
    }
    for (int i = nsamp; i--;) {                      // jitter to fuzz BSDF cells
        double vjit[3];
        bsdf_jitter(vjit, vray, tsr, specjitter, jkey, 40u + 2u * (unsigned)i);
        const double y = sd_eval(B, vjit, vsrc);
        if (y - diffY <= RB_FTINY) { ++scnt; continue; }      // still counts as 0 contribution
        double tomega2[2];
        sd_size(B, tomega2, vjit, vsrc, SDQ_MIN);              // check for variable resolution
        if (tomega2[0] < .12 * tomega[0]) continue;            // not safe to include
        float csmp[3];
        bsdf_gray(csmp, y);
        scval[0] += csmp[0]; scval[1] += csmp[1]; scval[2] += csmp[2];
        ++scnt;
    }
    if (!scnt) return false;                         // no valid specular samples?
    for (int k = 0; k < 3; k++) scval[k] = (float)(scval[k] * (1. / scnt));       // weighted average BSDF
    if (diffY > RB_FTINY)                            // subtract diffuse contribution
        for (int k = 0; k < 3; k++) if ((scval[k] -= cdiff[k]) < 0) scval[k] = 0;
    return true;
}

// m_bsdf.c:345-482 dir_bsdf() / dir_brdf() / dir_btdf(), chosen by np.dmode (patterns are not built: pcol = 1)
__device__ __noinline__ void dir_bsdf(float scval[3], const WaveArgs& A, const NormDat& np, const RayCtx& r, const double ldir[3],
                                      double omega, unsigned long long jkey) {
    scval[0] = scval[1] = scval[2] = 0.f;
    const double ldot = dot3(np.pnorm, ldir);
    if (np.dmode == 0) { if ((-RB_FTINY <= ldot) & (ldot <= RB_FTINY)) return; }
    else if (np.dmode == 1) { if (ldot <= RB_FTINY) return; }
    else if (ldot >= -RB_FTINY) return;
    if (np.dmode != 2 && ldot > 0 && max3(np.mcolor) > (float)RB_FTINY) {           // diffuse reflected component
        const double d = ldot * omega * (1. / RB_PI);
        for (int k = 0; k < 3; k++) scval[k] += (float)(np.mcolor[k] * d);
    }
    if (np.dmode != 1 && ldot < 0 && max3(np.scolor) > (float)RB_FTINY) {           // diffuse transmission
        const double d = -ldot * omega * (1. / RB_PI);
        for (int k = 0; k < 3; k++) scval[k] += (float)(np.scolor[k] * d);
    }
    const BsdfRef B = {A.S.bsdfs + np.bsdf, A.S.bsdfbases, A.S.bsdfpool};
    float sct[3];
    if (!bsdf_direct_specular(sct, B, np, r, ldir, omega, A.P.specjitter, jkey)) return;
    const double d = (ldot < 0 ? -ldot : ldot) * omega;
    for (int k = 0; k < 3; k++) scval[k] += (float)(sct[k] * d);
}

// m_bsdf.c:625-804 m_bsdf() with sample_sdf() :555-623 and sample_sdcomp() :484-553 in their single-sample form
// (-ss <= 1.5).  mkind = MK_BSDF (thickness: proxy surface) or MK_ABSDF (through component).
__device__ __noinline__ void m_bsdf(const WaveArgs& A, RayCtx& r, const MatRec& m) {
    const DParams& P = A.P;
    const bool hasthick = m.kind == MK_BSDF;
    const bool hitfront = r.rod > 0;
    if (!hitfront & !P.backvis) { raytrans(A, r); return; }       // check backface visibility
    const double thick = hasthick ? m.pad2 : 0.;
    if (thick != 0 && ((r.crtype & RT_SHADOW) || !(r.crtype & (RT_SPECULAR | RT_AMBIENT)) || ((thick > 0) ^ hitfront))) {
        raytrans(A, r);                              // hide our proxy
        return;
    }
    if (hasthick && (r.crtype & RT_SHADOW)) return;  // early shadow check #1
    const BsdfRef B = {A.S.bsdfs + m.pad[0], A.S.bsdfbases, A.S.bsdfpool};
    const BsdfRec& R = *B.rec;
    const bool anyt = R.c[BC_TF].present | R.c[BC_TB].present;
    if ((r.crtype & RT_SHADOW) && !anyt) return;     // early shadow check #2
    NormDat nd;
    nd.specfl = SP_BSDF; nd.bsdf = m.pad[0]; nd.dmode = 0;
    float a8 = 0.f;
    if (m.nargs >= 9) a8 = __int_as_float(m.pad[1]);
    // diffuse components (nd.mcolor = rdiff, nd.scolor = tdiff)
    bsdf_gray(nd.mcolor, hitfront ? R.lamb[0] : R.lamb[1]);
    if (hitfront) { if (m.nargs >= 3) for (int k = 0; k < 3; k++) nd.mcolor[k] += m.a[k]; }
    else if (m.nargs >= 6) for (int k = 0; k < 3; k++) nd.mcolor[k] += m.a[3 + k];
    bsdf_gray(nd.scolor, hitfront ? R.lamb[2] : R.lamb[3]);
    if (m.nargs >= 9) { nd.scolor[0] += m.a[6]; nd.scolor[1] += m.a[7]; nd.scolor[2] += a8; }
    // local BSDF coordinates
    double pert[3] = {0., 0., 0.};
    ray_pert(A.S, r, false, pert);
    raynormal(nd.pnorm, r, pert);
    double toloc[3][3], fromloc[3][3];
    bool ok = sd_comp_xform(toloc, nd.pnorm, m.u);
    if (ok) {
        const double nv[3] = {-r.dir[0], -r.dir[1], -r.dir[2]};
        ok = sd_map_dir(nd.prdir, toloc, nv);
    }
    if (!ok) return;                                 // "Illegal orientation vector" (a warning in the reference)
    for (int k = 0; k < 3; k++) { nd.u[k] = toloc[0][k]; nd.v[k] = toloc[1][k]; nd.w[k] = toloc[2][k]; }
    const double* vray = nd.prdir;
    for (int k = 0; k < 3; k++) nd.cthru[k] = nd.cthru_surr[k] = 0.f;
    if (m.kind == MK_ABSDF) {                        // consider through component
        bsdf_compute_through(B, nd, hitfront, nd.cthru, nd.cthru_surr);
        if (r.crtype & RT_SHADOW) {                  // attempt to pass shadow ray
            float rc[3] = {nd.cthru[0], nd.cthru[1], nd.cthru[2]};
            QRay q;
            if (rayorigin(P, r, RT_TRANS, rc, true, q)) {
                q.dir[0] = r.dir[0]; q.dir[1] = r.dir[1]; q.dir[2] = r.dir[2];
                push_ray(A, q);
            }
            return;
        }
    }
    if (!sd_inv_xform(fromloc, toloc)) return;
    double sr_vpsa[2];
    sd_size(B, sr_vpsa, vray, nullptr, SDQ_MIN + SDQ_MAX);        // determine BSDF resolution
    sr_vpsa[0] = sqrt(sr_vpsa[0]); sr_vpsa[1] = sqrt(sr_vpsa[1]);
    if (!hitfront) { nd.pnorm[0] = -nd.pnorm[0]; nd.pnorm[1] = -nd.pnorm[1]; nd.pnorm[2] = -nd.pnorm[2]; }   // perturb normal towards hit
    const unsigned long long bkey = child_key(r.key, r.nchild++);
    float unsamp[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};      // runsamp, tunsamp
    for (int xmit = 0; xmit < 2; xmit++) {           // sample_sdf(SDsampSpR), then sample_sdf(SDsampSpT)
        const int ck = xmit ? bsdf_tcomp(R, hitfront) : bsdf_rcomp(R, hitfront);
        if (ck < 0) continue;                        // no specular component?
        bool hasthru = xmit && !(r.crtype & (RT_SPECULAR | RT_AMBIENT)) && max3(nd.cthru) > (float)RB_FTINY;
        const bool hasthru0 = hasthru;
        double b = 0;
        if (hasthru) {                               // separate view sample?
            float rc[3] = {nd.cthru[0], nd.cthru[1], nd.cthru[2]};
            QRay q;
            if (rayorigin(P, r, RT_TRANS, rc, true, q)) {
                q.dir[0] = r.dir[0]; q.dir[1] = r.dir[1]; q.dir[2] = r.dir[2];
                push_ray(A, q);
                b = 0.2651058201058201 * nd.cthru[0] + 0.6701058201058201 * nd.cthru[1] + 0.0647883597883598 * nd.cthru[2];     // pbright(), color.h CIE_rf/gf/bf
            } else
                hasthru = false;
        }
        if (R.c[ck].maxHemi - b <= RB_FTINY) b = 0;  // have specular to sample?
        else {
            double vjit[3];
            bsdf_jitter(vjit, vray, sr_vpsa[1], P.specjitter, bkey, 10u + 20u * (unsigned)xmit);
            b = sd_direct_hemi(B, vjit, xmit != 0) - b;
            b *= (b > 0);
        }
        if (b <= P.specthresh + RB_FTINY) {          // below sampling threshold?
            if (b > RB_FTINY) unsamp[xmit][0] = unsamp[xmit][1] = unsamp[xmit][2] = (float)b;     // XXX no color from BSDF
            continue;
        }
        // sample_sdcomp(), one sample
        double xrand = rnd01(bkey, 12u + 20u * (unsigned)xmit);
        if (P.specjitter < 1.) xrand = .5 + P.specjitter * (xrand - .5);
        double vsmp[3];
        bsdf_jitter(vsmp, vray, sr_vpsa[0], P.specjitter, bkey, 13u + 20u * (unsigned)xmit);
        const double vinc[3] = {vsmp[0], vsmp[1], vsmp[2]};
        const double cieY = comp_sample(B, ck, vsmp, xrand, rnd01(bkey, 15u + 20u * (unsigned)xmit), rnd01(bkey, 16u + 20u * (unsigned)xmit));
        if (cieY <= RB_FTINY) continue;              // zero component?
        if (hasthru0) {                              // check for view ray
            const double dx = vinc[0] + vsmp[0], dy = vinc[1] + vsmp[1];
            if (dx * dx + dy * dy <= sr_vpsa[0] * sr_vpsa[0]) continue;            // exclude view sample
        }
        double sdir[3];
        if (!sd_map_dir(sdir, fromloc, vsmp)) continue;
        float rc[3];
        bsdf_gray(rc, cieY);
        QRay q;
        if (!rayorigin(P, r, xmit ? RT_TSPECULAR : RT_RSPECULAR, rc, true, q)) continue;
        q.dir[0] = sdir[0]; q.dir[1] = sdir[1]; q.dir[2] = sdir[2];
        if (xmit && thick != 0) for (int k = 0; k < 3; k++) q.org[k] += r.ron[k] * -thick;      // need to offset origin?
        push_ray(A, q);
    }
    // compute indirect diffuse
    float sct[3];
    for (int k = 0; k < 3; k++) sct[k] = nd.mcolor[k] + unsamp[0][k];
    if (max3(sct) > (float)RB_FTINY) multambient(A, r, sct, nd.pnorm);            // ambient from reflection
    for (int k = 0; k < 3; k++) sct[k] = nd.scolor[k] + unsamp[1][k];
    if (max3(sct) > (float)RB_FTINY) {               // ambient from other side
        const double bnorm[3] = {-nd.pnorm[0], -nd.pnorm[1], -nd.pnorm[2]};
        if (thick != 0) {                            // proxy with offset?
            const double keep[3] = {r.rop[0], r.rop[1], r.rop[2]};
            for (int k = 0; k < 3; k++) r.rop[k] = keep[k] + r.ron[k] * thick;
            multambient(A, r, sct, bnorm);
            for (int k = 0; k < 3; k++) r.rop[k] = keep[k];
        } else
            multambient(A, r, sct, bnorm);
    }
    // add direct component
    if (!anyt && max3(nd.scolor) <= (float)RB_FTINY) { nd.dmode = 1; direct_or_park(A, r, nd); }           // reflection only
    else if (thick == 0) { nd.dmode = 0; direct_or_park(A, r, nd); }                                   // thin surface scattering
    else {
        nd.dmode = 1; direct_or_park(A, r, nd);      // reflection first
        if (A.dout) r.nchild += (unsigned)A.S.nsrcs; // (a parked job numbers its shadow rays from the count it was parked with)
        const double keep[3] = {r.rop[0], r.rop[1], r.rop[2]};
        for (int k = 0; k < 3; k++) r.rop[k] = keep[k] + r.ron[k] * -thick;        // offset for transmitted
        nd.dmode = 2; direct_or_park(A, r, nd);      // separate transmission
        for (int k = 0; k < 3; k++) r.rop[k] = keep[k];
    }
}

// glass.c:46-165
template <bool FAST = false>
__device__ __forceinline__ void m_glass(const WaveArgs& A, RayCtx& r, const float* a, int nargs) {
    const DParams& P = A.P;
    double rindex = (nargs == 4) ? (double)a[3] : 1.52;
    if (!P.backvis && r.rod <= 0.0) { raytrans(A, r); return; }
    float mcolor[3] = {a[0], a[1], a[2]};
    bool hastrans = max3(mcolor) > 1e-15f;
    if (hastrans) { for (int k = 0; k < 3; k++) if (mcolor[k] < 1e-15f) mcolor[k] = 1e-15f; }
    else if (r.crtype & RT_SHADOW) return;
    bool flipped = false;
    if (r.rod < 0.0) { r.rod = -r.rod; r.ron[0] = -r.ron[0]; r.ron[1] = -r.ron[1]; r.ron[2] = -r.ron[2]; flipped = true; }
    double pdot = r.rod;
    double pnorm[3] = {r.ron[0], r.ron[1], r.ron[2]}, pert[3] = {0., 0., 0.};
    const bool hastexture = FAST ? false : ray_pert(A.S, r, flipped, pert);       // glass.c:91-98
    if (hastexture) pdot = raynormal(pnorm, r, pert);
    double cos2 = sqrt((1.0 - 1.0 / (rindex * rindex)) + pdot * pdot / (rindex * rindex));
    if (hastrans)
        for (int k = 0; k < 3; k++) mcolor[k] = (float)pow((double)mcolor[k], 1.0 / cos2);
    double r1e = (pdot - rindex * cos2) / (pdot + rindex * cos2);
    r1e *= r1e;
    double r1m = (1.0 / pdot - rindex / cos2) / (1.0 / pdot + rindex / cos2);
    r1m *= r1m;
    if (hastrans) {
        float rc[3];
        for (int k = 0; k < 3; k++) {
            double d = mcolor[k];
            rc[k] = (float)(.5 * (1.0 - r1e) * (1.0 - r1e) * d / (1.0 - r1e * r1e * d * d) +
                            .5 * (1.0 - r1m) * (1.0 - r1m) * d / (1.0 - r1m * r1m * d * d));
        }
        QRay q;
        if (rayorigin(P, r, RT_TRANS, rc, true, q)) {
            q.dir[0] = r.dir[0]; q.dir[1] = r.dir[1]; q.dir[2] = r.dir[2];
            if (hastexture && !(r.crtype & (RT_SHADOW | RT_AMBIENT)) && !(r.xfl & 2)) {      // glass.c:124-130
                double pd[3];
                for (int k = 0; k < 3; k++) pd[k] = r.dir[k] + pert[k] * (2. * (1. - rindex));
                if (normalize3(pd) != 0.0) { q.dir[0] = pd[0]; q.dir[1] = pd[1]; q.dir[2] = pd[2]; }
            }
            push_ray(A, q);
        }
    }
    if (r.crtype & RT_SHADOW) return;
    float rc[3];
    for (int k = 0; k < 3; k++) {
        double d = mcolor[k];
        d *= d;
        rc[k] = (float)(.5 * r1e * (1.0 + (1.0 - 2.0 * r1e) * d) / (1.0 - r1e * r1e * d) +
                        .5 * r1m * (1.0 + (1.0 - 2.0 * r1m) * d) / (1.0 - r1m * r1m * d));
    }
    QRay q;
    if (rayorigin(P, r, RT_REFLECTED, rc, true, q)) {
        for (int k = 0; k < 3; k++) q.dir[k] = r.dir[k] + pnorm[k] * (2. * pdot);
        normalize3(q.dir);
        push_ray(A, q);
    }
}

// gen/skybright.cal:21-44 (gensky) and gen/perezlum.cal:18-39 (gendaylight) as native code;
// rayinit.cal: Acos(x) = acos(bound(-1,x,1)), if(a,b,c) = a > 0 ? b : c, select(N,..) = argument int(N+.5).
// The function sees the ray direction in its own coordinates (func.c:455-458).
__device__ __noinline__ double sky_pattern(const PatRec& p, const double dir[3]) {
    const double Dx = dir[0] * p.xb[0] + dir[1] * p.xb[3] + dir[2] * p.xb[6];
    const double Dy = dir[0] * p.xb[1] + dir[1] * p.xb[4] + dir[2] * p.xb[7];
    const double Dz = dir[0] * p.xb[2] + dir[1] * p.xb[5] + dir[2] * p.xb[8];
    auto Acos = [](double x) { return acos(x < -1 ? -1 : x > 1 ? 1 : x); };
    const double* const pa = p.a;
#define A(i) pa[(i) - 1]                         /* the function file's A1..A10 */
    double sky, ground;
    if (p.kind == PAT_SKYBRIGHT) {
        const double cosgamma = Dx * A(5) + Dy * A(6) + Dz * A(7);
        const double gamma = Acos(cosgamma), zt = Acos(A(7)), eta = Acos(Dz);
        const int sel = (int)(A(1) + .5);
        if (sel == 1)
            sky = A(2) * (.91 + 10 * exp(-3 * gamma) + .45 * cosgamma * cosgamma) * (Dz - .01 > 0 ? 1.0 - exp(-.32 / Dz) : 1.0) / A(4);
        else if (sel == 2) sky = A(2) * (1 + 2 * Dz) / 3;
        else if (sel == 3) sky = A(2);
        else
            sky = A(2) * ((1.35 * sin(5.631 - 3.59 * eta) + 3.12) * sin(4.396 - 2.6 * zt) + 6.37 - eta) / 2.326 *
                  exp(gamma * -.563 * ((2.629 - eta) * (1.562 - zt) + .812)) / A(4);
        ground = A(3);
    } else {
        const double cosgamma = Dx * A(8) + Dy * A(9) + Dz * A(10);
        const double gamma = Acos(cosgamma);
        const double dz = (Dz - 0.01 > 0) ? Dz : 0.01;
        sky = A(1) * (1 + A(3) * exp(A(4) / dz)) * (1 + A(5) * exp(A(6) * gamma) + A(7) * cosgamma * cosgamma);
        ground = A(2);
    }
    const double a = pow(Dz + 1.01, 10.0), b = pow(Dz + 1.01, -10.0);
    return (a * sky + b * ground) / (a + b);
#undef A
}

// source.c:749-793 m_light.  Returns 1 and sets rcol when the ray sees the
// emitter, 0 when its coefficient is zeroed / it is passed on.
template <bool FAST = false, bool LEAN = false>
__device__ __forceinline__ int m_light(const WaveArgs& A, RayCtx& r, const MatRec& m, float rcol[3], bool& zeroed) {
    const DScene& S = A.S;
    zeroed = false;
    bool isglow = (m.kind == MK_GLOW);
    // distglow(m, r, d): glow too far away to act as a source
    auto distglow = [&](double d) { return isglow && m.a[3] >= -(float)RB_FTINY && d > (double)m.a[3]; };
    // badcomponent
    if ((r.crtype & (RT_AMBIENT | RT_SPECULAR)) &&
        !((r.crtype & RT_SHADOW) || r.rod < 0.0 || distglow(r.rot))) { zeroed = true; return 0; }
    // wrongsource
    if (r.rsrc >= 0 && S.srcs[r.rsrc].so != r.robj) {
        bool illumblock = false;
        if (m.kind == MK_ILLUM) {
            const MatRec& sm = S.mats[S.srcs[r.rsrc].mat];
            illumblock = r.rod > 0.0 && (sm.kind == MK_ILLUM || sm.kind == MK_GLOW);
        }
        if (m.kind != MK_ILLUM || illumblock) { zeroed = true; return 0; }
    }
    // passillum
    if (m.kind == MK_ILLUM && (r.rsrc < 0 || S.srcs[r.rsrc].so != r.robj)) return 2;
    // srcignore (-dv-): path length approximated by this ray's own length
    if (!(A.P.directvis || (r.crtype & RT_SHADOW) || distglow(r.rot))) { zeroed = true; return 0; }
    if (r.rod < 0.0) {
        if (!A.P.backvis) raytrans(A, r);
        return 0;
    }
    if (!LEAN && m.kind == MK_SPOT) {             // check for outside spot (source.c:778-779)
        if (r.rsrc >= 0 && (S.srcs[r.rsrc].flags & SF_SPOT)) {
            if (spotout(S.srcs[r.rsrc], r.org, r.dir)) return 0;
        } else {                                  // seen directly: makespot() from the material's reals
            SrcRec sp;
            sp.spot_siz = (float)(2.0 * RB_PI * (1.0 - cos(RB_PI / 180.0 / 2.0 * (double)m.a[3])));
            sp.spot_aim[0] = m.a[4]; sp.spot_aim[1] = m.a[5]; sp.spot_aim[2] = m.a[6];
            sp.spot_flen = (float)normalize3(sp.spot_aim);
            if (spotout(sp, r.org, r.dir)) return 0;
        }
    }
    if (!FAST && (m.flags & 1) && A.P.need_values) {       // pattern under an emitter only matters for values
        atomicOr(&A.C->errflag, RB_ERR_UNSUP_MOD); A.C->errobj = (unsigned)m.obj;
    }
    rcol[0] = m.a[0]; rcol[1] = m.a[1]; rcol[2] = m.a[2];
    if (!FAST && m.pat >= 0) {                    // raytexture(r, m->omod): sky brightness functions
        float pcol = 1.f;
        for (int pi = m.pat; pi >= 0; pi = S.pats[pi].next) pcol *= (float)sky_pattern(S.pats[pi], r.dir);
        rcol[0] *= pcol; rcol[1] *= pcol; rcol[2] *= pcol;
    }
    return 1;
}

// rayshade() + trace callback for one traced ray (raytrace.c:162-179,210-256)
// FAST: the materials of k_shade_mid / k_shade_fast only (shade_class() decides which rays may come here); LEAN
// (k_shade_fast): without glass, trans and spotlights as well.
template <bool FAST = false, bool LEAN = false>
__device__ __forceinline__ void shade_ray(const WaveArgs& A, RayCtx& r) {
    const DScene& S = A.S;
    int4 hd = __ldg(&S.objhdr[r.robj]);
    int ms = hd.z;
    float rcol[3] = {0.f, 0.f, 0.f};
    bool zeroed = false, have_rcol = false;
    if (ms < 0) {                       // no material: rayshade() returns 0 -> raytrans
        raytrans(A, r);
        trace_contrib(A, r, false, rcol, false);
        return;
    }
    const MatRec* m = &S.mats[ms];
    bool tst_irrad = A.P.do_irrad && !(r.crtype & ~(RT_PRIMARY | RT_TRANS));
    int nk = -1;                        // m_normal() material kind and reals, when that is what shades the ray
    float na[7];
    for (int guard = 0; guard < 8; guard++) {
        int k = m->kind;
        if (k == MK_UNSUPPORTED || ((m->flags & 1) && !(k >= MK_LIGHT && k <= MK_SPOT)) || (m->flags & 2)) {
            atomicOr(&A.C->errflag, (m->flags & 2) ? RB_ERR_LOCAL_SRC : (k == MK_UNSUPPORTED ? RB_ERR_UNSUP_MAT : RB_ERR_UNSUP_MOD));
            A.C->errobj = (unsigned)m->obj;
            return;
        }
        if (tst_irrad) {                // raytirrad(), raytrace.c:210-228
            if (k == MK_TRANS || k == MK_GLASS || k == MK_TRANS2 || k == MK_DIELECTRIC || k == MK_INTERFACE || k == MK_ABSDF ||
                (k == MK_BSDF && (m->flags & 8))) { raytrans(A, r); break; }      // istransp(m) || isBSDFproxy(m)
            if (!(k >= MK_LIGHT && k <= MK_SPOT)) { nk = MK_PLASTIC; for (int j = 0; j < 7; j++) na[j] = j < 3 ? (float)RB_PI : 0.f; break; }
        }
        if (k == MK_PLASTIC || k == MK_METAL || k == MK_TRANS) { nk = k; for (int j = 0; j < 7; j++) na[j] = m->a[j]; break; }
        if (!LEAN && k == MK_GLASS) { m_glass<FAST>(A, r, m->a, m->nargs); break; }
        if (!FAST) {
            if (k >= MK_PLASTIC2 && k <= MK_TRANS2) { m_aniso(A, r, k, m->a, m->u); break; }
            if (k == MK_DIELECTRIC || k == MK_INTERFACE) { m_dielectric(A, r, k, (int)(m - S.mats), m->a); break; }
            if (k == MK_BSDF || k == MK_ABSDF) { m_bsdf(A, r, *m); break; }
        }
        int rv = m_light<FAST, LEAN>(A, r, *m, rcol, zeroed);
        if (rv == 1) {
            have_rcol = true;
            add_value(A, r.row, r.coef, rcol[0], rcol[1], rcol[2]);
        } else if (rv == 0 && !(r.rod < 0.0 && !A.P.backvis)) {
            // the emitter answered black (seen from behind, outside its spot cone, wrong source ...):
            // the reference multiplies by rcol = 0 and adds nothing (rcontrib.c trace_contrib)
            have_rcol = true;
        } else if (rv == 2) {           // passed illum: alternate material or straight through
            if (m->alt >= 0) { m = &S.mats[m->alt]; continue; }
            raytrans(A, r);
        }
        break;
    }
    if (nk >= 0) m_normal<FAST, LEAN>(A, r, nk, na);
    trace_contrib(A, r, zeroed, rcol, have_rcol);
}

// Which kernel shades a queued ray (mirrors the decisions of shade_ray() / m_normal() above; anything unusual goes
// to the general kernel):
//   SC_NONE  nothing to do: no hit, or a ray that provably adds nothing and spawns nothing (below)
//   SC_DIFF  k_shade_fast itself: a surface of plastic / metal without specular reflection, in a scene whose sources
//            are all glow: all the material does is multambient() (shade_diffuse())
//   SC_SPEC  k_shade_spec: the same surfaces with specular reflection: m_normal<FAST, LEAN, NODIRECT> alone
//   SC_LEAN  k_shade_lean: plastic / metal without a sampled highlight, plain light / glow emitters, surfaces
//            without a material, the Lambertian stand-in of an irradiance ray (raytirrad) -- shade_ray<FAST, LEAN>
//   SC_MID   k_shade_mid: glass, trans without a sampled highlight, spotlights -- shade_ray<FAST>
//   SC_SLOW  k_shade: everything else
enum : int { SC_NONE = 0, SC_DIFF, SC_SPEC, SC_LEAN, SC_MID, SC_SLOW };
// (geomoff / mat: the hit object's record and material, for shade_diffuse())
__device__ __forceinline__ int shade_class(const WaveArgs& A, unsigned qinfo, unsigned qmed, const HitRec& hr, int& geomoff,
                                           const MatRec*& mat) {
    if (qmed) return SC_SLOW;                                 // absorbing medium: ray_medium()
    const int crtype = qinfo & 0x3ff;
    if (A.res && crtype == RT_PRIMARY) return SC_SLOW;        // primary-hit report (smooth_pert, flip flag)
    if (hr.robj < 0) return SC_NONE;                          // nothing to shade
    const int4 hd = __ldg(&A.S.objhdr[hr.robj]);
    if (hr.local && ((hd.x >> 13) & 3)) return SC_SLOW;       // vertex normals / Phong modifier
    if (hd.z < 0) return SC_LEAN;                             // no material: raytrans()
    const MatRec& m = A.S.mats[hd.z];
    geomoff = hd.w; mat = &m;
    const int k = m.kind;
    if (k == MK_UNSUPPORTED || (m.flags & 3)) return SC_SLOW; // error paths and patterns
    if (k == MK_BSDF || k == MK_ABSDF) return SC_SLOW;        // m_bsdf() is compiled into the general kernel only
    const bool emitter = k >= MK_LIGHT && k <= MK_SPOT;
    const bool tirrad = A.P.do_irrad && !(crtype & ~(RT_PRIMARY | RT_TRANS));
    if (tirrad && !emitter) return SC_LEAN;                   // raytirrad(): passes through or Lambertian
    if (k == MK_PLASTIC || k == MK_METAL) {
        double a2 = m.a[4]; a2 *= a2;
        if (!(a2 <= RB_FTINY)) return SC_SLOW;
        // A ray that ends on such a surface, adds nothing and spawns nothing -- a shadow ray (normal.c:190-191), or
        // a ray past the last ambient bounce on a surface without specular reflection in a scene whose sources are
        // all glow -- needs neither its hit frame nor m_normal(): multambient() is "dumb" with a black -av,
        // direct() has no source to test, and trace_contrib() returns for an untracked modifier.  (Back faces with
        // -bv- go through raytrans(), hence the test.)
        const bool untracked = !A.acc || __ldg(&A.S.otrack[hr.robj]) < 0;
        if (untracked && A.P.backvis) {
            if (crtype & RT_SHADOW) return SC_NONE;
            const int rdepth = (qinfo >> 16) & 0x3f;
            const bool dumb = (A.P.ambdiv <= 0) | (rdepth >= A.P.ambounce);
            const bool black_av = !A.vacc || !(A.P.ambval[0] > 0.f || A.P.ambval[1] > 0.f || A.P.ambval[2] > 0.f);
            if (dumb && black_av && A.nodirect && m.a[3] == 0.f) return SC_NONE;
            // no specular reflection, no Fresnel term (normal.c:229-235 needs rspec >= .018), nothing for direct() to do:
            // what is left of m_normal() is multambient() with the material's colour
#ifndef RB_NO_DIFF
            if (A.nodirect && m.a[3] == 0.f && hr.local) return SC_DIFF;
            // the same with a specular component (mirror ray, Fresnel term): m_normal() without direct() and
            // trace_contrib(), in a kernel of its own (k_shade_spec)
            if (A.nodirect && hr.local && !(crtype & RT_SHADOW)) return SC_SPEC;
#endif
        }
        return SC_LEAN;
    }
    if (k == MK_TRANS) { double a2 = m.a[4]; a2 *= a2; return a2 <= RB_FTINY ? SC_MID : SC_SLOW; }
    if (k == MK_GLASS) return SC_MID;
    if (emitter) return (k == MK_ILLUM || m.pat >= 0) ? SC_SLOW : k == MK_SPOT ? SC_MID : SC_LEAN;
    return SC_SLOW;
}

// SC_DIFF: m_normal() of a surface whose material is plastic / metal with rspec = 0 and roughness = 0, hit by a ray that
// is not a shadow ray, in a scene without sources for direct(), modifier untracked, back faces visible: flipsurface(),
// rdiff = 1, and multambient(mcolor * rdiff, ron) is all that happens (normal.c:192-200,215-248,337-345).
__device__ __forceinline__ void shade_diffuse(const WaveArgs& A, const QRay& q, const HitRec& hr, int geomoff, const MatRec& m) {
    RayCtx r;
    for (int k = 0; k < 3; k++) r.coef[k] = q.coef[k];
    r.rmax = q.rmax; r.rweight = q.rweight; r.row = q.row;
    r.crtype = q.info & 0x3ff; r.rlvl = (q.info >> 10) & 0x3f; r.rdepth = (q.info >> 16) & 0x3f;
    r.rsrc = q.rsrc;
    r.key = ((unsigned long long)q.key_hi << 32) | q.key_lo;
    r.nchild = 0; r.med = 0; r.re = 0.f;
    r.robj = hr.robj; r.rot = hr.rot;
    double rod;
    hit_frame(A.S, hr.robj, hr.rot, q.org, q.dir, r.rop, r.ron, rod);
    if (rod < 0.0) { r.ron[0] = -r.ron[0]; r.ron[1] = -r.ron[1]; r.ron[2] = -r.ron[2]; }       // flipsurface
    const float sct[3] = {m.a[0], m.a[1], m.a[2]};
    multambient(A, r, sct, r.ron);
}

}  // namespace rb
