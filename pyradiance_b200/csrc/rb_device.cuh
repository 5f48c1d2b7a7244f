// rb_device.cuh -- device-side tables and queue records (sm_100a).
//
// HBM layout (all arrays resident for the life of the loaded octree):
//   nodes[8*nnodes]  int32   child words; one node = 32 B = one sector
//   leafpool[]       int2    (count,0),(id, geom offset)... per full leaf, ids ascending
//   objhdr[nobjs]    int4    kind|flags|nv, omod, material slot, geom offset
//   geom[]           double  16-byte aligned primitive records (plane + 2-D
//                            vertices for faces, centre+radius, cone frame)
//   mats[], srcs[]           material and light-source tables
//   otrack[nobjs]    int32   tracked-modifier slot of each object's modifier
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "rb_scene.hpp"

namespace rb {

#define RB_FTINY 1e-6
#define RB_FHUGE 1e10
#define RB_PI 3.14159265358979323846
#define RB_MAXDEPTH 21          // octree levels held in a 64-bit coordinate walk
#define RB_STACK (RB_MAXDEPTH + 1)

// ray type flags (rt/ray.h:27-43)
enum : int {
    RT_PRIMARY = 01, RT_RSHADOW = 02, RT_REFLECTED = 04, RT_REFRACTED = 010,
    RT_TRANS = 020, RT_RAMBIENT = 040, RT_RSPECULAR = 0100, RT_TSHADOW = 0200,
    RT_TAMBIENT = 0400, RT_TSPECULAR = 01000,
    RT_SHADOW = RT_RSHADOW | RT_TSHADOW, RT_AMBIENT = RT_RAMBIENT | RT_TAMBIENT,
    RT_SPECULAR = RT_RSPECULAR | RT_TSPECULAR,
    RT_RAYREFL = RT_RSHADOW | RT_REFLECTED | RT_RAMBIENT | RT_RSPECULAR
};

struct DScene {
    double cuorg[3];
    double cusize;
    int root, nobjs, nsrcs, maxdepth;
    const int* __restrict__ nodes;
    const int* __restrict__ leafpool;
    const int4* __restrict__ objhdr;
    const double* __restrict__ geom;
    const MatRec* __restrict__ mats;
    const SrcRec* __restrict__ srcs;
    const PatRec* __restrict__ pats;
    const int* __restrict__ otrack;
    const BsdfRec* __restrict__ bsdfs;          // BSDF / aBSDF materials: Klems-matrix data (rb_bsdf.cuh)
    const BsdfBasis* __restrict__ bsdfbases;
    const unsigned* __restrict__ bsdfpool;
    // integer walk (rb_geom.cuh): positions are maxdepth-bit integers, inv_cell = 2^maxdepth / cusize; top[] holds,
    // for every cell of level topk (x | y << topk | z << 2 topk), the node word and the level of the cube that
    // contains the cell (an interior node at level topk, or the leaf / empty cube above it)
    const int2* __restrict__ top;
    int topk;
    double inv_cell;
};

// render options (rt/ray.h:157-190 RAYPARAMS subset used by this path)
struct DParams {
    int ambounce, ambdiv, maxdepth, backvis, directvis, do_irrad;
    int contrib;            // rcontrib -V+
    int need_values;        // radiance values are consumed (rtrace -ov, -V+)
    float minweight;
    float ambval[3];
    double dstrsrc, specthresh, specjitter, srcsizerat;
    unsigned long long seed;
};

// one tracked modifier (rt/rcontrib.h:62-74 MODCONT, bins as native code)
enum : int { BIN_CONST = 0, BIN_REINHARTB, BIN_REINHART, BIN_KLEMS_FULL, BIN_HEMI,
             BIN_KLEMS_HALF, BIN_KLEMS_QUARTER, BIN_SHIRCHIU };
struct DBinSpec {
    int fn, nbins, col0, mf;
    double n[3], u[3], rhs;
    int cbin;               // BIN_CONST value
    int pad;
};

// a ray waiting to be traced: 96 bytes
struct __align__(16) QRay {
    double org[3];
    double dir[3];
    double rmax;
    float coef[3];          // product of rcoef from the primary down to this ray
    float rweight;
    unsigned row;           // output record
    unsigned info;          // crtype (10 bits) | rlvl<<10 (6) | rdepth<<16 (6) | spare
    int rsrc;
    unsigned key_lo, key_hi;  // RNG path key
    unsigned med;           // medium the ray travels in: 0 = none, else (material slot + 1) << 1 | side (rb_shade.cuh medium_cext)
};
static_assert(sizeof(QRay) == 96, "QRay must be 96 bytes");

// a hemisphere waiting to be expanded into n*n rays (rt/ambcomp.c:350-422)
struct __align__(16) QHemi {
    double rop[3];
    double onrm[3];
    double rmax_rem;        // (parent rmax > FTINY) * (rmax - rot), for TAMBIENT children
    float acoef[3];         // per-division coefficient (rcol / n^2)
    float ccoef[3];         // parent's cumulative coefficient
    float rweight;          // parent's weight
    int n;
    unsigned row, info;     // parent's crtype / rlvl / rdepth
    unsigned key_lo, key_hi;
    int atype;              // RT_RAMBIENT or RT_TAMBIENT | medium of the parent ray << 10
    int rsrc;
};
static_assert(sizeof(QHemi) == 112, "QHemi must be 112 bytes");

// what k_trace hands to k_shade for each queued ray
struct HitRec {
    double rot, rod;
    int robj;               // object hit, source object for distant-source hits, -1 none
    int local;              // 1: local surface hit, 0: distant source / nothing
};

// per-ray result for rtrace-style queries
struct RayResult {
    double rop[3];
    double ron[3];
    double rot;
    double rod;
    int robj;               // surface object index or -1
    int omod;               // its modifier object index or -1
    float rweight;
    int pad;
    double pert[3];         // RAY.pert of the primary hit (o_mesh.c:201-209)
};

struct DCounters {
    unsigned long long nrays;      // calls to localhit
    unsigned long long nodes;      // octree node words read
    unsigned long long leafents;   // leaf set entries read
    unsigned long long prims;      // primitive tests
    unsigned long long contribs;   // accumulated contributions
    unsigned nq_out;               // rays pushed to the next queue
    unsigned nh_out;               // hemispheres pushed
    unsigned hemi_rays;            // ray slots reserved by pushed hemispheres
    unsigned overflow;             // queue overflow
    unsigned errflag;              // RB_ERR_* bits
    unsigned errobj;               // offending object
    unsigned badbin;               // bin >= nbins warnings
    unsigned next_ray;             // k_trace's persistent-thread fetch counter
    unsigned nd_out;               // parked direct() jobs (keep right after next_ray: reset together)
    unsigned nslow;                // rays k_shade_fast left to the general k_shade (reset with the two above)
    unsigned nmid;                 // rays it left to k_shade_mid
    unsigned nlean;                // rays it left to k_shade_lean
    unsigned nspec;                // rays it left to k_shade_spec
    // wave chaining without the host (rb_engine.cu k_gate / k_prepare): what the kernels of the current wave read
    unsigned nin;                  // rays in the input queue of this wave
    unsigned nh_in, nd_in;         // hemispheres / parked direct() jobs to expand before it
    unsigned long long rays_traced;   // sum of nin over the waves of the batch
    unsigned wave_nin[64];         // nin of the last waves (slot = wave index & 63), for the host's statistics
};
enum : unsigned { RB_ERR_UNSUP_MAT = 1, RB_ERR_UNSUP_PRIM = 2, RB_ERR_UNSUP_MOD = 4,
                  RB_ERR_LOCAL_SRC = 8, RB_ERR_DEPTH = 16, RB_ERR_CONTRIB_VALUE = 32 };

}  // namespace rb
