// rb_engine.cu -- kernels and host driver of the wavefront ray engine (sm_100a).
//
// Kernels (all hand-written, launched on one stream):
//   k_init    one thread per input ray: primary ray into the queue, or, for -I
//             sensors, the pretend Lambertian hit of rt/rtrace.c:415-432 /
//             rt/rcontrib.c:321-339 / rt/RtraceSimulManager.cpp:315-335, which
//             emits one hemisphere record and the sensor's shadow rays
//   k_expand  one CTA per hemisphere record: the n*n stratified Shirley-Chiu
//             samples of rt/ambcomp.c:350-422 become queue rays
//   k_wave    one thread per queued ray: octree walk + intersection
//             (rb_geom.cuh), then shading, contribution accumulation and
//             child-ray emission (rb_shade.cuh)
//   k_finish  accumulators (double) -> output rows (float32), rc2.c:293-335
// The host loops wave after wave until the queues are empty; batches of
// records are sized so that a wave fits the two ray queues in HBM.
#include "rb_engine.cuh"
#include "rb_shade.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace rb {

#ifndef WAVE_THREADS
#define WAVE_THREADS 128
#endif
#ifndef RB_MINBLOCKS
#define RB_MINBLOCKS 5
#endif

static const size_t kDefaultQueue = (size_t)48 << 20;      // rays per queue of a large job

#define CK(call)                                                                     \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call; \
            return false;                                                            \
        }                                                                            \
    } while (0)

// ------------------------------------------------------------- kernels -----
#if RB_WALK_STATS
__device__ __forceinline__ void flush_stats(DCounters* C, const WalkStats& ws) {
    unsigned a = __reduce_add_sync(__activemask(), ws.nodes);
    unsigned b = __reduce_add_sync(__activemask(), ws.leafents);
    unsigned c = __reduce_add_sync(__activemask(), ws.prims);
    unsigned lane = threadIdx.x & 31;
    unsigned leader = __ffs(__activemask()) - 1;
    if (lane == leader) {
        atomicAdd(&C->nodes, (unsigned long long)a);
        atomicAdd(&C->leafents, (unsigned long long)b);
        atomicAdd(&C->prims, (unsigned long long)c);
    }
}
#endif

// Trace: octree walk + intersection only (small code).  Persistent threads:
// the grid is sized to fill the machine once and every warp pulls rays from
// the queue until it is empty (rb_geom.cuh walk_rays).
#ifndef RB_SHADOW_ANYHIT
#define RB_SHADOW_ANYHIT 1
#endif
#ifdef RB_TRACE_MAXNREG        // developer knob: cap the registers directly (CTA sizes whose warps do not divide evenly)
__global__ void __maxnreg__(RB_TRACE_MAXNREG) k_trace(const WaveArgs A) {
#else
__global__ void __launch_bounds__(WAVE_THREADS, RB_MINBLOCKS) k_trace(const WaveArgs A) {
#endif
    __shared__ WalkSmem<WAVE_THREADS> sm;
    extern __shared__ int stk_dyn[];         // [maxdepth + 1][WAVE_THREADS]
    TraceIO io;
    io.qin = A.qin; io.nin = A.C->nin; io.hits = A.hits; io.next = &A.C->next_ray;
    io.anyhit = A.anyhit;
    WalkStats ws = {0, 0, 0};
    walk_rays<WAVE_THREADS>(A.S, io, sm, stk_dyn, ws, &A.C->errflag, &A.C->errobj);
#if RB_WALK_STATS
    flush_stats(A.C, ws);
#endif
}
#if RB_WALK_STATS
__global__ void k_dbg_print() {
    printf("[rb dbg] cyl pairs %llu, after the axis test %llu, after the end test %llu, candidates %llu; sphere pairs %llu, with roots %llu\n",
           g_dbg[0], g_dbg[1], g_dbg[5], g_dbg[2], g_dbg[3], g_dbg[4]);
    printf("[rb dbg] steps %llu, of which out of full leaves %llu; mean level of the leaf stepped out of %.2f\n", g_dbg[8], g_dbg[9],
           (double)g_dbg[10] / (double)(g_dbg[8] ? g_dbg[8] : 1));
}
#endif

// Shade: material evaluation, contribution accumulation, child-ray emission.
#ifndef RB_SHADE_MINBLOCKS
#define RB_SHADE_MINBLOCKS 6
#endif
#ifndef RB_SHADE_THREADS
#define RB_SHADE_THREADS 128
#endif
// One queued ray -> RayCtx (hit frame included).
__device__ __forceinline__ void load_ray(const WaveArgs& A, const QRay& q, const HitRec& hr, RayCtx& r) {
    for (int k = 0; k < 3; k++) { r.org[k] = q.org[k]; r.dir[k] = q.dir[k]; r.coef[k] = q.coef[k]; }
    r.rmax = q.rmax; r.rweight = q.rweight; r.row = q.row;
    r.crtype = q.info & 0x3ff; r.rlvl = (q.info >> 10) & 0x3f; r.rdepth = (q.info >> 16) & 0x3f;
    r.rsrc = q.rsrc;
    r.key = ((unsigned long long)q.key_hi << 32) | q.key_lo;
    r.nchild = 0;
    r.med = q.med; r.re = 0.f;
    r.robj = hr.robj; r.rot = hr.rot; r.rod = hr.rod; r.flat = false; r.xfl = 0;
    if (hr.local) {
        hit_frame(A.S, hr.robj, hr.rot, r.org, r.dir, r.rop, r.ron, r.rod);
        const int hx = __ldg(&A.S.objhdr[hr.robj]).x;
        const int kind = hx & 0xff;
        r.flat = ((kind == PK_FACE) | (kind == PK_RING)) & !(hx & PX_NOTFLAT);
        r.xfl = (unsigned char)((hx >> 13) & 3);          // PX_SMOOTH, PX_PHONG
    } else {
        for (int k = 0; k < 3; k++) { r.rop[k] = r.org[k]; r.ron[k] = -r.dir[k]; }
    }
}

// Shading is split by what a ray hit (shade_class(), rb_shade.cuh):
//   k_shade_fast  one thread per queued ray: reads the ray's type and its hit, classifies it, and shades the commonest
//                 class itself -- a diffuse polygon in a scene lit by glow sources only, where the material comes down
//                 to multambient() (shade_diffuse()); rays that end without effect (SC_NONE) stop right there; the
//                 others go, by queue slot, to
//   k_shade_lean  plastic / metal without a sampled highlight, plain emitters, surfaces without a material
//                 (shade_ray<true, true>: the rest of the material set compiled out);
//   k_shade_mid   glass, trans without a sampled highlight, spotlights (shade_ray<true, false>);
//   k_shade       every material.
#ifndef RB_FAST_MINBLOCKS
#define RB_FAST_MINBLOCKS 6
#endif
#ifndef RB_SHADE_EAGER
#define RB_SHADE_EAGER 0
#endif
#ifndef RB_DIFF_MINBLOCKS
#define RB_DIFF_MINBLOCKS 6
#endif
// (the bodies are functions of their own so that the grid-stride loop around them -- the ray count is known only on
//  the device -- does not add to the register pressure of the shading code: inlined, the loop tripled the spills)
__device__ __noinline__ void shade_fast_one(const WaveArgs& A, unsigned i) {
    // classification needs the last 32-byte sector of the queued ray only (type, depth, medium): a ray that ends here
    // never has the other two read
    const QRay* qp = A.qin + i;
#if RB_SHADE_EAGER
    const QRay q = *qp;                       // all three sectors at once: the chain below is latency, not bandwidth
    const unsigned qinfo = q.info, qmed = q.med;
#else
    const unsigned qinfo = __ldg(&qp->info), qmed = __ldg(&qp->med);
#endif
    const HitRec hr = A.hits[i];
    int geomoff = 0;
    const MatRec* mat = nullptr;
    const int cls = shade_class(A, qinfo, qmed, hr, geomoff, mat);
    if (cls == SC_NONE) return;
    // (one branch per class: reserve_slot() aggregates over the lanes that arrive together, which must share a counter)
    if (cls == SC_SPEC) { A.spec[reserve_slot(&A.C->nspec)] = i; return; }
    if (cls == SC_LEAN) { A.lean[reserve_slot(&A.C->nlean)] = i; return; }
    if (cls == SC_MID) { A.mid[reserve_slot(&A.C->nmid)] = i; return; }
    if (cls == SC_SLOW) { A.slow[reserve_slot(&A.C->nslow)] = i; return; }
#if !RB_SHADE_EAGER
    const QRay q = *qp;
#endif
    shade_diffuse(A, q, hr, geomoff, *mat);
}
__global__ void __launch_bounds__(RB_SHADE_THREADS, RB_DIFF_MINBLOCKS) k_shade_fast(const __grid_constant__ WaveArgs A) {
    const unsigned n = A.C->nin;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) shade_fast_one(A, i);
}
__device__ __noinline__ void shade_spec_one(const WaveArgs& A, unsigned i) {
    const QRay q = A.qin[i];
    const HitRec hr = A.hits[i];
    RayCtx r;
    load_ray(A, q, hr, r);
    const MatRec& m = A.S.mats[__ldg(&A.S.objhdr[hr.robj]).z];
    float na[7];
    for (int j = 0; j < 7; j++) na[j] = m.a[j];
    m_normal<true, true, true>(A, r, m.kind, na);
}
__global__ void __launch_bounds__(RB_SHADE_THREADS, RB_FAST_MINBLOCKS) k_shade_spec(const __grid_constant__ WaveArgs A) {
    const unsigned n = A.C->nspec;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) shade_spec_one(A, A.spec[j]);
}
__device__ __noinline__ void shade_lean_one(const WaveArgs& A, unsigned i) {
    const QRay q = A.qin[i];
    const HitRec hr = A.hits[i];
    RayCtx r;
    load_ray(A, q, hr, r);
    shade_ray<true, true>(A, r);
}
__global__ void __launch_bounds__(RB_SHADE_THREADS, RB_FAST_MINBLOCKS) k_shade_lean(const __grid_constant__ WaveArgs A) {
    const unsigned n = A.C->nlean;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) shade_lean_one(A, A.lean[j]);
}
__device__ __noinline__ void shade_mid_one(const WaveArgs& A, unsigned i) {
    const QRay q = A.qin[i];
    const HitRec hr = A.hits[i];
    RayCtx r;
    load_ray(A, q, hr, r);
    shade_ray<true, false>(A, r);
}
__global__ void __launch_bounds__(RB_SHADE_THREADS, RB_FAST_MINBLOCKS) k_shade_mid(const __grid_constant__ WaveArgs A) {
    const unsigned n = A.C->nmid;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) shade_mid_one(A, A.mid[j]);
}

// The general shading kernel: every material.  With A.slow it takes the queue slots k_shade_fast left over
// (grid-stride over a count that only the device knows), else the whole queue.
__global__ void __launch_bounds__(RB_SHADE_THREADS, RB_SHADE_MINBLOCKS) k_shade(const __grid_constant__ WaveArgs A) {
    const unsigned n = A.slow ? A.C->nslow : A.C->nin;
    for (unsigned j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const unsigned i = A.slow ? A.slow[j] : j;
        const QRay q = A.qin[i];
        const HitRec hr = A.hits[i];
        RayCtx r;
        load_ray(A, q, hr, r);
        if (A.res && r.crtype == RT_PRIMARY) {
            RayResult& o = A.res[r.row - A.row0];
            for (int k = 0; k < 3; k++) { o.rop[k] = r.rop[k]; o.ron[k] = r.ron[k]; }
            o.rot = r.rot; o.rod = r.rod; o.robj = r.robj;
            o.omod = r.robj >= 0 ? __ldg(&A.S.objhdr[r.robj]).y : -1;
            o.rweight = r.rweight; o.pad = 0;
            o.pert[0] = o.pert[1] = o.pert[2] = 0.0;
            if (r.xfl & 1) smooth_pert(A.S, r.robj, r.rop, r.ron, false, o.pert);
        }
        if (r.robj >= 0) {
            const bool front = r.rod > 0.0;
            if (r.med) ray_medium(A, r, -1);          // path extinction of an absorbing medium (rayparticipate)
            shade_ray<false>(A, r);
            // the material reversed a surface hit from behind (flipsurface): rtrace -on reports it that way
            if (A.res && r.crtype == RT_PRIMARY && front != (r.rod > 0.0)) A.res[r.row - A.row0].pad = 1;
        }
    }
}

struct InitArgs {
    const double* rays;     // [n][6] for this batch
    unsigned nrays;
    unsigned long long ray0;     // global index of the batch's first ray
    int accum;
    int irrad;
    int lim_dist;
};

__device__ void init_ray(const WaveArgs& A, const InitArgs& I, unsigned i) {
    const double* v = I.rays + (size_t)i * 6;
    double org[3] = {v[0], v[1], v[2]}, dir[3] = {v[3], v[4], v[5]};
    unsigned row = I.accum > 0 ? (unsigned)((I.ray0 + i) / I.accum - I.ray0 / I.accum) + A.row0 : A.row0;
    double d = normalize3(dir);
    if (d == 0.0) {                                  // dummy ray: blank record
        if (A.res) { RayResult& o = A.res[i]; memset(&o, 0, sizeof(o)); o.robj = -1; o.omod = -1; }
        return;
    }
    unsigned long long key = mix64(A.P.seed ^ mix64(I.ray0 + i));
    if (I.irrad == IRR_NONE) {
        unsigned slot = reserve_slot(&A.C->nq_out);
        if (slot >= A.qcap) { A.C->overflow = 1; return; }
        QRay q;
        for (int k = 0; k < 3; k++) { q.org[k] = org[k]; q.dir[k] = dir[k]; q.coef[k] = 1.f; }
        q.rmax = I.lim_dist ? d : 0.0;
        q.rweight = 1.f;
        q.row = A.res ? A.row0 + i : row;           // rtrace reports per ray
        q.info = pack_info(RT_PRIMARY, 0, 0);
        q.rsrc = -1;
        q.key_lo = (unsigned)key; q.key_hi = (unsigned)(key >> 32); q.med = 0;
        A.qout[slot] = q;
        return;
    }
    RayCtx r;
    r.coef[0] = r.coef[1] = r.coef[2] = 1.f;
    r.rweight = 1.f; r.row = A.res ? A.row0 + i : row;
    r.crtype = RT_PRIMARY; r.rlvl = 0; r.rdepth = 0; r.rsrc = -1;
    r.robj = -1; r.flat = false; r.xfl = 0; r.key = key; r.nchild = 0; r.rmax = 0.0;
    r.med = 0; r.re = 0.f;
    r.rod = 1.0;
    for (int k = 0; k < 3; k++) { r.ron[k] = dir[k]; r.dir[k] = -dir[k]; }
    if (I.irrad == IRR_RTRACE) {                     // rtrace.c:443-448,415-432
        r.rot = 1e-5;
        for (int k = 0; k < 3; k++) { r.org[k] = org[k] + 1.1e-4 * dir[k]; r.rop[k] = r.org[k] + r.dir[k] * r.rot; }
        for (int k = 0; k < 3; k++) r.ron[k] = -r.dir[k];
    } else if (I.irrad == IRR_RCONTRIB) {            // rcontrib.c:321-339
        r.rot = 1e-5;
        for (int k = 0; k < 3; k++) { r.org[k] = org[k] + 1.1e-4 * dir[k]; r.rop[k] = org[k] + 1e-4 * dir[k]; }
    } else {                                         // RtraceSimulManager.cpp:315-335
        r.rot = 1e-4;
        for (int k = 0; k < 3; k++) { r.rop[k] = org[k] + r.ron[k] * r.rot; r.org[k] = r.rop[k] + r.ron[k] * r.rot; }
    }
    if (A.res) {
        RayResult& o = A.res[i];
        for (int k = 0; k < 3; k++) { o.rop[k] = r.rop[k]; o.ron[k] = r.ron[k]; }
        o.rot = r.rot; o.rod = r.rod; o.robj = -1; o.omod = -1; o.rweight = 1.f; o.pad = 0;
        o.pert[0] = o.pert[1] = o.pert[2] = 0.0;
    }
    const float lamb[5] = {(float)RB_PI, (float)RB_PI, (float)RB_PI, 0.f, 0.f};
    m_normal(A, r, MK_PLASTIC, lamb);
}

__global__ void __launch_bounds__(WAVE_THREADS) k_init(const WaveArgs A, const InitArgs I) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < I.nrays) init_ray(A, I, i);
}

// direct() for scenes with many sources: one CTA per parked shading point, one
// source per thread and pass (source.c:398-556 with every source tested).
struct DirectArgs {
    const DirectJob* din;
};

__global__ void __launch_bounds__(256) k_direct(const __grid_constant__ WaveArgs A, const DirectArgs D) {
    __shared__ DirectJob sj;
    const int ns = A.S.nsrcs;
    const unsigned njobs = A.C->nd_in;
    for (unsigned job = blockIdx.x; job < njobs; job += gridDim.x) {
        __syncthreads();
        {   // cooperative copy of the job record
            const unsigned* src = reinterpret_cast<const unsigned*>(D.din + job);
            unsigned* dst = reinterpret_cast<unsigned*>(&sj);
            for (unsigned w = threadIdx.x; w < sizeof(DirectJob) / 4; w += blockDim.x) dst[w] = src[w];
        }
        __syncthreads();
        const unsigned base = sj.r.nchild;
        for (int sn = threadIdx.x; sn < ns; sn += blockDim.x) {
            if (A.S.srcs[sn].flags & SF_DISTANT) direct_one(A, sj.r, sj.nd, sn, base);
            else direct_local(A, sj.r, sj.nd, sn, base);
        }
    }
}

struct ExpandArgs {
    const QHemi* hin;
};

// The wave loop runs on the device's own counters: the host queues the kernels of several waves ahead and looks at
// the counters once per chunk (run_batch).  k_gate opens a wave: what the previous wave's shading produced becomes
// this wave's hemispheres / parked direct() jobs, unless the rays they will expand into would not fit the queue
// (then the batch stops with the overflow flag and the host retries it smaller).  k_prepare, after the expansion,
// turns the filled queue into the wave's input.
__global__ void k_gate(DCounters* C, unsigned qcap) {
    if (C->overflow || C->errflag || (size_t)C->nq_out + C->hemi_rays > qcap) {
        if (!C->errflag && !C->overflow) C->overflow = 1;
        C->nh_in = 0; C->nd_in = 0; C->nq_out = 0;
    } else {
        C->nh_in = C->nh_out; C->nd_in = C->nd_out;
    }
    C->nh_out = 0; C->nd_out = 0; C->hemi_rays = 0;
}
__global__ void k_prepare(DCounters* C, unsigned wave) {
    const unsigned n = (C->overflow || C->errflag) ? 0u : C->nq_out;
    C->nin = n; C->nq_out = 0; C->next_ray = 0; C->nslow = 0; C->nmid = 0; C->nlean = 0; C->nspec = 0;
    C->rays_traced += n;
    C->wave_nin[wave & 63] = n;
}

__global__ void __launch_bounds__(256) k_expand(const WaveArgs A, const ExpandArgs E) {
    const unsigned nh = A.C->nh_in;
    for (unsigned job = blockIdx.x; job < nh; job += gridDim.x) {
        const QHemi h = E.hin[job];
        unsigned long long hkey = ((unsigned long long)h.key_hi << 32) | h.key_lo;
        double onrm[3] = {h.onrm[0], h.onrm[1], h.onrm[2]}, ux[3], uy[3];
        if (!getperpendicular_rand(ux, onrm, hkey)) continue;
        uy[0] = onrm[1] * ux[2] - onrm[2] * ux[1];
        uy[1] = onrm[2] * ux[0] - onrm[0] * ux[2];
        uy[2] = onrm[0] * ux[1] - onrm[1] * ux[0];
        RayCtx par;
        for (int k = 0; k < 3; k++) { par.rop[k] = h.rop[k]; par.coef[k] = h.ccoef[k]; par.ron[k] = onrm[k]; }
        par.rot = 0.0; par.rmax = h.rmax_rem; par.rod = 1.0;
        par.rweight = h.rweight; par.row = h.row;
        par.crtype = h.info & 0x3ff; par.rlvl = (h.info >> 10) & 0x3f; par.rdepth = (h.info >> 16) & 0x3f;
        par.rsrc = h.rsrc; par.robj = -1; par.flat = false; par.key = hkey;
        par.med = (unsigned)h.atype >> 10; par.re = 0.f;      // (the weight estimate is already in h.rweight)
        const int atype = h.atype & 0x3ff;
        const float acoef[3] = {h.acoef[0], h.acoef[1], h.acoef[2]};
        const int n = h.n, nn = n * n;
        for (int idx = threadIdx.x; idx < nn; idx += blockDim.x) {
            par.nchild = (unsigned)idx;
            QRay q;
            if (ambsample(A.P, par, atype, acoef, onrm, ux, uy, n, idx / n, idx % n, q)) push_ray(A, q);
        }
    }
}

// rc2.c:293-335 put_contrib(): record value = accumulated sum / accumulate count
// A partial final record is averaged over the rays it really got
// (rcontrib.c:417-423 "partial accumulation in final record").
template <class T>
__global__ void k_finish(const double* __restrict__ acc, T* __restrict__ out, size_t n, double scale,
                         size_t tail_start, double tail_scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (T)(acc[i] * (i >= tail_start ? tail_scale : scale));
}

// ---------------------------------------------------------------- host -----
Engine::Engine(int device) : dev_(device) {
    cudaSetDevice(dev_);
    cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking);
    own_stream_ = stream_;
    cudaEventCreate(&ev0_);
    cudaEventCreate(&ev1_);
    cudaEventCreate(&ev2_);
    cudaEventCreate(&ev3_);
    cudaMalloc(&d_cnt_, sizeof(DCounters));
    cudaMallocHost(&h_cnt_, sizeof(DCounters));
}

Engine::~Engine() {
    cudaSetDevice(dev_);
    cudaDeviceSynchronize();
    void* ptrs[] = {d_nodes_, d_leaf_, d_hdr_, d_geom_, d_mats_, d_srcs_, d_pats_, d_bsdfs_, d_bsdfbases_, d_bsdfpool_, d_otrack_, d_top_, d_bins_, q_[0], q_[1],
                    h_[0], h_[1], d_hits_, dq_, d_cnt_, d_acc_, d_vacc_, d_rays_, d_out_, d_res_, d_slow_, d_mid_, d_lean_, d_spec_};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h_cnt_) cudaFreeHost(h_cnt_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (ev2_) cudaEventDestroy(ev2_);
    if (ev3_) cudaEventDestroy(ev3_);
    for (auto& e : wev_) if (e) cudaEventDestroy(e);
    if (own_stream_) cudaStreamDestroy(own_stream_);
}

bool Engine::recycle() {
    cudaSetDevice(dev_);
    if (cudaStreamSynchronize(stream_) != cudaSuccess) { cudaGetLastError(); return false; }
    stream_ = own_stream_; user_stream_ = false;
    stats = EngineStats();
    qcap_req_ = 0; nbins_ = 0; ncols_ = 0;
    objdesc.clear();
    // keep what is cheap to hold: queues up to 1 M rays (~250 MB) and scene tables up to 64 MB
    const size_t scene_bytes = cap_nodes_ + cap_leaf_ + cap_hdr_ + cap_geom_ + cap_top_;
    if (qcap_ > ((size_t)1 << 20) || scene_bytes > ((size_t)64 << 20) ||
        acc_bytes_ + out_bytes_ + vacc_bytes_ + rays_bytes_ + res_bytes_ > ((size_t)64 << 20))
        return false;
    return true;
}

template <class T>
static bool upload(void*& dptr, size_t& cap, const std::vector<T>& v, std::string& err) {
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    if (!dptr || cap < bytes) {              // an engine taken from the pool re-uses what it has
        if (dptr) { cudaFree(dptr); dptr = nullptr; cap = 0; }
        CK(cudaMalloc(&dptr, bytes));
        cap = bytes;
    }
    if (!v.empty()) CK(cudaMemcpy(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return true;
}

bool Engine::upload_scene(const FlatScene& fs, const Scene& sc, std::string& err) {
    CK(cudaSetDevice(dev_));
    if (!upload(d_nodes_, cap_nodes_, fs.nodes, err) || !upload(d_leaf_, cap_leaf_, fs.leaf2, err) ||
        !upload(d_hdr_, cap_hdr_, fs.objhdr, err) || !upload(d_geom_, cap_geom_, fs.geom, err) ||
        !upload(d_mats_, cap_mats_, fs.mats, err) || !upload(d_srcs_, cap_srcs_, fs.srcs, err) ||
        !upload(d_pats_, cap_pats_, fs.pats, err) || !upload(d_bsdfs_, cap_bsdfs_, fs.bsdfs, err) ||
        !upload(d_bsdfbases_, cap_bsdfbases_, fs.bsdfbases, err) || !upload(d_bsdfpool_, cap_bsdfpool_, fs.bsdfpool, err))
        return false;
    std::vector<int> ot(sc.objs.size(), -1);
    if (!upload(d_otrack_, cap_otrack_, ot, err)) return false;
    // ---- level-K cell table of the integer walk ----
    int depth = 0;
    {   // deepest cube level (root = 0), from the flat node array itself
        std::vector<std::pair<int, int>> st;
        if (fs.root >= 0) st.push_back({fs.root, 0});
        while (!st.empty()) {
            auto [nd, l] = st.back(); st.pop_back();
            depth = std::max(depth, l + 1);
            for (int k = 0; k < 8; k++) { int w = fs.nodes[(size_t)nd * 8 + k]; if (w >= 0) st.push_back({w, l + 1}); }
        }
    }
    depth = std::max(depth, 1);
    if (depth > RB_MAXDEPTH) {
        err = "octree is " + std::to_string(depth) + " levels deep; this engine walks at most " + std::to_string(RB_MAXDEPTH);
        return false;
    }
    int K = 8;                                  // 8^8 cells x 8 B = 134 MB; only the cells rays cross are ever read
    if (const char* e = getenv("RB_TOPK")) K = std::max(1, std::min(8, atoi(e)));      // developer knob
    K = std::min(K, depth);
    {
        const size_t ncell = (size_t)1 << (3 * K);
        std::vector<int2> top(ncell);
        // fill by walking the tree once: a cube at level l <= K covers 8^(K-l) cells
        struct It { int w, l; unsigned x, y, z; };
        std::vector<It> st;
        st.push_back({fs.root, 0, 0u, 0u, 0u});
        while (!st.empty()) {
            It t = st.back(); st.pop_back();
            if (t.w >= 0 && t.l < K) {
                for (int k = 0; k < 8; k++)
                    st.push_back({fs.nodes[(size_t)t.w * 8 + k], t.l + 1, (t.x << 1) | (unsigned)(k & 1),
                                  (t.y << 1) | (unsigned)((k >> 1) & 1), (t.z << 1) | (unsigned)((k >> 2) & 1)});
                continue;
            }
            const int s = K - t.l;
            const unsigned n = 1u << s;
            for (unsigned z = 0; z < n; z++)
                for (unsigned y = 0; y < n; y++)
                    for (unsigned x = 0; x < n; x++)
                        top[(size_t)((t.x << s) + x) | ((size_t)((t.y << s) + y) << K) | ((size_t)((t.z << s) + z) << (2 * K))] =
                            make_int2(t.w, t.l);
        }
        if (!upload(d_top_, cap_top_, top, err)) return false;
    }
    S_.top = (const int2*)d_top_; S_.topk = K;
    S_.inv_cell = (double)(1u << depth) / sc.cusize;
    for (int k = 0; k < 3; k++) S_.cuorg[k] = sc.cuorg[k];
    S_.cusize = sc.cusize;
    S_.root = fs.root; S_.nobjs = (int)sc.objs.size(); S_.nsrcs = (int)fs.srcs.size();
    nsrc_active_ = 0;
    for (const SrcRec& sr : fs.srcs)              // local sources: a few partitions each on average
        nsrc_active_ += (sr.flags & SF_SKIP) ? 0 : (sr.flags & SF_DISTANT) ? 1 : 4;
    S_.maxdepth = depth;
    S_.nodes = (const int*)d_nodes_; S_.leafpool = (const int*)d_leaf_;
    S_.objhdr = (const int4*)d_hdr_; S_.geom = (const double*)d_geom_;
    S_.mats = (const MatRec*)d_mats_; S_.srcs = (const SrcRec*)d_srcs_; S_.pats = (const PatRec*)d_pats_;
    S_.bsdfs = (const BsdfRec*)d_bsdfs_; S_.bsdfbases = (const BsdfBasis*)d_bsdfbases_; S_.bsdfpool = (const unsigned*)d_bsdfpool_;
    S_.otrack = (const int*)d_otrack_;
    has_local_sources_ = false;                  // local emitters: direct() always runs in k_direct
    for (const auto& s : fs.srcs)
        if (!(s.flags & SF_DISTANT) && !(s.flags & SF_SKIP)) has_local_sources_ = true;
    objdesc.clear();
    for (const auto& o : sc.objs) objdesc.push_back(o.tname + " \"" + o.name + "\"");
    nbins_ = 0; ncols_ = 0;
    return true;
}

std::string Engine::describe_obj(unsigned idx) const {
    if (idx < objdesc.size()) return objdesc[idx];
    return "object #" + std::to_string(idx);
}

bool Engine::set_bins(const std::vector<DBinSpec>& bins, const std::vector<int>& otrack, int ncols,
                      std::string& err) {
    CK(cudaSetDevice(dev_));
    if (!upload(d_bins_, cap_bins_, bins, err)) return false;
    if ((int)otrack.size() != S_.nobjs) { err = "internal: otrack size"; return false; }
    if (!otrack.empty()) CK(cudaMemcpy(d_otrack_, otrack.data(), otrack.size() * sizeof(int), cudaMemcpyHostToDevice));
    nbins_ = (int)bins.size();
    ncols_ = ncols;
    return true;
}

// `hint` = rays per queue that would let a batch hold a few hundred records of the coming job
// (0: none).  Queues only ever grow: 2 x 48 M rays by default, more when a job's records are
// wide (thousands of suns per shading point) and HBM is there -- 180 GB per B200.
bool Engine::ensure_queues(std::string& err, size_t hint) {
    if (park_direct() && !dq_ && q_[0]) {   // a many-source scene loaded after the queues were made
        CK(cudaMalloc(&dq_, dcap_ * sizeof(DirectJob)));
    }
    size_t want = qcap_req_ ? qcap_req_ : (hint ? hint : kDefaultQueue);     // rays per queue
    if (q_[0] && want <= qcap_) return true;
    if (q_[0]) {                            // grow: drop the old queues first
        CK(cudaStreamSynchronize(stream_));
        void* old[] = {q_[0], q_[1], h_[0], h_[1], d_hits_, dq_, d_slow_, d_mid_, d_lean_, d_spec_};
        for (void* p : old) if (p) cudaFree(p);
        q_[0] = q_[1] = nullptr; h_[0] = h_[1] = nullptr; d_hits_ = nullptr; dq_ = nullptr; d_slow_ = nullptr; d_mid_ = nullptr; d_lean_ = nullptr; d_spec_ = nullptr;
    }
    size_t freeb = 0, totalb = 0;
    CK(cudaMemGetInfo(&freeb, &totalb));
    size_t maxq = (freeb / 3) / (2 * sizeof(QRay));              // at most a third of free HBM
    if (want > maxq) want = maxq;
    if (want < 4096) { err = "not enough device memory for ray queues"; return false; }
    qcap_ = want;
    hcap_ = std::max<size_t>(qcap_ / 16, 4096);
    CK(cudaMalloc(&q_[0], qcap_ * sizeof(QRay)));
    CK(cudaMalloc(&q_[1], qcap_ * sizeof(QRay)));
    CK(cudaMalloc(&h_[0], hcap_ * sizeof(QHemi)));
    CK(cudaMalloc(&h_[1], hcap_ * sizeof(QHemi)));
    CK(cudaMalloc(&d_hits_, qcap_ * sizeof(HitRec)));
    CK(cudaMalloc(&d_slow_, qcap_ * sizeof(unsigned)));
    CK(cudaMalloc(&d_mid_, qcap_ * sizeof(unsigned)));
    CK(cudaMalloc(&d_lean_, qcap_ * sizeof(unsigned)));
    CK(cudaMalloc(&d_spec_, qcap_ * sizeof(unsigned)));
    dcap_ = std::max<size_t>(qcap_ / 32, 4096);
    if (park_direct()) {      // many or local sources: direct() runs as its own kernel from a job queue
        CK(cudaMalloc(&dq_, dcap_ * sizeof(DirectJob)));
    }
    return size_trace_grid(err);
}

// Persistent k_trace grid: every SM filled exactly once.  Depends on the octree
// depth through the dynamic shared memory of the ancestor stack.
bool Engine::size_trace_grid(std::string& err) {
    int per_sm = 0, nsm = 0;
    // ancestor stack of the levels below the cell table: interior nodes at levels topk .. maxdepth - 1
    trace_smem_ = (size_t)std::max(1, S_.maxdepth - S_.topk + 1) * WAVE_THREADS * sizeof(int);
    if (const char* e = getenv("RB_TRACE_CARVEOUT"))      // developer knob: shared-memory share of the SM's 256 KB, percent
        CK(cudaFuncSetAttribute(k_trace, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace, WAVE_THREADS, trace_smem_));
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev_));
    if (const char* e = getenv("RB_TRACE_CTAS_PER_SM")) per_sm = std::min(per_sm, std::max(1, atoi(e)));   // developer knob
    trace_blocks_ = std::max(1, per_sm) * std::max(1, nsm);
    if (getenv("RB_DEBUG_GRID")) fprintf(stderr, "[rb] k_trace: %d CTAs/SM x %d threads, %zu B dynamic smem\n", per_sm, WAVE_THREADS, trace_smem_);
    return true;
}

template <class T>
static bool ensure_buf(T*& p, size_t& have, size_t bytes, std::string& err) {
    if (bytes <= have && p) return true;
    if (p) cudaFree(p);
    p = nullptr; have = 0;
    CK(cudaMalloc(&p, std::max<size_t>(bytes, 256)));
    have = bytes;
    return true;
}

bool Engine::run_batch(const TraceJob& job, const DParams& P, size_t rec0, size_t nrec, std::string& err,
                       bool& overflow) {
    overflow = false;
    const int accum = job.accum;
    const size_t ray0 = accum > 0 ? rec0 * accum : 0;
    const size_t nray = accum > 0 ? std::min(job.nrays - ray0, nrec * (size_t)accum) : job.nrays;
    const bool per_ray = job.results != nullptr || (job.values != nullptr && job.cmat == nullptr);
    // rtrace-style jobs report per input ray: rows == rays
    const size_t nrows = per_ray ? nray : nrec;
    const bool want_c = job.cmat != nullptr && ncols_ > 0;
    const bool want_v = job.values != nullptr;
    // ---- stage inputs ----
    const double* d_rays = nullptr;
    if (job.rays_on_device) d_rays = job.rays + ray0 * 6;
    else {
        if (!ensure_buf(d_rays_, rays_bytes_, nray * 6 * sizeof(double), err)) return false;
        CK(cudaMemcpyAsync(d_rays_, job.rays + ray0 * 6, nray * 6 * sizeof(double), cudaMemcpyHostToDevice, stream_));
        d_rays = d_rays_;
    }
    size_t accn = want_c ? nrows * (size_t)ncols_ * 3 : 0;
    if (want_c) {
        if (!ensure_buf(d_acc_, acc_bytes_, accn * sizeof(double), err)) return false;
        CK(cudaMemsetAsync(d_acc_, 0, accn * sizeof(double), stream_));
    }
    if (want_v) {
        if (!ensure_buf(d_vacc_, vacc_bytes_, nrows * 3 * sizeof(double), err)) return false;
        CK(cudaMemsetAsync(d_vacc_, 0, nrows * 3 * sizeof(double), stream_));
    }
    if (job.results) {
        if (!ensure_buf(d_res_, res_bytes_, nray * sizeof(RayResult), err)) return false;
    }
    CK(cudaMemsetAsync(d_cnt_, 0, sizeof(DCounters), stream_));

    WaveArgs A;
    A.S = S_; A.P = P;
    A.bins = (const DBinSpec*)d_bins_; A.nbinspecs = nbins_;
    A.acc = want_c ? d_acc_ : nullptr; A.ncols = ncols_;
    A.vacc = want_v ? d_vacc_ : nullptr;
    A.row0 = 0;
    A.res = job.results ? d_res_ : nullptr;
    A.C = d_cnt_;
    A.inline_hemi_max = 16;
    A.qcap = (unsigned)qcap_; A.hcap = (unsigned)hcap_;
    A.hits = d_hits_;
    A.dout = park_direct() ? dq_ : nullptr; A.dcap = (unsigned)dcap_;
    A.slow = getenv("RB_NO_SHADE_SPLIT") ? nullptr : d_slow_;
    A.mid = d_mid_; A.lean = d_lean_; A.spec = d_spec_;
    A.nodirect = nsrc_active_ == 0 ? 1 : 0;
    A.anyhit = (RB_SHADOW_ANYHIT && !A.nodirect && P.backvis && !getenv("RB_NO_ANYHIT")) ? 1 : 0;      // (the variable: for the test that both give one matrix)

    auto sync_counters = [&](std::string& err) -> bool {
        CK(cudaMemcpyAsync(h_cnt_, d_cnt_, sizeof(DCounters), cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        return true;
    };
    auto timed = [&](double& bucket, std::string& err) -> bool {   // call after sync
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ev0_, ev1_));
        bucket += ms;
        return true;
    };

    int cur = 0;
    unsigned long long batch_rays = 0;
    // ---- k_init ----
    {
        InitArgs I;
        I.rays = d_rays; I.nrays = (unsigned)nray;
        I.ray0 = job.row_base * (unsigned long long)(accum > 0 ? accum : 1) + ray0;
        I.accum = per_ray ? 1 : accum; I.irrad = job.irrad; I.lim_dist = job.lim_dist;
        if (per_ray) I.accum = 1;
        A.qin = nullptr; A.nin = 0; A.qout = q_[cur]; A.hout = h_[cur];
        unsigned grid = (unsigned)((nray + WAVE_THREADS - 1) / WAVE_THREADS);
        CK(cudaEventRecord(ev0_, stream_));
        k_init<<<grid, WAVE_THREADS, 0, stream_>>>(A, I);
        CK(cudaEventRecord(ev1_, stream_));
        stats.launches++;
        CK(cudaGetLastError());
    }
    // ---- waves ----
    // Every kernel of a wave takes its counts from the device (k_gate / k_prepare), so the host may queue waves ahead.
    // It does so for SMALL waves, where a host round trip per wave would be most of the time (chunks of up to 8 waves,
    // grid-stride launches).  A LARGE wave is launched alone, with the one-thread-per-ray grid its known upper bound
    // gives (27 % faster shading than the grid-stride form) and one look at the counters after it.
    const int kChunkMax = 8;
    const size_t kSmallWave = (size_t)1 << 19;
    if (!wev_[0]) for (auto& e : wev_) CK(cudaEventCreate(&e));
    unsigned fgrid = 148u * 12u;                   // grid-stride kernels: the count is only known on the device
    if (const char* e = getenv("RB_FGRID")) fgrid = 148u * (unsigned)std::max(1, atoi(e));      // developer knob
    if (!sync_counters(err)) return false;         // what k_init produced
    { float ms = 0; CK(cudaEventElapsedTime(&ms, ev0_, ev1_)); stats.kernel_ms += ms; }
    bool finished = h_cnt_->nq_out == 0 && h_cnt_->nh_out == 0 && h_cnt_->nd_out == 0;
    int wave = 0;
    while (!finished && wave < 4096) {
        if (h_cnt_->overflow) { overflow = true; return true; }
        if (h_cnt_->errflag) break;
        // upper bound of the coming wave: queued rays + the rays its hemispheres reserved + a shadow ray per parked
        // direct() job and source sample (local sources split into at most 64 partitions, srcsamp.c MAXSPART)
        size_t ub = (size_t)h_cnt_->nq_out + h_cnt_->hemi_rays +
                    (size_t)h_cnt_->nd_out * (size_t)std::max(1, S_.nsrcs) * (has_local_sources_ ? 64 : 1);
        ub = std::min(ub, qcap_);
        const bool big = ub >= kSmallWave;
        const int chunk = big ? 1 : kChunkMax;
        const unsigned sgrid = big ? (unsigned)((ub + RB_SHADE_THREADS - 1) / RB_SHADE_THREADS) : fgrid;
        const unsigned egrid = std::max(1u, std::min<unsigned>(big ? h_cnt_->nh_out : 148u * 8u, 148u * 8u));
        const unsigned dgrid = std::max(1u, std::min<unsigned>(big ? h_cnt_->nd_out : 148u * 8u, 148u * 8u));
        const int w0 = wave;
        for (int k = 0; k < chunk; k++, wave++) {
            CK(cudaEventRecord(wev_[4 * k + 3], stream_));
            k_gate<<<1, 1, 0, stream_>>>(d_cnt_, (unsigned)qcap_);
            ExpandArgs E; E.hin = h_[cur];
            A.qout = q_[cur];
            k_expand<<<egrid, 256, 0, stream_>>>(A, E);
            stats.launches += 2;
            if (park_direct()) {
                DirectArgs D; D.din = dq_;
                k_direct<<<dgrid, 256, 0, stream_>>>(A, D);
                stats.launches++;
            }
            k_prepare<<<1, 1, 0, stream_>>>(d_cnt_, (unsigned)wave);
            A.qin = q_[cur]; A.nin = 0; A.qout = q_[cur ^ 1]; A.hout = h_[cur ^ 1];
            CK(cudaEventRecord(wev_[4 * k + 0], stream_));
            k_trace<<<(unsigned)trace_blocks_, WAVE_THREADS, trace_smem_, stream_>>>(A);
            CK(cudaEventRecord(wev_[4 * k + 1], stream_));
            if (A.slow) {
                k_shade_fast<<<sgrid, RB_SHADE_THREADS, 0, stream_>>>(A);
                k_shade_spec<<<big ? 148u * 24u : 148u * 6u, RB_SHADE_THREADS, 0, stream_>>>(A);
                k_shade_lean<<<big ? 148u * 12u : 148u * 6u, RB_SHADE_THREADS, 0, stream_>>>(A);
                k_shade_mid<<<148u * 6u, RB_SHADE_THREADS, 0, stream_>>>(A);
                k_shade<<<148u * 4u, RB_SHADE_THREADS, 0, stream_>>>(A);
                stats.launches += 4;
            } else
                k_shade<<<sgrid, RB_SHADE_THREADS, 0, stream_>>>(A);
            CK(cudaEventRecord(wev_[4 * k + 2], stream_));
            stats.launches += 3;
            cur ^= 1;
        }
        CK(cudaGetLastError());
        if (!sync_counters(err)) return false;
        for (int k = 0; k < chunk; k++) {
            const unsigned n = h_cnt_->wave_nin[(w0 + k) & 63];
            float ms = 0, ms2 = 0, ms0 = 0;
            CK(cudaEventElapsedTime(&ms0, wev_[4 * k + 3], wev_[4 * k + 0]));     // gate, expand, direct, prepare
            CK(cudaEventElapsedTime(&ms, wev_[4 * k + 0], wev_[4 * k + 1]));
            CK(cudaEventElapsedTime(&ms2, wev_[4 * k + 1], wev_[4 * k + 2]));
            stats.kernel_ms += ms0 + ms + ms2;
            if (n) { stats.wave_ms += ms; stats.shade_ms += ms2; stats.wave_launches++; stats.waves++; }
            if (getenv("RB_DEBUG_WAVES"))
                fprintf(stderr, "[rb] wave %d: %u rays (bound %zu) trace %.3f ms shade %.3f ms; last wave of the chunk: spec %u lean %u mid %u slow %u\n",
                        w0 + k, n, ub, ms, ms2, h_cnt_->nspec, h_cnt_->nlean, h_cnt_->nmid, h_cnt_->nslow);
        }
        finished = h_cnt_->nq_out == 0 && h_cnt_->nh_out == 0 && h_cnt_->nd_out == 0;
    }
    if (h_cnt_->overflow) { overflow = true; return true; }
    batch_rays = h_cnt_->rays_traced;
    const unsigned nq = h_cnt_->nq_out, nh = h_cnt_->nh_out, nd = h_cnt_->nd_out;
    if (!h_cnt_->errflag && (nq > 0 || nh > 0 || nd > 0)) {
        err = "ray generations did not end after 4096 waves (" + std::to_string(nq) + " rays still queued)";
        return false;
    }
    if (h_cnt_->errflag) {
        unsigned f = h_cnt_->errflag;
        std::string what = describe_obj(h_cnt_->errobj);
        if (f & RB_ERR_LOCAL_SRC) err = "unsupported: local light source material " + what + " (only distant sources are built)";
        else if (f & RB_ERR_UNSUP_MAT) err = "unsupported material " + what + " reached by a ray (no CPU fallback)";
        else if (f & RB_ERR_CONTRIB_VALUE)
            err = "unsupported: -V+ (contributions) with the tracked modifier on " + what +
                  ", which does not emit: the value a reflecting / transmitting surface returns is not available "
                  "to this engine (coefficients, -V-, are)";
        else if (f & RB_ERR_UNSUP_PRIM) err = "unsupported surface " + what + " reached by a ray (no CPU fallback)";
        else err = "unsupported modifier on " + what + " reached by a ray (patterns/textures/mixtures are not built)";
        return false;
    }
#if RB_WALK_STATS
    if (getenv("RB_DEBUG_DBG")) { k_dbg_print<<<1, 1, 0, stream_>>>(); cudaStreamSynchronize(stream_); }
#endif
    stats.nrays += batch_rays; stats.nodes += h_cnt_->nodes; stats.leafents += h_cnt_->leafents;
    stats.prims += h_cnt_->prims; stats.contribs += h_cnt_->contribs; stats.badbin += h_cnt_->badbin;
    // ---- outputs ----
    if (want_c) {
        double scale = accum > 1 ? 1.0 / accum : 1.0;
        const size_t esz = job.cmat_double ? sizeof(double) : sizeof(float);
        char* dst;
        size_t off = rec0 * (size_t)ncols_ * 3;
        if (job.cmat_on_device) dst = (char*)job.cmat + off * esz;
        else {
            if (!ensure_buf(d_out_, out_bytes_, accn * esz, err)) return false;
            dst = (char*)d_out_;
        }
        unsigned grid = (unsigned)std::min<size_t>((accn + 255) / 256, 148 * 16);
        CK(cudaEventRecord(ev0_, stream_));
        size_t tail_start = accn;
        double tail_scale = scale;
        if (accum > 1 && ray0 + nray == job.nrays && job.nrays % accum != 0) {
            tail_start = (nrows - 1) * (size_t)ncols_ * 3;
            tail_scale = 1.0 / (double)(job.nrays % accum);
        }
        if (job.cmat_double)
            k_finish<double><<<grid, 256, 0, stream_>>>(d_acc_, (double*)dst, accn, scale, tail_start, tail_scale);
        else
            k_finish<float><<<grid, 256, 0, stream_>>>(d_acc_, (float*)dst, accn, scale, tail_start, tail_scale);
        CK(cudaEventRecord(ev1_, stream_));
        stats.launches++;
        if (!job.cmat_on_device)
            CK(cudaMemcpyAsync((char*)job.cmat + off * esz, d_out_, accn * esz, cudaMemcpyDeviceToHost, stream_));
        CK(cudaStreamSynchronize(stream_));
        if (!timed(stats.kernel_ms, err)) return false;
    }
    if (want_v) {
        size_t off = (per_ray ? ray0 : rec0) * 3;
        CK(cudaMemcpyAsync(job.values + off, d_vacc_, nrows * 3 * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    }
    if (job.results)
        CK(cudaMemcpyAsync(job.results + ray0, d_res_, nray * sizeof(RayResult), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    stats.batches++;
    return true;
}

bool Engine::run(const TraceJob& job, const DParams& P, std::string& err) {
    CK(cudaSetDevice(dev_));
    if (!d_nodes_) { err = "no octree loaded"; return false; }
    if (job.nrays == 0) return true;
    const int accum = job.accum;
    const size_t nrec_total = accum > 0 ? (job.nrays + accum - 1) / accum : 1;
    // estimate rays per record of the largest wave to size batches
    double per_rec = 1.0;
    if (P.ambounce > 0 && P.ambdiv > 0) {
        double wt = 1.0, d = RB_PI * 0.8 / (P.ambdiv * (double)P.minweight + 1e-20);
        if (wt > d) wt = d;
        int n = (int)(sqrt(P.ambdiv * wt) + .5);
        if (n < 1) n = 1;
        per_rec = (double)n * n;
    }
    // every shading point also sends one shadow ray per active source; the
    // widest wave is the first-bounce one (about half the sources face a surface)
    per_rec = per_rec * (1.0 + 0.6 * nsrc_active_) + nsrc_active_;
    per_rec *= (accum > 0 ? accum : 1);
    {   // queues wide enough for the whole job when it is small (a 10 k-ray rtrace call must not pay for
        // 10 GB of cudaMalloc), else for ~256 records per batch or the default 2 x 48 M rays, memory permitting
        const double whole = per_rec * (double)nrec_total / 0.45 * 1.25;
        const double want_rec = (double)std::min<size_t>(nrec_total, 256);
        size_t hint = (size_t)(per_rec * want_rec / 0.45);
        if (whole < (double)kDefaultQueue) hint = std::max<size_t>((size_t)whole, (size_t)1 << 16);
        else hint = std::max<size_t>(hint, kDefaultQueue);
        if (!ensure_queues(err, hint) || !size_trace_grid(err)) return false;
    }
    size_t batch = (size_t)std::max(1.0, (double)qcap_ * 0.45 / per_rec);
    if (job.cmat && ncols_ > 0) {          // the batch's accumulators (double) and output rows must fit too
        size_t freeb = 0, totalb = 0;
        CK(cudaMemGetInfo(&freeb, &totalb));
        const size_t per_row = (size_t)ncols_ * 3 * (sizeof(double) + (job.cmat_on_device ? 0 : (job.cmat_double ? 8 : 4)));
        const size_t budget = (freeb + acc_bytes_ + out_bytes_) / 2;
        batch = std::min(batch, std::max<size_t>(1, budget / per_row));
    }
    if (accum <= 0) batch = 1;
    if (batch < nrec_total) {              // equal batches: a short last batch would run its waves on a half-empty GPU
        const size_t nb = (nrec_total + batch - 1) / batch;
        batch = (nrec_total + nb - 1) / nb;
    }
    size_t rec = 0;
    while (rec < nrec_total) {
        size_t n = std::min(batch, nrec_total - rec);
        bool ovf = false;
        if (!run_batch(job, P, rec, n, err, ovf)) return false;
        if (ovf) {
            stats.retries++;
            if (!qcap_req_ && qcap_ < kDefaultQueue) {       // a small job's queues were sized by an estimate: grow them first
                if (!ensure_queues(err, std::min<size_t>(kDefaultQueue, qcap_ * 4)) || !size_trace_grid(err)) return false;
                batch = (size_t)std::max(1.0, (double)qcap_ * 0.45 / per_rec);
                if (accum <= 0) batch = 1;
                continue;
            }
            if (n <= 1) { err = "ray queue overflow on a single record; raise the queue capacity"; return false; }
            batch = std::max<size_t>(1, n / 2);
            continue;
        }
        rec += n;
    }
    return true;
}

}  // namespace rb
