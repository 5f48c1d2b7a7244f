/* rb200.h -- C ABI of the B200-native rtrace / rcontrib hot path.
 *
 * Drop-in boundary for LBNL-ETA/pyradiance: these are the entry points a
 * binding (nanobind, ctypes, cgo ...) needs in order to replace, for this path
 * only, what src/binding/radiance_ext.cpp reaches in the reference:
 *
 *   rb_create / rb_destroy        RtraceSimulManager / RcontribSimulManager
 *                                 ctor + Cleanup  (src/binding/radiance_ext.cpp:170-231,298-445;
 *                                 src/radiance/rt/RtraceSimulManager.h:86-177)
 *   rb_load_octree                LoadOctree()  (rt/RtraceSimulManager.cpp:237-260,
 *                                 rt/RcontribSimulManager.h:235-242) = readoct() + marksources()
 *   rb_set_option                 getrenderopt() (rt/renderopts.c:123-349), set_option
 *                                 (radiance_ext.cpp:125-152)
 *   rb_get_params / rb_set_params ray_save / ray_restore (rt/raycalls.c:247-423),
 *                                 get_ray_params / set_ray_params (radiance_ext.cpp:125-152)
 *   rb_set_defaults               ray_defaults (rt/raycalls.c:380-423) vs rcontrib's own
 *                                 defaults (rt/rcontrib.c:24-58)
 *   rb_cal_load / rb_cal_set /    loadfunc / set_eparams + scompile / eval
 *   rb_cal_eval                   (radiance_ext.cpp:554-560; rt/func.c:76-119) -- only the
 *                                 names of the known bin-function files are understood
 *   rb_add_modifier               AddModifier() (rt/RcontribSimulManager.cpp:261-359),
 *                                 addmodifier() (rt/rcontrib.c:92-160)
 *   rb_rcontrib                   ComputeRecord() loop (rt/RcontribSimulManager.cpp:668-714),
 *                                 rcontrib() main loop (rt/rcontrib.c:379-427)
 *   rb_rtrace                     EnqueueBundle() (rt/RtraceSimulManager.cpp:352-403),
 *                                 rtcompute() (rt/rtrace.c:436-468)
 *
 * Plain pointers and sizes only.  Every function returns 0 on success or a
 * negative value on failure, with the message available from rb_last_error().
 * There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef RB200_H
#define RB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_ctx rb_ctx;

/* rendering parameters: field-for-field the RAYPARAMS of rt/ray.h:157-190 */
typedef struct rb_params {
    int do_irrad;          /* -i  */
    int rand_samp;         /* -u  */
    double dstrsrc;        /* -dj */
    double shadthresh;     /* -dt */
    double shadcert;       /* -dc */
    int directrelay;       /* -dr */
    int vspretest;         /* -dp */
    int directvis;         /* -dv */
    double srcsizerat;     /* -ds */
    double cextinction[3]; /* -me */
    double salbedo[3];     /* -ma */
    double seccg;          /* -mg */
    double ssampdist;      /* -ms */
    double specthresh;     /* -st */
    double specjitter;     /* -ss */
    int backvis;           /* -bv */
    int maxdepth;          /* -lr */
    double minweight;      /* -lw */
    double ambval[3];      /* -av */
    int ambvwt;            /* -aw */
    double ambacc;         /* -aa */
    int ambres;            /* -ar */
    int ambdiv;            /* -ad */
    int ambssamp;          /* -as */
    int ambounce;          /* -ab */
} rb_params;

/* per-ray report of rb_rtrace (the fields rtrace's -o spec can print) */
typedef struct rb_ray_result {
    double rop[3];         /* -op intersection point */
    double ron[3];         /* -oN unperturbed normal, as intersected */
    double rot;            /* -oL distance (1e10 = none) */
    double rod;
    int32_t robj;          /* -os surface object index, -1 none */
    int32_t omod;          /* -om modifier object index, -1 none */
    float rweight;         /* -ow */
    int32_t pad;           /* 1: the material reversed the surface (hit from behind; flipsurface), so -on = -(...) */
    double pert[3];        /* normal perturbation (smooth mesh triangles; 0 otherwise): -on = raynormal(ron + pert) */
} rb_ray_result;

/* counters of the last compute call (device-side, summed over launches) */
typedef struct rb_stats {
    uint64_t nrays;        /* rays walked through the octree (localhit calls) */
    uint64_t nodes;        /* octree child words read   } counted only by developer builds   */
    uint64_t leafents;     /* leaf-set entries read     } (-DRB_WALK_STATS=1); 0 otherwise:  */
    uint64_t prims;        /* primitive intersection tests } the walker keeps no statistics registers */
    uint64_t contribs;     /* contributions accumulated */
    uint64_t launches;     /* kernels launched */
    uint64_t wave_launches;/* launches of the dominant (trace) kernel */
    uint64_t waves, batches, retries, badbin;
    double kernel_ms;      /* device time of all kernels (CUDA events) */
    double wave_ms;        /* device time of the trace kernel (octree walk + intersection) */
    double shade_ms;       /* device time of the shade kernel */
} rb_stats;

/* flags of rb_rcontrib / rb_rtrace */
#define RB_IRRAD_NONE      0
#define RB_IRRAD_RTRACE    1   /* rtrace -I     (rt/rtrace.c:443-448,415-432) */
#define RB_IRRAD_RCONTRIB  2   /* rcontrib -I   (rt/rcontrib.c:321-339) */
#define RB_IRRAD_MANAGER   3   /* RTimmIrrad    (rt/RtraceSimulManager.cpp:315-335) */
#define RB_FLAG_IRRAD_MASK 3u
#define RB_FLAG_LIMDIST    4u  /* -ld / RTlimDist */
#define RB_FLAG_CONTRIB    8u  /* rcontrib -V+ */
#define RB_FLAG_RAYS_ON_DEVICE 16u
#define RB_FLAG_OUT_ON_DEVICE  32u
#define RB_FLAG_OUT_DOUBLE     64u  /* out is float64 instead of float32 (rcontrib -fd) */

#define RB_PROGRAM_RTRACE   0
#define RB_PROGRAM_RCONTRIB 1

rb_ctx* rb_create(int cuda_device);
void rb_destroy(rb_ctx* ctx);
const char* rb_last_error(rb_ctx* ctx);
const char* rb_version(void);
/* CUDA devices visible to this process (0 without a usable driver): what `-n N` can spread records over */
int rb_device_count(void);

int rb_set_defaults(rb_ctx* ctx, int program);
int rb_get_params(rb_ctx* ctx, rb_params* out);
int rb_set_params(rb_ctx* ctx, const rb_params* in);
/* parse one render option at argv[0]; returns the number of EXTRA arguments
 * consumed (>= 0) or -1 if argv[0] is not a render option of this path */
int rb_set_option(rb_ctx* ctx, int argc, const char* const* argv);

int rb_load_octree(rb_ctx* ctx, const char* path);
/* write the loaded (instance-expanded) scene as a frozen octree any Radiance program reads */
int rb_save_octree(rb_ctx* ctx, const char* path);
/* scene queries */
int rb_num_objects(rb_ctx* ctx);
const char* rb_object_name(rb_ctx* ctx, int obj);
const char* rb_object_type(rb_ctx* ctx, int obj);
int rb_object_modifier(rb_ctx* ctx, int obj);
int rb_num_header_lines(rb_ctx* ctx);
const char* rb_header_line(rb_ctx* ctx, int i);
const char* rb_scene_warnings(rb_ctx* ctx);

/* calcomp stand-in for the known bin-function files */
int rb_cal_load(rb_ctx* ctx, const char* calfile);
int rb_cal_set(rb_ctx* ctx, const char* assignments);   /* "MF:4" or "MF=4,rNx=0,..." */
int rb_cal_eval(rb_ctx* ctx, const char* expr, double* value);

int rb_clear_modifiers(rb_ctx* ctx);
/* returns the first column of this modifier in the output row, or < 0 */
int rb_add_modifier(rb_ctx* ctx, const char* modname, const char* params,
                    const char* binexpr, int nbins);
int rb_num_columns(rb_ctx* ctx);
/* host evaluation of a tracked modifier's bin function for direction D (the
 * same code the device runs; used by the CPU-side parity tests of the bins) */
int rb_bin_of_direction(rb_ctx* ctx, int modifier_index, const double dir[3], double* binval);

/* rays: [nrays][6] doubles, origin then direction (zero direction = dummy).
 * out:  [nrecords][ncols][3] float32 (float64 with RB_FLAG_OUT_DOUBLE),
 *       nrecords = ceil(nrays / accum); out_floats = number of elements.
 * row_base: global index of the first record (keeps RNG streams independent
 * of how records are sharded across GPUs). */
int rb_rcontrib(rb_ctx* ctx, const double* rays, size_t nrays, int accum,
                unsigned flags, uint64_t row_base, void* out, size_t out_floats);

/* values: [nrays][3] doubles (may be NULL), results: [nrays] (may be NULL) */
int rb_rtrace(rb_ctx* ctx, const double* rays, size_t nrays, unsigned flags,
              double* values, rb_ray_result* results);
/* rb_rtrace over a SHARD of a larger ray set: the global index of the shard's first ray, so that the random
 * streams (keyed by global ray index) do not depend on how the rays were split over calls or GPUs */
int rb_set_row_base(rb_ctx* ctx, uint64_t first_ray);

int rb_get_stats(rb_ctx* ctx, rb_stats* out);
int rb_reset_stats(rb_ctx* ctx);
int rb_set_stream(rb_ctx* ctx, void* cuda_stream);     /* launch on the caller's stream */
int rb_set_seed(rb_ctx* ctx, uint64_t seed);
int rb_set_queue_capacity(rb_ctx* ctx, size_t nrays);
/* device memory helpers so that callers without a CUDA binding can keep
 * inputs/outputs resident (used by bench.py's HBM-resident measurement) */
void* rb_device_alloc(rb_ctx* ctx, size_t bytes);
int rb_device_free(rb_ctx* ctx, void* p);
int rb_device_upload(rb_ctx* ctx, void* dst, const void* src, size_t bytes);
int rb_device_download(rb_ctx* ctx, void* dst, const void* src, size_t bytes);
int rb_device_sync(rb_ctx* ctx);
/* page-lock / unlock a caller-owned host buffer (faster, truly asynchronous copies) */
int rb_host_register(rb_ctx* ctx, void* p, size_t bytes);
int rb_host_unregister(rb_ctx* ctx, void* p);

/* Peer-memory window for the one exchange of the path, the gather of matrix rows (the reference hands
 * every child's rows to one caller, rt/RcontribSimulManager.cpp:677-689, rt/rc3.c:598-622).  The gathering rank
 * allocates the whole [nrecords][ncols][3] matrix with rb_device_alloc and exports it; every other rank (one
 * process per GPU) opens the 64-byte handle and passes `window + its row offset` as `out` of rb_rcontrib with
 * RB_FLAG_OUT_ON_DEVICE: the kernel that finishes a batch of records stores its rows straight into the gathering
 * GPU's HBM over NVLink, batch by batch while the next batch is traced -- no separate gather step. */
#define RB_IPC_HANDLE_BYTES 64
int rb_ipc_export(rb_ctx* ctx, void* device_ptr, void* handle_out /* RB_IPC_HANDLE_BYTES */);
int rb_ipc_open(rb_ctx* ctx, const void* handle, void** device_ptr_out);
int rb_ipc_close(rb_ctx* ctx, void* device_ptr);

/* View rays on the device: util/vwrays.c:245-300 putrays() over common/image.c:214-305 viewray() -- the step before
 * the path when the rays are the pixels of a view (the three-phase view matrix, BASELINE configs[3]: 2048 x 2048
 * rays that never need to exist in host memory).  `view` is a VIEW after setview() (common/image.c:24-127):
 * type letter ('v' perspective, 'l' parallel, 'c' cylinder, 'h' hemispherical, 'a' angular, 's' stereographic
 * fisheye), unit vdir, hvec / vvec already scaled, hn2 / vn2 squared.  Writes xres * yres * repeat rays, scanlines
 * from the top like vwrays (-Y yres +X xres), origin then direction, direction scaled by the aft distance when the
 * view has an aft plane (pass RB_FLAG_LIMDIST to the tracer then), zeros where the view has no ray.  pj = pixel
 * jitter (-pj), drawn from a counter-based generator keyed by `seed` and the ray index.  `out` is host memory, or
 * device memory with RB_FLAG_OUT_ON_DEVICE. */
typedef struct rb_view {
    int type;
    double vp[3], vdir[3], hvec[3], vvec[3];
    double horiz, vert, hoff, voff, vfore, vaft, hn2, vn2;
} rb_view;
int rb_view_rays(rb_ctx* ctx, const rb_view* view, int xres, int yres, int repeat, double pj, uint64_t seed,
                 double* out, unsigned flags);

/* own octree builder for synthetic scenes (next-row f3): text scene -> frozen .oct */
int rb_oconv(const char* rad_path, const char* oct_path, int objlim, int maxres,
             char* errbuf, size_t errlen);
/* oconv -f [-i include_octree] rad_paths...: what rfluxmtx runs to put its receivers into the scene
 * (util/rfluxmtx.c:139-187 oconv_command); include_octree may be NULL */
int rb_oconv_files(const char* const* rad_paths, int npaths, const char* include_octree, const char* oct_path,
                   int objlim, int maxres, char* errbuf, size_t errlen);

/* The matrix consumer right after the path: out[nrows][ncols][3] = a[nrows][ninner][3] x b[ninner][ncols][3],
 * colour channel by colour channel -- util/cmatrix.c:420-475 cm_multiply() as dctimestep / rmtxop use it on a
 * daylight-coefficient matrix and a sky matrix.  fp32 with two-level accumulation (within 1e-5 of the reference's
 * double accumulation).  Buffers are host memory unless flagged; kernel_ms (may be NULL) gets the device time. */
#define RB_MTX_A_ON_DEVICE 1u
#define RB_MTX_B_ON_DEVICE 2u
#define RB_MTX_OUT_ON_DEVICE 4u
int rb_mtx_multiply(rb_ctx* ctx, const float* a, size_t nrows, size_t ninner, const float* b, size_t ncols, float* out,
                    unsigned flags, double* kernel_ms);

/* ASCII output of a value matrix at C speed.  style 0: "%e\t" per value, "\n" per row (rc2.c:304-312
 * put_contrib, rtrace.c:907-918 puta); style 1: "%e %e %e" triplets separated by tabs, "\n" at the end of a row
 * (util/cmatrix.c:492-498 cm_write).  Returns the text length; the text is written when outlen is large enough. */
size_t rb_format_ascii(const void* values, int is_double, size_t nrows, size_t per_row, int style, char* out,
                       size_t outlen);

#ifdef __cplusplus
}
#endif
#endif /* RB200_H */
