// rb_engine.cuh -- host-side driver of the wavefront ray engine.
#pragma once
#include <string>
#include <vector>
#include "rb_device.cuh"
#include "../../include/rb200.h"

namespace rb {
struct DirectJob;


struct EngineStats {
    unsigned long long nrays = 0, nodes = 0, leafents = 0, prims = 0, contribs = 0;
    unsigned long long launches = 0;     // kernels launched
    unsigned long long waves = 0, batches = 0, retries = 0;
    double kernel_ms = 0;                // device time of all kernels (CUDA events)
    double wave_ms = 0;                  // device time of k_trace launches only
    double shade_ms = 0;                 // device time of k_shade launches
    unsigned long long wave_launches = 0;
    unsigned long long badbin = 0;
};

// how -I sensors are turned into a pretend hit (see SURVEY 8a "three entries")
enum : int { IRR_NONE = 0, IRR_RTRACE = 1, IRR_RCONTRIB = 2, IRR_MANAGER = 3 };

struct TraceJob {
    const double* rays = nullptr;     // [nrays][6] origin, direction (host or device)
    bool rays_on_device = false;
    size_t nrays = 0;
    int accum = 1;                    // rays per output record (0: all into one)
    int irrad = IRR_NONE;
    bool lim_dist = false;
    // outputs (any may be null)
    void* cmat = nullptr;             // [nrows][ncols][3] coefficients, float32 (or float64)
    bool cmat_double = false;
    bool cmat_on_device = false;
    double* values = nullptr;         // [nrows][3] radiance/irradiance (host)
    RayResult* results = nullptr;     // [nrays] primary-hit reports (host)
    unsigned long long row_base = 0;  // global index of row 0 (RNG keys, multi-GPU shards)
};

class Engine {
public:
    explicit Engine(int device);
    ~Engine();
    bool upload_scene(const FlatScene& fs, const Scene& sc, std::string& err);
    bool set_bins(const std::vector<DBinSpec>& bins, const std::vector<int>& otrack, int ncols, std::string& err);
    bool run(const TraceJob& job, const DParams& P, std::string& err);
    void set_stream(cudaStream_t s) {          // launch on the caller's stream (the engine's own one stays for later)
        if (stream_ && stream_ != s) { cudaSetDevice(dev_); cudaStreamSynchronize(stream_); }
        stream_ = s; user_stream_ = true;
    }
    void set_queue_capacity(size_t nrays) { qcap_req_ = nrays; }
    // A closed context hands its engine to a per-device pool (rb_api.cu) so that the next context does not pay for
    // streams, events, pinned counters and small queues again: recycle() forgets everything that belonged to the
    // old context and says whether the engine is worth keeping (large queues / scenes are released instead).
    bool recycle();
    EngineStats stats;
    int device() const { return dev_; }
    const std::vector<std::string>* objnames = nullptr;
    std::string describe_obj(unsigned idx) const;
    std::vector<std::string> objdesc;      // "type \"name\"" per object, for messages

private:
    bool ensure_queues(std::string& err, size_t hint = 0);
    bool size_trace_grid(std::string& err);
    bool run_batch(const TraceJob& job, const DParams& P, size_t rec0, size_t nrec, std::string& err, bool& overflow);
    int dev_;
    cudaStream_t stream_ = nullptr;
    bool user_stream_ = false;
    DScene S_{};
    void *d_nodes_ = nullptr, *d_leaf_ = nullptr, *d_hdr_ = nullptr, *d_geom_ = nullptr,
         *d_mats_ = nullptr, *d_srcs_ = nullptr, *d_pats_ = nullptr, *d_bsdfs_ = nullptr, *d_bsdfbases_ = nullptr, *d_bsdfpool_ = nullptr, *d_otrack_ = nullptr, *d_bins_ = nullptr, *d_top_ = nullptr;
    size_t cap_nodes_ = 0, cap_leaf_ = 0, cap_hdr_ = 0, cap_geom_ = 0, cap_mats_ = 0, cap_srcs_ = 0, cap_pats_ = 0, cap_bsdfs_ = 0, cap_bsdfbases_ = 0, cap_bsdfpool_ = 0,
           cap_otrack_ = 0, cap_bins_ = 0, cap_top_ = 0;      // bytes behind the scene pointers (re-used by the next upload)
    cudaStream_t own_stream_ = nullptr;    // the stream this engine created (stream_ may be the caller's)
    int nbins_ = 0, ncols_ = 0;
    QRay* q_[2] = {nullptr, nullptr};
    QHemi* h_[2] = {nullptr, nullptr};
    DirectJob* dq_ = nullptr;   // parked direct() calculations (only when the scene has many sources)
    size_t dcap_ = 0;
    size_t qcap_ = 0, hcap_ = 0, qcap_req_ = 0;
    DCounters* d_cnt_ = nullptr;
    DCounters* h_cnt_ = nullptr;          // pinned
    double* d_acc_ = nullptr; size_t acc_bytes_ = 0;
    double* d_vacc_ = nullptr; size_t vacc_bytes_ = 0;
    double* d_rays_ = nullptr; size_t rays_bytes_ = 0;
    char* d_out_ = nullptr; size_t out_bytes_ = 0;
    RayResult* d_res_ = nullptr; size_t res_bytes_ = 0;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr, ev2_ = nullptr, ev3_ = nullptr;
    cudaEvent_t wev_[32] = {};            // per-wave events of a chunk of queued-ahead waves (4 per wave)
    HitRec* d_hits_ = nullptr;
    unsigned* d_slow_ = nullptr;          // queue slots left to the general shading kernel
    unsigned* d_mid_ = nullptr;           // queue slots left to k_shade_mid
    unsigned* d_lean_ = nullptr;          // queue slots left to k_shade_lean
    unsigned* d_spec_ = nullptr;          // queue slots left to k_shade_spec
    int trace_blocks_ = 148;
    size_t trace_smem_ = 0;      // dynamic shared memory of k_trace (ancestor stack)
    bool has_local_sources_ = false;
    bool park_direct() const { return has_local_sources_ || nsrc_active_ >= 64; }   // 64 = RB_COOP_SRC_MIN
    int nsrc_active_ = 0;       // distant sources direct() samples (not glow skies)
    std::string local_source_note_;
};

// rb_views.cu: view rays on the device (vwrays)
bool view_rays(int device, cudaStream_t stream, const struct ::rb_view& v, int xres, int yres, int repeat, double pj,
               unsigned long long seed, double* out, bool out_dev, std::string& err);

// rb_mtx.cu: C[nr][nc][3] = A[nr][ni][3] x B[ni][nc][3] per colour channel (dctimestep's cm_multiply)
bool mtx_multiply(int device, cudaStream_t stream, const float* A, size_t nr, size_t ni, const float* B, size_t nc,
                  float* C, bool a_dev, bool b_dev, bool c_dev, double* kernel_ms, std::string& err);

}  // namespace rb
