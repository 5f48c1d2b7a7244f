// rb_views.cu -- vwrays on the device: one thread per view ray.
// common/image.c:214-305 viewray() for the six view types, util/vwrays.c:245-300 putrays() for the pixel order
// (scanlines from the top, `repeat` rays per pixel, optional jitter), expression by expression as
// pyradiance_b200/views.py states them on the host (the Python mirror stays the CPU-only path of vwrays()).
#include <cuda_runtime.h>
#include <string>
#include "rb_engine.cuh"

namespace rb {

__device__ __forceinline__ unsigned long long vw_mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double vw_rnd(unsigned long long key, unsigned dim) {
    return (double)(vw_mix64(key + 0x9e3779b97f4a7c15ULL * (dim + 1)) >> 11) * (1.0 / 9007199254740992.0);
}
// common/fvect.c:130-157 normalize(), the near-unit shortcut included
__device__ __forceinline__ double vw_normalize(double v[3]) {
    double d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], len;
    if (d == 0.0) return 0.0;
    if ((d <= 1.0 + 1e-6) & (d >= 1.0 - 1e-6)) { len = 0.5 + 0.5 * d; d = 2.0 - len; }
    else { len = sqrt(d); d = 1.0 / len; }
    v[0] *= d; v[1] *= d; v[2] *= d;
    return len;
}

__global__ void k_view_rays(const rb_view V, int xres, int yres, int repeat, double pj, unsigned long long seed,
                            double* __restrict__ out, size_t n) {
    const double PI = 3.14159265358979323846, FTINY = 1e-6;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i / (size_t)repeat;
        const int sx = (int)(pix % (size_t)xres), sy = (int)(pix / (size_t)xres);
        double lx = (sx + .5) / xres, ly = ((yres - 1 - sy) + .5) / yres;
        if (pj > FTINY) {
            const unsigned long long key = vw_mix64(seed ^ vw_mix64(i));
            lx = lx + pj * (.5 - vw_rnd(key, 0)) / xres;
            ly = ly + pj * (.5 - vw_rnd(key, 1)) / yres;
        }
        double x = lx + (V.hoff - 0.5), y = ly + (V.voff - 0.5);
        const double aft = V.vaft > FTINY ? V.vaft - V.vfore : 0.0;
        double org[3], dir[3], d = 0.0, z;
        switch (V.type) {
        case 'l':
            for (int k = 0; k < 3; k++) { org[k] = ((V.vp[k] + V.vfore * V.vdir[k]) + x * V.hvec[k]) + y * V.vvec[k]; dir[k] = V.vdir[k]; }
            d = aft;
            break;
        case 'v':
            for (int k = 0; k < 3; k++) { dir[k] = (V.vdir[k] + x * V.hvec[k]) + y * V.vvec[k]; org[k] = V.vp[k] + dir[k] * V.vfore; }
            d = vw_normalize(dir);
            d = V.vaft > FTINY ? aft * d : 0.0;
            break;
        case 'h':
            z = 1.0 - x * x * V.hn2 - y * y * V.vn2;
            if (z < 0.0) { d = -1.0; z = 0.0; } else d = aft;
            z = sqrt(z);
            for (int k = 0; k < 3; k++) { dir[k] = (z * V.vdir[k] + x * V.hvec[k]) + y * V.vvec[k]; org[k] = V.vp[k] + dir[k] * V.vfore; }
            break;
        case 'c': {
            const double a = x * V.horiz * (PI / 180.0);
            const double ca = cos(a), sa = sin(a);
            for (int k = 0; k < 3; k++) { dir[k] = (ca * V.vdir[k] + sa * V.hvec[k]) + y * V.vvec[k]; org[k] = V.vp[k] + dir[k] * V.vfore; }
            d = vw_normalize(dir);
            d = V.vaft > FTINY ? aft * d : 0.0;
            break;
        }
        case 'a': {
            x = x * ((1.0 / 180.0) * V.horiz);
            y = y * ((1.0 / 180.0) * V.vert);
            double r = x * x + y * y;
            const bool ok = r <= 1.0;
            r = sqrt(r);
            z = cos(PI * r);
            const double s = r <= FTINY ? PI : sqrt(1.0 - z * z) / r;
            for (int k = 0; k < 3; k++) { dir[k] = (z * V.vdir[k] + (x * s) * V.hvec[k]) + (y * s) * V.vvec[k]; org[k] = V.vp[k] + dir[k] * V.vfore; }
            d = ok ? aft : -1.0;
            break;
        }
        default: {      // 's'
            x = x * sqrt(V.hn2);
            y = y * sqrt(V.vn2);
            const double r = x * x + y * y;
            z = (1. - r) / (1. + r);
            for (int k = 0; k < 3; k++) { dir[k] = (z * V.vdir[k] + (x * (1. + z)) * V.hvec[k]) + (y * (1. + z)) * V.vvec[k]; org[k] = V.vp[k] + dir[k] * V.vfore; }
            d = aft;
        }
        }
        double* o = out + i * 6;
        if (d < -FTINY) { for (int k = 0; k < 6; k++) o[k] = 0.0; continue; }
        const double sc = d > FTINY ? d : 1.0;
        for (int k = 0; k < 3; k++) { o[k] = org[k]; o[3 + k] = dir[k] * sc; }
    }
}

bool view_rays(int device, cudaStream_t stream, const rb_view& v, int xres, int yres, int repeat, double pj,
               unsigned long long seed, double* out, bool out_dev, std::string& err) {
    const size_t n = (size_t)xres * yres * repeat;
    double* d = out;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess && !out_dev) e = cudaMalloc(&d, n * 6 * sizeof(double));
    if (e != cudaSuccess) { err = std::string("rb_view_rays: ") + cudaGetErrorString(e); return false; }
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 32);
    k_view_rays<<<grid, 256, 0, stream>>>(v, xres, yres, repeat, pj, seed, d, n);
    e = cudaGetLastError();
    if (e == cudaSuccess && !out_dev) e = cudaMemcpyAsync(out, d, n * 6 * sizeof(double), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (!out_dev) cudaFree(d);
    if (e != cudaSuccess) { err = std::string("rb_view_rays: ") + cudaGetErrorString(e); return false; }
    return true;
}

}  // namespace rb
