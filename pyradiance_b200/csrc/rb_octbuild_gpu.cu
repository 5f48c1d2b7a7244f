// rb_octbuild_gpu.cu -- the octree build on the device (reference ot/oconv.c:215-320).
//
// The reference inserts surfaces one by one and splits a cube when it overflows; the tree that results depends
// only on (surface set, cube, -n, -r): a cube is split iff more than `objlim` surfaces cross it and its children
// are not below the minimum size (rb_octbuild.cpp states the rule).  Here the tree is built top-down, LEVEL BY
// LEVEL, with the expensive part -- every (surface, child cube) overlap test, the reference's own tests of
// rb_octtests.hpp -- on the GPU:
//
//   frontier of level L = the cubes still to split, each with its candidate list (a segment of one array);
//   k_oct_count    one thread per (cube, candidate): tests the surface against the cube's eight children,
//                  stores the 8-bit result and counts the survivors per child (atomics);
//   host           reads the 8 counts per cube, decides empty / leaf / split per child (a few integer
//                  compares), lays the children's candidate segments out back to back (prefix sum);
//   k_oct_scatter  one thread per (cube, candidate): writes the surface into the segments of the children it
//                  crosses;
//   next level     the children that split.
// The host then assembles nodes and leaf sets depth-first -- the order of the serial builder, equal sibling
// leaves merged like common/octree.c:73-91 combine() -- so the result is the host builder's (and `oconv -f`'s)
// tree word for word; tests/test_gpu.py compares the two files byte by byte.
// Both builders run the same IEEE double arithmetic (no contraction), so they take the same decisions.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "rb_octtests.hpp"

namespace rb {
using octt::Prim;

#define OCK(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); goto fail; } \
    } while (0)

// surfaces that cross the root cube
__global__ void k_oct_root(const Prim* __restrict__ prims, int n, double ox, double oy, double oz, double size, double mincu,
                           unsigned char* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double org[3] = {ox, oy, oz};
    flag[i] = octt::overlaps(prims[i], org, size, mincu) ? 1 : 0;
}

__global__ void k_oct_count(const Prim* __restrict__ prims, const int* __restrict__ cand, const int* __restrict__ ecube, int n,
                            const double* __restrict__ corg, double size, double mincu, unsigned char* __restrict__ mask,
                            int* __restrict__ childcount) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int c = ecube[e];
    if (c < 0) { mask[e] = 0; return; }              // an entry of a finished (leaf) segment
    const Prim& p = prims[cand[e]];
    const double half = size * 0.5;
    const double o0 = corg[3 * c], o1 = corg[3 * c + 1], o2 = corg[3 * c + 2];
    unsigned m = 0;
    for (int k = 0; k < 8; k++) {
        const double ko[3] = {o0 + ((k & 1) ? half : 0.0), o1 + ((k & 2) ? half : 0.0), o2 + ((k & 4) ? half : 0.0)};
        if (octt::overlaps(p, ko, half, mincu)) { m |= 1u << k; atomicAdd(&childcount[8 * c + k], 1); }
    }
    mask[e] = (unsigned char)m;
}

__global__ void k_oct_scatter(const int* __restrict__ cand, const int* __restrict__ ecube, const unsigned char* __restrict__ mask,
                              int n, const int* __restrict__ childstart, const int* __restrict__ childcube, int* __restrict__ fill,
                              int* __restrict__ ncand, int* __restrict__ necube) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    unsigned m = mask[e];
    if (!m) return;
    const int c = ecube[e], pr = cand[e];
    for (; m; m &= m - 1) {
        const int k = __ffs(m) - 1;
        const int pos = childstart[8 * c + k] + atomicAdd(&fill[8 * c + k], 1);
        ncand[pos] = pr;
        necube[pos] = childcube[8 * c + k];
    }
}

bool octbuild_device_available() {
    if (getenv("RB_OCTBUILD_HOST")) return false;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return false; }
    return n > 0;
}

namespace {
struct Level {
    std::vector<int> child;          // 8 words per frontier cube: -1 empty, >= 0 frontier index of the next level, <= -2 leaf -(id) - 2
    std::vector<int> leaf_start, leaf_count;     // per leaf id: its segment in `cand` below
    std::vector<int> cand;           // the NEXT level's candidate array (surface indices), host copy: leaf sets are cut from it
};

struct Assembler {
    const std::vector<Prim>& prims;
    std::vector<Level>& lv;
    std::vector<int>&nodes, &pool;
    std::vector<int> tmp;
    std::string err;
    int leaf(const int* seg, int n) {
        if (n > octt::MAXSET) { err = "set overflow in octree build"; return -1; }
        tmp.resize(n);
        for (int i = 0; i < n; i++) tmp[i] = prims[seg[i]].obj;
        std::sort(tmp.begin(), tmp.end());
        const int off = (int)pool.size();
        pool.push_back(n);
        pool.insert(pool.end(), tmp.begin(), tmp.end());
        return -off - 2;
    }
    bool same_leaf(int a, int b) const {
        if (a == b) return true;
        const int *pa = &pool[-a - 2], *pb = &pool[-b - 2];
        return pa[0] == pb[0] && std::equal(pa + 1, pa + 1 + pa[0], pb + 1);
    }
    // tree word of frontier cube `c` of level `L`, children first (the serial builder's order)
    int cube(int L, int c) {
        int kids[8];
        const Level& l = lv[L];
        for (int k = 0; k < 8; k++) {
            const int w = l.child[8 * c + k];
            if (w == -1) kids[k] = -1;
            else if (w >= 0) kids[k] = cube(L + 1, w);
            else { const int id = -w - 2; kids[k] = leaf(&l.cand[l.leaf_start[id]], l.leaf_count[id]); }
            if (!err.empty()) return -1;
        }
        bool same = kids[0] < 0;                  // combine(): eight equal leaves (or eight empties) become one
        for (int i = 1; i < 8 && same; i++) same = kids[0] == -1 ? kids[i] == -1 : (kids[i] < -1 && same_leaf(kids[i], kids[0]));
        if (same) return kids[0];
        const int idx = (int)(nodes.size() / 8);
        nodes.insert(nodes.end(), kids, kids + 8);
        return idx;
    }
};
}  // namespace

// The tree of Builder::build() (rb_octbuild.cpp) over `prims` (polygon vertices still in host memory behind
// Prim::va), computed level by level on the current CUDA device.
bool build_tree_device(const std::vector<Prim>& prims, int objlim, double mincusize, const double cuorg[3], double cusize,
                       std::vector<int>& nodes, std::vector<int>& pool, int& root, std::string& err) {
    const int np = (int)prims.size();
    nodes.clear(); pool.clear(); root = -1;
    const bool dbg = getenv("RB_DEBUG_OCT") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now(), t_up = 0, t_count = 0, t_host = 0, t_scatter = 0, t_down = 0, t_asm = 0;
    if (np == 0) return true;
    Prim* d_prims = nullptr; double* d_verts = nullptr; unsigned char* d_mask = nullptr;
    int *d_cand = nullptr, *d_ecube = nullptr, *d_ncand = nullptr, *d_necube = nullptr, *d_cnt = nullptr, *d_start = nullptr,
        *d_ccube = nullptr, *d_fill = nullptr;
    double* d_corg = nullptr;
    size_t cap_e = 0, cap_ne = 0, cap_c = 0;
    std::vector<Level> lv;
    {
    // ---- surfaces and their vertices to the device ----
    std::vector<double> verts;
    std::vector<Prim> hp(prims);
    std::vector<size_t> voff(np, 0);
    for (int i = 0; i < np; i++)
        if (hp[i].kind == 0) { voff[i] = verts.size(); verts.insert(verts.end(), hp[i].va, hp[i].va + 3 * (size_t)hp[i].nv); }
    OCK(cudaMalloc(&d_verts, std::max<size_t>(verts.size(), 1) * sizeof(double)));
    OCK(cudaMemcpy(d_verts, verts.data(), verts.size() * sizeof(double), cudaMemcpyHostToDevice));
    for (int i = 0; i < np; i++) if (hp[i].kind == 0) hp[i].va = d_verts + voff[i];
    OCK(cudaMalloc(&d_prims, (size_t)np * sizeof(Prim)));
    OCK(cudaMemcpy(d_prims, hp.data(), (size_t)np * sizeof(Prim), cudaMemcpyHostToDevice));
    t_up = now() - t0;
    // ---- the root cube ----
    std::vector<int> in;
    {
        OCK(cudaMalloc(&d_mask, (size_t)np));
        k_oct_root<<<(np + 255) / 256, 256>>>(d_prims, np, cuorg[0], cuorg[1], cuorg[2], cusize, mincusize, d_mask);
        std::vector<unsigned char> fl(np);
        OCK(cudaMemcpy(fl.data(), d_mask, (size_t)np, cudaMemcpyDeviceToHost));
        for (int i = 0; i < np; i++) if (fl[i]) in.push_back(i);
        cudaFree(d_mask); d_mask = nullptr;
    }
    if (in.empty()) { cudaFree(d_prims); cudaFree(d_verts); return true; }
    auto is_leaf = [&](size_t n, double size, int depth) {          // Builder::build()'s rule for a cube of this size
        const double half = size * 0.5;
        const bool toosmall = half < ((int)n < octt::MAXSET ? mincusize : mincusize / 256.0);
        return (int)n <= objlim || toosmall || depth >= 20;
    };
    if (is_leaf(in.size(), cusize, 0)) {
        std::vector<Level> one;
        Assembler A{prims, one, nodes, pool, {}, {}};
        root = A.leaf(in.data(), (int)in.size());
        cudaFree(d_prims); cudaFree(d_verts);
        if (!A.err.empty()) { err = A.err; return false; }
        return true;
    }
    // ---- level loop ----
    std::vector<double> corg = {cuorg[0], cuorg[1], cuorg[2]};      // origins of the frontier cubes
    size_t ne = in.size();
    cap_e = ne;
    OCK(cudaMalloc(&d_cand, cap_e * sizeof(int)));
    OCK(cudaMalloc(&d_ecube, cap_e * sizeof(int)));
    OCK(cudaMemcpy(d_cand, in.data(), ne * sizeof(int), cudaMemcpyHostToDevice));
    OCK(cudaMemset(d_ecube, 0, ne * sizeof(int)));
    double size = cusize;
    std::vector<int> cnt, start, ccube;
    for (int depth = 0; !corg.empty(); depth++) {
        const size_t nc = corg.size() / 3;
        if (nc * 8 > cap_c) {
            for (void* p : {(void*)d_cnt, (void*)d_start, (void*)d_ccube, (void*)d_fill, (void*)d_corg}) if (p) cudaFree(p);
            d_cnt = d_start = d_ccube = d_fill = nullptr; d_corg = nullptr;
            cap_c = nc * 8 * 2;
            OCK(cudaMalloc(&d_cnt, cap_c * sizeof(int))); OCK(cudaMalloc(&d_start, cap_c * sizeof(int)));
            OCK(cudaMalloc(&d_ccube, cap_c * sizeof(int))); OCK(cudaMalloc(&d_fill, cap_c * sizeof(int)));
            OCK(cudaMalloc(&d_corg, cap_c / 8 * 3 * sizeof(double)));
        }
        if (d_mask) cudaFree(d_mask);
        d_mask = nullptr;
        OCK(cudaMalloc(&d_mask, ne));
        OCK(cudaMemcpy(d_corg, corg.data(), nc * 3 * sizeof(double), cudaMemcpyHostToDevice));
        OCK(cudaMemset(d_cnt, 0, nc * 8 * sizeof(int)));
        double ta = now();
        k_oct_count<<<(unsigned)((ne + 127) / 128), 128>>>(d_prims, d_cand, d_ecube, (int)ne, d_corg, size, mincusize, d_mask, d_cnt);
        OCK(cudaGetLastError());
        cnt.resize(nc * 8);
        OCK(cudaMemcpy(cnt.data(), d_cnt, nc * 8 * sizeof(int), cudaMemcpyDeviceToHost));
        t_count += now() - ta; ta = now();
        // ---- decide the children, lay out their segments ----
        lv.emplace_back();
        Level& l = lv.back();
        l.child.assign(nc * 8, -1);
        start.assign(nc * 8, 0); ccube.assign(nc * 8, -1);
        const double half = size * 0.5;
        std::vector<double> norg;
        size_t nne = 0;
        for (size_t c = 0; c < nc; c++)
            for (int k = 0; k < 8; k++) {
                const int n = cnt[8 * c + k];
                start[8 * c + k] = (int)nne;
                if (n == 0) continue;
                if (is_leaf((size_t)n, half, depth + 1)) {
                    l.child[8 * c + k] = -(int)l.leaf_start.size() - 2;
                    l.leaf_start.push_back((int)nne); l.leaf_count.push_back(n);
                } else {
                    const int idx = (int)(norg.size() / 3);
                    l.child[8 * c + k] = idx; ccube[8 * c + k] = idx;
                    for (int j = 0; j < 3; j++) norg.push_back(corg[3 * c + j] + (((1 << j) & k) ? half : 0.0));
                }
                nne += (size_t)n;
                if (nne > (size_t)0x7fff0000) { err = "octree build: candidate lists exceed 2^31 entries"; goto fail; }
            }
        t_host += now() - ta; ta = now();
        // ---- scatter ----
        if (nne > cap_ne) {
            if (d_ncand) cudaFree(d_ncand);
            if (d_necube) cudaFree(d_necube);
            d_ncand = d_necube = nullptr;
            cap_ne = nne + nne / 4;
            OCK(cudaMalloc(&d_ncand, cap_ne * sizeof(int))); OCK(cudaMalloc(&d_necube, cap_ne * sizeof(int)));
        }
        OCK(cudaMemcpy(d_start, start.data(), nc * 8 * sizeof(int), cudaMemcpyHostToDevice));
        OCK(cudaMemcpy(d_ccube, ccube.data(), nc * 8 * sizeof(int), cudaMemcpyHostToDevice));
        OCK(cudaMemset(d_fill, 0, nc * 8 * sizeof(int)));
        k_oct_scatter<<<(unsigned)((ne + 255) / 256), 256>>>(d_cand, d_ecube, d_mask, (int)ne, d_start, d_ccube, d_fill, d_ncand, d_necube);
        OCK(cudaGetLastError());
        OCK(cudaDeviceSynchronize());
        t_scatter += now() - ta; ta = now();
        l.cand.resize(nne);
        if (nne) OCK(cudaMemcpy(l.cand.data(), d_ncand, nne * sizeof(int), cudaMemcpyDeviceToHost));
        t_down += now() - ta;
        if (dbg) fprintf(stderr, "[rb oct] level %d: %zu cubes, %zu candidates -> %zu, %zu split\n", depth, nc, ne, nne, norg.size() / 3);
        // ---- next level ----
        std::swap(d_cand, d_ncand); std::swap(d_ecube, d_necube); std::swap(cap_e, cap_ne);
        ne = nne; corg.swap(norg); size = half;
    }
    }
    {
        double ta = now();
        Assembler A{prims, lv, nodes, pool, {}, {}};
        root = A.cube(0, 0);
        if (!A.err.empty()) { err = A.err; goto fail; }
        t_asm = now() - ta;
        if (dbg) fprintf(stderr, "[rb oct] device build of %d surfaces: upload %.3f, count kernels %.3f, host decisions %.3f, scatter %.3f, "
                                 "download %.3f, assembly %.3f, total %.3f s\n", np, t_up, t_count, t_host, t_scatter, t_down, t_asm, now() - t0);
    }
    for (void* p : {(void*)d_prims, (void*)d_verts, (void*)d_mask, (void*)d_cand, (void*)d_ecube, (void*)d_ncand, (void*)d_necube,
                    (void*)d_cnt, (void*)d_start, (void*)d_ccube, (void*)d_fill, (void*)d_corg})
        if (p) cudaFree(p);
    return true;
fail:
    for (void* p : {(void*)d_prims, (void*)d_verts, (void*)d_mask, (void*)d_cand, (void*)d_ecube, (void*)d_ncand, (void*)d_necube,
                    (void*)d_cnt, (void*)d_start, (void*)d_ccube, (void*)d_fill, (void*)d_corg})
        if (p) cudaFree(p);
    cudaGetLastError();
    return false;
}

}  // namespace rb
