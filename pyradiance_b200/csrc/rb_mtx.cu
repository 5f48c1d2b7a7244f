// rb_mtx.cu -- the matrix consumer right after the hot path:
// dctimestep's  result[s][t] = sum_b DC[s][b] * sky[b][t]  per colour channel
// (util/cmatrix.c:420-475 cm_multiply: float in, double accumulation, float out).
//
// RGB triplets are interleaved, so this is three GEMMs that share their index
// structure; a tile of the A rows / B columns is contiguous in memory for all three
// channels.  Hand-written fp32 SIMT kernel: one channel of a 128 x 128 output tile per
// CTA, 16-deep double-buffered k slices in shared memory, 8 x 8 accumulators per thread.
// fp32 on purpose: coefficients span many decades (TF32 / BF16 tensor-core inputs
// would cost 3 digits), and to stay inside 1e-5 of the reference's double
// accumulation the sums are two-level (128 products in one accumulator, then
// folded into a second): worst-case relative error (128 + K/128) eps/2 for the
// non-negative data of this domain.  The step is ~2 % of the matrix computation.
#include <cuda_runtime.h>
#include <algorithm>
#include <string>
#include "rb_engine.cuh"

namespace rb {

constexpr int MT_BM = 128, MT_BN = 128, MT_BK = 16, MT_THREADS = 256, MT_FOLD = 8;   // fold every 8 slices = 128 products
constexpr int MT_LD = MT_BM + 4;                 // padded row of a staged slice (the transposing stores of the A tile)

// One colour channel of a 128 x 128 output tile per CTA (blockIdx.x = channel, so the three CTAs that share
// the sectors of an interleaved output tile run side by side and their writes merge in L2).  256 threads,
// 8 x 8 accumulators per thread as 2 x 2 groups of 4 x 4 (rows ty*4.. and 64+ty*4.., columns tx*4.. and
// 64+tx*4..: every shared-memory read is a conflict-free LDS.128, 64 bytes per 64 FMAs -- the 4 x 4 x 3 tile
// of the first version read 96 bytes per 48 and was bound by shared-memory bandwidth at 23 % of the FMA peak).
// k slices of 16 are double-buffered: the next slice travels global -> registers while the current one is
// multiplied, then registers -> the other buffer.  The second accumulation level lives in shared memory
// (64 floats per thread, touched once per 128 products).
#ifndef MT_MINB
#define MT_MINB 2
#endif
__global__ void __launch_bounds__(MT_THREADS, MT_MINB) k_mtx3(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ C, int nr, int ni, int nc) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                               // [2][MT_BK][MT_LD]
    float* Bs = smem + 2 * MT_BK * MT_LD;           // [2][MT_BK][MT_LD]
    float* acc2 = smem + 4 * MT_BK * MT_LD;         // [64][MT_THREADS]
    // (a warp as 2 row groups x 16 column groups; 8 x 4 saves shared-memory wavefronts but measured no faster)
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ch = blockIdx.x;
    const size_t r0 = (size_t)blockIdx.z * MT_BM, c0 = (size_t)blockIdx.y * MT_BN;
    // accumulators as float2 pairs along the columns: the products go through the packed FFMA2 of sm_100
    // (two FMAs per issue slot; with scalar FFMA the kernel sat at 62 % of its issue slots with the FMA pipe at 47 %)
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);
    bool folded = false;                            // the second level is written by the first fold, not zeroed up front
    // staging assignment: A element e -> row e / 16, k e % 16 (16 consecutive threads walk one row's k run);
    // B element e -> k e / 128, column e % 128 (consecutive threads walk consecutive columns)
    float ra[8], rb[8];
    // the 8 + 8 elements a thread stages per slice: element t of A is row (tid >> 4) + 16 t at k = tid & 15,
    // element t of B is k = (tid >> 7) + 2 t at column tid & 127 -- two base pointers that advance by one
    // slice per fetch, row validity as a mask, column validity as one flag, only the k bound changes
    const int ka = tid & 15, kb0 = tid >> 7;
    const float* pa = A + ((r0 + (tid >> 4)) * ni + ka) * 3 + ch;
    const float* pb = B + ((size_t)kb0 * nc + c0 + (tid & 127)) * 3 + ch;
    const size_t astep = (size_t)16 * ni * 3, bstep = (size_t)2 * nc * 3;
    unsigned va = 0;
#pragma unroll
    for (int t = 0; t < 8; t++)
        if (r0 + (tid >> 4) + 16 * t < (size_t)nr) va |= 1u << t;
    const bool vb = c0 + (tid & 127) < (size_t)nc;
    auto fetch = [&](int k0) {
        const bool kin = k0 + ka < ni;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            ra[t] = (kin && (va >> t & 1)) ? __ldg(pa + t * astep) : 0.f;
            rb[t] = (vb && k0 + kb0 + 2 * t < ni) ? __ldg(pb + t * bstep) : 0.f;
        }
        pa += MT_BK * 3; pb += (size_t)MT_BK * nc * 3;
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int e = tid + t * MT_THREADS;
            As[(buf * MT_BK + (e & 15)) * MT_LD + (e >> 4)] = ra[t];
            Bs[(buf * MT_BK + (e >> 7)) * MT_LD + (e & 127)] = rb[t];
        }
    };
    fetch(0);
    stage(0);
    __syncthreads();
    int slice = 0;
    for (int k0 = 0; k0 < ni; k0 += MT_BK, slice++) {
        const int buf = slice & 1;
        const bool more = k0 + MT_BK < ni;
        if (more) fetch(k0 + MT_BK);
        const float* as = As + buf * MT_BK * MT_LD + ty * 4;
        const float* bs = Bs + buf * MT_BK * MT_LD + tx * 4;
#pragma unroll
        for (int kk = 0; kk < MT_BK; kk++) {
            const float4 a0 = *reinterpret_cast<const float4*>(as + kk * MT_LD);
            const float4 a1 = *reinterpret_cast<const float4*>(as + kk * MT_LD + 64);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * MT_LD);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + kk * MT_LD + 64);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = __ffma2_rn(aa, b[j], acc[i][j]);
            }
        }
        if ((slice & (MT_FOLD - 1)) == MT_FOLD - 1) {
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float* p0 = &acc2[(i * 8 + 2 * j) * MT_THREADS + tid];
                    float* p1 = &acc2[(i * 8 + 2 * j + 1) * MT_THREADS + tid];
                    *p0 = folded ? *p0 + acc[i][j].x : acc[i][j].x;
                    *p1 = folded ? *p1 + acc[i][j].y : acc[i][j].y;
                    acc[i][j] = make_float2(0.f, 0.f);
                }
            folded = true;
        }
        if (more) stage(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const size_t r = r0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= (size_t)nr) continue;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const size_t c = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c >= (size_t)nc) continue;
            const float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
            C[(r * nc + c) * 3 + ch] = folded ? acc2[(i * 8 + j) * MT_THREADS + tid] + v : v;
        }
    }
}
constexpr size_t MT_SMEM = (size_t)(4 * MT_BK * MT_LD + 64 * MT_THREADS) * sizeof(float);   // 33 KB of slices + 64 KB second level

#define MCK(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); goto done; } \
    } while (0)

// rb_mtx_tc.cu: the same product on the tensor cores (3xTF32, tcgen05 / TMEM / TMA)
bool mtx_multiply_tc(cudaStream_t stream, const float* A, size_t n, size_t ni, const float* Bplanes, size_t Np, int Kp,
                     size_t nc, float* C, float* Aplanes, size_t Mp_cap, double* kernel_ms, std::string& err);
bool tc_prepare_b(cudaStream_t stream, const float* B, size_t ni, size_t nc, float** planes, size_t* Np_out, int* Kp_out,
                  std::string& err);

// C[nr][nc][3] = A[nr][ni][3] x B[ni][nc][3], channel by channel.  Host or device buffers.
// Products large enough to fill tensor-core tiles go to k_mtx_tc (rb_mtx_tc.cu); small ones (rmtxop on a few rows,
// inner dimension below 8) stay on the SIMT kernel above, where the operand pre-pass would cost more than it saves.
bool mtx_multiply(int device, cudaStream_t stream, const float* A, size_t nr, size_t ni, const float* B, size_t nc,
                  float* C, bool a_dev, bool b_dev, bool c_dev, double* kernel_ms, std::string& err) {
    float *dA = nullptr, *dB = nullptr, *dC = nullptr, *Bpl = nullptr, *Apl = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = false;
    double ms_total = 0;
    if (nr == 0 || nc == 0) return true;
    if (ni == 0) { err = "matrix dimension mismatch"; return false; }
    {
    MCK(cudaSetDevice(device));
    MCK(cudaEventCreate(&e0)); MCK(cudaEventCreate(&e1));
    MCK(cudaFuncSetAttribute(k_mtx3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
    const float* Bd = B;
    if (!b_dev) {
        MCK(cudaMalloc(&dB, ni * nc * 3 * sizeof(float)));
        MCK(cudaMemcpyAsync(dB, B, ni * nc * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
        Bd = dB;
    }
    // rows per chunk: staging buffers of at most ~1 GiB each when a side lives on the host
    size_t chunk = std::min<size_t>(nr, (size_t)65535 * MT_BM);        // grid.z limit
    if (!a_dev || !c_dev) {
        size_t per_row = std::max(ni, nc) * 3 * sizeof(float);
        chunk = std::max<size_t>(MT_BM, std::min<size_t>(chunk, ((size_t)1 << 30) / per_row / MT_BM * MT_BM));
    }
    const bool tc = !getenv("RB_MTX_SIMT") && ni >= 8 && nr >= 64 && nc >= 64;
    size_t Np = 0, Mp_cap = 0; int Kp = 0;
    if (tc) {
        if (!tc_prepare_b(stream, Bd, ni, nc, &Bpl, &Np, &Kp, err)) goto done;
        // operand planes of a chunk of rows: at most ~2 GiB
        chunk = std::max<size_t>(128, std::min(chunk, ((size_t)2 << 30) / ((size_t)24 * Kp) / 128 * 128));
        Mp_cap = (std::min(chunk, nr) + 127) / 128 * 128;
        MCK(cudaMalloc(&Apl, 6 * Mp_cap * (size_t)Kp * sizeof(float)));
    }
    if (!a_dev) MCK(cudaMalloc(&dA, chunk * ni * 3 * sizeof(float)));
    if (!c_dev) MCK(cudaMalloc(&dC, chunk * nc * 3 * sizeof(float)));
    for (size_t r = 0; r < nr; r += chunk) {
        const size_t n = std::min(chunk, nr - r);
        const float* Ad = a_dev ? A + r * ni * 3 : dA;
        float* Cd = c_dev ? C + r * nc * 3 : dC;
        if (!a_dev) MCK(cudaMemcpyAsync(dA, A + r * ni * 3, n * ni * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
        if (tc) {
            if (!mtx_multiply_tc(stream, Ad, n, ni, Bpl, Np, Kp, nc, Cd, Apl, Mp_cap, &ms_total, err)) goto done;
            if (!c_dev) MCK(cudaMemcpyAsync(C + r * nc * 3, dC, n * nc * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
            MCK(cudaStreamSynchronize(stream));
            continue;
        }
        dim3 grid(3, (unsigned)((nc + MT_BN - 1) / MT_BN), (unsigned)((n + MT_BM - 1) / MT_BM));
        MCK(cudaEventRecord(e0, stream));
        k_mtx3<<<grid, MT_THREADS, MT_SMEM, stream>>>(Ad, Bd, Cd, (int)n, (int)ni, (int)nc);
        MCK(cudaEventRecord(e1, stream));
        MCK(cudaGetLastError());
        if (!c_dev) MCK(cudaMemcpyAsync(C + r * nc * 3, dC, n * nc * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
        MCK(cudaStreamSynchronize(stream));
        float ms = 0; MCK(cudaEventElapsedTime(&ms, e0, e1)); ms_total += ms;
    }
    ok = true;
    }
done:
    if (dA) cudaFree(dA);
    if (dB) cudaFree(dB);
    if (dC) cudaFree(dC);
    if (Bpl) cudaFree(Bpl);
    if (Apl) cudaFree(Apl);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms_total;
    return ok;
}

}  // namespace rb
