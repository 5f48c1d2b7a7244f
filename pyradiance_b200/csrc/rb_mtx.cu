// rb_mtx.cu -- the matrix consumer right after the hot path:
// dctimestep's  result[s][t] = sum_b DC[s][b] * sky[b][t]  per colour channel
// (util/cmatrix.c:420-475 cm_multiply: float in, double accumulation, float out).
//
// RGB triplets are interleaved, so this is three GEMMs that share their index
// structure; a tile of the A rows / B columns is contiguous in memory for all three
// channels.  Hand-written fp32 SIMT kernel: 64 x 64 output tile per CTA, 16-deep
// k slices staged through shared memory, 4 x 4 x 3 accumulators per thread.
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

constexpr int MT_BM = 64, MT_BN = 64, MT_BK = 16, MT_THREADS = 256, MT_FOLD = 8;   // fold every 8 slices = 128 products

// 64 x 64 output tile per CTA, 4 x 4 x 3 first-level accumulators per thread in registers; the second level
// lives in shared memory (48 floats per thread, touched once per 128 products).
#ifndef MT_MINB
#define MT_MINB 2
#endif
__global__ void __launch_bounds__(MT_THREADS, MT_MINB) k_mtx3(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ C, int nr, int ni, int nc) {
    extern __shared__ __align__(16) float smem[];
    float (*As)[MT_BM * 3] = reinterpret_cast<float (*)[MT_BM * 3]>(smem);                       // [MT_BK][MT_BM*3]
    float (*Bs)[MT_BN * 3] = reinterpret_cast<float (*)[MT_BN * 3]>(smem + MT_BK * MT_BM * 3);   // [MT_BK][MT_BN*3]
    float* acc2 = smem + MT_BK * (MT_BM + MT_BN) * 3;                                            // [48][MT_THREADS]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int r0 = blockIdx.y * MT_BM, c0 = blockIdx.x * MT_BN;
    float acc[4][4][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) { acc[i][j][k] = 0.f; acc2[((i * 4 + j) * 3 + k) * MT_THREADS + tid] = 0.f; }
    int slice = 0;
    for (int k0 = 0; k0 < ni; k0 += MT_BK, slice++) {
        // A tile: MT_BM rows x (MT_BK * 3) contiguous floats each
#pragma unroll
        for (int t = 0; t < (MT_BM * MT_BK * 3) / MT_THREADS; t++) {
            const int e = tid + t * MT_THREADS, row = e / (MT_BK * 3), rem = e % (MT_BK * 3);
            const int k = rem / 3, ch = rem % 3;
            float v = 0.f;
            if (r0 + row < nr && k0 + k < ni) v = __ldg(&A[((size_t)(r0 + row) * ni + k0) * 3 + rem]);
            As[k][row * 3 + ch] = v;
        }
        // B tile: MT_BK rows x (MT_BN * 3) contiguous floats each
#pragma unroll
        for (int t = 0; t < (MT_BK * MT_BN * 3) / MT_THREADS; t++) {
            const int e = tid + t * MT_THREADS, k = e / (MT_BN * 3), rem = e % (MT_BN * 3);
            float v = 0.f;
            if (k0 + k < ni && c0 + rem / 3 < nc) v = __ldg(&B[((size_t)(k0 + k) * nc + c0) * 3 + rem]);
            Bs[k][rem] = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int kk = 0; kk < MT_BK; kk++) {
            float a[4][3], b[4][3];
            const float4* ap = reinterpret_cast<const float4*>(&As[kk][ty * 12]);
            const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2];
            a[0][0] = a0.x; a[0][1] = a0.y; a[0][2] = a0.z; a[1][0] = a0.w; a[1][1] = a1.x; a[1][2] = a1.y;
            a[2][0] = a1.z; a[2][1] = a1.w; a[2][2] = a2.x; a[3][0] = a2.y; a[3][1] = a2.z; a[3][2] = a2.w;
            const float4* bp = reinterpret_cast<const float4*>(&Bs[kk][tx * 12]);
            const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2];
            b[0][0] = b0.x; b[0][1] = b0.y; b[0][2] = b0.z; b[1][0] = b0.w; b[1][1] = b1.x; b[1][2] = b1.y;
            b[2][0] = b1.z; b[2][1] = b1.w; b[2][2] = b2.x; b[3][0] = b2.y; b[3][1] = b2.z; b[3][2] = b2.w;
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int k = 0; k < 3; k++) acc[i][j][k] = fmaf(a[i][k], b[j][k], acc[i][j][k]);
        }
        __syncthreads();
        if ((slice & (MT_FOLD - 1)) == MT_FOLD - 1) {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        acc2[((i * 4 + j) * 3 + k) * MT_THREADS + tid] += acc[i][j][k];
                        acc[i][j][k] = 0.f;
                    }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = r0 + ty * 4 + i;
        if (r >= nr) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = c0 + tx * 4 + j;
            if (c >= nc) continue;
            float* o = C + ((size_t)r * nc + c) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) o[k] = acc2[((i * 4 + j) * 3 + k) * MT_THREADS + tid] + acc[i][j][k];
        }
    }
}
constexpr size_t MT_SMEM = (size_t)(MT_BK * (MT_BM + MT_BN) * 3 + 48 * MT_THREADS) * sizeof(float);   // 72 KB

#define MCK(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); goto done; } \
    } while (0)

// C[nr][nc][3] = A[nr][ni][3] x B[ni][nc][3], channel by channel.  Host or device buffers.
bool mtx_multiply(int device, cudaStream_t stream, const float* A, size_t nr, size_t ni, const float* B, size_t nc,
                  float* C, bool a_dev, bool b_dev, bool c_dev, double* kernel_ms, std::string& err) {
    float *dA = nullptr, *dB = nullptr, *dC = nullptr;
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
    size_t chunk = nr;
    if (!a_dev || !c_dev) {
        size_t per_row = std::max(ni, nc) * 3 * sizeof(float);
        chunk = std::max<size_t>(MT_BM, std::min<size_t>(nr, ((size_t)1 << 30) / per_row / MT_BM * MT_BM));
    }
    if (!a_dev) MCK(cudaMalloc(&dA, chunk * ni * 3 * sizeof(float)));
    if (!c_dev) MCK(cudaMalloc(&dC, chunk * nc * 3 * sizeof(float)));
    for (size_t r = 0; r < nr; r += chunk) {
        const size_t n = std::min(chunk, nr - r);
        const float* Ad = a_dev ? A + r * ni * 3 : dA;
        float* Cd = c_dev ? C + r * nc * 3 : dC;
        if (!a_dev) MCK(cudaMemcpyAsync(dA, A + r * ni * 3, n * ni * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
        dim3 grid((unsigned)((nc + MT_BN - 1) / MT_BN), (unsigned)((n + MT_BM - 1) / MT_BM));
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
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms_total;
    return ok;
}

}  // namespace rb
