// rb_mtx_tc.cu -- the matrix consumer on the 5th-generation tensor cores.
//
// result[s][t][c] = sum_b DC[s][b][c] * sky[b][t][c]  (util/cmatrix.c:420-475 cm_multiply: float in, double
// accumulation, float out) as three GEMMs, one per colour channel, with fp32 accuracy from TF32 hardware:
//
//   * split precision (3xTF32): every operand x is stored as hi = tf32(x) (round to nearest, 10-bit mantissa) and
//     lo = x - hi (exact in fp32; the hardware drops its last 2 of 13 bits), and the product is
//     hi*hi + hi*lo + lo*hi -- what is left out, lo*lo and the dropped bits, is below 2^-20 of the product;
//   * a pre-pass (k_split_*) writes the operands channel-planar and K-major: A planes [hi|lo][3][Mp][Kp], B planes
//     transposed [hi|lo][3][Np][Kp], zero padded to the tile sizes, so that a TMA box of 32 floats (128 bytes, the
//     swizzle atom) x 128 rows is one operand tile;
//   * k_mtx_tc: one CTA per 128 x 64 output tile and ALL THREE channels, two CTAs per SM (96 KB of shared memory and
//     256 TMEM columns each) so that one CTA's epilogue overlaps the other's MMAs.  Warp 0 (one lane) feeds a
//     2-stage ring of {A hi, A lo, B hi, B lo} tiles with cp.async.bulk.tensor (TMA, 128-byte swizzle, mbarrier
//     complete_tx); warp 1 (one lane) issues tcgen05.mma kind::tf32, M = 128, N = 64, K = 8 per instruction, 12 per
//     ring stage (4 k-steps x 3 products), and releases each stage with tcgen05.commit.
//   * two-level accumulation: the tensor core's fp32 accumulator loses ~2e-8 of the sum per product (measured:
//     3e-5 after 2305 non-negative products), so the MMAs accumulate SEGMENTS of 128 products in TMEM columns
//     [0, 64); warps 2-5 fold each finished segment into per-thread fp32 registers (round to nearest) with
//     tcgen05.ld, hand the segment buffer back (mbarrier), and park a finished channel in TMEM columns
//     [64 + 64 c, 128 + 64 c) with tcgen05.st.  Relative error stays at ~3e-6 for any inner dimension.
//   * epilogue: the three channel sums are read back, interleaved in registers, staged through the (now idle)
//     ring memory and written to HBM as contiguous runs of [column][channel] floats -- the output, 10.5 GB at
//     BASELINE configs[1] size, is what the kernel has to stream.
// Nothing of a library GEMM is used; descriptors are built by hand (bit layouts: PTX ISA "tcgen05 matrix / instruction
// descriptor").  Every mbarrier wait is bounded: a protocol error traps instead of hanging the GPU.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include "rb_engine.cuh"

namespace rb {

constexpr int TC_BM = 128, TC_BN = 64, TC_BK = 32, TC_STAGES = 2, TC_THREADS = 192;
// Thread-block cluster along the output columns (k_mtx_tc<TC_CL = 2, ...>): the CTAs of a cluster work on the same
// 128 rows of A, so each loads 128 / TC_CL rows of every A tile and MULTICASTS them into the shared memory of all of
// them (one L2 read feeds TC_CL SMs); a stage is handed back when the MMAs of every CTA of the cluster have read it
// (tcgen05.commit with a CTA mask on the `empty` barriers, which count TC_CL arrivals).
// Measured (tools/mtx_variants.sh): at K = 145 clusters of 2 / 4 are 5 / 12 % slower (multicast does not reduce the
// bytes delivered into each SM, and the kernel is not bound by the L2 reads); at K = 2305 a cluster of 2 is 7 % faster
// (152.6 TFLOP/s).  The host picks: clusters for K >= 512.
constexpr int TC_CL_MAX = 2;
// k-blocks per accumulation segment: 4 (128 products) in general; a product whose whole inner dimension fits 8 blocks
// (K <= 256, the MF:1 daylight-coefficient case) is one segment per channel (3.3e-6 instead of 2.7e-6, 2.5 % faster)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4, TC_B_BYTES = TC_BN * TC_BK * 4;     // operand tiles: rows of 128 bytes
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;                   // A hi, A lo, B hi, B lo = 48 KB
constexpr int TC_RING_BYTES = TC_STAGES * TC_STAGE_BYTES;    // 96 KB
constexpr int TC_SMEM_BYTES = TC_RING_BYTES + 256 + 1024;    // + barriers + alignment slack
constexpr int TC_STG_ROW = 32 * 3 * 4 + 16;                  // staged row: 32 columns x 3 channels, padded (conflict-free STS.128)

// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (unsigned spin = 0;; spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();              // a protocol error must not hang the device
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ... arriving on the barrier at the same offset in every CTA of the mask
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], M = 128, N = 128, K = 8 (TF32), one thread issues for the CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor of a K-major tile stored by TMA with the 128-byte swizzle: rows of 128 bytes,
// 8-row groups 1024 bytes apart (stride byte offset), descriptor version 1, layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);                  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                   // leading byte offset (unused with swizzle), [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                         // stride byte offset, [32,46)
    d |= (uint64_t)1 << 46;                                   // version, [46,48)
    d |= (uint64_t)2 << 61;                                   // layout type, [61,64)
    return d;
}
// instruction descriptor: D = F32 [4,6) = 1, A = B = TF32 [7,10), [10,13) = 2, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                   "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}

// ------------------------------------------------------------ pre-pass ----
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// A[nr][K][3] -> planes [hi|lo][3][Mp][Kp] (zeroed beforehand)
__global__ void k_split_a(const float* __restrict__ A, float* __restrict__ P, size_t nr, int K, size_t Mp, int Kp) {
    const size_t n = nr * (size_t)K * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % 3);
        const size_t rk = i / 3;
        const int k = (int)(rk % K);
        const size_t r = rk / K;
        const float x = A[i], hi = tf32_rn(x);
        const size_t o = ((size_t)ch * Mp + r) * Kp + k;
        P[o] = hi;
        P[o + 3 * Mp * (size_t)Kp] = x - hi;
    }
}
// B[K][nc][3] -> planes [hi|lo][3][Np][Kp], transposed to K-major (zeroed beforehand); a 32 x 32 tile per CTA and channel
__global__ void k_split_b(const float* __restrict__ B, float* __restrict__ P, int K, size_t nc, size_t Np, int Kp) {
    __shared__ float t[32][33];
    const int ch = blockIdx.z;
    const size_t c0 = (size_t)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int k = k0 + j;
        const size_t c = c0 + threadIdx.x;
        t[j][threadIdx.x] = (k < K && c < nc) ? B[((size_t)k * nc + c) * 3 + ch] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const size_t c = c0 + j;
        const int k = k0 + threadIdx.x;
        if (c < nc && k < K) {
            const float x = t[threadIdx.x][j], hi = tf32_rn(x);
            const size_t o = ((size_t)ch * Np + c) * Kp + k;
            P[o] = hi;
            P[o + 3 * Np * (size_t)Kp] = x - hi;
        }
    }
}

// ---------------------------------------------------------------- GEMM ----
template <int TC_CL, int TC_SEG_KB>
__global__ void __cluster_dims__(TC_CL, 1, 1) __launch_bounds__(TC_THREADS, 2)
k_mtx_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
         size_t nr, size_t nc, int nkb) {
    constexpr uint16_t CL_MASK = (uint16_t)((1u << TC_CL) - 1u);
    const uint32_t crank = TC_CL > 1 ? cluster_rank() : 0u;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_RING_BYTES);       // full[S], empty[S], seg_full, seg_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2);
    const uint32_t ring = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TC_STAGES);
    const uint32_t seg_full = smem_u32(bars + 2 * TC_STAGES), seg_empty = smem_u32(bars + 2 * TC_STAGES + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;

    if (threadIdx.x == 0) {
        // a stage is empty once the MMAs of EVERY CTA of the cluster have read it: each of them will be written by all
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, TC_CL); }
        mbar_init(seg_full, 1);
        mbar_init(seg_empty, 4);                      // one arrival per folding warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (TC_CL > 1) cluster_sync_all();                // the peers' barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            const int iters = 3 * nkb;
            for (int it = 0; it < iters; it++) {
                const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                const int ch = it / nkb, kb = it % nkb;
                mbar_wait(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, TC_STAGE_BYTES);
                const uint32_t st = ring + s * TC_STAGE_BYTES;
                if (TC_CL > 1) {                  // my 128 / TC_CL rows of the A tiles, into every CTA of the cluster
                    constexpr int AR = TC_BM / TC_CL, AB = AR * TC_BK * 4;
                    tma_load_3d_mc(st + crank * AB, &tmA, full0 + 8 * s, kb * TC_BK, m0 + (int)crank * AR, ch, CL_MASK);
                    tma_load_3d_mc(st + TC_A_BYTES + crank * AB, &tmA, full0 + 8 * s, kb * TC_BK, m0 + (int)crank * AR, 3 + ch, CL_MASK);
                } else {
                    tma_load_3d(st, &tmA, full0 + 8 * s, kb * TC_BK, m0, ch);
                    tma_load_3d(st + TC_A_BYTES, &tmA, full0 + 8 * s, kb * TC_BK, m0, 3 + ch);
                }
                tma_load_3d(st + 2 * TC_A_BYTES, &tmB, full0 + 8 * s, kb * TC_BK, n0, ch);
                tma_load_3d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tmB, full0 + 8 * s, kb * TC_BK, n0, 3 + ch);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(TC_BM, TC_BN);
            int it = 0, seg = 0;
            for (int ch = 0; ch < 3; ch++)
                for (int s0 = 0; s0 < nkb; s0 += TC_SEG_KB, seg++) {
                    mbar_wait(seg_empty, (seg & 1) ^ 1);      // the folding warps have read the previous segment
                    tc_fence_after();
                    const int s1 = min(s0 + TC_SEG_KB, nkb);
                    for (int kb = s0; kb < s1; kb++, it++) {
                        const int s = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                        mbar_wait(full0 + 8 * s, ph);
                        tc_fence_after();
                        const uint32_t st = ring + s * TC_STAGE_BYTES;
                        const uint64_t ah = tc_smem_desc(st), al = tc_smem_desc(st + TC_A_BYTES);
                        const uint64_t bh = tc_smem_desc(st + 2 * TC_A_BYTES), bl = tc_smem_desc(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; k++) {
                            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);      // 8 TF32 = 32 bytes along K inside the swizzle atom
                            tc_mma_tf32(tmem, ah + adv, bh + adv, idesc, (kb > s0 || k) ? 1u : 0u);
                            tc_mma_tf32(tmem, ah + adv, bl + adv, idesc, 1u);
                            tc_mma_tf32(tmem, al + adv, bh + adv, idesc, 1u);
                        }
                        if (TC_CL > 1) tc_commit_mc(empty0 + 8 * s, CL_MASK);      // the stage is free once these MMAs have read it:
                        else tc_commit(empty0 + 8 * s);                            // every producer of the cluster is told
                    }
                    tc_commit(seg_full);                    // ... and the segment is complete once they have finished
                }
        }
        __syncwarp();
    } else {
        // ---------------- fold segments (TMEM -> registers), then TMEM -> staged rows -> HBM ----------------
        const int q = warp & 3;                         // the TMEM lane quarter this warp may touch
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        int seg = 0;
        for (int ch = 0; ch < 3; ch++) {
            float acc[4][16];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int j = 0; j < 16; j++) acc[c][j] = 0.f;
            for (int s0 = 0; s0 < nkb; s0 += TC_SEG_KB, seg++) {
                mbar_wait(seg_full, seg & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t v[16];
                    tmem_ld16(tq + 16 * c, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[c][j] += __uint_as_float(v[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(seg_empty);
            }
#pragma unroll
            for (int c = 0; c < 4; c++) tmem_st16(tq + (uint32_t)(TC_BN + TC_BN * ch + 16 * c), acc[c]);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // all MMAs have completed (the last segment was folded), so the ring is idle: 32 staged rows per warp
        uint8_t* stg = smem + (size_t)q * 32 * TC_STG_ROW;
        const bool vec_ok = (nc % 4) == 0;              // 16-byte aligned rows
        for (int h = 0; h < 2; h++) {
            const size_t cbase = (size_t)n0 + 32 * h;
            if (cbase >= nc) break;
#pragma unroll
            for (int c16 = 0; c16 < 2; c16++) {
                uint32_t v0[16], v1[16], v2[16];
                const uint32_t ta = tq + (uint32_t)(TC_BN + 32 * h + 16 * c16);
                tmem_ld16(ta, v0);
                tmem_ld16(ta + TC_BN, v1);
                tmem_ld16(ta + 2 * TC_BN, v2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float4* dst = reinterpret_cast<float4*>(stg + (size_t)lane * TC_STG_ROW + (size_t)(16 * c16) * 12);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    dst[0] = make_float4(__uint_as_float(v0[j]), __uint_as_float(v1[j]), __uint_as_float(v2[j]), __uint_as_float(v0[j + 1]));
                    dst[1] = make_float4(__uint_as_float(v1[j + 1]), __uint_as_float(v2[j + 1]), __uint_as_float(v0[j + 2]), __uint_as_float(v1[j + 2]));
                    dst[2] = make_float4(__uint_as_float(v2[j + 2]), __uint_as_float(v0[j + 3]), __uint_as_float(v1[j + 3]), __uint_as_float(v2[j + 3]));
                    dst += 3;
                }
            }
            __syncwarp();
            const int ncol = (int)min((size_t)32, nc - cbase);
            const int nflt = ncol * 3;
            const size_t row0 = (size_t)m0 + q * 32;
            const int nrow = (int)min((size_t)32, nr > row0 ? nr - row0 : 0);
            if (vec_ok && nflt == 96) {
                // full tile: 24 float4 per row; a warp instruction covers 4 rows x 8 float4 ... written as 3 float4
                // columns per lane group so that every store instruction fills whole 128-byte lines
                const int sub = lane >> 3, f0 = lane & 7;             // 4 rows per pass, 8 lanes per row
#pragma unroll 2
                for (int r4 = 0; r4 < nrow; r4 += 4) {
                    const int rr = r4 + sub;
                    if (rr < nrow) {
                        const float4* src = reinterpret_cast<const float4*>(stg + (size_t)rr * TC_STG_ROW);
                        float4* dst = reinterpret_cast<float4*>(C + ((row0 + rr) * nc + cbase) * 3);
                        const float4 a = src[f0], b = src[f0 + 8], c = src[f0 + 16];
                        dst[f0] = a; dst[f0 + 8] = b; dst[f0 + 16] = c;
                    }
                }
            } else if (vec_ok && (nflt % 4) == 0) {
                const int per = nflt / 4;               // float4 per row
                for (int i = lane; i < nrow * per; i += 32) {
                    const int rr = i / per, f = i - rr * per;
                    reinterpret_cast<float4*>(C + ((row0 + rr) * nc + cbase) * 3)[f] =
                        reinterpret_cast<const float4*>(stg + (size_t)rr * TC_STG_ROW)[f];
                }
            } else {
                for (int i = lane; i < nrow * nflt; i += 32) {
                    const int rr = i / nflt, f = i - rr * nflt;
                    C[((row0 + rr) * nc + cbase) * 3 + f] = reinterpret_cast<const float*>(stg + (size_t)rr * TC_STG_ROW)[f];
                }
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (TC_CL > 1) cluster_sync_all();                // no CTA leaves while a peer may still arrive on its barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}


// ------------------------------------------------------- GEMM, CTA pairs ----
// k_mtx_tc2 (EXPERIMENTAL, off unless RB_MTX_2CTA is set): the same product on PAIRS of CTAs (tcgen05 cta_group::2,
// UMMA M = 256, N = 128) for inner dimensions of at most 8 k-blocks (K <= 256: one accumulation segment per channel,
// accumulated in place).  CTA r of a pair holds 128 rows of A and 64 of the tile's 128 columns of B in its shared
// memory; the leader's MMAs read both halves of B and write, in each CTA's tensor memory, that CTA's 128 x 128
// accumulator (three channels: 384 columns).  Per byte delivered into an SM the pair computes twice what k_mtx_tc does.
//   * every CTA: warp 0 feeds its own ring by TMA; warps 2-5 are the epilogue;
//   * CTA 1: warp 1 relays "my stage is full" to the leader (remote mbarrier arrive);
//   * CTA 0 (leader): warp 1 issues the MMAs when both halves of a stage are full, frees the stage in BOTH CTAs and
//     finally signals both epilogues with multicast commits.
// Measured at configs[1] size: identical results to k_mtx_tc<1, 8> (3.26e-6), 8.8 ms with a ring of 3 stages (10.0 with
// 2, 9.0 with 4) against 6.8 ms: with the whole tensor memory taken by one CTA per SM nothing overlaps the epilogue
// (196 KB of rows per CTA) or the per-tile set-up, which costs more than the halved operand stream saves.  What it
// needs is a persistent tile loop whose epilogue runs under the next tile's loads -- kept as the base for that.
constexpr int TC2_BN = 128;                       // columns of the pair's tile
#ifndef RB_TC2_STAGES
#define RB_TC2_STAGES 3
#endif
constexpr int TC2_STAGES = RB_TC2_STAGES;         // one CTA per SM: the whole shared memory is ring (bytes in flight are what the load pipeline lives on)
constexpr int TC2_RING_BYTES = TC2_STAGES * TC_STAGE_BYTES;
constexpr int TC2_SMEM_BYTES = TC2_RING_BYTES + 256 + 1024;
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (unsigned spin = 0;; spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
k_mtx_tc2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
          size_t nr, size_t nc, int nkb) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // barriers: full[S], empty[S], peerfull[S] (used in the leader), accfull
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC2_RING_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC2_STAGES + 1);
    const uint32_t ring = smem_u32(smem);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TC2_STAGES), peer0 = smem_u32(bars + 2 * TC2_STAGES);
    const uint32_t accfull = smem_u32(bars + 3 * TC2_STAGES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_rank();
    const int m0 = (int)(blockIdx.x >> 1) * 256 + (int)crank * TC_BM;       // my 128 rows
    const int n0 = (int)blockIdx.y * TC2_BN;                                // the pair's 128 columns
    const int nb0 = n0 + (int)crank * TC_BN;                                // my half of B

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC2_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); mbar_init(peer0 + 8 * s, 1); }
        mbar_init(accfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int iters = 3 * nkb;

    if (warp == 0) {
        if (lane == 0) {                              // TMA producer of this CTA's halves
            for (int it = 0; it < iters; it++) {
                const int s = it % TC2_STAGES, ph = (it / TC2_STAGES) & 1;
                const int ch = it / nkb, kb = it % nkb;
                mbar_wait_cluster(empty0 + 8 * s, ph ^ 1);
                mbar_expect_tx(full0 + 8 * s, TC_STAGE_BYTES);
                const uint32_t st = ring + s * TC_STAGE_BYTES;
                tma_load_3d(st, &tmA, full0 + 8 * s, kb * TC_BK, m0, ch);
                tma_load_3d(st + TC_A_BYTES, &tmA, full0 + 8 * s, kb * TC_BK, m0, 3 + ch);
                tma_load_3d(st + 2 * TC_A_BYTES, &tmB, full0 + 8 * s, kb * TC_BK, nb0, ch);
                tma_load_3d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tmB, full0 + 8 * s, kb * TC_BK, nb0, 3 + ch);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            if (crank != 0) {                         // relay: my stage is full -> the leader's peerfull barrier
                for (int it = 0; it < iters; it++) {
                    const int s = it % TC2_STAGES, ph = (it / TC2_STAGES) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    mbar_arrive_remote(mapa_rank(peer0 + 8 * s, 0));
                }
            } else {                                  // MMA issuer for the pair
                constexpr uint32_t idesc = tc_idesc(256, TC2_BN);
                for (int it = 0; it < iters; it++) {
                    const int s = it % TC2_STAGES, ph = (it / TC2_STAGES) & 1;
                    const int ch = it / nkb, kb = it % nkb;
                    mbar_wait(full0 + 8 * s, ph);
                    mbar_wait_cluster(peer0 + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t st = ring + s * TC_STAGE_BYTES;
                    const uint64_t ah = tc_smem_desc(st), al = tc_smem_desc(st + TC_A_BYTES);
                    const uint64_t bh = tc_smem_desc(st + 2 * TC_A_BYTES), bl = tc_smem_desc(st + 2 * TC_A_BYTES + TC_B_BYTES);
                    const uint32_t d = tmem + (uint32_t)(TC2_BN * ch);
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; k++) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                        tc_mma_tf32_2cta(d, ah + adv, bh + adv, idesc, (kb || k) ? 1u : 0u);
                        tc_mma_tf32_2cta(d, ah + adv, bl + adv, idesc, 1u);
                        tc_mma_tf32_2cta(d, al + adv, bh + adv, idesc, 1u);
                    }
                    tc_commit_2cta(empty0 + 8 * s, 3);     // the stage is free in both CTAs once these MMAs have read it
                }
                tc_commit_2cta(accfull, 3);                // ... and both accumulators are complete once they have finished
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: TMEM -> staged rows -> HBM (this CTA's 128 rows x 128 columns x 3 channels) ----------------
        const int q = warp & 3;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        mbar_wait_cluster(accfull, 0);
        tc_fence_after();
        uint8_t* stg = smem + (size_t)q * 32 * TC_STG_ROW;
        const bool vec_ok = (nc % 4) == 0;
        for (int h = 0; h < TC2_BN / 32; h++) {
            const size_t cbase = (size_t)n0 + 32 * h;
            if (cbase >= nc) break;
#pragma unroll
            for (int c16 = 0; c16 < 2; c16++) {
                uint32_t v0[16], v1[16], v2[16];
                const uint32_t ta = tq + (uint32_t)(32 * h + 16 * c16);
                tmem_ld16(ta, v0);
                tmem_ld16(ta + TC2_BN, v1);
                tmem_ld16(ta + 2 * TC2_BN, v2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float4* dst = reinterpret_cast<float4*>(stg + (size_t)lane * TC_STG_ROW + (size_t)(16 * c16) * 12);
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    dst[0] = make_float4(__uint_as_float(v0[j]), __uint_as_float(v1[j]), __uint_as_float(v2[j]), __uint_as_float(v0[j + 1]));
                    dst[1] = make_float4(__uint_as_float(v1[j + 1]), __uint_as_float(v2[j + 1]), __uint_as_float(v0[j + 2]), __uint_as_float(v1[j + 2]));
                    dst[2] = make_float4(__uint_as_float(v2[j + 2]), __uint_as_float(v0[j + 3]), __uint_as_float(v1[j + 3]), __uint_as_float(v2[j + 3]));
                    dst += 3;
                }
            }
            __syncwarp();
            const int ncol = (int)min((size_t)32, nc - cbase);
            const int nflt = ncol * 3;
            const size_t row0 = (size_t)m0 + q * 32;
            const int nrow = (int)min((size_t)32, nr > row0 ? nr - row0 : 0);
            if (vec_ok && nflt == 96) {
                const int sub = lane >> 3, f0 = lane & 7;
#pragma unroll 2
                for (int r4 = 0; r4 < nrow; r4 += 4) {
                    const int rr = r4 + sub;
                    if (rr < nrow) {
                        const float4* src = reinterpret_cast<const float4*>(stg + (size_t)rr * TC_STG_ROW);
                        float4* dst = reinterpret_cast<float4*>(C + ((row0 + rr) * nc + cbase) * 3);
                        const float4 a = src[f0], b = src[f0 + 8], c = src[f0 + 16];
                        dst[f0] = a; dst[f0 + 8] = b; dst[f0 + 16] = c;
                    }
                }
            } else if (vec_ok && (nflt % 4) == 0) {
                const int per = nflt / 4;
                for (int i = lane; i < nrow * per; i += 32) {
                    const int rr = i / per, f = i - rr * per;
                    reinterpret_cast<float4*>(C + ((row0 + rr) * nc + cbase) * 3)[f] =
                        reinterpret_cast<const float4*>(stg + (size_t)rr * TC_STG_ROW)[f];
                }
            } else {
                for (int i = lane; i < nrow * nflt; i += 32) {
                    const int rr = i / nflt, f = i - rr * nflt;
                    C[((row0 + rr) * nc + cbase) * 3 + f] = reinterpret_cast<const float*>(stg + (size_t)rr * TC_STG_ROW)[f];
                }
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();                               // neither CTA leaves while the other may still read its shared memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---------------------------------------------------------------- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_plane_map(EncodeTiledFn enc, CUtensorMap* map, float* base, size_t rows, int Kp, int box_rows, std::string& err) {
    const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, 6};
    const cuuint64_t strides[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)rows * Kp * 4};
    const cuuint32_t box[3] = {TC_BK, (cuuint32_t)box_rows, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return false; }
    return true;
}

#define TCK(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); goto done; } \
    } while (0)

// C[n][nc][3] = A[n][ni][3] x B[ni][nc][3]; all three buffers in device memory.  kernel_ms += device time of the
// pre-pass of A and of the GEMM (the pre-pass of B is done once per call by the caller through tc_prepare_b).
bool mtx_multiply_tc(cudaStream_t stream, const float* A, size_t n, size_t ni, const float* Bplanes, size_t Np, int Kp,
                     size_t nc, float* C, float* Aplanes, size_t Mp_cap, double* kernel_ms, std::string& err) {
    static EncodeTiledFn enc = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    bool ok = false;
    {
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        TCK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) { err = "cuTensorMapEncodeTiled is not available in this driver"; goto done; }
        enc = (EncodeTiledFn)fn;
    }
    const size_t Mp = (n + TC_BM - 1) / TC_BM * TC_BM;
    if (Mp > Mp_cap) { err = "internal: A plane buffer too small"; goto done; }
    TCK(cudaEventCreate(&e0)); TCK(cudaEventCreate(&e1));
    const int nkb = Kp / TC_BK;
    const int cl = nkb >= 16 ? 2 : 1;                // clusters (A tiles multicast) pay for long inner dimensions only
    CUtensorMap tmA, tmB;
    if (!make_plane_map(enc, &tmA, Aplanes, Mp, Kp, TC_BM / cl, err) ||
        !make_plane_map(enc, &tmB, const_cast<float*>(Bplanes), Np, Kp, TC_BN, err)) goto done;
    TCK(cudaEventRecord(e0, stream));
    TCK(cudaMemsetAsync(Aplanes, 0, 6 * Mp * (size_t)Kp * sizeof(float), stream));
    k_split_a<<<148 * 16, 256, 0, stream>>>(A, Aplanes, n, (int)ni, Mp, Kp);
    dim3 grid((unsigned)(Np / TC_BN), (unsigned)(Mp / TC_BM));      // Np is padded to whole clusters of column tiles
    if (cl == 2) {
        TCK(cudaFuncSetAttribute(k_mtx_tc<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        k_mtx_tc<2, 4><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tmA, tmB, C, n, nc, nkb);
    } else if (nkb <= 8 && getenv("RB_MTX_2CTA") && (Mp % 256) == 0 && (Np % TC2_BN) == 0) {
        TCK(cudaFuncSetAttribute(k_mtx_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
        dim3 grid2((unsigned)(2 * (Mp / 256)), (unsigned)(Np / TC2_BN));
        k_mtx_tc2<<<grid2, TC_THREADS, TC2_SMEM_BYTES, stream>>>(tmA, tmB, C, n, nc, nkb);
    } else if (nkb <= 8) {
        TCK(cudaFuncSetAttribute(k_mtx_tc<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        k_mtx_tc<1, 8><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tmA, tmB, C, n, nc, nkb);
    } else {
        TCK(cudaFuncSetAttribute(k_mtx_tc<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        k_mtx_tc<1, 4><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tmA, tmB, C, n, nc, nkb);
    }
    TCK(cudaEventRecord(e1, stream));
    TCK(cudaGetLastError());
    TCK(cudaStreamSynchronize(stream));
    float ms = 0; TCK(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms += ms;
    ok = true;
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return ok;
}

// B[ni][nc][3] (device) -> planes; returns the device buffer (caller frees) and its padded sizes
bool tc_prepare_b(cudaStream_t stream, const float* B, size_t ni, size_t nc, float** planes, size_t* Np_out, int* Kp_out,
                  std::string& err) {
    const int Kp = (int)((ni + TC_BK - 1) / TC_BK * TC_BK);
    const size_t Np = (nc + TC_BN * TC_CL_MAX - 1) / (TC_BN * TC_CL_MAX) * (TC_BN * TC_CL_MAX);      // whole clusters of column tiles
    float* P = nullptr;
    cudaError_t e = cudaMalloc(&P, 6 * Np * (size_t)Kp * sizeof(float));
    if (e != cudaSuccess) { err = std::string("cudaMalloc (B planes): ") + cudaGetErrorString(e); return false; }
    cudaMemsetAsync(P, 0, 6 * Np * (size_t)Kp * sizeof(float), stream);
    dim3 grid((unsigned)((nc + 31) / 32), (unsigned)((ni + 31) / 32), 3);
    k_split_b<<<grid, dim3(32, 8), 0, stream>>>(B, P, (int)ni, nc, Np, Kp);
    e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("k_split_b: ") + cudaGetErrorString(e); cudaFree(P); return false; }
    *planes = P; *Np_out = Np; *Kp_out = Kp;
    return true;
}

}  // namespace rb
