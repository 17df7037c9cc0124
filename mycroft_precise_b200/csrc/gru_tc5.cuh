// gru_tc5.cuh -- K2 on the 5th-generation tensor cores (tcgen05 + TMEM) for the default network
// (H <= 24, F <= 16; precise/model.py:77-82 with recurrent_units = 20, 13 MFCCs).
//
// A CTA owns 128 streams = the 128 rows (TMEM lanes) of the accumulator; thread i <-> stream row i for
// all element-wise work.  Per GRU step:
//
//   every thread : writes its row of the A operands [x_t | h] (TF32 hi and lo parts) into shared memory in the
//                  K-major no-swizzle canonical layout  A[k/4][row][k%4]  (one 16-byte store per 4 k's)
//   thread 0     : tcgen05.mma.kind::tf32, M=128:  D1[128x48] = [x|h] . [Wz|Wr]   (5 k-steps x 3 split terms)
//                                                  D2[128x32] =  x    .  Wh        (2 k-steps x 3)       -> commit
//   every thread : tcgen05.ld its row of D1 -> z, r ; writes r*h over the h operand
//   thread 0     : D2 += (r*h) . Uh  (3 k-steps x 3)                                                     -> commit
//   every thread : tcgen05.ld its row of D2 -> candidate ; h = z h + (1 - z) hh
//
// 3xTF32 (a_lo b_hi + a_hi b_lo + a_hi b_hi) keeps fp32-level accuracy, as in gru_mma_kernel.
// Weights (both split parts, 25 KB) stay resident in shared memory for the CTA's lifetime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gru_kernels.cuh"
#include "mfcc_fast.cuh"      // smem_u32, mbarrier helpers

namespace pb {

constexpr int TC5_THREADS = 128;     // row threads
constexpr int TC5_BLOCK = 160;       // + the MMA-issuer warp (a row thread parked in mbarrier.try_wait must not share a warp with it)
constexpr int TC5_KXC = 4;           // x: 16 k's = 4 chunks of 4
constexpr int TC5_KHC = 6;           // h: 24 k's = 6 chunks
constexpr int TC5_N1 = 48;           // z | r  (24 + 24)
constexpr int TC5_N2 = 32;           // candidate (20, padded to a legal N)
constexpr int TC5_TMEM_COLS = 128;   // 48 + 32 -> next power of two

struct GruTc5W {
    const float* b1_hi;   // [10 chunks][48][4]
    const float* b1_lo;
    const float* b2_hi;   // [10 chunks][32][4]
    const float* b2_lo;
    const float* bias;    // [48 + 32]: z(24) r(24) h(32)
    const float* wd;      // [24]
    float bd;
};

struct Tc5Smem {
    float ax_hi[TC5_KXC][128][4], ax_lo[TC5_KXC][128][4];
    float ah_hi[TC5_KHC][128][4], ah_lo[TC5_KHC][128][4];
    float b1_hi[10][TC5_N1][4], b1_lo[10][TC5_N1][4];
    float b2_hi[10][TC5_N2][4], b2_lo[10][TC5_N2][4];
    float bias[TC5_N1 + TC5_N2];
    float wd[24];
    unsigned long long bar[2], ready[2];
    uint32_t tmem_base;
};

// shared-memory matrix descriptor: K-major, no swizzle.  lbo = byte distance between the two 4-element k chunks of one
// K=8 step, sbo = byte distance between 8-row groups (cute/arch/mma_sm100_desc.hpp: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version = 1 at [46,48), layout_type = 0 at [61,64)).
__device__ __forceinline__ uint64_t tc5_desc(const void* p, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_u32(p) & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t tc5_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc5_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void tc5_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// write 4 consecutive k's of this thread's row as hi / lo TF32 parts
__device__ __forceinline__ void tc5_put4(float (*hi)[4], float (*lo)[4], int row, float a, float b, float c, float d) {
    const float v[4] = {a, b, c, d};
    float h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
        l[e] = v[e] - h[e];
    }
    *reinterpret_cast<float4*>(hi[row]) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo[row]) = make_float4(l[0], l[1], l[2], l[3]);
}

// PROJ (opt-in, pb_debug_gru_mode 8; written after round 1's GPU budget was spent -- not yet validated on hardware): the
// input projection x_t.[Wz|Wr|Wh] + b comes from the cache that input_proj_kernel maintains (gru_kernels.cuh), so the
// tensor core only runs the recurrent products: 18 instead of 30 MMAs per step, no x operand tiles.  The row threads fetch
// their 60 cached values while the MMAs run and add them where the non-PROJ kernel adds the bias.
template <int H, int F, bool RING, bool PROJ = false>
__global__ void __launch_bounds__(TC5_BLOCK)
gru_tc5_kernel(GruTc5W W, K2In in, long long n, DecodeParams dp, K2Out out) {
    static_assert(!PROJ || (RING && H == 20), "the cached projection rows hold 3 x 20 columns");
    static_assert(H <= 24 && F <= 16, "operand tiles are sized for the default network");
    extern __shared__ __align__(128) unsigned char tc5_raw[];
    Tc5Smem& sm = *reinterpret_cast<Tc5Smem*>(tc5_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    // ---- one-time setup: weights to shared memory, barriers, TMEM
    for (int e = tid; e < 10 * TC5_N1 * 4; e += TC5_BLOCK) { (&sm.b1_hi[0][0][0])[e] = __ldg(W.b1_hi + e); (&sm.b1_lo[0][0][0])[e] = __ldg(W.b1_lo + e); }
    for (int e = tid; e < 10 * TC5_N2 * 4; e += TC5_BLOCK) { (&sm.b2_hi[0][0][0])[e] = __ldg(W.b2_hi + e); (&sm.b2_lo[0][0][0])[e] = __ldg(W.b2_lo + e); }
    for (int e = tid; e < TC5_N1 + TC5_N2; e += TC5_BLOCK) sm.bias[e] = __ldg(W.bias + e);
    if (tid < 24) sm.wd[tid] = __ldg(W.wd + tid);
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1);
        mbar_init(&sm.ready[0], TC5_THREADS); mbar_init(&sm.ready[1], TC5_THREADS);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(TC5_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t idesc1 = tc5_idesc(TC5_N1), idesc2 = tc5_idesc(TC5_N2);
    if (warp == 4) {
        // ---- MMA issuer (one lane), free-running; hands results back through tcgen05.commit -> mbarrier
        if (tid == TC5_THREADS) {
#pragma unroll 1
            for (int step = 0; step < in.T; ++step) {
                mbar_wait(&sm.ready[0], step & 1);
                tc5_fence_after();
                // D1 = [x|h] . [Wz|Wr]   (PROJ: h part only)
#pragma unroll 1
                for (int s = PROJ ? 2 : 0; s < 5; ++s) {
                    const float* a_hi = s < 2 ? &sm.ax_hi[2 * s][0][0] : &sm.ah_hi[2 * (s - 2)][0][0];
                    const float* a_lo = s < 2 ? &sm.ax_lo[2 * s][0][0] : &sm.ah_lo[2 * (s - 2)][0][0];
                    const uint64_t dah = tc5_desc(a_hi, 2048, 128), dal = tc5_desc(a_lo, 2048, 128);
                    const uint64_t dbh = tc5_desc(sm.b1_hi[2 * s], TC5_N1 * 16, 128), dbl = tc5_desc(sm.b1_lo[2 * s], TC5_N1 * 16, 128);
                    tc5_mma(tmem, dal, dbh, idesc1, s > (PROJ ? 2 : 0));
                    tc5_mma(tmem, dah, dbl, idesc1, 1);
                    tc5_mma(tmem, dah, dbh, idesc1, 1);
                }
                // D2 = x . Wh   (PROJ: nothing, the candidate product starts fresh in the second phase)
#pragma unroll 1
                for (int s = 0; s < (PROJ ? 0 : 2); ++s) {
                    const uint64_t dah = tc5_desc(sm.ax_hi[2 * s], 2048, 128), dal = tc5_desc(sm.ax_lo[2 * s], 2048, 128);
                    const uint64_t dbh = tc5_desc(sm.b2_hi[2 * s], TC5_N2 * 16, 128), dbl = tc5_desc(sm.b2_lo[2 * s], TC5_N2 * 16, 128);
                    tc5_mma(tmem + TC5_N1, dal, dbh, idesc2, s > 0);
                    tc5_mma(tmem + TC5_N1, dah, dbl, idesc2, 1);
                    tc5_mma(tmem + TC5_N1, dah, dbh, idesc2, 1);
                }
                tc5_commit(&sm.bar[0]);
                mbar_wait(&sm.ready[1], step & 1);
                tc5_fence_after();
#pragma unroll 1
                for (int s = 2; s < 5; ++s) {
                    const uint64_t dah = tc5_desc(sm.ah_hi[2 * (s - 2)], 2048, 128), dal = tc5_desc(sm.ah_lo[2 * (s - 2)], 2048, 128);
                    const uint64_t dbh = tc5_desc(sm.b2_hi[2 * s], TC5_N2 * 16, 128), dbl = tc5_desc(sm.b2_lo[2 * s], TC5_N2 * 16, 128);
                    tc5_mma(tmem + TC5_N1, dal, dbh, idesc2, PROJ ? (s > 2) : 1);
                    tc5_mma(tmem + TC5_N1, dah, dbl, idesc2, 1);
                    tc5_mma(tmem + TC5_N1, dah, dbh, idesc2, 1);
                }
                tc5_commit(&sm.bar[1]);
            }
        }
        return;
    }
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);       // this warp's 32 TMEM lanes

    const long long i = (long long)blockIdx.x * TC5_THREADS + tid;
    const bool valid = i < n;
    int sid = 0;
    RingCursor cur;
    if (RING && valid) {
        sid = in.ids ? in.ids[i] : (int)i;
        const long long ns = in.n_samples[sid];
        const long long rel = ns >= in.window ? (ns - in.window) / in.hop + 1 : 0;
        if (PROJ) cur.init_proj(in, sid, rel, PROJ_BLOCK);
        else cur.init(in, sid, rel);
    }
    float h[24];
#pragma unroll
    for (int j = 0; j < 24; ++j) h[j] = 0.f;

#pragma unroll 1
    for (int step = 0; step < in.T; ++step) {
        // ---- this row's x_t and h as A operands (PROJ: h only; the cached projection row is fetched below, under the MMAs)
        const float* prow = nullptr;
        if (PROJ) {
            if (valid) prow = cur.next(step);
        } else {
            float x[16];
#pragma unroll
            for (int f = 0; f < 16; ++f) x[f] = 0.f;
            if (valid) {
                const float* row = RING ? cur.next(step) : in.inputs + (i * in.T + step) * F;
                if (row != nullptr) {
#pragma unroll
                    for (int f = 0; f < F; ++f) x[f] = __ldg(row + f);
                }
            }
#pragma unroll
            for (int c = 0; c < TC5_KXC; ++c) tc5_put4(sm.ax_hi[c], sm.ax_lo[c], tid, x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
        }
#pragma unroll
        for (int c = 0; c < TC5_KHC; ++c) tc5_put4(sm.ah_hi[c], sm.ah_lo[c], tid, h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
        fence_proxy_async();
        tc5_fence_before();
        mbar_arrive(&sm.ready[0]);
        // additive terms of the three gates: the cached projection (bias included) or, for rows before the stream's first
        // frame and in the non-PROJ kernel, the bias
        float pz[24], pr[24], ph[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) { pz[j] = sm.bias[j]; pr[j] = sm.bias[24 + j]; ph[j] = sm.bias[TC5_N1 + j]; }
        if (PROJ && prow != nullptr) {
            // the cache is laid out for the mma.sync scan's fragments (proj_off): unit pairs (2u, 2u + 1) of a gate are adjacent
#pragma unroll
            for (int u = 0; u < 10; ++u) {
                const int nt0 = (2 * u) / 8, tt = ((2 * u) & 7) >> 1;
                const float2 a = __ldg(reinterpret_cast<const float2*>(prow + proj_off(nt0, sid & 15, tt)));
                const float2 b = __ldg(reinterpret_cast<const float2*>(prow + proj_off(3 + nt0, sid & 15, tt)));
                const float2 c = __ldg(reinterpret_cast<const float2*>(prow + proj_off(6 + nt0, sid & 15, tt)));
                pz[2 * u] = a.x; pz[2 * u + 1] = a.y; pr[2 * u] = b.x; pr[2 * u + 1] = b.y; ph[2 * u] = c.x; ph[2 * u + 1] = c.y;
            }
        }
        mbar_wait(&sm.bar[0], step & 1);
        tc5_fence_after();
        float z[24];
        {
            float d[16];
            float zr[48];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                tc5_ld16(t_row + 16 * q, d);
#pragma unroll
                for (int e = 0; e < 16; ++e) zr[16 * q + e] = d[e];
            }
#pragma unroll
            for (int j = 0; j < 24; ++j) {
                z[j] = hard_sigmoid(zr[j] + pz[j]);
                const float r = hard_sigmoid(zr[24 + j] + pr[j]);
                zr[j] = r * h[j];                                    // r * h, reusing the array
            }
#pragma unroll
            for (int c = 0; c < TC5_KHC; ++c) tc5_put4(sm.ah_hi[c], sm.ah_lo[c], tid, zr[4 * c], zr[4 * c + 1], zr[4 * c + 2], zr[4 * c + 3]);
        }
        fence_proxy_async();
        tc5_fence_before();
        mbar_arrive(&sm.ready[1]);
        mbar_wait(&sm.bar[1], step & 1);
        tc5_fence_after();
        {
            float d[16], hh[32];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                tc5_ld16(t_row + TC5_N1 + 16 * q, d);
#pragma unroll
                for (int e = 0; e < 16; ++e) hh[16 * q + e] = d[e];
            }
#pragma unroll
            for (int j = 0; j < 24; ++j) h[j] = j < H ? z[j] * h[j] + (1.f - z[j]) * (hh[j] + ph[j]) : 0.f;
        }
    }
    float logit = W.bd;
#pragma unroll
    for (int j = 0; j < H; ++j) logit = fmaf(h[j], sm.wd[j], logit);
    epilogue(logit, valid, i, sid, dp, out);
    tc5_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC5_TMEM_COLS) : "memory");
}

}  // namespace pb
