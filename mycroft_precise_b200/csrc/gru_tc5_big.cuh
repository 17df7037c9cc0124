// gru_tc5_big.cuh -- K2 for wide networks (24 < H <= 128, F <= 64; BASELINE configs[2]: H = 128, F = 40) on
// tcgen05 + TMEM with the weights streamed through a TMA (cp.async.bulk) pipeline.
//
// A CTA owns 128 streams (the 128 TMEM lanes); thread i <-> stream row i.  TMEM (all 512 columns):
//     D  [0,256)   fp32 accumulators:  phase 1: z | r pre-activations;  phase 2 overwrites the r half with the candidate
//     Ah [256,384) TF32 hi part of the recurrent A operand (h in phase 1, r*h in phase 2)     -- A is read from TMEM
//     Al [384,512) TF32 lo part                                                                  (tcgen05.mma "TS" form)
// Shared memory: x_t operand (hi/lo, K-major canonical layout), h itself in fp32 [unit][row], and a ring of weight
// stages.  Per step:
//   phase 1: D[:, 0:256]   = x . [Wz|Wr]  (SS)  +  h . [Uz|Ur]  (TS)        (kx + 16) k-steps x 3 split terms, N = 256
//            rows: r = hs(D_r + b);  r*h -> TMEM A (hi, lo)                  (tcgen05.ld / tcgen05.st, thread = row)
//   phase 2: D[:, 128:256] = x . Wh (SS)  +  (r*h) . Uh (TS)                (kx + 16) k-steps x 3, N = 128
//            rows: z = hs(D_z + b);  h = z h + (1-z) act(D_h + b);  h -> shared fp32 + TMEM A (hi, lo)
// One thread streams the weight tiles (global, L2 resident, pre-split and laid out per k-step by the host) into the stage
// ring with cp.async.bulk; another issues the MMAs; tcgen05.commit releases each stage back to the producer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gru_tc5.cuh"

namespace pb {

constexpr int TCB_THREADS = 256;            // row threads: two per stream row
constexpr int TCB_BLOCK = 320;              // + a weight-producer warp and an MMA-issuer warp (one lane each), free-running
constexpr int TCB_HP = 128;                 // padded hidden width
constexpr int TCB_KH = TCB_HP / 8;          // recurrent k-steps
constexpr int TCB_MAX_KX = 5;               // F <= 40
constexpr int TCB_STAGES = 6;
constexpr int TCB_STAGE_FLOATS = 4096;      // 16 KB: phase-1 tile (hi + lo, N = 256); phase-2 tiles use half
constexpr int TCB_TMEM_AH = 256, TCB_TMEM_AL = 384;

struct GruTcbW {
    const float* b1;      // [(kx + 16)][4096]  phase-1 tiles
    const float* b2;      // [(kx + 16)][2048]  phase-2 tiles
    const float* bias;    // [3][128]
    const float* wd;      // [128]
    float bd;
    int kx;               // x k-steps = ceil(F / 8)
    int F, H;
    int act;              // candidate activation
    int ract;
    long long* dbg;       // optional: issuer-side cycle counters of CTA 0 (ready wait, full wait, issue, total)
};

struct TcbSmem {
    float stage[TCB_STAGES][TCB_STAGE_FLOATS];
    float ax_hi[2 * TCB_MAX_KX][128][4], ax_lo[2 * TCB_MAX_KX][128][4];
    float hs[TCB_HP][128];                   // h, fp32, [unit][row]
    float bias[3 * TCB_HP];
    float wd[TCB_HP];
    unsigned long long full[TCB_STAGES], empty[TCB_STAGES], done[2], ready[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ void tcb_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void tcb_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                   "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                   "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tcb_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// write 16 values of this row into the TMEM A operand (hi / lo split) at column `col`
__device__ __forceinline__ void tcb_put_a16(uint32_t t_row, int col, const float (&v)[16]) {
    float hi[16], lo[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        hi[e] = __uint_as_float(__float_as_uint(v[e]) & 0xffffe000u);
        lo[e] = v[e] - hi[e];
    }
    tcb_st16(t_row + TCB_TMEM_AH + col, hi);
    tcb_st16(t_row + TCB_TMEM_AL + col, lo);
}

__device__ __forceinline__ void tcb_row_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 256 row threads only

// asynchronous TMEM load of 16 columns and the matching wait (tied to the registers so nothing is consumed early)
__device__ __forceinline__ void tcb_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tcb_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// Block = 8 row warps (two threads per stream row: thread `half` owns hidden units [64 half, 64 half + 64)) + a weight
// producer warp + an MMA issuer warp (one lane each, free-running over the whole scan).
template <bool RING>
__global__ void __launch_bounds__(TCB_BLOCK, 1)
gru_tcb_kernel(GruTcbW W, K2In in, long long n, DecodeParams dp, K2Out out) {
    extern __shared__ __align__(128) unsigned char tcb_raw[];
    TcbSmem& sm = *reinterpret_cast<TcbSmem*>(tcb_raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int kx = W.kx, ksteps = kx + TCB_KH;
    for (int e = tid; e < 3 * TCB_HP; e += TCB_BLOCK) sm.bias[e] = __ldg(W.bias + e);
    for (int e = tid; e < TCB_HP; e += TCB_BLOCK) sm.wd[e] = __ldg(W.wd + e);
    for (int e = tid; e < TCB_HP * 128; e += TCB_BLOCK) (&sm.hs[0][0])[e] = 0.f;
    if (tid == 0) {
        for (int s = 0; s < TCB_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.done[0], 1); mbar_init(&sm.done[1], 1);
        mbar_init(&sm.ready[0], TCB_THREADS); mbar_init(&sm.ready[1], TCB_THREADS);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = sm.tmem_base;
    if (warp == TCB_THREADS / 32) {
        // ---- weight producer: free-running over steps x phases x k-steps, throttled only by the stage ring, so the
        // tiles of the next phase / step arrive while the row threads do their element-wise work
        if (tid == TCB_THREADS) {
            uint32_t it = 0;
            for (int step = 0; step < in.T; ++step)
                for (int phase = 0; phase < 2; ++phase) {
                    const uint32_t tile_bytes = phase == 0 ? 16384u : 8192u;
                    const float* src = phase == 0 ? W.b1 : W.b2;
                    for (int s = 0; s < ksteps; ++s, ++it) {
                        const int st = it % TCB_STAGES;
                        const uint32_t use = it / TCB_STAGES;
                        if (use > 0) mbar_wait(&sm.empty[st], (use - 1) & 1);
                        mbar_expect_tx(&sm.full[st], tile_bytes);
                        bulk_g2s(sm.stage[st], src + (size_t)s * (tile_bytes / 4), tile_bytes, &sm.full[st]);
                    }
                }
        }
        return;
    }
    const uint32_t idesc1 = tc5_idesc(256), idesc2 = tc5_idesc(128);
    if (warp == TCB_THREADS / 32 + 1) {
        // ---- MMA issuer: its own warp, so that no row thread parked in mbarrier.try_wait shares a warp with it
        if (tid == TCB_THREADS + 32) {
            uint32_t it = 0;
            long long c_ready = 0, c_full = 0, c_issue = 0;
            const long long c_begin = clock64();
            for (int step = 0; step < in.T; ++step)
                for (int phase = 0; phase < 2; ++phase) {
                    const int N = phase == 0 ? 256 : 128;
                    long long c0 = clock64();
                    mbar_wait(&sm.ready[phase], step & 1);             // operands of this phase are in place
                    c_ready += clock64() - c0;
                    tc5_fence_after();
                    const uint32_t d = tmem + (phase == 0 ? 0 : TCB_HP);
                    const uint32_t idesc = phase == 0 ? idesc1 : idesc2;
                    for (int s = 0; s < ksteps; ++s, ++it) {
                        const int st = it % TCB_STAGES;
                        c0 = clock64();
                        mbar_wait(&sm.full[st], (it / TCB_STAGES) & 1);
                        const long long c1 = clock64();
                        c_full += c1 - c0;
                        tc5_fence_after();
                        const uint64_t dbh = tc5_desc(sm.stage[st], N * 16, 128);
                        const uint64_t dbl = tc5_desc(sm.stage[st] + 2 * N * 4, N * 16, 128);
                        if (s < kx) {
                            const uint64_t dah = tc5_desc(sm.ax_hi[2 * s], 2048, 128), dal = tc5_desc(sm.ax_lo[2 * s], 2048, 128);
                            tc5_mma(d, dal, dbh, idesc, s > 0);
                            tc5_mma(d, dah, dbl, idesc, 1);
                            tc5_mma(d, dah, dbh, idesc, 1);
                        } else {
                            const uint32_t ah = tmem + TCB_TMEM_AH + 8 * (s - kx), al = tmem + TCB_TMEM_AL + 8 * (s - kx);
                            tcb_mma_ts(d, al, dbh, idesc, 1);
                            tcb_mma_ts(d, ah, dbl, idesc, 1);
                            tcb_mma_ts(d, ah, dbh, idesc, 1);
                        }
                        tc5_commit(&sm.empty[st]);                     // stage reusable once these MMAs have read it
                        c_issue += clock64() - c1;
                    }
                    tc5_commit(&sm.done[phase]);
                }
            if (W.dbg != nullptr && blockIdx.x == 0) {
                W.dbg[0] = c_ready; W.dbg[1] = c_full; W.dbg[2] = c_issue; W.dbg[3] = clock64() - c_begin;
            }
        }
        return;
    }
    // ---- row threads
    const int row = tid & 127, half = tid >> 7, u0 = 64 * half;
    const uint32_t t_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    {   // h = 0 in the TMEM A operand (this thread's 64 units)
        float zero[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) zero[e] = 0.f;
        for (int c = u0; c < u0 + 64; c += 16) { tcb_st16(t_row + TCB_TMEM_AH + c, zero); tcb_st16(t_row + TCB_TMEM_AL + c, zero); }
        tcb_wait_st();
    }
    const long long i = (long long)blockIdx.x * 128 + row;
    const bool valid = i < n;
    int sid = 0;
    RingCursor cur;
    if (RING && valid) {
        sid = in.ids ? in.ids[i] : (int)i;
        const long long ns = in.n_samples[sid];
        cur.init(in, sid, ns >= in.window ? (ns - in.window) / in.hop + 1 : 0);
    }
    // this thread stages the x chunks c with (c & 1) == half: features 4c .. 4c+3
    float xn[TCB_MAX_KX][4];
    auto fetch_x = [&](int step) {
        const float* rp = nullptr;
        if (valid) rp = RING ? cur.next(step) : in.inputs + (i * in.T + step) * W.F;
#pragma unroll
        for (int q = 0; q < TCB_MAX_KX; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int f = 4 * (2 * q + half) + e;
                xn[q][e] = (rp != nullptr && f < W.F) ? __ldg(rp + f) : 0.f;
            }
    };
    fetch_x(0);

#pragma unroll 1
    for (int step = 0; step < in.T; ++step) {
#pragma unroll
        for (int q = 0; q < TCB_MAX_KX; ++q)
            if (q < kx) tc5_put4(sm.ax_hi[2 * q + half], sm.ax_lo[2 * q + half], row, xn[q][0], xn[q][1], xn[q][2], xn[q][3]);
        fence_proxy_async();
        tc5_fence_before();
        mbar_arrive(&sm.ready[0]);                                     // x_t (and, from the previous step, h) are in place
        if (step + 1 < in.T) fetch_x(step + 1);                         // latency hides behind the MMAs
        // ---- phase 0 results: r gate -> r*h becomes the recurrent A operand
        mbar_wait(&sm.done[0], step & 1);
        tc5_fence_after();
        {
            uint32_t ra[16], rb[16];
            tcb_ld16_async(t_row + TCB_HP + u0, ra);
            tcb_ld_wait(ra);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = u0 + 16 * q;
                uint32_t (&cur_r)[16] = (q & 1) ? rb : ra;
                uint32_t (&nxt_r)[16] = (q & 1) ? ra : rb;
                if (q < 3) tcb_ld16_async(t_row + TCB_HP + c + 16, nxt_r);
                float d[16];
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    d[e] = apply_ract(__uint_as_float(cur_r[e]) + sm.bias[TCB_HP + c + e], W.ract) * sm.hs[c + e][row];
                tcb_put_a16(t_row, c, d);
                if (q < 3) tcb_ld_wait(nxt_r);
            }
        }
        tcb_wait_st();
        tc5_fence_before();
        mbar_arrive(&sm.ready[1]);                                     // r*h is in place
        // ---- phase 1 results: z, candidate -> new h (shared fp32 copy + TMEM A operand)
        mbar_wait(&sm.done[1], step & 1);
        tc5_fence_after();
        {
            uint32_t za[16], zb[16], ha[16], hb[16];
            tcb_ld16_async(t_row + u0, za);
            tcb_ld16_async(t_row + TCB_HP + u0, ha);
            tcb_ld_wait(za); tcb_ld_wait(ha);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = u0 + 16 * q;
                uint32_t (&cz)[16] = (q & 1) ? zb : za;
                uint32_t (&ch)[16] = (q & 1) ? hb : ha;
                uint32_t (&nz)[16] = (q & 1) ? za : zb;
                uint32_t (&nh)[16] = (q & 1) ? ha : hb;
                if (q < 3) { tcb_ld16_async(t_row + c + 16, nz); tcb_ld16_async(t_row + TCB_HP + c + 16, nh); }
                float d[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float z = apply_ract(__uint_as_float(cz[e]) + sm.bias[c + e], W.ract);
                    const float hh = apply_act(__uint_as_float(ch[e]) + sm.bias[2 * TCB_HP + c + e], W.act);
                    const float hp = sm.hs[c + e][row];
                    const float hn = (c + e) < W.H ? z * hp + (1.f - z) * hh : 0.f;
                    sm.hs[c + e][row] = hn;
                    d[e] = hn;
                }
                tcb_put_a16(t_row, c, d);
                if (q < 3) { tcb_ld_wait(nz); tcb_ld_wait(nh); }
            }
        }
        tcb_wait_st();
        tc5_fence_before();
    }
    // ---- Dense(1): two partial sums per row
    float part = 0.f;
    for (int j = u0; j < u0 + 64 && j < W.H; ++j) part = fmaf(sm.hs[j][row], sm.wd[j], part);
    float* scratch = &sm.stage[0][0];                                  // every weight tile has been consumed
    if (half == 1) scratch[row] = part;
    tcb_row_sync();
    if (half == 0) epilogue(part + scratch[row] + W.bd, valid, i, sid, dp, out);
    tc5_fence_before();
    tcb_row_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace pb
