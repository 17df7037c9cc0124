// mfcc_tc.cuh -- host tables, CPU model and device helpers of the tensor-core MFCC tick (the kernel itself: mfcc_tc2.cuh).
//
// Replaces, per frame, np.fft.rfft(frame, n=512), the power spectrum, the mel filterbank, log, DCT and c0 of
// sonopy.mfcc_spec as the reference calls it (precise/vectorization.py:36-39), for the stateful tick
// (Listener.update_vectors, precise/network_runner.py:125-146).
//
// Decomposition.  n = n2 + 32 q (n2 < 32, q < 16), k = 16 m + r:
//     Y_r[n2]     = sum_q x[n2 + 32 q] w16^(q r)                      16-point DFT of REAL data, CUDA cores (fp32)
//     X[16 m + r] = sum_n2 Y_r[n2] w512^(n2 r) w32^(n2 m)             eight 64 x 64 real GEMM blocks, tensor cores
// Block 0 takes [Y_0 | Y_8] (both real), block r = 1..7 takes [Re Y_r | Im Y_r] and also yields X[16 m + 16 - r] from
// conj(Y_r) = Y_(16-r).  Operands are fp16 hi + lo pieces, three passes (a_lo b_hi + a_hi b_lo + a_hi b_hi, fp32
// accumulate): 6e-7 of the peak bin against a float64 FFT, the accuracy of the fp32 FFT kernels.
// The CPU model below (same butterfly, same operand tables read through the same layout arithmetic) is what
// tests/test_tc_dft_host_model.py checks without a device.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

#include "gru_tc5.cuh"        // tc5_desc, tc5_commit, tcgen05 fences, tc5_ld16
#include "mfcc_fast.cuh"      // mbarrier helpers
#include "mfcc_kernels.cuh"   // StreamState, frames_ready, K1_EPS

namespace pb {

constexpr int TCD_BLOCKS = 8;                 // GEMM blocks per frame
constexpr int TCD_KSTEPS = 4;                 // K = 64 per block = 4 MMA K-steps of 16
constexpr int TCD_MAX_NEW = 4;                // frames a stream may release per tick
constexpr int TCD_MAX_FILT = 22;              // n_filt + 2 accumulator slots of 128 floats must fit
constexpr int TCD_MAX_OUT = 16;
constexpr float TCD_IN_SCALE = 0.015625f;     // 2^-6 folded into the int16 -> float conversion; the butterfly returns 2 Y
constexpr float TCD_A_SCALE = 0.03125f;       // => A operands hold Y * 2^-5: |A| <= 16384, and <= 32768 < fp16 max after the shift below
// The kernel transforms x - x[0] (a constant only moves X[0]; constant input then gives exact zeros everywhere else, like the
// float64 reference and the FFT kernels): Y_0 loses 16 x[0], i.e. the butterfly output 2 * 16 * IN_SCALE * x[0], and the
// X[0] accumulator gets 32 times that back.
constexpr float TCD_X0_Y = 32.f * TCD_IN_SCALE;          // 0.5
constexpr float TCD_X0_D = 32.f * TCD_X0_Y;              // 16

// ---------------------------------------------------------------------------------------------------------------------
// 16-point DFT of real data, outputs scaled by 2: yr[k] + i yi[k] = 2 * sum_q x[q] w16^(q k), k = 0..8 (yi[0] = yi[8] = 0).
// Packed as an 8-point complex FFT of z[n] = x[2n] + i x[2n+1] plus the real-input split.  Host + device: the CPU tests call
// the host build of exactly this function.
__host__ __device__ __forceinline__ void rdft16_x2(const float (&x)[16], float (&yr)[9], float (&yi)[9]) {
    const float R = 0.70710678118654752f, C1 = 0.92387953251128674f, S1 = 0.38268343236508977f;
    // FFT8 of z: even part (z0 z2 z4 z6), odd part (z1 z3 z5 z7)
    float er[4], ei[4], orr[4], oi[4];
    {
        const float ar = x[0] + x[8], ai = x[1] + x[9], br = x[0] - x[8], bi = x[1] - x[9];          // z0 +- z4
        const float cr = x[4] + x[12], ci = x[5] + x[13], dr = x[4] - x[12], di = x[5] - x[13];      // z2 +- z6
        er[0] = ar + cr; ei[0] = ai + ci; er[2] = ar - cr; ei[2] = ai - ci;
        er[1] = br + di; ei[1] = bi - dr; er[3] = br - di; ei[3] = bi + dr;                          // -i (z2 - z6) = (di, -dr)
    }
    {
        const float ar = x[2] + x[10], ai = x[3] + x[11], br = x[2] - x[10], bi = x[3] - x[11];      // z1 +- z5
        const float cr = x[6] + x[14], ci = x[7] + x[15], dr = x[6] - x[14], di = x[7] - x[15];      // z3 +- z7
        orr[0] = ar + cr; oi[0] = ai + ci; orr[2] = ar - cr; oi[2] = ai - ci;
        orr[1] = br + di; oi[1] = bi - dr; orr[3] = br - di; oi[3] = bi + dr;
    }
    // Z[k] = E[k] + w8^k O[k], Z[k+4] = E[k] - w8^k O[k];  w8 = (1 - i) / sqrt 2, w8^2 = -i, w8^3 = (-1 - i) / sqrt 2
    float zr[8], zi[8];
    {
        float tr = orr[0], ti = oi[0];
        zr[0] = er[0] + tr; zi[0] = ei[0] + ti; zr[4] = er[0] - tr; zi[4] = ei[0] - ti;
        tr = (orr[1] + oi[1]) * R; ti = (oi[1] - orr[1]) * R;
        zr[1] = er[1] + tr; zi[1] = ei[1] + ti; zr[5] = er[1] - tr; zi[5] = ei[1] - ti;
        tr = oi[2]; ti = -orr[2];
        zr[2] = er[2] + tr; zi[2] = ei[2] + ti; zr[6] = er[2] - tr; zi[6] = ei[2] - ti;
        tr = (oi[3] - orr[3]) * R; ti = -(orr[3] + oi[3]) * R;
        zr[3] = er[3] + tr; zi[3] = ei[3] + ti; zr[7] = er[3] - tr; zi[7] = ei[3] - ti;
    }
    // real-input split: 2 Y[k] = S + T, S = Z[k] + conj(Z[8-k]), T = -i w16^k (Z[k] - conj(Z[8-k]))
    yr[0] = 2.f * (zr[0] + zi[0]); yi[0] = 0.f;
    yr[8] = 2.f * (zr[0] - zi[0]); yi[8] = 0.f;
    yr[4] = 2.f * zr[4]; yi[4] = -2.f * zi[4];
    const float cs[3] = {C1, R, S1}, sn[3] = {S1, R, C1};                  // cos, sin of 2 pi k / 16, k = 1, 2, 3
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k <= 3; ++k) {
        const float sr = zr[k] + zr[8 - k], si = zi[k] - zi[8 - k];        // S = Z[k] + conj(Z[8-k])
        const float dr = zr[k] - zr[8 - k], di = zi[k] + zi[8 - k];        // D = Z[k] - conj(Z[8-k])
        const float p1 = sn[k - 1] * dr - cs[k - 1] * di, p2 = sn[k - 1] * di + cs[k - 1] * dr;
        yr[k] = sr - p1; yi[k] = si - p2;                                  // T_k = -p1 - i p2
        yr[8 - k] = sr + p1; yi[8 - k] = -si - p2;                         // S_(8-k) = conj(S), T_(8-k) = p1 - i p2
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Host tables.
//   B operand of block b, piece hi / lo: fp16 [kgroup 8][n 64][8]  (K-major canonical, no swizzle: element (n, k) at
//   ((k / 8) * 64 + n) * 8 + k % 8).  K-group g holds the inputs n2 = 4 g .. 4 g + 3, first and second input half of the
//   block.  K-step ks of the MMAs uses the groups ks and 4 + ks: the two producer lanes of a frame row each own one 32-byte
//   sector of every sample row (groups 0-3 and 4-7).  Element order inside the 16-byte group: tcd_kslot(j, second, g) --
//   two 8-byte halves, one per sample pair (j >> 1), swapped for the groups of the second lane (g >= 4) so that the two
//   lanes store to different bank halves.
//   Output column c of block b (accumulator column 64 b + c):
//     block 0: [Re X[16m] | Im X[16m] (m = 0 carries X[256]) | Re X[16m+8] | Im X[16m+8]],   m = c % 16
//     block r: [Re X[16m+r] | Im X[16m+r] | Re X[16m+16-r] | Im X[16m+16-r]]
struct TcdHostTables {
    std::vector<__half> b_hi, b_lo;            // [8][8][64][8] each
    std::vector<float4> etab;                  // [257]: entry 16 c + m of chunk c = 2 b + h (bin in tcd_chunk_bin), [256] = bin 256
    std::vector<float> dct;                    // [TCD_MAX_OUT][24]
};

// position (0..7) inside K-group g of input n2 = 4 g + j, first (second = 0) or second (second = 1) input half of the block
__host__ __device__ constexpr int tcd_kslot(int j, int second, int g) { return 4 * ((j >> 1) ^ (g >> 2)) + 2 * second + (j & 1); }

// bin held by element m of 32-column chunk c = 2 b + h (16 re columns then 16 im columns)
__host__ __device__ constexpr int tcd_chunk_bin_c(int c, int m) {
    return (c >> 1) == 0 ? ((c & 1) == 0 ? 16 * m : 16 * m + 8) : ((c & 1) == 0 ? 16 * m + (c >> 1) : 16 * m + 16 - (c >> 1));
}
static inline int tcd_chunk_bin(int c, int m) {
    const int b = c >> 1, h = c & 1;
    if (b == 0) return h == 0 ? 16 * m : 16 * m + 8;
    return h == 0 ? 16 * m + b : 16 * m + 16 - b;
}

static inline void tcd_build_b(std::vector<__half>& b_hi, std::vector<__half>& b_lo) {
    b_hi.assign((size_t)TCD_BLOCKS * 8 * 64 * 8, __float2half_rn(0.f));
    b_lo = b_hi;
    const double PI2 = 6.283185307179586476925286766559;
    for (int b = 0; b < TCD_BLOCKS; ++b)
        for (int k = 0; k < 64; ++k)
            for (int n = 0; n < 64; ++n) {
                const int g = k >> 3, e = k & 7;                                             // which input this row multiplies:
                const int second = (e >> 1) & 1, n2 = 4 * g + 2 * ((e >> 2) ^ (g >> 2)) + (e & 1);  // inverse of tcd_kslot
                const int quarter = n >> 4, m = n & 15;
                double v = 0.0;
                if (b == 0) {
                    // first half: Y_0 -> X[16 m] (quarters 0, 1); second half: Y_8 -> X[16 m + 8] (quarters 2, 3)
                    if (!second && quarter < 2) {
                        const double a = PI2 * n2 * m / 32.0;
                        v = quarter == 0 ? cos(a) : -sin(a);
                        if (quarter == 1 && m == 0) v = (n2 & 1) ? -1.0 : 1.0;             // slot Im X[0] := X[256]
                    } else if (second && quarter >= 2) {
                        const double a = PI2 * n2 * 8 / 512.0 + PI2 * n2 * m / 32.0;
                        v = quarter == 2 ? cos(a) : -sin(a);
                    }
                } else {
                    const int r = quarter < 2 ? b : 16 - b;
                    const double a = PI2 * n2 * r / 512.0 + PI2 * n2 * m / 32.0;
                    const double tr = cos(a), ti = -sin(a);                                 // T = w512^(n2 r) w32^(n2 m)
                    // quarters 0, 1: (a + i b) T -> re: a tr - b ti, im: a ti + b tr
                    // quarters 2, 3: (a - i b) T -> re: a tr + b ti, im: a ti - b tr
                    const bool im_out = quarter & 1;
                    if (quarter < 2) v = !second ? (im_out ? ti : tr) : (im_out ? tr : -ti);
                    else v = !second ? (im_out ? ti : tr) : (im_out ? -tr : ti);
                }
                const __half hi = __float2half_rn((float)v);
                const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
                const size_t o = (((size_t)b * 8 + g) * 64 + n) * 8 + e;
                b_hi[o] = hi; b_lo[o] = lo;
            }
}

// wrise / wfall / grid as built by api.cu (build_mel); pscale turns (re^2 + im^2) of the scaled accumulators into power / n_fft
static inline void tcd_build_etab(std::vector<float4>& etab, const std::vector<float>& wrise, const std::vector<float>& wfall,
                                  const std::vector<int>& grid, int n_filt, float pscale) {
    etab.assign(257, make_float4(0.f, 0.f, 0.f, 0.f));
    auto entry = [&](int k) {
        int s = 0;
        float wr = 0.f, wf = 0.f;
        if (k >= grid[0] && k < grid[n_filt + 1]) {
            while (s < n_filt && k >= grid[s + 1]) ++s;                 // segment: grid[s] <= k < grid[s + 1]
            if (s < n_filt) wr = wrise[k];                              // rising edge of filter s
            if (s > 0) wf = wfall[k];                                   // falling edge of filter s - 1
        }
        float sf;
        memcpy(&sf, &s, 4);
        return make_float4(wr * pscale, wf * pscale, sf, 0.f);
    };
    for (int c = 0; c < 16; ++c)
        for (int m = 0; m < 16; ++m) etab[16 * c + m] = entry(tcd_chunk_bin(c, m));
    etab[256] = entry(256);
}

// CPU model of the tensor-core path for ONE frame of 512 int16 samples: the same butterfly, the same tables read through the
// same layout arithmetic, fp16 products accumulated in fp32, the frame's first sample removed from Y_0 and restored in X[0].
// d[512] = the frame's accumulator row (TMEM lane) as the epilogue sees it after that correction.
static inline void tcd_host_accumulators(const int16_t* x, float* d) {
    static std::vector<__half> b_hi, b_lo;
    if (b_hi.empty()) tcd_build_b(b_hi, b_lo);
    std::vector<float> a((size_t)TCD_BLOCKS * 64);
    for (int g = 0; g < 8; ++g)
        for (int j = 0; j < 4; ++j) {
            const int n2 = 4 * g + j;
            float in[16], yr[9], yi[9];
            for (int q = 0; q < 16; ++q) in[q] = (float)x[n2 + 32 * q] * TCD_IN_SCALE;
            rdft16_x2(in, yr, yi);
            a[0 * 64 + 8 * g + tcd_kslot(j, 0, g)] = yr[0] - TCD_X0_Y * (float)x[0]; a[0 * 64 + 8 * g + tcd_kslot(j, 1, g)] = yr[8];
            for (int r = 1; r < 8; ++r) { a[r * 64 + 8 * g + tcd_kslot(j, 0, g)] = yr[r]; a[r * 64 + 8 * g + tcd_kslot(j, 1, g)] = yi[r]; }
        }
    for (int b = 0; b < TCD_BLOCKS; ++b)
        for (int n = 0; n < 64; ++n) {
            float acc = 0.f;
            for (int pass = 0; pass < 3; ++pass)
                for (int k = 0; k < 64; ++k) {
                    const float av = a[b * 64 + k];
                    const __half ah = __float2half_rn(av);
                    const __half al = __float2half_rn(av - __half2float(ah));
                    const size_t o = (((size_t)b * 8 + (k >> 3)) * 64 + n) * 8 + (k & 7);
                    const float pa = __half2float(pass == 0 ? al : ah), pb = __half2float(pass == 1 ? b_lo[o] : b_hi[o]);
                    acc += pa * pb;
                }
            d[64 * b + n] = acc;
        }
    d[0] += TCD_X0_D * (float)x[0];            // the kernel transforms x - x[0] (exact zeros for constant input) and restores X[0] here
}

// |X[k]|^2 of the raw samples, k = 0..256
static inline void tcd_host_power(const int16_t* x, double* power) {
    float d[512];
    tcd_host_accumulators(x, d);
    const double inv = 1.0 / ((double)TCD_A_SCALE * (double)TCD_A_SCALE);
    for (int c = 0; c < 16; ++c)
        for (int m = 0; m < 16; ++m) {
            const double re = d[32 * c + m], im = d[32 * c + 16 + m];
            const int k = tcd_chunk_bin(c, m);
            if (c == 0 && m == 0) { power[0] = re * re * inv; power[256] = im * im * inv; }
            else power[k] = (re * re + im * im) * inv;
        }
}

// The epilogue of mfcc_tc_stream_kernel for one accumulator row, statement for statement (fp32): table-driven mel sums with
// the run-length flush, log, DCT, c0.  acc: (n_filt + 2) floats of scratch.
static inline void tcd_host_epilogue(const float* d, const float4* etab, const float* dct, int n_filt, int n_out, float tot_scale,
                                     float* acc, float* out) {
    for (int j = 0; j < n_filt + 2; ++j) acc[j] = 0.f;
    float tot = 0.f, a_r = 0.f, a_f = 0.f, p256 = 0.f;
    int s_cur = 0;
    for (int c = 0; c < 16; ++c)
        for (int m = 0; m < 16; ++m) {
            const float re = d[32 * c + m], im = d[32 * c + 16 + m];
            float p = re * re;
            if (c == 0 && m == 0) p256 = im * im;
            else p = fmaf(im, im, p);
            const float4 e = etab[16 * c + m];
            int s;
            memcpy(&s, &e.z, 4);
            if (s != s_cur) { acc[s_cur + 1] += a_r; acc[s_cur] += a_f; s_cur = s; a_r = 0.f; a_f = 0.f; }
            tot += p;
            a_r = fmaf(e.x, p, a_r);
            a_f = fmaf(e.y, p, a_f);
        }
    {
        const float4 e = etab[256];
        int s;
        memcpy(&s, &e.z, 4);
        if (s != s_cur) { acc[s_cur + 1] += a_r; acc[s_cur] += a_f; s_cur = s; a_r = 0.f; a_f = 0.f; }
        tot += p256;
        a_r = fmaf(e.x, p256, a_r);
        a_f = fmaf(e.y, p256, a_f);
        acc[s_cur + 1] += a_r; acc[s_cur] += a_f;
    }
    const float eps = 2.220446049250313e-16f;
    for (int j = 0; j < n_filt; ++j) acc[j + 1] = logf(fmaxf(acc[j + 1], eps));
    for (int o = 0; o < n_out; ++o) {
        float v0 = 0.f, v1 = 0.f;
        const float* dr = dct + (size_t)o * 24;
        int j = 0;
        for (; j + 1 < n_filt; j += 2) { v0 = fmaf(dr[j], acc[j + 1], v0); v1 = fmaf(dr[j + 1], acc[j + 2], v1); }
        if (j < n_filt) v0 = fmaf(dr[j], acc[j + 1], v0);
        out[o] = o == 0 ? logf(fmaxf(tot * tot_scale, eps)) : v0 + v1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// instruction descriptor: kind::f16, A and B fp16 K-major, fp32 accumulate, M = 128
__device__ __forceinline__ uint32_t tcd_idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tcd_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}

// int16 pair (already XORed with 0x80008000) -> two floats scaled by 2^-6 (TCD_IN_SCALE), exact: drop the biased 16-bit value
// into the mantissa of 2^17 (ulp 2^-6) and subtract 2^17 + 32768 * 2^-6.
__device__ __forceinline__ void tcd_cvt2(uint32_t v, float& lo, float& hi) {
    lo = __uint_as_float(__byte_perm(v, 0x48000000u, 0x7610)) - 131584.f;
    hi = __uint_as_float(__byte_perm(v, 0x48000000u, 0x7632)) - 131584.f;
}

// four floats -> fp16 hi and lo pieces, one 8-byte store each
__device__ __forceinline__ void tcd_put4(__half* hi_dst, __half* lo_dst, float v0, float v1, float v2, float v3) {
    const __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v0 - f0.x, v1 - f0.y), l1 = __floats2half2_rn(v2 - f1.x, v3 - f1.y);
    *reinterpret_cast<uint2*>(hi_dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(lo_dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

}  // namespace pb
