// mfcc_tc.cuh -- EXPERIMENTAL K1 variant: the 512-point real DFT on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Status: written at the end of round 1 after the GPU budget was spent.  The host-side parts (twiddle tables in the UMMA
// operand layout, accumulator-column -> bin map, the 16-point real-DFT butterfly, the mel / DCT tables and the epilogue's
// arithmetic) are verified on the CPU (pb_debug_tc_dft_power, pb_debug_tc_mfcc_frame, tests/test_tc_dft_host_model.py); the
// kernel itself compiles for sm_100a but has NOT run on hardware yet.  It is
// therefore opt-in only (pb_debug_k1_mode(h, 1)); the default MFCC kernels are untouched.  Design and numbers: DESIGN.md
// section 6 ("Round-2 plan for K1"), numerical study: scripts/proto_tc_dft.py.
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
//
// Mapping.  A CTA owns groups of 128 streams; the frames a tick releases in a group are processed in tiles of 128 frames =
// the 128 TMEM lanes.  Warps 0-3: epilogue, thread <-> frame (TMEM lane): tcgen05.ld -> power -> mel (table driven) -> log
// -> DCT -> ring row.  Warps 4-11: producers, thread <-> (frame, four n2): 16 x LDG.64 of PCM, four real DFT-16s, fp16
// split, sixteen 16-byte stores into the K-major canonical A tiles.  Warp 12: one lane issues the 96 MMAs of a tile
// (4 K-steps x 8 blocks x 3 passes, M = 128, N = 64, K = 16).  Hand-offs are mbarriers only inside a group.
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
constexpr int TCD_EPI_WARPS = 4, TCD_PROD_WARPS = 8;
constexpr int TCD_THREADS = (TCD_EPI_WARPS + TCD_PROD_WARPS + 1) * 32;     // 416
constexpr int TCD_GROUP = 128;                // streams per group
constexpr int TCD_MAX_NEW = 4;                // frames a stream may release per tick
constexpr int TCD_MAX_FILT = 22;              // n_filt + 2 accumulator slots of 128 floats must fit
constexpr int TCD_MAX_OUT = 16;
constexpr float TCD_IN_SCALE = 0.03125f;      // 2^-5 folded into the int16 -> float conversion; the butterfly returns 2 Y
constexpr float TCD_A_SCALE = 0.0625f;        // => A operands hold Y * 2^-4: |A| <= 32768 < fp16 max

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
//   ((k / 8) * 64 + n) * 8 + k % 8).  k = 8 g + e: e < 4 -> first input half, n2 = 4 g + e; e >= 4 -> second half,
//   n2 = 4 g + e - 4  (the producer task (frame, g) writes exactly one 16-byte K-group per block).
//   Output column c of block b (accumulator column 64 b + c):
//     block 0: [Re X[16m] | Im X[16m] (m = 0 carries X[256]) | Re X[16m+8] | Im X[16m+8]],   m = c % 16
//     block r: [Re X[16m+r] | Im X[16m+r] | Re X[16m+16-r] | Im X[16m+16-r]]
struct TcdHostTables {
    std::vector<__half> b_hi, b_lo;            // [8][8][64][8] each
    std::vector<float4> etab;                  // [257]: entry 16 c + m of chunk c = 2 b + h (bin in tcd_chunk_bin), [256] = bin 256
    std::vector<float> dct;                    // [TCD_MAX_OUT][24]
};

// bin held by element m of 32-column chunk c = 2 b + h (16 re columns then 16 im columns)
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
                const int g = k >> 3, e = k & 7, n2 = 4 * g + (e & 3), second = e >> 2;      // which input this row multiplies
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
// same layout arithmetic, fp16 products accumulated in fp32.  d[512] = the frame's accumulator row (TMEM lane) as the
// epilogue sees it.
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
            a[0 * 64 + 8 * g + j] = yr[0]; a[0 * 64 + 8 * g + 4 + j] = yr[8];
            for (int r = 1; r < 8; ++r) { a[r * 64 + 8 * g + j] = yr[r]; a[r * 64 + 8 * g + 4 + j] = yi[r]; }
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
struct TcdTables {               // device pointers
    const uint4* b;              // [2][8][8][64] x 16 bytes: hi then lo
    const float4* etab;          // [257]
    const float* dct;            // [TCD_MAX_OUT][24]
    int n_filt, n_out;
    float tot_scale;             // pscale for the total power (c0)
};

struct TcdSmem {
    __half b_hi[TCD_BLOCKS][8][64][8];
    __half b_lo[TCD_BLOCKS][8][64][8];
    __half a_hi[TCD_BLOCKS][2][128][8];          // the current K-step: two 16-byte K-groups per block
    __half a_lo[TCD_BLOCKS][2][128][8];
    float acc[TCD_MAX_FILT + 2][128];            // mel accumulators, slot j + 1 = filter j, own column per epilogue thread
    float4 etab[257];
    float dct[TCD_MAX_OUT][24];
    long long st_n0[TCD_GROUP], st_c0[TCD_GROUP], st_ts0[TCD_GROUP];
    int st_id[TCD_GROUP], st_cnt[TCD_GROUP];
    short fr_stream[TCD_GROUP * TCD_MAX_NEW], fr_sub[TCD_GROUP * TCD_MAX_NEW];
    int warp_tot[4];
    int n_frames;
    unsigned long long a_full, a_empty, d_full, d_empty;
    uint32_t tmem_base;
};

// instruction descriptor: kind::f16, A and B fp16 K-major, fp32 accumulate, M = 128
__device__ __forceinline__ uint32_t tcd_idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tcd_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}

// int16 pair (already XORed with 0x80008000) -> two floats scaled by 2^-5, exact: drop the biased 16-bit value into the
// mantissa of 2^18 (ulp 2^-5) and subtract 2^18 + 32768 * 2^-5.
__device__ __forceinline__ void tcd_cvt2(uint32_t v, float& lo, float& hi) {
    lo = __uint_as_float(__byte_perm(v, 0x48800000u, 0x7610)) - 263168.f;
    hi = __uint_as_float(__byte_perm(v, 0x48800000u, 0x7632)) - 263168.f;
}

// eight floats -> fp16 hi and lo pieces, one 16-byte store each
__device__ __forceinline__ void tcd_put8(__half* hi_dst, __half* lo_dst, const float (&v)[8]) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 h = __floats2half2_rn(v[2 * p], v[2 * p + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * p] - hf.x, v[2 * p + 1] - hf.y);
        hw[p] = *reinterpret_cast<const uint32_t*>(&h);
        lw[p] = *reinterpret_cast<const uint32_t*>(&l);
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

__global__ void __launch_bounds__(TCD_THREADS, 1)
mfcc_tc_stream_kernel(const int16_t* __restrict__ pcm, const int* __restrict__ ids, int n, int chunk, int hop,
                      TcdTables tab, StreamState st) {
    extern __shared__ __align__(128) unsigned char tcd_raw[];
    TcdSmem& sm = *reinterpret_cast<TcdSmem*>(tcd_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int used = 512;

    // ---- one-time setup: twiddle operands and small tables to shared memory, barriers, TMEM (all 512 columns)
    {
        uint4* dst = reinterpret_cast<uint4*>(&sm.b_hi[0][0][0][0]);
        for (int e = tid; e < 2 * TCD_BLOCKS * 8 * 64; e += TCD_THREADS) dst[e] = __ldg(tab.b + e);     // b_hi then b_lo, contiguous
        for (int e = tid; e < 257; e += TCD_THREADS) sm.etab[e] = __ldg(tab.etab + e);
        for (int e = tid; e < TCD_MAX_OUT * 24; e += TCD_THREADS) (&sm.dct[0][0])[e] = __ldg(tab.dct + e);
    }
    if (tid == 0) {
        mbar_init(&sm.a_full, TCD_PROD_WARPS * 32); mbar_init(&sm.a_empty, 1);
        mbar_init(&sm.d_full, 1); mbar_init(&sm.d_empty, TCD_EPI_WARPS * 32);
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
    const uint32_t idesc = tcd_idesc(64);

    uint32_t n_ksteps = 0;            // K-steps handed over so far (producers, issuer): phase of a_full / a_empty
    uint32_t n_tiles_done = 0;        // tiles so far (issuer, epilogue): phase of d_full / d_empty

    const int n_groups = (n + TCD_GROUP - 1) / TCD_GROUP;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int base = grp * TCD_GROUP;
        // ---- bookkeeping: thread t < 128 <-> stream base + t; frame list by a block-wide exclusive scan
        int cnt = 0;
        if (tid < TCD_GROUP) {
            const int i = base + tid;
            int sid = -1;
            long long n0 = 0, c0 = 0, ts0 = 0;
            if (i < n) {
                sid = ids ? ids[i] : i;
                n0 = st.n_samples[sid];
                c0 = frames_ready(n0, used, hop);
                cnt = (int)(frames_ready(n0 + chunk, used, hop) - c0);
                ts0 = c0 * hop < n0 ? c0 * hop : n0;
            }
            sm.st_id[tid] = sid; sm.st_n0[tid] = n0; sm.st_c0[tid] = c0; sm.st_ts0[tid] = ts0; sm.st_cnt[tid] = cnt;
        }
        int incl = cnt;
        if (warp < 4) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
            if (lane == 31) sm.warp_tot[warp] = incl;
        }
        __syncthreads();
        if (tid < TCD_GROUP) {
            int off = incl - cnt;
            for (int w = 0; w < warp; ++w) off += sm.warp_tot[w];
            for (int j = 0; j < cnt; ++j) { sm.fr_stream[off + j] = (short)tid; sm.fr_sub[off + j] = (short)j; }
            if (tid == TCD_GROUP - 1) sm.n_frames = off + cnt;
        }
        __syncthreads();
        const int n_frames = sm.n_frames;
        const int n_tiles = (n_frames + 127) >> 7;

        for (int tile = 0; tile < n_tiles; ++tile) {
            if (warp >= TCD_EPI_WARPS && warp < TCD_EPI_WARPS + TCD_PROD_WARPS) {
                // ================= producers: thread <-> (frame row, K-group parity)
                const int pw = warp - TCD_EPI_WARPS, row = 32 * (pw & 3) + lane, gq = pw >> 2;
                const int f = tile * 128 + row;
                const bool active = f < n_frames;
                const int16_t *p0 = pcm, *p1 = pcm;
                int len0 = 0;
                if (active) {
                    const int t = sm.fr_stream[f];
                    const long long a0 = (sm.st_c0[t] + sm.fr_sub[f]) * hop, n0 = sm.st_n0[t];
                    const int16_t* chunk_p = pcm + (long long)(base + t) * chunk;
                    if (a0 >= n0) { len0 = 0; p1 = chunk_p + (a0 - n0); }
                    else {
                        len0 = (int)min((long long)used, n0 - a0);
                        p0 = st.tail + (long long)sm.st_id[t] * st.tail_cap + (a0 - sm.st_ts0[t]);
                        p1 = chunk_p;
                    }
                }
#pragma unroll 1
                for (int ks = 0; ks < TCD_KSTEPS; ++ks, ++n_ksteps) {
                    const int g = 2 * ks + gq;                     // K-group: n2 = 4 g .. 4 g + 3
                    float y[TCD_BLOCKS][8];                        // per block: the 8 values of this K-group
                    if (active) {
                        uint2 raw[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const int i = 4 * g + 32 * q;
                            const int16_t* src = i < len0 ? p0 + i : p1 + (i - len0);
                            raw[q] = __ldg(reinterpret_cast<const uint2*>(src));
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float x[16], yr[9], yi[9];
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const uint32_t w = ((j & 2) ? raw[q].y : raw[q].x) ^ 0x80008000u;
                                float a, b;
                                tcd_cvt2(w, a, b);
                                x[q] = (j & 1) ? b : a;
                            }
                            rdft16_x2(x, yr, yi);
                            y[0][j] = yr[0]; y[0][4 + j] = yr[8];
#pragma unroll
                            for (int r = 1; r < 8; ++r) { y[r][j] = yr[r]; y[r][4 + j] = yi[r]; }
                        }
                    }
                    // the tensor core has finished reading the previous K-step's tiles
                    mbar_wait(&sm.a_empty, (n_ksteps & 1) ^ 1);
                    if (active) {
#pragma unroll
                        for (int b = 0; b < TCD_BLOCKS; ++b) tcd_put8(&sm.a_hi[b][gq][row][0], &sm.a_lo[b][gq][row][0], y[b]);
                    }
                    fence_proxy_async();
                    mbar_arrive(&sm.a_full);
                }
            } else if (warp == TCD_EPI_WARPS + TCD_PROD_WARPS) {
                // ================= MMA issuer (one lane)
                if (lane == 0) {
                    mbar_wait(&sm.d_empty, (n_tiles_done & 1) ^ 1);          // the epilogue has drained the previous tile
                    tc5_fence_after();
#pragma unroll 1
                    for (int ks = 0; ks < TCD_KSTEPS; ++ks, ++n_ksteps) {
                        mbar_wait(&sm.a_full, n_ksteps & 1);
                        tc5_fence_after();
#pragma unroll 1
                        for (int b = 0; b < TCD_BLOCKS; ++b) {
                            const uint64_t dah = tc5_desc(&sm.a_hi[b][0][0][0], 2048, 128), dal = tc5_desc(&sm.a_lo[b][0][0][0], 2048, 128);
                            const uint64_t dbh = tc5_desc(&sm.b_hi[b][2 * ks][0][0], 1024, 128), dbl = tc5_desc(&sm.b_lo[b][2 * ks][0][0], 1024, 128);
                            const uint32_t d = tmem + 64 * b;
                            tcd_mma(d, dal, dbh, idesc, ks > 0);
                            tcd_mma(d, dah, dbl, idesc, 1);
                            tcd_mma(d, dah, dbh, idesc, 1);
                        }
                        tc5_commit(&sm.a_empty);                             // arrives when these MMAs have read the A tiles
                    }
                    tc5_commit(&sm.d_full);
                } else {
                    n_ksteps += TCD_KSTEPS;
                }
                __syncwarp();
                ++n_tiles_done;
            } else {
                // ================= epilogue: thread <-> frame = TMEM lane
                const int f = tile * 128 + tid;
                const bool active = f < n_frames;
                const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
                float* acc = &sm.acc[0][tid];
                for (int j = 0; j < tab.n_filt + 2; ++j) acc[j * 128] = 0.f;
                mbar_wait(&sm.d_full, n_tiles_done & 1);
                tc5_fence_after();
                float tot = 0.f, a_r = 0.f, a_f = 0.f, p256 = 0.f;
                int s_cur = 0;
#pragma unroll 1
                for (int c = 0; c < 16; ++c) {
                    float re[16], im[16];
                    tc5_ld16(t_row + 32 * c, re);
                    tc5_ld16(t_row + 32 * c + 16, im);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        float p = re[m] * re[m];
                        if (c == 0 && m == 0) p256 = im[0] * im[0];
                        else p = fmaf(im[m], im[m], p);
                        const float4 e = sm.etab[16 * c + m];
                        const int s = __float_as_int(e.z);
                        if (s != s_cur) {                                        // warp-uniform: the table is shared
                            acc[(s_cur + 1) * 128] += a_r; acc[s_cur * 128] += a_f;
                            s_cur = s; a_r = 0.f; a_f = 0.f;
                        }
                        tot += p;
                        a_r = fmaf(e.x, p, a_r);
                        a_f = fmaf(e.y, p, a_f);
                    }
                }
                {   // bin 256 (carried in the Im X[0] slot)
                    const float4 e = sm.etab[256];
                    const int s = __float_as_int(e.z);
                    if (s != s_cur) { acc[(s_cur + 1) * 128] += a_r; acc[s_cur * 128] += a_f; s_cur = s; a_r = 0.f; a_f = 0.f; }
                    tot += p256;
                    a_r = fmaf(e.x, p256, a_r);
                    a_f = fmaf(e.y, p256, a_f);
                    acc[(s_cur + 1) * 128] += a_r; acc[s_cur * 128] += a_f;
                }
                tc5_fence_before();
                mbar_arrive(&sm.d_empty);                                        // TMEM may be overwritten by the next tile
                ++n_tiles_done;
                if (active) {
                    const int t = sm.fr_stream[f];
                    const long long k = sm.st_c0[t] + sm.fr_sub[f];
                    float* row = st.ring + ((long long)sm.st_id[t] * st.ring_rows + (int)(k % st.ring_rows)) * st.row_stride;
                    for (int j = 0; j < tab.n_filt; ++j) acc[(j + 1) * 128] = logf(fmaxf(acc[(j + 1) * 128], K1_EPS));
                    for (int o = 0; o < tab.n_out; ++o) {
                        float v0 = 0.f, v1 = 0.f;
                        const float* d = sm.dct[o];
                        int j = 0;
                        for (; j + 1 < tab.n_filt; j += 2) { v0 = fmaf(d[j], acc[(j + 1) * 128], v0); v1 = fmaf(d[j + 1], acc[(j + 2) * 128], v1); }
                        if (j < tab.n_filt) v0 = fmaf(d[j], acc[(j + 1) * 128], v0);
                        row[o] = o == 0 ? logf(fmaxf(tot * tab.tot_scale, K1_EPS)) : v0 + v1;
                    }
                }
            }
        }
        __syncthreads();              // every frame of the group is done: all reads of the old tails are complete
        // ---- tails and sample counters (chunk >= 512 samples: nothing of the old tail survives)
        if (tid < TCD_GROUP && sm.st_id[tid] >= 0) {
            const long long n0 = sm.st_n0[tid], n1 = n0 + chunk;
            const long long c1 = sm.st_c0[tid] + sm.st_cnt[tid];
            const long long ts1 = c1 * hop < n1 ? c1 * hop : n1;
            sm.st_ts0[tid] = ts1 - n0;                                           // offset of the new tail inside the chunk
            sm.st_cnt[tid] = (int)(n1 - ts1) >> 3;                               // 16-byte vectors
            st.n_samples[sm.st_id[tid]] = n1;
        }
        __syncthreads();
        for (int e = tid; e < TCD_GROUP * 64; e += TCD_THREADS) {
            const int t = e >> 6, vi = e & 63;
            if (sm.st_id[t] >= 0 && vi < sm.st_cnt[t]) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(pcm + (long long)(base + t) * chunk + sm.st_ts0[t]) + vi);
                reinterpret_cast<int4*>(st.tail + (long long)sm.st_id[t] * st.tail_cap)[vi] = v;
            }
        }
        __syncthreads();
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace pb
