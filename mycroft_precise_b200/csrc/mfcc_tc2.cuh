// mfcc_tc2.cuh -- K1 on the 5th-generation tensor cores (tcgen05 + TMEM): the stateful MFCC tick for the reference's
// default geometry (n_fft 512, 20 mel filters at 16 kHz; precise/params.py:140-144).
//
// Replaces, per frame, np.fft.rfft(frame, n=512) -> power -> mel filterbank -> log -> DCT -> c0 of sonopy.mfcc_spec as the
// reference calls it (precise/vectorization.py:36-39) and the carry-buffer bookkeeping of Listener.update_vectors
// (precise/network_runner.py:125-146).
//
// Arithmetic (shared with mfcc_tc.cuh, whose host tables and CPU model this kernel uses unchanged): n = n2 + 32 q,
// k = 16 m + r.  CUDA cores: the 16-point real DFT over q (fp32, rdft16_x2).  Tensor cores: eight 64 x 64 real GEMM blocks
// per frame, X[16 m + r] = sum_n2 Y_r[n2] w512^(n2 r) w32^(n2 m), operands split into fp16 hi + lo pieces, three passes
// (a_lo b_hi + a_hi b_lo + a_hi b_hi) accumulated in fp32 in TMEM.
//
// What changed against the first version (mfcc_tc_stream_kernel, 438 us per 131 072-stream tick on the B200):
//   * super-groups of up to 1024 streams per CTA share one frame list, so tiles of 128 frames are full (93 % instead of
//     64 % with 128-stream groups) and there is one bookkeeping pass per CTA instead of seven;
//   * a producer lane owns one 32-byte sector of every 64-byte sample row of its frame (the K-groups 4 gq + ks): 16-byte
//     loads, each serving two K-steps (the old mapping, 32 frames x 8 bytes per instruction, fetched every sector four
//     times and visited every L1 line sixteen times); the next half's loads are issued as soon as the registers are free;
//   * the producer warp that delivers last issues the K-step's MMAs (descriptor arithmetic is one add per operand);
//   * the mel stage is compiled for the geometry: each of the 512 accumulator columns knows its bin, segment and edge
//     weights at compile time, so the 257 bins cost five FP32 instructions each into register accumulators (the old
//     epilogue did a table look-up and a dependent shared-memory read-modify-write per bin);
//   * new tails are copied by the producer warps at the end of each tile (four streams' loads in flight per warp), the
//     sample counters are bumped by the bookkeeping pass;
//   * exact zeros: the frame's first sample is subtracted from Y_0 (a constant only moves X[0], which is corrected in the
//     epilogue), so a constant input gives exactly zero in every other bin, as the float64 reference and the FFT kernels do.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "mfcc_tc.cuh"

namespace pb {

constexpr int TC2_EPI_WARPS = 4, TC2_PROD_WARPS = 8;
constexpr int TC2_THREADS = (TC2_EPI_WARPS + TC2_PROD_WARPS) * 32;    // 384: three warps per SM sub-partition -> up to 168 registers
constexpr int TC2_ISSUER_TID = TC2_EPI_WARPS * 32;                    // lane 0 of the first producer warp also issues the MMAs
constexpr int TC2_SG_MAX = 1024;                // streams per super-group
constexpr int TC2_MAX_NEW = 4;                  // frames a stream may complete per tick
constexpr int TC2_MAX_FRAMES = TC2_SG_MAX * TC2_MAX_NEW;

// ---------------------------------------------------------------------------------------------------------------------
// Compile-time mel geometry.  grid[] is sonopy's bin grid as api.cu builds it at run time (build_mel); pb_create enables
// this kernel only when the run-time grid and edge weights equal these (tc2_geo_matches).
struct Tc2Geo20 {
    static constexpr int n_filt = 20;
    static constexpr int n_bins = 257;
    __host__ __device__ static constexpr int grid(int i) {
        constexpr int g[22] = {0, 1, 3, 6, 9, 12, 16, 21, 26, 32, 39, 47, 57, 68, 81, 97, 114, 135, 159, 187, 219, 257};
        return g[i];
    }
};

template <class G>
__host__ __device__ constexpr bool tc2_in_grid(int k) { return k >= G::grid(0) && k < G::grid(G::n_filt + 1); }
// segment s: grid[s] <= k < grid[s + 1]; rising edge of filter s (s < n_filt), falling edge of filter s - 1 (s >= 1)
template <class G>
__host__ __device__ constexpr int tc2_seg(int k) {
    int s = 0;
    while (s < G::n_filt && k >= G::grid(s + 1)) ++s;
    return s;
}
// the same expressions as build_mel (api.cu): np.linspace(0, 1, n, endpoint=False) / np.linspace(1, 0, n, endpoint=False)
template <class G>
__host__ __device__ constexpr float tc2_wrise(int k) {
    const int s = tc2_seg<G>(k);
    if (!tc2_in_grid<G>(k) || s >= G::n_filt) return 0.f;
    const int lo = G::grid(s), mid = G::grid(s + 1);
    return (float)((double)(k - lo) * (1.0 / (double)(mid - lo)));
}
template <class G>
__host__ __device__ constexpr float tc2_wfall(int k) {
    const int s = tc2_seg<G>(k);
    if (!tc2_in_grid<G>(k) || s < 1) return 0.f;
    const int mid = G::grid(s), hi = G::grid(s + 1);
    return (float)((double)(k - mid) * (-1.0 / (double)(hi - mid)) + 1.0);
}
template <class G>
static inline bool tc2_geo_matches(int n_filt, int n_bins, const std::vector<int>& grid, const std::vector<float>& wrise,
                                   const std::vector<float>& wfall) {
    if (n_filt != G::n_filt || n_bins != G::n_bins || (int)grid.size() != n_filt + 2) return false;
    for (int i = 0; i < n_filt + 2; ++i) if (grid[i] != G::grid(i)) return false;
    for (int k = 0; k < n_bins; ++k)
        if (wrise[k] != tc2_wrise<G>(k) || wfall[k] != tc2_wfall<G>(k)) return false;
    return true;
}

template <int... I, class F>
__device__ __forceinline__ void tc2_static_for_impl(std::integer_sequence<int, I...>, F&& f) { (f(std::integral_constant<int, I>{}), ...); }
template <int N, class F>
__device__ __forceinline__ void tc2_static_for(F&& f) { tc2_static_for_impl(std::make_integer_sequence<int, N>{}, f); }

// ---------------------------------------------------------------------------------------------------------------------
struct Tc2Smem {
    __half b_hi[TCD_BLOCKS][8][64][8];           // twiddle operands, resident (64 KB + 64 KB)
    __half b_lo[TCD_BLOCKS][8][64][8];
    __half a_hi[TCD_BLOCKS][2][128][8];          // one K-step of A: two 16-byte K-groups per block (32 KB + 32 KB)
    __half a_lo[TCD_BLOCKS][2][128][8];
    float dct[TCD_MAX_OUT][24];
    float x0[2][128];                            // per tile parity: the constant subtracted from the frame (its first sample)
    int st_sid[TC2_SG_MAX];                      // stream id (only read when ids != null)
    short st_d[TC2_SG_MAX];                      // first new frame's start relative to the chunk: c0 * hop - n0
    unsigned short st_off[TC2_SG_MAX];           // index of the stream's first frame in the list
    unsigned char st_slot[TC2_SG_MAX];           // ring slot of the first new frame
    unsigned char st_cnt[TC2_SG_MAX];            // frames completed by this tick
    unsigned short fr[TC2_MAX_FRAMES];           // frame list: (local stream << 2) | sub-frame
    int lane_tot[32];
    int n_frames;
    unsigned int arrivals;                       // producer warps that have delivered their part of a K-step (monotonic)
    int b_loaded;
    unsigned long long a_empty, d_full, d_empty, b_ready;
    uint32_t tmem_base;
};

struct Tc2Tables {               // device pointers
    const uint4* b;              // [2][8][8][64] x 16 bytes: hi then lo (tcd_build_b)
    const float* dct;            // [TCD_MAX_OUT][24]
    int n_out;
    float pscale;                // (re^2 + im^2) of the scaled accumulators -> power / n_fft
};

// mbarrier wait for a role that is expected to wait long (the epilogue between tiles): back off between polls so the
// spinning threads do not take issue slots from the producers
__device__ __forceinline__ void tc2_mbar_wait_idle(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(128);
    }
}

__device__ __forceinline__ void tc2_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
// tcgen05.wait::ld, tied to the registers it makes valid so that no use can be scheduled above it
__device__ __forceinline__ void tc2_wait_ld(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

template <class G>
__global__ void __launch_bounds__(TC2_THREADS, 1)      

mfcc_tc2_stream_kernel(const int16_t* __restrict__ pcm, const int* __restrict__ ids, int n, int sg, int chunk, int hop,
                       Tc2Tables tab, StreamState st) {
    extern __shared__ __align__(128) unsigned char tc2_raw[];
    Tc2Smem& sm = *reinterpret_cast<Tc2Smem*>(tc2_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int used = 512;
    static_assert(G::n_filt % 4 == 0 && G::n_filt <= 24, "DCT rows are read as float4 from 24-float rows");

    // ---- one-time setup: barriers, TMEM (all 512 columns), twiddle operands by one bulk copy pair, DCT table
    if (tid == 0) {
        mbar_init(&sm.a_empty, 1);
        sm.arrivals = 0; sm.b_loaded = 0;
        mbar_init(&sm.d_full, 1); mbar_init(&sm.d_empty, TC2_EPI_WARPS * 32);
        mbar_init(&sm.b_ready, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < TCD_MAX_OUT * 24; e += TC2_THREADS) (&sm.dct[0][0])[e] = __ldg(tab.dct + e);
    fence_proxy_async();
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    if (tid == TC2_ISSUER_TID) {
        mbar_expect_tx(&sm.b_ready, 2u * 65536u);
        bulk_g2s(&sm.b_hi[0][0][0][0], tab.b, 65536u, &sm.b_ready);
        bulk_g2s(&sm.b_lo[0][0][0][0], reinterpret_cast<const char*>(tab.b) + 65536, 65536u, &sm.b_ready);
    }
    const uint32_t tmem = sm.tmem_base;
    const uint32_t idesc = tcd_idesc(64);

    uint32_t n_ksteps = 0;            // K-steps handed over so far (producers, issuer): phase of a_full / a_empty
    uint32_t n_tiles_done = 0;        // tiles so far (issuer, epilogue): phase of d_full / d_empty

    const int n_groups = (n + sg - 1) / sg;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int base = grp * sg;
        const int sg_n = min(sg, n - base);
        // ---- bookkeeping: per stream, the frames this tick completes (cf. Listener.update_vectors, network_runner.py:137-144)
        for (int s = tid; s < sg_n; s += TC2_THREADS) {
            const int sid = ids ? ids[base + s] : base + s;
            const long long n0 = st.n_samples[sid];
            const long long c0 = frames_ready(n0, used, hop);
            sm.st_sid[s] = sid;
            sm.st_cnt[s] = (unsigned char)(frames_ready(n0 + chunk, used, hop) - c0);
            sm.st_d[s] = (short)(c0 * hop - n0);
            sm.st_slot[s] = (unsigned char)(c0 % st.ring_rows);
            st.n_samples[sid] = n0 + chunk;          // nothing else in this kernel reads it; K2 runs after the kernel
        }
        __syncthreads();
        if (warp == 0) {              // exclusive scan of the counts: lane l owns streams 32 l .. 32 l + 31
            int tot = 0;
            for (int e = 0; e < 32; ++e) { const int s = 32 * lane + e; if (s < sg_n) tot += sm.st_cnt[s]; }
            int incl = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
            int off = incl - tot;
            for (int e = 0; e < 32; ++e) {
                const int s = 32 * lane + e;
                if (s < sg_n) { sm.st_off[s] = (unsigned short)off; off += sm.st_cnt[s]; }
            }
            if (lane == 31) sm.n_frames = incl;
        }
        __syncthreads();
        for (int s = tid; s < sg_n; s += TC2_THREADS) {
            const int off = sm.st_off[s], cnt = sm.st_cnt[s];
            for (int j = 0; j < cnt; ++j) sm.fr[off + j] = (unsigned short)((s << 2) | j);
        }
        __syncthreads();
        const int n_frames = sm.n_frames;
        const int n_tiles = (n_frames + 127) >> 7;

        if (warp >= TC2_EPI_WARPS) {
            // ================= producers: lane <-> (frame row, K-group parity); warp pw owns rows 16 pw .. 16 pw + 15
            // lane <-> (row, gq): a half-warp is 8 rows x both K-groups, so its 8-byte operand stores cover all 32 banks once
            const int pw = warp - TC2_EPI_WARPS, row = 16 * pw + 8 * (lane >> 4) + (lane & 7), gq = (lane >> 3) & 1;
            // Per frame: samples [0, len0) come from the stream's tail, the rest from the chunk.  P0 / P1 are byte pointers such
            // that sample i sits at P0 + 2 i (i < len0) or P1 + 2 i (i >= len0); len0 is a multiple of 8, so a 4-sample group
            // never straddles.
            struct Frame { const char* P0; const char* P1; int len0; int x0; bool active; };
            auto setup = [&](int tile) {
                Frame fr;
                fr.P0 = fr.P1 = reinterpret_cast<const char*>(pcm); fr.len0 = 0; fr.x0 = 0;
                const int f = tile * 128 + row;
                fr.active = tile < n_tiles && f < n_frames;
                if (fr.active) {
                    const int e = sm.fr[f], s = e >> 2, j = e & 3;
                    const int dj = sm.st_d[s] + j * hop;                // frame start relative to the chunk
                    const int16_t* chunk_p = pcm + (long long)(base + s) * chunk;
                    if (dj >= 0) { fr.P1 = reinterpret_cast<const char*>(chunk_p + dj); }
                    else {                                             // dj < 0: the tail starts at frame c0, this frame at j * hop inside it
                        fr.len0 = min(used, -dj);
                        fr.P0 = reinterpret_cast<const char*>(st.tail + (long long)sm.st_sid[s] * st.tail_cap + j * hop);
                        fr.P1 = reinterpret_cast<const char*>(chunk_p - fr.len0);
                    }
                    fr.x0 = __ldg(reinterpret_cast<const int16_t*>(fr.len0 > 0 ? fr.P0 : fr.P1));   // converted by its first use, a K-step later
                }
                return fr;
            };
            // This lane owns one 32-byte sector of every 64-byte row q of its frame: samples 16 gq + 32 q .. + 15, i.e. the
            // K-groups g = 4 gq + ks of the four K-steps.  One 16-byte load per row serves two K-steps (half h = ks >> 1).
            auto load = [&](const Frame& fr, int h, uint4 (&raw)[16]) {
                if (!fr.active) return;
                const int i0 = 16 * gq + 8 * h;                        // first sample of the half in row 0
                const int qs = (fr.len0 - i0 + 31) >> 5;               // rows q < qs come from the tail
                const char* a0 = fr.P0 + 2 * i0;
                const char* a1 = fr.P1 + 2 * i0;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    if (q < qs) raw[q] = __ldg(reinterpret_cast<const uint4*>(a0 + 64 * q));
                    else raw[q] = __ldg(reinterpret_cast<const uint4*>(a1 + 64 * q));
                }
            };
            Frame cur = setup(0), nxt = cur;
            uint4 raw[16];
            load(cur, 0, raw);
            for (int tile = 0; tile < n_tiles; ++tile) {
                const bool active = cur.active;
                const float x0f = (float)cur.x0;
                if (active && gq == 0) sm.x0[tile & 1][row] = x0f;
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    // the next tile's frame is looked up well before its first load
                    if (h == 1) nxt = setup(tile + 1);
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub, ++n_ksteps) {
                        const int ks = 2 * h + sub;
                        // two sample pairs (j = 0, 1 and j = 2, 3 of the K-group's four samples): two 16-point real DFTs each, then
                        // one 8-byte store per block and piece into this lane's K-group (half (jp ^ gq), see tcd_kslot)
#pragma unroll
                        for (int jp = 0; jp < 2; ++jp) {
                            float yv[TCD_BLOCKS][4];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {             // the pair's two samples: low / high half of the word
                                float x[16], yr[9], yi[9];
                                if (active) {
#pragma unroll
                                    for (int q = 0; q < 16; ++q) {
                                        const uint32_t w = (sub == 0 ? (jp == 0 ? raw[q].x : raw[q].y) : (jp == 0 ? raw[q].z : raw[q].w)) ^ 0x80008000u;
                                        float lo, hi;
                                        tcd_cvt2(w, lo, hi);
                                        x[q] = e ? hi : lo;
                                    }
                                }
                                // software pipeline: the registers are free once the last sample of this half is converted; the loads
                                // of the next half (of the next tile after the second one) fly while it is transformed and stored
                                if (sub == 1 && jp == 1 && e == 1) {
                                    if (h == 0) load(cur, 1, raw);
                                    else load(nxt, 0, raw);
                                }
                                if (active) {
                                    rdft16_x2(x, yr, yi);
                                    yv[0][e] = yr[0] - TCD_X0_Y * x0f; yv[0][2 + e] = yr[8];        // Y_0 of (x - x[0]): exact
#pragma unroll
                                    for (int r = 1; r < 8; ++r) { yv[r][e] = yr[r]; yv[r][2 + e] = yi[r]; }
                                }
                            }
                            // the tensor core has finished reading the previous K-step's tiles
                            if (jp == 0) mbar_wait(&sm.a_empty, (n_ksteps & 1) ^ 1);
                            if (active) {
#pragma unroll
                                for (int b = 0; b < TCD_BLOCKS; ++b)
                                    tcd_put4(&sm.a_hi[b][gq][row][4 * (jp ^ gq)], &sm.a_lo[b][gq][row][4 * (jp ^ gq)], yv[b][0], yv[b][1], yv[b][2], yv[b][3]);
                            }
                        }
                        // ---- hand-over: the warp that arrives last issues the K-step's MMAs (no thread waits for the others)
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            uint32_t prev;
                            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(smem_u32(&sm.arrivals)) : "memory");
                            if (prev == TC2_PROD_WARPS * (n_ksteps + 1) - 1) {
                                if (!sm.b_loaded) { mbar_wait(&sm.b_ready, 0); sm.b_loaded = 1; }
                                if (ks == 0) mbar_wait(&sm.d_empty, (n_tiles_done & 1) ^ 1);     // the epilogue has drained the previous tile
                                tc5_fence_after();
                                // descriptors of the first operand tiles; every other tile is a constant number of 16-byte units further on
                                const uint64_t dA_hi = tc5_desc(&sm.a_hi[0][0][0][0], 2048, 128), dA_lo = tc5_desc(&sm.a_lo[0][0][0][0], 2048, 128);
                                const uint64_t dbh0 = tc5_desc(&sm.b_hi[0][ks][0][0], 4096, 128), dbl0 = tc5_desc(&sm.b_lo[0][ks][0][0], 4096, 128);   // K-groups ks and 4 + ks
#pragma unroll
                                for (int b = 0; b < TCD_BLOCKS; ++b) {
                                    const uint64_t dah = dA_hi + b * 256, dal = dA_lo + b * 256;       // a block of A: 2 x 128 x 16 B
                                    const uint64_t dbh = dbh0 + b * 512, dbl = dbl0 + b * 512;         // a block of B: 8 x 64 x 16 B
                                    const uint32_t d = tmem + 64 * b;
                                    tcd_mma(d, dal, dbh, idesc, ks > 0);
                                    tcd_mma(d, dah, dbl, idesc, 1);
                                    tcd_mma(d, dah, dbh, idesc, 1);
                                }
                                tc5_commit(&sm.a_empty);                                 // arrives when these MMAs have read the A tiles
                                if (ks == TCD_KSTEPS - 1) tc5_commit(&sm.d_full);
                            }
                        }
                        __syncwarp();
                    }
                }
                ++n_tiles_done;
                cur = nxt;
                // ---- new tails (chunk >= 512: they lie inside the chunk).  Only a stream's FIRST new frame can reach into the old
                // tail (hop >= 512), so the warp that owns that frame's row is the tail's only reader; it has consumed those
                // loads by now and replaces the tail itself, four streams' loads in flight per batch -- no block-wide barrier.
                __syncwarp();
#pragma unroll 1
                for (int batch = 0; batch < 4; ++batch) {
                    int4 v[4][2];
                    int4* dst[4];
                    int nv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int fi = tile * 128 + 16 * pw + 4 * batch + u;
                        nv[u] = 0;
                        dst[u] = nullptr;
                        if (fi < n_frames) {
                            const int e = sm.fr[fi], s = e >> 2, cnt = sm.st_cnt[s];
                            if ((e & 3) == 0) {
                                const int off = min((int)sm.st_d[s] + cnt * hop, chunk);          // new tail = chunk[off, chunk)
                                nv[u] = (chunk - off) >> 3;
                                const int4* src = reinterpret_cast<const int4*>(pcm + (long long)(base + s) * chunk + off);
                                dst[u] = reinterpret_cast<int4*>(st.tail + (long long)sm.st_sid[s] * st.tail_cap);
                                if (lane < nv[u]) v[u][0] = __ldg(src + lane);
                                if (lane + 32 < nv[u]) v[u][1] = __ldg(src + lane + 32);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (lane < nv[u]) dst[u][lane] = v[u][0];
                        if (lane + 32 < nv[u]) dst[u][lane + 32] = v[u][1];
                    }
                }
            }
        } else {
            // ================= epilogue: thread <-> frame = TMEM lane
            for (int tile = 0; tile < n_tiles; ++tile) {
                const int f = tile * 128 + tid;
                const bool active = f < n_frames;
                const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
                float rise[G::n_filt + 1], fall[G::n_filt + 1];
#pragma unroll
                for (int j = 0; j <= G::n_filt; ++j) { rise[j] = 0.f; fall[j] = 0.f; }
                float tot = 0.f;
                tc2_mbar_wait_idle(&sm.d_full, n_tiles_done & 1);
                tc5_fence_after();
                const float x0f = sm.x0[tile & 1][tid];
                uint32_t buf[2][32];
                tc2_ld32(t_row, buf[0]);
                tc2_wait_ld(buf[0]);
                tc2_static_for<16>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    if constexpr (c + 1 < 16) tc2_ld32(t_row + 32 * (c + 1), buf[(c + 1) & 1]);
                    uint32_t (&v)[32] = buf[c & 1];
                    tc2_static_for<16>([&](auto mm) {
                        constexpr int m = decltype(mm)::value;
                        constexpr int bin = tcd_chunk_bin_c(c, m);
                        float re = __uint_as_float(v[m]);
                        const float im = __uint_as_float(v[16 + m]);
                        float p;
                        if constexpr (c == 0 && m == 0) {
                            re = fmaf(TCD_X0_D, x0f, re);                         // undo the constant subtracted from the frame
                            p = re * re;
                            // the Im X[0] slot carries X[256]
                            constexpr int s256 = tc2_seg<G>(256);
                            constexpr float wr256 = tc2_wrise<G>(256), wf256 = tc2_wfall<G>(256);
                            const float p256 = im * im;
                            tot += p256;
                            if constexpr (wr256 != 0.f) rise[s256] = fmaf(wr256, p256, rise[s256]);
                            if constexpr (wf256 != 0.f) fall[s256] = fmaf(wf256, p256, fall[s256]);
                        } else {
                            p = fmaf(im, im, re * re);
                        }
                        constexpr int s = tc2_seg<G>(bin);
                        constexpr float wr = tc2_wrise<G>(bin), wf = tc2_wfall<G>(bin);
                        tot += p;
                        if constexpr (wr != 0.f) rise[s] = fmaf(wr, p, rise[s]);
                        if constexpr (wf != 0.f) fall[s] = fmaf(wf, p, fall[s]);
                    });
                    if constexpr (c + 1 < 16) tc2_wait_ld(buf[(c + 1) & 1]);
                });
                tc5_fence_before();
                mbar_arrive(&sm.d_empty);                                        // TMEM may be overwritten by the next tile
                ++n_tiles_done;
                if (active) {
                    const int e = sm.fr[f], s = e >> 2, j = e & 3;
                    int slot = sm.st_slot[s] + j;
                    if (slot >= st.ring_rows) slot -= st.ring_rows;
                    float* rowp = st.ring + ((long long)sm.st_sid[s] * st.ring_rows + slot) * st.row_stride;
                    float lg[G::n_filt];
#pragma unroll
                    for (int q = 0; q < G::n_filt; ++q) lg[q] = __logf(fmaxf((rise[q] + fall[q + 1]) * tab.pscale, K1_EPS));
                    rowp[0] = __logf(fmaxf(tot * tab.pscale, K1_EPS));
#pragma unroll 1
                    for (int o = 1; o < tab.n_out; ++o) {
                        const float4* d4 = reinterpret_cast<const float4*>(sm.dct[o]);
                        float v0 = 0.f, v1 = 0.f;
#pragma unroll
                        for (int q = 0; q < G::n_filt / 4; ++q) {
                            const float4 d = d4[q];
                            v0 = fmaf(d.x, lg[4 * q], v0); v1 = fmaf(d.y, lg[4 * q + 1], v1);
                            v0 = fmaf(d.z, lg[4 * q + 2], v0); v1 = fmaf(d.w, lg[4 * q + 3], v1);
                        }
                        rowp[o] = v0 + v1;
                    }
                }
            }
        }
        __syncthreads();              // every role is done with this super-group's lists
        // streams that completed no frame this tick (chunk < hop only): nothing read their old tail, which the chunk replaces
        for (int s = warp; s < sg_n; s += TC2_THREADS / 32) {
            if (sm.st_cnt[s] != 0) continue;
            const int off = min((int)sm.st_d[s], chunk);
            const int nv = (chunk - off) >> 3;
            const int4* src = reinterpret_cast<const int4*>(pcm + (long long)(base + s) * chunk + off);
            int4* dst = reinterpret_cast<int4*>(st.tail + (long long)sm.st_sid[s] * st.tail_cap);
            for (int v = lane; v < nv; v += 32) dst[v] = __ldg(src + v);
        }
        __syncthreads();
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

}  // namespace pb
