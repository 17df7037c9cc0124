// mfcc_tc2.cuh -- K1 on the 5th-generation tensor cores (tcgen05 + TMEM): the stateful MFCC tick for the reference's
// default geometry (n_fft 512, 20 mel filters at 16 kHz; precise/params.py:140-144).
//
// Replaces, per frame, np.fft.rfft(frame, n=512) -> power -> mel filterbank -> log -> DCT -> c0 of sonopy.mfcc_spec as the
// reference calls it (precise/vectorization.py:36-39) and the carry-buffer bookkeeping of Listener.update_vectors
// (precise/network_runner.py:125-146).  Arithmetic, operand tables and CPU model: mfcc_tc.cuh.
//
// Organisation of a CTA (one per SM, 12 warps):
//   * super-groups of up to 1024 streams share one frame list; a TILE is 32 frames = 128 accumulator rows (frame, h):
//     row h of a frame carries up to three of its nine 64-column GEMM blocks (tcd_blk_h / tcd_blk_s), so a tile needs
//     192 TMEM columns and two tiles are in flight (MMAs of tile t + 1 under the epilogue of tile t);
//   * PCM arrives by 1-D bulk copies (cp.async.bulk, one or two per frame: tail part + chunk part) into a double-buffered
//     staging area, a tile ahead of its use -- no global-load instruction or register sits on the critical path;
//   * producers (warps 4-11): lane <-> (frame, K-group g of four inputs n2 = 4 g .. 4 g + 3): sixteen LDS.64, four 16-point
//     real DFTs, the w512 twiddles, fp16 hi / lo split, one 8-byte store per block, piece and sample pair.  Warps 4-7 own
//     g = 0..3 (K-steps 0, 1 = stage X), warps 8-11 g = 4..7 (stage Y); the warp of a stage that delivers last issues that
//     stage's 18 MMAs (M = 128, N = 64, K = 16; 3 slots x 2 K-steps x 3 passes), so nobody waits for a hand-over;
//   * epilogue (warps 0-3): warp h owns the rows (frame, h): tcgen05.ld, power, mel edge sums with compile-time bins and
//     weights, partial sums of the four rows of a frame exchanged through shared memory, log, DCT, ring row;
//   * new tails are copied by the stage-Y producers once the bulk copies that read the old tails have landed; sample
//     counters are bumped by the bookkeeping pass;
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
constexpr int TC2_SG_MAX = 1024;                // streams per super-group
constexpr int TC2_MAX_NEW = 4;                  // frames a stream may complete per tick
constexpr int TC2_MAX_FRAMES = TC2_SG_MAX * TC2_MAX_NEW;
constexpr int TC2_TILE = 32;                    // frames per tile
constexpr int TC2_SLOTS = 3;                    // 64-column MMA slots per tile row
constexpr int TC2_TMEM_COLS = 64 * TC2_SLOTS;   // accumulator columns per tile (two tiles resident)
constexpr int TC2_PCM_STRIDE = 1024 + 32;       // bytes per staged frame: 4 frames x 4 K-groups of a half-warp hit 32 distinct banks
constexpr int TC2_A_LBO = 2048 + 32;            // bytes between consecutive K-groups of an A tile: 4 K-groups x 4 rows, 8-byte stores, no conflict
constexpr int TC2_A_TILE = 4 * TC2_A_LBO;       // one (piece, stage, slot) operand tile: 4 K-groups x 128 rows x 16 bytes (+ padding)

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
    __half b_hi[8][64][8];                       // the 64 x 64 twiddle operand (K-major canonical layout), pieces hi / lo: 8 KB each
    __half b_lo[8][64][8];
    unsigned char a[2][2][TC2_SLOTS][TC2_A_TILE];     // A operands [piece hi / lo][stage X / Y][slot]: 4 K-groups x 128 rows each
    unsigned char pcm[2][TC2_TILE][TC2_PCM_STRIDE];   // staged frames (int16 x 512), two tiles
    float part[21][128];                         // mel sums (20) + total power of every tile row, to be added over the four rows of a frame
    float lgm[20][TC2_TILE];                     // log-mel values of the tile's frames
    float tw[32 * TCD_TW_STRIDE];                // w512^(n2 r), r = 1..8
    float dct[TCD_MAX_OUT][24];
    float x0[4][TC2_TILE];                       // ring over tiles: the constant subtracted from the frame (its first sample)
    int st_sid[TC2_SG_MAX];                      // stream id
    short st_d[TC2_SG_MAX];                      // first new frame's start relative to the chunk: c0 * hop - n0
    unsigned short st_off[TC2_SG_MAX];           // index of the stream's first frame in the list
    unsigned char st_slot[TC2_SG_MAX];           // ring slot of the first new frame
    unsigned char st_cnt[TC2_SG_MAX];            // frames completed by this tick
    unsigned short fr[TC2_MAX_FRAMES];           // frame list: (local stream << 2) | sub-frame
    int n_frames;
    unsigned int arrivals[2];                    // producer warps that have delivered a tile's stage X / Y (monotonic)
    unsigned int x_issued;                       // tiles whose stage-X MMAs have been issued (monotonic)
    unsigned long long pcm_full[2], pcm_empty[2], a_empty[2], d_full[2], d_empty[2], b_ready;
    uint32_t tmem_base;
};

struct Tc2Tables {               // device pointers
    const uint4* b;              // [2][8][64] x 16 bytes: hi then lo (tcd_build_b)
    const float* tw;             // [32][TCD_TW_STRIDE]
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

// The mel edge sums of one 64-column block B of a tile row: up to 32 bins, each with compile-time segment and weights.
template <class G, int B>
__device__ __forceinline__ void tc2_block_bins(uint32_t taddr, float x0f, float (&rise)[G::n_filt + 1], float (&fall)[G::n_filt + 1], float& tot) {
    tc2_static_for<2>([&](auto hh) {
        constexpr int half = decltype(hh)::value;
        constexpr bool any = tcd_col_bin(B, 32 * half) >= 0 || tcd_col_bin(B, 32 * half + 15) >= 0;
        if constexpr (any) {
            uint32_t v[32];
            tc2_ld32(taddr + 32 * half, v);
            tc2_wait_ld(v);
            tc2_static_for<16>([&](auto mm) {
                constexpr int m = decltype(mm)::value;
                constexpr int bin = tcd_col_bin(B, 32 * half + m);
                if constexpr (bin >= 0) {
                    float re = __uint_as_float(v[m]);
                    const float im = __uint_as_float(v[16 + m]);
                    float p;
                    if constexpr (bin == 0) { re = fmaf(TCD_X0_D, x0f, re); p = re * re; }          // undo the constant subtracted from the frame
                    else if constexpr (bin == 256) p = re * re;
                    else p = fmaf(im, im, re * re);
                    constexpr int s = tc2_seg<G>(bin);
                    constexpr float wr = tc2_wrise<G>(bin), wf = tc2_wfall<G>(bin);
                    tot += p;
                    if constexpr (wr != 0.f) rise[s] = fmaf(wr, p, rise[s]);
                    if constexpr (wf != 0.f) fall[s] = fmaf(wf, p, fall[s]);
                }
            });
        }
    });
}

template <class G, int H>
__device__ __forceinline__ void tc2_row_bins(uint32_t t_row, float x0f, float (&rise)[G::n_filt + 1], float (&fall)[G::n_filt + 1], float& tot) {
    tc2_static_for<TC2_SLOTS>([&](auto ss) {
        constexpr int s = decltype(ss)::value;
        constexpr int b = tcd_hs_blk(H, s);
        if constexpr (b >= 0) tc2_block_bins<G, b>(t_row + 64 * s, x0f, rise, fall, tot);
    });
}

template <class G>
__global__ void __launch_bounds__(TC2_THREADS, 1)
mfcc_tc2_stream_kernel(const int16_t* __restrict__ pcm, const int* __restrict__ ids, int n, int sg, int chunk, int hop,
                       Tc2Tables tab, StreamState st) {
    extern __shared__ __align__(128) unsigned char tc2_raw[];
    Tc2Smem& sm = *reinterpret_cast<Tc2Smem*>(tc2_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int used = 512;
    static_assert(G::n_filt == 20, "the epilogue splits 20 filters and the DCT rows over the four rows of a frame");

    // ---- one-time setup: barriers, TMEM (512 columns: two tiles of 192), operand / twiddle / DCT tables
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&sm.pcm_full[i], TC2_TILE); mbar_init(&sm.pcm_empty[i], TC2_PROD_WARPS);
            mbar_init(&sm.a_empty[i], 1);
            mbar_init(&sm.d_full[i], 1); mbar_init(&sm.d_empty[i], TC2_EPI_WARPS * 32);
        }
        mbar_init(&sm.b_ready, 1);
        sm.arrivals[0] = sm.arrivals[1] = 0; sm.x_issued = 0;
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < TCD_MAX_OUT * 24; e += TC2_THREADS) (&sm.dct[0][0])[e] = __ldg(tab.dct + e);
    for (int e = tid; e < 32 * TCD_TW_STRIDE; e += TC2_THREADS) sm.tw[e] = __ldg(tab.tw + e);
    fence_proxy_async();
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    if (tid == TC2_EPI_WARPS * 32) {
        mbar_expect_tx(&sm.b_ready, 2u * 8192u);
        bulk_g2s(&sm.b_hi[0][0][0], tab.b, 8192u, &sm.b_ready);
        bulk_g2s(&sm.b_lo[0][0][0], reinterpret_cast<const char*>(tab.b) + 8192, 8192u, &sm.b_ready);
    }
    const uint32_t tmem = sm.tmem_base;
    const uint32_t idesc = tcd_idesc(64);

    // Tiles are numbered across super-groups (T = tiles of the earlier groups): tile G uses staging / TMEM buffer G & 1 for the
    // (G >> 1)-th time, which fixes the phase parity of every barrier below.
    uint32_t T = 0;

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
        const int n_tiles = (n_frames + TC2_TILE - 1) / TC2_TILE;

        if (warp >= TC2_EPI_WARPS) {
            // ================= producers
            const int pw = warp - TC2_EPI_WARPS, stage = pw >> 2;
            const int g_l = lane & 3, fl = 8 * (pw & 3) + (lane >> 2), g = 4 * stage + g_l;       // lane <-> (frame fl of the tile, K-group g)
            // bulk copies of tile t's frames (global tile Gt) into staging buffer Gt & 1, by the stage-X lanes with g_l == 0
            // (one lane per frame), once every producer warp has read the tile that used the buffer before
            auto stage_pcm = [&](int t, uint32_t Gt) {
                mbar_wait(&sm.pcm_empty[Gt & 1], ((Gt >> 1) & 1) ^ 1);
                unsigned long long* bar = &sm.pcm_full[Gt & 1];
                const int f = t * TC2_TILE + fl;
                if (f >= n_frames) { mbar_arrive(bar); return; }
                const int e = sm.fr[f], s = e >> 2, j = e & 3;
                const int dj = sm.st_d[s] + j * hop;                    // frame start relative to the chunk
                const int16_t* chunk_p = pcm + (long long)(base + s) * chunk;
                unsigned char* dst = sm.pcm[Gt & 1][fl];
                mbar_expect_tx(bar, 1024u);
                if (dj >= 0) bulk_g2s(dst, chunk_p + dj, 1024u, bar);
                else {                                                  // dj < 0: the tail starts at frame c0, this frame at j * hop inside it
                    const int len0 = min(used, -dj);
                    bulk_g2s(dst, st.tail + (long long)sm.st_sid[s] * st.tail_cap + j * hop, 2u * len0, bar);
                    if (len0 < used) bulk_g2s(dst + 2 * len0, chunk_p, 2u * (used - len0), bar);
                }
            };
            const bool stager = stage == 0 && g_l == 0;
            if (stager) {                                               // prologue: the group's first two tiles
                if (n_tiles > 0) stage_pcm(0, T);
                if (n_tiles > 1) stage_pcm(1, T + 1);
            }
            for (int tile = 0; tile < n_tiles; ++tile) {
                const uint32_t Gt = T + tile;
                const int f = tile * TC2_TILE + fl;
                const bool active = f < n_frames;
                // ---- this lane's 64 samples: x[n2 + 32 q], n2 = 4 g .. 4 g + 3
                mbar_wait(&sm.pcm_full[Gt & 1], (Gt >> 1) & 1);
                uint2 raw[16];
                {
                    const unsigned char* src = sm.pcm[Gt & 1][fl] + 8 * g;
#pragma unroll
                    for (int q = 0; q < 16; ++q) raw[q] = *reinterpret_cast<const uint2*>(src + 64 * q);
                }
                const float x0f = (float)*reinterpret_cast<const int16_t*>(sm.pcm[Gt & 1][fl]);
                if (stager && active) sm.x0[Gt & 3][fl] = x0f;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.pcm_empty[Gt & 1]);     // this warp has read its part of the staged tile
                // the stager keeps one tile ahead: tile + 1 goes into the buffer tile - 1 used (read long ago)
                if (stager && tile >= 1 && tile + 1 < n_tiles) stage_pcm(tile + 1, Gt + 1);
                // ---- new tails (chunk >= 512: they lie inside the chunk).  The bulk copies that read the old tails of this tile's
                // frames have landed; the stage-Y warps (which stage nothing) replace the tails of the streams whose first new
                // frame is among their eight frames, four streams' loads in flight per batch.
                if (stage == 1) {
#pragma unroll 1
                    for (int batch = 0; batch < 2; ++batch) {
                        int4 v[4][2];
                        int4* dst[4];
                        int nv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int fi = tile * TC2_TILE + 8 * (pw & 3) + 4 * batch + u;
                            nv[u] = 0;
                            dst[u] = nullptr;
                            if (fi < n_frames) {
                                const int e = sm.fr[fi], s = e >> 2, cnt = sm.st_cnt[s];
                                if ((e & 3) == 0) {
                                    const int off = min((int)sm.st_d[s] + cnt * hop, chunk);      // new tail = chunk[off, chunk)
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
                // ---- two sample pairs (j = 0, 1 from the low words, j = 2, 3 from the high words): per sample a 16-point real DFT
                // and the twiddles, then one 8-byte store per block and piece into this lane's K-group (half (jp ^ (g & 1)))
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    float zv[TCD_BLOCKS][4];                            // per block: Re, Re, Im, Im of the pair's two inputs
                    if (active) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            float x[16], yr[9], yi[9], zr[9], zi[9];
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                float lo, hi;
                                tcd_cvt2((jp ? raw[q].y : raw[q].x) ^ 0x80008000u, lo, hi);
                                x[q] = e ? hi : lo;
                            }
                            rdft16_x2(x, yr, yi);
                            tcd_twiddle(yr, yi, sm.tw + (4 * g + 2 * jp + e) * TCD_TW_STRIDE, TCD_X0_Y * x0f, zr, zi);
#pragma unroll
                            for (int b = 0; b < TCD_BLOCKS; ++b) { zv[b][e] = zr[b]; zv[b][2 + e] = zi[b]; }
                        }
                    }
                    // the tensor core has finished reading this stage's operand tiles of the previous tile
                    if (jp == 0) mbar_wait(&sm.a_empty[stage], (Gt & 1) ^ 1);
                    if (active) {
#pragma unroll
                        for (int b = 0; b < TCD_BLOCKS; ++b) {
                            const int o = tcd_blk_s(b) * TC2_A_TILE + g_l * TC2_A_LBO + (fl + 32 * tcd_blk_h(b)) * 16 + 8 * (jp ^ (g & 1));
                            tcd_put4(reinterpret_cast<__half*>(&sm.a[0][stage][0][o]), reinterpret_cast<__half*>(&sm.a[1][stage][0][o]),
                                     zv[b][0], zv[b][1], zv[b][2], zv[b][3]);
                        }
                    }
                }
                // ---- hand-over: the warp of this stage that arrives last issues the stage's MMAs (no thread waits for the others)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    uint32_t prev;
                    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(smem_u32(&sm.arrivals[stage])) : "memory");
                    if (prev == 4 * (Gt + 1) - 1) {
                        const uint32_t tb = Gt & 1;
                        if (stage == 0) {
                            if (Gt == 0) mbar_wait(&sm.b_ready, 0);
                            mbar_wait(&sm.d_empty[tb], ((Gt >> 1) & 1) ^ 1);          // the epilogue has drained the tile that used these columns
                        } else {
                            while (*reinterpret_cast<volatile unsigned int*>(&sm.x_issued) < Gt + 1) { }  // stage X's MMAs go first
                            __threadfence_block();
                        }
                        tc5_fence_after();
                        const uint64_t da_hi = tc5_desc(&sm.a[0][stage][0][0], TC2_A_LBO, 128), da_lo = tc5_desc(&sm.a[1][stage][0][0], TC2_A_LBO, 128);
                        const uint64_t db_hi = tc5_desc(&sm.b_hi[4 * stage][0][0], 1024, 128), db_lo = tc5_desc(&sm.b_lo[4 * stage][0][0], 1024, 128);
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)                  // the stage's two K-steps: K-groups 2 kk, 2 kk + 1 of the stage
#pragma unroll
                            for (int s = 0; s < TC2_SLOTS; ++s) {
                                const uint64_t dah = da_hi + (uint64_t)((s * TC2_A_TILE + 2 * kk * TC2_A_LBO) >> 4);
                                const uint64_t dal = da_lo + (uint64_t)((s * TC2_A_TILE + 2 * kk * TC2_A_LBO) >> 4);
                                const uint64_t dbh = db_hi + (uint64_t)(2 * kk * 64), dbl = db_lo + (uint64_t)(2 * kk * 64);    // a K-group of B: 64 x 16 B
                                const uint32_t d = tmem + tb * TC2_TMEM_COLS + 64 * s;
                                tcd_mma(d, dal, dbh, idesc, (stage | kk) != 0);
                                tcd_mma(d, dah, dbl, idesc, 1);
                                tcd_mma(d, dah, dbh, idesc, 1);
                            }
                        tc5_commit(&sm.a_empty[stage]);                 // arrives when these MMAs have read the A tiles
                        if (stage == 0) {
                            tc5_fence_before();
                            asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&sm.x_issued)), "r"(Gt + 1) : "memory");
                        } else {
                            tc5_commit(&sm.d_full[tb]);
                        }
                    }
                }
                __syncwarp();
            }
        } else {
            // ================= epilogue: warp h <-> rows (frame, h) = TMEM lanes 32 h .. 32 h + 31
            const int h = warp;
            for (int tile = 0; tile < n_tiles; ++tile) {
                const uint32_t Gt = T + tile, tb = Gt & 1;
                const int f = tile * TC2_TILE + lane;
                const bool active = f < n_frames;
                const uint32_t t_row = tmem + ((uint32_t)(h * 32) << 16) + tb * TC2_TMEM_COLS;
                float rise[G::n_filt + 1], fall[G::n_filt + 1];
#pragma unroll
                for (int j = 0; j <= G::n_filt; ++j) { rise[j] = 0.f; fall[j] = 0.f; }
                float tot = 0.f;
                tc2_mbar_wait_idle(&sm.d_full[tb], (Gt >> 1) & 1);
                tc5_fence_after();
                const float x0f = sm.x0[Gt & 3][lane];
                if (h == 0) tc2_row_bins<G, 0>(t_row, x0f, rise, fall, tot);
                else if (h == 1) tc2_row_bins<G, 1>(t_row, x0f, rise, fall, tot);
                else if (h == 2) tc2_row_bins<G, 2>(t_row, x0f, rise, fall, tot);
                else tc2_row_bins<G, 3>(t_row, x0f, rise, fall, tot);
                tc5_fence_before();
                mbar_arrive(&sm.d_empty[tb]);                                    // these TMEM columns may be overwritten
                // ---- the four rows of a frame meet in shared memory: row h then finishes filters 5 h .. 5 h + 4 and its share of the DCT
#pragma unroll
                for (int q = 0; q < G::n_filt; ++q) sm.part[q][tid] = rise[q] + fall[q + 1];
                sm.part[G::n_filt][tid] = tot;
                asm volatile("bar.sync 1, %0;" ::"n"(TC2_EPI_WARPS * 32) : "memory");
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    const int j = 5 * h + q;
                    const float m4 = (sm.part[j][lane] + sm.part[j][32 + lane]) + (sm.part[j][64 + lane] + sm.part[j][96 + lane]);
                    sm.lgm[j][lane] = __logf(fmaxf(m4 * tab.pscale, K1_EPS));
                }
                float c0 = 0.f;
                if (h == 0) {
                    const float t4 = (sm.part[G::n_filt][lane] + sm.part[G::n_filt][32 + lane]) +
                                     (sm.part[G::n_filt][64 + lane] + sm.part[G::n_filt][96 + lane]);
                    c0 = __logf(fmaxf(t4 * tab.pscale, K1_EPS));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(TC2_EPI_WARPS * 32) : "memory");
                if (active) {
                    const int e = sm.fr[f], s = e >> 2, j = e & 3;
                    int slot = sm.st_slot[s] + j;
                    if (slot >= st.ring_rows) slot -= st.ring_rows;
                    float* rowp = st.ring + ((long long)sm.st_sid[s] * st.ring_rows + slot) * st.row_stride;
                    float lg[G::n_filt];
#pragma unroll
                    for (int q = 0; q < G::n_filt; ++q) lg[q] = sm.lgm[q][lane];
                    // DCT rows o = h, h + 4, h + 8, ... (row 0 is replaced by c0 = log of the total power)
                    for (int o = h; o < tab.n_out; o += 4) {
                        const float4* d4 = reinterpret_cast<const float4*>(sm.dct[o]);
                        float v0 = 0.f, v1 = 0.f;
#pragma unroll
                        for (int q = 0; q < G::n_filt / 4; ++q) {
                            const float4 d = d4[q];
                            v0 = fmaf(d.x, lg[4 * q], v0); v1 = fmaf(d.y, lg[4 * q + 1], v1);
                            v0 = fmaf(d.z, lg[4 * q + 2], v0); v1 = fmaf(d.w, lg[4 * q + 3], v1);
                        }
                        rowp[o] = o == 0 ? c0 : v0 + v1;
                    }
                }
            }
        }
        T += n_tiles;
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
