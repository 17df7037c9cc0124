// api.cu -- C ABI of libprecise_b200.so (see include/precise_b200.h for the contract and the
// reference interface each entry point replaces).  Host side: table construction (float64, then
// rounded once), per-stream state allocation, launch configuration, the pinned/pipelined host
// entry point and CUDA-event profiling.  No CPU compute path exists here by design.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <cuda_fp16.h>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/precise_b200.h"
#include "gru_kernels.cuh"
#include "mfcc_kernels.cuh"
#include "mfcc_fast.cuh"
#include "gru_tc5.cuh"
#include "gru_tc5_big.cuh"
#include "mfcc_tc.cuh"
#include "mfcc_tc2.cuh"
#include "mfcc_tc3.cuh"

using namespace pb;

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(PB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int N_PROFILE_SLOTS = 4;
constexpr int PROFILE_POOL = 2048;
constexpr int K2_WARP_PATH_MAX = 8192;       // streams: below this the warp-per-stream GRU kernel wins (latency-bound regime)
constexpr int64_t HOST_ZERO_COPY_MAX = 64;  // streams: at or below this pb_update_host works in place on pinned host buffers
constexpr int HOST_PIPE = 3;                 // internal streams of pb_update_host
constexpr int64_t HOST_SUB_BATCH = 16384;    // streams per pipelined sub-batch (32 MiB of PCM at 1024 samples)

struct ProfSlot {
    std::vector<cudaEvent_t> ev;   // pairs
    int used = 0;
    double ms = 0.0;
    uint64_t launches = 0;
};

struct pb_handle {
    pb_config cfg;
    int sm_count = 148;
    // derived
    int used = 0, n_bins = 0, n_out = 0, feat = 0, ring_rows = 0, row_stride = 0, tail_cap = 0, max_new = 0;
    int rel_window = 0;              // samples before frame 0 is released: window_samples (sonopy), window_samples + hop_samples (speechpy drops the last complete frame)
    bool has_proj = false;           // default network: a second ring caches the input projections (gru_kernels.cuh)
    float* d_proj_ring = nullptr;
    bool proj_dirty = true;          // some ring rows lack a valid cached projection (weights changed / projection skipped)
    bool host_tick_proj = false;     // inside a pb_update_host tick whose sub-batches use the cache: short sub-batches keep it valid too
    float *d_proj_w = nullptr, *d_proj_b = nullptr;
    size_t k1_batch_smem = 0, k1_stream_smem = 0, k1_fast_smem = 0;
    bool force_generic = false;      // tests: exercise the generic kernels on the aligned geometry
    int k1_mode = 0;                 // 0 = default (mfcc_tc3 from TC3_MIN_STREAMS streams per tick on, else the FFT kernel with 32-bit set-up), 2 = always the FFT kernel, 3 = FFT kernel with the original 64-bit set-up, 4 = tcgen05 stage-2 kernel (mfcc_tc2), 5 = both DFT stages on tcgen05 (mfcc_tc3), 100 + w = 5 with warp w's timeline
    bool tcd_ok = false;             // geometry the tensor-core DFT tables cover (CPU model + kernel)
    bool tc2_ok = false;             // ... and by mfcc_tc2_stream_kernel<Tc2Geo20> (run-time mel tables equal its compile-time ones)
    std::vector<float> h_wrise, h_wfall; std::vector<int> h_grid;      // host copies for the lazily built tcd tables
    uint4* d_tcd_b = nullptr; float* d_tcd_tw = nullptr; float* d_tcd_dct = nullptr;
    bool tc3_ok = false;             // ... and by mfcc_tc3_kernel<Tc2Geo20> (both DFT stages on the tensor cores; needs chunk >= hop)
    uint4 *d_tc3_b1 = nullptr, *d_tc3_b2 = nullptr; float* d_tc3_tw = nullptr;
    Tc3Rec* d_tc3_recs = nullptr; unsigned int* d_tc3_counters = nullptr;
    int tc3_parity = 0;              // which of the two frame counters the next tick's plan kernel fills
    float tcd_tot_scale = 0.f;
    bool fast_ok = false;            // aligned geometry: warp-autonomous kernels (mfcc_fast.cuh)
    int npl = 0, maxc = 0, nol = 0;
    float4* d_ptab = nullptr;
    unsigned char* d_ctab = nullptr;
    float* d_dct_t = nullptr;
    // host copies of tables
    std::vector<double> fb;        // [n_filt][n_bins]
    std::vector<double> cd;
    int min_out = 0, max_out = 0;
    // device tables
    float *d_wrise = nullptr, *d_wfall = nullptr, *d_dct = nullptr;
    int* d_grid = nullptr;
    float2 *d_tw_stage = nullptr, *d_tw_post = nullptr, *d_tw_any = nullptr;
    double* d_cd = nullptr;
    // state
    StreamState st{};
    // weights
    bool have_weights = false;
    bool small_path = false;
    GruSmallW<20, 13> w_small;
    float *d_wcat = nullptr, *d_bias = nullptr, *d_wd = nullptr;
    float4* d_bfrag = nullptr;       // tensor-core GRU: pre-split, fragment-ordered weights
    uint4* d_bfrag16 = nullptr;      // ... recurrent part as fp16 hi / lo fragments (gru_mma16_kernel)
    uint4* d_xfrag16 = nullptr;      // ... and the input part (the scan projects a tick's new frames itself)
    float *d_mma_bias = nullptr, *d_mma_wd = nullptr;
    long long* d_dbg = nullptr;       // optional debug counters (pb_debug_counters)
    float *d_tcb = nullptr;           // tcgen05 wide-network GRU: [b1 tiles | b2 tiles | bias(384) | wd(128)]
    bool tcb_ok = false;
    int tcb_kx = 0;
    float *d_tc5 = nullptr;           // tcgen05 GRU: [b1_hi | b1_lo | b2_hi | b2_lo | bias(80) | wd(24)]
    int gru_mode = 0;                // 0 = auto, 1 = force CUDA-core small kernel, 2 = force tensor-core kernel, 3 = tcgen05 scan, 7 = tensor-core kernel with 32-stream warp tiles, 8 = tcgen05 scan over cached projections (opt-in, unvalidated)
    float bd = 0.f;
    // host pipeline
    cudaStream_t pipe[HOST_PIPE] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ev[HOST_PIPE] = {nullptr, nullptr, nullptr};
    int16_t* d_stage_pcm[HOST_PIPE] = {nullptr, nullptr, nullptr};
    int* d_stage_ids[HOST_PIPE] = {nullptr, nullptr, nullptr};
    float* d_stage_raw[HOST_PIPE] = {nullptr, nullptr, nullptr};
    double* d_stage_conf[HOST_PIPE] = {nullptr, nullptr, nullptr};
    uint8_t* d_stage_fired[HOST_PIPE] = {nullptr, nullptr, nullptr};
    unsigned long long* d_count = nullptr;
    unsigned long long* h_count_pinned = nullptr;
    // profiling
    bool profiling = false;
    ProfSlot prof[N_PROFILE_SLOTS];
};

// ------------------------------------------------------------------------------------------------
// table construction (host, float64), restating sonopy.filterbanks as the reference calls it
// (precise/vectorization.py:36-39): grid up to sample_rate, int() truncation, duplicate bins pushed
// forward, np.linspace(endpoint=False) edge weights.
static int build_mel(pb_handle* h, std::vector<float>& wrise, std::vector<float>& wfall, std::vector<int>& grid) {
    const pb_config& c = h->cfg;
    const int nb = h->n_bins, nf = c.n_filt;
    const double top = 1127.0 * log(1.0 + (double)c.sample_rate / 700.0);
    grid.assign(nf + 2, 0);
    long long shift = 0, prev = -1;
    if (c.vectorizer == PB_VEC_SPEECHPY_MFCCS) {
        // speechpy.feature.filterbanks as speechpy.feature.mfe calls it (precise/vectorization.py:40-42; the package itself is not in
        // the reference tree: PARITY UNPINNED, its published algorithm is restated): mel points between 0 and sample_rate / 2, corner bins
        // floor((coefficients + 1) * hz / sample_rate) with coefficients = n_fft / 2 + 1, triangles without de-duplication.
        const double top2 = 1127.0 * log(1.0 + 0.5 * (double)c.sample_rate / 700.0);
        for (int i = 0; i < nf + 2; ++i) {
            const double m = (i == nf + 1) ? top2 : (double)i * (top2 / (double)(nf + 1));
            const double hz = 700.0 * (exp(m / 1127.0) - 1.0);
            grid[i] = (int)floor((double)(nb + 1) * hz / (double)c.sample_rate);
        }
    } else
    for (int i = 0; i < nf + 2; ++i) {
        // np.linspace(0, top, nf + 2): i * step, last element forced to stop
        double m = (i == nf + 1) ? top : (double)i * (top / (double)(nf + 1));
        double hz = 700.0 * (exp(m / 1127.0) - 1.0);
        long long raw = (long long)(hz * (double)nb / (double)c.sample_rate);
        if (i == 0) prev = raw - 1;
        shift = std::max(0LL, shift + prev + 1 - raw);
        grid[i] = (int)(raw + shift);
        prev = raw;
    }
    if (grid[nf + 1] > nb)
        return fail(PB_ERR_INVALID, "mel grid exceeds the spectrum (%d > %d bins): the reference's sonopy.filterbanks raises here", grid[nf + 1], nb);
    h->fb.assign((size_t)nf * nb, 0.0);
    wrise.assign(nb, 0.f);
    wfall.assign(nb, 0.f);
    for (int i = 0; i < nf; ++i) {
        int lo = grid[i], mid = grid[i + 1], hi = grid[i + 2];
        for (int k = lo; k < mid; ++k) h->fb[(size_t)i * nb + k] = (double)(k - lo) * (1.0 / (double)(mid - lo));
        for (int k = mid; k < hi; ++k) h->fb[(size_t)i * nb + k] = (double)(k - mid) * (-1.0 / (double)(hi - mid)) + 1.0;
        for (int k = lo; k < mid; ++k) wrise[k] = (float)h->fb[(size_t)i * nb + k];
        for (int k = mid; k < hi; ++k) wfall[k] = (float)h->fb[(size_t)i * nb + k];
    }
    return PB_OK;
}

static void build_cdf(pb_handle* h) {
    // precise/threshold_decoder.py:38-43, :68-70 and functions.pdf (:104-108)
    const pb_config& c = h->cfg;
    const int resolution = 200;
    double lo = 0, hi = 0;
    for (int i = 0; i < c.n_thresholds; ++i) {
        double a = c.threshold_mu[i] + -4 * c.threshold_std[i], b = c.threshold_mu[i] + 4 * c.threshold_std[i];
        if (i == 0 || a < lo) lo = a;
        if (i == 0 || b > hi) hi = b;
    }
    h->min_out = (int)lo;
    h->max_out = (int)hi;
    const int range = h->max_out - h->min_out;
    const int num = resolution * range;
    h->cd.assign(std::max(num, 0), 0.0);
    if (num <= 0) return;
    const double step = num > 1 ? (double)(h->max_out - h->min_out) / (double)(num - 1) : 0.0;
    double run = 0.0;
    for (int j = 0; j < num; ++j) {
        double x = (j == num - 1 && num > 1) ? (double)h->max_out : (double)j * step + (double)h->min_out;
        double s = 0.0;
        for (int i = 0; i < c.n_thresholds; ++i) {
            double mu = c.threshold_mu[i], sd = c.threshold_std[i];
            double p = sd == 0 ? 0.0 : (1.0 / (sd * sqrt(2 * M_PI))) * exp(-((x - mu) * (x - mu)) / (2 * (sd * sd)));
            s = (i == 0) ? p : s + p;
        }
        run += s / (double)(resolution * c.n_thresholds);
        h->cd[j] = run;
    }
}

template <typename T>
static cudaError_t upload(T** dst, const std::vector<T>& v) {
    cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(v.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel, not to a handle: handles with different geometries
// (n_filt, hidden, ...) coexist in one process, so only ever raise it -- to the largest size any handle has asked for.
template <typename K>
static cudaError_t ensure_dyn_smem(K kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> granted;     // per (device, kernel): the attribute lives in the context
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = granted[std::make_pair(dev, (const void*)kernel)];
    if (bytes <= cur) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

// ------------------------------------------------------------------------------------------------
#define PB_API extern "C" __attribute__((visibility("default")))

PB_API int pb_abi_version(void) { return PB_ABI_VERSION; }
PB_API const char* pb_last_error(void) { return g_err; }
PB_API const char* pb_build_info(void) { return "precise_b200 sm_100a " __DATE__ " " __TIME__; }

PB_API int pb_config_default(pb_config* cfg) {
    if (!cfg) return fail(PB_ERR_INVALID, "cfg is null");
    memset(cfg, 0, sizeof(*cfg));
    cfg->abi_version = PB_ABI_VERSION;
    cfg->device = 0;
    cfg->max_streams = 1;
    cfg->chunk_samples = 1024;
    cfg->sample_rate = 16000;
    cfg->window_samples = 1600;
    cfg->hop_samples = 800;
    cfg->n_fft = 512;
    cfg->n_filt = 20;
    cfg->n_mfcc = 13;
    cfg->n_features = 29;
    cfg->use_delta = 0;
    cfg->vectorizer = PB_VEC_MFCCS;
    cfg->hidden = 20;
    cfg->activation = PB_ACT_LINEAR;
    cfg->recurrent_activation = PB_RACT_HARD_SIGMOID;
    cfg->n_thresholds = 1;
    cfg->threshold_mu[0] = 6.0;
    cfg->threshold_std[0] = 4.0;
    cfg->threshold_center = 0.2;
    cfg->sensitivity = 0.5;
    cfg->trigger_level = 3;
    return PB_OK;
}

PB_API void pb_destroy(pb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->d_wrise); cudaFree(h->d_wfall); cudaFree(h->d_dct); cudaFree(h->d_grid);
    cudaFree(h->d_tcd_b); cudaFree(h->d_tcd_tw); cudaFree(h->d_tcd_dct);
    cudaFree(h->d_tc3_b1); cudaFree(h->d_tc3_b2); cudaFree(h->d_tc3_tw); cudaFree(h->d_tc3_recs); cudaFree(h->d_tc3_counters);
    cudaFree(h->d_tw_stage); cudaFree(h->d_tw_post); cudaFree(h->d_tw_any); cudaFree(h->d_cd); cudaFree(h->d_ptab); cudaFree(h->d_ctab); cudaFree(h->d_dct_t);
    cudaFree(h->st.n_samples); cudaFree(h->st.tail); cudaFree(h->st.ring); cudaFree(h->st.trig);
    cudaFree(h->d_wcat); cudaFree(h->d_bias); cudaFree(h->d_wd); cudaFree(h->d_count);
    cudaFree(h->d_bfrag16); cudaFree(h->d_xfrag16); cudaFree(h->d_bfrag); cudaFree(h->d_mma_bias); cudaFree(h->d_mma_wd); cudaFree(h->d_proj_w); cudaFree(h->d_proj_b); cudaFree(h->d_proj_ring); cudaFree(h->d_tc5); cudaFree(h->d_tcb); cudaFree(h->d_dbg);
    if (h->h_count_pinned) cudaFreeHost(h->h_count_pinned);
    for (int i = 0; i < HOST_PIPE; ++i) {
        cudaFree(h->d_stage_pcm[i]); cudaFree(h->d_stage_ids[i]); cudaFree(h->d_stage_raw[i]);
        cudaFree(h->d_stage_conf[i]); cudaFree(h->d_stage_fired[i]);
        if (h->pipe[i]) cudaStreamDestroy(h->pipe[i]);
        if (h->pipe_ev[i]) cudaEventDestroy(h->pipe_ev[i]);
    }
    for (auto& p : h->prof)
        for (auto e : p.ev) cudaEventDestroy(e);
    delete h;
}

PB_API int pb_create(const pb_config* cfg, pb_handle** out) {
    if (!cfg || !out) return fail(PB_ERR_INVALID, "null argument");
    *out = nullptr;
    const pb_config& c = *cfg;
    if (c.abi_version != PB_ABI_VERSION) return fail(PB_ERR_INVALID, "abi_version %d != %d", c.abi_version, PB_ABI_VERSION);
    if (c.max_streams < 1) return fail(PB_ERR_INVALID, "max_streams must be >= 1");
    if (c.chunk_samples < 1) return fail(PB_ERR_INVALID, "chunk_samples must be >= 1");
    if (c.sample_rate < 1 || c.window_samples < 1 || c.hop_samples < 1 || c.n_features < 1 || c.hidden < 1)
        return fail(PB_ERR_INVALID, "sample_rate, window_samples, hop_samples, n_features, hidden must be positive");
    if (c.vectorizer != PB_VEC_MFCCS && c.vectorizer != PB_VEC_MELS && c.vectorizer != PB_VEC_SPEECHPY_MFCCS) return fail(PB_ERR_INVALID, "unknown vectorizer %d", c.vectorizer);
    if (!is_pow2(c.n_fft) || c.n_fft < 64 || c.n_fft > 1024)
        return fail(PB_ERR_UNSUPPORTED, "n_fft %d: powers of two in [64, 1024] are implemented (512 is the reference default)", c.n_fft);
    if (c.n_filt < 1 || c.n_filt > 64 || c.n_mfcc < 1 || c.n_mfcc > 64) return fail(PB_ERR_UNSUPPORTED, "n_filt and n_mfcc must be in [1, 64]");
    if (c.n_thresholds < 1 || c.n_thresholds > PB_MAX_THRESHOLDS) return fail(PB_ERR_INVALID, "n_thresholds must be in [1, %d]", PB_MAX_THRESHOLDS);
    if (c.activation < 0 || c.activation > 1 || c.recurrent_activation < 0 || c.recurrent_activation > 1)
        return fail(PB_ERR_UNSUPPORTED, "unsupported GRU activation");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (c.device < 0 || c.device >= ndev) return fail(PB_ERR_CUDA, "device %d not available (%d visible)", c.device, ndev);
    CK(cudaSetDevice(c.device));

    pb_handle* h = new (std::nothrow) pb_handle();
    if (!h) return fail(PB_ERR_CUDA, "out of host memory");
    h->cfg = c;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, c.device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    h->used = std::min(c.n_fft, c.window_samples);
    h->n_bins = c.n_fft / 2 + 1;
    h->n_out = c.vectorizer == PB_VEC_MELS ? c.n_filt : std::min(c.n_filt, c.n_mfcc);
    h->feat = h->n_out * (c.use_delta ? 2 : 1);
    h->row_stride = (h->n_out + 3) & ~3;
    // default-sized networks cache the input projection of every frame in a second ring (gru_mma_kernel<.., PROJ>)
    h->has_proj = c.hidden == 20 && h->n_out == 13 && !c.use_delta && c.vectorizer == PB_VEC_MFCCS &&
                  c.chunk_samples / c.hop_samples + 2 <= 8;      // long chunks (fed as sub-chunks, launch_stream_mfcc) may add more rows than the cache logic tracks
    // speechpy's stack_frames yields floor((len - window) / hop) frames, one fewer than sonopy's framing: frame k is released one hop later
    h->rel_window = c.window_samples + (c.vectorizer == PB_VEC_SPEECHPY_MFCCS ? c.hop_samples : 0);
    h->ring_rows = c.n_features + (h->rel_window - h->used) / c.hop_samples + 2;
    h->tail_cap = (h->used + 7) & ~7;            // rows stay 16-byte aligned
    h->max_new = c.chunk_samples / c.hop_samples + 2;

    const size_t k1_big = c.n_fft > 512 ? k1_big_smem + 16 : 0;      // n_fft = 1024: power rows and FFT scratch in the dynamic tail
    h->k1_batch_smem = sizeof(K1Smem) + (size_t)h->n_out * c.n_filt * sizeof(float) + k1_big;
    h->k1_stream_smem = sizeof(K1StreamSmem) + (size_t)h->n_out * c.n_filt * sizeof(float) + k1_big;
    std::vector<float> wrise, wfall;
    std::vector<int> grid;
    int rc = build_mel(h, wrise, wfall, grid);
    if (rc != PB_OK) { delete h; return rc; }
    // piece schedule of the 16-lane mel stage (mfcc_fast.cuh): <= 8 bins each, never across a grid point
    std::vector<int> pieces, seg_first(c.n_filt + 2, 0);
    {
        auto add_range = [&](int lo, int hi) {
            for (int k = lo; k < hi; k += K1F_PIECE_LEN) pieces.push_back(k | (std::min(K1F_PIECE_LEN, hi - k) << 16));
        };
        for (int i = 0; i <= c.n_filt; ++i) {
            seg_first[i] = (int)pieces.size();
            add_range(grid[i], std::min(grid[i + 1], h->n_bins));
        }
        seg_first[c.n_filt + 1] = (int)pieces.size();
        add_range(0, std::min(grid[0], h->n_bins));                       // bins outside the grid: total power only
        add_range(grid[c.n_filt + 1], h->n_bins);
    }
    const int n_pieces = (int)pieces.size();
    h->npl = (n_pieces + 15) / 16;
    h->nol = (h->n_out + 15) / 16;
    h->maxc = 1;
    for (int j = 0; j < c.n_filt; ++j) h->maxc = std::max(h->maxc, seg_first[j + 2] - seg_first[j]);
    // ptab[q][e][lane]: piece p = lane + 16 q, entry e -> (byte offset of the bin in P, w_rise, w_fall, 0); padding -> zero bin
    std::vector<float4> ptab((size_t)h->npl * 128);
    for (int q = 0; q < h->npl; ++q)
        for (int e = 0; e < 8; ++e)
            for (int lane = 0; lane < 16; ++lane) {
                const int pidx = lane + 16 * q;
                int bin = K1F_ZERO_BIN;
                float wr = 0.f, wf = 0.f;
                if (pidx < n_pieces) {
                    const int start = pieces[pidx] & 0xffff, len = pieces[pidx] >> 16;
                    if (e < len) { bin = start + e; wr = wrise[bin]; wf = wfall[bin]; }
                }
                const int off = bin * 4;
                float offf; memcpy(&offf, &off, 4);
                ptab[((size_t)q * 8 + e) * 16 + lane] = make_float4(offf, wr, wf, 0.f);
            }
    // ctab[j][c]: partial slots summed into filter j: rise partials of segment j, fall partials (index + 64) of segment j + 1
    std::vector<unsigned char> ctab((size_t)c.n_filt * h->maxc, 128);
    for (int j = 0; j < c.n_filt; ++j) {
        int w = 0;
        for (int pp = seg_first[j]; pp < seg_first[j + 1]; ++pp) ctab[(size_t)j * h->maxc + w++] = (unsigned char)pp;
        for (int pp = seg_first[j + 1]; pp < seg_first[j + 2]; ++pp) ctab[(size_t)j * h->maxc + w++] = (unsigned char)(64 + pp);
    }
    h->fast_ok = c.n_fft == 512 && h->used == 512 && c.hop_samples % 8 == 0 && n_pieces <= 64 && h->npl <= 4;
    h->h_wrise = wrise; h->h_wfall = wfall; h->h_grid = grid;
    h->tcd_ok = h->fast_ok && c.vectorizer == PB_VEC_MFCCS && c.n_filt <= TCD_MAX_FILT && h->n_out <= TCD_MAX_OUT &&
                c.chunk_samples % 8 == 0 && c.chunk_samples >= 512 && (c.chunk_samples + c.hop_samples - 1) / c.hop_samples <= TC2_MAX_NEW;
    h->tc2_ok = h->tcd_ok && c.hop_samples >= 512 && c.hop_samples <= 16384 && h->ring_rows <= 255 && h->n_out >= 1 &&
                (c.chunk_samples + c.hop_samples - 1) / c.hop_samples <= TC2_MAX_NEW &&
                tc2_geo_matches<Tc2Geo20>(c.n_filt, h->n_bins, grid, wrise, wfall);
    h->tc3_ok = h->tc2_ok && c.chunk_samples >= c.hop_samples && c.chunk_samples <= 32760;
    h->k1_fast_smem = K1F_WARPS * sizeof(K1FWarp) + (size_t)h->npl * 128 * sizeof(float4) +
                      (size_t)c.n_filt * 16 * h->nol * sizeof(float) + (((size_t)c.n_filt * h->maxc + 15) & ~(size_t)15);
    // DCT-II, norm='ortho' (scipy.fftpack.dct as sonopy.mfcc_spec calls it), first n_out rows
    std::vector<float> dct((size_t)h->n_out * c.n_filt);
    for (int k = 0; k < h->n_out; ++k)
        for (int n = 0; n < c.n_filt; ++n) {
            double v = cos(M_PI * k * (2 * n + 1) / (2.0 * c.n_filt)) * sqrt(2.0 / c.n_filt);
            if (k == 0) v *= sqrt(0.5);
            dct[(size_t)k * c.n_filt + n] = (float)v;
        }
    std::vector<float2> tws(256), twp(16);
    for (int n2 = 0; n2 < 16; ++n2)
        for (int k1 = 0; k1 < 16; ++k1) {
            double a = 2.0 * M_PI * (double)(n2 * k1) / 256.0;
            tws[n2 * 16 + k1] = make_float2((float)cos(a), (float)-sin(a));
        }
    for (int k1 = 0; k1 < 16; ++k1) {
        double a = 2.0 * M_PI * (double)k1 / 512.0;
        twp[k1] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<float2> twa(c.n_fft / 2);
    for (int k = 0; k < c.n_fft / 2; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)c.n_fft;
        twa[k] = make_float2((float)cos(a), (float)-sin(a));
    }
    build_cdf(h);

#define CKH(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            pb_destroy(h);                                                                        \
            return fail(PB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));             \
        }                                                                                         \
    } while (0)
    CKH(upload(&h->d_wrise, wrise));
    CKH(upload(&h->d_wfall, wfall));
    CKH(upload(&h->d_grid, grid));
    CKH(upload(&h->d_dct, dct));
    CKH(upload(&h->d_tw_stage, tws));
    CKH(upload(&h->d_tw_post, twp));
    CKH(upload(&h->d_tw_any, twa));
    CKH(upload(&h->d_cd, h->cd));
    {
        std::vector<float> dct_t((size_t)c.n_filt * 16 * h->nol, 0.f);
        for (int k = 0; k < h->n_out; ++k)
            for (int n = 0; n < c.n_filt; ++n) dct_t[(size_t)n * 16 * h->nol + k] = dct[(size_t)k * c.n_filt + n];
        CKH(upload(&h->d_dct_t, dct_t));
    }
    CKH(upload(&h->d_ptab, ptab));
    CKH(upload(&h->d_ctab, ctab));
    const size_t S = (size_t)c.max_streams;
    h->st.tail_cap = h->tail_cap; h->st.ring_rows = h->ring_rows; h->st.row_stride = h->row_stride;
    CKH(cudaMalloc((void**)&h->st.n_samples, S * sizeof(long long)));
    CKH(cudaMalloc((void**)&h->st.tail, S * h->tail_cap * sizeof(int16_t)));
    CKH(cudaMalloc((void**)&h->st.ring, S * h->ring_rows * h->row_stride * sizeof(float)));
    CKH(cudaMalloc((void**)&h->st.trig, S * sizeof(int)));
    CKH(cudaMemset(h->st.n_samples, 0, S * sizeof(long long)));
    CKH(cudaMemset(h->st.tail, 0, S * h->tail_cap * sizeof(int16_t)));
    CKH(cudaMemset(h->st.ring, 0, S * h->ring_rows * h->row_stride * sizeof(float)));
    if (h->has_proj) CKH(cudaMalloc((void**)&h->d_proj_ring, ((S + 15) / 16) * h->ring_rows * PROJ_BLOCK * sizeof(float)));
    CKH(cudaMemset(h->st.trig, 0, S * sizeof(int)));
    CKH(cudaMalloc((void**)&h->d_count, sizeof(unsigned long long)));
    CKH(cudaMemset(h->d_count, 0, sizeof(unsigned long long)));
    CKH(ensure_dyn_smem(mfcc_batch_kernel<int16_t, true>, (size_t)(h->k1_batch_smem)));
    CKH(ensure_dyn_smem(mfcc_batch_kernel<int16_t, false>, (size_t)(h->k1_batch_smem)));
    CKH(ensure_dyn_smem(mfcc_batch_kernel<float, true>, (size_t)(h->k1_batch_smem)));
    CKH(ensure_dyn_smem(mfcc_batch_kernel<float, false>, (size_t)(h->k1_batch_smem)));
    CKH(ensure_dyn_smem(mfcc_fast_batch_kernel, (size_t)(h->k1_fast_smem)));
    CKH(ensure_dyn_smem(mfcc_fast_stream_kernel<false>, (size_t)(h->k1_fast_smem)));
    CKH(ensure_dyn_smem(mfcc_fast_stream_kernel<true>, (size_t)(h->k1_fast_smem)));
    CKH(ensure_dyn_smem(mfcc_stream_kernel<true>, (size_t)(h->k1_stream_smem)));
    CKH(ensure_dyn_smem(mfcc_stream_kernel<false>, (size_t)(h->k1_stream_smem)));
#undef CKH
    *out = h;
    return PB_OK;
}

PB_API int64_t pb_mfcc_frames(const pb_handle* h, int64_t n) {
    if (!h) return 0;
    return n < h->rel_window ? 0 : (n - h->rel_window) / h->cfg.hop_samples + 1;
}
PB_API int32_t pb_feature_size(const pb_handle* h) { return h ? h->feat : 0; }
PB_API int32_t pb_mfcc_width(const pb_handle* h) { return h ? h->n_out : 0; }

PB_API int pb_get_filterbank(const pb_handle* h, double* out) {
    if (!h || !out) return fail(PB_ERR_INVALID, "null argument");
    memcpy(out, h->fb.data(), h->fb.size() * sizeof(double));
    return PB_OK;
}

PB_API int64_t pb_get_cdf(const pb_handle* h, double* out, int64_t capacity, int32_t* min_out, int32_t* max_out) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (min_out) *min_out = h->min_out;
    if (max_out) *max_out = h->max_out;
    if (out) memcpy(out, h->cd.data(), std::min<int64_t>(capacity, (int64_t)h->cd.size()) * sizeof(double));
    return (int64_t)h->cd.size();
}

PB_API int pb_set_cdf(pb_handle* h, const double* cd, int64_t len) {
    if (!h || !cd) return fail(PB_ERR_INVALID, "null argument");
    if (len != (int64_t)h->cd.size()) return fail(PB_ERR_INVALID, "cdf length %lld != %zu", (long long)len, h->cd.size());
    CK(cudaSetDevice(h->cfg.device));
    memcpy(h->cd.data(), cd, len * sizeof(double));
    if (len) CK(cudaMemcpy(h->d_cd, cd, len * sizeof(double), cudaMemcpyHostToDevice));
    return PB_OK;
}

PB_API int pb_load_weights(pb_handle* h, const float* kernel, const float* recurrent, const float* bias,
                    const float* dense_w, float dense_b) {
    if (!h || !kernel || !recurrent || !bias || !dense_w) return fail(PB_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    const int H = h->cfg.hidden, F = h->feat, H3 = 3 * H;
    std::vector<float> wcat((size_t)(F + H) * H3);
    memcpy(wcat.data(), kernel, (size_t)F * H3 * sizeof(float));
    memcpy(wcat.data() + (size_t)F * H3, recurrent, (size_t)H * H3 * sizeof(float));
    cudaFree(h->d_wcat); cudaFree(h->d_bias); cudaFree(h->d_wd);
    h->d_wcat = h->d_bias = h->d_wd = nullptr;
    CK(upload(&h->d_wcat, wcat));
    CK(upload(&h->d_bias, std::vector<float>(bias, bias + H3)));
    CK(upload(&h->d_wd, std::vector<float>(dense_w, dense_w + H)));
    h->bd = dense_b;
    h->small_path = (H == 20 && F == 13 && !h->cfg.use_delta && h->cfg.activation == PB_ACT_LINEAR &&
                     h->cfg.recurrent_activation == PB_RACT_HARD_SIGMOID);
    if (h->small_path) {
        // tensor-core fragments (gru_mma_kernel): logical k slot (kt, t, j) -> input 8 kt + 2 t + j
        // (x feature for kt < 2, hidden unit 8 (kt - 2) + 2 t + j otherwise); column (nt, g) -> gate nt / 3,
        // unit 8 (nt % 3) + g.  Values are split into TF32 hi / lo parts (round to nearest, ties away).
        auto tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };
        std::vector<float4> bf((size_t)MMA_KT * MMA_NT * 32);
        for (int kt = 0; kt < MMA_KT; ++kt)
            for (int nt = 0; nt < MMA_NT; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3, gate = nt / 3, unit = 8 * (nt % 3) + g;
                    float b[2];
                    for (int j = 0; j < 2; ++j) {
                        const int k = 8 * kt + 2 * t + j;
                        float v = 0.f;
                        if (unit < H) {
                            if (kt < 2) { if (k < F) v = kernel[(size_t)k * H3 + gate * H + unit]; }
                            else { const int hu = k - 16; if (hu < H) v = recurrent[(size_t)hu * H3 + gate * H + unit]; }
                        }
                        b[j] = v;
                    }
                    const float b0h = tf32(b[0]), b1h = tf32(b[1]);
                    bf[((size_t)kt * MMA_NT + nt) * 32 + lane] = make_float4(b0h, b1h, tf32(b[0] - b0h), tf32(b[1] - b1h));
                }
        {   // fp16 fragments of the recurrent weights (gru_mma16_kernel): k-tile 0 = hidden units 0..15 as an m16n8k16 B fragment
            // (b0: k = 2t, 2t + 1; b1: k = 2t + 8, 2t + 9), k-tile 1 = units 16..23 as an m16n8k8 one (b0 only); column (nt, g) as above.
            auto h2 = [](float lo16, float hi16) {
                const __half a = __float2half_rn(lo16), b = __float2half_rn(hi16);
                uint16_t ua, ub; memcpy(&ua, &a, 2); memcpy(&ub, &b, 2);
                return (uint32_t)ua | ((uint32_t)ub << 16);
            };
            auto res = [](float v) { return v - __half2float(__float2half_rn(v)); };
            std::vector<uint4> bf16((size_t)2 * MMA_NT * 32);
            for (int kt = 0; kt < 2; ++kt)
                for (int nt = 0; nt < MMA_NT; ++nt)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int g = lane >> 2, t = lane & 3, gate = nt / 3, unit = 8 * (nt % 3) + g;
                        float b[2][2];
                        for (int r = 0; r < 2; ++r)
                            for (int j = 0; j < 2; ++j) {
                                const int hu = 16 * kt + 8 * r + 2 * t + j;
                                b[r][j] = (unit < H && hu < H && !(kt == 1 && r == 1)) ? recurrent[(size_t)hu * H3 + gate * H + unit] : 0.f;
                            }
                        bf16[((size_t)kt * MMA_NT + nt) * 32 + lane] = make_uint4(h2(b[0][0], b[0][1]), h2(b[1][0], b[1][1]),
                                                                                  h2(res(b[0][0]), res(b[0][1])), h2(res(b[1][0]), res(b[1][1])));
                    }
            cudaFree(h->d_bfrag16); h->d_bfrag16 = nullptr;
            CK(upload(&h->d_bfrag16, bf16));
            // ... and of the input weights: features 0..15 as one k16 fragment (b0: k = 2t, 2t + 1; b1: k = 2t + 8, 2t + 9)
            std::vector<uint4> xf16((size_t)MMA_NT * 32);
            for (int nt = 0; nt < MMA_NT; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3, gate = nt / 3, unit = 8 * (nt % 3) + g;
                    float b[2][2];
                    for (int r = 0; r < 2; ++r)
                        for (int j = 0; j < 2; ++j) {
                            const int f = 8 * r + 2 * t + j;
                            b[r][j] = (unit < H && f < F) ? kernel[(size_t)f * H3 + gate * H + unit] : 0.f;
                        }
                    xf16[(size_t)nt * 32 + lane] = make_uint4(h2(b[0][0], b[0][1]), h2(b[1][0], b[1][1]),
                                                              h2(res(b[0][0]), res(b[0][1])), h2(res(b[1][0]), res(b[1][1])));
                }
            cudaFree(h->d_xfrag16); h->d_xfrag16 = nullptr;
            CK(upload(&h->d_xfrag16, xf16));
            CK(ensure_dyn_smem(gru_mma16_kernel<20, 13, 4>, (size_t)K2_STAGED_SMEM));
            CK(ensure_dyn_smem(gru_mma16_kernel<20, 13, 5>, (size_t)K2_STAGED_SMEM));
        }
        {   // input projection table: wx[f][col], col = gate * 24 + unit (same column order as the accumulator tiles)
            std::vector<float> pw((size_t)F * PROJ_COLS, 0.f), pbias(PROJ_COLS, 0.f);
            for (int gate = 0; gate < 3; ++gate)
                for (int u = 0; u < H; ++u) {
                    pbias[gate * 24 + u] = bias[gate * H + u];
                    for (int f = 0; f < F; ++f) pw[(size_t)f * PROJ_COLS + gate * 24 + u] = kernel[(size_t)f * H3 + gate * H + u];
                }
            cudaFree(h->d_proj_w); cudaFree(h->d_proj_b); h->d_proj_w = h->d_proj_b = nullptr;
            CK(upload(&h->d_proj_w, pw));
            CK(upload(&h->d_proj_b, pbias));
        }
        std::vector<float> mb(72, 0.f), mw(24, 0.f);
        for (int gate = 0; gate < 3; ++gate)
            for (int u = 0; u < H; ++u) mb[gate * 24 + u] = bias[gate * H + u];
        for (int u = 0; u < H; ++u) mw[u] = dense_w[u];
        cudaFree(h->d_bfrag); cudaFree(h->d_mma_bias); cudaFree(h->d_mma_wd);
        h->d_bfrag = nullptr; h->d_mma_bias = h->d_mma_wd = nullptr;
        CK(upload(&h->d_bfrag, bf));
        CK(upload(&h->d_mma_bias, mb));
        CK(upload(&h->d_mma_wd, mw));
        {   // tcgen05 operand tiles (gru_tc5.cuh): B[k/4][n][k%4], k = x feature (0..15) then hidden unit (16..39)
            const int n1 = TC5_N1, n2 = TC5_N2, sz1 = 10 * n1 * 4, sz2 = 10 * n2 * 4;
            std::vector<float> t((size_t)2 * sz1 + 2 * sz2 + 80 + 24, 0.f);
            float *b1h = t.data(), *b1l = b1h + sz1, *b2h = b1l + sz1, *b2l = b2h + sz2, *tb = b2l + sz2, *tw = tb + 80;
            auto wv = [&](int k, int gate, int unit) -> float {
                if (unit >= H) return 0.f;
                if (k < 16) return k < F ? kernel[(size_t)k * H3 + gate * H + unit] : 0.f;
                const int hu = k - 16;
                return hu < H ? recurrent[(size_t)hu * H3 + gate * H + unit] : 0.f;
            };
            for (int k = 0; k < 40; ++k) {
                for (int c = 0; c < n1; ++c) {
                    const float v = wv(k, c / 24, c % 24), vh = tf32(v);
                    b1h[((size_t)(k / 4) * n1 + c) * 4 + k % 4] = vh;
                    b1l[((size_t)(k / 4) * n1 + c) * 4 + k % 4] = tf32(v - vh);
                }
                for (int c = 0; c < n2; ++c) {
                    const float v = wv(k, 2, c), vh = tf32(v);
                    b2h[((size_t)(k / 4) * n2 + c) * 4 + k % 4] = vh;
                    b2l[((size_t)(k / 4) * n2 + c) * 4 + k % 4] = tf32(v - vh);
                }
            }
            for (int u = 0; u < H; ++u) { tb[u] = bias[u]; tb[24 + u] = bias[H + u]; tb[48 + u] = bias[2 * H + u]; tw[u] = dense_w[u]; }
            cudaFree(h->d_tc5); h->d_tc5 = nullptr;
            CK(upload(&h->d_tc5, t));
            CK(ensure_dyn_smem(gru_mma_kernel<20, 13, true, true, 1, true>, (size_t)K2_STAGED_SMEM));
            CK(ensure_dyn_smem(gru_tc5_kernel<20, 13, true>, (size_t)(sizeof(Tc5Smem) + 128)));
            CK(ensure_dyn_smem(gru_tc5_kernel<20, 13, false>, (size_t)(sizeof(Tc5Smem) + 128)));
            CK(ensure_dyn_smem(gru_tc5_kernel<20, 13, true, true>, (size_t)(sizeof(Tc5Smem) + 128)));
        }
        memcpy(h->w_small.W, kernel, sizeof(h->w_small.W));
        memcpy(h->w_small.U, recurrent, sizeof(h->w_small.U));
        memcpy(h->w_small.b, bias, sizeof(h->w_small.b));
        memcpy(h->w_small.wd, dense_w, sizeof(h->w_small.wd));
        h->w_small.bd = dense_b;
    } else {
        h->tcb_ok = H <= TCB_HP && F <= 8 * TCB_MAX_KX && !h->cfg.use_delta;
        if (h->tcb_ok) {
            // weight tiles of gru_tcb_kernel: per k-step s one contiguous tile [hi: chunk0[N][4], chunk1[N][4] | lo: same];
            // k-step s < kx covers x features 8s..8s+7, s >= kx hidden units 8(s-kx)..; phase 1: N = 256 (z | r), phase 2: N = 128
            auto tf32 = [](float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xffffe000u; float r; memcpy(&r, &u, 4); return r; };
            const int kx = (F + 7) / 8, ks = kx + TCB_KH;
            h->tcb_kx = kx;
            std::vector<float> t((size_t)ks * 4096 + (size_t)ks * 2048 + 3 * TCB_HP + TCB_HP, 0.f);
            float* b1 = t.data(); float* b2 = b1 + (size_t)ks * 4096; float* tb = b2 + (size_t)ks * 2048; float* tw = tb + 3 * TCB_HP;
            auto wv = [&](int s, int j, int gate, int unit) -> float {
                if (unit >= H) return 0.f;
                if (s < kx) { const int f = 8 * s + j; return f < F ? kernel[(size_t)f * H3 + gate * H + unit] : 0.f; }
                const int hu = 8 * (s - kx) + j;
                return hu < H ? recurrent[(size_t)hu * H3 + gate * H + unit] : 0.f;
            };
            for (int s = 0; s < ks; ++s)
                for (int j = 0; j < 8; ++j) {
                    for (int c = 0; c < 256; ++c) {
                        const float v = wv(s, j, c / TCB_HP, c % TCB_HP), vh = tf32(v);
                        const size_t o = (size_t)s * 4096 + ((size_t)(j / 4) * 256 + c) * 4 + j % 4;
                        b1[o] = vh; b1[o + 2048] = tf32(v - vh);
                    }
                    for (int c = 0; c < 128; ++c) {
                        const float v = wv(s, j, 2, c), vh = tf32(v);
                        const size_t o = (size_t)s * 2048 + ((size_t)(j / 4) * 128 + c) * 4 + j % 4;
                        b2[o] = vh; b2[o + 1024] = tf32(v - vh);
                    }
                }
            for (int g = 0; g < 3; ++g)
                for (int u = 0; u < H; ++u) tb[g * TCB_HP + u] = bias[g * H + u];
            for (int u = 0; u < H; ++u) tw[u] = dense_w[u];
            cudaFree(h->d_tcb); h->d_tcb = nullptr;
            CK(upload(&h->d_tcb, t));
            CK(ensure_dyn_smem(gru_tcb_kernel<true>, (size_t)(sizeof(TcbSmem) + 128)));
            CK(ensure_dyn_smem(gru_tcb_kernel<false>, (size_t)(sizeof(TcbSmem) + 128)));
        }
        size_t smem = (size_t)(F + 3 * H) * K2_TILE_STREAMS * sizeof(float);
        if (smem > 200 * 1024) return fail(PB_ERR_UNSUPPORTED, "feature_size + 3*hidden = %d is too large for the tiled GRU kernel", F + 3 * H);
        CK(ensure_dyn_smem(gru_tiled_kernel<true>, (size_t)(smem)));
        CK(ensure_dyn_smem(gru_tiled_kernel<false>, (size_t)(smem)));
    }
    h->have_weights = true;
    h->proj_dirty = true;
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
// profiling helpers
struct ProfScope {
    pb_handle* h; int slot; cudaStream_t s; int idx = -1;
    ProfScope(pb_handle* h_, int slot_, cudaStream_t s_) : h(h_), slot(slot_), s(s_) {
        if (!h->profiling) return;
        ProfSlot& p = h->prof[slot];
        if (p.used + 2 > (int)p.ev.size()) {
            if ((int)p.ev.size() >= 2 * PROFILE_POOL) {          // pool full: fold what we have
                for (int i = 0; i + 1 < p.used; i += 2) {
                    cudaEventSynchronize(p.ev[i + 1]);
                    float ms = 0; cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]); p.ms += ms;
                }
                p.used = 0;
            } else {
                cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); p.ev.push_back(a); p.ev.push_back(b);
            }
        }
        idx = p.used; p.used += 2; p.launches++;
        cudaEventRecord(p.ev[idx], s);
    }
    ~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof[slot].ev[idx + 1], s); }
};

PB_API int pb_debug_counters(pb_handle* h, long long out[4]) {
    if (!h || !out) return fail(PB_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->cfg.device));
    if (!h->d_dbg) { CK(cudaMalloc((void**)&h->d_dbg, 4 * sizeof(long long))); CK(cudaMemset(h->d_dbg, 0, 4 * sizeof(long long))); }
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, h->d_dbg, 4 * sizeof(long long), cudaMemcpyDeviceToHost));
    return PB_OK;
}

PB_API int pb_debug_gru_mode(pb_handle* h, int mode) { if (!h) return fail(PB_ERR_INVALID, "null handle"); h->gru_mode = mode; return PB_OK; }

PB_API int pb_debug_k1_mode(pb_handle* h, int mode) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (mode >= 100 && mode <= 116 && h->tc3_ok) { h->k1_mode = mode; return PB_OK; }      // mode 5 with the phase timeline of warp (mode - 100) in pb_debug_counters
    if (mode < 0 || mode > 6 || mode == 1) return fail(PB_ERR_INVALID, "k1 mode must be 0 (automatic), 2 (FFT kernel, lean set-up), 3 (FFT kernel), 4 (tensor-core DFT kernel, stage 2 only) or 5 (both DFT stages on the tensor cores)");
    if (mode >= 5 && mode <= 6 && !h->tc3_ok) return fail(PB_ERR_UNSUPPORTED, "the two-stage tensor-core MFCC tick needs the default mel geometry, hop >= 512 and chunk >= hop, a multiple of 8");
    if (mode == 4 && !h->tc2_ok) return fail(PB_ERR_UNSUPPORTED, "the tensor-core MFCC tick needs the default mel geometry (20 filters, 16 kHz, n_fft 512), chunk >= 512 and a multiple of 8");
    if (mode == 2 && !h->fast_ok) return fail(PB_ERR_UNSUPPORTED, "k1 mode 2 needs the aligned geometry of the fast MFCC kernels");
    h->k1_mode = mode;
    return PB_OK;
}

// CPU model of the tensor-core DFT for one 512-sample frame (mfcc_tc.cuh: same butterfly, same operand tables and layout
// arithmetic as the kernel).  No device needed; used by the CPU tests to pin the host-side half of that design.
PB_API int pb_debug_tc_dft_power(const int16_t* x512, double* power257) {
    if (!x512 || !power257) return fail(PB_ERR_INVALID, "null argument");
    tcd_host_power(x512, power257);
    return PB_OK;
}

static void tcd_host_tables(const pb_handle* h, std::vector<float4>& etab, std::vector<float>& dct, float* tot_scale);

// CPU model of the whole experimental kernel for one frame: accumulator row (as above) + the epilogue (mel, log, DCT, c0) with
// the tables a handle of this configuration would upload.  Needs no device: only the mel-table part of pb_create runs.
PB_API int pb_debug_tc_mfcc_frame(const pb_config* cfg, const int16_t* x512, float* out) {
    if (!cfg || !x512 || !out) return fail(PB_ERR_INVALID, "null argument");
    if (cfg->n_fft != 512 || cfg->n_filt < 1 || cfg->n_filt > TCD_MAX_FILT || cfg->vectorizer != PB_VEC_MFCCS)
        return fail(PB_ERR_UNSUPPORTED, "the tensor-core MFCC model covers n_fft 512, n_filt <= %d, MFCC vectorizer", TCD_MAX_FILT);
    pb_handle* h = new (std::nothrow) pb_handle();
    if (!h) return fail(PB_ERR_CUDA, "out of host memory");
    h->cfg = *cfg;
    h->n_bins = cfg->n_fft / 2 + 1;
    h->n_out = std::min(cfg->n_filt, cfg->n_mfcc);
    int rc = h->n_out > TCD_MAX_OUT ? fail(PB_ERR_UNSUPPORTED, "n_mfcc > %d", TCD_MAX_OUT) : build_mel(h, h->h_wrise, h->h_wfall, h->h_grid);
    if (rc == PB_OK) {
        std::vector<float4> etab;
        std::vector<float> dct;
        float tot_scale = 0.f;
        tcd_host_tables(h, etab, dct, &tot_scale);
        float d[TCD_BLOCKS][64];
        tcd_host_accumulators(x512, d);
        tcd_host_epilogue(d, etab.data(), dct.data(), cfg->n_filt, h->n_out, tot_scale, out);
    }
    delete h;
    return rc;
}

// ... and of the kernel with both DFT stages on the tensor cores (mfcc_tc3.cuh): exact int16 split, stage-1 matrix passes, twiddle,
// fp16 split, stage-2 passes, its own epilogue order.  No device needed.  Test hook.
PB_API int pb_debug_tc3_mfcc_frame(const pb_config* cfg, const int16_t* x512, float* out, double* power257) {
    if (!cfg || !x512 || !out) return fail(PB_ERR_INVALID, "null argument");
    if (cfg->n_fft != 512 || cfg->n_filt < 1 || cfg->n_filt > TCD_MAX_FILT || cfg->vectorizer != PB_VEC_MFCCS)
        return fail(PB_ERR_UNSUPPORTED, "the tensor-core MFCC model covers n_fft 512, n_filt <= %d, MFCC vectorizer", TCD_MAX_FILT);
    pb_handle* h = new (std::nothrow) pb_handle();
    if (!h) return fail(PB_ERR_CUDA, "out of host memory");
    h->cfg = *cfg;
    h->n_bins = cfg->n_fft / 2 + 1;
    h->n_out = std::min(cfg->n_filt, cfg->n_mfcc);
    int rc = h->n_out > TCD_MAX_OUT ? fail(PB_ERR_UNSUPPORTED, "n_mfcc > %d", TCD_MAX_OUT) : build_mel(h, h->h_wrise, h->h_wfall, h->h_grid);
    if (rc == PB_OK) {
        std::vector<float4> etab;
        std::vector<float> dct;
        float tot_scale = 0.f;
        tcd_host_tables(h, etab, dct, &tot_scale);
        float d[TCD_BLOCKS][64];
        tc3_host_accumulators(x512, d);
        tc3_host_epilogue(d, h->h_wrise, h->h_wfall, h->h_grid, dct.data(), cfg->n_filt, h->n_out, tot_scale, out);
        if (power257) {
            const double inv = 1.0 / ((double)TCD_A_SCALE * (double)TCD_A_SCALE);
            for (int b = 0; b < TCD_BLOCKS; ++b)
                for (int half = 0; half < 2; ++half)
                    for (int m = 0; m < 16; ++m) {
                        const int k = tcd_col_bin(b, 32 * half + m);
                        if (k < 0) continue;
                        const double re = d[b][32 * half + m], im = d[b][32 * half + 16 + m];
                        power257[k] = (k == 0 || k == 256) ? re * re * inv : (re * re + im * im) * inv;
                    }
        }
    }
    delete h;
    return rc;
}

PB_API int pb_debug_force_generic(pb_handle* h, int on) { if (!h) return fail(PB_ERR_INVALID, "null handle"); h->force_generic = on != 0; return PB_OK; }

PB_API int pb_profile_enable(pb_handle* h, int on) { if (!h) return fail(PB_ERR_INVALID, "null handle"); h->profiling = on != 0; return PB_OK; }

PB_API int pb_profile_reset(pb_handle* h) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    for (auto& p : h->prof) { p.used = 0; p.ms = 0; p.launches = 0; }
    return PB_OK;
}

PB_API int pb_profile_read(pb_handle* h, double ms[4], uint64_t launches[4]) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    CK(cudaSetDevice(h->cfg.device));
    for (int s = 0; s < N_PROFILE_SLOTS; ++s) {
        ProfSlot& p = h->prof[s];
        for (int i = 0; i + 1 < p.used; i += 2) {
            CK(cudaEventSynchronize(p.ev[i + 1]));
            float t = 0; CK(cudaEventElapsedTime(&t, p.ev[i], p.ev[i + 1])); p.ms += t;
        }
        p.used = 0;
        if (ms) ms[s] = p.ms;
        if (launches) launches[s] = p.launches;
    }
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
static MelTables mel_tables(const pb_handle* h) {
    MelTables t;
    t.w_rise = h->d_wrise; t.w_fall = h->d_wfall; t.grid = h->d_grid; t.dct = h->d_dct;
    t.tw_stage = h->d_tw_stage; t.tw_post = h->d_tw_post;
    t.n_bins = h->n_bins; t.n_filt = h->cfg.n_filt; t.n_out = h->n_out;
    t.mels_only = h->cfg.vectorizer == PB_VEC_MELS;
    t.n_fft = h->cfg.n_fft; t.tw_any = h->d_tw_any;
    return t;
}

static FastTables fast_tables(const pb_handle* h) {
    FastTables f;
    f.ptab = h->d_ptab; f.ctab = h->d_ctab; f.dct_t = h->d_dct_t;
    f.npl = h->npl; f.maxc = h->maxc; f.nol = h->nol;
    return f;
}

static DecodeParams decode_params(const pb_handle* h) {
    DecodeParams d;
    d.cd = h->d_cd; d.cd_len = (int)h->cd.size();
    d.min_out = h->min_out; d.out_range = h->max_out - h->min_out;
    d.center = h->cfg.threshold_center;
    d.hot_threshold = 1.0 - h->cfg.sensitivity;
    d.trigger_level = h->cfg.trigger_level;
    const long long bytes = 2LL * h->cfg.chunk_samples;          // TriggerDetector.chunk_size is in bytes
    long long q = -(8 * 2048) / bytes;                            // C truncates toward zero ...
    if ((-(8 * 2048)) % bytes != 0) q -= 1;                       // ... python floors
    d.trigger_reset = (int)q;
    d.legacy_f64 = h->cfg.decode_legacy_f64 != 0;
    return d;
}

template <typename T>
static int mfcc_impl(pb_handle* h, const T* d_in, int64_t n_streams, int64_t L, float* d_out, cudaStream_t s, float scale) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (n_streams < 0 || L < 0) return fail(PB_ERR_INVALID, "negative size");
    if (L == 0 && n_streams > 0) return fail(PB_ERR_INVALID, "Cannot vectorize empty audio!");   // vectorization.py:48-49
    const int64_t nf = pb_mfcc_frames(h, L), total = nf * n_streams;
    if (total == 0) return PB_OK;
    if (!d_in || !d_out) return fail(PB_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    const bool pairs = (L % 2 == 0) && (h->cfg.hop_samples % 2 == 0) && (h->used % 2 == 0) &&
                       ((uintptr_t)d_in % (2 * sizeof(T)) == 0);
    const int64_t tiles = (total + K1_TILE - 1) / K1_TILE;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 4);
    ProfScope ps(h, 0, s);
    if (sizeof(T) == 2 && h->fast_ok && L % 8 == 0 && (uintptr_t)d_in % 16 == 0 && !h->force_generic) {
        const int64_t pairs2 = (total + 1) / 2;
        const int gridf = (int)std::min<int64_t>((pairs2 + K1F_WARPS - 1) / K1F_WARPS, (int64_t)h->sm_count * 4);
        mfcc_fast_batch_kernel<<<gridf, K1F_THREADS, h->k1_fast_smem, s>>>((const int16_t*)d_in, L, nf, total, h->cfg.hop_samples, scale,
                                                                           mel_tables(h), fast_tables(h), d_out);
    } else if (pairs)
        mfcc_batch_kernel<T, true><<<grid, K1_THREADS, h->k1_batch_smem, s>>>(d_in, L, nf, total, h->cfg.hop_samples, h->used, scale, mel_tables(h), d_out);
    else
        mfcc_batch_kernel<T, false><<<grid, K1_THREADS, h->k1_batch_smem, s>>>(d_in, L, nf, total, h->cfg.hop_samples, h->used, scale, mel_tables(h), d_out);
    CK(cudaGetLastError());
    return PB_OK;
}

PB_API int pb_mfcc(pb_handle* h, const int16_t* d_pcm, int64_t n_streams, int64_t L, float* d_out, void* stream) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    const float inv = 1.0f / 32768.0f;
    return mfcc_impl<int16_t>(h, d_pcm, n_streams, L, d_out, (cudaStream_t)stream, inv * inv / (float)h->cfg.n_fft);
}

PB_API int pb_mfcc_f32(pb_handle* h, const float* d_audio, int64_t n_streams, int64_t L, float* d_out, void* stream) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    return mfcc_impl<float>(h, d_audio, n_streams, L, d_out, (cudaStream_t)stream, 1.0f / (float)h->cfg.n_fft);
}

static int launch_gru(pb_handle* h, const K2In& in, bool ring, int64_t n, const DecodeParams& dp, const K2Out& o, cudaStream_t s) {
    ProfScope ps(h, 1, s);
    if (h->small_path && n <= K2_WARP_PATH_MAX && h->gru_mode == 0) {                 // latency path: a warp per stream
        const int grid = (int)((n + 3) / 4);
        if (ring) gru_warp_kernel<20, 13, true><<<grid, 128, 0, s>>>(h->w_small, in, n, dp, o);
        else gru_warp_kernel<20, 13, false><<<grid, 128, 0, s>>>(h->w_small, in, n, dp, o);
    } else if (h->small_path && (h->gru_mode == 3 || h->gru_mode == 8)) {   // tcgen05 + TMEM scan (8: over cached projections, opt-in)
        GruTc5W w;
        const int sz1 = 10 * TC5_N1 * 4, sz2 = 10 * TC5_N2 * 4;
        w.b1_hi = h->d_tc5; w.b1_lo = w.b1_hi + sz1; w.b2_hi = w.b1_lo + sz1; w.b2_lo = w.b2_hi + sz2;
        w.bias = w.b2_lo + sz2; w.wd = w.bias + 80; w.bd = h->bd;
        const int grid = (int)((n + TC5_THREADS - 1) / TC5_THREADS);
        const size_t smem = sizeof(Tc5Smem) + 128;
        if (ring && in.proj != nullptr && h->gru_mode == 8) gru_tc5_kernel<20, 13, true, true><<<grid, TC5_BLOCK, smem, s>>>(w, in, n, dp, o);
        else if (ring) gru_tc5_kernel<20, 13, true><<<grid, TC5_BLOCK, smem, s>>>(w, in, n, dp, o);
        else gru_tc5_kernel<20, 13, false><<<grid, TC5_BLOCK, smem, s>>>(w, in, n, dp, o);
    } else if (h->small_path && h->gru_mode != 1) {               // tensor-core scan (mma.sync TF32 x3)
        GruMmaW w;
        w.bfrag = h->d_bfrag; w.bias = h->d_mma_bias; w.wd = h->d_mma_wd; w.bd = h->bd;
        const int per_cta = (MMA_THREADS / 32) * 16 * MMA_MB;
        const int grid = (int)((n + per_cta - 1) / per_cta);
        // Steady-state stream scan: 16-stream tiles (MB = 1) unless forced (gru_mode 7).  With 32-stream tiles 131 072 streams
        // are 2.3 rounds of 3 CTAs/SM and the last, 31 % full round still costs most of a round (each warp's step is a
        // dependent chain); 16-stream tiles fit 5 CTAs/SM and leave a much shorter tail.
        if (ring && in.proj != nullptr && h->gru_mode != 7) {
            const int per1 = (MMA_THREADS / 32) * 16;
            if (h->gru_mode == 9)            // A/B: 3xTF32 scan without bulk-copy staging of the projection blocks
                gru_mma_kernel<20, 13, true, true, 1><<<(int)((n + per1 - 1) / per1), MMA_THREADS, 0, s>>>(w, in, n, dp, o);
            else if (h->gru_mode == 10)      // A/B: 3xTF32 scan with staging
                gru_mma_kernel<20, 13, true, true, 1, true><<<(int)((n + per1 - 1) / per1), MMA_THREADS, K2_STAGED_SMEM, s>>>(w, in, n, dp, o);
            else {                           // default: fp16x3 recurrent products (half the tensor-pipe time), staged projection blocks
                GruMma16W w16;
                w16.bfrag = h->d_bfrag16; w16.xfrag = h->d_xfrag16; w16.bias = h->d_mma_bias; w16.wd = h->d_mma_wd; w16.bd = h->bd;
                if (h->gru_mode == 11)       // A/B: 5 CTAs per SM (96 registers, a small spill): 215 vs 211 us
                    gru_mma16_kernel<20, 13, 5><<<(int)((n + per1 - 1) / per1), MMA_THREADS, K2_STAGED_SMEM, s>>>(w16, in, n, dp, o);
                else                         // 4 CTAs per SM, 128 registers
                    gru_mma16_kernel<20, 13, 4><<<(int)((n + per1 - 1) / per1), MMA_THREADS, K2_STAGED_SMEM, s>>>(w16, in, n, dp, o);
            }
        } else if (ring && in.proj != nullptr) gru_mma_kernel<20, 13, true, true><<<grid, MMA_THREADS, 0, s>>>(w, in, n, dp, o);
        else if (ring) gru_mma_kernel<20, 13, true, false><<<grid, MMA_THREADS, 0, s>>>(w, in, n, dp, o);
        else gru_mma_kernel<20, 13, false, false><<<grid, MMA_THREADS, 0, s>>>(w, in, n, dp, o);
    } else if (h->small_path) {
        const int per_cta = K2_SMALL_THREADS * K2_NS;
        const int grid = (int)((n + per_cta - 1) / per_cta);
        if (ring) gru_small_kernel<20, 13, true><<<grid, K2_SMALL_THREADS, 0, s>>>(h->w_small, in, n, dp, o);
        else gru_small_kernel<20, 13, false><<<grid, K2_SMALL_THREADS, 0, s>>>(h->w_small, in, n, dp, o);
    } else if (h->tcb_ok && h->gru_mode != 1) {                    // wide network on tcgen05 + TMEM, weights via TMA pipeline
        GruTcbW w;
        const int ks = h->tcb_kx + TCB_KH;
        w.b1 = h->d_tcb; w.b2 = w.b1 + (size_t)ks * 4096; w.bias = w.b2 + (size_t)ks * 2048; w.wd = w.bias + 3 * TCB_HP;
        w.bd = h->bd; w.kx = h->tcb_kx; w.F = h->feat; w.H = h->cfg.hidden; w.act = h->cfg.activation; w.ract = h->cfg.recurrent_activation;
        w.dbg = h->d_dbg;
        const int grid = (int)((n + 127) / 128);
        const size_t smem = sizeof(TcbSmem) + 128;
        if (ring) gru_tcb_kernel<true><<<grid, TCB_BLOCK, smem, s>>>(w, in, n, dp, o);
        else gru_tcb_kernel<false><<<grid, TCB_BLOCK, smem, s>>>(w, in, n, dp, o);
    } else {
        GruTiledW w;
        w.wcat = h->d_wcat; w.bias = h->d_bias; w.wd = h->d_wd; w.bd = h->bd;
        w.H = h->cfg.hidden; w.F_in = h->feat; w.act = h->cfg.activation; w.ract = h->cfg.recurrent_activation;
        const size_t smem = (size_t)(w.F_in + 3 * w.H) * K2_TILE_STREAMS * sizeof(float);
        const int grid = (int)((n + K2_TILE_STREAMS - 1) / K2_TILE_STREAMS);
        if (ring) gru_tiled_kernel<true><<<grid, K2_TILE_THREADS, smem, s>>>(w, in, n, dp, o);
        else gru_tiled_kernel<false><<<grid, K2_TILE_THREADS, smem, s>>>(w, in, n, dp, o);
    }
    CK(cudaGetLastError());
    return PB_OK;
}

PB_API int pb_predict(pb_handle* h, const float* d_inputs, int64_t n, float* d_out, float* d_logit, void* stream) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (!h->have_weights) return fail(PB_ERR_STATE, "pb_load_weights has not been called");
    if (n < 0) return fail(PB_ERR_INVALID, "negative n");
    if (n == 0) return PB_OK;
    if (!d_inputs || !d_out) return fail(PB_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    K2In in{};
    in.inputs = d_inputs; in.T = h->cfg.n_features; in.F_base = h->n_out; in.use_delta = 0;
    K2Out o{};
    o.raw = d_out; o.logit = d_logit;
    return launch_gru(h, in, false, n, decode_params(h), o, (cudaStream_t)stream);
}

PB_API int pb_decode(pb_handle* h, const float* d_raw, int64_t n, double* d_conf, void* stream) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (n < 0) return fail(PB_ERR_INVALID, "negative n");
    if (n == 0) return PB_OK;
    if (!d_raw || !d_conf) return fail(PB_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(h, 2, s);
    decode_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(d_raw, n, decode_params(h), d_conf);
    CK(cudaGetLastError());
    return PB_OK;
}

static int check_tick(pb_handle* h, const void* pcm, int64_t n) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (n < 0 || n > h->cfg.max_streams) return fail(PB_ERR_INVALID, "n = %lld outside [0, max_streams = %d]", (long long)n, h->cfg.max_streams);
    if (n > 0 && !pcm) return fail(PB_ERR_INVALID, "null pcm");
    return PB_OK;
}

// tables of the experimental tensor-core MFCC tick (host part: shared by the device upload and the CPU model)
static void tcd_host_tables(const pb_handle* h, std::vector<float4>& etab, std::vector<float>& dct, float* tot_scale) {
    const float inv = 1.0f / 32768.0f, scale = inv * inv / (float)h->cfg.n_fft, pscale = scale / (TCD_A_SCALE * TCD_A_SCALE);
    tcd_build_etab(etab, h->h_wrise, h->h_wfall, h->h_grid, h->cfg.n_filt, pscale);
    dct.assign((size_t)TCD_MAX_OUT * 24, 0.f);
    for (int k = 0; k < h->n_out; ++k)
        for (int j = 0; j < h->cfg.n_filt; ++j) {
            double v = cos(M_PI * k * (2 * j + 1) / (2.0 * h->cfg.n_filt)) * sqrt(2.0 / h->cfg.n_filt);
            if (k == 0) v *= sqrt(0.5);
            dct[(size_t)k * 24 + j] = (float)v;
        }
    *tot_scale = pscale;
}

// ... built on first use (never on the default path)
static int ensure_tcd_tables(pb_handle* h) {
    if (h->d_tcd_b) return PB_OK;
    std::vector<__half> bh, bl;
    tcd_build_b(bh, bl);
    std::vector<__half> both(bh);
    both.insert(both.end(), bl.begin(), bl.end());
    std::vector<float4> etab;
    std::vector<float> dct, tw;
    tcd_host_tables(h, etab, dct, &h->tcd_tot_scale);
    tcd_build_tw(tw);
    CK(cudaMalloc((void**)&h->d_tcd_b, both.size() * sizeof(__half)));
    CK(cudaMemcpy(h->d_tcd_b, both.data(), both.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CK(upload(&h->d_tcd_tw, tw));
    CK(upload(&h->d_tcd_dct, dct));
    CK(ensure_dyn_smem(mfcc_tc2_stream_kernel<Tc2Geo20>, sizeof(Tc2Smem) + 128));
    return PB_OK;
}

static int ensure_tc3_tables(pb_handle* h) {
    if (h->d_tc3_b1) return PB_OK;
    int rc = ensure_tcd_tables(h);                       // DCT table and the power scale are shared with mfcc_tc2
    if (rc != PB_OK) return rc;
    std::vector<__half> b1, b2;
    std::vector<float> tw;
    tc3_build_b1(b1); tc3_build_b2(b2); tc3_build_tw(tw);
    CK(cudaMalloc((void**)&h->d_tc3_b2, b2.size() * sizeof(__half)));
    CK(cudaMemcpy(h->d_tc3_b2, b2.data(), b2.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CK(upload(&h->d_tc3_tw, tw));
    CK(cudaMalloc((void**)&h->d_tc3_recs, (size_t)h->cfg.max_streams * (size_t)std::max(1, h->max_new) * sizeof(Tc3Rec)));
    CK(cudaMalloc((void**)&h->d_tc3_counters, 2 * sizeof(unsigned int)));
    CK(cudaMemset(h->d_tc3_counters, 0, 2 * sizeof(unsigned int)));
    CK(ensure_dyn_smem(mfcc_tc3_kernel<Tc2Geo20, false>, sizeof(Tc3Smem) + 128));
    CK(ensure_dyn_smem(mfcc_tc3_kernel<Tc2Geo20, true>, sizeof(Tc3Smem) + 128));
    CK(cudaMalloc((void**)&h->d_tc3_b1, b1.size() * sizeof(__half)));      // last: its presence marks the set as complete
    CK(cudaMemcpy(h->d_tc3_b1, b1.data(), b1.size() * sizeof(__half), cudaMemcpyHostToDevice));
    return PB_OK;
}

static int launch_stream_mfcc(pb_handle* h, const int16_t* d_pcm, const int32_t* d_ids, int64_t n, cudaStream_t s) {
    const bool pairs = (h->cfg.chunk_samples % 2 == 0) && (h->cfg.hop_samples % 2 == 0) && (h->used % 2 == 0) &&
                       ((uintptr_t)d_pcm % 4 == 0);
    const int64_t tiles = (n + K1_STREAMS_PER_CTA - 1) / K1_STREAMS_PER_CTA;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 4);
    const float inv = 1.0f / 32768.0f, scale = inv * inv / (float)h->cfg.n_fft;
    ProfScope ps(h, 0, s);
    // Default choice (k1_mode 0): from TC3_MIN_STREAMS streams on, the tick's MFCCs come from the kernel with both DFT stages on the
    // tensor cores (measured on B200: 116 vs 128 us at 65 536 streams, 208 vs 228 us at 131 072, 392 vs 414 us at 262 144; below
    // that its persistent pipeline does not fill and the FFT kernel wins: 71 vs 64 us at 32 768).
    const bool tc3_auto = h->k1_mode == 0 && n >= TC3_MIN_STREAMS;
    if ((h->k1_mode == 5 || h->k1_mode == 6 || h->k1_mode >= 100 || tc3_auto) && h->tc3_ok && (uintptr_t)d_pcm % 16 == 0 && !h->force_generic) {
        int rc = ensure_tc3_tables(h);
        if (rc != PB_OK) return rc;
        Tc3Tables t;
        t.b1 = h->d_tc3_b1; t.b2 = h->d_tc3_b2; t.tw = h->d_tc3_tw; t.dct = h->d_tcd_dct; t.n_out = h->n_out; t.pscale = h->tcd_tot_scale;
        const int par = h->tc3_parity;
        h->tc3_parity ^= 1;
        mfcc_tc3_plan_kernel<<<(int)((n + TC3_PLAN_THREADS - 1) / TC3_PLAN_THREADS), TC3_PLAN_THREADS, 0, s>>>(
            d_pcm, d_ids, (int)n, h->cfg.chunk_samples, h->cfg.hop_samples, h->st, h->d_tc3_recs, h->d_tc3_counters, par);
        const int64_t max_tiles = (n * std::max(1, h->max_new) + TC3_TILE - 1) / TC3_TILE;
        const int g3 = (int)std::min<int64_t>(max_tiles, h->sm_count);
        if (h->k1_mode >= 100)          // phase timeline of one warp (pb_debug_counters): separate instantiation, the counters cost registers
            mfcc_tc3_kernel<Tc2Geo20, true><<<g3, TC3_THREADS, sizeof(Tc3Smem) + 128, s>>>(t, h->d_tc3_recs, h->d_tc3_counters, par, h->k1_mode, h->d_dbg);
        else
            mfcc_tc3_kernel<Tc2Geo20, false><<<g3, TC3_THREADS, sizeof(Tc3Smem) + 128, s>>>(t, h->d_tc3_recs, h->d_tc3_counters, par, h->k1_mode == 6 ? 8 : 0, nullptr);   // 6: A/B of the epilogue's shuffle-gather tail
    } else if (h->k1_mode == 4 && h->tc2_ok && (uintptr_t)d_pcm % 16 == 0 && !h->force_generic) {
        int rc = ensure_tcd_tables(h);
        if (rc != PB_OK) return rc;
        Tc2Tables t;
        t.b = h->d_tcd_b; t.tw = h->d_tcd_tw; t.dct = h->d_tcd_dct; t.n_out = h->n_out; t.pscale = h->tcd_tot_scale;
        // super-groups: one per SM where the batch allows it, a multiple of 32 streams, at most TC2_SG_MAX
        int sg = (int)((n + h->sm_count - 1) / h->sm_count);
        sg = std::min(TC2_SG_MAX, std::max(128, (sg + 31) & ~31));
        const int groups = (int)((n + sg - 1) / sg);
        mfcc_tc2_stream_kernel<Tc2Geo20><<<std::min(groups, h->sm_count), TC2_THREADS, sizeof(Tc2Smem) + 128, s>>>(
            d_pcm, d_ids, (int)n, sg, h->cfg.chunk_samples, h->cfg.hop_samples, t, h->st);
    } else if (h->fast_ok && h->max_new <= 8 && h->cfg.chunk_samples % 8 == 0 && (uintptr_t)d_pcm % 16 == 0 && !h->force_generic) {
        // streams per warp tile: 16 at scale; fewer when the batch cannot fill the machine's warps
        const int64_t warps_total = (int64_t)h->sm_count * 4 * K1F_WARPS;
        const int spw = (int)std::max<int64_t>(1, std::min<int64_t>(K1F_STREAMS_PER_WARP, (n + warps_total - 1) / warps_total));
        const int64_t tilesf = (n + spw - 1) / spw;
        const int gridf = (int)std::min<int64_t>((tilesf + K1F_WARPS - 1) / K1F_WARPS, (int64_t)h->sm_count * 4);
        if (h->k1_mode != 3)                     // default: the 32-bit per-pass set-up (bit-identical rows, 3 % faster on B200); 3 = the 64-bit original
            mfcc_fast_stream_kernel<true><<<gridf, K1F_THREADS, h->k1_fast_smem, s>>>(d_pcm, d_ids, (int)n, h->cfg.chunk_samples, h->cfg.hop_samples, spw, scale,
                                                                                      mel_tables(h), fast_tables(h), h->st);
        else
            mfcc_fast_stream_kernel<false><<<gridf, K1F_THREADS, h->k1_fast_smem, s>>>(d_pcm, d_ids, (int)n, h->cfg.chunk_samples, h->cfg.hop_samples, spw, scale,
                                                                                       mel_tables(h), fast_tables(h), h->st);
    } else if (h->max_new > 8) {
        // A chunk that completes more than 8 frames per stream: consecutive sub-chunks of at most 6 hops (<= 8 frames each) through the
        // generic kernel -- to the state machine they are separate ticks (Listener.update_vectors is chunking-independent); the network
        // runs once, after the last one.
        const int sub = std::max(1, 6 * h->cfg.hop_samples);
        for (int off = 0; off < h->cfg.chunk_samples; off += sub)
            mfcc_stream_kernel<false><<<grid, K1_THREADS, h->k1_stream_smem, s>>>(d_pcm + off, d_ids, (int)n, std::min(sub, h->cfg.chunk_samples - off),
                                                                                    h->cfg.chunk_samples, h->cfg.hop_samples, h->used, scale, mel_tables(h), h->st);
    } else if (pairs)
        mfcc_stream_kernel<true><<<grid, K1_THREADS, h->k1_stream_smem, s>>>(d_pcm, d_ids, (int)n, h->cfg.chunk_samples, h->cfg.chunk_samples, h->cfg.hop_samples, h->used, scale, mel_tables(h), h->st);
    else
        mfcc_stream_kernel<false><<<grid, K1_THREADS, h->k1_stream_smem, s>>>(d_pcm, d_ids, (int)n, h->cfg.chunk_samples, h->cfg.chunk_samples, h->cfg.hop_samples, h->used, scale, mel_tables(h), h->st);
    CK(cudaGetLastError());
    return PB_OK;
}

PB_API int pb_update_vectors(pb_handle* h, const int16_t* d_pcm, const int32_t* d_ids, int64_t n, void* stream) {
    int rc = check_tick(h, d_pcm, n);
    if (rc != PB_OK || n == 0) return rc;
    CK(cudaSetDevice(h->cfg.device));
    h->proj_dirty = true;
    return launch_stream_mfcc(h, d_pcm, d_ids, n, (cudaStream_t)stream);
}

// Does a tick of n streams run the scan that reads cached input projections (gru_mma_kernel<.., PROJ>)?
static bool wants_projection(const pb_handle* h, int64_t n) {
    return h->has_proj && h->small_path && n > K2_WARP_PATH_MAX && (h->gru_mode == 0 || h->gru_mode == 2 || h->gru_mode == 7 || h->gru_mode == 8 || h->gru_mode == 9 || h->gru_mode == 10 || h->gru_mode == 11);
}

// Recompute the projection of every ring row once (all streams), then the cache is maintained incrementally.
static int rebuild_projections_if_dirty(pb_handle* h, cudaStream_t s) {
    if (!h->proj_dirty) return PB_OK;
    ProfScope ps(h, 3, s);
    const long long rows = (long long)h->cfg.max_streams * h->ring_rows;
    const int grid = (int)std::min<long long>((rows + PROJ_FRAMES_PER_CTA - 1) / PROJ_FRAMES_PER_CTA, (long long)h->sm_count * 16);
    input_proj_all_kernel<13><<<grid, 64 * PROJ_FRAMES_PER_CTA, 0, s>>>(h->d_proj_w, h->d_proj_b, rows, h->st.ring, h->ring_rows, h->row_stride, h->d_proj_ring,
                                                                       (h->cfg.max_streams + 15) / 16);
    CK(cudaGetLastError());
    h->proj_dirty = false;
    return PB_OK;
}

PB_API int pb_update(pb_handle* h, const int16_t* d_pcm, const int32_t* d_ids, int64_t n, float* d_raw, double* d_conf,
              uint8_t* d_fired, unsigned long long* d_count, void* stream) {
    int rc = check_tick(h, d_pcm, n);
    if (rc != PB_OK || n == 0) return rc;
    if (!h->have_weights) return fail(PB_ERR_STATE, "pb_load_weights has not been called");
    if (!d_conf) return fail(PB_ERR_INVALID, "null d_conf");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    // Default network, large batch: the tensor-core scan reads cached input projections.  A stale cache (weights changed,
    // or ticks that skipped the projection) is rebuilt for every existing row BEFORE this tick's MFCC kernel runs -- the
    // host-buffer path calls this up front on its first pipe, see pb_update_host -- and the rows the tick adds are projected
    // right after it, so the rebuild never reads or writes a row that another sub-batch of the same tick is producing.
    const bool want_proj = wants_projection(h, n);
    if (want_proj) { rc = rebuild_projections_if_dirty(h, s); if (rc != PB_OK) return rc; }
    rc = launch_stream_mfcc(h, d_pcm, d_ids, n, s);
    if (rc != PB_OK) return rc;
    // The default scan (gru_mma16_kernel, gru_mode 0) projects the frames a tick has added itself, in its prologue; the separate
    // projection kernel serves the other scans (A/B modes) and short sub-batches of a large host tick (second case below), whose
    // warp-per-stream kernel does not touch the cache.
    const bool scan_projects = want_proj && (h->gru_mode == 0 || h->gru_mode == 11);
    if ((want_proj && !scan_projects) || (!want_proj && h->host_tick_proj && !h->proj_dirty)) {
        ProfScope ps(h, 3, s);
        const long long items = (long long)n * h->max_new;
        const int grid = (int)((items + PROJ_THREADS - 1) / PROJ_THREADS);     // 32 frames per warp
        input_proj_kernel<13><<<grid, PROJ_THREADS, 0, s>>>(h->d_bfrag, h->d_proj_b, h->st.n_samples, d_ids, (int)n,
            h->cfg.chunk_samples, h->used, h->cfg.hop_samples, h->max_new, h->st.ring, h->ring_rows, h->row_stride, h->d_proj_ring,
            (h->cfg.max_streams + 15) / 16);
        CK(cudaGetLastError());
    } else if (!scan_projects && h->has_proj && h->small_path) {
        h->proj_dirty = true;                                  // this tick's frames get no projection
    }
    const bool use_proj = want_proj;
    K2In in{};
    in.ring = h->st.ring; in.n_samples = h->st.n_samples; in.ids = d_ids;
    in.ring_rows = h->ring_rows; in.row_stride = h->row_stride; in.window = h->rel_window; in.hop = h->cfg.hop_samples;
    in.T = h->cfg.n_features; in.F_base = h->n_out; in.use_delta = h->cfg.use_delta;
    K2Out o{};
    in.proj = use_proj ? h->d_proj_ring : nullptr;
    in.proj_tiles = (h->cfg.max_streams + 15) / 16;
    in.used = h->used;
    in.chunk = h->cfg.chunk_samples;
    o.raw = d_raw; o.conf = d_conf; o.fired = d_fired; o.count = d_count; o.trig = h->st.trig;
    return launch_gru(h, in, true, n, decode_params(h), o, s);
}

__global__ void read_window_kernel(K2In in, const int* ids, long long n, float* out) {
    // one thread per (item, row, column)
    const int Fb = in.F_base;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = n * in.T * Fb;
    if (e >= total) return;
    int f = (int)(e % Fb);
    long long q = e / Fb;
    int t = (int)(q % in.T);
    long long i = q / in.T;
    int sid = ids ? ids[i] : (int)i;
    long long ns = in.n_samples[sid];
    long long rel = ns >= in.window ? (ns - in.window) / in.hop + 1 : 0;
    const float* row = ring_row(in, sid, rel, t);
    out[e] = row ? row[f] : 0.f;
}

PB_API int pb_read_window(pb_handle* h, const int32_t* d_ids, int64_t n, float* d_out, void* stream) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (n < 0 || n > h->cfg.max_streams) return fail(PB_ERR_INVALID, "bad n");
    if (n == 0) return PB_OK;
    if (!d_out) return fail(PB_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(h->cfg.device));
    K2In in{};
    in.ring = h->st.ring; in.n_samples = h->st.n_samples; in.ids = d_ids;
    in.ring_rows = h->ring_rows; in.row_stride = h->row_stride; in.window = h->rel_window; in.hop = h->cfg.hop_samples;
    in.T = h->cfg.n_features; in.F_base = h->n_out;
    long long total = n * in.T * in.F_base;
    read_window_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, d_ids, n, d_out);
    CK(cudaGetLastError());
    return PB_OK;
}

__global__ void clear_kernel(StreamState st, const int* ids, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int sid = ids ? ids[i] : (int)i;
    st.n_samples[sid] = 0;
    st.trig[sid] = 0;
}

PB_API int pb_clear(pb_handle* h, const int32_t* d_ids, int64_t n, void* stream) {
    if (!h) return fail(PB_ERR_INVALID, "null handle");
    if (n < 0 || n > h->cfg.max_streams) return fail(PB_ERR_INVALID, "bad n");
    if (n == 0) return PB_OK;
    CK(cudaSetDevice(h->cfg.device));
    clear_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->st, d_ids, n);
    CK(cudaGetLastError());
    return PB_OK;
}

// ------------------------------------------------------------------------------------------------
PB_API int pb_host_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(PB_ERR_INVALID, "null argument");
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PB_OK;
}

PB_API int pb_host_free(void* p) {
    if (p) CK(cudaFreeHost(p));
    return PB_OK;
}

static int ensure_pipe(pb_handle* h) {
    if (h->pipe[0]) return PB_OK;
    const int64_t sb = std::min<int64_t>(HOST_SUB_BATCH, h->cfg.max_streams);
    for (int i = 0; i < HOST_PIPE; ++i) {
        CK(cudaStreamCreateWithFlags(&h->pipe[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->pipe_ev[i], cudaEventDisableTiming));
        CK(cudaMalloc((void**)&h->d_stage_pcm[i], sb * h->cfg.chunk_samples * sizeof(int16_t)));
        CK(cudaMalloc((void**)&h->d_stage_ids[i], sb * sizeof(int)));
        CK(cudaMalloc((void**)&h->d_stage_raw[i], sb * sizeof(float)));
        CK(cudaMalloc((void**)&h->d_stage_conf[i], sb * sizeof(double)));
        CK(cudaMalloc((void**)&h->d_stage_fired[i], sb * sizeof(uint8_t)));
    }
    CK(cudaHostAlloc((void**)&h->h_count_pinned, sizeof(unsigned long long), cudaHostAllocDefault));
    return PB_OK;
}

__global__ void iota_kernel(int* ids, int base, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = base + i;
}

PB_API int pb_update_host(pb_handle* h, const int16_t* h_pcm, const int32_t* h_ids, int64_t n, float* h_raw, double* h_conf,
                   uint8_t* h_fired, unsigned long long* h_count) {
    int rc = check_tick(h, h_pcm, n);
    if (rc != PB_OK) return rc;
    if (n == 0) { if (h_count) *h_count = 0; return PB_OK; }
    if (!h->have_weights) return fail(PB_ERR_STATE, "pb_load_weights has not been called");
    if (!h_conf) return fail(PB_ERR_INVALID, "null h_conf");
    CK(cudaSetDevice(h->cfg.device));
    rc = ensure_pipe(h);
    if (rc != PB_OK) return rc;
    const int64_t sb = std::min<int64_t>(HOST_SUB_BATCH, h->cfg.max_streams);
    const int chunk = h->cfg.chunk_samples;
    // ---- latency path (BASELINE configs[4]): a handful of streams in pinned host memory.  With unified addressing the
    // kernels read the PCM and write the results through the mapped host pointers directly: two launches and one
    // synchronisation instead of three staged copies.
    if (n <= HOST_ZERO_COPY_MAX) {
        auto pinned = [](const void* p) {
            if (!p) return true;
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
            return a.type == cudaMemoryTypeHost;
        };
        if (pinned(h_pcm) && pinned(h_ids) && pinned(h_raw) && pinned(h_conf) && pinned(h_fired)) {
            *h->h_count_pinned = 0;
            rc = pb_update(h, h_pcm, h_ids, n, h_raw, h_conf, h_fired, h->h_count_pinned, h->pipe[0]);
            if (rc != PB_OK) return rc;
            CK(cudaStreamSynchronize(h->pipe[0]));
            if (h_count) *h_count = *h->h_count_pinned;
            return PB_OK;
        }
    }
    // the counter is zeroed on pipe 0; the other pipes wait for that, pipe 0 waits for them at the end,
    // so the whole tick costs one host synchronisation
    CK(cudaMemsetAsync(h->d_count, 0, sizeof(unsigned long long), h->pipe[0]));
    // equal sub-batches (a short last one would drop below the size at which the tick uses the projection cache and
    // invalidate it every tick); step <= sb, a multiple of 32 except when one sub-batch takes everything
    const int64_t n_sub = (n + sb - 1) / sb;
    const int64_t step = n_sub == 1 ? n : std::min(sb, ((n + n_sub - 1) / n_sub + 31) & ~(int64_t)31);
    struct TickFlag { bool& f; ~TickFlag() { f = false; } } tick_flag{h->host_tick_proj};
    if (wants_projection(h, std::min(step, n))) {
        rc = rebuild_projections_if_dirty(h, h->pipe[0]);
        if (rc != PB_OK) return rc;
        h->host_tick_proj = true;
    }
    const int used_pipes = (int)std::min<int64_t>(HOST_PIPE, (n + step - 1) / step);
    if (used_pipes > 1) {
        CK(cudaEventRecord(h->pipe_ev[0], h->pipe[0]));
        for (int i = 1; i < used_pipes; ++i) CK(cudaStreamWaitEvent(h->pipe[i], h->pipe_ev[0], 0));
    }
    int p = 0;
    for (int64_t off = 0; off < n; off += step, p = (p + 1) % HOST_PIPE) {
        const int64_t m = std::min(step, n - off);
        cudaStream_t s = h->pipe[p];
        CK(cudaMemcpyAsync(h->d_stage_pcm[p], h_pcm + off * chunk, m * chunk * sizeof(int16_t), cudaMemcpyHostToDevice, s));
        if (h_ids) CK(cudaMemcpyAsync(h->d_stage_ids[p], h_ids + off, m * sizeof(int), cudaMemcpyHostToDevice, s));
        else { iota_kernel<<<(int)((m + 255) / 256), 256, 0, s>>>(h->d_stage_ids[p], (int)off, (int)m); CK(cudaGetLastError()); }
        rc = pb_update(h, h->d_stage_pcm[p], h->d_stage_ids[p], m, h_raw ? h->d_stage_raw[p] : nullptr, h->d_stage_conf[p],
                       h_fired ? h->d_stage_fired[p] : nullptr, h->d_count, s);
        if (rc != PB_OK) return rc;
        CK(cudaMemcpyAsync(h_conf + off, h->d_stage_conf[p], m * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (h_raw) CK(cudaMemcpyAsync(h_raw + off, h->d_stage_raw[p], m * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (h_fired) CK(cudaMemcpyAsync(h_fired + off, h->d_stage_fired[p], m * sizeof(uint8_t), cudaMemcpyDeviceToHost, s));
    }
    for (int i = 1; i < used_pipes; ++i) {
        CK(cudaEventRecord(h->pipe_ev[i], h->pipe[i]));
        CK(cudaStreamWaitEvent(h->pipe[0], h->pipe_ev[i], 0));
    }
    CK(cudaMemcpyAsync(h->h_count_pinned, h->d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->pipe[0]));
    CK(cudaStreamSynchronize(h->pipe[0]));
    if (h_count) *h_count = *h->h_count_pinned;
    return PB_OK;
}

