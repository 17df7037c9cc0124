/*
 * precise_b200.h -- C ABI of libprecise_b200.so: the B200 (sm_100a) implementation of the
 * Mycroft Precise streaming-inference hot path
 *
 *      int16 PCM -> MFCC -> GRU window scan + Dense + sigmoid -> threshold decode -> trigger
 *
 * Every entry point names the reference interface it stands in for (paths relative to the
 * mycroft-precise checkout, commit e1a635e).  The reference is pure Python and has no FFI for
 * this path; INTEGRATION.md shows the ctypes stubs a maintainer would add behind
 * precise.network_runner.Runner / Listener and precise_runner.Engine.
 *
 * Conventions
 *   - every function returns PB_OK (0) or a negative pb_status; nothing throws across the ABI;
 *     pb_last_error() returns a thread-local message for the last failure on this thread.
 *   - pointers named d_* are DEVICE pointers on the handle's device, h_* are HOST pointers.
 *     All buffers are caller-owned and not retained after the call returns (device work is
 *     ordered on `stream`; the caller keeps buffers alive until that stream reaches the work).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host unless stated otherwise.
 *   - a handle is not re-entrant (the reference drives one Listener from one thread,
 *     runner/precise_runner/runner.py:232-243); distinct handles / devices are independent.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     PB_ERR_CUDA.
 */
#ifndef PRECISE_B200_H
#define PRECISE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB_ABI_VERSION 1
#define PB_MAX_THRESHOLDS 8

typedef enum pb_status {
    PB_OK = 0,
    PB_ERR_INVALID = -1,      /* bad argument / shape; maps to ValueError            */
    PB_ERR_UNSUPPORTED = -2,  /* legal in the reference, not implemented here        */
    PB_ERR_CUDA = -3,         /* CUDA runtime / driver failure; maps to RuntimeError */
    PB_ERR_STATE = -4,        /* call order (e.g. predict before load_weights)       */
    PB_ERR_EOF = -5           /* empty chunk; maps to EOFError (network_runner.py:133-134) */
} pb_status;

/* Vectorizer ids, precise/params.py:121-133 */
enum { PB_VEC_MELS = 1, PB_VEC_MFCCS = 2, PB_VEC_SPEECHPY_MFCCS = 3 /* legacy vectoriser (precise/vectorization.py:40-42); restated from speechpy's published algorithm, parity unpinned */ };
/* activations of the GRU layer (precise/model.py:77-82 uses linear + Keras default hard_sigmoid) */
enum { PB_ACT_LINEAR = 0, PB_ACT_TANH = 1 };
enum { PB_RACT_HARD_SIGMOID = 0, PB_RACT_SIGMOID = 1 };

/*
 * Mirrors precise.params.ListenerParams (precise/params.py:29-118, defaults :140-144) with the
 * time fields already converted to samples the way the reference's properties do
 * (window_samples :85-87, hop_samples :90-92, n_features :80-82), plus the network size
 * (precise/model.py:40 recurrent_units), the ThresholdDecoder arguments
 * (precise/threshold_decoder.py:38) and the TriggerDetector arguments
 * (runner/precise_runner/runner.py:121).
 */
typedef struct pb_config {
    int32_t abi_version;       /* must be PB_ABI_VERSION                                     */
    int32_t device;            /* CUDA device ordinal                                        */
    int32_t max_streams;       /* capacity of the per-stream state (>= 1)                    */
    int32_t chunk_samples;     /* int16 samples per update per stream (runner.py:23: 2048 B = 1024) */
    /* ---- ListenerParams ---- */
    int32_t sample_rate;       /* 16000 */
    int32_t window_samples;    /* 1600  */
    int32_t hop_samples;       /* 800   */
    int32_t n_fft;             /* 512; power of two in [64, 1024]                            */
    int32_t n_filt;            /* 20;  <= 64                                                 */
    int32_t n_mfcc;            /* 13;  <= 64                                                 */
    int32_t n_features;        /* 29 rows per network input                                  */
    int32_t use_delta;         /* 0                                                          */
    int32_t vectorizer;        /* PB_VEC_MFCCS                                               */
    /* ---- network ---- */
    int32_t hidden;            /* GRU units, 20                                              */
    int32_t activation;        /* PB_ACT_LINEAR                                              */
    int32_t recurrent_activation; /* PB_RACT_HARD_SIGMOID                                    */
    /* ---- ThresholdDecoder ---- */
    int32_t n_thresholds;      /* number of (mu, std) pairs, 1..PB_MAX_THRESHOLDS            */
    double threshold_mu[PB_MAX_THRESHOLDS];   /* 6.0 */
    double threshold_std[PB_MAX_THRESHOLDS];  /* 4.0 */
    double threshold_center;   /* 0.2 */
    /* ---- TriggerDetector ---- */
    double sensitivity;        /* 0.5 */
    int32_t trigger_level;     /* 3   */
    int32_t decode_legacy_f64; /* 0 (default): asigmoid's `1 / x - 1` in float32, what the reference computes for the
                                * np.float32 that Runner.run returns (network_runner.py:73-74, functions.py:99-101) under
                                * NumPy >= 2 scalar promotion -- the behaviour of the reference run in this image;
                                * 1: the same expression in float64, what NumPy 1.16 (the reference's own pin, setup.py:74)
                                * evaluates, because legacy promotion makes `1 / np.float32` a float64 */
} pb_config;

typedef struct pb_handle pb_handle;

/* Fills *cfg with the reference defaults (precise/params.py:140-144, precise/model.py:40,
 * runner/precise_runner/runner.py:121,167), max_streams = 1, device = 0. */
int pb_config_default(pb_config* cfg);

/* Replaces Listener.__init__ (precise/network_runner.py:101-109): allocates per-stream state
 * (tail PCM, MFCC ring, trigger counters) for cfg->max_streams streams on cfg->device, builds
 * the mel filterbank / DCT / twiddle / CDF tables.  Derived feature width
 * feature_size = (vectorizer == MELS ? n_filt : min(n_filt, n_mfcc)) * (use_delta ? 2 : 1)
 * (precise/params.py:100-109). */
int pb_create(const pb_config* cfg, pb_handle** out);
void pb_destroy(pb_handle* h);

/* Replaces model loading (precise/model.py:48-54, network_runner.py:50-57 / :85-86).  HOST
 * pointers, Keras layout and gate order z,r,h: kernel[F][3H], recurrent[H][3H], bias[3H],
 * dense_w[H], dense_b; F = feature_size.  Synchronous. */
int pb_load_weights(pb_handle* h, const float* h_kernel, const float* h_recurrent,
                    const float* h_bias, const float* h_dense_w, float dense_b);

/* Number of MFCC frames vectorize_raw() yields for n samples:
 * n < window ? 0 : (n - window) / hop + 1  (sonopy framing, precise/vectorization.py:36-39). */
int64_t pb_mfcc_frames(const pb_handle* h, int64_t samples_per_stream);
int32_t pb_feature_size(const pb_handle* h);     /* network input width incl. deltas           */
int32_t pb_mfcc_width(const pb_handle* h);       /* columns vectorize_raw() returns            */

/* K1, stateless.  Replaces buffer_to_audio + vectorize_raw (precise/util.py:35-37,
 * precise/vectorization.py:46-50) for n_streams independent buffers:
 *   d_pcm [n_streams][samples_per_stream] int16 (scaled by 1/32768 like buffer_to_audio)
 *   d_out [n_streams][pb_mfcc_frames()][pb_mfcc_width()] float32 */
int pb_mfcc(pb_handle* h, const int16_t* d_pcm, int64_t n_streams, int64_t samples_per_stream,
            float* d_out, void* stream);
/* Same for float32 samples that are already scaled (the ndarray branch of
 * Listener.update_vectors, network_runner.py:126-127, and load_audio, util.py:65). */
int pb_mfcc_f32(pb_handle* h, const float* d_audio, int64_t n_streams, int64_t samples_per_stream,
                float* d_out, void* stream);

/* K2(+sigmoid), stateless.  Replaces Runner.predict (network_runner.py:35-37, :69-71, :88-92):
 *   d_inputs [n][n_features][feature_size] float32 -> d_out [n] float32 (the [n,1] column).
 * d_logit (optional, may be NULL) receives the pre-sigmoid Dense output. */
int pb_predict(pb_handle* h, const float* d_inputs, int64_t n, float* d_out, float* d_logit,
               void* stream);

/* K3, stateless.  Replaces ThresholdDecoder.decode (threshold_decoder.py:45-57) element-wise:
 *   d_raw [n] float32 -> d_conf [n] float64. */
int pb_decode(pb_handle* h, const float* d_raw, int64_t n, double* d_conf, void* stream);

/* Stateful tick.  Replaces, for n streams at once, Listener.update (network_runner.py:148-153)
 * followed by TriggerDetector.update (runner.py:127-142):
 *   d_pcm        [n][chunk_samples] int16, row i belongs to stream d_stream_ids[i]
 *   d_stream_ids [n] int32 in [0, max_streams), unique; NULL means 0..n-1
 *   d_raw        [n] float32  network output (optional, may be NULL)
 *   d_conf       [n] float64  decoded confidence (what Listener.update returns)
 *   d_fired      [n] uint8    TriggerDetector.update result (optional, may be NULL)
 *   d_count      [1] uint64   += number of streams that fired this tick (optional; the
 *                             caller zeroes it; this is the quantity all-reduced across GPUs) */
int pb_update(pb_handle* h, const int16_t* d_pcm, const int32_t* d_stream_ids, int64_t n,
              float* d_raw, double* d_conf, uint8_t* d_fired, unsigned long long* d_count,
              void* stream);

/* Listener.update_vectors only (network_runner.py:125-146): advance stream state, no network. */
int pb_update_vectors(pb_handle* h, const int16_t* d_pcm, const int32_t* d_stream_ids, int64_t n,
                      void* stream);

/* Same tick with HOST buffers (what Engine.get_prediction sees: runner.py:62-67).  Copies are
 * pipelined in sub-batches over internal streams; returns when h_conf/h_fired/h_count are
 * valid.  Pinned buffers (pb_host_alloc) are needed for full PCIe rate.  h_stream_ids may be
 * NULL; h_raw, h_fired, h_count may be NULL; *h_count receives this tick's count. */
int pb_update_host(pb_handle* h, const int16_t* h_pcm, const int32_t* h_stream_ids, int64_t n,
                   float* h_raw, double* h_conf, uint8_t* h_fired, unsigned long long* h_count);

/* The 29 x F window Listener.update_vectors returns (network_runner.py:146), gathered for the
 * given streams: d_out [n][n_features][pb_mfcc_width()] float32, oldest row first. */
int pb_read_window(pb_handle* h, const int32_t* d_stream_ids, int64_t n, float* d_out, void* stream);

/* Replaces Listener.clear (network_runner.py:121-123) and re-arms the stream's trigger counter.
 * d_stream_ids NULL => streams 0..n-1. */
int pb_clear(pb_handle* h, const int32_t* d_stream_ids, int64_t n, void* stream);

/* Pinned host memory for pb_update_host / benchmarks. */
int pb_host_alloc(void** out, uint64_t bytes);
int pb_host_free(void* p);

/* Per-kernel device timing (CUDA events on the launching stream), for bench.py's roofline.
 * slot 0 = MFCC kernel, 1 = GRU(+decode+trigger) kernel, 2 = decode-only kernel.
 * pb_profile_read synchronises the recorded events, returns accumulated ms and launch counts
 * since the last pb_profile_reset. */
int pb_profile_enable(pb_handle* h, int on);
int pb_profile_reset(pb_handle* h);
int pb_profile_read(pb_handle* h, double ms[4], uint64_t launches[4]);

/* Host copies of the device tables, for tests: mel filterbank [n_filt][n_fft/2+1] (float64),
 * decoder CDF (float64, length returned), and decoder range. */
int pb_get_filterbank(const pb_handle* h, double* h_out);
int64_t pb_get_cdf(const pb_handle* h, double* h_out, int64_t capacity, int32_t* min_out, int32_t* max_out);
/* Overrides the CDF table (same length as pb_get_cdf reports).  The Python host uploads the table
 * computed by numpy -- the library the reference builds it with (threshold_decoder.py:41,68-70) --
 * so decoded values are bit-identical to the reference's; the built-in table (libm exp) differs
 * from numpy's SIMD exp by at most an ulp per entry. */
int pb_set_cdf(pb_handle* h, const double* h_cd, int64_t len);

/* Test hook: route the aligned default geometry through the generic (any-alignment) MFCC kernels
 * instead of the warp-autonomous fast kernels, so both implementations are covered by parity tests. */
int pb_debug_force_generic(pb_handle* h, int on);
/* Test / A-B hook for the default network (H=20, F=13): 0 = automatic choice (warp-per-stream kernel up to 8192 streams per tick; above,
 * the fp16x3 mma.sync scan over bulk-copy-staged cached projections, which also projects the tick's new frames), 1 = CUDA-core
 * thread-per-stream kernel, 2 = tensor-core kernel also for small batches, 3 = tcgen05 scan, 7 = 3xTF32 scan with 32-stream warp
 * tiles, 8 = tcgen05 scan over cached projections, 9 = 3xTF32 scan over cached projections without staging, 10 = with staging,
 * 11 = the default scan at 5 CTAs per SM.  All variants are parity-tested on B200 (tests/test_gpu_parity.py). */
int pb_debug_gru_mode(pb_handle* h, int mode);
/* Test / A-B hook for the stateful tick's MFCC kernel (aligned default geometry).  0 = automatic: from 49 152 streams per tick on the
 * kernel with both DFT stages on the tensor cores (csrc/mfcc_tc3.cuh: int16 samples split exactly into two fp16 pieces, tcgen05
 * MMAs with TMEM accumulators), below that the FFT kernel on the CUDA cores (csrc/mfcc_fast.cuh); 2 = always the FFT kernel;
 * 3 = the FFT kernel with its original 64-bit set-up; 4 = csrc/mfcc_tc2.cuh (radix-16 butterflies on the CUDA cores, second DFT
 * stage on tcgen05); 5 = always mfcc_tc3; 100 + w = mfcc_tc3 with the phase timeline of warp w in pb_debug_counters.  All variants
 * are parity-tested on B200 (tests/test_gpu_parity.py). */
int pb_debug_k1_mode(pb_handle* h, int mode);
/* CPU model of that kernel's DFT for one frame of 512 int16 samples -> |X[k]|^2, k = 0..256 (same butterfly, operand tables
 * and layout arithmetic; no device needed).  Test hook. */
int pb_debug_tc_dft_power(const int16_t* x512, double* power257);
/* ... and of the whole kernel for one frame (accumulators + mel / log / DCT epilogue with the tables this configuration
 * would upload) -> out[min(n_filt, n_mfcc)].  No device needed.  Test hook. */
int pb_debug_tc_mfcc_frame(const pb_config* cfg, const int16_t* x512, float* out);
/* The same for k1 mode 5 (csrc/mfcc_tc3.cuh: both DFT stages on the tensor cores, int16 split exactly into two fp16 pieces);
 * power257 (optional) receives |X[k]|^2 of the raw samples as that kernel's accumulators hold it.  No device needed.  Test hook. */
int pb_debug_tc3_mfcc_frame(const pb_config* cfg, const int16_t* x512, float* out, double* power257);
/* Test/profiling hook: the first call arms, later calls read four device-side cycle counters of the wide-network
 * tensor-core kernel's MMA-issuer thread (operand wait, weight-tile wait, issue, total) for CTA 0. */
int pb_debug_counters(pb_handle* h, long long out[4]);

const char* pb_last_error(void);
int pb_abi_version(void);
const char* pb_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* PRECISE_B200_H */
