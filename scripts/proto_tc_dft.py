"""Numerical prototype (CPU, numpy) of a tensor-core formulation of K1's 512-point real DFT -- design study for round 2.

Idea: K1 is instruction-issue bound on the CUDA cores.  The DFT can instead run as fp16 GEMMs with fp32
accumulation on tcgen05 if both operands are split into fp16 pieces:
  * one exact radix-4 decimation step on the integer samples (adds only): s, t (real, 128 each) and c, d (128 each) give
    X[4m] = DFT128(s), X[4m+2] = DFT128(t * w256^n), X[4m+1] = DFT128((c - i d) * w512^n); X[4m+3] follows by symmetry;
  * inputs (<= 18-bit integers, scaled by 2^-15) split EXACTLY into two fp16 pieces; the twiddle matrices into hi + lo fp16;
  * three products hi*Whi + lo*Whi + hi*Wlo (the lo*Wlo term is below fp32 resolution).
This script measures the error of that scheme against a float64 FFT and against a float32 FFT (what K1 computes today).
"""
import numpy as np


def split16(v):
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float64)).astype(np.float16)
    return hi, lo


def gemm_f16(a, b):
    """fp16 x fp16 -> fp32 accumulate (products of two fp16 are exact in fp32)."""
    return a.astype(np.float32) @ b.astype(np.float32)


def matrices():
    n = np.arange(128)
    m = np.arange(128)
    w128 = np.exp(-2j * np.pi * np.outer(n, m) / 128)                 # [n, m]
    Ws = w128                                                         # s -> X[4m]
    Wt = np.exp(-2j * np.pi * n / 256)[:, None] * w128                # t -> X[4m+2]
    Wc = np.exp(-2j * np.pi * n / 512)[:, None] * w128                # (c - i d) -> X[4m+1]
    # real GEMM operands: inputs [s | t | c | d] (512), outputs re/im
    def ri(W):
        return np.concatenate([W.real, W.imag], axis=1)               # [128, 256]
    Bs, Bt = ri(Ws), ri(Wt)
    Bc = np.concatenate([ri(Wc), ri(-1j * Wc)], axis=0)               # rows: c then d  -> [256, 256]
    return Bs, Bt, Bc


def tc_power(x_int, passes=3):
    """x_int: [frames, 512] int16 -> power spectrum [frames, 257] (|X|^2, no scaling)."""
    x = x_int.astype(np.int64)
    x0, x1, x2, x3 = x[:, 0:128], x[:, 128:256], x[:, 256:384], x[:, 384:512]
    a, b = x0 + x2, x1 + x3
    s, t, c, d = a + b, a - b, x0 - x2, x1 - x3
    sc = 2.0 ** -15
    Bs, Bt, Bc = matrices()
    out = {}
    for name, A, B in (('s', s, Bs), ('t', t, Bt), ('c', np.concatenate([c, d], axis=1), Bc)):
        ah, al = split16(A * sc)
        assert np.array_equal(ah.astype(np.float64) + al.astype(np.float64), A * sc)       # exact split
        bh, bl = split16(B)
        acc = gemm_f16(ah, bh) + gemm_f16(al, bh) + gemm_f16(ah, bl)
        if passes == 4:
            acc = acc + gemm_f16(al, bl)
        if passes == 1:
            acc = gemm_f16(ah, bh)
        out[name] = (acc[:, :128].astype(np.float64) + 1j * acc[:, 128:].astype(np.float64)) / sc
    X = np.zeros((x.shape[0], 512), complex)
    X[:, 0::4] = out['s']; X[:, 2::4] = out['t']; X[:, 1::4] = out['c']
    k3 = np.arange(3, 512, 4)
    X[:, k3] = np.conj(X[:, (512 - k3) % 512])
    return (X.real ** 2 + X.imag ** 2)[:, :257]


def main():
    rs = np.random.RandomState(0)
    cases = {
        'noise sigma 3000': np.clip(rs.randn(64, 512) * 3000, -32768, 32767).astype(np.int16),
        'tone 1 kHz full scale': (32000 * np.sin(2 * np.pi * 1000 / 16000 * np.arange(512) + rs.rand(64, 1) * 6.28)).astype(np.int16),
        'quiet noise sigma 3': np.round(rs.randn(64, 512) * 3).astype(np.int16),
        'tone + noise floor': (20000 * np.sin(2 * np.pi * 440 / 16000 * np.arange(512)) + rs.randn(64, 512) * 2).astype(np.int16),
    }
    print('%-24s %12s %12s %12s %12s' % ('case', 'fp32 FFT', 'tc 3-pass', 'tc 4-pass', 'tc 1-pass'))
    for name, x in cases.items():
        ref = np.abs(np.fft.rfft(x.astype(np.float64), axis=1)) ** 2
        f32 = np.abs(np.fft.rfft(x.astype(np.float32), axis=1).astype(np.complex64)) ** 2      # numpy computes in f64 internally; see below
        # honest fp32 emulation: complex64 DFT by matrix product in float32
        W = np.exp(-2j * np.pi * np.outer(np.arange(512), np.arange(257)) / 512).astype(np.complex64)
        f32 = np.abs(x.astype(np.complex64) @ W) ** 2
        peak = ref.max(axis=1, keepdims=True)
        def err(p):
            return np.max(np.abs(p - ref) / peak)                     # error relative to the frame's largest bin
        def lerr(p):                                                  # worst absolute error of log(power) over bins above 1e-9 of the peak
            msk = ref > 1e-9 * peak
            return np.max(np.abs(np.log(np.maximum(p, 1e-300)[msk]) - np.log(ref[msk])))
        row = [f32, tc_power(x, 3), tc_power(x, 4), tc_power(x, 1)]
        print('%-24s ' % name + ' '.join('%12.3g' % err(p) for p in row) + '   | log err: ' + ' '.join('%9.2e' % lerr(p) for p in row))




# ------------------------------------------------------------------------------------------------------------------
# Variant II (the one sized to fit shared memory): radix-16 decimation on the CUDA cores, 32-point sub-DFTs on tcgen05.
#   n = n2 + 32 q (n2 < 32, q < 16),  k = 16 m + r:   X[16 m + r] = sum_n2 Y_r[n2] w512^(n2 r) w32^(n2 m),
#   Y_r[n2] = sum_q x[n2 + 32 q] w16^(q r)  (16-point DFT of REAL data: Y_0, Y_8 real, Y_(16-r) = conj(Y_r)).
# Eight 64 x 64 real GEMM blocks per frame (32 768 twiddle entries = 128 KB as fp16 hi + lo, resident in shared memory):
#   block 0: inputs [Y_0 | Y_8]            -> columns [Re X[16m] | Im X[16m] (slot m=0 carries X[256]) | Re X[16m+8] | Im X[16m+8]]
#   block r: inputs [Re Y_r | Im Y_r]      -> columns [Re X[16m+r] | Im X[16m+r] | Re X[16m+16-r] | Im X[16m+16-r]],  m = 0..15
# A thread of the epilogue owns one frame (= one TMEM lane) and reads its 512 columns; column -> bin is the table below.
def block_matrices():
    n2 = np.arange(32)[:, None]
    m = np.arange(16)[None, :]
    w32 = np.exp(-2j * np.pi * n2 * m / 32)
    B = np.zeros((8, 64, 64))
    # block 0
    T0 = w32.copy()                                                    # X[16m] from Y_0 (real input): rows n2
    T8 = np.exp(-2j * np.pi * n2 * 8 / 512) * w32                      # X[16m+8] from Y_8
    B[0, :32, 0:16] = T0.real; B[0, :32, 16:32] = T0.imag
    B[0, :32, 16] = ((-1.0) ** np.arange(32))                          # Im X[0] == 0: the slot carries X[256] = sum (-1)^n2 Y_0
    B[0, 32:, 32:48] = T8.real; B[0, 32:, 48:64] = T8.imag
    for r in range(1, 8):
        Ta = np.exp(-2j * np.pi * n2 * r / 512) * w32                  # multiplies Y_r
        Tb = np.exp(-2j * np.pi * n2 * (16 - r) / 512) * w32           # multiplies conj(Y_r)
        # (a + i b) T = (a Tr - b Ti) + i (a Ti + b Tr);   (a - i b) T = (a Tr + b Ti) + i (a Ti - b Tr)
        B[r, :32, 0:16] = Ta.real;  B[r, 32:, 0:16] = -Ta.imag
        B[r, :32, 16:32] = Ta.imag; B[r, 32:, 16:32] = Ta.real
        B[r, :32, 32:48] = Tb.real; B[r, 32:, 32:48] = Tb.imag
        B[r, :32, 48:64] = Tb.imag; B[r, 32:, 48:64] = -Tb.real
    return B


def column_bins():
    """(re column, im column or -1) of every bin k = 0..256 in the 512-column accumulator row (block b at columns 64 b)."""
    cols = {}
    for m in range(16):
        cols[16 * m] = (m, 16 + m if m else -1)
        cols[16 * m + 8] = (32 + m, 48 + m)
        for r in range(1, 8):
            cols[16 * m + r] = (64 * r + m, 64 * r + 16 + m)
            cols[16 * m + 16 - r] = (64 * r + 32 + m, 64 * r + 48 + m)
    cols[256] = (16, -1)
    return [cols[k] for k in range(257)]


def tc_power_radix16(x_int, scale=2.0 ** -4):
    """Emulates variant II: fp32 radix-16 butterflies (numpy fft in float32 precision), fp16 hi/lo split, 3-pass GEMMs."""
    F = x_int.shape[0]
    x = x_int.astype(np.float32).reshape(F, 16, 32)                    # [frame, q, n2]
    Y = np.fft.fft(x.astype(np.float64), axis=1).astype(np.complex64)  # over q; the kernel does this in fp32 registers
    A = np.zeros((F, 8, 64), np.float32)
    A[:, 0, :32] = Y[:, 0, :].real; A[:, 0, 32:] = Y[:, 8, :].real
    for r in range(1, 8):
        A[:, r, :32] = Y[:, r, :].real; A[:, r, 32:] = Y[:, r, :].imag
    A *= np.float32(scale)                                             # |values| <= 16 * 32768 / 16 < fp16 max
    B = block_matrices()
    bh, bl = split16(B)
    ah = A.astype(np.float16)
    al = (A - ah.astype(np.float32)).astype(np.float16)
    D = np.zeros((F, 512), np.float32)
    for b in range(8):
        D[:, 64 * b:64 * b + 64] = gemm_f16(ah[:, b], bh[b]) + gemm_f16(al[:, b], bh[b]) + gemm_f16(ah[:, b], bl[b])
    P = np.zeros((F, 257), np.float64)
    for k, (cr, ci) in enumerate(column_bins()):
        re = D[:, cr].astype(np.float64)
        im = D[:, ci].astype(np.float64) if ci >= 0 else 0.0
        P[:, k] = (re * re + im * im) / scale ** 2
    return P


def check_radix16():
    rs = np.random.RandomState(1)
    x = np.clip(rs.randn(32, 512) * 3000, -32768, 32767).astype(np.int16)
    x[0] = 32767; x[1] = -32768; x[2] = 0; x[3, ::2] = 32767; x[3, 1::2] = -32768      # DC extremes, silence, Nyquist
    ref = np.abs(np.fft.rfft(x.astype(np.float64), axis=1)) ** 2
    got = tc_power_radix16(x)
    peak = np.maximum(ref.max(axis=1, keepdims=True), 1.0)
    return float(np.max(np.abs(got - ref) / peak))


if __name__ == '__main__':
    main()
    print('variant II (radix-16 + eight 64x64 fp16x3 GEMM blocks): max |P - P64| / peak = %.3g' % check_radix16())
