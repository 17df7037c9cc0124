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
            return np.max(np.abs(np.log(np.maximum(p, 1e-300)) - np.log(np.maximum(ref, 1e-300)))[msk])
        row = [f32, tc_power(x, 3), tc_power(x, 4), tc_power(x, 1)]
        print('%-24s ' % name + ' '.join('%12.3g' % err(p) for p in row) + '   | log err: ' + ' '.join('%9.2e' % lerr(p) for p in row))


if __name__ == '__main__':
    main()
