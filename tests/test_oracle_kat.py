"""Known-answer tests and independent cross-checks for the UNPINNED oracle parts (MFCC, GRU)."""
import numpy as np
import pytest

from oracle import mfcc as om
from oracle import gru as og
from oracle.params import OracleParams

PR = OracleParams()
ARGS = (PR.sample_rate, PR.window_samples, PR.hop_samples, PR.n_fft, PR.n_filt, PR.n_mfcc)


def test_frame_count_and_empty():
    assert om.mfcc_spec(np.zeros(1599), *ARGS).shape == (0, 13)
    assert om.mfcc_spec(np.zeros(1600), *ARGS).shape == (1, 13)
    assert om.mfcc_spec(np.zeros(2399), *ARGS).shape == (1, 13)
    assert om.mfcc_spec(np.zeros(2400), *ARGS).shape == (2, 13)
    assert om.mfcc_spec(np.zeros(24000), *ARGS).shape == (29, 13)
    with pytest.raises(ValueError):
        om.vectorize_raw(np.zeros(0), PR)


def test_zero_and_ones_frames():
    # the reference's own test signals (test/scripts/dummy_audio_folder.py:37-45)
    z = om.mfcc_spec(np.zeros(1600), *ARGS)[0]
    assert abs(z[0] - (-36.04365338911715)) < 1e-12 and np.all(np.abs(z[1:]) < 1e-12)
    o = om.mfcc_spec(np.ones(1600), *ARGS)[0]
    # ones: only the DC bin is non-zero (=512); no filter covers bin 0 with non-zero weight,
    # so every mel is log(eps) and the AC cepstral terms vanish; c0 = log(512^2/512)
    assert abs(o[0] - np.log(512.0)) < 1e-12 and np.all(np.abs(o[1:]) < 1e-9)


def test_rfft_crops_to_first_nfft_samples():
    rs = np.random.RandomState(3)
    a = rs.randn(1600)
    b = a.copy()
    b[512:] = rs.randn(1600 - 512)          # samples 512.. never enter the FFT
    assert np.array_equal(om.mfcc_spec(a, *ARGS), om.mfcc_spec(b, *ARGS))
    assert np.allclose(np.fft.rfft(a, n=512), np.fft.rfft(a[:512]))


def test_power_against_direct_dft():
    rs = np.random.RandomState(4)
    a = rs.randn(2400)
    p = om.power_frames(a, 1600, 800, 512)
    n = np.arange(512)
    k = np.arange(257)[:, None]
    dft = np.exp(-2j * np.pi * k * n / 512)
    for f, start in enumerate((0, 800)):
        X = dft @ a[start:start + 512]
        assert np.allclose(p[f], (X.real ** 2 + X.imag ** 2) / 512, rtol=1e-10, atol=1e-12)


def test_impulse_has_flat_spectrum():
    a = np.zeros(1600)
    a[0] = 0.5
    p = om.power_frames(a, 1600, 800, 512)[0]
    assert np.allclose(p, 0.25 / 512)
    m = om.mfcc_spec(a, *ARGS)[0]
    assert abs(m[0] - np.log(257 * 0.25 / 512)) < 1e-12


def test_filterbank_shape_and_ranges():
    fb = om.filterbank(16000, 20, 257)
    assert fb.shape == (20, 257)
    got = [(int(np.nonzero(r)[0][0]), int(np.nonzero(r)[0][-1])) for r in fb]
    want = [(1, 2), (2, 5), (4, 8), (7, 11), (10, 15), (13, 20), (17, 25), (22, 31), (27, 38), (33, 46),
            (40, 56), (48, 67), (58, 80), (69, 96), (82, 113), (98, 134), (115, 158), (136, 186),
            (160, 218), (188, 256)]                      # SURVEY 8c, semantic (4)
    assert got == want
    # falling edge of filter i and rising edge of filter i+1 share bins and sum to one
    g = om.mel_grid(16000, 20, 257)
    for i in range(19):
        seg = slice(int(g[i + 1]), int(g[i + 2]))
        assert np.allclose(fb[i, seg] + fb[i + 1, seg], 1.0)


def test_grid_dedup_pushes_forward():
    g = om.mel_grid(16000, 40, 257)          # n_filt=40 has colliding low bins
    assert np.all(np.diff(g) >= 1) and g[0] == 0 and g[-1] <= 257
    fb = om.filterbank(16000, 40, 257)
    assert fb.shape == (40, 257) and np.all(fb.sum(axis=1) > 0)


def test_dct_matrix_matches_scipy():
    from scipy.fftpack import dct
    rs = np.random.RandomState(5)
    x = rs.randn(7, 20)
    assert np.allclose(x @ om.dct2_ortho_matrix(20, 13).T, dct(x, norm='ortho')[:, :13], atol=1e-12)
    x = rs.randn(3, 40)
    assert np.allclose(x @ om.dct2_ortho_matrix(40, 40).T, dct(x, norm='ortho'), atol=1e-12)


def test_mfcc_pipeline_against_scipy_composition():
    from scipy.fftpack import dct
    rs = np.random.RandomState(6)
    a = rs.randn(4000) * 0.1
    p = om.power_frames(a, 1600, 800, 512)
    mels = np.log(np.clip(p @ om.filterbank(16000, 20, 257).T, np.finfo(float).eps, None))
    want = dct(mels, norm='ortho')[:, :13]
    want[:, 0] = np.log(np.clip(p.sum(1), np.finfo(float).eps, None))
    assert np.allclose(om.mfcc_spec(a, *ARGS), want, atol=1e-12)


def test_vectorize_pad_and_crop():
    rs = np.random.RandomState(8)
    short = rs.randn(4000) * 0.1
    v = om.vectorize(short, PR)
    assert v.shape == (29, 13) and np.all(v[:25] == 0) and np.any(v[25:] != 0)
    long = rs.randn(40000) * 0.1
    assert np.array_equal(om.vectorize(long, PR), om.mfcc_spec(long[-24000:], *ARGS))
    d = om.add_deltas(v)
    assert d.shape == (29, 26) and np.array_equal(d[1:, 13:], v[1:] - v[:-1]) and np.all(d[0, 13:] == 0)


# ------------------------------------------------------------------------------------ GRU
def test_gru_single_step_by_hand():
    w = og.GruWeights.random(3, 2, seed=1, scale=0.5)
    x = np.array([[[0.3, -0.2, 0.9]]], dtype=np.float32)
    K, U, b = w.kernel.astype(np.float64), w.recurrent.astype(np.float64), w.bias.astype(np.float64)
    a = x[0, 0].astype(np.float64) @ K + b
    z = np.clip(0.2 * a[0:2] + 0.5, 0, 1)
    hh = a[4:6]                       # h0 = 0: reset gate has nothing to gate, activation linear
    h1 = (1 - z) * hh
    logit = h1 @ w.dense_w.astype(np.float64) + w.dense_b
    p, lg, h = og.gru_forward(w, x, np.float64, return_hidden=True)
    assert np.allclose(h[0], h1, atol=1e-12) and abs(lg[0] - logit) < 1e-12
    assert abs(p[0] - 1 / (1 + np.exp(-logit))) < 1e-12


def test_gru_second_step_uses_reset_before_matmul():
    w = og.GruWeights.random(2, 3, seed=2, scale=0.7)
    x = np.random.RandomState(0).randn(1, 2, 2)
    K, U, b = (m.astype(np.float64) for m in (w.kernel, w.recurrent, w.bias))
    hs = lambda v: np.clip(0.2 * v + 0.5, 0, 1)
    h = np.zeros(3)
    for t in range(2):
        a = x[0, t] @ K + b
        z = hs(a[0:3] + h @ U[:, 0:3])
        r = hs(a[3:6] + h @ U[:, 3:6])
        hh = a[6:9] + (r * h) @ U[:, 6:9]
        h = z * h + (1 - z) * hh
    _, _, hg = og.gru_forward(w, x, np.float64, return_hidden=True)
    assert np.allclose(hg[0], h, atol=1e-12)


def test_gru_against_torch_cell_rearranged():
    """Independent implementation: torch matmuls, Keras reset_after=False wiring, fp64."""
    import torch
    w = og.GruWeights.random(13, 20, seed=0, scale=0.3)
    x = np.random.RandomState(9).randn(5, 29, 13) * 2
    K, U, b = (torch.tensor(m, dtype=torch.float64) for m in (w.kernel, w.recurrent, w.bias))
    xt = torch.tensor(x)
    h = torch.zeros(5, 20, dtype=torch.float64)
    for t in range(29):
        a = torch.addmm(b, xt[:, t], K)
        z = torch.clamp(0.2 * (a[:, :20] + h @ U[:, :20]) + 0.5, 0, 1)
        r = torch.clamp(0.2 * (a[:, 20:40] + h @ U[:, 20:40]) + 0.5, 0, 1)
        hh = a[:, 40:] + (r * h) @ U[:, 40:]
        h = z * h + (1 - z) * hh
    want = torch.sigmoid(h @ torch.tensor(w.dense_w, dtype=torch.float64) + w.dense_b).numpy()
    got = og.gru_forward(w, x, np.float64)[0]
    assert np.allclose(got, want, atol=1e-12)


def test_gru_fp32_close_to_fp64_and_shapes():
    w = og.GruWeights.random(13, 20, seed=0, scale=0.1)
    x = np.random.RandomState(10).randn(64, 29, 13) * 3
    p32 = og.predict(w, x)
    assert p32.shape == (64, 1) and p32.dtype == np.float32
    p64 = og.gru_forward(w, x, np.float64)[0]
    assert np.max(np.abs(p32[:, 0] - p64)) < 1e-5
    assert abs(og.run(w, x[3]) - p32[3, 0]) < 1e-7
    assert 0.02 < p64.min() and p64.max() < 0.98          # weights at this scale are unsaturated


# ---------------------------------------------------------------------------------------------------------------------
# Vectorizer.speechpy_mfccs restatement (PARITY UNPINNED: speechpy is not in the reference tree): the properties of its published
# algorithm that the restatement and the CUDA path rely on
def test_speechpy_restatement_framing_and_grid():
    from oracle import mfcc as om
    from oracle.listener import OracleListener
    from oracle.gru import GruWeights
    from oracle.params import OracleParams
    x = np.random.RandomState(1).randn(5000) * 0.1
    for n, want in ((1599, 0), (1600, 0), (2399, 0), (2400, 1), (3200, 2), (5000, 4)):       # floor((n - window) / hop): one fewer than sonopy
        assert om.speechpy_mfcc(x[:n], 16000, 1600, 800, 512, 20, 13).shape == (want, 13), n
    assert om.mfcc_spec(x[:2400], 16000, 1600, 800, 512, 20, 13).shape[0] == 2
    g = om.speechpy_grid(16000, 20, 257)
    assert g[0] == 0 and g[-1] in (128, 129) and np.all(np.diff(g) >= 1)                     # corners up to sample_rate / 2 on (n_bins + 1) hz / sr (floor of 128.99..)
    # coefficient 0 = log of the frame energy; silence hits the eps floor everywhere before the DCT
    m = om.speechpy_mfcc(x, 16000, 1600, 800, 512, 20, 13)
    p = om.power_frames(x[:4 * 800 + 1600 - 800], 1600, 800, 512)
    assert np.allclose(m[:, 0], np.log(p.sum(axis=1)))
    z = om.speechpy_mfcc(np.zeros(4000), 16000, 1600, 800, 512, 20, 13)
    assert np.allclose(z[:, 0], np.log(om.EPS64)) and np.allclose(z[:, 1:], 0.0, atol=1e-9)
    # the listener state machine is chunking-independent with this framing too
    pr = OracleParams(vectorizer=3)
    w = GruWeights.random(13, 20, seed=0, scale=0.1)
    a, b = OracleListener(w, pr), OracleListener(w, pr)
    sig = (np.random.RandomState(2).randn(1024 * 40) * 0.1).astype(np.float32)
    for k in range(40):
        wa = a.update_vectors(sig[k * 1024:(k + 1) * 1024])
    for k in range(0, len(sig), 333):
        wb = b.update_vectors(sig[k:k + 333])
    assert np.allclose(wa, wb)
