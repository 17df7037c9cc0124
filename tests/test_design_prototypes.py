"""CPU: the numerical design study for the round-2 tensor-core DFT (scripts/proto_tc_dft.py) stays valid: the block twiddle
matrices, the accumulator-column -> bin map and the fp16 hi/lo three-pass scheme reproduce a float64 power spectrum."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _proto():
    spec = importlib.util.spec_from_file_location('proto_tc_dft', os.path.join(ROOT, 'scripts', 'proto_tc_dft.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_radix16_block_formulation_matches_float64_fft():
    p = _proto()
    assert p.check_radix16() < 5e-6
    cols = p.column_bins()
    used = sorted(c for pair in cols for c in pair if c >= 0)
    assert len(cols) == 257 and used == list(range(512))          # every accumulator column is exactly one bin component


def test_radix4_three_pass_matches_float64_fft():
    p = _proto()
    rs = np.random.RandomState(3)
    x = np.clip(rs.randn(8, 512) * 3000, -32768, 32767).astype(np.int16)
    ref = np.abs(np.fft.rfft(x.astype(np.float64), axis=1)) ** 2
    for passes, tol in ((3, 5e-6), (1, 1e-2)):
        got = p.tc_power(x, passes)
        err = np.max(np.abs(got - ref) / ref.max(axis=1, keepdims=True))
        assert err < tol
    assert np.max(np.abs(p.tc_power(x, 1) - ref) / ref.max(axis=1, keepdims=True)) > 1e-5      # one fp16 pass is not enough
