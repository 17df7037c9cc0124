"""Oracle restatement of the network forward pass.  TEST INFRASTRUCTURE ONLY.  ** PARITY UNPINNED **

The network is defined by the reference at ``precise/model.py:77-82``::

    Sequential([GRU(recurrent_units, activation='linear', input_shape=(n_features, feature_size),
                    dropout=..., name='net'),
                Dense(1, activation='sigmoid')])

and evaluated by ``KerasRunner.predict`` / ``TensorFlowRunner.predict``
(``precise/network_runner.py:88-92`` / ``:69-71``) -- both stateless: every call starts from
h0 = 0 and scans all ``n_features`` rows (``Listener.update``, ``network_runner.py:148-153``).

The arithmetic lives in Keras (<=2.1.5 per ``setup.py:78``; 2.2.4 per ``requirements.txt:11``)
on TensorFlow 1.13 CPU, neither of which is present.  Published Keras GRU semantics restated:

  * weights ``kernel[F, 3H]``, ``recurrent_kernel[H, 3H]``, ``bias[3H]``; gate order z, r, h
  * recurrent_activation = hard_sigmoid(x) = clip(0.2 x + 0.5, 0, 1)   (Keras default)
  * reset_after = False (Keras default): the reset gate multiplies h BEFORE the matmul
        z  = hs(x Wz + h Uz + bz)
        r  = hs(x Wr + h Ur + br)
        hh = act(x Wh + (r * h) Uh + bh)      act = linear here (model.py:78)
        h' = z * h + (1 - z) * hh
  * return_sequences = False, dropout inactive at inference
  * Dense: y = sigmoid(h_T Wd + bd);   Keras casts inputs to float32.

``activation`` / ``recurrent_activation`` are parameters because a saved model's config may
override them (the weight importer reads them).
"""
import numpy as np


def hard_sigmoid(x):
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def sigmoid(x):
    one = x.dtype.type(1)
    return one / (one + np.exp(-x))


_ACT = {
    'linear': lambda x: x,
    'tanh': np.tanh,
    'hard_sigmoid': hard_sigmoid,
    'sigmoid': sigmoid,
}


class GruWeights:
    """Plain container: kernel[F,3H], recurrent[H,3H], bias[3H], dense_w[H], dense_b."""

    def __init__(self, kernel, recurrent, bias, dense_w, dense_b,
                 activation='linear', recurrent_activation='hard_sigmoid'):
        self.kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        self.recurrent = np.ascontiguousarray(recurrent, dtype=np.float32)
        self.bias = np.ascontiguousarray(bias, dtype=np.float32).reshape(-1)
        self.dense_w = np.ascontiguousarray(dense_w, dtype=np.float32).reshape(-1)
        self.dense_b = float(np.float32(np.asarray(dense_b).reshape(-1)[0]))
        self.activation = activation
        self.recurrent_activation = recurrent_activation
        self.F = self.kernel.shape[0]
        self.H = self.recurrent.shape[0]
        assert self.kernel.shape == (self.F, 3 * self.H)
        assert self.recurrent.shape == (self.H, 3 * self.H)
        assert self.bias.shape == (3 * self.H,)
        assert self.dense_w.shape == (self.H,)

    @staticmethod
    def random(F=13, H=20, seed=0, scale=0.3):
        """Seeded synthetic weights (no trained model ships with the reference)."""
        rs = np.random.RandomState(seed)
        return GruWeights(rs.randn(F, 3 * H) * scale, rs.randn(H, 3 * H) * scale,
                          rs.randn(3 * H) * scale, rs.randn(H) * scale, rs.randn(1) * scale)


def gru_forward(w: GruWeights, x: np.ndarray, dtype=np.float32, return_hidden=False):
    """x[N, T, F] -> (prob[N], logit[N]) in ``dtype``.  Batched over N; sequential over T."""
    x = np.asarray(x).astype(dtype)
    if x.ndim == 2:
        x = x[None]
    N, T, F = x.shape
    assert F == w.F, (F, w.F)
    H = w.H
    K = w.kernel.astype(dtype)
    U = w.recurrent.astype(dtype)
    b = w.bias.astype(dtype)
    act = _ACT[w.activation]
    ract = _ACT[w.recurrent_activation]
    h = np.zeros((N, H), dtype=dtype)
    for t in range(T):
        a = x[:, t, :] @ K + b                       # input projection, all three gates
        zr = ract(a[:, :2 * H] + h @ U[:, :2 * H])
        z, r = zr[:, :H], zr[:, H:]
        hh = act(a[:, 2 * H:] + (r * h) @ U[:, 2 * H:])
        h = z * h + (dtype(1) - z) * hh
    logit = h @ w.dense_w.astype(dtype) + dtype(w.dense_b)
    prob = sigmoid(logit)
    if return_hidden:
        return prob, logit, h
    return prob, logit


def predict(w: GruWeights, inputs: np.ndarray) -> np.ndarray:
    """``Runner.predict`` contract (network_runner.py:35-37): [N,T,F] -> float32 [N,1]."""
    return gru_forward(w, inputs, np.float32)[0].astype(np.float32)[:, None]


def run(w: GruWeights, inp: np.ndarray) -> float:
    """``Runner.run`` contract (network_runner.py:73-74 / :94-95): [T,F] -> scalar."""
    return predict(w, inp[np.newaxis])[0][0]
