"""Model weights for the GRU + Dense network (precise/model.py:77-82).

Native format: ``<name>.npz`` with keys kernel[F,3H], recurrent[H,3H], bias[3H], dense_w[H],
dense_b, plus optional activation / recurrent_activation strings, next to the reference's own
``<name>.params`` JSON (precise/params.py:150-165).  Keras gate order z, r, h.
"""
import numpy as np


class GruModel:
    def __init__(self, kernel, recurrent, bias, dense_w, dense_b, activation='linear',
                 recurrent_activation='hard_sigmoid'):
        self.kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        self.recurrent = np.ascontiguousarray(recurrent, dtype=np.float32)
        self.bias = np.ascontiguousarray(bias, dtype=np.float32).reshape(-1)
        self.dense_w = np.ascontiguousarray(dense_w, dtype=np.float32).reshape(-1)
        self.dense_b = float(np.asarray(dense_b, dtype=np.float32).reshape(-1)[0])
        self.activation = str(activation)
        self.recurrent_activation = str(recurrent_activation)
        self.hidden = self.recurrent.shape[0]
        self.feature_size = self.kernel.shape[0]
        H = self.hidden
        if (self.kernel.shape[1] != 3 * H or self.recurrent.shape != (H, 3 * H) or
                self.bias.shape != (3 * H,) or self.dense_w.shape != (H,)):
            raise ValueError('inconsistent GRU weight shapes')

    @staticmethod
    def random(feature_size=13, hidden=20, seed=0, scale=0.3):
        """Seeded synthetic weights (the reference ships no trained model)."""
        rs = np.random.RandomState(seed)
        return GruModel(rs.randn(feature_size, 3 * hidden) * scale, rs.randn(hidden, 3 * hidden) * scale,
                        rs.randn(3 * hidden) * scale, rs.randn(hidden) * scale, rs.randn(1) * scale)


def save_weights(path: str, m: GruModel):
    np.savez(path, kernel=m.kernel, recurrent=m.recurrent, bias=m.bias, dense_w=m.dense_w,
             dense_b=np.float32(m.dense_b), activation=m.activation,
             recurrent_activation=m.recurrent_activation)


def load_weights(path: str) -> GruModel:
    if path.endswith('.npz'):
        z = np.load(path, allow_pickle=False)
        return GruModel(z['kernel'], z['recurrent'], z['bias'], z['dense_w'], z['dense_b'],
                        str(z['activation']) if 'activation' in z else 'linear',
                        str(z['recurrent_activation']) if 'recurrent_activation' in z else 'hard_sigmoid')
    if path.endswith('.pb'):
        from .pb_import import load_pb
        return load_pb(path)
    raise ValueError('File extension of ' + path + " must be: ['.npz', '.pb'] "
                     '(Keras .net / HDF5 import is not implemented: convert with precise-convert)')
