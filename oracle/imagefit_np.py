"""TEST INFRASTRUCTURE ONLY -- CPU restatement of BASELINE config C1 (``models/imagefit.py:11-71``, hash-grid arm;
``app/imagefit.py:131-141`` optimizer): 2-D hash-grid encoding (the pinned encoder restatement of hashgrid_np.py) ->
Dense(128) -> ReLU -> Dense(128) -> ReLU -> Dense(3) -> sigmoid with biases, mean squared error, optax.adam
(b1 0.9, b2 0.99, eps 1e-15).  The reference's app is bit-rotted (it cannot construct its encoder), so this model is
PARITY UNPINNED beyond the encoder; it is checked against central differences of its own forward."""
import numpy as np

from . import hashgrid_np as H

LAYERS = ("linear1", "linear2", "color_predictor")


def mlp_forward(params, enc):
    enc = np.asarray(enc, np.float64)
    a1 = enc @ params["linear1"][0] + params["linear1"][1]
    h1 = np.maximum(a1, 0)
    a2 = h1 @ params["linear2"][0] + params["linear2"][1]
    h2 = np.maximum(a2, 0)
    a3 = h2 @ params["color_predictor"][0] + params["color_predictor"][1]
    rgb = 1 / (1 + np.exp(-a3))
    return rgb, dict(enc=enc, a1=a1, h1=h1, a2=a2, h2=h2, rgb=rgb)


def encode(levels, table, uv):
    """uv in [0, 1]^2 -> float32 encoding [n, 32]; pos = 2 uv - 1 with bound 1 gives pos01 = uv (encoders.py:87)."""
    return H.encode(levels, (np.asarray(uv, np.float32) * 2 - 1).astype(np.float32), 1.0, table)


def forward(levels, table, params, uv):
    return mlp_forward(params, encode(levels, table, uv))


def loss_and_grads(levels, table, params, uv, target):
    rgb, c = forward(levels, table, params, uv)
    diff = rgb - target
    loss = float(np.mean(diff ** 2))
    d3 = 2 * diff / diff.size * rgb * (1 - rgb)
    g = {"color_predictor": (c["h2"].T @ d3, d3.sum(0))}
    d2 = (d3 @ params["color_predictor"][0].T) * (c["a2"] > 0)
    g["linear2"] = (c["h1"].T @ d2, d2.sum(0))
    d1 = (d2 @ params["linear2"][0].T) * (c["a1"] > 0)
    g["linear1"] = (c["enc"].T @ d1, d1.sum(0))
    d_enc = d1 @ params["linear1"][0].T
    g_table = H.backward(levels, (uv * 2 - 1).astype(np.float32), 1.0, d_enc.astype(np.float32), table.shape[1])
    return loss, g, np.asarray(g_table, np.float64), d_enc


def adam_update(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.99, eps=1e-15):
    """optax.adam step t (1-based): eps added outside the square root, eps_root = 0."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    p = p - lr * (m / (1 - b1 ** t)) / (np.sqrt(v / (1 - b2 ** t)) + eps)
    return p, m, v
