"""Numerics of the two fp16 split schemes the tensor-core decoder path uses, emulated in numpy (IEEE half incl. subnormals, exact
products, wide accumulation -- what tcgen05 kind::f16 with fp32 accumulators computes up to accumulation order):

* cluster decoder GEMMs (csrc/decoder_cluster.cu, api.cu split_f16): W = W1 + 2^-11 W2 with W2 = fp16((W - W1) * 2048), activations
  likewise; W.x ~ W1.x1 + 2^-11 (W1.x2 + W2.x1)
* projection / cross-attention K,V GEMMs (api.cu concat3_f16 + split3_act16_kernel): unscaled low parts, ONE GEMM of triple depth
  [A1 | A2 | A1] x [W1 | W1 | W2]

Both must sit at the level of fp32 arithmetic (the reference's decoder is fp32), far below what single fp16 operands give."""
import numpy as np


def _h(x):
    return x.astype(np.float16).astype(np.float64)


def _shapes(seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((64, 768)) * rng.uniform(0.05, 2.0, (1, 768))  # activations of mixed scale
    w = rng.uniform(-1, 1, (256, 768)) / np.sqrt(768)                      # nn.Linear default init range
    return a.astype(np.float32).astype(np.float64), w.astype(np.float32).astype(np.float64)


def _rel(got, ref):
    return float(np.abs(got - ref).max() / np.abs(ref).max())


def test_scaled_split_of_the_cluster_decoder_is_fp32_accurate():
    a, w = _shapes(0)
    ref = a @ w.T
    a1, w1 = _h(a), _h(w)
    a2, w2 = _h((a - a1) * 2048.0), _h((w - w1) * 2048.0)
    got = a1 @ w1.T + (a1 @ w2.T + a2 @ w1.T) / 2048.0
    fp32 = (a.astype(np.float32) @ w.T.astype(np.float32)).astype(np.float64)
    single = a1 @ w1.T
    assert _rel(got, ref) < 2e-6
    assert _rel(got, ref) < 4 * max(_rel(fp32, ref), 2e-7)
    assert _rel(single, ref) > 50 * _rel(got, ref)  # what one fp16 product per element would cost


def test_unscaled_triple_depth_gemm_of_the_projection_is_fp32_accurate():
    a, w = _shapes(1)
    ref = a @ w.T
    a1, w1 = _h(a), _h(w)
    a2, w2 = _h(a - a1), _h(w - w1)  # low parts unscaled: fp16 subnormals keep an absolute step of 2^-24
    a3 = np.concatenate([a1, a2, a1], axis=1)
    w3 = np.concatenate([w1, w1, w2], axis=1)
    got = a3 @ w3.T
    assert np.abs(w2).max() < 6.2e-5  # the weights' low parts really are subnormal: the case the comment in api.cu argues about
    assert _rel(got, ref) < 3e-6
    # second GEMM of dec_project: K = 256 on the ReLU output of the first
    m = np.maximum(got, 0.0).astype(np.float32).astype(np.float64)
    rng = np.random.default_rng(2)
    wk = (rng.uniform(-1, 1, (3072, 256)) / 16.0).astype(np.float32).astype(np.float64)
    m1, k1 = _h(m), _h(wk)
    got2 = np.concatenate([m1, _h(m - m1), m1], axis=1) @ np.concatenate([k1, k1, _h(wk - k1)], axis=1).T
    assert _rel(got2, m @ wk.T) < 3e-6
