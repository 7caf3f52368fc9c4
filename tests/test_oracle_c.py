"""The C restatement (bench.py's CPU baseline) agrees with the golden-pinned numpy oracle."""
import numpy as np

from oracle import bitdelta_oracle as O
from oracle import c_oracle as C


def test_c_codec_matches_reference_vectors(golden):
    g = golden("codec.npz")
    bits = g["bits"][0, 0]
    assert np.array_equal(C.pack_i32(bits), g["packed32"][0, 0])
    assert np.array_equal(C.unpack_i32(g["packed32"][1, 2]), g["bits"][1, 2])
    assert np.array_equal(C.unpack_i32(g["words"][0]), g["words_unpacked"][0])


def test_c_forward_matches_numpy_oracle(golden):
    g = golden("demo_modules.npz")
    y = C.fwd_batched_bf16(g["x"], g["weight"], g["masks"], O.bf16_bits_to_f32(g["coeffs"]), threads=2)
    exact = O.diffcompress_forward_exact(O.bf16_bits_to_f32(g["x"]), O.bf16_bits_to_f32(g["weight"]), g["masks"], O.bf16_bits_to_f32(g["coeffs"]))
    assert O.rel_mean_abs_err(y, exact) < 1e-5
    d = C.fwd_batched_bf16(g["x"], None, g["masks"], None)
    assert O.rel_mean_abs_err(d, O.binary_bmm_exact(O.bf16_bits_to_f32(g["x"]), g["masks"])) < 1e-5


def test_c_forward_ragged_shapes():
    rng = np.random.default_rng(0)
    T, m, K, N = 3, 2, 96, 70
    x = O.f32_to_bf16_bits(rng.standard_normal((T, m, K)).astype(np.float32))
    w = O.f32_to_bf16_bits((rng.standard_normal((N, K)) * 0.05).astype(np.float32))
    masks = rng.integers(-(2**31), 2**31 - 1, (T, K // 32, N)).astype(np.int32)
    coeff = rng.random(T).astype(np.float32) * 0.01
    y = C.fwd_batched_bf16(x, w, masks, coeff)
    exact = O.diffcompress_forward_exact(O.bf16_bits_to_f32(x), O.bf16_bits_to_f32(w), masks, coeff)
    assert O.rel_mean_abs_err(y, exact) < 1e-5
