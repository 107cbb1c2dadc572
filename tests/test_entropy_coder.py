"""CPU tests (-m "not gpu"): the entropy-coder stand-in ("next" row f-4, R:codec_processing.py:433,447,488-497,
509-536).  compressai is absent, so the reference's bitstream is not reproducible here (parity unpinned); what is
checked is the contract: bit-exact round trips through the C ABI, byte counts that follow the model's
cross-entropy, escapes for out-of-table symbols, and loud failures on malformed input."""
import ctypes

import numpy as np
import pytest

import rdvc_corr_b200 as rc

ec = rc.entropy_coder


@pytest.fixture(scope="module")
def lib():
    rc._build.build()
    return rc._cabi.load()


def test_quantized_cdf_properties():
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 255, 4000):
        pmf = rng.random(n) ** 4
        pmf[rng.integers(0, n)] = 0.0                       # a zero-probability symbol still gets a slot
        cdf = ec.pmf_to_quantized_cdf(pmf)
        assert cdf[0] == 0 and cdf[-1] == 65536 and cdf.size == n + 1
        assert np.all(np.diff(cdf.astype(np.int64)) >= 1)
        if n > 1:
            f = np.diff(cdf.astype(np.int64)) / 65536.0
            big = pmf / pmf.sum() > 0.01
            assert np.allclose(f[big], (pmf / pmf.sum())[big], rtol=0.05, atol=2e-4)
    with pytest.raises(ValueError):
        ec.pmf_to_quantized_cdf(np.array([]))
    with pytest.raises(ValueError):
        ec.pmf_to_quantized_cdf(np.array([0.5, -0.1]))


def test_round_trip_is_bit_exact_and_bytes_follow_the_model(lib):
    rng = np.random.default_rng(1)
    C, H, W = 8, 17, 30
    scales = np.linspace(0.3, 12.0, C)
    prior = ec.FactorizedPrior(C, scales=scales, medians=rng.normal(0, 2, C))
    with pytest.raises(RuntimeError, match="must be updated"):
        prior.compress(np.zeros((C, H, W), np.float32))
    assert prior.update() and not prior.update() and prior.update(force=True)
    y = (rng.logistic(0, 1, (C, H, W)) * scales[:, None, None] + prior.medians[:, None, None]).astype(np.float32)
    data = prior.compress(y)
    sym = prior.quantize(y)
    assert np.array_equal(prior.decompress_symbols(data, (H, W)), sym)
    assert np.allclose(prior.decompress(data, (H, W)), sym + prior.medians[:, None, None], atol=1e-5)
    ideal = prior.cross_entropy_bits(sym) / 8
    assert ideal <= len(data) <= ideal * 1.01 + 8            # rANS overhead: the 4-byte state + < 1 %
    # the prior matches the data, so the code length is near the source entropy -- far below raw int8
    assert len(data) < 0.75 * sym.size
    # a mismatched (too narrow) prior costs bytes: counts are driven by the statistics, not constant
    narrow = ec.FactorizedPrior(C, scales=np.full(C, 0.3), medians=prior.medians)
    narrow.update()
    assert len(narrow.compress(y)) > 1.3 * len(data)
    assert np.array_equal(narrow.decompress_symbols(narrow.compress(y), (H, W)), sym)


def test_escapes_for_out_of_table_symbols(lib):
    prior = ec.FactorizedPrior(2, scales=[0.5, 0.5], tail_mass=1e-4)
    prior.update()
    half = -int(prior._offsets[0])
    sym = np.zeros((2, 4, 6), np.int32)
    sym[0, 0] = [half, half + 1, -half - 1, 10 ** 6, -10 ** 6, 2 ** 31 - 1]
    sym[1, 1] = [-2 ** 31, half - 1, -half, 0, 1, -1]
    data = prior.compress_symbols(sym)
    assert np.array_equal(prior.decompress_symbols(data, (4, 6)), sym)


def test_flow_coder_and_edge_cases(lib):
    coder = ec.FlowCoder()
    q = np.zeros((2, 135, 240), np.int8)
    q[0] += 12; q[1] += 8                                   # the GOP bench's constant translation, x4 quantiser
    data = coder.compress(q)
    assert np.array_equal(coder.decompress(data, (135, 240)), q.astype(np.int32))
    assert len(data) < q.size                               # cheaper than the raw int8 dump it replaces
    rng = np.random.default_rng(3)
    noisy = rng.integers(-127, 128, (2, 16, 16)).astype(np.int8)
    assert np.array_equal(coder.decompress(coder.compress(noisy), (16, 16)), noisy.astype(np.int32))
    assert len(coder.compress(noisy)) > len(coder.compress(np.zeros((2, 16, 16), np.int8)))
    empty = coder.compress(np.zeros((2, 0, 5), np.int8))
    assert len(empty) == 4 and coder.decompress(empty, (0, 5)).shape == (2, 0, 5)
    with pytest.raises(ValueError):
        coder.compress_symbols(np.zeros((3, 4, 4), np.int32))


def test_malformed_input_fails_loudly(lib):
    coder = ec.FlowCoder()
    q = np.arange(2 * 8 * 8, dtype=np.int32).reshape(2, 8, 8) % 9 - 4
    data = coder.compress(q)
    with pytest.raises(ValueError, match="decode failed"):
        coder.decompress(data[: len(data) // 2], (8, 8))                    # truncated stream
    cdfs = np.array([[0, 10, 5, 65536]], np.uint32)                         # not increasing
    lens, offs = np.array([4], np.int32), np.array([0], np.int32)
    sym, idx = np.zeros(4, np.int32) + 1, np.zeros(4, np.int32)
    out = np.zeros(256, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.rdvc_ec_encode_with_indexes(p(sym), p(idx), 4, p(cdfs), p(lens), p(offs), 1, 4, p(out), 256) == 0
    assert b"malformed" in lib.rdvc_corr_last_error()
    good = np.array([[0, 30000, 60000, 65536]], np.uint32)
    assert lib.rdvc_ec_encode_with_indexes(p(sym), p(idx + 3), 4, p(good), p(lens), p(offs), 1, 4, p(out), 256) == 0   # bad index
    assert lib.rdvc_ec_encode_with_indexes(p(sym), p(idx), 4, p(good), p(lens), p(offs), 1, 4, p(out), 3) == 0         # no room
    n = lib.rdvc_ec_encode_with_indexes(p(sym), p(idx), 4, p(good), p(lens), p(offs), 1, 4, p(out), 256)
    back = np.zeros(4, np.int32)
    assert n >= 4 and lib.rdvc_ec_decode_with_indexes(p(out), n, p(idx), 4, p(good), p(lens), p(offs), 1, 4, p(back)) == 0
    assert np.array_equal(back, sym)


def test_known_answer_vector(lib):
    """A hand-checkable stream: one table {0: 1/2, 1: 1/4, escape: 1/4}, symbols [0, 1, 0].  rANS state after
    encoding in reverse from x = 2^16: pins the byte layout (big-endian 32-bit state, then 16-bit words)."""
    cdfs = np.array([[0, 32768, 49152, 65536]], np.uint32)
    lens, offs = np.array([4], np.int32), np.array([0], np.int32)
    sym, idx = np.array([0, 1, 0], np.int32), np.zeros(3, np.int32)
    out = np.zeros(64, np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    n = lib.rdvc_ec_encode_with_indexes(p(sym), p(idx), 3, p(cdfs), p(lens), p(offs), 1, 4, p(out), 64)
    # by hand: x = 65536; code 0 (start 0, freq 32768): x = (65536 // 32768 << 16) + 0 + 0 = 131072
    #          code 1 (start 32768, freq 16384): x = (131072 // 16384 << 16) + 0 + 32768 = 557056
    #          code 0: x = (557056 // 32768 << 16) + 0 + 0 = 1114112 = 0x00110000; no word was emitted
    assert n == 4 and bytes(out[:4]) == bytes([0x00, 0x11, 0x00, 0x00])
