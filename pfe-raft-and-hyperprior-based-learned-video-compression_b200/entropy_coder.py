"""Entropy-coder stand-in for the hyperprior codecs' bottlenecks ("next" row f-4, last piece).

The reference's ``VideoCodec`` holds two ``compressai.entropy_models.EntropyBottleneck`` modules
(R:codec_processing.py:433 motion, :447 residual) and turns latents into byte strings with
``bottleneck.compress(latent)`` (:488-497) / back with ``decompress(strings, size)`` (:509-536).  compressai is
not in this image and is not vendored by the reference, so the exact bitstream cannot be reproduced or checked
here: **bitstream parity is unpinned**.  This module restates the same contract from scratch:

* a FACTORISED prior: every channel has its own fixed symbol distribution, independent of position -- here a
  discretised logistic per channel (median, scale) where the reference has a small learned density network;
  ``update()`` turns the densities into quantised 16-bit CDF tables with an escape slot for the tails, which is what
  ``EntropyBottleneck.update(force=True)`` does (R:codec_processing.py:464-478);
* ``compress`` quantises to integers around the channel median and codes them channel-major with the range coder
  behind the C ABI (``rdvc_ec_encode_with_indexes``, csrc/entropy_coder.h: rANS, 32-bit state, 16-bit words, 4-bit
  bypass digits for out-of-table symbols); ``decompress`` is its exact inverse;
* byte counts follow the statistics: len(compress(y)) is the cross-entropy of the symbols under the tables + 4 bytes.

The coder itself is host C (the reference's runs on the CPU too); PyTorch / numpy only carry the arrays.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _cabi

PRECISION = 16


def pmf_to_quantized_cdf(pmf: np.ndarray, precision: int = PRECISION) -> np.ndarray:
    """Probabilities (any positive scale) -> integer CDF of len(pmf) + 1 entries, first 0, last 2^precision, strictly
    increasing (every symbol keeps a frequency >= 1 so it stays codable)."""
    pmf = np.asarray(pmf, np.float64)
    if pmf.ndim != 1 or pmf.size < 1 or pmf.size >= (1 << precision) or not np.all(np.isfinite(pmf)) or np.any(pmf < 0):
        raise ValueError("pmf must be a non-empty 1-D array of finite non-negative numbers shorter than 2^precision")
    total = 1 << precision
    p = pmf / max(pmf.sum(), 1e-300)
    freq = np.maximum(1, np.floor(p * total)).astype(np.int64)
    diff = int(total - freq.sum())
    if diff > 0:                                   # rounding surplus: largest remainders first (diff <= len(pmf))
        rem = p * total - np.floor(p * total)
        freq[np.argsort(-rem, kind="stable")[:diff]] += 1
        freq[int(np.argmax(freq))] += total - int(freq.sum())      # (only for a degenerate all-zero pmf)
    elif diff < 0:                                 # the >= 1 floor overshot: take it back from the largest entries
        for k in np.argsort(-freq, kind="stable"):
            take = min(-diff, int(freq[k]) - 1)
            freq[k] -= take
            diff += take
            if diff == 0:
                break
        if diff != 0:
            raise ValueError("too many symbols for this precision")
    cdf = np.zeros(pmf.size + 1, np.uint32)
    cdf[1:] = np.cumsum(freq)
    assert cdf[-1] == total and np.all(np.diff(cdf.astype(np.int64)) >= 1)
    return cdf


class FactorizedPrior:
    """Per-channel discretised-logistic prior + range coder: the role of ``EntropyBottleneck`` in the reference.

    ``medians`` / ``scales``: one value per channel.  ``tail_mass``: probability left to the escape symbol; the table of
    a channel covers the integers whose total mass is 1 - tail_mass."""

    def __init__(self, channels: int, scales: Optional[Sequence[float]] = None,
                 medians: Optional[Sequence[float]] = None, tail_mass: float = 1e-9):
        if channels <= 0:
            raise ValueError("channels must be positive")
        self.channels = channels
        self.scales = np.full(channels, 1.0) if scales is None else np.asarray(scales, np.float64).reshape(channels)
        self.medians = np.zeros(channels) if medians is None else np.asarray(medians, np.float64).reshape(channels)
        if np.any(self.scales <= 0):
            raise ValueError("scales must be positive")
        self.tail_mass = float(tail_mass)
        self._cdfs = None
        self._lengths = None
        self._offsets = None

    # -- tables ----------------------------------------------------------------------------------------------
    def update(self, force: bool = False) -> bool:
        """Build the quantised CDF tables (like EntropyBottleneck.update, R:codec_processing.py:464-478)."""
        if self._cdfs is not None and not force:
            return False
        half = []
        for s in self.scales:                       # symmetric support: logistic quantile of tail_mass / 2
            q = s * np.log(2.0 / self.tail_mass - 1.0)
            half.append(int(min(max(np.ceil(q), 1), 30000)))
        max_len = 2 * max(half) + 1 + 2             # symbols + escape + closing entry
        cdfs = np.zeros((self.channels, max_len), np.uint32)
        lengths = np.zeros(self.channels, np.int32)
        offsets = np.zeros(self.channels, np.int32)
        sig = lambda z: 0.5 * (1.0 + np.tanh(0.5 * z))
        for c in range(self.channels):
            k = np.arange(-half[c], half[c] + 1, dtype=np.float64)
            pmf = sig((k + 0.5) / self.scales[c]) - sig((k - 0.5) / self.scales[c])
            tail = max(1.0 - pmf.sum(), self.tail_mass)
            cdf = pmf_to_quantized_cdf(np.concatenate([pmf, [tail]]))
            cdfs[c, :cdf.size] = cdf
            lengths[c] = cdf.size
            offsets[c] = -half[c]
        self._cdfs, self._lengths, self._offsets = np.ascontiguousarray(cdfs), lengths, offsets
        return True

    def _tables(self):
        if self._cdfs is None:
            raise RuntimeError("Entropy bottleneck must be updated: call update() before compress / decompress")
        return self._cdfs, self._lengths, self._offsets

    # -- quantisation ----------------------------------------------------------------------------------------
    def quantize(self, y) -> np.ndarray:
        """(C, H, W) latents -> int32 symbols round(y - median) (the "symbols" mode of the reference's bottleneck)."""
        y = np.asarray(y)
        if y.ndim != 3 or y.shape[0] != self.channels:
            raise ValueError(f"expected a (C={self.channels}, H, W) array, got {y.shape}")
        return np.round(y.astype(np.float64) - self.medians[:, None, None]).astype(np.int32)

    def dequantize(self, symbols: np.ndarray) -> np.ndarray:
        return (symbols.astype(np.float64) + self.medians[:, None, None]).astype(np.float32)

    # -- coding ----------------------------------------------------------------------------------------------
    def compress_symbols(self, symbols: np.ndarray) -> bytes:
        cdfs, lengths, offsets = self._tables()
        sym = np.ascontiguousarray(symbols, np.int32)
        if sym.ndim != 3 or sym.shape[0] != self.channels:
            raise ValueError(f"expected (C={self.channels}, H, W) symbols, got {sym.shape}")
        idx = np.ascontiguousarray(np.broadcast_to(np.arange(self.channels, dtype=np.int32)[:, None, None], sym.shape))
        lib = _cabi.load()
        cap = lib.rdvc_ec_max_encoded_bytes(sym.size)
        out = np.empty(cap, np.uint8)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        n = lib.rdvc_ec_encode_with_indexes(p(sym), p(idx), sym.size, p(cdfs), p(lengths), p(offsets), self.channels,
                                            cdfs.shape[1], p(out), cap)
        if n == 0 and sym.size:
            raise ValueError("rdvc_ec_encode_with_indexes: " + _cabi.last_error())
        return out[:n].tobytes()

    def decompress_symbols(self, data: bytes, size: Tuple[int, int]) -> np.ndarray:
        cdfs, lengths, offsets = self._tables()
        shape = (self.channels, int(size[0]), int(size[1]))
        idx = np.ascontiguousarray(np.broadcast_to(np.arange(self.channels, dtype=np.int32)[:, None, None], shape))
        out = np.empty(shape, np.int32)
        buf = np.frombuffer(data, np.uint8) if len(data) else np.zeros(1, np.uint8)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib = _cabi.load()
        _cabi.check(lib.rdvc_ec_decode_with_indexes(p(buf), len(data), p(idx), out.size, p(cdfs), p(lengths), p(offsets),
                                                    self.channels, cdfs.shape[1], p(out)), "rdvc_ec_decode_with_indexes")
        return out

    def compress(self, y) -> bytes:
        """(C, H, W) latents -> one byte string (the reference keeps ``strings[0]``, R:codec_processing.py:492)."""
        return self.compress_symbols(self.quantize(y))

    def decompress(self, data: bytes, size: Tuple[int, int]) -> np.ndarray:
        """Byte string + latent (H, W) -> dequantised (C, H, W) latents (R:codec_processing.py:509-536)."""
        return self.dequantize(self.decompress_symbols(data, size))

    # -- what the byte count should be ---------------------------------------------------------------------
    def cross_entropy_bits(self, symbols: np.ndarray) -> float:
        """Ideal code length of in-table symbols under the quantised tables, in bits (escapes count their table slot
        plus 4 bits per bypass digit)."""
        cdfs, lengths, offsets = self._tables()
        bits = 0.0
        for c in range(self.channels):
            v = symbols[c].astype(np.int64).reshape(-1) - offsets[c]
            mx = lengths[c] - 2
            esc = (v < 0) | (v >= mx)
            vv = np.where(esc, mx, v)
            f = (cdfs[c, vv + 1].astype(np.int64) - cdfs[c, vv].astype(np.int64)).astype(np.float64)
            bits += float(np.sum(PRECISION - np.log2(f)))
            if esc.any():
                raw = np.where(v < 0, -2 * v - 1, 2 * (v - mx))[esc]
                digits = np.ceil(np.log2(raw.astype(np.float64) + 1) / 4).astype(np.int64)
                bits += float(np.sum(4 * (digits + 1 + digits // 15)))
        return bits


class FlowCoder(FactorizedPrior):
    """The stand-in used by bench_gop.py for the motion bitstream of a P-frame: a 2-channel factorised prior over the
    quantised 1/8-resolution flow (int8 symbols)."""

    def __init__(self, scale: float = 6.0):
        super().__init__(2, scales=[scale, scale])
        self.update()

    def compress(self, q) -> bytes:              # q: (2, h, w) integer array
        return self.compress_symbols(np.asarray(q).astype(np.int32))

    def decompress(self, data: bytes, size: Tuple[int, int]) -> np.ndarray:
        return self.decompress_symbols(data, size)
