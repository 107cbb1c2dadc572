"""`.rdvc` container: the byte layout RDVC's encoder writes and its decoder reads.

Needed on the multi-GPU path only: each rank produces the frame records of its GOPs and rank 0
stitches them into one stream.  Layout (all integers big-endian), from
R:codec_processing.py:88-96 (markers), :1398-1418 (I record), :1485-1495 (P record),
:1556-1568 (file assembly) and the reader at :1698-1704, :1749-1772, :1806-1813:

    "RDVCMETA" u32 json_len  json
    repeat:  "RDVCFRME" u32 frame_index  'I'|'P'  u64 payload_len  payload
    "RDVCEND_"

    I payload:  u8 ext_len, ext (".jpg"), image bytes
    P payload:  i32 mH, i32 mW, u32 mLen, motion bytes, i32 rH, i32 rW, u32 rLen, residual bytes

Records are self-delimiting and carry the GLOBAL frame index, so per-GOP byte strings can be
concatenated in GOP order without rewriting.
"""
from __future__ import annotations

import io
import json
import struct
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Tuple

META_MARKER = b"RDVCMETA"
FRAME_MARKER = b"RDVCFRME"
EOF_MARKER = b"RDVCEND_"
_U8, _U32, _I32, _U64 = ">B", ">I", ">i", ">Q"


@dataclass
class FrameRecord:
    index: int
    kind: str                       # "I" or "P"
    payload: bytes

    def pack(self) -> bytes:
        if self.kind not in ("I", "P"):
            raise ValueError(f"frame type must be 'I' or 'P', got {self.kind!r}")
        return b"".join((FRAME_MARKER, struct.pack(_U32, self.index), self.kind.encode("ascii"),
                         struct.pack(_U64, len(self.payload)), self.payload))


def iframe_payload(image_bytes: bytes, ext: str = ".jpg") -> bytes:
    e = ext.encode("utf-8")
    return struct.pack(_U8, len(e)) + e + image_bytes


def parse_iframe_payload(payload: bytes) -> Tuple[str, bytes]:
    n = struct.unpack_from(_U8, payload, 0)[0]
    return payload[1:1 + n].decode("utf-8"), payload[1 + n:]


def pframe_payload(motion_hw: Tuple[int, int], motion: bytes, residual_hw: Tuple[int, int],
                   residual: bytes) -> bytes:
    return b"".join((struct.pack(_I32, motion_hw[0]), struct.pack(_I32, motion_hw[1]),
                     struct.pack(_U32, len(motion)), motion,
                     struct.pack(_I32, residual_hw[0]), struct.pack(_I32, residual_hw[1]),
                     struct.pack(_U32, len(residual)), residual))


def parse_pframe_payload(payload: bytes):
    s = io.BytesIO(payload)

    def part():
        h = struct.unpack(_I32, s.read(4))[0]
        w = struct.unpack(_I32, s.read(4))[0]
        n = struct.unpack(_U32, s.read(4))[0]
        b = s.read(n)
        if len(b) != n:
            raise EOFError("truncated P-frame payload")
        return (h, w), b

    motion_hw, motion = part()
    residual_hw, residual = part()
    return motion_hw, motion, residual_hw, residual


def write_stream(metadata: dict, frame_bytes: Iterable[bytes]) -> bytes:
    """Assemble a complete `.rdvc` byte string from metadata + already-packed frame records."""
    meta = json.dumps(metadata, indent=4).encode("utf-8")
    return b"".join((META_MARKER, struct.pack(_U32, len(meta)), meta, *frame_bytes, EOF_MARKER))


class StreamParts:
    """A complete `.rdvc` stream as the pieces its writer emits in order -- header, packed frame records (any
    bytes-like objects, e.g. memory-mapped per-rank buffers), end marker -- without concatenating them: the reference
    itself writes metadata, frame buffer and end marker with three `f.write` calls (R:codec_processing.py:1556-1568)."""

    def __init__(self, metadata: dict, frame_bytes: Iterable):
        meta = json.dumps(metadata, indent=4).encode("utf-8")
        self.header = b"".join((META_MARKER, struct.pack(_U32, len(meta)), meta))
        self.chunks = list(frame_bytes)

    def __len__(self) -> int:
        return len(self.header) + sum(len(c) for c in self.chunks) + len(EOF_MARKER)

    def write_to(self, f) -> int:
        f.write(self.header)
        for c in self.chunks:
            f.write(c)
        f.write(EOF_MARKER)
        return len(self)

    def tobytes(self) -> bytes:
        return b"".join((self.header, *self.chunks, EOF_MARKER))


def read_stream(data: bytes) -> Tuple[dict, List[FrameRecord]]:
    s = io.BytesIO(data)
    if s.read(len(META_MARKER)) != META_MARKER:
        raise ValueError("Invalid RDVC file: Missing or incorrect METADATA marker at the beginning.")
    n = struct.unpack(_U32, s.read(4))[0]
    meta = json.loads(s.read(n).decode("utf-8"))
    return meta, list(iter_frames(s))


def iter_frames(s) -> Iterator[FrameRecord]:
    while True:
        marker = s.read(len(FRAME_MARKER))
        if not marker or marker == EOF_MARKER:
            return
        if marker != FRAME_MARKER:
            raise ValueError(f"Invalid RDVC file: Missing or incorrect FRAME marker. Found: {marker!r}")
        idx = struct.unpack(_U32, s.read(4))[0]
        kind = s.read(1).decode("ascii")
        n = struct.unpack(_U64, s.read(8))[0]
        payload = s.read(n)
        if len(payload) != n:
            raise EOFError(f"Could not read full frame content for frame {idx}.")
        yield FrameRecord(idx, kind, payload)


def pframe_payload_bytes(records: Iterable[FrameRecord]) -> int:
    """`total_pframe_payload_bytes` of the header: motion + residual bitstream lengths
    (R:codec_processing.py:1535)."""
    total = 0
    for r in records:
        if r.kind == "P":
            _, m, _, res = parse_pframe_payload(r.payload)
            total += len(m) + len(res)
    return total


def scan_frames(data) -> List[Tuple[int, str, int, int]]:
    """(frame index, kind, payload offset, payload length) of every record in a byte string of packed frame records,
    WITHOUT copying the payloads (the gather on rank 0 only needs to check order and add up lengths: at 44 KB per
    P-frame the copying reader cost more than the gather itself)."""
    mv = memoryview(data)
    out = []
    pos, n = 0, len(mv)
    head = len(FRAME_MARKER) + 4 + 1 + 8
    while pos < n:
        marker = bytes(mv[pos:pos + len(FRAME_MARKER)])
        if marker == EOF_MARKER:
            break
        if marker != FRAME_MARKER:
            raise ValueError(f"Invalid RDVC file: Missing or incorrect FRAME marker. Found: {marker!r}")
        if pos + head > n:
            raise EOFError("truncated frame record header")
        idx = struct.unpack_from(_U32, mv, pos + 8)[0]
        kind = chr(mv[pos + 12])
        plen = struct.unpack_from(_U64, mv, pos + 13)[0]
        if pos + head + plen > n:
            raise EOFError(f"Could not read full frame content for frame {idx}.")
        out.append((idx, kind, pos + head, plen))
        pos += head + plen
    return out


def pframe_bitstream_bytes(data, payload_offset: int) -> int:
    """Motion + residual bitstream lengths of the P payload starting at `payload_offset` (no copies)."""
    m_len = struct.unpack_from(_U32, data, payload_offset + 8)[0]
    r_len = struct.unpack_from(_U32, data, payload_offset + 12 + m_len + 8)[0]
    return m_len + r_len
