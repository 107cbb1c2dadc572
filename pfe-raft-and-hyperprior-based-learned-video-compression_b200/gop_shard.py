"""GOP sharding of RDVC's encoder over the GPUs of one box.

The encoder is open loop -- the reference frame of a P-frame is the ORIGINAL previous frame
(R:codec_processing.py:1498-1499) -- so a GOP (an I-frame plus the P-frames up to the next
I-frame, `iframe_interval` frames, R:codec_processing.py:634,1392) depends on no other GOP.
One process per GPU takes whole GOPs; the motion branch of every P-frame runs RAFT with the
B200 correlation block.  There is NO collective on the data path: the only cross-rank step is
a host-side gather of per-GOP byte strings into the `.rdvc` writer on rank 0.
"""
from __future__ import annotations

import logging
import traceback
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

log = logging.getLogger("rdvc_corr_b200.gop_shard")

from . import rdvc_format as fmt


@dataclass(frozen=True)
class Gop:
    index: int          # GOP number
    start: int          # global index of its I-frame
    stop: int           # one past its last frame

    @property
    def num_pframes(self) -> int:
        return max(0, self.stop - self.start - 1)


def split_gops(num_frames: int, iframe_interval: int) -> List[Gop]:
    """Frame t is an I-frame iff t % iframe_interval == 0 (R:codec_processing.py:1392)."""
    if num_frames < 0 or iframe_interval <= 0:
        raise ValueError("num_frames must be >= 0 and iframe_interval > 0")
    return [Gop(g, s, min(s + iframe_interval, num_frames))
            for g, s in enumerate(range(0, num_frames, iframe_interval))]


def assign_gops(gops: Sequence[Gop], world_size: int) -> List[List[Gop]]:
    """Static longest-processing-time assignment by P-frame count: deterministic, identical on
    every rank without communication, and balanced to within one GOP.  (60 GOPs on 8 GPUs
    cannot beat 7.5x with whole GOPs; ragged tails are spread instead of piling on one rank.)"""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    load = [0] * world_size
    out: List[List[Gop]] = [[] for _ in range(world_size)]
    for g in sorted(gops, key=lambda g: (-g.num_pframes, g.index)):
        r = min(range(world_size), key=lambda r: (load[r], r))
        out[r].append(g)
        load[r] += max(1, g.num_pframes)
    for lst in out:
        lst.sort(key=lambda g: g.index)
    return out


def _is_fatal(exc: BaseException) -> bool:
    """Errors a retry on the next frame cannot fix: CUDA runtime failures (sticky) and memory exhaustion."""
    name = type(exc).__name__
    msg = str(exc)
    return isinstance(exc, MemoryError) or name in ("OutOfMemoryError", "AcceleratorError") or "CUDA error" in msg


def encode_gop(gop: Gop, frames: Callable[[int], object], encode_iframe: Callable[[object], bytes],
               encode_pframe: Callable[[object, object], bytes], failures: Optional[List[int]] = None) -> bytes:
    """Frame records of one GOP.  `frames(t)` returns frame t; `encode_iframe(frame)` returns the I
    payload; `encode_pframe(prev_original, cur)` the P payload (motion branch + codec).  A failing
    P-frame turns the NEXT frame into an I-frame, as the reference does
    (R:codec_processing.py:1501-1506); the indices of failed frames are appended to `failures` when given."""
    parts: List[bytes] = []
    prev = None
    force_i = True
    for t in range(gop.start, gop.stop):
        cur = frames(t)
        if force_i or prev is None:
            parts.append(fmt.FrameRecord(t, "I", encode_iframe(cur)).pack())
            force_i = False
        else:
            try:
                parts.append(fmt.FrameRecord(t, "P", encode_pframe(prev, cur)).pack())
            except Exception as exc:
                # the reference prints the error and its traceback and goes on (R:codec_processing.py:1501-1506);
                # here the failed frame additionally gets an EMPTY P record (0x0 shapes, no bitstreams) so that the
                # frame indices of the stream stay dense -- a deliberate divergence: a decoder must treat an empty
                # P record as "repeat the previous reconstruction".  Errors that will not go away by themselves
                # (a sticky CUDA error, out of memory) are re-raised instead of degrading the whole stream silently.
                log.error("P-frame %d failed (%s: %s); next frame forced to I\n%s", t, type(exc).__name__, exc,
                          traceback.format_exc())
                if _is_fatal(exc):
                    raise
                if failures is not None:
                    failures.append(t)
                parts.append(fmt.FrameRecord(t, "P", fmt.pframe_payload((0, 0), b"", (0, 0), b"")).pack())
                force_i = True
        prev = cur   # open loop: the ORIGINAL frame is the next reference
    return b"".join(parts)


def encode_gop_batched(gop: Gop, frames: Callable[[int], object], encode_iframe: Callable[[object], bytes],
                       encode_pframes: Callable[[Sequence[object], Sequence[object]], Sequence[bytes]],
                       encode_pframe: Optional[Callable[[object, object], bytes]] = None,
                       failures: Optional[List[int]] = None) -> bytes:
    """Like :func:`encode_gop`, but all P-frames of the GOP go through ONE call
    `encode_pframes([prev originals], [current frames]) -> [payloads]`: the encoder is open loop, so every
    (previous original, current) pair of a GOP is known up front and RAFT can run them as one batch.  If the
    batched call fails and a per-frame `encode_pframe` is given, the GOP is redone frame by frame with the
    reference's failure rule (:func:`encode_gop`)."""
    ts = list(range(gop.start, gop.stop))
    if not ts:
        return b""
    fr = [frames(t) for t in ts]
    parts: List[bytes] = [fmt.FrameRecord(ts[0], "I", encode_iframe(fr[0])).pack()]
    if len(ts) > 1:
        try:
            payloads = list(encode_pframes(fr[:-1], fr[1:]))
            if len(payloads) != len(ts) - 1:
                raise RuntimeError("encode_pframes returned the wrong number of payloads")
        except Exception as exc:
            log.error("batched P-frames of GOP %d failed (%s: %s); redoing the GOP frame by frame\n%s", gop.index,
                      type(exc).__name__, exc, traceback.format_exc())
            if encode_pframe is None or _is_fatal(exc):
                raise
            return encode_gop(gop, lambda t: fr[t - gop.start], encode_iframe, encode_pframe, failures)
        parts += [fmt.FrameRecord(t, "P", pl).pack() for t, pl in zip(ts[1:], payloads)]
    return b"".join(parts)


def gather_stream(local: Dict[int, bytes], num_gops: int, metadata: dict, rank: int = 0,
                  world_size: int = 1, group=None) -> Optional[bytes]:
    """Host-side gather of {gop index: bytes} from every rank; rank 0 returns the `.rdvc` stream.
    `total_frames_processed` / `total_pframe_payload_bytes` are plain host sums
    (R:codec_processing.py:1528,1535)."""
    if world_size > 1:
        import torch.distributed as dist
        gathered: List[Optional[Dict[int, bytes]]] = [None] * world_size if rank == 0 else None
        dist.gather_object(local, gathered, dst=0, group=group)
        if rank != 0:
            return None
        merged: Dict[int, bytes] = {}
        for d in gathered:
            merged.update(d)
    else:
        merged = dict(local)
    missing = [g for g in range(num_gops) if g not in merged]
    if missing:
        raise RuntimeError(f"GOPs missing from the gather: {missing}")
    ordered = [merged[g] for g in range(num_gops)]
    import io
    records = [r for b in ordered for r in fmt.iter_frames(io.BytesIO(b))]
    meta = dict(metadata)
    meta["total_frames_processed"] = len(records)
    meta["total_pframe_payload_bytes"] = fmt.pframe_payload_bytes(records)
    return fmt.write_stream(meta, ordered)
