"""GOP sharding of RDVC's encoder over the GPUs of one box.

The encoder is open loop -- the reference frame of a P-frame is the ORIGINAL previous frame
(R:codec_processing.py:1498-1499) -- so a GOP (an I-frame plus the P-frames up to the next
I-frame, `iframe_interval` frames, R:codec_processing.py:634,1392) depends on no other GOP.
One process per GPU takes whole GOPs; the motion branch of every P-frame runs RAFT with the
B200 correlation block.  There is NO collective on the data path: the only cross-rank step is
a host-side gather of per-GOP byte strings into the `.rdvc` writer on rank 0.

Two partitions of the same stream (both give a byte-identical `.rdvc` file to the 1-rank encode):

* whole GOPs (`split_gops` / `assign_gops` / `encode_gop[_batched]` / `gather_stream`): the north star's unit.
  60 GOPs on 8 ranks is 8 + 7 x ... = at best 7.5x.
* frame spans (`assign_frames` / `encode_span` / `gather_spans`): because the encoder is open loop, a P-frame
  needs only the two ORIGINAL frames (t-1, t), so the cut between two ranks may fall anywhere -- every rank
  takes one contiguous span of frames holding (almost) the same number of P-frames (540 / 8 -> 67 or 68: 7.94x),
  reads the one frame before its span, and batches its P-frames `batch` at a time across GOP boundaries.
  The only cross-rank coupling is the reference's failure rule (a failed P-frame forces the next frame to I):
  a span reports whether its LAST frame failed and rank 0 re-encodes the first frame of the next span as an
  I-frame in that (rare) case, so the stream stays identical to the serial encode even then.
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


# ---------------------------------------------------------------- frame-level sharding ("frame-level
# sharding is also legal for the encoder")
def is_iframe(t: int, iframe_interval: int) -> bool:
    """R:codec_processing.py:1392."""
    return t % iframe_interval == 0


def assign_frames(num_frames: int, iframe_interval: int, world_size: int) -> List[range]:
    """One contiguous span of frames per rank, balanced by P-frame count (the I-frames cost the GPU nothing):
    rank r owns the P-frames number floor(r P / W) .. floor((r+1) P / W) - 1 of the sequence and the I-frames
    between them; an I-frame sitting right before a span's first P-frame goes with that span.  Deterministic,
    identical on every rank without communication; spans may be empty when there are more ranks than P-frames."""
    if num_frames < 0 or iframe_interval <= 0 or world_size <= 0:
        raise ValueError("num_frames must be >= 0, iframe_interval and world_size > 0")
    p_frames = [t for t in range(num_frames) if not is_iframe(t, iframe_interval)]
    P = len(p_frames)
    cuts = [0]
    for r in range(1, world_size):
        k = r * P // world_size                      # first P-frame (by P index) of rank r
        if k >= P:
            cuts.append(num_frames)
            continue
        c = p_frames[k]
        if c > 0 and is_iframe(c - 1, iframe_interval):
            c -= 1                                   # keep the I-frame with the P-frames that follow it
        cuts.append(max(c, cuts[-1]))
    cuts.append(num_frames)
    if P == 0:                                       # I-frames only: spread them evenly
        cuts = [r * num_frames // world_size for r in range(world_size + 1)]
    return [range(cuts[r], cuts[r + 1]) for r in range(world_size)]


def pframe_batches(ts_p: Sequence[int], batch: int, consecutive_runs: bool = False) -> List[List[int]]:
    """The P-frame indices `ts_p` (ascending) cut into batches of at most `batch`: plain slices, or -- with
    `consecutive_runs` -- runs of consecutive indices (a batch ends at an I-frame or after `batch` frames)."""
    if batch <= 0:
        raise ValueError("batch must be positive")
    ts_p = list(ts_p)
    if not consecutive_runs:
        return [ts_p[k:k + batch] for k in range(0, len(ts_p), batch)]
    out: List[List[int]] = []
    for t in ts_p:
        if out and out[-1][-1] == t - 1 and len(out[-1]) < batch:
            out[-1].append(t)
        else:
            out.append([t])
    return out


def encode_span(span: range, iframe_interval: int, frames: Callable[[int], object],
                encode_iframe: Callable[[object], bytes],
                encode_pframe: Optional[Callable[[object, object], bytes]] = None,
                encode_pframes: Optional[Callable[[Sequence[object], Sequence[object]], Sequence[bytes]]] = None,
                batch: int = 9, failures: Optional[List[int]] = None,
                force_first_i: bool = False, consecutive_runs: bool = False) -> Tuple[bytes, bool]:
    """Frame records of the frames in `span` and whether the span's LAST frame was a failed P-frame (the next
    span's first frame must then become an I-frame, `gather_spans` does that).  With `encode_pframes` all
    P-frames of the span are encoded `batch` at a time, across GOP boundaries (only the last batch may be
    short); if a batch fails and `encode_pframe` is given, the span is redone frame by frame with the
    reference's failure rule (R:codec_processing.py:1501-1506).  `frames(span.start - 1)` is read when the
    span starts with a P-frame: the previous ORIGINAL frame, whichever rank encodes it.
    `encode_pframes` may return the payloads or a zero-argument callable that returns them: the callable is called
    only after the NEXT batch has been submitted, so host-side work (entropy coding after a device -> host copy) of
    one batch overlaps the device work of the next.
    With `consecutive_runs` a batch never straddles an I-frame: it is a run of up to `batch` CONSECUTIVE P-frames
    t0 .. t1, its frames t0 - 1 .. t1 are fetched ONCE and `encode_pframes(fr[:-1], fr[1:])` gets the same objects
    on both sides (`prevs[i + 1] is curs[i]`), so the callee can run the feature encoder once per frame
    (`raft_flow_sequence`).  Same records, same order, either way."""
    if encode_pframe is None and encode_pframes is None:
        raise ValueError("need encode_pframe and/or encode_pframes")
    if batch <= 0:
        raise ValueError("batch must be positive")
    ts = list(span)
    if not ts:
        return b"", False

    def kind(t):
        return "I" if (is_iframe(t, iframe_interval) or (force_first_i and t == ts[0])) else "P"

    if encode_pframes is not None:
        try:
            ts_p = [t for t in ts if kind(t) == "P"]
            payloads: Dict[int, bytes] = {}
            def settle(chunk, out):
                out = list(out() if callable(out) else out)
                if len(out) != len(chunk):
                    raise RuntimeError("encode_pframes returned the wrong number of payloads")
                payloads.update(zip(chunk, out))

            pending = None      # a batch whose payloads are still being produced (encode_pframes returned a callable)
            for chunk in pframe_batches(ts_p, batch, consecutive_runs):
                if consecutive_runs:
                    fr = [frames(t) for t in range(chunk[0] - 1, chunk[-1] + 1)]
                    out = encode_pframes(fr[:-1], fr[1:])
                else:
                    out = encode_pframes([frames(t - 1) for t in chunk], [frames(t) for t in chunk])
                if pending is not None:     # ... finish the previous batch while this one runs on the device
                    settle(*pending)
                    pending = None
                if callable(out):
                    pending = (chunk, out)
                else:
                    settle(chunk, out)
            if pending is not None:
                settle(*pending)
            recs = [fmt.FrameRecord(t, "I", encode_iframe(frames(t))).pack() if kind(t) == "I"
                    else fmt.FrameRecord(t, "P", payloads[t]).pack() for t in ts]
            return b"".join(recs), False
        except Exception as exc:
            log.error("batched P-frames of span [%d, %d) failed (%s: %s); redoing the span frame by frame\n%s",
                      ts[0], ts[-1] + 1, type(exc).__name__, exc, traceback.format_exc())
            if encode_pframe is None or _is_fatal(exc):
                raise
    parts: List[bytes] = []
    force_i = False
    for t in ts:
        if kind(t) == "I" or force_i:
            parts.append(fmt.FrameRecord(t, "I", encode_iframe(frames(t))).pack())
            force_i = False
            continue
        try:
            parts.append(fmt.FrameRecord(t, "P", encode_pframe(frames(t - 1), frames(t))).pack())
        except Exception as exc:                      # same policy as encode_gop
            log.error("P-frame %d failed (%s: %s); next frame forced to I\n%s", t, type(exc).__name__, exc,
                      traceback.format_exc())
            if _is_fatal(exc):
                raise
            if failures is not None:
                failures.append(t)
            parts.append(fmt.FrameRecord(t, "P", fmt.pframe_payload((0, 0), b"", (0, 0), b"")).pack())
            force_i = True
    return b"".join(parts), force_i


_gather_seq = 0


def _single_node(world_size: int) -> bool:
    import os
    try:
        return int(os.environ.get("LOCAL_WORLD_SIZE", "0")) == world_size and os.access("/dev/shm", os.W_OK)
    except ValueError:
        return False


def _gather_bytes(local: bytes, flag: bool, rank: int, world_size: int, group=None, use_shm: Optional[bool] = None):
    """Every rank's (byte string, flag) on rank 0.  Returns a list of (bytes-like, flag) on rank 0, None elsewhere.

    `gather_object` pickles, copies and un-pickles each 3 MB string several times and gloo moves it over loopback TCP:
    at 8 ranks that was 57 ms on top of 1.45 s of encoding.  Two cheaper host-side routes:
      * all ranks on ONE node (the design point: one box, one process per GPU; LOCAL_WORLD_SIZE == world): every rank
        writes its string to /dev/shm, a tiny all_gather of (length, flag) doubles as the "my file is complete"
        barrier, rank 0 reads the files and removes them -- no socket carries the payload;
      * otherwise: the same all_gather + a gather of zero-padded uint8 tensors."""
    import os
    import numpy as np
    import torch
    import torch.distributed as dist
    global _gather_seq
    _gather_seq += 1
    if use_shm is None:
        use_shm = _single_node(world_size)
    meta = torch.tensor([len(local), int(bool(flag)), int(bool(use_shm))], dtype=torch.int64)
    prefix = f"/dev/shm/rdvc_gather_{os.environ.get('MASTER_PORT', '0')}_{_gather_seq}_"
    if use_shm and rank != 0:
        with open(prefix + str(rank), "wb") as f:
            f.write(local)
    metas = [torch.zeros(3, dtype=torch.int64) for _ in range(world_size)]
    dist.all_gather(metas, meta, group=group)
    shm_all = all(int(m[2]) for m in metas)
    if shm_all:
        if rank != 0:
            return None
        import mmap
        out = [(local, bool(flag))]
        for r in range(1, world_size):
            n = int(metas[r][0])
            with open(prefix + str(r), "rb") as f:
                if os.fstat(f.fileno()).st_size != n:
                    raise RuntimeError(f"rank {r} announced {n} bytes, its file holds {os.fstat(f.fileno()).st_size}")
                # mapped, not read: the pages are already in memory (tmpfs), copying them into a fresh buffer cost
                # ~0.7 ms per MB in page faults; the mapping outlives the unlink
                data = memoryview(mmap.mmap(f.fileno(), 0, prot=mmap.PROT_READ)) if n else memoryview(b"")
            os.unlink(prefix + str(r))
            out.append((data, bool(int(metas[r][1]))))
        return out
    if use_shm and rank != 0:                     # somebody could not use shared memory: fall back together
        os.unlink(prefix + str(rank))
    max_len = max(1, max(int(m[0]) for m in metas))
    mine = torch.zeros(max_len, dtype=torch.uint8)
    if len(local):
        mine.numpy()[:len(local)] = np.frombuffer(local, np.uint8)          # one copy, into the send buffer
    bufs = [torch.empty(max_len, dtype=torch.uint8) for _ in range(world_size)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return [(memoryview(bufs[r].numpy())[:int(metas[r][0])], bool(int(metas[r][1]))) for r in range(world_size)]


def gather_spans(local: bytes, tail_failed: bool, spans: Sequence[range], metadata: dict, rank: int = 0,
                 world_size: int = 1, group=None,
                 reencode_iframe: Optional[Callable[[int], bytes]] = None, as_parts: bool = False):
    """Host-side gather of every rank's span records (rank order = frame order); rank 0 returns the `.rdvc`
    stream -- as one byte string, or with `as_parts=True` as an `rdvc_format.StreamParts` (header, per-rank record
    buffers, end marker: what a writer emits with consecutive `f.write` calls, no 24 MB concatenation).  If span r-1
    ended on a failed P-frame, the first frame of span r is re-encoded as an I-frame with
    `reencode_iframe(t) -> I payload` (the rule a serial encode applies in line)."""
    if world_size > 1:
        gathered = _gather_bytes(local, tail_failed, rank, world_size, group)
        if rank != 0:
            return None
    else:
        gathered = [(local, bool(tail_failed))]
    chunks: List[bytes] = []
    prev_failed = False
    n_records, p_bytes, expect = 0, 0, 0
    want = [t for sp in spans for t in sp]
    for r, (data, failed) in enumerate(gathered):
        recs = fmt.scan_frames(data)                      # headers only: payloads are not copied
        if prev_failed and len(spans[r]) > 0 and recs and recs[0][1] == "P":
            if reencode_iframe is None:
                raise RuntimeError(f"frame {recs[0][0]} must become an I-frame (the frame before it failed) "
                                   "but no reencode_iframe callback was given")
            first = fmt.FrameRecord(recs[0][0], "I", reencode_iframe(recs[0][0])).pack()
            data = first + bytes(data[recs[0][2] + recs[0][3]:])
            recs = fmt.scan_frames(data)
        if len(spans[r]) > 0:
            prev_failed = failed
        for (idx, kind, off, _plen) in recs:
            if expect >= len(want) or idx != want[expect]:
                raise RuntimeError("frame records missing or out of order in the gather")
            expect += 1
            if kind == "P":
                p_bytes += fmt.pframe_bitstream_bytes(data, off)
        n_records += len(recs)
        chunks.append(data)
    if expect != len(want):
        raise RuntimeError("frame records missing or out of order in the gather")
    meta = dict(metadata)
    meta["total_frames_processed"] = n_records
    meta["total_pframe_payload_bytes"] = p_bytes
    parts = fmt.StreamParts(meta, chunks)
    return parts if as_parts else parts.tobytes()
