"""CPU tests (-m "not gpu"): `.rdvc` container round trips and the GOP-sharded gather, including a
real world_size=2 gloo run (the N>1 host path; there is no data-path collective to test)."""
import io
import os
import socket
import struct

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rdvc_corr_b200 as rc

fmt = rc.rdvc_format
gs = rc.gop_shard


# ------------------------------------------------------------------ container format
def test_record_layout_is_the_references():
    """Byte layout of R:codec_processing.py:1398-1418 / :1485-1495, spelled out by hand."""
    i_payload = fmt.iframe_payload(b"\xff\xd8JPEG", ".jpg")
    assert i_payload == b"\x04.jpg\xff\xd8JPEG"
    rec = fmt.FrameRecord(7, "I", i_payload).pack()
    assert rec == b"RDVCFRME" + struct.pack(">I", 7) + b"I" + struct.pack(">Q", len(i_payload)) + i_payload
    p_payload = fmt.pframe_payload((17, 30), b"mm", (68, 120), b"rrr")
    assert p_payload == (struct.pack(">i", 17) + struct.pack(">i", 30) + struct.pack(">I", 2) + b"mm" +
                         struct.pack(">i", 68) + struct.pack(">i", 120) + struct.pack(">I", 3) + b"rrr")
    assert fmt.parse_pframe_payload(p_payload) == ((17, 30), b"mm", (68, 120), b"rrr")
    assert fmt.parse_iframe_payload(i_payload) == (".jpg", b"\xff\xd8JPEG")


def test_stream_round_trip_and_errors():
    recs = [fmt.FrameRecord(0, "I", fmt.iframe_payload(b"abc")),
            fmt.FrameRecord(1, "P", fmt.pframe_payload((2, 3), b"m" * 5, (4, 6), b"r" * 9)),
            fmt.FrameRecord(2, "P", fmt.pframe_payload((2, 3), b"", (4, 6), b""))]
    data = fmt.write_stream({"rdvc_version": "1.0", "iframe_interval": 10}, [r.pack() for r in recs])
    assert data.startswith(b"RDVCMETA") and data.endswith(b"RDVCEND_")
    meta, out = fmt.read_stream(data)
    assert meta["rdvc_version"] == "1.0" and out == recs
    assert fmt.pframe_payload_bytes(out) == 14
    with pytest.raises(ValueError, match="METADATA marker"):
        fmt.read_stream(b"XXXXXXXX" + data[8:])
    with pytest.raises(ValueError, match="FRAME marker"):
        list(fmt.iter_frames(io.BytesIO(b"BADMARK_" + b"\0" * 20)))
    with pytest.raises(EOFError):
        list(fmt.iter_frames(io.BytesIO(recs[1].pack()[:-3])))
    with pytest.raises(ValueError):
        fmt.FrameRecord(0, "B", b"").pack()
    empty_meta, empty = fmt.read_stream(fmt.write_stream({}, []))
    assert empty == [] and empty_meta == {}


# ------------------------------------------------------------------ GOP partition
def test_split_gops():
    g = gs.split_gops(600, 10)
    assert len(g) == 60 and g[0] == gs.Gop(0, 0, 10) and g[-1] == gs.Gop(59, 590, 600)
    r = gs.split_gops(25, 10)                       # ragged tail
    assert [(x.start, x.stop, x.num_pframes) for x in r] == [(0, 10, 9), (10, 20, 9), (20, 25, 4)]
    assert gs.split_gops(0, 10) == []
    assert gs.split_gops(1, 5)[0].num_pframes == 0
    with pytest.raises(ValueError):
        gs.split_gops(10, 0)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_assign_gops_is_a_balanced_partition(world):
    gops = gs.split_gops(600, 10)
    parts = gs.assign_gops(gops, world)
    flat = sorted(g.index for p in parts for g in p)
    assert flat == list(range(60))                                  # partition: every GOP exactly once
    loads = [sum(g.num_pframes for g in p) for p in parts]
    assert max(loads) - min(loads) <= 9                             # within one GOP
    assert gs.assign_gops(gops, world) == parts                     # deterministic on every rank
    assert all(p == sorted(p, key=lambda g: g.index) for p in parts)
    # more ranks than GOPs: the extras get nothing
    few = gs.assign_gops(gs.split_gops(15, 10), 8)
    assert sum(len(p) for p in few) == 2 and sum(1 for p in few if not p) == 6


def _fake_encoders():
    frames = lambda t: bytes([t % 251]) * 4
    enc_i = lambda f: fmt.iframe_payload(b"J" + f)
    enc_p = lambda prev, cur: fmt.pframe_payload((1, 2), prev, (3, 4), cur)
    return frames, enc_i, enc_p


def test_encode_gop_open_loop_and_failure_recovery():
    frames, enc_i, enc_p = _fake_encoders()
    recs = list(fmt.iter_frames(io.BytesIO(gs.encode_gop(gs.Gop(2, 20, 25), frames, enc_i, enc_p))))
    assert [(r.index, r.kind) for r in recs] == [(20, "I"), (21, "P"), (22, "P"), (23, "P"), (24, "P")]
    # open loop: the P-frame of t references the ORIGINAL frame t-1
    assert fmt.parse_pframe_payload(recs[2].payload)[1] == frames(21)

    def flaky(prev, cur):
        if cur == frames(22):
            raise RuntimeError("boom")
        return enc_p(prev, cur)
    recs = list(fmt.iter_frames(io.BytesIO(gs.encode_gop(gs.Gop(2, 20, 25), frames, enc_i, flaky))))
    assert [(r.index, r.kind) for r in recs] == [(20, "I"), (21, "P"), (22, "P"), (23, "I"), (24, "P")]


def _serial_stream(num_frames, interval, meta):
    frames, enc_i, enc_p = _fake_encoders()
    gops = gs.split_gops(num_frames, interval)
    local = {g.index: gs.encode_gop(g, frames, enc_i, enc_p) for g in gops}
    return gs.gather_stream(local, len(gops), meta)


def test_single_rank_stream():
    data = _serial_stream(25, 10, {"rdvc_version": "1.0"})
    meta, recs = fmt.read_stream(data)
    assert [r.index for r in recs] == list(range(25))
    assert [r.kind for r in recs] == ["I" if t % 10 == 0 else "P" for t in range(25)]
    assert meta["total_frames_processed"] == 25
    assert meta["total_pframe_payload_bytes"] == 22 * 8
    with pytest.raises(RuntimeError, match="missing"):
        gs.gather_stream({0: b""}, 2, {})


# ------------------------------------------------------------------ world_size = 2 over gloo
def _worker(rank, world, port, num_frames, interval, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames, enc_i, enc_p = _fake_encoders()
    gops = gs.split_gops(num_frames, interval)
    mine = gs.assign_gops(gops, world)[rank]
    local = {g.index: gs.encode_gop(g, frames, enc_i, enc_p) for g in mine}
    data = gs.gather_stream(local, len(gops), {"rdvc_version": "1.0"}, rank, world)
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(data)
    else:
        assert data is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(180)
def test_two_rank_gather_equals_serial(tmp_path):
    out = str(tmp_path / "two_rank.rdvc")
    mp.spawn(_worker, args=(2, _free_port(), 47, 5, out), nprocs=2, join=True)
    with open(out, "rb") as f:
        sharded = f.read()
    assert sharded == _serial_stream(47, 5, {"rdvc_version": "1.0"})   # byte-identical to 1-rank encode


def test_encode_gop_batched_matches_per_frame_and_falls_back():
    """encode_gop_batched: one call for all P-frames of a GOP gives the same records as the per-frame walk;
    a failing batch falls back to the per-frame path (with the reference's failure rule)."""
    gop = gs.Gop(3, 30, 36)
    frames = lambda t: t
    enc_i = lambda f: fmt.iframe_payload(b"I%d" % f)
    enc_p = lambda prev, cur: fmt.pframe_payload((1, 2), b"m%d-%d" % (prev, cur), (3, 4), b"r")
    want = gs.encode_gop(gop, frames, enc_i, enc_p)
    calls = []

    def enc_batch(prevs, curs):
        calls.append((list(prevs), list(curs)))
        return [enc_p(a, b) for a, b in zip(prevs, curs)]

    assert gs.encode_gop_batched(gop, frames, enc_i, enc_batch) == want
    assert calls == [([30, 31, 32, 33, 34], [31, 32, 33, 34, 35])]
    def broken(prevs, curs):
        raise RuntimeError("batch failed")
    assert gs.encode_gop_batched(gop, frames, enc_i, broken, enc_p) == want
    with pytest.raises(RuntimeError, match="batch failed"):
        gs.encode_gop_batched(gop, frames, enc_i, broken)
    assert gs.encode_gop_batched(gs.Gop(0, 5, 5), frames, enc_i, enc_batch) == b""
    lone = gs.encode_gop_batched(gs.Gop(0, 7, 8), frames, enc_i, enc_batch)
    assert [r.kind for r in fmt.iter_frames(io.BytesIO(lone))] == ["I"]


# ------------------------------------------------------------------ frame-level sharding (spans)
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("frames_interval", [(600, 10), (47, 5), (25, 10), (7, 32), (9, 1)])
def test_assign_frames_is_a_balanced_contiguous_partition(world, frames_interval):
    n, I = frames_interval
    spans = gs.assign_frames(n, I, world)
    assert len(spans) == world and [t for sp in spans for t in sp] == list(range(n))     # contiguous partition
    loads = [sum(1 for t in sp if t % I != 0) for sp in spans]
    assert max(loads) - min(loads) <= 1                                                  # P-frames balanced to one
    assert gs.assign_frames(n, I, world) == spans
    for sp in spans:                      # a span never starts right after "its" I-frame went to the previous rank
        if len(sp) and sp.start > 0 and sp.start % I != 0:
            assert (sp.start - 1) % I != 0
    if (n, I, world) == (600, 10, 8):
        assert max(loads) == 68 and min(loads) == 67                                     # 540 / 8: 7.94x, not 7.5x


def _span_stream(num_frames, interval, world, enc_p_factory=None, batch=4, batched=True, runs=False, seen=None):
    frames, enc_i, enc_p = _fake_encoders()
    if enc_p_factory is not None:
        enc_p = enc_p_factory(frames, enc_p)

    def enc_batch_fn(prevs, curs):
        if seen is not None:
            seen.append((list(prevs), list(curs)))
        return [enc_p(a, b) for a, b in zip(prevs, curs)]
    enc_batch = enc_batch_fn if batched else None
    spans = gs.assign_frames(num_frames, interval, world)
    out, failed = [], []
    for sp in spans:
        data, tail = gs.encode_span(sp, interval, frames, enc_i, enc_p, enc_batch, batch=batch, consecutive_runs=runs)
        out.append(data); failed.append(tail)
    if world == 1:
        return gs.gather_spans(out[0], failed[0], spans, {"rdvc_version": "1.0"}, 0, 1)
    # emulate the multi-rank gather in one process: rank 0 receives every rank's (bytes, tail_failed)
    import unittest.mock as um
    fake = list(zip(out, failed))
    with um.patch.object(gs, "_gather_bytes", lambda local, flag, rank, world, group=None: fake):
        return gs.gather_spans(out[0], failed[0], spans, {"rdvc_version": "1.0"}, 0, world,
                               reencode_iframe=lambda t: enc_i(frames(t)))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_span_sharding_equals_gop_serial_stream(world):
    """Frame spans (batched across GOP boundaries) give the SAME bytes as the serial whole-GOP encode."""
    for (n, I) in [(47, 5), (25, 10), (12, 4)]:
        assert _span_stream(n, I, world) == _serial_stream(n, I, {"rdvc_version": "1.0"}), (n, I, world)
        assert _span_stream(n, I, world, batched=False) == _serial_stream(n, I, {"rdvc_version": "1.0"})


def test_pframe_batches():
    assert gs.pframe_batches([1, 2, 3, 4, 6, 7], 4) == [[1, 2, 3, 4], [6, 7]]
    assert gs.pframe_batches([1, 2, 3, 4, 6, 7], 3) == [[1, 2, 3], [4, 6, 7]]
    assert gs.pframe_batches([1, 2, 3, 4, 6, 7], 3, True) == [[1, 2, 3], [4], [6, 7]]
    assert gs.pframe_batches([8, 9, 11, 12, 13], 9, True) == [[8, 9], [11, 12, 13]]
    assert gs.pframe_batches([], 9, True) == []
    with pytest.raises(ValueError):
        gs.pframe_batches([1], 0)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_consecutive_run_batches_give_the_same_stream(world):
    """`consecutive_runs`: every batch is a run of consecutive frames whose frames are handed over ONCE (the same
    object ends one pair and starts the next, so a callee can share per-frame work), batches never straddle an
    I-frame, and the stream is byte-identical to the serial encode."""
    for (n, I, batch) in [(47, 5, 4), (25, 10, 9), (12, 4, 2), (30, 10, 4)]:
        seen = []
        assert _span_stream(n, I, world, batch=batch, runs=True, seen=seen) == _serial_stream(n, I, {"rdvc_version": "1.0"})
        p_total = 0
        for prevs, curs in seen:
            assert 1 <= len(curs) <= batch and len(prevs) == len(curs)
            assert all(prevs[i + 1] is curs[i] for i in range(len(curs) - 1))
            p_total += len(curs)
        assert p_total == n - len(gs.split_gops(n, I))


def test_deferred_payloads_are_settled_after_the_next_batch_is_submitted():
    """encode_pframes may return a callable: encode_span calls it only after submitting the next batch (so host work
    overlaps device work) and the stream is the same; an exception raised by the deferred half takes the usual route."""
    n, I = 30, 10
    frames, enc_i, enc_p = _fake_encoders()
    events = []

    def enc_batch(prevs, curs):
        k = len([e for e in events if e[0] == "submit"])
        events.append(("submit", k))
        prevs, curs = list(prevs), list(curs)

        def finish():
            events.append(("finish", k))
            return [enc_p(a, b) for a, b in zip(prevs, curs)]
        return finish
    data, tail = gs.encode_span(range(0, n), I, frames, enc_i, enc_p, enc_batch, batch=4, consecutive_runs=True)
    ref, _ = gs.encode_span(range(0, n), I, frames, enc_i, enc_p, None)
    assert data == ref and tail is False
    n_b = len([e for e in events if e[0] == "submit"])
    assert n_b == 9 and [e for e in events if e[0] == "finish"] == [("finish", k) for k in range(n_b)]
    for k in range(n_b - 1):
        assert events.index(("submit", k + 1)) < events.index(("finish", k))       # overlap: next submitted first

    def enc_batch_bad(prevs, curs):
        def finish():
            raise RuntimeError("late failure")
        return finish
    data2, _ = gs.encode_span(range(0, n), I, frames, enc_i, enc_p, enc_batch_bad, batch=4)
    assert data2 == ref                               # redone frame by frame through encode_pframe


def test_consecutive_runs_failure_rule_across_a_cut():
    n, I = 20, 10
    frames, enc_i, enc_p = _fake_encoders()
    for bad in (3, 9, 11, 19):
        def factory(frames_, enc_p_, bad=bad):
            def flaky(prev, cur):
                if cur == frames_(bad):
                    raise RuntimeError("boom")
                return enc_p_(prev, cur)
            return flaky
        gops = gs.split_gops(n, I)
        serial = gs.gather_stream({g.index: gs.encode_gop(g, frames, enc_i, factory(frames, enc_p)) for g in gops},
                                  len(gops), {"rdvc_version": "1.0"})
        for world in (2, 3):
            assert _span_stream(n, I, world, factory, runs=True) == serial, (bad, world)


def test_span_sharding_failure_rule_across_a_cut():
    """A failing P-frame forces the next frame to I even when the cut between two ranks falls right after it."""
    n, I = 20, 10
    frames, enc_i, enc_p = _fake_encoders()
    for bad in range(1, n):
        if bad % I == 0:
            continue

        def factory(frames_, enc_p_, bad=bad):
            def flaky(prev, cur):
                if cur == frames_(bad):
                    raise RuntimeError("boom")
                return enc_p_(prev, cur)
            return flaky
        gops = gs.split_gops(n, I)
        serial = gs.gather_stream({g.index: gs.encode_gop(g, frames, enc_i, factory(frames, enc_p)) for g in gops},
                                  len(gops), {"rdvc_version": "1.0"})
        for world in (2, 3, 6):
            assert _span_stream(n, I, world, factory) == serial, (bad, world)


def _span_worker(rank, world, port, num_frames, interval, out_path, shm=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if shm:
        os.environ["LOCAL_WORLD_SIZE"] = str(world)     # what torchrun sets on a single node
    else:
        os.environ.pop("LOCAL_WORLD_SIZE", None)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames, enc_i, enc_p = _fake_encoders()
    spans = gs.assign_frames(num_frames, interval, world)
    data, tail = gs.encode_span(spans[rank], interval, frames, enc_i, enc_p,
                                lambda prevs, curs: [enc_p(a, b) for a, b in zip(prevs, curs)], batch=3)
    out = gs.gather_spans(data, tail, spans, {"rdvc_version": "1.0"}, rank, world,
                          reencode_iframe=lambda t: enc_i(frames(t)))
    if rank == 0:
        with open(out_path, "wb") as f:
            f.write(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("shm", [False, True], ids=["tensor-gather", "dev-shm"])
def test_two_rank_span_gather_equals_serial(tmp_path, shm):
    """Both host-side routes of the byte gather (gloo tensors; /dev/shm files on a single node) over real processes."""
    if shm and not os.access("/dev/shm", os.W_OK):
        pytest.skip("/dev/shm not writable")
    out = str(tmp_path / "two_rank_spans.rdvc")
    port = _free_port()
    mp.spawn(_span_worker, args=(2, port, 47, 5, out, shm), nprocs=2, join=True)
    with open(out, "rb") as f:
        assert f.read() == _serial_stream(47, 5, {"rdvc_version": "1.0"})
    import glob
    assert not glob.glob(f"/dev/shm/rdvc_gather_{port}_*")          # rank 0 removed what it read


def test_encode_gop_logs_and_reraises_fatal_errors(caplog):
    """A failed P-frame is logged with its frame index (the reference prints the traceback,
    R:codec_processing.py:1501-1506); errors a retry cannot fix are re-raised instead of silently degrading."""
    frames, enc_i, enc_p = _fake_encoders()

    def flaky(prev, cur):
        raise RuntimeError("boom")
    failures = []
    with caplog.at_level("ERROR", logger="rdvc_corr_b200.gop_shard"):
        gs.encode_gop(gs.Gop(0, 0, 3), frames, enc_i, flaky, failures)
    assert failures == [1] and "P-frame 1 failed" in caplog.text and "boom" in caplog.text

    def fatal(prev, cur):
        raise RuntimeError("CUDA error: an illegal memory access was encountered")
    with pytest.raises(RuntimeError, match="CUDA error"):
        gs.encode_gop(gs.Gop(0, 0, 3), frames, enc_i, fatal)


def test_scan_frames_matches_the_copying_reader():
    recs = [fmt.FrameRecord(0, "I", fmt.iframe_payload(b"abc")),
            fmt.FrameRecord(1, "P", fmt.pframe_payload((2, 3), b"m" * 5, (4, 6), b"r" * 9)),
            fmt.FrameRecord(2, "P", fmt.pframe_payload((2, 3), b"", (4, 6), b""))]
    data = b"".join(r.pack() for r in recs)
    scanned = fmt.scan_frames(data)
    assert [(i, k) for i, k, _, _ in scanned] == [(0, "I"), (1, "P"), (2, "P")]
    assert [data[o:o + n] for _, _, o, n in scanned] == [r.payload for r in recs]
    assert sum(fmt.pframe_bitstream_bytes(data, o) for _, k, o, _ in scanned if k == "P") == fmt.pframe_payload_bytes(recs) == 14
    assert fmt.scan_frames(data + fmt.EOF_MARKER) == scanned and fmt.scan_frames(b"") == []
    with pytest.raises(EOFError):
        fmt.scan_frames(data[:-3])
    with pytest.raises(ValueError, match="FRAME marker"):
        fmt.scan_frames(b"BADMARK_" + b"\0" * 20)


def test_stream_parts_equal_write_stream():
    recs = [fmt.FrameRecord(0, "I", fmt.iframe_payload(b"abc")).pack(), fmt.FrameRecord(1, "P", fmt.pframe_payload((2, 3), b"mm", (4, 6), b"r")).pack()]
    want = fmt.write_stream({"rdvc_version": "1.0", "n": 2}, recs)
    parts = fmt.StreamParts({"rdvc_version": "1.0", "n": 2}, [memoryview(recs[0]), bytearray(recs[1])])
    assert parts.tobytes() == want and len(parts) == len(want)
    buf = io.BytesIO()
    assert parts.write_to(buf) == len(want) and buf.getvalue() == want
    frames, enc_i, enc_p = _fake_encoders()
    spans = gs.assign_frames(12, 4, 1)
    data, tail = gs.encode_span(spans[0], 4, frames, enc_i, enc_p)
    p = gs.gather_spans(data, tail, spans, {"rdvc_version": "1.0"}, as_parts=True)
    assert isinstance(p, fmt.StreamParts) and p.tobytes() == _serial_stream(12, 4, {"rdvc_version": "1.0"})
