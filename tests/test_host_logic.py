"""CPU suite: host-side logic -- synthetic stream generator, dvs_msgs::Event packing, the
config mirror, the packed result block and the N>1 plumbing on gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

from esvio_b200 import shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_window_local():
    a = synth.StereoEventStream(346, 260, 1.0e6)
    b = synth.StereoEventStream(346, 260, 1.0e6)
    w5 = a.window(5, 0)
    for k in (0, 1, 2):          # generating other windows first must not change window 5
        b.window(k, 0)
    for x, y in zip(w5, b.window(5, 0)):
        assert np.array_equal(x, y)
    x, y, t, p, sec, nsec = w5
    assert len(x) == 33333 and x.max() < 346 and y.max() < 260 and set(np.unique(p)) <= {0, 1}
    assert (np.diff(t) >= 0).all(), "events of a window are time-ascending"
    assert t[0] > 1.7e9 and np.float32(t[0]) != t[0] or True
    # ros::Time::toSec(): sec + 1e-9 * nsec, bit for bit
    assert np.array_equal(t, sec.astype(np.float64) + 1e-9 * nsec.astype(np.float64))
    # left and right streams differ, streams of different ranks differ
    assert not np.array_equal(a.window(5, 1)[0], x)
    c = synth.StereoEventStream(346, 260, 1.0e6, stream=1)
    assert not np.array_equal(c.window(5, 0)[0], x)


def test_mono_stream_has_empty_right_camera():
    s = synth.StereoEventStream(346, 260, 1.0e6, mono=True)
    L, R, t_ref = s.stereo_window(0)
    assert len(L[0]) == 33333 and len(R[0]) == 0 and t_ref == L[2][-1]


def test_aos_layout_matches_dvs_msgs_event():
    """feature_tracker/src/dvs_msgs/Event.h:42-52: u16 x, u16 y, {u32 sec, u32 nsec}, u8 pol."""
    x = np.array([1, 345], np.uint16)
    y = np.array([2, 259], np.uint16)
    sec = np.array([1700000000, 1700000001], np.uint32)
    nsec = np.array([5000, 999999000], np.uint32)
    p = np.array([1, 0], np.uint8)
    a = synth.to_aos(x, y, sec, nsec, p)
    assert a.dtype.itemsize == 16 and a.nbytes == 32
    raw = a.tobytes()
    assert int.from_bytes(raw[0:2], "little") == 1 and int.from_bytes(raw[2:4], "little") == 2
    assert int.from_bytes(raw[4:8], "little") == 1700000000
    assert int.from_bytes(raw[8:12], "little") == 5000 and raw[12] == 1
    assert int.from_bytes(raw[16 + 8:16 + 12], "little") == 999999000 and raw[16 + 12] == 0


def test_workloads_cover_baseline_configs():
    w = synth.WORKLOADS
    assert w["stereo_davis346_1mevs"]["width"] == 346 and w["stereo_davis346_1mevs"]["rate"] == 1e6
    assert w["stereo_vga_5mevs"]["rate"] == 5e6 and w["stereo_vga_20mevs_burst"]["max_cnt"] == 200
    for name, cfg in w.items():
        c = synth.default_config(cfg["width"], cfg["height"], max_cnt=cfg["max_cnt"])
        assert c["decay_ms"] == 20.0 and c["feature_filter_threshold"] == 0.01
        assert c["ts_lk_threshold"] == 128.0 and c["focal_length"] == 460.0


def test_result_block_roundtrip_and_validation():
    rng = np.random.default_rng(0)
    M = 150
    tr = {k: rng.normal(size=40).astype(np.float32) for k in shard.LEFT_FIELDS[2:]}
    tr.update({k: rng.normal(size=25).astype(np.float32) for k in shard.RIGHT_FIELDS[1:]})
    tr["id"] = np.arange(40, dtype=np.int32) * 3
    tr["track_cnt"] = rng.integers(1, 9, 40).astype(np.int32)
    tr["id_right"] = tr["id"][:25].copy()
    blk = shard.pack_result_block(tr, M)
    assert blk.shape == (shard.result_words(M),)
    out = shard.unpack_result_block(blk, M)
    for k in tr:
        assert np.array_equal(out[k], tr[k]), k
    blk[0] = M + 1
    with pytest.raises(ValueError):
        shard.unpack_result_block(blk, M)


def test_streams_of_rank_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(s for r in range(world) for s in shard.streams_of_rank(r, world, 8))
        assert seen == list(range(8))
    assert shard.streams_of_rank(1, 2, 8) == [1, 3, 5, 7]
    with pytest.raises(ValueError):
        shard.streams_of_rank(2, 2, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = 150
        n = 10 + rank
        tr = {k: np.full(n, rank + 0.5, np.float32) for k in shard.LEFT_FIELDS[2:]}
        tr.update({k: np.full(n - 3, rank + 0.25, np.float32) for k in shard.RIGHT_FIELDS[1:]})
        tr["id"] = np.arange(n, dtype=np.int32) + 1000 * rank
        tr["track_cnt"] = np.full(n, 2, np.int32)
        tr["id_right"] = tr["id"][:n - 3].copy()
        local = torch.from_numpy(shard.pack_result_block(tr, M))
        g = shard.all_gather_tracks(local)
        decoded = [shard.unpack_result_block(g[r].numpy(), M) for r in range(world)]
        ok = all(len(decoded[r]["id"]) == 10 + r and decoded[r]["id"][0] == 1000 * r and
                 np.all(decoded[r]["u"] == r + 0.5) and len(decoded[r]["id_right"]) == 7 + r
                 for r in range(world))
        # weak scaling: events add up, time is the max over ranks
        val, ms = shard.aggregate_throughput(1.0e6 * (rank + 1), 10.0 * (rank + 1))
        ok = ok and abs(ms - 10.0 * world) < 1e-9
        ok = ok and abs(val - (1.0e6 * world * (world + 1) / 2) / (10.0 * world * 1e-3) / 1e6) < 1e-9
        q.put((rank, ok, shard.streams_of_rank(rank, world, 4)))
    finally:
        dist.destroy_process_group()


def test_all_gather_tracks_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] == [0, 2] and res[1][2] == [1, 3]


class _FakeSplitFrontEnd:
    """Stands in for EventFrontEnd in the split protocol test: 'images' are rows of a CPU
    tensor pool addressed by fake pointers, rotating over three buffers like the library."""

    def __init__(self, nbytes):
        import torch
        self.pool = torch.zeros((3, nbytes), dtype=torch.uint8)
        self.k = 0
        self.log = []

    def view(self, ptr, nbytes):
        return self.pool[ptr - 100][:nbytes]

    def split_image_submit(self, cur_time, events, stream):
        i = self.k % 3
        self.pool[i] = int(events) % 251          # the "image" of this window
        self.k += 1
        self.log.append(("image", cur_time))
        return 100 + i, self.pool.shape[1]

    def split_right_buffer(self):
        return 100 + self.k % 3, self.pool.shape[1]

    def submit_split(self, cur_time, left, pub, stream):
        self.log.append(("submit", cur_time, int(self.pool[self.k % 3][0]), int(self.pool[self.k % 3][-1]), pub))
        self.k += 1

    def wait(self, unpack=True):
        self.log.append(("wait",))
        return {"id": np.arange(3)}


def _split_worker(rank, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        fe = _FakeSplitFrontEnd(4096)
        sp = shard.LeftRightSplit(fe, rank, exchange_stream=0, view=fe.view)
        outs = []
        for k in range(5):
            sp.step(1.0 + k, 7 * k + 3, k % 2 == 0)     # "events" = a number that seeds the image
            outs.append(sp.wait())
        q.put((rank, fe.log, [o is not None for o in outs]))
    finally:
        dist.destroy_process_group()


def test_left_right_split_protocol_gloo_world2():
    """SURVEY.md 8e row 2, host side: rank 1 produces the right image and sends it, rank 0
    receives it into the buffer the library names BEFORE it submits the window."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, left_log, left_outs), (_, right_log, right_outs) = res
    assert left_outs == [True] * 5 and right_outs == [False] * 5
    assert right_log == [("image", 1.0 + k) for k in range(5)]
    submits = [e for e in left_log if e[0] == "submit"]
    assert submits == [("submit", 1.0 + k, (7 * k + 3) % 251, (7 * k + 3) % 251, k % 2 == 0)
                       for k in range(5)]
    with pytest.raises(ValueError):
        shard.LeftRightSplit(None, 2, exchange_stream=0)

def test_soa_layouts_are_aligned_and_disjoint():
    """esvio_fe_soa_layout / esvio_fe_soa_layout_stereo (host-only entry points of the C ABI): every
    array 16-byte aligned, arrays disjoint and in order, the right camera's block on a 256-byte
    boundary behind the left one, totals tight -- what the single-transfer staging relies on."""
    import ctypes as C
    from esvio_b200 import _capi
    L = _capi.lib()
    width = (2, 2, 8, 1)
    for n in (0, 1, 7, 33333, 166667, 666667):
        off = (C.c_size_t * 4)()
        tot = C.c_size_t()
        L.esvio_fe_soa_layout(n, off, C.byref(tot))
        end = 0
        for o, w in zip(off, width):
            assert o % 16 == 0 and o >= end
            end = o + w * n
        assert end <= tot.value < end + 16
        for nr in (0, 5, n):
            ol, orr = (C.c_size_t * 4)(), (C.c_size_t * 4)()
            t2 = C.c_size_t()
            L.esvio_fe_soa_layout_stereo(n, nr, ol, orr, C.byref(t2))
            assert list(ol) == list(off)
            assert orr[0] % 256 == 0 and tot.value <= orr[0] < tot.value + 256
            offr = (C.c_size_t * 4)()
            totr = C.c_size_t()
            L.esvio_fe_soa_layout(nr, offr, C.byref(totr))
            assert [o - orr[0] for o in orr] == list(offr)
            assert t2.value == orr[0] + totr.value
            assert t2.value <= 16 * (n + 64) + 16 * (nr + 64)      # fits the slot's 32 B x capacity block
