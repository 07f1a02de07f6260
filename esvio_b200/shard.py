"""Multi-GPU plumbing of the front-end: the path shards by independent
stereo streams -- one process per GPU, one or more streams per process, no data-path
collective -- plus ONE all-gather of the packed track records per window so that every rank
(or rank 0's adapter) can publish the feature clouds of all streams.

Only `torch.distributed` is used here; the same functions run on NCCL (device tensors that
alias the library's result block) and on gloo (CPU tensors, used by the tests).
"""
from __future__ import annotations

import numpy as np

RESULT_HDR = 32        # int32 words in front of the arrays (csrc/common.cuh kResultHdr)
RESULT_ARRAYS = 15     # id, cnt, u, v, un_x, un_y, vx, vy | id_r, ru, rv, run_x, run_y, rvx, rvy
LEFT_FIELDS = ("id", "track_cnt", "u", "v", "un_x", "un_y", "vx", "vy")
RIGHT_FIELDS = ("id_right", "ru", "rv", "run_x", "run_y", "rvx", "rvy")
INT_FIELDS = ("id", "track_cnt", "id_right")


def result_words(max_cnt: int) -> int:
    return RESULT_HDR + RESULT_ARRAYS * max_cnt


def streams_of_rank(rank: int, world: int, n_streams: int) -> list[int]:
    """Round-robin assignment of stereo streams to ranks (stream s -> rank s % world)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return [s for s in range(n_streams) if s % world == rank]


def pack_result_block(tracks: dict, max_cnt: int) -> np.ndarray:
    """dict of arrays -> the int32 block layout the library writes (for tests / replay)."""
    blk = np.zeros(result_words(max_cnt), np.int32)
    nl, nr = len(tracks["id"]), len(tracks["id_right"])
    blk[0], blk[1] = nl, nr
    for k, name in enumerate(LEFT_FIELDS + RIGHT_FIELDS):
        n = nl if k < len(LEFT_FIELDS) else nr
        a = np.asarray(tracks[name])
        a = a.astype(np.int32) if name in INT_FIELDS else a.astype(np.float32).view(np.int32)
        blk[RESULT_HDR + k * max_cnt: RESULT_HDR + k * max_cnt + n] = a[:n]
    return blk


def unpack_result_block(blk, max_cnt: int) -> dict:
    """One packed block (int32[result_words]) -> dict of per-feature arrays."""
    blk = np.asarray(blk, np.int32)
    nl, nr = int(blk[0]), int(blk[1])
    if not (0 <= nl <= max_cnt and 0 <= nr <= max_cnt):
        raise ValueError(f"corrupt result block: n_left={nl} n_right={nr} max_cnt={max_cnt}")
    out = {}
    for k, name in enumerate(LEFT_FIELDS + RIGHT_FIELDS):
        n = nl if k < len(LEFT_FIELDS) else nr
        a = blk[RESULT_HDR + k * max_cnt: RESULT_HDR + k * max_cnt + n]
        out[name] = a.copy() if name in INT_FIELDS else a.view(np.float32).copy()
    out["stats"] = dict(n_prev=int(blk[2]), n_after_temporal=int(blk[3]),
                        n_after_ransac=int(blk[4]), n_after_mask=int(blk[5]),
                        n_new=int(blk[6]), ransac_iters=int(blk[8]), next_id=int(blk[9]))
    return out


def all_gather_tracks(local_block, gathered=None):
    """The per-window collective: every rank contributes its packed block (a torch int32
    tensor, on the GPU for NCCL / CPU for gloo) and receives all of them, [world, words]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if gathered is None:
        gathered = torch.empty((world, local_block.numel()), dtype=local_block.dtype,
                               device=local_block.device)
    if world == 1:
        gathered[0].copy_(local_block)
    else:
        dist.all_gather_into_tensor(gathered.view(-1), local_block.view(-1))
    return gathered


def aggregate_throughput(events_local: float, ms_local: float):
    """Whole-job Mevents/s: events summed over ranks / max over ranks of the timed region."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return events_local / (ms_local * 1e-3) / 1e6, ms_local
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    e = torch.tensor([events_local], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    return float(e.item()) / (float(t.item()) * 1e-3) / 1e6, float(t.item())


# ---------------------------------------------------------------------------------------
# left/right split of ONE stereo stream over two ranks
# ---------------------------------------------------------------------------------------
LEFT_RANK, RIGHT_RANK = 0, 1


class _DeviceBytes:
    """A raw device pointer seen as a uint8 vector through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 3}


def device_bytes(ptr: int, nbytes: int):
    """torch uint8 tensor aliasing `nbytes` of device memory at `ptr` (no copy)."""
    import torch
    return torch.as_tensor(_DeviceBytes(ptr, nbytes), device="cuda")


def exchange_right_image(rank: int, image, group=None):
    """The one data-path exchange of the split: the right rank sends its image block (all
    pyramid levels of the right time surface), the left rank receives it in place.  `image` is
    a uint8 tensor (device for NCCL, CPU for gloo); runs on the caller's current stream."""
    import torch.distributed as dist
    if rank == RIGHT_RANK:
        dist.send(image, dst=LEFT_RANK, group=group)
    elif rank == LEFT_RANK:
        dist.recv(image, src=RIGHT_RANK, group=group)
    else:
        raise ValueError("the left/right split uses ranks 0 (left) and 1 (right)")
    return image


class LeftRightSplit:
    """One stereo stream on two GPUs: rank 1 owns the right camera's SAE / time surface /
    pyramid, rank 0 everything else (the right camera has no detector and no temporal tracker
    in the reference, feature_tracker.cpp:475-575, so rank 1 is lightly loaded by construction).
    `fe` is this rank's EventFrontEnd; per window call `step(cur_time, events, pub)` with the
    left events on rank 0 and the right events on rank 1, then `wait()` on rank 0."""

    def __init__(self, fe, rank: int, exchange_stream=None, view=device_bytes):
        if rank not in (LEFT_RANK, RIGHT_RANK):
            raise ValueError("the left/right split uses ranks 0 (left) and 1 (right)")
        self.fe, self.rank = fe, rank
        if exchange_stream is None:
            import torch
            exchange_stream = torch.cuda.current_stream().cuda_stream
        self._stream = exchange_stream
        self._make_view = view        # (pointer, bytes) -> uint8 tensor; the tests use CPU tensors
        self._views = {}

    def _view(self, ptr, nbytes):
        v = self._views.get(ptr)
        if v is None:
            v = self._views[ptr] = self._make_view(ptr, nbytes)
        return v

    def step(self, cur_time, events, pub_this_frame=True):
        if self.rank == RIGHT_RANK:
            ptr, n = self.fe.split_image_submit(cur_time, events, self._stream)
            exchange_right_image(self.rank, self._view(ptr, n))
        else:
            ptr, n = self.fe.split_right_buffer()
            exchange_right_image(self.rank, self._view(ptr, n))
            self.fe.submit_split(cur_time, events, pub_this_frame, self._stream)

    def wait(self, unpack=True):
        if self.rank != LEFT_RANK:
            return None
        return self.fe.wait(unpack)


# ---------------------------------------------------------------------------------------
# time-window shard of ONE stereo stream (BASELINE configs[3])
# ---------------------------------------------------------------------------------------
class TimeShardRank:
    """One rank of the time-window shard: consecutive windows' SAE / time-surface / corner stages
    run on different GPUs, the serial track chain (feature_tracker.cpp:405-410,585-586) on rank 0.

    createSAE_*'s acceptance test reads only sae_latest_ ("time of the last event per pixel and
    polarity", event_detector.cc:149-166), and times only grow, so the state BEFORE window w is
    the element-wise maximum of the state at the round's start and of the per-window "last
    event" planes of the windows before w.  A round handles `world` consecutive windows, rank r
    window r of the round:

      phase_a   replay the window on a zeroed state        -> L_r = last event per pixel / polarity
      (all-gather of L)
      phase_c   carry-in = max(G_lat, L_0..L_{r-1}); sae := 0; replay again
                -> sae_latest after the window (exact) and S_r = last ACCEPTED time in the window
      (all-gather of S)
      phase_d   sae := max(G_sae, S_0..S_r): the exact state after window r; time surface,
                pyramids and Arc* candidates from it -> this window's products (one packed block)
      (all-gather of the products; rank 0 tracks the windows in order)
      G := max(G, L_*), max(G, S_*) on every rank.

    Every plane and image equals the sequential ones bit for bit (tests: a single-process run
    with `world` handles on one GPU against esvio_fe_track).  The collectives are the caller's
    (torch.distributed all_gather_into_tensor in TimeWindowShard.run_round, plain concatenation
    in the single-process test)."""

    def __init__(self, fe, rank: int, world: int, stream=None):
        import torch
        self.torch, self.fe, self.rank, self.world = torch, fe, rank, world
        self.stream = stream or torch.cuda.current_stream()
        self.s = self.stream.cuda_stream
        sae, lat, nbytes = fe.state_device_ptrs()
        self.nd = nbytes // 8
        f64 = lambda p: torch.as_tensor(_DeviceF64(p, self.nd), device="cuda")   # noqa: E731
        self.sae, self.lat = f64(sae), f64(lat)
        self.sae_ptr, self.lat_ptr = sae, lat
        dev = self.sae.device
        self.g_sae = torch.zeros(self.nd, dtype=torch.float64, device=dev)
        self.g_lat = torch.zeros(self.nd, dtype=torch.float64, device=dev)
        self.loc = torch.zeros(self.nd, dtype=torch.float64, device=dev)        # send buffer
        self.img_bytes, self.cand_bytes, self.cnt_bytes = fe.shard_sizes()
        self.prod_bytes = 2 * self.img_bytes + self.cand_bytes + self.cnt_bytes
        self.prod = torch.zeros(self.prod_bytes, dtype=torch.uint8, device=dev)
        self.window = None

    def _merge(self, dst_ptr, srcs):
        self.fe.shard_merge_max(dst_ptr, [t.data_ptr() for t in srcs], self.nd, self.s)

    def phase_a(self, t_ref, left, right):
        """left / right: device-resident events (frontend._Ev of DeviceEvents)."""
        torch = self.torch
        self.window = (t_ref, left, right)
        with torch.cuda.stream(self.stream):
            self.sae.zero_()
            self.lat.zero_()
        self.fe.shard_event_stage(t_ref, left, right, self.s)
        with torch.cuda.stream(self.stream):
            self.loc.copy_(self.lat)
        return self.loc

    def phase_c(self, gathered):
        """gathered: [world, nd] last-event planes of the round's windows."""
        torch = self.torch
        t_ref, left, right = self.window
        self.L = gathered
        with torch.cuda.stream(self.stream):
            self._merge(self.lat_ptr, [self.g_lat] + [gathered[k] for k in range(self.rank)])
            self.sae.zero_()
        self.fe.shard_event_stage(t_ref, left, right, self.s)
        with torch.cuda.stream(self.stream):
            self.loc.copy_(self.sae)
        return self.loc

    def phase_d(self, gathered, pub: bool, empty_left, empty_right):
        """gathered: [world, nd] accepted-time planes.  Returns this window's packed products
        (left image | right image | candidate lists | candidate counts) as a uint8 tensor."""
        torch = self.torch
        t_ref, left, right = self.window
        self.S = gathered
        with torch.cuda.stream(self.stream):
            self._merge(self.sae_ptr, [self.g_sae] + [gathered[k] for k in range(self.rank + 1)])
        # time surface + pyramids from the exact state: the event stage with no events
        self.fe.shard_event_stage(t_ref, empty_left, empty_right, self.s)
        li, ri = self.fe.shard_images()
        ib, cb, nb = self.img_bytes, self.cand_bytes, self.cnt_bytes
        with torch.cuda.stream(self.stream):
            self.prod[:ib].copy_(device_bytes(li, ib))
            self.prod[ib:2 * ib].copy_(device_bytes(ri, ib))
        if pub:
            cand, cnt = self.fe.shard_corner_candidates(left, self.s)
            with torch.cuda.stream(self.stream):
                self.prod[2 * ib:2 * ib + cb].copy_(device_bytes(cand, cb))
                self.prod[2 * ib + cb:].copy_(device_bytes(cnt, nb))
        # the state every rank starts the next round from
        with torch.cuda.stream(self.stream):
            self._merge(self.g_lat.data_ptr(), [self.g_lat] + [self.L[k] for k in range(self.world)])
            self._merge(self.g_sae.data_ptr(), [self.g_sae] + [self.S[k] for k in range(self.world)])
        return self.prod


class _DeviceF64:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 3}


class TimeShardTracker:
    """Rank 0's serial half: takes the gathered products of a round and tracks its windows in
    order through esvio_fe_track_submit_external."""

    def __init__(self, fe, shard_rank: TimeShardRank):
        self.fe, self.r = fe, shard_rank

    def submit(self, products, cur_time, n_left_events, pub):
        torch = self.r.torch
        li, ri, cand, cnt = self.fe.external_buffers()
        ib, cb, nb = self.r.img_bytes, self.r.cand_bytes, self.r.cnt_bytes
        with torch.cuda.stream(self.r.stream):
            device_bytes(li, ib).copy_(products[:ib])
            device_bytes(ri, ib).copy_(products[ib:2 * ib])
            if pub:
                device_bytes(cand, cb).copy_(products[2 * ib:2 * ib + cb])
                device_bytes(cnt, nb).copy_(products[2 * ib + cb:])
        self.fe.submit_external(cur_time, n_left_events, pub, self.r.s)
