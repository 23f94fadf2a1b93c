"""Row-strip decomposition helpers for the multi-GPU path (one process per GPU, one handle per process).

The reference is a single process (Agents.jl `step!` is serial); this is new: strip s owns the rows
y in [y0, y0 + ny_local) of the GridSpace and all agents standing there.  `abm_step` itself exchanges the boundary
sub-tile rows with the ring neighbours every tick (include/abm.h, "multi-GPU").  torch.distributed is only plumbing
here: it ships NCCL's unique id, or — for hosts without NCCL — carries the exchange itself through the
`abm_transport` callbacks (host memory; used by the CPU tests over gloo).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

GHOST = 32  # ghost rows above and below the strip in every grid array (one tile row)


def strip_bounds(ny: int, n_ranks: int, rank: int):
    """Rows [y0, y0 + n) of `rank`: whole 32-row cores, as even as possible."""
    cores = ny // 32
    if ny % 32 or cores < n_ranks:
        raise ValueError("ny must be a multiple of 32 with at least one core row per rank")
    lo = (cores * rank) // n_ranks
    hi = (cores * (rank + 1)) // n_ranks
    return 32 * lo, 32 * (hi - lo)


def strip_rows(a, y0: int, n: int, periodic: bool, fill=0):
    """Rows y0-32 .. y0+n+32 of a global (ny, nx) array; rows outside a non-periodic grid are `fill`."""
    a = np.asarray(a)
    ny = a.shape[0]
    rows = np.arange(y0 - GHOST, y0 + n + GHOST)
    if periodic:
        return np.ascontiguousarray(a[rows % ny])
    out = np.full((rows.size,) + a.shape[1:], fill, a.dtype)
    ok = (rows >= 0) & (rows < ny)
    out[ok] = a[rows[ok]]
    return out


def load_strip(lib: capi.CLib, cfgkw: dict, world: dict, n_ranks: int, rank: int, next_id=None):
    """Creates the handle of strip `rank` of a global world (dict of abm_b200.worlds arrays)."""
    ny, periodic = cfgkw["ny"], bool(cfgkw.get("periodic", 0))
    y0, n = strip_bounds(ny, n_ranks, rank)
    e = lib.create(lib.default_config(strip_y0=y0, strip_ny=n, n_ranks=n_ranks, rank=rank, **cfgkw))
    e.ny = n + 2 * GHOST
    wm, el = world.get("walkmap"), world.get("elevation")
    e.set_grid(strip_rows(world["fully_grown"], y0, n, periodic), strip_rows(world["countdown"], y0, n, periodic),
               None if wm is None else strip_rows(wm, y0, n, periodic), None if el is None else strip_rows(el, y0, n, periodic))
    mine = (world["y"] > y0) & (world["y"] <= y0 + n)
    if next_id is None:
        next_id = int(world["ids"].max()) + 1 if world["ids"].size else 1
    e.set_agents(world["ids"][mine], world["x"][mine], world["y"][mine], world["energy"][mine], next_id)
    e.strip_y0, e.strip_ny = y0, n
    return e


def strip_snapshot(e):
    """Own agents and own grid rows of a strip handle."""
    ids, x, y, en = e.get_agents()
    fg, cd = e.get_grid()
    st = e.get_global_state()
    return dict(ids=ids, x=x, y=y, energy=en, fully_grown=fg[GHOST:GHOST + e.strip_ny], countdown=cd[GHOST:GHOST + e.strip_ny], **st)


def merge_snapshots(parts):
    """Global state from the strips' snapshots (in rank order)."""
    ids = np.concatenate([p["ids"] for p in parts])
    o = np.argsort(ids, kind="stable")
    out = dict(ids=ids[o], x=np.concatenate([p["x"] for p in parts])[o], y=np.concatenate([p["y"] for p in parts])[o],
               energy=np.concatenate([p["energy"] for p in parts])[o],
               fully_grown=np.concatenate([p["fully_grown"] for p in parts], 0), countdown=np.concatenate([p["countdown"] for p in parts], 0))
    for k in ("tick", "behav", "behav_counter", "next_id"):
        out[k] = parts[0][k]
    return out


class DistTransport:
    """`abm_transport` over a torch.distributed process group with CPU tensors (gloo)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._cbs = (capi.SENDRECV_FN(self._sendrecv), capi.ALLREDUCE_FN(self._allreduce), capi.ALLGATHER_FN(self._allgather))
        self.struct = capi.AbmTransport(None, *self._cbs)

    def _view(self, ptr, nbytes):
        return self.torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(int(nbytes),)))

    def _sendrecv(self, user, to_prev, b_tp, to_next, b_tn, from_prev, b_fp, from_next, b_fn):
        try:
            dist, prev, nxt = self.dist, (self.rank - 1) % self.world, (self.rank + 1) % self.world
            ops = []
            # issue order matters when both neighbours are the same rank (n_ranks == 2): see include/abm.h
            if b_tp:
                ops.append(dist.P2POp(dist.isend, self._view(to_prev, b_tp), prev, self.group, tag=1))
            if b_tn:
                ops.append(dist.P2POp(dist.isend, self._view(to_next, b_tn), nxt, self.group, tag=2))
            if b_fn:
                ops.append(dist.P2POp(dist.irecv, self._view(from_next, b_fn), nxt, self.group, tag=1))
            if b_fp:
                ops.append(dist.P2POp(dist.irecv, self._view(from_prev, b_fp), prev, self.group, tag=2))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
            return 0
        except Exception as ex:  # never let an exception cross the C boundary
            print("DistTransport.sendrecv failed:", ex)
            return 1

    def _allreduce(self, user, values, n):
        try:
            t = self.torch.from_numpy(np.ctypeslib.as_array(values, shape=(int(n),)).view(np.int64))
            self.dist.all_reduce(t, group=self.group)
            return 0
        except Exception as ex:
            print("DistTransport.allreduce failed:", ex)
            return 1

    def _allgather(self, user, send, recv, n):
        try:
            s = self.torch.from_numpy(np.ctypeslib.as_array(send, shape=(int(n),)).view(np.int32))
            r = self.torch.from_numpy(np.ctypeslib.as_array(recv, shape=(int(n) * self.world,)).view(np.int32))
            self.dist.all_gather_into_tensor(r, s, group=self.group) if hasattr(self.dist, "all_gather_into_tensor") and self.dist.get_backend(self.group) != "gloo" \
                else self._gather_list(r, s, int(n))
            return 0
        except Exception as ex:
            print("DistTransport.allgather failed:", ex)
            return 1

    def _gather_list(self, r, s, n):
        parts = [self.torch.empty(n, dtype=self.torch.int32) for _ in range(self.world)]
        self.dist.all_gather(parts, s, group=self.group)
        for i, p in enumerate(parts):
            r[i * n:(i + 1) * n] = p


def init_nccl(engine, device=None):
    """NCCL ring for a strip handle: rank 0's unique id is broadcast through the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    buf = C.create_string_buffer(128)
    if dist.get_rank() == 0:
        rc = engine.lib.fn["nccl_unique_id"](buf)
        if rc != 0:
            raise capi.AbmError(rc, "abm_nccl_unique_id failed")
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
    dist.broadcast(t, 0)
    engine.comm_init_nccl(bytes(t.cpu().numpy().tobytes()))


def init_p2p(engine):
    """Peer-memory transport (NVLink): every rank exports the CUDA IPC handle of its mailbox, the handles are
    all-gathered through the default torch.distributed group, every rank maps its peers' mailboxes."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    mine = torch.frombuffer(bytearray(engine.p2p_export()), dtype=torch.uint8)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    parts = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine.to(dev))
    failure = None
    try:
        engine.p2p_connect(b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts), world)
    except Exception as exc:  # noqa: BLE001 -- keep the collective sequence identical on every rank, then report
        failure = exc
    dist.barrier()
    if failure is not None:
        raise failure
