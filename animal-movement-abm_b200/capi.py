"""ctypes binding of the C ABI in include/abm.h.

The same binder is used for three libraries that export the ABI under different prefixes:
  * ``libabm.so``       prefix ``abm_``    — the product (CUDA, sm_100a); the only one this package loads itself;
  * ``liboracle.so``    prefix ``oracle_`` — the CPU restatement of the reference (tests / bench baseline only);
  * ``libabm_emu.so``   prefix ``emu_``    — the engine sources compiled for the host (tests only).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ABM_OK = 0
ABM_E_BADARG, ABM_E_CUDA, ABM_E_SMALL, ABM_E_STATE, ABM_E_NOWALKABLE = -1, -2, -3, -4, -5
ABM_E_CAPACITY, ABM_E_NOMEM, ABM_E_INTERNAL = -6, -7, -8

RANDOM_WALK, RANDOM_WALK_ATTRACTOR, RANDOM_WALK_GREGARIOUS, ALTERNATED_WALK, RANDOM_WALKMAP = range(5)
ASTAR_DIRECT, ASTAR_MAX = 0, 1
REDUCE_COUNT_AGENTS, REDUCE_COUNT_GROWN, REDUCE_SUM_ENERGY, REDUCE_MAX_ENERGY, REDUCE_MIN_ENERGY = range(5)


class AbmConfig(C.Structure):
    """struct abm_config (include/abm.h)."""

    _fields_ = [
        ("struct_size", C.c_int32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("periodic", C.c_int32),
        ("walk_type", C.c_int32),
        ("eat", C.c_int32),
        ("reproduce", C.c_int32),
        ("walkmap_regrowth", C.c_int32),
        ("regrowth_time", C.c_int32),
        ("visual_distance", C.c_int32),
        ("counter", C.c_int32),
        ("randomwalk_ifempty", C.c_int32),
        ("astar_metric", C.c_int32),
        ("astar_penalty", C.c_int32),
        ("device", C.c_int32),
        ("reserved0", C.c_int32),
        ("prob_random_walk", C.c_double),
        ("delta_energy", C.c_double),
        ("movement_cost", C.c_double),
        ("reproduction_prob", C.c_double),
        ("attractor_x", C.c_double),
        ("attractor_y", C.c_double),
        ("seed", C.c_uint64),
        ("strip_y0", C.c_int32),
        ("strip_ny", C.c_int32),
        ("n_ranks", C.c_int32),
        ("rank", C.c_int32),
        ("scheduler", C.c_int32),
        ("reserved1", C.c_int32),
    ]


SENDRECV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_int32)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32)


class AbmTransport(C.Structure):
    """struct abm_transport (include/abm.h)."""

    _fields_ = [("user", C.c_void_p), ("sendrecv", SENDRECV_FN), ("allreduce_sum_u64", ALLREDUCE_FN), ("allgather_u32", ALLGATHER_FN)]


class AbmFilter(C.Structure):
    """struct abm_filter (include/abm.h): a box of (energy, x, y), inclusive, 1-based."""

    _fields_ = [("energy_lo", C.c_double), ("energy_hi", C.c_double), ("x_lo", C.c_int32), ("x_hi", C.c_int32), ("y_lo", C.c_int32), ("y_hi", C.c_int32)]


class AbmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"abm error {code}: {msg}")
        self.code = code


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_U8P = C.POINTER(C.c_uint8)
_F64P = C.POINTER(C.c_double)
_U32P = C.POINTER(C.c_uint32)

# name -> (restype, argtypes); mirrors include/abm.h one to one
SIGNATURES = {
    "config_default": (None, [C.POINTER(AbmConfig)]),
    "config_size": (C.c_int, []),
    "create": (C.c_int, [C.POINTER(AbmConfig), C.POINTER(_P)]),
    "destroy": (None, [_P]),
    "set_grid": (C.c_int, [_P, _U8P, _I64P, _U8P, _I64P]),
    "set_agents": (C.c_int, [_P, C.c_int64, _I64P, _I64P, _I64P, _F64P, C.c_int64]),
    "init_random": (C.c_int, [_P, C.c_int64, C.c_uint64]),
    "set_global_walkmap_bits": (C.c_int, [_P, _U32P]),
    "set_global_state": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64]),
    "get_global_state": (C.c_int, [_P, _I64P, _I64P, _I64P, _I64P]),
    "set_weight_lut": (C.c_int, [_P, _F64P, C.c_int32]),
    "step": (C.c_int, [_P, C.c_int64]),
    "count_agents": (C.c_int, [_P, _I64P]),
    "get_agents": (C.c_int, [_P, _I64P, _I64P, _I64P, _I64P, _F64P]),
    "get_grid": (C.c_int, [_P, _U8P, _I64P]),
    "set_grass_packed": (C.c_int, [_P, _U8P]),
    "get_grass_packed": (C.c_int, [_P, _U8P]),
    "set_agents_packed": (C.c_int, [_P, C.c_int64, _U32P, _U32P, _F64P, C.c_int64]),
    "get_agents_packed": (C.c_int, [_P, _I64P, _U32P, _U32P, _F64P]),
    "reduce": (C.c_int, [_P, C.c_int32, _F64P]),
    "reduce_where": (C.c_int, [_P, C.c_int32, C.POINTER(AbmFilter), _F64P]),
    "get_heat": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_float)]),
    "plan_routes": (C.c_int, [_P, C.c_int64, _I64P, _I64P, _I64P]),
    "get_paths": (C.c_int, [_P, C.c_int64, _I64P, _I64P, _I64P, _I64P, _I64P, _I64P]),
    "get_timing": (C.c_int, [_P, _F64P, C.c_int32]),
    "set_profiling": (C.c_int, [_P, C.c_int32]),
    "get_kernel_stat": (C.c_int, [_P, C.c_int32, C.c_char_p, C.c_int32, _F64P, _I64P, _I64P]),
    "set_transport": (C.c_int, [_P, C.POINTER(AbmTransport)]),
    "nccl_unique_id": (C.c_int, [C.c_void_p]),
    "comm_init_nccl": (C.c_int, [_P, C.c_void_p]),
    "p2p_export": (C.c_int, [_P, C.c_void_p]),
    "p2p_connect": (C.c_int, [_P, C.c_void_p, C.c_int32]),
    "last_error": (C.c_char_p, [_P]),
    "group_create": (C.c_int, [C.POINTER(AbmConfig), C.c_int32, C.POINTER(C.c_int32), C.POINTER(_P)]),
    "group_destroy": (None, [_P]),
    "group_set_grid": (C.c_int, [_P, _U8P, _I64P, _U8P, _I64P]),
    "group_set_agents": (C.c_int, [_P, C.c_int64, _I64P, _I64P, _I64P, _F64P, C.c_int64]),
    "group_init_random": (C.c_int, [_P, C.c_int64, C.c_uint64]),
    "group_set_global_state": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int64]),
    "group_get_global_state": (C.c_int, [_P, _I64P, _I64P, _I64P, _I64P]),
    "group_step": (C.c_int, [_P, C.c_int64]),
    "group_count_agents": (C.c_int, [_P, _I64P]),
    "group_get_agents": (C.c_int, [_P, _I64P, _I64P, _I64P, _I64P, _F64P]),
    "group_get_grid": (C.c_int, [_P, _U8P, _I64P]),
    "group_reduce": (C.c_int, [_P, C.c_int32, _F64P]),
    "group_member": (_P, [_P, C.c_int32]),
    "group_last_error": (C.c_char_p, [_P]),
}
# entry points the CPU oracle does not have (it is one process on one device, like the reference)
ENGINE_ONLY = {n for n in SIGNATURES if n.startswith("group_")}


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(ty)


class CLib:
    """A loaded shared library exporting the ABI under ``prefix``."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        self.fn = {}
        for name, (res, args) in SIGNATURES.items():
            if name in ENGINE_ONLY and not hasattr(self.dll, prefix + name):
                continue
            f = getattr(self.dll, prefix + name)
            f.restype, f.argtypes = res, args
            self.fn[name] = f
        if self.fn["config_size"]() != C.sizeof(AbmConfig):
            raise RuntimeError(f"{path}: abm_config is {self.fn['config_size']()} bytes, the ctypes mirror {C.sizeof(AbmConfig)}")

    def default_config(self, **kw) -> AbmConfig:
        cfg = AbmConfig()
        self.fn["config_default"](C.byref(cfg))
        for k, v in kw.items():
            if not hasattr(cfg, k):
                raise AttributeError(f"abm_config has no field {k!r}")
            setattr(cfg, k, v)
        return cfg

    def create(self, cfg: AbmConfig) -> "Engine":
        return Engine(self, cfg)

    def create_group(self, cfg: AbmConfig, device_ids) -> "Group":
        return Group(self, cfg, device_ids)


class Group:
    """One abm_group: n strip handles on n devices of one process, whole-GridSpace arrays in and out (include/abm.h)."""

    def __init__(self, lib: "CLib", cfg: AbmConfig, device_ids):
        self.lib, self.cfg = lib, cfg
        ids = (C.c_int32 * len(device_ids))(*device_ids)
        self.g = _P()
        rc = lib.fn["group_create"](C.byref(cfg), len(device_ids), ids, C.byref(self.g))
        if rc != ABM_OK:
            msg = lib.fn["group_last_error"](None)
            self.g = None
            raise AbmError(rc, (msg or b"").decode())
        self.nx, self.ny, self.n_devices = cfg.nx, cfg.ny, len(device_ids)

    def _ck(self, rc):
        if rc != ABM_OK:
            raise AbmError(rc, (self.lib.fn["group_last_error"](self.g) or b"").decode())

    def close(self):
        if self.g is not None and self.g.value is not None:
            self.lib.fn["group_destroy"](self.g)
        self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_grid(self, fully_grown, countdown, walkmap=None, elevation=None):
        c = lambda a, dt: None if a is None else np.ascontiguousarray(a, dtype=dt).reshape(-1)
        fg, cd, wm, el = c(fully_grown, np.uint8), c(countdown, np.int64), c(walkmap, np.uint8), c(elevation, np.int64)
        self._ck(self.lib.fn["group_set_grid"](self.g, _ptr(fg, _U8P), _ptr(cd, _I64P), _ptr(wm, _U8P), _ptr(el, _I64P)))

    def set_agents(self, ids, x, y, energy, next_id=None):
        ids = np.ascontiguousarray(ids, dtype=np.int64); x = np.ascontiguousarray(x, dtype=np.int64)
        y = np.ascontiguousarray(y, dtype=np.int64); e = np.ascontiguousarray(energy, dtype=np.float64)
        if next_id is None:
            next_id = int(ids.max()) + 1 if ids.size else 1
        self._ck(self.lib.fn["group_set_agents"](self.g, ids.size, _ptr(ids, _I64P), _ptr(x, _I64P), _ptr(y, _I64P), _ptr(e, _F64P), int(next_id)))

    def init_random(self, n_sheep, init_seed):
        self._ck(self.lib.fn["group_init_random"](self.g, int(n_sheep), int(init_seed)))

    def set_global_state(self, tick=0, behav=0, behav_counter=0):
        self._ck(self.lib.fn["group_set_global_state"](self.g, int(tick), int(behav), int(behav_counter)))

    def get_global_state(self):
        v = [C.c_int64() for _ in range(4)]
        self._ck(self.lib.fn["group_get_global_state"](self.g, *[C.byref(a) for a in v]))
        return dict(tick=v[0].value, behav=v[1].value, behav_counter=v[2].value, next_id=v[3].value)

    def step(self, n=1):
        self._ck(self.lib.fn["group_step"](self.g, int(n)))

    def count_agents(self):
        n = C.c_int64()
        self._ck(self.lib.fn["group_count_agents"](self.g, C.byref(n)))
        return n.value

    def get_agents(self):
        n = self.count_agents()
        ids = np.empty(n, np.int64); x = np.empty(n, np.int64); y = np.empty(n, np.int64); e = np.empty(n, np.float64)
        cap = C.c_int64(n)
        self._ck(self.lib.fn["group_get_agents"](self.g, C.byref(cap), _ptr(ids, _I64P), _ptr(x, _I64P), _ptr(y, _I64P), _ptr(e, _F64P)))
        return ids, x, y, e

    def get_grid(self):
        fg = np.empty((self.ny, self.nx), np.uint8); cd = np.empty((self.ny, self.nx), np.int64)
        self._ck(self.lib.fn["group_get_grid"](self.g, _ptr(fg, _U8P), _ptr(cd, _I64P)))
        return fg, cd

    def reduce(self, what):
        out = C.c_double()
        self._ck(self.lib.fn["group_reduce"](self.g, int(what), C.byref(out)))
        return out.value

    def snapshot(self):
        ids, x, y, e = self.get_agents()
        fg, cd = self.get_grid()
        return dict(ids=ids, x=x, y=y, energy=e, fully_grown=fg, countdown=cd, **self.get_global_state())


class Engine:
    """One abm_handle.  Method names follow the ABI; array arguments are numpy arrays."""

    def __init__(self, lib: CLib, cfg: AbmConfig):
        self.lib, self.cfg = lib, cfg
        self.h = _P()
        rc = lib.fn["create"](C.byref(cfg), C.byref(self.h))
        if rc != ABM_OK:
            msg = lib.fn["last_error"](None)
            self.h = None
            raise AbmError(rc, (msg or b"").decode())
        self.nx, self.ny = cfg.nx, cfg.ny

    def _ck(self, rc: int):
        if rc != ABM_OK:
            raise AbmError(rc, (self.lib.fn["last_error"](self.h) or b"").decode())

    def close(self):
        if self.h is not None and self.h.value is not None:
            self.lib.fn["destroy"](self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # grids are (ny, nx) C-ordered numpy arrays == Julia (nx, ny) column-major: a[y, x]
    def set_grid(self, fully_grown, countdown, walkmap=None, elevation=None):
        G = self.nx * self.ny
        fg = None if fully_grown is None else np.ascontiguousarray(fully_grown, dtype=np.uint8).reshape(-1)
        cd = None if countdown is None else np.ascontiguousarray(countdown, dtype=np.int64).reshape(-1)
        wm = None if walkmap is None else np.ascontiguousarray(walkmap, dtype=np.uint8).reshape(-1)
        el = None if elevation is None else np.ascontiguousarray(elevation, dtype=np.int64).reshape(-1)
        for a in (fg, cd, wm, el):
            if a is not None and a.size != G:
                raise ValueError("grid array has the wrong size")
        self._ck(self.lib.fn["set_grid"](self.h, _ptr(fg, _U8P), _ptr(cd, _I64P), _ptr(wm, _U8P), _ptr(el, _I64P)))

    def set_agents(self, ids, x, y, energy, next_id=None):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        x = np.ascontiguousarray(x, dtype=np.int64)
        y = np.ascontiguousarray(y, dtype=np.int64)
        e = np.ascontiguousarray(energy, dtype=np.float64)
        n = ids.size
        if not (x.size == y.size == e.size == n):
            raise ValueError("agent arrays differ in length")
        if next_id is None:
            next_id = int(ids.max()) + 1 if n else 1
        self._ck(self.lib.fn["set_agents"](self.h, n, _ptr(ids, _I64P), _ptr(x, _I64P), _ptr(y, _I64P), _ptr(e, _F64P), int(next_id)))

    def init_random(self, n_sheep, init_seed):
        """The random part of `initialize_model`, built by the engine itself (abm_init_random)."""
        self._ck(self.lib.fn["init_random"](self.h, int(n_sheep), int(init_seed)))

    def set_global_walkmap(self, walkmap):
        """Strip handles: the walkmap of the whole GridSpace ((ny, nx) array of 0/1), packed to one bit per cell."""
        if walkmap is None:
            self._ck(self.lib.fn["set_global_walkmap_bits"](self.h, None))
            return
        bits = np.packbits(np.ascontiguousarray(walkmap, dtype=np.uint8).reshape(-1) != 0, bitorder="little")
        bits = np.concatenate([bits, np.zeros((-bits.size) % 4, np.uint8)]).view(np.uint32)
        self._ck(self.lib.fn["set_global_walkmap_bits"](self.h, _ptr(bits, _U32P)))

    def set_global_walkmap_bits(self, bits):
        """The same from bits the caller packed already (uint32 words, bit l & 31 of word l >> 5, l = (y-1)*nx + (x-1))."""
        bits = np.ascontiguousarray(bits, dtype=np.uint32).reshape(-1)
        if bits.size * 32 < self.cfg.nx * self.cfg.ny:
            raise ValueError("global walkmap bits are too short")
        self._ck(self.lib.fn["set_global_walkmap_bits"](self.h, _ptr(bits, _U32P)))

    def set_global_state(self, tick=0, behav=0, behav_counter=0):
        self._ck(self.lib.fn["set_global_state"](self.h, int(tick), int(behav), int(behav_counter)))

    def get_global_state(self):
        v = [C.c_int64() for _ in range(4)]
        self._ck(self.lib.fn["get_global_state"](self.h, *[C.byref(a) for a in v]))
        return dict(tick=v[0].value, behav=v[1].value, behav_counter=v[2].value, next_id=v[3].value)

    def set_weight_lut(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self._ck(self.lib.fn["set_weight_lut"](self.h, _ptr(w, _F64P), w.size))

    def step(self, n=1):
        self._ck(self.lib.fn["step"](self.h, int(n)))

    def count_agents(self) -> int:
        n = C.c_int64()
        self._ck(self.lib.fn["count_agents"](self.h, C.byref(n)))
        return n.value

    def get_agents(self, out=None):
        """ids, x, y, energy in id order.  `out`: optional (ids, x, y, energy) buffers (e.g. pinned) of sufficient length."""
        n = self.count_agents()
        if out is None:
            ids = np.empty(n, np.int64); x = np.empty(n, np.int64); y = np.empty(n, np.int64); e = np.empty(n, np.float64)
        else:
            ids, x, y, e = out
            if min(ids.size, x.size, y.size, e.size) < n:
                raise AbmError(ABM_E_SMALL, f"agent buffers hold fewer than {n} entries")
        cap = C.c_int64(ids.size)
        self._ck(self.lib.fn["get_agents"](self.h, C.byref(cap), _ptr(ids, _I64P), _ptr(x, _I64P), _ptr(y, _I64P), _ptr(e, _F64P)))
        return ids[:n], x[:n], y[:n], e[:n]

    def get_grid(self, out=None):
        if out is None:
            fg = np.empty((self.ny, self.nx), np.uint8)
            cd = np.empty((self.ny, self.nx), np.int64)
        else:
            fg, cd = out
        self._ck(self.lib.fn["get_grid"](self.h, _ptr(fg, _U8P), _ptr(cd, _I64P)))
        return fg, cd

    # narrow transfer forms: one grass byte per cell (fully_grown << 7 | countdown), agents as (id u32, x | y << 16, energy)
    def set_grass_packed(self, state):
        st = np.ascontiguousarray(state, dtype=np.uint8).reshape(-1)
        if st.size != self.nx * self.ny:
            raise ValueError("grid array has the wrong size")
        self._ck(self.lib.fn["set_grass_packed"](self.h, _ptr(st, _U8P)))

    def get_grass_packed(self, out=None):
        st = np.empty((self.ny, self.nx), np.uint8) if out is None else out
        self._ck(self.lib.fn["get_grass_packed"](self.h, _ptr(st, _U8P)))
        return st

    def set_agents_packed(self, ids, xy, energy, next_id=None):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        xy = np.ascontiguousarray(xy, dtype=np.uint32)
        e = np.ascontiguousarray(energy, dtype=np.float64)
        n = ids.size
        if not (xy.size == e.size == n):
            raise ValueError("agent arrays differ in length")
        if next_id is None:
            next_id = int(ids.max()) + 1 if n else 1
        self._ck(self.lib.fn["set_agents_packed"](self.h, n, _ptr(ids, _U32P), _ptr(xy, _U32P), _ptr(e, _F64P), int(next_id)))

    def get_agents_packed(self, out=None):
        """ids (u32), x | y << 16 (u32), energy in id order; `out`: optional (ids, xy, energy) buffers."""
        n = self.count_agents()
        if out is None:
            ids = np.empty(n, np.uint32); xy = np.empty(n, np.uint32); e = np.empty(n, np.float64)
        else:
            ids, xy, e = out
            if min(ids.size, xy.size, e.size) < n:
                raise AbmError(ABM_E_SMALL, f"agent buffers hold fewer than {n} entries")
        cap = C.c_int64(ids.size)
        self._ck(self.lib.fn["get_agents_packed"](self.h, C.byref(cap), _ptr(ids, _U32P), _ptr(xy, _U32P), _ptr(e, _F64P)))
        return ids[:n], xy[:n], e[:n]

    def reduce(self, what: int) -> float:
        out = C.c_double()
        self._ck(self.lib.fn["reduce"](self.h, int(what), C.byref(out)))
        return out.value

    def reduce_where(self, what: int, energy=(-np.inf, np.inf), x=None, y=None) -> float:
        """abm_reduce_where: `what` over the agents with energy, x, y inside the given inclusive ranges."""
        x = x or (1, self.cfg.nx)
        y = y or (1, self.cfg.ny)
        f = AbmFilter(float(energy[0]), float(energy[1]), int(x[0]), int(x[1]), int(y[0]), int(y[1]))
        out = C.c_double()
        self._ck(self.lib.fn["reduce_where"](self.h, int(what), C.byref(f), C.byref(out)))
        return out.value

    def get_heat(self, which=0):
        """0: countdown ./ regrowth_time (grasscolor); 1: the penaltymap (elevation) — float32 (ny, nx), computed on the device."""
        out = np.empty((self.ny, self.nx), np.float32)
        self._ck(self.lib.fn["get_heat"](self.h, int(which), out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def plan_routes(self, agent_ids, goal_x, goal_y):
        a = np.ascontiguousarray(agent_ids, dtype=np.int64)
        gx = np.ascontiguousarray(goal_x, dtype=np.int64)
        gy = np.ascontiguousarray(goal_y, dtype=np.int64)
        self._ck(self.lib.fn["plan_routes"](self.h, a.size, _ptr(a, _I64P), _ptr(gx, _I64P), _ptr(gy, _I64P)))

    def get_paths(self, agent_ids):
        a = np.ascontiguousarray(agent_ids, dtype=np.int64)
        b = a.size
        off = np.zeros(b + 1, np.int64)
        cost = np.zeros(b, np.int64)
        cap = C.c_int64(0)
        px = np.empty(0, np.int64); py = np.empty(0, np.int64)
        rc = self.lib.fn["get_paths"](self.h, b, _ptr(a, _I64P), _ptr(off, _I64P), C.byref(cap), _ptr(px, _I64P), _ptr(py, _I64P), _ptr(cost, _I64P))
        if rc == ABM_E_SMALL:
            px = np.empty(cap.value, np.int64); py = np.empty(cap.value, np.int64)
            rc = self.lib.fn["get_paths"](self.h, b, _ptr(a, _I64P), _ptr(off, _I64P), C.byref(cap), _ptr(px, _I64P), _ptr(py, _I64P), _ptr(cost, _I64P))
        self._ck(rc)
        return off, px[: cap.value], py[: cap.value], cost

    def timing(self):
        out = np.zeros(8, np.float64)
        self._ck(self.lib.fn["get_timing"](self.h, _ptr(out, _F64P), out.size))
        return dict(ms_total=out[0], ms_agents=out[1], ms_grass=out[2], rounds=out[3], launches=out[4], astar_expansions=out[5],
                    agent_updates=out[6])

    def set_transport(self, transport: "AbmTransport"):
        self._transport = transport  # keep the callbacks alive
        self._ck(self.lib.fn["set_transport"](self.h, C.byref(transport)))

    def comm_init_nccl(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.fn["comm_init_nccl"](self.h, buf))

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.fn["p2p_export"](self.h, buf))
        return buf.raw

    def p2p_connect(self, handles: bytes, n_ranks: int):
        buf = C.create_string_buffer(handles, 64 * n_ranks)
        self._ck(self.lib.fn["p2p_connect"](self.h, buf, n_ranks))

    def set_profiling(self, on=True):
        self._ck(self.lib.fn["set_profiling"](self.h, int(bool(on))))

    def kernel_stats(self):
        """Per-kernel device time of the last step(): {name: dict(ms, launches, units)}."""
        out, i = {}, 0
        while True:
            name = C.create_string_buffer(64)
            ms, ln, un = C.c_double(), C.c_int64(), C.c_int64()
            rc = self.lib.fn["get_kernel_stat"](self.h, i, name, 64, C.byref(ms), C.byref(ln), C.byref(un))
            if rc == ABM_E_SMALL:
                return out
            self._ck(rc)
            out[name.value.decode()] = dict(ms=ms.value, launches=ln.value, units=un.value)
            i += 1

    def snapshot(self):
        """Full observable state, for parity comparisons."""
        ids, x, y, e = self.get_agents()
        fg, cd = self.get_grid()
        return dict(ids=ids, x=x, y=y, energy=e, fully_grown=fg, countdown=cd, **self.get_global_state())
