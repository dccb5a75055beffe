"""ctypes binding of oracle/libsforacle.so (oracle/sf_oracle.c, the plain-C restatement of the
reference tick engine).  Test infrastructure only: imported by tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline legs, never by the product package.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from strikeforce_b200 import config as sfcfg  # noqa: E402  (struct layouts only)

LIB_PATH = os.path.join(HERE, "libsforacle%s.so" % sfcfg.GEOMETRY_TAG)
OBS_LEN = sfcfg.OBS_LEN
_lib = None


def build():
    if sfcfg.GEOMETRY_TAG:  # the same source for a larger arena (SF_GEOMETRY): dimensions are compile-time constants
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter",
                               "-I" + os.path.join(ROOT, "include")] + sfcfg.GEOMETRY_CFLAGS +
                              ["-shared", os.path.join(HERE, "sf_oracle.c"), "-o", LIB_PATH, "-lm"])
        return
    subprocess.check_call(["make", "-s", "-C", HERE, "libsforacle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.sfo_create.argtypes = [C.POINTER(sfcfg.SfConfig)]
        L.sfo_create.restype = C.c_void_p
        L.sfo_destroy.argtypes = [C.c_void_p]
        L.sfo_reset.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.sfo_step.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.sfo_step_a.argtypes = [C.c_void_p]
        L.sfo_step_b.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.sfo_status.argtypes = [C.c_void_p]
        L.sfo_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.sfo_population.argtypes = [C.c_void_p, C.c_void_p]
        L.sfo_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
        L.sfo_dump.restype = C.c_long
        L.sfo_hash.argtypes = [C.c_void_p]
        L.sfo_hash.restype = C.c_uint64
        L.sfo_step_out.argtypes = [C.c_void_p, C.POINTER(sfcfg.StepOut)]
        L.sfo_rng_draws.argtypes = [C.c_void_p]
        L.sfo_rng_draws.restype = C.c_int64
        L.sfo_observe.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfo_observe_raw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sfo_srand.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.sfo_rand.argtypes = [C.c_void_p]
        L.sfo_compute_damage.argtypes = [C.c_int, C.c_int]
        L.sfo_obs_transform.argtypes = [C.c_float]
        L.sfo_obs_transform.restype = C.c_float
        L.sfo_run_stream.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_char_p, C.c_int, C.c_long, C.c_int,
                                     C.POINTER(C.c_uint64)]
        L.sfo_run_stream.restype = C.c_long
        L.sfo_run_trace.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_char_p, C.c_int, C.c_long,
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p]
        L.sfo_run_trace.restype = C.c_long
        _lib = L
    return _lib


class Rng:
    """random.hpp:54-76 on a stand-alone generator."""

    def __init__(self, tb, serial):
        self._buf = (C.c_int64 * 55)()
        lib().sfo_srand(self._buf, tb, serial)

    def rand(self):
        return lib().sfo_rand(self._buf)

    def state(self):
        return list(self._buf[0:18]), int(self._buf[54])


def compute_damage(x, y):
    return lib().sfo_compute_damage(x, y)


def obs_transform(x):
    return float(lib().sfo_obs_transform(C.c_float(x)))


class Arena:
    """One arena of the C model; ``cfg`` is an ``SfConfig`` (strikeforce_b200.config.make_config)."""

    def __init__(self, cfg):
        self._cfg = cfg
        self._h = lib().sfo_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("sfo_create failed (caps too large for the model?)")

    def __del__(self):
        try:
            if getattr(self, "_h", None) and _lib is not None:
                _lib.sfo_destroy(self._h)
                self._h = None
        except Exception:
            pass  # interpreter shutdown

    def reset(self, level, tb, serial):
        lib().sfo_reset(self._h, level, tb, serial)

    def step(self, actions):
        return lib().sfo_step(self._h, bytes(actions), len(actions))

    def step_a(self):
        return lib().sfo_step_a(self._h)

    def step_b(self, actions):
        return lib().sfo_step_b(self._h, bytes(actions), len(actions))

    def status(self):
        return lib().sfo_status(self._h)

    def counters(self):
        out = np.zeros(8, dtype=np.int64)
        lib().sfo_counters(self._h, out.ctypes.data)
        return dict(zip(["frame", "kills", "teams_kills", "loot", "chest", "steps", "status", "hp"], out.tolist()))

    def population(self):
        out = np.zeros(6, dtype=np.int32)
        lib().sfo_population(self._h, out.ctypes.data)
        return dict(zip(["humans", "zombies", "bullets", "chests", "built", "portals"], out.tolist()))

    def dump(self):
        buf = np.empty(1 << 18, dtype=np.int32)
        n = lib().sfo_dump(self._h, buf.ctypes.data, buf.size)
        if n < 0:
            raise RuntimeError("sfo_dump: buffer too small")
        return buf[:n].copy()

    def state_hash(self):
        return int(lib().sfo_hash(self._h))

    def step_out(self):
        o = sfcfg.StepOut()
        lib().sfo_step_out(self._h, C.byref(o))
        return {n: getattr(o, n) for n, _ in sfcfg.StepOut._fields_}

    def rng_draws(self):
        return int(lib().sfo_rng_draws(self._h))

    def observe(self, slot=0, raw=False):
        out = np.empty(OBS_LEN, dtype=np.float32)
        fn = lib().sfo_observe_raw if raw else lib().sfo_observe
        if fn(self._h, slot, out.ctypes.data) != OBS_LEN:
            raise RuntimeError("slot %d has no active agent" % slot)
        return out

    def run_stream(self, env, level, n_steps, table, with_obs=False):
        h = C.c_uint64(0)
        n = lib().sfo_run_stream(self._h, env, level, bytes(table), len(table), n_steps, int(with_obs), C.byref(h))
        return n, int(h.value)

    def run_trace(self, env, level, n_steps, table, with_obs=False):
        """(episodes ended, checksum over the state hash after every step, last hash, observation of
        slot 0 after the last step or None): the synthetic workload of include/sf_synth.h with auto-reset."""
        chk, last = C.c_uint64(0), C.c_uint64(0)
        obs = np.empty(OBS_LEN, dtype=np.float32) if with_obs else None
        n = lib().sfo_run_trace(self._h, env, level, bytes(table), len(table), n_steps, C.byref(chk), C.byref(last),
                                obs.ctypes.data if with_obs else None)
        return n, int(chk.value), int(last.value), obs


def parse_record(rec):
    """Split a canonical record (include/sf_canon.h) into {(kind, index): fields}."""
    out, i = {}, 0
    rec = np.asarray(rec)
    while i + 3 <= len(rec):
        kind, index, nf = int(rec[i]), int(rec[i + 1]), int(rec[i + 2])
        out[(kind, index)] = rec[i + 3:i + 3 + nf].tolist()
        i += 3 + nf
    return out


def diff_records(a, b, limit=10):
    """Human-readable differences between two canonical records."""
    pa, pb = parse_record(a), parse_record(b)
    names = {1: "header", 2: "rng", 3: "human", 4: "zombie", 5: "bullet", 6: "portal", 7: "cell"}
    msgs = []
    for key in sorted(set(pa) | set(pb)):
        if pa.get(key) != pb.get(key):
            kind, index = key
            extra = ""
            if kind == 7:
                extra = " (f,r,c)=(%d,%d,%d)" % (index // (sfcfg.ROWS * sfcfg.COLS), index // sfcfg.COLS % sfcfg.ROWS, index % sfcfg.COLS)
            msgs.append("%s[%d]%s: %s != %s" % (names.get(kind, kind), index, extra, pa.get(key), pb.get(key)))
            if len(msgs) >= limit:
                break
    return msgs
