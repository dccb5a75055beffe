"""ctypes binding of tests/hostcheck/libhostcheck.so: the device tick (csrc/sf_core.cuh)
compiled for the host.  Debugging aid for the CPU test-suite only (see hostcheck.cpp)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from strikeforce_b200 import config as sfcfg  # noqa: E402

# extra compiler flags build a variant next to the default one (e.g. "-DSF_BT_SLOTS=7 -DSF_BT_FULL=5": a
# bullet-flag table so small that most flags spill into the overlay); HOSTCHECK_CFLAGS sets the default variant
EXTRA = tuple(os.environ.get("HOSTCHECK_CFLAGS", "").split()) + tuple(sfcfg.GEOMETRY_CFLAGS)
_libs = {}


def lib_path(extra=EXTRA):
    tag = "_" + "".join(c for c in "".join(extra) if c.isalnum()) if extra else ""
    return os.path.join(HERE, "libhostcheck%s.so" % tag)


def build(force=False, extra=EXTRA):
    srcs = [os.path.join(HERE, "hostcheck.cpp")] + [
        os.path.join(ROOT, "strikeforce_b200", "csrc", f)
        for f in ("sf_core.cuh", "sf_canon_dev.cuh", "sf_host_setup.h", "sf_state.h", "sf_obs.cuh")]
    path = lib_path(extra)
    if not force and os.path.exists(path) and all(os.path.getmtime(path) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call([
        "g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
        "-I" + os.path.join(ROOT, "strikeforce_b200", "csrc")] + list(extra) + ["-x", "c++", srcs[0], "-o", path])


def lib(extra=EXTRA):
    extra = tuple(extra)
    if extra not in _libs:
        build(extra=extra)
        L = C.CDLL(lib_path(extra))
        L.hc_create.argtypes = [C.POINTER(sfcfg.SfConfig)]
        L.hc_create.restype = C.c_void_p
        L.hc_destroy.argtypes = [C.c_void_p]
        L.hc_reset.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong]
        L.hc_step.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.hc_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long]
        L.hc_dump.restype = C.c_long
        L.hc_hash.argtypes = [C.c_void_p, C.c_int]
        L.hc_hash.restype = C.c_ulonglong
        L.hc_step_out.argtypes = [C.c_void_p, C.c_int, C.POINTER(sfcfg.StepOut)]
        L.hc_status.argtypes = [C.c_void_p, C.c_int]
        L.hc_misc.argtypes = [C.c_void_p, C.c_int]
        L.hc_misc.restype = C.c_uint
        L.hc_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.hc_observe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.hc_n_agents.argtypes = [C.c_void_p]
        L.hc_compute_damage.argtypes = [C.c_int, C.c_int]
        L.hc_obs_transform_milli.argtypes = [C.c_int]
        L.hc_obs_transform_milli.restype = C.c_float
        _libs[extra] = L
    return _libs[extra]


class HostSim:
    def __init__(self, cfg, extra=EXTRA):
        self._cfg = cfg
        self._lib = lib(extra)
        self._h = self._lib.hc_create(C.byref(cfg))
        if not self._h:
            raise RuntimeError("hc_create failed")
        self.n_envs = cfg.n_envs
        self.n_agents = self._lib.hc_n_agents(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.hc_destroy(self._h)
            self._h = None

    def reset(self, env, tb, serial):
        self._lib.hc_reset(self._h, env, tb, serial)

    def step(self, actions, half=0):
        a = None if actions is None else bytes(actions)
        if a is not None:
            assert len(a) == self.n_envs * self.n_agents
        self._lib.hc_step(self._h, a, half)

    def dump(self, env):
        buf = np.empty(1 << 18, dtype=np.int32)
        n = self._lib.hc_dump(self._h, env, buf.ctypes.data, buf.size)
        assert n >= 0
        return buf[:n].copy()

    def state_hash(self, env):
        return int(self._lib.hc_hash(self._h, env))

    def status(self, env):
        return self._lib.hc_status(self._h, env)

    def misc(self, env):
        """the packed header word (sf_state.h SfDev::misc)"""
        return self._lib.hc_misc(self._h, env)

    def step_out(self, env):
        o = sfcfg.StepOut()
        self._lib.hc_step_out(self._h, env, C.byref(o))
        return {n: getattr(o, n) for n, _ in sfcfg.StepOut._fields_}

    def observe(self, env, slot=0, raw=False):
        out = np.empty(sfcfg.OBS_LEN, dtype=np.float32)
        if self._lib.hc_observe(self._h, env, slot, out.ctypes.data, int(raw)) != sfcfg.OBS_LEN:
            raise RuntimeError("no such human slot")
        return out

    def stats(self):
        out = np.zeros(16, dtype=np.uint64)
        self._lib.hc_stats(self._h, out.ctypes.data)
        return dict(zip(sfcfg.STAT_NAMES, out.astype(np.int64).tolist()))
