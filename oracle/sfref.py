"""ctypes binding of oracle/_ref/libsfref.so -- the UNMODIFIED reference tick engine driven
headless (oracle/ref_harness/harness.cpp).  Test infrastructure only: imported by tests/,
__graft_entry__.smoke() and bench.py's reference arm, never by the product package.

One process == one arena (the reference keeps its state in globals, gameplay.hpp:37-55).
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFREF_LIB") or os.path.join(HERE, "_ref", "libsfref.so")
NOGUARD_PATH = os.path.join(HERE, "_ref", "libsfref_noguard.so")  # timing build: the harness's own checks compiled out
RUNDIR = os.path.join(HERE, "_ref", "rundir")
ACCOUNT1 = os.path.join(RUNDIR, "player_account1.txt")
OBS_LEN = 32 * 31 * 31

_lib = None


def available():
    return os.path.exists(LIB_PATH) and os.path.isdir(RUNDIR)


def lib():
    """Load the library and chdir-initialise it.  NOTE: the reference opens its data files
    by relative path, so the harness chdir()s into oracle/_ref/rundir for good."""
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH)
        L.sfref_init.argtypes = [ctypes.c_char_p]
        L.sfref_reset.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong,
                                  ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long]
        L.sfref_reset_ex.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_long, ctypes.c_int, ctypes.c_char_p]
        L.sfref_seeds.argtypes = [ctypes.c_void_p]
        L.sfref_close_log.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.sfref_step.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.sfref_observe.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.sfref_set_capture.argtypes = [ctypes.c_int, ctypes.c_int]
        L.sfref_get_captured.argtypes = [ctypes.c_int, ctypes.c_void_p]
        L.sfref_counters.argtypes = [ctypes.c_void_p]
        L.sfref_population.argtypes = [ctypes.c_void_p]
        L.sfref_dump.argtypes = [ctypes.c_void_p, ctypes.c_long]
        L.sfref_dump.restype = ctypes.c_long
        L.sfref_hash.restype = ctypes.c_ulonglong
        L.sfref_srand.argtypes = [ctypes.c_longlong, ctypes.c_longlong]
        L.sfref_rng_state.argtypes = [ctypes.c_void_p]
        L.sfref_compute_damage.argtypes = [ctypes.c_int, ctypes.c_int]
        L.sfref_run_stream.argtypes = [ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_long, ctypes.c_char_p,
                                       ctypes.c_int, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]
        L.sfref_run_stream.restype = ctypes.c_long
        cwd = os.getcwd()
        if L.sfref_init(RUNDIR.encode()) != 0:
            raise RuntimeError("sfref_init failed for " + RUNDIR)
        os.chdir(cwd)  # python side keeps its cwd; the harness re-enters rundir per call below
        _lib = L
    return _lib


class _InRundir:
    def __enter__(self):
        self.cwd = os.getcwd()
        os.chdir(RUNDIR)

    def __exit__(self, *a):
        os.chdir(self.cwd)


def _caps_arr(caps):
    if caps is None:
        return None
    return (ctypes.c_int * 6)(*caps)


def reset(mode, level, tb, serial, template=ACCOUNT1, squad_agents=False, caps=None, max_steps=0):
    L = lib()
    with _InRundir():
        rc = L.sfref_reset(mode, level, tb, serial, os.path.abspath(template).encode(),
                           int(squad_agents), _caps_arr(caps), max_steps)
    if rc != 0:
        raise RuntimeError("sfref_reset failed")


def reset_logging(mode, level, template=ACCOUNT1, squad_agents=False, caps=None, max_steps=0):
    """Start a match with the reference's own .sf_sample logging on (gameplay.hpp:1784-1794, 966-967).
    The seeds are the reference's (time based): returns (tb, serial)."""
    L = lib()
    with _InRundir():
        rc = L.sfref_reset_ex(mode, level, os.path.abspath(template).encode(), int(squad_agents), _caps_arr(caps),
                              max_steps, 1, None)
    if rc != 0:
        raise RuntimeError("sfref_reset_ex failed")
    out = np.zeros(2, dtype=np.int64)
    L.sfref_seeds(out.ctypes.data)
    return int(out[0]), int(out[1])


def close_log():
    """Close the running log and return its absolute path."""
    buf = ctypes.create_string_buffer(4096)
    lib().sfref_close_log(buf, 4096)
    return os.path.join(RUNDIR, buf.value.decode())


def reset_replay(mode, level, path, squad_agents=False, caps=None, max_steps=0):
    """Start a match that the reference replays from a .sf_sample file with its own reader
    (gameplay.hpp:1749-1783, 968-993): seeds, player sheet and every command come from the file."""
    L = lib()
    with _InRundir():
        rc = L.sfref_reset_ex(mode, level, os.path.abspath(ACCOUNT1).encode(), int(squad_agents), _caps_arr(caps),
                              max_steps, 0, os.path.abspath(path).encode())
    if rc != 0:
        raise RuntimeError("sfref_reset_ex failed")


def reset_online(port, password, user="1", ip="127.0.0.1", caps=None, max_steps=0):
    """Join a live online match with the reference's own client code (gameplay.hpp:66-193, 1807-1859);
    returns this player's index.  Blocks until the server has everybody's sheet."""
    L = lib()
    with _InRundir():
        rc = L.sfref_reset_online(ip.encode(), int(port), password.encode(), user.encode(), _caps_arr(caps), max_steps)
    if rc < 0:
        raise RuntimeError("sfref_reset_online failed (%d)" % rc)
    return rc


def step(actions):
    """actions: bytes, one command symbol per human slot (slot 0 = the player)."""
    L = lib()
    with _InRundir():
        return L.sfref_step(bytes(actions), len(actions))


def status():
    return lib().sfref_status()


def observe(slot=0):
    out = np.empty(OBS_LEN, dtype=np.float32)
    n = lib().sfref_observe(slot, out.ctypes.data)
    if n != OBS_LEN:
        raise RuntimeError("sfref_observe: slot %d has no active agent" % slot)
    return out


def set_capture(slot, on=True):
    lib().sfref_set_capture(slot, int(on))


def get_captured(slot):
    out = np.empty(OBS_LEN, dtype=np.float32)
    n = lib().sfref_get_captured(slot, out.ctypes.data)
    return out if n == OBS_LEN else None


def counters():
    out = np.zeros(8, dtype=np.int64)
    lib().sfref_counters(out.ctypes.data)
    return dict(zip(["frame", "kills", "teams_kills", "loot", "chest", "steps", "status", "hp"], out.tolist()))


def population():
    out = np.zeros(6, dtype=np.int32)
    lib().sfref_population(out.ctypes.data)
    return dict(zip(["humans", "zombies", "bullets", "chests", "built", "portals"], out.tolist()))


def dump():
    buf = np.empty(1 << 20, dtype=np.int32)
    n = lib().sfref_dump(buf.ctypes.data, buf.size)
    if n < 0:
        raise RuntimeError("sfref_dump: buffer too small")
    return buf[:n].copy()


def state_hash():
    return int(lib().sfref_hash())


def srand(tb, serial):
    lib().sfref_srand(tb, serial)


def rand():
    return lib().sfref_rand()


def rng_state():
    out = np.zeros(19, dtype=np.int64)
    lib().sfref_rng_state(out.ctypes.data)
    return out


def compute_damage(x, y):
    return lib().sfref_compute_damage(x, y)


def run_stream(env, mode, level, n_steps, table, template=ACCOUNT1, squad_agents=False, caps=None,
               max_steps=0, with_obs=False):
    """Free-running synthetic workload (include/sf_synth.h) for the CPU baseline."""
    L = lib()
    h = ctypes.c_ulonglong(0)
    with _InRundir():
        n = L.sfref_run_stream(env, mode, level, os.path.abspath(template).encode(), int(squad_agents),
                               _caps_arr(caps), max_steps, table.encode(), len(table), n_steps,
                               int(with_obs), ctypes.byref(h))
    return n, int(h.value)
