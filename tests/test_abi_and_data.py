"""C-ABI surface, data loaders and the host-side mirror of the plugin interface (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from strikeforce_b200 import config as sfcfg
from strikeforce_b200 import data as sfdata
from strikeforce_b200 import lib as sflib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """Every function include/strikeforce_b200.h declares is exported by the shared library."""
    hdr = open(os.path.join(ROOT, "include", "strikeforce_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sf_[a-z_0-9]+)\s*\(", hdr)) - {"sf_mix64"})
    assert len(declared) >= 15
    L = sflib.lib()
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    assert sorted(sflib.EXPORTS) == declared
    assert L.sf_abi_version() == sfcfg.ABI_VERSION


def test_no_cpu_fallback(arena_data):
    """Without a CUDA device the product fails loudly (SF_ERR_NO_DEVICE); nothing simulates on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = sfcfg.make_config(arena_data, n_envs=4)
    h = C.c_void_p()
    rc = sflib.lib().sf_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU path" in sflib.lib().sf_last_error(None)
    from strikeforce_b200.sim import BatchedArena
    with pytest.raises(RuntimeError):
        BatchedArena(4)
    src = open(os.path.join(ROOT, "strikeforce_b200", "sim.py")).read() + open(
        os.path.join(ROOT, "strikeforce_b200", "lib.py")).read()
    assert "oracle" not in src and "hostcheck" not in src, "the product must not reach into the test models"


def test_python_mirror_of_the_header_constants():
    """The Python mirror (strikeforce_b200/config.py) carries the header's values."""
    hdr = open(os.path.join(ROOT, "include", "strikeforce_b200.h")).read()

    def const(name):
        m = re.search(r"\b%s\s*=?\s*(0x[0-9A-Fa-f]+|-?\d+)" % name, hdr)
        assert m, name + " not found in the header"
        return int(m.group(1), 0)

    assert const("SF_ABI_VERSION") == sfcfg.ABI_VERSION
    assert (const("SF_OBS_P1"), const("SF_OBS_P2"), const("SF_OBS_NHWC")) == (sfcfg.OBS_P1, sfcfg.OBS_P2, sfcfg.OBS_NHWC)
    assert sfcfg.OBS_NHWC & (sfcfg.OBS_P1 | sfcfg.OBS_P2) == 0  # a flag or-ed into the phase
    assert (const("SF_OBS_CH"), const("SF_OBS_WIN")) == (sfcfg.OBS_CH, sfcfg.OBS_WIN)


def test_struct_layout_matches_header():
    assert C.sizeof(sfcfg.StepOut) == 32
    assert sfcfg.SfConfig.map_cells.offset % 8 == 0
    assert sfcfg.SfConfig.royale_players.offset == sfcfg.SfConfig.npc_sheet.offset + 4 * sfcfg.SHEET_LEN
    assert sfcfg.SfConfig.royale_sheets.offset == sfcfg.SfConfig.royale_teams.offset + 4 * sfcfg.MAX_PLAYERS
    assert C.sizeof(sfcfg.SfConfig) == sfcfg.SfConfig.royale_sheets.offset + 4 * sfcfg.MAX_PLAYERS * sfcfg.SHEET_LEN + 4  # tail padding to 8


def test_default_arena_and_reference_parser(arena_data, tmp_path):
    d = arena_data
    cells = d.map_cells.reshape(3, 30, 100)
    assert (cells == ord("O")).sum() == 6 and (cells[:, 1, 38] == ord("O")).all() and (cells[:, 1, 93] == ord("O")).all()
    assert d.player_sheet("account1")[:3].tolist() == [15000, 1000, 15000]
    assert d.npc_sheet[2] == 1000000
    # write the arena back in the reference's text formats (CRLF, '^ n' tokens) and parse it again
    for k in range(3):
        os.makedirs(tmp_path / "map", exist_ok=True)
        with open(tmp_path / "map" / ("floor%d.txt" % (k + 1)), "wb") as f:
            for r in range(30):
                row = b""
                for c in range(100):
                    ch = bytes([cells[k, r, c]])
                    row += ch + (b" %d " % d.map_portal.reshape(3, 30, 100)[k, r, c] if ch in b"^v" else b"")
                f.write(row + b"\r\n")
    os.makedirs(tmp_path / "Items"), os.makedirs(tmp_path / "character")
    for i in range(4):
        (tmp_path / "Items" / ("cons%d.txt" % i)).write_text("c%d 1 1 1 %d\r\n%d %d\r\n" % ((i,) + tuple(d.consumables[i])))
        (tmp_path / "Items" / ("throw%d.txt" % i)).write_text("t%d 1 1 1 %d\r\n%d %d %d\r\n" % ((i,) + tuple(d.throwables[i])))
    for i in range(8):
        (tmp_path / "Items" / ("w%d.txt" % i)).write_text("w%d 1 1 0 %d\r\n%d %d %d\r\n" % ((i,) + tuple(d.weapons[i])))
    (tmp_path / "character" / "human_enemy.txt").write_text("\r\n".join(str(v) for v in d.npc_sheet))
    (tmp_path / "character" / "human.txt").write_text("\r\n".join(str(v) for v in d.player_sheet("new_player")))
    (tmp_path / "acct.txt").write_text("bob\n" + "\n".join(str(v) for v in d.player_sheet("account1")))
    back = sfdata.load_reference_dir(str(tmp_path), {"account1": str(tmp_path / "acct.txt")})
    assert (back.map_cells == d.map_cells).all() and (back.map_portal == d.map_portal).all()
    assert (back.weapons == d.weapons).all() and (back.player_sheet("account1") == d.player_sheet("account1")).all()


def test_bad_configs_are_rejected(arena_data):
    import hostcheck
    for kw in (dict(caps=dict(cap_humans=200)), dict(caps=dict(cap_zombies=0)), dict(level_min=0), dict(level_min=99)):
        cfg = sfcfg.make_config(arena_data, n_envs=1, **kw)
        with pytest.raises(RuntimeError):
            hostcheck.HostSim(cfg)
    broken = sfdata.from_json(sfdata.to_json(arena_data))
    broken.map_cells = broken.map_cells.copy()
    broken.map_cells[0] = ord(".")  # open the border
    with pytest.raises(AssertionError):
        broken.validate()
