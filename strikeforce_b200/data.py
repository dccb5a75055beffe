"""Static inputs of the simulator: arena map, item tables, character sheets.

The reference reads these from text files every match (``map/floorK.txt`` in
``gameplay::setup`` gameplay.hpp:1252-1274, ``Items/*.txt`` in ``Item::download_items``
Item.hpp:179-188, ``character/*.txt`` and the account sheet in ``Human::build``
Character.hpp:650-709).  Here they are parsed once on the host into plain arrays
(:class:`ArenaData`) which ``sf_create`` uploads as constants shared by every arena.

Two sources:

* :func:`load_reference_dir` parses a ``StrikeForce-client`` style directory in the
  reference's own text formats (drop-in: point it at an existing checkout);
* :func:`load_default` loads ``data/default_arena.json``, a run-length encoded snapshot of
  the shipped arena produced by ``tools/make_default_data.py``.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

# The arena's dimensions are compile-time constants of the library, as they are in the reference
# (gameplay.hpp:37: 3 x 30 x 100).  SF_GEOMETRY=ROWSxCOLS in the environment selects a library built for a
# larger arena (csrc/build.sh) and makes load_default() embed the reference's map in one of that size.
REF_ROWS, REF_COLS = 30, 100
_geo = os.environ.get("SF_GEOMETRY", "")
FLOORS = 3
ROWS, COLS = (int(_geo.split("x")[0]), int(_geo.split("x")[1])) if _geo else (REF_ROWS, REF_COLS)
assert ROWS >= REF_ROWS and COLS >= REF_COLS and FLOORS * ((ROWS + 3) // 4) * ((COLS + 7) // 8) * 32 <= 16384, \
    "SF_GEOMETRY: at least 30x100, and cell ids must fit 14 bits"
GEOMETRY_TAG = "" if (ROWS, COLS) == (REF_ROWS, REF_COLS) else "_%dx%d" % (ROWS, COLS)
GEOMETRY_CFLAGS = ["-DSF_ROWS=%d" % ROWS, "-DSF_COLS=%d" % COLS] if GEOMETRY_TAG else []
CELLS = FLOORS * ROWS * COLS
SHEET_LEN = 32

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_JSON = os.path.join(_HERE, "data", "default_arena.json")


@dataclass
class ArenaData:
    map_cells: np.ndarray  # uint8 [CELLS], one of b"#.^vO"
    map_portal: np.ndarray  # int16 [CELLS], destination exit index of '^' / 'v', else -1
    consumables: np.ndarray  # int32 [4, 3]  stamina, hp, effect
    throwables: np.ndarray  # int32 [4, 4]  stamina, damage, effect, range
    weapons: np.ndarray  # int32 [8, 4]  (level-0 stats)
    npc_sheet: np.ndarray  # int32 [32]  character/human_enemy.txt
    player_sheets: dict = field(default_factory=dict)  # name -> int32 [32]

    def player_sheet(self, name="account1"):
        return self.player_sheets[name]

    def validate(self):
        assert self.map_cells.shape == (CELLS,) and self.map_cells.dtype == np.uint8
        assert self.map_portal.shape == (CELLS,) and self.map_portal.dtype == np.int16
        assert set(np.unique(self.map_cells).tolist()) <= set(b"#.^vO")
        cells = self.map_cells.reshape(FLOORS, ROWS, COLS)
        # the engine indexes neighbours without bounds tests (gameplay.hpp:664, 682); the
        # border must be solid apart from portal entrances
        for edge in (cells[:, 0, :], cells[:, -1, :], cells[:, :, 0], cells[:, :, -1]):
            assert not np.isin(edge, list(b".O")).any(), "arena border must be closed"
        n_exits = int((self.map_cells == ord("O")).sum())
        tgt = self.map_portal[np.isin(self.map_cells, list(b"^v"))]
        assert ((tgt >= 0) & (tgt < n_exits)).all(), "portal entrance without an exit"
        return self


def _tokens(path):
    with open(path, "rb") as f:
        return f.read().decode("latin-1").split()


def parse_floor(path):
    """One ``map/floorK.txt``: a stream of cell symbols where '^' and 'v' are followed by
    the index of their exit (gameplay.hpp:1254-1271).  Whitespace and CR are skipped the way
    ``ifstream >> char`` / ``>> int`` skip them."""
    with open(path, "rb") as f:
        text = f.read().decode("latin-1")
    cells = np.zeros(ROWS * COLS, dtype=np.uint8)
    portal = np.full(ROWS * COLS, -1, dtype=np.int16)
    pos, n = 0, 0
    while n < ROWS * COLS:
        while text[pos].isspace():
            pos += 1
        c = text[pos]
        pos += 1
        cells[n] = ord(c)
        if c in "^v":
            while text[pos].isspace():
                pos += 1
            start = pos
            while pos < len(text) and (text[pos].isdigit() or (pos == start and text[pos] in "+-")):
                pos += 1
            portal[n] = int(text[start:pos])
        n += 1
    return cells, portal


def parse_sheet(path, has_name):
    tok = _tokens(path)
    if has_name:
        tok = tok[1:]
    sheet = np.array([int(t) for t in tok[:SHEET_LEN]], dtype=np.int32)
    assert sheet.shape == (SHEET_LEN,), path
    return sheet


def load_reference_dir(client_dir, accounts=None):
    """Parse a reference client directory.  ``accounts`` maps a name to an account sheet
    path (``accounts/game/<user>/info, <user>.txt``, Character.hpp:657-660)."""
    cells, portal = [], []
    for k in range(FLOORS):
        c, p = parse_floor(os.path.join(client_dir, "map", "floor%d.txt" % (k + 1)))
        cells.append(c)
        portal.append(p)
    cons = np.zeros((4, 3), dtype=np.int32)
    for i in range(4):  # name price vol lvl stamina | Hp effect        (Item.hpp:70-75)
        t = _tokens(os.path.join(client_dir, "Items", "cons%d.txt" % i))
        cons[i] = [int(t[4]), int(t[5]), int(t[6])]
    thr = np.zeros((4, 4), dtype=np.int32)
    for i in range(4):  # name price vol lvl stamina | damage effect range (Item.hpp:149-154)
        t = _tokens(os.path.join(client_dir, "Items", "throw%d.txt" % i))
        thr[i] = [int(t[4]), int(t[5]), int(t[6]), int(t[7])]
    wpn = np.zeros((8, 4), dtype=np.int32)
    for i in range(8):
        t = _tokens(os.path.join(client_dir, "Items", "w%d.txt" % i))
        wpn[i] = [int(t[4]), int(t[5]), int(t[6]), int(t[7])]
    data = ArenaData(
        map_cells=np.concatenate(cells),
        map_portal=np.concatenate(portal),
        consumables=cons,
        throwables=thr,
        weapons=wpn,
        npc_sheet=parse_sheet(os.path.join(client_dir, "character", "human_enemy.txt"), has_name=False),
    )
    data.player_sheets["new_player"] = parse_sheet(os.path.join(client_dir, "character", "human.txt"), has_name=False)
    for name, path in (accounts or {}).items():
        data.player_sheets[name] = parse_sheet(path, has_name=True)
    return data.validate()


# ---------------------------------------------------------------- run-length encoded snapshot

def _rle_encode(arr):
    out, i = [], 0
    arr = list(arr)
    while i < len(arr):
        j = i
        while j < len(arr) and arr[j] == arr[i]:
            j += 1
        out.append([int(arr[i]), j - i])
        i = j
    return out


def _rle_decode(runs, dtype):
    return np.concatenate([np.full(n, v, dtype=dtype) for v, n in runs])


def to_json(data: ArenaData):
    return {
        "format": "strikeforce_b200.arena/1",
        "dims": [FLOORS, ROWS, COLS],
        "cells_rle": [[chr(v), n] for v, n in _rle_encode(data.map_cells)],
        "portal_rle": _rle_encode(data.map_portal),
        "consumables": data.consumables.tolist(),
        "throwables": data.throwables.tolist(),
        "weapons": data.weapons.tolist(),
        "npc_sheet": data.npc_sheet.tolist(),
        "player_sheets": {k: v.tolist() for k, v in data.player_sheets.items()},
    }


def enlarge(cells, portal):
    """The reference's 3 x 30 x 100 map as the top-left part of a ROWS x COLS one (a "large custom map",
    BASELINE.json configs[4]): the rest is open floor inside a closed border, reached through doors cut into the
    reference map's own bottom and right walls.  The static exits keep their scan order, so every '^' / 'v'
    still leads where it led."""
    big = np.full((FLOORS, ROWS, COLS), ord("."), dtype=np.uint8)
    bigp = np.full((FLOORS, ROWS, COLS), -1, dtype=np.int16)
    big[:, :REF_ROWS, :REF_COLS] = cells.reshape(FLOORS, REF_ROWS, REF_COLS)
    bigp[:, :REF_ROWS, :REF_COLS] = portal.reshape(FLOORS, REF_ROWS, REF_COLS)
    wall = ord("#")
    if ROWS > REF_ROWS:
        for c in range(5, REF_COLS - 1, 10):
            door = big[:, REF_ROWS - 1, c] == wall
            big[:, REF_ROWS - 1, c] = np.where(door, ord("."), big[:, REF_ROWS - 1, c])
    if COLS > REF_COLS:
        for r in range(3, REF_ROWS - 1, 7):
            door = big[:, r, REF_COLS - 1] == wall
            big[:, r, REF_COLS - 1] = np.where(door, ord("."), big[:, r, REF_COLS - 1])
    big[:, 0, :] = np.where(np.isin(big[:, 0, :], list(b".O")), wall, big[:, 0, :])
    big[:, :, 0] = np.where(np.isin(big[:, :, 0], list(b".O")), wall, big[:, :, 0])
    big[:, ROWS - 1, :] = wall
    big[:, :, COLS - 1] = wall
    return big.reshape(-1), bigp.reshape(-1)


def from_json(obj):
    assert obj["format"] == "strikeforce_b200.arena/1" and obj["dims"] in ([FLOORS, ROWS, COLS], [FLOORS, REF_ROWS, REF_COLS])
    cells = _rle_decode([[ord(v), n] for v, n in obj["cells_rle"]], np.uint8)
    portal = _rle_decode(obj["portal_rle"], np.int16)
    if obj["dims"] != [FLOORS, ROWS, COLS]:
        cells, portal = enlarge(cells, portal)
    data = ArenaData(
        map_cells=cells,
        map_portal=portal,
        consumables=np.array(obj["consumables"], dtype=np.int32),
        throwables=np.array(obj["throwables"], dtype=np.int32),
        weapons=np.array(obj["weapons"], dtype=np.int32),
        npc_sheet=np.array(obj["npc_sheet"], dtype=np.int32),
        player_sheets={k: np.array(v, dtype=np.int32) for k, v in obj["player_sheets"].items()},
    )
    return data.validate()


def load_default():
    with open(DEFAULT_JSON) as f:
        return from_json(json.load(f))


def synthetic_player_sheet():
    """a sheet with >= 8 of every consumable / throwable and every
    weapon at level 1, otherwise the new-player defaults of character/human.txt."""
    s = np.array([1000, 100, 1000, 1, 1, 1, 1000, 0, 0, 0, 0] + [8] * 4 + [1, 8] * 4 + [1] * 8 + [1], dtype=np.int32)
    assert s.shape == (SHEET_LEN,)
    return s
