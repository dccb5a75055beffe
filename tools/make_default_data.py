#!/usr/bin/env python
"""Snapshot the shipped arena of a reference checkout into strikeforce_b200/data/default_arena.json.

Reads the reference's DATA files (map/, Items/, character/, the test account sheet) through
strikeforce_b200.data.load_reference_dir and stores them run-length encoded.  Run in the
build container (needs /root/reference or oracle/_ref/rundir); the JSON travels with the repo.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strikeforce_b200 import data as sfdata  # noqa: E402


def main():
    ref = os.environ.get("SF_REFERENCE", "/root/reference") + "/StrikeForce-client"
    if os.path.isdir(ref):
        acct = os.path.join(ref, "accounts", "game", "1", "info, 1.txt")
        d = sfdata.load_reference_dir(ref, {"account1": acct})
    else:
        run = os.path.join(ROOT, "oracle", "_ref", "rundir")
        d = sfdata.load_reference_dir(run, {"account1": os.path.join(run, "player_account1.txt")})
    d.player_sheets["synthetic"] = sfdata.synthetic_player_sheet()
    with open(sfdata.DEFAULT_JSON, "w") as f:
        json.dump(sfdata.to_json(d), f, separators=(",", ":"))
        f.write("\n")
    back = sfdata.load_default()
    assert (back.map_cells == d.map_cells).all() and (back.map_portal == d.map_portal).all()
    print("wrote", sfdata.DEFAULT_JSON, os.path.getsize(sfdata.DEFAULT_JSON), "bytes")


if __name__ == "__main__":
    main()
