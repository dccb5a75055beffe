#!/usr/bin/env python
"""The parity soak of tests/test_gpu_soak.py at several times its length (thousands of steps per
arena); usage: python tools/gpu_soak.py [case ...]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import test_gpu_soak as soak  # noqa: E402
from strikeforce_b200 import data as sfdata  # noqa: E402

arena = sfdata.load_default()
t0 = time.time()
for name in (sys.argv[1:] or sorted(soak.CASES)):
    print("soak ok:", soak.run_case(torch, arena, name, long=True), "(%.0f s)" % (time.time() - t0), flush=True)
print("ALL SOAKS OK")
