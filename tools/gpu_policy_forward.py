#!/usr/bin/env python
"""What does the batched AgentModel forward of configs[4] cost, and which library settings move it?
One chunk of 32,768 observations (the royale16 tick runs 16 of them): fp32 with TF32 convolutions (cuDNN default),
+ cudnn.benchmark, + channels_last, strict fp32, and (for scale only: reduced precision) bf16 autocast."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strikeforce_b200 import policy  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = policy.AgentModel().to(dev).eval()
x = (torch.rand(B, 32, 31, 31, device=dev) < 0.08).float() * torch.rand(B, 32, 31, 31, device=dev)
state = model.initial_state(B, dev)


def timed(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd(inp=x, m=model):
    with torch.no_grad():
        return m(inp, state)


def stages():
    with torch.no_grad():
        c = model.backbone.cnn
        t0 = timed(lambda: c.conv0(x))
        y0 = c.conv0(x)
        t1 = timed(lambda: c.conv1(y0))
        y1 = c.conv1(y0)
        t2 = timed(lambda: c.conv3(c.conv2(y1)))
    return t0, t1, t2


torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = True
print("B = %d observations (%.1f GB)" % (B, x.numel() * 4 / 1e9))
print("tf32 convolutions (cuDNN default): %.1f ms   conv0 %.1f conv1 %.1f conv2+3 %.1f" % ((timed(fwd),) + stages()))
torch.backends.cudnn.benchmark = True
print("  + cudnn.benchmark:                %.1f ms   conv0 %.1f conv1 %.1f conv2+3 %.1f" % ((timed(fwd),) + stages()))
xcl = x.contiguous(memory_format=torch.channels_last)
mcl = policy.AgentModel().to(dev).eval().to(memory_format=torch.channels_last)
print("  + channels_last input and weights: %.1f ms (conversion of the input not included: %.1f ms)" % (
    timed(lambda: fwd(xcl, mcl)), timed(lambda: x.contiguous(memory_format=torch.channels_last))))
torch.backends.cudnn.allow_tf32 = False
print("strict fp32 convolutions (+benchmark): %.1f ms" % timed(fwd, 1))
torch.backends.cudnn.allow_tf32 = True
with torch.autocast("cuda", dtype=torch.bfloat16):
    print("bf16 autocast (reduced precision, for scale only): %.1f ms" % timed(fwd))

# where the time of one forward goes, kernel by kernel (torch.profiler, CUDA time)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

torch.backends.cudnn.allow_tf32 = True
fwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fwd()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
