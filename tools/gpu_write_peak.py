#!/usr/bin/env python
"""What does this GPU sustain for a PURE WRITE stream?  The observation kernel writes 16 GB per launch and reads
almost nothing; the peak in MEASURED_PEAKS.json is a copy (half reads, half writes).  Times, on 16 GB:
a device-to-device copy (bytes = read + written), torch's fill kernel, cudaMemsetAsync, and a plain 16-byte
streaming-store kernel shaped like the observation kernel's copy-out (compiled here with nvcc, loaded with ctypes)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import torch

GB = 1 << 30
N = 16 * GB
dev = torch.device("cuda", 0)
a = torch.empty(N, dtype=torch.uint8, device=dev)
b = torch.empty(N // 2, dtype=torch.uint8, device=dev)
c = torch.empty(N // 2, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
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


ms = timed(lambda: c.copy_(b))
print("copy 8 GB -> 8 GB : %.3f ms = %.0f GB/s (read + written)" % (ms, N / ms / 1e6))
ms = timed(lambda: a.zero_())
print("torch fill 16 GB  : %.3f ms = %.0f GB/s written" % (ms, N / ms / 1e6))
rt = C.CDLL("libcudart.so.12")
rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
ms = timed(lambda: rt.cudaMemsetAsync(a.data_ptr(), 0, N, torch.cuda.current_stream().cuda_stream))
print("cudaMemsetAsync   : %.3f ms = %.0f GB/s written" % (ms, N / ms / 1e6))

SRC = r"""
#include <cuda_runtime.h>
extern "C" __global__ void __launch_bounds__(128, 8) fill(float4 *dst, size_t n_chunks, size_t per_cta)
{   /* every CTA streams contiguous blocks of per_cta chunks (an observation is 7,688), 512 B per warp instruction */
    for (size_t blk = blockIdx.x; blk * per_cta < n_chunks; blk += gridDim.x) {
        float4 *d = dst + blk * per_cta;
        for (size_t q = threadIdx.x; q < per_cta; q += 128) __stcs(d + q, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}
extern "C" __global__ void __launch_bounds__(128, 8) fill8(float4 *dst, size_t n_obs, int run)
{   /* the [32][31][31] copy-out order: chunk q of each of the eight 4-channel groups (961 chunks apart) per iteration;
       a CTA writes `run` consecutive observations before it moves on */
    for (size_t r = blockIdx.x; r * run < n_obs; r += gridDim.x)
        for (size_t o = r * run; o < r * run + run && o < n_obs; ++o) {
            float4 *d = dst + o * 7688;
            for (int q = threadIdx.x; q < 961; q += 128)
#pragma unroll
                for (int g = 0; g < 8; ++g) __stcs(d + g * 961 + q, make_float4(0.f, 0.f, 0.f, 0.f));
        }
}
extern "C" __global__ void __launch_bounds__(128, 8) fillg(float4 *dst, size_t n_obs, int run)
{   /* front to back, but group by group with a thread's chunk fixed within the group (q = tid + 128 i): the warp
       stores of group g start 16 * g bytes off a 512-byte boundary */
    for (size_t r = blockIdx.x; r * run < n_obs; r += gridDim.x)
        for (size_t o = r * run; o < r * run + run && o < n_obs; ++o) {
            float4 *d = dst + o * 7688;
            for (int g = 0; g < 8; ++g)
                for (int q = threadIdx.x; q < 961; q += 128) __stcs(d + g * 961 + q, make_float4(0.f, 0.f, 0.f, 0.f));
        }
}
extern "C" __global__ void __launch_bounds__(128, 8) fill32(float *dst, size_t n_units, size_t per_cta)
{   /* the same walk with 32-byte stores (st.global.v8, sm_100): 1 KB per warp instruction */
    for (size_t blk = blockIdx.x; blk * per_cta < n_units; blk += gridDim.x) {
        float *d = dst + blk * per_cta * 8;
        for (size_t q = threadIdx.x; q < per_cta; q += 128)
            asm volatile("st.global.cs.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(d + q * 8), "f"(0.f) : "memory");
    }
}
extern "C" void launch(void *dst, size_t n_chunks, size_t per_cta, int grid, void *stream)
{
    fill<<<grid, 128, 0, (cudaStream_t)stream>>>((float4 *)dst, n_chunks, per_cta);
}
extern "C" void launch32(void *dst, size_t n_units, size_t per_cta, int grid, void *stream)
{
    fill32<<<grid, 128, 0, (cudaStream_t)stream>>>((float *)dst, n_units, per_cta);
}
extern "C" void launchg(void *dst, size_t n_obs, int run, int grid, void *stream)
{
    fillg<<<grid, 128, 0, (cudaStream_t)stream>>>((float4 *)dst, n_obs, run);
}
extern "C" void launch8(void *dst, size_t n_obs, int run, int grid, void *stream)
{
    fill8<<<grid, 128, 0, (cudaStream_t)stream>>>((float4 *)dst, n_obs, run);
}
"""
with tempfile.TemporaryDirectory() as td:
    open(os.path.join(td, "k.cu"), "w").write(SRC)
    so = os.path.join(td, "k.so")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared",
                           "-Xcompiler", "-fPIC", os.path.join(td, "k.cu"), "-o", so])
    k = C.CDLL(so)
    k.launch.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for per in (7688, 4 * 7688, 65536):
        n_chunks = (N // 16) // per * per
        for ctas in (8, 16):
            ms = timed(lambda: k.launch(a.data_ptr(), n_chunks, per, ctas * sms, torch.cuda.current_stream().cuda_stream))
            print("store kernel, %6d chunks per block, %2d CTAs/SM: %.3f ms = %.0f GB/s written" % (per, ctas, ms, n_chunks * 16 / ms / 1e6))
    k.launch8.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    k.launchg.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    n_obs = 131072
    for run in (1, 4, 16):
        ms = timed(lambda: k.launch8(a.data_ptr(), n_obs, run, 8 * sms, torch.cuda.current_stream().cuda_stream))
        print("store kernel, eight-group order, runs of %2d observations: %.3f ms = %.0f GB/s written" % (run, ms, n_obs * 123008 / ms / 1e6))
        ms = timed(lambda: k.launchg(a.data_ptr(), n_obs, run, 8 * sms, torch.cuda.current_stream().cuda_stream))
        print("store kernel, front to back by groups (misaligned warp stores), runs of %2d: %.3f ms = %.0f GB/s written" % (run, ms, n_obs * 123008 / ms / 1e6))
        ms = timed(lambda: k.launch(a.data_ptr(), n_obs * 7688, 7688 * run, 8 * sms, torch.cuda.current_stream().cuda_stream))
        print("store kernel, front-to-back order, runs of %2d observations: %.3f ms = %.0f GB/s written" % (run, ms, n_obs * 123008 / ms / 1e6))
    k.launch32.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
    for run in (1, 16):
        ms = timed(lambda: k.launch32(a.data_ptr(), n_obs * 3844, 3844 * run, 8 * sms, torch.cuda.current_stream().cuda_stream))
        print("store kernel, front to back, 32-byte stores (st.global.v8), runs of %2d observations: %.3f ms = %.0f GB/s written"
              % (run, ms, n_obs * 123008 / ms / 1e6))
    # the same front-to-back walk with every warp store starting `shift` chunks (16 bytes each) off a 512-byte boundary
    for shift in (0, 1, 2, 3, 4, 8, 16):
        ms = timed(lambda: k.launch(a.data_ptr() + 16 * shift, n_obs * 7688 - 32, 7688 * 16, 8 * sms, torch.cuda.current_stream().cuda_stream))
        print("store kernel, front to back, runs of 16, warp stores %2d chunks off a 512-byte boundary: %.3f ms = %.0f GB/s written"
              % (shift, ms, n_obs * 123008 / ms / 1e6))
sys.stdout.flush()
