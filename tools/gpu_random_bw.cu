// Development aid: how many scattered 64-byte granules per second HBM delivers (the access pattern
// of the cell overlay), against a streaming read of the same buffer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/gpu_random_bw.cu -o build_variants/random_bw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <int MLP, int BYTES>
__global__ void random_read(const uint8_t *buf, uint64_t granules, int iters, uint64_t *sink, int write_every)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint64_t v[MLP];
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            uint64_t g = mix(tid * 1315423911ull + (uint64_t)it * MLP + j) % granules;
            v[j] = *(const uint16_t *)(buf + g * BYTES + (tid & 31) * 2 % BYTES);
            if (write_every && ((it * MLP + j) % write_every) == 0) *(uint16_t *)(buf + g * BYTES) = (uint16_t)tid;
        }
#pragma unroll
        for (int j = 0; j < MLP; ++j) acc += v[j];
    }
    if (acc == 0x1234567) *sink = acc;
}

__global__ void stream_read(const uint4 *buf, uint64_t n, uint64_t *sink)
{
    uint64_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = buf[i];
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x1234567) *sink = acc;
}

// pattern of the step kernel: the 32 lanes of a warp read one granule each from 32 neighbouring
// 20 KB blocks (the overlays of a chunk of arenas); the chunk is fixed per warp (RANDOM_CHUNK 0)
// or drawn per access (1)
template <int MLP, int RANDOM_CHUNK>
__global__ void chunk_read(const uint8_t *buf, uint64_t chunks, int iters, uint64_t *sink)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, acc = 0, warp = tid >> 5, lane = tid & 31;
    for (int it = 0; it < iters; ++it) {
        uint64_t v[MLP];
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            uint64_t chunk = RANDOM_CHUNK ? mix(warp * 77777ull + (uint64_t)it * MLP + j) % chunks : warp % chunks;
            uint64_t g = mix(tid * 1315423911ull + (uint64_t)it * MLP + j) % 312;
            v[j] = *(const uint16_t *)(buf + (chunk * 32 + lane) * 19968 + g * 64);
        }
#pragma unroll
        for (int j = 0; j < MLP; ++j) acc += v[j];
    }
    if (acc == 0x1234567) *sink = acc;
}

// the zombie pattern: own cell and its four neighbours in a 4x8-tiled overlay (sf_state.h), issued
// back to back.  PLUS 0: the own cell only; 1: five scalar loads; 2: own tile row as one 16-byte load
// plus the two vertical neighbours
__device__ __forceinline__ int step_cell(int t, int d)
{
    const int inr = (t >> 3) & 3, inc = t & 7;
    if (d == 0) return inr != 3 ? t + 8 : t + 13 * 32 - 24;
    if (d == 1) return inc != 7 ? t + 1 : t + 32 - 7;
    if (d == 2) return inr != 0 ? t - 8 : t - 13 * 32 + 24;
    return inc != 0 ? t - 1 : t - 32 + 7;
}
template <int PLUS>
__global__ void plus_read(const uint8_t *buf, uint64_t chunks, int iters, uint64_t *sink)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, acc = 0, warp = tid >> 5, lane = tid & 31;
    const uint16_t *g = (const uint16_t *)(buf + ((warp % chunks) * 32 + lane) * 19968);
    for (int it = 0; it < iters; ++it) {
        int cell = 13 * 32 + 32 + (int)(mix(tid * 1315423911ull + it) % (6 * 11 * 32)); // an interior tile
        cell = (cell / (11 * 32) + 1) * 13 * 32 + (cell % (11 * 32)) + 32;
        if (PLUS == 0) acc += g[cell];
        if (PLUS == 1) {
            uint32_t v0 = g[cell], v1 = g[step_cell(cell, 0)], v2 = g[step_cell(cell, 1)], v3 = g[step_cell(cell, 2)],
                     v4 = g[step_cell(cell, 3)];
            acc += v0 + v1 + v2 + v3 + v4;
        }
        if (PLUS == 2) {
            uint4 row = *(const uint4 *)(g + (cell & ~7));
            uint32_t v1 = g[step_cell(cell, 0)], v3 = g[step_cell(cell, 2)];
            uint32_t v2 = (cell & 7) != 7 ? 0u : g[step_cell(cell, 1)], v4 = (cell & 7) != 0 ? 0u : g[step_cell(cell, 3)];
            acc += row.x + row.y + row.z + row.w + v1 + v2 + v3 + v4;
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
// read-modify-write of one 2-byte cell per access, the way the tick updates the overlay: MODE 0 read
// only, 1 read then write the same cell (the store waits for the load), 2 write only, 3 read, then write
// a different cell of the same granule
template <int MODE>
__global__ void rmw_cells(uint8_t *buf, uint64_t chunks, int iters, uint64_t *sink)
{
    uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x, acc = 0, warp = tid >> 5, lane = tid & 31;
    uint16_t *g = (uint16_t *)(buf + ((warp % chunks) * 32 + lane) * 19968);
    for (int it = 0; it < iters; ++it) {
        int cell = (int)(mix(tid * 1315423911ull + it) % 9984);
        uint32_t v = 0;
        if (MODE != 2) v = g[cell];
        if (MODE == 1) g[cell] = (uint16_t)(v ^ 0x400u);
        if (MODE == 2) g[cell] = (uint16_t)it;
        if (MODE == 3) g[cell ^ 1] = (uint16_t)(v ^ 0x400u);
        acc += v;
    }
    if (acc == 0x1234567) *sink = acc;
}
template <int MODE> void run_rmw(uint8_t *buf, size_t bytes, uint64_t *sink)
{
    int cta = 128, grid = 148 * 896 / cta, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    uint64_t chunks = bytes / (32 * 19968);
    rmw_cells<MODE><<<grid, cta>>>(buf, chunks, iters, sink);
    cudaEventRecord(e0);
    rmw_cells<MODE><<<grid, cta>>>(buf, chunks, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)grid * cta * iters;
    printf("cell update %d (0 read, 1 read+write same cell, 2 write only, 3 read+write neighbour): %.2f G cells/s\n", MODE,
           n / ms / 1e6);
}

template <int PLUS> void run_plus(const uint8_t *buf, size_t bytes, uint64_t *sink)
{
    int cta = 128, grid = 148 * 896 / cta, iters = 2048;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    uint64_t chunks = bytes / (32 * 19968);
    plus_read<PLUS><<<grid, cta>>>(buf, chunks, iters, sink);
    cudaEventRecord(e0);
    plus_read<PLUS><<<grid, cta>>>(buf, chunks, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)grid * cta * iters;
    printf("plus pattern %d (0 own cell, 1 five scalar loads, 2 row vector + verticals): %.2f G neighbourhoods/s\n", PLUS,
           n / ms / 1e6);
}

template <int MLP, int RANDOM_CHUNK> void run_chunk(const uint8_t *buf, size_t bytes, uint64_t *sink, int threads_per_sm)
{
    int cta = 128, grid = 148 * threads_per_sm / cta, iters = 4096 / MLP;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    uint64_t chunks = bytes / (32 * 19968);
    chunk_read<MLP, RANDOM_CHUNK><<<grid, cta>>>(buf, chunks, iters, sink);
    cudaEventRecord(e0);
    chunk_read<MLP, RANDOM_CHUNK><<<grid, cta>>>(buf, chunks, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)grid * cta * iters * MLP;
    printf("chunk pattern (%s chunk per warp), footprint %.1f GB, %4d threads/SM, MLP %d: %.2f G accesses/s\n",
           RANDOM_CHUNK ? "random" : "fixed", bytes / 1e9, threads_per_sm, MLP, n / ms / 1e6);
}

template <int MLP, int BYTES> void run(const uint8_t *buf, size_t bytes, uint64_t *sink, int threads_per_sm, int write_every)
{
    int cta = 256, grid = 148 * threads_per_sm / cta, iters = 4096 / MLP;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    random_read<MLP, BYTES><<<grid, cta>>>(buf, bytes / BYTES, iters, sink, write_every);
    cudaEventRecord(e0);
    random_read<MLP, BYTES><<<grid, cta>>>(buf, bytes / BYTES, iters, sink, write_every);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)grid * cta * iters * MLP;
    printf("granule %3d B, footprint %.2f GB, %4d threads/SM, MLP %d, write 1/%d: %.2f G accesses/s = %.0f GB/s of granules\n",
           BYTES, bytes / 1e9, threads_per_sm, MLP, write_every, n / ms / 1e6, n * BYTES / ms / 1e6);
}

int main()
{
    size_t bytes = (size_t)3500 << 20;
    uint8_t *buf;
    uint64_t *sink;
    cudaMalloc(&buf, bytes), cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    stream_read<<<148 * 8, 512>>>((const uint4 *)buf, bytes / 16, sink);
    cudaEventRecord(e0);
    stream_read<<<148 * 8, 512>>>((const uint4 *)buf, bytes / 16, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("streaming read: %.0f GB/s\n", bytes / ms / 1e6);
    run<1, 64>(buf, bytes, sink, 896, 0);
    run<2, 64>(buf, bytes, sink, 896, 0);
    run<4, 64>(buf, bytes, sink, 896, 0);
    run<8, 64>(buf, bytes, sink, 896, 0);
    run<8, 64>(buf, bytes, sink, 2048, 0);
    run<8, 32>(buf, bytes, sink, 2048, 0);
    run<8, 128>(buf, bytes, sink, 2048, 0);
    run<4, 64>(buf, bytes, sink, 896, 4);
    run<8, 64>(buf, bytes, sink, 2048, 4);
    run<8, 64>(buf, (size_t)64 << 20, sink, 2048, 0);
    run<8, 64>(buf, (size_t)256 << 20, sink, 2048, 0);
    run<8, 64>(buf, (size_t)1024 << 20, sink, 2048, 0);
    run_chunk<4, 0>(buf, bytes, sink, 896);
    run_chunk<4, 1>(buf, bytes, sink, 896);
    run_chunk<1, 0>(buf, bytes, sink, 896);
    run_chunk<4, 0>(buf, (size_t)1024 << 20, sink, 896);
    run_chunk<4, 0>(buf, (size_t)2560 << 20, sink, 896);
    run_rmw<0>(buf, bytes, sink);
    run_rmw<1>(buf, bytes, sink);
    run_rmw<2>(buf, bytes, sink);
    run_rmw<3>(buf, bytes, sink);
    run_plus<0>(buf, bytes, sink);
    run_plus<1>(buf, bytes, sink);
    run_plus<2>(buf, bytes, sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
