// probe_atom_linear.cu — DRAM bytes moved per random read of 4 / 32 / 64 / 128 contiguous, aligned bytes from plain
// global memory (cudaMalloc, 16 GB >> L2), with cudaLimitMaxL2FetchGranularity at its default and at 32.  Companion of
// probe_atom.cu (CUDA arrays).  Run under `ncu --metrics dram__bytes_read.sum`; launches print their order.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }
template <int BYTES, int MODE>
__global__ void rd(const float* __restrict__ buf, uint64_t nchunks, int n, uint32_t salt, float* out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    const uint64_t r = ((uint64_t)mix(tid * 2654435761u + salt) * nchunks) >> 32;       // distinct-ish 128-byte cells
    const float* p = buf + r * 32;                                                     // 128-byte aligned
    float acc = 0.f;
    if (BYTES == 4) {
        if (MODE == 0) acc = __ldg(p); else asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(acc) : "l"(p));
    } else {
#pragma unroll
        for (int k = 0; k < BYTES / 16; ++k) { const float4 v = __ldg(reinterpret_cast<const float4*>(p) + k); acc += v.x + v.w; }
    }
    if (acc == 1234.5f) out[0] = acc;
}
int main(int argc, char** argv) {
    if (argc > 1) { cudaFree(0); printf("set granularity %s: %d\n", argv[1], (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]))); }
    size_t g = 0; cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("cudaLimitMaxL2FetchGranularity = %zu\n", g);
    const size_t bytes = (size_t)16 << 30; const int n = 1 << 22;
    float *buf, *out;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 4); cudaMemset(buf, 0, bytes); cudaDeviceSynchronize();
    const uint64_t cells = bytes / 128;
    rd<4, 0><<<n / 256, 256>>>(buf, cells, n, 1, out);   cudaDeviceSynchronize(); printf("launch: 4-byte ld.global.nc\n");
    rd<4, 1><<<n / 256, 256>>>(buf, cells, n, 2, out);   cudaDeviceSynchronize(); printf("launch: 4-byte ld.global.nc.L1::no_allocate\n");
    rd<32, 0><<<n / 256, 256>>>(buf, cells, n, 3, out);  cudaDeviceSynchronize(); printf("launch: 32-byte\n");
    rd<64, 0><<<n / 256, 256>>>(buf, cells, n, 4, out);  cudaDeviceSynchronize(); printf("launch: 64-byte\n");
    rd<128, 0><<<n / 256, 256>>>(buf, cells, n, 5, out); cudaDeviceSynchronize(); printf("launch: 128-byte\n");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
