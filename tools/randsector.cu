// Random-access granularity microbenchmark: GB/s and Gaccess/s of independent random reads of
// 32 / 64 / 128 contiguous bytes from a buffer far larger than L2.  Decides whether a layout that
// turns a trilinear sample into ONE 32-byte sector can beat the texture layout (~64 B/sample).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }
template <int BYTES>
__global__ void rd(const float4* __restrict__ buf, uint64_t nchunks, int per_thread, float* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (int it = 0; it < per_thread; it += 4) {
        float4 v[4][BYTES / 16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t r = ((uint64_t)mix(tid * 977u + (it + u) * 7919u) << 32 | mix(tid + (it + u) * 104729u)) % nchunks;
#pragma unroll
            for (int k = 0; k < BYTES / 16; ++k) v[u][k] = __ldg(buf + r * (BYTES / 16) + k);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < BYTES / 16; ++k) acc += v[u][k].x + v[u][k].w;
    }
    if (acc == 1234.5f) out[0] = acc;
}
template <int BYTES> void run(const float4* buf, size_t bytes, float* out) {
    const int threads = 256, blocks = 148 * 32, per = 64;
    const uint64_t nchunks = bytes / BYTES;
    rd<BYTES><<<blocks, threads>>>(buf, nchunks, per, out);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) rd<BYTES><<<blocks, threads>>>(buf, nchunks, per, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const double acc = (double)blocks * threads * per;
    printf("random %3d-byte reads: %7.1f Gaccess/s  %7.1f GB/s useful  (%.3f ms)\n", BYTES, acc / ms / 1e6, acc * BYTES / ms / 1e6, ms);
}
int main() {
    const size_t bytes = (size_t)32 << 30;
    float4* buf; float* out;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 4); cudaMemset(buf, 0, bytes);
    run<32>(buf, bytes, out); run<64>(buf, bytes, out); run<128>(buf, bytes, out); run<16>(buf, bytes, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
