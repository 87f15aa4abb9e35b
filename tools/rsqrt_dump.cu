// rsqrt_dump.cu — what rsqrt.approx.f32 (MUFU.RSQ) returns on this GPU for every float in [lo, hi].
//
// The reference's d_render normalises the eye ray with helper_math.h's normalize() = v * rsqrtf(dot(v, v))
// (/root/reference/volumeRender_kernel.cu:295), which nvcc turns into rsqrt.approx.f32.  Its argument is
// u*u + v*v + 4 with |u|, |v| <= 1, i.e. a float in [4, 6].  A CPU cannot compute the instruction's result, but it can
// look it up: this tool dumps the result for all 4 194 305 floats of that interval so that the oracle can restate
// the reference build's ray set-up bit for bit (tools/make_rsqrt_table.py -> tests/golden/rsqrt_approx_b200_v1.npz).
// Test infrastructure; not part of libvrdd.so.
//     rsqrt_dump <out.u32> [lo_bits_hex hi_bits_hex]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__global__ void rsqrt_kernel(uint32_t lo, uint32_t n, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float r;
        asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(__uint_as_float(lo + i)));
        out[i] = __float_as_uint(r);
    }
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: rsqrt_dump <out.u32> [lo_bits_hex hi_bits_hex]\n"); return 2; }
    const uint32_t lo = argc > 3 ? (uint32_t)std::strtoul(argv[2], nullptr, 16) : 0x40800000u;   // 4.0f
    const uint32_t hi = argc > 3 ? (uint32_t)std::strtoul(argv[3], nullptr, 16) : 0x40C00000u;   // 6.0f
    const uint32_t n = hi - lo + 1;
    uint32_t* d = nullptr;
    if (cudaMalloc(&d, (size_t)n * 4) != cudaSuccess) { std::fprintf(stderr, "cudaMalloc failed\n"); return 1; }
    rsqrt_kernel<<<(n + 255) / 256, 256>>>(lo, n, d);
    std::vector<uint32_t> h(n);
    if (cudaMemcpy(h.data(), d, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { std::fprintf(stderr, "kernel/copy failed\n"); return 1; }
    FILE* f = std::fopen(argv[1], "wb");
    if (!f || std::fwrite(h.data(), 4, n, f) != n) { std::fprintf(stderr, "cannot write %s\n", argv[1]); return 1; }
    std::fclose(f);
    std::printf("rsqrt_dump: %u values from 0x%08X to 0x%08X -> %s\n", n, lo, hi, argv[1]);
    return 0;
}
