// l1tex_peak.cu — peak rate of the operation the ray caster is made of: one trilinear fp32 tex3D fetch per sample
// (SURVEY.md §8d asks for the ray caster's roofline against a MEASURED L1TEX / L2 peak, 32 algorithmic bytes per fetch).
//
// Every thread marches a ray through a cubic fp32 3-D array like raycast_kernel does (8x4 pixel tiles per warp,
// normalised coordinates, linear filter, 8 fetches in flight), with nothing else in the loop.  Three working sets:
//   L1 : 32^3 texels (128 KB)  — every fetch hits L1TEX: the texture pipe's own peak
//   L2 : 256^3 texels (64 MB)  — L1 misses served by L2 (126 MB on B200)
//   HBM: 1024^3 texels (4.3 GB) at the ray caster's own sampling lattice (step 5 voxels, rays 2 voxels apart)
// and the same for tld4 on a layered array (the gather path: two tld4 per sample).  Prints Gfetch/s; 32 B per trilinear fetch.
// Measurement infrastructure, not product code.   l1tex_peak [json]
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>   // 0: tex3D linear   1: two tld4 on a layered array
__global__ void __launch_bounds__(256) march(cudaTextureObject_t tex, int N, int steps, float spread, float step, float* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = blockIdx.x % 64, by = blockIdx.x / 64;
    const int px = bx * 16 + (warp & 1) * 8 + (lane & 7), py = (by % 64) * 16 + (warp >> 1) * 4 + (lane >> 3);
    // rays along z, `spread` voxels apart in x and y, `step` voxels per step; positions wrap inside the volume
    float x = fmodf(px * spread, (float)N), y = fmodf(py * spread, (float)N), z = (float)((by / 64) * 37 % N);
    const float inv = 1.0f / (float)N;
    float acc = 0.f;
    for (int i = 0; i < steps; i += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) {
                v[k] = tex3D<float>(tex, x * inv, y * inv, z * inv);
            } else {
                float4 a, b;
                const int l = (int)z;
                asm volatile("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(tex), "r"(l), "f"(x), "f"(y));
                asm volatile("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(tex), "r"(min(l + 1, N - 1)), "f"(x), "f"(y));
                v[k] = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
            }
            z += step; if (z >= (float)N) z -= (float)N;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k];
    }
    if (acc == 1234.5f) out[0] = acc;
}

static double run(int mode, int N, float spread, float step, int steps, int blocks) {
    cudaArray_t arr; cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    if (cudaMalloc3DArray(&arr, &desc, make_cudaExtent(N, N, N), mode ? cudaArrayLayered : 0) != cudaSuccess) return -1.0;
    cudaResourceDesc rd; std::memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td; std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = mode ? cudaFilterModePoint : cudaFilterModeLinear; td.readMode = cudaReadModeElementType; td.normalizedCoords = mode ? 0 : 1;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    float* out; cudaMalloc(&out, 4);
    auto launch = [&] { if (mode) march<1><<<blocks, 256>>>(tex, N, steps, spread, step, out); else march<0><<<blocks, 256>>>(tex, N, steps, spread, step, out); };
    launch(); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) launch(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    cudaDestroyTextureObject(tex); cudaFreeArray(arr); cudaFree(out);
    return (double)blocks * 256 * steps / ms / 1e6;      // Gfetch/s (samples/s)
}

int main(int argc, char** argv) {
    const bool json = argc > 1 && !std::strcmp(argv[1], "json");
    struct { const char* name; int mode, N; float spread, step; int steps, blocks; } cfg[] = {
        {"tex3d_linear_l1", 0, 32, 0.37f, 0.61f, 512, 148 * 64}, {"tex3d_linear_l2", 0, 256, 1.9f, 5.1f, 256, 148 * 64},
        {"tex3d_linear_hbm_lattice", 0, 1024, 2.0f, 5.12f, 64, 4096 * 4},
        {"tld4_pair_l1", 1, 32, 0.37f, 0.61f, 512, 148 * 64}, {"tld4_pair_l2", 1, 256, 1.9f, 5.1f, 256, 148 * 64},
        {"tld4_pair_hbm_lattice", 1, 1024, 2.0f, 5.12f, 64, 4096 * 4}};
    if (json) printf("{");
    for (unsigned i = 0; i < sizeof(cfg) / sizeof(cfg[0]); ++i) {
        const double g = run(cfg[i].mode, cfg[i].N, cfg[i].spread, cfg[i].step, cfg[i].steps, cfg[i].blocks);
        if (json) printf("%s\"%s_gsamples_per_s\": %.1f", i ? ", " : "", cfg[i].name, g);
        else printf("%-28s N=%4d: %8.1f Gsamples/s = %8.1f GB/s at 32 B per sample\n", cfg[i].name, cfg[i].N, g, g * 32);
    }
    if (json) printf("}\n");
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
