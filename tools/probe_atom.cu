// probe_atom.cu — which texels of a CUDA array arrive together from DRAM.
//
// Each thread point-fetches ONE texel pair: a base texel at a random position aligned to (32, 8, 1) texels and the
// texel at base + (ox, oy, oz).  The array (4.3 GB of fp32) is far larger than L2 and the 4 M bases are distinct, so
// nothing is re-used: under `ncu --metrics dram__bytes_read.sum` the bytes per pair are one DRAM access unit when
// both texels share it and two otherwise.  Run for a 3-D array and for a layered 2-D array of the same extents.
// This is what decides whether a sector-aligned copy of the sampled plane can lower the ray caster's DRAM traffic
// (DESIGN.md §4).  Test/measurement infrastructure, not product code.
//     probe_atom <3d|layered>       (one launch per offset; the offsets are printed in launch order)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t h) { h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h; }

template <bool LAYERED>
__global__ void probe(cudaTextureObject_t tex, int n, int N, int ox, int oy, int oz, uint32_t salt, float* out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    // distinct aligned bases: a bijective scramble of tid over the (N/32) x (N/8) x N grid of aligned positions
    const uint32_t cells = (uint32_t)(N / 32) * (N / 8) * N;
    const uint32_t c = (uint32_t)(((uint64_t)mix(tid * 2654435761u + salt) * cells) >> 32);
    const int bx = (c % (N / 32)) * 32, by = ((c / (N / 32)) % (N / 8)) * 8, bz = c / ((N / 32) * (N / 8));
    float a, b;
    if (LAYERED) {
        a = tex2DLayered<float>(tex, bx + 0.5f, by + 0.5f, bz);
        b = tex2DLayered<float>(tex, bx + ox + 0.5f, by + oy + 0.5f, min(bz + oz, N - 1));
    } else {
        a = tex3D<float>(tex, bx + 0.5f, by + 0.5f, bz + 0.5f);
        b = tex3D<float>(tex, bx + ox + 0.5f, by + oy + 0.5f, bz + oz + 0.5f);
    }
    if (a + b == 1234.5f) out[0] = a;
}

int main(int argc, char** argv) {
    const bool layered = argc > 1 && !std::strcmp(argv[1], "layered");
    const int N = 1024, n = 1 << 22;
    cudaArray_t arr;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    if (cudaMalloc3DArray(&arr, &desc, make_cudaExtent(N, N, N), layered ? cudaArrayLayered : 0) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaResourceDesc rd; std::memset(&rd, 0, sizeof(rd)); rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td; std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    float* out; cudaMalloc(&out, 4);
    const int offs[][3] = {{0, 0, 0}, {1, 0, 0}, {2, 0, 0}, {3, 0, 0}, {4, 0, 0}, {7, 0, 0}, {8, 0, 0}, {15, 0, 0}, {16, 0, 0}, {31, 0, 0},
                           {0, 1, 0}, {0, 2, 0}, {0, 3, 0}, {0, 4, 0}, {0, 7, 0}, {0, 0, 1}, {1, 1, 0}, {3, 1, 0}, {7, 1, 0}, {15, 1, 0}, {4, 2, 0}, {8, 2, 0}};
    const int nk = sizeof(offs) / sizeof(offs[0]);
    for (int k = 0; k < nk; ++k) {
        if (layered) probe<true><<<n / 256, 256>>>(tex, n, N, offs[k][0], offs[k][1], offs[k][2], 1000u + k, out);
        else probe<false><<<n / 256, 256>>>(tex, n, N, offs[k][0], offs[k][1], offs[k][2], 1000u + k, out);
        cudaDeviceSynchronize();
        printf("launch %2d: %s offset (%2d,%2d,%2d)  pairs %d\n", k, layered ? "layered" : "3d", offs[k][0], offs[k][1], offs[k][2], n);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
