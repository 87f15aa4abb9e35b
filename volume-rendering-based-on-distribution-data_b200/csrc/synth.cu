// synth.cu — device-side generation of the seeded synthetic inputs (include/vrdd_synth.h).
//
// Benchmark volumes (a 1024^3 histogram volume is 137 GB) cannot be uploaded from the
// host, so they are generated where they are consumed.  The generator uses only integer
// hashing and explicitly rounded fp32 operations, so these kernels write exactly the bits
// the host generator (and therefore the oracle) sees; tests/test_synth.py checks that.
// Not part of the decode / ray-cast algorithm and never timed.
#include "common.cuh"
#include "../../include/vrdd_synth.h"

#include <vector>

namespace vrdd {

namespace {

__global__ void synth_hist_kernel(uint32_t seed, int W, int H, int D, int z0, long long nvox, float4* out) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const long long wh = (long long)W * H;
    const int z = (int)(v / wh);
    const int r = (int)(v - (long long)z * wh);
    const int y = r / W, x = r - y * W;
    float p[VRDD_BINS];
    vrdd_synth_histogram(seed, x, y, z0 + z, W, H, D, VRDD_BINS, p);
#pragma unroll
    for (int j = 0; j < VRDD_BINS / 4; ++j)
        out[v * (VRDD_BINS / 4) + j] = make_float4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

// brick of a larger volume: local voxel (x,y,z) is global voxel (ox+x, oy+y, oz+z)
__global__ void synth_hist_region_kernel(uint32_t seed, int gw, int gh, int gd, int ox, int oy, int oz, int W, int H,
                                         int z0, long long nvox, float4* out) {
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvox) return;
    const long long wh = (long long)W * H;
    const int z = (int)(v / wh);
    const int r = (int)(v - (long long)z * wh);
    const int y = r / W, x = r - y * W;
    float p[VRDD_BINS];
    vrdd_synth_histogram(seed, ox + x, oy + y, oz + z0 + z, gw, gh, gd, VRDD_BINS, p);
#pragma unroll
    for (int j = 0; j < VRDD_BINS / 4; ++j)
        out[v * (VRDD_BINS / 4) + j] = make_float4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
}

__global__ void synth_templates_kernel(uint32_t seed, int T, float* out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    float p[VRDD_BINS];
    vrdd_synth_template(seed, k, T, VRDD_BINS, p);
    for (int i = 0; i < VRDD_BINS; ++i) out[(size_t)k * VRDD_BINS + i] = p[i];
}

// pass 1: codebook + number of errors per 32-voxel chunk (one warp)
constexpr int kSynthThreads = 256;
__global__ void __launch_bounds__(kSynthThreads)
synth_fractal_codes_kernel(uint32_t seed, int W, int H, int D, int T, int max_ne, int z0, long long nvox,
                           int4* codebook, unsigned int* chunk_ne) {
    const long long v = (long long)blockIdx.x * kSynthThreads + threadIdx.x;
    int ne = 0;
    if (v < nvox) {
        const long long wh = (long long)W * H;
        const int z = (int)(v / wh);
        const int r = (int)(v - (long long)z * wh);
        const int y = r / W, x = r - y * W;
        int code[4], eb[VRDD_BINS];
        float ev[VRDD_BINS];
        vrdd_synth_fractal_code(seed, x, y, z0 + z, W, H, D, VRDD_BINS, T, max_ne, code, eb, ev);
        codebook[v] = make_int4(code[0], code[1], code[2], code[3]);
        ne = code[3];
    }
    int s = ne;
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const long long chunk = v / VRDD_ERR_CHUNK;
    if ((threadIdx.x & 31) == 0 && chunk * VRDD_ERR_CHUNK < nvox) chunk_ne[chunk] = (unsigned int)s;
}

// pass 2: errors in the compact form: per 32-voxel chunk, round k = the k-th error of every voxel with NE > k
__global__ void __launch_bounds__(kSynthThreads)
synth_fractal_errors_kernel(uint32_t seed, int W, int H, int D, int T, int max_ne, int z0, long long nvox,
                            const unsigned long long* chunk_off, vrdd_error_entry* errs) {
    const long long v = (long long)blockIdx.x * kSynthThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int code[4] = {0, 0, 0, 0}, eb[VRDD_BINS];
    float ev[VRDD_BINS];
    if (v < nvox) {
        const long long wh = (long long)W * H;
        const int z = (int)(v / wh);
        const int r = (int)(v - (long long)z * wh);
        const int y = r / W, x = r - y * W;
        vrdd_synth_fractal_code(seed, x, y, z0 + z, W, H, D, VRDD_BINS, T, max_ne, code, eb, ev);
    }
    const int ne = code[3];
    if (v - lane >= nvox) return;
    unsigned long long pos = chunk_off[v / VRDD_ERR_CHUNK];
    const unsigned lt = (1u << lane) - 1u;
    for (int k = 0;; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, k < ne);
        if (m == 0u) break;
        if (k < ne) {
            vrdd_error_entry e;
            e.bin = eb[k]; e.value = ev[k];
            errs[pos + __popc(m & lt)] = e;
        }
        pos += __popc(m);
    }
}

}  // namespace

int launch_synth_hist(vrdd_context* c, uint32_t seed, int z0, int nz, float* d_hist) {
    const long long nvox = (long long)c->W * c->H * nz;
    if (nvox <= 0) return VRDD_OK;
    const long long grid = (nvox + 127) / 128;
    synth_hist_kernel<<<(unsigned)grid, 128, 0, c->stream>>>(seed, c->W, c->H, c->D, z0, nvox,
                                                              reinterpret_cast<float4*>(d_hist));
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_synth_hist_region(vrdd_context* c, uint32_t seed, int gw, int gh, int gd, int ox, int oy, int oz, int z0,
                             int nz, float* d_hist) {
    const long long nvox = (long long)c->W * c->H * nz;
    if (nvox <= 0) return VRDD_OK;
    synth_hist_region_kernel<<<(unsigned)((nvox + 127) / 128), 128, 0, c->stream>>>(
        seed, gw, gh, gd, ox, oy, oz, c->W, c->H, z0, nvox, reinterpret_cast<float4*>(d_hist));
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_synth_fractal(vrdd_context* c, uint32_t seed, int T, int max_ne, int z0, int nz, int32_t* d_cb,
                         vrdd_error_entry* d_err, uint64_t* d_off, float* d_tmpl, uint64_t* total_ne) {
    const long long nvox = (long long)c->W * c->H * nz;
    if (nvox <= 0) return VRDD_OK;
    const long long nchunks = (nvox + VRDD_ERR_CHUNK - 1) / VRDD_ERR_CHUNK;
    synth_templates_kernel<<<(T + 127) / 128, 128, 0, c->stream>>>(seed, T, d_tmpl);
    unsigned int* d_chunk_ne = nullptr;
    VRDD_CUDA(c, cudaMalloc(&d_chunk_ne, sizeof(unsigned int) * nchunks));
    const unsigned nblk = (unsigned)((nvox + kSynthThreads - 1) / kSynthThreads);
    synth_fractal_codes_kernel<<<nblk, kSynthThreads, 0, c->stream>>>(
        seed, c->W, c->H, c->D, T, max_ne, z0, nvox, reinterpret_cast<int4*>(d_cb), d_chunk_ne);
    c->launches += 2;
    // exclusive scan of the per-chunk counts on the host (set-up path, not timed)
    std::vector<unsigned int> cnt(nchunks);
    cudaError_t e = cudaMemcpyAsync(cnt.data(), d_chunk_ne, sizeof(unsigned int) * nchunks, cudaMemcpyDeviceToHost,
                                    c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_chunk_ne);
    if (e != cudaSuccess) return fail_cuda(c, e, "synth_fractal: chunk counts");
    std::vector<uint64_t> off(nchunks + 1);
    uint64_t run = 0;
    for (long long i = 0; i < nchunks; ++i) { off[i] = run; run += cnt[i]; }
    off[nchunks] = run;
    if (total_ne) *total_ne = run;
    VRDD_CUDA(c, cudaMemcpyAsync(d_off, off.data(), sizeof(uint64_t) * (nchunks + 1), cudaMemcpyHostToDevice,
                                 c->stream));
    synth_fractal_errors_kernel<<<nblk, kSynthThreads, 0, c->stream>>>(
        seed, c->W, c->H, c->D, T, max_ne, z0, nvox, reinterpret_cast<const unsigned long long*>(d_off),
        d_err);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));      // `off` must outlive the copy
    return VRDD_OK;
}

}  // namespace vrdd
