// decode_fractal.cu — P1b: fractal code -> reconstructed histogram -> (mean, variance, entropy).
//
// Replaces the fractal half of d_basicDataProcessing
// (/root/reference/volumeRender_kernel.cu:775-871) and fractalDecoding (:195-222).
//
// Per voxel:  (id, shift, flip, NE) = codebook[v]                               (:777-786)
//             src = flip ? reverse(template[id]) : template[id]                 (:200-220)
//             cur[(i + shift) mod B] = src[i]
//             for k < NE: cur[bin_k] += val_k; cur[bin_k] = max(cur[bin_k], 0)  (:806-825)
//             tot = sum cur; if tot > 0: cur /= tot                             (:828-839)
//             mean/variance about the bin CENTRE, entropy                       (:841-867)
// The integer part (template row, permutation, which bin each error lands in, the order of
// the additions) is reproduced exactly; vrdd_reconstruct_fractal_device exposes `cur`
// before normalisation so tests can compare it bit for bit.  Statistics are fp32 with
// MUFU.LG2 (see decode_hist.cu for why not fp64).
//
// Inputs are the compact form: 16 B codebook entry + 8*NE B of errors per voxel (the
// reference's dense float2[V][32] error table is 256 B per voxel, mostly unused), with an
// error-offset entry every 256 voxels; offsets inside a tile come from a block scan of NE.
//
// Kernel shape: persistent CTAs of 512 threads, one voxel per thread, 512 voxels per
// iteration.  The template table (T x 32 floats, 79.6 KB at T = 622) is loaded into shared
// memory once per CTA; a tile's errors are staged into shared memory with coalesced loads;
// each thread's 32-bin working histogram lives in shared memory as cur[bin][thread], which
// is bank-conflict-free for the data-dependent error updates.
#include "common.cuh"

namespace vrdd {

namespace {

constexpr int kThreads = 512;
constexpr int kErrCap = 4096;          // staged error entries per tile (32 KB)

__global__ void __launch_bounds__(kThreads, 1)
decode_fractal_dense_kernel(const int4* __restrict__ codebook, const float2* __restrict__ errs,
                            const unsigned long long* __restrict__ chunk_off,
                            const float* __restrict__ tmpl_g, int T, int tmpl_in_smem, long long nvox,
                            DecodeOut out, float* __restrict__ recon) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* cur = reinterpret_cast<float*>(smem_raw);                         // [32][kThreads]
    float2* estage = reinterpret_cast<float2*>(cur + VRDD_BINS * kThreads);  // [kErrCap]
    int* wsum = reinterpret_cast<int*>(estage + kErrCap);                    // [kThreads/32]
    float* tmpl_s = reinterpret_cast<float*>(wsum + kThreads / 32);          // [T*32] if tmpl_in_smem

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tmpl_in_smem) {
        const float4* src = reinterpret_cast<const float4*>(tmpl_g);
        float4* dst = reinterpret_cast<float4*>(tmpl_s);
        for (int i = tid; i < T * (VRDD_BINS / 4); i += kThreads) dst[i] = src[i];
    }
    const float* tmpl = tmpl_in_smem ? tmpl_s : tmpl_g;
    __syncthreads();

    const float bw = VRDD_MAX_HISTOGRAM / (float)VRDD_BINS;
    const float hb = 0.5f * bw;
    const long long ntiles = (nvox + kThreads - 1) / kThreads;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long v = tile * kThreads + tid;
        const bool live = v < nvox;
        int4 code = make_int4(0, 0, 0, 0);
        if (live) code = ldg_stream_i4(codebook + v);
        // the reference only prints when a code is out of range (:781-789); here it is made safe
        int id = min(max(code.x, 0), T - 1);
        const int shift = code.y & (VRDD_BINS - 1);       // shift == B is the identity, like the single wrap
        const int flip = code.z;
        const int ne = live ? min(max(code.w, 0), VRDD_BINS) : 0;

        // exclusive scan of NE over the tile -> this thread's first error
        int incl = ne;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wbase = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            const int s = wsum[w];
            if (w < warp) wbase += s;
            tile_total += s;
        }
        const int my_off = wbase + incl - ne;
        // chunk_off has one entry per 256 voxels; a tile spans two chunks
        const unsigned long long tile_base = chunk_off[tile * (kThreads / VRDD_ERR_CHUNK)];
        const bool staged = tile_total <= kErrCap;
        if (staged)
            for (int i = tid; i < tile_total; i += kThreads) estage[i] = errs[tile_base + i];

        // template row, flipped and rotated, into this thread's column of cur[][]
        const float* row = tmpl + (size_t)id * VRDD_BINS;
#pragma unroll
        for (int m = 0; m < VRDD_BINS; ++m) {
            int si = (m - shift) & (VRDD_BINS - 1);
            if (flip) si = VRDD_BINS - 1 - si;
            cur[m * kThreads + tid] = row[si];
        }
        __syncthreads();                                   // estage complete (cur is thread-private)

        for (int k = 0; k < ne; ++k) {
            const float2 e = staged ? estage[my_off + k] : errs[tile_base + my_off + k];
            const int bin = (int)e.x;
            if (bin >= 0 && bin < VRDD_BINS) {
                float x = cur[bin * kThreads + tid] + e.y;
                cur[bin * kThreads + tid] = (x < 0.f) ? 0.f : x;
            }
        }

        float p[VRDD_BINS];
        float tot = 0.f;
#pragma unroll
        for (int m = 0; m < VRDD_BINS; ++m) {
            p[m] = cur[m * kThreads + tid];
            tot += p[m];                                   // sequential, like :829-831
        }
        if (recon != nullptr && live) {
#pragma unroll
            for (int m = 0; m < VRDD_BINS; ++m) recon[v * VRDD_BINS + m] = p[m];
        }
        if (live) {
            const float inv = (tot > 0.f) ? (1.0f / tot) : 1.0f;
            float mean_raw = 0.f, E0 = 0.f, E1 = 0.f;
#pragma unroll
            for (int m = 0; m < VRDD_BINS; ++m) {
                p[m] *= inv;
                mean_raw = fmaf(p[m], fmaf(bw, (float)m, hb), mean_raw);
                if (m & 1) E1 += plog2p(p[m]); else E0 += plog2p(p[m]);
            }
            float va = 0.f, vb = 0.f;
#pragma unroll
            for (int m = 0; m < VRDD_BINS; ++m) {
                const float d = fmaf(bw, (float)m, hb) - mean_raw;          // bin centre (:851-853)
                if (m & 1) vb = fmaf(p[m] * d, d, vb); else va = fmaf(p[m] * d, d, va);
            }
            emit_decoded(out, v, mean_raw * (float)(1.0 / VRDD_MEAN_NORM), (va + vb) * (float)(1.0 / VRDD_VAR_NORM),
                         -(E0 + E1) * (1.0f / 5.0f));
        }
        __syncthreads();                                   // estage / wsum reused by the next tile
    }
}

size_t fractal_smem_bytes(int T, bool tmpl_in_smem) {
    return (size_t)VRDD_BINS * kThreads * sizeof(float) + (size_t)kErrCap * sizeof(float2) +
           (kThreads / 32) * sizeof(int) + (tmpl_in_smem ? (size_t)T * VRDD_BINS * sizeof(float) : 0);
}

}  // namespace

int launch_decode_fractal(vrdd_context* c, const int32_t* cb, const float* errs, const uint64_t* off,
                          const float* tmpl, int T, long long nvox, const DecodeOut& out, float* d_recon) {
    if (nvox <= 0) return VRDD_OK;
    if (T <= 0) return fail(c, VRDD_ERR_INVALID, "decode_fractal: no templates");
    bool in_smem = fractal_smem_bytes(T, true) <= 227 * 1024;
    const size_t smem = fractal_smem_bytes(T, in_smem);
    VRDD_CUDA(c, cudaFuncSetAttribute(decode_fractal_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const long long ntiles = (nvox + kThreads - 1) / kThreads;
    const int grid = (int)((ntiles < c->num_sms) ? ntiles : c->num_sms);
    decode_fractal_dense_kernel<<<grid, kThreads, smem, c->stream>>>(
        reinterpret_cast<const int4*>(cb), reinterpret_cast<const float2*>(errs),
        reinterpret_cast<const unsigned long long*>(off), tmpl, T, in_smem ? 1 : 0, nvox, out, d_recon);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

}  // namespace vrdd
