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
        // chunk_off has one entry per VRDD_ERR_CHUNK (32) voxels; a tile spans 16 of them
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

        // Uniform trip count + predicated body: with `k < ne` as the loop bound, lanes leave the loop
        // at different iterations and ptxas does not reconverge them before the statistics below,
        // which then run once per distinct NE with a sliver of the warp (ncu: 8.8x the instructions).
        const int ne_warp = __reduce_max_sync(0xffffffffu, ne);
        for (int k = 0; k < ne_warp; ++k) {
            if (k < ne) {
                const float2 e = staged ? estage[my_off + k] : errs[tile_base + my_off + k];
                const int bin = (int)e.x;
                if (bin >= 0 && bin < VRDD_BINS) {
                    float x = cur[bin * kThreads + tid] + e.y;
                    cur[bin * kThreads + tid] = (x < 0.f) ? 0.f : x;
                }
            }
        }
        __syncwarp();

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

// ---- "moments" variant: O(NE) per voxel instead of O(B) ------------------------------------
//
// Everything d_basicDataProcessing computes from the reconstructed histogram is a function of
// four sums over its bins m:  A0 = sum cur[m],  A1 = sum m cur[m],  A2 = sum m^2 cur[m],
// AH = sum cur[m] log2 cur[m]:
//     tot = A0;  mean1 = bw*A1/tot + bw/2;  variance1 = bw^2 (A2/tot - (A1/tot)^2);
//     entropy1 = -(AH/tot - log2 tot) / log2(32)
// (volumeRender_kernel.cu:828-867 with p = cur/tot).  For cur = shift(flip(template)) those
// sums follow from per-template prefix moments — a circular shift by s moves every bin by s
// except the last s source bins, which wrap by -32; a flip maps j -> 31-j — and AH / A0 do not
// depend on the permutation at all.  Each of the NE sparse errors then changes ONE bin from
// `old` (a template entry) to `new = max(old + val, 0)`, which updates the four sums by
// (new-old), m(new-old), m^2(new-old), g(new)-g(old).  The per-bin values are the same floats
// the reference forms; only the order of the (double precision) summation differs, so results
// agree with the oracle to rounding.  A voxel whose errors hit the same bin twice (the
// reference applies them in order, with a clamp in between) takes the dense route in-thread.
//
// Per voxel: 16 B codebook + 8*NE B errors + 12 B out from/to HBM; ~5 table words from L2/L1.
constexpr int kMomStride = 72;          // doubles per template: pre0[33], pre1[33], P2, PH, pad
constexpr int kMomThreads = 256;
constexpr int kMomWarps = kMomThreads / 32;
constexpr int kMomErrCap = 256;         // staged error entries per warp (2 KB)

__global__ void build_moments_kernel(const float* __restrict__ tmpl, int T, double* __restrict__ mom) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const float* r = tmpl + (size_t)t * VRDD_BINS;
    double* m = mom + (size_t)t * kMomStride;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, ah = 0.0;
    for (int j = 0; j < VRDD_BINS; ++j) {
        m[j] = a0; m[33 + j] = a1;
        const double v = (double)r[j];
        a0 += v; a1 += j * v; a2 += (double)(j * j) * v;
        if (v > 0.0) ah += v * log2(v);
    }
    m[32] = a0; m[65] = a1; m[66] = a2; m[67] = ah;
    m[68] = m[69] = m[70] = m[71] = 0.0;
}

__device__ __forceinline__ float xlog2x(float x) { return x * fast_log2(fmaxf(x, 1.0e-37f)); }

// Warps are independent: a warp owns 32 consecutive voxels (one entry of the error-offset
// table), scans their NE with shuffles, stages its own errors in its own slice of shared
// memory and never meets a CTA-wide barrier.  The sparse corrections are accumulated in fp32
// (they are small against the template sums); only the final combination is fp64.
__global__ void __launch_bounds__(kMomThreads)
decode_fractal_moments_kernel(const int4* __restrict__ codebook, const float2* __restrict__ errs,
                              const unsigned long long* __restrict__ chunk_off, const float* __restrict__ tmpl,
                              const double* __restrict__ mom, int T, long long nvox, DecodeOut out) {
    __shared__ float2 estage_all[kMomWarps][kMomErrCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2* estage = estage_all[warp];
    const double bw = (double)(VRDD_MAX_HISTOGRAM / (float)VRDD_BINS);
    const long long nwt = (nvox + 31) / 32;
    const long long wstride = (long long)gridDim.x * kMomWarps;

    for (long long wt = (long long)blockIdx.x * kMomWarps + warp; wt < nwt; wt += wstride) {
        const long long v = wt * 32 + lane;
        const bool live = v < nvox;
        int4 code = make_int4(0, 0, 0, 0);
        if (live) code = ldg_stream_i4(codebook + v);
        const unsigned long long base = chunk_off[wt];          // same address in every lane: one broadcast load
        const int id = min(max(code.x, 0), T - 1);
        const int s = code.y & (VRDD_BINS - 1);
        const bool flip = code.z != 0;
        const int ne = live ? min(max(code.w, 0), VRDD_BINS) : 0;

        // sums of the permuted template from its prefix moments
        const double* m = mom + (size_t)id * kMomStride;
        const double P0 = m[32], P1 = m[65], P2 = m[66], PH = m[67];
        const double q0 = flip ? m[s] : m[32 - s], q1 = flip ? m[33 + s] : m[65 - s];

        int incl = ne;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += n;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int my_off = incl - ne;
        const bool staged = total <= kMomErrCap;
        __syncwarp();                                           // previous tile's readers are done
        if (staged)
            for (int i = lane; i < total; i += 32) estage[i] = errs[base + i];
        __syncwarp();

        // sparse corrections: d0 = sum (new-old), d1 = sum bin (new-old), d2 = sum bin^2 (new-old),
        // dh = sum g(new)-g(old); uniform trip count keeps the warp converged for the tail
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, dh = 0.f;
        unsigned touched = 0u;
        bool dup = false;
        const float* row = tmpl + (size_t)id * VRDD_BINS;
        const float2* my_err = staged ? (estage + my_off) : (errs + base + my_off);
        const int ne_warp = __reduce_max_sync(0xffffffffu, ne);
        for (int k = 0; k < ne_warp; ++k) {
            if (k < ne) {
                const float2 e = my_err[k];
                const int bin = (int)e.x;
                if (bin >= 0 && bin < VRDD_BINS) {
                    dup = dup || ((touched >> bin) & 1u);
                    touched |= 1u << bin;
                    const int si = (bin - s) & (VRDD_BINS - 1);
                    const float oldv = __ldg(row + (flip ? VRDD_BINS - 1 - si : si));
                    const float newv = fmaxf(oldv + e.y, 0.f);
                    const float d = newv - oldv, fb = (float)bin;
                    d0 += d; d1 = fmaf(fb, d, d1); d2 = fmaf(fb * fb, d, d2);
                    dh += xlog2x(newv) - xlog2x(oldv);
                }
            }
        }
        __syncwarp();

        float mean_n = 0.f, var_n = 0.f, ent_n = 0.f;
        if (!dup) {
            double S1, S2, U0, U1;                   // moments of src[], and of its last s bins
            if (!flip) { S1 = P1; S2 = P2; U0 = P0 - q0; U1 = P1 - q1; }
            else { S1 = 31.0 * P0 - P1; S2 = 961.0 * P0 - 62.0 * P1 + P2; U0 = q0; U1 = 31.0 * q0 - q1; }
            const double sd = (double)s;
            const double A0 = P0 + (double)d0;
            const double A1 = S1 + sd * P0 - 32.0 * U0 + (double)d1;
            const double A2 = S2 + 2.0 * sd * S1 + sd * sd * P0 - 64.0 * (U1 + sd * U0) + 1024.0 * U0 + (double)d2;
            const double AH = PH + (double)dh;
            if (A0 > 0.0) {
                double inv = (double)(1.0f / (float)A0);       // fp32 seed + one Newton step: ~1e-14
                inv = inv * (2.0 - A0 * inv);
                const double mi = A1 * inv;
                mean_n = (float)((bw * mi + 0.5 * bw) * (1.0 / VRDD_MEAN_NORM));
                var_n = (float)(bw * bw * fmax(A2 * inv - mi * mi, 0.0) * (1.0 / VRDD_VAR_NORM));
                ent_n = (float)(-(AH * inv - (double)__log2f((float)A0)) * 0.2);
            }                                         // else: all-zero histogram stays unnormalised (:833)
        } else {
            // a bin hit twice: do what the reference does, in order, on a private histogram
            float cur[VRDD_BINS];
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                const int si = (mm - s) & (VRDD_BINS - 1);
                cur[mm] = __ldg(row + (flip ? VRDD_BINS - 1 - si : si));
            }
            for (int k = 0; k < ne; ++k) {
                const float2 e = my_err[k];
                const int bin = (int)e.x;
                if (bin < 0 || bin >= VRDD_BINS) continue;
                const float x = cur[bin] + e.y;
                cur[bin] = (x < 0.f) ? 0.f : x;
            }
            float tot = 0.f;
            for (int mm = 0; mm < VRDD_BINS; ++mm) tot += cur[mm];
            const float inv = (tot > 0.f) ? 1.0f / tot : 1.0f;
            const float bwf = (float)bw;
            float mean_raw = 0.f, E = 0.f, var = 0.f;
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                cur[mm] *= inv;
                mean_raw = fmaf(cur[mm], fmaf(bwf, (float)mm, 0.5f * bwf), mean_raw);
                E += plog2p(cur[mm]);
            }
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                const float dd = fmaf(bwf, (float)mm, 0.5f * bwf) - mean_raw;
                var = fmaf(cur[mm] * dd, dd, var);
            }
            mean_n = mean_raw * (float)(1.0 / VRDD_MEAN_NORM);
            var_n = var * (float)(1.0 / VRDD_VAR_NORM);
            ent_n = -E * 0.2f;
        }
        __syncwarp();
        if (live) emit_decoded(out, v, mean_n, var_n, ent_n);
    }
}

size_t fractal_smem_bytes(int T, bool tmpl_in_smem) {
    return (size_t)VRDD_BINS * kThreads * sizeof(float) + (size_t)kErrCap * sizeof(float2) +
           (kThreads / 32) * sizeof(int) + (tmpl_in_smem ? (size_t)T * VRDD_BINS * sizeof(float) : 0);
}

}  // namespace

int build_template_moments(vrdd_context* c, const float* d_tmpl, int T) {
    if (c->tmpl_mom) { cudaFree(c->tmpl_mom); c->tmpl_mom = nullptr; }
    VRDD_CUDA(c, cudaMalloc(&c->tmpl_mom, sizeof(double) * (size_t)T * kMomStride));
    build_moments_kernel<<<(T + 127) / 128, 128, 0, c->stream>>>(d_tmpl, T, c->tmpl_mom);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_decode_fractal(vrdd_context* c, const int32_t* cb, const float* errs, const uint64_t* off,
                          const float* tmpl, int T, long long nvox, const DecodeOut& out, float* d_recon) {
    if (nvox <= 0) return VRDD_OK;
    if (T <= 0) return fail(c, VRDD_ERR_INVALID, "decode_fractal: no templates");
    if (c->var_fractal == 1 && d_recon == nullptr && c->tmpl_mom != nullptr) {
        const long long nblk = (nvox + kMomThreads - 1) / kMomThreads;
        const long long cap = (long long)c->num_sms * 8;
        const int grid = (int)((nblk < cap) ? nblk : cap);
        decode_fractal_moments_kernel<<<grid, kMomThreads, 0, c->stream>>>(
            reinterpret_cast<const int4*>(cb), reinterpret_cast<const float2*>(errs),
            reinterpret_cast<const unsigned long long*>(off), tmpl, c->tmpl_mom, T, nvox, out);
        c->launches += 1;
        VRDD_CUDA(c, cudaGetLastError());
        return VRDD_OK;
    }
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
