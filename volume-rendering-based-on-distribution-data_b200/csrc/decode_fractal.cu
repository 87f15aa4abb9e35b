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
// Inputs are the compact form (include/vrdd.h, vrdd_set_fractal_device): 16 B codebook entry +
// 8*NE B of errors per voxel (the reference's dense float2[V][32] error table is 256 B per
// voxel, mostly unused).  Errors are grouped per chunk of 32 consecutive voxels — one warp —
// and stored ROUND-MAJOR inside a chunk: first the first error of every voxel that has one (in
// voxel order), then the second errors, ...  In round k the lanes with NE > k therefore read
// consecutive entries: one coalesced load per round straight from global memory, no staging,
// no prefix scan; a lane's slot in the round is its rank in the ballot of (k < NE).
#include "common.cuh"

namespace vrdd {

namespace {

struct ErrEntry { int bin; float val; };     // vrdd_error_entry

__device__ __forceinline__ unsigned long long ldg_stream_u64(const unsigned long long* p) {
    unsigned long long r;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ ErrEntry ldg_stream_err(const ErrEntry* p) {
    ErrEntry r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.bin), "=f"(r.val) : "l"(p));
    return r;
}

constexpr int kThreads = 512;

// ---- "dense" variant: the reference's order of operations on a private 32-bin histogram -----
// Persistent CTAs of 512 threads, one voxel per thread; the template table (T x 32 floats,
// 79.6 KB at T = 622) sits in shared memory when it fits; each thread's working histogram lives
// in shared memory as cur[bin][thread], bank-conflict-free for the data-dependent updates.
__global__ void __launch_bounds__(kThreads, 1)
decode_fractal_dense_kernel(const int4* __restrict__ codebook, const ErrEntry* __restrict__ errs,
                            const unsigned long long* __restrict__ chunk_off,
                            const float* __restrict__ tmpl_g, int T, int tmpl_in_smem, long long nvox,
                            DecodeOut out, float* __restrict__ recon) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* cur = reinterpret_cast<float*>(smem_raw);                         // [32][kThreads]
    float* tmpl_s = cur + VRDD_BINS * kThreads;                              // [T*32] if tmpl_in_smem

    const int tid = threadIdx.x;
    const unsigned lt = (1u << (tid & 31)) - 1u;

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

        // template row, flipped and rotated, into this thread's column of cur[][]
        const float* row = tmpl + (size_t)id * VRDD_BINS;
#pragma unroll
        for (int m = 0; m < VRDD_BINS; ++m) {
            int si = (m - shift) & (VRDD_BINS - 1);
            if (flip) si = VRDD_BINS - 1 - si;
            cur[m * kThreads + tid] = row[si];
        }

        // rounds of the warp's chunk; warp-uniform trip count (the ballot) keeps the lanes converged
        unsigned long long pos = (v - (tid & 31) < nvox) ? chunk_off[v >> 5] : 0ull;
        for (int k = 0;; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, k < ne);
            if (m == 0u) break;
            if (k < ne) {
                const ErrEntry e = errs[pos + __popc(m & lt)];
                if ((unsigned)e.bin < (unsigned)VRDD_BINS) {
                    float x = cur[e.bin * kThreads + tid] + e.val;
                    cur[e.bin * kThreads + tid] = (x < 0.f) ? 0.f : x;
                }
            }
            pos += __popc(m);
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
    }
}

// ---- "moments" variant: O(NE) per voxel instead of O(B) ------------------------------------
//
// Everything d_basicDataProcessing computes from the reconstructed histogram cur[] is a function of four
// sums over its bins m, taken about any reference point c:
//     A0 = sum cur[m],  B1 = sum (m-c) cur[m],  B2 = sum (m-c)^2 cur[m],  AH = sum cur[m] log2 cur[m]
//     tot = A0;  mean1 = bw (c + B1/tot) + bw/2;  variance1 = bw^2 (B2/tot - (B1/tot)^2);
//     entropy1 = -(AH/tot - log2 tot) / log2(32)
// (volumeRender_kernel.cu:828-867 with p = cur/tot).  For cur = shift(flip(template)) there are only
// T x 2 x 32 different histograms, so their sums are tabulated once (fp64, build_moments_kernel) with c = the
// permuted template's own mean, which makes B1 vanish and B2 its central second moment: one 16-byte entry
// {c, B2, A0, AH} per (template, flip, shift).  Each of the NE sparse errors then changes ONE bin from `old`
// (a template entry) to `new = max(old + val, 0)`, which updates the sums by (new-old), (m-c)(new-old),
// (m-c)^2 (new-old), g(new)-g(old) with g(x) = x log2 x.  Because the sums are centred, the corrections and
// the final combination need no more than fp32: B1/tot is the small shift of the mean the errors cause, so
// B2/tot - (B1/tot)^2 does not cancel.  The per-bin values are the same floats the reference forms; only the
// order of the summation differs, so results agree with the oracle to rounding.  A voxel whose errors hit
// the same bin twice (the reference applies them in order, with a clamp in between) takes the dense route.
//
// Per voxel: 16 B codebook + 8*NE B errors + 12 B out from/to HBM, one 16-B table entry from L2.
//
// Table layout in vrdd_context::tmpl_mom: float4 [T][2][32] {c, B2, A0, AH} indexed (id, flip, shift), then
// float2 [T][32] {t[j], g(t[j])}.
constexpr int kMomBytesPerTemplate = 2 * VRDD_BINS * 16 + VRDD_BINS * 8;      // 1280
constexpr int kMomThreads = 256;        // global-table kernel
constexpr int kMomWarps = kMomThreads / 32;

__device__ __forceinline__ float xlog2x(float x) { return __fmul_rn(x, fast_log2(fmaxf(x, 1.0e-37f))); }

__global__ void build_moments_kernel(const float* __restrict__ tmpl, int T, float4* __restrict__ perm,
                                     float2* __restrict__ vg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;          // (t, flip, s)
    if (i >= T * 2 * VRDD_BINS) return;
    const int t = i / (2 * VRDD_BINS), s = i & (VRDD_BINS - 1);
    const int fm = ((i / VRDD_BINS) & 1) ? VRDD_BINS - 1 : 0;
    const float* r = tmpl + (size_t)t * VRDD_BINS;
    double a0 = 0.0, a1 = 0.0, ah = 0.0;
    for (int m = 0; m < VRDD_BINS; ++m) {
        const double v = (double)r[((m - s) & (VRDD_BINS - 1)) ^ fm];
        a0 += v; a1 += m * v;
        if (v > 0.0) ah += v * log2(v);
    }
    const float c = (a0 > 0.0) ? (float)(a1 / a0) : 0.f;
    double b2 = 0.0;
    for (int m = 0; m < VRDD_BINS; ++m) {
        const double dm = (double)m - (double)c;
        b2 += dm * dm * (double)r[((m - s) & (VRDD_BINS - 1)) ^ fm];
    }
    // the residual first moment a1 - c*a0 (|.| < 2e-6 a0) is dropped: it moves the mean by < 1e-7 bins
    perm[i] = make_float4(c, (float)b2, (float)a0, (float)ah);
    if (i < T * VRDD_BINS) vg[i] = make_float2(tmpl[i], xlog2x(tmpl[i]));
}

// What one lane does for its voxel (shared by both moments kernels).  `row` is the voxel's
// {value, g(value)} template row (shared or global memory), `ent` its table entry, `pos` the
// warp's chunk offset.  Trip counts come from ballots, so they are warp-uniform and the lanes
// stay converged for the tail.
template <class AfterLoads>
__device__ __forceinline__ void moments_voxel(const float2* __restrict__ row, const ErrEntry* __restrict__ errs,
                                              unsigned long long pos, int ne, int s, int fm, const float4 ent,
                                              unsigned lt, float& mean_n, float& var_n, float& ent_n,
                                              AfterLoads&& after_loads) {
    const float c = ent.x;
    float d0 = 0.f, b1 = 0.f, b2 = 0.f, dh = 0.f;
    unsigned touched = 0u, dup = 0u;
    // branch-free: an absent entry is (bin 32, value 0): it reads a valid row slot, leaves new == old, so
    // every increment below is exactly zero, and sets no bit (shl.b32 clamps the shift).  Bins are trusted to
    // be in [0, 32) (vrdd_set_fractal_host / vrdd_pack_fractal_errors reject others); whatever they hold, the
    // row index is masked, so no access leaves the table.
    auto apply = [&](int bin, float val) {
        unsigned bit;
        asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(bin));
        dup |= touched & bit;
        touched |= bit;
        const float2 og = row[((bin - s) & (VRDD_BINS - 1)) ^ fm];
        const float newv = fmaxf(og.x + val, 0.f);
        const float d = newv - og.x;
        const float fc = (__int_as_float(0x4b000000 | bin) - 8388608.f) - c;    // (float)bin - c without the conversion pipe
        d0 += d; b1 = fmaf(fc, d, b1); b2 = fmaf(fc * fc, d, b2);
        dh += __fadd_rn(xlog2x(newv), -og.y);                                   // exactly 0 when new == old
    };
    // The first kRounds rounds are loaded before any of them is used (their addresses depend only on the
    // ballots), so a tile pays one memory latency for its errors, not one per round.
    constexpr int kRounds = 8;
    ErrEntry e[kRounds];
    const char* eb = reinterpret_cast<const char*>(errs + pos);
    unsigned off = 0u;                                           // < 32 * 32 entries per chunk
    const int rounds = min(__reduce_max_sync(0xffffffffu, ne), kRounds);          // warp-uniform
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, k < ne);
        e[k].bin = VRDD_BINS; e[k].val = 0.f;
        if (k < ne) e[k] = ldg_stream_err(reinterpret_cast<const ErrEntry*>(eb + ((off + __popc(m & lt)) << 3)));
        off += __popc(m);
    }
    after_loads();
#pragma unroll
    for (int k = 0; k < kRounds; ++k) {
        if (k < rounds) apply(e[k].bin, e[k].val);
    }
    if (rounds == kRounds) {
        for (int k = kRounds;; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, k < ne);
            if (m == 0u) break;
            ErrEntry x; x.bin = VRDD_BINS; x.val = 0.f;
            if (k < ne) x = ldg_stream_err(reinterpret_cast<const ErrEntry*>(eb + ((off + __popc(m & lt)) << 3)));
            apply(x.bin, x.val);
            off += __popc(m);
        }
    }

    const float bw = VRDD_MAX_HISTOGRAM / (float)VRDD_BINS;
    mean_n = 0.f; var_n = 0.f; ent_n = 0.f;
    const float A0 = ent.z + d0;
    // Heavy cancellation: when the errors remove (nearly) all of the template's mass, A0 = table sum (fp64, rounded) +
    // sequential fp32 corrections is rounding noise — about 1e-7 where the reference's tot is 0 (it then writes
    // 0, 0, 0; 1/1e-7 here would give garbage), or a small value with too few correct digits: the noise is ~1e-7 of
    // the mass, so below 1/64 of the mass the 2e-5 tolerance of the statistics is at risk.  Such voxels take the
    // reference's in-order dense route below, like the voxels with a bin hit twice.
    if (A0 <= 0.015625f * (ent.z + fabsf(d0)) && (ent.z + fabsf(d0)) > 0.f) dup |= 0x80000000u;
    if (!dup && A0 > 0.f) {
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(A0));   // one MUFU.RCP, 1 ulp; A0 is a sum of frequencies, ~1
        const float dm = b1 * inv;                               // shift of the mean caused by the errors
        mean_n = fmaf(bw, c + dm, 0.5f * bw) * (float)(1.0 / VRDD_MEAN_NORM);
        var_n = bw * bw * fmaxf(fmaf(ent.y + b2, inv, -dm * dm), 0.f) * (float)(1.0 / VRDD_VAR_NORM);
        ent_n = -(fmaf(ent.w + dh, inv, -fast_log2(A0))) * 0.2f;
    }                                             // A0 == 0: the all-zero histogram stays unnormalised (:833)
    // A bin hit twice: do what the reference does, in order, on a private histogram.  The rounds are walked
    // again by the whole warp (the ballots need every lane); only the lanes that saw a duplicate apply them.
    if (__any_sync(0xffffffffu, dup != 0u)) {
        float cur[VRDD_BINS];
        if (dup)
            for (int mm = 0; mm < VRDD_BINS; ++mm) cur[mm] = row[((mm - s) & (VRDD_BINS - 1)) ^ fm].x;
        off = 0u;
        for (int k = 0;; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, k < ne);
            if (m == 0u) break;
            if (dup && k < ne) {
                const ErrEntry x = *reinterpret_cast<const ErrEntry*>(eb + ((off + __popc(m & lt)) << 3));
                if ((unsigned)x.bin < (unsigned)VRDD_BINS) {
                    const float y = cur[x.bin] + x.val;
                    cur[x.bin] = (y < 0.f) ? 0.f : y;
                }
            }
            off += __popc(m);
        }
        if (dup) {
            float tot = 0.f;
            for (int mm = 0; mm < VRDD_BINS; ++mm) tot += cur[mm];
            const float inv = (tot > 0.f) ? 1.0f / tot : 1.0f;
            float mean_raw = 0.f, E = 0.f, var = 0.f;
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                cur[mm] *= inv;
                mean_raw = fmaf(cur[mm], fmaf(bw, (float)mm, 0.5f * bw), mean_raw);
                E += plog2p(cur[mm]);
            }
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                const float dd = fmaf(bw, (float)mm, 0.5f * bw) - mean_raw;
                var = fmaf(cur[mm] * dd, dd, var);
            }
            mean_n = mean_raw * (float)(1.0 / VRDD_MEAN_NORM);
            var_n = var * (float)(1.0 / VRDD_VAR_NORM);
            ent_n = -E * 0.2f;
        }
        __syncwarp();
    }
}

// Tables in global memory (any T).  A warp owns 32 consecutive voxels (one entry of the error-offset table).
__global__ void __launch_bounds__(kMomThreads)
decode_fractal_moments_kernel(const int4* __restrict__ codebook, const ErrEntry* __restrict__ errs,
                              const unsigned long long* __restrict__ chunk_off,
                              const float4* __restrict__ perm, const float2* __restrict__ vg_g, int T, long long nvox,
                              DecodeOut out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
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
        const int fl = code.z != 0;
        const int ne = live ? min(max(code.w, 0), VRDD_BINS) : 0;
        const float4 ent = __ldg(perm + ((id * 2 + fl) * VRDD_BINS + s));
        float mean_n, var_n, ent_n;
        moments_voxel(vg_g + (size_t)id * VRDD_BINS, errs, base, ne, s, fl ? VRDD_BINS - 1 : 0, ent, lt, mean_n, var_n, ent_n,
                      [] {});
        if (live) emit_decoded(out, v, mean_n, var_n, ent_n);
    }
}

// The same algorithm with the {value, g(value)} rows in shared memory (T*256 B, 159 KB at T = 622): the
// data-dependent row reads were 32-way divergent global loads and made the kernel L1-tag bound (ncu,
// profiles/summary_r1d.md: l1tex 72 %, 230 tag requests per warp iteration); in shared memory they cost a few
// bank-conflict wavefronts each.  Only the one table entry per voxel stays a global (L2-resident) load.
// One persistent CTA of independent warps per SM; the next tile's codebook entry and error offset are loaded
// before the current tile is processed; the voxel coordinate advances incrementally when tiles do not
// straddle rows (W % 32 == 0) instead of being divided out per voxel.
size_t moments_smem_bytes(int T) { return (size_t)T * VRDD_BINS * sizeof(float2); }

template <int kSmThreads>
__global__ void __launch_bounds__(kSmThreads, 1)
decode_fractal_moments_smem_kernel(const int4* __restrict__ codebook, const ErrEntry* __restrict__ errs,
                                   const unsigned long long* __restrict__ chunk_off,
                                   const float4* __restrict__ perm, const float2* __restrict__ vg_g, int T,
                                   long long nvox, DecodeOut out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* vg_s = reinterpret_cast<float2*>(smem_raw);                                   // [T][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    {
        const float4* src = reinterpret_cast<const float4*>(vg_g);
        float4* dst = reinterpret_cast<float4*>(vg_s);
        for (int i = threadIdx.x; i < T * (VRDD_BINS / 2); i += kSmThreads) dst[i] = src[i];
    }
    __syncthreads();

    constexpr int kSmWarps = kSmThreads / 32;
    const long long nwt = (nvox + 31) / 32;
    const long long wstride = (long long)gridDim.x * kSmWarps;
    long long wt = (long long)blockIdx.x * kSmWarps + warp;
    if (wt >= nwt) return;

    // voxel coordinate of lane 0 of this warp's tile and of the stride between its tiles
    const bool rows32 = ((out.W & 31) == 0) && ((out.v_base & 31) == 0) && (out.use_surf || out.brick[0]);
    int x0 = 0, y0 = 0, z0 = 0, sx = 0, sy = 0, sz = 0;
    if (rows32) {
        split_voxel(out, out.v_base + wt * 32, x0, y0, z0);
        split_voxel(out, wstride * 32, sx, sy, sz);
    }

    int4 code = make_int4(0, 0, 0, 0);
    if (wt * 32 + lane < nvox) code = ldg_stream_i4(codebook + wt * 32 + lane);
    unsigned long long base = chunk_off[wt];

    while (true) {
        const long long v = wt * 32 + lane;
        const bool live = v < nvox;
        const long long wt_n = wt + wstride;
        int4 code_n = make_int4(0, 0, 0, 0);
        unsigned long long base_n = 0ull;
        const int id = min(max(code.x, 0), T - 1);
        const int s = code.y & (VRDD_BINS - 1);
        const int fl = code.z != 0;
        const int ne = live ? min(max(code.w, 0), VRDD_BINS) : 0;
        const float4 ent = __ldg(perm + ((id * 2 + fl) * VRDD_BINS + s));
        float mean_n, var_n, ent_n;
        moments_voxel(vg_s + id * VRDD_BINS, errs, base, ne, s, fl ? VRDD_BINS - 1 : 0, ent, lt, mean_n, var_n, ent_n, [&] {
            // next tile's inputs: issued behind this tile's loads, in flight while it is processed
            if (wt_n < nwt) {
                if (wt_n * 32 + lane < nvox) code_n = ldg_stream_i4(codebook + wt_n * 32 + lane);
                base_n = ldg_stream_u64(chunk_off + wt_n);
            }
        });
        if (live) {
            if (rows32) emit_decoded_xyz(out, v, x0 + lane, y0, z0, mean_n, var_n, ent_n);
            else emit_decoded(out, v, mean_n, var_n, ent_n);
        }
        if (wt_n >= nwt) break;
        wt = wt_n; code = code_n; base = base_n;
        if (rows32) {
            x0 += sx; if (x0 >= out.W) { x0 -= out.W; ++y0; }
            y0 += sy; if (y0 >= out.H) { y0 -= out.H; ++z0; }
            z0 += sz;
        }
    }
}

// ---- second generation of the shared-memory moments kernel ----------------------------------
// Same algorithm and the same numbers per voxel as moments_voxel; what changes is how the work is issued.
// The r1f capture (profiles/summary_r1f.md) shows the kernel at 75 % issue-slot and 89 % LSU-data-pipe
// utilisation, 170 of its 449 instructions per 32 voxels in the address phase of the eight rounds and 44 of
// its 60 shared-memory wavefronts in bank conflicts.  Hence:
//  * SCAN: a lane's slot in round k is (entries of rounds < k) + (lanes below it with NE > k).  The eight
//    indicator bits (NE > k) are packed one per byte into two words and ONE warp scan (5 steps, 10 SHFL)
//    yields the eight ranks together; the round offsets are the byte-wise prefix sums of the totals, a
//    multiply by 0x01010100.  Replaces 8 ballots + 16 POPC (quarter-rate pipe) + 64-bit address chains.
//  * row reads are predicated on k < NE: an absent entry used to read a (valid, random) slot, taking part
//    in the bank conflicts of the round; now it contributes {0, 0}, which still leaves every increment
//    exactly zero (new = max(0 + 0, 0) = old, g(0) - 0 = 0).
//  * RECOMP: rows hold only the template values (float, the reference's own table, 80 KB at T = 622);
//    g(old) = old*log2(old) is recomputed with the function the table was built with (identical bits), one
//    MUFU more per round against a 32-bit instead of a 64-bit shared load (half the wavefronts).
//  * a bin hit twice is detected once per voxel, popc(touched) != NE, not per round.
//  * rounds are skipped in pairs (warp-uniform), absent entries being no-ops anyway.
__device__ __forceinline__ unsigned shl_clamp(unsigned x, unsigned n) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));            // PTX shl clamps n > 31 (result 0)
    return r;
}
// inclusive warp scan step: x += (lane >= d) ? x of lane - d : 0, using the shuffle's own in-range predicate
__device__ __forceinline__ unsigned scan_step(unsigned x, int d) {
    unsigned r;
    asm("{\n\t.reg .u32 t;\n\t.reg .pred p;\n\t"
        "shfl.sync.up.b32 t|p, %1, %2, 0, 0xffffffff;\n\t"      // out of range: t = own value, p = false
        "@p add.u32 t, t, %1;\n\t"
        "mov.u32 %0, t;\n\t}"
        : "=r"(r) : "r"(x), "r"(d));
    return r;
}

template <bool RECOMP>
struct RowT { using type = float2; };          // {t, t*log2 t}
template <>
struct RowT<true> { using type = float; };     // t only: g(t) is recomputed

template <bool SCAN, bool RECOMP, class AfterLoads>
__device__ __forceinline__ void moments_voxel2(const typename RowT<RECOMP>::type* __restrict__ row,
                                               const ErrEntry* __restrict__ errs, unsigned long long pos, int ne, int s,
                                               int fm, const float4 ent, unsigned lt, float& mean_n, float& var_n,
                                               float& ent_n, AfterLoads&& after_loads) {
    const float c = ent.x;
    float d0 = 0.f, b1 = 0.f, b2 = 0.f, dh = 0.f;
    unsigned touched = 0u;
    auto row_at = [&](int bin) -> float2 {
        const int j = ((bin - s) & (VRDD_BINS - 1)) ^ fm;
        if constexpr (RECOMP) { const float t = row[j]; return make_float2(t, xlog2x(t)); }
        else return row[j];
    };
    auto apply = [&](int bin, float val) {
        touched |= shl_clamp(1u, (unsigned)bin);
        float2 og = make_float2(0.f, 0.f);
        if ((unsigned)bin < (unsigned)VRDD_BINS) og = row_at(bin);     // absent entry: bin == 32 (a short-lived predicate;
                                                                        // keeping (k < NE) alive per round made ptxas sink the loads)
        const float newv = fmaxf(og.x + val, 0.f);
        const float d = newv - og.x;
        const float fc = (__int_as_float(0x4b000000 | bin) - 8388608.f) - c;    // (float)bin - c without the conversion pipe
        d0 += d; b1 = fmaf(fc, d, b1); b2 = fmaf(fc * fc, d, b2);
        dh += __fadd_rn(xlog2x(newv), -og.y);                                   // exactly 0 when new == old
    };
    constexpr int kRounds = 8;
    ErrEntry e[kRounds];
    const char* eb = reinterpret_cast<const char*>(errs + pos);
    unsigned off = 0u;
    const int rounds = min(__reduce_max_sync(0xffffffffu, ne), kRounds);          // warp-uniform
    if constexpr (SCAN) {
        const unsigned ones = 0x01010101u;
        const unsigned lo = ones & ~shl_clamp(0xffffffffu, 8u * (unsigned)ne);                  // byte k: NE > k
        const unsigned hi = ones & ~shl_clamp(0xffffffffu, 8u * (unsigned)max(ne - 4, 0));      // byte k: NE > 4 + k
        unsigned il = lo, ih = hi;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { il = scan_step(il, d); ih = scan_step(ih, d); }
        const unsigned tl = __shfl_sync(0xffffffffu, il, 31), th = __shfl_sync(0xffffffffu, ih, 31);
        const unsigned sum_lo = (tl * ones) >> 24;                               // entries of rounds 0..3 (<= 128)
        // byte k = entries of the rounds before round k (+ lanes below with an entry in round k): < 256
        const unsigned slot_lo = tl * 0x01010100u + (il - lo);
        const unsigned slot_hi = th * 0x01010100u + sum_lo * ones + (ih - hi);
        off = sum_lo + ((th * ones) >> 24);
#pragma unroll
        for (int k = 0; k < kRounds; ++k) {
            const unsigned idx = __byte_perm(k < 4 ? slot_lo : slot_hi, 0u, 0x4440u + (k & 3));
            e[k].bin = VRDD_BINS; e[k].val = 0.f;
            if (k < ne) e[k] = ldg_stream_err(reinterpret_cast<const ErrEntry*>(eb + ((size_t)idx << 3)));
        }
    } else {
#pragma unroll
        for (int k = 0; k < kRounds; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, k < ne);
            e[k].bin = VRDD_BINS; e[k].val = 0.f;
            if (k < ne) e[k] = ldg_stream_err(reinterpret_cast<const ErrEntry*>(eb + ((off + __popc(m & lt)) << 3)));
            off += __popc(m);
        }
    }
    after_loads();
    // Pin the order "all loads, then all uses": the empty volatile statements stay behind the (volatile) loads
    // and every use below depends on them; without this ptxas sinks the later rounds' loads next to their
    // use and the tile pays a memory latency per pair of rounds (5.4 instead of 4.3 ms per 2^28 voxels).
#pragma unroll
    for (int k = 0; k < kRounds; ++k) asm volatile("" : "+r"(e[k].bin), "+f"(e[k].val));
#pragma unroll
    for (int k = 0; k < kRounds; k += 2) {
        if (k < rounds) { apply(e[k].bin, e[k].val); apply(e[k + 1].bin, e[k + 1].val); }
    }
    if (rounds == kRounds) {
        for (int k = kRounds;; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, k < ne);
            if (m == 0u) break;
            ErrEntry x; x.bin = VRDD_BINS; x.val = 0.f;
            if (k < ne) x = ldg_stream_err(reinterpret_cast<const ErrEntry*>(eb + ((off + __popc(m & lt)) << 3)));
            apply(x.bin, x.val);
            off += __popc(m);
        }
    }
    const float bw = VRDD_MAX_HISTOGRAM / (float)VRDD_BINS;
    mean_n = 0.f; var_n = 0.f; ent_n = 0.f;
    const float A0 = ent.z + d0;
    // NE distinct bins set NE bits (bins are in [0, 32), see moments_voxel); a bin hit twice, or errors that cancel
    // (nearly) all of the template's mass (A0 is then rounding noise, see moments_voxel), take the dense route
    const float mass = ent.z + fabsf(d0);
    const bool dup = (__popc(touched) != ne) || (A0 <= 0.015625f * mass && mass > 0.f);
    if (!dup && A0 > 0.f) {
        float inv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(A0));
        const float dm = b1 * inv;
        mean_n = fmaf(bw, c + dm, 0.5f * bw) * (float)(1.0 / VRDD_MEAN_NORM);
        var_n = bw * bw * fmaxf(fmaf(ent.y + b2, inv, -dm * dm), 0.f) * (float)(1.0 / VRDD_VAR_NORM);
        ent_n = -(fmaf(ent.w + dh, inv, -fast_log2(A0))) * 0.2f;
    }
    // A bin hit twice: the reference's ordered route on a private histogram, as in moments_voxel.
    if (__any_sync(0xffffffffu, dup)) {
        float cur[VRDD_BINS];
        if (dup)
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                const int j = ((mm - s) & (VRDD_BINS - 1)) ^ fm;
                if constexpr (RECOMP) cur[mm] = row[j]; else cur[mm] = row[j].x;
            }
        off = 0u;
        for (int k = 0;; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, k < ne);
            if (m == 0u) break;
            if (dup && k < ne) {
                const ErrEntry x = *reinterpret_cast<const ErrEntry*>(eb + ((off + __popc(m & lt)) << 3));
                if ((unsigned)x.bin < (unsigned)VRDD_BINS) {
                    const float y = cur[x.bin] + x.val;
                    cur[x.bin] = (y < 0.f) ? 0.f : y;
                }
            }
            off += __popc(m);
        }
        if (dup) {
            float tot = 0.f;
            for (int mm = 0; mm < VRDD_BINS; ++mm) tot += cur[mm];
            const float inv = (tot > 0.f) ? 1.0f / tot : 1.0f;
            float mean_raw = 0.f, E = 0.f, var = 0.f;
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                cur[mm] *= inv;
                mean_raw = fmaf(cur[mm], fmaf(bw, (float)mm, 0.5f * bw), mean_raw);
                E += plog2p(cur[mm]);
            }
            for (int mm = 0; mm < VRDD_BINS; ++mm) {
                const float dd = fmaf(bw, (float)mm, 0.5f * bw) - mean_raw;
                var = fmaf(cur[mm] * dd, dd, var);
            }
            mean_n = mean_raw * (float)(1.0 / VRDD_MEAN_NORM);
            var_n = var * (float)(1.0 / VRDD_VAR_NORM);
            ent_n = -E * 0.2f;
        }
        __syncwarp();
    }
}

// SURF_ONLY: the sink is known at compile time to be the three 3-D arrays and nothing else (one GPU, texture sampler:
// no linear / bricked planes, no peers) with whole 32-voxel rows: the generic sink's uniform tests — lin?, brick?, seven
// peers x three planes — are ~35 of the ~430 instructions a tile costs in a kernel that is bound by instruction issue.
// (Measured and rejected, round 2: two / four texels per surface store gathered with shuffles, +0.5 % / -4 %; the moment
// gather through the texture pipe, -3.5 %: profiles/tuning_r2.md.)
template <bool SCAN, bool RECOMP, bool SURF_ONLY = false>
__global__ void __launch_bounds__(1024, 1)
decode_fractal_moments2_kernel(const int4* __restrict__ codebook, const ErrEntry* __restrict__ errs,
                               const unsigned long long* __restrict__ chunk_off, const float4* __restrict__ perm,
                               const void* __restrict__ rows_g, int T, long long nvox, DecodeOut out, int pf_lines) {
    using Row = typename RowT<RECOMP>::type;
    constexpr int kSmThreads = 1024;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Row* rows_s = reinterpret_cast<Row*>(smem_raw);                                       // [T][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    {
        const float4* src = reinterpret_cast<const float4*>(rows_g);
        float4* dst = reinterpret_cast<float4*>(smem_raw);
        const int n16 = T * VRDD_BINS * (int)sizeof(Row) / 16;
        for (int i = threadIdx.x; i < n16; i += kSmThreads) dst[i] = src[i];
    }
    __syncthreads();

    // tile indices are 32-bit (the launcher checks nvox < 2^35): the per-tile bookkeeping is a handful of
    // integer instructions instead of 64-bit compare / min chains in a kernel that is issue-bound
    constexpr int kSmWarps = kSmThreads / 32;
    const int nwt = (int)((nvox + 31) / 32);
    const int wstride = (int)gridDim.x * kSmWarps;
    int wt = (int)blockIdx.x * kSmWarps + warp;
    if (wt >= nwt) return;
    const int tail = (int)(nvox - (long long)(nwt - 1) * 32);        // live lanes of the very last tile (1..32)
    const int last_vox_lane = tail - 1;

    const bool rows32 = SURF_ONLY || (((out.W & 31) == 0) && ((out.v_base & 31) == 0) && (out.use_surf || out.brick[0]));
    int x0 = 0, y0 = 0, z0 = 0, sx = 0, sy = 0, sz = 0;
    if (rows32) {
        split_voxel(out, out.v_base + (long long)wt * 32, x0, y0, z0);
        split_voxel(out, (long long)wstride * 32, sx, sy, sz);
    }

    // lane's voxel of tile t, clamped into the volume (only the last tile can be partial)
    auto vox_clamped = [&](int t) { return (long long)t * 32 + ((t == nwt - 1) ? min(lane, last_vox_lane) : lane); };
    int4 code = ldg_stream_i4(codebook + vox_clamped(wt));
    unsigned long long base = chunk_off[wt];
    // The error offsets run two tiles ahead, so the NEXT tile's errors (the bulk of a voxel's bytes, and the
    // one input whose address is data-dependent) can be pulled into L2 while this tile is processed: the
    // first pf_lines lanes touch one 128-byte line each behind errs[base_n].  ncu had the kernel at 4.5
    // warps per issue stalled on the long scoreboard with issue slots and LSU pipe no longer the limit.
    unsigned long long base_n = chunk_off[min(wt + wstride, nwt)];
    const unsigned long long err_total = chunk_off[nwt];

    while (true) {
        const long long v = (long long)wt * 32 + lane;
        const bool live = (wt != nwt - 1) || (lane < tail);
        const int wt_n = wt + wstride;
        int4 code_n;
        unsigned long long base_nn;
        if (lane < pf_lines && wt_n < nwt) {
            const unsigned long long first = base_n + 16ull * (unsigned)lane;           // 16 entries per line
            if (first < err_total) asm volatile("prefetch.global.L2 [%0];" ::"l"(errs + first));
        }
        const int id = min(max(code.x, 0), T - 1);
        const int s = code.y & (VRDD_BINS - 1);
        // flip as ARITHMETIC, not as a predicate: ptxas hoists `setp.ne code.z, 0` of the next tile across the
        // back edge to right behind the prefetch that defines code.z (a predicate frees a register), and
        // every tile then waits a full DRAM latency on the prefetch it has just issued — 32 % of all stall
        // samples sat on that one ISETP (ncu source page, r1f kernel and this one alike).
        int fl;
        asm("min.u32 %0, %1, 1;" : "=r"(fl) : "r"(code.z));        // (code.z != 0) as 0 / 1, opaque to the optimiser
        const int ne = live ? min(max(code.w, 0), VRDD_BINS) : 0;
        const float4 ent = __ldg(perm + ((id * 2 + fl) * VRDD_BINS + s));
        float mean_n, var_n, ent_n;
        moments_voxel2<SCAN, RECOMP>(rows_s + id * VRDD_BINS, errs, base, ne, s, fl * (VRDD_BINS - 1), ent, lt, mean_n,
                                     var_n, ent_n, [&] {
            // unconditional, from clamped (always valid) addresses: a predicated load merged with a default
            // value made ptxas copy one of the four registers right behind the load — the same stall again
            code_n = ldg_stream_i4(codebook + vox_clamped(min(wt_n, nwt - 1)));
            base_nn = ldg_stream_u64(chunk_off + min(wt_n + wstride, nwt));
        });
        if (live) {
            if constexpr (SURF_ONLY) {
                const int xb = (x0 + lane) * 4;
                surf3Dwrite(mean_n, out.surf[0], xb, y0, z0);
                surf3Dwrite(var_n, out.surf[1], xb, y0, z0);
                surf3Dwrite(ent_n, out.surf[2], xb, y0, z0);
            } else if (rows32) emit_decoded_xyz(out, v, x0 + lane, y0, z0, mean_n, var_n, ent_n);
            else emit_decoded(out, v, mean_n, var_n, ent_n);
        }
        if (wt_n >= nwt) break;
        wt = wt_n; code = code_n; base = base_n; base_n = base_nn;
        if (rows32) {
            x0 += sx; if (x0 >= out.W) { x0 -= out.W; ++y0; }
            y0 += sy; if (y0 >= out.H) { y0 -= out.H; ++z0; }
            z0 += sz;
        }
    }
}

size_t fractal_smem_bytes(int T, bool tmpl_in_smem) {
    return (size_t)VRDD_BINS * kThreads * sizeof(float) + (tmpl_in_smem ? (size_t)T * VRDD_BINS * sizeof(float) : 0);
}

}  // namespace

int build_template_moments(vrdd_context* c, const float* d_tmpl, int T) {
    if (c->tmpl_mom) { cudaFree(c->tmpl_mom); c->tmpl_mom = nullptr; }
    VRDD_CUDA(c, cudaMalloc(&c->tmpl_mom, (size_t)T * kMomBytesPerTemplate));
    float4* perm = reinterpret_cast<float4*>(c->tmpl_mom);
    float2* vg = reinterpret_cast<float2*>(perm + (size_t)T * 2 * VRDD_BINS);
    build_moments_kernel<<<(T * 2 * VRDD_BINS + 127) / 128, 128, 0, c->stream>>>(d_tmpl, T, perm, vg);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_decode_fractal(vrdd_context* c, const int32_t* cb, const void* errs, const uint64_t* off,
                          const float* tmpl, int T, long long nvox, const DecodeOut& out, float* d_recon) {
    if (nvox <= 0) return VRDD_OK;
    if (T <= 0) return fail(c, VRDD_ERR_INVALID, "decode_fractal: no templates");
    const int4* cb4 = reinterpret_cast<const int4*>(cb);
    const ErrEntry* er = reinterpret_cast<const ErrEntry*>(errs);
    const unsigned long long* of = reinterpret_cast<const unsigned long long*>(off);
    if (c->var_fractal >= 1 && d_recon == nullptr && c->tmpl_mom != nullptr) {
        const int threads = (c->var_fractal == 3) ? 768 : 1024;
        const size_t smem = moments_smem_bytes(T);
        const float4* perm = reinterpret_cast<const float4*>(c->tmpl_mom);
        const float2* vg = reinterpret_cast<const float2*>(perm + (size_t)T * 2 * VRDD_BINS);
        const bool gen2 = c->var_fractal >= 4 && c->var_fractal <= 7;
        const bool recomp = gen2 && (c->var_fractal & 1) != 0;
        const size_t smem2 = recomp ? (size_t)T * VRDD_BINS * sizeof(float) : smem;
        if (gen2 && smem2 <= 227 * 1024 && nvox < (1ll << 35)) {        // second-generation kernel (moments_voxel2); its tile indices are
                                                                         // 32-bit: nvox / 32 plus two grid strides must stay below 2^31
            const bool scan = c->var_fractal <= 5;
            const bool surf_only = scan && !recomp && c->var_fractal_sink != 0 && out.use_surf && !out.lin[0] && !out.brick[0] &&
                                   out.n_peers == 0 && (out.W & 31) == 0 && (out.v_base & 31) == 0 && (nvox & 31) == 0;
            auto kern = surf_only ? decode_fractal_moments2_kernel<true, false, true>
                        : scan ? (recomp ? decode_fractal_moments2_kernel<true, true> : decode_fractal_moments2_kernel<true, false>)
                               : (recomp ? decode_fractal_moments2_kernel<false, true> : decode_fractal_moments2_kernel<false, false>);
            VRDD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            const long long nblk = (nvox + 1023) / 1024;
            const int grid = (int)((nblk < c->num_sms) ? nblk : c->num_sms);
            kern<<<grid, 1024, smem2, c->stream>>>(cb4, er, of, perm, recomp ? (const void*)tmpl : (const void*)vg, T, nvox, out,
                                                   c->var_fractal_pf);
        } else if (c->var_fractal != 2 && smem <= 227 * 1024) {        // the rows fit in shared memory (T <= 908)
            auto kern = (threads == 768) ? decode_fractal_moments_smem_kernel<768> : decode_fractal_moments_smem_kernel<1024>;
            VRDD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const long long nblk = (nvox + threads - 1) / threads;
            const int grid = (int)((nblk < c->num_sms) ? nblk : c->num_sms);
            kern<<<grid, threads, smem, c->stream>>>(cb4, er, of, perm, vg, T, nvox, out);
        } else {
            const long long nblk = (nvox + kMomThreads - 1) / kMomThreads;
            const long long cap = (long long)c->num_sms * 8;
            const int grid = (int)((nblk < cap) ? nblk : cap);
            decode_fractal_moments_kernel<<<grid, kMomThreads, 0, c->stream>>>(cb4, er, of, perm, vg, T, nvox, out);
        }
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
    decode_fractal_dense_kernel<<<grid, kThreads, smem, c->stream>>>(cb4, er, of, tmpl, T, in_smem ? 1 : 0, nvox, out,
                                                                      d_recon);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

}  // namespace vrdd
