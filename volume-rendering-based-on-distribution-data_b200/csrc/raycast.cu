// raycast.cu — P2: volume ray caster, sm_100a.
//
// Replaces d_render (/root/reference/volumeRender_kernel.cu:272-717), intersectBox
// (:136-156), mul (:168-184) and rgbaFloatToInt (:186-193) for queryMethod 1..6.
//
// One thread per pixel, as in the reference, but
//   * a warp covers an 8x4 pixel tile (the reference's 16x16 block puts a warp on 16x2),
//     so the eight texels around the 32 samples of one step share cache lines;
//   * the sampled volume is ONE fp32 plane (mean, variance or entropy), 4 B per texel,
//     instead of the reference's float4 volume of which one lane is used (16 B per texel);
//   * the unconditional 8x33-fetch prologue (:354-367) and the dead 32-fetch per-step loop
//     (:605-612) are not reproduced — they cannot affect the image in these modes;
//   * image-space partitions (tiles, round-robin over ranks) and an optional fused clear
//     of missed pixels are part of the launch.
//
// Ray set-up (eye ray, slab test, first position: common.cuh, eye_ray) and the t / pos recurrences use
// explicitly rounded operations, by default with the rounding of the reference's own nvcc build
// (FMA pattern of its PTX, rsqrt.approx), so tnear, tfar and every sample position are bit-identical
// to the reference binary's and to the oracle's (Oracle.set_reference_build) and all sides take the
// same number of steps.  Compositing is plain fp32 and may contract to FMA, as in the reference's
// build; images agree with the oracle within +-1 LSB per channel.
//
// Samplers:
//   texture — tex3D<float> on a 3-D cudaArray, linear / normalised / clamp: the texture
//             unit applies the same 8-bit-weight trilinear filter the reference relies on.
//   bricked — manual trilinear on 4x4x4-texel bricks with the texture unit's exact integer
//             weight scheme (split_hw / sample_bricked below).
// Transfer function: tex1D<float4> (the reference's path, :683) or a shared-memory table
// with the same filter, chosen with vrdd_set_variant("raycast_tf", ...).
#include "common.cuh"

#include <cmath>
#include <cstring>
#include <map>
#include <utility>

namespace vrdd {

namespace {

struct RayArgs {
    cudaTextureObject_t vol_tex;
    const float* vol_brick;
    int W, H, D, bW, bH;
    cudaTextureObject_t tf_tex;
    const float4* tf_tab;
    int tf_n;
    uint32_t* out;
    int iw, ih;
    float m[12];
    float density, brightness, t_offset, t_scale, tstep, thresh;
    int max_steps;
    TileMap tiles;                   // which pixels this launch renders (common.cuh)
    FrameSignal done;                // frame-complete signal (common.cuh)
    int pow2_shift[3];               // gather path, all extents powers of two <= 4096: 13 - log2(extent) per axis
    int clear_misses;
    int ref_rounding;                // ray set-up rounded like the reference's nvcc build (common.cuh, eye_ray)
    unsigned long long* samples;
};

constexpr int kBlock = 256;          // 8 warps = 16x16 pixels, each warp an 8x4 sub-tile

// The texture unit's linear filter, restated in integer arithmetic (measured on B200 with
// one-hot / ramp volumes, tools/probe_texture*.py; the oracle carries the same scheme):
//   U = trunc(sat(u) * 2^21);  q = ((U * N * 256 + 2^20) >> 21) - 128, clamped to [0,(N-1)*256];
//   texel i = q >> 8, weight A = q & 255 (out of 256).
__device__ __forceinline__ void split_hw(float u, int n256, int& i, int& a) {
    const unsigned U = (unsigned)(__saturatef(u) * 2097152.0f);
    const unsigned long long p = (unsigned long long)U * (unsigned)n256 + (1u << 20);
    int q = (int)(p >> 21) - 128;
    q = min(max(q, 0), n256 - 256);
    i = q >> 8;
    a = q & 255;
}

// Measured and rejected (round 2): the table resampled per block at the filter's full fixed-point resolution
// ((n - 1) * 256 + 1 float4 entries, 32 KB for the nine-entry rainbow), so that a lookup is one LDS.128 — 18 of the
// ~30 instructions of a lookup go, but the extra 16 KB per block come out of L1: the 3-D array path takes 0.41 instead
// of 0.30 ms on oblique views, the gather path gains 2 %.
__device__ __forceinline__ float4 tf_lookup_smem(const float4* tab, int n, float u) {
    int i, a;
    split_hw(u, n << 8, i, a);
    const float4 c0 = tab[i], c1 = tab[min(i + 1, n - 1)];
    const float w0 = (float)(256 - a) * (1.0f / 256.0f), w1 = (float)a * (1.0f / 256.0f);
    return make_float4(w0 * c0.x + w1 * c1.x, w0 * c0.y + w1 * c1.y, w0 * c0.z + w1 * c1.z, w0 * c0.w + w1 * c1.w);
}

// Manual trilinear sample with the texture unit's eight integer weights (they always sum to
// 256): split z, then x, then y, rounding half up on the branches the hardware rounds.
__device__ __forceinline__ float sample_bricked(const RayArgs& A, float u, float v, float w) {
    int i, j, k, a, b, c;
    split_hw(u, A.W << 8, i, a);
    split_hw(v, A.H << 8, j, b);
    split_hw(w, A.D << 8, k, c);
    const int i1 = min(i + 1, A.W - 1), j1 = min(j + 1, A.H - 1), k1 = min(k + 1, A.D - 1);
    const float* B = A.vol_brick;
    const float t000 = __ldg(B + brick_index(i, j, k, A.bW, A.bH));
    const float t100 = __ldg(B + brick_index(i1, j, k, A.bW, A.bH));
    const float t010 = __ldg(B + brick_index(i, j1, k, A.bW, A.bH));
    const float t110 = __ldg(B + brick_index(i1, j1, k, A.bW, A.bH));
    const float t001 = __ldg(B + brick_index(i, j, k1, A.bW, A.bH));
    const float t101 = __ldg(B + brick_index(i1, j, k1, A.bW, A.bH));
    const float t011 = __ldg(B + brick_index(i, j1, k1, A.bW, A.bH));
    const float t111 = __ldg(B + brick_index(i1, j1, k1, A.bW, A.bH));
    const int z1 = c, z0 = 256 - c;
    const int x10 = (z0 * a + 128) >> 8, x00 = z0 - x10;          // x-split of the lower z slice
    const int x11 = (z1 * a + 128) >> 8, x01 = z1 - x11;          // ... and of the upper one
    const int w000 = (x00 * (256 - b) + 128) >> 8, w010 = x00 - w000;
    const int w110 = (x10 * b + 128) >> 8, w100 = x10 - w110;
    const int w001 = (x01 * (256 - b) + 128) >> 8, w011 = x01 - w001;
    const int w111 = (x11 * b + 128) >> 8, w101 = x11 - w111;
    float acc = (float)w000 * t000;
    acc = fmaf((float)w010, t010, acc);
    acc = fmaf((float)w100, t100, acc);
    acc = fmaf((float)w110, t110, acc);
    acc = fmaf((float)w001, t001, acc);
    acc = fmaf((float)w011, t011, acc);
    acc = fmaf((float)w101, t101, acc);
    acc = fmaf((float)w111, t111, acc);
    return acc * (1.0f / 256.0f);
}

__device__ __forceinline__ uint32_t pack_rgba(float r, float g, float b, float a) {
    // saturate, scale, truncate, pack (volumeRender_kernel.cu:186-193)
    return ((uint32_t)(__saturatef(a) * 255.0f) << 24) | ((uint32_t)(__saturatef(b) * 255.0f) << 16) |
           ((uint32_t)(__saturatef(g) * 255.0f) << 8) | (uint32_t)(__saturatef(r) * 255.0f);
}

// The march is software-pipelined in batches of U steps.  A ray's sample positions follow
// from geometry alone (pos += step, t += tstep, stop when t > tfar or after max_steps), so
// the U volume fetches of a batch are issued back to back, then the U transfer-function
// lookups, and only the compositing — with the reference's early exit at alpha > threshold
// (:698) — is sequential.  Samples fetched beyond an early exit are discarded and never
// counted, so images and sample counts are those of the step-by-step loop; what changes is
// that U texture requests per thread are in flight instead of one (ncu: the one-step loop
// idles on long-scoreboard stalls with the L1TEX pipe at 41 %).
template <int SAMPLER, int TFMODE, bool COUNT, int U>
__global__ void __launch_bounds__(kBlock) raycast_kernel(const RayArgs A) {
    // the transfer-function table, sized to the function (dynamic shared memory: tf_n * 16 B; a static 1024-entry array
    // would take 16 KB per block, 80 KB per SM, out of the L1 the volume fetches live in)
    extern __shared__ float4 tf_s[];
    if (TFMODE == 1) {
        for (int i = threadIdx.x; i < A.tf_n; i += kBlock) tf_s[i] = A.tf_tab[i];
        __syncthreads();
    }

    const int lane = threadIdx.x & 31;
    unsigned long long nsamp = 0;
    int it = 0;
    for (int item = A.done.queue_begin(); item < A.tiles.n_items; item = A.done.queue_next(it++)) {
        A.done.queue_prefetch(it);
        int x, y;
        if (A.tiles.pixel(item, x, y)) {
            // ---- eye ray, slab test (volumeRender_kernel.cu:288-303): common.cuh, eye_ray ---------
            const EyeRay R = eye_ray(A.m, x, y, A.iw, A.ih, A.ref_rounding);
            const float dx = R.dx, dy = R.dy, dz = R.dz, tfar = R.tfar;
            float tnear = R.tnear;

            if (tfar > tnear) {
                if (tnear < 0.0f) tnear = 0.0f;                                          // :305-306
                float sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
                float t = tnear;
                float px, py, pz;
                eye_ray_start(R, tnear, A.ref_rounding, px, py, pz);                     // :311
                const float stx = __fmul_rn(dx, A.tstep), sty = __fmul_rn(dy, A.tstep), stz = __fmul_rn(dz, A.tstep);
                int i = 0;
                bool alive = A.max_steps > 0;                 // the geometric state (i, t, p) is a live step
                while (alive) {
                    // -- geometry of the next U steps (:381, :701-706), exactly as the one-step loop
                    float cu[U], cv[U], cw[U];
                    bool valid[U];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        valid[k] = alive;
                        // pos*0.5 is exact, so fma == mul-then-add here
                        cu[k] = fmaf(px, 0.5f, 0.5f); cv[k] = fmaf(py, 0.5f, 0.5f); cw[k] = fmaf(pz, 0.5f, 0.5f);
                        const float tn = __fadd_rn(t, A.tstep);
                        const bool cont = alive && !(tn > tfar) && (i + 1 < A.max_steps);
                        if (cont) {
                            t = tn; ++i;
                            px = __fadd_rn(px, stx); py = __fadd_rn(py, sty); pz = __fadd_rn(pz, stz);
                        }
                        alive = cont;
                    }
                    // -- U volume fetches in flight (:601-651)
                    float s[U];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        s[k] = 0.f;
                        if (valid[k]) {
                            if (SAMPLER == 0) s[k] = tex3D<float>(A.vol_tex, cu[k], cv[k], cw[k]);
                            else s[k] = sample_bricked(A, cu[k], cv[k], cw[k]);
                        }
                    }
                    // -- U transfer-function lookups (:683-684)
                    float4 col[U];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        col[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid[k]) {
                            const float tu = (s[k] - A.t_offset) * A.t_scale;
                            if (TFMODE == 0) col[k] = tex1D<float4>(A.tf_tex, tu);
                            else col[k] = tf_lookup_smem(tf_s, A.tf_n, tu);
                        }
                    }
                    // -- front-to-back compositing, in order, with the early exit (:685-699)
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        if (!valid[k]) { alive = false; break; }
                        if (COUNT) ++nsamp;
                        float4 c = col[k];
                        c.w *= A.density;
                        c.x *= c.w; c.y *= c.w; c.z *= c.w;
                        const float kk = 1.0f - sa;
                        sr += c.x * kk; sg += c.y * kk; sb += c.z * kk; sa += c.w * kk;
                        if (sa > A.thresh) { alive = false; break; }
                    }
                }
                A.out[(size_t)y * A.iw + x] =
                    pack_rgba(sr * A.brightness, sg * A.brightness, sb * A.brightness, sa * A.brightness);  // :713-716
            } else if (A.clear_misses) {
                A.out[(size_t)y * A.iw + x] = 0u;                                        // volumeRender.cpp:208
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nsamp += __shfl_xor_sync(0xffffffffu, nsamp, d);
        if (lane == 0 && nsamp) atomicAdd(A.samples, nsamp);
    }
    A.done.block_done();
}

// ---- queryMethod 7: interpolated mean (volumeRender_kernel.cu:253-270, 320-367, 395-480) ---------------
// The eight corners of the cell around the sample are floor/ceil(pos01*dim)/dim; each corner point-samples
// the block-index texture (nearest-texel rule of the texture unit: idx = (trunc(sat(u)*2^21) * N) >> 21,
// measured with tools/probe_texture3.py) and takes that block's un-normalised mean; the sample is their
// trilinear blend in double precision times 50.  The corner cache is refreshed when the sample leaves the
// cell.  When pos01*dim is an integer the blend divides by zero and the sample is NaN — the reference's
// "vertical and horizontal line" artefact (ver1.9.6.txt:166); a NaN transfer-function coordinate reads
// texel 0, as on the hardware.  The reference runs this mode below 5 fps (ver1.9.6.txt:168) because it
// re-reads 8x32 histogram bins at every refresh; here the means are decoded once.
__device__ __forceinline__ int point_index_hw(float u, int n) {
    const unsigned U = (unsigned)(__saturatef(u) * 2097152.0f);
    const unsigned long long i = ((unsigned long long)U * (unsigned)n) >> 21;
    return (int)min(i, (unsigned long long)(n - 1));
}

// The march of mode 7 was bound by the conversion/special-function pipe (ncu: XU 63 %, everything else
// < 30 %): 6 FRND, 6 F2I, 9 MUFU.RCP and 24 fp64<->fp32 conversions per sample.  The helpers below give the
// same values without that pipe.
// floor() for |v| < 2^22: adding 1.5 * 2^23 rounds to an integer, one compare fixes round-up.
__device__ __forceinline__ float floor_small(float v) {
    const float f = __fadd_rn(__fadd_rn(v, 12582912.0f), -12582912.0f);
    return (f > v) ? __fadd_rn(f, -1.0f) : f;
}
// point_index_hw without F2I (0 <= sat(u) * 2^21 <= 2^21)
__device__ __forceinline__ int point_index_fast(float u, int n) {
    const float f = floor_small(__saturatef(u) * 2097152.0f);
    const unsigned U = (unsigned)(__float_as_int(__fadd_rn(f, 12582912.0f)) - 0x4B400000);
    const unsigned long long i = ((unsigned long long)U * (unsigned)n) >> 21;
    return (int)min(i, (unsigned long long)(n - 1));
}
// k / n correctly rounded, for an integer-valued k in [-1, n + 1] and n <= 8192, from the correctly rounded
// reciprocal and two FMAs (checked exhaustively against the division for every such k and n).
__device__ __forceinline__ float div_small(float k, float n, float inv_n) {
    const float q = __fmul_rn(k, inv_n);
    return fmaf(fmaf(-q, n, k), inv_n, q);
}

struct Mode7Args {
    const float* mean_raw;
    cudaTextureObject_t mean_tex;    // the same plane as a point-sampled 3-D array
    cudaTextureObject_t mean_gather; // the same plane as a layered 2-D array for tld4 (GATHER kernels)
    int W, H, D;
    const float4* tf_tab;
    int tf_n;
    uint32_t* out;
    int iw, ih;
    float m[12];
    float density, brightness, t_offset, t_scale, tstep, thresh;
    int max_steps, clear_misses;
    TileMap tiles;
    FrameSignal done;
    int use_tab, idx32;
    int ref_rounding;                // ray set-up rounded like the reference's nvcc build (common.cuh, eye_ray)
    unsigned long long* samples;
};

// tld4 on a layered 2-D texture: the four texels a bilinear fetch at (x, y) of layer `layer` would blend,
// unfiltered: .x = (i, j+1), .y = (i+1, j+1), .z = (i+1, j), .w = (i, j) with i = floor(x - .5), j = floor(y - .5).
__device__ __forceinline__ float4 gather_layer(cudaTextureObject_t tex, float x, float y, int layer) {
    float4 r;
    asm volatile("tld4.r.a2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6, %7, %7}];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(tex), "r"(layer), "f"(x), "f"(y));
    return r;
}

// GATHER: the 2x2 (x, y) corners of a cell come from ONE tld4 per z layer (two fetches per sample instead of
// eight).  Valid when the texel the texture unit's point rule gives boundary k is k itself on the x and y axes
// (checked on the host for the volume's extents; true for powers of two, where k/n is exact): the corners are
// then texels (k, k+1), which is the footprint of a fetch at the texel corner k+1.  The two cases where the
// reference fetches the SAME texel twice are covered too: at the last cell the clamp returns texel n-1 for
// n, and in a degenerate cell (floor == ceil) the blend is 0/0 = NaN whatever the means are.
template <bool COUNT, int U, bool GATHER>
__global__ void __launch_bounds__(kBlock, 2) raycast_mode7_kernel(const Mode7Args A) {
    // dynamic shared memory: the transfer-function table (tf_n float4), then the cell table.
    // Per axis and per cell boundary k in [-1, n+1]: {fl(k/n), index the texture unit's point rule gives it}.
    // Both depend only on (k, n), so every block tabulates them once (A.use_tab: they fit in shared memory)
    // and the march replaces 6 divisions and 6 index computations per sample by 6 shared-memory reads.
    extern __shared__ float4 tf_s[];
    float2* cell_tab = reinterpret_cast<float2*>(tf_s + A.tf_n);
    const float2* tabx = cell_tab;
    const float2* taby = cell_tab + (A.W + 3);
    const float2* tabz = taby + (A.H + 3);
    for (int i = threadIdx.x; i < A.tf_n; i += kBlock) tf_s[i] = A.tf_tab[i];
    if (A.use_tab) {
        const int nx = A.W + 3, ny = A.H + 3, nz = A.D + 3;
        for (int e = threadIdx.x; e < nx + ny + nz; e += kBlock) {
            const int n = (e < nx) ? A.W : (e < nx + ny) ? A.H : A.D;
            const int k = ((e < nx) ? e : (e < nx + ny) ? e - nx : e - nx - ny) - 1;
            const float fn = (float)n;
            const float v = div_small((float)k, fn, __fdiv_rn(1.0f, fn));
            cell_tab[e] = make_float2(v, (float)point_index_fast(v, n) + 0.5f);      // texel centre
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned long long nsamp = 0;
    int it = 0;
    for (int item = A.done.queue_begin(); item < A.tiles.n_items; item = A.done.queue_next(it++)) {
        A.done.queue_prefetch(it);
        int x, y;
        if (A.tiles.pixel(item, x, y)) {
            const EyeRay R = eye_ray(A.m, x, y, A.iw, A.ih, A.ref_rounding);
            const float dx = R.dx, dy = R.dy, dz = R.dz, tfar = R.tfar;
            float tnear = R.tnear;
            if (tfar > tnear) {
                if (tnear < 0.0f) tnear = 0.0f;
                float sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f, t = tnear;
                float px, py, pz;
                eye_ray_start(R, tnear, A.ref_rounding, px, py, pz);
                const float stx = __fmul_rn(dx, A.tstep), sty = __fmul_rn(dy, A.tstep), stz = __fmul_rn(dz, A.tstep);
                const float fW = (float)A.W, fH = (float)A.H, fD = (float)A.D;
                // The reference keeps the eight corner means of the current cell and refreshes them when a sample
                // leaves it (:396).  Here the march runs in batches of U steps like raycast_kernel: for each of
                // the next U samples the cell a refresh WOULD produce at that sample is computed and its eight
                // loads are issued up front (8U loads in flight); the in-order pass below then either keeps the
                // cached cell or adopts the sample's own, which is exactly the refresh the reference would do.
                float botx = 0.f, boty = 0.f, botz = 0.f, topx = 0.f, topy = 0.f, topz = 0.f, mean[8];
    #pragma unroll
                for (int j = 0; j < 8; ++j) mean[j] = 0.f;
                bool have = false;
                int i = 0;
                bool alive = A.max_steps > 0;
                while (alive) {
                    float cu[U], cv[U], cw[U];
                    bool valid[U];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        valid[k] = alive;
                        cu[k] = fmaf(px, 0.5f, 0.5f); cv[k] = fmaf(py, 0.5f, 0.5f); cw[k] = fmaf(pz, 0.5f, 0.5f);
                        const float tn = __fadd_rn(t, A.tstep);
                        const bool cont = alive && !(tn > tfar) && (i + 1 < A.max_steps);
                        if (cont) {
                            t = tn; ++i;
                            px = __fadd_rn(px, stx); py = __fadd_rn(py, sty); pz = __fadd_rn(pz, stz);
                        }
                        alive = cont;
                    }
                    float sbx[U], sby[U], sbz[U], stx_[U], sty_[U], stz_[U], sm[U][8];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {                                       // :322-367, :398-463
                        int x0, x1, y0, y1, z0, z1;
                        float xf0, xf1, yf0, yf1, zf0, zf1;              // texel centres
                        float gfx = 0.f, gfy = 0.f;                      // floor(pos01 * dim) in x and y (GATHER)
                        if (A.use_tab) {
                            const float vx = __fmul_rn(cu[k], fW), vy = __fmul_rn(cv[k], fH), vz = __fmul_rn(cw[k], fD);
                            const float fx = floor_small(vx), fy = floor_small(vy), fz = floor_small(vz);
                            gfx = fx; gfy = fy;
                            // k + 1 for floor and ceil, as integers (|f| < 2^22), clamped to the table
                            const int ex0 = min(max(__float_as_int(__fadd_rn(fx, 12582912.0f)) - 0x4B400000 + 1, 0), A.W + 1);
                            const int ey0 = min(max(__float_as_int(__fadd_rn(fy, 12582912.0f)) - 0x4B400000 + 1, 0), A.H + 1);
                            const int ez0 = min(max(__float_as_int(__fadd_rn(fz, 12582912.0f)) - 0x4B400000 + 1, 0), A.D + 1);
                            const float2 ax = tabx[ex0], bx_ = tabx[ex0 + (fx < vx)];
                            const float2 ay = taby[ey0], by_ = taby[ey0 + (fy < vy)];
                            const float2 az = tabz[ez0], bz_ = tabz[ez0 + (fz < vz)];
                            sbx[k] = ax.x; stx_[k] = bx_.x; xf0 = ax.y; xf1 = bx_.y;
                            sby[k] = ay.x; sty_[k] = by_.x; yf0 = ay.y; yf1 = by_.y;
                            sbz[k] = az.x; stz_[k] = bz_.x; zf0 = az.y; zf1 = bz_.y;
                            x0 = (int)xf0; x1 = (int)xf1; y0 = (int)yf0; y1 = (int)yf1; z0 = (int)zf0; z1 = (int)zf1;   // dead on the texture path
                        } else {
                            sbx[k] = __fdiv_rn(floorf(__fmul_rn(cu[k], fW)), fW); stx_[k] = __fdiv_rn(ceilf(__fmul_rn(cu[k], fW)), fW);
                            sby[k] = __fdiv_rn(floorf(__fmul_rn(cv[k], fH)), fH); sty_[k] = __fdiv_rn(ceilf(__fmul_rn(cv[k], fH)), fH);
                            sbz[k] = __fdiv_rn(floorf(__fmul_rn(cw[k], fD)), fD); stz_[k] = __fdiv_rn(ceilf(__fmul_rn(cw[k], fD)), fD);
                            x0 = point_index_hw(sbx[k], A.W); x1 = point_index_hw(stx_[k], A.W);
                            y0 = point_index_hw(sby[k], A.H); y1 = point_index_hw(sty_[k], A.H);
                            z0 = point_index_hw(sbz[k], A.D); z1 = point_index_hw(stz_[k], A.D);
                            xf0 = (float)x0 + 0.5f; xf1 = (float)x1 + 0.5f; yf0 = (float)y0 + 0.5f; yf1 = (float)y1 + 0.5f;
                            zf0 = (float)z0 + 0.5f; zf1 = (float)z1 + 0.5f;
                        }
    #pragma unroll
                        for (int j = 0; j < 8; ++j) sm[k][j] = 0.f;
                        if (GATHER) {
                            if (valid[k]) {
                                // the texel corner at the cell's upper boundary, floor + 1 clamped to [0, n]: texels
                                // (k, k+1) inside, (n-1, n-1) at and beyond the last boundary, (0, 0) for the cell
                                // below zero that rounding of the box entry can produce (floor == -1)
                                const float gx = fminf(fmaxf(__fadd_rn(gfx, 1.0f), 0.0f), fW);
                                const float gy = fminf(fmaxf(__fadd_rn(gfy, 1.0f), 0.0f), fH);
                                // layer = texel centre - 0.5 as an integer, without the conversion pipe
                                const int l0 = __float_as_int(__fadd_rn(__fadd_rn(zf0, -0.5f), 12582912.0f)) - 0x4B400000;
                                const int l1 = __float_as_int(__fadd_rn(__fadd_rn(zf1, -0.5f), 12582912.0f)) - 0x4B400000;
                                const float4 a = gather_layer(A.mean_gather, gx, gy, l0);
                                const float4 b = gather_layer(A.mean_gather, gx, gy, l1);
                                sm[k][0] = a.w; sm[k][1] = a.z; sm[k][2] = a.x; sm[k][3] = a.y;
                                sm[k][4] = b.w; sm[k][5] = b.z; sm[k][6] = b.x; sm[k][7] = b.y;
                            }
                        } else if (valid[k] && A.mean_tex) {
                            sm[k][0] = tex3D<float>(A.mean_tex, xf0, yf0, zf0); sm[k][1] = tex3D<float>(A.mean_tex, xf1, yf0, zf0);
                            sm[k][2] = tex3D<float>(A.mean_tex, xf0, yf1, zf0); sm[k][3] = tex3D<float>(A.mean_tex, xf1, yf1, zf0);
                            sm[k][4] = tex3D<float>(A.mean_tex, xf0, yf0, zf1); sm[k][5] = tex3D<float>(A.mean_tex, xf1, yf0, zf1);
                            sm[k][6] = tex3D<float>(A.mean_tex, xf0, yf1, zf1); sm[k][7] = tex3D<float>(A.mean_tex, xf1, yf1, zf1);
                        } else if (valid[k]) {
                            if (A.idx32) {                           // W*H*D < 2^31: 32-bit index arithmetic
                                const unsigned zb0 = (unsigned)A.H * z0, zb1 = (unsigned)A.H * z1;
                                const unsigned r00 = (unsigned)A.W * (y0 + zb0), r10 = (unsigned)A.W * (y1 + zb0);
                                const unsigned r01 = (unsigned)A.W * (y0 + zb1), r11 = (unsigned)A.W * (y1 + zb1);
                                sm[k][0] = __ldg(A.mean_raw + (r00 + x0)); sm[k][1] = __ldg(A.mean_raw + (r00 + x1));
                                sm[k][2] = __ldg(A.mean_raw + (r10 + x0)); sm[k][3] = __ldg(A.mean_raw + (r10 + x1));
                                sm[k][4] = __ldg(A.mean_raw + (r01 + x0)); sm[k][5] = __ldg(A.mean_raw + (r01 + x1));
                                sm[k][6] = __ldg(A.mean_raw + (r11 + x0)); sm[k][7] = __ldg(A.mean_raw + (r11 + x1));
                            } else {
                                const size_t r00 = (size_t)A.W * (y0 + (size_t)A.H * z0), r10 = (size_t)A.W * (y1 + (size_t)A.H * z0);
                                const size_t r01 = (size_t)A.W * (y0 + (size_t)A.H * z1), r11 = (size_t)A.W * (y1 + (size_t)A.H * z1);
                                sm[k][0] = __ldg(A.mean_raw + r00 + x0); sm[k][1] = __ldg(A.mean_raw + r00 + x1);
                                sm[k][2] = __ldg(A.mean_raw + r10 + x0); sm[k][3] = __ldg(A.mean_raw + r10 + x1);
                                sm[k][4] = __ldg(A.mean_raw + r01 + x0); sm[k][5] = __ldg(A.mean_raw + r01 + x1);
                                sm[k][6] = __ldg(A.mean_raw + r11 + x0); sm[k][7] = __ldg(A.mean_raw + r11 + x1);
                            }
                        }
                    }
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        if (!valid[k]) { alive = false; break; }
                        const float cx = cu[k], cy = cv[k], cz = cw[k];
                        if (!have || cx < botx || cy < boty || cz < botz || cx > topx || cy > topy || cz > topz) {   // :396
                            botx = sbx[k]; boty = sby[k]; botz = sbz[k]; topx = stx_[k]; topy = sty_[k]; topz = stz_[k];
    #pragma unroll
                            for (int j = 0; j < 8; ++j) mean[j] = sm[k][j];
                            have = true;
                        }
                        // :466-479.  The reference blends in double (its literals promote); fp32 with the same
                        // expression shape keeps the special values (0/0 -> NaN, x/0 -> inf, inf - inf -> NaN: the
                        // degenerate-cell artefact) and differs by an ulp at most, far inside the +-1 LSB bar.
                        const float xd = __fdividef(__fsub_rn(cx, botx), __fsub_rn(topx, botx));   // a * rcp(b): 0/0 and x/0
                        const float yd = __fdividef(__fsub_rn(cy, boty), __fsub_rn(topy, boty));   // stay NaN and inf
                        const float zd = __fdividef(__fsub_rn(cz, botz), __fsub_rn(topz, botz));
                        const float wx = 1.0f - xd, wy = 1.0f - yd, wz = 1.0f - zd;
                        const float m00 = fmaf(mean[1], xd, mean[0] * wx);
                        const float m10 = fmaf(mean[3], xd, mean[2] * wx);
                        const float m01 = fmaf(mean[5], xd, mean[4] * wx);
                        const float m11 = fmaf(mean[7], xd, mean[6] * wx);
                        const float m0 = fmaf(m10, yd, m00 * wy);
                        const float m1 = fmaf(m11, yd, m01 * wy);
                        const float s = fmaf(m1, zd, m0 * wz) * 50.0f;
                        if (COUNT) ++nsamp;
                        float4 col = tf_lookup_smem(tf_s, A.tf_n, (s - A.t_offset) * A.t_scale);
                        col.w *= A.density;
                        col.x *= col.w; col.y *= col.w; col.z *= col.w;
                        const float kk = 1.0f - sa;
                        sr += col.x * kk; sg += col.y * kk; sb += col.z * kk; sa += col.w * kk;
                        if (sa > A.thresh) { alive = false; break; }
                    }
                }
                A.out[(size_t)y * A.iw + x] = pack_rgba(sr * A.brightness, sg * A.brightness, sb * A.brightness, sa * A.brightness);
            } else if (A.clear_misses) {
                A.out[(size_t)y * A.iw + x] = 0u;
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nsamp += __shfl_xor_sync(0xffffffffu, nsamp, d);
        if (lane == 0 && nsamp) atomicAdd(A.samples, nsamp);
    }
    A.done.block_done();
}

#ifdef VRDD_PROBE_EXPORTS
__global__ void debug_sample_kernel(cudaTextureObject_t tex, const float* __restrict__ uvw, int n,
                                    float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex3D<float>(tex, uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2]);
}

__global__ void debug_sample_tf_kernel(cudaTextureObject_t tex, const float* __restrict__ u, int n,
                                       float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex1D<float4>(tex, u[i]);
}
#endif  // VRDD_PROBE_EXPORTS


// ---- gather path: the plane as a layered 2-D array whose LAYERS are stacked along x or y -----------------------
// What DRAM delivers is the full 128-byte line: 8 (x) x 4 (y) x 1 (z) texels of a CUDA array, 3-D or layered, whatever
// the load asks for (tools/probe_atom.cu, probe_atom_linear.cu: a random 4-byte read moves 127 B; the L2 fetch-
// granularity limit changes nothing).  With the reference's fixed step a 1024^3 volume is sampled every 5 voxels
// along the ray and every ~2 across, so almost every line of the region the rays cross holds a texel of some sample:
// 64 B per sample, compulsory — EXCEPT when the rays run along the axis the lines are thin in, because then all rays
// sample the same planes and three slices out of five are never touched (frontal views of the 3-D array, thin in z:
// 24 B per sample; ncu, profiles/raycast_orbit_r1h.csv).  A copy whose layers are stacked along x (or y) gives the
// views that look along x (or y) the same advantage.  Only the identity mapping of the axes keeps the texture unit's
// weight split (z, then x, then y), so on a copy the filter moves into the kernel: the eight raw texels come from
// two tld4 fetches (one per layer) and are blended with the unit's integer weights, bit for bit the scheme of
// split_hw / sample_bricked above.
// AXIS = 1: layers along x — array (x' = y, y' = z, layer = x);  AXIS = 2: layers along y — array (x' = z, y' = x, layer = y).
template <bool POW2>
__device__ __forceinline__ int split_q(float u, unsigned n256, int sh) {
    // q = 256 * texel + weight of split_hw, without the conversion pipe: 2^23 + trunc(sat(u) * 2^21) in one FFMA.RZ
    const float r = __fmaf_rz(__saturatef(u), 2097152.0f, 8388608.0f);
    const unsigned U = __float_as_uint(r) & 0x7fffffu;
    int q;
    if (POW2) {
        // N = 2^m (m <= 12): (U * 2^(m+8) + 2^20) >> 21 = (U + 2^(12-m)) >> (13-m), sh = 13 - m
        q = (int)((U + ((1u << sh) >> 1)) >> sh) - 128;
    } else {
        const long long p = (long long)((unsigned long long)U * n256) + ((1ll << 20) - (128ll << 21));
        q = (int)(p >> 21);
    }
    return min(max(q, 0), (int)n256 - 256);
}
__device__ __forceinline__ float small_int_as_float(int i) {        // 0 <= i < 2^23, without I2F
    return __uint_as_float(0x4b000000u | (unsigned)i) - 8388608.0f;
}

struct GatherFetch { float4 l0, l1; int abc; };                      // raw texels of both layers, weights packed a | b<<8 | c<<16

template <int AXIS, bool POW2>
__device__ __forceinline__ GatherFetch gather_issue(cudaTextureObject_t tex, float cu, float cv, float cw, int W, int H, int D, int shx,
                                                    int shy, int shz) {
    const int qx = split_q<POW2>(cu, (unsigned)W << 8, shx), qy = split_q<POW2>(cv, (unsigned)H << 8, shy),
              qz = split_q<POW2>(cw, (unsigned)D << 8, shz);
    const int i = qx >> 8, j = qy >> 8, k = qz >> 8;
    GatherFetch G;
    // a | b << 8 | c << 16 in two PRMT (q < 2^24, so byte 3 of qx is the zero byte the selectors borrow)
    G.abc = (int)__byte_perm(__byte_perm((unsigned)qx, (unsigned)qy, 0x3340), (unsigned)qz, 0x3410);
    // texel corner (row index + 1, column index + 1): the 2x2 footprint {idx, idx + 1}^2, clamped at the far edge
    // (where the weight of the clamped texel is 0: q <= 256 * (N - 1))
    if (AXIS == 1) {
        const float gx = small_int_as_float(j + 1), gy = small_int_as_float(k + 1);
        G.l0 = gather_layer(tex, gx, gy, i);
        G.l1 = gather_layer(tex, gx, gy, min(i + 1, W - 1));
    } else {
        const float gx = small_int_as_float(k + 1), gy = small_int_as_float(i + 1);
        G.l0 = gather_layer(tex, gx, gy, j);
        G.l1 = gather_layer(tex, gx, gy, min(j + 1, H - 1));
    }
    return G;
}

// tld4 returns .w = (x', y'), .z = (x'+1, y'), .x = (x', y'+1), .y = (x'+1, y'+1)
template <int AXIS>
__device__ __forceinline__ float gather_blend(const GatherFetch& G) {
    const int a = G.abc & 255, b = (int)__byte_perm((unsigned)G.abc, 0u, 0x4441), c = G.abc >> 16;
    float t000, t100, t010, t110, t001, t101, t011, t111;            // t[x][y][z]
    if (AXIS == 1) {       // x' = y, y' = z, layer = x
        t000 = G.l0.w; t010 = G.l0.z; t001 = G.l0.x; t011 = G.l0.y;
        t100 = G.l1.w; t110 = G.l1.z; t101 = G.l1.x; t111 = G.l1.y;
    } else {               // x' = z, y' = x, layer = y
        t000 = G.l0.w; t001 = G.l0.z; t100 = G.l0.x; t101 = G.l0.y;
        t010 = G.l1.w; t011 = G.l1.z; t110 = G.l1.x; t111 = G.l1.y;
    }
    // the unit's eight integer weights: split z, then x, then y (sample_bricked).  Measured and rejected: the same
    // weights in fp32 (each rounded product as fma(fma(p, m, .5), 2^-8, 1.5 * 2^23) - 1.5 * 2^23, exact; three
    // conversions instead of eight, the work moved from the integer to the FMA pipe): 54 instead of 48 registers, four
    // instead of five blocks per SM, 0.205 instead of 0.197 ms on a side view.
    const int z1 = c, z0 = 256 - c;
    const int x10 = (z0 * a + 128) >> 8, x00 = z0 - x10;
    const int x11 = (z1 * a + 128) >> 8, x01 = z1 - x11;
    const int w000 = (x00 * (256 - b) + 128) >> 8, w010 = x00 - w000;
    const int w110 = (x10 * b + 128) >> 8, w100 = x10 - w110;
    const int w001 = (x01 * (256 - b) + 128) >> 8, w011 = x01 - w001;
    const int w111 = (x11 * b + 128) >> 8, w101 = x11 - w111;
    float acc = (float)w000 * t000;
    acc = fmaf((float)w010, t010, acc);
    acc = fmaf((float)w100, t100, acc);
    acc = fmaf((float)w110, t110, acc);
    acc = fmaf((float)w001, t001, acc);
    acc = fmaf((float)w011, t011, acc);
    acc = fmaf((float)w101, t101, acc);
    acc = fmaf((float)w111, t111, acc);
    return acc * (1.0f / 256.0f);
}

// Same march as raycast_kernel (batches of U steps, in-order compositing with the reference's early exit); only the
// sampler differs.  RayArgs::vol_tex is the layered copy.
template <int AXIS, int TFMODE, bool POW2, bool COUNT, int U>
__global__ void __launch_bounds__(kBlock) raycast_gather_kernel(const RayArgs A) {
    // the transfer-function table, sized to the function (dynamic shared memory: tf_n * 16 B; a static 1024-entry array
    // would take 16 KB per block, 80 KB per SM, out of the L1 the volume fetches live in)
    extern __shared__ float4 tf_s[];
    if (TFMODE == 1) {
        for (int i = threadIdx.x; i < A.tf_n; i += kBlock) tf_s[i] = A.tf_tab[i];
        __syncthreads();
    }
    const int shx = A.pow2_shift[0], shy = A.pow2_shift[1], shz = A.pow2_shift[2];
    const int lane = threadIdx.x & 31;
    unsigned long long nsamp = 0;
    int it = 0;
    for (int item = A.done.queue_begin(); item < A.tiles.n_items; item = A.done.queue_next(it++)) {
        A.done.queue_prefetch(it);
        int x, y;
        if (A.tiles.pixel(item, x, y)) {
            const EyeRay R = eye_ray(A.m, x, y, A.iw, A.ih, A.ref_rounding);
            const float tfar = R.tfar;
            float tnear = R.tnear;
            if (tfar > tnear) {
                if (tnear < 0.0f) tnear = 0.0f;
                float sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
                float t = tnear;
                float px, py, pz;
                eye_ray_start(R, tnear, A.ref_rounding, px, py, pz);
                const float stx = __fmul_rn(R.dx, A.tstep), sty = __fmul_rn(R.dy, A.tstep), stz = __fmul_rn(R.dz, A.tstep);
                int i = 0;
                bool alive = A.max_steps > 0;
                while (alive) {
                    GatherFetch G[U];
                    bool valid[U];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        valid[k] = alive;
                        if (alive) G[k] = gather_issue<AXIS, POW2>(A.vol_tex, fmaf(px, 0.5f, 0.5f), fmaf(py, 0.5f, 0.5f), fmaf(pz, 0.5f, 0.5f), A.W, A.H, A.D, shx, shy, shz);
                        const float tn = __fadd_rn(t, A.tstep);
                        const bool cont = alive && !(tn > tfar) && (i + 1 < A.max_steps);
                        if (cont) {
                            t = tn; ++i;
                            px = __fadd_rn(px, stx); py = __fadd_rn(py, sty); pz = __fadd_rn(pz, stz);
                        }
                        alive = cont;
                    }
                    float4 col[U];
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        col[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid[k]) {
                            const float tu = (gather_blend<AXIS>(G[k]) - A.t_offset) * A.t_scale;
                            if (TFMODE == 0) col[k] = tex1D<float4>(A.tf_tex, tu);
                            else col[k] = tf_lookup_smem(tf_s, A.tf_n, tu);
                        }
                    }
    #pragma unroll
                    for (int k = 0; k < U; ++k) {
                        if (!valid[k]) { alive = false; break; }
                        if (COUNT) ++nsamp;
                        float4 c = col[k];
                        c.w *= A.density;
                        c.x *= c.w; c.y *= c.w; c.z *= c.w;
                        const float kk = 1.0f - sa;
                        sr += c.x * kk; sg += c.y * kk; sb += c.z * kk; sa += c.w * kk;
                        if (sa > A.thresh) { alive = false; break; }
                    }
                }
                A.out[(size_t)y * A.iw + x] = pack_rgba(sr * A.brightness, sg * A.brightness, sb * A.brightness, sa * A.brightness);
            } else if (A.clear_misses) {
                A.out[(size_t)y * A.iw + x] = 0u;
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nsamp += __shfl_xor_sync(0xffffffffu, nsamp, d);
        if (lane == 0 && nsamp) atomicAdd(A.samples, nsamp);
    }
    A.done.block_done();
}

// 3-D array (x fastest) -> layered copy: 32x32 tiles through shared memory, coalesced on both sides.
// AXIS 1: tiles over (x, y) at fixed z -> copy(x' = y, y' = z, layer = x);  grid = (ceil(W/32), ceil(H/32), D)
// AXIS 2: tiles over (x, z) at fixed y -> copy(x' = z, y' = x, layer = y);  grid = (ceil(W/32), ceil(D/32), H)
template <int AXIS>
__global__ void build_gather_copy_kernel(cudaSurfaceObject_t src, cudaSurfaceObject_t dst, int W, int H, int D) {
    __shared__ float tile[32][33];
    const int x0 = blockIdx.x * 32, r0 = blockIdx.y * 32, fixed = blockIdx.z;
    const int nrow = (AXIS == 1) ? H : D;
#pragma unroll
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int x = x0 + threadIdx.x, row = r0 + r;
        if (x < W && row < nrow)
            tile[r][threadIdx.x] = (AXIS == 1) ? surf3Dread<float>(src, x * 4, row, fixed) : surf3Dread<float>(src, x * 4, fixed, row);
    }
    __syncthreads();
#pragma unroll
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int row = r0 + threadIdx.x, x = x0 + r;                // now the row index is the fast one
        if (x < W && row < nrow) {
            if (AXIS == 1) surf2DLayeredwrite(tile[threadIdx.x][r], dst, row * 4, fixed, x);
            else surf2DLayeredwrite(tile[threadIdx.x][r], dst, row * 4, x, fixed);
        }
    }
}

// Persistent launch: as many blocks as the device holds at once (never more than there are items); the blocks share
// the items through the queue of FrameSignal.
// Distance between neighbouring rays at the centre of the volume, in voxels: the frame spans 4 world units there (eye 4
// away, image plane |u| <= 1 at distance 2), the box 2 world units over the volume's longest edge.
float ray_spacing_voxels(const vrdd_context* c, int iw, int ih) {
    const int n = c->W > c->H ? (c->W > c->D ? c->W : c->D) : (c->H > c->D ? c->H : c->D);
    const int m = iw > ih ? iw : ih;
    return m > 0 ? 2.0f * (float)n / (float)m : 0.0f;
}

template <class Kernel, class Args>
void launch_persistent(vrdd_context* c, Kernel kernel, const Args& A, long long items, size_t smem, int max_per_sm = 0) {
    static std::map<std::pair<const void*, size_t>, int> per_sm_cache;
    const std::pair<const void*, size_t> key(reinterpret_cast<const void*>(kernel), smem);
    auto it = per_sm_cache.find(key);
    if (it == per_sm_cache.end()) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        it = per_sm_cache.emplace(key, per_sm).first;
    }
    // var_persist_pct: resident blocks launched, in percent of what the device holds (100 = one wave of persistent blocks;
    // 0 = one block per item, i.e. the hardware's block scheduler instead of the queue)
    long long cap = (long long)it->second * c->num_sms * c->var_persist_pct / 100;
    if (c->var_persist_pct <= 0 || cap < 1) cap = items;
    // max_per_sm: fewer resident blocks than fit.  The 3-D array path is fastest with TWO blocks (16 warps) per SM: with
    // the five that fit, the rays of 1280 pixels march through one SM's L1 at once and evict one another's lines
    // (profiles/tuning_r2.md: oblique views 0.285 -> 0.259 ms, frontal views unchanged; more blocks under a register cap
    // are slower still).  The gather path is bound by instruction issue and wants all the warps it can get.
    // (not for small frames: with fewer than four items per block the tail of the last round costs more than the cap gains —
    // 512^2 frames: 0.048 ms uncapped, 0.051 ms with three blocks per SM)
    if (max_per_sm > 0 && c->var_persist_pct == 100 && (long long)max_per_sm * c->num_sms < cap && items >= 4ll * max_per_sm * c->num_sms)
        cap = (long long)max_per_sm * c->num_sms;
    kernel<<<(unsigned)(items < cap ? items : cap), kBlock, smem, c->stream>>>(A);
}

template <int SAMPLER, int TFMODE, int U>
void launch_u(vrdd_context* c, bool count, long long items, const RayArgs& A) {
    const size_t smem = (TFMODE == 1) ? sizeof(float4) * (size_t)A.tf_n : 0;
    // resident blocks per SM of the 3-D array kernel: two where the samples are sparse in the volume — neighbouring rays
    // about two voxels apart AND steps of several voxels (the headline: every DRAM line the rays cross is fetched, and five
    // blocks' rays evict one another's lines from L1) —, three otherwise (2048^2 frames, smaller volumes, the
    // resolution-matched step: the texture pipe is the limit and wants more requests); -1 = this rule, 0 = all that fit
    int per_sm = 0;
    if (SAMPLER == 0) {
        const int n = A.W > A.H ? (A.W > A.D ? A.W : A.D) : (A.H > A.D ? A.H : A.D);
        const bool sparse = ray_spacing_voxels(c, A.iw, A.ih) >= 1.5f && A.tstep * 0.5f * (float)n >= 2.5f;
        // ... and not looking along z (within 25 degrees): there whole slices are skipped, the traffic is 25 B per sample and the
        // texture pipe is the limit again (frontal views 0.145 ms with three blocks, 0.150 with two)
        const bool along_z = std::fabs(c->view[10]) >= 0.906f;
        per_sm = (c->var_array_blocks_per_sm >= 0) ? c->var_array_blocks_per_sm : ((sparse && !along_z) ? 2 : 3);
    }
    if (count) launch_persistent(c, raycast_kernel<SAMPLER, TFMODE, true, U>, A, items, smem, per_sm);
    else launch_persistent(c, raycast_kernel<SAMPLER, TFMODE, false, U>, A, items, smem, per_sm);
}

// The march batch U is a measured knob (vrdd_set_variant("raycast_unroll")) of the default path only: texture
// sampler, shared-memory transfer function; the two alternatives (transfer function through the texture unit, bricked
// manual sampler) exist for the comparison in DESIGN.md §4 and are built at U = 4.
template <int SAMPLER, int TFMODE>
void launch_variant(vrdd_context* c, bool count, long long items, const RayArgs& A, int unroll) {
    if (SAMPLER == 0 && TFMODE == 1) {
        switch (unroll) {
            case 1: launch_u<0, 1, 1>(c, count, items, A); return;
            case 2: launch_u<0, 1, 2>(c, count, items, A); return;
            case 8: launch_u<0, 1, 8>(c, count, items, A); return;
            default: launch_u<0, 1, 4>(c, count, items, A); return;
        }
    }
    launch_u<SAMPLER, TFMODE, 4>(c, count, items, A);
}

}  // namespace

// Host restatement of point_index_hw for the boundary coordinates fl(k / n): true when boundary k fetches
// texel min(k, n - 1) for every k in [0, n] (the precondition of the tld4 path of queryMethod 7).
bool point_rule_is_regular(int n) {
    for (int k = 0; k <= n; ++k) {
        float u = (float)k / (float)n;
        u = u < 0.f ? 0.f : (u > 1.f ? 1.f : u);
        const unsigned U = (unsigned)(u * 2097152.0f);
        unsigned long long i = ((unsigned long long)U * (unsigned)n) >> 21;
        if (i > (unsigned long long)(n - 1)) i = (unsigned long long)(n - 1);
        if ((int)i != (k < n ? k : n - 1)) return false;
    }
    return true;
}


long long make_tile_map(int iw, int ih, const vrdd_tile_partition& part, TileMap* tm) {
    if (iw <= 0 || ih <= 0 || part.parts < 1 || part.part < 0 || part.part >= part.parts || part.tile_w < 1 || part.tile_h < 1)
        return -1;
    tm->iw = iw; tm->ih = ih;
    tm->tile_w = part.tile_w; tm->tile_h = part.tile_h;
    tm->tiles_x = (iw + part.tile_w - 1) / part.tile_w;
    const int tiles_y = (ih + part.tile_h - 1) / part.tile_h;
    const long long ntiles = (long long)tm->tiles_x * tiles_y;
    tm->part = part.part; tm->parts = part.parts;
    const long long mine = (ntiles - part.part + part.parts - 1) / part.parts;
    tm->blocks_x = (part.tile_w + 15) / 16;
    tm->blocks_per_tile = tm->blocks_x * ((part.tile_h + 15) / 16);
    tm->n_items = 0;
    if (mine <= 0) return 0;
    const long long items = mine * tm->blocks_per_tile;
    if (items > 0x7fffffffLL) return -1;
    tm->n_items = (int)items;
    return items;
}

namespace {
// The wait gives up after 10 s (a rank that died must not hang the device); a frame is then simply not ordered.
__global__ void stream_wait_flag_kernel(const unsigned* flag, unsigned at_least, unsigned* post) {
    unsigned v;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while ((int)(v - at_least) < 0 && t1 - t0 < 10000000000ull);     // wrap-safe "v >= at_least"
    if (post) {                                                          // ... then publish (e.g. "frame consumed")
        __threadfence_system();
        asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(post) : "memory");
    }
}
__global__ void stream_post_flag_kernel(unsigned* flag) {
    __threadfence_system();
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
}
}  // namespace

int launch_stream_wait_flag(vrdd_context* c, const unsigned* d_flag, unsigned at_least, unsigned* d_post) {
    stream_wait_flag_kernel<<<1, 1, 0, c->stream>>>(d_flag, at_least, d_post);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_stream_post_flag(vrdd_context* c, unsigned* d_flag) {
    stream_post_flag_kernel<<<1, 1, 0, c->stream>>>(d_flag);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

void invalidate_gather_copies(vrdd_decoded_volume& v) {
    for (int i = 0; i < 3; ++i)
        for (int a = 0; a < 3; ++a) v.gvalid[i][a] = false;
}

namespace {

// The layered copy of plane `comp` with its layers stacked along `axis` (1 = x, 2 = y): created on first use, refilled
// from the 3-D array after every decode.  Returns VRDD_ERR_UNSUPPORTED when the extents do not fit a layered array.
int ensure_gather_copy(vrdd_context* c, vrdd_decoded_volume& v, int comp, int axis) {
    const int nrow = (axis == 1) ? c->H : c->D, ncol = (axis == 1) ? c->D : c->W, nlayer = (axis == 1) ? c->W : c->H;
    if (nlayer > 2048 || nrow > 32768 || ncol > 32768) return VRDD_ERR_UNSUPPORTED;
    if (!v.arr[comp] || !v.surf[comp]) return VRDD_ERR_UNSUPPORTED;
    if (!v.garr[comp][axis]) {
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        if (cudaMalloc3DArray(&v.garr[comp][axis], &desc, make_cudaExtent(nrow, ncol, nlayer),
                              cudaArrayLayered | cudaArraySurfaceLoadStore) != cudaSuccess) {
            cudaGetLastError();
            v.garr[comp][axis] = nullptr;
            return VRDD_ERR_UNSUPPORTED;                       // e.g. out of memory: the 3-D array path still works
        }
    }
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = v.garr[comp][axis];
    if (!v.gtex[comp][axis]) {
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        VRDD_CUDA(c, cudaCreateTextureObject(&v.gtex[comp][axis], &rd, &td, nullptr));
    }
    if (!v.gsurf[comp][axis]) VRDD_CUDA(c, cudaCreateSurfaceObject(&v.gsurf[comp][axis], &rd));
    if (!v.gvalid[comp][axis]) {
        const dim3 grid((c->W + 31) / 32, (nrow + 31) / 32, (axis == 1) ? c->D : c->H), block(32, 8);
        if (axis == 1) build_gather_copy_kernel<1><<<grid, block, 0, c->stream>>>(v.surf[comp], v.gsurf[comp][axis], c->W, c->H, c->D);
        else build_gather_copy_kernel<2><<<grid, block, 0, c->stream>>>(v.surf[comp], v.gsurf[comp][axis], c->W, c->H, c->D);
        c->launches += 1;
        VRDD_CUDA(c, cudaGetLastError());
        v.gvalid[comp][axis] = true;
    }
    return VRDD_OK;
}

// Which array serves this view: 0 = the 3-D array (lines thin in z, filtered by the texture unit: fewest instructions),
// 1 / 2 = the layered copy stacked along x / y (lines thin in x / y, filtered in the kernel).  Measured over the orbit
// (tools/bench_layouts.py, 1024^3, tstep 0.01): the copy stacked along x is the faster one from 45 degrees off the z
// axis on (0.234 against 0.243 ms at 45 degrees, 0.190 against 0.240 ms looking straight along x; 0.245 against 0.213 ms
// at 39 degrees) — it halves the lines touched when the view is within ~20 degrees of its axis and is no worse in
// between, where every line of the region the rays cross is touched whatever its shape.  So: the stacking axis the
// centre ray is most parallel to, if it is within acos(var_layout_cos) = 42 degrees of it and a step advances by more
// than var_layout_min_step voxels along it (otherwise no slice is ever skipped and the cheaper texture-unit path wins).
int choose_sector_axis(const vrdd_context* c, const vrdd_render_params& p, int iw, int ih) {
    if (c->var_layout == 1) return 0;
    if (c->var_layout == 2) return 1;
    if (c->var_layout == 3) return 2;
    // Rays one voxel apart or closer (a 2048^2 frame of the 1024^3 volume, the 512^3 volume at 1024^2) share their lines
    // across the frame whatever the layout: the texture-unit path wins on every view there (side view of 1024^3 at
    // 2048^2: 0.49 against 0.58 ms; of 512^3 at 1024^2: 0.139 against 0.195 ms).
    if (ray_spacing_voxels(c, iw, ih) < c->var_layout_min_spacing) return 0;
    const float dir[2] = {std::fabs(c->view[2]), std::fabs(c->view[6])};       // |x|, |y| of the centre ray (third column of M)
    const float adv[2] = {dir[0] * p.tstep * 0.5f * (float)c->W, dir[1] * p.tstep * 0.5f * (float)c->H};
    const int a = dir[0] >= dir[1] ? 0 : 1;
    if (dir[a] >= c->var_layout_cos && adv[a] > c->var_layout_min_step) return a + 1;
    return 0;
}

template <int AXIS, int TFMODE, bool POW2>
void launch_gather_u(vrdd_context* c, bool count, long long items, const RayArgs& A, int unroll) {
    const size_t smem = (TFMODE == 1) ? sizeof(float4) * (size_t)A.tf_n : 0;
    if (unroll <= 2) {
        if (count) launch_persistent(c, raycast_gather_kernel<AXIS, TFMODE, POW2, true, 2>, A, items, smem);
        else launch_persistent(c, raycast_gather_kernel<AXIS, TFMODE, POW2, false, 2>, A, items, smem);
    } else {
        if (count) launch_persistent(c, raycast_gather_kernel<AXIS, TFMODE, POW2, true, 4>, A, items, smem);
        else launch_persistent(c, raycast_gather_kernel<AXIS, TFMODE, POW2, false, 4>, A, items, smem);
    }
}

template <int AXIS>
void launch_gather(vrdd_context* c, bool count, long long items, RayArgs& A, int unroll, int tfm) {
    bool pow2 = true;
    const int ext[3] = {A.W, A.H, A.D};
    for (int i = 0; i < 3; ++i) {
        int m = 0;
        while ((1 << m) < ext[i]) ++m;
        if ((1 << m) != ext[i] || m > 12) pow2 = false;
        A.pow2_shift[i] = 13 - m;
    }
    // The transfer function always comes from the shared-memory table here: through the texture unit (a third fetch
    // per sample next to the two tld4) the same views take 0.29 instead of 0.20 ms.
    (void)tfm;
    if (pow2) launch_gather_u<AXIS, 1, true>(c, count, items, A, unroll);
    else launch_gather_u<AXIS, 1, false>(c, count, items, A, unroll);
}

}  // namespace

int launch_raycast(vrdd_context* c, uint32_t* d_out, int iw, int ih, const vrdd_render_params& p,
                   const vrdd_tile_partition& part, int clear_misses) {
    const int qm = p.query_method;
    if (qm == 8 || qm == 9 || qm == 0)                             // flexible blocks (:654-680)
        return launch_raycast_flex(c, d_out, iw, ih, p, part, clear_misses);
    if (qm == 7) {
        vrdd_decoded_volume& v0 = c->vol[VRDD_SRC_ORIGINAL];
        if (!v0.decoded || !v0.mean_raw)
            return fail(c, VRDD_ERR_INVALID, "render: queryMethod 7 needs vrdd_enable_interpolated_mean before the decode");
        if (iw <= 0 || ih <= 0 || !d_out) return fail(c, VRDD_ERR_INVALID, "render: bad image");
        Mode7Args A;
        const long long grid7 = make_tile_map(iw, ih, part, &A.tiles);
        if (grid7 < 0) return fail(c, VRDD_ERR_INVALID, "render: bad tile partition or image too large");
        if (grid7 == 0) return c->frame_signal ? launch_stream_post_flag(c, c->frame_signal) : VRDD_OK;
        A.done.flag = c->frame_signal; A.done.tickets = c->d_tickets;
        // fetch path: tld4 on the layered copy (default when it exists), else point fetches on the 3-D array,
        // else (or when asked for) plain loads from the linear plane — the same texels every time
        A.mean_raw = v0.mean_raw; A.mean_tex = (c->var_mode7 != 1) ? v0.mean_tex : 0; A.W = c->W; A.H = c->H; A.D = c->D;
        A.mean_gather = v0.mean_gather;
        A.tf_tab = reinterpret_cast<const float4*>(c->tf_dev); A.tf_n = c->tf_n;
        A.out = d_out; A.iw = iw; A.ih = ih;
        for (int i = 0; i < 12; ++i) A.m[i] = c->view[i];
        A.density = p.density; A.brightness = p.brightness; A.t_offset = p.transfer_offset; A.t_scale = p.transfer_scale;
        A.tstep = p.tstep; A.thresh = p.opacity_threshold; A.max_steps = p.max_steps; A.clear_misses = clear_misses;
        A.samples = c->d_samples;
        A.ref_rounding = c->var_ray_setup;
        const size_t tab_bytes = sizeof(float2) * ((size_t)c->W + c->H + c->D + 9);
        A.use_tab = (c->W <= 8192 && c->H <= 8192 && c->D <= 8192 && tab_bytes <= 48 * 1024) ? 1 : 0;   // div_small's checked range
        A.idx32 = ((unsigned long long)c->W * c->H * c->D < (1ull << 31)) ? 1 : 0;
        const size_t smem7 = sizeof(float4) * (size_t)A.tf_n + (A.use_tab ? tab_bytes : 0);
        // tld4 path: asked for, its layered copy exists, and the point rule maps every x / y boundary to itself
        const bool gather = c->var_mode7 == 2 && v0.mean_gather && A.use_tab;   // regular x / y rule: checked when it was made
        const bool cnt = c->count_samples && c->d_samples;
        auto k7 = gather ? (cnt ? raycast_mode7_kernel<true, 4, true> : raycast_mode7_kernel<false, 4, true>)
                         : (cnt ? raycast_mode7_kernel<true, 4, false> : raycast_mode7_kernel<false, 4, false>);
        VRDD_CUDA(c, cudaFuncSetAttribute(k7, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));   // <= 16 KB table + 48 KB cells
        launch_persistent(c, k7, A, grid7, smem7);
        c->launches += 1;
        VRDD_CUDA(c, cudaGetLastError());
        return VRDD_OK;
    }
    if (qm < 1 || qm > 6)
        return fail(c, VRDD_ERR_UNSUPPORTED, "render: queryMethod must be 0..9 (volumeRender.cpp:129)");
    const int source = (qm >= 4) ? VRDD_SRC_FRACTAL : VRDD_SRC_ORIGINAL;
    const int comp = (qm - 1) % 3;
    vrdd_decoded_volume& vol = c->vol[source];
    if (!vol.decoded) return fail(c, VRDD_ERR_INVALID, "render: the requested volume has not been decoded");
    if (iw <= 0 || ih <= 0 || !d_out) return fail(c, VRDD_ERR_INVALID, "render: bad image");

    RayArgs A;
    A.vol_tex = vol.tex[comp];
    A.vol_brick = vol.brick[comp];
    A.W = c->W; A.H = c->H; A.D = c->D; A.bW = c->bW; A.bH = c->bH;
    A.tf_tex = c->tf_tex; A.tf_tab = reinterpret_cast<const float4*>(c->tf_dev); A.tf_n = c->tf_n;
    A.out = d_out; A.iw = iw; A.ih = ih;
    for (int i = 0; i < 12; ++i) A.m[i] = c->view[i];
    A.density = p.density; A.brightness = p.brightness; A.t_offset = p.transfer_offset;
    A.t_scale = p.transfer_scale; A.tstep = p.tstep; A.thresh = p.opacity_threshold; A.max_steps = p.max_steps;
    const long long grid = make_tile_map(iw, ih, part, &A.tiles);
    if (grid < 0) return fail(c, VRDD_ERR_INVALID, "render: bad tile partition or image too large");
    A.done.flag = c->frame_signal; A.done.tickets = c->d_tickets;
    A.clear_misses = clear_misses;
    A.ref_rounding = c->var_ray_setup;
    A.samples = c->d_samples;
    if (grid == 0) return c->frame_signal ? launch_stream_post_flag(c, c->frame_signal) : VRDD_OK;   // no tile of mine: still counted

    if (c->sampler == VRDD_SAMPLER_LINEAR)
        return fail(c, VRDD_ERR_INVALID, "render: VRDD_SAMPLER_LINEAR volumes are rendered with vrdd_render_brick_*");
    const bool count = c->count_samples && c->d_samples;
    const int sampler = c->sampler, tfm = (c->sampler == VRDD_SAMPLER_BRICKED) ? 1 : c->var_tf;
    if (sampler == VRDD_SAMPLER_TEXTURE && !A.vol_tex) return fail(c, VRDD_ERR_INVALID, "render: no texture volume");
    if (sampler == VRDD_SAMPLER_BRICKED && !A.vol_brick) return fail(c, VRDD_ERR_INVALID, "render: no bricked volume");
    if (sampler == VRDD_SAMPLER_TEXTURE) {
        int axis = choose_sector_axis(c, p, iw, ih);
        if (axis != 0) {
            const int rc = ensure_gather_copy(c, vol, comp, axis);
            if (rc == VRDD_ERR_UNSUPPORTED && c->var_layout == 0) axis = 0;       // auto: the 3-D array always works
            else if (rc != VRDD_OK) return rc == VRDD_ERR_UNSUPPORTED ? fail(c, rc, "render: no layered copy for this volume") : rc;
        }
        if (axis != 0) {
            A.vol_tex = vol.gtex[comp][axis];
            const int gtf = (c->var_gather_tf >= 0) ? c->var_gather_tf : tfm;
            if (axis == 1) launch_gather<1>(c, count, grid, A, c->var_gather_unroll, gtf);
            else launch_gather<2>(c, count, grid, A, c->var_gather_unroll, gtf);
        } else if (tfm == 0) launch_variant<0, 0>(c, count, grid, A, c->var_unroll);
        else launch_variant<0, 1>(c, count, grid, A, c->var_unroll);
    } else {
        launch_variant<1, 1>(c, count, grid, A, c->var_unroll);
    }
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

#ifdef VRDD_PROBE_EXPORTS
int launch_debug_sample(vrdd_context* c, cudaTextureObject_t tex, const float* d_uvw, int n, float* d_out) {
    if (n <= 0) return VRDD_OK;
    debug_sample_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(tex, d_uvw, n, d_out);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_debug_sample_tf(vrdd_context* c, const float* d_u, int n, float* d_out4) {
    if (n <= 0) return VRDD_OK;
    debug_sample_tf_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->tf_tex, d_u, n, reinterpret_cast<float4*>(d_out4));
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

#endif  // VRDD_PROBE_EXPORTS

}  // namespace vrdd
