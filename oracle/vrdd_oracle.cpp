/*
 * vrdd_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain C++/OpenMP restatement of the two hot paths of
 * ykou/Volume-Rendering-Based-on-Distribution-Data:
 *   P1  distribution decode   (volumeRender_kernel.cu:195-222, 722-872)
 *   P2  volume ray casting    (volumeRender_kernel.cu:136-193, 272-717)
 * plus the texture-unit semantics those kernels lean on (modes set at
 * volumeRender_kernel.cu:1865-1876, 2161-2171, 2337-2339) and the view matrix the host
 * builds with OpenGL (volumeRender.cpp:224-246, 1024-1043).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library, and only as the checker or the timed CPU baseline.  The product
 * (libvrdd.so) never links, loads or calls it.
 *
 * PARITY PINNED BY THE REFERENCE'S OWN DEVICE CODE, RUN ON A B200.  The reference ships no input data, no
 * reference image and no unit tests, and as it stands it does not compile with CUDA 12.9
 * (texture references were removed; helper_math.h / helper_cuda.h are not vendored).  oracle/ref_shim gives CUDA 12
 * those spellings back on top of texture objects, oracle/ref_driver.cu includes volumeRender_kernel.cu where it
 * lies and replays the reference's call sequence (initCuda, basicDataProcessing, copyInvViewMatrix,
 * render_kernel) on seeded inputs (oracle/Makefile `ref`, tools/ref_pin.py); what it computed on the GPU is
 * tests/golden/ref_gpu_v1.npz.  tests/test_reference_pin.py holds this restatement to it:
 *   - raw-histogram decode: the mean bit for bit, variance and entropy to 3e-7;
 *   - d_render, queryMethod 1..3: every byte of two 256x256 frames within 1 LSB, at most 12 bytes differing;
 *   - fractal decode: every voxel to 1e-6, and the frames of queryMethod 4..6, once the stores the reference's
 *     build drops are modelled — its fractalDecoding() returns a pointer to a local array (:196-221), and nvcc 12.9
 *     keeps only the stores of template bins 0..22 (unflipped) resp. 8..31 (flipped) into it
 *     (tools/ref_pin.as_the_reference_build_decodes); this restatement keeps the source's intent, all 32 bins;
 *   - queryMethod 7 is discontinuous at cell boundaries, so the last bit of a sample position decides single
 *     samples: see g_fma_contract below.
 * Besides that it is checked against hand-computed known answers and its own invariants (tests/test_oracle.py), and
 * the texture-filter model (WQ_HW below) was fitted to, and is checked against, the B200 texture unit itself
 * (tools/probe_texture*.py, tests/test_gpu_render.py::test_texture_unit_matches_the_filter_model).
 *
 * Arithmetic that lives outside /root/reference and is restated from its public definition:
 *   helper_math.h (CUDA Samples common/inc, CUDA 5.0 era, no pin file):
 *       dot = a.x*b.x + a.y*b.y + a.z*b.z (+ a.w*b.w), left to right;
 *       normalize(v) = v * rsqrtf(dot(v,v))  -> restated as v * (1/sqrtf(dot)), IEEE;
 *       float3/float4 operators are componentwise.
 *   Texture filtering (CUDA C Programming Guide, "Texture Fetching"):
 *       point:  T[floor(x)];
 *       linear: xB = x - 0.5, i = floor(xB), a = frac(xB) held in 9-bit fixed point with
 *       8 fractional bits; normalised coordinates are multiplied by the extent first;
 *       clamp addressing clamps i and i+1 into [0, N-1].
 *     The guide does not say how the fixed-point coordinate is formed or how the eight
 *     trilinear weights are combined; d_render's result depends on both.  WQ_HW restates
 *     what the texture unit of the B200 actually does, measured with one-hot and ramp
 *     volumes (0 mismatches in ~300 000 weight probes, <= 3e-8 on random volumes):
 *       - the normalised coordinate is clamped to [0,1] and TRUNCATED to 21 fractional bits,
 *         U = floor(u * 2^21);
 *       - q = ((U * N * 256 + 2^20) >> 21) - 128 is the texel-centre-relative coordinate in
 *         1/256 texels, clamped to [0, (N-1)*256];  i = q >> 8,  A = q & 255;
 *       - the eight weights are integers out of 256 that always sum to 256, split z, then x,
 *         then y:  Z1 = C, Z0 = 256 - C;  X1 = (Z*A + 128) >> 8, X0 = Z - X1;
 *         on X0: Y0 = (X0*(256-B) + 128) >> 8, Y1 = X0 - Y0;  on X1: Y1 = (X1*B + 128) >> 8,
 *         Y0 = X1 - Y1.
 *
 * Build with -ffp-contract=off: every float operation below rounds once, in source
 * order, so the oracle is a fixed point of itself on any host.
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include <array>
#include <map>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vrdd_synth.h"

namespace {

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

/* ---- texture-unit emulation ---------------------------------------------------- */

enum WeightQuant {
    WQ_EXACT = 0,   /* fp32 weights, no quantisation                                        */
    WQ_ROUND8 = 1,  /* the programming guide's prose: per-axis 8-bit weight, nearest        */
    WQ_TRUNC8 = 2,  /* same, truncated                                                      */
    WQ_HW = 3       /* the measured texture-unit scheme described in the header (default)   */
};

/* WQ_HW: texel index and 8-bit weight of one axis */
inline void split_hw(float u, int N, int* i, int* a256) {
    float uc = fminf(fmaxf(u, 0.0f), 1.0f);                 /* NaN -> 0 */
    long long U = (long long)std::floor(uc * 2097152.0f);  /* 2^21; the scaling is exact */
    long long q = ((U * (long long)N * 256 + (1LL << 20)) >> 21) - 128;
    long long qmax = (long long)(N - 1) * 256;
    if (q < 0) q = 0;
    if (q > qmax) q = qmax;
    *i = (int)(q >> 8);
    *a256 = (int)(q & 255);
}

/* WQ_HW: the eight integer weights, index = x | y<<1 | z<<2 */
inline void weights_hw(int A, int B, int C, int* w) {
    const int Z[2] = {256 - C, C};
    for (int z = 0; z < 2; ++z) {
        int x1 = (Z[z] * A + 128) >> 8, x0 = Z[z] - x1;
        int y0 = (x0 * (256 - B) + 128) >> 8;              /* lower-x column: the lower-y weight is rounded */
        int y1 = (x1 * B + 128) >> 8;                      /* upper-x column: the upper-y weight is rounded */
        w[0 | 0 | (z << 2)] = y0;       w[0 | 2 | (z << 2)] = x0 - y0;
        w[1 | 0 | (z << 2)] = x1 - y1;  w[1 | 2 | (z << 2)] = y1;
    }
}

/* Split an unnormalised linear-filter coordinate into (i, alpha).  CUDA Programming
 * Guide "Linear Filtering": xB = x - 0.5; i = floor(xB); alpha = frac(xB) in 1.8 fixed point. */
inline void split_linear(float x_unnorm, int wq, int* i, float* a) {
    float xb = x_unnorm - 0.5f;
    if (wq == WQ_EXACT) {
        float fl = std::floor(xb);
        *i = (int)fl;
        *a = xb - fl;
        return;
    }
    /* fixed-point coordinate with 8 fractional bits */
    float scaled = xb * 256.0f;
    float q = (wq == WQ_ROUND8) ? std::floor(scaled + 0.5f) : std::floor(scaled);
    long long qi = (long long)q;
    long long ii = qi >> 8;                     /* arithmetic shift == floor division */
    *i = (int)ii;
    *a = (float)(qi - (ii << 8)) * (1.0f / 256.0f);
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* tex3D on a float4 volume, linear, normalised, clamp (originalQueryTex/fractalQueryTex,
 * volumeRender_kernel.cu:1865-1876), returning one component. */
struct Volume4 {
    const float* data;   /* float4[V], x fastest (volumeRender_kernel.cu:740, 771-773) */
    int W, H, D;
    int stride = 4;      /* floats per texel: 4 = the reference's float4 volume; 1 = ONE component as a plane (the headline
                            sizes: a 1024^3 float4 volume is 17 GB, its sampled component 4.3 GB) */
};

inline float tex3d_linear_comp(const Volume4& v, int comp, float u, float vv, float w, int wq) {
    if (wq == WQ_HW) {
        int i, j, k, A, B, C, wt[8];
        split_hw(u, v.W, &i, &A);
        split_hw(vv, v.H, &j, &B);
        split_hw(w, v.D, &k, &C);
        weights_hw(A, B, C, wt);
        float acc = 0.0f;
        for (int n = 0; n < 8; ++n) {
            if (wt[n] == 0) continue;                       /* also keeps i+1 == N out of reach */
            int x = i + (n & 1), y = j + ((n >> 1) & 1), z = k + (n >> 2);
            float t = v.data[(size_t)v.stride * ((size_t)x + (size_t)v.W * ((size_t)y + (size_t)v.H * (size_t)z)) + comp];
            acc = acc + ((float)wt[n] * (1.0f / 256.0f)) * t;
        }
        return acc;
    }
    int i, j, k; float a, b, c;
    split_linear(u * (float)v.W, wq, &i, &a);
    split_linear(vv * (float)v.H, wq, &j, &b);
    split_linear(w * (float)v.D, wq, &k, &c);
    int i0 = clampi(i, 0, v.W - 1), i1 = clampi(i + 1, 0, v.W - 1);
    int j0 = clampi(j, 0, v.H - 1), j1 = clampi(j + 1, 0, v.H - 1);
    int k0 = clampi(k, 0, v.D - 1), k1 = clampi(k + 1, 0, v.D - 1);
    auto T = [&](int x, int y, int z) -> float {
        return v.data[(size_t)v.stride * ((size_t)x + (size_t)v.W * ((size_t)y + (size_t)v.H * (size_t)z)) + comp];
    };
    /* (1-a)(1-b)(1-c) T000 + ... evaluated as nested lerps, x then y then z */
    float oa = 1.0f - a, ob = 1.0f - b, oc = 1.0f - c;
    float x00 = oa * T(i0, j0, k0) + a * T(i1, j0, k0);
    float x10 = oa * T(i0, j1, k0) + a * T(i1, j1, k0);
    float x01 = oa * T(i0, j0, k1) + a * T(i1, j0, k1);
    float x11 = oa * T(i0, j1, k1) + a * T(i1, j1, k1);
    float y0 = ob * x00 + b * x10;
    float y1 = ob * x01 + b * x11;
    return oc * y0 + c * y1;
}

/* tex1D on the float4 transfer function, linear, normalised, clamp
 * (transferTex, volumeRender_kernel.cu:2337-2339) */
inline f4 tex1d_linear4(const float* tf, int n, float u, int wq) {
    if (wq == WQ_HW) {
        int i, A;
        split_hw(u, n, &i, &A);
        int i1 = (i + 1 < n) ? i + 1 : i;
        float w0 = (float)(256 - A) * (1.0f / 256.0f), w1 = (float)A * (1.0f / 256.0f);
        f4 r;
        r.x = w0 * tf[4 * i + 0] + w1 * tf[4 * i1 + 0];
        r.y = w0 * tf[4 * i + 1] + w1 * tf[4 * i1 + 1];
        r.z = w0 * tf[4 * i + 2] + w1 * tf[4 * i1 + 2];
        r.w = w0 * tf[4 * i + 3] + w1 * tf[4 * i1 + 3];
        return r;
    }
    int i; float a;
    split_linear(u * (float)n, wq, &i, &a);
    int i0 = clampi(i, 0, n - 1), i1 = clampi(i + 1, 0, n - 1);
    float oa = 1.0f - a;
    f4 r;
    r.x = oa * tf[4 * i0 + 0] + a * tf[4 * i1 + 0];
    r.y = oa * tf[4 * i0 + 1] + a * tf[4 * i1 + 1];
    r.z = oa * tf[4 * i0 + 2] + a * tf[4 * i1 + 2];
    r.w = oa * tf[4 * i0 + 3] + a * tf[4 * i1 + 3];
    return r;
}

/* ---- P1: decode ------------------------------------------------------------------ */

/* Statistics of a raw histogram, volumeRender_kernel.cu:736-769.
 * Types follow the reference expression by expression:
 *   mean     += p * (binWidth*i + binWidth/2.0)        -> double term, float accumulator   (:746)
 *   variance += p * ((i/B)*Max - mean) * (...)         -> all float, bin LEFT edge         (:753-754)
 *   mean /= 0.0217; variance /= 0.000021               -> double divide                    (:758-759)
 *   entropy  += p * (p<=0 ? 0 : log(p)/log(2.0))       -> logf, double divide              (:765-766)
 *   entropy   = -entropy / (logf(B)/logf(2))                                               (:768-769) */
inline void stats_original(const float* p, int B, float* out3) {
    float MaxHistogram = 0.0217;
    float MinHistogram = 0.0;
    float binWidth = (MaxHistogram - MinHistogram) / (float)B;
    float mean = 0;
    for (int i = 0; i < B; ++i)
        mean = (float)((double)mean + (double)p[i] * ((double)(binWidth * (float)i) + (double)binWidth / 2.0));
    float variance = 0;
    for (int i = 0; i < B; ++i) {
        float d = ((float)i / (float)B) * MaxHistogram - mean;
        variance = variance + (p[i] * d) * d;
    }
    mean = (float)((double)mean / 0.0217);
    variance = (float)((double)variance / 0.000021);
    float entropy = 0;
    for (int i = 0; i < B; ++i) {
        double term = (p[i] <= 0) ? 0.0 : ((double)logf(p[i]) / std::log(2.0));
        entropy = (float)((double)entropy + (double)p[i] * term);
    }
    entropy = -entropy;
    entropy = entropy / (logf((float)B) / logf(2.0f));
    out3[0] = mean; out3[1] = variance; out3[2] = entropy;
}

/* Optional reversal then circular right shift, volumeRender_kernel.cu:195-222.
 * Single wrap only: valid for 0 <= shift <= B (the reference indexes out of bounds
 * otherwise); callers validate. */
inline void fractal_transform(const float* tmpl, int B, int flip, int shift, float* cur) {
    for (int i = 0; i < B; ++i) {
        int m = i + shift;
        if (m >= B) m -= B;
        cur[m] = (flip == 0) ? tmpl[i] : tmpl[B - 1 - i];
    }
}

/* Statistics of a reconstructed histogram, volumeRender_kernel.cu:841-867.  Here the
 * variance uses the bin CENTRE in double (:851-853), unlike stats_original. */
inline void stats_fractal(const float* cur, int B, float* out3) {
    float MaxHistogram1 = 0.0217;
    float MinHistogram1 = 0.0;
    float mean1 = 0;
    float binWidth1 = (MaxHistogram1 - MinHistogram1) / (float)B;
    for (int i = 0; i < B; ++i)
        mean1 = (float)((double)mean1 + (double)cur[i] * ((double)(binWidth1 * (float)i) + (double)binWidth1 / 2.0));
    float variance1 = 0;
    for (int i = 0; i < B; ++i) {
        double d = ((double)(binWidth1 * (float)i) + (double)binWidth1 / 2.0) - (double)mean1;
        variance1 = (float)((double)variance1 + ((double)cur[i] * d) * d);
    }
    mean1 = (float)((double)mean1 / 0.0217);
    variance1 = (float)((double)variance1 / 0.000021);
    float entropy1 = 0;
    for (int i = 0; i < B; ++i) {
        double term = (cur[i] <= 0) ? 0.0 : ((double)logf(cur[i]) / std::log(2.0));
        entropy1 = (float)((double)entropy1 + (double)cur[i] * term);
    }
    entropy1 = -entropy1;
    entropy1 = entropy1 / (logf((float)B) / logf(2.0f));
    out3[0] = mean1; out3[1] = variance1; out3[2] = entropy1;
}

/* ---- P2: ray casting -------------------------------------------------------------- */

inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* slab test, volumeRender_kernel.cu:136-156 */
inline int intersect_box(f3 o, f3 d, f3 bmin, f3 bmax, float* tnear, float* tfar) {
    f3 invR = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    f3 tbot = {invR.x * (bmin.x - o.x), invR.y * (bmin.y - o.y), invR.z * (bmin.z - o.z)};
    f3 ttop = {invR.x * (bmax.x - o.x), invR.y * (bmax.y - o.y), invR.z * (bmax.z - o.z)};
    f3 tmin = {fminf(ttop.x, tbot.x), fminf(ttop.y, tbot.y), fminf(ttop.z, tbot.z)};
    f3 tmax = {fmaxf(ttop.x, tbot.x), fmaxf(ttop.y, tbot.y), fmaxf(ttop.z, tbot.z)};
    float largest_tmin = fmaxf(fmaxf(tmin.x, tmin.y), fmaxf(tmin.x, tmin.z));
    float smallest_tmax = fminf(fminf(tmax.x, tmax.y), fminf(tmax.x, tmax.z));
    *tnear = largest_tmin;
    *tfar = smallest_tmax;
    return smallest_tmax > largest_tmin;
}

inline float saturate(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }  /* NaN -> 0 like __saturatef */

/* How the ray set-up and the compositing are ROUNDED.  The source (:282-312, :695) does not say: it leaves the
 * contraction of a*b + c into one fused multiply-add to the compiler.  0 (default): no contraction, 1/sqrtf — the
 * order the CUDA kernels of this repository reproduce bit for bit.  1: what nvcc 12.9 makes of the reference's
 * d_render with its default -fmad=true, read off the PTX of the reference compiled where it lies
 * (oracle/Makefile `ref`): u*u + v*v fused, + 4 added, rsqrt; each component of M*dir as
 * fma(dir.z, m.z, fma(dir.x, m.x, dir.y*m.y)); pos = fma(d, tnear, o); sum = fma(col, 1 - sum.w, sum).  rsqrt is
 * the GPU's rsqrt.approx.f32 looked up in a table dumped on a B200 (g_rsqrt_delta below; the correctly rounded
 * 1/sqrt, within two ulps of it, when no table is installed).  The two settings give
 * the same frames to +-1 LSB in queryMethod 1..6; queryMethod 7 amplifies the last bit of a position wherever a
 * sample sits on a cell boundary (its "vertical and horizontal line" artefact, ver1.9.6.txt:166), and only
 * setting 1 reproduces the frames of the reference's own binary there (tests/test_reference_pin.py). */
static int g_fma_contract = 0;
/* rsqrt.approx.f32 as the B200 computes it, for every float of [lo, hi] (the argument u*u + v*v + 4 of the eye ray's
 * normalize() lies in [4, 6]): distance in ulps from (float)(1.0 / sqrt((double)x)), dumped on the GPU by
 * tools/rsqrt_dump.cu (tests/golden/rsqrt_approx_b200_v1.npz).  With it setting 1 reproduces the ray set-up of the
 * reference's own binary BIT FOR BIT; without it (or outside the interval) the correctly rounded value stands in. */
static std::vector<int8_t> g_rsqrt_delta;
static uint32_t g_rsqrt_lo = 0;
inline float rsqrt_as_the_gpu(float x) {
    float r = (float)(1.0 / std::sqrt((double)x));
    uint32_t xb, rb;
    std::memcpy(&xb, &x, 4);
    if (!g_rsqrt_delta.empty() && xb >= g_rsqrt_lo && (size_t)(xb - g_rsqrt_lo) < g_rsqrt_delta.size()) {
        std::memcpy(&rb, &r, 4);
        rb = (uint32_t)((int64_t)rb + g_rsqrt_delta[xb - g_rsqrt_lo]);
        std::memcpy(&r, &rb, 4);
    }
    return r;
}

struct RaySetup { f3 o, d; };
inline RaySetup make_ray(const float* m12, int x, int y, int imageW, int imageH) {
    RaySetup R;
    float u = ((float)x / (float)imageW) * 2.0f - 1.0f;                    /* :288 (the same value fused or not) */
    float v = ((float)y / (float)imageH) * 2.0f - 1.0f;                    /* :289 */
    R.o = {m12[3], m12[7], m12[11]};               /* mul(M, (0,0,0,1)) == the translation column exactly (:293-294) */
    f3 r0 = {m12[0], m12[1], m12[2]}, r1 = {m12[4], m12[5], m12[6]}, r2 = {m12[8], m12[9], m12[10]};
    if (!g_fma_contract) {
        f3 dv = {u, v, -2.0f};                                             /* :295 */
        float inv_len = 1.0f / sqrtf(dot3(dv, dv));
        dv.x = dv.x * inv_len; dv.y = dv.y * inv_len; dv.z = dv.z * inv_len;
        R.d = {dot3(dv, r0), dot3(dv, r1), dot3(dv, r2)};                  /* :296, 168-174 */
    } else {
        float dd = fmaf(u, u, v * v) + 4.0f;
        float inv_len = rsqrt_as_the_gpu(dd);
        f3 dv = {u * inv_len, v * inv_len, inv_len * -2.0f};
        R.d = {fmaf(dv.z, r0.z, fmaf(dv.x, r0.x, dv.y * r0.y)), fmaf(dv.z, r1.z, fmaf(dv.x, r1.x, dv.y * r1.y)),
               fmaf(dv.z, r2.z, fmaf(dv.x, r2.x, dv.y * r2.y))};
    }
    return R;
}
inline f3 first_pos(f3 o, f3 d, float tnear) {                             /* :311 */
    if (g_fma_contract) return {fmaf(d.x, tnear, o.x), fmaf(d.y, tnear, o.y), fmaf(d.z, tnear, o.z)};
    return {o.x + d.x * tnear, o.y + d.y * tnear, o.z + d.z * tnear};
}
inline void composite(f4& sum, f4 col) {                                   /* :695 */
    float k = 1.0f - sum.w;
    if (g_fma_contract) {
        sum.x = fmaf(col.x, k, sum.x); sum.y = fmaf(col.y, k, sum.y); sum.z = fmaf(col.z, k, sum.z); sum.w = fmaf(col.w, k, sum.w);
    } else {
        sum.x = sum.x + col.x * k; sum.y = sum.y + col.y * k; sum.z = sum.z + col.z * k; sum.w = sum.w + col.w * k;
    }
}

/* volumeRender_kernel.cu:186-193: saturate, x255, truncate, pack A<<24|B<<16|G<<8|R */
inline uint32_t pack_rgba(f4 c) {
    float r = saturate(c.x), g = saturate(c.y), b = saturate(c.z), a = saturate(c.w);
    return ((uint32_t)(a * 255) << 24) | ((uint32_t)(b * 255) << 16) | ((uint32_t)(g * 255) << 8) |
           (uint32_t)(r * 255);
}

}  // namespace

extern "C" {

struct vrdd_oracle_render_params {
    int image_w, image_h;
    float density, brightness, transfer_offset, transfer_scale;
    int query_method;        /* 1,2,3: component 0,1,2 of vol_original; 4,5,6: of vol_fractal */
    float tstep;             /* reference: 0.01f      (volumeRender_kernel.cu:277) */
    int max_steps;           /* reference: 500        (:276) */
    float opacity_threshold; /* reference: 0.95f      (:278) */
    int weight_quant;        /* 3 = the measured B200 texture-unit filter (default); 0/1/2 = textbook variants */
    int y0, y1;              /* rows [y0, y1) to render (whole image: 0, image_h) */
};

/* 0 / 1: see g_fma_contract.  Process-wide; returns the previous setting. */
int vrdd_oracle_set_fma_contract(int on) { const int prev = g_fma_contract; g_fma_contract = on ? 1 : 0; return prev; }

/* Installs (n > 0) or removes (n == 0) the rsqrt.approx table: delta[i] belongs to the float with bit pattern lo_bits + i. */
void vrdd_oracle_set_rsqrt_table(uint32_t lo_bits, int64_t n, const int8_t* delta) {
    g_rsqrt_lo = lo_bits;
    if (n > 0 && delta) g_rsqrt_delta.assign(delta, delta + n);
    else g_rsqrt_delta.clear();
}

int vrdd_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void vrdd_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* a1 — raw-histogram half of d_basicDataProcessing (volumeRender_kernel.cu:722-773).
 * hist: float[V][B]; out: float4[V] = (mean, variance, entropy, 0). */
void vrdd_oracle_decode_hist(const float* hist, int64_t V, int B, float* out4) {
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < V; ++v) {
        float s[3];
        stats_original(hist + v * B, B, s);
        out4[4 * v + 0] = s[0]; out4[4 * v + 1] = s[1]; out4[4 * v + 2] = s[2]; out4[4 * v + 3] = 0.0f;
    }
}

/* a2 — fractal half of d_basicDataProcessing (volumeRender_kernel.cu:775-871).
 * codebook: int4[V] = (templateId, shift, flip, NE)           (:777-786)
 * errors:   float2[V][B] dense, first NE entries valid        (:806-825, volumeRender.cpp:582)
 * templates: float[T][B]; the reference fetches layer 1 of a 1-layer array, which the
 *            hardware clamps to layer 0 (:792-793)
 * recon (optional): float[V][B], the histogram after errors+clamp and BEFORE
 *            normalisation — every value is a permuted template entry plus exactly the
 *            additions the reference performs in order, so it is bit-comparable.
 * Returns the number of voxels whose code violates the reference's run-time guards. */
int64_t vrdd_oracle_decode_fractal(const int32_t* codebook, const float* errors, const float* templates,
                                   int T, int64_t V, int B, float* out4, float* recon) {
    int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (int64_t v = 0; v < V; ++v) {
        float cur[VRDD_SYNTH_MAX_BINS];
        int id = codebook[4 * v + 0], shift = codebook[4 * v + 1], flip = codebook[4 * v + 2],
            ne = codebook[4 * v + 3];
        if (id < 0 || id >= T || shift < 0 || shift > B || ne < 0 || ne > B) {
            bad += 1;
            out4[4 * v + 0] = out4[4 * v + 1] = out4[4 * v + 2] = out4[4 * v + 3] = 0.0f;
            if (recon) std::memset(recon + v * B, 0, sizeof(float) * B);
            continue;
        }
        fractal_transform(templates + (size_t)id * B, B, flip, shift, cur);
        for (int k = 0; k < ne; ++k) {                                    /* :806-825 */
            int bin = (int)errors[2 * (v * B + k) + 0];
            float val = errors[2 * (v * B + k) + 1];
            if (bin < 0 || bin >= B) { bad += 1; continue; }
            cur[bin] = cur[bin] + val;
            if (cur[bin] < 0) cur[bin] = 0;
        }
        if (recon) std::memcpy(recon + v * B, cur, sizeof(float) * B);
        float tot = 0;                                                    /* :828-839 */
        for (int i = 0; i < B; ++i) tot = tot + cur[i];
        if (tot > 0)
            for (int i = 0; i < B; ++i) cur[i] = cur[i] / tot;
        float s[3];
        stats_fractal(cur, B, s);
        out4[4 * v + 0] = s[0]; out4[4 * v + 1] = s[1]; out4[4 * v + 2] = s[2]; out4[4 * v + 3] = 0.0f;
    }
    return bad;
}

/* The hard-coded nine-entry rainbow, volumeRender_kernel.cu:2323-2326 */
void vrdd_oracle_default_transfer_function(float* tf4 /* [9*4] */) {
    static const float k[9][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {1, 0.5f, 0, 1}, {1, 1, 0, 1}, {0, 1, 0, 1},
                                  {0, 1, 1, 1}, {0, 0, 1, 1}, {1, 0, 1, 1}, {0, 0, 0, 0}};
    std::memcpy(tf4, k, sizeof(k));
}

/* Inverse view matrix without OpenGL (volumeRender.cpp:224-246).
 * GL post-multiplies: M = Rx(-rot.x) * Ry(-rot.y) * T(-trans); the host keeps the top
 * three rows of M row-major with the translation in the 4th column.  Angles in degrees.
 * rot = 0, trans = (0,0,-4) gives the self-test matrix of volumeRender.cpp:1024-1043. */
void vrdd_oracle_view_matrix(float rot_x_deg, float rot_y_deg, float tx, float ty, float tz, float* m12) {
    const double d2r = 3.14159265358979323846 / 180.0;
    double ax = -(double)rot_x_deg * d2r, ay = -(double)rot_y_deg * d2r;
    float cx = (float)std::cos(ax), sx = (float)std::sin(ax), cy = (float)std::cos(ay), sy = (float)std::sin(ay);
    /* Rx = [1 0 0; 0 cx -sx; 0 sx cx],  Ry = [cy 0 sy; 0 1 0; -sy 0 cy] */
    float R[3][3] = {{cy, 0.0f, sy}, {sx * sy, cx, -(sx * cy)}, {-(cx * sy), sx, cx * cy}};
    float t[3] = {-tx, -ty, -tz};
    for (int r = 0; r < 3; ++r) {
        m12[4 * r + 0] = R[r][0]; m12[4 * r + 1] = R[r][1]; m12[4 * r + 2] = R[r][2];
        m12[4 * r + 3] = R[r][0] * t[0] + R[r][1] * t[1] + R[r][2] * t[2];
    }
}

/* d_render for queryMethod 1..6 (volumeRender_kernel.cu:272-312, 381-387, 601-651, 683-716).
 * The unconditional 8x33-fetch prologue (:354-367) and the per-step mean loop (:605-612)
 * have no effect on the output for these modes and are not restated.
 * out: uint32[image_h][image_w]; only hit pixels are written (:302-303) — the caller
 * pre-clears (volumeRender.cpp:208, 1022).  Returns the number of loop iterations that
 * performed a transfer-function lookup (the S of "Gsamples/s"). */
int64_t vrdd_oracle_render(const float* vol_original4, const float* vol_fractal4, int W, int H, int D,
                           const float* tf4, int tf_n, const float* m12, uint32_t* out,
                           const vrdd_oracle_render_params* P) {
    const f3 boxMin = {-1.0f, -1.0f, -1.0f}, boxMax = {1.0f, 1.0f, 1.0f};
    const int qm = P->query_method;
    /* query_method + 100: the volume argument of that query method is ONE plane float[V] holding its component */
    const bool planar = qm >= 100;
    const int comp = planar ? 0 : (qm - 1) % 3;
    Volume4 vol = {((qm % 100) >= 4) ? vol_fractal4 : vol_original4, W, H, D, planar ? 1 : 4};
    const int imageW = P->image_w, imageH = P->image_h;
    int64_t samples = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : samples)
    for (int y = P->y0; y < P->y1; ++y) {
        for (int x = 0; x < imageW; ++x) {
            const RaySetup ray = make_ray(m12, x, y, imageW, imageH);                /* :288-296 */
            const f3 o = ray.o, d = ray.d;
            float tnear, tfar;
            if (!intersect_box(o, d, boxMin, boxMax, &tnear, &tfar)) continue;    /* :300-303 */
            if (tnear < 0.0f) tnear = 0.0f;                                        /* :305-306 */
            f4 sum = {0, 0, 0, 0};
            float t = tnear;
            f3 pos = first_pos(o, d, tnear);                                       /* :311 */
            f3 step = {d.x * P->tstep, d.y * P->tstep, d.z * P->tstep};            /* :312 */
            for (int i = 0; i < P->max_steps; ++i) {                              /* :381 */
                float sample = tex3d_linear_comp(vol, comp, pos.x * 0.5f + 0.5f, pos.y * 0.5f + 0.5f,
                                                 pos.z * 0.5f + 0.5f, P->weight_quant);
                samples += 1;
                f4 col = tex1d_linear4(tf4, tf_n, (sample - P->transfer_offset) * P->transfer_scale,
                                       P->weight_quant);                           /* :683-684 */
                col.w = col.w * P->density;                                        /* :685 */
                col.x = col.x * col.w; col.y = col.y * col.w; col.z = col.z * col.w; /* :691-693 */
                composite(sum, col);                                               /* :695 */
                if (sum.w > P->opacity_threshold) break;                           /* :698 */
                t = t + P->tstep;                                                  /* :701 */
                if (t > tfar) break;                                               /* :703 */
                pos.x = pos.x + step.x; pos.y = pos.y + step.y; pos.z = pos.z + step.z; /* :706 */
            }
            sum.x = sum.x * P->brightness; sum.y = sum.y * P->brightness;          /* :713 */
            sum.z = sum.z * P->brightness; sum.w = sum.w * P->brightness;
            out[(size_t)y * imageW + x] = pack_rgba(sum);                          /* :716 */
        }
    }
    return samples;
}

/* Nearest-texel rule of the texture unit for a normalised coordinate (measured on B200,
 * tools/probe_texture3.py, 0 mismatches): the coordinate is clamped to [0,1] and truncated to 21
 * fractional bits like in the linear filter, idx = (U * N) >> 21, clamped to N-1.  A coordinate that
 * sits exactly on a texel boundary, u = fl(k/N), therefore usually lands in texel k-1. */
static inline int point_index_hw(float u, int N) {
    float uc = fminf(fmaxf(u, 0.0f), 1.0f);
    long long U = (long long)std::floor(uc * 2097152.0f);
    long long i = (U * (long long)N) >> 21;
    return (int)(i > N - 1 ? N - 1 : i);
}

/* d_render with queryMethod 7, "interpolated mean" (volumeRender_kernel.cu:253-270, 320-367,
 * 395-480): the eight corners of the cell around the sample are floor/ceil(pos01*dim)/dim; each
 * corner point-samples the block-index texture `tex` (:361, 456; point, normalised, clamp,
 * :2161-2165) and takes that block's un-normalised bin-centre mean (:362-366); the sample is the
 * trilinear blend of the eight means (double arithmetic, :472-478) times 50 (:479).  The corner
 * cache is refreshed when the sample leaves [corner0, corner7] (:396, inInterpolation).  When
 * pos01*dim is an integer, ceil == floor and the blend divides by zero (:466-471): the sample is
 * NaN, the transfer-function fetch of a NaN coordinate returns texel 0 (measured), and the
 * reference's "vertical and horizontal line" artefact (ver1.9.6.txt:166) appears — kept.
 * hist: float[V][B].  Other arguments as vrdd_oracle_render. */
int64_t vrdd_oracle_render_mode7(const float* hist, int W, int H, int D, int B, const float* tf4, int tf_n,
                                 const float* m12, uint32_t* out, const vrdd_oracle_render_params* P) {
    const f3 boxMin = {-1.0f, -1.0f, -1.0f}, boxMax = {1.0f, 1.0f, 1.0f};
    const int imageW = P->image_w, imageH = P->image_h;
    const int64_t V = (int64_t)W * H * D;
    /* per-block mean, the expression of :362-366 / :742-747 */
    std::vector<float> mean_raw(V);
#pragma omp parallel for schedule(static)
    for (int64_t v = 0; v < V; ++v) {
        float MaxHistogram = 0.0217, MinHistogram = 0.0;
        float binWidth = (MaxHistogram - MinHistogram) / (float)B;
        float m = 0;
        for (int i = 0; i < B; ++i)
            m = (float)((double)m + (double)hist[v * B + i] * ((double)(binWidth * (float)i) + (double)binWidth / 2.0));
        mean_raw[v] = m;
    }
    int64_t samples = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : samples)
    for (int y = P->y0; y < P->y1; ++y) {
        for (int x = 0; x < imageW; ++x) {
            const RaySetup ray = make_ray(m12, x, y, imageW, imageH);                /* :288-296 */
            const f3 o = ray.o, d = ray.d;
            float tnear, tfar;
            if (!intersect_box(o, d, boxMin, boxMax, &tnear, &tfar)) continue;
            if (tnear < 0.0f) tnear = 0.0f;
            f4 sum = {0, 0, 0, 0};
            float t = tnear;
            f3 pos = first_pos(o, d, tnear);                                       /* :311 */
            f3 step = {d.x * P->tstep, d.y * P->tstep, d.z * P->tstep};
            f3 bot, top;                 /* interPos[0] and interPos[7]; the other six corners mix their components */
            float mean[8];
            auto refresh = [&](const f3& p) {                                        /* :322-367 and :398-463 */
                float px = p.x * 0.5f + 0.5f, py = p.y * 0.5f + 0.5f, pz = p.z * 0.5f + 0.5f;
                bot.x = std::floor(px * (float)W) / (float)W; top.x = std::ceil(px * (float)W) / (float)W;
                bot.y = std::floor(py * (float)H) / (float)H; top.y = std::ceil(py * (float)H) / (float)H;
                bot.z = std::floor(pz * (float)D) / (float)D; top.z = std::ceil(pz * (float)D) / (float)D;
                for (int j = 0; j < 8; ++j) {
                    float cx = (j & 1) ? top.x : bot.x, cy = (j & 2) ? top.y : bot.y, cz = (j & 4) ? top.z : bot.z;
                    int64_t index = point_index_hw(cx, W) + (int64_t)W * (point_index_hw(cy, H) + (int64_t)H * point_index_hw(cz, D));
                    mean[j] = mean_raw[index];
                }
            };
            refresh(pos);
            for (int i = 0; i < P->max_steps; ++i) {
                float px = pos.x * 0.5f + 0.5f, py = pos.y * 0.5f + 0.5f, pz = pos.z * 0.5f + 0.5f;
                if (px < bot.x || py < bot.y || pz < bot.z || px > top.x || py > top.y || pz > top.z) refresh(pos);   /* :253-270, 396 */
                float xd = (px - bot.x) / (top.x - bot.x);                          /* :466-471 */
                float yd = (py - bot.y) / (top.y - bot.y);
                float zd = (pz - bot.z) / (top.z - bot.z);
                /* :472-478.  `a * (1.0 - w) + b * w` with float a, b, w: the first product is promoted by the double
                 * literal, the second is a FLOAT product, the sum is double (fused by nvcc: contraction setting 1) */
                auto lerp7 = [](float a, float b, float w) {
                    const double wa = 1.0 - (double)w, pb = (double)(b * w);
                    return (float)(g_fma_contract ? std::fma(wa, (double)a, pb) : (double)a * wa + pb);
                };
                float mean00 = lerp7(mean[0], mean[1], xd), mean10 = lerp7(mean[2], mean[3], xd);
                float mean01 = lerp7(mean[4], mean[5], xd), mean11 = lerp7(mean[6], mean[7], xd);
                float mean0 = lerp7(mean00, mean10, yd), mean1 = lerp7(mean01, mean11, yd);
                float interMean = lerp7(mean0, mean1, zd);
                float sample = interMean * 50;                                      /* :479 */
                samples += 1;
                f4 col = tex1d_linear4(tf4, tf_n, (sample - P->transfer_offset) * P->transfer_scale, WQ_HW);
                col.w = col.w * P->density;
                col.x = col.x * col.w; col.y = col.y * col.w; col.z = col.z * col.w;
                composite(sum, col);                                               /* :695 */
                if (sum.w > P->opacity_threshold) break;
                t = t + P->tstep;
                if (t > tfar) break;
                pos.x = pos.x + step.x; pos.y = pos.y + step.y; pos.z = pos.z + step.z;
            }
            sum.x = sum.x * P->brightness; sum.y = sum.y * P->brightness;
            sum.z = sum.z * P->brightness; sum.w = sum.w * P->brightness;
            out[(size_t)y * imageW + x] = pack_rgba(sum);
        }
    }
    return samples;
}

int vrdd_oracle_point_index(float u, int N) { return point_index_hw(u, N); }

/* ---- flexible-block-size query chain (volumeRender_kernel.cu:892-1796), SURVEY.md §8f row 1 --------------
 * dataProcessing() = d_divideBlock -> d_queryBlockNew -> d_querySpanNew -> d_computeBlock -> bindToTex.
 * The raw volume (64^3 in the reference) is described by 64-bin histograms of power-of-two-aligned boxes
 * ("spans", 1-based inclusive); boxes of >= 8 voxels are fractal-coded, smaller ones are sparse "simple"
 * histograms with 0-based coordinates (:1464-1471).  A block's histogram is the 8-corner inclusion-exclusion
 * of prefix-box histograms (integral histogram), each prefix box being the weighted sum of the spans its
 * corner decomposes into (:1248-1315).  Quirks kept: the lower corners use `low`, not `low - 1` (:1157-1227),
 * so a block covers (low, high]; statistics use MaxHistogram = 255 and are not normalised (:1084-1098). */
struct vrdd_oracle_flex_tables {
    int raw_w, raw_h, raw_d, bins;
    int n_fractal; const int32_t* span_low; const int32_t* span_high; const int32_t* codebook; const float* errors;
    int n_simple;  const int32_t* simple_low; const int32_t* simple_high; const int32_t* simple_count; const float* simple_hist;
    int n_templates; const float* templates;
};

static int prefix_pieces(int x, int (*out)[2]) {                  /* :1248-1259 */
    int n = 0;
    for (int i = 0; i < 31 && x != 0; ++i)
        if (x & (1 << i)) { out[n][1] = x; x &= ~(1 << i); out[n][0] = x + 1; ++n; }
    return n;
}

/* out4: float4[nx*ny*nz] = (mean, variance, entropy, 0), blocks x fastest; dims3 receives (nx, ny, nz).
 * Returns the number of spans that were not found in the tables (the reference prints and reads garbage;
 * here they contribute nothing). */
int64_t vrdd_oracle_flex_process(const vrdd_oracle_flex_tables* T, int block, float* out4, int* dims3) {
    const int B = T->bins;
    typedef std::array<int, 6> Key;
    std::map<Key, int> fr, si;
    for (int i = 0; i < T->n_fractal; ++i)
        fr.emplace(Key{T->span_low[4 * i], T->span_low[4 * i + 1], T->span_low[4 * i + 2], T->span_high[4 * i],
                       T->span_high[4 * i + 1], T->span_high[4 * i + 2]}, i);      /* first match wins, like the linear scan */
    for (int i = 0; i < T->n_simple; ++i)
        si.emplace(Key{T->simple_low[4 * i], T->simple_low[4 * i + 1], T->simple_low[4 * i + 2], T->simple_high[4 * i],
                       T->simple_high[4 * i + 1], T->simple_high[4 * i + 2]}, i);
    const int vd[3] = {T->raw_w, T->raw_h, T->raw_d};
    int nb[3];
    for (int a = 0; a < 3; ++a) nb[a] = (vd[a] + block - 1) / block;              /* d_divideBlock, :892-1031 */
    dims3[0] = nb[0]; dims3[1] = nb[1]; dims3[2] = nb[2];
    const int64_t nblocks = (int64_t)nb[0] * nb[1] * nb[2];
    int64_t missing = 0;
    std::vector<float> corner_sum((size_t)nblocks * 8 * B);
    for (int64_t b = 0; b < nblocks; ++b) {
        const int bx = (int)(b % nb[0]), by = (int)((b / nb[0]) % nb[1]), bz = (int)(b / ((int64_t)nb[0] * nb[1]));
        const int bi[3] = {bx, by, bz};
        int lo[3], hi[3];
        for (int a = 0; a < 3; ++a) { lo[a] = 1 + bi[a] * block; hi[a] = (bi[a] == nb[a] - 1) ? vd[a] : (bi[a] + 1) * block; }
        for (int j = 0; j < 8; ++j) {                                              /* d_queryBlockNew, :1142-1315 */
            const int c[3] = {(j & 1) ? hi[0] : lo[0], (j & 2) ? hi[1] : lo[1], (j & 4) ? hi[2] : lo[2]};
            int px[32][2], py[32][2], pz[32][2];
            const int nx = prefix_pieces(c[0], px), ny = prefix_pieces(c[1], py), nz = prefix_pieces(c[2], pz);
            float* sum = &corner_sum[((size_t)b * 8 + j) * B];
            for (int k = 0; k < B; ++k) sum[k] = 0.0f;
            for (int ix = 0; ix < nx; ++ix)
                for (int iy = 0; iy < ny; ++iy)
                    for (int iz = 0; iz < nz; ++iz) {                             /* d_querySpanNew, :1318-1544 */
                        const int weight = (px[ix][1] - px[ix][0] + 1) * (py[iy][1] - py[iy][0] + 1) * (pz[iz][1] - pz[iz][0] + 1);
                        float cur[VRDD_SYNTH_MAX_BINS];
                        if (weight >= 8) {
                            auto it = fr.find(Key{px[ix][0], py[iy][0], pz[iz][0], px[ix][1], py[iy][1], pz[iz][1]});
                            if (it == fr.end()) { ++missing; continue; }
                            const int idx = it->second;
                            const int id = T->codebook[4 * idx], shift = T->codebook[4 * idx + 1], flip = T->codebook[4 * idx + 2],
                                      ne = T->codebook[4 * idx + 3];
                            if (id < 0 || id >= T->n_templates || shift < 0 || shift > B || ne < 0 || ne > B) { ++missing; continue; }
                            fractal_transform(T->templates + (size_t)id * B, B, flip, shift, cur);      /* :225-251 */
                            for (int e = 0; e < ne; ++e) {
                                const int bin = (int)T->errors[2 * ((size_t)idx * B + e)];
                                const float val = T->errors[2 * ((size_t)idx * B + e) + 1];
                                if (bin < 0 || bin >= B) continue;
                                cur[bin] = cur[bin] + val;
                                if (cur[bin] < 0) cur[bin] = 0;
                            }
                            float tot = 0;
                            for (int k = 0; k < B; ++k) tot = tot + cur[k];
                            for (int k = 0; k < B; ++k) cur[k] = cur[k] / tot;                             /* no guard, :1437-1439 */
                        } else {
                            auto it = si.find(Key{px[ix][0] - 1, py[iy][0] - 1, pz[iz][0] - 1, px[ix][1] - 1, py[iy][1] - 1, pz[iz][1] - 1});
                            if (it == si.end()) { ++missing; continue; }
                            const int idx = it->second;
                            for (int k = 0; k < B; ++k) cur[k] = 0;
                            for (int e = 0; e < T->simple_count[idx] && e < B; ++e) {
                                const int bin = (int)T->simple_hist[2 * ((size_t)idx * B + e)];
                                if (bin >= 0 && bin < B) cur[bin] = T->simple_hist[2 * ((size_t)idx * B + e) + 1];
                            }
                        }
                        for (int k = 0; k < B; ++k) sum[k] = sum[k] + cur[k] * (float)weight;            /* :1449, 1517 */
                    }
        }
    }
    for (int64_t b = 0; b < nblocks; ++b) {                                      /* d_computeBlock, :1033-1126 */
        const float* c = &corner_sum[(size_t)b * 8 * B];
        float h[VRDD_SYNTH_MAX_BINS];
        for (int s = 0; s < B; ++s) {
            h[s] = c[0 * B + s] + c[3 * B + s] + c[4 * B + s] + c[7 * B + s] - c[1 * B + s] - c[2 * B + s] - c[5 * B + s] - c[6 * B + s];
            if (h[s] < 0) h[s] = 0;
        }
        float total = 0;
        for (int s = 0; s < B; ++s) total = total + h[s];
        if (total > 0)
            for (int s = 0; s < B; ++s) { h[s] = h[s] / total; if (h[s] < 0) h[s] = 0; if (h[s] > 1) h[s] = 1; }
        float MaxHistogram = 255.0, MinHistogram = 0.0, mean = 0;
        float binWidth = (MaxHistogram - MinHistogram) / (float)B;
        for (int i = 0; i < B; ++i)
            mean = (float)((double)mean + (double)h[i] * ((double)(binWidth * (float)i) + (double)binWidth / 2.0));
        float variance = 0;
        for (int i = 0; i < B; ++i) {
            double d = ((double)(binWidth * (float)i) + (double)binWidth / 2.0) - (double)mean;
            variance = (float)((double)variance + ((double)h[i] * d) * d);
        }
        float entropy = 0;
        for (int i = 0; i < B; ++i) {
            double term = (h[i] <= 0) ? 0.0 : ((double)logf(h[i]) / std::log(2.0));
            entropy = (float)((double)entropy + (double)h[i] * term);
        }
        entropy = -entropy;
        entropy = entropy / (logf((float)B) / logf(2.0f));
        out4[4 * b] = mean; out4[4 * b + 1] = variance; out4[4 * b + 2] = entropy; out4[4 * b + 3] = 0.0f;
    }
    return missing;
}

/* tex3D(flexBlockTex, x, y, z): float4 volume of nx*ny*nz blocks inside a zero-filled 500^3 array
 * (bindToTex, :1572-1696), linear, UN-normalised, clamp.  Un-normalised coordinates are not truncated to
 * 21 bits: q = floor(x*256 + 0.5) - 128 (measured, tools/probe_texture4.py); weights as in weights_hw. */
static float tex3d_flex(const float* blocks4, int nx, int ny, int nz, int comp, float x, float y, float z) {
    const int N = 500;
    int ii[3], aa[3];
    const float xyz[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        float v = xyz[a];
        if (!(v == v)) v = 0.0f;
        double q = std::floor((double)v * 256.0 + 0.5) - 128.0;
        if (q < 0) q = 0;
        if (q > (double)(N - 1) * 256) q = (double)(N - 1) * 256;
        long long qi = (long long)q;
        ii[a] = (int)(qi >> 8); aa[a] = (int)(qi & 255);
    }
    int wt[8];
    weights_hw(aa[0], aa[1], aa[2], wt);
    float acc = 0.0f;
    for (int n = 0; n < 8; ++n) {
        if (wt[n] == 0) continue;
        int tx = ii[0] + (n & 1), ty = ii[1] + ((n >> 1) & 1), tz = ii[2] + (n >> 2);
        float t = (tx < nx && ty < ny && tz < nz) ? blocks4[4 * ((size_t)tx + (size_t)nx * ((size_t)ty + (size_t)ny * (size_t)tz)) + comp] : 0.0f;
        acc = acc + ((float)wt[n] * (1.0f / 256.0f)) * t;
    }
    return acc;
}

/* d_render for queryMethod 8 / 9 / 0: entropy / mean / variance of the flexible blocks (:654-680). */
int64_t vrdd_oracle_render_flex(const float* blocks4, int nx, int ny, int nz, const float* tf4, int tf_n,
                                const float* m12, uint32_t* out, const vrdd_oracle_render_params* P) {
    const f3 boxMin = {-1.0f, -1.0f, -1.0f}, boxMax = {1.0f, 1.0f, 1.0f};
    const int qm = P->query_method;
    const int comp = (qm == 8) ? 2 : (qm == 9) ? 0 : 1;
    const int imageW = P->image_w, imageH = P->image_h;
    int64_t samples = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : samples)
    for (int y = P->y0; y < P->y1; ++y) {
        for (int x = 0; x < imageW; ++x) {
            const RaySetup ray = make_ray(m12, x, y, imageW, imageH);                /* :288-296 */
            const f3 o = ray.o, d = ray.d;
            float tnear, tfar;
            if (!intersect_box(o, d, boxMin, boxMax, &tnear, &tfar)) continue;
            if (tnear < 0.0f) tnear = 0.0f;
            f4 sum = {0, 0, 0, 0};
            float t = tnear;
            f3 pos = first_pos(o, d, tnear);                                       /* :311 */
            f3 step = {d.x * P->tstep, d.y * P->tstep, d.z * P->tstep};
            for (int i = 0; i < P->max_steps; ++i) {
                float sample = tex3d_flex(blocks4, nx, ny, nz, comp, (pos.x * 0.5f + 0.5f) * (float)nx,
                                          (pos.y * 0.5f + 0.5f) * (float)ny, (pos.z * 0.5f + 0.5f) * (float)nz);
                samples += 1;
                f4 col = tex1d_linear4(tf4, tf_n, (sample - P->transfer_offset) * P->transfer_scale, WQ_HW);
                col.w = col.w * P->density;
                col.x = col.x * col.w; col.y = col.y * col.w; col.z = col.z * col.w;
                composite(sum, col);                                               /* :695 */
                if (sum.w > P->opacity_threshold) break;
                t = t + P->tstep;
                if (t > tfar) break;
                pos.x = pos.x + step.x; pos.y = pos.y + step.y; pos.z = pos.z + step.z;
            }
            sum.x = sum.x * P->brightness; sum.y = sum.y * P->brightness;
            sum.z = sum.z * P->brightness; sum.w = sum.w * P->brightness;
            out[(size_t)y * imageW + x] = pack_rgba(sum);
        }
    }
    return samples;
}

/* Direct access to the filter model, for the texture-unit conformance test. */
float vrdd_oracle_tex3d(const float* vol4, int W, int H, int D, int comp, float u, float v, float w, int wq) {
    Volume4 vol = {vol4, W, H, D};
    return tex3d_linear_comp(vol, comp, u, v, w, wq);
}
void vrdd_oracle_tex1d4(const float* tf4, int n, float u, int wq, float* out4) {
    f4 r = tex1d_linear4(tf4, n, u, wq);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

/* ---- synthetic inputs (include/vrdd_synth.h), host side --------------------------- */

/* hist[(z-z0)*H*W + y*W + x][B] for z in [z0, z0+nz) of a W x H x D volume */
void vrdd_oracle_synth_histograms(uint32_t seed, int W, int H, int D, int B, int z0, int nz, float* hist) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int z = z0; z < z0 + nz; ++z)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t v = (size_t)x + (size_t)W * ((size_t)y + (size_t)H * (size_t)(z - z0));
                vrdd_synth_histogram(seed, x, y, z, W, H, D, B, hist + v * B);
            }
}

void vrdd_oracle_synth_templates(uint32_t seed, int T, int B, float* templates) {
    for (int k = 0; k < T; ++k) vrdd_synth_template(seed, k, T, B, templates + (size_t)k * B);
}

/* codebook int4[V]; errors float2[V][B] dense like the reference's errorsbook, entries
 * beyond NE zero-filled (the reference leaves them uninitialised, volumeRender.cpp:582) */
void vrdd_oracle_synth_fractal(uint32_t seed, int W, int H, int D, int B, int T, int max_ne, int z0, int nz,
                               int32_t* codebook, float* errors) {
#pragma omp parallel for schedule(static) collapse(2)
    for (int z = z0; z < z0 + nz; ++z)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                size_t v = (size_t)x + (size_t)W * ((size_t)y + (size_t)H * (size_t)(z - z0));
                int code[4], eb[VRDD_SYNTH_MAX_BINS];
                float ev[VRDD_SYNTH_MAX_BINS];
                vrdd_synth_fractal_code(seed, x, y, z, W, H, D, B, T, max_ne, code, eb, ev);
                for (int k = 0; k < 4; ++k) codebook[4 * v + k] = code[k];
                for (int k = 0; k < B; ++k) {
                    errors[2 * (v * B + k) + 0] = (k < code[3]) ? (float)eb[k] : 0.0f;
                    errors[2 * (v * B + k) + 1] = (k < code[3]) ? ev[k] : 0.0f;
                }
            }
}

}  // extern "C"
