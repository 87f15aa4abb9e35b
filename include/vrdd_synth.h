/*
 * vrdd_synth.h — seeded synthetic distribution volumes.
 *
 * The reference ships none of the nine .bin inputs it loads
 * (/root/reference/volumeRender.cpp:76-84), so every test and benchmark input is
 * generated.  This header is the single definition of those inputs.  It is compiled
 * both by g++ (oracle, tests) and by nvcc (device-side generation of volumes too large
 * to upload), and must give IDENTICAL BITS on both.  It therefore uses only
 *   - 32-bit integer hashing, and
 *   - IEEE-754 round-to-nearest +, -, *, / on float, one rounding per operation
 *     (device: __fadd_rn/__fmul_rn/__fdiv_rn, which nvcc never contracts into FMA;
 *      host: plain operators, and the translation unit MUST be built with
 *      -ffp-contract=off).
 * No transcendental functions.
 *
 * The generated data honours the invariants the reference checks at run time
 * (volumeRender_kernel.cu:781-838, volumeRender.cpp:611-614, 678-684):
 *   0 <= templateId < T, 0 <= NE <= B, 0 <= shift < B, every frequency in [0,1],
 *   error bin in [0,B), error value in [-1,1], histograms sum to 1 (fp32).
 *
 * This is test/benchmark data, not part of the decode or ray-cast algorithm.
 */
#ifndef VRDD_SYNTH_H_
#define VRDD_SYNTH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define VRDD_HD __host__ __device__ __forceinline__
#else
#define VRDD_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define VRDD_ADD(a, b) __fadd_rn((a), (b))
#define VRDD_SUB(a, b) __fsub_rn((a), (b))
#define VRDD_MUL(a, b) __fmul_rn((a), (b))
#define VRDD_DIV(a, b) __fdiv_rn((a), (b))
#else
#define VRDD_ADD(a, b) ((float)(a) + (float)(b))
#define VRDD_SUB(a, b) ((float)(a) - (float)(b))
#define VRDD_MUL(a, b) ((float)(a) * (float)(b))
#define VRDD_DIV(a, b) ((float)(a) / (float)(b))
#endif

#define VRDD_SYNTH_MAX_BINS 64

/* ---- integer hashing ------------------------------------------------------------ */

VRDD_HD uint32_t vrdd_mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

VRDD_HD uint32_t vrdd_hash4(uint32_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = vrdd_mix32(seed ^ 0x9e3779b9u);
    h = vrdd_mix32(h ^ (a * 0x9e3779b1u + 0x7f4a7c15u));
    h = vrdd_mix32(h ^ (b * 0x85ebca77u + 0x165667b1u));
    h = vrdd_mix32(h ^ (c * 0xc2b2ae3du + 0x27d4eb2fu));
    h = vrdd_mix32(h ^ (d * 0x27d4eb2fu + 0x9e3779b9u));
    return h;
}

/* uniform in [0,1): 24 random mantissa bits, conversion and scaling are exact */
VRDD_HD float vrdd_u01(uint32_t h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

/* ---- smooth scalar fields (value noise, exact ops only) ------------------------- */

VRDD_HD float vrdd_lerp_exact(float a, float b, float s) {
    return VRDD_ADD(a, VRDD_MUL(s, VRDD_SUB(b, a)));
}

/* value noise on an integer lattice; (px,py,pz) >= 0 in lattice units */
VRDD_HD float vrdd_value_noise3(uint32_t seed, float px, float py, float pz) {
    int ix = (int)px, iy = (int)py, iz = (int)pz;          /* truncation == floor for >= 0 */
    float tx = VRDD_SUB(px, (float)ix), ty = VRDD_SUB(py, (float)iy), tz = VRDD_SUB(pz, (float)iz);
    /* smoothstep s = t*t*(3 - 2t) */
    float sx = VRDD_MUL(VRDD_MUL(tx, tx), VRDD_SUB(3.0f, VRDD_MUL(2.0f, tx)));
    float sy = VRDD_MUL(VRDD_MUL(ty, ty), VRDD_SUB(3.0f, VRDD_MUL(2.0f, ty)));
    float sz = VRDD_MUL(VRDD_MUL(tz, tz), VRDD_SUB(3.0f, VRDD_MUL(2.0f, tz)));
    float c[8];
    for (int k = 0; k < 8; ++k)
        c[k] = vrdd_u01(vrdd_hash4(seed, (uint32_t)(ix + (k & 1)), (uint32_t)(iy + ((k >> 1) & 1)),
                                   (uint32_t)(iz + (k >> 2)), 0x51u));
    float x00 = vrdd_lerp_exact(c[0], c[1], sx), x10 = vrdd_lerp_exact(c[2], c[3], sx);
    float x01 = vrdd_lerp_exact(c[4], c[5], sx), x11 = vrdd_lerp_exact(c[6], c[7], sx);
    float y0 = vrdd_lerp_exact(x00, x10, sy), y1 = vrdd_lerp_exact(x01, x11, sy);
    return vrdd_lerp_exact(y0, y1, sz);
}

/* The two fields that shape a voxel's histogram:
 *   f in [0,1]   -> centre bin  (drives the decoded mean)
 *   g in [0,1]   -> bump width  (drives variance and entropy)
 * (x,y,z) is the GLOBAL voxel coordinate, (W,H,D) the GLOBAL volume size, so a slab or
 * a brick of a larger volume generates exactly the voxels of the whole. */
VRDD_HD void vrdd_synth_fields(uint32_t seed, int x, int y, int z, int W, int H, int D,
                               float* f_out, float* g_out) {
    float ux = VRDD_DIV(VRDD_ADD((float)x, 0.5f), (float)W);
    float uy = VRDD_DIV(VRDD_ADD((float)y, 0.5f), (float)H);
    float uz = VRDD_DIV(VRDD_ADD((float)z, 0.5f), (float)D);
    float n1 = vrdd_value_noise3(seed + 11u, VRDD_MUL(ux, 3.0f), VRDD_MUL(uy, 3.0f), VRDD_MUL(uz, 3.0f));
    float n2 = vrdd_value_noise3(seed + 23u, VRDD_MUL(ux, 7.0f), VRDD_MUL(uy, 7.0f), VRDD_MUL(uz, 7.0f));
    float n3 = vrdd_value_noise3(seed + 37u, VRDD_MUL(ux, 17.0f), VRDD_MUL(uy, 17.0f), VRDD_MUL(uz, 17.0f));
    float n = VRDD_ADD(VRDD_ADD(VRDD_MUL(0.55f, n1), VRDD_MUL(0.30f, n2)), VRDD_MUL(0.15f, n3));
    /* radial falloff: dense core, empty corners (keeps part of every ray transparent) */
    float cx = VRDD_SUB(VRDD_MUL(2.0f, ux), 1.0f), cy = VRDD_SUB(VRDD_MUL(2.0f, uy), 1.0f),
          cz = VRDD_SUB(VRDD_MUL(2.0f, uz), 1.0f);
    float r2 = VRDD_ADD(VRDD_ADD(VRDD_MUL(cx, cx), VRDD_MUL(cy, cy)), VRDD_MUL(cz, cz));
    float fall = VRDD_SUB(1.0f, VRDD_MUL(0.45f, r2));
    if (fall < 0.0f) fall = 0.0f;
    float f = VRDD_MUL(n, fall);
    /* per-voxel jitter so neighbouring voxels never carry identical histograms */
    float j = vrdd_u01(vrdd_hash4(seed + 41u, (uint32_t)x, (uint32_t)y, (uint32_t)z, 0x77u));
    f = VRDD_ADD(VRDD_MUL(f, 0.97f), VRDD_MUL(j, 0.03f));
    if (f > 1.0f) f = 1.0f;
    float g = vrdd_value_noise3(seed + 53u, VRDD_MUL(ux, 5.0f), VRDD_MUL(uy, 5.0f), VRDD_MUL(uz, 5.0f));
    *f_out = f;
    *g_out = g;
}

/* Compact-support bump  q_i = max(0, 1 - ((i-c)/w)^2)^2, normalised to sum 1 (fp32,
 * sequential sum).  Bins outside the support are exactly 0, which exercises the
 * `p <= 0` branch of the entropy (volumeRender_kernel.cu:765-766). */
VRDD_HD void vrdd_synth_bump(float c, float w, int B, float* p /* [B] */) {
    float tot = 0.0f;
    for (int i = 0; i < B; ++i) {
        float d = VRDD_DIV(VRDD_SUB((float)i, c), w);
        float q = VRDD_SUB(1.0f, VRDD_MUL(d, d));
        if (q < 0.0f) q = 0.0f;
        q = VRDD_MUL(q, q);
        p[i] = q;
        tot = VRDD_ADD(tot, q);
    }
    for (int i = 0; i < B; ++i) p[i] = VRDD_DIV(p[i], tot);
}

/* Raw histogram of voxel (x,y,z): the reference's `float hist[V][B]` row
 * (volumeRender.cpp:538-556, layout x-fastest, volumeRender_kernel.cu:740). */
VRDD_HD void vrdd_synth_histogram(uint32_t seed, int x, int y, int z, int W, int H, int D, int B,
                                  float* p /* [B] */) {
    float f, g;
    vrdd_synth_fields(seed, x, y, z, W, H, D, &f, &g);
    float c = VRDD_MUL(f, (float)(B - 1));
    float w = VRDD_ADD(1.5f, VRDD_MUL(g, 9.0f));
    vrdd_synth_bump(c, w, B, p);
}

/* Template k of T (volumeRender.cpp:644-691): bumps whose centre sweeps the bin range */
VRDD_HD void vrdd_synth_template(uint32_t seed, int k, int T, int B, float* p /* [B] */) {
    uint32_t h = vrdd_hash4(seed + 61u, (uint32_t)k, 0u, 0u, 0x13u);
    float c = VRDD_MUL(VRDD_DIV((float)k, (float)(T > 1 ? T - 1 : 1)), (float)(B - 1));
    float w = VRDD_ADD(1.5f, VRDD_MUL(vrdd_u01(h), 9.0f));
    vrdd_synth_bump(c, w, B, p);
}

/* Fractal code of voxel (x,y,z) (volumeRender.cpp:558-642): (templateId, shift, flip, NE)
 * and NE sparse (bin,value) corrections with distinct bins.  The template id follows the
 * smooth field so the decoded volume is spatially coherent; shift is small and signed
 * through the modular wrap (0..2 or B-2..B-1), flip is a coin toss. */
VRDD_HD void vrdd_synth_fractal_code(uint32_t seed, int x, int y, int z, int W, int H, int D,
                                     int B, int T, int max_ne,
                                     int* code /* [4] */, int* err_bin /* [<=B] */,
                                     float* err_val /* [<=B] */) {
    float f, g;
    vrdd_synth_fields(seed, x, y, z, W, H, D, &f, &g);
    uint32_t h = vrdd_hash4(seed + 71u, (uint32_t)x, (uint32_t)y, (uint32_t)z, 0x29u);
    int id = (int)VRDD_MUL(f, (float)(T - 1));
    id += (int)(h & 3u) - 1;                         /* -1..2 jitter */
    if (id < 0) id = 0;
    if (id > T - 1) id = T - 1;
    int sh = (int)((h >> 2) % 5u) - 2;               /* -2..2 */
    if (sh < 0) sh += B;                             /* wrap: 0 <= shift < B */
    int flip = (int)((h >> 8) & 1u);
    if (max_ne > B) max_ne = B;
    int ne = (int)((h >> 9) % (uint32_t)(max_ne + 1));
    code[0] = id; code[1] = sh; code[2] = flip; code[3] = ne;
    uint32_t start = (h >> 16) % (uint32_t)B;
    uint32_t stride = ((h >> 21) | 1u) % (uint32_t)B; /* odd -> coprime with power-of-two B */
    if ((stride & 1u) == 0u) stride += 1u;
    for (int k = 0; k < ne; ++k) {
        uint32_t hk = vrdd_hash4(seed + 83u, (uint32_t)x, (uint32_t)y, (uint32_t)z, (uint32_t)k);
        err_bin[k] = (int)((start + (uint32_t)k * stride) % (uint32_t)B);
        /* uniform in [-0.05, 0.05) */
        err_val[k] = VRDD_MUL(VRDD_SUB(vrdd_u01(hk), 0.5f), 0.1f);
    }
}

#endif /* VRDD_SYNTH_H_ */
