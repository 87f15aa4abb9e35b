// sortlast.cu — sort-last rendering of a volume that is brick-decomposed over GPUs
// (BASELINE.json configs[4]).  New work: the reference is single-GPU.
//
// Each handle stores ONE brick of the global volume (its voxels plus a one-voxel ghost
// layer) as linear fp32 planes.  The reference stops a ray when the accumulated alpha
// exceeds 0.95 (volumeRender_kernel.cu:698); plain "over" compositing of independently
// rendered bricks cannot reproduce that (the dropped tail is worth up to 12 LSB), so the
// accumulated alpha is forwarded exactly, but without serialising the bricks:
//
//   pass 1  vrdd_render_brick_alpha   every rank marches its own samples and accumulates
//                                     alpha only                         -> A_seg[H][W]
//           all-gather of A_seg over ranks (host side, NCCL)
//   compose vrdd_compose_alpha_in     per pixel, the alpha entering this brick is the
//                                     front-to-back composite of the A_seg of the bricks the
//                                     ray crossed before it              -> A_in[H][W]
//   pass 2  vrdd_render_brick_color   marches again from A_in, with the reference's early
//                                     exit, and emits the increments (dR, dG, dB, dA)
//           SUM reduction of the increments over ranks (NCCL), then vrdd_pack_frame.
//
// Direct-send form (round 2; vrdd_render_brick_alpha_send / _color_send / vrdd_pack_frame_slots): the two exchanges travel
// in the kernels' own stores over NVLink instead of in NCCL calls.  Pass 1 writes the rows of its screen window straight
// into slot [brick] of EVERY rank's table of segment alphas (CUDA IPC mappings), pass 2 writes its increments into slot
// [brick] of the ROOT's table, the last block of each launch bumps a counter next to the table(s) with release.sys, and
// the reader's stream waits for the count (vrdd_stream_wait_flag).  The root sums the slots in brick order and packs.
// Only the window rows are launched at all.  Tables are double-buffered by frame parity: a rank that sees every other
// rank's pass-1 counter of frame k + 1 knows they have finished reading the tables of frame k (stream order there), so
// frame k + 2 may overwrite them.
//
// Band owners (vrdd_render_brick_color_send_bands / vrdd_pack_band_slots): instead of all increments converging on the
// root (N x 25 MB per 2048^2 frame into one GPU's NVLink port), image rows are dealt out in bands of band_rows rows, band
// o to rank o.  Pass 2 stores every pixel's increment into the table of the band's owner; each owner sums the slots of
// its band in brick order, packs to RGBA8 and stores the packed rows (4 B per pixel) into the root's frame.  Every rank
// then receives 1/N of the increments, and the sum + pack is spread over all GPUs.
//
// Fused first segment (direct-send form): where nothing precedes a brick on a ray the incoming alpha is exactly 0, and
// pass 2 would repeat pass 1's march step for step.  So pass 1 of the direct-send form accumulates colour as well (from
// alpha 0, with the reference's early exit on its own alpha) and keeps (dR, dG, dB, dA) per pixel in a scratch buffer of
// the context; pass 2 forwards that record for every pixel whose incoming alpha came out as 0.0f and marches only the
// others.  With the reference's constants rays die within the first brick they cross, so pass 2 all but disappears.  The
// record is bit for bit what pass 2 would compute (same code, same order).
//
// A sample belongs to the brick whose half-open texture-coordinate box contains it; every
// rank walks the SAME ray recurrence from the global box entry (pos += step is not
// restarted mid-ray, cf. SURVEY.md §7 "incremental stepping"), and filters with the texture
// unit's integer weight scheme evaluated on GLOBAL coordinates, so each sample value is the
// one the single-GPU manual sampler produces.  What differs from the single-GPU image is
// fp32 rounding in A_in and in the final sum: within +-1 LSB (tests/test_gpu_sortlast.py).
#include "common.cuh"

#include <cstring>

namespace vrdd {

namespace {

constexpr int kBlock = 256;

struct BrickArgs {
    cudaTextureObject_t tex;        // this brick's plane as a 3-D array, linear filter, UN-normalised coordinates; or 0
    const float* plane;             // ... or as a plane in global memory
    int bricked, bW, bH;            // layout of `plane`: 4x4x4-bricked (common.cuh) or linear, x fastest
    int sx, sy, sz;                 // local (stored) size in voxels, ghost included
    int gw, gh, gd;                 // global volume size
    int ox, oy, oz;                 // global voxel coordinate of local (0,0,0)
    float lo[3], hi[3];             // owned samples: lo <= texcoord < hi
    const float4* tf_tab;
    int tf_n;
    int iw, ih;
    float m[12];
    float density, t_offset, t_scale, tstep, thresh;
    int max_steps;
    int ref_rounding;               // ray set-up rounded like the reference's nvcc build (common.cuh, eye_ray)
    const float* alpha_in;          // pass 2
    float* alpha_seg;               // pass 1 out
    float4* partial;                // pass 2 out
    unsigned long long* samples;
    // direct-send form: n_dst > 0
    int n_dst;                      // tables to write (pass 1: every rank's; pass 2: the root's only)
    float* dst[VRDD_MAX_PEERS + 1]; // base of slot [brick] in each table: float[rows][iw] (pass 1) or float4[rows][iw] (pass 2)
    unsigned* flags[VRDD_MAX_PEERS + 1];   // counter next to each table, bumped when the launch is complete
    unsigned* tickets;              // this context's block counter
    int row0, rows;                 // my screen window: only these rows are launched and written
    int n_items;                    // 16x16-pixel items of the launch (the window's rows)
    int band_rows;                  // pass 2, band owners: > 0 -> dst[o] is slot [brick] of the owner of rows [o * band_rows, (o + 1) * band_rows)
    float4* first4;                 // FUSE: (dR, dG, dB, dA) of the march from alpha 0, float4[rows][iw] (pass 1 writes, pass 2 reads)
};

__device__ __forceinline__ void split_hw(float u, int n256, int& i, int& a) {
    const unsigned U = (unsigned)(__saturatef(u) * 2097152.0f);
    const unsigned long long p = (unsigned long long)U * (unsigned)n256 + (1u << 20);
    int q = (int)(p >> 21) - 128;
    q = min(max(q, 0), n256 - 256);
    i = q >> 8;
    a = q & 255;
}

// the texture unit's trilinear filter on GLOBAL coordinates, texels read from the local brick
__device__ __forceinline__ float sample_brick(const BrickArgs& A, float u, float v, float w) {
    int i, j, k, a, b, c;
    split_hw(u, A.gw << 8, i, a);
    split_hw(v, A.gh << 8, j, b);
    split_hw(w, A.gd << 8, k, c);
    const int i1 = min(i + 1, A.gw - 1), j1 = min(j + 1, A.gh - 1), k1 = min(k + 1, A.gd - 1);
    // local indices; clamped so that a weight-0 neighbour outside the ghost layer stays in bounds
    const int x0 = min(max(i - A.ox, 0), A.sx - 1), x1 = min(max(i1 - A.ox, 0), A.sx - 1);
    const int y0 = min(max(j - A.oy, 0), A.sy - 1), y1 = min(max(j1 - A.oy, 0), A.sy - 1);
    const int z0 = min(max(k - A.oz, 0), A.sz - 1), z1 = min(max(k1 - A.oz, 0), A.sz - 1);
    const float* P = A.plane;
    float t000, t100, t010, t110, t001, t101, t011, t111;
    if (A.bricked) {
        t000 = __ldg(P + brick_index(x0, y0, z0, A.bW, A.bH)); t100 = __ldg(P + brick_index(x1, y0, z0, A.bW, A.bH));
        t010 = __ldg(P + brick_index(x0, y1, z0, A.bW, A.bH)); t110 = __ldg(P + brick_index(x1, y1, z0, A.bW, A.bH));
        t001 = __ldg(P + brick_index(x0, y0, z1, A.bW, A.bH)); t101 = __ldg(P + brick_index(x1, y0, z1, A.bW, A.bH));
        t011 = __ldg(P + brick_index(x0, y1, z1, A.bW, A.bH)); t111 = __ldg(P + brick_index(x1, y1, z1, A.bW, A.bH));
    } else {
        const size_t r00 = ((size_t)z0 * A.sy + y0) * A.sx, r10 = ((size_t)z0 * A.sy + y1) * A.sx;
        const size_t r01 = ((size_t)z1 * A.sy + y0) * A.sx, r11 = ((size_t)z1 * A.sy + y1) * A.sx;
        t000 = __ldg(P + r00 + x0); t100 = __ldg(P + r00 + x1);
        t010 = __ldg(P + r10 + x0); t110 = __ldg(P + r10 + x1);
        t001 = __ldg(P + r01 + x0); t101 = __ldg(P + r01 + x1);
        t011 = __ldg(P + r11 + x0); t111 = __ldg(P + r11 + x1);
    }
    const int zz1 = c, zz0 = 256 - c;
    const int x10 = (zz0 * a + 128) >> 8, x00 = zz0 - x10;
    const int x11 = (zz1 * a + 128) >> 8, x01 = zz1 - x11;
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

// The same filter from the texture unit.  The fixed-point coordinate q = 256 * texel + fraction is formed
// from the GLOBAL coordinate exactly as the unit would for the whole volume (split_hw), moved into the
// brick, and handed over as the un-normalised coordinate (q + 128) / 256 — exact in fp32 — from which the
// unit recovers q_local = floor(x * 256 + 0.5) - 128 (measured, tools/probe_texture4.py) and applies its
// own weight split: one fetch instead of eight loads and forty integer operations.
__device__ __forceinline__ float sample_brick_tex(const BrickArgs& A, float u, float v, float w) {
    int i, j, k, a, b, c;
    split_hw(u, A.gw << 8, i, a);
    split_hw(v, A.gh << 8, j, b);
    split_hw(w, A.gd << 8, k, c);
    const float xl = (float)(((i - A.ox) << 8) + a + 128) * (1.0f / 256.0f);
    const float yl = (float)(((j - A.oy) << 8) + b + 128) * (1.0f / 256.0f);
    const float zl = (float)(((k - A.oz) << 8) + c + 128) * (1.0f / 256.0f);
    return tex3D<float>(A.tex, xl, yl, zl);
}

__device__ __forceinline__ float4 tf_lookup(const float4* tab, int n, float u) {
    int i, a;
    split_hw(u, n << 8, i, a);
    const float4 c0 = tab[i], c1 = tab[min(i + 1, n - 1)];
    const float w0 = (float)(256 - a) * (1.0f / 256.0f), w1 = (float)a * (1.0f / 256.0f);
    return make_float4(w0 * c0.x + w1 * c1.x, w0 * c0.y + w1 * c1.y, w0 * c0.z + w1 * c1.z, w0 * c0.w + w1 * c1.w);
}

// Eye ray of pixel (x, y): common.cuh, eye_ray — the same function as raycast.cu, so every rank and the
// single-GPU kernel agree bit for bit on direction, tnear and tfar.
struct RaySetup { float ox, oy, oz, dx, dy, dz, tnear, tfar; bool hit; };
__device__ __forceinline__ RaySetup make_ray(const float* m, int x, int y, int iw, int ih, int ref_rounding) {
    const EyeRay E = eye_ray(m, x, y, iw, ih, ref_rounding);
    RaySetup R;
    R.ox = E.ox; R.oy = E.oy; R.oz = E.oz; R.dx = E.dx; R.dy = E.dy; R.dz = E.dz; R.tnear = E.tnear; R.tfar = E.tfar;
    R.hit = R.tfar > R.tnear;
    if (R.tnear < 0.0f) R.tnear = 0.0f;
    return R;
}

// PASS 1: alpha of this brick's segment.  PASS 2: colour increments from the incoming alpha.
template <int PASS, bool COUNT, bool TEX, bool FUSE = false>
__global__ void __launch_bounds__(kBlock) raycast_brick_kernel(const BrickArgs A) {
    constexpr bool COLOR = (PASS == 2) || FUSE;
    constexpr int U = 4;
    extern __shared__ float4 tf_s[];                 // tf_n entries (dynamic: sized to the function, not to VRDD_MAX_TF)
    for (int i = threadIdx.x; i < A.tf_n; i += kBlock) tf_s[i] = A.tf_tab[i];
    __syncthreads();
    const int blocks_x = (A.iw + 15) / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long nsamp = 0;
    // 16x16-pixel items, block-strided: the launch holds as many blocks as the caller wants resident (launch_brick_pass)
    for (int item = blockIdx.x; item < A.n_items; item += gridDim.x) {
    const int by = item / blocks_x, bx = item - by * blocks_x;
    const int x = bx * 16 + (warp & 1) * 8 + (lane & 7);
    const int y = A.row0 + by * 16 + (warp >> 1) * 4 + (lane >> 3);          // row0 = 0, rows = ih without a window
    if (x < A.iw && y < A.row0 + A.rows) {
        const size_t pix = (size_t)y * A.iw + x;
        const RaySetup R = make_ray(A.m, x, y, A.iw, A.ih, A.ref_rounding);
        float sr = 0.f, sg = 0.f, sb = 0.f;
        const float a_in = (PASS == 2) ? A.alpha_in[pix] : 0.f;
        float sa = a_in;
        bool march = R.hit && !(sa > A.thresh);
        if (PASS == 2 && FUSE && a_in == 0.0f) {             // nothing in front of this brick: pass 1 has marched this pixel
            const float4 v = A.first4[(size_t)(y - A.row0) * A.iw + x];
            sr = v.x; sg = v.y; sb = v.z; sa = v.w;
            march = false;
        }
        if (march) {
            float t = R.tnear;
            float px, py, pz;
            {
                EyeRay E; E.ox = R.ox; E.oy = R.oy; E.oz = R.oz; E.dx = R.dx; E.dy = R.dy; E.dz = R.dz; E.tnear = R.tnear; E.tfar = R.tfar;
                eye_ray_start(E, R.tnear, A.ref_rounding, px, py, pz);
            }
            const float stx = __fmul_rn(R.dx, A.tstep), sty = __fmul_rn(R.dy, A.tstep), stz = __fmul_rn(R.dz, A.tstep);
            // Steps that can fall into this brick: slab test of the ray against the brick's box in
            // world coordinates (texcoord c <-> world 2c-1), widened by two steps.  Outside that range
            // only the recurrence itself is replayed (it must not be restarted, see the header), so a
            // rank pays for the samples it owns rather than for every step of every ray.
            int i_first = 0, i_last = A.max_steps - 1;
            {
                float t0 = -INFINITY, t1 = INFINITY;
                const float o3[3] = {R.ox, R.oy, R.oz}, d3[3] = {R.dx, R.dy, R.dz};
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const float wlo = 2.0f * A.lo[ax] - 1.0f, whi = 2.0f * A.hi[ax] - 1.0f;   // +-inf at the volume faces
                    if (fabsf(d3[ax]) > 1e-12f) {
                        const float a = (wlo - o3[ax]) / d3[ax], b = (whi - o3[ax]) / d3[ax];
                        t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
                    } else if (o3[ax] < wlo - 1e-3f || o3[ax] > whi + 1e-3f) {
                        t1 = -INFINITY;
                    }
                }
                if (t1 >= t0) {
                    const float k0 = floorf((t0 - R.tnear) / A.tstep) - 2.0f, k1 = ceilf((t1 - R.tnear) / A.tstep) + 2.0f;
                    i_first = (int)fminf(fmaxf(k0, 0.0f), (float)A.max_steps);
                    i_last = (int)fminf(fmaxf(k1, -1.0f), (float)(A.max_steps - 1));
                } else {
                    i_last = -1;                                   // the ray never enters this brick
                }
            }
            int i = 0;
            bool alive = i_last >= 0;
            for (; alive && i < i_first; ++i) {                    // replay the recurrence up to the brick (:701-706)
                t = __fadd_rn(t, A.tstep);
                if (t > R.tfar) { alive = false; break; }
                px = __fadd_rn(px, stx); py = __fadd_rn(py, sty); pz = __fadd_rn(pz, stz);
            }
            alive = alive && i <= i_last;
            // the march, in batches of U steps like raycast_kernel: geometry, U fetches in flight, then
            // in-order compositing with the early exit
            while (alive) {
                float cu[U], cv[U], cw[U];
                bool valid[U], mine[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    valid[k] = alive;
                    cu[k] = fmaf(px, 0.5f, 0.5f); cv[k] = fmaf(py, 0.5f, 0.5f); cw[k] = fmaf(pz, 0.5f, 0.5f);
                    mine[k] = alive && cu[k] >= A.lo[0] && cu[k] < A.hi[0] && cv[k] >= A.lo[1] && cv[k] < A.hi[1] &&
                              cw[k] >= A.lo[2] && cw[k] < A.hi[2];
                    const float tn = __fadd_rn(t, A.tstep);
                    const bool cont = alive && !(tn > R.tfar) && (i + 1 <= i_last);
                    if (cont) {
                        t = tn; ++i;
                        px = __fadd_rn(px, stx); py = __fadd_rn(py, sty); pz = __fadd_rn(pz, stz);
                    }
                    alive = cont;
                }
                float s[U];
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    s[k] = 0.f;
                    if (mine[k]) s[k] = TEX ? sample_brick_tex(A, cu[k], cv[k], cw[k]) : sample_brick(A, cu[k], cv[k], cw[k]);
                }
#pragma unroll
                for (int k = 0; k < U; ++k) {
                    if (!valid[k]) { alive = false; break; }
                    if (mine[k]) {
                        float4 col = tf_lookup(tf_s, A.tf_n, (s[k] - A.t_offset) * A.t_scale);
                        if (COUNT) ++nsamp;
                        col.w *= A.density;
                        const float kk = 1.0f - sa;
                        if (COLOR) {
                            col.x *= col.w; col.y *= col.w; col.z *= col.w;
                            sr += col.x * kk; sg += col.y * kk; sb += col.z * kk;
                        }
                        sa += col.w * kk;
                        if (sa > A.thresh) { alive = false; break; }               // :698, on the GLOBAL alpha in pass 2
                    }
                }
            }
        }
        if (PASS == 1 && FUSE) A.first4[(size_t)(y - A.row0) * A.iw + x] = make_float4(sr, sg, sb, sa);
        if (PASS == 2 && A.band_rows > 0) {                  // band owners: this row's increments go to the owner of its band
            const int o = y / A.band_rows;
            reinterpret_cast<float4*>(A.dst[o])[(size_t)(y - o * A.band_rows) * A.iw + x] = make_float4(sr, sg, sb, sa - a_in);
        } else if (A.n_dst > 0) {                            // direct send: my window's rows into slot [brick] of the table(s)
            const size_t wpix = (size_t)(y - A.row0) * A.iw + x;
#pragma unroll 1
            for (int d = 0; d < A.n_dst; ++d) {
                if (PASS == 1) A.dst[d][wpix] = sa;
                else reinterpret_cast<float4*>(A.dst[d])[wpix] = make_float4(sr, sg, sb, sa - a_in);
            }
        } else if (PASS == 1) A.alpha_seg[pix] = sa;
        else A.partial[pix] = make_float4(sr, sg, sb, sa - a_in);
    }
    }
    if (COUNT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nsamp += __shfl_xor_sync(0xffffffffu, nsamp, d);
        if (lane == 0 && nsamp) atomicAdd(A.samples, nsamp);
    }
    if (A.n_dst > 0) {                                       // launch complete -> bump every reader's counter (cf. FrameSignal)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(A.tickets, 1u) == gridDim.x - 1) {
                *A.tickets = 0u;
                __threadfence_system();
#pragma unroll 1
                for (int d = 0; d < A.n_dst; ++d)
                    if (A.flags[d]) asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(A.flags[d]) : "memory");
            }
        }
    }
}

// alpha entering brick (qx,qy,qz): composite, front to back, of the segment alphas of every brick
// that precedes it along the ray in all three axes.  Bricks the ray does not cross hold 0 and
// leave the accumulator untouched, so any linear extension of the axis-wise order is exact.
// Every brick contributes a window of `rows` image rows starting at row0[brick] (the rows its screen footprint
// can touch; the whole image when rows == ih and row0 == 0): seg_rows = float[brick][rows][iw].
struct ViewMatrix { float m[12]; };
constexpr int kMaxBricks = 64;
struct RowWindows { int row0[kMaxBricks]; int rows; };
__global__ void compose_alpha_in_kernel(const float* __restrict__ seg_rows, int gx, int gy, int gz, int qx, int qy, int qz,
                                        float* __restrict__ alpha_in, int iw, int ih, const ViewMatrix M,
                                        const RowWindows Wn, int ref_rounding) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= iw || y >= ih) return;
    const RaySetup R = make_ray(M.m, x, y, iw, ih, ref_rounding);
    const size_t pix = (size_t)y * iw + x;
    const bool fx = R.dx >= 0.f, fy = R.dy >= 0.f, fz = R.dz >= 0.f;
    const int rqx = fx ? qx : gx - 1 - qx, rqy = fy ? qy : gy - 1 - qy, rqz = fz ? qz : gz - 1 - qz;
    float a = 0.f;
    for (int rz = 0; rz <= rqz; ++rz)
        for (int ry = 0; ry <= rqy; ++ry)
            for (int rx = 0; rx <= rqx; ++rx) {
                if (rx == rqx && ry == rqy && rz == rqz) continue;
                const int bx = fx ? rx : gx - 1 - rx, by = fy ? ry : gy - 1 - ry, bz = fz ? rz : gz - 1 - rz;
                const int b = (bz * gy + by) * gx + bx;
                const int yy = y - Wn.row0[b];
                const float s = ((unsigned)yy < (unsigned)Wn.rows) ? seg_rows[((size_t)b * Wn.rows + yy) * iw + x] : 0.0f;
                a += s * (1.0f - a);
            }
    alpha_in[pix] = a;
}

__global__ void pack_frame_kernel(const float4* __restrict__ sum, uint32_t* __restrict__ out, size_t n, float brightness) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = sum[i];
    out[i] = ((uint32_t)(__saturatef(c.w * brightness) * 255.0f) << 24) | ((uint32_t)(__saturatef(c.z * brightness) * 255.0f) << 16) |
             ((uint32_t)(__saturatef(c.y * brightness) * 255.0f) << 8) | (uint32_t)(__saturatef(c.x * brightness) * 255.0f);
}

// root, direct-send form: the frame is the sum, in brick order, of the slots whose window holds the row
__global__ void pack_frame_slots_kernel(const float4* __restrict__ slots, int nb, const RowWindows Wn, uint32_t* __restrict__ out,
                                        int iw, int ih, float brightness) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= iw || y >= ih) return;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < nb; ++b) {
        const int yy = y - Wn.row0[b];
        if ((unsigned)yy < (unsigned)Wn.rows) {
            const float4 v = slots[((size_t)b * Wn.rows + yy) * iw + x];
            c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
        }
    }
    out[(size_t)y * iw + x] = ((uint32_t)(__saturatef(c.w * brightness) * 255.0f) << 24) | ((uint32_t)(__saturatef(c.z * brightness) * 255.0f) << 16) |
                              ((uint32_t)(__saturatef(c.y * brightness) * 255.0f) << 8) | (uint32_t)(__saturatef(c.x * brightness) * 255.0f);
}

// band owner: rows [band0, band0 + band_rows) of the frame = sum, in brick order, of the slots whose window holds the row,
// packed into the (root's, peer-mapped) frame; the last block tells the root
__global__ void pack_band_slots_kernel(const float4* __restrict__ slots, int nb, const RowWindows Wn, int band0, int band_rows,
                                       uint32_t* __restrict__ frame, unsigned* frame_flag, unsigned* tickets, int iw, int ih,
                                       float brightness) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, yb = blockIdx.y, y = band0 + yb;
    if (x < iw && y < ih) {
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int b = 0; b < nb; ++b) {
            if ((unsigned)(y - Wn.row0[b]) < (unsigned)Wn.rows) {
                const float4 v = slots[((size_t)b * band_rows + yb) * iw + x];
                c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
            }
        }
        frame[(size_t)y * iw + x] = ((uint32_t)(__saturatef(c.w * brightness) * 255.0f) << 24) | ((uint32_t)(__saturatef(c.z * brightness) * 255.0f) << 16) |
                                    ((uint32_t)(__saturatef(c.y * brightness) * 255.0f) << 8) | (uint32_t)(__saturatef(c.x * brightness) * 255.0f);
    }
    if (frame_flag) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (atomicAdd(tickets, 1u) == gridDim.x * gridDim.y - 1) {
                *tickets = 0u;
                __threadfence_system();
                asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(frame_flag) : "memory");
            }
        }
    }
}

}  // namespace

int launch_brick_pass(vrdd_context* c, int pass, const float* d_alpha_in, float* d_out, int iw, int ih,
                      const vrdd_render_params& p, const vrdd_brick& b, const BrickSend* send) {
    const int qm = p.query_method;
    if (qm < 1 || qm > 6) return fail(c, VRDD_ERR_UNSUPPORTED, "render_brick: queryMethod must be 1..6");
    const int source = (qm >= 4) ? VRDD_SRC_FRACTAL : VRDD_SRC_ORIGINAL, comp = (qm - 1) % 3;
    vrdd_decoded_volume& vol = c->vol[source];
    if (!vol.decoded || !(vol.lin[comp] || vol.brick[comp] || vol.arr[comp]))
        return fail(c, VRDD_ERR_INVALID, "render_brick: decode the brick first");
    if (iw <= 0 || ih <= 0 || (!d_out && !send) || (pass == 2 && !d_alpha_in)) return fail(c, VRDD_ERR_INVALID, "render_brick: bad arguments");
    if (send && (send->n_dst < 1 || send->n_dst > VRDD_MAX_PEERS + 1 || send->row0 < 0 || send->rows < 1 || send->row0 + send->rows > ih))
        return fail(c, VRDD_ERR_INVALID, "render_brick: bad send window");
    BrickArgs A;
    A.tex = 0;
    if (vol.arr[comp] && c->sampler == VRDD_SAMPLER_TEXTURE) {
        if (!vol.tex_un[comp]) {                               // same array, un-normalised coordinates
            cudaResourceDesc rd;
            std::memset(&rd, 0, sizeof(rd));
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = vol.arr[comp];
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof(td));
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            VRDD_CUDA(c, cudaCreateTextureObject(&vol.tex_un[comp], &rd, &td, nullptr));
        }
        A.tex = vol.tex_un[comp];
    }
    A.bricked = vol.brick[comp] != nullptr;
    A.plane = A.bricked ? vol.brick[comp] : vol.lin[comp];
    if (!A.tex && !A.plane) return fail(c, VRDD_ERR_INVALID, "render_brick: no plane for the current sampler");
    A.bW = c->bW; A.bH = c->bH;
    A.sx = c->W; A.sy = c->H; A.sz = c->D;
    A.gw = b.gw; A.gh = b.gh; A.gd = b.gd; A.ox = b.ox; A.oy = b.oy; A.oz = b.oz;
    for (int i = 0; i < 3; ++i) { A.lo[i] = b.lo[i]; A.hi[i] = b.hi[i]; }
    A.tf_tab = reinterpret_cast<const float4*>(c->tf_dev); A.tf_n = c->tf_n;
    A.iw = iw; A.ih = ih;
    for (int i = 0; i < 12; ++i) A.m[i] = c->view[i];
    A.density = p.density; A.t_offset = p.transfer_offset; A.t_scale = p.transfer_scale; A.tstep = p.tstep;
    A.thresh = p.opacity_threshold; A.max_steps = p.max_steps; A.ref_rounding = c->var_ray_setup;
    A.alpha_in = d_alpha_in; A.alpha_seg = d_out; A.partial = reinterpret_cast<float4*>(d_out);
    A.samples = c->d_samples;
    A.n_dst = 0; A.row0 = 0; A.rows = ih; A.tickets = c->d_tickets;
    for (int d = 0; d <= VRDD_MAX_PEERS; ++d) { A.dst[d] = nullptr; A.flags[d] = nullptr; }
    A.band_rows = 0;
    if (send) {
        if (send->band_rows > 0) {
            if (pass != 2 || (long long)send->band_rows * send->n_dst < ih) return fail(c, VRDD_ERR_INVALID, "render_brick: bands do not cover the frame");
            A.band_rows = send->band_rows;
        }
        A.n_dst = send->n_dst; A.row0 = send->row0; A.rows = send->rows;
        for (int d = 0; d < send->n_dst; ++d) { A.dst[d] = send->dst[d]; A.flags[d] = send->flags[d]; }
    }
    A.n_items = ((iw + 15) / 16) * ((A.rows + 15) / 16);           // only the rows of the window are launched
    // resident blocks: all items at once (0), or var_sortlast_blocks_per_sm per SM striding over the items
    const long long cap = (long long)c->var_sortlast_blocks_per_sm * c->num_sms;
    const int grid = (cap > 0 && cap < A.n_items) ? (int)cap : A.n_items;
    const bool count = c->count_samples && pass == 2;
    // Fused first segment: only in the direct-send form (pass 2 is known to follow pass 1 of the same frame), not while
    // samples are counted (the count is taken in pass 2).  The record is tagged with everything the march depends on; a
    // pass 2 whose tag differs marches every pixel.
    A.first4 = nullptr;
    bool fuse = false;
    if (send && c->var_sortlast_fuse && !c->count_samples) {
        uint64_t tag = 1469598103934665603ull;
        auto mix = [&tag](const void* q, size_t n) {
            const unsigned char* u = static_cast<const unsigned char*>(q);
            for (size_t i = 0; i < n; ++i) { tag ^= u[i]; tag *= 1099511628211ull; }
        };
        mix(c->view, sizeof(c->view)); mix(&p, sizeof(p)); mix(&b, sizeof(b)); mix(&iw, sizeof(iw)); mix(&ih, sizeof(ih));
        mix(&A.row0, sizeof(int)); mix(&A.rows, sizeof(int)); mix(&source, sizeof(source)); mix(&A.tex, sizeof(A.tex));
        tag |= 1ull;                                              // 0 = invalid (api.cu clears the tag when volume / function change)
        const size_t need = (size_t)A.rows * iw;
        if (pass == 1) {
            if (c->first4_cap < need) {
                if (c->d_first4) { cudaFree(c->d_first4); c->d_first4 = nullptr; c->first4_cap = 0; }
                VRDD_CUDA(c, cudaMalloc(&c->d_first4, need * sizeof(float4)));
                c->first4_cap = need;
            }
            c->first4_tag = tag;
            fuse = true;
        } else {
            fuse = c->d_first4 && c->first4_cap >= need && c->first4_tag == tag;
        }
        A.first4 = reinterpret_cast<float4*>(c->d_first4);
    }
    if (fuse) {
        const size_t sm = sizeof(float4) * (size_t)A.tf_n;
        if (A.tex) {
            if (pass == 1) raycast_brick_kernel<1, false, true, true><<<grid, kBlock, sm, c->stream>>>(A);
            else raycast_brick_kernel<2, false, true, true><<<grid, kBlock, sm, c->stream>>>(A);
        } else {
            if (pass == 1) raycast_brick_kernel<1, false, false, true><<<grid, kBlock, sm, c->stream>>>(A);
            else raycast_brick_kernel<2, false, false, true><<<grid, kBlock, sm, c->stream>>>(A);
        }
    } else if (A.tex) {
        if (pass == 1) raycast_brick_kernel<1, false, true><<<grid, kBlock, sizeof(float4) * (size_t)A.tf_n, c->stream>>>(A);
        else if (count) raycast_brick_kernel<2, true, true><<<grid, kBlock, sizeof(float4) * (size_t)A.tf_n, c->stream>>>(A);
        else raycast_brick_kernel<2, false, true><<<grid, kBlock, sizeof(float4) * (size_t)A.tf_n, c->stream>>>(A);
    } else {
        if (pass == 1) raycast_brick_kernel<1, false, false><<<grid, kBlock, sizeof(float4) * (size_t)A.tf_n, c->stream>>>(A);
        else if (count) raycast_brick_kernel<2, true, false><<<grid, kBlock, sizeof(float4) * (size_t)A.tf_n, c->stream>>>(A);
        else raycast_brick_kernel<2, false, false><<<grid, kBlock, sizeof(float4) * (size_t)A.tf_n, c->stream>>>(A);
    }
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_compose_alpha_in(vrdd_context* c, const float* d_seg_rows, int gx, int gy, int gz, int qx, int qy, int qz,
                            const int* row0, int rows, float* d_alpha_in, int iw, int ih) {
    if (!d_seg_rows || !d_alpha_in || iw <= 0 || ih <= 0 || gx < 1 || gy < 1 || gz < 1 || qx < 0 || qx >= gx || qy < 0 ||
        qy >= gy || qz < 0 || qz >= gz)
        return fail(c, VRDD_ERR_INVALID, "compose_alpha_in: bad arguments");
    const long long nb = (long long)gx * gy * gz;
    if (nb > kMaxBricks) return fail(c, VRDD_ERR_UNSUPPORTED, "compose_alpha_in: more than 64 bricks");
    RowWindows Wn;
    Wn.rows = row0 ? rows : ih;
    if (Wn.rows <= 0 || Wn.rows > ih) return fail(c, VRDD_ERR_INVALID, "compose_alpha_in: bad row window");
    for (int b = 0; b < kMaxBricks; ++b) Wn.row0[b] = (row0 && b < nb) ? row0[b] : 0;
    ViewMatrix M;
    for (int i = 0; i < 12; ++i) M.m[i] = c->view[i];
    dim3 grid((iw + 127) / 128, ih);
    compose_alpha_in_kernel<<<grid, 128, 0, c->stream>>>(d_seg_rows, gx, gy, gz, qx, qy, qz, d_alpha_in, iw, ih, M, Wn, c->var_ray_setup);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_pack_frame_slots(vrdd_context* c, const float* d_slots4, int nbricks, const int* row0, int rows, uint32_t* d_out, int iw,
                            int ih, float brightness) {
    if (!d_slots4 || !d_out || !row0 || iw <= 0 || ih <= 0 || nbricks < 1 || nbricks > kMaxBricks || rows < 1 || rows > ih)
        return fail(c, VRDD_ERR_INVALID, "pack_frame_slots: bad arguments");
    RowWindows Wn;
    Wn.rows = rows;
    for (int b = 0; b < kMaxBricks; ++b) Wn.row0[b] = (b < nbricks) ? row0[b] : 0;
    dim3 grid((iw + 127) / 128, ih);
    pack_frame_slots_kernel<<<grid, 128, 0, c->stream>>>(reinterpret_cast<const float4*>(d_slots4), nbricks, Wn, d_out, iw, ih, brightness);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_pack_band_slots(vrdd_context* c, const float* d_slots4, int nbricks, const int* row0, int rows, int band_index, int band_rows,
                           uint32_t* d_frame, uint32_t* d_frame_flag, int iw, int ih, float brightness) {
    if (!d_slots4 || !d_frame || !row0 || iw <= 0 || ih <= 0 || nbricks < 1 || nbricks > kMaxBricks || rows < 1 || rows > ih ||
        band_rows < 1 || band_index < 0)
        return fail(c, VRDD_ERR_INVALID, "pack_band_slots: bad arguments");
    RowWindows Wn;
    Wn.rows = rows;
    for (int b = 0; b < kMaxBricks; ++b) Wn.row0[b] = (b < nbricks) ? row0[b] : 0;
    const int band0 = band_index * band_rows;
    const int nrows = (band0 >= ih) ? 0 : ((band0 + band_rows <= ih) ? band_rows : ih - band0);
    // an empty band (more ranks than bands) still owes the root its signal: one block, no rows
    dim3 grid((iw + 127) / 128, nrows > 0 ? nrows : 1);
    pack_band_slots_kernel<<<grid, 128, 0, c->stream>>>(reinterpret_cast<const float4*>(d_slots4), nbricks, Wn, nrows > 0 ? band0 : ih, band_rows,
                                                        d_frame, d_frame_flag, c->d_tickets + 2, iw, ih, brightness);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

int launch_pack_frame(vrdd_context* c, const float* d_sum4, uint32_t* d_out, int iw, int ih, float brightness) {
    const size_t n = (size_t)iw * ih;
    pack_frame_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const float4*>(d_sum4), d_out, n,
                                                                         brightness);
    c->launches += 1;
    VRDD_CUDA(c, cudaGetLastError());
    return VRDD_OK;
}

}  // namespace vrdd
