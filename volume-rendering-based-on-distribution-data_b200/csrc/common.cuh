// common.cuh — internal declarations shared by the translation units of libvrdd.so.
// sm_100a only.  Nothing here is part of the ABI (see include/vrdd.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/vrdd.h"

#define VRDD_BINS 32                 // volumeRender_kernel.cu:91 (nBins); the only supported value
#define VRDD_ERR_CHUNK 32            // voxels per entry of the fractal error-offset table (one warp)
#define VRDD_MAX_TF 1024             // transfer-function entries kept in shared memory
#define VRDD_MAX_PEERS 7             // other ranks of an 8-GPU node

// Dataset scale constants of the reference (volumeRender_kernel.cu:736, 758-759)
#define VRDD_MAX_HISTOGRAM 0.0217f
#define VRDD_MEAN_NORM 0.0217        // double in the reference
#define VRDD_VAR_NORM 0.000021       // double in the reference

namespace vrdd {

// ---- decoded-volume sink -----------------------------------------------------------------
// Where a decode kernel puts (mean, variance, entropy) of global voxel gv.  Any subset of
// the three destinations may be active.
struct DecodeOut {
    float* lin[3];                   // linear planes float[V], x fastest (or nullptr)
    cudaSurfaceObject_t surf[3];     // 3-D cudaArray planes (valid iff use_surf)
    float* brick[3];                 // bricked planes (or nullptr)
    float* mean_raw;                 // un-normalised block means, linear (queryMethod 7), or nullptr
    int W, H, D;
    int use_surf;
    long long v_base;                // global index of local voxel 0
    int bW, bH;                      // bricks per row / per slice (bricked layout)
    float inv_wh;                    // 1 / (W*H), for the index split in emit_decoded
    // Replication fused into the decode (N > 1, vrdd_set_peer_planes): the linear planes of the OTHER ranks of the node,
    // mapped over CUDA IPC; every decoded value is also stored at its global index in each of them, over NVLink, so the
    // all-gather of the slabs costs no extra pass.  Planes a rank does not want replicated are nullptr.
    int n_peers;
    float* peer[VRDD_MAX_PEERS][3];
    float* peer_mean_raw[VRDD_MAX_PEERS];   // the peers' un-normalised mean planes (queryMethod 7), or nullptr
};

// Bricked layout for the manual sampler: 4x4x4-texel bricks, 256 B each (two 128-B lines),
// bricks stored x-fastest.  The volume is padded up to a multiple of 4 in each axis.
#define VRDD_BRICK 4
#define VRDD_BRICK_SHIFT 2
__host__ __device__ __forceinline__ size_t brick_index(int x, int y, int z, int bW, int bH) {
    size_t b = (size_t)(x >> VRDD_BRICK_SHIFT) +
               (size_t)bW * ((size_t)(y >> VRDD_BRICK_SHIFT) + (size_t)bH * (size_t)(z >> VRDD_BRICK_SHIFT));
    return (b << (3 * VRDD_BRICK_SHIFT)) + (size_t)(((z & 3) << 4) | ((y & 3) << 2) | (x & 3));
}

// (x, y, z) of a global voxel index.  A 64-bit divide costs ~100 instructions, so z comes from a
// float estimate of gv / (W*H) corrected by at most two (exact for gv < 2^47).
__device__ __forceinline__ void split_voxel(const DecodeOut& o, long long gv, int& x, int& y, int& z) {
    const long long wh = (long long)o.W * o.H;
    z = (int)((float)gv * o.inv_wh);
    long long rem = gv - (long long)z * wh;
    if (rem < 0) { --z; rem += wh; }
    else if (rem >= wh) { ++z; rem -= wh; }
    if (rem < 0) { --z; rem += wh; }
    else if (rem >= wh) { ++z; rem -= wh; }
    const int r = (int)rem;
    y = r / o.W;
    x = r - y * o.W;
}

// Stores into the other ranks' linear planes (NVLink; consecutive lanes hold consecutive voxels, so a warp writes
// whole 128-byte lines).  n_peers == 0 on a single GPU: one uniform branch.
__device__ __forceinline__ void emit_to_peers(const DecodeOut& o, long long gv, float mean, float var, float ent) {
    if (o.n_peers == 0) return;                                       // single GPU: one uniform branch per voxel
#pragma unroll 1
    for (int p = 0; p < o.n_peers; ++p) {
        if (o.peer[p][0]) o.peer[p][0][gv] = mean;
        if (o.peer[p][1]) o.peer[p][1][gv] = var;
        if (o.peer[p][2]) o.peer[p][2][gv] = ent;
    }
}

// Same sink with the voxel coordinate supplied by the caller (kernels that walk consecutive
// tiles advance (x, y, z) incrementally instead of dividing per voxel).
__device__ __forceinline__ void emit_decoded_xyz(const DecodeOut& o, long long v_local, int x, int y, int z, float mean,
                                                 float var, float ent) {
    const long long gv = o.v_base + v_local;
    if (o.lin[0]) {
        o.lin[0][gv] = mean; o.lin[1][gv] = var; o.lin[2][gv] = ent;
    }
    if (o.use_surf) {
        surf3Dwrite(mean, o.surf[0], x * 4, y, z);
        surf3Dwrite(var, o.surf[1], x * 4, y, z);
        surf3Dwrite(ent, o.surf[2], x * 4, y, z);
    }
    if (o.brick[0]) {
        const size_t bi = brick_index(x, y, z, o.bW, o.bH);
        o.brick[0][bi] = mean; o.brick[1][bi] = var; o.brick[2][bi] = ent;
    }
    emit_to_peers(o, gv, mean, var, ent);
}

__device__ __forceinline__ void emit_decoded(const DecodeOut& o, long long v_local, float mean, float var,
                                             float ent) {
    const long long gv = o.v_base + v_local;
    if (o.lin[0]) {
        o.lin[0][gv] = mean; o.lin[1][gv] = var; o.lin[2][gv] = ent;
    }
    if (o.use_surf || o.brick[0]) {
        int x, y, z;
        split_voxel(o, gv, x, y, z);
        if (o.use_surf) {
            surf3Dwrite(mean, o.surf[0], x * 4, y, z);
            surf3Dwrite(var, o.surf[1], x * 4, y, z);
            surf3Dwrite(ent, o.surf[2], x * 4, y, z);
        }
        if (o.brick[0]) {
            const size_t bi = brick_index(x, y, z, o.bW, o.bH);
            o.brick[0][bi] = mean; o.brick[1][bi] = var; o.brick[2][bi] = ent;
        }
    }
    emit_to_peers(o, gv, mean, var, ent);
}

// ---- mbarrier / bulk-copy (TMA 1-D) PTX wrappers ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy (SASS: UBLKCP), completion signalled on `bar` in bytes.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// streaming loads that do not pollute L1
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ldg_stream_i4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// p * log2(p) with the reference's `p <= 0 ? 0` guard (volumeRender_kernel.cu:765-766):
// log2(max(p, tiny)) is finite, and 0 * finite == 0.  MUFU.LG2 replaces logf(p)/log(2.0).
// lg2.approx.ftz is one MUFU.LG2; __log2f (non-ftz) wraps it in a denormal rescue (FSETP + 2 FMUL + FADD)
// that 32 bins per voxel pay for in an issue-bound kernel.  The argument is clamped to >= 1e-37 first.
__device__ __forceinline__ float fast_log2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float plog2p(float p) { return p * fast_log2(fmaxf(p, 1.0e-37f)); }

// ---- which pixel a thread renders: image-space tiles, 16x16-pixel blocks, 8x4 pixels per warp ----------------------
// The image is cut into tile_w x tile_h tiles numbered row-major; a launch renders the tiles with index % parts == part
// (vrdd_tile_partition).  Every tile is covered by blocks of 256 threads = 16x16 pixels, each warp an 8x4 sub-tile, so
// the eight texels around the 32 samples of one step share cache lines.  Shared by every ray kernel of vrdd_render.
struct TileMap {
    int iw, ih;
    int tile_w, tile_h, tiles_x, part, parts;
    int blocks_x, blocks_per_tile;                                   // 16x16-pixel blocks inside a tile
    int n_items;                                                     // 16x16-pixel blocks this launch renders
    __device__ __forceinline__ bool pixel(int item, int& x, int& y) const {
        const int lt = item / blocks_per_tile;                       // my tile number
        const int bt = item - lt * blocks_per_tile;                  // block inside the tile
        const int gt = part + lt * parts;                            // global tile index
        const int ty = gt / tiles_x, tx = gt - ty * tiles_x;
        const int by = bt / blocks_x, bx = bt - by * blocks_x;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int lx = bx * 16 + (warp & 1) * 8 + (lane & 7);
        const int ly = by * 16 + (warp >> 1) * 4 + (lane >> 3);
        x = tx * tile_w + lx; y = ty * tile_h + ly;
        return lx < tile_w && ly < tile_h && x < iw && y < ih;
    }
};

// ---- persistent blocks, work queue and frame-complete signal ------------------------------------------------------
// A ray kernel is launched with as many 256-thread blocks as the device holds at once; each block takes 16x16-pixel
// items from a queue (tickets[1]) until none is left, then draws a ticket (tickets[0]); the block that draws the last
// one resets both counters for the next launch of this context's stream.
// Frame-complete signal (vrdd_set_frame_signal): N ranks store their tiles into one frame in rank 0's memory; instead of
// a host barrier per frame, the last block of each rank's launch bumps a counter next to that frame with system-scope
// release semantics, and the owner's stream waits for the count (vrdd_stream_wait_flag).  Every block: stores ->
// __syncthreads -> system fence -> ticket; the block that draws the last ticket has (cumulatively) all stores of the
// launch before it and publishes.  Persistent blocks keep that to one fence per resident block instead of one per item.
struct FrameSignal {
    unsigned* flag;            // device pointer, possibly into a peer's memory; nullptr = no signal
    unsigned* tickets;         // this context's counters: [0] finished blocks, [1] next item (both zero between launches)
    // The queue: thread 0 draws the NEXT item while the block works on the current one, so the atomic's latency is hidden
    // and an item costs one __syncthreads.  for (item = queue_begin(); item < n; item = queue_next(it++)) { queue_prefetch(it); ... }
    __device__ __forceinline__ int* queue_slots() const {
        __shared__ int s_item[2];
        return s_item;
    }
    __device__ __forceinline__ int queue_begin() const {             // the same value for every thread of the block
        int* s = queue_slots();
        if (threadIdx.x == 0) s[0] = (int)atomicAdd(tickets + 1, 1u);
        __syncthreads();
        return s[0];
    }
    __device__ __forceinline__ void queue_prefetch(int it) const {
        if (threadIdx.x == 0) queue_slots()[(it + 1) & 1] = (int)atomicAdd(tickets + 1, 1u);
    }
    __device__ __forceinline__ int queue_next(int it) const {
        __syncthreads();                                               // the slot is written; everybody is done with the other one
        return queue_slots()[(it + 1) & 1];
    }
    __device__ __forceinline__ void block_done() const {
        __syncthreads();
        if (threadIdx.x == 0) {
            if (flag) __threadfence_system(); else __threadfence();
            if (atomicAdd(tickets, 1u) == gridDim.x - 1) {
                tickets[0] = 0u; tickets[1] = 0u;                      // every block has left its loop: the next launch starts from zero
                if (flag) {
                    __threadfence_system();
                    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
                }
            }
        }
    }
};

// ---- eye ray of a pixel (volumeRender_kernel.cu:282-312, intersectBox :136-156, mul :168-184) -----------------
// Shared by every ray kernel (raycast.cu, sortlast.cu, flex.cu), so all of them — and every rank — agree bit for bit
// on direction, tnear, tfar and the first sample position.  Two roundings, chosen at run time:
//   ref_rounding = 1 (default, variant ray_setup = "nvcc"): what nvcc 12.9 generates for the reference's d_render
//     with its default -fmad=true, read off the PTX of the reference compiled where it lies (oracle/Makefile `ref`):
//     u*u + v*v fused, + 4, rsqrt.approx (helper_math.h normalize()); each component of M*dir as
//     fma(dir.z, m.z, fma(dir.x, m.x, dir.y*m.y)); pos = fma(d, tnear, o).  These are the instructions of the
//     reference's own binary, hence its bits: measured on a B200 against tests/golden/ref_gpu_v1.npz, queryMethod 7
//     (which amplifies the last bit of a position at every cell boundary) has no byte off by more than 1 LSB with
//     this rounding and 3-4 % of its bytes off with the other one.
//   ref_rounding = 0 ("source"): the source's expressions without contraction and with an IEEE 1/sqrt — the
//     oracle's default order.
// The slab test is the same in both: IEEE reciprocals and products, as in the reference's PTX.
struct EyeRay { float ox, oy, oz, dx, dy, dz, tnear, tfar; };      // tnear not yet clamped to 0 (:305-306)
__device__ __forceinline__ EyeRay eye_ray(const float* m, int x, int y, int iw, int ih, int ref_rounding) {
    EyeRay R;
    const float u = __fsub_rn(__fmul_rn(__fdiv_rn((float)x, (float)iw), 2.0f), 1.0f);     // :288 (the same value fused or not)
    const float v = __fsub_rn(__fmul_rn(__fdiv_rn((float)y, (float)ih), 2.0f), 1.0f);
    R.ox = m[3]; R.oy = m[7]; R.oz = m[11];
    if (ref_rounding) {
        const float dd = __fadd_rn(fmaf(u, u, __fmul_rn(v, v)), 4.0f);
        float inv;
        asm("rsqrt.approx.f32 %0, %1;" : "=f"(inv) : "f"(dd));                            // as in the reference's build
        const float a = __fmul_rn(u, inv), b = __fmul_rn(v, inv), c = __fmul_rn(inv, -2.0f);
        R.dx = fmaf(c, m[2], fmaf(a, m[0], __fmul_rn(b, m[1])));
        R.dy = fmaf(c, m[6], fmaf(a, m[4], __fmul_rn(b, m[5])));
        R.dz = fmaf(c, m[10], fmaf(a, m[8], __fmul_rn(b, m[9])));
    } else {
        float dx0 = u, dy0 = v, dz0 = -2.0f;
        const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(dx0, dx0), __fmul_rn(dy0, dy0)), __fmul_rn(dz0, dz0));
        const float inv_len = __fdiv_rn(1.0f, __fsqrt_rn(len2));
        dx0 = __fmul_rn(dx0, inv_len); dy0 = __fmul_rn(dy0, inv_len); dz0 = __fmul_rn(dz0, inv_len);
        R.dx = __fadd_rn(__fadd_rn(__fmul_rn(dx0, m[0]), __fmul_rn(dy0, m[1])), __fmul_rn(dz0, m[2]));
        R.dy = __fadd_rn(__fadd_rn(__fmul_rn(dx0, m[4]), __fmul_rn(dy0, m[5])), __fmul_rn(dz0, m[6]));
        R.dz = __fadd_rn(__fadd_rn(__fmul_rn(dx0, m[8]), __fmul_rn(dy0, m[9])), __fmul_rn(dz0, m[10]));
    }
    // slab test against [-1,1]^3 (:136-156)
    const float ix = __fdiv_rn(1.0f, R.dx), iy = __fdiv_rn(1.0f, R.dy), iz = __fdiv_rn(1.0f, R.dz);
    const float bx0 = __fmul_rn(ix, __fsub_rn(-1.0f, R.ox)), bx1 = __fmul_rn(ix, __fsub_rn(1.0f, R.ox));
    const float by0 = __fmul_rn(iy, __fsub_rn(-1.0f, R.oy)), by1 = __fmul_rn(iy, __fsub_rn(1.0f, R.oy));
    const float bz0 = __fmul_rn(iz, __fsub_rn(-1.0f, R.oz)), bz1 = __fmul_rn(iz, __fsub_rn(1.0f, R.oz));
    const float tminx = fminf(bx1, bx0), tminy = fminf(by1, by0), tminz = fminf(bz1, bz0);
    const float tmaxx = fmaxf(bx1, bx0), tmaxy = fmaxf(by1, by0), tmaxz = fmaxf(bz1, bz0);
    R.tnear = fmaxf(fmaxf(tminx, tminy), fmaxf(tminx, tminz));
    R.tfar = fminf(fminf(tmaxx, tmaxy), fminf(tmaxx, tmaxz));
    return R;
}
// first sample position o + d * tnear (:311)
__device__ __forceinline__ void eye_ray_start(const EyeRay& R, float tnear, int ref_rounding, float& px, float& py, float& pz) {
    if (ref_rounding) {
        px = fmaf(R.dx, tnear, R.ox); py = fmaf(R.dy, tnear, R.oy); pz = fmaf(R.dz, tnear, R.oz);
    } else {
        px = __fadd_rn(R.ox, __fmul_rn(R.dx, tnear)); py = __fadd_rn(R.oy, __fmul_rn(R.dy, tnear)); pz = __fadd_rn(R.oz, __fmul_rn(R.dz, tnear));
    }
}

}  // namespace vrdd

struct vrdd_flex_state;                  // flexible-block chain (flex.cu)

// ---- the context behind vrdd_handle --------------------------------------------------------
struct vrdd_decoded_volume {
    float* lin[3] = {nullptr, nullptr, nullptr};
    cudaArray_t arr[3] = {nullptr, nullptr, nullptr};
    cudaTextureObject_t tex[3] = {0, 0, 0};
    cudaTextureObject_t tex_un[3] = {0, 0, 0};   // same arrays, un-normalised coordinates (sort-last bricks)
    cudaSurfaceObject_t surf[3] = {0, 0, 0};
    float* brick[3] = {nullptr, nullptr, nullptr};
    float* mean_raw = nullptr;       // un-normalised bin-centre mean per block, linear (queryMethod 7)
    cudaArray_t mean_arr = nullptr;  // the same plane as a 3-D array read with point fetches: the march of
    cudaTextureObject_t mean_tex = 0;   // queryMethod 7 wants the texture path's 3-D locality (side views)
    cudaArray_t mean_lay = nullptr;  // and as a layered 2-D array (layer = z) read with tld4: the 2x2 (x, y) corners
    cudaTextureObject_t mean_gather = 0;   // of a cell in one fetch (variant raycast_mode7 = gather)
    // Copies of a plane for the gather path of the ray caster (raycast.cu, raycast_gather_kernel): DRAM delivers whole
    // 128-byte lines = 8 x 4 x 1 texels of the 3-D array above, thin in z only, so only views along z skip the slices
    // between two samples.  A copy is a layered 2-D array stacked along x (index 1) or y (index 2), read with tld4,
    // for views along that axis; it is built lazily from the 3-D array the first time a view selects it and refilled
    // after the source is decoded again.
    cudaArray_t garr[3][3] = {};                 // [plane][stacking axis]; [.][0] unused (z = the 3-D array itself)
    cudaTextureObject_t gtex[3][3] = {};
    cudaSurfaceObject_t gsurf[3][3] = {};
    bool gvalid[3][3] = {};
    bool decoded = false;
};

struct vrdd_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int num_sms = 148;

    int W = 0, H = 0, D = 0, B = VRDD_BINS;
    size_t V = 0;
    int bW = 0, bH = 0, bD = 0;      // bricks per axis

    // raw histograms (attached slab)
    float* hist_owned = nullptr;
    const float* hist = nullptr;
    int hist_z0 = 0, hist_nz = 0;

    // fractal codes (attached slab)
    int32_t* cb_owned = nullptr;
    void* err_owned = nullptr;
    uint64_t* off_owned = nullptr;
    float* tmpl_owned = nullptr;
    const int32_t* cb = nullptr;
    const void* errs = nullptr;          // vrdd_error_entry[], round-major per 32-voxel chunk
    const uint64_t* err_off = nullptr;
    const float* tmpl = nullptr;
    int num_templates = 0;
    int fr_z0 = 0, fr_nz = 0;
    void* tmpl_mom = nullptr;        // per-(template, flip, shift) centred moments + {value, g(value)} rows (decode_fractal.cu)

    vrdd_decoded_volume vol[2];
    int sampler = VRDD_SAMPLER_TEXTURE;
    bool keep_linear = false;
    bool keep_mean_raw = false;      // vrdd_enable_interpolated_mean

    // transfer function
    cudaArray_t tf_arr = nullptr;
    cudaTextureObject_t tf_tex = 0;
    float* tf_dev = nullptr;         // float4[tf_n]
    int tf_n = 0;

    float view[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 4};
    bool legacy_linear_filter = true;

    unsigned long long* d_samples = nullptr;
    bool count_samples = false;

    void* d_first4 = nullptr;        // sort-last, fused first segment: float4[rows][iw] record of pass 1 (sortlast.cu)
    size_t first4_cap = 0;           // ... its capacity in pixels
    unsigned long long first4_tag = 0;   // ... what it was computed for (0 = invalid)
    unsigned* d_tickets = nullptr;   // {finished blocks, next item} of the persistent ray kernels (FrameSignal); [2] = finished blocks
                                     // of pack_band_slots_kernel, which may run on another stream beside a ray kernel
    unsigned* frame_signal = nullptr;   // vrdd_set_frame_signal: bumped by the last block of every vrdd_render launch

    int n_peers[2] = {0, 0};         // vrdd_set_peer_planes: the other ranks' linear planes, per source
    float* peer_planes[2][VRDD_MAX_PEERS][3] = {};
    float* peer_mean_raw[VRDD_MAX_PEERS] = {};   // vrdd_set_peer_mean_raw (same peers as source 0)

    vrdd_flex_state* flex = nullptr; // span store + block volume of the flexible-block chain

    uint32_t* frame = nullptr;       // device frame of vrdd_render_host, kept between calls
    size_t frame_bytes = 0;
    uint32_t* pframe[2] = {nullptr, nullptr};   // vrdd_render_host_async: two device frames, a copy stream and
    size_t pframe_bytes = 0;                    // the events that order render k+2 after copy k
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_rendered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    unsigned pslot = 0;

    // kernel variants (vrdd_set_variant)
    int var_decode_hist = 0;         // 0 tma, 1 ldg
    int var_decode_order = 1;        // tma tile order: 0 interleaved over CTAs, 1 one contiguous run per CTA (TLB-friendly, default)
    int var_tf = 1;                  // 0 texture unit, 1 shared-memory table (default: frees the TEX pipe)
    int var_unroll = 4;              // ray-march batch: steps whose fetches are in flight together (1,2,4,8)
    int var_gather_unroll = 2;       // march batch of raycast_gather_kernel (2: 48 registers, five blocks per SM — measured faster than 4)
    int var_gather_tf = -1;          // transfer function of raycast_gather_kernel: -1 follow var_tf, 0 texture unit, 1 shared-memory table
    int var_sortlast_blocks_per_sm = 3;   // sort-last brick kernel: resident blocks per SM striding over the 16x16-pixel items (0 = one block
                                          // per item; N = 2, same box: 0 -> 1432, 2 -> 1413, 3 -> 1562, 4 -> 1559, 5 -> 1370 fps)
    int var_sortlast_fuse = 1;       // direct-send sort-last: pass 1 keeps the colour of the march from alpha 0, pass 2 skips those pixels
    int var_array_blocks_per_sm = -1; // 3-D array ray kernel: resident 256-thread blocks per SM (-1 = by ray spacing, 0 = as many as fit); at persist_pct 100
    float var_layout_min_spacing = 1.5f;  // auto: a layered copy only if neighbouring rays are at least this many voxels apart
    int var_persist_pct = 100;       // ray kernels: blocks launched in percent of the resident capacity (0: one block per item)
    int var_layout = 0;              // array the ray caster samples: 0 auto (per view, launch_raycast), 1 the 3-D array (texture unit
                                     // filters), 2 / 3 the layered copy stacked along x / y (tld4 + the unit's integer weights in the kernel)
    float var_layout_min_step = 2.5f;   // auto: a copy only if a ray advances more than this many voxels per step along its stacking axis
    float var_layout_cos = 0.74f;       // ... and the view direction is within acos(this) of that axis (42 degrees; 47 before the 3-D array
                                        // kernel ran with two blocks per SM)
    int var_ray_setup = 1;           // 1 "nvcc" (default): the rounding of the reference's own build; 0 "source": the source's uncontracted order (eye_ray above)
    int var_mode7 = 2;               // 2 layered array + tld4 where the volume allows it (default), 0 point-sampled 3-D array, 1 linear plane
    int var_fractal_sink = 1;        // moments2: 1 = the surfaces-only instance where the sink is just the three 3-D arrays; 0 = generic
    int var_fractal_pf = 12;         // moments2: 128-byte lines of the next tile's errors prefetched into L2 (0..32)
    int var_fractal = 4;             // 0 dense (O(B) per voxel); moments (O(NE) per voxel): 1 r1f kernel, 2 tables in global memory,
                                     // 3 768 threads, 4 moments2 (default), 5 moments2r, 6 moments2b, 7 moments2br (decode_fractal.cu)
};

namespace vrdd {

int fail(vrdd_context* c, int code, const char* what);
int fail_cuda(vrdd_context* c, cudaError_t e, const char* what);
#define VRDD_CUDA(c, call)                                                      \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return ::vrdd::fail_cuda((c), e__, #call);      \
    } while (0)

DecodeOut make_decode_out(vrdd_context* c, int source, long long v_base);
int ensure_volume_storage(vrdd_context* c, int source);

// kernels' host launchers (one per .cu)
int launch_decode_hist(vrdd_context* c, const float* d_hist, long long nvox, const DecodeOut& out);
int launch_decode_fractal(vrdd_context* c, const int32_t* cb, const void* errs, const uint64_t* off,
                          const float* tmpl, int T, long long nvox, const DecodeOut& out, float* d_recon);
int build_template_moments(vrdd_context* c, const float* d_tmpl, int T);
bool point_rule_is_regular(int n);       // raycast.cu: does the point rule map boundary k to texel min(k, n - 1)?
int launch_raycast(vrdd_context* c, uint32_t* d_out, int iw, int ih, const vrdd_render_params& p,
                   const vrdd_tile_partition& part, int clear_misses);
// host side of TileMap: fills `tm`, returns the grid size (0: this part owns no tile, -1: bad partition / too large)
long long make_tile_map(int iw, int ih, const vrdd_tile_partition& part, TileMap* tm);
int launch_stream_wait_flag(vrdd_context* c, const unsigned* d_flag, unsigned at_least, unsigned* d_post);
int launch_stream_post_flag(vrdd_context* c, unsigned* d_flag);
void invalidate_gather_copies(vrdd_decoded_volume& v);   // after a decode / commit: the copies no longer match the 3-D arrays
int launch_synth_hist(vrdd_context* c, uint32_t seed, int z0, int nz, float* d_hist);
int launch_synth_fractal(vrdd_context* c, uint32_t seed, int T, int max_ne, int z0, int nz, int32_t* d_cb,
                         vrdd_error_entry* d_err, uint64_t* d_off, float* d_tmpl, uint64_t* total_ne);
int launch_synth_hist_region(vrdd_context* c, uint32_t seed, int gw, int gh, int gd, int ox, int oy, int oz, int z0,
                             int nz, float* d_hist);
// direct-send form of the sort-last passes (sortlast.cu): where this launch's window rows go, and whom it tells
struct BrickSend {
    int n_dst;
    float* dst[VRDD_MAX_PEERS + 1];
    unsigned* flags[VRDD_MAX_PEERS + 1];
    int row0, rows;
    int band_rows = 0;               // pass 2, band owners: dst[o] = slot [brick] of the owner of image rows [o * band_rows, (o + 1) * band_rows)
};
int launch_brick_pass(vrdd_context* c, int pass, const float* d_alpha_in, float* d_out, int iw, int ih,
                      const vrdd_render_params& p, const vrdd_brick& b, const BrickSend* send = nullptr);
int launch_pack_band_slots(vrdd_context* c, const float* d_slots4, int nbricks, const int* row0, int rows, int band_index, int band_rows,
                           uint32_t* d_frame, uint32_t* d_frame_flag, int iw, int ih, float brightness);
int launch_pack_frame_slots(vrdd_context* c, const float* d_slots4, int nbricks, const int* row0, int rows, uint32_t* d_out, int iw,
                            int ih, float brightness);
int launch_compose_alpha_in(vrdd_context* c, const float* d_seg_rows, int gx, int gy, int gz, int qx, int qy, int qz,
                            const int* row0, int rows, float* d_alpha_in, int iw, int ih);
int launch_pack_frame(vrdd_context* c, const float* d_sum4, uint32_t* d_out, int iw, int ih, float brightness);
void destroy_flex(vrdd_context* c);
int launch_raycast_flex(vrdd_context* c, uint32_t* d_out, int iw, int ih, const vrdd_render_params& p,
                        const vrdd_tile_partition& part, int clear_misses);
int launch_debug_sample(vrdd_context* c, cudaTextureObject_t tex, const float* d_uvw, int n, float* d_out);
int launch_debug_sample_tf(vrdd_context* c, const float* d_u, int n, float* d_out4);

}  // namespace vrdd
