/*
 * vrdd.h — C ABI of libvrdd.so: B200-native distribution decode (P1) and volume ray
 * casting (P2).  Plain pointers, sizes and ints only; no C++ or torch types.
 *
 * The library replaces the device side of ykou/Volume-Rendering-Based-on-Distribution-Data
 * behind the reference's own host-facing surface.  Two layers are exported:
 *
 *   1. the handle-based vrdd_* API below — size-generic, explicit stream, error codes,
 *      multi-GPU partitions — which is what new host code binds; and
 *   2. the seven legacy `extern "C"` symbols with the reference's exact signatures
 *      (include/vrdd_legacy.h), implemented on one default handle, so the unmodified
 *      /root/reference/volumeRender.cpp links against libvrdd.so instead of
 *      volumeRender_kernel.cu.
 *
 * Each entry point names the reference interface it replaces (file:line in
 * /root/reference).  All functions return VRDD_OK (0) or a negative vrdd_status; the
 * reference aborts the process on any CUDA error (checkCudaErrors), this library never
 * does.  There is no CPU fallback: without a CUDA device every compute call fails with
 * VRDD_ERR_NO_DEVICE.
 *
 * Threading: a handle is not re-entrant (the reference keeps all state in file-scope
 * globals, volumeRender_kernel.cu:22-43); distinct handles may be used from distinct
 * threads.  All work is enqueued on the handle's stream and is asynchronous unless a
 * function is documented to copy to host memory.
 */
#ifndef VRDD_H_
#define VRDD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRDD_VERSION 100

typedef struct vrdd_context* vrdd_handle;

typedef enum vrdd_status {
    VRDD_OK = 0,
    VRDD_ERR_INVALID = -1,      /* bad argument / call order */
    VRDD_ERR_CUDA = -2,         /* CUDA runtime error; see vrdd_last_error() */
    VRDD_ERR_NO_DEVICE = -3,    /* no usable CUDA device (no CPU fallback exists) */
    VRDD_ERR_RANGE = -4,        /* input violates the reference's run-time guards
                                   (volumeRender_kernel.cu:781-816, volumeRender.cpp:611-614) */
    VRDD_ERR_UNSUPPORTED = -5   /* valid in the reference but outside this build's scope */
} vrdd_status;

/* Which decoded volume: the reference keeps two float4 arrays
 * (originalHistogramData / fractalHistogramData, volumeRender_kernel.cu:719-720). */
typedef enum vrdd_source { VRDD_SRC_ORIGINAL = 0, VRDD_SRC_FRACTAL = 1 } vrdd_source;

/* Ray-caster sampling path (both must agree with the oracle within +-1 LSB). */
typedef enum vrdd_sampler {
    VRDD_SAMPLER_TEXTURE = 0,   /* decoded planes live in 3-D cudaArrays; the texture unit
                                   filters (the reference's own path, :601-651) */
    VRDD_SAMPLER_BRICKED = 1,   /* decoded planes live in a bricked linear layout; manual
                                   trilinear with the texture unit's 8-bit weights */
    VRDD_SAMPLER_LINEAR = 2     /* decoded planes are kept as linear fp32 planes only (x fastest):
                                   what the sort-last brick renderer samples */
} vrdd_sampler;

/* Ray-marching parameters.  Defaults = the reference's compile-time constants
 * (volumeRender_kernel.cu:276-278) and run-time defaults (volumeRender.cpp:130-133). */
typedef struct vrdd_render_params {
    float density;            /* 0.05f  */
    float brightness;         /* 1.0f   */
    float transfer_offset;    /* 0.0f   */
    float transfer_scale;     /* 1.0f   */
    float tstep;              /* 0.01f  */
    int max_steps;            /* 500    */
    float opacity_threshold;  /* 0.95f  */
    int query_method;         /* 1,2,3 = mean/variance/entropy of the original histograms,
                                 4,5,6 = same of the fractal-decoded ones, 7 = interpolated mean,
                                 8/9/0 = entropy/mean/variance of the flexible blocks (volumeRender.cpp:129) */
} vrdd_render_params;

/* Image-space partition for multi-GPU rendering: the image is cut into tile_w x tile_h
 * tiles, numbered row-major, and this call renders tiles with index % parts == part.
 * parts = 1 renders everything.  Pixels of other parts are not touched. */
typedef struct vrdd_tile_partition {
    int tile_w, tile_h;
    int part, parts;
} vrdd_tile_partition;

/* ---- lifetime ----------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` (-1 = current device).  Replaces the
 * file-scope globals of volumeRender_kernel.cu:22-43 and chooseCudaDevice
 * (volumeRender.cpp:1000-1014). */
int vrdd_create(int device, vrdd_handle* out);
/* Releases every device allocation.  Replaces freeCudaBuffers (volumeRender_kernel.cu:2360). */
int vrdd_destroy(vrdd_handle h);
/* Work is enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream,
 * which is what the reference uses everywhere). */
int vrdd_set_stream(vrdd_handle h, void* cuda_stream);
int vrdd_synchronize(vrdd_handle h);
/* Human-readable description of the last error on this handle (never NULL). */
const char* vrdd_last_error(vrdd_handle h);
/* Number of kernels this library has launched on this handle since creation. */
int64_t vrdd_kernel_launches(vrdd_handle h);

/* ---- volume geometry and inputs (replaces initCuda, volumeRender_kernel.cu:1893) ----- */

/* Size of the distribution volume in voxels ("blocks" in the reference, 50x50x10 there,
 * volumeRender.cpp:86) and bins per histogram (32, volumeRender_kernel.cu:91; this build
 * supports bins == 32 only).  Drops any previously decoded data. */
int vrdd_set_volume(vrdd_handle h, int width, int height, int depth, int bins);

/* Raw histograms, float hist[depth*height*width][bins], x fastest
 * (the h_volume argument of initCuda; layout volumeRender_kernel.cu:740-745).
 * _host copies synchronously from host memory like the reference does
 * (volumeRender_kernel.cu:1893-2158); the pointer is not retained.
 * _device borrows a device pointer covering z-slices [z0, z0+nz) only — slab-wise
 * streaming of volumes that do not fit (a 1024^3 volume is 137 GB).  The caller keeps the
 * memory alive until the slab has been decoded. */
int vrdd_set_histograms_host(vrdd_handle h, const float* hist);
int vrdd_set_histograms_device(vrdd_handle h, const float* d_hist, int z0, int nz);

/* Fractal codes (h_codebook, h_templates, h_errorsbook of initCuda).
 *   codebook:  int32[V][4] = (templateId, shift, flip, NE)     volumeRender.cpp:617
 *   errors_dense: float[V][bins][2] = (binId, value), first NE entries of each row valid
 *                                                              volumeRender.cpp:582, 625-635
 *   templates: float[num_templates][bins]                      volumeRender.cpp:661-681
 * The dense error table is compacted on upload to 8*NE bytes per voxel.  Codes are
 * validated against the reference's guards; VRDD_ERR_RANGE reports a violation. */
int vrdd_set_fractal_host(vrdd_handle h, const int32_t* codebook, const float* errors_dense,
                          const float* templates, int num_templates);
/* Compact form of the errors: one 8-byte entry per error, grouped per chunk of 32 consecutive
 * voxels (one warp) and stored ROUND-MAJOR inside a chunk: round k holds the k-th error of every
 * voxel of the chunk with NE > k, in voxel order; round k+1 follows round k.  In round k the
 * lanes that still have an error read consecutive entries, so the decode kernels take the table
 * with one coalesced load per round and no staging. */
typedef struct vrdd_error_entry {
    int32_t bin;        /* (int)binId of the reference's float2, 0 <= bin < bins */
    float value;
} vrdd_error_entry;
/* Host-side packer from the reference's dense table to the compact form (what vrdd_set_fractal_host
 * does internally).  chunk_offsets has ceil(nvox/32)+1 entries; entries may be NULL to only count
 * (*total_ne = sum of NE).  VRDD_ERR_RANGE for NE/shift/bin outside the reference's guards. */
int vrdd_pack_fractal_errors(const int32_t* codebook, const float* errors_dense, int64_t nvox, int bins,
                             vrdd_error_entry* entries, uint64_t* chunk_offsets, uint64_t* total_ne);
/* Compact device-resident form for slab streaming:
 *   d_codebook: int32[nvox][4];  d_errors: vrdd_error_entry[total_ne], round-major per chunk;
 *   d_chunk_offsets: uint64[ceil(nvox/32)+1], entry c = index into d_errors of the first
 *   entry of chunk c (exclusive prefix sum of NE taken every 32 voxels);
 *   d_templates: float[num_templates][bins].  Covers z-slices [z0, z0+nz). */
int vrdd_set_fractal_device(vrdd_handle h, const int32_t* d_codebook, const vrdd_error_entry* d_errors,
                            const uint64_t* d_chunk_offsets, const float* d_templates,
                            int num_templates, int z0, int nz);

/* ---- P1: decode (replaces basicDataProcessing, volumeRender_kernel.cu:1798-1887) ----- */

/* Selects where decoded volumes are kept; call before decoding.  Default TEXTURE. */
int vrdd_set_sampler(vrdd_handle h, int sampler /* vrdd_sampler */);

/* Decodes z-slices [z0, z0+nz) of `source` into the (mean, variance, entropy) volume the
 * ray caster samples.  Stays on the device: the reference's D->H->D round trip
 * (volumeRender_kernel.cu:1805-1862) is gone.  nz <= 0 decodes the slab most recently
 * attached for that source. */
int vrdd_decode(vrdd_handle h, int source /* vrdd_source */, int z0, int nz);

/* Copies the decoded volume of `source` to host memory as float4[V] = (mean, variance,
 * entropy, 0), x fastest — the layout of originalHistogramData / fractalHistogramData
 * (volumeRender_kernel.cu:719-720, 771-773).  Synchronous. */
int vrdd_get_decoded_host(vrdd_handle h, int source, float* out4);
/* Device pointers of the three decoded planes (float[V] each, x fastest), or NULL when the
 * handle keeps them only in cudaArrays (TEXTURE sampler without linear planes).  For
 * NCCL all-gather of z-slabs across ranks. */
int vrdd_get_decoded_planes_device(vrdd_handle h, int source, float** mean, float** variance,
                                   float** entropy);
/* queryMethod 7 (the reference's "interpolated mean", volumeRender_kernel.cu:320-367, 395-480) blends
 * the UN-normalised block means.  Enable before decoding the raw histograms to keep them (4 B/voxel for the
 * linear plane + 4 B/voxel for the array the ray caster fetches from). */
int vrdd_enable_interpolated_mean(vrdd_handle h, int enable);
/* Keep linear planes next to the texture arrays (needed for the call above and for
 * vrdd_commit_planes).  Off by default: the decode then writes the arrays directly. */
int vrdd_keep_linear_planes(vrdd_handle h, int keep);
/* After writing z-slices [z0, z0+nz) of the linear planes from outside (e.g. an NCCL
 * all-gather), republish them to the sampler's layout. */
int vrdd_commit_planes(vrdd_handle h, int source, int z0, int nz);
/* Same for the planes in plane_mask only (bit 0 mean, bit 1 variance, bit 2 entropy): with vrdd_set_peer_planes a
 * rank can replicate the plane the next frames sample first and the others later. */
int vrdd_commit_planes_mask(vrdd_handle h, int source, int z0, int nz, int plane_mask);

/* Bit-comparable view of the integer part of the fractal decode: the reconstructed
 * histogram float[nvox][bins] after template lookup, flip, shift, error merge and clamp,
 * BEFORE normalisation (volumeRender_kernel.cu:798-825), for the attached slab.
 * d_out is device memory. */
int vrdd_reconstruct_fractal_device(vrdd_handle h, float* d_out);

/* ---- P2: ray casting ------------------------------------------------------------------ */

/* 1-D RGBA transfer function, float[n][4], linear, normalised, clamp
 * (baked into initCuda in the reference, volumeRender_kernel.cu:2322-2344).
 * tf == NULL restores the reference's nine-entry rainbow. */
int vrdd_set_transfer_function(vrdd_handle h, const float* tf, int n);
/* Inverse view matrix, row-major 3x4 (copyInvViewMatrix, volumeRender_kernel.cu:2403). */
int vrdd_set_view(vrdd_handle h, const float* m12);
void vrdd_default_render_params(vrdd_render_params* p);

/* Renders into device memory d_output = uint32[image_h][image_w], packed
 * A<<24|B<<16|G<<8|R, row 0 at the bottom (render_kernel, volumeRender_kernel.cu:2387;
 * d_render :272-717).  Like the reference only pixels whose ray hits the volume are
 * written; with clear_misses != 0 missed pixels of this call's partition are written as 0,
 * which folds the caller's cudaMemset (volumeRender.cpp:208) into the kernel.
 * part == NULL renders the whole image.  Asynchronous. */
int vrdd_render(vrdd_handle h, uint32_t* d_output, int image_w, int image_h,
                const vrdd_render_params* params, const vrdd_tile_partition* part,
                int clear_misses);
/* Same, then copies the image to host memory and synchronises: the end-to-end call
 * (render() + the read-back of runSingleTest, volumeRender.cpp:194-217, 1073-1074). */
int vrdd_render_host(vrdd_handle h, uint32_t* h_output, int image_w, int image_h,
                     const vrdd_render_params* params);
/* Pipelined form for a sequence of frames (an orbit, an animation): queues the frame and returns.
 * Frames alternate between two device buffers; a second stream copies frame k to h_output while
 * frame k+1 renders (the render of frame k+2 waits for that copy).  h_output should be page-locked
 * (cudaHostAlloc / cudaHostRegister / vrdd_host_register) for the copy to overlap, and holds the
 * image once vrdd_render_host_wait has returned; use a different h_output for frames that are in
 * flight together.  The view and parameters are taken at the call, so they may change between calls.
 * part == NULL: the whole frame.  Otherwise the partition must consist of full-width row bands
 * (tile_w >= image_w): only this rank's bands are rendered and copied, into their place in the
 * full-frame buffer h_output — with h_output in shared host memory, N ranks fill one frame over
 * N PCIe links (bench.py --gpus N, e2e). */
int vrdd_render_host_async(vrdd_handle h, uint32_t* h_output, int image_w, int image_h,
                           const vrdd_render_params* params, const vrdd_tile_partition* part);
/* Device-side fence: the handle's stream waits until the read-back queued `lag` (0 or 1) calls ago is
 * done; the host does not block.  Put it in front of a per-frame barrier between ranks. */
int vrdd_render_host_fence(vrdd_handle h, int lag);
int vrdd_render_host_wait(vrdd_handle h);
/* cudaHostRegister (portable, mapped) / cudaHostUnregister for caller memory (e.g. a frame in POSIX shared memory).  Mapped:
 * under unified addressing the same pointer is valid in kernels, so vrdd_pack_band_slots may be given a frame in registered
 * host memory as d_frame and stores its rows there directly. */
int vrdd_host_register(void* p, size_t bytes);
int vrdd_host_unregister(void* p);
/* Enables counting of transfer-function lookups (the S of Gsamples/s) in vrdd_render and
 * reads / resets the counter.  Reading synchronises. */
int vrdd_count_samples(vrdd_handle h, int enable);
int vrdd_get_sample_count(vrdd_handle h, int64_t* out, int reset);

/* ---- peer-visible frames: image-space partitions assembled by the ray-cast kernels themselves ------------
 * A frame allocated here can be exported to the other ranks of the node (CUDA IPC).  A rank that opens it
 * gets a device pointer into the owner's HBM and passes it to vrdd_render as d_output together with its
 * vrdd_tile_partition: the kernel then stores its tiles straight into the owner's frame over NVLink — the
 * gather of tiles costs no extra pass and no collective.  The caller orders frames (e.g. one barrier per
 * frame) before the owner reads them. */
#define VRDD_IPC_HANDLE_BYTES 64
int vrdd_frame_alloc(vrdd_handle h, size_t bytes, void** d_frame);
int vrdd_frame_free(vrdd_handle h, void* d_frame);
int vrdd_frame_export(vrdd_handle h, const void* d_frame, unsigned char* ipc_handle /* [64] */);
int vrdd_frame_open(vrdd_handle h, const unsigned char* ipc_handle /* [64] */, void** d_peer_frame);
int vrdd_frame_close(vrdd_handle h, void* d_peer_frame);

/* Frame-complete signal, so that no host barrier is needed per frame: after vrdd_set_frame_signal(h, d_flag) the last
 * block of every vrdd_render launch of this handle adds 1 to *d_flag with system-scope release semantics, once all
 * tiles of the launch are stored — d_flag may point into the frame owner's memory (vrdd_frame_alloc / _open), next to
 * the frame.  The owner orders its stream behind the N ranks with vrdd_stream_wait_flag(h, d_flag, generation * N):
 * the stream (not the host) waits until *d_flag >= at_least.  vrdd_stream_post_flag adds 1 from the stream (e.g. "frame
 * consumed, its buffer may be overwritten", which the other ranks wait for before they render into it again).
 * NULL switches the signal off.  Counters only grow; the caller keeps track of the generation.  A wait gives up after
 * 10 s, so that a rank that died cannot hang the device. */
int vrdd_set_frame_signal(vrdd_handle h, uint32_t* d_flag);
int vrdd_stream_wait_flag(vrdd_handle h, const uint32_t* d_flag, uint32_t at_least);
int vrdd_stream_post_flag(vrdd_handle h, uint32_t* d_flag);
/* vrdd_stream_wait_flag(d_wait, at_least) followed by vrdd_stream_post_flag(d_post), as one stream operation. */
int vrdd_stream_wait_post_flag(vrdd_handle h, const uint32_t* d_wait, uint32_t at_least, uint32_t* d_post);

/* Replication of the decoded planes fused into the decode (N ranks, each decoding its z-slab of a volume that every
 * rank samples): d_planes[3 * q + i] is plane i (mean, variance, entropy) of peer q's LINEAR planes
 * (vrdd_keep_linear_planes + vrdd_get_decoded_planes_device there, exported with vrdd_frame_export and opened here
 * with vrdd_frame_open), or NULL for a plane that is not to be replicated now.  Every later vrdd_decode of `source`
 * also stores each value at its global voxel index in those planes, over NVLink, from the decode kernel itself — no
 * all-gather pass.  When all ranks have synchronised, each commits the slabs it received with vrdd_commit_planes.
 * n_peers = 0 switches it off.  At most 7 peers. */
int vrdd_set_peer_planes(vrdd_handle h, int source, int n_peers, float* const* d_planes);
/* The same for the un-normalised block-mean plane queryMethod 7 interpolates (vrdd_enable_interpolated_mean on every
 * rank): vrdd_get_mean_raw_device gives this rank's plane (to export), vrdd_set_peer_mean_raw attaches the peers'
 * (d_mean_raw[q], same peers and order as vrdd_set_peer_planes; NULL entries are skipped), and after the ranks have
 * synchronised vrdd_commit_mean_raw republishes the received z-slices to the array the ray caster fetches from. */
int vrdd_get_mean_raw_device(vrdd_handle h, float** d_mean_raw);
int vrdd_set_peer_mean_raw(vrdd_handle h, int n_peers, float* const* d_mean_raw);
int vrdd_commit_mean_raw(vrdd_handle h, int z0, int nz);

/* ---- sort-last rendering of a brick-decomposed volume (volumes larger than one GPU's HBM;
 *      new work, the reference is single-GPU; scheme in csrc/sortlast.cu) ------------------------- */

/* This handle's volume (vrdd_set_volume dims; VRDD_SAMPLER_BRICKED, or VRDD_SAMPLER_LINEAR) is one
 * brick of a larger one. */
typedef struct vrdd_brick {
    int gw, gh, gd;           /* size of the GLOBAL volume in voxels                                   */
    int ox, oy, oz;           /* global voxel coordinate of this handle's voxel (0,0,0), ghost included */
    float lo[3], hi[3];       /* this brick owns the samples with lo <= texture coordinate < hi per axis;
                                 use -INFINITY / +INFINITY at the faces of the global volume.  The stored
                                 voxels must cover [lo*N - 1.5, hi*N + 0.5] (one ghost voxel each side) */
} vrdd_brick;

/* Pass 1: alpha accumulated by this brick's own samples, float[image_h][image_w]. */
int vrdd_render_brick_alpha(vrdd_handle h, float* d_alpha_seg, int image_w, int image_h,
                            const vrdd_render_params* params, const vrdd_brick* brick);
/* Alpha entering brick (qx,qy,qz) of a gx x gy x gz grid, from the gathered pass-1 images
 * d_alpha_seg_all = float[gz][gy][gx][image_h][image_w] (brick index x fastest). */
int vrdd_compose_alpha_in(vrdd_handle h, const float* d_alpha_seg_all, int gx, int gy, int gz, int qx, int qy,
                          int qz, float* d_alpha_in, int image_w, int image_h);
/* The same from row windows: brick b contributed only rows [row0[b], row0[b] + rows) of its pass-1 image
 * (the rows its screen footprint can touch, padded to one common count so that an all-gather applies):
 * d_alpha_seg_rows = float[gz][gy][gx][rows][image_w]; row0 = host int[gx*gy*gz]; rows outside a brick's
 * window count as 0.  At most 64 bricks.  vrdd_compose_alpha_in is the case row0 = 0, rows = image_h. */
int vrdd_compose_alpha_in_rows(vrdd_handle h, const float* d_alpha_seg_rows, int gx, int gy, int gz, int qx, int qy,
                               int qz, const int* row0, int rows, float* d_alpha_in, int image_w, int image_h);
/* Pass 2: colour increments (dR, dG, dB, dA) of this brick, float4[image_h][image_w], starting
 * from d_alpha_in with the reference's early exit on the global alpha. */
int vrdd_render_brick_color(vrdd_handle h, const float* d_alpha_in, float* d_partial4, int image_w, int image_h,
                            const vrdd_render_params* params, const vrdd_brick* brick);
/* Sum of all bricks' increments -> RGBA8 frame (brightness, saturate, truncate, pack: :713-716). */
int vrdd_pack_frame(vrdd_handle h, const float* d_sum4, uint32_t* d_output, int image_w, int image_h, float brightness);

/* Direct-send form of the same scheme: the two exchanges travel in the kernels' own stores over NVLink instead of in
 * collectives.  Every rank holds a table of segment alphas float[nbricks][rows][W] and a counter; the root also holds a
 * table of increments float4[nbricks][rows][W] and a counter (vrdd_frame_alloc + _export / _open; double-buffer them by
 * frame parity).  `rows` is the common window height and row0 this brick's first row (vrdd_b200.dist.brick_row_windows).
 *   pass 1   vrdd_render_brick_alpha_send   marches rows [row0, row0 + rows) only and stores them into slot [brick_index] of
 *                                           each of the n_tables tables (d_seg_tables[i], mine included); the last block of
 *                                           the launch adds 1 to every d_flags[i] (release.sys)
 *            vrdd_stream_wait_flag(my counter, nranks * generation), then vrdd_compose_alpha_in_rows on my table
 *   pass 2   vrdd_render_brick_color_send   stores the increments into slot [brick_index] of the root's table and adds 1 to
 *                                           the root's counter
 *   root     vrdd_stream_wait_flag(counter, nranks * generation), vrdd_pack_frame_slots: the frame is the sum of the
 *                                           slots in brick order, packed to RGBA8.
 * Fused first segment: vrdd_render_brick_alpha_send also accumulates colour from alpha 0 and keeps (dR, dG, dB, dA) per
 * pixel in a scratch buffer of the handle; the vrdd_render_brick_color_send* call that follows it for the same view,
 * parameters, window and brick forwards that record for every pixel whose incoming alpha is exactly 0 (nothing precedes
 * the brick on that ray: pass 2 would repeat pass 1's march bit for bit) and marches only the others.  The record is
 * tagged with everything it depends on and dropped when the volume, the transfer function or a variant changes; a pass 2
 * without a matching record marches every pixel.  Off while samples are counted (vrdd_count_samples) and with
 * vrdd_set_variant("sortlast_fuse", "off"). */
int vrdd_render_brick_alpha_send(vrdd_handle h, float* const* d_seg_tables, uint32_t* const* d_flags, int n_tables, int brick_index,
                                 int row0, int rows, int image_w, int image_h, const vrdd_render_params* params,
                                 const vrdd_brick* brick);
int vrdd_render_brick_color_send(vrdd_handle h, const float* d_alpha_in, float* d_root_slots4, uint32_t* d_root_flag, int brick_index,
                                 int row0, int rows, int image_w, int image_h, const vrdd_render_params* params,
                                 const vrdd_brick* brick);
int vrdd_pack_frame_slots(vrdd_handle h, const float* d_slots4, int nbricks, const int* row0, int rows, uint32_t* d_output,
                          int image_w, int image_h, float brightness);

/* Band owners: the same pass 2 with the increments dealt out over ALL ranks instead of converging on the root.  Image rows
 * [o * band_rows, (o + 1) * band_rows) belong to owner o (n_owners * band_rows >= image_h); every owner holds a table
 * float4[nbricks][band_rows][W] and a counter.
 *   pass 2   vrdd_render_brick_color_send_bands   stores each pixel's increment into slot [brick_index] of the table of the
 *                                                 owner of its row, then adds 1 to every owner's counter
 *   owner o  vrdd_stream_wait_flag(my counter, nranks * generation), then vrdd_pack_band_slots: sums the slots of band o in
 *            brick order, packs to RGBA8 and stores the rows into d_frame (the root's frame, peer-mapped: 4 bytes per
 *            pixel cross NVLink); the last block adds 1 to *d_frame_flag (may be NULL)
 *   root     vrdd_stream_wait_flag(frame counter, n_owners * generation): the frame is complete. */
int vrdd_render_brick_color_send_bands(vrdd_handle h, const float* d_alpha_in, float* const* d_owner_slots4, uint32_t* const* d_owner_flags,
                                       int n_owners, int band_rows, int brick_index, int row0, int rows, int image_w, int image_h,
                                       const vrdd_render_params* params, const vrdd_brick* brick);
int vrdd_pack_band_slots(vrdd_handle h, const float* d_slots4, int nbricks, const int* row0, int rows, int band_index, int band_rows,
                         uint32_t* d_frame, uint32_t* d_frame_flag, int image_w, int image_h, float brightness);

/* Host helper: the inverse view matrix the reference builds with OpenGL
 * (volumeRender.cpp:224-246): M = Rx(-rot_x) * Ry(-rot_y) * T(-trans), top three rows,
 * row-major.  Angles in degrees.  (0, 0, (0,0,-4)) is the self-test view (:1024-1043). */
void vrdd_view_matrix(float rot_x_deg, float rot_y_deg, float tx, float ty, float tz, float* m12);

/* ---- synthetic inputs, generated on the device (include/vrdd_synth.h) ------------------ */

/* Fills d_hist = float[nz*height*width][bins] with the seeded synthetic histograms of
 * z-slices [z0, z0+nz) of the handle's volume; bit-identical to the host generator. */
int vrdd_synth_histograms_device(vrdd_handle h, uint32_t seed, int z0, int nz, float* d_hist);
/* Same for a brick: the handle's volume is the sub-box at global voxel offset (ox,oy,oz) of a
 * gw x gh x gd volume; fills local z-slices [z0, z0+nz) with the histograms of the GLOBAL voxels. */
int vrdd_synth_histograms_region_device(vrdd_handle h, uint32_t seed, int gw, int gh, int gd, int ox, int oy,
                                        int oz, int z0, int nz, float* d_hist);
/* Fractal counterpart, in the compact form of vrdd_set_fractal_device.  d_errors must hold max_ne*nvox entries; *total_ne (host) receives
 * the number actually written.  Synchronises. */
int vrdd_synth_fractal_device(vrdd_handle h, uint32_t seed, int num_templates, int max_ne, int z0,
                              int nz, int32_t* d_codebook, vrdd_error_entry* d_errors,
                              uint64_t* d_chunk_offsets, float* d_templates, uint64_t* total_ne);

/* ---- flexible-block-size query chain (csrc/flex.cu) -----------------------------------
 * Replaces dataProcessing() (volumeRender_kernel.cu:1735-1796) and the queryMethod 8/9/0 sampling of d_render.
 * The span tables are the last nine arguments of initCuda (volumeRender_kernel.cu:1896-1900); the reference
 * hard-codes their sizes (131 072 spans, 469 templates, 64 bins, 64^3 raw volume, :96-101), here they are explicit. */
typedef struct vrdd_flex_tables {
    int raw_w, raw_h, raw_d;        /* raw volume the spans refer to (rawVolumeDim, :104); each <= 1023           */
    int bins;                       /* flexNBin = 64 (:97); the only supported value                              */
    int n_fractal;                  /* spans of >= 8 voxels, fractal-coded                                        */
    const int32_t* span_low;        /* int32[n_fractal][4] = (x, y, z, 0), 1-based inclusive   (h_codebookSpanLow)  */
    const int32_t* span_high;       /* int32[n_fractal][4]                                     (h_codebookSpanHigh) */
    const int32_t* codebook;        /* int32[n_fractal][4] = (templateId, shift, flip, NE)     (h_flexibleCodebook) */
    const float* errors;            /* float[n_fractal][bins][2] = (binId, value), NE valid    (h_flexibleErrorsbook) */
    int n_simple;                   /* spans of < 8 voxels, sparse histograms                                      */
    const int32_t* simple_low;      /* int32[n_simple][4], 0-BASED inclusive (:1464-1471)      (h_simpleLow)        */
    const int32_t* simple_high;     /* int32[n_simple][4]                                      (h_simpleHigh)       */
    const int32_t* simple_count;    /* int32[n_simple]                                         (h_simpleCount)      */
    const float* simple_hist;       /* float[n_simple][bins][2] = (binId, frequency)           (h_simpleHistogram)  */
    int n_templates;
    const float* templates;         /* float[n_templates][bins]                                (h_flexibleTemplates) */
} vrdd_flex_tables;

/* Copies and indexes the tables (hash on (low, high) instead of the reference's linear scans); rows whose
 * span_low.x < 0 are padding.  Codes are checked against the loader's guards (VRDD_ERR_RANGE). */
int vrdd_flex_set_tables_host(vrdd_handle h, const vrdd_flex_tables* tables);
/* dataProcessing(): partitions the raw volume into blocks of edge block_size (the reference hard-codes 6, :1737),
 * builds every block's histogram from the spans and leaves (mean, variance, entropy) per block ready for
 * queryMethod 8 (entropy) / 9 (mean) / 0 (variance).  *spans_not_found (may be NULL; reading it synchronises)
 * counts sub-spans missing from the tables (the reference prints "didn't find ..."). */
int vrdd_flex_process(vrdd_handle h, int block_size, int64_t* spans_not_found);
/* float4[nx*ny*nz] = (mean, variance, entropy, 0), blocks x fastest (flexBlockData, :890); dims3 = (nx, ny, nz).
 * out4 may be NULL to query the dimensions. */
int vrdd_flex_get_blocks_host(vrdd_handle h, float* out4, int* dims3);

/* Host stages of the chain, usable on their own:
 * vrdd_flex_divide_blocks — d_divideBlock (:892-1031): partition a vx x vy x vz raw volume into blocks of
 *   edge `block`; spans are 1-based and inclusive, the last block of an axis is clipped; blocks are numbered
 *   x fastest.  spans = int32[n][6] = (lowX, lowY, lowZ, highX, highY, highZ); returns n (or the count
 *   needed when spans == NULL), negative on bad arguments.  Known answer: ver1.9.6.txt:77.
 * vrdd_flex_prefix_spans — the decomposition inside d_queryBlockNew (:1248-1259): the prefix [1, x] as
 *   power-of-two-aligned pieces, peeled from the least significant set bit: 25 -> [25,25], [17,24], [1,16]
 *   (presentation.pdf p.14).  spans = int32[n][2] = (low, high); returns n <= 32. */
int vrdd_flex_divide_blocks(int vx, int vy, int vz, int block, int32_t* spans, int capacity);
int vrdd_flex_prefix_spans(int x, int32_t* spans);

/* ---- diagnostics ------------------------------------------------------------------------ */

/* Selects a kernel variant by name for A/B measurement: "decode_hist" -> "tma" | "ldg";
 * "decode_order" -> "chunked" | "interleaved"; "decode_fractal" -> "moments2" (default) | "moments2r" | "moments2b" |
 * "moments2br" | "moments" | "moments768" | "moments_global" | "dense"; "decode_fractal_prefetch" -> "0".."32"
 * (128-byte lines of the next tile's errors the moments2 kernels pull into L2, default 12);
 * "raycast_tf" -> "texture" | "smem"; "raycast_mode7" -> "gather" (default) | "texture" | "linear": where queryMethod 7
 * reads the block means from — a layered 2-D array with tld4 (where the extents allow it), a point-sampled 3-D
 * array, or the linear plane; "gather" / "texture" decide which array the DECODE fills, so set them before it;
 * "raycast_unroll" -> "1" | "2" | "4" | "8" (ray-march steps whose fetches are in flight
 * together).  Results do not depend on these variants.
 * "ray_setup" -> "nvcc" (default) | "source" is different: it selects how the eye ray of every ray kernel is ROUNDED —
 * what nvcc's default FMA contraction and rsqrt.approx make of it in the reference's own build (the same instructions,
 * hence the same bits: queryMethod 7 then matches the reference binary to +-1 LSB), or the uncontracted order of the
 * reference's source (csrc/common.cuh, eye_ray).  Frames agree to +-1 LSB except in queryMethod 7, whose cell logic
 * is discontinuous.
 * "raycast_layout" -> "auto" (default) | "array" | "layers_x" | "layers_y": which copy of the sampled plane
 * queryMethod 1..6 read — the 3-D array filtered by the texture unit, or a layered copy stacked along x / y read with
 * tld4 and filtered in the kernel with the unit's integer weights (same samples; csrc/raycast.cu,
 * raycast_gather_kernel); auto chooses per view.  Unknown names return VRDD_ERR_INVALID. */
int vrdd_set_variant(vrdd_handle h, const char* what, const char* variant);
/* ---- texture-unit probes: exported only when the library is built with -DVRDD_PROBE_EXPORTS (csrc/Makefile PROBES=1) ---- */
/* Samples the texture unit: out[i] = tex3D(plane `comp` of `source`, u[i], v[i], w[i]) with
 * the ray caster's texture object.  For the filter-model conformance test.  Device ptrs. */
int vrdd_debug_sample_texture(vrdd_handle h, int source, int comp, const float* d_uvw, int n,
                              float* d_out);
/* Same with a POINT-filtered view of the same array (what the reference's block-index texture `tex`
 * is, volumeRender_kernel.cu:2161-2165): used to measure the unit's nearest-texel rule. */
int vrdd_debug_sample_texture_point(vrdd_handle h, int source, int comp, const float* d_uvw, int n,
                                    float* d_out);
/* Same with UN-normalised coordinates and linear filtering (what flexBlockTex is,
 * volumeRender_kernel.cu:1680-1686): d_uvw holds texel-space coordinates. */
int vrdd_debug_sample_texture_unnorm(vrdd_handle h, int source, int comp, const float* d_uvw, int n,
                                     float* d_out);
/* Same for the transfer-function texture: out4[i] = tex1D(transferTex, u[i]).  Device ptrs. */
int vrdd_debug_sample_transfer_function(vrdd_handle h, const float* d_u, int n, float* d_out4);

#ifdef __cplusplus
}
#endif
#endif /* VRDD_H_ */
