// api.cu — the handle-based C ABI of include/vrdd.h: context, device storage, uploads,
// decode / render orchestration.  Host code only; kernels live in the sibling files.
#include "common.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace vrdd {

int fail(vrdd_context* c, int code, const char* what) {
    if (c) c->err = what;
    return code;
}
int fail_cuda(vrdd_context* c, cudaError_t e, const char* what) {
    if (c) c->err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return VRDD_ERR_CUDA;
}

namespace {

// The hard-coded rainbow of the reference (volumeRender_kernel.cu:2323-2326)
const float kDefaultTf[9][4] = {{0, 0, 0, 0}, {1, 0, 0, 1}, {1, 0.5f, 0, 1}, {1, 1, 0, 1}, {0, 1, 0, 1},
                                {0, 1, 1, 1}, {0, 0, 1, 1}, {1, 0, 1, 1}, {0, 0, 0, 0}};

void free_volume(vrdd_decoded_volume& v) {
    for (int i = 0; i < 3; ++i) {
        if (v.tex[i]) cudaDestroyTextureObject(v.tex[i]);
        if (v.tex_un[i]) cudaDestroyTextureObject(v.tex_un[i]);
        if (v.surf[i]) cudaDestroySurfaceObject(v.surf[i]);
        if (v.arr[i]) cudaFreeArray(v.arr[i]);
        if (v.lin[i]) cudaFree(v.lin[i]);
        if (v.brick[i]) cudaFree(v.brick[i]);
    }
    for (int i = 0; i < 3; ++i)
        for (int a = 0; a < 3; ++a) {
            if (v.gtex[i][a]) cudaDestroyTextureObject(v.gtex[i][a]);
            if (v.gsurf[i][a]) cudaDestroySurfaceObject(v.gsurf[i][a]);
            if (v.garr[i][a]) cudaFreeArray(v.garr[i][a]);
        }
    if (v.mean_raw) cudaFree(v.mean_raw);
    if (v.mean_tex) cudaDestroyTextureObject(v.mean_tex);
    if (v.mean_arr) cudaFreeArray(v.mean_arr);
    if (v.mean_gather) cudaDestroyTextureObject(v.mean_gather);
    if (v.mean_lay) cudaFreeArray(v.mean_lay);
    v = vrdd_decoded_volume();
}

void free_inputs(vrdd_context* c) {
    if (c->hist_owned) cudaFree(c->hist_owned);
    if (c->cb_owned) cudaFree(c->cb_owned);
    if (c->err_owned) cudaFree(c->err_owned);
    if (c->off_owned) cudaFree(c->off_owned);
    if (c->tmpl_owned) cudaFree(c->tmpl_owned);
    c->hist_owned = nullptr; c->cb_owned = nullptr; c->err_owned = nullptr; c->off_owned = nullptr;
    c->tmpl_owned = nullptr;
    c->hist = nullptr; c->cb = nullptr; c->errs = nullptr; c->err_off = nullptr; c->tmpl = nullptr;
    c->hist_nz = 0; c->fr_nz = 0; c->num_templates = 0;
    if (c->tmpl_mom) cudaFree(c->tmpl_mom);
    c->tmpl_mom = nullptr;
}

void free_tf(vrdd_context* c) {
    if (c->tf_tex) cudaDestroyTextureObject(c->tf_tex);
    if (c->tf_arr) cudaFreeArray(c->tf_arr);
    if (c->tf_dev) cudaFree(c->tf_dev);
    c->tf_tex = 0; c->tf_arr = nullptr; c->tf_dev = nullptr; c->tf_n = 0;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t brick_elems(const vrdd_context* c) {
    return (size_t)c->bW * c->bH * c->bD * (VRDD_BRICK * VRDD_BRICK * VRDD_BRICK);
}

}  // namespace

// Allocate whatever the current sampler / keep_linear setting needs for `source`.
int ensure_volume_storage(vrdd_context* c, int source) {
    vrdd_decoded_volume& v = c->vol[source];
    const bool want_tex = c->sampler == VRDD_SAMPLER_TEXTURE;
    const bool want_brick = c->sampler == VRDD_SAMPLER_BRICKED;
    const bool want_lin = c->keep_linear || c->sampler == VRDD_SAMPLER_LINEAR;
    for (int i = 0; i < 3; ++i) {
        if (want_tex && (!v.arr[i] || !v.tex[i] || !v.surf[i])) {
            cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
            if (!v.arr[i])
                VRDD_CUDA(c, cudaMalloc3DArray(&v.arr[i], &desc, make_cudaExtent(c->W, c->H, c->D),
                                               cudaArraySurfaceLoadStore));
            cudaResourceDesc rd;
            std::memset(&rd, 0, sizeof(rd));
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = v.arr[i];
            // linear, normalised, clamp: originalQueryTex / fractalQueryTex
            // (volumeRender_kernel.cu:1865-1876; the third axis keeps the default clamp)
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof(td));
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 1;
            if (!v.tex[i]) VRDD_CUDA(c, cudaCreateTextureObject(&v.tex[i], &rd, &td, nullptr));
            if (!v.surf[i]) VRDD_CUDA(c, cudaCreateSurfaceObject(&v.surf[i], &rd));
        }
        if (want_lin && !v.lin[i]) VRDD_CUDA(c, cudaMalloc(&v.lin[i], sizeof(float) * c->V));
        if (want_brick && !v.brick[i]) {
            VRDD_CUDA(c, cudaMalloc(&v.brick[i], sizeof(float) * brick_elems(c)));
            VRDD_CUDA(c, cudaMemsetAsync(v.brick[i], 0, sizeof(float) * brick_elems(c), c->stream));
        }
    }
    // queryMethod 7 reads the block means through ONE of two arrays: a layered 2-D array fetched with tld4 when
    // the extents allow it (layer limit 2048, regular point rule in x and y: raycast_mode7_kernel<GATHER>),
    // else a point-sampled 3-D array; the linear plane is kept in both cases (decode target, fallback).
    const bool want_gather = c->var_mode7 == 2 && c->D <= 2048 && c->W <= 32768 && c->H <= 32768 &&
                             point_rule_is_regular(c->W) && point_rule_is_regular(c->H);
    // every resource below is created only if it does not exist yet, so a call that failed half-way can be repeated
    if (c->keep_mean_raw && source == VRDD_SRC_ORIGINAL && !v.mean_raw && want_gather && !v.mean_gather) {
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        // (cudaArrayTextureGather is rejected together with cudaArrayLayered; tld4.a2d itself does not need it)
        if (v.mean_lay || cudaMalloc3DArray(&v.mean_lay, &desc, make_cudaExtent(c->W, c->H, c->D), cudaArrayLayered) == cudaSuccess) {
            cudaResourceDesc rd;
            std::memset(&rd, 0, sizeof(rd));
            rd.resType = cudaResourceTypeArray;
            rd.res.array.array = v.mean_lay;
            cudaTextureDesc td;
            std::memset(&td, 0, sizeof(td));
            td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 0;
            VRDD_CUDA(c, cudaCreateTextureObject(&v.mean_gather, &rd, &td, nullptr));
        } else {
            cudaGetLastError();                              // no layered array: the 3-D array below takes over
            v.mean_lay = nullptr;
        }
    }
    if (c->keep_mean_raw && source == VRDD_SRC_ORIGINAL && !v.mean_raw && !v.mean_lay && !v.mean_tex) {
        // point-sampled, un-normalised coordinates: texel (x, y, z) is fetched at (x + .5, y + .5, z + .5)
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        if (!v.mean_arr) VRDD_CUDA(c, cudaMalloc3DArray(&v.mean_arr, &desc, make_cudaExtent(c->W, c->H, c->D), 0));
        cudaResourceDesc rd;
        std::memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = v.mean_arr;
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        VRDD_CUDA(c, cudaCreateTextureObject(&v.mean_tex, &rd, &td, nullptr));
    }
    if (c->keep_mean_raw && source == VRDD_SRC_ORIGINAL && !v.mean_raw)      // last: its presence marks the set-up as done
        VRDD_CUDA(c, cudaMalloc(&v.mean_raw, sizeof(float) * c->V));
    return VRDD_OK;
}

DecodeOut make_decode_out(vrdd_context* c, int source, long long v_base) {
    vrdd_decoded_volume& v = c->vol[source];
    DecodeOut o;
    for (int i = 0; i < 3; ++i) {
        o.lin[i] = v.lin[i];
        o.surf[i] = v.surf[i];
        o.brick[i] = v.brick[i];
    }
    o.mean_raw = (source == VRDD_SRC_ORIGINAL) ? v.mean_raw : nullptr;
    o.use_surf = v.surf[0] != 0;
    o.W = c->W; o.H = c->H; o.D = c->D;
    o.v_base = v_base;
    o.bW = c->bW; o.bH = c->bH;
    o.inv_wh = 1.0f / ((float)c->W * (float)c->H);
    o.n_peers = c->n_peers[source];
    for (int q = 0; q < VRDD_MAX_PEERS; ++q) {
        for (int i = 0; i < 3; ++i) o.peer[q][i] = (q < o.n_peers) ? c->peer_planes[source][q][i] : nullptr;
        o.peer_mean_raw[q] = (q < o.n_peers && source == VRDD_SRC_ORIGINAL) ? c->peer_mean_raw[q] : nullptr;
    }
    return o;
}

}  // namespace vrdd

// linear plane slab -> 3-D array
__global__ void commit_array_kernel(const float* __restrict__ lin, cudaSurfaceObject_t surf, int W, int H, int z0) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = z0 + blockIdx.z;
    if (x < W) surf3Dwrite(lin[((size_t)z * H + y) * W + x], surf, x * 4, y, z);
}

// linear plane slab -> sampler layout
__global__ void commit_brick_kernel(const float* __restrict__ lin, float* __restrict__ brick, int W, int H, int bW,
                                    int bH, long long v0, long long nvox) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvox) return;
    const long long gv = v0 + i, wh = (long long)W * H;
    const int z = (int)(gv / wh);
    const int r = (int)(gv - (long long)z * wh);
    const int y = r / W, x = r - y * W;
    brick[vrdd::brick_index(x, y, z, bW, bH)] = lin[gv];
}

__global__ void gather_brick_kernel(const float* __restrict__ brick, float* __restrict__ lin, int W, int H, int bW,
                                    int bH, long long nvox) {
    const long long gv = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gv >= nvox) return;
    const long long wh = (long long)W * H;
    const int z = (int)(gv / wh);
    const int r = (int)(gv - (long long)z * wh);
    const int y = r / W, x = r - y * W;
    lin[gv] = brick[vrdd::brick_index(x, y, z, bW, bH)];
}

using namespace vrdd;

#define CHECK_HANDLE(h)                      \
    if (!(h)) return VRDD_ERR_INVALID;       \
    vrdd_context* c = (h);                   \
    DeviceGuard guard__(c->device)

extern "C" {

int vrdd_create(int device, vrdd_handle* out) {
    if (!out) return VRDD_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return VRDD_ERR_NO_DEVICE;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return VRDD_ERR_NO_DEVICE;
    if (device >= n) return VRDD_ERR_INVALID;
    vrdd_context* c = new vrdd_context();
    c->device = device;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return VRDD_ERR_NO_DEVICE; }
    c->num_sms = prop.multiProcessorCount;
    if (prop.major != 10 || prop.minor != 0) {       // arch-specific ("a") code runs on exactly this compute capability
        std::fprintf(stderr, "libvrdd: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
                     prop.major, prop.minor);
        delete c;
        return VRDD_ERR_NO_DEVICE;
    }
    if (cudaMalloc(&c->d_samples, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(c->d_samples, 0, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&c->d_tickets, 4 * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(c->d_tickets, 0, 4 * sizeof(unsigned)) != cudaSuccess) {
        delete c;
        return VRDD_ERR_CUDA;
    }
    *out = c;
    int rc = vrdd_set_transfer_function(c, nullptr, 0);
    if (rc != VRDD_OK) { vrdd_destroy(c); *out = nullptr; return rc; }
    return VRDD_OK;
}

int vrdd_destroy(vrdd_handle h) {
    CHECK_HANDLE(h);
    cudaStreamSynchronize(c->stream);
    free_volume(c->vol[0]);
    free_volume(c->vol[1]);
    free_inputs(c);
    free_tf(c);
    if (c->d_samples) cudaFree(c->d_samples);
    if (c->d_tickets) cudaFree(c->d_tickets);
    if (c->d_first4) cudaFree(c->d_first4);
    if (c->frame) cudaFree(c->frame);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (int i = 0; i < 2; ++i) {
            if (c->pframe[i]) cudaFree(c->pframe[i]);
            if (c->ev_rendered[i]) cudaEventDestroy(c->ev_rendered[i]);
            if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        }
        cudaStreamDestroy(c->copy_stream);
    }
    destroy_flex(c);
    delete c;
    return VRDD_OK;
}

int vrdd_set_stream(vrdd_handle h, void* cuda_stream) {
    CHECK_HANDLE(h);
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    return VRDD_OK;
}

int vrdd_synchronize(vrdd_handle h) {
    CHECK_HANDLE(h);
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
    return VRDD_OK;
}

const char* vrdd_last_error(vrdd_handle h) { return h ? h->err.c_str() : "null handle"; }
int64_t vrdd_kernel_launches(vrdd_handle h) { return h ? h->launches : 0; }

int vrdd_set_volume(vrdd_handle h, int width, int height, int depth, int bins) {
    CHECK_HANDLE(h);
    c->first4_tag = 0;                                  // sort-last: the fused record of pass 1 is stale
    if (width <= 0 || height <= 0 || depth <= 0) return fail(c, VRDD_ERR_INVALID, "set_volume: bad size");
    if (bins != VRDD_BINS) return fail(c, VRDD_ERR_UNSUPPORTED, "set_volume: this build supports bins == 32 only");
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
    free_volume(c->vol[0]);
    free_volume(c->vol[1]);
    free_inputs(c);
    c->W = width; c->H = height; c->D = depth; c->B = bins;
    c->V = (size_t)width * height * depth;
    c->bW = (width + VRDD_BRICK - 1) / VRDD_BRICK;
    c->bH = (height + VRDD_BRICK - 1) / VRDD_BRICK;
    c->bD = (depth + VRDD_BRICK - 1) / VRDD_BRICK;
    return VRDD_OK;
}

int vrdd_set_histograms_host(vrdd_handle h, const float* hist) {
    CHECK_HANDLE(h);
    if (!c->V || !hist) return fail(c, VRDD_ERR_INVALID, "set_histograms_host: set_volume first");
    const size_t bytes = c->V * c->B * sizeof(float);
    if (!c->hist_owned) VRDD_CUDA(c, cudaMalloc(&c->hist_owned, bytes));
    VRDD_CUDA(c, cudaMemcpyAsync(c->hist_owned, hist, bytes, cudaMemcpyHostToDevice, c->stream));
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));      // the caller frees right after (volumeRender.cpp:1205)
    c->hist = c->hist_owned; c->hist_z0 = 0; c->hist_nz = c->D;
    return VRDD_OK;
}

int vrdd_set_histograms_device(vrdd_handle h, const float* d_hist, int z0, int nz) {
    CHECK_HANDLE(h);
    if (!c->V || !d_hist || z0 < 0 || nz <= 0 || z0 + nz > c->D)
        return fail(c, VRDD_ERR_INVALID, "set_histograms_device: bad slab");
    if ((reinterpret_cast<uintptr_t>(d_hist) & 15u) != 0)
        return fail(c, VRDD_ERR_INVALID, "set_histograms_device: pointer must be 16-byte aligned");
    c->hist = d_hist; c->hist_z0 = z0; c->hist_nz = nz;
    return VRDD_OK;
}

// The compact error form (include/vrdd.h): per chunk of 32 consecutive voxels, round k holds the k-th error
// of every voxel of the chunk with NE > k, in voxel order; rounds follow one another.  Returns 0, or 1 / 2
// when a codebook entry / an error bin is out of range (the reference's guards, volumeRender_kernel.cu:781-816).
static int pack_errors(const int32_t* codebook, const float* errors_dense, size_t V, int B, int num_templates,
                       std::vector<vrdd_error_entry>* entries, std::vector<uint64_t>* offsets) {
    const size_t nchunks = (V + VRDD_ERR_CHUNK - 1) / VRDD_ERR_CHUNK;
    offsets->assign(nchunks + 1, 0);
    uint64_t total = 0;
    for (size_t v = 0; v < V; ++v) {
        if (v % VRDD_ERR_CHUNK == 0) (*offsets)[v / VRDD_ERR_CHUNK] = total;
        const int32_t* e = codebook + 4 * v;
        if ((num_templates > 0 && (e[0] < 0 || e[0] >= num_templates)) || e[1] < 0 || e[1] > B || e[3] < 0 || e[3] > B)
            return 1;
        total += (uint64_t)e[3];
    }
    (*offsets)[nchunks] = total;
    entries->resize(total);
    uint64_t w = 0;
    for (size_t c0 = 0; c0 < V; c0 += VRDD_ERR_CHUNK) {
        const size_t c1 = (c0 + VRDD_ERR_CHUNK < V) ? c0 + VRDD_ERR_CHUNK : V;
        for (int k = 0; k < B; ++k) {
            bool any = false;
            for (size_t v = c0; v < c1; ++v) {
                if (codebook[4 * v + 3] <= k) continue;
                any = true;
                const float* row = errors_dense + 2 * (v * B + k);
                const int bin = (int)row[0];
                if (bin < 0 || bin >= B) return 2;
                (*entries)[w].bin = bin; (*entries)[w].value = row[1];
                ++w;
            }
            if (!any) break;
        }
    }
    return 0;
}

int vrdd_pack_fractal_errors(const int32_t* codebook, const float* errors_dense, int64_t nvox, int bins,
                             vrdd_error_entry* entries, uint64_t* chunk_offsets, uint64_t* total_ne) {
    if (!codebook || !errors_dense || nvox <= 0 || bins <= 0 || !chunk_offsets) return VRDD_ERR_INVALID;
    std::vector<vrdd_error_entry> e;
    std::vector<uint64_t> off;
    if (pack_errors(codebook, errors_dense, (size_t)nvox, bins, 0, &e, &off) != 0) return VRDD_ERR_RANGE;
    std::memcpy(chunk_offsets, off.data(), sizeof(uint64_t) * off.size());
    if (entries && !e.empty()) std::memcpy(entries, e.data(), sizeof(vrdd_error_entry) * e.size());
    if (total_ne) *total_ne = e.size();
    return VRDD_OK;
}

int vrdd_set_fractal_host(vrdd_handle h, const int32_t* codebook, const float* errors_dense, const float* templates,
                          int num_templates) {
    CHECK_HANDLE(h);
    if (!c->V || !codebook || !errors_dense || !templates || num_templates <= 0)
        return fail(c, VRDD_ERR_INVALID, "set_fractal_host: bad arguments");
    const int B = c->B;
    const size_t V = c->V;
    const size_t nchunks = (V + VRDD_ERR_CHUNK - 1) / VRDD_ERR_CHUNK;
    std::vector<vrdd_error_entry> compact;
    std::vector<uint64_t> off;
    const int bad = pack_errors(codebook, errors_dense, V, B, num_templates, &compact, &off);
    if (bad == 1) return fail(c, VRDD_ERR_RANGE, "set_fractal_host: codebook entry out of range");
    if (bad == 2) return fail(c, VRDD_ERR_RANGE, "set_fractal_host: error bin out of range");
    if (compact.empty()) compact.resize(1);
    for (int i = 0; i < num_templates * B; ++i)
        if (!(templates[i] >= 0.0f && templates[i] <= 1.0f))
            return fail(c, VRDD_ERR_RANGE, "set_fractal_host: template frequency outside [0,1]");
    if (c->cb_owned) cudaFree(c->cb_owned);
    if (c->err_owned) cudaFree(c->err_owned);
    if (c->off_owned) cudaFree(c->off_owned);
    if (c->tmpl_owned) cudaFree(c->tmpl_owned);
    c->cb_owned = nullptr; c->err_owned = nullptr; c->off_owned = nullptr; c->tmpl_owned = nullptr;
    // nothing may keep pointing at the freed buffers if one of the calls below fails
    c->cb = nullptr; c->errs = nullptr; c->err_off = nullptr; c->tmpl = nullptr; c->fr_nz = 0; c->num_templates = 0;
    VRDD_CUDA(c, cudaMalloc(&c->cb_owned, sizeof(int32_t) * 4 * V));
    VRDD_CUDA(c, cudaMalloc(&c->err_owned, sizeof(vrdd_error_entry) * compact.size()));
    VRDD_CUDA(c, cudaMalloc(&c->off_owned, sizeof(uint64_t) * (nchunks + 1)));
    VRDD_CUDA(c, cudaMalloc(&c->tmpl_owned, sizeof(float) * num_templates * B));
    VRDD_CUDA(c, cudaMemcpyAsync(c->cb_owned, codebook, sizeof(int32_t) * 4 * V, cudaMemcpyHostToDevice, c->stream));
    VRDD_CUDA(c, cudaMemcpyAsync(c->err_owned, compact.data(), sizeof(vrdd_error_entry) * compact.size(),
                                 cudaMemcpyHostToDevice, c->stream));
    VRDD_CUDA(c, cudaMemcpyAsync(c->off_owned, off.data(), sizeof(uint64_t) * (nchunks + 1), cudaMemcpyHostToDevice,
                                 c->stream));
    VRDD_CUDA(c, cudaMemcpyAsync(c->tmpl_owned, templates, sizeof(float) * num_templates * B, cudaMemcpyHostToDevice,
                                 c->stream));
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->cb = c->cb_owned; c->errs = c->err_owned; c->err_off = c->off_owned; c->tmpl = c->tmpl_owned;
    c->num_templates = num_templates; c->fr_z0 = 0; c->fr_nz = c->D;
    return build_template_moments(c, c->tmpl, num_templates);
}

int vrdd_set_fractal_device(vrdd_handle h, const int32_t* d_codebook, const vrdd_error_entry* d_errors,
                            const uint64_t* d_chunk_offsets, const float* d_templates, int num_templates, int z0,
                            int nz) {
    CHECK_HANDLE(h);
    if (!c->V || !d_codebook || !d_errors || !d_chunk_offsets || !d_templates || num_templates <= 0 || z0 < 0 ||
        nz <= 0 || z0 + nz > c->D)
        return fail(c, VRDD_ERR_INVALID, "set_fractal_device: bad arguments");
    if (((reinterpret_cast<uintptr_t>(d_codebook) | reinterpret_cast<uintptr_t>(d_templates)) & 15u) != 0 ||
        (reinterpret_cast<uintptr_t>(d_errors) & 7u) != 0)
        return fail(c, VRDD_ERR_INVALID, "set_fractal_device: misaligned pointer");
    c->cb = d_codebook; c->errs = d_errors; c->err_off = d_chunk_offsets; c->tmpl = d_templates;
    c->num_templates = num_templates; c->fr_z0 = z0; c->fr_nz = nz;
    // The per-(template, flip, shift) moment table is rebuilt on EVERY call: it is keyed on the table's contents, and a
    // pointer tells nothing about those (a caller may rewrite the templates in place, or an allocator may hand the
    // same address out again).  T x 64 threads: microseconds next to a slab decode.
    return build_template_moments(c, d_templates, num_templates);
}

int vrdd_set_sampler(vrdd_handle h, int sampler) {
    CHECK_HANDLE(h);
    c->first4_tag = 0;                                  // sort-last: the fused record of pass 1 is stale
    if (sampler != VRDD_SAMPLER_TEXTURE && sampler != VRDD_SAMPLER_BRICKED && sampler != VRDD_SAMPLER_LINEAR)
        return fail(c, VRDD_ERR_INVALID, "set_sampler: unknown sampler");
    c->sampler = sampler;
    return VRDD_OK;
}

int vrdd_enable_interpolated_mean(vrdd_handle h, int enable) {
    CHECK_HANDLE(h);
    c->keep_mean_raw = enable != 0;
    return VRDD_OK;
}

int vrdd_keep_linear_planes(vrdd_handle h, int keep) {
    CHECK_HANDLE(h);
    c->keep_linear = keep != 0;
    return VRDD_OK;
}

// the slab [z0, z0 + nz) of un-normalised block means, linear -> the array queryMethod 7 fetches from: slices of the 3-D
// array, or the same layers of the layered one (only when queryMethod 7 was asked for)
static int commit_mean_raw(vrdd_context* c, int z0, int nz) {
    vrdd_decoded_volume& v = c->vol[VRDD_SRC_ORIGINAL];
    if (!v.mean_raw || !(v.mean_arr || v.mean_lay)) return VRDD_OK;
    const size_t slice = (size_t)c->W * c->H;
    cudaMemcpy3DParms cp;
    std::memset(&cp, 0, sizeof(cp));
    cp.srcPtr = make_cudaPitchedPtr(v.mean_raw + (size_t)z0 * slice, sizeof(float) * c->W, c->W, c->H);
    cp.dstArray = v.mean_lay ? v.mean_lay : v.mean_arr;
    cp.dstPos = make_cudaPos(0, 0, z0);
    cp.extent = make_cudaExtent(c->W, c->H, nz);
    cp.kind = cudaMemcpyDeviceToDevice;
    VRDD_CUDA(c, cudaMemcpy3DAsync(&cp, c->stream));
    return VRDD_OK;
}

int vrdd_get_mean_raw_device(vrdd_handle h, float** d_mean_raw) {
    CHECK_HANDLE(h);
    if (!d_mean_raw) return fail(c, VRDD_ERR_INVALID, "get_mean_raw_device: null");
    if (c->keep_mean_raw && c->V) {
        int rc = ensure_volume_storage(c, VRDD_SRC_ORIGINAL);
        if (rc != VRDD_OK) return rc;
    }
    *d_mean_raw = c->vol[VRDD_SRC_ORIGINAL].mean_raw;
    return VRDD_OK;
}

int vrdd_set_peer_mean_raw(vrdd_handle h, int n_peers, float* const* d_mean_raw) {
    CHECK_HANDLE(h);
    if (n_peers < 0 || n_peers > VRDD_MAX_PEERS || (n_peers > 0 && !d_mean_raw))
        return fail(c, VRDD_ERR_INVALID, "set_peer_mean_raw: bad arguments");
    for (int q = 0; q < VRDD_MAX_PEERS; ++q) c->peer_mean_raw[q] = (q < n_peers) ? d_mean_raw[q] : nullptr;
    if (n_peers > c->n_peers[VRDD_SRC_ORIGINAL]) c->n_peers[VRDD_SRC_ORIGINAL] = n_peers;
    return VRDD_OK;
}

int vrdd_commit_mean_raw(vrdd_handle h, int z0, int nz) {
    CHECK_HANDLE(h);
    if (z0 < 0 || nz <= 0 || z0 + nz > c->D) return fail(c, VRDD_ERR_INVALID, "commit_mean_raw: bad arguments");
    return commit_mean_raw(c, z0, nz);
}

int vrdd_decode(vrdd_handle h, int source, int z0, int nz) {
    CHECK_HANDLE(h);
    c->first4_tag = 0;                                  // sort-last: the fused record of pass 1 is stale
    if (source != VRDD_SRC_ORIGINAL && source != VRDD_SRC_FRACTAL) return fail(c, VRDD_ERR_INVALID, "decode: bad source");
    if (!c->V) return fail(c, VRDD_ERR_INVALID, "decode: set_volume first");
    const bool orig = source == VRDD_SRC_ORIGINAL;
    const int az0 = orig ? c->hist_z0 : c->fr_z0, anz = orig ? c->hist_nz : c->fr_nz;
    if (anz <= 0) return fail(c, VRDD_ERR_INVALID, "decode: no input attached for this source");
    if (nz <= 0) { z0 = az0; nz = anz; }
    if (z0 < az0 || z0 + nz > az0 + anz) return fail(c, VRDD_ERR_INVALID, "decode: slab outside the attached input");
    int rc = ensure_volume_storage(c, source);
    if (rc != VRDD_OK) return rc;
    const long long slice = (long long)c->W * c->H;
    const long long local0 = (long long)(z0 - az0) * slice;     // first voxel inside the attached slab
    const long long nvox = (long long)nz * slice;
    DecodeOut out = make_decode_out(c, source, (long long)z0 * slice);
    if (orig) {
        rc = launch_decode_hist(c, c->hist + local0 * c->B, nvox, out);
    } else {
        // the error table is grouped per chunk of 32 voxels (round-major inside a chunk), so a sub-slab has to be
        // made of whole chunks; only the very end of the attached input may fall inside one
        if (local0 % VRDD_ERR_CHUNK != 0 || ((local0 + nvox) % VRDD_ERR_CHUNK != 0 && z0 + nz != az0 + anz))
            return fail(c, VRDD_ERR_INVALID, "decode: fractal sub-slab must start and end on a 32-voxel boundary");
        rc = launch_decode_fractal(c, c->cb + 4 * local0, c->errs, c->err_off + local0 / VRDD_ERR_CHUNK, c->tmpl,
                                   c->num_templates, nvox, out, nullptr);
    }
    if (rc == VRDD_OK && orig) rc = commit_mean_raw(c, z0, nz);
    if (rc == VRDD_OK) c->vol[source].decoded = true;
    invalidate_gather_copies(c->vol[source]);
    return rc;
}

int vrdd_reconstruct_fractal_device(vrdd_handle h, float* d_out) {
    CHECK_HANDLE(h);
    if (c->fr_nz <= 0 || !d_out) return fail(c, VRDD_ERR_INVALID, "reconstruct_fractal: no fractal input attached");
    int rc = ensure_volume_storage(c, VRDD_SRC_FRACTAL);
    if (rc != VRDD_OK) return rc;
    const long long slice = (long long)c->W * c->H;
    DecodeOut out = make_decode_out(c, VRDD_SRC_FRACTAL, (long long)c->fr_z0 * slice);
    rc = launch_decode_fractal(c, c->cb, c->errs, c->err_off, c->tmpl, c->num_templates, (long long)c->fr_nz * slice,
                               out, d_out);
    if (rc == VRDD_OK) c->vol[VRDD_SRC_FRACTAL].decoded = true;
    invalidate_gather_copies(c->vol[VRDD_SRC_FRACTAL]);
    return rc;
}

int vrdd_get_decoded_planes_device(vrdd_handle h, int source, float** mean, float** variance, float** entropy) {
    CHECK_HANDLE(h);
    if (source < 0 || source > 1) return fail(c, VRDD_ERR_INVALID, "get_decoded_planes_device: bad source");
    if (c->keep_linear && c->V) {
        int rc = ensure_volume_storage(c, source);
        if (rc != VRDD_OK) return rc;
    }
    if (mean) *mean = c->vol[source].lin[0];
    if (variance) *variance = c->vol[source].lin[1];
    if (entropy) *entropy = c->vol[source].lin[2];
    return VRDD_OK;
}

int vrdd_commit_planes(vrdd_handle h, int source, int z0, int nz) { return vrdd_commit_planes_mask(h, source, z0, nz, 7); }

int vrdd_commit_planes_mask(vrdd_handle h, int source, int z0, int nz, int plane_mask) {
    CHECK_HANDLE(h);
    c->first4_tag = 0;                                  // sort-last: the fused record of pass 1 is stale
    if (source < 0 || source > 1 || z0 < 0 || nz <= 0 || z0 + nz > c->D || (plane_mask & ~7) || !plane_mask)
        return fail(c, VRDD_ERR_INVALID, "commit_planes: bad arguments");
    vrdd_decoded_volume& v = c->vol[source];
    if (!v.lin[0]) return fail(c, VRDD_ERR_INVALID, "commit_planes: linear planes are not kept");
    int rc = ensure_volume_storage(c, source);
    if (rc != VRDD_OK) return rc;
    const long long slice = (long long)c->W * c->H;
    for (int i = 0; i < 3; ++i) {
        if (!(plane_mask & (1 << i))) continue;
        if (v.arr[i] && v.surf[i]) {
            // linear slab -> 3-D array through the surface: one thread per voxel, coalesced reads, x-fastest writes
            const dim3 grid((c->W + 255) / 256, c->H, nz);
            commit_array_kernel<<<grid, 256, 0, c->stream>>>(v.lin[i], v.surf[i], c->W, c->H, z0);
            c->launches += 1;
            VRDD_CUDA(c, cudaGetLastError());
        }
        if (v.brick[i]) {
            const long long nvox = (long long)nz * slice;
            commit_brick_kernel<<<(unsigned)((nvox + 255) / 256), 256, 0, c->stream>>>(
                v.lin[i], v.brick[i], c->W, c->H, c->bW, c->bH, (long long)z0 * slice, nvox);
            c->launches += 1;
            VRDD_CUDA(c, cudaGetLastError());
        }
    }
    v.decoded = true;
    invalidate_gather_copies(v);
    return VRDD_OK;
}

int vrdd_get_decoded_host(vrdd_handle h, int source, float* out4) {
    CHECK_HANDLE(h);
    if (source < 0 || source > 1 || !out4) return fail(c, VRDD_ERR_INVALID, "get_decoded_host: bad arguments");
    vrdd_decoded_volume& v = c->vol[source];
    if (!v.decoded) return fail(c, VRDD_ERR_INVALID, "get_decoded_host: nothing decoded");
    std::vector<float> plane(c->V);
    float* d_tmp = nullptr;
    for (int i = 0; i < 3; ++i) {
        if (v.lin[i]) {
            VRDD_CUDA(c, cudaMemcpyAsync(plane.data(), v.lin[i], sizeof(float) * c->V, cudaMemcpyDeviceToHost, c->stream));
        } else if (v.arr[i]) {
            cudaMemcpy3DParms p;
            std::memset(&p, 0, sizeof(p));
            p.srcArray = v.arr[i];
            p.dstPtr = make_cudaPitchedPtr(plane.data(), c->W * sizeof(float), c->W, c->H);
            p.extent = make_cudaExtent(c->W, c->H, c->D);
            p.kind = cudaMemcpyDeviceToHost;
            VRDD_CUDA(c, cudaMemcpy3DAsync(&p, c->stream));
        } else if (v.brick[i]) {
            if (!d_tmp) VRDD_CUDA(c, cudaMalloc(&d_tmp, sizeof(float) * c->V));
            gather_brick_kernel<<<(unsigned)((c->V + 255) / 256), 256, 0, c->stream>>>(v.brick[i], d_tmp, c->W, c->H,
                                                                                      c->bW, c->bH, (long long)c->V);
            c->launches += 1;
            VRDD_CUDA(c, cudaMemcpyAsync(plane.data(), d_tmp, sizeof(float) * c->V, cudaMemcpyDeviceToHost, c->stream));
        } else {
            return fail(c, VRDD_ERR_INVALID, "get_decoded_host: no storage");
        }
        VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
        for (size_t k = 0; k < c->V; ++k) out4[4 * k + i] = plane[k];
    }
    if (d_tmp) cudaFree(d_tmp);
    for (size_t k = 0; k < c->V; ++k) out4[4 * k + 3] = 0.0f;   // the reference never writes .w (:771-773)
    return VRDD_OK;
}

int vrdd_set_transfer_function(vrdd_handle h, const float* tf, int n) {
    CHECK_HANDLE(h);
    c->first4_tag = 0;                                  // sort-last: the fused record of pass 1 is stale
    if (!tf) { tf = &kDefaultTf[0][0]; n = 9; }
    if (n < 1 || n > VRDD_MAX_TF) return fail(c, VRDD_ERR_INVALID, "set_transfer_function: 1..1024 entries");
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
    free_tf(c);
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    VRDD_CUDA(c, cudaMallocArray(&c->tf_arr, &desc, n, 1));
    VRDD_CUDA(c, cudaMemcpy2DToArray(c->tf_arr, 0, 0, tf, n * sizeof(float4), n * sizeof(float4), 1,
                                     cudaMemcpyHostToDevice));
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->tf_arr;
    cudaTextureDesc td;                                   // volumeRender_kernel.cu:2337-2339
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    VRDD_CUDA(c, cudaCreateTextureObject(&c->tf_tex, &rd, &td, nullptr));
    VRDD_CUDA(c, cudaMalloc(&c->tf_dev, n * sizeof(float4)));
    VRDD_CUDA(c, cudaMemcpy(c->tf_dev, tf, n * sizeof(float4), cudaMemcpyHostToDevice));
    c->tf_n = n;
    return VRDD_OK;
}

int vrdd_set_view(vrdd_handle h, const float* m12) {
    if (!h || !m12) return VRDD_ERR_INVALID;
    std::memcpy(h->view, m12, sizeof(float) * 12);      // travels as a kernel parameter
    return VRDD_OK;
}

void vrdd_default_render_params(vrdd_render_params* p) {
    if (!p) return;
    p->density = 0.05f; p->brightness = 1.0f; p->transfer_offset = 0.0f; p->transfer_scale = 1.0f;
    p->tstep = 0.01f; p->max_steps = 500; p->opacity_threshold = 0.95f; p->query_method = 1;
}

int vrdd_render(vrdd_handle h, uint32_t* d_output, int image_w, int image_h, const vrdd_render_params* params,
                const vrdd_tile_partition* part, int clear_misses) {
    CHECK_HANDLE(h);
    vrdd_render_params p;
    if (params) p = *params; else vrdd_default_render_params(&p);
    vrdd_tile_partition tp;
    if (part) tp = *part; else { tp.tile_w = image_w; tp.tile_h = image_h; tp.part = 0; tp.parts = 1; }
    return launch_raycast(c, d_output, image_w, image_h, p, tp, clear_misses);
}

int vrdd_render_host(vrdd_handle h, uint32_t* h_output, int image_w, int image_h, const vrdd_render_params* params) {
    CHECK_HANDLE(h);
    if (!h_output || image_w <= 0 || image_h <= 0) return fail(c, VRDD_ERR_INVALID, "render_host: bad image");
    const size_t bytes = sizeof(uint32_t) * (size_t)image_w * image_h;
    if (c->frame_bytes < bytes) {                       // the frame buffer lives as long as the handle
        if (c->frame) cudaFree(c->frame);
        c->frame = nullptr; c->frame_bytes = 0;
        VRDD_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->frame), bytes));
        c->frame_bytes = bytes;
    }
    int rc = vrdd_render(h, c->frame, image_w, image_h, params, nullptr, 1);
    if (rc == VRDD_OK) {
        cudaError_t e = cudaMemcpyAsync(h_output, c->frame, bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) rc = fail_cuda(c, e, "render_host: read-back");
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (rc == VRDD_OK && e != cudaSuccess) rc = fail_cuda(c, e, "render_host: synchronize");
    return rc;
}

// Pipelined read-back: frame k is copied to the host by a second stream while frame k+1 renders.
// With a partition of full-width row bands (tile_w >= image_w) only this rank's bands are rendered and only
// they are copied, into their place in the caller's full-frame host buffer: N ranks fill one frame in shared
// host memory over N PCIe links.
int vrdd_render_host_async(vrdd_handle h, uint32_t* h_output, int image_w, int image_h,
                           const vrdd_render_params* params, const vrdd_tile_partition* part) {
    CHECK_HANDLE(h);
    if (!h_output || image_w <= 0 || image_h <= 0) return fail(c, VRDD_ERR_INVALID, "render_host_async: bad image");
    const bool banded = part && part->parts > 1;
    if (banded && (part->tile_w < image_w || part->tile_h < 1 || part->part < 0 || part->part >= part->parts))
        return fail(c, VRDD_ERR_UNSUPPORTED, "render_host_async: a partition must consist of full-width row bands");
    const size_t bytes = sizeof(uint32_t) * (size_t)image_w * image_h;
    if (!c->copy_stream) {
        VRDD_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            VRDD_CUDA(c, cudaEventCreateWithFlags(&c->ev_rendered[i], cudaEventDisableTiming));
            VRDD_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
        }
    }
    if (c->pframe_bytes < bytes) {
        VRDD_CUDA(c, cudaStreamSynchronize(c->copy_stream));
        for (int i = 0; i < 2; ++i) {
            if (c->pframe[i]) cudaFree(c->pframe[i]);
            c->pframe[i] = nullptr;
        }
        c->pframe_bytes = 0;
        for (int i = 0; i < 2; ++i) VRDD_CUDA(c, cudaMalloc(reinterpret_cast<void**>(&c->pframe[i]), bytes));
        c->pframe_bytes = bytes;
        c->pslot = 0;
    }
    const int slot = (int)(c->pslot & 1u);
    if (c->pslot >= 2) VRDD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));   // frame k-2 has left this slot
    const int rc = vrdd_render(h, c->pframe[slot], image_w, image_h, params, banded ? part : nullptr, 1);
    if (rc != VRDD_OK) return rc;
    VRDD_CUDA(c, cudaEventRecord(c->ev_rendered[slot], c->stream));
    VRDD_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rendered[slot], 0));
    if (!banded) {
        VRDD_CUDA(c, cudaMemcpyAsync(h_output, c->pframe[slot], bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    } else {
        // bands part, part + parts, ...: the complete ones are one strided copy, a ragged last band a second
        const size_t row = sizeof(uint32_t) * (size_t)image_w, band = row * part->tile_h;
        const int nbands = (image_h + part->tile_h - 1) / part->tile_h;
        const int nfull = image_h / part->tile_h;                      // bands [0, nfull) are complete
        const int mine_full = (nfull > part->part) ? (nfull - part->part + part->parts - 1) / part->parts : 0;
        const size_t first = band * part->part;
        if (mine_full > 0)
            VRDD_CUDA(c, cudaMemcpy2DAsync(reinterpret_cast<char*>(h_output) + first, band * part->parts,
                                           reinterpret_cast<const char*>(c->pframe[slot]) + first, band * part->parts, band,
                                           mine_full, cudaMemcpyDeviceToHost, c->copy_stream));
        if (nbands > nfull && (nfull % part->parts) == part->part) {
            const size_t off = band * nfull;
            VRDD_CUDA(c, cudaMemcpyAsync(reinterpret_cast<char*>(h_output) + off,
                                         reinterpret_cast<const char*>(c->pframe[slot]) + off, bytes - off,
                                         cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    VRDD_CUDA(c, cudaEventRecord(c->ev_copied[slot], c->copy_stream));
    c->pslot += 1;
    return VRDD_OK;
}

// The handle's stream waits (on the device, the host does not block) until the read-back of the frame queued
// `lag` calls ago has finished: lag 0 = the last frame, 1 = the one before (what a per-frame barrier between
// ranks should wait for, so that the last frame's copy still overlaps the next render).
int vrdd_render_host_fence(vrdd_handle h, int lag) {
    CHECK_HANDLE(h);
    if (lag < 0 || lag > 1) return fail(c, VRDD_ERR_INVALID, "render_host_fence: lag is 0 or 1");
    if (!c->copy_stream || c->pslot <= (unsigned)lag) return VRDD_OK;
    const int slot = (int)((c->pslot - 1u - (unsigned)lag) & 1u);
    VRDD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[slot], 0));
    return VRDD_OK;
}

int vrdd_render_host_wait(vrdd_handle h) {
    CHECK_HANDLE(h);
    if (c->copy_stream) VRDD_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
    return VRDD_OK;
}

// Page-locks caller memory (for instance a frame in POSIX shared memory that several ranks fill) so that
// copies into it are asynchronous and run at full PCIe rate.
int vrdd_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return VRDD_ERR_INVALID;
    return cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess ? VRDD_OK : VRDD_ERR_CUDA;
}

int vrdd_host_unregister(void* p) {
    if (!p) return VRDD_ERR_INVALID;
    return cudaHostUnregister(p) == cudaSuccess ? VRDD_OK : VRDD_ERR_CUDA;
}

int vrdd_count_samples(vrdd_handle h, int enable) {
    if (!h) return VRDD_ERR_INVALID;
    h->count_samples = enable != 0;
    return VRDD_OK;
}

int vrdd_get_sample_count(vrdd_handle h, int64_t* out, int reset) {
    CHECK_HANDLE(h);
    if (!out) return fail(c, VRDD_ERR_INVALID, "get_sample_count: null");
    unsigned long long v = 0;
    VRDD_CUDA(c, cudaMemcpyAsync(&v, c->d_samples, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
    if (reset) VRDD_CUDA(c, cudaMemsetAsync(c->d_samples, 0, sizeof(v), c->stream));
    VRDD_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = (int64_t)v;
    return VRDD_OK;
}

// volumeRender.cpp:224-246 without OpenGL.  glRotatef(-rot.x, 1,0,0); glRotatef(-rot.y, 0,1,0);
// glTranslatef(-trans): GL post-multiplies, so M = Rx * Ry * T and the stored 3x4 is the
// top three rows of M (rotation | rotation * (-trans)).
void vrdd_view_matrix(float rot_x_deg, float rot_y_deg, float tx, float ty, float tz, float* m12) {
    const double k = 3.14159265358979323846 / 180.0;
    const double ax = -(double)rot_x_deg * k, ay = -(double)rot_y_deg * k;
    const float cx = (float)std::cos(ax), sx = (float)std::sin(ax);
    const float cy = (float)std::cos(ay), sy = (float)std::sin(ay);
    const float R[3][3] = {{cy, 0.0f, sy}, {sx * sy, cx, -(sx * cy)}, {-(cx * sy), sx, cx * cy}};
    const float t[3] = {-tx, -ty, -tz};
    for (int r = 0; r < 3; ++r) {
        for (int q = 0; q < 3; ++q) m12[4 * r + q] = R[r][q];
        m12[4 * r + 3] = R[r][0] * t[0] + R[r][1] * t[1] + R[r][2] * t[2];
    }
}

int vrdd_synth_histograms_device(vrdd_handle h, uint32_t seed, int z0, int nz, float* d_hist) {
    CHECK_HANDLE(h);
    if (!c->V || !d_hist || z0 < 0 || nz <= 0 || z0 + nz > c->D)
        return fail(c, VRDD_ERR_INVALID, "synth_histograms_device: bad arguments");
    return launch_synth_hist(c, seed, z0, nz, d_hist);
}

int vrdd_synth_fractal_device(vrdd_handle h, uint32_t seed, int num_templates, int max_ne, int z0, int nz,
                              int32_t* d_codebook, vrdd_error_entry* d_errors, uint64_t* d_chunk_offsets,
                              float* d_templates, uint64_t* total_ne) {
    CHECK_HANDLE(h);
    if (!c->V || !d_codebook || !d_errors || !d_chunk_offsets || !d_templates || num_templates <= 0 || max_ne < 0 ||
        max_ne > VRDD_BINS || z0 < 0 || nz <= 0 || z0 + nz > c->D)
        return fail(c, VRDD_ERR_INVALID, "synth_fractal_device: bad arguments");
    return launch_synth_fractal(c, seed, num_templates, max_ne, z0, nz, d_codebook, d_errors, d_chunk_offsets,
                                d_templates, total_ne);
}

int vrdd_frame_alloc(vrdd_handle h, size_t bytes, void** d_frame) {
    CHECK_HANDLE(h);
    if (!d_frame || !bytes) return fail(c, VRDD_ERR_INVALID, "frame_alloc: bad arguments");
    VRDD_CUDA(c, cudaMalloc(d_frame, bytes));            // a plain cudaMalloc block: exportable as a whole
    VRDD_CUDA(c, cudaMemset(*d_frame, 0, bytes));
    return VRDD_OK;
}

int vrdd_frame_free(vrdd_handle h, void* d_frame) {
    CHECK_HANDLE(h);
    VRDD_CUDA(c, cudaFree(d_frame));
    return VRDD_OK;
}

int vrdd_frame_export(vrdd_handle h, const void* d_frame, unsigned char* ipc_handle) {
    CHECK_HANDLE(h);
    static_assert(sizeof(cudaIpcMemHandle_t) == VRDD_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d_frame || !ipc_handle) return fail(c, VRDD_ERR_INVALID, "frame_export: bad arguments");
    cudaIpcMemHandle_t hnd;
    VRDD_CUDA(c, cudaIpcGetMemHandle(&hnd, const_cast<void*>(d_frame)));
    std::memcpy(ipc_handle, &hnd, sizeof(hnd));
    return VRDD_OK;
}

int vrdd_frame_open(vrdd_handle h, const unsigned char* ipc_handle, void** d_peer_frame) {
    CHECK_HANDLE(h);
    if (!d_peer_frame || !ipc_handle) return fail(c, VRDD_ERR_INVALID, "frame_open: bad arguments");
    cudaIpcMemHandle_t hnd;
    std::memcpy(&hnd, ipc_handle, sizeof(hnd));
    VRDD_CUDA(c, cudaIpcOpenMemHandle(d_peer_frame, hnd, cudaIpcMemLazyEnablePeerAccess));
    return VRDD_OK;
}

int vrdd_frame_close(vrdd_handle h, void* d_peer_frame) {
    CHECK_HANDLE(h);
    VRDD_CUDA(c, cudaIpcCloseMemHandle(d_peer_frame));
    return VRDD_OK;
}

int vrdd_set_frame_signal(vrdd_handle h, uint32_t* d_flag) {
    CHECK_HANDLE(h);
    if (d_flag && (reinterpret_cast<uintptr_t>(d_flag) & 3u)) return fail(c, VRDD_ERR_INVALID, "set_frame_signal: misaligned flag");
    c->frame_signal = d_flag;
    return VRDD_OK;
}

int vrdd_stream_wait_flag(vrdd_handle h, const uint32_t* d_flag, uint32_t at_least) {
    CHECK_HANDLE(h);
    if (!d_flag) return fail(c, VRDD_ERR_INVALID, "stream_wait_flag: null flag");
    return launch_stream_wait_flag(c, d_flag, at_least, nullptr);
}

int vrdd_stream_wait_post_flag(vrdd_handle h, const uint32_t* d_wait, uint32_t at_least, uint32_t* d_post) {
    CHECK_HANDLE(h);
    if (!d_wait || !d_post) return fail(c, VRDD_ERR_INVALID, "stream_wait_post_flag: null flag");
    return launch_stream_wait_flag(c, d_wait, at_least, d_post);
}

int vrdd_stream_post_flag(vrdd_handle h, uint32_t* d_flag) {
    CHECK_HANDLE(h);
    if (!d_flag) return fail(c, VRDD_ERR_INVALID, "stream_post_flag: null flag");
    return launch_stream_post_flag(c, d_flag);
}

int vrdd_set_peer_planes(vrdd_handle h, int source, int n_peers, float* const* d_planes) {
    CHECK_HANDLE(h);
    if (source < 0 || source > 1 || n_peers < 0 || n_peers > VRDD_MAX_PEERS || (n_peers > 0 && !d_planes))
        return fail(c, VRDD_ERR_INVALID, "set_peer_planes: bad arguments");
    c->n_peers[source] = n_peers;
    for (int q = 0; q < VRDD_MAX_PEERS; ++q)
        for (int i = 0; i < 3; ++i) c->peer_planes[source][q][i] = (q < n_peers) ? d_planes[3 * q + i] : nullptr;
    return VRDD_OK;
}

int vrdd_synth_histograms_region_device(vrdd_handle h, uint32_t seed, int gw, int gh, int gd, int ox, int oy, int oz,
                                        int z0, int nz, float* d_hist) {
    CHECK_HANDLE(h);
    if (!c->V || !d_hist || z0 < 0 || nz <= 0 || z0 + nz > c->D || ox < 0 || oy < 0 || oz < 0 || ox + c->W > gw ||
        oy + c->H > gh || oz + c->D > gd)
        return fail(c, VRDD_ERR_INVALID, "synth_histograms_region_device: the brick must lie inside the global volume");
    return launch_synth_hist_region(c, seed, gw, gh, gd, ox, oy, oz, z0, nz, d_hist);
}

static int check_brick(vrdd_context* c, const vrdd_brick* b) {
    if (!b) return fail(c, VRDD_ERR_INVALID, "render_brick: null brick");
    if (b->ox < 0 || b->oy < 0 || b->oz < 0 || b->ox + c->W > b->gw || b->oy + c->H > b->gh || b->oz + c->D > b->gd)
        return fail(c, VRDD_ERR_INVALID, "render_brick: the brick must lie inside the global volume");
    return VRDD_OK;
}

int vrdd_render_brick_alpha(vrdd_handle h, float* d_alpha_seg, int image_w, int image_h, const vrdd_render_params* params,
                            const vrdd_brick* brick) {
    CHECK_HANDLE(h);
    int rc = check_brick(c, brick);
    if (rc != VRDD_OK) return rc;
    vrdd_render_params p;
    if (params) p = *params; else vrdd_default_render_params(&p);
    return launch_brick_pass(c, 1, nullptr, d_alpha_seg, image_w, image_h, p, *brick);
}

int vrdd_render_brick_color(vrdd_handle h, const float* d_alpha_in, float* d_partial4, int image_w, int image_h,
                            const vrdd_render_params* params, const vrdd_brick* brick) {
    CHECK_HANDLE(h);
    int rc = check_brick(c, brick);
    if (rc != VRDD_OK) return rc;
    vrdd_render_params p;
    if (params) p = *params; else vrdd_default_render_params(&p);
    return launch_brick_pass(c, 2, d_alpha_in, d_partial4, image_w, image_h, p, *brick);
}

int vrdd_render_brick_alpha_send(vrdd_handle h, float* const* d_seg_tables, uint32_t* const* d_flags, int n_tables, int brick_index,
                                 int row0, int rows, int image_w, int image_h, const vrdd_render_params* params,
                                 const vrdd_brick* brick) {
    CHECK_HANDLE(h);
    int rc = check_brick(c, brick);
    if (rc != VRDD_OK) return rc;
    if (!d_seg_tables || !d_flags || n_tables < 1 || n_tables > VRDD_MAX_PEERS + 1 || brick_index < 0 || rows < 1)
        return fail(c, VRDD_ERR_INVALID, "render_brick_alpha_send: bad arguments");
    vrdd_render_params p;
    if (params) p = *params; else vrdd_default_render_params(&p);
    BrickSend s;
    s.n_dst = n_tables; s.row0 = row0; s.rows = rows;
    for (int d = 0; d < n_tables; ++d) {
        if (!d_seg_tables[d]) return fail(c, VRDD_ERR_INVALID, "render_brick_alpha_send: null table");
        s.dst[d] = d_seg_tables[d] + (size_t)brick_index * rows * image_w;           // slot [brick]: float[rows][W]
        s.flags[d] = d_flags[d];
    }
    return launch_brick_pass(c, 1, nullptr, nullptr, image_w, image_h, p, *brick, &s);
}

int vrdd_render_brick_color_send(vrdd_handle h, const float* d_alpha_in, float* d_root_slots4, uint32_t* d_root_flag, int brick_index,
                                 int row0, int rows, int image_w, int image_h, const vrdd_render_params* params,
                                 const vrdd_brick* brick) {
    CHECK_HANDLE(h);
    int rc = check_brick(c, brick);
    if (rc != VRDD_OK) return rc;
    if (!d_root_slots4 || brick_index < 0 || rows < 1) return fail(c, VRDD_ERR_INVALID, "render_brick_color_send: bad arguments");
    vrdd_render_params p;
    if (params) p = *params; else vrdd_default_render_params(&p);
    BrickSend s;
    s.n_dst = 1; s.row0 = row0; s.rows = rows;
    s.dst[0] = d_root_slots4 + (size_t)brick_index * rows * image_w * 4;              // slot [brick]: float4[rows][W]
    s.flags[0] = d_root_flag;
    return launch_brick_pass(c, 2, d_alpha_in, nullptr, image_w, image_h, p, *brick, &s);
}

int vrdd_render_brick_color_send_bands(vrdd_handle h, const float* d_alpha_in, float* const* d_owner_slots4, uint32_t* const* d_owner_flags,
                                       int n_owners, int band_rows, int brick_index, int row0, int rows, int image_w, int image_h,
                                       const vrdd_render_params* params, const vrdd_brick* brick) {
    CHECK_HANDLE(h);
    int rc = check_brick(c, brick);
    if (rc != VRDD_OK) return rc;
    if (!d_owner_slots4 || !d_owner_flags || n_owners < 1 || n_owners > VRDD_MAX_PEERS + 1 || band_rows < 1 || brick_index < 0 || rows < 1)
        return fail(c, VRDD_ERR_INVALID, "render_brick_color_send_bands: bad arguments");
    vrdd_render_params p;
    if (params) p = *params; else vrdd_default_render_params(&p);
    BrickSend s;
    s.n_dst = n_owners; s.row0 = row0; s.rows = rows; s.band_rows = band_rows;
    for (int o = 0; o < n_owners; ++o) {
        if (!d_owner_slots4[o]) return fail(c, VRDD_ERR_INVALID, "render_brick_color_send_bands: null table");
        s.dst[o] = d_owner_slots4[o] + (size_t)brick_index * band_rows * image_w * 4;   // slot [brick]: float4[band_rows][W]
        s.flags[o] = d_owner_flags[o];
    }
    return launch_brick_pass(c, 2, d_alpha_in, nullptr, image_w, image_h, p, *brick, &s);
}

int vrdd_pack_band_slots(vrdd_handle h, const float* d_slots4, int nbricks, const int* row0, int rows, int band_index, int band_rows,
                         uint32_t* d_frame, uint32_t* d_frame_flag, int image_w, int image_h, float brightness) {
    CHECK_HANDLE(h);
    return launch_pack_band_slots(c, d_slots4, nbricks, row0, rows, band_index, band_rows, d_frame, d_frame_flag, image_w, image_h, brightness);
}

int vrdd_pack_frame_slots(vrdd_handle h, const float* d_slots4, int nbricks, const int* row0, int rows, uint32_t* d_output,
                          int image_w, int image_h, float brightness) {
    CHECK_HANDLE(h);
    return launch_pack_frame_slots(c, d_slots4, nbricks, row0, rows, d_output, image_w, image_h, brightness);
}

int vrdd_compose_alpha_in(vrdd_handle h, const float* d_alpha_seg_all, int gx, int gy, int gz, int qx, int qy, int qz,
                          float* d_alpha_in, int image_w, int image_h) {
    CHECK_HANDLE(h);
    return launch_compose_alpha_in(c, d_alpha_seg_all, gx, gy, gz, qx, qy, qz, nullptr, image_h, d_alpha_in, image_w, image_h);
}

int vrdd_compose_alpha_in_rows(vrdd_handle h, const float* d_alpha_seg_rows, int gx, int gy, int gz, int qx, int qy, int qz,
                               const int* row0, int rows, float* d_alpha_in, int image_w, int image_h) {
    CHECK_HANDLE(h);
    if (!row0) return fail(c, VRDD_ERR_INVALID, "compose_alpha_in_rows: no row table");
    return launch_compose_alpha_in(c, d_alpha_seg_rows, gx, gy, gz, qx, qy, qz, row0, rows, d_alpha_in, image_w, image_h);
}

int vrdd_pack_frame(vrdd_handle h, const float* d_sum4, uint32_t* d_output, int image_w, int image_h, float brightness) {
    CHECK_HANDLE(h);
    if (!d_sum4 || !d_output || image_w <= 0 || image_h <= 0) return fail(c, VRDD_ERR_INVALID, "pack_frame: bad arguments");
    return launch_pack_frame(c, d_sum4, d_output, image_w, image_h, brightness);
}

int vrdd_set_variant(vrdd_handle h, const char* what, const char* variant) {
    if (!h || !what || !variant) return VRDD_ERR_INVALID;
    vrdd_context* c = h;
    const std::string w(what), v(variant);
    c->first4_tag = 0;                                  // sort-last: the fused record of pass 1 is stale
    if (w == "decode_hist") {
        if (v == "tma") c->var_decode_hist = 0;
        else if (v == "ldg") c->var_decode_hist = 1;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: decode_hist is tma|ldg");
    } else if (w == "decode_fractal") {
        if (v == "dense") c->var_fractal = 0;
        else if (v == "moments") c->var_fractal = 1;
        else if (v == "moments_global") c->var_fractal = 2;
        else if (v == "moments768") c->var_fractal = 3;
        else if (v == "moments2") c->var_fractal = 4;          // scan + predicated rows
        else if (v == "moments2r") c->var_fractal = 5;         // ... + float rows, g(old) recomputed
        else if (v == "moments2b") c->var_fractal = 6;         // ballots + predicated rows
        else if (v == "moments2br") c->var_fractal = 7;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: decode_fractal is dense|moments|moments_global");
    } else if (w == "decode_fractal_prefetch") {
        const int n = std::atoi(variant);
        if (n < 0 || n > 32) return fail(c, VRDD_ERR_INVALID, "set_variant: decode_fractal_prefetch is 0..32 lines");
        c->var_fractal_pf = n;
    } else if (w == "decode_fractal_sink") {
        if (v == "generic") c->var_fractal_sink = 0;
        else if (v == "auto") c->var_fractal_sink = 1;          // the surfaces-only instance where the sink allows it
        else return fail(c, VRDD_ERR_INVALID, "set_variant: decode_fractal_sink is auto|generic");
    } else if (w == "sortlast_blocks_per_sm") {
        const int n = std::atoi(variant);
        if (n < 0 || n > 16) return fail(c, VRDD_ERR_INVALID, "set_variant: sortlast_blocks_per_sm is 0..16");
        c->var_sortlast_blocks_per_sm = n;
    } else if (w == "sortlast_fuse") {
        if (v == "on") c->var_sortlast_fuse = 1;
        else if (v == "off") c->var_sortlast_fuse = 0;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: sortlast_fuse is on|off");
    } else if (w == "ray_setup") {
        if (v == "source") c->var_ray_setup = 0;
        else if (v == "nvcc") c->var_ray_setup = 1;            // queryMethod 1..7 of vrdd_render; see raycast.cu, ray_dir_nvcc
        else return fail(c, VRDD_ERR_INVALID, "set_variant: ray_setup is source|nvcc");
    } else if (w == "raycast_mode7") {
        if (v == "texture") c->var_mode7 = 0;
        else if (v == "linear") c->var_mode7 = 1;
        else if (v == "gather") c->var_mode7 = 2;             // set before the decode (the layered copy is made there)
        else return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_mode7 is texture|linear|gather");
    } else if (w == "decode_order") {
        if (v == "interleaved") c->var_decode_order = 0;
        else if (v == "chunked") c->var_decode_order = 1;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: decode_order is interleaved|chunked");
    } else if (w == "raycast_array_blocks_per_sm") {
        const int n = std::atoi(variant);
        if (n < -1 || n > 8) return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_array_blocks_per_sm is -1 (auto), 0 (all that fit), 1..8");
        c->var_array_blocks_per_sm = n;
    } else if (w == "raycast_unroll") {
        if (v == "1" || v == "2" || v == "4" || v == "8") c->var_unroll = v[0] - '0';
        else return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_unroll is 1|2|4|8");
    } else if (w == "raycast_layout") {
        if (v == "auto") c->var_layout = 0;
        else if (v == "array") c->var_layout = 1;
        else if (v == "layers_x") c->var_layout = 2;
        else if (v == "layers_y") c->var_layout = 3;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_layout is auto|array|layers_x|layers_y");
    } else if (w == "raycast_layout_min_step") {
        const float f = (float)std::atof(variant);
        if (!(f >= 0.f)) return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_layout_min_step is a number of voxels >= 0");
        c->var_layout_min_step = f;
    } else if (w == "raycast_layout_min_spacing") {
        const float f = (float)std::atof(variant);
        if (!(f >= 0.f)) return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_layout_min_spacing is a number of voxels >= 0");
        c->var_layout_min_spacing = f;
    } else if (w == "raycast_gather_unroll") {
        if (v == "2" || v == "4") c->var_gather_unroll = v[0] - '0';
        else return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_gather_unroll is 2|4");
    } else if (w == "raycast_gather_tf") {
        if (v == "texture") c->var_gather_tf = 0;
        else if (v == "smem") c->var_gather_tf = 1;
        else if (v == "follow") c->var_gather_tf = -1;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_gather_tf is texture|smem|follow");
    } else if (w == "raycast_persist_pct") {
        const int n = std::atoi(variant);
        if (n < 0 || n > 400) return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_persist_pct is 0..400");
        c->var_persist_pct = n;
    } else if (w == "raycast_layout_cos") {
        const float f = (float)std::atof(variant);
        if (!(f >= 0.f && f <= 1.f)) return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_layout_cos is in [0, 1]");
        c->var_layout_cos = f;
    } else if (w == "raycast_tf") {
        if (v == "texture") c->var_tf = 0;
        else if (v == "smem") c->var_tf = 1;
        else return fail(c, VRDD_ERR_INVALID, "set_variant: raycast_tf is texture|smem");
    } else {
        return fail(c, VRDD_ERR_INVALID, "set_variant: unknown kernel");
    }
    return VRDD_OK;
}

// ---- texture-unit probes: for the filter-model conformance tests and tools/probe_texture*.py only.  Built when the
// library is compiled with -DVRDD_PROBE_EXPORTS (csrc/Makefile: PROBES=1, the default of this repository's test build;
// `make PROBES=0` gives the library without them).
#ifdef VRDD_PROBE_EXPORTS
int vrdd_debug_sample_texture(vrdd_handle h, int source, int comp, const float* d_uvw, int n, float* d_out) {
    CHECK_HANDLE(h);
    if (source < 0 || source > 1 || comp < 0 || comp > 2 || !c->vol[source].tex[comp])
        return fail(c, VRDD_ERR_INVALID, "debug_sample_texture: no texture volume");
    return launch_debug_sample(c, c->vol[source].tex[comp], d_uvw, n, d_out);
}

int vrdd_debug_sample_texture_point(vrdd_handle h, int source, int comp, const float* d_uvw, int n, float* d_out) {
    CHECK_HANDLE(h);
    if (source < 0 || source > 1 || comp < 0 || comp > 2 || !c->vol[source].arr[comp])
        return fail(c, VRDD_ERR_INVALID, "debug_sample_texture_point: no texture volume");
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->vol[source].arr[comp];
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    cudaTextureObject_t tex = 0;
    VRDD_CUDA(c, cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    int rc = launch_debug_sample(c, tex, d_uvw, n, d_out);
    cudaStreamSynchronize(c->stream);
    cudaDestroyTextureObject(tex);
    return rc;
}

int vrdd_debug_sample_texture_unnorm(vrdd_handle h, int source, int comp, const float* d_uvw, int n, float* d_out) {
    CHECK_HANDLE(h);
    if (source < 0 || source > 1 || comp < 0 || comp > 2 || !c->vol[source].arr[comp])
        return fail(c, VRDD_ERR_INVALID, "debug_sample_texture_unnorm: no texture volume");
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->vol[source].arr[comp];
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex = 0;
    VRDD_CUDA(c, cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    int rc = launch_debug_sample(c, tex, d_uvw, n, d_out);
    cudaStreamSynchronize(c->stream);
    cudaDestroyTextureObject(tex);
    return rc;
}

int vrdd_debug_sample_transfer_function(vrdd_handle h, const float* d_u, int n, float* d_out4) {
    CHECK_HANDLE(h);
    if (!c->tf_tex || !d_u || !d_out4) return fail(c, VRDD_ERR_INVALID, "debug_sample_transfer_function: bad arguments");
    return launch_debug_sample_tf(c, d_u, n, d_out4);
}
#endif  // VRDD_PROBE_EXPORTS

}  // extern "C"
