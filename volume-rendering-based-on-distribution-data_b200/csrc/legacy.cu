// legacy.cu — the reference's seven extern "C" entry points (volumeRender.cpp:156-170) on top
// of one default vrdd handle, so volumeRender.cpp links against libvrdd.so unchanged.
// See include/vrdd_legacy.h for the deliberate differences.
#include "common.cuh"
#include "../../include/vrdd_legacy.h"

#include <cstdio>

namespace {

vrdd_handle g_handle = nullptr;
bool g_have_fractal = false;
bool g_have_flex = false;

void report(const char* where, int rc) {
    if (rc == VRDD_OK) return;
    std::fprintf(stderr, "libvrdd: %s failed (%d): %s\n", where, rc, g_handle ? vrdd_last_error(g_handle) : "no handle");
}

}  // namespace

extern "C" {

vrdd_handle vrdd_legacy_handle(void) { return g_handle; }

// volumeRender_kernel.cu:1893-2358.  histogramSize is (bins, W*H, D) (volumeRender.cpp:87);
// templatesSize is (bins, numTemplates, 1) (volumeRender.cpp:89).  Host pointers are copied
// synchronously; the caller frees them right after (volumeRender.cpp:1205-1218).
void initCuda(void* h_volume, cudaExtent volumeSize, cudaExtent histogramSize, int4* h_codebook,
              cudaExtent codebookSize, float* h_templates, cudaExtent templatesSize, float2* h_errorsbook,
              cudaExtent errorsbookSize, int4* h_codebookSpanLow, int4* h_codebookSpanHigh, int4* h_flexibleCodebook,
              float2* h_flexibleErrorsbook, int4* h_simpleLow, int4* h_simpleHigh, int* h_simpleCount,
              float2* h_simpleHistogram, float* h_flexibleTemplates) {
    (void)codebookSize; (void)errorsbookSize;
    if (!g_handle) {
        int rc = vrdd_create(-1, &g_handle);
        if (rc != VRDD_OK) { std::fprintf(stderr, "libvrdd: initCuda: no usable CUDA device (%d)\n", rc); return; }
    }
    int rc = vrdd_set_volume(g_handle, (int)volumeSize.width, (int)volumeSize.height, (int)volumeSize.depth,
                             (int)histogramSize.width);
    report("initCuda/set_volume", rc);
    if (rc != VRDD_OK) return;
    vrdd_enable_interpolated_mean(g_handle, 1);                                  // queryMethod 7 stays available
    if (h_volume) report("initCuda/histograms", vrdd_set_histograms_host(g_handle, static_cast<const float*>(h_volume)));
    g_have_fractal = false;
    if (h_codebook && h_templates && h_errorsbook) {
        rc = vrdd_set_fractal_host(g_handle, reinterpret_cast<const int32_t*>(h_codebook),
                                   reinterpret_cast<const float*>(h_errorsbook), h_templates, (int)templatesSize.height);
        report("initCuda/fractal", rc);
        g_have_fractal = rc == VRDD_OK;
    }
    report("initCuda/transfer_function", vrdd_set_transfer_function(g_handle, nullptr, 0));  // :2322-2344
    // the flexible-block tables, with the sizes the reference hard-codes (volumeRender_kernel.cu:96-101):
    // 64 x 64 x 32 = 131 072 fractal spans and as many simple spans, 64 bins, 469 templates, a 64^3 raw volume
    g_have_flex = false;
    if (h_codebookSpanLow && h_codebookSpanHigh && h_flexibleCodebook && h_flexibleErrorsbook && h_simpleLow && h_simpleHigh &&
        h_simpleCount && h_simpleHistogram && h_flexibleTemplates) {
        vrdd_flex_tables t;
        t.raw_w = t.raw_h = t.raw_d = 64; t.bins = 64;
        t.n_fractal = 64 * 64 * 32; t.n_simple = 64 * 64 * 32; t.n_templates = 469;
        t.span_low = reinterpret_cast<const int32_t*>(h_codebookSpanLow); t.span_high = reinterpret_cast<const int32_t*>(h_codebookSpanHigh);
        t.codebook = reinterpret_cast<const int32_t*>(h_flexibleCodebook); t.errors = reinterpret_cast<const float*>(h_flexibleErrorsbook);
        t.simple_low = reinterpret_cast<const int32_t*>(h_simpleLow); t.simple_high = reinterpret_cast<const int32_t*>(h_simpleHigh);
        t.simple_count = h_simpleCount; t.simple_hist = reinterpret_cast<const float*>(h_simpleHistogram);
        t.templates = h_flexibleTemplates;
        rc = vrdd_flex_set_tables_host(g_handle, &t);
        report("initCuda/flexible tables", rc);
        g_have_flex = rc == VRDD_OK;
    }
}

// volumeRender_kernel.cu:1798-1887: decode both volumes and leave them sampleable.
void basicDataProcessing(void) {
    if (!g_handle) { std::fprintf(stderr, "libvrdd: basicDataProcessing before initCuda\n"); return; }
    report("basicDataProcessing/original", vrdd_decode(g_handle, VRDD_SRC_ORIGINAL, 0, 0));
    if (g_have_fractal) report("basicDataProcessing/fractal", vrdd_decode(g_handle, VRDD_SRC_FRACTAL, 0, 0));
}

// volumeRender_kernel.cu:1735-1796: the flexible-block-size chain with the reference's hard-coded block size 6
// (:1737), timed like the reference times its stages (:1739-1783).
void dataProcessing(void) {
    if (!g_handle || !g_have_flex) {
        std::fprintf(stderr, "libvrdd: dataProcessing(): no flexible-block tables were given to initCuda; "
                             "queryMethod 8/9/0 are unavailable\n");
        return;
    }
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, g_handle->stream);
    int64_t missing = 0;
    const int rc = vrdd_flex_process(g_handle, 6, nullptr);
    cudaEventRecord(t1, g_handle->stream);
    cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    report("dataProcessing", rc);
    if (rc == VRDD_OK) {
        int dims[3];
        vrdd_flex_get_blocks_host(g_handle, nullptr, dims);
        std::printf("x %d, y %d, z %d, total %d\n", dims[0], dims[1], dims[2], dims[0] * dims[1] * dims[2]);   // bindToTex, :1586
        std::printf("d_queryBlockNew() + d_querySpanNew() + d_computeBlock(): %f ms\n", ms);
    }
    (void)missing;
}

// volumeRender_kernel.cu:2403-2406
void copyInvViewMatrix(float* invViewMatrix, size_t sizeofMatrix) {
    if (!g_handle || !invViewMatrix || sizeofMatrix < sizeof(float) * 12) return;
    vrdd_set_view(g_handle, invViewMatrix);
}

// volumeRender_kernel.cu:2387-2401.  gridSize/blockSize only had to cover the image in the
// reference (one thread per pixel, :282-286); every pixel is still computed exactly once,
// with this library's own launch shape.  volumeSize is what the reference forwards to its
// mode-7 code; the decoded volume's own size is authoritative here.  Asynchronous.
void render_kernel(dim3 gridSize, dim3 blockSize, unsigned int* d_output, unsigned int imageW, unsigned int imageH,
                   float density, float brightness, float transferOffset, float transferScale, int queryMethod,
                   cudaExtent volumeSize) {
    (void)gridSize; (void)blockSize; (void)volumeSize;
    if (!g_handle) return;
    vrdd_render_params p;
    vrdd_default_render_params(&p);
    p.density = density; p.brightness = brightness; p.transfer_offset = transferOffset; p.transfer_scale = transferScale;
    p.query_method = queryMethod;
    report("render_kernel", vrdd_render(g_handle, d_output, (int)imageW, (int)imageH, &p, nullptr, 0));
}

// volumeRender_kernel.cu:1889-1891 toggles the filter of the block-index texture `tex`,
// which only the interpolated-mean mode (7) and a dead fetch (:386) read.  Recorded only.
void setTextureFilterMode(bool bLinearFilter) {
    if (g_handle) g_handle->legacy_linear_filter = bLinearFilter;
}

// volumeRender_kernel.cu:2360-2385
void freeCudaBuffers(void) {
    if (g_handle) vrdd_destroy(g_handle);
    g_handle = nullptr;
    g_have_fractal = false;
    g_have_flex = false;
}

}  // extern "C"
