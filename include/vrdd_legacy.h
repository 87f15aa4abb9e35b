/*
 * vrdd_legacy.h — the reference's own host-facing surface, exported unchanged by
 * libvrdd.so so that /root/reference/volumeRender.cpp links against it in place of
 * volumeRender_kernel.cu.  Declarations mirror volumeRender.cpp:156-170; definitions they
 * replace are cited per symbol.  Requires the CUDA runtime headers for cudaExtent, dim3,
 * int4 and float2, exactly as the reference's translation unit does.
 *
 * All seven operate on one process-wide default vrdd handle (the reference keeps its state
 * in file-scope globals, volumeRender_kernel.cu:22-43).  Differences from the reference,
 * all deliberate:
 *   - any volumeSize is accepted, not only 50x50x10 (d_basicDataProcessing hard-codes it,
 *     volumeRender_kernel.cu:727-729);
 *   - CUDA errors are reported on stderr and through vrdd_last_error(vrdd_legacy_handle())
 *     instead of exit() inside checkCudaErrors;
 *   - freeCudaBuffers frees only what was allocated (the reference frees a never-allocated
 *     array and a shadowed global, volumeRender_kernel.cu:2362, 2366);
 *   - the flexible-block tables (last nine arguments of initCuda) may be NULL; then
 *     dataProcessing() says so and queryMethod 8/9/0 are unavailable.  When given they must have
 *     the sizes the reference hard-codes (131 072 spans each, 469 templates of 64 bins,
 *     volumeRender_kernel.cu:96-101); rows with span_low.x < 0 are padding.
 */
#ifndef VRDD_LEGACY_H_
#define VRDD_LEGACY_H_

#include <cuda_runtime_api.h>
#include <vector_types.h>
#include "vrdd.h"

#ifdef __cplusplus
extern "C" {
#endif

/* volumeRender_kernel.cu:1893 */
void initCuda(void* h_volume, cudaExtent volumeSize, cudaExtent histogramSize, int4* h_codebook,
              cudaExtent codebookSize, float* h_templates, cudaExtent templatesSize,
              float2* h_errorsbook, cudaExtent errorsbookSize, int4* h_codebookSpanLow,
              int4* h_codebookSpanHigh, int4* h_flexibleCodebook, float2* h_flexibleErrorsbook,
              int4* h_simpleLow, int4* h_simpleHigh, int* h_simpleCount, float2* h_simpleHistogram,
              float* h_flexibleTemplates);
/* volumeRender_kernel.cu:1798 */
void basicDataProcessing(void);
/* volumeRender_kernel.cu:1735 */
void dataProcessing(void);
/* volumeRender_kernel.cu:2403 */
void copyInvViewMatrix(float* invViewMatrix, size_t sizeofMatrix);
/* volumeRender_kernel.cu:2387 */
void render_kernel(dim3 gridSize, dim3 blockSize, unsigned int* d_output, unsigned int imageW,
                   unsigned int imageH, float density, float brightness, float transferOffset,
                   float transferScale, int queryMethod, cudaExtent volumeSize);
/* volumeRender_kernel.cu:1889 */
void setTextureFilterMode(bool bLinearFilter);
/* volumeRender_kernel.cu:2360 */
void freeCudaBuffers(void);

/* The default handle the seven symbols above act on (NULL before initCuda). */
vrdd_handle vrdd_legacy_handle(void);

#ifdef __cplusplus
}
#endif
#endif /* VRDD_LEGACY_H_ */
