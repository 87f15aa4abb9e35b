/*
 * vrdd_io.h — readers (and writers, for tests) of the reference's on-disk formats
 * (SURVEY.md §8f row 2).  Host-only functions of libvrdd.so; no GPU needed.  They replace
 * loadRawFile / loadCodebook / loadTemplates of /root/reference/volumeRender.cpp:538-691 and
 * return the arrays in the layout initCuda expects (volumeRender.cpp:1200-1203).
 *
 * None of the reference's nine .bin files ship with it, so these formats are restated from the
 * loader code alone:
 *   raw histograms   float32[V*bins]                                        (:538-556)
 *   codebook         int32 nSteps, int32 nBlocks, then per block:
 *                    int32 spanId (ignored), int32 templateId, int32 shift, 1-byte bool flip,
 *                    int32 NE, int32 binIds[NE], float64 values[NE]          (:558-642)
 *   templates        int32 n, then per template: float64 limits[6] (ignored),
 *                    float64 freq[bins]                                      (:644-691)
 * All little-endian, packed (the loader freads field by field).
 */
#ifndef VRDD_IO_H_
#define VRDD_IO_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Each reader returns 0 on success, a negative vrdd_status otherwise (VRDD_ERR_INVALID: cannot
 * open / short file; VRDD_ERR_RANGE: NE > bins, volumeRender.cpp:611-614). */

/* float hist[nvox*bins] (caller-allocated). */
int vrdd_io_read_histograms(const char* path, size_t nvox, int bins, float* hist);
/* Number of blocks (voxels) announced by a codebook file, or a negative status. */
int64_t vrdd_io_codebook_blocks(const char* path);
/* codebook int32[nblocks][4] = (templateId, shift, flip, NE); errors_dense float[nblocks][bins][2],
 * rows zero-filled beyond NE (the reference leaves them uninitialised, :582). */
int vrdd_io_read_codebook(const char* path, int bins, int64_t nblocks, int32_t* codebook, float* errors_dense);
/* Number of templates announced by a template file, or a negative status. */
int vrdd_io_template_count(const char* path, int bins);
/* float templates[n][bins] (doubles narrowed like volumeRender.cpp:681). */
int vrdd_io_read_templates(const char* path, int bins, int n, float* templates);

/* Writers of the same formats (test fixtures / exporting synthetic volumes). */
int vrdd_io_write_histograms(const char* path, size_t nvox, int bins, const float* hist);
int vrdd_io_write_codebook(const char* path, int bins, int64_t nblocks, const int32_t* codebook, const float* errors_dense);
int vrdd_io_write_templates(const char* path, int bins, int n, const float* templates);

/* Binary PPM (P6) of an RGBA8 frame, rows in memory order, alpha dropped — what
 * sdkSavePPM4ub does for runSingleTest (volumeRender.cpp:1076). */
int vrdd_io_write_ppm(const char* path, const uint32_t* rgba, int width, int height);
int vrdd_io_read_ppm(const char* path, uint8_t* rgb, int width, int height);

#ifdef __cplusplus
}
#endif
#endif
