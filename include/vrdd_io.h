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
 *   span list        int32 n, then per span int32 lowX, highX, lowY, highY, lowZ, highZ
 *                    (interleaved like that, volumeRender.cpp:744-749)               (:709-771)
 *   flexible codebook  the codebook format with 64 bins; spanId selects the row of the span list
 *                    whose (low, high) the code belongs to                           (:773-875)
 *   flexible templates the template format with 64 bins                             (:951-997)
 *   simple histograms  three files: counts  int32 n, then per span int32 low[3], high[3], count;
 *                    bin ids  int32 stream;  frequencies  float64 stream             (:877-949)
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

/* Span list: low/high int32[n][4] = (x, y, z, 0).  vrdd_io_span_count returns n or a negative status. */
int vrdd_io_span_count(const char* path);
int vrdd_io_read_span_list(const char* path, int n, int32_t* low, int32_t* high);
/* Flexible codebook: like vrdd_io_read_codebook plus the span id of every row (codebookSpanLow[i] =
 * spanLow[spanId], volumeRender.cpp:838-839).  Row count: vrdd_io_codebook_blocks. */
int vrdd_io_read_flex_codebook(const char* path, int bins, int64_t n, int32_t* span_ids, int32_t* codebook,
                               float* errors_dense);
/* Simple histograms: low/high int32[n][4] (0-based), count int32[n], hist float[n][bins][2].
 * vrdd_io_simple_count returns n or a negative status. */
int vrdd_io_simple_count(const char* counts_path);
int vrdd_io_read_simple(const char* counts_path, const char* ids_path, const char* freqs_path, int bins, int n,
                        int32_t* low, int32_t* high, int32_t* count, float* hist);

/* Writers of the same formats (test fixtures / exporting synthetic volumes). */
int vrdd_io_write_span_list(const char* path, int n, const int32_t* low, const int32_t* high);
int vrdd_io_write_flex_codebook(const char* path, int bins, int64_t n, const int32_t* span_ids, const int32_t* codebook,
                                const float* errors_dense);
int vrdd_io_write_simple(const char* counts_path, const char* ids_path, const char* freqs_path, int bins, int n,
                         const int32_t* low, const int32_t* high, const int32_t* count, const float* hist);
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
