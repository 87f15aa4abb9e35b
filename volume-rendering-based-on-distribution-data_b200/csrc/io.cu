// io.cu — host-only readers/writers of the reference's on-disk formats (include/vrdd_io.h).
// Restated from the loader code of /root/reference/volumeRender.cpp:538-691; no device code.
#include "../../include/vrdd.h"
#include "../../include/vrdd_io.h"

#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

namespace {

struct File {
    FILE* f;
    explicit File(const char* path, const char* mode) : f(path ? std::fopen(path, mode) : nullptr) {}
    ~File() { if (f) std::fclose(f); }
    template <typename T> bool get(T* v, size_t n = 1) { return std::fread(v, sizeof(T), n, f) == n; }
    template <typename T> bool put(const T* v, size_t n = 1) { return std::fwrite(v, sizeof(T), n, f) == n; }
};

}  // namespace

extern "C" {

int vrdd_io_read_histograms(const char* path, size_t nvox, int bins, float* hist) {
    File fp(path, "rb");
    if (!fp.f || !hist || bins <= 0) return VRDD_ERR_INVALID;
    return fp.get(hist, nvox * (size_t)bins) ? VRDD_OK : VRDD_ERR_INVALID;      // volumeRender.cpp:549-550
}

int64_t vrdd_io_codebook_blocks(const char* path) {
    File fp(path, "rb");
    int32_t steps = 0, blocks = 0;
    if (!fp.f || !fp.get(&steps) || !fp.get(&blocks) || blocks < 0) return VRDD_ERR_INVALID;   // :569-577
    return blocks;
}

int vrdd_io_read_codebook(const char* path, int bins, int64_t nblocks, int32_t* codebook, float* errors_dense) {
    File fp(path, "rb");
    int32_t steps = 0, blocks = 0;
    if (!fp.f || !codebook || !errors_dense || bins <= 0) return VRDD_ERR_INVALID;
    if (!fp.get(&steps) || !fp.get(&blocks) || blocks != nblocks) return VRDD_ERR_INVALID;
    std::vector<int32_t> ids(bins);
    std::vector<double> vals(bins);
    std::memset(errors_dense, 0, sizeof(float) * 2 * (size_t)nblocks * bins);
    for (int64_t i = 0; i < nblocks; ++i) {
        int32_t span = 0, tid = 0, shift = 0, ne = 0;
        unsigned char flip = 0;                                                // `bool` on disk is one byte (:603-604)
        if (!fp.get(&span) || !fp.get(&tid) || !fp.get(&shift) || !fp.get(&flip) || !fp.get(&ne)) return VRDD_ERR_INVALID;
        if (ne > bins || ne < 0) return VRDD_ERR_RANGE;                        // :611-614
        codebook[4 * i + 0] = tid; codebook[4 * i + 1] = shift; codebook[4 * i + 2] = flip ? 1 : 0; codebook[4 * i + 3] = ne;
        if (!fp.get(ids.data(), ne) || !fp.get(vals.data(), ne)) return VRDD_ERR_INVALID;      // :619-637
        for (int k = 0; k < ne; ++k) {
            errors_dense[2 * (i * bins + k) + 0] = (float)ids[k];
            errors_dense[2 * (i * bins + k) + 1] = (float)vals[k];
        }
    }
    return VRDD_OK;
}

int vrdd_io_template_count(const char* path, int bins) {
    (void)bins;
    File fp(path, "rb");
    int32_t n = 0;
    if (!fp.f || !fp.get(&n) || n < 0) return VRDD_ERR_INVALID;                 // :656-659
    return n;
}

int vrdd_io_read_templates(const char* path, int bins, int n, float* templates) {
    File fp(path, "rb");
    int32_t cnt = 0;
    if (!fp.f || !templates || bins <= 0 || !fp.get(&cnt) || cnt != n) return VRDD_ERR_INVALID;
    std::vector<double> row(bins);
    double limits[6];
    for (int i = 0; i < n; ++i) {
        if (!fp.get(limits, 6) || !fp.get(row.data(), bins)) return VRDD_ERR_INVALID;          // :664-675
        for (int k = 0; k < bins; ++k) templates[(size_t)i * bins + k] = (float)row[k];         // :681
    }
    return VRDD_OK;
}

int vrdd_io_span_count(const char* path) {
    File fp(path, "rb");
    int32_t n = 0;
    if (!fp.f || !fp.get(&n) || n < 0) return VRDD_ERR_INVALID;                 // volumeRender.cpp:721-723
    return n;
}

int vrdd_io_read_span_list(const char* path, int n, int32_t* low, int32_t* high) {
    File fp(path, "rb");
    int32_t cnt = 0;
    if (!fp.f || !low || !high || !fp.get(&cnt) || cnt != n) return VRDD_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        int32_t v[6];                                                            // lowX, highX, lowY, highY, lowZ, highZ (:744-749)
        if (!fp.get(v, 6)) return VRDD_ERR_INVALID;
        low[4 * i] = v[0]; low[4 * i + 1] = v[2]; low[4 * i + 2] = v[4]; low[4 * i + 3] = 0;
        high[4 * i] = v[1]; high[4 * i + 1] = v[3]; high[4 * i + 2] = v[5]; high[4 * i + 3] = 0;
        if (v[0] > v[1] || v[2] > v[3] || v[4] > v[5] || v[0] < 0 || v[2] < 0 || v[4] < 0) return VRDD_ERR_RANGE;   // checkSpanLimit, :693-699
    }
    return VRDD_OK;
}

int vrdd_io_read_flex_codebook(const char* path, int bins, int64_t n, int32_t* span_ids, int32_t* codebook, float* errors_dense) {
    File fp(path, "rb");
    int32_t steps = 0, cnt = 0;
    if (!fp.f || !span_ids || !codebook || !errors_dense || bins <= 0) return VRDD_ERR_INVALID;
    if (!fp.get(&steps) || !fp.get(&cnt) || cnt != n) return VRDD_ERR_INVALID;              // :784-788
    std::vector<int32_t> ids(bins);
    std::vector<double> vals(bins);
    std::memset(errors_dense, 0, sizeof(float) * 2 * (size_t)n * bins);
    for (int64_t i = 0; i < n; ++i) {
        int32_t span = -1, tid = -1, shift = -1, ne = -1;
        unsigned char flip = 0;
        if (!fp.get(&span) || !fp.get(&tid) || !fp.get(&shift) || !fp.get(&flip) || !fp.get(&ne)) return VRDD_ERR_INVALID;
        if (span < 0 || span > 2 * n || tid < 0 || ne > bins || ne < 0) return VRDD_ERR_RANGE;   // :799-836
        span_ids[i] = span;
        codebook[4 * i] = tid; codebook[4 * i + 1] = shift; codebook[4 * i + 2] = flip ? 1 : 0; codebook[4 * i + 3] = ne;
        if (!fp.get(ids.data(), ne) || !fp.get(vals.data(), ne)) return VRDD_ERR_INVALID;
        for (int k = 0; k < ne; ++k) {
            errors_dense[2 * (i * bins + k)] = (float)ids[k];
            errors_dense[2 * (i * bins + k) + 1] = (float)vals[k];
        }
    }
    return VRDD_OK;
}

int vrdd_io_simple_count(const char* counts_path) {
    File fp(counts_path, "rb");
    int32_t n = -1;
    if (!fp.f || !fp.get(&n) || n < 0) return VRDD_ERR_INVALID;                 // :893-895
    return n;
}

int vrdd_io_read_simple(const char* counts_path, const char* ids_path, const char* freqs_path, int bins, int n, int32_t* low,
                        int32_t* high, int32_t* count, float* hist) {
    File fc(counts_path, "rb"), fi(ids_path, "rb"), ff(freqs_path, "rb");
    int32_t cnt = -1;
    if (!fc.f || !fi.f || !ff.f || !low || !high || !count || !hist || bins <= 0) return VRDD_ERR_INVALID;
    if (!fc.get(&cnt) || cnt != n) return VRDD_ERR_INVALID;
    std::memset(hist, 0, sizeof(float) * 2 * (size_t)n * bins);
    for (int i = 0; i < n; ++i) {
        int32_t v[7];                                                            // low xyz, high xyz, count (:906-925)
        if (!fc.get(v, 7)) return VRDD_ERR_INVALID;
        low[4 * i] = v[0]; low[4 * i + 1] = v[1]; low[4 * i + 2] = v[2]; low[4 * i + 3] = 0;
        high[4 * i] = v[3]; high[4 * i + 1] = v[4]; high[4 * i + 2] = v[5]; high[4 * i + 3] = 0;
        count[i] = v[6];
        if (v[6] < 0 || v[6] > bins) return VRDD_ERR_RANGE;                      // :926-929
        for (int k = 0; k < v[6]; ++k) {
            int32_t id = -1; double fr = -1.0;
            if (!fi.get(&id) || !ff.get(&fr)) return VRDD_ERR_INVALID;
            if (id < 0 || id > bins || fr < 0 || fr > 1.0) return VRDD_ERR_RANGE;   // checkHistogram, :701-707
            hist[2 * ((size_t)i * bins + k)] = (float)id;
            hist[2 * ((size_t)i * bins + k) + 1] = (float)fr;
        }
    }
    return VRDD_OK;
}

int vrdd_io_write_span_list(const char* path, int n, const int32_t* low, const int32_t* high) {
    File fp(path, "wb");
    const int32_t cnt = n;
    if (!fp.f || !low || !high || !fp.put(&cnt)) return VRDD_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        const int32_t v[6] = {low[4 * i], high[4 * i], low[4 * i + 1], high[4 * i + 1], low[4 * i + 2], high[4 * i + 2]};
        if (!fp.put(v, 6)) return VRDD_ERR_INVALID;
    }
    return VRDD_OK;
}

int vrdd_io_write_flex_codebook(const char* path, int bins, int64_t n, const int32_t* span_ids, const int32_t* codebook,
                                const float* errors_dense) {
    File fp(path, "wb");
    if (!fp.f || !span_ids || !codebook || !errors_dense) return VRDD_ERR_INVALID;
    const int32_t steps = 1, cnt = (int32_t)n;
    if (!fp.put(&steps) || !fp.put(&cnt)) return VRDD_ERR_INVALID;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t span = span_ids[i], tid = codebook[4 * i], shift = codebook[4 * i + 1], ne = codebook[4 * i + 3];
        const unsigned char flip = codebook[4 * i + 2] ? 1 : 0;
        if (ne < 0 || ne > bins) return VRDD_ERR_RANGE;
        if (!fp.put(&span) || !fp.put(&tid) || !fp.put(&shift) || !fp.put(&flip) || !fp.put(&ne)) return VRDD_ERR_INVALID;
        for (int k = 0; k < ne; ++k) { const int32_t id = (int32_t)errors_dense[2 * (i * bins + k)]; if (!fp.put(&id)) return VRDD_ERR_INVALID; }
        for (int k = 0; k < ne; ++k) { const double v = (double)errors_dense[2 * (i * bins + k) + 1]; if (!fp.put(&v)) return VRDD_ERR_INVALID; }
    }
    return VRDD_OK;
}

int vrdd_io_write_simple(const char* counts_path, const char* ids_path, const char* freqs_path, int bins, int n,
                         const int32_t* low, const int32_t* high, const int32_t* count, const float* hist) {
    File fc(counts_path, "wb"), fi(ids_path, "wb"), ff(freqs_path, "wb");
    const int32_t cnt = n;
    if (!fc.f || !fi.f || !ff.f || !low || !high || !count || !hist || !fc.put(&cnt)) return VRDD_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        const int32_t v[7] = {low[4 * i], low[4 * i + 1], low[4 * i + 2], high[4 * i], high[4 * i + 1], high[4 * i + 2], count[i]};
        if (count[i] < 0 || count[i] > bins || !fc.put(v, 7)) return VRDD_ERR_INVALID;
        for (int k = 0; k < count[i]; ++k) {
            const int32_t id = (int32_t)hist[2 * ((size_t)i * bins + k)];
            const double fr = (double)hist[2 * ((size_t)i * bins + k) + 1];
            if (!fi.put(&id) || !ff.put(&fr)) return VRDD_ERR_INVALID;
        }
    }
    return VRDD_OK;
}

int vrdd_io_write_histograms(const char* path, size_t nvox, int bins, const float* hist) {
    File fp(path, "wb");
    if (!fp.f || !hist) return VRDD_ERR_INVALID;
    return fp.put(hist, nvox * (size_t)bins) ? VRDD_OK : VRDD_ERR_INVALID;
}

int vrdd_io_write_codebook(const char* path, int bins, int64_t nblocks, const int32_t* codebook, const float* errors_dense) {
    File fp(path, "wb");
    if (!fp.f || !codebook || !errors_dense) return VRDD_ERR_INVALID;
    const int32_t steps = 1, blocks = (int32_t)nblocks;
    if (!fp.put(&steps) || !fp.put(&blocks)) return VRDD_ERR_INVALID;
    for (int64_t i = 0; i < nblocks; ++i) {
        const int32_t span = (int32_t)i, tid = codebook[4 * i], shift = codebook[4 * i + 1], ne = codebook[4 * i + 3];
        const unsigned char flip = codebook[4 * i + 2] ? 1 : 0;
        if (ne < 0 || ne > bins) return VRDD_ERR_RANGE;
        if (!fp.put(&span) || !fp.put(&tid) || !fp.put(&shift) || !fp.put(&flip) || !fp.put(&ne)) return VRDD_ERR_INVALID;
        for (int k = 0; k < ne; ++k) {
            const int32_t id = (int32_t)errors_dense[2 * (i * bins + k)];
            if (!fp.put(&id)) return VRDD_ERR_INVALID;
        }
        for (int k = 0; k < ne; ++k) {
            const double v = (double)errors_dense[2 * (i * bins + k) + 1];
            if (!fp.put(&v)) return VRDD_ERR_INVALID;
        }
    }
    return VRDD_OK;
}

int vrdd_io_write_templates(const char* path, int bins, int n, const float* templates) {
    File fp(path, "wb");
    if (!fp.f || !templates) return VRDD_ERR_INVALID;
    const int32_t cnt = n;
    if (!fp.put(&cnt)) return VRDD_ERR_INVALID;
    const double limits[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        if (!fp.put(limits, 6)) return VRDD_ERR_INVALID;
        for (int k = 0; k < bins; ++k) {
            const double v = (double)templates[(size_t)i * bins + k];
            if (!fp.put(&v)) return VRDD_ERR_INVALID;
        }
    }
    return VRDD_OK;
}

// d_divideBlock, volumeRender_kernel.cu:892-1031 (per-axis; the reference mixes the axes up for
// non-cubic volumes, :949-1013, which only its cubic 64^3 data set never notices)
int vrdd_flex_divide_blocks(int vx, int vy, int vz, int block, int32_t* spans, int capacity) {
    if (vx <= 0 || vy <= 0 || vz <= 0 || block <= 0) return VRDD_ERR_INVALID;
    const int nx = (vx + block - 1) / block, ny = (vy + block - 1) / block, nz = (vz + block - 1) / block;
    const long long n = (long long)nx * ny * nz;
    if (n > 0x7fffffff) return VRDD_ERR_INVALID;
    if (!spans) return (int)n;
    if (capacity < n) return VRDD_ERR_INVALID;
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                int32_t* s = spans + 6 * ((size_t)z * nx * ny + (size_t)y * nx + x);
                s[0] = 1 + x * block; s[1] = 1 + y * block; s[2] = 1 + z * block;          // :937, 1-based
                s[3] = (x == nx - 1) ? vx : (x + 1) * block;                               // :940-946, clipped
                s[4] = (y == ny - 1) ? vy : (y + 1) * block;
                s[5] = (z == nz - 1) ? vz : (z + 1) * block;
            }
    return (int)n;
}

// the bit trick of d_queryBlockNew, volumeRender_kernel.cu:1248-1259
int vrdd_flex_prefix_spans(int x, int32_t* spans) {
    if (x < 0 || !spans) return VRDD_ERR_INVALID;
    int n = 0;
    for (int i = 0; i < 31 && x != 0; ++i)
        if (x & (1 << i)) {
            spans[2 * n + 1] = x;
            x &= ~(1 << i);
            spans[2 * n] = x + 1;
            ++n;
        }
    return n;
}

int vrdd_io_write_ppm(const char* path, const uint32_t* rgba, int width, int height) {
    File fp(path, "wb");
    if (!fp.f || !rgba || width <= 0 || height <= 0) return VRDD_ERR_INVALID;
    std::fprintf(fp.f, "P6\n%d %d\n255\n", width, height);
    std::vector<unsigned char> row(3 * (size_t)width);
    for (int y = 0; y < height; ++y) {
        for (int x = 0; x < width; ++x) {
            const uint32_t p = rgba[(size_t)y * width + x];
            row[3 * x] = p & 255; row[3 * x + 1] = (p >> 8) & 255; row[3 * x + 2] = (p >> 16) & 255;
        }
        if (!fp.put(row.data(), row.size())) return VRDD_ERR_INVALID;
    }
    return VRDD_OK;
}

int vrdd_io_read_ppm(const char* path, uint8_t* rgb, int width, int height) {
    File fp(path, "rb");
    int w = 0, h = 0, mx = 0;
    if (!fp.f || !rgb || std::fscanf(fp.f, "P6 %d %d %d", &w, &h, &mx) != 3 || w != width || h != height || mx != 255)
        return VRDD_ERR_INVALID;
    std::fgetc(fp.f);                                                           // the single whitespace after maxval
    return fp.get(rgb, 3 * (size_t)w * h) ? VRDD_OK : VRDD_ERR_INVALID;
}

}  // extern "C"
