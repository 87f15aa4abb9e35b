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
