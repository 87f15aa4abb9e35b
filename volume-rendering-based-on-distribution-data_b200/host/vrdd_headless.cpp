// vrdd_headless.cpp — headless host driver: the reference's volumeRender.cpp without the
// GLUT/GLEW shell.  Host C++ only; it reaches the GPU exclusively through
// the reference's own seven extern "C" entry points (volumeRender.cpp:156-170), now exported by
// libvrdd.so, in the order main() and runSingleTest() call them:
//
//   initCuda (:1200) -> dataProcessing (:1220) -> basicDataProcessing (:1221) ->
//   per view: copyInvViewMatrix (:1046) -> cudaMemset (:1022) -> render_kernel (:1059) ->
//   cudaMemcpy D2H (:1074) -> PPM (:1076) [-> compare (:1077)] -> freeCudaBuffers (:462)
//
// What replaces the window: an orbit camera built with vrdd_view_matrix (the GL matrix of
// :224-246 without OpenGL), PPM frames on disk, and the --file self-test with a +-1 LSB rule
// instead of sdkComparePPM's eps 5 / 30 % (:57-58).
//
// Inputs: the reference's .bin formats (include/vrdd_io.h) or a seeded synthetic volume
// (include/vrdd_synth.h) when no files are given — none of the reference's data files ship.
//
// Flags (the reference's own, volumeRender.cpp:1100-1153, plus headless ones):
//   --file=<ref.ppm>  --volume=<hist.bin>  --size=N --xsize=N --ysize=N --zsize=N
//   --codebook=<f> --templates=<f>  --seed=N  --width=N --height=N  --views=N  --rotx=deg --roty=deg
//   --query=1..6  --density=f --brightness=f --offset=f --scale=f  --out=<prefix>  --iters=N
#include <cuda_runtime_api.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vrdd_io.h"
#include "../../include/vrdd_legacy.h"
#include "../../include/vrdd_synth.h"

namespace {

const char* arg_str(int argc, char** argv, const char* name, const char* dflt) {
    const size_t n = std::strlen(name);
    for (int i = 1; i < argc; ++i) {
        const char* a = argv[i];
        while (*a == '-') ++a;                                   // the SDK helper accepts -x and --x
        if (std::strncmp(a, name, n) == 0 && a[n] == '=') return a + n + 1;
    }
    return dflt;
}
bool arg_flag(int argc, char** argv, const char* name) {
    return arg_str(argc, argv, name, nullptr) != nullptr;
}
double arg_num(int argc, char** argv, const char* name, double dflt) {
    const char* s = arg_str(argc, argv, name, nullptr);
    return s ? std::atof(s) : dflt;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            std::fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e__));          \
            return EXIT_FAILURE;                                                              \
        }                                                                                     \
    } while (0)

}  // namespace

int main(int argc, char** argv) {
    std::printf("vrdd_headless: CUDA 3D Volume Render (distribution data), headless\n");
    // ---- sizes: the reference's defaults (volumeRender.cpp:86-90, 121) and flags (:1131-1153)
    size_t W = 50, H = 50, D = 10;
    const int bins = 32;
    if (arg_flag(argc, argv, "size")) W = H = D = (size_t)arg_num(argc, argv, "size", 50);
    W = (size_t)arg_num(argc, argv, "xsize", (double)W);
    H = (size_t)arg_num(argc, argv, "ysize", (double)H);
    D = (size_t)arg_num(argc, argv, "zsize", (double)D);
    const unsigned width = (unsigned)arg_num(argc, argv, "width", 512), height = (unsigned)arg_num(argc, argv, "height", 512);
    const int views = (int)arg_num(argc, argv, "views", 1), iters = (int)arg_num(argc, argv, "iters", 10);
    const int query = (int)arg_num(argc, argv, "query", 1);                      // queryMethod, :129
    const float density = (float)arg_num(argc, argv, "density", 0.05), brightness = (float)arg_num(argc, argv, "brightness", 1.0);
    const float offset = (float)arg_num(argc, argv, "offset", 0.0), scale = (float)arg_num(argc, argv, "scale", 1.0);   // :130-133
    const float rotx = (float)arg_num(argc, argv, "rotx", 0.0), roty0 = (float)arg_num(argc, argv, "roty", 0.0);
    const uint32_t seed = (uint32_t)arg_num(argc, argv, "seed", 1234);
    const char* ref_file = arg_str(argc, argv, "file", nullptr);
    const char* out_prefix = arg_str(argc, argv, "out", "volume");
    const size_t V = W * H * D;

    // ---- load or synthesise the inputs (L2 of the reference: volumeRender.cpp:1180-1187)
    std::vector<float> hist(V * bins);
    const char* vol_file = arg_str(argc, argv, "volume", nullptr);
    if (vol_file) {
        if (vrdd_io_read_histograms(vol_file, V, bins, hist.data()) != VRDD_OK) {
            std::fprintf(stderr, "Error opening file '%s'\n", vol_file);
            return EXIT_FAILURE;
        }
        std::printf("Read '%s', %zu float\n", vol_file, V * bins);
    } else {
        for (size_t z = 0; z < D; ++z)
            for (size_t y = 0; y < H; ++y)
                for (size_t x = 0; x < W; ++x)
                    vrdd_synth_histogram(seed, (int)x, (int)y, (int)z, (int)W, (int)H, (int)D, bins,
                                         &hist[(x + W * (y + H * z)) * bins]);
        std::printf("synthetic histograms: %zux%zux%zu x %d bins, seed %u\n", W, H, D, bins, seed);
    }
    std::vector<int32_t> codebook;
    std::vector<float> errors, templates;
    int T = 0;
    const char* cb_file = arg_str(argc, argv, "codebook", nullptr);
    const char* tm_file = arg_str(argc, argv, "templates", nullptr);
    if (cb_file && tm_file) {
        T = vrdd_io_template_count(tm_file, bins);
        if (T <= 0 || vrdd_io_codebook_blocks(cb_file) != (int64_t)V) {
            std::fprintf(stderr, "Wrong Codebook or Templates!\n");                // :1189-1192
            return EXIT_FAILURE;
        }
        codebook.resize(4 * V); errors.resize(2 * V * bins); templates.resize((size_t)T * bins);
        if (vrdd_io_read_codebook(cb_file, bins, (int64_t)V, codebook.data(), errors.data()) != VRDD_OK ||
            vrdd_io_read_templates(tm_file, bins, T, templates.data()) != VRDD_OK) {
            std::fprintf(stderr, "Wrong Codebook or Templates!\n");
            return EXIT_FAILURE;
        }
        std::printf("nBlocks: %zu\nnTemplates: %d\n", V, T);
    } else {
        T = 622;                                                                   // :89
        codebook.resize(4 * V); errors.assign(2 * V * bins, 0.f); templates.resize((size_t)T * bins);
        for (int k = 0; k < T; ++k) vrdd_synth_template(seed, k, T, bins, &templates[(size_t)k * bins]);
        for (size_t z = 0; z < D; ++z)
            for (size_t y = 0; y < H; ++y)
                for (size_t x = 0; x < W; ++x) {
                    const size_t v = x + W * (y + H * z);
                    int code[4], eb[VRDD_SYNTH_MAX_BINS];
                    float evv[VRDD_SYNTH_MAX_BINS];
                    vrdd_synth_fractal_code(seed, (int)x, (int)y, (int)z, (int)W, (int)H, (int)D, bins, T, 8, code, eb, evv);
                    for (int k = 0; k < 4; ++k) codebook[4 * v + k] = code[k];
                    for (int k = 0; k < code[3]; ++k) {
                        errors[2 * (v * bins + k)] = (float)eb[k];
                        errors[2 * (v * bins + k) + 1] = evv[k];
                    }
                }
    }

    // ---- device set-up through the reference's own entry points (:1200-1221)
    auto extent = [](size_t w, size_t h, size_t d) { cudaExtent e; e.width = w; e.height = h; e.depth = d; return e; };
    const cudaExtent volumeSize = extent(W, H, D), histogramSize = extent(bins, W * H, D);         // :86-87
    const cudaExtent templatesSize = extent(bins, T, 1);                                           // :89
    initCuda(hist.data(), volumeSize, histogramSize, reinterpret_cast<int4*>(codebook.data()), volumeSize, templates.data(),
             templatesSize, reinterpret_cast<float2*>(errors.data()), histogramSize, nullptr, nullptr, nullptr, nullptr, nullptr,
             nullptr, nullptr, nullptr, nullptr);
    if (!vrdd_legacy_handle()) return EXIT_FAILURE;
    dataProcessing();
    basicDataProcessing();

    const dim3 blockSize(16, 16);                                                  // :122
    const dim3 gridSize((width + 15) / 16, (height + 15) / 16);                    // :1231
    std::printf("GridSize: %u, %u, %u\n", gridSize.x, gridSize.y, gridSize.z);
    unsigned int* d_output = nullptr;
    CK(cudaMalloc(reinterpret_cast<void**>(&d_output), (size_t)width * height * 4));   // :1021
    std::vector<uint32_t> frame((size_t)width * height);

    int status = EXIT_SUCCESS;
    for (int vidx = 0; vidx < views; ++vidx) {
        float m[12];
        vrdd_view_matrix(rotx, roty0 + (float)vidx * (360.0f / (float)views), 0.f, 0.f, -4.f, m);    // :126, 229-246
        copyInvViewMatrix(m, sizeof(float) * 12);                                  // :1046
        CK(cudaMemset(d_output, 0, (size_t)width * height * 4));                   // :1022
        // nIter timed launches after one warm-up, like :1049-1063
        std::chrono::steady_clock::time_point t0;
        for (int i = -1; i < iters; ++i) {
            if (i == 0) { CK(cudaDeviceSynchronize()); t0 = std::chrono::steady_clock::now(); }
            render_kernel(gridSize, blockSize, d_output, width, height, density, brightness, offset, scale, query, volumeSize);
        }
        CK(cudaDeviceSynchronize());
        const double avg = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / (iters > 0 ? iters : 1);
        std::printf("volumeRender, Throughput = %.4f MTexels/s, Time = %.5f s, Size = %u Texels, NumDevsUsed = %u, Workgroup = %u\n",
                    (1.0e-6 * width * height) / avg, avg, width * height, 1u, blockSize.x * blockSize.y);   // :1066
        CK(cudaGetLastError());                                                    // :1070
        CK(cudaMemcpy(frame.data(), d_output, (size_t)width * height * 4, cudaMemcpyDeviceToHost));   // :1074
        std::string name = std::string(out_prefix) + (views > 1 ? "_" + std::to_string(vidx) : "") + ".ppm";
        if (vrdd_io_write_ppm(name.c_str(), frame.data(), (int)width, (int)height) != VRDD_OK)
            std::fprintf(stderr, "cannot write %s\n", name.c_str());
        else
            std::printf("saved %s\n", name.c_str());
        if (ref_file && vidx == 0) {                                               // :1077, with a +-1 LSB rule
            std::vector<uint8_t> ref(3 * (size_t)width * height);
            if (vrdd_io_read_ppm(ref_file, ref.data(), (int)width, (int)height) != VRDD_OK) {
                std::fprintf(stderr, "cannot read reference image %s\n", ref_file);
                status = EXIT_FAILURE;
            } else {
                size_t bad = 0; int worst = 0;
                for (size_t p = 0; p < (size_t)width * height; ++p)
                    for (int ch = 0; ch < 3; ++ch) {
                        const int d = std::abs((int)((frame[p] >> (8 * ch)) & 255) - (int)ref[3 * p + ch]);
                        worst = d > worst ? d : worst;
                        bad += d > 1;
                    }
                std::printf("compare with %s: max |diff| = %d LSB, %zu channels differ by more than 1 -> %s\n", ref_file,
                            worst, bad, bad ? "FAILED" : "PASSED");
                if (bad) status = EXIT_FAILURE;
            }
        }
    }
    cudaFree(d_output);
    freeCudaBuffers();                                                             // cleanup(), :462
    return status;
}
