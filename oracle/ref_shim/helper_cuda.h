// helper_cuda.h — stand-in for the NVIDIA CUDA Samples header of that name, which the reference includes
// (volumeRender_kernel.cu:16) but does not vendor.  Only what the reference's device file uses.  Test
// infrastructure (oracle/): lets the reference's own source be compiled where it lies, see oracle/Makefile.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define checkCudaErrors(call)                                                                             \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            std::fprintf(stderr, "CUDA error at %s:%d: %s (%s)\n", __FILE__, __LINE__, cudaGetErrorName(e__), \
                         cudaGetErrorString(e__));                                                        \
            std::exit(EXIT_FAILURE);                                                                      \
        }                                                                                                 \
    } while (0)

#define getLastCudaError(msg)                                                                             \
    do {                                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                                             \
        if (e__ != cudaSuccess) {                                                                         \
            std::fprintf(stderr, "%s: CUDA error at %s:%d: %s\n", (msg), __FILE__, __LINE__,              \
                         cudaGetErrorString(e__));                                                        \
            std::exit(EXIT_FAILURE);                                                                      \
        }                                                                                                 \
    } while (0)
