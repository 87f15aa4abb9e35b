// ref_texture_shim.h — force-included (-include) in front of the reference's device file when it is compiled for
// oracle/_ref.  CUDA 12 removed texture REFERENCES (`texture<T, dim, mode> name;`, cudaBindTextureToArray,
// tex3D(texref, ...)), which is all that keeps /root/reference/volumeRender_kernel.cu from compiling.  The shim
// gives those spellings back on top of texture OBJECTS, so the reference's own arithmetic runs, unmodified, with
// the real texture unit doing the filtering:
//   * `texture` becomes a managed aggregate with the fields the reference sets (normalized, filterMode,
//     addressMode[]) plus the texture object created when the reference binds an array to it;
//   * tex1D / tex2DLayered / tex3D overloads on that aggregate forward to the texture-object fetches;
//   * a texture reference's address mode defaults to clamp (the reference never asks for anything else).
// Test infrastructure only (oracle/): nothing in the product includes this.
#pragma once
#include <cuda_runtime.h>
#include <cstring>

template <class T, int Dim, cudaTextureReadMode Mode>
struct ref_texture {
    cudaTextureObject_t obj;
    int normalized;
    cudaTextureFilterMode filterMode;
    cudaTextureAddressMode addressMode[3];
};

template <class T, int Dim, cudaTextureReadMode Mode>
cudaError_t cudaBindTextureToArray(ref_texture<T, Dim, Mode>& t, cudaArray* array, const cudaChannelFormatDesc&) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    if (t.obj) cudaDestroyTextureObject(t.obj);
    cudaResourceDesc rd;
    std::memset(&rd, 0, sizeof(rd));
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = array;
    cudaTextureDesc td;
    std::memset(&td, 0, sizeof(td));
    for (int i = 0; i < 3; ++i) td.addressMode[i] = cudaAddressModeClamp;
    td.filterMode = t.filterMode;
    td.readMode = Mode;
    td.normalizedCoords = t.normalized ? 1 : 0;
    cudaTextureObject_t obj = 0;
    e = cudaCreateTextureObject(&obj, &rd, &td, nullptr);
    t.obj = obj;
    return e;
}

template <class T, cudaTextureReadMode Mode>
__device__ __forceinline__ T tex1D(const ref_texture<T, 1, Mode>& t, float x) { return tex1D<T>(t.obj, x); }
template <class T, cudaTextureReadMode Mode>
__device__ __forceinline__ T tex3D(const ref_texture<T, cudaTextureType3D, Mode>& t, float x, float y, float z) {
    return tex3D<T>(t.obj, x, y, z);
}
template <class T, cudaTextureReadMode Mode>
__device__ __forceinline__ T tex2DLayered(const ref_texture<T, cudaTextureType2DLayered, Mode>& t, float x, float y, int layer) {
    return tex2DLayered<T>(t.obj, x, y, layer);
}

// every `texture<...> name;` of the reference is a namespace-scope definition
#define texture __device__ __managed__ ref_texture
