// TEST INFRASTRUCTURE — lets a CUDA translation unit whose kernels use neither shared memory nor barriers be compiled
// with g++ and executed on the CPU, one emulated thread after the other, so that its LOGIC (indexing, tables, flux
// correction) can be checked against the golden fixtures on a box without a GPU.  Used only by
// tests/test_amr_ops_host_emulation.py for csrc/amr_ops.cu (a path that has not been run on hardware yet); never part of
// the product library, and not a CPU fallback of it.
#pragma once
#include "../../include/cup2d_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct EmuIdx { int x = 0; };
inline thread_local EmuIdx blockIdx, threadIdx, blockDim, gridDim;

template <class T> cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)malloc(n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }

// kernel<<<grid, block, smem, stream>>>(args)  ->  emu_launch(grid, block, [&] { kernel(args); })   (rewritten by build.py)
inline void emu_launch(int grid, int block, const std::function<void()> &body) {
  gridDim.x = grid, blockDim.x = block;
  for (int b = 0; b < grid; b++)
    for (int t = 0; t < block; t++) {
      blockIdx.x = b, threadIdx.x = t;
      body();
    }
}

using std::fabs;
using std::fmax;
struct cup2d_sim;

// what csrc/amr_ops.cu takes from sim.h
namespace cup2d {
void set_error(const std::string &msg);
int dim_of(int field);
}
#define CUP2D_CUDA(call)                                                                          \
  do {                                                                                            \
    if ((call) != cudaSuccess) {                                                                  \
      cup2d::set_error(#call);                                                                    \
      return CUP2D_ECUDA;                                                                         \
    }                                                                                             \
  } while (0)
