// TEST INFRASTRUCTURE — lets a CUDA translation unit whose kernels use neither shared memory nor barriers be compiled
// with g++ and executed on the CPU, one emulated thread after the other, so that its LOGIC (indexing, tables, flux
// correction) can be checked against the golden fixtures on a box without a GPU.  Used only by
// tests/test_amr_ops_host_emulation.py for csrc/amr_ops.cu (a path that has not been run on hardware yet); never part of
// the product library, and not a CPU fallback of it.
#pragma once
#include "../../include/cup2d_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__

typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
struct EmuIdx { int x = 0; };
inline thread_local EmuIdx blockIdx, threadIdx, blockDim, gridDim;

// device allocations: 256-byte aligned like cudaMalloc, exact size (so that AddressSanitizer sees overruns), and filled with
// 0xFF bytes (NaN doubles, -1 indices): cudaMalloc does not zero memory, a fresh malloc often does
template <class T> cudaError_t cudaMalloc(T **p, size_t n) {
  void *q = nullptr;
  if (posix_memalign(&q, 256, n ? n : 1)) return 2;
  memset(q, 0xFF, n ? n : 1);
  *p = (T *)q;
  return 0;
}
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
// cp.async.bulk: both addresses and the size must be multiples of 16 bytes
inline void emu_bulk_copy(void *dst, const void *src, size_t bytes) {
  if (((uintptr_t)dst | (uintptr_t)src | bytes) & 15) {
    fprintf(stderr, "emulation: misaligned bulk copy %p <- %p, %zu bytes\n", dst, src, bytes);
    abort();
  }
  memcpy(dst, src, bytes);
}
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
// Serial form: one emulated thread after the other — enough for kernels without shared memory, barriers or shuffles.
inline void emu_launch(int grid, int block, const std::function<void()> &body) {
  gridDim.x = grid, blockDim.x = block;
  for (int b = 0; b < grid; b++)
    for (int t = 0; t < block; t++) {
      blockIdx.x = b, threadIdx.x = t;
      body();
    }
}

// Cooperative form: the threads of a block are OS threads that really run concurrently and meet at __syncthreads() /
// __syncwarp() / warp shuffles; blocks run one after the other, so `__shared__` can simply be `static` storage.
// A thread that returns from the kernel drops out of the barriers (as an exited CUDA thread does).
#include <barrier>
#include <memory>
#include <thread>
#include <vector>
#include <map>
#include <mutex>
struct EmuBlock {
  std::mutex shm_lock;
  std::map<int, void *> shm; // block-local "shared memory": one buffer per __shared__ declaration site (build.py rewrites them)
  ~EmuBlock() { for (auto &e : shm) free(e.second); }
  std::barrier<> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<unsigned long long> slot; // shuffle exchange, 32 per warp
  explicit EmuBlock(int n) : block_bar(n), slot((size_t)((n + 31) / 32) * 32) {
    for (int w = 0; w < (n + 31) / 32; w++) warp_bar.emplace_back(new std::barrier<>(std::min(32, n - 32 * w)));
  }
};
inline thread_local EmuBlock *emu_block = nullptr;
inline void __syncthreads() { emu_block->block_bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_block->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
template <class T> T emu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  unsigned long long bits = 0;
  memcpy(&bits, &v, sizeof(T));
  emu_block->slot[w * 32 + l] = bits;
  __syncwarp();
  bits = emu_block->slot[w * 32 + (src_lane & 31)];
  __syncwarp();
  T r;
  memcpy(&r, &bits, sizeof(T));
  return r;
}
template <class T> T __shfl_xor_sync(unsigned, T v, int mask) { return emu_shfl(v, (threadIdx.x & 31) ^ mask); }
template <class T> T __shfl_sync(unsigned, T v, int lane) { return emu_shfl(v, lane); }
template <class T> T __shfl_down_sync(unsigned, T v, int d) { return emu_shfl(v, std::min(31, (int)(threadIdx.x & 31) + d)); }
// storage of the __shared__ variable declared at site `id`, common to the threads of the current block only (several
// emulated ranks may run kernels at the same time in one process)
inline void *emu_shared(int id, size_t bytes) {
  std::lock_guard<std::mutex> g(emu_block->shm_lock);
  void *&p = emu_block->shm[id];
  if (!p) { // exact size + garbage, like the real thing
    if (posix_memalign(&p, 128, bytes ? bytes : 1)) abort();
    memset(p, 0xFF, bytes ? bytes : 1);
  }
  return p;
}
inline void emu_launch_coop(int grid, int block, const std::function<void()> &body) {
  for (int b = 0; b < grid; b++) {
    EmuBlock blk(block);
    std::vector<std::thread> pool;
    for (int t = 0; t < block; t++)
      pool.emplace_back([&, t] {
        gridDim.x = grid, blockDim.x = block, blockIdx.x = b, threadIdx.x = t;
        emu_block = &blk;
        body();
        blk.block_bar.arrive_and_drop();
        blk.warp_bar[t >> 5]->arrive_and_drop();
      });
    for (auto &th : pool) th.join();
  }
}
// kernels whose threads never interact (build.py: serial_safe_kernels) run one thread after the other, except in the
// sanitizer builds, where every CUDA thread stays an OS thread so that ThreadSanitizer sees all of them
inline void emu_launch_auto(int grid, int block, const std::function<void()> &body) {
#ifdef EMU_ALL_COOP
  emu_launch_coop(grid, block, body);
#else
  emu_launch(grid, block, body);
#endif
}
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
// vector types carry the alignment the hardware demands of their loads and stores (-fsanitize=alignment reports a violation)
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
inline double2 make_double2(double x, double y) { return {x, y}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }

using std::fabs;
using std::fma;
using std::fmax;
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
template <class T, class U> cudaError_t cudaMemcpyToSymbol(T &sym, const U &src, size_t n) { memcpy(&sym, &src, n); return 0; }
#ifndef CUP2D_FULL_EMU
namespace cup2d {
inline bool is_pos(double x) { return !(x <= 0); } // U > 0, NaN counted as positive (common.cuh)
inline int __double2hiint(double x) { long long b; memcpy(&b, &x, 8); return (int)(b >> 32); } // (the full build gets it from cuda_runtime.h)
}
#endif
struct cup2d_sim;

#ifndef CUP2D_FULL_EMU
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
#endif
