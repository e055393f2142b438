// Micro-benchmark of grid-wide barriers for a cooperative (co-resident) launch on B200: how long is one barrier?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o grid_barrier grid_barrier.cu ; run: ./grid_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ldv(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

// A: one counter, threadfence + atomicAdd + volatile spin (what select.cu uses)
__device__ __forceinline__ void bar_flat(unsigned* c, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) { target += gridDim.x; __threadfence(); atomicAdd(c, 1u); while (ldv(c) < target) {} __threadfence(); }
  __syncthreads();
}
// B: release/acquire PTX instead of fences
__device__ __forceinline__ void bar_relacq(unsigned* c, unsigned& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory"); } while (v < target);
  }
  __syncthreads();
}
// C: hierarchical: 8 group counters (blockIdx % 8), last arriver of a group arrives at the root, root's last arriver bumps the epoch
__device__ __forceinline__ void bar_tree(unsigned* c, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    const unsigned g = blockIdx.x & 7u, gsize = (gridDim.x + 7u - g) / 8u;
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(c + 32 * (1 + g)) : "memory");
    if ((old + 1u) % gsize == 0u) {
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(c + 32 * 9) : "memory");
      if ((old + 1u) % 8u == 0u) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
    }
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory"); } while (v < epoch);
  }
  __syncthreads();
}

template <int MODE>
__global__ void k(unsigned* c, int iters, unsigned long long* out) {
  unsigned t = 0;
  cg::grid_group grid = cg::this_grid();
  unsigned long long t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) bar_flat(c, t);
    if (MODE == 1) bar_relacq(c, t);
    if (MODE == 2) bar_tree(c, t);
    if (MODE == 3) grid.sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); *out = t1 - t0; }
}

int main() {
  unsigned* c; unsigned long long* out; cudaMalloc(&c, 4096); cudaMalloc(&out, 8);
  const int iters = 2000;
  const char* names[4] = {"flat fence+atomic+volatile spin", "flat red.release / ld.acquire", "tree (8 groups)", "cooperative_groups grid.sync()"};
  for (int cfg = 0; cfg < 3; cfg++) {
    const int blocks = cfg == 0 ? 296 : 148, threads = cfg == 2 ? 1024 : 512;
    for (int mode = 0; mode < 4; mode++) {
      cudaMemset(c, 0, 4096);
      void* args[] = {&c, (void*)&iters, &out};
      void* fn = mode == 0 ? (void*)k<0> : mode == 1 ? (void*)k<1> : mode == 2 ? (void*)k<2> : (void*)k<3>;
      for (int rep = 0; rep < 2; rep++) { cudaMemset(c, 0, 4096); cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, 0, 0); cudaDeviceSynchronize(); }
      unsigned long long ns = 0; cudaMemcpy(&ns, out, 8, cudaMemcpyDeviceToHost);
      printf("%3d x %4d  %-36s %.3f us per barrier (%s)\n", blocks, threads, names[mode], ns / 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
