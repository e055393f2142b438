// Throughput of warp-collective primitives on sm_100a (per SM, all 4 sub-partitions busy): cycles per warp-instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_prims warp_prims.cu ; run: ./warp_prims
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* cyc) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t v = seed + (OP == 1 ? 0u : lane * 2654435761u), acc = 0;
  unsigned long long v64 = ((unsigned long long)v << 32) | (lane * 40503u + seed);
  __shared__ uint32_t sm[1024];
  sm[threadIdx.x] = v;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < ITERS; i++) {
    if (OP == 0) acc += __match_any_sync(0xffffffffu, v + i);                       // 32 distinct values
    if (OP == 1) acc += __match_any_sync(0xffffffffu, v + i);                       // 1 distinct value
    if (OP == 2) acc += (uint32_t)__match_any_sync(0xffffffffu, v64 + i);           // 64-bit, distinct
    if (OP == 3) acc += __reduce_or_sync(0xffffffffu, v + i);
    if (OP == 4) acc += __ballot_sync(0xffffffffu, ((v + i) >> 3) & 1u);
    if (OP == 5) acc += __shfl_xor_sync(0xffffffffu, v + i, 5);
    if (OP == 6) { __syncwarp(); acc += v + i; }
    if (OP == 7) acc += __popc(v + i) + __brev(v ^ i);
    if (OP == 8) { acc += sm[(threadIdx.x * 33 + i) & 1023]; }
    if (OP == 9) acc += __match_any_sync(0xffffffffu, (v + i) & 3u);                // 4 distinct values
    if (OP == 10) acc += __any_sync(0xffffffffu, ((v + i) & 0xff) == 0u);
    if (OP == 11) acc += atomicAdd(&sm[(threadIdx.x * 33 + i) & 1023], 1u);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, uint32_t* out, long long* cyc) {
  for (int warps : {4, 16, 32}) {   // warps per SM (one block per SM)
    k<OP><<<148, warps * 32>>>(out, 7, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<148, warps * 32>>>(out, 7, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/SM %2d: %8.2f cycles per warp-instr per SM (block0 clock %lld, %.3f ms)\n", name, warps, (double)c / ((double)ITERS * warps), c, ms);
  }
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<0>("match_any b32 32 distinct", out, cyc);
  run<9>("match_any b32 4 distinct", out, cyc);
  run<1>("match_any b32 1 distinct", out, cyc);
  run<2>("match_any b64 32 distinct", out, cyc);
  run<3>("reduce_or (REDUX)", out, cyc);
  run<4>("ballot (VOTE)", out, cyc);
  run<10>("any_sync (VOTE.ANY)", out, cyc);
  run<5>("shfl_xor", out, cyc);
  run<6>("syncwarp", out, cyc);
  run<7>("popc+brev", out, cyc);
  run<8>("lds.32 (conflict-free)", out, cyc);
  run<11>("atoms.add.32", out, cyc);
  return 0;
}
