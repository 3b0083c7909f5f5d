// Micro-benchmark (measurement only, not part of the library): throughput of shared-memory atomics on sm_100a,
// the number that decides whether a linked-list inverse map (one ATOMS.EXCH per source pixel) can beat red.global.
#include <cstdio>
#include <cuda_runtime.h>
template <int kOp>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, int spread) {
  __shared__ unsigned head[4096];
  __shared__ float facc[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) { head[i] = 0; facc[i] = 0.f; }
  __syncthreads();
  unsigned acc = 0;
  unsigned idx = (threadIdx.x * spread) & 4095;
  for (int it = 0; it < iters; ++it) {
    if (kOp == 0) acc += atomicExch(&head[idx], it);             // ATOMS.EXCH with return
    if (kOp == 1) atomicAdd(&head[idx], 1u);                     // no return
    if (kOp == 2) acc += atomicAdd(&head[idx], 1u);              // with return
    if (kOp == 3) atomicAdd(&facc[idx], 1.0f);                   // float add (CAS loop?)
    if (kOp == 4) { unsigned v = head[idx]; head[idx] = v + it; } // plain LDS + STS
    if (kOp == 5) acc += atomicCAS(&head[idx], 0u, (unsigned)it);
    idx = (idx + 257 * spread) & 4095;
  }
  if (acc == 0xdeadbeef) out[0] = acc + (unsigned)facc[threadIdx.x];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = head[5] + (unsigned)facc[7];
}
template <int kOp>
void run(const char* name, int spread) {
  unsigned* out; cudaMalloc(&out, 64);
  int iters = 4096; int grid = 148 * 4;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<kOp><<<grid, 256>>>(out, 16, spread);
  cudaEventRecord(a); k<kOp><<<grid, 256>>>(out, iters, spread); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = (double)grid * 256 * iters;
  printf("%-28s spread=%d: %.3f ms, %.1f G lane-ops/s, %.2f SM-cycles/warp-instr @1.9GHz\n", name, spread, ms, ops / ms * 1e-6,
         ms * 1e-3 * 1.9e9 * 148 / (ops / 32));
  cudaFree(out);
}
int main() {
  for (int spread : {1, 33}) {
    run<0>("ATOMS.EXCH (ret)", spread); run<1>("atomicAdd u32 (no ret)", spread); run<2>("atomicAdd u32 (ret)", spread);
    run<3>("atomicAdd f32", spread); run<4>("LDS+STS", spread); run<5>("atomicCAS", spread);
  }
  return 0;
}
