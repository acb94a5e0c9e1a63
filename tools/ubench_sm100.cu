// Micro-benchmarks of the per-SM rates the attention / epilogue warps are bound by on sm_100a (tools only, not the product):
//   tcgen05.ld throughput (bytes / clk / SM) for 4, 8, 16 warps and 1 or 2 loads in flight per warp,
//   MUFU.EX2 throughput (lanes / clk / SM), cvt.rn.bf16x2.f32 throughput, and tcgen05.ld under a concurrent MUFU load.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_sm100 tools/ubench_sm100.cu ; run: tools/ubench_sm100
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: one load in flight; 1: two loads in flight; 2: one load + 32 ex2 per load (the softmax pattern); 3: ex2 only;
// 4: cvt.rn.bf16x2 only; 5: one load + 32 fmax per load (the row-max pattern)
template <int MODE>
__global__ void __launch_bounds__(512, 1) ubench(int iters, long long* clk_out, float* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tslot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tslot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  uint32_t xa = 0;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = -0.01f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0 || MODE == 5) {
    uint32_t r[32];
    for (int i = 0; i < iters; ++i) {
      tmem_ld32(base + ((i * 32) & 511), r);
      tmem_wait_ld();
      if (MODE == 5) {
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r[j]));
      } else {
        xa ^= r[0] ^ r[31];
      }
    }
  } else if (MODE == 1) {
    uint32_t r[32], s[32];
    for (int i = 0; i < iters; i += 2) {
      tmem_ld32(base + ((i * 32) & 511), r);
      tmem_ld32(base + ((i * 32 + 32) & 511), s);
      tmem_wait_ld();
      xa ^= r[0] ^ s[31];
    }
  } else if (MODE == 2) {
    uint32_t r[32], s[32];
    tmem_ld32(base, r);
    for (int i = 0; i < iters; i += 2) {
      tmem_wait_ld();
      tmem_ld32(base + ((i * 32 + 32) & 511), s);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(r[j]) * 1e-30f - 1.0f));
        acc += e;
      }
      tmem_wait_ld();
      tmem_ld32(base + ((i * 32 + 64) & 511), r);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__uint_as_float(s[j]) * 1e-30f - 1.0f));
        acc += e;
      }
    }
    tmem_wait_ld();
  } else if (MODE == 3) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[j]));
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += f[j];
  } else if (MODE == 4) {
    uint32_t p[16];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[j]) : "f"(f[2 * j]), "f"(f[2 * j + 1]));
        f[2 * j] += __uint_as_float(p[j] << 16);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += f[j];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f || xa == 0xdeadbeefu) sink[threadIdx.x] = acc + xa;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(512) : "memory");
}

template <int MODE>
static void run(const char* what, int warps, int iters, double units_per_warp_iter, const char* unit) {
  long long* d_clk;
  float* d_sink;
  cudaMalloc(&d_clk, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 512 * sizeof(float));
  ubench<MODE><<<148, warps * 32>>>(iters, d_clk, d_sink);
  ubench<MODE><<<148, warps * 32>>>(iters, d_clk, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: %s\n", what, cudaGetErrorString(e));
    return;
  }
  long long h[148];
  cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-58s warps=%2d  %9.0f clk  ->  %8.2f %s / clk / SM\n", what, warps, avg, units_per_warp_iter * iters * warps / avg, unit);
  cudaFree(d_clk);
  cudaFree(d_sink);
}

int main() {
  const int it = 4096;
  for (int w : {4, 8, 16}) run<0>("tcgen05.ld 32x32b.x32, one in flight", w, it, 4096.0, "B");
  for (int w : {4, 8, 16}) run<1>("tcgen05.ld 32x32b.x32, two in flight", w, it, 4096.0, "B");
  for (int w : {4, 8, 16}) run<5>("tcgen05.ld x32 + 32 fmax per load (row max)", w, it, 4096.0, "B");
  for (int w : {4, 8, 16}) run<2>("tcgen05.ld x32 prefetched + 32 ex2 per load (softmax)", w, it, 32.0 * 32.0, "ex2");
  for (int w : {4, 8, 16}) run<3>("ex2.approx.ftz.f32 only", w, it, 32.0 * 32.0, "ex2");
  for (int w : {4, 8, 16}) run<4>("cvt.rn.bf16x2.f32 only", w, it, 16.0 * 32.0, "cvt");
  return 0;
}
