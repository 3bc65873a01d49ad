// Lab micro-benchmarks (not part of the product): tensor-memory read/write bandwidth and MUFU.EX2 throughput per SM
// on sm_100a, the two limits of the attention softmax role (csrc/attention_tcgen05.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(addr, v)                                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"        \
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"                            \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),        \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),  \
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),             \
                 "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),             \
                 "=r"(v[30]), "=r"(v[31])                                                                               \
               : "r"(addr)                                                                                              \
               : "memory")
#define ST32(addr, v)                                                                                                  \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(addr),                       \
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),       \
               "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),           \
               "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),          \
               "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])                       \
               : "memory")

// mode 0: tcgen05.ld only; 1: tcgen05.st only; 2: MUFU.EX2 fp32 only; 3: ld + 32 ex2 per chunk, load of the next chunk
// issued first (the attention pattern); 4: ex2.approx.f16x2 only; 5: ld + ex2 + st (apply pattern)
__global__ void __launch_bounds__(512, 1) k(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t a[32], b[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { a[j] = __float_as_uint(0.001f * (threadIdx.x + j)); b[j] = a[j]; }
  // initialise the columns this warp will read
  for (int c = 0; c < 512; c += 32) ST32(base + c, a);
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
  if (mode == 0) {
    for (int i = 0; i < iters; i += 2) {
      LD32(base + ((i * 32) & 480), a);
      LD32(base + ((i * 32 + 32) & 480), b);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      acc += __uint_as_float(a[0]) + __uint_as_float(b[31]);
    }
  } else if (mode == 1) {
    for (int i = 0; i < iters; i += 2) {
      ST32(base + ((i * 32) & 480), a);
      ST32(base + ((i * 32 + 32) & 480), b);
      asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
  } else if (mode == 2) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(__uint_as_float(a[j])));
        a[j] = __float_as_uint(y * 0.5f);
      }
    }
    acc = __uint_as_float(a[3]);
  } else if (mode == 4) {
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {  // 32 packed ops = 64 exponentials
        uint32_t y;
        asm volatile("ex2.approx.f16x2 %0, %1;\n" : "=r"(y) : "r"(a[j]));
        a[j] = y ^ 0x04000400u;
      }
    }
    acc = __uint_as_float(a[3]);
  } else {
    LD32(base, a);
    for (int i = 0; i < iters; i += 2) {
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      LD32(base + ((i * 32 + 32) & 480), b);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(__uint_as_float(a[j]) * 0.25f));
        a[j] = __float_as_uint(y);
      }
      if (mode == 5) ST32(base + ((i * 32) & 480), a);
      acc += __uint_as_float(a[5]);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      LD32(base + ((i * 32 + 64) & 480), a);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(__uint_as_float(b[j]) * 0.25f));
        b[j] = __float_as_uint(y);
      }
      if (mode == 5) ST32(base + ((i * 32 + 32) & 480), b);
      acc += __uint_as_float(b[7]);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  const long long t1 = clock64();
  if (acc == 1.2345e-33f) sink[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  long long* out;
  float* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, 4096);
  const char* names[] = {"tcgen05.ld 32x32b.x32", "tcgen05.st 32x32b.x32", "ex2.approx.ftz.f32", "ld(next) + 32 ex2 (stats pattern)",
                         "ex2.approx.f16x2 (2 exps per op)", "ld(next) + 32 ex2 + st (apply pattern)"};
  const int iters = 4096;
  for (int mode = 0; mode < 6; ++mode)
    for (int warps : {4, 8, 16}) {
      k<<<148, warps * 32>>>(mode, iters, out, sink);
      k<<<148, warps * 32>>>(mode, iters, out, sink);
      long long clk = 0;
      cudaError_t e = cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      const double chunks = (double)iters * warps;            // 32 lanes x 32 columns x 4 B each
      const double per_chunk = (double)clk / iters;
      if (mode == 2 || mode == 4)
        printf("%-42s %2d warps: %8.1f clk per 32 ops per warp, %6.2f exps/clk/SM\n", names[mode], warps, per_chunk,
               (mode == 4 ? 2.0 : 1.0) * 32.0 * 32.0 * chunks / clk);
      else
        printf("%-42s %2d warps: %8.1f clk per chunk per warp, %7.1f B/clk/SM tmem, %6.2f exps/clk/SM\n", names[mode], warps,
               per_chunk, 4096.0 * chunks / clk, mode >= 3 ? 1024.0 * chunks / clk : 0.0);
    }
  return 0;
}
