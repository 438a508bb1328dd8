// HBM copy-kernel probe: which access shape reaches the copy peak on B200?  Development aid.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bw_probe bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct W8 { uint32_t w[8]; };
__device__ __forceinline__ W8 ldg256(const void* p) {
  W8 r;
  asm volatile("ld.global.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(void* p, const W8& r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r.w[0]), "r"(r.w[1]), "r"(r.w[2]),
               "r"(r.w[3]), "r"(r.w[4]), "r"(r.w[5]), "r"(r.w[6]), "r"(r.w[7]) : "memory");
}
template <int U>
__global__ void __launch_bounds__(256) copy256(const uint8_t* in, uint8_t* out, int64_t chunks) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t c0 = tid; c0 < chunks; c0 += U * nthreads) {
    W8 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (c0 + u * nthreads < chunks) v[u] = ldg256(in + (c0 + u * nthreads) * 32);
#pragma unroll
    for (int u = 0; u < U; ++u) if (c0 + u * nthreads < chunks) { v[u].w[0] ^= 1; stg256(out + (c0 + u * nthreads) * 32, v[u]); }
  }
}
template <int U>
__global__ void __launch_bounds__(256) copy128(const uint4* in, uint4* out, int64_t chunks) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t c0 = tid; c0 < chunks; c0 += U * nthreads) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (c0 + u * nthreads < chunks) v[u] = in[c0 + u * nthreads];
#pragma unroll
    for (int u = 0; u < U; ++u) if (c0 + u * nthreads < chunks) { v[u].x ^= 1; out[c0 + u * nthreads] = v[u]; }
  }
}
// non-persistent: one chunk per thread
__global__ void __launch_bounds__(256) copy256_flat(const uint8_t* in, uint8_t* out, int64_t chunks) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < chunks) { W8 v = ldg256(in + c * 32); v.w[0] ^= 1; stg256(out + c * 32, v); }
}
__global__ void __launch_bounds__(256) copy128_flat(const uint4* in, uint4* out, int64_t chunks) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < chunks) { uint4 v = in[c]; v.x ^= 1; out[c] = v; }
}

template <typename F>
float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main() {
  const int64_t bytes = 402653184;  // 403 MB like the fake-quant probe
  uint8_t *in, *out; cudaMalloc(&in, bytes); cudaMalloc(&out, bytes); cudaMemset(in, 1, bytes);
  auto rep = [&](const char* n, float ms) { printf("%-28s %.4f ms  %.0f GB/s\n", n, ms, 2.0 * bytes / ms / 1e6); };
  rep("cudaMemcpy d2d", timeit([&] { cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice); }));
  for (int mult : {2, 4, 6, 8, 16}) {
    char nm[64];
    snprintf(nm, 64, "copy256<U=1> grid=148x%d", mult); rep(nm, timeit([&] { copy256<1><<<148 * mult, 256>>>(in, out, bytes / 32); }));
    snprintf(nm, 64, "copy256<U=2> grid=148x%d", mult); rep(nm, timeit([&] { copy256<2><<<148 * mult, 256>>>(in, out, bytes / 32); }));
    snprintf(nm, 64, "copy256<U=4> grid=148x%d", mult); rep(nm, timeit([&] { copy256<4><<<148 * mult, 256>>>(in, out, bytes / 32); }));
    snprintf(nm, 64, "copy128<U=2> grid=148x%d", mult); rep(nm, timeit([&] { copy128<2><<<148 * mult, 256>>>((uint4*)in, (uint4*)out, bytes / 16); }));
    snprintf(nm, 64, "copy128<U=4> grid=148x%d", mult); rep(nm, timeit([&] { copy128<4><<<148 * mult, 256>>>((uint4*)in, (uint4*)out, bytes / 16); }));
    snprintf(nm, 64, "copy128<U=8> grid=148x%d", mult); rep(nm, timeit([&] { copy128<8><<<148 * mult, 256>>>((uint4*)in, (uint4*)out, bytes / 16); }));
  }
  rep("copy256_flat", timeit([&] { copy256_flat<<<(unsigned)((bytes / 32 + 255) / 256), 256>>>(in, out, bytes / 32); }));
  rep("copy128_flat", timeit([&] { copy128_flat<<<(unsigned)((bytes / 16 + 255) / 256), 256>>>((uint4*)in, (uint4*)out, bytes / 16); }));
  return 0;
}
