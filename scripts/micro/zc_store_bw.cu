// Zero-copy (SM store -> pinned host memory) bandwidth vs store width, and SM loads from pinned host memory.
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void wr(T* out, size_t n, T v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}
template <typename T>
__global__ void rd(const T* in, size_t n, int* sink) {
  int acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    T v = in[i];
    acc += *reinterpret_cast<int*>(&v);
  }
  if (acc == 12345) *sink = acc;
}
template <typename F>
float timeit(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 5; ++i) f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps * 1e3f;
}
int main() {
  int* sink; cudaMalloc(&sink, 4);
  for (size_t bytes : {(size_t)589824, (size_t)2359296, (size_t)16 << 20}) {
    void* h; cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    void* d; cudaMalloc(&d, bytes);
    for (int grid : {148, 1024, 4096}) {
      for (int threads : {64, 256}) {
        float t4 = timeit([&] { wr<int><<<grid, threads>>>((int*)h, bytes / 4, 1); }, 20);
        float t8 = timeit([&] { wr<int2><<<grid, threads>>>((int2*)h, bytes / 8, make_int2(1, 2)); }, 20);
        float t16 = timeit([&] { wr<int4><<<grid, threads>>>((int4*)h, bytes / 16, make_int4(1, 2, 3, 4)); }, 20);
        float t1 = timeit([&] { wr<unsigned char><<<grid, threads>>>((unsigned char*)h, bytes / 4, 1); }, 20);  // a quarter of the bytes
        float r4 = timeit([&] { rd<int><<<grid, threads>>>((const int*)h, bytes / 4, sink); }, 20);
        float r16 = timeit([&] { rd<int4><<<grid, threads>>>((const int4*)h, bytes / 16, sink); }, 20);
        printf("bytes %8zu grid %4d thr %3d  store 4B %7.1f us %5.1f GB/s | 8B %7.1f us %5.1f | 16B %7.1f us %5.1f | 1B(x1/4) %7.1f us %5.1f | load 4B %7.1f us %5.1f | 16B %7.1f us %5.1f\n",
               bytes, grid, threads, t4, bytes / t4 / 1e3, t8, bytes / t8 / 1e3, t16, bytes / t16 / 1e3, t1, bytes / 4 / t1 / 1e3,
               r4, bytes / r4 / 1e3, r16, bytes / r16 / 1e3);
      }
    }
    float tc = timeit([&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, 0); }, 20);
    float th = timeit([&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0); }, 20);
    printf("bytes %8zu DMA d2h %7.1f us %5.1f GB/s | h2d %7.1f us %5.1f GB/s\n", bytes, tc, bytes / tc / 1e3, th, bytes / th / 1e3);
    cudaFreeHost(h); cudaFree(d);
  }
  return 0;
}
