// Pipe-rate microbenchmarks that size the DSTD-GC kernel design on B200 (sm_100a): 3-register FFMA, mma.sync TF32 /
// BF16 (legacy tensor path), the ex2+rcp tanh used by the dynamic adjacency, and shared-memory float4 reads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
  float x = a + threadIdx.x * 1e-6f, y = b;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ float fast_tanh(float x) {
  float e = exp2f(x * 2.885390081777927f);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

__global__ void __launch_bounds__(256) k_tanh(float* out, float a) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 0.001f + i * a;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = fast_tanh(r[i] + a);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_tanhf(float* out, float a) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 0.001f + i * a;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = tanhf(r[i] + a);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_tanh_approx(float* out, float a) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 0.001f + i * a;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t;
      asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(r[i] + a));
      r[i] = t;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4 independent accumulator tiles per warp
__global__ void __launch_bounds__(256) k_mma_tf32(float* out, uint32_t a0) {
  float c[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {a0 + 4, a0 + 5};
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mma_bf16(float* out, uint32_t a0) {
  float c[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  uint32_t a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {a0 + 4, a0 + 5};
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory read bandwidth: conflict-free float4 reads
__global__ void __launch_bounds__(256) k_lds(float* out) {
  __shared__ float4 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float4(i, i, i, i);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = sm[(idx + j * 256) & 2047];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx = (idx + 1) & 2047;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  float* out;
  CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
  const int blocks = sms * 8;   // 8 CTAs x 256 thr = 64 warps / SM
  float ms;
  ms = time_ms([&] { k_ffma<<<blocks, 256>>>(out, 1.0001f, 0.5f); });
  double fl = 2.0 * blocks * 256.0 * ITERS * 16;
  printf("FFMA 3-reg       : %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM @1.9GHz)\n", ms, fl / ms * 1e-9, fl / 2 / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_tanh<<<blocks, 256>>>(out, 0.01f); });
  double nt = (double)blocks * 256.0 * ITERS * 8;
  printf("tanh (ex2+rcp)   : %8.3f ms  %7.2f Gtanh/s   (%.2f /clk/SM @1.9GHz)\n", ms, nt / ms * 1e-6, nt / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_tanhf<<<blocks, 256>>>(out, 0.01f); });
  printf("tanhf (libdevice): %8.3f ms  %7.2f Gtanh/s   (%.2f /clk/SM @1.9GHz)\n", ms, nt / ms * 1e-6, nt / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_tanh_approx<<<blocks, 256>>>(out, 0.01f); });
  printf("tanh.approx      : %8.3f ms  %7.2f Gtanh/s   (%.2f /clk/SM @1.9GHz)\n", ms, nt / ms * 1e-6, nt / (ms * 1e-3) / sms / 1.9e9);
  ms = time_ms([&] { k_mma_tf32<<<blocks, 256>>>(out, 0x3f800000u); });
  double mf = 2.0 * 16 * 8 * 8 * 4.0 * ITERS * blocks * 8;
  printf("mma.sync tf32    : %8.3f ms  %7.2f TFLOP/s\n", ms, mf / ms * 1e-9);
  ms = time_ms([&] { k_mma_bf16<<<blocks, 256>>>(out, 0x3f803f80u); });
  mf = 2.0 * 16 * 8 * 16 * 4.0 * ITERS * blocks * 8;
  printf("mma.sync bf16    : %8.3f ms  %7.2f TFLOP/s\n", ms, mf / ms * 1e-9);
  ms = time_ms([&] { k_lds<<<blocks, 256>>>(out); });
  double by = (double)blocks * 256.0 * ITERS * 8 * 16;
  printf("LDS.128          : %8.3f ms  %7.2f TB/s  (%.1f B/clk/SM @1.9GHz)\n", ms, by / ms * 1e-9, by / (ms * 1e-3) / sms / 1.9e9);
  cudaFree(out);
  return 0;
}
