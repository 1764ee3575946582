// Micro-benchmarks that size the CTC lattice kernel's per-frame chain on B200 (sm_100a):
// fp64 add/fma latency+throughput, MUFU ex2, 64-bit shuffle, bar.sync, smem ld.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

__global__ void lat_dfma(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < n; i++) a = fma(a, b, b);
  long long t1 = clock64();
  out[2] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_dadd(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  #pragma unroll 16
  for (int i = 0; i < n; i++) a = a + b;
  long long t1 = clock64();
  out[2] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: 8 independent chains per thread, many warps
__global__ void thr_dfma(double* out, long long* cyc, int n) {
  double a0=out[0],a1=out[1],a2=out[2],a3=out[3],a4=out[4],a5=out[5],a6=out[6],a7=out[7], b=out[8];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    a0=fma(a0,b,b);a1=fma(a1,b,b);a2=fma(a2,b,b);a3=fma(a3,b,b);
    a4=fma(a4,b,b);a5=fma(a5,b,b);a6=fma(a6,b,b);a7=fma(a7,b,b);
  }
  __syncthreads();
  long long t1 = clock64();
  out[9+threadIdx.x%2] = a0+a1+a2+a3+a4+a5+a6+a7; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void thr_ffma(float* out, long long* cyc, int n) {
  float a0=out[0],a1=out[1],a2=out[2],a3=out[3],a4=out[4],a5=out[5],a6=out[6],a7=out[7], b=out[8];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    a0=fmaf(a0,b,b);a1=fmaf(a1,b,b);a2=fmaf(a2,b,b);a3=fmaf(a3,b,b);
    a4=fmaf(a4,b,b);a5=fmaf(a5,b,b);a6=fmaf(a6,b,b);a7=fmaf(a7,b,b);
  }
  __syncthreads();
  long long t1 = clock64();
  out[9+threadIdx.x%2] = a0+a1+a2+a3+a4+a5+a6+a7; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void thr_ex2(float* out, long long* cyc, int n) {
  float a0=out[0],a1=out[1],a2=out[2],a3=out[3];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a0=exp2f(a0);a1=exp2f(a1);a2=exp2f(a2);a3=exp2f(a3); }
  __syncthreads();
  long long t1 = clock64();
  out[9+threadIdx.x%2] = a0+a1+a2+a3; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void thr_f2d(float* in, double* out, long long* cyc, int n) {
  float a0=in[0],a1=in[1],a2=in[2],a3=in[3]; double s0=0,s1=0,s2=0,s3=0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { s0=(double)a0; a0=(float)s0+1.f; s1=(double)a1; a1=(float)s1+1.f; s2=(double)a2; a2=(float)s2+1.f; s3=(double)a3; a3=(float)s3+1.f;}
  __syncthreads();
  long long t1 = clock64();
  out[threadIdx.x%2] = s0+s1+s2+s3; if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_shfl64(double* out, long long* cyc, int n) {
  double a = out[threadIdx.x & 31];
  long long t0 = clock64();
  #pragma unroll 8
  for (int i = 0; i < n; i++) a = __shfl_up_sync(0xffffffffu, a, 1);
  long long t1 = clock64();
  out[32 + threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_shfl32(int* out, long long* cyc, int n) {
  int a = out[threadIdx.x & 31];
  long long t0 = clock64();
  #pragma unroll 8
  for (int i = 0; i < n; i++) a = __shfl_up_sync(0xffffffffu, a, 1) + 1;
  long long t1 = clock64();
  out[32 + threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_bar(long long* cyc, int n) {
  long long t0 = clock64();
  for (int i = 0; i < n; i++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_lds(int* out, long long* cyc, int n) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  int a = threadIdx.x;
  long long t0 = clock64();
  #pragma unroll 8
  for (int i = 0; i < n; i++) a = s[a];
  long long t1 = clock64();
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// chain step like the lattice: shfl64 -> dmul -> dfma -> dmul
__global__ void lat_chain(double* out, long long* cyc, int n) {
  double a = out[threadIdx.x & 31], p = out[40];
  long long t0 = clock64();
  #pragma unroll 4
  for (int i = 0; i < n; i++) { double b = __shfl_up_sync(0xffffffffu, a, 1); b = b * p; double y = fma(a, p, b); a = y * p; }
  long long t1 = clock64();
  out[64 + threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ void sts_v4(volatile void* p, int a, int b, int c, int d) {
  asm volatile("st.volatile.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"((unsigned)__cvta_generic_to_shared((const void*)p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ int4 lds_v4(volatile void* p) {
  int4 r;
  asm volatile("ld.volatile.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"((unsigned)__cvta_generic_to_shared((const void*)p)) : "memory");
  return r;
}
// two warps ping-pong through a tagged 16-byte shared-memory slot (volatile polling): one-way hop latency
__global__ void lat_pingpong(double* out, long long* cyc, int n) {
  __shared__ __align__(16) volatile int slot[2][4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 8) ((volatile int*)slot)[threadIdx.x] = -1;
  __syncthreads();
  long long t0 = clock64();
  double acc = out[0];
  for (int i = 0; i < n; i++) {
    if (w == 0) {
      if (lane == 0) sts_v4(&slot[0][0], __double2loint(acc), __double2hiint(acc), 0, i);
      int tag; int4 r;
      do { r = lds_v4(&slot[1][0]); tag = r.w; } while (tag != i);
      acc = __hiloint2double(r.y, r.x) + 1.0;
    } else {
      int tag; int4 r;
      do { r = lds_v4(&slot[0][0]); tag = r.w; } while (tag != i);
      acc = __hiloint2double(r.y, r.x) + 1.0;
      if (lane == 0) sts_v4(&slot[1][0], __double2loint(acc), __double2hiint(acc), 0, i);
    }
  }
  long long t1 = clock64();
  out[64 + threadIdx.x] = acc; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_f2f(double* out, long long* cyc, int n) {
  double a = out[threadIdx.x & 31];
  long long t0 = clock64();
  #pragma unroll 8
  for (int i = 0; i < n; i++) { float f = (float)a; a = (double)f; }
  long long t1 = clock64();
  out[64 + threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// named barrier between 2 / 4 warps
__global__ void lat_namedbar(long long* cyc, int n, int nthr) {
  long long t0 = clock64();
  if (threadIdx.x < nthr) for (int i = 0; i < n; i++) asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d; long long* c; CK(cudaMalloc(&d, 1 << 16)); CK(cudaMalloc(&c, 1 << 16)); CK(cudaMemset(d, 0, 1 << 16));
  float* f = (float*)d; int* ii = (int*)d;
  long long h[1024]; int n = 4096;
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); printf("%s SMs=%d clock=%d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
  for (int rep = 0; rep < 2; rep++) {
    lat_dfma<<<1,32>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("DFMA latency       %.2f cyc\n", (double)h[0]/n);
    lat_dadd<<<1,32>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("DADD latency       %.2f cyc\n", (double)h[0]/n);
    for (int nt = 128; nt <= 1024; nt *= 2) {
      thr_dfma<<<1,nt>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost));
      printf("DFMA thr %4d thr/SM: %.2f DFMA/clk/SM\n", nt, (double)nt*8*n/h[0]);
    }
    thr_ffma<<<1,1024>>>(f, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("FFMA thr 1024: %.2f /clk/SM\n", 1024.0*8*n/h[0]);
    thr_ex2<<<1,1024>>>(f, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("EX2 thr 1024: %.2f /clk/SM\n", 1024.0*4*n/h[0]);
    thr_f2d<<<1,1024>>>(f, d+64, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("F2D+D2F pair thr 1024: %.2f pairs/clk/SM\n", 1024.0*4*n/h[0]);
    lat_shfl64<<<1,32>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("SHFL64 latency     %.2f cyc\n", (double)h[0]/n);
    lat_shfl32<<<1,32>>>(ii, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("SHFL32+IADD lat    %.2f cyc\n", (double)h[0]/n);
    for (int nt = 32; nt <= 1024; nt *= 2) { lat_bar<<<1,nt>>>(c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("BAR.SYNC %4d thr   %.2f cyc\n", nt, (double)h[0]/n); }
    lat_lds<<<1,32>>>(ii, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("LDS latency        %.2f cyc\n", (double)h[0]/n);
    lat_pingpong<<<1,64>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("smem flag ping-pong round trip %.2f cyc (one-way hop = half)\n", (double)h[0]/n);
    lat_f2f<<<1,32>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("F2F d->f->d pair latency %.2f cyc\n", (double)h[0]/n);
    for (int nt = 64; nt <= 256; nt *= 2) { lat_namedbar<<<1,256>>>(c, n, nt); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("named bar.sync %3d thr %.2f cyc\n", nt, (double)h[0]/n); }
    lat_chain<<<1,32>>>(d, c, n); CK(cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost)); printf("chain shfl64+dmul+dfma+dmul %.2f cyc\n", (double)h[0]/n);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
