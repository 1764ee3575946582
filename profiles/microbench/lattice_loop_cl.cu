// Micro-benchmark of the CTC lattice warp's frame loop on B200 (sm_100a): what does ONE warp pay per frame for the
// recurrence alone, and for each thing added around it (emission loads, val-ring stores, exponent snapshot, publish)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o lattice_loop lattice_loop.cu ; run: ./lattice_loop
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kNegExp = -(1 << 28);
__device__ __forceinline__ double pow2i(int d) {
  d = min(d, 1023);
  const double r = __hiloint2double((d + 1023) << 20, 0);
  return d < -1022 ? 0.0 : r;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_arrive_if(unsigned long long* bar, int on) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 st;\n\tsetp.ne.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(s32(bar)), "r"(on) : "memory");
}

// FLAGS bits: 1 LDS emissions, 2 val-ring STS, 4 snapshot+apply, 8 publish (mbarrier arrive per frame), 16 poll words per 8 frames
template <int NBU, int FLAGS, int UNROLL, int SPIN = 0, int RT = 0>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(128) lat(long long* cyc, double* sink, int T, int V, unsigned spin_mask = 0, int es_rt = 33, int vframe_rt = 0, int rm = 127, int rvm = 15) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* E = reinterpret_cast<double*>(smem);                      // [128][33]
  unsigned char* val = smem + 128 * 33 * 8;                         // [16][NBU*640 + 512]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(val + 16 * (NBU * 640 + 512));
  volatile int* words = reinterpret_cast<volatile int*>(bars + 32);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < 128 * 33; k += blockDim.x) E[k] = 0.3 + 0.001 * (k % 97);
  if (threadIdx.x < 32) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[threadIdx.x])));
  if (threadIdx.x < 8) words[threadIdx.x] = 1 << 30;
  __syncthreads();
  if (w != 0) {
    if (SPIN == 0 || !((spin_mask >> w) & 1)) return;
    if (SPIN == 1) {   // wait (mbarrier try_wait loop) until the lattice warp is done
      uint32_t done = 0;
      while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(s32(&bars[31])), "r"(0) : "memory");
    } else if (SPIN >= 4) {   // a warp that streams through SPIN KB of straight-line code per iteration (instruction-cache pressure)
      float a = (float)threadIdx.x, b2 = 1.0001f;
      while (words[7] != 12345) {
#pragma unroll
        for (int q = 0; q < SPIN * 64; q++) a = fmaf(a, b2, 0.5f);
      }
      sink[4096 + threadIdx.x] = a;
    } else if (SPIN == 2) {
      while (words[7] != 12345) __nanosleep(32);
    } else {
      while (words[7] != 12345) {}
    }
    return;
  }
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int P = 512 + 16 * ((8 + NBU - 1) / NBU), NBP = (NBU + 3) & ~3;
  const int vframe = RT ? vframe_rt : NBU * 640 + 512;
  const int es = RT ? es_rt : 33;
  const int Rm = RT ? rm : 127, RVm = RT ? rvm : 15;
  int ecol[NBU][2];
  unsigned skipm = 0x55555555u * (lane & 1) | 0x2;
#pragma unroll
  for (int j = 0; j < NBU; j++) { ecol[j][0] = ((lane * 7 + j * 3) % V) * 8; ecol[j][1] = ((lane * 5 + j * 11 + 1) % V) * 8; }
  double x[NBU][4]; int e[NBU], en_next[NBU]; double fb[NBU];
#pragma unroll
  for (int j = 0; j < NBU; j++) {
#pragma unroll
    for (int c = 0; c < 4; c++) x[j][c] = (lane * NBU + j < 3) ? 1.0 : 0.0;
    e[j] = 0; en_next[j] = 0; fb[j] = (j == 0 && lane == 0) ? 0.0 : 1.0;
  }
  int nb_en0 = 0;
  auto load_em = [&](int i, double& mb, double (&ml)[NBU][2]) {
    if (FLAGS & 1) {
      const double* Erow = E + (size_t)(i & Rm) * es;
      mb = Erow[0];
#pragma unroll
      for (int j = 0; j < NBU; j++) {
        ml[j][0] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(Erow) + ecol[j][0]);
        ml[j][1] = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(Erow) + ecol[j][1]);
      }
    } else {
      mb = 0.31;
#pragma unroll
      for (int j = 0; j < NBU; j++) { ml[j][0] = 0.3 + 0.01 * j; ml[j][1] = 0.29 + 0.01 * j; }
    }
  };
  auto snapshot = [&]() {
    int own[NBU]; bool alive[NBU];
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      const int mhi = max(max(__double2hiint(x[j][0]), __double2hiint(x[j][1])), max(__double2hiint(x[j][2]), __double2hiint(x[j][3])));
      alive[j] = mhi != 0;
      own[j] = e[j] + ((mhi >> 20) - 1023);
    }
    int Dj[NBU], vin[NBU];
    int run = 2 * kNegExp;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      Dj[j] = 192 * (lane * NBU + j);
      const int v = alive[j] ? own[j] + Dj[j] : 2 * kNegExp;
      run = max(run, v);
      vin[j] = run;
    }
    int tot = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(FULL, tot, d);
      if (lane >= d) tot = max(tot, t);
    }
    int excl = __shfl_up_sync(FULL, tot, 1);
    if (lane == 0) excl = 2 * kNegExp;
#pragma unroll
    for (int j = 0; j < NBU; j++) {
      const int pre = max(excl, vin[j]);
      en_next[j] = pre < kNegExp ? e[j] : pre - Dj[j];
    }
    nb_en0 = __shfl_up_sync(FULL, en_next[NBU - 1], 1);
  };
  auto frame = [&](int i, int k, double mb, double (&ml)[NBU][2]) {
    if ((FLAGS & 4) && k == 2) snapshot();
    const bool apply = (FLAGS & 4) && k == 0;
    double mbj[NBU];
#pragma unroll
    for (int j = 0; j < NBU; j++) mbj[j] = mb;
    if (apply) {
#pragma unroll
      for (int j = 0; j < NBU; j++) { const double f = pow2i(e[j] - en_next[j]); mbj[j] *= f; ml[j][0] *= f; ml[j][1] *= f; }
    }
    const double top = __shfl_up_sync(FULL, x[NBU - 1][3], 1);
    unsigned char* const vf = val + (size_t)(i & RVm) * vframe;
#pragma unroll
    for (int j = NBU - 1; j >= 0; j--) {
      const double in = j > 0 ? x[j - 1][3] : top;
      const double x0 = x[j][0], x1 = x[j][1], x2 = x[j][2], x3 = x[j][3];
      double s3 = x3 + x2;
      if (skipm & (1u << (2 * j + 1))) s3 += x1;
      const double s2 = x2 + x1;
      double s1 = x1 + x0;
      if (skipm & (1u << (2 * j))) s1 = fma(in, fb[j], s1);
      const double s0 = fma(in, fb[j], x0);
      x[j][3] = s3 * ml[j][1]; x[j][2] = s2 * mbj[j]; x[j][1] = s1 * ml[j][0]; x[j][0] = s0 * mbj[j];
      if (FLAGS & 2)
        *reinterpret_cast<uint4*>(vf + j * P + lane * 16) = make_uint4((uint32_t)__double2hiint(x[j][0]), (uint32_t)__double2hiint(x[j][1]), (uint32_t)__double2hiint(x[j][2]), (uint32_t)__double2hiint(x[j][3]));
    }
    if (FLAGS & 2) {
      int* const ve = reinterpret_cast<int*>(vf + NBU * P) + lane * NBP;
      int ev[NBP];
#pragma unroll
      for (int j = 0; j < NBP; j++) ev[j] = j < NBU ? (apply ? en_next[j] : e[j]) : 0;
#pragma unroll
      for (int u = 0; u < NBP / 4; u++) reinterpret_cast<int4*>(ve)[u] = make_int4(ev[4 * u], ev[4 * u + 1], ev[4 * u + 2], ev[4 * u + 3]);
    }
    if (apply) {
#pragma unroll
      for (int j = NBU - 1; j >= 1; j--) fb[j] = pow2i(en_next[j - 1] - en_next[j]);
      fb[0] = lane == 0 ? 0.0 : pow2i(nb_en0 - en_next[0]);
#pragma unroll
      for (int j = 0; j < NBU; j++) e[j] = en_next[j];
    }
    if (FLAGS & 8) { __syncwarp(); mbar_arrive_if(&bars[i & 15], lane == 0); }
  };
  const long long t0 = clock64();
  double mb_n, ml_n[NBU][2];
  load_em(0, mb_n, ml_n);
  for (int i0 = 0; i0 < T; i0 += UNROLL) {
    if ((FLAGS & 16) && (i0 & 7) == 0) {
      int m = words[0];
      for (int q = 1; q < 5; q++) m = min(m, words[q]);
      while (m < i0) { m = words[0]; }
    }
#pragma unroll
    for (int k = 0; k < UNROLL; k++) {
      const double mb = mb_n;
      double ml[NBU][2];
#pragma unroll
      for (int j = 0; j < NBU; j++) { ml[j][0] = ml_n[j][0]; ml[j][1] = ml_n[j][1]; }
      load_em(i0 + k + 1, mb_n, ml_n);
      frame(i0 + k, (i0 + k) & 3, mb, ml);
    }
  }
  const long long t1 = clock64();
  double acc = 0;
#pragma unroll
  for (int j = 0; j < NBU; j++) acc += x[j][0] + x[j][1] + x[j][2] + x[j][3] + e[j];
  sink[blockIdx.x * 32 + lane] = acc + mb_n;
  if (lane == 0) cyc[blockIdx.x] = t1 - t0;
  __syncwarp();
  if (lane == 0) { words[7] = 12345; asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(s32(&bars[31])) : "memory"); }
}

template <int NBU, int FLAGS, int UNROLL, int SPIN = 0, int RT = 0>
void run(const char* name, int nwarps, unsigned spin_mask = 0) {
  const int T = 400;
  long long* cyc; double* sink;
  cudaMalloc(&cyc, 256 * 8); cudaMalloc(&sink, 256 * 32 * 8 + 65536);
  const size_t smem = 180 * 1024;
  cudaFuncSetAttribute(lat<NBU, FLAGS, UNROLL, SPIN, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++) lat<NBU, FLAGS, UNROLL, SPIN, RT><<<128, 32 * nwarps, smem>>>(cyc, sink, T, 29, spin_mask, 33, NBU * 640 + 512, 127, 15);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[128];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s spin %d mask %02x NBU %d flags %2d unroll %d: %7.1f cycles/frame  (%s)\n", name, SPIN, spin_mask, NBU, FLAGS, UNROLL, (double)h[5] / T, cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  run<3, 31, 4, 0, 1>("CLUSTER(2) + 180 KB smem, runtime strides", 8);
  run<3, 31, 4, 1, 1>("CLUSTER(2) + 180 KB smem + try_wait spinners", 8, 0xee);
  run<3, 31, 4, 0, 1>("runtime strides + 128 regs", 8);
  run<3, 31, 4, 0, 0>("const strides + 128 regs", 8);
  run<3, 31, 4, 8>("8 KB code streamer: warp 1", 8, 0x02);
  run<3, 31, 4, 32>("32 KB code streamer: warp 1", 8, 0x02);
  run<3, 31, 4, 32>("32 KB code streamers: warps 1,2,3", 8, 0x0e);
  run<3, 31, 4, 64>("64 KB code streamer: warp 1", 8, 0x02);
  run<3, 31, 4, 64>("64 KB code streamers: warps 1,2,3", 8, 0x0e);
  run<3, 31, 4, 128>("128 KB code streamers: warps 1,2,3", 8, 0x0e);
  run<3, 31, 4, 1>("try_wait spinners: warps 1-3,5-7", 8, 0xee);
  run<3, 31, 4, 1>("try_wait spinner: warp 4 only", 8, 0x10);
  run<3, 31, 4, 1>("try_wait spinner: warp 1 only", 8, 0x02);
  run<3, 31, 4, 2>("nanosleep pollers: warps 1-3,5-7", 8, 0xee);
  run<3, 31, 4, 2>("nanosleep poller: warp 4 only", 8, 0x10);
  run<3, 31, 4, 3>("busy pollers: warps 1-3,5-7", 8, 0xee);
  run<3, 31, 4, 3>("busy poller: warp 4 only", 8, 0x10);
  run<3, 31, 4, 3>("busy poller: warp 1 only", 8, 0x02);
  run<3, 0, 4>("recurrence only", 8);
  run<3, 1, 4>("+ LDS emissions", 8);
  run<3, 3, 4>("+ val-ring STS", 8);
  run<3, 7, 4>("+ snapshot/apply", 8);
  run<3, 15, 4>("+ publish per frame", 8);
  run<3, 31, 4>("+ progress-word poll per 8 frames", 8);
  run<3, 31, 1>("everything, per-frame loop", 8);
  run<3, 31, 8>("everything, unroll 8", 8);
  run<3, 27, 4>("everything but snapshot", 8);
  run<1, 0, 4>("recurrence only", 8);
  run<1, 31, 4>("everything", 8);
  run<4, 0, 4>("recurrence only", 8);
  run<4, 31, 4>("everything", 8);
  run<2, 31, 4>("everything", 8);
  return 0;
}
