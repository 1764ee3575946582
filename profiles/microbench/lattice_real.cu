// The REAL lattice-warp function (end2end_b200/csrc/ctc_fused_impl.cuh) in isolation: a fake producer warp only
// releases the emission blocks, nobody drains the val ring.  Prints cycles per frame of the lattice warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I../../end2end_b200/csrc -o lattice_real lattice_real.cu
#include <cstdio>
#include "ctc_fused_impl.cuh"
namespace e2e { void set_error(const char*, ...) {} void launch_begin(int, cudaStream_t) {} void launch_end(int, cudaStream_t) {} }
using namespace e2e;

template <int NBU, bool BWD, bool SCALER>
__global__ void __maxnreg__(128) solo(const FzParams p, long long* cyc, int Ti, int Li) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FzView sv = fz_carve(smem_raw, p.L);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 4 + 1; i += blockDim.x) sv.lab[i] = i < Li ? 1 + (i * 7) % 28 : 0;
  for (int k = tid; k < p.L.R * p.L.es; k += blockDim.x) sv.E[k] = 0.2 + 0.001 * (k % 89);
  uint32_t* z = reinterpret_cast<uint32_t*>(smem_raw + p.L.off_ctl);
  for (int k = tid; k < (int)(sizeof(FzCtl) >> 2); k += blockDim.x) z[k] = 0u;
  __syncthreads();
  if (tid < 8) sv.ctl->comb_done[tid] = 1 << 30;
  if (tid < 32) fz_mbar_init(&sv.ctl->full[tid], 1);
  else if (tid < 48) fz_mbar_init(&sv.ctl->fullE[tid - 32], 1);
  else if (tid < 50) fz_mbar_init(&sv.ctl->sc_full[tid - 48], 1);
  __syncthreads();
  if (w == 0) {
    const long long t0 = clock64();
    fz_lattice<NBU, BWD, false, SCALER>(p, sv, Ti, Li, lane);
    if (lane == 0) cyc[blockIdx.x] = clock64() - t0;
  } else if (w == 4) {
    if (SCALER) fz_scaler<NBU>(p, sv, Ti, lane);
  } else if (w == 1) {
    const int nblocks = (Ti + 7) / 8;
    for (int bi = 0; bi < nblocks; bi++) {
      while (sv.ctl->lat_prog < bi * 8 + 8 - p.L.R) __nanosleep(32);
      __syncwarp();
      if (lane == 0) fz_mbar_arrive(&sv.ctl->fullE[bi & p.L.neb_mask]);
    }
  }
}

template <int NBU, bool BWD, bool SCALER>
void run(int Li, const char* name) {
  FzParams p; memset(&p, 0, sizeof(p));
  FzLayout& L = p.L;
  L.NB = 4; L.NP = 1; L.NC = 1; L.PF = 2; L.nwarps = 8; L.nap = 32; L.R = 128; L.RV = 16; L.CF = 8; L.es = 31;
  L.vframe = 4 * 640 + 512; L.rv_log2 = 4; L.neb_log2 = 4; L.neb_mask = 15; L.PB = 8; L.pb_log2 = 3;
  L.off_lab = 0; L.off_occ = 1040; L.off_E = 1040 + 2 * 20 * 128; L.off_val = L.off_E + 128 * 31 * 8; L.off_stage = L.off_val + 16 * L.vframe;
  L.off_post = L.off_stage; L.off_ctl = L.off_stage + 64; L.total = L.off_ctl + 1024;
  p.V = 29; p.blank = 0; p.B = 128; p.T = 400;
  long long* cyc; cudaMalloc(&cyc, 256 * 8);
  cudaFuncSetAttribute(solo<NBU, BWD, SCALER>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
  const int T = 400;
  for (int rep = 0; rep < 2; rep++) solo<NBU, BWD, SCALER><<<128, 256, L.total>>>(p, cyc, T, Li);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[128]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-30s scaler %d NBU %d bwd %d: %7.1f cycles/frame (%s)\n", name, (int)SCALER, NBU, (int)BWD, (double)h[7] / T, cudaGetErrorString(e));
  cudaFree(cyc);
}

int main() {
  run<3, false, false>(150, "real fz_lattice");
  run<3, false, true>(150, "real fz_lattice");
  run<3, true, true>(150, "real fz_lattice");
  run<4, false, true>(200, "real fz_lattice");
  run<1, false, true>(25, "real fz_lattice");
  run<1, false, false>(25, "real fz_lattice");
  return 0;
}
