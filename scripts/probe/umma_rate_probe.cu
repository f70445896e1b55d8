// umma_rate_probe.cu — hardware probe (not product code): how fast does ONE issuing thread stream tcgen05.mma
// (kind::f16, bf16 operands, M = 128, K = 16) as a function of N and of the accumulator dependency pattern?
//
// Why: the gather kernels issue 4 x (taps x chunks) MMAs per tile into ONE accumulator.  If back-to-back MMAs into the
// same TMEM accumulator are spaced by a fixed pipeline latency rather than by their work (M*N*K / 4096 MAC/clk), small-N
// layers (32 / 64 output channels: 16 / 32 clk of work per MMA) are latency-bound and the cure is independent
// accumulator chains, not fewer bytes.  The probe times R MMAs (operands = zeros already in shared memory, no TMA) for
// N in {32, 64, 128, 256} with 1, 2 and 4 interleaved accumulators, with 1 and 2 CTAs per SM.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I nemar_b200/csrc -I include -o scripts/probe/umma_rate_probe.bin scripts/probe/umma_rate_probe.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace tc;

// ELECT = true: the issuing warp stays converged and one lane is chosen by elect.sync per group of 4 MMAs (what the
// engine does since this probe); false: everything inside `threadIdx.x == 0` (what it did before: ptxas then wraps every
// UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop).
template <int N, bool ELECT>
__global__ void __launch_bounds__(128) rate_kernel(int reps, int naccs, int kper, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                          // 128 rows x 64 bf16, SWIZZLE_128B K-major (zeros)
  uint8_t* sB = smem + 16384;                  // N rows x 64 bf16
  uint64_t* bar = (uint64_t*)(sB + 32768);
  uint32_t* tmem_slot = (uint32_t*)(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 512 / 2);          // 256 columns: two CTAs may share an SM
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (ELECT && warp == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t adesc = make_smem_desc(smem_u32(sA), 16, 1024, LAYOUT_SW128);
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, LAYOUT_SW128);
    if (elect_one()) {
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k > 0);
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
      const uint32_t acc = tmem + (uint32_t)((r / kper) % naccs) * (uint32_t)N;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 1);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  if (!ELECT && threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t adesc = make_smem_desc(smem_u32(sA), 16, 1024, LAYOUT_SW128);
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, LAYOUT_SW128);
    // warm-up
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k > 0);
    umma_commit(bar);
    mbar_wait(bar, 0);
    tc_fence_after();
    const long long t0 = clock64();
    // `kper` consecutive MMAs go to one accumulator (one k-step of the conv kernels), then the next accumulator
    for (int r = 0; r < reps; ++r) {
      const uint32_t acc = tmem + (uint32_t)((r / kper) % naccs) * (uint32_t)N;
      umma_bf16(acc, adesc + (uint64_t)((r & 3) * 2), bdesc + (uint64_t)((r & 3) * 2), idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 1);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

template <int N, bool ELECT>
static void run(int grid, int reps, int naccs, int kper, long long* d_cyc) {
  if (naccs * N > 256) return;
  cudaFuncSetAttribute(rate_kernel<N, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  rate_kernel<N, ELECT><<<grid, 128, 56 * 1024>>>(reps, naccs, kper, d_cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d naccs=%d: %s\n", N, naccs, cudaGetErrorString(e)); exit(1); }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double avg = 0, mx = 0;
  for (long long v : h) { avg += (double)v; if ((double)v > mx) mx = (double)v; }
  avg /= grid;
  const double work = 128.0 * N * 16 / 4096.0;     // clk per MMA at 4096 MAC/clk/SM
  printf("RATE %s N=%3d ctas/SM=%d accs=%d kper=%d : %7.1f clk/MMA (max %7.1f)  work %5.1f clk  -> %5.1f %% of the MMA rate per CTA\n", ELECT ? "elect " : "lane==0", N,
         grid > 148 ? 2 : 1, naccs, kper, avg / reps, mx / reps, work, 100.0 * work / (avg / reps));
}

int main() {
  long long* d_cyc;
  cudaMalloc(&d_cyc, sizeof(long long) * 296);
  const int reps = 2048;
  for (int grid : {148, 296}) {
    for (int naccs : {1, 2}) {
      run<32, false>(grid, reps, naccs, 4, d_cyc);
      run<128, false>(grid, reps, naccs, 4, d_cyc);
      run<256, false>(grid, reps, naccs, 4, d_cyc);
    }
    for (int naccs : {1, 2, 4}) {
      run<32, true>(grid, reps, naccs, 4, d_cyc);
      run<64, true>(grid, reps, naccs, 4, d_cyc);
      run<128, true>(grid, reps, naccs, 4, d_cyc);
      run<256, true>(grid, reps, naccs, 4, d_cyc);
    }
  }
  printf("SUMMARY done\n");
  return 0;
}
