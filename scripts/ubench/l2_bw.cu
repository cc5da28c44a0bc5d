// l2_bw.cu — how much bandwidth do L2 hits give on B200, for bulk (TMA) loads / stores issued by persistent CTAs?
// Decides whether handing the dZ stash from the dgrad chain to wgrad through L2 (instead of HBM) can pay.
//   modes: R  = every CTA streams bulk loads (chunk bytes each, ring of stages) over a region of S bytes (S <= L2: hits)
//          W  = bulk stores over a region of S bytes
//          RW = both at once: loads over region A (size S), stores over region B (size S)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_bw l2_bw.cu
#include "../../mvip_nerf_b200/csrc/common.cuh"
#include <vector>
#include <stdlib.h>
void mvip_set_error(const char*, ...) {}

constexpr int kStages = 6;

__global__ void __launch_bounds__(128, 1) bw_kernel(const uint8_t* src, uint8_t* dst, size_t region, uint32_t chunk, int iters, int mode,
                                                    int evict_first) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
  __syncthreads();
  const size_t n_chunks = region / chunk;
  if (warp == 0 && lane == 0 && (mode & 1)) {
    const uint64_t pol = l2_policy_evict_first();
    // prologue: fill the ring, then wait/refill in order
    size_t c = (size_t)blockIdx.x * 977 % n_chunks;
    for (int i = 0; i < iters + kStages; ++i) {
      const int s = i % kStages;
      if (i >= kStages) mbar_wait(&full[s], ((i / kStages) - 1) & 1);
      if (i < iters) {
        mbar_arrive_expect_tx(&full[s], chunk);
        if (evict_first) tma_load_1d_hint(smem + (size_t)s * chunk, src + c * chunk, chunk, &full[s], pol);
        else tma_load_1d(smem + (size_t)s * chunk, src + c * chunk, chunk, &full[s]);
        c += gridDim.x; if (c >= n_chunks) c -= n_chunks;
      }
    }
  }
  if (warp == 1 && lane == 0 && (mode & 2)) {
    size_t c = (size_t)blockIdx.x * 977 % n_chunks;
    uint8_t* stg = smem + (size_t)kStages * chunk;      // never written: content irrelevant
    for (int i = 0; i < iters; ++i) {
      tma_store_1d(dst + c * chunk, stg, chunk);
      tma_store_commit();
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      c += gridDim.x; if (c >= n_chunks) c -= n_chunks;
    }
    tma_store_wait_all0();
  }
}

int main(int argc, char** argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = (size_t)4 << 30;
  uint8_t *a, *b;
  cudaMalloc(&a, big); cudaMalloc(&b, big);
  cudaMemset(a, 1, big); cudaMemset(b, 2, big);
  const uint32_t chunk = 32768;
  const size_t smem = (size_t)(kStages + 1) * chunk + 1024;
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"", "R ", "W ", "RW"};
  printf("mode region_MB  chunk  grid  evict_first  GB/s(read)  GB/s(write)  GB/s(total)\n");
  for (int mode = 1; mode <= 3; ++mode)
    for (size_t mb : {8, 24, 48, 96, 192, 4096})
      for (int ef = 0; ef <= ((mode & 1) ? 1 : 0); ++ef) {
        const size_t region = mb << 20;
        const int iters = 6000;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          bw_kernel<<<sms, 128, smem>>>(a, b, region, chunk, iters, mode, ef);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
        }
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)sms * iters * chunk;
        const double r = (mode & 1) ? bytes / ms / 1e6 : 0, w = (mode & 2) ? bytes / ms / 1e6 : 0;
        printf("%s   %6zu  %6u  %4d  %d  %10.0f  %10.0f  %10.0f\n", names[mode], mb, chunk, sms, ef, r, w, r + w);
      }
  return 0;
}
