// l2_handoff.cu — does data written with bulk (TMA) stores stay in L2 long enough to be read back from L2, and can the
// write-back to HBM be avoided altogether (discard.global.L2) once the consumer has read it?  Sizes the dZ hand-over of the
// fused backward: the chain writes 4.9 KB / point of dZ, the wgrad role reads it back a few 10 us later.
//
// Every CTA runs three single-thread roles on its own slice of the buffers:
//   producer   bulk-stores 32 KB chunks (content irrelevant) into a ring of `lag + 2` chunks per CTA -> footprint = grid * (lag + 2) * 32 KB
//   consumer   bulk-loads the chunk the producer completed `lag` chunks ago; optionally discards its lines afterwards
//   background bulk-loads `bg` x 32 KB per produced chunk from a 4 GB region (evict_first): the activation stash streaming past
// Run under ncu for the DRAM bytes:  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ./l2_handoff
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o l2_handoff l2_handoff.cu
#include "../../mvip_nerf_b200/csrc/common.cuh"
#include <vector>
#include <stdlib.h>
void mvip_set_error(const char*, ...) {}

constexpr uint32_t kChunk = 32768;
constexpr int kBgStages = 4;

__device__ __forceinline__ void tma_store_1d_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

struct Cfg { int lag, store_hint, load_hint, discard, bg, iters; };

__global__ void __launch_bounds__(128, 1) handoff_kernel(uint8_t* ring, const uint8_t* big, size_t big_bytes, Cfg c, volatile uint32_t* progress) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_c[2], bar_b[kBgStages];
  __shared__ volatile uint32_t produced;     // chunks whose bulk stores are complete
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_c[0], 1); mbar_init(&bar_c[1], 1);
    for (int i = 0; i < kBgStages; ++i) mbar_init(&bar_b[i], 1);
    mbar_fence_init();
    produced = 0;
  }
  __syncthreads();
  const int ring_n = c.lag + 2;
  uint8_t* my_ring = ring + (size_t)blockIdx.x * ring_n * kChunk;
  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_last();
      for (int i = 0; i < c.iters; ++i) {
        uint8_t* dst = my_ring + (size_t)(i % ring_n) * kChunk;
        // the ring slot is free: the consumer is at most lag + 1 behind... enforce it
        while ((int)(i - *(volatile uint32_t*)&progress[blockIdx.x]) > c.lag + 1) {}
        if (c.store_hint) tma_store_1d_hint(dst, smem, kChunk, pol); else tma_store_1d(dst, smem, kChunk);
        tma_store_commit();
        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");      // chunk i - 1 is complete (written, not only read)
        if (i >= 1) { __threadfence(); produced = i; }
      }
      tma_store_wait_all0();
      __threadfence();
      produced = c.iters;
    }
  } else if (warp == 1) {
    const uint64_t pol = l2_policy_evict_first();
    for (int i = 0; i + c.lag < c.iters + c.lag; ++i) {
      if (i >= c.iters) break;
      const uint8_t* src = my_ring + (size_t)(i % ring_n) * kChunk;
      if (lane == 0) {
        // consume chunk i once chunk i + lag has been produced (or everything has)
        const uint32_t need = (uint32_t)((i + c.lag + 1 < c.iters) ? i + c.lag + 1 : c.iters);
        while (produced < need) {}
        fence_proxy_async_all();
        mbar_arrive_expect_tx(&bar_c[i & 1], kChunk);
        if (c.load_hint) tma_load_1d_hint(smem + kChunk, src, kChunk, &bar_c[i & 1], pol);
        else tma_load_1d(smem + kChunk, src, kChunk, &bar_c[i & 1]);
      }
      mbar_wait(&bar_c[i & 1], (i >> 1) & 1);
      if (c.discard) {
        for (int l = lane; l < (int)(kChunk / 128); l += 32)
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(src + (size_t)l * 128) : "memory");
      }
      __syncwarp();
      if (lane == 0) { __threadfence(); progress[blockIdx.x] = i + 1; }
    }
  } else if (warp == 2 && c.bg > 0) {
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      const size_t n_chunks = big_bytes / kChunk;
      size_t ch = (size_t)blockIdx.x * 7919 % n_chunks;
      const int total = c.iters * c.bg;
      for (int i = 0; i < total + kBgStages; ++i) {
        const int s = i % kBgStages;
        if (i >= kBgStages) mbar_wait(&bar_b[s], ((i / kBgStages) - 1) & 1);
        if (i < total) {
          // keep pace with the producer: never more than bg * (produced + 2) chunks
          while (i > (int)(produced + 2) * c.bg && produced < (uint32_t)c.iters) {}
          mbar_arrive_expect_tx(&bar_b[s], kChunk);
          tma_load_1d_hint(smem + (size_t)(2 + s) * kChunk, big + ch * kChunk, kChunk, &bar_b[s], pol);
          ch += gridDim.x; if (ch >= n_chunks) ch -= n_chunks;
        }
      }
    }
  }
}

// evicts everything: reads 1 GB with plain loads (forces the write-back of whatever is still dirty)
__global__ void flush_kernel(const uint4* p, size_t n, uint4* out) {
  uint4 a = make_uint4(0, 0, 0, 0);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
  }
  if (a.x == 0x12345678u) *out = a;
}

int main(int argc, char** argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t big = (size_t)4 << 30;
  uint8_t *ring, *bigbuf; uint32_t* progress;
  cudaMalloc(&ring, (size_t)1 << 30); cudaMalloc(&bigbuf, big); cudaMalloc(&progress, sms * sizeof(uint32_t));
  cudaMemset(ring, 1, (size_t)1 << 30); cudaMemset(bigbuf, 2, big);
  const size_t smem = (size_t)(2 + kBgStages) * kChunk + 1024;
  cudaFuncSetAttribute(handoff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("launch  lag  footprint_MB  store_hint  load_hint  discard  bg   ms    produced_GB  GB/s(produced)\n");
  int n = 0;
  const int iters = 1500;
  for (int lag : {2, 6, 12, 20})
    for (int bg : {0, 1})
      for (int variant = 0; variant < 4; ++variant) {
        Cfg c; c.lag = lag; c.bg = bg; c.iters = iters;
        c.store_hint = (variant == 1 || variant == 3); c.load_hint = (variant >= 2); c.discard = (variant == 3);
        if (variant == 2) c.discard = 1;
        cudaMemset(progress, 0, sms * sizeof(uint32_t));
        cudaEventRecord(e0);
        handoff_kernel<<<sms, 128, smem>>>(ring, bigbuf, big, c, progress);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { printf("error %s\n", cudaGetErrorString(err)); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double gb = (double)sms * iters * kChunk / 1e9;
        printf("%4d  %4d  %8.1f  %d  %d  %d  %d  %7.3f  %6.2f  %8.0f\n", n, lag, (double)sms * (lag + 2) * kChunk / 1e6, c.store_hint, c.load_hint,
               c.discard, bg, ms, gb, gb / ms * 1e3);
        ++n;
        flush_kernel<<<sms * 8, 256>>>(reinterpret_cast<const uint4*>(bigbuf), ((size_t)1 << 30) / 16, reinterpret_cast<uint4*>(ring));
        cudaDeviceSynchronize();
      }
  return 0;
}
