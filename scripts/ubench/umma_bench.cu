// umma_bench.cu — microbenchmarks that size the fused-MLP kernel design on B200 (not part of the library):
//   mma:   tcgen05.mma issue rate, A from smem (SS) or TMEM (TS), cta_group 1/2, N = 128/256, with optional
//          concurrent shared-memory store traffic from 8 "epilogue" warps
//   ldtm:  tcgen05.ld throughput per SM for 4/8/16 warps;  sttm: tcgen05.st throughput
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_bench umma_bench.cu
#include "../../mvip_nerf_b200/csrc/common.cuh"
#include <vector>
#include <stdlib.h>

void mvip_set_error(const char*, ...) {}

__device__ int g_commit_every = 16;
__device__ int g_commit_mask = 1;
__device__ int g_alt = 0;
__device__ int g_acol = 256;
__device__ int g_dcol = 0;
__device__ int g_mn = 0;   // 1: both smem operands MN-major (wgrad layout)
struct Res { unsigned long long cyc; unsigned long long noise_bytes; };

// kCta: 1 or 2; ts: A from TMEM; N: MMA N; noise_gap: cycles of spin between noise stores (0 = no noise)
template <int kCta>
__global__ void __launch_bounds__(384, 1) mma_bench(int ts, int N, int iters, int noise_gap, Res* res) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_done, bar_sink;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = kCta == 2 ? cluster_ctarank() : 0;
  // A: 4 chunk images (64 KB) at 0; B: 4 chunk images of (N/kCta) rows at 64 KB; noise scratch at 192 KB (16 KB)
  for (int i = tid; i < (192 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x007f007fu);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_sink, 1); mbar_fence_init(); stop_flag = 0; }
  if (warp == 2) { if (kCta == 2) tmem_alloc_2cta(&tmem_base_s, 512); else tmem_alloc(&tmem_base_s, 512); }
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t b_img = (uint32_t)(N / kCta) * 128u;

  if (warp == 1) {
    if (rank == 0) {
      const int mn = g_mn;
      const uint32_t idesc = umma_idesc_bf16(128 * kCta, N, mn, mn);
      const uint32_t sb = smem_u32(smem);
      long long t0 = clock64();
      const int ce = g_commit_every, cmask = g_commit_mask, alt = g_alt, acol = g_acol, dcol = g_dcol;
      for (int it = 0; it < iters; ++it) {
        if (elect_one_sync()) {
          const uint32_t d_t = tmem_base + dcol + ((alt && (it & 1)) ? 128 : 0);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              uint64_t db = mn ? umma_desc_sw128(sb + 65536 + (c * 4 + kk) * 2048, 8192, 1024) : umma_desc_sw128(sb + 65536 + c * b_img + kk * 32, 16, 1024);
              if (ts) {
                uint32_t ta = tmem_base + acol + c * 32 + kk * 8;
                if (kCta == 2) umma_bf16_ts_2cta(d_t, ta, db, idesc, 1u); else umma_bf16_ts(d_t, ta, db, idesc, 1u);
              } else {
                uint64_t da = mn ? umma_desc_sw128(sb + (c * 4 + kk) * 2048, 8192, 1024) : umma_desc_sw128(sb + c * 16384 + kk * 32, 16, 1024);
                if (kCta == 2) umma_bf16_2cta(d_t, da, db, idesc, 1u); else umma_bf16(d_t, da, db, idesc, 1u);
              }
              if (((c * 4 + kk + 1) % ce) == 0) { if (kCta == 2) umma_commit_2cta(&bar_sink, (uint16_t)cmask); else umma_commit(&bar_sink); }
            }
          }
        }
        __syncwarp();
      }
      if (elect_one_sync()) { if (kCta == 2) umma_commit_2cta(&bar_done, 1); else umma_commit(&bar_done); }
      __syncwarp();
      mbar_wait(&bar_done, 0);
      long long t1 = clock64();
      if ((tid & 31) == 0) res[blockIdx.x].cyc = (unsigned long long)(t1 - t0);
    }
    if ((tid & 31) == 0) stop_flag = 1;
  } else if (warp >= 4 && noise_gap < 0) {
    // 8 warps: 4 x tcgen05.ld.x16 (64 columns of the accumulator region) + pack + one tcgen05.st.x32 per batch
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 64;
    unsigned long long n = 0; uint32_t accx = 0;
    while (!stop_flag) {
      uint32_t a[16], b[16], c[16], d[16];
      tmem_ld16(taddr, a); tmem_ld16(taddr + 16, b); tmem_ld16(taddr + 32, c); tmem_ld16(taddr + 48, d);
      tmem_ld_wait();
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) { pk[i] = a[2*i] ^ a[2*i+1]; pk[8+i] = b[2*i] ^ b[2*i+1]; pk[16+i] = c[2*i] ^ c[2*i+1]; pk[24+i] = d[2*i] ^ d[2*i+1]; }
      if (noise_gap <= -1000) { tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 384 + ((warp - 4) >> 2) * 32, pk); tmem_st_wait(); }
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) accx ^= pk[i];
      }
      n += 32 * 64 * 4;
      long long t = clock64();
      const int gap = (-noise_gap) % 1000;
      while (clock64() - t < gap) {}
    }
    if (accx == 0x1234567u) n += 1;
    if (kCta == 1 || rank == 0) atomicAdd(&res[blockIdx.x].noise_bytes, n);
  } else if (warp >= 4 && noise_gap > 0) {
    // 8 warps x 512 B per store instruction, conflict-free
    uint4 v = make_uint4(tid, tid + 1, tid + 2, tid + 3);
    uint8_t* dst = smem + 192 * 1024 + (tid - 128) * 16;
    unsigned long long n = 0;
    while (!stop_flag) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(smem_u32(dst + (j & 1) * 4096)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); v.x += j; }
      n += 64;
      long long t = clock64();
      while (clock64() - t < noise_gap) {}
    }
    if (kCta == 1 || rank == 0) atomicAdd(&res[blockIdx.x].noise_bytes, n);
  }
  tc_fence_before();
  __syncthreads();
  if (kCta == 2) cluster_sync_all();
  if (warp == 2) { if (kCta == 2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

// tcgen05.ld / st throughput: nwarps warps, each `iters` x (x32 op on its lane quarter)
__global__ void __launch_bounds__(512, 1) tmem_bench(int store, int iters, int batch, Res* res) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = tid + i;
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (store) {
      for (int b = 0; b < batch; ++b) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
          :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
             "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
      for (int b = 0; b < batch; ++b) {
        uint32_t q[32];
        tmem_ld32(taddr, q);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= q[i];
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (tid == 0) res[blockIdx.x].cyc = (unsigned long long)(t1 - t0);
  if (acc == 0x12345u) res[blockIdx.x].noise_bytes = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}
// pipelined loads: issue `batch` x16 loads back to back, one wait
__global__ void __launch_bounds__(512, 1) tmem_bench_pipe(int iters, Res* res) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t a[16], b[16], c[16], d[16];
    tmem_ld16(taddr, a); tmem_ld16(taddr + 16, b); tmem_ld16(taddr + 32, c); tmem_ld16(taddr + 48, d);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= a[i] ^ b[i] ^ c[i] ^ d[i];
  }
  long long t1 = clock64();
  __syncthreads();
  if (tid == 0) res[blockIdx.x].cyc = (unsigned long long)(t1 - t0);
  if (acc == 0x12345u) res[blockIdx.x].noise_bytes = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <int kCta>
void run_mma(int ts, int N, int gap, Res* d_res, int grid) {
  const int iters = 2000;
  const size_t smem = 208 * 1024 + 1024;
  CK(cudaMemset(d_res, 0, sizeof(Res) * 148));
  if (kCta == 2) {
    CK(cudaFuncSetAttribute(mma_bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, mma_bench<2>, ts, N, iters, gap, d_res));
  } else {
    CK(cudaFuncSetAttribute(mma_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_bench<1><<<grid, 384, smem>>>(ts, N, iters, gap, d_res);
  }
  CK(cudaDeviceSynchronize());
  std::vector<Res> h(148);
  CK(cudaMemcpy(h.data(), d_res, sizeof(Res) * 148, cudaMemcpyDeviceToHost));
  double cyc = 0, nb = 0; int n = 0;
  for (int i = 0; i < grid; i += kCta) { cyc += h[i].cyc; nb += h[i].noise_bytes; ++n; }
  cyc /= n; nb /= n;
  const double per_mma = cyc / (iters * 16.0);
  const double floor_cyc = 128.0 * N / 256.0;   // per SM: 128 x N x 16 MACs at 4096 MAC/cyc
  printf("mma cta_group=%d %s N=%3d noise_gap=%4d: %.1f cyc/MMA (floor %.0f) -> %.1f%% of tensor peak; noise %.1f B/cyc/SM\n", kCta, ts ? "TS" : "SS", N,
         gap, per_mma, floor_cyc, 100 * floor_cyc / per_mma, nb / cyc);
}

int main() {
  Res* d_res; CK(cudaMalloc(&d_res, sizeof(Res) * 148));
  for (int mn = 0; mn < 2; ++mn) {
    CK(cudaMemcpyToSymbol(g_mn, &mn, 4));
    printf("smem operands %s: ", mn ? "MN-major (wgrad)" : "K-major");
    run_mma<1>(0, 256, 0, d_res, 148);
    printf("smem operands %s: ", mn ? "MN-major (wgrad)" : "K-major");
    run_mma<2>(0, 256, 0, d_res, 148);
    printf("smem operands %s: ", mn ? "MN-major (wgrad)" : "K-major");
    run_mma<1>(0, 128, 0, d_res, 148);
  }
  { int mn = 0; CK(cudaMemcpyToSymbol(g_mn, &mn, 4)); }
  { int ce = 16, mask = 1, alt = 0; CK(cudaMemcpyToSymbol(g_commit_every, &ce, 4)); CK(cudaMemcpyToSymbol(g_commit_mask, &mask, 4)); CK(cudaMemcpyToSymbol(g_alt, &alt, 4)); }
  if (getenv("UBENCH_ALL"))
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {256, 128})
      for (int gap : {0, 400, 100, 1}) {
        run_mma<1>(ts, N, gap, d_res, 148);
        run_mma<2>(ts, N, gap, d_res, 148);
      }
  if (!getenv("UBENCH_ALL")) return 0;
  for (int store = 0; store < 2; ++store)
    for (int nw : {4, 8, 16}) {
      CK(cudaMemset(d_res, 0, sizeof(Res) * 148));
      const int iters = 2000, batch = 4;
      tmem_bench<<<148, nw * 32, 0>>>(store, iters, batch, d_res);
      CK(cudaDeviceSynchronize());
      Res h; CK(cudaMemcpy(&h, d_res, sizeof(Res), cudaMemcpyDeviceToHost));
      printf("tmem %s x32 (serial wait) %2d warps: %.1f B/cyc/SM, %.1f cyc per op per warp\n", store ? "st" : "ld", nw, (double)nw * iters * batch * 4096.0 / h.cyc, (double)h.cyc / (iters * batch));
    }
  for (int nw : {4, 8, 16}) {
    CK(cudaMemset(d_res, 0, sizeof(Res) * 148));
    tmem_bench_pipe<<<148, nw * 32, 0>>>(4000, d_res);
    CK(cudaDeviceSynchronize());
    Res h; CK(cudaMemcpy(&h, d_res, sizeof(Res), cudaMemcpyDeviceToHost));
    printf("tmem ld 4 x x16 pipelined %2d warps: %.1f B/cyc/SM\n", nw, (double)nw * 4000 * 4 * 2048.0 / h.cyc);
  }
  return 0;
}
