// mlp_backward.cu — backward of the fused NeRF MLP (what autograd computes for NeRF.forward,
// DS_NeRF/run_nerf_helpers.py:104-127), from the activation stash written by mlp_forward.cu.
//
// Three launches per call:
//   1. backward_fused_kernel  ONE persistent tcgen05 kernel, both roles in every CTA pair:
//        chain  d_raw -> d hidden_pre -> d feature -> dZ_7 .. dZ_0 (grad w.r.t. each layer's pre-activation), tile-major like
//               the forward: gradient tiles resident in tensor memory, transposed weights streamed by TMA, ReLU masks from the
//               stash; every dZ tile image is bulk-stored to the dZ stash;
//        wgrad  dW_l = dZ_l^T X_l (K = points) with MN-major operands: the stash images are used as-is (the bytes of a K-major
//               [points x features] tile are an MN-major operand when points are the contraction); the pair is bound to one
//               layer, its 256 x 256 fp32 accumulator stays in TMEM for the whole launch and is flushed once; the dZ tiles are
//               picked up out of L2 right after the chain (on any SM) has published them; two warps sum the dZ columns
//               (bias grads) from the staged operand.
//   2. head_grads_kernel      CUDA cores: alpha_linear / rgb_linear weight + bias grads (N = 1 and 3).
//   3. reduce_kernel          deterministic (fixed-order) sum of the partials into the caller's grad tensors.
// No gradient flows to pts / viewdirs (none is required by the reference: z_samples is detached, run.py:1812).
#include "mlp_pair.cuh"
#include <stdlib.h>

namespace {
using namespace mlp;

constexpr int kCSteps = 9;          // chain steps: 0 = d feature, s >= 1: dZ_{8-s}

struct ChainParams {
  const uint8_t* packed;
  const float4* d_raw;
  const uint8_t* stash;
  uint8_t* dz;
  int64_t n_points;
  int64_t n_tiles;
  uint32_t* flags;      // [n_tiles][kFlagsPerTile] "dZ unit is in the stash" flags, set by the chain's store warp
  uint32_t* cons_stamp; // [n_tiles][kFlagsPerTile] %globaltimer_lo when the consuming pair started on the unit (debug aid)
  uint32_t* next_pair;  // [0] tile pairs handed out beyond the first one of every CTA pair; [16 + 4 c + (k & 3)]: rank 0 -> rank 1 of pair c
  uint32_t* credit;     // [0] 2 x dZ units published, [1] 2 x dZ units picked up (a unit read by two work items counts 1 per item)
  int throttle_units;   // a chain does not start a tile while more than this many units are published but not picked up ...
  int throttle_cycles;  // ... for at most this many cycles (a soft limit: it can delay, never block)
  int throttle_gain;    // > 0: the delay is (excess units) x gain cycles, decided once per tile; 0: wait until the excess is gone
  int stagger;          // cycles by which the chain of cluster c starts after that of cluster c - 1 (0: all at once)
};
constexpr int kFlagsPerTile = 10;   // 0: d hidden_pre (input stage), 1 + s: output of chain step s

__device__ __forceinline__ int chain_nchunks(int s) { return s == 0 ? 2 : 4; }

constexpr int kStagePts = 64;                            // points per stage of the wgrad ring (whole-tile stages, 2 x 64 KB, measured slower: 1.41 vs 1.36 ms)
constexpr int kSubStages = kTile / kStagePts;            // stages per tile
constexpr uint32_t kHalf = kStagePts * 128;              // bytes of one operand piece: kStagePts points of a chunk image
constexpr int kNumItems = 11;
constexpr size_t kPartialSlotBytes = (size_t)256 * 320 * sizeof(float);
constexpr int kFusedSlots = 80;                          // partial slots: one per CTA pair (74 on a B200)

// cycle counters of cluster 0 (debug aid, mvip_debug_wgrad_profile): [0] chain issuer total, [1] waiting for the epilogue,
// [2] waiting for weights, [3] wgrad issuer total, [4] waiting for operands, [5] wgrad producer: flag wait, [6] stage wait, [7] total
#ifndef MVIP_PROF_CLUSTER
#define MVIP_PROF_CLUSTER 0      // the CTA pair whose cycle counters are recorded
#endif
__device__ unsigned long long g_wprof[8];
// hand-over lag of the dZ units (debug aid, mvip_debug_bwd_lag): per CTA pair [0] sum, [1] max of (consumer picks the unit up) -
// (chain published it) in ns of %globaltimer, [2] units consumed, [3] units that were already published when the consumer asked
__device__ unsigned long long g_lag[80][4];
__device__ __forceinline__ uint32_t globaltimer_lo() {
  uint32_t t;
  asm volatile("mov.u32 %0, %%globaltimer_lo;" : "=r"(t));
  return t;
}
#ifdef MVIP_TRACE_BWD
// event stamps of epilogue warp 0 (lane 0) of cluster 0, leader CTA, for the 4th tile pair: [half-step 0..18][event]
__device__ long long g_btrace[20][12];
#define BTR(hs, ev) do { if (tr_on) g_btrace[hs][ev] = clock64(); } while (0)
#else
#define BTR(hs, ev)
#endif

// ---- geometry of the weight-gradient partial slots, one entry per (parameter, column block): used by the reduce ----
struct WItem {
  int a_chunk;      // first dZ-stash chunk of the M side
  int m_blocks;     // 1 (128 output rows) or 2 (256)
  int nb;           // number of 64-wide B chunks: a slot row has 64 nb floats
  int b_chunk[5];   // forward-stash chunk index of each
  int dst;          // parameter index of the weight
  int ld;           // its leading dimension
  int col[5];       // destination column of each B chunk
  int valid[5];     // valid columns of each B chunk
  int bias;         // parameter index of the bias grad computed with this item, or -1
  int cost;         // (unused by the fused backward)
};

__constant__ WItem kItems[kNumItems] = {
    // views_linears.0: d hidden_pre^T [feature | PE(viewdir)]
    {kDzHidden, 1, 5, {33, 34, 35, 36, 37}, kPViewsW, 283, {0, 64, 128, 192, 256}, {64, 64, 64, 64, 27}, kPViewsB, 7},
    // feature_linear: d feature^T h8
    {kDzFeat, 2, 4, {29, 30, 31, 32, 0}, kPFeatW, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, kPFeatB, 8},
    // pts_linears 7, 6
    {kDzTrunk + 0, 2, 4, {25, 26, 27, 28, 0}, 14, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, 15, 8},
    {kDzTrunk + 4, 2, 4, {21, 22, 23, 24, 0}, 12, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, 13, 8},
    // pts_linears 5: h part (columns 63..318) and PE part (columns 0..62)
    {kDzTrunk + 8, 2, 4, {17, 18, 19, 20, 0}, 10, 319, {63, 127, 191, 255, 0}, {64, 64, 64, 64, 0}, 11, 8},
    {kDzTrunk + 8, 2, 1, {0, 0, 0, 0, 0}, 10, 319, {0, 0, 0, 0, 0}, {63, 0, 0, 0, 0}, -1, 5},
    // pts_linears 4..1
    {kDzTrunk + 12, 2, 4, {13, 14, 15, 16, 0}, 8, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, 9, 8},
    {kDzTrunk + 16, 2, 4, {9, 10, 11, 12, 0}, 6, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, 7, 8},
    {kDzTrunk + 20, 2, 4, {5, 6, 7, 8, 0}, 4, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, 5, 8},
    {kDzTrunk + 24, 2, 4, {1, 2, 3, 4, 0}, 2, 256, {0, 64, 128, 192, 0}, {64, 64, 64, 64, 0}, 3, 8},
    // pts_linears 0: dZ_0^T PE
    {kDzTrunk + 28, 2, 1, {0, 0, 0, 0, 0}, 0, 63, {0, 0, 0, 0, 0}, {63, 0, 0, 0, 0}, 1, 5},
};
constexpr int kRealItems = kNumItems;

struct Segment {
  int item;       // kItems index of the slot; -1 or t1 == t0 (the zeroed table) = empty
  int t0, t1;     // tiles t0, t0 + stride, ... < t1
  int stride;
};

struct WParams {
  const uint8_t* stash;
  const uint8_t* dz;
  int64_t n_tiles;
  float* partials;        // [kFusedSlots] slots of kPartialSlotBytes
  float* bias_partials;   // [kFusedSlots][2 bias warps][256]
  Segment* segs;          // [kFusedSlots]
};

// =================================================================================================
// 1. FUSED backward: the dgrad chain and ALL weight-gradient GEMMs in ONE persistent launch, BOTH roles in EVERY CTA pair,
//    dZ handed from the chain to wgrad through L2.
//
//     With separate kernels wgrad re-reads the whole dZ stash (4.9 KB / point) from HBM after the chain has written it, and
//     the chain leaves the tensor pipe idle while its epilogue warps work (it is bound by their latency, not by HBM).
//     Here every CTA pair runs
//       * the chain on ONE tile slot per CTA (TMEM columns [0, 256): A operand + one accumulator half), tile pairs taken
//         round-robin over the clusters, the chain of cluster c starting c x stagger cycles late.  The epilogue warps write
//         packed bf16 rows into ONE 32 KB staging buffer; a dedicated store thread sends it to the dZ stash as one bulk
//         store and publishes each of the 10 dZ UNITS of a tile (the output of one chain step) as soon as its stores are
//         complete (st.release.gpu of a flag; only that thread pays for the fences: a red.release.gpu per group on the
//         epilogue warps cost them 3,000 - 5,000 cycles each time);
//       * a cta_group::2 wgrad (M = 256: 128 rows per CTA, N <= 256: B split over the CTAs, K = points) whose fp32
//         accumulator lives in TMEM columns [256, 512) of both CTAs for the whole launch; the pair is bound to ONE work
//         item (kFItems) and consumes its operands of every n-th tile, in tile order (bitwise reproducible sums).  A
//         watcher warp polls the flags of the pair's next 32 tiles (one ld.acquire.gpu round trip) and keeps the count of
//         published tiles in shared memory; the converged TMA warp (lane l issues the l-th 8 KB load of a 64-point
//         stage; 4 stages) issues one fence.proxy.async.global per batch of published tiles, prefetches the stash
//         operands into L2 two tiles ahead and loads the dZ unit ~30 us after its publication: out of L2.
//       * a soft credit throttle keeps the pairs that close: a chain delays the start of a tile by
//         (units published - units picked up - throttle_units) x throttle_gain cycles (bounded: it can delay, never block).
//     HBM sees the forward stash once (read), the dZ write-back, and the ~30 % of dZ reads that still miss L2
//     (profiles/r02_backward_handover.md).  The chain never waits for wgrad (dZ has its full-size buffer), wgrad only
//     waits for flags, all CTAs are resident (grid <= #SMs): no deadlock.
//     Shared memory per CTA: chain weight ring 8 x 8 KB, dZ staging 32 KB, wgrad ring 4 x 32 KB.
// =================================================================================================
constexpr int kFThreads = 512;          // warps 0-7 chain epilogue (0-3 also drain the wgrad accumulator at the end), 8 chain TMA,
                                        // 9 TMEM alloc / chain relay, 10 chain MMA, 11 wgrad TMA, 12 wgrad MMA / relay, 13-14 bias sums,
                                        // 15 dZ bulk stores + publication  (512 threads: 128 registers per thread, no setmaxnreg needed)
constexpr int kFSlots = 8, kFLag = 2;                                     // chain weight ring: two groups of <= 4 slots in flight (12 slots / 3 groups
                                                                          // at the expense of the third wgrad stage measured slower: 1.27 vs 1.19 ms)
constexpr int kFStages = (4 * 64) / kStagePts;        // 128 KB of ring
constexpr uint32_t kFStageBytes = 4 * kHalf;                              // A0 A1 B0 B1: kStagePts points x 64 features each
constexpr uint32_t kFSmemW = 0;
constexpr uint32_t kFSmemStg = kFSmemW + kFSlots * kSlotBytes2;           //  65,536
constexpr uint32_t kFSmemWg = kFSmemStg + 2 * kActChunk;                  //  98,304
constexpr uint32_t kFSmemBytes = kFSmemWg + kFStages * kFStageBytes;      // 229,376
// ---- work items of the wgrad role: every one is a cta_group::2 MMA  D[256 x N] += A^T[256 x 64 points] B[64 points x N]
//      with A = 2 half chunk images per CTA (its 128 M rows) and B = n_b half chunk images per CTA (its N / 2 columns).
//   full   (8 x): A = dZ_l (256 features), B = X_l (256 features), N = 256.
//   L0 / L5 PE part: A = dZ_0 / dZ_5, B = PE(pts) given by BOTH CTAs (N = 128: columns [64, 128) duplicate [0, 64)).
//   views, feature part: TRANSPOSED - A = feature (256), B = d hidden_pre (128 = 64 per CTA), D[m = feature][n = hidden].
//   views, PE(viewdir) part: A = d hidden_pre given by BOTH CTAs (rows [128, 256) duplicate), B = PE(viewdir) by both, N = 128.
struct FItem {
  int slot_item;         // kItems index the partial slot belongs to (geometry of the reduce)
  int flag;              // dZ group the item waits for
  int a_src, b_src;      // 0: dZ stash, 1: forward stash
  int a_chunk[2][2];     // [rank][j]
  int b_chunk[2][2];     // [rank][j], j < n_b
  int n_b, n_mma;        // B half chunks per CTA; MMA N
  int transposed;        // drain: D[m][n] -> part[n * ld + 128 rank + m]  (else part[(128 rank + m) * ld + col0 + n])
  int ld, col0, n_cols;  // slot geometry
  int drain_ranks;       // bit r: CTA r writes its D rows
  int bias_ranks;        // bit r: CTA r writes its 128 bias sums (column sums of ITS A operand); other CTAs write zeros
  int zero_lo, zero_hi;  // columns [zero_lo, zero_hi) of the 128 rows of the slot are zero-filled (the other views item owns them)
  int cost;              // relative cost (pairs are dealt out in proportion)
};
constexpr int kFusedItems = 12;
#ifndef MVIP_SMALL_COST
#define MVIP_SMALL_COST 10     // measured: per tile a small item takes a pair about as long as a full one (operand loads, not MMAs, set the pace)
#endif
#define MVIP_FULL_ITEM(it, fl, a0, b0) {it, fl, 0, 1, {{a0, a0 + 1}, {a0 + 2, a0 + 3}}, {{b0, b0 + 1}, {b0 + 2, b0 + 3}}, 2, 256, 0, 256, 0, 256, 3, 3, 0, 0, 10}
__constant__ FItem kFItems[kFusedItems] = {
    MVIP_FULL_ITEM(1, 1, kDzFeat, 29),             // feature_linear   : d feature^T h8
    MVIP_FULL_ITEM(2, 2, kDzTrunk + 0, 25),        // pts_linears.7
    MVIP_FULL_ITEM(3, 3, kDzTrunk + 4, 21),        // pts_linears.6
    MVIP_FULL_ITEM(4, 4, kDzTrunk + 8, 17),        // pts_linears.5, h part
    MVIP_FULL_ITEM(6, 5, kDzTrunk + 12, 13),       // pts_linears.4
    MVIP_FULL_ITEM(7, 6, kDzTrunk + 16, 9),        // pts_linears.3
    MVIP_FULL_ITEM(8, 7, kDzTrunk + 20, 5),        // pts_linears.2
    MVIP_FULL_ITEM(9, 8, kDzTrunk + 24, 1),        // pts_linears.1
    // pts_linears.5, PE part (no bias: it comes with the h part)
    {5, 4, 0, 1, {{kDzTrunk + 8, kDzTrunk + 9}, {kDzTrunk + 10, kDzTrunk + 11}}, {{0, 0}, {0, 0}}, 1, 128, 0, 64, 0, 64, 3, 0, 0, 0, MVIP_SMALL_COST},
    // pts_linears.0
    {10, 9, 0, 1, {{kDzTrunk + 28, kDzTrunk + 29}, {kDzTrunk + 30, kDzTrunk + 31}}, {{0, 0}, {0, 0}}, 1, 128, 0, 64, 0, 64, 3, 3, 0, 0, MVIP_SMALL_COST},
    // views_linears.0, feature part (transposed)
    {0, 0, 1, 0, {{33, 34}, {35, 36}}, {{kDzHidden, 0}, {kDzHidden + 1, 0}}, 1, 128, 1, 320, 0, 128, 3, 0, 256, 320, MVIP_SMALL_COST},
    // views_linears.0, PE(viewdir) part + bias
    {0, 0, 0, 1, {{kDzHidden, kDzHidden + 1}, {kDzHidden, kDzHidden + 1}}, {{37, 0}, {37, 0}}, 1, 128, 0, 320, 256, 64, 1, 1, 0, 256, MVIP_SMALL_COST},
};
struct FusedPlan {          // host-computed: which item each CTA pair works on, as the k-th of n pairs on that item
  unsigned char item[kFusedSlots], k[kFusedSlots], n[kFusedSlots];
};


__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFThreads, 1)
backward_fused_kernel(const ChainParams p, const WParams wp, const FusedPlan plan) {
  constexpr int kG = kGroupBars2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_gfull[kG], bar_gempty[kG], bar_acc, bar_act, bar_hi;          // chain
  __shared__ uint64_t bar_wfull[kFStages], bar_wempty[kFStages], bar_wacc;              // wgrad
  __shared__ uint64_t bar_sfull[2], bar_sfree[2];                                        // dZ staging buffers: epilogue -> store warp
  __shared__ uint32_t tmem_base_s;
  // Tile pairs are handed out DYNAMICALLY (an atomic counter): with a static round-robin the chains drift apart by several
  // waves (some SM pairs are persistently faster), units are published up to +-40 us out of tile order, and the wgrad pairs,
  // which consume in tile order, wait for the stragglers while everything else piles up.  it_q[k & 3] = the k-th tile pair of
  // this CTA pair, published one iteration ahead by epilogue warp 0 (rank 1 receives it from rank 0 through global memory).
  __shared__ uint32_t it_q[4];
  __shared__ uint64_t bar_it[4];
  __shared__ uint32_t w_ready_s;                                                         // wgrad: leading tiles of this pair's sequence whose dZ unit is published
  volatile uint32_t* w_ready = &w_ready_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int64_t n_pairs = (p.n_tiles + 1) / 2;
  const uint8_t* wT = p.packed + kFwdBytes;

  // wgrad assignment of this pair: the wk-th of wn pairs on fused item wj: tiles wk, wk + wn, ...
  const int wj = plan.item[cluster], wk = plan.k[cluster], wn = plan.n[cluster];
  const FItem& fit = kFItems[wj];
  const bool w_active = wk < p.n_tiles;
  // ===================== wgrad flag watcher (one otherwise idle warp per CTA) =====================
  // Lane l polls the flag of the (ready + l)-th tile of this pair's sequence: one acquire round trip (~1,500 cycles) covers 32
  // tiles and stays off the TMA producer's path.  The count of leading published tiles goes to shared memory.
  auto watch_flags = [&]() {
    if (!w_active) return;
    const int n_my = ((int)p.n_tiles - wk + wn - 1) / wn;
    unsigned long long l_sum = 0, l_max = 0;
    int ready = 0;
    while (ready < n_my) {
      const int idx = ready + lane;
      uint32_t stamp = 0;
      if (idx < n_my) stamp = ld_acquire_gpu(p.flags + (size_t)(wk + idx * wn) * kFlagsPerTile + fit.flag);
      const uint32_t got = __ballot_sync(0xffffffffu, stamp != 0u);
      const int prefix = (got == 0xffffffffu) ? 32 : __ffs(~got) - 1;
      if (prefix > 0) {
        if (lane < prefix) { const uint32_t lag = globaltimer_lo() - (stamp & ~1u); l_sum += lag; l_max = lag > l_max ? lag : l_max; }
        ready += prefix;
        __threadfence_block();
        if (lane == 0) *w_ready = (uint32_t)ready;
      } else {
        __nanosleep(100);
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      l_sum += __shfl_xor_sync(0xffffffffu, l_sum, o);
      const unsigned long long m = __shfl_xor_sync(0xffffffffu, l_max, o);
      l_max = m > l_max ? m : l_max;
    }
    if (rank == 0 && lane == 0) { g_lag[cluster][0] = l_sum; g_lag[cluster][1] = l_max; }
  };

  if (tid == 0) {
    for (int i = 0; i < kG; ++i) {
      mbar_init(&bar_gfull[i], rank == 0 ? 2 : 1);   // leader: own producer + peer relay
      mbar_init(&bar_gempty[i], 1);                  // multicast tcgen05.commit of the chain issuer
    }
    mbar_init(&bar_acc, 1); mbar_init(&bar_act, 16); mbar_init(&bar_hi, 16);
    for (int i = 0; i < kFStages; ++i) {
      mbar_init(&bar_wfull[i], rank == 0 ? 2 : 1);   // leader: own producer + peer relay
      mbar_init(&bar_wempty[i], 1 + 2);              // multicast tcgen05.commit + the two bias-sum warps
    }
    mbar_init(&bar_wacc, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&bar_it[i], 1);
    it_q[0] = (uint32_t)cluster;                     // the first tile pair is static
    w_ready_s = 0;
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_sfull[i], 8); mbar_init(&bar_sfree[i], 1); }
    mbar_fence_init();
    mbar_arrive(&bar_it[0]);                         // it_q[0] is valid
    if (rank == 0) {
      Segment sg;
      sg.item = w_active ? fit.slot_item : -1; sg.t0 = wk; sg.t1 = w_active ? (int)p.n_tiles : wk; sg.stride = wn;
      wp.segs[cluster] = sg;
    }
  }
  if (warp == 9) tmem_alloc_2cta(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  // k-th tile pair of this CTA pair's chain (>= n_pairs: none left)
  auto get_it = [&](uint32_t k) -> int64_t {
    mbar_wait(&bar_it[k & 3u], (k >> 2) & 1u);
    return (int64_t)reinterpret_cast<volatile uint32_t*>(it_q)[k & 3u];
  };

  if (warp < 8) {
    // ===================== chain epilogue: TMEM lane quarter q = warp%4, column half ch of every N-half =====================
    const int q = warp & 3, ch = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t tA = tmem_base + lane_base;
    const uint32_t tD = tA + 128 + ch * 64;
    const float* small = reinterpret_cast<const float*>(p.packed + kSmallOff);
    // staging: one 32 KB buffer (the two chunk images of one accumulator half); this warp writes rows [32 q, 32 q + 32) of image ch
    const uint32_t stg_row = smem_u32(smem + kFSmemStg + ch * kActChunk + q * 4096) + (uint32_t)(lane >> 3) * 1024u + (uint32_t)(lane & 7) * 128u;
    const uint32_t stg_x7 = (uint32_t)(lane & 7) << 4;
    uint32_t acc_phase = 0, so_n = 0;
    auto act_arrive = [&]() {
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&bar_act);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&bar_act), 0));
      }
    };
    auto hi_arrive = [&]() {
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&bar_hi);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&bar_hi), 0));
      }
    };

    // Staggered start: cluster c begins c * stagger cycles late, so that at any time the chains are spread evenly over the
    // nine steps and every wgrad item sees a steady stream of units in tile order (the order its pairs consume them in),
    // instead of a whole wave of units of the same layer at once.
    if (p.stagger > 0) {
      const long long t0 = clock64(), delay = (long long)cluster * p.stagger;
      while (clock64() - t0 < delay) __nanosleep(200);
    }
    for (uint32_t kq = 0;; ++kq) {
      const int64_t it = get_it(kq);
      if (it >= n_pairs) break;
      // the next tile pair: rank 0 draws it now (the atomic's latency hides behind the first chain steps) ...
      uint32_t nxt = 0;
      if (warp == 0 && lane == 0 && rank == 0) nxt = (uint32_t)n_clusters + atom_add_relaxed_gpu(p.next_pair, 1u);
      const int64_t tile = 2 * it + (int64_t)rank;
      const bool tile_valid = tile < p.n_tiles;
      const int64_t g = tile * kTile + r;
      const bool valid = tile_valid && g < p.n_points;
      const uint32_t* masks = reinterpret_cast<const uint32_t*>(p.stash + (size_t)(tile_valid ? tile : 0) * kStashTileBytes + kStashMaskOff);
#ifdef MVIP_TRACE_BWD
      const bool tr_on = cluster == 0 && rank == 0 && warp == 0 && lane == 0 && kq == 3;
      int tr_hs = 0;
#endif

      // This warp's 32 packed rows -> its 4 KB piece of the current staging buffer; the store warp sends the whole 32 KB
      // buffer (two chunk images) to the dZ stash as ONE bulk store and publishes the groups.  No global-memory operation, no
      // bulk-group wait and no release fence on the epilogue warps (a red.release.gpu here cost 3,000 - 5,000 cycles).
      auto stage_out = [&](const uint32_t (&pk)[32]) {
        const uint32_t sb = so_n & 1u;
        // ONE 32 KB staging buffer: the previous half-step's bulk store (issued ~2,900 cycles ago, ~1,300 cycles to read its
        // source) has finished with it.  The 32 KB a second buffer took are a fourth stage of the wgrad ring.
        if (so_n > 0) mbar_wait(&bar_sfree[sb ^ 1u], ((so_n - 1) >> 1) & 1u);
        ++so_n;
        BTR(tr_hs, 7);
#pragma unroll
        for (int gq = 0; gq < 8; ++gq)     // chunk_off16(lane, gq) = row offset | ((gq ^ (lane & 7)) << 4)
          sts128(stg_row | (((uint32_t)gq << 4) ^ stg_x7), pk[4 * gq], pk[4 * gq + 1], pk[4 * gq + 2], pk[4 * gq + 3]);
        BTR(tr_hs, 8);
        fence_proxy_async_smem();
        __syncwarp();
        BTR(tr_hs, 9);
        if (lane == 0) mbar_arrive(&bar_sfull[sb]);
        BTR(tr_hs, 10);
      };
      // ReLU mask words of (step s, half h), loaded one accumulator half ahead: they come from HBM
      auto load_mask = [&](int s, int h) -> uint2 {
        if (!valid) return make_uint2(0u, 0u);
        if (s < 1) return make_uint2(~0u, ~0u);
        return __ldg(reinterpret_cast<const uint2*>(masks + mask_word_index(8 - s, r, h, ch)));
      };

      uint32_t pk[32];
      const float4 dr = valid ? __ldg(p.d_raw + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      {  // ---- input stage: d hidden_pre = (W_rgb^T d_rgb) * [hidden > 0] -> A columns [0,64) (K = 128) of step 0
        const uint2 mw = valid ? __ldg(reinterpret_cast<const uint2*>(masks + mask_word_index(8, r, 0, ch))) : make_uint2(0u, 0u);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int c0 = ch * 64 + 16 * b;
          float v[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(small + kSmWRgb + c0 + 4 * j4));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(small + kSmWRgb + 128 + c0 + 4 * j4));
            const float4 w2 = __ldg(reinterpret_cast<const float4*>(small + kSmWRgb + 256 + c0 + 4 * j4));
            v[4 * j4 + 0] = dr.x * w0.x + dr.y * w1.x + dr.z * w2.x;
            v[4 * j4 + 1] = dr.x * w0.y + dr.y * w1.y + dr.z * w2.y;
            v[4 * j4 + 2] = dr.x * w0.z + dr.y * w1.z + dr.z * w2.z;
            v[4 * j4 + 3] = dr.x * w0.w + dr.y * w1.w + dr.z * w2.w;
          }
          const uint32_t m = (b < 2 ? mw.x : mw.y) >> (8 * (b & 1));
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[8 * b + i] = mask_bf16x2(pack_bf16x2(v[2 * i], v[2 * i + 1]), (m >> i) & 0x00010001u);
        }
        tmem_st32(tA + ch * 32, pk);            // the previous tile's last MMAs are complete (its last accumulator was drained)
        tmem_st_wait();
        tc_fence_before();
        act_arrive();                           // step 0 may start
        stage_out(pk);
      }

      uint2 mw_next = load_mask(0, 0);
#pragma unroll 1
      for (int s = 0; s < kCSteps; ++s) {
        // ... and publishes it half-way through the tile: to the peer CTA through global memory (tagged with the iteration
        // number), to the other warps of this CTA through it_q / bar_it
        if (warp == 0 && lane == 0 && ((rank == 0 && s == 4) || (rank == 1 && s == 6))) {
          uint32_t* slot = p.next_pair + 16 + 4 * cluster + ((kq + 1) & 3u);
          if (rank == 0) {
            if (nxt > 0xFFFFFu) nxt = 0xFFFFFu;                  // (>= n_pairs either way: n_pairs < 2^20 is checked on the host)
            st_relaxed_gpu(slot, ((kq + 1) << 20) | nxt);
          } else {
            uint32_t v = ld_relaxed_gpu(slot);
            const long long t0 = clock64();
            while ((v >> 20) != ((kq + 1) & 0xFFFu)) {
              if (clock64() - t0 > 8000000000LL) { printf("mvip: fused backward tile hand-out timeout cluster %d\n", cluster); __trap(); }
              v = ld_relaxed_gpu(slot);
            }
            nxt = v & 0xFFFFFu;
          }
          reinterpret_cast<volatile uint32_t*>(it_q)[(kq + 1) & 3u] = nxt;
          mbar_arrive(&bar_it[(kq + 1) & 3u]);
        }
        // s == 0: d feature (no activation).  s >= 1: dZ_{8-s} = acc [+ d_alpha * w_alpha] masked by h_{9-s} > 0
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const uint2 mw = mw_next;
          mw_next = (h == 0) ? load_mask(s, 1) : load_mask(s + 1 < kCSteps ? s + 1 : 0, 0);
          {
            uint32_t raw[4][16];
            BTR(tr_hs, 0);
            mbar_wait(&bar_acc, acc_phase);
            BTR(tr_hs, 1);
            acc_phase ^= 1;
            tc_fence_after();
            if (h == 1 && s < kCSteps - 1) tmem_st32(tA + ch * 32, pk);      // half 0 of dZ -> next A operand, K columns [0,128)
            load_half(tD, raw);
            if (h == 1 && s < kCSteps - 1) tmem_st_wait();
            BTR(tr_hs, 2);
            tc_fence_before();
            if (h == 0 || s < kCSteps - 1) act_arrive();   // h = 0: accumulator drained; h = 1: A[0,128) ready + accumulator drained
            BTR(tr_hs, 3);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              float v[16];
              if (s == 1) {
                const int c0 = h * 128 + ch * 64 + 16 * b;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 w = __ldg(reinterpret_cast<const float4*>(small + kSmWAlpha + c0 + 4 * j4));
                  v[4 * j4 + 0] = __uint_as_float(raw[b][4 * j4 + 0]) + dr.w * w.x;
                  v[4 * j4 + 1] = __uint_as_float(raw[b][4 * j4 + 1]) + dr.w * w.y;
                  v[4 * j4 + 2] = __uint_as_float(raw[b][4 * j4 + 2]) + dr.w * w.z;
                  v[4 * j4 + 3] = __uint_as_float(raw[b][4 * j4 + 3]) + dr.w * w.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[b][j]);
              }
              const uint32_t m = (b < 2 ? mw.x : mw.y) >> (8 * (b & 1));
#pragma unroll
              for (int i = 0; i < 8; ++i) pk[8 * b + i] = mask_bf16x2(pack_bf16x2(v[2 * i], v[2 * i + 1]), (m >> i) & 0x00010001u);
            }
          }
          BTR(tr_hs, 4);
          if (h == 1 && s < kCSteps - 1) {
            tmem_st32(tA + 64 + ch * 32, pk);
            tmem_st_wait();
            tc_fence_before();
            hi_arrive();                        // A[128,256) ready
          }
          BTR(tr_hs, 5);
          stage_out(pk);
          BTR(tr_hs, 6);
#ifdef MVIP_TRACE_BWD
          ++tr_hs;
#endif
        }
      }
    }
    if (warp < 4 && w_active) {
      // drain of the wgrad accumulator: TMEM lane i of CTA `rank` = M row 128 rank + i, columns = N
      mbar_wait(&bar_wacc, 0);
      tc_fence_after();
      float* part = wp.partials + (size_t)cluster * (kPartialSlotBytes / sizeof(float));
      if ((fit.drain_ranks >> rank) & 1) {
        for (int c0 = 0; c0 < fit.n_cols; c0 += 32) {
          uint32_t acc[32];
          tmem_ld32(tmem_base + lane_base + 256 + c0, acc);
          tmem_ld_wait();
          if (fit.transposed) {       // D[m][n] -> part[n][128 rank + m]: consecutive lanes write consecutive floats
#pragma unroll
            for (int j = 0; j < 32; ++j) part[(size_t)(c0 + j) * fit.ld + 128 * rank + r] = __uint_as_float(acc[j]);
          } else {
            float* prow = part + (size_t)(128 * rank + r) * fit.ld + fit.col0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(prow + j) = make_float4(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1]),
                                                                 __uint_as_float(acc[j + 2]), __uint_as_float(acc[j + 3]));
          }
        }
      }
      if (fit.zero_hi > fit.zero_lo) {   // the two views items share kItems[0]'s slot geometry: each zero-fills the other's columns
        const int row = 64 * (int)rank + (r & 63), half = r >> 6, w = (fit.zero_hi - fit.zero_lo) / 2;
        float* z = part + (size_t)row * fit.ld + fit.zero_lo + half * w;
        for (int j = 0; j < w; j += 4) *reinterpret_cast<float4*>(z + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  } else if (warp == 8) {
    // ===================== chain TMA producer: this CTA's 64 rows of every (step, N-half, K chunk), once per tile pair =====================
    if (lane == 0) {
      uint32_t j = 0; int slot = 0;
      for (uint32_t kq = 0; get_it(kq) < n_pairs; ++kq) {
        int cbase = 0;
        for (int s = 0; s < kCSteps; ++s) {
          const int n = chain_nchunks(s);
          for (int h = 0; h < 2; ++h, ++j) {
            if (j >= (uint32_t)kFLag) mbar_wait(&bar_gempty[(j - kFLag) % kG], ((j - kFLag) / kG) & 1u);
            uint64_t* full = &bar_gfull[j % kG];
            mbar_arrive_expect_tx(full, (uint32_t)n * kSlotBytes2);
            for (int ci = 0; ci < n; ++ci) {
              const uint8_t* src = wT + (size_t)(cbase + ci) * kW256 + (size_t)(2 * h + (int)rank) * kSlotBytes2;
              tma_load_1d(smem + kFSmemW + slot * kSlotBytes2, src, kSlotBytes2, full);
              if (++slot == kFSlots) slot = 0;
            }
          }
          cbase += n;
        }
      }
    }
  } else if (warp == 9) {
    // ===================== chain relay (peer CTA): "my part of the weight group has landed" =====================
    if (rank == 1 && lane == 0) {
      uint32_t j = 0;
      for (uint32_t kq = 0; get_it(kq) < n_pairs; ++kq) {
        for (int c = 0; c < 2 * kCSteps; ++c, ++j) {
          mbar_wait(&bar_gfull[j % kG], (j / kG) & 1u);
          mbar_arrive_cluster(mapa_u32(smem_u32(&bar_gfull[j % kG]), 0));
        }
      }
    } else if (rank == 0) {
      watch_flags();
    }
  } else if (warp == 10) {
    // ===================== chain MMA issuer (leader CTA) =====================
    if (rank == 0) {
      uint32_t act_phase = 0, hi_phase = 0, j = 0;
      long long c_t0 = clock64(), c_act = 0, c_full = 0, c_x = 0;
      const uint32_t idesc = umma_idesc_bf16(256, 128, 0, 0);
      const uint32_t tA = tmem_base;
      const uint32_t tDm = tA + 128;
      const uint32_t w_lo = desc_lo2(smem_u32(smem) + kFSmemW);
      const uint32_t w_end = w_lo + kFSlots * (kSlotBytes2 >> 4);
      uint32_t bpos = w_lo;
      auto next_slot = [&](uint32_t b) { b += (kSlotBytes2 >> 4); return b == w_end ? w_lo : b; };
      for (uint32_t kq = 0; get_it(kq) < n_pairs; ++kq) {
#pragma unroll 1
        for (int s = 0; s < kCSteps; ++s) {
#pragma unroll 1
          for (int h = 0; h < 2; ++h, ++j) {
            const uint32_t b1 = bpos, b2 = next_slot(b1), b3 = next_slot(b2), b4 = next_slot(b3);
            bpos = (s == 0) ? b3 : next_slot(b4);
            uint64_t* gempty = &bar_gempty[j % kG];
            c_x = clock64();
            mbar_wait(&bar_gfull[j % kG], (j / kG) & 1u);
            c_full += clock64() - c_x; c_x = clock64();
            mbar_wait(&bar_act, act_phase);
            c_act += clock64() - c_x;
            act_phase ^= 1;
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma2_ts(tDm, tA + kk * 8, b1 + kk * 2, idesc, kk > 0 ? 1u : 0u);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) mma2_ts(tDm, tA + 32 + kk * 8, b2 + kk * 2, idesc, 1u);
              if (s == 0) {
                umma_commit_2cta(gempty, 3);
                umma_commit_2cta(&bar_acc, 3);
              }
            }
            __syncwarp();
            if (s > 0) {
              if (h == 0) {
                c_x = clock64();
                mbar_wait(&bar_hi, hi_phase);
                c_act += clock64() - c_x;
                hi_phase ^= 1;
                tc_fence_after();
              }
              if (elect_one_sync()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma2_ts(tDm, tA + 64 + kk * 8, b3 + kk * 2, idesc, 1u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma2_ts(tDm, tA + 96 + kk * 8, b4 + kk * 2, idesc, 1u);
                umma_commit_2cta(gempty, 3);
                umma_commit_2cta(&bar_acc, 3);
              }
              __syncwarp();
            }
          }
        }
      }
      if (cluster == MVIP_PROF_CLUSTER && lane == 0) { g_wprof[0] = clock64() - c_t0; g_wprof[1] = c_act; g_wprof[2] = c_full; }
    } else {
      watch_flags();
    }
  } else if (warp == 11) {
    // ===================== wgrad TMA producer (both CTAs): this CTA's 128 dZ features and 128 input features of a 64-point stage =====================
    // The whole warp runs converged: lane l < 2 + n_b issues the l-th 8 KB load of a stage, so a stage costs ONE issue slot
    // instead of four to six (a single thread needs ~150 cycles per bulk-copy instruction; with the flag poll on the same
    // thread the producer took 3,900 cycles per tile against 1,400 for the MMAs and the pairs fell ever further behind the
    // chains, every dZ byte coming back from HBM).  The flags are watched by another warp (below): this one polls a counter
    // in shared memory.
    if (w_active) {
      const uint64_t pol = l2_policy_evict_first();
      const int n_loads = 2 + fit.n_b;
      const uint32_t stage_tx = (uint32_t)n_loads * kHalf;
      const bool loader = lane < n_loads;
      const int lj = lane < 2 ? lane : lane - 2;
      const bool from_stash = loader && (lane < 2 ? fit.a_src : fit.b_src);
      const int my_chunk = loader ? (lane < 2 ? fit.a_chunk[rank][lj] : fit.b_chunk[rank][lj]) : 0;
      const int n_my = ((int)p.n_tiles - wk + wn - 1) / wn;
      int stage = 0; uint32_t phase = 0;
      long long p_flag = 0, p_empty = 0, p_t0 = clock64();
      unsigned long long l_ready = 0;
      uint32_t fenced = 0;
      // the forward-stash operands of a tile do not depend on the chain: they are pulled into L2 two tiles ahead
      if (from_stash) {
        tma_prefetch_l2(wp.stash + (size_t)wk * kStashTileBytes + (size_t)my_chunk * kActChunk, kActChunk);
        if (n_my > 1) tma_prefetch_l2(wp.stash + (size_t)(wk + wn) * kStashTileBytes + (size_t)my_chunk * kActChunk, kActChunk);
      }
      for (int i = 0; i < n_my; ++i) {
        const int t = wk + i * wn;
        const uint8_t* src = (from_stash ? wp.stash + (size_t)t * kStashTileBytes : wp.dz + (size_t)t * kDzTileBytes) + (size_t)my_chunk * kActChunk;
        if (from_stash && i + 2 < n_my) tma_prefetch_l2(src + (size_t)2 * wn * kStashTileBytes, kActChunk);
        {  // the chain that owns tile t (some SM of this launch) has published this item's dZ unit
          long long t0 = clock64();
          if ((uint32_t)i >= fenced) {                 // one proxy fence covers every tile the watcher had reported before it
            uint32_t r = *w_ready;
            if (r > (uint32_t)i) ++l_ready;
            while (r <= (uint32_t)i) {
              if (clock64() - t0 > 8000000000LL) { if (lane == 0) printf("mvip: fused wgrad flag timeout cluster %d tile %d flag %d\n", cluster, t, fit.flag); __trap(); }
              r = *w_ready;
            }
            fence_proxy_async_global();
            fenced = r;
          } else {
            ++l_ready;
          }
          p_flag += clock64() - t0;
          if (rank == 0 && lane == 0) {
            p.cons_stamp[(size_t)t * kFlagsPerTile + fit.flag] = globaltimer_lo();
            red_add_relaxed_gpu(p.credit + 1, (fit.flag == 0 || fit.flag == 4) ? 1u : 2u);
          }
        }
        for (int h = 0; h < kSubStages; ++h) {
          uint8_t* sbase = smem + kFSmemWg + stage * kFStageBytes;
          long long te = clock64();
          mbar_wait(&bar_wempty[stage], phase ^ 1);
          p_empty += clock64() - te;
          if (lane == 0) mbar_arrive_expect_tx(&bar_wfull[stage], stage_tx);
          __syncwarp();
          if (loader) tma_load_1d_hint(sbase + lane * kHalf, src + h * kHalf, kHalf, &bar_wfull[stage], pol);
          if (++stage == kFStages) { stage = 0; phase ^= 1; }
        }
      }
      if (cluster == MVIP_PROF_CLUSTER && rank == 0 && lane == 0) { g_wprof[5] = p_flag; g_wprof[6] = p_empty; g_wprof[7] = clock64() - p_t0; }
      if (rank == 0 && lane == 0) { g_lag[cluster][2] = n_my; g_lag[cluster][3] = l_ready; }
    }
  } else if (warp == 12) {
    // ===================== wgrad MMA issuer (leader) / relay (peer) =====================
    if (w_active) {
      int stage = 0; uint32_t phase = 0;
      if (rank == 0) {
        const uint32_t idesc = umma_idesc_bf16(256, fit.n_mma, 1, 1);
        const uint32_t d0 = tmem_base + 256;
        bool first = true;
        long long w_t0 = clock64(), w_full = 0;
        for (int t = wk; t < p.n_tiles; t += wn) {
          for (int h = 0; h < kSubStages; ++h) {
            const uint32_t sbase = smem_u32(smem) + kFSmemWg + stage * kFStageBytes;
            const uint32_t a_lo = desc_lo_mn(sbase, kHalf), b_lo = desc_lo_mn(sbase + 2 * kHalf, kHalf);
            long long tw = clock64();
            mbar_wait(&bar_wfull[stage], phase);
            w_full += clock64() - tw;
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
              for (int ks = 0; ks < kStagePts / 16; ++ks) {
                const uint32_t ko = (uint32_t)ks * (2048u >> 4);
                mma2_ss(d0, a_lo + ko, b_lo + ko, idesc, (first && ks == 0) ? 0u : 1u);
              }
              umma_commit_2cta(&bar_wempty[stage], 3);
            }
            __syncwarp();
            first = false;
            if (++stage == kFStages) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one_sync()) umma_commit_2cta(&bar_wacc, 3);
        __syncwarp();
        if (cluster == MVIP_PROF_CLUSTER && lane == 0) { g_wprof[3] = clock64() - w_t0; g_wprof[4] = w_full; }
      } else if (lane == 0) {
        for (int t = wk; t < p.n_tiles; t += wn) {
          for (int h = 0; h < kSubStages; ++h) {
            mbar_wait(&bar_wfull[stage], phase);
            mbar_arrive_cluster(mapa_u32(smem_u32(&bar_wfull[stage]), 0));
            if (++stage == kFStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if ((warp == 13 || warp == 14) && w_active) {
    // ===================== bias column sums of this CTA's 128 dZ features, from the staged A operand =====================
    // thread t4 (0..63): 8-feature group fg = t4 % 16 (piece fg / 8, 16-byte group fg % 8), rows [16 rg, 16 rg + 16) of the stage with
    // rg = t4 / 16: sixteen 16-byte loads per stage (the first version read 64 words per thread and took as long as the four MMAs
    // of a full item - longer than those of a small one - before it released the stage)
    static_assert(kStagePts == 64, "the bias warps assume 64-point stages");
    const int t4 = (warp - 13) * 32 + lane;
    const int fg = t4 & 15, rg = t4 >> 4;
    const uint32_t boff = (uint32_t)(fg >> 3) * kHalf + (uint32_t)rg * 16u * 128u;
    const uint32_t g7 = (uint32_t)(fg & 7);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool do_bias = (fit.bias_ranks >> rank) & 1;
    int stage = 0; uint32_t phase = 0;
    for (int t = wk; t < p.n_tiles; t += wn) {
      for (int h = 0; h < kSubStages; ++h) {
        mbar_wait(&bar_wfull[stage], phase);
        if (do_bias) {
          const uint32_t a32 = smem_u32(smem) + kFSmemWg + stage * kFStageBytes + boff;
#pragma unroll
          for (int i = 0; i < 16; ++i) {          // row 16 rg + i: (row & 7) = i & 7
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(a32 + (uint32_t)i * 128u + ((g7 ^ (uint32_t)(i & 7)) << 4)));
            acc[0] += __uint_as_float(w0 << 16); acc[1] += __uint_as_float(w0 & 0xffff0000u);
            acc[2] += __uint_as_float(w1 << 16); acc[3] += __uint_as_float(w1 & 0xffff0000u);
            acc[4] += __uint_as_float(w2 << 16); acc[5] += __uint_as_float(w2 & 0xffff0000u);
            acc[6] += __uint_as_float(w3 << 16); acc[7] += __uint_as_float(w3 & 0xffff0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_wempty[stage]);
        if (++stage == kFStages) { stage = 0; phase ^= 1; }
      }
    }
    // the two row groups of a warp meet by shuffle; the two warps write separate partials (the reduce kernel adds them)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    if (lane < 16) {
      float* bias_out = wp.bias_partials + (size_t)cluster * 512 + (size_t)(warp - 13) * 256 + 128 * rank + 8 * fg;     // zeros where this CTA contributes nothing
#pragma unroll
      for (int e = 0; e < 8; ++e) bias_out[e] = acc[e];
    }
  } else if (warp == 15) {
    // ===================== dZ store warp: one 32 KB bulk store per staging buffer; publication of every dZ unit =====================
    // A unit (the output of one chain step of one tile: flag 0 = d hidden_pre, 1 + s = step s) is published as soon as its
    // stores are complete: the wgrad pair that consumes it picks it up out of L2 a few microseconds later, long before the
    // line would be evicted (measured with scripts/ubench/l2_handoff.cu: bulk-stored data is read back from L2 as long as
    // less than ~70 MB are written between the store and the load; the first version published three times per tile and its
    // TMA thread needed 3,900 cycles per tile: the pairs fell ever further behind and every dZ byte was re-read from HBM).
    // The store of buffer c is given until buffer c + 2 has been issued to complete (cp.async.bulk.wait_group 2), then the
    // thread orders it (async proxy) before its generic-proxy flag store and releases the flag at gpu scope.
    // Only this thread ever pays for the fences.
    if (lane == 0) {
      uint32_t n = 0;
      uint32_t* pend[2] = {nullptr, nullptr};       // flags of the units whose last store is the bulk group n - 1 / n - 2
      auto publish = [&](uint32_t*& f) {
        if (f) { fence_proxy_async_global(); st_release_gpu(f, globaltimer_lo() | 1u); red_add_relaxed_gpu(p.credit, 2u); f = nullptr; }   // non-zero; the value is a time stamp (debug)
      };
      for (uint32_t kq = 0;; ++kq) {
        const int64_t it = get_it(kq);
        if (it >= n_pairs) break;
        const int64_t tile = 2 * it + (int64_t)rank;
        const bool tile_valid = tile < p.n_tiles;
        uint8_t* dz_tile = p.dz + (size_t)(tile_valid ? tile : 0) * kDzTileBytes;
        uint32_t* tile_flags = p.flags + (size_t)(tile_valid ? tile : 0) * kFlagsPerTile;
        for (int c = 0; c < 1 + 2 * kCSteps; ++c, ++n) {
          // c == 0: d hidden_pre; c = 1 + 2 s + h: N-half h of chain step s
          const int s = (c - 1) >> 1, h = (c - 1) & 1;
          const int chunk = (c == 0) ? kDzHidden : ((s == 0 ? kDzFeat : kDzTrunk + 4 * (s - 1)) + 2 * h);
          const uint32_t sb = n & 1u;
          if (c == 0 && p.throttle_cycles > 0) {
            // soft back-pressure: hold the tile back (for a bounded time) while the wgrad pairs are more than ~50 MB behind,
            // so that they read dZ out of L2 and stay fast enough to keep up
            const long long t0 = clock64();
            if (p.throttle_gain > 0) {        // proportional: one look, a delay that grows with the excess (no stop-and-go waves)
              const int excess = (int)(ld_relaxed_gpu(p.credit) - ld_relaxed_gpu(p.credit + 1)) / 2 - p.throttle_units;
              long long delay = (long long)excess * p.throttle_gain;
              if (delay > p.throttle_cycles) delay = p.throttle_cycles;
              while (clock64() - t0 < delay) __nanosleep(200);
            } else {
              while ((int)(ld_relaxed_gpu(p.credit) - ld_relaxed_gpu(p.credit + 1)) > 2 * p.throttle_units && clock64() - t0 < p.throttle_cycles) __nanosleep(500);
            }
          }
          mbar_wait(&bar_sfull[sb], (n >> 1) & 1u);
          if (tile_valid) tma_store_1d(dz_tile + (size_t)chunk * kActChunk, smem + kFSmemStg, 2 * kActChunk);   // (an evict_last hint: 1.236 -> 1.244 ms)
          tma_store_commit();
          tma_store_wait_read0();
          mbar_arrive(&bar_sfree[sb]);
          tma_store_wait_all2();                    // every store but the last two is complete
          publish(pend[1]);
          pend[1] = pend[0];
          pend[0] = (tile_valid && (c == 0 || h == 1)) ? tile_flags + (c == 0 ? 0 : 1 + s) : nullptr;
        }
      }
      tma_store_wait_all0();
      publish(pend[1]);
      publish(pend[0]);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) tmem_dealloc_2cta(tmem_base, 512);
}

// =================================================================================================
// 3. alpha / rgb head grads on CUDA cores
// =================================================================================================
constexpr int kHeadFloats = 256 + 4 + 384 + 4;   // dW_alpha, db_alpha(+pad), dW_rgb, db_rgb(+pad)
constexpr int kHeadMaxBlocks = 1184;             // partial slots in the workspace (one per block; the grid is one block per SM)
constexpr int kHgRows = 64;                      // points per pipeline stage (half a tile)
constexpr int kHgStages = 4;
constexpr uint32_t kHgPiece = kHgRows * 128;     // 64 rows of one chunk image (rows 0-63 / 64-127 of a chunk are contiguous)
constexpr uint32_t kHgStageBytes = 6 * kHgPiece + kHgRows * 16;   // h8 (4 chunks) + hidden (2 chunks) + d_raw
constexpr int kHgThreads = 288;                  // 8 consumer warps + 1 bulk-copy warp
constexpr size_t kHgSmemBytes = (size_t)kHgStages * kHgStageBytes + 1024;

__device__ __forceinline__ void unpack_bf16x8(const uint4 v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

// dW_alpha[j] = sum_p d_alpha[p] h8[p][j];  dW_rgb[c][j] = sum_p d_rgb[p][c] hidden[p][j];  db = sum_p d_raw[p].
// One block per SM walks half tiles (64 points) in a fixed order.  Warp 8 streams the operands of a half tile - the h8 and
// hidden rows of the stash (six 8 KB pieces) and 1 KB of d_raw - into a four-stage shared-memory ring with bulk copies
// (evict-first: this is the last reader of the stash); consumer warp w owns rows [8w, 8w+8) of every stage.  h8: lane l
// holds the 8 features [8l, 8l+8) of a row; hidden (128 features): lane l holds features [8(l&15), +8) of the rows of
// parity l>>4, the two half warps are added at the end.  Fixed summation order => bitwise reproducible.
__global__ void __launch_bounds__(kHgThreads, 1) head_grads_kernel(const float4* __restrict__ d_raw, const uint8_t* __restrict__ stash,
                                                                int64_t n_points, int64_t n_half, float* __restrict__ out) {
  extern __shared__ uint8_t hg_smem_raw[];
  __shared__ float red[8][kHeadFloats];
  __shared__ uint64_t full[kHgStages], empty[kHgStages];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hg_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kHgStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == 8) {
    const uint64_t pol = l2_policy_evict_first();
    int it = 0;
    for (int64_t h = blockIdx.x; h < n_half; h += gridDim.x, ++it) {
      const int s = it % kHgStages;
      if (it >= kHgStages) mbar_wait(&empty[s], ((it / kHgStages) & 1) ^ 1);
      uint8_t* dst = ring + (size_t)s * kHgStageBytes;
      const uint8_t* tile = stash + (size_t)(h >> 1) * kStashTileBytes + (size_t)(h & 1) * kHgPiece;
      const int64_t left = n_points - h * kHgRows;                       // >= 1: n_half = ceil(n_points / 64)
      const uint32_t d_bytes = (uint32_t)(left < kHgRows ? left : kHgRows) * 16u;
      if (lane == 0) mbar_arrive_expect_tx(&full[s], 6 * kHgPiece + d_bytes);
      __syncwarp();
      if (lane < 4) tma_load_1d_hint(dst + lane * kHgPiece, tile + (size_t)(kStashH + 28 + lane) * kActChunk, kHgPiece, &full[s], pol);
      else if (lane < 6) tma_load_1d_hint(dst + lane * kHgPiece, tile + (size_t)(kStashHidden + lane - 4) * kActChunk, kHgPiece, &full[s], pol);
      else if (lane == 6) tma_load_1d(dst + 6 * kHgPiece, d_raw + h * kHgRows, d_bytes, &full[s]);
    }
  } else {
    float aa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float ar[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ag[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ab[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t h8_piece = (uint32_t)(lane >> 3) * kHgPiece, hid_piece = (uint32_t)(4 + ((lane & 15) >> 3)) * kHgPiece;
    const int g8 = lane & 7, par = lane >> 4;
    int it = 0;
    for (int64_t h = blockIdx.x; h < n_half; h += gridDim.x, ++it) {
      const int s = it % kHgStages;
      mbar_wait(&full[s], (it / kHgStages) & 1);
      const uint8_t* st = ring + (size_t)s * kHgStageBytes;
      const float4* dsm = reinterpret_cast<const float4*>(st + 6 * kHgPiece);
      const int64_t left = n_points - h * kHgRows;
      const int valid = (int)(left < kHgRows ? left : kHgRows);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 8 + i;
        const float4 d = row < valid ? dsm[row] : make_float4(0.f, 0.f, 0.f, 0.f);
        float f[8];
        unpack_bf16x8(*reinterpret_cast<const uint4*>(st + h8_piece + chunk_off16(row, g8)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) aa[e] += d.w * f[e];
        ds.x += d.x; ds.y += d.y; ds.z += d.z; ds.w += d.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = warp * 8 + 2 * i + par;
        const float4 d = row < valid ? dsm[row] : make_float4(0.f, 0.f, 0.f, 0.f);
        float f[8];
        unpack_bf16x8(*reinterpret_cast<const uint4*>(st + hid_piece + chunk_off16(row, g8)), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { ar[e] += d.x * f[e]; ag[e] += d.y * f[e]; ab[e] += d.z * f[e]; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ar[e] += __shfl_xor_sync(FULL_MASK, ar[e], 16);
      ag[e] += __shfl_xor_sync(FULL_MASK, ag[e], 16);
      ab[e] += __shfl_xor_sync(FULL_MASK, ab[e], 16);
      red[warp][lane * 8 + e] = aa[e];
      if (lane < 16) {
        red[warp][260 + lane * 8 + e] = ar[e];
        red[warp][260 + 128 + lane * 8 + e] = ag[e];
        red[warp][260 + 256 + lane * 8 + e] = ab[e];
      }
    }
    if (lane == 31) { red[warp][256] = ds.w; red[warp][644] = ds.x; red[warp][645] = ds.y; red[warp][646] = ds.z; }
    if (lane == 30) { red[warp][257] = red[warp][258] = red[warp][259] = 0.f; red[warp][647] = 0.f; }
  }
  __syncthreads();
  float* o = out + (size_t)blockIdx.x * kHeadFloats;
  for (int e = tid; e < kHeadFloats; e += kHgThreads) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][e];
    o[e] = s;
  }
}

// =================================================================================================
// 4. deterministic reduction of all partials into the gradient tensors
// =================================================================================================
struct ReduceParams {
  const float* partials;
  const float* bias_partials;
  const Segment* segs;
  int n_slots;            // partial slots in use (one per CTA pair of the fused backward)
  const float* head_partials;
  int head_grid;
  GradPtrs grads;
  int accumulate;
};

constexpr int kReduceGridX = 80;     // 256 x 320 floats / 4 per thread / 256 threads

// Fixed summation order everywhere (slot order; for the heads: stripe-of-8 order, then stripe order) => bitwise
// reproducible gradients.  Every thread keeps 8 independent 16-byte (weights) / 4-byte (heads) loads in flight: the
// first version walked ~30 slots (weights) and up to 1184 partials (heads) with a dependent add per load and took
// 70 us per launch at 0.5 TB/s.
__global__ void __launch_bounds__(256) reduce_kernel(const ReduceParams p) {
  __shared__ int slots[kFusedSlots];
  __shared__ int nslots;
  __shared__ float hred[8][32];
  const int item = blockIdx.y;
  if (item == kRealItems) {
    // heads: alpha_linear / rgb_linear.  Block x owns elements [32x, 32x+32); warp w sums partials w, w+8, ... (lane = element)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e0 = blockIdx.x * 32; e0 < kHeadFloats; e0 += gridDim.x * 32) {
      const int e = e0 + lane;
      float s = 0.f;
      if (e < kHeadFloats) {
        int b = warp;
        for (; b + 56 < p.head_grid; b += 64) {
          float a[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) a[k] = p.head_partials[(size_t)(b + 8 * k) * kHeadFloats + e];
#pragma unroll
          for (int k = 0; k < 8; ++k) s += a[k];
        }
        for (; b < p.head_grid; b += 8) s += p.head_partials[(size_t)b * kHeadFloats + e];
      }
      hred[warp][lane] = s;
      __syncthreads();
      if (warp == 0 && e < kHeadFloats) {
        float t = hred[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) t += hred[k][lane];
        float* dst = nullptr;
        if (e < 256) dst = p.grads.p[kPAlphaW] + e;
        else if (e == 256) dst = p.grads.p[kPAlphaB];
        else if (e >= 260 && e < 644) dst = p.grads.p[kPRgbW] + (e - 260);
        else if (e >= 644 && e < 647) dst = p.grads.p[kPRgbB] + (e - 644);
        if (dst) *dst = p.accumulate ? *dst + t : t;
      }
      __syncthreads();
    }
    return;
  }
  // slots of this item, in slot order: the segment table is read by all threads at once (one thread walking the 296
  // entries in global memory was 15 us of latency), then compacted from shared memory
  __shared__ unsigned char mine[kFusedSlots];
  for (int i = threadIdx.x; i < p.n_slots; i += blockDim.x) {
    const Segment sg = p.segs[i];
    mine[i] = (sg.item == item && sg.t1 > sg.t0) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < p.n_slots; ++i)
      if (mine[i]) slots[n++] = i;
    nslots = n;
  }
  __syncthreads();
  const WItem& itm = kItems[item];
  const int rows = 128 * itm.m_blocks, ntot = 64 * itm.nb;
  const size_t slot_floats = kPartialSlotBytes / sizeof(float);
  for (int e = 4 * (blockIdx.x * blockDim.x + threadIdx.x); e < rows * ntot; e += 4 * gridDim.x * blockDim.x) {
    const int row = e / ntot, n = e % ntot;     // ntot is a multiple of 64: the four elements share a B chunk
    const int j = n >> 6, w = n & 63;
    if (w >= itm.valid[j]) continue;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = 0;
    for (; k + 8 <= nslots; k += 8) {   // eight loads in flight, summed in slot order (deterministic)
      float4 a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = *reinterpret_cast<const float4*>(p.partials + (size_t)slots[k + q] * slot_floats + e);
#pragma unroll
      for (int q = 0; q < 8; ++q) { s.x += a[q].x; s.y += a[q].y; s.z += a[q].z; s.w += a[q].w; }
    }
    for (; k < nslots; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(p.partials + (size_t)slots[k] * slot_floats + e);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    float* dst = p.grads.p[itm.dst] + (size_t)row * itm.ld + itm.col[j] + w;   // ld = 63 / 283 / 319: scalar stores
    const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (w + q < itm.valid[j]) dst[q] = p.accumulate ? dst[q] + v[q] : v[q];
  }
  if (itm.bias >= 0 && blockIdx.x == 0) {
    for (int e = threadIdx.x; e < rows; e += blockDim.x) {
      float s = 0.f;
      int k = 0;
      for (; k + 8 <= nslots; k += 8) {
        float a[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] = p.bias_partials[(size_t)slots[k + q] * 512 + e] + p.bias_partials[(size_t)slots[k + q] * 512 + 256 + e];
#pragma unroll
        for (int q = 0; q < 8; ++q) s += a[q];
      }
      for (; k < nslots; ++k) s += p.bias_partials[(size_t)slots[k] * 512 + e] + p.bias_partials[(size_t)slots[k] * 512 + 256 + e];
      float* dst = p.grads.p[itm.bias] + e;
      *dst = p.accumulate ? *dst + s : s;
    }
  }
}

// workspace carve-up (all offsets 1024-aligned)
struct Workspace {
  size_t dz, partials, bias, segs, heads, flags, stamps, total;
};
Workspace carve(int64_t n_points) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) & ~(size_t)1023; return o; };
  w.dz = take((size_t)num_tiles(n_points) * kDzTileBytes);
  w.partials = take((size_t)kFusedSlots * kPartialSlotBytes);
  w.bias = take((size_t)kFusedSlots * 512 * sizeof(float));
  w.heads = take((size_t)kHeadMaxBlocks * kHeadFloats * sizeof(float));
  w.segs = take((size_t)kFusedSlots * sizeof(Segment));      // segs and flags are adjacent: one memset clears both
  w.flags = take((size_t)num_tiles(n_points) * kFlagsPerTile * sizeof(uint32_t) + 64 + (16 + 4 * kFusedSlots) * sizeof(uint32_t));   // + credit counters + tile hand-out
  w.stamps = take((size_t)num_tiles(n_points) * kFlagsPerTile * sizeof(uint32_t));
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int mvip_debug_wgrad_profile(unsigned long long* out8) {
  MVIP_CUDA_OK(cudaDeviceSynchronize());
  MVIP_CUDA_OK(cudaMemcpyFromSymbol(out8, g_wprof, sizeof(unsigned long long) * 8));
  return MVIP_OK;
}

int mvip_debug_bwd_lag(unsigned long long* out320) {
  MVIP_CUDA_OK(cudaDeviceSynchronize());
  MVIP_CUDA_OK(cudaMemcpyFromSymbol(out320, g_lag, sizeof(unsigned long long) * 320));
  return MVIP_OK;
}

int mvip_debug_bwd_trace(long long* out240) {
#ifdef MVIP_TRACE_BWD
  MVIP_CUDA_OK(cudaDeviceSynchronize());
  MVIP_CUDA_OK(cudaMemcpyFromSymbol(out240, g_btrace, sizeof(long long) * 240));
  return MVIP_OK;
#else
  (void)out240;
  mvip_set_error("mvip_debug_bwd_trace: library built without -DMVIP_TRACE_BWD");
  return MVIP_E_UNSUPPORTED;
#endif
}

// cycles between the chain starts of consecutive CTA pairs (tuning aid; < 0 restores the default)
static int g_stagger = -1;
static int g_throttle_units = 150, g_throttle_cycles = 60000, g_throttle_gain = 40;   // swept at P = 524,288 (units, gain): (150, 40) median 1.249 ms, (500, 20) 1.266, (200, 20) 1.257, no throttle 1.263
int mvip_debug_set_bwd_throttle(int units, int cycles, int gain) { g_throttle_units = units; g_throttle_cycles = cycles; g_throttle_gain = gain; return MVIP_OK; }
int mvip_debug_set_bwd_stagger(int cycles) { g_stagger = cycles; return MVIP_OK; }

// byte offsets of the publication stamps ([n_tiles][10] u32) and the pick-up stamps inside the workspace (debug aid)
int mvip_debug_bwd_stamp_offsets(int64_t n_points, size_t* pub, size_t* pick) {
  const Workspace w = carve(n_points);
  *pub = w.flags; *pick = w.stamps;
  return MVIP_OK;
}

size_t mvip_mlp_backward_workspace_bytes(int64_t n_points) { return carve(n_points).total; }

int mvip_mlp_backward(const void* packed, const float* d_raw, int64_t n_points, const void* stash, void* workspace,
                      float* const* grads, int accumulate, void* stream) {
  return mvip_mlp_backward_phases(packed, d_raw, n_points, stash, workspace, grads, accumulate, 15, stream);
}

int mvip_mlp_backward_phases(const void* packed, const float* d_raw, int64_t n_points, const void* stash,
                             void* workspace, float* const* grads, int accumulate, int phase_mask, void* stream) {
  MVIP_REQUIRE(n_points >= 0, MVIP_E_INVALID, "mvip_mlp_backward: negative n_points");
  MVIP_REQUIRE(grads, MVIP_E_INVALID, "mvip_mlp_backward: null grads");
  GradPtrs gp;
  for (int i = 0; i < MVIP_MLP_NUM_PARAMS; ++i) {
    MVIP_REQUIRE(grads[i], MVIP_E_INVALID, "mvip_mlp_backward: grads[%d] is null", i);
    gp.p[i] = grads[i];
  }
  if (n_points == 0) return MVIP_OK;  // nothing to add (caller zero-initialises when not accumulating)
  MVIP_REQUIRE(packed && d_raw && stash && workspace, MVIP_E_INVALID, "mvip_mlp_backward: null pointer");
  MVIP_REQUIRE(mvip_aligned(packed, 1024) && mvip_aligned(stash, 1024) && mvip_aligned(workspace, 1024) &&
                   mvip_aligned(d_raw, 16),
               MVIP_E_INVALID, "mvip_mlp_backward: packed/stash/workspace need 1024-byte and d_raw 16-byte alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const Workspace ws = carve(n_points);
  uint8_t* wsb = static_cast<uint8_t*>(workspace);
  const int64_t n_tiles = num_tiles(n_points);
  const int sms = mvip_num_sms();

  ChainParams cp;
  cp.packed = static_cast<const uint8_t*>(packed);
  cp.d_raw = reinterpret_cast<const float4*>(d_raw);
  cp.stash = static_cast<const uint8_t*>(stash);
  cp.dz = wsb + ws.dz;
  cp.n_points = n_points;
  cp.n_tiles = n_tiles;
  cp.flags = reinterpret_cast<uint32_t*>(wsb + ws.flags);
  cp.cons_stamp = reinterpret_cast<uint32_t*>(wsb + ws.stamps);
  cp.credit = cp.flags + (size_t)n_tiles * kFlagsPerTile;
  cp.next_pair = cp.credit + 16;
  cp.throttle_units = g_throttle_units;
  cp.throttle_cycles = g_throttle_cycles;
  cp.throttle_gain = g_throttle_gain;
  WParams wp;
  wp.stash = static_cast<const uint8_t*>(stash);
  wp.dz = wsb + ws.dz;
  wp.n_tiles = n_tiles;
  wp.partials = reinterpret_cast<float*>(wsb + ws.partials);
  wp.bias_partials = reinterpret_cast<float*>(wsb + ws.bias);
  wp.segs = reinterpret_cast<Segment*>(wsb + ws.segs);
  const int clusters = sms / 2 < kFusedSlots ? sms / 2 : kFusedSlots;
  // one tile pair takes a chain ~53,000 cycles (profiles/r02_bwd_trace.md): spread the chain starts over one such period;
  // pointless when there is a single wave of tile pairs
  cp.stagger = ((n_tiles + 1) / 2 > clusters) ? (g_stagger >= 0 ? g_stagger : 53000 / clusters) : 0;
  MVIP_REQUIRE((n_tiles + 1) / 2 + clusters < (1 << 20), MVIP_E_UNSUPPORTED, "mvip_mlp_backward: more than 2^20 tile pairs (268 M points) per call");
  MVIP_REQUIRE(clusters >= kFusedItems, MVIP_E_UNSUPPORTED, "mvip_mlp_backward: needs at least %d SM pairs", kFusedItems);

  // 1. fused dgrad chain + weight gradients (phase bit 1; bit 2 is kept for ABI compatibility and launches nothing)
  if (phase_mask & 1) {
    // CTA pairs dealt out to the work items in proportion to their cost (largest remainder); always ALL pairs, whatever the
    // number of tiles: every item needs its consumers
    static FusedPlan plan;
    static int plan_clusters = 0;
    if (plan_clusters != clusters) {
      FItem items_h[kFusedItems];
      MVIP_CUDA_OK(cudaMemcpyFromSymbol(items_h, kFItems, sizeof(items_h)));
      int total = 0, n_i[kFusedItems], assigned = 0;
      for (int i = 0; i < kFusedItems; ++i) total += items_h[i].cost;
      for (int i = 0; i < kFusedItems; ++i) { n_i[i] = clusters * items_h[i].cost / total; if (n_i[i] < 1) n_i[i] = 1; assigned += n_i[i]; }
      while (assigned < clusters) {   // next pair to the item with the largest cost per pair
        int best = 0;
        for (int i = 1; i < kFusedItems; ++i)
          if ((long long)items_h[i].cost * n_i[best] > (long long)items_h[best].cost * n_i[i]) best = i;
        ++n_i[best]; ++assigned;
      }
      while (assigned > clusters) {   // (only if the floor of 1 pair per item overshot)
        int worst = -1;
        for (int i = 0; i < kFusedItems; ++i)
          if (n_i[i] > 1 && (worst < 0 || (long long)items_h[i].cost * n_i[worst] < (long long)items_h[worst].cost * n_i[i])) worst = i;
        --n_i[worst]; --assigned;
      }
      // interleave the items over the cluster index, so that neighbouring SM pairs work on different layers
      int given[kFusedItems] = {0};
      int c = 0;
      while (c < clusters)
        for (int i = 0; i < kFusedItems && c < clusters; ++i)
          if (given[i] < n_i[i]) { plan.item[c] = (unsigned char)i; plan.k[c] = (unsigned char)given[i]; plan.n[c] = (unsigned char)n_i[i]; ++given[i]; ++c; }
      plan_clusters = clusters;
    }
    // flags, credit counters, tile hand-out and the segment table (all-zero segment = unused slot: t1 == t0) in one memset
    MVIP_CUDA_OK(cudaMemsetAsync(wsb + ws.segs, 0, ws.stamps - ws.segs, st));
    const size_t smem = kFSmemBytes + 1024;
    MVIP_SMEM_OPT_IN(backward_fused_kernel, smem);
    backward_fused_kernel<<<2 * clusters, kFThreads, smem, st>>>(cp, wp, plan);
    MVIP_LAUNCH_OK("backward_fused_kernel");
  }
  // 3. heads
  const int64_t n_half = (n_points + kHgRows - 1) / kHgRows;
  const int head_cap = sms < kHeadMaxBlocks ? sms : kHeadMaxBlocks;
  const int head_grid = (int)(n_half < head_cap ? n_half : head_cap);
  if (phase_mask & 4) {
    MVIP_SMEM_OPT_IN(head_grads_kernel, kHgSmemBytes);
    head_grads_kernel<<<head_grid, kHgThreads, kHgSmemBytes, st>>>(reinterpret_cast<const float4*>(d_raw), static_cast<const uint8_t*>(stash),
                                                                   n_points, n_half, reinterpret_cast<float*>(wsb + ws.heads));
    MVIP_LAUNCH_OK("head_grads_kernel");
  }
  // 4. reduce
  if (phase_mask & 8) {
    ReduceParams rp;
    rp.partials = reinterpret_cast<const float*>(wsb + ws.partials);
    rp.bias_partials = reinterpret_cast<const float*>(wsb + ws.bias);
    rp.segs = reinterpret_cast<const Segment*>(wsb + ws.segs);
    rp.n_slots = kFusedSlots;
    rp.head_partials = reinterpret_cast<const float*>(wsb + ws.heads);
    rp.head_grid = head_grid;
    rp.grads = gp;
    rp.accumulate = accumulate;
    reduce_kernel<<<dim3(kReduceGridX, kRealItems + 1), 256, 0, st>>>(rp);
    MVIP_LAUNCH_OK("reduce_kernel");
  }
  return MVIP_OK;
}

}  // extern "C"
