// vi_gemm_tc.cu - bf16 GEMM on the 5th-gen tensor cores: Y = epi(X W^T + bias) + residual.
//
// Persistent, warp-specialised, one CTA (or one CTA pair) per SM walking output tiles:
//   warp 0      : TMA producer  - cp.async.bulk.tensor loads of the X (128 x 64) and W (BN x 64) k-blocks into
//                 a STAGES-deep shared-memory ring (128B swizzle), completion counted on mbarriers;
//   warp 1      : MMA issuer    - one thread issues tcgen05.mma (K=16 per instruction) accumulating fp32 in
//                 TMEM; tcgen05.commit frees ring slots and publishes finished accumulators;
//   warp 2      : TMEM allocator (2 x BN columns: the accumulator is double buffered, so the epilogue of
//                 tile i overlaps the main loop of tile i+1);
//   warps 4..11 : epilogue      - tcgen05.ld (32 lanes x 32 columns per warp), + bias, GELU / ReLU, + residual,
//                 staged through swizzled shared memory and written with TMA stores; the fp32 residual tile
//                 is fetched by TMA into the same staging buffer, so every global access is a bulk copy.
// PAIR mode (cta_group::2): two CTAs of a cluster work on one 256 x BN tile; each loads its own 128 rows of X
// and half of the W tile, the leader issues M=256 MMAs that read both CTAs' shared memory, and each CTA drains
// its own half of the accumulator.  This halves the W traffic per flop, which is what bounds a 128 x BN tile
// (L2 -> SM bandwidth, B300_MICROARCH.md "TMA chip-throughput").
// The grouped form lets consecutive row ranges of X use different weight blocks of a stacked W (DUET's
// global / local encoders, HAMT's language / vision streams) in a single launch.
//
// Replaces nn.Linear + its following activation / residual in the reference
// (VLN-DUET/map_nav_src/models/vilmodel.py:93-95,147,172,186,315-317; transformer.py:178,181).
#include "vi_common.cuh"

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;           // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int MAX_GROUPS = 4;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int MAX_CPW = 4;                             // 32-column chunks per epilogue warp per tile (BN <= 256)
constexpr int BIAS_BYTES = EPI_WARPS * 2 * 32 * 4;     // per epilogue warp: the bias slice of the current and the next chunk
// staging per epilogue warp: two buffers of 32 rows x 32 columns (fp32: 128-byte rows, bf16: 64-byte rows)
__host__ __device__ constexpr int epi_buf_bytes(bool f32out) { return f32out ? 4096 : 2048; }
__host__ __device__ constexpr int epi_bytes(bool f32out) { return EPI_WARPS * 2 * epi_buf_bytes(f32out); }
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;            // clears the CTA-rank bit of a shared::cluster address (pair leader)

struct GemmParams {
  const float* bias;
  int has_residual;
  int y_f32;
  int M, N, K;
  int epilogue;
  int n_groups;
  int group_tile_end[MAX_GROUPS];      // in units of 128-row tiles
  int num_m_tiles, num_n_tiles;        // m tiles in units of the launch's tile height (128, or 256 in PAIR mode)
};

__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
  // UMMA shared-memory descriptor, K-major operand, 128-byte swizzle:
  //   bits [0,14) start address >> 4; [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1);
  //   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups; [46,48) descriptor version 1 (sm_100);
  //   [61,64) layout type 2 = SWIZZLE_128B.
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc(int mma_m, int mma_n) {
  // kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both operands
  // K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(mma_m >> 4) << 24);
}

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): two MUFU ops instead of erff's branchy polynomial.  The
// GELU output of this kernel is rounded to bf16 (relative step 4e-3), so the approximation is invisible.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));      // MUFU.RCP, 1 ulp
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, __expf(-z * z), 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// ---- PTX helpers that exist only for this kernel ---------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {          // one lane of a converged warp
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// pair mode: this CTA's TMA load completes on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tmap), "r"(bar & PEER_MASK), "r"(c0), "r"(c1), "l"(0x1000000000000000ull)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
template <bool PAIR> __device__ __forceinline__ void tc_commit_t(uint32_t bar) {
  if constexpr (PAIR) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  } else {
    tc_commit(bar);
  }
}
template <bool PAIR>
__device__ __forceinline__ void tc_mma_t(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  } else {
    tc_mma_bf16(d_tmem, adesc, bdesc, idesc, acc);
  }
}
template <bool PAIR> __device__ __forceinline__ void tmem_alloc_t(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (PAIR) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    tmem_alloc(dst_smem, ncols);
    tmem_relinquish();
  }
}
template <bool PAIR> __device__ __forceinline__ void tmem_dealloc_t(uint32_t taddr, uint32_t ncols) {
  if constexpr (PAIR) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  } else {
    tmem_dealloc(taddr, ncols);
  }
}

__host__ __device__ constexpr uint32_t tmem_cols_for(int bn) { return 2 * bn <= 128 ? 128u : (2 * bn <= 256 ? 256u : 512u); }

template <int BN, int STAGES, bool PAIR, bool F32OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                    const GemmParams p) {
  constexpr int BN_CTA = PAIR ? BN / 2 : BN;         // rows of the W tile this CTA stages
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN_CTA * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(BN);
  constexpr int NCHUNK = BN / 32;                    // 32-column epilogue chunks per tile
  constexpr int EPI_BUF = epi_buf_bytes(F32OUT);
  constexpr int TILES_PER_M = PAIR ? 2 : 1;          // 128-row tiles per m index
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "tile N");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "operand tiles must keep the 1024-byte swizzle alignment");

  // The kernel has no static shared memory, so the dynamic window starts at offset 0 of the CTA's shared memory
  // and is 1024-byte aligned, which the 128B-swizzled tiles need (checked below: a misaligned base traps).
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  float* bias_smem = reinterpret_cast<float*>(epi_smem + epi_bytes(F32OUT));
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + epi_bytes(F32OUT) + BIAS_BYTES);
  // barrier slots: full[S], empty[S], tfull[2], tempty[2], rbar[EPI_WARPS][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * EPI_WARPS);

  const uint32_t smem_base = smem_u32(smem);
  if (smem_base & 1023u) __trap();
  const uint32_t epi_base = smem_u32(epi_smem);
  const uint32_t bar_base = smem_u32(bars);
  auto a_addr = [&](int s) { return smem_base + (uint32_t)(s * STAGE_BYTES); };
  auto b_addr = [&](int s) { return smem_base + (uint32_t)(s * STAGE_BYTES + A_BYTES); };
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * STAGES + 2 + a); };
  auto res_bar = [&](int w, int b) { return bar_base + 8u * (uint32_t)(2 * STAGES + 4 + 2 * w + b); };

  pdl_launch_dependents();                            // the next kernel may begin its own prologue
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // tile walker id (CTA or pair)
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (p.has_residual) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 2 * EPI_WARPS : EPI_WARPS);      // one arrive per epilogue warp (of both CTAs)
    }
    for (int w = 0; w < EPI_WARPS; ++w) {
      mbar_init(res_bar(w, 0), 1);
      mbar_init(res_bar(w, 1), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_t<PAIR>(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                         // everything above overlapped the previous kernel's tail

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = p.K / BK;

  auto group_of = [&](int mt128) {
    int g = 0;
    while (g < p.n_groups - 1 && mt128 >= p.group_tile_end[g]) ++g;
    return g;
  };

  if (warp == 0) {
    // ------------------------------- TMA producer (whole warp converged, one elected lane issues) -----------
    int s = 0;
    uint32_t ph = 0;
    for (int t = worker; t < total_tiles; t += n_workers) {
      const int mt = t / p.num_n_tiles, nt = t - mt * p.num_n_tiles;
      const int arow = (mt * TILES_PER_M + (int)rank) * BM;
      const int wrow = group_of(mt * TILES_PER_M) * p.N + nt * BN + (int)rank * BN_CTA;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        if (elect_one()) {
          if constexpr (PAIR) {
            if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);       // both CTAs' bytes land on the leader
            tma_load_2d_pair(a_addr(s), &tmA, full_bar(s), kb * BK, arow);
            tma_load_2d_pair(b_addr(s), &tmB, full_bar(s), kb * BK, wrow);
          } else {
            mbar_expect_tx(full_bar(s), STAGE_BYTES);
            tma_load_2d(a_addr(s), &tmA, full_bar(s), kb * BK, arow);
            tma_load_2d(b_addr(s), &tmB, full_bar(s), kb * BK, wrow);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (pair: leader only; warp converged, elected lane issues) ----
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(PAIR ? 256 : 128, BN);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int t = worker; t < total_tiles; t += n_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);                 // TMA bytes have landed (in both CTAs)
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = make_sw128_kmajor_desc(a_addr(s));
            const uint64_t bdesc = make_sw128_kmajor_desc(b_addr(s));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in 16-byte units
              tc_mma_t<PAIR>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            }
            tc_commit_t<PAIR>(empty_bar(s));          // frees the ring slot (of both CTAs) when the MMAs retire
            if (kb == num_kb - 1) tc_commit_t<PAIR>(tfull_bar(acc));   // accumulator complete -> epilogue (both CTAs)
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------------------
    const int ew = warp - 4;                          // 0..7
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
    const int half = ew >> 2;                         // which interleaved half of the column chunks
    constexpr int CPW0 = (NCHUNK + 1) / 2, CPW1 = NCHUNK / 2;
    const int cpw = half ? CPW1 : CPW0;               // chunks per tile for this warp: half, half+2, ...
    const uint32_t my_buf = epi_base + (uint32_t)(ew * 2 * EPI_BUF);
    float* my_bias = bias_smem + ew * 2 * 32;         // [2][32]: bias slice of chunk n in buffer n & 1
    const uint32_t my_bias_u = smem_u32(my_bias);
    const bool res = p.has_residual != 0;
    const int row_in_tile = (int)rank * BM + quarter * 32;

    auto issue_residual = [&](int n, int row0, int col0) {          // lane 0 only
      const uint32_t bar = res_bar(ew, n & 1);
      mbar_expect_tx(bar, 4096);
      tma_load_2d(my_buf + (uint32_t)((n & 1) * EPI_BUF), &tmR, bar, col0, row0);
    };

    int n = 0, it = 0;
    if (worker < total_tiles) {
      const int mt = worker / p.num_n_tiles, nt = worker - mt * p.num_n_tiles;
      if (res && lane == 0) issue_residual(0, mt * TILES_PER_M * BM + row_in_tile, nt * BN + half * 32);
      if (p.bias) my_bias[lane] = (*(p.bias + (long long)group_of(mt * TILES_PER_M) * p.N + nt * BN + half * 32 + lane));
      __syncwarp();
    }
    for (int t = worker; t < total_tiles; t += n_workers, ++it) {
      const int mt = t / p.num_n_tiles, nt = t - mt * p.num_n_tiles;
      const int g = group_of(mt * TILES_PER_M);
      const int row0 = mt * TILES_PER_M * BM + row_in_tile;
      const int tcol0 = nt * BN + half * 32;          // first column chunk of this warp; the next ones are +64 apart
      // first chunk of the NEXT tile (for the residual prefetch that crosses the tile boundary)
      const int tn = t + n_workers;
      const int mtn = tn / p.num_n_tiles, ntn = tn - mtn * p.num_n_tiles;
      const int row0n = mtn * TILES_PER_M * BM + row_in_tile, tcol0n = ntn * BN + half * 32;
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const float* bias_g = p.bias ? p.bias + (long long)g * p.N : nullptr;
      const float* bias_gn = (p.bias && tn < total_tiles) ? p.bias + (long long)group_of(mtn * TILES_PER_M) * p.N : nullptr;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + half * 32);
      uint32_t r[2][32];
      tmem_ld_32x32(tbase, r[0]);
#pragma unroll
      for (int k = 0; k < MAX_CPW; ++k) {
        if (k < cpw) {
          const uint32_t buf = my_buf + (uint32_t)((n & 1) * EPI_BUF);
          const int col0 = tcol0 + k * 64;
          if (lane == 0) {
            if (res) {
              const bool last = k == cpw - 1;
              if (!last || tn < total_tiles) {
                bulk_wait_read<0>();                  // the store that last used the other buffer has read it
                issue_residual(n + 1, last ? row0n : row0, last ? tcol0n : col0 + 64);
              }
            } else {
              bulk_wait_read<1>();                    // the store issued two chunks ago has read this buffer
            }
          }
          // bias slice of the NEXT chunk (possibly the first one of the next tile): loaded now, parked at the end
          float bias_next = 0.f;
          if (k + 1 < cpw) { if (bias_g) bias_next = (*(bias_g + col0 + 64 + lane)); }
          else if (bias_gn) bias_next = (*(bias_gn + tcol0n + lane));
          tmem_ld_wait();
          if (k + 1 < cpw) {
            tmem_ld_32x32(tbase + (uint32_t)((k + 1) * 64), r[(k + 1) & 1]);   // next chunk in flight during the math
          } else {                                    // accumulator fully read by this warp: hand it back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc));
            }
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[k & 1][j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 b;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(my_bias_u + (uint32_t)((n & 1) * 128 + j * 16)));
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (p.epilogue == VI_EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
          } else if (p.epilogue == VI_EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          __syncwarp();                               // lane 0's wait on the staging buffer covers the whole warp
          if (res) {
            mbar_wait(res_bar(ew, n & 1), (uint32_t)(n >> 1) & 1u);
            // residual tile: 32 rows x 128 B, 128B-swizzled: 16-byte chunk j of row r sits at j ^ (r & 7)
            const uint32_t rowb = buf + (uint32_t)(lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 q;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "r"(rowb + (uint32_t)((j ^ (lane & 7)) << 4)));
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          }
          if constexpr (F32OUT) {
            const uint32_t rowb = buf + (uint32_t)(lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                           ::"r"(rowb + (uint32_t)((j ^ (lane & 7)) << 4)), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]),
                           "f"(v[4 * j + 3]) : "memory");
          } else {
            // bf16 rows are 64 B, 64B-swizzled: chunk j of row r sits at j ^ ((r >> 1) & 3)
            const uint32_t rowb = buf + (uint32_t)(lane * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};"
                           ::"r"(rowb + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4)), "r"(pack_bf16x2(v[8 * j], v[8 * j + 1])),
                           "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])), "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                           "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7])) : "memory");
          }
          if (p.bias) my_bias[((n + 1) & 1) * 32 + lane] = bias_next;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmY, buf, col0, row0);      // rows beyond M are clipped by the tensor map
            bulk_commit();
          }
          ++n;
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_t<PAIR>(tmem_base, TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;

int resolve_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!g_encode) {
    vi_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VI_ERR_CUDA;
  }
  return VI_OK;
}

// 2D row-major tensor map: inner dim contiguous, `box_inner x box_rows` boxes
int make_map(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int elem_bytes, uint64_t inner, uint64_t rows,
             uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle sw) {
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vi_set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p inner=%llu rows=%llu ld=%llu box=%ux%u)",
                 (int)r, ptr, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)ld_elems, box_inner,
                 box_rows);
    return VI_ERR_CUDA;
  }
  return VI_OK;
}

template <int BN, int STAGES, bool PAIR, bool F32OUT>
constexpr int smem_bytes() {
  return STAGES * (BM * BK * 2 + (PAIR ? BN / 2 : BN) * BK * 2) + epi_bytes(F32OUT) + BIAS_BYTES +
         (2 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
}

template <int BN, int STAGES, bool PAIR, bool F32OUT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR,
           const GemmParams& p, int grid, cudaStream_t st) {
  static_assert(smem_bytes<BN, STAGES, PAIR, F32OUT>() <= 232448, "shared memory budget");
  static bool attr_set = false;          // idempotent; races only repeat the same call
  auto kern = gemm_bf16_tc_kernel<BN, STAGES, PAIR, F32OUT>;
  if (!attr_set) {
    VI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN, STAGES, PAIR, F32OUT>()));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes<BN, STAGES, PAIR, F32OUT>();
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  if (vi_pdl_enabled()) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  if (PAIR) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = 2;
    attr[nattr].val.clusterDim.y = 1;
    attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  VI_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmY, tmR, p));
  return VI_OK;
}

struct Choice { int bn; bool pair; };

// Tile choice by a two-term cost model (cycles): the tensor pipe (waves x MMA time of a tile) against the
// L2 -> SM operand traffic of the whole problem (about 6300 B/cycle chip-wide), plus one exposed epilogue.
Choice pick_tile(int M, int N, int K, int nsm, bool pair_ok) {
  Choice best = {64, false};
  if (const char* e = getenv("VI_GEMM_TILE")) {        // e.g. "192" or "256p" (tests / sweeps)
    const int v = atoi(e);
    const bool pr = strchr(e, 'p') != nullptr;
    if ((v == 64 || v == 96 || v == 128 || v == 192 || v == 256) && N % v == 0 && (!pr || (pair_ok && v >= 128)))
      return Choice{v, pr};
  }
  double best_cost = 1e30;
  const int cands[5] = {256, 192, 128, 96, 64};
  for (int pr = 1; pr >= 0; --pr) {
    if (pr && !pair_ok) continue;
    for (int i = 0; i < 5; ++i) {
      const int bn = cands[i];
      if (N % bn) continue;
      if (pr && bn < 128) continue;
      const int tm = pr ? 256 : 128;
      const long long m_tiles = (M + tm - 1) / tm;
      const long long tiles = m_tiles * (N / bn);
      const long long workers = pr ? nsm / 2 : nsm;
      const long long waves = (tiles + workers - 1) / workers;
      const double kb = K / 64.0;
      const double mma = (double)waves * kb * 4.0 * (bn / 2.0);                 // cycles: K=16 MMA = bn/2 (M=128 per SM)
      const double bytes = (double)tiles * kb * (tm * 128.0 + bn * 128.0);
      const double l2 = bytes / 6300.0;
      const double epi = 1500.0 + 10.0 * bn;
      const double cost = (mma > l2 ? mma : l2) + epi + (double)waves * 300.0;
      if (cost < best_cost - 1e-9) { best_cost = cost; best = Choice{bn, pr != 0}; }
    }
  }
  return best;
}

}  // namespace

extern "C" int vi_gemm_bf16(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual,
                            int64_t ldr, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int epilogue,
                            int n_groups, const int32_t* group_row_end, vi_stream_t stream) {
  return vi_gemm_bf16_tiled(x, ldx, w, bias, residual, ldr, y, ldy, y_dtype, M, N, K, epilogue, n_groups, group_row_end, 0,
                            stream);
}

extern "C" int vi_gemm_bf16_tiled(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual,
                                  int64_t ldr, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int epilogue,
                                  int n_groups, const int32_t* group_row_end, int tile, vi_stream_t stream) {
  VI_CHECK_ARG(x && w && y, "vi_gemm_bf16: null operand");
  VI_CHECK_ARG(M > 0 && N > 0 && K > 0, "vi_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  VI_CHECK_ARG(K % BK == 0, "vi_gemm_bf16: K=%d must be a multiple of %d", K, BK);
  VI_CHECK_ARG(N % 64 == 0, "vi_gemm_bf16: N=%d must be a multiple of 64", N);
  VI_CHECK_ARG(ldx % 8 == 0 && ldx >= K, "vi_gemm_bf16: ldx=%lld must be >= K and a multiple of 8", (long long)ldx);
  VI_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0,
               "vi_gemm_bf16: operands must be 16-byte aligned");
  VI_CHECK_ARG(ldy >= N && ldy % 8 == 0, "vi_gemm_bf16: ldy=%lld must be >= N and a multiple of 8", (long long)ldy);
  VI_CHECK_ARG(!residual || (ldr >= N && ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0),
               "vi_gemm_bf16: residual must be 16-byte aligned with ldr >= N, ldr %% 4 == 0");
  VI_CHECK_ARG(!bias || ((uintptr_t)bias & 15) == 0, "vi_gemm_bf16: bias must be 16-byte aligned");
  VI_CHECK_ARG(!residual || y_dtype == VI_DT_F32,
               "vi_gemm_bf16: a residual needs an fp32 output (the residual stream of the model is fp32)");
  VI_CHECK_ARG(y_dtype == VI_DT_BF16 || y_dtype == VI_DT_F32, "vi_gemm_bf16: bad y_dtype %d", y_dtype);
  VI_CHECK_ARG(epilogue >= VI_EPI_NONE && epilogue <= VI_EPI_RELU, "vi_gemm_bf16: bad epilogue %d", epilogue);
  VI_CHECK_ARG(n_groups >= 1 && n_groups <= MAX_GROUPS, "vi_gemm_bf16: n_groups=%d out of range", n_groups);
  VI_CHECK_ARG(n_groups == 1 || group_row_end, "vi_gemm_bf16: grouped call without group_row_end");
  if (int rc = resolve_encode()) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.has_residual = residual != nullptr; p.y_f32 = (y_dtype == VI_DT_F32);
  p.M = M; p.N = N; p.K = K; p.epilogue = epilogue; p.n_groups = n_groups;
  bool pair_ok = true;                   // a 256-row pair tile must not straddle two weight groups
  const int m_tiles128 = (M + BM - 1) / BM;
  for (int g = 0; g < n_groups; ++g) {
    if (n_groups == 1) { p.group_tile_end[g] = m_tiles128; break; }
    const int e = group_row_end[g];
    VI_CHECK_ARG(e > 0 && e <= M && (g == 0 || e > group_row_end[g - 1]), "vi_gemm_bf16: bad group_row_end[%d]=%d", g, e);
    VI_CHECK_ARG(g == n_groups - 1 || e % BM == 0, "vi_gemm_bf16: group %d must end on a multiple of %d rows (got %d)", g, BM, e);
    if (g < n_groups - 1 && e % (2 * BM) != 0) pair_ok = false;
    p.group_tile_end[g] = (e + BM - 1) / BM;
  }
  VI_CHECK_ARG(n_groups == 1 || group_row_end[n_groups - 1] == M, "vi_gemm_bf16: last group must end at M");

  const int nsm = vi_num_sms();
  Choice c = pick_tile(M, N, K, nsm, pair_ok);
  if (tile != 0) {                       // caller-selected tile (the Python host autotunes per shape)
    const int bn = tile & 0xFFF;
    const bool pr = (tile & VI_TILE_PAIR) != 0;
    VI_CHECK_ARG((bn == 64 || bn == 96 || bn == 128 || bn == 192 || bn == 256) && N % bn == 0,
                 "vi_gemm_bf16_tiled: tile width %d does not divide N=%d", bn, N);
    VI_CHECK_ARG(!pr || (pair_ok && bn >= 128), "vi_gemm_bf16_tiled: CTA-pair tiles need width >= 128 and 256-row group ends");
    c = Choice{bn, pr};
  }
  const int tm = c.pair ? 2 * BM : BM;
  p.num_m_tiles = (M + tm - 1) / tm;
  p.num_n_tiles = N / c.bn;
  const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
  int grid;
  if (c.pair) {
    const long long workers = tiles < nsm / 2 ? tiles : nsm / 2;
    grid = (int)(2 * workers);
  } else {
    grid = (int)(tiles < nsm ? tiles : nsm);
  }

  CUtensorMap tmA, tmB, tmY, tmR;
  memset(&tmR, 0, sizeof(tmR));
  if (int rc = make_map(&tmA, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, BK, BM,
                        CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map(&tmB, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)n_groups * N, (uint64_t)K, BK,
                        (uint32_t)(c.pair ? c.bn / 2 : c.bn), CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (p.y_f32) {
    if (int rc = make_map(&tmY, y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)N, (uint64_t)M, (uint64_t)ldy, 32, 32,
                          CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  } else {
    if (int rc = make_map(&tmY, y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ldy, 32, 32,
                          CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  }
  if (residual) {
    if (int rc = make_map(&tmR, residual, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (uint64_t)N, (uint64_t)M, (uint64_t)ldr, 32, 32,
                          CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  } else {
    tmR = tmY;
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define VI_LAUNCH(BN_, SB_, SF_, PAIR_)                                                          \
  (p.y_f32 ? launch<BN_, SF_, PAIR_, true>(tmA, tmB, tmY, tmR, p, grid, st)                      \
           : launch<BN_, SB_, PAIR_, false>(tmA, tmB, tmY, tmR, p, grid, st))
  // ring depth = what fits beside the epilogue staging (bf16 out: 32 KB, fp32 out: 64 KB) in 227 KB
  if (c.pair) {
    switch (c.bn) {
      case 256: return VI_LAUNCH(256, 6, 5, true);
      case 192: return VI_LAUNCH(192, 6, 5, true);
      default:  return VI_LAUNCH(128, 8, 6, true);
    }
  }
  switch (c.bn) {
    case 256: return VI_LAUNCH(256, 4, 3, false);
    case 192: return VI_LAUNCH(192, 4, 4, false);
    case 128: return VI_LAUNCH(128, 6, 5, false);
    case 96:  return VI_LAUNCH(96, 6, 5, false);
    default:  return VI_LAUNCH(64, 8, 6, false);
  }
#undef VI_LAUNCH
}
