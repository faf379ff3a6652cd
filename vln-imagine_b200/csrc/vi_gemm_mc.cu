// vi_gemm_mc.cu - EXPERIMENTAL, not on the default path: the tcgen05 GEMM of vi_gemm_tc.cu with the weight tile shared across
// a thread-block cluster by TMA multicast.  Kept as the measurement instrument it was built to be.
//
// Hypothesis it tested: the wide GEMMs of the navigation step (FFN1, QKV, FFN2, context K | V) are bound by operand delivery into
// the SMs (profiles/r01b_gemm_source_level.md: a 128 x 256 tile pulls 48 KB per k-block per SM; tensor pipe 41 % active).  Here a
// cluster of TWO CTAs works on two vertically adjacent 128-row tiles of the same BN columns: both need the same W tile, so each
// CTA TMA-loads HALF of it and multicasts that half into the shared memory of both (cp.async.bulk.tensor ... .multicast::cluster).
// Every SM then pulls 16 KB (its X tile) + BN / 2 x 128 B (its half of W) per k-block: 32 KB instead of 48 KB at BN = 256.
//   * each CTA issues its own cta_group::1 MMAs (M = 128, N = BN) on the full W tile the two halves form in ITS shared memory;
//   * a ring slot is recycled when BOTH CTAs' MMAs on it have retired: every MMA warp commits with a multicast arrive on the
//     "empty" barrier of both CTAs (count 2) - the scheme vi_gemm_rb.cu already uses for its activation tile;
//   * accumulators, epilogue (bias, GELU / ReLU, TMA-fetched fp32 residual, swizzled staging, TMA store) and the tile walk are
//     those of the CTA-pair mode of vi_gemm_tc.cu: a cluster walks 256-row x BN tiles, rank r owns rows [128 r, 128 r + 128).
// Result on B200 (tools/gemm_mc_check.py): bit-identical to vi_gemm_bf16_tiled on every shape and 3-13 % SLOWER - halving the W
// traffic per SM buys nothing, so operand delivery is not the bound.  The same kernel, FFN1 shape, isolates the activation:
// no activation 20.5 us, erf-form GELU 27.0 us, one-MUFU tanh-form GELU 21.9 us (`--gelu`).
#include "vi_common.cuh"

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int MAX_GROUPS = 4;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int MAX_CPW = 4;
constexpr int BIAS_BYTES = EPI_WARPS * 2 * 32 * 4;
constexpr int NCL = 2;                                  // cluster size (CTAs that share a W tile)
constexpr uint16_t CL_MASK = 0x3;
__host__ __device__ constexpr int epi_buf_bytes(bool f32out) { return f32out ? 4096 : 2048; }
__host__ __device__ constexpr int epi_bytes(bool f32out) { return EPI_WARPS * 2 * epi_buf_bytes(f32out); }

struct McParams {
  const float* bias;
  int has_residual;
  int M, N, K;
  int epilogue;
  int n_groups;
  int group_tile_end[MAX_GROUPS];      // in 128-row tiles
  int num_m_tiles, num_n_tiles;        // m tiles in units of 256 rows (one cluster tile)
};

__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int mma_m, int mma_n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(mma_m >> 4) << 24);
}
__device__ __forceinline__ float gelu_fast(float x) {      // as in vi_gemm_tc.cu (A&S 7.1.26 erf, two MUFU ops)
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  poly *= t;
  const float erf_abs = fmaf(-poly, __expf(-z * z), 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
// one-MUFU GELU (tanh form) - measurement only (epilogue code 3 of THIS kernel, tools/gemm_mc_check.py --gelu): its distance from
// the erf form (< 5e-4 absolute) is below the bf16 rounding of the output, but whether it may replace the erf form is a parity
// decision for the default kernel, not taken here
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// this CTA's half of the W tile, delivered to the same shared-memory offset (and mbarrier) of every CTA in the mask
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__host__ __device__ constexpr uint32_t tmem_cols_for(int bn) { return 2 * bn <= 128 ? 128u : (2 * bn <= 256 ? 256u : 512u); }

template <int BN, int STAGES, bool F32OUT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_mc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, const McParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;               // the FULL W tile lives in every CTA's shared memory
  constexpr int B_HALF = B_BYTES / NCL;              // what this CTA loads (and multicasts)
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = tmem_cols_for(BN);
  constexpr int NCHUNK = BN / 32;
  constexpr int EPI_BUF = epi_buf_bytes(F32OUT);
  static_assert(BN % 64 == 0 && BN >= 128 && BN <= 256, "tile N");
  static_assert(A_BYTES % 1024 == 0 && B_HALF % 1024 == 0, "operand tiles must keep the 1024-byte swizzle alignment");

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

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int worker = (int)(blockIdx.x / NCL);          // cluster id: walks 256-row x BN cluster tiles
  const int n_workers = (int)(gridDim.x / NCL);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
    if (p.has_residual) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NCL);                    // one multicast commit from each CTA's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS);
    }
    for (int w = 0; w < EPI_WARPS; ++w) {
      mbar_init(res_bar(w, 0), 1);
      mbar_init(res_bar(w, 1), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();                                  // both CTAs' barriers exist before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = p.K / BK;
  auto group_of = [&](int mt128) {
    int g = 0;
    while (g < p.n_groups - 1 && mt128 >= p.group_tile_end[g]) ++g;
    return g;
  };

  if (warp == 0) {
    // ------------------------------- TMA producer: own X tile + own half of the shared W tile (multicast)
    int s = 0;
    uint32_t ph = 0;
    for (int t = worker; t < total_tiles; t += n_workers) {
      const int mt = t / p.num_n_tiles, nt = t - mt * p.num_n_tiles;
      const int arow = (mt * NCL + (int)rank) * BM;
      const int wrow = group_of(mt * NCL) * p.N + nt * BN + (int)rank * (BN / NCL);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);              // BOTH CTAs have consumed this slot
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), STAGE_BYTES);    // own X tile + both halves of W land on this CTA's barrier
          tma_load_2d(a_addr(s), &tmA, full_bar(s), kb * BK, arow);
          tma_load_2d_mc(b_addr(s) + rank * (uint32_t)B_HALF, &tmB, full_bar(s), kb * BK, wrow, CL_MASK);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (every CTA issues its own M = 128 MMAs)
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int t = worker; t < total_tiles; t += n_workers, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr(s));
          const uint64_t bdesc = make_sw128_kmajor_desc(b_addr(s));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          tc_commit_mc(empty_bar(s), CL_MASK);         // frees the slot in both CTAs once these MMAs retire
          if (kb == num_kb - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue (vi_gemm_tc.cu, CTA-pair addressing: rank r owns rows [128 r, +128) of the tile)
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    constexpr int CPW0 = (NCHUNK + 1) / 2, CPW1 = NCHUNK / 2;
    const int cpw = half ? CPW1 : CPW0;
    const uint32_t my_buf = epi_base + (uint32_t)(ew * 2 * EPI_BUF);
    float* my_bias = bias_smem + ew * 2 * 32;
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
      if (res && lane == 0) issue_residual(0, mt * NCL * BM + row_in_tile, nt * BN + half * 32);
      if (p.bias) my_bias[lane] = (*(p.bias + (long long)group_of(mt * NCL) * p.N + nt * BN + half * 32 + lane));
      __syncwarp();
    }
    for (int t = worker; t < total_tiles; t += n_workers, ++it) {
      const int mt = t / p.num_n_tiles, nt = t - mt * p.num_n_tiles;
      const int g = group_of(mt * NCL);
      const int row0 = mt * NCL * BM + row_in_tile;
      const int tcol0 = nt * BN + half * 32;
      const int tn = t + n_workers;
      const int mtn = tn / p.num_n_tiles, ntn = tn - mtn * p.num_n_tiles;
      const int row0n = mtn * NCL * BM + row_in_tile, tcol0n = ntn * BN + half * 32;
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const float* bias_g = p.bias ? p.bias + (long long)g * p.N : nullptr;
      const float* bias_gn = (p.bias && tn < total_tiles) ? p.bias + (long long)group_of(mtn * NCL) * p.N : nullptr;
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
                bulk_wait_read<0>();
                issue_residual(n + 1, last ? row0n : row0, last ? tcol0n : col0 + 64);
              }
            } else {
              bulk_wait_read<1>();
            }
          }
          float bias_next = 0.f;
          if (k + 1 < cpw) { if (bias_g) bias_next = (*(bias_g + col0 + 64 + lane)); }
          else if (bias_gn) bias_next = (*(bias_gn + tcol0n + lane));
          tmem_ld_wait();
          if (k + 1 < cpw) {
            tmem_ld_32x32(tbase + (uint32_t)((k + 1) * 64), r[(k + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
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
          } else if (p.epilogue == 3) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
          }
          __syncwarp();
          if (res) {
            mbar_wait(res_bar(ew, n & 1), (uint32_t)(n >> 1) & 1u);
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
  cluster_sync_all();                                 // no CTA leaves while its peer may still multicast / signal into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

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

int make_map(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int elem_bytes, uint64_t inner, uint64_t rows,
             uint64_t ld_elems, uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle sw) {
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {ld_elems * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vi_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VI_ERR_CUDA;
  }
  return VI_OK;
}

template <int BN, int STAGES, bool F32OUT>
constexpr int smem_bytes() {
  return STAGES * (BM * BK * 2 + BN * BK * 2) + epi_bytes(F32OUT) + BIAS_BYTES + (2 * STAGES + 4 + 2 * EPI_WARPS) * 8 + 16;
}

template <int BN, int STAGES, bool F32OUT>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR, const McParams& p,
           long long cluster_tiles, cudaStream_t st) {
  static_assert(smem_bytes<BN, STAGES, F32OUT>() <= 232448, "shared memory budget");
  static bool attr_set = false;
  static int max_clusters = 0;
  auto kern = gemm_bf16_mc_kernel<BN, STAGES, F32OUT>;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  if (!attr_set) {
    VI_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN, STAGES, F32OUT>()));
    cudaLaunchConfig_t q;
    memset(&q, 0, sizeof(q));
    q.gridDim = dim3((unsigned)(vi_num_sms() / NCL * NCL));
    q.blockDim = dim3(NUM_THREADS);
    q.dynamicSmemBytes = smem_bytes<BN, STAGES, F32OUT>();
    q.attrs = attr;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = vi_num_sms() / NCL;
    }
    max_clusters = n;
    attr_set = true;
  }
  const long long clusters = cluster_tiles < max_clusters ? cluster_tiles : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * NCL));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes<BN, STAGES, F32OUT>();
  cfg.stream = st;
  int nattr = 1;
  if (vi_pdl_enabled()) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  VI_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmY, tmR, p));
  return VI_OK;
}

}  // namespace

extern "C" int vi_gemm_bf16_mc(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual, int64_t ldr,
                               void* y, int64_t ldy, int y_dtype, int M, int N, int K, int epilogue, int n_groups,
                               const int32_t* group_row_end, int tile, vi_stream_t stream) {
  VI_CHECK_ARG(x && w && y, "vi_gemm_bf16_mc: null operand");
  VI_CHECK_ARG(M > 0 && N > 0 && K > 0 && K % BK == 0, "vi_gemm_bf16_mc: bad sizes M=%d N=%d K=%d (K %% 64 == 0)", M, N, K);
  VI_CHECK_ARG(tile == 128 || tile == 192 || tile == 256, "vi_gemm_bf16_mc: tile width must be 128, 192 or 256 (got %d)", tile);
  VI_CHECK_ARG(N % tile == 0, "vi_gemm_bf16_mc: tile width %d does not divide N=%d", tile, N);
  VI_CHECK_ARG(ldx % 8 == 0 && ldx >= K && ldy >= N && ldy % 8 == 0, "vi_gemm_bf16_mc: bad leading dimensions");
  VI_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0),
               "vi_gemm_bf16_mc: operands must be 16-byte aligned");
  VI_CHECK_ARG(!residual || (y_dtype == VI_DT_F32 && ldr >= N && ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0),
               "vi_gemm_bf16_mc: a residual needs an fp32 output, ldr >= N, 16-byte alignment");
  VI_CHECK_ARG(y_dtype == VI_DT_BF16 || y_dtype == VI_DT_F32, "vi_gemm_bf16_mc: bad y_dtype %d", y_dtype);
  VI_CHECK_ARG(epilogue >= VI_EPI_NONE && epilogue <= 3, "vi_gemm_bf16_mc: bad epilogue %d (3 = tanh-form GELU, measurement only)", epilogue);
  VI_CHECK_ARG(n_groups >= 1 && n_groups <= MAX_GROUPS && (n_groups == 1 || group_row_end), "vi_gemm_bf16_mc: bad row groups");
  if (int rc = resolve_encode()) return rc;

  McParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.has_residual = residual != nullptr;
  p.M = M; p.N = N; p.K = K; p.epilogue = epilogue; p.n_groups = n_groups;
  const int m_tiles128 = (M + BM - 1) / BM;
  for (int g = 0; g < n_groups; ++g) {
    if (n_groups == 1) { p.group_tile_end[g] = m_tiles128; break; }
    const int e = group_row_end[g];
    VI_CHECK_ARG(e > 0 && e <= M && (g == 0 || e > group_row_end[g - 1]), "vi_gemm_bf16_mc: bad group_row_end[%d]=%d", g, e);
    VI_CHECK_ARG(g == n_groups - 1 || e % (NCL * BM) == 0, "vi_gemm_bf16_mc: group %d must end on a multiple of %d rows", g, NCL * BM);
    p.group_tile_end[g] = (e + BM - 1) / BM;
  }
  VI_CHECK_ARG(n_groups == 1 || group_row_end[n_groups - 1] == M, "vi_gemm_bf16_mc: last group must end at M");
  p.num_m_tiles = (M + NCL * BM - 1) / (NCL * BM);
  p.num_n_tiles = N / tile;
  const long long cluster_tiles = (long long)p.num_m_tiles * p.num_n_tiles;
  const bool f32 = y_dtype == VI_DT_F32;

  CUtensorMap tmA, tmB, tmY, tmR;
  memset(&tmR, 0, sizeof(tmR));
  if (int rc = make_map(&tmA, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, BK, BM,
                        CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = make_map(&tmB, w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (uint64_t)K, (uint64_t)n_groups * N, (uint64_t)K, BK,
                        (uint32_t)(tile / NCL), CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (f32) {
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
  // ring depth = what fits beside the epilogue staging (bf16 out: 32 KB, fp32 out: 64 KB) in 227 KB
  switch (tile) {
    case 256: return f32 ? launch<256, 3, true>(tmA, tmB, tmY, tmR, p, cluster_tiles, st) : launch<256, 4, false>(tmA, tmB, tmY, tmR, p, cluster_tiles, st);
    case 192: return f32 ? launch<192, 4, true>(tmA, tmB, tmY, tmR, p, cluster_tiles, st) : launch<192, 4, false>(tmA, tmB, tmY, tmR, p, cluster_tiles, st);
    default:  return f32 ? launch<128, 5, true>(tmA, tmB, tmY, tmR, p, cluster_tiles, st) : launch<128, 6, false>(tmA, tmB, tmY, tmR, p, cluster_tiles, st);
  }
}
