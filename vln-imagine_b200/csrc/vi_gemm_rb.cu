// vi_gemm_rb.cu - "row-block" GEMM for the 768-wide layers with the residual add and LayerNorm fused into the epilogue:
//
//     pre = X W^T + bias + residual          (fp32, optional output: the pre-norm residual stream of the panorama encoder)
//     y   = LayerNorm(pre) * gamma + beta    (fp32 and / or bf16 outputs)
//
// A thread-block CLUSTER of 4 CTAs owns a 128-row block of the output over all 768 columns (CTA rank r computes columns
// [192 r, 192 r + 192) with tcgen05.mma M=128 N=192, fp32 accumulators in TMEM):
//   * the X tile is the same for the four CTAs, so each CTA TMA-loads a 32-row slice of it and MULTICASTS it into the
//     shared memory of all four (cp.async.bulk.tensor ... .multicast::cluster): X crosses L2 -> SM once per cluster instead
//     of four times, which is what bounds the N = 768 GEMMs of this path (L2 -> SM operand traffic);
//   * ring slots are recycled when all four CTAs' MMAs have retired (tcgen05.commit ... .multicast::cluster on the
//     "empty" barrier of every CTA);
//   * LayerNorm needs full-row statistics: every epilogue thread owns one row x 96 columns in registers, the two column
//     halves are combined through shared memory, the four CTAs exchange (sum, M2) per row through DISTRIBUTED shared
//     memory (st.shared::cluster + mbarrier.arrive.release.cluster) and combine them with Chan's parallel-variance formula,
//     so the normalisation is as stable as the two-pass row kernel it replaces.
// The accumulator is double buffered (epilogue of row block i overlaps the main loop of row block i + 1).
//
// Replaces  nn.Linear -> dropout(p = 0) -> + input -> LayerNorm  of BertSelfOutput / BertOutput / BertXAttention's output
// (VLN-DUET/map_nav_src/models/vilmodel.py:147-155,183-194,355-364) and out_proj / linear2 + norm of the pre-norm
// TransformerEncoderLayer (models/transformer.py:170-182).
#include "vi_common.cuh"

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BN = 192;            // columns per CTA; 4 CTAs of a cluster cover N = 768
constexpr int BK = 64;
constexpr int NCL = 4;             // cluster size
constexpr int NTOT = BN * NCL;     // 768
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int MAX_GROUPS = 4;
constexpr int A_BYTES = BM * BK * 2;             // 16 KB
constexpr int A_SLICE = A_BYTES / NCL;           // 4 KB: 32 rows
constexpr int B_BYTES = BN * BK * 2;             // 24 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t TMEM_COLS = 512;              // 2 x 192 accumulator columns
constexpr uint16_t CL_MASK = 0xF;

struct RbParams {
  const float* bias;       // [n_groups * 768] or null
  const float* residual;   // [M, ldr] fp32 or null
  long long ldr;
  const float* gamma;      // [n_groups * 768]
  const float* beta;
  float eps;
  float* pre32;            // [M, 768] or null
  float* y32;              // [M, 768] or null
  bf16* y16;               // [M, 768] or null
  int M, K;
  int n_groups;
  int group_tile_end[MAX_GROUPS];   // in 128-row tiles
  int num_m_tiles;
  int debug;      // VI_RB_DEBUG bit mask (diagnostics): 1 = skip residual loads, 2 = skip output stores, 4 = skip the cluster exchange
};

// shared-memory layout after the operand ring
struct RbShared {
  float2 cl_stats[2][NCL][BM];     // per row-block parity: (sum, M2) of every CTA of the cluster        8 KB
  float2 half_stats[2][BM];        // the two column halves of this CTA                                  2 KB
  float bias[BN], gamma[BN], beta[BN];
  uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2], stats[2];
  uint32_t tmem_slot;
};

constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + (int)sizeof(RbShared) + 1024;

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
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
// this CTA's slice of the X tile, delivered to the same shared-memory offset (and mbarrier) of every CTA in the mask
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
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t raddr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// bounded wait with cluster-scope acquire (the data was written by other CTAs of the cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  long long t0 = 0;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done) {
      if (++spins == 1024u) t0 = clock64();
      if (spins > 1024u && (spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
  } while (!done);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __cluster_dims__(NCL, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_rowblock_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const RbParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the 128B-swizzled operand tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  RbShared* sh = reinterpret_cast<RbShared*>(smem + STAGES * STAGE_BYTES);
  const uint32_t smem_base = smem_u32(smem);
  auto a_addr = [&](int s) { return smem_base + (uint32_t)(s * STAGE_BYTES); };
  auto b_addr = [&](int s) { return smem_base + (uint32_t)(s * STAGE_BYTES + A_BYTES); };

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)(blockIdx.x / NCL), n_clusters = (int)(gridDim.x / NCL);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&sh->full[s]), 1);
      mbar_init(smem_u32(&sh->empty[s]), NCL);                // one multicast commit from each CTA's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&sh->tfull[a]), 1);
      mbar_init(smem_u32(&sh->tempty[a]), EPI_WARPS);
      mbar_init(smem_u32(&sh->stats[a]), BM * NCL);           // 128 row-owner threads of each of the 4 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&sh->tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();                                        // every CTA's barriers exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_slot;
  pdl_wait();

  const int num_kb = p.K / BK;
  auto group_of = [&](int mt) {
    int g = 0;
    while (g < p.n_groups - 1 && mt >= p.group_tile_end[g]) ++g;
    return g;
  };

  if (warp == 0) {
    // ------------------------------- TMA producer
    int s = 0;
    uint32_t ph = 0;
    for (int mt = cluster_id; mt < p.num_m_tiles; mt += n_clusters) {
      const int arow = mt * BM + (int)rank * (BM / NCL);
      const int wrow = group_of(mt) * NTOT + (int)rank * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&sh->empty[s]), ph ^ 1u);          // all four CTAs have consumed this slot
        if (elect_one()) {
          const uint32_t fb = smem_u32(&sh->full[s]);
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_2d_mc(a_addr(s) + rank * A_SLICE, &tmA, fb, kb * BK, arow, CL_MASK);
          tma_load_2d(b_addr(s), &tmB, fb, kb * BK, wrow);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc(BM, BN);
    int s = 0, it = 0;
    uint32_t ph = 0;
    for (int mt = cluster_id; mt < p.num_m_tiles; mt += n_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(smem_u32(&sh->tempty[acc]), acc_ph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&sh->full[s]), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr(s));
          const uint64_t bdesc = make_sw128_kmajor_desc(b_addr(s));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          tc_commit_mc(smem_u32(&sh->empty[s]), CL_MASK);     // frees the slot in all four CTAs once these MMAs retire
          if (kb == num_kb - 1) tc_commit(smem_u32(&sh->tfull[acc]));
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue: bias + residual, row statistics across the cluster, LayerNorm
    const int ew = warp - 4;
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
    const int half = ew >> 2;                         // column chunks half, half + 2, half + 4 (32 columns each)
    const int etid = threadIdx.x - 128;               // 0..255
    const int row_in_tile = quarter * 32 + lane;
    int it = 0;
    for (int mt = cluster_id; mt < p.num_m_tiles; mt += n_clusters, ++it) {
      const int g = group_of(mt);
      const int acc = it & 1, buf = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      const long long row = (long long)mt * BM + row_in_tile;
      const bool valid = row < p.M;
      // stage this CTA's slice of bias / gamma / beta (the previous row block is done with the old values)
      epi_bar_sync();
      for (int i = etid; i < BN; i += 256) {
        const int c = g * NTOT + (int)rank * BN + i;
        sh->bias[i] = p.bias ? p.bias[c] : 0.f;
        sh->gamma[i] = p.gamma[c];
        sh->beta[i] = p.beta[c];
      }
      epi_bar_sync();
      mbar_wait(smem_u32(&sh->tfull[acc]), acc_ph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      float v[3][32];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int c0 = (half + 2 * k) * 32;           // column offset inside this CTA's 192
        uint32_t r[32];
        tmem_ld_32x32(tbase + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[k][j] = __uint_as_float(r[j]) + sh->bias[c0 + j];
        if (p.residual && valid && !(p.debug & 1)) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual + row * p.ldr + (long long)rank * BN + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q = rp[j];
            v[k][4 * j] += q.x; v[k][4 * j + 1] += q.y; v[k][4 * j + 2] += q.z; v[k][4 * j + 3] += q.w;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sh->tempty[acc]));          // accumulator drained: the next main loop may start
      if (p.pre32 && valid && !(p.debug & 2)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float4* op = reinterpret_cast<float4*>(p.pre32 + row * NTOT + (long long)rank * BN + (half + 2 * k) * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) op[j] = make_float4(v[k][4 * j], v[k][4 * j + 1], v[k][4 * j + 2], v[k][4 * j + 3]);
        }
      }
      // local statistics of this thread's 96 values
      float s1 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 32; ++j) s1 += v[k][j];
      const float ml = s1 * (1.0f / 96.0f);
      float m2 = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[k][j] - ml; m2 = fmaf(d, d, m2); }
      sh->half_stats[half][row_in_tile] = make_float2(s1, m2);
      epi_bar_sync();
      if (half == 0 && !(p.debug & 4)) {
        // combine the two halves of the row (Chan), publish to every CTA of the cluster
        const float2 a = sh->half_stats[0][row_in_tile], b = sh->half_stats[1][row_in_tile];
        const float s = a.x + b.x;
        const float mean = s * (1.0f / (float)BN);
        const float da = a.x * (1.0f / 96.0f) - mean, db = b.x * (1.0f / 96.0f) - mean;
        const float M2 = a.y + b.y + 96.0f * (da * da + db * db);
        const uint32_t slot = smem_u32(&sh->cl_stats[buf][rank][row_in_tile]);
        const uint32_t sbar = smem_u32(&sh->stats[buf]);
#pragma unroll
        for (uint32_t dst = 0; dst < NCL; ++dst) {
          st_cluster_f2(mapa(slot, dst), s, M2);
          mbar_arrive_remote(mapa(sbar, dst));        // release.cluster: the store above is visible to the waiter
        }
      }
      if (!(p.debug & 4)) mbar_wait_cluster(smem_u32(&sh->stats[buf]), acc_ph);
      float S = 0.f;
      float2 cs[NCL];
#pragma unroll
      for (int c = 0; c < NCL; ++c) { cs[c] = sh->cl_stats[buf][c][row_in_tile]; S += cs[c].x; }
      const float mean = S * (1.0f / (float)NTOT);
      float M2 = 0.f;
#pragma unroll
      for (int c = 0; c < NCL; ++c) {
        const float d = cs[c].x * (1.0f / (float)BN) - mean;
        M2 += cs[c].y + (float)BN * d * d;
      }
      const float rstd = rsqrtf(M2 * (1.0f / (float)NTOT) + p.eps);
      if (valid && !(p.debug & 2)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int c0 = (half + 2 * k) * 32;
          float y[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaf((v[k][j] - mean) * rstd, sh->gamma[c0 + j], sh->beta[c0 + j]);
          const long long off = row * NTOT + (long long)rank * BN + c0;
          if (p.y32) {
            float4* op = reinterpret_cast<float4*>(p.y32 + off);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
          }
          if (p.y16) {
            uint4* op = reinterpret_cast<uint4*>(p.y16 + off);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              op[j] = make_uint4(pack_bf16x2(y[8 * j], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                                 pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7]));
          }
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                                 // no CTA leaves while peers may still multicast / signal into it
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

int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_inner,
             uint32_t box_rows) {
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vi_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VI_ERR_CUDA;
  }
  return VI_OK;
}

}  // namespace

extern "C" int vi_gemm_ln_bf16(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual, int64_t ldr,
                               const float* gamma, const float* beta, float eps, float* pre32, float* y32, void* y16, int M, int K,
                               int n_groups, const int32_t* group_row_end, vi_stream_t stream) {
  VI_CHECK_ARG(x && w && gamma && beta && (y32 || y16), "vi_gemm_ln_bf16: null operand");
  VI_CHECK_ARG(M > 0 && K > 0 && K % BK == 0, "vi_gemm_ln_bf16: bad sizes M=%d K=%d (K %% 64 == 0)", M, K);
  VI_CHECK_ARG(ldx % 8 == 0 && ldx >= K && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0,
               "vi_gemm_ln_bf16: operands must be 16-byte aligned, ldx >= K and a multiple of 8");
  VI_CHECK_ARG(!residual || (ldr >= NTOT && ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0),
               "vi_gemm_ln_bf16: residual must be 16-byte aligned with ldr >= 768, ldr %% 4 == 0");
  VI_CHECK_ARG((((uintptr_t)pre32 | (uintptr_t)y32 | (uintptr_t)y16) & 15) == 0, "vi_gemm_ln_bf16: outputs must be 16-byte aligned");
  VI_CHECK_ARG(n_groups >= 1 && n_groups <= MAX_GROUPS && (n_groups == 1 || group_row_end), "vi_gemm_ln_bf16: bad row groups");
  if (int rc = resolve_encode()) return rc;
  RbParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.pre32 = pre32; p.y32 = y32; p.y16 = reinterpret_cast<bf16*>(y16);
  p.M = M; p.K = K; p.n_groups = n_groups;
  p.num_m_tiles = (M + BM - 1) / BM;
  if (const char* e = getenv("VI_RB_DEBUG")) p.debug = atoi(e);
  for (int g = 0; g < n_groups; ++g) {
    if (n_groups == 1) { p.group_tile_end[g] = p.num_m_tiles; break; }
    const int e = group_row_end[g];
    VI_CHECK_ARG(e > 0 && e <= M && (g == 0 || e > group_row_end[g - 1]), "vi_gemm_ln_bf16: bad group_row_end[%d]=%d", g, e);
    VI_CHECK_ARG(g == n_groups - 1 || e % BM == 0, "vi_gemm_ln_bf16: group %d must end on a multiple of %d rows", g, BM);
    p.group_tile_end[g] = (e + BM - 1) / BM;
  }
  VI_CHECK_ARG(n_groups == 1 || group_row_end[n_groups - 1] == M, "vi_gemm_ln_bf16: last group must end at M");
  CUtensorMap tmA, tmB;
  if (int rc = make_map(&tmA, x, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, BK, BM / NCL)) return rc;
  if (int rc = make_map(&tmB, w, (uint64_t)K, (uint64_t)n_groups * NTOT, (uint64_t)K, BK, BN)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    VI_CUDA(cudaFuncSetAttribute(gemm_rowblock_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  // Clusters of 4 cannot use every SM (GPCs hold 16 / 18 / 20 SMs: 33 co-resident clusters on a B200, not 148 / 4 = 37).
  // The grid is clamped to what is co-resident so that the row blocks beyond it are walked by the persistent loop
  // (main loop of block i + 1 under the epilogue of block i) instead of waiting for a whole cluster to retire.
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t q;
    memset(&q, 0, sizeof(q));
    q.gridDim = dim3((unsigned)(vi_num_sms() / NCL * NCL));
    q.blockDim = dim3(NUM_THREADS);
    q.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = NCL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa;
    q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_rowblock_ln_kernel, &q) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = vi_num_sms() / NCL;
    }
    if (const char* e = getenv("VI_RB_CLUSTERS")) { const int v = atoi(e); if (v > 0) n = v; }
    max_clusters = n;
    if (getenv("VI_RB_VERBOSE")) fprintf(stderr, "vi_gemm_ln_bf16: %d co-resident clusters of %d CTAs\n", n, NCL);
  }
  const int n_clusters = p.num_m_tiles < max_clusters ? p.num_m_tiles : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(n_clusters * NCL));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  if (vi_pdl_enabled()) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  VI_CUDA(cudaLaunchKernelEx(&cfg, gemm_rowblock_ln_kernel, tmA, tmB, p));
  return VI_OK;
}
