// vi_gemm_tc.cu - bf16 GEMM on the 5th-gen tensor cores: Y = epi(X W^T + bias) + residual.
//
// One persistent CTA per SM walks output tiles (static round-robin).  Per CTA:
//   warp 0      : TMA producer  - cp.async.bulk.tensor 2D loads of the X (128 x 64) and W (BN x 64)
//                 k-blocks into a STAGES-deep shared-memory ring (128B swizzle), mbarrier tx counts;
//   warp 1      : MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per k-block,
//                 accumulating fp32 in TMEM; tcgen05.commit releases ring slots / publishes tiles;
//   warp 2      : TMEM allocator (2 x BN columns = double-buffered accumulator);
//   warps 4..7  : epilogue      - tcgen05.ld 32 lanes x 32 columns, +bias, GELU/ReLU, +residual,
//                 16-byte stores as bf16 or fp32; overlaps the next tile's main loop.
// The grouped form lets row ranges of X use different weight blocks of a stacked W (DUET's
// global/local encoders, HAMT's language/vision streams) in a single launch.
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
constexpr int NUM_THREADS = 256;

struct GemmParams {
  const float* bias;
  const float* residual;
  long long ldr;
  void* y;
  long long ldy;
  int y_f32;
  int M, N, K;
  int epilogue;
  int n_groups;
  int group_tile_end[MAX_GROUPS];
  int num_m_tiles, num_n_tiles;
};

__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t saddr) {
  // UMMA shared-memory descriptor, K-major operand, 128-byte swizzle:
  //   bits [0,14) start address >> 4; [16,30) leading byte offset >> 4 (unused for swizzled
  //   K-major, 1); [32,46) stride byte offset >> 4 = 1024 B between 8-row groups;
  //   [46,48) descriptor version 1 (sm_100); [61,64) layout type 2 = SWIZZLE_128B.
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  // kind::f16 instruction descriptor: D=f32 (bits 4-5 =1), A=B=bf16 (bits 7-9, 10-12 = 1),
  // both operands K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;           // 128 / 256 / 512: powers of two >= 32
  static_assert(BN == 64 || BN == 128 || BN == 256, "tile N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto a_addr = [&](int s) { return smem_base + (uint32_t)(s * STAGE_BYTES); };
  auto b_addr = [&](int s) { return smem_base + (uint32_t)(s * STAGE_BYTES + A_BYTES); };
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * STAGES + 2 + a); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);     // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = p.K / BK;

  auto group_of = [&](int mt) {
    int g = 0;
    while (g < p.n_groups - 1 && mt >= p.group_tile_end[g]) ++g;
    return g;
  };

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t / p.num_n_tiles, nt = t % p.num_n_tiles;
        const int wrow = group_of(mt) * p.N + nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), STAGE_BYTES);
          tma_load_2d(a_addr(s), &tmA, full_bar(s), kb * BK, mt * BM);
          tma_load_2d(b_addr(s), &tmB, full_bar(s), kb * BK, wrow);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc<BN>();
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_ph ^ 1u);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(s), ph);                 // TMA bytes have landed
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(a_addr(s));
          const uint64_t bdesc = make_sw128_kmajor_desc(b_addr(s));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in 16-byte units
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                        (uint32_t)((kb | k) != 0));
          }
          tc_commit(empty_bar(s));                    // frees the ring slot when the MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        tc_commit(tfull_bar(acc));                    // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------------------
    const int ew = warp & 3;                          // TMEM lane quarter this warp may read
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int mt = t / p.num_n_tiles, nt = t % p.num_n_tiles;
      const int g = group_of(mt);
      const int acc = it & 1;
      const uint32_t acc_ph = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_ph);
      tc_fence_after();
      const long long row = (long long)mt * BM + ew * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        tmem_ld_wait();
        const int col = nt * BN + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + (long long)g * p.N + col);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if (p.epilogue == VI_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (p.epilogue == VI_EPI_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (row_ok) {
          if (p.residual) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + row * p.ldr + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(r4 + j);
              v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
          }
          if (p.y_f32) {
            float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + row * p.ldy + col);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.y) + row * p.ldy + col);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
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

// 2D bf16 tensor map: inner dim = K (contiguous), outer = rows, 128B swizzle, box = 64 x box_rows
int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld_elems, uint32_t box_rows) {
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstride[1] = {ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vi_set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p inner=%llu rows=%llu ld=%llu box_rows=%u)",
                 (int)r, ptr, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)ld_elems, box_rows);
    return VI_ERR_CUDA;
  }
  return VI_OK;
}

template <int BN, int STAGES>
constexpr int smem_bytes() { return STAGES * (BM * BK * 2 + BN * BK * 2) + (2 * STAGES + 4) * 8 + 16 + 1024; }

template <int BN, int STAGES>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid, cudaStream_t st) {
  static bool attr_set = false;          // idempotent; races only repeat the same call
  if (!attr_set) {
    VI_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 smem_bytes<BN, STAGES>()));
    attr_set = true;
  }
  gemm_bf16_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, smem_bytes<BN, STAGES>(), st>>>(tmA, tmB, p);
  VI_LAUNCH_CHECK();
  return VI_OK;
}

int pick_bn(int m_tiles, int N, int nsm) {
  if (const char* e = getenv("VI_GEMM_BN")) {
    int v = atoi(e);
    if ((v == 64 || v == 128 || v == 256) && N % v == 0) return v;
  }
  int best = 64;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (N % bn) continue;
    const long long tiles = (long long)m_tiles * (N / bn);
    const long long waves = (tiles + nsm - 1) / nsm;
    const double cost = (double)waves * (bn + 48);   // per-tile time ~ BN plus a fixed prologue/epilogue share
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

extern "C" int vi_gemm_bf16(const void* x, int64_t ldx, const void* w, const float* bias, const float* residual,
                            int64_t ldr, void* y, int64_t ldy, int y_dtype, int M, int N, int K, int epilogue,
                            int n_groups, const int32_t* group_row_end, vi_stream_t stream) {
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
  VI_CHECK_ARG(y_dtype == VI_DT_BF16 || y_dtype == VI_DT_F32, "vi_gemm_bf16: bad y_dtype %d", y_dtype);
  VI_CHECK_ARG(epilogue >= VI_EPI_NONE && epilogue <= VI_EPI_RELU, "vi_gemm_bf16: bad epilogue %d", epilogue);
  VI_CHECK_ARG(n_groups >= 1 && n_groups <= MAX_GROUPS, "vi_gemm_bf16: n_groups=%d out of range", n_groups);
  VI_CHECK_ARG(n_groups == 1 || group_row_end, "vi_gemm_bf16: grouped call without group_row_end");
  if (int rc = resolve_encode()) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.y = y; p.ldy = ldy; p.y_f32 = (y_dtype == VI_DT_F32);
  p.M = M; p.N = N; p.K = K; p.epilogue = epilogue; p.n_groups = n_groups;
  p.num_m_tiles = (M + BM - 1) / BM;
  for (int g = 0; g < n_groups; ++g) {
    if (n_groups == 1) { p.group_tile_end[g] = p.num_m_tiles; break; }
    const int e = group_row_end[g];
    VI_CHECK_ARG(e > 0 && e <= M && (g == 0 || e > group_row_end[g - 1]), "vi_gemm_bf16: bad group_row_end[%d]=%d", g, e);
    VI_CHECK_ARG(g == n_groups - 1 || e % BM == 0, "vi_gemm_bf16: group %d must end on a multiple of %d rows (got %d)", g, BM, e);
    p.group_tile_end[g] = (e + BM - 1) / BM;
  }
  VI_CHECK_ARG(n_groups == 1 || group_row_end[n_groups - 1] == M, "vi_gemm_bf16: last group must end at M");

  const int nsm = vi_num_sms();
  const int bn = pick_bn(p.num_m_tiles, N, nsm);
  p.num_n_tiles = N / bn;
  const long long tiles = (long long)p.num_m_tiles * p.num_n_tiles;
  const int grid = (int)(tiles < nsm ? tiles : nsm);

  CUtensorMap tmA, tmB;
  if (int rc = make_map(&tmA, x, (uint64_t)K, (uint64_t)M, (uint64_t)ldx, BM)) return rc;
  if (int rc = make_map(&tmB, w, (uint64_t)K, (uint64_t)n_groups * N, (uint64_t)K, (uint32_t)bn)) return rc;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return launch<256, 4>(tmA, tmB, p, grid, st);
    case 128: return launch<128, 6>(tmA, tmB, p, grid, st);
    default:  return launch<64, 8>(tmA, tmB, p, grid, st);
  }
}
