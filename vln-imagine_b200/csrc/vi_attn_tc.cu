// vi_attn_tc.cu - fused masked multi-head attention on the 5th-gen tensor cores (inference path).
//
//   O = softmax(Q K^T / sqrt(64) + key_padding + (w * pair_dist + b)) V        per (episode, head)
//
// The sequences of this path are short (30 - 212 tokens), so one (episode, head) problem is a 64 x Lk score tile.
// tcgen05.mma with M = 64 fills only lanes 0-15 of every TMEM lane quarter; a SECOND M = 64 accumulator placed 16 lanes
// higher fills the other half, so one work item is a PAIR of adjacent heads of one episode (same masks, same lengths):
// all 128 TMEM lanes and all 128 softmax threads are busy, thread t of warp w owning query row 16 w + (t & 15) of head
// (t >> 4) of the pair.
//
//   one elected thread : TMA loads (cp.async.bulk.tensor, 128B swizzle) of Q (64 x 64), K and V (LkP x 64) of both heads;
//                        S = Q K^T  : tcgen05.mma kind::f16, M = 64, N = LkP, K = 64, Q / K as K-major shared-memory operands;
//                        O = P V    : M = 64, N = 64, K = LkP, P from shared memory (K-major), V as an MN-major operand
//                                     (the [key][d] tile TMA wrote is used as it lies: no transpose anywhere);
//   128 threads        : tcgen05.ld of their score row in 32-column chunks, twice (row maximum, then exponentials): scale,
//                        key-padding mask, GASA bias, exp2; the probabilities go to shared memory in the 16-bit operand format
//                        (swizzled K-major) for the second contraction; after it the O row is read back from TMEM, divided by
//                        the row sum and written to global memory (128 contiguous bytes per thread).
// Scores and probabilities never reach HBM.  Two or three CTAs share an SM (64 - 112 KB of shared memory, 128 - 256 TMEM
// columns each) and overlap each other's load / MMA / softmax phases; a CTA also issues the loads of its next item as soon as
// the second contraction has released the operand tiles.
//
// Not handled here (vi_attn.cu keeps them): the fp32 check mode, attention-probability dropout and the log-sum-exp output of
// the training path, sequences beyond 256 keys.
//
// Reference: BertSelfAttention / BertOutAttention (VLN-DUET/map_nav_src/models/vilmodel.py:118-134, 336-349), GASA bias
// (:392-394, :1145-1149), nn.MultiheadAttention with key_padding_mask (models/transformer.py:176-177).
#include "vi_common.cuh"

namespace {

constexpr int DH = 64;
constexpr int MAXP = VI_ATTN_MAX_PROBLEMS;
constexpr int TILE_Q = 64;

struct TcProblem {
  void* o; long long ldo;
  const uint8_t* key_mask;
  const float* pair_dist;
  const float* bias_affine;
  int B, Lq, Lk, LkP;          // LkP = Lk rounded up to 16
  int q_tiles;                 // ceil(Lq / 64)
  int item0;                   // first work item of this problem
};
struct TcParams {
  CUtensorMap tmQ[MAXP], tmK[MAXP], tmV[MAXP];
  TcProblem pr[MAXP];
  int n_problems, H, mask_mode, total_items, f16;
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  // shared-memory matrix descriptor, 128-byte swizzle, 1024 bytes between 8-row groups (K-major: groups along M / N;
  // MN-major: groups along K), descriptor version 1 (sm_100)
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_m64(int n, bool f16, bool b_mn_major) {
  // kind::f16: D = f32, A / B formats (0 = fp16, 1 = bf16), A K-major, B K-major or MN-major (bit 16), N >> 3, M >> 4
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | (b_mn_major ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(64 >> 4) << 24);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// KB = number of 64-key blocks the tiles are sized for (LkP <= 64 KB)
template <int KB>
__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int Q_BYTES = TILE_Q * 128;                  // one head: 64 rows x 64 x 16 bit
  constexpr int KV_BYTES = KB * 64 * 128;                // one head: up to 64 KB keys
  constexpr int P_BYTES = KB * TILE_Q * 128;             // one head: KB key blocks of 64 rows x 64 keys
  constexpr uint32_t S_COLS = 64 * KB;                   // score columns in TMEM; O follows at S_COLS
  constexpr uint32_t TMEM_COLS = (S_COLS + 64 <= 128) ? 128u : ((S_COLS + 64 <= 256) ? 256u : 512u);
  // no static shared memory: the dynamic window then starts at offset 0 and is 1024-byte aligned (128B-swizzled tiles)
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int TILE_BYTES = 2 * Q_BYTES + 4 * KV_BYTES + 2 * P_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TILE_BYTES);
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(smem + TILE_BYTES + 16);
  float* madd = reinterpret_cast<float*>(smem + TILE_BYTES + 32);           // [64 KB]

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t q_s = base, k_s = q_s + 2 * Q_BYTES, v_s = k_s + 2 * KV_BYTES, p_s = v_s + 2 * KV_BYTES;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
    for (int i = 0; i < p.n_problems; ++i) {
      tma_prefetch_desc(&p.tmQ[i]);
      tma_prefetch_desc(&p.tmK[i]);
      tma_prefetch_desc(&p.tmV[i]);
    }
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();

  const int HP = p.H >> 1;
  auto decode = [&](int item, int& pi, int& b, int& hp, int& qt) {
    pi = 0;
    while (pi < p.n_problems - 1 && item >= p.pr[pi + 1].item0) ++pi;
    int r = item - p.pr[pi].item0;
    qt = r % p.pr[pi].q_tiles; r /= p.pr[pi].q_tiles;
    hp = r % HP;
    b = r / HP;
  };
  auto issue_loads = [&](int item) {                     // one thread
    int pi, b, hp, qt;
    decode(item, pi, b, hp, qt);
    const TcProblem& pr = p.pr[pi];
    mbar_expect_tx(bar_full, (uint32_t)(2 * Q_BYTES + 4 * pr.LkP * 128));
    const int qrow = b * pr.Lq + qt * TILE_Q, krow = b * pr.Lk;
    for (int it = 0; it < 2; ++it) {
      const int col = (2 * hp + it) * DH;
      tma_load_2d(q_s + (uint32_t)(it * Q_BYTES), &p.tmQ[pi], bar_full, col, qrow);
      tma_load_2d(k_s + (uint32_t)(it * KV_BYTES), &p.tmK[pi], bar_full, col, krow);
      tma_load_2d(v_s + (uint32_t)(it * KV_BYTES), &p.tmV[pi], bar_full, col, krow);
    }
  };

  const int it_head = lane >> 4;                         // which head of the pair this thread serves
  const int row_in_tile = warp * 16 + (lane & 15);       // query row inside the 64-row tile
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);          // this warp's lane quarter
  const float LOG2E = 1.4426950408889634f;

  int n_done = 0;
  if (tid == 0 && (int)blockIdx.x < p.total_items) issue_loads(blockIdx.x);
  for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++n_done) {
    int pi, b, hp, qt;
    decode(item, pi, b, hp, qt);
    const TcProblem& pr = p.pr[pi];
    const int LkP = pr.LkP, Lk = pr.Lk;
    const int q = qt * TILE_Q + row_in_tile;             // query index inside the episode
    const bool q_ok = q < pr.Lq;
    // additive key-padding term of this episode (shared by both heads), in shared memory for broadcast reads
    for (int key = tid; key < LkP; key += 128) {
      float m = 0.f;
      if (key >= Lk) m = -INFINITY;
      else if (pr.key_mask && !pr.key_mask[(long long)b * Lk + key]) m = (p.mask_mode == VI_MASK_NEG_INF) ? -INFINITY : -10000.0f;
      madd[key] = m;
    }
    const float* dist = nullptr;
    float bw = 0.f, bb = 0.f;
    if (pr.pair_dist) {
      bw = pr.bias_affine[0];
      bb = pr.bias_affine[1];
      dist = pr.pair_dist + ((long long)b * pr.Lq + (q_ok ? q : 0)) * Lk;
    }
    mbar_wait(bar_full, (uint32_t)n_done & 1u);
    __syncthreads();                                     // madd is complete; everybody saw the operand tiles land
    tc_fence_after();
    if (tid == 0) {
      // S = Q K^T for both heads: accumulators at lanes +0 / +16, columns [0, LkP)
      const uint32_t idesc = idesc_m64(LkP, p.f16 != 0, false);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const uint64_t a = desc_sw128(q_s + (uint32_t)(it * Q_BYTES)), bd = desc_sw128(k_s + (uint32_t)(it * KV_BYTES));
        const uint32_t d = tmem_base + ((uint32_t)(it * 16) << 16);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) tc_mma_bf16(d, a + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
      }
      tc_commit(bar_mma);
    }
    mbar_wait(bar_mma, 0u);
    tc_fence_after();

    // ---- pass 1: row maximum of scale * s + mask + bias (log2 domain) ----------------------------------------------
    float mx = -INFINITY;
    for (int c0 = 0; c0 < LkP; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + (uint32_t)c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = c0 + j;
        if (key < LkP) {
          float s = fmaf(__uint_as_float(r[j]), 0.125f, madd[key]);
          if (dist && key < Lk) s += fmaf(bw, dist[key], bb);
          mx = fmaxf(mx, s);
        }
      }
    }
    const float mu = (mx == -INFINITY) ? 0.f : mx * LOG2E;
    // ---- pass 2: probabilities -> shared memory (operand A of the second contraction), row sum ----------------------
    float sum = 0.f;
    const uint32_t p_row = p_s + (uint32_t)(it_head * P_BYTES) + (uint32_t)((row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128);
    for (int c0 = 0; c0 < LkP; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + (uint32_t)c0, r);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int key = c0 + j;
        float s = fmaf(__uint_as_float(r[j]), 0.125f, key < LkP ? madd[key] : -INFINITY);
        if (dist && key < Lk) s += fmaf(bw, dist[key], bb);
        e[j] = ex2(fmaf(s, LOG2E, -mu));
        sum += e[j];
      }
      // 32 keys = 64 bytes = four 16-byte chunks of the 128-byte row of key block c0 / 64
      const uint32_t blk = p_row + (uint32_t)((c0 >> 6) * (TILE_Q * 128));
      const int ch0 = (c0 & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        if (c0 + jj * 8 < LkP) {
          const uint32_t addr = blk + (uint32_t)(((ch0 + jj) ^ (row_in_tile & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};"
                       ::"r"(addr), "r"(pack_h16x2(e[8 * jj], e[8 * jj + 1], p.f16 != 0)), "r"(pack_h16x2(e[8 * jj + 2], e[8 * jj + 3], p.f16 != 0)),
                       "r"(pack_h16x2(e[8 * jj + 4], e[8 * jj + 5], p.f16 != 0)), "r"(pack_h16x2(e[8 * jj + 6], e[8 * jj + 7], p.f16 != 0))
                       : "memory");
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      // O = P V: A = P (K-major, key blocks of 64), B = V as it lies in shared memory ([key][d], MN-major), N = 64
      const uint32_t idesc = idesc_m64(DH, p.f16 != 0, true);
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const uint32_t d = tmem_base + ((uint32_t)(it * 16) << 16) + S_COLS;
        for (int k = 0; k < LkP / 16; ++k) {
          const uint64_t a = desc_sw128(p_s + (uint32_t)(it * P_BYTES + (k >> 2) * (TILE_Q * 128))) + (uint64_t)(2 * (k & 3));
          const uint64_t bd = desc_sw128(v_s + (uint32_t)(it * KV_BYTES + k * 2048));
          tc_mma_bf16(d, a, bd, idesc, (uint32_t)(k != 0));
        }
      }
      tc_commit(bar_mma);
    }
    mbar_wait(bar_mma, 1u);
    tc_fence_after();
    // the operand tiles are free again: start the loads of this CTA's next item under the epilogue
    const int next = item + (int)gridDim.x;
    if (tid == 0 && next < p.total_items) issue_loads(next);

    // ---- epilogue: O row / row sum -> global ------------------------------------------------------------------------
    const float inv = 1.0f / sum;
    uint8_t* orow = reinterpret_cast<uint8_t*>(pr.o) + (((long long)b * pr.Lq + q) * pr.ldo + (2 * hp + it_head) * DH) * 2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld_32x32(t_lane + S_COLS + (uint32_t)(half * 32), r);
      tmem_ld_wait();
      if (q_ok) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 v;
          v.x = pack_h16x2(__uint_as_float(r[8 * jj]) * inv, __uint_as_float(r[8 * jj + 1]) * inv, p.f16 != 0);
          v.y = pack_h16x2(__uint_as_float(r[8 * jj + 2]) * inv, __uint_as_float(r[8 * jj + 3]) * inv, p.f16 != 0);
          v.z = pack_h16x2(__uint_as_float(r[8 * jj + 4]) * inv, __uint_as_float(r[8 * jj + 5]) * inv, p.f16 != 0);
          v.w = pack_h16x2(__uint_as_float(r[8 * jj + 6]) * inv, __uint_as_float(r[8 * jj + 7]) * inv, p.f16 != 0);
          *reinterpret_cast<uint4*>(orow + half * 64 + jj * 16) = v;
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                     // TMEM and madd may be overwritten by the next item
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int KB> constexpr int tc_smem_bytes() { return 2 * TILE_Q * 128 + 4 * KB * 64 * 128 + 2 * KB * TILE_Q * 128 + 32 + 64 * KB * 4; }

template <int KB>
int launch_tc(const TcParams& p, int ctas_per_sm, cudaStream_t st) {
  const int cap = vi_num_sms() * ctas_per_sm;
  const int grid = p.total_items < cap ? p.total_items : cap;
  VI_CUDA(vi_launch(attn_fwd_tc_kernel<KB>, dim3((unsigned)grid), dim3(128), (size_t)tc_smem_bytes<KB>(), st, p));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

}  // namespace

// 1 when the tcgen05 kernel can run these problems (shape / alignment / feature limits)
int vi_attn_tc_supported(const vi_attn_problem* pr, int n, int H, int dtype) {
  if (dtype != VI_DT_BF16 && dtype != VI_DT_F16) return 0;
  if (H % 2 != 0) return 0;
  for (int i = 0; i < n; ++i) {
    if (pr[i].Lk > 256 || pr[i].drop_p > 0.f || pr[i].lse) return 0;
    if ((pr[i].ldq % 8) || (pr[i].ldk % 8) || (pr[i].ldv % 8) || (pr[i].ldo % 8)) return 0;
    if ((((uintptr_t)pr[i].q | (uintptr_t)pr[i].k | (uintptr_t)pr[i].v | (uintptr_t)pr[i].o) & 15) != 0) return 0;
  }
  return 1;
}
// 1 when vi_attn_fwd_multi should route these problems to the tcgen05 kernel.  Measured on B200 at the cfg-2 shapes (30 / 37
// queries x 85 keys): 38 us per launch against 22 us for the mma.sync kernel - a 64-row tcgen05 tile leaves 45 % of its
// softmax threads without a query row and two CTAs of four warps per SM cannot hide the TMEM / MUFU latencies of a whole
// score row per thread - so the dispatcher keeps the mma.sync kernel unless VI_ATTN_TC=1 asks for this one.
int vi_attn_tc_eligible(const vi_attn_problem* pr, int n, int H, int dtype) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VI_ATTN_TC"); on = (e && e[0] == '1') ? 1 : 0; }
  return on && vi_attn_tc_supported(pr, n, H, dtype);
}

int vi_attn_tc_launch(const vi_attn_problem* problems, int n_problems, int H, int dtype, int mask_mode, cudaStream_t st) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_problems = n_problems; p.H = H; p.mask_mode = mask_mode; p.f16 = dtype == VI_DT_F16;
  int items = 0, max_lkp = 0;
  for (int i = 0; i < n_problems; ++i) {
    const vi_attn_problem& a = problems[i];
    TcProblem& o = p.pr[i];
    o.o = a.o; o.ldo = a.ldo; o.key_mask = a.key_mask; o.pair_dist = a.pair_dist; o.bias_affine = a.bias_affine;
    o.B = a.B; o.Lq = a.Lq; o.Lk = a.Lk; o.LkP = (a.Lk + 15) & ~15;
    o.q_tiles = (a.Lq + TILE_Q - 1) / TILE_Q;
    o.item0 = items;
    items += a.B * (H / 2) * o.q_tiles;
    max_lkp = o.LkP > max_lkp ? o.LkP : max_lkp;
    const uint64_t cols = (uint64_t)H * DH;
    if (int rc = vi_make_tmap_h16(&p.tmQ[i], a.q, p.f16, cols, (uint64_t)a.B * a.Lq, (uint64_t)a.ldq, TILE_Q)) return rc;
    if (int rc = vi_make_tmap_h16(&p.tmK[i], a.k, p.f16, cols, (uint64_t)a.B * a.Lk, (uint64_t)a.ldk, (uint32_t)o.LkP)) return rc;
    if (int rc = vi_make_tmap_h16(&p.tmV[i], a.v, p.f16, cols, (uint64_t)a.B * a.Lk, (uint64_t)a.ldv, (uint32_t)o.LkP)) return rc;
  }
  p.total_items = items;
  if (max_lkp <= 64) return launch_tc<1>(p, 3, st);
  if (max_lkp <= 128) return launch_tc<2>(p, 2, st);
  if (max_lkp <= 192) return launch_tc<3>(p, 1, st);
  return launch_tc<4>(p, 1, st);
}

int vi_attn_tc_init() {
  VI_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<1>()));
  VI_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<2>()));
  VI_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<3>()));
  VI_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes<4>()));
  return VI_OK;
}

// The tcgen05 kernel by name (tests, A/B timing): same contract as vi_attn_fwd_multi for the problems it supports.
extern "C" int vi_attn_fwd_tc(const vi_attn_problem* problems, int n_problems, int H, int dtype, int mask_mode, vi_stream_t stream) {
  VI_CHECK_ARG(problems && n_problems >= 1 && n_problems <= VI_ATTN_MAX_PROBLEMS, "vi_attn_fwd_tc: 1..%d problems", VI_ATTN_MAX_PROBLEMS);
  VI_CHECK_ARG(mask_mode == VI_MASK_ADD_NEG10000 || mask_mode == VI_MASK_NEG_INF, "vi_attn_fwd_tc: bad mask_mode");
  for (int i = 0; i < n_problems; ++i) {
    const vi_attn_problem& a = problems[i];
    VI_CHECK_ARG(a.q && a.k && a.v && a.o && a.B > 0 && a.Lq > 0 && a.Lk > 0, "vi_attn_fwd_tc: bad problem %d", i);
    VI_CHECK_ARG(!a.pair_dist || a.bias_affine, "vi_attn_fwd_tc: pair_dist needs bias_affine {w,b}");
  }
  VI_CHECK_ARG(vi_attn_tc_supported(problems, n_problems, H, dtype),
               "vi_attn_fwd_tc: needs 16-bit operands, an even head count, <= 256 keys, 16-byte aligned rows, no dropout / lse");
  return vi_attn_tc_launch(problems, n_problems, H, dtype, mask_mode, reinterpret_cast<cudaStream_t>(stream));
}
