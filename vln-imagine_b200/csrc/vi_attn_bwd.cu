// vi_attn_bwd.cu - attention backward on bf16 operands with warp-level tensor-core tiles (fine-tuning path).
//
// One CTA (4 warps) per (episode, head).  Q, K, V, dO of the head are staged in shared memory with cp.async.
//   phase 1 (a warp per 16-query tile):  S = Q K^T / 8 + mask (+ GASA bias), P = softmax(S), dP = dO V^T,
//            D = rowsum(P o dP), dS = P o (dP - D), dQ = dS K / 8;  P and dS are parked in shared memory (bf16);
//   phase 2 (a warp per 16-key tile):    dV = P^T dO,  dK = dS^T Q / 8  (ldmatrix.trans gives the transposed operands).
// All contractions are mma.sync m16n8k16 (bf16 in, fp32 accumulate); softmax and the dS algebra are fp32.
// Sequences on this path are short (<= 100 queries, <= 128 keys), so nothing is tiled over keys: the whole score row of
// a query tile lives in registers.  Longer key sequences take the fp32 kernel of vi_bwd.cu.
// GASA: d(w), d(b) of the affine bias w * dist + b are accumulated with two fp32 atomics per CTA.
//
// Reference semantics: torch.autograd of BertSelfAttention / BertOutAttention (VLN-DUET/map_nav_src/models/vilmodel.py:
// 118-134, 336-349) and of nn.MultiheadAttention (models/transformer.py:176-177).
#include "vi_common.cuh"

namespace {

constexpr int DH = 64;
constexpr int ROW = 72;            // bf16 elements per staged row (64 + 8): conflict-free ldmatrix

struct AttnBwdTcParams {
  const bf16* q; long long ldq;
  const bf16* k; long long ldk;
  const bf16* v; long long ldv;
  const bf16* dout; long long ldo;
  bf16* dq; long long lddq;
  bf16* dk; long long lddk;
  bf16* dv; long long lddv;
  const uint8_t* key_mask;
  const float* pair_dist;
  const float* bias_affine;
  float* d_affine;
  int B, H, Lq, Lk, LqP, LkP, mask_mode;
  uint32_t drop_thresh;        // 0 = no attention dropout
  float drop_scale;
  uint32_t drop_site;
  const uint32_t* drop_seed;
};

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;       // src-size 0 zero-fills the 16 bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

// NT = 8-key score tiles per query row held in registers (LkP <= 8 * NT)
// CTAs per SM asked of the register allocator: the grid is episodes x heads (768 CTAs at the fine-tuning shapes), and at the
// 195 registers NT = 12 takes unconstrained only two CTAs fit an SM (2.6 waves); 167 registers (no spills) fit three.
template <int NT>
__global__ void __launch_bounds__(128, NT == 12 ? 3 : NT == 6 ? 4 : 1) attn_bwd_bf16_kernel(const AttnBwdTcParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t smem[];
  const int LqP = p.LqP, LkP = p.LkP;
  const int PP = LkP + 8;                                    // pitch of the P / dS planes (bf16 elements)
  bf16* Ks = reinterpret_cast<bf16*>(smem);                  // [LkP][72]
  bf16* Vs = Ks + (size_t)LkP * ROW;                         // [LkP][72]
  bf16* Qs = Vs + (size_t)LkP * ROW;                         // [LqP][72]
  bf16* Os = Qs + (size_t)LqP * ROW;                         // [LqP][72]   dO
  bf16* Ps = Os + (size_t)LqP * ROW;                         // [LqP][PP]
  bf16* Ss = Ps + (size_t)LqP * PP;                          // [LqP][PP]   dS
  float* madd = reinterpret_cast<float*>(Ss + (size_t)LqP * PP);   // [LkP]
  const int b = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const bf16* qg = p.q + (long long)b * p.Lq * p.ldq + h * DH;
  const bf16* kg = p.k + (long long)b * p.Lk * p.ldk + h * DH;
  const bf16* vg = p.v + (long long)b * p.Lk * p.ldv + h * DH;
  const bf16* og = p.dout + (long long)b * p.Lq * p.ldo + h * DH;
  const uint32_t ks_u = smem_u32(Ks), vs_u = smem_u32(Vs), qs_u = smem_u32(Qs), os_u = smem_u32(Os);
  const uint32_t ps_u = smem_u32(Ps), ss_u = smem_u32(Ss);
  for (int e = tid; e < LkP * 8; e += 128) {
    const int key = e >> 3, ch = e & 7;
    const bool ok = key < p.Lk;
    const int kk = ok ? key : 0;
    cp_async16(ks_u + (uint32_t)(key * ROW + ch * 8) * 2, kg + (long long)kk * p.ldk + ch * 8, ok);
    cp_async16(vs_u + (uint32_t)(key * ROW + ch * 8) * 2, vg + (long long)kk * p.ldv + ch * 8, ok);
  }
  for (int e = tid; e < LqP * 8; e += 128) {
    const int r = e >> 3, ch = e & 7;
    const bool ok = r < p.Lq;
    const int rr = ok ? r : 0;
    cp_async16(qs_u + (uint32_t)(r * ROW + ch * 8) * 2, qg + (long long)rr * p.ldq + ch * 8, ok);
    cp_async16(os_u + (uint32_t)(r * ROW + ch * 8) * 2, og + (long long)rr * p.ldo + ch * 8, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int key = tid; key < LkP; key += 128) {
    float m = 0.f;
    if (key >= p.Lk) m = -INFINITY;
    else if (p.key_mask && !p.key_mask[(long long)b * p.Lk + key]) m = (p.mask_mode == VI_MASK_NEG_INF) ? -INFINITY : -10000.0f;
    madd[key] = m;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int g = lane >> 2, tg = lane & 3;
  const int lm = lane >> 3, lr = lane & 7;                   // ldmatrix: this lane addresses row lr of matrix lm
  const int nt_run = LkP >> 3;                               // 8-key tiles actually present (even)
  float bw = 0.f, bb = 0.f;
  if (p.pair_dist) { bw = p.bias_affine[0]; bb = p.bias_affine[1]; }
  float aw = 0.f, ab = 0.f;                                  // GASA affine gradients of this thread
  const bool drop = p.drop_thresh != 0;
  const uint32_t dkey = drop ? vi_drop_key(p.drop_seed, p.drop_site) : 0u;
  // dropout element id = ((b * H + h) * Lq + q) * LkP + key, the same as in the forward kernel
  const uint32_t dbase = (uint32_t)((b * p.H + h) * p.Lq) * (uint32_t)LkP;

  // ------------------------------------------------ phase 1: a warp per 16-query tile
  for (int qt = warp; qt < (LqP >> 4); qt += 4) {
    const int q0 = qt * 16;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    uint32_t qa[4][4], da[4][4];                             // Q and dO fragments of the tile, 4 k-steps over d
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t off = (uint32_t)((q0 + (lm & 1) * 8 + lr) * ROW + (lm >> 1) * 8 + ks * 16) * 2;
      ldsm_x4(qa[ks], qs_u + off);
      ldsm_x4(da[ks], os_u + off);
    }
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    }
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      if (2 * np < nt_run) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t kb[4], vb[4];   // (keys 0-7, d 0-7) (keys 0-7, d 8-15) (keys 8-15, d 0-7) (keys 8-15, d 8-15)
          const uint32_t off = (uint32_t)((np * 16 + (lm >> 1) * 8 + lr) * ROW + ks * 16 + (lm & 1) * 8) * 2;
          ldsm_x4(kb, ks_u + off);
          ldsm_x4(vb, vs_u + off);
          mma_16816(s[2 * np], qa[ks], kb[0], kb[1]);
          mma_16816(s[2 * np + 1], qa[ks], kb[2], kb[3]);
          mma_16816(dp[2 * np], da[ks], vb[0], vb[1]);
          mma_16816(dp[2 * np + 1], da[ks], vb[2], vb[3]);
        }
      }
    }
    const float* pd0 = nullptr;
    const float* pd1 = nullptr;
    if (p.pair_dist) {
      pd0 = p.pair_dist + ((long long)b * p.Lq + (r0 < p.Lq ? r0 : 0)) * p.Lk;
      pd1 = p.pair_dist + ((long long)b * p.Lq + (r1 < p.Lq ? r1 : 0)) * p.Lk;
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_run) {
        const int key = nt * 8 + 2 * tg;
        const float2 ma = *reinterpret_cast<const float2*>(madd + key);
        float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
        if (pd0) {
          if (key < p.Lk) { b00 = fmaf(bw, (*(pd0 + key)), bb); b10 = fmaf(bw, (*(pd1 + key)), bb); }
          if (key + 1 < p.Lk) { b01 = fmaf(bw, (*(pd0 + key + 1)), bb); b11 = fmaf(bw, (*(pd1 + key + 1)), bb); }
        }
        s[nt][0] = s[nt][0] * 0.125f + ma.x + b00;
        s[nt][1] = s[nt][1] * 0.125f + ma.y + b01;
        s[nt][2] = s[nt][2] * 0.125f + ma.x + b10;
        s[nt][3] = s[nt][3] * 0.125f + ma.y + b11;
        m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
        m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float mu0 = (m0 == -INFINITY) ? 0.f : m0, mu1 = (m1 == -INFINITY) ? 0.f : m1;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_run) {
        s[nt][0] = __expf(s[nt][0] - mu0); s[nt][1] = __expf(s[nt][1] - mu0);
        s[nt][2] = __expf(s[nt][2] - mu1); s[nt][3] = __expf(s[nt][3] - mu1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = l0 > 0.f ? 1.0f / l0 : 0.f, i1 = l1 > 0.f ? 1.0f / l1 : 0.f;
    float d0 = 0.f, d1 = 0.f;                                 // D = rowsum(P o dP)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_run) {
        s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1;
        if (drop) {      // O = (P o mask / (1 - p)) V: the mask scales dP here and P where it feeds dV below
          const uint32_t e0 = dbase + (uint32_t)r0 * (uint32_t)LkP + (uint32_t)(nt * 8 + 2 * tg);
          const uint32_t e1 = dbase + (uint32_t)r1 * (uint32_t)LkP + (uint32_t)(nt * 8 + 2 * tg);
          dp[nt][0] *= vi_hash32(e0, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
          dp[nt][1] *= vi_hash32(e0 + 1, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
          dp[nt][2] *= vi_hash32(e1, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
          dp[nt][3] *= vi_hash32(e1 + 1, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
        }
        d0 = fmaf(s[nt][0], dp[nt][0], fmaf(s[nt][1], dp[nt][1], d0));
        d1 = fmaf(s[nt][2], dp[nt][2], fmaf(s[nt][3], dp[nt][3], d1));
      }
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    // dS = P o (dP - D); park P and dS (bf16) for phase 2; GASA affine gradients
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_run) {
        const int key = nt * 8 + 2 * tg;
        dp[nt][0] = s[nt][0] * (dp[nt][0] - d0);
        dp[nt][1] = s[nt][1] * (dp[nt][1] - d0);
        dp[nt][2] = s[nt][2] * (dp[nt][2] - d1);
        dp[nt][3] = s[nt][3] * (dp[nt][3] - d1);
        if (pd0) {
          if (r0 < p.Lq) {
            if (key < p.Lk) { aw = fmaf(dp[nt][0], (*(pd0 + key)), aw); ab += dp[nt][0]; }
            if (key + 1 < p.Lk) { aw = fmaf(dp[nt][1], (*(pd0 + key + 1)), aw); ab += dp[nt][1]; }
          }
          if (r1 < p.Lq) {
            if (key < p.Lk) { aw = fmaf(dp[nt][2], (*(pd1 + key)), aw); ab += dp[nt][2]; }
            if (key + 1 < p.Lk) { aw = fmaf(dp[nt][3], (*(pd1 + key + 1)), aw); ab += dp[nt][3]; }
          }
        }
        if (drop) {      // dV = (P o mask / (1 - p))^T dO
          const uint32_t e0 = dbase + (uint32_t)r0 * (uint32_t)LkP + (uint32_t)key;
          const uint32_t e1 = dbase + (uint32_t)r1 * (uint32_t)LkP + (uint32_t)key;
          s[nt][0] *= vi_hash32(e0, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
          s[nt][1] *= vi_hash32(e0 + 1, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
          s[nt][2] *= vi_hash32(e1, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
          s[nt][3] *= vi_hash32(e1 + 1, dkey) >= p.drop_thresh ? p.drop_scale : 0.f;
        }
        *reinterpret_cast<uint32_t*>(Ps + (size_t)r0 * PP + key) = pack_bf16x2(s[nt][0], s[nt][1]);
        *reinterpret_cast<uint32_t*>(Ps + (size_t)r1 * PP + key) = pack_bf16x2(s[nt][2], s[nt][3]);
        *reinterpret_cast<uint32_t*>(Ss + (size_t)r0 * PP + key) = pack_bf16x2(dp[nt][0], dp[nt][1]);
        *reinterpret_cast<uint32_t*>(Ss + (size_t)r1 * PP + key) = pack_bf16x2(dp[nt][2], dp[nt][3]);
      }
    }
    // dQ = dS K / 8 : dS (accumulator layout) re-packed as the A operand; K^T fragments through ldmatrix.trans
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (2 * kk < nt_run) {
        uint32_t sa[4];
        sa[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
        sa[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
        sa[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
        sa[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
        for (int d2 = 0; d2 < 4; ++d2) {
          uint32_t kb[4];     // (keys 0-7, d 0-7)^T (keys 8-15, d 0-7)^T (keys 0-7, d 8-15)^T (keys 8-15, d 8-15)^T
          ldsm_x4_trans(kb, ks_u + (uint32_t)((kk * 16 + (lm & 1) * 8 + lr) * ROW + d2 * 16 + (lm >> 1) * 8) * 2);
          mma_16816(dq[2 * d2], sa, kb[0], kb[1]);
          mma_16816(dq[2 * d2 + 1], sa, kb[2], kb[3]);
        }
      }
    }
    bf16* dqg = p.dq + (long long)b * p.Lq * p.lddq + h * DH;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + 2 * tg;
      if (r0 < p.Lq) *reinterpret_cast<uint32_t*>(dqg + (long long)r0 * p.lddq + c) = pack_bf16x2(dq[dt][0] * 0.125f, dq[dt][1] * 0.125f);
      if (r1 < p.Lq) *reinterpret_cast<uint32_t*>(dqg + (long long)r1 * p.lddq + c) = pack_bf16x2(dq[dt][2] * 0.125f, dq[dt][3] * 0.125f);
    }
  }
  if (p.pair_dist && p.d_affine) {
    aw = warp_sum(aw);
    ab = warp_sum(ab);
    if (lane == 0) { atomicAdd(p.d_affine, aw); atomicAdd(p.d_affine + 1, ab); }
  }
  __syncthreads();

  // ------------------------------------------------ phase 2: a warp per 16-key tile
  bf16* dkg = p.dk + (long long)b * p.Lk * p.lddk + h * DH;
  bf16* dvg = p.dv + (long long)b * p.Lk * p.lddv + h * DH;
  for (int kt = warp; kt < (LkP >> 4); kt += 4) {
    const int k0 = kt * 16;
    float dv[8][4], dk[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    }
    for (int qb = 0; qb < (LqP >> 4); ++qb) {
      // A = P^T / dS^T (m = key, k = query): stored [query][key] -> transposed 8x8 blocks
      uint32_t pa[4], sa[4];
      const uint32_t aoff = (uint32_t)((qb * 16 + (lm >> 1) * 8 + lr) * PP + k0 + (lm & 1) * 8) * 2;
      ldsm_x4_trans(pa, ps_u + aoff);
      ldsm_x4_trans(sa, ss_u + aoff);
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t ob[4], qb4[4];   // B = dO / Q (k = query, n = d): stored [query][d] -> .trans
        const uint32_t boff = (uint32_t)((qb * 16 + (lm & 1) * 8 + lr) * ROW + d2 * 16 + (lm >> 1) * 8) * 2;
        ldsm_x4_trans(ob, os_u + boff);
        ldsm_x4_trans(qb4, qs_u + boff);
        mma_16816(dv[2 * d2], pa, ob[0], ob[1]);
        mma_16816(dv[2 * d2 + 1], pa, ob[2], ob[3]);
        mma_16816(dk[2 * d2], sa, qb4[0], qb4[1]);
        mma_16816(dk[2 * d2 + 1], sa, qb4[2], qb4[3]);
      }
    }
    const int r0 = k0 + g, r1 = k0 + g + 8;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int c = dt * 8 + 2 * tg;
      if (r0 < p.Lk) {
        *reinterpret_cast<uint32_t*>(dvg + (long long)r0 * p.lddv + c) = pack_bf16x2(dv[dt][0], dv[dt][1]);
        *reinterpret_cast<uint32_t*>(dkg + (long long)r0 * p.lddk + c) = pack_bf16x2(dk[dt][0] * 0.125f, dk[dt][1] * 0.125f);
      }
      if (r1 < p.Lk) {
        *reinterpret_cast<uint32_t*>(dvg + (long long)r1 * p.lddv + c) = pack_bf16x2(dv[dt][2], dv[dt][3]);
        *reinterpret_cast<uint32_t*>(dkg + (long long)r1 * p.lddk + c) = pack_bf16x2(dk[dt][2] * 0.125f, dk[dt][3] * 0.125f);
      }
    }
  }
}

template <int NT>
int launch_tc(const AttnBwdTcParams& p, size_t smem, cudaStream_t st) {
  static bool set = false;
  if (!set) {
    VI_CUDA(cudaFuncSetAttribute(attn_bwd_bf16_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    set = true;
  }
  VI_CUDA(vi_launch(attn_bwd_bf16_kernel<NT>, dim3(p.H, p.B), dim3(128), smem, st, p));
  return VI_OK;
}

}  // namespace

// Returns VI_OK after launching, or 1 when the problem does not fit the tensor-core kernel (the caller then uses the
// fp32-arithmetic kernel of vi_bwd.cu).
int vi_attn_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout,
                   int64_t ldo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                   const uint8_t* key_mask, const float* pair_dist, const float* bias_affine, float* d_affine, int B, int H,
                   int Lq, int Lk, int mask_mode, float drop_p, uint32_t drop_site, const uint32_t* drop_seed, cudaStream_t st) {
  const int LqP = (Lq + 15) & ~15, LkP = (Lk + 15) & ~15;
  if (LkP > 128 || LqP > 256) return 1;
  const uintptr_t al = (uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)dout;
  if ((al & 15) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return 1;
  if ((((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 3) || (lddq % 2) || (lddk % 2) || (lddv % 2)) return 1;
  const size_t smem = ((size_t)(2 * LkP + 2 * LqP) * ROW + (size_t)2 * LqP * (LkP + 8)) * 2 + (size_t)LkP * 4;
  if (smem > 200 * 1024) return 1;
  AttnBwdTcParams p;
  p.q = reinterpret_cast<const bf16*>(q); p.ldq = ldq; p.k = reinterpret_cast<const bf16*>(k); p.ldk = ldk;
  p.v = reinterpret_cast<const bf16*>(v); p.ldv = ldv; p.dout = reinterpret_cast<const bf16*>(dout); p.ldo = ldo;
  p.dq = reinterpret_cast<bf16*>(dq); p.lddq = lddq; p.dk = reinterpret_cast<bf16*>(dk); p.lddk = lddk;
  p.dv = reinterpret_cast<bf16*>(dv); p.lddv = lddv;
  p.key_mask = key_mask; p.pair_dist = pair_dist; p.bias_affine = bias_affine; p.d_affine = d_affine;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.LqP = LqP; p.LkP = LkP; p.mask_mode = mask_mode;
  p.drop_thresh = drop_p > 0.f ? vi_drop_threshold(drop_p) : 0u;
  if (drop_p > 0.f && p.drop_thresh == 0u) p.drop_thresh = 1u;
  p.drop_scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  p.drop_site = drop_site; p.drop_seed = drop_seed;
  if (LkP <= 48) return launch_tc<6>(p, smem, st);
  if (LkP <= 96) return launch_tc<12>(p, smem, st);
  return launch_tc<16>(p, smem, st);
}
