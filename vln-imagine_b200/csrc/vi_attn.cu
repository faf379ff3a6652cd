// vi_attn.cu - fused masked multi-head attention for the navigation step (scores never reach HBM).
//
//   O = softmax(Q K^T / sqrt(64) + key_padding + (w * pair_dist + b)) V        per (episode, head)
//
// Sequence lengths on this path are tiny (<= 37 queries x <= 85 keys at cfg-2, <= 212 at cfg-5;
// SURVEY.md section 5), so a whole K/V head fits one CTA's shared memory and attention is ~1.7% of the
// step's FLOPs: the bf16 kernel uses warp-level mma.sync m16n8k16 tiles (a 16-query tile per
// warp) with an fp32 online softmax rather than 128-row tcgen05 tiles that would be >70% padding.
// The fp32 kernel is the check mode.
//
// Reference: BertSelfAttention / BertOutAttention (VLN-DUET/map_nav_src/models/vilmodel.py:118-134,
// 336-349), GASA bias (:392-394, :1145-1149), nn.MultiheadAttention with key_padding_mask
// (models/transformer.py:176-177).
#include "vi_common.cuh"

namespace {

struct AttnParams {
  const void* q; long long ldq;
  const void* k; long long ldk;
  const void* v; long long ldv;
  void* o; long long ldo;
  const uint8_t* key_mask;
  const float* pair_dist;
  const float* bias_affine;
  float* lse;
  int B, H, Lq, Lk, LkP, mask_mode;
};

constexpr int DH = 64;
constexpr int KS_STRIDE = 72;      // bf16 elements per K row in smem (64 + 8: conflict-free fragment loads)

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (ceil(Lq/64), H, B), 128 threads: warp w owns query rows [64*bx + 16w, +16)
__global__ void __launch_bounds__(128) attn_fwd_bf16_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int LkP = p.LkP;
  const int vt_stride = LkP + 8;
  bf16* Ks = reinterpret_cast<bf16*>(smem);                           // [LkP][72]
  bf16* Vt = Ks + (size_t)LkP * KS_STRIDE;                            // [64][LkP+8]   (V transposed)
  float* madd = reinterpret_cast<float*>(Vt + (size_t)DH * vt_stride);  // [LkP]

  const int b = blockIdx.z, h = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16* kg = reinterpret_cast<const bf16*>(p.k) + (long long)b * p.Lk * p.ldk + h * DH;
  const bf16* vg = reinterpret_cast<const bf16*>(p.v) + (long long)b * p.Lk * p.ldv + h * DH;

  // ---- stage K (row-major, padded stride) and V (transposed) for this (episode, head) ----
  for (int e = tid; e < LkP * 8; e += 128) {
    const int key = e >> 3, ch = e & 7;
    uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
    if (key < p.Lk) {
      kv = *reinterpret_cast<const uint4*>(kg + (long long)key * p.ldk + ch * 8);
      vv = *reinterpret_cast<const uint4*>(vg + (long long)key * p.ldv + ch * 8);
    }
    *reinterpret_cast<uint4*>(Ks + (size_t)key * KS_STRIDE + ch * 8) = kv;
    const bf16* ve = reinterpret_cast<const bf16*>(&vv);
#pragma unroll
    for (int i = 0; i < 8; ++i) Vt[(size_t)(ch * 8 + i) * vt_stride + key] = ve[i];
  }
  for (int key = tid; key < LkP; key += 128) {
    float m = 0.f;
    if (key >= p.Lk) m = -INFINITY;
    else if (p.key_mask && !p.key_mask[(long long)b * p.Lk + key]) m = (p.mask_mode == VI_MASK_NEG_INF) ? -INFINITY : -10000.0f;
    madd[key] = m;
  }
  __syncthreads();

  const int q0 = blockIdx.x * 64 + warp * 16;
  if (q0 >= p.Lq) return;
  const int g = lane >> 2, tg = lane & 3;
  const int r0 = q0 + g, r1 = q0 + g + 8;

  // ---- Q fragments (16 rows x 64) straight from global ----
  uint32_t qa[4][4];
  {
    const bf16* qg = reinterpret_cast<const bf16*>(p.q) + (long long)b * p.Lq * p.ldq + h * DH;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 16 + 2 * tg;
      qa[ks][0] = r0 < p.Lq ? *reinterpret_cast<const uint32_t*>(qg + (long long)r0 * p.ldq + c) : 0u;
      qa[ks][1] = r1 < p.Lq ? *reinterpret_cast<const uint32_t*>(qg + (long long)r1 * p.ldq + c) : 0u;
      qa[ks][2] = r0 < p.Lq ? *reinterpret_cast<const uint32_t*>(qg + (long long)r0 * p.ldq + c + 8) : 0u;
      qa[ks][3] = r1 < p.Lq ? *reinterpret_cast<const uint32_t*>(qg + (long long)r1 * p.ldq + c + 8) : 0u;
    }
  }
  float bw = 0.f, bb = 0.f;
  const float* pd0 = nullptr;
  const float* pd1 = nullptr;
  if (p.pair_dist) {
    bw = p.bias_affine[0];
    bb = p.bias_affine[1];
    pd0 = p.pair_dist + ((long long)b * p.Lq + (r0 < p.Lq ? r0 : 0)) * p.Lk;
    pd1 = p.pair_dist + ((long long)b * p.Lq + (r1 < p.Lq ? r1 : 0)) * p.Lk;
  }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kc = 0; kc < LkP; kc += 64) {
    const int ntiles = min(8, (LkP - kc) >> 3);       // LkP % 16 == 0 -> even
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < ntiles) {
        const bf16* kr = Ks + (size_t)(kc + nt * 8 + g) * KS_STRIDE + 2 * tg;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + ks * 16);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 8);
          mma_16816(s[nt], qa[ks], b0, b1);
        }
      }
    }
    // scale, mask, bias; chunk row max
    float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < ntiles) {
        const int key = kc + nt * 8 + 2 * tg;
        const float ma = madd[key], mb = madd[key + 1];
        float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
        if (pd0) {
          if (key < p.Lk) { b00 = fmaf(bw, pd0[key], bb); b10 = fmaf(bw, pd1[key], bb); }
          if (key + 1 < p.Lk) { b01 = fmaf(bw, pd0[key + 1], bb); b11 = fmaf(bw, pd1[key + 1], bb); }
        }
        s[nt][0] = s[nt][0] * 0.125f + ma + b00;
        s[nt][1] = s[nt][1] * 0.125f + mb + b01;
        s[nt][2] = s[nt][2] * 0.125f + ma + b10;
        s[nt][3] = s[nt][3] * 0.125f + mb + b11;
        cm0 = fmaxf(cm0, fmaxf(s[nt][0], s[nt][1]));
        cm1 = fmaxf(cm1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float mn0 = fmaxf(m0, cm0), mn1 = fmaxf(m1, cm1);
    const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0, mu1 = (mn1 == -INFINITY) ? 0.f : mn1;
    const float sc0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - mu0);
    const float sc1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - mu1);
    m0 = mn0; m1 = mn1;
    l0 *= sc0; l1 *= sc1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { o[dt][0] *= sc0; o[dt][1] *= sc0; o[dt][2] *= sc1; o[dt][3] *= sc1; }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < ntiles) {
        s[nt][0] = __expf(s[nt][0] - mu0); s[nt][1] = __expf(s[nt][1] - mu0);
        s[nt][2] = __expf(s[nt][2] - mu1); s[nt][3] = __expf(s[nt][3] - mu1);
        l0 += s[nt][0] + s[nt][1];
        l1 += s[nt][2] + s[nt][3];
      }
    }
    // O += P V : P (accumulator layout) re-packed as the A operand, V^T rows give the B operand
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (2 * kk < ntiles) {
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const bf16* vr = Vt + (size_t)g * vt_stride + kc + kk * 16 + 2 * tg;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(vr + (size_t)dt * 8 * vt_stride);
          const uint32_t b1 = *reinterpret_cast<const uint32_t*>(vr + (size_t)dt * 8 * vt_stride + 8);
          mma_16816(o[dt], pa, b0, b1);
        }
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  bf16* og = reinterpret_cast<bf16*>(p.o) + (long long)b * p.Lq * p.ldo + h * DH;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int c = dt * 8 + 2 * tg;
    if (r0 < p.Lq) *reinterpret_cast<uint32_t*>(og + (long long)r0 * p.ldo + c) = pack_bf16x2(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < p.Lq) *reinterpret_cast<uint32_t*>(og + (long long)r1 * p.ldo + c) = pack_bf16x2(o[dt][2] * i1, o[dt][3] * i1);
  }
  if (p.lse && tg == 0) {
    float* lg = p.lse + ((long long)b * p.H + h) * p.Lq;
    if (r0 < p.Lq) lg[r0] = m0 + logf(l0);
    if (r1 < p.Lq) lg[r1] = m1 + logf(l1);
  }
}

// fp32 check mode: grid (H, B), 128 threads; warp w handles query rows w, w+4, ...
__global__ void __launch_bounds__(128) attn_fwd_f32_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* Ks = reinterpret_cast<float*>(smem);            // [Lk][65]
  float* Vs = Ks + (size_t)p.Lk * 65;                    // [Lk][64]
  float* madd = Vs + (size_t)p.Lk * 64;                  // [Lk]
  float* sc = madd + p.Lk;                               // [4][Lk]
  float* qs = sc + 4 * (size_t)p.Lk;                     // [4][64]
  const int b = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* kg = reinterpret_cast<const float*>(p.k) + (long long)b * p.Lk * p.ldk + h * DH;
  const float* vg = reinterpret_cast<const float*>(p.v) + (long long)b * p.Lk * p.ldv + h * DH;
  for (int e = tid; e < p.Lk * DH; e += 128) {
    const int key = e >> 6, d = e & 63;
    Ks[(size_t)key * 65 + d] = kg[(long long)key * p.ldk + d];
    Vs[(size_t)key * 64 + d] = vg[(long long)key * p.ldv + d];
  }
  for (int key = tid; key < p.Lk; key += 128) {
    float m = 0.f;
    if (p.key_mask && !p.key_mask[(long long)b * p.Lk + key]) m = (p.mask_mode == VI_MASK_NEG_INF) ? -INFINITY : -10000.0f;
    madd[key] = m;
  }
  __syncthreads();
  float bw = 0.f, bb = 0.f;
  if (p.pair_dist) { bw = p.bias_affine[0]; bb = p.bias_affine[1]; }
  const float* qg = reinterpret_cast<const float*>(p.q) + (long long)b * p.Lq * p.ldq + h * DH;
  float* og = reinterpret_cast<float*>(p.o) + (long long)b * p.Lq * p.ldo + h * DH;
  float* myq = qs + warp * 64;
  float* mysc = sc + (size_t)warp * p.Lk;
  for (int r = warp; r < p.Lq; r += 4) {
    myq[lane] = qg[(long long)r * p.ldq + lane];
    myq[lane + 32] = qg[(long long)r * p.ldq + lane + 32];
    __syncwarp();
    float mx = -INFINITY;
    for (int key = lane; key < p.Lk; key += 32) {
      float dot = 0.f;
#pragma unroll 16
      for (int d = 0; d < DH; ++d) dot = fmaf(myq[d], Ks[(size_t)key * 65 + d], dot);
      // same association as the reference: (qk / 8 + mask) + bias  (mask+sprels is pre-added there;
      // the difference is below fp32 resolution of the -10000 term only for masked keys)
      float s = dot * 0.125f;
      float add = madd[key];
      if (p.pair_dist) add += fmaf(bw, p.pair_dist[((long long)b * p.Lq + r) * p.Lk + key], bb);
      s += add;
      mysc[key] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    const float mu = (mx == -INFINITY) ? 0.f : mx;
    float sum = 0.f;
    for (int key = lane; key < p.Lk; key += 32) {
      const float e = expf(mysc[key] - mu);
      mysc[key] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int key = 0; key < p.Lk; ++key) {
      const float pj = mysc[key];
      a0 = fmaf(pj, Vs[(size_t)key * 64 + lane], a0);
      a1 = fmaf(pj, Vs[(size_t)key * 64 + lane + 32], a1);
    }
    const float inv = 1.0f / sum;
    og[(long long)r * p.ldo + lane] = a0 * inv;
    og[(long long)r * p.ldo + lane + 32] = a1 * inv;
    if (p.lse && lane == 0) p.lse[((long long)b * p.H + h) * p.Lq + r] = mx + logf(sum);
    __syncwarp();
  }
}

}  // namespace

extern "C" int vi_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                           int64_t ldo, int dtype, const uint8_t* key_mask, const float* pair_dist,
                           const float* bias_affine, float* lse, int B, int H, int Lq, int Lk, int mask_mode,
                           vi_stream_t stream) {
  VI_CHECK_ARG(q && k && v && o, "vi_attn_fwd: null operand");
  VI_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, "vi_attn_fwd: empty problem B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  VI_CHECK_ARG(Lk <= 512, "vi_attn_fwd: Lk=%d exceeds the single-pass limit of 512 keys", Lk);
  VI_CHECK_ARG(mask_mode == VI_MASK_ADD_NEG10000 || mask_mode == VI_MASK_NEG_INF, "vi_attn_fwd: bad mask_mode");
  VI_CHECK_ARG(!pair_dist || bias_affine, "vi_attn_fwd: pair_dist needs bias_affine {w,b}");
  VI_CHECK_ARG(ldq >= (int64_t)H * DH && ldk >= (int64_t)H * DH && ldv >= (int64_t)H * DH && ldo >= (int64_t)H * DH,
               "vi_attn_fwd: leading dimensions smaller than H*64");
  AttnParams p;
  p.q = q; p.ldq = ldq; p.k = k; p.ldk = ldk; p.v = v; p.ldv = ldv; p.o = o; p.ldo = ldo;
  p.key_mask = key_mask; p.pair_dist = pair_dist; p.bias_affine = bias_affine; p.lse = lse;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.LkP = (Lk + 15) & ~15; p.mask_mode = mask_mode;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == VI_DT_BF16) {
    VI_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "vi_attn_fwd: bf16 leading dims must be multiples of 8");
    VI_CHECK_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0 && ((uintptr_t)o & 3) == 0, "vi_attn_fwd: misaligned bf16 operands");
    const size_t smem = (size_t)p.LkP * KS_STRIDE * 2 + (size_t)DH * (p.LkP + 8) * 2 + (size_t)p.LkP * 4;
    dim3 grid((Lq + 63) / 64, H, B);
    attn_fwd_bf16_kernel<<<grid, 128, smem, st>>>(p);
  } else if (dtype == VI_DT_F32) {
    const size_t smem = ((size_t)Lk * 65 + (size_t)Lk * 64 + Lk + 4 * (size_t)Lk + 4 * 64) * 4;
    dim3 grid(H, B);
    attn_fwd_f32_kernel<<<grid, 128, smem, st>>>(p);
  } else {
    vi_set_error("vi_attn_fwd: bad dtype %d", dtype);
    return VI_ERR_ARG;
  }
  VI_LAUNCH_CHECK();
  return VI_OK;
}

int vi_attn_init() {
  VI_CUDA(cudaFuncSetAttribute(attn_fwd_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  VI_CUDA(cudaFuncSetAttribute(attn_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return VI_OK;
}
