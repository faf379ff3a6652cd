// vi_bwd.cu - backward-pass kernels of the navigation hot path (fine-tuning, BASELINE.json cfg-4).
//
// The dense gradients (dgrad / wgrad) reuse the tcgen05 GEMM of vi_gemm_tc.cu on transposed operands; this file
// holds everything around them: transposes, bias / LayerNorm parameter reductions, LayerNorm / GELU / ReLU input
// gradients, the attention backward, gather / scatter adjoints, and the adjoints of the logit heads and of the
// alignment loss.  Reductions over rows are deterministic (fixed summation order, no floating-point atomics) except
// where noted.  Reference semantics: torch.autograd of the reference modules (VLN-DUET/map_nav_src/models/
// vilmodel.py, transformer.py; VLN-HAMT/finetune_src/models/vilmodel_cmt.py).
#include "vi_common.cuh"

namespace {

constexpr int D = VI_HIDDEN;

template <typename T> __device__ __forceinline__ float ldf(const T* p, long long i) { return to_f32<T>(p[i]); }

// ---------------------------------------------------------------------------------------------
// dst[c, r] = src[r, c]  (rows x cols -> cols x ldd, columns r >= rows of dst are zero up to pad_rows)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ src, long long ld, T* __restrict__ dst,
                                                        long long ldd, int rows, int cols, int pad_rows) {
  pdl_enter();
  __shared__ T tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? src[(long long)r * ld + c] : from_f32<T>(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;
    if (c < cols && r < pad_rows) dst[(long long)c * ldd + r] = tile[tx][ty + 8 * i];
  }
}

// ---------------------------------------------------------------------------------------------
// out[c] = sum_r x[r, c] over a row range; one CTA per 32-column slab, fixed order -> deterministic
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out,
                                                     long long rows, int cols) {
  pdl_enter();
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < cols)
    for (long long r = ty; r < rows; r += 8) s += ldf(x, r * ld + c);
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    out[c] = t;
  }
}

// ---------------------------------------------------------------------------------------------
// elementwise activations on [rows, cols] views
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, int act) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = ldf(x, i);
  y[i] = from_f32<T>(act == VI_EPI_GELU ? gelu_erf(v) : fmaxf(v, 0.f));
}
// dx = dy * act'(x)
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                      long long n, int act) {
  pdl_enter();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = ldf(x, i), g = ldf(dy, i);
  dx[i] = from_f32<T>(act == VI_EPI_GELU ? g * gelu_grad(v) : (v > 0.f ? g : 0.f));
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  y = LN(a [+ b]) * gamma + beta;  dy = dy32 [+ dy16].
//   dx[r] = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma      (one warp per row)
//   stats[r] = {mean, rstd} for the parameter-gradient pass
// ---------------------------------------------------------------------------------------------
struct RowGroups {
  int n;
  int end[4];
};
__device__ __forceinline__ int group_of_row(const RowGroups& g, long long row) {
  int i = 0;
  while (i < g.n - 1 && row >= g.end[i]) ++i;
  return i;
}

__global__ void __launch_bounds__(128) ln_bwd_dx_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        const float* __restrict__ gamma, float eps,
                                                        const float* __restrict__ dy32, const bf16* __restrict__ dy16,
                                                        float* __restrict__ dx32, bf16* __restrict__ dx16,
                                                        float* __restrict__ stats, long long rows, const RowGroups grp) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  gamma += group_of_row(grp, row) * D;
  float x[24], g[24];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = (lane + 32 * j) * 4;
    float4 t = *reinterpret_cast<const float4*>(a + row * D + c);
    if (b) {
      const float4 u = *reinterpret_cast<const float4*>(b + row * D + c);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
    s += t.x + t.y + t.z + t.w;
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = x[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = (lane + 32 * j) * 4;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    if (dy32) {
      const float4 t = *reinterpret_cast<const float4*>(dy32 + row * D + c);
      d[0] = t.x; d[1] = t.y; d[2] = t.z; d[3] = t.w;
    }
    if (dy16) {
#pragma unroll
      for (int e = 0; e < 4; ++e) d[e] += __bfloat162float(dy16[row * D + c + e]);
    }
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    const float gmv[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (x[4 * j + e] - mean) * rstd;
      const float gg = d[e] * gmv[e];
      x[4 * j + e] = xh;
      g[4 * j + e] = gg;
      sg += gg;
      sgx = fmaf(gg, xh, sgx);
    }
  }
  const float mg = warp_sum(sg) * (1.0f / D), mgx = warp_sum(sgx) * (1.0f / D);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int c = (lane + 32 * j) * 4;
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = rstd * (g[4 * j + e] - mg - x[4 * j + e] * mgx);
    if (dx32) *reinterpret_cast<float4*>(dx32 + row * D + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (dx16) *reinterpret_cast<uint2*>(dx16 + row * D + c) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
  }
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dgamma[g, c] = sum_r dy * xhat, dbeta[g, c] = sum_r dy   over the rows of group g; grid (24, n_groups)
__global__ void __launch_bounds__(256) ln_bwd_param_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           const float* __restrict__ dy32, const bf16* __restrict__ dy16,
                                                           const float* __restrict__ stats, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, long long rows, const RowGroups grp) {
  pdl_enter();
  __shared__ float pg[8][33], pb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int gi = blockIdx.y;
  const long long r0 = gi == 0 ? 0 : grp.end[gi - 1];
  long long r1 = gi == grp.n - 1 ? rows : grp.end[gi];
  if (r1 > rows) r1 = rows;
  float sg = 0.f, sb = 0.f;
  for (long long r = r0 + ty; r < r1; r += 8) {
    float x = a[r * D + c];
    if (b) x += b[r * D + c];
    float d = dy32 ? dy32[r * D + c] : 0.f;
    if (dy16) d += __bfloat162float(dy16[r * D + c]);
    const float xh = (x - stats[2 * r]) * stats[2 * r + 1];
    sg = fmaf(d, xh, sg);
    sb += d;
  }
  pg[ty][tx] = sg;
  pb[ty][tx] = sb;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t += pg[i][tx]; u += pb[i][tx]; }
    dgamma[gi * D + c] = t;
    dbeta[gi * D + c] = u;
  }
}

// ---------------------------------------------------------------------------------------------
// small-feature linear  t = feat W^T + b  (feat_dim <= 16):  dW[c, k] = sum_r dt[r, c] feat[r, k], db[c] = sum_r dt[r, c]
// one CTA per 32-column slab, fixed order
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) feat_wgrad_kernel(const float* __restrict__ dt, const float* __restrict__ feat, int fd,
                                                         float* __restrict__ dW, float* __restrict__ db, long long rows) {
  pdl_enter();
  __shared__ float part[8][17][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) acc[k] = 0.f;
  for (long long r = ty; r < rows; r += 8) {
    const float d = dt[r * D + c];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < fd) acc[k] = fmaf(d, __ldg(feat + r * fd + k), acc[k]);
    acc[16] += d;
  }
#pragma unroll
  for (int k = 0; k < 17; ++k) part[ty][k][tx] = acc[k];
  __syncthreads();
  if (ty == 0) {
    for (int k = 0; k < 17; ++k) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += part[i][k][tx];
      if (k < fd) dW[(long long)c * fd + k] = t;
      else if (k == 16 && db) db[c] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dst[idx[r]] += src[r]   (embedding-table adjoint; fp32 atomics: rows may repeat)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) scatter_add_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                               float* __restrict__ dst, long long rows, int period) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long d = idx ? idx[row] : (row % period);
  for (int c = lane; c < D; c += 32) atomicAdd(dst + d * D + c, src[row * D + c]);
}

// ---------------------------------------------------------------------------------------------
// out[r] = x[r] . w + b  adjoint:  dx[r, :] = dout[r] * w[g],  dw[g, c] = sum_r dout[r] x[r, c],  db[g] = sum_r dout[r]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rowdot_bwd_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                            float* __restrict__ dx, long long rows, const RowGroups grp) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* wg = w + group_of_row(grp, row) * D;
  const float g = dout[row];
  for (int c = lane; c < D; c += 32) dx[row * D + c] = g * wg[c];
}
__global__ void __launch_bounds__(256) rowdot_bwd_w_kernel(const float* __restrict__ dout, const float* __restrict__ x,
                                                           float* __restrict__ dw, float* __restrict__ db, long long rows,
                                                           const RowGroups grp) {
  pdl_enter();
  __shared__ float pw[8][33], pb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int gi = blockIdx.y;
  const long long r0 = gi == 0 ? 0 : grp.end[gi - 1];
  long long r1 = gi == grp.n - 1 ? rows : grp.end[gi];
  if (r1 > rows) r1 = rows;
  float sw = 0.f, sb = 0.f;
  for (long long r = r0 + ty; r < r1; r += 8) {
    const float g = dout[r];
    sw = fmaf(g, x[r * D + c], sw);
    sb += g;
  }
  pw[ty][tx] = sw;
  pb[ty][tx] = sb;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f, u = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t += pw[i][tx]; u += pb[i][tx]; }
    dw[gi * D + c] = t;
    if (blockIdx.x == 0 && tx == 0 && db) db[gi] = u;
  }
}

// ---------------------------------------------------------------------------------------------
// attention backward, one CTA (8 warps) per (episode, head); fp32 arithmetic on fp32 or bf16 tensors.
//   S = Q K^T / 8 + mask + (w dist + b);  P = softmax(S);  dP = dO V^T;  Dr = rowsum(dO o O);
//   dS = P o (dP - Dr);  dQ = dS K / 8;  dK = dS^T Q / 8;  dV = P^T dO;
//   GASA:  dw += sum dS dist,  db += sum dS  (two fp32 atomics per CTA).
// K, V and the dK / dV accumulators live in shared memory; each warp walks query rows; the accumulators take
// shared-memory atomics (rows of different warps collide), so dK / dV sums are order-dependent in the last bits.
// ---------------------------------------------------------------------------------------------
struct AttnBwdParams {
  const void* q; long long ldq;
  const void* k; long long ldk;
  const void* v; long long ldv;
  const void* dout; long long ldo;
  void* dq; long long lddq;
  void* dk; long long lddk;
  void* dv; long long lddv;
  const uint8_t* key_mask;
  const float* pair_dist;
  const float* bias_affine;
  float* d_affine;
  int B, H, Lq, Lk, mask_mode;
};

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const AttnBwdParams p) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  const int Lk = p.Lk;
  float* Ks = sm;                        // [Lk][65]
  float* Vs = Ks + (size_t)Lk * 65;      // [Lk][65]
  float* dKs = Vs + (size_t)Lk * 65;     // [Lk][64]
  float* dVs = dKs + (size_t)Lk * 64;    // [Lk][64]
  float* madd = dVs + (size_t)Lk * 64;   // [Lk]
  float* wbuf = madd + Lk;               // per warp: q[64], do[64], prob[Lk], ds[Lk]
  const int b = blockIdx.y, h = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* kg = reinterpret_cast<const T*>(p.k) + (long long)b * Lk * p.ldk + h * 64;
  const T* vg = reinterpret_cast<const T*>(p.v) + (long long)b * Lk * p.ldv + h * 64;
  for (int e = tid; e < Lk * 64; e += 256) {
    const int key = e >> 6, d = e & 63;
    Ks[(size_t)key * 65 + d] = ldf(kg, (long long)key * p.ldk + d);
    Vs[(size_t)key * 65 + d] = ldf(vg, (long long)key * p.ldv + d);
    dKs[e] = 0.f;
    dVs[e] = 0.f;
  }
  for (int key = tid; key < Lk; key += 256) {
    float m = 0.f;
    if (p.key_mask && !p.key_mask[(long long)b * Lk + key]) m = (p.mask_mode == VI_MASK_NEG_INF) ? -INFINITY : -10000.0f;
    madd[key] = m;
  }
  __syncthreads();
  float bw = 0.f, bb = 0.f;
  if (p.pair_dist) { bw = p.bias_affine[0]; bb = p.bias_affine[1]; }
  float* myq = wbuf + (size_t)warp * (128 + 2 * Lk);
  float* mydo = myq + 64;
  float* prob = mydo + 64;
  float* ds = prob + Lk;
  const T* qg = reinterpret_cast<const T*>(p.q) + (long long)b * p.Lq * p.ldq + h * 64;
  const T* dog = reinterpret_cast<const T*>(p.dout) + (long long)b * p.Lq * p.ldo + h * 64;
  T* dqg = reinterpret_cast<T*>(p.dq) + (long long)b * p.Lq * p.lddq + h * 64;
  float aw = 0.f, ab = 0.f;              // GASA affine gradients of this warp
  for (int r = warp; r < p.Lq; r += 8) {
    myq[lane] = ldf(qg, (long long)r * p.ldq + lane);
    myq[lane + 32] = ldf(qg, (long long)r * p.ldq + lane + 32);
    mydo[lane] = ldf(dog, (long long)r * p.ldo + lane);
    mydo[lane + 32] = ldf(dog, (long long)r * p.ldo + lane + 32);
    __syncwarp();
    const float* pd = p.pair_dist ? p.pair_dist + ((long long)b * p.Lq + r) * Lk : nullptr;
    float mx = -INFINITY;
    for (int key = lane; key < Lk; key += 32) {
      float dot = 0.f, dp = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) {
        dot = fmaf(myq[d], Ks[(size_t)key * 65 + d], dot);
        dp = fmaf(mydo[d], Vs[(size_t)key * 65 + d], dp);
      }
      float add = madd[key];
      if (pd) add += fmaf(bw, pd[key], bb);
      const float s = dot * 0.125f + add;
      prob[key] = s;
      ds[key] = dp;                      // dP for now
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    const float mu = (mx == -INFINITY) ? 0.f : mx;
    float sum = 0.f;
    for (int key = lane; key < Lk; key += 32) {
      const float e = expf(prob[key] - mu);
      prob[key] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    // Dr = sum_j P_j dP_j  (= rowsum(dO o O))
    float dr = 0.f;
    for (int key = lane; key < Lk; key += 32) {
      const float pj = prob[key] * inv;
      prob[key] = pj;
      dr = fmaf(pj, ds[key], dr);
    }
    dr = warp_sum(dr);
    for (int key = lane; key < Lk; key += 32) {
      const float g = prob[key] * (ds[key] - dr);
      ds[key] = g;
      if (pd) { aw = fmaf(g, pd[key], aw); ab += g; }
    }
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    const float qa = myq[lane] * 0.125f, qb = myq[lane + 32] * 0.125f;
    const float oa = mydo[lane], ob = mydo[lane + 32];
    for (int key = 0; key < Lk; ++key) {
      const float g = ds[key], pj = prob[key];
      q0 = fmaf(g, Ks[(size_t)key * 65 + lane], q0);
      q1 = fmaf(g, Ks[(size_t)key * 65 + lane + 32], q1);
      atomicAdd(dKs + (size_t)key * 64 + lane, g * qa);
      atomicAdd(dKs + (size_t)key * 64 + lane + 32, g * qb);
      atomicAdd(dVs + (size_t)key * 64 + lane, pj * oa);
      atomicAdd(dVs + (size_t)key * 64 + lane + 32, pj * ob);
    }
    dqg[(long long)r * p.lddq + lane] = from_f32<T>(q0 * 0.125f);
    dqg[(long long)r * p.lddq + lane + 32] = from_f32<T>(q1 * 0.125f);
    __syncwarp();
  }
  if (p.pair_dist && p.d_affine) {
    aw = warp_sum(aw);
    ab = warp_sum(ab);
    if (lane == 0) { atomicAdd(p.d_affine, aw); atomicAdd(p.d_affine + 1, ab); }
  }
  __syncthreads();
  T* dkg = reinterpret_cast<T*>(p.dk) + (long long)b * Lk * p.lddk + h * 64;
  T* dvg = reinterpret_cast<T*>(p.dv) + (long long)b * Lk * p.lddv + h * 64;
  for (int e = tid; e < Lk * 64; e += 256) {
    const int key = e >> 6, d = e & 63;
    dkg[(long long)key * p.lddk + d] = from_f32<T>(dKs[e]);
    dvg[(long long)key * p.lddv + d] = from_f32<T>(dVs[e]);
  }
}

// ---------------------------------------------------------------------------------------------
// adjoint of vi_duet_fuse_logits (one warp per episode, same id matching as the forward kernel)
// ---------------------------------------------------------------------------------------------
constexpr int FUSE_MAX = 512;
__global__ void __launch_bounds__(32) duet_fuse_logits_bwd_kernel(
    const float* __restrict__ g_raw, const float* __restrict__ l_raw, const float* __restrict__ fuse_raw,
    const uint8_t* __restrict__ gmap_masks, const uint8_t* __restrict__ gmap_visited, const uint8_t* __restrict__ vp_nav,
    const int32_t* __restrict__ gmap_ids, const int32_t* __restrict__ cand_ids, const float* __restrict__ d_global,
    const float* __restrict__ d_local, const float* __restrict__ d_fused, float* __restrict__ dg_raw,
    float* __restrict__ dl_raw, float* __restrict__ dfuse_raw, int G, int P) {
  pdl_enter();
  __shared__ float dll[FUSE_MAX];          // gradient w.r.t. the masked local logits
  __shared__ int gid[FUSE_MAX], cid[FUSE_MAX];
  __shared__ uint8_t gvis[FUSE_MAX], cvis[FUSE_MAX];
  __shared__ float dbw_s;
  const int b = blockIdx.x, lane = threadIdx.x;
  const float fw = fuse_raw ? 1.0f / (1.0f + expf(-fuse_raw[b])) : 0.5f;
  for (int v = lane; v < P; v += 32) {
    cid[v] = cand_ids[(long long)b * P + v];
    dll[v] = d_local ? d_local[(long long)b * P + v] : 0.f;
  }
  for (int j = lane; j < G; j += 32) gid[j] = gmap_ids[(long long)b * G + j];
  __syncwarp();
  for (int j = lane; j < G; j += 32) {
    bool in_set = false;
    if (gid[j] != -1)
      for (int k = 0; k < G; ++k) in_set |= (gid[k] == gid[j]) && gid[k] != -1 && gmap_visited[(long long)b * G + k];
    gvis[j] = in_set;
  }
  for (int v = lane; v < P; v += 32) {
    bool in_set = false;
    if (cid[v] != -2)
      for (int k = 0; k < G; ++k) in_set |= (gid[k] == cid[v]) && gid[k] != -1 && gmap_visited[(long long)b * G + k];
    cvis[v] = in_set;
  }
  __syncwarp();
  // fused[j] = global[j] + (j == 0 ? local[0] : hit >= 0 ? local[hit] : bw) -> scatter d_fused into dll / dbw (lane 0, in order)
  float dfw = 0.f;                          // d loss / d fuse weight
  if (lane == 0) {
    float dbw = 0.f;
    for (int j = 0; j < G; ++j) {
      const float df = d_fused ? d_fused[(long long)b * G + j] : 0.f;
      if (j == 0) { dll[0] += df; continue; }
      if (gid[j] == -1 || gvis[j]) continue;
      int hit = -1;
      for (int v = 1; v < P; ++v)
        if (cid[v] != -2 && !cvis[v] && cid[v] == gid[j]) hit = v;
      if (hit >= 0) dll[hit] += df; else dbw += df;
    }
    for (int v = 1; v < P; ++v)
      if (cid[v] != -2 && cvis[v]) dll[v] += dbw;
    dbw_s = dbw;
  }
  __syncwarp();
  for (int v = lane; v < P; v += 32) {
    const long long i = (long long)b * P + v;
    float g = 0.f;
    if (vp_nav[i]) {                        // masked positions are constants (-inf)
      g = dll[v] * (1.0f - fw);
      dfw -= dll[v] * l_raw[i];
    }
    dl_raw[i] = g;
  }
  for (int j = lane; j < G; j += 32) {
    const long long i = (long long)b * G + j;
    float g = 0.f;
    if (!(gmap_visited[i] || !gmap_masks[i])) {
      const float dgl = (d_global ? d_global[i] : 0.f) + (d_fused ? d_fused[i] : 0.f);
      g = dgl * fw;
      dfw += dgl * g_raw[i];
    }
    dg_raw[i] = g;
  }
  dfw = warp_sum(dfw);
  if (lane == 0 && dfuse_raw) dfuse_raw[b] = dfw * fw * (1.0f - fw);
}

// ---------------------------------------------------------------------------------------------
// cosine alignment loss adjoint:  loss = mean_r (1 - cos(p_r, t_r));  dp_r = -(dloss / R) d cos / d p_r
// (torch clamps each norm at eps = 1e-8; the clamp is inactive for non-degenerate rows and is treated as such)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cosine_loss_bwd_kernel(const float* __restrict__ proj, const float* __restrict__ tgt,
                                                              const float* __restrict__ dloss, float* __restrict__ dproj,
                                                              float* __restrict__ dtgt, long long rows) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float a[24], t[24];
  float ab = 0.f, aa = 0.f, bb = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    a[i] = proj[row * D + lane + 32 * i];
    t[i] = tgt[row * D + lane + 32 * i];
    ab = fmaf(a[i], t[i], ab); aa = fmaf(a[i], a[i], aa); bb = fmaf(t[i], t[i], bb);
  }
  ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
  const float na = fmaxf(sqrtf(aa), 1e-8f), nb = fmaxf(sqrtf(bb), 1e-8f);
  const float cosv = ab / (na * nb);
  const float scale = -dloss[0] / (float)rows;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    if (dproj) dproj[row * D + lane + 32 * i] = scale * (t[i] / (na * nb) - cosv * a[i] / (na * na));
    if (dtgt) dtgt[row * D + lane + 32 * i] = scale * (a[i] / (na * nb) - cosv * t[i] / (nb * nb));
  }
}

inline bool make_groups(RowGroups& g, int n_groups, const int32_t* ends) {
  if (n_groups < 1 || n_groups > 4 || (n_groups > 1 && !ends)) return false;
  g.n = n_groups;
  for (int i = 0; i < 4; ++i) g.end[i] = 0x7fffffff;
  for (int i = 0; i < n_groups && n_groups > 1; ++i) {
    if (ends[i] <= 0 || (i > 0 && ends[i] <= ends[i - 1])) return false;
    g.end[i] = ends[i];
  }
  return true;
}
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int vi_transpose(const void* src, int64_t ld, void* dst, int64_t ldd, int rows, int cols, int pad_rows, int dtype,
                            vi_stream_t stream) {
  VI_CHECK_ARG(src && dst && rows > 0 && cols > 0 && pad_rows >= rows && ld >= cols && ldd >= pad_rows, "vi_transpose: bad operands");
  dim3 grid((cols + 31) / 32, (pad_rows + 31) / 32);
  if (dtype == VI_DT_BF16)
    VI_CUDA(vi_launch(transpose_kernel<bf16>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(src), (long long)ld,
                      reinterpret_cast<bf16*>(dst), (long long)ldd, rows, cols, pad_rows));
  else
    VI_CUDA(vi_launch(transpose_kernel<float>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const float*>(src), (long long)ld,
                      reinterpret_cast<float*>(dst), (long long)ldd, rows, cols, pad_rows));
  return VI_OK;
}

extern "C" int vi_colsum(const void* x, int64_t ld, int dtype, float* out, int64_t rows, int cols, vi_stream_t stream) {
  VI_CHECK_ARG(x && out && cols > 0 && ld >= cols, "vi_colsum: bad operands");
  dim3 grid((cols + 31) / 32);
  if (dtype == VI_DT_BF16)
    VI_CUDA(vi_launch(colsum_kernel<bf16>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(x), (long long)ld, out,
                      (long long)rows, cols));
  else
    VI_CUDA(vi_launch(colsum_kernel<float>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const float*>(x), (long long)ld, out,
                      (long long)rows, cols));
  return VI_OK;
}

extern "C" int vi_act_fwd(const void* x, void* y, int64_t n, int act, int dtype, vi_stream_t stream) {
  VI_CHECK_ARG(x && y && (act == VI_EPI_GELU || act == VI_EPI_RELU), "vi_act_fwd: bad operands");
  if (n <= 0) return VI_OK;
  dim3 grid((unsigned)((n + 255) / 256));
  if (dtype == VI_DT_BF16)
    VI_CUDA(vi_launch(act_fwd_kernel<bf16>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(x),
                      reinterpret_cast<bf16*>(y), (long long)n, act));
  else
    VI_CUDA(vi_launch(act_fwd_kernel<float>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const float*>(x),
                      reinterpret_cast<float*>(y), (long long)n, act));
  return VI_OK;
}

extern "C" int vi_act_bwd(const void* x, const void* dy, void* dx, int64_t n, int act, int dtype, vi_stream_t stream) {
  VI_CHECK_ARG(x && dy && dx && (act == VI_EPI_GELU || act == VI_EPI_RELU), "vi_act_bwd: bad operands");
  if (n <= 0) return VI_OK;
  dim3 grid((unsigned)((n + 255) / 256));
  if (dtype == VI_DT_BF16)
    VI_CUDA(vi_launch(act_bwd_kernel<bf16>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const bf16*>(x),
                      reinterpret_cast<const bf16*>(dy), reinterpret_cast<bf16*>(dx), (long long)n, act));
  else
    VI_CUDA(vi_launch(act_bwd_kernel<float>, grid, dim3(256), 0, ST(stream), reinterpret_cast<const float*>(x),
                      reinterpret_cast<const float*>(dy), reinterpret_cast<float*>(dx), (long long)n, act));
  return VI_OK;
}

extern "C" int vi_add_ln_bwd(const float* a, const float* b, const float* gamma, float eps, const float* dy32, const void* dy16,
                             float* dx32, void* dx16, float* dgamma, float* dbeta, float* stats, int64_t rows, int n_groups,
                             const int32_t* group_row_end, vi_stream_t stream) {
  RowGroups grp;
  VI_CHECK_ARG(make_groups(grp, n_groups, group_row_end), "vi_add_ln_bwd: bad row groups");
  VI_CHECK_ARG(a && gamma && (dy32 || dy16) && (dx32 || dx16) && stats, "vi_add_ln_bwd: null operand");
  VI_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(gamma) && aligned16(dy32) && aligned16(dx32) &&
                   ((uintptr_t)dx16 & 7) == 0, "vi_add_ln_bwd: misaligned operands");
  if (rows <= 0) return VI_OK;
  VI_CUDA(vi_launch(ln_bwd_dx_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, ST(stream), a, b, gamma, eps, dy32,
                    reinterpret_cast<const bf16*>(dy16), dx32, reinterpret_cast<bf16*>(dx16), stats, (long long)rows, grp));
  if (dgamma && dbeta)
    VI_CUDA(vi_launch(ln_bwd_param_kernel, dim3(D / 32, n_groups), dim3(256), 0, ST(stream), a, b, dy32,
                      reinterpret_cast<const bf16*>(dy16), (const float*)stats, dgamma, dbeta, (long long)rows, grp));
  return VI_OK;
}

extern "C" int vi_feat_wgrad(const float* dt, const float* feat, int feat_dim, float* dW, float* db, int64_t rows,
                             vi_stream_t stream) {
  VI_CHECK_ARG(dt && feat && dW && feat_dim > 0 && feat_dim <= 16, "vi_feat_wgrad: bad operands");
  VI_CUDA(vi_launch(feat_wgrad_kernel, dim3(D / 32), dim3(256), 0, ST(stream), dt, feat, feat_dim, dW, db, (long long)rows));
  return VI_OK;
}

extern "C" int vi_scatter_add_rows(const float* src, const int64_t* idx, int period, float* dst, int64_t rows, vi_stream_t stream) {
  VI_CHECK_ARG(src && dst && (idx || period > 0), "vi_scatter_add_rows: bad operands");
  if (rows <= 0) return VI_OK;
  VI_CUDA(vi_launch(scatter_add_rows_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, ST(stream), src, idx, dst,
                    (long long)rows, period));
  return VI_OK;
}

extern "C" int vi_rowdot_bwd(const float* dout, const float* x, const float* w, float* dx, float* dw, float* db, int64_t rows,
                             int n_groups, const int32_t* group_row_end, vi_stream_t stream) {
  RowGroups grp;
  VI_CHECK_ARG(make_groups(grp, n_groups, group_row_end), "vi_rowdot_bwd: bad row groups");
  VI_CHECK_ARG(dout && x && w && dx && dw, "vi_rowdot_bwd: null operand");
  if (rows <= 0) return VI_OK;
  VI_CUDA(vi_launch(rowdot_bwd_dx_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, ST(stream), dout, w, dx, (long long)rows, grp));
  VI_CUDA(vi_launch(rowdot_bwd_w_kernel, dim3(D / 32, n_groups), dim3(256), 0, ST(stream), dout, x, dw, db, (long long)rows, grp));
  return VI_OK;
}

extern "C" int vi_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* dout,
                           int64_t ldo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int dtype,
                           const uint8_t* key_mask, const float* pair_dist, const float* bias_affine, float* d_affine, int B,
                           int H, int Lq, int Lk, int mask_mode, vi_stream_t stream) {
  VI_CHECK_ARG(q && k && v && dout && dq && dk && dv, "vi_attn_bwd: null operand");
  VI_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, "vi_attn_bwd: bad sizes B=%d H=%d Lq=%d Lk=%d", B, H, Lq, Lk);
  VI_CHECK_ARG(((size_t)Lk * (65 * 2 + 64 * 2 + 1) + 8 * (size_t)(128 + 2 * Lk)) * sizeof(float) <= 220 * 1024,
               "vi_attn_bwd: Lk=%d keys do not fit the 220 KB shared-memory tile", Lk);
  VI_CHECK_ARG(!pair_dist || bias_affine, "vi_attn_bwd: pair_dist needs bias_affine");
  AttnBwdParams p;
  p.q = q; p.ldq = ldq; p.k = k; p.ldk = ldk; p.v = v; p.ldv = ldv; p.dout = dout; p.ldo = ldo;
  p.dq = dq; p.lddq = lddq; p.dk = dk; p.lddk = lddk; p.dv = dv; p.lddv = lddv;
  p.key_mask = key_mask; p.pair_dist = pair_dist; p.bias_affine = bias_affine; p.d_affine = d_affine;
  p.B = B; p.H = H; p.Lq = Lq; p.Lk = Lk; p.mask_mode = mask_mode;
  const size_t smem = ((size_t)Lk * (65 * 2 + 64 * 2 + 1) + 8 * (size_t)(128 + 2 * Lk)) * sizeof(float);
  if (dtype == VI_DT_BF16) {
    static bool set = false;
    if (!set) { VI_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); set = true; }
    VI_CUDA(vi_launch(attn_bwd_kernel<bf16>, dim3(H, B), dim3(256), smem, ST(stream), p));
  } else {
    static bool set = false;
    if (!set) { VI_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); set = true; }
    VI_CUDA(vi_launch(attn_bwd_kernel<float>, dim3(H, B), dim3(256), smem, ST(stream), p));
  }
  return VI_OK;
}

extern "C" int vi_duet_fuse_logits_bwd(const float* g_raw, const float* l_raw, const float* fuse_raw, const uint8_t* gmap_masks,
                                       const uint8_t* gmap_visited, const uint8_t* vp_nav_masks, const int32_t* gmap_ids,
                                       const int32_t* cand_ids, const float* d_global, const float* d_local,
                                       const float* d_fused, float* dg_raw, float* dl_raw, float* dfuse_raw, int B, int G, int P,
                                       vi_stream_t stream) {
  VI_CHECK_ARG(g_raw && l_raw && gmap_masks && gmap_visited && vp_nav_masks && gmap_ids && cand_ids && dg_raw && dl_raw,
               "vi_duet_fuse_logits_bwd: null operand");
  VI_CHECK_ARG(B > 0 && G > 0 && P > 0 && G <= FUSE_MAX && P <= FUSE_MAX, "vi_duet_fuse_logits_bwd: bad sizes");
  VI_CUDA(vi_launch(duet_fuse_logits_bwd_kernel, dim3(B), dim3(32), 0, ST(stream), g_raw, l_raw, fuse_raw, gmap_masks,
                    gmap_visited, vp_nav_masks, gmap_ids, cand_ids, d_global, d_local, d_fused, dg_raw, dl_raw, dfuse_raw, G, P));
  return VI_OK;
}

extern "C" int vi_cosine_loss_bwd(const float* proj, const float* tgt, const float* dloss, float* dproj, float* dtgt, int R,
                                  vi_stream_t stream) {
  VI_CHECK_ARG(proj && tgt && dloss && (dproj || dtgt), "vi_cosine_loss_bwd: null operand");
  if (R <= 0) return VI_OK;
  VI_CUDA(vi_launch(cosine_loss_bwd_kernel, dim3((unsigned)((R + 3) / 4)), dim3(128), 0, ST(stream), proj, tgt, dloss, dproj, dtgt,
                    (long long)R));
  return VI_OK;
}
