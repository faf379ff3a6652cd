// vi_graph.cu - the per-step graph glue of a DUET rollout on the device.
//
// Between `panorama` and `navigation` the reference agent keeps, per episode, a Python GraphMap: dict-of-dict
// shortest-path distances relaxed through every visited viewpoint, a running mean embedding per node, and at every
// step an O(G^2) Python double loop that rebuilds the pair-distance matrix plus per-node trigonometric position
// features (VLN-DUET/map_nav_src/models/graph_utils.py:42-148, r2r/agent.py:98-207, 466-479).  Here the B graph maps
// of a batch are dense arrays in HBM (fp64 graph state: the reference computes in Python floats and only the final
// features are cast to fp32, so distances / path lengths / features come out bit-identical) and a step is three small
// launches with one CTA per episode:
//   vi_graph_update      add_edge for the current viewpoint's candidates + FloydGraph.update(k) as a whole-matrix relax
//   vi_graph_embed_step  masked panorama mean -> node running sums -> gathered gmap_img_embeds and [stop | pano] rows
//   vi_graph_features    7-d gmap / 14-d vp position features and the raw-metre pair distances
// HBM-bound, a few hundred KB per step: the point is not bandwidth but removing ~B*G^2 Python iterations and B
// host->device copies from the step (latency-bound kernels; the grid is one CTA per episode).
#include "vi_common.cuh"

namespace {

constexpr double UNREACHED = 95959595.0;     // graph_utils.py:44
constexpr int MAX_HOPS_STACK = 160;

// No FMA contraction anywhere in the fp64 geometry: Python evaluates dx**2 + dy**2 + dz**2 with separately rounded
// products, and the distances must match it to the last bit.
__device__ __forceinline__ double sq3(double dx, double dy, double dz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void __launch_bounds__(256) graph_init_kernel(double* dis, int32_t* point, uint8_t* visited, float* ecnt,
                                                         long long n_pairs, long long n_nodes) {
  pdl_enter();
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < n_pairs; i += stride) { dis[i] = UNREACHED; point[i] = -1; }
  for (long long i = i0; i < n_nodes; i += stride) { visited[i] = 0; ecnt[i] = 0.f; }
}

// graph_utils.py:109-115: positions of the current viewpoint and its candidates, add_edge (:53-58), update(k) (:60-70)
__global__ void __launch_bounds__(256) graph_update_kernel(double* pos, double* dis, int32_t* point, uint8_t* visited, int N,
                                                           const int32_t* cur_node, const double* cur_pos,
                                                           const int32_t* cand_node, const double* cand_pos, int C,
                                                           const int32_t* n_nodes) {
  pdl_enter();
  const int b = blockIdx.x;
  const int k = cur_node[b];
  if (k < 0) return;                                   // ended episode: its graph is frozen (agent.py:601-603)
  const int n = n_nodes[b];
  double* P = pos + (long long)b * N * 3;
  double* D = dis + (long long)b * N * N;
  int32_t* T = point + (long long)b * N * N;
  const double ax = cur_pos[b * 3 + 0], ay = cur_pos[b * 3 + 1], az = cur_pos[b * 3 + 2];
  if (threadIdx.x == 0) { P[k * 3 + 0] = ax; P[k * 3 + 1] = ay; P[k * 3 + 2] = az; }
  for (int j = threadIdx.x; j < C; j += blockDim.x) {  // candidates of one panorama are distinct viewpoints
    const int c = cand_node[(long long)b * C + j];
    if (c < 0) continue;
    const double bx = cand_pos[((long long)b * C + j) * 3 + 0], by = cand_pos[((long long)b * C + j) * 3 + 1],
                 bz = cand_pos[((long long)b * C + j) * 3 + 2];
    P[c * 3 + 0] = bx; P[c * 3 + 1] = by; P[c * 3 + 2] = bz;
    const double d = sqrt(sq3(bx - ax, by - ay, bz - az));
    if (d < D[(long long)k * N + c]) {
      D[(long long)k * N + c] = d; D[(long long)c * N + k] = d;
      T[(long long)k * N + c] = -1; T[(long long)c * N + k] = -1;
    }
  }
  __syncthreads();
  // Row / column k cannot improve during the pass (D[k][k] stays UNREACHED), so every pair relaxes independently and
  // the result equals the reference's in-place double loop (oracle/graph_oracle.py::relax, pinned to the real class).
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int x = e / n, y = e - x * n;
    if (x == y) continue;
    const double nd = __dadd_rn(D[(long long)x * N + k], D[(long long)k * N + y]);
    if (nd < D[(long long)x * N + y]) {
      D[(long long)x * N + y] = nd;
      T[(long long)x * N + y] = k;
    }
  }
  if (threadIdx.x == 0) visited[(long long)b * N + k] = 1;
}

// agent.py:466-479 (masked mean, GraphMap.update_node_embed), :125-129 (gather of the node means behind a zero [stop]
// row, zero padding), :176-178 ([stop] row in front of the panorama).  One CTA per episode, a thread owns 4 columns.
__global__ void __launch_bounds__(256) graph_embed_kernel(const float* pano, const uint8_t* pano_masks, int V, int H,
                                                          const int32_t* cur_node, const int32_t* cand_node, int C,
                                                          const uint8_t* visited, float* esum, float* ecnt, int N,
                                                          const int32_t* gmap_node, int G, float* gmap_out, float* vp_out) {
  pdl_enter();
  const int b = blockIdx.x;
  const float* Pb = pano + (long long)b * V * H;
  float* S = esum + (long long)b * N * H;
  float* Cn = ecnt + (long long)b * N;
  const int cur = cur_node[b];
  __shared__ float s_cnt;
  if (cur >= 0) {
    float nvalid = 0.f;
    for (int v = 0; v < V; ++v) nvalid += pano_masks[(long long)b * V + v] ? 1.f : 0.f;
    for (int c4 = threadIdx.x * 4; c4 < H; c4 += blockDim.x * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int v = 0; v < V; ++v) {
        if (!pano_masks[(long long)b * V + v]) continue;
        const float4 x = *reinterpret_cast<const float4*>(Pb + (long long)v * H + c4);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
      *reinterpret_cast<float4*>(S + (long long)cur * H + c4) = make_float4(acc.x / nvalid, acc.y / nvalid, acc.z / nvalid, acc.w / nvalid);
    }
    __syncthreads();
    if (threadIdx.x == 0) Cn[cur] = 1.f;               // rewrite=True (graph_utils.py:118-119)
    for (int j = 0; j < C; ++j) {
      const int c = cand_node[(long long)b * C + j];
      if (c < 0 || visited[(long long)b * N + c]) continue;      // block-uniform
      __syncthreads();
      if (threadIdx.x == 0) s_cnt = Cn[c];
      __syncthreads();
      const float cnt = s_cnt;
      for (int c4 = threadIdx.x * 4; c4 < H; c4 += blockDim.x * 4) {
        const float4 x = *reinterpret_cast<const float4*>(Pb + (long long)j * H + c4);
        float4* dst = reinterpret_cast<float4*>(S + (long long)c * H + c4);
        if (cnt == 0.f) { *dst = x; }
        else { float4 a = *dst; a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w; *dst = a; }
      }
      if (threadIdx.x == 0) Cn[c] = cnt + 1.f;
    }
    __syncthreads();
  }
  if (gmap_out) {
    for (int g = 0; g < G; ++g) {
      const int n = gmap_node[(long long)b * G + g];
      const float cnt = n >= 0 ? Cn[n] : 0.f;
      for (int c4 = threadIdx.x * 4; c4 < H; c4 += blockDim.x * 4) {
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n >= 0 && cnt > 0.f) {
          const float4 a = *reinterpret_cast<const float4*>(S + (long long)n * H + c4);
          y = make_float4(a.x / cnt, a.y / cnt, a.z / cnt, a.w / cnt);
        }
        *reinterpret_cast<float4*>(gmap_out + ((long long)b * G + g) * H + c4) = y;
      }
    }
  }
  if (vp_out) {
    float* O = vp_out + (long long)b * (V + 1) * H;
    for (int c4 = threadIdx.x * 4; c4 < H; c4 += blockDim.x * 4) *reinterpret_cast<float4*>(O + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long e = threadIdx.x * 4; e < (long long)V * H; e += blockDim.x * 4)
      *reinterpret_cast<float4*>(O + H + e) = *reinterpret_cast<const float4*>(Pb + e);
  }
}

// len(FloydGraph.path(x, y)) (graph_utils.py:74-90) without recursion
__device__ int path_hops(const int32_t* T, int N, int x, int y) {
  if (x == y) return 0;
  uint32_t stack[MAX_HOPS_STACK];
  int sp = 0, hops = 0;
  stack[sp++] = ((uint32_t)x << 16) | (uint32_t)y;
  while (sp > 0) {
    const uint32_t e = stack[--sp];
    const int a = (int)(e >> 16), c = (int)(e & 0xFFFFu);
    if (a == c) continue;
    const int k = T[(long long)a * N + c];
    if (k < 0) { ++hops; continue; }
    if (sp + 2 > MAX_HOPS_STACK) return -1;            // cannot happen for N <= 128 (a path visits a node once)
    stack[sp++] = ((uint32_t)k << 16) | (uint32_t)c;
    stack[sp++] = ((uint32_t)a << 16) | (uint32_t)k;
  }
  return hops;
}

// graph_utils.py:14-40 + 131-148: (sin h, cos h, sin e, cos e, line / 30, shortest / 30, hops / 10)
__device__ void rel_pos_fts(const double* P, const double* D, const int32_t* T, int N, int cur, int v, double heading,
                            double elevation, float* out) {
  if (v < 0) {                                         // the [stop] slot (None): zero angles, zero distances
    out[0] = 0.f; out[1] = 1.f; out[2] = 0.f; out[3] = 1.f; out[4] = 0.f; out[5] = 0.f; out[6] = 0.f;
    return;
  }
  const double dx = P[v * 3 + 0] - P[cur * 3 + 0], dy = P[v * 3 + 1] - P[cur * 3 + 1], dz = P[v * 3 + 2] - P[cur * 3 + 2];
  const double xy = fmax(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))), 1e-8);
  const double xyz = fmax(sqrt(sq3(dx, dy, dz)), 1e-8);
  double h = asin(dx / xy);
  if (P[v * 3 + 1] < P[cur * 3 + 1]) h = 3.141592653589793 - h;
  h -= heading;
  const double el = asin(dz / xyz) - elevation;
  const float hf = (float)h, ef = (float)el;           // the reference casts the angles to fp32 before sin / cos
  out[0] = (float)sin((double)hf); out[1] = (float)cos((double)hf);
  out[2] = (float)sin((double)ef); out[3] = (float)cos((double)ef);
  out[4] = (float)(xyz / 30.0);
  out[5] = (float)((cur == v ? 0.0 : D[(long long)cur * N + v]) / 30.0);
  out[6] = (float)((double)path_hops(T, N, cur, v) / 10.0);
}

__global__ void __launch_bounds__(128) graph_features_kernel(const double* pos, const double* dis, const int32_t* point, int N,
                                                             const int32_t* cur_node, const double* heading,
                                                             const double* elevation, const int32_t* gmap_node,
                                                             const int32_t* gmap_lens, int G, float* gmap_pos_fts,
                                                             float* gmap_pair_dists, const int32_t* cand_node, int C,
                                                             const int32_t* start_node, int Pn, float* vp_pos_fts) {
  pdl_enter();
  const int b = blockIdx.x;
  const double* P = pos + (long long)b * N * 3;
  const double* D = dis + (long long)b * N * N;
  const int32_t* T = point + (long long)b * N * N;
  const int cur = cur_node[b];
  const double hd = heading[b], el = elevation[b];
  const int len = gmap_lens[b];
  __shared__ float s_start[7];
  if (gmap_pos_fts) {
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      float f[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // rows beyond the episode's own length: zero padding
      if (g < len) rel_pos_fts(P, D, T, N, cur, gmap_node[(long long)b * G + g], hd, el, f);
#pragma unroll
      for (int i = 0; i < 7; ++i) gmap_pos_fts[((long long)b * G + g) * 7 + i] = f[i];
    }
  }
  if (gmap_pair_dists) {
    for (int e = threadIdx.x; e < G * G; e += blockDim.x) {       // agent.py:137-141: raw metres, zero row / column 0 and diagonal
      const int i = e / G, j = e - i * G;
      float d = 0.f;
      if (i >= 1 && j >= 1 && i != j && i < len && j < len) {
        const int ni = gmap_node[(long long)b * G + i], nj = gmap_node[(long long)b * G + j];
        d = (float)(ni == nj ? 0.0 : D[(long long)ni * N + nj]);
      }
      gmap_pair_dists[(long long)b * G * G + e] = d;
    }
  }
  if (vp_pos_fts) {                                               // agent.py:182-196
    if (threadIdx.x == 0) rel_pos_fts(P, D, T, N, cur, start_node[b], hd, el, s_start);
    __syncthreads();
    for (int r = threadIdx.x; r < Pn; r += blockDim.x) {
      float f[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int j = r - 1;
      if (j >= 0 && j < C && cand_node[(long long)b * C + j] >= 0) rel_pos_fts(P, D, T, N, cur, cand_node[(long long)b * C + j], hd, el, f);
      float* o = vp_pos_fts + ((long long)b * Pn + r) * 14;
#pragma unroll
      for (int i = 0; i < 7; ++i) { o[i] = s_start[i]; o[7 + i] = f[i]; }
    }
  }
}

}  // namespace

extern "C" int vi_graph_init(double* dis, int32_t* point, uint8_t* visited, float* ecnt, int B, int N, vi_stream_t stream) {
  VI_CHECK_ARG(dis && point && visited && ecnt, "vi_graph_init: null buffer");
  VI_CHECK_ARG(B > 0 && N > 0 && N <= 0xFFFF, "vi_graph_init: bad sizes B=%d N=%d", B, N);
  const long long pairs = (long long)B * N * N, nodes = (long long)B * N;
  const int grid = (int)((pairs + 255) / 256 < 1184 ? (pairs + 255) / 256 : 1184);      // 8 CTAs per SM at most
  VI_CUDA(vi_launch(graph_init_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), dis, point, visited,
                    ecnt, pairs, nodes));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_graph_update(double* pos, double* dis, int32_t* point, uint8_t* visited, int B, int N, const int32_t* cur_node,
                               const double* cur_pos, const int32_t* cand_node, const double* cand_pos, int C,
                               const int32_t* n_nodes, vi_stream_t stream) {
  VI_CHECK_ARG(pos && dis && point && visited && cur_node && cur_pos && n_nodes, "vi_graph_update: null buffer");
  VI_CHECK_ARG(B > 0 && N > 0 && C >= 0 && (C == 0 || (cand_node && cand_pos)), "vi_graph_update: bad sizes B=%d N=%d C=%d", B, N, C);
  VI_CUDA(vi_launch(graph_update_kernel, dim3(B), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), pos, dis, point, visited,
                    N, cur_node, cur_pos, cand_node, cand_pos, C, n_nodes));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_graph_embed_step(const float* pano_embeds, const uint8_t* pano_masks, int B, int V, int H,
                                   const int32_t* cur_node, const int32_t* cand_node, int C, const uint8_t* visited,
                                   float* esum, float* ecnt, int N, const int32_t* gmap_node, int G, float* gmap_img_embeds,
                                   float* vp_img_embeds, vi_stream_t stream) {
  VI_CHECK_ARG(pano_embeds && pano_masks && cur_node && visited && esum && ecnt, "vi_graph_embed_step: null buffer");
  VI_CHECK_ARG(B > 0 && V > 0 && H > 0 && H % 4 == 0 && N > 0 && C >= 0 && C <= V && (C == 0 || cand_node),
               "vi_graph_embed_step: bad sizes B=%d V=%d H=%d N=%d C=%d (H %% 4 == 0, C <= V)", B, V, H, N, C);
  VI_CHECK_ARG(!gmap_img_embeds || (gmap_node && G > 0), "vi_graph_embed_step: gmap_img_embeds needs gmap_node and G > 0");
  VI_CHECK_ARG((((uintptr_t)pano_embeds | (uintptr_t)esum | (uintptr_t)gmap_img_embeds | (uintptr_t)vp_img_embeds) & 15) == 0,
               "vi_graph_embed_step: embedding buffers must be 16-byte aligned");
  VI_CUDA(vi_launch(graph_embed_kernel, dim3(B), dim3(192), 0, reinterpret_cast<cudaStream_t>(stream), pano_embeds, pano_masks, V,
                    H, cur_node, cand_node, C, visited, esum, ecnt, N, gmap_node, G, gmap_img_embeds, vp_img_embeds));
  VI_LAUNCH_CHECK();
  return VI_OK;
}

extern "C" int vi_graph_features(const double* pos, const double* dis, const int32_t* point, int B, int N, const int32_t* cur_node,
                                 const double* heading, const double* elevation, const int32_t* gmap_node,
                                 const int32_t* gmap_lens, int G, float* gmap_pos_fts, float* gmap_pair_dists,
                                 const int32_t* cand_node, int C, const int32_t* start_node, int P, float* vp_pos_fts,
                                 vi_stream_t stream) {
  VI_CHECK_ARG(pos && dis && point && cur_node && heading && elevation, "vi_graph_features: null buffer");
  VI_CHECK_ARG(B > 0 && N > 0 && N <= 128, "vi_graph_features: bad sizes B=%d N=%d (N <= 128)", B, N);
  VI_CHECK_ARG((!gmap_pos_fts && !gmap_pair_dists) || (gmap_node && gmap_lens && G > 0), "vi_graph_features: gmap outputs need gmap_node, gmap_lens, G");
  VI_CHECK_ARG(!vp_pos_fts || (start_node && P > 0 && C >= 0 && (C == 0 || cand_node)), "vi_graph_features: vp output needs start_node, P, cand_node");
  VI_CUDA(vi_launch(graph_features_kernel, dim3(B), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), pos, dis, point, N,
                    cur_node, heading, elevation, gmap_node, gmap_lens, G, gmap_pos_fts, gmap_pair_dists, cand_node, C, start_node,
                    P, vp_pos_fts));
  VI_LAUNCH_CHECK();
  return VI_OK;
}
