// Fused DeepFM tower for the dac_ctr workload (include/b200_deepfm.h).
//
//   k_tower_prep : zero loss / grads / the live rows of the per-unique-id gradient buffers
//   k_tower_a    : lane = sample.  gather rows through inv, FM + DNN forward, loss, backward
//                  to the activations; embedding gradients are reduced per unique id on the
//                  fly (warp-level id dedup + one vector red per distinct row per warp)
//   k_tower_b    : thread = input column.  parameter gradients (dW1 is a [16 x B] x [B x IN]
//                  contraction over the batch: each thread owns one column and 16 accumulators,
//                  per-sample backward state staged through shared memory)
//
// fp32 throughout; no tensor cores (0.3 GFLOP per batch -- the kernel is bound by the
// gather/scatter of embedding rows, not by math).
#include <cuda_runtime.h>

#include <string>

#include "../../include/b200_deepfm.h"

namespace {

constexpr int ND = B200_DEEPFM_NDENSE, D = B200_DEEPFM_DIM, H1 = B200_DEEPFM_H1, H2 = B200_DEEPFM_H2;
constexpr int SCR = B200_DEEPFM_SCRATCH;
constexpr int WD_PAD = 16;  // w_dense padded to 16 floats

struct Layout {
  int in;  // ND + G*D
  int o_wd, o_w1, o_b1, o_w2, o_b2, o_w3, total;
};
__host__ __device__ inline Layout layout(int G) {
  Layout l;
  l.in = ND + G * D;
  l.o_wd = 0;
  l.o_w1 = WD_PAD;
  l.o_b1 = l.o_w1 + H1 * l.in;
  l.o_w2 = l.o_b1 + H1;
  l.o_b2 = l.o_w2 + H2 * H1;
  l.o_w3 = l.o_b2 + H2;
  l.total = l.o_w3 + H2;
  return l;
}

long long g_launches = 0;
thread_local std::string g_msg;

__global__ void __launch_bounds__(256) k_tower_prep(b200_deepfm_args_t a, int n_params) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0) *a.loss = 0.f;
  for (long long i = tid; i < n_params; i += stride) a.grads[i] = 0.f;
  // live rows only: rows >= n_unique[g] are never read by the push
  for (int g = blockIdx.y; g < a.G; g += gridDim.y) {
    const int u = a.n_unique[g];
    float* gw = a.gsum_wide + (long long)g * a.B;
    float4* gd = reinterpret_cast<float4*>(a.gsum_deep + (long long)g * a.B * D);
    for (long long i = tid; i < u; i += stride) gw[i] = 0.f;
    for (long long i = tid; i < 2LL * u; i += stride) gd[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// shared-memory copy of the parameters: W1 transposed to [IN][16] so that the 16 outputs of
// one input are four broadcast LDS.128.
struct SmemParams {
  float* w1t;  // [IN][H1]
  float* b1;   // [H1]
  float* w2;   // [H2][H1]
  float* b2;   // [H2]
  float* w3;   // [H2]
  float* wd;   // [ND]
};
__device__ inline SmemParams carve(float* smem, int in) {
  SmemParams s;
  s.w1t = smem;
  s.b1 = s.w1t + in * H1;
  s.w2 = s.b1 + H1;
  s.b2 = s.w2 + H2 * H1;
  s.w3 = s.b2 + H2;
  s.wd = s.w3 + H2;
  return s;
}
__host__ __device__ inline size_t smem_floats(int in) { return (size_t)in * H1 + H1 + H2 * H1 + H2 + H2 + WD_PAD; }

__device__ inline void load_params(const float* __restrict__ p, const Layout& l, SmemParams s) {
  for (int i = threadIdx.x; i < H1 * l.in; i += blockDim.x) {
    int j = i / l.in, e = i - j * l.in;  // coalesced read of w1[j][e]
    s.w1t[e * H1 + j] = p[l.o_w1 + i];
  }
  for (int i = threadIdx.x; i < H1; i += blockDim.x) s.b1[i] = p[l.o_b1 + i];
  for (int i = threadIdx.x; i < H2 * H1; i += blockDim.x) s.w2[i] = p[l.o_w2 + i];
  for (int i = threadIdx.x; i < H2; i += blockDim.x) { s.b2[i] = p[l.o_b2 + i]; s.w3[i] = p[l.o_w3 + i]; }
  for (int i = threadIdx.x; i < ND; i += blockDim.x) s.wd[i] = p[l.o_wd + i];
  __syncthreads();
}

__device__ __forceinline__ void axpy16(float (&h)[H1], const float* __restrict__ w, float x) {
  const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 v = w4[q];
    h[4 * q + 0] = fmaf(v.x, x, h[4 * q + 0]);
    h[4 * q + 1] = fmaf(v.y, x, h[4 * q + 1]);
    h[4 * q + 2] = fmaf(v.z, x, h[4 * q + 2]);
    h[4 * q + 3] = fmaf(v.w, x, h[4 * q + 3]);
  }
}
__device__ __forceinline__ float dot16(const float (&h)[H1], const float* __restrict__ w) {
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float4 v = w4[q];
    acc = fmaf(v.x, h[4 * q + 0], acc);
    acc = fmaf(v.y, h[4 * q + 1], acc);
    acc = fmaf(v.z, h[4 * q + 2], acc);
    acc = fmaf(v.w, h[4 * q + 3], acc);
  }
  return acc;
}

// k_tower_a work split: a block = SPB samples x NPART group-parts.  Warp w handles the samples
// (w / NPART) * 32 + lane and the id groups g = w % NPART, + NPART, ...  (lane = sample keeps every
// weight read a shared-memory broadcast; splitting the 38 groups over 4 warps quadruples the
// number of independent gather chains in flight -- the kernel is latency-bound, not math-bound).
constexpr int NPART = 4;
constexpr int SPB = 64;                        // samples per block
constexpr int TA_THREADS = SPB * NPART;        // 256
constexpr int GC = 3;                          // groups gathered per round per warp
constexpr int RW = H1 + D + 2;                 // partial-sum record: h[16] | s[8] | q | lin

template <bool BACKWARD>
__global__ void __launch_bounds__(TA_THREADS, 2) k_tower_a(b200_deepfm_args_t a) {
  extern __shared__ __align__(16) float smem[];
  const Layout l = layout(a.G);
  SmemParams sp = carve(smem, l.in);
  float* red = smem + smem_floats(l.in);                       // [NPART][RW][SPB]
  int* sinv = reinterpret_cast<int*>(red + NPART * RW * SPB);  // [G][SPB]
  load_params(a.params, l, sp);
  const int B = a.B, G = a.G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int part = warp % NPART;
  const int sl = (warp / NPART) * 32 + lane;  // sample slot within the block
  float loss_acc = 0.f;
  const long long nblk = ((long long)B + SPB - 1) / SPB;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long b = blk * SPB + sl;
    const bool live = b < B;
    const long long bb = live ? b : B - 1;  // dead lanes replay the last sample, results discarded
    float h[H1];
    float lin = 0.f;
    if (part == 0) {
#pragma unroll
      for (int j = 0; j < H1; ++j) h[j] = sp.b1[j];
#pragma unroll
      for (int e = 0; e < ND; ++e) {
        const float x = a.dense[bb * ND + e];
        lin = fmaf(sp.wd[e], x, lin);
        axpy16(h, sp.w1t + e * H1, x);
      }
    } else {
#pragma unroll
      for (int j = 0; j < H1; ++j) h[j] = 0.f;
    }
    float s[D];
#pragma unroll
    for (int d = 0; d < D; ++d) s[d] = 0.f;
    float q = 0.f;
    // pass 1 over this warp's groups: GC rank loads, then 3*GC independent row loads in flight
    for (int g0 = part; g0 < G; g0 += NPART * GC) {
      int r[GC];
#pragma unroll
      for (int u = 0; u < GC; ++u) {
        const int g = g0 + u * NPART;
        r[u] = g < G ? a.inv[(long long)g * B + bb] : 0;
      }
      float4 e0[GC], e1[GC];
      float wv[GC];
#pragma unroll
      for (int u = 0; u < GC; ++u) {
        const int g = g0 + u * NPART < G ? g0 + u * NPART : part;
        const float4* row = reinterpret_cast<const float4*>(a.bet_deep + ((long long)g * B + r[u]) * D);
        e0[u] = row[0];
        e1[u] = row[1];
        wv[u] = a.bet_wide[(long long)g * B + r[u]];
      }
#pragma unroll
      for (int u = 0; u < GC; ++u) {
        const int g = g0 + u * NPART;
        if (g < G) {
          sinv[g * SPB + sl] = r[u];
          lin += wv[u];
          const float ev[D] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
          const float* w = sp.w1t + (ND + g * D) * H1;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            s[d] += ev[d];
            q = fmaf(ev[d], ev[d], q);
            axpy16(h, w + d * H1, ev[d]);
          }
        }
      }
    }
    // combine the NPART partial sums of each sample through shared memory
    {
      float* mine = red + (part * RW) * SPB + sl;
#pragma unroll
      for (int j = 0; j < H1; ++j) mine[j * SPB] = h[j];
#pragma unroll
      for (int d = 0; d < D; ++d) mine[(H1 + d) * SPB] = s[d];
      mine[(H1 + D) * SPB] = q;
      mine[(H1 + D + 1) * SPB] = lin;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < H1; ++j) h[j] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) s[d] = 0.f;
    q = 0.f;
    lin = 0.f;
#pragma unroll
    for (int pt = 0; pt < NPART; ++pt) {
      const float* src = red + (pt * RW) * SPB + sl;
#pragma unroll
      for (int j = 0; j < H1; ++j) h[j] += src[j * SPB];
#pragma unroll
      for (int d = 0; d < D; ++d) s[d] += src[(H1 + d) * SPB];
      q += src[(H1 + D) * SPB];
      lin += src[(H1 + D + 1) * SPB];
    }
    float a1[H1];
#pragma unroll
    for (int j = 0; j < H1; ++j) a1[j] = fmaxf(h[j], 0.f);
    float h2[H2], dnn = 0.f;
#pragma unroll
    for (int k = 0; k < H2; ++k) {
      h2[k] = fmaxf(sp.b2[k] + dot16(a1, sp.w2 + k * H1), 0.f);
      dnn = fmaf(sp.w3[k], h2[k], dnn);
    }
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) ss = fmaf(s[d], s[d], ss);
    const float z = lin + dnn + 0.5f * (ss - q);
    if (part == 0 && live && a.logits != nullptr) a.logits[b] = z;
    if (BACKWARD) {
      const float y = a.labels[bb];
      const float lb = fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));  // BCE with logits
      if (part == 0 && live) loss_acc += lb;
      const float p = 1.f / (1.f + expf(-z));
      const float dz = live ? (p - y) / (float)B : 0.f;
      float dh2[H2];
#pragma unroll
      for (int k = 0; k < H2; ++k) dh2[k] = h2[k] > 0.f ? dz * sp.w3[k] : 0.f;
      float dh1[H1];
#pragma unroll
      for (int j = 0; j < H1; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < H2; ++k) acc = fmaf(sp.w2[k * H1 + j], dh2[k], acc);
        dh1[j] = h[j] > 0.f ? acc : 0.f;
      }
      if (part == 0 && live) {  // per-sample backward state for k_tower_b: [dh1 16 | a1 16 | dh2 4 | h2 4 | dz | pad 3]
        float4* sc = reinterpret_cast<float4*>(a.scratch + b * SCR);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) sc[qd] = make_float4(dh1[4 * qd], dh1[4 * qd + 1], dh1[4 * qd + 2], dh1[4 * qd + 3]);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) sc[4 + qd] = make_float4(a1[4 * qd], a1[4 * qd + 1], a1[4 * qd + 2], a1[4 * qd + 3]);
        sc[8] = make_float4(dh2[0], dh2[1], dh2[2], dh2[3]);
        sc[9] = make_float4(h2[0], h2[1], h2[2], h2[3]);
        sc[10] = make_float4(dz, 0.f, 0.f, 0.f);
      }
      // pass 2: d loss / d embedding rows of this warp's groups, reduced per unique id
      // (deduplicate_indexed_slices' sum)
      for (int g0 = part; g0 < G; g0 += NPART * GC) {
        int r[GC];
        float4 e0[GC], e1[GC];
#pragma unroll
        for (int u = 0; u < GC; ++u) {
          const int g = g0 + u * NPART < G ? g0 + u * NPART : part;
          r[u] = sinv[g * SPB + sl];
          const float4* row = reinterpret_cast<const float4*>(a.bet_deep + ((long long)g * B + r[u]) * D);
          e0[u] = row[0];
          e1[u] = row[1];
        }
#pragma unroll
        for (int u = 0; u < GC; ++u) {
          const int g = g0 + u * NPART;
          if (g >= G) break;  // warp-uniform
          const float ev[D] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
          const float* w = sp.w1t + (ND + g * D) * H1;
          float x[D + 1];
#pragma unroll
          for (int d = 0; d < D; ++d) x[d] = fmaf(dz, s[d] - ev[d], dot16(dh1, w + d * H1));
          x[D] = dz;  // wide row gradient
          // warp-level id dedup: lanes hitting the same row combine (lane order), lowest lane writes
          const int key = live ? r[u] : -1 - lane;
          const unsigned peers = __match_any_sync(0xffffffffu, key);
          const bool leader = (__ffs(peers) - 1) == lane;
          unsigned rest = peers & ~(1u << lane);
          const int maxn = __reduce_max_sync(0xffffffffu, (unsigned)__popc(peers));
          for (int it = 1; it < maxn; ++it) {
            const int src = rest ? __ffs(rest) - 1 : lane;
#pragma unroll
            for (int e = 0; e <= D; ++e) {
              const float yv = __shfl_sync(0xffffffffu, x[e], src);
              if (leader && rest) x[e] += yv;
            }
            rest &= rest - 1;
          }
          if (live && leader) {
            float* od = a.gsum_deep + ((long long)g * B + r[u]) * D;
            atomicAdd(reinterpret_cast<float4*>(od), make_float4(x[0], x[1], x[2], x[3]));
            atomicAdd(reinterpret_cast<float4*>(od + 4), make_float4(x[4], x[5], x[6], x[7]));
            atomicAdd(a.gsum_wide + (long long)g * B + r[u], x[D]);
          }
        }
      }
    }
    __syncthreads();  // red / sinv are reused by the next block of samples
  }
  if (BACKWARD) {
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, o);
    if (lane == 0 && part == 0) atomicAdd(a.loss, loss_acc / (float)B);
  }
}

__host__ inline size_t tower_a_smem(int G) {
  return (smem_floats(ND + G * D) + (size_t)NPART * RW * SPB) * sizeof(float) + (size_t)G * SPB * sizeof(int);
}

// parameter gradients.  dW1 = dH1^T [16 x B] . X [B x IN] is a contraction over the batch: a block
// stages a chunk of S samples in shared memory -- the per-sample backward state written by
// k_tower_a and the gathered input rows X (cooperative, fully independent 32 B row gathers) --
// then thread e owns input column e with 16 accumulators; the remaining threads own db1 and
// the small gradients (dW2, db2, dw3, dw_dense).  One atomicAdd per output per block.
constexpr int S_CHUNK = 64;
constexpr int TB_THREADS = 416;
constexpr int N_SMALL = H2 * H1 + H2 + H2 + ND;  // dW2 64 | db2 4 | dw3 4 | dw_dense 13

__host__ __device__ inline int tb_xpad(int G) { return (G * D + ND + 3) / 4 * 4; }  // tile row: [deep G*8 | dense 13 | pad]
__host__ inline size_t tower_b_smem(int G) {
  return ((size_t)S_CHUNK * SCR + (size_t)S_CHUNK * tb_xpad(G)) * sizeof(float) + (size_t)G * S_CHUNK * sizeof(int);
}

__global__ void __launch_bounds__(TB_THREADS) k_tower_b(b200_deepfm_args_t a) {
  extern __shared__ __align__(16) float smem_b[];
  const Layout l = layout(a.G);
  const int B = a.B, G = a.G, t = threadIdx.x;
  const int IN = l.in, XP = tb_xpad(G), NDEEP = G * D;
  float* sc = smem_b;                                   // [S][SCR]
  float* xt = sc + S_CHUNK * SCR;                       // [S][XP]
  int* sinv = reinterpret_cast<int*>(xt + S_CHUNK * XP);  // [G][S]
  // role of this thread
  const bool is_col = t < IN;      // tile column t -> dW1 column (deep columns first in the tile)
  const bool is_b1 = t == IN;      // db1
  const int small = t - (IN + 1);  // small outputs
  const bool is_small = small >= 0 && small < N_SMALL;
  float acc[H1];
#pragma unroll
  for (int j = 0; j < H1; ++j) acc[j] = 0.f;
  float sacc = 0.f;
  const long long nchunk = ((long long)B + S_CHUNK - 1) / S_CHUNK;
  for (long long c = blockIdx.x; c < nchunk; c += gridDim.x) {
    const long long b0 = c * S_CHUNK;
    const int n = (int)min((long long)S_CHUNK, B - b0);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(a.scratch + b0 * SCR);
      float4* dst = reinterpret_cast<float4*>(sc);
      for (int i = t; i < n * (SCR / 4); i += TB_THREADS) dst[i] = src[i];
      for (int i = t; i < n * ND; i += TB_THREADS) {
        const int s_ = i / ND, e = i - s_ * ND;
        xt[s_ * XP + NDEEP + e] = a.dense[b0 * ND + i];
      }
      for (int i = t; i < G * n; i += TB_THREADS) {  // coalesced rank loads
        const int g = i / n, s_ = i - g * n;
        sinv[g * S_CHUNK + s_] = a.inv[(long long)g * B + b0 + s_];
      }
    }
    __syncthreads();
    for (int i = t; i < n * G; i += TB_THREADS) {  // independent 32 B row gathers, group fastest
      const int s_ = i / G, g = i - s_ * G;
      const float4* row = reinterpret_cast<const float4*>(a.bet_deep + ((long long)g * B + sinv[g * S_CHUNK + s_]) * D);
      const float4 e0 = row[0], e1 = row[1];
      float4* dst = reinterpret_cast<float4*>(xt + s_ * XP + g * D);
      dst[0] = e0;
      dst[1] = e1;
    }
    __syncthreads();
    if (is_col) {
#pragma unroll 4
      for (int s_ = 0; s_ < n; ++s_) axpy16(acc, sc + s_ * SCR, xt[s_ * XP + t]);
    } else if (is_b1) {
      for (int s_ = 0; s_ < n; ++s_) axpy16(acc, sc + s_ * SCR, 1.0f);
    } else if (is_small) {
      for (int s_ = 0; s_ < n; ++s_) {
        const float* r = sc + s_ * SCR;
        float v;
        if (small < H2 * H1) v = r[32 + small / H1] * r[16 + small % H1];            // dh2[k] * a1[j]
        else if (small < H2 * H1 + H2) v = r[32 + small - H2 * H1];                   // dh2[k]
        else if (small < H2 * H1 + 2 * H2) v = r[40] * r[36 + small - H2 * H1 - H2];  // dz * h2[k]
        else v = r[40] * xt[s_ * XP + NDEEP + small - H2 * H1 - 2 * H2];              // dz * dense[e]
        sacc += v;
      }
    }
  }
  if (is_col) {
    const int col = t < NDEEP ? ND + t : t - NDEEP;  // tile column -> W1 input index (dense first)
#pragma unroll
    for (int j = 0; j < H1; ++j) atomicAdd(a.grads + l.o_w1 + j * IN + col, acc[j]);
  } else if (is_b1) {
#pragma unroll
    for (int j = 0; j < H1; ++j) atomicAdd(a.grads + l.o_b1 + j, acc[j]);
  } else if (is_small) {
    int off;
    if (small < H2 * H1) off = l.o_w2 + small;
    else if (small < H2 * H1 + H2) off = l.o_b2 + small - H2 * H1;
    else if (small < H2 * H1 + 2 * H2) off = l.o_w3 + small - H2 * H1 - H2;
    else off = l.o_wd + small - H2 * H1 - 2 * H2;
    atomicAdd(a.grads + off, sacc);
  }
}

int check_args(const b200_deepfm_args_t* a, bool backward) {
  if (!a || a->G < 1 || a->B < 1) { g_msg = "bad shape"; return -1; }
  if (ND + a->G * D + 1 + N_SMALL > TB_THREADS) { g_msg = "too many id groups for the fused tower (max 39)"; return -1; }
  if (!a->inv || !a->bet_wide || !a->bet_deep || !a->dense || !a->params) { g_msg = "null input"; return -1; }
  if (backward && (!a->labels || !a->grads || !a->gsum_wide || !a->gsum_deep || !a->loss || !a->scratch || !a->n_unique)) {
    g_msg = "null output";
    return -1;
  }
  return 0;
}

}  // namespace

extern "C" {

size_t b200_deepfm_param_count(int G) { return (size_t)layout(G).total; }
int64_t b200_deepfm_launch_count(void) { return g_launches; }

int b200_deepfm_fwd_bwd(const b200_deepfm_args_t* args, void* stream) {
  if (check_args(args, true)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const Layout l = layout(args->G);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = tower_a_smem(args->G);
  static bool attr_done[64] = {false};
  if (dev < 64 && !attr_done[dev]) {
    cudaFuncSetAttribute(k_tower_a<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_tower_a<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done[dev] = true;
  }
  dim3 gp(8, args->G);
  k_tower_prep<<<gp, 256, 0, st>>>(*args, l.total);
  long long blocks = ((long long)args->B + SPB - 1) / SPB;
  long long cap = (long long)n_sm * 8;
  k_tower_a<true><<<(unsigned)(blocks < cap ? blocks : cap), TA_THREADS, smem, st>>>(*args);
  long long chunks = ((long long)args->B + S_CHUNK - 1) / S_CHUNK;
  cap = (long long)n_sm * 2;
  static bool attr_b[64] = {false};
  if (dev < 64 && !attr_b[dev]) {
    cudaFuncSetAttribute(k_tower_b, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    attr_b[dev] = true;
  }
  k_tower_b<<<(unsigned)(chunks < cap ? chunks : cap), TB_THREADS, tower_b_smem(args->G), st>>>(*args);
  g_launches += 3;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_msg = cudaGetErrorString(e); return -2; }
  return 0;
}

int b200_deepfm_forward(const b200_deepfm_args_t* args, void* stream) {
  if (check_args(args, false) || !args->logits) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const Layout l = layout(args->G);
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = tower_a_smem(args->G);
  cudaFuncSetAttribute(k_tower_a<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long blocks = ((long long)args->B + SPB - 1) / SPB;
  long long cap = (long long)n_sm * 8;
  k_tower_a<false><<<(unsigned)(blocks < cap ? blocks : cap), TA_THREADS, smem, st>>>(*args);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_msg = cudaGetErrorString(e); return -2; }
  return 0;
}

}  // extern "C"
