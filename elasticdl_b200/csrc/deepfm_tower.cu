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

constexpr int TA_THREADS = 128;  // samples per block (lane = sample)
constexpr int GC = 4;            // id groups gathered per batch of loads (memory-level parallelism)

template <bool BACKWARD>
__global__ void __launch_bounds__(TA_THREADS) k_tower_a(b200_deepfm_args_t a) {
  extern __shared__ __align__(16) float smem[];
  const Layout l = layout(a.G);
  SmemParams sp = carve(smem, l.in);
  int* sinv = reinterpret_cast<int*>(smem + smem_floats(l.in));  // [G][TA_THREADS] ranks of this block's samples
  load_params(a.params, l, sp);
  const int B = a.B, G = a.G;
  const int lane = threadIdx.x & 31;
  float loss_acc = 0.f;
  const long long nblk = ((long long)B + TA_THREADS - 1) / TA_THREADS;
  for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long b = blk * TA_THREADS + threadIdx.x;
    const bool live = b < B;
    const long long bb = live ? b : B - 1;  // dead lanes replay the last sample, results discarded
    float h[H1];
#pragma unroll
    for (int j = 0; j < H1; ++j) h[j] = sp.b1[j];
    float lin = 0.f;
#pragma unroll
    for (int e = 0; e < ND; ++e) {
      const float x = a.dense[bb * ND + e];
      lin = fmaf(sp.wd[e], x, lin);
      axpy16(h, sp.w1t + e * H1, x);
    }
    float s[D];
#pragma unroll
    for (int d = 0; d < D; ++d) s[d] = 0.f;
    float q = 0.f;
    // pass 1: GC groups per round -- GC rank loads, then 3*GC independent row loads in flight
    for (int g0 = 0; g0 < G; g0 += GC) {
      int r[GC];
#pragma unroll
      for (int u = 0; u < GC; ++u) r[u] = (g0 + u < G) ? a.inv[(long long)(g0 + u) * B + bb] : 0;
      float4 e0[GC], e1[GC];
      float wv[GC];
#pragma unroll
      for (int u = 0; u < GC; ++u) {
        const int g = (g0 + u < G) ? g0 + u : G - 1;
        const float4* row = reinterpret_cast<const float4*>(a.bet_deep + ((long long)g * B + r[u]) * D);
        e0[u] = row[0];
        e1[u] = row[1];
        wv[u] = a.bet_wide[(long long)g * B + r[u]];
      }
#pragma unroll
      for (int u = 0; u < GC; ++u) {
        if (g0 + u < G) {
          sinv[(g0 + u) * TA_THREADS + threadIdx.x] = r[u];
          lin += wv[u];
          const float ev[D] = {e0[u].x, e0[u].y, e0[u].z, e0[u].w, e1[u].x, e1[u].y, e1[u].z, e1[u].w};
          const float* w = sp.w1t + (ND + (g0 + u) * D) * H1;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            s[d] += ev[d];
            q = fmaf(ev[d], ev[d], q);
            axpy16(h, w + d * H1, ev[d]);
          }
        }
      }
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
    if (live && a.logits != nullptr) a.logits[b] = z;
    if (BACKWARD) {
      const float y = a.labels[bb];
      const float lb = fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));  // BCE with logits
      if (live) loss_acc += lb;
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
      if (live) {  // per-sample backward state for k_tower_b: [dh1 16 | a1 16 | dh2 4 | h2 4 | dz | pad 3]
        float4* sc = reinterpret_cast<float4*>(a.scratch + b * SCR);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) sc[qd] = make_float4(dh1[4 * qd], dh1[4 * qd + 1], dh1[4 * qd + 2], dh1[4 * qd + 3]);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) sc[4 + qd] = make_float4(a1[4 * qd], a1[4 * qd + 1], a1[4 * qd + 2], a1[4 * qd + 3]);
        sc[8] = make_float4(dh2[0], dh2[1], dh2[2], dh2[3]);
        sc[9] = make_float4(h2[0], h2[1], h2[2], h2[3]);
        sc[10] = make_float4(dz, 0.f, 0.f, 0.f);
      }
      // pass 2: d loss / d embedding rows, reduced per unique id (deduplicate_indexed_slices' sum)
      for (int g0 = 0; g0 < G; g0 += GC) {
        int r[GC];
        float4 e0[GC], e1[GC];
#pragma unroll
        for (int u = 0; u < GC; ++u) {
          const int g = (g0 + u < G) ? g0 + u : G - 1;
          r[u] = sinv[g * TA_THREADS + threadIdx.x];
          const float4* row = reinterpret_cast<const float4*>(a.bet_deep + ((long long)g * B + r[u]) * D);
          e0[u] = row[0];
          e1[u] = row[1];
        }
#pragma unroll
        for (int u = 0; u < GC; ++u) {
          if (g0 + u >= G) break;  // warp-uniform
          const int g = g0 + u;
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
  }
  if (BACKWARD) {
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) atomicAdd(a.loss, loss_acc / (float)B);
  }
}

// parameter gradients.  Block = 416 threads, chunk of S samples staged in shared memory.
constexpr int S_CHUNK = 128;
constexpr int TB_THREADS = 416;
constexpr int N_SMALL = H2 * H1 + H2 + H2 + ND;  // dW2 64 | db2 4 | dw3 4 | dw_dense 13

__global__ void __launch_bounds__(TB_THREADS) k_tower_b(b200_deepfm_args_t a) {
  __shared__ __align__(16) float sc[S_CHUNK * SCR];
  __shared__ float dn[S_CHUNK * ND];
  const Layout l = layout(a.G);
  const int B = a.B, t = threadIdx.x;
  const int IN = l.in;
  // role of this thread
  const bool is_col = t < IN;         // dW1 column t
  const bool is_b1 = t == IN;         // db1
  const int small = t - (IN + 1);     // small outputs
  const bool is_small = small >= 0 && small < N_SMALL;
  const int g = is_col && t >= ND ? (t - ND) / D : 0;
  const int d = is_col && t >= ND ? (t - ND) % D : 0;
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
      for (int i = t; i < n * ND; i += TB_THREADS) dn[i] = a.dense[b0 * ND + i];
    }
    __syncthreads();
    if (is_col) {
      constexpr int U = 8;  // samples in flight per thread
      for (int s0 = 0; s0 < n; s0 += U) {
        float x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int s = s0 + u;
          if (s < n) {
            if (t < ND) x[u] = dn[s * ND + t];
            else {
              const int r = a.inv[(long long)g * B + b0 + s];
              x[u] = a.bet_deep[((long long)g * B + r) * D + d];
            }
          } else x[u] = 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (s0 + u < n) axpy16(acc, sc + (s0 + u) * SCR, x[u]);
        }
      }
    } else if (is_b1) {
      for (int s = 0; s < n; ++s) axpy16(acc, sc + s * SCR, 1.0f);
    } else if (is_small) {
      for (int s = 0; s < n; ++s) {
        const float* r = sc + s * SCR;
        float v;
        if (small < H2 * H1) v = r[32 + small / H1] * r[16 + small % H1];           // dh2[k] * a1[j]
        else if (small < H2 * H1 + H2) v = r[32 + small - H2 * H1];                  // dh2[k]
        else if (small < H2 * H1 + 2 * H2) v = r[40] * r[36 + small - H2 * H1 - H2];  // dz * h2[k]
        else v = r[40] * dn[s * ND + small - H2 * H1 - 2 * H2];                       // dz * dense[e]
        sacc += v;
      }
    }
  }
  if (is_col) {
#pragma unroll
    for (int j = 0; j < H1; ++j) atomicAdd(a.grads + l.o_w1 + j * IN + t, acc[j]);
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
  const size_t smem = smem_floats(l.in) * sizeof(float) + (size_t)args->G * TA_THREADS * sizeof(int);
  static bool attr_done[64] = {false};
  if (smem > 48 * 1024 && dev < 64 && !attr_done[dev]) {
    cudaFuncSetAttribute(k_tower_a<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k_tower_a<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done[dev] = true;
  }
  dim3 gp(8, args->G);
  k_tower_prep<<<gp, 256, 0, st>>>(*args, l.total);
  long long blocks = ((long long)args->B + TA_THREADS - 1) / TA_THREADS;
  long long cap = (long long)n_sm * 4;
  k_tower_a<true><<<(unsigned)(blocks < cap ? blocks : cap), TA_THREADS, smem, st>>>(*args);
  long long chunks = ((long long)args->B + S_CHUNK - 1) / S_CHUNK;
  cap = (long long)n_sm * 2;
  k_tower_b<<<(unsigned)(chunks < cap ? chunks : cap), TB_THREADS, 0, st>>>(*args);
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
  const size_t smem = smem_floats(l.in) * sizeof(float) + (size_t)args->G * TA_THREADS * sizeof(int);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_tower_a<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long blocks = ((long long)args->B + TA_THREADS - 1) / TA_THREADS;
  long long cap = (long long)n_sm * 4;
  k_tower_a<false><<<(unsigned)(blocks < cap ? blocks : cap), TA_THREADS, smem, st>>>(*args);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_msg = cudaGetErrorString(e); return -2; }
  return 0;
}

}  // extern "C"
