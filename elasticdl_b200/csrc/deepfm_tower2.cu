// DeepFM tower, tile version (include/b200_deepfm.h: b200_deepfm_fwd_bwd_tile / b200_deepfm_forward_tile).
//
// Round 1's tower gathered every embedding row three times (forward, backward, parameter-gradient
// kernel) with one dependent global gather chain per lane at ~20 % occupancy: 113 us of a 198 us
// step.  Here a CTA owns a tile of 32 samples and gathers the tile's 38 x 32 rows ONCE, with
// cp.async (16 B, L2 -> shared memory, no registers held while in flight), into
//     X[s] = [ deep 38 x 8 | dense 13 | 1 (bias column) | 0 0 | pad ]     (row stride 324 floats = 81 x 16 B,
//            odd, so a warp's 32 rows are bank-conflict free for 128-bit reads; wide values in WV[g][s])
// and everything else -- H1 = X W1x, the FM sums, the per-sample middle, dX = dH1 W1x^T + FM term with the
// per-unique-id reduction, dW1x += dH1^T X and the small gradients -- runs out of shared memory:
//   G  gather             5 (group, sample) pairs per thread: rank load, 2 cp.async + 1 wide load
//   F  forward            warp = 40 tile columns (5 id groups), lane = sample: 16 accumulators + FM partials
//   M  middle             partial records summed through shared memory; every warp redoes the tiny
//                         per-sample MLP tail (it needs dh1 / dz / FM sums in registers for B1 anyway)
//   B1 embedding grads    warp = its 5 groups, lane = sample; equal rows of a warp combine up a tree
//                         (warp-level id dedup), one vector red per distinct row
//   B2 parameter grads    thread = tile column t x 16 hidden units + 4 hidden units of column 256 + t % 64
//                         (all 8 warps busy), accumulated in registers across the CTA's tiles, one atomic
//                         per output per CTA at the end
// W1 is presented in TILE COLUMN ORDER (W1x[c][j]: deep columns first, then dense, then b1 as the weight
// of the constant-1 column) by the prep kernel, so the bias needs no special case and every CTA loads it
// with straight 16-byte copies.  fp32 SIMT: the step has 1 GFLOP of first-layer work -- the kernel is
// bound by the row gather and shared-memory bandwidth, not by FMA throughput; tensor cores would need
// 3xTF32 splitting to keep the fp32-grade parity with the torch reference the tests check.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "../../include/b200_deepfm.h"

namespace {

constexpr int ND = B200_DEEPFM_NDENSE, D = B200_DEEPFM_DIM, H1 = B200_DEEPFM_H1, H2 = B200_DEEPFM_H2;
constexpr int WD_PAD = 16;
constexpr int TS = 32;        // samples per tile
constexpr int THREADS = 256;  // 8 warps
constexpr int NKP = 8;        // column parts (warps) of phases F / B1
constexpr int KPC = 40;       // tile columns per part = 5 id groups
constexpr int NCOL = 320;     // deep 304 | dense 13 | 1 | 0 0   (G <= 38)
constexpr int XS = 324;       // row stride: 320 + pad = 81 * 4 (odd number of 16 B quads: conflict-free 128-bit row reads)
constexpr int RS = 39;        // rank row stride (odd): Rt[s][g]
constexpr int RW = H1 + D + 2;  // partial record: h[16] | s[8] | q | lin
constexpr int MS = 44;        // per-sample backward state: dh1 16 | a1 16 | dh2 4 | h2 4 | dz | pad 3
constexpr int MAXG = 38;

struct Layout {
  int in, o_wd, o_w1, o_b1, o_w2, o_b2, o_w3, total;
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

// shared memory carve-up (floats)
constexpr int SM_W1X = NCOL * H1;          // 5120
constexpr int SM_X = TS * XS;              // 11648
constexpr int SM_P = NKP * RW * TS;        // 6656   partial records [kp][r][s]; reused as T[r][s] + M[s][MS]
constexpr int SM_R = 2 * TS * RS;          // 2496   ranks Rt[buffer][s][g]: this tile's and the prefetched next tile's
constexpr int SM_WV = MAXG * TS;           // 1216   wide values WV[g][s]
constexpr int SM_SMALL = 128;              // b... w2 64 | b2 4 | w3 4 | wd 16
constexpr size_t SMEM_BYTES = (size_t)(SM_W1X + SM_X + SM_P + SM_R + SM_WV + SM_SMALL) * sizeof(float);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Tile column c of W1x -> index into the flat parameter / gradient layout, or -1 (padding).
__host__ __device__ inline int w1x_src(const Layout& l, int G, int c, int j) {
  const int ndeep = G * D;
  if (c < ndeep) return l.o_w1 + j * l.in + ND + c;
  if (c >= MAXG * D && c < MAXG * D + ND) return l.o_w1 + j * l.in + (c - MAXG * D);
  if (c == MAXG * D + ND) return l.o_b1 + j;
  return -1;
}

__global__ void __launch_bounds__(256) k_tile_prep(b200_deepfm_args_t a, int n_params, float* w1x, int main_grid) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid == 0 && a.loss) *a.loss = 0.f;
  if (tid == 0 && blockIdx.y == 0) reinterpret_cast<int*>(w1x + NCOL * H1)[0] = main_grid;  // dynamic tile counter
  if (blockIdx.y == 0) {
    if (a.grads)
      for (long long i = tid; i < n_params; i += stride) a.grads[i] = 0.f;
    const Layout l = layout(a.G);
    for (long long i = tid; i < (long long)NCOL * H1; i += stride) {
      const int c = (int)(i / H1), j = (int)(i - (long long)c * H1);
      const int src = w1x_src(l, a.G, c, j);
      w1x[i] = src >= 0 ? a.params[src] : 0.f;
    }
  }
  if (a.gsum_deep == nullptr) return;
  for (int g = blockIdx.y; g < a.G; g += gridDim.y) {  // live rows only: rows >= n_unique[g] are never read by the push
    const int u = a.n_unique[g];
    float* gw = a.gsum_wide + (long long)g * a.B;
    float4* gd = reinterpret_cast<float4*>(a.gsum_deep + (long long)g * a.B * D);
    for (long long i = tid; i < u; i += stride) gw[i] = 0.f;
    for (long long i = tid; i < 2LL * u; i += stride) gd[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool BACKWARD>
__global__ void __launch_bounds__(THREADS, 2) k_tower_tile(b200_deepfm_args_t a, const float* __restrict__ w1x_g) {
  extern __shared__ __align__(16) float smem[];
  float* W1x = smem;                 // [NCOL][16]
  float* X = W1x + SM_W1X;           // [TS][XS]
  float* P = X + SM_X;               // [NKP][RW][TS], later T[RW][TS] and Mst[TS][MS]
  int* Rt0 = reinterpret_cast<int*>(P + SM_P);  // [2][TS][RS] ranks
  float* WV = reinterpret_cast<float*>(Rt0 + SM_R);  // [G][TS] wide values
  float* small = WV + SM_WV;
  float* s_w2 = small;        // [H2][H1]
  float* s_b2 = small + 64;   // [H2]
  float* s_w3 = small + 68;   // [H2]
  float* s_wd = small + 72;   // [ND]
  float* T = P;                          // [RW][TS] summed records (first RW*TS floats of P)
  float* Mst = P + RW * TS + 64;         // [TS][MS] per-sample backward state for B2 (after T, 16 B aligned)

  const Layout l = layout(a.G);
  const int B = a.B, G = a.G;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  {
    const float4* src = reinterpret_cast<const float4*>(w1x_g);
    float4* dst = reinterpret_cast<float4*>(W1x);
    for (int i = t; i < SM_W1X / 4; i += THREADS) dst[i] = src[i];
    for (int i = t; i < H2 * H1; i += THREADS) s_w2[i] = a.params[l.o_w2 + i];
    if (t < H2) { s_b2[t] = a.params[l.o_b2 + t]; s_w3[t] = a.params[l.o_w3 + t]; }
    if (t < ND) s_wd[t] = a.params[l.o_wd + t];
  }
  // B2 accumulators, persistent over the CTA's tiles, all 8 warps busy: thread t owns tile column t (16 hidden
  // units) and, of the remaining 64 columns, column 256 + t%64 x hidden units 4*(t/64) .. +3
  float2 acc0[H1 / 2], acc1[2];
  float sacc = 0.f;  // threads t < 85: one small output each (dW2 64 | db2 4 | dw3 4 | dwd 13)
#pragma unroll
  for (int j = 0; j < H1 / 2; ++j) acc0[j] = make_float2(0.f, 0.f);
  acc1[0] = acc1[1] = make_float2(0.f, 0.f);
  float loss_acc = 0.f;
  __syncthreads();

  const long long ntile = ((long long)B + TS - 1) / TS;
  // ranks of a tile: 4-byte cp.async, coalesced reads of inv (32 consecutive samples of one group per warp)
  auto fetch_ranks = [&](const long long tl, int* dst) {
    const long long bb0 = tl * TS;
    for (int i = t; i < G * TS; i += THREADS) {
      long long b = bb0 + (i & 31);
      if (b >= B) b = B - 1;  // dead samples replay the last one, their results are masked
      cp_async4(dst + (i & 31) * RS + (i >> 5), a.inv + (long long)(i >> 5) * B + b);
    }
  };
  // Tiles are handed out dynamically (one atomic per tile on a counter the prologue set to gridDim.x): with
  // 1024 tiles on 296 resident CTAs a static round-robin leaves SMs with 8 tiles next to SMs with 6.  A CTA
  // needs the index of its NEXT tile while it gathers the current one (rank prefetch), so two slots rotate:
  // s_tile[k & 1] = tile of iteration k, written during iteration k - 2 (slot reuse) by thread 0.
  __shared__ int s_tile[2];
  int* tile_counter = reinterpret_cast<int*>(const_cast<float*>(w1x_g) + SM_W1X);
  if (t == 0) {
    s_tile[0] = (int)blockIdx.x;
    s_tile[1] = atomicAdd(tile_counter, 1);
  }
  int cur = 0;
  if ((long long)blockIdx.x < ntile) fetch_ranks(blockIdx.x, Rt0);
  cp_async_wait_all();
  __syncthreads();
  for (int iter = 0;; ++iter, cur ^= 1) {
    const int tile = s_tile[iter & 1], tile_next = s_tile[(iter + 1) & 1];
    if (tile >= ntile) break;
    const long long b0 = (long long)tile * TS;
    int* Rt = Rt0 + cur * (TS * RS);
    // ---------------- G: gather the tile once ----------------
    // The tile's ranks are already in Rt (prefetched while the previous tile was computed), so the gather is
    // ONE global round trip: the row copies of this tile and the rank prefetch of the next one fly together.
    // Rows: lane = consecutive 16 B chunk of ONE sample's row (contiguous shared-memory destinations).
    {
      const int npair = G * TS;
      constexpr int NFILL = (TS * 16 + THREADS - 1) / THREADS;  // dense | 1 | 0 0 columns: 2 per thread
      float dv[NFILL];
#pragma unroll
      for (int jj = 0; jj < NFILL; ++jj) {  // tile columns 304..319 of sample s: dense 13 | 1 | 0 | 0
        const int i = t + THREADS * jj;
        const int s = i >> 4, c = i & 15;
        long long b = b0 + s;
        if (b >= B) b = B - 1;
        float v = c == ND ? 1.0f : 0.f;
        if (i < TS * 16 && c < ND) v = a.dense[b * ND + c];
        dv[jj] = v;
      }
      const int nchunk = G * 2;  // 16 B chunks per sample row
      for (int i = t; i < TS * nchunk; i += THREADS) {
        const int s = i / nchunk, ch = i - s * nchunk;
        const int g = ch >> 1;
        const int r = Rt[s * RS + g];
        cp_async16(X + s * XS + ch * 4, a.bet_deep + ((long long)g * B + r) * D + (ch & 1) * 4);
      }
      for (int i = t; i < npair; i += THREADS) {  // wide values: lane = sample
        const int g = i >> 5, s = i & 31;
        cp_async4(WV + g * TS + s, a.bet_wide + (long long)g * B + Rt[s * RS + g]);
      }
      if (tile_next < ntile) fetch_ranks(tile_next, Rt0 + (cur ^ 1) * (TS * RS));
#pragma unroll
      for (int jj = 0; jj < NFILL; ++jj) {
        const int i = t + THREADS * jj;
        if (i < TS * 16) X[(i >> 4) * XS + MAXG * D + (i & 15)] = dv[jj];
      }
      for (int i = t; i < TS * (MAXG - G) * D; i += THREADS) {  // unused group columns (G < 38): zeros
        const int s = i / ((MAXG - G) * D), c = G * D + i % ((MAXG - G) * D);
        X[s * XS + c] = 0.f;
      }
      cp_async_wait_all();
    }
    __syncthreads();
    if (t == 0) s_tile[iter & 1] = atomicAdd(tile_counter, 1);  // tile of iteration iter + 2
    const bool live = b0 + lane < B;
    // ---------------- F: partial H1 / FM sums over this warp's 40 columns ----------------
    {
      float2 h2[H1 / 2];
#pragma unroll
      for (int j = 0; j < H1 / 2; ++j) h2[j] = make_float2(0.f, 0.f);
      float sd[D], q = 0.f, lin = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) sd[d] = 0.f;
      const float* xrow = X + lane * XS;
      const int c0 = warp * KPC;
#pragma unroll
      for (int gi = 0; gi < KPC / D; ++gi) {
        const int c = c0 + gi * D;
        const float4 xa = *reinterpret_cast<const float4*>(xrow + c);
        const float4 xb = *reinterpret_cast<const float4*>(xrow + c + 4);
        const float xv[D] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const int g = c / D;
        if (g < G) {  // an embedding group: FM sums + the wide value
#pragma unroll
          for (int d = 0; d < D; ++d) { sd[d] += xv[d]; q = fmaf(xv[d], xv[d], q); }
          lin += WV[g * TS + lane];
        } else if (c == MAXG * D) {  // dense columns 304..311 and (next chunk) 312..316: the linear part
#pragma unroll
          for (int d = 0; d < D; ++d) lin = fmaf(s_wd[d], xv[d], lin);
        } else if (c == MAXG * D + D) {
#pragma unroll
          for (int d = 0; d < ND - D; ++d) lin = fmaf(s_wd[D + d], xv[d], lin);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float4* w4 = reinterpret_cast<const float4*>(W1x + (c + d) * H1);
          const float2 xx = make_float2(xv[d], xv[d]);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {  // packed fp32 FMA (FFMA2, sm_100): two hidden units per instruction
            const float4 w = w4[qd];
            h2[2 * qd] = __ffma2_rn(make_float2(w.x, w.y), xx, h2[2 * qd]);
            h2[2 * qd + 1] = __ffma2_rn(make_float2(w.z, w.w), xx, h2[2 * qd + 1]);
          }
        }
      }
      float* mine = P + (warp * RW) * TS + lane;
#pragma unroll
      for (int j = 0; j < H1 / 2; ++j) {
        mine[(2 * j) * TS] = h2[j].x;
        mine[(2 * j + 1) * TS] = h2[j].y;
      }
#pragma unroll
      for (int d = 0; d < D; ++d) mine[(H1 + d) * TS] = sd[d];
      mine[(H1 + D) * TS] = q;
      mine[(H1 + D + 1) * TS] = lin;
    }
    __syncthreads();
    // ---------------- M: sum the 8 partial records ----------------
    float tsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = warp + 8 * i;
      float v = 0.f;
      if (r < RW) {
#pragma unroll
        for (int kp = 0; kp < NKP; ++kp) v += P[(kp * RW + r) * TS + lane];
      }
      tsum[i] = v;
    }
    __syncthreads();  // all partials read before T (aliasing P) is written
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = warp + 8 * i;
      if (r < RW) T[r * TS + lane] = tsum[i];
    }
    __syncthreads();
    // every warp: the per-sample tail (lane = sample)
    float hpre[H1], sd[D];
#pragma unroll
    for (int j = 0; j < H1; ++j) hpre[j] = T[j * TS + lane];
#pragma unroll
    for (int d = 0; d < D; ++d) sd[d] = T[(H1 + d) * TS + lane];
    const float q = T[(H1 + D) * TS + lane], lin = T[(H1 + D + 1) * TS + lane];
    float a1[H1];
#pragma unroll
    for (int j = 0; j < H1; ++j) a1[j] = fmaxf(hpre[j], 0.f);
    float h2[H2], dnn = 0.f;
#pragma unroll
    for (int k = 0; k < H2; ++k) {
      float accv = s_b2[k];
#pragma unroll
      for (int j = 0; j < H1; ++j) accv = fmaf(s_w2[k * H1 + j], a1[j], accv);
      h2[k] = fmaxf(accv, 0.f);
      dnn = fmaf(s_w3[k], h2[k], dnn);
    }
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) ss = fmaf(sd[d], sd[d], ss);
    const float z = lin + dnn + 0.5f * (ss - q);
    if (warp == 0 && live && a.logits != nullptr) a.logits[b0 + lane] = z;
    float dz = 0.f, dh1[H1];
    if (BACKWARD) {
      const long long bb = live ? b0 + lane : B - 1;
      const float y = a.labels[bb];
      const float lb = fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));  // BCE with logits
      if (warp == 0 && live) loss_acc += lb;
      const float pr = 1.f / (1.f + expf(-z));
      dz = live ? (pr - y) / (float)B : 0.f;
      float dh2[H2];
#pragma unroll
      for (int k = 0; k < H2; ++k) dh2[k] = h2[k] > 0.f ? dz * s_w3[k] : 0.f;
#pragma unroll
      for (int j = 0; j < H1; ++j) {
        float accv = 0.f;
#pragma unroll
        for (int k = 0; k < H2; ++k) accv = fmaf(s_w2[k * H1 + j], dh2[k], accv);
        dh1[j] = hpre[j] > 0.f ? accv : 0.f;
      }
      if (warp == 0) {  // backward state for B2 (T's region is not overwritten: Mst lies behind it)
        float4* m4 = reinterpret_cast<float4*>(Mst + lane * MS);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) m4[qd] = make_float4(dh1[4 * qd], dh1[4 * qd + 1], dh1[4 * qd + 2], dh1[4 * qd + 3]);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) m4[4 + qd] = make_float4(a1[4 * qd], a1[4 * qd + 1], a1[4 * qd + 2], a1[4 * qd + 3]);
        m4[8] = make_float4(dh2[0], dh2[1], dh2[2], dh2[3]);
        m4[9] = make_float4(h2[0], h2[1], h2[2], h2[3]);
        m4[10] = make_float4(dz, 0.f, 0.f, 0.f);
      }
      // ---------------- B1: d loss / d embedding rows of this warp's groups, reduced per unique id ----------------
      const float* xrow = X + lane * XS;
#pragma unroll 1
      for (int gi = 0; gi < KPC / D; ++gi) {
        const int g = warp * (KPC / D) + gi;
        if (g >= G) break;  // warp-uniform
        const int c = g * D;
        const float4 xa = *reinterpret_cast<const float4*>(xrow + c);
        const float4 xb = *reinterpret_cast<const float4*>(xrow + c + 4);
        const float ev[D] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        float x[D + 1];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float4* w4 = reinterpret_cast<const float4*>(W1x + (c + d) * H1);
          float2 dt2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const float4 w = w4[qd];
            dt2 = __ffma2_rn(make_float2(w.x, w.y), make_float2(dh1[4 * qd + 0], dh1[4 * qd + 1]), dt2);
            dt2 = __ffma2_rn(make_float2(w.z, w.w), make_float2(dh1[4 * qd + 2], dh1[4 * qd + 3]), dt2);
          }
          x[d] = fmaf(dz, sd[d] - ev[d], dt2.x + dt2.y);
        }
        x[D] = dz;  // wide row gradient
        // warp-level id dedup: lanes hitting the same row combine pairwise up a tree threaded through the
        // peer mask (pointer doubling): ceil(log2(max multiplicity)) rounds for the whole warp
        const int r = Rt[lane * RS + g];
        const int key = live ? r : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int rank = __popc(peers & ((1u << lane) - 1));
        const int maxn = __reduce_max_sync(0xffffffffu, (unsigned)__popc(peers));
        const unsigned above = peers & ~((2u << lane) - 1);
        int nxt = above ? __ffs(above) - 1 : -1;
        for (int step = 1; step < maxn; step <<= 1) {
          const int src = nxt >= 0 ? nxt : lane;
          const bool take = nxt >= 0 && (rank & (2 * step - 1)) == 0;
#pragma unroll
          for (int e = 0; e <= D; ++e) {
            const float yv = __shfl_sync(0xffffffffu, x[e], src);
            if (take) x[e] += yv;
          }
          const int nn = __shfl_sync(0xffffffffu, nxt, src);
          nxt = nxt >= 0 ? nn : -1;
        }
        if (live && rank == 0) {
          float* od = a.gsum_deep + ((long long)g * B + r) * D;
          atomicAdd(reinterpret_cast<float4*>(od), make_float4(x[0], x[1], x[2], x[3]));
          atomicAdd(reinterpret_cast<float4*>(od + 4), make_float4(x[4], x[5], x[6], x[7]));
          atomicAdd(a.gsum_wide + (long long)g * B + r, x[D]);
        }
      }
      __syncthreads();  // Mst complete
      // ---------------- B2: parameter gradients, accumulated over the CTA's tiles ----------------
      {
        // second column: 256 + t % 64, hidden units 4*(t/64) .. +3 -- warp-uniform, so its dh1 quad is one of the
        // four broadcast loads above (a per-lane quad cost four shared-memory wavefronts per load)
        const int c1 = 256 + (t & 63), jsel = t >> 6;
#pragma unroll 4
        for (int s = 0; s < TS; ++s) {
          const float4* d4 = reinterpret_cast<const float4*>(Mst + s * MS);
          const float4 q0 = d4[0], q1 = d4[1], q2 = d4[2], q3 = d4[3];
          const float dh[H1] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
          const float x0 = X[s * XS + t], x1 = X[s * XS + c1];
          const float4 dq = jsel == 0 ? q0 : (jsel == 1 ? q1 : (jsel == 2 ? q2 : q3));
          const float2 xx0 = make_float2(x0, x0), xx1 = make_float2(x1, x1);
#pragma unroll
          for (int j = 0; j < H1 / 2; ++j) acc0[j] = __ffma2_rn(make_float2(dh[2 * j], dh[2 * j + 1]), xx0, acc0[j]);
          acc1[0] = __ffma2_rn(make_float2(dq.x, dq.y), xx1, acc1[0]);
          acc1[1] = __ffma2_rn(make_float2(dq.z, dq.w), xx1, acc1[1]);
        }
        if (t < H2 * H1 + 2 * H2 + ND) {
          const int o = t;
          for (int s = 0; s < TS; ++s) {
            const float* m = Mst + s * MS;
            float v;
            if (o < H2 * H1) v = m[32 + o / H1] * m[16 + o % H1];                         // dh2[k] * a1[j]
            else if (o < H2 * H1 + H2) v = m[32 + o - H2 * H1];                            // dh2[k]
            else if (o < H2 * H1 + 2 * H2) v = m[40] * m[36 + o - H2 * H1 - H2];           // dz * h2[k]
            else v = m[40] * X[s * XS + MAXG * D + o - H2 * H1 - 2 * H2];                  // dz * dense[e]
            sacc += v;
          }
        }
      }
    }
    __syncthreads();  // X / R / P are rewritten by the next tile's gather
  }
  if (BACKWARD) {
#pragma unroll
    for (int j = 0; j < H1; ++j) {
      const int dst = w1x_src(l, G, t, j);
      if (dst >= 0) atomicAdd(a.grads + dst, (j & 1) ? acc0[j >> 1].y : acc0[j >> 1].x);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int dst = w1x_src(l, G, 256 + (t & 63), (t >> 6) * 4 + j);
      if (dst >= 0) atomicAdd(a.grads + dst, (j & 1) ? acc1[j >> 1].y : acc1[j >> 1].x);
    }
    if (t < H2 * H1 + 2 * H2 + ND) {
      const int o = t;
      int off;
      if (o < H2 * H1) off = l.o_w2 + o;
      else if (o < H2 * H1 + H2) off = l.o_b2 + o - H2 * H1;
      else if (o < H2 * H1 + 2 * H2) off = l.o_w3 + o - H2 * H1 - H2;
      else off = l.o_wd + o - H2 * H1 - 2 * H2;
      atomicAdd(a.grads + off, sacc);
    }
    if (warp == 0) {
      for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_down_sync(0xffffffffu, loss_acc, o);
      if (lane == 0) atomicAdd(a.loss, loss_acc / (float)B);
    }
  }
}

// The step's loss to the host without a copy-engine operation: one thread stores the value into a ring in
// MAPPED pinned host memory (a 4-byte write over PCIe) and bumps the ring cursor.  A cudaMemcpyAsync of the
// scalar between two graph launches costs a copy-engine round trip on the critical path of every step.
__global__ void k_publish_loss(const float* __restrict__ loss, float* ring, int ring_len, unsigned* cursor) {
  const unsigned at = *cursor;
  *reinterpret_cast<volatile float*>(ring + (at % (unsigned)ring_len)) = *loss;
  *cursor = at + 1;
}

int check_args(const b200_deepfm_args_t* a, bool backward) {
  if (!a || a->G < 1 || a->B < 1) { g_msg = "bad shape"; return -1; }
  if (a->G > MAXG) { g_msg = "the tile tower holds at most 38 id groups"; return -1; }
  if (!a->inv || !a->bet_wide || !a->bet_deep || !a->dense || !a->params || !a->scratch) { g_msg = "null input"; return -1; }
  if (backward && (!a->labels || !a->grads || !a->gsum_wide || !a->gsum_deep || !a->loss || !a->n_unique)) {
    g_msg = "null output";
    return -1;
  }
  return 0;
}

// what: 1 = prologue only (k_tile_prep), 2 = tile kernel only, 3 = both
template <bool BACKWARD>
int launch(const b200_deepfm_args_t* args, void* stream, int what = 3) {
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  static bool attr_done[64][2] = {};
  if (dev < 64 && !attr_done[dev][BACKWARD]) {
    cudaFuncSetAttribute(k_tower_tile<BACKWARD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    attr_done[dev][BACKWARD] = true;
  }
  b200_deepfm_args_t a = *args;
  if (!BACKWARD) { a.grads = nullptr; a.gsum_deep = nullptr; a.gsum_wide = nullptr; a.loss = nullptr; }
  float* w1x = a.scratch;  // NCOL*16 floats at the start of the scratch buffer, then the tile counter
  const long long ntile = ((long long)a.B + TS - 1) / TS;
  static const int per_sm = [] { const char* e = getenv("B200_TILE_CTAS"); const int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
  long long grid = (long long)n_sm * per_sm;
  if (grid > ntile) grid = ntile;
  if (what & 1) {
    k_tile_prep<<<dim3(4, BACKWARD ? a.G : 1), 256, 0, st>>>(a, layout(a.G).total, w1x, (int)grid);
    g_launches += 1;
  }
  if (!(what & 2)) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_msg = cudaGetErrorString(e); return -2; }
    return 0;
  }
  k_tower_tile<BACKWARD><<<(unsigned)grid, THREADS, SMEM_BYTES, st>>>(*args, w1x);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_msg = cudaGetErrorString(e); return -2; }
  return 0;
}

}  // namespace

extern "C" {

int64_t b200_deepfm_tile_launch_count(void) { return g_launches; }
const char* b200_deepfm_tile_last_error(void) { return g_msg.c_str(); }

int b200_deepfm_fwd_bwd_tile(const b200_deepfm_args_t* args, void* stream) {
  if (check_args(args, true)) return -1;
  return launch<true>(args, stream);
}

int b200_deepfm_tile_prologue(const b200_deepfm_args_t* args, void* stream) {
  if (check_args(args, true)) return -1;
  return launch<true>(args, stream, 1);
}

int b200_deepfm_tile_main(const b200_deepfm_args_t* args, void* stream) {
  if (check_args(args, true)) return -1;
  return launch<true>(args, stream, 2);
}

int b200_deepfm_publish_loss(const float* loss_dev, float* ring_host_mapped, int ring_len, unsigned* cursor_dev, void* stream) {
  if (!loss_dev || !ring_host_mapped || ring_len < 1 || !cursor_dev) { g_msg = "bad publish_loss arguments"; return -1; }
  k_publish_loss<<<1, 1, 0, (cudaStream_t)stream>>>(loss_dev, ring_host_mapped, ring_len, cursor_dev);
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_msg = cudaGetErrorString(e); return -2; }
  return 0;
}

int b200_deepfm_forward_tile(const b200_deepfm_args_t* args, void* stream) {
  if (check_args(args, false) || !args->logits) return -1;
  return launch<false>(args, stream);
}

}  // extern "C"
