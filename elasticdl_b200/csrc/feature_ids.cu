// Feature-id generation on the GPU (include/b200_features.h): Hashing / Discretization /
// ConcatenateWithOffset / Normalizer of elasticdl_preprocessing/layers, fused into one launch that
// writes the [G][B] id matrix b200ps_unique consumes -- the worker's feature transform
// (model_zoo/dac_ctr/feature_transform.py:36-118) no longer runs on the host.
//
// HBM-bound integer / byte work: one thread per (group, sample); inputs are feature-major so a warp
// reads 32 consecutive values of one feature (coalesced), ids are written coalesced.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/b200_features.h"

namespace {

thread_local std::string g_err;
long long g_launches = 0;
int fail(int code, const std::string& m) {
  g_err = m;
  return code;
}

constexpr int kMaxGroups = 64, kMaxDense = 32, kMaxBoundaries = 256, kMaxW = 64;

// ---------------------------------------------------------------------------
// FarmHash Fingerprint64 = farmhashna::Hash64 (FarmHash 1.1, public domain / MIT), lengths 0..64.
// TensorFlow's to_hash_bucket_fast is Fingerprint64(s) % num_buckets (unsigned).  Restated from the
// published algorithm; pinned by elasticdl_preprocessing/layers/hashing.py:35-39.
// ---------------------------------------------------------------------------
constexpr uint64_t k0 = 0xc3a5c85c97cb3127ULL, k1 = 0xb492b66fbe98f273ULL, k2 = 0x9ae16a3b2f90404fULL;

__host__ __device__ inline uint64_t rot(uint64_t v, int s) { return s == 0 ? v : (v >> s) | (v << (64 - s)); }
__host__ __device__ inline uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
__host__ __device__ inline uint64_t fetch64(const unsigned char* p) {
  uint64_t v = 0;
#pragma unroll
  for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
  return v;
}
__host__ __device__ inline uint64_t fetch32(const unsigned char* p) {
  return (uint64_t)p[0] | ((uint64_t)p[1] << 8) | ((uint64_t)p[2] << 16) | ((uint64_t)p[3] << 24);
}
__host__ __device__ inline uint64_t hash_len16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= (a >> 47);
  uint64_t b = (v ^ a) * mul;
  b ^= (b >> 47);
  return b * mul;
}
__host__ __device__ inline uint64_t hash_0to16(const unsigned char* s, int len) {
  if (len >= 8) {
    const uint64_t mul = k2 + (uint64_t)len * 2;
    const uint64_t a = fetch64(s) + k2;
    const uint64_t b = fetch64(s + len - 8);
    const uint64_t c = rot(b, 37) * mul + a;
    const uint64_t d = (rot(a, 25) + b) * mul;
    return hash_len16(c, d, mul);
  }
  if (len >= 4) {
    const uint64_t mul = k2 + (uint64_t)len * 2;
    const uint64_t a = fetch32(s);
    return hash_len16((uint64_t)len + (a << 3), fetch32(s + len - 4), mul);
  }
  if (len > 0) {
    const uint32_t a = s[0], b = s[len >> 1], c = s[len - 1];
    const uint32_t y = a + (b << 8);
    const uint32_t z = (uint32_t)len + (c << 2);
    return shift_mix((uint64_t)y * k2 ^ (uint64_t)z * k0) * k2;
  }
  return k2;
}
__host__ __device__ inline uint64_t hash_17to32(const unsigned char* s, int len) {
  const uint64_t mul = k2 + (uint64_t)len * 2;
  const uint64_t a = fetch64(s) * k1;
  const uint64_t b = fetch64(s + 8);
  const uint64_t c = fetch64(s + len - 8) * mul;
  const uint64_t d = fetch64(s + len - 16) * k2;
  return hash_len16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + k2, 18) + c, mul);
}
__host__ __device__ inline uint64_t hash_33to64(const unsigned char* s, int len) {
  const uint64_t mul = k2 + (uint64_t)len * 2;
  const uint64_t a = fetch64(s) * k2;
  const uint64_t b = fetch64(s + 8);
  const uint64_t c = fetch64(s + len - 8) * mul;
  const uint64_t d = fetch64(s + len - 16) * k2;
  const uint64_t y = rot(a + b, 43) + rot(c, 30) + d;
  const uint64_t z = hash_len16(y, a + rot(b + k2, 18) + c, mul);
  const uint64_t e = fetch64(s + 16) * mul;
  const uint64_t f = fetch64(s + 24);
  const uint64_t g = (y + fetch64(s + len - 32)) * mul;
  const uint64_t h = (z + fetch64(s + len - 24)) * mul;
  return hash_len16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
}
__host__ __device__ inline uint64_t fingerprint64(const unsigned char* s, int len) {
  if (len <= 16) return hash_0to16(s, len);
  if (len <= 32) return hash_17to32(s, len);
  return hash_33to64(s, len);
}

// tf.as_string(int64): plain decimal, '-' for negatives
__device__ inline int as_string(long long v, unsigned char* out) {
  unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
  unsigned char tmp[20];
  int n = 0;
  do {
    tmp[n++] = (unsigned char)('0' + u % 10);
    u /= 10;
  } while (u);
  int len = 0;
  if (v < 0) out[len++] = '-';
  while (n) out[len++] = tmp[--n];
  return len;
}

__device__ inline int load_string(const uint8_t* src, int W, unsigned char* buf) {
  int len = 0;
  bool open = true;
  for (int i = 0; i < W; ++i) {
    const unsigned char c = src[i];
    buf[i] = c;
    if (open && c != 0) len = i + 1; else open = false;  // zero-padded: the string ends at the first NUL
  }
  return len;
}

// math_ops._bucketize: std::upper_bound over the boundaries (bins include their left boundary)
__device__ inline int bucketize(float x, const float* b, int n) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (x < b[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

struct Plan {  // passed by value (< 4 KB): the launch is self-contained and capturable in a CUDA graph
  b200feat_group_t group[kMaxGroups];
  b200feat_dense_t dense[kMaxDense];
  float bnd[kMaxBoundaries];
  int G, n_dense;
};

__global__ void __launch_bounds__(256) k_feature_transform(Plan p, const void* numeric,
                                                           int numeric_is_float, const uint8_t* __restrict__ strings, int W,
                                                           long long B, void* ids_out, int ids32, float* dense_out) {
  __shared__ float s_bnd[kMaxBoundaries];
  const int g = blockIdx.y;
  if (g < p.G) {
    const b200feat_group_t& gr = p.group[g];
    if (gr.kind == B200FEAT_DISCRETIZE) {
      for (int i = threadIdx.x; i < gr.n_boundaries; i += blockDim.x) s_bnd[i] = p.bnd[gr.boundary_off + i];
      __syncthreads();
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
      long long id;
      if (gr.kind == B200FEAT_DISCRETIZE) {
        // int64 inputs compare against float boundaries in float, like TF's BucketizeOp<T>
        const float x = numeric_is_float ? reinterpret_cast<const float*>(numeric)[(long long)gr.column * B + b]
                                         : (float)reinterpret_cast<const long long*>(numeric)[(long long)gr.column * B + b];
        id = bucketize(x, s_bnd, gr.n_boundaries);
      } else {
        unsigned char buf[kMaxW];
        int len;
        if (gr.kind == B200FEAT_HASH_STRING) {
          len = load_string(strings + ((long long)gr.column * B + b) * W, W, buf);
        } else {
          const long long v = numeric_is_float ? (long long)reinterpret_cast<const float*>(numeric)[(long long)gr.column * B + b]
                                               : reinterpret_cast<const long long*>(numeric)[(long long)gr.column * B + b];
          len = as_string(v, buf);
        }
        id = (long long)(fingerprint64(buf, len) % (uint64_t)gr.num_bins);
      }
      id += gr.offset;
      if (ids32) reinterpret_cast<int*>(ids_out)[(long long)g * B + b] = (int)id;
      else reinterpret_cast<long long*>(ids_out)[(long long)g * B + b] = id;
    }
  } else {  // blockIdx.y >= G: the Normalizer columns
    const int j = g - p.G;
    const b200feat_dense_t& d = p.dense[j];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
      const double x = numeric_is_float ? (double)reinterpret_cast<const float*>(numeric)[(long long)d.column * B + b]
                                        : (double)reinterpret_cast<const long long*>(numeric)[(long long)d.column * B + b];
      dense_out[b * p.n_dense + j] = (float)((x - d.subtractor) / d.divisor);
    }
  }
}

__global__ void __launch_bounds__(256) k_fingerprint(const uint8_t* __restrict__ strings, int W, long long n, uint64_t* out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned char buf[kMaxW];
    const int len = load_string(strings + i * W, W, buf);
    out[i] = fingerprint64(buf, len);
  }
}

int grid_x(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

const char* b200feat_last_error(void) { return g_err.c_str(); }
int64_t b200feat_launch_count(void) { return g_launches; }

int b200feat_transform(const b200feat_group_t* groups, int G, const float* boundaries, int n_boundaries,
                       const b200feat_dense_t* dense, int n_dense, const void* numeric_dev, int n_numeric,
                       int numeric_is_float, const uint8_t* strings_dev, int n_string, int W, int64_t B,
                       void* ids_out_dev, int ids32, float* dense_out_dev, void* stream) {
  if (G < 0 || G > kMaxGroups || n_dense < 0 || n_dense > kMaxDense || B < 1) return fail(-1, "bad transform shape");
  if (n_boundaries < 0 || n_boundaries > kMaxBoundaries) return fail(-1, "too many boundaries (max 256)");
  if (G > 0 && (!groups || !ids_out_dev)) return fail(-1, "groups / ids_out is null");
  if (n_dense > 0 && (!dense || !dense_out_dev)) return fail(-1, "dense / dense_out is null");
  if (W < 0 || W > kMaxW) return fail(-1, "strings longer than 64 bytes are not supported (FarmHash restated for lengths 0..64)");
  Plan p;
  memset(&p, 0, sizeof(p));
  p.G = ids_out_dev ? G : 0;
  p.n_dense = dense_out_dev ? n_dense : 0;
  for (int g = 0; g < p.G; ++g) {
    const b200feat_group_t& gr = groups[g];
    if (gr.kind == B200FEAT_DISCRETIZE) {
      if (gr.column < 0 || gr.column >= n_numeric || !numeric_dev) return fail(-1, "discretize: bad numeric column");
      if (gr.n_boundaries < 0 || gr.boundary_off < 0 || gr.boundary_off + gr.n_boundaries > n_boundaries || (gr.n_boundaries && !boundaries))
        return fail(-1, "discretize: boundaries out of range");
    } else if (gr.kind == B200FEAT_HASH_STRING) {
      if (gr.column < 0 || gr.column >= n_string || !strings_dev || W < 1) return fail(-1, "hash: bad string column");
      if (gr.num_bins <= 0) return fail(-1, "`num_bins` cannot be `None` or non-positive values.");  // hashing.py:53-56
    } else if (gr.kind == B200FEAT_HASH_INT) {
      if (gr.column < 0 || gr.column >= n_numeric || !numeric_dev) return fail(-1, "hash: bad numeric column");
      if (gr.num_bins <= 0) return fail(-1, "`num_bins` cannot be `None` or non-positive values.");
    } else {
      return fail(-1, "unknown group kind");
    }
    p.group[g] = gr;
  }
  for (int j = 0; j < p.n_dense; ++j) {
    if (dense[j].column < 0 || dense[j].column >= n_numeric || !numeric_dev) return fail(-1, "normalizer: bad numeric column");
    if (dense[j].divisor == 0.0) return fail(-1, "The divisor cannot be 0");  // normalizer.py build()
    p.dense[j] = dense[j];
  }
  if (p.G + p.n_dense == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_boundaries; ++i) p.bnd[i] = boundaries[i];
  int gx = grid_x(B);
  const int per = (148 * 8 + p.G + p.n_dense - 1) / (p.G + p.n_dense);
  if (gx > per) gx = per < 1 ? 1 : per;
  dim3 grid(gx, p.G + p.n_dense);
  k_feature_transform<<<grid, 256, 0, st>>>(p, numeric_dev, numeric_is_float, strings_dev, W, (long long)B, ids_out_dev,
                                            ids32, dense_out_dev);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-2, cudaGetErrorString(e));
  return 0;
}

int b200feat_hash_strings(const uint8_t* strings_dev, int W, int64_t n, int64_t num_bins, int64_t* out_dev, void* stream) {
  b200feat_group_t g{};
  g.kind = B200FEAT_HASH_STRING;
  g.num_bins = num_bins;
  return b200feat_transform(&g, 1, nullptr, 0, nullptr, 0, nullptr, 0, 0, strings_dev, 1, W, n, out_dev, 0, nullptr, stream);
}

int b200feat_hash_ints(const int64_t* values_dev, int64_t n, int64_t num_bins, int64_t* out_dev, void* stream) {
  b200feat_group_t g{};
  g.kind = B200FEAT_HASH_INT;
  g.num_bins = num_bins;
  return b200feat_transform(&g, 1, nullptr, 0, nullptr, 0, values_dev, 1, 0, nullptr, 0, 0, n, out_dev, 0, nullptr, stream);
}

int b200feat_bucketize(const float* x_dev, int64_t n, const float* boundaries, int n_boundaries, int64_t* out_dev, void* stream) {
  b200feat_group_t g{};
  g.kind = B200FEAT_DISCRETIZE;
  g.n_boundaries = n_boundaries;
  return b200feat_transform(&g, 1, boundaries, n_boundaries, nullptr, 0, x_dev, 1, 1, nullptr, 0, 0, n, out_dev, 0, nullptr, stream);
}

int b200feat_fingerprint64(const uint8_t* strings_dev, int W, int64_t n, uint64_t* out_dev, void* stream) {
  if (!strings_dev || !out_dev || n < 1 || W < 1 || W > kMaxW) return fail(-1, "bad fingerprint arguments");
  k_fingerprint<<<grid_x(n), 256, 0, (cudaStream_t)stream>>>(strings_dev, W, (long long)n, out_dev);
  g_launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-2, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
