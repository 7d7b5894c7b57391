// Flat-scheduled row kernels (round 2): pull / set / push for every direct-indexed table and
// dense-parameter row view, any dim, up to kMaxSegs (table, ids) segments in ONE launch.
//
// Why they replace the per-segment grids of ps_kernels.cuh (k_rows_copy_d8 / k_push_rows_d8 / _d1):
//   * a DeepFM step addresses 76 tables whose live sizes (device-side unique counts) range from 3 to
//     25 000 rows.  blockIdx.y = segment gave the big segments 3+ grid-stride passes of dependent
//     latencies while the blocks of the small ones exited at once.  Here the lane-items of all
//     segments form ONE index space (prefix over the live counts, built per block in shared memory
//     from the device-side counts) that a persistent grid walks in whole-warp runs, so every SM
//     carries the same share whatever the split;
//   * everything a row access depends on besides the id -- the table's per-shard base pointers,
//     bitmap pointers, strides, slot offsets, this push's lr / Adam alpha per shard -- is staged in
//     shared memory by the block prologue.  The per-row dependent chain is id -> record -> store
//     (it was count -> id -> TableView -> runtime scalars -> record -> store);
//   * each thread keeps U independent rows in flight: all id loads, then all record / gradient
//     loads, then the arithmetic and the stores (Little's law: 6.5 TB/s x ~1 us needs ~44 KB in
//     flight per SM; one 16 B load per thread at 50 % occupancy was 16 KB);
//   * the update is "lane = (row, column)": the 8 lanes of a dim-8 row read 32 contiguous bytes of
//     param, of each slot and of the gradient -- one full sector per array per row, no shuffles --
//     and every lane executes exactly one element update (the d8 kernel ran four on 2 of 8 lanes).
// Arithmetic is opt_update<> of ps_kernels.cuh: results stay bit-identical to the oracle.
// Replaces go/pkg/kernel/kernel.go:35-199 (Sparse* / Indexed* row loops) and
// go/pkg/common/embedding_table.go:61-77 (Get / SetEmbeddingVectors).
#pragma once
#include "ps_kernels.cuh"

namespace b200ps_impl {

struct FlatMeta {                 // host-side knowledge about the segments of one launch
  unsigned char vec[kMaxSegs];    // copy kernels: 1 = dim % 4 == 0 and rows_dev 16 B aligned -> 16 B chunks
};

struct FlatShared {
  long long prefix[kMaxSegs + 1];  // lane-items before segment s (prefix[nseg] = total)
  float* base[kMaxSegs][kMaxShards];
  uint32_t* present[kMaxSegs][kMaxShards];
  long long rows_cap[kMaxSegs];
  long long stride[kMaxSegs];
  long long soff[kMaxSegs][kMaxSlots + 1];
  int lpr[kMaxSegs];    // lanes per row
  int shift[kMaxSegs];  // log2(lpr) or -1
  int dim[kMaxSegs];
  int owner[kMaxSegs];
  float lr[kMaxShards], alpha[kMaxShards], l2adj[kMaxShards];
};

__device__ __forceinline__ float ld_f1(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_f1(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}

// Largest s < nseg with prefix[s] <= w (empty segments repeat their prefix and are skipped).
__device__ __forceinline__ int flat_seg_of(const long long* prefix, int nseg, long long w) {
  int lo = 0, hi = nseg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= w) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool PUSH>
__device__ __forceinline__ void flat_prologue(const GroupView& gv, const SegBatch& sb, const FlatMeta* fm, FlatShared& fs) {
  const int nseg = sb.nseg, ns = gv.n_shards;
  for (int s = threadIdx.x; s < nseg; s += blockDim.x) {
    const b200ps_seg_t& sg = sb.seg[s];
    const TableView& tv = gv.tables[sg.table];
    const int dim = tv.dim;
    const int lpr = PUSH ? dim : (fm->vec[s] ? dim / 4 : dim);
    fs.lpr[s] = lpr;
    fs.shift[s] = (lpr & (lpr - 1)) == 0 ? 31 - __clz(lpr) : -1;
    fs.dim[s] = dim;
    fs.owner[s] = tv.owner;
    fs.rows_cap[s] = tv.rows;
    fs.stride[s] = tv.row_stride;
#pragma unroll
    for (int k = 0; k <= kMaxSlots; ++k) fs.soff[s][k] = tv.slot_off[k];
    fs.prefix[s + 1] = (long long)seg_count(sg) * lpr;  // count for now, scanned below
  }
  for (int i = threadIdx.x; i < nseg * ns; i += blockDim.x) {
    const int s = i / ns, sh = i - s * ns;
    const TableView& tv = gv.tables[sb.seg[s].table];
    fs.base[s][sh] = tv.base[sh];
    fs.present[s][sh] = tv.present[sh];
  }
  if (PUSH && threadIdx.x < ns) {
    fs.lr[threadIdx.x] = gv.rt->lr[threadIdx.x];
    fs.alpha[threadIdx.x] = gv.rt->alpha[threadIdx.x];
    fs.l2adj[threadIdx.x] = gv.rt->l2adj[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // exclusive scan of <= 96 counts: three per lane
    constexpr int PER = (kMaxSegs + 31) / 32;
    const int lane = threadIdx.x;
    long long v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int idx = lane * PER + j;
      v[j] = idx < nseg ? fs.prefix[idx + 1] : 0;
      sum += v[j];
    }
    long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    long long run = incl - sum;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int idx = lane * PER + j;
      run += v[j];
      if (idx < nseg) fs.prefix[idx + 1] = run;
    }
    if (lane == 0) fs.prefix[0] = 0;
  }
  __syncthreads();
}

struct FlatLoc {
  float* rec;
  long long slot;
  int shard;
  bool ok;
};

__device__ __forceinline__ FlatLoc flat_locate(const GroupView& gv, const FlatShared& fs, int seg, long long id) {
  FlatLoc r;
  const bool nonneg = id >= 0;
  const long long uid = nonneg ? id : 0;
  if (fs.owner[seg] >= 0) {
    r.shard = fs.owner[seg];
    r.slot = uid;
  } else if (gv.shard_shift >= 0) {
    r.shard = (int)(uid & (long long)(gv.n_shards - 1));
    r.slot = uid >> gv.shard_shift;
  } else {
    r.slot = uid / gv.n_shards;
    r.shard = (int)(uid - r.slot * gv.n_shards);
  }
  r.ok = nonneg && r.slot < fs.rows_cap[seg];
  if (!r.ok) r.slot = 0;
  r.rec = fs.base[seg][r.shard] + r.slot * fs.stride[seg];
  return r;
}

__device__ __forceinline__ void flat_mark_present(const FlatShared& fs, int seg, const FlatLoc& r) {
  uint32_t* bm = fs.present[seg][r.shard];
  if (bm == nullptr) return;
  uint32_t* w = bm + (r.slot >> 5);
  const uint32_t bit = 1u << (r.slot & 31);
  if (!(*(volatile uint32_t*)w & bit)) atomicOr_system(w, bit);
}

// item w of the flat space -> (segment, row, column); `seg` is a running hint (items of one thread ascend)
struct FlatItem {
  long long row;
  int seg, col;
  bool live;
};
__device__ __forceinline__ FlatItem flat_item(const FlatShared& fs, long long w, long long total, int& seg) {
  FlatItem it;
  it.live = w < total;
  if (it.live)
    while (w >= fs.prefix[seg + 1]) ++seg;
  it.seg = seg;
  const long long local = it.live ? w - fs.prefix[seg] : 0;
  const int sh = fs.shift[seg];
  it.row = sh >= 0 ? local >> sh : local / fs.lpr[seg];
  it.col = (int)(local - it.row * fs.lpr[seg]);
  return it;
}

// ---------------------------------------------------------------------------
// pull (WRITE = false, PullEmbeddingVectors) / set (WRITE = true, SetEmbeddingVectors / slot access):
// item = (row, 16 B chunk) when the segment is vectorisable, else (row, float).
// ---------------------------------------------------------------------------
template <bool WRITE, int U>
__global__ void __launch_bounds__(256) k_copy_flat(GroupView gv, SegBatch sb, FlatMeta fm, int slot) {
  __shared__ FlatShared fs;
  flat_prologue<false>(gv, sb, &fm, fs);
  const int nseg = sb.nseg;
  const long long total = fs.prefix[nseg];
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wb = warp * (32 * U); wb < total; wb += nwarp * (32 * U)) {
    const long long w0 = wb + lane;
    int seg = flat_seg_of(fs.prefix, nseg, w0 < total ? w0 : total - 1);
    FlatItem it[U];
    long long id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      it[k] = flat_item(fs, wb + k * 32 + lane, total, seg);
      id[k] = it[k].live ? sb.seg[it[k].seg].ids_dev[it[k].row] : 0;
    }
    FlatLoc loc[U];
    float4 x[U];
    float* dst[U];
    bool vec[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      dst[k] = nullptr;
      if (!it[k].live) continue;
      const int s = it[k].seg;
      loc[k] = flat_locate(gv, fs, s, id[k]);
      if (!loc[k].ok) {
        if (it[k].col == 0) atomicOr(gv.err, kErrRange);
        continue;
      }
      vec[k] = fs.lpr[s] != fs.dim[s];
      const int off = vec[k] ? 4 * it[k].col : it[k].col;
      float* rec = loc[k].rec + fs.soff[s][slot] + off;
      float* user = sb.seg[s].rows_dev + it[k].row * fs.dim[s] + off;
      const float* src = WRITE ? user : rec;
      dst[k] = WRITE ? rec : user;
      if (vec[k]) x[k] = ld_f4(src);
      else x[k].x = ld_f1(src);
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (dst[k] != nullptr && it[k].col == 0) flat_mark_present(fs, it[k].seg, loc[k]);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (dst[k] == nullptr) continue;
      if (vec[k]) st_f4(dst[k], x[k]);
      else st_f1(dst[k], x[k].x);
    }
  }
}

// ---------------------------------------------------------------------------
// push: lane = (row, column).  dim-1 tables whose record is one float4 [p, s0, s1, s2] keep the
// single 16 B access.
// ---------------------------------------------------------------------------
template <int OPT, int U>
__global__ void __launch_bounds__(256) k_push_flat(GroupView gv, SegBatch sb, OptParams o) {
  constexpr int S = opt_slots(OPT);
  __shared__ FlatShared fs;
  flat_prologue<true>(gv, sb, nullptr, fs);
  const int nseg = sb.nseg;
  const long long total = fs.prefix[nseg];
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long wb = warp * (32 * U); wb < total; wb += nwarp * (32 * U)) {
    const long long w0 = wb + lane;
    int seg = flat_seg_of(fs.prefix, nseg, w0 < total ? w0 : total - 1);
    FlatItem it[U];
    long long id[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      it[k] = flat_item(fs, wb + k * 32 + lane, total, seg);
      id[k] = it[k].live ? sb.seg[it[k].seg].ids_dev[it[k].row] : 0;
    }
    FlatLoc loc[U];
    float g[U], p[U], s0[U], s1[U], s2[U];
    float* rp[U];
    bool rec4[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      rp[k] = nullptr;
      s0[k] = s1[k] = s2[k] = 0.f;
      if (!it[k].live) continue;
      const int s = it[k].seg;
      loc[k] = flat_locate(gv, fs, s, id[k]);
      if (!loc[k].ok) {
        if (it[k].col == 0) atomicOr(gv.err, kErrRange);
        continue;
      }
      const int dim = fs.dim[s];
      g[k] = ld_f1(sb.seg[s].rows_dev + it[k].row * dim + it[k].col);
      rp[k] = loc[k].rec + it[k].col;
      rec4[k] = S > 0 && dim == 1 && fs.stride[s] == 4 && fs.soff[s][1] == 1;
      if (rec4[k]) {
        const float4 r = ld_f4(rp[k]);
        p[k] = r.x; s0[k] = r.y; s1[k] = r.z; s2[k] = r.w;
      } else {
        p[k] = ld_f1(rp[k]);
        if (S > 0) s0[k] = ld_f1(rp[k] + fs.soff[s][1]);
        if (S > 1) s1[k] = ld_f1(rp[k] + fs.soff[s][2]);
        if (S > 2) s2[k] = ld_f1(rp[k] + fs.soff[s][3]);
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k)
      if (rp[k] != nullptr && it[k].col == 0) flat_mark_present(fs, it[k].seg, loc[k]);
#pragma unroll
    for (int k = 0; k < U; ++k) {
      if (rp[k] == nullptr) continue;
      const int s = it[k].seg, sh = loc[k].shard;
      opt_update<OPT>(g[k], p[k], s0[k], s1[k], s2[k], fs.lr[sh], fs.alpha[sh], fs.l2adj[sh], o);
      if (rec4[k]) {
        st_f4(rp[k], make_float4(p[k], s0[k], s1[k], s2[k]));
      } else {
        st_f1(rp[k], p[k]);
        if (S > 0) st_f1(rp[k] + fs.soff[s][1], s0[k]);
        if (S > 1) st_f1(rp[k] + fs.soff[s][2], s1[k]);
        if (S > 2) st_f1(rp[k] + fs.soff[s][3], s2[k]);
      }
    }
  }
}

}  // namespace b200ps_impl
