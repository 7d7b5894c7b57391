// Owner-computes exchange for rank-per-GPU groups: the all-to-all-v the reference performs with 2*M
// RPCs per step (python/worker/ps_client.py:105-130 pull, :243-277 push), as NVLink traffic that is
// contiguous in both directions.
//
// Measured on this box (tools/mgpu_probe.py): scattered 32 B accesses to a peer GPU -- reads AND
// writes -- run at ~130-160 GB/s (a few G sectors/s, independent of request size), coalesced peer
// traffic at the link rate.  So nothing scattered crosses NVLink:
//   pull  k_x_post      requester copies its unique-id lists into its exchange buffer, posts a flag
//         k_x_serve     every OWNER streams every requester's id lists over NVLink (coalesced
//                       reads), keeps the ids it owns (id % N), gathers those rows from its own HBM
//                       and appends {row, dst} CONTIGUOUSLY to the requester's response region; it
//                       remembers (id, group) of what it served, in order
//         k_x_unscatter requester copies the rows of each owner's region to bet[dst]
//   push  k_x_send_upd  requester walks each owner's response region again (its order == the owner's
//                       serve order) and writes the matching gradient rows CONTIGUOUSLY into the
//                       owner's inbox: no bucketing, no atomics
//         k_x_apply     the owner applies the fused optimizer to its own shard: entry i of source s
//                       updates the row it served as entry i (with the source's lr / Adam alpha)
// A push therefore rides on the routing of the pull that precedes it (the training step's order).
// Cross-GPU ordering: epoch flags in HBM written with st.release.sys after a per-block system fence
// (last-block-done), read with ld.acquire.sys; everything a peer wrote or owns is read with
// ld.global.cg (no stale L1 lines).  All ranks launch the same kernels in the same order
// (bulk-synchronous step); waits time out after ~2 s and raise instead of hanging the GPU.
#pragma once
#include "ps_kernels.cuh"

namespace b200ps_impl {

constexpr int kXEntry = 48;    // bytes: response {float deep[8]; float wide; int32 dst; pad 2}; update {float g[8]; float gw; pad 3}
constexpr int kXServed = 16;   // bytes: {int64 id; int32 grp; int32 pad}   (owner-local)
constexpr int kXChunk = 512;   // ids scanned per block iteration in k_x_serve (256 threads x 2)

struct XHeader {  // first page of every rank's exchange buffer
  // written by source ranks into the OWNER's header
  int nuniq[kMaxShards][kMaxSegs];  // [src][g] live id count per group of src's request
  int flag_ids[kMaxShards];  // == epoch when src's id lists are published
  int flag_upd[kMaxShards];  // == epoch when src's gradient rows have landed
  float upd_lr[kMaxShards], upd_alpha[kMaxShards], upd_l2adj[kMaxShards];
  // written by owner ranks into the REQUESTER's header
  int flag_resp[kMaxShards];     // [owner] == epoch when that owner's rows have landed
  int resp_cnt[kMaxShards];      // [owner] rows that owner returned
  int flag_applied[kMaxShards];  // [owner] == epoch when that owner has applied my updates
  // local
  int epoch;
  int cursor[kMaxShards];      // owner: rows served so far to each source (this epoch)
  int served_cnt[kMaxShards];  // owner: rows served to each source by the last pull
  unsigned done_blocks;
  unsigned done_src[kMaxShards];
};

struct XView {
  char* buf[kMaxShards];  // every rank's exchange buffer (peer-mapped); buf[me] is local
  long long off_ids, off_resp, off_upd, off_served;  // byte offsets of the regions
  long long cap;                                     // entries per (owner, source) lane = G * B
  const int* deep_tab;                               // [G] table ids (device)
  const int* wide_tab;                               // [G]
  int n, me, G, B;
};

__device__ __forceinline__ XHeader* xhdr(const XView& x, int r) { return reinterpret_cast<XHeader*>(x.buf[r]); }
__device__ __forceinline__ long long* xids(const XView& x, int r) { return reinterpret_cast<long long*>(x.buf[r] + x.off_ids); }
__device__ __forceinline__ char* xresp(const XView& x, int requester, int owner) {
  return x.buf[requester] + x.off_resp + ((long long)owner * x.cap) * kXEntry;
}
__device__ __forceinline__ char* xupd(const XView& x, int owner, int src) {
  return x.buf[owner] + x.off_upd + ((long long)src * x.cap) * kXEntry;
}
__device__ __forceinline__ char* xserved(const XView& x, int owner, int src) {
  return x.buf[owner] + x.off_served + ((long long)src * x.cap) * kXServed;
}

constexpr unsigned kErrTimeout = 8u;

// The flags are posted with a system-scope release after the data and read with a system-scope
// ACQUIRE load; what a peer wrote is then read with ld.global.cg (L2 only).  A full
// __threadfence_system() on the reader side costs an L1 invalidate per block and dominated the
// first version of these kernels.
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ldcg_f4(const void* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldcg_i(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldcg_f(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ long long ldcg_ll(const long long* p) {
  long long v;
  asm volatile("ld.global.cg.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void wait_flag(const int* flag, int epoch, unsigned* err) {
  for (long long spins = 0; spins < (1LL << 24); ++spins) {  // ~2 s with the sleep below
    if (ld_acquire_sys(flag) >= epoch) return;
    __nanosleep(64);
  }
  atomicOr(err, kErrTimeout);
}

// Largest g with prefix[g] <= w (prefix[0] == 0, prefix[G] == total > w).
__device__ __forceinline__ int group_of(const int* prefix, int G, int w) {
  int lo = 0, hi = G;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prefix[mid] <= w) lo = mid; else hi = mid;
  }
  return lo;
}

// Requester: publish this step's unique-id lists (the live prefix of every group) and open a new epoch.
__global__ void __launch_bounds__(256) k_x_post(XView x, const long long* uniq, const int* __restrict__ n_unique) {
  XHeader* h = xhdr(x, x.me);
  long long* ids = xids(x, x.me);
  __shared__ int s_prefix[kMaxSegs + 1];
  __shared__ int s_u[kMaxSegs];
  if (threadIdx.x < x.G) {
    const int u = n_unique[threadIdx.x];
    s_u[threadIdx.x] = u < x.B ? u : x.B;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int g = 0; g < x.G; ++g) {
      s_prefix[g] = acc;
      acc += s_u[g];
    }
    s_prefix[x.G] = acc;
  }
  __syncthreads();
  const int total = s_prefix[x.G];
  if (uniq != ids) {  // not published in place (b200ps_xchg_ids): copy the live prefixes
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
      const int g = group_of(s_prefix, x.G, w);
      const long long slot = (long long)g * x.B + (w - s_prefix[g]);
      ids[slot] = uniq[slot];
    }
  }
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&h->done_blocks, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  const int epoch = *(volatile int*)&h->epoch + 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    h->epoch = epoch;
    h->done_blocks = 0;
  }
  if (threadIdx.x < kMaxShards) {
    h->cursor[threadIdx.x] = 0;
    h->done_src[threadIdx.x] = 0;
  }
  for (int i = threadIdx.x; i < x.n * x.G; i += blockDim.x)  // counts travel to the owners with the flag
    xhdr(x, i / x.G)->nuniq[x.me][i % x.G] = s_u[i % x.G];
  __syncthreads();  // the releases below are cumulative over everything this block has synchronised with
  if (threadIdx.x < x.n) st_release_sys(&xhdr(x, threadIdx.x)->flag_ids[x.me], epoch);
}

// Owner: stream source blockIdx.y's id lists, keep the ids this shard owns, gather and return the rows.
__global__ void __launch_bounds__(256) k_x_serve(XView x, GroupView gv) {
  constexpr int PER = kXChunk / 256;
  const int src = blockIdx.y;
  XHeader* h = xhdr(x, x.me);
  __shared__ int s_prefix[kMaxSegs + 1];
  __shared__ int s_u[kMaxSegs];
  __shared__ int s_epoch, s_n, s_base;
  __shared__ long long s_id[kXChunk];
  __shared__ int s_slot[kXChunk];
  if (threadIdx.x == 0) {
    s_epoch = h->epoch;
    wait_flag(&h->flag_ids[src], s_epoch, gv.err);
    s_n = 0;
  }
  __syncthreads();
  if (threadIdx.x < x.G) s_u[threadIdx.x] = ldcg_i(&h->nuniq[src][threadIdx.x]);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int g = 0; g < x.G; ++g) {
      s_prefix[g] = acc;
      acc += s_u[g];
    }
    s_prefix[x.G] = acc;
  }
  __syncthreads();
  const int total = s_prefix[x.G];
  const long long* ids = xids(x, src);  // the source's id lists (remote unless src == me), read coalesced
  char* resp = xresp(x, src, x.me);     // remote, appended contiguously
  char* served = xserved(x, x.me, src);
  const int lane = threadIdx.x & 31, lane4 = threadIdx.x & 3;
  const bool pow2 = (x.n & (x.n - 1)) == 0;
  long long nid[PER];
  int nslot[PER];
  auto fetch = [&](int chunk) {  // issue the (remote) id loads of a chunk; consumed one iteration later
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int w = chunk + j * 256 + threadIdx.x;
      nslot[j] = -1;
      nid[j] = 0;
      if (w < total) {
        const int g = group_of(s_prefix, x.G, w);
        nslot[j] = g * x.B + (w - s_prefix[g]);
        nid[j] = ldcg_ll(ids + nslot[j]);
      }
    }
  };
  int chunk = blockIdx.x * kXChunk;
  if (chunk < total) fetch(chunk);
  for (; chunk < total; chunk += gridDim.x * kXChunk) {
    // pass 1: compact the ids this shard owns into shared memory
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const long long id = nid[j];
      const int slot = nslot[j];
      const int owner = pow2 ? (int)(id & (x.n - 1)) : (int)(((id % x.n) + x.n) % x.n);
      const bool mine = slot >= 0 && owner == x.me;
      const unsigned m = __ballot_sync(0xffffffffu, mine);
      int base = 0;
      if (lane == 0 && m) base = atomicAdd(&s_n, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (mine) {
        const int p = base + __popc(m & ((1u << lane) - 1));
        s_id[p] = id;
        s_slot[p] = slot;
      }
    }
    const int next = chunk + gridDim.x * kXChunk;
    if (next < total) fetch(next);
    __syncthreads();
    const int n = s_n;
    if (threadIdx.x == 0) s_base = n ? atomicAdd(&h->cursor[src], n) : 0;
    __syncthreads();
    const int base = s_base;
    // pass 2: four lanes per kept id -- deep lo, deep hi, {wide, dst}, served record
    for (int e = threadIdx.x >> 2; e < n; e += 64) {
      const long long id = s_id[e];
      const int slot = s_slot[e];
      const int g = slot / x.B;
      char* out = resp + (long long)(base + e) * kXEntry;
      if (lane4 < 2) {
        const TableView& td = gv.tables[x.deep_tab[g]];
        RowLoc loc = locate(gv, td, id);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (loc.ok) {
          v = ld_f4(loc.rec + 4 * lane4);
          if (lane4 == 0) mark_present(td, loc);
        } else if (lane4 == 0) {
          atomicOr(gv.err, kErrRange);
        }
        *reinterpret_cast<float4*>(out + 16 * lane4) = v;
      } else if (lane4 == 2) {
        const TableView& tw = gv.tables[x.wide_tab[g]];
        RowLoc lw = locate(gv, tw, id);
        const float w = lw.ok ? *lw.rec : 0.f;
        if (lw.ok) mark_present(tw, lw);
        *reinterpret_cast<float4*>(out + 32) = make_float4(w, __int_as_float(slot), 0.f, 0.f);
      } else {
        int4 rec;
        rec.x = (int)(id & 0xffffffffLL);
        rec.y = (int)(id >> 32);
        rec.z = g;
        rec.w = 0;
        *reinterpret_cast<int4*>(served + (long long)(base + e) * kXServed) = rec;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
  }
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&h->done_src[src], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const int cnt = *(volatile int*)&h->cursor[src];
    h->served_cnt[src] = cnt;
    XHeader* rh = xhdr(x, src);
    rh->resp_cnt[x.me] = cnt;
    st_release_sys(&rh->flag_resp[x.me], s_epoch);
    h->done_src[src] = 0;
  }
}

// Requester: rows of owner blockIdx.y have landed in my response region -> bet_d / bet_w.
__global__ void __launch_bounds__(256) k_x_unscatter(XView x, GroupView gv, float* bet_d, float* bet_w) {
  const int owner = blockIdx.y;
  XHeader* h = xhdr(x, x.me);
  __shared__ int s_cnt;
  if (threadIdx.x == 0) {
    wait_flag(&h->flag_resp[owner], h->epoch, gv.err);
    s_cnt = ldcg_i(&h->resp_cnt[owner]);
  }
  __syncthreads();
  const int cnt = s_cnt;
  const char* resp = xresp(x, x.me, owner);
  const int lane4 = threadIdx.x & 3;
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; i < cnt; i += stride) {
    const char* in = resp + i * kXEntry;
    const float4 tail = ldcg_f4(in + 32);
    const int dst = __float_as_int(tail.y);
    if (lane4 < 2) *reinterpret_cast<float4*>(bet_d + (long long)dst * 8 + 4 * lane4) = ldcg_f4(in + 16 * lane4);
    else if (lane4 == 2) bet_w[dst] = tail.x;
  }
}

// Requester: gradient rows to owner blockIdx.y, in the order that owner served my rows.
__global__ void __launch_bounds__(256) k_x_send_upd(XView x, GroupView gv, const float* __restrict__ gsum_d,
                                                    const float* __restrict__ gsum_w) {
  const int owner = blockIdx.y;
  XHeader* h = xhdr(x, x.me);
  const int cnt = ldcg_i(&h->resp_cnt[owner]);
  const char* resp = xresp(x, x.me, owner);
  char* upd = xupd(x, owner, x.me);  // remote, contiguous
  const int lane4 = threadIdx.x & 3;
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; i0 < cnt; i0 += 2 * stride) {
    // two entries in flight per lane group: dst -> gradient row -> contiguous remote store
    const long long i1 = i0 + stride;
    const bool two = i1 < cnt;
    const int d0 = __float_as_int(ldcg_f(reinterpret_cast<const float*>(resp + i0 * kXEntry + 36)));
    const int d1 = two ? __float_as_int(ldcg_f(reinterpret_cast<const float*>(resp + i1 * kXEntry + 36))) : d0;
    float4 v0, v1;
    if (lane4 < 2) {
      v0 = *reinterpret_cast<const float4*>(gsum_d + (long long)d0 * 8 + 4 * lane4);
      v1 = *reinterpret_cast<const float4*>(gsum_d + (long long)d1 * 8 + 4 * lane4);
    } else {
      v0 = make_float4(gsum_w[d0], 0.f, 0.f, 0.f);
      v1 = make_float4(gsum_w[d1], 0.f, 0.f, 0.f);
    }
    if (lane4 < 3) {
      *reinterpret_cast<float4*>(upd + i0 * kXEntry + 16 * lane4) = v0;
      if (two) *reinterpret_cast<float4*>(upd + i1 * kXEntry + 16 * lane4) = v1;
    }
  }
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&h->done_blocks, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < x.n) {
    const int o = threadIdx.x;
    XHeader* oh = xhdr(x, o);
    oh->upd_lr[x.me] = gv.rt->lr[o];
    oh->upd_alpha[x.me] = gv.rt->alpha[o];
    oh->upd_l2adj[x.me] = gv.rt->l2adj[o];
    st_release_sys(&oh->flag_upd[x.me], *(volatile int*)&h->epoch);
  }
  if (threadIdx.x == 0) h->done_blocks = 0;
}

__device__ __forceinline__ long long entry_id(int4 e) {
  return (long long)(((unsigned long long)(unsigned)e.y << 32) | (unsigned)e.x);
}

// Owner: apply the updates of source blockIdx.y to the rows served to it (its own ApplyGradients:
// lr / Adam alpha were fixed by the source's push_begin on this shard).
template <int OPT>
__global__ void __launch_bounds__(256) k_x_apply(XView x, GroupView gv, OptParams o) {
  constexpr int S = opt_slots(OPT);
  const int src = blockIdx.y;
  XHeader* h = xhdr(x, x.me);
  __shared__ int s_epoch, s_cnt;
  __shared__ float s_lr, s_alpha, s_l2;
  if (threadIdx.x == 0) {
    s_epoch = h->epoch;
    wait_flag(&h->flag_upd[src], s_epoch, gv.err);
    s_cnt = h->served_cnt[src];
    s_lr = ldcg_f(&h->upd_lr[src]);
    s_alpha = ldcg_f(&h->upd_alpha[src]);
    s_l2 = ldcg_f(&h->upd_l2adj[src]);
  }
  __syncthreads();
  const int cnt = s_cnt;
  const float lr = s_lr, alpha = s_alpha, l2adj = s_l2;
  const char* upd = xupd(x, x.me, src);
  const char* served = xserved(x, x.me, src);
  const int lane4 = threadIdx.x & 3;  // deep lo, deep hi, wide, idle
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; i < cnt; i += stride) {
    const int4 e = *reinterpret_cast<const int4*>(served + i * kXServed);
    const long long id = entry_id(e);
    const char* in = upd + i * kXEntry;
    if (lane4 < 2) {
      const TableView& td = gv.tables[x.deep_tab[e.z]];
      RowLoc loc = locate(gv, td, id);
      if (!loc.ok) continue;  // already reported by the serving pass
      float* rec = loc.rec + 4 * lane4;
      float4 g = ldcg_f4(in + 16 * lane4), p = ld_f4(rec), s0 = p, s1 = p, s2 = p;
      if (S > 0) s0 = ld_f4(rec + td.slot_off[1]);
      if (S > 1) s1 = ld_f4(rec + td.slot_off[2]);
      if (S > 2) s2 = ld_f4(rec + td.slot_off[3]);
      float* gf = reinterpret_cast<float*>(&g);
      float* pf = reinterpret_cast<float*>(&p);
      float* af = reinterpret_cast<float*>(&s0);
      float* bf = reinterpret_cast<float*>(&s1);
      float* cf = reinterpret_cast<float*>(&s2);
#pragma unroll
      for (int k = 0; k < 4; ++k) opt_update<OPT>(gf[k], pf[k], af[k], bf[k], cf[k], lr, alpha, l2adj, o);
      st_f4(rec, p);
      if (S > 0) st_f4(rec + td.slot_off[1], s0);
      if (S > 1) st_f4(rec + td.slot_off[2], s1);
      if (S > 2) st_f4(rec + td.slot_off[3], s2);
    } else if (lane4 == 2) {
      const TableView& tw = gv.tables[x.wide_tab[e.z]];
      RowLoc lw = locate(gv, tw, id);
      if (!lw.ok) continue;
      float* rec = lw.rec;
      float g = ldcg_f(reinterpret_cast<const float*>(in + 32)), p = *rec, s0 = 0.f, s1 = 0.f, s2 = 0.f;
      if (S > 0) s0 = rec[tw.slot_off[1]];
      if (S > 1) s1 = rec[tw.slot_off[2]];
      if (S > 2) s2 = rec[tw.slot_off[3]];
      opt_update<OPT>(g, p, s0, s1, s2, lr, alpha, l2adj, o);
      *rec = p;
      if (S > 0) rec[tw.slot_off[1]] = s0;
      if (S > 1) rec[tw.slot_off[2]] = s1;
      if (S > 2) rec[tw.slot_off[3]] = s2;
    }
  }
  __shared__ bool last;
  __threadfence();  // rows are local; the flag's system-scope release orders them for the source's next pull
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&h->done_src[src], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    st_release_sys(&xhdr(x, src)->flag_applied[x.me], s_epoch);
    h->done_src[src] = 0;
  }
}

// Requester: all owners have applied my updates (my next pull observes my own push).
__global__ void k_x_wait_applied(XView x, GroupView gv) {
  XHeader* h = xhdr(x, x.me);
  if (threadIdx.x < x.n) wait_flag(&h->flag_applied[threadIdx.x], h->epoch, gv.err);
}

}  // namespace b200ps_impl
