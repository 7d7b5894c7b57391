// Owner-computes exchange for rank-per-GPU groups ("all-to-all-v" of the reference's per-shard RPC
// fan-out, python/worker/ps_client.py:105-130,243-277, over NVLink).
//
// Measured on this box (tools/mgpu_probe.py): scattered 32 B accesses to a peer GPU -- reads AND
// writes -- run at ~130-160 GB/s (a few G sectors/s), while coalesced peer traffic runs at the
// link rate.  So nothing scattered crosses NVLink here: a requester buckets its unique ids by owner
// and writes them CONTIGUOUSLY into the owner's inbox; the owner gathers rows from its own HBM and
// writes them back CONTIGUOUSLY in request order; the requester un-scatters locally.  Updates travel
// the same way and the owner applies the optimizer to its own shard.  Cross-GPU ordering uses
// epoch flags in HBM (system-scope fences, last-block-done), all kernels are launched by every rank
// in the same order (bulk-synchronous step), waits carry a timeout so a dead peer cannot hang a GPU.
#pragma once
#include "ps_kernels.cuh"

namespace b200ps_impl {

constexpr int kXEntryReq = 16;   // bytes: {int64 id; int32 dst; int32 grp}
constexpr int kXEntryResp = 48;  // bytes: {float deep[8]; float wide; int32 dst; pad 2}
constexpr int kXEntryUpd = 64;   // bytes: {int64 id; int32 grp; int32 pad; float g_deep[8]; float g_wide; pad 3}

struct XHeader {  // at the start of every rank's exchange buffer
  // owner side, written by the source ranks
  int req_cnt[kMaxShards];
  int upd_cnt[kMaxShards];
  float upd_lr[kMaxShards], upd_alpha[kMaxShards], upd_l2adj[kMaxShards];
  int flag_req[kMaxShards];  // == epoch when src's requests have landed
  int flag_upd[kMaxShards];
  // requester side, written by the owner ranks
  int flag_resp[kMaxShards];     // [owner] == epoch when that owner's rows have landed
  int flag_applied[kMaxShards];  // [owner] == epoch when that owner has applied my updates
  // local scratch (never written remotely)
  int epoch;
  int cursor[kMaxShards];  // bucket cursors of the running send kernel
  int sent[kMaxShards];    // entries sent to each owner by the last request pass
  unsigned done_blocks;
  unsigned done_src[kMaxShards];
  int pad[9];
};

struct XView {
  char* buf[kMaxShards];  // every rank's exchange buffer (peer-mapped); buf[me] is local
  long long off_req, off_resp, off_upd;  // byte offsets of the three regions
  long long cap;                         // entries per (owner, source) lane = G * B
  const int* deep_tab;                   // [G] table ids (device)
  const int* wide_tab;                   // [G]
  int n, me, G, B;
};

__device__ __forceinline__ XHeader* xhdr(const XView& x, int r) { return reinterpret_cast<XHeader*>(x.buf[r]); }
__device__ __forceinline__ char* xreq(const XView& x, int owner, int src) {
  return x.buf[owner] + x.off_req + ((long long)src * x.cap) * kXEntryReq;
}
__device__ __forceinline__ char* xresp(const XView& x, int requester, int owner) {
  return x.buf[requester] + x.off_resp + ((long long)owner * x.cap) * kXEntryResp;
}
__device__ __forceinline__ char* xupd(const XView& x, int owner, int src) {
  return x.buf[owner] + x.off_upd + ((long long)src * x.cap) * kXEntryUpd;
}

constexpr unsigned kErrTimeout = 8u;

// Spin until *flag >= epoch.  The flag is posted with a system-scope release after the data; it is
// read with a system-scope ACQUIRE load, and everything a peer wrote is then read with ld.global.cg
// (L2 only -- peer writes land in this GPU's L2/HBM, a stale L1 line must not be hit).  A full
// __threadfence_system() here costs an L1 invalidate per block and dominated these kernels.
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int4 ldcg_i4(const void* p) {
  int4 v;
  asm volatile("ld.global.cg.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
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

__device__ __forceinline__ void wait_flag(const int* flag, int epoch, unsigned* err) {
  for (long long spins = 0; spins < (1LL << 24); ++spins) {  // ~2 s with the sleep below
    if (ld_acquire_sys(flag) >= epoch) return;
    __nanosleep(64);
  }
  atomicOr(err, kErrTimeout);
}

__global__ void k_x_begin(XView x) {  // one thread: new epoch, clear the send cursors
  XHeader* h = xhdr(x, x.me);
  if (threadIdx.x == 0) {
    h->epoch += 1;
    h->done_blocks = 0;
  }
  if (threadIdx.x < kMaxShards) {
    h->cursor[threadIdx.x] = 0;
    h->done_src[threadIdx.x] = 0;
  }
}

// Requester: bucket (id, dst) by owner into the owners' inboxes.  UPD: also ship the gradient rows.
// Work is indexed over the LIVE (group, rank) pairs only (prefix of n_unique in shared memory); a
// block reserves its output range per owner with ONE global atomic per owner (count pass, then
// write pass), so the eight cursors are not hammered by every warp.
constexpr int kXChunk = 512;  // live entries per block iteration (256 threads x 2): ~240 busy blocks at 122 K entries

template <bool UPD>
__global__ void __launch_bounds__(256) k_x_send(XView x, GroupView gv, const int64_t* __restrict__ uniq,
                                                const int* __restrict__ n_unique, const float* __restrict__ gsum_d,
                                                const float* __restrict__ gsum_w) {
  XHeader* h = xhdr(x, x.me);
  __shared__ int s_prefix[kMaxSegs + 1];
  __shared__ int s_cnt[kMaxShards], s_base[kMaxShards];
  const int lane = threadIdx.x & 31;
  // exclusive prefix of the live counts: G independent loads in parallel, then a short serial scan in
  // shared memory (a serial chain of G global loads would cost ~0.3 us each in every block)
  __shared__ int s_u[kMaxSegs];
  if (threadIdx.x < x.G) {
    const int u = n_unique[threadIdx.x];
    s_u[threadIdx.x] = u < x.B ? u : x.B;
  }
  if (threadIdx.x < kMaxShards) s_cnt[threadIdx.x] = 0;
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
  for (int chunk = blockIdx.x * kXChunk; chunk < total; chunk += gridDim.x * kXChunk) {
    long long id_[kXChunk / 256];
    int slot_[kXChunk / 256];  // g * B + r, -1 = dead
    // pass 1: locate the live entries of this chunk and count them per owner
#pragma unroll
    for (int j = 0; j < kXChunk / 256; ++j) {
      const int w = chunk + j * 256 + threadIdx.x;
      slot_[j] = -1;
      id_[j] = 0;
      int owner = -1 - lane;
      if (w < total) {
        int lo = 0, hi = x.G;  // largest g with prefix[g] <= w
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_prefix[mid] <= w) lo = mid; else hi = mid;
        }
        slot_[j] = lo * x.B + (w - s_prefix[lo]);
        id_[j] = uniq[slot_[j]];
        owner = gv.shard_shift >= 0 ? (int)(id_[j] & (x.n - 1)) : (int)(id_[j] % x.n);
      }
      const unsigned peers = __match_any_sync(0xffffffffu, owner);
      if (owner >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&s_cnt[owner], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < x.n) {
      s_base[threadIdx.x] = atomicAdd(&h->cursor[threadIdx.x], s_cnt[threadIdx.x]);
      s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    // pass 2: write the entries; lanes of a warp with the same owner get consecutive positions
#pragma unroll
    for (int j = 0; j < kXChunk / 256; ++j) {
      const bool live = slot_[j] >= 0;
      const long long id = id_[j];
      int owner = -1 - lane;
      if (live) owner = gv.shard_shift >= 0 ? (int)(id & (x.n - 1)) : (int)(id % x.n);
      const unsigned peers = __match_any_sync(0xffffffffu, owner);
      const int leader = __ffs(peers) - 1;
      int loc = 0;
      if (live && lane == leader) loc = atomicAdd(&s_cnt[owner], __popc(peers));
      loc = __shfl_sync(0xffffffffu, loc, leader);
      if (!live) continue;
      const int pos = s_base[owner] + loc + __popc(peers & ((1u << lane) - 1));
      int4 e;
      e.x = (int)(id & 0xffffffffLL);
      e.y = (int)(id >> 32);
      if (!UPD) {
        e.z = slot_[j];  // dst = g * B + r
        e.w = slot_[j] / x.B;
        *reinterpret_cast<int4*>(xreq(x, owner, x.me) + (long long)pos * kXEntryReq) = e;
      } else {
        e.z = slot_[j] / x.B;
        e.w = 0;
        float4* dst = reinterpret_cast<float4*>(xupd(x, owner, x.me) + (long long)pos * kXEntryUpd);
        const float4* gd = reinterpret_cast<const float4*>(gsum_d + (long long)slot_[j] * 8);
        dst[0] = *reinterpret_cast<float4*>(&e);
        dst[1] = gd[0];
        dst[2] = gd[1];
        dst[3] = make_float4(gsum_w[slot_[j]], 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    if (threadIdx.x < kMaxShards) s_cnt[threadIdx.x] = 0;
    __syncthreads();
  }
  // publish: every block fences its writes; the last block to finish posts the counts and the flags
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&h->done_blocks, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  const int epoch = h->epoch;
  if (threadIdx.x < x.n) {
    const int o = threadIdx.x;
    const int cnt = *(volatile int*)&h->cursor[o];
    XHeader* oh = xhdr(x, o);
    if (!UPD) {
      h->sent[o] = cnt;
      oh->req_cnt[x.me] = cnt;
    } else {
      oh->upd_cnt[x.me] = cnt;
      oh->upd_lr[x.me] = gv.rt->lr[o];
      oh->upd_alpha[x.me] = gv.rt->alpha[o];
      oh->upd_l2adj[x.me] = gv.rt->l2adj[o];
    }
    if (!UPD) st_release_sys(&oh->flag_req[x.me], epoch);
    else st_release_sys(&oh->flag_upd[x.me], epoch);
  }
  if (threadIdx.x == 0) h->done_blocks = 0;
  if (threadIdx.x < kMaxShards) h->cursor[threadIdx.x] = 0;
}

__device__ __forceinline__ long long entry_id(int4 e) {
  return (long long)(((unsigned long long)(unsigned)e.y << 32) | (unsigned)e.x);
}

// Owner: serve the requests of source rank blockIdx.y from the local shard, rows back in request order.
__global__ void __launch_bounds__(256) k_x_serve(XView x, GroupView gv) {
  const int src = blockIdx.y;
  XHeader* h = xhdr(x, x.me);
  __shared__ int s_epoch, s_cnt;
  if (threadIdx.x == 0) {
    s_epoch = h->epoch;
    wait_flag(&h->flag_req[src], s_epoch, gv.err);
    s_cnt = ldcg_i(&h->req_cnt[src]);
  }
  __syncthreads();
  const int cnt = s_cnt;
  const char* req = xreq(x, x.me, src);
  char* resp = xresp(x, src, x.me);  // remote (or local when src == me), contiguous
  const int lane4 = threadIdx.x & 3;  // 4 lanes per entry: deep lo, deep hi, {wide, dst}, idle
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; i < cnt; i += stride) {
    const int4 e = ldcg_i4(req + i * kXEntryReq);
    const long long id = entry_id(e);
    const TableView& td = gv.tables[x.deep_tab[e.w]];
    RowLoc loc = locate(gv, td, id);
    if (!loc.ok) {
      if (lane4 == 0) atomicOr(gv.err, kErrRange);
      continue;
    }
    float4* out = reinterpret_cast<float4*>(resp + i * kXEntryResp);
    if (lane4 < 2) {
      out[lane4] = ld_f4(loc.rec + 4 * lane4);
      if (lane4 == 0) mark_present(td, loc);
    } else if (lane4 == 2) {
      const TableView& tw = gv.tables[x.wide_tab[e.w]];
      RowLoc lw = locate(gv, tw, id);
      const float w = lw.ok ? *lw.rec : 0.f;
      if (lw.ok) mark_present(tw, lw);
      out[2] = make_float4(w, __int_as_float(e.z), 0.f, 0.f);
    }
  }
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&h->done_src[src], 1u) == gridDim.x - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    st_release_sys(&xhdr(x, src)->flag_resp[x.me], s_epoch);
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
    s_cnt = h->sent[owner];
  }
  __syncthreads();
  const int cnt = s_cnt;
  const char* resp = xresp(x, x.me, owner);
  const int lane4 = threadIdx.x & 3;
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; i < cnt; i += stride) {
    const char* in = resp + i * kXEntryResp;
    const float4 tail = ldcg_f4(in + 32);
    const int dst = __float_as_int(tail.y);
    if (lane4 < 2) *reinterpret_cast<float4*>(bet_d + (long long)dst * 8 + 4 * lane4) = ldcg_f4(in + 16 * lane4);
    else if (lane4 == 2) bet_w[dst] = tail.x;
  }
}

// Owner: apply the updates of source rank blockIdx.y to the local shard (its own ApplyGradients:
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
    s_cnt = ldcg_i(&h->upd_cnt[src]);
    s_lr = ldcg_f(&h->upd_lr[src]);
    s_alpha = ldcg_f(&h->upd_alpha[src]);
    s_l2 = ldcg_f(&h->upd_l2adj[src]);
  }
  __syncthreads();
  const int cnt = s_cnt;
  const float lr = s_lr, alpha = s_alpha, l2adj = s_l2;
  const char* upd = xupd(x, x.me, src);
  const int lane4 = threadIdx.x & 3;  // deep lo, deep hi, wide, idle
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; i < cnt; i += stride) {
    const char* in = upd + i * kXEntryUpd;
    const int4 e = ldcg_i4(in);
    const long long id = entry_id(e);
    const TableView& td = gv.tables[x.deep_tab[e.z]];
    RowLoc loc = locate(gv, td, id);
    if (!loc.ok) {
      if (lane4 == 0) atomicOr(gv.err, kErrRange);
      continue;
    }
    if (lane4 < 2) {
      float* rec = loc.rec + 4 * lane4;
      float4 g = ldcg_f4(in + 16 + 16 * lane4), p = ld_f4(rec), s0 = p, s1 = p, s2 = p;
      if (S > 0) s0 = ld_f4(rec + td.slot_off[1]);
      if (S > 1) s1 = ld_f4(rec + td.slot_off[2]);
      if (S > 2) s2 = ld_f4(rec + td.slot_off[3]);
      if (lane4 == 0) mark_present(td, loc);
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
      float g = ldcg_f(reinterpret_cast<const float*>(in + 48)), p = *rec, s0 = 0.f, s1 = 0.f, s2 = 0.f;
      if (S > 0) s0 = rec[tw.slot_off[1]];
      if (S > 1) s1 = rec[tw.slot_off[2]];
      if (S > 2) s2 = rec[tw.slot_off[3]];
      mark_present(tw, lw);
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
