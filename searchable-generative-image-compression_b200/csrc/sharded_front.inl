// Single-process multi-GPU index (include/sgic.h: sgic_index_create_sharded).
//
// The reference's caller is ONE process (src/search.py:149-162, webapp.py:246-248), so the row-sharded index must
// also exist behind one handle: a "front" sgic_index that owns one ordinary single-device index per GPU.  Rows are
// appended in contiguous runs ("segments": global start, local start, count), queries are replicated, every GPU
// scans its shard on its own stream, and the shards' final writers (K3's last CTA, the merge kernels of the
// tensor-core path) store their (nq, k) answer — already carrying global row numbers — STRAIGHT INTO THE HOME
// GPU's gather buffer over NVLink (cudaDeviceEnablePeerAccess, plain stores; no push kernel, no IPC, no NCCL).
// The home stream waits for one event per shard and runs the K5 merge (score desc, global id asc: identical to the
// single-GPU answer).  Host buffers are staged through pinned memory here, in C; a search costs the caller one call.
//
// This file is included at the end of sgic_api.cu (it uses the single-device internals directly).

namespace sgic {

// ---- a small pool: one worker per shard beyond the first, so that the G launches do not queue up behind one
// host thread (each launch sequence costs ~5 us; at 8 GPUs the last shard would start ~40 us late) -------------
class ShardPool {
 public:
  // devices[g]: the GPU worker g launches on; made current once, so the DeviceGuards inside the jobs are no-ops
  ShardPool(int n, const std::vector<int>& devices) : n_(n) {
    for (int g = 1; g < n_; ++g)
      threads_.emplace_back([this, g, dev = devices[static_cast<size_t>(g)]] {
        cudaSetDevice(dev);
        loop(g);
      });
  }
  ~ShardPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      epoch_.fetch_add(1, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  // fn(g) for every g in [0, n): g = 0 on the caller, the rest on the workers.  Returns when all are done.
  void run(const std::function<void(int)>& fn) {
    fn_ = &fn;
    pending_.store(n_ - 1, std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);   // a worker between its last spin and its wait() must not miss the epoch
      epoch_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    fn(0);
    while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
  }

 private:
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  void loop(int g) {
    uint64_t seen = 0;
    for (;;) {
      // spin for a while after the last job (searches arrive back to back), then sleep on the condition variable
      const auto t0 = std::chrono::steady_clock::now();
      uint64_t e;
      int spins = 0;
      while ((e = epoch_.load(std::memory_order_acquire)) == seen) {
        cpu_relax();
        if ((++spins & 1023) == 0 &&
            std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(500)) {
          std::unique_lock<std::mutex> lk(mu_);
          sleepers_.fetch_add(1, std::memory_order_release);
          cv_.wait(lk, [&] { return epoch_.load(std::memory_order_acquire) != seen; });
          sleepers_.fetch_sub(1, std::memory_order_release);
        }
      }
      seen = e;
      if (stop_) return;
      (*fn_)(g);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  int n_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::atomic<uint64_t> epoch_{0};
  std::atomic<int> pending_{0}, sleepers_{0};
  const std::function<void(int)>* fn_ = nullptr;
  bool stop_ = false;
};

struct FrontState {
  std::vector<std::vector<ShardSeg>> segs;       // per shard, ascending in both coordinates
  std::vector<uint8_t> direct;                   // shard g can store into the home GPU's memory
  std::vector<cudaEvent_t> ev_done;              // per shard (created on its device): answer has landed at home
  std::vector<cudaEvent_t> ev_in;                // per shard: it has consumed the caller's device buffers
  cudaEvent_t ev_q = nullptr;                    // home device: the caller's stream has produced the queries
  cudaEvent_t ev_merge = nullptr;                // home device: the previous search's merge has read the gather buffer
  bool merged_once = false;
  std::vector<void*> seg_dev;                    // per shard: device copy of {l0[n], off[n]} (multi-segment only)
  std::vector<size_t> seg_dev_n;
  std::vector<uint8_t> seg_dirty;
  void* gather = nullptr;                        // home device: D lists [G][cand] f32, then I lists [G][cand] i64
  size_t gather_bytes = 0;
  void* pin_q = nullptr;                         // pinned host staging: queries in
  size_t pin_q_bytes = 0;
  void* pin_o = nullptr;                         // pinned host staging: answers out
  size_t pin_o_bytes = 0;
  void* out_dev = nullptr;                       // home device: merged answer of a host-buffer search
  size_t out_dev_bytes = 0;
  std::unique_ptr<ShardPool> pool;
  int sticky = 0;                                // shard that takes the small appends at the moment
  int64_t opt_workers = 1;
  int64_t stat_merge_launches = 0;
  std::vector<std::string> errs;                 // per shard: message of a failed step (set_error is thread-local)
};

static FrontState* fs(sgic_index* f) { return static_cast<FrontState*>(f->front); }
static const FrontState* fs(const sgic_index* f) { return static_cast<const FrontState*>(f->front); }

// runs fn(g) -> rc for every shard, on the pool when it pays; first non-zero rc wins and its message is restored
static int for_shards(sgic_index* f, bool parallel, const std::function<int(int)>& fn) {
  FrontState* S = fs(f);
  const int G = static_cast<int>(f->shards.size());
  std::vector<int> rcs(static_cast<size_t>(G), 0);
  S->errs.assign(static_cast<size_t>(G), std::string());
  auto body = [&](int g) {
    rcs[static_cast<size_t>(g)] = fn(g);
    if (rcs[static_cast<size_t>(g)]) S->errs[static_cast<size_t>(g)] = g_err;
  };
  if (parallel && G > 1 && S->opt_workers) {
    if (!S->pool) {
      std::vector<int> devs;
      for (const sgic_index* s : f->shards) devs.push_back(s->device);
      S->pool.reset(new ShardPool(G, devs));
    }
    S->pool->run(body);
  } else {
    for (int g = 0; g < G; ++g) body(g);
  }
  for (int g = 0; g < G; ++g)
    if (rcs[static_cast<size_t>(g)]) {
      set_error("shard " + std::to_string(g) + " (device " + std::to_string(f->shards[static_cast<size_t>(g)]->device) +
                "): " + S->errs[static_cast<size_t>(g)]);
      return rcs[static_cast<size_t>(g)];
    }
  return 0;
}

static void front_note_segment(sgic_index* f, int g, int64_t global_start, int64_t local_start, int64_t count) {
  if (count <= 0) return;
  FrontState* S = fs(f);
  auto& v = S->segs[static_cast<size_t>(g)];
  if (!v.empty() && v.back().g0 + v.back().cnt == global_start && v.back().l0 + v.back().cnt == local_start)
    v.back().cnt += count;
  else
    v.push_back({global_start, local_start, count});
  S->seg_dirty[static_cast<size_t>(g)] = 1;
}

// local row numbers -> global ones for a shard that holds several runs: binary search over the run starts
__global__ void remap_push_kernel(const float* srcD, const long long* srcI, uint32_t n, const long long* seg_l0,
                                  const long long* seg_off, uint32_t n_seg, float* dstD, long long* dstI) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long id = srcI[i];
  if (id >= 0) {
    uint32_t lo = 0, hi = n_seg;  // last run with l0 <= id
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (seg_l0[mid] <= id) lo = mid;
      else hi = mid;
    }
    id += seg_off[lo];
  }
  dstD[i] = srcD[i];
  dstI[i] = id;
}

static int front_sync_seg_table(sgic_index* f, int g) {
  FrontState* S = fs(f);
  const size_t gi = static_cast<size_t>(g);
  if (!S->seg_dirty[gi]) return 0;
  const auto& v = S->segs[gi];
  if (v.size() > 1) {
    std::vector<long long> host(2 * v.size());
    for (size_t i = 0; i < v.size(); ++i) {
      host[i] = v[i].l0;
      host[v.size() + i] = v[i].g0 - v[i].l0;
    }
    if (v.size() > S->seg_dev_n[gi]) {
      if (S->seg_dev[gi]) SGIC_CUDA(cudaFree(S->seg_dev[gi]));
      S->seg_dev[gi] = nullptr;
      const size_t cap = std::max<size_t>(64, 2 * v.size());
      SGIC_CUDA(cudaMalloc(&S->seg_dev[gi], cap * 16));
      S->seg_dev_n[gi] = cap;
    }
    // layout {l0[cap], off[cap]} so that the table can grow in place up to cap
    sgic_index* s = f->shards[gi];
    SGIC_CUDA(cudaMemcpyAsync(S->seg_dev[gi], host.data(), v.size() * 8, cudaMemcpyHostToDevice, s->stream));
    SGIC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(S->seg_dev[gi]) + S->seg_dev_n[gi] * 8, host.data() + v.size(),
                              v.size() * 8, cudaMemcpyHostToDevice, s->stream));
    SGIC_CUDA(cudaStreamSynchronize(s->stream));  // `host` goes out of scope
  }
  S->seg_dirty[gi] = 0;
  return 0;
}

// One shard's part of a search: queries in, local scan, answer (global ids) into slot g of the home gather buffer.
//   q_pinned != nullptr: host-buffer search, the queries sit in pinned host memory
//   q_home   != nullptr: device search, the queries sit on the home GPU (ev_q recorded on the caller's stream)
static int front_shard_search(sgic_index* f, int g, int64_t nq, const float* q_pinned, const float* q_home, int64_t k,
                              float* dstD, int64_t* dstI) {
  FrontState* S = fs(f);
  const size_t gi = static_cast<size_t>(g);
  sgic_index* s = f->shards[gi];
  sgic_index* home = f->shards[0];
  std::lock_guard<std::mutex> lk(s->mu);
  DeviceGuard dg(s->device);
  cudaStream_t st = s->stream;
  int rc = order_begin(s, st);
  if (rc) return rc;
  if (S->merged_once) SGIC_CUDA(cudaStreamWaitEvent(st, S->ev_merge, 0));  // (long complete after a host-buffer search)
  const size_t qbytes = static_cast<size_t>(nq) * s->d * 4, cand = static_cast<size_t>(nq) * k;
  const float* q_dev = nullptr;
  if (q_pinned) {
    // host queries: one or two ride in the kernel parameters of the streaming scan, larger batches are copied from
    // the (shared, pinned) staging buffer by this GPU's copy engine — search_hostq_impl decides
  } else {
    SGIC_CUDA(cudaStreamWaitEvent(st, S->ev_q, 0));
    if (s->device == home->device || (S->direct[gi] && qbytes <= kZeroCopyBytes)) {
      q_dev = q_home;  // small batches are read out of the home GPU's memory over NVLink (K3: once per CTA)
    } else {
      if ((rc = ensure_buf(&s->qdev, &s->qdev_bytes, qbytes, false))) return rc;
      SGIC_CUDA(cudaMemcpyPeerAsync(s->qdev, s->device, q_home, home->device, qbytes, st));
      q_dev = static_cast<const float*>(s->qdev);
    }
  }
  if ((rc = front_sync_seg_table(f, g))) return rc;
  const auto& segs = S->segs[gi];
  const bool multi = segs.size() > 1;
  const int64_t id_base = segs.size() == 1 ? segs[0].g0 - segs[0].l0 : 0;
  if (!multi && S->direct[gi]) {
    // the shard's final writer stores straight into the home GPU's gather slot
    rc = q_pinned ? search_hostq_impl(s, nq, q_pinned, true, k, dstD, dstI, id_base, st)
                  : search_dev_impl(s, nq, q_dev, k, dstD, dstI, id_base, st);
    if (rc) return rc;
  } else {
    const size_t half = (cand * 12 + 15) & ~size_t(15);  // local answer | remapped answer, each ids (8 B) then scores
    if ((rc = ensure_buf(&s->odev, &s->odev_bytes, 2 * half, false))) return rc;
    int64_t* lI = reinterpret_cast<int64_t*>(s->odev);
    float* lD = reinterpret_cast<float*>(static_cast<uint8_t*>(s->odev) + cand * 8);
    rc = q_pinned ? search_hostq_impl(s, nq, q_pinned, true, k, lD, lI, id_base, st)
                  : search_dev_impl(s, nq, q_dev, k, lD, lI, id_base, st);
    if (rc) return rc;
    float* oD = lD;
    int64_t* oI = lI;
    if (multi) {
      oI = reinterpret_cast<int64_t*>(static_cast<uint8_t*>(s->odev) + half);
      oD = reinterpret_cast<float*>(static_cast<uint8_t*>(s->odev) + half + cand * 8);
      float* tD = S->direct[gi] ? dstD : oD;
      int64_t* tI = S->direct[gi] ? dstI : oI;
      const long long* tab = static_cast<const long long*>(S->seg_dev[gi]);
      remap_push_kernel<<<static_cast<unsigned>((cand + 255) / 256), 256, 0, st>>>(
          lD, reinterpret_cast<const long long*>(lI), static_cast<uint32_t>(cand), tab, tab + S->seg_dev_n[gi],
          static_cast<uint32_t>(segs.size()), tD, reinterpret_cast<long long*>(tI));
      s->stat_launches++;
      SGIC_CUDA(cudaGetLastError());
    }
    if (!S->direct[gi]) {  // no peer mapping: the copy engines carry the answer home
      SGIC_CUDA(cudaMemcpyPeerAsync(dstD, home->device, oD, s->device, cand * 4, st));
      SGIC_CUDA(cudaMemcpyPeerAsync(dstI, home->device, oI, s->device, cand * 8, st));
    }
  }
  SGIC_CUDA(cudaEventRecord(S->ev_done[gi], st));
  return order_end(s, st);
}

// gather buffer + the merge on `st` (home device).  Caller holds the home shard's mutex and device.
static int front_merge(sgic_index* f, int64_t nq, int64_t k, float* dev_D, int64_t* dev_I, cudaStream_t st) {
  FrontState* S = fs(f);
  sgic_index* home = f->shards[0];
  const int G = static_cast<int>(f->shards.size());
  const size_t cand = static_cast<size_t>(nq) * k;
  for (int g = 0; g < G; ++g)
    if (!(g == 0 && st == home->stream)) SGIC_CUDA(cudaStreamWaitEvent(st, S->ev_done[static_cast<size_t>(g)], 0));
  const float* Dl = static_cast<const float*>(S->gather);
  const int64_t* Il = reinterpret_cast<const int64_t*>(static_cast<const uint8_t*>(S->gather) +
                                                       ((static_cast<size_t>(G) * cand * 4 + 127) & ~size_t(127)));
  // ids beyond 32 bits cannot ride in the key: ties are then broken by list position, which equals the global row
  // order only when every shard holds ONE run and the runs ascend with the shard number
  int by_pos = 0;
  if (f->ntotal >= (1ll << 32) - 1) {
    int64_t prev_end = -1;
    for (int g = 0; g < G; ++g) {
      const auto& v = S->segs[static_cast<size_t>(g)];
      SGIC_REQUIRE(v.size() <= 1 && (v.empty() || v[0].g0 >= prev_end),
                   "more than 2^32-2 rows need one ascending row range per shard (reserve + one bulk load)");
      if (!v.empty()) prev_end = v[0].g0 + v[0].cnt;
    }
    by_pos = 1;
  }
  S->stat_merge_launches++;
  int rc = sgic_merge_topk_dev(home->device, nq, G, k, Dl, Il, dev_D, dev_I, by_pos, st);
  if (rc) return rc;
  SGIC_CUDA(cudaEventRecord(S->ev_merge, st));  // the next search's shards may not overwrite the slots before this
  S->merged_once = true;
  return 0;
}

static int front_ensure_gather(sgic_index* f, int64_t nq, int64_t k, float** D0, int64_t** I0) {
  FrontState* S = fs(f);
  const size_t G = f->shards.size(), cand = static_cast<size_t>(nq) * k;
  const size_t i_off = (G * cand * 4 + 127) & ~size_t(127);
  const size_t need = i_off + G * cand * 8;
  if (need > S->gather_bytes) {
    // nobody may still be writing into or merging out of the old buffer: a device-resident search returns without
    // a synchronise, so drain every shard stream and everything on the home GPU before the buffer is replaced
    for (sgic_index* s : f->shards) {
      DeviceGuard dg(s->device);
      SGIC_CUDA(cudaStreamSynchronize(s->stream));
    }
    SGIC_CUDA(cudaDeviceSynchronize());
    int rc = ensure_buf(&S->gather, &S->gather_bytes, need, false);
    if (rc) return rc;
  }
  *D0 = static_cast<float*>(S->gather);
  *I0 = reinterpret_cast<int64_t*>(static_cast<uint8_t*>(S->gather) + i_off);
  return 0;
}

static int front_search_host(sgic_index* f, int64_t nq, const float* host_q, int64_t k, float* host_D, int64_t* host_I) {
  FrontState* S = fs(f);
  SGIC_REQUIRE(k >= 1, "k must be >= 1");
  SGIC_REQUIRE(nq >= 0, "nq must be >= 0");
  if (nq == 0) return 0;
  SGIC_REQUIRE(host_q && host_D && host_I, "NULL buffer");
  std::lock_guard<std::mutex> lk(f->mu);
  sgic_index* home = f->shards[0];
  const size_t qbytes = static_cast<size_t>(nq) * f->d * 4, cand = static_cast<size_t>(nq) * k;
  float* gD = nullptr;
  int64_t* gI = nullptr;
  int rc;
  {
    DeviceGuard dg(home->device);
    if ((rc = ensure_buf(&S->pin_q, &S->pin_q_bytes, qbytes, true))) return rc;
    if ((rc = ensure_buf(&S->pin_o, &S->pin_o_bytes, cand * 12, true))) return rc;
    if ((rc = ensure_buf(&S->out_dev, &S->out_dev_bytes, cand * 12, false))) return rc;
    if ((rc = front_ensure_gather(f, nq, k, &gD, &gI))) return rc;
  }
  std::memcpy(S->pin_q, host_q, qbytes);
  rc = for_shards(f, true, [&](int g) {
    return front_shard_search(f, g, nq, static_cast<const float*>(S->pin_q), nullptr, k, gD + static_cast<size_t>(g) * cand,
                              gI + static_cast<size_t>(g) * cand);
  });
  if (rc) return rc;
  {
    std::lock_guard<std::mutex> hl(home->mu);
    DeviceGuard dg(home->device);
    if (cand * 12 <= kZeroCopyBytes) {  // the merge kernel stores the answer into the pinned host buffer itself
      int64_t* oI = reinterpret_cast<int64_t*>(S->pin_o);
      float* oD = reinterpret_cast<float*>(static_cast<uint8_t*>(S->pin_o) + cand * 8);
      if ((rc = front_merge(f, nq, k, oD, oI, home->stream))) return rc;
    } else {
      int64_t* oI = reinterpret_cast<int64_t*>(S->out_dev);
      float* oD = reinterpret_cast<float*>(static_cast<uint8_t*>(S->out_dev) + cand * 8);
      if ((rc = front_merge(f, nq, k, oD, oI, home->stream))) return rc;
      SGIC_CUDA(cudaMemcpyAsync(S->pin_o, S->out_dev, cand * 12, cudaMemcpyDeviceToHost, home->stream));
    }
    SGIC_CUDA(cudaStreamSynchronize(home->stream));
  }
  std::memcpy(host_I, S->pin_o, cand * 8);
  std::memcpy(host_D, static_cast<uint8_t*>(S->pin_o) + cand * 8, cand * 4);
  return 0;
}

// queries / outputs on the HOME device; `stream` is a stream of the home device (NULL: the home shard's own)
static int front_search_dev(sgic_index* f, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                            int64_t id_base, cudaStream_t stream) {
  FrontState* S = fs(f);
  SGIC_REQUIRE(k >= 1, "k must be >= 1");
  SGIC_REQUIRE(nq >= 0, "nq must be >= 0");
  SGIC_REQUIRE(id_base == 0, "id_base is not supported on a multi-GPU index (row numbers are global already)");
  if (nq == 0) return 0;
  std::lock_guard<std::mutex> lk(f->mu);
  sgic_index* home = f->shards[0];
  cudaStream_t st = stream ? stream : home->stream;
  const size_t cand = static_cast<size_t>(nq) * k;
  float* gD = nullptr;
  int64_t* gI = nullptr;
  int rc;
  {
    DeviceGuard dg(home->device);
    if ((rc = front_ensure_gather(f, nq, k, &gD, &gI))) return rc;
    SGIC_CUDA(cudaEventRecord(S->ev_q, st));
  }
  rc = for_shards(f, true, [&](int g) {
    return front_shard_search(f, g, nq, nullptr, dev_q, k, gD + static_cast<size_t>(g) * cand,
                              gI + static_cast<size_t>(g) * cand);
  });
  if (rc) return rc;
  std::lock_guard<std::mutex> hl(home->mu);
  DeviceGuard dg(home->device);
  return front_merge(f, nq, k, dev_D, dev_I, st);
}

// ---- appends ------------------------------------------------------------------------------------------------
struct FrontPiece {
  int g;
  int64_t lo, n;  // rows [lo, lo + n) of the block go to shard g
};

// Placement of a block of n rows.  Large blocks are cut into G contiguous, equal slices (contiguous row
// ranges per GPU); small ones go whole to the shard that is currently taking small appends, which is replaced by
// the least-full shard once it is 64k rows ahead — one row at a time (compress.py:300-305 adds vectors singly) must
// neither pile up on one GPU nor leave a run of one row per call behind.
static std::vector<FrontPiece> front_place(sgic_index* f, int64_t n) {
  FrontState* S = fs(f);
  const int G = static_cast<int>(f->shards.size());
  std::vector<FrontPiece> out;
  if (n <= 0) return out;
  if (n >= static_cast<int64_t>(G) * 4096) {
    const int64_t per = (n + G - 1) / G;
    for (int g = 0; g < G; ++g) {
      const int64_t lo = std::min<int64_t>(n, g * per), hi = std::min<int64_t>(n, lo + per);
      if (hi > lo) out.push_back({g, lo, hi - lo});
    }
    return out;
  }
  int least = 0;
  for (int g = 1; g < G; ++g)
    if (f->shards[static_cast<size_t>(g)]->ntotal < f->shards[static_cast<size_t>(least)]->ntotal) least = g;
  if (f->shards[static_cast<size_t>(S->sticky)]->ntotal - f->shards[static_cast<size_t>(least)]->ntotal >= 65536)
    S->sticky = least;
  out.push_back({S->sticky, 0, n});
  return out;
}

// An append that failed on one shard must not leave rows behind on the others: rows outside every run would still be
// scanned and come back under the row number of a neighbouring run.  Every shard is cut back to the row count it
// had before the call (the rows' bytes stay in HBM beyond ntotal and are overwritten by the next append).
static void front_rollback(sgic_index* f, const std::vector<int64_t>& before) {
  for (size_t g = 0; g < f->shards.size(); ++g) {
    sgic_index* s = f->shards[g];
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->ntotal <= before[g]) continue;
    {
      DeviceGuard dg(s->device);
      cudaStreamSynchronize(s->stream);  // nothing of the failed call is still in flight when the caller retries
      (void)cudaGetLastError();
    }
    const size_t keep = static_cast<size_t>(before[g]) * static_cast<size_t>(s->d);
    if (s->retain_ok && s->retained.size() > keep) s->retained.resize(keep);
    if (s->codes_ok && s->codes.size() > keep) s->codes.resize(keep);
    s->ntotal = before[g];
  }
}

// host rows: kind 0 = fp32 (add_f32), 1 = u8 codes (add_u8)
static int front_add_host(sgic_index* f, int64_t n, const void* host, int kind) {
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  SGIC_REQUIRE(host != nullptr, "x is NULL");
  std::lock_guard<std::mutex> lk(f->mu);
  const std::vector<FrontPiece> pieces = front_place(f, n);
  const size_t row_bytes = static_cast<size_t>(f->d) * (kind == 0 ? 4 : 1);
  std::vector<int64_t> before(f->shards.size(), 0);
  for (size_t g = 0; g < f->shards.size(); ++g) before[g] = f->shards[g]->ntotal;
  std::vector<const FrontPiece*> of_shard(f->shards.size(), nullptr);
  for (const auto& p : pieces) of_shard[static_cast<size_t>(p.g)] = &p;
  int rc = for_shards(f, pieces.size() > 1, [&](int g) -> int {
    const FrontPiece* p = of_shard[static_cast<size_t>(g)];
    if (!p) return 0;
    const uint8_t* src = static_cast<const uint8_t*>(host) + static_cast<size_t>(p->lo) * row_bytes;
    sgic_index* s = f->shards[static_cast<size_t>(g)];
    return kind == 0 ? sgic_index_add_f32(s, p->n, reinterpret_cast<const float*>(src)) : sgic_index_add_u8(s, p->n, src);
  });
  if (rc) {
    const std::string msg = g_err;
    front_rollback(f, before);
    set_error(msg);
    return rc;
  }
  for (const auto& p : pieces) front_note_segment(f, p.g, f->ntotal + p.lo, before[static_cast<size_t>(p.g)], p.n);
  f->ntotal += n;
  return 0;
}

// device rows on the HOME device: kind 0 = fp32, 1 = u8 codes, 2 = rows already in the storage type.  Shards
// read their slice out of the home GPU's memory over NVLink.
static int front_add_dev(sgic_index* f, int64_t n, const void* dev, int kind, cudaStream_t stream) {
  FrontState* S = fs(f);
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(f->mu);
  sgic_index* home = f->shards[0];
  cudaStream_t st = stream ? stream : home->stream;
  const std::vector<FrontPiece> pieces = front_place(f, n);
  const size_t row_bytes = static_cast<size_t>(f->d) * (kind == 0 ? 4 : kind == 1 ? 1 : 2);
  for (const auto& p : pieces)
    SGIC_REQUIRE(S->direct[static_cast<size_t>(p.g)],
                 "device-resident appends need peer access from every shard to the home GPU");
  std::vector<int64_t> before(f->shards.size(), 0);
  for (size_t g = 0; g < f->shards.size(); ++g) before[g] = f->shards[g]->ntotal;
  {
    DeviceGuard dg(home->device);
    SGIC_CUDA(cudaEventRecord(S->ev_q, st));
  }
  auto one_piece = [&](const FrontPiece& p) -> int {
    const size_t gi = static_cast<size_t>(p.g);
    sgic_index* s = f->shards[gi];
    const uint8_t* src = static_cast<const uint8_t*>(dev) + static_cast<size_t>(p.lo) * row_bytes;
    {
      DeviceGuard dg(s->device);
      SGIC_CUDA(cudaStreamWaitEvent(s->stream, S->ev_q, 0));
    }
    int rc = kind == 0   ? sgic_index_add_f32_dev(s, p.n, reinterpret_cast<const float*>(src), nullptr)
             : kind == 1 ? sgic_index_add_u8_dev(s, p.n, src, nullptr)
                         : sgic_index_add_packed_dev(s, p.n, src, nullptr);
    if (rc) return rc;
    {
      DeviceGuard dg(s->device);
      SGIC_CUDA(cudaEventRecord(S->ev_in[gi], s->stream));
    }
    DeviceGuard dg(home->device);
    SGIC_CUDA(cudaStreamWaitEvent(st, S->ev_in[gi], 0));  // the caller may reuse / free its buffer in stream order
    return 0;
  };
  for (const auto& p : pieces) {
    const int rc = one_piece(p);
    if (rc) {
      const std::string msg = g_err;
      front_rollback(f, before);
      set_error(msg);
      return rc;
    }
  }
  for (const auto& p : pieces) front_note_segment(f, p.g, f->ntotal + p.lo, before[static_cast<size_t>(p.g)], p.n);
  f->ntotal += n;
  return 0;
}

// build.py:80-88 over G GPUs: the file list is cut into G contiguous slices, every shard walks / decodes its slice
// with its share of the host threads, failed files are skipped, and the global row number of a file is the number
// of good files before it — known once every slice has reported its count.
static int front_add_c2df(sgic_index* f, const uint8_t* blob, const int64_t* offsets, int64_t n, int32_t* status_out,
                          int64_t* n_added, int n_threads) {
  SGIC_REQUIRE(n >= 0 && offsets != nullptr && status_out != nullptr, "bad arguments");
  if (n_added) *n_added = 0;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(f->mu);
  const int G = static_cast<int>(f->shards.size());
  const int64_t per = (n + G - 1) / G;
  if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
  const int per_threads = std::max(1, n_threads / G);
  std::vector<int64_t> added(static_cast<size_t>(G), 0), before(static_cast<size_t>(G), 0);
  for (int g = 0; g < G; ++g) before[static_cast<size_t>(g)] = f->shards[static_cast<size_t>(g)]->ntotal;
  int rc = for_shards(f, true, [&](int g) -> int {
    const int64_t lo = std::min<int64_t>(n, g * per), hi = std::min<int64_t>(n, lo + per);
    if (hi <= lo) return 0;
    return sgic_index_add_c2df(f->shards[static_cast<size_t>(g)], blob, offsets + lo, hi - lo, status_out + lo,
                               &added[static_cast<size_t>(g)], per_threads);
  });
  if (rc) {
    const std::string msg = g_err;
    front_rollback(f, before);
    set_error(msg);
    return rc;
  }
  int64_t start = f->ntotal;
  for (int g = 0; g < G; ++g) {
    front_note_segment(f, g, start, before[static_cast<size_t>(g)], added[static_cast<size_t>(g)]);
    start += added[static_cast<size_t>(g)];
  }
  if (n_added) *n_added = start - f->ntotal;
  f->ntotal = start;
  return 0;
}

// ---- rows back out --------------------------------------------------------------------------------------------
struct FrontRun {
  int64_t g0, cnt, l0;
  int g;
};
static std::vector<FrontRun> front_runs(const sgic_index* f) {
  const FrontState* S = fs(f);
  std::vector<FrontRun> runs;
  for (size_t g = 0; g < S->segs.size(); ++g)
    for (const auto& s : S->segs[g]) runs.push_back({s.g0, s.cnt, s.l0, static_cast<int>(g)});
  std::sort(runs.begin(), runs.end(), [](const FrontRun& a, const FrontRun& b) { return a.g0 < b.g0; });
  return runs;
}

// calls fn(shard, local_start, count, offset_in_output_rows) for the pieces of global rows [i0, i0 + n)
template <typename F>
static int front_for_rows(sgic_index* f, int64_t i0, int64_t n, F fn) {
  const std::vector<FrontRun> runs = front_runs(f);
  int64_t done = 0;
  size_t r = static_cast<size_t>(std::upper_bound(runs.begin(), runs.end(), i0,
                                                  [](int64_t v, const FrontRun& a) { return v < a.g0; }) -
                                 runs.begin());
  r = r ? r - 1 : 0;
  while (done < n) {
    SGIC_REQUIRE(r < runs.size() && runs[r].g0 <= i0 + done && i0 + done < runs[r].g0 + runs[r].cnt,
                 "row map of the multi-GPU index is inconsistent");
    const int64_t in_run = i0 + done - runs[r].g0;
    const int64_t take = std::min(n - done, runs[r].cnt - in_run);
    int rc = fn(f->shards[static_cast<size_t>(runs[r].g)], runs[r].l0 + in_run, take, done);
    if (rc) return rc;
    done += take;
    ++r;
  }
  return 0;
}

static int front_reconstruct(sgic_index* f, int64_t i0, int64_t n, float* host_out) {
  SGIC_REQUIRE(i0 >= 0 && n >= 0 && i0 + n <= f->ntotal, "row range out of bounds");
  std::lock_guard<std::mutex> lk(f->mu);
  return front_for_rows(f, i0, n, [&](sgic_index* s, int64_t l0, int64_t cnt, int64_t off) {
    return sgic_index_reconstruct(s, l0, cnt, host_out + static_cast<size_t>(off) * f->d);
  });
}

static int front_codes(sgic_index* f, int64_t i0, int64_t n, uint8_t* host_out) {
  SGIC_REQUIRE(i0 >= 0 && n >= 0 && i0 + n <= f->ntotal, "row range out of bounds");
  std::lock_guard<std::mutex> lk(f->mu);
  return front_for_rows(f, i0, n, [&](sgic_index* s, int64_t l0, int64_t cnt, int64_t off) {
    return sgic_index_codes(s, l0, cnt, host_out + static_cast<size_t>(off) * f->d);
  });
}

// fp32 rows [i0, i0 + n) for the IxFI writer: from the host copies (fp32 rows / u8 codes) when EVERY shard still
// holds them for all of its rows, from HBM otherwise.  mode: 1 retained fp32, 2 retained codes, 0 HBM.
static int front_write_mode(const sgic_index* f) {
  bool f32 = true, codes = true;
  for (const sgic_index* s : f->shards) {
    const size_t want = static_cast<size_t>(s->ntotal) * static_cast<size_t>(s->d);
    f32 = f32 && s->retain_ok && s->retained.size() == want;
    codes = codes && s->codes_ok && s->codes.size() == want;
  }
  return f32 ? 1 : codes ? 2 : 0;
}

static int front_rows_for_write(sgic_index* f, int mode, int64_t i0, int64_t n, float* out) {
  return front_for_rows(f, i0, n, [&](sgic_index* s, int64_t l0, int64_t cnt, int64_t off) -> int {
    float* dst = out + static_cast<size_t>(off) * f->d;
    if (mode == 1) {
      std::memcpy(dst, s->retained.data() + static_cast<size_t>(l0) * s->d, static_cast<size_t>(cnt) * s->d * 4);
      return 0;
    }
    if (mode == 2) {
      rows_from_codes_f32(s->codes.data() + static_cast<size_t>(l0) * s->d, cnt, s->d, dst);
      return 0;
    }
    return sgic_index_reconstruct(s, l0, cnt, dst);
  });
}

static void front_destroy(sgic_index* f) {
  FrontState* S = fs(f);
  if (S) {
    S->pool.reset();
    for (size_t g = 0; g < f->shards.size(); ++g) {
      sgic_index* s = f->shards[g];
      if (!s) continue;
      DeviceGuard dg(s->device);
      cudaStreamSynchronize(s->stream);
      if (g < S->ev_done.size() && S->ev_done[g]) cudaEventDestroy(S->ev_done[g]);
      if (g < S->ev_in.size() && S->ev_in[g]) cudaEventDestroy(S->ev_in[g]);
      if (g < S->seg_dev.size() && S->seg_dev[g]) cudaFree(S->seg_dev[g]);
    }
    if (!f->shards.empty() && f->shards[0]) {
      DeviceGuard dg(f->shards[0]->device);
      if (S->ev_q) cudaEventDestroy(S->ev_q);
      if (S->ev_merge) cudaEventDestroy(S->ev_merge);
      if (S->gather) cudaFree(S->gather);
      if (S->out_dev) cudaFree(S->out_dev);
      if (S->pin_q) cudaFreeHost(S->pin_q);
      if (S->pin_o) cudaFreeHost(S->pin_o);
    }
  }
  for (sgic_index* s : f->shards) sgic_index_destroy(s);
  f->shards.clear();
  delete S;
  f->front = nullptr;
}

// one run per shard, in shard order: the layout of a bulk load (shards filled directly, or read from shard files)
static void front_adopt_contiguous(sgic_index* f) {
  FrontState* S = fs(f);
  int64_t start = 0;
  for (size_t g = 0; g < f->shards.size(); ++g) {
    S->segs[g].clear();
    if (f->shards[g]->ntotal > 0) S->segs[g].push_back({start, 0, f->shards[g]->ntotal});
    S->seg_dirty[g] = 1;
    start += f->shards[g]->ntotal;
  }
  f->ntotal = start;
}

static int front_init(sgic_index* f, const std::vector<sgic_index*>& shards) {
  FrontState* S = new FrontState();
  f->front = S;
  f->shards = shards;
  const size_t G = shards.size();
  S->segs.resize(G);
  S->direct.assign(G, 0);
  S->ev_done.assign(G, nullptr);
  S->ev_in.assign(G, nullptr);
  S->seg_dev.assign(G, nullptr);
  S->seg_dev_n.assign(G, 0);
  S->seg_dirty.assign(G, 0);
  if (const char* w = std::getenv("SGIC_FRONT_WORKERS")) S->opt_workers = std::atoi(w) ? 1 : 0;  // A/B runs
  sgic_index* home = shards[0];
  f->d = home->d;
  f->dtype = home->dtype;
  f->device = home->device;
  f->sm_count = home->sm_count;
  {
    DeviceGuard dg(home->device);
    SGIC_CUDA(cudaEventCreateWithFlags(&S->ev_q, cudaEventDisableTiming));
    SGIC_CUDA(cudaEventCreateWithFlags(&S->ev_merge, cudaEventDisableTiming));
  }
  for (size_t g = 0; g < G; ++g) {
    sgic_index* s = shards[g];
    DeviceGuard dg(s->device);
    SGIC_CUDA(cudaEventCreateWithFlags(&S->ev_done[g], cudaEventDisableTiming));
    SGIC_CUDA(cudaEventCreateWithFlags(&S->ev_in[g], cudaEventDisableTiming));
    if (s->device == home->device) {
      S->direct[g] = 1;
      continue;
    }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, s->device, home->device) == cudaSuccess && can) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(home->device, 0);
      if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) S->direct[g] = 1;
      (void)cudaGetLastError();
    }
  }
  return 0;
}

}  // namespace sgic
