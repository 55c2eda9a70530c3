// C-ABI implementation (include/sgic.h): index object, HBM database, ingest, search dispatch,
// IxFI (de)serialisation.  Host logic only; the kernels live in the .cuh files next to this.
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sgic.h"
#include "ingest.cuh"
#include "scan_dense.cuh"
#include "scan_small.cuh"
#include "topk_common.cuh"

namespace sgic {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

int c2df_parse_batch(const uint8_t* blob, const int64_t* offsets, int64_t n, int dim, uint8_t* out_u8,
                     int32_t* status_out, int32_t* dim_out, int n_threads, int64_t* frame_off = nullptr,
                     uint32_t* frame_len = nullptr);
int c2df_pack_batch(const uint8_t* blob, int64_t cnt, int dim, const int32_t* status, const int64_t* frame_off,
                    const uint32_t* frame_len, const uint8_t* rows_in, uint8_t* rows_out, uint8_t* frames_out,
                    size_t frames_cap, ZlDesc* desc_out, int64_t* n_rows_out, int64_t* n_frames_out,
                    int64_t* n_host_rows_out, size_t* frames_bytes_out, int n_threads);

#define SGIC_CUDA(call)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (call);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                std::to_string(__LINE__) + ")");                                                \
      return 2;                                                                                 \
    }                                                                                           \
  } while (0)

#define SGIC_REQUIRE(cond, msg) \
  do {                          \
    if (!(cond)) {              \
      set_error(msg);           \
      return 1;                 \
    }                           \
  } while (0)

constexpr size_t kSmemBudget = 232448;  // 227 KB opt-in maximum per CTA on sm_100
constexpr int kMaxD = 2048;
constexpr int64_t kMaxK = 1024;
constexpr size_t kStageChunkBytes = 32u << 20;  // host->device staging granularity for add()
constexpr size_t kZeroCopyBytes = 64u << 10;     // queries / answers up to this size are read / written in place in pinned host memory

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace sgic

namespace sgic {
struct ShardSeg {  // a run of consecutive rows of a shard: global rows [g0, g0 + cnt) = local rows [l0, l0 + cnt)
  int64_t g0, l0, cnt;
};
}  // namespace sgic

struct sgic_index {
  // multi-GPU front (sgic_index_create_sharded): `shards` holds one ordinary index per GPU and `front` the
  // gather / staging state (sharded_front.inl); d, dtype, device (= home GPU) and ntotal describe the whole index
  std::vector<sgic_index*> shards;
  void* front = nullptr;
  int d = 0, dtype = SGIC_F16, device = 0, flags = 0;
  int64_t ntotal = 0, capacity = 0;
  void* db = nullptr;
  // The database lives in ONE virtual address range reserved up front (cuMemAddressReserve, the size of the GPU's
  // memory) that physical chunks are mapped into as the index grows (cuMemCreate / cuMemMap): growing never copies
  // and never needs old + new side by side, so an index can grow past half of HBM.  vmm_base == 0: the driver
  // calls are not available (or SGIC_VMM=0) and the database is a cudaMalloc that is re-allocated to grow.
  unsigned long long vmm_base = 0;
  size_t vmm_va_bytes = 0, vmm_mapped = 0, vmm_gran = 0;
  std::vector<unsigned long long> vmm_handles;
  std::vector<size_t> vmm_sizes;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::mutex mu;
  // staging (double-buffered) for add()
  void* pin[2] = {nullptr, nullptr};
  void* dstage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  // search workspaces
  void* ws = nullptr;  // partial keys (+ the fused merge's CTA counter)
  void* ws2 = nullptr;  // k > 1024: the bound handed from one pass to the next
  size_t ws2_bytes = 0;
  uint32_t* ws_counter = nullptr;
  size_t ws_bytes = 0;
  void* qh = nullptr;  // queries rounded to the storage dtype (dense path operand A)
  size_t qh_bytes = 0;
  void* lists_ws = nullptr;  // dense path, k > 32: per-query reservoirs
  size_t lists_ws_bytes = 0;
  void* seed_ws = nullptr;   // dense path, k > 32: answer of the sample pass + the per-query start thresholds
  size_t seed_ws_bytes = 0;
  // device-side clip_stream decode (K0): u8 row matrix, packed frames, descriptors, per-frame status
  // ingest pipeline (device-side clip_stream decode): kZlDepth slabs in flight — the host walks / packs slab i
  // while slab i-1 crosses PCIe on the copy stream and K0 decodes slab i-2 on the index's stream
  struct ZlSlab {
    void *pin_rows = nullptr, *pin_frames = nullptr, *pin_desc = nullptr, *status_host = nullptr;  // pinned
    size_t pin_rows_bytes = 0, pin_frames_bytes = 0, pin_desc_bytes = 0, status_host_bytes = 0;
    void *rows = nullptr, *frames = nullptr, *desc = nullptr, *status = nullptr;                    // device
    size_t rows_bytes = 0, frames_bytes = 0, desc_bytes = 0, status_bytes = 0;
    cudaEvent_t copied = nullptr, decoded = nullptr, freed = nullptr;
    cudaEvent_t t_c0 = nullptr, t_c1 = nullptr, t_k0 = nullptr, t_k1 = nullptr, t_q0 = nullptr, t_q1 = nullptr;  // "timing" option
    bool used = false, timed = false;
    // the slab currently in the set
    int64_t s0 = 0, cnt = 0, w = 0, nf = 0, nh = 0, full = 0;  // full: files of a full slab (buffer sizing)
    size_t fbytes = 0;
    int mode = 0;  // 0 nothing to append, 1 device decode, 2 host route (nothing in the device profile)
  };
  static constexpr int kZlDepth = 3;
  ZlSlab zl[kZlDepth];
  cudaStream_t copy_stream = nullptr;
  void* qdev = nullptr;  // queries / outputs for the host-buffer search
  size_t qdev_bytes = 0;
  void* qpin = nullptr;  // pinned staging of host queries on their way to qdev
  size_t qpin_bytes = 0;
  void* odev = nullptr;
  size_t odev_bytes = 0;
  void* opin = nullptr;
  size_t opin_bytes = 0;
  cudaEvent_t t0 = nullptr, t1 = nullptr, tm = nullptr;
  // "timing" = 2: the scan kernel of every search is bracketed by a pair of events from this ring WITHOUT a
  // synchronise, so that a caller can time its own loop and read the kernels' durations afterwards
  // (sgic_index_scan_times) — the roofline figure of bench.py comes from inside its timed region this way
  static constexpr int kTimeRing = 256;
  std::vector<cudaEvent_t> ring_a, ring_b;
  int64_t ring_n = 0;
  // Cross-stream ordering.  The *_dev entry points enqueue on the CALLER's stream, the host-buffer calls on the
  // index's own stream, and all of them share the database and the search workspaces: every call first makes its
  // stream wait for the work the previous call left on a different stream (order_begin / order_end below).
  cudaEvent_t ev_order = nullptr;
  cudaStream_t last_stream = nullptr;
  bool last_stream_valid = false;
  // retained fp32 rows (SGIC_RETAIN_F32)
  std::vector<float> retained;
  bool retain_ok = false;
  // retained u8 codes (SGIC_RETAIN_U8): rows that came in as clip_stream codes (add_u8 / add_c2df) are kept on the
  // host, 1 byte per element, so that sgic_index_write can regenerate the fp32 rows the reference's
  // dequantize_clip_u8 + l2n (src/build.py:18-24) would have handed to faiss — bit for bit
  std::vector<uint8_t> codes;
  bool codes_ok = false;
  // shard placement recorded in / restored from an SGI2 file (single index: 0, ntotal, 0, 1)
  int64_t shard_row_start = 0, shard_total_rows = -1;
  int shard_id = 0, shard_count = 1;
  // options / stats
  int64_t opt_timing = 0, opt_evict_first = 1, opt_grid = 0, opt_stages = 0, opt_rb = 0, opt_fused = 1, opt_dense_min_nq = 0, opt_debug = 0, opt_dense_mode = 0, opt_device_zstd = 1, opt_dense_l2_mb = 0, opt_dense_b_min_mb = 90000, opt_dense_gthr = 1, opt_trace = 0, opt_steal = 1, opt_t_n8 = 1, opt_t_max_nq = 88, opt_dense_seed = 1, opt_dense_slices = 0, opt_epi_wait_ns = 0, opt_scan_wait_ns = 0, opt_fail_appends = 0;
  int64_t stat_ingest_h2d_ns = 0, stat_ingest_k0_ns = 0, stat_ingest_k1_ns = 0;
  int64_t stat_ingest_parse_ns = 0, stat_ingest_pack_ns = 0, stat_ingest_gpu_ns = 0;
  int64_t stat_zl_device_frames = 0, stat_zl_host_rows = 0, stat_zl_fallback_slabs = 0;
  int64_t stat_launches = 0, stat_last_search_ns = 0, stat_last_scan_ns = 0, stat_last_grid = 0, stat_last_stages = 0, stat_last_kernel = 0;
};

namespace sgic {

static size_t elt_rows_bytes(const sgic_index* h, int64_t rows) { return static_cast<size_t>(rows) * h->d * 2; }

// Stream hand-over (call with the index mutex held).  order_begin(st): if the previous call enqueued on another
// stream, an event is recorded on THAT stream now (everything enqueued there so far, which includes our work) and
// `st` waits for it.  The previous stream may be a caller's stream that no longer exists: then the whole device is
// synchronised instead, which is always safe.  order_end(st) only remembers the stream — calls that keep using one
// stream (the common case: a search loop) pay nothing.
static int order_begin(sgic_index* h, cudaStream_t st) {
  if (h->last_stream_valid && h->last_stream != st) {
    if (cudaEventRecord(h->ev_order, h->last_stream) == cudaSuccess) {
      SGIC_CUDA(cudaStreamWaitEvent(st, h->ev_order, 0));
    } else {
      (void)cudaGetLastError();
      SGIC_CUDA(cudaDeviceSynchronize());
    }
  }
  return 0;
}
static int order_end(sgic_index* h, cudaStream_t st) {
  h->last_stream = st;
  h->last_stream_valid = true;
  return 0;
}

static int ring_mark(sgic_index* h, bool begin, cudaStream_t st) {
  if (h->opt_timing != 2) return 0;
  if (h->ring_a.empty()) {
    h->ring_a.resize(sgic_index::kTimeRing);
    h->ring_b.resize(sgic_index::kTimeRing);
    for (int i = 0; i < sgic_index::kTimeRing; ++i) {
      SGIC_CUDA(cudaEventCreate(&h->ring_a[i]));
      SGIC_CUDA(cudaEventCreate(&h->ring_b[i]));
    }
  }
  const int slot = static_cast<int>(h->ring_n % sgic_index::kTimeRing);
  if (begin) {
    SGIC_CUDA(cudaEventRecord(h->ring_a[slot], st));
  } else {
    SGIC_CUDA(cudaEventRecord(h->ring_b[slot], st));
    h->ring_n++;
  }
  return 0;
}

// ---- virtual-memory backed database ------------------------------------------------------------------------
struct VmmApi {
  CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*addr_free)(CUdeviceptr, size_t) = nullptr;
  CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  bool ok = false;
};

static const VmmApi& vmm_api() {
  static VmmApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* env = std::getenv("SGIC_VMM");
    if (env && std::atoi(env) == 0) return;
    auto get = [](const char* name, void** out) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, out, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess &&
             *out != nullptr;
    };
    api.ok = get("cuMemAddressReserve", reinterpret_cast<void**>(&api.reserve)) &&
             get("cuMemAddressFree", reinterpret_cast<void**>(&api.addr_free)) &&
             get("cuMemCreate", reinterpret_cast<void**>(&api.create)) &&
             get("cuMemRelease", reinterpret_cast<void**>(&api.release)) &&
             get("cuMemMap", reinterpret_cast<void**>(&api.map)) && get("cuMemUnmap", reinterpret_cast<void**>(&api.unmap)) &&
             get("cuMemSetAccess", reinterpret_cast<void**>(&api.set_access)) &&
             get("cuMemGetAllocationGranularity", reinterpret_cast<void**>(&api.granularity));
    (void)cudaGetLastError();
  });
  return api;
}

static CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  return prop;
}

// reserve the address range once (first growth); false: fall back to cudaMalloc
static bool vmm_init(sgic_index* h) {
  const VmmApi& v = vmm_api();
  if (h->vmm_base) return true;
  if (!v.ok || h->db != nullptr) return false;  // (an index that already holds a cudaMalloc database keeps that scheme)
  const CUmemAllocationProp prop = vmm_prop(h->device);
  size_t gran = 0;
  if (v.granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return false;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
  const size_t va = ((total_b + gran - 1) / gran) * gran;
  CUdeviceptr base = 0;
  if (v.reserve(&base, va, 0, 0, 0) != CUDA_SUCCESS) return false;
  h->vmm_base = base;
  h->vmm_va_bytes = va;
  h->vmm_gran = gran;
  return true;
}

// map `bytes` (rounded up to the granularity) more physical memory behind what is mapped already
static int vmm_grow(sgic_index* h, size_t bytes) {
  const VmmApi& v = vmm_api();
  bytes = ((bytes + h->vmm_gran - 1) / h->vmm_gran) * h->vmm_gran;
  if (h->vmm_mapped + bytes > h->vmm_va_bytes) return 3;
  const CUmemAllocationProp prop = vmm_prop(h->device);
  CUmemGenericAllocationHandle handle = 0;
  if (v.create(&handle, bytes, &prop, 0) != CUDA_SUCCESS) return 3;  // out of memory: the caller may retry smaller
  if (v.map(h->vmm_base + h->vmm_mapped, bytes, 0, handle, 0) != CUDA_SUCCESS) {
    v.release(handle);
    return 2;
  }
  CUmemAccessDesc acc;
  std::memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = h->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (v.set_access(h->vmm_base + h->vmm_mapped, bytes, &acc, 1) != CUDA_SUCCESS) {
    v.unmap(h->vmm_base + h->vmm_mapped, bytes);
    v.release(handle);
    return 2;
  }
  h->vmm_handles.push_back(handle);
  h->vmm_sizes.push_back(bytes);
  h->vmm_mapped += bytes;
  return 0;
}

static void vmm_destroy(sgic_index* h) {
  if (!h->vmm_base) return;
  const VmmApi& v = vmm_api();
  size_t off = 0;
  for (size_t i = 0; i < h->vmm_handles.size(); ++i) {
    v.unmap(h->vmm_base + off, h->vmm_sizes[i]);
    v.release(h->vmm_handles[i]);
    off += h->vmm_sizes[i];
  }
  v.addr_free(h->vmm_base, h->vmm_va_bytes);
  h->vmm_handles.clear();
  h->vmm_sizes.clear();
  h->vmm_base = 0;
  h->vmm_mapped = 0;
  h->db = nullptr;
}

// exact = true: an explicit reserve() — no geometric slack
static int grow_database(sgic_index* h, int64_t rows, cudaStream_t st, bool exact) {
  if (h->opt_fail_appends > 0) {  // fault injection for the tests of the error paths (option "fail_appends" = count)
    --h->opt_fail_appends;
    set_error("out of device memory (injected by the fail_appends option)");
    return 3;
  }
  if (rows <= h->capacity) return 0;
  const size_t row_bytes = static_cast<size_t>(h->d) * 2;
  if (vmm_init(h)) {
    const size_t need = static_cast<size_t>(rows) * row_bytes;
    if (need > h->vmm_mapped) {
      size_t more = need - h->vmm_mapped;
      // appends without a reserve(): grow by half of what is there (at least 32 MB) so that n small adds map O(log n) chunks
      const size_t slack = exact ? more : std::max<size_t>(more, std::max<size_t>(h->vmm_mapped / 2, size_t(32) << 20));
      int rc = vmm_grow(h, slack);
      if (rc == 3 && slack > more) rc = vmm_grow(h, more);
      if (rc) {
        set_error("mapping " + std::to_string(more) + " more bytes for the database failed (" +
                  std::to_string(h->vmm_mapped) + " mapped): out of device memory");
        return 2;
      }
    }
    h->db = reinterpret_cast<void*>(static_cast<uintptr_t>(h->vmm_base));
    h->capacity = static_cast<int64_t>(h->vmm_mapped / row_bytes);
    return 0;
  }
  int64_t cap = exact ? rows : std::max<int64_t>(rows, std::max<int64_t>(1024, h->capacity + h->capacity / 2));
  void* nb = nullptr;
  cudaError_t e = cudaMalloc(&nb, elt_rows_bytes(h, cap));
  if (e != cudaSuccess && cap > rows) {
    cudaGetLastError();
    cap = rows;
    e = cudaMalloc(&nb, elt_rows_bytes(h, cap));
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc of " + std::to_string(elt_rows_bytes(h, cap)) + " bytes for the database failed: " +
              cudaGetErrorString(e));
    return 2;
  }
  if (h->ntotal > 0) {
    SGIC_CUDA(cudaMemcpyAsync(nb, h->db, elt_rows_bytes(h, h->ntotal), cudaMemcpyDeviceToDevice, st));
    SGIC_CUDA(cudaStreamSynchronize(st));
  }
  if (h->db) SGIC_CUDA(cudaFree(h->db));
  h->db = nb;
  h->capacity = cap;
  return 0;
}

static int ensure_capacity(sgic_index* h, int64_t rows, cudaStream_t st) { return grow_database(h, rows, st, false); }

static int ensure_buf(void** p, size_t* cur, size_t need, bool pinned) {
  if (need <= *cur) return 0;
  if (*p) {
    if (pinned) SGIC_CUDA(cudaFreeHost(*p));
    else SGIC_CUDA(cudaFree(*p));
    *p = nullptr;
    *cur = 0;
  }
  need = (need + 4095) & ~static_cast<size_t>(4095);
  if (pinned) SGIC_CUDA(cudaMallocHost(p, need));
  else SGIC_CUDA(cudaMalloc(p, need));
  *cur = need;
  return 0;
}

static int ensure_staging(sgic_index* h) {
  for (int i = 0; i < 2; ++i) {
    if (!h->pin[i]) SGIC_CUDA(cudaMallocHost(&h->pin[i], kStageChunkBytes));
    if (!h->dstage[i]) SGIC_CUDA(cudaMalloc(&h->dstage[i], kStageChunkBytes));
    if (!h->ev[i]) SGIC_CUDA(cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming));
  }
  return 0;
}

static unsigned grid_for(size_t n_items, int threads, int sm_count) {
  size_t blocks = (n_items + threads - 1) / threads;
  size_t cap = static_cast<size_t>(sm_count) * 8;
  return static_cast<unsigned>(std::max<size_t>(1, std::min(blocks, cap)));
}

// ---- K2 / K1 launches -------------------------------------------------------------------
static int launch_pack_f32(sgic_index* h, const float* dev_src, int64_t row0, int64_t n, cudaStream_t st) {
  const size_t n8 = static_cast<size_t>(n) * h->d / 8;
  void* dst = static_cast<uint8_t*>(h->db) + elt_rows_bytes(h, row0);
  const unsigned g = grid_for(n8, 256, h->sm_count);
  if (h->dtype == SGIC_F16) pack_f32_kernel<__half><<<g, 256, 0, st>>>(dev_src, dst, n8);
  else pack_f32_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(dev_src, dst, n8);
  h->stat_launches++;
  SGIC_CUDA(cudaGetLastError());
  return 0;
}

static int launch_dequant_u8(sgic_index* h, const uint8_t* dev_src, int64_t row0, int64_t n, cudaStream_t st) {
  void* dst = static_cast<uint8_t*>(h->db) + elt_rows_bytes(h, row0);
  const unsigned g = grid_for(static_cast<size_t>(n) * 32, 256, h->sm_count);
  if (h->dtype == SGIC_F16)
    dequant_u8_kernel<__half><<<g, 256, 0, st>>>(dev_src, dst, static_cast<uint32_t>(n), static_cast<uint32_t>(h->d));
  else
    dequant_u8_kernel<__nv_bfloat16>
        <<<g, 256, 0, st>>>(dev_src, dst, static_cast<uint32_t>(n), static_cast<uint32_t>(h->d));
  h->stat_launches++;
  SGIC_CUDA(cudaGetLastError());
  return 0;
}

// Streams `n` host rows of `row_bytes` each through the pinned/device double buffer and
// calls `consume(dev_ptr, first_row_of_chunk, rows)` for each chunk.
template <typename F>
static int stream_rows_h2d(sgic_index* h, const uint8_t* host, int64_t n, size_t row_bytes, F consume) {
  int rc = ensure_staging(h);
  if (rc) return rc;
  const int64_t rows_per_chunk = std::max<int64_t>(1, static_cast<int64_t>(kStageChunkBytes / row_bytes));
  SGIC_REQUIRE(row_bytes <= kStageChunkBytes, "row too large for the staging buffer");
  int64_t done = 0;
  int b = 0;
  while (done < n) {
    const int64_t rows = std::min(rows_per_chunk, n - done);
    const size_t bytes = static_cast<size_t>(rows) * row_bytes;
    SGIC_CUDA(cudaEventSynchronize(h->ev[b]));  // previous use of this buffer pair has drained
    std::memcpy(h->pin[b], host + static_cast<size_t>(done) * row_bytes, bytes);
    SGIC_CUDA(cudaMemcpyAsync(h->dstage[b], h->pin[b], bytes, cudaMemcpyHostToDevice, h->stream));
    rc = consume(h->dstage[b], done, rows);
    if (rc) return rc;
    SGIC_CUDA(cudaEventRecord(h->ev[b], h->stream));
    done += rows;
    b ^= 1;
  }
  SGIC_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- K3 dispatch -------------------------------------------------------------------------
struct ScanCfg {
  int cpl, rb;
};
static ScanCfg scan_cfg_for_d(int d) {
  if (d <= 256) return {1, 8};
  if (d <= 512) return {2, 4};
  if (d <= 768) return {3, 4};
  if (d <= 1024) return {4, 2};
  return {8, 1};
}

template <typename T, int NQ, int CPL, int RB>
static cudaError_t launch_scan_inst(const ScanSmallParams& p, unsigned grid, size_t smem, cudaStream_t st) {
  auto kern = scan_small_kernel<T, NQ, CPL, RB>;
  static size_t configured[64] = {0};  // per device: the attribute lives in the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured[dev & 63] < smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured[dev & 63] = smem;
  }
  kern<<<grid, kScanThreads, smem, st>>>(p);
  return cudaGetLastError();
}

template <typename T, int NQ>
static cudaError_t launch_scan_nq(const ScanSmallParams& p, int cpl, int rb, unsigned grid, size_t smem,
                                  cudaStream_t st) {
  if constexpr (NQ == 1) {  // tile-height variants of the d<=512 batch-1 kernel (tuning knob "rb")
    if (cpl == 2 && rb == 2) return launch_scan_inst<T, NQ, 2, 2>(p, grid, smem, st);
    if (cpl == 2 && rb == 8) return launch_scan_inst<T, NQ, 2, 8>(p, grid, smem, st);
  }
  switch (cpl) {
    case 1: return launch_scan_inst<T, NQ, 1, 8>(p, grid, smem, st);
    case 2: return launch_scan_inst<T, NQ, 2, 4>(p, grid, smem, st);
    case 3: return launch_scan_inst<T, NQ, 3, 4>(p, grid, smem, st);
    case 4: return launch_scan_inst<T, NQ, 4, 2>(p, grid, smem, st);
    default: return launch_scan_inst<T, NQ, 8, 1>(p, grid, smem, st);
  }
}

template <typename T>
static cudaError_t launch_scan_t(const ScanSmallParams& p, int NQ, int cpl, int rb, unsigned grid, size_t smem,
                                 cudaStream_t st) {
  switch (NQ) {
    case 1: return launch_scan_nq<T, 1>(p, cpl, rb, grid, smem, st);
    case 2: return launch_scan_nq<T, 2>(p, cpl, rb, grid, smem, st);
    default: return launch_scan_nq<T, 4>(p, cpl, rb, grid, smem, st);
  }
}

// sorted_lists: every list is ordered best-first (the multiway merge needs that); the bitonic kernel does not care
static int launch_merge_keys(sgic_index* h, const uint64_t* partial, uint32_t nq, uint32_t n_lists, uint32_t k,
                             float* D, int64_t* I, int64_t id_base, cudaStream_t st, bool sorted_lists = true) {
  MergeKeysParams mp;
  mp.partial = partial;
  mp.n_lists = n_lists;
  mp.k = k;
  const uint32_t kp = next_pow2_u32(k);
  uint32_t chunk = std::max<uint32_t>(2 * kp, std::min<uint32_t>(16384, next_pow2_u32(n_lists * k)));
  chunk = std::max<uint32_t>(chunk, 64);
  mp.chunk = chunk;
  mp.D = D;
  mp.I = reinterpret_cast<long long*>(I);
  mp.id_base = id_base;
  const size_t all_bytes = static_cast<size_t>(n_lists) * k * 8;
  if (sorted_lists && n_lists <= 32u * kMergeMaxLpl && all_bytes <= 200u * 1024) {
    // every list fits in shared memory: multiway merge by one warp (k short rounds)
    if (all_bytes > 48 * 1024)
      SGIC_CUDA(cudaFuncSetAttribute(merge_keys_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(all_bytes)));
    merge_keys_small_kernel<<<nq, 128, all_bytes, st>>>(mp);
    h->stat_launches++;
    SGIC_CUDA(cudaGetLastError());
    return 0;
  }
  const size_t smem = static_cast<size_t>(chunk) * 8;
  if (smem > 48 * 1024)
    SGIC_CUDA(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  const int threads = static_cast<int>(std::min<uint32_t>(1024, std::max<uint32_t>(32, chunk / 2)));
  merge_keys_kernel<<<nq, threads, smem, st>>>(mp);
  h->stat_launches++;
  SGIC_CUDA(cudaGetLastError());
  return 0;
}

// ---- K4 dispatch (tcgen05 dense path) ------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Row-major [rows][d] 16-bit matrix, box = 64 elements (128 B) x box_rows, 128-byte swizzle,
// out-of-bounds elements read as zero (ragged last tile / d not a multiple of 64).
static int make_tmap_rows(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t d, uint32_t box_rows, int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  SGIC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {d, rows};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(d) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kDenseBK), box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, dtype == SGIC_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                         const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 2;
  }
  return 0;
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {
  while (b) {
    const uint32_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}

constexpr int kDenseBN = 256;
constexpr int64_t kDenseQueryBlock = 4096;  // queries per launch (FAISS blocks queries by 4096 too)

// Per-query start thresholds for the reservoir epilogue (k > 32): seed[q] = the order-preserving word just below the
// k-th best score of query q over a SAMPLE of the rows (0 = no bound).  At least k rows score above it, so every list
// of the query may start from it instead of -inf — the same protocol as DenseParams::gthr (ties stay admissible).
__global__ void seed_gthr_kernel(const float* D, const long long* I, uint32_t nq, uint32_t k, uint32_t* seed) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const size_t last = static_cast<size_t>(q) * k + (k - 1);
  seed[q] = I[last] >= 0 ? score_to_ord(D[last]) - 1u : 0u;
}

// rows_limit > 0: only the first rows_limit rows are searched (the sample pass below); such a call never seeds itself
static int search_dense_impl(sgic_index* h, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                             int64_t id_base, cudaStream_t st, int64_t rows_limit = 0) {
  using Cfg = DenseCfg<kDenseBN>;
  const uint32_t n_rows = static_cast<uint32_t>(rows_limit > 0 ? std::min<int64_t>(rows_limit, h->ntotal) : h->ntotal);
  // Sample pass.  With k > 32 every (query, slice) list fills from -inf: ~(C-k) ln(n/C)/ln(C/k) appends and a
  // compaction of every lane's reservoir in lockstep at the start of each item.  A search of the first rows of the
  // shard alone (0.5 % of the work at 10M rows) gives each query a bound that at least k rows beat; lists that start
  // from it skip most of the warm-up.  "dense_seed" = 0 switches it off (A/B); the answer is the same either way.
  const uint32_t* seed_all = nullptr;
  const bool huge_shard = ((static_cast<size_t>(n_rows) * h->d * 2) >> 20) >= static_cast<size_t>(h->opt_dense_b_min_mb);
  const bool seed_forced = h->opt_dense_seed == 2 && h->ntotal >= 8192;  // tests: the sample pass on small shards too
  if (rows_limit == 0 && k > 32 && h->opt_dense_seed && h->opt_dense_gthr &&
      ((nq >= 64 && h->ntotal >= (4ll << 20)) || seed_forced) &&
      h->opt_dense_mode != 4 && !(h->opt_dense_mode == 0 && huge_shard && nq > 256 && h->d <= 512)) {
    const int64_t sample = std::min<int64_t>(65536, std::max<int64_t>(1024, (h->ntotal / 8) & ~int64_t(255)));
    const size_t cand = static_cast<size_t>(nq) * k;
    const size_t need = cand * 12 + static_cast<size_t>(nq) * 4 + 64;
    if (need > h->seed_ws_bytes) {
      SGIC_CUDA(cudaStreamSynchronize(st));
      int rc0 = ensure_buf(&h->seed_ws, &h->seed_ws_bytes, need, false);
      if (rc0) return rc0;
    }
    int64_t* sI = reinterpret_cast<int64_t*>(h->seed_ws);
    float* sD = reinterpret_cast<float*>(static_cast<uint8_t*>(h->seed_ws) + cand * 8);
    uint32_t* seed = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(h->seed_ws) + cand * 12);
    int rc0 = search_dense_impl(h, nq, dev_q, k, sD, sI, 0, st, sample);
    if (rc0) return rc0;
    seed_gthr_kernel<<<static_cast<unsigned>((nq + 255) / 256), 256, 0, st>>>(sD, reinterpret_cast<const long long*>(sI),
                                                                              static_cast<uint32_t>(nq),
                                                                              static_cast<uint32_t>(k), seed);
    h->stat_launches++;
    SGIC_CUDA(cudaGetLastError());
    seed_all = seed;
  }
  const uint32_t kp = std::max<uint32_t>(2, next_pow2_u32(static_cast<uint32_t>(k)));
  // k <= 32: thread-private sorted lists in shared memory (k KB per CTA); larger k: per-query reservoirs of
  // res_cap keys in an L2-resident workspace, compacted by the warp when full (lists come out unsorted)
  const bool tp = k <= 32;
  const uint32_t res_cap = std::max<uint32_t>(256, 2 * kp);
  const size_t epi_bytes = dense_epi_bytes(static_cast<uint32_t>(k), tp);
  // the three kernels' shared memory: alignment slack + operand ring (+ resident query tile) + epilogue + barriers
  const size_t smem_1cta = 1024 + Cfg::kStageBytes + epi_bytes + 256;  // at least one full stage (ring sized below)
  const size_t smem_2r = 1024 + static_cast<size_t>(kD2MaxKc + 4) * kD2HalfBytes + epi_bytes + 256;
  const size_t smem_2s = 1024 + static_cast<size_t>(6) * 2 * kD2HalfBytes + epi_bytes + 256;
  SGIC_REQUIRE(std::max(smem_1cta, std::max(smem_2r, smem_2s)) <= kSmemBudget,
               "dense path: shared memory budget exceeded");
  auto kern = scan_dense_kernel<kDenseBN>;
  auto kern2r = scan_dense2_kernel<true, 4>;   // CTA pairs, query tile resident (d <= 512)
  auto kern2s = scan_dense2_kernel<false, 6>;  // CTA pairs, query tile streamed with the database
  static bool configured[64] = {false};
  if (!configured[h->device & 63]) {
    SGIC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
    SGIC_CUDA(cudaFuncSetAttribute(kern2r, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
    SGIC_CUDA(cudaFuncSetAttribute(kern2s, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
    configured[h->device & 63] = true;
  }
  // cta_group::2 (CTA pairs, 256 queries per tile) for batches of more than 128 queries; up to 128 queries
  // fit one CTA's TMEM lanes and the scan is HBM-bound: single-CTA kernel with the query box cut to the
  // batch.  "dense_mode" = 1 forces the single-CTA kernel, 3 forces pairs (A/B comparisons).
  const bool pairs = h->sm_count >= 2 && h->opt_dense_mode != 1 && (nq > kDenseBM || h->opt_dense_mode == 3);
  const bool a_resident = h->d <= kD2MaxKc * kDenseBK && h->opt_dense_mode != 2;
  const uint32_t q_tile = pairs ? 256u : static_cast<uint32_t>(kDenseBM);
  const uint32_t n_units = pairs ? static_cast<uint32_t>(h->sm_count) / 2 : static_cast<uint32_t>(h->sm_count);
  const uint32_t n_tiles = (n_rows + kDenseBN - 1) / kDenseBN;

  // queries -> storage dtype, once for the whole call
  const size_t qh_need = static_cast<size_t>(nq) * h->d * 2;
  if (qh_need > h->qh_bytes) {
    SGIC_CUDA(cudaStreamSynchronize(st));
    int rc = ensure_buf(&h->qh, &h->qh_bytes, qh_need, false);
    if (rc) return rc;
  }
  {
    const size_t n8 = static_cast<size_t>(nq) * h->d / 8;
    const unsigned g = grid_for(n8, 256, h->sm_count);
    if (h->dtype == SGIC_F16) pack_f32_kernel<__half><<<g, 256, 0, st>>>(dev_q, h->qh, n8);
    else pack_f32_kernel<__nv_bfloat16><<<g, 256, 0, st>>>(dev_q, h->qh, n8);
    h->stat_launches++;
    SGIC_CUDA(cudaGetLastError());
  }
  CUtensorMap tm_db;
  int rc = make_tmap_rows(&tm_db, h->db, n_rows, static_cast<uint32_t>(h->d), pairs ? kDenseBM : kDenseBN, h->dtype);
  if (rc) return rc;

  if (h->opt_timing == 1) SGIC_CUDA(cudaEventRecord(h->t0, st));
  for (int64_t q0 = 0; q0 < nq; q0 += kDenseQueryBlock) {
    const uint32_t nqb = static_cast<uint32_t>(std::min<int64_t>(kDenseQueryBlock, nq - q0));
    const uint32_t m_tiles = (nqb + q_tile - 1) / q_tile;
    // Slices.  Default: the fewest slices that make the items a multiple of the CTAs (pairs) — every CTA gets
    // the same number of equal items and visits each of its query tiles once.  "dense_l2_mb" > 0 switches to
    // SHORT slices sized so that the ~units/m_tiles slices in flight fit that L2 budget (a CTA then resumes a
    // query tile's list many times, see scan_dense.cuh).  Measured on 30M x 512, nq = 4096 (profiles/): short
    // slices read MORE from DRAM (144.8 GB vs 84.2 GB; L2 hit 64.6 % vs 77.6 %) because the CTAs drift over
    // thousands of items without a lockstep, and are 5 % slower — so it is off by default.
    const uint32_t step = n_units / gcd_u32(m_tiles, n_units);
    uint32_t n_slices = std::max<uint32_t>(1, std::min(step, n_tiles));
    // Reservoir epilogue (k > 32): every (query, slice) list pays its own warm-up — about (C-k) ln(n/C)/ln(C/k)
    // appends, ~1150 at k = 100 — and the epilogue is the bottleneck (10M x 512, k = 100, nq = 4096: 46 ms against
    // 33 ms with the epilogue off).  Fewer, longer slices cut that work in proportion; the price is an item count
    // that no longer divides the CTAs.  Pick the slice count that minimises (rounds * units / items) * (1 + 0.008
    // slices): 16 query tiles on 74 pairs -> 9 slices (144 items, 2 rounds, 2.7 % idle) instead of 37.
    // ("dense_slices" > 0 forces a slice count.  Measured on C3 with the sample pass on, one box: 9 slices 60.5 ms,
    //  18: 61.5, 37: 62.6, 74: 64.6 — shorter slices do not pay although they re-align the pairs that share a slice.)
    if (!tp && m_tiles * n_slices > n_units && h->opt_dense_slices == 0) {
      double best_cost = 1e30;
      uint32_t best = n_slices;
      for (uint32_t ns = 1; ns <= std::min(step, n_tiles); ++ns) {
        const uint32_t items = m_tiles * ns;
        if (items < n_units && ns < std::min(step, n_tiles)) continue;  // at least one item per CTA (pair)
        const uint32_t rounds = (items + n_units - 1) / n_units;
        const double cost = static_cast<double>(rounds) * n_units / items * (1.0 + 0.008 * ns);
        if (cost < best_cost - 1e-12) {
          best_cost = cost;
          best = ns;
        }
      }
      n_slices = best;
    }
    if (h->opt_dense_slices > 0) n_slices = std::max<uint32_t>(1, std::min<uint32_t>(static_cast<uint32_t>(h->opt_dense_slices), n_tiles));
    const size_t tile_bytes = static_cast<size_t>(kDenseBN) * h->d * 2;
    const size_t l2_budget = static_cast<size_t>(h->opt_dense_l2_mb) << 20;
    if (l2_budget > 0 && m_tiles > 1 && static_cast<size_t>(n_tiles) * tile_bytes > l2_budget) {
      const uint32_t in_flight = (n_units + m_tiles - 1) / m_tiles + 1;
      const uint32_t tps0 = std::max<uint32_t>(8, static_cast<uint32_t>(l2_budget / (in_flight * tile_bytes)));
      n_slices = (n_tiles + tps0 - 1) / tps0;
      n_slices = ((n_slices + step - 1) / step) * step;  // items stay a multiple of the CTAs
      n_slices = std::max<uint32_t>(1, std::min(n_slices, n_tiles));
    }
    const uint32_t tiles_per_slice = (n_tiles + n_slices - 1) / n_slices;
    n_slices = (n_tiles + tiles_per_slice - 1) / tiles_per_slice;
    const uint32_t n_items = m_tiles * n_slices;
    const uint32_t grid_units = std::min<uint32_t>(n_units, n_items);
    const uint32_t grid = grid_units * (pairs ? 2u : 1u);

    // default schedule: every (query tile, slice) item is visited exactly once -> one list per slice; the
    // short-slice schedule revisits query tiles -> one list per CTA (pair), parked and resumed
    const bool by_slice = n_slices <= grid_units;
    const uint32_t n_lists = by_slice ? n_slices : grid_units;
    const size_t partial_bytes = static_cast<size_t>(nqb) * std::max(n_lists, grid_units) * static_cast<size_t>(k) * 8;
    const size_t gthr_off = (partial_bytes + 15) & ~size_t(15), gthr_bytes = static_cast<size_t>(nqb) * 4;
    if (gthr_off + gthr_bytes + 16 > h->ws_bytes) {
      SGIC_CUDA(cudaStreamSynchronize(st));
      rc = ensure_buf(&h->ws, &h->ws_bytes, gthr_off + gthr_bytes + 16, false);
      if (rc) return rc;
      h->ws_counter = nullptr;
    }
    if (!tp) {
      const size_t need = static_cast<size_t>(grid) * kDenseBM * res_cap * 8;
      if (need > h->lists_ws_bytes) {
        SGIC_CUDA(cudaStreamSynchronize(st));
        rc = ensure_buf(&h->lists_ws, &h->lists_ws_bytes, need, false);
        if (rc) return rc;
      }
    }
    // Several query tiles over a large shard: the pair kernel that keeps a DATABASE tile resident and streams the
    // query block past it — the database leaves DRAM exactly once per block (scan_dense.cuh).  Below
    // "dense_b_min_mb" MB of database the query-resident kernel's slices still meet in L2 and its in-register /
    // shared-memory lists are cheaper than reservoirs.  "dense_mode" = 4 forces it wherever it is eligible.
    {
      const uint32_t cap_b = k <= 16 ? 32u : k <= 32 ? 64u : res_cap;
      const uint32_t units_b = std::min<uint32_t>(n_units, n_tiles);
      const size_t ws_b = static_cast<size_t>(units_b) * 2 * m_tiles * kDenseBM * cap_b * 8;
      const size_t db_mb = (static_cast<size_t>(n_rows) * h->d * 2) >> 20;
      const bool eligible = pairs && h->d <= kD2MaxKc * kDenseBK && m_tiles >= 2 &&
                            dense_b_smem_bytes(m_tiles) <= kSmemBudget && ws_b <= (size_t(1) << 30);
      if (eligible && (h->opt_dense_mode == 4 ||
                       (h->opt_dense_mode == 0 && db_mb >= static_cast<size_t>(h->opt_dense_b_min_mb)))) {
        const size_t pbytes = static_cast<size_t>(nqb) * units_b * static_cast<size_t>(k) * 8;
        if (pbytes + 16 > h->ws_bytes) {
          SGIC_CUDA(cudaStreamSynchronize(st));
          rc = ensure_buf(&h->ws, &h->ws_bytes, pbytes + 16, false);
          if (rc) return rc;
        }
        if (ws_b > h->lists_ws_bytes) {
          SGIC_CUDA(cudaStreamSynchronize(st));
          rc = ensure_buf(&h->lists_ws, &h->lists_ws_bytes, ws_b, false);
          if (rc) return rc;
        }
        h->ws_counter = nullptr;  // the workspace is overwritten: K3 must re-zero its counter
        CUtensorMap tm_qb;
        rc = make_tmap_rows(&tm_qb, static_cast<const uint8_t*>(h->qh) + static_cast<size_t>(q0) * h->d * 2, nqb,
                            static_cast<uint32_t>(h->d), static_cast<uint32_t>(kDenseBM), h->dtype);
        if (rc) return rc;
        DenseBParams bp;
        bp.partial = static_cast<uint64_t*>(h->ws);
        bp.lists_ws = static_cast<uint64_t*>(h->lists_ws);
        bp.n_rows = n_rows;
        bp.nq = nqb;
        bp.k = static_cast<uint32_t>(k);
        bp.res_cap = cap_b;
        bp.m_tiles = m_tiles;
        bp.n_tiles = n_tiles;
        bp.n_lists = units_b;
        bp.kc = (static_cast<uint32_t>(h->d) + kDenseBK - 1) / kDenseBK;
        bp.idesc = ptx::umma_idesc_f16(256, kDenseBN, h->dtype == SGIC_BF16 ? 1u : 0u);
        bp.debug = static_cast<uint32_t>(h->opt_debug);
        bp.epi_wait_ns = static_cast<uint32_t>(h->opt_epi_wait_ns);
        static bool b_configured[64] = {false};
        if (!b_configured[h->device & 63]) {
          SGIC_CUDA(cudaFuncSetAttribute(scan_dense2b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSmemBudget)));
          b_configured[h->device & 63] = true;
        }
        if (q0 == 0 && (rc = ring_mark(h, true, st))) return rc;
        scan_dense2b_kernel<<<units_b * 2, kDenseThreads, dense_b_smem_bytes(m_tiles), st>>>(tm_qb, tm_db, bp);
        h->stat_launches++;
        SGIC_CUDA(cudaGetLastError());
        if (q0 == 0 && (rc = ring_mark(h, false, st))) return rc;
        h->stat_last_grid = units_b * 2;
        h->stat_last_stages = kD2bStages;
        h->stat_last_kernel = 5;
        if (h->opt_timing == 1 && q0 == 0) SGIC_CUDA(cudaEventRecord(h->tm, st));
        rc = launch_merge_keys(h, static_cast<const uint64_t*>(h->ws), nqb, units_b, static_cast<uint32_t>(k),
                               dev_D + static_cast<size_t>(q0) * k, dev_I + static_cast<size_t>(q0) * k, id_base, st,
                               /*sorted_lists=*/cap_b <= 64u);
        if (rc) return rc;
        continue;
      }
    }
    CUtensorMap tm_q;
    // Small batches: the transposed kernel (database rows on M, queries on N), as long as the resident query
    // matrix + the (warp, query) lists leave room for a ring of at least 3 database stages.
    // "dense_mode" = 1 keeps the queries-on-M single-CTA kernel (A/B comparisons).
    const uint32_t kc_chunks = (static_cast<uint32_t>(h->d) + kDenseBK - 1) / kDenseBK;
    const uint32_t n_pad = (nqb + 15u) & ~15u;
    const uint32_t n_mma = (nqb <= 8u && h->opt_t_n8) ? 8u : n_pad;  // tensor work proportional to the batch down to N = 8
    uint32_t t_stages = 0;
    // Measured per-query cost on top of the stream (50M x 512, 1000 W cap, profiles/r02_mid_batches.log): K4t 29-33 us
    // (its MMAs are narrow: N = the batch), the queries-on-M kernel 20.5 us at a FULL 128-query tile — from ~88
    // queries the padded full tile is the cheaper one (9.5 ms flat against 6.9 + 0.031 nq).  "t_max_nq" overrides.
    if (!pairs && h->opt_dense_mode != 1 && nqb <= static_cast<uint32_t>(h->opt_t_max_nq)) {
      const size_t fixed = dense_t_fixed_bytes(n_pad, kc_chunks, static_cast<uint32_t>(k)) + 1024 + 256;
      // measured at 100M x 512, nq = 8: 3-4 stages (96-128 KB in flight) 14.2 ms, 5 stages 15.2 ms, 6 stages 15.5 ms
      if (fixed < kSmemBudget) t_stages = static_cast<uint32_t>(std::min<size_t>((kSmemBudget - fixed) / kDtStageBytes, 4));
      if (h->opt_stages > 0) t_stages = std::min<uint32_t>(t_stages, static_cast<uint32_t>(h->opt_stages));
      if (t_stages < 3) t_stages = 0;
    }
    const bool transposed = t_stages > 0;
    rc = make_tmap_rows(&tm_q, static_cast<const uint8_t*>(h->qh) + static_cast<size_t>(q0) * h->d * 2, nqb,
                        static_cast<uint32_t>(h->d),
                        transposed ? n_mma
                                   : pairs ? static_cast<uint32_t>(kDenseBM) : std::min<uint32_t>(kDenseBM, (nqb + 7u) & ~7u),
                        h->dtype);
    if (rc) return rc;
    DenseParams p;
    p.partial = static_cast<uint64_t*>(h->ws);
    p.lists_ws = static_cast<uint64_t*>(h->lists_ws);
    p.n_rows = n_rows;
    p.nq = nqb;
    p.k = static_cast<uint32_t>(k);
    p.res_cap = res_cap;
    p.m_tiles = m_tiles;
    p.n_slices = n_slices;
    p.tiles_per_slice = tiles_per_slice;
    p.n_tiles = n_tiles;
    p.n_lists = n_lists;
    p.by_slice = by_slice ? 1u : 0u;
    // items of one query tile that run one after the other inherit each other's bounds ("dense_gthr" = 0: off)
    p.gthr = (by_slice && (n_slices > 1 || seed_all != nullptr) && h->opt_dense_gthr)
                 ? reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(h->ws) + gthr_off) : nullptr;
    p.kc = (static_cast<uint32_t>(h->d) + kDenseBK - 1) / kDenseBK;
    p.tp = tp ? 1u : 0u;
    p.a_rows = pairs ? static_cast<uint32_t>(kDenseBM) : std::min<uint32_t>(kDenseBM, (nqb + 7u) & ~7u);
    // single-CTA ring: the query box only reserves what the batch needs, the rest of shared memory goes to
    // database stages ("stages" option overrides the depth for experiments)
    p.a_region = (p.a_rows * static_cast<uint32_t>(kDenseBK * 2) + 1023u) & ~1023u;
    {
      const size_t per_stage = p.a_region + Cfg::kBBytes;
      size_t ns = (kSmemBudget - 1024 - epi_bytes - 256) / per_stage;
      ns = std::min<size_t>(ns, kDenseMaxStages);
      if (h->opt_stages > 0) ns = std::min<size_t>(ns, static_cast<size_t>(h->opt_stages));
      p.n_stages = static_cast<uint32_t>(std::max<size_t>(ns, 1));
    }
    p.idesc = ptx::umma_idesc_f16(q_tile, kDenseBN, h->dtype == SGIC_BF16 ? 1u : 0u);
    p.db_evict_first = (m_tiles == 1 && h->opt_evict_first) ? 1u : 0u;
    p.debug = static_cast<uint32_t>(h->opt_debug);
    p.epi_wait_ns = static_cast<uint32_t>(h->opt_epi_wait_ns);
    h->ws_counter = nullptr;  // this launch overwrites the workspace: K3 must re-zero its "CTAs done" counter
    // empty lists where a CTA resumes its own slot; the by-slice schedule writes every list in full
    if (transposed || !by_slice) SGIC_CUDA(cudaMemsetAsync(h->ws, 0, partial_bytes, st));
    if (!transposed && p.gthr) {
      if (seed_all) SGIC_CUDA(cudaMemcpyAsync(p.gthr, seed_all + q0, gthr_bytes, cudaMemcpyDeviceToDevice, st));
      else SGIC_CUDA(cudaMemsetAsync(p.gthr, 0, gthr_bytes, st));
    }
    if (q0 == 0 && (rc = ring_mark(h, true, st))) return rc;
    if (transposed) {
      DenseTParams tp_;
      tp_.partial = p.partial;
      tp_.n_rows = n_rows;
      tp_.nq = nqb;
      tp_.k = static_cast<uint32_t>(k);
      tp_.n_pad = n_pad;
      tp_.n_mma = n_mma;
      tp_.n_slices = n_slices;  // m_tiles == 1: one slice per CTA
      tp_.tiles_per_slice = tiles_per_slice;
      tp_.n_tiles = n_tiles;
      tp_.n_lists = grid_units;
      tp_.kc = kc_chunks;
      tp_.n_stages = t_stages;
      tp_.idesc = ptx::umma_idesc_f16(128, n_mma, h->dtype == SGIC_BF16 ? 1u : 0u);
      tp_.db_evict_first = p.db_evict_first;
      tp_.debug = p.debug;
      tp_.epi_wait_ns = p.epi_wait_ns;
      static bool t_configured[64] = {false};
      if (!t_configured[h->device & 63]) {
        SGIC_CUDA(cudaFuncSetAttribute(scan_dense_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kSmemBudget)));
        t_configured[h->device & 63] = true;
      }
      scan_dense_t_kernel<<<grid, kDenseThreads, kSmemBudget, st>>>(tm_q, tm_db, tp_);
    } else if (!pairs) kern<<<grid, kDenseThreads, kSmemBudget, st>>>(tm_q, tm_db, p);
    else if (a_resident) kern2r<<<grid, kDenseThreads, kSmemBudget, st>>>(tm_q, tm_db, p);
    else kern2s<<<grid, kDenseThreads, kSmemBudget, st>>>(tm_q, tm_db, p);
    h->stat_launches++;
    SGIC_CUDA(cudaGetLastError());
    h->stat_last_grid = grid;
    h->stat_last_stages = transposed ? t_stages : pairs ? 4 : p.n_stages;
    h->stat_last_kernel = transposed ? 2 : !pairs ? 1 : a_resident ? 3 : 4;
    if (q0 == 0 && (rc = ring_mark(h, false, st))) return rc;
    if (h->opt_timing == 1 && q0 == 0) SGIC_CUDA(cudaEventRecord(h->tm, st));
    rc = launch_merge_keys(h, static_cast<const uint64_t*>(h->ws), nqb, transposed ? grid_units : n_lists,
                           static_cast<uint32_t>(k), dev_D + static_cast<size_t>(q0) * k,
                           dev_I + static_cast<size_t>(q0) * k, id_base, st, /*sorted_lists=*/tp || transposed);
    if (rc) return rc;
  }
  if (h->opt_timing == 1) {
    SGIC_CUDA(cudaEventRecord(h->t1, st));
    SGIC_CUDA(cudaEventSynchronize(h->t1));
    float ms = 0.f;
    SGIC_CUDA(cudaEventElapsedTime(&ms, h->t0, h->t1));
    h->stat_last_search_ns = static_cast<int64_t>(static_cast<double>(ms) * 1e6);
    SGIC_CUDA(cudaEventElapsedTime(&ms, h->t0, h->tm));
    h->stat_last_scan_ns = static_cast<int64_t>(static_cast<double>(ms) * 1e6);
  }
  return 0;
}

// Where a search's answer goes when it is one shard's part of a multi-GPU search: `n` gather slots (in the home GPU's
// or every peer's memory) and their flags.  A search that can fold the push into its final writer (K3's last CTA) does
// so and reports *pushed = true; otherwise it writes (dev_D, dev_I) as usual and the caller pushes.
struct XDst {
  uint32_t n = 0, epoch = 0;
  float* D[16];
  long long* I[16];
  uint32_t* F[16];
};

// host_q_inline != nullptr (and dev_q == nullptr): the queries are taken from host memory and travel in the kernel
// parameters (nq * d <= kScanInlineQ)
static int search_small_impl(sgic_index* h, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                             int64_t id_base, cudaStream_t st, const uint64_t* bound = nullptr,
                             const float* host_q_inline = nullptr, const XDst* xd = nullptr, bool* pushed = nullptr);

// Smallest batch that goes to the tensor-core kernels ("dense_min_nq" overrides).  K3 owns lanes by 256-element
// slabs of a row (8 elements per lane and chunk): at d = 512 / 768 / 1024 / 2048 every lane works and one query
// streams at 7.4 TB/s, but narrow rows (d <= 128: 5.4 TB/s, d = 64: 2.7 TB/s) and widths that fill the last
// template instance badly (d = 1280: 6.1 TB/s) leave lanes idle — there even a single query is better off on the
// transposed kernel (7.1-7.3 TB/s at every width, profiles/r01_dims.log) once the shard is large enough to hide
// its extra launches.
static int64_t auto_dense_min_nq(const sgic_index* h) {
  if (h->opt_dense_min_nq != 0) return h->opt_dense_min_nq;
  const ScanCfg cfg = scan_cfg_for_d(h->d);
  const double lane_util = static_cast<double>(h->d) / (cfg.cpl * 256.0);
  if (h->ntotal >= (2ll << 20) && lane_util < 0.7) return 1;
  return (h->ntotal >= (8ll << 20)) ? 2 : 3;
}

// Searches `nq` device-resident fp32 queries; results to device buffers.  Regime choice: a few
// queries -> K3 (CUDA-core streaming scan, HBM-bound); batches -> K4 (tcgen05 dense contraction).
static int search_dev_impl(sgic_index* h, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                           int64_t id_base, cudaStream_t st, const XDst* xd = nullptr, bool* pushed = nullptr) {
  if (pushed) *pushed = false;
  SGIC_REQUIRE(k >= 1, "k must be >= 1");
  SGIC_REQUIRE(nq >= 0, "nq must be >= 0");
  if (nq == 0) return 0;
  SGIC_REQUIRE(h->ntotal < (1ll << 32) - 1, "more than 2^32-2 rows in one shard");
  if (k > kMaxK) {
    // FAISS takes any k (do_search asks for min(topk, ntotal), src/search.py:114).  Beyond 1024 the answer is
    // enumerated in passes of <= 1024 by K3, one query at a time: pass p only admits keys strictly below the last
    // key of pass p-1 (the bound stays on the device), so the concatenation is the exact ordered top-k.
    SGIC_REQUIRE(k <= (1ll << 24), "k too large");
    const int64_t valid = std::min<int64_t>(k, h->ntotal);
    if (32 > h->ws2_bytes) {
      SGIC_CUDA(cudaStreamSynchronize(st));
      int rc = ensure_buf(&h->ws2, &h->ws2_bytes, 32, false);
      if (rc) return rc;
    }
    uint64_t* bound = static_cast<uint64_t*>(h->ws2);
    for (int64_t q = 0; q < nq; ++q) {
      float* Dq = dev_D + q * k;
      int64_t* Iq = dev_I + q * k;
      for (int64_t got = 0; got < valid; got += kMaxK) {
        const int64_t kk = std::min<int64_t>(kMaxK, valid - got);
        int rc = search_small_impl(h, 1, dev_q + q * h->d, kk, Dq + got, Iq + got, id_base, st, got ? bound : nullptr);
        if (rc) return rc;
        if (got + kk < valid)
          next_bound_kernel<<<1, 1, 0, st>>>(Dq + got, reinterpret_cast<const long long*>(Iq + got),
                                             static_cast<uint32_t>(kk - 1), id_base, bound);
      }
      if (valid < k) {
        const uint32_t pad = static_cast<uint32_t>(k - valid);
        fill_padding_kernel<<<(pad + 255) / 256, 256, 0, st>>>(Dq + valid, reinterpret_cast<long long*>(Iq + valid), pad);
      }
      SGIC_CUDA(cudaGetLastError());
    }
    return 0;
  }
  // Regime choice (measured, 512-d fp16): one query -> K3.  Two queries -> K3 on small shards (one launch), the
  // transposed tensor-core kernel from ~8M rows (100M rows: 14.2 ms against K3's 15.8 ms).  Three or more ->
  // tensor cores always (K3 with 4 queries is FMA-bound: 18.5 ms at 100M rows against 15.4 ms).
  const int64_t min_nq = auto_dense_min_nq(h);
  if (h->ntotal > 0 && nq >= min_nq) return search_dense_impl(h, nq, dev_q, k, dev_D, dev_I, id_base, st);
  return search_small_impl(h, nq, dev_q, k, dev_D, dev_I, id_base, st, nullptr, nullptr, xd, pushed);
}

// The same for queries in HOST memory.  One or two queries in the streaming regime ride in the kernel parameters (no
// copy at all); anything else is staged through pinned memory (`pinned`: host_q already is pinned) and copied.
static int search_hostq_impl(sgic_index* h, int64_t nq, const float* host_q, bool pinned, int64_t k, float* dev_D,
                             int64_t* dev_I, int64_t id_base, cudaStream_t st, const XDst* xd = nullptr,
                             bool* pushed = nullptr) {
  if (pushed) *pushed = false;
  SGIC_REQUIRE(k >= 1, "k must be >= 1");
  SGIC_REQUIRE(nq >= 0, "nq must be >= 0");
  if (nq == 0) return 0;
  const bool small = h->ntotal == 0 || nq < auto_dense_min_nq(h);
  if (small && k <= kMaxK && nq * h->d <= kScanInlineQ)
    return search_small_impl(h, nq, nullptr, k, dev_D, dev_I, id_base, st, nullptr, host_q, xd, pushed);
  const size_t qbytes = static_cast<size_t>(nq) * h->d * 4;
  int rc = ensure_buf(&h->qdev, &h->qdev_bytes, qbytes, false);
  if (rc) return rc;
  const void* src = host_q;
  if (!pinned) {
    if ((rc = ensure_buf(&h->qpin, &h->qpin_bytes, qbytes, true))) return rc;
    std::memcpy(h->qpin, host_q, qbytes);
    src = h->qpin;
  }
  SGIC_CUDA(cudaMemcpyAsync(h->qdev, src, qbytes, cudaMemcpyHostToDevice, st));
  return search_dev_impl(h, nq, static_cast<const float*>(h->qdev), k, dev_D, dev_I, id_base, st, xd, pushed);
}

static int search_small_impl(sgic_index* h, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                             int64_t id_base, cudaStream_t st, const uint64_t* bound, const float* host_q_inline,
                             const XDst* xd, bool* pushed) {
  if (pushed) *pushed = false;
  SGIC_REQUIRE(k >= 1, "k must be >= 1");
  SGIC_REQUIRE(k <= kMaxK, "k > 1024 is not supported by this build");
  SGIC_REQUIRE(nq >= 0, "nq must be >= 0");
  if (nq == 0) return 0;
  SGIC_REQUIRE(h->ntotal < (1ll << 32) - 1, "more than 2^32-2 rows in one shard");
  ScanCfg cfg = scan_cfg_for_d(h->d);
  if (nq == 1 && cfg.cpl == 2 && (h->opt_rb == 2 || h->opt_rb == 8)) cfg.rb = static_cast<int>(h->opt_rb);
  const uint32_t kp = std::max<uint32_t>(2, next_pow2_u32(static_cast<uint32_t>(k)));
  int NQ = (nq == 1) ? 1 : (nq == 2) ? 2 : 4;
  while (NQ > 1 && static_cast<uint32_t>(NQ) * kp > 1024) NQ >>= 1;

  const uint32_t R = kScanConsumerWarps * cfg.rb;
  const uint32_t stage_bytes = ((R * h->d * 2u) + 127u) & ~127u;
  const size_t list_bytes = static_cast<size_t>(NQ) * kScanConsumerWarps * kp * 8;
  // full / empty barriers + the stages' tile numbers + the staged queries (fp32)
  const size_t bar_bytes = 2 * kScanMaxStages * 8 + kScanMaxStages * 4 + static_cast<size_t>(NQ) * h->d * 4;
  SGIC_REQUIRE(list_bytes + bar_bytes + 2 * stage_bytes <= kSmemBudget, "k too large for shared memory");
  // Measured on B200 (profiles/r01_k3_stage_sweep.md): ~128 KB of bulk copies in flight per SM is
  // the sweet spot (7.4-7.5 TB/s); 160 KB and more drops to ~6.8 TB/s, 64 KB to ~6.4 TB/s.
  const uint32_t max_stages =
      std::min<uint32_t>(kScanMaxStages, static_cast<uint32_t>((kSmemBudget - list_bytes - bar_bytes) / stage_bytes));
  uint32_t stages = std::max<uint32_t>(2, ((128u << 10) + stage_bytes / 2) / stage_bytes);
  if (h->opt_stages > 0) stages = static_cast<uint32_t>(h->opt_stages);
  stages = std::max<uint32_t>(2, std::min(stages, max_stages));
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + list_bytes + bar_bytes;

  const uint32_t n_rows = static_cast<uint32_t>(h->ntotal);
  const uint32_t n_tiles = (n_rows + R - 1) / R;
  uint32_t grid = static_cast<uint32_t>(h->opt_grid > 0 ? h->opt_grid : h->sm_count);
  grid = std::max<uint32_t>(1, std::min(grid, std::max<uint32_t>(n_tiles, 1)));
  h->stat_last_grid = grid;
  h->stat_last_stages = stages;
  h->stat_last_kernel = 0;

  // fused grid-level merge (last CTA) when all partial lists fit in the freed stage ring
  const size_t partial_bytes = static_cast<size_t>(NQ) * grid * static_cast<size_t>(k) * 8;
  const bool fused = h->opt_fused && n_rows > 0 && grid <= 32u * kMergeMaxLpl &&
                     partial_bytes <= static_cast<size_t>(stages) * stage_bytes && NQ <= kScanConsumerWarps;
  const size_t ws_need = partial_bytes + 16;  // + the "CTAs done" counter behind the lists
  if (ws_need > h->ws_bytes) {
    SGIC_CUDA(cudaStreamSynchronize(st));
    int rc = ensure_buf(&h->ws, &h->ws_bytes, ws_need, false);
    if (rc) return rc;
    h->ws_counter = nullptr;
  }
  uint32_t* counter = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(h->ws) + ((partial_bytes + 7) & ~size_t(7)));
  if (h->ws_counter != counter) {  // (re)allocated or a different layout: zero the counter once
    SGIC_CUDA(cudaMemsetAsync(counter, 0, 8, st));
    h->ws_counter = counter;
  }

  if (h->opt_timing == 1) SGIC_CUDA(cudaEventRecord(h->t0, st));
  for (int64_t q0 = 0; q0 < nq; q0 += NQ) {
    const uint32_t nq_here = static_cast<uint32_t>(std::min<int64_t>(NQ, nq - q0));
    ScanSmallParams p;
    p.db = h->db;
    if (dev_q != nullptr) {
      p.q = dev_q + static_cast<size_t>(q0) * h->d;
    } else {
      p.q = nullptr;
      std::memcpy(p.q_inline, host_q_inline + static_cast<size_t>(q0) * h->d, static_cast<size_t>(nq_here) * h->d * 4);
    }
    p.partial = static_cast<uint64_t*>(h->ws);
    p.n_rows = n_rows;
    p.d = static_cast<uint32_t>(h->d);
    p.nq = nq_here;
    p.k = static_cast<uint32_t>(k);
    p.kp = kp;
    p.n_tiles = n_tiles;
    p.n_stages = stages;
    p.stage_bytes = stage_bytes;
    p.evict_first = h->opt_evict_first ? 1u : 0u;
    p.wait_ns = static_cast<uint32_t>(h->opt_scan_wait_ns);
    p.fused = fused ? 1u : 0u;
    p.counter = counter;
    // equal round-robin shares for the first 7/8 of the tiles, the rest is claimed dynamically ("steal" = 0: all static)
    p.n_static = h->opt_steal ? static_cast<uint32_t>((static_cast<uint64_t>(n_tiles) * 7 / 8) / grid) * grid : n_tiles;
    p.steal = counter + 1;
    if (!fused && n_rows > 0) SGIC_CUDA(cudaMemsetAsync(p.steal, 0, 4, st));  // nobody resets it in the kernel then
    p.D = dev_D + static_cast<size_t>(q0) * k;
    p.I = reinterpret_cast<long long*>(dev_I + static_cast<size_t>(q0) * k);
    p.id_base = id_base;
    p.bound = bound;
    p.x_n = 0;
    p.x_epoch = 0;
    if (xd != nullptr && xd->n > 0 && fused && nq <= NQ && n_rows > 0) {  // one launch answers the whole batch: fold the push
      p.x_n = xd->n;
      p.x_epoch = xd->epoch;
      for (uint32_t j = 0; j < xd->n; ++j) {
        p.xD[j] = xd->D[j];
        p.xI[j] = xd->I[j];
        p.xF[j] = xd->F[j];
      }
      if (pushed) *pushed = true;
    }
    p.trace = reinterpret_cast<unsigned long long*>(static_cast<uintptr_t>(h->opt_trace));
    cudaError_t e;
    if (q0 == 0) {
      int rrc = ring_mark(h, true, st);
      if (rrc) return rrc;
    }
    if (n_rows > 0) {
      if (h->dtype == SGIC_F16) e = launch_scan_t<__half>(p, NQ, cfg.cpl, cfg.rb, grid, smem, st);
      else e = launch_scan_t<__nv_bfloat16>(p, NQ, cfg.cpl, cfg.rb, grid, smem, st);
      h->stat_launches++;
      if (e != cudaSuccess) {
        set_error(std::string("scan_small launch failed: ") + cudaGetErrorString(e));
        return 2;
      }
    } else {
      SGIC_CUDA(cudaMemsetAsync(h->ws, 0, static_cast<size_t>(NQ) * grid * k * 8, st));
    }
    if (q0 == 0) {
      int rrc = ring_mark(h, false, st);
      if (rrc) return rrc;
    }
    if (h->opt_timing == 1 && q0 == 0) SGIC_CUDA(cudaEventRecord(h->tm, st));  // first scan launch alone
    if (fused) continue;
    int rc = launch_merge_keys(h, static_cast<const uint64_t*>(h->ws), nq_here, grid, static_cast<uint32_t>(k),
                               dev_D + static_cast<size_t>(q0) * k, dev_I + static_cast<size_t>(q0) * k, id_base, st);
    if (rc) return rc;
  }
  if (h->opt_timing == 1) {
    SGIC_CUDA(cudaEventRecord(h->t1, st));
    SGIC_CUDA(cudaEventSynchronize(h->t1));
    float ms = 0.f;
    SGIC_CUDA(cudaEventElapsedTime(&ms, h->t0, h->t1));
    h->stat_last_search_ns = static_cast<int64_t>(static_cast<double>(ms) * 1e6);
    SGIC_CUDA(cudaEventElapsedTime(&ms, h->t0, h->tm));
    h->stat_last_scan_ns = static_cast<int64_t>(static_cast<double>(ms) * 1e6);
  }
  return 0;
}

// ---- retained rows ---------------------------------------------------------------------------
// a row arrived by a route that does not keep its source: the host copies no longer describe the whole index
static void drop_codes(sgic_index* h) {
  h->codes_ok = false;
  h->codes.clear();
  h->codes.shrink_to_fit();
}
static void keep_codes(sgic_index* h, const uint8_t* q, int64_t n) {
  if (!h->codes_ok) return;
  try {
    h->codes.insert(h->codes.end(), q, q + static_cast<size_t>(n) * h->d);
  } catch (const std::bad_alloc&) {
    drop_codes(h);
  }
}

// numpy's add.reduce over a contiguous float32 axis (pairwise_sum in loops_utils.h: blocks of <= 128 elements
// summed with 8 interleaved accumulators, halves split on multiples of 8), which is what
// np.linalg.norm(x, axis=-1) of l2n (src/search.py:16-18 == src/build.py:18-20) runs on x*x.
// tests/test_oracle.py checks this restatement against numpy itself, bit for bit.
static float np_pairwise_sum_f32(const float* a, int64_t n) {
  if (n < 8) {
    float res = 0.f;
    for (int64_t i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
}

// u8 codes -> the fp32 unit rows of the reference, same operations in the same order:
//   z = (q.astype(float32) / 255.0) * 2.0 - 1.0 ;  z / np.maximum(np.linalg.norm(z, axis=-1, keepdims=True), 1e-9)
// (this translation unit is compiled with -ffp-contract=off: no fused multiply-add may merge two roundings)
void rows_from_codes_f32(const uint8_t* q, int64_t n, int d, float* out) {
  float lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = (static_cast<float>(v) / 255.0f) * 2.0f - 1.0f;
  std::vector<float> sq(static_cast<size_t>(d));
  for (int64_t r = 0; r < n; ++r) {
    const uint8_t* qr = q + static_cast<size_t>(r) * d;
    float* zr = out + static_cast<size_t>(r) * d;
    for (int i = 0; i < d; ++i) {
      zr[i] = lut[qr[i]];
      sq[static_cast<size_t>(i)] = zr[i] * zr[i];
    }
    const float nrm = std::sqrt(np_pairwise_sum_f32(sq.data(), d));
    const float den = nrm > 1e-9f ? nrm : 1e-9f;
    for (int i = 0; i < d; ++i) zr[i] = zr[i] / den;
  }
}

// ---- IxFI ---------------------------------------------------------------------------------
#pragma pack(push, 1)
struct IxfiHeader {
  char fourcc[4];
  int32_t d;
  int64_t ntotal;
  int64_t dummy0;
  int64_t dummy1;
  uint8_t is_trained;
  int32_t metric_type;
  uint64_t count;
};
#pragma pack(pop)
static_assert(sizeof(IxfiHeader) == 45, "IxFI header is 45 bytes (SURVEY.md §8a F3)");

}  // namespace sgic

using namespace sgic;

namespace sgic {
static bool is_front(const sgic_index* h) { return h != nullptr && !h->shards.empty(); }
static void front_destroy(sgic_index* f);
static int front_search_host(sgic_index* f, int64_t nq, const float* host_q, int64_t k, float* host_D, int64_t* host_I);
static int front_search_dev(sgic_index* f, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                            int64_t id_base, cudaStream_t stream);
static int front_add_host(sgic_index* f, int64_t n, const void* host, int kind);
static int front_add_dev(sgic_index* f, int64_t n, const void* dev, int kind, cudaStream_t stream);
static int front_add_c2df(sgic_index* f, const uint8_t* blob, const int64_t* offsets, int64_t n, int32_t* status_out,
                          int64_t* n_added, int n_threads);
static int front_reconstruct(sgic_index* f, int64_t i0, int64_t n, float* host_out);
static int front_codes(sgic_index* f, int64_t i0, int64_t n, uint8_t* host_out);
static int front_write_mode(const sgic_index* f);
static int front_rows_for_write(sgic_index* f, int mode, int64_t i0, int64_t n, float* out);
static int front_set_option(sgic_index* f, const std::string& name, int64_t value);
static int64_t front_get_stat(const sgic_index* f, const std::string& name, bool* handled);
static int front_reset(sgic_index* f);
}  // namespace sgic

extern "C" {

const char* sgic_last_error(void) { return g_err.c_str(); }
int sgic_version(void) { return 200; }

int sgic_index_create(int d, int dtype, int device, int64_t capacity_rows, int flags, sgic_index** out) {
  SGIC_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  SGIC_REQUIRE(d > 0 && d % 8 == 0 && d <= kMaxD, "d must be a positive multiple of 8 and <= 2048");
  SGIC_REQUIRE(dtype == SGIC_F16 || dtype == SGIC_BF16, "dtype must be SGIC_F16 or SGIC_BF16");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
              "); this library has no CPU fallback");
    return 2;
  }
  SGIC_REQUIRE(device >= 0 && device < n_dev, "device ordinal out of range");
  DeviceGuard g(device);
  cudaDeviceProp prop;
  SGIC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
              "; this library is built for sm_100a (B200) only");
    return 2;
  }
  sgic_index* h = new sgic_index();
  h->d = d;
  h->dtype = dtype;
  h->device = device;
  h->flags = flags;
  h->sm_count = prop.multiProcessorCount;
  h->retain_ok = (flags & SGIC_RETAIN_F32) != 0;
  h->codes_ok = (flags & SGIC_RETAIN_U8) != 0;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&h->t0) != cudaSuccess || cudaEventCreate(&h->t1) != cudaSuccess ||
      cudaEventCreate(&h->tm) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming) != cudaSuccess) {
    set_error("stream/event creation failed");
    delete h;
    return 2;
  }
  if (capacity_rows > 0) {
    int rc = ensure_capacity(h, capacity_rows, h->stream);
    if (rc) {
      sgic_index_destroy(h);
      return rc;
    }
  }
  *out = h;
  return 0;
}

int sgic_index_destroy(sgic_index* h) {
  if (!h) return 0;
  if (is_front(h) || h->front) {
    front_destroy(h);
    delete h;
    return 0;
  }
  DeviceGuard g(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int i = 0; i < 2; ++i) {
    if (h->pin[i]) cudaFreeHost(h->pin[i]);
    if (h->dstage[i]) cudaFree(h->dstage[i]);
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  }
  if (h->ws) cudaFree(h->ws);
  if (h->qdev) cudaFree(h->qdev);
  if (h->qh) cudaFree(h->qh);
  if (h->lists_ws) cudaFree(h->lists_ws);
  if (h->seed_ws) cudaFree(h->seed_ws);
  if (h->ws2) cudaFree(h->ws2);
  for (auto& z : h->zl) {
    if (z.rows) cudaFree(z.rows);
    if (z.frames) cudaFree(z.frames);
    if (z.desc) cudaFree(z.desc);
    if (z.status) cudaFree(z.status);
    if (z.status_host) cudaFreeHost(z.status_host);
    if (z.pin_rows) cudaFreeHost(z.pin_rows);
    if (z.pin_frames) cudaFreeHost(z.pin_frames);
    if (z.pin_desc) cudaFreeHost(z.pin_desc);
    for (cudaEvent_t e : {z.copied, z.decoded, z.freed, z.t_c0, z.t_c1, z.t_k0, z.t_k1, z.t_q0, z.t_q1})
      if (e) cudaEventDestroy(e);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->odev) cudaFree(h->odev);
  if (h->opin) cudaFreeHost(h->opin);
  if (h->qpin) cudaFreeHost(h->qpin);
  if (h->vmm_base) vmm_destroy(h);
  else if (h->db) cudaFree(h->db);
  if (h->t0) cudaEventDestroy(h->t0);
  if (h->t1) cudaEventDestroy(h->t1);
  if (h->tm) cudaEventDestroy(h->tm);
  if (h->ev_order) cudaEventDestroy(h->ev_order);
  for (cudaEvent_t e : h->ring_a) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ring_b) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int64_t sgic_index_ntotal(const sgic_index* h) { return h ? h->ntotal : -1; }
int sgic_index_d(const sgic_index* h) { return h ? h->d : -1; }
int sgic_index_dtype(const sgic_index* h) { return h ? h->dtype : -1; }
int sgic_index_device(const sgic_index* h) { return h ? h->device : -1; }
const void* sgic_index_data_dev(const sgic_index* h) { return h ? h->db : nullptr; }

int sgic_index_reserve(sgic_index* h, int64_t rows) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) {  // equal shares: what a bulk load or large appends fill
    const int64_t G = static_cast<int64_t>(h->shards.size());
    for (sgic_index* s : h->shards) {
      int rc = sgic_index_reserve(s, (rows + G - 1) / G);
      if (rc) return rc;
    }
    return 0;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  if (rows <= h->capacity) return 0;
  int orc = order_begin(h, h->stream);
  if (orc) return orc;
  return grow_database(h, rows, h->stream, /*exact=*/true);
}

int sgic_index_reset(sgic_index* h) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_reset(h);
  std::lock_guard<std::mutex> lk(h->mu);
  h->ntotal = 0;
  h->retained.clear();
  h->retain_ok = (h->flags & SGIC_RETAIN_F32) != 0;
  h->codes.clear();
  h->codes_ok = (h->flags & SGIC_RETAIN_U8) != 0;
  return 0;
}

int sgic_index_add_f32(sgic_index* h, int64_t n, const float* host_x) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_add_host(h, n, host_x, 0);
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  SGIC_REQUIRE(host_x != nullptr, "x is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  int rc = order_begin(h, h->stream);
  if (rc) return rc;
  rc = ensure_capacity(h, h->ntotal + n, h->stream);
  if (rc) return rc;
  const int64_t base = h->ntotal;
  rc = stream_rows_h2d(h, reinterpret_cast<const uint8_t*>(host_x), n, static_cast<size_t>(h->d) * 4,
                       [&](void* dev, int64_t first, int64_t rows) {
                         return launch_pack_f32(h, static_cast<const float*>(dev), base + first, rows, h->stream);
                       });
  if (rc) return rc;
  if ((rc = order_end(h, h->stream))) return rc;
  drop_codes(h);
  if (h->retain_ok) {
    try {
      h->retained.insert(h->retained.end(), host_x, host_x + static_cast<size_t>(n) * h->d);
    } catch (const std::bad_alloc&) {
      h->retained.clear();
      h->retained.shrink_to_fit();
      h->retain_ok = false;
    }
  }
  h->ntotal += n;
  return 0;
}

int sgic_index_add_f32_dev(sgic_index* h, int64_t n, const float* dev_x, void* stream) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_add_dev(h, n, dev_x, 0, static_cast<cudaStream_t>(stream));
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  int rc = order_begin(h, st);
  if (rc) return rc;
  rc = ensure_capacity(h, h->ntotal + n, st);
  if (rc) return rc;
  rc = launch_pack_f32(h, dev_x, h->ntotal, n, st);
  if (rc) return rc;
  if ((rc = order_end(h, st))) return rc;
  h->retain_ok = false;
  drop_codes(h);
  h->retained.clear();
  h->ntotal += n;
  return 0;
}

int sgic_index_add_packed_dev(sgic_index* h, int64_t n, const void* dev_rows, void* stream) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_add_dev(h, n, dev_rows, 2, static_cast<cudaStream_t>(stream));
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  int rc = order_begin(h, st);
  if (rc) return rc;
  rc = ensure_capacity(h, h->ntotal + n, st);
  if (rc) return rc;
  SGIC_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(h->db) + elt_rows_bytes(h, h->ntotal), dev_rows,
                            elt_rows_bytes(h, n), cudaMemcpyDeviceToDevice, st));
  if ((rc = order_end(h, st))) return rc;
  h->retain_ok = false;
  drop_codes(h);
  h->retained.clear();
  h->ntotal += n;
  return 0;
}

int sgic_index_add_u8(sgic_index* h, int64_t n, const uint8_t* host_q) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_add_host(h, n, host_q, 1);
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  SGIC_REQUIRE(host_q != nullptr, "q is NULL");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  int rc = order_begin(h, h->stream);
  if (rc) return rc;
  rc = ensure_capacity(h, h->ntotal + n, h->stream);
  if (rc) return rc;
  const int64_t base = h->ntotal;
  rc = stream_rows_h2d(h, host_q, n, static_cast<size_t>(h->d), [&](void* dev, int64_t first, int64_t rows) {
    return launch_dequant_u8(h, static_cast<const uint8_t*>(dev), base + first, rows, h->stream);
  });
  if (rc) return rc;
  if ((rc = order_end(h, h->stream))) return rc;
  h->retain_ok = false;
  keep_codes(h, host_q, n);
  h->retained.clear();
  h->ntotal += n;
  return 0;
}

int sgic_index_add_u8_dev(sgic_index* h, int64_t n, const uint8_t* dev_q, void* stream) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_add_dev(h, n, dev_q, 1, static_cast<cudaStream_t>(stream));
  SGIC_REQUIRE(n >= 0, "n must be >= 0");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  int rc = order_begin(h, st);
  if (rc) return rc;
  rc = ensure_capacity(h, h->ntotal + n, st);
  if (rc) return rc;
  rc = launch_dequant_u8(h, dev_q, h->ntotal, n, st);
  if (rc) return rc;
  if ((rc = order_end(h, st))) return rc;
  h->retain_ok = false;
  drop_codes(h);
  h->retained.clear();
  h->ntotal += n;
  return 0;
}

int sgic_c2df_parse(const uint8_t* blob, const int64_t* offsets, int64_t n, int dim, uint8_t* out_u8,
                    int32_t* status_out, int32_t* dim_out, int n_threads) {
  SGIC_REQUIRE(blob != nullptr || n == 0, "blob is NULL");
  SGIC_REQUIRE(offsets != nullptr && status_out != nullptr && (out_u8 != nullptr || n == 0), "NULL argument");
  return c2df_parse_batch(blob, offsets, n, dim, out_u8, status_out, dim_out, n_threads);
}

// ---- ingest pipeline stages (device-side clip_stream decode) ---------------------------------
// A slab goes through: host (TLV walk + classification + pinned packing, n_threads) -> copy (H2D on the copy
// stream) -> decode (K0 on the index's stream, status back) -> finalize (host checks the statuses, K1 expands the
// u8 rows into the database, or libzstd redoes the slab).  Slabs are finalized in order, so rows keep file order.
static int zl_slab_prepare(sgic_index* h, sgic_index::ZlSlab& z) {
  using namespace sgic;
  if (!h->copy_stream) SGIC_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&z.copied, &z.decoded, &z.freed})
    if (!*e) SGIC_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  for (cudaEvent_t* e : {&z.t_c0, &z.t_c1, &z.t_k0, &z.t_k1, &z.t_q0, &z.t_q1})
    if (!*e) SGIC_CUDA(cudaEventCreate(e));
  return 0;
}

// "timing" option: per-phase device times of the slab that last used this set (events long complete by now)
static int zl_slab_harvest(sgic_index* h, sgic_index::ZlSlab& z) {
  if (!z.timed) return 0;
  z.timed = false;
  float ms = 0.f;
  SGIC_CUDA(cudaEventSynchronize(z.t_q1));
  SGIC_CUDA(cudaEventElapsedTime(&ms, z.t_c0, z.t_c1));
  h->stat_ingest_h2d_ns += static_cast<int64_t>(ms * 1e6);
  SGIC_CUDA(cudaEventElapsedTime(&ms, z.t_k0, z.t_k1));
  h->stat_ingest_k0_ns += static_cast<int64_t>(ms * 1e6);
  SGIC_CUDA(cudaEventElapsedTime(&ms, z.t_q0, z.t_q1));
  h->stat_ingest_k1_ns += static_cast<int64_t>(ms * 1e6);
  return 0;
}

// H2D of a packed slab on the copy stream; the set's device buffers are free once K1 of their previous slab ran
static int zl_slab_copy(sgic_index* h, sgic_index::ZlSlab& z) {
  using namespace sgic;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  int rc;
  const size_t d = static_cast<size_t>(h->d);
  const uint32_t nf = static_cast<uint32_t>(z.nf);
  cudaStream_t cs = h->copy_stream;
  if (z.used) SGIC_CUDA(cudaStreamWaitEvent(cs, z.freed, 0));
  // device buffers sized for a FULL slab of typical frames the first time (a short first slab must not cost a
  // reallocation — cudaFree / cudaFreeHost synchronise — when the next one is full)
  const size_t n_full = static_cast<size_t>(std::max<int64_t>(z.full, z.w));
  const size_t need_rows = n_full * d, need_frames = std::max(z.fbytes + 16, n_full * 640 + (1u << 20));
  const size_t need_desc = std::max<size_t>(nf, n_full) * sizeof(ZlDesc), need_status = std::max<size_t>(nf, n_full) * 4;
  const bool grow = need_rows > z.rows_bytes || need_frames > z.frames_bytes || need_desc > z.desc_bytes ||
                    need_status > z.status_bytes;
  if (grow) SGIC_CUDA(cudaDeviceSynchronize());  // (re)allocation: nothing may still use the old buffers
  if ((rc = ensure_buf(&z.rows, &z.rows_bytes, need_rows, false))) return rc;
  if ((rc = ensure_buf(&z.frames, &z.frames_bytes, need_frames, false))) return rc;
  if ((rc = ensure_buf(&z.desc, &z.desc_bytes, need_desc, false))) return rc;
  if ((rc = ensure_buf(&z.status, &z.status_bytes, need_status, false))) return rc;
  if ((rc = ensure_buf(&z.status_host, &z.status_host_bytes, need_status, true))) return rc;
  if (h->opt_timing) SGIC_CUDA(cudaEventRecord(z.t_c0, cs));
  if (z.nh > 0) SGIC_CUDA(cudaMemcpyAsync(z.rows, z.pin_rows, static_cast<size_t>(z.w) * d, cudaMemcpyHostToDevice, cs));
  SGIC_CUDA(cudaMemcpyAsync(z.frames, z.pin_frames, z.fbytes, cudaMemcpyHostToDevice, cs));
  SGIC_CUDA(cudaMemcpyAsync(z.desc, z.pin_desc, static_cast<size_t>(nf) * sizeof(ZlDesc), cudaMemcpyHostToDevice, cs));
  if (h->opt_timing) SGIC_CUDA(cudaEventRecord(z.t_c1, cs));
  SGIC_CUDA(cudaEventRecord(z.copied, cs));
  z.used = true;
  return 0;
}

// K0 on the index's stream once the slab has landed; statuses come back to pinned memory
static int zl_slab_decode(sgic_index* h, sgic_index::ZlSlab& z) {
  using namespace sgic;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const uint32_t nf = static_cast<uint32_t>(z.nf);
  {
    int orc = order_begin(h, st);
    if (orc) return orc;
    if ((orc = order_end(h, st))) return orc;
  }
  SGIC_CUDA(cudaStreamWaitEvent(st, z.copied, 0));
  const size_t smem = static_cast<size_t>(kZlWarpsPerBlock) * sizeof(ZlWarpSmem);  // ~71 KB: 3 CTAs per SM
  static bool zl_configured[64] = {false};
  if (!zl_configured[h->device & 63]) {
    SGIC_CUDA(cudaFuncSetAttribute(zstd_lit_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
    zl_configured[h->device & 63] = true;
  }
  const unsigned grid = std::min<unsigned>((nf + kZlWarpsPerBlock - 1) / kZlWarpsPerBlock,
                                           static_cast<unsigned>(h->sm_count) * 3u);
  if (h->opt_timing) SGIC_CUDA(cudaEventRecord(z.t_k0, st));
  zstd_lit_decode_kernel<<<grid, kZlWarpsPerBlock * 32, smem, st>>>(
      static_cast<const uint8_t*>(z.frames), static_cast<const ZlDesc*>(z.desc), nf, static_cast<uint32_t>(h->d),
      static_cast<uint8_t*>(z.rows), static_cast<int32_t*>(z.status));
  h->stat_launches++;
  SGIC_CUDA(cudaGetLastError());
  if (h->opt_timing) SGIC_CUDA(cudaEventRecord(z.t_k1, st));
  // SGIC_RETAIN_U8: the slab's u8 rows come back into the pinned buffer they partly left from (the H2D of the
  // host-decoded rows is long done: K0 waited for it), 1 byte per element; finalize appends them to h->codes
  if (h->codes_ok)
    SGIC_CUDA(cudaMemcpyAsync(z.pin_rows, z.rows, static_cast<size_t>(z.w) * h->d, cudaMemcpyDeviceToHost, st));
  SGIC_CUDA(cudaMemcpyAsync(z.status_host, z.status, static_cast<size_t>(nf) * 4, cudaMemcpyDeviceToHost, st));
  SGIC_CUDA(cudaEventRecord(z.decoded, st));
  return 0;
}

// Waits for the slab's statuses; all good: K1 appends its rows (asynchronously, stream-ordered before whatever
// touches the database next).  *bad = true (nothing appended) if the device rejected a frame: the caller redoes
// the slab with libzstd, which stays the judge of malformed input.
static int zl_slab_finalize(sgic_index* h, sgic_index::ZlSlab& z, bool* bad) {
  using namespace sgic;
  *bad = false;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  SGIC_CUDA(cudaEventSynchronize(z.decoded));
  const int32_t* stt = static_cast<const int32_t*>(z.status_host);
  for (int64_t i = 0; i < z.nf; ++i)
    if (stt[i] != zl::ZL_OK) {
      *bad = true;
      SGIC_CUDA(cudaEventRecord(z.freed, st));
      return 0;
    }
  int rc = order_begin(h, st);
  if (rc) return rc;
  if ((rc = order_end(h, st))) return rc;
  rc = ensure_capacity(h, h->ntotal + z.w, st);
  if (rc) return rc;
  if (h->opt_timing) SGIC_CUDA(cudaEventRecord(z.t_q0, st));
  rc = launch_dequant_u8(h, static_cast<const uint8_t*>(z.rows), h->ntotal, z.w, st);
  if (rc) return rc;
  if (h->opt_timing) {
    SGIC_CUDA(cudaEventRecord(z.t_q1, st));
    z.timed = true;
  }
  SGIC_CUDA(cudaEventRecord(z.freed, st));
  h->retain_ok = false;
  keep_codes(h, static_cast<const uint8_t*>(z.pin_rows), z.w);  // copied back behind K0 (zl_slab_decode)
  h->retained.clear();
  h->ntotal += z.w;
  h->stat_zl_device_frames += z.nf;
  h->stat_zl_host_rows += z.nh;
  return 0;
}

int sgic_index_add_c2df(sgic_index* h, const uint8_t* blob, const int64_t* offsets, int64_t n, int32_t* status_out,
                        int64_t* n_added, int n_threads) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_add_c2df(h, blob, offsets, n, status_out, n_added, n_threads);
  SGIC_REQUIRE(n >= 0 && offsets != nullptr && status_out != nullptr, "bad arguments");
  if (n_added) *n_added = 0;
  if (n == 0) return 0;
  const size_t d = static_cast<size_t>(h->d);
  const bool dev_dec = h->opt_device_zstd != 0;
  // Slabs keep host memory bounded.  Host-only decode: 256 MB of rows per slab (zstd overlaps the H2D + K1 of the
  // previous chunk inside add_u8).  Device decode: the host walks the TLV container, lets libzstd decode only the
  // frames outside the device profile, and packs ~300-byte frames into pinned memory (all on n_threads threads);
  // one H2D + K0 + K1 per slab.
  const int64_t slab = dev_dec ? 65536 : std::max<int64_t>(1, static_cast<int64_t>((256u << 20) / d));
  struct RawBuf {  // uninitialised host rows (a std::vector would memset 64 MB per slab)
    uint8_t* p = nullptr;
    size_t cap = 0;
    ~RawBuf() { std::free(p); }
    uint8_t* data() { return p; }
    bool resize(size_t n) {
      if (n <= cap) return true;
      std::free(p);
      p = static_cast<uint8_t*>(std::malloc(n));
      cap = p ? n : 0;
      return p != nullptr;
    }
  } rows;
  std::vector<int64_t> foff;
  std::vector<uint32_t> flen;
  int64_t added = 0;
  auto host_route = [&](int64_t s0, int64_t cnt, bool reparse) -> int {
    int rc = 0;
    if (reparse) {
      rc = sgic::c2df_parse_batch(blob, offsets + s0, cnt, h->d, rows.data(), status_out + s0, nullptr, n_threads);
      if (rc) return rc;
    }
    // compact the good rows in place (order preserved, as build.py's `keep` list does)
    int64_t w = 0;
    for (int64_t i = 0; i < cnt; ++i) {
      if (status_out[s0 + i] != SGIC_C2DF_OK) continue;
      if (w != i) std::memmove(rows.data() + static_cast<size_t>(w) * d, rows.data() + static_cast<size_t>(i) * d, d);
      ++w;
    }
    if (w > 0) {
      rc = sgic_index_add_u8(h, w, rows.data());
      if (rc) return rc;
      added += w;
    }
    return 0;
  };
  auto now_ns = [] {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch())
        .count();
  };
  if (!dev_dec) {
    for (int64_t s0 = 0; s0 < n; s0 += slab) {
      const int64_t cnt = std::min(slab, n - s0);
      SGIC_REQUIRE(rows.resize(static_cast<size_t>(cnt) * d), "out of host memory for the row slab");
      int rc = host_route(s0, cnt, true);
      if (rc) return rc;
    }
    if (n_added) *n_added = added;
    return 0;
  }
  // Device decode, pipelined: step i = host work of slab i, then finalize slab i-2 (its K0 ran under the host work of
  // slabs i-1 and i), then enqueue copy + K0 of slab i (the copy overlaps K0 of slab i-1).
  constexpr int kDepth = sgic_index::kZlDepth;
  const int64_t n_slabs = (n + slab - 1) / slab;
  auto finalize = [&](sgic_index::ZlSlab& z) -> int {
    int rc = 0;
    if (z.mode == 1) {
      const int64_t t0 = now_ns();
      bool bad = false;
      rc = zl_slab_finalize(h, z, &bad);
      h->stat_ingest_gpu_ns += now_ns() - t0;   // host time spent waiting for the device
      if (rc) return rc;
      if (bad) {  // a frame the device could not decode: libzstd decides for the whole slab
        h->stat_zl_fallback_slabs++;
        SGIC_REQUIRE(rows.resize(static_cast<size_t>(z.cnt) * d), "out of host memory for the row slab");
        rc = host_route(z.s0, z.cnt, true);
      } else {
        added += z.w;
      }
    } else if (z.mode == 2) {  // nothing for the device in this slab
      SGIC_REQUIRE(rows.resize(static_cast<size_t>(z.cnt) * d), "out of host memory for the row slab");
      rc = host_route(z.s0, z.cnt, true);
    }
    z.mode = 0;
    return rc;
  };
  for (int64_t i = 0; i < n_slabs + (kDepth - 1); ++i) {
    if (i < n_slabs) {
      sgic_index::ZlSlab& z = h->zl[i % kDepth];
      const int64_t s0 = i * slab, cnt = std::min(slab, n - s0);
      z.s0 = s0;
      z.cnt = cnt;
      z.full = slab;
      z.mode = 0;
      SGIC_REQUIRE(rows.resize(static_cast<size_t>(cnt) * d), "out of host memory for the row slab");
      foff.resize(static_cast<size_t>(cnt));
      flen.resize(static_cast<size_t>(cnt));
      const int64_t t_a = now_ns();
      int rc = sgic::c2df_parse_batch(blob, offsets + s0, cnt, h->d, rows.data(), status_out + s0, nullptr, n_threads,
                                      foff.data(), flen.data());
      if (rc) return rc;
      const int64_t t_b = now_ns();
      h->stat_ingest_parse_ns += t_b - t_a;
      {
        std::lock_guard<std::mutex> lk(h->mu);
        sgic::DeviceGuard g(h->device);
        if ((rc = zl_slab_prepare(h, z))) return rc;
        if ((rc = zl_slab_harvest(h, z))) return rc;
        // pinned slab buffers: sized once for a full slab of typical frames (grown only if a slab needs more);
        // the previous slab of this set was finalized two steps ago, so nothing reads them any more
        const size_t frames_need = static_cast<size_t>(slab) * 640 + (1u << 20);
        if ((rc = sgic::ensure_buf(&z.pin_rows, &z.pin_rows_bytes, static_cast<size_t>(slab) * d, true))) return rc;
        if ((rc = sgic::ensure_buf(&z.pin_desc, &z.pin_desc_bytes, static_cast<size_t>(slab) * sizeof(sgic::ZlDesc), true)))
          return rc;
        size_t need = 0;
        for (int64_t j = 0; j < cnt; ++j)
          if (status_out[s0 + j] == SGIC_C2DF_OK && foff[static_cast<size_t>(j)] >= 0)
            need += (static_cast<size_t>(flen[static_cast<size_t>(j)]) + 15u) & ~static_cast<size_t>(15);
        if ((rc = sgic::ensure_buf(&z.pin_frames, &z.pin_frames_bytes, std::max(need + 16, frames_need), true))) return rc;
      }
      rc = sgic::c2df_pack_batch(blob, cnt, h->d, status_out + s0, foff.data(), flen.data(), rows.data(),
                                 static_cast<uint8_t*>(z.pin_rows), static_cast<uint8_t*>(z.pin_frames),
                                 z.pin_frames_bytes, static_cast<sgic::ZlDesc*>(z.pin_desc), &z.w, &z.nf, &z.nh, &z.fbytes,
                                 n_threads);
      if (rc) return rc;
      h->stat_ingest_pack_ns += now_ns() - t_b;
      z.mode = z.w == 0 ? 0 : (z.nf == 0 ? 2 : 1);
    }
    if (i >= kDepth - 1) {
      int rc = finalize(h->zl[(i - (kDepth - 1)) % kDepth]);
      if (rc) return rc;
    }
    if (i < n_slabs && h->zl[i % kDepth].mode == 1) {
      int rc = zl_slab_copy(h, h->zl[i % kDepth]);
      if (rc) return rc;
      if ((rc = zl_slab_decode(h, h->zl[i % kDepth]))) return rc;
    }
  }
  {
    std::lock_guard<std::mutex> lk(h->mu);
    sgic::DeviceGuard g(h->device);
    SGIC_CUDA(cudaStreamSynchronize(h->stream));  // the last K1s: the rows are in the database when the call returns
    for (auto& z : h->zl) {
      int rc = zl_slab_harvest(h, z);
      if (rc) return rc;
    }
  }
  if (n_added) *n_added = added;
  return 0;
}

int sgic_index_search_dev(sgic_index* h, int64_t nq, const float* dev_q, int64_t k, float* dev_D, int64_t* dev_I,
                          int64_t id_base, void* stream) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_search_dev(h, nq, dev_q, k, dev_D, dev_I, id_base, static_cast<cudaStream_t>(stream));
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
  int rc = order_begin(h, st);
  if (rc) return rc;
  rc = search_dev_impl(h, nq, dev_q, k, dev_D, dev_I, id_base, st);
  if (rc) return rc;
  return order_end(h, st);
}

int sgic_index_search(sgic_index* h, int64_t nq, const float* host_q, int64_t k, float* host_D, int64_t* host_I) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_search_host(h, nq, host_q, k, host_D, host_I);
  SGIC_REQUIRE(k >= 1, "k must be >= 1");
  SGIC_REQUIRE(nq >= 0, "nq must be >= 0");
  if (nq == 0) return 0;
  SGIC_REQUIRE(host_q && host_D && host_I, "NULL buffer");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const size_t qbytes = static_cast<size_t>(nq) * h->d * 4;
  const size_t dbytes = static_cast<size_t>(nq) * k * 4, ibytes = static_cast<size_t>(nq) * k * 8;
  int rc = ensure_buf(&h->qdev, &h->qdev_bytes, qbytes, false);
  if (rc) return rc;
  rc = ensure_buf(&h->odev, &h->odev_bytes, dbytes + ibytes, false);
  if (rc) return rc;
  rc = ensure_buf(&h->opin, &h->opin_bytes, std::max(qbytes, dbytes + ibytes), true);
  if (rc) return rc;
  if ((rc = order_begin(h, h->stream))) return rc;
  if ((rc = order_end(h, h->stream))) return rc;
  // The answer of a small search is stored by the final writer (K3's last CTA / the merge kernel) straight into
  // pinned host memory — no D2H copy engine transfer (~8 us of start-up latency) behind the kernel; large answers
  // come back through cudaMemcpyAsync.  The queries go in through search_hostq_impl (kernel parameters or H2D).
  if (dbytes + ibytes <= kZeroCopyBytes) {
    int64_t* pI = reinterpret_cast<int64_t*>(h->opin);
    float* pD = reinterpret_cast<float*>(static_cast<uint8_t*>(h->opin) + ibytes);
    rc = search_hostq_impl(h, nq, host_q, false, k, pD, pI, 0, h->stream);
    if (rc) return rc;
    SGIC_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(host_I, pI, ibytes);
    std::memcpy(host_D, pD, dbytes);
    return 0;
  }
  int64_t* dI = reinterpret_cast<int64_t*>(h->odev);
  float* dD = reinterpret_cast<float*>(static_cast<uint8_t*>(h->odev) + ibytes);
  rc = search_hostq_impl(h, nq, host_q, false, k, dD, dI, 0, h->stream);
  if (rc) return rc;
  SGIC_CUDA(cudaMemcpyAsync(h->opin, h->odev, dbytes + ibytes, cudaMemcpyDeviceToHost, h->stream));
  SGIC_CUDA(cudaStreamSynchronize(h->stream));
  std::memcpy(host_I, h->opin, ibytes);
  std::memcpy(host_D, static_cast<uint8_t*>(h->opin) + ibytes, dbytes);
  return 0;
}

int sgic_merge_topk_dev(int device, int64_t nq, int n_lists, int64_t k, const float* dev_D_lists,
                        const int64_t* dev_I_lists, float* dev_D, int64_t* dev_I, int tie_by_position,
                        void* stream) {
  SGIC_REQUIRE(nq >= 0 && n_lists >= 1 && k >= 1, "bad arguments");
  SGIC_REQUIRE(static_cast<int64_t>(n_lists) * k < (1ll << 31), "too many candidates");
  if (nq == 0) return 0;
  DeviceGuard g(device);
  MergeListsParams mp;
  mp.D_lists = dev_D_lists;
  mp.I_lists = reinterpret_cast<const long long*>(dev_I_lists);
  mp.n_lists = static_cast<uint32_t>(n_lists);
  mp.nq = static_cast<uint32_t>(nq);
  mp.k = static_cast<uint32_t>(k);
  const uint32_t kp = next_pow2_u32(mp.k);
  uint32_t chunk = std::max<uint32_t>(2 * kp, std::min<uint32_t>(16384, next_pow2_u32(mp.n_lists * mp.k)));
  chunk = std::max<uint32_t>(chunk, 64);
  SGIC_REQUIRE(static_cast<size_t>(chunk) * 8 <= kSmemBudget, "k too large for the merge kernel");
  mp.chunk = chunk;
  mp.tie_by_position = tie_by_position ? 1u : 0u;
  mp.D = dev_D;
  mp.I = reinterpret_cast<long long*>(dev_I);
  const size_t smem = static_cast<size_t>(chunk) * 8;
  if (smem > 48 * 1024)
    SGIC_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  const int threads = static_cast<int>(std::min<uint32_t>(1024, std::max<uint32_t>(32, chunk / 2)));
  merge_lists_kernel<<<static_cast<unsigned>(nq), threads, smem, static_cast<cudaStream_t>(stream)>>>(mp);
  SGIC_CUDA(cudaGetLastError());
  return 0;
}

// ---- K5x: exchange of the per-shard answers by peer stores over NVLink (SURVEY.md §8e "optional fusion") --------
// Each rank owns one buffer (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by every other rank):
//   flags [2 parities][world] one 128-byte line each | error word | D_lists [2][world][max_cands] f32 |
//   I_lists [2][world][max_cands] i64
// A search step pushes this rank's (nq,k) answer into slot `rank` of every rank's buffer (xchg_push_kernel: plain
// stores into the peers' HBM + one system-scope release flag per destination) and merges the `world` lists of its
// own buffer as soon as their flags carry the step's epoch (merge_lists_kernel with wait_flags).  Two launches, no
// NCCL call, no packing.  Parity double-buffering is enough: a rank can only start step s+1's push after its merge
// of step s, which waited for every peer's push of step s, which follows that peer's merge of step s-1.
struct sgic_xchg {
  int device = 0, world = 0, rank = 0;
  int64_t max_cands = 0;
  uint8_t* base = nullptr;
  size_t bytes = 0, d_off = 0, i_off = 0, err_off = 0;
  uint8_t* peer[16] = {nullptr};
  uint32_t epoch = 0;
  bool opened = false;
};

int sgic_xchg_create(int device, int world, int rank, int64_t max_cands, sgic_xchg** out) {
  SGIC_REQUIRE(out != nullptr, "out is NULL");
  SGIC_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world, "world must be 2..16 and 0 <= rank < world");
  SGIC_REQUIRE(max_cands >= 1 && max_cands <= (1ll << 24), "max_cands out of range");
  DeviceGuard g(device);
  auto* x = new sgic_xchg();
  x->device = device;
  x->world = world;
  x->rank = rank;
  x->max_cands = max_cands;
  x->err_off = static_cast<size_t>(2) * world * 128;
  x->d_off = x->err_off + 128;
  x->i_off = x->d_off + static_cast<size_t>(2) * world * max_cands * 4;
  x->i_off = (x->i_off + 127) & ~size_t(127);
  x->bytes = x->i_off + static_cast<size_t>(2) * world * max_cands * 8;
  cudaError_t e = cudaMalloc(&x->base, x->bytes);
  if (e != cudaSuccess) {
    delete x;
    set_error(std::string("exchange buffer: cudaMalloc failed: ") + cudaGetErrorString(e));
    return 3;
  }
  e = cudaMemset(x->base, 0, x->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(x->base);
    delete x;
    set_error(std::string("exchange buffer: memset failed: ") + cudaGetErrorString(e));
    return 2;
  }
  *out = x;
  return 0;
}

int sgic_xchg_export(sgic_xchg* x, uint8_t* handle64) {
  SGIC_REQUIRE(x != nullptr && handle64 != nullptr, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard g(x->device);
  cudaIpcMemHandle_t h;
  SGIC_CUDA(cudaIpcGetMemHandle(&h, x->base));
  std::memcpy(handle64, &h, 64);
  return 0;
}

int sgic_xchg_open(sgic_xchg* x, const uint8_t* handles) {
  SGIC_REQUIRE(x != nullptr && handles != nullptr, "NULL argument");
  SGIC_REQUIRE(!x->opened, "exchange already opened");
  DeviceGuard g(x->device);
  for (int r = 0; r < x->world; ++r) {
    if (r == x->rank) {
      x->peer[r] = x->base;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + static_cast<size_t>(r) * 64, 64);
    void* ptr = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int r2 = 0; r2 < r; ++r2)
        if (r2 != x->rank && x->peer[r2]) cudaIpcCloseMemHandle(x->peer[r2]);
      for (auto& pp : x->peer) pp = nullptr;
      (void)cudaGetLastError();
      set_error(std::string("cudaIpcOpenMemHandle failed for rank ") + std::to_string(r) + ": " + cudaGetErrorString(e));
      return 2;
    }
    x->peer[r] = static_cast<uint8_t*>(ptr);
  }
  x->opened = true;
  return 0;
}

// one step of the exchange, in three parts so that a search can fold the push into its own final writer:
//   xchg_begin  next epoch; where this rank's (nq, k) answer goes in every rank's buffer, and the flags to raise
//   xchg_push   the push as a kernel of its own (the answer sits in local buffers)
//   xchg_merge  merge of this rank's own buffer as soon as the `world` flags carry the epoch
static void xchg_begin(sgic_xchg* x, XDst* xd) {
  const uint32_t epoch = ++x->epoch;
  const size_t par = epoch & 1u, W = static_cast<size_t>(x->world), MC = static_cast<size_t>(x->max_cands);
  xd->n = static_cast<uint32_t>(x->world);
  xd->epoch = epoch;
  for (int r = 0; r < x->world; ++r) {
    const size_t slot = par * W + static_cast<size_t>(x->rank);
    xd->D[r] = reinterpret_cast<float*>(x->peer[r] + x->d_off) + slot * MC;
    xd->I[r] = reinterpret_cast<long long*>(x->peer[r] + x->i_off) + slot * MC;
    xd->F[r] = reinterpret_cast<uint32_t*>(x->peer[r] + slot * 128);
  }
}

static int xchg_push(sgic_xchg* x, const XDst& xd, int64_t n, const float* dev_D_local, const int64_t* dev_I_local,
                     cudaStream_t st) {
  XchgPushParams pp;
  pp.D = dev_D_local;
  pp.I = reinterpret_cast<const long long*>(dev_I_local);
  pp.n = static_cast<uint32_t>(n);
  pp.epoch = xd.epoch;
  for (int r = 0; r < x->world; ++r) {
    pp.dstD[r] = xd.D[r];
    pp.dstI[r] = xd.I[r];
    pp.dstFlag[r] = xd.F[r];
  }
  xchg_push_kernel<<<static_cast<unsigned>(x->world), 256, 0, st>>>(pp);
  SGIC_CUDA(cudaGetLastError());
  return 0;
}

static int xchg_merge(sgic_xchg* x, uint32_t epoch, int64_t nq, int64_t k, float* dev_D, int64_t* dev_I,
                      int tie_by_position, cudaStream_t st) {
  const size_t par = epoch & 1u, W = static_cast<size_t>(x->world), MC = static_cast<size_t>(x->max_cands);
  MergeListsParams mp;
  mp.D_lists = reinterpret_cast<const float*>(x->base + x->d_off) + par * W * MC;
  mp.I_lists = reinterpret_cast<const long long*>(x->base + x->i_off) + par * W * MC;
  mp.n_lists = static_cast<uint32_t>(x->world);
  mp.nq = static_cast<uint32_t>(nq);
  mp.k = static_cast<uint32_t>(k);
  const uint32_t kp = next_pow2_u32(mp.k);
  uint32_t chunk = std::max<uint32_t>(2 * kp, std::min<uint32_t>(16384, next_pow2_u32(mp.n_lists * mp.k)));
  chunk = std::max<uint32_t>(chunk, 64);
  SGIC_REQUIRE(static_cast<size_t>(chunk) * 8 <= kSmemBudget, "k too large for the merge kernel");
  mp.chunk = chunk;
  mp.tie_by_position = tie_by_position ? 1u : 0u;
  mp.D = dev_D;
  mp.I = reinterpret_cast<long long*>(dev_I);
  mp.list_stride = MC;
  mp.wait_flags = reinterpret_cast<const uint32_t*>(x->base + par * W * 128);
  mp.wait_epoch = epoch;
  mp.err = reinterpret_cast<uint32_t*>(x->base + x->err_off);
  const size_t smem = static_cast<size_t>(chunk) * 8;
  if (smem > 48 * 1024)
    SGIC_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  const int threads = static_cast<int>(std::min<uint32_t>(1024, std::max<uint32_t>(32, chunk / 2)));
  merge_lists_kernel<<<static_cast<unsigned>(nq), threads, smem, st>>>(mp);
  SGIC_CUDA(cudaGetLastError());
  return 0;
}

int sgic_xchg_merge_dev(sgic_xchg* x, int64_t nq, int64_t k, const float* dev_D_local, const int64_t* dev_I_local,
                        float* dev_D, int64_t* dev_I, int tie_by_position, void* stream) {
  SGIC_REQUIRE(x != nullptr && x->opened, "exchange is not open");
  SGIC_REQUIRE(nq >= 1 && k >= 1 && nq * k <= x->max_cands, "nq * k exceeds the exchange buffer");
  SGIC_REQUIRE(static_cast<int64_t>(x->world) * k < (1ll << 31), "too many candidates");
  DeviceGuard g(x->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XDst xd;
  xchg_begin(x, &xd);
  int rc = xchg_push(x, xd, nq * k, dev_D_local, dev_I_local, st);
  if (rc) return rc;
  return xchg_merge(x, xd.epoch, nq, k, dev_D, dev_I, tie_by_position, st);
}

// One rank's whole search step with HOST buffers, in one call: pinned staging, H2D of the queries, the local scan
// with global row numbers, the peer exchange + merge, D2H of the merged answer, one synchronise.  (The Python
// host did the staging with torch calls before: ~70 us per step, more than the exchange itself.)
int sgic_xchg_search(sgic_xchg* x, sgic_index* h, int64_t nq, const float* host_q, int64_t k, float* host_D,
                     int64_t* host_I, int64_t id_base, int tie_by_position) {
  SGIC_REQUIRE(x != nullptr && h != nullptr, "NULL argument");
  SGIC_REQUIRE(x->opened, "exchange is not open");
  SGIC_REQUIRE(!is_front(h), "the peer exchange joins single-GPU indexes of different processes");
  SGIC_REQUIRE(k >= 1 && nq >= 1 && host_q && host_D && host_I, "bad arguments");
  SGIC_REQUIRE(nq * k <= x->max_cands, "nq * k exceeds the exchange buffer");
  SGIC_REQUIRE(static_cast<int64_t>(x->world) * k < (1ll << 31), "too many candidates");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const size_t qbytes = static_cast<size_t>(nq) * h->d * 4, cand = static_cast<size_t>(nq) * k;
  int rc = ensure_buf(&h->qdev, &h->qdev_bytes, qbytes, false);
  if (rc) return rc;
  const size_t half = (cand * 12 + 15) & ~size_t(15);  // merged answer | local answer, each ids (8 B) then scores
  if ((rc = ensure_buf(&h->odev, &h->odev_bytes, 2 * half, false))) return rc;
  if ((rc = ensure_buf(&h->opin, &h->opin_bytes, std::max(qbytes, cand * 12), true))) return rc;
  if ((rc = order_begin(h, h->stream))) return rc;
  if ((rc = order_end(h, h->stream))) return rc;
  uint8_t* o = static_cast<uint8_t*>(h->odev);
  int64_t* lI = reinterpret_cast<int64_t*>(o + half);          // this rank's local answer
  float* lD = reinterpret_cast<float*>(o + half + cand * 8);
  // the local scan: K3's last CTA stores the answer straight into every rank's buffer and raises the flags when it
  // can (one or two queries); otherwise the answer lands in (lD, lI) and a push kernel carries it
  XDst xd;
  xchg_begin(x, &xd);
  bool pushed = false;
  static const bool fold = [] { const char* e = std::getenv("SGIC_XCHG_FOLD"); return !(e && e[0] == '0'); }();
  rc = search_hostq_impl(h, nq, host_q, false, k, lD, lI, id_base, h->stream, fold ? &xd : nullptr, &pushed);
  if (rc) return rc;
  if (!pushed && (rc = xchg_push(x, xd, static_cast<int64_t>(cand), lD, lI, h->stream))) return rc;
  if (cand * 12 <= kZeroCopyBytes) {
    // small step: the merged answer is stored into pinned host memory by the merge kernel itself (no D2H copy)
    int64_t* pI = reinterpret_cast<int64_t*>(h->opin);
    float* pD = reinterpret_cast<float*>(static_cast<uint8_t*>(h->opin) + cand * 8);
    if ((rc = xchg_merge(x, xd.epoch, nq, k, pD, pI, tie_by_position, h->stream))) return rc;
    SGIC_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(host_I, pI, cand * 8);
    std::memcpy(host_D, pD, cand * 4);
    return 0;
  }
  int64_t* mI = reinterpret_cast<int64_t*>(o);                 // merged answer first: ids then scores, contiguous
  float* mD = reinterpret_cast<float*>(o + cand * 8);
  if ((rc = xchg_merge(x, xd.epoch, nq, k, mD, mI, tie_by_position, h->stream))) return rc;
  SGIC_CUDA(cudaMemcpyAsync(h->opin, h->odev, cand * 12, cudaMemcpyDeviceToHost, h->stream));
  SGIC_CUDA(cudaStreamSynchronize(h->stream));
  std::memcpy(host_I, h->opin, cand * 8);
  std::memcpy(host_D, static_cast<uint8_t*>(h->opin) + cand * 8, cand * 4);
  return 0;
}

int sgic_xchg_error(sgic_xchg* x) {
  SGIC_REQUIRE(x != nullptr, "exchange is NULL");
  DeviceGuard g(x->device);
  uint32_t v = 0;
  SGIC_CUDA(cudaMemcpy(&v, x->base + x->err_off, 4, cudaMemcpyDeviceToHost));
  if (v) set_error("peer exchange: a rank's candidates did not arrive within the time-out");
  return v ? 2 : 0;
}

int sgic_xchg_destroy(sgic_xchg* x) {
  if (x == nullptr) return 0;
  DeviceGuard g(x->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < x->world; ++r)
    if (r != x->rank && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
  if (x->base) cudaFree(x->base);
  delete x;
  return 0;
}

int sgic_index_reconstruct(sgic_index* h, int64_t i0, int64_t n, float* host_out) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (is_front(h)) return front_reconstruct(h, i0, n, host_out);
  SGIC_REQUIRE(i0 >= 0 && n >= 0 && i0 + n <= h->ntotal, "row range out of bounds");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  int rc = ensure_staging(h);
  if (rc) return rc;
  if ((rc = order_begin(h, h->stream))) return rc;
  if ((rc = order_end(h, h->stream))) return rc;
  const size_t row_bytes = static_cast<size_t>(h->d) * 4;
  const int64_t rows_per_chunk = std::max<int64_t>(1, static_cast<int64_t>(kStageChunkBytes / row_bytes));
  for (int64_t done = 0; done < n; done += rows_per_chunk) {
    const int64_t rows = std::min(rows_per_chunk, n - done);
    const size_t n8 = static_cast<size_t>(rows) * h->d / 8;
    const void* src = static_cast<const uint8_t*>(h->db) + elt_rows_bytes(h, i0 + done);
    const unsigned gsz = grid_for(n8, 256, h->sm_count);
    if (h->dtype == SGIC_F16)
      unpack_rows_kernel<__half><<<gsz, 256, 0, h->stream>>>(src, static_cast<float*>(h->dstage[0]), n8);
    else
      unpack_rows_kernel<__nv_bfloat16><<<gsz, 256, 0, h->stream>>>(src, static_cast<float*>(h->dstage[0]), n8);
    h->stat_launches++;
    SGIC_CUDA(cudaGetLastError());
    SGIC_CUDA(cudaMemcpyAsync(h->pin[0], h->dstage[0], static_cast<size_t>(rows) * row_bytes, cudaMemcpyDeviceToHost,
                              h->stream));
    SGIC_CUDA(cudaStreamSynchronize(h->stream));
    std::memcpy(reinterpret_cast<uint8_t*>(host_out) + static_cast<size_t>(done) * row_bytes, h->pin[0],
                static_cast<size_t>(rows) * row_bytes);
  }
  return 0;
}

int sgic_index_write(sgic_index* h, const char* path) {
  SGIC_REQUIRE(h != nullptr && path != nullptr, "NULL argument");
  FILE* f = std::fopen(path, "wb");
  if (!f) {
    set_error(std::string("could not open ") + path + " for writing");  // faiss: "could not open %s for writing"
    return 3;
  }
  IxfiHeader hd;
  std::memcpy(hd.fourcc, "IxFI", 4);
  hd.d = h->d;
  hd.ntotal = h->ntotal;
  hd.dummy0 = 1 << 20;
  hd.dummy1 = 1 << 20;
  hd.is_trained = 1;
  hd.metric_type = 0;  // METRIC_INNER_PRODUCT
  hd.count = static_cast<uint64_t>(h->ntotal) * h->d;
  bool ok = std::fwrite(&hd, sizeof(hd), 1, f) == 1;
  if (is_front(h)) {  // rows in global order, from the shards' host copies when all of them still hold theirs
    std::lock_guard<std::mutex> lk(h->mu);
    const int mode = front_write_mode(h);
    const int64_t chunk = std::max<int64_t>(1, static_cast<int64_t>((16u << 20) / (h->d * 4)));
    std::vector<float> buf(static_cast<size_t>(std::min<int64_t>(chunk, std::max<int64_t>(h->ntotal, 1))) * h->d);
    for (int64_t i0 = 0; ok && i0 < h->ntotal; i0 += chunk) {
      const int64_t rows = std::min(chunk, h->ntotal - i0);
      int rc = front_rows_for_write(h, mode, i0, rows, buf.data());
      if (rc) {
        std::fclose(f);
        return rc;
      }
      ok = std::fwrite(buf.data(), sizeof(float), static_cast<size_t>(rows) * h->d, f) == static_cast<size_t>(rows) * h->d;
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) {
      set_error(std::string("write error on ") + path);
      return 3;
    }
    return 0;
  }
  const bool from_host =
      h->retain_ok && h->retained.size() == static_cast<size_t>(h->ntotal) * static_cast<size_t>(h->d);
  if (ok && h->ntotal > 0) {
    const bool from_codes =
        h->codes_ok && h->codes.size() == static_cast<size_t>(h->ntotal) * static_cast<size_t>(h->d);
    if (from_host) {
      ok = std::fwrite(h->retained.data(), sizeof(float), h->retained.size(), f) == h->retained.size();
    } else if (from_codes) {
      // rows that came from clip_stream codes: what build.py:82-94 hands to faiss is dequantize_clip_u8(q) in
      // fp32, regenerated here with the reference's operation order — the file equals the reference's byte for byte
      const int64_t chunk = std::max<int64_t>(1, static_cast<int64_t>((16u << 20) / (h->d * 4)));
      std::vector<float> buf(static_cast<size_t>(std::min(chunk, h->ntotal)) * h->d);
      for (int64_t i0 = 0; ok && i0 < h->ntotal; i0 += chunk) {
        const int64_t rows = std::min(chunk, h->ntotal - i0);
        rows_from_codes_f32(h->codes.data() + static_cast<size_t>(i0) * h->d, rows, h->d, buf.data());
        ok = std::fwrite(buf.data(), sizeof(float), static_cast<size_t>(rows) * h->d, f) ==
             static_cast<size_t>(rows) * h->d;
      }
    } else {
      const int64_t chunk = std::max<int64_t>(1, static_cast<int64_t>((64u << 20) / (h->d * 4)));
      std::vector<float> buf(static_cast<size_t>(std::min(chunk, h->ntotal)) * h->d);
      for (int64_t i0 = 0; ok && i0 < h->ntotal; i0 += chunk) {
        const int64_t rows = std::min(chunk, h->ntotal - i0);
        int rc = sgic_index_reconstruct(h, i0, rows, buf.data());
        if (rc) {
          std::fclose(f);
          return rc;
        }
        ok = std::fwrite(buf.data(), sizeof(float), static_cast<size_t>(rows) * h->d, f) ==
             static_cast<size_t>(rows) * h->d;
      }
    }
  }
  ok = (std::fclose(f) == 0) && ok;
  if (!ok) {
    set_error(std::string("write error on ") + path);
    return 3;
  }
  return 0;
}


// ---- on-disk format v2 "SGI2"  ------------------------------------------------------------
// The rows exactly as they sit in HBM (fp16 / bf16, row-major) behind a 4 KB header page, so that loading a
// shard is read() into pinned memory + one H2D copy per chunk — half the bytes of the fp32 "IxFI" file and no
// conversion.  IxFI stays the interchange format (sgic_index_write); sgic_index_read recognises both.
struct Sgi2Header {
  char magic[4];          // "SGI2"
  uint32_t version;       // 1
  uint32_t dtype;         // SGIC_F16 | SGIC_BF16
  uint32_t d;
  int64_t ntotal;         // rows in this file
  int64_t row_start;      // global number of this file's first row
  int64_t total_rows;     // rows of the whole logical index
  uint32_t shard, n_shards;
  uint64_t payload_offset;  // 4096
  uint8_t reserved[8];
};
static_assert(sizeof(Sgi2Header) == 64, "SGI2 header is 64 bytes");
constexpr uint64_t kSgi2Payload = 4096;

int sgic_index_write_v2(sgic_index* h, const char* path, int64_t row_start, int64_t total_rows, int shard,
                        int n_shards) {
  SGIC_REQUIRE(h != nullptr && path != nullptr, "NULL argument");
  SGIC_REQUIRE(!is_front(h), "a multi-GPU index is written with sgic_index_save_shards (one SGI2 file per GPU)");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  int rc = ensure_staging(h);
  if (rc) return rc;
  if ((rc = order_begin(h, h->stream))) return rc;
  if ((rc = order_end(h, h->stream))) return rc;
  FILE* f = std::fopen(path, "wb");
  if (!f) {
    set_error(std::string("could not open ") + path + " for writing");
    return 3;
  }
  struct Closer {  // every early return below (SGIC_CUDA) closes the file
    FILE** f;
    ~Closer() {
      if (*f) std::fclose(*f);
    }
  } closer{&f};
  std::vector<uint8_t> page(kSgi2Payload, 0);
  Sgi2Header hd;
  std::memset(&hd, 0, sizeof(hd));
  std::memcpy(hd.magic, "SGI2", 4);
  hd.version = 1;
  hd.dtype = static_cast<uint32_t>(h->dtype);
  hd.d = static_cast<uint32_t>(h->d);
  hd.ntotal = h->ntotal;
  hd.row_start = row_start;
  hd.total_rows = total_rows < 0 ? h->ntotal : total_rows;
  hd.shard = static_cast<uint32_t>(shard);
  hd.n_shards = static_cast<uint32_t>(n_shards < 1 ? 1 : n_shards);
  hd.payload_offset = kSgi2Payload;
  std::memcpy(page.data(), &hd, sizeof(hd));
  bool ok = std::fwrite(page.data(), 1, page.size(), f) == page.size();
  // D2H through the pinned double buffer: the copy of chunk i+1 runs while chunk i is written
  const size_t total = elt_rows_bytes(h, h->ntotal);
  size_t issued = 0, written = 0;
  size_t len[2] = {0, 0};
  int b = 0;
  SGIC_CUDA(cudaStreamSynchronize(h->stream));
  if (total > 0) {
    len[0] = std::min(total, kStageChunkBytes);
    SGIC_CUDA(cudaMemcpyAsync(h->pin[0], static_cast<const uint8_t*>(h->db), len[0], cudaMemcpyDeviceToHost, h->stream));
    SGIC_CUDA(cudaEventRecord(h->ev[0], h->stream));
    issued = len[0];
  }
  while (ok && written < total) {
    if (issued < total) {
      len[b ^ 1] = std::min(total - issued, kStageChunkBytes);
      SGIC_CUDA(cudaMemcpyAsync(h->pin[b ^ 1], static_cast<const uint8_t*>(h->db) + issued, len[b ^ 1],
                                cudaMemcpyDeviceToHost, h->stream));
      SGIC_CUDA(cudaEventRecord(h->ev[b ^ 1], h->stream));
      issued += len[b ^ 1];
    }
    SGIC_CUDA(cudaEventSynchronize(h->ev[b]));
    ok = std::fwrite(h->pin[b], 1, len[b], f) == len[b];
    written += len[b];
    b ^= 1;
  }
  {
    FILE* ff = f;
    f = nullptr;
    ok = (std::fclose(ff) == 0) && ok;
  }
  if (!ok) {
    set_error(std::string("write error on ") + path);
    return 3;
  }
  return 0;
}

static int read_v2_body(FILE* f, const char* path, const Sgi2Header& hd, int device, int flags, sgic_index** out) {
  if (hd.version != 1 || hd.dtype > 1 || hd.d == 0 || hd.ntotal < 0 || hd.payload_offset < sizeof(Sgi2Header)) {
    set_error(std::string("corrupt SGI2 header in ") + path);
    return 3;
  }
  sgic_index* h = nullptr;
  int rc = sgic_index_create(static_cast<int>(hd.d), static_cast<int>(hd.dtype), device, hd.ntotal, flags & ~SGIC_RETAIN_F32, &h);
  if (rc) return rc;
  // (the index is destroyed only after the block that holds its mutex has been left)
  std::string fail_msg;
  int fail_code = 0;
  auto fail = [&](const std::string& msg, int code) {
    fail_msg = msg;
    fail_code = code;
  };
  [&] {
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    // Several readers, each with its own pair of pinned chunks and its own stream: chunk c is pread() by reader
    // c % T into pinned memory while that reader's previous chunk crosses PCIe.  One reader is bound by the copy out
    // of the page cache (~6 GB/s measured); four keep a PCIe 5 x16 link busy.
    const size_t total = elt_rows_bytes(h, hd.ntotal);
    const int T = static_cast<int>(std::min<size_t>(4, std::max<size_t>(1, total / (8u << 20))));
    // 4 MB chunks: pinning memory costs about as much per byte as reading it (the first version pinned 8 x 32 MB per
    // load and was slower than one reader); 8 x 4 MB keeps four preads and four H2D copies in flight
    const size_t chunk = std::min<size_t>(size_t(4) << 20,
                                          std::max<size_t>(size_t(1) << 20, ((total / (2 * static_cast<size_t>(T)) + 4095) & ~size_t(4095))));
    const size_t n_chunks = (total + chunk - 1) / chunk;
    const int fd = fileno(f);
    std::vector<int> trc(static_cast<size_t>(T), 0);
    std::vector<std::thread> readers;
    const int device = h->device;
    uint8_t* db = static_cast<uint8_t*>(h->db);
    const uint64_t payload = hd.payload_offset;
    for (int t = 0; t < T; ++t)
      readers.emplace_back([&, t] {
        cudaSetDevice(device);
        void* pin[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaStream_t st = nullptr;
        int& rc_t = trc[static_cast<size_t>(t)];
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) rc_t = 2;
        for (int b = 0; b < 2 && !rc_t; ++b)
          if (cudaMallocHost(&pin[b], chunk) != cudaSuccess || cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) != cudaSuccess)
            rc_t = 2;
        int b = 0;
        for (size_t c = static_cast<size_t>(t); c < n_chunks && !rc_t; c += static_cast<size_t>(T), b ^= 1) {
          const size_t off = c * chunk, n = std::min(chunk, total - off);
          if (cudaEventSynchronize(ev[b]) != cudaSuccess) {
            rc_t = 2;
            break;
          }
          size_t got = 0;
          while (got < n) {
            const ssize_t r = pread(fd, static_cast<uint8_t*>(pin[b]) + got, n - got, static_cast<off_t>(payload + off + got));
            if (r <= 0) break;
            got += static_cast<size_t>(r);
          }
          if (got != n) {
            rc_t = 3;
            break;
          }
          if (cudaMemcpyAsync(db + off, pin[b], n, cudaMemcpyHostToDevice, st) != cudaSuccess ||
              cudaEventRecord(ev[b], st) != cudaSuccess)
            rc_t = 2;
        }
        if (st) cudaStreamSynchronize(st);
        for (int i = 0; i < 2; ++i) {
          if (ev[i]) cudaEventDestroy(ev[i]);
          if (pin[i]) cudaFreeHost(pin[i]);
        }
        if (st) cudaStreamDestroy(st);
      });
    for (auto& th : readers) th.join();
    (void)cudaGetLastError();
    for (int t = 0; t < T; ++t) {
      if (trc[static_cast<size_t>(t)] == 3)
        return fail(std::string("read error in ") + path + ": file shorter than its header says", 3);
      if (trc[static_cast<size_t>(t)]) return fail("H2D copy failed while loading", 2);
    }
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return fail("stream synchronise failed while loading", 2);
    h->ntotal = hd.ntotal;
    h->retain_ok = false;
    drop_codes(h);
    h->shard_row_start = hd.row_start;
    h->shard_total_rows = hd.total_rows;
    h->shard_id = static_cast<int>(hd.shard);
    h->shard_count = static_cast<int>(hd.n_shards);
  }();
  if (fail_code) {
    sgic_index_destroy(h);
    set_error(fail_msg);
    return fail_code;
  }
  *out = h;
  return 0;
}

int sgic_index_shard_info(const sgic_index* h, int64_t* out4) {
  SGIC_REQUIRE(h != nullptr && out4 != nullptr, "NULL argument");
  if (is_front(h)) {
    out4[0] = 0;
    out4[1] = h->ntotal;
    out4[2] = 0;
    out4[3] = static_cast<int64_t>(h->shards.size());
    return 0;
  }
  out4[0] = h->shard_row_start;
  out4[1] = h->shard_total_rows < 0 ? h->ntotal : h->shard_total_rows;
  out4[2] = h->shard_id;
  out4[3] = h->shard_count;
  return 0;
}

int sgic_index_read(const char* path, int dtype, int device, int flags, sgic_index** out) {
  SGIC_REQUIRE(path != nullptr && out != nullptr, "NULL argument");
  *out = nullptr;
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    set_error(std::string("could not open ") + path + " for reading");
    return 3;
  }
  char magic[4] = {0, 0, 0, 0};
  if (std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "SGI2", 4) == 0) {  // v2 shard file: rows as stored in HBM
    Sgi2Header h2;
    std::rewind(f);
    if (std::fread(&h2, sizeof(h2), 1, f) != 1) {
      std::fclose(f);
      set_error(std::string("read error in ") + path + ": truncated header");
      return 3;
    }
    const int rc2 = read_v2_body(f, path, h2, device, flags, out);
    std::fclose(f);
    return rc2;
  }
  std::rewind(f);
  IxfiHeader hd;
  if (std::fread(&hd, sizeof(hd), 1, f) != 1) {
    std::fclose(f);
    set_error(std::string("read error in ") + path + ": truncated header");
    return 3;
  }
  if (std::memcmp(hd.fourcc, "IxFI", 4) != 0) {
    std::fclose(f);
    set_error(std::string("Index type 0x") + std::to_string(*reinterpret_cast<uint32_t*>(hd.fourcc)) +
              " not recognized: only IxFI (IndexFlatIP) is supported");
    return 3;
  }
  if (hd.d <= 0 || hd.ntotal < 0 || hd.count != static_cast<uint64_t>(hd.ntotal) * hd.d) {
    std::fclose(f);
    set_error(std::string("corrupt IxFI header in ") + path);
    return 3;
  }
  sgic_index* h = nullptr;
  const bool dbg_load = std::getenv("SGIC_DEBUG_LOAD") != nullptr;
  const auto t_load0 = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_load0).count(); };
  int rc = sgic_index_create(hd.d, dtype, device, hd.ntotal, flags, &h);
  if (rc) {
    std::fclose(f);
    return rc;
  }
  if (dbg_load) std::fprintf(stderr, "[load] index created + %lld rows mapped: %.1f ms\n", static_cast<long long>(hd.ntotal), since());
  // The fp32 rows come in through several readers (the reference re-reads this file in every query process,
  // src/search.py:69,76): reader t preads chunk c = t, t + T, ... into one of its two pinned buffers, copies it to
  // its own device staging buffer and packs it (K2) into the database rows the chunk stands for — all on the reader's
  // stream, so the pread of one chunk overlaps the H2D copy + pack of the previous one.  The host copy that makes
  // write_index bit-exact (SGIC_RETAIN_F32) is filled by the same readers.
  {
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard g(h->device);
    const size_t row_bytes = static_cast<size_t>(hd.d) * 4;
    const int64_t rows_per_chunk = std::max<int64_t>(1, static_cast<int64_t>((size_t(4) << 20) / row_bytes));
    const int64_t n_chunks = (hd.ntotal + rows_per_chunk - 1) / rows_per_chunk;
    const int T = static_cast<int>(std::min<int64_t>(4, std::max<int64_t>(1, n_chunks / 2)));
    const bool retain = h->retain_ok;
    bool retain_alloc = false;
    if (retain) {
      try {
        h->retained.resize(static_cast<size_t>(hd.ntotal) * hd.d);
        retain_alloc = true;
      } catch (const std::bad_alloc&) {
        h->retained.clear();
        h->retain_ok = false;
      }
    }
    const int fd = fileno(f);
    const size_t chunk_bytes = static_cast<size_t>(rows_per_chunk) * row_bytes;
    std::vector<int> trc(static_cast<size_t>(T), 0);
    std::vector<std::thread> readers;
    const int dev = h->device;
    for (int t = 0; t < T; ++t)
      readers.emplace_back([&, t] {
        cudaSetDevice(dev);
        int& rc_t = trc[static_cast<size_t>(t)];
        void* pin[2] = {nullptr, nullptr};
        void* stage[2] = {nullptr, nullptr};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        cudaStream_t st = nullptr;
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) rc_t = 2;
        for (int b = 0; b < 2 && !rc_t; ++b)
          if (cudaMallocHost(&pin[b], chunk_bytes) != cudaSuccess || cudaMalloc(&stage[b], chunk_bytes) != cudaSuccess ||
              cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) != cudaSuccess)
            rc_t = 2;
        if (dbg_load && t == 0) std::fprintf(stderr, "[load] reader 0 has its buffers: %.1f ms\n", since());
        int b = 0;
        for (int64_t c = t; c < n_chunks && !rc_t; c += T, b ^= 1) {
          const int64_t row0 = c * rows_per_chunk, rows = std::min(rows_per_chunk, hd.ntotal - row0);
          const size_t n = static_cast<size_t>(rows) * row_bytes;
          if (cudaEventSynchronize(ev[b]) != cudaSuccess) {
            rc_t = 2;
            break;
          }
          size_t got = 0;
          while (got < n) {
            const ssize_t r = pread(fd, static_cast<uint8_t*>(pin[b]) + got, n - got,
                                    static_cast<off_t>(sizeof(IxfiHeader) + static_cast<size_t>(row0) * row_bytes + got));
            if (r <= 0) break;
            got += static_cast<size_t>(r);
          }
          if (got != n) {
            rc_t = 3;
            break;
          }
          if (retain_alloc) std::memcpy(h->retained.data() + static_cast<size_t>(row0) * hd.d, pin[b], n);
          if (cudaMemcpyAsync(stage[b], pin[b], n, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            rc_t = 2;
            break;
          }
          // K2 on this reader's stream (launch_pack_f32 counts launches on the index: not from several threads)
          const size_t n8 = static_cast<size_t>(rows) * hd.d / 8;
          void* dst = static_cast<uint8_t*>(h->db) + static_cast<size_t>(row0) * hd.d * 2;
          const unsigned grid = grid_for(n8, 256, h->sm_count);
          if (h->dtype == SGIC_F16) pack_f32_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const float*>(stage[b]), dst, n8);
          else pack_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const float*>(stage[b]), dst, n8);
          if (cudaGetLastError() != cudaSuccess || cudaEventRecord(ev[b], st) != cudaSuccess) rc_t = 2;
        }
        if (st) cudaStreamSynchronize(st);
        if (dbg_load && t == 0) std::fprintf(stderr, "[load] reader 0 drained: %.1f ms\n", since());
        for (int i = 0; i < 2; ++i) {
          if (ev[i]) cudaEventDestroy(ev[i]);
          if (pin[i]) cudaFreeHost(pin[i]);
          if (stage[i]) cudaFree(stage[i]);
        }
        if (st) cudaStreamDestroy(st);
      });
    for (auto& th : readers) th.join();
    if (dbg_load) std::fprintf(stderr, "[load] %d readers done: %.1f ms\n", T, since());
    (void)cudaGetLastError();
    int bad = 0;
    for (int t = 0; t < T; ++t) bad = std::max(bad, trc[static_cast<size_t>(t)]);
    if (bad) {
      std::fclose(f);
      // (unlock before destroying: the guard goes out of scope with this block — destroy below)
      h->ntotal = 0;
    }
    if (!bad) {
      h->ntotal = hd.ntotal;
      h->stat_launches += n_chunks;
      drop_codes(h);
    }
    if (bad) {
      // fall through to the error return outside the lock
      rc = bad;
    }
  }
  if (rc) {
    sgic_index_destroy(h);
    if (rc == 3) set_error(std::string("read error in ") + path + ": file shorter than its header says");
    else set_error(std::string("loading ") + path + " onto the device failed");
    return rc;
  }
  std::fclose(f);
  *out = h;
  return 0;
}

int sgic_index_scan_times(sgic_index* h, int max_n, float* ms_out, int* n_out) {
  SGIC_REQUIRE(h != nullptr && ms_out != nullptr && n_out != nullptr && max_n >= 0, "bad arguments");
  if (is_front(h)) return sgic_index_scan_times(h->shards[0], max_n, ms_out, n_out);  // the home GPU's scans
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard g(h->device);
  const int64_t have = std::min<int64_t>(h->ring_n, sgic_index::kTimeRing);
  const int n = static_cast<int>(std::min<int64_t>(have, max_n));
  for (int i = 0; i < n; ++i) {  // oldest of the last n first
    const int slot = static_cast<int>((h->ring_n - n + i) % sgic_index::kTimeRing);
    SGIC_CUDA(cudaEventSynchronize(h->ring_b[slot]));
    SGIC_CUDA(cudaEventElapsedTime(&ms_out[i], h->ring_a[slot], h->ring_b[slot]));
  }
  *n_out = n;
  h->ring_n = 0;
  return 0;
}

int sgic_index_codes(sgic_index* h, int64_t i0, int64_t n, uint8_t* host_out) {
  SGIC_REQUIRE(h != nullptr && (host_out != nullptr || n == 0), "NULL argument");
  if (is_front(h)) return front_codes(h, i0, n, host_out);
  std::lock_guard<std::mutex> lk(h->mu);
  SGIC_REQUIRE(h->codes_ok && h->codes.size() == static_cast<size_t>(h->ntotal) * static_cast<size_t>(h->d),
               "this index does not hold the u8 codes of all its rows (SGIC_RETAIN_U8, rows added as codes only)");
  SGIC_REQUIRE(i0 >= 0 && n >= 0 && i0 + n <= h->ntotal, "row range out of bounds");
  std::memcpy(host_out, h->codes.data() + static_cast<size_t>(i0) * h->d, static_cast<size_t>(n) * h->d);
  return 0;
}

int sgic_codes_to_f32(const uint8_t* host_q, int64_t n, int d, float* host_out) {
  SGIC_REQUIRE(n >= 0 && d > 0 && (n == 0 || (host_q != nullptr && host_out != nullptr)), "bad arguments");
  rows_from_codes_f32(host_q, n, d, host_out);
  return 0;
}

int sgic_index_set_option(sgic_index* h, const char* name, int64_t value) {
  SGIC_REQUIRE(h != nullptr && name != nullptr, "NULL argument");
  const std::string n(name);
  if (is_front(h)) return front_set_option(h, n, value);
  if (n == "timing") h->opt_timing = value;
  else if (n == "evict_first") h->opt_evict_first = value;
  else if (n == "grid") h->opt_grid = value;
  else if (n == "stages") h->opt_stages = value;
  else if (n == "rb") h->opt_rb = value;
  else if (n == "fused") h->opt_fused = value;
  else if (n == "dense_min_nq") h->opt_dense_min_nq = value;
  else if (n == "debug") h->opt_debug = value;
  else if (n == "fail_appends") h->opt_fail_appends = std::max<int64_t>(0, value);
  else if (n == "dense_mode") h->opt_dense_mode = value;
  else if (n == "device_zstd") h->opt_device_zstd = value;
  else if (n == "dense_l2_mb") h->opt_dense_l2_mb = std::max<int64_t>(0, value);
  else if (n == "dense_b_min_mb") h->opt_dense_b_min_mb = std::max<int64_t>(0, value);
  else if (n == "dense_gthr") h->opt_dense_gthr = value ? 1 : 0;
  else if (n == "steal") h->opt_steal = value ? 1 : 0;
  else if (n == "t_n8") h->opt_t_n8 = value ? 1 : 0;
  else if (n == "t_max_nq") h->opt_t_max_nq = std::max<int64_t>(0, value);
  else if (n == "epi_wait_ns") h->opt_epi_wait_ns = std::max<int64_t>(0, std::min<int64_t>(10000, value));
  else if (n == "scan_wait_ns") h->opt_scan_wait_ns = std::max<int64_t>(0, std::min<int64_t>(10000, value));
  else if (n == "dense_slices") h->opt_dense_slices = std::max<int64_t>(0, value);
  else if (n == "dense_seed") h->opt_dense_seed = std::max<int64_t>(0, std::min<int64_t>(2, value));  // 2: force (tests)
  else if (n == "trace") h->opt_trace = value;  // device address of a [grid][4] u64 buffer (0 = off), see scan_small.cuh
  else if (n == "drop_retained") {
    std::lock_guard<std::mutex> lk(h->mu);
    h->retain_ok = false;
    h->retained.clear();
    h->retained.shrink_to_fit();
  }
  else {
    set_error("unknown option " + n);
    return 1;
  }
  return 0;
}

int64_t sgic_index_get_stat(const sgic_index* h, const char* name) {
  if (!h || !name) return -1;
  const std::string n(name);
  if (is_front(h)) {
    bool handled = false;
    const int64_t v = front_get_stat(h, n, &handled);
    if (handled) return v;
    return sgic_index_get_stat(h->shards[0], name);
  }
  if (n == "n_shards") return 1;
  if (n == "launches") return h->stat_launches;
  if (n == "last_search_ns") return h->stat_last_search_ns;
  if (n == "last_scan_ns") return h->stat_last_scan_ns;
  if (n == "last_grid") return h->stat_last_grid;
  if (n == "last_stages") return h->stat_last_stages;
  // scan kernel of the last search: 0 scan_small (K3), 1 scan_dense (1 CTA), 2 scan_dense_t (K4t), 3 scan_dense2
  // with the query tile resident, 4 scan_dense2 streamed, 5 scan_dense2b (database tile resident, K4b)
  if (n == "last_kernel") return h->stat_last_kernel;
  if (n == "capacity") return h->capacity;
  if (n == "sm_count") return h->sm_count;
  if (n == "dense_min_nq") return auto_dense_min_nq(h);
  if (n == "ingest_h2d_ns") return h->stat_ingest_h2d_ns;
  if (n == "ingest_k0_ns") return h->stat_ingest_k0_ns;
  if (n == "ingest_k1_ns") return h->stat_ingest_k1_ns;
  if (n == "ingest_parse_ns") return h->stat_ingest_parse_ns;
  if (n == "ingest_pack_ns") return h->stat_ingest_pack_ns;
  if (n == "ingest_gpu_ns") return h->stat_ingest_gpu_ns;
  if (n == "zl_device_frames") return h->stat_zl_device_frames;
  if (n == "zl_host_rows") return h->stat_zl_host_rows;
  if (n == "zl_fallback_slabs") return h->stat_zl_fallback_slabs;
  if (n == "retained_rows") return h->retain_ok ? static_cast<int64_t>(h->retained.size() / h->d) : -1;
  if (n == "retained_code_rows") return h->codes_ok ? static_cast<int64_t>(h->codes.size() / h->d) : -1;
  return -1;
}

}  // extern "C"

#include "sharded_front.inl"

namespace sgic {
static int front_reset(sgic_index* f) {
  std::lock_guard<std::mutex> lk(f->mu);
  for (sgic_index* s : f->shards) {
    int rc = sgic_index_reset(s);
    if (rc) return rc;
  }
  FrontState* S = fs(f);
  for (size_t g = 0; g < S->segs.size(); ++g) {
    S->segs[g].clear();
    S->seg_dirty[g] = 1;
  }
  f->ntotal = 0;
  return 0;
}
static int front_set_option(sgic_index* f, const std::string& name, int64_t value) {
  if (name == "workers") {  // 1: one launch thread per GPU (default); 0: the caller launches every shard itself
    fs(f)->opt_workers = value ? 1 : 0;
    return 0;
  }
  for (sgic_index* s : f->shards) {
    int rc = sgic_index_set_option(s, name.c_str(), value);
    if (rc) return rc;
  }
  return 0;
}
static int64_t front_get_stat(const sgic_index* f, const std::string& name, bool* handled) {
  *handled = true;
  if (name == "n_shards") return static_cast<int64_t>(f->shards.size());
  if (name == "launches") {  // the shards' kernels + the merges on the home GPU
    int64_t v = fs(f)->stat_merge_launches;
    for (const sgic_index* s : f->shards) v += s->stat_launches;
    return v;
  }
  if (name == "capacity") {
    int64_t v = 0;
    for (const sgic_index* s : f->shards) v += s->capacity;
    return v;
  }
  if (name == "n_segments") {
    int64_t v = 0;
    for (const auto& sv : fs(f)->segs) v += static_cast<int64_t>(sv.size());
    return v;
  }
  if (name == "direct_shards") {
    int64_t v = 0;
    for (uint8_t d : fs(f)->direct) v += d;
    return v;
  }
  if (name == "retained_rows") return front_write_mode(f) == 1 ? f->ntotal : -1;
  if (name == "retained_code_rows") return front_write_mode(f) == 2 ? f->ntotal : -1;
  if (name == "zl_device_frames" || name == "zl_host_rows" || name == "zl_fallback_slabs") {
    int64_t v = 0;
    for (const sgic_index* s : f->shards) v += sgic_index_get_stat(s, name.c_str());
    return v;
  }
  *handled = false;
  return -1;
}
}  // namespace sgic

extern "C" {

int sgic_index_create_sharded(int d, int dtype, int n_dev, const int* dev_ids, int64_t capacity_rows, int flags,
                              sgic_index** out) {
  SGIC_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  SGIC_REQUIRE(n_dev >= 1 && n_dev <= 64 && dev_ids != nullptr, "n_dev must be 1..64 with a device list");
  std::vector<sgic_index*> shards;
  for (int g = 0; g < n_dev; ++g) {
    sgic_index* s = nullptr;
    int rc = sgic_index_create(d, dtype, dev_ids[g], capacity_rows > 0 ? (capacity_rows + n_dev - 1) / n_dev : 0, flags, &s);
    if (rc) {
      for (sgic_index* t : shards) sgic_index_destroy(t);
      return rc;
    }
    shards.push_back(s);
  }
  sgic_index* f = new sgic_index();
  f->flags = flags;
  int rc = front_init(f, shards);
  if (rc) {
    sgic_index_destroy(f);
    return rc;
  }
  *out = f;
  return 0;
}

int sgic_index_n_shards(const sgic_index* h) { return h ? (is_front(h) ? static_cast<int>(h->shards.size()) : 1) : -1; }

int sgic_index_shard(sgic_index* h, int g, sgic_index** out) {
  SGIC_REQUIRE(h != nullptr && out != nullptr, "NULL argument");
  if (!is_front(h)) {
    SGIC_REQUIRE(g == 0, "shard number out of range");
    *out = h;
    return 0;
  }
  SGIC_REQUIRE(g >= 0 && g < static_cast<int>(h->shards.size()), "shard number out of range");
  *out = h->shards[static_cast<size_t>(g)];
  return 0;
}

int sgic_index_adopt_shards(sgic_index* h) {
  SGIC_REQUIRE(h != nullptr, "index is NULL");
  if (!is_front(h)) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  front_adopt_contiguous(h);
  return 0;
}

static std::string shard_file(const char* dir, int g, int G) {
  char name[64];
  std::snprintf(name, sizeof(name), "shard-%05d-of-%05d.sgi2", g, G);
  return std::string(dir) + "/" + name;
}

int sgic_index_save_shards(sgic_index* h, const char* dir) {
  SGIC_REQUIRE(h != nullptr && dir != nullptr, "NULL argument");
  if (!is_front(h)) return sgic_index_write_v2(h, shard_file(dir, 0, 1).c_str(), 0, h->ntotal, 0, 1);
  std::lock_guard<std::mutex> lk(h->mu);
  FrontState* S = fs(h);
  const int G = static_cast<int>(h->shards.size());
  for (int g = 0; g < G; ++g)
    SGIC_REQUIRE(S->segs[static_cast<size_t>(g)].size() <= 1,
                 "save needs one contiguous row range per GPU (reserve + bulk load, or large appends)");
  return for_shards(h, true, [&](int g) -> int {
    const auto& v = S->segs[static_cast<size_t>(g)];
    return sgic_index_write_v2(h->shards[static_cast<size_t>(g)], shard_file(dir, g, G).c_str(), v.empty() ? 0 : v[0].g0,
                               h->ntotal, g, G);
  });
}

int sgic_index_load_shards(const char* dir, int n_dev, const int* dev_ids, int flags, sgic_index** out) {
  SGIC_REQUIRE(dir != nullptr && out != nullptr && n_dev >= 1 && n_dev <= 64 && dev_ids != nullptr, "bad arguments");
  *out = nullptr;
  std::vector<sgic_index*> shards(static_cast<size_t>(n_dev), nullptr);
  std::vector<int> rcs(static_cast<size_t>(n_dev), 0);
  std::vector<std::string> errs(static_cast<size_t>(n_dev));
  {
    std::vector<std::thread> th;  // one reader per GPU: every shard streams over its own PCIe link
    for (int g = 0; g < n_dev; ++g)
      th.emplace_back([&, g] {
        rcs[static_cast<size_t>(g)] = sgic_index_read(shard_file(dir, g, n_dev).c_str(), SGIC_F16, dev_ids[g], flags,
                                                       &shards[static_cast<size_t>(g)]);
        if (rcs[static_cast<size_t>(g)]) errs[static_cast<size_t>(g)] = g_err;
      });
    for (auto& t : th) t.join();
  }
  auto fail = [&](int rc, const std::string& msg) {
    for (sgic_index* s : shards) sgic_index_destroy(s);
    set_error(msg);
    return rc;
  };
  for (int g = 0; g < n_dev; ++g)
    if (rcs[static_cast<size_t>(g)]) return fail(rcs[static_cast<size_t>(g)], errs[static_cast<size_t>(g)]);
  int64_t start = 0;
  for (int g = 0; g < n_dev; ++g) {
    const sgic_index* s = shards[static_cast<size_t>(g)];
    if (s->shard_count != n_dev || s->shard_id != g || s->d != shards[0]->d || s->dtype != shards[0]->dtype ||
        s->shard_row_start != start)
      return fail(3, std::string(dir) + ": shard files do not describe one index split over " + std::to_string(n_dev) +
                         " GPUs");
    start += s->ntotal;
  }
  sgic_index* f = new sgic_index();
  f->flags = flags;
  int rc = front_init(f, shards);
  if (rc) {
    sgic_index_destroy(f);
    return rc;
  }
  front_adopt_contiguous(f);
  *out = f;
  return 0;
}

}  // extern "C"
