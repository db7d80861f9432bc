// Host side of host-buffer calls: a pinned bounce ring drained by host threads (declared in blmm_ctx.cuh).
//
// Why it exists.  The callers of the C-ABI (a Julia `Array{Float64}`, a numpy array) hand the library ordinary
// pageable memory.  cudaMemcpyAsync to pageable memory is not asynchronous: the runtime stages it through an internal
// buffer on the calling thread, so the chunked copy-back of an alt-grid scan (2 x 2 GB at BXD size) would serialise
// with the scan launches and run at the speed of one memcpy thread.  Here the DMA goes device -> pinned ring slot at
// PCIe speed, and `nthreads` drain threads move each slot into the caller's array with non-temporal stores (no
// read-for-ownership of the destination lines: the host's DRAM write bandwidth is what several GPUs feeding one
// host end up sharing).  One-byte h2 grid indices travel through the same ring and are expanded to grid[index].
#include <emmintrin.h>
#include <string.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "blmm_ctx.cuh"

namespace blmm {

namespace {

constexpr size_t SLOT_BYTES_DEFAULT = 4u << 20;  // Float64 pieces: 4 MB of DMA, 4 MB of host copy per task
constexpr int NSLOTS_DEFAULT = 24;
constexpr int NSLOTS_MAX = 256;

struct Task {
  int slot;            // ring slot to release afterwards, or -1: the piece sits in caller-owned pinned staging
  const uint8_t* src;  // where the piece lands on the host (ring slot or staging)
  cudaEvent_t ev;      // recorded after the device-to-host copy that fills `src`
  double* dst;       // first destination element
  int64_t ld_dst;    // doubles between destination columns
  int64_t rows;      // elements per column in this piece
  int64_t cols;
  const double* grid;  // nullptr: Float64 copy; else expansion of one-byte indices
  void* in_dst = nullptr;  // host-to-host piece of an input upload (hostpipe_gather_input): memcpy(in_dst, src, rows bytes)
};

// dst[0..n) = src[0..n) with streaming stores (dst 8-byte aligned)
inline void copy_nt(double* dst, const double* src, int64_t n) {
  int64_t i = 0;
  if (n > 0 && ((uintptr_t)dst & 15)) {
    _mm_stream_si64(reinterpret_cast<long long*>(dst), reinterpret_cast<const long long*>(src)[0]);
    i = 1;
  }
  for (; i + 8 <= n; i += 8) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 2));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 6));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 2), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 6), d);
  }
  for (; i + 2 <= n; i += 2)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i)));
  if (i < n) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(src)[i]);
}

// dst[i] = grid[src[i]] with streaming stores
inline void expand_nt(double* dst, const uint8_t* src, int64_t n, const double* grid) {
  int64_t i = 0;
  if (n > 0 && ((uintptr_t)dst & 15)) {
    _mm_stream_si64(reinterpret_cast<long long*>(dst), reinterpret_cast<const long long*>(grid)[src[0]]);
    i = 1;
  }
  for (; i + 4 <= n; i += 4) {
    const __m128d a = _mm_set_pd(grid[src[i + 1]], grid[src[i]]);
    const __m128d b = _mm_set_pd(grid[src[i + 3]], grid[src[i + 2]]);
    _mm_stream_pd(dst + i, a);
    _mm_stream_pd(dst + i + 2, b);
  }
  for (; i < n; ++i) _mm_stream_si64(reinterpret_cast<long long*>(dst + i), reinterpret_cast<const long long*>(grid)[src[i]]);
}

}  // namespace

struct HostPipe {
  int device = 0;
  uint8_t* ring = nullptr;  // nslots x slot_bytes, pinned
  size_t slot_bytes = SLOT_BYTES_DEFAULT;
  int nslots = NSLOTS_DEFAULT;
  cudaEvent_t ev[NSLOTS_MAX] = {};
  std::mutex mu;
  std::condition_variable cv_task, cv_slot, cv_done;
  std::deque<Task> tasks;
  std::vector<int> free_slots;
  int64_t pending = 0;  // queued or in progress
  bool quit = false;
  std::string error;
  std::vector<std::thread> workers;

  int active = 1 << 30;  // workers with id >= active leave the queue alone (hostpipe_set_active)

  void worker(int id) {
    cudaSetDevice(device);
    for (;;) {
      Task t;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_task.wait(lk, [&] { return quit || (!tasks.empty() && id < active); });
        if (quit && (tasks.empty() || id >= active)) return;
        t = tasks.front();
        tasks.pop_front();
      }
      if (t.in_dst) {  // input upload: plain host copy into pinned staging, no event to wait for
        memcpy(t.in_dst, t.src, (size_t)t.rows);
        {
          std::lock_guard<std::mutex> lk(mu);
          --pending;
        }
        cv_done.notify_all();
        continue;
      }
      const cudaError_t e = cudaEventSynchronize(t.ev);
      if (e == cudaSuccess) {
        const uint8_t* src = t.src;
        for (int64_t c = 0; c < t.cols; ++c) {
          if (t.grid)
            expand_nt(t.dst + c * t.ld_dst, src + c * t.rows, t.rows, t.grid);
          else
            copy_nt(t.dst + c * t.ld_dst, reinterpret_cast<const double*>(src) + c * t.rows, t.rows);
        }
        _mm_sfence();
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (e != cudaSuccess && error.empty()) error = std::string("device-to-host copy: ") + cudaGetErrorString(e);
        if (t.slot >= 0) free_slots.push_back(t.slot);
        --pending;
      }
      cv_slot.notify_one();
      cv_done.notify_all();
    }
  }
};

bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int default_host_threads() {
  if (const char* ht = getenv("BLMM_B200_HOST_THREADS")) return std::max(1, std::min(64, atoi(ht)));
  const unsigned hc = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(16u, hc > 1 ? hc - 1 : 1u));
}

HostPipe* hostpipe_create(int device, int nthreads) {
  HostPipe* hp = new HostPipe();
  hp->device = device;
  // development knobs (tools/e2e_probe.py): ring geometry
  if (const char* e = getenv("BLMM_B200_RING_SLOT_KB")) hp->slot_bytes = (size_t)std::max(64, atoi(e)) << 10;
  if (const char* e = getenv("BLMM_B200_RING_SLOTS")) hp->nslots = std::max(4, std::min(NSLOTS_MAX, atoi(e)));
  const int NSLOTS = hp->nslots;
  if (cudaMallocHost(&hp->ring, (size_t)NSLOTS * hp->slot_bytes) != cudaSuccess) {
    cudaGetLastError();
    delete hp;
    throw Fail{BLMM_E_CUDA, "cudaMallocHost of the host result ring failed"};
  }
  for (int i = 0; i < NSLOTS; ++i) {
    // blocking sync: a drain thread waiting for its DMA sleeps instead of spinning on a core another drain thread
    // (or another GPU's) could be copying with
    if (cudaEventCreateWithFlags(&hp->ev[i], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) {
      hostpipe_destroy(hp);
      throw Fail{BLMM_E_CUDA, "cudaEventCreate failed"};
    }
    hp->free_slots.push_back(NSLOTS - 1 - i);
  }
  nthreads = std::max(1, std::min(nthreads, NSLOTS - 2));
  for (int i = 0; i < nthreads; ++i) hp->workers.emplace_back([hp, i] { hp->worker(i); });
  return hp;
}

void hostpipe_destroy(HostPipe* hp) {
  if (!hp) return;
  {
    std::lock_guard<std::mutex> lk(hp->mu);
    hp->quit = true;
  }
  hp->cv_task.notify_all();
  for (auto& t : hp->workers) t.join();
  for (int i = 0; i < NSLOTS_MAX; ++i)
    if (hp->ev[i]) cudaEventDestroy(hp->ev[i]);
  if (hp->ring) cudaFreeHost(hp->ring);
  delete hp;
}

void hostpipe_push(HostPipe* hp, cudaStream_t stream, double* dst, int64_t ld_dst, const void* src_dev, int64_t ld_src,
                   int64_t rows, int64_t cols, const double* grid) {
  if (rows <= 0 || cols <= 0) return;
  const size_t elem = grid ? 1 : 8;
  // index pieces expand 8x on the host: an eighth of a slot per piece keeps the work per task the same
  const size_t budget = grid ? std::max<size_t>(hp->slot_bytes / 8, 65536) : hp->slot_bytes;
  const int64_t rows_per_piece = std::min<int64_t>(rows, (int64_t)(budget / elem));
  const int64_t cols_per_piece = (rows_per_piece == rows) ? std::max<int64_t>(1, (int64_t)(budget / (rows * elem))) : 1;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(src_dev);
  for (int64_t c0 = 0; c0 < cols; c0 += cols_per_piece) {
    const int64_t nc = std::min(cols_per_piece, cols - c0);
    for (int64_t r0 = 0; r0 < rows; r0 += rows_per_piece) {
      const int64_t nr = std::min(rows_per_piece, rows - r0);
      int slot;
      {
        std::unique_lock<std::mutex> lk(hp->mu);
        hp->cv_slot.wait(lk, [&] { return !hp->free_slots.empty(); });
        slot = hp->free_slots.back();
        hp->free_slots.pop_back();
        ++hp->pending;
      }
      uint8_t* s = hp->ring + (size_t)slot * hp->slot_bytes;
      const uint8_t* from = src + ((size_t)c0 * ld_src + r0) * elem;
      const cudaError_t e1 =
          (ld_src == nr || nc == 1)
              ? cudaMemcpyAsync(s, from, (size_t)nr * nc * elem, cudaMemcpyDeviceToHost, stream)
              : cudaMemcpy2DAsync(s, nr * elem, from, ld_src * elem, nr * elem, nc, cudaMemcpyDeviceToHost, stream);
      const cudaError_t e2 = cudaEventRecord(hp->ev[slot], stream);
      if (e1 != cudaSuccess || e2 != cudaSuccess) {
        {
          std::lock_guard<std::mutex> lk(hp->mu);
          hp->free_slots.push_back(slot);
          --hp->pending;
        }
        hp->cv_done.notify_all();
        throw Fail{BLMM_E_CUDA, std::string("device-to-host copy: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2)};
      }
      {
        std::lock_guard<std::mutex> lk(hp->mu);
        hp->tasks.push_back(Task{slot, s, hp->ev[slot], dst + c0 * ld_dst + r0, ld_dst, nr, nc, grid});
      }
      hp->cv_task.notify_all();  // all: a sleeping worker beyond the active limit must not swallow the wake-up
    }
  }
}

void hostpipe_push_staged(HostPipe* hp, cudaEvent_t ev, double* dst, int64_t ld_dst, const void* staged, int64_t rows,
                          int64_t cols, const double* grid) {
  if (rows <= 0 || cols <= 0) return;
  const size_t elem = grid ? 1 : 8;
  // pieces of ~4 MB of destination stores each, so that every drain thread gets a share of every chunk
  const int64_t cols_per_piece = std::max<int64_t>(1, (int64_t)((4u << 20) / (rows * 8)));
  const uint8_t* src = reinterpret_cast<const uint8_t*>(staged);
  {
    std::lock_guard<std::mutex> lk(hp->mu);
    for (int64_t c0 = 0; c0 < cols; c0 += cols_per_piece) {
      const int64_t nc = std::min(cols_per_piece, cols - c0);
      hp->tasks.push_back(Task{-1, src + (size_t)c0 * rows * elem, ev, dst + c0 * ld_dst, ld_dst, rows, nc, grid});
      ++hp->pending;
    }
  }
  hp->cv_task.notify_all();
}

void hostpipe_gather_input(HostPipe* hp, void* pinned_dst, const void* src, size_t bytes) {
  constexpr size_t PIECE = 1u << 20;
  {
    std::lock_guard<std::mutex> lk(hp->mu);
    hp->active = 1 << 30;
    for (size_t off = 0; off < bytes; off += PIECE) {
      Task t{};
      t.slot = -1;
      t.src = static_cast<const uint8_t*>(src) + off;
      t.rows = (int64_t)std::min(PIECE, bytes - off);
      t.in_dst = static_cast<uint8_t*>(pinned_dst) + off;
      hp->tasks.push_back(t);
      ++hp->pending;
    }
  }
  hp->cv_task.notify_all();
  std::unique_lock<std::mutex> lk(hp->mu);
  hp->cv_done.wait(lk, [&] { return hp->pending == 0; });
}

void hostpipe_set_active(HostPipe* hp, int n) {
  if (!hp) return;
  {
    std::lock_guard<std::mutex> lk(hp->mu);
    hp->active = n < 1 ? 1 : n;
  }
  hp->cv_task.notify_all();
}

void hostpipe_wait(HostPipe* hp) {
  if (!hp) return;
  std::unique_lock<std::mutex> lk(hp->mu);
  hp->cv_done.wait(lk, [&] { return hp->pending == 0; });
  if (!hp->error.empty()) {
    std::string e = hp->error;
    hp->error.clear();
    throw Fail{BLMM_E_CUDA, e};
  }
}

// Streaming-store bandwidth of `nthreads` host threads into a pageable buffer of `bytes` (touched first, best of 3):
// the rate at which this host can take Float64 results at all — the ceiling of every host-buffer call next to PCIe.
double host_write_gbs(int nthreads, size_t bytes) {
  nthreads = std::max(1, std::min(nthreads, 256));
  const int64_t n = (int64_t)(bytes / 8);
  double* buf = static_cast<double*>(aligned_alloc(64, (size_t)n * 8));
  if (!buf) return -1.0;
  memset(buf, 0, (size_t)n * 8);
  std::vector<double> src(4096, 1.0);
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
      th.emplace_back([&, t] {
        const int64_t a = n * t / nthreads, b = n * (t + 1) / nthreads;
        for (int64_t c = a; c < b; c += 4096) copy_nt(buf + c, src.data(), std::min<int64_t>(4096, b - c));
        _mm_sfence();
      });
    for (auto& x : th) x.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    best = std::max(best, (double)n * 8 / sec / 1e9);
  }
  free(buf);
  return best;
}

}  // namespace blmm
