// zipnn_b200.cu -- C ABI (include/zipnn_b200.h) over the sm_100a kernels.
//
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
//              -Xcompiler -fPIC -shared -o libzipnn_b200.so zipnn_b200.cu
#include "../../include/zipnn_b200.h"

#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "decode.cuh"
#include "encode.cuh"
#include "stage1.cuh"

using namespace zb;

namespace {

std::atomic<int> g_last_cuda_error{0};
std::atomic<unsigned long long> g_launches{0};

inline bool cuda_ok(cudaError_t e) {
  if (e != cudaSuccess) {
    g_last_cuda_error.store((int)e);
    return false;
  }
  return true;
}
#define ZB_CUDA(x)                          \
  do {                                      \
    if (!cuda_ok((x))) return ZIPNN_B200_E_CUDA; \
  } while (0)
#define ZB_LAUNCHED()                                       \
  do {                                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);     \
    if (!cuda_ok(cudaGetLastError())) return ZIPNN_B200_E_CUDA; \
  } while (0)

// ---- optional per-kernel timing (CUDA events on the launching stream) ----------------
// Off by default.  bench.py turns it on to attribute the step time to kernels; the events
// sit between launches on the same stream, so they do not change the schedule.
enum KernelId { kKDecodeMeta = 0, kKHufDecode, kKHufDecodePlanar, kKRegroup, kKEncodeStats, kKEncodeTable, kKEncodeScan, kKEncodeWrite, kKEncodeWriteRagged, kKSplit,
                kKRegroupPlanar, kKCount };
const char* const kKernelNames[kKCount] = {"k_decode_meta", "k_huf_decode_fused", "k_huf_decode_planar", "k_regroup", "k_encode_hist", "k_encode_table",
                                           "k_encode_scan", "k_encode_write_warp", "k_encode_write_ragged", "k_split_planar", "k_regroup_planar"};
struct TimedSpan {
  int id;
  cudaEvent_t a, b;
};
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
std::vector<TimedSpan> g_spans;
std::vector<cudaEvent_t> g_event_pool;

cudaEvent_t take_event() {
  cudaEvent_t e = nullptr;
  if (!g_event_pool.empty()) {
    e = g_event_pool.back();
    g_event_pool.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}
struct ScopedTimer {
  int id;
  cudaStream_t st;
  cudaEvent_t a = nullptr;
  ScopedTimer(int id_, cudaStream_t st_) : id(id_), st(st_) {
    if (g_timing.load(std::memory_order_relaxed)) {
      std::lock_guard<std::mutex> lk(g_timing_mu);
      a = take_event();
      cudaEventRecord(a, st);
    }
  }
  ~ScopedTimer() {
    if (a) {
      std::lock_guard<std::mutex> lk(g_timing_mu);
      cudaEvent_t b = take_event();
      cudaEventRecord(b, st);
      g_spans.push_back({id, a, b});
    }
  }
};

int sm_count_cached() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

inline bool valid_layout(int num_buf, int bytes_mode, size_t chunk) {
  if (!(num_buf == 1 || num_buf == 2 || num_buf == 4)) return false;
  // reference: mode 10 for one or two groups (dtype16.c:44,81), 220 for four (dtype32.c:241)
  if (num_buf == 4 ? bytes_mode != 220 : bytes_mode != 10) return false;
  if (chunk == 0 || (chunk & (chunk - 1)) != 0 || chunk > (1ull << 31)) return false;
  if (chunk % (size_t)num_buf) return false;
  return true;
}

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline uint64_t num_chunks(size_t n, size_t chunk) { return (n + chunk - 1) / chunk; }

// ---- decompress workspace layout ----
//   [Ctrl 256][ItemDesc G*K][mode u8 K][slot u32 K][fill 64*G*K][planes slots*G*pstride]
// `planes` is only used by chunks the fused kernel cannot take (several coded groups, or the
// ragged last chunk); the default size provides kDefaultSlots of them, the "full" size K.
constexpr uint64_t kDefaultSlots = 64;
struct DecWs {
  size_t items_off, mode_off, slot_off, rlist_off, fill_off, planes_off, pstride, fixed;
};
inline DecWs dec_ws_layout(size_t orig, int G, size_t chunk) {
  DecWs L;
  const uint64_t K = num_chunks(orig, chunk);
  L.items_off = kCtrlBytes;
  L.mode_off = round_up(L.items_off + sizeof(ItemDesc) * (size_t)G * K, 256);
  L.slot_off = round_up(L.mode_off + K, 256);
  L.rlist_off = round_up(L.slot_off + 4 * K, 256);
  L.fill_off = round_up(L.rlist_off + 4 * K, 256);
  L.planes_off = round_up(L.fill_off + (size_t)kFillBytes * G * K, 256);
  L.pstride = round_up(chunk / (size_t)G, 16) + 16;
  L.fixed = L.planes_off + 256;
  return L;
}

template <typename F>
int dispatch_G(int G, F&& f) {
  switch (G) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    default: return f(std::integral_constant<int, 4>());
  }
}

int read_ctrl_error(void* d_ws, cudaStream_t st) {
  uint32_t err = 0;
  ZB_CUDA(cudaMemcpyAsync(&err, d_ws, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ZB_CUDA(cudaStreamSynchronize(st));
  if (err & kErrCorrupt) return ZIPNN_B200_E_CORRUPT;
  if (err & kErrUnsupported) return ZIPNN_B200_E_UNSUPPORTED;
  if (err & kErrWorkspace) return ZIPNN_B200_E_CAPACITY;
  return ZIPNN_B200_OK;
}

}  // namespace

extern "C" {

int zipnn_b200_version(void) { return 0x000100; }

const char* zipnn_b200_strerror(int s) {
  switch (s) {
    case ZIPNN_B200_OK: return "ok";
    case ZIPNN_B200_E_ARG: return "invalid argument";
    case ZIPNN_B200_E_CAPACITY: return "output or workspace too small";
    case ZIPNN_B200_E_CORRUPT: return "corrupt ZipNN stream";
    case ZIPNN_B200_E_CUDA: return "CUDA runtime error";
    case ZIPNN_B200_E_UNSUPPORTED: return "unsupported stream feature (Huffman table log 12)";
    default: return "unknown status";
  }
}

int zipnn_b200_last_cuda_error(void) { return g_last_cuda_error.load(); }
int zipnn_b200_sm_count(void) { return sm_count_cached(); }
unsigned long long zipnn_b200_launch_count(void) { return g_launches.load(); }

void zipnn_b200_timing_enable(int on) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  g_timing.store(on ? 1 : 0);
  for (auto& sp : g_spans) {
    g_event_pool.push_back(sp.a);
    g_event_pool.push_back(sp.b);
  }
  g_spans.clear();
}

int zipnn_b200_timing_kernel_count(void) { return kKCount; }
const char* zipnn_b200_timing_kernel_name(int id) { return (id >= 0 && id < kKCount) ? kKernelNames[id] : ""; }

int zipnn_b200_timing_collect(double* ms_total, unsigned long long* launches, int n) {
  if (!ms_total || !launches || n < kKCount) return ZIPNN_B200_E_ARG;
  ZB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int i = 0; i < n; i++) {
    ms_total[i] = 0;
    launches[i] = 0;
  }
  for (auto& sp : g_spans) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
      ms_total[sp.id] += ms;
      launches[sp.id] += 1;
    }
    g_event_pool.push_back(sp.a);
    g_event_pool.push_back(sp.b);
  }
  g_spans.clear();
  return ZIPNN_B200_OK;
}

int zipnn_b200_compress_bound(size_t n, int num_buf, size_t chunk, size_t hdr_len, size_t* out) {
  if (!out || chunk == 0 || !(num_buf == 1 || num_buf == 2 || num_buf == 4)) return ZIPNN_B200_E_ARG;
  *out = hdr_len + 9 * (size_t)num_buf * num_chunks(n, chunk) + n;
  return ZIPNN_B200_OK;
}

int zipnn_b200_decompress_workspace_size(size_t orig, int num_buf, size_t chunk, size_t* out) {
  if (!out || chunk == 0 || !(num_buf == 1 || num_buf == 2 || num_buf == 4)) return ZIPNN_B200_E_ARG;
  const DecWs L = dec_ws_layout(orig, num_buf, chunk);
  const uint64_t K = num_chunks(orig, chunk);
  *out = L.fixed + (size_t)std::min<uint64_t>(K, kDefaultSlots) * num_buf * L.pstride;
  return ZIPNN_B200_OK;
}

int zipnn_b200_decompress_workspace_size_full(size_t orig, int num_buf, size_t chunk, size_t* out) {
  if (!out || chunk == 0 || !(num_buf == 1 || num_buf == 2 || num_buf == 4)) return ZIPNN_B200_E_ARG;
  const DecWs L = dec_ws_layout(orig, num_buf, chunk);
  *out = L.fixed + (size_t)num_chunks(orig, chunk) * num_buf * L.pstride;
  return ZIPNN_B200_OK;
}

int zipnn_b200_decompress(const void* d_body, size_t body_len, int num_buf, int bits_mode, int bytes_mode,
                          size_t chunk, size_t orig, void* d_out, void* d_ws, size_t ws_bytes, void* cuda_stream,
                          int check) {
  if (!valid_layout(num_buf, bytes_mode, chunk)) return ZIPNN_B200_E_ARG;
  if (orig == 0) return ZIPNN_B200_OK;
  if (!d_body || !d_out || !d_ws) return ZIPNN_B200_E_ARG;
  if (((uintptr_t)d_out & 15) || ((uintptr_t)d_ws & 255)) return ZIPNN_B200_E_ARG;
  const int G = num_buf;
  const uint64_t K = num_chunks(orig, chunk);
  if (body_len < 9ull * G * K) return ZIPNN_B200_E_CORRUPT;
  const DecWs L = dec_ws_layout(orig, G, chunk);
  if (ws_bytes < L.fixed) return ZIPNN_B200_E_CAPACITY;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  uint8_t* ws = (uint8_t*)d_ws;
  DecodeCfg cfg;
  cfg.body = (const uint8_t*)d_body;
  cfg.body_len = body_len;
  cfg.G = G;
  cfg.K = K;
  cfg.chunk = (uint32_t)chunk;
  cfg.orig = orig;
  cfg.bits_mode = bits_mode;
  cfg.ctrl = (Ctrl*)ws;
  cfg.items = (ItemDesc*)(ws + L.items_off);
  cfg.mode = ws + L.mode_off;
  cfg.slot = (uint32_t*)(ws + L.slot_off);
  cfg.rlist = (uint32_t*)(ws + L.rlist_off);
  cfg.fill = ws + L.fill_off;
  cfg.planes = ws + L.planes_off;
  cfg.pstride = L.pstride;
  cfg.tail_cap = 0;
  cfg.max_slots = (uint32_t)std::min<uint64_t>((ws_bytes - L.fixed) / ((size_t)G * L.pstride), K);
  const uint64_t nitems = (uint64_t)G * K;

  ZB_CUDA(cudaMemsetAsync(cfg.ctrl, 0, kCtrlBytes, st));
  {
    const int threads = 128;
    const int blocks = (int)std::min<uint64_t>((K + threads - 1) / threads, 4096);
    ScopedTimer tm(kKDecodeMeta, st);
    k_decode_meta<<<blocks, threads, 0, st>>>(cfg);
    ZB_LAUNCHED();
  }
  {
    const uint64_t warps = (K + kDecItemsPerWarp - 1) / kDecItemsPerWarp;
    if (warps > 0x7fffffffull) return ZIPNN_B200_E_ARG;
    // Short-code planes (the exponent plane of the rotated types, ~2.6 bits per symbol) use
    // conflict-free private 5-bit table columns and a 1024-entry tail pool (8 chunks x ~64 entries
    // for the codes longer than 5 bits, so 2x slack): 13 KiB per warp with one side plane, 16 warps
    // per SM.  fp16 / fp8 planes (6-7 bits per symbol, 90-150 entries of > 8 bits per chunk) keep the
    // shared 8-bit primaries.  A chunk whose tail does not fit the pool takes the general path.
    const bool short_codes = (G >= 2 && bits_mode == 1);
    cfg.tail_cap = short_codes ? 1024u : 2048u;
    ScopedTimer tm(kKHufDecode, st);
    int rc = dispatch_G(G, [&](auto g) -> int {
      constexpr int GG = decltype(g)::value;
      if (short_codes)
        k_huf_decode_fused<GG, 5><<<(unsigned)warps, 32, fused_smem_bytes(cfg.tail_cap, 5, GG), st>>>(cfg, (uint8_t*)d_out);
      else
        k_huf_decode_fused<GG, 0><<<(unsigned)warps, 32, fused_smem_bytes(cfg.tail_cap, 0, GG), st>>>(cfg, (uint8_t*)d_out);
      ZB_LAUNCHED();
      return ZIPNN_B200_OK;
    });
    if (rc) return rc;
  }
  {
    const uint64_t warps = (nitems + kDecItemsPerWarp - 1) / kDecItemsPerWarp;
    if (warps > 0x7fffffffull) return ZIPNN_B200_E_ARG;
    ScopedTimer tm(kKHufDecodePlanar, st);
    k_huf_decode_planar<<<(unsigned)warps, 32, sizeof(DecodeSmem), st>>>(cfg);
    ZB_LAUNCHED();
  }
  {
    const uint32_t tiles_per_chunk = (uint32_t)((chunk + kMergeTile - 1) / kMergeTile);
    const uint64_t ntiles = K * tiles_per_chunk;
    const int blocks = (int)std::min<uint64_t>(ntiles, (uint64_t)sm_count_cached() * 16);
    ScopedTimer tm(kKRegroup, st);
    int rc = dispatch_G(G, [&](auto g) -> int {
      k_regroup<decltype(g)::value><<<blocks, kMergeThreads, 0, st>>>(cfg, (uint8_t*)d_out);
      ZB_LAUNCHED();
      return ZIPNN_B200_OK;
    });
    if (rc) return rc;
  }
  if (check) return read_ctrl_error(d_ws, st);
  return ZIPNN_B200_OK;
}

int zipnn_b200_split(const void* d_in, size_t n, int num_buf, int bits_mode, void* d_planes, size_t stride,
                     void* cuda_stream) {
  if (!(num_buf == 1 || num_buf == 2 || num_buf == 4)) return ZIPNN_B200_E_ARG;
  if (n == 0) return ZIPNN_B200_OK;
  if (!d_in || !d_planes || ((uintptr_t)d_in & 15) || ((uintptr_t)d_planes & 15) || (stride & 15)) return ZIPNN_B200_E_ARG;
  if (stride < (n + num_buf - 1) / num_buf) return ZIPNN_B200_E_CAPACITY;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const uint64_t units = n / (16ull * num_buf);
  const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((units + kStage1Threads - 1) / kStage1Threads,
                                                                    (uint64_t)sm_count_cached() * 32));
  ScopedTimer tm(kKSplit, st);
  return dispatch_G(num_buf, [&](auto g) -> int {
    k_split_planar<decltype(g)::value><<<blocks, kStage1Threads, 0, st>>>((const uint8_t*)d_in, n, bits_mode,
                                                                          (uint8_t*)d_planes, stride);
    ZB_LAUNCHED();
    return ZIPNN_B200_OK;
  });
}

int zipnn_b200_regroup(const void* d_planes, size_t stride, size_t n, int num_buf, int bits_mode, void* d_out,
                       void* cuda_stream) {
  if (!(num_buf == 1 || num_buf == 2 || num_buf == 4)) return ZIPNN_B200_E_ARG;
  if (n == 0) return ZIPNN_B200_OK;
  if (!d_out || !d_planes || ((uintptr_t)d_out & 15) || ((uintptr_t)d_planes & 15) || (stride & 15)) return ZIPNN_B200_E_ARG;
  if (stride < (n + num_buf - 1) / num_buf) return ZIPNN_B200_E_CAPACITY;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const uint64_t units = n / (16ull * num_buf);
  const int blocks = (int)std::max<uint64_t>(1, std::min<uint64_t>((units + kStage1Threads - 1) / kStage1Threads,
                                                                    (uint64_t)sm_count_cached() * 32));
  ScopedTimer tm(kKRegroupPlanar, st);
  return dispatch_G(num_buf, [&](auto g) -> int {
    k_regroup_planar<decltype(g)::value><<<blocks, kStage1Threads, 0, st>>>((const uint8_t*)d_planes, stride, n,
                                                                            bits_mode, (uint8_t*)d_out);
    ZB_LAUNCHED();
    return ZIPNN_B200_OK;
  });
}

}  // extern "C"

#include "api_compress.inc"
#include "api_host.inc"
