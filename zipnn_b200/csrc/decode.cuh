// decode.cuh -- decompress side:  stream metadata -> item table, Huffman bit-unpack,
// byte-group regroup (+ sign-bit un-rotate).
//
// Replaces reference csrc/zipnn_core.c:881-1142 (py_combine_dtype), :768-861
// (decompression_chunk_worker), huf_decompress.c:118-354 (table + 4-stream decode),
// data_manipulation_dtype16.c:167-216 and data_manipulation_dtype32.c:391-456 (combine).
#pragma once
#include "common.cuh"

namespace zb {

// ====================================================================================
// Kernel 1: parse + validate the per-(group,chunk) metadata, emit the item table.
// Stream body layout (csrc/zipnn_core.c:105-244):
//   types u8[G][K] | cum u64le[G][K] (inclusive, per group) | group-major payload
// ====================================================================================
__global__ void k_decode_meta(const uint8_t* __restrict__ body, uint64_t body_len, int G, uint64_t K,
                              uint32_t chunk, uint64_t orig, Ctrl* ctrl, ItemDesc* items) {
  const uint64_t nitems = (uint64_t)G * K;
  const uint8_t* types = body;
  const uint8_t* cum = body + nitems;
  const uint64_t payload0 = 9 * nitems;
  const uint64_t payload_len = body_len - payload0;
  uint64_t base[4] = {0, 0, 0, 0};
  for (int g = 1; g < G; g++) base[g] = base[g - 1] + ld_u64_bytes(cum + 8 * ((uint64_t)(g - 1) * K + (K - 1)));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int g = 0; g < 4; g++) ctrl->base[g] = payload0 + base[g];
    uint64_t all = base[G - 1] + ld_u64_bytes(cum + 8 * ((uint64_t)(G - 1) * K + (K - 1)));
    if (all > payload_len) atomicOr(&ctrl->error, kErrCorrupt);
  }
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nitems;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i / K);
    const uint64_t c = i - (uint64_t)g * K;
    const uint64_t hi = ld_u64_bytes(cum + 8 * i);
    const uint64_t lo = c ? ld_u64_bytes(cum + 8 * (i - 1)) : 0;
    const uint32_t chunk_len = (c == K - 1) ? (uint32_t)(orig - c * (uint64_t)chunk) : chunk;
    const uint32_t dlen = plane_len(chunk_len, G, g);
    const uint8_t type = types[i];
    ItemDesc d;
    d.src_off = payload0 + base[g] + lo;
    d.dec_len = dlen;
    d.pad = 0;
    bool bad = (hi < lo) || (base[g] + hi > payload_len) || (type > 1) || (hi - lo > 0xFFFFFFFFull);
    const uint32_t slen = (uint32_t)(hi - lo);
    d.src_len = slen;
    if (type == 0) {
      d.kind = kRaw;
      bad = bad || (slen != dlen);
    } else {
      // HUF_decompress (huf_decompress.c:1056-1081): csize > dst -> error; == -> copy; 1 -> RLE
      if (dlen == 0) {
        d.kind = kRaw;  // the reference never decodes an empty plane
        d.src_len = 0;
      } else if (slen > dlen || slen == 0) {
        bad = true;
        d.kind = kRaw;
      } else if (slen == dlen) {
        d.kind = kRaw;
      } else if (slen == 1) {
        d.kind = kRle;
      } else {
        d.kind = kHuf;
        bad = bad || (dlen > (uint32_t)kHufBlockMax);
      }
    }
    if (bad) {
      atomicOr(&ctrl->error, kErrCorrupt);
      d.kind = kRaw;
      d.src_len = 0;
      d.dec_len = 0;
    }
    items[i] = d;
  }
}

// ====================================================================================
// Kernel 2: Huffman decode of kHuf items into planar byte planes.
//
// One thread per bitstream (a huff0 block is 4 independent backward bitstreams,
// huf_decompress.c:283-298), one warp = 8 items.  Each item's single-symbol decode table
// (2^tableLog x {symbol, length}) lives in shared memory; the stream is read through a
// left-aligned 64-bit window refilled one aligned 32-bit word at a time with the next
// word prefetched, so the global-load latency is off the symbol-to-symbol chain.
// ====================================================================================
constexpr int kDecItemsPerWarp = 8;
constexpr int kDecLutLog = 11;  // the reference encoder never exceeds 11 (HUF_TABLELOG_DEFAULT)
constexpr int kDecLutEntries = 1 << kDecLutLog;

struct DecodeSmem {
  uint16_t lut[kDecItemsPerWarp][kDecLutEntries];  // also scratch for the table parse
  uint8_t weights[kDecItemsPerWarp][256];
};
static_assert(sizeof(FseDec) <= sizeof(uint16_t) * kDecLutEntries, "FseDec must fit in one LUT slot");

struct BitWindow {
  uint64_t w;           // unread bits, left aligned
  int avail;            // valid bits in w
  const uint32_t* wp;   // next word to prefetch (moves down)
  uint32_t nxt;         // prefetched word
  const uint32_t* wp0;  // wp right after init (for the exact-consumption check)
  int loaded0;          // bits in the window right after init
  uint32_t unread;      // bits between stream start and the end mark
};

__device__ __forceinline__ bool window_init(BitWindow& b, const uint8_t* s, uint32_t len, const uint8_t* lo,
                                            const uint8_t* hi) {
  const uint8_t lastb = s[len - 1];
  if (lastb == 0) return false;
  const uint64_t mark = 8ull * (uint64_t)(uintptr_t)(s + len - 1) + (uint64_t)hb32(lastb);  // end-mark bit
  b.unread = (uint32_t)(mark - 8ull * (uint64_t)(uintptr_t)s);
  if (b.unread == 0) return false;
  const uintptr_t top_byte = (uintptr_t)((mark - 1) >> 3);
  const uint32_t* wt = reinterpret_cast<const uint32_t*>(top_byte & ~(uintptr_t)3);
  const int k = (int)(mark - 8ull * (uint64_t)(uintptr_t)wt);  // 1..32 unread bits in the top word
  const uint32_t topw = ld_word_guarded(wt, lo, hi);
  const uint32_t low1 = ld_word_guarded(wt - 1, lo, hi);
  b.w = ((uint64_t)(topw << (32 - k)) << 32) | ((uint64_t)low1 << (32 - k));
  b.avail = k + 32;
  b.loaded0 = k + 32;
  b.nxt = ld_word_guarded(wt - 2, lo, hi);
  b.wp = wt - 3;
  b.wp0 = b.wp;
  return true;
}

__device__ __forceinline__ void window_refill(BitWindow& b, const uint8_t* lo_aligned) {
  if (b.avail <= 32) {
    b.w |= (uint64_t)b.nxt << (32 - b.avail);
    b.avail += 32;
    b.nxt = (reinterpret_cast<const uint8_t*>(b.wp) >= lo_aligned) ? __ldg(b.wp) : 0u;
    b.wp--;
  }
}

__device__ __forceinline__ uint32_t window_decode(BitWindow& b, const uint16_t* lut, int lg) {
  const uint32_t e = lut[(uint32_t)(b.w >> (64 - lg))];
  const int nb = (int)(e >> 8);
  b.w <<= nb;
  b.avail -= nb;
  return e & 0xFFu;
}

// Decode `count` symbols of one stream into dst (global).  Returns false if the stream
// was not consumed exactly.
__device__ __forceinline__ bool decode_stream_planar(BitWindow& b, const uint16_t* lut, int lg, uint8_t* dst,
                                                     uint32_t count, const uint8_t* lo_aligned) {
  uint32_t done = 0;
  if ((((uintptr_t)dst) & 15) == 0) {
    const uint32_t n16 = count >> 4;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (uint32_t it = 0; it < n16; it++) {
      uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        window_refill(b, lo_aligned);
        o[j >> 2] |= window_decode(b, lut, lg) << (8 * (j & 3));
        o[(j + 1) >> 2] |= window_decode(b, lut, lg) << (8 * ((j + 1) & 3));
      }
      d4[it] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    done = n16 << 4;
  }
  for (; done < count; done++) {
    window_refill(b, lo_aligned);
    dst[done] = (uint8_t)window_decode(b, lut, lg);
  }
  const int refills = (int)(b.wp0 - b.wp);
  const int consumed = b.loaded0 + 32 * refills - b.avail;
  return consumed == (int)b.unread;
}

// Serial single-symbol table fill, one lane per item (huf_decompress.c:151-183): weights
// ascending, symbols ascending within a weight, 2^(w-1) consecutive entries each.
__device__ __forceinline__ void fill_lut(uint16_t* lut, const uint8_t* weights, int nsym, int lg) {
  uint32_t cnt[kHufLogMax + 2];
#pragma unroll
  for (int i = 0; i < kHufLogMax + 2; i++) cnt[i] = 0;
  for (int n = 0; n < nsym; n++) cnt[weights[n]]++;
  uint32_t start[kHufLogMax + 2];
  uint32_t at = 0;
  start[0] = 0;
  for (int w = 1; w <= lg; w++) {
    start[w] = at;
    at += cnt[w] << (w - 1);
  }
  for (int n = 0; n < nsym; n++) {
    const int w = weights[n];
    if (w == 0) continue;
    const uint32_t len = 1u << (w - 1);
    const uint16_t e = (uint16_t)(n | ((lg + 1 - w) << 8));
    uint32_t u = start[w];
    start[w] = u + len;
    if (len >= 4 && (u & 1) == 0) {
      const uint32_t ee = (uint32_t)e | ((uint32_t)e << 16);
      uint32_t* p = reinterpret_cast<uint32_t*>(lut + u);
      for (uint32_t q = 0; q < (len >> 1); q++) p[q] = ee;
    } else {
      for (uint32_t q = 0; q < len; q++) lut[u + q] = e;
    }
  }
}

__global__ void __launch_bounds__(32) k_huf_decode_planar(const uint8_t* __restrict__ body, uint64_t body_len,
                                                          const ItemDesc* __restrict__ items, uint64_t nitems,
                                                          uint8_t* planes, uint64_t plane_stride, Ctrl* ctrl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  DecodeSmem& S = *reinterpret_cast<DecodeSmem*>(smem_raw);
  const int lane = threadIdx.x;
  const int slot = lane >> 2;    // item within the warp
  const int stream = lane & 3;   // bitstream within the item
  const uint64_t item = (uint64_t)blockIdx.x * kDecItemsPerWarp + slot;
  const uint8_t* lo = body;
  const uint8_t* hi = body + body_len;
  const uint8_t* lo_aligned = reinterpret_cast<const uint8_t*>(((uintptr_t)lo + 3) & ~(uintptr_t)3);

  ItemDesc d;
  d.kind = kRaw;
  d.src_off = 0;
  d.src_len = 0;
  d.dec_len = 0;
  if (item < nitems) d = items[item];
  const bool active = (d.kind == kHuf);
  if (__ballot_sync(0xffffffffu, active) == 0) return;

  // ---- table description -> weights -> LUT (one lane per item) ----
  int lg = 0, hsize = -1;
  if (active && stream == 0) {
    int nsym = 0;
    FseDec& D = *reinterpret_cast<FseDec*>(&S.lut[slot][0]);
    hsize = huf_read_weights(S.weights[slot], &nsym, &lg, body + d.src_off, d.src_len, D);
    if (hsize >= 0 && lg > kDecLutLog) {
      atomicOr(&ctrl->error, kErrUnsupported);
      hsize = -1;
    } else if (hsize < 0) {
      atomicOr(&ctrl->error, kErrCorrupt);
    }
    if (hsize >= 0) fill_lut(S.lut[slot], S.weights[slot], nsym, lg);
  }
  __syncwarp();
  lg = __shfl_sync(0xffffffffu, lg, lane & ~3);
  hsize = __shfl_sync(0xffffffffu, hsize, lane & ~3);
  if (!active || hsize < 0) return;

  // ---- jump table (huf_decompress.c:283-290) ----
  const uint8_t* p = body + d.src_off + hsize;
  const uint32_t rest = d.src_len - (uint32_t)hsize;
  if (rest < 10) {
    atomicOr(&ctrl->error, kErrCorrupt);
    return;
  }
  const uint32_t l0 = p[0] | (p[1] << 8), l1 = p[2] | (p[3] << 8), l2 = p[4] | (p[5] << 8);
  if (l0 + l1 + l2 + 6 > rest) {
    atomicOr(&ctrl->error, kErrCorrupt);
    return;
  }
  const uint32_t l3 = rest - (l0 + l1 + l2 + 6);
  const uint32_t seg = (d.dec_len + 3) >> 2;
  if (3 * seg > d.dec_len || l0 == 0 || l1 == 0 || l2 == 0 || l3 == 0) {
    atomicOr(&ctrl->error, kErrCorrupt);
    return;
  }
  uint32_t s_off = 6, s_len = l0;
  if (stream == 1) { s_off += l0; s_len = l1; }
  if (stream == 2) { s_off += l0 + l1; s_len = l2; }
  if (stream == 3) { s_off += l0 + l1 + l2; s_len = l3; }
  const uint32_t out_off = (uint32_t)stream * seg;
  const uint32_t count = (stream == 3) ? d.dec_len - 3 * seg : seg;

  BitWindow b;
  bool ok = window_init(b, p + s_off, s_len, lo, hi);
  if (ok) ok = decode_stream_planar(b, S.lut[slot], lg, planes + item * plane_stride + out_off, count, lo_aligned);
  if (!ok) atomicOr(&ctrl->error, kErrCorrupt);
}

// ====================================================================================
// Kernel 3: regroup byte planes into the element stream (+ un-rotate the sign bit).
// Sources per (group, chunk): raw bytes inside the stream (unaligned), one RLE byte, or a
// decoded plane in the workspace.  Each thread produces 16 output bytes per step.
// ====================================================================================
struct PlaneSrc {
  const uint8_t* ptr;  // first plane byte (any alignment); nullptr => constant fill
  uint32_t fill;       // the RLE byte replicated into 4 lanes
  uint32_t len;        // plane bytes
};

// n <= 16 consecutive bytes of a plane starting at byte index j (n multiple of 4).
template <int NW>
__device__ __forceinline__ void load_plane_words(const PlaneSrc& s, uint32_t j, uint32_t (&out)[NW]) {
  if (s.ptr == nullptr) {
#pragma unroll
    for (int i = 0; i < NW; i++) out[i] = s.fill;
    return;
  }
  const uintptr_t a = (uintptr_t)(s.ptr + j);
  const uint32_t* base = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t w[NW + 1];
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = __ldg(base + i);
  w[NW] = sh ? __ldg(base + NW) : 0u;  // only touched when the run really spills into it
#pragma unroll
  for (int i = 0; i < NW; i++) out[i] = __funnelshift_r(w[i], w[i + 1], sh);
}

__device__ __forceinline__ uint8_t plane_byte(const PlaneSrc& s, uint32_t j) {
  return s.ptr ? s.ptr[j] : (uint8_t)s.fill;
}

constexpr int kMergeThreads = 256;
constexpr uint32_t kMergeTile = kMergeThreads * 16 * 4;  // bytes of output per block step (16 KiB)

template <int G>
__global__ void __launch_bounds__(kMergeThreads) k_regroup(const uint8_t* __restrict__ body,
                                                           const ItemDesc* __restrict__ items, uint64_t K,
                                                           const uint8_t* __restrict__ planes, uint64_t plane_stride,
                                                           uint32_t chunk, uint64_t orig, int bits_mode,
                                                           uint8_t* __restrict__ out) {
  __shared__ PlaneSrc src[G];
  const uint32_t tiles_per_chunk = (chunk + kMergeTile - 1) / kMergeTile;
  const uint64_t ntiles = K * tiles_per_chunk;
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint64_t c = t / tiles_per_chunk;
    const uint32_t tile = (uint32_t)(t - c * tiles_per_chunk);
    const uint32_t chunk_len = (c == K - 1) ? (uint32_t)(orig - c * (uint64_t)chunk) : chunk;
    const uint32_t o_begin = tile * kMergeTile;
    if (o_begin >= chunk_len) continue;
    __syncthreads();
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      const ItemDesc d = items[(uint64_t)g * K + c];
      PlaneSrc s;
      s.len = d.dec_len;
      s.fill = 0;
      if (d.kind == kRaw) {
        s.ptr = body + d.src_off;
      } else if (d.kind == kRle) {
        s.ptr = nullptr;
        s.fill = 0x01010101u * (uint32_t)body[d.src_off];
      } else {
        s.ptr = planes + ((uint64_t)g * K + c) * plane_stride;
      }
      src[g] = s;
    }
    __syncthreads();
    uint8_t* out_c = out + c * (uint64_t)chunk;
    const uint32_t o_end = min(chunk_len, o_begin + kMergeTile);
    const uint32_t rot_words = (bits_mode == 1 && G > 1) ? (chunk_len >> 2) : 0;  // words that get un-rotated
    for (uint32_t o = o_begin + threadIdx.x * 16; o < o_end; o += kMergeThreads * 16) {
      if (o + 16 <= o_end) {
        uint32_t r[4];
        if (G == 1) {
          load_plane_words<4>(src[0], o, r);
        } else if (G == 2) {
          uint32_t a[2], b2[2];
          load_plane_words<2>(src[0], o >> 1, a);
          load_plane_words<2>(src[1], o >> 1, b2);
          r[0] = __byte_perm(a[0], b2[0], 0x5140);
          r[1] = __byte_perm(a[0], b2[0], 0x7362);
          r[2] = __byte_perm(a[1], b2[1], 0x5140);
          r[3] = __byte_perm(a[1], b2[1], 0x7362);
        } else {
          uint32_t p0[1], p1[1], p2[1], p3[1];
          load_plane_words<1>(src[0], o >> 2, p0);
          load_plane_words<1>(src[1 % G], o >> 2, p1);
          load_plane_words<1>(src[2 % G], o >> 2, p2);
          load_plane_words<1>(src[3 % G], o >> 2, p3);
          const uint32_t t0 = __byte_perm(p0[0], p1[0], 0x5140), t1 = __byte_perm(p2[0], p3[0], 0x5140);
          const uint32_t t2 = __byte_perm(p0[0], p1[0], 0x7362), t3 = __byte_perm(p2[0], p3[0], 0x7362);
          r[0] = __byte_perm(t0, t1, 0x5410);
          r[1] = __byte_perm(t0, t1, 0x7632);
          r[2] = __byte_perm(t2, t3, 0x5410);
          r[3] = __byte_perm(t2, t3, 0x7632);
        }
        const uint32_t w0 = o >> 2;
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (w0 + i < rot_words) r[i] = unrot_word<G>(r[i]);
        *reinterpret_cast<uint4*>(out_c + o) = make_uint4(r[0], r[1], r[2], r[3]);
      } else {
        // ragged tail of the last chunk: byte by byte, whole words still get un-rotated
        for (uint32_t q = o; q < o_end; q += 4) {
          uint32_t w = 0;
          const uint32_t nb = min(4u, o_end - q);
          for (uint32_t i = 0; i < nb; i++) {
            const uint32_t pos = q + i;
            w |= (uint32_t)plane_byte(src[pos % G], pos / G) << (8 * i);
          }
          if ((q >> 2) < rot_words) w = unrot_word<G>(w);
          for (uint32_t i = 0; i < nb; i++) out_c[q + i] = (uint8_t)(w >> (8 * i));
        }
      }
    }
  }
}

}  // namespace zb
