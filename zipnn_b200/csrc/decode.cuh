// decode.cuh -- decompress side:  stream metadata -> item table, Huffman bit-unpack fused
// with byte-group regroup (+ sign-bit un-rotate).
//
// Replaces reference csrc/zipnn_core.c:881-1142 (py_combine_dtype), :768-861
// (decompression_chunk_worker), huf_decompress.c:118-354 (table + 4-stream decode),
// data_manipulation_dtype16.c:167-216 and data_manipulation_dtype32.c:391-456 (combine).
//
// Kernels
//   k_decode_meta        one thread per chunk: item table, per-chunk mode, RLE fill blocks
//   k_huf_decode_fused   chunks with exactly one Huffman-coded group (the normal case:
//                        the exponent plane) -> decoded, merged with the raw/RLE planes,
//                        un-rotated and written as elements, in one pass
//   k_huf_decode_planar  chunks with several Huffman-coded groups, or a ragged last chunk:
//                        decode each coded plane into a workspace plane ...
//   k_regroup            ... and regroup planes (raw / RLE / workspace) into elements; also
//                        handles chunks with no coded group at all
#pragma once
#include "common.cuh"

namespace zb {

enum : uint32_t { kModePlain = 0, kModeFused = 1, kModeGeneral = 2, kModeSkip = 3 };
constexpr uint32_t kFillBytes = 64;  // replicated RLE byte block per item (read with stride 0)

struct DecodeCfg {
  const uint8_t* body;
  uint64_t body_len;
  int G;
  uint64_t K;
  uint32_t chunk;
  uint64_t orig;
  int bits_mode;
  Ctrl* ctrl;
  ItemDesc* items;      // [G*K]
  uint8_t* mode;        // [K]
  uint32_t* slot;       // [K] workspace plane slot of a general-mode chunk (G planes per slot)
  uint32_t* rlist;      // [K] chunks that k_regroup has to write (plain and general mode), ctrl->regroup_count of them
  uint8_t* fill;        // [G*K*kFillBytes]
  uint8_t* planes;      // [slots][G][pstride]
  uint64_t pstride;
  uint32_t max_slots;
  uint32_t tail_cap;    // entries in the per-warp tail pool of k_huf_decode_fused (multiple of 8)
};

// ====================================================================================
// Kernel 1: parse + validate the per-(group,chunk) metadata.
// Stream body layout (csrc/zipnn_core.c:105-244):
//   types u8[G][K] | cum u64le[G][K] (inclusive, per group) | group-major payload
// ====================================================================================
__global__ void k_decode_meta(DecodeCfg cfg) {
  const int G = cfg.G;
  const uint64_t K = cfg.K;
  const uint64_t nitems = (uint64_t)G * K;
  const uint8_t* types = cfg.body;
  const uint8_t* cum = cfg.body + nitems;
  const uint64_t payload0 = 9 * nitems;
  const uint64_t payload_len = cfg.body_len - payload0;
  uint64_t base[4] = {0, 0, 0, 0};
  for (int g = 1; g < G; g++) base[g] = base[g - 1] + ld_u64_bytes(cum + 8 * ((uint64_t)(g - 1) * K + (K - 1)));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int g = 0; g < 4; g++) cfg.ctrl->base[g] = payload0 + base[g];
    const uint64_t all = base[G - 1] + ld_u64_bytes(cum + 8 * ((uint64_t)(G - 1) * K + (K - 1)));
    if (all > payload_len) atomicOr(&cfg.ctrl->error, kErrCorrupt);
  }
  for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < K; c += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t chunk_len = (c == K - 1) ? (uint32_t)(cfg.orig - c * (uint64_t)cfg.chunk) : cfg.chunk;
    int nhuf = 0, last_huf = -1;
    bool bad_chunk = false;
    for (int g = 0; g < G; g++) {
      const uint64_t i = (uint64_t)g * K + c;
      const uint64_t hi = ld_u64_bytes(cum + 8 * i);
      const uint64_t lo = c ? ld_u64_bytes(cum + 8 * (i - 1)) : 0;
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
          d.kind = kRaw;
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
        atomicOr(&cfg.ctrl->error, kErrCorrupt);
        d.kind = kRaw;
        d.src_len = 0;
        d.dec_len = 0;
        bad_chunk = true;
      }
      if (d.kind == kHuf) {
        nhuf++;
        last_huf = g;
      }
      if (d.kind == kRle) {
        const uint32_t v = 0x01010101u * (uint32_t)cfg.body[d.src_off];
        uint4* f = reinterpret_cast<uint4*>(cfg.fill + i * kFillBytes);
#pragma unroll
        for (int q = 0; q < (int)(kFillBytes / 16); q++) f[q] = make_uint4(v, v, v, v);
      }
      cfg.items[i] = d;
    }
    // The fused kernel wants the one coded plane to be the top byte plane (the exponent side: what
    // float tensors produce), and whole, equally long planes whose quarter streams are a whole
    // number of 128-byte output rows: chunk_len a multiple of 512.  Anything else goes the general way.
    uint32_t m = kModePlain;
    if (bad_chunk) {
      m = kModeSkip;  // rejected: nothing may be read through its (untrusted) offsets
    } else if (nhuf == 1 && last_huf == G - 1 && (chunk_len % 512u) == 0) {
      m = kModeFused;
    } else if (nhuf >= 1) {
      m = kModeGeneral;
      const uint32_t s = atomicAdd(&cfg.ctrl->work_counter, 1u);
      if (s >= cfg.max_slots) {
        atomicOr(&cfg.ctrl->error, kErrWorkspace);
        m = kModeSkip;  // the caller retries with the full workspace
      } else {
        cfg.slot[c] = s;
      }
    }
    cfg.mode[c] = (uint8_t)m;
    if (m == kModePlain || m == kModeGeneral) cfg.rlist[atomicAdd(&cfg.ctrl->regroup_count, 1u)] = (uint32_t)c;
  }
}

// ====================================================================================
// Shared pieces of the two Huffman kernels.
//
// One thread per bitstream (a huff0 block is 4 independent backward bitstreams,
// huf_decompress.c:283-298); one warp = 8 blocks.  Each block's single-symbol decode
// table (2^tableLog x {symbol, length}) lives in shared memory.
//
// Stream bytes reach the thread through a private 128-byte ring in shared memory that is
// filled with cp.async (16-byte, L2-only) two iterations ahead of use.  Registers never
// wait on a global load: a per-lane "prefetch into a register" does not work on a GPU,
// because lanes refill at different symbols while the scoreboard is per warp, so every
// refill ends up waiting for some other lane's load (measured: 234 clk/symbol).
// The refill itself is branch-free: every second symbol all lanes execute the same
// select/shift/LDS sequence, whether or not their window needed a word.
// ====================================================================================
constexpr int kDecItemsPerWarp = 8;
constexpr int kDecLutLog = 11;  // the reference encoder never exceeds 11 (HUF_TABLELOG_DEFAULT)
constexpr int kDecLutEntries = 1 << kDecLutLog;
constexpr uint32_t kRingBytes = 64;

struct DecodeSmem {
  uint16_t lut[kDecItemsPerWarp][kDecLutEntries];  // also scratch for the table parse
  __align__(16) uint8_t ring[32][kRingBytes];      // per lane; weights[8][256] alias it during the parse
};
static_assert(sizeof(FseDec) <= sizeof(uint16_t) * kDecLutEntries, "FseDec must fit in one LUT slot");
static_assert(32 * kRingBytes >= kDecItemsPerWarp * 256, "weights alias the ring");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// Bit window: a 64-bit container `cont` holding bytes [q, q+8) of the stream (q 4-aligned in the
// ring's offset space) and a count `c` of bits already consumed from its top, as in
// bitstream.h:272-443, but refilled one aligned 32-bit word at a time: when 32 or more bits are
// gone, cont = cont << 32 | next, where `next` (the word below q) was read from the ring at the
// PREVIOUS refill.  A symbol costs peek = cont >> (53 - c) and c += length; the refill check runs
// every 2 symbols (31 + 2 x 11 <= 64 - 11).  The shared-memory read is off the dependent chain
// (its result is needed one refill later) and happens once per 32 stream bits instead of three
// times per 4 symbols.
struct BitWindow {
  uint64_t cont;         // bytes [q, q+8) of the stream, little endian
  int32_t s;             // 53 - (bits consumed from the top of `cont`): (uint32_t)(cont >> s) has the next 11 bits in [10:0]
  uint32_t q;            // byte offset (from gbase) of the container's lowest byte; moves down by 4
  uint32_t next;         // the word at q - 4
  uint32_t rd;           // shared-space address of the word at q - 8 (read by the next refill)
  uint32_t ring_s;       // shared-space address of the ring (64-byte aligned)
  uint32_t fetch;        // byte offset (from gbase) of the lowest 16-byte block already requested
  uint32_t start_bit;    // bit offset (from gbase) of the first stream bit (exact-consumption check)
  const uint8_t* gbase;  // 128-byte aligned global address the offsets are relative to
  uint32_t floor_off;    // do not request blocks below this offset (start of the stream buffer)
  const uint8_t* ring;
};

__device__ __forceinline__ uint32_t ring_word(const uint8_t* ring, uint32_t off) {
  return *reinterpret_cast<const uint32_t*>(ring + (off & (kRingBytes - 4)));
}

// Request every 16-byte block that fits in the ring below what is still needed (<= `maxn`).
__device__ __forceinline__ void ring_top_up(BitWindow& b, int maxn) {
#pragma unroll 2
  for (int i = 0; i < maxn; i++) {
    const uint32_t f = b.fetch - 16;
    // block [f, f+16) replaces ring bytes [f+64, f+80): free once they lie at or above q
    // (the container and `next` are in registers; later reads are at q - 8 and below)
    if (b.fetch >= 16 + b.floor_off && f + kRingBytes >= b.q) {
      cp_async16(const_cast<uint8_t*>(b.ring) + (f & (kRingBytes - 1)), b.gbase + f);
      b.fetch = f;
    }
  }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void window_refill(BitWindow& b) {
  if (b.s <= 21) {  // 32 or more bits consumed
    b.cont = (b.cont << 32) | b.next;
    b.s += 32;
    b.q -= 4u;
    b.next = lds_u32(b.rd);
    b.rd = b.ring_s | ((b.rd - 4u) & (kRingBytes - 4u));
  }
}

// s points at the stream (len bytes); `lo` is the first readable byte of the buffer.
__device__ __forceinline__ bool window_init(BitWindow& b, const uint8_t* s, uint32_t len, const uint8_t* lo, uint8_t* ring) {
  const uint8_t lastb = s[len - 1];
  if (lastb == 0) return false;
  b.ring = ring;
  b.gbase = reinterpret_cast<const uint8_t*>(((uintptr_t)s & ~(uintptr_t)(kRingBytes - 1)) - kRingBytes);
  // (one ring below the stream start keeps every offset the decoder forms non-negative)
  b.floor_off = (b.gbase < lo) ? (uint32_t)(((uintptr_t)lo - (uintptr_t)b.gbase + 15) & ~(uintptr_t)15) : 0u;
  const uint32_t s_off = (uint32_t)((uintptr_t)s - (uintptr_t)b.gbase);
  const uint32_t mark = 8u * (s_off + len - 1) + (uint32_t)hb32(lastb);  // bit offset of the end mark
  b.start_bit = 8u * s_off;
  if (mark == b.start_bit) return false;
  const uint32_t top_byte = (mark - 1) >> 3;
  b.q = (top_byte & ~3u) - 4u;
  b.s = 53 - (int32_t)(8u * (b.q + 8u) - mark);  // 1..32 bits lie above the first unread bit
  b.fetch = (top_byte & ~15u) + 16;
  ring_top_up(b, (int)(kRingBytes / 16));
  cp_async_commit();
  cp_async_wait<0>();
  b.cont = ((uint64_t)ring_word(ring, b.q + 4u) << 32) | ring_word(ring, b.q);
  b.next = ring_word(ring, b.q - 4u);
  b.ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  b.rd = b.ring_s | ((b.q - 8u) & (kRingBytes - 4u));
  return true;
}

__device__ __forceinline__ bool window_exact(const BitWindow& b) {
  return 8u * (b.q + 8u) - (uint32_t)(53 - b.s) == b.start_bit;  // every bit down to the stream start consumed, none below
}

// ---- decode tables ------------------------------------------------------------------
// Full table: 2^lg entries {symbol, length} (what huf_decompress.c:151-183 builds).
struct LutFull {
  const uint16_t* lut;
  int lg;
  __device__ __forceinline__ int32_t get(uint32_t x) const {
    return (int32_t)reinterpret_cast<const int16_t*>(lut)[(x & 0x7FFu) >> (11 - lg)];
  }
};
// Two-level table in an 11-bit index space (shorter table logs are replicated into it):
// codes of <= 8 bits resolve in a 256-entry primary indexed by the top 8 bits; longer codes
// sit at the bottom of the canonical order (index < x_long) and resolve in a tail table
// indexed by all 11 bits (read only by the lanes that need it).
// 1 KiB per block instead of 4 KiB: three times as many bitstreams resident per SM.
__device__ __forceinline__ int32_t lds_s16(uint32_t saddr) {
  int32_t v;  // sign-extended: byte 1 of an entry is minus the code length
  asm("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}
struct LutTwo {
  uint32_t prim_s;  // shared-space byte address of the 256-entry primary
  uint32_t tail_s;  // ... of the x_long-entry tail (32-bit shared addresses: a generic pointer makes
                    // the compiler rebuild the shared window base for every lookup)
  uint32_t x_long;
  __device__ __forceinline__ int32_t get(uint32_t x) const {
    const uint32_t idx = x & 0x7FFu;
    int32_t e = lds_s16(((x >> 2) & 0x1FEu) | prim_s);  // 2 * (top 8 of the 11 bits); every primary is 512-byte aligned
    if (idx < x_long) e = lds_s16(tail_s + idx + idx);
    return e;
  }
};

// Private-column table for planes with short codes (the exponent plane of bf16 / fp32: 98 % of the
// symbols have codes of <= 5 bits).  The shared 256-entry primary above is read at random words by
// 32 lanes: ~3.5 bank conflicts per lookup, and the LSU pipe becomes the limit of the whole kernel.
// Here every lane owns a column of a [32][32] u16 array indexed by the top 5 window bits (two lanes
// share a 4-byte bank word, which is not a conflict), 2 KiB per warp.  The array is 2 KiB ALIGNED in
// the shared address space, so that (x & 0x7C0) | column address is the entry's address in ONE LOP3:
// row stride 64 bytes = bit 6, the 5 index bits are x[10:6].  Codes longer than 5 bits resolve in
// the shared tail (index < x_long), which few lanes touch.
template <int PB>
struct LutCol {
  static_assert(PB == 5, "the one-instruction address needs 2^(11-PB) = 64 bytes = one row of 32 u16");
  uint32_t col_s;   // shared-space address of this lane's entry 0 (2 KiB aligned array + 2 * lane)
  uint32_t tail_s;
  uint32_t x_long;
  __device__ __forceinline__ int32_t get(uint32_t x) const {
    const uint32_t idx = x & 0x7FFu;
    int32_t e = lds_s16((x & 0x7C0u) | col_s);
    if (idx < x_long) e = lds_s16(tail_s + idx + idx);
    return e;
  }
};

// Tail size for a PB-bit primary: index bound of the codes longer than PB bits (or -1).
__device__ __forceinline__ int lut_tail_size(const uint8_t* weights, int nsym, int lg, int pb) {
  if (lg > kDecLutLog) return -1;
  uint32_t cnt[kHufLogMax + 2];
#pragma unroll
  for (int i = 0; i < kHufLogMax + 2; i++) cnt[i] = 0;
  for (int n = 0; n < nsym; n++) cnt[weights[n]]++;
  uint32_t at = 0, x_long = 0;
  for (int w = 1; w <= lg; w++) {
    at += (cnt[w] << (w - 1)) << (kDecLutLog - lg);
    if (lg + 1 - w > pb) x_long = at;
  }
  return (int)x_long;
}

// Fill one lane's column (all 4 lanes of a chunk run it) and, when `with_tail`, the chunk's tail.
template <int PB>
__device__ __forceinline__ void fill_lut_col(uint16_t* col /* entry k at col[32 * k] */, uint16_t* tail, bool with_tail,
                                             const uint8_t* weights, int nsym, int lg) {
  uint32_t cnt[kHufLogMax + 2];
#pragma unroll
  for (int i = 0; i < kHufLogMax + 2; i++) cnt[i] = 0;
  for (int n = 0; n < nsym; n++) cnt[weights[n]]++;
  const int up = kDecLutLog - lg;
  uint32_t start[kHufLogMax + 2];
  uint32_t at = 0;
  start[0] = 0;
  for (int w = 1; w <= lg; w++) {
    start[w] = at;
    at += (cnt[w] << (w - 1)) << up;
  }
  for (int n = 0; n < nsym; n++) {
    const int w = weights[n];
    if (w == 0) continue;
    const int len = lg + 1 - w;
    const uint32_t span = 1u << (kDecLutLog - len);
    const uint32_t e = (uint32_t)n | (((256u - (uint32_t)len) & 0xFFu) << 8);  // symbol | -length
    const uint32_t u = start[w];
    start[w] = u + span;
    if (len > PB) {
      if (with_tail)
        for (uint32_t q = 0; q < span; q++) tail[u + q] = (uint16_t)e;
    } else {
      const uint32_t p0 = u >> (kDecLutLog - PB), pn = span >> (kDecLutLog - PB);
      for (uint32_t q = 0; q < pn; q++) col[32 * (p0 + q)] = (uint16_t)e;
    }
  }
}

// Table entries are 16 bits: symbol in byte 0, MINUS the code length in byte 1 (two's complement), read
// with a sign-extending load.  With s = 53 - consumed the dependent chain per symbol is
//   SHF.R.U64 (cont >> s)  ->  LOP3 (table address)  ->  LDS.S16  ->  LEA.HI.SX32 (s += e >> 8)
// three ALU operations and the load; the straightforward (cont << c) >> 32, index, address, c += len
// takes four.
template <class LUT>
__device__ __forceinline__ uint32_t window_decode(BitWindow& b, const LUT& lut) {
  const uint32_t x = (uint32_t)(b.cont >> b.s);  // next 11 stream bits in [10:0], later bits below... above them: older bits
  const int32_t e = lut.get(x);
  b.s += e >> 8;  // s -= length
  return (uint32_t)e;  // symbol in byte 0
}

// 16 symbols -> 4 words (symbol j in byte j).  Ring maintenance for the NEXT iterations is
// issued first so the copies overlap the decode.
template <class LUT>
__device__ __forceinline__ void decode16(BitWindow& b, const LUT& lut, uint32_t (&o)[4]) {
  // The ring is 64 bytes: one block is requested per 8 symbols (<= 11 bytes consumed), and a block
  // is first read at least one half-iteration after the wait that covers it (see ring_top_up).
#pragma unroll
  for (int h = 0; h < 2; h++) {
    ring_top_up(b, 1);
    cp_async_commit();
#pragma unroll
    for (int q = 2 * h; q < 2 * h + 2; q++) {
      window_refill(b);
      const uint32_t e0 = window_decode(b, lut), e1 = window_decode(b, lut);
      window_refill(b);
      const uint32_t e2 = window_decode(b, lut), e3 = window_decode(b, lut);
      o[q] = __byte_perm(__byte_perm(e0, e1, 0x0040), __byte_perm(e2, e3, 0x0040), 0x5410);
    }
    cp_async_wait<1>();  // everything but the group just committed has landed
  }
}

template <class LUT>
__device__ __forceinline__ uint32_t decode1(BitWindow& b, const LUT& lut) {
  ring_top_up(b, 1);
  cp_async_commit();
  window_refill(b);
  const uint32_t s = window_decode(b, lut) & 0xFFu;
  cp_async_wait<0>();
  return s;
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
    const uint16_t e = (uint16_t)(n | (((256 - (lg + 1 - w)) & 0xFF) << 8));  // symbol | -length
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

// Two-level table, step 1: the tail size (index bound of the codes longer than 8 bits) in the
// 11-bit index space, or -1 when the table log exceeds 11 (the caller demotes the chunk).
__device__ __forceinline__ int lut2_tail_size(const uint8_t* weights, int nsym, int lg) {
  if (lg > kDecLutLog) return -1;
  uint32_t cnt[kHufLogMax + 2];
#pragma unroll
  for (int i = 0; i < kHufLogMax + 2; i++) cnt[i] = 0;
  for (int n = 0; n < nsym; n++) cnt[weights[n]]++;
  uint32_t at = 0, x_long = 0;
  for (int w = 1; w <= lg; w++) {
    at += (cnt[w] << (w - 1)) << (kDecLutLog - lg);
    if (lg + 1 - w > 8) x_long = at;
  }
  return (int)x_long;
}

// Step 2: fill the 256-entry primary and the x_long-entry tail.
__device__ __forceinline__ void fill_lut2(uint16_t* prim, uint16_t* tail, const uint8_t* weights, int nsym, int lg) {
  uint32_t cnt[kHufLogMax + 2];
#pragma unroll
  for (int i = 0; i < kHufLogMax + 2; i++) cnt[i] = 0;
  for (int n = 0; n < nsym; n++) cnt[weights[n]]++;
  const int up = kDecLutLog - lg;  // replicate into the 11-bit index space
  uint32_t start[kHufLogMax + 2];
  uint32_t at = 0;
  start[0] = 0;
  for (int w = 1; w <= lg; w++) {
    start[w] = at;
    at += (cnt[w] << (w - 1)) << up;
  }
  for (int n = 0; n < nsym; n++) {
    const int w = weights[n];
    if (w == 0) continue;
    const int len = lg + 1 - w;
    const uint32_t span = 1u << (kDecLutLog - len);
    const uint16_t e = (uint16_t)(n | (((256 - len) & 0xFF) << 8));  // symbol | -length
    const uint32_t u = start[w];
    start[w] = u + span;
    if (len > 8) {
      for (uint32_t q = 0; q < span; q++) tail[u + q] = e;
    } else {
      const uint32_t p0 = u >> 3, pn = span >> 3;
      for (uint32_t q = 0; q < pn; q++) prim[p0 + q] = e;
    }
  }
}

struct StreamSetup {
  int lg;
  uint32_t s_off, s_len;   // stream position inside the item (after the table description)
  uint32_t out_off, count; // first symbol index and symbol count of this stream
  const uint8_t* p;        // item payload after the table description
};

// Table description -> LUT (lane 0 of each item), then the jump table.  All 32 lanes call it;
// returns false for lanes that have nothing to decode.
__device__ __forceinline__ bool setup_item(DecodeSmem& S, const uint8_t* body, const ItemDesc& d, bool active, int slot,
                                           int stream, Ctrl* ctrl, StreamSetup& st) {
  const int lane = threadIdx.x;
  int lg = 0, hsize = -1;
  uint8_t* weights = &S.ring[0][0] + slot * 256;
  if (active && stream == 0) {
    int nsym = 0;
    FseDec& D = *reinterpret_cast<FseDec*>(&S.lut[slot][0]);
    hsize = huf_read_weights(weights, &nsym, &lg, body + d.src_off, d.src_len, D);
    if (hsize >= 0 && lg > kDecLutLog) {
      atomicOr(&ctrl->error, kErrUnsupported);
      hsize = -1;
    } else if (hsize < 0) {
      atomicOr(&ctrl->error, kErrCorrupt);
    }
    if (hsize >= 0) fill_lut(S.lut[slot], weights, nsym, lg);
  }
  __syncwarp();
  lg = __shfl_sync(0xffffffffu, lg, lane & ~3);
  hsize = __shfl_sync(0xffffffffu, hsize, lane & ~3);
  __syncwarp();  // the ring (aliased by weights) is free from here on
  if (!active || hsize < 0) return false;
  const uint8_t* p = body + d.src_off + hsize;
  const uint32_t rest = d.src_len - (uint32_t)hsize;
  if (rest < 10) {
    atomicOr(&ctrl->error, kErrCorrupt);
    return false;
  }
  const uint32_t l0 = p[0] | (p[1] << 8), l1 = p[2] | (p[3] << 8), l2 = p[4] | (p[5] << 8);
  if (l0 + l1 + l2 + 6 > rest) {
    atomicOr(&ctrl->error, kErrCorrupt);
    return false;
  }
  const uint32_t l3 = rest - (l0 + l1 + l2 + 6);
  const uint32_t seg = (d.dec_len + 3) >> 2;
  if (3 * seg > d.dec_len || l0 == 0 || l1 == 0 || l2 == 0 || l3 == 0) {
    atomicOr(&ctrl->error, kErrCorrupt);
    return false;
  }
  uint32_t s_off = 6, s_len = l0;
  if (stream == 1) { s_off += l0; s_len = l1; }
  if (stream == 2) { s_off += l0 + l1; s_len = l2; }
  if (stream == 3) { s_off += l0 + l1 + l2; s_len = l3; }
  st.lg = lg;
  st.p = p;
  st.s_off = s_off;
  st.s_len = s_len;
  st.out_off = (uint32_t)stream * seg;
  st.count = (stream == 3) ? d.dec_len - 3 * seg : seg;
  return true;
}

// ====================================================================================
// Kernel 2a: general mode -- decode coded planes into workspace planes.
// ====================================================================================
__global__ void __launch_bounds__(32) k_huf_decode_planar(DecodeCfg cfg) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  DecodeSmem& S = *reinterpret_cast<DecodeSmem*>(smem_raw);
  const int lane = threadIdx.x, slot = lane >> 2, stream = lane & 3;
  const uint64_t nitems = (uint64_t)cfg.G * cfg.K;
  const uint64_t item = (uint64_t)blockIdx.x * kDecItemsPerWarp + slot;
  ItemDesc d;
  d.kind = kRaw;
  d.src_off = 0;
  d.src_len = d.dec_len = 0;
  bool active = false;
  uint64_t c = 0;
  int g = 0;
  if (item < nitems) {
    g = (int)(item / cfg.K);
    c = item - (uint64_t)g * cfg.K;
    if (cfg.mode[c] == kModeGeneral) {
      d = cfg.items[item];
      active = (d.kind == kHuf);
    }
  }
  if (__ballot_sync(0xffffffffu, active) == 0) return;
  StreamSetup st;
  if (!setup_item(S, cfg.body, d, active, slot, stream, cfg.ctrl, st)) return;
  uint8_t* dst = cfg.planes + ((uint64_t)cfg.slot[c] * cfg.G + g) * cfg.pstride + st.out_off;
  BitWindow b;
  const LutFull lut{S.lut[slot], st.lg};
  bool ok = window_init(b, st.p + st.s_off, st.s_len, cfg.body, S.ring[lane]);
  if (ok) {
    uint32_t done = 0;
    if ((((uintptr_t)dst) & 15) == 0) {
      const uint32_t n16 = st.count >> 4;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
      for (uint32_t it = 0; it < n16; it++) {
        uint32_t o[4];
        decode16(b, lut, o);
        d4[it] = make_uint4(o[0], o[1], o[2], o[3]);
      }
      done = n16 << 4;
    }
    for (; done < st.count; done++) dst[done] = (uint8_t)decode1(b, lut);
    ok = window_exact(b);
  }
  if (!ok) atomicOr(&cfg.ctrl->error, kErrCorrupt);
}

// ====================================================================================
// Kernel 2b: fused mode.  The lane that decodes 16 symbols of the coded plane also fetches
// the 16 matching bytes of each other plane (raw bytes in the stream at any alignment, or
// a replicated RLE block read with stride 0), interleaves, un-rotates and stores 16*G bytes
// of elements.  The other planes are loaded as aligned 16-byte blocks one iteration ahead.
// Shared memory per warp: 8 x (256-entry primary + 256-entry tail) + 32 rings = 12 KiB.
// ====================================================================================
// Shared memory of one warp (dynamic, tail_cap is a launch parameter):
//   prim  [8][256] u16   4 KiB   shared primary tables (PB = 0), or 3 KiB holding [32][32] u16 private columns (PB = 5)
//   tail  [tail_cap] u16         tail tables of the 8 chunks packed back to back (bf16 / fp32
//                                exponent planes need ~16 entries each, fp16 ~90, fp8 ~150)
//   ring  [32][64]       2 KiB   per-lane stream ring; weights[8][256] alias it during the parse
//   stage [32][128]      4 KiB   one 128-byte output row per lane, 16-byte units XOR-swizzled;
//                                the tANS scratch of the parse aliases it (512 B per chunk)
//   side  [G-1][32][64]  2 KiB per side plane: blocks of the other planes in flight (cp.async)
// bf16: 3 KiB table area (2 KiB of private 5-bit u16 columns) + tail_cap 1024 + one side plane = 13 KiB
// -> 16 warps per SM.
struct FusedSmem {
  uint16_t (*prim)[256];
  uint16_t* tail;
  uint8_t (*ring)[kRingBytes];
  uint8_t (*stage)[128];
  uint8_t* side;   // [G-1][32][64]: four 16-byte slots per lane and side plane for the blocks in flight
};
// PB = 0: shared 256-entry u16 primaries (4 KiB).  PB = 5: private u16 columns, 2 KiB, which must start
// on a 2 KiB boundary of the shared address space (LutCol); the dynamic buffer is 1 KiB aligned, so the
// area is 3 KiB and the columns start 0 or 1 KiB into it.
__host__ __device__ constexpr size_t fused_prim_bytes(int pb) { return pb == 0 ? (size_t)kDecItemsPerWarp * 512 : (size_t)3072; }
__host__ __device__ inline size_t fused_smem_bytes(uint32_t tail_cap, int pb, int G) {
  return fused_prim_bytes(pb) + (size_t)tail_cap * 2 + 32 * kRingBytes + 32 * 128 + (size_t)(G - 1) * 32 * 64;
}
__device__ __forceinline__ FusedSmem fused_smem_carve(unsigned char* raw, uint32_t tail_cap, int pb) {
  FusedSmem S;
  const size_t pbytes = fused_prim_bytes(pb);
  S.prim = reinterpret_cast<uint16_t (*)[256]>(raw);
  S.tail = reinterpret_cast<uint16_t*>(raw + pbytes);
  S.ring = reinterpret_cast<uint8_t (*)[kRingBytes]>(raw + pbytes + (size_t)tail_cap * 2);
  S.stage = reinterpret_cast<uint8_t (*)[128]>(raw + pbytes + (size_t)tail_cap * 2 + 32 * kRingBytes);
  S.side = raw + pbytes + (size_t)tail_cap * 2 + 32 * kRingBytes + 32 * 128;
  return S;
}
static_assert(sizeof(FseDecSmall) <= 512, "small tANS scratch must fit in 4 stage rows");

struct SidePlane {
  const uint4* blk;  // aligned block holding the plane byte that pairs with the lane's next symbol
  uint32_t shift;    // byte offset (0..15) of that byte inside the block
  uint32_t step;     // 1 for stream bytes, 0 for an RLE fill block
  uint4 a, b;        // blocks k, k+1
  uint32_t slots_s;  // shared address of this lane's four 16-byte slots; block j waits in
  uint32_t swz;      // slot (j & 3) ^ swz -- the XOR spreads the lanes of a quarter warp over all banks
};

// The blocks in flight (k+2, k+3) are NOT held in registers.  A register load has a first use, and ptxas
// schedules the load right in front of it whatever the source order says (even for ld.volatile): with
// a register rotation a = b, b = c, c = d the first use of `d` is that move, at the end of the very
// iteration that requested it -- 12 % of all stall samples sat on that one instruction, and an L2
// prefetch only shortened the wait.  cp.async has no destination register: block k+3 is requested
// at the top of iteration k, joins the commit groups of the stream ring, and an LDS picks it up at
// the end of iteration k+1.  64 bytes of shared memory per lane and plane.  Measured: bf16 10.43 ->
// 9.69 ms (16 GiB), fp32 1285 -> 1456 GB/s (4 GiB).
__device__ __forceinline__ void cp_async16_s(uint32_t saddr, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t side_slot(const SidePlane& sp, uint32_t j) { return sp.slots_s + (((j & 3u) ^ sp.swz) << 4); }

__device__ __forceinline__ uint4 ldg128(const uint4* p) { return __ldg(p); }
// 16 bytes starting `shift` bytes into the 32-byte pair (a, b).
__device__ __forceinline__ void take16(const uint4& a, const uint4& b, uint32_t shift, uint32_t (&out)[4]) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  const uint32_t bs = (shift & 3) * 8;
  uint32_t t[7];
#pragma unroll
  for (int i = 0; i < 7; i++) t[i] = __funnelshift_r(w[i], w[i + 1], bs);
  const bool s4 = shift & 4, s8 = shift & 8;
  uint32_t u[5];
#pragma unroll
  for (int i = 0; i < 5; i++) u[i] = s8 ? t[i + 2] : t[i];
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = s4 ? u[i + 1] : u[i];
}

// Un-rotate at plane level: the two top byte planes hold [sign|mant7] (lo) and [exponent] (hi)
// of every element; the element's real top bytes are hi' = sign|exp>>1, lo' = exp<<7|mant7.
// 4 ops per 4 elements instead of 5 per 32-bit word after interleaving.
__device__ __forceinline__ void unrotate_planes(uint32_t& lo, uint32_t& hi) {
  const uint32_t sm = lo, e = hi;
  hi = (sm & 0x80808080u) | ((e >> 1) & 0x7F7F7F7Fu);
  lo = ((e << 7) & 0x80808080u) | (sm & 0x7F7F7F7Fu);
}

// One iteration: 16 symbols of the coded (top) plane + the matching bytes of the G-1 other
// planes -> 16*G bytes of elements.  guard = clamp the look-ahead block loads to the end of
// the stream buffer (only the last iterations of a stream can reach past it).
// Output rows.  A lane produces 16*G bytes per iteration, 64 KiB away from its neighbours'
// data, so direct stores cost one LSU wavefront per lane.  Instead each lane fills a 128-byte
// row in shared memory (16-byte units XOR-swizzled by the lane so that the 128-bit stores of 8
// lanes cover 32 banks), and every 8/G iterations the warp writes the 32 rows out with 8 stores
// that each cover four whole 128-byte lines.
__device__ __forceinline__ uint4* stage_unit(uint8_t (*stage)[128], int row, int unit) {
  return reinterpret_cast<uint4*>(&stage[row][((unit ^ row) & 7) * 16]);
}

// One iteration: 16 symbols of the coded (top) plane + the matching bytes of the G-1 other
// planes -> 16*G bytes of elements into units [unit0, unit0+G) of the lane's stage row.
// guard = clamp the look-ahead block loads to the end of the stream buffer (only the last
// iterations of a stream can reach past it).
template <int G, class LUT>
__device__ __forceinline__ void fused_iteration(BitWindow& b, const LUT& lut, SidePlane (&side)[(G > 1) ? G - 1 : 1],
                                                const uint4* hi_block, bool guard, bool rot, uint8_t (*stage)[128], int lane, int unit0, uint32_t it) {
  if (G > 1) {
#pragma unroll
    for (int g = 0; g < G - 1; g++) {  // block k+3 of every side plane, used two iterations later
      const uint4* nb = side[g].blk + 3 * side[g].step;
      if (guard && side[g].step && nb > hi_block) nb = hi_block;  // (an RLE fill block lives in the workspace)
      cp_async16_s(side_slot(side[g], it + 3u), nb);  // joins the next commit group of decode16
    }
  }
  uint32_t pl[G][4];
  decode16(b, lut, pl[G - 1]);
  if (G == 1) {
    *stage_unit(stage, lane, unit0) = make_uint4(pl[0][0], pl[0][1], pl[0][2], pl[0][3]);
    return;
  }
#pragma unroll
  for (int g = 0; g < G - 1; g++) take16(side[g].a, side[g].b, side[g].shift, pl[g]);
  if (rot) {
#pragma unroll
    for (int q = 0; q < 4; q++) unrotate_planes(pl[(G - 2) % G][q], pl[G - 1][q]);
  }
  uint32_t w[4 * G];
  if (G == 2) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      w[2 * q] = __byte_perm(pl[0][q], pl[1 % G][q], 0x5140);
      w[2 * q + 1] = __byte_perm(pl[0][q], pl[1 % G][q], 0x7362);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t t0 = __byte_perm(pl[0][q], pl[1 % G][q], 0x5140), t1 = __byte_perm(pl[2 % G][q], pl[3 % G][q], 0x5140);
      const uint32_t t2 = __byte_perm(pl[0][q], pl[1 % G][q], 0x7362), t3 = __byte_perm(pl[2 % G][q], pl[3 % G][q], 0x7362);
      w[4 * q] = __byte_perm(t0, t1, 0x5410);
      w[4 * q + 1] = __byte_perm(t0, t1, 0x7632);
      w[4 * q + 2] = __byte_perm(t2, t3, 0x5410);
      w[4 * q + 3] = __byte_perm(t2, t3, 0x7632);
    }
  }
#pragma unroll
  for (int q = 0; q < G; q++) *stage_unit(stage, lane, unit0 + q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
#pragma unroll
  for (int g = 0; g < G - 1; g++) {
    side[g].a = side[g].b;
    side[g].b = lds_u128(side_slot(side[g], it + 2u));  // requested in the previous iteration, landed since
    side[g].blk += side[g].step;
  }
}

template <int G, int PB>
__global__ void __launch_bounds__(32) k_huf_decode_fused(DecodeCfg cfg, uint8_t* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const FusedSmem S = fused_smem_carve(smem_raw, cfg.tail_cap, PB);
  using LUT = typename std::conditional<PB == 0, LutTwo, LutCol<(PB ? PB : 5)>>::type;
  // private columns: first 2 KiB boundary of the shared address space inside the 3 KiB table area
  const uint32_t col_off = PB ? ((2048u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 2047u)) & 2047u) : 0u;
  const int lane = threadIdx.x, slot = lane >> 2, stream = lane & 3;
  const uint64_t K = cfg.K;
  const uint64_t c = (uint64_t)blockIdx.x * kDecItemsPerWarp + slot;
  const bool active = (c < K) && cfg.mode[c] == kModeFused;
  if (__ballot_sync(0xffffffffu, active) == 0) return;
  if (PB != 0 && col_off > 1024u) {  // the dynamic buffer is not 1 KiB aligned: cannot happen, but never decode wrongly
    if (lane == 0) atomicOr(&cfg.ctrl->error, kErrUnsupported);
    return;
  }

  ItemDesc d;  // the coded plane: always group G-1 in fused mode
  d.kind = kRaw;
  d.src_off = 0;
  d.src_len = d.dec_len = 0;
  if (active) d = cfg.items[(uint64_t)(G - 1) * K + c];

  // ---- table description -> two-level table (lane 0 of each chunk) ----
  int lg = 0, hsize = -1, x_long = 0;
  uint32_t tail_at = 0;
  {
    uint8_t* weights = &S.ring[0][0] + slot * 256;
    const bool builder = active && stream == 0;
    int nsym = 0;
    if (builder) {
      FseDecSmall& D = *reinterpret_cast<FseDecSmall*>(&S.stage[4 * slot][0]);
      hsize = huf_read_weights(weights, &nsym, &lg, cfg.body + d.src_off, d.src_len, D);
      if (hsize >= 0) {
        x_long = PB == 0 ? lut2_tail_size(weights, nsym, lg) : lut_tail_size(weights, nsym, lg, PB);
        if (x_long < 0) hsize = -1;
      }
    }
    // the 8 tails share one pool: exclusive prefix over the chunks of the warp
    {
      const uint32_t mine = (builder && hsize >= 0) ? (uint32_t)x_long : 0u;
      uint32_t run = mine;
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, run, o);
        if (lane >= o) run += v;
      }
      tail_at = run - mine;
      if (builder && hsize >= 0 && run > cfg.tail_cap) hsize = -1;  // does not fit: general path
    }
    if (builder) {
      if (hsize >= 0) {
        if (PB == 0) fill_lut2(S.prim[slot], S.tail + tail_at, weights, nsym, lg);
      } else {
        // Not an error yet: a table that needs the big scratch, a long tail or log 12, or a
        // corrupt one.  Hand the chunk to the general kernels, which decide.
        const uint32_t s = atomicAdd(&cfg.ctrl->work_counter, 1u);
        if (s >= cfg.max_slots) {
          atomicOr(&cfg.ctrl->error, kErrWorkspace);
          cfg.mode[c] = (uint8_t)kModeSkip;
        } else {
          cfg.slot[c] = s;
          cfg.mode[c] = (uint8_t)kModeGeneral;
          cfg.rlist[atomicAdd(&cfg.ctrl->regroup_count, 1u)] = (uint32_t)c;
        }
      }
    }
    __syncwarp();
    lg = __shfl_sync(0xffffffffu, lg, lane & ~3);
    hsize = __shfl_sync(0xffffffffu, hsize, lane & ~3);
    x_long = __shfl_sync(0xffffffffu, x_long, lane & ~3);
    tail_at = __shfl_sync(0xffffffffu, tail_at, lane & ~3);
    if (PB != 0) {  // private columns: the 4 lanes of a chunk fill their own copy in parallel
      nsym = __shfl_sync(0xffffffffu, nsym, lane & ~3);
      if (active && hsize >= 0)
        fill_lut_col<(PB ? PB : 5)>(reinterpret_cast<uint16_t*>(smem_raw + col_off) + lane, S.tail + tail_at, stream == 0, weights, nsym, lg);
    }
    __syncwarp();  // the ring and the stage (aliased by the parse scratch) are free from here on
  }
  // From here on no lane leaves early: the output flush is a warp-wide exchange.
  bool live = active && hsize >= 0;

  // ---- jump table (huf_decompress.c:283-290) ----
  const uint8_t* p = cfg.body + d.src_off + (live ? hsize : 0);
  const uint32_t rest = live ? d.src_len - (uint32_t)hsize : 0;
  uint32_t l0 = 0, l1 = 0, l2 = 0, l3 = 0;
  if (live) {
    bool ok = rest >= 10;
    if (ok) {
      l0 = p[0] | (p[1] << 8);
      l1 = p[2] | (p[3] << 8);
      l2 = p[4] | (p[5] << 8);
      ok = l0 + l1 + l2 + 6 <= rest;
      l3 = rest - (l0 + l1 + l2 + 6);
      ok = ok && l0 && l1 && l2 && l3;
    }
    if (!ok) {
      atomicOr(&cfg.ctrl->error, kErrCorrupt);
      live = false;
    }
  }
  const uint32_t seg = d.dec_len >> 2;  // fused chunks: dec_len % 128 == 0
  uint32_t s_off = 6, s_len = l0;
  if (stream == 1) { s_off += l0; s_len = l1; }
  if (stream == 2) { s_off += l0 + l1; s_len = l2; }
  if (stream == 3) { s_off += l0 + l1 + l2; s_len = l3; }
  const uint32_t out_off = (uint32_t)stream * seg;

  // ---- the other planes: groups 0 .. G-2 ----
  constexpr int NS = (G > 1) ? G - 1 : 1;
  SidePlane side[NS];
  const uint4* hi_block = reinterpret_cast<const uint4*>(((uintptr_t)(cfg.body + cfg.body_len) - 1) & ~(uintptr_t)15);
  if (G > 1 && live) {
#pragma unroll
    for (int g = 0; g < G - 1; g++) {
      const uint64_t i = (uint64_t)g * K + c;
      const ItemDesc t = cfg.items[i];
      const uint8_t* q;
      if (t.kind == kRle) {
        q = cfg.fill + i * kFillBytes;
        side[g].step = 0;
      } else {
        q = cfg.body + t.src_off + out_off;
        side[g].step = 1;
      }
      side[g].shift = (uint32_t)((uintptr_t)q & 15);
      side[g].blk = reinterpret_cast<const uint4*>((uintptr_t)q & ~(uintptr_t)15);
      side[g].a = ldg128(side[g].blk);
      const uint4* nb = side[g].blk + side[g].step;
      side[g].b = ldg128((side[g].step && nb > hi_block) ? hi_block : nb);
      nb += side[g].step;
      // block 2 waits in slot 2 (the window setup below commits and waits for it)
      side[g].swz = ((uint32_t)lane >> 1) & 3u;
      side[g].slots_s = (uint32_t)__cvta_generic_to_shared(S.side) + 64u * (uint32_t)(32 * g + lane);
      cp_async16_s(side_slot(side[g], 2u), (side[g].step && nb > hi_block) ? hi_block : nb);
    }
  }

  const bool rot = (cfg.bits_mode == 1) && (G > 1);
  LUT lut;
  if constexpr (PB == 0) {
    lut.prim_s = (uint32_t)__cvta_generic_to_shared(S.prim[slot]);
  } else {
    lut.col_s = (uint32_t)__cvta_generic_to_shared(smem_raw) + col_off + 2u * (uint32_t)lane;
  }
  lut.tail_s = (uint32_t)__cvta_generic_to_shared(S.tail + tail_at);
  lut.x_long = (uint32_t)x_long;
  BitWindow b;
  if (live && !window_init(b, p + s_off, s_len, cfg.body, S.ring[lane])) {
    atomicOr(&cfg.ctrl->error, kErrCorrupt);
    live = false;
  }

  // ---- rows: 128 bytes of output = 128/G elements = kIters iterations of 16 symbols ----
  constexpr int kIters = 8 / G;
  const uint32_t my_rows = live ? (seg >> 4) / kIters : 0;
  uint32_t max_rows = my_rows;
#pragma unroll
  for (int o = 16; o; o >>= 1) max_rows = max(max_rows, __shfl_xor_sync(0xffffffffu, max_rows, o));
  // the 8 stage rows this lane writes out each round: row r*4 + lane/8, 16-byte unit lane%8
  const uint64_t my_out = (uint64_t)(uintptr_t)(out + c * (uint64_t)cfg.chunk + (uint64_t)out_off * G);
  uint64_t row_out[8];
  uint32_t row_cnt[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int src = r * 4 + (lane >> 3);
    row_out[r] = __shfl_sync(0xffffffffu, my_out, src) + (uint64_t)(lane & 7) * 16;
    row_cnt[r] = __shfl_sync(0xffffffffu, my_rows, src);
  }

  for (uint32_t row = 0; row < max_rows; row++) {
    if (row < my_rows) {
      const bool guard = row + 2 >= my_rows;  // look-ahead of 3 blocks: clamp in the last two rows
      // unrolled: with it = row * kIters + k the slot phases (it + 2) & 3, (it + 3) & 3 fold to constants for
      // the 16-bit types (kIters = 4)
#pragma unroll
      for (int k = 0; k < kIters; k++)
        fused_iteration<G, LUT>(b, lut, side, hi_block, guard, rot, S.stage, lane, k * G, row * (uint32_t)kIters + (uint32_t)k);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int src = r * 4 + (lane >> 3);
      if (row < row_cnt[r]) {
        const uint4 v = *stage_unit(S.stage, src, lane & 7);
        *reinterpret_cast<uint4*>((uintptr_t)(row_out[r] + (uint64_t)row * 128)) = v;
      }
    }
    __syncwarp();
  }
  if (live && !window_exact(b)) atomicOr(&cfg.ctrl->error, kErrCorrupt);
}

// ====================================================================================
// Kernel 3: regroup byte planes into the element stream (+ un-rotate the sign bit), for
// chunks the fused kernel did not take.  Sources per (group, chunk): raw bytes inside the
// stream (unaligned), one RLE byte, or a decoded plane in the workspace.
// ====================================================================================
struct PlaneSrc {
  const uint8_t* ptr;  // first plane byte (any alignment); nullptr => constant fill
  uint32_t fill;       // the RLE byte replicated into 4 lanes
  uint32_t len;        // plane bytes
};

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
__global__ void __launch_bounds__(kMergeThreads) k_regroup(DecodeCfg cfg, uint8_t* __restrict__ out) {
  __shared__ PlaneSrc src[G];
  const uint64_t K = cfg.K;
  const uint32_t chunk = cfg.chunk;
  const uint32_t tiles_per_chunk = (chunk + kMergeTile - 1) / kMergeTile;
  const uint64_t ntiles = (uint64_t)cfg.ctrl->regroup_count * tiles_per_chunk;  // usually none: the fused kernel wrote everything
  for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint64_t c = cfg.rlist[t / tiles_per_chunk];
    const uint32_t tile = (uint32_t)(t % tiles_per_chunk);
    const uint32_t chunk_len = (c == K - 1) ? (uint32_t)(cfg.orig - c * (uint64_t)chunk) : chunk;
    const uint32_t o_begin = tile * kMergeTile;
    if (o_begin >= chunk_len) continue;
    __syncthreads();
    if (threadIdx.x < G) {
      const int g = threadIdx.x;
      const ItemDesc d = cfg.items[(uint64_t)g * K + c];
      PlaneSrc s;
      s.len = d.dec_len;
      s.fill = 0;
      if (d.kind == kRaw) {
        s.ptr = cfg.body + d.src_off;
      } else if (d.kind == kRle) {
        s.ptr = nullptr;
        s.fill = 0x01010101u * (uint32_t)cfg.body[d.src_off];
      } else {
        s.ptr = cfg.planes + ((uint64_t)cfg.slot[c] * G + g) * cfg.pstride;
      }
      src[g] = s;
    }
    __syncthreads();
    uint8_t* out_c = out + c * (uint64_t)chunk;
    const uint32_t o_end = min(chunk_len, o_begin + kMergeTile);
    const uint32_t rot_words = (cfg.bits_mode == 1 && G > 1) ? (chunk_len >> 2) : 0;  // words that get un-rotated
    for (uint32_t o = o_begin + threadIdx.x * 16; o < o_end; o += kMergeThreads * 16) {
      if (o + 16 <= o_end) {
        uint32_t r[4];
        if (G == 1) {
          load_plane_words<4>(src[0], o, r);
        } else if (G == 2) {
          uint32_t a[2], b2[2];
          load_plane_words<2>(src[0], o >> 1, a);
          load_plane_words<2>(src[1 % G], o >> 1, b2);
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
        for (uint32_t q = o; q < min(o + 16, o_end); q += 4) {
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
