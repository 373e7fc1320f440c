// Host-compiled unit-test shim for zipnn_b200/csrc/huf_serial.cuh.
// TEST ONLY: lets `pytest -m "not gpu"` fuzz the single-thread device routines
// (tree build, table header encode/decode) against the oracle without a GPU.
// It is not part of the product library and is never loaded by zipnn_b200.
#include "../../zipnn_b200/csrc/huf_serial.cuh"
#include <cstring>

extern "C" {

// counts[256] -> nb[256], val[256], header bytes.  Returns header length or -1.
int emu_table_from_counts(const uint32_t* count, uint32_t n, uint8_t* nb, uint16_t* val, uint8_t* hdr, int* lg_out) {
  static thread_local zb::TreeScratch T;
  int max_sym = 255;
  while (max_sym > 0 && count[max_sym] == 0) max_sym--;
  int last = zb::huf_sort_serial(T, count, max_sym);
  int want = zb::fse_pick_log(zb::kHufLogDefault, n, (uint32_t)max_sym, 1);
  int lg = zb::huf_lengths_from_sorted(T, last, want, nb);
  zb::huf_assign_values(nb, max_sym, lg, val);
  *lg_out = lg;
  int h = zb::huf_write_table(T, nb, max_sym, lg);
  if (h > 0) memcpy(hdr, T.hdr, (size_t)h);
  return h;
}

int emu_read_weights(const uint8_t* src, uint32_t size, uint8_t* weights, int* nsym, int* lg) {
  static thread_local zb::FseDec D;
  return zb::huf_read_weights(weights, nsym, lg, src, size, D);
}

// Host model of k_huf_decode_pair's table + step: decodes ONE backward bitstream of `nsyms`
// symbols (src[0..len), last byte holds the end mark) with the pair table built from `weights`.
// Returns 0 when the stream is consumed exactly, -1 on a table the kernel would demote, -2 otherwise.
int emu_pair_decode(const uint8_t* weights, int nsym, int lg, const uint8_t* src, uint32_t len, uint8_t* out,
                    uint32_t nsyms, uint32_t* steps_out) {
  static thread_local uint16_t prim[256], tail[zb::kPairTailEntries], t1[256];
  uint32_t sl[4];
  const int x_long = zb::huf_fill_pair_table(prim, tail, t1, sl, weights, nsym, lg);
  if (x_long < 0) return -1;
  if (len == 0 || src[len - 1] == 0) return -2;
  // bits of the stream, in consumption order: from the bit below the end mark downwards
  int64_t pos = (int64_t)(len - 1) * 8 + (31 - __builtin_clz((uint32_t)src[len - 1]));  // index of the end mark
  auto peek11 = [&](int64_t at) {  // the 11 bits below position `at` (exclusive), zero padded
    uint32_t v = 0;
    for (int i = 1; i <= zb::kPairIndexBits; i++) {
      const int64_t b = at - i;
      const uint32_t bit = b >= 0 ? (src[b >> 3] >> (b & 7)) & 1u : 0u;
      v = (v << 1) | bit;
    }
    return v;
  };
  uint32_t total = 0, steps = 0;
  while (total < nsyms) {
    const uint32_t idx = peek11(pos);
    uint32_t e = prim[idx >> 3], l = e & 15u;
    if (l == 0) {
      if ((int)idx >= x_long) return -2;  // cannot happen with a complete code
      const uint32_t et = tail[idx];
      l = et >> 8;
      e = (et & 0xFFu) << 4;
    }
    const bool is_pair = (e & 0x8000u) != 0;
    const bool take2 = is_pair && total + 2 <= nsyms;
    if (is_pair) {
      const uint32_t i0 = (e >> 4) & 7u, i1 = (e >> 8) & 7u;
      const uint8_t s0 = (uint8_t)(sl[i0 >> 2] >> (8 * (i0 & 3))), s1 = (uint8_t)(sl[i1 >> 2] >> (8 * (i1 & 3)));
      if (!take2) l = (sl[2] >> (4 * i0)) & 15u;
      out[total++] = s0;
      if (take2) out[total++] = s1;
    } else {
      out[total++] = (uint8_t)(e >> 4);
    }
    pos -= l;
    steps++;
  }
  if (steps_out) *steps_out = steps;
  return pos == 0 ? 0 : -2;
}
}
