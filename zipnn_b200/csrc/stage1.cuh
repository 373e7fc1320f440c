// stage1.cuh -- byte-group split / regroup on flat planar buffers (+ sign-bit rotate).
// Stand-alone form of stage 1 (reference data_manipulation_dtype16.c:64-138,167-216 and
// data_manipulation_dtype32.c:78-133,391-456); the codec kernels fuse the same byte
// shuffles into their own loads and stores.
//
// Memory shape: the element side moves as 128-bit loads/stores of 16 consecutive elements
// per thread; each plane side moves as one 128-bit store (load) per thread, so both
// sides are fully coalesced (a warp covers 512 B of every plane and 512*G B of elements).
#pragma once
#include "common.cuh"

namespace zb {

constexpr int kStage1Threads = 256;

// 16 elements held as 4*G words -> G plane words-of-4 (each plane gets 16 bytes).
template <int G>
__device__ __forceinline__ void split16(const uint32_t* w, uint4* plane_out) {
  if (G == 1) {
    plane_out[0] = make_uint4(w[0], w[1], w[2], w[3]);
  } else if (G == 2) {
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      lo[i] = __byte_perm(w[2 * i], w[2 * i + 1], 0x6420);
      hi[i] = __byte_perm(w[2 * i], w[2 * i + 1], 0x7531);
    }
    plane_out[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    plane_out[1] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  } else {
    uint32_t p[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t x = __byte_perm(w[4 * j], w[4 * j + 1], 0x5140), x2 = __byte_perm(w[4 * j + 2], w[4 * j + 3], 0x5140);
      const uint32_t y = __byte_perm(w[4 * j], w[4 * j + 1], 0x7362), y2 = __byte_perm(w[4 * j + 2], w[4 * j + 3], 0x7362);
      p[0][j] = __byte_perm(x, x2, 0x5410);
      p[1][j] = __byte_perm(x, x2, 0x7632);
      p[2][j] = __byte_perm(y, y2, 0x5410);
      p[3][j] = __byte_perm(y, y2, 0x7632);
    }
#pragma unroll
    for (int g = 0; g < 4; g++) plane_out[g] = make_uint4(p[g][0], p[g][1], p[g][2], p[g][3]);
  }
}

// Inverse: G plane vectors (16 bytes each) -> 4*G element words.
template <int G>
__device__ __forceinline__ void join16(const uint4* pl, uint32_t* w) {
  if (G == 1) {
    w[0] = pl[0].x; w[1] = pl[0].y; w[2] = pl[0].z; w[3] = pl[0].w;
  } else if (G == 2) {
    const uint32_t a[4] = {pl[0].x, pl[0].y, pl[0].z, pl[0].w};
    const uint32_t b[4] = {pl[1].x, pl[1].y, pl[1].z, pl[1].w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      w[2 * i] = __byte_perm(a[i], b[i], 0x5140);
      w[2 * i + 1] = __byte_perm(a[i], b[i], 0x7362);
    }
  } else {
    const uint32_t p0[4] = {pl[0].x, pl[0].y, pl[0].z, pl[0].w};
    const uint32_t p1[4] = {pl[1 % G].x, pl[1 % G].y, pl[1 % G].z, pl[1 % G].w};
    const uint32_t p2[4] = {pl[2 % G].x, pl[2 % G].y, pl[2 % G].z, pl[2 % G].w};
    const uint32_t p3[4] = {pl[3 % G].x, pl[3 % G].y, pl[3 % G].z, pl[3 % G].w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t t0 = __byte_perm(p0[j], p1[j], 0x5140), t1 = __byte_perm(p2[j], p3[j], 0x5140);
      const uint32_t t2 = __byte_perm(p0[j], p1[j], 0x7362), t3 = __byte_perm(p2[j], p3[j], 0x7362);
      w[4 * j] = __byte_perm(t0, t1, 0x5410);
      w[4 * j + 1] = __byte_perm(t0, t1, 0x7632);
      w[4 * j + 2] = __byte_perm(t2, t3, 0x5410);
      w[4 * j + 3] = __byte_perm(t2, t3, 0x7632);
    }
  }
}

template <int G>
__global__ void __launch_bounds__(kStage1Threads) k_split_planar(const uint8_t* __restrict__ in, uint64_t n, int bits_mode,
                                                                 uint8_t* __restrict__ planes, uint64_t stride) {
  const uint64_t unit = 16ull * G;  // bytes of input per thread step
  const uint64_t nunits = n / unit;
  const uint64_t rot_words = (bits_mode == 1 && G > 1) ? (n >> 2) : 0;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < nunits; u += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t w[4 * G];
    const uint4* src = reinterpret_cast<const uint4*>(in + u * unit);
#pragma unroll
    for (int i = 0; i < G; i++) {
      const uint4 v = __ldg(src + i);
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
    if (rot_words) {  // full units are always inside the rotated prefix
#pragma unroll
      for (int i = 0; i < 4 * G; i++) w[i] = rot_word<G>(w[i]);
    }
    uint4 pv[G];
    split16<G>(w, pv);
#pragma unroll
    for (int g = 0; g < G; g++) *reinterpret_cast<uint4*>(planes + (uint64_t)g * stride + u * 16) = pv[g];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (uint64_t q = nunits * unit; q < n; q += 4) {
      const uint32_t nb = (uint32_t)((n - q) < 4 ? (n - q) : 4);
      uint32_t w = 0;
      for (uint32_t i = 0; i < nb; i++) w |= (uint32_t)in[q + i] << (8 * i);
      if ((q >> 2) < rot_words) w = rot_word<G>(w);
      for (uint32_t i = 0; i < nb; i++) {
        const uint64_t pos = q + i;
        planes[(pos % G) * stride + pos / G] = (uint8_t)(w >> (8 * i));
      }
    }
  }
}

template <int G>
__global__ void __launch_bounds__(kStage1Threads) k_regroup_planar(const uint8_t* __restrict__ planes, uint64_t stride, uint64_t n,
                                                                   int bits_mode, uint8_t* __restrict__ out) {
  const uint64_t unit = 16ull * G;
  const uint64_t nunits = n / unit;
  const uint64_t rot_words = (bits_mode == 1 && G > 1) ? (n >> 2) : 0;
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < nunits; u += (uint64_t)gridDim.x * blockDim.x) {
    uint4 pv[G];
#pragma unroll
    for (int g = 0; g < G; g++) pv[g] = __ldg(reinterpret_cast<const uint4*>(planes + (uint64_t)g * stride + u * 16));
    uint32_t w[4 * G];
    join16<G>(pv, w);
    if (rot_words) {
#pragma unroll
      for (int i = 0; i < 4 * G; i++) w[i] = unrot_word<G>(w[i]);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + u * unit);
#pragma unroll
    for (int i = 0; i < G; i++) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (uint64_t q = nunits * unit; q < n; q += 4) {
      const uint32_t nb = (uint32_t)((n - q) < 4 ? (n - q) : 4);
      uint32_t w = 0;
      for (uint32_t i = 0; i < nb; i++) {
        const uint64_t pos = q + i;
        w |= (uint32_t)planes[(pos % G) * stride + pos / G] << (8 * i);
      }
      if ((q >> 2) < rot_words) w = unrot_word<G>(w);
      for (uint32_t i = 0; i < nb; i++) out[q + i] = (uint8_t)(w >> (8 * i));
    }
  }
}

}  // namespace zb
