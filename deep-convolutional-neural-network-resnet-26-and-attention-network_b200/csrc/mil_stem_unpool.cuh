// Max-pool routing of the stem (gbm/model.py:26,53: MaxPool2d(3, stride 2, pad 1) after conv1 + LeakyReLU) in the
// space-to-depth-by-4 form of mil_stem_tc.cu: arg-max record layout and the un-pool arithmetic shared by the fused
// weight-gradient kernel (mil_stem_wgrad.cu) and the stand-alone un-pool kernel (mil_stem_tc.cu) -- ONE routine, so
// the fused and the un-fused backward produce the same bits.
//
// Arg-max record: uint2 am[pooled chunk pc (3)][gp.PS flat pixels], same flat geometry (lead guard, padded rows) as the
// pooled map; the four 16-bit fields are the channel pairs cp = 4 pc + j, each (code of channel 2cp) | (code of 2cp+1)
// << 8 with code = 0x40 | window position (0..8, ATen scan order).  A code byte moved into the HIGH byte of a 16-bit
// lane is a NORMAL fp16 number (0x4000 + pos * 256), so one packed fp16 compare yields the 0xFFFF / 0 lane mask of
// "this window's maximum sits at position pos" for both channels of a pair.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define MIL_AM_CODE 0x4040u  // OR-ed onto (am0 | am1 << 8)

// 16-bit field j (0..3) of an arg-max record -> (code0 << 8) | (code1 << 24)
__device__ __forceinline__ uint32_t mil_am_lanes(const uint2& a, int j) {
  return __byte_perm((j & 2) ? a.y : a.x, 0u, (j & 1) ? 0x3424 : 0x1404);
}
__device__ __forceinline__ uint32_t mil_am_is(uint32_t lanes, uint32_t pos) {
  const uint32_t want = 0x40004000u | (pos << 8) | (pos << 24);
  return __heq2_mask(*reinterpret_cast<const __half2*>(&lanes), *reinterpret_cast<const __half2*>(&want));
}
__device__ __forceinline__ uint32_t mil_bf2_add(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// One channel pair at phase-map pixel (Y, X): g.. = the pair's pooled-gradient word (bf16x2: channel 2cp | 2cp+1) and
// a.. = its arg-max lanes at the four pooled windows that can point into the pixel -- (Y,X), (Y,X+1), (Y+1,X),
// (Y+1,X+1).  Conv position (2Y+a, 2X+b) seen from window (Y+dyy, X+dxx) is window position
// (a + 1 - 2 dyy) * 3 + (b + 1 - 2 dxx).  Returns the chunk of dY4: channel 2cp phases 00 01 10 11, channel 2cp+1 ...
// Sums of colliding windows are bf16 additions (pairwise for the four-way case).
__device__ __forceinline__ uint4 mil_unpool_pair(uint32_t g00, uint32_t g01, uint32_t g10, uint32_t g11, uint32_t a00,
                                                 uint32_t a01, uint32_t a10, uint32_t a11) {
  const uint32_t r00 = g00 & mil_am_is(a00, 4);
  const uint32_t r01 = mil_bf2_add(g00 & mil_am_is(a00, 5), g01 & mil_am_is(a01, 3));
  const uint32_t r10 = mil_bf2_add(g00 & mil_am_is(a00, 7), g10 & mil_am_is(a10, 1));
  const uint32_t r11 = mil_bf2_add(mil_bf2_add(g00 & mil_am_is(a00, 8), g01 & mil_am_is(a01, 6)),
                                   mil_bf2_add(g10 & mil_am_is(a10, 2), g11 & mil_am_is(a11, 0)));
  uint4 o;
  o.x = __byte_perm(r00, r01, 0x5410);  // channel 2cp:   phases (0,0) (0,1)
  o.y = __byte_perm(r10, r11, 0x5410);  //                phases (1,0) (1,1)
  o.z = __byte_perm(r00, r01, 0x7632);  // channel 2cp+1
  o.w = __byte_perm(r10, r11, 0x7632);
  return o;
}

__device__ __forceinline__ uint32_t mil_word(const uint4& v, int j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}
