// cvt_math.cuh -- per-pixel conversion arithmetic shared by cvt.cu and the fused strip ops.
#pragma once

#include <stdint.h>

namespace rcv {

// OpenCV 4.13 BGR2GRAY (15-bit coefficients), bit-exact with cv2.cvtColor.
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) {
  return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// ---- BT.601 for the vector kernel ----------------------------------------------------------
// The reference formula (videoio/mod.rs:352-369) per channel is clamp((298*c + k1*u + k2*v + 128) >> 8)
// with c = y-16, u = U-128, v = V-128.  Folding the offsets into one constant per channel gives
//   B = 298*y + 516*U - 70688      G = 298*y - 100*U - 208*V + 34784      R = 298*y + 409*V - 56992
// (all exact in i32): 2 IMAD for the luma terms, 4 for the chroma terms shared by both pixels,
// 6 adds.  clamp(x >> 8) equals byte 1 of min(max(x, 0), 65535) (x < 0 -> 0; x >= 65536 -> 0xFFFF
// -> 255; else floor(x / 256)), so shift + clamp is ONE VIMNMX.RELU (__vimin_s32_relu) and the
// result bytes are gathered with PRMT (dp2a / cvt.pack are emulated on sm_100a -- measured slower).
struct Px6 {
  uint32_t b0, g0, r0, b1, g1, r1;  // clamped to [0, 65535]; the channel value is byte 1
};

template <bool UYVY>
__device__ __forceinline__ Px6 yuv_word(uint32_t w) {
  const int y0 = (int)__byte_perm(w, 0, UYVY ? 0x4441 : 0x4440);
  const int u = (int)__byte_perm(w, 0, UYVY ? 0x4440 : 0x4441);
  const int y1 = (int)__byte_perm(w, 0, UYVY ? 0x4443 : 0x4442);
  const int v = (int)__byte_perm(w, 0, UYVY ? 0x4442 : 0x4443);
  const int cy0 = y0 * 298, cy1 = y1 * 298;
  const int db = u * 516 - 70688;
  const int dr = v * 409 - 56992;
  const int dg = u * -100 + (v * -208 + 34784);
  Px6 o;
  o.b0 = (uint32_t)__vimin_s32_relu(cy0 + db, 65535);
  o.g0 = (uint32_t)__vimin_s32_relu(cy0 + dg, 65535);
  o.r0 = (uint32_t)__vimin_s32_relu(cy0 + dr, 65535);
  o.b1 = (uint32_t)__vimin_s32_relu(cy1 + db, 65535);
  o.g1 = (uint32_t)__vimin_s32_relu(cy1 + dg, 65535);
  o.r1 = (uint32_t)__vimin_s32_relu(cy1 + dr, 65535);
  return o;
}

// The same conversion handing out each channel of the macro-pixel's two pixels as ONE register of two 16-bit
// lanes (value of pixel 0, value of pixel 1), values 0..255.  x >> 8 lies in [-277, 534] for every (Y, U, V), so
// the two shifted sums fit signed 16-bit lanes: one PRMT takes bytes 1-2 of both sums (= x >> 8 truncated to 16
// bits) and one VIMNMX.S16x2.RELU clamps both to [0, 255] -- 2 instructions per channel instead of 2 clamps + a
// pack.  Exactly clamp((...) >> 8) of the reference formula (floor shift first, then clamp).
__device__ __forceinline__ uint32_t mad298(int y, int d) {  // 298 * y + d as ONE multiply-add (NVVM would share 298 * y and add three times)
  int r;
  asm("mad.lo.s32 %0, %1, 298, %2;" : "=r"(r) : "r"(y), "r"(d));
  return (uint32_t)r;
}

template <bool UYVY>
__device__ __forceinline__ void yuv_word_pairs(uint32_t w, uint32_t &pb, uint32_t &pg, uint32_t &pr) {
  const int y0 = (int)__byte_perm(w, 0, UYVY ? 0x4441 : 0x4440);
  const int u = (int)__byte_perm(w, 0, UYVY ? 0x4440 : 0x4441);
  const int y1 = (int)__byte_perm(w, 0, UYVY ? 0x4443 : 0x4442);
  const int v = (int)__byte_perm(w, 0, UYVY ? 0x4442 : 0x4443);
  const int db = u * 516 - 70688;
  const int dr = v * 409 - 56992;
  const int dg = u * -100 + (v * -208 + 34784);
  pb = __vimin_s16x2_relu(__byte_perm(mad298(y0, db), mad298(y1, db), 0x6521), 0x00FF00FFu);
  pg = __vimin_s16x2_relu(__byte_perm(mad298(y0, dg), mad298(y1, dg), 0x6521), 0x00FF00FFu);
  pr = __vimin_s16x2_relu(__byte_perm(mad298(y0, dr), mad298(y1, dr), 0x6521), 0x00FF00FFu);
}

// BGR2GRAY of two pixels at once, on the (pixel 0, pixel 1) 16-bit-lane pairs yuv_word_pairs hands out.  The 15-bit
// coefficients split into bytes -- 3735 = 14*256 + 151, 19235 = 75*256 + 35, 9798 = 38*256 + 70, high parts summing
// to 127 and low parts to 256 -- so A = 14b + 75g + 38r <= 32385 and B = 151b + 35g + 70r <= 65280 both fit a lane, and
//   (3735b + 19235g + 9798r + 16384) >> 15  ==  (256A + B + 16384) >> 15  ==  (A + (B >> 8) + 64) >> 7
// exactly (16384 = 64 * 256, and dropping B's low byte cannot carry across a multiple of 128 of an integer sum).
// 9 operations per pixel pair instead of 2 x 7.  Result: (gray 0, gray 1) as 16-bit lanes, values 0..255.
__device__ __forceinline__ uint32_t gray_pair(uint32_t pb, uint32_t pg, uint32_t pr) {
  const uint32_t A = pb * 14u + pg * 75u + pr * 38u;
  const uint32_t B = pb * 151u + pg * 35u + pr * 70u;
  const uint32_t u = A + __byte_perm(B, 0, 0x4341) + 0x00400040u;
  return (u >> 7) & 0x00FF00FFu;
}

}  // namespace rcv
