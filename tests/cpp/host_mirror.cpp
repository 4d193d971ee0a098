// Exercises the C++ host mirror (rustcv_b200/hostcpp/rustcv_b200.hpp) against the CPU oracle.
// Reads like the reference's own tests (rustcv-camera/src/decode.rs:234-273).
//   host_mirror            full run on cuda:0 (exit 0 = parity)
//   host_mirror --no-gpu   checks the loud failure mode only
#include <cstdio>
#include <cstring>
#include <vector>

#include "rustcv_b200.hpp"
#include "../../oracle/rcv_oracle.h"

using namespace rustcv;

#define EXPECT(cond)                                                  \
  do {                                                                \
    if (!(cond)) {                                                    \
      std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                       \
    }                                                                 \
  } while (0)

int main(int argc, char **argv) {
  if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) {
    core::Mat a = core::Mat::create(8, 8, 3), b;
    Result r = imgproc::gaussian_blur(a, b, {5, 5}, 0.0);
    EXPECT(!r.is_ok() && r.code == RCV_ERR_NOT_INIT);
    EXPECT(r.message.find("no CPU fallback") != std::string::npos);
    std::puts("host_mirror --no-gpu ok");
    return 0;
  }
  EXPECT(init(0).is_ok());

  // decode.rs:235-252 / :255-265 through decode_frame
  {
    const uint8_t white[4] = {235, 128, 235, 128}, black[4] = {16, 128, 16, 128};
    core::Mat m;
    EXPECT(videoio::decode_frame(white, 4, 2, 1, videoio::YUYV, m).is_ok());
    for (uint8_t v : m.data) EXPECT(v > 240);
    EXPECT(videoio::decode_frame(black, 4, 2, 1, videoio::YUYV, m).is_ok());
    for (uint8_t v : m.data) EXPECT(v < 10);
    // the facade silently returns on a short source (videoio/mod.rs:346-348); the ABI reports it
    EXPECT(videoio::decode_frame(white, 3, 2, 1, videoio::YUYV, m).code == RCV_ERR_SIZE);
    // MJPG frames go to nvJPEG: four bytes of YUYV are not a JPEG
    EXPECT(videoio::decode_frame(white, 4, 2, 1, videoio::MJPEG, m).code == RCV_ERR_ARG);
  }
  // config 1: SplitMix64 seed 1, 640x480
  {
    std::vector<uint8_t> src(640 * 480 * 2), want(640 * 480 * 3);
    orc_fill_u8(1, src.data(), src.size());
    core::Mat m;
    EXPECT(videoio::decode_frame(src.data(), src.size(), 640, 480, videoio::YUYV, m).is_ok());
    EXPECT(orc_yuyv_to_bgr_facade(src.data(), src.size(), want.data(), want.size(), 640, 480) == 0);
    EXPECT(m.rows == 480 && m.cols == 640 && m.channels == 3 && m.step == 1920);
    EXPECT(std::memcmp(m.data.data(), want.data(), want.size()) == 0);
    EXPECT(orc_crc32(m.data.data(), m.data.size()) == 0x0BF66518u);
  }
  // read() into a device-resident Mat: raw YUYV up (2 B/px), BGR stays in HBM
  {
    std::vector<uint8_t> src(640 * 480 * 2);
    orc_fill_u8(1, src.data(), src.size());
    core::DeviceMat dm;
    EXPECT(core::DeviceMat::create(dm, 480, 640, 3).is_ok());
    EXPECT(videoio::decode_frame(src.data(), src.size(), 640, 480, videoio::YUYV, dm).is_ok());
    core::Mat back;
    EXPECT(dm.download(back).is_ok());
    EXPECT(orc_crc32(back.data.data(), back.data.size()) == 0x0BF66518u);
    EXPECT(videoio::decode_frame(src.data(), 100, 640, 480, videoio::YUYV, dm).code == RCV_ERR_SIZE);
  }
  // GaussianBlur 5x5 on host Mats and on device-resident Mats
  {
    core::Mat src = core::Mat::create(270, 480, 3), dst, want = core::Mat::create(270, 480, 3);
    orc_fill_u8(2, src.data.data(), src.data.size());
    orc_gaussian_blur_u8(src.data.data(), src.step, want.data.data(), want.step, 270, 480, 3, 5, 5, 0.0, 0.0);
    EXPECT(imgproc::gaussian_blur(src, dst, {5, 5}, 0.0).is_ok());
    EXPECT(dst.data == want.data);
    core::DeviceMat ds, dd;
    EXPECT(core::DeviceMat::create(ds, 270, 480, 3).is_ok() && core::DeviceMat::create(dd, 270, 480, 3).is_ok());
    EXPECT(ds.upload(src).is_ok());
    EXPECT(imgproc::gaussian_blur(ds, dd, {5, 5}, 0.0).is_ok());
    core::Mat back;
    EXPECT(dd.download(back).is_ok());
    EXPECT(back.data == want.data);
    // dst geometry is the callee's job for host Mats, exactly like read() (videoio/mod.rs:192-199)
    EXPECT(dst.rows == 270 && dst.cols == 480 && dst.step == 1440);
  }
  // Sobel + magnitude, resize, BGR->Gray
  {
    core::Mat f = core::Mat::create(135, 240, 1, core::F32), mag, want = core::Mat::create(135, 240, 1, core::F32);
    orc_fill_f32(3, (float *)f.data.data(), 135 * 240);
    orc_sobel3_f32((const float *)f.data.data(), f.step, nullptr, 0, nullptr, 0, (float *)want.data.data(), want.step, 135, 240);
    EXPECT(imgproc::sobel_magnitude(f, mag).is_ok());
    EXPECT(mag.data == want.data);
    core::Mat img = core::Mat::create(64, 96, 3), small, wsmall = core::Mat::create(16, 24, 3), gray, wgray = core::Mat::create(64, 96, 1);
    orc_fill_u8(11, img.data.data(), img.data.size());
    orc_resize_bilinear_u8(img.data.data(), img.step, 64, 96, wsmall.data.data(), wsmall.step, 16, 24, 3);
    EXPECT(imgproc::resize(img, small, {24, 16}).is_ok());
    EXPECT(small.data == wsmall.data);
    orc_bgr_to_gray_strided(img.data.data(), img.step, wgray.data.data(), wgray.step, 64, 96);
    EXPECT(imgproc::cvt_color(img, gray, imgproc::COLOR_BGR2GRAY).is_ok());
    EXPECT(gray.data == wgray.data);
  }
  // fused decode -> process chain: raw YUYV -> BGR -> Gray -> f32 -> Sobel magnitude in one kernel, against the
  // oracle's stand-alone stages run one after the other
  {
    const int h = 120, w = 480;
    core::Mat yuyv = core::Mat::create(h, w, 2), mag;
    orc_fill_u8(21, yuyv.data.data(), yuyv.data.size());
    core::Mat bgr = core::Mat::create(h, w, 3), gray = core::Mat::create(h, w, 1);
    core::Mat gf = core::Mat::create(h, w, 1, core::F32), want = core::Mat::create(h, w, 1, core::F32);
    orc_yuyv_to_bgr_strided(yuyv.data.data(), yuyv.step, bgr.data.data(), bgr.step, h, w);
    orc_bgr_to_gray_strided(bgr.data.data(), bgr.step, gray.data.data(), gray.step, h, w);
    orc_convert_to(gray.data.data(), gray.step, 0, gf.data.data(), gf.step, 1, h, w, 1.0, 0.0);
    orc_sobel3_f32((const float *)gf.data.data(), gf.step, nullptr, 0, nullptr, 0, (float *)want.data.data(), want.step, h, w);
    EXPECT(imgproc::yuyv_to_sobel_magnitude(yuyv, mag).is_ok());
    EXPECT(mag.rows == h && mag.cols == w && mag.channels == 1 && mag.depth == core::F32);
    EXPECT(mag.data == want.data);
  }
  // the reference's reused Vec<u8>, page-locked in place once; then a batch fanned out over every GPU of the box
  // from this one thread (frame j -> GPU j mod N)
  {
    EXPECT(init_multi(0).is_ok());
    const int n = 5, h = 96, w = 160;
    std::vector<core::Mat> srcs(n), dsts, wants(n);
    for (int j = 0; j < n; ++j) {
      srcs[j] = core::Mat::create(h, w, 3);
      wants[j] = core::Mat::create(h, w, 3);
      orc_fill_u8(40 + j, srcs[j].data.data(), srcs[j].data.size());
      orc_gaussian_blur_u8(srcs[j].data.data(), srcs[j].step, wants[j].data.data(), wants[j].step, h, w, 3, 5, 5, 0.0, 0.0);
    }
    EXPECT(core::pin_in_place(srcs[0]).is_ok());
    EXPECT(core::pin_in_place(srcs[0]).is_ok());  // idempotent
    EXPECT(imgproc::gaussian_blur_batch(srcs, dsts, {5, 5}, 0.0, 0.0, 0).is_ok());
    for (int j = 0; j < n; ++j) EXPECT(dsts[j].data == wants[j].data);
    core::Mat one;
    EXPECT(imgproc::gaussian_blur(srcs[0], one, {5, 5}, 0.0).is_ok());  // the registered Mat through the single-frame call
    EXPECT(one.data == wants[0].data);
    core::release(srcs[0]);
    core::release(srcs[0]);
  }
  std::puts("host_mirror ok");
  return 0;
}
