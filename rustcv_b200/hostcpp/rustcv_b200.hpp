// rustcv_b200.hpp -- C++ host-side mirror of the reference API for the imgproc hot path,
// over the C ABI of include/rcv_imgproc.h.  Header-only; link with -lrcv_imgproc.
//
// The reference is Rust (no rustc in this image), so per the project rules the host side
// above the C ABI is C++ with the reference's names, argument meaning and error behaviour:
//   rustcv::core::Mat            rustcv/src/core/mat.rs:6-51   (pub fields, new/empty/is_empty/row_bytes)
//   rustcv::imgproc::*           free functions over Mat, OpenCV argument order, in the style of
//                                rectangle(mat, ..)  rustcv/src/imgproc/drawing.rs:67
//   rustcv::videoio::decode_frame  the FourCC dispatch of VideoCapture::read,
//                                rustcv/src/videoio/mod.rs:181-260 / rustcv-camera/src/decode.rs:36-86
// Errors: the reference returns anyhow::Result (videoio/mod.rs:5); here every function
// returns rustcv::Result {code, message} (code 0 = Ok), never throws, never falls back to a CPU path.
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "rcv_imgproc.h"

namespace rustcv {

struct Result {
  int code = RCV_OK;
  std::string message;
  bool is_ok() const { return code == RCV_OK; }
  explicit operator bool() const { return is_ok(); }
};

inline Result check(int rc) {
  Result r;
  r.code = rc;
  if (rc != RCV_OK) r.message = rcv_last_error();
  return r;
}

namespace core {

enum Depth : uint8_t { U8 = RCV_U8, F32 = RCV_F32 };

// rustcv/src/core/mat.rs:6-15 -- owned, strided.  `depth` is the f32 extension (TODO at mat.rs:53).
struct Mat {
  std::vector<uint8_t> data;
  int32_t rows = 0;
  int32_t cols = 0;
  size_t step = 0;  // bytes per row; packed: cols*channels*elemsize, padded: larger
  uint8_t channels = 0;
  uint8_t depth = U8;

  // Mat::new (mat.rs:18-28): packed, zero-filled
  static Mat create(int32_t rows, int32_t cols, uint8_t channels, uint8_t depth = U8) {
    Mat m;
    m.rows = rows;
    m.cols = cols;
    m.channels = channels;
    m.depth = depth;
    m.step = (size_t)cols * channels * (depth == F32 ? 4 : 1);
    m.data.assign((size_t)rows * m.step, 0);
    return m;
  }
  static Mat empty() { return Mat(); }                                               // mat.rs:31-39
  bool is_empty() const { return data.empty() || rows == 0 || cols == 0; }            // mat.rs:42-44
  size_t elem_size() const { return depth == F32 ? 4 : 1; }
  // mat.rs:47-51: the valid bytes of a row (padding dropped)
  std::pair<const uint8_t *, size_t> row_bytes(int32_t row) const {
    return {data.data() + (size_t)row * step, (size_t)cols * channels * elem_size()};
  }
  // what VideoCapture::read does to its output before converting (videoio/mod.rs:192-199)
  // A buffer that was page-locked in place (pin_in_place) is released before it is reallocated.
  void ensure_size(int32_t r, int32_t c, uint8_t cn, uint8_t d = U8) {
    size_t st = (size_t)c * cn * (d == F32 ? 4 : 1);
    if (data.size() != (size_t)r * st) {
      if (!data.empty()) rcv_host_unregister(data.data());  // no-op when never registered
      data.assign((size_t)r * st, 0);
    }
    rows = r;
    cols = c;
    channels = cn;
    depth = d;
    step = st;
  }
  RcvMat pod() const {
    RcvMat m;
    std::memset(&m, 0, sizeof(m));
    m.data = const_cast<uint8_t *>(data.data());
    m.rows = rows;
    m.cols = cols;
    m.step = step;
    m.channels = channels;
    m.depth = depth;
    m.loc = RCV_HOST;
    return m;
  }
};

// The reference reuses one Vec<u8> per Mat frame after frame (videoio/mod.rs:192-199): page-lock it in place once
// and every later call DMAs it directly (rcv_host_register).  `release` must run before the Mat is destroyed or
// its buffer replaced by hand -- the Rust wrapper does it in Drop.
inline Result pin_in_place(Mat &m) {
  if (m.data.empty()) return Result();
  return check(rcv_host_register(m.data.data(), m.data.size()));
}
inline void release(Mat &m) {
  if (!m.data.empty()) rcv_host_unregister(m.data.data());
}

// Device-resident storage variant (HBM).  Move-only RAII, freed like the reference frees its
// native handle in Drop (rustcv-camera/src/backend/macos/mod.rs:264-272).
class DeviceMat {
 public:
  DeviceMat() { std::memset(&m_, 0, sizeof(m_)); }
  DeviceMat(const DeviceMat &) = delete;
  DeviceMat &operator=(const DeviceMat &) = delete;
  DeviceMat(DeviceMat &&o) noexcept : m_(o.m_) { o.m_.data = nullptr; }
  DeviceMat &operator=(DeviceMat &&o) noexcept {
    if (this != &o) {
      release();
      m_ = o.m_;
      o.m_.data = nullptr;
    }
    return *this;
  }
  ~DeviceMat() { release(); }
  static Result create(DeviceMat &out, int32_t rows, int32_t cols, int channels, int depth = U8, int device = -1) {
    out.release();
    return check(rcv_mat_alloc_device(&out.m_, rows, cols, channels, depth, device));
  }
  Result upload(const Mat &host) {
    RcvMat h = host.pod();
    return check(rcv_mat_upload(&h, &m_));
  }
  Result download(Mat &host) const {
    host.ensure_size(m_.rows, m_.cols, m_.channels, m_.depth);
    RcvMat h = host.pod();
    return check(rcv_mat_download(&m_, &h));
  }
  const RcvMat &pod() const { return m_; }
  RcvMat &pod() { return m_; }

 private:
  void release() {
    if (m_.data) rcv_mat_free_device(&m_);
    m_.data = nullptr;
  }
  RcvMat m_;
};

}  // namespace core

inline Result init(int device = 0) { return check(rcv_init(device)); }
// every GPU of the box (or the first `ngpus`), one library worker thread per GPU
inline Result init_multi(int ngpus = 0) { return check(rcv_init_multi(ngpus)); }

namespace imgproc {
using core::Mat;

struct Size {
  int32_t width, height;
};

enum ColorCode {
  COLOR_YUYV2BGR = RCV_COLOR_YUYV2BGR,
  COLOR_UYVY2BGR = RCV_COLOR_UYVY2BGR,
  COLOR_BGRA2BGR = RCV_COLOR_BGRA2BGR,
  COLOR_RGB2BGR = RCV_COLOR_RGB2BGR,
  COLOR_BGR2RGB = RCV_COLOR_BGR2RGB,
  COLOR_BGR2GRAY = RCV_COLOR_BGR2GRAY,
  COLOR_BGR2XRGB32 = RCV_COLOR_BGR2XRGB32,
  COLOR_YUYV2GRAY = RCV_COLOR_YUYV2GRAY,
};

inline Result cvt_color(const Mat &src, Mat &dst, ColorCode code) {
  static const uint8_t dst_cn[] = {3, 3, 3, 3, 1, 4, 1};
  if ((int)code < 0 || (int)code > 6) return check(rcv_cvt_color(nullptr, nullptr, code));
  dst.ensure_size(src.rows, src.cols, dst_cn[code]);
  RcvMat s = src.pod(), d = dst.pod();
  return check(rcv_cvt_color(&s, &d, code));
}

inline Result gaussian_blur(const Mat &src, Mat &dst, Size ksize, double sigma_x, double sigma_y = 0.0) {
  dst.ensure_size(src.rows, src.cols, src.channels, src.depth);
  RcvMat s = src.pod(), d = dst.pod();
  return check(rcv_gaussian_blur(&s, &d, ksize.width, ksize.height, sigma_x, sigma_y));
}
inline Result gaussian_blur(const core::DeviceMat &src, core::DeviceMat &dst, Size ksize, double sigma_x,
                            double sigma_y = 0.0) {
  return check(rcv_gaussian_blur(&src.pod(), &dst.pod(), ksize.width, ksize.height, sigma_x, sigma_y));
}

// A batch of independent frames sharded over `ngpus` GPUs (0 = all initialised) from this ONE calling thread:
// frame j runs on GPU j mod ngpus; returns when every frame is done.
inline Result gaussian_blur_batch(const std::vector<Mat> &srcs, std::vector<Mat> &dsts, Size ksize, double sigma_x,
                                  double sigma_y = 0.0, int ngpus = 0) {
  dsts.resize(srcs.size());
  std::vector<RcvMat> s(srcs.size()), d(srcs.size());
  for (size_t i = 0; i < srcs.size(); ++i) {
    dsts[i].ensure_size(srcs[i].rows, srcs[i].cols, srcs[i].channels, srcs[i].depth);
    s[i] = srcs[i].pod();
    d[i] = dsts[i].pod();
  }
  return check(rcv_gaussian_blur_batch_multi(s.data(), d.data(), (int32_t)s.size(), ngpus, ksize.width, ksize.height, sigma_x,
                                             sigma_y));
}

inline Result filter2d(const Mat &src, Mat &dst, const std::vector<float> &kernel, int kw, int kh, float delta = 0.f) {
  if ((size_t)kw * kh != kernel.size()) return check(rcv_filter2d(nullptr, nullptr, nullptr, 0, 0, 0.f));
  dst.ensure_size(src.rows, src.cols, src.channels, src.depth);
  RcvMat s = src.pod(), d = dst.pod();
  return check(rcv_filter2d(&s, &d, kernel.data(), kw, kh, delta));
}

inline Result sep_filter2d(const Mat &src, Mat &dst, const std::vector<float> &kx, const std::vector<float> &ky) {
  dst.ensure_size(src.rows, src.cols, src.channels, src.depth);
  RcvMat s = src.pod(), d = dst.pod();
  return check(rcv_sep_filter2d(&s, &d, kx.data(), (int)kx.size(), ky.data(), (int)ky.size()));
}

// Sobel 3x3 + gradient magnitude, single-channel f32
inline Result sobel_magnitude(const Mat &src, Mat &mag) {
  mag.ensure_size(src.rows, src.cols, 1, core::F32);
  RcvMat s = src.pod(), d = mag.pod();
  return check(rcv_sobel_mag(&s, &d, nullptr, nullptr));
}

// Fused decode -> process chain: the raw YUYV frame (channels = 2) -> BGR -> Gray -> f32 -> Sobel magnitude in
// one kernel; bit-identical to cvt_color x2 + convert_to + sobel_magnitude.
inline Result yuyv_to_sobel_magnitude(const Mat &src_yuyv, Mat &mag) {
  mag.ensure_size(src_yuyv.rows, src_yuyv.cols, 1, core::F32);
  RcvMat s = src_yuyv.pod(), d = mag.pod();
  return check(rcv_yuyv_to_sobel_mag(&s, &d));
}

inline Result resize(const Mat &src, Mat &dst, Size dsize) {
  dst.ensure_size(dsize.height, dsize.width, src.channels, src.depth);
  RcvMat s = src.pod(), d = dst.pod();
  return check(rcv_resize_bilinear(&s, &d));
}

inline Result get_rotation_matrix_2d(double cx, double cy, double angle_deg, double scale, double M[6]) {
  return check(rcv_get_rotation_matrix_2d(cx, cy, angle_deg, scale, M));
}

inline Result warp_affine(const Mat &src, Mat &dst, const double M[6], Size dsize, double border_value = 0.0) {
  dst.ensure_size(dsize.height, dsize.width, src.channels, src.depth);
  RcvMat s = src.pod(), d = dst.pod();
  return check(rcv_warp_affine(&s, &d, M, 0, border_value));
}

}  // namespace imgproc

namespace videoio {

inline constexpr uint32_t fourcc(char a, char b, char c, char d) {  // rustcv-core/src/pixel_format.rs:6-33
  return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
}
constexpr uint32_t YUYV = fourcc('Y', 'U', 'Y', 'V');
constexpr uint32_t BGRA = fourcc('B', 'G', 'R', 'A');
constexpr uint32_t MJPEG = fourcc('M', 'J', 'P', 'G');

// The conversion step of VideoCapture::read (videoio/mod.rs:181-260): size `mat`, branch on
// FourCC, convert a PACKED frame.  MJPG frames are decoded by nvJPEG (header first, like TurboJPEG's read_header).
inline Result decode_frame(const uint8_t *data, size_t len, uint32_t width, uint32_t height, uint32_t fcc,
                           core::Mat &mat) {
  if (fcc == MJPEG) {  // videoio/mod.rs:205-232: the frame's own header gives the size
    int32_t w = 0, h = 0;
    Result r = check(rcv_mjpeg_info(data, len, &w, &h));
    if (!r.is_ok()) return r;
    mat.ensure_size(h, w, 3);
    RcvMat d = mat.pod();
    return check(rcv_mjpeg_to_bgr(data, len, &d));
  }
  mat.ensure_size((int32_t)height, (int32_t)width, 3);
  if (fcc == YUYV) return check(rcv_yuyv_to_bgr_packed(data, len, mat.data.data(), mat.data.size(), width, height));
  if (fcc == BGRA) return check(rcv_bgra_to_bgr_packed(data, len, mat.data.data(), mat.data.size(), width, height));
  if (len == mat.data.size()) std::memcpy(mat.data.data(), data, len);  // "Assume RGB/BGR or copy" (:253-257)
  return Result();
}

// Device-resident form: only the raw frame crosses PCIe (2 B/px for
// YUYV), the conversion runs on the GPU and the BGR stays in HBM for the imgproc calls that follow.
// `mat` must already be height x width x 3.  stride = bytes between source rows (0 = packed).
inline Result decode_frame(const uint8_t *data, size_t len, uint32_t width, uint32_t height, uint32_t fcc,
                           core::DeviceMat &mat, size_t stride = 0) {
  if (fcc == MJPEG) return check(rcv_mjpeg_to_bgr(data, len, &mat.pod()));  // only the compressed frame crosses PCIe
  const int bpp = fcc == YUYV ? 2 : fcc == BGRA ? 4 : 0;
  Result r;
  if (!bpp) {
    r.code = RCV_ERR_UNSUPPORTED;
    r.message = "format not converted on this path";
    return r;
  }
  if (!stride) stride = (size_t)width * bpp;
  if (height && len < (size_t)(height - 1) * stride + (size_t)width * bpp) {
    r.code = RCV_ERR_SIZE;
    r.message = "frame buffer shorter than height x stride";
    return r;
  }
  RcvMat src;
  std::memset(&src, 0, sizeof(src));
  src.data = const_cast<uint8_t *>(data);
  src.rows = (int32_t)height;
  src.cols = (int32_t)width;
  src.step = stride;
  src.channels = (uint8_t)bpp;
  src.depth = RCV_U8;
  src.loc = RCV_HOST;
  return check(rcv_cvt_color(&src, &mat.pod(), fcc == YUYV ? RCV_COLOR_YUYV2BGR : RCV_COLOR_BGRA2BGR));
}

}  // namespace videoio
}  // namespace rustcv
