//! `rustcv::imgproc` filters / colour conversion / geometry on the B200 backend.
//!
//! SOURCE TEXT ONLY (no rustc in the build image; never compiled).  Free functions over
//! `Mat` in the style of `rectangle(mat: &mut Mat, ..)` (rustcv/src/imgproc/drawing.rs:67),
//! `anyhow::Result` like the rest of the facade (rustcv/src/videoio/mod.rs:5).  `dst` is
//! sized by the callee exactly as `VideoCapture::read` sizes its output
//! (rustcv/src/videoio/mod.rs:192-199).  Add `pub mod cuda;` to rustcv/src/imgproc/mod.rs and
//! re-export: `pub use cuda::*;`.
use anyhow::{anyhow, Result};
use std::ffi::CStr;

use super::sys;
use crate::core::mat::Mat;

fn check(rc: i32) -> Result<()> {
    if rc == sys::RCV_OK {
        return Ok(());
    }
    // error mapping in the idiom of rustcv-camera/src/backend/macos/mod.rs:145-164
    let msg = unsafe { CStr::from_ptr(sys::rcv_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow!("rcv_imgproc error {}: {}", rc, msg))
}

fn pod(m: &Mat) -> sys::RcvMat {
    sys::RcvMat {
        data: m.data.as_ptr() as *mut _,
        rows: m.rows,
        cols: m.cols,
        step: m.step,
        channels: m.channels,
        depth: sys::RCV_U8, // rustcv::core::Mat is u8-only today (mat.rs:53 TODO)
        loc: sys::RCV_HOST,
        reserved: 0,
        device: 0,
    }
}

/// videoio/mod.rs:192-199: reallocate only when the byte length changes, then set geometry.
fn ensure_size(m: &mut Mat, rows: i32, cols: i32, channels: u8) {
    let step = cols as usize * channels as usize;
    let len = rows as usize * step;
    if m.data.len() != len {
        m.data = vec![0; len];
    }
    m.rows = rows;
    m.cols = cols;
    m.channels = channels;
    m.step = step;
}

/// Call once per process (e.g. from `VideoCapture::new`); binds GPU `device`.
pub fn init(device: i32) -> Result<()> {
    check(unsafe { sys::rcv_init(device) })
}

pub fn gaussian_blur(src: &Mat, dst: &mut Mat, ksize: (i32, i32), sigma: f64) -> Result<()> {
    ensure_size(dst, src.rows, src.cols, src.channels);
    let (s, mut d) = (pod(src), pod(dst));
    check(unsafe { sys::rcv_gaussian_blur(&s, &mut d, ksize.0, ksize.1, sigma, sigma) })
}

pub fn filter2d(src: &Mat, dst: &mut Mat, kernel: &[f32], ksize: (i32, i32), delta: f32) -> Result<()> {
    if kernel.len() != (ksize.0 * ksize.1) as usize {
        return Err(anyhow!("kernel length does not match ksize"));
    }
    ensure_size(dst, src.rows, src.cols, src.channels);
    let (s, mut d) = (pod(src), pod(dst));
    check(unsafe { sys::rcv_filter2d(&s, &mut d, kernel.as_ptr(), ksize.0, ksize.1, delta) })
}

pub fn cvt_color(src: &Mat, dst: &mut Mat, code: i32) -> Result<()> {
    let dst_channels = match code {
        sys::RCV_COLOR_BGR2GRAY | sys::RCV_COLOR_YUYV2GRAY => 1,
        sys::RCV_COLOR_BGR2XRGB32 => 4,
        _ => 3,
    };
    ensure_size(dst, src.rows, src.cols, dst_channels);
    let (s, mut d) = (pod(src), pod(dst));
    check(unsafe { sys::rcv_cvt_color(&s, &mut d, code) })
}

pub fn resize(src: &Mat, dst: &mut Mat, dsize: (i32, i32)) -> Result<()> {
    ensure_size(dst, dsize.1, dsize.0, src.channels);
    let (s, mut d) = (pod(src), pod(dst));
    check(unsafe { sys::rcv_resize_bilinear(&s, &mut d) })
}

pub fn warp_affine(src: &Mat, dst: &mut Mat, m: &[f64; 6], dsize: (i32, i32)) -> Result<()> {
    ensure_size(dst, dsize.1, dsize.0, src.channels);
    let (s, mut d) = (pod(src), pod(dst));
    check(unsafe { sys::rcv_warp_affine(&s, &mut d, m.as_ptr(), 0, 0.0) })
}

/// Replaces the two scalar loops called from `VideoCapture::read`
/// (rustcv/src/videoio/mod.rs:203 and :205) -- same packed-buffer contract.
pub fn yuyv_to_bgr(src: &[u8], dest: &mut [u8], width: usize, height: usize) -> Result<()> {
    check(unsafe { sys::rcv_yuyv_to_bgr_packed(src.as_ptr(), src.len(), dest.as_mut_ptr(), dest.len(), width, height) })
}

pub fn bgra_to_bgr(src: &[u8], dest: &mut [u8], width: usize, height: usize) -> Result<()> {
    check(unsafe { sys::rcv_bgra_to_bgr_packed(src.as_ptr(), src.len(), dest.as_mut_ptr(), dest.len(), width, height) })
}

/// Replaces the TurboJPEG branch of `VideoCapture::read` (rustcv/src/videoio/mod.rs:205-232; `decode_mjpeg`,
/// rustcv-camera/src/decode.rs:93-121): header first, then decompress to BGR at the Mat's pitch.
pub fn mjpeg_to_bgr(data: &[u8], mat: &mut Mat) -> Result<()> {
    let (mut w, mut h) = (0i32, 0i32);
    check(unsafe { sys::rcv_mjpeg_info(data.as_ptr(), data.len(), &mut w, &mut h) })?;
    ensure_size(mat, h, w, 3);
    let mut d = pod(mat);
    check(unsafe { sys::rcv_mjpeg_to_bgr(data.as_ptr(), data.len(), &mut d) })
}

/// The raw YUYV frame of `read()` (videoio/mod.rs:201-203) as a 2-channel Mat view: no copy.
fn yuyv_pod(data: &[u8], width: i32, height: i32, stride: usize) -> sys::RcvMat {
    sys::RcvMat {
        data: data.as_ptr() as *mut _,
        rows: height,
        cols: width,
        step: if stride != 0 { stride } else { width as usize * 2 },
        channels: 2,
        depth: sys::RCV_U8,
        loc: sys::RCV_HOST,
        reserved: 0,
        device: 0,
    }
}

/// Fused decode -> process: GaussianBlur5x5(YUYV2BGR(frame)) in one kernel, the BGR intermediate never exists.
pub fn yuyv_to_bgr_gaussian5(data: &[u8], width: i32, height: i32, stride: usize, dst: &mut Mat) -> Result<()> {
    ensure_size(dst, height, width, 3);
    let (s, mut d) = (yuyv_pod(data, width, height, stride), pod(dst));
    check(unsafe { sys::rcv_yuyv_to_bgr_gaussian5(&s, &mut d) })
}

/// Fused decode -> process: Sobel magnitude (f32) of the frame's gray image in one kernel.  `mag` holds
/// rows x cols f32 values in `data` (step in bytes), the f32 Mat convention of include/rcv_imgproc.h.
pub fn yuyv_to_sobel_magnitude(data: &[u8], width: i32, height: i32, stride: usize, mag: &mut Mat) -> Result<()> {
    let step = width as usize * 4;
    if mag.data.len() != height as usize * step {
        mag.data = vec![0; height as usize * step];
    }
    mag.rows = height;
    mag.cols = width;
    mag.channels = 1;
    mag.step = step;
    let s = yuyv_pod(data, width, height, stride);
    let mut d = pod(mag);
    d.depth = sys::RCV_F32;
    check(unsafe { sys::rcv_yuyv_to_sobel_mag(&s, &mut d) })
}
