//! Raw FFI bindings to librcv_imgproc.so (include/rcv_imgproc.h).
//!
//! SOURCE TEXT ONLY: the build image has no rustc/cargo, so this file has never been
//! compiled.  It mirrors, line for line in style, the reference's own FFI module
//! `rustcv-camera/src/backend/macos/mod.rs:42-80` (`mod sys { #[repr(C)] ...; extern "C" {...} }`).
//! Drop it into `rustcv/src/imgproc/cuda/sys.rs`.
#![allow(non_camel_case_types, dead_code)]

use std::os::raw::{c_char, c_int, c_void};

pub const RCV_OK: c_int = 0;
pub const RCV_ERR_ARG: c_int = -1;
pub const RCV_ERR_SIZE: c_int = -2;
pub const RCV_ERR_DEPTH: c_int = -3;
pub const RCV_ERR_CUDA: c_int = -4;
pub const RCV_ERR_UNSUPPORTED: c_int = -5;
pub const RCV_ERR_NOT_INIT: c_int = -6;
pub const RCV_ERR_NOMEM: c_int = -7;
pub const RCV_ERR_NCCL: c_int = -8;

pub const RCV_U8: u8 = 0;
pub const RCV_F32: u8 = 1;
pub const RCV_HOST: u8 = 0;
pub const RCV_DEVICE: u8 = 1;
pub const RCV_HOST_PINNED: u8 = 2;

pub const RCV_COLOR_YUYV2BGR: i32 = 0;
pub const RCV_COLOR_UYVY2BGR: i32 = 1;
pub const RCV_COLOR_BGRA2BGR: i32 = 2;
pub const RCV_COLOR_RGB2BGR: i32 = 3;
pub const RCV_COLOR_BGR2GRAY: i32 = 4;
pub const RCV_COLOR_BGR2XRGB32: i32 = 5;
pub const RCV_COLOR_YUYV2GRAY: i32 = 6;

/// POD mirror of `rustcv::core::Mat` (rustcv/src/core/mat.rs:6-15) handed across the ABI.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct RcvMat {
    pub data: *mut c_void,
    pub rows: i32,
    pub cols: i32,
    pub step: usize,
    pub channels: u8,
    pub depth: u8,
    pub loc: u8,
    pub reserved: u8,
    pub device: i32,
}

#[link(name = "rcv_imgproc", kind = "dylib")]
extern "C" {
    pub fn rcv_init(device: c_int) -> c_int;
    pub fn rcv_init_multi(ngpus: i32) -> c_int;
    pub fn rcv_shutdown() -> c_int;
    pub fn rcv_last_error() -> *const c_char;
    pub fn rcv_sync(device: c_int) -> c_int;

    pub fn rcv_mat_alloc_device(m: *mut RcvMat, rows: i32, cols: i32, channels: i32, depth: i32, device: i32) -> c_int;
    pub fn rcv_mat_free_device(m: *mut RcvMat) -> c_int;
    pub fn rcv_mat_upload(host: *const RcvMat, dev: *mut RcvMat) -> c_int;
    pub fn rcv_mat_download(dev: *const RcvMat, host: *mut RcvMat) -> c_int;
    pub fn rcv_pinned_alloc(ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn rcv_pinned_alloc_on(device: i32, ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn rcv_pinned_free(ptr: *mut c_void) -> c_int;
    pub fn rcv_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    pub fn rcv_host_unregister(ptr: *mut c_void) -> c_int;

    pub fn rcv_cvt_color(src: *const RcvMat, dst: *mut RcvMat, code: i32) -> c_int;
    pub fn rcv_yuyv_to_bgr(src: *const RcvMat, dst: *mut RcvMat) -> c_int;
    pub fn rcv_yuyv_to_bgr_packed(src: *const u8, src_len: usize, dst: *mut u8, dst_len: usize, width: usize, height: usize) -> c_int;
    pub fn rcv_bgra_to_bgr_packed(src: *const u8, src_len: usize, dst: *mut u8, dst_len: usize, width: usize, height: usize) -> c_int;
    pub fn rcv_nv12_to_bgr(y: *const RcvMat, uv: *const RcvMat, dst: *mut RcvMat) -> c_int;

    pub fn rcv_mjpeg_info(jpeg: *const u8, len: usize, width: *mut i32, height: *mut i32) -> c_int;
    pub fn rcv_mjpeg_to_bgr(jpeg: *const u8, len: usize, dst: *mut RcvMat) -> c_int;

    pub fn rcv_convert_to(src: *const RcvMat, dst: *mut RcvMat, alpha: f64, beta: f64) -> c_int;

    pub fn rcv_gaussian_blur(src: *const RcvMat, dst: *mut RcvMat, kw: i32, kh: i32, sigma_x: f64, sigma_y: f64) -> c_int;
    pub fn rcv_sep_filter2d(src: *const RcvMat, dst: *mut RcvMat, kx: *const f32, kw: i32, ky: *const f32, kh: i32) -> c_int;
    pub fn rcv_sep_filter2d_q8(src: *const RcvMat, dst: *mut RcvMat, kx: *const i32, kw: i32, ky: *const i32, kh: i32) -> c_int;
    pub fn rcv_filter2d(src: *const RcvMat, dst: *mut RcvMat, kernel: *const f32, kw: i32, kh: i32, delta: f32) -> c_int;
    pub fn rcv_sobel_mag(src: *const RcvMat, mag: *mut RcvMat, gx: *mut RcvMat, gy: *mut RcvMat) -> c_int;
    pub fn rcv_resize_bilinear(src: *const RcvMat, dst: *mut RcvMat) -> c_int;
    pub fn rcv_warp_affine(src: *const RcvMat, dst: *mut RcvMat, m: *const f64, inverse_map: i32, border_value: f64) -> c_int;
    pub fn rcv_get_rotation_matrix_2d(cx: f64, cy: f64, angle_deg: f64, scale: f64, m: *mut f64) -> c_int;

    pub fn rcv_yuyv_to_bgr_gaussian5(src_yuyv: *const RcvMat, dst_bgr: *mut RcvMat) -> c_int;
    pub fn rcv_yuyv_to_sobel_mag(src_yuyv: *const RcvMat, mag_f32: *mut RcvMat) -> c_int;
    pub fn rcv_yuyv_to_bgr_gaussian5_batch(srcs_yuyv: *const RcvMat, dsts_bgr: *mut RcvMat, n: i32) -> c_int;
    pub fn rcv_yuyv_to_sobel_mag_batch(srcs_yuyv: *const RcvMat, mags_f32: *mut RcvMat, n: i32) -> c_int;

    pub fn rcv_gaussian_blur_batch(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, kw: i32, kh: i32, sigma_x: f64, sigma_y: f64) -> c_int;
    pub fn rcv_resize_bilinear_batch(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32) -> c_int;
    pub fn rcv_warp_affine_batch(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, m: *const f64, inverse_map: i32, border_value: f64) -> c_int;
    pub fn rcv_sobel_mag_batch(srcs: *const RcvMat, mags: *mut RcvMat, n: i32) -> c_int;
    pub fn rcv_cvt_color_batch(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, code: i32) -> c_int;
    pub fn rcv_sep_filter2d_q8_batch(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, kx: *const i32, kw: i32, ky: *const i32, kh: i32) -> c_int;

    // the same batches sharded over several GPUs from the ONE calling thread (frame j -> GPU j mod ngpus)
    pub fn rcv_gaussian_blur_batch_multi(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, ngpus: i32, kw: i32, kh: i32, sigma_x: f64, sigma_y: f64) -> c_int;
    pub fn rcv_sobel_mag_batch_multi(srcs: *const RcvMat, mags: *mut RcvMat, n: i32, ngpus: i32) -> c_int;
    pub fn rcv_resize_bilinear_batch_multi(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, ngpus: i32) -> c_int;
    pub fn rcv_warp_affine_batch_multi(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, ngpus: i32, m: *const f64, inverse_map: i32, border_value: f64) -> c_int;
    pub fn rcv_cvt_color_batch_multi(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, ngpus: i32, code: i32) -> c_int;
    pub fn rcv_yuyv_to_sobel_mag_batch_multi(srcs_yuyv: *const RcvMat, mags_f32: *mut RcvMat, n: i32, ngpus: i32) -> c_int;
    pub fn rcv_yuyv_to_bgr_gaussian5_batch_multi(srcs_yuyv: *const RcvMat, dsts_bgr: *mut RcvMat, n: i32, ngpus: i32) -> c_int;
    pub fn rcv_filter2d_batch(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, kernel: *const f32, kw: i32, kh: i32, delta: f32) -> c_int;
    pub fn rcv_filter2d_batch_multi(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, ngpus: i32, kernel: *const f32, kw: i32, kh: i32, delta: f32) -> c_int;
    pub fn rcv_sep_filter2d_q8_batch_multi(srcs: *const RcvMat, dsts: *mut RcvMat, n: i32, ngpus: i32, kx: *const i32, kw: i32, ky: *const i32, kh: i32) -> c_int;
    pub fn rcv_set_kernel_broadcast(coeffs: *const f32, count: i32, root_device: i32, ngpus: i32, received: *mut f32) -> c_int;
}
