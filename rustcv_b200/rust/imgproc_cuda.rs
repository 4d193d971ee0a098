//! `rustcv::imgproc` filters / colour conversion / geometry on the B200 backend.
//!
//! SOURCE TEXT ONLY (no rustc in the build image; never compiled).  Free functions over
//! `Mat` in the style of `rectangle(mat: &mut Mat, ..)` (rustcv/src/imgproc/drawing.rs:67),
//! `anyhow::Result` like the rest of the facade (rustcv/src/videoio/mod.rs:5).  `dst` is
//! sized by the callee exactly as `VideoCapture::read` sizes its output
//! (rustcv/src/videoio/mod.rs:192-199).  Add `pub mod cuda;` to rustcv/src/imgproc/mod.rs and
//! re-export: `pub use cuda::*;`.
//!
//! Storage: the reference `Mat` owns a pageable `Vec<u8>` (rustcv/src/core/mat.rs:6-15) that
//! `read()` reuses frame after frame.  `MatStorage` below is the shape the three storage
//! variants of include/rcv_imgproc.h take on the Rust side; every variant releases what it
//! holds in `Drop`, the idiom of rustcv-camera/src/backend/macos/mod.rs:264-272.
use anyhow::{anyhow, Result};
use std::ffi::CStr;
use std::os::raw::c_void;

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

// ---------------------------------------------------------------------------------------------
// storage variants
// ---------------------------------------------------------------------------------------------

/// Page-locked host bytes owned by the library (rcv_pinned_alloc_on: on the NUMA node of the GPU
/// that will DMA them).  Derefs to a byte slice, so CPU code keeps working on it.
pub struct PinnedBuf {
    ptr: *mut u8,
    len: usize,
}
unsafe impl Send for PinnedBuf {}

impl PinnedBuf {
    pub fn new(len: usize, device: i32) -> Result<Self> {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { sys::rcv_pinned_alloc_on(device, &mut p, len) })?;
        Ok(Self { ptr: p as *mut u8, len })
    }
    pub fn as_slice(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl Drop for PinnedBuf {
    fn drop(&mut self) {
        unsafe { sys::rcv_pinned_free(self.ptr as *mut c_void) };
    }
}

/// Device-resident rows (HBM), 256-byte pitch; freed in Drop.
pub struct DeviceBuf {
    pod: sys::RcvMat,
}
unsafe impl Send for DeviceBuf {}

impl DeviceBuf {
    pub fn new(rows: i32, cols: i32, channels: u8, depth: u8, device: i32) -> Result<Self> {
        let mut pod: sys::RcvMat = unsafe { std::mem::zeroed() };
        check(unsafe { sys::rcv_mat_alloc_device(&mut pod, rows, cols, channels as i32, depth as i32, device) })?;
        Ok(Self { pod })
    }
}
impl Drop for DeviceBuf {
    fn drop(&mut self) {
        unsafe { sys::rcv_mat_free_device(&mut self.pod) };
    }
}

/// A `Vec<u8>` page-locked IN PLACE (rcv_host_register): the reference's own buffer, unchanged
/// for every CPU user, DMA-able for the GPU.  Unregistered before the Vec is freed or regrown.
pub struct RegisteredVec {
    data: Vec<u8>,
    registered: bool,
}
impl RegisteredVec {
    pub fn new(data: Vec<u8>) -> Self {
        let mut v = Self { data, registered: false };
        v.register();
        v
    }
    fn register(&mut self) {
        if !self.data.is_empty() {
            // failure is not an error: the Mat then travels through the library's bounce ring
            self.registered = unsafe { sys::rcv_host_register(self.data.as_mut_ptr() as *mut c_void, self.data.len()) } == sys::RCV_OK;
        }
    }
    fn unregister(&mut self) {
        if self.registered {
            unsafe { sys::rcv_host_unregister(self.data.as_mut_ptr() as *mut c_void) };
            self.registered = false;
        }
    }
    /// videoio/mod.rs:192-199 / rustcv-camera/src/mat.rs:65-74: reallocate only on a length change.
    pub fn ensure_len(&mut self, len: usize) {
        if self.data.len() != len {
            self.unregister();
            self.data = vec![0; len];
            self.register();
        }
    }
}
impl Drop for RegisteredVec {
    fn drop(&mut self) {
        self.unregister();
    }
}

/// What `Mat.data` becomes: the existing pageable Vec (unchanged default), the same Vec
/// page-locked in place, library-owned pinned bytes, or device-resident rows.
pub enum MatStorage {
    Host(Vec<u8>),
    Registered(RegisteredVec),
    Pinned(PinnedBuf),
    Device(DeviceBuf),
}

impl MatStorage {
    /// (pointer, loc, device) for the POD handed across the ABI.
    fn raw(&self) -> (*mut c_void, u8, i32) {
        match self {
            MatStorage::Host(v) => (v.as_ptr() as *mut c_void, sys::RCV_HOST, 0),
            MatStorage::Registered(r) => (r.data.as_ptr() as *mut c_void, sys::RCV_HOST, 0),
            MatStorage::Pinned(p) => (p.ptr as *mut c_void, sys::RCV_HOST_PINNED, 0),
            MatStorage::Device(d) => (d.pod.data, sys::RCV_DEVICE, d.pod.device),
        }
    }
    fn raw_mut(&mut self) -> (*mut c_void, u8, i32) {
        match self {
            MatStorage::Host(v) => (v.as_mut_ptr() as *mut c_void, sys::RCV_HOST, 0),
            MatStorage::Registered(r) => (r.data.as_mut_ptr() as *mut c_void, sys::RCV_HOST, 0),
            MatStorage::Pinned(p) => (p.ptr as *mut c_void, sys::RCV_HOST_PINNED, 0),
            MatStorage::Device(d) => (d.pod.data, sys::RCV_DEVICE, d.pod.device),
        }
    }
}

/// A Mat over any storage variant: the reference's public fields (mat.rs:6-15) plus the depth
/// tag f32 images need (mat.rs:53 TODO).
pub struct GpuMat {
    pub storage: MatStorage,
    pub rows: i32,
    pub cols: i32,
    pub step: usize,
    pub channels: u8,
    pub depth: u8,
}

impl GpuMat {
    fn pod(&self) -> sys::RcvMat {
        let (data, loc, device) = self.storage.raw();
        sys::RcvMat { data, rows: self.rows, cols: self.cols, step: self.step, channels: self.channels, depth: self.depth, loc, reserved: 0, device }
    }
    fn pod_mut(&mut self) -> sys::RcvMat {
        let (data, loc, device) = self.storage.raw_mut();
        sys::RcvMat { data, rows: self.rows, cols: self.cols, step: self.step, channels: self.channels, depth: self.depth, loc, reserved: 0, device }
    }
}

pub fn gaussian_blur_gpumat(src: &GpuMat, dst: &mut GpuMat, ksize: (i32, i32), sigma: f64) -> Result<()> {
    let (s, mut d) = (src.pod(), dst.pod_mut());
    check(unsafe { sys::rcv_gaussian_blur(&s, &mut d, ksize.0, ksize.1, sigma, sigma) })
}

// ---------------------------------------------------------------------------------------------
// the reference Mat (pageable Vec<u8>), unchanged API
// ---------------------------------------------------------------------------------------------

/// A source: read-only view of the Mat's bytes.
fn pod(m: &Mat) -> sys::RcvMat {
    sys::RcvMat {
        data: m.data.as_ptr() as *mut _, // never written through: every entry point takes src as `const RcvMat *`
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

/// A destination: the pointer is derived from `&mut`, so the library's writes are writes the
/// borrow checker knows about (writing through a pointer taken from `&Mat` would be UB).
fn pod_mut(m: &mut Mat) -> sys::RcvMat {
    sys::RcvMat {
        data: m.data.as_mut_ptr() as *mut _,
        rows: m.rows,
        cols: m.cols,
        step: m.step,
        channels: m.channels,
        depth: sys::RCV_U8,
        loc: sys::RCV_HOST,
        reserved: 0,
        device: 0,
    }
}

/// videoio/mod.rs:192-199: reallocate only when the byte length changes, then set geometry.
/// A buffer that was page-locked in place is released first and the new one registered.
fn ensure_size(m: &mut Mat, rows: i32, cols: i32, channels: u8, elem: usize) {
    let step = cols as usize * channels as usize * elem;
    let len = rows as usize * step;
    if m.data.len() != len {
        if !m.data.is_empty() {
            unsafe { sys::rcv_host_unregister(m.data.as_mut_ptr() as *mut c_void) }; // no-op if never registered
        }
        m.data = vec![0; len];
    }
    m.rows = rows;
    m.cols = cols;
    m.channels = channels;
    m.step = step;
}

/// Page-locks the Mat's current buffer in place: call once after the first `read()` sized it
/// (the same Vec is then reused for every frame).  `release` MUST run before the Mat is dropped
/// -- in a real integration `Mat` gets `impl Drop { release(self) }`.
pub fn pin_in_place(m: &mut Mat) -> Result<()> {
    if m.data.is_empty() {
        return Ok(());
    }
    check(unsafe { sys::rcv_host_register(m.data.as_mut_ptr() as *mut c_void, m.data.len()) })
}
pub fn release(m: &mut Mat) {
    if !m.data.is_empty() {
        unsafe { sys::rcv_host_unregister(m.data.as_mut_ptr() as *mut c_void) };
    }
}

/// Call once per process (e.g. from `VideoCapture::new`); binds GPU `device` (-1: the GPU named
/// by the environment variable RCV_DEVICE, default 0).
pub fn init(device: i32) -> Result<()> {
    check(unsafe { sys::rcv_init(device) })
}

/// Every GPU of the box (or the first `ngpus`), one library worker thread per GPU.
pub fn init_multi(ngpus: i32) -> Result<()> {
    check(unsafe { sys::rcv_init_multi(ngpus) })
}

pub fn gaussian_blur(src: &Mat, dst: &mut Mat, ksize: (i32, i32), sigma: f64) -> Result<()> {
    ensure_size(dst, src.rows, src.cols, src.channels, 1);
    let (s, mut d) = (pod(src), pod_mut(dst));
    check(unsafe { sys::rcv_gaussian_blur(&s, &mut d, ksize.0, ksize.1, sigma, sigma) })
}

/// A batch of independent frames sharded over `ngpus` GPUs (0 = all initialised) from this one
/// thread: frame j runs on GPU j mod ngpus; returns when every frame is done.
pub fn gaussian_blur_batch(srcs: &[Mat], dsts: &mut [Mat], ksize: (i32, i32), sigma: f64, ngpus: i32) -> Result<()> {
    if srcs.len() != dsts.len() {
        return Err(anyhow!("srcs and dsts differ in length"));
    }
    for (s, d) in srcs.iter().zip(dsts.iter_mut()) {
        ensure_size(d, s.rows, s.cols, s.channels, 1);
    }
    let s: Vec<sys::RcvMat> = srcs.iter().map(pod).collect();
    let mut d: Vec<sys::RcvMat> = dsts.iter_mut().map(pod_mut).collect();
    check(unsafe { sys::rcv_gaussian_blur_batch_multi(s.as_ptr(), d.as_mut_ptr(), s.len() as i32, ngpus, ksize.0, ksize.1, sigma, sigma) })
}

pub fn filter2d(src: &Mat, dst: &mut Mat, kernel: &[f32], ksize: (i32, i32), delta: f32) -> Result<()> {
    if ksize.0 < 1 || ksize.1 < 1 || kernel.len() != (ksize.0 as usize) * (ksize.1 as usize) {
        return Err(anyhow!("kernel length does not match ksize"));
    }
    ensure_size(dst, src.rows, src.cols, src.channels, 1);
    let (s, mut d) = (pod(src), pod_mut(dst));
    check(unsafe { sys::rcv_filter2d(&s, &mut d, kernel.as_ptr(), ksize.0, ksize.1, delta) })
}

pub fn cvt_color(src: &Mat, dst: &mut Mat, code: i32) -> Result<()> {
    let dst_channels = match code {
        sys::RCV_COLOR_BGR2GRAY | sys::RCV_COLOR_YUYV2GRAY => 1,
        sys::RCV_COLOR_BGR2XRGB32 => 4,
        _ => 3,
    };
    ensure_size(dst, src.rows, src.cols, dst_channels, 1);
    let (s, mut d) = (pod(src), pod_mut(dst));
    check(unsafe { sys::rcv_cvt_color(&s, &mut d, code) })
}

pub fn resize(src: &Mat, dst: &mut Mat, dsize: (i32, i32)) -> Result<()> {
    if dsize.0 < 0 || dsize.1 < 0 {
        return Err(anyhow!("negative dsize"));
    }
    ensure_size(dst, dsize.1, dsize.0, src.channels, 1);
    let (s, mut d) = (pod(src), pod_mut(dst));
    check(unsafe { sys::rcv_resize_bilinear(&s, &mut d) })
}

pub fn warp_affine(src: &Mat, dst: &mut Mat, m: &[f64; 6], dsize: (i32, i32)) -> Result<()> {
    if dsize.0 < 0 || dsize.1 < 0 {
        return Err(anyhow!("negative dsize"));
    }
    ensure_size(dst, dsize.1, dsize.0, src.channels, 1);
    let (s, mut d) = (pod(src), pod_mut(dst));
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
    ensure_size(mat, h, w, 3, 1);
    let mut d = pod_mut(mat);
    check(unsafe { sys::rcv_mjpeg_to_bgr(data.as_ptr(), data.len(), &mut d) })
}

/// The raw YUYV frame of `read()` (videoio/mod.rs:201-203) as a 2-channel Mat view: no copy.
/// The slice must hold every byte the C side will read: (height-1)*step + width*2 -- a short
/// slice is an error here, never an out-of-bounds read there (the reference's own loop returns
/// silently on a short source, videoio/mod.rs:346-348).
fn yuyv_pod(data: &[u8], width: i32, height: i32, stride: usize) -> Result<sys::RcvMat> {
    if width < 0 || height < 0 {
        return Err(anyhow!("negative frame size {}x{}", width, height));
    }
    let row = width as usize * 2;
    if stride != 0 && stride < row {
        return Err(anyhow!("stride {} is smaller than a row of {} bytes", stride, row));
    }
    let step = if stride != 0 { stride } else { row };
    let need = if width == 0 || height == 0 { 0 } else { (height as usize - 1) * step + row };
    if data.len() < need {
        return Err(anyhow!("YUYV frame needs {} bytes, the slice has {}", need, data.len()));
    }
    Ok(sys::RcvMat {
        data: data.as_ptr() as *mut _, // source only
        rows: height,
        cols: width,
        step,
        channels: 2,
        depth: sys::RCV_U8,
        loc: sys::RCV_HOST,
        reserved: 0,
        device: 0,
    })
}

/// Fused decode -> process: GaussianBlur5x5(YUYV2BGR(frame)) in one kernel, the BGR intermediate never exists.
pub fn yuyv_to_bgr_gaussian5(data: &[u8], width: i32, height: i32, stride: usize, dst: &mut Mat) -> Result<()> {
    let s = yuyv_pod(data, width, height, stride)?;
    ensure_size(dst, height, width, 3, 1);
    let mut d = pod_mut(dst);
    check(unsafe { sys::rcv_yuyv_to_bgr_gaussian5(&s, &mut d) })
}

/// Fused decode -> process: Sobel magnitude (f32) of the frame's gray image in one kernel.  `mag` holds
/// rows x cols f32 values in `data` (step in bytes), the f32 Mat convention of include/rcv_imgproc.h.
pub fn yuyv_to_sobel_magnitude(data: &[u8], width: i32, height: i32, stride: usize, mag: &mut Mat) -> Result<()> {
    let s = yuyv_pod(data, width, height, stride)?;
    ensure_size(mag, height, width, 1, 4);
    let mut d = pod_mut(mag);
    d.depth = sys::RCV_F32;
    check(unsafe { sys::rcv_yuyv_to_sobel_mag(&s, &mut d) })
}
