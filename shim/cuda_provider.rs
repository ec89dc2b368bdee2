//! shim/cuda_provider.rs — `impl AccelProvider for CudaProvider`, forwarding to the C ABI in include/rm_accel.h.
//!
//! This is the binding a RunMat maintainer adds under the dormant `cuda` feature of `crates/runmat-accelerate`
//! (`Cargo.toml:12`, `DeviceKind::Cuda` at `src/lib.rs:373`). It is SOURCE ONLY in this repository: the image has no
//! `cargo`/`rustc`, so it is not compiled or tested here (INTEGRATION.md says how it slots into the reference).
//! Every method is a 1:1 forward; anything the C library reports as RM_UNSUPPORTED keeps the trait's default
//! ("... not supported by provider"), so callers fall back to host exactly as they do today.
//! `tests/test_shim_coverage.py` parses this file, include/rm_accel.h and (when present) the reference trait, and fails when an
//! exported entry point with a trait counterpart is not forwarded here or a forwarded method's arity differs from the trait's.

use anyhow::{anyhow, Result};
use runmat_accelerate_api::{
    AccelDownloadFuture, AccelProvider, AccelProviderFuture, ApiDeviceInfo, GpuTensorHandle, GpuTensorStorage,
    CovNormalization, CovRows, CovarianceOptions, FindDirection, HostTensorOwned, HostTensorView, ImageNormalizeDescriptor,
    ImfilterMode, ImfilterOptions, ImfilterPadding, ImfilterShape, MatmulEpilogue, PowerStepEpilogue, ProviderConvMode,
    ProviderDispatchStats, ProviderFindResult, ProviderLinsolveOptions, ProviderLinsolveResult, ProviderMoments2, ProviderPrecision,
    KernelAttrTelemetry, KernelLaunchTelemetry, ProviderTelemetry, ReduceDimResult, ReductionFlavor, ScaleOp, SpawnHandleConcurrency,
};
use std::ffi::{c_char, c_int, c_void, CStr, CString};

pub const RM_MAX_RANK: usize = 16;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RmHandle {
    pub buffer_id: u64,
    pub device_id: u32,
    pub rank: u32,
    pub shape: [u64; RM_MAX_RANK],
}

#[repr(C)]
pub struct RmMatmulEpilogue {
    pub alpha: f64,
    pub beta: f64,
    pub row_scale: *const RmHandle,
    pub col_scale: *const RmHandle,
    pub row_op: c_int,
    pub col_op: c_int,
    pub has_clamp_min: c_int,
    pub clamp_min: f64,
    pub has_clamp_max: c_int,
    pub clamp_max: f64,
    pub has_pow: c_int,
    pub pow_exponent: f64,
    pub diag_output: *const RmHandle,
}

#[repr(C)]
pub struct RmImageNormalizeDesc {
    pub batch: u64,
    pub height: u64,
    pub width: u64,
    pub epsilon: f64,
    pub has_gain: c_int,
    pub gain: f64,
    pub has_bias: c_int,
    pub bias: f64,
    pub has_gamma: c_int,
    pub gamma: f64,
    pub clamp_zero: c_int,
}

#[repr(C)]
pub struct RmImfilterOptions {
    pub padding: c_int,        // rm_imfilter_padding: constant, replicate, symmetric, circular
    pub constant_value: f64,
    pub shape: c_int,          // rm_imfilter_shape: same, full, valid
    pub mode: c_int,           // rm_imfilter_mode: correlation, convolution
}

#[repr(C)]
pub struct RmLinsolveOptions {
    pub lower: c_int,
    pub upper: c_int,
    pub rectangular: c_int,
    pub transposed: c_int,
    pub conjugate: c_int,
    pub symmetric: c_int,
    pub posdef: c_int,
    pub need_rcond: c_int,
    pub has_rcond: c_int,
    pub rcond: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct RmDispatchStats {
    pub count: u64,
    pub total_wall_time_ns: u64,
}

#[repr(C)]
#[derive(Default)]
pub struct RmTelemetry {
    pub fused_elementwise: RmDispatchStats,
    pub fused_reduction: RmDispatchStats,
    pub matmul: RmDispatchStats,
    pub linsolve: RmDispatchStats,
    pub mldivide: RmDispatchStats,
    pub mrdivide: RmDispatchStats,
    pub upload_bytes: u64,
    pub download_bytes: u64,
    pub fusion_cache_hits: u64,
    pub fusion_cache_misses: u64,
    pub kernel_launches: u64,
}

/// rm_kernel_launch_event (include/rm_accel.h): one entry of the bounded launch log
#[repr(C)]
#[derive(Clone, Copy)]
pub struct RmKernelAttr { pub key: [u8; 16], pub value: u64 }
#[repr(C)]
#[derive(Clone, Copy)]
pub struct RmKernelLaunchEvent {
    pub kernel: [u8; 32],
    pub precision: c_int,
    pub n_shape: u32,
    pub n_tuning: u32,
    pub shape: [RmKernelAttr; 4],
    pub tuning: [RmKernelAttr; 4],
}
fn c_chars(b: &[u8]) -> String { String::from_utf8_lossy(&b[..b.iter().position(|&c| c == 0).unwrap_or(b.len())]).into_owned() }

#[link(name = "rm_accel_b200")]
extern "C" {
    fn rm_provider_create(cuda_ordinal: c_int, device_id: u32, precision: c_int, out: *mut *mut c_void) -> c_int;
    fn rm_provider_destroy(p: *mut c_void) -> c_int;
    fn rm_last_error() -> *const c_char;
    fn rm_device_info_string(p: *mut c_void, buf: *mut c_char, len: usize) -> c_int;
    fn rm_upload(p: *mut c_void, data: *const f64, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_download(p: *mut c_void, h: *const RmHandle, out: *mut f64, n: u64) -> c_int;
    fn rm_free(p: *mut c_void, h: *const RmHandle) -> c_int;
    fn rm_read_scalar(p: *mut c_void, h: *const RmHandle, idx: u64, out: *mut f64) -> c_int;
    fn rm_zeros(p: *mut c_void, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_fill(p: *mut c_void, shape: *const u64, rank: u32, v: f64, out: *mut RmHandle) -> c_int;
    fn rm_linspace(p: *mut c_void, start: f64, stop: f64, count: u64, out: *mut RmHandle) -> c_int;
    fn rm_transpose(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_gather_linear(p: *mut c_void, src: *const RmHandle, idx: *const u32, n: u64, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_scatter_linear(p: *mut c_void, dst: *const RmHandle, idx: *const u32, n: u64, vals: *const RmHandle) -> c_int;
    fn rm_elem_binary(p: *mut c_void, op: c_int, a: *const RmHandle, b: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_unary(p: *mut c_void, op: c_int, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_scalar_op_apply(p: *mut c_void, op: c_int, a: *const RmHandle, s: f64, out: *mut RmHandle) -> c_int;
    fn rm_fused_elementwise(p: *mut c_void, shader: *const c_char, inputs: *const RmHandle, n_inputs: u32, shape: *const u64, rank: u32, len: u64, out: *mut RmHandle) -> c_int;
    fn rm_fused_elementwise_multi(p: *mut c_void, shader: *const c_char, inputs: *const RmHandle, n_inputs: u32, shape: *const u64, rank: u32, len: u64, n_out: u32, outs: *mut RmHandle) -> c_int;
    fn rm_fused_reduction(p: *mut c_void, shader: *const c_char, inputs: *const RmHandle, n_inputs: u32, shape: *const u64, rank: u32, reduce_len: u64, num_slices: u64, wg: u32, flavor: c_int, custom_scale: f64, out: *mut RmHandle) -> c_int;
    fn rm_reduce_sum(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_reduce_sum_dim(p: *mut c_void, a: *const RmHandle, dim: u32, out: *mut RmHandle) -> c_int;
    fn rm_reduce_mean(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_reduce_max_dim(p: *mut c_void, a: *const RmHandle, dim: u32, vals: *mut RmHandle, idx: *mut RmHandle) -> c_int;
    fn rm_matmul(p: *mut c_void, a: *const RmHandle, b: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_matmul_epilogue_apply(p: *mut c_void, a: *const RmHandle, b: *const RmHandle, ep: *const RmMatmulEpilogue, out: *mut RmHandle) -> c_int;
    fn rm_image_normalize(p: *mut c_void, a: *const RmHandle, d: *const RmImageNormalizeDesc, out: *mut RmHandle) -> c_int;
    fn rm_stochastic_evolution(p: *mut c_void, s: *const RmHandle, drift: f64, scale: f64, steps: u32, out: *mut RmHandle) -> c_int;
    fn rm_set_rng_state(p: *mut c_void, state: u64) -> c_int;
    fn rm_telemetry_snapshot(p: *mut c_void, out: *mut RmTelemetry) -> c_int;
    fn rm_reset_telemetry(p: *mut c_void) -> c_int;
    fn rm_kernel_launch_log(p: *mut c_void, out: *mut RmKernelLaunchEvent, cap: u32, count: *mut u32) -> c_int;
    fn rm_spawn_handle_concurrency_policy(p: *mut c_void) -> c_int;
    fn rm_fused_cache_counters(p: *mut c_void, hits: *mut u64, misses: *mut u64);
    // "next" rows: solves, indexing/layout class, fusion-pattern hooks, filters
    fn rm_mldivide(p: *mut c_void, lhs: *const RmHandle, rhs: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_mrdivide(p: *mut c_void, lhs: *const RmHandle, rhs: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_syrk(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_matmul_power_step(p: *mut c_void, lhs: *const RmHandle, rhs: *const RmHandle, epsilon: f64, out: *mut RmHandle) -> c_int;
    fn rm_covariance(p: *mut c_void, m: *const RmHandle, normalization_biased: c_int, out: *mut RmHandle) -> c_int;
    fn rm_diag_extract(p: *mut c_void, m: *const RmHandle, offset: i64, out: *mut RmHandle) -> c_int;
    fn rm_imfilter(p: *mut c_void, image: *const RmHandle, kernel: *const RmHandle, opt: *const RmImfilterOptions, out: *mut RmHandle) -> c_int;
    fn rm_conv2d(p: *mut c_void, signal: *const RmHandle, kernel: *const RmHandle, mode: c_int, out: *mut RmHandle) -> c_int;
    fn rm_permute(p: *mut c_void, a: *const RmHandle, order: *const u32, n: u32, out: *mut RmHandle) -> c_int;
    fn rm_repmat(p: *mut c_void, a: *const RmHandle, reps: *const u64, n: u32, out: *mut RmHandle) -> c_int;
    fn rm_cat(p: *mut c_void, dim_one_based: u32, inputs: *const RmHandle, n: u32, out: *mut RmHandle) -> c_int;
    fn rm_eye(p: *mut c_void, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_find(p: *mut c_void, a: *const RmHandle, has_limit: c_int, limit: u64, direction_last: c_int,
               linear: *mut RmHandle, rows: *mut RmHandle, cols: *mut RmHandle, values: *mut RmHandle) -> c_int;
    fn rm_scatter_column(p: *mut c_void, m: *const RmHandle, col: u64, values: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_scatter_row(p: *mut c_void, m: *const RmHandle, row: u64, values: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_sub2ind(p: *mut c_void, dims: *const u64, strides: *const u64, ndims: u32, inputs: *const RmHandle, scalar_mask: *const u8,
                  len: u64, output_shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_ind2sub(p: *mut c_void, dims: *const u64, strides: *const u64, ndims: u32, indices: *const RmHandle, total: u64, len: u64,
                  output_shape: *const u64, rank: u32, outs: *mut RmHandle) -> c_int;
    fn rm_reduce_prod(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_reduce_mean_dim(p: *mut c_void, a: *const RmHandle, dim: u32, out: *mut RmHandle) -> c_int;
    fn rm_reduce_mean_nd(p: *mut c_void, a: *const RmHandle, dims: *const u32, n: u32, out: *mut RmHandle) -> c_int;
    fn rm_reduce_moments_nd(p: *mut c_void, a: *const RmHandle, dims: *const u32, n: u32, mean: *mut RmHandle, ex2: *mut RmHandle) -> c_int;
    fn rm_warmup(p: *mut c_void) -> c_int;
    fn rm_ones(p: *mut c_void, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_reshape(p: *mut c_void, h: *const RmHandle, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_random_uniform(p: *mut c_void, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_random_normal(p: *mut c_void, shape: *const u64, rank: u32, out: *mut RmHandle) -> c_int;
    fn rm_reduce_max(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_reduce_min(p: *mut c_void, a: *const RmHandle, out: *mut RmHandle) -> c_int;
    fn rm_reduce_min_dim(p: *mut c_void, a: *const RmHandle, dim: u32, vals: *mut RmHandle, idx: *mut RmHandle) -> c_int;
    fn rm_linsolve(p: *mut c_void, lhs: *const RmHandle, rhs: *const RmHandle, opt: *const RmLinsolveOptions, solution: *mut RmHandle, rcond: *mut f64) -> c_int;
    fn rm_default_reduction_workgroup_size(p: *mut c_void) -> u32;
    fn rm_two_pass_threshold(p: *mut c_void) -> u64;
}

pub struct CudaProvider {
    raw: *mut c_void,
    device_id: u32,
}
// The C library is thread-safe (one mutex around the buffer table, stream-ordered allocation): rm_accel.h header.
unsafe impl Send for CudaProvider {}
unsafe impl Sync for CudaProvider {}

fn err() -> anyhow::Error {
    let msg = unsafe { CStr::from_ptr(rm_last_error()) }.to_string_lossy().into_owned();
    anyhow!(msg)
}
fn check(status: c_int) -> Result<()> {
    if status == 0 { Ok(()) } else { Err(err()) }
}
fn to_raw(h: &GpuTensorHandle) -> Result<RmHandle> {
    if h.shape.len() > RM_MAX_RANK {
        return Err(anyhow!("tensor rank {} exceeds RM_MAX_RANK", h.shape.len()));
    }
    let mut shape = [0u64; RM_MAX_RANK];
    for (d, &s) in h.shape.iter().enumerate() {
        shape[d] = s as u64;
    }
    Ok(RmHandle { buffer_id: h.buffer_id, device_id: h.device_id, rank: h.shape.len() as u32, shape })
}
fn from_raw(h: &RmHandle) -> GpuTensorHandle {
    let handle = GpuTensorHandle {
        shape: h.shape[..h.rank as usize].iter().map(|&d| d as usize).collect(),
        device_id: h.device_id,
        buffer_id: h.buffer_id,
    };
    // the side tables the rest of the runtime consults (accelerate-api/src/lib.rs:132-245)
    runmat_accelerate_api::set_handle_precision(&handle, ProviderPrecision::F64);
    runmat_accelerate_api::set_handle_storage(&handle, GpuTensorStorage::Real);
    runmat_accelerate_api::set_handle_logical(&handle, false);
    handle
}
fn empty_raw() -> RmHandle {
    RmHandle { buffer_id: 0, device_id: 0, rank: 0, shape: [0; RM_MAX_RANK] }
}

impl CudaProvider {
    /// ProviderTelemetry::kernel_launches: the library's bounded log (64 events, newest last)
    fn launch_log(&self) -> Vec<KernelLaunchTelemetry> {
        let mut ev: Vec<RmKernelLaunchEvent> = Vec::with_capacity(64);
        let mut n: u32 = 0;
        unsafe {
            if rm_kernel_launch_log(self.raw, ev.as_mut_ptr(), 64, &mut n) != 0 { return Vec::new(); }
            ev.set_len(n as usize);
        }
        let attrs = |a: &[RmKernelAttr], k: u32| a[..k as usize].iter().map(|x| KernelAttrTelemetry { key: c_chars(&x.key), value: x.value }).collect::<Vec<_>>();
        ev.iter().map(|e| KernelLaunchTelemetry {
            kernel: c_chars(&e.kernel),
            precision: Some(if e.precision == 1 { "f64".to_string() } else { "f32".to_string() }),
            shape: attrs(&e.shape, e.n_shape),
            tuning: attrs(&e.tuning, e.n_tuning),
        }).collect()
    }
    /// `cuda_ordinal`: CUDA device; `device_id`: the id from `runmat_accelerate_api::next_device_id()` (lib.rs:3279).
    pub fn new(cuda_ordinal: i32, device_id: u32) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { rm_provider_create(cuda_ordinal, device_id, 1 /* RM_F64 */, &mut raw) })?;
        Ok(Self { raw, device_id })
    }
    fn binary(&self, op: c_int, a: &GpuTensorHandle, b: &GpuTensorHandle) -> Result<GpuTensorHandle> {
        let (ra, rb, mut out) = (to_raw(a)?, to_raw(b)?, empty_raw());
        check(unsafe { rm_elem_binary(self.raw, op, &ra, &rb, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn unary(&self, op: c_int, a: &GpuTensorHandle) -> Result<GpuTensorHandle> {
        let (ra, mut out) = (to_raw(a)?, empty_raw());
        check(unsafe { rm_unary(self.raw, op, &ra, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn ctor(&self, f: unsafe extern "C" fn(*mut c_void, *const u64, u32, *mut RmHandle) -> c_int, shape: &[usize]) -> Result<GpuTensorHandle> {
        let s: Vec<u64> = shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { f(self.raw, s.as_ptr(), s.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn scalar(&self, op: c_int, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> {
        let (ra, mut out) = (to_raw(a)?, empty_raw());
        check(unsafe { rm_scalar_op_apply(self.raw, op, &ra, s, &mut out) })?;
        Ok(from_raw(&out))
    }
}

impl Drop for CudaProvider {
    fn drop(&mut self) {
        unsafe { rm_provider_destroy(self.raw) };
    }
}

macro_rules! ready { ($e:expr) => { Box::pin(async move { $e }) }; }

impl AccelProvider for CudaProvider {
    fn upload(&self, host: &HostTensorView) -> Result<GpuTensorHandle> {
        let shape: Vec<u64> = host.shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_upload(self.raw, host.data.as_ptr(), shape.as_ptr(), shape.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn download<'a>(&'a self, h: &'a GpuTensorHandle) -> AccelDownloadFuture<'a> {
        ready!({
            let raw = to_raw(h)?;
            let n: usize = h.shape.iter().product();
            let mut data = vec![0.0f64; n];
            check(unsafe { rm_download(self.raw, &raw, data.as_mut_ptr(), n as u64) })?;
            Ok(HostTensorOwned { data, shape: h.shape.clone(), storage: GpuTensorStorage::Real })
        })
    }
    fn free(&self, h: &GpuTensorHandle) -> Result<()> {
        check(unsafe { rm_free(self.raw, &to_raw(h)?) })?;
        runmat_accelerate_api::clear_handle_precision(h);
        runmat_accelerate_api::clear_handle_class_name(h);
        runmat_accelerate_api::clear_handle_logical(h);
        runmat_accelerate_api::clear_handle_storage(h);
        Ok(())
    }
    fn device_info(&self) -> String {
        let mut buf = vec![0 as c_char; 512];
        unsafe { rm_device_info_string(self.raw, buf.as_mut_ptr(), buf.len()) };
        unsafe { CStr::from_ptr(buf.as_ptr()) }.to_string_lossy().into_owned()
    }
    fn device_id(&self) -> u32 { self.device_id }
    fn precision(&self) -> ProviderPrecision { ProviderPrecision::F64 }
    fn device_info_struct(&self) -> ApiDeviceInfo {
        ApiDeviceInfo { device_id: self.device_id, name: self.device_info(), vendor: "NVIDIA".into(), memory_bytes: None, backend: Some("cuda-sm_100a".into()) }
    }
    fn read_scalar(&self, h: &GpuTensorHandle, idx: usize) -> Result<f64> {
        let mut v = 0.0;
        check(unsafe { rm_read_scalar(self.raw, &to_raw(h)?, idx as u64, &mut v) })?;
        Ok(v)
    }
    fn zeros(&self, shape: &[usize]) -> Result<GpuTensorHandle> {
        let s: Vec<u64> = shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_zeros(self.raw, s.as_ptr(), s.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn fill(&self, shape: &[usize], value: f64) -> Result<GpuTensorHandle> {
        let s: Vec<u64> = shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_fill(self.raw, s.as_ptr(), s.len() as u32, value, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn ones(&self, shape: &[usize]) -> Result<GpuTensorHandle> { self.ctor(rm_ones, shape) }
    fn zeros_like(&self, prototype: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.zeros(&prototype.shape) }
    fn ones_like(&self, prototype: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.ctor(rm_ones, &prototype.shape) }
    fn fill_like(&self, prototype: &GpuTensorHandle, value: f64) -> Result<GpuTensorHandle> { self.fill(&prototype.shape, value) }
    fn random_uniform(&self, shape: &[usize]) -> Result<GpuTensorHandle> { self.ctor(rm_random_uniform, shape) }
    fn random_uniform_like(&self, prototype: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.ctor(rm_random_uniform, &prototype.shape) }
    fn random_normal(&self, shape: &[usize]) -> Result<GpuTensorHandle> { self.ctor(rm_random_normal, shape) }
    fn random_normal_like(&self, prototype: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.ctor(rm_random_normal, &prototype.shape) }
    fn reshape(&self, handle: &GpuTensorHandle, new_shape: &[usize]) -> Result<GpuTensorHandle> {
        // metadata-only, like the trait default (lib.rs:2676-2684); the library additionally rejects element-count changes
        let s: Vec<u64> = new_shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_reshape(self.raw, &to_raw(handle)?, s.as_ptr(), s.len() as u32, &mut out) })?;
        Ok(GpuTensorHandle { shape: new_shape.to_vec(), device_id: handle.device_id, buffer_id: handle.buffer_id })
    }
    fn linspace(&self, start: f64, stop: f64, count: usize) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_linspace(self.raw, start, stop, count as u64, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn transpose(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_transpose(self.raw, &to_raw(a)?, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn gather_linear(&self, source: &GpuTensorHandle, indices: &[u32], output_shape: &[usize]) -> Result<GpuTensorHandle> {
        let s: Vec<u64> = output_shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_gather_linear(self.raw, &to_raw(source)?, indices.as_ptr(), indices.len() as u64, s.as_ptr(), s.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn scatter_linear(&self, target: &GpuTensorHandle, indices: &[u32], values: &GpuTensorHandle) -> Result<()> {
        check(unsafe { rm_scatter_linear(self.raw, &to_raw(target)?, indices.as_ptr(), indices.len() as u64, &to_raw(values)?) })
    }

    // rm_binary_op / rm_unary_op / rm_scalar_op numbering: include/rm_accel.h
    fn elem_add<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(0, a, b)) }
    fn elem_sub<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(1, a, b)) }
    fn elem_mul<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(2, a, b)) }
    fn elem_div<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(3, a, b)) }
    fn elem_pow<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(4, a, b)) }
    fn elem_max<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(5, a, b)) }
    fn elem_min<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(6, a, b)) }
    fn unary_sin<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(0, a)) }
    fn unary_cos<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(1, a)) }
    fn unary_tanh<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(8, a)) }
    fn unary_exp<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(12, a)) }
    fn unary_log<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(14, a)) }
    fn unary_sqrt<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(18, a)) }
    fn unary_abs<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(19, a)) }
    fn unary_tan<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(2, a)) }
    fn unary_asin<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(3, a)) }
    fn unary_acos<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(4, a)) }
    fn unary_atan<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(5, a)) }
    fn unary_sinh<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(6, a)) }
    fn unary_cosh<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(7, a)) }
    fn unary_asinh<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(9, a)) }
    fn unary_acosh<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(10, a)) }
    fn unary_atanh<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(11, a)) }
    fn unary_expm1<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(13, a)) }
    fn unary_log2<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(15, a)) }
    fn unary_log10<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(16, a)) }
    fn unary_log1p<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(17, a)) }
    fn unary_sign<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(20, a)) }
    fn unary_floor<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(21, a)) }
    fn unary_ceil<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(22, a)) }
    fn unary_round<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(23, a)) }
    fn unary_fix<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(24, a)) }
    fn unary_pow2<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(26, a)) }
    fn unary_heaviside<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(27, a)) }
    fn unary_single<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(28, a)) }
    fn unary_double<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(29, a)) }
    fn unary_erf<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(35, a)) }
    fn unary_gamma<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(36, a)) }
    fn unary_gammaln<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.unary(37, a)) }
    fn logical_isnan(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.unary(30, a) }
    fn logical_isinf(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.unary(31, a) }
    fn logical_isfinite(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.unary(32, a) }
    fn map_nan_to_zero(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.unary(33, a) }
    fn not_nan_mask(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> { self.unary(34, a) }
    fn elem_hypot<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(7, a, b)) }
    fn elem_atan2<'a>(&'a self, y: &'a GpuTensorHandle, x: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(8, y, x)) }
    fn elem_ge<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(11, a, b)) }
    fn elem_le<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(12, a, b)) }
    fn elem_lt<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(13, a, b)) }
    fn elem_gt<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(14, a, b)) }
    fn elem_eq<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(15, a, b)) }
    fn elem_ne<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> { ready!(self.binary(16, a, b)) }
    fn scalar_add(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(0, a, s) }
    fn scalar_sub(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(1, a, s) }
    fn scalar_mul(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(2, a, s) }
    fn scalar_div(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(3, a, s) }
    fn scalar_rsub(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(4, a, s) }
    fn scalar_rdiv(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(5, a, s) }
    fn scalar_max(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(6, a, s) }
    fn scalar_min(&self, a: &GpuTensorHandle, s: f64) -> Result<GpuTensorHandle> { self.scalar(7, a, s) }

    fn fused_elementwise(&self, shader: &str, inputs: &[GpuTensorHandle], output_shape: &[usize], len: usize) -> Result<GpuTensorHandle> {
        let sh = CString::new(shader)?;
        let raws: Vec<RmHandle> = inputs.iter().map(to_raw).collect::<Result<_>>()?;
        let s: Vec<u64> = output_shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_fused_elementwise(self.raw, sh.as_ptr(), raws.as_ptr(), raws.len() as u32, s.as_ptr(), s.len() as u32, len as u64, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn fused_elementwise_multi(&self, shader: &str, inputs: &[GpuTensorHandle], output_shape: &[usize], len: usize, num_outputs: usize) -> Result<Vec<GpuTensorHandle>> {
        let sh = CString::new(shader)?;
        let raws: Vec<RmHandle> = inputs.iter().map(to_raw).collect::<Result<_>>()?;
        let s: Vec<u64> = output_shape.iter().map(|&d| d as u64).collect();
        let mut outs = vec![empty_raw(); num_outputs];
        check(unsafe { rm_fused_elementwise_multi(self.raw, sh.as_ptr(), raws.as_ptr(), raws.len() as u32, s.as_ptr(), s.len() as u32, len as u64, num_outputs as u32, outs.as_mut_ptr()) })?;
        Ok(outs.iter().map(from_raw).collect())
    }
    #[allow(clippy::too_many_arguments)]
    fn fused_reduction(&self, shader: &str, inputs: &[GpuTensorHandle], output_shape: &[usize], reduce_len: usize, num_slices: usize, workgroup_size: u32, flavor: ReductionFlavor) -> Result<GpuTensorHandle> {
        let sh = CString::new(shader)?;
        let raws: Vec<RmHandle> = inputs.iter().map(to_raw).collect::<Result<_>>()?;
        let s: Vec<u64> = output_shape.iter().map(|&d| d as u64).collect();
        let (kind, scale) = match flavor { ReductionFlavor::Sum => (0, 1.0), ReductionFlavor::Mean => (1, 1.0), ReductionFlavor::CustomScale(v) => (2, v) };
        let mut out = empty_raw();
        check(unsafe { rm_fused_reduction(self.raw, sh.as_ptr(), raws.as_ptr(), raws.len() as u32, s.as_ptr(), s.len() as u32, reduce_len as u64, num_slices as u64, workgroup_size, kind, scale, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn fused_cache_counters(&self) -> (u64, u64) {
        let (mut h, mut m) = (0u64, 0u64);
        unsafe { rm_fused_cache_counters(self.raw, &mut h, &mut m) };
        (h, m)
    }

    fn reduce_sum<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_sum(self.raw, &to_raw(a)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_sum_dim<'a>(&'a self, a: &'a GpuTensorHandle, dim: usize) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_sum_dim(self.raw, &to_raw(a)?, dim as u32, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_mean<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_mean(self.raw, &to_raw(a)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_max<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_max(self.raw, &to_raw(a)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_min<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_min(self.raw, &to_raw(a)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_min_dim<'a>(&'a self, a: &'a GpuTensorHandle, dim: usize) -> AccelProviderFuture<'a, ReduceDimResult> {
        ready!({
            let (mut v, mut i) = (empty_raw(), empty_raw());
            check(unsafe { rm_reduce_min_dim(self.raw, &to_raw(a)?, dim as u32, &mut v, &mut i) })?;
            Ok(ReduceDimResult { values: from_raw(&v), indices: from_raw(&i) })
        })
    }
    fn default_reduction_workgroup_size(&self) -> u32 { unsafe { rm_default_reduction_workgroup_size(self.raw) } }
    fn two_pass_threshold(&self) -> usize { unsafe { rm_two_pass_threshold(self.raw) as usize } }
    fn reduce_max_dim<'a>(&'a self, a: &'a GpuTensorHandle, dim: usize) -> AccelProviderFuture<'a, ReduceDimResult> {
        ready!({
            let (mut v, mut i) = (empty_raw(), empty_raw());
            check(unsafe { rm_reduce_max_dim(self.raw, &to_raw(a)?, dim as u32, &mut v, &mut i) })?;
            Ok(ReduceDimResult { values: from_raw(&v), indices: from_raw(&i) })
        })
    }

    fn matmul<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_matmul(self.raw, &to_raw(a)?, &to_raw(b)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn matmul_epilogue<'a>(&'a self, a: &'a GpuTensorHandle, b: &'a GpuTensorHandle, ep: &'a MatmulEpilogue) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({
            let row = ep.row_scale.as_ref().map(to_raw).transpose()?;
            let col = ep.col_scale.as_ref().map(to_raw).transpose()?;
            let diag = ep.diag_output.as_ref().map(to_raw).transpose()?;
            let raw = RmMatmulEpilogue {
                alpha: ep.alpha, beta: ep.beta,
                row_scale: row.as_ref().map_or(std::ptr::null(), |h| h as *const _),
                col_scale: col.as_ref().map_or(std::ptr::null(), |h| h as *const _),
                row_op: matches!(ep.row_op, ScaleOp::Divide) as c_int, col_op: matches!(ep.col_op, ScaleOp::Divide) as c_int,
                has_clamp_min: ep.clamp_min.is_some() as c_int, clamp_min: ep.clamp_min.unwrap_or(0.0),
                has_clamp_max: ep.clamp_max.is_some() as c_int, clamp_max: ep.clamp_max.unwrap_or(0.0),
                has_pow: ep.pow_exponent.is_some() as c_int, pow_exponent: ep.pow_exponent.unwrap_or(0.0),
                diag_output: diag.as_ref().map_or(std::ptr::null(), |h| h as *const _),
            };
            let mut out = empty_raw();
            check(unsafe { rm_matmul_epilogue_apply(self.raw, &to_raw(a)?, &to_raw(b)?, &raw, &mut out) })?;
            Ok(from_raw(&out))
        })
    }
    fn image_normalize<'a>(&'a self, input: &'a GpuTensorHandle, d: &'a ImageNormalizeDescriptor) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({
            let raw = RmImageNormalizeDesc {
                batch: d.batch as u64, height: d.height as u64, width: d.width as u64, epsilon: d.epsilon,
                has_gain: d.gain.is_some() as c_int, gain: d.gain.unwrap_or(1.0), has_bias: d.bias.is_some() as c_int, bias: d.bias.unwrap_or(0.0),
                has_gamma: d.gamma.is_some() as c_int, gamma: d.gamma.unwrap_or(1.0), clamp_zero: d.clamp_zero as c_int,
            };
            let mut out = empty_raw();
            check(unsafe { rm_image_normalize(self.raw, &to_raw(input)?, &raw, &mut out) })?;
            Ok(from_raw(&out))
        })
    }
    fn stochastic_evolution(&self, state: &GpuTensorHandle, drift: f64, scale: f64, steps: u32) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_stochastic_evolution(self.raw, &to_raw(state)?, drift, scale, steps, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn set_rng_state(&self, state: u64) -> Result<()> { check(unsafe { rm_set_rng_state(self.raw, state) }) }

    // ---- "next" rows: every method below is again a 1:1 forward -------------------------------------------------
    fn mldivide<'a>(&'a self, lhs: &'a GpuTensorHandle, rhs: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        // RM_UNSUPPORTED (least-squares / singular / ill-conditioned) surfaces as Err => host SVD path, as with wgpu today
        ready!({ let mut out = empty_raw(); check(unsafe { rm_mldivide(self.raw, &to_raw(lhs)?, &to_raw(rhs)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn linsolve<'a>(&'a self, lhs: &'a GpuTensorHandle, rhs: &'a GpuTensorHandle, o: &'a ProviderLinsolveOptions) -> AccelProviderFuture<'a, ProviderLinsolveResult> {
        // triangular systems and (without an rcond request) square general systems stay on the device; everything else is Err => host
        ready!({
            let raw = RmLinsolveOptions {
                lower: o.lower as c_int, upper: o.upper as c_int, rectangular: o.rectangular as c_int, transposed: o.transposed as c_int,
                conjugate: o.conjugate as c_int, symmetric: o.symmetric as c_int, posdef: o.posdef as c_int, need_rcond: o.need_rcond as c_int,
                has_rcond: o.rcond.is_some() as c_int, rcond: o.rcond.unwrap_or(0.0),
            };
            let (mut out, mut rcond) = (empty_raw(), f64::NAN);
            check(unsafe { rm_linsolve(self.raw, &to_raw(lhs)?, &to_raw(rhs)?, &raw, &mut out, &mut rcond) })?;
            Ok(ProviderLinsolveResult { solution: from_raw(&out), reciprocal_condition: rcond })
        })
    }
    fn mrdivide<'a>(&'a self, lhs: &'a GpuTensorHandle, rhs: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_mrdivide(self.raw, &to_raw(lhs)?, &to_raw(rhs)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn syrk(&self, a: &GpuTensorHandle) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_syrk(self.raw, &to_raw(a)?, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn matmul_power_step<'a>(&'a self, lhs: &'a GpuTensorHandle, rhs: &'a GpuTensorHandle, ep: &'a PowerStepEpilogue) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_matmul_power_step(self.raw, &to_raw(lhs)?, &to_raw(rhs)?, ep.epsilon, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn covariance<'a>(&'a self, matrix: &'a GpuTensorHandle, second: Option<&'a GpuTensorHandle>, weights: Option<&'a GpuTensorHandle>,
                      options: &'a CovarianceOptions) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({
            if second.is_some() || weights.is_some() || options.has_weight_vector || options.rows != CovRows::All {
                return Err(anyhow!("covariance: this form is not supported by provider"));  // host fallback
            }
            let mut out = empty_raw();
            check(unsafe { rm_covariance(self.raw, &to_raw(matrix)?, (options.normalization == CovNormalization::Biased) as c_int, &mut out) })?;
            Ok(from_raw(&out))
        })
    }
    fn diag_extract(&self, matrix: &GpuTensorHandle, offset: isize) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_diag_extract(self.raw, &to_raw(matrix)?, offset as i64, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn imfilter<'a>(&'a self, image: &'a GpuTensorHandle, kernel: &'a GpuTensorHandle, o: &'a ImfilterOptions) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({
            let raw = RmImfilterOptions {
                padding: match o.padding { ImfilterPadding::Constant => 0, ImfilterPadding::Replicate => 1, ImfilterPadding::Symmetric => 2, ImfilterPadding::Circular => 3 },
                constant_value: o.constant_value,
                shape: match o.shape { ImfilterShape::Same => 0, ImfilterShape::Full => 1, ImfilterShape::Valid => 2 },
                mode: match o.mode { ImfilterMode::Correlation => 0, ImfilterMode::Convolution => 1 },
            };
            let mut out = empty_raw();
            check(unsafe { rm_imfilter(self.raw, &to_raw(image)?, &to_raw(kernel)?, &raw, &mut out) })?;
            Ok(from_raw(&out))
        })
    }
    fn conv2d(&self, signal: &GpuTensorHandle, kernel: &GpuTensorHandle, mode: ProviderConvMode) -> Result<GpuTensorHandle> {
        let m = match mode { ProviderConvMode::Full => 0, ProviderConvMode::Same => 1, ProviderConvMode::Valid => 2 };
        let mut out = empty_raw();
        check(unsafe { rm_conv2d(self.raw, &to_raw(signal)?, &to_raw(kernel)?, m, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn permute(&self, handle: &GpuTensorHandle, order: &[usize]) -> Result<GpuTensorHandle> {
        let o: Vec<u32> = order.iter().map(|&d| d as u32).collect();  // zero-based, as the trait passes it
        let mut out = empty_raw();
        check(unsafe { rm_permute(self.raw, &to_raw(handle)?, o.as_ptr(), o.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn repmat(&self, handle: &GpuTensorHandle, reps: &[usize]) -> Result<GpuTensorHandle> {
        let r: Vec<u64> = reps.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_repmat(self.raw, &to_raw(handle)?, r.as_ptr(), r.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn cat(&self, dim: usize, inputs: &[GpuTensorHandle]) -> Result<GpuTensorHandle> {
        let raws: Vec<RmHandle> = inputs.iter().map(to_raw).collect::<Result<_>>()?;
        let mut out = empty_raw();
        check(unsafe { rm_cat(self.raw, dim as u32, raws.as_ptr(), raws.len() as u32, &mut out) })?;  // dim is one-based (lib.rs:2686)
        Ok(from_raw(&out))
    }
    fn eye(&self, shape: &[usize]) -> Result<GpuTensorHandle> {
        let s: Vec<u64> = shape.iter().map(|&d| d as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_eye(self.raw, s.as_ptr(), s.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn find(&self, a: &GpuTensorHandle, limit: Option<usize>, direction: FindDirection) -> Result<ProviderFindResult> {
        let (mut lin, mut rows, mut cols, mut vals) = (empty_raw(), empty_raw(), empty_raw(), empty_raw());
        check(unsafe {
            rm_find(self.raw, &to_raw(a)?, limit.is_some() as c_int, limit.unwrap_or(0) as u64, (direction == FindDirection::Last) as c_int,
                    &mut lin, &mut rows, &mut cols, &mut vals)
        })?;
        Ok(ProviderFindResult { linear: from_raw(&lin), rows: from_raw(&rows), cols: from_raw(&cols), values: Some(from_raw(&vals)) })
    }
    fn scatter_column(&self, matrix: &GpuTensorHandle, col_index: usize, values: &GpuTensorHandle) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_scatter_column(self.raw, &to_raw(matrix)?, col_index as u64, &to_raw(values)?, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn scatter_row(&self, matrix: &GpuTensorHandle, row_index: usize, values: &GpuTensorHandle) -> Result<GpuTensorHandle> {
        let mut out = empty_raw();
        check(unsafe { rm_scatter_row(self.raw, &to_raw(matrix)?, row_index as u64, &to_raw(values)?, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn sub2ind(&self, dims: &[usize], strides: &[usize], inputs: &[&GpuTensorHandle], scalar_mask: &[bool], len: usize, output_shape: &[usize]) -> Result<GpuTensorHandle> {
        let (d, st): (Vec<u64>, Vec<u64>) = (dims.iter().map(|&v| v as u64).collect(), strides.iter().map(|&v| v as u64).collect());
        let raws: Vec<RmHandle> = inputs.iter().map(|h| to_raw(h)).collect::<Result<_>>()?;
        let mask: Vec<u8> = scalar_mask.iter().map(|&b| b as u8).collect();
        let os: Vec<u64> = output_shape.iter().map(|&v| v as u64).collect();
        let mut out = empty_raw();
        check(unsafe { rm_sub2ind(self.raw, d.as_ptr(), st.as_ptr(), d.len() as u32, raws.as_ptr(), mask.as_ptr(), len as u64, os.as_ptr(), os.len() as u32, &mut out) })?;
        Ok(from_raw(&out))
    }
    fn supports_ind2sub(&self) -> bool { true }
    fn ind2sub(&self, dims: &[usize], strides: &[usize], indices: &GpuTensorHandle, total: usize, len: usize, output_shape: &[usize]) -> Result<Vec<GpuTensorHandle>> {
        let (d, st): (Vec<u64>, Vec<u64>) = (dims.iter().map(|&v| v as u64).collect(), strides.iter().map(|&v| v as u64).collect());
        let os: Vec<u64> = output_shape.iter().map(|&v| v as u64).collect();
        let mut outs = vec![empty_raw(); d.len()];
        check(unsafe { rm_ind2sub(self.raw, d.as_ptr(), st.as_ptr(), d.len() as u32, &to_raw(indices)?, total as u64, len as u64, os.as_ptr(), os.len() as u32, outs.as_mut_ptr()) })?;
        Ok(outs.iter().map(from_raw).collect())
    }
    fn reduce_prod<'a>(&'a self, a: &'a GpuTensorHandle) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_prod(self.raw, &to_raw(a)?, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_mean_dim<'a>(&'a self, a: &'a GpuTensorHandle, dim: usize) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({ let mut out = empty_raw(); check(unsafe { rm_reduce_mean_dim(self.raw, &to_raw(a)?, dim as u32, &mut out) })?; Ok(from_raw(&out)) })
    }
    fn reduce_mean_nd<'a>(&'a self, a: &'a GpuTensorHandle, dims_zero_based: &'a [usize]) -> AccelProviderFuture<'a, GpuTensorHandle> {
        ready!({
            let d: Vec<u32> = dims_zero_based.iter().map(|&v| v as u32).collect();
            let mut out = empty_raw();
            check(unsafe { rm_reduce_mean_nd(self.raw, &to_raw(a)?, d.as_ptr(), d.len() as u32, &mut out) })?;
            Ok(from_raw(&out))
        })
    }
    fn reduce_moments_nd<'a>(&'a self, a: &'a GpuTensorHandle, dims_zero_based: &'a [usize]) -> AccelProviderFuture<'a, ProviderMoments2> {
        ready!({
            let d: Vec<u32> = dims_zero_based.iter().map(|&v| v as u32).collect();
            let (mut mean, mut ex2) = (empty_raw(), empty_raw());
            check(unsafe { rm_reduce_moments_nd(self.raw, &to_raw(a)?, d.as_ptr(), d.len() as u32, &mut mean, &mut ex2) })?;
            Ok(ProviderMoments2 { mean: from_raw(&mean), ex2: from_raw(&ex2) })
        })
    }
    fn warmup(&self) { unsafe { rm_warmup(self.raw) }; }

    fn spawn_handle_concurrency(&self) -> SpawnHandleConcurrency {
        match unsafe { rm_spawn_handle_concurrency_policy(self.raw) } {
            0 => SpawnHandleConcurrency::ImmutableShare,
            1 => SpawnHandleConcurrency::CopyOnWrite,
            2 => SpawnHandleConcurrency::SynchronizedMutation,
            _ => SpawnHandleConcurrency::Reject,
        }
    }

    fn telemetry_snapshot(&self) -> ProviderTelemetry {
        let mut t = RmTelemetry::default();
        unsafe { rm_telemetry_snapshot(self.raw, &mut t) };
        let cv = |s: RmDispatchStats| ProviderDispatchStats { count: s.count, total_wall_time_ns: s.total_wall_time_ns };
        ProviderTelemetry {
            fused_elementwise: cv(t.fused_elementwise), fused_reduction: cv(t.fused_reduction), matmul: cv(t.matmul),
            linsolve: cv(t.linsolve), mldivide: cv(t.mldivide), mrdivide: cv(t.mrdivide),
            upload_bytes: t.upload_bytes, download_bytes: t.download_bytes, solve_fallbacks: Vec::new(),
            fusion_cache_hits: t.fusion_cache_hits, fusion_cache_misses: t.fusion_cache_misses,
            bind_group_cache_hits: 0, bind_group_cache_misses: 0, bind_group_cache_by_layout: None, kernel_launches: self.launch_log(),
        }
    }
    fn reset_telemetry(&self) { unsafe { rm_reset_telemetry(self.raw) }; }
}

/// Registration, next to `register_wgpu_provider` (backend/wgpu/provider.rs:13):
pub fn register_cuda_provider(cuda_ordinal: i32) -> Result<&'static CudaProvider> {
    let id = runmat_accelerate_api::next_device_id();
    let provider: &'static CudaProvider = Box::leak(Box::new(CudaProvider::new(cuda_ordinal, id)?));
    unsafe { runmat_accelerate_api::register_provider(provider) };
    Ok(provider)
}
