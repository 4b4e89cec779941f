//! The `cuda_backend`: layout-compatible element types + slice functions over `ffi` (generated from the C header).
//!
//! In the reference the four types are plain tuple structs / a struct of four fields with no `repr` attribute
//! (src/backend/u64/field.rs:31-32, scalar.rs:26-27, src/edwards.rs:336-342, src/ristretto.rs:157-158).  To hand `&[T]` to C
//! as `*const u64` the layout must be guaranteed, so an in-tree integration adds
//!     #[cfg_attr(feature = "cuda_backend", repr(transparent))] pub struct FieldElement(pub [u64; 5]);
//!     #[cfg_attr(feature = "cuda_backend", repr(transparent))] pub struct Scalar(pub [u64; 5]);
//!     #[cfg_attr(feature = "cuda_backend", repr(C))]           pub struct EdwardsPoint { pub X, pub Y, pub Z, pub T }
//!     #[cfg_attr(feature = "cuda_backend", repr(transparent))] pub struct RistrettoPoint(pub EdwardsPoint);
//! and replaces the `standalone` declarations below by `use crate::{field::FieldElement, scalar::Scalar, ...}`.
pub mod ffi;

use core::ffi::c_void;
use core::ptr;

#[cfg(feature = "standalone")]
mod types {
    /// radix-2^52 limbs, little-endian limb order, canonical (field.rs:31-32)
    #[repr(transparent)] #[derive(Clone, Copy, Debug, PartialEq, Eq)] pub struct FieldElement(pub [u64; 5]);
    /// radix-2^52 limbs mod L (scalar.rs:26-27)
    #[repr(transparent)] #[derive(Clone, Copy, Debug, PartialEq, Eq)] pub struct Scalar(pub [u64; 5]);
    /// extended twisted-Edwards coordinates (edwards.rs:336-342)
    #[repr(C)] #[derive(Clone, Copy, Debug)] pub struct EdwardsPoint { pub X: FieldElement, pub Y: FieldElement, pub Z: FieldElement, pub T: FieldElement }
    /// newtype over EdwardsPoint (ristretto.rs:157-158)
    #[repr(transparent)] #[derive(Clone, Copy, Debug)] pub struct RistrettoPoint(pub EdwardsPoint);
    impl FieldElement { pub const fn zero() -> Self { FieldElement([0; 5]) } pub const fn one() -> Self { FieldElement([1, 0, 0, 0, 0]) } }
    impl EdwardsPoint {
        /// (0, 1, 1, 0)  edwards.rs:381-391
        pub const fn identity() -> Self { EdwardsPoint { X: FieldElement::zero(), Y: FieldElement::one(), Z: FieldElement::one(), T: FieldElement::zero() } }
    }
    impl RistrettoPoint { pub const fn identity() -> Self { RistrettoPoint(EdwardsPoint::identity()) } }
}
#[cfg(feature = "standalone")]
pub use types::{EdwardsPoint, FieldElement, RistrettoPoint, Scalar};

/// A non-zero status of the C ABI: > 0 argument errors (ZC_ERR_*), < 0 -(cudaError_t) / NCCL failures.
#[derive(Debug, Clone, PartialEq, Eq)]
pub struct ZcError { pub status: i32, pub message: String }

/// One device + one stream + scratch arenas.  Not `Sync`: a context is bound to one caller thread at a time.
pub struct Gpu { ctx: *mut ffi::zc_ctx }
unsafe impl Send for Gpu {}

impl Gpu {
    /// There is no CPU fallback: without a CUDA device this is an error.
    pub fn new(device: i32) -> Result<Gpu, ZcError> {
        let mut ctx = ptr::null_mut();
        match unsafe { ffi::zc_ctx_create(device, ptr::null_mut(), &mut ctx) } {
            0 => Ok(Gpu { ctx }),
            s => Err(ZcError { status: s, message: "zc_ctx_create failed (no CUDA device?)".into() }),
        }
    }
    fn status(&self, s: i32) -> Result<(), ZcError> {
        if s == 0 { return Ok(()); }
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::zc_last_error_string(self.ctx)) }.to_string_lossy().into_owned();
        Err(ZcError { status: s, message: msg })
    }
    /// Check every input of the hot-path calls on the device (limbs < 2^52, value < modulus): ZC_ERR_NONCANONICAL instead of
    /// a silently wrong result.  The reference's types expose their limbs, so a caller CAN build such values.
    pub fn set_validation(&self, on: bool) -> Result<(), ZcError> { self.status(unsafe { ffi::zc_ctx_set_validation(self.ctx, on as i32) }) }
    /// Pin a slice the caller owns so the host-pointer calls overlap H2D / kernel / D2H.
    pub fn pin<T>(&self, v: &mut [T]) -> Result<(), ZcError> {
        self.status(unsafe { ffi::zc_host_register(v.as_mut_ptr() as *mut c_void, core::mem::size_of_val(v)) })
    }
    pub fn unpin<T>(&self, v: &mut [T]) -> Result<(), ZcError> { self.status(unsafe { ffi::zc_host_unregister(v.as_mut_ptr() as *mut c_void) }) }

    // ---- FieldElement (field.rs:191-315) ------------------------------------------------------------------------
    /// out[i] = a[i] * b[i]   (batched `impl Mul<&FieldElement> for &FieldElement`, field.rs:250-262)
    pub fn fe_mul(&self, a: &[FieldElement], b: &[FieldElement], out: &mut [FieldElement]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == out.len());
        self.status(unsafe { ffi::zc_fe_mul_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }
    /// out[i] = a[i]^2   (Square, field.rs:302-315)
    pub fn fe_square(&self, a: &[FieldElement], out: &mut [FieldElement]) -> Result<(), ZcError> {
        assert!(a.len() == out.len());
        self.status(unsafe { ffi::zc_fe_square_batch(self.ctx, a.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }
    /// prod[i] = a[i] * b[i], sq[i] = a[i]^2 in one launch (BASELINE config 2)
    pub fn fe_mul_square(&self, a: &[FieldElement], b: &[FieldElement], prod: &mut [FieldElement], sq: &mut [FieldElement]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == prod.len() && a.len() == sq.len());
        self.status(unsafe { ffi::zc_fe_mul_square_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64,
                                                         prod.as_mut_ptr() as *mut u64, sq.as_mut_ptr() as *mut u64, a.len()) })
    }
    pub fn fe_add(&self, a: &[FieldElement], b: &[FieldElement], out: &mut [FieldElement]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == out.len());
        self.status(unsafe { ffi::zc_fe_add_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }
    pub fn fe_sub(&self, a: &[FieldElement], b: &[FieldElement], out: &mut [FieldElement]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == out.len());
        self.status(unsafe { ffi::zc_fe_sub_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }
    /// out[i] = a[i] / b[i]   (Div, field.rs:277-299; a zero divisor gives 0 where the reference asserts)
    pub fn fe_div(&self, a: &[FieldElement], b: &[FieldElement], out: &mut [FieldElement]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == out.len());
        self.status(unsafe { ffi::zc_fe_div_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }

    // ---- Scalar (scalar.rs:184-283) -----------------------------------------------------------------------------
    pub fn scalar_mul(&self, a: &[Scalar], b: &[Scalar], out: &mut [Scalar]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == out.len());
        self.status(unsafe { ffi::zc_scalar_mul_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }
    pub fn scalar_add(&self, a: &[Scalar], b: &[Scalar], out: &mut [Scalar]) -> Result<(), ZcError> {
        assert!(a.len() == b.len() && a.len() == out.len());
        self.status(unsafe { ffi::zc_scalar_add_batch(self.ctx, a.as_ptr() as *const u64, b.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, a.len()) })
    }
    /// Scalar::into_bits (scalar.rs:352-366) for a slice
    pub fn scalar_into_bits(&self, a: &[Scalar], out: &mut [[u8; 256]]) -> Result<(), ZcError> {
        assert!(a.len() == out.len());
        self.status(unsafe { ffi::zc_scalar_into_bits_batch(self.ctx, a.as_ptr() as *const u64, out.as_mut_ptr() as *mut u8, a.len()) })
    }

    // ---- EdwardsPoint / RistrettoPoint (edwards.rs:440-592, ristretto.rs:224-392) -------------------------------
    /// out[i] = p[i] + q[i]   (batched `impl Add<&EdwardsPoint> for &EdwardsPoint`, limb-exact, edwards.rs:465-489)
    pub fn point_add(&self, p: &[EdwardsPoint], q: &[EdwardsPoint], out: &mut [EdwardsPoint]) -> Result<(), ZcError> {
        assert!(p.len() == q.len() && p.len() == out.len());
        self.status(unsafe { ffi::zc_point_add_batch(self.ctx, p.as_ptr() as *const u64, q.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, p.len()) })
    }
    /// out[i] = 2 p[i]   (Double = self + self, edwards.rs:579-592, limb-exact)
    pub fn point_double(&self, p: &[EdwardsPoint], out: &mut [EdwardsPoint]) -> Result<(), ZcError> {
        assert!(p.len() == out.len());
        self.status(unsafe { ffi::zc_point_double_batch(self.ctx, p.as_ptr() as *const u64, out.as_mut_ptr() as *mut u64, p.len()) })
    }
    /// out[i] = points[i] * scalars[i]   (batched `double_and_add`, edwards.rs:102-120; strict = the CPU backend's limbs)
    pub fn scalar_mul_points(&self, points: &[RistrettoPoint], scalars: &[Scalar], out: &mut [RistrettoPoint], strict: bool) -> Result<(), ZcError> {
        assert!(points.len() == scalars.len() && points.len() == out.len());
        self.status(unsafe { ffi::zc_point_scalar_mul_batch(self.ctx, points.as_ptr() as *const u64, scalars.as_ptr() as *const u64,
                                                            out.as_mut_ptr() as *mut u64, points.len(),
                                                            if strict { ffi::ZC_SCALAR_MUL_STRICT } else { ffi::ZC_SCALAR_MUL_FAST }) })
    }
    /// out[i] = points[i].compress()   (ristretto.rs:398-425, the CPU backend's 32 bytes)
    pub fn compress(&self, points: &[RistrettoPoint], out: &mut [[u8; 32]]) -> Result<(), ZcError> {
        assert!(points.len() == out.len());
        self.status(unsafe { ffi::zc_ristretto_compress_batch(self.ctx, points.as_ptr() as *const u64, out.as_mut_ptr() as *mut u8, points.len()) })
    }
    /// sum_i scalars[i] * points[i]: what a bulletproofs caller writes today as zip / map / fold over `Mul` and `Add`
    pub fn msm(&self, points: &[RistrettoPoint], scalars: &[Scalar]) -> Result<RistrettoPoint, ZcError> {
        assert!(points.len() == scalars.len());
        let mut out = RistrettoPoint::identity();
        self.status(unsafe { ffi::zc_msm(self.ctx, points.as_ptr() as *const u64, scalars.as_ptr() as *const u64, points.len(), 16,
                                         &mut out as *mut RistrettoPoint as *mut u64) })?;
        Ok(out)
    }
    /// Fixed generators (bulletproofs G_i, H_i): upload once, prepare once.  `fixed_base` trades memory (16 x n rows on one
    /// GPU) for time (one bucket reduction, no doubling chain).
    pub fn generators(&self, points: &[RistrettoPoint], fixed_base: bool) -> Result<MsmGenerators<'_>, ZcError> {
        MsmGenerators::new(self, points, fixed_base)
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { ffi::zc_ctx_destroy(self.ctx); } } }

/// Owns the prepared operands / fixed-base tables of a generator vector on the device (zc_msm_generators).  The library keeps
/// its own copy of everything derived from the points: the slice passed to `new` is not referenced afterwards.
pub struct MsmGenerators<'g> { gpu: &'g Gpu, h: *mut ffi::zc_msm_generators, n: usize, scalars_dev: *mut c_void, out_dev: *mut c_void }

extern "C" {
    // the CUDA runtime calls the shim itself needs to stage the scalars (libcudart is already a dependency of the library)
    fn cudaMalloc(p: *mut *mut c_void, bytes: usize) -> i32;
    fn cudaFree(p: *mut c_void) -> i32;
    fn cudaMemcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: i32) -> i32;
}
const H2D: i32 = 1;
const D2H: i32 = 2;

impl<'g> MsmGenerators<'g> {
    fn new(gpu: &'g Gpu, points: &[RistrettoPoint], fixed_base: bool) -> Result<Self, ZcError> {
        let n = points.len();
        let mut dp = ptr::null_mut();
        let mut ds = ptr::null_mut();
        let mut dout = ptr::null_mut();
        let cuda = |e: i32| if e == 0 { Ok(()) } else { Err(ZcError { status: -e, message: "CUDA runtime error".into() }) };
        unsafe {
            cuda(cudaMalloc(&mut dp, n * 160))?;
            cuda(cudaMemcpy(dp, points.as_ptr() as *const c_void, n * 160, H2D))?;
            cuda(cudaMalloc(&mut ds, n * 40))?;
            cuda(cudaMalloc(&mut dout, 160))?;
        }
        let mut h = ptr::null_mut();
        let kind = if fixed_base { ffi::ZC_GEN_FIXED_BASE } else { ffi::ZC_GEN_PREPARED };
        let st = unsafe { ffi::zc_msm_generators_create_dev(gpu.ctx, dp as *const u64, n, kind, 16, 0, 1, &mut h) };
        unsafe { cudaFree(dp); }                       // the handle owns what it derived; the points are no longer needed
        gpu.status(st)?;
        Ok(MsmGenerators { gpu, h, n, scalars_dev: ds, out_dev: dout })
    }
    /// sum_i scalars[i] * G_i
    pub fn msm(&self, scalars: &[Scalar]) -> Result<RistrettoPoint, ZcError> {
        assert!(scalars.len() == self.n);
        let mut out = RistrettoPoint::identity();
        unsafe { cudaMemcpy(self.scalars_dev, scalars.as_ptr() as *const c_void, self.n * 40, H2D); }
        self.gpu.status(unsafe { ffi::zc_msm_gen_dev(self.gpu.ctx, self.h, self.scalars_dev as *const u64, 16, self.out_dev as *mut u64) })?;
        self.gpu.status(unsafe { ffi::zc_ctx_sync(self.gpu.ctx) })?;
        unsafe { cudaMemcpy(&mut out as *mut RistrettoPoint as *mut c_void, self.out_dev, 160, D2H); }
        Ok(out)
    }
}
impl<'g> Drop for MsmGenerators<'g> {
    fn drop(&mut self) {
        unsafe { ffi::zc_msm_generators_destroy(self.gpu.ctx, self.h); cudaFree(self.scalars_dev); cudaFree(self.out_dev); }
    }
}
