//! FFI shim over `libozl_b200` (C ABI: `include/ozl.h`).
//!
//! UNCOMPILED in this repository (no Rust toolchain in the build image).  It mirrors, call for
//! call, what `tests/test_gpu_msm.py` / `tests/test_gpu_ntt.py` do through ctypes, and replaces the
//! two arkworks entry points `plugins/arkworks` re-exports:
//!
//! * `ark_ec::msm::VariableBaseMSM::multi_scalar_mul`  (`pub use ec`,   plugins/arkworks/src/lib.rs:28-29)
//! * `ark_poly::EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place` (`pub use poly`, lib.rs:70-71)
//!
//! Memory conventions are arkworks' own: `Fp256/Fp384` hold Montgomery limbs (`fe.0.0`),
//! `into_repr()` yields the canonical `BigInteger256` the MSM takes, so slices cross the boundary
//! without conversion.  `GroupAffine` is not `repr(C)` (x, y, infinity + padding), hence the packing.

pub mod groth16; // Groth16B200<E>: ProofSystem (prove = one ozl_groth16_prove call; compile / verify with arkworks)

use ark_ec::AffineCurve;
use ark_ff::{BigInteger256, BigInteger384, PrimeField};
use std::os::raw::c_int;

#[repr(C)]
pub struct OzlCtx {
    _private: [u8; 0],
}

/// `ozl_csr` (include/ozl.h): one R1CS matrix, coefficients as indices into a shared table.
#[repr(C)]
pub(crate) struct OzlCsr {
    pub n_rows: u32,
    pub row_ptr: *const u32,
    pub col_idx: *const u32,
    pub coef_idx: *const u32,
}

/// The C ABI (`include/ozl.h`), declared once for the whole crate.
pub(crate) mod ffi {
    use super::{OzlComm, OzlCsr, OzlCtx};
    use std::os::raw::c_int;
    extern "C" {
        pub fn ozl_ctx_create(device: c_int, out: *mut *mut OzlCtx) -> c_int;
        pub fn ozl_ctx_destroy(ctx: *mut OzlCtx);
        pub fn ozl_msm_bases_upload(ctx: *mut OzlCtx, curve: c_int, bases: *const u64, inf_mask: *const u8, n: usize, handle: *mut u32) -> c_int;
        pub fn ozl_msm_bases_precompute(ctx: *mut OzlCtx, handle: u32, factor: c_int) -> c_int;
        pub fn ozl_msm_bases_free(ctx: *mut OzlCtx, handle: u32) -> c_int;
        pub fn ozl_msm(ctx: *mut OzlCtx, handle: u32, scalars: *const u64, n: usize, out_jacobian: *mut u64) -> c_int;
        pub fn ozl_ntt(ctx: *mut OzlCtx, field: c_int, data: *mut u64, log_n: u32, inverse: c_int, coset: c_int) -> c_int;
        // multi-GPU: one process (rank) per GPU, each holding its point range of the query vector
        pub fn ozl_comm_unique_id(id_out: *mut u8) -> c_int;
        pub fn ozl_comm_create(ctx: *mut OzlCtx, id: *const u8, rank: c_int, world: c_int, out: *mut *mut OzlComm) -> c_int;
        pub fn ozl_comm_destroy(comm: *mut OzlComm) -> c_int;
        pub fn ozl_msm_sharded(ctx: *mut OzlCtx, comm: *mut OzlComm, handle: u32, scalars: *const u64, n: usize, out_jacobian: *mut u64) -> c_int;
        // Groth16 prover resident on the device
        pub fn ozl_groth16_pk_create(
            ctx: *mut OzlCtx, pairing: c_int, n_constraints: u32, n_instance: u32, n_vars: u32,
            a: *const OzlCsr, b: *const OzlCsr, c: *const OzlCsr, coef_table: *const u64, n_coef: u32,
            a_query: u32, b_g1_query: u32, b_g2_query: u32, h_query: u32, l_query: u32,
            alpha_g1: *const u64, beta_g1: *const u64, delta_g1: *const u64, beta_g2: *const u64, delta_g2: *const u64,
            pk_handle: *mut u32,
        ) -> c_int;
        pub fn ozl_groth16_pk_destroy(ctx: *mut OzlCtx, pk_handle: u32) -> c_int;
        pub fn ozl_groth16_prove(
            ctx: *mut OzlCtx, pk_handle: u32, z: *const u64, r: *const u64, s: *const u64,
            proof_a: *mut u64, proof_b: *mut u64, proof_c: *mut u64, h_out: *mut u64,
        ) -> c_int;
    }
}
use ffi::*;

#[repr(C)]
pub struct OzlComm {
    _private: [u8; 0],
}
pub const OZL_COMM_ID_BYTES: usize = 128;

pub const OZL_BLS12_381_G1: c_int = 0;
pub const OZL_BLS12_381_G2: c_int = 1;
pub const OZL_BN254_G1: c_int = 2;
pub const OZL_BN254_G2: c_int = 3;
pub const OZL_BN254_FR: c_int = 0;
pub const OZL_BLS12_381_FR: c_int = 1;

/// Same opaque unit error as `plugins/arkworks/src/groth16.rs:35-45`.
#[derive(Clone, Copy, Debug, Default, Eq, PartialEq)]
pub struct Error;

/// One context per (device, stream); not `Sync` (use one per thread).
pub struct Context(*mut OzlCtx);

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut p = core::ptr::null_mut();
        match unsafe { ozl_ctx_create(device, &mut p) } {
            0 => Ok(Self(p)),
            _ => Err(Error),
        }
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ozl_ctx_destroy(self.0) }
    }
}

/// Device-resident bases of one query vector of a proving key (`a_query`, `b_g1_query`, ...):
/// they are constant across proofs (`ProvingContext<E>(pub ProvingKey<E>)`, groth16.rs:127-129).
pub struct G1Bases381<'c> {
    ctx: &'c Context,
    handle: u32,
    n: usize,
}

impl<'c> G1Bases381<'c> {
    pub fn upload(ctx: &'c Context, bases: &[ark_bls12_381::G1Affine], precompute: i32) -> Result<Self, Error> {
        let mut packed = Vec::<u64>::with_capacity(bases.len() * 12);
        let mut inf = vec![0u8; (bases.len() + 7) / 8];
        for (i, p) in bases.iter().enumerate() {
            if p.is_zero() {
                inf[i / 8] |= 1 << (i % 8);
                packed.extend_from_slice(&[0u64; 12]);
            } else {
                packed.extend_from_slice(&p.x.0 .0); // Montgomery limbs, as stored
                packed.extend_from_slice(&p.y.0 .0);
            }
        }
        let mut handle = 0u32;
        if unsafe { ozl_msm_bases_upload(ctx.0, OZL_BLS12_381_G1, packed.as_ptr(), inf.as_ptr(), bases.len(), &mut handle) } != 0 {
            return Err(Error);
        }
        if precompute > 1 && unsafe { ozl_msm_bases_precompute(ctx.0, handle, precompute) } != 0 {
            return Err(Error);
        }
        Ok(Self { ctx, handle, n: bases.len() })
    }

    /// Drop-in for `VariableBaseMSM::multi_scalar_mul(bases, scalars)`.
    pub fn multi_scalar_mul(&self, scalars: &[<ark_bls12_381::Fr as PrimeField>::BigInt]) -> Result<ark_bls12_381::G1Projective, Error> {
        let n = core::cmp::min(self.n, scalars.len()); // ark: size = min(bases.len(), scalars.len())
        let mut out = [0u64; 18];
        // BigInteger256 is a transparent [u64; 4]: the slice already is n x 4 canonical limbs
        if unsafe { ozl_msm(self.ctx.0, self.handle, scalars.as_ptr() as *const u64, n, out.as_mut_ptr()) } != 0 {
            return Err(Error);
        }
        let fq = |o: usize| ark_bls12_381::Fq::new(BigInteger384([out[o], out[o + 1], out[o + 2], out[o + 3], out[o + 4], out[o + 5]]));
        Ok(ark_bls12_381::G1Projective::new(fq(0), fq(6), fq(12))) // Fp::new takes the Montgomery representation
    }
}

/// The library's NCCL communicator over the ranks that share one MSM (one rank per GPU).
/// Rank 0 calls `Comm::unique_id()` and ships the 128 bytes to the other ranks by the
/// application's own means; every rank then calls `Comm::join`.
pub struct Comm<'c> {
    ctx: &'c Context,
    raw: *mut OzlComm,
}

impl<'c> Comm<'c> {
    pub fn unique_id() -> Result<[u8; OZL_COMM_ID_BYTES], Error> {
        let mut id = [0u8; OZL_COMM_ID_BYTES];
        match unsafe { ozl_comm_unique_id(id.as_mut_ptr()) } {
            0 => Ok(id),
            _ => Err(Error),
        }
    }
    pub fn join(ctx: &'c Context, id: &[u8; OZL_COMM_ID_BYTES], rank: i32, world: i32) -> Result<Self, Error> {
        let mut raw = core::ptr::null_mut();
        match unsafe { ozl_comm_create(ctx.0, id.as_ptr(), rank, world, &mut raw) } {
            0 => Ok(Self { ctx, raw }),
            _ => Err(Error),
        }
    }
}

impl Drop for Comm<'_> {
    fn drop(&mut self) {
        unsafe { ozl_comm_destroy(self.raw) };
    }
}

impl<'c> G1Bases381<'c> {
    /// `multi_scalar_mul` over a query vector split by point range across the ranks of `comm`:
    /// `self` holds THIS rank's range, `scalars` the matching slice; the full sum comes back on every rank.
    pub fn multi_scalar_mul_sharded(&self, comm: &Comm<'c>, scalars: &[<ark_bls12_381::Fr as PrimeField>::BigInt]) -> Result<ark_bls12_381::G1Projective, Error> {
        let n = core::cmp::min(self.n, scalars.len());
        let mut out = [0u64; 18];
        if unsafe { ozl_msm_sharded(self.ctx.0, comm.raw, self.handle, scalars.as_ptr() as *const u64, n, out.as_mut_ptr()) } != 0 {
            return Err(Error);
        }
        let fq = |o: usize| ark_bls12_381::Fq::new(BigInteger384([out[o], out[o + 1], out[o + 2], out[o + 3], out[o + 4], out[o + 5]]));
        Ok(ark_bls12_381::G1Projective::new(fq(0), fq(6), fq(12)))
    }
}

impl Drop for G1Bases381<'_> {
    fn drop(&mut self) {
        unsafe { ozl_msm_bases_free(self.ctx.0, self.handle) };
    }
}

/// Drop-in for `domain.{fft,ifft,coset_fft,coset_ifft}_in_place(&mut v)` with `v.len() == domain.size()`.
pub fn ntt_in_place_bn254(ctx: &Context, v: &mut [ark_bn254::Fr], inverse: bool, coset: bool) -> Result<(), Error> {
    debug_assert!(v.len().is_power_of_two());
    let _: &BigInteger256 = &v[0].0; // layout check: Fp256 wraps BigInteger256([u64; 4])
    match unsafe { ozl_ntt(ctx.0, OZL_BN254_FR, v.as_mut_ptr() as *mut u64, v.len().trailing_zeros(), inverse as c_int, coset as c_int) } {
        0 => Ok(()),
        _ => Err(Error),
    }
}

/// Curve-generic bases handle for stand-alone MSMs (BN254: 4 u64 limbs per Fq).  It BORROWS its context:
/// a handle cannot outlive the context that owns its device memory (`ozl_ctx_destroy` frees every bases
/// buffer), so the drop order of a struct's fields can no longer produce a dangling `ozl_msm_bases_free`.
/// `groth16::Groth16B200` does not use it: there the device pk owns the five query handles.
pub struct Bases<'c> {
    ctx: &'c Context,
    handle: u32,
    n: usize,
}

impl<'c> Bases<'c> {
    fn upload(ctx: &'c Context, curve: c_int, packed: &[u64], inf: &[u8], n: usize, precompute: i32) -> Result<Self, Error> {
        let mut handle = 0u32;
        if unsafe { ozl_msm_bases_upload(ctx.0, curve, packed.as_ptr(), inf.as_ptr(), n, &mut handle) } != 0 {
            return Err(Error);
        }
        if precompute > 1 && unsafe { ozl_msm_bases_precompute(ctx.0, handle, precompute) } != 0 {
            return Err(Error);
        }
        Ok(Self { ctx, handle, n })
    }

    pub fn upload_g1_bn254(ctx: &'c Context, bases: &[ark_bn254::G1Affine], precompute: i32) -> Result<Self, Error> {
        let mut packed = Vec::<u64>::with_capacity(bases.len() * 8);
        let mut inf = vec![0u8; (bases.len() + 7) / 8];
        for (i, p) in bases.iter().enumerate() {
            if p.is_zero() {
                inf[i / 8] |= 1 << (i % 8);
                packed.extend_from_slice(&[0u64; 8]);
            } else {
                packed.extend_from_slice(&p.x.0 .0);
                packed.extend_from_slice(&p.y.0 .0);
            }
        }
        Self::upload(ctx, OZL_BN254_G1, &packed, &inf, bases.len(), precompute)
    }

    pub fn upload_g2_bn254(ctx: &'c Context, bases: &[ark_bn254::G2Affine], precompute: i32) -> Result<Self, Error> {
        let mut packed = Vec::<u64>::with_capacity(bases.len() * 16);
        let mut inf = vec![0u8; (bases.len() + 7) / 8];
        for (i, p) in bases.iter().enumerate() {
            if p.is_zero() {
                inf[i / 8] |= 1 << (i % 8);
                packed.extend_from_slice(&[0u64; 16]);
            } else {
                for c in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1] {
                    packed.extend_from_slice(&c.0 .0);   // x.c0 || x.c1 || y.c0 || y.c1, Montgomery limbs
                }
            }
        }
        Self::upload(ctx, OZL_BN254_G2, &packed, &inf, bases.len(), precompute)
    }

    fn msm_raw<const LIMBS: usize>(&self, scalars: &[BigInteger256]) -> Result<[u64; LIMBS], Error> {
        let n = core::cmp::min(self.n, scalars.len());
        let mut out = [0u64; LIMBS];
        match unsafe { ozl_msm(self.ctx.0, self.handle, scalars.as_ptr() as *const u64, n, out.as_mut_ptr()) } {
            0 => Ok(out),
            _ => Err(Error),
        }
    }

    pub fn msm_g1_bn254(&self, scalars: &[BigInteger256]) -> Result<ark_bn254::G1Projective, Error> {
        let o = self.msm_raw::<12>(scalars)?;
        let fq = |k: usize| ark_bn254::Fq::new(BigInteger256([o[k], o[k + 1], o[k + 2], o[k + 3]]));
        Ok(ark_bn254::G1Projective::new(fq(0), fq(4), fq(8)))
    }

    pub fn msm_g2_bn254(&self, scalars: &[BigInteger256]) -> Result<ark_bn254::G2Projective, Error> {
        let o = self.msm_raw::<24>(scalars)?;
        let fq = |k: usize| ark_bn254::Fq::new(BigInteger256([o[k], o[k + 1], o[k + 2], o[k + 3]]));
        let fq2 = |k: usize| ark_bn254::Fq2::new(fq(k), fq(k + 4));
        Ok(ark_bn254::G2Projective::new(fq2(0), fq2(8), fq2(16)))
    }
}

impl Drop for Bases<'_> {
    fn drop(&mut self) {
        unsafe { ozl_msm_bases_free(self.ctx.0, self.handle) };
    }
}
