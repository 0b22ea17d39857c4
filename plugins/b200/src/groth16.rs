//! `Groth16B200<E>: ProofSystem` -- the plugin a maintainer places next to `plugins/arkworks`
//! (mirror of `plugins/arkworks/src/groth16.rs:399-467`, generic over the pairing like
//! `Groth16<E>(PhantomData<E>)` there).  UNCOMPILED here (no Rust toolchain in the build image); the
//! identical call sequence is replayed from plain C with pageable buffers by
//! `examples/ozl_groth16_replay.c` (`tests/test_c_replay.py`) and driven through ctypes by
//! `openzl_b200/groth16.py` (`tests/test_gpu_groth16.py`, bit-for-bit against the oracle).
//!
//! * `compile` runs arkworks' circuit-specific setup on the CPU exactly like the reference
//!   (`groth16.rs:428-443`), then moves the proving key to the device ONCE: the five query vectors as
//!   MSM bases (`ozl_msm_bases_upload` + one shifted copy per Pippenger window), the R1CS matrices in CSR
//!   form and alpha/beta/delta (`ozl_groth16_pk_create`).  They are constant across proofs
//!   (`ProvingContext<E>(pub ProvingKey<E>)`, `groth16.rs:127-129`).
//! * `prove` keeps the reference's contract (`groth16.rs:446-457`: borrows the proving context, consumes
//!   the compiler, draws r and s from the caller's rng FIRST, opaque `Error`) and makes ONE device call,
//!   `ozl_groth16_prove`: the full assignment goes in, the device runs `R1CStoQAP::witness_map` (three
//!   sparse mat-vecs, seven domain transforms, the pointwise quotient), the five
//!   `VariableBaseMSM::multi_scalar_mul`s on three streams and the proof assembly, and three affine points
//!   come back.  This is the path `bench.py` times (25 ms per proof at 2^20 constraints).
//! * `verify` stays with arkworks on the CPU (`groth16.rs:459-466`): three pairings, milliseconds.
//!
//! Threading: the reference's `Groth16<E>` has no state, so concurrent `prove` calls sharing one
//! `&ProvingContext` are legal.  A device context is thread-compatible, not thread-safe, hence the
//! `Mutex` around it; provers that want several proofs in flight on one GPU compile one context each.

use crate::ffi::{
    ozl_ctx_create, ozl_ctx_destroy, ozl_groth16_pk_create, ozl_groth16_pk_destroy, ozl_groth16_prove, ozl_msm_bases_free,
    ozl_msm_bases_precompute, ozl_msm_bases_upload,
};
use crate::{Error, OzlCsr, OzlCtx};
use ark_ec::{AffineCurve, PairingEngine};
use ark_ff::{PrimeField, UniformRand, Zero};
use ark_groth16::{Groth16 as ArkGroth16, PreparedVerifyingKey, Proof, ProvingKey};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, ConstraintSystemRef, OptimizationGoal, SynthesisMode};
use ark_snark::SNARK;
use core::marker::PhantomData;
use openzl_crypto::constraint::ProofSystem;
use openzl_plugin_arkworks::constraint::R1CS;
use openzl_util::rand::{CryptoRng, RngCore, SizedRng};
use std::collections::HashMap;
use std::os::raw::c_int;
use std::sync::Mutex;

/// What the shim needs to know about a pairing: the C ABI's identifiers and how arkworks lays its
/// points out in memory.  Implemented below for BN254 and BLS12-381, the two families the plugin
/// enables (`plugins/arkworks/Cargo.toml:116-117`).
pub trait B200Pairing: PairingEngine {
    /// `ozl_pairing`.
    const PAIRING: c_int;
    /// `ozl_curve` of G1 / G2.
    const G1: c_int;
    const G2: c_int;
    /// u64 limbs per Fq element (4 for BN254, 6 for BLS12-381); a G1 point is 2, a G2 point 4 of them.
    const FQ_LIMBS: usize;
    /// `x || y` Montgomery limbs exactly as `Fp::0.0` holds them (zeros for the point at infinity).
    fn pack_g1(p: &Self::G1Affine, out: &mut Vec<u64>);
    /// `x.c0 || x.c1 || y.c0 || y.c1`.
    fn pack_g2(p: &Self::G2Affine, out: &mut Vec<u64>);
    fn unpack_g1(limbs: &[u64]) -> Self::G1Affine;
    fn unpack_g2(limbs: &[u64]) -> Self::G2Affine;
    /// The Montgomery limbs of a scalar-field element (`Fp256::0.0`): what `ozl_groth16_prove` takes for z.
    fn fr_mont(x: &Self::Fr) -> [u64; 4];
}

macro_rules! impl_b200_pairing {
    ($engine:ty, $module:ident, $big:ident, $limbs:expr, $pairing:expr, $g1:expr, $g2:expr) => {
        impl B200Pairing for $engine {
            const PAIRING: c_int = $pairing;
            const G1: c_int = $g1;
            const G2: c_int = $g2;
            const FQ_LIMBS: usize = $limbs;
            fn pack_g1(p: &Self::G1Affine, out: &mut Vec<u64>) {
                if p.is_zero() {
                    out.extend(core::iter::repeat(0u64).take(2 * $limbs));
                } else {
                    out.extend_from_slice(&p.x.0 .0);
                    out.extend_from_slice(&p.y.0 .0);
                }
            }
            fn pack_g2(p: &Self::G2Affine, out: &mut Vec<u64>) {
                if p.is_zero() {
                    out.extend(core::iter::repeat(0u64).take(4 * $limbs));
                } else {
                    for c in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1] {
                        out.extend_from_slice(&c.0 .0);
                    }
                }
            }
            fn unpack_g1(l: &[u64]) -> Self::G1Affine {
                let fq = |k: usize| {
                    let mut w = [0u64; $limbs];
                    w.copy_from_slice(&l[k..k + $limbs]);
                    $module::Fq::new(ark_ff::$big(w)) // Fp::new takes the Montgomery representation
                };
                let zero = l.iter().all(|v| *v == 0);
                $module::G1Affine::new(fq(0), fq($limbs), zero)
            }
            fn unpack_g2(l: &[u64]) -> Self::G2Affine {
                let fq = |k: usize| {
                    let mut w = [0u64; $limbs];
                    w.copy_from_slice(&l[k..k + $limbs]);
                    $module::Fq::new(ark_ff::$big(w))
                };
                let fq2 = |k: usize| $module::Fq2::new(fq(k), fq(k + $limbs));
                let zero = l.iter().all(|v| *v == 0);
                $module::G2Affine::new(fq2(0), fq2(2 * $limbs), zero)
            }
            fn fr_mont(x: &Self::Fr) -> [u64; 4] {
                x.0 .0
            }
        }
    };
}

impl_b200_pairing!(ark_bn254::Bn254, ark_bn254, BigInteger256, 4, 0, crate::OZL_BN254_G1, crate::OZL_BN254_G2);
impl_b200_pairing!(ark_bls12_381::Bls12_381, ark_bls12_381, BigInteger384, 6, 1, crate::OZL_BLS12_381_G1, crate::OZL_BLS12_381_G2);

/// The device half of a proving key: a context and the pk resident in it.  The pk owns its five bases
/// handles (`ozl_groth16_pk_create` takes them over), so there is exactly one thing to release and it
/// is released BEFORE the context (see `Drop`): no handle outlives the context it belongs to.
struct DeviceKey {
    ctx: *mut OzlCtx,
    pk: u32,
}

// The raw pointer is only ever handed to libozl_b200 while the surrounding Mutex is held; the library
// keeps no thread-local state, so the key may move between threads.
unsafe impl Send for DeviceKey {}

impl Drop for DeviceKey {
    fn drop(&mut self) {
        unsafe {
            ozl_groth16_pk_destroy(self.ctx, self.pk);
            ozl_ctx_destroy(self.ctx); // would also release the pk: the registry lives in the context
        }
    }
}

/// `ProvingContext<E>` with the key also resident on the device.  The host copy stays because the
/// reference's codecs (`groth16.rs:142-179`) serialize it.
pub struct ProvingContextB200<E: B200Pairing> {
    pub key: ProvingKey<E>,
    num_instance: usize,
    num_vars: usize,
    device: Mutex<DeviceKey>,
}

pub struct VerifyingContextB200<E: PairingEngine>(pub PreparedVerifyingKey<E>);

#[derive(Clone, Copy, Debug, Default, Eq, PartialEq)]
pub struct Groth16B200<E>(PhantomData<E>);

/// CSR form of one constraint matrix over a coefficient table shared by A, B and C (gadget-built
/// systems have a handful of distinct constants).
fn to_csr<F: PrimeField>(rows: &[Vec<(F, usize)>], table: &mut Vec<F>, index: &mut HashMap<F, u32>) -> (Vec<u32>, Vec<u32>, Vec<u32>) {
    let mut row_ptr = Vec::with_capacity(rows.len() + 1);
    let (mut col, mut cidx) = (Vec::new(), Vec::new());
    row_ptr.push(0u32);
    for row in rows {
        for (coeff, j) in row {
            let k = *index.entry(*coeff).or_insert_with(|| {
                table.push(*coeff);
                (table.len() - 1) as u32
            });
            col.push(*j as u32);
            cidx.push(k);
        }
        row_ptr.push(col.len() as u32);
    }
    (row_ptr, col, cidx)
}

fn upload<E: B200Pairing>(ctx: *mut OzlCtx, curve: c_int, packed: &[u64], inf: &[u8], n: usize) -> Result<u32, Error> {
    let mut handle = 0u32;
    if unsafe { ozl_msm_bases_upload(ctx, curve, packed.as_ptr(), inf.as_ptr(), n, &mut handle) } != 0 {
        return Err(Error);
    }
    // factor 32 >= the number of Pippenger windows: one shifted copy per window, a single bucket set
    if unsafe { ozl_msm_bases_precompute(ctx, handle, 32) } != 0 {
        unsafe { ozl_msm_bases_free(ctx, handle) };
        return Err(Error);
    }
    Ok(handle)
}

fn upload_g1<E: B200Pairing>(ctx: *mut OzlCtx, pts: &[E::G1Affine]) -> Result<u32, Error> {
    let mut packed = Vec::with_capacity(pts.len() * 2 * E::FQ_LIMBS);
    let mut inf = vec![0u8; (pts.len() + 7) / 8 + 8];
    for (i, p) in pts.iter().enumerate() {
        if p.is_zero() {
            inf[i / 8] |= 1 << (i % 8);
        }
        E::pack_g1(p, &mut packed);
    }
    upload::<E>(ctx, E::G1, &packed, &inf, pts.len())
}

fn upload_g2<E: B200Pairing>(ctx: *mut OzlCtx, pts: &[E::G2Affine]) -> Result<u32, Error> {
    let mut packed = Vec::with_capacity(pts.len() * 4 * E::FQ_LIMBS);
    let mut inf = vec![0u8; (pts.len() + 7) / 8 + 8];
    for (i, p) in pts.iter().enumerate() {
        if p.is_zero() {
            inf[i / 8] |= 1 << (i % 8);
        }
        E::pack_g2(p, &mut packed);
    }
    upload::<E>(ctx, E::G2, &packed, &inf, pts.len())
}

impl<E> ProofSystem for Groth16B200<E>
where
    E: B200Pairing,
{
    type Compiler = R1CS<E::Fr>;
    type PublicParameters = ();
    type ProvingContext = ProvingContextB200<E>;
    type VerifyingContext = VerifyingContextB200<E>;
    type Input = Vec<E::Fr>;
    type Proof = Proof<E>;
    type Error = Error;

    fn context_compiler() -> Self::Compiler {
        Self::Compiler::for_contexts()
    }

    fn proof_compiler() -> Self::Compiler {
        Self::Compiler::for_proofs()
    }

    fn compile<R>(_: &(), compiler: Self::Compiler, rng: &mut R) -> Result<(Self::ProvingContext, Self::VerifyingContext), Error>
    where
        R: CryptoRng + RngCore + ?Sized,
    {
        // The matrices of the circuit: synthesize once in setup mode (the compiler hands its precomputed
        // system over, constraint/mod.rs:186-196), keep A, B, C, and give the same system to arkworks' setup.
        let cs: ConstraintSystemRef<E::Fr> = ConstraintSystem::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        cs.set_mode(SynthesisMode::Setup);
        compiler.generate_constraints(cs.clone()).map_err(|_| Error)?;
        cs.finalize();
        let matrices = cs.to_matrices().ok_or(Error)?;
        let (num_constraints, num_instance) = (cs.num_constraints(), cs.num_instance_variables());
        let num_vars = num_instance + cs.num_witness_variables();
        let (key, vk) = ArkGroth16::<E>::circuit_specific_setup(R1CS::new_unchecked(cs), &mut SizedRng(rng)).map_err(|_| Error)?;
        let pvk = ArkGroth16::<E>::process_vk(&vk).map_err(|_| Error)?;

        let mut table = Vec::new();
        let mut index = HashMap::new();
        let a = to_csr(&matrices.a, &mut table, &mut index);
        let b = to_csr(&matrices.b, &mut table, &mut index);
        let c = to_csr(&matrices.c, &mut table, &mut index);
        let coef: Vec<u64> = table.iter().flat_map(|x| E::fr_mont(x)).collect();
        let csr = |m: &(Vec<u32>, Vec<u32>, Vec<u32>)| OzlCsr { n_rows: num_constraints as u32, row_ptr: m.0.as_ptr(), col_idx: m.1.as_ptr(), coef_idx: m.2.as_ptr() };

        let mut ctx = core::ptr::null_mut();
        if unsafe { ozl_ctx_create(0, &mut ctx) } != 0 {
            return Err(Error);
        }
        // from here on every failure destroys the context, which releases whatever was uploaded into it
        let built = (|| -> Result<u32, Error> {
            // the whole vectors: index 0 belongs to the constant 1, which z carries as its first entry
            // (ark's calculate_coeff adds query[0] separately; the sum is the same)
            let a_q = upload_g1::<E>(ctx, &key.a_query)?;
            let b1_q = upload_g1::<E>(ctx, &key.b_g1_query)?;
            let b2_q = upload_g2::<E>(ctx, &key.b_g2_query)?;
            let h_q = upload_g1::<E>(ctx, &key.h_query)?;
            let l_q = upload_g1::<E>(ctx, &key.l_query)?;
            let g1 = |p: &E::G1Affine| { let mut v = Vec::new(); E::pack_g1(p, &mut v); v };
            let g2 = |p: &E::G2Affine| { let mut v = Vec::new(); E::pack_g2(p, &mut v); v };
            let (alpha1, beta1, delta1) = (g1(&key.vk.alpha_g1), g1(&key.beta_g1), g1(&key.delta_g1));
            let (beta2, delta2) = (g2(&key.vk.beta_g2), g2(&key.vk.delta_g2));
            let mut pk = 0u32;
            let rc = unsafe {
                ozl_groth16_pk_create(
                    ctx, E::PAIRING, num_constraints as u32, num_instance as u32, num_vars as u32,
                    &csr(&a), &csr(&b), &csr(&c), coef.as_ptr(), table.len() as u32,
                    a_q, b1_q, b2_q, h_q, l_q,
                    alpha1.as_ptr(), beta1.as_ptr(), delta1.as_ptr(), beta2.as_ptr(), delta2.as_ptr(), &mut pk,
                )
            };
            if rc != 0 { Err(Error) } else { Ok(pk) }
        })();
        let pk = match built {
            Ok(pk) => pk,
            Err(e) => {
                unsafe { ozl_ctx_destroy(ctx) };
                return Err(e);
            }
        };
        Ok((
            ProvingContextB200 { key, num_instance, num_vars, device: Mutex::new(DeviceKey { ctx, pk }) },
            VerifyingContextB200(pvk),
        ))
    }

    fn prove<R>(context: &Self::ProvingContext, compiler: Self::Compiler, rng: &mut R) -> Result<Self::Proof, Error>
    where
        R: CryptoRng + RngCore + ?Sized,
    {
        // create_random_proof: r, s first, from the caller's rng, two Fr::rand draws
        let mut rng = SizedRng(rng);
        let r = E::Fr::rand(&mut rng);
        let s = E::Fr::rand(&mut rng);

        // create_proof's synthesis step: the compiler moves its precomputed system (with assignments) in
        let cs = ConstraintSystem::<E::Fr>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        compiler.generate_constraints(cs.clone()).map_err(|_| Error)?;
        cs.finalize();
        let prover = cs.borrow().ok_or(Error)?;
        if prover.instance_assignment.len() != context.num_instance
            || prover.instance_assignment.len() + prover.witness_assignment.len() != context.num_vars
        {
            return Err(Error);
        }
        // z = (1, instance.., witness..) as Montgomery limbs: the one buffer that crosses PCIe per proof
        let z: Vec<u64> = prover.instance_assignment.iter().chain(prover.witness_assignment.iter()).flat_map(|x| E::fr_mont(x)).collect();
        let (rr, ss) = (r.into_repr(), s.into_repr()); // canonical, like the MSM scalars
        let (l1, l2) = (2 * E::FQ_LIMBS, 4 * E::FQ_LIMBS);
        let (mut a, mut b, mut c) = (vec![0u64; l1], vec![0u64; l2], vec![0u64; l1]);
        {
            let dev = context.device.lock().map_err(|_| Error)?;
            let rc = unsafe {
                ozl_groth16_prove(dev.ctx, dev.pk, z.as_ptr(), rr.as_ref().as_ptr(), ss.as_ref().as_ptr(),
                                  a.as_mut_ptr(), b.as_mut_ptr(), c.as_mut_ptr(), core::ptr::null_mut())
            };
            if rc != 0 {
                return Err(Error);
            }
        }
        Ok(Proof { a: E::unpack_g1(&a), b: E::unpack_g2(&b), c: E::unpack_g1(&c) })
    }

    fn verify(context: &Self::VerifyingContext, input: &Self::Input, proof: &Self::Proof) -> Result<bool, Error> {
        // three pairings, milliseconds on the CPU: stays with arkworks (groth16.rs:459-466)
        ArkGroth16::<E>::verify_with_processed_vk(&context.0, input, proof).map_err(|_| Error)
    }
}
