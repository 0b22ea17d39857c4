//! `Groth16B200: ProofSystem` -- the plugin a maintainer places next to `plugins/arkworks`
//! (mirror of `plugins/arkworks/src/groth16.rs:399-467`).  UNCOMPILED here (no Rust toolchain in the
//! build image); the same call sequence is what `openzl_b200/groth16.py` drives through ctypes and
//! `tests/test_gpu_groth16.py` checks bit-for-bit against the oracle.
//!
//! * `compile` and `verify` delegate to arkworks on the CPU exactly like the reference
//!   (`groth16.rs:428-443`, `:460-466`); `compile` additionally uploads the five query vectors of the
//!   proving key to the device once -- they are constant across proofs (`groth16.rs:127-129`).
//! * `prove` keeps the reference's contract (`groth16.rs:446-457`: borrows the proving context,
//!   consumes the compiler, draws r and s from the caller's rng first, opaque `Error`) but runs
//!   `ark_groth16::create_proof`'s two hot loops on the GPU: the seven domain transforms of
//!   `R1CStoQAP::witness_map` and the five `VariableBaseMSM::multi_scalar_mul` calls.
//!
//! Shown for BN254 (the curve the reference has Poseidon constants for,
//! `plugins/arkworks/src/poseidon/mod.rs:300-322`); BLS12-381 differs only in the limb counts.

use crate::{ntt_in_place_bn254, Context, Error, OZL_BN254_G1, OZL_BN254_G2};
use ark_bn254::{Bn254, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ec::{AffineCurve, ProjectiveCurve};
use ark_ff::{Field, One, PrimeField, UniformRand, Zero};
use ark_groth16::{Groth16 as ArkGroth16, PreparedVerifyingKey, Proof, ProvingKey};
use ark_poly::{EvaluationDomain, Radix2EvaluationDomain};
use ark_relations::r1cs::{ConstraintSynthesizer, ConstraintSystem, OptimizationGoal};
use ark_snark::SNARK;
use openzl_crypto::constraint::ProofSystem;
use openzl_plugin_arkworks::constraint::R1CS;
use openzl_util::rand::{CryptoRng, RngCore, SizedRng};

/// Device-resident `ProvingKey`: the host copy (for the few single points) plus the five query
/// vectors as bases handles.  `a_query[0]`, `b_g1_query[0]`, `b_g2_query[0]` belong to the constant
/// term and stay on the host, as in ark's `calculate_coeff`.
pub struct ProvingContextB200 {
    pub key: ProvingKey<Bn254>,
    ctx: Context,
    a_query: crate::Bases,      // a_query[1..]
    b_g1_query: crate::Bases,   // b_g1_query[1..]
    b_g2_query: crate::Bases,   // b_g2_query[1..]   (G2)
    h_query: crate::Bases,
    l_query: crate::Bases,
}

pub struct VerifyingContextB200(pub PreparedVerifyingKey<Bn254>);

#[derive(Clone, Copy, Debug, Default, Eq, PartialEq)]
pub struct Groth16B200;

impl ProofSystem for Groth16B200 {
    type Compiler = R1CS<Fr>;
    type PublicParameters = ();
    type ProvingContext = ProvingContextB200;
    type VerifyingContext = VerifyingContextB200;
    type Input = Vec<Fr>;
    type Proof = Proof<Bn254>;
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
        let (key, vk) = ArkGroth16::<Bn254>::circuit_specific_setup(compiler, &mut SizedRng(rng)).map_err(|_| Error)?;
        let pvk = ArkGroth16::<Bn254>::process_vk(&vk).map_err(|_| Error)?;
        let ctx = Context::new(0)?;
        // one shifted copy per Pippenger window (factor 32 >= the window count): a single bucket set
        let full = 32;
        let a_query = crate::Bases::upload_g1_bn254(&ctx, &key.a_query[1..], full)?;
        let b_g1_query = crate::Bases::upload_g1_bn254(&ctx, &key.b_g1_query[1..], full)?;
        let b_g2_query = crate::Bases::upload_g2_bn254(&ctx, &key.b_g2_query[1..], full)?;
        let h_query = crate::Bases::upload_g1_bn254(&ctx, &key.h_query, full)?;
        let l_query = crate::Bases::upload_g1_bn254(&ctx, &key.l_query, full)?;
        Ok((ProvingContextB200 { key, ctx, a_query, b_g1_query, b_g2_query, h_query, l_query }, VerifyingContextB200(pvk)))
    }

    fn prove<R>(context: &Self::ProvingContext, compiler: Self::Compiler, rng: &mut R) -> Result<Self::Proof, Error>
    where
        R: CryptoRng + RngCore + ?Sized,
    {
        // create_random_proof: r, s first, from the caller's rng
        let mut rng = SizedRng(rng);
        let r = Fr::rand(&mut rng);
        let s = Fr::rand(&mut rng);

        // create_proof: synthesize (R1CS::generate_constraints moves the precomputed system in,
        // constraint/mod.rs:186-196), finalize
        let cs = ConstraintSystem::<Fr>::new_ref();
        cs.set_optimization_goal(OptimizationGoal::Constraints);
        compiler.generate_constraints(cs.clone()).map_err(|_| Error)?;
        cs.finalize();

        // R1CStoQAP::witness_map with the seven transforms on the device
        let matrices = cs.to_matrices().ok_or(Error)?;
        let prover = cs.borrow().ok_or(Error)?;
        let num_inputs = prover.instance_assignment.len();
        let num_constraints = cs.num_constraints();
        let full: Vec<Fr> = prover.instance_assignment.iter().chain(prover.witness_assignment.iter()).copied().collect();
        let domain = Radix2EvaluationDomain::<Fr>::new(num_constraints + num_inputs).ok_or(Error)?;
        let n = domain.size();
        let dot = |row: &[(Fr, usize)]| row.iter().fold(Fr::zero(), |acc, (coeff, j)| acc + *coeff * full[*j]);
        let mut a = vec![Fr::zero(); n];
        let mut b = vec![Fr::zero(); n];
        let mut c = vec![Fr::zero(); n];
        for i in 0..num_constraints {
            a[i] = dot(&matrices.a[i]);
            b[i] = dot(&matrices.b[i]);
            c[i] = dot(&matrices.c[i]);
        }
        a[num_constraints..num_constraints + num_inputs].copy_from_slice(&full[..num_inputs]);
        let ctx = &context.ctx;
        for v in [&mut a, &mut b, &mut c] {
            ntt_in_place_bn254(ctx, v, true, false)?;   // ifft_in_place
            ntt_in_place_bn254(ctx, v, false, true)?;   // coset_fft_in_place
        }
        // (a * b - c) / Z on the coset: Z(g w^i) = g^n - 1 is the same for every i
        let z_inv = (Fr::multiplicative_generator().pow([n as u64]) - Fr::one()).inverse().ok_or(Error)?;
        let mut h: Vec<Fr> = a.iter().zip(&b).zip(&c).map(|((a, b), c)| (*a * *b - *c) * z_inv).collect();
        ntt_in_place_bn254(ctx, &mut h, true, true)?;  // coset_ifft_in_place

        // the five MSMs (scalars in canonical form: into_repr)
        let repr = |v: &[Fr]| v.iter().map(|x| x.into_repr()).collect::<Vec<_>>();
        let h_acc: G1Projective = context.h_query.msm_g1_bn254(&repr(&h[..n - 1]))?;
        let aux = repr(&prover.witness_assignment);
        let l_acc: G1Projective = context.l_query.msm_g1_bn254(&aux)?;
        let assignment = [repr(&prover.instance_assignment[1..]), aux].concat();
        let key = &context.key;
        // calculate_coeff(initial, query, vk_param, assignment) = initial + query[0] + MSM(query[1..]) + vk_param
        let coeff_g1 = |initial: G1Projective, q0: &G1Affine, acc: G1Projective, vk: &G1Affine| {
            let mut res = initial;
            res.add_assign_mixed(q0);
            res += &acc;
            res.add_assign_mixed(vk);
            res
        };
        let g_a = coeff_g1(key.delta_g1.mul(r), &key.a_query[0], context.a_query.msm_g1_bn254(&assignment)?, &key.vk.alpha_g1);
        let g1_b = if r.is_zero() {
            G1Projective::zero()
        } else {
            coeff_g1(key.delta_g1.mul(s), &key.b_g1_query[0], context.b_g1_query.msm_g1_bn254(&assignment)?, &key.beta_g1)
        };
        let g2_b: G2Projective = {
            let mut res = key.vk.delta_g2.mul(s);
            res.add_assign_mixed(&key.b_g2_query[0]);
            res += &context.b_g2_query.msm_g2_bn254(&assignment)?;
            res.add_assign_mixed(&key.vk.beta_g2);
            res
        };
        let mut g_c = g_a.mul(s.into_repr());
        g_c += &g1_b.mul(r.into_repr());
        g_c -= &key.delta_g1.mul(r).mul(s.into_repr());
        g_c += &l_acc;
        g_c += &h_acc;
        Ok(Proof { a: g_a.into_affine(), b: g2_b.into_affine(), c: g_c.into_affine() })
    }

    fn verify(context: &Self::VerifyingContext, input: &Self::Input, proof: &Self::Proof) -> Result<bool, Error> {
        // three pairings, milliseconds on the CPU: stays with arkworks (groth16.rs:460-466)
        ArkGroth16::<Bn254>::verify_with_processed_vk(&context.0, input, proof).map_err(|_| Error)
    }
}

// `crate::Bases` (lib.rs) is the curve-generic form of `G1Bases381`: `upload_g1_bn254` / `upload_g2_bn254`
// pack `x.0.0 || y.0.0` (G2: `x.c0 || x.c1 || y.c0 || y.c1`) Montgomery limbs plus the infinity bitset,
// call `ozl_msm_bases_upload(ctx, OZL_BN254_G1 | OZL_BN254_G2, ...)` then `ozl_msm_bases_precompute`;
// `msm_g1_bn254` / `msm_g2_bn254` call `ozl_msm` and rebuild `G1Projective::new(X, Y, Z)` /
// `G2Projective::new(..)` from the Jacobian limbs with `Fq::new(BigInteger256(..))`.
#[allow(dead_code)]
const _CURVE_IDS: (i32, i32) = (OZL_BN254_G1, OZL_BN254_G2);
