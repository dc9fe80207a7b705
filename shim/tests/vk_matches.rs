//! ACCEPTANCE TEST of the drop-in claim (uncompiled in this repository's image: no Rust toolchain).
//!
//! 1. `keygen_vk` (stock halo2, CPU) of the reference's pkcs1v15 circuit and `b2r_rsa_keygen` (GPU) must produce the same
//!    15 fixed-column and 6 permutation commitments for the same `ParamsKZG`  -> pins the circuit layout the GPU
//!    recorder restates (oracle/EXT_ASSUMPTIONS.md A1-A3, B5-B6, C1-C5, C8-C9).
//! 2. a proof made by `b2r_rsa_prove_batch_ex` must be accepted by the STOCK `verify_proof` with the Rust verifying key
//!    -> pins the transcript, the proof order and every prover formula (A3-A5, B1-B4, C6-C7, D1, D3-D8).
//! 3. the CPU proof of the same instance (stock `create_proof`) and the GPU proof verify under the same vk; a GPU proof
//!    of a wrong signature is rejected.
//!
//! The circuit below is the SHA-disabled instantiation of the reference's bench circuit (benches/bench.rs:88-226, macro
//! arguments k = 17, n_bits = 2048, sha2 disabled), written against the reference's PUBLIC chip API because the bench's
//! own struct is private to its file: same `configure` calls in the same order, same three regions in `synthesize`.
use b2rsa_shim::{Gpu, RsaProver};
use halo2_rsa::{
    big_integer::{BigIntConfig, BigIntInstructions, UnassignedInteger},
    RSAChip, RSAConfig, RSAInstructions, RSAPubE, RSAPublicKey, RSASignature,
};
use halo2wrong::curves::bn256::{Bn256, Fr, G1Affine};
use halo2wrong::curves::FieldExt;
use halo2wrong::halo2::{
    circuit::{Layouter, SimpleFloorPlanner, Value},
    plonk::{create_proof, keygen_pk, keygen_vk, verify_proof, Circuit, ConstraintSystem, Error},
    poly::kzg::{
        commitment::{KZGCommitmentScheme, ParamsKZG},
        multiopen::{ProverGWC, VerifierGWC},
        strategy::SingleStrategy,
    },
    transcript::{Blake2bRead, Blake2bWrite, Challenge255, TranscriptReadBuffer, TranscriptWriterBuffer},
};
use maingate::{decompose_big, MainGate, MainGateInstructions, RangeChip, RangeInstructions, RegionCtx};
use num_bigint::BigUint;
use rand::rngs::OsRng;
use rsa::{Hash, PaddingScheme, PublicKeyParts, RsaPrivateKey, RsaPublicKey};
use sha2::{Digest, Sha256};
use std::marker::PhantomData;

const K: u32 = 17;
const BITS: usize = 2048;
const E: u128 = 65537;

#[derive(Clone)]
struct Cfg { rsa: RSAConfig }

struct Pkcs1v15Circuit<F: FieldExt> {
    signature: RSASignature<F>,
    public_key: RSAPublicKey<F>,
    hashed: Vec<u8>, // SHA-256 digest of the message, big-endian bytes as Sha256::digest returns them
    _f: PhantomData<F>,
}
impl<F: FieldExt> Default for Pkcs1v15Circuit<F> {
    fn default() -> Self {
        let nl = BITS / RSAChip::<F>::LIMB_WIDTH;
        Self { signature: RSASignature::without_witness(nl), public_key: RSAPublicKey::without_witness(nl, BigUint::from(E)), hashed: vec![0; 32], _f: PhantomData }
    }
}
impl<F: FieldExt> Circuit<F> for Pkcs1v15Circuit<F> {
    type Config = Cfg;
    type FloorPlanner = SimpleFloorPlanner;
    fn without_witnesses(&self) -> Self { Self::default() }
    fn configure(meta: &mut ConstraintSystem<F>) -> Cfg {
        let main_gate = MainGate::<F>::configure(meta);
        let (comp, over) = RSAChip::<F>::compute_range_lens(BITS / RSAChip::<F>::LIMB_WIDTH);
        let range = RangeChip::<F>::configure(meta, &main_gate, comp, over);
        Cfg { rsa: RSAConfig::new(BigIntConfig::new(range, main_gate)) }
    }
    fn synthesize(&self, cfg: Cfg, mut layouter: impl Layouter<F>) -> Result<(), Error> {
        let rsa_chip = RSAChip::<F>::new(cfg.rsa, BITS, 5);
        let bigint_chip = rsa_chip.bigint_chip();
        let main_gate = rsa_chip.main_gate();
        bigint_chip.range_chip().load_table(&mut layouter)?;
        let (public_key, signature) = layouter.assign_region(|| "keys", |region| {
            let ctx = &mut RegionCtx::new(region, 0);
            let sign = rsa_chip.assign_signature(ctx, self.signature.clone())?;
            let pk = rsa_chip.assign_public_key(ctx, self.public_key.clone())?;
            Ok((pk, sign))
        })?;
        let is_valid = layouter.assign_region(|| "verify", |region| {
            let ctx = &mut RegionCtx::new(region, 0);
            let mut h = self.hashed.clone();
            h.reverse();
            let limbs = decompose_big::<F>(BigUint::from_bytes_le(&h), 4, RSAChip::<F>::LIMB_WIDTH);
            let hashed = bigint_chip.assign_integer(ctx, UnassignedInteger::from(limbs))?;
            rsa_chip.verify_pkcs1v15_signature(ctx, &public_key, &hashed, &signature)
        })?;
        layouter.assign_region(|| "assert", |region| {
            let ctx = &mut RegionCtx::new(region, 0);
            main_gate.assert_one(ctx, &is_valid)
        })
    }
}

fn limbs64(x: &BigUint, n: usize) -> Vec<u64> {
    let mut v = x.to_u64_digits();
    v.resize(n, 0);
    v
}

#[test]
fn vk_matches_and_stock_verifier_accepts_gpu_proofs() {
    let params = ParamsKZG::<Bn256>::setup(K, OsRng);
    let empty = Pkcs1v15Circuit::<Fr>::default();
    let vk = keygen_vk(&params, &empty).expect("keygen_vk");
    let pk = keygen_pk(&params, vk.clone(), &empty).expect("keygen_pk");

    // ---- 1. verifying-key commitments
    let gpu = Gpu::new(0).expect("no sm_100 device: libb2rsa has no CPU fallback");
    let mut prover = RsaProver::new(&gpu, &params, BITS as u32, &BigUint::from(E)).expect("b2r_rsa_keygen");
    let (fixed, sigma) = prover.vk_commitments().unwrap();
    let rust_fixed: &Vec<G1Affine> = vk.fixed_commitments();
    let rust_sigma: &Vec<G1Affine> = vk.permutation().commitments();
    assert_eq!(rust_fixed.len(), fixed.len(), "number of fixed columns (EXT_ASSUMPTIONS C1/C5)");
    for (i, (a, b)) in rust_fixed.iter().zip(fixed.iter()).enumerate() {
        assert_eq!(a, b, "fixed commitment {} differs: 0-8 = MainGate columns (C1/C3), 9-14 = RangeChip columns (C4/C5)", i);
    }
    assert_eq!(rust_sigma.len(), sigma.len(), "number of permutation columns (C2)");
    for (i, (a, b)) in rust_sigma.iter().zip(sigma.iter()).enumerate() {
        assert_eq!(a, b, "permutation commitment {} differs (C2 column order / C9 cycle construction / C3 copy constraints)", i);
    }

    // ---- 2. a GPU proof under the stock verifier
    prover.bind_vk(&vk).unwrap(); // the real transcript_repr (D2)
    let mut rng = rand::thread_rng();
    let sk = RsaPrivateKey::new(&mut rng, BITS).unwrap();
    let pubk = RsaPublicKey::from(&sk);
    let msg = b"b2rsa acceptance test";
    let digest = Sha256::digest(msg).to_vec();
    let mut sig = sk.sign(PaddingScheme::PKCS1v15Sign { hash: Some(Hash::SHA2_256) }, &digest).unwrap();
    sig.reverse();
    let sig_big = BigUint::from_bytes_le(&sig);
    let n_big = BigUint::from_bytes_le(&pubk.n().to_bytes_le());
    let nl = BITS / 64;
    let mut h_le = digest.clone();
    h_le.reverse();
    let hashed_big = BigUint::from_bytes_le(&h_le);
    let mut seed = [0u8; 32];
    rand::RngCore::fill_bytes(&mut rng, &mut seed);
    let (proofs, status) = prover.prove_batch(&limbs64(&n_big, nl), &limbs64(&sig_big, nl), &limbs64(&hashed_big, 4), &seed).unwrap();
    assert_eq!(status, vec![1]);
    let verify = |proof: &[u8]| {
        let strategy = SingleStrategy::new(&params);
        let mut tr = Blake2bRead::<_, _, Challenge255<_>>::init(proof);
        verify_proof::<KZGCommitmentScheme<Bn256>, VerifierGWC<_>, _, _, _>(&params, &vk, strategy, &[&[&[]]], &mut tr).is_ok()
    };
    assert!(verify(&proofs[..prover.proof_bytes]), "stock verify_proof rejected the GPU proof");

    // ---- 3. the CPU prover on the same instance, and a wrong signature on the GPU
    let limb_width = RSAChip::<Fr>::LIMB_WIDTH;
    let circuit = Pkcs1v15Circuit::<Fr> {
        signature: RSASignature::new(Value::known(sig_big.clone())),
        public_key: RSAPublicKey::new(Value::known(n_big.clone()), RSAPubE::Fix(BigUint::from(E))),
        hashed: digest.clone(),
        _f: PhantomData,
    };
    let _ = limb_width;
    let cpu_proof = {
        let mut tr = Blake2bWrite::<_, G1Affine, Challenge255<_>>::init(vec![]);
        create_proof::<KZGCommitmentScheme<_>, ProverGWC<_>, _, _, _, _>(&params, &pk, &[circuit], &[&[&[]]], OsRng, &mut tr).unwrap();
        tr.finalize()
    };
    assert_eq!(cpu_proof.len(), prover.proof_bytes, "proof length (D3/D4)");
    assert!(verify(&cpu_proof));
    let mut bad_hash = limbs64(&hashed_big, 4);
    bad_hash[0] ^= 1;
    let (bad, st) = prover.prove_batch(&limbs64(&n_big, nl), &limbs64(&sig_big, nl), &bad_hash, &seed).unwrap();
    assert_eq!(st, vec![0]);
    assert!(!verify(&bad[..prover.proof_bytes]));
}

/// path (b) alone: the two functions halo2 would call through patches/halo2_proofs_arithmetic.patch
#[test]
fn best_multiexp_and_best_fft_match_halo2() {
    use halo2wrong::halo2::arithmetic::{best_fft as cpu_fft, best_multiexp as cpu_msm, Field};
    use halo2wrong::halo2::poly::commitment::ParamsProver;
    let params = ParamsKZG::<Bn256>::setup(12, OsRng);
    let g: Vec<G1Affine> = params.get_g().to_vec();
    let coeffs: Vec<Fr> = (0..g.len()).map(|_| Fr::random(OsRng)).collect();
    use halo2wrong::curves::group::Curve;
    assert_eq!(b2rsa_shim::best_multiexp(&coeffs, &g).to_affine(), cpu_msm(&coeffs, &g).to_affine());
    let omega = Fr::root_of_unity().pow_vartime(&[1u64 << (Fr::S - 12)]);
    let (mut a, mut b) = (coeffs.clone(), coeffs);
    b2rsa_shim::best_fft(&mut a, omega, 12);
    cpu_fft(&mut b, omega, 12);
    assert_eq!(a, b);
}
