//! SOURCE ONLY -- NOT COMPILED IN THIS REPOSITORY'S IMAGE (no cargo / rustc there).
//!
//! Drop-in for `crypto_ops::verify_merkle_proof` (crypto-ops/src/lib.rs:8-23) plus the batched
//! `verify_merkle_proofs`, over the C ABI of `include/mptv.h`.  Inputs are borsh-serialised and handed to
//! `mptv_flatten_borsh`, which lays them out as the CSR arena in page-locked memory (multi-threaded, buffers
//! recycled through the handle), so `mptv_verify_batch` streams them at PCIe speed.
use alloy_primitives::B256;
use crypto_ops::types::{MerkleProofInput, StorageProofInput};
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct MptvBatch {
    pub node_bytes: *const u8,
    pub node_bytes_len: u64,
    pub node_off: *const u64,
    pub node_len: *const u32,
    pub n_nodes: u64,
    pub proof_first: *const u32,
    pub n_proofs: u64,
    pub roots: *const u8,
    pub key_bytes: *const u8,
    pub key_off: *const u32,
    pub root_from_proof: *const i32,
}
#[repr(C)]
pub struct MptvResult {
    pub status: *mut u8,
    pub value_off: *mut u64,
    pub value_len: *mut u32,
}
// the layout include/mptv.h pins (MPTV_ABI_PIN): a mismatch fails the build on either side
const _: () = assert!(std::mem::size_of::<MptvBatch>() == 88 && std::mem::size_of::<MptvResult>() == 24);
pub enum MptvCtx {}
pub enum MptvHostBatch {}

extern "C" {
    fn mptv_create(device_ids: *const c_int, n_devices: c_int, out: *mut *mut MptvCtx) -> c_int;
    fn mptv_destroy(ctx: *mut MptvCtx);
    fn mptv_strerror(err: c_int) -> *const c_char;
    fn mptv_verify_batch(ctx: *mut MptvCtx, input: *const MptvBatch, out: *mut MptvResult) -> c_int;
    fn mptv_verify_batch_hashed_keys(ctx: *mut MptvCtx, input: *const MptvBatch, hash_key: *const u8, out: *mut MptvResult) -> c_int;
    fn mptv_verify_borsh(ctx: *mut MptvCtx, blobs: *const u8, blob_off: *const u64, n: u64, n_threads: c_int, out: *mut MptvResult) -> c_int;
    fn mptv_verify_storage_borsh(ctx: *mut MptvCtx, blobs: *const u8, blob_off: *const u64, n_inputs: u64, n_threads: c_int,
                                 proof_first: *mut u64, input_status: *mut u8, results_cap: u64, out: *mut MptvResult) -> c_int;
    fn mptv_flatten_borsh(blobs: *const u8, blob_off: *const u64, n: u64, n_threads: c_int, pinned: c_int, out: *mut *mut MptvHostBatch) -> c_int;
    fn mptv_host_batch_view(hb: *const MptvHostBatch) -> *const MptvBatch;
    fn mptv_host_batch_bad_root(hb: *const MptvHostBatch) -> *const u8;
    fn mptv_host_batch_free(hb: *mut MptvHostBatch);
}

/// What the reference would have panicked with (MPTV_ST_* of mptv.h).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum VerifyError {
    InvalidStateRoot,   // lib.rs:14  "Invalid merkle proof: ..."
    RootNotCanonical,   // lib.rs:19  assert_eq!(root_hash, trie.root_hash())
    InvalidProof,       // lib.rs:21  "Failed to verify Merkle Proof: InvalidProof"
    KeyNotFound,        // lib.rs:22  "Key does not exist!"
    PanicOther,         // raw panics inside eth_trie
    BadRootLen,         // root_hash.len() != 32 (the guests' try_into().unwrap())
    DependencyFailed,   // storage proof whose account proof failed
}

impl VerifyError {
    fn from_status(s: u8) -> Option<Self> {
        match s {
            0 => None,
            1 => Some(Self::InvalidStateRoot),
            2 => Some(Self::RootNotCanonical),
            3 => Some(Self::InvalidProof),
            4 => Some(Self::KeyNotFound),
            6 => Some(Self::BadRootLen),
            7 => Some(Self::DependencyFailed),
            _ => Some(Self::PanicOther),
        }
    }
    fn panic_message(self) -> &'static str {
        match self {
            Self::InvalidStateRoot => "Invalid merkle proof: InvalidStateRoot",
            Self::RootNotCanonical => "assertion `left == right` failed",
            Self::InvalidProof => "Failed to verify Merkle Proof: InvalidProof",
            Self::KeyNotFound => "Key does not exist!",
            Self::BadRootLen => "called `Result::unwrap()` on an `Err` value: TryFromSliceError",
            Self::DependencyFailed => "account proof rejected or not an Account RLP",
            Self::PanicOther => "panicked inside eth_trie (invalid data / index out of bounds)",
        }
    }
}

/// Owns an `mptv_ctx` (device memory + streams) and the recycled pinned flatten buffers.
/// One context is not thread-safe; separate contexts are.  There is no CPU fallback: `new` fails without a B200.
pub struct Verifier {
    ctx: *mut MptvCtx,
    flat: *mut MptvHostBatch,
}
unsafe impl Send for Verifier {}

impl Verifier {
    pub fn new() -> Result<Self, String> {
        let mut ctx: *mut MptvCtx = std::ptr::null_mut();
        let rc = unsafe { mptv_create(std::ptr::null(), 0, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(mptv_strerror(rc)) }.to_string_lossy().into_owned();
            return Err(msg);
        }
        Ok(Self { ctx, flat: std::ptr::null_mut() })
    }

    /// The batched entry of the north star: `verify_merkle_proofs(&[MerkleProofInput])`.
    pub fn verify_merkle_proofs(&mut self, inputs: &[MerkleProofInput]) -> Vec<Result<Vec<u8>, VerifyError>> {
        self.run(inputs, None, None)
    }

    /// The same for inputs that already exist as borsh blobs (the prover's input file, prover/src/bin/main.rs:41,67):
    /// blob i = `blobs[off[i]..off[i + 1]]`.  One pipelined call: a chunk is flattened while the previous ones are
    /// copied and verified; every value comes back as a slice of `blobs`.
    pub fn verify_borsh_blobs(&mut self, blobs: &[u8], off: &[u64]) -> Vec<Result<Vec<u8>, VerifyError>> {
        let n = off.len().saturating_sub(1);
        if n == 0 {
            return Vec::new();
        }
        let (mut status, mut voff, mut vlen) = (vec![0u8; n], vec![0u64; n], vec![0u32; n]);
        let mut res = MptvResult { status: status.as_mut_ptr(), value_off: voff.as_mut_ptr(), value_len: vlen.as_mut_ptr() };
        let rc = unsafe { mptv_verify_borsh(self.ctx, blobs.as_ptr(), off.as_ptr(), n as u64, 0, &mut res) };
        assert_eq!(rc, 0, "mptv_verify_borsh failed (malformed borsh?)");
        (0..n)
            .map(|p| match VerifyError::from_status(status[p]) {
                None => Ok(blobs[voff[p] as usize..voff[p] as usize + vlen[p] as usize].to_vec()),
                Some(e) => Err(e),
            })
            .collect()
    }

    /// The storage guest (storage-circuit/src/main.rs:6-31) over inputs that exist as borsh(StorageProofInput) bytes:
    /// per input the committed storage values, or what the guest would have panicked with.  `n_proofs` = the sum over the
    /// inputs of 1 + min(storage_proofs.len(), storage_keys.len()) when the caller knows it (else 0: one sizing call more).
    pub fn verify_storage_borsh_blobs(&mut self, blobs: &[u8], off: &[u64], n_proofs: usize) -> Vec<Result<Vec<Vec<u8>>, VerifyError>> {
        let n = off.len().saturating_sub(1);
        if n == 0 {
            return Vec::new();
        }
        let mut first = vec![0u64; n + 1];
        let mut ist = vec![0u8; n];
        let mut cap = n_proofs;
        loop {
            let (mut status, mut voff, mut vlen) = (vec![0u8; cap], vec![0u64; cap], vec![0u32; cap]);
            let mut res = MptvResult { status: status.as_mut_ptr(), value_off: voff.as_mut_ptr(), value_len: vlen.as_mut_ptr() };
            let rc = unsafe {
                mptv_verify_storage_borsh(self.ctx, blobs.as_ptr(), off.as_ptr(), n as u64, 0, first.as_mut_ptr(), ist.as_mut_ptr(),
                                          cap as u64, &mut res)
            };
            if rc == -4 && first[n] as usize > cap {
                cap = first[n] as usize; // MPTV_ERR_NOMEM: proof_first holds the layout, nothing was verified
                continue;
            }
            assert_eq!(rc, 0, "mptv_verify_storage_borsh failed (malformed borsh?)");
            return (0..n)
                .map(|i| match VerifyError::from_status(ist[i]) {
                    Some(e) => Err(e),
                    None => Ok((first[i] as usize + 1..first[i + 1] as usize)
                        .map(|p| blobs[voff[p] as usize..voff[p] as usize + vlen[p] as usize].to_vec())
                        .collect()),
                })
                .collect();
        }
    }

    /// `root_from_proof[p] = d >= 0`: proof p is verified under the storage_root of the account returned by the
    /// earlier proof d; `hash_key[p] != 0`: proof p is looked up under keccak256(key) (computed on the device).
    fn run(&mut self, inputs: &[MerkleProofInput], root_from_proof: Option<&[i32]>, hash_key: Option<&[u8]>)
        -> Vec<Result<Vec<u8>, VerifyError>> {
        if inputs.is_empty() {
            return Vec::new();
        }
        let mut blobs: Vec<u8> = Vec::new();
        let mut off: Vec<u64> = vec![0];
        for inp in inputs {
            blobs.extend_from_slice(&borsh::to_vec(inp).expect("borsh"));
            off.push(blobs.len() as u64);
        }
        let rc = unsafe { mptv_flatten_borsh(blobs.as_ptr(), off.as_ptr(), inputs.len() as u64, 0, 1, &mut self.flat) };
        assert_eq!(rc, 0, "mptv_flatten_borsh failed");
        let mut b = unsafe { std::ptr::read(mptv_host_batch_view(self.flat)) };
        if let Some(r) = root_from_proof {
            b.root_from_proof = r.as_ptr();
        }
        let n = inputs.len();
        let (mut status, mut voff, mut vlen) = (vec![0u8; n], vec![0u64; n], vec![0u32; n]);
        let mut res = MptvResult { status: status.as_mut_ptr(), value_off: voff.as_mut_ptr(), value_len: vlen.as_mut_ptr() };
        let rc = unsafe {
            match hash_key {
                Some(h) => mptv_verify_batch_hashed_keys(self.ctx, &b, h.as_ptr(), &mut res),
                None => mptv_verify_batch(self.ctx, &b, &mut res),
            }
        };
        assert_eq!(rc, 0, "mptv_verify_batch failed");
        let bad = unsafe { std::slice::from_raw_parts(mptv_host_batch_bad_root(self.flat), n) };
        let arena = unsafe { std::slice::from_raw_parts(b.node_bytes, b.node_bytes_len as usize) };
        (0..n)
            .map(|p| {
                if bad[p] != 0 {
                    return Err(VerifyError::BadRootLen);
                }
                match VerifyError::from_status(status[p]) {
                    None => Ok(arena[voff[p] as usize..voff[p] as usize + vlen[p] as usize].to_vec()),
                    Some(e) => Err(e),
                }
            })
            .collect()
    }

    /// The risc0 storage guest (storage-circuit/src/main.rs:6-31) for one input: the verified storage values.
    pub fn verify_storage_proof_input(&mut self, inp: &StorageProofInput) -> Result<Vec<Vec<u8>>, VerifyError> {
        let mut items = vec![MerkleProofInput {
            proof: inp.account_proof.clone(),
            root_hash: inp.root_hash.clone(),
            key: inp.address_keccak.to_vec(),
        }];
        let (mut rfp, mut hk) = (vec![-1i32], vec![0u8]);
        for (proof, key) in inp.storage_proofs.iter().zip(inp.storage_keys.iter()) {
            items.push(MerkleProofInput { proof: proof.clone(), root_hash: vec![0u8; 32], key: key.clone() });
            rfp.push(0);
            hk.push(1);
        }
        let mut out = Vec::new();
        for (i, r) in self.run(&items, Some(&rfp), Some(&hk)).into_iter().enumerate() {
            let v = r?;
            if i > 0 {
                out.push(v);
            }
        }
        Ok(out)
    }
}

impl Drop for Verifier {
    fn drop(&mut self) {
        unsafe {
            mptv_host_batch_free(self.flat);
            mptv_destroy(self.ctx);
        }
    }
}

/// Same signature and panic behaviour as `crypto_ops::verify_merkle_proof` (crypto-ops/src/lib.rs:8-23).
pub fn verify_merkle_proof(root_hash: B256, proof: Vec<Vec<u8>>, key: &[u8]) -> Vec<u8> {
    thread_local!(static V: std::cell::RefCell<Verifier> = std::cell::RefCell::new(Verifier::new().expect("no B200: there is no CPU fallback")));
    let inp = MerkleProofInput { proof, root_hash: root_hash.to_vec(), key: key.to_vec() };
    match V.with(|v| v.borrow_mut().verify_merkle_proofs(std::slice::from_ref(&inp)).pop().unwrap()) {
        Ok(v) => v,
        Err(e) => panic!("{}", e.panic_message()),
    }
}
