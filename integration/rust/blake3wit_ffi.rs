//! blake3wit_ffi.rs -- optional FFI hook for rust_fold: feeds Nova step witnesses from libblake3wit.so instead of
//! `circom_scotia::calculate_witness` (rust_fold/src/blake3_circuit.rs:303-305).  cargo is not available in this
//! repository's build image, so this file is documentation-grade: it is written against the C ABI in
//! include/blake3wit.h (whose entry points are exercised through ctypes by tests/) but has not been built here.
//!
//! Two seams, smallest first.
//!
//! (1) One step at a time -- in `Blake3BlockCompressCircuit::synthesize` (blake3_circuit.rs:297-313) replace
//!         let cfg = load_cfg::<G>(&self.circom_path_wasm, &self.circom_path_r1cs);   // re-reads wasm + r1cs every step
//!         let input = self.format_input(z)?;
//!         let witness = calculate_witness(&cfg, input, true).expect("msg");
//!     by
//!         let row = self.input_row(z);                       // the same values, as 32 u32 in circuit declaration order
//!         let witness: Vec<G::Scalar> = ctx.witness_one(&row)?;
//!     and keep `utils::synthesize_with_vec` (utils.rs:17-88), which only needs `w[i]` in witness order.
//!
//! (2) The whole chunk path at once -- `Ctx::nova_chain` computes, in ONE call, every step
//!     witness of every chunk of a file plus the z_i chain, i.e. everything the `prove_step` loop
//!     (main.rs:166-171: prove_step + update_for_step, `num_steps = n_blocks + total_depth - 1` times per chunk) asks
//!     `synthesize` for.  Replacement text for rust_fold:
//!
//!     main.rs, before the loop at :166 --
//!         let chain = ctx.nova_chain::<<E1 as Engine>::Scalar>(&input_bytes)?;     // all chunks, all steps
//!         let steps = chain.steps_of_chunk(chunk_idx as usize);                   // this chunk's range of step indices
//!         circuit_primary.attach_witnesses(chain.clone(), steps.start);            // Arc<NovaChain<F>> + cursor
//!     main.rs:166-171 stays as it is (prove_step, then update_for_step, which now also advances the cursor).
//!
//!     blake3_circuit.rs:297-313, `synthesize` --
//!         let witness = match &self.chain {                                        // attached: no wasm, no r1cs re-parse
//!             Some(chain) => chain.witness(self.cursor).to_vec(),
//!             None => calculate_witness(&load_cfg::<G>(..), self.format_input(z)?, true).expect("msg"),
//!         };
//!         utils::synthesize_with_vec::<G::Scalar, _>(&mut cs.namespace(|| "blake3_circom"), self.r1cs.clone(), Some(witness), self.arity())
//!     (`self.r1cs` parsed once in `new`; `synthesize_with_vec` still enforces every row on the vector it is handed,
//!     utils.rs:78-85 -- and `NovaChain::z_out(step)` equals its return value `vars[0..15]`, utils.rs:62, so a caller can
//!     also cross-check the chaining without synthesizing.)
//!
//!     The sibling values along each chunk's path are the TRUE BLAKE3 siblings by default; create the context with
//!     `B3W_FLAG_REFERENCE_SIBLINGS` to get `hash_with_path`'s choice (blake3_hash.rs:60-78) bit for bit -- the two agree on
//!     every 2^k-chunk file, the only shape rust_fold's tests assert on (main.rs:447-476).
//!
//!     A prover that runs on the same GPU takes `b3w_nova_chain_device` instead: the step witnesses stay in device memory
//!     (one launch), only the file crosses PCIe.
use std::ops::Range;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct B3wConfig { pub circuit: u32, pub device: i32, pub chunk: u32, pub flags: u32 }
#[repr(C)]
pub struct B3wInfo { pub witness_size: u32, pub n_inputs: u32, pub n32: u32, pub n_public: u32, pub version: [u32; 3], pub prime: [u8; 32] }
#[repr(C)]
pub struct B3wBatchExtras { pub sums: *mut u64, pub sample_idx: *const u64, pub n_samples: u32, pub sample_out: *mut u8, pub first_bad: *mut u32 }
pub enum B3wCtx {}

#[link(name = "blake3wit")]
extern "C" {
    pub fn b3w_create(cfg: *const B3wConfig, out: *mut *mut B3wCtx) -> c_int;
    pub fn b3w_destroy(ctx: *mut B3wCtx);
    pub fn b3w_circuit_info(circuit: u32, info: *mut B3wInfo) -> c_int;
    pub fn b3w_witness_one(ctx: *mut B3wCtx, input: *const u32, out: *mut u8) -> c_int;
    pub fn b3w_witness_batch(ctx: *mut B3wCtx, input: *const u32, n: u64, out: *mut u8, status: *mut u8, publ: *mut u32) -> c_int;
    pub fn b3w_witness_batch_ex(ctx: *mut B3wCtx, input: *const u32, n: u64, out: *mut u8, status: *mut u8, publ: *mut u32,
                                extras: *const B3wBatchExtras) -> c_int;
    /// inputs as little-endian field elements (`F::to_repr()`), n rows of n_inputs x 32 bytes; any field element is taken
    pub fn b3w_witness_batch_fr(ctx: *mut B3wCtx, input_fr: *const u8, n: u64, out: *mut u8, status: *mut u8, publ: *mut u32) -> c_int;
    pub fn b3w_assert_trace_fr(circuit: u32, input_fr: *const u8, buf: *mut c_char, cap: usize) -> c_int;
    pub fn b3w_nova_chain_size(len: u64, n_chunks: *mut u64, total_steps: *mut u64) -> c_int;
    pub fn b3w_nova_chain(ctx: *mut B3wCtx, data: *const u8, len: u64, out: *mut u8, status: *mut u8, publ: *mut u32,
                          rows: *mut u32, step_off: *mut u64, root: *mut u8) -> c_int;
    /// as b3w_nova_chain with DEVICE buffers for out / status / publ / rows (a prover on the same GPU)
    pub fn b3w_nova_chain_device(ctx: *mut B3wCtx, data: *const u8, len: u64, d_out: *mut u8, d_status: *mut u8, d_publ: *mut u32,
                                 d_rows: *mut u32, step_off: *mut u64, root: *mut u8) -> c_int;
    pub fn b3w_host_alloc(bytes: usize) -> *mut u8;
    pub fn b3w_host_free(p: *mut u8);
    pub fn b3w_last_error() -> *const c_char;
}

pub const B3W_NOVA_PASTA_O2: u32 = 2; // ../build/blake3_nova_pasta_js/blake3_nova_pasta.wasm (main.rs:364-365)
pub const B3W_FLAG_FUSED_CHECK: u32 = 1;
pub const B3W_FLAG_REFERENCE_SIBLINGS: u32 = 8;
/// every chunk is re-read from the HBM ring and all rows are evaluated on its bytes (what `synthesize_with_vec`, utils.rs:78-85,
/// would enforce on the vector) before the witnesses leave the GPU
pub const B3W_FLAG_BYTE_CHECK: u32 = 16;
pub const IO_ARITY: usize = 15; // blake3_circuit.rs:15

fn last_error() -> String {
    unsafe { std::ffi::CStr::from_ptr(b3w_last_error()).to_string_lossy().into_owned() }
}

pub struct Ctx { raw: *mut B3wCtx, witness_size: usize }

/// Pinned host buffer from the library (full PCIe rate, overlaps with kernels); freed on drop.
struct Pinned { p: *mut u8, len: usize }
impl Pinned {
    fn new(len: usize) -> Result<Self, String> {
        let p = unsafe { b3w_host_alloc(len.max(1)) };
        if p.is_null() { Err(last_error()) } else { Ok(Pinned { p, len }) }
    }
    fn bytes(&self) -> &[u8] { unsafe { std::slice::from_raw_parts(self.p, self.len) } }
}
impl Drop for Pinned { fn drop(&mut self) { unsafe { b3w_host_free(self.p) } } }

/// Every step witness of every chunk of one file, as the library left them in a pinned buffer, plus the z chain.
/// `witness(step)` converts lazily: `synthesize_with_vec` wants a `Vec<F>` per step, not 20 GB of them at once.
pub struct NovaChain<F> {
    witness_size: usize,
    out: Pinned,                  // total_steps x witness_size x 32 bytes, canonical little-endian
    pub status: Vec<u8>,          // per step: 0, or 4 = "Assert Failed."
    pub z: Vec<[u32; IO_ARITY]>,  // z_{i+1} of every step = the circuit's 15 outputs (all fit u32 for u32 inputs)
    pub rows: Vec<[u32; 32]>,     // the step inputs format_input() would have built (blake3_circuit.rs:197-289)
    pub step_off: Vec<u64>,       // n_chunks + 1: first step of each chunk
    pub root: [u8; 32],           // h_out of chunk 0's last step
    _f: std::marker::PhantomData<F>,
}

impl<F: ff::PrimeField<Repr = [u8; 32]>> NovaChain<F> {
    pub fn steps_of_chunk(&self, chunk: usize) -> Range<usize> { self.step_off[chunk] as usize..self.step_off[chunk + 1] as usize }
    /// the witness vector of one step in witness order: w[0] = 1, w[1..16] = z_{i+1}, w[16..28] = the public inputs
    /// (what utils.rs:33-56 splits into `public_*` and `aux_*`)
    pub fn witness(&self, step: usize) -> Vec<F> {
        let b = &self.out.bytes()[step * self.witness_size * 32..(step + 1) * self.witness_size * 32];
        b.chunks_exact(32).map(|c| { let mut r = [0u8; 32]; r.copy_from_slice(c); F::from_repr(r).unwrap() }).collect()
    }
    /// z_{i+1} as field elements (= `synthesize`'s return value, utils.rs:62)
    pub fn z_out(&self, step: usize) -> [F; IO_ARITY] { self.z[step].map(|x| F::from(x as u64)) }
}

impl Ctx {
    pub fn new(circuit: u32) -> Result<Self, String> { Self::with_flags(circuit, 0) }

    pub fn with_flags(circuit: u32, flags: u32) -> Result<Self, String> {
        let cfg = B3wConfig { circuit, device: -1, chunk: 0, flags };
        let mut raw = std::ptr::null_mut();
        let mut info = unsafe { std::mem::zeroed::<B3wInfo>() };
        unsafe {
            if b3w_create(&cfg, &mut raw) != 0 || b3w_circuit_info(circuit, &mut info) != 0 {
                return Err(last_error());
            }
        }
        Ok(Ctx { raw, witness_size: info.witness_size as usize })
    }

    /// One step witness as field elements: every slot is a canonical little-endian 32-byte repr -> `F::from_repr`.
    pub fn witness_one<F: ff::PrimeField<Repr = [u8; 32]>>(&self, row: &[u32; 32]) -> Result<Vec<F>, String> {
        let mut bytes = vec![0u8; self.witness_size * 32];
        let rc = unsafe { b3w_witness_one(self.raw, row.as_ptr(), bytes.as_mut_ptr()) };
        if rc != 0 {
            return Err(if rc == 4 { "Assert Failed.".into() } else { last_error() });
        }
        Ok(bytes.chunks_exact(32).map(|c| { let mut r = [0u8; 32]; r.copy_from_slice(c); F::from_repr(r).unwrap() }).collect())
    }

    /// All Nova step witnesses of `data` in one call (the batched form of main.rs:166-171 + blake3_circuit.rs:297-313).
    pub fn nova_chain<F: ff::PrimeField<Repr = [u8; 32]>>(&self, data: &[u8]) -> Result<std::sync::Arc<NovaChain<F>>, String> {
        let (mut nc, mut ns) = (0u64, 0u64);
        if unsafe { b3w_nova_chain_size(data.len() as u64, &mut nc, &mut ns) } != 0 { return Err(last_error()); }
        let (nc, ns) = (nc as usize, ns as usize);
        let out = Pinned::new(ns * self.witness_size * 32)?;
        let mut chain = NovaChain::<F> {
            witness_size: self.witness_size, out, status: vec![0u8; ns], z: vec![[0u32; IO_ARITY]; ns], rows: vec![[0u32; 32]; ns],
            step_off: vec![0u64; nc + 1], root: [0u8; 32], _f: std::marker::PhantomData,
        };
        let rc = unsafe {
            b3w_nova_chain(self.raw, data.as_ptr(), data.len() as u64, chain.out.p, chain.status.as_mut_ptr(), chain.z.as_mut_ptr() as *mut u32,
                           chain.rows.as_mut_ptr() as *mut u32, chain.step_off.as_mut_ptr(), chain.root.as_mut_ptr())
        };
        if rc != 0 { return Err(last_error()); }
        if let Some(i) = chain.status.iter().position(|&s| s != 0) { return Err(format!("step {}: Assert Failed.", i)); }
        Ok(std::sync::Arc::new(chain))
    }
}

impl Drop for Ctx {
    fn drop(&mut self) { unsafe { b3w_destroy(self.raw) } }
}
