//! blake3wit_ffi.rs -- optional FFI hook for rust_fold: feeds Nova step witnesses from libblake3wit.so instead of
//! `circom_scotia::calculate_witness` (rust_fold/src/blake3_circuit.rs:303-305).  cargo is not available in this
//! repository's build image, so this file is documentation-grade: it compiles against the C ABI in
//! include/blake3wit.h but has not been built here.
//!
//! In `Blake3BlockCompressCircuit::synthesize` replace
//!     let cfg = load_cfg::<G>(&self.circom_path_wasm, &self.circom_path_r1cs);   // re-reads wasm + r1cs every step
//!     let input = self.format_input(z)?;
//!     let witness = calculate_witness(&cfg, input, true).expect("msg");
//! by
//!     let row = self.input_row(z);                       // the same values, as 32 u32 in circuit declaration order
//!     let witness: Vec<G::Scalar> = ctx.witness_one(&row);
//! and keep `utils::synthesize_with_vec` (rust_fold/src/utils.rs:17-88), which only needs `w[i]` in witness order.
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct B3wConfig { pub circuit: u32, pub device: i32, pub chunk: u32, pub flags: u32 }
#[repr(C)]
pub struct B3wInfo { pub witness_size: u32, pub n_inputs: u32, pub n32: u32, pub n_public: u32, pub version: [u32; 3], pub prime: [u8; 32] }
pub enum B3wCtx {}

#[link(name = "blake3wit")]
extern "C" {
    pub fn b3w_create(cfg: *const B3wConfig, out: *mut *mut B3wCtx) -> c_int;
    pub fn b3w_destroy(ctx: *mut B3wCtx);
    pub fn b3w_circuit_info(circuit: u32, info: *mut B3wInfo) -> c_int;
    pub fn b3w_witness_one(ctx: *mut B3wCtx, input: *const u32, out: *mut u8) -> c_int;
    pub fn b3w_witness_batch(ctx: *mut B3wCtx, input: *const u32, n: u64, out: *mut u8, status: *mut u8, publ: *mut u32) -> c_int;
    /// inputs as canonical little-endian field elements (`F::to_repr()`), n rows of n_inputs x 32 bytes
    pub fn b3w_witness_batch_fr(ctx: *mut B3wCtx, input_fr: *const u8, n: u64, out: *mut u8, status: *mut u8, publ: *mut u32) -> c_int;
    pub fn b3w_assert_trace_fr(circuit: u32, input_fr: *const u8, buf: *mut c_char, cap: usize) -> c_int;
    pub fn b3w_nova_chain_size(len: u64, n_chunks: *mut u64, total_steps: *mut u64) -> c_int;
    pub fn b3w_nova_chain(ctx: *mut B3wCtx, data: *const u8, len: u64, out: *mut u8, status: *mut u8, publ: *mut u32,
                          rows: *mut u32, step_off: *mut u64, root: *mut u8) -> c_int;
    pub fn b3w_last_error() -> *const c_char;
}

pub const B3W_NOVA_PASTA_O2: u32 = 2; // ../build/blake3_nova_pasta_js/blake3_nova_pasta.wasm (main.rs:364-365)

pub struct Ctx { raw: *mut B3wCtx, witness_size: usize }

impl Ctx {
    pub fn new(circuit: u32) -> Result<Self, String> {
        let cfg = B3wConfig { circuit, device: -1, chunk: 0, flags: 0 };
        let mut raw = std::ptr::null_mut();
        let mut info = unsafe { std::mem::zeroed::<B3wInfo>() };
        unsafe {
            if b3w_create(&cfg, &mut raw) != 0 || b3w_circuit_info(circuit, &mut info) != 0 {
                return Err(std::ffi::CStr::from_ptr(b3w_last_error()).to_string_lossy().into_owned());
            }
        }
        Ok(Ctx { raw, witness_size: info.witness_size as usize })
    }

    /// One step witness as field elements: every slot is a canonical little-endian 32-byte repr -> `F::from_repr`.
    pub fn witness_one<F: ff::PrimeField<Repr = [u8; 32]>>(&self, row: &[u32; 32]) -> Result<Vec<F>, String> {
        let mut bytes = vec![0u8; self.witness_size * 32];
        let rc = unsafe { b3w_witness_one(self.raw, row.as_ptr(), bytes.as_mut_ptr()) };
        if rc != 0 {
            return Err(if rc == 4 { "Assert Failed.".into() } else { unsafe { std::ffi::CStr::from_ptr(b3w_last_error()).to_string_lossy().into_owned() } });
        }
        Ok(bytes.chunks_exact(32).map(|c| { let mut r = [0u8; 32]; r.copy_from_slice(c); F::from_repr(r).unwrap() }).collect())
    }
}

impl Drop for Ctx {
    fn drop(&mut self) { unsafe { b3w_destroy(self.raw) } }
}
