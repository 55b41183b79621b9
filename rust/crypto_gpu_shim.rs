// crypto/src/gpu.rs — FFI to libhs_crypto.so (include/hs_crypto.h)
use std::os::raw::{c_int, c_void};
use std::sync::OnceLock;

#[repr(C)] pub struct HsCtx { _private: [u8; 0] }
#[repr(C)] pub struct HsRec128 { pub sig: [u8; 64], pub pk: [u8; 32], pub msg: [u8; 32] }   // = (Signature, PublicKey, Digest)
#[repr(C)] pub struct HsVote   { pub pk: [u8; 32], pub sig: [u8; 64] }                        // = QC.votes element, messages.rs:168

#[link(name = "hs_crypto")]
extern "C" {
    fn hs_ctx_create(out: *mut *mut HsCtx, device: c_int, flags: u32) -> c_int;
    fn hs_verify_strict_batch(ctx: *mut HsCtx, recs: *const HsRec128, n: usize, out_bitmap: *mut u32) -> c_int;
    fn hs_verify_batch_shared_msg(ctx: *mut HsCtx, digest: *const u8, votes: *const HsVote, n: usize,
                                  all_ok: *mut c_int, out_bitmap_or_null: *mut u32) -> c_int;
    fn hs_committee_register(ctx: *mut HsCtx, pks: *const u8, n: usize, out_valid_bitmap: *mut u32) -> c_int;
    fn hs_digest32_batch(ctx: *mut HsCtx, data: *const u8, off: *const u64, n: usize, out: *mut u8) -> c_int;
}

struct Ctx(*mut HsCtx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}            // the C side serialises calls on an internal mutex
static CTX: OnceLock<Ctx> = OnceLock::new();
fn ctx() -> *mut HsCtx {
    CTX.get_or_init(|| { let mut p = std::ptr::null_mut(); let rc = unsafe { hs_ctx_create(&mut p, 0, 0) };
                         assert!(rc == 0, "hs_ctx_create failed: no GPU"); Ctx(p) }).0
}

/// Called once per epoch from node/src/node.rs after the committee file is read (consensus/src/config.rs:28-60).
pub fn register_committee(keys: &[[u8; 32]]) { unsafe { hs_committee_register(ctx(), keys.as_ptr() as *const u8, keys.len(), std::ptr::null_mut()); } }

pub fn verify_strict(sig: &[u8; 64], pk: &[u8; 32], digest: &[u8; 32]) -> bool {
    let rec = HsRec128 { sig: *sig, pk: *pk, msg: *digest };
    let mut word = 0u32;
    let rc = unsafe { hs_verify_strict_batch(ctx(), &rec, 1, &mut word) };
    rc == 0 && (word & 1) == 1                      // engine failure => reject (core.rs:434-439 drops the message on Err)
}
pub fn verify_batch(digest: &[u8; 32], votes: &[HsVote]) -> bool {
    let mut ok: c_int = 0;
    let rc = unsafe { hs_verify_batch_shared_msg(ctx(), digest.as_ptr(), votes.as_ptr(), votes.len(), &mut ok, std::ptr::null_mut()) };
    rc == 0 && ok == 1
}
pub fn digest32(data: &[u8]) -> [u8; 32] {
    let off = [0u64, data.len() as u64]; let mut out = [0u8; 32];
    let rc = unsafe { hs_digest32_batch(ctx(), data.as_ptr(), off.as_ptr(), 1, out.as_mut_ptr()) };
    assert!(rc == 0); out
}
