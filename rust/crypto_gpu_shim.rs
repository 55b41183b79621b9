// crypto/src/gpu.rs — FFI to libhs_crypto.so (include/hs_crypto.h) for asonnino/hotstuff's `crypto` crate.
//
// STATUS: source only — this image has no cargo/rustc, so this file has never been compiled here.  The same C ABI is exercised
// end to end by tests/ (ctypes) and tests/cpp/crypto_tests.cpp.  INTEGRATION.md §2 shows the edits in crypto/src/lib.rs.
//
// What it adds over a bare FFI pass-through:
//   * CPU/GPU cut-over: a lone signature or a single large Digest stays on the reference's own dalek / sha2 path;
//     the GPU takes every call with >= GPU_MIN_SIGS signatures and every multi-message digest.  Thresholds come from
//     tools/replay_config5 on B200 (profiles/r02_replay_config5.json): 1 verify 72 us GPU vs ~62 us one CPU core; 3 votes
//     72 us GPU vs ~185 us CPU; one 15 kB batch digest 327 us GPU vs ~40 us CPU.
//   * every status code is propagated: a failed registration or engine call is an Err / a rejected message, never an accept.
//   * batch front ends for the consensus call sites: many QCs (view-change burst), TC votes, whole frames.
use std::os::raw::c_int;
use std::sync::OnceLock;

#[repr(C)] pub struct HsCtx { _private: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy)] pub struct HsRec128 { pub sig: [u8; 64], pub pk: [u8; 32], pub msg: [u8; 32] } // (Signature, PublicKey, Digest)
#[repr(C)] #[derive(Clone, Copy)] pub struct HsVote   { pub pk: [u8; 32], pub sig: [u8; 64] }                     // QC.votes element, messages.rs:168

pub const HS_OK: c_int = 0;
/// Smallest signature count sent to the GPU (below it the dalek path is faster on this hardware; see the header comment).
pub const GPU_MIN_SIGS: usize = 2;
/// A Digest call goes to the GPU only with at least this many messages in flight (SHA-512 is sequential inside one message).
pub const GPU_MIN_DIGEST_MSGS: usize = 8;

#[link(name = "hs_crypto")]
extern "C" {
    fn hs_ctx_create(out: *mut *mut HsCtx, device: c_int, flags: u32) -> c_int;
    fn hs_last_error(ctx: *const HsCtx) -> *const std::os::raw::c_char;
    fn hs_committee_register(ctx: *mut HsCtx, pks: *const u8, n: usize, out_valid_bitmap: *mut u32) -> c_int;
    fn hs_committee_update(ctx: *mut HsCtx, add_pks: *const u8, n_add: usize, remove_idx: *const u32, n_remove: usize, out_add_idx: *mut u32) -> c_int;
    fn hs_verify_strict_batch(ctx: *mut HsCtx, recs: *const HsRec128, n: usize, out_bitmap: *mut u32) -> c_int;
    fn hs_verify_batch_shared_msg(ctx: *mut HsCtx, digest: *const u8, votes: *const HsVote, n: usize,
                                  all_ok: *mut c_int, out_bitmap_or_null: *mut u32) -> c_int;
    fn hs_verify_qcs(ctx: *mut HsCtx, preimages: *const u8, n_qc: usize, pk: *const u8, vidx: *const u32, sig: *const u8,
                     qc_idx: *const u32, n_votes: usize, out_vote_bitmap: *mut u32, out_qc_bitmap: *mut u32) -> c_int;
    fn hs_verify_tcs(ctx: *mut HsCtx, tc_rounds: *const u64, n_tc: usize, pk: *const u8, vidx: *const u32, sig: *const u8,
                     high_qc_rounds: *const u64, tc_idx: *const u32, n_votes: usize, out_vote_bitmap: *mut u32, out_tc_bitmap: *mut u32) -> c_int;
    fn hs_digest32_batch(ctx: *mut HsCtx, data: *const u8, off: *const u64, n: usize, out: *mut u8) -> c_int;
}

struct Ctx(*mut HsCtx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}            // host-pointer entry points are serialised on the context's mutex
static CTX: OnceLock<Option<Ctx>> = OnceLock::new();

/// None when no GPU / the library failed to initialise: every caller below then stays on the CPU path.
fn ctx() -> Option<*mut HsCtx> {
    CTX.get_or_init(|| {
        let mut p = std::ptr::null_mut();
        if unsafe { hs_ctx_create(&mut p, 0, 0) } == HS_OK && !p.is_null() { Some(Ctx(p)) } else { None }
    }).as_ref().map(|c| c.0)
}
fn last_error(c: *mut HsCtx) -> String {
    unsafe { std::ffi::CStr::from_ptr(hs_last_error(c)).to_string_lossy().into_owned() }
}

#[derive(Debug)]
pub enum GpuError { Unavailable, Engine(String), InvalidKeys(Vec<usize>) }

/// Once per epoch from node/src/node.rs after the committee file is read (consensus/src/config.rs:28-60).  Keys that do not
/// decompress are reported (PublicKey::from_bytes would fail on them at first use, crypto/src/lib.rs:202).
pub fn register_committee(keys: &[[u8; 32]]) -> Result<(), GpuError> {
    let c = ctx().ok_or(GpuError::Unavailable)?;
    let mut valid = vec![0u32; (keys.len() + 31) / 32];
    let rc = unsafe { hs_committee_register(c, keys.as_ptr() as *const u8, keys.len(), valid.as_mut_ptr()) };
    if rc != HS_OK { return Err(GpuError::Engine(last_error(c))); }
    let bad: Vec<usize> = (0..keys.len()).filter(|i| valid[i / 32] >> (i % 32) & 1 == 0).collect();
    if bad.is_empty() { Ok(()) } else { Err(GpuError::InvalidKeys(bad)) }
}
/// Incremental epoch change: returns the table indices of the added validators.
pub fn update_committee(add: &[[u8; 32]], remove_idx: &[u32]) -> Result<Vec<u32>, GpuError> {
    let c = ctx().ok_or(GpuError::Unavailable)?;
    let mut out = vec![0u32; add.len().max(1)];
    let rc = unsafe { hs_committee_update(c, add.as_ptr() as *const u8, add.len(), remove_idx.as_ptr(), remove_idx.len(), out.as_mut_ptr()) };
    if rc != HS_OK { return Err(GpuError::Engine(last_error(c))); }
    out.truncate(add.len());
    Ok(out)
}

/// Signature::verify for n triples.  None = "use the CPU path" (no GPU, or too few signatures to pay for a launch);
/// Some(bits) = verdicts.  An engine failure rejects everything (core.rs:434-439 drops a message on any Err).
pub fn verify_strict_many(recs: &[HsRec128]) -> Option<Vec<bool>> {
    if recs.len() < GPU_MIN_SIGS { return None; }
    let c = ctx()?;
    let mut bm = vec![0u32; (recs.len() + 31) / 32];
    let rc = unsafe { hs_verify_strict_batch(c, recs.as_ptr(), recs.len(), bm.as_mut_ptr()) };
    Some((0..recs.len()).map(|i| rc == HS_OK && bm[i / 32] >> (i % 32) & 1 == 1).collect())
}
/// Signature::verify_batch (one digest, n votes).  None = use dalek.
pub fn verify_batch(digest: &[u8; 32], votes: &[HsVote]) -> Option<bool> {
    if votes.len() < GPU_MIN_SIGS { return None; }
    let c = ctx()?;
    let mut ok: c_int = 0;
    let rc = unsafe { hs_verify_batch_shared_msg(c, digest.as_ptr(), votes.as_ptr(), votes.len(), &mut ok, std::ptr::null_mut()) };
    Some(rc == HS_OK && ok == 1)
}
/// QC::verify for many certificates at once (the view-change burst: every Timeout carries a high_qc, core.rs:227).
/// `preimages[j]` = hash || round_le (40 bytes); vote i = (pk[i], sig[i]) of certificate qc_idx[i].  Stake / duplicate checks stay
/// with the caller (messages.rs:182-194).
pub fn verify_qcs(preimages: &[[u8; 40]], pk: &[[u8; 32]], sig: &[[u8; 64]], qc_idx: &[u32]) -> Option<Vec<bool>> {
    let c = ctx()?;
    let mut qbm = vec![0u32; (preimages.len() + 31) / 32 + 1];
    let rc = unsafe { hs_verify_qcs(c, preimages.as_ptr() as *const u8, preimages.len(), pk.as_ptr() as *const u8, std::ptr::null(),
                                    sig.as_ptr() as *const u8, qc_idx.as_ptr(), sig.len(), std::ptr::null_mut(), qbm.as_mut_ptr()) };
    Some((0..preimages.len()).map(|j| rc == HS_OK && qbm[j / 32] >> (j % 32) & 1 == 1).collect())
}
/// TC::verify (tc_idx = Some) / n Timeout signatures (tc_idx = None): the 16-byte digests are built on the GPU.
pub fn verify_tcs(tc_rounds: &[u64], pk: &[[u8; 32]], sig: &[[u8; 64]], high_qc_rounds: &[u64], tc_idx: Option<&[u32]>) -> Option<Vec<bool>> {
    let c = ctx()?;
    let mut tbm = vec![0u32; (tc_rounds.len() + 31) / 32 + 1];
    let rc = unsafe { hs_verify_tcs(c, tc_rounds.as_ptr(), tc_rounds.len(), pk.as_ptr() as *const u8, std::ptr::null(), sig.as_ptr() as *const u8,
                                    high_qc_rounds.as_ptr(), tc_idx.map_or(std::ptr::null(), |t| t.as_ptr()), sig.len(), std::ptr::null_mut(), tbm.as_mut_ptr()) };
    Some((0..tc_rounds.len()).map(|j| rc == HS_OK && tbm[j / 32] >> (j % 32) & 1 == 1).collect())
}
/// Digest = SHA-512[..32] of many messages (mempool/src/processor.rs:30 with several batches in flight).  None = hash on the CPU.
pub fn digest32_many(msgs: &[&[u8]]) -> Option<Vec<[u8; 32]>> {
    if msgs.len() < GPU_MIN_DIGEST_MSGS { return None; }
    let c = ctx()?;
    let mut off = Vec::with_capacity(msgs.len() + 1);
    let mut data = Vec::new();
    off.push(0u64);
    for m in msgs { data.extend_from_slice(m); off.push(data.len() as u64); }
    let mut out = vec![[0u8; 32]; msgs.len()];
    let rc = unsafe { hs_digest32_batch(c, data.as_ptr(), off.as_ptr(), msgs.len(), out.as_mut_ptr() as *mut u8) };
    if rc == HS_OK { Some(out) } else { None }   // a failed digest call falls back to the CPU hash: a digest has no "reject"
}
