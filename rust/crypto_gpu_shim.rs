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
//   * batch front ends for the consensus call sites: many QCs (view-change burst), TC votes, and whole bincode frames
//     (ingest_frames + verify_ingested: the receiver path of consensus.rs:138 without building the message structs first).
use std::os::raw::c_int;
use std::sync::OnceLock;

#[repr(C)] pub struct HsCtx { _private: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy)] pub struct HsRec128 { pub sig: [u8; 64], pub pk: [u8; 32], pub msg: [u8; 32] } // (Signature, PublicKey, Digest)
#[repr(C)] #[derive(Clone, Copy)] pub struct HsVote   { pub pk: [u8; 32], pub sig: [u8; 64] }                     // QC.votes element, messages.rs:168

/// Mirror of hs_frame_info (48 bytes): which items of the ingest output belong to frame i's author signature / QC votes / TC votes.
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct HsFrameInfo { pub kind: u8, pub has_tc: u8, pub qc_is_genesis: u8, pub pad: u8, pub author_item: u32,
                         pub qc_lo: u32, pub qc_hi: u32, pub tc_lo: u32, pub tc_hi: u32, pub round: u64, pub qc_round: u64, pub tc_round: u64 }
/// Mirror of hs_ingest_out: caller-owned arrays (capacities in, counts out).
#[repr(C)]
pub struct HsIngestOut { pub cap_items: usize, pub cap_msgs: usize, pub cap_pre_bytes: usize,
                         pub sig: *mut u8, pub pk: *mut u8, pub msg_idx: *mut u32, pub group_idx: *mut u32, pub mode: *mut u8,
                         pub preimages: *mut u8, pub pre_off: *mut u64, pub n_items: usize, pub n_msgs: usize, pub pre_bytes: usize }
pub const HS_FRAME_MALFORMED: u8 = 255;

pub const HS_OK: c_int = 0;
pub const HS_ERR_NOMEM: c_int = 3;
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
    fn hs_verify_groups(ctx: *mut HsCtx, preimages: *const u8, pre_off: *const u64, n_msgs: usize, sig: *const u8, pk: *const u8, vidx: *const u32,
                        msg_idx: *const u32, group_idx: *const u32, mode: *const u8, n_items: usize, n_groups: usize,
                        out_item_bitmap: *mut u32, out_group_bitmap: *mut u32) -> c_int;
    fn hs_ingest_consensus_frames(frames: *const u8, off: *const u64, n: usize, info: *mut HsFrameInfo, out: *mut HsIngestOut) -> c_int;
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

/// Everything `ingest_frames` extracted: frame i is group i of one hs_verify_groups pass.
pub struct Ingested { pub info: Vec<HsFrameInfo>, pub sig: Vec<u8>, pub pk: Vec<u8>, pub msg_idx: Vec<u32>, pub group_idx: Vec<u32>, pub mode: Vec<u8>,
                      pub preimages: Vec<u8>, pub pre_off: Vec<u64> }

/// Receiver side (consensus/src/consensus.rs:138): bincode `ConsensusMessage` frames -> flat arrays, without building the message
/// structs first.  Host-only (works without a GPU).  Malformed frames come back with kind == HS_FRAME_MALFORMED and no items.
pub fn ingest_frames(frames: &[&[u8]]) -> Result<Ingested, GpuError> {
    let n = frames.len();
    let mut off = Vec::with_capacity(n + 1);
    let mut blob = Vec::new();
    off.push(0u64);
    for f in frames { blob.extend_from_slice(f); off.push(blob.len() as u64); }
    let mut info = vec![HsFrameInfo::default(); n.max(1)];
    // every item costs >= 116 frame bytes and every preimage is copied from frame bytes: generous first guess, exact retry on NOMEM
    let (mut ci, mut cm, mut cp) = (blob.len() / 116 + n + 1, blob.len() / 60 + n + 1, blob.len() + 64 * n + 64);
    for _ in 0..2 {
        let mut g = Ingested { info: Vec::new(), sig: vec![0; ci * 64], pk: vec![0; ci * 32], msg_idx: vec![0; ci], group_idx: vec![0; ci], mode: vec![0; ci],
                               preimages: vec![0; cp], pre_off: vec![0; cm + 1] };
        let mut o = HsIngestOut { cap_items: ci, cap_msgs: cm, cap_pre_bytes: cp, sig: g.sig.as_mut_ptr(), pk: g.pk.as_mut_ptr(), msg_idx: g.msg_idx.as_mut_ptr(),
                                  group_idx: g.group_idx.as_mut_ptr(), mode: g.mode.as_mut_ptr(), preimages: g.preimages.as_mut_ptr(), pre_off: g.pre_off.as_mut_ptr(),
                                  n_items: 0, n_msgs: 0, pre_bytes: 0 };
        let rc = unsafe { hs_ingest_consensus_frames(blob.as_ptr(), off.as_ptr(), n, info.as_mut_ptr(), &mut o) };
        if rc == HS_OK {
            g.sig.truncate(o.n_items * 64); g.pk.truncate(o.n_items * 32); g.msg_idx.truncate(o.n_items); g.group_idx.truncate(o.n_items);
            g.mode.truncate(o.n_items); g.preimages.truncate(o.pre_bytes); g.pre_off.truncate(o.n_msgs + 1);
            info.truncate(n);
            g.info = info;
            return Ok(g);
        }
        if rc != HS_ERR_NOMEM { return Err(GpuError::Engine(format!("hs_ingest_consensus_frames: status {}", rc))); }
        ci = o.n_items + 1; cm = o.n_msgs + 1; cp = o.pre_bytes + 1;
    }
    Err(GpuError::Engine("ingest capacity retry failed".into()))
}
/// One GPU pass over everything `ingest_frames` found: bit j of the result = every signature of frame j verified (author strict, QC
/// votes by the batch equation, TC votes strict).  The stake / duplicate / genesis pre-checks of messages.rs stay with the caller, who
/// has the item ranges in `info`.  None = no GPU.
pub fn verify_ingested(g: &Ingested) -> Option<Vec<bool>> {
    let c = ctx()?;
    let n_groups = g.info.len();
    let mut gbm = vec![0u32; (n_groups + 31) / 32 + 1];
    let rc = unsafe { hs_verify_groups(c, g.preimages.as_ptr(), g.pre_off.as_ptr(), g.pre_off.len().saturating_sub(1), g.sig.as_ptr(), g.pk.as_ptr(), std::ptr::null(),
                                       g.msg_idx.as_ptr(), g.group_idx.as_ptr(), g.mode.as_ptr(), g.msg_idx.len(), n_groups, std::ptr::null_mut(), gbm.as_mut_ptr()) };
    Some((0..n_groups).map(|j| rc == HS_OK && g.info[j].kind != HS_FRAME_MALFORMED && gbm[j / 32] >> (j % 32) & 1 == 1).collect())
}
