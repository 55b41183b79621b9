/*
 * hs_crypto.h — C ABI of the B200 batch Ed25519 verification / SHA-512 digest engine.
 *
 * This is the drop-in boundary for ONE path of asonnino/hotstuff: the `crypto` crate's verify / verify_batch /
 * Digest surface.  The reference has no FFI today (it calls ed25519-dalek directly, crypto/Cargo.toml:10); each
 * entry point below names the reference interface it replaces (paths relative to the reference repo root).
 * INTEGRATION.md shows the Rust `extern "C"` binding a maintainer would add to crypto/src/lib.rs.
 *
 * Conventions
 *   - Every function returns 0 (HS_OK) when the engine ran; verdicts are in the output buffers.  Non-zero = engine
 *     failure (CUDA error, bad argument): the caller must treat every signature of that call as REJECTED
 *     (reference behaviour: any Err drops the message, consensus/src/core.rs:434-439).  There is no CPU fallback.
 *   - Malformed inputs (S >= l, non-decompressible A or R, ...) are verdict 0, never an error.
 *   - Host-pointer entry points copy inputs to the device, run, and copy results back before returning; nothing is
 *     retained.  `_dev` entry points take device pointers and a cudaStream_t (as void*) and return after enqueueing.
 *   - A context is bound to one CUDA device and is thread-safe (calls are serialised on an internal mutex).
 *   - Bitmaps: bit (i & 31) of word (i >> 5) is the verdict of item i; unused high bits of the last word are 0.
 */
#ifndef HS_CRYPTO_H
#define HS_CRYPTO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HS_FLAG_NO_KEY_CACHE 0x10000u /* hs_ctx_create flag: never learn unregistered keys */
#define HS_OK 0
#define HS_ERR_CUDA 1
#define HS_ERR_ARG 2
#define HS_ERR_NOMEM 3

/* verdict selector for the verify entry points */
#define HS_MODE_STRICT 0u /* Signature::verify semantics  = dalek verify_strict          (crypto/src/lib.rs:200-204) */
#define HS_MODE_BATCH_EQ 1u /* per-signature condition of Signature::verify_batch          (crypto/src/lib.rs:206-219) */

typedef struct hs_ctx hs_ctx;

/* One (Signature, PublicKey, Digest) triple exactly as Signature::verify receives it:
 * sig = part1 || part2 (crypto/src/lib.rs:179-182,193-198), pk = PublicKey.0 (:66), msg = Digest.0 (:22). */
typedef struct {
  uint8_t sig[64];
  uint8_t pk[32];
  uint8_t msg[32];
} hs_rec128;

/* One QC vote = (PublicKey, Signature), consensus/src/messages.rs:168 `votes: Vec<(PublicKey, Signature)>`. */
typedef struct {
  uint8_t pk[32];
  uint8_t sig[64];
} hs_vote;

/* ---- lifecycle ------------------------------------------------------------------------------------------------ */
/* device: CUDA ordinal.  Builds the base-point comb table on the GPU (default window 24 bits = 8.9 GB of HBM).
 * flags: 0 = defaults; bits 0-7 = base-point window width (even, 8..26; 26 = 10 windows in 32 GB, one addition fewer per verify), bits 8-15 = forced per-key window width
 * (8..17; 0 = widest that fits ~62 % of device memory); HS_FLAG_NO_KEY_CACHE disables the key cache (also env HS_KEY_CACHE=0). */
int hs_ctx_create(hs_ctx **out, int device, uint32_t flags);
void hs_ctx_destroy(hs_ctx *ctx);
/* Human-readable description of the last failure on this context (never NULL). */
const char *hs_last_error(const hs_ctx *ctx);
/* Key cache: when NO committee is registered, keys that show up in calls carrying key bytes are learned between calls (up to
 * 4,096 keys, 14 MB of table each, allocated on first use): the first sighting of a key takes the generic path, later
 * ones the table path.  Verdicts are identical either way.  A full cache that misses on more than half of a pass is reset
 * and relearns (validator-set rotation); there is no per-key eviction, so registering the committee remains the robust
 * choice against floods of one-off keys.  Registering switches learning off; hs_committee_register(.., 0, ..) clears the committee
 * and re-enables it.  Returns the number of keys currently cached. */
size_t hs_cached_keys(const hs_ctx *ctx);
/* Comb window widths in use: per-key tables (0 when no committee is registered) and the base-point table. */
void hs_window_bits(const hs_ctx *ctx, int *key_bits, int *base_bits);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t hs_kernel_launches(const hs_ctx *ctx);
/* Measurement hook (bench.py's roofline): with profiling on, CUDA events bracket the k_verify_main<committee> launch of every
 * verify pass, on the stream the pass runs on; hs_profile_main_ms() waits for the last pass and returns that kernel's duration. */
int hs_profile_enable(hs_ctx *ctx, int on);
double hs_profile_main_ms(hs_ctx *ctx);
/* Pinned host memory helpers (optional; any host pointer is accepted by the host entry points). */
void *hs_host_alloc(size_t bytes);
void hs_host_free(void *p);

/* ---- Signature::verify (crypto/src/lib.rs:200-204), n independent triples -------------------------------------- */
/* Callers: Block::verify consensus/src/messages.rs:64, Vote::verify :144, Timeout::verify :258, TC::verify :312. */
int hs_verify_strict_batch(hs_ctx *ctx, const hs_rec128 *recs, size_t n, uint32_t *out_bitmap);
/* Same with an explicit verdict mode (HS_MODE_*). */
int hs_verify_rec128(hs_ctx *ctx, const hs_rec128 *recs, size_t n, uint32_t mode, uint32_t *out_bitmap);
/* Variable-length messages (PureEdDSA over the raw bytes): sig[n][64], pk[n][32], msgs concatenated, off[n+1]. */
int hs_verify_var(hs_ctx *ctx, const uint8_t *sig, const uint8_t *pk, const uint8_t *msgs, const uint64_t *off, size_t n,
                  uint32_t mode, uint32_t *out_bitmap);

/* ---- Signature::verify_batch (crypto/src/lib.rs:206-219): one digest, n votes ---------------------------------- */
/* Caller: QC::verify consensus/src/messages.rs:197.  *all_ok = 1 iff every vote parses and satisfies the
 * cofactorless equation (deterministic restatement of dalek::verify_batch). */
int hs_verify_batch_shared_msg(hs_ctx *ctx, const uint8_t digest[32], const hs_vote *votes, size_t n, int *all_ok,
                               uint32_t *out_bitmap_or_null);

/* ---- QC::verify for many certificates in one pass (consensus/src/messages.rs:180-208) --------------------------------- */
/* preimages: n_qc x 40 bytes = hash[32] || round_le[8]; the engine computes QC::digest (messages.rs:201-208) on the GPU.
 * Vote i = (key_i, sig_i) belongs to certificate qc_idx[i]; key_i = pk[i] (32 B each) or, when pk is NULL, committee key
 * validator_idx[i].  out_qc_bitmap bit j = AND over certificate j's votes of the verify_batch condition (no votes -> 1);
 * out_vote_bitmap (nullable) = per-vote verdicts.  The stake / duplicate checks of messages.rs:182-194 stay on the host.
 * This is the view-change burst of SURVEY §3D (every Timeout carries a high_qc) as ONE engine call. */
int hs_verify_qcs(hs_ctx *ctx, const uint8_t *preimages, size_t n_qc, const uint8_t *pk_or_null, const uint32_t *validator_idx_or_null,
                  const uint8_t *sig /* n_votes x 64 */, const uint32_t *qc_idx, size_t n_votes, uint32_t *out_vote_bitmap_or_null,
                  uint32_t *out_qc_bitmap);

/* ---- TC::verify / Timeout::verify for many certificates (consensus/src/messages.rs:250-265,290-315) ------------------------ */
/* Vote i = (key_i, sig_i, high_qc_rounds[i]) of certificate tc_idx[i]; its message is SHA-512(tc_rounds[tc_idx[i]]_le ||
 * high_qc_rounds[i]_le)[..32], built and hashed ON THE GPU (messages.rs:307-311: n digests that differ in 8 bytes), judged with
 * Signature::verify (strict).  out_tc_bitmap bit j = AND over certificate j's votes (no votes -> 1).  tc_idx == NULL: vote i is
 * its own certificate (n_tc == n_votes) — the shape of n Timeout messages (Timeout::digest, messages.rs:268-275; the embedded
 * high_qc goes through hs_verify_qcs).  The stake / duplicate checks of messages.rs:292-304 stay on the host. */
int hs_verify_tcs(hs_ctx *ctx, const uint64_t *tc_rounds, size_t n_tc, const uint8_t *pk_or_null, const uint32_t *validator_idx_or_null,
                  const uint8_t *sig /* n_votes x 64 */, const uint64_t *high_qc_rounds, const uint32_t *tc_idx_or_null, size_t n_votes,
                  uint32_t *out_vote_bitmap_or_null, uint32_t *out_tc_bitmap);

/* ---- mixed groups: Block::verify for many blocks in one pass (consensus/src/messages.rs:54-76) ------------------------------ */
/* Item i signs Digest(preimages[pre_off[msg_idx[i]] .. pre_off[msg_idx[i]+1])) (hashed on the GPU), belongs to group group_idx[i]
 * and is judged by mode[i] (HS_MODE_*; NULL = all strict): a block is one group holding its author signature (strict, Block::digest
 * preimage :79-90), its QC's votes (batch-eq, 40-byte preimage) and its TC's votes (strict, 16-byte preimages).
 * out_group_bitmap bit j = AND over group j's items. */
int hs_verify_groups(hs_ctx *ctx, const uint8_t *preimages, const uint64_t *pre_off /* n_msgs + 1 */, size_t n_msgs, const uint8_t *sig /* n_items x 64 */,
                     const uint8_t *pk_or_null, const uint32_t *validator_idx_or_null, const uint32_t *msg_idx, const uint32_t *group_idx,
                     const uint8_t *mode_or_null, size_t n_items, size_t n_groups, uint32_t *out_item_bitmap_or_null, uint32_t *out_group_bitmap);

/* ---- wire-format ingest: bincode ConsensusMessage frames -> the arrays hs_verify_groups consumes ------------------------------ */
/* Replaces `bincode::deserialize::<ConsensusMessage>` + the per-signature walk of Block/Vote/Timeout/TC::verify
 * (consensus/src/consensus.rs:33-39,138; crypto/src/lib.rs:94-112 PublicKey = base64 *string*; :178-182 Signature = 2 x 32 raw bytes)
 * for the crypto path: frame i (frames[off[i] .. off[i+1])) becomes group i; every signature in it becomes one item (signature,
 * decoded key bytes, preimage index, verdict mode) and every digest preimage is laid out for on-GPU hashing:
 *   Propose(Block): author item (strict, Block::digest preimage) + QC votes (batch-eq; skipped for the genesis QC) + TC votes (strict)
 *   Vote: one strict item      Timeout: author item (strict, 16-byte preimage) + high_qc votes      TC: its votes (strict)
 * Host-only, stateless, thread-safe.  Output buffers are caller-owned (hs_host_alloc() them to make the following H2D copies DMA
 * directly).  Returns HS_OK, or HS_ERR_NOMEM when a capacity is too small — n_items / n_msgs / pre_bytes then hold the required
 * sizes.  A malformed frame (truncated, bad tag, bad base64, absurd length) gets kind HS_FRAME_MALFORMED and contributes no items
 * (the reference drops it with SerializationError).  The stake / duplicate pre-checks of messages.rs stay with the caller: the
 * item ranges in hs_frame_info say which keys belong to which certificate. */
#define HS_NO_ITEM 0xffffffffu
#define HS_FRAME_MALFORMED 255
typedef struct {
  uint8_t kind;          /* ConsensusMessage tag: 0 Propose, 1 Vote, 2 Timeout, 3 TC, 4 SyncRequest; HS_FRAME_MALFORMED */
  uint8_t has_tc;        /* Block.tc is Some / the frame is a TC */
  uint8_t qc_is_genesis; /* embedded QC == QC::genesis(): not verified upstream (messages.rs:67,261) */
  uint8_t pad;
  uint32_t author_item;  /* item of the author's signature (HS_NO_ITEM for TC / SyncRequest) */
  uint32_t qc_lo, qc_hi; /* items of the embedded QC's votes [lo, hi) */
  uint32_t tc_lo, tc_hi; /* items of the TC's votes [lo, hi) */
  uint64_t round, qc_round, tc_round;
} hs_frame_info;
typedef struct {
  size_t cap_items, cap_msgs, cap_pre_bytes; /* in: capacities */
  uint8_t *sig;        /* cap_items x 64 */
  uint8_t *pk;         /* cap_items x 32 */
  uint32_t *msg_idx;   /* cap_items */
  uint32_t *group_idx; /* cap_items (= frame index) */
  uint8_t *mode;       /* cap_items: HS_MODE_* */
  uint8_t *preimages;  /* cap_pre_bytes */
  uint64_t *pre_off;   /* cap_msgs + 1 */
  size_t n_items, n_msgs, pre_bytes; /* out */
} hs_ingest_out;
int hs_ingest_consensus_frames(const uint8_t *frames, const uint64_t *off /* n + 1 */, size_t n, hs_frame_info *info /* n */, hs_ingest_out *out);

/* ---- committee mode: keys registered once per epoch (consensus/src/config.rs:28-60 Committee) ------------------- */
/* Decompresses every key and builds its comb table in HBM (window 17 bits: 94 MB, 16: 50 MB, 15: 27 MB, 14: 14 MB, 12: 4.1 MB per key).  out_valid_bitmap (nullable): bit i = key i
 * decompresses.  Replaces the per-call PublicKey::from_bytes of crypto/src/lib.rs:202,216. */
int hs_committee_register(hs_ctx *ctx, const uint8_t *pks /* N x 32 */, size_t N, uint32_t *out_valid_bitmap);
/* Incremental epoch change: validators remove_idx[] stop verifying (their indices become free), keys add_pks[] take a free
 * or spare slot (registration reserves N/16, at least 16, spare slots) and only their tables are built; out_add_idx[i]
 * receives the index of add_pks[i] (an already-registered key returns its existing index).  Everyone else keeps index and
 * table.  HS_ERR_NOMEM when no slot is free: re-register.  Requires a registered committee. */
int hs_committee_update(hs_ctx *ctx, const uint8_t *add_pks /* n_add x 32 */, size_t n_add, const uint32_t *remove_idx, size_t n_remove,
                        uint32_t *out_add_idx);
/* Memory budget (bytes) for the per-key tables of the NEXT registration / key-cache allocation (0 = default, ~62 % of the
 * device; also env HS_TABLE_BUDGET_MB at context creation).  The engine picks the widest window that fits: e.g. 4,096 keys in
 * 18 GB -> 12-bit windows.  Lets the engine sit beside another tenant on the same GPU. */
int hs_set_table_budget(hs_ctx *ctx, size_t bytes);
/* Vote i is (validator_idx[i], sig[i]) over digests[msg_idx[i]].  msg_idx may be NULL when n_msgs == 1. */
int hs_verify_committee(hs_ctx *ctx, const uint32_t *validator_idx, const uint8_t *sig /* n x 64 */, const uint32_t *msg_idx,
                        const uint8_t *digests /* n_msgs x 32 */, size_t n_msgs, size_t n, uint32_t mode, uint32_t *out_bitmap);

/* ---- Digest surface: out[i] = SHA-512(data[off[i] .. off[i+1]))[0..32] ------------------------------------------ */
/* Replaces Sha512::digest(..)[..32] at mempool/src/processor.rs:30 and consensus/src/messages.rs:81,151,203,270,308. */
int hs_digest32_batch(hs_ctx *ctx, const uint8_t *data, const uint64_t *off, size_t n, uint8_t *out /* n x 32 */);

/* ---- reference-shaped end-to-end call: verdict_i = Signature::verify(Digest(msg_i), key_i) ----------------------- */
/* n fixed-size messages (msg_len bytes each, concatenated); key_i = pk[i] (pk != NULL) or committee key validator_idx[i].
 * Computes the 32-byte Digest on the GPU (mempool/src/processor.rs:30 / consensus/src/messages.rs digests) and verifies
 * over it, overlapping the host->device copy of one chunk with the kernels of the previous one. */
int hs_verify_msgs(hs_ctx *ctx, const uint8_t *sig /* n x 64 */, const uint8_t *pk_or_null /* n x 32 */, const uint32_t *validator_idx_or_null,
                   const uint8_t *msgs, size_t msg_len, size_t n, uint32_t mode, uint32_t *out_bitmap);

/* ---- device-resident entry points (inputs already in HBM; enqueue on `stream`, a cudaStream_t) ------------------- */
int hs_verify_rec128_dev(hs_ctx *ctx, const void *d_recs, size_t n, uint32_t mode, void *d_bitmap, void *stream);
int hs_verify_var_dev(hs_ctx *ctx, const void *d_sig, const void *d_pk, const void *d_msgs, const void *d_off, size_t n,
                      uint32_t mode, void *d_bitmap, void *stream);
int hs_verify_committee_dev(hs_ctx *ctx, const void *d_validator_idx, const void *d_sig, const void *d_msg_idx,
                            const void *d_digests, size_t n, uint32_t mode, void *d_bitmap, void *stream);
int hs_digest32_dev(hs_ctx *ctx, const void *d_data, const void *d_off, size_t n, void *d_out, void *stream);
/* Digest of n fixed-size messages (msg_len bytes each, concatenated): the transaction / payload shape.  16-byte aligned sizes
 * >= 128 take the staged kernel (coalesced loads; multiples of 128 also skip the padding block's message schedule). */
int hs_digest32_fixed_dev(hs_ctx *ctx, const void *d_msgs, size_t msg_len, size_t n, void *d_out, void *stream);
/* d_digests: n x 32 bytes of scratch that receives Digest(msg_i). */
int hs_verify_msgs_dev(hs_ctx *ctx, const void *d_sig, const void *d_pk_or_null, const void *d_validator_idx_or_null, const void *d_msgs,
                       size_t msg_len, size_t n, uint32_t mode, void *d_digests, void *d_bitmap, void *stream);

/* QC votes of this rank's shard (per-vote verify_batch condition) over precomputed QC digests (hs_digest32_fixed_dev over the
 * 40-byte preimages), and the per-QC AND over a (possibly all-gathered) vote bitmap — the device-resident pieces of
 * hs_verify_qcs, used when the votes of many QCs are sharded across GPUs (BASELINE config[3]). */
int hs_verify_qc_votes_dev(hs_ctx *ctx, const void *d_qc_digests, const void *d_pk_or_null, const void *d_validator_idx_or_null, const void *d_sig,
                           const void *d_qc_idx, size_t n_votes, void *d_vote_bitmap, void *stream);
int hs_qc_and_dev(hs_ctx *ctx, const void *d_vote_bitmap, const void *d_qc_idx, size_t n_votes, size_t n_qc, void *d_qc_bitmap, void *stream);

/* ---- load generation: RFC 8032 key generation and signing of 32-byte digests ON THE GPU ---------------------------------------
 * generate_keypair / Signature::new (crypto/src/lib.rs:167-175,185-191) for input synthesis only: the node itself signs one
 * message per request on the CPU (SignatureService) and keeps doing so.  Deterministic, byte-identical to dalek / OpenSSL.
 * Signature i is over digests[i] with key key_idx[i] (NULL: key i).  Secret seeds travel to the device: test / benchmark use. */
int hs_keygen_batch(hs_ctx *ctx, const uint8_t *seeds /* n x 32 */, size_t n, uint8_t *out_pks /* n x 32 */);
int hs_sign_digests(hs_ctx *ctx, const uint8_t *seeds, const uint8_t *pks, size_t n_keys, const uint32_t *key_idx_or_null,
                    const uint8_t *digests /* n x 32 */, size_t n, uint8_t *out_sig /* n x 64 */);
int hs_keygen_batch_dev(hs_ctx *ctx, const void *d_seeds, size_t n, void *d_pks, void *stream);
int hs_sign_digests_dev(hs_ctx *ctx, const void *d_seeds, const void *d_pks, size_t n_keys, const void *d_key_idx_or_null, const void *d_digests,
                        size_t n, void *d_sig, void *stream);

/* Deferred-results mode for STREAMS of `_dev` verify passes (e.g. one pass per QC burst): the latency-bound tail of a pass — finish
 * kernel incl. the peer exchange, hs_qc_and_dev — runs on an internal stream and overlaps the main kernel of the next pass (a serial
 * field inversion per block makes the tail ~80 us however small the pass).  Bitmaps are complete only after hs_results_wait(ctx, stream),
 * which makes `stream` wait for every tail enqueued so far; the signatures of a pass must stay valid until then.  Host-pointer entry points
 * must not be mixed with deferred passes in flight. */
int hs_set_deferred(hs_ctx *ctx, int on);
int hs_results_wait(hs_ctx *ctx, void *stream);

/* ---- multi-GPU: fused all-gather of the accept bitmap (one process per GPU, same node, NVLink) ------------------------
 * Each rank creates a result buffer for the GLOBAL bitmap (total_words) and exports a 64-byte CUDA-IPC handle; the host
 * exchanges handles (e.g. torch.distributed.all_gather_object) and opens every peer's.  hs_peer_next() then arms the next
 * `_dev` verify call: its finish kernel stores each bitmap word it produces directly into EVERY rank's buffer at
 * word_offset (P2P stores over NVLink), signals the peers and waits for theirs — after the call (stream order) the buffer
 * returned by hs_peer_bitmap() holds every rank's verdicts for that epoch.  Epochs must increase by one per armed call.
 * The buffer is double-buffered by epoch parity (hs_peer_bitmap() follows the most recently armed epoch): consume epoch e's
 * bitmap on the same stream before enqueueing the verify of epoch e+1 and no barrier between ranks is needed.  The flag
 * exchange runs inside the finish kernel (no extra launches).  total_words must be world x (words per rank); a peer that
 * never signals makes hs_peer_timed_out() return 1 and its shard read as all-rejected.  hs_peer_setup may be called again
 * (new total_words): the old buffers are released — every rank must have drained its stream and re-exchange handles. */
int hs_peer_setup(hs_ctx *ctx, int rank, int world, size_t total_words, uint8_t handle_out[64]);
int hs_peer_open(hs_ctx *ctx, int peer_rank, const uint8_t handle[64]);
int hs_peer_next(hs_ctx *ctx, size_t word_offset, uint32_t epoch);
void *hs_peer_bitmap(hs_ctx *ctx);
int hs_peer_timed_out(hs_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
