/*
 * hs_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the Ed25519 / SHA-512 decision procedures that the reference's
 * `crypto` crate obtains from ed25519-dalek 1.0.1 (crypto/Cargo.toml:10 — the crate is NOT
 * vendored under /root/reference and there is no Rust toolchain here, so this restates the
 * crate's published algorithm; see oracle/README.md for how it is pinned).
 *
 * PARITY UNPINNED against a running dalek (cannot be built here); pinned on RFC 8032 / FIPS 180-4 known answers, OpenSSL + libsodium
 * differentials, reference-derived fixtures and dalek's published speccheck verdicts (oracle/README.md).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this. The product path (hotstuff_b200/ + include/hs_crypto.h) never does.
 */
#ifndef HS_ORACLE_H
#define HS_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* SHA-512 (FIPS 180-4). Reference call sites: consensus/src/messages.rs:81,151,203,270,308,
 * mempool/src/processor.rs:30, crypto/src/tests/crypto_tests.rs:8-12. */
void hso_sha512(const uint8_t *msg, size_t len, uint8_t out[64]);
/* Digest = first 32 bytes of SHA-512 (crypto/src/lib.rs:22; processor.rs:30). */
void hso_digest32(const uint8_t *msg, size_t len, uint8_t out[32]);
/* out[i] = digest32(data[off[i] .. off[i+1])) */
void hso_digest32_batch(const uint8_t *data, const uint64_t *off, size_t n, uint8_t *out);
void hso_digest32_batch_mt(const uint8_t *data, const uint64_t *off, size_t n, int nthreads, uint8_t *out);

/* RFC 8032 5.1.5 / dalek Keypair::generate (crypto/src/lib.rs:167-175): pk from 32-byte seed. */
void hso_keygen(const uint8_t seed[32], uint8_t pk[32]);
/* RFC 8032 5.1.6 / dalek Keypair::sign (crypto/src/lib.rs:185-191). */
void hso_sign(const uint8_t seed[32], const uint8_t *msg, size_t len, uint8_t sig[64]);

/* Bulk input synthesis (fixtures, benchmark inputs): item i signs msgs[off[i]..off[i+1]) with key key_idx[i]. */
void hso_keygen_batch(const uint8_t *seeds, size_t n, uint8_t *pks);
void hso_sign_batch(const uint8_t *seeds, const uint8_t *pks, const uint32_t *key_idx, const uint8_t *msgs, const uint64_t *off,
                    size_t n, int nthreads, uint8_t *sigs);
void hso_scalarmult_base_simple(const uint8_t sc[32], uint8_t out[32]);
void hso_scalarmult_base_comb(const uint8_t sc[32], uint8_t out[32]);

/* Per-signature decision bits. */
#define HSO_PARSE_OK 1u   /* S < l, A decompresses (Signature::from_bytes + PublicKey::from_bytes, lib.rs:201-202) */
#define HSO_R_OK 2u       /* R decompresses */
#define HSO_EQ_OK 4u      /* PARSE_OK & R_OK & [S]B - [k]A == R as points (cofactorless) */
#define HSO_SMALL 8u      /* R or A has small order ([8]P == identity) */
#define HSO_STRICT 16u    /* EQ_OK & !SMALL  == dalek verify_strict == Signature::verify (lib.rs:200-204) */
unsigned hso_verify_flags(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len);
/* Same decision through the windowed vartime double-scalar path (the fast path used as CPU baseline). */
unsigned hso_verify_flags_fast(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len);

/* Signature::verify semantics: 1 = Ok, 0 = Err. */
int hso_verify_strict(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len);

/* recs: n packed 128-byte records {sig[64], pk[32], msg[32]}. bit i of bitmap = verdict.
 * mode 0: strict (Signature::verify), mode 1: cofactorless equation only (deterministic restatement
 * of verify_batch's per-signature condition). nthreads >= 1 (pthreads). */
void hso_verify_rec128_batch(const uint8_t *recs, size_t n, int mode, int nthreads, uint32_t *bitmap);
/* Variable-length messages: sig[n][64], pk[n][32], msgs concatenated with off[n+1]. */
void hso_verify_var_batch(const uint8_t *sig, const uint8_t *pk, const uint8_t *msgs, const uint64_t *off,
                          size_t n, int mode, int nthreads, uint32_t *bitmap);
/* Signature::verify_batch shape (lib.rs:206-219): one digest, votes = n x {pk[32], sig[64]}.
 * Returns 1 iff every vote parses and satisfies the cofactorless equation. */
int hso_verify_batch_shared_msg(const uint8_t digest[32], const uint8_t *votes, size_t n, int nthreads,
                                uint32_t *bitmap_or_null);

/* Point helpers used to build adversarial fixtures (tests only). */
int hso_point_decompress_ok(const uint8_t enc[32]);
int hso_point_is_small_order(const uint8_t enc[32]);                    /* -1 if not decompressible */
int hso_point_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]); /* 0 on failure */
int hso_point_scalarmult(const uint8_t scalar[32], const uint8_t pt[32], uint8_t out[32]);
void hso_sc_reduce64(const uint8_t in[64], uint8_t out[32]);
void hso_sc_muladd(const uint8_t a[32], const uint8_t b[32], const uint8_t c[32], uint8_t out[32]); /* a*b+c mod l */

#ifdef __cplusplus
}
#endif
#endif
