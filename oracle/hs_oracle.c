/*
 * hs_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY (see hs_oracle.h).
 *
 * What it restates (the arithmetic is in the third-party crate ed25519-dalek 1.0.1, pulled in at
 * /root/reference/crypto/Cargo.toml:10 and not vendored; transitive curve25519-dalek 3.x, sha2 0.9):
 *   - Signature::verify        /root/reference/crypto/src/lib.rs:200-204  -> dalek PublicKey::verify_strict
 *   - Signature::verify_batch  /root/reference/crypto/src/lib.rs:206-219  -> per-signature cofactorless equation
 *   - Signature::new           /root/reference/crypto/src/lib.rs:185-191  -> RFC 8032 sign
 *   - generate_keypair         /root/reference/crypto/src/lib.rs:167-175  -> RFC 8032 key derivation
 *   - Digest = SHA-512[..32]   /root/reference/mempool/src/processor.rs:30, consensus/src/messages.rs:81-208
 * Decision procedures follow SURVEY.md Appendix A (dalek 1.0.1 semantics):
 *   parse: S must be < l; A must decompress (non-canonical y accepted, "x=0 with sign bit" accepted);
 *   strict: R must decompress, R and A must not be small order, k = SHA512(R||A||M) mod l,
 *           accept iff [S]B + [k](-A) == R as projective points (no cofactor).
 *
 * PARITY UNPINNED (in the task's strict sense): the reference holds no golden vectors for this path and dalek cannot be
 * built or run here (no Rust toolchain, crate not vendored).  What the oracle IS pinned on: RFC 8032 §7.1 vectors, FIPS 180-4
 * vectors, OpenSSL and libsodium differentials, the fixtures derived from the reference's own tests, and
 * dalek's PUBLISHED verdicts for the 12 ed25519-speccheck case classes — see tests/test_oracle_pins.py and oracle/README.md.
 *
 * Representation here is deliberately different from the CUDA code (5x51-bit limbs, __int128) so the
 * two implementations do not share bugs.
 */
#include "hs_oracle.h"
#include "hs_constants.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ SHA-512 */
static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static inline uint64_t load_be64(const uint8_t *p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
  return v;
}
static inline void store_be64(uint8_t *p, uint64_t v) {
  for (int i = 7; i >= 0; i--) { p[i] = (uint8_t)v; v >>= 8; }
}

static void sha512_block(uint64_t st[8], const uint8_t blk[128]) {
  uint64_t w[80];
  for (int i = 0; i < 16; i++) w[i] = load_be64(blk + 8 * i);
  for (int i = 16; i < 80; i++) {
    uint64_t s0 = rotr64(w[i - 15], 1) ^ rotr64(w[i - 15], 8) ^ (w[i - 15] >> 7);
    uint64_t s1 = rotr64(w[i - 2], 19) ^ rotr64(w[i - 2], 61) ^ (w[i - 2] >> 6);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
  for (int i = 0; i < 80; i++) {
    uint64_t S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
    uint64_t ch = (e & f) ^ (~e & g);
    uint64_t t1 = h + S1 + ch + HS_SHA512_K[i] + w[i];
    uint64_t S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
    uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint64_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

typedef struct { uint64_t st[8]; uint8_t buf[128]; size_t fill; uint64_t total; } sha512_ctx;
static void sha512_init(sha512_ctx *c) { memcpy(c->st, HS_SHA512_H0, 64); c->fill = 0; c->total = 0; }
static void sha512_update(sha512_ctx *c, const uint8_t *m, size_t len) {
  c->total += len;
  while (len) {
    size_t take = 128 - c->fill;
    if (take > len) take = len;
    memcpy(c->buf + c->fill, m, take);
    c->fill += take; m += take; len -= take;
    if (c->fill == 128) { sha512_block(c->st, c->buf); c->fill = 0; }
  }
}
static void sha512_final(sha512_ctx *c, uint8_t out[64]) {
  uint64_t bits = c->total * 8;
  c->buf[c->fill++] = 0x80;
  if (c->fill > 112) { memset(c->buf + c->fill, 0, 128 - c->fill); sha512_block(c->st, c->buf); c->fill = 0; }
  memset(c->buf + c->fill, 0, 120 - c->fill);
  store_be64(c->buf + 120, bits); /* message lengths here are < 2^61 bytes; the high 64 length bits are zero */
  sha512_block(c->st, c->buf);
  for (int i = 0; i < 8; i++) store_be64(out + 8 * i, c->st[i]);
}
void hso_sha512(const uint8_t *msg, size_t len, uint8_t out[64]) {
  sha512_ctx c; sha512_init(&c); sha512_update(&c, msg, len); sha512_final(&c, out);
}
void hso_digest32(const uint8_t *msg, size_t len, uint8_t out[32]) {
  uint8_t h[64]; hso_sha512(msg, len, h); memcpy(out, h, 32);
}
void hso_digest32_batch(const uint8_t *data, const uint64_t *off, size_t n, uint8_t *out) {
  for (size_t i = 0; i < n; i++) hso_digest32(data + off[i], (size_t)(off[i + 1] - off[i]), out + 32 * i);
}
typedef struct { const uint8_t *data; const uint64_t *off; uint8_t *out; size_t lo, hi; } djob_t;
static void *djob_run(void *arg) {
  djob_t *j = (djob_t *)arg;
  for (size_t i = j->lo; i < j->hi; i++) hso_digest32(j->data + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->out + 32 * i);
  return NULL;
}
void hso_digest32_batch_mt(const uint8_t *data, const uint64_t *off, size_t n, int nthreads, uint8_t *out) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; djob_t jobs[256]; int started = 0; size_t per = (n + (size_t)nthreads - 1) / (size_t)nthreads;
  for (int t = 0; t < nthreads; t++) {
    size_t lo = (size_t)t * per, hi = lo + per; if (lo >= n) break; if (hi > n) hi = n;
    jobs[t] = (djob_t){data, off, out, lo, hi};
    if (nthreads == 1) djob_run(&jobs[t]); else pthread_create(&th[t], NULL, djob_run, &jobs[t]);
    started++;
  }
  if (nthreads > 1) for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ GF(2^255-19), 5 x 51-bit limbs */
typedef struct { uint64_t v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static void fe_from_limbs(fe *r, const uint64_t l[5]) { memcpy(r->v, l, 40); }
static void fe_0(fe *r) { memset(r, 0, sizeof *r); }
static void fe_1(fe *r) { fe_0(r); r->v[0] = 1; }

/* dalek FieldElement::from_bytes: bit 255 ignored, value NOT checked for canonicity. */
static void fe_frombytes(fe *r, const uint8_t s[32]) {
  uint64_t w[4];
  for (int i = 0; i < 4; i++) { w[i] = 0; for (int j = 7; j >= 0; j--) w[i] = (w[i] << 8) | s[8 * i + j]; }
  r->v[0] = w[0] & M51;
  r->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
  r->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
  r->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
  r->v[4] = (w[3] >> 12) & M51;
}
static void fe_carry(fe *r) {
  for (int k = 0; k < 2; k++) {
    for (int i = 0; i < 4; i++) { r->v[i + 1] += r->v[i] >> 51; r->v[i] &= M51; }
    r->v[0] += 19 * (r->v[4] >> 51); r->v[4] &= M51;
  }
}
/* canonical little-endian encoding */
static void fe_tobytes(uint8_t s[32], const fe *a) {
  fe t = *a; fe_carry(&t); fe_carry(&t);
  /* now t < 2^255 + small; subtract p if t >= p: compute t + 19 and see whether bit 255 sets */
  uint64_t q = (t.v[0] + 19) >> 51;
  q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
  t.v[0] += 19 * q;
  for (int i = 0; i < 4; i++) { t.v[i + 1] += t.v[i] >> 51; t.v[i] &= M51; }
  t.v[4] &= M51;
  uint64_t w[4];
  w[0] = t.v[0] | (t.v[1] << 51);
  w[1] = (t.v[1] >> 13) | (t.v[2] << 38);
  w[2] = (t.v[2] >> 26) | (t.v[3] << 25);
  w[3] = (t.v[3] >> 39) | (t.v[4] << 12);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}
static void fe_add(fe *r, const fe *a, const fe *b) {
  for (int i = 0; i < 5; i++) r->v[i] = a->v[i] + b->v[i];
  fe_carry(r);
}
static void fe_sub(fe *r, const fe *a, const fe *b) {
  /* a + 8p - b keeps every limb non-negative for carried inputs (< 2^52) */
  r->v[0] = a->v[0] + 0x3FFFFFFFFFFF68ULL - b->v[0];
  for (int i = 1; i < 5; i++) r->v[i] = a->v[i] + 0x3FFFFFFFFFFFF8ULL - b->v[i];
  fe_carry(r);
}
static void fe_neg(fe *r, const fe *a) { fe z; fe_0(&z); fe_sub(r, &z, a); }
static void fe_mul(fe *r, const fe *a, const fe *b) {
  const uint64_t *x = a->v, *y = b->v;
  uint64_t y1 = 19 * y[1], y2 = 19 * y[2], y3 = 19 * y[3], y4 = 19 * y[4];
  u128 t0 = (u128)x[0] * y[0] + (u128)x[1] * y4 + (u128)x[2] * y3 + (u128)x[3] * y2 + (u128)x[4] * y1;
  u128 t1 = (u128)x[0] * y[1] + (u128)x[1] * y[0] + (u128)x[2] * y4 + (u128)x[3] * y3 + (u128)x[4] * y2;
  u128 t2 = (u128)x[0] * y[2] + (u128)x[1] * y[1] + (u128)x[2] * y[0] + (u128)x[3] * y4 + (u128)x[4] * y3;
  u128 t3 = (u128)x[0] * y[3] + (u128)x[1] * y[2] + (u128)x[2] * y[1] + (u128)x[3] * y[0] + (u128)x[4] * y4;
  u128 t4 = (u128)x[0] * y[4] + (u128)x[1] * y[3] + (u128)x[2] * y[2] + (u128)x[3] * y[1] + (u128)x[4] * y[0];
  t1 += (uint64_t)(t0 >> 51); r->v[0] = (uint64_t)t0 & M51;
  t2 += (uint64_t)(t1 >> 51); r->v[1] = (uint64_t)t1 & M51;
  t3 += (uint64_t)(t2 >> 51); r->v[2] = (uint64_t)t2 & M51;
  t4 += (uint64_t)(t3 >> 51); r->v[3] = (uint64_t)t3 & M51;
  r->v[0] += 19 * (uint64_t)(t4 >> 51); r->v[4] = (uint64_t)t4 & M51;
  r->v[1] += r->v[0] >> 51; r->v[0] &= M51;
}
static void fe_sq(fe *r, const fe *a) { /* dedicated squaring: 15 products instead of 25 */
  const uint64_t *x = a->v;
  uint64_t d0 = 2 * x[0], d1 = 2 * x[1], d2 = 2 * x[2], x3_19 = 19 * x[3], x4_19 = 19 * x[4];
  u128 t0 = (u128)x[0] * x[0] + (u128)d1 * x4_19 + (u128)d2 * x3_19;
  u128 t1 = (u128)d0 * x[1] + (u128)d2 * x4_19 + (u128)x[3] * x3_19;
  u128 t2 = (u128)d0 * x[2] + (u128)x[1] * x[1] + (u128)(2 * x[3]) * x4_19;
  u128 t3 = (u128)d0 * x[3] + (u128)d1 * x[2] + (u128)x[4] * x4_19;
  u128 t4 = (u128)d0 * x[4] + (u128)d1 * x[3] + (u128)x[2] * x[2];
  t1 += (uint64_t)(t0 >> 51); r->v[0] = (uint64_t)t0 & M51;
  t2 += (uint64_t)(t1 >> 51); r->v[1] = (uint64_t)t1 & M51;
  t3 += (uint64_t)(t2 >> 51); r->v[2] = (uint64_t)t2 & M51;
  t4 += (uint64_t)(t3 >> 51); r->v[3] = (uint64_t)t3 & M51;
  r->v[0] += 19 * (uint64_t)(t4 >> 51); r->v[4] = (uint64_t)t4 & M51;
  r->v[1] += r->v[0] >> 51; r->v[0] &= M51;
}
static void fe_sqn(fe *r, const fe *a, int n) { fe_sq(r, a); for (int i = 1; i < n; i++) fe_sq(r, r); }

/* z^(2^250 - 1) and z^11 through the usual 2,9,11,2^5-1,2^10-1,... ladder */
static void fe_pow2_250_1(fe *out, fe *z11_out, const fe *z) {
  fe z2, z9, z11, t, z2_5, z2_10, z2_20, z2_40, z2_50, z2_100, z2_200;
  fe_sq(&z2, z);
  fe_sqn(&t, &z2, 2); fe_mul(&z9, &t, z);
  fe_mul(&z11, &z9, &z2);
  fe_sq(&t, &z11); fe_mul(&z2_5, &t, &z9);
  fe_sqn(&t, &z2_5, 5); fe_mul(&z2_10, &t, &z2_5);
  fe_sqn(&t, &z2_10, 10); fe_mul(&z2_20, &t, &z2_10);
  fe_sqn(&t, &z2_20, 20); fe_mul(&z2_40, &t, &z2_20);
  fe_sqn(&t, &z2_40, 10); fe_mul(&z2_50, &t, &z2_10);
  fe_sqn(&t, &z2_50, 50); fe_mul(&z2_100, &t, &z2_50);
  fe_sqn(&t, &z2_100, 100); fe_mul(&z2_200, &t, &z2_100);
  fe_sqn(&t, &z2_200, 50); fe_mul(out, &t, &z2_50);
  if (z11_out) *z11_out = z11;
}
static void fe_invert(fe *r, const fe *z) { /* z^(p-2) = z^(2^255-21) */
  fe t, z11; fe_pow2_250_1(&t, &z11, z); fe_sqn(&t, &t, 5); fe_mul(r, &t, &z11);
}
static void fe_pow_p58(fe *r, const fe *z) { /* z^((p-5)/8) = z^(2^252-3) */
  fe t; fe_pow2_250_1(&t, NULL, z); fe_sqn(&t, &t, 2); fe_mul(r, &t, z);
}
static int fe_iszero(const fe *a) {
  uint8_t s[32]; fe_tobytes(s, a); uint8_t acc = 0;
  for (int i = 0; i < 32; i++) acc |= s[i];
  return acc == 0;
}
static int fe_eq(const fe *a, const fe *b) { fe t; fe_sub(&t, a, b); return fe_iszero(&t); }
static int fe_isneg(const fe *a) { uint8_t s[32]; fe_tobytes(s, a); return s[0] & 1; }

/* dalek FieldElement::sqrt_ratio_i(u, v): returns 1 and r = +sqrt(u/v) (the "non-negative" root) when u/v is
 * square; 0 otherwise (r is then sqrt(i*u/v), unused here). */
static int fe_sqrt_ratio_i(fe *r, const fe *u, const fe *v) {
  fe v3, v7, t, chk, neg_u, neg_u_i, sqm1;
  fe_from_limbs(&sqm1, HS_FE_SQRTM1_51);
  fe_sq(&t, v); fe_mul(&v3, &t, v);           /* v^3 */
  fe_sq(&t, &v3); fe_mul(&v7, &t, v);         /* v^7 */
  fe_mul(&t, u, &v7); fe_pow_p58(&t, &t);     /* (u v^7)^((p-5)/8) */
  fe_mul(&t, &t, &v3); fe_mul(r, &t, u);      /* r = u v^3 (u v^7)^((p-5)/8) */
  fe_sq(&t, r); fe_mul(&chk, &t, v);          /* v r^2 */
  fe_neg(&neg_u, u); fe_mul(&neg_u_i, &neg_u, &sqm1);
  int correct = fe_eq(&chk, u), flipped = fe_eq(&chk, &neg_u), flipped_i = fe_eq(&chk, &neg_u_i);
  if (flipped || flipped_i) { fe_mul(&t, r, &sqm1); *r = t; }
  if (fe_isneg(r)) fe_neg(r, r);
  return correct || flipped;
}

/* ------------------------------------------------------------------ Edwards points, extended coordinates */
typedef struct { fe X, Y, Z, T; } ge;
static void ge_identity(ge *p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }
static void ge_base(ge *p) {
  fe_from_limbs(&p->X, HS_FE_BX_51); fe_from_limbs(&p->Y, HS_FE_BY_51); fe_1(&p->Z); fe_mul(&p->T, &p->X, &p->Y);
}
/* dalek CompressedEdwardsY::decompress — tolerant rules. */
static int ge_decompress(ge *p, const uint8_t s[32]) {
  fe y, yy, u, v, d, one, x;
  fe_from_limbs(&d, HS_FE_D_51); fe_1(&one);
  fe_frombytes(&y, s);
  fe_sq(&yy, &y); fe_sub(&u, &yy, &one); fe_mul(&v, &yy, &d); fe_add(&v, &v, &one);
  if (!fe_sqrt_ratio_i(&x, &u, &v)) return 0;
  if (s[31] >> 7) fe_neg(&x, &x); /* conditional_negate(sign): -0 == 0 is accepted */
  p->X = x; p->Y = y; fe_1(&p->Z); fe_mul(&p->T, &x, &y);
  return 1;
}
static void ge_compress(uint8_t s[32], const ge *p) {
  fe zi, x, y; fe_invert(&zi, &p->Z); fe_mul(&x, &p->X, &zi); fe_mul(&y, &p->Y, &zi);
  fe_tobytes(s, &y); s[31] ^= (uint8_t)(fe_isneg(&x) << 7);
}
/* complete unified addition on -x^2+y^2 = 1+d x^2 y^2 (valid for doubling, identity, torsion points) */
static void ge_add(ge *r, const ge *p, const ge *q) {
  fe a, b, c, dd, e, f, g, h, t, d2;
  fe_from_limbs(&d2, HS_FE_D2_51);
  fe_sub(&a, &p->Y, &p->X); fe_sub(&t, &q->Y, &q->X); fe_mul(&a, &a, &t);
  fe_add(&b, &p->Y, &p->X); fe_add(&t, &q->Y, &q->X); fe_mul(&b, &b, &t);
  fe_mul(&c, &p->T, &q->T); fe_mul(&c, &c, &d2);
  fe_mul(&dd, &p->Z, &q->Z); fe_add(&dd, &dd, &dd);
  fe_sub(&e, &b, &a); fe_sub(&f, &dd, &c); fe_add(&g, &dd, &c); fe_add(&h, &b, &a);
  fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}
static void ge_neg(ge *r, const ge *p) { fe_neg(&r->X, &p->X); r->Y = p->Y; r->Z = p->Z; fe_neg(&r->T, &p->T); }
/* dedicated doubling (a = -1): A=X^2 B=Y^2 C=2Z^2 H=A+B E=H-(X+Y)^2 G=A-B F=C+G */
static void ge_dbl(ge *r, const ge *p) {
  fe a, b, c, e, f, g, h, t;
  fe_sq(&a, &p->X); fe_sq(&b, &p->Y); fe_sq(&c, &p->Z); fe_add(&c, &c, &c);
  fe_add(&h, &a, &b); fe_add(&t, &p->X, &p->Y); fe_sq(&t, &t); fe_sub(&e, &h, &t);
  fe_sub(&g, &a, &b); fe_add(&f, &c, &g);
  fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->T, &e, &h); fe_mul(&r->Z, &f, &g);
}
/* projective equality: X1 Z2 == X2 Z1 and Y1 Z2 == Y2 Z1 (dalek EdwardsPoint ct_eq) */
static int ge_eq(const ge *p, const ge *q) {
  fe a, b; int ok;
  fe_mul(&a, &p->X, &q->Z); fe_mul(&b, &q->X, &p->Z); ok = fe_eq(&a, &b);
  fe_mul(&a, &p->Y, &q->Z); fe_mul(&b, &q->Y, &p->Z); return ok & fe_eq(&a, &b);
}
static int ge_is_identity(const ge *p) { ge id; ge_identity(&id); return ge_eq(p, &id); }
/* dalek is_small_order: mul_by_cofactor().is_identity() */
static int ge_is_small_order(const ge *p) {
  ge t; ge_dbl(&t, p); ge_dbl(&t, &t); ge_dbl(&t, &t); return ge_is_identity(&t);
}
/* plain MSB-first double-and-add with the complete addition: the slow, obviously-right path */
static void ge_scalarmult_simple(ge *r, const uint8_t sc[32], const ge *p) {
  ge acc; ge_identity(&acc);
  for (int i = 255; i >= 0; i--) {
    ge_add(&acc, &acc, &acc);
    if ((sc[i >> 3] >> (i & 7)) & 1) ge_add(&acc, &acc, p);
  }
  *r = acc;
}

/* ------------------------------------------------------------------ scalars mod l (64-bit limbs, Barrett) */
static void limbs_from_le(uint64_t *w, int nw, const uint8_t *s, int nbytes) {
  for (int i = 0; i < nw; i++) w[i] = 0;
  for (int i = 0; i < nbytes; i++) w[i >> 3] |= (uint64_t)s[i] << (8 * (i & 7));
}
static void limbs_to_le(uint8_t *s, const uint64_t *w, int nbytes) {
  for (int i = 0; i < nbytes; i++) s[i] = (uint8_t)(w[i >> 3] >> (8 * (i & 7)));
}
static void limbs_mul(uint64_t *out, const uint64_t *a, int na, const uint64_t *b, int nb) {
  for (int i = 0; i < na + nb; i++) out[i] = 0;
  for (int i = 0; i < na; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < nb; j++) {
      u128 t = (u128)a[i] * b[j] + out[i + j] + carry;
      out[i + j] = (uint64_t)t; carry = (uint64_t)(t >> 64);
    }
    out[i + nb] = carry;
  }
}
static int limbs_geq(const uint64_t *a, const uint64_t *b, int n) {
  for (int i = n - 1; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
  return 1;
}
static void limbs_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
  uint64_t borrow = 0;
  for (int i = 0; i < n; i++) {
    u128 t = (u128)a[i] - b[i] - borrow; r[i] = (uint64_t)t; borrow = (uint64_t)(t >> 64) & 1;
  }
}
/* x (8 limbs, < 2^512) mod l -> 4 limbs. Barrett with mu = floor(2^512 / l). (dalek Scalar::from_bytes_mod_order_wide) */
static void sc_reduce_limbs(uint64_t r[4], const uint64_t x[8]) {
  uint64_t q2[10], q3l[10], l5[5] = {HS_SC_L_64[0], HS_SC_L_64[1], HS_SC_L_64[2], HS_SC_L_64[3], 0}, rr[5], t[5];
  limbs_mul(q2, x + 3, 5, HS_SC_MU_64, 5);  /* floor(x / 2^192) * mu */
  limbs_mul(q3l, q2 + 5, 5, l5, 5);         /* floor(q2 / 2^320) * l   (only low 5 limbs needed) */
  limbs_sub(rr, x, q3l, 5);                 /* x - q3*l mod 2^320; true value < 3l */
  for (int k = 0; k < 3; k++) if (limbs_geq(rr, l5, 5)) { limbs_sub(t, rr, l5, 5); memcpy(rr, t, 40); }
  memcpy(r, rr, 32);
}
void hso_sc_reduce64(const uint8_t in[64], uint8_t out[32]) {
  uint64_t x[8], r[4]; limbs_from_le(x, 8, in, 64); sc_reduce_limbs(r, x); limbs_to_le(out, r, 32);
}
void hso_sc_muladd(const uint8_t a[32], const uint8_t b[32], const uint8_t c[32], uint8_t out[32]) {
  uint64_t x[4], y[4], z[4], prod[8], r[4];
  limbs_from_le(x, 4, a, 32); limbs_from_le(y, 4, b, 32); limbs_from_le(z, 4, c, 32);
  limbs_mul(prod, x, 4, y, 4);
  uint64_t carry = 0;
  for (int i = 0; i < 8; i++) { u128 t = (u128)prod[i] + (i < 4 ? z[i] : 0) + carry; prod[i] = (uint64_t)t; carry = (uint64_t)(t >> 64); }
  sc_reduce_limbs(r, prod); limbs_to_le(out, r, 32);
}
/* dalek check_scalar: accept iff S < l */
static int sc_is_canonical(const uint8_t s[32]) {
  uint64_t x[4]; limbs_from_le(x, 4, s, 32);
  return !limbs_geq(x, HS_SC_L_64, 4);
}

/* ------------------------------------------------------------------ fixed-base [s]B for input synthesis */
/* radix-16 unsigned comb: g_comb[i][j] = j * 16^i * B.  Only keygen/sign use it (bulk fixture generation);
 * tests check it against ge_scalarmult_simple. */
static ge g_comb[64][16];
static pthread_once_t g_comb_once = PTHREAD_ONCE_INIT;
static void comb_init(void) {
  ge base; ge_base(&base);
  for (int i = 0; i < 64; i++) {
    ge_identity(&g_comb[i][0]);
    for (int j = 1; j < 16; j++) ge_add(&g_comb[i][j], &g_comb[i][j - 1], &base);
    for (int k = 0; k < 4; k++) ge_dbl(&base, &base);
  }
}
static void ge_scalarmult_base(ge *r, const uint8_t sc[32]) {
  pthread_once(&g_comb_once, comb_init);
  ge acc; ge_identity(&acc);
  for (int i = 0; i < 64; i++) {
    int nib = (sc[i >> 1] >> ((i & 1) * 4)) & 15;
    if (nib) ge_add(&acc, &acc, &g_comb[i][nib]);
  }
  *r = acc;
}

/* ------------------------------------------------------------------ keygen / sign (RFC 8032 §5.1.5-5.1.6) */
static void expand_seed(const uint8_t seed[32], uint8_t a[32], uint8_t prefix[32]) {
  uint8_t h[64]; hso_sha512(seed, 32, h);
  h[0] &= 248; h[31] &= 127; h[31] |= 64;
  memcpy(a, h, 32); memcpy(prefix, h + 32, 32);
}
void hso_keygen(const uint8_t seed[32], uint8_t pk[32]) {
  uint8_t a[32], prefix[32]; ge B, A; expand_seed(seed, a, prefix);
  (void)B; ge_scalarmult_base(&A, a); ge_compress(pk, &A);
}
static void sign_with_pk(const uint8_t seed[32], const uint8_t pk[32], const uint8_t *msg, size_t len, uint8_t sig[64]);
void hso_sign(const uint8_t seed[32], const uint8_t *msg, size_t len, uint8_t sig[64]) {
  uint8_t pk[32]; hso_keygen(seed, pk); sign_with_pk(seed, pk, msg, len, sig);
}
static void sign_with_pk(const uint8_t seed[32], const uint8_t pk[32], const uint8_t *msg, size_t len, uint8_t sig[64]) {
  uint8_t a[32], prefix[32], h[64], r[32], k[32]; ge P; sha512_ctx c;
  expand_seed(seed, a, prefix);
  sha512_init(&c); sha512_update(&c, prefix, 32); sha512_update(&c, msg, len); sha512_final(&c, h);
  hso_sc_reduce64(h, r);
  ge_scalarmult_base(&P, r); ge_compress(sig, &P);
  sha512_init(&c); sha512_update(&c, sig, 32); sha512_update(&c, pk, 32); sha512_update(&c, msg, len); sha512_final(&c, h);
  hso_sc_reduce64(h, k);
  hso_sc_muladd(k, a, r, sig + 32);
}

/* ------------------------------------------------------------------ verification */
/* Fast vartime double-scalar multiplication: width-5 signed sliding window (NAF) for the variable point
 * and for B — the shape of dalek's vartime_double_scalar_mul_basepoint. */
static void sc_naf(int8_t naf[257], const uint8_t s[32], int w) {
  uint64_t x[5] = {0, 0, 0, 0, 0}; limbs_from_le(x, 4, s, 32);
  int width = 1 << w, pos = 0;
  memset(naf, 0, 257);
  while (pos < 257) {
    int idx = pos >> 6, bit = pos & 63;
    uint64_t chunk = x[idx] >> bit;
    if (bit && idx < 4) chunk |= x[idx + 1] << (64 - bit);
    if (!(chunk & 1)) { pos++; continue; }
    int win = (int)(chunk & (uint64_t)(width - 1));
    if (win >= width / 2) {
      win -= width;
      /* add 2^pos * width back: propagate +1 at bit pos+w */
      int p2 = pos + w, i2 = p2 >> 6; uint64_t add = 1ULL << (p2 & 63);
      /* first clear the window bits */
      uint64_t mask = (uint64_t)(width - 1);
      x[idx] &= ~(mask << bit);
      if (bit + w > 64 && idx < 4) x[idx + 1] &= ~(mask >> (64 - bit));
      while (i2 < 5) { uint64_t o = x[i2]; x[i2] += add; if (x[i2] >= o) break; add = 1; i2++; }
    } else {
      uint64_t mask = (uint64_t)(width - 1);
      x[idx] &= ~(mask << bit);
      if (bit + w > 64 && idx < 4) x[idx + 1] &= ~(mask >> (64 - bit));
    }
    naf[pos] = (int8_t)win;
    pos += w;
  }
}
static ge g_btab[8]; /* 1B,3B,...,15B */
static pthread_once_t g_btab_once = PTHREAD_ONCE_INIT;
static void btab_init(void) {
  ge B, B2; ge_base(&B); ge_dbl(&B2, &B); g_btab[0] = B;
  for (int i = 1; i < 8; i++) ge_add(&g_btab[i], &g_btab[i - 1], &B2);
}
/* r = [a]P + [b]B */
static void ge_double_scalarmult_vartime(ge *r, const uint8_t a[32], const ge *P, const uint8_t b[32]) {
  int8_t na[257], nb[257]; ge tab[8], P2, acc, t;
  pthread_once(&g_btab_once, btab_init);
  sc_naf(na, a, 5); sc_naf(nb, b, 5);
  tab[0] = *P; ge_dbl(&P2, P);
  for (int i = 1; i < 8; i++) ge_add(&tab[i], &tab[i - 1], &P2);
  int i = 256;
  while (i >= 0 && !na[i] && !nb[i]) i--;
  ge_identity(&acc);
  for (; i >= 0; i--) {
    ge_dbl(&acc, &acc);
    if (na[i] > 0) ge_add(&acc, &acc, &tab[na[i] >> 1]);
    else if (na[i] < 0) { ge_neg(&t, &tab[(-na[i]) >> 1]); ge_add(&acc, &acc, &t); }
    if (nb[i] > 0) ge_add(&acc, &acc, &g_btab[nb[i] >> 1]);
    else if (nb[i] < 0) { ge_neg(&t, &g_btab[(-nb[i]) >> 1]); ge_add(&acc, &acc, &t); }
  }
  *r = acc;
}

static unsigned verify_flags_impl(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len, int fast) {
  unsigned fl = 0; ge A, R, negA, Rp; uint8_t h[64], k[32]; sha512_ctx c;
  int s_ok = sc_is_canonical(sig + 32);       /* Signature::from_bytes, lib.rs:201 */
  int a_ok = ge_decompress(&A, pk);           /* PublicKey::from_bytes, lib.rs:202 */
  if (s_ok && a_ok) fl |= HSO_PARSE_OK;
  int r_ok = ge_decompress(&R, sig);
  if (r_ok) fl |= HSO_R_OK;
  if ((r_ok && ge_is_small_order(&R)) || (a_ok && ge_is_small_order(&A))) fl |= HSO_SMALL;
  if (!(s_ok && a_ok && r_ok)) return fl;
  sha512_init(&c); sha512_update(&c, sig, 32); sha512_update(&c, pk, 32); sha512_update(&c, msg, len); sha512_final(&c, h);
  hso_sc_reduce64(h, k);
  ge_neg(&negA, &A);
  if (fast) ge_double_scalarmult_vartime(&Rp, k, &negA, sig + 32);
  else { ge B, t1, t2; ge_base(&B); ge_scalarmult_simple(&t1, k, &negA); ge_scalarmult_simple(&t2, sig + 32, &B); ge_add(&Rp, &t1, &t2); }
  if (ge_eq(&Rp, &R)) fl |= HSO_EQ_OK;
  if ((fl & HSO_EQ_OK) && !(fl & HSO_SMALL)) fl |= HSO_STRICT;
  return fl;
}
unsigned hso_verify_flags(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len) {
  return verify_flags_impl(sig, pk, msg, len, 0);
}
unsigned hso_verify_flags_fast(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len) {
  return verify_flags_impl(sig, pk, msg, len, 1);
}
int hso_verify_strict(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, size_t len) {
  return (verify_flags_impl(sig, pk, msg, len, 1) & HSO_STRICT) != 0;
}

/* ------------------------------------------------------------------ batch drivers (pthreads) */
typedef struct {
  const uint8_t *sig, *pk, *msgs; const uint64_t *off; size_t sig_stride, pk_stride, msg_stride, msg_len;
  size_t lo, hi; int mode; uint32_t *bitmap;
} job_t;
static void *job_run(void *arg) {
  job_t *j = (job_t *)arg;
  for (size_t w = j->lo; w < j->hi; w += 32) { /* lo is a multiple of 32: words are thread-private */
    uint32_t word = 0;
    for (size_t i = w; i < w + 32 && i < j->hi; i++) {
      const uint8_t *m = j->off ? j->msgs + j->off[i] : j->msgs + i * j->msg_stride;
      size_t ml = j->off ? (size_t)(j->off[i + 1] - j->off[i]) : j->msg_len;
      unsigned fl = verify_flags_impl(j->sig + i * j->sig_stride, j->pk + i * j->pk_stride, m, ml, 1);
      unsigned ok = j->mode == 0 ? (fl & HSO_STRICT) : (fl & HSO_EQ_OK);
      if (ok) word |= 1u << (i & 31);
    }
    j->bitmap[w >> 5] = word;
  }
  return NULL;
}
static void run_jobs(job_t proto, size_t n, int nthreads, uint32_t *bitmap) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  size_t words = (n + 31) / 32, per = (words + (size_t)nthreads - 1) / (size_t)nthreads;
  pthread_t th[256]; job_t jobs[256]; int started = 0;
  for (int t = 0; t < nthreads; t++) {
    size_t lo = (size_t)t * per * 32, hi = lo + per * 32;
    if (lo >= n) break;
    if (hi > n) hi = n;
    jobs[t] = proto; jobs[t].lo = lo; jobs[t].hi = hi; jobs[t].bitmap = bitmap;
    if (nthreads == 1) job_run(&jobs[t]);
    else { pthread_create(&th[t], NULL, job_run, &jobs[t]); }
    started++;
  }
  if (nthreads > 1) for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}
void hso_verify_rec128_batch(const uint8_t *recs, size_t n, int mode, int nthreads, uint32_t *bitmap) {
  job_t j; memset(&j, 0, sizeof j);
  j.sig = recs; j.pk = recs + 64; j.msgs = recs + 96; j.sig_stride = j.pk_stride = j.msg_stride = 128; j.msg_len = 32; j.mode = mode;
  run_jobs(j, n, nthreads, bitmap);
}
void hso_verify_var_batch(const uint8_t *sig, const uint8_t *pk, const uint8_t *msgs, const uint64_t *off,
                          size_t n, int mode, int nthreads, uint32_t *bitmap) {
  job_t j; memset(&j, 0, sizeof j);
  j.sig = sig; j.pk = pk; j.msgs = msgs; j.off = off; j.sig_stride = 64; j.pk_stride = 32; j.mode = mode;
  run_jobs(j, n, nthreads, bitmap);
}
int hso_verify_batch_shared_msg(const uint8_t digest[32], const uint8_t *votes, size_t n, int nthreads, uint32_t *bitmap_or_null) {
  size_t words = (n + 31) / 32;
  uint32_t *bm = bitmap_or_null ? bitmap_or_null : (uint32_t *)calloc(words ? words : 1, 4);
  job_t j; memset(&j, 0, sizeof j);
  j.pk = votes; j.sig = votes + 32; j.pk_stride = j.sig_stride = 96; j.msgs = digest; j.msg_stride = 0; j.msg_len = 32; j.mode = 1;
  run_jobs(j, n, nthreads, bm);
  int all = 1;
  for (size_t i = 0; i < n; i++) if (!((bm[i >> 5] >> (i & 31)) & 1)) { all = 0; break; }
  if (!bitmap_or_null) free(bm);
  return all;
}

/* bulk signing for fixture / benchmark-input synthesis: item i signs msgs[off[i]..off[i+1]) with key key_idx[i] */
typedef struct { const uint8_t *seeds, *pks, *msgs; const uint32_t *key_idx; const uint64_t *off; uint8_t *sigs; size_t lo, hi; } sjob_t;
static void *sjob_run(void *arg) {
  sjob_t *j = (sjob_t *)arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    uint32_t k = j->key_idx[i];
    sign_with_pk(j->seeds + 32 * (size_t)k, j->pks + 32 * (size_t)k, j->msgs + j->off[i], (size_t)(j->off[i + 1] - j->off[i]), j->sigs + 64 * i);
  }
  return NULL;
}
void hso_sign_batch(const uint8_t *seeds, const uint8_t *pks, const uint32_t *key_idx, const uint8_t *msgs, const uint64_t *off,
                    size_t n, int nthreads, uint8_t *sigs) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_once(&g_comb_once, comb_init);
  pthread_t th[256]; sjob_t jobs[256]; int started = 0; size_t per = (n + (size_t)nthreads - 1) / (size_t)nthreads;
  for (int t = 0; t < nthreads; t++) {
    size_t lo = (size_t)t * per, hi = lo + per; if (lo >= n) break; if (hi > n) hi = n;
    jobs[t] = (sjob_t){seeds, pks, msgs, key_idx, off, sigs, lo, hi};
    if (nthreads == 1) sjob_run(&jobs[t]); else pthread_create(&th[t], NULL, sjob_run, &jobs[t]);
    started++;
  }
  if (nthreads > 1) for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}
void hso_keygen_batch(const uint8_t *seeds, size_t n, uint8_t *pks) { for (size_t i = 0; i < n; i++) hso_keygen(seeds + 32 * i, pks + 32 * i); }
/* [s]B through the slow double-and-add, for checking the comb used by keygen/sign */
void hso_scalarmult_base_simple(const uint8_t sc[32], uint8_t out[32]) { ge B, r; ge_base(&B); ge_scalarmult_simple(&r, sc, &B); ge_compress(out, &r); }
void hso_scalarmult_base_comb(const uint8_t sc[32], uint8_t out[32]) { ge r; ge_scalarmult_base(&r, sc); ge_compress(out, &r); }

/* ------------------------------------------------------------------ fixture helpers */
int hso_point_decompress_ok(const uint8_t enc[32]) { ge p; return ge_decompress(&p, enc); }
int hso_point_is_small_order(const uint8_t enc[32]) { ge p; if (!ge_decompress(&p, enc)) return -1; return ge_is_small_order(&p); }
int hso_point_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  ge p, q, r; if (!ge_decompress(&p, a) || !ge_decompress(&q, b)) return 0;
  ge_add(&r, &p, &q); ge_compress(out, &r); return 1;
}
int hso_point_scalarmult(const uint8_t scalar[32], const uint8_t pt[32], uint8_t out[32]) {
  ge p, r; if (!ge_decompress(&p, pt)) return 0;
  ge_scalarmult_simple(&r, scalar, &p); ge_compress(out, &r); return 1;
}
