// hostemu.cpp — TEST-ONLY host build of the device headers (HS_HOST_EMU): the exact curve / scalar / window /
// table logic of the CUDA kernels compiled by g++ with the PTX field primitives swapped for portable C, so it can
// be checked against the oracle on a machine without a GPU.  Never linked into the product library.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../hotstuff_b200/csrc/verify_core.cuh"

static std::vector<ge_niels> g_btable;
static void build_table(std::vector<ge_niels> &t, const ge_ext &P, int W, int windows) {
  const int entries = 1 << (W - 1), block = 64;
  t.resize((size_t)windows * comb_window_stride(W));
  std::vector<fe> prod(block);
  for (int w = 0; w < windows; w++)
    for (int b = 0; b < entries / block; b++) comb_build_block(t.data(), P, W, w, b * block, block, prod.data());
}
static comb_params g_cp;
// window widths used by the host emulation (small enough to build on a CPU; the logic is width-independent)
extern "C" void emu_set_windows(int wa, int wb) {
  g_cp.wa = wa; g_cp.na = sc_ndigits_rt(wa);
  g_cp.wb = wb; g_cp.nb = sc_ndigits_rt(wb);
  sc_bias_rt(g_cp.bias_a, wa);
  sc_bias_rt(g_cp.bias_b, wb);
  g_btable.clear();
}
static void ensure_btable() {
  if (g_cp.wa == 0) emu_set_windows(10, 12);
  if (!g_btable.empty()) return;
  ge_ext B;
  ge_basepoint(B);
  build_table(g_btable, B, g_cp.wb, g_cp.nb);
}
static unsigned finish_one(const ge_ext &acc, const uint32_t (&R)[8], uint32_t meta) {
  fe zinv;
  ge_ext a = acc;
  if (!(meta & HS_META_PARSE_OK)) { fe_set0(a.X); fe_set1(a.Y); fe_set1(a.Z); }
  fe_invert(zinv, a.Z);
  return verify_flags_from(a.X, a.Y, zinv, R, meta);
}
static void load_words(uint32_t (&w)[8], const uint8_t *p) { memcpy(w, p, 32); }

extern "C" {
void emu_fe_op(int op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  fe x, y, r;
  memcpy(x.v, a, 32);
  memcpy(y.v, b, 32);
  switch (op) {
    case 0: fe_mul(r, x, y); break;
    case 1: fe_sqr(r, x); break;
    case 2: fe_add(r, x, y); break;
    case 3: fe_sub(r, x, y); break;
    case 4: fe_canon(r, x); break;
    case 5: fe_invert(r, x); break;
    case 6: fe_pow_p58(r, x); break;
    case 7: fe_neg(r, x); break;
    default: fe_set0(r);
  }
  memcpy(out, r.v, 32);
}
void emu_sc_reduce512(const uint8_t in[64], uint8_t out[32]) {
  uint32_t x[16], r[8];
  memcpy(x, in, 64);
  sc_reduce512(r, x);
  memcpy(out, r, 32);
}
int emu_sc_is_canonical(const uint8_t s[32]) {
  uint32_t w[8];
  load_words(w, s);
  return (int)sc_is_canonical(w);
}
// digits of the signed recoding: W == 4 with msb != 0 uses the radix-16 msb stream, everything else the runtime recoder
int emu_sc_digits(int W, int msb, const uint8_t s[32], int *out) {
  uint32_t w[8];
  load_words(w, s);
  if (W == 4 && msb) { digits_msb<4> d; d.init(w); for (int i = 0; i < 64; i++) out[63 - i] = d.next(); return 64; }
  uint32_t bias[9];
  sc_bias_rt(bias, W);
  int n = sc_ndigits_rt(W);
  int32_t dig[80];
  sc_digits_rt(dig, 1, w, bias, W, n);
  for (int i = 0; i < n; i++) out[i] = dig[i];
  return n;
}
void emu_sha512(const uint8_t *msg, uint64_t len, uint8_t out[64]) {
  uint32_t o[16];
  uint64_t pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  sha512_prefix_msg(o, pre, 0, msg, len);
  memcpy(out, o, 64);
}
void emu_sha512_ram(const uint8_t R[32], const uint8_t A[32], const uint8_t *msg, uint64_t len, uint8_t out[64]) {
  uint32_t o[16], r[8], a[8];
  load_words(r, R);
  load_words(a, A);
  if (len == 32) {
    uint32_t m[8];
    load_words(m, msg);
    sha512_ram32(o, r, a, m);
  } else {
    uint64_t pre[8];
    for (int i = 0; i < 4; i++) { pre[i] = be64_from_le32(r[2 * i], r[2 * i + 1]); pre[4 + i] = be64_from_le32(a[2 * i], a[2 * i + 1]); }
    sha512_prefix_msg(o, pre, 8, msg, len);
  }
  memcpy(out, o, 64);
}
int emu_decompress(const uint8_t enc[32], uint8_t x_out[32], uint8_t y_out[32]) {
  uint32_t w[8];
  load_words(w, enc);
  ge_ext p;
  uint32_t ok = ge_decompress(p, w);
  fe cx, cy;
  fe_canon(cx, p.X);
  fe_canon(cy, p.Y);
  memcpy(x_out, cx.v, 32);
  memcpy(y_out, cy.v, 32);
  return (int)ok;
}
int emu_enc_is_small_order(const uint8_t enc[32]) {
  uint32_t w[8];
  load_words(w, enc);
  return (int)ge_enc_is_small_order(w);
}
// generic-key verify of one (sig, pk, msg); returns HS_F_* flags
unsigned emu_verify_generic(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, uint64_t len) {
  ensure_btable();
  uint32_t R[8], S[8], A[8], h[16];
  load_words(R, sig);
  load_words(S, sig + 32);
  load_words(A, pk);
  uint8_t hb[64];
  emu_sha512_ram(sig, pk, msg, len, hb);
  memcpy(h, hb, 64);
  ge_cached tab[9];
  ge_ext acc;
  int32_t dig[HS_MAX_DIGITS];
  uint32_t meta = verify_generic_main(acc, R, S, A, h, g_btable.data(), tab, dig, 1, g_cp);
  return finish_one(acc, R, meta);
}
// committee path: builds -A's comb table on the fly (slow; tests only)
unsigned emu_verify_committee(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, uint64_t len) {
  ensure_btable();
  uint32_t R[8], S[8], A[8], h[16];
  load_words(R, sig);
  load_words(S, sig + 32);
  load_words(A, pk);
  ge_ext Apt, negA;
  uint32_t a_ok = ge_decompress(Apt, A);
  uint32_t a_small = ge_enc_is_small_order(A);
  ge_neg(negA, Apt);
  std::vector<ge_niels> at;
  if (!a_ok) ge_identity(negA);
  build_table(at, negA, g_cp.wa, g_cp.na);
  uint8_t hb[64];
  emu_sha512_ram(sig, pk, msg, len, hb);
  memcpy(h, hb, 64);
  ge_ext acc;
  int32_t dig[HS_MAX_DIGITS];
  uint32_t meta = verify_committee_main(acc, R, S, h, g_btable.data(), at.data(), (a_ok & 1u) | (a_small << 1), dig, 1, g_cp);
  return finish_one(acc, R, meta);
}
// latency path: 32 emulated lanes fetch one table entry each (lane j = digit j), convert it, and a 5-level tree adds them;
// R is decompressed separately and compared projectively — must give the same flags as the serial committee path.
unsigned emu_verify_committee_tree(const uint8_t sig[64], const uint8_t pk[32], const uint8_t *msg, uint64_t len) {
  ensure_btable();
  uint32_t R[8], S[8], A[8], h[16];
  load_words(R, sig);
  load_words(S, sig + 32);
  load_words(A, pk);
  ge_ext Apt, negA;
  uint32_t a_ok = ge_decompress(Apt, A);
  uint32_t a_small = ge_enc_is_small_order(A);
  ge_neg(negA, Apt);
  std::vector<ge_niels> at;
  if (!a_ok) ge_identity(negA);
  build_table(at, negA, g_cp.wa, g_cp.na);
  uint8_t hb[64];
  emu_sha512_ram(sig, pk, msg, len, hb);
  memcpy(h, hb, 64);
  uint32_t k[8];
  sc_reduce512(k, h);
  int32_t dig[HS_MAX_DIGITS];
  sc_digits_rt(dig, 1, k, g_cp.bias_a, g_cp.wa, g_cp.na);
  sc_digits_rt(dig + g_cp.na, 1, S, g_cp.bias_b, g_cp.wb, g_cp.nb);
  const int NT = g_cp.na + g_cp.nb;
  int lanes = 1;
  while (lanes < NT) lanes <<= 1;
  std::vector<ge_ext> lane(lanes);
  for (int j = 0; j < lanes; j++) {
    if (j < NT) {
      uint32_t neg;
      bool is_a;
      const ge_niels *e = comb_entry(at.data(), g_btable.data(), dig, 1, j, g_cp, neg, is_a);
      niels_signed q;
      niels_load_signed(q, e, neg, is_a);
      ge_from_signed_niels(lane[j], q.m0, q.m1);
    } else {
      ge_identity(lane[j]);
    }
  }
  for (int step = 1; step < lanes; step <<= 1)
    for (int j = 0; j + step < lanes; j += 2 * step) ge_add_ext(lane[j], lane[j], lane[j + step]);
  ge_ext Rpt;
  uint32_t r_ok = ge_decompress(Rpt, R);
  uint32_t parse_ok = sc_is_canonical(S) & a_ok & 1u;
  uint32_t small = ge_enc_is_small_order(R) | a_small;
  uint32_t eq = parse_ok & r_ok & ge_proj_equals_affine(lane[0].X, lane[0].Y, lane[0].Z, Rpt.X, Rpt.Y);
  unsigned fl = 0;
  if (parse_ok) fl |= HS_F_PARSE_OK;
  if (eq) fl |= HS_F_EQ;
  if (small) fl |= HS_F_SMALL;
  if (eq && !small) fl |= HS_F_STRICT;
  return fl;
}
const uint8_t *emu_btable_bytes(uint64_t *nbytes) {
  ensure_btable();
  *nbytes = g_btable.size() * sizeof(ge_niels);
  return reinterpret_cast<const uint8_t *>(g_btable.data());
}
}
// long-message digest path: schedules expanded per block into a strided K+W table, rounds run from the table
extern "C" void emu_sha512_kw_path(const uint8_t *msg, uint64_t len, uint8_t out[64]) {
  sha512_state s;
  sha512_init(s);
  const uint64_t nblk = sha512_nblocks(len);
  std::vector<uint64_t> kw(80 * 32);
  for (uint64_t g0 = 0; g0 < nblk; g0 += 32) {
    const int cnt = (int)((nblk - g0 < 32) ? (nblk - g0) : 32);
    for (int l = 0; l < cnt; l++) {
      uint64_t w[16];
      sha512_block_words(w, msg, len, g0 + l);
      sha512_expand_kw(kw.data() + l, 32, w);
    }
    for (int l = 0; l < cnt; l++) sha512_compress_kw_strided(s, kw.data() + l, 32);
  }
  uint32_t o[16];
  sha512_output_words(s, o);
  memcpy(out, o, 64);
}
// padding-only last block through the precomputed schedule (len % 128 == 0)
extern "C" void emu_sha512_padkw_path(const uint8_t *msg, uint64_t len, uint8_t out[64]) {
  sha512_state s;
  sha512_init(s);
  const uint64_t pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  sha512_absorb_blocks(s, pre, 0, msg, len, 0, len / 128);
  sha512_kw t;
  sha512_pad_schedule(t, len);
  sha512_compress_kw(s, t);
  uint32_t o[16];
  sha512_output_words(s, o);
  memcpy(out, o, 64);
}

// load-generation signer (sign_digest_core / keygen_core): must reproduce RFC 8032 signatures byte for byte
extern "C" void emu_sign_digest(const uint8_t seed[32], const uint8_t msg32[32], uint8_t pk_out[32], uint8_t sig_out[64]) {
  ensure_btable();
  uint32_t sd[8], M[8], A[8], R[8], S[8];
  load_words(sd, seed);
  load_words(M, msg32);
  int32_t dig[HS_MAX_DIGITS];
  keygen_core(A, sd, g_btable.data(), dig, 1, g_cp);
  sign_digest_core(R, S, sd, A, M, g_btable.data(), dig, 1, g_cp);
  memcpy(pk_out, A, 32);
  memcpy(sig_out, R, 32);
  memcpy(sig_out + 32, S, 32);
}
