// verify_core.cuh — the per-signature decision procedure (one thread = one signature).
//
// Implements, for the reference's crypto crate:
//   Signature::verify       (crypto/src/lib.rs:200-204)  -> flag HS_F_STRICT  (dalek verify_strict)
//   Signature::verify_batch (crypto/src/lib.rs:206-219)  -> AND over votes of flag HS_F_EQ (deterministic restatement of
//                                                            the per-signature equation)
// Decision procedure:
//   parse_ok = (S < l) and A decompresses                       [Signature::from_bytes / PublicKey::from_bytes]
//   k        = SHA-512(R || A || M) mod l
//   R'       = [S]B + [k](-A)
//   eq       = parse_ok and R decompresses and R' == R as points (no cofactor)
//   small    = R or A is one of the 8 torsion points
//   strict   = eq and not small
//
// Scalar multiplication layout (B200-first, not dalek's vartime NAF — see DESIGN.md):
//   [S]B      : radix-2^8 signed comb over a precomputed table of j * 256^i * B (32 windows x 128 affine-Niels
//               entries, 384 KB, L2-resident) -> 32 mixed additions, no doublings.
//   [k](-A)   : generic key  -> radix-2^4 signed fixed window, 8-entry per-thread table, 252 doublings + 64 additions;
//               committee key -> the same comb as for B, over that validator's table in HBM -> 32 mixed additions.
//   Every lane executes the same operation sequence (identity entry for digit 0), so warps never diverge.
#pragma once
#include <cstdint>
#include "fe.cuh"
#include "ge.cuh"
#include "sc.cuh"
#include "sha512.cuh"

#define HS_F_PARSE_OK 1u
#define HS_F_EQ 4u
#define HS_F_SMALL 8u
#define HS_F_STRICT 16u

#define HS_COMB_W 8
#define HS_COMB_WINDOWS 32
#define HS_COMB_ENTRIES 128
#define HS_COMB_TABLE_NIELS (HS_COMB_WINDOWS * HS_COMB_ENTRIES)  // ge_niels per point

// ---- digit streams over a recoded scalar (static register indexing: the scalar is shifted, not indexed)
template <int W>
struct digits_lsb {  // least-significant digit first
  uint32_t u[8];
  HS_HD void init(const uint32_t (&s)[8]) {
    sc_recoded<W> r;
    sc_recode<W>(r, s);
    for (int i = 0; i < 8; i++) u[i] = r.u[i];
  }
  HS_HD int next() {
    int d = (int)(u[0] & ((1u << W) - 1u)) - (1 << (W - 1));
    for (int i = 0; i < 7; i++) u[i] = (u[i] >> W) | (u[i + 1] << (32 - W));
    u[7] >>= W;
    return d;
  }
};
template <int W>
struct digits_msb {  // most-significant digit first; requires W * ndigits == 256
  uint32_t u[8];
  HS_HD void init(const uint32_t (&s)[8]) {
    static_assert(W * sc_ndigits<W>() == 256, "msb stream needs W | 256");
    sc_recoded<W> r;
    sc_recode<W>(r, s);
    for (int i = 0; i < 8; i++) u[i] = r.u[i];
  }
  HS_HD int next() {
    int d = (int)(u[7] >> (32 - W)) - (1 << (W - 1));
    for (int i = 7; i > 0; i--) u[i] = (u[i] << W) | (u[i - 1] >> (32 - W));
    u[0] <<= W;
    return d;
  }
};

// ---- table-entry loads
HS_HD void niels_load(ge_niels &q, const ge_niels *p) {
#if defined(__CUDA_ARCH__)
  // 96 B, 32 B-aligned: six 16-byte read-only loads (three full sectors)
  const uint4 *s = reinterpret_cast<const uint4 *>(p);
  uint4 a = __ldg(s + 0), b = __ldg(s + 1), c = __ldg(s + 2), d = __ldg(s + 3), e = __ldg(s + 4), f = __ldg(s + 5);
  q.ypx.v[0] = a.x; q.ypx.v[1] = a.y; q.ypx.v[2] = a.z; q.ypx.v[3] = a.w;
  q.ypx.v[4] = b.x; q.ypx.v[5] = b.y; q.ypx.v[6] = b.z; q.ypx.v[7] = b.w;
  q.ymx.v[0] = c.x; q.ymx.v[1] = c.y; q.ymx.v[2] = c.z; q.ymx.v[3] = c.w;
  q.ymx.v[4] = d.x; q.ymx.v[5] = d.y; q.ymx.v[6] = d.z; q.ymx.v[7] = d.w;
  q.xy2d.v[0] = e.x; q.xy2d.v[1] = e.y; q.xy2d.v[2] = e.z; q.xy2d.v[3] = e.w;
  q.xy2d.v[4] = f.x; q.xy2d.v[5] = f.y; q.xy2d.v[6] = f.z; q.xy2d.v[7] = f.w;
#else
  q = *p;
#endif
}

// acc += sum_i digit_i(s) * 256^i * P   using P's comb table (32 x 128 affine Niels entries)
HS_HD void ge_comb_accumulate(ge_ext &acc, const ge_niels *table, const uint32_t (&s)[8]) {
  digits_lsb<HS_COMB_W> ds;
  ds.init(s);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < HS_COMB_WINDOWS; i++) {
    int d = ds.next();
    uint32_t neg = (uint32_t)(d < 0);
    int mag = d < 0 ? -d : d;
    ge_niels q;
    if (mag == 0) ge_niels_identity(q);
    else niels_load(q, table + (size_t)i * HS_COMB_ENTRIES + (mag - 1));
    ge_niels_cneg(q, neg);
    ge_madd(acc, acc, q);
  }
}

// acc = [k]P for an arbitrary point P (already negated by the caller when -A is wanted): radix-16 signed fixed window.
// tab: 9 cached entries of thread-private scratch (tab[j] = j*P, tab[0] = identity).
HS_HD void ge_scalarmult_window4(ge_ext &acc, const ge_ext &P, const uint32_t (&k)[8], ge_cached *tab) {
  ge_cached_identity(tab[0]);
  ge_to_cached(tab[1], P);
  ge_ext m;
  m = P;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int j = 2; j <= 8; j++) {
    ge_add_cached(m, m, tab[1]);  // complete formula: also correct for m == P (doubling) and torsion points
    ge_to_cached(tab[j], m);
  }
  digits_msb<4> ds;
  ds.init(k);
  ge_identity(acc);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 64; i++) {
    if (i != 0) {
      ge_p1p1 c;
      ge_dbl_p1p1(c, acc); ge_p1p1_to_proj(acc, c);
      ge_dbl_p1p1(c, acc); ge_p1p1_to_proj(acc, c);
      ge_dbl_p1p1(c, acc); ge_p1p1_to_proj(acc, c);
      ge_dbl_p1p1(c, acc); ge_p1p1_to_ext(acc, c);
    }
    int d = ds.next();
    uint32_t neg = (uint32_t)(d < 0);
    int mag = d < 0 ? -d : d;
    ge_cached q = tab[mag];
    ge_cached_cneg(q, neg);
    ge_add_cached(acc, acc, q);
  }
}

// Final comparison + flag assembly shared by both key modes.
HS_HD uint32_t verify_finish(const ge_ext &acc, const uint32_t (&R)[8], uint32_t parse_ok, uint32_t small) {
  fe zinv;
  fe_invert(zinv, acc.Z);
  uint32_t eq = ge_matches_encoding(acc.X, acc.Y, zinv, R) & parse_ok;
  uint32_t fl = 0;
  if (parse_ok) fl |= HS_F_PARSE_OK;
  if (eq) fl |= HS_F_EQ;
  if (small) fl |= HS_F_SMALL;
  if (eq && !small) fl |= HS_F_STRICT;
  return fl;
}

// Generic key: A is given by its 32-byte encoding and decompressed here.
//   h: the 64-byte SHA-512(R||A||M) as 16 LE words.  tab: 9-entry thread-private scratch.
HS_HD uint32_t verify_generic_core(const uint32_t (&R)[8], const uint32_t (&S)[8], const uint32_t (&A)[8],
                                   const uint32_t (&h)[16], const ge_niels *btable, ge_cached *tab) {
  uint32_t k[8];
  sc_reduce512(k, h);
  uint32_t s_ok = sc_is_canonical(S);
  ge_ext Apt, negA, acc;
  uint32_t a_ok = ge_decompress(Apt, A);
  uint32_t small = ge_enc_is_small_order(R) | ge_enc_is_small_order(A);
  ge_neg(negA, Apt);
  ge_scalarmult_window4(acc, negA, k, tab);
  ge_comb_accumulate(acc, btable, S);
  return verify_finish(acc, R, s_ok & a_ok, small);
}

// Committee key: -A's comb table was built at registration (hs_committee_register); a_flags bit0 = A decompressed,
// bit1 = A is small order.
HS_HD uint32_t verify_committee_core(const uint32_t (&R)[8], const uint32_t (&S)[8], const uint32_t (&h)[16],
                                     const ge_niels *btable, const ge_niels *neg_a_table, uint32_t a_flags) {
  uint32_t k[8];
  sc_reduce512(k, h);
  uint32_t s_ok = sc_is_canonical(S);
  uint32_t small = ge_enc_is_small_order(R) | ((a_flags >> 1) & 1u);
  ge_ext acc;
  ge_identity(acc);
  ge_comb_accumulate(acc, neg_a_table, k);
  ge_comb_accumulate(acc, btable, S);
  return verify_finish(acc, R, s_ok & (a_flags & 1u), small);
}

// ---- table construction (runs on the GPU at context creation / committee registration; also under host emu)
// One call fills window `win` of the comb table of point P: entries j * 256^win * P for j = 1..128, affine Niels.
HS_HD void comb_build_window(ge_niels *table, const ge_ext &P, int win) {
  ge_ext base = P;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < 8 * win; i++) ge_dbl(base, base);
  ge_cached cb;
  ge_to_cached(cb, base);
  ge_ext m = base;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int j = 1; j <= HS_COMB_ENTRIES; j++) {
    fe zinv, x, y;
    fe_invert(zinv, m.Z);
    fe_mul(x, m.X, zinv);
    fe_mul(y, m.Y, zinv);
    ge_niels q;
    ge_affine_to_niels(q, x, y);
    // store canonical field elements so table bytes are representation-independent
    fe_canon(q.ypx, q.ypx);
    fe_canon(q.ymx, q.ymx);
    fe_canon(q.xy2d, q.xy2d);
    table[(size_t)win * HS_COMB_ENTRIES + (j - 1)] = q;
    if (j < HS_COMB_ENTRIES) ge_add_cached(m, m, cb);
  }
}
