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

// Comb window widths (bits).  B's table is shared by every signature and lives in L2 (16 x 32768 x 96 B = 48 MB);
// each committee key gets its own table in HBM (22 x 2048 x 96 B = 4.1 MB per key; 10,000 validators = 41 GB of the
// B200's 180 GB).  Wider windows trade HBM capacity / gather traffic for fewer field multiplications — the right trade
// on a part whose integer-multiply pipe, not its memory system, is the binding resource.
#ifndef HS_A_W
#define HS_A_W 12
#endif
#ifndef HS_B_W
#define HS_B_W 16
#endif
#define HS_A_WINDOWS (sc_ndigits<HS_A_W>())
#define HS_B_WINDOWS (sc_ndigits<HS_B_W>())
#define HS_A_ENTRIES (1 << (HS_A_W - 1))
#define HS_B_ENTRIES (1 << (HS_B_W - 1))
#define HS_A_TABLE_NIELS ((size_t)HS_A_WINDOWS * HS_A_ENTRIES)  // ge_niels per committee key
#define HS_B_TABLE_NIELS ((size_t)HS_B_WINDOWS * HS_B_ENTRIES)

// ---- digit streams over a recoded scalar (static register indexing: the scalar is shifted, not indexed)
template <int W>
struct digits_lsb {  // least-significant digit first
  uint32_t u[9];
  HS_HD void init(const uint32_t (&s)[8]) {
    sc_recoded<W> r;
    sc_recode<W>(r, s);
    for (int i = 0; i < 9; i++) u[i] = r.u[i];
  }
  HS_HD int next() {
    int d = (int)(u[0] & ((1u << W) - 1u)) - (1 << (W - 1));
    for (int i = 0; i < 8; i++) u[i] = (u[i] >> W) | (u[i + 1] << (32 - W));
    u[8] >>= W;
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

// acc += sum_i digit_i(s) * 2^(W i) * P   using P's comb table (windows x 2^(W-1) affine Niels entries)
template <int W>
HS_HD void ge_comb_accumulate(ge_ext &acc, const ge_niels *table, const uint32_t (&s)[8]) {
  constexpr int NW = sc_ndigits<W>();
  constexpr int NE = 1 << (W - 1);
  digits_lsb<W> ds;
  ds.init(s);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < NW; i++) {
    int d = ds.next();
    uint32_t neg = (uint32_t)(d < 0);
    int mag = d < 0 ? -d : d;
    ge_niels q;
    if (mag == 0) ge_niels_identity(q);
    else niels_load(q, table + (size_t)i * NE + (mag - 1));
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

// ---- phase 1 ("main"): everything up to the projective result R' = [S]B + [k](-A) = (X : Y : Z).
// meta bit0 = parse_ok (S canonical, A decompresses), bit1 = small (R or A is a torsion point).
#define HS_META_PARSE_OK 1u
#define HS_META_SMALL 2u
#define HS_META_MISS 0x80u  // committee lookup miss: record is re-run through the generic path

HS_HD uint32_t verify_generic_main(ge_ext &acc, const uint32_t (&R)[8], const uint32_t (&S)[8], const uint32_t (&A)[8],
                                   const uint32_t (&h)[16], const ge_niels *btable, ge_cached *tab) {
  uint32_t k[8];
  sc_reduce512(k, h);
  uint32_t s_ok = sc_is_canonical(S);
  ge_ext Apt, negA;
  uint32_t a_ok = ge_decompress(Apt, A);
  uint32_t small = ge_enc_is_small_order(R) | ge_enc_is_small_order(A);
  ge_neg(negA, Apt);
  ge_scalarmult_window4(acc, negA, k, tab);
  ge_comb_accumulate<HS_B_W>(acc, btable, S);
  return ((s_ok & a_ok) ? HS_META_PARSE_OK : 0u) | (small ? HS_META_SMALL : 0u);
}
// Committee key: -A's comb table was built at registration; a_flags bit0 = A decompressed, bit1 = A is small order.
HS_HD uint32_t verify_committee_main(ge_ext &acc, const uint32_t (&R)[8], const uint32_t (&S)[8], const uint32_t (&h)[16],
                                     const ge_niels *btable, const ge_niels *neg_a_table, uint32_t a_flags) {
  uint32_t k[8];
  sc_reduce512(k, h);
  uint32_t s_ok = sc_is_canonical(S);
  uint32_t small = ge_enc_is_small_order(R) | ((a_flags >> 1) & 1u);
  ge_identity(acc);
  ge_comb_accumulate<HS_A_W>(acc, neg_a_table, k);
  ge_comb_accumulate<HS_B_W>(acc, btable, S);
  return ((s_ok & a_flags & 1u) ? HS_META_PARSE_OK : 0u) | (small ? HS_META_SMALL : 0u);
}

// ---- phase 2 ("finish"): affine comparison with R's encoding given 1/Z (the inversion is batched by the caller)
HS_HD uint32_t verify_flags_from(const fe &X, const fe &Y, const fe &zinv, const uint32_t (&R)[8], uint32_t meta) {
  uint32_t parse_ok = meta & HS_META_PARSE_OK, small = (meta & HS_META_SMALL) ? 1u : 0u;
  uint32_t eq = (ge_matches_encoding(X, Y, zinv, R) && parse_ok) ? 1u : 0u;
  uint32_t fl = 0;
  if (parse_ok) fl |= HS_F_PARSE_OK;
  if (eq) fl |= HS_F_EQ;
  if (small) fl |= HS_F_SMALL;
  if (eq && !small) fl |= HS_F_STRICT;
  return fl;
}

// ---- table construction (runs on the GPU at context creation / committee registration; also under host emu)
// Fills entries [first, first+count) of window `win` of P's comb table: entry e (0-based) = (e+1) * 2^(W win) * P as an
// affine Niels point.  The forward pass parks (X, Y, Z) in the destination slots and the running product of the Z's in
// `prod` (count entries of scratch); one inversion then serves the whole block (Montgomery's trick).
HS_HD void comb_build_block(ge_niels *table, const ge_ext &P, int W, int win, int first, int count, fe *prod) {
  ge_ext base = P;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < W * win; i++) ge_dbl(base, base);
  ge_cached cb;
  ge_to_cached(cb, base);
  // m = (first + 1) * base by double-and-add (first + 1 <= 2^(W-1))
  ge_ext m;
  ge_identity(m);
  const int mult = first + 1;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int b = 16; b >= 0; b--) {
    ge_dbl(m, m);
    if ((mult >> b) & 1) ge_add_cached(m, m, cb);
  }
  ge_niels *slot = table + ((size_t)win << (W - 1)) + first;
  fe run;
  fe_set1(run);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int c = 0; c < count; c++) {
    slot[c].ypx = m.X;
    slot[c].ymx = m.Y;
    slot[c].xy2d = m.Z;
    fe_mul(run, run, m.Z);
    prod[c] = run;
    if (c + 1 < count) ge_add_cached(m, m, cb);
  }
  fe u;
  fe_invert(u, run);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int c = count - 1; c >= 0; c--) {
    fe X = slot[c].ypx, Y = slot[c].ymx, Z = slot[c].xy2d, zinv, x, y;
    if (c > 0) fe_mul(zinv, u, prod[c - 1]);
    else zinv = u;
    fe_mul(u, u, Z);
    fe_mul(x, X, zinv);
    fe_mul(y, Y, zinv);
    ge_niels q;
    ge_affine_to_niels(q, x, y);
    fe_canon(q.ypx, q.ypx);  // canonical field elements: table bytes are representation-independent
    fe_canon(q.ymx, q.ymx);
    fe_canon(q.xy2d, q.xy2d);
    slot[c] = q;
  }
}
