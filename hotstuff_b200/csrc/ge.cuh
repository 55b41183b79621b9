// ge.cuh — twisted Edwards curve -x^2 + y^2 = 1 + d x^2 y^2 over GF(2^255-19): the point arithmetic behind
// Signature::verify / verify_batch (crypto/src/lib.rs:200-219, i.e. dalek PublicKey::from_bytes, verify_strict and the
// per-signature equation of verify_batch).  Accept/reject semantics follow SURVEY.md Appendix A.
#pragma once
#include <cstdint>
#include "fe.cuh"
#include "hs_constants.cuh"

struct ge_ext {  // extended coordinates: x = X/Z, y = Y/Z, T = XY/Z
  fe X, Y, Z, T;
};
struct ge_niels {  // affine precomputed point: (y+x, y-x, 2dxy) — 96 bytes, the table-entry format
  fe ypx, ymx, xy2d;
};
struct ge_cached {  // projective precomputed point: (Y+X, Y-X, Z, 2dT)
  fe YpX, YmX, Z, T2d;
};

HS_HD void fe_const(fe &r, const uint32_t *c) {
  for (int i = 0; i < 8; i++) r.v[i] = c[i];
}
#if defined(__CUDACC__)
__device__ __constant__ uint32_t HS_D_DEV[8] = {HS_D_INIT};
__device__ __constant__ uint32_t HS_D2_DEV[8] = {HS_D2_INIT};
__device__ __constant__ uint32_t HS_SQRTM1_DEV[8] = {HS_SQRTM1_INIT};
__device__ __constant__ uint32_t HS_BX_DEV[8] = {HS_BX_INIT};
__device__ __constant__ uint32_t HS_BY_DEV[8] = {HS_BY_INIT};
#endif
#if defined(__CUDA_ARCH__)
#define HS_CONST(name) name##_DEV
#else
#define HS_CONST(name) name##_32
#endif

HS_HD void ge_identity(ge_ext &p) {
  fe_set0(p.X);
  fe_set1(p.Y);
  fe_set1(p.Z);
  fe_set0(p.T);
}
HS_HD void ge_basepoint(ge_ext &p) {
  fe_const(p.X, HS_CONST(HS_BX));
  fe_const(p.Y, HS_CONST(HS_BY));
  fe_set1(p.Z);
  fe_mul(p.T, p.X, p.Y);
}
HS_HD void ge_neg(ge_ext &r, const ge_ext &p) {
  fe_neg(r.X, p.X);
  fe_copy(r.Y, p.Y);
  fe_copy(r.Z, p.Z);
  fe_neg(r.T, p.T);
}

// ---- "completed" intermediate (E,F,G,H): X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
struct ge_p1p1 {
  fe E, F, G, H;
};
HS_HD void ge_p1p1_to_ext(ge_ext &r, const ge_p1p1 &c) {
  fe_mul(r.X, c.E, c.F);
  fe_mul(r.Y, c.G, c.H);
  fe_mul(r.Z, c.F, c.G);
  fe_mul(r.T, c.E, c.H);
}
// X, Y, Z only (T left untouched) — enough when the next operation is a doubling
HS_HD void ge_p1p1_to_proj(ge_ext &r, const ge_p1p1 &c) {
  fe_mul(r.X, c.E, c.F);
  fe_mul(r.Y, c.G, c.H);
  fe_mul(r.Z, c.F, c.G);
}

// doubling, a = -1 (reads X, Y, Z only): A = X^2, B = Y^2, C = 2 Z^2, H = A + B, E = H - (X+Y)^2, G = A - B, F = C + G
HS_HD void ge_dbl_p1p1(ge_p1p1 &c, const ge_ext &p) {
  fe a, b, t;
  fe_sqr(a, p.X);
  fe_sqr(b, p.Y);
  fe_add(t, p.X, p.Y);
  fe_sqr(t, t);
  fe_add(c.H, a, b);
  fe_sub(c.E, c.H, t);
  fe_sub(c.G, a, b);
  fe_sqr(t, p.Z);
  fe_add(t, t, t);
  fe_add(c.F, t, c.G);
}
HS_HD void ge_dbl(ge_ext &r, const ge_ext &p) {
  ge_p1p1 c;
  ge_dbl_p1p1(c, p);
  ge_p1p1_to_ext(r, c);
}

// Mixed addition with an affine Niels point (7M); complete for every input on the curve.
// r = p + (neg ? -q : q) without negating a field element: -q = (ymx, ypx, -xy2d), and -xy2d only swaps F and G.
HS_HD void ge_madd_signed_p1p1(ge_p1p1 &c, const ge_ext &p, const ge_niels &q, uint32_t neg) {
  fe a, b, t, dd, m0, m1;
  fe_select(m0, q.ymx, q.ypx, neg);
  fe_select(m1, q.ypx, q.ymx, neg);
  fe_sub(t, p.Y, p.X);
  fe_mul(a, t, m0);
  fe_add(t, p.Y, p.X);
  fe_mul(b, t, m1);
  fe_mul(t, p.T, q.xy2d);
  fe_add(dd, p.Z, p.Z);
  fe_sub(c.E, b, a);
  fe_add(c.H, b, a);
  fe_sub(c.F, dd, t);
  fe_add(c.G, dd, t);
  fe_cswap(c.F, c.G, neg);
}
HS_HD void ge_madd_signed(ge_ext &r, const ge_ext &p, const ge_niels &q, uint32_t neg) {
  ge_p1p1 c;
  ge_madd_signed_p1p1(c, p, q, neg);
  ge_p1p1_to_ext(r, c);
}
// addition with a projective cached point (8M)
HS_HD void ge_add_cached_p1p1(ge_p1p1 &c, const ge_ext &p, const ge_cached &q) {
  fe a, b, t, dd;
  fe_sub(t, p.Y, p.X);
  fe_mul(a, t, q.YmX);
  fe_add(t, p.Y, p.X);
  fe_mul(b, t, q.YpX);
  fe_mul(t, p.T, q.T2d);
  fe_mul(dd, p.Z, q.Z);
  fe_add(dd, dd, dd);
  fe_sub(c.E, b, a);
  fe_add(c.H, b, a);
  fe_sub(c.F, dd, t);
  fe_add(c.G, dd, t);
}
HS_HD void ge_add_cached(ge_ext &r, const ge_ext &p, const ge_cached &q) {
  ge_p1p1 c;
  ge_add_cached_p1p1(c, p, q);
  ge_p1p1_to_ext(r, c);
}
HS_HD void ge_to_cached(ge_cached &r, const ge_ext &p) {
  fe d2;
  fe_const(d2, HS_CONST(HS_D2));
  fe_add(r.YpX, p.Y, p.X);
  fe_sub(r.YmX, p.Y, p.X);
  fe_copy(r.Z, p.Z);
  fe_mul(r.T2d, p.T, d2);
}
HS_HD void ge_cached_identity(ge_cached &r) {
  fe_set1(r.YpX);
  fe_set1(r.YmX);
  fe_set1(r.Z);
  fe_set0(r.T2d);
}
// r = neg ? -q : q   (negating a cached point swaps Y+X / Y-X and negates 2dT)
HS_HD void ge_cached_cneg(ge_cached &q, uint32_t neg) {
  fe_cswap(q.YpX, q.YmX, neg);
  fe nt;
  fe_neg(nt, q.T2d);
  fe_select(q.T2d, q.T2d, nt, neg);
}
HS_HD void ge_niels_identity(ge_niels &r) {
  fe_set1(r.ypx);
  fe_set1(r.ymx);
  fe_set0(r.xy2d);
}
// affine (x, y) -> Niels
HS_HD void ge_affine_to_niels(ge_niels &r, const fe &x, const fe &y) {
  fe d2, t;
  fe_const(d2, HS_CONST(HS_D2));
  fe_add(r.ypx, y, x);
  fe_sub(r.ymx, y, x);
  fe_mul(t, x, y);
  fe_mul(r.xy2d, t, d2);
}

// ---- decompression: dalek CompressedEdwardsY::decompress (tolerant rules):
//   y = bytes with bit 255 cleared, NOT checked < p; u = y^2-1, v = d y^2+1; x = sqrt_ratio_i(u, v) (the even root);
//   fail iff u/v is not a square; negate x when the sign bit is set ("x = 0, sign = 1" is accepted).
// Returns 1 on success.  Output is affine (Z = 1).
HS_HD uint32_t ge_decompress(ge_ext &p, const uint32_t (&enc)[8]) {
  fe y, yy, u, v, v3, v7, r, t, chk, dconst, sqm1, one;
  fe_const(dconst, HS_CONST(HS_D));
  fe_const(sqm1, HS_CONST(HS_SQRTM1));
  fe_set1(one);
  fe_from_words(y, enc);
  fe_sqr(yy, y);
  fe_sub(u, yy, one);
  fe_mul(v, yy, dconst);
  fe_add(v, v, one);
  fe_sqr(t, v);
  fe_mul(v3, t, v);
  fe_sqr(t, v3);
  fe_mul(v7, t, v);
  fe_mul(t, u, v7);
  fe_pow_p58(t, t);
  fe_mul(t, t, v3);
  fe_mul(r, t, u);  // r = u v^3 (u v^7)^((p-5)/8)
  fe_sqr(t, r);
  fe_mul(chk, t, v);  // v r^2
  fe neg_u, neg_u_i, ri;
  fe_neg(neg_u, u);
  fe_mul(neg_u_i, neg_u, sqm1);
  uint32_t correct = fe_eq(chk, u);
  uint32_t flipped = fe_eq(chk, neg_u);
  uint32_t flipped_i = fe_eq(chk, neg_u_i);
  fe_mul(ri, r, sqm1);
  fe_select(r, r, ri, flipped | flipped_i);
  // choose the non-negative root, then apply the encoded sign
  uint32_t neg = fe_is_neg(r) ^ (enc[7] >> 31);
  fe nr;
  fe_neg(nr, r);
  fe_select(p.X, r, nr, neg);
  fe_copy(p.Y, y);
  fe_set1(p.Z);
  fe_mul(p.T, p.X, p.Y);
  return correct | flipped;
}

// Small-order test on the *encoding*: the eight torsion points have y in {0, 1, -1, y8, -y8}; dalek reduces a
// non-canonical y (y + p, only possible for y < 19) before use, and ignores the sign bit when x = 0, so
// [8]P == identity  <=>  (enc mod 2^255) in {0, 1, p-1, p, p+1, y8, p-y8}  (every one of these decompresses, with
// either sign bit).  Equivalent to dalek is_small_order() = mul_by_cofactor().is_identity() for decompressible input;
// tests/test_oracle_pins.py::test_small_order_equivalence and tests/test_hostemu.py::test_decompress_and_small_order check the
// equivalence against the oracle's [8]P computation.
HS_HD uint32_t ge_enc_is_small_order(const uint32_t (&enc)[8]) {
  const uint32_t top = enc[7] & 0x7fffffffu;
  // y = 0 / 1 (canonical): limbs 1..7 zero, limb0 in {0,1}
  uint32_t mid_zero = (enc[1] | enc[2] | enc[3] | enc[4] | enc[5] | enc[6]) == 0;
  uint32_t small01 = mid_zero & (top == 0) & (enc[0] <= 1u);
  // y in {p-1, p, p+1} = 2^255 - {20, 19, 18}: limbs 1..6 all ones, top = 0x7fffffff, limb0 in {0xffffffec, ed, ee}
  uint32_t mid_ones = (enc[1] & enc[2] & enc[3] & enc[4] & enc[5] & enc[6]) == 0xffffffffu;
  uint32_t nearp = mid_ones & (top == 0x7fffffffu) & (enc[0] >= 0xffffffecu) & (enc[0] <= 0xffffffeeu);
  // y8 = 0x05fc536d880238b13933c6d305acdfd5f098eff289f4c345b027b2c28f95e826 and p - y8
  const uint32_t y8a[8] = {0x8f95e826u, 0xb027b2c2u, 0x89f4c345u, 0xf098eff2u, 0x05acdfd5u, 0x3933c6d3u, 0x880238b1u, 0x05fc536du};
  const uint32_t y8b[8] = {0x706a17c7u, 0x4fd84d3du, 0x760b3cbau, 0x0f67100du, 0xfa53202au, 0xc6cc392cu, 0x77fdc74eu, 0x7a03ac92u};
  uint32_t da = 0, db = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t w = (i == 7) ? top : enc[i];
    da |= w ^ y8a[i];
    db |= w ^ y8b[i];
  }
  return small01 | nearp | (da == 0) | (db == 0);
}

// Does the projective point (X:Y:Z) equal the point the 32-byte encoding `enc` decompresses to?
// Given x' = X/Z, y' = Y/Z (zinv = 1/Z): equal iff y' == y_enc (mod p) and (x' == 0 or sign(x') == sign bit).
// If `enc` does not decompress, no curve point has that y, so the y comparison already fails — this is exactly
// "R decompresses and [S]B-[k]A == R as points" (dalek verify_strict step 5 / the verify_batch equation) without a
// square root for R.  tests/hostemu checks it against explicit decompress + projective compare.
HS_HD uint32_t ge_matches_encoding(const fe &X, const fe &Y, const fe &zinv, const uint32_t (&enc)[8]) {
  fe x, y, ye;
  fe_mul(x, X, zinv);
  fe_mul(y, Y, zinv);
  fe_from_words(ye, enc);
  uint32_t y_ok = fe_eq(y, ye);
  fe xc;
  fe_canon(xc, x);
  uint32_t xz = 0;
  for (int i = 0; i < 8; i++) xz |= xc.v[i];
  uint32_t sign_ok = ((xc.v[0] & 1u) == (enc[7] >> 31)) | (xz == 0);
  return y_ok & sign_ok;
}
