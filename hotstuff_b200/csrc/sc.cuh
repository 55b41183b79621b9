// sc.cuh — scalars mod l = 2^252 + 27742317777372353535851937790883648493 (the Ed25519 group order).
//
// Replaces dalek's Scalar::from_hash / from_bytes_mod_order_wide (k = SHA512(R||A||M) mod l) and the
// canonical-S check done by Signature::from_bytes, both reached from crypto/src/lib.rs:201-203,215-218.
// Portable C++ (32-bit limbs, 64-bit accumulators): runs once per signature, < 1 % of a verify.
#pragma once
#include <cstdint>
#include "fe.cuh"

// l, little-endian 32-bit limbs
#define HS_L0 0x5cf5d3edu
#define HS_L1 0x5812631au
#define HS_L2 0xa2f79cd6u
#define HS_L3 0x14def9deu
#define HS_L4 0u
#define HS_L5 0u
#define HS_L6 0u
#define HS_L7 0x10000000u

HS_HD uint32_t sc_l_limb(int i) {
  switch (i) {
    case 0: return HS_L0;
    case 1: return HS_L1;
    case 2: return HS_L2;
    case 3: return HS_L3;
    case 7: return HS_L7;
    default: return 0u;
  }
}
// mu = floor(2^512 / l), 9 limbs (Barrett constant)
HS_HD uint32_t sc_mu_limb(int i) {
  switch (i) {
    case 0: return 0x0a2c131bu;
    case 1: return 0xed9ce5a3u;
    case 2: return 0x086329a7u;
    case 3: return 0x2106215du;
    case 4: return 0xffffffebu;
    case 5: return 0xffffffffu;
    case 6: return 0xffffffffu;
    case 7: return 0xffffffffu;
    default: return 0xfu;
  }
}

// S < l ?  (dalek check_scalar: non-canonical S is a parse error)
HS_HD uint32_t sc_is_canonical(const uint32_t (&s)[8]) {
  // compute s - l, canonical iff it borrows
  int64_t acc = 0;
  for (int i = 0; i < 8; i++) {
    acc += (int64_t)s[i] - (int64_t)sc_l_limb(i);
    acc >>= 32;
  }
  return (uint32_t)(acc & 1);  // acc == -1 on borrow
}

// x (16 limbs, the 64-byte hash as a little-endian integer) mod l -> 8 limbs.  Barrett, b = 2^32, k = 8 (HAC 14.42).
HS_HD void sc_reduce512(uint32_t (&r)[8], const uint32_t (&x)[16]) {
  // q1 = floor(x / b^7): limbs x[7..15] (9 limbs); q2 = q1 * mu (18 limbs); q3 = floor(q2 / b^9) (limbs 9..17)
  uint32_t q2[18];
  for (int i = 0; i < 18; i++) q2[i] = 0;
  for (int i = 0; i < 9; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < 9; j++) {
      uint64_t t = (uint64_t)x[7 + i] * sc_mu_limb(j) + q2[i + j] + carry;
      q2[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    q2[i + 9] = (uint32_t)carry;
  }
  // r2 = q3 * l mod b^9
  uint32_t r2[9];
  for (int i = 0; i < 9; i++) r2[i] = 0;
  for (int i = 0; i < 9; i++) {
    uint64_t carry = 0;
    for (int j = 0; j + i < 9; j++) {
      uint32_t lj = (j < 8) ? sc_l_limb(j) : 0u;
      uint64_t t = (uint64_t)q2[9 + i] * lj + r2[i + j] + carry;
      r2[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
  }
  // r = (x mod b^9) - r2 mod b^9  (true remainder < 3l, so the wrap-around is exact)
  uint32_t t9[9];
  int64_t acc = 0;
  for (int i = 0; i < 9; i++) {
    acc += (int64_t)x[i] - (int64_t)r2[i];
    t9[i] = (uint32_t)acc;
    acc >>= 32;
  }
  // at most two conditional subtractions of l
  for (int pass = 0; pass < 2; pass++) {
    uint32_t u[9];
    int64_t a2 = 0;
    for (int i = 0; i < 9; i++) {
      uint32_t li = (i < 8) ? sc_l_limb(i) : 0u;
      a2 += (int64_t)t9[i] - (int64_t)li;
      u[i] = (uint32_t)a2;
      a2 >>= 32;
    }
    uint32_t keep = (uint32_t)(a2 & 1);  // borrow -> t9 < l -> keep t9
    uint32_t m = 0u - keep;
    for (int i = 0; i < 9; i++) t9[i] = (t9[i] & m) | (u[i] & ~m);
  }
  for (int i = 0; i < 8; i++) r[i] = t9[i];
}

// Signed fixed-window recoding.  For a scalar s < 2^253 and window width w (w | 32: 4 or 8; general w also fine),
// u = s + sum_i 2^(w-1) * 2^(w i) is computed once; digit_i = bits(u, w*i, w) - 2^(w-1) is in [-2^(w-1), 2^(w-1)-1]
// and sum digit_i 2^(w i) = s.  (No data-dependent branches: every lane does the same adds.)
template <int W>
HS_HD constexpr int sc_ndigits() {
  // the top digit must have room for the +2^(W-1) bias and a carry
  return (253 + W - 1) / W + ((253 % W == 0 || 253 % W == W - 1) ? 1 : 0);
}
template <int W>
HS_HD constexpr uint32_t sc_bias_limb(int i) {
  uint32_t c = 0;
  for (int d = 0; d < sc_ndigits<W>(); d++) {
    int bit = W - 1 + W * d;
    if ((bit >> 5) == i) c |= 1u << (bit & 31);
  }
  return c;
}

template <int W>
struct sc_recoded {
  uint32_t u[9];
};

template <int W>
HS_HD void sc_recode(sc_recoded<W> &out, const uint32_t (&s)[8]) {
  static_assert(W * sc_ndigits<W>() <= 288, "window too wide");
  constexpr uint32_t C[9] = {sc_bias_limb<W>(0), sc_bias_limb<W>(1), sc_bias_limb<W>(2), sc_bias_limb<W>(3), sc_bias_limb<W>(4),
                             sc_bias_limb<W>(5), sc_bias_limb<W>(6), sc_bias_limb<W>(7), sc_bias_limb<W>(8)};
  uint64_t acc = 0;
  for (int i = 0; i < 9; i++) {
    acc += (uint64_t)((i < 8) ? s[i] : 0u) + C[i];
    out.u[i] = (uint32_t)acc;
    acc >>= 32;
  }
}

// ---- runtime-width variant (the comb tables' window widths are chosen per context / per committee size)
HS_HD int sc_ndigits_rt(int W) {
  int r = 253 % W;
  return (253 + W - 1) / W + ((r == 0 || r == W - 1) ? 1 : 0);
}
// bias = sum_{i < ndigits} 2^(W-1 + W i); computed on the host once per table and passed to the kernels
HS_HD void sc_bias_rt(uint32_t (&b)[9], int W) {
  for (int i = 0; i < 9; i++) b[i] = 0;
  const int n = sc_ndigits_rt(W);
  for (int d = 0; d < n; d++) {
    int bit = W - 1 + W * d;
    b[bit >> 5] |= 1u << (bit & 31);
  }
}
// dig[i * stride] = signed digit i of s in radix 2^W (i < n), each in [-2^(W-1), 2^(W-1) - 1]; W <= 26 (hostemu checks every width 8 .. 26)
HS_HD void sc_digits_rt(int32_t *dig, int stride, const uint32_t (&s)[8], const uint32_t (&bias)[9], int W, int n) {
  uint32_t u[9];
  uint64_t acc = 0;
  for (int i = 0; i < 9; i++) {
    acc += (uint64_t)((i < 8) ? s[i] : 0u) + bias[i];
    u[i] = (uint32_t)acc;
    acc >>= 32;
  }
  const uint32_t mask = (1u << W) - 1u;
  const int half = 1 << (W - 1);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int i = 0; i < n; i++) {
    dig[i * stride] = (int32_t)(u[0] & mask) - half;
    for (int j = 0; j < 8; j++) u[j] = (u[j] >> W) | (u[j + 1] << (32 - W));
    u[8] >>= W;
  }
}
