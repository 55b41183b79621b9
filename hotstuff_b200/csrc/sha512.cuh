// sha512.cuh — SHA-512 (FIPS 180-4) for the verify kernels and the Digest surface.
//
// Replaces: ed25519_dalek::Sha512 as used for k = H(R || A || M) inside Signature::verify / verify_batch
// (crypto/src/lib.rs:200-219) and for Digest = SHA-512(..)[..32] at consensus/src/messages.rs:81,151,203,270,308
// and mempool/src/processor.rs:30.
//
// One thread hashes one message (Merkle–Damgård is sequential within a message; parallelism is across
// signatures / messages).  The 80 rounds are fully unrolled over a rolling 16-word schedule so every index
// is a compile-time constant and the schedule lives in registers.
#pragma once
#include <cstdint>
#include "fe.cuh"
#include "hs_constants.cuh"

#if defined(__CUDACC__)
__device__ __constant__ uint64_t HS_SHA512_K_DEV[80] = {HS_SHA512_K_INIT};
__device__ __constant__ uint32_t HS_ONE_DEV = 1;  // opaque multiplier: keeps ptxas from folding mad.wide(x, 1, y) back into ALU adds
#endif

HS_HD uint64_t sha_k(int i) {
#if defined(__CUDA_ARCH__)
  return HS_SHA512_K_DEV[i];
#else
  return HS_SHA512_K_HOST[i];
#endif
}

HS_HD uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
// SHA-512 on a 32-bit datapath is ALU-pipe bound (rotates, xors and add/add-with-carry all issue there: ncu shows the
// digest kernel at 86 % ALU-pipe utilisation).  Experiment HS_SHA_FMA_ADD moves half of every 64-bit add to the FMA pipe
// (x + y = mad.wide(x_lo, 1, y) + (x_hi << 32)); measured on B200 it is SLOWER (1.95 ms vs 1.33 ms per 2^20 x 512 B:
// IMAD.WIDE costs ~2.4 issue cycles and drags register moves along), so the plain adds stay the default.
HS_HD uint64_t add64_fma(uint64_t x, uint64_t y) {
#if defined(__CUDA_ARCH__) && defined(HS_SHA_FMA_ADD)
  uint64_t t;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"((uint32_t)x), "r"(HS_ONE_DEV), "l"(y));
  return t + (x & 0xffffffff00000000ULL);
#else
  return x + y;
#endif
}
HS_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(x, 0, 0x0123);
#else
  return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}
// big-endian 64-bit message word from two little-endian 32-bit memory words (lo = bytes 0..3, hi = bytes 4..7)
HS_HD uint64_t be64_from_le32(uint32_t lo, uint32_t hi) { return ((uint64_t)bswap32(lo) << 32) | bswap32(hi); }

struct sha512_state {
  uint64_t h[8];
};

HS_HD void sha512_init(sha512_state &s) {
  const uint64_t h0[8] = {HS_SHA512_H0_INIT};
  for (int i = 0; i < 8; i++) s.h[i] = h0[i];
}

// One compression; w[16] is consumed (used as the rolling schedule).
// Code shape: 16 fully unrolled rounds over the loaded block, then a 4-trip loop whose body is 16 unrolled rounds with
// the schedule update — every w[] / state index is a compile-time constant (registers), but the hot body is ~2.5 k
// instructions instead of the ~8.5 k of an 80-round unroll, which thrashed the instruction cache (ncu: 6 of every 10
// stall cycles of the digest kernel were "no instruction").
#define HS_SHA_ROUND(a, b, c, d, e, f, g, h, kw)                                 \
  {                                                                              \
    uint64_t t1_ = add64_fma(add64_fma(h, rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41)), add64_fma((e & f) ^ (~e & g), (kw))); \
    uint64_t t2_ = add64_fma(rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39), (a & b) ^ (a & c) ^ (b & c));   \
    d = add64_fma(d, t1_);                                                       \
    h = add64_fma(t1_, t2_);                                                     \
  }
#define HS_SHA_SCHED(w, j)                                                                          \
  {                                                                                                 \
    uint64_t w15_ = w[((j) + 1) & 15], w2_ = w[((j) + 14) & 15];                                    \
    w[(j) & 15] = add64_fma(add64_fma(w[(j) & 15], rotr64(w15_, 1) ^ rotr64(w15_, 8) ^ (w15_ >> 7)),  \
                            add64_fma(w[((j) + 9) & 15], rotr64(w2_, 19) ^ rotr64(w2_, 61) ^ (w2_ >> 6))); \
  }
#define HS_SHA_8ROUNDS(w, base, j0)                                   \
  HS_SHA_ROUND(a, b, c, d, e, f, g, h, sha_k((base) + (j0) + 0) + w[(j0) + 0]) \
  HS_SHA_ROUND(h, a, b, c, d, e, f, g, sha_k((base) + (j0) + 1) + w[(j0) + 1]) \
  HS_SHA_ROUND(g, h, a, b, c, d, e, f, sha_k((base) + (j0) + 2) + w[(j0) + 2]) \
  HS_SHA_ROUND(f, g, h, a, b, c, d, e, sha_k((base) + (j0) + 3) + w[(j0) + 3]) \
  HS_SHA_ROUND(e, f, g, h, a, b, c, d, sha_k((base) + (j0) + 4) + w[(j0) + 4]) \
  HS_SHA_ROUND(d, e, f, g, h, a, b, c, sha_k((base) + (j0) + 5) + w[(j0) + 5]) \
  HS_SHA_ROUND(c, d, e, f, g, h, a, b, sha_k((base) + (j0) + 6) + w[(j0) + 6]) \
  HS_SHA_ROUND(b, c, d, e, f, g, h, a, sha_k((base) + (j0) + 7) + w[(j0) + 7])

HS_HD void sha512_compress(sha512_state &s, uint64_t (&w)[16]) {
  uint64_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
  HS_SHA_8ROUNDS(w, 0, 0)
  HS_SHA_8ROUNDS(w, 0, 8)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int base = 16; base < 80; base += 16) {
    HS_SHA_SCHED(w, 0) HS_SHA_SCHED(w, 1) HS_SHA_SCHED(w, 2) HS_SHA_SCHED(w, 3)
    HS_SHA_SCHED(w, 4) HS_SHA_SCHED(w, 5) HS_SHA_SCHED(w, 6) HS_SHA_SCHED(w, 7)
    HS_SHA_8ROUNDS(w, base, 0)
    HS_SHA_SCHED(w, 8) HS_SHA_SCHED(w, 9) HS_SHA_SCHED(w, 10) HS_SHA_SCHED(w, 11)
    HS_SHA_SCHED(w, 12) HS_SHA_SCHED(w, 13) HS_SHA_SCHED(w, 14) HS_SHA_SCHED(w, 15)
    HS_SHA_8ROUNDS(w, base, 8)
  }
  s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

// state -> 64 output bytes as 16 little-endian u32 words (word j = bytes 4j..4j+3 of the digest)
HS_HD void sha512_output_words(const sha512_state &s, uint32_t (&out)[16]) {
  for (int i = 0; i < 8; i++) {
    out[2 * i] = bswap32((uint32_t)(s.h[i] >> 32));
    out[2 * i + 1] = bswap32((uint32_t)s.h[i]);
  }
}

// k-hash fast path: SHA-512(R[32] || A[32] || M[32]) — exactly one block (every message the reference signs is a
// 32-byte Digest: crypto/src/lib.rs:185,200,206).  Inputs are little-endian u32 words as loaded from memory.
HS_HD void sha512_ram32(uint32_t (&out)[16], const uint32_t (&R)[8], const uint32_t (&A)[8], const uint32_t (&M)[8]) {
  uint64_t w[16];
  for (int i = 0; i < 4; i++) {
    w[i] = be64_from_le32(R[2 * i], R[2 * i + 1]);
    w[4 + i] = be64_from_le32(A[2 * i], A[2 * i + 1]);
    w[8 + i] = be64_from_le32(M[2 * i], M[2 * i + 1]);
  }
  w[12] = 0x8000000000000000ULL;
  w[13] = 0;
  w[14] = 0;
  w[15] = 96 * 8;
  sha512_state s;
  sha512_init(s);
  sha512_compress(s, w);
  sha512_output_words(s, out);
}

// Byte-granular reader for arbitrary-length / arbitrarily-aligned messages.
HS_HD uint64_t load_be64_bytes(const uint8_t *p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
  return v;
}

// One 64-bit big-endian message word at message offset m (multiple of 8): real bytes, then the 0x80 terminator,
// then zeros.  Keeps the schedule index static so w[] stays in registers.
HS_HD uint64_t sha512_msg_word(const uint8_t *msg, uint64_t len, uint64_t m, bool aligned8) {
  if (m + 8 <= len) {
    if (aligned8) {
      uint64_t le = *reinterpret_cast<const uint64_t *>(msg + m);
      return ((uint64_t)bswap32((uint32_t)le) << 32) | bswap32((uint32_t)(le >> 32));
    }
    return load_be64_bytes(msg + m);
  }
  if (m > len) return 0;
  uint64_t v = 0;
  int nb = (int)(len - m);
  for (int i = 0; i < nb; i++) v |= (uint64_t)msg[m + i] << (56 - 8 * i);
  return v | ((uint64_t)0x80 << (56 - 8 * nb));
}

// General hash of prefix || msg[0..len): the prefix is n_prefix_words (0 or 8) big-endian 64-bit words already in
// registers (R||A for the k-hash; none for Digest); message bytes stream from global memory.
HS_HD void sha512_prefix_msg(uint32_t (&out)[16], const uint64_t (&prefix_words)[8], int n_prefix_words, const uint8_t *msg,
                             uint64_t len) {
  sha512_state s;
  sha512_init(s);
  const uint64_t P = (uint64_t)n_prefix_words * 8;
  const uint64_t total = P + len;
  const uint64_t nblk = (total + 17 + 127) / 128;
  const bool aligned8 = ((reinterpret_cast<uintptr_t>(msg) & 7u) == 0);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (uint64_t b = 0; b < nblk; b++) {
    uint64_t w[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 16; j++) {
      uint64_t o = b * 128 + 8 * (uint64_t)j;
      uint64_t v;
      if (o < P) v = prefix_words[j & 7];  // only reachable for b == 0, j < 8
      else v = sha512_msg_word(msg, len, o - P, aligned8);
      if (b == nblk - 1 && j == 14) v = total >> 61;
      if (b == nblk - 1 && j == 15) v = total << 3;
      w[j] = v;
    }
    sha512_compress(s, w);
  }
  sha512_output_words(s, out);
}
