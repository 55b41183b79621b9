// hs_engine.cu — sm_100a kernels + the C ABI of include/hs_crypto.h.
//
// Hot path of asonnino/hotstuff's crypto crate (crypto/src/lib.rs:200-219 + the SHA-512 Digest call sites) rebuilt for
// B200.  No CPU path: if CUDA fails the call returns an error and the caller must reject.
//
// Throughput pipeline of one verify call (any n):
//   k_key_lookup      pk bytes -> committee index through a device hash table; misses -> compacted list
//                     (skipped when the caller gives validator indices)
//   k_verify_main<C>  registered / learned keys: SHA-512(R||A||M) -> k mod l -> [k](-A) + [S]B by table gathers only
//                     (17 + 11 mixed additions for 4,096 keys; no doublings, no decompression) -> (X:Y:Z) + meta
//   k_verify_main<G>  records whose key has no table: decompress A, radix-16 window for [k](-A); runs over the
//                     compacted list on a high-priority side stream, beside the pass above
//   k_verify_finish   two-level Montgomery-batched inversion, affine compare with R's encoding, small-order rule,
//                     32 verdicts -> one bitmap word, stored locally or into every peer GPU's buffer (fused all-gather,
//                     epoch flags exchanged by the last block); optionally on the context's tail stream (deferred mode)
// Latency pipeline (n <= 64, every key has a table): k_verify_small — ONE launch, a warp per signature sums the table entries
//   with a shuffle tree while a second warp decompresses R; inputs / verdicts in mapped pinned memory.
// Digest: k_digest32_fixed (staged coalesced loads, constant padding schedule), k_digest32 (any length), k_digest32_long
//   (one warp per long message, schedules expanded across lanes).
// Front ends: QC / TC / Timeout / Block groups with on-GPU digests and per-certificate AND; load-generation keygen / signer.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/hs_crypto.h"
#include "verify_core.cuh"

#define HS_THREADS 128
#define HS_FINISH_GROUP 16  // signatures whose Z's share one inversion
#define HS_NO_KEY 0xffffffffu
#define HS_LEARN_MAX 8192u  // unknown keys examined per call
#define HS_LEARN_PER_CALL 1024u  // new tables built per call at most (bounds the latency a flood of one-off keys can add to one call)
#define HS_CACHE_RESET_MIN_CALLS 64u  // a full cache is cleared at most once per this many verify calls

// ------------------------------------------------------------------------------------------------ input layout
// One descriptor covers every caller-facing layout: packed hs_rec128 records, separate sig/pk arrays with
// variable-length messages, QC votes sharing one digest, committee-indexed votes.
struct in_layout {
  const uint8_t *sig;   // 64-byte signature of record i at sig + i * sig_stride
  size_t sig_stride;
  const uint8_t *pk;    // 32-byte key of record i at pk + i * pk_stride (nullptr when only indices are given)
  size_t pk_stride;
  const uint32_t *vidx; // committee index per record (caller-given or produced by k_key_lookup); nullptr = none
  const uint8_t *msg;   // message i: off ? msg + off[i] : msg + (midx ? midx[i] : i) * msg_stride
  size_t msg_stride;
  const uint32_t *midx;
  const uint64_t *off;
  uint64_t fixed_len;   // message length when off == nullptr
  int aos128;           // 1: sig/pk/msg are the fields of packed 128-byte records at `sig` (coalesced staged loads)
};

__device__ __forceinline__ void load32(uint32_t (&w)[8], const uint8_t *p) {
  const uint32_t *s = reinterpret_cast<const uint32_t *>(p);  // every 32-byte field is 4-byte aligned in all layouts
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = __ldg(s + i);
}

// A warp loads 32 rows of 128 bytes (row r at base + (warp_first + r) * stride, 16-byte aligned) with fully coalesced
// 16-byte accesses, parks them in shared memory under an XOR swizzle, and every lane then reads back its own row
// conflict-free.  Rows past n read as zeros.  Every lane of the warp must call it.
__device__ __forceinline__ void warp_stage_rows128(uint4 (&q)[8], const uint8_t *__restrict__ base, size_t stride, size_t n, size_t warp_first,
                                                   uint4 *smem_warp) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int c = j * 32 + lane;  // chunk index inside the warp's 4 KB
    int rec = c >> 3, part = c & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (warp_first + rec < n) v = __ldg(reinterpret_cast<const uint4 *>(base + (warp_first + rec) * stride) + part);
    smem_warp[rec * 8 + (part ^ (rec & 7))] = v;
  }
  __syncwarp();
#pragma unroll
  for (int part = 0; part < 8; part++) q[part] = smem_warp[lane * 8 + (part ^ (lane & 7))];
  __syncwarp();
}
// packed hs_rec128 records: sig.R | sig.S | pk | msg
__device__ __forceinline__ void warp_load_rec128(uint32_t (&sig_r)[8], uint32_t (&sig_s)[8], uint32_t (&pk)[8], uint32_t (&msg)[8],
                                                 const uint4 *__restrict__ recs, size_t n, size_t warp_first, uint4 *smem_warp) {
  uint4 q[8];
  warp_stage_rows128(q, reinterpret_cast<const uint8_t *>(recs), 128, n, warp_first, smem_warp);
  sig_r[0] = q[0].x; sig_r[1] = q[0].y; sig_r[2] = q[0].z; sig_r[3] = q[0].w; sig_r[4] = q[1].x; sig_r[5] = q[1].y; sig_r[6] = q[1].z; sig_r[7] = q[1].w;
  sig_s[0] = q[2].x; sig_s[1] = q[2].y; sig_s[2] = q[2].z; sig_s[3] = q[2].w; sig_s[4] = q[3].x; sig_s[5] = q[3].y; sig_s[6] = q[3].z; sig_s[7] = q[3].w;
  pk[0] = q[4].x; pk[1] = q[4].y; pk[2] = q[4].z; pk[3] = q[4].w; pk[4] = q[5].x; pk[5] = q[5].y; pk[6] = q[5].z; pk[7] = q[5].w;
  msg[0] = q[6].x; msg[1] = q[6].y; msg[2] = q[6].z; msg[3] = q[6].w; msg[4] = q[7].x; msg[5] = q[7].y; msg[6] = q[7].z; msg[7] = q[7].w;
}

// ------------------------------------------------------------------------------------------------ committee key lookup
// Open-addressing hash table over the registered keys: slot -> key index (HS_NO_KEY = empty).  Built on the host at
// registration (hashing only), probed here by one thread per record.
struct key_table {
  const uint32_t *slots;
  uint32_t mask;  // capacity - 1 (power of two)
  const uint8_t *pks;
  uint32_t n_keys;
};
__host__ __device__ inline uint32_t key_hash(const uint32_t *w) {
  uint32_t h = 0x9e3779b9u;
  for (int i = 0; i < 8; i++) {
    h ^= w[i];
    h *= 0x85ebca6bu;
    h ^= h >> 15;
  }
  return h;
}
__global__ void __launch_bounds__(256) k_key_lookup(in_layout L, size_t n, key_table T, uint32_t *__restrict__ out_vidx,
                                                    uint32_t *__restrict__ miss_list, uint32_t *__restrict__ miss_count) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  load32(k, L.pk + i * L.pk_stride);
  uint32_t found = HS_NO_KEY;
  uint32_t h = key_hash(k) & T.mask;
  for (uint32_t probe = 0; probe <= T.mask; probe++) {
    uint32_t idx = __ldg(T.slots + h);
    if (idx == HS_NO_KEY) break;
    const uint32_t *cand = reinterpret_cast<const uint32_t *>(T.pks + (size_t)idx * 32);
    uint32_t diff = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) diff |= __ldg(cand + j) ^ k[j];
    if (diff == 0) {
      found = idx;
      break;
    }
    h = (h + 1) & T.mask;
  }
  out_vidx[i] = found;
  if (found == HS_NO_KEY) miss_list[atomicAdd(miss_count, 1u)] = (uint32_t)i;  // compacted list for the generic pass
}

// ------------------------------------------------------------------------------------------------ key cache: collect unknown keys
// Copies the key bytes of (up to `max_keys`) records that missed the lookup — or, when nothing is cached yet, of the first
// records of the call — into a compact buffer that is read back asynchronously; the host dedupes them before the NEXT call
// and builds their tables (hs_engine.cu: learn_process).  miss_count == nullptr: take records 0 .. n-1 directly.
__global__ void __launch_bounds__(256) k_gather_keys(in_layout L, size_t n, const uint32_t *__restrict__ miss_list,
                                                     const uint32_t *__restrict__ miss_count, uint32_t max_keys, uint8_t *__restrict__ out_keys,
                                                     uint32_t *__restrict__ out_n) {
  const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t avail = miss_count ? (size_t)*miss_count : n;
  const size_t take = avail < max_keys ? avail : max_keys;
  if (t == 0) *out_n = (uint32_t)take;
  if (t >= take) return;
  const size_t i = miss_count ? miss_list[t] : t;
  uint32_t k[8];
  load32(k, L.pk + i * L.pk_stride);
  uint32_t *dst = reinterpret_cast<uint32_t *>(out_keys + t * 32);
#pragma unroll
  for (int j = 0; j < 8; j++) dst[j] = k[j];
}

// ------------------------------------------------------------------------------------------------ phase 1: main
struct main_out {
  fe *xyz;         // 3 field elements per record: X, Y, Z of R' = [S]B + [k](-A)
  uint8_t *meta;   // HS_META_* per record
  int side_pass;   // 1: records without a registered key are handled by the concurrent generic pass
};
struct committee_tables {
  const uint8_t *pks;
  const uint8_t *key_flags;
  uint32_t n_keys;
  const ge_niels *atables;
  size_t table_entries;  // ge_niels per key
};

// Grid-stride over records so the same kernel serves a full launch (one pass) and the compacted miss list, whose length
// only the device knows (n_ptr): no host round trip between the committee pass and the generic pass.
template <bool COMMITTEE>
#ifndef HS_MAIN_MINBLOCKS
#define HS_MAIN_MINBLOCKS 4
#endif
#ifndef HS_GENERIC_MINBLOCKS
#define HS_GENERIC_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(HS_THREADS, COMMITTEE ? HS_MAIN_MINBLOCKS : HS_GENERIC_MINBLOCKS) k_verify_main(in_layout L, size_t n_arg, const uint32_t *__restrict__ n_ptr,
                                                             const uint32_t *__restrict__ index_list, const ge_niels *__restrict__ btable,
                                                             committee_tables C, main_out O, const comb_params cp) {
  // one buffer, two lives: record staging while loading, then the signed digits [digit][thread] (conflict-free columns)
  __shared__ __align__(16) unsigned char smem_raw[HS_MAX_DIGITS * HS_THREADS * 4];
  static_assert(sizeof(smem_raw) >= (HS_THREADS / 32) * 256 * sizeof(uint4), "staging does not fit");
  uint4(*stage)[256] = reinterpret_cast<uint4(*)[256]>(smem_raw);
  int32_t *digits = reinterpret_cast<int32_t *>(smem_raw) + threadIdx.x;
  const size_t n = n_ptr ? (size_t)*n_ptr : n_arg;
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {  // blockDim <= HS_THREADS
  const size_t t = base + threadIdx.x;
  const size_t warp_first = t & ~(size_t)31;
  // (no early exit for warps past the end: every warp of the block reaches the barriers below; they clamp and do not store)
  const bool active = t < n;
  size_t i = active ? t : n - 1;
  uint32_t R[8], S[8], A[8], h[16];
  uint32_t v = 0;
  bool have_key = true;
  if (L.aos128 && !index_list) {
    uint32_t M[8];
    __syncthreads();  // the previous grid-stride iteration's digits live in the same shared bytes
    warp_load_rec128(R, S, A, M, reinterpret_cast<const uint4 *>(L.sig), n, warp_first, stage[threadIdx.x >> 5]);
    __syncthreads();
    if (COMMITTEE) {
      v = __ldg(L.vidx + i);
      have_key = v < C.n_keys;
      if (!have_key) v = 0;
      load32(A, C.pks + (size_t)v * 32);  // hash the registered key bytes (identical to the record's on a lookup hit)
    }
    sha512_ram32(h, R, A, M);
  } else {
    if (index_list) i = index_list[i];
    load32(R, L.sig + i * L.sig_stride);
    load32(S, L.sig + i * L.sig_stride + 32);
    if (COMMITTEE) {
      v = __ldg(L.vidx + i);
      have_key = v < C.n_keys;
      if (!have_key) v = 0;
      load32(A, C.pks + (size_t)v * 32);
    } else {
      load32(A, L.pk + i * L.pk_stride);
    }
    const uint8_t *m = L.off ? L.msg + L.off[i] : L.msg + (size_t)(L.midx ? __ldg(L.midx + i) : i) * L.msg_stride;
    const uint64_t len = L.off ? (L.off[i + 1] - L.off[i]) : L.fixed_len;
    if (len == 32 && ((reinterpret_cast<uintptr_t>(m) & 3u) == 0)) {
      uint32_t M[8];
      load32(M, m);
      sha512_ram32(h, R, A, M);
    } else {
      uint64_t pre[8];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        pre[j] = be64_from_le32(R[2 * j], R[2 * j + 1]);
        pre[4 + j] = be64_from_le32(A[2 * j], A[2 * j + 1]);
      }
      sha512_prefix_msg(h, pre, 8, m, len);
    }
  }
  ge_ext acc;
  uint32_t meta;
  if (COMMITTEE) {
    meta = verify_committee_main(acc, R, S, h, btable, C.atables + (size_t)v * C.table_entries, have_key ? C.key_flags[v] : 0u, digits,
                                 HS_THREADS, cp);
    if (!have_key) {
      if (O.side_pass) continue;  // unknown key bytes: the generic pass on the side stream owns this record's outputs
      meta = 0;                   // unknown authority INDEX: reject (messages.rs:57-61 rejects it before any crypto)
    }
  } else {
    ge_cached tab[9];
    meta = verify_generic_main(acc, R, S, A, h, btable, tab, digits, HS_THREADS, cp);
  }
  if (!active) continue;
  if (!(meta & HS_META_PARSE_OK)) {  // keep the batched inversion well-defined for rejected records
    fe_set0(acc.X);
    fe_set1(acc.Y);
    fe_set1(acc.Z);
  }
  uint4 *dst = reinterpret_cast<uint4 *>(O.xyz + i * 3);
  dst[0] = make_uint4(acc.X.v[0], acc.X.v[1], acc.X.v[2], acc.X.v[3]);
  dst[1] = make_uint4(acc.X.v[4], acc.X.v[5], acc.X.v[6], acc.X.v[7]);
  dst[2] = make_uint4(acc.Y.v[0], acc.Y.v[1], acc.Y.v[2], acc.Y.v[3]);
  dst[3] = make_uint4(acc.Y.v[4], acc.Y.v[5], acc.Y.v[6], acc.Y.v[7]);
  dst[4] = make_uint4(acc.Z.v[0], acc.Z.v[1], acc.Z.v[2], acc.Z.v[3]);
  dst[5] = make_uint4(acc.Z.v[4], acc.Z.v[5], acc.Z.v[6], acc.Z.v[7]);
  O.meta[i] = (uint8_t)meta;
  }
}

// ------------------------------------------------------------------------------------------------ latency path (n <= 64)
// One Block::verify / Vote::verify / 3-vote QC of a 4-node deployment is 1 .. 64 signatures (BASELINE config[4]): the
// throughput kernels above would spend 5 launches and ~28 serial mixed additions + an inversion on it (r1: 155 us per
// verify).  Here ONE launch of 64-thread blocks does a signature per block: warp 0 hashes, recodes, lets lane j fetch the table
// entry of digit j and sums the lanes' points with a shuffle tree (5 levels); warp 1 decompresses R meanwhile; thread 0
// compares projectively.  Inputs and verdicts live in mapped pinned host memory (no copy calls); the last block raises a
// completion word the host polls.  Registered / cached keys only (the host resolves key bytes to indices first).
struct small_rec {
  uint8_t sig[64];
  uint8_t msg[32];
  uint32_t vidx;
  uint32_t pad[7];
};
static_assert(sizeof(small_rec) == 128, "small_rec is 128 bytes");
#define HS_SMALL_MAX 64
__device__ __forceinline__ void fe_shfl_down(fe &r, const fe &a, int delta) {
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], delta);
}
__global__ void __launch_bounds__(64) k_verify_small(const small_rec *__restrict__ in, uint32_t n, const ge_niels *__restrict__ btable,
                                                      committee_tables C, const comb_params cp, uint8_t *out_flags, uint32_t *counter,
                                                      volatile uint32_t *done, uint32_t seq) {
  __shared__ int32_t dig[HS_MAX_DIGITS];
  __shared__ fe sh_acc[3], sh_r[2];
  __shared__ uint32_t sh_meta[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const small_rec *rec = in + blockIdx.x;
  uint32_t R[8], S[8];
  load32(R, rec->sig);
  load32(S, rec->sig + 32);
  uint32_t v = rec->vidx;
  const bool have_key = v < C.n_keys;
  if (!have_key) v = 0;
  const uint32_t a_flags = have_key ? C.key_flags[v] : 0u;
  if (warp == 0) {
    uint32_t A[8], M[8], h[16], k[8];
    load32(A, C.pks + (size_t)v * 32);
    load32(M, rec->msg);
    sha512_ram32(h, R, A, M);
    sc_reduce512(k, h);
    if (lane == 0) {  // one writer for the shared digit array
      sc_digits_rt(dig, 1, k, cp.bias_a, cp.wa, cp.na);
      sc_digits_rt(dig + cp.na, 1, S, cp.bias_b, cp.wb, cp.nb);
    }
    __syncwarp();
    const int NT = cp.na + cp.nb;
    ge_ext acc;
    ge_identity(acc);
#pragma unroll 1
    for (int j = lane; j < NT; j += 32) {  // NT <= 32 for every production geometry: one entry per lane
      uint32_t neg;
      bool is_a;
      const ge_niels *e = comb_entry(C.atables + (size_t)v * C.table_entries, btable, dig, 1, j, cp, neg, is_a);
      niels_signed q;
      niels_load_signed(q, e, neg, false);
      ge_ext p;
      ge_from_signed_niels(p, q.m0, q.m1);
      if (j == lane) acc = p;
      else ge_add_ext(acc, acc, p);
    }
#pragma unroll 1
    for (int step = 1; step < 32; step <<= 1) {
      ge_ext o;
      fe_shfl_down(o.X, acc.X, step);
      fe_shfl_down(o.Y, acc.Y, step);
      fe_shfl_down(o.Z, acc.Z, step);
      fe_shfl_down(o.T, acc.T, step);
      ge_add_ext(acc, acc, o);
    }
    if (lane == 0) {
      sh_acc[0] = acc.X;
      sh_acc[1] = acc.Y;
      sh_acc[2] = acc.Z;
    }
  } else {
    ge_ext Rpt;
    const uint32_t r_ok = ge_decompress(Rpt, R);
    const uint32_t parse_ok = sc_is_canonical(S) & a_flags & 1u & (have_key ? 1u : 0u);
    const uint32_t small = ge_enc_is_small_order(R) | ((a_flags >> 1) & 1u);
    if (lane == 0) {
      sh_r[0] = Rpt.X;
      sh_r[1] = Rpt.Y;
      sh_meta[0] = parse_ok & r_ok;
      sh_meta[1] = (parse_ok ? HS_F_PARSE_OK : 0u) | (small ? HS_F_SMALL : 0u);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t fl = sh_meta[1];
    const uint32_t eq = sh_meta[0] & ge_proj_equals_affine(sh_acc[0], sh_acc[1], sh_acc[2], sh_r[0], sh_r[1]);
    if (eq) fl |= HS_F_EQ;
    if (eq && !(fl & HS_F_SMALL)) fl |= HS_F_STRICT;
    out_flags[blockIdx.x] = (uint8_t)fl;
    __threadfence_system();
    if (atomicAdd(counter, 1u) == n - 1) {
      *counter = 0;
      __threadfence_system();
      *done = seq;
    }
  }
}

// A few LONG messages (one mempool batch is ~15 kB = 120 blocks, mempool/src/processor.rs:30): SHA-512 is sequential in its
// 80 x nblk rounds, but the message schedule (45 % of the work) of different blocks is independent — lane l of the warp
// expands block g + l into a shared K+W table, then the rounds run back to back from that table.  One warp per message.
__global__ void __launch_bounds__(32) k_digest32_long(const uint8_t *__restrict__ data, const uint64_t *__restrict__ off, size_t n,
                                                       uint32_t *__restrict__ out) {
  __shared__ uint64_t kw[80 * 32];
  const size_t i = blockIdx.x;
  if (i >= n) return;
  const int lane = threadIdx.x;
  const uint8_t *m = data + off[i];
  const uint64_t len = off[i + 1] - off[i];
  const uint64_t nblk = sha512_nblocks(len);
  sha512_state s;
  sha512_init(s);
#pragma unroll 1
  for (uint64_t g0 = 0; g0 < nblk; g0 += 32) {
    if (g0 + lane < nblk) {
      uint64_t w[16];
      sha512_block_words(w, m, len, g0 + lane);
      sha512_expand_kw(kw + lane, 32, w);
    }
    __syncwarp();
    const int cnt = (int)((nblk - g0 < 32) ? (nblk - g0) : 32);
#pragma unroll 1
    for (int j = 0; j < cnt; j++) sha512_compress_kw_strided(s, kw + j, 32);  // every lane runs the same rounds (broadcast reads)
    __syncwarp();
  }
  if (lane == 0) {
    uint32_t h[16];
    sha512_output_words(s, h);
    uint4 *dst = reinterpret_cast<uint4 *>(out + i * 8);
    dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
    dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
  }
}

// ------------------------------------------------------------------------------------------------ multi-GPU epilogue
// The accept bitmap of a sharded verify has to reach every rank (each validator process needs every verdict).  Instead of a
// separate all-gather collective after the kernel, the finish kernel stores each bitmap word it produces straight into EVERY
// peer's result buffer over NVLink (P2P stores through CUDA-IPC mapped pointers), then a release flag per (writer, reader)
// pair tells the reader the shard has landed.  Payload is n/8 bytes per rank: latency, not bandwidth.
#define HS_MAX_PEERS 16
// Result buffer of one rank (cudaMalloc'd, exported over CUDA IPC):
//   [2][total_words]  the global bitmap, double-buffered by epoch parity: a fast rank's epoch e+1 words land in the OTHER
//                     half, so a slower rank that is still reading epoch e never sees them (r1's single buffer had a
//                     write-after-read hazard); a rank can run at most one epoch ahead, because finishing epoch e+1
//                     needs every peer's epoch e+1 flag, which a peer publishes only after its own epoch-e readers ran
//                     (stream order: consume epoch e's bitmap before enqueueing the verify of epoch e+1)
//   [HS_MAX_PEERS]    flags[w] = last epoch whose words from writer w have landed here (release/acquire, system scope)
//   [0] timeout flag, [1] finish-kernel block counter
struct peer_route {
  uint32_t *buf[HS_MAX_PEERS];  // buf[p] = base of rank p's result buffer as mapped in THIS process
  int n;                        // world size (0 = route disabled: plain local bitmap)
  int my_rank;
  uint32_t epoch;
  size_t total_words;           // words of the global bitmap; rank p owns [p * total_words / n, (p + 1) * total_words / n)
  size_t word_offset;           // this rank's first word in the global bitmap
};
#define HS_PEER_FLAGS(P) ((P).total_words * 2)
#define HS_PEER_CTRL(P) ((P).total_words * 2 + HS_MAX_PEERS)
// Executed by ONE block after all of this rank's words of the epoch are stored (and fenced): thread p publishes this rank's
// flag in rank p's buffer, then waits for rank p's flag here.  A peer that never shows up is an error, not a hang: the spin
// is bounded, the sticky timeout flag is raised and that peer's shard is cleared (every verdict reads "reject").
__device__ __forceinline__ void peer_signal_and_wait(const peer_route &P) {
  const int p = threadIdx.x;
  if (p >= P.n) return;
  uint32_t *own = P.buf[P.my_rank];
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(P.buf[p] + HS_PEER_FLAGS(P) + P.my_rank), "r"(P.epoch) : "memory");
  uint32_t v = 0;
  bool ok = false;
  for (long long spin = 0; spin < (1ll << 24); spin++) {  // bounded (~5 s)
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(own + HS_PEER_FLAGS(P) + p) : "memory");
    if ((int32_t)(v - P.epoch) >= 0) {
      ok = true;
      break;
    }
    __nanosleep(200);
  }
  if (!ok) {
    own[HS_PEER_CTRL(P)] = 1;
    const size_t per = P.total_words / P.n;
    uint32_t *w = own + (P.epoch & 1u) * P.total_words + (size_t)p * per;
    for (size_t k = 0; k < per; k++) w[k] = 0;
  }
}
// shard without records: nothing to verify, but the peers still wait for this rank's flag
__global__ void k_peer_sync_only(const peer_route P) { peer_signal_and_wait(P); }

// ------------------------------------------------------------------------------------------------ phase 2: finish
__device__ __forceinline__ void fe_load_global(fe &r, const fe *p) {
  const uint4 *s = reinterpret_cast<const uint4 *>(p);
  uint4 a = s[0], b = s[1];
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
}
// One thread owns `group` (16, 8 or 4: fewer for small batches, so that enough blocks exist to hide the one serial field inversion
// each block waits for — at 126 k records the 16-record form ran 62 blocks for 110 us) consecutive records.  Montgomery's trick at two levels so that ONE
// field inversion per 64 records is executed (by warp 0, lane l inverting the product of threads 4l..4l+3) instead of one
// per thread: phase A prefix products of the thread's 16 Z's; phase B the block-level inversion through shared memory;
// phase C back-substitution + affine comparison with R's encoding.  Two neighbouring lanes combine their 16 verdicts into
// one bitmap word, which goes to the local bitmap or — armed by hs_peer_next — straight into every peer's buffer, after
// which the last block of the grid exchanges the epoch flags with the peers (no separate signal / wait launches).
__global__ void __launch_bounds__(HS_THREADS) k_verify_finish(in_layout L, size_t n, const fe *__restrict__ xyz, const uint8_t *__restrict__ meta,
                                                               uint32_t mode, uint32_t *__restrict__ bitmap, uint8_t *flags_out, const peer_route P,
                                                               const int group) {
  __shared__ fe tot[HS_THREADS];
  const size_t t = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t first = t * (size_t)group;
  const int cnt = (first < n) ? (int)((n - first < (size_t)group) ? (n - first) : group) : 0;
  fe prod[HS_FINISH_GROUP];
  fe run;
  fe_set1(run);
#pragma unroll 1
  for (int c = 0; c < cnt; c++) {
    fe Z;
    fe_load_global(Z, xyz + (first + c) * 3 + 2);
    if (fe_is_zero(Z)) fe_set1(Z);  // cannot happen for curve points; keeps one bad record from poisoning the group
    fe_mul(run, run, Z);
    prod[c] = run;
  }
  tot[threadIdx.x] = run;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int b = threadIdx.x * 4;
    fe q0 = tot[b], q1, q2, q3, inv, u;
    fe_mul(q1, q0, tot[b + 1]);
    fe_mul(q2, q1, tot[b + 2]);
    fe_mul(q3, q2, tot[b + 3]);
    fe_invert(inv, q3);
    fe_mul(u, inv, q2);        // 1 / tot[b+3]
    fe_mul(inv, inv, tot[b + 3]);
    tot[b + 3] = u;
    fe_mul(u, inv, q1);        // 1 / tot[b+2]
    fe_mul(inv, inv, tot[b + 2]);
    tot[b + 2] = u;
    fe_mul(u, inv, q0);        // 1 / tot[b+1]
    fe_mul(inv, inv, tot[b + 1]);
    tot[b + 1] = u;
    tot[b] = inv;              // 1 / tot[b]
  }
  __syncthreads();
  uint32_t bits = 0;
  if (cnt) {
    fe u = tot[threadIdx.x];
#pragma unroll 1
    for (int c = cnt - 1; c >= 0; c--) {
      const size_t i = first + c;
      fe X, Y, Z, zinv;
      fe_load_global(X, xyz + i * 3 + 0);
      fe_load_global(Y, xyz + i * 3 + 1);
      fe_load_global(Z, xyz + i * 3 + 2);
      const uint32_t zero_z = fe_is_zero(Z);
      if (zero_z) fe_set1(Z);
      if (c > 0) fe_mul(zinv, u, prod[c - 1]);
      else zinv = u;
      fe_mul(u, u, Z);
      uint32_t R[8];
      load32(R, L.sig + i * L.sig_stride);
      uint32_t m = meta[i];
      if (zero_z) m &= ~HS_META_PARSE_OK;
      const uint32_t fl = verify_flags_from(X, Y, zinv, R, m);
      if (flags_out) flags_out[i] = (uint8_t)fl;
      const uint32_t ok = (mode == HS_MODE_STRICT) ? (fl & HS_F_STRICT) : (fl & HS_F_EQ);
      if (ok) bits |= 1u << c;
    }
  }
  // 32 / group neighbouring lanes hold the verdicts of one bitmap word: butterfly-OR them together
  const int lanes_per_word = 32 / group;
  uint32_t word = bits << (group * (threadIdx.x & (lanes_per_word - 1)));
  for (int m = 1; m < lanes_per_word; m <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, m);
  if ((threadIdx.x & (lanes_per_word - 1)) == 0 && first < n) {
    const size_t widx = t / lanes_per_word;
    if (P.n == 0) {
      bitmap[widx] = word;
    } else {
      const size_t at = (P.epoch & 1u) * P.total_words + P.word_offset + widx;
#pragma unroll 1
      for (int p = 0; p < P.n; p++) P.buf[p][at] = word;  // fused all-gather: one NVLink store per peer
    }
  }
  if (P.n) {
    // the last block to get here publishes the epoch flag to every peer and waits for theirs: the exchange costs no launch
    __shared__ int is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(P.buf[P.my_rank] + HS_PEER_CTRL(P) + 1, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
      peer_signal_and_wait(P);
      __syncthreads();
      if (threadIdx.x == 0) P.buf[P.my_rank][HS_PEER_CTRL(P) + 1] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------ table construction
// thread = (point p, window w, block b of HS_BUILD_BLOCK entries)
#define HS_BUILD_BLOCK 64
__global__ void __launch_bounds__(HS_THREADS) k_build_comb(const uint8_t *__restrict__ encs, size_t n_points, int negate, int W, int n_windows,
                                                            ge_niels *tables, uint8_t *key_flags) {
  const int entries = 1 << (W - 1);
  const int blocks_per_window = entries / HS_BUILD_BLOCK;
  const size_t t = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t per_point = (size_t)n_windows * blocks_per_window;
  const size_t p = t / per_point;
  if (p >= n_points) return;
  const int w = (int)((t % per_point) / blocks_per_window);
  const int b = (int)(t % blocks_per_window);
  ge_ext P;
  if (encs) {
    uint32_t e[8];
    load32(e, encs + p * 32);
    uint32_t ok = ge_decompress(P, e);
    uint32_t small = ge_enc_is_small_order(e);
    if (w == 0 && b == 0 && key_flags) key_flags[p] = (uint8_t)((ok & 1u) | (small << 1));
    if (!ok) ge_identity(P);  // the table of a rejected key is never used for an accept (flag bit0 = 0)
  } else {
    ge_basepoint(P);
  }
  if (negate) {
    ge_ext Q;
    ge_neg(Q, P);
    P = Q;
  }
  fe prod[HS_BUILD_BLOCK];
  comb_build_block(tables + p * ((size_t)n_windows * comb_window_stride(W)), P, W, w, b * HS_BUILD_BLOCK, HS_BUILD_BLOCK, prod);
}

// ------------------------------------------------------------------------------------------------ Digest kernels
__global__ void __launch_bounds__(HS_THREADS) k_digest32(const uint8_t *__restrict__ data, const uint64_t *__restrict__ off, uint64_t fixed_len,
                                                          size_t n, uint32_t *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  if (i >= n) return;
  uint32_t h[16];
  uint64_t pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint8_t *m = off ? data + off[i] : data + i * fixed_len;
  const uint64_t len = off ? off[i + 1] - off[i] : fixed_len;
  sha512_prefix_msg(h, pre, 0, m, len);
  uint4 *dst = reinterpret_cast<uint4 *>(out + i * 8);
  dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
  dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// Fixed-size, 16-byte aligned messages (the transaction / payload shape of BASELINE config[1]): every full 128-byte block is
// fetched by the warp with coalesced 16-byte loads through shared memory (a per-thread 8-byte walk touches 32 sectors per
// load instruction), and when the length is a multiple of 128 the padding-only last block runs without its message
// schedule (sha512_compress_kw).  Other lengths finish through the generic reader.
__global__ void __launch_bounds__(HS_THREADS) k_digest32_fixed(const uint8_t *__restrict__ data, uint64_t len, size_t n, uint32_t *__restrict__ out,
                                                                const sha512_kw padkw, int pad_is_const) {
  __shared__ __align__(16) uint4 stage[HS_THREADS / 32][256];
  const size_t i = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t warp_first = i & ~(size_t)31;
  if (warp_first >= n) return;  // whole warp past the end
  sha512_state s;
  sha512_init(s);
  const uint64_t nfull = len >> 7;
#pragma unroll 1
  for (uint64_t b = 0; b < nfull; b++) {
    uint4 q[8];
    warp_stage_rows128(q, data + b * 128, len, n, warp_first, stage[threadIdx.x >> 5]);
    uint64_t w[16];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      w[2 * k] = be64_from_le32(q[k].x, q[k].y);
      w[2 * k + 1] = be64_from_le32(q[k].z, q[k].w);
    }
    sha512_compress(s, w);
  }
  if (i >= n) return;
  if (pad_is_const) {
    sha512_compress_kw(s, padkw);
  } else {
    const uint64_t pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sha512_absorb_blocks(s, pre, 0, data + i * len, len, nfull, sha512_nblocks(len));
  }
  uint32_t h[16];
  sha512_output_words(s, h);
  uint4 *dst = reinterpret_cast<uint4 *>(out + i * 8);
  dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
  dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// ------------------------------------------------------------------------------------------------ per-QC AND
// vote i belongs to certificate qc_idx[i]; a rejected vote clears its certificate's bit (qc bitmap pre-set to all ones)
__global__ void __launch_bounds__(256) k_qc_and(const uint32_t *__restrict__ vote_bitmap, const uint32_t *__restrict__ qc_idx, size_t n_votes,
                                                size_t n_qc, uint32_t *__restrict__ qc_bitmap) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_votes) return;
  if (!((vote_bitmap[i >> 5] >> (i & 31)) & 1u)) {
    const uint32_t j = qc_idx[i];
    if (j < n_qc) atomicAnd(qc_bitmap + (j >> 5), ~(1u << (j & 31)));
  }
}

// item i belongs to group grp[i]; its verdict is flag bit STRICT or EQ by mode[i] (nullptr = all strict).  A rejected item
// clears its group's bit (group bitmap pre-set to ones); the item verdicts are also packed into item_bitmap (nullable).
__global__ void __launch_bounds__(256) k_group_and(const uint8_t *__restrict__ flags, const uint8_t *__restrict__ mode, const uint32_t *__restrict__ grp,
                                                   size_t n_items, size_t n_groups, uint32_t *__restrict__ item_bitmap, uint32_t *__restrict__ group_bitmap) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  uint32_t ok = 0;
  if (i < n_items) {
    const uint32_t want = (mode && mode[i] == HS_MODE_BATCH_EQ) ? HS_F_EQ : HS_F_STRICT;
    ok = (flags[i] & want) ? 1u : 0u;
    if (!ok) {
      const uint32_t j = grp[i];
      if (j < n_groups) atomicAnd(group_bitmap + (j >> 5), ~(1u << (j & 31)));
    }
  }
  const uint32_t word = __ballot_sync(0xffffffffu, ok);
  if (item_bitmap && (threadIdx.x & 31) == 0 && i < n_items) item_bitmap[i >> 5] = word;
}
// all-ones bitmap over n bits (unused high bits of the last word 0)
__global__ void k_bitmap_ones(uint32_t *bm, size_t n) {
  const size_t w = (size_t)blockIdx.x * 256 + threadIdx.x;
  const size_t words = (n + 31) / 32;
  if (w < words) bm[w] = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
}
// TC::verify (consensus/src/messages.rs:307-311) and Timeout::digest (:268-275): the message of vote i is
// SHA-512(tc_round_le || high_qc_round_le)[..32] — 16 bytes of which 8 differ per vote; built and hashed here from the two
// integers (one padded block), so the host ships 8 bytes per vote instead of a digest.
__global__ void __launch_bounds__(HS_THREADS) k_tc_digests(const uint64_t *__restrict__ tc_round, const uint32_t *__restrict__ tc_idx,
                                                            const uint64_t *__restrict__ high_qc_round, size_t n, size_t n_tc, uint32_t *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = tc_idx ? tc_idx[i] : (uint32_t)i;
  const uint64_t r = t < n_tc ? tc_round[t] : 0, hq = high_qc_round[i];
  uint64_t w[16];
  // to_le_bytes() then read as big-endian message words = byte swap
  w[0] = ((uint64_t)bswap32((uint32_t)r) << 32) | bswap32((uint32_t)(r >> 32));
  w[1] = ((uint64_t)bswap32((uint32_t)hq) << 32) | bswap32((uint32_t)(hq >> 32));
  w[2] = 0x8000000000000000ULL;
#pragma unroll
  for (int j = 3; j < 15; j++) w[j] = 0;
  w[15] = 16 * 8;
  sha512_state s;
  sha512_init(s);
  sha512_compress(s, w);
  uint32_t h[16];
  sha512_output_words(s, h);
  uint4 *dst = reinterpret_cast<uint4 *>(out + i * 8);
  dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
  dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
}

// ------------------------------------------------------------------------------------------------ load generation: keygen / sign
// RFC 8032 key generation and signing of 32-byte digests (verify_core.cuh: keygen_core / sign_digest_core).  Not on the node's
// path (the reference signs one message per request on the CPU); used to synthesise benchmark and test inputs.
__global__ void __launch_bounds__(HS_THREADS) k_keygen(const uint8_t *__restrict__ seeds, size_t n, const ge_niels *__restrict__ btable,
                                                        const comb_params cp, uint8_t *__restrict__ pks) {
  __shared__ int32_t digits_s[HS_MAX_DIGITS * HS_THREADS];
  const size_t i = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  if (i >= n) return;
  uint32_t sd[8], A[8];
  load32(sd, seeds + i * 32);
  keygen_core(A, sd, btable, digits_s + threadIdx.x, HS_THREADS, cp);
  uint32_t *dst = reinterpret_cast<uint32_t *>(pks + i * 32);
#pragma unroll
  for (int j = 0; j < 8; j++) dst[j] = A[j];
}
__global__ void __launch_bounds__(HS_THREADS) k_sign_digests(const uint8_t *__restrict__ seeds, const uint8_t *__restrict__ pks,
                                                              const uint32_t *__restrict__ key_idx, const uint8_t *__restrict__ digests, size_t n,
                                                              size_t n_keys, const ge_niels *__restrict__ btable, const comb_params cp,
                                                              uint8_t *__restrict__ sig) {
  __shared__ int32_t digits_s[HS_MAX_DIGITS * HS_THREADS];
  const size_t i = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  if (i >= n) return;
  uint32_t k = key_idx ? key_idx[i] : (uint32_t)i;
  if (k >= n_keys) k = 0;
  uint32_t sd[8], A[8], M[8], R[8], S[8];
  load32(sd, seeds + (size_t)k * 32);
  load32(A, pks + (size_t)k * 32);
  load32(M, digests + i * 32);
  sign_digest_core(R, S, sd, A, M, btable, digits_s + threadIdx.x, HS_THREADS, cp);
  uint4 *dst = reinterpret_cast<uint4 *>(sig + i * 64);
  dst[0] = make_uint4(R[0], R[1], R[2], R[3]);
  dst[1] = make_uint4(R[4], R[5], R[6], R[7]);
  dst[2] = make_uint4(S[0], S[1], S[2], S[3]);
  dst[3] = make_uint4(S[4], S[5], S[6], S[7]);
}

// ================================================================================================ host side
struct dev_buf {
  void *p = nullptr;
  size_t cap = 0;
};
struct hs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr, stream_side = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_side[2] = {nullptr, nullptr};
  ge_niels *d_btable = nullptr;
  comb_params cp{};
  size_t a_table_entries = 0;
  int wa_forced = 0;
  // committee
  size_t n_keys = 0;
  uint8_t *d_pks = nullptr;
  uint8_t *d_key_flags = nullptr;
  ge_niels *d_atables = nullptr;
  uint32_t *d_slots = nullptr;
  uint32_t slot_mask = 0;
  // grow-only device scratch
  dev_buf in[2], digest[2], xyz, meta, vidx, miss, out;
  uint32_t *d_miss_count = nullptr;
  uint32_t *h_miss_count = nullptr;  // pinned
  // key cache: tables for keys that were never registered but keep showing up (learned between calls)
  bool explicit_committee = false;   // hs_committee_register was called with keys: the set is fixed, nothing is learned
  bool cache_wanted = true, cache_enabled = true;
  size_t cache_cap = 4096;           // keys
  std::vector<uint8_t> h_pks;        // host mirrors of d_pks / d_slots (key cache and hs_committee_update)
  std::vector<uint32_t> h_slots;
  std::vector<uint8_t> h_key_live;   // explicit committee: 1 = slot holds a live validator, 0 = removed (free for reuse)
  size_t table_budget = 0;           // bytes the per-key tables may use (0 = ~62 % of the device)
  size_t key_capacity = 0;           // explicit committee: table slots allocated (>= n_keys; spare slots serve hs_committee_update)
  uint8_t *d_learn_keys = nullptr, *h_learn_keys = nullptr;
  uint32_t *d_learn_n = nullptr, *h_learn_n = nullptr;
  cudaEvent_t ev_learn = nullptr;
  cudaEvent_t ev_tables = nullptr;   // recorded after the latest table build of the key cache; every pass waits for it (any stream)
  bool learn_pending = false;
  bool cache_full = false;           // no free slot: only the miss RATE is watched (a mostly-missing full cache is reset)
  uint64_t calls_since_reset = HS_CACHE_RESET_MIN_CALLS;
  size_t learn_records = 0;          // records of the pass whose misses are parked in h_learn_*
  uint32_t *h_miss_total = nullptr;  // pinned: total misses of that pass
  // multi-GPU peer routing
  peer_route peers{};
  int peer_rank = 0;
  size_t peer_total_words = 0;
  uint32_t *peer_own = nullptr;       // cudaMalloc'd: [2][total_words][HS_MAX_PEERS flags][timeout flag, block counter]
  void *peer_mapped[HS_MAX_PEERS] = {};
  bool peer_armed = false;
  uint32_t peer_epoch = 0;
  // latency path: mapped pinned staging (inputs, verdict flags, completion word) + device block counter
  small_rec *h_small_in = nullptr;
  uint8_t *h_small_out = nullptr;
  uint32_t *h_small_done = nullptr;
  uint32_t *d_small_counter = nullptr;
  uint32_t small_seq = 0;
  bool small_enabled = true;
  // deferred-results mode (hs_set_deferred): the latency-bound tail of a `_dev` verify pass (finish kernel, peer exchange, per-QC AND)
  // runs on an internal stream so that it overlaps the main kernel of the NEXT pass; two scratch sets alternate
  bool deferred = false;
  cudaStream_t stream_tail = nullptr;
  cudaEvent_t ev_main_done = nullptr, ev_tail[2] = {nullptr, nullptr}, ev_results = nullptr;
  dev_buf xyz2, meta2;
  int flip = 0;
  // optional timing of the dominant kernel alone (bench.py's roofline): events around k_verify_main<committee>
  bool profile_main = false;
  cudaEvent_t ev_prof[2] = {nullptr, nullptr};
  std::atomic<uint64_t> launches{0};
  std::mutex mu;
  std::mutex err_mu;                  // guards err only: fail() is also reached from argument checks taken before `mu`
  std::string err = "ok";
};

static int fail(hs_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
  if (c) {
    std::lock_guard<std::mutex> g(c->err_mu);
    c->err = what;
    if (e != cudaSuccess) {
      c->err += ": ";
      c->err += cudaGetErrorString(e);
    }
  }
  return code;
}
#define HS_CUDA(c, call)                                               \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return fail((c), HS_ERR_CUDA, #call, e__); \
  } while (0)
#define HS_TRY(expr)          \
  do {                        \
    int rc__ = (expr);        \
    if (rc__) return rc__;    \
  } while (0)

static int ensure(hs_ctx *c, dev_buf &b, size_t need) {
  if (need <= b.cap) return HS_OK;
  if (b.p) {
    cudaDeviceSynchronize();  // the old block may still be in flight on another stream
    cudaFree(b.p);
  }
  b.p = nullptr;
  b.cap = 0;
  size_t want = need + need / 4 + 4096;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) return fail(c, HS_ERR_NOMEM, "cudaMalloc scratch", e);
  b.cap = want;
  return HS_OK;
}
static inline unsigned blocks_for(size_t n, unsigned per = HS_THREADS) { return (unsigned)((n + per - 1) / per); }


static int launch_build(hs_ctx *c, const uint8_t *d_encs, size_t n_points, int negate, int W, int n_windows, ge_niels *tables, uint8_t *flags) {
  size_t threads = n_points * (size_t)n_windows * ((1u << (W - 1)) / HS_BUILD_BLOCK);
  k_build_comb<<<blocks_for(threads), HS_THREADS, 0, c->stream>>>(d_encs, n_points, negate, W, n_windows, tables, flags);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}

static void set_window(comb_params &cp, bool a, int w) {
  if (a) {
    cp.wa = w;
    cp.na = sc_ndigits_rt(w);
    sc_bias_rt(cp.bias_a, w);
  } else {
    cp.wb = w;
    cp.nb = sc_ndigits_rt(w);
    sc_bias_rt(cp.bias_b, w);
  }
}

// ---- key cache -----------------------------------------------------------------------------------------------------
static void cache_release(hs_ctx *c) {
  cudaFree(c->d_pks);
  cudaFree(c->d_key_flags);
  cudaFree(c->d_atables);
  cudaFree(c->d_slots);
  c->d_pks = nullptr;
  c->d_key_flags = nullptr;
  c->d_atables = nullptr;
  c->d_slots = nullptr;
  c->n_keys = 0;
  c->h_pks.clear();
  c->h_slots.clear();
  c->h_key_live.clear();
  c->key_capacity = 0;
}
// lazily allocate the store for cache_cap learned keys (14-bit windows: 14 MB per key, narrower if memory is short)
static int cache_allocate(hs_ctx *c) {
  size_t free_b = 0, total_b = 0;
  HS_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  size_t lim = free_b / 2;
  if (c->table_budget && c->table_budget < lim) lim = c->table_budget;
  int wa = 8;
  for (int w : {14, 12, 10, 8}) {
    if (c->wa_forced && w != c->wa_forced && w != 8) continue;
    wa = w;
    if (c->cache_cap * comb_table_entries(w) * sizeof(ge_niels) <= lim) break;
  }
  if (c->cache_cap * comb_table_entries(wa) * sizeof(ge_niels) > lim || sc_ndigits_rt(wa) + c->cp.nb > HS_MAX_DIGITS) {
    c->cache_enabled = false;  // not enough memory: stay on the generic path
    return HS_OK;
  }
  uint32_t cap = 16;
  while (cap < 2 * c->cache_cap) cap <<= 1;
  set_window(c->cp, true, wa);
  c->a_table_entries = comb_table_entries(wa);
  HS_CUDA(c, cudaMalloc(&c->d_pks, c->cache_cap * 32));
  HS_CUDA(c, cudaMalloc(&c->d_key_flags, c->cache_cap));
  HS_CUDA(c, cudaMalloc(&c->d_slots, (size_t)cap * 4));
  HS_CUDA(c, cudaMalloc(&c->d_atables, c->cache_cap * sizeof(ge_niels) * c->a_table_entries));
  HS_CUDA(c, cudaMemset(c->d_slots, 0xff, (size_t)cap * 4));
  c->slot_mask = cap - 1;
  c->h_slots.assign(cap, HS_NO_KEY);
  c->h_pks.clear();
  return HS_OK;
}
// Called at the start of a verify pass: if the previous pass left unknown keys behind (already copied to pinned host memory),
// dedupe them, append the new ones to the store and build their comb tables on `stream` before this pass's lookup runs.
static int learn_process(hs_ctx *c, cudaStream_t stream) {
  if (!c->learn_pending) return HS_OK;
  if (cudaEventQuery(c->ev_learn) != cudaSuccess) return HS_OK;  // copy still in flight: try again on the next call
  c->learn_pending = false;
  c->calls_since_reset++;
  if (!c->cache_enabled || c->explicit_committee) return HS_OK;
  if (c->cache_full) {
    // Full cache that no longer matches the traffic (e.g. the validator set rotated): more than half of the last pass missed.
    // Start over — the next passes relearn the keys that are actually in use.  (No per-key eviction)
    if (c->learn_records >= 64 && (size_t)*c->h_miss_total * 2 > c->learn_records && c->calls_since_reset >= HS_CACHE_RESET_MIN_CALLS) {
      c->calls_since_reset = 0;
      c->n_keys = 0;
      c->h_pks.clear();
      std::fill(c->h_slots.begin(), c->h_slots.end(), HS_NO_KEY);
      HS_CUDA(c, cudaMemsetAsync(c->d_slots, 0xff, c->h_slots.size() * 4, stream));
      c->cache_full = false;
    }
    return HS_OK;
  }
  const uint32_t got = *c->h_learn_n < HS_LEARN_MAX ? *c->h_learn_n : HS_LEARN_MAX;
  if (got == 0) return HS_OK;
  if (!c->d_atables) {
    HS_TRY(cache_allocate(c));
    if (!c->cache_enabled) return HS_OK;
  }
  const size_t old_n = c->n_keys;
  size_t n_new = 0;
  const uint32_t mask = c->slot_mask;
  for (uint32_t t = 0; t < got && old_n + n_new < c->cache_cap && n_new < HS_LEARN_PER_CALL; t++) {
    const uint8_t *key = c->h_learn_keys + 32 * (size_t)t;
    uint32_t w[8];
    memcpy(w, key, 32);
    uint32_t h = key_hash(w) & mask;
    bool present = false;
    while (c->h_slots[h] != HS_NO_KEY) {
      if (memcmp(c->h_pks.data() + 32 * (size_t)c->h_slots[h], key, 32) == 0) {
        present = true;
        break;
      }
      h = (h + 1) & mask;
    }
    if (present) continue;
    c->h_slots[h] = (uint32_t)(old_n + n_new);
    c->h_pks.insert(c->h_pks.end(), key, key + 32);
    n_new++;
  }
  if (n_new == 0) return HS_OK;
  HS_CUDA(c, cudaMemcpyAsync(c->d_pks + old_n * 32, c->h_pks.data() + old_n * 32, n_new * 32, cudaMemcpyHostToDevice, stream));
  HS_CUDA(c, cudaMemcpyAsync(c->d_slots, c->h_slots.data(), c->h_slots.size() * 4, cudaMemcpyHostToDevice, stream));
  size_t threads = n_new * (size_t)c->cp.na * ((1u << (c->cp.wa - 1)) / HS_BUILD_BLOCK);
  k_build_comb<<<blocks_for(threads), HS_THREADS, 0, stream>>>(c->d_pks + old_n * 32, n_new, 1, c->cp.wa, c->cp.na,
                                                                c->d_atables + old_n * c->a_table_entries, c->d_key_flags + old_n);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  // (no synchronisation: copies from pageable memory return once the source is staged, so the host vectors may change afterwards)
  HS_CUDA(c, cudaEventRecord(c->ev_tables, stream));  // passes on OTHER streams (host entry points vs a _dev caller's stream) wait for the build
  c->n_keys = old_n + n_new;
  if (c->n_keys >= c->cache_cap) c->cache_full = true;  // no free slot: unknown keys stay on the generic path until a reset
  return HS_OK;
}
// After the lookup of a pass: park the unknown keys for learn_process().
static int learn_collect(hs_ctx *c, const in_layout &L, size_t n, bool have_lookup, cudaStream_t stream) {
  if (!c->cache_enabled || c->explicit_committee || c->learn_pending || !L.pk) return HS_OK;
  if (!c->d_learn_keys) {
    HS_CUDA(c, cudaMalloc(&c->d_learn_keys, (size_t)HS_LEARN_MAX * 32));
    HS_CUDA(c, cudaMalloc(&c->d_learn_n, 4));
    HS_CUDA(c, cudaMallocHost(&c->h_learn_keys, (size_t)HS_LEARN_MAX * 32));
    HS_CUDA(c, cudaMallocHost(&c->h_learn_n, 4));
    HS_CUDA(c, cudaMallocHost(&c->h_miss_total, 4));
    HS_CUDA(c, cudaEventCreateWithFlags(&c->ev_learn, cudaEventDisableTiming));
    HS_CUDA(c, cudaEventCreateWithFlags(&c->ev_tables, cudaEventDisableTiming));
  }
  c->learn_records = n;
  if (c->cache_full) {  // only watch the miss rate
    if (!have_lookup) return HS_OK;
    HS_CUDA(c, cudaMemcpyAsync(c->h_miss_total, c->d_miss_count, 4, cudaMemcpyDeviceToHost, stream));
    HS_CUDA(c, cudaEventRecord(c->ev_learn, stream));
    c->learn_pending = true;
    return HS_OK;
  }
  k_gather_keys<<<blocks_for(HS_LEARN_MAX, 256), 256, 0, stream>>>(L, n, have_lookup ? (const uint32_t *)c->miss.p : nullptr,
                                                                     have_lookup ? c->d_miss_count : nullptr, HS_LEARN_MAX, c->d_learn_keys, c->d_learn_n);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  HS_CUDA(c, cudaMemcpyAsync(c->h_learn_n, c->d_learn_n, 4, cudaMemcpyDeviceToHost, stream));
  HS_CUDA(c, cudaMemcpyAsync(c->h_learn_keys, c->d_learn_keys, (size_t)HS_LEARN_MAX * 32, cudaMemcpyDeviceToHost, stream));
  HS_CUDA(c, cudaEventRecord(c->ev_learn, stream));
  c->learn_pending = true;
  return HS_OK;
}

// Runs lookup (optional) -> main (committee and/or generic) -> finish on `stream` for a device-resident layout.
// use_lookup: L.pk is valid and a committee is registered -> resolve indices on the device.
static int run_verify(hs_ctx *c, in_layout L, size_t n, uint32_t mode, uint32_t *d_bitmap, cudaStream_t stream, bool indexed,
                      uint8_t *d_flags_out = nullptr) {
  if (n == 0) {
    if (c->peer_armed) {  // an empty shard still owes its peers the epoch flag
      c->peer_armed = false;
      k_peer_sync_only<<<1, HS_MAX_PEERS, 0, stream>>>(c->peers);
      c->launches++;
      HS_CUDA(c, cudaGetLastError());
    }
    return HS_OK;
  }
  if (indexed && (!c->explicit_committee || c->n_keys == 0)) return fail(c, HS_ERR_ARG, "committee-indexed verify without a registered committee");
  if (!indexed) HS_TRY(learn_process(c, stream));
  if (c->ev_tables && !c->explicit_committee) HS_CUDA(c, cudaStreamWaitEvent(stream, c->ev_tables, 0));
  const bool defer = c->deferred && stream != c->stream;  // host-pointer entry points (internal stream) always complete in stream order
  dev_buf &XYZ = (defer && c->flip) ? c->xyz2 : c->xyz, &META = (defer && c->flip) ? c->meta2 : c->meta;
  const int set = defer ? c->flip : 0;
  if (defer) {
    c->flip ^= 1;
    HS_CUDA(c, cudaStreamWaitEvent(stream, c->ev_tail[set], 0));  // the finish kernel that last read this scratch set is done
  }
  HS_TRY(ensure(c, XYZ, n * 3 * sizeof(fe)));
  HS_TRY(ensure(c, META, n));
  main_out O{(fe *)XYZ.p, (uint8_t *)META.p, 0};
  committee_tables C{c->d_pks, c->d_key_flags, (uint32_t)c->n_keys, c->d_atables, c->a_table_entries};
  const bool committee = c->n_keys > 0 && (indexed || L.pk);
  if (!committee && !L.pk) return fail(c, HS_ERR_ARG, "verify without keys");
  if (committee) {
    if (!indexed) {
      HS_TRY(ensure(c, c->vidx, n * 4));
      HS_TRY(ensure(c, c->miss, n * 4));
      key_table T{c->d_slots, c->slot_mask, c->d_pks, (uint32_t)c->n_keys};
      HS_CUDA(c, cudaMemsetAsync(c->d_miss_count, 0, 4, stream));
      k_key_lookup<<<blocks_for(n, 256), 256, 0, stream>>>(L, n, T, (uint32_t *)c->vidx.p, (uint32_t *)c->miss.p, c->d_miss_count);
      c->launches++;
      HS_CUDA(c, cudaGetLastError());
      L.vidx = (const uint32_t *)c->vidx.p;
      O.side_pass = 1;
      HS_TRY(learn_collect(c, L, n, true, stream));
      // Records whose key is not registered take the generic path over the compacted list.  One generic verify has a
      // ~0.8 ms single-warp latency, so the pass runs on the high-priority side stream CONCURRENTLY with the committee
      // pass (disjoint outputs); its length stays on the device (no host round trip).
      HS_CUDA(c, cudaEventRecord(c->ev_side[0], stream));
      HS_CUDA(c, cudaStreamWaitEvent(c->stream_side, c->ev_side[0], 0));
      unsigned grid = blocks_for(n, 32);
      if (grid > 148u * 8u) grid = 148u * 8u;
      k_verify_main<false><<<grid, 32, 0, c->stream_side>>>(L, 0, c->d_miss_count, (const uint32_t *)c->miss.p, c->d_btable, C, O, c->cp);
      c->launches++;
      HS_CUDA(c, cudaGetLastError());
      HS_CUDA(c, cudaEventRecord(c->ev_side[1], c->stream_side));
    }
    if (c->profile_main) HS_CUDA(c, cudaEventRecord(c->ev_prof[0], stream));
    k_verify_main<true><<<blocks_for(n), HS_THREADS, 0, stream>>>(L, n, nullptr, nullptr, c->d_btable, C, O, c->cp);
    if (c->profile_main) HS_CUDA(c, cudaEventRecord(c->ev_prof[1], stream));
    c->launches++;
    HS_CUDA(c, cudaGetLastError());
    if (!indexed) HS_CUDA(c, cudaStreamWaitEvent(stream, c->ev_side[1], 0));
  } else {
    HS_TRY(learn_collect(c, L, n, false, stream));
    k_verify_main<false><<<blocks_for(n), HS_THREADS, 0, stream>>>(L, n, nullptr, nullptr, c->d_btable, C, O, c->cp);
    c->launches++;
    HS_CUDA(c, cudaGetLastError());
  }
  // small batches: small groups, so that enough blocks exist to hide each block's serial inversion; when the tail overlaps the next pass
  // (deferred mode) latency is hidden anyway and the 16-record group costs the fewest inversions
  const int fin_group = (defer || n >= (1u << 19)) ? 16 : (n >= (1u << 18) ? 8 : 4);
  const size_t fin_threads = (n + fin_group - 1) / fin_group;
  peer_route P{};
  if (c->peer_armed) {
    P = c->peers;
    c->peer_armed = false;
  }
  cudaStream_t fin_stream = stream;
  if (defer) {  // tail on the internal stream: the caller's stream is free for the next pass's digest / main kernels
    HS_CUDA(c, cudaEventRecord(c->ev_main_done, stream));
    HS_CUDA(c, cudaStreamWaitEvent(c->stream_tail, c->ev_main_done, 0));
    fin_stream = c->stream_tail;
  }
  k_verify_finish<<<blocks_for(fin_threads), HS_THREADS, 0, fin_stream>>>(L, n, (const fe *)XYZ.p, (const uint8_t *)META.p, mode, d_bitmap, d_flags_out, P, fin_group);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  if (defer) HS_CUDA(c, cudaEventRecord(c->ev_tail[set], c->stream_tail));
  return HS_OK;
}

// ---- latency path (host side)
// key bytes -> table index through the host mirror of the device hash table (registered committee or learned cache)
static uint32_t host_key_lookup(const hs_ctx *c, const uint8_t *key) {
  if (c->n_keys == 0 || c->h_slots.empty()) return HS_NO_KEY;
  uint32_t w[8];
  memcpy(w, key, 32);
  uint32_t h = key_hash(w) & c->slot_mask;
  for (uint32_t probe = 0; probe <= c->slot_mask; probe++) {
    const uint32_t idx = c->h_slots[h];
    if (idx == HS_NO_KEY) return HS_NO_KEY;
    if (idx < c->n_keys && memcmp(c->h_pks.data() + 32 * (size_t)idx, key, 32) == 0) return idx;
    h = (h + 1) & c->slot_mask;
  }
  return HS_NO_KEY;
}
static bool small_eligible(const hs_ctx *c, size_t n) { return c->small_enabled && n >= 1 && n <= HS_SMALL_MAX && c->n_keys > 0 && c->d_atables; }
// c->h_small_in[0 .. n) is filled: one launch, then poll the completion word the last block writes to mapped host memory.
static int run_small(hs_ctx *c, size_t n, uint32_t mode, uint32_t *out_bitmap, uint8_t *out_flags_or_null) {
  small_rec *d_in = nullptr;
  uint8_t *d_out = nullptr;
  uint32_t *d_done = nullptr;
  HS_CUDA(c, cudaHostGetDevicePointer(&d_in, c->h_small_in, 0));
  HS_CUDA(c, cudaHostGetDevicePointer(&d_out, c->h_small_out, 0));
  HS_CUDA(c, cudaHostGetDevicePointer(&d_done, c->h_small_done, 0));
  const uint32_t seq = ++c->small_seq ? c->small_seq : ++c->small_seq;  // never 0
  committee_tables C{c->d_pks, c->d_key_flags, (uint32_t)c->n_keys, c->d_atables, c->a_table_entries};
  if (c->ev_tables && !c->explicit_committee) HS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tables, 0));
  k_verify_small<<<(unsigned)n, 64, 0, c->stream>>>(d_in, (uint32_t)n, c->d_btable, C, c->cp, d_out, c->d_small_counter, d_done, seq);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  volatile uint32_t *done = c->h_small_done;
  bool finished = false;
  for (uint64_t spin = 0; spin < (1ull << 34); spin++) {
    if (*done == seq) {
      finished = true;
      break;
    }
    if ((spin & 0xfff) == 0xfff) {
      cudaError_t q = cudaStreamQuery(c->stream);
      if (q != cudaSuccess && q != cudaErrorNotReady) return fail(c, HS_ERR_CUDA, "k_verify_small", q);
      if (q == cudaSuccess && *done != seq && spin > (1u << 20)) break;  // stream drained without the completion word
    }
  }
  if (!finished) {
    HS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (*done != seq) return fail(c, HS_ERR_CUDA, "k_verify_small did not complete");
  }
  for (size_t w = 0; w < (n + 31) / 32; w++) out_bitmap[w] = 0;
  const uint32_t want = (mode == HS_MODE_STRICT) ? HS_F_STRICT : HS_F_EQ;
  for (size_t i = 0; i < n; i++) {
    const uint8_t fl = ((volatile uint8_t *)c->h_small_out)[i];
    if (out_flags_or_null) out_flags_or_null[i] = fl;
    if (fl & want) out_bitmap[i >> 5] |= 1u << (i & 31);
  }
  return HS_OK;
}

// Digest of n fixed-size messages: staged/coalesced kernel when every message starts 16-byte aligned and has at least one
// full block, the generic per-thread reader otherwise.
static int launch_digest_fixed(hs_ctx *c, const uint8_t *d_msgs, size_t msg_len, size_t n, uint32_t *d_out, cudaStream_t stream) {
  if (msg_len >= 128 && (msg_len & 15) == 0 && (reinterpret_cast<uintptr_t>(d_msgs) & 15) == 0) {
    sha512_kw kw;
    const int pad_is_const = (msg_len & 127) == 0;
    if (pad_is_const) sha512_pad_schedule(kw, msg_len);
    else memset(&kw, 0, sizeof(kw));
    k_digest32_fixed<<<blocks_for(n), HS_THREADS, 0, stream>>>(d_msgs, msg_len, n, d_out, kw, pad_is_const);
  } else {
    k_digest32<<<blocks_for(n), HS_THREADS, 0, stream>>>(d_msgs, nullptr, msg_len, n, d_out);
  }
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}

// off[0] == 0 and off[i] <= off[i+1]: a decreasing offset would make a length wrap to ~2^64 on the device
static bool offsets_ok(const uint64_t *off, size_t n) {
  if (off[0] != 0) return false;
  for (size_t i = 0; i < n; i++)
    if (off[i] > off[i + 1]) return false;
  return true;
}
static in_layout layout_rec128(const void *d_recs) {
  const uint8_t *r = (const uint8_t *)d_recs;
  return in_layout{r, 128, r + 64, 128, nullptr, r + 96, 128, nullptr, nullptr, 32, 1};
}

extern "C" {

int hs_ctx_create(hs_ctx **out, int device, uint32_t flags) {
  if (!out) return HS_ERR_ARG;
  int wb = (int)(flags & 0xffu);
  if (wb == 0) wb = 24;  // 11 windows x 2^23 entries x 96 B = 8.9 GB of HBM for 11 instead of 16+ additions per [S]B
  if (wb < 8 || wb > 26 || (wb % 2)) return HS_ERR_ARG;  // 26 bits: 10 windows, 32 GB — one addition fewer per verify than the 8.9 GB default
  *out = nullptr;
  hs_ctx *c = new (std::nothrow) hs_ctx();
  if (!c) return HS_ERR_NOMEM;
  c->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    e = cudaStreamCreateWithPriority(&c->stream_side, cudaStreamNonBlocking, hi);
  }
  for (int i = 0; i < 2 && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_side[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMalloc(&c->d_miss_count, 4);
  if (e == cudaSuccess) e = cudaMallocHost(&c->h_miss_count, 4);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream_tail, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_main_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_results, cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&c->ev_tail[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaHostAlloc(&c->h_small_in, HS_SMALL_MAX * sizeof(small_rec), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostAlloc(&c->h_small_out, 256, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostAlloc(&c->h_small_done, 64, cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_small_counter, 4);
  if (e == cudaSuccess) e = cudaMemset(c->d_small_counter, 0, 4);
  if (e == cudaSuccess) *c->h_small_done = 0;
  c->small_enabled = !(getenv("HS_SMALL_PATH") && getenv("HS_SMALL_PATH")[0] == '0');
  set_window(c->cp, false, wb);
  set_window(c->cp, true, 12);
  c->wa_forced = (int)((flags >> 8) & 0xffu);
  if (const char *b = getenv("HS_TABLE_BUDGET_MB")) c->table_budget = (size_t)strtoull(b, nullptr, 10) << 20;
  c->cache_wanted = c->cache_enabled = !(flags & HS_FLAG_NO_KEY_CACHE) && !(getenv("HS_KEY_CACHE") && getenv("HS_KEY_CACHE")[0] == '0');
  if (e == cudaSuccess) e = cudaMalloc(&c->d_btable, sizeof(ge_niels) * comb_table_entries(wb));
  if (e == cudaSuccess) {
    if (launch_build(c, nullptr, 1, 0, wb, c->cp.nb, c->d_btable, nullptr) != HS_OK) e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) {
    fprintf(stderr, "hs_ctx_create: CUDA failure: %s\n", cudaGetErrorString(e));
    hs_ctx_destroy(c);
    return HS_ERR_CUDA;
  }
  *out = c;
  return HS_OK;
}

void hs_ctx_destroy(hs_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  cudaFree(c->d_btable);
  cudaFree(c->d_pks);
  cudaFree(c->d_key_flags);
  cudaFree(c->d_atables);
  cudaFree(c->d_slots);
  cudaFree(c->d_miss_count);
  if (c->h_miss_count) cudaFreeHost(c->h_miss_count);
  cudaFree(c->xyz2.p);
  cudaFree(c->meta2.p);
  for (cudaEvent_t ev : {c->ev_main_done, c->ev_tail[0], c->ev_tail[1], c->ev_results})
    if (ev) cudaEventDestroy(ev);
  if (c->stream_tail) cudaStreamDestroy(c->stream_tail);
  if (c->h_small_in) cudaFreeHost(c->h_small_in);
  if (c->h_small_out) cudaFreeHost(c->h_small_out);
  if (c->h_small_done) cudaFreeHost(c->h_small_done);
  cudaFree(c->d_small_counter);
  for (dev_buf *b : {&c->in[0], &c->in[1], &c->digest[0], &c->digest[1], &c->xyz, &c->meta, &c->vidx, &c->miss, &c->out}) cudaFree(b->p);
  cudaFree(c->d_learn_keys);
  cudaFree(c->d_learn_n);
  if (c->h_learn_keys) cudaFreeHost(c->h_learn_keys);
  if (c->h_learn_n) cudaFreeHost(c->h_learn_n);
  if (c->h_miss_total) cudaFreeHost(c->h_miss_total);
  if (c->ev_learn) cudaEventDestroy(c->ev_learn);
  if (c->ev_tables) cudaEventDestroy(c->ev_tables);
  for (int i = 0; i < 2; i++)
    if (c->ev_prof[i]) cudaEventDestroy(c->ev_prof[i]);
  for (int p = 0; p < HS_MAX_PEERS; p++)
    if (c->peer_mapped[p]) cudaIpcCloseMemHandle(c->peer_mapped[p]);
  cudaFree(c->peer_own);
  for (int i = 0; i < 2; i++) {
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
    if (c->ev_side[i]) cudaEventDestroy(c->ev_side[i]);
  }
  if (c->stream_side) cudaStreamDestroy(c->stream_side);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  delete c;
}

const char *hs_last_error(const hs_ctx *c) { return c ? c->err.c_str() : "null context"; }
size_t hs_cached_keys(const hs_ctx *c) { return (c && !c->explicit_committee) ? c->n_keys : 0; }
void hs_window_bits(const hs_ctx *c, int *key_bits, int *base_bits) {
  if (key_bits) *key_bits = (c && c->n_keys) ? c->cp.wa : 0;
  if (base_bits) *base_bits = c ? c->cp.wb : 0;
}
uint64_t hs_kernel_launches(const hs_ctx *c) { return c ? c->launches.load() : 0; }
/* Measurement hook: with profiling on, CUDA events bracket the k_verify_main<committee> launch of every verify pass on the stream the
 * pass runs on; hs_profile_main_ms() waits for the last pass and returns that kernel's duration in ms (< 0: nothing recorded). */
/* Deferred-results mode for streams of `_dev` verify passes: the tail of a pass (finish kernel + peer exchange, hs_qc_and_dev) runs on an
 * internal stream and overlaps the NEXT pass's kernels; verdict bitmaps are complete only after hs_results_wait(ctx, stream) (which makes
 * `stream` wait for every tail enqueued so far).  Inputs of a pass (signatures) must stay valid until then.  Host-pointer entry points are
 * unaffected: switch the mode only while no pass is in flight. */
int hs_set_deferred(hs_ctx *c, int on) {
  if (!c) return HS_ERR_ARG;
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  HS_CUDA(c, cudaDeviceSynchronize());
  c->deferred = on != 0;
  c->flip = 0;
  return HS_OK;
}
int hs_results_wait(hs_ctx *c, void *stream) {
  if (!c) return HS_ERR_ARG;
  HS_CUDA(c, cudaSetDevice(c->device));
  HS_CUDA(c, cudaEventRecord(c->ev_results, c->stream_tail));
  HS_CUDA(c, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_results, 0));
  return HS_OK;
}
int hs_profile_enable(hs_ctx *c, int on) {
  if (!c) return HS_ERR_ARG;
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  for (int i = 0; i < 2 && on; i++)
    if (!c->ev_prof[i]) HS_CUDA(c, cudaEventCreate(&c->ev_prof[i]));
  c->profile_main = on != 0;
  return HS_OK;
}
double hs_profile_main_ms(hs_ctx *c) {
  if (!c || !c->ev_prof[1]) return -1.0;
  cudaSetDevice(c->device);
  if (cudaEventSynchronize(c->ev_prof[1]) != cudaSuccess) return -1.0;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, c->ev_prof[0], c->ev_prof[1]) != cudaSuccess) return -1.0;
  return (double)ms;
}

void *hs_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void hs_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- committee registration
static int committee_register_locked(hs_ctx *c, const uint8_t *pks, size_t N, uint32_t *out_valid_bitmap) {
  HS_CUDA(c, cudaSetDevice(c->device));
  HS_CUDA(c, cudaDeviceSynchronize());
  cache_release(c);
  c->learn_pending = false;
  c->cache_full = false;
  c->explicit_committee = false;   // set only once the new tables are complete: a failed registration leaves NO committee
  c->cache_enabled = c->cache_wanted;
  if (N == 0) return HS_OK;        // clears the committee and hands key handling back to the cache (if enabled)
  // host-side hash table (hashing only; first occurrence of a duplicated key wins)
  uint32_t cap = 16;
  std::vector<uint32_t> slots;
  // widest per-key window whose tables fit in the budget: ~62 % of the device by default (B200: 16 bits up to ~2.2 k keys,
  // 15 up to ~4.2 k, 14 up to ~7.6 k, 12 up to ~26 k), or HS_TABLE_BUDGET_MB / hs_set_table_budget for a shared device
  size_t free_b = 0, total_b = 0;
  HS_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  size_t budget = total_b / 100 * 62;
  if (c->table_budget) budget = c->table_budget;
  if (budget > free_b - free_b / 8) budget = free_b - free_b / 8;
  // spare slots (1/16 of the set, at least 16) let hs_committee_update add validators without rebuilding anything
  const size_t capk = N + (N / 16 > 16 ? N / 16 : 16);
  while (cap < 2 * capk) cap <<= 1;
  slots.clear();
  slots.assign(cap, HS_NO_KEY);
  for (size_t i = 0; i < N; i++) {
    uint32_t w[8];
    memcpy(w, pks + 32 * i, 32);
    uint32_t h = key_hash(w) & (cap - 1);
    bool dup = false;
    while (slots[h] != HS_NO_KEY) {
      if (memcmp(pks + 32 * (size_t)slots[h], pks + 32 * i, 32) == 0) {
        dup = true;
        break;
      }
      h = (h + 1) & (cap - 1);
    }
    if (!dup) slots[h] = (uint32_t)i;
  }
  int wa = 8;
  for (int w : {17, 16, 15, 14, 13, 12, 11, 10, 9, 8}) {  // 17 bits: 15 windows (94 MB per key: committees up to ~1,200 keys); 18 would still need 15
    if (c->wa_forced && w != c->wa_forced) continue;
    wa = w;
    if (capk * comb_table_entries(w) * sizeof(ge_niels) <= budget) break;
  }
  if (sc_ndigits_rt(wa) + c->cp.nb > HS_MAX_DIGITS) return fail(c, HS_ERR_ARG, "window combination exceeds HS_MAX_DIGITS");
  cudaError_t e = cudaMalloc(&c->d_pks, capk * 32);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_key_flags, capk);
  if (e == cudaSuccess) e = cudaMemset(c->d_key_flags, 0, capk);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_slots, (size_t)cap * 4);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_atables, capk * sizeof(ge_niels) * comb_table_entries(wa));
  if (e != cudaSuccess) {
    cudaGetLastError();
    cache_release(c);
    return fail(c, HS_ERR_NOMEM, "committee tables do not fit in device memory", e);
  }
  set_window(c->cp, true, wa);
  c->a_table_entries = comb_table_entries(wa);
  int rc = HS_OK;
  e = cudaMemcpyAsync(c->d_pks, pks, N * 32, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_slots, slots.data(), (size_t)cap * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) rc = launch_build(c, c->d_pks, N, 1, wa, c->cp.na, c->d_atables, c->d_key_flags);
  if (e == cudaSuccess && rc == HS_OK) e = cudaStreamSynchronize(c->stream);
  std::vector<uint8_t> fl(N);
  if (e == cudaSuccess && rc == HS_OK) e = cudaMemcpy(fl.data(), c->d_key_flags, N, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess || rc != HS_OK) {
    cache_release(c);
    return rc != HS_OK ? rc : fail(c, HS_ERR_CUDA, "committee registration", e);
  }
  c->slot_mask = cap - 1;
  c->n_keys = N;
  c->key_capacity = capk;
  c->explicit_committee = true;
  c->h_pks.assign(pks, pks + N * 32);  // host mirror: hs_committee_update edits the set incrementally
  c->h_slots = slots;
  c->h_key_live.assign(N, 1);
  if (out_valid_bitmap) {
    for (size_t w = 0; w < (N + 31) / 32; w++) out_valid_bitmap[w] = 0;
    for (size_t i = 0; i < N; i++)
      if (fl[i] & 1) out_valid_bitmap[i >> 5] |= 1u << (i & 31);
  }
  return HS_OK;
}
int hs_committee_register(hs_ctx *c, const uint8_t *pks, size_t N, uint32_t *out_valid_bitmap) {
  if (!c || (N && !pks) || N >= HS_NO_KEY) return fail(c, HS_ERR_ARG, "hs_committee_register: bad argument");
  std::lock_guard<std::mutex> g(c->mu);
  return committee_register_locked(c, pks, N, out_valid_bitmap);
}

// Incremental epoch change (consensus/src/config.rs Committee: a few validators join / leave): removed indices stop
// verifying (flag cleared, hash slot dropped), added keys take a free slot — a removed one or a spare — and only THEIR tables
// are built (~0.25 ms per key); every other validator keeps its index and its table.
int hs_committee_update(hs_ctx *c, const uint8_t *add_pks, size_t n_add, const uint32_t *remove_idx, size_t n_remove, uint32_t *out_add_idx) {
  if (!c || (n_add && (!add_pks || !out_add_idx)) || (n_remove && !remove_idx)) return fail(c, HS_ERR_ARG, "hs_committee_update: bad argument");
  std::lock_guard<std::mutex> g(c->mu);
  if (!c->explicit_committee) return fail(c, HS_ERR_ARG, "hs_committee_update: no committee registered");
  HS_CUDA(c, cudaSetDevice(c->device));
  HS_CUDA(c, cudaDeviceSynchronize());  // epoch boundary: nothing of the old set may be in flight
  for (size_t i = 0; i < n_remove; i++)
    if (remove_idx[i] >= c->n_keys) return fail(c, HS_ERR_ARG, "hs_committee_update: remove index out of range");
  for (size_t i = 0; i < n_remove; i++) {
    c->h_key_live[remove_idx[i]] = 0;
    HS_CUDA(c, cudaMemsetAsync(c->d_key_flags + remove_idx[i], 0, 1, c->stream));
  }
  auto find = [&](const uint8_t *key) -> uint32_t {
    uint32_t w[8];
    memcpy(w, key, 32);
    uint32_t h = key_hash(w) & c->slot_mask;
    while (c->h_slots[h] != HS_NO_KEY) {
      const uint32_t idx = c->h_slots[h];
      if (c->h_key_live[idx] && memcmp(c->h_pks.data() + 32 * (size_t)idx, key, 32) == 0) return idx;
      h = (h + 1) & c->slot_mask;
    }
    return HS_NO_KEY;
  };
  size_t next_free = 0;
  std::vector<uint32_t> added;
  for (size_t i = 0; i < n_add; i++) {
    const uint8_t *key = add_pks + 32 * i;
    uint32_t idx = find(key);
    for (uint32_t a : added)  // the same key twice in one call
      if (idx == HS_NO_KEY && memcmp(c->h_pks.data() + 32 * (size_t)a, key, 32) == 0) idx = a;
    if (idx == HS_NO_KEY) {
      while (next_free < c->n_keys && c->h_key_live[next_free]) next_free++;
      if (next_free < c->n_keys) idx = (uint32_t)next_free;
      else if (c->n_keys < c->key_capacity) {
        idx = (uint32_t)c->n_keys++;
        c->h_key_live.push_back(0);
        c->h_pks.resize(c->n_keys * 32);
      } else return fail(c, HS_ERR_NOMEM, "hs_committee_update: no free table slot (re-register the committee)");
      memcpy(c->h_pks.data() + 32 * (size_t)idx, key, 32);
      c->h_key_live[idx] = 1;
      HS_CUDA(c, cudaMemcpyAsync(c->d_pks + 32 * (size_t)idx, key, 32, cudaMemcpyHostToDevice, c->stream));
      HS_TRY(launch_build(c, c->d_pks + 32 * (size_t)idx, 1, 1, c->cp.wa, c->cp.na, c->d_atables + (size_t)idx * c->a_table_entries, c->d_key_flags + idx));
      added.push_back(idx);
    }
    out_add_idx[i] = idx;
  }
  // rebuild the open-addressing table from the live keys (deletions leave no tombstones) and publish it
  std::fill(c->h_slots.begin(), c->h_slots.end(), HS_NO_KEY);
  for (size_t i = 0; i < c->n_keys; i++) {
    if (!c->h_key_live[i]) continue;
    uint32_t w[8];
    memcpy(w, c->h_pks.data() + 32 * i, 32);
    uint32_t h = key_hash(w) & c->slot_mask;
    bool dup = false;
    while (c->h_slots[h] != HS_NO_KEY) {
      if (memcmp(c->h_pks.data() + 32 * (size_t)c->h_slots[h], c->h_pks.data() + 32 * i, 32) == 0) {
        dup = true;
        break;
      }
      h = (h + 1) & c->slot_mask;
    }
    if (!dup) c->h_slots[h] = (uint32_t)i;
  }
  HS_CUDA(c, cudaMemcpyAsync(c->d_slots, c->h_slots.data(), c->h_slots.size() * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}
/* Upper bound (bytes) for the per-key tables of the NEXT registration / key-cache allocation; 0 = default (~62 % of the device). */
int hs_set_table_budget(hs_ctx *c, size_t bytes) {
  if (!c) return HS_ERR_ARG;
  std::lock_guard<std::mutex> g(c->mu);
  c->table_budget = bytes;
  return HS_OK;
}

// ---- device-resident entry points (one stream at a time per context: they share the context's scratch)
int hs_verify_rec128_dev(hs_ctx *c, const void *d_recs, size_t n, uint32_t mode, void *d_bitmap, void *stream) {
  if (!c || (n && (!d_recs || !d_bitmap)) || mode > 1) return fail(c, HS_ERR_ARG, "hs_verify_rec128_dev: bad argument");
  HS_CUDA(c, cudaSetDevice(c->device));
  return run_verify(c, layout_rec128(d_recs), n, mode, (uint32_t *)d_bitmap, (cudaStream_t)stream, false);
}
int hs_verify_var_dev(hs_ctx *c, const void *d_sig, const void *d_pk, const void *d_msgs, const void *d_off, size_t n, uint32_t mode,
                      void *d_bitmap, void *stream) {
  if (!c || mode > 1 || (n && (!d_sig || !d_pk || !d_off || !d_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_var_dev: bad argument");
  HS_CUDA(c, cudaSetDevice(c->device));
  in_layout L{(const uint8_t *)d_sig, 64, (const uint8_t *)d_pk, 32, nullptr, (const uint8_t *)d_msgs, 0, nullptr, (const uint64_t *)d_off, 0, 0};
  return run_verify(c, L, n, mode, (uint32_t *)d_bitmap, (cudaStream_t)stream, false);
}
int hs_verify_committee_dev(hs_ctx *c, const void *d_vidx, const void *d_sig, const void *d_midx, const void *d_digests, size_t n,
                            uint32_t mode, void *d_bitmap, void *stream) {
  if (!c || mode > 1 || (n && (!d_vidx || !d_sig || !d_digests || !d_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_committee_dev: bad argument");
  HS_CUDA(c, cudaSetDevice(c->device));
  // d_midx == NULL: every vote is over digests[0]
  in_layout L{(const uint8_t *)d_sig, 64, nullptr, 0, (const uint32_t *)d_vidx, (const uint8_t *)d_digests, d_midx ? (size_t)32 : (size_t)0,
              (const uint32_t *)d_midx, nullptr, 32, 0};
  return run_verify(c, L, n, mode, (uint32_t *)d_bitmap, (cudaStream_t)stream, true);
}
int hs_digest32_dev(hs_ctx *c, const void *d_data, const void *d_off, size_t n, void *d_out, void *stream) {
  if (!c || (n && (!d_off || !d_out))) return fail(c, HS_ERR_ARG, "hs_digest32_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  k_digest32<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t *)d_data, (const uint64_t *)d_off, 0, n, (uint32_t *)d_out);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}
int hs_digest32_fixed_dev(hs_ctx *c, const void *d_msgs, size_t msg_len, size_t n, void *d_out, void *stream) {
  if (!c || (n && (!d_msgs || !d_out || msg_len == 0))) return fail(c, HS_ERR_ARG, "hs_digest32_fixed_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  return launch_digest_fixed(c, (const uint8_t *)d_msgs, msg_len, n, (uint32_t *)d_out, (cudaStream_t)stream);
}
int hs_verify_msgs_dev(hs_ctx *c, const void *d_sig, const void *d_pk, const void *d_vidx, const void *d_msgs, size_t msg_len, size_t n,
                       uint32_t mode, void *d_digests, void *d_bitmap, void *stream) {
  if (!c || mode > 1 || (n && (!d_sig || (!d_pk && !d_vidx) || !d_msgs || !d_digests || !d_bitmap)))
    return fail(c, HS_ERR_ARG, "hs_verify_msgs_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  // Digest(msg_i) in its own kernel: fusing it into k_verify_main was measured SLOWER on B200 (4.53 vs 4.13 ms per 2^20: the
  // SHA phase then runs at the curve kernel's 128-register occupancy and the two phases do not overlap across pipes in practice).
  HS_TRY(launch_digest_fixed(c, (const uint8_t *)d_msgs, msg_len, n, (uint32_t *)d_digests, (cudaStream_t)stream));
  in_layout L{(const uint8_t *)d_sig, 64, (const uint8_t *)d_pk, 32, (const uint32_t *)d_vidx, (const uint8_t *)d_digests, 32, nullptr, nullptr, 32, 0};
  return run_verify(c, L, n, mode, (uint32_t *)d_bitmap, (cudaStream_t)stream, d_vidx != nullptr);
}

// ---- QC::verify for many certificates (consensus/src/messages.rs:180-208)
int hs_verify_qcs(hs_ctx *c, const uint8_t *preimages, size_t n_qc, const uint8_t *pk, const uint32_t *vidx, const uint8_t *sig,
                  const uint32_t *qc_idx, size_t n_votes, uint32_t *out_vote_bitmap, uint32_t *out_qc_bitmap) {
  if (!c || !out_qc_bitmap || (n_qc && !preimages) || (n_votes && (!sig || !qc_idx || (!pk && !vidx) || n_qc == 0)))
    return fail(c, HS_ERR_ARG, "hs_verify_qcs: bad argument");
  const size_t qc_words = (n_qc + 31) / 32, vote_words = (n_votes + 31) / 32;
  for (size_t w = 0; w < qc_words; w++) out_qc_bitmap[w] = (w == qc_words - 1 && (n_qc & 31)) ? ((1u << (n_qc & 31)) - 1u) : 0xffffffffu;
  if (n_votes == 0) return HS_OK;
  for (size_t i = 0; i < n_votes; i++)
    if (qc_idx[i] >= n_qc) return fail(c, HS_ERR_ARG, "hs_verify_qcs: qc_idx out of range");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  const size_t key_bytes = pk ? 32 : 4;
  const size_t o_pre = 0, o_dig = (n_qc * 40 + 15) & ~(size_t)15, o_sig = o_dig + n_qc * 32, o_key = o_sig + n_votes * 64,
               o_qi = o_key + ((n_votes * key_bytes + 15) & ~(size_t)15), total = o_qi + n_votes * 4;
  HS_TRY(ensure(c, c->in[0], total));
  HS_TRY(ensure(c, c->out, (vote_words + qc_words) * 4));
  uint8_t *d = (uint8_t *)c->in[0].p;
  uint32_t *d_votes = (uint32_t *)c->out.p, *d_qc = d_votes + vote_words;
  HS_CUDA(c, cudaMemcpyAsync(d + o_pre, preimages, n_qc * 40, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n_votes * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_key, pk ? (const void *)pk : (const void *)vidx, n_votes * key_bytes, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_qi, qc_idx, n_votes * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d_qc, out_qc_bitmap, qc_words * 4, cudaMemcpyHostToDevice, c->stream));  // all ones (tail cleared)
  // QC::digest = SHA-512(hash || round_le)[..32] for every certificate (messages.rs:201-208)
  k_digest32<<<blocks_for(n_qc), HS_THREADS, 0, c->stream>>>(d + o_pre, nullptr, 40, n_qc, (uint32_t *)(d + o_dig));
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  in_layout L{d + o_sig, 64, pk ? d + o_key : nullptr, 32, pk ? nullptr : (const uint32_t *)(d + o_key), d + o_dig, 32, (const uint32_t *)(d + o_qi),
              nullptr, 32, 0};
  HS_TRY(run_verify(c, L, n_votes, HS_MODE_BATCH_EQ, d_votes, c->stream, pk == nullptr));
  k_qc_and<<<blocks_for(n_votes, 256), 256, 0, c->stream>>>(d_votes, (const uint32_t *)(d + o_qi), n_votes, n_qc, d_qc);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  if (out_vote_bitmap) HS_CUDA(c, cudaMemcpyAsync(out_vote_bitmap, d_votes, vote_words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(out_qc_bitmap, d_qc, qc_words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

// ---- device-resident QC verification (strong-scaling path of BASELINE config[3]: the votes of many QCs sharded over ranks)
// Per-vote verdicts of this rank's shard (batch-eq condition).  d_qc_digests: n_qc x 32 (hs_digest32_fixed_dev over the 40-byte
// preimages); d_qc_idx selects each vote's digest.  Combine with the (all-gathered) vote bitmap through hs_qc_and_dev.
int hs_verify_qc_votes_dev(hs_ctx *c, const void *d_qc_digests, const void *d_pk, const void *d_vidx, const void *d_sig, const void *d_qc_idx,
                           size_t n_votes, void *d_vote_bitmap, void *stream) {
  if (!c || (n_votes && (!d_qc_digests || (!d_pk && !d_vidx) || !d_sig || !d_qc_idx || !d_vote_bitmap)))
    return fail(c, HS_ERR_ARG, "hs_verify_qc_votes_dev: bad argument");
  HS_CUDA(c, cudaSetDevice(c->device));
  in_layout L{(const uint8_t *)d_sig, 64, (const uint8_t *)d_pk, 32, (const uint32_t *)d_vidx, (const uint8_t *)d_qc_digests, 32,
              (const uint32_t *)d_qc_idx, nullptr, 32, 0};
  return run_verify(c, L, n_votes, HS_MODE_BATCH_EQ, (uint32_t *)d_vote_bitmap, (cudaStream_t)stream, d_pk == nullptr);
}
// d_qc_bitmap bit j = AND of the verdict bits of the votes with qc_idx == j (no votes -> 1), over a vote bitmap of n_votes bits.
int hs_qc_and_dev(hs_ctx *c, const void *d_vote_bitmap, const void *d_qc_idx, size_t n_votes, size_t n_qc, void *d_qc_bitmap, void *stream) {
  if (!c || !d_qc_bitmap || (n_votes && (!d_vote_bitmap || !d_qc_idx))) return fail(c, HS_ERR_ARG, "hs_qc_and_dev: bad argument");
  HS_CUDA(c, cudaSetDevice(c->device));
  cudaStream_t st = (c->deferred && (cudaStream_t)stream != c->stream) ? c->stream_tail : (cudaStream_t)stream;  // deferred: after the finish kernel on the tail stream
  if (n_qc) k_bitmap_ones<<<blocks_for((n_qc + 31) / 32, 256), 256, 0, st>>>((uint32_t *)d_qc_bitmap, n_qc);
  if (n_votes)
    k_qc_and<<<blocks_for(n_votes, 256), 256, 0, st>>>((const uint32_t *)d_vote_bitmap, (const uint32_t *)d_qc_idx, n_votes, n_qc, (uint32_t *)d_qc_bitmap);
  c->launches += (n_qc ? 1 : 0) + (n_votes ? 1 : 0);
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}

// ---- TC::verify / Timeout::verify for many certificates (consensus/src/messages.rs:250-265,290-315)
// Vote i = (key_i, sig_i, high_qc_round_i) of certificate tc_idx[i] (NULL: vote i is its own certificate — the Timeout
// shape, rounds[i] = Timeout.round); message = SHA-512(tc_round || high_qc_round)[..32] built on the GPU; Signature::verify
// (strict) per vote; out_tc_bitmap bit j = AND over certificate j.  Stake / duplicate checks (messages.rs:292-304) stay on the host.
int hs_verify_tcs(hs_ctx *c, const uint64_t *tc_rounds, size_t n_tc, const uint8_t *pk, const uint32_t *vidx, const uint8_t *sig,
                  const uint64_t *high_qc_rounds, const uint32_t *tc_idx, size_t n_votes, uint32_t *out_vote_bitmap, uint32_t *out_tc_bitmap) {
  if (!c || !out_tc_bitmap || (n_tc && !tc_rounds) || (n_votes && (!sig || !high_qc_rounds || (!pk && !vidx) || n_tc == 0)) || (!tc_idx && n_votes && n_tc != n_votes))
    return fail(c, HS_ERR_ARG, "hs_verify_tcs: bad argument");
  const size_t tc_words = (n_tc + 31) / 32, vote_words = (n_votes + 31) / 32;
  for (size_t w = 0; w < tc_words; w++) out_tc_bitmap[w] = (w == tc_words - 1 && (n_tc & 31)) ? ((1u << (n_tc & 31)) - 1u) : 0xffffffffu;
  if (n_votes == 0) return HS_OK;
  if (tc_idx)
    for (size_t i = 0; i < n_votes; i++)
      if (tc_idx[i] >= n_tc) return fail(c, HS_ERR_ARG, "hs_verify_tcs: tc_idx out of range");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  const size_t key_bytes = pk ? 32 : 4;
  const size_t o_r = 0, o_hq = (n_tc * 8 + 15) & ~(size_t)15, o_dig = o_hq + ((n_votes * 8 + 15) & ~(size_t)15), o_sig = o_dig + n_votes * 32,
               o_key = o_sig + n_votes * 64, o_ti = o_key + ((n_votes * key_bytes + 15) & ~(size_t)15), total = o_ti + n_votes * 4;
  HS_TRY(ensure(c, c->in[0], total));
  HS_TRY(ensure(c, c->out, (vote_words + tc_words) * 4));
  uint8_t *d = (uint8_t *)c->in[0].p;
  uint32_t *d_votes = (uint32_t *)c->out.p, *d_tc = d_votes + vote_words;
  HS_CUDA(c, cudaMemcpyAsync(d + o_r, tc_rounds, n_tc * 8, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_hq, high_qc_rounds, n_votes * 8, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n_votes * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_key, pk ? (const void *)pk : (const void *)vidx, n_votes * key_bytes, cudaMemcpyHostToDevice, c->stream));
  if (tc_idx) HS_CUDA(c, cudaMemcpyAsync(d + o_ti, tc_idx, n_votes * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d_tc, out_tc_bitmap, tc_words * 4, cudaMemcpyHostToDevice, c->stream));
  k_tc_digests<<<blocks_for(n_votes), HS_THREADS, 0, c->stream>>>((const uint64_t *)(d + o_r), tc_idx ? (const uint32_t *)(d + o_ti) : nullptr,
                                                                  (const uint64_t *)(d + o_hq), n_votes, n_tc, (uint32_t *)(d + o_dig));
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  in_layout L{d + o_sig, 64, pk ? d + o_key : nullptr, 32, pk ? nullptr : (const uint32_t *)(d + o_key), d + o_dig, 32, nullptr, nullptr, 32, 0};
  HS_TRY(run_verify(c, L, n_votes, HS_MODE_STRICT, d_votes, c->stream, pk == nullptr));
  if (tc_idx) {
    k_qc_and<<<blocks_for(n_votes, 256), 256, 0, c->stream>>>(d_votes, (const uint32_t *)(d + o_ti), n_votes, n_tc, d_tc);
    c->launches++;
    HS_CUDA(c, cudaGetLastError());
    HS_CUDA(c, cudaMemcpyAsync(out_tc_bitmap, d_tc, tc_words * 4, cudaMemcpyDeviceToHost, c->stream));
  } else {
    HS_CUDA(c, cudaMemcpyAsync(out_tc_bitmap, d_votes, vote_words * 4, cudaMemcpyDeviceToHost, c->stream));  // one vote per certificate
  }
  if (out_vote_bitmap) HS_CUDA(c, cudaMemcpyAsync(out_vote_bitmap, d_votes, vote_words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

// ---- mixed groups: Block::verify for many blocks (messages.rs:54-76) = author signature (strict) + QC votes (batch-eq) + TC
// votes (strict), all in ONE pass.  Item i signs Digest(preimage[msg_idx[i]]) (variable-length preimages, hashed on the GPU),
// belongs to group group_idx[i] and is judged by mode[i]; out_group_bitmap bit j = AND over group j's items.
int hs_verify_groups(hs_ctx *c, const uint8_t *preimages, const uint64_t *pre_off, size_t n_msgs, const uint8_t *sig, const uint8_t *pk,
                     const uint32_t *vidx, const uint32_t *msg_idx, const uint32_t *group_idx, const uint8_t *mode, size_t n_items, size_t n_groups,
                     uint32_t *out_item_bitmap, uint32_t *out_group_bitmap) {
  if (!c || !out_group_bitmap || (n_msgs && !pre_off) || (n_items && (!sig || !msg_idx || !group_idx || (!pk && !vidx) || n_msgs == 0 || n_groups == 0)))
    return fail(c, HS_ERR_ARG, "hs_verify_groups: bad argument");
  const size_t g_words = (n_groups + 31) / 32, i_words = (n_items + 31) / 32;
  for (size_t w = 0; w < g_words; w++) out_group_bitmap[w] = (w == g_words - 1 && (n_groups & 31)) ? ((1u << (n_groups & 31)) - 1u) : 0xffffffffu;
  if (n_items == 0) return HS_OK;
  if (!offsets_ok(pre_off, n_msgs) || (pre_off[n_msgs] && !preimages)) return fail(c, HS_ERR_ARG, "hs_verify_groups: bad preimage offsets");
  for (size_t i = 0; i < n_items; i++)
    if (msg_idx[i] >= n_msgs || group_idx[i] >= n_groups || (mode && mode[i] > 1)) return fail(c, HS_ERR_ARG, "hs_verify_groups: index out of range");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  const size_t key_bytes = pk ? 32 : 4, pre_bytes = pre_off[n_msgs];
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_off = 0, o_pre = al((n_msgs + 1) * 8), o_dig = o_pre + al(pre_bytes + 8), o_sig = o_dig + n_msgs * 32, o_key = o_sig + n_items * 64,
               o_mi = o_key + al(n_items * key_bytes), o_gi = o_mi + al(n_items * 4), o_mo = o_gi + al(n_items * 4), o_fl = o_mo + al(n_items),
               total = o_fl + al(n_items);
  HS_TRY(ensure(c, c->in[0], total));
  HS_TRY(ensure(c, c->out, (i_words + g_words) * 4));
  uint8_t *d = (uint8_t *)c->in[0].p;
  uint32_t *d_items = (uint32_t *)c->out.p, *d_groups = d_items + i_words;
  HS_CUDA(c, cudaMemcpyAsync(d + o_off, pre_off, (n_msgs + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (pre_bytes) HS_CUDA(c, cudaMemcpyAsync(d + o_pre, preimages, pre_bytes, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n_items * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_key, pk ? (const void *)pk : (const void *)vidx, n_items * key_bytes, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_mi, msg_idx, n_items * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_gi, group_idx, n_items * 4, cudaMemcpyHostToDevice, c->stream));
  if (mode) HS_CUDA(c, cudaMemcpyAsync(d + o_mo, mode, n_items, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d_groups, out_group_bitmap, g_words * 4, cudaMemcpyHostToDevice, c->stream));
  HS_TRY(hs_digest32_dev(c, d + o_pre, d + o_off, n_msgs, d + o_dig, c->stream));
  in_layout L{d + o_sig, 64, pk ? d + o_key : nullptr, 32, pk ? nullptr : (const uint32_t *)(d + o_key), d + o_dig, 32, (const uint32_t *)(d + o_mi),
              nullptr, 32, 0};
  HS_TRY(run_verify(c, L, n_items, HS_MODE_STRICT, d_items, c->stream, pk == nullptr, d + o_fl));
  k_group_and<<<blocks_for(n_items, 256), 256, 0, c->stream>>>(d + o_fl, mode ? d + o_mo : nullptr, (const uint32_t *)(d + o_gi), n_items, n_groups, d_items,
                                                               d_groups);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  if (out_item_bitmap) HS_CUDA(c, cudaMemcpyAsync(out_item_bitmap, d_items, i_words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(out_group_bitmap, d_groups, g_words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

// ---- load generation: RFC 8032 keygen / signing of 32-byte digests on the GPU
int hs_keygen_batch_dev(hs_ctx *c, const void *d_seeds, size_t n, void *d_pks, void *stream) {
  if (!c || (n && (!d_seeds || !d_pks))) return fail(c, HS_ERR_ARG, "hs_keygen_batch_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  k_keygen<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t *)d_seeds, n, c->d_btable, c->cp, (uint8_t *)d_pks);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}
int hs_sign_digests_dev(hs_ctx *c, const void *d_seeds, const void *d_pks, size_t n_keys, const void *d_key_idx, const void *d_digests, size_t n,
                        void *d_sig, void *stream) {
  if (!c || (n && (!d_seeds || !d_pks || !d_digests || !d_sig || n_keys == 0)) || (!d_key_idx && n > n_keys))
    return fail(c, HS_ERR_ARG, "hs_sign_digests_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  k_sign_digests<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t *)d_seeds, (const uint8_t *)d_pks, (const uint32_t *)d_key_idx,
                                                                          (const uint8_t *)d_digests, n, n_keys, c->d_btable, c->cp, (uint8_t *)d_sig);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}
int hs_keygen_batch(hs_ctx *c, const uint8_t *seeds, size_t n, uint8_t *out_pks) {
  if (!c || (n && (!seeds || !out_pks))) return fail(c, HS_ERR_ARG, "hs_keygen_batch: bad argument");
  if (n == 0) return HS_OK;
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  HS_TRY(ensure(c, c->in[0], n * 32));
  HS_TRY(ensure(c, c->out, n * 32));
  HS_CUDA(c, cudaMemcpyAsync(c->in[0].p, seeds, n * 32, cudaMemcpyHostToDevice, c->stream));
  HS_TRY(hs_keygen_batch_dev(c, c->in[0].p, n, c->out.p, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(out_pks, c->out.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}
int hs_sign_digests(hs_ctx *c, const uint8_t *seeds, const uint8_t *pks, size_t n_keys, const uint32_t *key_idx, const uint8_t *digests, size_t n,
                    uint8_t *out_sig) {
  if (!c || (n && (!seeds || !pks || !digests || !out_sig || n_keys == 0)) || (!key_idx && n > n_keys)) return fail(c, HS_ERR_ARG, "hs_sign_digests: bad argument");
  if (n == 0) return HS_OK;
  if (key_idx)
    for (size_t i = 0; i < n; i++)
      if (key_idx[i] >= n_keys) return fail(c, HS_ERR_ARG, "hs_sign_digests: key index out of range");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  const size_t o_seed = 0, o_pk = n_keys * 32, o_ki = o_pk + n_keys * 32, o_d = o_ki + ((n * 4 + 15) & ~(size_t)15), total = o_d + n * 32;
  HS_TRY(ensure(c, c->in[0], total));
  HS_TRY(ensure(c, c->out, n * 64));
  uint8_t *d = (uint8_t *)c->in[0].p;
  HS_CUDA(c, cudaMemcpyAsync(d + o_seed, seeds, n_keys * 32, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_pk, pks, n_keys * 32, cudaMemcpyHostToDevice, c->stream));
  if (key_idx) HS_CUDA(c, cudaMemcpyAsync(d + o_ki, key_idx, n * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_d, digests, n * 32, cudaMemcpyHostToDevice, c->stream));
  HS_TRY(hs_sign_digests_dev(c, d + o_seed, d + o_pk, n_keys, key_idx ? d + o_ki : nullptr, d + o_d, n, c->out.p, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(out_sig, c->out.p, n * 64, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

// ---- multi-GPU peer routing (one process per GPU; handles are exchanged by the host, e.g. torch.distributed.all_gather_object)
int hs_peer_setup(hs_ctx *c, int rank, int world, size_t total_words, uint8_t handle_out[64]) {
  if (!c || world < 1 || world > HS_MAX_PEERS || rank < 0 || rank >= world || !handle_out || total_words % (size_t)world)
    return fail(c, HS_ERR_ARG, "hs_peer_setup: bad argument (total_words must be a multiple of world)");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  if (c->peer_own) {  // a new geometry (another workload): drop the old buffers.  Every rank must have drained its stream first.
    HS_CUDA(c, cudaDeviceSynchronize());
    for (int p = 0; p < HS_MAX_PEERS; p++)
      if (c->peer_mapped[p]) {
        cudaIpcCloseMemHandle(c->peer_mapped[p]);
        c->peer_mapped[p] = nullptr;
      }
    cudaFree(c->peer_own);
    c->peer_own = nullptr;
    c->peer_epoch = 0;
    c->peer_armed = false;
  }
  const size_t bytes = (2 * total_words + HS_MAX_PEERS + 16) * 4;
  HS_CUDA(c, cudaMalloc(&c->peer_own, bytes));
  HS_CUDA(c, cudaMemset(c->peer_own, 0, bytes));
  cudaIpcMemHandle_t h;
  HS_CUDA(c, cudaIpcGetMemHandle(&h, c->peer_own));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle_out, &h, 64);
  c->peer_rank = rank;
  c->peer_total_words = total_words;
  c->peers = peer_route{};
  c->peers.n = world;
  c->peers.my_rank = rank;
  c->peers.total_words = total_words;
  c->peers.buf[rank] = c->peer_own;
  return HS_OK;
}
int hs_peer_open(hs_ctx *c, int peer_rank, const uint8_t handle[64]) {
  if (!c || !handle || peer_rank < 0 || peer_rank >= c->peers.n || peer_rank == c->peer_rank) return fail(c, HS_ERR_ARG, "hs_peer_open: bad argument");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void *p = nullptr;
  HS_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer_mapped[peer_rank] = p;
  c->peers.buf[peer_rank] = (uint32_t *)p;
  return HS_OK;
}
/* Arms the NEXT `_dev` verify call on this context: its bitmap goes to every rank's buffer at word_offset (fused all-gather). */
int hs_peer_next(hs_ctx *c, size_t word_offset, uint32_t epoch) {
  if (!c || !c->peer_own) return fail(c, HS_ERR_ARG, "hs_peer_next: peers not set up");
  std::lock_guard<std::mutex> g(c->mu);
  for (int p = 0; p < c->peers.n; p++)
    if (!c->peers.buf[p]) return fail(c, HS_ERR_ARG, "hs_peer_next: a peer buffer is not mapped");
  if (c->peer_epoch != 0 && epoch != c->peer_epoch + 1) return fail(c, HS_ERR_ARG, "hs_peer_next: epochs must increase by one");
  if (word_offset >= c->peer_total_words && c->peer_total_words) return fail(c, HS_ERR_ARG, "hs_peer_next: word_offset out of range");
  c->peers.word_offset = word_offset;
  c->peers.epoch = epoch;
  c->peer_epoch = epoch;
  c->peer_armed = true;
  return HS_OK;
}
/* Device pointer of this rank's copy of the full bitmap of the most recently armed epoch, and whether a peer wait ever timed out. */
void *hs_peer_bitmap(hs_ctx *c) { return (c && c->peer_own) ? (void *)(c->peer_own + (size_t)(c->peer_epoch & 1u) * c->peer_total_words) : nullptr; }
int hs_peer_timed_out(hs_ctx *c) {
  if (!c || !c->peer_own) return 0;
  uint32_t v = 0;
  cudaSetDevice(c->device);
  cudaMemcpy(&v, c->peer_own + 2 * c->peer_total_words + HS_MAX_PEERS, 4, cudaMemcpyDeviceToHost);
  return (int)v;
}

// ---- host-pointer entry points
static int finish_bitmap(hs_ctx *c, size_t n, uint32_t *out_bitmap) {
  size_t words = (n + 31) / 32;
  HS_CUDA(c, cudaMemcpyAsync(out_bitmap, c->out.p, words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

int hs_verify_rec128(hs_ctx *c, const hs_rec128 *recs, size_t n, uint32_t mode, uint32_t *out_bitmap) {
  if (!c || mode > 1 || (n && (!recs || !out_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_rec128: bad argument");
  if (n == 0) return HS_OK;
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  if (small_eligible(c, n)) {  // latency path: every key must already have a table (registered or learned)
    bool all = true;
    for (size_t i = 0; i < n && all; i++) {
      const uint32_t idx = host_key_lookup(c, recs[i].pk);
      if (idx == HS_NO_KEY) all = false;
      else {
        memcpy(c->h_small_in[i].sig, recs[i].sig, 64);
        memcpy(c->h_small_in[i].msg, recs[i].msg, 32);
        c->h_small_in[i].vidx = idx;
      }
    }
    if (all) return run_small(c, n, mode, out_bitmap, nullptr);
  }
  HS_TRY(ensure(c, c->in[0], n * sizeof(hs_rec128)));
  HS_TRY(ensure(c, c->out, ((n + 31) / 32) * 4));
  HS_CUDA(c, cudaMemcpyAsync(c->in[0].p, recs, n * sizeof(hs_rec128), cudaMemcpyHostToDevice, c->stream));
  HS_TRY(run_verify(c, layout_rec128(c->in[0].p), n, mode, (uint32_t *)c->out.p, c->stream, false));
  return finish_bitmap(c, n, out_bitmap);
}
int hs_verify_strict_batch(hs_ctx *c, const hs_rec128 *recs, size_t n, uint32_t *out_bitmap) {
  return hs_verify_rec128(c, recs, n, HS_MODE_STRICT, out_bitmap);
}

int hs_verify_var(hs_ctx *c, const uint8_t *sig, const uint8_t *pk, const uint8_t *msgs, const uint64_t *off, size_t n, uint32_t mode,
                  uint32_t *out_bitmap) {
  if (!c || mode > 1 || (n && (!sig || !pk || !off || !out_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_var: bad argument");
  if (n == 0) return HS_OK;
  if (!offsets_ok(off, n)) return fail(c, HS_ERR_ARG, "hs_verify_var: offsets must start at 0 and be non-decreasing");
  if (off[n] && !msgs) return fail(c, HS_ERR_ARG, "hs_verify_var: null msgs");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  // device layout: [sig n*64][pk n*32][off (n+1)*8][msgs] — every section 8-byte aligned
  size_t o_sig = 0, o_pk = n * 64, o_off = o_pk + n * 32, o_msg = o_off + (n + 1) * 8, total = o_msg + off[n];
  HS_TRY(ensure(c, c->in[0], total + 8));
  HS_TRY(ensure(c, c->out, ((n + 31) / 32) * 4));
  uint8_t *d = (uint8_t *)c->in[0].p;
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_pk, pk, n * 32, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_off, off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (off[n]) HS_CUDA(c, cudaMemcpyAsync(d + o_msg, msgs, off[n], cudaMemcpyHostToDevice, c->stream));
  in_layout L{d + o_sig, 64, d + o_pk, 32, nullptr, d + o_msg, 0, nullptr, (const uint64_t *)(d + o_off), 0, 0};
  HS_TRY(run_verify(c, L, n, mode, (uint32_t *)c->out.p, c->stream, false));
  return finish_bitmap(c, n, out_bitmap);
}

int hs_verify_batch_shared_msg(hs_ctx *c, const uint8_t digest[32], const hs_vote *votes, size_t n, int *all_ok, uint32_t *out_bitmap_or_null) {
  if (!c || !all_ok || !digest || (n && !votes)) return fail(c, HS_ERR_ARG, "hs_verify_batch_shared_msg: bad argument");
  *all_ok = 0;
  if (n == 0) {  // dalek::verify_batch on empty input is Ok
    *all_ok = 1;
    return HS_OK;
  }
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  size_t words = (n + 31) / 32;
  if (small_eligible(c, n)) {  // latency path (the 2f+1 = 3 votes of a 4-node QC)
    bool all = true;
    for (size_t i = 0; i < n && all; i++) {
      const uint32_t idx = host_key_lookup(c, votes[i].pk);
      if (idx == HS_NO_KEY) all = false;
      else {
        memcpy(c->h_small_in[i].sig, votes[i].sig, 64);
        memcpy(c->h_small_in[i].msg, digest, 32);
        c->h_small_in[i].vidx = idx;
      }
    }
    if (all) {
      uint32_t bm2[(HS_SMALL_MAX + 31) / 32];
      HS_TRY(run_small(c, n, HS_MODE_BATCH_EQ, bm2, nullptr));
      int ok2 = 1;
      for (size_t w = 0; w < words; w++) {
        const uint32_t want = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
        if (bm2[w] != want) ok2 = 0;
        if (out_bitmap_or_null) out_bitmap_or_null[w] = bm2[w];
      }
      *all_ok = ok2;
      return HS_OK;
    }
  }
  HS_TRY(ensure(c, c->in[0], n * sizeof(hs_vote) + 32));
  HS_TRY(ensure(c, c->out, words * 4));
  uint8_t *d = (uint8_t *)c->in[0].p;
  HS_CUDA(c, cudaMemcpyAsync(d, digest, 32, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + 32, votes, n * sizeof(hs_vote), cudaMemcpyHostToDevice, c->stream));
  in_layout L{d + 32 + 32, sizeof(hs_vote), d + 32, sizeof(hs_vote), nullptr, d, 0, nullptr, nullptr, 32, 0};
  HS_TRY(run_verify(c, L, n, HS_MODE_BATCH_EQ, (uint32_t *)c->out.p, c->stream, false));
  std::vector<uint32_t> tmp;
  uint32_t *bm = out_bitmap_or_null;
  if (!bm) {
    tmp.resize(words);
    bm = tmp.data();
  }
  HS_TRY(finish_bitmap(c, n, bm));
  int ok = 1;
  for (size_t w = 0; w < words; w++) {
    uint32_t want = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
    if (bm[w] != want) ok = 0;
  }
  *all_ok = ok;
  return HS_OK;
}

int hs_verify_committee(hs_ctx *c, const uint32_t *vidx, const uint8_t *sig, const uint32_t *midx, const uint8_t *digests, size_t n_msgs,
                        size_t n, uint32_t mode, uint32_t *out_bitmap) {
  if (!c || mode > 1 || (n && (!vidx || !sig || !digests || !out_bitmap || n_msgs == 0)) || (n && !midx && n_msgs != 1))
    return fail(c, HS_ERR_ARG, "hs_verify_committee: bad argument");
  if (n == 0) return HS_OK;
  if (midx)
    for (size_t i = 0; i < n; i++)
      if (midx[i] >= n_msgs) return fail(c, HS_ERR_ARG, "hs_verify_committee: msg_idx out of range");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  if (small_eligible(c, n) && c->explicit_committee) {  // latency path
    for (size_t i = 0; i < n; i++) {
      memcpy(c->h_small_in[i].sig, sig + 64 * i, 64);
      memcpy(c->h_small_in[i].msg, digests + 32 * (size_t)(midx ? midx[i] : 0), 32);
      c->h_small_in[i].vidx = vidx[i];
    }
    return run_small(c, n, mode, out_bitmap, nullptr);
  }
  size_t o_sig = 0, o_v = n * 64, o_m = o_v + n * 4, o_d = o_m + (midx ? n * 4 : 0), total = o_d + n_msgs * 32;
  HS_TRY(ensure(c, c->in[0], total));
  HS_TRY(ensure(c, c->out, ((n + 31) / 32) * 4));
  uint8_t *d = (uint8_t *)c->in[0].p;
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_v, vidx, n * 4, cudaMemcpyHostToDevice, c->stream));
  if (midx) HS_CUDA(c, cudaMemcpyAsync(d + o_m, midx, n * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_d, digests, n_msgs * 32, cudaMemcpyHostToDevice, c->stream));
  HS_TRY(hs_verify_committee_dev(c, d + o_v, d + o_sig, midx ? d + o_m : nullptr, d + o_d, n, mode, c->out.p, c->stream));
  return finish_bitmap(c, n, out_bitmap);
}

int hs_digest32_batch(hs_ctx *c, const uint8_t *data, const uint64_t *off, size_t n, uint8_t *out) {
  if (!c || (n && (!off || !out))) return fail(c, HS_ERR_ARG, "hs_digest32_batch: bad argument");
  if (n == 0) return HS_OK;
  if (!offsets_ok(off, n)) return fail(c, HS_ERR_ARG, "hs_digest32_batch: offsets must start at 0 and be non-decreasing");
  if (off[n] && !data) return fail(c, HS_ERR_ARG, "hs_digest32_batch: null data");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  size_t o_off = 0, o_data = (n + 1) * 8, total = o_data + off[n];
  HS_TRY(ensure(c, c->in[0], total + 8));
  HS_TRY(ensure(c, c->out, n * 32));
  uint8_t *d = (uint8_t *)c->in[0].p;
  HS_CUDA(c, cudaMemcpyAsync(d + o_off, off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (off[n]) HS_CUDA(c, cudaMemcpyAsync(d + o_data, data, off[n], cudaMemcpyHostToDevice, c->stream));
  if (n <= 64 && off[n] / n >= 1024) {
    // a few long messages (mempool batches): one warp per message, schedules expanded in parallel across lanes
    k_digest32_long<<<(unsigned)n, 32, 0, c->stream>>>(d + o_data, (const uint64_t *)(d + o_off), n, (uint32_t *)c->out.p);
    c->launches++;
    HS_CUDA(c, cudaGetLastError());
  } else {
    HS_TRY(hs_digest32_dev(c, d + o_data, d + o_off, n, c->out.p, c->stream));
  }
  HS_CUDA(c, cudaMemcpyAsync(out, c->out.p, n * 32, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

// Reference-shaped end-to-end call: verdict_i = Signature::verify(Digest(msg_i), key_i) for fixed-size messages, with the
// H2D copy of chunk j+1 overlapped with the kernels of chunk j (two streams, two staging buffers).
int hs_verify_msgs(hs_ctx *c, const uint8_t *sig, const uint8_t *pk, const uint32_t *vidx, const uint8_t *msgs, size_t msg_len, size_t n,
                   uint32_t mode, uint32_t *out_bitmap) {
  if (!c || mode > 1 || (n && (!sig || (!pk && !vidx) || !msgs || !out_bitmap || msg_len == 0)))
    return fail(c, HS_ERR_ARG, "hs_verify_msgs: bad argument");
  if (n == 0) return HS_OK;
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  size_t CH = 1u << 17;  // records per chunk (multiple of 32).  Measured on B200 (512 B messages, 0.3 % unknown keys): 2^15 -> 2.9e7/s, 2^16 -> 4.5e7,
                         // 2^17 -> 7.4e7, 2^18 -> 7.5e7 verifies/s end to end: every chunk with an unknown key waits ~0.84 ms for the generic pass,
                         // so chunks must be long enough for the PCIe copy of the next chunk to cover it
  if (const char *e = getenv("HS_CHUNK_RECORDS")) {
    size_t v = strtoull(e, nullptr, 10);
    if (v >= 1024) CH = v & ~(size_t)31;
  }
  const size_t key_bytes = vidx ? 4 : 32;
  const size_t per_rec = 64 + key_bytes + msg_len;
  const size_t chunk_cap = (n < CH ? n : CH);
  HS_TRY(ensure(c, c->out, ((n + 31) / 32) * 4));
  // scratch shared by both chunks' verify passes is stream-ordered: verify of chunk j+1 is enqueued on the same compute
  // stream after chunk j, only the copies run ahead on the second stream.
  for (int b = 0; b < 2; b++) {
    HS_TRY(ensure(c, c->in[b], chunk_cap * per_rec + 64));
    HS_TRY(ensure(c, c->digest[b], chunk_cap * 32));
  }
  HS_TRY(ensure(c, c->xyz, chunk_cap * 3 * sizeof(fe)));
  HS_TRY(ensure(c, c->meta, chunk_cap));
  if (!vidx && c->n_keys) {
    HS_TRY(ensure(c, c->vidx, chunk_cap * 4));
    HS_TRY(ensure(c, c->miss, chunk_cap * 4));
  }
  // (A short "ramp" first chunk was tried and measured slower — 7.1e7 vs 7.6e7 verifies/s: every extra chunk costs one more
  // generic-pass latency when the batch contains unknown keys.)
  const size_t first = 0;
  size_t lo = 0;
  for (size_t j = 0; lo < n; j++) {
    const int b = (int)(j & 1);
    const size_t want = (j == 0 && first) ? first : CH;
    const size_t cnt = (n - lo < want) ? (n - lo) : want;
    uint8_t *d = (uint8_t *)c->in[b].p;
    const size_t o_sig = 0, o_key = cnt * 64, o_msg = o_key + ((cnt * key_bytes + 15) & ~(size_t)15);
    if (j >= 2) HS_CUDA(c, cudaStreamWaitEvent(c->stream2, c->ev_done[b], 0));  // staging buffer b is free again
    HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig + lo * 64, cnt * 64, cudaMemcpyHostToDevice, c->stream2));
    if (vidx) HS_CUDA(c, cudaMemcpyAsync(d + o_key, vidx + lo, cnt * 4, cudaMemcpyHostToDevice, c->stream2));
    else HS_CUDA(c, cudaMemcpyAsync(d + o_key, pk + lo * 32, cnt * 32, cudaMemcpyHostToDevice, c->stream2));
    HS_CUDA(c, cudaMemcpyAsync(d + o_msg, msgs + lo * msg_len, cnt * msg_len, cudaMemcpyHostToDevice, c->stream2));
    HS_CUDA(c, cudaEventRecord(c->ev[b], c->stream2));
    HS_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev[b], 0));
    HS_TRY(hs_verify_msgs_dev(c, d + o_sig, vidx ? nullptr : d + o_key, vidx ? d + o_key : nullptr, d + o_msg, msg_len, cnt, mode,
                              c->digest[b].p, (uint32_t *)c->out.p + lo / 32, c->stream));
    HS_CUDA(c, cudaEventRecord(c->ev_done[b], c->stream));
    lo += cnt;
  }
  return finish_bitmap(c, n, out_bitmap);
}

}  // extern "C"
