// hs_engine.cu — sm_100a kernels + the C ABI of include/hs_crypto.h.
//
// Hot path of asonnino/hotstuff's crypto crate (crypto/src/lib.rs:200-219 + the SHA-512 Digest call sites) rebuilt for
// B200: one thread verifies one signature end to end (SHA-512 -> mod l -> decompress -> double-scalar mult -> compare),
// a warp ballots 32 verdicts into one bitmap word.  No CPU path: if CUDA fails the call returns an error.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "../../include/hs_crypto.h"
#include "verify_core.cuh"

#define HS_THREADS 128

// ------------------------------------------------------------------------------------------------ record staging
// A warp loads its 32 packed 128-byte records (4 KB) with fully coalesced 16-byte accesses, parks them in shared
// memory under an XOR swizzle, and every lane then reads back its own record conflict-free.
__device__ __forceinline__ void warp_load_rec128(uint32_t (&sig_r)[8], uint32_t (&sig_s)[8], uint32_t (&pk)[8], uint32_t (&msg)[8],
                                                 const uint4 *__restrict__ recs, size_t n, size_t warp_first, uint4 *smem_warp) {
  const int lane = threadIdx.x & 31;
  const uint4 *src = recs + warp_first * 8;  // 8 x 16 B per record
#pragma unroll
  for (int j = 0; j < 8; j++) {
    int c = j * 32 + lane;  // chunk index inside the warp's 4 KB
    int rec = c >> 3, part = c & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (warp_first + rec < n) v = __ldg(src + c);
    smem_warp[rec * 8 + (part ^ (rec & 7))] = v;
  }
  __syncwarp();
  uint4 q[8];
#pragma unroll
  for (int part = 0; part < 8; part++) q[part] = smem_warp[lane * 8 + (part ^ (lane & 7))];
  __syncwarp();
  sig_r[0] = q[0].x; sig_r[1] = q[0].y; sig_r[2] = q[0].z; sig_r[3] = q[0].w; sig_r[4] = q[1].x; sig_r[5] = q[1].y; sig_r[6] = q[1].z; sig_r[7] = q[1].w;
  sig_s[0] = q[2].x; sig_s[1] = q[2].y; sig_s[2] = q[2].z; sig_s[3] = q[2].w; sig_s[4] = q[3].x; sig_s[5] = q[3].y; sig_s[6] = q[3].z; sig_s[7] = q[3].w;
  pk[0] = q[4].x; pk[1] = q[4].y; pk[2] = q[4].z; pk[3] = q[4].w; pk[4] = q[5].x; pk[5] = q[5].y; pk[6] = q[5].z; pk[7] = q[5].w;
  msg[0] = q[6].x; msg[1] = q[6].y; msg[2] = q[6].z; msg[3] = q[6].w; msg[4] = q[7].x; msg[5] = q[7].y; msg[6] = q[7].z; msg[7] = q[7].w;
}

__device__ __forceinline__ void load32(uint32_t (&w)[8], const uint8_t *p) {
  // 32-byte field at a 4-byte-aligned address (sig/pk arrays of the var / vote / committee layouts)
  const uint32_t *s = reinterpret_cast<const uint32_t *>(p);
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = __ldg(s + i);
}

__device__ __forceinline__ void emit_verdict(uint32_t fl, uint32_t mode, bool active, size_t idx, uint32_t *bitmap, uint8_t *flags_out) {
  uint32_t bit = active && ((mode == HS_MODE_STRICT) ? (fl & HS_F_STRICT) : (fl & HS_F_EQ));
  uint32_t word = __ballot_sync(0xffffffffu, bit);
  if ((threadIdx.x & 31) == 0 && active) bitmap[idx >> 5] = word;
  if (flags_out && active) flags_out[idx] = (uint8_t)fl;
}

// ------------------------------------------------------------------------------------------------ generic-key kernels
// packed records {sig, pk, msg32}: Signature::verify over n independent triples
__global__ void __launch_bounds__(HS_THREADS) k_verify_rec128(const uint4 *__restrict__ recs, size_t n, const ge_niels *__restrict__ btable,
                                                               uint32_t mode, uint32_t *__restrict__ bitmap, uint8_t *flags_out) {
  __shared__ uint4 stage[HS_THREADS / 32][256];
  const size_t idx = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t warp_first = idx & ~(size_t)31;
  if (warp_first >= n) return;
  const bool active = idx < n;
  uint32_t R[8], S[8], A[8], M[8], h[16];
  warp_load_rec128(R, S, A, M, recs, n, warp_first, stage[threadIdx.x >> 5]);
  sha512_ram32(h, R, A, M);
  ge_cached tab[9];
  uint32_t fl = verify_generic_core(R, S, A, h, btable, tab);
  emit_verdict(fl, mode, active, idx, bitmap, flags_out);
}

// variable-length messages; also serves the shared-message vote layout through strides
struct var_layout {
  const uint8_t *sig;
  const uint8_t *pk;
  const uint8_t *msgs;
  const uint64_t *off;  // nullptr -> every item uses msgs[0 .. fixed_len)
  size_t sig_stride, pk_stride;
  uint64_t fixed_len;
};
__global__ void __launch_bounds__(HS_THREADS) k_verify_var(var_layout L, size_t n, const ge_niels *__restrict__ btable, uint32_t mode,
                                                            uint32_t *__restrict__ bitmap, uint8_t *flags_out) {
  const size_t idx = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t warp_first = idx & ~(size_t)31;
  if (warp_first >= n) return;
  const bool active = idx < n;
  const size_t i = active ? idx : n - 1;
  uint32_t R[8], S[8], A[8], h[16];
  load32(R, L.sig + i * L.sig_stride);
  load32(S, L.sig + i * L.sig_stride + 32);
  load32(A, L.pk + i * L.pk_stride);
  const uint8_t *m = L.off ? L.msgs + L.off[i] : L.msgs;
  uint64_t len = L.off ? (L.off[i + 1] - L.off[i]) : L.fixed_len;
  uint64_t pre[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    pre[j] = be64_from_le32(R[2 * j], R[2 * j + 1]);
    pre[4 + j] = be64_from_le32(A[2 * j], A[2 * j + 1]);
  }
  sha512_prefix_msg(h, pre, 8, m, len);
  ge_cached tab[9];
  uint32_t fl = verify_generic_core(R, S, A, h, btable, tab);
  emit_verdict(fl, mode, active, idx, bitmap, flags_out);
}

// ------------------------------------------------------------------------------------------------ committee kernel
__global__ void __launch_bounds__(HS_THREADS) k_verify_committee(const uint32_t *__restrict__ vidx, const uint8_t *__restrict__ sig,
                                                                  const uint32_t *__restrict__ midx, const uint8_t *__restrict__ digests, size_t n,
                                                                  const uint8_t *__restrict__ pks, const uint8_t *__restrict__ key_flags, uint32_t n_keys,
                                                                  const ge_niels *__restrict__ btable, const ge_niels *__restrict__ atables, uint32_t mode,
                                                                  uint32_t *__restrict__ bitmap, uint8_t *flags_out) {
  const size_t idx = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t warp_first = idx & ~(size_t)31;
  if (warp_first >= n) return;
  const bool active = idx < n;
  const size_t i = active ? idx : n - 1;
  uint32_t v = __ldg(vidx + i);
  const bool known = v < n_keys;  // unknown authority: reject (messages.rs:57-61 rejects it before crypto)
  if (!known) v = 0;
  uint32_t R[8], S[8], A[8], M[8], h[16];
  load32(R, sig + i * 64);
  load32(S, sig + i * 64 + 32);
  load32(A, pks + (size_t)v * 32);
  load32(M, digests + (size_t)(midx ? __ldg(midx + i) : 0u) * 32);
  sha512_ram32(h, R, A, M);
  uint32_t fl = verify_committee_core(R, S, h, btable, atables + (size_t)v * HS_COMB_TABLE_NIELS, known ? key_flags[v] : 0u);
  emit_verdict(fl, mode, active, idx, bitmap, flags_out);
}

// ------------------------------------------------------------------------------------------------ table construction
// thread (point p, window w): decompress point p (or take B when encs == nullptr), optionally negate, fill one window
__global__ void __launch_bounds__(HS_THREADS) k_build_comb(const uint8_t *__restrict__ encs, size_t n_points, int negate, ge_niels *tables,
                                                            uint8_t *key_flags) {
  const size_t t = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  const size_t p = t / HS_COMB_WINDOWS;
  const int w = (int)(t % HS_COMB_WINDOWS);
  if (p >= n_points) return;
  ge_ext P;
  if (encs) {
    uint32_t e[8];
    load32(e, encs + p * 32);
    uint32_t ok = ge_decompress(P, e);
    uint32_t small = ge_enc_is_small_order(e);
    if (w == 0 && key_flags) key_flags[p] = (uint8_t)((ok & 1u) | (small << 1));
    if (!ok) ge_identity(P);  // table of a rejected key is never used for an accept (flag bit0 = 0)
  } else {
    ge_basepoint(P);
  }
  if (negate) {
    ge_ext Q;
    ge_neg(Q, P);
    P = Q;
  }
  comb_build_window(tables + p * HS_COMB_TABLE_NIELS, P, w);
}

// ------------------------------------------------------------------------------------------------ Digest kernel
__global__ void __launch_bounds__(HS_THREADS) k_digest32(const uint8_t *__restrict__ data, const uint64_t *__restrict__ off, size_t n,
                                                          uint32_t *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * HS_THREADS + threadIdx.x;
  if (i >= n) return;
  uint32_t h[16];
  uint64_t pre[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  sha512_prefix_msg(h, pre, 0, data + off[i], off[i + 1] - off[i]);
#pragma unroll
  for (int j = 0; j < 8; j++) out[i * 8 + j] = h[j];
}

// ================================================================================================ host side
struct hs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  ge_niels *d_btable = nullptr;
  // committee
  size_t n_keys = 0;
  uint8_t *d_pks = nullptr;
  uint8_t *d_key_flags = nullptr;
  ge_niels *d_atables = nullptr;
  // grow-only device scratch for the host-pointer entry points
  void *d_in = nullptr;
  size_t d_in_cap = 0;
  void *d_in2 = nullptr;
  size_t d_in2_cap = 0;
  void *d_out = nullptr;
  size_t d_out_cap = 0;
  uint64_t launches = 0;
  std::mutex mu;
  std::string err = "ok";
};

static int fail(hs_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess) {
  if (c) {
    c->err = what;
    if (e != cudaSuccess) {
      c->err += ": ";
      c->err += cudaGetErrorString(e);
    }
  }
  return code;
}
#define HS_CUDA(c, call)                                          \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return fail((c), HS_ERR_CUDA, #call, e__); \
  } while (0)

static int ensure(hs_ctx *c, void **p, size_t *cap, size_t need) {
  if (need <= *cap) return HS_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = need + need / 4 + 4096;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) return fail(c, HS_ERR_NOMEM, "cudaMalloc scratch", e);
  *cap = want;
  return HS_OK;
}
static inline unsigned blocks_for(size_t n) { return (unsigned)((n + HS_THREADS - 1) / HS_THREADS); }

extern "C" {

int hs_ctx_create(hs_ctx **out, int device, uint32_t flags) {
  (void)flags;
  if (!out) return HS_ERR_ARG;
  *out = nullptr;
  hs_ctx *c = new (std::nothrow) hs_ctx();
  if (!c) return HS_ERR_NOMEM;
  c->device = device;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_btable, sizeof(ge_niels) * HS_COMB_TABLE_NIELS);
  if (e == cudaSuccess) {
    k_build_comb<<<blocks_for(HS_COMB_WINDOWS), HS_THREADS, 0, c->stream>>>(nullptr, 1, 0, c->d_btable, nullptr);
    c->launches++;
    e = cudaGetLastError();
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
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_btable);
  cudaFree(c->d_pks);
  cudaFree(c->d_key_flags);
  cudaFree(c->d_atables);
  cudaFree(c->d_in);
  cudaFree(c->d_in2);
  cudaFree(c->d_out);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *hs_last_error(const hs_ctx *c) { return c ? c->err.c_str() : "null context"; }
uint64_t hs_kernel_launches(const hs_ctx *c) { return c ? c->launches : 0; }

void *hs_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
  return p;
}
void hs_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- device-resident entry points
int hs_verify_rec128_dev(hs_ctx *c, const void *d_recs, size_t n, uint32_t mode, void *d_bitmap, void *stream) {
  if (!c || (!d_recs && n) || (!d_bitmap && n) || mode > 1) return fail(c, HS_ERR_ARG, "hs_verify_rec128_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  k_verify_rec128<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>((const uint4 *)d_recs, n, c->d_btable, mode, (uint32_t *)d_bitmap, nullptr);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}
int hs_verify_var_dev(hs_ctx *c, const void *d_sig, const void *d_pk, const void *d_msgs, const void *d_off, size_t n, uint32_t mode,
                      void *d_bitmap, void *stream) {
  if (!c || mode > 1 || (n && (!d_sig || !d_pk || !d_off || !d_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_var_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  var_layout L{(const uint8_t *)d_sig, (const uint8_t *)d_pk, (const uint8_t *)d_msgs, (const uint64_t *)d_off, 64, 32, 0};
  k_verify_var<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>(L, n, c->d_btable, mode, (uint32_t *)d_bitmap, nullptr);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}
int hs_verify_committee_dev(hs_ctx *c, const void *d_vidx, const void *d_sig, const void *d_midx, const void *d_digests, size_t n,
                            uint32_t mode, void *d_bitmap, void *stream) {
  if (!c || mode > 1 || (n && (!d_vidx || !d_sig || !d_digests || !d_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_committee_dev: bad argument");
  if (n == 0) return HS_OK;
  if (c->n_keys == 0) return fail(c, HS_ERR_ARG, "hs_verify_committee: no committee registered");
  HS_CUDA(c, cudaSetDevice(c->device));
  k_verify_committee<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>((const uint32_t *)d_vidx, (const uint8_t *)d_sig, (const uint32_t *)d_midx,
                                                                               (const uint8_t *)d_digests, n, c->d_pks, c->d_key_flags,
                                                                               (uint32_t)c->n_keys, c->d_btable, c->d_atables, mode,
                                                                               (uint32_t *)d_bitmap, nullptr);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}
int hs_digest32_dev(hs_ctx *c, const void *d_data, const void *d_off, size_t n, void *d_out, void *stream) {
  if (!c || (n && (!d_off || !d_out))) return fail(c, HS_ERR_ARG, "hs_digest32_dev: bad argument");
  if (n == 0) return HS_OK;
  HS_CUDA(c, cudaSetDevice(c->device));
  k_digest32<<<blocks_for(n), HS_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t *)d_data, (const uint64_t *)d_off, n, (uint32_t *)d_out);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  return HS_OK;
}

// ---- host-pointer entry points
static int finish_bitmap(hs_ctx *c, size_t n, uint32_t *out_bitmap) {
  size_t words = (n + 31) / 32;
  HS_CUDA(c, cudaMemcpyAsync(out_bitmap, c->d_out, words * 4, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

int hs_verify_rec128(hs_ctx *c, const hs_rec128 *recs, size_t n, uint32_t mode, uint32_t *out_bitmap) {
  if (!c || mode > 1 || (n && (!recs || !out_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_rec128: bad argument");
  if (n == 0) return HS_OK;
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure(c, &c->d_in, &c->d_in_cap, n * sizeof(hs_rec128)))) return rc;
  if ((rc = ensure(c, &c->d_out, &c->d_out_cap, ((n + 31) / 32) * 4))) return rc;
  HS_CUDA(c, cudaMemcpyAsync(c->d_in, recs, n * sizeof(hs_rec128), cudaMemcpyHostToDevice, c->stream));
  if ((rc = hs_verify_rec128_dev(c, c->d_in, n, mode, c->d_out, c->stream))) return rc;
  return finish_bitmap(c, n, out_bitmap);
}
int hs_verify_strict_batch(hs_ctx *c, const hs_rec128 *recs, size_t n, uint32_t *out_bitmap) {
  return hs_verify_rec128(c, recs, n, HS_MODE_STRICT, out_bitmap);
}

int hs_verify_var(hs_ctx *c, const uint8_t *sig, const uint8_t *pk, const uint8_t *msgs, const uint64_t *off, size_t n, uint32_t mode,
                  uint32_t *out_bitmap) {
  if (!c || mode > 1 || (n && (!sig || !pk || !off || !out_bitmap))) return fail(c, HS_ERR_ARG, "hs_verify_var: bad argument");
  if (n == 0) return HS_OK;
  if (off[n] && !msgs) return fail(c, HS_ERR_ARG, "hs_verify_var: null msgs");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  // device layout: [sig n*64][pk n*32][off (n+1)*8][msgs] — every section 8-byte aligned
  size_t o_sig = 0, o_pk = n * 64, o_off = o_pk + n * 32, o_msg = o_off + (n + 1) * 8, total = o_msg + off[n];
  int rc;
  if ((rc = ensure(c, &c->d_in, &c->d_in_cap, total + 8))) return rc;
  if ((rc = ensure(c, &c->d_out, &c->d_out_cap, ((n + 31) / 32) * 4))) return rc;
  uint8_t *d = (uint8_t *)c->d_in;
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_pk, pk, n * 32, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_off, off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (off[n]) HS_CUDA(c, cudaMemcpyAsync(d + o_msg, msgs, off[n], cudaMemcpyHostToDevice, c->stream));
  if ((rc = hs_verify_var_dev(c, d + o_sig, d + o_pk, d + o_msg, d + o_off, n, mode, c->d_out, c->stream))) return rc;
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
  int rc;
  if ((rc = ensure(c, &c->d_in, &c->d_in_cap, n * sizeof(hs_vote) + 32))) return rc;
  if ((rc = ensure(c, &c->d_out, &c->d_out_cap, words * 4))) return rc;
  uint8_t *d = (uint8_t *)c->d_in;
  HS_CUDA(c, cudaMemcpyAsync(d, digest, 32, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + 32, votes, n * sizeof(hs_vote), cudaMemcpyHostToDevice, c->stream));
  var_layout L{d + 32 + 32, d + 32, d, nullptr, sizeof(hs_vote), sizeof(hs_vote), 32};
  k_verify_var<<<blocks_for(n), HS_THREADS, 0, c->stream>>>(L, n, c->d_btable, HS_MODE_BATCH_EQ, (uint32_t *)c->d_out, nullptr);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  uint32_t *bm = out_bitmap_or_null;
  uint32_t *tmp = nullptr;
  if (!bm) {
    tmp = new (std::nothrow) uint32_t[words];
    if (!tmp) return fail(c, HS_ERR_NOMEM, "host bitmap");
    bm = tmp;
  }
  rc = finish_bitmap(c, n, bm);
  if (rc == HS_OK) {
    int ok = 1;
    for (size_t w = 0; w < words; w++) {
      uint32_t want = (w == words - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
      if (bm[w] != want) ok = 0;
    }
    *all_ok = ok;
  }
  delete[] tmp;
  return rc;
}

int hs_committee_register(hs_ctx *c, const uint8_t *pks, size_t N, uint32_t *out_valid_bitmap) {
  if (!c || (N && !pks)) return fail(c, HS_ERR_ARG, "hs_committee_register: bad argument");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->d_pks);
  cudaFree(c->d_key_flags);
  cudaFree(c->d_atables);
  c->d_pks = nullptr;
  c->d_key_flags = nullptr;
  c->d_atables = nullptr;
  c->n_keys = 0;
  if (N == 0) return HS_OK;
  HS_CUDA(c, cudaMalloc(&c->d_pks, N * 32));
  HS_CUDA(c, cudaMalloc(&c->d_key_flags, N));
  HS_CUDA(c, cudaMalloc(&c->d_atables, N * sizeof(ge_niels) * HS_COMB_TABLE_NIELS));
  HS_CUDA(c, cudaMemcpyAsync(c->d_pks, pks, N * 32, cudaMemcpyHostToDevice, c->stream));
  k_build_comb<<<blocks_for(N * HS_COMB_WINDOWS), HS_THREADS, 0, c->stream>>>(c->d_pks, N, 1, c->d_atables, c->d_key_flags);
  c->launches++;
  HS_CUDA(c, cudaGetLastError());
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_keys = N;
  if (out_valid_bitmap) {
    uint8_t *fl = new (std::nothrow) uint8_t[N];
    if (!fl) return fail(c, HS_ERR_NOMEM, "host flags");
    cudaError_t e = cudaMemcpy(fl, c->d_key_flags, N, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
      delete[] fl;
      return fail(c, HS_ERR_CUDA, "copy key flags", e);
    }
    for (size_t w = 0; w < (N + 31) / 32; w++) out_valid_bitmap[w] = 0;
    for (size_t i = 0; i < N; i++)
      if (fl[i] & 1) out_valid_bitmap[i >> 5] |= 1u << (i & 31);
    delete[] fl;
  }
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
  size_t o_sig = 0, o_v = n * 64, o_m = o_v + n * 4, o_d = o_m + (midx ? n * 4 : 0), total = o_d + n_msgs * 32;
  int rc;
  if ((rc = ensure(c, &c->d_in, &c->d_in_cap, total))) return rc;
  if ((rc = ensure(c, &c->d_out, &c->d_out_cap, ((n + 31) / 32) * 4))) return rc;
  uint8_t *d = (uint8_t *)c->d_in;
  HS_CUDA(c, cudaMemcpyAsync(d + o_sig, sig, n * 64, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_v, vidx, n * 4, cudaMemcpyHostToDevice, c->stream));
  if (midx) HS_CUDA(c, cudaMemcpyAsync(d + o_m, midx, n * 4, cudaMemcpyHostToDevice, c->stream));
  HS_CUDA(c, cudaMemcpyAsync(d + o_d, digests, n_msgs * 32, cudaMemcpyHostToDevice, c->stream));
  if ((rc = hs_verify_committee_dev(c, d + o_v, d + o_sig, midx ? d + o_m : nullptr, d + o_d, n, mode, c->d_out, c->stream))) return rc;
  return finish_bitmap(c, n, out_bitmap);
}

int hs_digest32_batch(hs_ctx *c, const uint8_t *data, const uint64_t *off, size_t n, uint8_t *out) {
  if (!c || (n && (!off || !out))) return fail(c, HS_ERR_ARG, "hs_digest32_batch: bad argument");
  if (n == 0) return HS_OK;
  if (off[n] && !data) return fail(c, HS_ERR_ARG, "hs_digest32_batch: null data");
  std::lock_guard<std::mutex> g(c->mu);
  HS_CUDA(c, cudaSetDevice(c->device));
  size_t o_off = 0, o_data = (n + 1) * 8, total = o_data + off[n];
  int rc;
  if ((rc = ensure(c, &c->d_in, &c->d_in_cap, total + 8))) return rc;
  if ((rc = ensure(c, &c->d_out, &c->d_out_cap, n * 32))) return rc;
  uint8_t *d = (uint8_t *)c->d_in;
  HS_CUDA(c, cudaMemcpyAsync(d + o_off, off, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  if (off[n]) HS_CUDA(c, cudaMemcpyAsync(d + o_data, data, off[n], cudaMemcpyHostToDevice, c->stream));
  if ((rc = hs_digest32_dev(c, d + o_data, d + o_off, n, c->d_out, c->stream))) return rc;
  HS_CUDA(c, cudaMemcpyAsync(out, c->d_out, n * 32, cudaMemcpyDeviceToHost, c->stream));
  HS_CUDA(c, cudaStreamSynchronize(c->stream));
  return HS_OK;
}

}  // extern "C"
