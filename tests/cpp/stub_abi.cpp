// TEST ONLY: the subset of include/hs_crypto.h that the C++ mirror (hs_crypto.hpp / hs_consensus.hpp) calls, answered by the CPU oracle.
// Linked INSTEAD of libhs_crypto.so (together with csrc/hs_ingest.cpp, which is host-only) so that tests/cpp/crypto_tests.cpp and
// tests/cpp/consensus_tests.cpp — the C++ ports of the reference's crypto_tests.rs / messages_tests.rs — also run on a box without a GPU.
// The product never sees this file; the same binaries run against the CUDA engine in the `-m gpu` tests.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/hs_crypto.h"
extern "C" {
#include "../../oracle/hs_oracle.h"
}

struct hs_ctx {
  uint64_t launches = 0;
  std::vector<uint8_t> committee;  // registered keys (validator_idx form)
};

static bool item_ok(const uint8_t *sig, const uint8_t *pk, const uint8_t digest[32], uint32_t mode) {
  const unsigned fl = hso_verify_flags(sig, pk, digest, 32);
  return mode == HS_MODE_BATCH_EQ ? (fl & HSO_EQ_OK) != 0 : (fl & HSO_STRICT) != 0;
}
static void ones(uint32_t *bm, size_t n) {
  for (size_t w = 0; w < (n + 31) / 32; w++) bm[w] = (w == (n + 31) / 32 - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
}
static const uint8_t *key_of(hs_ctx *c, const uint8_t *pk, const uint32_t *vidx, size_t i) {
  return pk ? pk + 32 * i : c->committee.data() + 32 * (size_t)vidx[i];
}

extern "C" {
int hs_ctx_create(hs_ctx **out, int, uint32_t) {
  *out = new hs_ctx;
  return HS_OK;
}
void hs_ctx_destroy(hs_ctx *c) { delete c; }
const char *hs_last_error(const hs_ctx *) { return "oracle stub"; }
uint64_t hs_kernel_launches(const hs_ctx *c) { return c->launches; }
int hs_committee_register(hs_ctx *c, const uint8_t *pks, size_t n, uint32_t *valid) {
  c->committee.assign(pks, pks + 32 * n);
  if (valid) {
    memset(valid, 0, ((n + 31) / 32) * 4);
    for (size_t i = 0; i < n; i++)
      if (hso_point_decompress_ok(pks + 32 * i)) valid[i / 32] |= 1u << (i % 32);
  }
  return HS_OK;
}
int hs_digest32_batch(hs_ctx *c, const uint8_t *data, const uint64_t *off, size_t n, uint8_t *out) {
  c->launches += 1;
  hso_digest32_batch(data, off, n, out);
  return HS_OK;
}
int hs_verify_strict_batch(hs_ctx *c, const hs_rec128 *recs, size_t n, uint32_t *bm) {
  c->launches += 3;
  memset(bm, 0, ((n + 31) / 32) * 4);
  for (size_t i = 0; i < n; i++)
    if (item_ok(recs[i].sig, recs[i].pk, recs[i].msg, HS_MODE_STRICT)) bm[i / 32] |= 1u << (i % 32);
  return HS_OK;
}
int hs_verify_batch_shared_msg(hs_ctx *c, const uint8_t digest[32], const hs_vote *votes, size_t n, int *all_ok, uint32_t *bm) {
  c->launches += 3;
  int ok = 1;
  if (bm) memset(bm, 0, ((n + 31) / 32) * 4);
  for (size_t i = 0; i < n; i++) {
    const bool v = item_ok(votes[i].sig, votes[i].pk, digest, HS_MODE_BATCH_EQ);
    if (v && bm) bm[i / 32] |= 1u << (i % 32);
    ok &= v ? 1 : 0;
  }
  *all_ok = ok;
  return HS_OK;
}
int hs_verify_qcs(hs_ctx *c, const uint8_t *pre, size_t n_qc, const uint8_t *pk, const uint32_t *vidx, const uint8_t *sig, const uint32_t *qc_idx,
                  size_t n_votes, uint32_t *vote_bm, uint32_t *qc_bm) {
  c->launches += 5;
  ones(qc_bm, n_qc);
  if (vote_bm) memset(vote_bm, 0, ((n_votes + 31) / 32) * 4);
  for (size_t i = 0; i < n_votes; i++) {
    uint8_t d[32];
    hso_digest32(pre + 40 * (size_t)qc_idx[i], 40, d);
    const bool v = item_ok(sig + 64 * i, key_of(c, pk, vidx, i), d, HS_MODE_BATCH_EQ);
    if (v && vote_bm) vote_bm[i / 32] |= 1u << (i % 32);
    if (!v) qc_bm[qc_idx[i] / 32] &= ~(1u << (qc_idx[i] % 32));
  }
  return HS_OK;
}
int hs_verify_tcs(hs_ctx *c, const uint64_t *tc_rounds, size_t n_tc, const uint8_t *pk, const uint32_t *vidx, const uint8_t *sig,
                  const uint64_t *hq, const uint32_t *tc_idx, size_t n_votes, uint32_t *vote_bm, uint32_t *tc_bm) {
  c->launches += 4;
  ones(tc_bm, n_tc);
  if (vote_bm) memset(vote_bm, 0, ((n_votes + 31) / 32) * 4);
  for (size_t i = 0; i < n_votes; i++) {
    const size_t t = tc_idx ? tc_idx[i] : i;
    uint8_t pre[16], d[32];
    memcpy(pre, &tc_rounds[t], 8);  // little-endian host, like every box this runs on
    memcpy(pre + 8, &hq[i], 8);
    hso_digest32(pre, 16, d);
    const bool v = item_ok(sig + 64 * i, key_of(c, pk, vidx, i), d, HS_MODE_STRICT);
    if (v && vote_bm) vote_bm[i / 32] |= 1u << (i % 32);
    if (!v) tc_bm[t / 32] &= ~(1u << (t % 32));
  }
  return HS_OK;
}
int hs_verify_groups(hs_ctx *c, const uint8_t *pre, const uint64_t *pre_off, size_t n_msgs, const uint8_t *sig, const uint8_t *pk, const uint32_t *vidx,
                     const uint32_t *msg_idx, const uint32_t *group_idx, const uint8_t *mode, size_t n_items, size_t n_groups, uint32_t *item_bm,
                     uint32_t *group_bm) {
  c->launches += 5;
  std::vector<uint8_t> d(32 * (n_msgs ? n_msgs : 1));
  hso_digest32_batch(pre, pre_off, n_msgs, d.data());
  ones(group_bm, n_groups);
  if (item_bm) memset(item_bm, 0, ((n_items + 31) / 32) * 4);
  for (size_t i = 0; i < n_items; i++) {
    const bool v = item_ok(sig + 64 * i, key_of(c, pk, vidx, i), d.data() + 32 * (size_t)msg_idx[i], mode ? mode[i] : HS_MODE_STRICT);
    if (v && item_bm) item_bm[i / 32] |= 1u << (i % 32);
    if (!v) group_bm[group_idx[i] / 32] &= ~(1u << (group_idx[i] % 32));
  }
  return HS_OK;
}
}
