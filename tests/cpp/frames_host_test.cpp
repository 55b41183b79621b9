// CPU-side test of hs::verify_frames' host logic (include/hs_consensus.hpp): real hs_ingest_consensus_frames from libhs_crypto.so, the
// reference's pre-checks and error order in C++, and — TEST ONLY — the CPU oracle standing in for hs_verify_groups, so that the
// compiled host code can be held against hotstuff_b200/wire.py::verify_frames (and through it against struct-level verification)
// on a box without a GPU.  Input file: u32 n_keys, n_keys x {32-byte key, u32 stake}, u32 n_frames, n_frames x {u32 len, bytes}.
// Output: one line per frame ("OK" or the error name).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>

#include "../../include/hs_consensus.hpp"
extern "C" {
#include "../../oracle/hs_oracle.h"
}

// `frames_host_test --block-preimage <author hex> <round> <qc hash hex> <payload digest hex>...` prints hs::Block::preimage() as hex
static int block_preimage(int argc, char **argv) {
  auto unhex = [](const char *h, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (uint8_t)std::stoi(std::string(h + 2 * i, 2), nullptr, 16);
  };
  hs::Block b;
  unhex(argv[2], b.author.bytes.data(), 32);
  b.round = std::strtoull(argv[3], nullptr, 10);
  unhex(argv[4], b.qc.hash.bytes.data(), 32);
  for (int i = 5; i < argc; i++) {
    hs::Digest d;
    unhex(argv[i], d.bytes.data(), 32);
    b.payload.push_back(d);
  }
  for (uint8_t x : b.preimage()) std::printf("%02x", x);
  std::printf("\n");
  return 0;
}

static bool read_input(const char *path, hs::Committee &c, std::vector<std::vector<uint8_t>> &frames) {
  std::ifstream in(path, std::ios::binary);
  if (!in) return false;
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  size_t at = 0;
  auto u32 = [&] {
    uint32_t v;
    std::memcpy(&v, buf.data() + at, 4);
    at += 4;
    return v;
  };
  const uint32_t n_keys = u32();
  for (uint32_t i = 0; i < n_keys; i++) {
    std::array<uint8_t, 32> k;
    std::memcpy(k.data(), buf.data() + at, 32);
    at += 32;
    c.stakes[k] = u32();
  }
  frames.resize(u32());
  for (auto &f : frames) {
    const uint32_t len = u32();
    f.assign(buf.begin() + at, buf.begin() + at + len);
    at += len;
  }
  return true;
}

// `frames_host_test --qcs <file>`: rebuilds every embedded certificate as an hs::QC from the ingested arrays and runs hs::verify_qcs_with
// (struct-level batch front end: pre-checks through CommitteeIndex, one grouped call) with the oracle as the vote verifier.
// Output: one line per frame: "-" (no certificate / genesis), "OK" or "BAD".
static int qcs_mode(const char *path) {
  hs::Committee c;
  std::vector<std::vector<uint8_t>> frames;
  if (!read_input(path, c, frames)) return 2;
  const hs::IngestedFrames g = hs::ingest_frames(frames);
  std::vector<hs::QC> qcs;
  std::vector<int> qc_of(frames.size(), -1);
  for (size_t j = 0; j < frames.size(); j++) {
    const hs_frame_info &f = g.info[j];
    if (f.kind == HS_FRAME_MALFORMED || f.qc_hi == f.qc_lo) continue;
    hs::QC q;
    const uint8_t *pre = g.preimages.data() + g.pre_off[g.msg_idx[f.qc_lo]];
    std::memcpy(q.hash.bytes.data(), pre, 32);
    std::memcpy(&q.round, pre + 32, 8);
    for (uint32_t i = f.qc_lo; i < f.qc_hi; i++) {
      hs::PublicKey k;
      std::memcpy(k.bytes.data(), g.pk.data() + 32 * i, 32);
      q.votes.push_back({k, hs::Signature::from_bytes(g.sig.data() + 64 * i)});
    }
    qc_of[j] = (int)qcs.size();
    qcs.push_back(q);
  }
  const auto ok = hs::verify_qcs_with(c, qcs, [&](const uint8_t *pre, size_t n_live, const uint8_t *pk, const uint8_t *sig, const uint32_t *qi, size_t n_votes) {
    std::vector<bool> got(n_live, true);
    for (size_t i = 0; i < n_votes; i++) {
      uint8_t d[32];
      hso_digest32(pre + 40 * qi[i], 40, d);
      if (!(hso_verify_flags(sig + 64 * i, pk + 32 * i, d, 32) & HSO_EQ_OK)) got[qi[i]] = false;
    }
    return got;
  });
  for (size_t j = 0; j < frames.size(); j++) std::printf("%s\n", qc_of[j] < 0 ? "-" : (ok[qc_of[j]] ? "OK" : "BAD"));
  return 0;
}

int main(int argc, char **argv) {
  if (argc >= 5 && std::string(argv[1]) == "--block-preimage") return block_preimage(argc, argv);
  if (argc == 3 && std::string(argv[1]) == "--qcs") return qcs_mode(argv[2]);
  if (argc != 2) return 2;
  hs::Committee c;
  std::vector<std::vector<uint8_t>> frames;
  if (!read_input(argv[1], c, frames)) return 2;
  size_t judged = 0;
#ifdef HS_TEST_ENGINE_PATH
  // the production wrapper (hs::verify_frames -> hs_verify_groups through hs::Engine); built against tests/cpp/stub_abi.cpp on a CPU box
  hs::Engine e(0);
  const auto out = hs::verify_frames(e, c, frames);
  (void)judged;
#else
  const auto out = hs::verify_frames_with(c, frames, [&](const hs::IngestedFrames &k) {
    std::vector<bool> got(k.n_items());
    for (size_t i = 0; i < k.n_items(); i++) {
      const uint32_t m = k.msg_idx[i];
      uint8_t d[32];
      hso_digest32(k.preimages.data() + k.pre_off[m], (size_t)(k.pre_off[m + 1] - k.pre_off[m]), d);
      const unsigned fl = hso_verify_flags(k.sig.data() + i * 64, k.pk.data() + i * 32, d, 32);
      got[i] = k.mode[i] == HS_MODE_BATCH_EQ ? (fl & HSO_EQ_OK) != 0 : (fl & HSO_STRICT) != 0;
    }
    judged += got.size();
    return got;
  });
#endif
  for (auto &s : out) std::printf("%s\n", s.empty() ? "OK" : s.c_str());
  std::fprintf(stderr, "judged %zu items\n", judged);
  return 0;
}
