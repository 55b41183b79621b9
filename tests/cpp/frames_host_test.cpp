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

int main(int argc, char **argv) {
  if (argc >= 5 && std::string(argv[1]) == "--block-preimage") return block_preimage(argc, argv);
  if (argc != 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  std::vector<uint8_t> buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  size_t at = 0;
  auto u32 = [&] {
    uint32_t v;
    std::memcpy(&v, buf.data() + at, 4);
    at += 4;
    return v;
  };
  hs::Committee c;
  const uint32_t n_keys = u32();
  for (uint32_t i = 0; i < n_keys; i++) {
    std::array<uint8_t, 32> k;
    std::memcpy(k.data(), buf.data() + at, 32);
    at += 32;
    c.stakes[k] = u32();
  }
  std::vector<std::vector<uint8_t>> frames(u32());
  for (auto &f : frames) {
    const uint32_t len = u32();
    f.assign(buf.begin() + at, buf.begin() + at + len);
    at += len;
  }
  size_t judged = 0;
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
  for (auto &s : out) std::printf("%s\n", s.empty() ? "OK" : s.c_str());
  std::fprintf(stderr, "judged %zu items\n", judged);
  return 0;
}
