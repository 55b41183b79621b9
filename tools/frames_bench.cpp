// Host-side cost of the C++ receiver path around the engine call (hs::verify_frames_with, include/hs_consensus.hpp): ingest + stake /
// duplicate / quorum pre-checks + gathering the items to judge + mapping verdicts back to frames, with a no-op item verifier (every
// signature "valid") in place of hs_verify_groups.  Same synthetic Propose frames as tools/ingest_bench.cpp.
//   g++ -O2 -std=c++17 -o tools/frames_bench tools/frames_bench.cpp hotstuff_b200/libhs_crypto.so -Wl,-rpath,$PWD/hotstuff_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../include/hs_consensus.hpp"

static void put64(std::vector<uint8_t> &b, uint64_t v) {
  for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i)));
}
static void put_key(std::vector<uint8_t> &b, uint64_t seed) {
  static const char *A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  uint8_t k[33];
  for (int i = 0; i < 32; i++) k[i] = (uint8_t)((seed = seed * 6364136223846793005ULL + 1442695040888963407ULL) >> 56);
  k[32] = 0;
  put64(b, 44);
  for (int i = 0; i < 33; i += 3) {
    const uint32_t v = (k[i] << 16) | (i + 1 < 33 ? k[i + 1] << 8 : 0) | (i + 2 < 33 ? k[i + 2] : 0);
    b.push_back(A[v >> 18]);
    b.push_back(A[(v >> 12) & 63]);
    b.push_back(i == 30 ? A[(v >> 6) & 60] : A[(v >> 6) & 63]);
    b.push_back(i == 30 ? '=' : A[v & 63]);
  }
}
static std::vector<uint8_t> propose_frame(int votes, uint64_t author, uint64_t seed) {  // the same validator set signs every QC
  std::vector<uint8_t> b = {0, 0, 0, 0};
  for (int i = 0; i < 32; i++) b.push_back((uint8_t)(seed + i));
  put64(b, 7);
  put64(b, votes);
  for (int v = 0; v < votes; v++) {
    put_key(b, 1000 + v);
    for (int i = 0; i < 64; i++) b.push_back((uint8_t)(v + i));
  }
  b.push_back(0);
  put_key(b, 1000 + author);
  put64(b, 8);
  put64(b, 1);
  for (int i = 0; i < 32; i++) b.push_back((uint8_t)i);
  for (int i = 0; i < 64; i++) b.push_back((uint8_t)(i * 3));
  return b;
}

int main(int argc, char **argv) {
  const int votes = argc > 1 ? atoi(argv[1]) : 667, n_frames = argc > 2 ? atoi(argv[2]) : 2000;
  std::vector<std::vector<uint8_t>> frames;
  for (int f = 0; f < n_frames; f++) frames.push_back(propose_frame(votes, f % votes, 17 + f % 16));
  hs::Committee c;
  {
    const hs::IngestedFrames g = hs::ingest_frames({frames[0]});
    for (size_t i = 0; i < g.n_items(); i++) {
      std::array<uint8_t, 32> k;
      std::memcpy(k.data(), g.pk.data() + 32 * i, 32);
      c.stakes[k] = 1;
    }
    for (int extra = 0; (int)c.stakes.size() < votes * 3 / 2; extra++) {  // the validators that did not sign: committee = 3/2 x votes
      std::array<uint8_t, 32> k{};
      std::memcpy(k.data(), &extra, sizeof(extra));
      k[31] = 0xEE;
      c.stakes[k] = 1;
    }
  }
  double best = 1e30;
  size_t judged = 0, ok_frames = 0;
  hs::IngestedFrames reuse;  // a receiver keeps one per worker: steady state touches no fresh memory
  for (int rep = 0; rep < 6; rep++) {
    const auto t0 = std::chrono::steady_clock::now();
    const auto out = hs::verify_frames_with(c, frames, [&](const hs::IngestedFrames &k) {
      judged = k.n_items();
      return std::vector<bool>(k.n_items(), true);
    }, &reuse);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dt < best) best = dt;
    ok_frames = 0;
    for (auto &s : out) ok_frames += s.empty();
  }
  if (ok_frames != (size_t)n_frames) {
    std::fprintf(stderr, "only %zu of %d frames passed the pre-checks\n", ok_frames, n_frames);
    return 1;
  }
  std::printf("{\"metric\": \"receiver host path items/s (ingest + pre-checks + gather, engine stubbed)\", \"votes_per_qc\": %d, \"committee\": %zu, \"frames\": %d, "
              "\"items\": %zu, \"seconds\": %.6f, \"items_per_s\": %.4e, \"threads\": 1}\n",
              votes, c.stakes.size(), n_frames, judged, best, judged / best);
  return 0;
}
