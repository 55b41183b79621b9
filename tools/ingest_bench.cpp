// Host-side throughput of the wire ingest (hs_ingest_consensus_frames, host-only): synthetic Propose frames whose QC carries V votes
// (the shape a committee of 3V/2 validators produces), parsed on T threads (the entry point is stateless and thread-safe: each thread
// takes a contiguous range of frames and its own output arrays).  Prints one JSON line: items (= signatures) per second.
//   g++ -O2 -std=c++17 -pthread -o tools/ingest_bench tools/ingest_bench.cpp hotstuff_b200/libhs_crypto.so -Wl,-rpath,$PWD/hotstuff_b200
//   tools/ingest_bench <votes per QC> <frames> <threads>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/hs_crypto.h"

static void put64(std::vector<uint8_t> &b, uint64_t v) {
  for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i)));
}
static void put_key(std::vector<uint8_t> &b, uint64_t seed) {  // 32 pseudo-random bytes as a 44-character base64 string
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
static std::vector<uint8_t> propose_frame(int votes, uint64_t seed) {  // ConsensusMessage::Propose(Block{qc, tc: None, author, round, payload: [d], signature})
  std::vector<uint8_t> b = {0, 0, 0, 0};
  for (int i = 0; i < 32; i++) b.push_back((uint8_t)(seed + i));  // qc.hash
  put64(b, 7);                                                     // qc.round
  put64(b, votes);
  for (int v = 0; v < votes; v++) {
    put_key(b, seed * 1000 + v);
    for (int i = 0; i < 64; i++) b.push_back((uint8_t)(v + i));
  }
  b.push_back(0);  // tc: None
  put_key(b, seed);
  put64(b, 8);
  put64(b, 1);
  for (int i = 0; i < 32; i++) b.push_back((uint8_t)i);
  for (int i = 0; i < 64; i++) b.push_back((uint8_t)(i * 3));
  return b;
}

int main(int argc, char **argv) {
  const int votes = argc > 1 ? atoi(argv[1]) : 667, n_frames = argc > 2 ? atoi(argv[2]) : 2000, threads = argc > 3 ? atoi(argv[3]) : 1;
  std::vector<uint8_t> blob;
  std::vector<uint64_t> off = {0};
  for (int f = 0; f < n_frames; f++) {
    auto fr = propose_frame(votes, 17 + f % 16);
    blob.insert(blob.end(), fr.begin(), fr.end());
    off.push_back(blob.size());
  }
  const size_t items_per_frame = votes + 1;
  double best = 1e30;
  size_t total_items = 0;
  struct lane {  // one thread's frame range and (reused, already touched) output arrays: a receiver keeps these as a ring of pinned buffers
    size_t lo, n;
    std::vector<uint8_t> sig, pk, mode, pre;
    std::vector<uint32_t> mi, gi;
    std::vector<uint64_t> po, o;
    std::vector<hs_frame_info> info;
  };
  std::vector<lane> lanes(threads);
  for (int t = 0; t < threads; t++) {
    lane &L = lanes[t];
    L.lo = (size_t)n_frames * t / threads;
    L.n = (size_t)n_frames * (t + 1) / threads - L.lo;
    const size_t ci = L.n * items_per_frame + 1, cm = L.n * 2 + 1, cp = L.n * 200 + 64;
    L.sig.assign(ci * 64, 1); L.pk.assign(ci * 32, 1); L.mode.assign(ci, 1); L.pre.assign(cp, 1);
    L.mi.assign(ci, 1); L.gi.assign(ci, 1); L.po.assign(cm + 1, 1); L.o.resize(L.n + 1); L.info.resize(L.n);
    for (size_t i = 0; i <= L.n; i++) L.o[i] = off[L.lo + i] - off[L.lo];
  }
  for (int rep = 0; rep < 5; rep++) {
    std::vector<std::thread> th;
    std::vector<size_t> got(threads, 0);
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; t++)
      th.emplace_back([&, t] {
        lane &L = lanes[t];
        hs_ingest_out out{};
        out.cap_items = L.mi.size(); out.cap_msgs = L.po.size() - 1; out.cap_pre_bytes = L.pre.size();
        out.sig = L.sig.data(); out.pk = L.pk.data(); out.msg_idx = L.mi.data(); out.group_idx = L.gi.data(); out.mode = L.mode.data();
        out.preimages = L.pre.data(); out.pre_off = L.po.data();
        if (hs_ingest_consensus_frames(blob.data() + off[L.lo], L.o.data(), L.n, L.info.data(), &out) != HS_OK) abort();
        got[t] = out.n_items;
      });
    for (auto &x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (dt < best) best = dt;
    total_items = 0;
    for (size_t g : got) total_items += g;
  }
  if (total_items != (size_t)n_frames * items_per_frame) return 1;
  std::printf("{\"metric\": \"ingest items/s\", \"votes_per_qc\": %d, \"frames\": %d, \"threads\": %d, \"frame_bytes\": %zu, \"items\": %zu, "
              "\"seconds\": %.6f, \"items_per_s\": %.4e, \"frame_GBps\": %.3f, \"includes\": \"thread start/join; output arrays reused\"}\n",
              votes, n_frames, threads, (size_t)off[1], total_items, best, total_items / best, blob.size() / best / 1e9);
  return 0;
}
