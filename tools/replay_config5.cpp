// replay_config5.cpp — BASELINE config[4] substitute (SURVEY §8d "Config 5"), C++ against the C ABI.
//
// The Rust node cannot be built here (no cargo), so this REPLAYS the per-node crypto call stream of `fab local`
// (benchmark/fabfile.py:14-33 at 50,000 tx/s, 512 B tx, 4 nodes, 15,000 B batches) and reports per-call latency:
//   per second : ~1,707 Digest calls over ~15.3 kB serialized batches            (mempool/src/processor.rs:30)
//   per round  : 1 strict verify                      (Block::verify, consensus/src/messages.rs:64)
//                1 verify_batch of 3 votes            (QC::verify,    messages.rs:197)
//                3 strict verifies at the leader      (Vote::verify,  messages.rs:144)
// Timed with steady_clock around each C-ABI call (host pointers in, verdicts out) — what the Rust shim would see.
// The CPU column times the oracle (the restatement of the reference's dalek path) on one core for the same calls:
// it is the number the shim's CPU/GPU cut-over is chosen against.  This is a replay of the call pattern, NOT a fab run.
//
// build: g++ -O2 -std=c++17 tools/replay_config5.cpp -Iinclude -Ioracle -Lhotstuff_b200 -lhs_crypto -Loracle -lhs_oracle -o tools/replay_config5
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hs_crypto.h"
#include "hs_oracle.h"  // input synthesis (signing) and the CPU column only

using clk = std::chrono::steady_clock;
static double us_since(clk::time_point t0) { return std::chrono::duration<double, std::micro>(clk::now() - t0).count(); }
struct series {
  std::vector<double> v;
  double pct(double p) {
    std::vector<double> s = v;
    std::sort(s.begin(), s.end());
    return s.empty() ? 0 : s[(size_t)(p * (s.size() - 1))];
  }
  double mean() {
    double t = 0;
    for (double x : v) t += x;
    return v.empty() ? 0 : t / v.size();
  }
};
static void emit(const char *name, series &g, series &c, bool last) {
  printf("\"%s\": {\"gpu_p50_us\": %.1f, \"gpu_p99_us\": %.1f, \"gpu_mean_us\": %.1f, \"cpu_oracle_1core_p50_us\": %.1f, \"calls\": %zu}%s", name, g.pct(0.5), g.pct(0.99),
         g.mean(), c.pct(0.5), g.v.size(), last ? "" : ", ");
}

int main(int argc, char **argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 1000;
  hs_ctx *ctx = nullptr;
  if (hs_ctx_create(&ctx, 0, 0) != HS_OK) {
    fprintf(stderr, "hs_ctx_create failed (no GPU?)\n");
    return 1;
  }
  uint8_t seeds[4][32], pks[4][32];
  for (int i = 0; i < 4; i++) {
    for (int j = 0; j < 32; j++) seeds[i][j] = (uint8_t)(17 * i + 3 * j + 1);
    hso_keygen(seeds[i], pks[i]);
  }
  uint32_t valid = 0;
  if (hs_committee_register(ctx, &pks[0][0], 4, &valid) != HS_OK || valid != 0xf) return 2;
  // one serialized MempoolMessage::Batch (bincode: u32 tag, u64 count, per tx u64 len + bytes), 29 x 512 B
  const int tx = 512, per_batch = 15000 / tx;
  std::vector<uint8_t> batch(12 + (size_t)per_batch * (8 + tx));
  memset(batch.data(), 0, batch.size());
  uint64_t cnt = per_batch;
  memcpy(batch.data() + 4, &cnt, 8);
  for (int t = 0; t < per_batch; t++) {
    uint64_t l = tx;
    uint8_t *p = batch.data() + 12 + (size_t)t * (8 + tx);
    memcpy(p, &l, 8);
    for (int j = 0; j < tx; j++) p[8 + j] = (uint8_t)(t * 31 + j * 7);
  }
  std::vector<uint8_t> two(batch.size() * 2);
  memcpy(two.data(), batch.data(), batch.size());
  memcpy(two.data() + batch.size(), batch.data(), batch.size());
  two[batch.size() + 20] ^= 1;
  const uint64_t off1[2] = {0, batch.size()}, off2[3] = {0, batch.size(), 2 * batch.size()};
  series g_d1, g_d2, g_v1, g_b3, g_v3, c_d1, c_d2, c_v1, c_b3, c_v3;
  uint8_t dig[64], want[64];
  int bad = 0;
  for (int r = 0; r < rounds + 20; r++) {
    const bool timed = r >= 20;  // warm-up rounds
    uint8_t blk[40], bd[32];
    memset(blk, 0, sizeof(blk));
    memcpy(blk, &r, 4);
    hso_digest32(blk, 40, bd);
    hs_rec128 recs[4];
    hs_vote votes[3];
    for (int i = 0; i < 4; i++) {
      hso_sign(seeds[i], bd, 32, recs[i].sig);
      memcpy(recs[i].pk, pks[i], 32);
      memcpy(recs[i].msg, bd, 32);
    }
    for (int i = 0; i < 3; i++) {
      memcpy(votes[i].pk, pks[i + 1], 32);
      memcpy(votes[i].sig, recs[i + 1].sig, 64);
    }
    if (r % 7 == 3) recs[2].sig[9] ^= 4;  // an invalid vote now and then: verdicts must follow
    uint32_t bm = 0;
    int all_ok = 0;
    auto t0 = clk::now();
    if (hs_digest32_batch(ctx, batch.data(), off1, 1, dig) != HS_OK) return 3;
    if (timed) g_d1.v.push_back(us_since(t0));
    t0 = clk::now();
    if (hs_digest32_batch(ctx, two.data(), off2, 2, dig) != HS_OK) return 3;  // the ~1.7 batches of a round submitted together
    if (timed) g_d2.v.push_back(us_since(t0));
    t0 = clk::now();
    if (hs_verify_strict_batch(ctx, &recs[0], 1, &bm) != HS_OK) return 4;
    if (timed) g_v1.v.push_back(us_since(t0));
    bad += (bm & 1) != 1;
    t0 = clk::now();
    if (hs_verify_batch_shared_msg(ctx, bd, votes, 3, &all_ok, nullptr) != HS_OK) return 5;
    if (timed) g_b3.v.push_back(us_since(t0));
    bad += all_ok != 1;
    t0 = clk::now();
    if (hs_verify_strict_batch(ctx, &recs[1], 3, &bm) != HS_OK) return 6;
    if (timed) g_v3.v.push_back(us_since(t0));
    bad += bm != ((r % 7 == 3) ? 0x5u : 0x7u);
    // CPU column (oracle, one core)
    t0 = clk::now();
    hso_digest32(batch.data(), batch.size(), want);
    if (timed) c_d1.v.push_back(us_since(t0));
    t0 = clk::now();
    hso_digest32(two.data(), batch.size(), want);
    hso_digest32(two.data() + batch.size(), batch.size(), want + 32);
    if (timed) c_d2.v.push_back(us_since(t0));
    bad += memcmp(want, dig, 64) != 0;
    t0 = clk::now();
    int ok = hso_verify_strict(recs[0].sig, recs[0].pk, recs[0].msg, 32);
    if (timed) c_v1.v.push_back(us_since(t0));
    bad += ok != 1;
    t0 = clk::now();
    ok = hso_verify_batch_shared_msg(bd, (const uint8_t *)votes, 3, 1, nullptr);
    if (timed) c_b3.v.push_back(us_since(t0));
    t0 = clk::now();
    for (int i = 1; i < 4; i++) ok += hso_verify_strict(recs[i].sig, recs[i].pk, recs[i].msg, 32);
    if (timed) c_v3.v.push_back(us_since(t0));
  }
  const double per_round = g_d2.mean() + g_v1.mean() + g_b3.mean() + g_v3.mean();
  const double per_round_cpu = c_d2.mean() + c_v1.mean() + c_b3.mean() + c_v3.mean();
  printf("{\"what\": \"replay of the per-node crypto call stream of fab local @ 50k tx/s, 512 B tx, 4 nodes through the C ABI (not a fab run)\", \"rounds\": %d, "
         "\"mismatches\": %d, \"calls\": {",
         rounds, bad);
  emit("digest_one_15kB_batch", g_d1, c_d1, false);
  emit("digest_two_15kB_batches_one_call", g_d2, c_d2, false);
  emit("verify_strict_1", g_v1, c_v1, false);
  emit("verify_batch_3", g_b3, c_b3, false);
  emit("verify_strict_3", g_v3, c_v3, true);
  printf("}, \"crypto_us_per_round_gpu\": %.1f, \"rounds_per_s_sustainable_single_caller_gpu\": %.0f, \"crypto_us_per_round_cpu_oracle_1core\": %.1f, "
         "\"kernel_launches\": %llu}\n",
         per_round, 1e6 / per_round, per_round_cpu, (unsigned long long)hs_kernel_launches(ctx));
  hs_ctx_destroy(ctx);
  return bad ? 9 : 0;
}
