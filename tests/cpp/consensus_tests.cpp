// C++ port of consensus/src/tests/messages_tests.rs:8-53 (verify_valid_qc, verify_qc_authority_reuse, verify_qc_unknown_authority,
// verify_qc_insufficient_stake) with the committee() / qc() fixtures of consensus/src/tests/common.rs:23-36,129-144, plus Vote / Timeout /
// TC and the batched QC front end, through include/hs_consensus.hpp.  Fixtures arrive as hex on the command line (tests/test_cpp_mirror.py).
#include <cstdio>
#include <cstdlib>

#include "../../include/hs_consensus.hpp"

static std::vector<uint8_t> unhex(const std::string &h) {
  std::vector<uint8_t> out(h.size() / 2);
  for (size_t i = 0; i < out.size(); i++) out[i] = (uint8_t)std::stoi(h.substr(2 * i, 2), nullptr, 16);
  return out;
}
#define REQUIRE(cond)                                                \
  do {                                                               \
    if (!(cond)) {                                                   \
      std::fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); \
      return 1;                                                      \
    }                                                                \
  } while (0)
template <class F>
static std::string error_of(F f) {
  try {
    f();
  } catch (const hs::ConsensusError &e) {
    return e.what();
  }
  return "";
}
static hs::PublicKey key(const char *hex) {
  hs::PublicKey k;
  std::memcpy(k.bytes.data(), unhex(hex).data(), 32);
  return k;
}

int main(int argc, char **argv) {
  // argv: pk0 pk1 pk2 pk3 | qc votes: pkA sigA pkB sigB pkC sigC | vote_hash vote_sig(key3, round 1) | timeout_sig(key2, round 9, high_qc = qc) | tc sigs (keys 0,1,2; round 7; hq 3,5,4)
  if (argc != 17) return 2;
  hs::Engine e(0);
  hs::Committee c;
  for (int i = 1; i <= 4; i++) c.stakes[key(argv[i]).bytes] = 1;  // stake 1 each -> quorum 3
  REQUIRE(c.quorum_threshold() == 3);
  c.register_with(e);
  hs::QC qc;
  qc.round = 1;  // hash = Digest::default()
  for (int i = 0; i < 3; i++) qc.votes.push_back({key(argv[5 + 2 * i]), hs::Signature::from_bytes(unhex(argv[6 + 2 * i]).data())});
  qc.verify(e, c);  // verify_valid_qc
  {
    hs::QC q = qc;  // verify_qc_authority_reuse
    q.votes[1] = q.votes[0];
    REQUIRE(error_of([&] { q.verify(e, c); }) == "AuthorityReuse");
  }
  {
    hs::QC q = qc;  // verify_qc_unknown_authority
    for (int i = 0; i < 32; i++) q.votes[0].first.bytes[i] = (uint8_t)i;
    REQUIRE(error_of([&] { q.verify(e, c); }) == "UnknownAuthority");
  }
  {
    hs::QC q = qc;  // verify_qc_insufficient_stake
    q.votes.pop_back();
    REQUIRE(error_of([&] { q.verify(e, c); }) == "QCRequiresQuorum");
  }
  {
    hs::QC bad = qc, other_round = qc, shortq = qc;
    bad.votes[2].second = hs::Signature{};
    other_round.round = 2;
    shortq.votes.pop_back();
    REQUIRE(error_of([&] { bad.verify(e, c); }) == "InvalidSignature");
    const auto ok = hs::verify_qcs(e, c, {qc, bad, other_round, shortq, qc});
    REQUIRE(ok == (std::vector<bool>{true, false, false, false, true}));
  }
  hs::Vote v;
  std::memcpy(v.hash.bytes.data(), unhex(argv[11]).data(), 32);
  v.round = 1;
  v.author = key(argv[4]);
  v.signature = hs::Signature::from_bytes(unhex(argv[12]).data());
  v.verify(e, c);
  v.round = 2;
  REQUIRE(error_of([&] { v.verify(e, c); }) == "InvalidSignature");
  hs::Timeout t;
  t.high_qc = qc;
  t.round = 9;
  t.author = key(argv[3]);
  t.signature = hs::Signature::from_bytes(unhex(argv[13]).data());
  t.verify(e, c);
  t.high_qc.votes[0].second = hs::Signature{};
  REQUIRE(error_of([&] { t.verify(e, c); }) == "InvalidSignature");
  {  // view-change burst: four timeouts with the SAME high_qc + one whose certificate differs in one vote: the shared one is verified once
    t.high_qc = qc;
    std::vector<hs::Timeout> burst(4, t);
    hs::Timeout forged = t;
    forged.high_qc.votes[2].second = hs::Signature{};
    burst.push_back(forged);
    hs::VerifiedQcCache cache;
    const uint64_t l0 = hs_kernel_launches(e.raw());
    auto res = hs::verify_timeouts(e, c, burst, cache);
    REQUIRE(res == (std::vector<std::string>{"", "", "", "", "InvalidSignature"}));
    const uint64_t first = hs_kernel_launches(e.raw()) - l0;
    res = hs::verify_timeouts(e, c, std::vector<hs::Timeout>(burst.begin(), burst.begin() + 4), cache);
    REQUIRE(res == (std::vector<std::string>(4, "")) && cache.hits >= 4);
    REQUIRE(hs_kernel_launches(e.raw()) - l0 - first < first);  // all four certificates came from the cache: only the timeout signatures ran
    REQUIRE(hs::verify_timeouts(e, c, {forged}, cache)[0] == "InvalidSignature");  // a hit needs identical bytes
  }
  hs::TC tc;
  tc.round = 7;
  const uint64_t hqs[3] = {3, 5, 4};
  for (int i = 0; i < 3; i++) tc.votes.push_back({key(argv[1 + i]), hs::Signature::from_bytes(unhex(argv[14 + i]).data()), hqs[i]});
  tc.verify(e, c);
  std::get<2>(tc.votes[1]) = 6;
  REQUIRE(error_of([&] { tc.verify(e, c); }) == "InvalidSignature");
  tc.votes.pop_back();
  REQUIRE(error_of([&] { tc.verify(e, c); }) == "TCRequiresQuorum");
  std::puts("cpp consensus mirror ok");
  return 0;
}
