// hs_consensus.hpp — header-only C++ mirror of the crypto-relevant parts of consensus/src/messages.rs over the C ABI: the digest
// layouts, the verify() pre-checks that decide what reaches the engine, and their error names (consensus/src/error.rs).
//   Committee (config.rs:28-72), QC (messages.rs:165-208), TC (:283-315), Vote (:104-156), Timeout (:223-275).
// Same behaviour and error order as the reference; the signature work goes through the engine's batch front ends
// (hs_verify_batch_shared_msg, hs_verify_qcs, hs_verify_tcs).  tests/cpp/consensus_tests.cpp ports messages_tests.rs.
#pragma once
#include <list>
#include <map>
#include <set>
#include <string>
#include <tuple>

#include "hs_crypto.hpp"

namespace hs {

// consensus/src/error.rs: AuthorityReuse / UnknownAuthority / QCRequiresQuorum / TCRequiresQuorum / InvalidSignature
struct ConsensusError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
using Round = uint64_t;
using Stake = uint32_t;

struct Committee {  // config.rs:28-72
  std::map<std::array<uint8_t, 32>, Stake> stakes;
  Stake stake(const PublicKey &name) const {
    auto it = stakes.find(name.bytes);
    return it == stakes.end() ? 0 : it->second;
  }
  Stake quorum_threshold() const {  // 2 * total / 3 + 1  (config.rs:66-72)
    Stake total = 0;
    for (auto &kv : stakes) total += kv.second;
    return 2 * total / 3 + 1;
  }
  // hand the validator set to the engine once per epoch (per-validator tables, latency path)
  void register_with(const Engine &e) const {
    std::vector<uint8_t> keys;
    for (auto &kv : stakes) keys.insert(keys.end(), kv.first.begin(), kv.first.end());
    e.check(hs_committee_register(e.raw(), keys.data(), stakes.size(), nullptr), "hs_committee_register");
  }
};

inline void put_le64(uint8_t *p, uint64_t v) {
  for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i));
}

struct QC {  // messages.rs:165-169
  Digest hash;
  Round round = 0;
  std::vector<std::pair<PublicKey, Signature>> votes;
  bool is_genesis() const { return round == 0 && hash == Digest{}; }  // PartialEq compares hash and round (messages.rs:216-220)
  std::array<uint8_t, 40> preimage() const {                           // QC::digest preimage = hash || round_le (messages.rs:201-208)
    std::array<uint8_t, 40> p;
    std::memcpy(p.data(), hash.bytes.data(), 32);
    put_le64(p.data() + 32, round);
    return p;
  }
  void check_quorum(const Committee &c) const {  // messages.rs:182-194 — BEFORE any crypto
    Stake weight = 0;
    std::set<std::array<uint8_t, 32>> used;
    for (auto &v : votes) {
      if (used.count(v.first.bytes)) throw ConsensusError("AuthorityReuse");
      Stake s = c.stake(v.first);
      if (s == 0) throw ConsensusError("UnknownAuthority");
      used.insert(v.first.bytes);
      weight += s;
    }
    if (weight < c.quorum_threshold()) throw ConsensusError("QCRequiresQuorum");
  }
  void verify(const Engine &e, const Committee &c) const {  // messages.rs:180-198
    check_quorum(c);
    const auto pre = preimage();
    try {
      Signature::verify_batch(e, Digest::of(e, pre.data(), pre.size()), votes);
    } catch (const CryptoError &) {
      throw ConsensusError("InvalidSignature");
    }
  }
};

// Many QCs in ONE engine call (the view-change burst, core.rs:227): pre-checks on the host, digests + votes + per-QC AND on the GPU.
inline std::vector<bool> verify_qcs(const Engine &e, const Committee &c, const std::vector<QC> &qcs) {
  std::vector<bool> ok(qcs.size(), true);
  std::vector<uint8_t> pre, pk, sig;
  std::vector<uint32_t> qi, live;
  for (size_t j = 0; j < qcs.size(); j++) {
    try {
      qcs[j].check_quorum(c);
    } catch (const ConsensusError &) {
      ok[j] = false;
      continue;
    }
    const auto p = qcs[j].preimage();
    pre.insert(pre.end(), p.begin(), p.end());
    for (auto &v : qcs[j].votes) {
      pk.insert(pk.end(), v.first.bytes.begin(), v.first.bytes.end());
      const auto f = v.second.flatten();
      sig.insert(sig.end(), f.begin(), f.end());
      qi.push_back((uint32_t)live.size());
    }
    live.push_back((uint32_t)j);
  }
  if (live.empty()) return ok;
  std::vector<uint32_t> bm((live.size() + 31) / 32 + 1);
  e.check(hs_verify_qcs(e.raw(), pre.data(), live.size(), pk.data(), nullptr, sig.data(), qi.data(), qi.size(), nullptr, bm.data()), "hs_verify_qcs");
  for (size_t k = 0; k < live.size(); k++) ok[live[k]] = (bm[k / 32] >> (k % 32)) & 1u;
  return ok;
}

// Verified-QC cache: during a view change every Timeout carries its sender's high_qc, so a node re-verifies the same
// certificate up to N times (core.rs:227 -> Timeout::verify -> QC::verify).  A hit requires the candidate's BYTES — (hash, round) and every
// (name, signature) — to equal a certificate that already verified, so the verdict is exactly the reference's; bounded LRU keyed by (hash, round).
class VerifiedQcCache {
 public:
  explicit VerifiedQcCache(size_t capacity = 1024) : cap_(capacity) {}
  static std::vector<uint8_t> wire(const QC &q) {
    std::vector<uint8_t> w(q.hash.bytes.begin(), q.hash.bytes.end());
    uint8_t r[8];
    put_le64(r, q.round);
    w.insert(w.end(), r, r + 8);
    for (auto &v : q.votes) {
      w.insert(w.end(), v.first.bytes.begin(), v.first.bytes.end());
      const auto f = v.second.flatten();
      w.insert(w.end(), f.begin(), f.end());
    }
    return w;
  }
  bool known_valid(const QC &q) {
    auto it = map_.find(key(q));
    if (it == map_.end() || it->second->second != wire(q)) {
      misses++;
      return false;
    }
    lru_.splice(lru_.begin(), lru_, it->second);
    hits++;
    return true;
  }
  void remember(const QC &q) {
    auto it = map_.find(key(q));
    if (it != map_.end()) lru_.erase(it->second);
    lru_.emplace_front(key(q), wire(q));
    map_[key(q)] = lru_.begin();
    while (lru_.size() > cap_) {
      map_.erase(lru_.back().first);
      lru_.pop_back();
    }
  }
  size_t hits = 0, misses = 0;

 private:
  using Key = std::pair<std::array<uint8_t, 32>, Round>;
  static Key key(const QC &q) { return {q.hash.bytes, q.round}; }
  size_t cap_;
  std::list<std::pair<Key, std::vector<uint8_t>>> lru_;
  std::map<Key, std::list<std::pair<Key, std::vector<uint8_t>>>::iterator> map_;
};

struct TC {  // messages.rs:283-287
  Round round = 0;
  std::vector<std::tuple<PublicKey, Signature, Round>> votes;
  void verify(const Engine &e, const Committee &c) const {  // messages.rs:290-315
    Stake weight = 0;
    std::set<std::array<uint8_t, 32>> used;
    for (auto &v : votes) {
      if (used.count(std::get<0>(v).bytes)) throw ConsensusError("AuthorityReuse");
      Stake s = c.stake(std::get<0>(v));
      if (s == 0) throw ConsensusError("UnknownAuthority");
      used.insert(std::get<0>(v).bytes);
      weight += s;
    }
    if (weight < c.quorum_threshold()) throw ConsensusError("TCRequiresQuorum");
    std::vector<uint8_t> pk, sig;
    std::vector<uint64_t> hq;
    std::vector<uint32_t> ti(votes.size(), 0);
    for (auto &v : votes) {
      pk.insert(pk.end(), std::get<0>(v).bytes.begin(), std::get<0>(v).bytes.end());
      const auto f = std::get<1>(v).flatten();
      sig.insert(sig.end(), f.begin(), f.end());
      hq.push_back(std::get<2>(v));
    }
    uint32_t tbm[2] = {0, 0};
    const uint64_t r = round;
    e.check(hs_verify_tcs(e.raw(), &r, 1, pk.data(), nullptr, sig.data(), hq.data(), ti.data(), votes.size(), nullptr, tbm), "hs_verify_tcs");
    if (!(tbm[0] & 1u)) throw ConsensusError("InvalidSignature");  // messages.rs:307-313, one engine call instead of n
  }
};

struct Vote {  // messages.rs:104-110
  Digest hash;
  Round round = 0;
  PublicKey author;
  Signature signature;
  void verify(const Engine &e, const Committee &c) const {  // messages.rs:136-146
    if (c.stake(author) == 0) throw ConsensusError("UnknownAuthority");
    uint8_t pre[40];
    std::memcpy(pre, hash.bytes.data(), 32);
    put_le64(pre + 32, round);
    try {
      signature.verify(e, Digest::of(e, pre, 40), author);
    } catch (const CryptoError &) {
      throw ConsensusError("InvalidSignature");
    }
  }
};

struct Timeout {  // messages.rs:223-228
  QC high_qc;
  Round round = 0;
  PublicKey author;
  Signature signature;
  void verify(const Engine &e, const Committee &c) const {  // messages.rs:250-265
    if (c.stake(author) == 0) throw ConsensusError("UnknownAuthority");
    const uint64_t r = round, hq = high_qc.round;
    const auto f = signature.flatten();
    uint32_t bm[2] = {0, 0};
    e.check(hs_verify_tcs(e.raw(), &r, 1, author.bytes.data(), nullptr, f.data(), &hq, nullptr, 1, nullptr, bm), "hs_verify_tcs");
    if (!(bm[0] & 1u)) throw ConsensusError("InvalidSignature");
    if (!high_qc.is_genesis()) high_qc.verify(e, c);
  }
};

// Timeout::verify for a burst (core.rs:227: one timeout per validator during a view change): the n author signatures in ONE hs_verify_tcs
// call (digests built on the GPU), the embedded high_qcs — mostly the same certificate n times — through verify_qcs with the exact-match
// cache.  Returns "" (valid) or the name of the ConsensusError the reference raises first, per timeout.
inline std::vector<std::string> verify_timeouts(const Engine &e, const Committee &c, const std::vector<Timeout> &ts, VerifiedQcCache &cache) {
  std::vector<std::string> out(ts.size());
  std::vector<uint64_t> rounds, hq;
  std::vector<uint8_t> pk, sig;
  std::vector<size_t> live;
  for (size_t j = 0; j < ts.size(); j++) {
    if (c.stake(ts[j].author) == 0) {
      out[j] = "UnknownAuthority";
      continue;
    }
    rounds.push_back(ts[j].round);
    hq.push_back(ts[j].high_qc.round);
    pk.insert(pk.end(), ts[j].author.bytes.begin(), ts[j].author.bytes.end());
    const auto f = ts[j].signature.flatten();
    sig.insert(sig.end(), f.begin(), f.end());
    live.push_back(j);
  }
  if (live.empty()) return out;
  std::vector<uint32_t> bm((live.size() + 31) / 32 + 1);
  e.check(hs_verify_tcs(e.raw(), rounds.data(), live.size(), pk.data(), nullptr, sig.data(), hq.data(), nullptr, live.size(), nullptr, bm.data()), "hs_verify_tcs");
  std::vector<QC> todo;
  std::vector<size_t> todo_of;       // timeout index of each certificate still to verify
  std::vector<std::pair<size_t, size_t>> alias;  // (timeout, index into todo) for byte-identical certificates inside this burst
  for (size_t k = 0; k < live.size(); k++) {
    const size_t j = live[k];
    if (!((bm[k / 32] >> (k % 32)) & 1u)) {
      out[j] = "InvalidSignature";
      continue;
    }
    const QC &q = ts[j].high_qc;
    if (q.is_genesis()) continue;
    try {
      q.check_quorum(c);
    } catch (const ConsensusError &ex) {
      out[j] = ex.what();
      continue;
    }
    if (cache.known_valid(q)) continue;
    const auto w = VerifiedQcCache::wire(q);
    size_t same = todo.size();
    for (size_t t = 0; t < todo.size(); t++)
      if (VerifiedQcCache::wire(todo[t]) == w) same = t;
    if (same < todo.size()) {
      alias.push_back({j, same});
    } else {
      todo.push_back(q);
      todo_of.push_back(j);
    }
  }
  if (!todo.empty()) {
    const auto ok = verify_qcs(e, c, todo);
    for (size_t t = 0; t < todo.size(); t++) {
      if (ok[t]) cache.remember(todo[t]);
      else out[todo_of[t]] = "InvalidSignature";
    }
    for (auto &a : alias)
      if (!ok[a.second]) out[a.first] = "InvalidSignature";
  }
  return out;
}

}  // namespace hs
