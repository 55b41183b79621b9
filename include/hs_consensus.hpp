// hs_consensus.hpp — header-only C++ mirror of the crypto-relevant parts of consensus/src/messages.rs over the C ABI: the digest
// layouts, the verify() pre-checks that decide what reaches the engine, and their error names (consensus/src/error.rs).
//   Committee (config.rs:28-72), QC (messages.rs:165-208), TC (:283-315), Vote (:104-156), Timeout (:223-275), Block (:17-90).
// Same behaviour and error order as the reference; the signature work goes through the engine's batch front ends
// (hs_verify_batch_shared_msg, hs_verify_qcs, hs_verify_tcs), and the receiver path — bincode frames in, one verdict per frame out —
// through hs_ingest_consensus_frames + hs_verify_groups (verify_frames).  tests/cpp/consensus_tests.cpp ports messages_tests.rs;
// tests/cpp/frames_host_test.cpp runs verify_frames' host logic on a CPU box.
#pragma once
#include <list>
#include <map>
#include <set>
#include <string>
#include <memory>
#include <tuple>
#include <type_traits>
#include <unordered_map>

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

// Per-batch acceleration of the stake / duplicate / quorum pre-checks (messages.rs:182-194, :292-304).  Per vote the reference does a
// HashSet insert and a Committee lookup on 32-byte keys; with ordered containers that costs more host time than the GPU needs for the
// signatures once certificates carry hundreds of votes (measured: tools/frames_bench.cpp).  One hash index over the committee per batch
// (keys are curve points: their first 8 bytes are hash enough) and a per-validator stamp holding the number of the last certificate the
// validator appeared in give the same answers in O(1) per vote.
class CommitteeIndex {
 public:
  explicit CommitteeIndex(const Committee &c) : threshold_(c.quorum_threshold()) {
    index_.reserve(c.stakes.size() * 2);
    for (auto &kv : c.stakes) index_.emplace(kv.first, member{kv.second, (uint32_t)index_.size()});
    stamp_.assign(index_.size(), 0);
  }
  Stake stake(const uint8_t key[32]) const {
    auto it = index_.find(load(key));
    return it == index_.end() ? (Stake)0 : it->second.stake;
  }
  // The ConsensusError name the reference raises first for a certificate of `n` votes whose i-th key is key_of(i), or nullptr.
  // (A key without stake is never inserted into the reference's `used` set — its first appearance already fails — so for it the reuse
  // check that precedes the stake check can never fire.)
  template <class KeyOf>
  const char *certificate_error(size_t n, KeyOf &&key_of, const char *quorum_error) {
    Stake weight = 0;
    certificate_++;
    for (size_t i = 0; i < n; i++) {
      auto it = index_.find(load(key_of(i)));
      if (it == index_.end() || it->second.stake == 0) return "UnknownAuthority";
      if (stamp_[it->second.slot] == certificate_) return "AuthorityReuse";
      stamp_[it->second.slot] = certificate_;
      weight += it->second.stake;
    }
    return weight >= threshold_ ? nullptr : quorum_error;
  }

 private:
  struct member {
    Stake stake;
    uint32_t slot;
  };
  struct key_hash {
    size_t operator()(const std::array<uint8_t, 32> &k) const {
      uint64_t h;
      std::memcpy(&h, k.data(), 8);
      return (size_t)(h * 0x9E3779B97F4A7C15ull);
    }
  };
  static std::array<uint8_t, 32> load(const uint8_t *p) {
    std::array<uint8_t, 32> k;
    std::memcpy(k.data(), p, 32);
    return k;
  }
  std::unordered_map<std::array<uint8_t, 32>, member, key_hash> index_;
  std::vector<uint32_t> stamp_;
  uint32_t certificate_ = 0;
  Stake threshold_;
};

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
// `verify_votes(pre /* 40 B per live QC */, n_live, pk, sig, qi /* live QC of vote i */, n_votes) -> std::vector<bool>` (one verdict per live
// QC) is the engine in production (verify_qcs below); the tests substitute a CPU checker to run this host logic without a GPU.
template <class VoteVerifier>
inline std::vector<bool> verify_qcs_with(const Committee &c, const std::vector<QC> &qcs, VoteVerifier &&verify_votes) {
  std::vector<bool> ok(qcs.size(), true);
  size_t total = 0;
  for (auto &q : qcs) total += q.votes.size();
  std::vector<uint8_t> pre, pk, sig;
  std::vector<uint32_t> qi, live;
  pre.reserve(qcs.size() * 40);
  pk.reserve(total * 32);
  sig.reserve(total * 64);
  qi.reserve(total);
  CommitteeIndex index(c);
  for (size_t j = 0; j < qcs.size(); j++) {
    const QC &q = qcs[j];
    if (index.certificate_error(q.votes.size(), [&](size_t i) { return q.votes[i].first.bytes.data(); }, "QCRequiresQuorum")) {
      ok[j] = false;  // same outcome as QC::check_quorum (messages.rs:182-194): rejected before any crypto
      continue;
    }
    const auto p = q.preimage();
    pre.insert(pre.end(), p.begin(), p.end());
    for (auto &v : q.votes) {
      pk.insert(pk.end(), v.first.bytes.begin(), v.first.bytes.end());
      const auto f = v.second.flatten();
      sig.insert(sig.end(), f.begin(), f.end());
      qi.push_back((uint32_t)live.size());
    }
    live.push_back((uint32_t)j);
  }
  if (live.empty()) return ok;
  const std::vector<bool> got = verify_votes(pre.data(), live.size(), pk.data(), sig.data(), qi.data(), qi.size());
  if (got.size() != live.size()) throw EngineError("verify_qcs: vote verifier returned the wrong number of verdicts");
  for (size_t k = 0; k < live.size(); k++) ok[live[k]] = got[k];
  return ok;
}
inline std::vector<bool> verify_qcs(const Engine &e, const Committee &c, const std::vector<QC> &qcs) {
  return verify_qcs_with(c, qcs, [&](const uint8_t *pre, size_t n_live, const uint8_t *pk, const uint8_t *sig, const uint32_t *qi, size_t n_votes) {
    std::vector<uint32_t> bm((n_live + 31) / 32 + 1);
    e.check(hs_verify_qcs(e.raw(), pre, n_live, pk, nullptr, sig, qi, n_votes, nullptr, bm.data()), "hs_verify_qcs");
    std::vector<bool> got(n_live);
    for (size_t k = 0; k < n_live; k++) got[k] = (bm[k / 32] >> (k % 32)) & 1u;
    return got;
  });
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

struct Block {  // messages.rs:17-25 (crypto-relevant fields)
  QC qc;
  bool has_tc = false;
  TC tc;
  PublicKey author;
  Round round = 0;
  std::vector<Digest> payload;
  Signature signature;
  std::vector<uint8_t> preimage() const {  // Block::digest (messages.rs:79-90): author || round_le || payload digests || qc.hash
    std::vector<uint8_t> p(32 + 8 + 32 * payload.size() + 32);
    std::memcpy(p.data(), author.bytes.data(), 32);
    put_le64(p.data() + 32, round);
    for (size_t i = 0; i < payload.size(); i++) std::memcpy(p.data() + 40 + 32 * i, payload[i].bytes.data(), 32);
    std::memcpy(p.data() + 40 + 32 * payload.size(), qc.hash.bytes.data(), 32);
    return p;
  }
  void verify(const Engine &e, const Committee &c) const {  // messages.rs:54-76, same order: stake, signature, QC (unless genesis), TC
    if (c.stake(author) == 0) throw ConsensusError("UnknownAuthority");
    const auto pre = preimage();
    try {
      signature.verify(e, Digest::of(e, pre.data(), pre.size()), author);
    } catch (const CryptoError &) {
      throw ConsensusError("InvalidSignature");
    }
    if (!qc.is_genesis()) qc.verify(e, c);
    if (has_tc) tc.verify(e, c);
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

// ---- receiver side (consensus/src/consensus.rs:33-39,138): bincode ConsensusMessage frames -> one verdict per frame ----------------
// What the reference does per frame — bincode::deserialize, then Block / Vote / Timeout / TC::verify with their stake, duplicate and
// quorum pre-checks (messages.rs:54-76,136-156,250-275,290-315) — for MANY frames with ONE engine pass: hs_ingest_consensus_frames lays
// every signature out as an item of group (= frame) i, the pre-checks below run on the item ranges and decide which items are judged at
// all, hs_verify_groups hashes every preimage and verifies every kept item on the GPU, and the per-frame result is the error the
// reference would have raised FIRST ("" = Ok, "Malformed" = the SerializationError the receiver logs and drops).
// (Same host logic as hotstuff_b200/wire.py::verify_frames, which the tests hold against struct-level verification.)
// std::vector that does not zero-fill on resize: the ingest output arrays are hundreds of megabytes for a burst of large certificates
// and every byte the caller reads is written by hs_ingest_consensus_frames first (zero-filling them cost 3x the parsing).
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = default_init_allocator<U>;
  };
  using std::allocator<T>::allocator;
  template <class U>
  void construct(U *p) noexcept(std::is_nothrow_default_constructible<U>::value) {
    ::new (static_cast<void *>(p)) U;
  }
  template <class U, class... Args>
  void construct(U *p, Args &&...args) {
    ::new (static_cast<void *>(p)) U(std::forward<Args>(args)...);
  }
};
template <class T>
using raw_vector = std::vector<T, default_init_allocator<T>>;
struct IngestedFrames {
  std::vector<hs_frame_info> info;
  raw_vector<uint8_t> sig, pk, mode, preimages;
  raw_vector<uint32_t> msg_idx, group_idx;
  raw_vector<uint64_t> pre_off;
  size_t n_items() const { return msg_idx.size(); }
  // scratch of ingest_frames (the frames laid end to end); kept so that a receiver that reuses one IngestedFrames per worker touches
  // no fresh memory in steady state (first-touch page faults on ~300 MB per burst cost twice the parsing)
  raw_vector<uint8_t> blob;
  std::vector<uint64_t> off;
};
// `g` is overwritten; pass the same object again to reuse its buffers.
inline void ingest_frames(const std::vector<std::vector<uint8_t>> &frames, IngestedFrames &g) {
  const size_t n = frames.size();
  std::vector<uint64_t> &off = g.off;
  off.assign(n + 1, 0);
  for (size_t i = 0; i < n; i++) off[i + 1] = off[i] + frames[i].size();
  raw_vector<uint8_t> &blob = g.blob;
  blob.resize(off[n] ? off[n] : 1);
  for (size_t i = 0; i < n; i++)
    if (!frames[i].empty()) std::memcpy(blob.data() + off[i], frames[i].data(), frames[i].size());
  g.info.resize(n ? n : 1);
  // Every item costs >= 116 frame bytes; a frame holds at most 2 + (its items) preimages, 16 or 40 bytes each except the one Block
  // preimage (72 bytes + its payload digests).  First guess from that with room for modest payloads, exact retry on HS_ERR_NOMEM.
  size_t ci = blob.size() / 116 + n + 1, cm = ci + 2 * n + 1, cp = 16 * ci + 512 * n + 64;
  for (int attempt = 0; attempt < 2; attempt++) {
    g.sig.resize(ci * 64);
    g.pk.resize(ci * 32);
    g.mode.resize(ci);
    g.msg_idx.resize(ci);
    g.group_idx.resize(ci);
    g.preimages.resize(cp);
    g.pre_off.resize(cm + 1);
    hs_ingest_out o{};
    o.cap_items = ci;
    o.cap_msgs = cm;
    o.cap_pre_bytes = cp;
    o.sig = g.sig.data();
    o.pk = g.pk.data();
    o.msg_idx = g.msg_idx.data();
    o.group_idx = g.group_idx.data();
    o.mode = g.mode.data();
    o.preimages = g.preimages.data();
    o.pre_off = g.pre_off.data();
    const int rc = hs_ingest_consensus_frames(blob.data(), off.data(), n, g.info.data(), &o);
    if (rc == HS_OK) {
      g.sig.resize(o.n_items * 64);
      g.pk.resize(o.n_items * 32);
      g.mode.resize(o.n_items);
      g.msg_idx.resize(o.n_items);
      g.group_idx.resize(o.n_items);
      g.preimages.resize(o.pre_bytes);
      g.pre_off.resize(o.n_msgs + 1);
      g.info.resize(n);
      return;
    }
    if (rc != HS_ERR_NOMEM) throw EngineError("hs_ingest_consensus_frames: bad argument");
    ci = o.n_items + 1;
    cm = o.n_msgs + 1;
    cp = o.pre_bytes + 1;
  }
  throw EngineError("hs_ingest_consensus_frames: capacity retry failed");
}
inline IngestedFrames ingest_frames(const std::vector<std::vector<uint8_t>> &frames) {
  IngestedFrames g;
  ingest_frames(frames, g);
  return g;
}

// `verify_items(g) -> std::vector<bool>` judges every item of g (already reduced to the items that passed the pre-checks): the engine
// in production (verify_frames below); the tests substitute a CPU checker to run this host logic without a GPU.
template <class ItemVerifier>
inline std::vector<std::string> verify_frames_with(const Committee &c, const std::vector<std::vector<uint8_t>> &frames, ItemVerifier &&verify_items,
                                                   IngestedFrames *reuse = nullptr) {
  IngestedFrames local;
  IngestedFrames &g = reuse ? *reuse : local;
  ingest_frames(frames, g);
  const size_t n = frames.size(), items = g.n_items();
  std::vector<std::string> out(n), qc_err(n), tc_err(n);
  std::vector<char> skip(items, 0), decided(n, 0);
  CommitteeIndex index(c);
  auto stake_of = [&](size_t i) { return index.stake(g.pk.data() + i * 32); };
  auto quorum = [&](uint32_t lo, uint32_t hi, const char *err) -> std::string {  // messages.rs:182-194 / :292-304
    const char *e = index.certificate_error(hi - lo, [&](size_t i) { return g.pk.data() + (lo + i) * 32; }, err);
    return e ? e : "";
  };
  auto skip_range = [&](uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo; i < hi; i++) skip[i] = 1;
  };
  for (size_t j = 0; j < n; j++) {
    const hs_frame_info &f = g.info[j];
    if (f.kind == HS_FRAME_MALFORMED) {
      out[j] = "Malformed";
      decided[j] = 1;
      continue;
    }
    if (f.author_item != HS_NO_ITEM && stake_of(f.author_item) == 0) {  // ensure!(voting_power > 0, UnknownAuthority) comes first
      out[j] = "UnknownAuthority";
      decided[j] = 1;
      skip_range(f.qc_lo, f.qc_hi);
      skip_range(f.tc_lo, f.tc_hi);
      skip[f.author_item] = 1;
      continue;
    }
    if ((f.kind == 0 || f.kind == 2) && !f.qc_is_genesis) {  // Propose / Timeout carry a QC; genesis is not verified (messages.rs:67,261)
      qc_err[j] = quorum(f.qc_lo, f.qc_hi, "QCRequiresQuorum");
      if (!qc_err[j].empty()) {
        skip_range(f.qc_lo, f.qc_hi);
        skip_range(f.tc_lo, f.tc_hi);
      }
    }
    if (f.has_tc && qc_err[j].empty()) {
      tc_err[j] = quorum(f.tc_lo, f.tc_hi, "TCRequiresQuorum");
      if (!tc_err[j].empty()) skip_range(f.tc_lo, f.tc_hi);
    }
  }
  // one pass over the items that are still to be judged (normally all of them: then the ingest arrays go to the engine as they are)
  std::vector<char> ok(items, 0);
  size_t n_skip = 0;
  for (size_t i = 0; i < items; i++) n_skip += skip[i] ? 1 : 0;
  if (n_skip == 0) {
    if (items) {
      const std::vector<bool> got = verify_items(g);
      if (got.size() != items) throw EngineError("verify_frames: item verifier returned the wrong number of verdicts");
      for (size_t i = 0; i < items; i++) ok[i] = got[i] ? 1 : 0;
    }
  } else if (n_skip < items) {
    IngestedFrames kept;
    std::vector<size_t> kept_of;
    const size_t nk = items - n_skip;
    kept_of.reserve(nk);
    kept.sig.resize(nk * 64);
    kept.pk.resize(nk * 32);
    kept.mode.resize(nk);
    kept.msg_idx.resize(nk);
    kept.group_idx.resize(nk);
    for (size_t i = 0; i < items; i++) {
      if (skip[i]) continue;
      const size_t k = kept_of.size();
      kept_of.push_back(i);
      std::memcpy(kept.sig.data() + k * 64, g.sig.data() + i * 64, 64);
      std::memcpy(kept.pk.data() + k * 32, g.pk.data() + i * 32, 32);
      kept.mode[k] = g.mode[i];
      kept.msg_idx[k] = g.msg_idx[i];
      kept.group_idx[k] = g.group_idx[i];
    }
    kept.preimages = g.preimages;
    kept.pre_off = g.pre_off;
    kept.info = g.info;
    const std::vector<bool> got = verify_items(kept);
    if (got.size() != nk) throw EngineError("verify_frames: item verifier returned the wrong number of verdicts");
    for (size_t k = 0; k < nk; k++) ok[kept_of[k]] = got[k] ? 1 : 0;
  }
  auto all_ok = [&](uint32_t lo, uint32_t hi) {
    for (uint32_t i = lo; i < hi; i++)
      if (!ok[i]) return false;
    return true;
  };
  for (size_t j = 0; j < n; j++) {
    const hs_frame_info &f = g.info[j];
    if (decided[j] || f.kind == 4) continue;  // SyncRequest: nothing to verify
    if (f.author_item != HS_NO_ITEM && !ok[f.author_item]) out[j] = "InvalidSignature";
    else if (!qc_err[j].empty()) out[j] = qc_err[j];
    else if (!all_ok(f.qc_lo, f.qc_hi)) out[j] = "InvalidSignature";
    else if (!tc_err[j].empty()) out[j] = tc_err[j];
    else if (!all_ok(f.tc_lo, f.tc_hi)) out[j] = "InvalidSignature";
  }
  return out;
}
inline std::vector<std::string> verify_frames(const Engine &e, const Committee &c, const std::vector<std::vector<uint8_t>> &frames,
                                              IngestedFrames *reuse = nullptr) {
  return verify_frames_with(c, frames, [&](const IngestedFrames &k) {
    const size_t ni = k.n_items(), ng = k.info.size();
    std::vector<uint32_t> item_bits((ni + 31) / 32 + 1, 0), group_bits((ng + 31) / 32 + 1, 0);
    e.check(hs_verify_groups(e.raw(), k.preimages.data(), k.pre_off.data(), k.pre_off.size() - 1, k.sig.data(), k.pk.data(), nullptr, k.msg_idx.data(),
                             k.group_idx.data(), k.mode.data(), ni, ng, item_bits.data(), group_bits.data()),
            "hs_verify_groups");
    std::vector<bool> got(ni);
    for (size_t i = 0; i < ni; i++) got[i] = (item_bits[i / 32] >> (i % 32)) & 1u;
    return got;
  }, reuse);
}

}  // namespace hs
