// hs_crypto.hpp — header-only C++ mirror of asonnino/hotstuff's `crypto` crate surface over the C ABI (hs_crypto.h).
//
// The reference is compiled Rust (crypto/src/lib.rs); with no Rust toolchain in the build image, this is the compiled-language
// host side above the ABI: same type names, argument meaning and error behaviour as the crate —
//   Digest (lib.rs:22), PublicKey (lib.rs:66), Signature{verify (lib.rs:200-204), verify_batch (lib.rs:206-219)}, CryptoError (lib.rs:18).
// Signing (Signature::new, SignatureService) stays on the CPU in the reference node and is not part of the GPU path.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <utility>
#include <vector>

#include "hs_crypto.h"

namespace hs {

// `CryptoError = ed25519::Error`: opaque, only Ok/Err is observable (consensus maps it to InvalidSignature, error.rs:39).
struct CryptoError : std::runtime_error {
  CryptoError() : std::runtime_error("signature error") {}
};
// Engine failure (CUDA error / bad argument): NOT a verdict.  Callers treat it as reject (core.rs:434-439 drops the message).
struct EngineError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

class Engine {
 public:
  explicit Engine(int device = 0, uint32_t flags = 0) {
    if (hs_ctx_create(&ctx_, device, flags) != HS_OK) throw EngineError("hs_ctx_create failed (no GPU?) — there is no CPU fallback");
  }
  ~Engine() { hs_ctx_destroy(ctx_); }
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;
  hs_ctx *raw() const { return ctx_; }
  void check(int rc, const char *what) const {
    if (rc != HS_OK) throw EngineError(std::string(what) + ": " + hs_last_error(ctx_));
  }

 private:
  hs_ctx *ctx_ = nullptr;
};

struct Digest {  // crypto/src/lib.rs:22
  std::array<uint8_t, 32> bytes{};
  size_t size() const { return 32; }
  std::vector<uint8_t> to_vec() const { return {bytes.begin(), bytes.end()}; }
  bool operator==(const Digest &o) const { return bytes == o.bytes; }
  // Digest(SHA-512(data)[..32]) — the Hash impls of consensus/src/messages.rs and mempool/src/processor.rs:30
  static Digest of(const Engine &e, const uint8_t *data, size_t len) {
    Digest d;
    const uint64_t off[2] = {0, (uint64_t)len};
    e.check(hs_digest32_batch(e.raw(), data, off, 1, d.bytes.data()), "hs_digest32_batch");
    return d;
  }
};

struct PublicKey {  // crypto/src/lib.rs:66
  std::array<uint8_t, 32> bytes{};
  bool operator==(const PublicKey &o) const { return bytes == o.bytes; }
};

struct Signature {  // crypto/src/lib.rs:179-182; default = 64 zero bytes (the "invalid" signature of crypto_tests.rs:111)
  std::array<uint8_t, 32> part1{}, part2{};
  static Signature from_bytes(const uint8_t b[64]) {
    Signature s;
    std::memcpy(s.part1.data(), b, 32);
    std::memcpy(s.part2.data(), b + 32, 32);
    return s;
  }
  std::array<uint8_t, 64> flatten() const {  // lib.rs:193-198
    std::array<uint8_t, 64> f;
    std::memcpy(f.data(), part1.data(), 32);
    std::memcpy(f.data() + 32, part2.data(), 32);
    return f;
  }
  // Signature::verify (lib.rs:200-204): dalek verify_strict.  Throws CryptoError on Err.
  void verify(const Engine &e, const Digest &digest, const PublicKey &public_key) const {
    hs_rec128 rec;
    const auto f = flatten();
    std::memcpy(rec.sig, f.data(), 64);
    std::memcpy(rec.pk, public_key.bytes.data(), 32);
    std::memcpy(rec.msg, digest.bytes.data(), 32);
    uint32_t word = 0;
    e.check(hs_verify_strict_batch(e.raw(), &rec, 1, &word), "hs_verify_strict_batch");
    if (!(word & 1u)) throw CryptoError();
  }
  // Signature::verify_batch (lib.rs:206-219): one digest, votes = (PublicKey, Signature) pairs.
  static void verify_batch(const Engine &e, const Digest &digest, const std::vector<std::pair<PublicKey, Signature>> &votes) {
    std::vector<hs_vote> v(votes.size());
    for (size_t i = 0; i < votes.size(); i++) {
      std::memcpy(v[i].pk, votes[i].first.bytes.data(), 32);
      const auto f = votes[i].second.flatten();
      std::memcpy(v[i].sig, f.data(), 64);
    }
    int ok = 0;
    e.check(hs_verify_batch_shared_msg(e.raw(), digest.bytes.data(), v.data(), v.size(), &ok, nullptr), "hs_verify_batch_shared_msg");
    if (!ok) throw CryptoError();
  }
};

}  // namespace hs
