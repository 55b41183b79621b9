// C++ mirror of the reference's crypto/src/tests/crypto_tests.rs (verify_valid_signature :50-61, verify_invalid_signature :64-77,
// verify_valid_batch :80-94, verify_invalid_batch :97-115) and the mempool batch_digest fixture, through include/hs_crypto.hpp.
// Fixtures (keys(), signatures) come on the command line as hex, produced by tests/golden/gen_golden.py — see
// tests/test_cpp_mirror.py.  Exit code 0 = all assertions held.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/hs_crypto.hpp"

static std::vector<uint8_t> unhex(const std::string &h) {
  std::vector<uint8_t> out(h.size() / 2);
  for (size_t i = 0; i < out.size(); i++) out[i] = (uint8_t)std::stoi(h.substr(2 * i, 2), nullptr, 16);
  return out;
}
#define REQUIRE(cond)                                              \
  do {                                                             \
    if (!(cond)) {                                                 \
      std::fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond); \
      return 1;                                                    \
    }                                                              \
  } while (0)

int main(int argc, char **argv) {
  // argv: hello_digest bad_digest sig_key3 pk3 pk2 pk1 sig2_hello sig1_hello serialized_batch batch_digest
  if (argc != 11) return 2;
  hs::Engine e(0);
  const std::string hello = "Hello, world!", bad = "Bad message!";
  hs::Digest d = hs::Digest::of(e, (const uint8_t *)hello.data(), hello.size());
  hs::Digest dbad = hs::Digest::of(e, (const uint8_t *)bad.data(), bad.size());
  REQUIRE(d.to_vec() == unhex(argv[1]) && dbad.to_vec() == unhex(argv[2]));
  hs::PublicKey pk3, pk2, pk1;
  std::memcpy(pk3.bytes.data(), unhex(argv[4]).data(), 32);
  std::memcpy(pk2.bytes.data(), unhex(argv[5]).data(), 32);
  std::memcpy(pk1.bytes.data(), unhex(argv[6]).data(), 32);
  hs::Signature s3 = hs::Signature::from_bytes(unhex(argv[3]).data());
  s3.verify(e, d, pk3);  // verify_valid_signature
  bool threw = false;
  try { s3.verify(e, dbad, pk3); } catch (const hs::CryptoError &) { threw = true; }  // verify_invalid_signature
  REQUIRE(threw);
  hs::Signature s2 = hs::Signature::from_bytes(unhex(argv[7]).data()), s1 = hs::Signature::from_bytes(unhex(argv[8]).data());
  hs::Signature::verify_batch(e, d, {{pk3, s3}, {pk2, s2}, {pk1, s1}});  // verify_valid_batch
  threw = false;
  try { hs::Signature::verify_batch(e, d, {{pk3, s3}, {pk2, s2}, {pk1, hs::Signature{}}}); } catch (const hs::CryptoError &) { threw = true; }
  REQUIRE(threw);  // verify_invalid_batch (Signature::default())
  auto batch = unhex(argv[9]);
  REQUIRE(hs::Digest::of(e, batch.data(), batch.size()).to_vec() == unhex(argv[10]));  // mempool batch_digest()
  std::puts("cpp mirror ok");
  return 0;
}
