"""Mirror of the reference's own crypto tests (crypto/src/tests/crypto_tests.rs) and the call-site fixtures
(consensus/src/tests/messages_tests.rs:8-10, mempool/src/tests/processor_tests.rs) through the Python mirror of the
crate API (hotstuff_b200.crypto), running on the GPU."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(engine):
    from hotstuff_b200 import crypto
    crypto.set_default_engine(engine)
    return crypto


def keys(golden):
    """keys(): crypto_tests.rs:26-29 (StdRng::from_seed([0;32]), 4 keypairs) — derived in tests/golden/gen_golden.py."""
    r = golden["reference"]
    return [(bytes.fromhex(p), bytes.fromhex(s)) for p, s in zip(r["pks"], r["seeds"])]


def test_import_export_public_key(api, golden):  # crypto_tests.rs:32-38
    pk = api.PublicKey(keys(golden)[-1][0])
    assert api.PublicKey.decode_base64(pk.encode_base64()) == pk
    assert pk.encode_base64() == "vq2gYSbHjZi0oaafbuYYlpTw9HUVONqCTxrcixShtWI="


def test_verify_valid_signature(api, oracle, golden):  # crypto_tests.rs:50-61
    pk, sk = keys(golden)[-1]
    digest = api.Digest.of(b"Hello, world!")
    assert digest.b.hex() == golden["reference"]["hello_digest"]
    sig = api.Signature(oracle.sign(sk, digest.b))
    assert sig.flatten().hex() == golden["reference"]["hello_sig_key3"]
    sig.verify(digest, api.PublicKey(pk))


def test_verify_invalid_signature(api, oracle, golden):  # crypto_tests.rs:64-77
    pk, sk = keys(golden)[-1]
    sig = api.Signature(oracle.sign(sk, api.Digest.of(b"Hello, world!").b))
    with pytest.raises(api.CryptoError):
        sig.verify(api.Digest.of(b"Bad message!"), api.PublicKey(pk))


def test_verify_valid_batch(api, oracle, golden):  # crypto_tests.rs:80-94
    digest = api.Digest.of(b"Hello, world!")
    ks = keys(golden)
    votes = [(api.PublicKey(pk), api.Signature(oracle.sign(sk, digest.b))) for pk, sk in (ks[3], ks[2], ks[1])]
    api.Signature.verify_batch(digest, votes)


def test_verify_invalid_batch(api, oracle, golden):  # crypto_tests.rs:97-115
    digest = api.Digest.of(b"Hello, world!")
    ks = keys(golden)
    votes = [(api.PublicKey(pk), api.Signature(oracle.sign(sk, digest.b))) for pk, sk in (ks[3], ks[2])]
    votes.append((api.PublicKey(ks[1][0]), api.Signature()))  # Signature::default()
    with pytest.raises(api.CryptoError):
        api.Signature.verify_batch(digest, votes)


def test_qc_fixture_verifies(api, golden):  # consensus/src/tests/common.rs:129-144 + messages_tests.rs:8-10
    r = golden["reference"]
    qc_digest = api.Digest.of(bytes(32) + (1).to_bytes(8, "little"))  # QC::digest, messages.rs:201-208
    assert qc_digest.b.hex() == r["qc_digest"]
    votes = [(api.PublicKey(bytes.fromhex(v["pk"])), api.Signature(bytes.fromhex(v["sig"]))) for v in r["qc_votes"]]
    api.Signature.verify_batch(qc_digest, votes)


def test_mempool_batch_digest(api, golden):  # mempool/src/tests/processor_tests.rs:8-38
    r = golden["reference"]
    assert api.Digest.of(bytes.fromhex(r["serialized_batch"])).b.hex() == r["batch_digest"]
