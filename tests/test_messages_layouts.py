"""Digest preimage layouts of consensus/src/messages.rs (CPU-only: checked with the oracle's SHA-512 against the fixtures derived from
the reference's own tests)."""
from hotstuff_b200.crypto import Digest, PublicKey
from hotstuff_b200 import messages


def test_qc_and_vote_preimage_is_hash_then_round_le(oracle, golden):      # messages.rs:149-156, 201-208
    pre = messages.vote_preimage(Digest(), 1)
    assert pre == bytes(32) + (1).to_bytes(8, "little") and len(pre) == 40
    assert oracle.digest32(pre).hex() == golden["reference"]["qc_digest"]


def test_timeout_and_block_preimages(oracle):                               # messages.rs:268-275, 79-90
    assert messages.timeout_preimage(7, 3) == (7).to_bytes(8, "little") + (3).to_bytes(8, "little")
    author, payload, qch = PublicKey(bytes(range(32))), [Digest(bytes([1] * 32)), Digest(bytes([2] * 32))], Digest(bytes([9] * 32))
    pre = messages.block_preimage(author, 5, payload, qch)
    assert pre == bytes(range(32)) + (5).to_bytes(8, "little") + bytes([1] * 32) + bytes([2] * 32) + bytes([9] * 32)


def test_committee_quorum_threshold():                                      # consensus/src/config.rs:63-72
    c = messages.Committee({PublicKey(bytes([i] * 32)): 1 for i in range(4)})
    assert c.quorum_threshold() == 3 and c.stake(PublicKey(bytes([9] * 32))) == 0
    assert messages.Committee({PublicKey(bytes([i % 256, i // 256] + [0] * 30)): 1 for i in range(1000)}).quorum_threshold() == 667
