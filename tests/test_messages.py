"""Call-site mirror (hotstuff_b200/messages.py) against the reference's own consensus tests:
consensus/src/tests/messages_tests.rs:8-53 (verify_valid_qc, verify_qc_authority_reuse, verify_qc_unknown_authority,
verify_qc_insufficient_stake) with the committee()/qc() fixtures of consensus/src/tests/common.rs:23-36,129-144."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fx(engine, golden):
    from hotstuff_b200 import crypto, messages
    crypto.set_default_engine(engine)
    r = golden["reference"]
    pks = [crypto.PublicKey(bytes.fromhex(p)) for p in r["pks"]]
    committee = messages.Committee({p: 1 for p in pks})                       # stake 1 each -> quorum 3
    votes = [(crypto.PublicKey(bytes.fromhex(v["pk"])), crypto.Signature(bytes.fromhex(v["sig"]))) for v in r["qc_votes"]]
    return crypto, messages, committee, votes, pks


def test_verify_valid_qc(fx):                                                  # messages_tests.rs:8-10
    crypto, messages, committee, votes, _ = fx
    assert committee.quorum_threshold() == 3
    messages.QC(crypto.Digest(), 1, votes).verify(committee)


def test_verify_qc_authority_reuse(fx):                                        # messages_tests.rs:13-24
    crypto, messages, committee, votes, _ = fx
    v = list(votes)
    v[1] = v[0]
    with pytest.raises(messages.ConsensusError, match="AuthorityReuse"):
        messages.QC(crypto.Digest(), 1, v).verify(committee)


def test_verify_qc_unknown_authority(fx):                                      # messages_tests.rs:27-39
    crypto, messages, committee, votes, _ = fx
    v = list(votes)
    v[0] = (crypto.PublicKey(bytes(range(32))), v[0][1])
    with pytest.raises(messages.ConsensusError, match="UnknownAuthority"):
        messages.QC(crypto.Digest(), 1, v).verify(committee)


def test_verify_qc_insufficient_stake(fx):                                     # messages_tests.rs:42-53
    crypto, messages, committee, votes, _ = fx
    with pytest.raises(messages.ConsensusError, match="QCRequiresQuorum"):
        messages.QC(crypto.Digest(), 1, votes[:2]).verify(committee)


def test_qc_with_bad_signature_and_batched_qcs(fx):
    crypto, messages, committee, votes, _ = fx
    bad = list(votes)
    bad[2] = (bad[2][0], crypto.Signature())
    with pytest.raises(messages.ConsensusError, match="InvalidSignature"):
        messages.QC(crypto.Digest(), 1, bad).verify(committee)
    qcs = [messages.QC(crypto.Digest(), 1, votes), messages.QC(crypto.Digest(), 1, bad), messages.QC(crypto.Digest(), 2, votes),
           messages.QC(crypto.Digest(), 1, votes[:2])]
    assert messages.verify_qcs(qcs, committee) == [True, False, False, False]


def test_tc_verify(fx, oracle, golden):                                        # messages.rs:290-315
    crypto, messages, committee, _, pks = fx
    seeds = [bytes.fromhex(s) for s in golden["reference"]["seeds"]]
    votes = []
    for i, hq in ((0, 3), (1, 5), (2, 4)):
        d = oracle.digest32(messages.timeout_preimage(7, hq))
        votes.append((pks[i], crypto.Signature(oracle.sign(seeds[i], d)), hq))
    messages.TC(7, votes).verify(committee)
    votes[1] = (votes[1][0], votes[1][1], 6)                                   # wrong high_qc_round for that signature
    with pytest.raises(messages.ConsensusError, match="InvalidSignature"):
        messages.TC(7, votes).verify(committee)
