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


# ---- Block / Vote / Timeout / TC and the batched front end (same scenarios as tests/test_messages_host.py, on the CUDA engine)
def test_block_vote_timeout_verify_gpu(engine, oracle, golden):
    import messages_scenarios as sc
    sc.scenario_block_vote_timeout(sc.Fixtures(oracle, golden, engine))


def test_blocks_batched_error_order_gpu(engine, oracle, golden):
    import messages_scenarios as sc
    sc.scenario_blocks_batched(sc.Fixtures(oracle, golden, engine))


def test_tcs_and_timeout_burst_gpu(engine, oracle, golden):
    import messages_scenarios as sc
    sc.scenario_tcs_and_timeout_burst(sc.Fixtures(oracle, golden, engine))


def test_verify_tcs_large_against_oracle(engine, oracle):
    """hs_verify_tcs at committee scale: 40 TCs x 67 votes, digests built on the GPU from (round, high_qc_round); key bytes and
    validator indices; some votes signed for another high_qc_round."""
    import numpy as np
    rng = np.random.default_rng(31)
    N, T, V = 100, 40, 67
    seeds = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    pks = oracle.keygen_batch(seeds)
    rounds = rng.integers(1, 2**40, T).astype(np.uint64)
    ti = np.repeat(np.arange(T, dtype=np.uint32), V)
    vidx = np.concatenate([rng.choice(N, V, replace=False) for _ in range(T)]).astype(np.uint32)
    hq = rng.integers(0, 2**33, T * V).astype(np.uint64)
    pre = b"".join(int(rounds[t]).to_bytes(8, "little") + int(h).to_bytes(8, "little") for t, h in zip(ti, hq))
    d = oracle.digest32_batch(pre, np.arange(T * V + 1, dtype=np.uint64) * 16)
    sig = oracle.sign_batch(seeds, pks, vidx, d.reshape(-1), np.arange(T * V + 1, dtype=np.uint64) * 32)
    bad = rng.choice(T * V, 25, replace=False)
    hq[bad] += 1
    want_votes = np.ones(T * V, dtype=bool)
    want_votes[bad] = False
    want = np.ones(T, dtype=bool)
    np.logical_and.at(want, ti, want_votes)
    engine.committee_register(np.zeros((0, 32), np.uint8))
    got, gv = engine.verify_tcs(rounds, sig, hq, tc_idx=ti, pk=pks[vidx], want_votes=True)
    assert (gv == want_votes).all() and (got == want).all() and (~want).sum() >= 10
    engine.committee_register(pks)
    assert (engine.verify_tcs(rounds, sig, hq, tc_idx=ti, validator_idx=vidx) == want).all()
    # Timeout shape: one vote per certificate
    got = engine.verify_tcs(rounds[ti], sig, hq, pk=pks[vidx])
    assert (got == want_votes).all()
    engine.committee_register(np.zeros((0, 32), np.uint8))


def test_wire_frames_ingest_and_verify_gpu(engine, oracle, golden):
    """bincode frames -> hs_ingest_consensus_frames -> hs_verify_groups on the GPU == struct-level verification (tests/test_wire_ingest.py)."""
    import messages_scenarios as sc
    from test_wire_ingest import run_verify_frames
    run_verify_frames(sc.Fixtures(oracle, golden, engine))
