"""Shared scenarios for the call-site mirror (hotstuff_b200/messages.py): run on the GPU engine by tests/test_messages.py and on an
oracle-backed stub engine (host logic only: pre-check order, grouping, cache) by tests/test_messages_host.py.
Fixtures follow consensus/src/tests/common.rs: keys() (:17-20), block() (:116-119), vote() (:122-125), qc() (:128-144), chain() (:147-180)."""
import numpy as np

from hotstuff_b200 import crypto, messages


class OracleStubEngine:
    """TEST ONLY: the Engine methods messages.py uses, answered by the CPU oracle (so the host logic runs without a GPU)."""

    def __init__(self, oracle):
        self.o = oracle
        self.calls = {"verify_qcs": 0, "verify_tcs": 0, "verify_groups": 0, "qc_votes": 0}

    def digest32_batch(self, data, off):
        return self.o.digest32_batch(bytes(data) if not isinstance(data, np.ndarray) else data, np.asarray(off, dtype=np.uint64))

    def verify_strict_batch(self, recs):
        return self.o.verify_rec128(recs, mode=0)

    def verify_batch_shared_msg(self, digest, votes, want_bitmap=False):
        ok, bits = self.o.verify_batch_shared_msg(bytes(digest), votes)
        return (ok, bits) if want_bitmap else ok

    def _items(self, sig, pk, digests, mode):
        sig = np.asarray(sig, np.uint8).reshape(-1, 64)
        pk = np.asarray(pk, np.uint8).reshape(-1, 32)
        recs = np.concatenate([sig, pk, digests], axis=1)
        s, e = self.o.verify_rec128(recs, mode=0), self.o.verify_rec128(recs, mode=1)
        return np.where(np.asarray(mode) == 1, e, s)

    def verify_qcs(self, preimages, sig, qc_idx, pk=None, validator_idx=None, want_votes=False):
        self.calls["verify_qcs"] += 1
        pre = np.asarray(preimages, np.uint8).reshape(-1, 40)
        d = self.o.digest32_batch(pre.reshape(-1), np.arange(len(pre) + 1, dtype=np.uint64) * 40)
        qi = np.asarray(qc_idx)
        self.calls["qc_votes"] += len(qi)
        bits = self._items(sig, pk, d[qi], np.ones(len(qi)))
        out = np.ones(len(pre), dtype=bool)
        np.logical_and.at(out, qi, bits)
        return (out, bits) if want_votes else out

    def verify_tcs(self, tc_rounds, sig, high_qc_rounds, tc_idx=None, pk=None, validator_idx=None, want_votes=False):
        self.calls["verify_tcs"] += 1
        n = len(high_qc_rounds)
        ti = np.arange(n) if tc_idx is None else np.asarray(tc_idx)
        pre = b"".join(int(tc_rounds[t]).to_bytes(8, "little") + int(h).to_bytes(8, "little") for t, h in zip(ti, high_qc_rounds))
        d = self.o.digest32_batch(pre, np.arange(n + 1, dtype=np.uint64) * 16)
        bits = self._items(sig, pk, d, np.zeros(n))
        out = np.ones(len(tc_rounds), dtype=bool)
        np.logical_and.at(out, ti, bits)
        return (out, bits) if want_votes else out

    def verify_groups(self, preimages, pre_off, sig, msg_idx, group_idx, n_groups, mode=None, pk=None, validator_idx=None, want_items=False):
        self.calls["verify_groups"] += 1
        d = self.o.digest32_batch(np.asarray(preimages, np.uint8), np.asarray(pre_off, dtype=np.uint64))
        mi, gi = np.asarray(msg_idx), np.asarray(group_idx)
        bits = self._items(sig, pk, d[mi], np.zeros(len(mi)) if mode is None else mode)
        out = np.ones(n_groups, dtype=bool)
        np.logical_and.at(out, gi, bits)
        return (out, bits) if want_items else out


class Fixtures:
    def __init__(self, oracle, golden, engine):
        r = golden["reference"]
        self.o, self.e = oracle, engine
        self.seeds = [bytes.fromhex(s) for s in r["seeds"]]
        self.pks = [crypto.PublicKey(bytes.fromhex(p)) for p in r["pks"]]
        self.committee = messages.Committee({p: 1 for p in self.pks})          # common.rs:23-36: stake 1 each -> quorum 3

    def sign(self, i, digest):
        return crypto.Signature(self.o.sign(self.seeds[i], digest.b if isinstance(digest, crypto.Digest) else digest))

    def d(self, pre):
        return crypto.Digest(self.o.digest32(pre))

    def block(self, i, round_, qc=None, tc=None, payload=()):
        qc = qc or messages.QC.genesis()
        b = messages.Block(qc, tc, self.pks[i], round_, list(payload), crypto.Signature())
        b.signature = self.sign(i, self.d(b.preimage()))
        return b

    def qc_for(self, hash_, round_, signers=(0, 1, 2, 3)):
        dig = self.d(messages.vote_preimage(hash_, round_))
        return messages.QC(hash_, round_, [(self.pks[i], self.sign(i, dig)) for i in signers])

    def chain(self, n=4):                                                      # common.rs:147-180
        latest, out = messages.QC.genesis(), []
        for i in range(n):
            b = self.block(i % 4, 1 + i, qc=latest)
            latest = self.qc_for(self.d(b.preimage()), b.round)
            out.append(b)
        return out

    def tc(self, round_, hqs=((0, 3), (1, 5), (2, 4))):
        return messages.TC(round_, [(self.pks[i], self.sign(i, self.d(messages.timeout_preimage(round_, hq))), hq) for i, hq in hqs])

    def timeout(self, i, round_, high_qc):
        return messages.Timeout(high_qc, round_, self.pks[i], self.sign(i, self.d(messages.timeout_preimage(round_, high_qc.round))))


def scenario_block_vote_timeout(fx):
    e, c = fx.e, fx.committee
    b = fx.block(3, 1)                                                         # common.rs block(): keys().pop(), genesis QC, round 1
    b.verify(c, e)
    v = messages.Vote(fx.d(b.preimage()), 1, fx.pks[3], crypto.Signature())
    v.signature = fx.sign(3, fx.d(messages.vote_preimage(v.hash, v.round)))
    v.verify(c, e)
    v.round = 2
    _raises(lambda: v.verify(c, e), "InvalidSignature")
    outsider = messages.Vote(v.hash, 1, crypto.PublicKey(bytes(range(32))), v.signature)
    _raises(lambda: outsider.verify(c, e), "UnknownAuthority")                 # messages.rs:138-141
    t = fx.timeout(2, 9, fx.qc_for(fx.d(b.preimage()), 1))
    t.verify(c, e)
    t_bad_qc = fx.timeout(2, 9, fx.qc_for(fx.d(b.preimage()), 1))
    t_bad_qc.high_qc.votes[1] = (t_bad_qc.high_qc.votes[1][0], crypto.Signature())
    _raises(lambda: t_bad_qc.verify(c, e), "InvalidSignature")
    t_genesis = fx.timeout(1, 4, messages.QC.genesis())                        # messages.rs:261: genesis high_qc is not verified
    t_genesis.verify(c, e)
    t_short = fx.timeout(1, 9, fx.qc_for(fx.d(b.preimage()), 1, signers=(0, 1)))
    _raises(lambda: t_short.verify(c, e), "QCRequiresQuorum")


def scenario_blocks_batched(fx):
    e, c = fx.e, fx.committee
    chain = fx.chain(4)
    assert messages.verify_blocks(chain, c, e) == [None] * 4
    blocks = list(chain)
    with_tc = fx.block(1, 9, qc=chain[3].qc, tc=fx.tc(8))
    blocks.append(with_tc)
    bad_sig = fx.block(2, 6, qc=chain[2].qc)
    bad_sig.round = 7                                                          # digest no longer matches the signature
    blocks.append(bad_sig)
    bad_qc = fx.block(0, 6, qc=fx.qc_for(fx.d(b"x" * 5), 5))
    bad_qc.qc.votes[2] = (bad_qc.qc.votes[2][0], bad_qc.qc.votes[0][1])
    blocks.append(bad_qc)
    reuse = fx.block(0, 6, qc=fx.qc_for(fx.d(b"y"), 5))
    reuse.qc.votes[1] = reuse.qc.votes[0]
    blocks.append(reuse)
    bad_tc = fx.block(3, 9, qc=chain[3].qc, tc=fx.tc(8))
    bad_tc.tc.votes[0] = (bad_tc.tc.votes[0][0], bad_tc.tc.votes[0][1], 99)
    blocks.append(bad_tc)
    short_tc = fx.block(3, 9, qc=chain[3].qc, tc=fx.tc(8, hqs=((0, 3), (1, 5))))
    blocks.append(short_tc)
    outsider = fx.block(3, 2)
    outsider.author = crypto.PublicKey(bytes(range(32)))
    blocks.append(outsider)
    both = fx.block(2, 6, qc=fx.qc_for(fx.d(b"z"), 5, signers=(0, 1)))          # bad author signature AND a QC without quorum:
    both.round = 8                                                             # the signature is checked first (messages.rs:63-66)
    blocks.append(both)
    got = messages.verify_blocks(blocks, c, e)
    assert got == [None] * 5 + ["InvalidSignature", "InvalidSignature", "AuthorityReuse", "InvalidSignature", "TCRequiresQuorum", "UnknownAuthority",
                               "InvalidSignature"], got
    for b, want in zip(blocks, got):                                           # one at a time gives the same answers
        if want is None:
            b.verify(c, e)
        else:
            _raises(lambda: b.verify(c, e), want)


def scenario_tcs_and_timeout_burst(fx, count_votes=None):
    e, c = fx.e, fx.committee
    good, bad = fx.tc(7), fx.tc(7)
    bad.votes[1] = (bad.votes[1][0], bad.votes[1][1], 6)                        # wrong high_qc_round for that signature
    reuse = fx.tc(7)
    reuse.votes[2] = reuse.votes[0]
    assert messages.verify_tcs([good, bad, fx.tc(11), reuse], c, e) == [True, False, True, False]
    good.verify(c, e)
    _raises(lambda: bad.verify(c, e), "InvalidSignature")
    _raises(lambda: reuse.verify(c, e), "AuthorityReuse")
    # view change: every validator's Timeout carries the SAME high_qc (core.rs:227 -> the O(N^2) path)
    hq = fx.qc_for(fx.d(b"high"), 6)
    tos = [fx.timeout(i, 7, fx.qc_for(fx.d(b"high"), 6)) for i in range(4)]
    forged = fx.timeout(1, 7, fx.qc_for(fx.d(b"high"), 6))
    forged.high_qc.votes[3] = (forged.high_qc.votes[3][0], crypto.Signature())  # same (hash, round), one vote replaced
    tos.append(forged)
    stale = fx.timeout(2, 7, hq)
    stale.round = 8                                                            # signature no longer matches
    tos.append(stale)
    cache = messages.VerifiedQcCache()
    before = count_votes() if count_votes else 0
    got = messages.verify_timeouts(tos, c, e, qc_cache=cache)
    assert got == [None, None, None, None, "InvalidSignature", "InvalidSignature"], got
    if count_votes:                                                            # 4 identical QCs verified once + the forged one: 8 votes, not 20
        assert count_votes() - before == 8
    again = messages.verify_timeouts(tos[:4], c, e, qc_cache=cache)
    assert again == [None] * 4 and cache.hits >= 4
    if count_votes:
        assert count_votes() - before == 8                                     # all cache hits: no QC vote reached the engine again
    assert messages.verify_timeouts([forged], c, e, qc_cache=cache) == ["InvalidSignature"]   # a hit needs identical bytes


def _raises(fn, name):
    try:
        fn()
    except messages.ConsensusError as ex:
        assert str(ex) == name, (str(ex), name)
        return
    raise AssertionError("expected ConsensusError(%s)" % name)
