"""Call-site mirror of consensus/src/messages.rs for the crypto path: the digest layouts that feed Signature::verify* and
the verify() pre-checks that decide WHAT reaches the engine ("next" row f1).

Only the crypto-relevant parts are mirrored (digests, stake/duplicate checks, which verify entry point is called);
block storage, networking and the protocol state machine are out of scope.
"""
import numpy as np

from .crypto import CryptoError, Digest, PublicKey, Signature, default_engine, digest_many


class ConsensusError(Exception):
    """consensus/src/error.rs: AuthorityReuse / UnknownAuthority / QCRequiresQuorum / TCRequiresQuorum / InvalidSignature."""


def _le64(x):
    return int(x).to_bytes(8, "little")


# ---- digest layouts (all SHA-512[..32]) -------------------------------------------------------------------------
def block_preimage(author, round_, payload, qc_hash):     # messages.rs:79-90
    return author.b + _le64(round_) + b"".join(d.b for d in payload) + qc_hash.b


def vote_preimage(hash_, round_):                          # messages.rs:149-156 (identical to QC::digest, :201-208)
    return hash_.b + _le64(round_)


def timeout_preimage(round_, high_qc_round):               # messages.rs:268-275 and the per-vote digest of TC::verify :307-311
    return _le64(round_) + _le64(high_qc_round)


class Committee:
    """consensus/src/config.rs:28-72: name -> stake; quorum_threshold = 2 * total / 3 + 1."""

    def __init__(self, stakes):
        self.stakes = {k.b if isinstance(k, PublicKey) else bytes(k): int(v) for k, v in stakes.items()}

    def stake(self, name):
        return self.stakes.get(name.b, 0)

    def quorum_threshold(self):
        return 2 * sum(self.stakes.values()) // 3 + 1


class QC:
    """consensus/src/messages.rs:165-208."""

    def __init__(self, hash_=None, round_=0, votes=()):
        self.hash, self.round, self.votes = (hash_ if hash_ is not None else Digest()), int(round_), list(votes)

    @staticmethod
    def genesis():                                          # messages.rs:172-174 = QC::default()
        return QC()

    def __eq__(self, other):                                # messages.rs:216-220: hash and round only
        return isinstance(other, QC) and self.hash == other.hash and self.round == other.round

    def __hash__(self):
        return hash((self.hash.b, self.round))

    def wire_bytes(self):
        """Everything verify() looks at, in a fixed layout — the exact-match key of VerifiedQcCache."""
        return self.hash.b + _le64(self.round) + b"".join(name.b + sig.flatten() for name, sig in self.votes)

    def digest(self, engine=None):
        return Digest.of(vote_preimage(self.hash, self.round), engine)

    def check_quorum(self, committee):                      # messages.rs:182-194 — runs BEFORE any crypto
        weight, used = 0, set()
        for name, _ in self.votes:
            if name.b in used:
                raise ConsensusError("AuthorityReuse")
            stake = committee.stake(name)
            if stake <= 0:
                raise ConsensusError("UnknownAuthority")
            used.add(name.b)
            weight += stake
        if weight < committee.quorum_threshold():
            raise ConsensusError("QCRequiresQuorum")

    def verify(self, committee, engine=None):               # messages.rs:180-198
        self.check_quorum(committee)
        try:
            Signature.verify_batch(self.digest(engine), self.votes, engine)   # messages.rs:197
        except CryptoError as e:
            raise ConsensusError("InvalidSignature") from e


class TC:
    """consensus/src/messages.rs:283-315: votes = (author, signature, high_qc_round); verified INDIVIDUALLY upstream."""

    def __init__(self, round_, votes):
        self.round, self.votes = int(round_), list(votes)

    def check_quorum(self, committee):                      # messages.rs:292-304 — runs BEFORE any crypto
        weight, used = 0, set()
        for name, _, _ in self.votes:
            if name.b in used:
                raise ConsensusError("AuthorityReuse")
            stake = committee.stake(name)
            if stake <= 0:
                raise ConsensusError("UnknownAuthority")
            used.add(name.b)
            weight += stake
        if weight < committee.quorum_threshold():
            raise ConsensusError("TCRequiresQuorum")

    def verify(self, committee, engine=None):
        self.check_quorum(committee)
        if not verify_tcs([self], committee, engine, prechecked=True)[0]:   # messages.rs:307-313 as ONE engine call
            raise ConsensusError("InvalidSignature")

    def high_qc_rounds(self):                               # messages.rs:317-319
        return [r for _, _, r in self.votes]


class Vote:
    """consensus/src/messages.rs:112-156."""

    def __init__(self, hash_, round_, author, signature):
        self.hash, self.round, self.author, self.signature = hash_, int(round_), author, signature

    def digest(self, engine=None):
        return Digest.of(vote_preimage(self.hash, self.round), engine)

    def verify(self, committee, engine=None):               # messages.rs:136-146
        if committee.stake(self.author) <= 0:
            raise ConsensusError("UnknownAuthority")
        try:
            self.signature.verify(self.digest(engine), self.author, engine)
        except CryptoError as e:
            raise ConsensusError("InvalidSignature") from e


class Timeout:
    """consensus/src/messages.rs:223-275."""

    def __init__(self, high_qc, round_, author, signature):
        self.high_qc, self.round, self.author, self.signature = high_qc, int(round_), author, signature

    def digest(self, engine=None):
        return Digest.of(timeout_preimage(self.round, self.high_qc.round), engine)

    def verify(self, committee, engine=None, qc_cache=None):   # messages.rs:250-265
        if committee.stake(self.author) <= 0:
            raise ConsensusError("UnknownAuthority")
        err = verify_timeouts([self], committee, engine, qc_cache=qc_cache, prechecked=True)[0]
        if err:
            raise ConsensusError(err)


class Block:
    """consensus/src/messages.rs:17-90 (crypto-relevant fields)."""

    def __init__(self, qc, tc, author, round_, payload, signature):
        self.qc, self.tc, self.author, self.round, self.payload, self.signature = qc, tc, author, int(round_), list(payload), signature

    def preimage(self):
        return block_preimage(self.author, self.round, self.payload, self.qc.hash)

    def digest(self, engine=None):
        return Digest.of(self.preimage(), engine)

    def verify(self, committee, engine=None):               # messages.rs:54-76
        err = verify_blocks([self], committee, engine)[0]
        if err:
            raise ConsensusError(err)


class VerifiedQcCache:
    """Verified-QC cache: during a view change every Timeout carries its sender's high_qc, so a node re-verifies
    the same certificate up to N times (consensus/src/core.rs:227 -> Timeout::verify -> QC::verify, the O(N^2) path).
    A hit requires the candidate's bytes — (hash, round) AND every (name, signature) — to equal a certificate that already
    verified, so the verdict is exactly the reference's (a different vote set for the same (hash, round) is verified afresh);
    rejected certificates are not cached.  Keyed by (hash, round), bounded LRU."""

    def __init__(self, capacity=1024):
        from collections import OrderedDict
        self.capacity, self.map, self.hits, self.misses = capacity, OrderedDict(), 0, 0

    def known_valid(self, qc):
        got = self.map.get((qc.hash.b, qc.round))
        if got is not None and got == qc.wire_bytes():
            self.map.move_to_end((qc.hash.b, qc.round))
            self.hits += 1
            return True
        self.misses += 1
        return False

    def remember(self, qc):
        self.map[(qc.hash.b, qc.round)] = qc.wire_bytes()
        self.map.move_to_end((qc.hash.b, qc.round))
        while len(self.map) > self.capacity:
            self.map.popitem(last=False)


def verify_qcs(qcs, committee, engine=None, qc_cache=None):
    """Many QCs in one engine pass (the view-change burst of SURVEY §3D: every received Timeout carries a high_qc).
    Host keeps the duplicate / stake checks; digests, signature checks and the per-QC AND run on the GPU.
    Returns a list of booleans (True = the QC verifies), never raising for crypto failures.  With a VerifiedQcCache,
    certificates byte-identical to one that already verified skip the engine, as do duplicates inside this call."""
    e = engine or default_engine()
    ok = []
    for qc in qcs:
        try:
            qc.check_quorum(committee)
            ok.append(True)
        except ConsensusError:
            ok.append(False)
    if qc_cache is not None:
        first, alias, todo = {}, {}, []
        for j, qc in enumerate(qcs):
            if not ok[j] or qc_cache.known_valid(qc):
                continue
            wb = qc.wire_bytes()
            if wb in first:
                alias[j] = first[wb]
            else:
                first[wb] = j
                todo.append(j)
        if todo:
            res = verify_qcs([qcs[j] for j in todo], committee, e)
            for j, r in zip(todo, res):
                ok[j] = r
                if r:
                    qc_cache.remember(qcs[j])
        for j, k in alias.items():
            ok[j] = ok[k]
        return ok
    pre, sig, pk, midx = [], [], [], []
    live = [j for j, o in enumerate(ok) if o]
    for new_j, j in enumerate(live):
        qc = qcs[j]
        pre.append(vote_preimage(qc.hash, qc.round))
        for name, s_ in qc.votes:
            sig.append(s_.flatten())
            pk.append(name.b)
            midx.append(new_j)
    if not live:
        return ok
    good = e.verify_qcs(np.frombuffer(b"".join(pre), np.uint8), np.frombuffer(b"".join(sig), np.uint8), np.asarray(midx, dtype=np.uint32),
                        pk=np.frombuffer(b"".join(pk), np.uint8))   # digests, votes and the per-QC AND all on the GPU
    for new_j, j in enumerate(live):
        ok[j] = bool(good[new_j])
    return ok


def verify_tcs(tcs, committee, engine=None, prechecked=False):
    """TC::verify (messages.rs:290-315) for many TCs in one engine call: the host runs the stake / duplicate checks, the GPU builds
    the n 16-byte digests (they differ only in high_qc_round), verifies every vote with Signature::verify and ANDs per TC."""
    e = engine or default_engine()
    ok = [True] * len(tcs)
    if not prechecked:
        for j, tc in enumerate(tcs):
            try:
                tc.check_quorum(committee)
            except ConsensusError:
                ok[j] = False
    live = [j for j, o in enumerate(ok) if o]
    if not live:
        return ok
    rounds, sig, pk, hq, ti = [], [], [], [], []
    for new_j, j in enumerate(live):
        rounds.append(tcs[j].round)
        for name, s_, r in tcs[j].votes:
            sig.append(s_.flatten())
            pk.append(name.b)
            hq.append(r)
            ti.append(new_j)
    good = e.verify_tcs(np.asarray(rounds, dtype=np.uint64), np.frombuffer(b"".join(sig), np.uint8), np.asarray(hq, dtype=np.uint64),
                        tc_idx=np.asarray(ti, dtype=np.uint32), pk=np.frombuffer(b"".join(pk), np.uint8))
    for new_j, j in enumerate(live):
        ok[j] = bool(good[new_j])
    return ok


def verify_timeouts(timeouts, committee, engine=None, qc_cache=None, prechecked=False):
    """Timeout::verify (messages.rs:250-265) for a burst of timeouts (core.rs:227, one per validator during a view change):
    the n timeout signatures in one hs_verify_tcs call (digest = SHA-512(round || high_qc.round)[..32] built on the GPU) and the
    embedded high_qcs — mostly the SAME certificate n times — through verify_qcs with the exact-match cache.
    Returns a list with None (valid) or the ConsensusError name the reference would raise first."""
    e = engine or default_engine()
    out = [None] * len(timeouts)
    for j, t in enumerate(timeouts):
        if not prechecked and committee.stake(t.author) <= 0:
            out[j] = "UnknownAuthority"
    live = [j for j, o in enumerate(out) if o is None]
    if not live:
        return out
    good = e.verify_tcs(np.asarray([timeouts[j].round for j in live], dtype=np.uint64),
                        np.frombuffer(b"".join(timeouts[j].signature.flatten() for j in live), np.uint8),
                        np.asarray([timeouts[j].high_qc.round for j in live], dtype=np.uint64),
                        pk=np.frombuffer(b"".join(timeouts[j].author.b for j in live), np.uint8))
    for j, g in zip(live, good):
        if not g:
            out[j] = "InvalidSignature"
    genesis = QC.genesis()
    need = [j for j in live if out[j] is None and timeouts[j].high_qc != genesis]
    for j in need:                                           # QC pre-checks raise their own error names (messages.rs:182-194)
        try:
            timeouts[j].high_qc.check_quorum(committee)
        except ConsensusError as ex:
            out[j] = str(ex)
    need = [j for j in need if out[j] is None]
    if need:
        res = verify_qcs([timeouts[j].high_qc for j in need], committee, e, qc_cache=qc_cache if qc_cache is not None else VerifiedQcCache())
        for j, r in zip(need, res):
            if not r:
                out[j] = "InvalidSignature"
    return out


def verify_blocks(blocks, committee, engine=None):
    """Block::verify (messages.rs:54-76) for many blocks in ONE engine pass (hs_verify_groups): per block the author signature
    over Block::digest (strict), the embedded QC's votes over QC::digest (verify_batch condition) unless it is the genesis QC, and
    the embedded TC's votes (strict, 16-byte digests).  All preimages are hashed on the GPU.  Returns a list with None (valid) or
    the name of the error the reference raises FIRST, in its order: author stake, author signature, QC pre-checks, QC
    signatures, TC pre-checks, TC signatures."""
    e = engine or default_engine()
    out = [None] * len(blocks)
    genesis = QC.genesis()
    pre, sig, pk, mi, gi, mode = [], [], [], [], [], []
    plan = []          # per engine group: (block index, author item, qc error | None, qc item range, tc error | None, tc item range)

    def item(sig_b, pk_b, msg, g, m):
        sig.append(sig_b); pk.append(pk_b); mi.append(msg); gi.append(g); mode.append(m)
        return len(sig) - 1

    for j, b in enumerate(blocks):
        if committee.stake(b.author) <= 0:
            out[j] = "UnknownAuthority"
            continue
        g = len(plan)
        pre.append(b.preimage())
        author_item = item(b.signature.flatten(), b.author.b, len(pre) - 1, g, 0)
        qc_err = tc_err = None
        qc_rng = tc_rng = (0, 0)
        if b.qc != genesis:
            try:
                b.qc.check_quorum(committee)
                pre.append(vote_preimage(b.qc.hash, b.qc.round))
                lo = len(sig)
                for name, s_ in b.qc.votes:
                    item(s_.flatten(), name.b, len(pre) - 1, g, 1)
                qc_rng = (lo, len(sig))
            except ConsensusError as ex:
                qc_err = str(ex)
        if b.tc is not None and qc_err is None:
            try:
                b.tc.check_quorum(committee)
                lo = len(sig)
                for name, s_, r in b.tc.votes:
                    pre.append(timeout_preimage(b.tc.round, r))
                    item(s_.flatten(), name.b, len(pre) - 1, g, 0)
                tc_rng = (lo, len(sig))
            except ConsensusError as ex:
                tc_err = str(ex)
        plan.append((j, author_item, qc_err, qc_rng, tc_err, tc_rng))
    if not plan:
        return out
    off = np.zeros(len(pre) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in pre])
    _, items = e.verify_groups(np.frombuffer(b"".join(pre), np.uint8), off, np.frombuffer(b"".join(sig), np.uint8), np.asarray(mi, dtype=np.uint32),
                               np.asarray(gi, dtype=np.uint32), len(plan), mode=np.asarray(mode, dtype=np.uint8),
                               pk=np.frombuffer(b"".join(pk), np.uint8), want_items=True)
    for j, author_item, qc_err, qc_rng, tc_err, tc_rng in plan:
        if not items[author_item]:
            out[j] = "InvalidSignature"
        elif qc_err:
            out[j] = qc_err
        elif not items[qc_rng[0]:qc_rng[1]].all():
            out[j] = "InvalidSignature"
        elif tc_err:
            out[j] = tc_err
        elif not items[tc_rng[0]:tc_rng[1]].all():
            out[j] = "InvalidSignature"
    return out
