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

    def __init__(self, hash_, round_, votes):
        self.hash, self.round, self.votes = hash_, int(round_), list(votes)

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

    def verify(self, committee, engine=None):
        weight, used = 0, set()
        for name, _, _ in self.votes:                       # messages.rs:292-304
            if name.b in used:
                raise ConsensusError("AuthorityReuse")
            stake = committee.stake(name)
            if stake <= 0:
                raise ConsensusError("UnknownAuthority")
            used.add(name.b)
            weight += stake
        if weight < committee.quorum_threshold():
            raise ConsensusError("TCRequiresQuorum")
        e = engine or default_engine()
        # messages.rs:307-313: one 16-byte digest and one strict verify per vote -> one batched GPU call here
        digests = digest_many([timeout_preimage(self.round, hq) for _, _, hq in self.votes], e)
        recs = np.frombuffer(b"".join(sig.flatten() + name.b + d.b for (name, sig, _), d in zip(self.votes, digests)), dtype=np.uint8)
        if not e.verify_strict_batch(recs.reshape(-1, 128)).all():
            raise ConsensusError("InvalidSignature")


def verify_qcs(qcs, committee, engine=None):
    """Many QCs in one engine pass (the view-change burst of SURVEY §3D: every received Timeout carries a high_qc).
    Host keeps the duplicate / stake checks; digests, signature checks and the per-QC AND run on the GPU.
    Returns a list of booleans (True = the QC verifies), never raising for crypto failures."""
    e = engine or default_engine()
    ok = []
    for qc in qcs:
        try:
            qc.check_quorum(committee)
            ok.append(True)
        except ConsensusError:
            ok.append(False)
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
