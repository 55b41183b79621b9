"""TEST ONLY: bincode 1.3 serializer for the reference's ConsensusMessage (consensus/src/consensus.rs:33-39) from the Python mirror
objects — used to re-serialise the reference's fixtures (consensus/src/tests/common.rs) for the ingest tests."""
import base64


def _u32(x):
    return int(x).to_bytes(4, "little")


def _u64(x):
    return int(x).to_bytes(8, "little")


def pk(p):                       # Serialize for PublicKey: serialize_str(base64) (crypto/src/lib.rs:94-101)
    s = base64.b64encode(p.b)
    return _u64(len(s)) + s


def qc(q):
    return q.hash.b + _u64(q.round) + _u64(len(q.votes)) + b"".join(pk(n) + s.flatten() for n, s in q.votes)


def tc(t):
    return _u64(t.round) + _u64(len(t.votes)) + b"".join(pk(n) + s.flatten() + _u64(r) for n, s, r in t.votes)


def block(b):
    return (qc(b.qc) + (b"\x00" if b.tc is None else b"\x01" + tc(b.tc)) + pk(b.author) + _u64(b.round) + _u64(len(b.payload))
            + b"".join(d.b for d in b.payload) + b.signature.flatten())


def propose(b):
    return _u32(0) + block(b)


def vote(v):
    return _u32(1) + v.hash.b + _u64(v.round) + pk(v.author) + v.signature.flatten()


def timeout(t):
    return _u32(2) + qc(t.high_qc) + _u64(t.round) + pk(t.author) + t.signature.flatten()


def tc_msg(t):
    return _u32(3) + tc(t)


def sync_request(digest, origin):
    return _u32(4) + digest.b + pk(origin)
