"""Randomised differential test of the receiver path's HOST logic (CPU): a thousand structurally mutated consensus messages
(duplicated / foreign / missing / corrupted votes, foreign authors, corrupted signatures, rounds changed after signing, genesis and
non-genesis certificates, blocks with and without a TC) are judged three ways —
  (1) a line-by-line sequential restatement of Block / Vote / Timeout / TC / QC::verify (consensus/src/messages.rs:54-76,136-146,
      180-198,250-265,290-315) written here, one oracle verify per signature,
  (2) hotstuff_b200/wire.py::verify_frames on the bincode frames (ingest + pre-checks + ONE grouped pass),
  (3) hs::verify_frames (C++, include/hs_consensus.hpp) on the same frames —
and all three must name the same first error for every message.  The engine is the oracle-backed stub in all three (test only)."""
import copy
import os
import struct
import subprocess

import numpy as np

import bincode_ref as bc
import messages_scenarios as sc
from hotstuff_b200 import crypto, messages

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class SequentialReference:
    """messages.rs restated statement by statement; `strict` = Signature::verify, `batch` = Signature::verify_batch."""

    def __init__(self, oracle, committee):
        self.o, self.c = oracle, committee

    def strict(self, digest, pk, sig):
        rec = np.frombuffer(sig.flatten() + pk.b + digest, dtype=np.uint8).reshape(1, 128)
        return bool(self.o.verify_rec128(rec, mode=0)[0])

    def batch(self, digest, votes):
        recs = np.frombuffer(b"".join(s.flatten() + p.b + digest for p, s in votes), dtype=np.uint8).reshape(-1, 128)
        return bool(self.o.verify_rec128(recs, mode=1).all())

    def qc(self, q):                                     # messages.rs:180-198
        weight, used = 0, set()
        for name, _ in q.votes:
            if name.b in used:
                return "AuthorityReuse"
            if self.c.stake(name) <= 0:
                return "UnknownAuthority"
            used.add(name.b)
            weight += self.c.stake(name)
        if weight < self.c.quorum_threshold():
            return "QCRequiresQuorum"
        return None if self.batch(self.o.digest32(messages.vote_preimage(q.hash, q.round)), q.votes) else "InvalidSignature"

    def tc(self, t):                                     # messages.rs:290-315
        weight, used = 0, set()
        for name, _, _ in t.votes:
            if name.b in used:
                return "AuthorityReuse"
            if self.c.stake(name) <= 0:
                return "UnknownAuthority"
            used.add(name.b)
            weight += self.c.stake(name)
        if weight < self.c.quorum_threshold():
            return "TCRequiresQuorum"
        for name, sig, hq in t.votes:
            if not self.strict(self.o.digest32(messages.timeout_preimage(t.round, hq)), name, sig):
                return "InvalidSignature"
        return None

    def block(self, b):                                  # messages.rs:54-76
        if self.c.stake(b.author) <= 0:
            return "UnknownAuthority"
        if not self.strict(self.o.digest32(b.preimage()), b.author, b.signature):
            return "InvalidSignature"
        if b.qc != messages.QC.genesis():
            e = self.qc(b.qc)
            if e:
                return e
        return self.tc(b.tc) if b.tc is not None else None

    def vote(self, v):                                   # messages.rs:136-146
        if self.c.stake(v.author) <= 0:
            return "UnknownAuthority"
        return None if self.strict(self.o.digest32(messages.vote_preimage(v.hash, v.round)), v.author, v.signature) else "InvalidSignature"

    def timeout(self, t):                                # messages.rs:250-265
        if self.c.stake(t.author) <= 0:
            return "UnknownAuthority"
        if not self.strict(self.o.digest32(messages.timeout_preimage(t.round, t.high_qc.round)), t.author, t.signature):
            return "InvalidSignature"
        return self.qc(t.high_qc) if t.high_qc != messages.QC.genesis() else None


def _flip(sig, rng):
    b = bytearray(sig.flatten())
    b[int(rng.integers(0, 64))] ^= 1 << int(rng.integers(0, 8))
    return crypto.Signature(bytes(b))


def _mutate_qc(q, rng, outsider):
    k = int(rng.integers(0, 6))
    if not q.votes:
        return
    i = int(rng.integers(0, len(q.votes)))
    if k == 0:
        q.votes[i] = q.votes[(i + 1) % len(q.votes)]                 # AuthorityReuse
    elif k == 1:
        q.votes[i] = (outsider, q.votes[i][1])                       # UnknownAuthority
    elif k == 2:
        del q.votes[2:]                                              # QCRequiresQuorum
    elif k == 3:
        q.votes[i] = (q.votes[i][0], _flip(q.votes[i][1], rng))      # InvalidSignature
    elif k == 4:
        q.round += 1                                                 # every vote signed another digest
    # k == 5: leave it valid


def _mutate_tc(t, rng, outsider):
    k = int(rng.integers(0, 6))
    i = int(rng.integers(0, len(t.votes)))
    n, s, r = t.votes[i]
    if k == 0:
        t.votes[i] = t.votes[(i + 1) % len(t.votes)]
    elif k == 1:
        t.votes[i] = (outsider, s, r)
    elif k == 2:
        del t.votes[2:]
    elif k == 3:
        t.votes[i] = (n, _flip(s, rng), r)
    elif k == 4:
        t.votes[i] = (n, s, r + 1)                                   # signed (round, r), carries r + 1


def _random_messages(fx, rng, count):
    outsider = crypto.PublicKey(fx.o.keygen(bytes([7]) * 32))
    chain = fx.chain(4)
    out = []
    for _ in range(count):
        kind = int(rng.integers(0, 4))
        author = int(rng.integers(0, 4))
        if kind == 0:
            qc = copy.deepcopy(chain[int(rng.integers(1, 4))].qc) if rng.random() < 0.8 else messages.QC.genesis()
            tc = fx.tc(int(rng.integers(2, 50))) if rng.random() < 0.5 else None
            b = fx.block(author, int(rng.integers(1, 100)), qc=qc, tc=tc, payload=[fx.d(bytes([j])) for j in range(int(rng.integers(0, 3)))])
            m = int(rng.integers(0, 8))
            if m == 0:
                b.signature = _flip(b.signature, rng)
            elif m == 1:
                b.author = outsider
            elif m == 2:
                b.round += 1
            if rng.random() < 0.6 and qc.votes:
                _mutate_qc(b.qc, rng, outsider)
            if tc is not None and rng.random() < 0.6:
                _mutate_tc(b.tc, rng, outsider)
            out.append(("block", b, bc.propose(b)))
        elif kind == 1:
            h = fx.d(bytes(rng.integers(0, 256, 8, dtype=np.uint8)))
            r = int(rng.integers(1, 100))
            v = messages.Vote(h, r, fx.pks[author], fx.sign(author, fx.d(messages.vote_preimage(h, r))))
            m = int(rng.integers(0, 5))
            if m == 0:
                v.signature = _flip(v.signature, rng)
            elif m == 1:
                v.author = outsider
            elif m == 2:
                v.round += 1
            out.append(("vote", v, bc.vote(v)))
        elif kind == 2:
            hq = copy.deepcopy(chain[int(rng.integers(1, 4))].qc) if rng.random() < 0.8 else messages.QC.genesis()
            t = fx.timeout(author, int(rng.integers(5, 100)), hq)
            m = int(rng.integers(0, 6))
            if m == 0:
                t.signature = _flip(t.signature, rng)
            elif m == 1:
                t.author = outsider
            elif m == 2:
                t.round += 1
            if rng.random() < 0.6 and hq.votes:
                _mutate_qc(t.high_qc, rng, outsider)
            out.append(("timeout", t, bc.timeout(t)))
        else:
            t = fx.tc(int(rng.integers(2, 100)))
            if rng.random() < 0.7:
                _mutate_tc(t, rng, outsider)
            out.append(("tc", t, bc.tc_msg(t)))
    return out


def test_three_implementations_name_the_same_first_error(oracle, golden, tmp_path):
    from hotstuff_b200 import build, wire
    fx = sc.Fixtures(oracle, golden, sc.OracleStubEngine(oracle))
    rng = np.random.default_rng(2024)
    msgs = _random_messages(fx, rng, 1000)
    ref = SequentialReference(oracle, fx.committee)
    want = [getattr(ref, kind)(m) for kind, m, _ in msgs]
    frames = [f for _, _, f in msgs]
    got_py = wire.verify_frames(frames, fx.committee, fx.e)
    diff = [(i, msgs[i][0], want[i], got_py[i]) for i in range(len(msgs)) if want[i] != got_py[i]]
    assert not diff, diff[:5]
    # every outcome is well represented (the mutations reach each error path)
    from collections import Counter
    hist = Counter(want)
    for name in (None, "InvalidSignature", "AuthorityReuse", "UnknownAuthority", "QCRequiresQuorum", "TCRequiresQuorum"):
        assert hist[name] >= 12, hist
    # C++
    blob = struct.pack("<I", len(fx.committee.stakes))
    for k, st in fx.committee.stakes.items():
        blob += k + struct.pack("<I", st)
    blob += struct.pack("<I", len(frames)) + b"".join(struct.pack("<I", len(f)) + f for f in frames)
    path = tmp_path / "frames.bin"
    path.write_bytes(blob)
    lib, olib = build.build_engine(), build.build_oracle()
    exe = os.path.join(ROOT, "tests", "cpp", "frames_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "frames_host_test.cpp"), lib, olib,
                           "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.dirname(olib)])
    out = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got_cpp = [None if s == "OK" else s for s in out.stdout.split("\n")[:-1]]
    diff = [(i, msgs[i][0], want[i], got_cpp[i]) for i in range(len(msgs)) if want[i] != got_cpp[i]]
    assert len(got_cpp) == len(want) and not diff, diff[:5]
    # the production wrapper hs::verify_frames(Engine, ...) -> hs_verify_groups, with the engine's entry points answered by the oracle stub
    exe2 = os.path.join(ROOT, "tests", "cpp", "frames_engine_path_stub")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-DHS_TEST_ENGINE_PATH", "-o", exe2, os.path.join(ROOT, "tests", "cpp", "frames_host_test.cpp"),
                           os.path.join(ROOT, "tests", "cpp", "stub_abi.cpp"), os.path.join(ROOT, "hotstuff_b200", "csrc", "hs_ingest.cpp"), olib,
                           "-Wl,-rpath," + os.path.dirname(olib)])
    out = subprocess.run([exe2, str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert [None if s == "OK" else s for s in out.stdout.split("\n")[:-1]] == want
    # the struct-level batch front end (hs::verify_qcs_with: CommitteeIndex pre-checks + one grouped call) on every embedded certificate
    out = subprocess.run([exe, "--qcs", str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got_qc = out.stdout.split("\n")[:-1]
    want_qc = []
    for kind, m, _ in msgs:
        q = m.qc if kind == "block" else m.high_qc if kind == "timeout" else None
        # (a certificate that equals genesis — hash and round zero — is never verified upstream and carries no items)
        want_qc.append("-" if q is None or not q.votes or q == messages.QC.genesis() else ("OK" if ref.qc(q) is None else "BAD"))
    assert got_qc == want_qc, [(i, want_qc[i], got_qc[i]) for i in range(len(msgs)) if want_qc[i] != got_qc[i]][:5]
    assert want_qc.count("OK") > 100 and want_qc.count("BAD") > 100


def test_timeout_bursts_with_the_verified_qc_cache_match_the_sequential_restatement(oracle, golden):
    """View-change bursts (core.rs:227): many timeouts carrying the same high_qc — or a tampered copy of it with the same (hash, round) —
    through messages.verify_timeouts with ONE VerifiedQcCache across bursts.  The cache may only ever skip work, never change a verdict:
    every timeout gets the error the sequential restatement names, in every burst, in any order of valid and tampered certificates."""
    fx = sc.Fixtures(oracle, golden, sc.OracleStubEngine(oracle))
    rng = np.random.default_rng(77)
    ref = SequentialReference(oracle, fx.committee)
    outsider = crypto.PublicKey(oracle.keygen(bytes([9]) * 32))
    chain = fx.chain(4)
    cache = messages.VerifiedQcCache(capacity=8)
    total, hits_before = 0, 0
    for burst in range(12):
        ts = []
        for _ in range(int(rng.integers(3, 40))):
            hq = copy.deepcopy(chain[int(rng.integers(1, 4))].qc) if rng.random() < 0.9 else messages.QC.genesis()
            t = fx.timeout(int(rng.integers(0, 4)), 50 + burst, hq)
            if rng.random() < 0.35 and hq.votes:
                _mutate_qc(t.high_qc, rng, outsider)          # same (hash, round) unless the round mutation hit; different bytes
            m = int(rng.integers(0, 8))
            if m == 0:
                t.signature = _flip(t.signature, rng)
            elif m == 1:
                t.author = outsider
            ts.append(t)
        want = [ref.timeout(t) for t in ts]
        got = messages.verify_timeouts(ts, fx.committee, fx.e, qc_cache=cache)
        assert got == want, [(i, want[i], got[i]) for i in range(len(ts)) if want[i] != got[i]][:5]
        total += len(ts)
    assert cache.hits > 20 and total > 100      # the cache did skip repeated certificates
