"""Wire-format ingest (hs_ingest_consensus_frames): the reference's fixtures (consensus/src/tests/common.rs: block(), vote(), qc(),
chain()) re-serialised with bincode's layout, parsed straight into verify-ready arrays, checked item by item against the Python
mirror, and verified through the same grouped engine call (oracle-backed stub here; the CUDA engine in test_messages.py).  Host-only."""
import numpy as np

import bincode_ref as bc
import messages_scenarios as sc
from hotstuff_b200 import crypto, messages, wire


def _fx(oracle, golden):
    return sc.Fixtures(oracle, golden, sc.OracleStubEngine(oracle))


def _messages(fx):
    chain = fx.chain(4)
    blk_tc = fx.block(1, 9, qc=chain[3].qc, tc=fx.tc(8), payload=[fx.d(b"p1"), fx.d(b"p2")])
    v = messages.Vote(fx.d(chain[0].preimage()), 1, fx.pks[3], crypto.Signature())
    v.signature = fx.sign(3, fx.d(messages.vote_preimage(v.hash, v.round)))
    to = fx.timeout(2, 9, chain[2].qc)
    to_gen = fx.timeout(1, 4, messages.QC.genesis())
    return chain, blk_tc, v, to, to_gen


def test_ingest_matches_the_python_mirror(oracle, golden):
    fx = _fx(oracle, golden)
    chain, blk_tc, v, to, to_gen = _messages(fx)
    frames = [bc.propose(b) for b in chain] + [bc.propose(blk_tc), bc.vote(v), bc.timeout(to), bc.timeout(to_gen), bc.tc_msg(fx.tc(7)),
                                              bc.sync_request(fx.d(b"missing"), fx.pks[0])]
    g = wire.ingest_frames(frames)
    info = g["info"]
    assert list(info["kind"]) == [0, 0, 0, 0, 0, 1, 2, 2, 3, 4]

    def pre(i):
        m = g["msg_idx"][i]
        return g["preimages"][int(g["pre_off"][m]):int(g["pre_off"][m + 1])].tobytes()

    # block 0: genesis QC -> author item only
    f = info[0]
    assert f["qc_is_genesis"] and f["qc_lo"] == f["qc_hi"] and f["round"] == 1
    a = f["author_item"]
    assert pre(a) == chain[0].preimage() and g["pk"][a].tobytes() == chain[0].author.b and g["sig"][a].tobytes() == chain[0].signature.flatten()
    assert g["mode"][a] == 0 and g["group_idx"][a] == 0
    # block 1: QC over block 0 with 4 votes under the verify_batch condition
    f = info[1]
    assert not f["qc_is_genesis"] and f["qc_hi"] - f["qc_lo"] == 4 and f["qc_round"] == 1
    for k, (name, sig) in enumerate(chain[1].qc.votes):
        i = f["qc_lo"] + k
        assert g["pk"][i].tobytes() == name.b and g["sig"][i].tobytes() == sig.flatten() and g["mode"][i] == 1 and g["group_idx"][i] == 1
        assert pre(i) == messages.vote_preimage(chain[1].qc.hash, chain[1].qc.round)
    # block with TC and payload
    f = info[4]
    assert f["has_tc"] and f["tc_hi"] - f["tc_lo"] == 3 and f["tc_round"] == 8
    assert pre(f["author_item"]) == blk_tc.preimage() and len(blk_tc.preimage()) == 32 + 8 + 64 + 32
    for k, (name, sig, r) in enumerate(blk_tc.tc.votes):
        i = f["tc_lo"] + k
        assert pre(i) == messages.timeout_preimage(8, r) and g["mode"][i] == 0 and g["pk"][i].tobytes() == name.b
    # vote, timeouts, TC, sync request
    assert pre(info[5]["author_item"]) == messages.vote_preimage(v.hash, v.round)
    assert pre(info[6]["author_item"]) == messages.timeout_preimage(9, to.high_qc.round) and info[6]["qc_hi"] - info[6]["qc_lo"] == 4
    assert info[7]["qc_is_genesis"] and info[7]["qc_lo"] == info[7]["qc_hi"]
    assert info[8]["author_item"] == wire.NO_ITEM and info[8]["tc_hi"] - info[8]["tc_lo"] == 3
    assert info[9]["author_item"] == wire.NO_ITEM
    assert len(g["sig"]) == 1 + 5 + 5 + 5 + (1 + 4 + 3) + 1 + (1 + 4) + 1 + 3


def test_verify_frames_equals_struct_level_verification(oracle, golden):
    run_verify_frames(_fx(oracle, golden))


def run_verify_frames(fx):
    chain, blk_tc, v, to, to_gen = _messages(fx)
    bad_sig = fx.block(2, 6, qc=chain[2].qc)
    bad_sig.round = 7
    reuse = fx.block(0, 6, qc=fx.qc_for(fx.d(b"y"), 5))
    reuse.qc.votes[1] = reuse.qc.votes[0]
    bad_vote = messages.Vote(v.hash, 2, v.author, v.signature)
    short_tc = fx.tc(8, hqs=((0, 3), (1, 5)))
    outsider = messages.Vote(v.hash, 1, crypto.PublicKey(bytes(range(32))), v.signature)
    blocks = chain + [blk_tc, bad_sig, reuse]
    frames = [bc.propose(b) for b in blocks] + [bc.vote(v), bc.vote(bad_vote), bc.vote(outsider), bc.timeout(to), bc.timeout(to_gen),
                                                bc.tc_msg(fx.tc(7)), bc.tc_msg(short_tc), bc.sync_request(fx.d(b"m"), fx.pks[1])]
    got = wire.verify_frames(frames, fx.committee, fx.e)
    want = messages.verify_blocks(blocks, fx.committee, fx.e) + [None, "InvalidSignature", "UnknownAuthority", None, None, None, "TCRequiresQuorum", None]
    assert got == want, (got, want)
    assert want[:7] == [None] * 5 + ["InvalidSignature", "AuthorityReuse"]


def test_malformed_frames_are_isolated(oracle, golden):
    """Truncations at every length, bad tags, bad base64, absurd vector lengths: kind = malformed, no items, neighbours untouched."""
    fx = _fx(oracle, golden)
    chain, blk_tc, v, to, _ = _messages(fx)
    good = bc.propose(blk_tc)
    ref = wire.ingest_frames([good])
    n_items = len(ref["sig"])
    rng = np.random.default_rng(3)
    cuts = sorted(set([0, 1, 3, 4, 5, 40, 47, 48, len(good) - 1] + [int(x) for x in rng.integers(0, len(good), 60)]))
    frames = []
    for c in cuts:
        frames += [good, good[:c]]
    g = wire.ingest_frames(frames)
    assert list(g["info"]["kind"][0::2]) == [0] * len(cuts) and list(g["info"]["kind"][1::2]) == [255] * len(cuts)
    assert len(g["sig"]) == n_items * len(cuts)
    assert (g["group_idx"].reshape(len(cuts), n_items) == (2 * np.arange(len(cuts)))[:, None]).all()
    # trailing bytes are allowed (bincode::deserialize default), a wrong enum tag is not
    assert list(wire.ingest_frames([good + b"xyz", b"\x05\x00\x00\x00" + good[4:]])["info"]["kind"]) == [0, 255]
    # base64 damage inside a key string: bad alphabet, missing padding, non-zero trailing bits, too short
    vt = bc.vote(v)
    at = 4 + 32 + 8 + 8                       # first character of the author's base64 string
    for mut in (lambda b: b[:at] + b"*" + b[at + 1:], lambda b: b[:at + 10] + b"=" + b[at + 11:], lambda b: b[:at + 42] + b"B" + b[at + 43:]):
        assert wire.ingest_frames([mut(vt)])["info"]["kind"][0] == 255
    nopad = vt[:at + 43] + b"A" + vt[at + 44:]   # 44 characters without padding decode to 33 bytes: the reference keeps the first 32
    assert wire.ingest_frames([nopad])["info"]["kind"][0] == 1 and (wire.ingest_frames([nopad])["pk"][0] == wire.ingest_frames([vt])["pk"][0]).all()
    short = vt[:at - 8] + (40).to_bytes(8, "little") + vt[at:at + 40] + vt[at + 44:]
    assert wire.ingest_frames([short])["info"]["kind"][0] == 255
    # a vote count far beyond the frame
    huge = b"\x03\x00\x00\x00" + (7).to_bytes(8, "little") + (2**61).to_bytes(8, "little")
    assert wire.ingest_frames([huge])["info"]["kind"][0] == 255
    assert wire.ingest_frames([])["n_frames"] == 0
    # random garbage never crashes and never yields items
    junk = [rng.bytes(int(n)) for n in rng.integers(0, 400, 200)]
    gj = wire.ingest_frames(junk)
    assert len(gj["sig"]) == 0 or (gj["info"]["kind"] != 255).any()


def test_ingest_mutation_fuzz_under_address_sanitizer(oracle, golden, tmp_path):
    """The parser reads untrusted network bytes: hs_ingest.cpp + a mutation fuzzer (tests/cpp/ingest_fuzz.cpp) are built with
    -fsanitize=address,undefined and run over 60,000 mutated batches of the re-serialised reference fixtures with deliberately tight,
    exactly-sized input and output buffers.  Any out-of-bounds read / write, overflow or broken invariant fails the test."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fx = _fx(oracle, golden)
    chain, blk_tc, v, to, to_gen = _messages(fx)
    frames = [bc.propose(b) for b in chain] + [bc.propose(blk_tc), bc.vote(v), bc.timeout(to), bc.timeout(to_gen), bc.tc_msg(fx.tc(7)),
                                              bc.sync_request(fx.d(b"missing"), fx.pks[0])]
    seeds = tmp_path / "seeds.bin"
    with open(seeds, "wb") as f:
        for fr in frames:
            f.write(struct.pack("<I", len(fr)) + fr)
    exe = str(tmp_path / "ingest_fuzz")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe,
                           os.path.join(root, "tests", "cpp", "ingest_fuzz.cpp"), os.path.join(root, "hotstuff_b200", "csrc", "hs_ingest.cpp")])
    out = subprocess.run([exe, str(seeds), "60000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ingest fuzz ok" in out.stdout, (out.returncode, out.stdout[-500:], out.stderr[-3000:])
    parsed = int(out.stdout.split("frames parsed")[0].split(",")[-1])
    assert parsed > 10000          # the mutations leave plenty of frames parseable: both sides of the parser are exercised


def test_key_string_decoder_against_a_strict_base64_reference(oracle, golden):
    """The 44-character fast path and the general path of the key decoder (crypto/src/lib.rs:103-112: base64::decode, then bytes[..32])
    against Python's base64 with the same strictness (standard alphabet, canonical padding and trailing bits): 20 000 vote frames whose
    author string has 0-3 characters replaced by arbitrary bytes, plus 48- and 88-character strings."""
    import base64
    fx = _fx(oracle, golden)
    _, _, v, _, _ = _messages(fx)
    vt = bc.vote(v)
    at = 4 + 32 + 8 + 8
    head, tail = vt[:at - 8], vt[at + 44:]
    rng = np.random.default_rng(11)
    alphabet = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/="

    def ref(s):
        if len(s) % 4 or len(s) < 44:
            return None
        try:
            d = base64.b64decode(s, validate=True)
        except Exception:
            return None
        return d[:32] if base64.b64encode(d) == s and len(d) >= 32 else None

    strings = []
    for _ in range(20000):
        key = rng.bytes(32 if rng.random() < 0.9 else int(rng.choice([33, 34, 35, 36, 64, 66])))
        s = bytearray(base64.b64encode(key))
        for _ in range(int(rng.integers(0, 4))):
            s[int(rng.integers(0, len(s)))] = alphabet[int(rng.integers(0, 65))] if rng.random() < 0.8 else int(rng.integers(0, 256))
        strings.append(bytes(s))
    frames = [head + len(s).to_bytes(8, "little") + s + tail for s in strings]
    g = wire.ingest_frames(frames)
    want = [ref(s) for s in strings]
    kinds = g["info"]["kind"]
    assert [k != 255 for k in kinds] == [w is not None for w in want]
    assert sum(w is not None for w in want) > 3000 and sum(w is None for w in want) > 3000
    items = iter(range(len(g["pk"])))
    for j, w in enumerate(want):
        if w is not None:
            assert g["pk"][next(items)].tobytes() == w, j
