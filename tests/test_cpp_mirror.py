"""include/hs_crypto.hpp — the compiled-language (C++) mirror of the crate API: compiles against the C ABI on any box; on a GPU box
the C++ port of crypto_tests.rs runs against the engine."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "crypto_tests")


def _build(name="crypto_tests"):
    from hotstuff_b200 import build
    lib = build.build_engine()
    out = os.path.join(ROOT, "tests", "cpp", name)
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
                           lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return out


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(_build())
    assert os.path.exists(_build("consensus_tests"))


def _crypto_args(oracle, golden):
    r = golden["reference"]
    seeds = [bytes.fromhex(s) for s in r["seeds"]]
    hello = bytes.fromhex(r["hello_digest"])
    return [r["hello_digest"], r["bad_digest"], r["hello_sig_key3"], r["pks"][3], r["pks"][2], r["pks"][1],
            oracle.sign(seeds[2], hello).hex(), oracle.sign(seeds[1], hello).hex(), r["serialized_batch"], r["batch_digest"]]


def _consensus_args(oracle, golden):
    r = golden["reference"]
    seeds = [bytes.fromhex(s) for s in r["seeds"]]
    args = list(r["pks"])
    for v in r["qc_votes"]:
        args += [v["pk"], v["sig"]]
    vote_hash = oracle.digest32(b"some block")
    args += [vote_hash.hex(), oracle.sign(seeds[3], oracle.digest32(vote_hash + (1).to_bytes(8, "little"))).hex()]
    args += [oracle.sign(seeds[2], oracle.digest32((9).to_bytes(8, "little") + (1).to_bytes(8, "little"))).hex()]
    for i, hq in ((0, 3), (1, 5), (2, 4)):
        args.append(oracle.sign(seeds[i], oracle.digest32((7).to_bytes(8, "little") + hq.to_bytes(8, "little"))).hex())
    return args


@pytest.mark.gpu
def test_cpp_port_of_reference_crypto_tests(oracle, golden):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = subprocess.run([_build()] + _crypto_args(oracle, golden), capture_output=True, text=True)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stderr


@pytest.mark.gpu
def test_cpp_port_of_reference_messages_tests(oracle, golden):
    """include/hs_consensus.hpp against consensus/src/tests/messages_tests.rs (QC cases) + Vote / Timeout / TC, compiled C++."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = subprocess.run([_build("consensus_tests")] + _consensus_args(oracle, golden), capture_output=True, text=True)
    assert out.returncode == 0 and "cpp consensus mirror ok" in out.stdout, out.stderr + out.stdout


def _build_on_stub(name):
    """The same test source linked against tests/cpp/stub_abi.cpp (the C ABI subset answered by the CPU oracle) instead of libhs_crypto.so."""
    from hotstuff_b200 import build
    olib = build.build_oracle()
    out = os.path.join(ROOT, "tests", "cpp", name + "_stub")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", out, os.path.join(ROOT, "tests", "cpp", name + ".cpp"), os.path.join(ROOT, "tests", "cpp", "stub_abi.cpp"),
                           os.path.join(ROOT, "hotstuff_b200", "csrc", "hs_ingest.cpp"), olib, "-Wl,-rpath," + os.path.dirname(olib)])
    return out


def test_cpp_ports_of_the_reference_tests_run_on_the_oracle_stub(oracle, golden):
    """The host logic of the compiled mirror (types, digest layouts, pre-checks, error names, batching, the verified-QC cache) does not
    need a GPU to be wrong: the C++ ports of crypto_tests.rs and messages_tests.rs also run here, with the engine's entry points answered
    by the oracle.  (On a GPU box the same sources run against the CUDA engine.)"""
    out = subprocess.run([_build_on_stub("crypto_tests")] + _crypto_args(oracle, golden), capture_output=True, text=True)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stderr + out.stdout
    out = subprocess.run([_build_on_stub("consensus_tests")] + _consensus_args(oracle, golden), capture_output=True, text=True)
    assert out.returncode == 0 and "cpp consensus mirror ok" in out.stdout, out.stderr + out.stdout


def test_cpp_verify_frames_host_logic_matches_the_python_mirror(oracle, golden, tmp_path):
    """hs::verify_frames (C++: ingest -> pre-checks -> one grouped pass -> first error per frame) on the frames of test_wire_ingest plus
    malformed ones, with the CPU oracle standing in for the engine (test only): same outcome per frame as wire.verify_frames, which is
    itself held against struct-level verification.  Also checks that items of certificates that fail a pre-check are never judged."""
    import struct
    import bincode_ref as bc
    import messages_scenarios as sc
    from hotstuff_b200 import build, crypto, messages, wire
    fx = sc.Fixtures(oracle, golden, sc.OracleStubEngine(oracle))
    chain = fx.chain(4)
    blk_tc = fx.block(1, 9, qc=chain[3].qc, tc=fx.tc(8), payload=[fx.d(b"p1"), fx.d(b"p2")])
    v = messages.Vote(fx.d(chain[0].preimage()), 1, fx.pks[3], crypto.Signature())
    v.signature = fx.sign(3, fx.d(messages.vote_preimage(v.hash, v.round)))
    to = fx.timeout(2, 9, chain[2].qc)
    to_gen = fx.timeout(1, 4, messages.QC.genesis())
    bad_sig = fx.block(2, 6, qc=chain[2].qc)
    bad_sig.round = 7
    reuse = fx.block(0, 6, qc=fx.qc_for(fx.d(b"y"), 5))
    reuse.qc.votes[1] = reuse.qc.votes[0]
    bad_qc_vote = fx.block(3, 6, qc=fx.qc_for(fx.d(b"z"), 5))
    n0, s0 = bad_qc_vote.qc.votes[2]
    bad_qc_vote.qc.votes[2] = (n0, crypto.Signature(bytes([s0.flatten()[0] ^ 1]) + s0.flatten()[1:]))
    bad_vote = messages.Vote(v.hash, 2, v.author, v.signature)
    short_tc = fx.tc(8, hqs=((0, 3), (1, 5)))
    outsider = messages.Vote(v.hash, 1, crypto.PublicKey(bytes(range(32))), v.signature)
    good = bc.propose(blk_tc)
    frames = [bc.propose(b) for b in chain + [blk_tc, bad_sig, reuse, bad_qc_vote]] + [
        bc.vote(v), bc.vote(bad_vote), bc.vote(outsider), bc.timeout(to), bc.timeout(to_gen), bc.tc_msg(fx.tc(7)), bc.tc_msg(short_tc),
        bc.sync_request(fx.d(b"m"), fx.pks[1]), good[:40], b"\x05\x00\x00\x00" + good[4:], b"", good + b"trailing"]
    want = ["OK" if w is None else w for w in wire.verify_frames(frames, fx.committee, fx.e)]
    assert want[:8] == ["OK"] * 5 + ["InvalidSignature", "AuthorityReuse", "InvalidSignature"] and want[-4:] == ["Malformed", "Malformed", "Malformed", "OK"]
    blob = struct.pack("<I", len(fx.committee.stakes))
    for k, st in fx.committee.stakes.items():
        blob += k + struct.pack("<I", st)
    blob += struct.pack("<I", len(frames)) + b"".join(struct.pack("<I", len(f)) + f for f in frames)
    path = tmp_path / "frames.bin"
    path.write_bytes(blob)
    lib = build.build_engine()
    olib = build.build_oracle()
    exe = os.path.join(ROOT, "tests", "cpp", "frames_host_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "frames_host_test.cpp"), lib, olib,
                           "-Wl,-rpath," + os.path.dirname(lib), "-Wl,-rpath," + os.path.dirname(olib)])
    out = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split("\n")[:-1] == want, (out.stdout, want)
    # hs::Block::preimage (Block::digest layout, messages.rs:79-90) equals the Python mirror's, whose digest the ingest test pins to the frame bytes
    pre = subprocess.run([exe, "--block-preimage", blk_tc.author.b.hex(), str(blk_tc.round), blk_tc.qc.hash.b.hex()] + [d.b.hex() for d in blk_tc.payload],
                         capture_output=True, text=True)
    assert pre.returncode == 0 and pre.stdout.strip() == blk_tc.preimage().hex()
    # items judged: everything except the certificates that failed a pre-check (reuse: 4 QC votes; short_tc: 2 votes; outsider: 1)
    total = len(wire.ingest_frames(frames)["sig"])
    assert "judged %d items" % (total - 4 - 2 - 1) in out.stderr, (out.stderr, total)
