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


@pytest.mark.gpu
def test_cpp_port_of_reference_crypto_tests(oracle, golden):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = golden["reference"]
    seeds = [bytes.fromhex(s) for s in r["seeds"]]
    hello = bytes.fromhex(r["hello_digest"])
    args = [r["hello_digest"], r["bad_digest"], r["hello_sig_key3"], r["pks"][3], r["pks"][2], r["pks"][1],
            oracle.sign(seeds[2], hello).hex(), oracle.sign(seeds[1], hello).hex(), r["serialized_batch"], r["batch_digest"]]
    out = subprocess.run([_build()] + args, capture_output=True, text=True)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stderr


@pytest.mark.gpu
def test_cpp_port_of_reference_messages_tests(oracle, golden):
    """include/hs_consensus.hpp against consensus/src/tests/messages_tests.rs (QC cases) + Vote / Timeout / TC, compiled C++."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
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
    out = subprocess.run([_build("consensus_tests")] + args, capture_output=True, text=True)
    assert out.returncode == 0 and "cpp consensus mirror ok" in out.stdout, out.stderr + out.stdout
