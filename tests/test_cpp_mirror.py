"""include/hs_crypto.hpp — the compiled-language (C++) mirror of the crate API: compiles against the C ABI on any box; on a GPU box
the C++ port of crypto_tests.rs runs against the engine."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "crypto_tests")


def _build():
    from hotstuff_b200 import build
    lib = build.build_engine()
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", BIN, os.path.join(ROOT, "tests", "cpp", "crypto_tests.cpp"),
                           lib, "-Wl,-rpath," + os.path.dirname(lib)])
    return BIN


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(_build())


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
