"""bench.py's CPU-only legs and the JSON contract of its line (the GPU arm is exercised on the box).  The reference arm
(`--impl reference`) must run without the engine, print ONE JSON line and carry the keys the driver compares across arms."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "3000", "--ref-sample", "3000", "--keys", "64",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Ed25519 verifies/s" and d["unit"] == "verifies/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 1000 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "verifies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "config[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True,
                         text=True, cwd=ROOT, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_input_synthesis_with_the_oracle_signer_matches_openssl():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import hashlib
    import bench
    from oracle_api import Oracle
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PublicKey
    o = Oracle()
    inp = bench.make_inputs(500, 16, 512, seed=9, corrupt_frac=0.02, oracle=o)
    assert inp["sig"].shape == (500, 64) and inp["msgs"].shape == (500, 512) and int(inp["corrupted"].sum()) == 10
    for i in range(0, 500, 37):
        if inp["corrupted"][i]:
            continue
        d = hashlib.sha512(inp["msgs"][i].tobytes()).digest()[:32]
        Ed25519PublicKey.from_public_bytes(inp["pk"][i].tobytes()).verify(inp["sig"][i].tobytes(), d)   # raises on a bad signature
    ok = o.verify_rec128(__import__("numpy").concatenate([inp["sig"], inp["pk"], o.digest32_batch(inp["msgs"].reshape(-1), __import__("numpy").arange(501, dtype="uint64") * 512)], axis=1))
    assert (ok == ~inp["corrupted"]).all()
