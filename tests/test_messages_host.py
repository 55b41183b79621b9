"""Host logic of the call-site mirror (pre-check order, grouping into engine calls, verified-QC cache) on an oracle-backed stub
engine — runs without a GPU.  The same scenarios run against the CUDA engine in tests/test_messages.py."""
import messages_scenarios as sc


def _fx(oracle, golden):
    return sc.Fixtures(oracle, golden, sc.OracleStubEngine(oracle))


def test_block_vote_timeout_verify(oracle, golden):
    sc.scenario_block_vote_timeout(_fx(oracle, golden))


def test_blocks_batched_error_order(oracle, golden):
    sc.scenario_blocks_batched(_fx(oracle, golden))


def test_tcs_and_timeout_burst_with_qc_cache(oracle, golden):
    fx = _fx(oracle, golden)
    sc.scenario_tcs_and_timeout_burst(fx, count_votes=lambda: fx.e.calls["qc_votes"])
