"""Property-based checks (hypothesis) of the device field / scalar code under host emulation: algebraic identities that must hold
for EVERY 256-bit input, not just the sampled ones of test_hostemu.py."""
import ctypes

from hypothesis import given, settings, strategies as st

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
u256 = st.integers(min_value=0, max_value=2**256 - 1)
# bias towards the edges where lost carries hide
edgy = st.one_of(u256, st.sampled_from([0, 1, P - 1, P, P + 1, 2 * P, 2**256 - 1, 2**256 - 38, 2**255, 2**32 - 1, (2**256 - 1) ^ (2**32 - 1)]),
                 st.integers(min_value=0, max_value=2**40).map(lambda x: 2**256 - 1 - x))


def _fe(lib, op, a, b=0):
    o = ctypes.create_string_buffer(32)
    lib.emu_fe_op(op, int(a).to_bytes(32, "little"), int(b).to_bytes(32, "little"), o)
    return int.from_bytes(o.raw, "little")


@settings(max_examples=400, deadline=None)
@given(edgy, edgy, edgy)
def test_field_ring_identities(hostemu, a, b, c):
    mul = lambda x, y: _fe(hostemu, 0, x, y)   # noqa: E731
    add = lambda x, y: _fe(hostemu, 2, x, y)   # noqa: E731
    sub = lambda x, y: _fe(hostemu, 3, x, y)   # noqa: E731
    assert mul(a, b) % P == a * b % P
    assert _fe(hostemu, 1, a) % P == a * a % P
    assert add(a, b) % P == (a + b) % P and sub(a, b) % P == (a - b) % P
    assert mul(a, add(b, c)) % P == add(mul(a, b), mul(a, c)) % P            # distributivity through the reduced forms
    assert sub(add(a, b), b) % P == a % P
    assert _fe(hostemu, 4, a) == a % P                                        # canonical form is the unique representative


@settings(max_examples=60, deadline=None)
@given(edgy)
def test_field_inverse_and_sqrt_chain(hostemu, a):
    inv = _fe(hostemu, 5, a)
    assert (inv * a) % P == (0 if a % P == 0 else 1)
    assert _fe(hostemu, 6, a) % P == pow(a % P, (P - 5) // 8, P)


@settings(max_examples=300, deadline=None)
@given(st.integers(min_value=0, max_value=2**512 - 1))
def test_scalar_reduction(hostemu, x):
    o = ctypes.create_string_buffer(32)
    hostemu.emu_sc_reduce512(x.to_bytes(64, "little"), o)
    assert int.from_bytes(o.raw, "little") == x % L


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=0, max_value=L - 1), st.sampled_from([4, 8, 10, 12, 13, 14, 15, 16, 20, 24]))
def test_signed_window_recoding_reconstructs_the_scalar(hostemu, s, w):
    out = (ctypes.c_int * 80)()
    n = hostemu.emu_sc_digits(w, 0, s.to_bytes(32, "little"), out)
    d = list(out)[:n]
    assert all(-(1 << (w - 1)) <= x < (1 << (w - 1)) for x in d)
    assert sum(x << (w * i) for i, x in enumerate(d)) == s
