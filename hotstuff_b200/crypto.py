"""Host-side mirror of the reference's `crypto` crate surface (crypto/src/lib.rs) over the C ABI.

Same names and argument meaning as the Rust API so the parity tests read like crypto/src/tests/crypto_tests.rs:
    Digest, PublicKey, Signature.verify(digest, public_key), Signature.verify_batch(digest, votes), CryptoError.
Signing (Signature::new, SignatureService, generate_keypair: lib.rs:167-191,225-250) stays on the CPU in the
reference node and is out of the GPU path; tests synthesise signatures with the oracle.
"""
import base64
import threading

import numpy as np

from .engine import Engine


class CryptoError(Exception):
    """Mirror of `CryptoError = ed25519::Error` (crypto/src/lib.rs:18): opaque, only Ok/Err is observable."""


_engine = None
_engine_lock = threading.Lock()


def default_engine():
    global _engine
    with _engine_lock:
        if _engine is None:
            _engine = Engine(0)
        return _engine


def set_default_engine(e):
    global _engine
    _engine = e


class Digest:
    """32-byte hash digest (crypto/src/lib.rs:22)."""

    def __init__(self, b=bytes(32)):
        b = bytes(b)
        if len(b) != 32:
            raise ValueError("Digest must be 32 bytes")  # TryFrom<&[u8]> error, lib.rs:53-58
        self.b = b

    def to_vec(self):
        return self.b

    def size(self):
        return 32

    def __eq__(self, o):
        return isinstance(o, Digest) and o.b == self.b

    def __hash__(self):
        return hash(self.b)

    def __repr__(self):
        return base64.b64encode(self.b).decode()

    @staticmethod
    def of(data, engine=None):
        """Digest(SHA-512(data)[..32]) on the GPU — the Hash impls of messages.rs:79-90,149-156 and processor.rs:30."""
        return digest_many([data], engine)[0]


def digest_many(items, engine=None):
    e = engine or default_engine()
    off = np.zeros(len(items) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in items])
    out = e.digest32_batch(b"".join(bytes(x) for x in items), off)
    return [Digest(out[i].tobytes()) for i in range(len(items))]


class PublicKey:
    """32-byte compressed Edwards point (crypto/src/lib.rs:66); base64 import/export as lib.rs:69-80."""

    def __init__(self, b=bytes(32)):
        b = bytes(b)
        if len(b) != 32:
            raise ValueError("PublicKey must be 32 bytes")
        self.b = b

    def encode_base64(self):
        return base64.b64encode(self.b).decode()

    @staticmethod
    def decode_base64(s):
        raw = base64.b64decode(s, validate=True)
        if len(raw) < 32:
            raise ValueError("InvalidLength")
        return PublicKey(raw[:32])

    def __eq__(self, o):
        return isinstance(o, PublicKey) and o.b == self.b

    def __hash__(self):
        return hash(self.b)

    def __repr__(self):
        return self.encode_base64()


class Signature:
    """Ed25519 signature {part1, part2} (crypto/src/lib.rs:179-182); Default = 64 zero bytes."""

    def __init__(self, b=bytes(64)):
        b = bytes(b)
        if len(b) != 64:
            raise ValueError("Signature must be 64 bytes")
        self.part1, self.part2 = b[:32], b[32:]

    def flatten(self):
        return self.part1 + self.part2

    def verify(self, digest, public_key, engine=None):
        """Signature::verify (lib.rs:200-204): dalek verify_strict.  Returns None, raises CryptoError on Err."""
        e = engine or default_engine()
        rec = np.frombuffer(self.flatten() + public_key.b + digest.b, dtype=np.uint8).reshape(1, 128)
        if not e.verify_strict_batch(rec)[0]:
            raise CryptoError("signature error")

    @staticmethod
    def verify_batch(digest, votes, engine=None):
        """Signature::verify_batch (lib.rs:206-219): one digest, votes = iterable of (PublicKey, Signature)."""
        e = engine or default_engine()
        votes = list(votes)
        buf = np.frombuffer(b"".join(pk.b + sig.flatten() for pk, sig in votes), dtype=np.uint8).reshape(-1, 96)
        if not e.verify_batch_shared_msg(digest.b, buf):
            raise CryptoError("signature error")
