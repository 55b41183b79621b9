"""Engine: Python handle on one hs_ctx (one CUDA device).  Thin: marshals numpy / torch buffers into the C ABI.

Host-side mirror of the reference interface lives in crypto.py; this module is the batch surface the consensus call
sites would use (QC/TC vote sets, mempool batch digests) plus device-resident entry points for the benchmark.
"""
import ctypes

import numpy as np

from . import _lib

MODE_STRICT = 0    # Signature::verify      (crypto/src/lib.rs:200-204)
MODE_BATCH_EQ = 1  # Signature::verify_batch (crypto/src/lib.rs:206-219), per-signature condition


class EngineError(RuntimeError):
    pass


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(a, shape_last=None):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray, memoryview)) else a, dtype=np.uint8)
    if shape_last is not None and a.size % shape_last:
        raise ValueError("buffer length is not a multiple of %d" % shape_last)
    return a


def bitmap_to_bools(bitmap, n):
    return np.unpackbits(bitmap.view(np.uint8), bitorder="little")[:n].astype(bool)


class Engine:
    def __init__(self, device=0, base_window=0, key_window=0, key_cache=True):
        """base_window / key_window: comb window widths in bits (0 = engine defaults: 24 and the widest that fits).
        key_cache: learn tables for unregistered keys between calls (include/hs_crypto.h, hs_cached_keys)."""
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.hs_ctx_create(ctypes.byref(h), int(device), (int(base_window) & 0xff) | ((int(key_window) & 0xff) << 8) | (0 if key_cache else 0x10000))
        if rc != 0 or not h:
            raise EngineError("hs_ctx_create(device=%d) failed with status %d (no GPU / CUDA error); there is no CPU fallback" % (device, rc))
        self.h = h
        self.device = int(device)
        self.n_keys = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.hs_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed: status %d: %s" % (what, rc, self.lib.hs_last_error(self.h).decode()))

    @property
    def cached_keys(self):
        return int(self.lib.hs_cached_keys(self.h))

    @property
    def window_bits(self):
        """(per-key comb window bits or 0, base-point comb window bits)"""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        self.lib.hs_window_bits(self.h, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    @property
    def kernel_launches(self):
        return int(self.lib.hs_kernel_launches(self.h))

    # ---- host-buffer API -------------------------------------------------------------------------------------
    def verify_rec128(self, recs, mode=MODE_STRICT):
        """recs: (n,128) uint8 [sig64|pk32|msg32] -> bool[n]"""
        recs = _u8(recs, 128).reshape(-1, 128)
        n = recs.shape[0]
        bm = np.zeros((n + 31) // 32, dtype=np.uint32)
        self._check(self.lib.hs_verify_rec128(self.h, _ptr(recs), n, mode, _ptr(bm)), "hs_verify_rec128")
        return bitmap_to_bools(bm, n)

    def verify_strict_batch(self, recs):
        recs = _u8(recs, 128).reshape(-1, 128)
        n = recs.shape[0]
        bm = np.zeros((n + 31) // 32, dtype=np.uint32)
        self._check(self.lib.hs_verify_strict_batch(self.h, _ptr(recs), n, _ptr(bm)), "hs_verify_strict_batch")
        return bitmap_to_bools(bm, n)

    def verify_var(self, sig, pk, msgs, off, mode=MODE_STRICT):
        sig = _u8(sig, 64).reshape(-1, 64)
        pk = _u8(pk, 32).reshape(-1, 32)
        n = sig.shape[0]
        off = np.ascontiguousarray(off, dtype=np.uint64)
        assert pk.shape[0] == n and off.shape[0] == n + 1
        msgs = _u8(msgs)
        bm = np.zeros((n + 31) // 32, dtype=np.uint32)
        self._check(self.lib.hs_verify_var(self.h, _ptr(sig), _ptr(pk), _ptr(msgs) if msgs.size else None, _ptr(off), n, mode, _ptr(bm)),
                    "hs_verify_var")
        return bitmap_to_bools(bm, n)

    def verify_batch_shared_msg(self, digest, votes, want_bitmap=False):
        """votes: (n,96) uint8 [pk32|sig64].  Returns all_ok (and bool[n] when want_bitmap)."""
        digest = _u8(digest)
        assert digest.size == 32
        votes = _u8(votes, 96).reshape(-1, 96)
        n = votes.shape[0]
        ok = ctypes.c_int(0)
        bm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32) if want_bitmap else None
        self._check(self.lib.hs_verify_batch_shared_msg(self.h, _ptr(digest), _ptr(votes) if n else None, n, ctypes.byref(ok), _ptr(bm)),
                    "hs_verify_batch_shared_msg")
        return (bool(ok.value), bitmap_to_bools(bm, n)) if want_bitmap else bool(ok.value)

    def verify_qcs(self, preimages, sig, qc_idx, pk=None, validator_idx=None, want_votes=False):
        """Many QCs in one pass: preimages (n_qc,40) = hash||round_le; vote i -> certificate qc_idx[i].  Returns bool[n_qc]
        (and bool[n_votes] when want_votes)."""
        pre = _u8(preimages, 40).reshape(-1, 40)
        sig = _u8(sig, 64).reshape(-1, 64)
        n_qc, n = pre.shape[0], sig.shape[0]
        qi = np.ascontiguousarray(qc_idx, dtype=np.uint32)
        assert qi.shape[0] == n and (pk is None) != (validator_idx is None)
        pk = None if pk is None else _u8(pk, 32).reshape(-1, 32)
        vidx = None if validator_idx is None else np.ascontiguousarray(validator_idx, dtype=np.uint32)
        qbm = np.zeros(max(1, (n_qc + 31) // 32), dtype=np.uint32)
        vbm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32) if want_votes else None
        self._check(self.lib.hs_verify_qcs(self.h, _ptr(pre) if n_qc else None, n_qc, _ptr(pk), _ptr(vidx), _ptr(sig) if n else None,
                                           _ptr(qi) if n else None, n, _ptr(vbm), _ptr(qbm)), "hs_verify_qcs")
        out = bitmap_to_bools(qbm, n_qc)
        return (out, bitmap_to_bools(vbm, n)) if want_votes else out

    def verify_tcs(self, tc_rounds, sig, high_qc_rounds, tc_idx=None, pk=None, validator_idx=None, want_votes=False):
        """TC::verify for many TCs (tc_idx given) or Timeout signatures (tc_idx None: one vote per certificate): the 16-byte
        digests are built on the GPU from (tc_round, high_qc_round).  Returns bool[n_tc] (and bool[n_votes])."""
        tr = np.ascontiguousarray(tc_rounds, dtype=np.uint64)
        hq = np.ascontiguousarray(high_qc_rounds, dtype=np.uint64)
        sig = _u8(sig, 64).reshape(-1, 64)
        n, n_tc = sig.shape[0], tr.shape[0]
        assert hq.shape[0] == n and (pk is None) != (validator_idx is None)
        ti = None if tc_idx is None else np.ascontiguousarray(tc_idx, dtype=np.uint32)
        pk = None if pk is None else _u8(pk, 32).reshape(-1, 32)
        vidx = None if validator_idx is None else np.ascontiguousarray(validator_idx, dtype=np.uint32)
        tbm = np.zeros(max(1, (n_tc + 31) // 32), dtype=np.uint32)
        vbm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32) if want_votes else None
        self._check(self.lib.hs_verify_tcs(self.h, _ptr(tr) if n_tc else None, n_tc, _ptr(pk), _ptr(vidx), _ptr(sig) if n else None, _ptr(hq) if n else None,
                                           _ptr(ti), n, _ptr(vbm), _ptr(tbm)), "hs_verify_tcs")
        out = bitmap_to_bools(tbm, n_tc)
        return (out, bitmap_to_bools(vbm, n)) if want_votes else out

    def verify_groups(self, preimages, pre_off, sig, msg_idx, group_idx, n_groups, mode=None, pk=None, validator_idx=None, want_items=False):
        """hs_verify_groups: items over GPU-hashed variable-length preimages, per-item verdict mode, per-group AND."""
        pre = _u8(preimages)
        off = np.ascontiguousarray(pre_off, dtype=np.uint64)
        sig = _u8(sig, 64).reshape(-1, 64)
        n = sig.shape[0]
        mi = np.ascontiguousarray(msg_idx, dtype=np.uint32)
        gi = np.ascontiguousarray(group_idx, dtype=np.uint32)
        mo = None if mode is None else np.ascontiguousarray(mode, dtype=np.uint8)
        assert (pk is None) != (validator_idx is None) and mi.shape[0] == n == gi.shape[0]
        pk = None if pk is None else _u8(pk, 32).reshape(-1, 32)
        vidx = None if validator_idx is None else np.ascontiguousarray(validator_idx, dtype=np.uint32)
        gbm = np.zeros(max(1, (n_groups + 31) // 32), dtype=np.uint32)
        ibm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32) if want_items else None
        self._check(self.lib.hs_verify_groups(self.h, _ptr(pre) if pre.size else None, _ptr(off), off.shape[0] - 1, _ptr(sig) if n else None, _ptr(pk), _ptr(vidx),
                                              _ptr(mi) if n else None, _ptr(gi) if n else None, _ptr(mo), n, n_groups, _ptr(ibm), _ptr(gbm)), "hs_verify_groups")
        out = bitmap_to_bools(gbm, n_groups)
        return (out, bitmap_to_bools(ibm, n)) if want_items else out

    def keygen_batch(self, seeds):
        """RFC 8032 public keys of n 32-byte seeds, computed on the GPU (load generation; hs_keygen_batch)."""
        seeds = _u8(seeds, 32).reshape(-1, 32)
        out = np.zeros_like(seeds)
        self._check(self.lib.hs_keygen_batch(self.h, _ptr(seeds), seeds.shape[0], _ptr(out)), "hs_keygen_batch")
        return out

    def sign_digests(self, seeds, pks, digests, key_idx=None):
        """RFC 8032 signatures over 32-byte digests, made on the GPU (load generation; hs_sign_digests)."""
        seeds = _u8(seeds, 32).reshape(-1, 32)
        pks = _u8(pks, 32).reshape(-1, 32)
        digests = _u8(digests, 32).reshape(-1, 32)
        ki = None if key_idx is None else np.ascontiguousarray(key_idx, dtype=np.uint32)
        n = digests.shape[0]
        out = np.zeros((n, 64), dtype=np.uint8)
        self._check(self.lib.hs_sign_digests(self.h, _ptr(seeds), _ptr(pks), seeds.shape[0], _ptr(ki), _ptr(digests), n, _ptr(out)), "hs_sign_digests")
        return out

    def committee_register(self, pks):
        pks = _u8(pks, 32).reshape(-1, 32)
        n = pks.shape[0]
        bm = np.zeros(max(1, (n + 31) // 32), dtype=np.uint32)
        self._check(self.lib.hs_committee_register(self.h, _ptr(pks) if n else None, n, _ptr(bm)), "hs_committee_register")
        self.n_keys = n
        return bitmap_to_bools(bm, n)

    def committee_update(self, add=None, remove=None):
        """Incremental epoch change (hs_committee_update): returns the indices assigned to the added keys."""
        add = np.zeros((0, 32), np.uint8) if add is None else _u8(add, 32).reshape(-1, 32)
        rem = np.zeros(0, np.uint32) if remove is None else np.ascontiguousarray(remove, dtype=np.uint32)
        out = np.zeros(max(1, add.shape[0]), dtype=np.uint32)
        self._check(self.lib.hs_committee_update(self.h, _ptr(add) if add.shape[0] else None, add.shape[0], _ptr(rem) if rem.shape[0] else None,
                                                 rem.shape[0], _ptr(out)), "hs_committee_update")
        return out[: add.shape[0]]

    def set_table_budget(self, nbytes):
        self._check(self.lib.hs_set_table_budget(self.h, int(nbytes)), "hs_set_table_budget")

    def verify_committee(self, validator_idx, sig, digests, msg_idx=None, mode=MODE_STRICT):
        vidx = np.ascontiguousarray(validator_idx, dtype=np.uint32)
        sig = _u8(sig, 64).reshape(-1, 64)
        digests = _u8(digests, 32).reshape(-1, 32)
        n = sig.shape[0]
        assert vidx.shape[0] == n
        midx = None if msg_idx is None else np.ascontiguousarray(msg_idx, dtype=np.uint32)
        bm = np.zeros((n + 31) // 32, dtype=np.uint32)
        self._check(self.lib.hs_verify_committee(self.h, _ptr(vidx), _ptr(sig), _ptr(midx), _ptr(digests), digests.shape[0], n, mode, _ptr(bm)),
                    "hs_verify_committee")
        return bitmap_to_bools(bm, n)

    def digest32_batch(self, data, off):
        data = _u8(data)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        n = off.shape[0] - 1
        out = np.zeros((n, 32), dtype=np.uint8)
        self._check(self.lib.hs_digest32_batch(self.h, _ptr(data) if data.size else None, _ptr(off), n, _ptr(out)), "hs_digest32_batch")
        return out

    def verify_msgs(self, sig, msgs, msg_len, pk=None, validator_idx=None, mode=MODE_STRICT):
        """Reference-shaped call: verdict_i = Signature::verify(Digest(msg_i), key_i); msgs = n fixed-size messages."""
        sig = _u8(sig, 64).reshape(-1, 64)
        n = sig.shape[0]
        msgs = _u8(msgs)
        assert msgs.size == n * msg_len and (pk is None) != (validator_idx is None)
        pk = None if pk is None else _u8(pk, 32).reshape(-1, 32)
        vidx = None if validator_idx is None else np.ascontiguousarray(validator_idx, dtype=np.uint32)
        bm = np.zeros((n + 31) // 32, dtype=np.uint32)
        self._check(self.lib.hs_verify_msgs(self.h, _ptr(sig), _ptr(pk), _ptr(vidx), _ptr(msgs), msg_len, n, mode, _ptr(bm)), "hs_verify_msgs")
        return bitmap_to_bools(bm, n)

    # ---- device-resident API (torch tensors on this engine's device; enqueued on torch's current stream) -------
    @staticmethod
    def _stream():
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def verify_rec128_dev(self, d_recs, d_bitmap, n, mode=MODE_STRICT):
        self._check(self.lib.hs_verify_rec128_dev(self.h, d_recs.data_ptr(), n, mode, d_bitmap.data_ptr(), self._stream()), "hs_verify_rec128_dev")

    def verify_var_dev(self, d_sig, d_pk, d_msgs, d_off, d_bitmap, n, mode=MODE_STRICT):
        self._check(self.lib.hs_verify_var_dev(self.h, d_sig.data_ptr(), d_pk.data_ptr(), d_msgs.data_ptr(), d_off.data_ptr(), n, mode,
                                               d_bitmap.data_ptr(), self._stream()), "hs_verify_var_dev")

    def verify_committee_dev(self, d_vidx, d_sig, d_digests, d_bitmap, n, d_midx=None, mode=MODE_STRICT):
        self._check(self.lib.hs_verify_committee_dev(self.h, d_vidx.data_ptr(), d_sig.data_ptr(), None if d_midx is None else d_midx.data_ptr(),
                                                     d_digests.data_ptr(), n, mode, d_bitmap.data_ptr(), self._stream()),
                    "hs_verify_committee_dev")

    def verify_msgs_dev(self, d_sig, d_msgs, msg_len, d_digests, d_bitmap, n, d_pk=None, d_vidx=None, mode=MODE_STRICT):
        self._check(self.lib.hs_verify_msgs_dev(self.h, d_sig.data_ptr(), None if d_pk is None else d_pk.data_ptr(),
                                                None if d_vidx is None else d_vidx.data_ptr(), d_msgs.data_ptr(), msg_len, n, mode,
                                                d_digests.data_ptr(), d_bitmap.data_ptr(), self._stream()), "hs_verify_msgs_dev")

    def verify_qc_votes_dev(self, d_qc_digests, d_sig, d_qc_idx, d_vote_bitmap, n, d_pk=None, d_vidx=None):
        self._check(self.lib.hs_verify_qc_votes_dev(self.h, d_qc_digests.data_ptr(), None if d_pk is None else d_pk.data_ptr(),
                                                    None if d_vidx is None else d_vidx.data_ptr(), d_sig.data_ptr(), d_qc_idx.data_ptr(), n,
                                                    d_vote_bitmap.data_ptr(), self._stream()), "hs_verify_qc_votes_dev")

    def qc_and_dev(self, d_vote_bitmap, d_qc_idx, n_votes, n_qc, d_qc_bitmap):
        self._check(self.lib.hs_qc_and_dev(self.h, d_vote_bitmap.data_ptr(), d_qc_idx.data_ptr(), n_votes, n_qc, d_qc_bitmap.data_ptr(), self._stream()),
                    "hs_qc_and_dev")

    def keygen_batch_dev(self, d_seeds, d_pks, n):
        self._check(self.lib.hs_keygen_batch_dev(self.h, d_seeds.data_ptr(), n, d_pks.data_ptr(), self._stream()), "hs_keygen_batch_dev")

    def sign_digests_dev(self, d_seeds, d_pks, n_keys, d_digests, d_sig, n, d_key_idx=None):
        self._check(self.lib.hs_sign_digests_dev(self.h, d_seeds.data_ptr(), d_pks.data_ptr(), n_keys, None if d_key_idx is None else d_key_idx.data_ptr(),
                                                 d_digests.data_ptr(), n, d_sig.data_ptr(), self._stream()), "hs_sign_digests_dev")

    def set_deferred(self, on):
        """Deferred-results mode for streams of `_dev` passes (hs_set_deferred): call results_wait() before reading bitmaps."""
        self._check(self.lib.hs_set_deferred(self.h, 1 if on else 0), "hs_set_deferred")

    def results_wait(self):
        self._check(self.lib.hs_results_wait(self.h, self._stream()), "hs_results_wait")

    def digest32_fixed_dev(self, d_msgs, msg_len, d_out, n):
        self._check(self.lib.hs_digest32_fixed_dev(self.h, d_msgs.data_ptr(), msg_len, n, d_out.data_ptr(), self._stream()), "hs_digest32_fixed_dev")

    def digest32_dev(self, d_data, d_off, d_out, n):
        self._check(self.lib.hs_digest32_dev(self.h, d_data.data_ptr(), d_off.data_ptr(), n, d_out.data_ptr(), self._stream()), "hs_digest32_dev")
