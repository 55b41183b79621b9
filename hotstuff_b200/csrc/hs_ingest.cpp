// hs_ingest.cpp — wire-format ingest: bincode-serialized ConsensusMessage frames -> flat arrays for hs_verify_groups.
//
// The reference hands every received frame to `bincode::deserialize::<ConsensusMessage>` (consensus/src/consensus.rs:138), builds
// Block / Vote / Timeout / TC structs and only then walks them, one signature at a time.  At 10^7+ verifies/s the per-vote
// host work (base64 key strings, Vec allocations) is the bottleneck, so this parser goes straight from frame bytes to the
// structure-of-arrays the GPU entry point consumes: signatures, key bytes, digest preimages, item -> message / group maps.
//
// bincode 1.3 `serialize` / `deserialize` (fixed-width little-endian integers, u64 lengths, u32 enum tags, trailing bytes allowed):
//   ConsensusMessage (consensus.rs:33-39)  u32 tag: 0 Propose(Block) 1 Vote 2 Timeout 3 TC 4 SyncRequest(Digest, PublicKey)
//   Block   (messages.rs:17-24)   qc: QC | tc: Option<TC> (u8 tag) | author: PublicKey | round: u64 | payload: Vec<Digest> | signature
//   Vote    (messages.rs:104-110) hash: Digest | round | author | signature
//   QC      (messages.rs:165-169) hash | round | votes: Vec<(PublicKey, Signature)>
//   Timeout (messages.rs:223-228) high_qc: QC | round | author | signature
//   TC      (messages.rs:283-287) round | votes: Vec<(PublicKey, Signature, Round)>
//   Digest = 32 raw bytes (crypto/src/lib.rs:22); Signature = 32 + 32 raw bytes (:178-182);
//   PublicKey = String holding standard base64 (:94-112): u64 length, bytes; decode, first 32 bytes of the result.
// Host-only code (no CUDA); part of libhs_crypto.so.
#include <cstdint>
#include <cstring>

#include "../../include/hs_crypto.h"

namespace {

struct reader {
  const uint8_t *p, *end;
  bool ok = true;
  bool need(uint64_t n) {
    if (!ok || (uint64_t)(end - p) < n) ok = false;
    return ok;
  }
  uint64_t u64() {
    if (!need(8)) return 0;
    uint64_t v;
    memcpy(&v, p, 8);
    p += 8;
    return v;
  }
  uint32_t u32() {
    if (!need(4)) return 0;
    uint32_t v;
    memcpy(&v, p, 4);
    p += 4;
    return v;
  }
  uint8_t u8() {
    if (!need(1)) return 0;
    return *p++;
  }
  const uint8_t *bytes(uint64_t n) {
    if (!need(n)) return nullptr;
    const uint8_t *q = p;
    p += n;
    return q;
  }
};

// standard base64 alphabet -> 6-bit value, -1 for everything else ('=' included).  A table: keys are random bytes, so a compare
// ladder mispredicts on every other character (measured: 760 ns per vote with the ladder, see tools/ingest_bench.cpp).
struct b64_table {
  int8_t v[256];
  constexpr b64_table() : v() {
    for (int i = 0; i < 256; i++) v[i] = -1;
    for (int i = 0; i < 26; i++) {
      v['A' + i] = (int8_t)i;
      v['a' + i] = (int8_t)(26 + i);
    }
    for (int i = 0; i < 10; i++) v['0' + i] = (int8_t)(52 + i);
    v[(int)'+'] = 62;
    v[(int)'/'] = 63;
  }
};
constexpr b64_table B64{};
inline int b64val(uint8_t c) { return B64.v[c]; }
// PublicKey::deserialize (crypto/src/lib.rs:103-112): String, base64::decode (standard alphabet, canonical padding, no
// trailing bits), then bytes[..32].  A decoded length < 32 makes the reference's slice expression panic; here it is a
// malformed frame.
bool read_public_key(reader &r, uint8_t out[32]) {
  const uint64_t len = r.u64();
  const uint8_t *s = r.bytes(len);
  if (!s || len % 4 != 0 || len < 44) return r.ok = false;
  if (len == 44 && s[43] == '=') {
    // The one canonical spelling of a 32-byte key — 43 symbols and one '=' — without per-character branches: ten full quanta, then
    // three symbols whose last two bits must be zero.  Any foreign character (a second '=' included) turns `bad` negative.
    int bad = 0;
    for (int q = 0; q < 10; q++) {
      const int a = B64.v[s[4 * q]], b = B64.v[s[4 * q + 1]], c = B64.v[s[4 * q + 2]], d = B64.v[s[4 * q + 3]];
      bad |= a | b | c | d;
      const uint32_t ua = (uint32_t)a, ub = (uint32_t)b, uc = (uint32_t)c, ud = (uint32_t)d;  // (shifts on unsigned: a foreign character is -1)
      out[3 * q] = (uint8_t)((ua << 2) | (ub >> 4));
      out[3 * q + 1] = (uint8_t)((ub << 4) | (uc >> 2));
      out[3 * q + 2] = (uint8_t)((uc << 6) | ud);
    }
    const int a = B64.v[s[40]], b = B64.v[s[41]], c = B64.v[s[42]];
    bad |= a | b | c;
    const uint32_t ua = (uint32_t)a, ub = (uint32_t)b, uc = (uint32_t)c;
    out[30] = (uint8_t)((ua << 2) | (ub >> 4));
    out[31] = (uint8_t)((ub << 4) | (uc >> 2));
    if (bad < 0 || (c & 3)) return r.ok = false;
    return true;
  }
  uint64_t n_out = len / 4 * 3;
  if (s[len - 1] == '=') n_out--;
  if (s[len - 2] == '=') n_out--;
  if (n_out < 32) return r.ok = false;
  uint8_t buf[3];
  uint64_t produced = 0;
  for (uint64_t i = 0; i < len; i += 4) {
    int v[4];
    int pad = 0;
    for (int k = 0; k < 4; k++) {
      const uint8_t c = s[i + k];
      if (c == '=') {
        if (i + 4 != len || k < 2) return r.ok = false;  // padding only in the last quantum, at most two
        v[k] = 0;
        pad++;
      } else {
        if (pad) return r.ok = false;  // data after padding
        v[k] = b64val(c);
        if (v[k] < 0) return r.ok = false;
      }
    }
    if (pad == 1 && (v[2] & 3)) return r.ok = false;   // non-zero trailing bits: InvalidLastSymbol
    if (pad == 2 && (v[1] & 15)) return r.ok = false;
    buf[0] = (uint8_t)((v[0] << 2) | (v[1] >> 4));
    buf[1] = (uint8_t)((v[1] << 4) | (v[2] >> 2));
    buf[2] = (uint8_t)((v[2] << 6) | v[3]);
    for (int k = 0; k < 3 - pad; k++) {
      if (produced < 32) out[produced] = buf[k];
      produced++;
    }
  }
  return produced >= 32;
}

struct sink {
  hs_ingest_out *o;
  size_t n_items = 0, n_msgs = 0, pre_bytes = 0;
  bool overflow = false;
  // returns the message index of a new preimage of `len` bytes and a pointer to fill (nullptr when over capacity)
  uint8_t *new_msg(size_t len, uint32_t &idx) {
    idx = (uint32_t)n_msgs;
    uint8_t *dst = nullptr;
    if (n_msgs < o->cap_msgs && pre_bytes + len <= o->cap_pre_bytes) {
      dst = o->preimages + pre_bytes;
      o->pre_off[n_msgs] = pre_bytes;
      o->pre_off[n_msgs + 1] = pre_bytes + len;
    } else {
      overflow = true;
    }
    n_msgs++;
    pre_bytes += len;
    return dst;
  }
  void item(const uint8_t *sig, const uint8_t *pk, uint32_t msg, uint32_t group, uint8_t mode) {
    if (n_items < o->cap_items) {
      memcpy(o->sig + 64 * n_items, sig, 64);
      memcpy(o->pk + 32 * n_items, pk, 32);
      o->msg_idx[n_items] = msg;
      o->group_idx[n_items] = group;
      o->mode[n_items] = mode;
    } else {
      overflow = true;
    }
    n_items++;
  }
};

struct qc_view {
  const uint8_t *hash;
  uint64_t round;
  uint64_t n_votes;
  const uint8_t *votes;  // start of the serialized vote list
  bool genesis;
};
// QC: hash | round | Vec<(PublicKey, Signature)>.  Remembers where the votes start; they are re-walked when emitted.
bool skip_qc(reader &r, qc_view &q) {
  q.hash = r.bytes(32);
  q.round = r.u64();
  q.n_votes = r.u64();
  q.votes = r.p;
  if (!r.ok || q.n_votes > (uint64_t)(r.end - r.p) / (8 + 44 + 64)) return r.ok = false;
  uint8_t pk[32];
  for (uint64_t i = 0; i < q.n_votes; i++) {
    if (!read_public_key(r, pk)) return false;
    if (!r.bytes(64)) return false;
  }
  static const uint8_t zero[32] = {0};
  q.genesis = q.round == 0 && memcmp(q.hash, zero, 32) == 0;  // QC::genesis() = QC::default(); PartialEq compares hash and round (messages.rs:216-220)
  return r.ok;
}
// QC::verify (messages.rs:180-198): every vote under the verify_batch condition over QC::digest = SHA-512(hash || round_le)[..32]
void emit_qc(sink &s, const qc_view &q, const uint8_t *end, uint32_t group, uint32_t &lo, uint32_t &hi) {
  uint32_t m;
  if (uint8_t *dst = s.new_msg(40, m)) {
    memcpy(dst, q.hash, 32);
    memcpy(dst + 32, &q.round, 8);
  }
  reader r{q.votes, end};
  lo = (uint32_t)s.n_items;
  uint8_t pk[32];
  for (uint64_t i = 0; i < q.n_votes; i++) {
    read_public_key(r, pk);
    s.item(r.bytes(64), pk, m, group, HS_MODE_BATCH_EQ);
  }
  hi = (uint32_t)s.n_items;
}
// TC: round | Vec<(PublicKey, Signature, Round)>; TC::verify (messages.rs:290-315): strict verify per vote over
// SHA-512(round_le || high_qc_round_le)[..32]
bool parse_tc(reader &r, sink *s, uint32_t group, uint64_t &round, uint32_t &lo, uint32_t &hi) {
  round = r.u64();
  const uint64_t n = r.u64();
  if (!r.ok || n > (uint64_t)(r.end - r.p) / (8 + 44 + 64 + 8)) return r.ok = false;
  if (s) lo = (uint32_t)s->n_items;
  uint8_t pk[32];
  for (uint64_t i = 0; i < n; i++) {
    if (!read_public_key(r, pk)) return false;
    const uint8_t *sig = r.bytes(64);
    const uint64_t hq = r.u64();
    if (!r.ok) return false;
    if (s) {
      uint32_t m;
      if (uint8_t *dst = s->new_msg(16, m)) {
        memcpy(dst, &round, 8);
        memcpy(dst + 8, &hq, 8);
      }
      s->item(sig, pk, m, group, HS_MODE_STRICT);
    }
  }
  if (s) hi = (uint32_t)s->n_items;
  return true;
}

// One frame.  Two passes over composite messages: validate the whole frame first (a malformed frame must contribute NOTHING),
// then emit.
bool parse_frame(const uint8_t *p, const uint8_t *end, sink &s, uint32_t group, hs_frame_info &fi) {
  reader r{p, end};
  const uint32_t tag = r.u32();
  if (!r.ok || tag > 4) return false;
  fi.kind = (uint8_t)tag;
  uint8_t author[32];
  if (tag == 0) {  // Propose(Block)
    qc_view q;
    if (!skip_qc(r, q)) return false;
    const uint8_t has_tc = r.u8();
    if (!r.ok || has_tc > 1) return false;
    const uint8_t *tc_at = r.p;
    uint64_t tc_round = 0;
    uint32_t d0, d1;
    if (has_tc && !parse_tc(r, nullptr, 0, tc_round, d0, d1)) return false;
    if (!read_public_key(r, author)) return false;
    const uint64_t round = r.u64();
    const uint64_t n_payload = r.u64();
    if (!r.ok || n_payload > (uint64_t)(r.end - r.p) / 32) return false;
    const uint8_t *payload = r.bytes(n_payload * 32);
    const uint8_t *sig = r.bytes(64);
    if (!r.ok) return false;
    // emit: author signature over Block::digest preimage = author || round_le || payload digests || qc.hash (messages.rs:79-90)
    uint32_t m;
    if (uint8_t *dst = s.new_msg(32 + 8 + n_payload * 32 + 32, m)) {
      memcpy(dst, author, 32);
      memcpy(dst + 32, &round, 8);
      if (n_payload) memcpy(dst + 40, payload, n_payload * 32);
      memcpy(dst + 40 + n_payload * 32, q.hash, 32);
    }
    fi.author_item = (uint32_t)s.n_items;
    s.item(sig, author, m, group, HS_MODE_STRICT);
    fi.round = round;
    fi.qc_is_genesis = q.genesis;
    fi.qc_round = q.round;
    if (!q.genesis) emit_qc(s, q, end, group, fi.qc_lo, fi.qc_hi);
    fi.has_tc = has_tc;
    if (has_tc) {
      reader rt{tc_at, end};
      parse_tc(rt, &s, group, tc_round, fi.tc_lo, fi.tc_hi);
      fi.tc_round = tc_round;
    }
    return true;
  }
  if (tag == 1) {  // Vote: hash | round | author | signature; digest preimage = hash || round_le (messages.rs:149-156)
    const uint8_t *hash = r.bytes(32);
    const uint64_t round = r.u64();
    if (!read_public_key(r, author)) return false;
    const uint8_t *sig = r.bytes(64);
    if (!r.ok) return false;
    uint32_t m;
    if (uint8_t *dst = s.new_msg(40, m)) {
      memcpy(dst, hash, 32);
      memcpy(dst + 32, &round, 8);
    }
    fi.author_item = (uint32_t)s.n_items;
    s.item(sig, author, m, group, HS_MODE_STRICT);
    fi.round = round;
    return true;
  }
  if (tag == 2) {  // Timeout: high_qc | round | author | signature; digest preimage = round_le || high_qc.round_le (messages.rs:268-275)
    qc_view q;
    if (!skip_qc(r, q)) return false;
    const uint64_t round = r.u64();
    if (!read_public_key(r, author)) return false;
    const uint8_t *sig = r.bytes(64);
    if (!r.ok) return false;
    uint32_t m;
    if (uint8_t *dst = s.new_msg(16, m)) {
      memcpy(dst, &round, 8);
      memcpy(dst + 8, &q.round, 8);
    }
    fi.author_item = (uint32_t)s.n_items;
    s.item(sig, author, m, group, HS_MODE_STRICT);
    fi.round = round;
    fi.qc_is_genesis = q.genesis;
    fi.qc_round = q.round;
    if (!q.genesis) emit_qc(s, q, end, group, fi.qc_lo, fi.qc_hi);
    return true;
  }
  if (tag == 3) {  // TC
    reader probe = r;
    uint64_t round = 0;
    uint32_t d0, d1;
    if (!parse_tc(probe, nullptr, 0, round, d0, d1)) return false;
    fi.author_item = HS_NO_ITEM;
    fi.has_tc = 1;
    parse_tc(r, &s, group, round, fi.tc_lo, fi.tc_hi);
    fi.tc_round = fi.round = round;
    return true;
  }
  // SyncRequest(Digest, PublicKey): nothing to verify
  if (!r.bytes(32) || !read_public_key(r, author)) return false;
  fi.author_item = HS_NO_ITEM;
  return true;
}

}  // namespace

extern "C" int hs_ingest_consensus_frames(const uint8_t *frames, const uint64_t *off, size_t n, hs_frame_info *info, hs_ingest_out *out) {
  if (!out || (n && (!off || !info)) || (n && off[n] && !frames)) return HS_ERR_ARG;
  if (n == 0) {
    out->n_items = out->n_msgs = out->pre_bytes = 0;
    if (out->pre_off) out->pre_off[0] = 0;
    return HS_OK;
  }
  if (out->cap_items && (!out->sig || !out->pk || !out->msg_idx || !out->group_idx || !out->mode)) return HS_ERR_ARG;
  if (out->cap_msgs && (!out->pre_off || (out->cap_pre_bytes && !out->preimages))) return HS_ERR_ARG;
  if (!out->pre_off) return HS_ERR_ARG;
  for (size_t i = 0; i < n; i++)
    if (off[i] > off[i + 1]) return HS_ERR_ARG;
  sink s{out};
  out->pre_off[0] = 0;
  for (size_t i = 0; i < n; i++) {
    hs_frame_info fi;
    memset(&fi, 0, sizeof(fi));
    fi.author_item = HS_NO_ITEM;
    const sink before = s;
    if (!parse_frame(frames + off[i], frames + off[i + 1], s, (uint32_t)i, fi)) {
      s = before;  // a malformed frame contributes nothing (the reference drops it: SerializationError, consensus.rs:138)
      memset(&fi, 0, sizeof(fi));
      fi.kind = HS_FRAME_MALFORMED;
      fi.author_item = HS_NO_ITEM;
    }
    info[i] = fi;
  }
  out->n_items = s.n_items;
  out->n_msgs = s.n_msgs;
  out->pre_bytes = s.pre_bytes;
  return s.overflow ? HS_ERR_NOMEM : HS_OK;  // on NOMEM the three counts say what the capacities must be
}
