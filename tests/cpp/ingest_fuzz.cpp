// Mutation fuzzer for the wire ingest (hotstuff_b200/csrc/hs_ingest.cpp parses UNTRUSTED network bytes).  Built with
// -fsanitize=address,undefined by tests/test_wire_ingest.py and run on the re-serialised reference fixtures: every mutated frame must
// either parse into items that stay inside the output capacities or be reported malformed — never read or write out of bounds.
// usage: ingest_fuzz <seed-frames-file> <iterations>      file = repeated { u32 length, bytes }
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/hs_crypto.h"

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() {
  rng_state ^= rng_state << 13;
  rng_state ^= rng_state >> 7;
  rng_state ^= rng_state << 17;
  return rng_state;
}

int main(int argc, char **argv) {
  if (argc != 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<std::vector<uint8_t>> seeds;
  uint32_t len;
  while (fread(&len, 4, 1, f) == 1) {
    std::vector<uint8_t> b(len);
    if (len && fread(b.data(), 1, len, f) != len) return 2;
    seeds.push_back(b);
  }
  fclose(f);
  const long iters = atol(argv[2]);
  long malformed = 0, parsed = 0, items = 0;
  for (long it = 0; it < iters; it++) {
    // a batch of 1..4 frames, each a seed with 0..6 mutations (bit flips, byte sets, truncation, extension, length-field edits)
    const int nf = 1 + (int)(rnd() % 4);
    std::vector<uint8_t> blob;
    std::vector<uint64_t> off(1, 0);
    for (int k = 0; k < nf; k++) {
      std::vector<uint8_t> fr = seeds[rnd() % seeds.size()];
      const int nm = (int)(rnd() % 7);
      for (int m = 0; m < nm && !fr.empty(); m++) {
        switch (rnd() % 6) {
          case 0: fr[rnd() % fr.size()] ^= (uint8_t)(1u << (rnd() % 8)); break;
          case 1: fr[rnd() % fr.size()] = (uint8_t)rnd(); break;
          case 2: fr.resize(rnd() % (fr.size() + 1)); break;
          case 3: fr.insert(fr.end(), (size_t)(rnd() % 40), (uint8_t)rnd()); break;
          case 4: {  // overwrite 8 bytes with an extreme length value
            if (fr.size() >= 8) {
              const uint64_t v = (rnd() & 1) ? ~0ull >> (rnd() % 40) : rnd() % 100000;
              memcpy(fr.data() + rnd() % (fr.size() - 7), &v, 8);
            }
            break;
          }
          default: {
            if (fr.size() >= 4) {
              const uint32_t v = (uint32_t)(rnd() % 7);
              memcpy(fr.data(), &v, 4);  // enum tag
            }
          }
        }
      }
      blob.insert(blob.end(), fr.begin(), fr.end());
      off.push_back(blob.size());
    }
    // deliberately tight, exactly-sized heap buffers: ASAN sees any overrun
    const size_t cap_items = (size_t)(rnd() % 64), cap_msgs = (size_t)(rnd() % 48), cap_pre = (size_t)(rnd() % 4096);
    std::vector<uint8_t> sig(cap_items * 64 + 1), pk(cap_items * 32 + 1), mode(cap_items + 1), pre(cap_pre + 1);
    std::vector<uint32_t> mi(cap_items + 1), gi(cap_items + 1);
    std::vector<uint64_t> po(cap_msgs + 1);
    std::vector<hs_frame_info> info(nf);
    hs_ingest_out o;
    memset(&o, 0, sizeof(o));
    o.cap_items = cap_items; o.cap_msgs = cap_msgs; o.cap_pre_bytes = cap_pre;
    o.sig = sig.data(); o.pk = pk.data(); o.msg_idx = mi.data(); o.group_idx = gi.data(); o.mode = mode.data(); o.preimages = pre.data(); o.pre_off = po.data();
    const std::vector<uint8_t> exact(blob);   // exactly-sized copy: reads past the end of the input are caught too
    const int rc = hs_ingest_consensus_frames(exact.empty() ? (const uint8_t *)"" : exact.data(), off.data(), (size_t)nf, info.data(), &o);
    if (rc != HS_OK && rc != HS_ERR_NOMEM) return 3;
    for (int k = 0; k < nf; k++) {
      if (info[k].kind == HS_FRAME_MALFORMED) malformed++;
      else parsed++;
      if (info[k].kind != HS_FRAME_MALFORMED && info[k].kind > 4) return 4;
    }
    if (rc == HS_OK) {
      if (o.n_items > cap_items || o.n_msgs > cap_msgs || o.pre_bytes > cap_pre) return 5;
      for (size_t i = 0; i < o.n_items; i++)
        if (mi[i] >= o.n_msgs || gi[i] >= (uint32_t)nf || mode[i] > 1) return 6;
      for (size_t m = 0; m < o.n_msgs; m++)
        if (po[m] > po[m + 1] || po[m + 1] > o.pre_bytes) return 7;
      items += (long)o.n_items;
    }
  }
  printf("ingest fuzz ok: %ld iterations, %ld frames parsed, %ld malformed, %ld items\n", iters, parsed, malformed, items);
  return 0;
}
