// Host-side check of the supporting-surfel arrival key (csrc/sm_kernels.cuh: tb_encode / supporting_index):
// decode(encode(slot)) == slot, the key never collides with the invalid marker, and the multiply-based
// division (Barrett) agrees with the plain formulas. Built and run by tests/test_tiebreak_host.py.
#include <cstdio>
#include <cstdlib>
#include "sm_handle.cuh"
using namespace smb;
namespace smb { int SetError(int c, const char*) { return c; } }
int main() {
  TieBreakConfig cfg{};
  // host re-implementation of SetTieBreakWave pieces
  for (int variant = 0; variant < 8; ++variant) {
    const u32 waves[8] = {303104u, 303104u, 1u << 30, 7u, 1000003u, 4096u, 303104u, 1000003u};
    const u32 shifts[8] = {5, 0, 5, 0, 0, 10, 5, 0};
    const u32 offsets[8] = {0, 0, 0, 0, 0, 0, 1, 1};
    const u32 wave = waves[variant], shift = shifts[variant], groups = wave >> shift;
    cfg.wave = wave; cfg.early_fraction = 0.3; cfg.index_order_fraction = 0.44;
    const unsigned long long prime = 2654435761ull; cfg.mul = prime % groups; if (!cfg.mul) cfg.mul = 1;
    // modular inverse
    long long t = 0, nt = 1, r = groups, nr = cfg.mul % groups;
    while (nr) { long long q = r / nr; long long tmp = t - q * nt; t = nt; nt = tmp; tmp = r - q * nr; r = nr; nr = tmp; }
    if (t < 0) t += groups; cfg.mul_inv = (u32)t;
    TieBreak tb{}; tb.wave = wave; tb.lane_shift = shift; tb.groups = groups; tb.mul = cfg.mul; tb.mul_inv = cfg.mul_inv;
    tb.add = 12345 % groups; tb.salt = 0xdeadbeef;
    tb.early_threshold = 0x40000000u; tb.index_order_threshold = 0x70000000u; tb.wave_reciprocal = ~0ull / wave;
    tb.early_threshold_later = 0x60000000u; tb.index_order_threshold_later = 0x30000000u; tb.early_threshold_second = 0x20000000u;
    tb.group_reciprocal = ~0ull / groups;
    tb.wave_offset = offsets[variant];
    unsigned long long bad = 0, n = 0;
    for (u32 idx = 0; idx < 20000000u; idx += 7) for (int sec = 0; sec < 2; ++sec) {
      const u32 pixel = (idx * 2654435761u) % 307200u;
      const u32 key = tb_encode(tb, idx, sec, pixel);
      if (supporting_index(tb, key, pixel) != idx || key == 0xFFFFFFFFu) ++bad;
      // reference formulas with real division
      const u32 phase = tb.wave_offset ? ((tb_hash(pixel ^ tb.salt ^ 0x5bd1e995u) % groups) << shift) : 0u;
      u32 w = (u32)(((u64)idx + phase) / wave), rr = (u32)(((u64)idx + phase) % wave);
      u32 rp = tb_index_order(tb, pixel, w) ? rr
                                         : (u32)(((((u64)(rr >> shift) * tb.mul + tb.add) % groups) << shift) | (rr & ((1u << shift) - 1)));
      bool late = sec && !(tb_hash(idx ^ tb.salt) < (w == 0 ? tb.early_threshold : (w == 1 ? tb.early_threshold_second : tb.early_threshold_later)));
      u32 expect = w * (2u * wave) + (late ? wave : 0u) + rp;
      if (wave < (1u<<30) || idx < wave) if (expect != key) ++bad;
      ++n;
    }
    // the lanes of one group keep their order under the shuffle
    if (shift) for (u32 idx = 0; idx + 1 < 5000000u; idx += 13) {
      if (tb.wave_offset) break;   // (the phase moves whole groups: covered by the variants without it)
      if ((idx % wave) >> shift != ((idx + 1) % wave) >> shift || idx / wave != (idx + 1) / wave) continue;
      for (u32 pixel = 0; pixel < 3; ++pixel) if (!(tb_encode(tb, idx, false, pixel) < tb_encode(tb, idx + 1, false, pixel))) ++bad;
    }
    printf("wave %u lanes %u offset %u: %llu keys, %llu bad\n", wave, 1u << shift, tb.wave_offset, n, bad);
    if (bad) return 1;
  }
  return 0;
}
