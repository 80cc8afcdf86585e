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
  for (u32 wave : {303104u, 1u << 30, 7u, 1000003u}) {
    cfg.wave = wave; cfg.early_fraction = 0.3; cfg.index_order_fraction = 0.44;
    const unsigned long long prime = 2654435761ull; cfg.mul = prime % wave; if (!cfg.mul) cfg.mul = 1;
    // modular inverse
    long long t = 0, nt = 1, r = wave, nr = cfg.mul % wave;
    while (nr) { long long q = r / nr; long long tmp = t - q * nt; t = nt; nt = tmp; tmp = r - q * nr; r = nr; nr = tmp; }
    if (t < 0) t += wave; cfg.mul_inv = (u32)t;
    TieBreak tb{}; tb.wave = wave; tb.mul = cfg.mul; tb.mul_inv = cfg.mul_inv; tb.add = 12345 % wave; tb.salt = 0xdeadbeef;
    tb.early_threshold = 0x40000000u; tb.index_order_threshold = 0x70000000u; tb.wave_reciprocal = ~0ull / wave;
    unsigned long long bad = 0, n = 0;
    for (u32 idx = 0; idx < 20000000u; idx += 7) for (int sec = 0; sec < 2; ++sec) {
      const u32 pixel = (idx * 2654435761u) % 307200u;
      const u32 key = tb_encode(tb, idx, sec, pixel);
      if (supporting_index(tb, key, pixel) != idx || key == 0xFFFFFFFFu) ++bad;
      // reference formulas with real division
      u32 w = idx / wave, rr = idx % wave;
      u32 rp = tb_index_order(tb, pixel) ? rr : (u32)(((u64)rr * tb.mul + tb.add) % wave);
      bool late = sec && !(tb_hash(idx ^ tb.salt) < tb.early_threshold);
      u32 expect = w * (2u * wave) + (late ? wave : 0u) + rp;
      if (wave < (1u<<30) || idx < wave) if (expect != key) ++bad;
      ++n;
    }
    printf("wave %u: %llu keys, %llu bad\n", wave, n, bad);
    if (bad) return 1;
  }
  return 0;
}
