// Host check of the TMA imfilter kernel's incremental tile iterator (image.cu, block "imf_tile_iter"). Test infrastructure:
// tests/test_host_logic.py cuts the struct out of image.cu into imf_tile_iter.inc, compiles this file and runs it. For a set of
// tile grids and grid sizes every CTA's walk must equal the division-based decomposition of t = b + k * stride, and the CTAs
// together must visit every tile exactly once.
#include <cstdint>
#include <cstdio>
#include <set>
#include <tuple>
#define __device__
#define __forceinline__ inline
#include "imf_tile_iter.inc"

int main() {
  int bad = 0;
  const uint32_t cfgs[][4] = {{60, 34, 3, 444}, {60, 34, 3, 296}, {60, 68, 3, 592}, {7, 100, 2, 592}, {100, 3, 5, 148}, {1, 1500, 1, 444},
                              {61, 20, 1, 1184}, {5, 7, 40, 1184}, {64, 2, 9, 1000}, {2, 700, 1, 444}, {1500, 1, 1, 444}, {1, 1, 1200, 1184},
                              {444, 3, 1, 444}, {37, 12, 3, 444}, {60, 17, 3, 445}};
  for (const auto& c : cfgs) {
    const uint32_t ntx = c[0], nty = c[1], planes = c[2], nt = ntx * nty * planes, G = c[3] > nt ? nt : c[3];
    std::set<std::tuple<uint32_t, uint32_t, uint32_t>> seen;
    for (uint32_t b = 0; b < G; ++b) {
      ImfTileIter it;
      it.init(b, G, ntx, nty);
      const uint32_t mine = (nt - b + G - 1) / G;
      for (uint32_t k = 0; k < mine; ++k, it.next()) {
        const uint32_t t = b + k * G;
        if (it.tx != t % ntx || it.ty != (t / ntx) % nty || it.plane != t / ntx / nty) ++bad;
        if (!seen.insert({it.tx, it.ty, it.plane}).second) ++bad;
      }
    }
    if (seen.size() != nt) ++bad;
    printf("ntx=%u nty=%u planes=%u grid=%u: %zu of %u tiles, mismatches so far %d\n", ntx, nty, planes, G, seen.size(), nt, bad);
  }
  return bad != 0;
}
