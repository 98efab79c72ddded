// Host model of the shared-memory wavefronts of the Hex8 tile kernel's FLUSH (hex8_tile_kernel.cuh, helper warps): builds the tile lists of
// a structured n^3 Hex8 mesh with the library's own builder (tiles.cpp) and replays the three LDS.64 of every 32-item flush group.
// A 64-bit shared load of a warp = two half-warps; each costs max over the 16 eight-byte banks of the distinct words it touches.
// Validation: the model gives 6.19 wavefronts per load for the shipped layout; ncu (profiles/r01/tile_v6_final_ncu_summary.txt, source
// page of the same capture) measured 5.6 - 6.6 (1.41 - 1.66 wavefronts per element at 0.25 executions per element).
//   g++ -O2 -std=c++17 -I/usr/local/cuda/include scripts/flush_bank_model.cpp fenris_b200/csrc/tiles.cpp -o /tmp/flush_model -lpthread
//   /tmp/flush_model <n> <variant> <layout>
//     variant 0 shipped order | 1 row index rotated by the entry number | 3 rotation chosen greedily per entry on the host |
//             5 complete entries permuted inside each 32-item group (local search; same global addresses per instruction) | 4 = 5 + 3
//     layout  0 shipped (pos * 9 + 3 i + j) | 1 row planes (i * PLANE + 3 pos + j, PLANE = 1 mod 16)
// Results (n = 32): see profiles/r01/README.md, "Flush bank-conflict model".
#include "../fenris_b200/csrc/fb200_internal.h"
#include <algorithm>
#include <set>
#include <map>
using namespace fb200;
static int g_layout = 0;
static inline int waddr(int pos, bool tr, int ii, int j) { const int PLANE = 3 * 1216 + 1; if (!g_layout) return pos * 9 + (tr ? j * 3 + ii : ii * 3 + j); return tr ? j * PLANE + 3 * pos + ii : ii * PLANE + 3 * pos + j; }
// wavefronts of one 64-bit shared load: two half-warps, each max over the 16 banks of the distinct words in it
static int wavefronts(const int* word, const bool* act) {
  int tot = 0;
  for (int h = 0; h < 2; ++h) {
    std::set<int> per[16]; bool any = false;
    for (int l = 16 * h; l < 16 * h + 16; ++l) if (act[l]) { per[word[l] & 15].insert(word[l]); any = true; }
    int m = 0; for (int b = 0; b < 16; ++b) m = std::max<int>(m, per[b].size());
    tot += any ? m : 0;
  }
  return tot;
}
int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 32; int variant = argc > 2 ? atoi(argv[2]) : 0; g_layout = argc > 3 ? atoi(argv[3]) : 0;
  int N1 = n + 1; uint64_t E = (uint64_t)n * n * n;
  std::vector<int32_t> conn(E * 8); std::vector<uint16_t> map(E * 64);
  auto nid = [&](int x, int y, int z) { return (int32_t)((z * N1 + y) * N1 + x); };
  auto rank = [&](int X, int Y, int Z, int VX, int VY, int VZ) {
    int r = 0;
    for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
      int x = X + dx, y = Y + dy, z = Z + dz; if (x < 0 || y < 0 || z < 0 || x > n || y > n || z > n) continue;
      if (x == VX && y == VY && z == VZ) return r; ++r; }
    return -1; };
  static const int off[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
  std::vector<std::pair<uint64_t, uint64_t>> keys(E);
  for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < n; ++x) {
    uint64_t e = ((uint64_t)z * n + y) * n + x;
    for (int a = 0; a < 8; ++a) conn[e * 8 + a] = nid(x + off[a][0], y + off[a][1], z + off[a][2]);
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) map[e * 64 + a * 8 + b] = rank(x + off[a][0], y + off[a][1], z + off[a][2], x + off[b][0], y + off[b][1], z + off[b][2]);
    uint64_t code = 0; for (int bt = 20; bt >= 0; --bt) { code = (code << 1) | ((z >> bt) & 1); code = (code << 1) | ((y >> bt) & 1); code = (code << 1) | ((x >> bt) & 1); }
    keys[e] = {code, e}; }
  std::sort(keys.begin(), keys.end());
  std::vector<int32_t> order(E); std::vector<uint64_t> codes(E);
  for (uint64_t i = 0; i < E; ++i) { order[i] = keys[i].second; codes[i] = keys[i].first; }
  TileShape sh{6, 64, 8, 128, 1216};
  HostTiles ht;
  build_tile_lists(sh, E, order.data(), codes.data(), conn.data(), E, (uint64_t)N1 * N1 * N1, map.data(), ht);
  size_t nt = ht.hdr.size() / 8;
  uint64_t instr = 0, wf = 0, ideal = 0;
  for (size_t t = 0; t < nt; ++t) {
    const uint32_t fb = ht.hdr[t * 8 + 5], nf = ht.hdr[t * 8 + 6];
    const uint32_t items = nf * 3;
    std::vector<int> rotv(nf, 0);
    if (variant == 3) {
      auto cost = [&](uint32_t e_last) {  // wavefronts of the half-warps touched by entry e_last, counting entries <= e_last only
        int tot = 0; uint32_t h0 = (e_last * 3) / 16, h1 = (e_last * 3 + 2) / 16;
        for (uint32_t h = h0; h <= h1; ++h) for (int i = 0; i < 3; ++i) {
          std::set<int> per[16];
          for (uint32_t it = h * 16; it < h * 16 + 16 && it / 3 <= e_last; ++it) {
            uint32_t a = ht.flush[fb + it / 3]; int j = it % 3; int pos = a & 0x7ff; bool tr = (a >> 11) & 1; int ii = (i + rotv[it / 3]) % 3;
            int w = waddr(pos, tr, ii, j); per[w & 15].insert(w); }
          int m = 0; for (int b = 0; b < 16; ++b) m = std::max<int>(m, per[b].size()); tot += m; }
        return tot; };
      for (uint32_t e = 0; e < nf; ++e) { int best = 0, bc = 1 << 30; for (int r = 0; r < 3; ++r) { rotv[e] = r; int c = cost(e); if (c < bc) { bc = c; best = r; } } rotv[e] = best; }
    }
    std::vector<uint32_t> fl(ht.flush.begin() + fb, ht.flush.begin() + fb + nf);
    if (variant == 4 || variant == 5) {
      auto group_cost = [&](uint32_t g) {
        int tot = 0;
        for (int i = 0; i < 3; ++i) { int word[32]; bool act[32];
          for (int l = 0; l < 32; ++l) { uint32_t it = g + l; act[l] = it < items; word[l] = 0; if (!act[l]) continue;
            uint32_t a = fl[it / 3]; int j = it % 3; int pos = a & 0x7ff; bool tr = (a >> 11) & 1; int ii = (i + rotv[it / 3]) % 3;
            word[l] = waddr(pos, tr, ii, j); }
          tot += wavefronts(word, act); }
        return tot; };
      for (uint32_t g = 0; g < items; g += 32) {
        uint32_t e0 = (g + 2) / 3, e1 = std::min<uint32_t>((g + 32) / 3, nf);  // entries wholly inside [g, g + 32)
        int best = group_cost(g);
        // local search: swaps of two complete entries (and rotations for variant 4) while the cost goes down
        for (int pass = 0; pass < 6; ++pass) { bool improved = false;
          for (uint32_t x = e0; x < e1; ++x) {
            for (uint32_t y = x + 1; y < e1; ++y) { std::swap(fl[x], fl[y]); std::swap(rotv[x], rotv[y]); int c = group_cost(g);
              if (c < best) { best = c; improved = true; } else { std::swap(fl[x], fl[y]); std::swap(rotv[x], rotv[y]); } }
            if (variant == 4) for (int r = 0; r < 3; ++r) { int old = rotv[x]; rotv[x] = r; int c = group_cost(g); if (c < best) { best = c; improved = true; } else rotv[x] = old; }
          }
          if (!improved) break; }
      }
    }
    for (uint32_t g = 0; g < items; g += 32) {
      for (int i = 0; i < 3; ++i) {
        int word[32]; bool act[32];
        for (int l = 0; l < 32; ++l) {
          uint32_t it = g + l; act[l] = it < items; word[l] = 0; if (!act[l]) continue;
          uint32_t a = fl[it / 3]; int j = it % 3; int pos = a & 0x7ff; bool tr = (a >> 11) & 1;
          int ii = i;
          if (variant == 4) ii = (i + rotv[it / 3]) % 3;
          if (variant == 1) ii = (i + (int)(it / 3)) % 3;           // row index rotated by the entry number
          if (variant == 3) ii = (i + rotv[it / 3]) % 3;
          if (variant == 2) ii = (i + (int)((a >> 19) & 0x1fff)) % 3;  // ... by the column position k
          word[l] = waddr(pos, tr, ii, j);
        }
        ++instr; wf += wavefronts(word, act); ideal += 2;
      }
    }
  }
  printf("n=%d variant=%d tiles=%zu flush entries/elem=%.2f  load instr/elem=%.3f  wavefronts/instr=%.3f  wavefronts/elem=%.2f (ideal %.2f)  accumulate conflict share=%.4f\n",
         n, variant, nt, (double)ht.flush.size() / E, (double)instr / E, (double)wf / instr, (double)wf / E, (double)ideal / E, ht.bank_conflict_share);
}
namespace fb200 { void morton_order(int, int, uint64_t, const double*, uint64_t, const uint64_t*, std::vector<int32_t>&, std::vector<uint64_t>&) {} }
