// Host-only check (no GPU): the closed-form tile layout helpers that the persistent kernels use (block_meta_of,
// row_tile_start_of, block_tiles_of, s2k_legendre.cuh) against a direct restatement of the tabulated layout
// (build_layout, plan.cu; RowSize, cospml.c:250-258).  Built and run by tests/test_boundary.py.
#include <algorithm>
#include <cstdio>

#include "../../s2kit_b200/csrc/s2k_legendre.cuh"

using namespace s2k;

static int row_size(int m, int l) {
    if (l < m) return 0;
    return (m & 1) ? (l - 1) / 2 + 1 : l / 2 + 1;
}

int main() {
    int bad = 0;
    const int bws[] = {2, 3, 8, 16, 17, 23, 24, 64, 128, 256, 512, 1024, 2048};
    for (int bw : bws) {
        for (int m = 0; m < bw; ++m) {
            unsigned cur = 0;
            for (int par = 0; par < 2; ++par) {
                int rows = (bw - m - par + 1) / 2;
                if (rows < 0) rows = 0;
                const int len0 = row_size(m, m + par), nrt = (rows + 7) / 8;
                const BlockMeta mb = block_meta_of(m, par, bw);
                if (mb.rows != rows || mb.len0 != len0 || mb.nrt != nrt) ++bad;
                const unsigned base = cur;
                if (par == 1 && base != block_tiles_of(block_meta_of(m, 0, bw))) ++bad;
                for (int rt = 0; rt < nrt; ++rt) {
                    if (cur - base != row_tile_start_of(mb, rt)) ++bad;
                    const int last_row = std::min(8 * rt + 7, rows - 1);
                    cur += (unsigned)((len0 + last_row + 7) >> 3);
                }
            }
        }
    }
    std::printf("layout mismatches: %d\n", bad);
    return bad != 0;
}
