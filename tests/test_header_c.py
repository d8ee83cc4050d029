"""The drop-in boundary is a C header: include/mrp_b200.h must compile as plain C (gcc -std=c99)
and its inline accessors (mrp_packed_value, mrp_compact_value) must read fields the way the numpy
mirrors of the tests do.  No device needed: mrp_free_cell_index is host code."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r"""
#include <stdio.h>
#include <stdlib.h>
#include "mrp_b200.h"

int main(void) {
  /* 5 x 3 map, obstacles at (1,0) and (3,1); goal (0,0) */
  const int dimx = 5, dimy = 3;
  const int32_t obst[] = {1, 0, 3, 1};
  uint32_t bits[1];
  int32_t prefix[2];
  const int n_free = mrp_free_cell_index(dimx, dimy, obst, 2, bits, prefix);
  if (n_free != 13 || prefix[1] != 13) return 1;
  /* BFS distances from (0,0) by hand: row 0: 0 X 4 5 6 / row 1: 1 2 3 X 7 / row 2: 2 3 4 5 6 */
  const int32_t want[15] = {0, MRP_INF, 4, 5, 6, 1, 2, 3, MRP_INF, 7, 2, 3, 4, 5, 6};
  uint8_t packed[15], compact[13];
  int k = 0;
  for (int c = 0; c < 15; ++c) {
    const int x = c % dimx, y = c / dimx;
    packed[c] = want[c] == MRP_INF ? 255 : (uint8_t)((want[c] - x - y) / 2);
    if ((bits[0] >> c) & 1u) compact[k++] = packed[c];
  }
  if (k != n_free) return 2;
  for (int c = 0; c < 15; ++c) {
    const int x = c % dimx, y = c / dimx;
    if (mrp_packed_value(packed, dimx, x, y, 0, 0) != want[c]) return 3;
    if (mrp_compact_value(compact, bits, prefix, dimx, x, y, 0, 0) != want[c]) return 4;
  }
  /* a goal on an obstacle reads 0 at the goal itself */
  if (mrp_compact_value(compact, bits, prefix, dimx, 1, 0, 1, 0) != 0) return 5;
  printf("ok %d\n", mrp_max_lanes());
  return 0;
}
"""


def test_header_is_plain_c_and_accessors_work(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(PROG)
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "libmultirobotplanning_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", str(exe), "-L", libdir, "-lmrp_b200", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout, r.stderr)
