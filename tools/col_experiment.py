#!/usr/bin/env python
"""Is the strided (COL) pass limited by the kernel or by DRAM access granularity?
Run the SAME COL kernel (tile = 4 or 8 pencils) on (a) the real 3-D layout
(pencil stride 16 KiB / 16 MiB) and (b) a layout where the tile's rows are
contiguous ([1024][tile] blocks), i.e. identical instruction stream but
DRAM-friendly addresses."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from fftw3_b200 import binding as B  # noqa: E402

lib = B.load()
n = 1024
total = n ** 3
x = torch.zeros(total, 2, dtype=torch.float64, device="cuda")


def run(name, dims, how, variant):
    os.environ["FFTW3_B200_FORCE_VARIANT"] = str(variant)
    p = lib.plan_guru_dft("d", dims, how, x.data_ptr(), x.data_ptr(), -1, B.FFTW_ESTIMATE)
    assert p, name
    lib.lib.fftw_b200_set_async(1)
    for _ in range(3):
        lib.execute("d", p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lib.execute("d", p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("%-52s variant %2d  %.3f ms  %.0f GB/s   %s" % (name, variant, ms, 32.0 * total / ms / 1e6,
                                                      " ".join(lib.sprint_plan("d", p).split())[-70:]), flush=True)
    lib.lib.fftw_b200_set_async(0)
    lib.destroy_plan("d", p)


for tile, variant in ((4, 14), (8, 15)):
    # (a) real dim-1 pass: stride n, pencils adjacent, planes as outer batch
    run("dim1 real layout (stride 16 KiB)", [(n, n, n)], [(n, 1, 1), (n, n * n, n * n)], variant)
    # (b) dim-0 pass: stride n*n
    run("dim0 real layout (stride 16 MiB)", [(n, n * n, n * n)], [(n * n, 1, 1)], variant)
    # (c) contiguous tiles: [1024][tile] blocks back to back
    run("contiguous tiles [1024][%d]" % tile, [(n, tile, tile)], [(tile, 1, 1), (total // (n * tile), n * tile, n * tile)],
        variant)
    # (d) medium stride: [1024][64] blocks (pencil stride 1 KiB)
    run("blocks [1024][64] (stride 1 KiB)", [(n, 64, 64)], [(64, 1, 1), (total // (n * 64), n * 64, n * 64)], variant)
run("ROW reference", [(n, 1, 1)], [(n * n, n, n)], 13)
# L2 prefetch-size flavours of the tile-4 / tile-8 COL kernels (variants 12 + 6*flavor + log2(tile))
for flavor in (4, 5, 6):
    for tile, lg in ((4, 2), (8, 3)):
        v = 12 + 6 * flavor + lg
        run("dim1 flavor %d tile %d" % (flavor, tile), [(n, n, n)], [(n, 1, 1), (n, n * n, n * n)], v)
        run("dim0 flavor %d tile %d" % (flavor, tile), [(n, n * n, n * n)], [(n * n, 1, 1)], v)
