"""External pin for the restated CirclePolygon plan (TEST INFRASTRUCTURE; needs /root/reference and PIL; run once).

matplotlib is not installed here, so the static 2D / 3D plans (Env/2D/DMP_Env_2D_static.py:31-52) rest on
oracle/dmp_oracle.py's restatement of ``matplotlib.patches.CirclePolygon.contains_point``.  The reference ships ONE
artefacts that show what real matplotlib produced: docs/2d_example_crop.png and the right panel of
docs/3d_example_crop.png, two renderings of the sparse design (plan_choose=1) on its 20 x 20 grid with different labels in
the way.  This script decodes both figures cell by cell -- ring colour (green / the yellow
highlighted brick) = plan cell, black = empty, anything else (the two text labels, the observation box, arrows, the red
brick) = occluded --, checks that they agree wherever both show a cell, and commits their union as
tests/golden/sparse_ring_from_docs.npz:
    cell  u8 [20][20]  1 = plan cell, 0 = empty (row 0 = top row of the figure)
    known u8 [20][20]  1 where the figure shows the cell unoccluded
tests/test_oracle_golden.py asserts that the oracle's sparse mask equals `cell` wherever `known` is set (the mask is
symmetric under both flips, so the figure's y-axis direction does not matter; the test checks that too).

    python tests/golden/make_polygon_pin.py
"""
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/docs/2d_example_crop.png"
SRC2 = "/root/reference/docs/3d_example_crop.png"      # its right panel shows the same design with other occlusions
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sparse_ring_from_docs.npz")


def decode(path=SRC, left_fraction=0.0):
    im = np.asarray(Image.open(path).convert("RGB")).astype(int)
    r, g, b = im[..., 0], im[..., 1], im[..., 2]
    black = (r < 40) & (g < 40) & (b < 40)
    ring = ((g > 200) & (r < 200) & (b < 120)) | ((r > 220) & (g > 220) & (b < 120))      # green | yellow
    area = black.copy()
    area[:, :int(im.shape[1] * left_fraction)] = False                                  # (3D figure: ignore the left panel)
    cols, rows = area.sum(0), area.sum(1)                                               # the solid black square, not stray text
    xs, ys = np.where(cols > 0.5 * cols.max())[0], np.where(rows > 0.5 * rows.max())[0]
    x0, x1, y0, y1 = xs.min(), xs.max() + 1, ys.min(), ys.max() + 1                     # the 20 x 20 grid area
    cw, ch = (x1 - x0) / 20.0, (y1 - y0) / 20.0
    assert abs(cw - ch) < 0.5, (cw, ch)
    cell = np.zeros((20, 20), np.uint8)
    known = np.zeros((20, 20), np.uint8)
    for i in range(20):
        for j in range(20):
            ya, yb = int(y0 + (i + 0.2) * ch), int(y0 + (i + 0.8) * ch)
            xa, xb = int(x0 + (j + 0.2) * cw), int(x0 + (j + 0.8) * cw)
            fr, fb = ring[ya:yb, xa:xb].mean(), black[ya:yb, xa:xb].mean()
            if fr > 0.97:
                cell[i, j], known[i, j] = 1, 1
            elif fb > 0.97:
                known[i, j] = 1
    return cell, known


if __name__ == "__main__":
    cell, known = decode()
    cell2, known2 = decode(SRC2, left_fraction=0.5)
    both = (known & known2).astype(bool)
    assert np.array_equal(cell[both], cell2[both]), "the two figures disagree"
    print("2D figure: %d cells known; 3D figure: %d; both: %d, identical there" % (known.sum(), known2.sum(), both.sum()))
    cell, known = cell | cell2, known | known2
    np.savez_compressed(OUT, cell=cell, known=known)
    print("plan cells seen: %d, cells known: %d of 400" % (cell.sum(), known.sum()))
    for i in range(20):
        print("".join("#" if c else ("." if k else "?") for c, k in zip(cell[i], known[i])))
