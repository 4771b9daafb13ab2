"""python -m sister_b200.cli <input folder/> <dmax> [output folder/] -- the reference's sample CLI (cpp/src/compute_disp.cpp)
for PNG folders, on top of cv2: reads <folder>center.png ... bottom.png with cv2.imread (compute_disp.cpp:19-23, the
folder string is concatenated as is), runs compute_disparities on the GPU and writes the three CV_16UC1 maps as 16-bit
PNG plus the pictures the reference shows in windows (min-max normalise, MAGMA, 0.1/0.9 blend; compute_disp.cpp:38-56).
tools/compute_disp.cpp is the same tool in C++ for PPM/PGM folders."""
from __future__ import annotations

import sys


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) not in (2, 3):
        print("expected <input folder> <dmax> [output folder]", file=sys.stderr)
        return 1
    import cv2

    from . import SisterMultiviewDisparities

    folder, disp_count = argv[0], int(argv[1])
    out = argv[2] if len(argv) == 3 else folder
    views = []
    for name in ("center", "right", "top", "left", "bottom"):
        im = cv2.imread(folder + name + ".png")
        if im is None:
            print(f"cannot read {folder}{name}.png", file=sys.stderr)
            return 1
        views.append(im)
    mv, hz, vt = SisterMultiviewDisparities(*views).compute_disparities(disp_count)
    colored = {}
    for name, m in (("disp_multiview", mv), ("disp_horizontal", hz), ("disp_vertical", vt)):
        cv2.imwrite(out + name + ".png", m)
        g = cv2.normalize(m, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1)
        colored[name] = cv2.applyColorMap(g, cv2.COLORMAP_MAGMA)
        cv2.imwrite(out + name + "_magma.png", colored[name])
    cv2.imwrite(out + "blended.png", cv2.addWeighted(views[0], 0.1, colored["disp_multiview"], 0.9, 0.0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
