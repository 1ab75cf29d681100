#!/usr/bin/env python3
"""writes vkdt_b200/csrc/kernels/munsell_table.h: the 40 hues x 21 chromas of munsell renotation chromaticities (adapted to D65)
that filmcurv's colour mode 2 looks up, as the reference packs them (two halfs per word: shared/munsell.glsl:4-15).
data only; run in a container that has the reference checkout:  scripts/make_munsell_table.py /root/reference"""
import re
import sys
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
src = open(ref + "/src/pipe/modules/shared/munsell.glsl").read()
body = re.search(r"munsell_xy\[21\*40\]\s*=\s*\{([^}]*)\}", src).group(1)
words = np.array([int(x) for x in body.replace("\n", " ").split(",") if x.strip()], dtype=np.uint32)
assert words.size == 21 * 40
xy = words.view(np.float16).reshape(40, 21, 2).astype(np.float32)
with open("vkdt_b200/csrc/kernels/munsell_table.h", "w") as f:
    f.write("// munsell renotation chromaticities, hue major: word [21 * hue + chroma] = half(x) | half(y) << 16  (40 hues, 21 chroma steps;\n")
    f.write("// D65 adapted).  the data of the reference's shared/munsell.glsl:4-15, written by scripts/make_munsell_table.py.\n")
    f.write("// x ranges %.4f .. %.4f, y %.4f .. %.4f; chroma 0 of every hue is the white point (%.5f, %.5f)\n" % (
        xy[..., 0].min(), xy[..., 0].max(), xy[..., 1].min(), xy[..., 1].max(), xy[0, 0, 0], xy[0, 0, 1]))
    f.write("#pragma once\n#define VKB_MUNSELL_HDIM 40\n#define VKB_MUNSELL_CDIM 21\n#define VKB_MUNSELL_WORDS \\\n")
    for h in range(40):
        f.write("  /* hue %2d */ " % h + " ".join("0x%08xu," % w for w in words[21 * h:21 * h + 21]) + (" \\\n" if h < 39 else "\n"))
print("wrote %d words" % words.size)
