"""regenerates the committed golden vectors.  run in the build container (needs /root/reference for the MLV part):

    python tests/golden/make_golden.py

 - mlv_unpack_{10,12,14}.npz: seeded pixels packed into an MLV file and decoded by the REFERENCE's own
   video_mlv.c (oracle/_ref/libmlvref.so).  pins the oracle's o_mlv_unpack and the CUDA unpack kernel bit exactly.
 - darkroom_*.npz: outputs of the CPU oracle for the default darkroom graph on a small synthetic frame.  these pin
   the oracle against accidental edits (the float path has no reference-made vectors: parity unpinned).
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from vkdt_b200 import synth          # noqa: E402
from oracle import oracle_py as O    # noqa: E402


def mlv_goldens():
    O.build()
    for bpp in (10, 12, 14):
        rng = np.random.default_rng(0xC0FFEE + bpp)
        w, h = 136, 50
        pix = rng.integers(0, 1 << bpp, (h, w), dtype=np.uint16)
        pix[0, :4] = [0, (1 << bpp) - 1, 1, (1 << bpp) - 2]
        with tempfile.TemporaryDirectory() as td:
            fn = os.path.join(td, "g.mlv")
            synth.write_mlv(fn, [pix], bpp=bpp, black=64, white=(1 << bpp) - 1)
            out, info = O.ref_mlv_decode(fn, 0)
            payload = open(fn, "rb").read()
        assert info["bpp"] == bpp and (out == pix).all(), "reference decoder disagrees with the packer"
        words = synth.pack_bits(pix, bpp)
        np.savez_compressed(os.path.join(HERE, "mlv_unpack_%d.npz" % bpp), words=words, expected=out, width=w, height=h, bpp=bpp)
        print("mlv golden", bpp, out.shape, len(payload))


def darkroom_goldens():
    w, h = 168, 126
    raw = synth.mosaic(w, h, seed=77)
    for name, strength in (("default", 0.0), ("denoise", 0.4)):
        d = O.darkroom_defaults(w, h)
        for k, v in enumerate((2.0, 1.0, 1.5)):
            d.whitebalance[k] = v
        d.denoise.strength = strength
        d.noise_a, d.noise_b = 100.0, 2.0
        out = O.darkroom_run(d, raw)
        np.savez_compressed(os.path.join(HERE, "darkroom_%s.npz" % name), raw=raw, out=out.astype(np.float32), strength=strength)
        print("darkroom golden", name, out.shape, float(out[..., :3].mean()))


if __name__ == "__main__":
    mlv_goldens()
    darkroom_goldens()
