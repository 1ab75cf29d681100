"""float kernels: the CPU oracle against the REFERENCE's own compute shaders, compiled as C++ (oracle/glsl/comp2cpp.py rewrites the
interface declarations of <module>/<kernel>.comp where it lies under /root/reference, oracle/glsl/glsl_shim.h is the language
runtime; `make -C oracle ref` -> oracle/_ref/libshaderref.so).  the arithmetic, its order and its constants are the shader's;
every kernel here has to agree BIT FOR BIT with the oracle's restatement on seeded inputs (f16 valued, like the edges of the
graph), image borders, clipped and negative values included.  tests/golden/shader_ref.npz holds the shaders' outputs for the
same inputs (written by tests/golden/make_golden.py), so the pin also holds where the reference is absent."""
import ctypes as C
import os
import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "shader_ref.npz")
BAYER, XTRANS = 0x5d5d5d5d, 9


def f16(a):
    return np.ascontiguousarray(a.astype(np.float16).astype(np.float32))


def mosaic(rng, w, h, hot=True):
    m = rng.uniform(0.0, 1.1, (h, w))
    m[h // 3:h // 2, w // 4:w // 2] = rng.uniform(0.95, 1.3, (h // 2 - h // 3, w // 2 - w // 4))   # a clipped region
    if hot:
        m[5, 7] = 4.0
    return f16(m)


def rgba(rng, w, h, lo=-0.05, hi=1.5):
    a = rng.uniform(lo, hi, (h, w, 4))
    a[..., 3] = 1.0
    a[0, 0, :3] = [np.nan, 0.5, 0.2]
    a[0, 1, :3] = 0.0
    a[0, 2, :3] = [-1.0, -2.0, 3.0]
    return f16(a)


def push_wb(wb, filters, extra=()):
    return np.array(list(wb), np.float32).tobytes() + np.array([filters] + list(extra), np.uint32).tobytes()


def cases(O):
    """name -> function returning (outputs of the oracle, outputs of the shader), each a list of float32 arrays"""
    L = O.lib()
    out = {}

    def img_out(h, w, c):
        return O.new_img(h, w, c)

    def add(name):
        def deco(fn):
            out[name] = fn
            return fn
        return deco

    for mode in (0, 1):
        @add("grade.main mode %d" % mode)
        def _(mode=mode):
            rng = np.random.default_rng(10 + mode)
            a = rgba(rng, 96, 64)
            gp = O.GradeParams((C.c_float * 4)(0.01, 0.0, 0.02, 0.0), (C.c_float * 4)(1.1, 1.0, 0.9, 0.05), (C.c_float * 4)(1.0, 1.1, 0.95, 0.0), (C.c_float * 4)(0.0, 0.01, 0.0, 0.0), mode, 0.3, 0.4)
            want, wi = img_out(64, 96, 4)
            L.o_grade_main(C.byref(O.img(a)), C.byref(wi), C.byref(gp), 0)
            got = np.zeros((64, 96, 4), np.float32)
            O.ref_shader("grade", "main", bytes(gp), b"", [(a, 0), (got, 0)], 96, 64)
            return [want], [got]

    @add("llap.curve")
    def _():
        rng = np.random.default_rng(20)
        a = rgba(rng, 96, 64)
        lp = O.LlapParams(0.12, 0.8, 1.3, 0.4)
        outs = [np.zeros((64, 96), np.float32) for _ in range(11)]
        O.ref_shader("llap", "curve", bytes(lp), np.array([10], np.uint32).tobytes(), [(a, 0), [(o, 1) for o in outs]], 96, 64)
        want = [img_out(64, 96, 1) for _ in range(11)]
        L.o_llap_curve(C.byref(O.img(a)), (O.OImg * 11)(*[x[1] for x in want]), C.byref(lp))
        return [x[0].reshape(64, 96) for x in want], outs

    for (w, h) in ((96, 64), (37, 23)):
        @add("llap.reduce %dx%d" % (w, h))
        def _(w=w, h=h):
            rng = np.random.default_rng(21)
            ins = [f16(rng.uniform(0, 1.2, (h, w))) for _ in range(11)]
            cw, ch = (w - 1) // 2 + 1, (h - 1) // 2 + 1
            outs = [np.zeros((ch, cw), np.float32) for _ in range(11)]
            O.ref_shader("llap", "reduce", bytes(O.LlapParams(0.12, 1, 1, 0.2)), b"", [[(i, 0) for i in ins], [(o, 1) for o in outs]], cw, ch, 11)
            want = [img_out(ch, cw, 1) for _ in range(11)]
            for i, x in zip(ins, want):     # the oracle's is one layer per call
                L.o_llap_reduce(C.byref(O.img(i)), C.byref(x[1]))
            return [x[0].reshape(ch, cw) for x in want], outs

        for first in (0, 1):
            @add("llap.assemble %dx%d first %d" % (w, h, first))
            def _(w=w, h=h, first=first):
                rng = np.random.default_rng(22 + first)
                cw, ch = (w - 1) // 2 + 1, (h - 1) // 2 + 1
                l0 = [f16(rng.uniform(0, 1.2, (h, w))) for _ in range(11)]
                l1 = [f16(rng.uniform(0, 1.2, (ch, cw))) for _ in range(11)]
                coarse = l1[10] if first else f16(rng.uniform(0, 1.2, (ch, cw)))
                got = np.zeros((h, w), np.float32)
                # with `first` the coarse input is a dummy binding (llap/main.c:84-86) and the shader reads l1[num_gamma] instead
                O.ref_shader("llap", "assemble", b"", np.array([10, first], np.uint32).tobytes(), [(coarse, 0), [(i, 0) for i in l0], [(i, 0) for i in l1], (got, 1)], w, h)
                want, wi = img_out(h, w, 1)
                L.o_llap_assemble(C.byref(O.img(coarse)), (O.OImg * 11)(*[O.img(i) for i in l0]), (O.OImg * 11)(*[O.img(i) for i in l1]), C.byref(wi), first)
                return [want.reshape(h, w)], [got]

    @add("llap.colour")
    def _():
        rng = np.random.default_rng(24)
        a, lum = rgba(rng, 96, 64), f16(rng.uniform(0, 1.2, (64, 96)))
        want, wi = img_out(64, 96, 4)
        L.o_llap_colour(C.byref(O.img(lum)), C.byref(O.img(a)), C.byref(wi), 1)
        got = np.zeros((64, 96, 4), np.float32)
        O.ref_shader("llap", "colour", bytes(O.LlapParams(0.12, 1, 1, 0.2)), b"", [(lum, 0), (a, 0), (got, 1)], 96, 64)
        return [want], [got]

    for filters, (w, h) in ((BAYER, (96, 64)), (XTRANS, (96, 66)), (BAYER, (38, 26))):
        blk = 3 if filters == XTRANS else 2
        tag = "%s %dx%d" % ("xtrans" if filters == XTRANS else "bayer", w, h)
        hp = O.HiliteParams(0.9, 0.3, 0.6)
        wb = (2.0, 1.0, 1.5, 1.0)

        @add("hilite.half " + tag)
        def _(filters=filters, w=w, h=h, blk=blk, hp=hp, wb=wb):
            m = mosaic(np.random.default_rng(30), w, h)
            want, wi = img_out(h // blk, w // blk, 4)
            L.o_hilite_half(C.byref(O.img(m)), C.byref(wi), C.byref(hp), C.c_uint32(filters))
            got = np.zeros((h // blk, w // blk, 4), np.float32)
            O.ref_shader("hilite", "half", bytes(hp), push_wb(wb, filters), [(m, 0), (got, 1)], w // blk, h // blk)
            return [want], [got]

        @add("hilite.reduce " + tag)
        def _(w=w // blk, h=h // blk, hp=hp, wb=wb, filters=filters):
            a = rgba(np.random.default_rng(31), w, h, 0.0, 1.2)
            a[..., 3] = f16(np.random.default_rng(32).uniform(0, 1, (h, w)))
            cw, ch = (w - 1) // 2 + 1, (h - 1) // 2 + 1
            want, wi = img_out(ch, cw, 4)
            L.o_hilite_reduce(C.byref(O.img(a)), C.byref(wi), C.byref(hp), (C.c_float * 4)(*wb))
            got = np.zeros((ch, cw, 4), np.float32)
            O.ref_shader("hilite", "reduce", bytes(hp), push_wb(wb, filters), [(a, 0), (got, 1)], cw, ch)
            return [want], [got]

        @add("hilite.assemble " + tag)
        def _(w=w // blk, h=h // blk, hp=hp, wb=wb, filters=filters):
            rng = np.random.default_rng(33)
            cw, ch = (w - 1) // 2 + 1, (h - 1) // 2 + 1
            fine, coarse = rgba(rng, w, h, 0.0, 1.2), rgba(rng, cw, ch, 0.0, 1.2)
            fine[..., 3] = f16(rng.uniform(0, 1, (h, w)))
            coarse[..., 3] = f16(rng.uniform(0, 1, (ch, cw)))
            want, wi = img_out(h, w, 4)
            L.o_hilite_assemble(C.byref(O.img(fine)), C.byref(O.img(coarse)), C.byref(wi), C.byref(hp))
            got = np.zeros((h, w, 4), np.float32)
            O.ref_shader("hilite", "assemble", bytes(hp), push_wb(wb, filters, (0,)), [(fine, 0), (coarse, 0), (got, 1)], w, h)
            return [want], [got]

        @add("hilite.doub " + tag)
        def _(filters=filters, w=w, h=h, blk=blk, hp=hp, wb=wb):
            rng = np.random.default_rng(34)
            m, coarse = mosaic(rng, w, h), rgba(rng, w // blk, h // blk, 0.0, 1.2)
            want, wi = img_out(h, w, 1)
            L.o_hilite_doub(C.byref(O.img(m)), C.byref(O.img(coarse)), C.byref(wi), C.byref(hp), C.c_uint32(filters))
            got = np.zeros((h, w), np.float32)
            O.ref_shader("hilite", "doub", bytes(hp), push_wb(wb, filters), [(m, 0), (coarse, 0), (got, 1)], w // blk, h // blk)
            return [want.reshape(h, w)], [got]

        @add("demosaic.down " + tag)
        def _(filters=filters, w=w, h=h, blk=blk):
            m = mosaic(np.random.default_rng(40), w, h)
            want, wi = img_out(h // blk, w // blk, 1)
            L.o_demosaic_down(C.byref(O.img(m)), C.byref(wi), C.c_uint32(filters))
            got = np.zeros((h // blk, w // blk), np.float32)
            O.ref_shader("demosaic", "down", b"", push_wb((1, 1, 1, 1), filters), [(m, 0), (got, 1)], w // blk, h // blk)
            return [want.reshape(h // blk, w // blk)], [got]

        @add("demosaic.halfsize " + tag)
        def _(filters=filters, w=w, h=h, blk=blk):
            m = mosaic(np.random.default_rng(41), w, h)
            want, wi = img_out(h // blk, w // blk, 4)
            L.o_demosaic_halfsize(C.byref(O.img(m)), C.byref(wi), C.c_uint32(filters))
            got = np.zeros((h // blk, w // blk, 4), np.float32)
            O.ref_shader("demosaic", "halfsize", b"", push_wb((1, 1, 1, 1), filters), [(m, 0), (got, 1)], w // blk, h // blk)
            return [want], [got]

        @add("demosaic.splat+fix " + tag)
        def _(filters=filters, w=w, h=h, blk=blk):
            rng = np.random.default_rng(42)
            m = mosaic(rng, w, h, hot=False)
            cov, ci = img_out(h // blk, w // blk, 4)
            L.o_demosaic_gauss(C.byref(O.img(m)), C.byref(ci), C.c_uint32(filters))       # the oracle's covariance image feeds both
            cov = np.ascontiguousarray(cov.reshape(h // blk, w // blk, 4))
            wg, wgi = img_out(h, w, 1)
            L.o_demosaic_splat(C.byref(O.img(m)), C.byref(O.img(cov)), C.byref(wgi), C.c_uint32(filters))
            gg = np.zeros((h, w), np.float32)
            O.ref_shader("demosaic", "splat", b"", push_wb((1, 1, 1, 1), filters), [(m, 0), (cov, 0), (gg, 1)], w, h)
            res_o, res_s = [wg.reshape(h, w)], [gg]
            green = np.ascontiguousarray(wg.reshape(h, w))
            for fixup in (0, 1):
                wo, woi = img_out(h, w, 4)
                L.o_demosaic_fix(C.byref(O.img(m)), C.byref(O.img(green)), C.byref(O.img(cov)), C.byref(woi), C.c_uint32(filters), fixup)
                go = np.zeros((h, w, 4), np.float32)
                O.ref_shader("demosaic", "fix", np.array([fixup], np.int32).tobytes(), push_wb((1, 1, 1, 1), filters), [(m, 0), (green, 0), (cov, 0), (go, 1)], w, h)
                res_o.append(wo)
                res_s.append(go)
            return res_o, res_s
    # ---- pointwise chain: crop, colour, filmcurv ------------------------------------------------------------------------
    for mode in (0, 1, 2, 3, 4, 5):
        @add("filmcurv.main mode %d" % mode)
        def _(mode=mode):
            a = rgba(np.random.default_rng(50 + mode), 96, 64, -0.02, 2.0)
            fp = O.FilmcurvParams(2.5, 1.3, 0.01, mode, 1.1, 0.2, 0.1, -0.1, 0.05, 0.1)
            res_o, res_s = [], []
            for out16 in (1, 0):
                want, wi = img_out(64, 96, 4)
                L.o_filmcurv_main(C.byref(O.img(a)), C.byref(wi), C.byref(fp), out16)
                got = np.zeros((64, 96, 4), np.float32)
                O.ref_shader("filmcurv", "main", bytes(fp), b"", [(a, 0), (got, out16)], 96, 64)
                res_o.append(want)
                res_s.append(got)
            return res_o, res_s

    for k, (matrix, gamut, sat, clip, cnt, mode) in enumerate(((1, 0, 1.0, 0, 4, 0), (2, 0, 1.3, 1, 4, 0), (1, 0, 0.7, 0, 6, 1), (0, 0, 1.0, 0, 4, 0))):
        @add("colour.main set %d" % k)
        def _(k=k, matrix=matrix, gamut=gamut, sat=sat, clip=clip, cnt=cnt, mode=mode):
            rng = np.random.default_rng(60 + k)
            a = rgba(rng, 96, 64, -0.02, 2.0)
            d = O.darkroom_defaults(64, 64)
            p = d.colour
            p.exposure, p.sat, p.matrix, p.gamut, p.clip, p.clipmax, p.cnt, p.mode = 0.4, sat, matrix, gamut, clip, 1.2, cnt, mode
            for i, v in enumerate((1.1, -0.05, -0.05, 0.02, 0.9, 0.08, 0.0, -0.1, 1.1)):
                p.mat[i] = v
            if mode == 1:
                for i in range(6 * cnt):
                    p.rbmap[i] = float(rng.uniform(0.1, 0.8))
            f, _wb = O.colour_commit_oracle(bytes(p), [2.1, 1.0, 1.6, 1.0], [0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8], 0, 0)
            f = np.ascontiguousarray(f, np.float32)
            res_o, res_s = [], []
            for out16 in (1,):
                want, wi = img_out(64, 96, 4)
                L.o_colour_main(C.byref(O.img(a)), C.byref(wi), O.fptr(f), out16)
                got = np.zeros((64, 96, 4), np.float32)
                # clut / pick / abney / spectra / auto_temp are unconnected: the reference binds the input there as a dummy
                O.ref_shader("colour", "main", f.tobytes(), np.zeros(3, np.int32).tobytes(), [(a, 0), (got, out16), (a, 0), (a, 0), (a, 0), (a, 0), (a, 0)], 96, 64)
                res_o.append(want)
                res_s.append(got)
            return res_o, res_s

    # input transfer curves (709 .. gamma, then the camera log curves of shared/oetf.glsl) and input gamuts (camera wide gamuts
    # go through xyz): main-impl.glsl:104-198
    for trc, prim, clip in ((1, 1, 0), (2, 3, 1), (3, 4, 0), (4, 5, 0), (5, 6, 0), (6, 7, 1), (7, 16, 0), (8, 14, 1), (9, 7, 0), (10, 8, 0), (11, 9, 1),
                            (12, 10, 0), (13, 13, 0), (14, 11, 1), (14, 12, 0), (15, 15, 0), (8, 2, 0)):
        @add("colour.main trc %d prim %d" % (trc, prim))
        def _(trc=trc, prim=prim, clip=clip):
            rng = np.random.default_rng(600 + 20 * trc + prim)
            a = rgba(rng, 96, 64, -0.05, 1.0)
            a[1, :, :3] = f16(np.linspace(-0.2, 1.2, 96))[:, None]     # a ramp across every curve's knee
            d = O.darkroom_defaults(64, 64)
            p = d.colour
            p.exposure, p.sat, p.matrix, p.clip, p.clipmax = 0.0, 1.0, 1, clip, 0.9
            f, _wb = O.colour_commit_oracle(bytes(p), [1.0, 1.0, 1.0, 1.0], [1, 0, 0, 0, 1, 0, 0, 0, 1], prim, trc)
            f = np.ascontiguousarray(f, np.float32)
            want, wi = img_out(64, 96, 4)
            L.o_colour_main(C.byref(O.img(a)), C.byref(wi), O.fptr(f), 1)
            got = np.zeros((64, 96, 4), np.float32)
            O.ref_shader("colour", "main", f.tobytes(), np.zeros(3, np.int32).tobytes(), [(a, 0), (got, 1), (a, 0), (a, 0), (a, 0), (a, 0), (a, 0)], 96, 64)
            return [want], [got]

    # lut inputs (main-impl.glsl:76-102, clut.glsl; :287-335): synthetic luts of the shapes the reference's tools write
    def luts(rng, nbands):
        ch = 32
        clut = f16(rng.uniform(0.05, 0.6, (ch, nbands * ch, 2)))
        spectra = np.zeros((48, 48, 4), np.float32)
        sx = rng.uniform(0.5, 2.0, (48, 48)) * np.where(rng.uniform(0, 1, (48, 48)) < 0.5, -1.0, 1.0)
        lam = rng.uniform(380.0, 720.0, (48, 48))
        spectra[..., 0], spectra[..., 1], spectra[..., 2], spectra[..., 3] = sx, -2.0 * sx * lam, rng.uniform(0, 1, (48, 48)), rng.uniform(0.0, 0.9, (48, 48))
        abney = f16(rng.uniform(0.1, 0.6, (40, 64, 2)))
        abney[:, -2:, 0] = f16(rng.uniform(0.5, 0.9, (40, 2)))      # the last two columns: gamut bounds (rec709, then spectral locus / rec2020)
        abney[:, -2:, 1] = f16(rng.uniform(0.6, 1.0, (40, 2)))
        return clut, np.ascontiguousarray(spectra), abney

    for k, (nbands, temp, use_clut, use_abney, sat, gamut, clip) in enumerate(((3, 0.3, 1, 0, 1.0, 0, 0), (6, 0.62, 1, 0, 1.2, 0, 0), (3, 1.0, 1, 1, 1.3, 0, 0),
                                                                             (3, 0.0, 0, 1, 1.0, 1, 0), (3, 0.0, 0, 1, 1.4, 2, 0), (3, 0.0, 0, 1, 0.7, 3, 0),
                                                                             (6, 0.2, 1, 1, 1.5, 3, 1), (9, 0.85, 1, 0, 1.0, 0, 0))):
        @add("colour.main lut set %d" % k)
        def _(k=k, nbands=nbands, temp=temp, use_clut=use_clut, use_abney=use_abney, sat=sat, gamut=gamut, clip=clip):
            rng = np.random.default_rng(640 + k)
            a = rgba(rng, 96, 64, 0.01, 1.2)
            a[0, 0, :3] = [0.3, 0.5, 0.2]                            # the nan of rgba() would be one here and there alike; keep the case finite
            clut, spectra, abney = luts(rng, nbands)
            d = O.darkroom_defaults(64, 64)
            p = d.colour
            p.exposure, p.sat, p.matrix, p.gamut, p.clip, p.clipmax = 0.2, sat, 4 if use_clut else 1, gamut, clip, 0.9
            f, _wb = O.colour_commit_oracle(bytes(p), [2.1, 1.0, 1.6, 1.0], [0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8], 0, 0)
            f = np.ascontiguousarray(f, np.float32)
            f[224] = temp                                           # the anchor blend of colour/main.c:272-292, set directly
            want, wi = img_out(64, 96, 4)
            L.o_colour_main_lut(C.byref(O.img(a)), C.byref(wi), O.fptr(f), 1, C.byref(O.img(clut)) if use_clut else None,
                                C.byref(O.img(abney)) if use_abney else None, C.byref(O.img(spectra)) if use_abney else None, C.c_float(0.0))
            got = np.zeros((64, 96, 4), np.float32)
            push = np.array([use_clut, 0, use_abney], np.int32).tobytes()
            O.ref_shader("colour", "main", f.tobytes(), push, [(a, 0), (got, 1), (clut if use_clut else a, 0), (a, 0),
                                                               (abney if use_abney else a, 0), (spectra if use_abney else a, 0), (a, 0)], 96, 64)
            return [want], [got]

    for nbands in (3, 6, 9):
        @add("colour.autotemp %d bands" % nbands)
        def _(nbands=nbands):
            rng = np.random.default_rng(680 + nbands)
            clut, _s, _a = luts(rng, nbands)
            # chroma anchors that walk across the neutral point, so that the best segment is an inner one
            for b in range(nbands if nbands > 3 else 3):
                clut[:, b * 32:(b + 1) * 32, 0] += f16(np.float32(0.05 * b - 0.1))
            clut = f16(clut)
            res_o, res_s = [], []
            for temp, wb in ((-1.0, (2.1, 1.0, 1.6)), (-1.0, (1.3, 1.0, 2.4)), (5000.0, (2.1, 1.0, 1.6))):
                d = O.darkroom_defaults(64, 64)
                d.colour.temp, d.colour.matrix = temp, 4
                f, _wb = O.colour_commit_oracle(bytes(d.colour), list(wb) + [1.0], [0.8, 0.15, 0.05, 0.1, 0.85, 0.05, 0.02, 0.18, 0.8], 0, 0)
                f = np.ascontiguousarray(f, np.float32)
                L.o_colour_autotemp.restype = C.c_float
                want = np.full((1, 1), L.o_colour_autotemp(C.byref(O.img(clut)), O.fptr(f)), np.float32)
                got = np.zeros((1, 1), np.float32)
                O.ref_shader("colour", "autotemp", f.tobytes(), np.zeros(1, np.int32).tobytes(), [(clut, 0), (got, 0), (clut, 0)], 1, 1)
                res_o.append(want)
                res_s.append(got)
            return res_o, res_s

    # the export's colour encoding (colenc/main.comp): output primaries and transfer curves
    for prim, trc in ((1, 1), (1, 2), (3, 6), (4, 3), (5, 0), (6, 4), (7, 5), (10, 2), (2, 1), (0, 0)):
        @add("colenc.main prim %d trc %d" % (prim, trc))
        def _(prim=prim, trc=trc):
            a = rgba(np.random.default_rng(700 + 10 * prim + trc), 96, 64, -0.05, 1.5)
            L.o_colenc_main.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
            res_o, res_s = [], []
            for out16 in (1, 0):
                want, wi = img_out(64, 96, 4)
                L.o_colenc_main(C.byref(O.img(a)), C.byref(wi), prim, trc, out16)
                got = np.zeros((64, 96, 4), np.float32)
                O.ref_shader("colenc", "main", np.array([prim, trc], np.int32).tobytes(), b"", [(a, 0), (got, out16)], 96, 64)
                res_o.append(want)
                res_s.append(got)
            return res_o, res_s

    for k, (rot, crop, ori) in enumerate(((1337.0, (1.0, 3.0, 3.0, 7.0), 0), (90.0, (0.1, 0.9, 0.2, 0.8), 0), (7.5, (0.1, 0.9, 0.2, 0.8), 0), (1337.0, (1.0, 3.0, 3.0, 7.0), 6))):
        @add("crop.main set %d" % k)
        def _(k=k, rot=rot, crop=crop, ori=ori):
            w, h = 500, 420                                        # > 400: the automatic 3 px micro crop is active
            a = rgba(np.random.default_rng(70 + k), w, h, 0.0, 1.5)
            persp = [0.25, 0.25, 0.75, 0.25, 0.75, 0.75, 0.25, 0.75]
            ow, oh, f = O.crop_oracle(ori, w, h, persp, crop, rot)
            want, wi = img_out(oh, ow, 4)
            L.o_crop_main(C.byref(O.img(a)), C.byref(wi), O.fptr(f))
            got = np.zeros((oh, ow, 4), np.float32)
            blob = np.zeros(28, np.float32)                        # std140: mat3 H = three vec4 columns, then r0..r3, crop window
            for c in range(3):
                blob[4 * c:4 * c + 3] = f[4 * c:4 * c + 3]
            blob[12:20] = f[12:20]
            O.ref_shader("crop", "main", f.tobytes(), b"", [(a, 0), (got, 1)], ow, oh)
            return [want], [got]

    # ---- denoise -----------------------------------------------------------------------------------------------------------
    for filters, (w, h), crop in ((BAYER, (96, 64), (0, 0, 96, 64)), (XTRANS, (96, 66), (0, 0, 96, 66)), (BAYER, (100, 68), (4, 2, 96, 62))):
        blk = 3 if filters == XTRANS else 2
        tag = "%s %dx%d" % ("xtrans" if filters == XTRANS else "bayer", w, h)
        black, white = [2048.0 / 65535.0] * 4, [15000.0 / 65535.0] * 4
        wb, na, nb = (2.0, 1.0, 1.5, 1.0), 100.0, 2.0
        dp = O.DenoiseParams(0.4, 0.6, 1.0, 0.0, (C.c_float * 4)(0, 0, 0, 0), 1)
        cw, ch = crop[2] - crop[0], crop[3] - crop[1]

        def raw_unorm(seed, w=w, h=h):
            rng = np.random.default_rng(seed)
            v = rng.integers(1500, 16383, (h, w)).astype(np.float32)
            v[h // 3:h // 2, w // 4:w // 2] = 16383
            return np.ascontiguousarray((v / np.float32(65535.0)).astype(np.float32))

        def f4(v):
            return np.array(v, np.float32).tobytes()

        def i4(v):
            return np.array(v, np.int32).tobytes()

        @add("denoise.noop " + tag)
        def _(filters=filters, crop=crop, cw=cw, ch=ch, black=black, white=white, dp=dp, raw_unorm=raw_unorm, f4=f4, i4=i4):
            m = raw_unorm(80)
            want, wi = img_out(ch, cw, 1)
            L.o_denoise_noop(C.byref(O.img(m)), C.byref(wi), (C.c_int * 4)(*crop), (C.c_float * 4)(*black), (C.c_float * 4)(*white))
            got = np.zeros((ch, cw, 4), np.float32)
            push = i4(crop) + f4(black) + f4(white) + f4([0, 0, 0, 0]) + np.array([filters], np.uint32).tobytes() + i4([0])
            O.ref_shader("denoise", "noop", bytes(dp) + b"\0" * 12, push, [(m, 0), (got, 1), (m, 0)], cw, ch)
            return [want.reshape(ch, cw)], [got[..., 0]]            # the shader stores (v, 0, 0, 1); consumers read .r

        if filters == BAYER:   # the DNG GainMap branch (noop.comp:48-57): a 9 x 7 rgba f32 gain texture over the uncropped image
            @add("denoise.noop gainmap " + tag)
            def _(filters=filters, crop=crop, cw=cw, ch=ch, black=black, white=white, dp=dp, raw_unorm=raw_unorm, f4=f4, i4=i4):
                m = raw_unorm(82)
                gm = np.ascontiguousarray(np.random.default_rng(83).uniform(0.8, 1.6, (7, 9, 4)).astype(np.float32))
                mos = [0.02, 0.01, 1.05, 0.98]
                want, wi = img_out(ch, cw, 1)
                L.o_denoise_noop_gm(C.byref(O.img(m)), C.byref(wi), (C.c_int * 4)(*crop), (C.c_float * 4)(*black), (C.c_float * 4)(*white),
                                    C.byref(O.img(gm)), (C.c_float * 4)(*mos))
                got = np.zeros((ch, cw, 4), np.float32)
                push = i4(crop) + f4(black) + f4(white) + f4(mos) + np.array([filters], np.uint32).tobytes() + i4([1])
                O.ref_shader("denoise", "noop", bytes(dp) + b"\0" * 12, push, [(m, 0), (got, 1), (gm, 0)], cw, ch)
                return [want.reshape(ch, cw)], [got[..., 0]]

        @add("denoise.half..doub " + tag)
        def _(filters=filters, blk=blk, crop=crop, cw=cw, ch=ch, black=black, white=white, wb=wb, na=na, nb=nb, dp=dp, raw_unorm=raw_unorm, f4=f4, i4=i4):
            m = raw_unorm(81)
            hw, hh = cw // blk, ch // blk
            fbits = np.array([filters], np.uint32).tobytes()
            params = bytes(dp) + b"\0" * 12
            head = f4(wb) + f4(black) + f4(white) + i4(crop)
            cI, bI, wI, wbI = (C.c_int * 4)(*crop), (C.c_float * 4)(*black), (C.c_float * 4)(*white), (C.c_float * 4)(*wb)
            res_o, res_s = [], []
            # each stage: the oracle's output of the previous stage feeds both sides
            half, hi = img_out(hh, hw, 4)
            L.o_denoise_half(C.byref(O.img(m)), C.byref(hi), cI, wI, C.c_uint32(filters))
            g = np.zeros((hh, hw, 4), np.float32)
            O.ref_shader("denoise", "half", params, head + fbits, [(m, 0), (g, 1)], hw, hh)
            res_o.append(half); res_s.append(g)
            dn = [img_out(hh, hw, 4) for _ in range(4)]
            cov, ci = img_out(hh, hw, 4)
            L.o_denoise_downcov(C.byref(hi), C.byref(dn[0][1]), C.byref(ci))
            g0, gc = np.zeros((hh, hw, 4), np.float32), np.zeros((hh, hw, 4), np.float32)
            # the crop window travels in the push constants of half and doub only: for a mosaic the reference hands down / downcov / assemble zeros (denoise/main.c:230-262)
            O.ref_shader("denoise", "downcov", params, head[:48] + i4([0, 0, 0, 0]) + f4([na, nb]) + i4([0, blk]), [(half, 0), (g0, 1), (gc, 1)], hw, hh)
            res_o += [dn[0][0], cov]; res_s += [g0, gc]
            for i in range(1, 4):
                L.o_denoise_down(C.byref(dn[i - 1][1]), C.byref(dn[i][1]), C.byref(dp), bI, wI, C.c_float(na), C.c_float(nb), i, C.c_uint32(blk))
                gi = np.zeros((hh, hw, 4), np.float32)
                O.ref_shader("denoise", "down", params, head[:48] + i4([0, 0, 0, 0]) + f4([na, nb]) + i4([i, blk]), [(dn[i - 1][0], 0), (gi, 1)], hw, hh)
                res_o.append(dn[i][0]); res_s.append(gi)
            asm, ai = img_out(hh, hw, 4)
            L.o_denoise_assemble(C.byref(hi), C.byref(dn[0][1]), C.byref(dn[1][1]), C.byref(dn[2][1]), C.byref(dn[3][1]), C.byref(ai), C.byref(dp), wbI, bI, wI,
                                 C.c_float(na), C.c_float(nb), C.c_uint32(filters))
            ga = np.zeros((hh, hw, 4), np.float32)
            O.ref_shader("denoise", "assemble", params, head[:48] + i4([0, 0, 0, 0]) + f4([na, nb]) + fbits,
                         [(half, 0), (dn[0][0], 0), (dn[1][0], 0), (dn[2][0], 0), (dn[3][0], 0), (ga, 1)], hw, hh)
            res_o.append(asm); res_s.append(ga)
            out, oi = img_out(ch, cw, 1)
            L.o_denoise_doub(C.byref(O.img(m)), C.byref(ai), C.byref(hi), C.byref(oi), C.byref(dp), cI, bI, wI, C.c_float(na), C.c_float(nb), C.c_uint32(filters))
            gd = np.zeros((ch, cw), np.float32)
            O.ref_shader("denoise", "doub", params, head + fbits + f4([na, nb]) + i4([0]) + f4([0, 0, 0, 0]), [(m, 0), (asm, 0), (half, 0), (gd, 1), (half, 0)], cw, ch)
            res_o.append(out.reshape(ch, cw)); res_s.append(gd)
            if filters == BAYER:   # doub.comp:106-114: the same stage with a gain map
                gm = np.ascontiguousarray(np.random.default_rng(84).uniform(0.8, 1.6, (7, 9, 4)).astype(np.float32))
                mos = [0.02, 0.01, 1.05, 0.98]
                out2, oi2 = img_out(ch, cw, 1)
                L.o_denoise_doub_gm(C.byref(O.img(m)), C.byref(ai), C.byref(hi), C.byref(oi2), C.byref(dp), cI, bI, wI, C.c_float(na), C.c_float(nb), C.c_uint32(filters),
                                    C.byref(O.img(gm)), (C.c_float * 4)(*mos))
                gd2 = np.zeros((ch, cw), np.float32)
                O.ref_shader("denoise", "doub", params, head + fbits + f4([na, nb]) + i4([1]) + f4(mos), [(m, 0), (asm, 0), (half, 0), (gd2, 1), (gm, 0)], cw, ch)
                res_o.append(out2.reshape(ch, cw)); res_s.append(gd2)
            return res_o, res_s

    for filters, (w, h) in ((BAYER, (96, 64)), (XTRANS, (96, 66))):
        @add("demosaic.gauss %s" % ("xtrans" if filters == XTRANS else "bayer"))
        def _(filters=filters, w=w, h=h):
            blk = 3 if filters == XTRANS else 2
            m = mosaic(np.random.default_rng(43), w, h, hot=False)
            want, wi = img_out(h // blk, w // blk, 4)
            L.o_demosaic_gauss(C.byref(O.img(m)), C.byref(wi), C.c_uint32(filters))
            got = np.zeros((h // blk, w // blk, 4), np.float32)
            dummy = np.zeros((h // blk, w // blk), np.float32)
            O.ref_shader("demosaic", "gauss", b"", push_wb((1, 1, 1, 1), filters), [(dummy, 0), (m, 0), (got, 1)], w // blk, h // blk)
            return [want], [got]
    @add("demosaic.rcd_conv")
    def _():
        w, h = 96, 64
        m = mosaic(np.random.default_rng(44), w, h, hot=False)
        wo = [img_out(h, w, 1), img_out(h, w // 2, 1), img_out(h, w // 2, 1)]
        L.o_rcd_conv(C.byref(O.img(m)), C.byref(wo[0][1]), C.byref(wo[1][1]), C.byref(wo[2][1]))
        go = [np.zeros((h, w), np.float32), np.zeros((h, w // 2), np.float32), np.zeros((h, w // 2), np.float32)]
        O.ref_shader("demosaic", "rcd_conv", b"", b"", [(m, 0), (go[0], 1), (go[1], 1), (go[2], 1)], w, h)
        return [x[0].reshape(g.shape) for x, g in zip(wo, go)], go

    @add("demosaic.rcd_fill")
    def _():
        w, h = 96, 64
        m = mosaic(np.random.default_rng(46), w, h, hot=False)
        planes = [img_out(h, w, 1), img_out(h, w // 2, 1), img_out(h, w // 2, 1)]
        L.o_rcd_conv(C.byref(O.img(m)), C.byref(planes[0][1]), C.byref(planes[1][1]), C.byref(planes[2][1]))
        wb = (2.0, 1.0, 1.5, 1.0)
        want, wi = img_out(h, w, 4)
        L.o_rcd_fill(C.byref(O.img(m)), C.byref(planes[0][1]), C.byref(planes[1][1]), C.byref(planes[2][1]), C.byref(wi), (C.c_float * 4)(*wb))
        got = np.zeros((h, w, 4), np.float32)
        # 58 x 26 px tiles of 32 x 32 invocations, dispatched as tiles * 8 (demosaic/main.c:125-131); runs one thread per invocation
        O.ref_shader("demosaic", "rcd_fill", b"", push_wb(wb, BAYER), [(m, 0)] + [(np.ascontiguousarray(x[0].reshape(x[0].shape[0], -1)), 0) for x in planes] + [(got, 1)],
                     (w + 57) // 58 * 8, (h + 25) // 26 * 8)
        # the reference keeps 58 x 26 px of every 64 x 32 tile (DT_RCD_BORDER = 3) while the second and third step reach 4 px through
        # the shared memory tile (rcd_fill.comp:96-99, 114-135): within 3 px of a tile seam, and within 6 px of the image border, its
        # output depends on the tiling (reads beyond the tile wrap into neighbouring rows of the shared arrays).  the oracle and the
        # product are tiling independent (deviation 5); everywhere else the two have to agree bit for bit
        yy, xx = np.mgrid[0:h, 0:w]
        keep = ((xx % 58 >= 3) & (xx % 58 < 55) & (yy % 26 >= 3) & (yy % 26 < 23) & (xx >= 6) & (xx < w - 6) & (yy >= 6) & (yy < h - 6))
        assert keep.sum() > 3000
        return [np.asarray(want).reshape(h, w, 4)[keep]], [got[keep]]

    for (ow, oh) in ((96, 64), (90, 60)):             # 1:1 (what vkdt-cli inserts behind demosaic) and a mild downscale
        @add("shared.resample to %dx%d" % (ow, oh))
        def _(ow=ow, oh=oh):
            a = rgba(np.random.default_rng(45), 96, 64, 0.0, 1.5)
            want, wi = img_out(oh, ow, 4)
            L.o_resample(C.byref(O.img(a)), C.byref(wi))
            got = np.zeros((oh, ow, 4), np.float32)
            O.ref_shader("shared", "resample", b"", b"", [[(a, 0)], [(got, 1)]], ow, oh)
            return [want], [got]
    # the export side's resize module: slice / flower (magnify) / catmull-rom (minify), f16 and f32 outputs, and the separable blur
    for mode, (ow, oh), f16o in ((1, (30, 20), 0), (2, (60, 40), 1), (2, (41, 29), 1), (0, (150, 100), 1)):
        @add("resize.main mode %d to %dx%d" % (mode, ow, oh))
        def _(mode=mode, ow=ow, oh=oh, f16o=f16o):
            a = rgba(np.random.default_rng(50 + mode), 96, 64, 0.0, 1.5)
            want, wi = img_out(oh, ow, 4)
            L.o_resize_main(C.byref(O.img(a)), C.byref(wi), mode, f16o)
            got = np.zeros((oh, ow, 4), np.float32)
            O.ref_shader("resize", "main", b"", np.array([mode], np.int32).tobytes(), [(a, 0), (got, f16o)], ow, oh)
            return [want], [got]
    for radius in (3.67, 4.0, 1.5, 0.7):
        for vert in (0, 1):
            @add("shared.blur%s radius %g" % ("v" if vert else "h", radius))
            def _(radius=radius, vert=vert):
                a = rgba(np.random.default_rng(55), 96, 64, 0.0, 1.5)
                want, wi = img_out(64, 96, 4)
                L.o_blur_sep.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int]
                L.o_blur_sep(C.byref(O.img(a)), C.byref(wi), radius, vert, 1)
                got = np.zeros((64, 96, 4), np.float32)
                O.ref_shader("shared", "blurv" if vert else "blurh", b"", np.array([radius], np.float32).tobytes(), [[(a, 0)], [(got, 1)]], 96, 64)
                return [want], [got]
    return out


def _same(a, b):
    a, b = np.asarray(a, np.float32).ravel(), np.asarray(b, np.float32).ravel()
    return a.shape == b.shape and np.array_equal(np.where(np.isnan(a), np.float32(-7), a).view(np.uint32), np.where(np.isnan(b), np.float32(-7), b).view(np.uint32))


def _f16_ulps(a, b):
    """distance in representable f16 values"""
    def key(x):
        u = x.astype(np.float16).view(np.uint16).astype(np.int32)
        return np.where(u & 0x8000, 0x8000 - u, u)
    return np.abs(key(a) - key(b))


# kernels that FILTER (texture() at fractional coordinates): the shader computes its texture coordinates in fp32, the oracle is an
# ideal sampler that carries them in double (oracle/o_common.h:122-143), so a weight can differ in its last bits and an
# f16 store can then round the other way.  everything else is bit exact.
SAMPLED = ("llap.reduce", "llap.assemble", "denoise.half..doub", "denoise.noop gainmap", "shared.resample", "resize.main mode 0", "resize.main mode 2", "shared.blur")


def _report(name, want, got):
    bad = []
    for k, (a, b) in enumerate(zip(want, got)):
        a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
        if name.startswith("demosaic.halfsize xtrans"):
            a, b = a[..., :3], b[..., :3]       # the shader leaves alpha unwritten for x-trans (demosaic/halfsize.comp:20-33)
        if _same(a, b):
            continue
        d = np.abs(a.astype(np.float64) - b.astype(np.float64)).ravel()
        n = int((d > 0).sum() + (np.isnan(a) != np.isnan(b)).sum())
        if name.startswith(SAMPLED) and np.array_equal(np.isnan(a), np.isnan(b)) and _f16_ulps(a, b).max() <= 1 and n <= 0.02 * d.size:
            continue
        bad.append("%s[%d]: %d of %d values differ, max abs %.3g" % (name, k, n, d.size, np.nanmax(d)))
    return bad


def test_oracle_matches_reference_shaders_live(oracle):
    if oracle.ref_shader_lib() is None:
        pytest.skip("oracle/_ref/libshaderref.so not built (needs /root/reference: make -C oracle ref)")
    bad, n = [], 0
    for name, fn in cases(oracle).items():
        want, got = fn()
        bad += _report(name, want, got)
        n += len(want)
    assert not bad, "\n".join(bad)
    assert n >= 110


def digest(name, a):
    """sha256 over the float32 bytes of what _report compares bit for bit (NaNs canonical, x-trans halfsize without its unwritten alpha)"""
    import hashlib
    a = np.asarray(a, np.float32)
    if name.startswith("demosaic.halfsize xtrans"):
        a = a[..., :3]
    a = np.where(np.isnan(a), np.float32(-7), a).astype(np.float32)
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() + " %s" % (a.shape,)


def test_oracle_matches_reference_shader_goldens(oracle):
    """the same against what the shaders gave when the reference was present (the inputs are seeded): images for the filtering
    kernels, sha256 digests for the bit exact ones (tests/golden/shader_ref.npz, shader_ref.json)"""
    import json
    G = np.load(GOLDEN)
    D = json.load(open(GOLDEN.replace(".npz", ".json")))
    bad, n = [], 0
    real = oracle.ref_shader
    try:
        oracle.ref_shader = lambda *a, **k: None          # the shader side of each case stays zero: only the oracle side is used
        for name, fn in cases(oracle).items():
            want, _ = fn()
            for k, w in enumerate(want):
                key = "%s/%d" % (name, k)
                if name.startswith(SAMPLED):
                    bad += _report(key, [w], [G[key].astype(np.float32).reshape(np.asarray(w).shape)])
                elif digest(name, w) != D[key]:
                    bad.append("%s: digest differs" % key)
                n += 1
    finally:
        oracle.ref_shader = real
    assert not bad, "\n".join(bad)
    assert n >= 110 and n == len(G.files) + len(D)
