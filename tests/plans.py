"""module-level launch sequences expressed through the C-ABI dispatch (the same sequences the C++ executor
records); used by the per-module parity tests."""
from helpers import dev_f16, dev_f32, fbits, ubits, ibits, levels


def hilite(api, d_in, w, h, params, wb=(1, 1, 1, 1), filters=0x5d5d5d5d):
    """d_in: f16 mosaic (h,w) on device. params = (white, desat, soft). returns f16 mosaic tensor."""
    blk = 3 if filters == 9 else 2
    push = fbits(*wb) + ubits(filters)
    par = fbits(*params)
    I = api.image
    lv = levels(w // blk, h // blk, 15)
    pyr = [dev_f16(lh, lw, 4) for (lw, lh) in lv]
    api.dispatch("hilite", "half", [I(d_in, w, h, 1, "f16"), I(pyr[0], lv[0][0], lv[0][1], 4, "f16")], push, par)
    for l in range(1, len(lv)):
        api.dispatch("hilite", "reduce", [I(pyr[l - 1], *lv[l - 1], 4, "f16"), I(pyr[l], *lv[l], 4, "f16")], push, par)
    coarse = pyr[-1]
    keep = []
    for l in range(len(lv) - 1, 0, -1):
        out = dev_f16(lv[l - 1][1], lv[l - 1][0], 4)
        api.dispatch("hilite", "assemble", [I(pyr[l - 1], *lv[l - 1], 4, "f16"), I(coarse, *lv[l], 4, "f16"),
                                             I(out, *lv[l - 1], 4, "f16")], push, par)
        keep.append(coarse)
        coarse = out
    d_out = dev_f16(h, w)
    api.dispatch("hilite", "doub", [I(d_in, w, h, 1, "f16"), I(coarse, *lv[0], 4, "f16"), I(d_out, w, h, 1, "f16")], push, par)
    return d_out


def demosaic(api, d_in, w, h, filters=0x5d5d5d5d, fixup=0):
    blk = 3 if filters == 9 else 2
    push = fbits(1, 1, 1, 1) + ubits(filters)
    I = api.image
    cov = dev_f16(h // blk, w // blk, 4)
    green = dev_f16(h, w)
    out = dev_f16(h, w, 4)
    api.dispatch("demosaic", "gauss", [I(None, 0, 0, 1, "f16"), I(d_in, w, h, 1, "f16"), I(cov, w // blk, h // blk, 4, "f16")], push)
    api.dispatch("demosaic", "splat", [I(d_in, w, h, 1, "f16"), I(cov, w // blk, h // blk, 4, "f16"), I(green, w, h, 1, "f16")], push)
    api.dispatch("demosaic", "fix", [I(d_in, w, h, 1, "f16"), I(green, w, h, 1, "f16"), I(cov, w // blk, h // blk, 4, "f16"),
                                     I(out, w, h, 4, "f16")], push, ibits(fixup, 0))
    return out, cov, green


def llap(api, d_in, w, h, params, grade=None, out_f32=False, final_kernel="llapfin"):
    """d_in: rgba f16 (h,w,4). params = (sigma, shadows, hilights, clarity); grade: 19-value tuple packed bytes or None."""
    I = api.image
    par = fbits(*params)
    lv = levels(w, h, 12)
    nl = len(lv)
    stacks = [None] + [dev_f16(lh, lw, 1, layers=11) for (lw, lh) in lv[1:]]
    api.dispatch("b200", "llapr0", [I(d_in, w, h, 4, "f16"), I(stacks[1], *lv[1], 1, "f16", layers=11)], b"", par)
    for l in range(2, nl):
        api.dispatch("llap", "reduce", [I(stacks[l - 1], *lv[l - 1], 1, "f16", layers=11), I(stacks[l], *lv[l], 1, "f16", layers=11)])
    coarse = None
    for l in range(nl - 1, 1, -1):  # assemble node l writes level l-1
        out = dev_f16(lv[l - 1][1], lv[l - 1][0])
        first = 1 if l == nl - 1 else 0
        api.dispatch("llap", "assemble", [I(coarse, *lv[l], 1, "f16"), I(stacks[l - 1], *lv[l - 1], 1, "f16", layers=11),
                                          I(stacks[l], *lv[l], 1, "f16", layers=11), I(out, *lv[l - 1], 1, "f16")], ubits(10, first))
        coarse = out
    out = dev_f32(h, w) if out_f32 else dev_f16(h, w, 4)
    first = 1 if nl == 2 else 0
    api.dispatch("b200", final_kernel, [I(d_in, w, h, 4, "f16"), I(coarse, *lv[1], 1, "f16"), I(stacks[1], *lv[1], 1, "f16", layers=11),
                                     I(out, w, h, 4, "f32" if out_f32 else "f16")], ubits(first, 1 if grade else 0),
                 par + (grade or b""))
    return out
