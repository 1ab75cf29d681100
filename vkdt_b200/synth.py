"""Synthetic raw inputs of the shapes BASELINE.json names and a minimal MLV writer
(container layout per SURVEY.md Appendix E: MLVI + RAWI + N x VIDF, little endian, packed(1)).

There is no network and no camera file in the image, so every benchmark/test input is generated here:
a smooth scene (low-frequency sinusoids) + sharp rectangles + two discs driven 1.5x over the white level
(so that hilite has work to do), sampled through the CFA, scaled to 14 bit with black 2048 / white 15000,
plus Poisson-Gaussian noise sigma^2 = a + b (x - black).
"""
import struct
import numpy as np

BLACK = 2048
WHITE = 15000

# canonical 6x6 x-trans layout after alignment (0 r, 1 g, 2 b), rows top to bottom
XTRANS = np.array([[1, 0, 1, 1, 2, 1],
                   [2, 1, 2, 0, 1, 0],
                   [1, 0, 1, 1, 2, 1],
                   [1, 2, 1, 1, 0, 1],
                   [0, 1, 0, 2, 1, 2],
                   [1, 2, 1, 1, 0, 1]], dtype=np.int32)


def cfa_pattern(width, height, xtrans=False):
    """(H,W) int32 map of colour index 0 r / 1 g / 2 b; bayer is RGGB starting at (0,0)."""
    if xtrans:
        return np.tile(XTRANS, ((height + 5) // 6, (width + 5) // 6))[:height, :width]
    return np.tile(np.array([[0, 1], [1, 2]], dtype=np.int32), ((height + 1) // 2, (width + 1) // 2))[:height, :width]


def scene_rgb(width, height, seed=0x5EED0000):
    """linear camera rgb in [0, ~1.5], float32 (H,W,3); deterministic in (width, height, seed)."""
    rng = np.random.default_rng(seed & 0xFFFFFFFF)
    y = np.linspace(0.0, 1.0, height, dtype=np.float32)[:, None]
    x = np.linspace(0.0, 1.0, width, dtype=np.float32)[None, :]
    img = np.zeros((height, width, 3), dtype=np.float32)
    base = np.full((height, width), 0.28, dtype=np.float32)
    for _ in range(6):
        fx, fy = rng.uniform(0.5, 6.0, 2)
        ph = rng.uniform(0, 2 * np.pi)
        amp = rng.uniform(0.02, 0.08)
        base += np.float32(amp) * np.sin(np.float32(2 * np.pi) * (np.float32(fx) * x + np.float32(fy) * y) + np.float32(ph))
    tint = np.array([0.9, 1.0, 0.8], dtype=np.float32)
    img[:] = base[:, :, None] * tint
    for k in range(3):  # sharp edged rectangles
        x0, y0 = rng.uniform(0.05, 0.6, 2)
        w, h = rng.uniform(0.1, 0.3, 2)
        col = rng.uniform(0.05, 0.9, 3).astype(np.float32)
        xa, xb = int(x0 * width), int(min(1.0, x0 + w) * width)
        ya, yb = int(y0 * height), int(min(1.0, y0 + h) * height)
        img[ya:yb, xa:xb] = col
    for k in range(2):  # clipped discs
        cx, cy = rng.uniform(0.2, 0.8, 2)
        r = rng.uniform(0.03, 0.08)
        asp = width / float(height)
        d2 = ((x - np.float32(cx)) * np.float32(asp)) ** 2 + (y - np.float32(cy)) ** 2
        m = d2 < np.float32(r * r)
        img[m] = np.array([1.5, 1.45, 1.2], dtype=np.float32)
    return np.maximum(img, 0.0)


def mosaic(width, height, seed=0x5EED0000, xtrans=False, noise_a=100.0, noise_b=2.0, wb=(2.0, 1.0, 1.5)):
    """(H,W) uint16 14-bit mosaic. wb: camera white balance multipliers; the sensor sees scene/wb."""
    rgb = scene_rgb(width, height, seed)
    pat = cfa_pattern(width, height, xtrans)
    val = np.take_along_axis(rgb, pat[:, :, None], axis=2)[:, :, 0]
    val = val / np.asarray(wb, dtype=np.float32)[pat]
    lin = val * np.float32(WHITE - BLACK)
    rng = np.random.default_rng((seed ^ 0xA5A5A5A5) & 0xFFFFFFFF)
    sigma = np.sqrt(np.float32(noise_a) + np.float32(noise_b) * np.maximum(lin, 0.0))
    lin = lin + sigma * rng.standard_normal(lin.shape, dtype=np.float32)
    return np.clip(np.rint(lin + BLACK), 0, 16383).astype(np.uint16)


def pack_bits(pix, bpp=14):
    """pack (N,) uint16 pixels MSB-first into a stream of little-endian 16-bit words (raw.h:30-43,
    video_mlv.c:261-273).  returns uint16 words incl. 2 spare words of padding."""
    pix = np.ascontiguousarray(pix, dtype=np.uint16).ravel()
    n = pix.size
    nbits = n * bpp
    nwords = (nbits + 15) // 16
    shifts = np.arange(bpp - 1, -1, -1, dtype=np.uint16)
    bits = ((pix[:, None] >> shifts) & 1).astype(np.uint8).ravel()
    pad = nwords * 16 - nbits
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)])
    by = np.packbits(bits)  # big-endian bit order within bytes, byte stream = big-endian words
    words = by.reshape(-1, 2)
    words = (words[:, 0].astype(np.uint16) << 8) | words[:, 1].astype(np.uint16)
    return np.concatenate([words, np.zeros(2, dtype=np.uint16)])


def pack_bits_fast14(pix):
    """vectorised 14-bit packer: 8 pixels -> 7 words. pix.size must be a multiple of 8."""
    p = np.ascontiguousarray(pix, dtype=np.uint16).ravel().astype(np.uint32).reshape(-1, 8)
    acc = np.zeros((p.shape[0], 7), dtype=np.uint32)
    # bit position of pixel i in the 112-bit group: [14 i, 14 i + 14)
    for i in range(8):
        start = 14 * i
        wi, sh = divmod(start, 16)
        # pixel occupies bits sh..sh+13 of word wi (MSB first), may spill into word wi+1
        room = 16 - sh
        if room >= 14:
            acc[:, wi] |= p[:, i] << (room - 14)
        else:
            acc[:, wi] |= p[:, i] >> (14 - room)
            acc[:, wi + 1] |= (p[:, i] << (16 - (14 - room))) & 0xFFFF
    words = (acc & 0xFFFF).astype(np.uint16).ravel()
    return np.concatenate([words, np.zeros(2, dtype=np.uint16)])


RAWI_SIZE = 4 + 4 + 8 + 2 + 2 + 160


def write_mlv(filename, frames, bpp=14, black=BLACK, white=WHITE, fps=(24000, 1000), camera_name=None, lossless=False):
    """frames: list of (H,W) uint16 arrays. writes MLVI + RAWI [+ IDNT] + VIDF*N, uncompressed (packed bits) or, with
    lossless=True, one lossless jpeg stream per frame (MLV_VIDEO_CLASS_FLAG_LJ92, two interleaved components like
    Magic Lantern writes them)."""
    h, w = frames[0].shape
    with open(filename, "wb") as f:
        f.write(struct.pack("<4sI8sQHHIHHIIII", b"MLVI", 52, b"v2.0\0\0\0\0", 0x1234, 0, 1, 0, 0x21 if lossless else 1, 0,
                            len(frames), 0, fps[0], fps[1]))
        raw_info = struct.pack("<II3iiiii4i4i2iii18ii", 1, 0, h, w, w * bpp // 8, w * h * bpp // 8, bpp, black, white,
                               0, 0, w, h, 0, 0, h, w, 0, 0, 0x02010100, 21, *([0] * 18), 1100)
        assert len(raw_info) == 160, len(raw_info)
        f.write(struct.pack("<4sIQHH", b"RAWI", RAWI_SIZE, 1, w, h) + raw_info)
        if camera_name:
            name = camera_name.encode()[:31]
            f.write(struct.pack("<4sIQ32sI32s", b"IDNT", 4 + 4 + 8 + 32 + 4 + 32, 2, name, 0x80000285, b""))
        for i, fr in enumerate(frames):
            assert fr.shape == (h, w)
            if lossless:
                payload = lj92_encode(fr, bits=bpp, components=2 if w % 2 == 0 else 1, predictor=1,
                                      lengths=[2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12])
                # the reference reads the stream into a buffer of the UNCOMPRESSED frame size (video_mlv.c:215)
                assert len(payload) <= w * h * bpp // 8, "lossless frame larger than the packed one: use smoother test data"
            else:
                words = pack_bits_fast14(fr) if (bpp == 14 and fr.size % 8 == 0) else pack_bits(fr, bpp)
                payload = words[:(w * h * bpp // 8 + 1) // 2].tobytes()[: w * h * bpp // 8]
            f.write(struct.pack("<4sIQIHHHHI", b"VIDF", 32 + len(payload), 10 + i, i, 0, 0, 0, 0, 0))
            f.write(payload)
    return filename


def write_dng(filename, raw, cfa=((0, 1), (1, 2)), black=2048, white=15000, neutral=(0.5, 1.0, 0.6),
              color_matrix=None, illuminant=21, active_area=None, make="b200", model="synth", iso=100,
              orientation=1, big_endian=False, opcode_list2=None):
    """Write `raw` (h x w uint16) as an uncompressed 16-bit CFA DNG (TIFF/EP + DNG 1.4 tags, one strip, raw data in IFD0).
    cfa: 2x2 or 6x6 nested sequence of 0 r / 1 g / 2 b as stored.  Test input for the i-raw file path."""
    import struct
    raw = np.ascontiguousarray(raw, dtype=np.uint16)
    h, w = raw.shape
    cfa = np.asarray(cfa, dtype=np.uint8)
    dim = cfa.shape[0]
    if color_matrix is None:
        color_matrix = (0.9, -0.3, -0.1, -0.4, 1.2, 0.2, -0.1, 0.2, 0.7)
    e = ">" if big_endian else "<"

    def rat(vals, signed=False):
        out = b""
        for v in vals:
            out += struct.pack(e + ("ii" if signed else "II"), int(round(v * 10000)), 10000)
        return out

    entries = []  # (tag, type, count, payload bytes)

    def add(tag, typ, count, payload):
        entries.append((tag, typ, count, payload))

    def short(v): return struct.pack(e + "H", v)
    def long_(v): return struct.pack(e + "I", v)
    add(254, 4, 1, long_(0))
    add(256, 4, 1, long_(w))
    add(257, 4, 1, long_(h))
    add(258, 3, 1, short(16))
    add(259, 3, 1, short(1))
    add(262, 3, 1, short(32803))
    add(271, 2, len(make) + 1, make.encode() + b"\0")
    add(272, 2, len(model) + 1, model.encode() + b"\0")
    add(273, 4, 1, None)  # strip offset, patched below
    add(274, 3, 1, short(orientation))
    add(277, 3, 1, short(1))
    add(278, 4, 1, long_(h))
    add(279, 4, 1, long_(w * h * 2))
    add(33421, 3, 2, short(dim) + short(dim))
    add(33422, 1, dim * dim, cfa.tobytes())
    add(34855, 3, 1, short(iso))
    add(50706, 1, 4, bytes((1, 4, 0, 0)))
    bl = np.broadcast_to(np.asarray(black, dtype=np.float64).ravel(), (4,)) if np.size(black) in (1, 4) else None
    add(50713, 3, 2, short(2) + short(2))
    add(50714, 5, 4, rat(bl))
    add(50717, 4, 1, long_(int(white)))
    add(50721, 10, 9, rat(color_matrix, True))
    add(50728, 5, 3, rat(neutral))
    add(50778, 3, 1, short(illuminant))
    if active_area is not None:
        add(50829, 4, 4, b"".join(long_(v) for v in active_area))
    if opcode_list2 is not None:      # OpcodeList2, type UNDEFINED: big endian bytes whatever the file's order (dng_opcode_list())
        add(51009, 7, len(opcode_list2), bytes(opcode_list2))
    entries.sort(key=lambda t: t[0])
    n = len(entries)
    ifd_off = 8
    extra_off = ifd_off + 2 + 12 * n + 4
    extra = b""
    body = b""
    strip_patch = None
    for tag, typ, count, payload in entries:
        if payload is None:
            strip_patch = len(body) + 8
            payload = long_(0)
        if len(payload) <= 4:
            val = payload.ljust(4, b"\0")
        else:
            if (extra_off + len(extra)) & 1:
                extra += b"\0"
            val = long_(extra_off + len(extra))
            extra += payload
        body += struct.pack(e + "HHI", tag, typ, count) + val
    if (extra_off + len(extra)) & 1:
        extra += b"\0"
    data_off = extra_off + len(extra)
    body = body[:strip_patch] + long_(data_off) + body[strip_patch + 4:]
    hdr = (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, ifd_off)
    pix = raw.astype(">u2" if big_endian else "<u2").tobytes()
    with open(filename, "wb") as f:
        f.write(hdr + struct.pack(e + "H", n) + body + long_(0) + extra + pix)


def write_pfm(filename, img):
    """(H,W,3|4) or (H,W) float32 -> PFM as o-pfm writes it (o-pfm/main.c:27-40: header padded with '0' so that the data
    starts 16-byte aligned, rows top to bottom, scale -1.0 = little endian)."""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    hdr = ("PF\n%d %d\n-1.0" if img.ndim == 3 else "Pf\n%d %d\n-1.0") % (w, h)
    pad = 0
    while (len(hdr) + 1 + pad) & 0xf:
        pad += 1
    with open(filename, "wb") as f:
        f.write((hdr + "0" * pad + "\n").encode())
        f.write(np.ascontiguousarray(img[..., :3] if img.ndim == 3 else img).astype("<f4").tobytes())


def lj92_encode(img, bits=14, components=1, predictor=1, lengths=None):
    """(H,W) uint16 -> lossless jpeg (T.81 annex H) byte stream.  components 2 packs pairs of columns into interleaved
    components the way Magic Lantern's lossless MLV frames do (frame width W = X * components).  Test helper: plain python,
    use small images.  lengths: code length per difference category 0..16 (default: a fixed 5-bit code for all 17)."""
    import struct
    img = np.asarray(img, dtype=np.int64)
    h, wtot = img.shape
    assert wtot % components == 0
    x = wtot // components
    lengths = list(lengths) if lengths is not None else [5] * 17
    assert len(lengths) == 17
    order = sorted(range(17), key=lambda s: (lengths[s], s))
    counts = [sum(1 for s in range(17) if lengths[s] == l) for l in range(1, 17)]
    codes, code, prev = {}, 0, lengths[order[0]]
    for s in order:            # canonical code assignment, T.81 annex C
        code <<= lengths[s] - prev
        prev = lengths[s]
        codes[s] = (code, lengths[s])
        code += 1
    out = bytearray(b"\xff\xd8")
    out += b"\xff\xc4" + struct.pack(">H", 2 + 17 + 17) + bytes([0]) + bytes(counts) + bytes(order)
    out += b"\xff\xc3" + struct.pack(">HBHHB", 8 + 3 * components, bits, h, x, components)
    for c in range(components):
        out += bytes([c + 1, 0x11, 0])
    out += b"\xff\xda" + struct.pack(">HB", 6 + 2 * components, components)
    for c in range(components):
        out += bytes([c + 1, 0x00])
    out += bytes([predictor, 0, 0])
    acc, nbits = 0, 0
    body = bytearray()

    def put(v, n):
        nonlocal acc, nbits
        acc = (acc << n) | (v & ((1 << n) - 1)); nbits += n
        while nbits >= 8:
            b = (acc >> (nbits - 8)) & 0xff
            body.append(b)
            if b == 0xff:
                body.append(0)
            nbits -= 8
    nc = components
    for y in range(h):
        for i in range(wtot):
            xx = i // nc
            if y == 0:
                pred = (1 << (bits - 1)) if xx == 0 else img[y, i - nc]
            elif xx == 0:
                pred = img[y - 1, i]
            else:
                ra, rb, rc = int(img[y, i - nc]), int(img[y - 1, i]), int(img[y - 1, i - nc])
                pred = [ra, rb, rc, ra + rb - rc, ra + ((rb - rc) >> 1), rb + ((ra - rc) >> 1), (ra + rb) >> 1][predictor - 1]
            d = (int(img[y, i]) - int(pred)) & 0xffff
            if d >= 32768:
                d -= 65536
            if d == -32768:
                d = 32768
            ssss = 16 if d == 32768 else (abs(d)).bit_length()
            put(*codes[ssss])
            if 0 < ssss < 16:
                put(d if d >= 0 else d + (1 << ssss) - 1, ssss)
    if nbits:
        put((1 << (8 - nbits)) - 1, 8 - nbits)      # pad with ones
    return bytes(out + body + b"\xff\xd9")


def dng_gain_map_opcode(gains, top, left, bottom, right, spacing=(None, None), origin=(0.0, 0.0), row_pitch=2, col_pitch=2):
    """one GainMap opcode (DNG 1.4, opcode id 9) for a single plane: `gains` is map_points_v x map_points_h float32"""
    import struct
    g = np.asarray(gains, dtype=np.float32)
    pv, ph = g.shape
    sv = spacing[0] if spacing[0] is not None else 1.0 / (pv - 1)
    sh = spacing[1] if spacing[1] is not None else 1.0 / (ph - 1)
    body = struct.pack(">8I2I4dI", top, left, bottom, right, 0, 1, row_pitch, col_pitch, pv, ph, sv, sh, origin[0], origin[1], 1)
    body += g.astype(">f4").tobytes()
    return struct.pack(">4I", 9, 0x01030000, 0, len(body)) + body


def dng_opcode_list(opcodes):
    import struct
    return struct.pack(">I", len(opcodes)) + b"".join(opcodes)
