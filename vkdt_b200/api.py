"""ctypes bindings of libvkdt_b200.so (include/vkdt_b200.h).  No fallback: a missing library is an ImportError,
a missing GPU makes every compute call raise VkbError(VKB_ERR_NO_DEVICE)."""
import ctypes as C
import os
import struct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvkdt_b200.so")
if not os.path.exists(LIB_PATH):
    raise ImportError("vkdt_b200: %s is missing - build it with `make -C vkdt_b200` (no CPU fallback exists)" % LIB_PATH)
lib = C.CDLL(LIB_PATH)


class VkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vkdt_b200 error %d: %s" % (code, msg))
        self.code = code


class Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("wd", C.c_uint32), ("ht", C.c_uint32), ("chan", C.c_uint32),
                ("layers", C.c_uint32), ("format", C.c_uint64)]


class RawParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("filters", C.c_uint32), ("crop_aabb", C.c_uint32 * 4),
                ("black", C.c_float * 4), ("white", C.c_float * 4), ("whitebalance", C.c_float * 4),
                ("cam_to_rec2020", C.c_float * 9), ("noise_a", C.c_float), ("noise_b", C.c_float),
                ("orientation", C.c_uint32), ("packed_bpp", C.c_uint32)]


lib.vkb_token.restype = C.c_uint64
lib.vkb_token.argtypes = [C.c_char_p]
lib.vkb_last_error.restype = C.c_char_p
lib.vkb_version.restype = C.c_char_p
lib.vkb_launch_count.restype = C.c_uint64
lib.vkb_dispatch.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32,
                             C.c_void_p, C.c_uint32, C.POINTER(Image), C.c_uint32, C.c_void_p]

# every symbol include/vkdt_b200.h declares (checked by tests/test_cabi.py)
DECLARED = """vkb_token vkb_init vkb_cleanup vkb_device_count vkb_last_error vkb_version vkb_malloc vkb_free
vkb_memcpy_h2d vkb_memcpy_d2h vkb_stream_sync vkb_host_alloc vkb_host_free vkb_event_create vkb_event_record
vkb_event_sync vkb_event_elapsed_ms vkb_event_destroy vkb_dispatch vkb_kernel_count
vkb_kernel_name vkb_launch_count vkb_launch_count_reset vkb_graph_new vkb_graph_free vkb_graph_read_config_ascii
vkb_graph_read_config_line vkb_graph_replace_display vkb_graph_set_source vkb_graph_set_sink_buffer
vkb_graph_sink_size vkb_graph_set_frame vkb_graph_run vkb_graph_plan vkb_graph_perf vkb_graph_dump_nodes
vkb_graph_set_source_device vkb_graph_sink_device vkb_graph_pool_bytes vkb_graph_stream vkb_graph_set_device vkb_dng_info vkb_graph_set_sink_layout vkb_lj92_decode vkb_graph_committed_params vkb_graph_describe vkb_graph_state vkb_graph_set_perf vkb_set_mode vkb_get_mode vkb_graph_set_mode vkb_graph_set_bands vkb_graph_band_plan vkb_graph_band_stats vkb_graph_band_mark vkb_graph_band_elapsed_ms vkb_graph_replace_display_ex vkb_graph_replace_display_sized vkb_graph_set_dng_opcodes vkb_dng_opcodes_describe vkb_jpeg_write vkb_graph_frame_count vkb_set_basedir vkb_module_describe vkb_graph_apply_keyframes vkb_graph_has_feedback vkb_register_module vkb_register_kernel""".split()


def token(s):
    return lib.vkb_token(s.encode())


def check(r):
    if r != 0:
        raise VkbError(r, lib.vkb_last_error().decode(errors="replace"))
    return r


def init(device=0):
    return check(lib.vkb_init(device))


def image(t, wd, ht, chan, fmt, layers=1):
    """describe a torch tensor (or raw device pointer) as a connector image."""
    ptr = t if isinstance(t, int) else (t.data_ptr() if t is not None else 0)
    im = Image(ptr, wd, ht, chan, layers, token(fmt))
    im._keep = t  # keep the tensor alive as long as the descriptor (a temporary would be recycled by torch's allocator)
    return im


def dispatch(name, kernel, conns, push=b"", params=b"", wd=0, ht=0, dp=1, stream=0):
    arr = (Image * len(conns))(*conns)
    pb = C.create_string_buffer(push, len(push)) if push else None
    qb = C.create_string_buffer(params, len(params)) if params else None
    return check(lib.vkb_dispatch(token(name), token(kernel), wd, ht, dp, pb, len(push), qb, len(params), arr, len(conns), stream))


def kernels():
    out = []
    n, k = C.c_uint64(), C.c_uint64()
    for i in range(lib.vkb_kernel_count()):
        lib.vkb_kernel_name(i, C.byref(n), C.byref(k))
        out.append((struct.pack("<Q", n.value).rstrip(b"\0").decode(), struct.pack("<Q", k.value).rstrip(b"\0").decode()))
    return out


def launch_count():
    return int(lib.vkb_launch_count())


MODE_STRICT, MODE_FAST = 0, 1


def set_mode(mode):
    """process default arithmetic mode of the kernels (MODE_STRICT: bit compatible with the CPU restatement; MODE_FAST: SFU)."""
    return check(lib.vkb_set_mode(mode))


def set_basedir(path):
    """<path>/modules/<name>/{connectors,params} of a vkdt installation or checkout become the module tables."""
    lib.vkb_set_basedir.argtypes = [C.c_char_p]
    return check(lib.vkb_set_basedir(path.encode() if path else None))


def module_describe(name):
    lib.vkb_module_describe.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    b = C.create_string_buffer(1 << 16)
    check(lib.vkb_module_describe(name.encode(), b, len(b)))
    return b.value.decode()


def get_mode():
    return int(lib.vkb_get_mode())


# ---- graph layer -------------------------------------------------------------------------------------------
RUN_ALL = -1
SINK_RGBA_F32, SINK_RGB_F32, SINK_RGBA_UI8, SINK_RGB_UI8 = 0, 1, 2, 3
PRIM = {"custom": 0, "srgb": 1, "bt2020": 2, "adobergb": 3, "p3": 4, "xyz": 5}      # cli/main.c:14-23
TRC = {"linear": 0, "709": 1, "srgb": 2, "pq": 3, "dci": 4, "hlg": 5, "gamma2.2": 6}  # cli/main.c:25-35
RUN_ROI, RUN_CREATE_NODES, RUN_ALLOC, RUN_RECORD, RUN_UPLOAD, RUN_DOWNLOAD, RUN_WAIT, RUN_PERF = 1, 2, 4, 8, 16, 32, 64, 1 << 16

lib.vkb_graph_new.restype = C.c_void_p
lib.vkb_graph_free.argtypes = [C.c_void_p]
lib.vkb_graph_read_config_ascii.argtypes = [C.c_void_p, C.c_char_p]
lib.vkb_graph_read_config_line.argtypes = [C.c_void_p, C.c_char_p]
lib.vkb_graph_replace_display.argtypes = [C.c_void_p, C.c_char_p]
lib.vkb_graph_replace_display_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
lib.vkb_jpeg_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_float]
lib.vkb_graph_set_source.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(RawParams)]
lib.vkb_graph_set_source_device.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(RawParams)]
lib.vkb_graph_set_sink_buffer.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
lib.vkb_graph_sink_size.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
lib.vkb_graph_sink_device.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
lib.vkb_graph_set_frame.argtypes = [C.c_void_p, C.c_uint32]
lib.vkb_graph_frame_count.argtypes = [C.c_void_p]
lib.vkb_graph_run.argtypes = [C.c_void_p, C.c_int]
lib.vkb_graph_plan.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.vkb_graph_perf.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.vkb_graph_set_perf.argtypes = [C.c_void_p, C.c_int]
lib.vkb_graph_set_mode.argtypes = [C.c_void_p, C.c_int]
lib.vkb_set_mode.argtypes = [C.c_int]
lib.vkb_graph_set_bands.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
lib.vkb_graph_band_plan.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.vkb_graph_band_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.vkb_graph_band_mark.argtypes = [C.c_void_p, C.c_int]
lib.vkb_graph_band_elapsed_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
lib.vkb_graph_dump_nodes.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.vkb_graph_pool_bytes.argtypes = [C.c_void_p]
lib.vkb_graph_pool_bytes.restype = C.c_uint64
lib.vkb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
lib.vkb_host_free.argtypes = [C.c_void_p]
lib.vkb_graph_stream.argtypes = [C.c_void_p]
lib.vkb_graph_stream.restype = C.c_void_p
lib.vkb_graph_set_device.argtypes = [C.c_void_p, C.c_int]
lib.vkb_graph_set_sink_layout.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
lib.vkb_graph_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.vkb_graph_state.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
lib.vkb_graph_committed_params.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p]
lib.vkb_lj92_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.vkb_dng_info.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.vkb_event_create.argtypes = [C.POINTER(C.c_void_p)]
lib.vkb_event_record.argtypes = [C.c_void_p, C.c_void_p]
lib.vkb_event_sync.argtypes = [C.c_void_p]
lib.vkb_event_elapsed_ms.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
lib.vkb_event_destroy.argtypes = [C.c_void_p]
lib.vkb_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
lib.vkb_free.argtypes = [C.c_void_p]
lib.vkb_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.vkb_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
lib.vkb_stream_sync.argtypes = [C.c_void_p]


class Event:
    def __init__(self):
        self.e = C.c_void_p()
        check(lib.vkb_event_create(C.byref(self.e)))

    def record(self, stream):
        check(lib.vkb_event_record(self.e, C.c_void_p(stream)))

    def sync(self):
        check(lib.vkb_event_sync(self.e))

    def elapsed_ms(self, later):
        ms = C.c_float()
        check(lib.vkb_event_elapsed_ms(self.e, later.e, C.byref(ms)))
        return ms.value


def host_alloc(nbytes):
    """pinned host memory as a ctypes pointer value (free with host_free)."""
    p = C.c_void_p()
    check(lib.vkb_host_alloc(C.byref(p), nbytes))
    return p.value


def host_free(ptr):
    check(lib.vkb_host_free(C.c_void_p(ptr)))


def dev_alloc(nbytes):
    p = C.c_void_p()
    check(lib.vkb_malloc(C.byref(p), nbytes))
    return p.value


def dev_free(ptr):
    check(lib.vkb_free(C.c_void_p(ptr)))

# the reference's bin/default-darkroom.i-raw / .i-mlv module and connection lines (gui coordinates dropped)
DARKROOM_CFG = """module:{src}:main
module:denoise:01
module:hilite:01
module:demosaic:01
module:colour:01
module:filmcurv:01
module:llap:01
module:grade:01
module:hist:01
module:zones:01
module:crop:01
module:lens:01
module:pick:01
module:display:hist
module:display:main
connect:{src}:main:output:denoise:01:input
connect:denoise:01:output:hilite:01:input
connect:hilite:01:output:demosaic:01:input
connect:demosaic:01:output:crop:01:input
connect:crop:01:output:colour:01:input
connect:colour:01:output:filmcurv:01:input
connect:filmcurv:01:output:llap:01:input
connect:llap:01:output:grade:01:input
connect:grade:01:output:display:main:input
connect:grade:01:output:hist:01:input
connect:hist:01:output:display:hist:input
param:colour:01:exposure:0
param:llap:01:sigma:0.12
param:llap:01:shadows:1
param:llap:01:hilights:1
param:llap:01:clarity:0.2
"""


class Graph:
    """dt_graph_t behind the C-ABI: read cfg lines, feed a source from memory, run, fetch the sink."""

    def __init__(self, cfg_text=None, cfg_file=None, sink="o-pfm", prim=None, trc=None, max_width=0, max_height=0):
        self.h = C.c_void_p(lib.vkb_graph_new())
        self._keep = []
        if cfg_file:
            check(lib.vkb_graph_read_config_ascii(self.h, cfg_file.encode()))
        if cfg_text:
            for line in cfg_text.splitlines():
                r = lib.vkb_graph_read_config_line(self.h, line.encode())
                if r < 0:
                    raise VkbError(r, "config line failed: " + line)
        if sink and (max_width or max_height):   # cli --width / --height: a resize module in front of the sink
            lib.vkb_graph_replace_display_sized.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
            check(lib.vkb_graph_replace_display_sized(self.h, b"main", sink.encode(), 2 if prim is None else prim, 0 if trc is None else trc,
                                                      int(max_width), int(max_height)))
        elif sink and (prim is not None or trc is not None):
            check(lib.vkb_graph_replace_display_ex(self.h, b"main", sink.encode(), 2 if prim is None else prim, 0 if trc is None else trc))
        elif sink:
            check(lib.vkb_graph_replace_display(self.h, sink.encode()))

    def line(self, text):
        return lib.vkb_graph_read_config_line(self.h, text.encode())

    def set_source(self, data_ptr, params, inst="main", device=False):
        fn = lib.vkb_graph_set_source_device if device else lib.vkb_graph_set_source
        check(fn(self.h, inst.encode(), C.c_void_p(data_ptr), C.byref(params)))

    def set_dng_opcodes(self, blob, ox=0, oy=0, inst="main"):
        """OpcodeList2 of an in-memory i-raw source (bytes as in the dng tag) + the cfa offset of the window"""
        lib.vkb_graph_set_dng_opcodes.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        check(lib.vkb_graph_set_dng_opcodes(self.h, inst.encode(), bytes(blob), len(blob), ox, oy))

    def set_sink_buffer(self, ptr, nbytes, inst="main"):
        check(lib.vkb_graph_set_sink_buffer(self.h, inst.encode(), C.c_void_p(ptr) if ptr else None, nbytes))

    def set_sink_layout(self, layout, inst="main"):
        """SINK_RGBA_F32 (16 B/px, the reference's mapped sink buffer) or SINK_RGB_F32 (12 B/px, the PFM payload)."""
        check(lib.vkb_graph_set_sink_layout(self.h, inst.encode(), layout))

    def committed_params(self, module, inst="01"):
        """float32 view of the uniform block commit_params() of that module produces for the current graph (host only)."""
        import numpy as np
        buf = np.zeros(4096, dtype=np.uint8)
        n = C.c_size_t(buf.nbytes)
        check(lib.vkb_graph_committed_params(self.h, module.encode(), inst.encode(), buf.ctypes.data_as(C.c_void_p), C.byref(n)))
        return buf[:n.value].view(np.float32).copy()

    def set_frame(self, frame):
        check(lib.vkb_graph_set_frame(self.h, int(frame)))

    def apply_keyframes(self):
        lib.vkb_graph_apply_keyframes.argtypes = [C.c_void_p]
        check(lib.vkb_graph_apply_keyframes(self.h))

    def has_feedback(self):
        lib.vkb_graph_has_feedback.argtypes = [C.c_void_p]
        return bool(lib.vkb_graph_has_feedback(self.h))

    def run(self, flags=RUN_ALL):
        check(lib.vkb_graph_run(self.h, flags))

    def sink_size(self, inst="main"):
        w, h = C.c_uint32(), C.c_uint32()
        check(lib.vkb_graph_sink_size(self.h, inst.encode(), C.byref(w), C.byref(h)))
        return w.value, h.value

    def sink_device(self, inst="main"):
        p = C.c_void_p()
        check(lib.vkb_graph_sink_device(self.h, inst.encode(), C.byref(p)))
        return p.value

    def plan(self):
        b = C.create_string_buffer(1 << 18)
        check(lib.vkb_graph_plan(self.h, b, len(b)))
        return b.value.decode()

    def perf(self):
        b = C.create_string_buffer(1 << 16)
        lib.vkb_graph_perf(self.h, b, len(b))
        return b.value.decode()

    def set_mode(self, mode):
        check(lib.vkb_graph_set_mode(self.h, int(mode)))

    def set_bands(self, devices):
        """band split of one frame over these CUDA devices (an ordinal may repeat: bands sharing a GPU)."""
        arr = (C.c_int * len(devices))(*devices)
        check(lib.vkb_graph_set_bands(self.h, len(devices), arr))

    def band_plan(self):
        b = C.create_string_buffer(1 << 22)
        check(lib.vkb_graph_band_plan(self.h, b, len(b)))
        return b.value.decode()

    def band_stats(self):
        t, m, np_, nk = C.c_uint64(), C.c_uint64(), C.c_int(), C.c_int()
        check(lib.vkb_graph_band_stats(self.h, C.byref(t), C.byref(m), C.byref(np_), C.byref(nk)))
        return dict(bytes_total=t.value, bytes_max_device=m.value, pulls=np_.value, launches=nk.value)

    def band_mark(self, which):
        check(lib.vkb_graph_band_mark(self.h, which))

    def band_elapsed_ms(self):
        ms = C.c_float()
        check(lib.vkb_graph_band_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def set_perf(self, on=True):
        check(lib.vkb_graph_set_perf(self.h, int(on)))

    def dump_nodes(self):
        b = C.create_string_buffer(1 << 18)
        lib.vkb_graph_dump_nodes(self.h, b, len(b))
        return b.value.decode()

    def describe(self):
        """module and node layer as text (host only): see vkb_graph_describe."""
        b = C.create_string_buffer(1 << 20)
        check(lib.vkb_graph_describe(self.h, b, len(b)))
        return b.value.decode()

    def state(self):
        """frame count, parameter blocks and connections of all modules as text (host only): see vkb_graph_state."""
        b = C.create_string_buffer(1 << 20)
        check(lib.vkb_graph_state(self.h, b, len(b)))
        return b.value.decode()

    def stream(self):
        return lib.vkb_graph_stream(self.h) or 0

    def set_device(self, dev):
        check(lib.vkb_graph_set_device(self.h, dev))

    def perf_entries(self):
        """[(label, ms, unique bytes in+out)] of the last synchronised run."""
        out = []
        for l in self.perf().splitlines():
            if l.startswith("[perf] total") or not l.startswith("[perf]"):
                continue
            label, rest = l[7:].split(":\t", 1) if ":\t" in l else (l, "")
            f = rest.split("\t")
            if len(f) >= 2:
                out.append((label.strip(), float(f[0].split()[0]), int(f[1].split()[0])))
        return out

    def pool_bytes(self):
        return int(lib.vkb_graph_pool_bytes(self.h))

    def close(self):
        if self.h:
            lib.vkb_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lj92_decode(data):
    """bytes of a lossless jpeg stream -> (height, width * components) uint16 array, bit depth.  Host only."""
    import numpy as np
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    w, h, bits, comps = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib.vkb_lj92_decode(buf, len(data), None, 0, C.byref(w), C.byref(h), C.byref(bits), C.byref(comps)))
    out = np.zeros((h.value, w.value * comps.value), dtype=np.uint16)
    check(lib.vkb_lj92_decode(buf, len(data), out.ctypes.data_as(C.c_void_p), out.size, C.byref(w), C.byref(h), C.byref(bits), C.byref(comps)))
    return out, bits.value


def dng_info(filename):
    """(RawParams, cfa_off_x, cfa_off_y) of an uncompressed cfa dng, as i-raw resolves it.  Host only."""
    p = RawParams()
    ox, oy = C.c_uint32(0), C.c_uint32(0)
    check(lib.vkb_dng_info(filename.encode(), C.byref(p), C.byref(ox), C.byref(oy)))
    return p, ox.value, oy.value


def raw_params(width, height, black=2048.0, white=15000.0, wb=(1.0, 1.0, 1.0), cam_to_rec2020=None, filters=0x5d5d5d5d,
               noise_a=1.0, noise_b=1.0, packed_bpp=0, crop_aabb=None):
    p = RawParams()
    p.width, p.height, p.filters = width, height, filters
    ca = crop_aabb or (0, 0, width, height)
    for k in range(4):
        p.crop_aabb[k] = ca[k]
        p.black[k] = black
        p.white[k] = white
        p.whitebalance[k] = wb[k] if k < 3 else 1.0
    m = cam_to_rec2020 or (1, 0, 0, 0, 1, 0, 0, 0, 1)
    for k in range(9):
        p.cam_to_rec2020[k] = m[k]
    p.noise_a, p.noise_b, p.orientation, p.packed_bpp = noise_a, noise_b, 0, packed_bpp
    return p
