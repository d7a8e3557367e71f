"""Streaming zstd COMPRESSION on the host, through the system's libzstd (ctypes).

Compression is not on the accelerated path: the reference keeps it on the CPU (`zstd::Encoder`, crate zstd -> libzstd,
nafcodec/src/encoder/mod.rs:147-154) and so does this package.  The call pattern is the reference's: `write` =
ZSTD_compressStream, `flush` = ZSTD_flushStream (closes the current block), `finish` = ZSTD_endStream, magicless frames
(`include_magicbytes(false)`).  Decompression never goes through this module: that is what the CUDA kernels are for.
"""
import ctypes as C

_ZSTD_c_compressionLevel = 100
_ZSTD_c_format = 10             # ZSTD_c_experimentalParam2
_ZSTD_f_zstd1_magicless = 1


class _Buf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


class _InBuf(C.Structure):             # ZSTD_inBuffer over a bytes object: no copy, no cast (the per-record calls are the cost)
    _fields_ = [("ptr", C.c_char_p), ("size", C.c_size_t), ("pos", C.c_size_t)]


_lib = None


def _zstd():
    global _lib
    if _lib is None:
        L = C.CDLL("libzstd.so.1")
        L.ZSTD_createCCtx.restype = C.c_void_p
        L.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        L.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ZSTD_CCtx_setParameter.restype = C.c_size_t
        for f in ("ZSTD_compressStream",):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(_Buf), C.c_void_p]        # (input: byref of _Buf or _InBuf)
            getattr(L, f).restype = C.c_size_t
        for f in ("ZSTD_flushStream", "ZSTD_endStream"):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(_Buf)]
            getattr(L, f).restype = C.c_size_t
        L.ZSTD_isError.argtypes = [C.c_size_t]
        L.ZSTD_getErrorName.argtypes = [C.c_size_t]
        L.ZSTD_getErrorName.restype = C.c_char_p
        _lib = L
    return _lib


class StreamEncoder:
    """One `zstd::Encoder` of the reference: level (0 = zstd's default), magicless, no pledged size."""

    def __init__(self, level: int = 0):
        self._z = _zstd()
        self._c = self._z.ZSTD_createCCtx()
        if not self._c:
            raise MemoryError("ZSTD_createCCtx")
        self._z.ZSTD_CCtx_setParameter(self._c, _ZSTD_c_compressionLevel, level)
        self._z.ZSTD_CCtx_setParameter(self._c, _ZSTD_c_format, _ZSTD_f_zstd1_magicless)
        self._out = bytearray()
        self._scratch = (C.c_uint8 * (1 << 17))()
        self._o = _Buf(C.addressof(self._scratch), len(self._scratch), 0)     # reused for every call
        self._o_ref = C.byref(self._o)
        self.written = 0                      # WriteCounter (encoder/counter.rs:25-34): bytes accepted

    def _check(self, r):
        if self._z.ZSTD_isError(r):
            raise OSError("zstd: " + self._z.ZSTD_getErrorName(r).decode())
        return r

    def _drain(self):
        n = self._o.pos
        if n:
            self._out += C.string_at(self._scratch, n)
            self._o.pos = 0

    def write(self, data) -> None:
        n = len(data)
        if n == 0:
            return
        if isinstance(data, C.Array):
            i = _Buf(C.addressof(data), n, 0)
        else:
            i = _InBuf(data if isinstance(data, bytes) else bytes(data), n, 0)
        i_ref = C.byref(i)
        while i.pos < n:
            self._check(self._z.ZSTD_compressStream(self._c, self._o_ref, i_ref))
            self._drain()
        self.written += n

    def flush(self) -> None:
        r = 1
        while r:
            r = self._check(self._z.ZSTD_flushStream(self._c, self._o_ref))
            self._drain()

    def finish(self) -> bytes:
        r = 1
        while r:
            r = self._check(self._z.ZSTD_endStream(self._c, self._o_ref))
            self._drain()
        self._z.ZSTD_freeCCtx(self._c)
        self._c = None
        return bytes(self._out)

    def __del__(self):
        if getattr(self, "_c", None):
            self._z.ZSTD_freeCCtx(self._c)
