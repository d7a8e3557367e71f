"""Pageable against pinned INPUT for the single-archive call: nafgpu_decode of one archive whose bytes
live in ordinary memory (ctypes copy, what Decoder(bytes) hands over) or in nafgpu_host_alloc'd memory.  One JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _cases  # noqa: E402
import nafcodec_b200 as N  # noqa: E402
from nafcodec_b200 import _ffi  # noqa: E402


def main():
    lib = _ffi.default_library()
    ctx = N.Context(0, lib)
    out = {}
    for label, n in (("5Mbp", 5_000_000), ("50Mbp", 50_000_000)):
        data = _cases.genome(21, n, level=3)
        p = lib.dll.nafgpu_host_alloc(len(data))
        C.memmove(p, data, len(data))
        pinned = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(len(data),))
        arch = {"pageable": N.parse_archive(data, lib), "pinned": N.parse_archive(pinned, lib)}
        res = (_ffi.Result * 1)()
        row = {"archive_bytes": len(data)}
        for kind, a in arch.items():
            arr = (_ffi.Archive * 1)(a)
            best = 1e9
            for _ in range(12):
                t0 = time.perf_counter()
                rc = lib.dll.nafgpu_decode_batch(ctx._ctx, arr, 1, _ffi.WANT_ALL, res)
                best = min(best, time.perf_counter() - t0)
                assert rc == 0 and res[0].total_residues == n
            row[kind + "_ms"] = round(best * 1e3, 3)
        out[label] = row
        lib.dll.nafgpu_host_free(p)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
