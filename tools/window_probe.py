"""Time to the first record and host memory of the two fetch modes (run on the GPU box):
  whole:   Decoder(file)                  -> the whole result crosses PCIe before record 0
  windows: Decoder(file, buffer_size=N)   -> the archive is decoded into HBM, records cross PCIe a window at a time
Prints one JSON line.  The archive is a synthetic multi-record genome (tests/_cases.genome)."""
import io
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _cases  # noqa: E402
import nafcodec_b200 as N  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    data = _cases.genome(11, n, level=3, records=64)
    out = {"residues": n, "records": 64, "archive_bytes": len(data)}
    for label, kw in (("whole", {}), ("windows_1MiB", dict(buffer_size=1 << 20)), ("windows_16MiB", dict(buffer_size=16 << 20))):
        best_first, best_all = 1e9, 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            dec = N.Decoder(io.BytesIO(data), **kw)
            first = next(dec)
            t1 = time.perf_counter()
            total = len(first.sequence)
            for r in dec:
                total += len(r.sequence)
            t2 = time.perf_counter()
            assert total == n
            best_first, best_all = min(best_first, t1 - t0), min(best_all, t2 - t0)
        out[label] = {"first_record_ms": round(best_first * 1e3, 3), "all_records_ms": round(best_all * 1e3, 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
