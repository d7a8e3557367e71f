"""Decodes one cfg4-shape FASTQ archive (all fields) a few times: the command ncu wraps for the text-like kernels
(k_lz_finish, k_frame_scan, k_decode_sequences on 10^5 tiny blocks); see profiles/README.md."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nafcodec_b200 as N
import _cases as K

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
data = K.fastq_reads(11, n_reads)
for _ in range(reps):
    res = N.decode_batch([data])
print("decoded", res[0].n_records, "reads")
