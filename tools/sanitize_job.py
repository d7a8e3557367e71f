"""A small but complete tour of the device code for compute-sanitizer: fixtures (all fields, text formatter), a FASTQ
archive that takes the finisher path, a masked genome, a zstd frame with long matches."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nafcodec_b200 as N
import _cases as K
import _oracle as O
from _harness import assert_same_as_oracle
from conftest import read_golden

datas = [read_golden(n) for n in ("masked.naf", "phix.naf", "LuxC.naf", "CP040672.naf")]
datas += [K.fastq_reads(3, 3000), K.genome(4, 200_000, level=19, gaps=2, gap_len=20000)]
res = N.decode_batch(datas)
for r, d in zip(res, datas):
    assert_same_as_oracle(r, O.decode(d), "sanitize")
for d in datas:
    assert N.to_text(d) == O.format_text(d)
print("sanitize tour ok")
