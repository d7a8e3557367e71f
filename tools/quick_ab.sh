# A/B helper for kernel work: parity tests, then the headline bench line in short form.  usage: bash tools/quick_ab.sh <tag> [EXTRA nvcc defs to rebuild with]
set -e
tag=${1:-quick}
if [ -n "$2" ]; then touch nafcodec_b200/csrc/zstd_kernels.cu; make -s -j8 -C nafcodec_b200/csrc EXTRA="$2" 2>&1 | grep -v "^$" | head -3; fi
python -m pytest tests/test_parity.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/$tag.json 2> gpurun_out/$tag.err || tail -5 gpurun_out/$tag.err
python -c "
import json;d=json.load(open('gpurun_out/$tag.json'));print('$tag', round(d['value'],1),'GB/s', round(d['ms_per_step'],3),'ms; huf kernel', round(d['roofline']['kernel_ms'],3), 'single', round(d['single_archive']['device_us'],1), 'us; stages', d['roofline']['stage_ms'])"
