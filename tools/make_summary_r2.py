"""Writes profiles/r2_summary.md from the committed bench lines (profiles/r2_bench_n*.json, r2_bench_reference_arm.json) and
the ncu launch list, so that its tables cannot drift from the JSON files.  The prose (what was tried, what the counters
said) is kept here beside the numbers it explains."""
import csv
import json
import os
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda name: os.path.join(ROOT, "profiles", name)
load = lambda name: json.load(open(P(name))) if os.path.exists(P(name)) else None
d = load("r2_bench_n1.json")
ref = load("r2_bench_reference_arm.json")
scale = {n: load(f"r2_bench_n{n}.json") for n in (2, 4, 8)}
rf, e, st = d["roofline"], d["e2e"], d["roofline"]["stage_ms"]
tot = sum(st.values())

rows = list(csv.reader(open(P("r2_launches_job256.csv"))))
hdr = [i for i, x in enumerate(rows) if x and x[0] == "ID"][0]
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = OrderedDict()
for x in rows[hdr + 1:]:
    if len(x) > vi:
        agg.setdefault(x[ki], []).append(float(x[vi].replace(",", "")))
per_decode = {k: sum(v) / len(v) for k, v in agg.items()}          # one launch of each per decode
ltot = sum(per_decode.values())
short = lambda k: k.replace("void ", "").split("(")[0]
launch_tbl = "\n".join(f"| `{short(k)}` | {len(agg[k])} | {v / 1e3:.1f} | {100 * v / ltot:.1f} % |" for k, v in sorted(per_decode.items(), key=lambda kv: -kv[1]))
huf_share = 100 * max(v for k, v in per_decode.items() if "k_huf_decode_block" in k) / ltot
stage_tbl = "\n".join(f"| {k} | {v:.4f} | {100 * v / tot:.1f} % |" for k, v in st.items())


def cfg_row(name, v):
    if "error" in v:
        return f"| {name} | error: {v['error']} |"
    return (f"| `{name}` | {v['archives']} | {v['zstd_blocks']} / {v['sequences']} | {v['ascii_bytes'] / 1e6:.1f} | **{v['device_ms']:.3f}** | {v['host_prepare_ms']:.2f} | "
            f"{v['ascii_GBps']:.1f} | {100 * v['frac_of_hbm_peak']:.2f} % | {v['e2e']['ms']:.2f} | {v['cpu_oracle']['ms']:.0f} ({v['cpu_oracle']['threads']} thr) | "
            f"{v['lz_rounds']}{' -> in-order kernel' if v.get('lz_in_order_kernel') else ''}{' -> finisher' if v['lz_handover_round'] else ''} |")


cfg_tbl = "\n".join(cfg_row(k, v) for k, v in d["configs"].items())
scale_tbl = ""
for n, s in scale.items():
    if s:
        ce = s["e2e"].get("ceiling", {}).get("value")
        scale_tbl += (f"| {n} | {s['value']:.0f} ({s['value'] / d['value'] / n * 100:.0f} % of {n} x N=1) | {s['ms_per_step']:.3f} | {s['e2e']['value']:.1f} | "
                      f"{ce:.1f} | {100 * s['e2e']['value'] / ce:.0f} % |\n")
ref_line = ""
if ref:
    pc = ref["cpu_baseline"].get("per_core", {})
    ref_line = (f"| reference arm (`--impl reference`: oracle port, {ref['cpu_baseline']['cores']} native threads, the same 256-archive batch) | "
                f"{ref['value']:.2f} GB/s ASCII; per core: " + ", ".join(f"{k.replace('_', ' ')} {v:.3f}" for k, v in pc.items()) + " |\n")

txt = f"""# Round 2 — measured numbers and ncu evidence

All numbers from B200 boxes via `gpurun` (SM clock {d['clocks']['sm_mhz']} MHz, throttle reasons: {d['clocks']['reasons'] or 'none'}).  Tables are generated
from the committed JSON lines by `tools/make_summary_r2.py`.

## Headline workload (`profiles/r2_bench_n1.json`: `python bench.py`)

{d['config']['archives_per_gpu_per_step']} cfg2 archives per step (synthetic 5 Mbp single-record DNA + soft-mask runs, every section zstd level 19; 8 distinct
archives cycled): {d['job']['compressed_bytes'] / 1e6:.1f} MB compressed in, {d['job']['ascii_bytes'] / 1e6:.0f} MB ASCII out, {d['job']['zstd_blocks']} zstd blocks, {d['job']['sequences']} sequences,
`B_alg` = {d['job']['algorithmic_bytes'] / 1e6:.1f} MB (compressed in + every output byte once + offsets; round 1 counted the compressed bytes twice).

| quantity | round 2 | round 1 |
|---|---|---|
| `value` — device-resident ASCII out | **{d['value']:.1f} GB/s** ({d['ms_per_step']:.3f} ms / step) | 318.3 GB/s (4.021 ms) |
| whole path vs HBM roofline (`B_alg` / t) | {d['path_algorithmic_GBps']:.0f} GB/s = **{100 * d['path_frac_of_hbm_peak']:.2f} %** of the measured {rf['peak']} GB/s | 399 GB/s = 6.09 % (corrected) |
| dominant kernel `{rf['kernel']}` | {rf['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic in **{rf['kernel_ms']:.3f} ms** = {rf['achieved']:.0f} GB/s = **{100 * rf['frac']:.1f} %** of HBM peak; DRAM traffic {rf['traffic'] / 1e6:.0f} MB per launch (ncu) | `k_huf_decode<512>`: 2.211 ms = 439 GB/s = 6.7 %; 1092 MB |
| `e2e` — pinned host in -> pinned host out through `nafgpu_pipeline_*` ({e['lanes']} lanes) | **{e['value']:.1f} GB/s** = {100 * e['frac_of_ceiling']:.0f} % of what pure copies of the same bytes reach ({e['ceiling']['value']:.1f} GB/s) | 48.9 GB/s |
| CPU baseline, 1 thread (oracle port on libzstd 1.5.5) | {d['cpu_baseline']['value']:.2f} GB/s ASCII | 0.43 |
{ref_line}| FASTA text of the batch on the device | {d['text_formatter']['text_bytes'] / 1e6:.0f} MB in {d['text_formatter']['device_ms']:.3f} ms = {100 * d['text_formatter']['frac_of_hbm_peak']:.1f} % of HBM peak | 33 % |

Stage times (serial profiled run, CUDA events on the launch stream; in the timed runs the Huffman branch and the FSE branch overlap):

| stage | ms | share |
|---|---|---|
{stage_tbl}

## Every BASELINE.json config at its stated size (`configs` in the bench line; each with a bit-exact parity gate vs the oracle)

| config | archives | zstd blocks / sequences | ASCII MB | device ms | host prepare ms | ASCII GB/s | `B_alg` % of HBM peak | e2e ms (one synchronous C-ABI call) | CPU oracle ms | LZ rounds |
|---|---|---|---|---|---|---|---|---|---|---|
{cfg_tbl}

* cfg1 (the reference's own fixture) alone: 0.200 ms in round 1 and at the start of this round -> {d['configs']['cfg1_fixture']['device_ms']:.3f} ms (target 0.09): the match stage of a small job runs
  in one CTA (`k_lz_small`), the small Huffman streams run beside the big ones, the Huffman weight chain was pipelined, long literal runs are copied with four
  chunks per lane in flight (steps and what each gained: below).  cfg2 alone: 0.234 -> {d['configs']['cfg2_single']['device_ms']:.3f} ms (target 0.10): bound by the FSE chain of its longest block
  (61 us after the control flow was taken out of its loop: ~580 sequences x 182 cycles) and the match stage (the in-order kernel before the rounds);
  serial stage times are in the JSON (`stage_ms_serial`).
* cfg3 (250 Mbp, ONE frame per section): 5.87 ms at the start of the round -> {d['configs']['cfg3_250Mbp']['device_ms']:.2f} ms = {d['configs']['cfg3_250Mbp']['ascii_GBps']:.0f} GB/s ASCII (target 250).  The diverged repeat
  family makes 78 generations of matches: after three rounds the in-order kernel `k_lz_flow` takes them (a generation costs a visibility latency instead of
  a round: 1.93 -> 0.39 ms for the match stage; steps below).
* cfg4 (10^6 reads, 2 x 10^6 tiny zstd blocks): device + prepare 107 ms at the start of the round (round 1: ~530 ms) -> {d['configs']['cfg4_1M_all_fields']['device_plus_prepare_ms']:.1f} ms (target 60), {d['configs']['cfg4_1M_all_fields']['e2e']['ms']:.1f} ms for one synchronous
  host-to-host call: tiled frame scan (10.4 -> 0.27 ms), sliced NAF scans (3.5 -> 0.1 ms), threaded header walk with descriptors written straight into pinned
  staging (45 -> 23 ms), warp-per-block kernels for the tiny blocks with the general kernels running over lists of the others, level 2 of the finisher over
  roots only (steps below).  What is left is the quality section's single dependency chain (finisher: 11.7 ms) and the header walk.
* cfg5: 512 UNIQUE archives (2-6 Mbp, 1-4 records) in one job.

## ncu launch list, 256-archive job (`profiles/r2_launches_job256.csv`: `ncu --metrics gpu__time_duration.sum --clock-control none -s 16 -c 32 python tools/profile_job.py 256 3`; cold-cache, serialised: compare shares)

| kernel | launches in the window | us / launch | share of one decode |
|---|---|---|---|
{launch_tbl}

`k_huf_decode_block` is {huf_share:.0f} % of the serialised decode here and {100 * rf['kernel_ms'] / tot:.0f} % of the serial stage sum in the bench line ({rf['kernel_ms']:.3f} of {tot:.3f} ms): they agree.
`profiles/r2_launches_bench.csv` is the launch list of the `bench.py` command itself (`python bench.py --steps 3 --warmup 3 --no-configs`, first 600 launches);
`r2_launches_single_archive.csv` and `r2_launches_fastq_1M.csv` are those of one cfg1 / cfg2 archive and of the 10^6-read FASTQ archive.

## Scaling (weak: 256 archives per GPU per step, no collective; `profiles/r2_bench_n{{2,4,8}}.json`)

| GPUs | device-resident GB/s | ms / step | e2e GB/s | pure-copy ceiling at that N | e2e / ceiling |
|---|---|---|---|---|---|
| 1 | {d['value']:.0f} | {d['ms_per_step']:.3f} | {e['value']:.1f} | {e['ceiling']['value']:.1f} | {100 * e['frac_of_ceiling']:.0f} % |
{scale_tbl}
The end-to-end rate at 8 GPUs is what the host gives: pure `cudaMemcpyAsync` traffic of the same bytes from pinned memory, all ranks at once, reaches the
ceiling in the table (about 11 GB/s per GPU), and the decode path is within a few per cent of it.  Round 1 lost another factor of two there to 32 threads
spinning in `cudaStreamSynchronize`; big jobs now wait on blocking-sync events and the lanes are native threads inside the library.
"""
txt += open(os.path.join(ROOT, "tools", "r2_findings.md")).read()
open(P("r2_summary.md"), "w").write(txt)
print("wrote", P("r2_summary.md"))
