#!/usr/bin/env python
"""Prints the headline metrics, opcode mix and stall breakdown of an .ncu-rep (first profiled launch)."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
KFILTER = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []   # which kernel of a multi-kernel report
KERNEL_INDEX = 0
raw = subprocess.run(["ncu", "-i", rep] + KFILTER + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
R = rows[2 + KERNEL_INDEX]
print("kernel:", R[hdr.index('Kernel Name')][:100])
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"  {w:70s} {units[i]:10s} {R[i]}")
src = subprocess.run(["ncu", "-i", rep] + KFILTER + ["--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
k = min(KERNEL_INDEX, len(hi) - 1)
h = rows[hi[k]]
end = hi[k + 1] - 1 if len(hi) > k + 1 else len(rows)
body = [r for r in rows[hi[k] + 1:end] if len(r) == len(h)]
ci = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ci['Instructions Executed']]) for r in body)
op = collections.Counter()
for r in body:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ci['Source']])
    op[m.group(2).split('.')[0] if m else '?'] += int(r[ci['Instructions Executed']])
print(f"  executed warp instructions {tot}; top opcodes:", ", ".join(f"{o} {100*c/tot:.1f}%" for o, c in op.most_common(14)))
st = {n: sum(int(r[ci[n]] or 0) for r in body) for n in h if n.startswith('stall_') and 'Not Issued' not in n}
s = sum(st.values())
print("  stalls:", ", ".join(f"{k[6:]} {100*v/s:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
body.sort(key=lambda r: -int(r[ci['# Samples']]))
print("  hottest instructions (samples):")
for r in body[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    top = max(((n, int(r[ci[n]] or 0)) for n in st), key=lambda kv: kv[1])
    print(f"    {r[ci['# Samples']]:>6s} {r[ci['Source']].strip()[:80]:80s} {top[0][6:]}={top[1]}")
