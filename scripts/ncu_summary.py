"""Summarise an .ncu-rep: key metrics + executed-instruction bands from the source page."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:70])
    for k in keys:
        if k in hdr:
            print(f"  {k:70s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
    st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[hdr.index(k)]))
          for k in hdr if "issue_stalled" in k and "per_issue_active" in k]
    print("  stalls:", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st, key=lambda kv: -kv[1])[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for n, start in enumerate(hi):
    end = hi[n + 1] - 1 if n + 1 < len(hi) else len(rows)
    h = rows[start]
    ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
    data = [(int(r[ie]), int(r[isamp])) for r in rows[start + 1:end] if len(r) > ie and r[ie].isdigit()]
    tot = sum(d[0] for d in data)
    bands, sb = collections.Counter(), collections.Counter()
    for e, s in data:
        bands[e] += 1
        sb[e] += s
    print(f"-- kernel {n}: {tot} warp instructions, {sum(d[1] for d in data)} samples")
    for e, c in sorted(bands.items(), key=lambda kv: -kv[0] * kv[1])[:6]:
        print(f"   exec={e:10d} x {c:4d} instrs -> {e * c / max(tot, 1) * 100:5.1f}% inst, samples {sb[e]}")
