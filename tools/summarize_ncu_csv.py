"""Summarise an `ncu --set full` capture exported on the GPU box as CSV (raw page + source page) into profiles/.

    ncu -i prof.ncu-rep --page raw --csv > X_raw.csv ; ncu -i prof.ncu-rep --page source --csv > X_source.csv     (on the box)
    python tools/summarize_ncu_csv.py gpurun_out/.../X profiles/r02_X                                                (here)

(The .ncu-rep files are 40 MB each -- over what gpurun copies back -- hence the on-box export.)"""
import collections, csv, json, re, sys

src_prefix, out_prefix = sys.argv[1], sys.argv[2]
raw_rows = list(csv.reader(open(src_prefix + "_raw.csv")))
raw = {h: (v, u) for h, u, v in zip(raw_rows[0], raw_rows[1], raw_rows[2])}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum"]
metrics = {k: {"value": raw[k][0], "unit": raw[k][1]} for k in keys if k in raw}
stalls = {k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(v[0]) for k, v in raw.items()
          if "average_warps_issue_stalled" in k and "per_issue_active" in k and v[0] not in ("", "n/a")}
rows = list(csv.reader(open(src_prefix + "_source.csv")))
kernel = rows[0][1] if len(rows[0]) > 1 else ""
hdr, body = rows[1], rows[2:]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
scols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = sum(int(r[si] or 0) for r in body) or 1
tot_i = sum(int(r[ie] or 0) for r in body)
hot = []
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:16]:
    st = sorted(((h, int(r[hdr.index(h)] or 0)) for h in scols), key=lambda kv: -kv[1])[:2]
    hot.append({"pct_of_stall_samples": round(100 * int(r[si] or 0) / tot_s, 1), "executed": int(r[ie] or 0), "sass": r[so].strip(),
                "top_stalls": [f"{h}={v}" for h, v in st if v]})
mn = collections.Counter()
dyn = collections.Counter()
for r in body:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", r[so])
    if m:
        base = m.group(1)
        mn[base] += 1
        dyn[base.split(".")[0]] += int(r[ie] or 0)
watch = ["LDGSTS", "UBLKCP", "UTMALDG", "UTCHMMA", "LDTM", "LDS", "STS", "STG", "LDG", "FADD2", "FMUL2", "FMNMX3", "FMNMX", "FFMA", "FMUL", "FADD",
         "MUFU", "SHFL", "IMAD", "BRA", "DEPBAR", "LDGDEPBAR", "REDG", "ATOMG", "BAR"]
summary = {"kernel": kernel, "metrics": metrics, "stall_cycles_per_issued_instruction": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10]),
           "warp_instructions_executed": tot_i, "hottest_sass": hot,
           "executed_by_mnemonic": {k: dyn[k] for k in watch if dyn.get(k)},
           "static_mnemonics": {k: v for k, v in mn.most_common(40)}}
json.dump(summary, open(out_prefix + ".json", "w"), indent=1)
with open(out_prefix + ".txt", "w") as f:
    f.write(kernel + "\n")
    for k, v in metrics.items():
        f.write(f"  {k:70s} {v['value']} {v['unit']}\n")
    f.write("stall cycles per issued instruction: " + ", ".join(f"{k} {v:.2f}" for k, v in list(summary["stall_cycles_per_issued_instruction"].items())) + "\n")
    f.write(f"warp instructions executed: {tot_i}\nexecuted by mnemonic: " + ", ".join(f"{k} {v}" for k, v in summary["executed_by_mnemonic"].items()) + "\n")
    f.write("hottest SASS (share of stall samples, times executed):\n")
    for h in hot:
        f.write(f"  {h['pct_of_stall_samples']:5.1f}%  x{h['executed']:<10d} {h['sass'][:70]:70s} {' '.join(h['top_stalls'])}\n")
print(open(out_prefix + ".txt").read()[:2500])
