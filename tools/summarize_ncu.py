"""Summarise an .ncu-rep (read here, without a GPU) into profiles/<name>.json + .txt: duration, DRAM bytes, stall
reasons, the hottest SASS lines.  usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_k_rows"""
import csv, io, json, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
kernels = []
for vals in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    def num(k):
        try:
            return float(d[k][0].replace(",", ""))
        except Exception:
            return None
    def bytes_of(k):
        v = num(k)
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d[k][1], 1)
    stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: num(k) for k in d
              if "smsp__average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k}
    kernels.append({
        "kernel": d.get("Kernel Name", ("?",))[0], "grid": d.get("launch__grid_size", ("?",))[0],
        "block": d.get("launch__block_size", ("?",))[0], "registers_per_thread": num("launch__registers_per_thread"),
        "duration_us": (num("gpu__time_duration.sum") or 0) / 1e3 if d.get("gpu__time_duration.sum", ("", ""))[1] in ("nsecond", "ns") else num("gpu__time_duration.sum"),
        "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
        "dram_throughput_pct_of_hw_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "sm_warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "inst_executed": num("smsp__inst_executed.sum"), "ipc_active": num("sm__inst_executed.avg.per_cycle_active"),
        "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
        "stall_cycles_per_issue": dict(sorted(((k, v) for k, v in stalls.items() if v), key=lambda kv: -kv[1])[:8]),
    })
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hot = []
if len(srows) > 2:
    h = srows[1]
    ia, iex, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    data = [r for r in srows[2:] if len(r) > isamp and r[isamp].isdigit()]
    ts = sum(int(r[isamp]) for r in data) or 1
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:15]:
        hot.append({"sass": r[ia].strip(), "stall_samples_pct": round(100 * int(r[isamp]) / ts, 1), "executed": int(r[iex])})
    mn = collections.Counter()
    for r in data:
        mn[r[ia].split()[0].lstrip("@!P0123456789U ") if not r[ia].strip().startswith("@") else r[ia].split()[1]] += 1
    tma = {k: v for k, v in mn.items() if any(t in k for t in ("UBLKCP", "UTMA", "LDGSTS", "SYNCS", "FADD2", "FMUL2", "FMNMX3", "LDS", "STG"))}
else:
    tma = {}
json.dump({"report": rep, "kernels": kernels, "hottest_sass_first_kernel": hot, "sass_mnemonics_first_kernel": tma}, open(out + ".json", "w"), indent=1)
with open(out + ".txt", "w") as f:
    for k in kernels:
        f.write(f"{k['kernel']}\n  grid {k['grid']} block {k['block']} regs {k['registers_per_thread']}\n")
        tot = (k['dram_bytes_read'] or 0) + (k['dram_bytes_write'] or 0)
        f.write(f"  duration {k['duration_us']:.1f} us   DRAM read {(k['dram_bytes_read'] or 0)/1e6:.1f} MB  write {(k['dram_bytes_write'] or 0)/1e6:.1f} MB  total {tot/1e6:.1f} MB\n")
        if k['duration_us']:
            f.write(f"  DRAM traffic rate {tot/k['duration_us']/1e3:.0f} GB/s   (ncu: {k['dram_throughput_pct_of_hw_peak']} % of HW peak)\n")
        f.write(f"  warps active {k['sm_warps_active_pct']} %   IPC(active) {k['ipc_active']}   L2 hit {k['l2_hit_rate_pct']} %\n")
        f.write(f"  stall cycles per issued instruction: {k['stall_cycles_per_issue']}\n")
    f.write("hottest SASS (stall samples):\n")
    for hline in hot:
        f.write(f"  {hline['stall_samples_pct']:5.1f}%  x{hline['executed']:<9d} {hline['sass']}\n")
    f.write(f"mnemonics: {tma}\n")
print(open(out + ".txt").read())
