"""Per-kernel SASS mnemonic histogram of libpna_sm100.so (cuobjdump -sass), written to profiles/.

    python tools/sass_histogram.py [--out profiles/r02_sass_histogram]

The PTX names never appear in SASS (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit ->
UTCBAR, cp.async.bulk -> UBLKCP, cp.async.bulk.tensor -> UTMALDG/UTMASTG, cp.async -> LDGSTS, red.global -> REDG.
Runs on the CPU box (no GPU needed)."""
import argparse, collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "LDGDEPBAR",
         "SYNCS", "HMMA", "REDG", "ATOMG", "RED", "FADD2", "FMUL2", "FFMA2", "FMNMX3", "FMNMX", "FFMA", "FADD", "FMUL", "MUFU", "LDG", "STG",
         "LDS", "STS", "SHFL", "BAR", "LDL", "STL"]
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(ROOT, "pna_b200", "libpna_sm100.so"))
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_sass_histogram"))
args = ap.parse_args()
sass = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels, cur = collections.OrderedDict(), None
arch = set(re.findall(r"arch = (sm_\w+)", sass))
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        kernels[m.group(1)] = cur
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
rows = []
for name, cnt in kernels.items():
    short = demangle(name)
    short = short[: short.rfind("(")] if "(" in short else short          # drop the parameter list
    short = short.replace("void ", "").replace("(int)", "").replace("(bool)", "").replace("(unsigned int)", "")
    short = re.sub(r"pna::CfgStatic<([^>]*)>", r"Cfg<\1>", short).replace("pna::", "")
    rec = {"kernel": short[:160], "instructions": sum(cnt.values())}
    rec.update({k: cnt[k] for k in WATCH if cnt.get(k)})
    rows.append(rec)
rows.sort(key=lambda r: -r["instructions"])
summary = {"lib": os.path.relpath(args.lib, ROOT), "arch": sorted(arch), "kernels": len(rows),
           "totals": {k: sum(r.get(k, 0) for r in rows) for k in WATCH if any(r.get(k) for r in rows)}}
json.dump({"summary": summary, "kernels": rows}, open(args.out + ".json", "w"), indent=1)
with open(args.out + ".txt", "w") as f:
    f.write(f"SASS mnemonic histogram of {summary['lib']} (cuobjdump -sass); ELF arch: {', '.join(summary['arch'])}; {len(rows)} kernels\n")
    f.write("totals: " + ", ".join(f"{k} {v}" for k, v in summary["totals"].items()) + "\n\n")
    pick = lambda r: any(r.get(k) for k in ("UTCHMMA", "LDTM", "UBLKCP", "UTMALDG", "LDGSTS", "REDG", "ATOMG")) or r["instructions"] > 1500
    for r in rows:
        if pick(r):
            f.write(f"{r['kernel']}\n    " + ", ".join(f"{k} {v}" for k, v in r.items() if k != "kernel") + "\n")
print(open(args.out + ".txt").read()[:3000])
