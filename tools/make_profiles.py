#!/usr/bin/env python
"""Turn the ncu artefacts in gpurun_out/ into the committed summaries under profiles/ (developer tool).

    python tools/make_profiles.py <tag> <launches.csv> <full.ncu-rep> [<full2.ncu-rep> ...]

Writes profiles/<tag>_launches_raw.csv (copy), <tag>_launches_summary.csv (per-kernel share),
<tag>_<kernel>_full.txt (ncu --page details), <tag>_<kernel>_stalls.txt (hottest SASS + stall mix) and
<tag>_ncu_full_summary.json (the metrics the roofline argument uses, per captured launch).
Runs in the CPU container: `ncu -i` needs no GPU.
"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PROF = os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def short(name):
    m = re.search(r"((?:corr|mcn)_\w+?)_kernel", name)
    return m.group(1) if m else re.sub(r"\W+", "_", name)[:30]


def launches(tag, path):
    shutil.copy(path, os.path.join(PROF, f"{tag}_launches_raw.csv"))
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        d[r[ik]].append(v / 1000 if r[iu].startswith("n") else v)
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(PROF, f"{tag}_launches_summary.csv"), "w") as f:
        f.write(f"# {tag}: ncu launch list summary (gpu__time_duration.sum, --clock-control none); every launch of the process\n"
                "# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n"
                f"# {sum(len(v) for v in d.values())} launches, {tot / 1000:.3f} ms total\n"
                "kernel,launches,total_us,avg_us,share_pct\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k[:90]}\",{len(v)},{sum(v):.1f},{sum(v) / len(v):.1f},{100 * sum(v) / tot:.1f}\n")


def full(tag, rep, summary):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    names = []
    for r in rows[2:]:
        m = {k: f"{r[h.index(k)]} {units[h.index(k)]}".strip() for k in WANT if k in h}
        m["report"] = os.path.basename(rep)
        summary.append(m)
        names.append(short(r[h.index("Kernel Name")]))
    key = names[0] if names else "kernel"
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(PROF, f"{tag}_{key}_kernel_full.txt"), "w").write(det)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = os.path.join("/tmp", f"{tag}_{key}_src.csv")
    open(tmp, "w").write(src)
    st = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_src.py"), tmp, "24"], capture_output=True, text=True)
    open(os.path.join(PROF, f"{tag}_{key}_kernel_stalls.txt"), "w").write(st.stdout)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag, sys.argv[2])
    summ = []
    for rep in sys.argv[3:]:
        full(tag, rep, summ)
    json.dump(summ, open(os.path.join(PROF, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    print("wrote", sorted(f for f in os.listdir(PROF) if f.startswith(tag)))
