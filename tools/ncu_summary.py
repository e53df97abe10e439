"""Turn ncu outputs into the markdown tables kept under profiles/ (runs on the CPU box: `ncu -i` needs no GPU).

  python tools/ncu_summary.py launches gpurun_out/launches.csv        # per-kernel totals / shares of a launch list
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep            # key metrics of a --set full capture, one column per kernel
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict, defaultdict

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__cluster_size", "launch__occupancy_limit_shared_mem",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("mg::", "")


def rows_of(text):
    lines = [l for l in text.splitlines() if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("\n".join(lines))))


def launches(path):
    rows = rows_of(open(path).read())
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values())
    print(f"{sum(cnt.values())} launches, {total:.2f} ms in total.\n")
    print("| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"| `{k}` | {cnt[k]} | {tot[k]:.3f} | {tot[k] / cnt[k]:.4f} | {100 * tot[k] / total:.1f} % |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = OrderedDict()
    for r in rows[2:]:
        seen.setdefault(short(r[ki]), r)          # first captured launch of every kernel
    names = list(seen)
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---|" * len(names))
    for m in KEY:
        if m not in hdr:
            continue
        i = hdr.index(m)
        print(f"| {m} [{units[i]}] | " + " | ".join(seen[n][i] for n in names) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
