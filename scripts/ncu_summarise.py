"""Turn ncu CSV logs (gpurun_out/*.csv) into the small summaries committed under profiles/.

    python scripts/ncu_summarise.py launches gpurun_out/r2_launches.csv profiles/r2_ncu_launch_list_summary.csv "<command>"
    python scripts/ncu_summarise.py traffic  gpurun_out/r2_gemm_traffic.csv profiles/r2_gemm_dram_traffic.json "<command>"
    python scripts/ncu_summarise.py raw      gpurun_out/r2_channel_raw.csv profiles/r2_ncu_channel_summary.csv "<command>"
"""
import collections, csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(path):
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    return H, [r for r in rows[h + 1:] if len(r) == len(H)]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    return name.split("(")[0]


def commit():
    return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()


def val(r, H):
    v = float(r[H.index("Metric Value")].replace(",", ""))
    u = r[H.index("Metric Unit")]
    scale = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v * scale


def launches(src, dst, cmd):
    H, rows = rows_of(src)
    agg = collections.OrderedDict()
    for r in rows:
        if r[H.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r[H.index("Kernel Name")]), [0, 0.0])
        a[0] += 1
        a[1] += val(r, H)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {cmd} (commit {commit()}; serialised, cold cache: compare SHARES, not absolute times)\n")
        f.write("kernel,launches_in_capture,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.1f},{t / tot:.4f}\n")
    print(open(dst).read())


def traffic(src, dst, cmd):
    H, rows = rows_of(src)
    rd = wr = t = 0.0
    ids = set()
    for r in rows:
        m = r[H.index("Metric Name")]
        ids.add(r[0])
        if m == "dram__bytes_read.sum":
            rd += val(r, H)
        elif m == "dram__bytes_write.sum":
            wr += val(r, H)
        elif m == "gpu__time_duration.sum":
            t += val(r, H)
    n = len(ids)
    out = {"source": cmd, "commit": commit(), "gemm_launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "dram_bytes_per_launch_avg": (rd + wr) / max(n, 1), "ncu_time_s": t * 1e-6}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def raw(src, dst, cmd):
    rows = list(csv.reader(open(src, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H, units, data = rows[h], rows[h + 1], rows[h + 2:]
    with open(dst, "w") as f:
        f.write(f"# {cmd} (commit {commit()})\n")
        for d in data:
            if len(d) != len(H):
                continue
            f.write(f"kernel,{short(d[H.index('Kernel Name')])}\nmetric,unit,value\n")
            for k in KEEP:
                if k in H:
                    f.write(f"{k},{units[H.index(k)]},{d[H.index(k)]}\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
