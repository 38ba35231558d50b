#!/usr/bin/env python
"""Correlate an ncu report (--import-source on, built with -lineinfo) with source lines.
usage: ncu_hot_lines.py report.ncu-rep kernel_substring [topN] [ncu_kernel_regex]
Prints headline metrics and the hottest source lines (by executed warp instructions)."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))


def run(cmd):
    return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode(errors="replace")


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    # ncu-side kernel filter (reports with several kernels): a kernel-name regex, or "#N" = the N-th launch of the report
    kf = ""
    if len(sys.argv) > 4:
        kf = (" --launch-skip %s --launch-count 1" % sys.argv[4][1:]) if sys.argv[4].startswith("#") else (" -k regex:%s" % sys.argv[4])
    raw = list(csv.reader(run("ncu -i %s --page raw --csv%s" % (rep, kf)).splitlines()))
    hdr, units, vals = raw[0], raw[1], raw[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic"]
    for i, h in enumerate(hdr):
        if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.2):
            print("%-90s %-10s %s" % (h, units[i], vals[i]))
    tmp = tempfile.mkdtemp()
    run("cd %s && cuobjdump -xelf all %s/nextpolish_b200/lib/nextpolish1.so" % (tmp, ROOT))
    dis = run("nvdisasm -g -c %s/%s.sm_100a.cubin" % (tmp, os.environ.get("NCU_CUBIN", "engine"))).split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
    cur, seq = None, {}
    for l in dis[start + 1:]:
        if l.startswith(".text.") and seq:
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            seq[int(m.group(1), 16)] = cur
    rows = list(csv.reader(run("ncu -i %s --page source --csv%s" % (rep, kf)).splitlines()))
    h = rows[1]
    ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    data = []
    for r in rows[2:]:                       # the page repeats itself per view / kernel: keep the first block
        if r and r[0] in ("Kernel Name", "Address"):
            break
        if len(r) > ie:
            data.append(r)
    base = int(data[0][ia], 16)
    agg, samp = collections.Counter(), collections.Counter()
    for r in data:
        fl = seq.get(int(r[ia], 16) - base)
        agg[fl] += float(r[ie] or 0)
        samp[fl] += float(r[isamp] or 0)
    tot, ts = sum(agg.values()), sum(samp.values())
    src = {}
    for f in os.listdir(os.path.join(ROOT, "nextpolish_b200", "csrc")):
        if f.endswith((".h", ".cu")):
            src[f] = open(os.path.join(ROOT, "nextpolish_b200", "csrc", f)).read().split("\n")
    # by function: line ranges of top-level definitions in each source file
    import bisect
    starts = {}
    for f, lines in src.items():
        st = [(i + 1, l.strip()[:70]) for i, l in enumerate(lines) if re.match(r"(NP_HD|template|struct|__global__|static|inline)\b", l)]
        starts[f] = st
    fagg, fsamp = collections.Counter(), collections.Counter()
    for fl, v in agg.items():
        key = fl
        if fl and fl[0] in starts and starts[fl[0]]:
            st = starts[fl[0]]
            j = bisect.bisect_right([a for a, _ in st], fl[1]) - 1
            # a "template <...>" line is followed by the definition it belongs to
            name = st[j][1] if j >= 0 else "?"
            if name.startswith("template") and j + 1 < len(st) and st[j + 1][0] == st[j][0] + 1: name = st[j + 1][1]
            key = (fl[0], name)
        fagg[key] += v; fsamp[key] += samp[fl]
    print("\nby function (warp instructions / samples):")
    for k, v in fagg.most_common(25):
        print("%5.1f%% inst %5.1f%% samp  %s" % (100 * v / tot, 100 * fsamp[k] / ts, k))
    print("\nwarp instructions: %.0f, samples: %.0f" % (tot, ts))
    for fl, v in agg.most_common(top):
        t = src.get(fl[0], [""] * 100000)[fl[1] - 1].strip()[:90] if fl and fl[0] in src else ""
        print("%5.1f%% inst %5.1f%% samp  %s:%s | %s" % (100 * v / tot, 100 * samp[fl] / ts, fl[0] if fl else None, fl[1] if fl else None, t))


if __name__ == "__main__":
    main()
