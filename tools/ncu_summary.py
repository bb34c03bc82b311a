"""Summarise an `ncu --set full` report and an `ncu --metrics gpu__time_duration.sum` launch list
into markdown for profiles/.   python tools/ncu_summary.py <prof.ncu-rep> <launches.csv> > profiles/<name>.md"""
import collections, csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__cycles_elapsed.avg", "elapsed cycles"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "mio_throttle", "lg_throttle", "math_pipe_throttle",
          "not_selected", "branch_resolving", "no_instruction", "sleeping", "membar", "dispatch_stall"]


def main(rep, launches):
    print("# ncu summary: %s\n" % rep)
    if launches:
        rows = [l for l in open(launches) if l.startswith('"')]
        r = list(csv.DictReader(io.StringIO("".join(rows))))
        agg = collections.OrderedDict()
        for x in r:
            k = x["Kernel Name"].split("(")[0][-70:]
            v = float(x["Metric Value"].replace(",", ""))
            u = x["Metric Unit"]
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            agg.setdefault(k, [0, 0.0])
            agg[k][0] += 1
            agg[k][1] += v
        tot = sum(v[1] for v in agg.values())
        print("## launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
        print("| kernel | launches | avg us | share |\n|---|---:|---:|---:|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("| `%s` | %d | %.1f | %.3f |" % (k, v[0], v[1] / v[0], v[1] / tot))
        print()
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    seen = set()
    print("## per-kernel metrics (`--set full --clock-control none`, one launch each)\n")
    for row in r[2:]:
        name = row[hdr.index("Kernel Name")]
        short = name.split("(")[0][-80:]
        if short in seen:
            continue
        seen.add(short)
        print("### `%s`\n" % short)
        print("| metric | value |\n|---|---|")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print("| %s | %s %s |" % (label, row[i], units[i]))
        st = []
        for s in STALLS:
            key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
            if key in hdr:
                st.append((float(row[hdr.index(key)] or 0), s))
        print("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (s, v) for v, s in sorted(st, reverse=True)[:5]))
        print()


# kernel name fragments -> (variant, bench kernel kind) for profiles/traffic.json
KINDS = [("k_v3_c1_reg", ("v3", "front")), ("k_conv_slab<ConvTcCfg<30, 2,", ("v3", "conv2")), ("k_conv_slab<ConvTcCfg<28, 3,", ("v3", "conv3")),
         ("k_fc4_tc", ("v3", "fc4")), ("k_tail_tc", ("v3", "tail")), ("k_slim_c1_reg", ("v3_slim", "front")),
         ("k_conv_slab<ConvTcCfg<35, 3,", ("v3_slim", "conv2")), ("k_conv_slab<ConvTcCfg<37, 5,", ("v3_slim", "conv3")),
         ("k_gemm_tc<48,", ("v3_slim", "fc4")), ("k_tail<", ("v3_slim", "tail")), ("k_tail_site<", ("v3_slim", "tail"))]


def _norm(name):
    """kernel name as ncu prints it, without namespaces and C-style casts: `k_conv_slab<ConvTcCfg<30, 2, ...`"""
    for t in ("(int)", "(bool)", "cvb::tc::", "cvb::", "tc::"):
        name = name.replace(t, "")
    return name


def traffic(rep, out_fn, source_hash):
    """profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch of every forward kernel in an
    `ncu --set full` report (the LARGEST launch of each kernel = a full chunk), stamped with the hash of the build"""
    import json
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    ir, iw, it, ik = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "Kernel Name"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    res = {"source_hash": source_hash, "report": rep, "v3": {}, "v3_slim": {}, "_detail": {}}
    for row in r[2:]:
        for frag, (variant, kind) in KINDS:
            if frag in _norm(row[ik]):
                b = float(row[ir]) * scale[units[ir]] + float(row[iw]) * scale[units[iw]]
                if b > res[variant].get(kind, 0):
                    res[variant][kind] = b
                    res["_detail"]["%s/%s" % (variant, kind)] = dict(dram_read=float(row[ir]) * scale[units[ir]],
                                                                     dram_write=float(row[iw]) * scale[units[iw]],
                                                                     duration=row[it] + " " + units[it])
                break
    json.dump(res, open(out_fn, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--traffic":      # ncu_summary.py --traffic <rep> <out.json> <source_hash>
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
