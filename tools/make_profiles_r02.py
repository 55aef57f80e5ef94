"""Turn the round-2 ncu captures in gpurun_out/ into the committed evidence under profiles/:

  profiles/r02_launches_{c4,c5}.csv   the launch lists of `bench.py --steps 1 --warmup 3` (copied) and
  profiles/r02_launches.md            their per-kernel totals (the kernel's SHARE of a step is what has to agree with the
                                      CUDA-event times of the bench line -- under ncu launches are serialised and cold);
  profiles/r02_ncu.md                 the counters of the `--set full` captures of the dominant kernels;
  profiles/r02_traffic.json           dram bytes per launch, keyed by kernel and launch shape (bench.py reads it).

    python tools/make_profiles_r02.py
"""
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe"),
    ("sm__inst_executed_pipe_fp64_op_dmma.avg.pct_of_peak_sustained_active", "fp64 tensor pipe (DMMA)"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 / shared wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
]

# (report, what it is, traffic key or None)
CAPTURES = [
    ("r02_far_c4slab.ncu-rep", "idw_far_kernel, one IDW field on a C4 slab (1024 x 10000 cells, 200 stations, 720 steps, i.i.d. dropouts) "
                               "-- `python tools/idw_far_perf.py 1024 10000 200 720`",
     {"kernel": "idw_far_kernel", "shape": "cells10240000_T720_S200"}),
    ("r02_far_c5.ncu-rep", "idw_far_kernel on the dense C5 network (4000 x 4000 cells of 10 m, the same 200 stations over 40 km: 12 x 12 nodes "
                           "from theta = 4, 4.8 near stations per patch, 512 steps) -- `python tools/idw_far_perf.py 4000 4000 200 512 10 4000`",
     {"kernel": "idw_far_kernel", "shape": "cells16000000_T512_S200"}),
    ("r02_grid_fold.ncu-rep", "grid_local_fold_staged_kernel, GRID grid_local (folded form, vertex values staged per CTA) on the C5 grid "
                              "(4000 x 4000 cells, 4900 source points, 256 steps) -- `python tools/grid_perf.py 4000 4000 256`",
     {"kernel": "grid_local_fold_staged_kernel", "shape": "cells16000000_T256_P4900"}),
    ("r02_far_node.ncu-rep", "idw_far_node_mma_kernel, the node sums of one IDW field on a C4 slab (632 patches x 100 nodes x 200 stations x 1440 "
                             "columns, DMMA) -- `python tools/idw_far_perf.py 1024 10000 200 720`", None),
    ("r02_dk_mma.ncu-rep", "dk_distribute_mma_kernel, DK steady state on a C4 row band (300 x 10000 cells, 720 steps) "
                           "-- `python tools/dk_steady_perf.py`",
     {"kernel": "dk_distribute_mma_kernel", "shape": "cells3000000_T720_S200"}),
    ("r02_dk_steady.ncu-rep", "dk_distribute_kernel (the ELL kernel the tensor-core kernel replaced on meshes), same launch",
     {"kernel": "dk_distribute_kernel", "shape": "cells3000000_T720_S200"}),
    ("r02_dk_sets.ncu-rep", "dk_sets_kernel, the one-off DK weight build (64 x 1024 cells of the C4 basin, 200 stations) "
                            "-- `python tools/dk_perf.py 64 1024 200`", None),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None
    return rows[0], rows[1], rows[2]


def launches(wl, md):
    src = os.path.join(GO, "r02_launches_%s.csv" % wl)
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PR, "r02_launches_%s.csv" % wl))
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("smrfb::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    steady = {k: v for k, v in agg.items() if not k.startswith(("dk_chain", "dk_build", "dk_final", "dk_sets"))}
    tot = sum(v[1] for v in steady.values())
    md.append("\n### %s: `ncu --metrics gpu__time_duration.sum --clock-control none` around `python bench.py%s --steps 1 --warmup 3 "
              "--no-cpu-baseline --no-e2e --no-extras`\n" % (wl.upper(), "" if wl == "c4" else " --workload c5"))
    md.append("%d launches.  One-off DK weight build (outside the timed region): %s.\n" % (
        sum(v[0] for v in agg.values()),
        ", ".join("%s %.1f ms" % (k, v[1] / 1e6) for k, v in agg.items() if k not in steady) or "none"))
    md.append("| kernel | launches | total ms | share of the distribution kernels | ms / launch |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(steady.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 5e-5:
            continue
        md.append("| `%s` | %d | %.2f | %.1f %% | %.3f |" % (k, n, t / 1e6, 100 * t / tot, t / n / 1e6))


def main():
    os.makedirs(PR, exist_ok=True)
    md = ["# Round 2: ncu launch lists of the bench commands\n",
          "Per-launch times under ncu are serialised and cold-cache; what has to agree with the bench line is each kernel's share."]
    for wl in ("c4", "c5"):
        launches(wl, md)
    open(os.path.join(PR, "r02_launches.md"), "w").write("\n".join(md) + "\n")

    md = ["# Round 2: `ncu --set full --clock-control none` captures of the dominant kernels\n"]
    traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the ncu --set full captures named in "
                       "profiles/r02_ncu.md; bench.py looks its roofline.traffic up here by kernel and launch shape", "captures": []}
    for rep, what, tkey in CAPTURES:
        p = os.path.join(GO, rep)
        if not os.path.exists(p):
            continue
        r = raw(p)
        if r is None:
            continue
        hdr, units, vals = r
        md.append("\n## %s\n\n%s\n\n| counter | value |\n|---|---|" % (rep, what))
        got = {}
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                got[key] = (vals[i], units[i])
                md.append("| %s (`%s`) | %s %s |" % (label, key, vals[i], units[i]))
        if tkey and "dram__bytes_read.sum" in got:
            def tobytes(v, u):
                f = float(v.replace(",", ""))
                return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
            e = dict(tkey)
            e["dram_bytes_read"] = tobytes(*got["dram__bytes_read.sum"])
            e["dram_bytes_write"] = tobytes(*got["dram__bytes_write.sum"])
            e["duration_ms"] = float(got["gpu__time_duration.sum"][0].replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[got["gpu__time_duration.sum"][1]]
            e["source"] = "gpurun_out/" + rep
            traffic["captures"].append(e)
    open(os.path.join(PR, "r02_ncu.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(os.path.join(PR, "r02_traffic.json"), "w"), indent=1)
    print("wrote profiles/r02_launches.md, r02_ncu.md, r02_traffic.json")


if __name__ == "__main__":
    main()
