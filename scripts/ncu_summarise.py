"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the JSON summaries kept under profiles/.

usage: python scripts/ncu_summarise.py raw.csv > summary.json     (one object per profiled launch)
"""
import csv, json, sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def summarise(path):
    rows = list(csv.reader(open(path)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(head)}

    def get(r, name, scale=1.0):
        i = col.get(name)
        if i is None:
            return None
        v = num(r[i])
        return None if v is None else v * scale

    out = []
    for r in body:
        dur_unit = units[col["gpu__time_duration.sum"]]
        dur = get(r, "gpu__time_duration.sum")
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(dur_unit, 1e-6)

        def bytes_of(name):
            v, u = get(r, name), units[col[name]]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

        stalls = {}
        for name, i in col.items():
            pre, post = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
            if name.startswith(pre) and name.endswith(post) and "not_issued" not in name:
                v = num(r[i])
                if v:
                    stalls[name[len(pre):-len(post)]] = v
        tot = sum(stalls.values()) or 1.0
        top = dict(sorted(((k, round(100 * v / tot, 2)) for k, v in stalls.items()), key=lambda kv: -kv[1])[:7])
        rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
        out.append({
            "kernel": r[col["Kernel Name"]],
            "grid": r[col["Grid Size"]],
            "block": r[col["Block Size"]],
            "duration_ms": dur_ms,
            "registers_per_thread": get(r, "launch__registers_per_thread"),
            "dram_read_bytes": rd,
            "dram_write_bytes": wr,
            "dram_bytes_per_launch": rd + wr,
            "dram_throughput_pct": (get(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") or 0) + (get(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed") or 0),
            "fmaheavy_pipe_active_pct_elapsed": get(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "alu_pipe_active_pct_elapsed": get(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "inst_executed": get(r, "smsp__inst_executed.sum"),
            "local_load_inst": get(r, "smsp__sass_inst_executed_op_local_ld.sum"),
            "local_store_inst": get(r, "smsp__sass_inst_executed_op_local_st.sum"),
            "ldgsts_inst": get(r, "smsp__inst_executed_op_ldgsts.sum"),
            "register_spill_inst": get(r, "sass__inst_executed_register_spilling"),
            "shared_bank_conflicts_ld": get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"),
            "shared_bank_conflicts_st": get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum"),
            "top_stalls_pct": top,
        })
    return out


if __name__ == "__main__":
    json.dump(summarise(sys.argv[1]), sys.stdout, indent=1)
    print()
