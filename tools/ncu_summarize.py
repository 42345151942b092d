"""Summarise `ncu --set full` reports into profiles/ncu_summary.json: per kernel (short name) the launch time, DRAM bytes per
launch, tensor-pipe and issue activity, registers, grid.  Usage: python tools/ncu_summarize.py <source-tag> rep1.ncu-rep [rep2 ...]
(run in the authoring container: ncu -i needs no GPU)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHORT = [("pool_fwd_kernel", "pool_patch_fwd"), ("pool_bwd_kernel", "pool_patch_bwd"), ("walk_pairs_fwd", "walk_pairs_fwd"),
         ("walk_chain_cluster", "walk_chain_cluster"), ("walk_pairs_bwd", "walk_pairs_bwd"), ("gemm_tf32_kernel", "gemm_tf32"),
         ("splitk_reduce", "splitk_reduce"), ("lp_topk_tc_kernel<(int)0>", "lp_topk_tc_pre"), ("lp_topk_tc_kernel<(int)1>", "lp_topk_tc_exact"),
         ("lp_rescore_kernel<(int)1>", "lp_rescore_check"), ("lp_rescore_kernel<(int)0>", "lp_rescore_listed"), ("lp_split_kernel", "lp_split"),
         ("lp_prepare_kernel", "lp_prepare"), ("lp_gather_kernel", "lp_gather"), ("segmean_accum_tma", "segmean_accum_tma"), ("segmean_mma_kernel", "segmean_mma"), ("segdil_runs", "segdil_runs"),
         ("slic_assign", "slic_assign"), ("slic_connect", "slic_connect"), ("slic_features", "slic_features"), ("slic_minmax", "slic_minmax"),
         ("segmean_count", "segmean_count"), ("segmean_csr", "segmean_csr"), ("segmean_bwd", "segmean_bwd"), ("segdil_count", "segdil_count"),
         ("gemm_tc_kernel", "gemm_tc"), ("tc_split_kernel", "tc_split"), ("patch_grid_kernel", "patch_grid"), ("lp_post_kernel", "lp_post")]
WANT = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "launch__registers_per_thread": "registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__grid_size": "grid", "launch__block_size": "block",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "smsp__inst_executed.sum": "warp_instructions"}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def short(name):
    name = name.replace("(int)", "")
    for k, v in SHORT:
        if k.replace("(int)", "") in name:
            return v
    return re.sub(r"\(.*", "", name).split("::")[-1][:40]


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    out = {"_source": tag}
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = short(d["Kernel Name"])
            ent = {"kernel": d["Kernel Name"][:120], "report": os.path.basename(rep)}
            for k, v in WANT.items():
                if k in d and d[k] != "":
                    val = float(d[k].replace(",", ""))
                    ent[v] = val * UNIT.get(u[k], 1.0) if v in ("us", "dram_read", "dram_write") else val
            if "dram_read" in ent:
                ent["dram_bytes_per_launch"] = ent.pop("dram_read") + ent.pop("dram_write", 0.0)
            if name in out and out[name].get("report") == ent["report"]:      # several launches in one report: keep the longest, count them
                out[name]["launches_captured"] = out[name].get("launches_captured", 1) + 1
                if ent.get("us", 0) <= out[name].get("us", 0):
                    continue
                ent["launches_captured"] = out[name]["launches_captured"]
            # (a kernel that also appears in an EARLIER report is replaced: reports are listed oldest first)
            out[name] = ent
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in sorted(out.items()):
        if k != "_source":
            print("%-22s %9.1f us  dram %8.1f MB  tensor %5.1f%%  issue %5.1f%%  regs %s" % (
                k, v.get("us", 0), v.get("dram_bytes_per_launch", 0) / 1e6, v.get("tensor_pipe_active_pct", 0), v.get("issue_active_pct", 0), v.get("registers")))


if __name__ == "__main__":
    main()
