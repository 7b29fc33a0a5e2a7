"""profiles/r2_traffic.json from the ncu summaries of tools/ncu_r2.sh: per-launch DRAM traffic and pipe activity of the step
kernel in the regimes bench.py reports, stamped with the hash of the kernel sources they were captured with (bench.py quotes
them only while that hash matches the tree).  usage: python tools/make_traffic_json.py [dir with r2_*.txt]"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles")


def grab(path):
    t = open(path).read()
    num = lambda key: float(re.search(re.escape(key) + r"\s+([0-9.,]+)", t).group(1).replace(",", ""))
    unit = lambda key: re.search(re.escape(key) + r"\s+[0-9.,]+\s+(\S+)", t).group(1)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = num("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
    wr = num("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    return {"kernel": re.search(r"Kernel Name\s+(.*?)\s*$", t, re.M).group(1).strip(), "dram_bytes_read": rd, "dram_bytes_write": wr,
            "dram_bytes_per_launch": rd + wr,
            "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "local_loads_executed": num("sass__inst_executed_local_loads"), "local_stores_executed": num("sass__inst_executed_local_stores"),
            "source": "profiles/" + os.path.basename(path) + " (ncu --set full --clock-control none, one launch)"}


out = {"csrc_sha16": open(os.path.join(src, "r2_csrc_sha16.txt")).read().strip(),
       "graded_fp64_B4096": grab(os.path.join(src, "r2_lat_graded_B4096.txt")),
       "fixed_fp64_B4096": grab(os.path.join(src, "r2_lat_fixed600_B4096.txt")),
       "graded_fp64_B262144": grab(os.path.join(src, "r2_tput_graded_B262144.txt")),
       "note": "at B = 4096 the outputs (4.3 MB of observations + state) stay in the 126 MB L2 inside ncu's replay: DRAM writes read 0; "
               "algorithmic bytes are 7.2 MB per launch"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:1500])
