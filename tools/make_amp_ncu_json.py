"""profiles/r02_amp_ncu.json from one evidence run (tools/gpu_evidence.sh <tag>): the DRAM traffic per AMP launch of the
ncu launch list and the pipe activity of the ncu --set full captures, stamped with the commit they were taken at.
Usage: python tools/make_amp_ncu_json.py <tag> <commit> [out.json]"""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, commit = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "r02_amp_ncu.json")
G = os.path.join(ROOT, "gpurun_out")
tab = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_table.py"), os.path.join(G, f"{tag}_launches_time_dram.csv")],
                     capture_output=True, text=True, check=True).stdout
open(os.path.join(ROOT, "profiles", f"r02_launch_table_{commit}_time_dram.txt"), "w").write(
    f"# ncu launch list (gpu__time_duration.sum, dram__bytes_*.sum, --clock-control none) of one 16 x 10 s bf16 decode at commit {commit}\n" + tab)
j = json.loads(tab.strip().splitlines()[-1])
def metric(name, key):
    p = os.path.join(G, f"{tag}_ncu_{name}.txt")
    if not os.path.exists(p): return None
    m = re.search(re.escape(key) + r" = ([\d.]+)", open(p).read())
    return float(m.group(1)) if m else None
wide = {n: metric(n, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for n in ("s0k3A", "s0k7A", "s0k11A", "s1k3A")}
narrow_issue = {n: metric(n, "smsp__issue_active.avg.pct_of_peak_sustained_active") for n in ("s3k7A", "s5k3B", "s5k11B")}
narrow_dram = {n: metric(n, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") for n in ("s3k7A", "s5k3B", "s5k11B")}
res = {"commit": commit,
       "source": f"profiles/r02_launch_table_{commit}_time_dram.txt (ncu launch list) + profiles/r02_ncu_{commit}_*.txt (ncu --set full), tools/gpu_evidence.sh",
       "amp_dram_bytes_per_launch": j["amp_dram_bytes_per_launch"], "amp_alg_bytes_per_launch": j["amp_alg_bytes_per_launch"],
       "amp_flops_per_launch": j["amp_flops_per_launch"], "amp_time_us_under_ncu": j["amp_time_us"],
       "wide_tensor_pipe_active": wide, "narrow_issue_active": narrow_issue, "narrow_dram_throughput_pct": narrow_dram}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
