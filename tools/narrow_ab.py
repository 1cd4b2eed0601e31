"""Per-class serialised times for a list of BVG_DBG values is not possible in one process (env read once), so this
prints one line per invocation: class times + a few stage-5 per-launch times."""
import os, sys, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "per_launch.py"), "0"], capture_output=True, text=True).stderr
v = [(int(m[1]), float(m[2])) for m in re.findall(r"bvg_prof (\d+) cls (\d+) us ([\d.]+)", out)]
cls = [0.0] * 4
for c, us in v: cls[c] += us
s5 = v[2 + 5 * 19 + 1:2 + 5 * 19 + 19]
print("DBG", os.environ.get("BVG_DBG"), "RINGS", os.environ.get("BVG_RINGS"), "ms by class", [round(x / 1e3, 2) for x in cls],
      "s5 k3 A/B", s5[0][1], s5[1][1], "k11 A/B", s5[12][1], s5[13][1])
