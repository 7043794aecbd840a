"""One-screen view of a bench.py JSON line:  python tools/show_bench.py file_with_the_line"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
out = [f"hot ms {round(d['hot_path']['ms_per_step'], 3)} full {round(d['ms_per_step'], 3)} e2e {round(d['e2e']['value'], 2)}"]
for k, v in d["hot_path"]["kernels"].items():
    if v["ms"] > 0.1:
        out.append(f"   {k} {v['ms']} {v['launches']}")
try:
    print("\n".join(out))
except BrokenPipeError:
    pass
