"""profiles/r2_ncu_traffic.json from `ncu --set full` captures: DRAM bytes (read + write) per launch of each kernel, keyed by
the C entry point bench.py reports.     python tools/ncu_traffic.py gpurun_out/a.ncu-rep [b.ncu-rep ...]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
ENTRY = {"mlp_chain_kernel": "nmrf_mlp_chain", "token_gemm_ra_kernel": "nmrf_token_gemm", "token_gemm_tc6_kernel<0, 0, 1>": "nmrf_conv2d",
         "stripe_attention_tc_kernel": "nmrf_stripe_attention", "window_attention_mma_kernel": "nmrf_window_attention",
         "proposal_attention_kernel": "nmrf_proposal_attention", "cost_volume_topk_kernel": "nmrf_cost_volume_topk"}
acc = {}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    for r in rows[2:]:
        name = next((v for k, v in ENTRY.items() if k in r[ix["Kernel Name"]]), None)
        if name is None:
            continue
        b = sum(float(r[ix[m]]) * UNIT[u[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(r[ix["gpu__time_duration.sum"]])
        a = acc.setdefault(name, dict(bytes=0.0, n=0, us=0.0, src=os.path.basename(rep)))
        a["bytes"] += b; a["n"] += 1; a["us"] += t
res = {k: {"dram_bytes_per_launch": round(a["bytes"] / a["n"]), "launches_captured": a["n"], "avg_us_under_ncu": round(a["us"] / a["n"], 1),
           "source": f"ncu --set full --clock-control none, {a['src']} (cold L2: every launch re-reads its inputs from HBM)"} for k, a in acc.items()}
json.dump(res, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
