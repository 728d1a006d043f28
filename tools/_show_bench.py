import json,sys
m=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/bench_is{m}.json").read().strip().splitlines()[-1])
    print("item_sort %s: value %.1f M, e2e %.1f M, ms/step %.2f, launches %s" % (m, d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d.get("gpu_launches")), d.get("kernel_share"), d.get("parity_vs_oracle"))
except Exception as e:
    print(m, "failed", e)
