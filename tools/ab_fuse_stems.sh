for w in i3d64_ens r3d34_ens twostream64_ens i3d20_ens; do
  for f in true false; do
    CSE_LOWER_KW="{\"fuse_stems\": $f}" timeout 300 python bench.py --workload $w --no-workloads --no-cpu-baseline --steps 10 --warmup 4 > gpurun_out/ab_${w}_$f.json 2> gpurun_out/ab_${w}_$f.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${w}_$f.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("$w fuse=$f", "value %.1f"%d["value"], "ms %.2f"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], "frac %.3f"%r["frac"], [(o["op"],o["ms"],o["tflops"]) for o in r.get("top_ops",[])[:2]])
PY
  done
done
