# A/B inside one gpurun call (same box, same clocks): A = SC_FUSED_MLP=0 (two-GEMM MLP backward), B = fused MLP dgrad
set -x
for i in 1 2; do
SC_FUSED_MLP=0 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_A$i.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_B$i.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("A1","B1","A2","B2"):
    d=json.load(open(f"gpurun_out/ab_{n}.json"))
    print(n, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["roofline"]["frac"])
PY
