# A/B inside one gpurun call (same box, same clocks).  usage: bash scripts/gpu_ab.sh "<env for A>" "<env for B>"
A_ENV=${1:-SC_FUSED_LN=0}
B_ENV=${2:-SC_FUSED_LN=1}
for i in 1 2; do
env $A_ENV python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_A$i.json 2>/dev/null
env $B_ENV python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ab_B$i.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("A1","B1","A2","B2"):
    d=json.load(open(f"gpurun_out/ab_{n}.json"))
    print(n, round(d["value"]), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]))
PY
