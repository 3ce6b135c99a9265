# does a smaller working set (L2-resident producer->consumer chains) beat the 4096-sample step per sample?
for mb in 296 592 1184 2368 4096; do
python bench.py --steps 10 --warmup 3 --no-cpu --minibatch $mb > gpurun_out/mb_$mb.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/mb_$mb.json"))
print($mb, round(d["value"]), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], "e2e", round(d["e2e"]["value"]))
PY
done
