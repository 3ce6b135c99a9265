set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v6_n1.json 2> gpurun_out/bench_v6.err; echo rc=$?
cat gpurun_out/bench_v6_n1.json | head -c 1500
