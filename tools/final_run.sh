# regenerates the evidence of a round in one gpurun call: tests, smoke, the four bench lines, launch lists
set -x
R=${1:-r2}
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/${R}_tests.log 2>&1; tail -3 gpurun_out/${R}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log | cut -c1-300
for c in 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/${R}_bench_config$c.json 2> gpurun_out/${R}_bench_config$c.err; tail -1 gpurun_out/${R}_bench_config$c.json | cut -c1-250
done
for c in 2 3 4 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${R}_launches_config$c.csv python bench.py --config $c --steps 1 --warmup 1 --no-graph --no-cpu > /dev/null 2>&1
done
