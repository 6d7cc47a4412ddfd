set -x
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/f_tests.log 2>&1; tail -3 gpurun_out/f_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/f_bench.log 2>&1; tail -1 gpurun_out/f_bench.log | cut -c1-300
timeout 300 python tools/bench_ct.py --profile --cpu > gpurun_out/f_ct.log 2>&1; tail -1 gpurun_out/f_ct.log | cut -c1-250
timeout 300 python tools/bench_uamt.py --profile > gpurun_out/f_uamt.log 2>&1; tail -1 gpurun_out/f_uamt.log | cut -c1-250
timeout 400 python tools/bench_unetr.py --profile --cpu > gpurun_out/f_unetr.log 2>&1; tail -1 gpurun_out/f_unetr.log | cut -c1-250
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/f_ct_launches.csv python tools/bench_ct.py --steps 1 --warmup 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/f_unetr_launches.csv python tools/bench_unetr.py --steps 1 --warmup 1 > /dev/null 2>&1
